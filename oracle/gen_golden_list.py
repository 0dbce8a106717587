"""Golden vectors for QaryPolarEncoderDecoder.listDecode from the LIVE reference (build container only).

Run through `python oracle/gen_golden.py list`.  Inputs are continuous-valued (BI-AWGN / jittered QSC) so the
metrics are tie-free and the reference's numpy-implementation-defined candidate order (see
oracle/polar_oracle_list.c header) is observable: on this AVX-512 host it is ascending by metric.
"""
import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def bec_z_order(n, eps=0.5):
    z = [eps]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return np.argsort(-np.array(z), kind="stable")


def main(ref):
    out = {}
    names = []

    def run_case(name, q, n, frozenSet, L, kind, frames, rng_seed, frozen_random=False):
        N = 1 << n
        rng = random.Random(rng_seed)
        nrng = np.random.default_rng(rng_seed)
        ed = ref.QPED.QaryPolarEncoderDecoder(q, N, set(frozenSet), 1)
        k = ed.k
        recs = {key: [] for key in ("xy", "fv", "ainfo", "info", "pr", "lsize", "linfo", "lprob", "aprob")}
        for f in range(frames):
            info = np.array([rng.randrange(q) for _ in range(k)], dtype=np.int64)
            fv = np.array([rng.randrange(q) if frozen_random else 0 for _ in range(N - k)], dtype=np.int64)
            u = ed.mergeInfoAndFrozen(info, fv)
            cw = np.asarray(ref.QPED.polarTransformOfQudits(q, u), dtype=np.int64)
            if kind == "awgn":  # q == 2
                sigma = 0.85
                y = (1.0 - 2.0 * cw) + sigma * nrng.standard_normal(N)
                l0, l1 = -(y - 1) ** 2 / (2 * sigma ** 2), -(y + 1) ** 2 / (2 * sigma ** 2)
                m = np.maximum(l0, l1)
                xy = np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=1)
            else:  # jittered q-ary symmetric channel: continuous values, no exact ties
                p = 0.12
                xy = np.empty((N, q))
                for i in range(N):
                    yv = cw[i] if rng.random() > p else (cw[i] + rng.randrange(1, q)) % q
                    row = np.array([1.0 - p if x == yv else p / (q - 1) for x in range(q)])
                    xy[i] = row * (1.0 + 0.2 * nrng.random(q))
            vd = ref.QMVD.QaryMemorylessVectorDistribution(q, N)
            vd.probs[:] = xy
            res, pr = ed.listDecode(vd, fv, L, np.zeros((k, 0), dtype=np.int64), np.zeros(0, dtype=np.int64),
                                    actualInformation=info)
            # the final list is still in ed.prob_list / the returned info list is internal: recompute via a second call
            # that exposes recursiveListDecode's outputs
            ed.actualInformation = info
            ed.actual_prob = 1.0
            ed.prob_list = np.array([1.0])
            ed.info_time = ed.transform_time = ed.encoding_time = 0
            il = np.full((L * q, k), -1, dtype=np.int64)
            it = np.nditer(fv, flags=['f_index']) if len(fv) else None
            (il, encl, nu, ni, fsize, omap, aenc) = ed.recursiveListDecode(il, 0, 0, [vd], it, inListSize=1, maxListSize=L,
                                                                           actualXyVectorDistribution=vd)
            lin = np.full((L, k), -1, dtype=np.int64)
            lin[:fsize] = il[:fsize]
            lpr = np.zeros(L)
            lpr[:fsize] = ed.prob_list
            recs["xy"].append(xy), recs["fv"].append(fv), recs["ainfo"].append(info)
            recs["info"].append(np.asarray(res, dtype=np.int64)), recs["pr"].append(pr.value)
            recs["lsize"].append(fsize), recs["linfo"].append(lin), recs["lprob"].append(lpr)
            recs["aprob"].append(ed.actual_prob)
        fm = np.zeros(N, dtype=np.uint8)
        if len(frozenSet):
            fm[list(frozenSet)] = 1
        out[name + "/q"], out[name + "/n"], out[name + "/L"] = np.int64(q), np.int64(n), np.int64(L)
        out[name + "/frozen"] = fm
        for key, dt in (("xy", np.float64), ("fv", np.int64), ("ainfo", np.int64), ("info", np.int64), ("pr", np.int64),
                        ("lsize", np.int64), ("linfo", np.int64), ("lprob", np.float64), ("aprob", np.float64)):
            out[name + "/" + key] = np.array(recs[key], dtype=dt)
        names.append(name)
        print(name, "results", recs["pr"], flush=True)

    for n in (1, 2, 3, 4, 5, 6, 8):
        N = 1 << n
        for L in ((1, 2, 4, 8) if n in (4, 6) else (4, 8)):
            fs = set(int(i) for i in bec_z_order(n)[:N // 2])
            run_case("q2_n%d_L%d_awgn" % (n, L), 2, n, fs, L, "awgn", 6 if n <= 6 else 3, 1000 + 10 * n + L)
    rr = random.Random(5)
    run_case("q2_n6_L8_randfrozen", 2, 6, set(rr.sample(range(64), 30)), 8, "awgn", 6, 77, frozen_random=True)
    run_case("q2_n7_L16_awgn", 2, 7, set(int(i) for i in bec_z_order(7)[:70]), 16, "awgn", 4, 78, frozen_random=True)
    run_case("q2_n9_L8_awgn", 2, 9, set(int(i) for i in bec_z_order(9)[:256]), 8, "awgn", 2, 79)
    for n in (2, 4, 6):
        N = 1 << n
        fs = set(int(i) for i in bec_z_order(n)[:N // 2])
        run_case("q3_n%d_L4_jqsc" % n, 3, n, fs, 4, "jqsc", 5, 300 + n, frozen_random=True)
    run_case("q3_n7_L8_jqsc", 3, 7, set(int(i) for i in bec_z_order(7)[:64]), 8, "jqsc", 3, 310)
    run_case("q5_n5_L4_jqsc", 5, 5, set(int(i) for i in bec_z_order(5)[:16]), 4, "jqsc", 3, 311, frozen_random=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "scl.npz"), **out)
    print("wrote scl.npz with", len(names), "cases")

"""Golden vectors for the LOG DOMAIN (`use_log=True`) of QaryPolarEncoderDecoder.decode and .listDecode from the LIVE
reference (build container only): python oracle/gen_golden.py log -> tests/golden/qlog.npz

The inputs are natural logarithms of continuous-valued (tie-free) channel probabilities, as
makeQaryMemorylessVectorDistribution(..., use_log=True) builds them (math.log, -inf for 0:
ScalarDistributions/QaryMemorylessDistribution.py:757-776).  The log branches go through numpy.logaddexp and
scipy.special.logsumexp, whose exp / log1p are the host libm's: the oracle and the kernels reproduce DECISIONS exactly on
these vectors and the float64 metrics to a stated tolerance (tests/test_oracle_golden_log.py, tests/test_gpu_log.py).
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


def _channel(q, cw, kind, rng, nrng, zeros=False):
    N = len(cw)
    if kind == "awgn":  # q == 2
        sigma = 0.85
        y = (1.0 - 2.0 * cw) + sigma * nrng.standard_normal(N)
        l0, l1 = -(y - 1) ** 2 / (2 * sigma ** 2), -(y + 1) ** 2 / (2 * sigma ** 2)
        m = np.maximum(l0, l1)
        return np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=1)
    p = 0.12
    xy = np.empty((N, q))
    for i in range(N):
        yv = cw[i] if rng.random() > p else (cw[i] + rng.randrange(1, q)) % q
        row = np.array([1.0 - p if x == yv else p / (q - 1) for x in range(q)])
        xy[i] = row * (1.0 + 0.2 * nrng.random(q))
        if zeros and rng.random() < 0.15:  # an impossible symbol: probability 0 -> -inf
            xy[i, (yv + 1) % q] = 0.0
    return xy


def _log(xy):
    with np.errstate(divide="ignore"):
        out = np.where(xy != 0, np.log(np.where(xy != 0, xy, 1.0)), -np.inf)
    # math.log, element by element, is what the reference uses; numpy's vector log may differ in the last bit
    import math
    flat = out.reshape(-1)
    src = xy.reshape(-1)
    for i in range(flat.shape[0]):
        flat[i] = math.log(src[i]) if src[i] != 0 else -math.inf
    return out


def main(ref):
    out = {}
    names = []

    def sc_case(name, q, n, frozenSet, kind, frames, seed, zeros=False):
        N = 1 << n
        rng, nrng = random.Random(seed), np.random.default_rng(seed)
        ed = ref.QPED.QaryPolarEncoderDecoder(q, N, set(frozenSet), 1, use_log=True)
        enc = ref.QPED.QaryPolarEncoderDecoder(q, N, set(frozenSet), 1)
        k = ed.k
        xv = ref.QMVD.QaryMemorylessVectorDistribution(q, N)
        xv.probs[:] = 1.0 / q
        xvl = ref.QMVD.QaryMemorylessVectorDistribution(q, N, use_log=True)
        xvl.probs[:] = -np.log(q)
        recs = {"xyl": [], "tx": [], "info": []}
        for _ in range(frames):
            info = [rng.randrange(q) for _ in range(k)]
            cw = np.asarray(enc.encode(xv, info), dtype=np.int64)
            xyl = _log(_channel(q, cw, kind, rng, nrng, zeros))
            vd = ref.QMVD.QaryMemorylessVectorDistribution(q, N, use_log=True)
            vd.probs[:] = xyl
            dec = ed.decode(xvl, vd)
            recs["xyl"].append(xyl), recs["tx"].append(np.array(info, dtype=np.int64))
            recs["info"].append(np.asarray(dec, dtype=np.int64))
        fm = np.zeros(N, dtype=np.uint8)
        if len(frozenSet):
            fm[list(frozenSet)] = 1
        out[name + "/kind"] = np.array("sc")
        out[name + "/q"], out[name + "/n"], out[name + "/frozen"] = np.int64(q), np.int64(n), fm
        out[name + "/xyl"] = np.array(recs["xyl"])
        out[name + "/tx"] = np.array(recs["tx"], dtype=np.int64).reshape(frames, k)
        out[name + "/info"] = np.array(recs["info"], dtype=np.int64).reshape(frames, k)
        names.append(name)
        print(name, "symbol errors", int((out[name + "/tx"] != out[name + "/info"]).sum()), flush=True)

    def list_case(name, q, n, frozenSet, L, kind, frames, seed, frozen_random=False):
        N = 1 << n
        rng, nrng = random.Random(seed), np.random.default_rng(seed)
        ed = ref.QPED.QaryPolarEncoderDecoder(q, N, set(frozenSet), 1, use_log=True)
        k = ed.k
        recs = {key: [] for key in ("xyl", "fv", "ainfo", "info", "pr", "lsize", "linfo", "lprob", "aprob")}
        for _ in range(frames):
            info = np.array([rng.randrange(q) for _ in range(k)], dtype=np.int64)
            fv = np.array([rng.randrange(q) if frozen_random else 0 for _ in range(N - k)], dtype=np.int64)
            u = ed.mergeInfoAndFrozen(info, fv)
            cw = np.asarray(ref.QPED.polarTransformOfQudits(q, u), dtype=np.int64)
            xyl = _log(_channel(q, cw, kind, rng, nrng))
            vd = ref.QMVD.QaryMemorylessVectorDistribution(q, N, use_log=True)
            vd.probs[:] = xyl
            res, pr = ed.listDecode(vd, fv, L, np.zeros((k, 0), dtype=np.int64), np.zeros(0, dtype=np.int64),
                                    actualInformation=info)
            ed.actualInformation = info
            ed.actual_prob = 0.0
            ed.prob_list = np.array([0.0])
            ed.info_time = ed.transform_time = ed.encoding_time = 0
            il = np.full((L * q, k), -1, dtype=np.int64)
            it = np.nditer(fv, flags=['f_index']) if len(fv) else None
            (il, encl, nu, ni, fsize, omap, aenc) = ed.recursiveListDecode(il, 0, 0, [vd], it, inListSize=1, maxListSize=L,
                                                                           actualXyVectorDistribution=vd)
            lin = np.full((L, k), -1, dtype=np.int64)
            lin[:fsize] = il[:fsize]
            lpr = np.full(L, -np.inf)
            lpr[:fsize] = ed.prob_list
            recs["xyl"].append(xyl), recs["fv"].append(fv), recs["ainfo"].append(info)
            recs["info"].append(np.asarray(res, dtype=np.int64)), recs["pr"].append(pr.value)
            recs["lsize"].append(fsize), recs["linfo"].append(lin), recs["lprob"].append(lpr)
            recs["aprob"].append(ed.actual_prob)
        fm = np.zeros(N, dtype=np.uint8)
        if len(frozenSet):
            fm[list(frozenSet)] = 1
        out[name + "/kind"] = np.array("list")
        out[name + "/q"], out[name + "/n"], out[name + "/L"] = np.int64(q), np.int64(n), np.int64(L)
        out[name + "/frozen"] = fm
        for key, dt in (("xyl", np.float64), ("fv", np.int64), ("ainfo", np.int64), ("info", np.int64), ("pr", np.int64),
                        ("lsize", np.int64), ("linfo", np.int64), ("lprob", np.float64), ("aprob", np.float64)):
            out[name + "/" + key] = np.array(recs[key], dtype=dt)
        names.append(name)
        print(name, "results", recs["pr"], flush=True)

    half = lambda n: set(int(i) for i in bec_z_order(n)[:(1 << n) // 2])  # noqa: E731
    for n in (1, 3, 6, 8):
        sc_case("sc_q2_n%d_awgn" % n, 2, n, half(n), "awgn", 6, 2000 + n)
    for n in (2, 4, 6, 7):
        sc_case("sc_q3_n%d_jqsc" % n, 3, n, half(n), "jqsc", 5, 2100 + n)
    sc_case("sc_q3_n5_zeros", 3, 5, half(5), "jqsc", 6, 2150, zeros=True)
    sc_case("sc_q5_n4_jqsc", 5, 4, half(4), "jqsc", 4, 2160)
    sc_case("sc_q4_n5_jqsc", 4, 5, half(5), "jqsc", 4, 2161)
    sc_case("sc_q3_n9_jqsc", 3, 9, half(9), "jqsc", 2, 2170)
    for n, Ls in ((2, (4,)), (4, (1, 2, 4, 8)), (6, (4, 8)), (8, (8,))):
        for L in Ls:
            list_case("list_q2_n%d_L%d_awgn" % (n, L), 2, n, half(n), L, "awgn", 5 if n <= 6 else 3, 2200 + 10 * n + L)
    rr = random.Random(6)
    list_case("list_q2_n6_L8_randfrozen", 2, 6, set(rr.sample(range(64), 30)), 8, "awgn", 5, 2277, frozen_random=True)
    list_case("list_q2_n7_L16_awgn", 2, 7, set(int(i) for i in bec_z_order(7)[:70]), 16, "awgn", 3, 2278, frozen_random=True)
    for n in (2, 4, 6):
        list_case("list_q3_n%d_L4_jqsc" % n, 3, n, half(n), 4, "jqsc", 4, 2300 + n, frozen_random=True)
    list_case("list_q3_n7_L8_jqsc", 3, 7, set(int(i) for i in bec_z_order(7)[:64]), 8, "jqsc", 2, 2310)
    list_case("list_q5_n5_L4_jqsc", 5, 5, set(int(i) for i in bec_z_order(5)[:16]), 4, "jqsc", 2, 2311, frozen_random=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "qlog.npz"), **out)
    print("wrote qlog.npz with", len(names), "cases")

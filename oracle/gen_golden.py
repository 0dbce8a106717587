"""Generate golden input/output vectors by running the LIVE reference (build container only).

TEST INFRASTRUCTURE.  Needs /root/reference (read-only) and the shim in oracle/refshim.py.  The outputs
(tests/golden/*.npz, small) are committed together with this script; nothing on the GPU box reads the
reference.

  python oracle/gen_golden.py sc        -> tests/golden/sc_binary.npz, tests/golden/sc_qary.npz
  python oracle/gen_golden.py list      -> tests/golden/scl.npz
  python oracle/gen_golden.py log       -> tests/golden/qlog.npz (use_log=True)
  python oracle/gen_golden.py trellis   -> tests/golden/trellis.npz
  python oracle/gen_golden.py c1        -> tests/golden/c1_n1024.npz   (needs constructions/bsc_p0.11_n10_L100_pe.npy)

Every case stores the exact inputs (frozen mask, common-randomness seed, prior, channel probabilities,
information) and what the reference returned (codewords, information, first-level probabilities, leaf
marginals).
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------------------------------------------
def bec_z_order(n, eps=0.5):
    """Heuristic reliability order (BEC Bhattacharyya, MSB-first minus/plus = reference u order)."""
    z = [eps]
    for _ in range(n):
        nz = []
        for v in z:
            nz.append(2 * v - v * v)
            nz.append(v * v)
        z = nz
    return np.argsort(-np.array(z), kind="stable")  # worst first


def frozen_from_order(n, k):
    N = 1 << n
    order = bec_z_order(n)
    return set(int(i) for i in order[:N - k])


def bmvd(ref, probs):
    v = ref.BMVD.BinaryMemorylessVectorDistribution(probs.shape[0])
    v.probs[:] = probs
    return v


def qmvd(ref, q, probs, use_log=False):
    v = ref.QMVD.QaryMemorylessVectorDistribution(q, probs.shape[0], use_log=use_log)
    v.probs[:] = probs
    return v


def bin_channel_probs(kind, rng, cw, N):
    """Return xy probs [N,2] for one frame given the transmitted codeword."""
    if kind == "bsc":
        p = 0.11
        y = np.array([c ^ (1 if rng.random() < p else 0) for c in cw])
        tab = np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])  # makeBSC, BMD:485-490
        return tab[y]
    if kind == "bsc_hard":  # many errors: exercises saturation / contradictions
        p = 0.25
        y = np.array([c ^ (1 if rng.random() < p else 0) for c in cw])
        tab = np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])
        return tab[y]
    if kind == "bec":
        p = 0.4
        tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])  # makeBEC, BMD:493-499
        y = np.array([2 if rng.random() < p else c for c in cw])
        return tab[y]
    if kind == "bec_lossy":  # erasures AND flips: contradiction states (0,0) appear
        p = 0.3
        tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
        y = np.array([2 if rng.random() < p else (c ^ (1 if rng.random() < 0.05 else 0)) for c in cw])
        return tab[y]
    if kind == "awgn":
        sigma = 0.9
        y = np.array([(1.0 - 2.0 * c) + rng.gauss(0.0, sigma) for c in cw])
        l0 = -(y - 1.0) ** 2 / (2 * sigma * sigma)
        l1 = -(y + 1.0) ** 2 / (2 * sigma * sigma)
        m = np.maximum(l0, l1)
        return np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=1)
    if kind == "random":
        return np.array([[rng.random(), rng.random()] for _ in range(N)])
    raise ValueError(kind)


def gen_sc_binary(ref):
    out = {}
    names = []

    def run_case(name, n, frozenSet, seed, kind, frames, prior=None, rng_seed=0):
        N = 1 << n
        rng = random.Random(rng_seed)
        enc = ref.BPED.BinaryPolarEncoderDecoder(N, set(frozenSet), seed)
        xprobs = np.tile(np.array([[0.5, 0.5]]), (N, 1)) if prior is None else np.asarray(prior, dtype=np.float64)
        infos, cws, xys, dcws, dinfos = [], [], [], [], []
        for f in range(frames):
            info = [0 if rng.random() < 0.5 else 1 for _ in range(enc.k)]
            cw = enc.encode(bmvd(ref, xprobs), info)
            xy = bin_channel_probs(kind, rng, [int(c) for c in cw], N)
            if prior is not None:  # joint P(x,y) = P(x) W(y|x)-like weighting
                xy = xy * xprobs * 2.0
            dcw, dinfo = enc.decode(bmvd(ref, xprobs), bmvd(ref, xy))
            infos.append(info), cws.append(cw), xys.append(xy), dcws.append(dcw), dinfos.append(dinfo)
        # intermediate probabilities of frame 0: first-level children and leaf marginals
        xy0 = bmvd(ref, xys[0])
        marg = []
        information = np.full(enc.k, -1, dtype=np.int64)
        enc.recursiveEncodeDecode(information, 0, 0, enc.randomlyGeneratedNumbers, bmvd(ref, xprobs), xy0, marg)
        if n >= 1:
            mchild = xy0.minusTransform()
            mchild.normalize(mchild.calcNormalizationVector())
            # minus-branch codeword of frame 0 = polar transform structure: take from decoded cw
            dcw0 = np.asarray(dcws[0])
            em = (dcw0[0::2] + dcw0[1::2]) % 2
            pchild = xy0.plusTransform(em)
            pchild.normalize(pchild.calcNormalizationVector())
            out[name + "/lvl1_minus"] = mchild.probs.copy()
            out[name + "/lvl1_plus"] = pchild.probs.copy()
        fm = np.zeros(N, dtype=np.uint8)
        if len(frozenSet):
            fm[list(frozenSet)] = 1
        out[name + "/n"] = np.int64(n)
        out[name + "/frozen"] = fm
        out[name + "/seed"] = np.int64(seed)
        out[name + "/r"] = np.asarray(enc.randomlyGeneratedNumbers, dtype=np.float64)
        out[name + "/xprobs"] = xprobs
        out[name + "/info"] = np.array(infos, dtype=np.int64).reshape(frames, enc.k)
        out[name + "/cw"] = np.array(cws, dtype=np.int64)
        out[name + "/xy"] = np.array(xys, dtype=np.float64)
        out[name + "/dec_cw"] = np.array(dcws, dtype=np.int64)
        out[name + "/dec_info"] = np.array(dinfos, dtype=np.int64).reshape(frames, enc.k)
        out[name + "/marg0"] = np.array(marg, dtype=np.float64)
        names.append(name)
        print(name, "frame errors", sum(int(not np.array_equal(a, b)) for a, b in zip(infos, dinfos)), "/", frames,
              flush=True)

    # SURVEY 8c vector: N=8, frozen {0,1,2,4}
    for seed in (1, -1):
        run_case("n3_survey_seed%d" % seed, 3, {0, 1, 2, 4}, seed, "bsc", 4, rng_seed=11)
    for n in (0, 1, 2, 4, 6, 7, 8):
        N = 1 << n
        for kind in ("bsc", "bsc_hard", "bec", "bec_lossy", "awgn", "random"):
            for seed in ((1, -1) if n <= 6 else (1,)):
                k = N // 2
                fs = frozen_from_order(n, k) if n > 0 else set()
                frames = 6 if n <= 6 else 3
                run_case("n%d_%s_seed%d" % (n, kind, seed), n, fs, seed, kind, frames, rng_seed=100 * n + seed + 5)
    # edge cases: nothing frozen, everything frozen, random frozen sets, odd rates
    rr = random.Random(99)
    run_case("n5_allinfo", 5, set(), 3, "bsc", 3, rng_seed=1)
    run_case("n5_allfrozen", 5, set(range(32)), 3, "bsc", 3, rng_seed=2)
    run_case("n6_randfrozen", 6, set(rr.sample(range(64), 23)), 5, "awgn", 4, rng_seed=3)
    run_case("n7_randfrozen", 7, set(rr.sample(range(128), 77)), -1, "bsc_hard", 3, rng_seed=4)
    run_case("n7_rate_hi", 7, frozen_from_order(7, 112), 2, "bec", 3, rng_seed=5)
    # non-uniform prior (Honda-Yamamoto shaping): frozen bits become data dependent, BPED:258-262
    pr = np.tile(np.array([[0.7, 0.3]]), (64, 1))
    run_case("n6_prior", 6, frozen_from_order(6, 24), 1, "bsc", 4, prior=pr, rng_seed=6)
    pr2 = np.array([[rr.random(), rr.random()] for _ in range(32)])
    run_case("n5_prior_rand", 5, frozen_from_order(5, 12), 4, "awgn", 4, prior=pr2, rng_seed=7)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "sc_binary.npz"), **out)
    print("wrote sc_binary.npz with", len(names), "cases")


def q_channel_probs(q, kind, rng, cw, N):
    if kind == "qsc":
        p = 0.1
        tab = np.array([[1.0 - p if x == y else p / (q - 1) for x in range(q)] for y in range(q)])  # makeQSC QMD:780
        y = []
        for c in cw:
            if rng.random() < p:
                c = (c + rng.randrange(1, q)) % q
            y.append(c)
        return tab[np.array(y)]
    if kind == "qsc_hard":
        p = 0.3
        tab = np.array([[1.0 - p if x == y else p / (q - 1) for x in range(q)] for y in range(q)])
        y = []
        for c in cw:
            if rng.random() < p:
                c = (c + rng.randrange(1, q)) % q
            y.append(c)
        return tab[np.array(y)]
    if kind == "qec":
        p = 0.4
        rows = [[(1.0 - p) / q if x == y else 0.0 for x in range(q)] for y in range(q)] + [[p / q] * q]  # makeQEC
        tab = np.array(rows)
        y = [q if rng.random() < p else c for c in cw]
        return tab[np.array(y)]
    if kind == "random":
        return np.array([[rng.random() for _ in range(q)] for _ in range(N)])
    raise ValueError(kind)


def gen_sc_qary(ref):
    out = {}
    names = []

    def run_case(name, q, n, frozenSet, kind, frames, rng_seed=0):
        N = 1 << n
        rng = random.Random(rng_seed)
        enc = ref.QPED.QaryPolarEncoderDecoder(q, N, set(frozenSet), 1)
        xprobs = np.full((N, q), 1.0 / q)
        infos, cws, xys, dinfos = [], [], [], []
        for f in range(frames):
            info = [rng.randrange(q) for _ in range(enc.k)]
            cw = enc.encode(qmvd(ref, q, xprobs), info)
            xy = q_channel_probs(q, kind, rng, [int(c) for c in cw], N)
            dinfo = enc.decode(qmvd(ref, q, xprobs), qmvd(ref, q, xy))
            infos.append(info), cws.append(cw), xys.append(xy), dinfos.append(dinfo)
        xy0 = qmvd(ref, q, xys[0])
        marg = []
        information = np.full(enc.k, -1, dtype=np.int64)
        dcw0, _, _ = enc.recursiveEncodeDecode(information, 0, 0, qmvd(ref, q, xprobs), xy0, marg)
        if n >= 1:
            mchild = xy0.minusTransform()
            mchild.normalize()
            # minus codeword from the decoded codeword: m = x[2h] + x[2h+1] (since x[2h+1] = -p)
            em = (dcw0[0::2] + dcw0[1::2]) % q
            pchild = xy0.plusTransform(em)
            pchild.normalize()
            out[name + "/lvl1_minus"] = mchild.probs.copy()
            out[name + "/lvl1_plus"] = pchild.probs.copy()
        fm = np.zeros(N, dtype=np.uint8)
        if len(frozenSet):
            fm[list(frozenSet)] = 1
        out[name + "/q"] = np.int64(q)
        out[name + "/n"] = np.int64(n)
        out[name + "/frozen"] = fm
        out[name + "/xprobs"] = xprobs
        out[name + "/info"] = np.array(infos, dtype=np.int64).reshape(frames, enc.k)
        out[name + "/cw"] = np.array(cws, dtype=np.int64)
        out[name + "/xy"] = np.array(xys, dtype=np.float64)
        out[name + "/dec_info"] = np.array(dinfos, dtype=np.int64).reshape(frames, enc.k)
        out[name + "/dec_cw0"] = np.asarray(dcw0, dtype=np.int64)
        out[name + "/marg0"] = np.array([np.asarray(m, dtype=np.float64) for m in marg])
        names.append(name)
        print(name, "frame errors", sum(int(not np.array_equal(a, b)) for a, b in zip(infos, dinfos)), "/", frames,
              flush=True)

    run_case("q3_n3_survey", 3, 3, {0, 1, 2, 4}, "qsc", 4, rng_seed=21)
    for q in (2, 3, 4, 5, 7):
        for n in ((1, 2, 4, 6, 8) if q == 3 else (2, 5, 6)):
            N = 1 << n
            for kind in ("qsc", "qsc_hard", "qec", "random"):
                fs = frozen_from_order(n, N // 2)
                run_case("q%d_n%d_%s" % (q, n, kind), q, n, fs, kind, 3 if n >= 6 else 5, rng_seed=q * 1000 + n * 10)
    run_case("q3_n5_allinfo", 3, 5, set(), "qsc", 3, rng_seed=31)
    run_case("q3_n5_allfrozen", 3, 5, set(range(32)), "qsc", 3, rng_seed=32)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "sc_qary.npz"), **out)
    print("wrote sc_qary.npz with", len(names), "cases")


def main():
    ref = refshim.load()
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else "sc"
    if what == "sc":
        gen_sc_binary(ref)
        gen_sc_qary(ref)
    elif what == "list":
        from oracle import gen_golden_list
        gen_golden_list.main(ref)
    elif what == "log":
        from oracle import gen_golden_log
        gen_golden_log.main(ref)
    elif what == "trellis":
        from oracle import gen_golden_trellis
        gen_golden_trellis.main(ref)
    elif what == "c1":
        from oracle import gen_golden_c1
        gen_golden_c1.main(ref)
    else:
        raise SystemExit("unknown target")


if __name__ == "__main__":
    main()

"""Golden vectors for the Tal-Vardy degrading construction, produced by the LIVE reference.

TEST INFRASTRUCTURE, build container only (needs /root/reference).  Output: tests/golden/tv_construct.npz.  For each case:
the channel table (BinaryMemorylessDistribution.probs), n, L, the Pe vector of
`dist.minusTransform().degrade(L)` / `dist.plusTransform().degrade(L)` per level + `errorProb()` per leaf
(ScalarDistributions/BinaryMemorylessDistribution.py:657-677), the frozen set calcFrozenSet_degradingUpgrading returns for the
stated epsilon, and the symbol table of the first degraded minus channel (so a mismatch can be localised).
The large constructions of the BASELINE configs are under tests/golden/constructions/ (oracle/gen_constructions.py).
"""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tv_construct.npz")


def main():
    ref = refshim.load()
    BMD = ref.BMD
    rng = np.random.default_rng(2024)

    def dist_of(rows):
        d = BMD.BinaryMemorylessDistribution()
        for r in rows:
            d.append([float(r[0]), float(r[1])])
        return d

    def random_channel(Y):
        t = rng.random((Y, 2)) ** 2
        return t / t.sum()

    dup = np.array([[0.2, 0.05], [0.1, 0.025], [0.0, 0.0], [0.05, 0.2], [0.3, 0.075]])  # equal LLRs and a zero-probability output
    cases = [  # name, table, n, L, epsilon
        ("bsc0.2_n5_L8", [[0.4, 0.1], [0.1, 0.4]], 5, 8, 0.3),
        ("bsc0.05_n6_L16", [[0.475, 0.025], [0.025, 0.475]], 6, 16, 0.1),
        ("bec0.3_n6_L4", [[0.35, 0.0], [0.0, 0.35], [0.15, 0.15]], 6, 4, 0.2),
        ("bec0.5_n4_L3", [[0.25, 0.0], [0.0, 0.25], [0.25, 0.25]], 4, 3, 0.5),
        ("rand5_n4_L33", random_channel(5), 4, 33, 0.4),
        ("rand9_n3_L7", random_channel(9), 3, 7, 0.4),
        ("dup_zero_n4_L5", dup / dup.sum(), 4, 5, 0.3),
        ("bsc0.11_n4_L2", [[0.445, 0.055], [0.055, 0.445]], 4, 2, 0.5),
        ("bsc0.11_n3_L1", [[0.445, 0.055], [0.055, 0.445]], 3, 1, 0.5),
        ("n0_L10", [[0.3, 0.1], [0.2, 0.4]], 0, 10, 0.5),
    ]
    out = {"names": np.array([c[0] for c in cases])}
    for nm, tab, n, L, eps in cases:
        tab = np.asarray(tab, dtype=np.float64)
        dists = [dist_of(tab)]
        first = None
        for m in range(1, n + 1):
            nxt = []
            for d in dists:
                nxt.append(d.minusTransform().degrade(L))
                nxt.append(d.plusTransform().degrade(L))
            dists = nxt
            if first is None:
                first = np.array(dists[0].probs, dtype=np.float64)
        pe = np.array([d.errorProb() for d in dists], dtype=np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            fs = BMD.calcFrozenSet_degradingUpgrading(n, L, eps, None, dist_of(tab))
        out[nm + "/table"] = tab
        out[nm + "/params"] = np.array([n, L], dtype=np.int64)
        out[nm + "/eps"] = np.float64(eps)
        out[nm + "/pe"] = pe
        out[nm + "/frozen"] = np.array(sorted(fs), dtype=np.int64)
        out[nm + "/first_minus"] = first if first is not None else np.zeros((0, 2))
        print(nm, "N", 1 << n, "L", L, "frozen", len(fs), "sum Pe", pe.sum())
    # ---- q-ary: QaryMemorylessDistribution.minusTransform().degrade(L) / plusTransform().degrade(L), errorProb() ----------
    QMD = ref.QMD

    def qdist_of(q, rows):
        d = QMD.QaryMemorylessDistribution(q)
        for r in rows:
            d.append([float(v) for v in r])
        return d

    def qsc(q, p):
        return np.array(QMD.makeQSC(q, p).probs, dtype=np.float64)

    def qec(q, p):
        return np.array(QMD.makeQEC(q, p).probs, dtype=np.float64)

    def rand_q(Y, q):
        t = rng.random((Y, q)) ** 2
        return t / t.sum()

    qcases = [  # name, q, table, n, L
        ("q3_qsc0.1_n4_L16", 3, qsc(3, 0.1), 4, 16),
        ("q3_qsc0.02_n5_L100", 3, qsc(3, 0.02), 5, 100),
        ("q3_qec0.3_n4_L9", 3, qec(3, 0.3), 4, 9),
        ("q2_qsc0.11_n5_L8", 2, qsc(2, 0.11), 5, 8),
        ("q4_qsc0.1_n3_L27", 4, qsc(4, 0.1), 3, 27),
        ("q5_qsc0.05_n2_L16", 5, qsc(5, 0.05), 2, 16),
        ("q3_rand4_n3_L25", 3, rand_q(4, 3), 3, 25),
        ("q3_qsc0.1_n3_L1", 3, qsc(3, 0.1), 3, 1),
    ]
    out["qnames"] = np.array([c[0] for c in qcases])
    for nm, q, tab, n, L in qcases:
        dists = [qdist_of(q, tab)]
        for m in range(1, n + 1):
            nxt = []
            for d in dists:
                nxt.append(d.minusTransform().degrade(L))
                nxt.append(d.plusTransform().degrade(L))
            dists = nxt
        pe = np.array([d.errorProb() for d in dists], dtype=np.float64)
        out[nm + "/table"] = np.asarray(tab, dtype=np.float64)
        out[nm + "/params"] = np.array([q, n, L], dtype=np.int64)
        out[nm + "/pe"] = pe
        # QaryPolarEncoderDecoder.frozenSetFromTVAndPe (:1156-1190), both selection rules, uniform input (TV = 0)
        out[nm + "/frozen_eps"] = np.array(sorted(ref.QPED.frozenSetFromTVAndPe([0.0] * len(pe), list(pe), 0.2, None)), dtype=np.int64)
        kk = max(0, (1 << n) // 2 - 1)
        out[nm + "/frozen_k"] = np.array(sorted(ref.QPED.frozenSetFromTVAndPe([0.0] * len(pe), list(pe), None, kk)), dtype=np.int64)
        print(nm, "q", q, "N", 1 << n, "L", L, "sum Pe", pe.sum())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

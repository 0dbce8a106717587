"""Generate the code constructions (Pe vectors -> frozen sets) of the BASELINE configs with the LIVE reference.

TEST INFRASTRUCTURE, build container only (needs /root/reference).  Output: tests/golden/constructions/*.npy
(small float64 vectors, committed).  The frozen sets used by tests/ and bench.py are derived from these
vectors by `polarcub_b200.construction.frozen_set_from_pe` (stable sort by Pe, BinaryPolarEncoderDecoder.py:525).

  python oracle/gen_constructions.py bsc 10 0.11        # C1: BSC(0.11), N=1024, L=100   (~11 min)
  python oracle/gen_constructions.py biawgn 12 2.0      # C2: BI-AWGN Eb/N0=2 dB R=1/2, N=4096
  python oracle/gen_constructions.py qsc 3 11 0.02      # C3: QSC(q=3, p=0.02), N=2048
  python oracle/gen_constructions.py bec 10 0.1         # validation of the closed-form BEC recursion

Each level follows BinaryMemorylessDistribution.py:657-677 exactly:
  dist.minusTransform().degrade(L), dist.plusTransform().degrade(L), errorProb() per leaf.
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "constructions")
L = 100


def pe_binary(ref, n, dist0):
    dists = [dist0]
    for m in range(1, n + 1):
        nxt = []
        for d in dists:
            nxt.append(d.minusTransform().degrade(L))
            nxt.append(d.plusTransform().degrade(L))
        dists = nxt
        print("level", m, "done", time.time(), flush=True)
    return np.array([d.errorProb() for d in dists], dtype=np.float64)


def make_biawgn(ref, ebn0_db, rate, bins=400):
    """BI-AWGN quantised into `bins` output cells of [p(y,0), p(y,1)] (SURVEY.md 8d, C2 inputs).

    The reference has no AWGN construction (QaryMemorylessDistribution.py:800-804 is an empty stub), so the
    channel is built with BinaryMemorylessDistribution.append (BinaryMemorylessDistribution.py:36) from the
    exact cell probabilities of y = (1-2x) + N(0, sigma^2), sigma^2 = 1/(2 R 10^(EbN0/10)).
    """
    sigma = math.sqrt(1.0 / (2.0 * rate * 10.0 ** (ebn0_db / 10.0)))
    lo, hi = -6.0 * sigma - 1.0, 6.0 * sigma + 1.0
    edges = np.linspace(lo, hi, bins + 1)
    edges[0], edges[-1] = -np.inf, np.inf
    from scipy.stats import norm
    d = ref.BMD.BinaryMemorylessDistribution()
    for b in range(bins):
        p0 = 0.5 * (norm.cdf((edges[b + 1] - 1.0) / sigma) - norm.cdf((edges[b] - 1.0) / sigma))
        p1 = 0.5 * (norm.cdf((edges[b + 1] + 1.0) / sigma) - norm.cdf((edges[b] + 1.0) / sigma))
        d.append([float(p0), float(p1)])
    return d


def main():
    ref = refshim.load()
    os.makedirs(OUT, exist_ok=True)
    kind = sys.argv[1]
    t0 = time.time()
    if kind == "bsc":
        n, p = int(sys.argv[2]), float(sys.argv[3])
        pe = pe_binary(ref, n, ref.BMD.makeBSC(p))
        name = "bsc_p%s_n%d_L%d_pe.npy" % (sys.argv[3], n, L)
    elif kind == "bec":
        n, p = int(sys.argv[2]), float(sys.argv[3])
        pe = pe_binary(ref, n, ref.BMD.makeBEC(p))
        name = "bec_p%s_n%d_L%d_pe.npy" % (sys.argv[3], n, L)
    elif kind == "biawgn":
        n, ebn0 = int(sys.argv[2]), float(sys.argv[3])
        pe = pe_binary(ref, n, make_biawgn(ref, ebn0, 0.5))
        name = "biawgn_ebn0%s_n%d_L%d_pe.npy" % (sys.argv[3], n, L)
    elif kind == "qsc":
        q, n, p = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
        d = "/tmp/polarcub_qsc_q%d_n%d_p%s/" % (q, n, sys.argv[4])
        tv, pe = ref.QMD.calcTVAndPe_degradingUpgrading(n, L, None, ref.QMD.makeQSC(q, p), d, verbosity=True)
        pe = np.asarray(pe, dtype=np.float64)
        name = "qsc_q%d_p%s_n%d_L%d_pe.npy" % (q, sys.argv[4], n, L)
    else:
        raise SystemExit("unknown kind")
    np.save(os.path.join(OUT, name), pe)
    print("saved", name, "in %.1f s" % (time.time() - t0), "sum Pe=", pe.sum())


if __name__ == "__main__":
    main()

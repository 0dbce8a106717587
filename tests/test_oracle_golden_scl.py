"""Pins oracle.list_decode (oracle/polar_oracle_list.c) against QaryPolarEncoderDecoder.listDecode run live.

Checked bit-exactly: returned information, ProbResult, final list size, every surviving path IN ORDER, the
normalised path metrics (float64 equality) and the genie path's metric.
"""
import os

import numpy as np

import oracle


def test_list_decode_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "scl.npz"))
    names = [str(s) for s in g["names"]]
    assert len(names) >= 25
    bad = []
    for nm in names:
        q, n, L = int(g[nm + "/q"]), int(g[nm + "/n"]), int(g[nm + "/L"])
        N = 1 << n
        fm = g[nm + "/frozen"]
        for f in range(g[nm + "/xy"].shape[0]):
            info, pr, ls, linfo, lprob, ap = oracle.list_decode(q, N, L, fm, g[nm + "/xy"][f], g[nm + "/fv"][f],
                                                                g[nm + "/ainfo"][f], want_list=True)
            ok = (np.array_equal(info, g[nm + "/info"][f]) and pr == int(g[nm + "/pr"][f]) and ls == int(g[nm + "/lsize"][f])
                  and np.array_equal(linfo, g[nm + "/linfo"][f]) and np.array_equal(lprob, g[nm + "/lprob"][f])
                  and ap == float(g[nm + "/aprob"][f]))
            if not ok:
                bad.append((nm, f))
    assert not bad, bad

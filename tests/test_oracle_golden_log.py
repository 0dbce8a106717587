"""The oracle's LOG DOMAIN (`use_log=True`) against golden vectors from the live reference (oracle/gen_golden_log.py).

Bar: decisions (decoded symbols, list contents and order, ProbResult) identical; float64 list metrics and actual_prob within
LOG_TOL -- the reference's log branches run numpy.logaddexp (libm exp / log1p, which the C oracle shares) and
scipy.special.logsumexp (numpy's vectorised exp, which may differ from libm in the last bit), so the metrics are not a
bit-exact target (third-party arithmetic)."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qlog.npz")
LOG_TOL = 1e-12  # absolute on log metrics whose magnitude is O(1) .. O(100): far below any decision margin of the vectors


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _names(kind):
    g = np.load(GOLD)
    return [str(n) for n in g["names"] if str(g[str(n) + "/kind"]) == kind]


@pytest.mark.parametrize("name", _names("sc"))
def test_sc_log_vs_reference(gold, name):
    q, n = int(gold[name + "/q"]), int(gold[name + "/n"])
    N = 1 << n
    fm = gold[name + "/frozen"]
    xv = np.full((N, q), -np.log(q))
    for f in range(gold[name + "/xyl"].shape[0]):
        cw, info = oracle.q_decode(q, N, fm, xv, gold[name + "/xyl"][f], use_log=True)
        assert np.array_equal(info, gold[name + "/info"][f]), (name, f)


@pytest.mark.parametrize("name", _names("list"))
def test_list_log_vs_reference(gold, name):
    q, n, L = int(gold[name + "/q"]), int(gold[name + "/n"]), int(gold[name + "/L"])
    N = 1 << n
    fm = gold[name + "/frozen"]
    for f in range(gold[name + "/xyl"].shape[0]):
        info, pr, ls, linfo, lprob, ap = oracle.list_decode(q, N, L, fm, gold[name + "/xyl"][f], gold[name + "/fv"][f],
                                                            gold[name + "/ainfo"][f], want_list=True, use_log=True)
        assert np.array_equal(info, gold[name + "/info"][f]), (name, f)
        assert pr == int(gold[name + "/pr"][f]), (name, f)
        assert ls == int(gold[name + "/lsize"][f]), (name, f)
        assert np.array_equal(linfo[:ls], gold[name + "/linfo"][f][:ls]), (name, f)
        assert np.allclose(lprob[:ls], gold[name + "/lprob"][f][:ls], rtol=0, atol=LOG_TOL), (name, f)
        assert abs(ap - float(gold[name + "/aprob"][f])) <= LOG_TOL * max(1.0, abs(ap)), (name, f)

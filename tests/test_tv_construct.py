"""The native Tal-Vardy degrading construction (pc_tv_degrade_pe, csrc/tv_construct.cu -- host code, no GPU needed) against
the live reference: the small cases of tests/golden/tv_construct.npz (oracle/gen_golden_tv.py: BSC / BEC / random / duplicate-LLR
and zero-probability outputs, L from 1 to 33) and the BASELINE-size Pe vectors under tests/golden/constructions/
(oracle/gen_constructions.py: 11 minutes of the reference's Python at N = 1024).  Bar: float64-identical Pe vectors."""
import math
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_small_cases_match_the_live_reference():
    from polarcub_b200 import construction as c
    d = np.load(os.path.join(GOLD, "tv_construct.npz"), allow_pickle=True)
    for nm in d["names"]:
        n, L = (int(v) for v in d[nm + "/params"])
        tab = d[nm + "/table"]
        pe = c.tal_vardy_pe(n, L, tab, threads=3)
        np.testing.assert_array_equal(pe, d[nm + "/pe"], err_msg=str(nm))
        np.testing.assert_array_equal(c.tal_vardy_pe(n, L, tab, threads=1), pe)  # the thread split changes nothing
        fs = c.calcFrozenSet_degradingUpgrading(n, L, float(d[nm + "/eps"]), None, tab)
        assert sorted(fs) == [int(i) for i in d[nm + "/frozen"]], nm
        if n >= 1:  # the first degraded minus channel is the n = 1 construction's first leaf: its Pe pins the symbol table
            acc = 0.0
            for r in d[nm + "/first_minus"].tolist():
                acc += min(r)  # errorProb's plain running sum (builtin sum() would compensate)
            assert c.tal_vardy_pe(1, L, tab)[0] == acc


@pytest.mark.parametrize("name,n,kind", [("bsc_p0.11_n7_L100_pe.npy", 7, "bsc"), ("bsc_p0.11_n10_L100_pe.npy", 10, "bsc"),
                                         ("bec_p0.1_n8_L100_pe.npy", 8, "bec"), ("biawgn_ebn02.0_n8_L100_pe.npy", 8, "biawgn")])
def test_baseline_constructions_match_the_live_reference(name, n, kind):
    from polarcub_b200 import construction as c
    if kind == "bsc":
        tab = c.make_bsc(0.11)
    elif kind == "bec":
        tab = c.make_bec(0.1)
    else:  # the 400-bin quantised BI-AWGN of oracle/gen_constructions.py (Eb/N0 = 2 dB, R = 1/2)
        from scipy.stats import norm
        sigma = math.sqrt(1.0 / (2.0 * 0.5 * 10.0 ** (2.0 / 10.0)))
        edges = np.linspace(-6.0 * sigma - 1.0, 6.0 * sigma + 1.0, 401)
        edges[0], edges[-1] = -np.inf, np.inf
        tab = np.array([[float(0.5 * (norm.cdf((edges[b + 1] - 1.0) / sigma) - norm.cdf((edges[b] - 1.0) / sigma))),
                         float(0.5 * (norm.cdf((edges[b + 1] + 1.0) / sigma) - norm.cdf((edges[b] + 1.0) / sigma)))]
                        for b in range(400)])
    pe = c.tal_vardy_pe(n, 100, tab)
    np.testing.assert_array_equal(pe, c.load_pe(name))


def test_survey_known_answer_k361():
    """BSC(0.11), L = 100, epsilon = 0.1, N = 1024 -> K = 361 information indices."""
    from polarcub_b200 import construction as c
    fs = c.calcFrozenSet_degradingUpgrading(10, 100, 0.1, None, c.make_bsc(0.11))
    assert 1024 - len(fs) == 361


def test_argument_validation():
    from polarcub_b200 import construction as c
    from polarcub_b200._lib import PolarcubError
    with pytest.raises(PolarcubError):
        c.tal_vardy_pe(3, 0, c.make_bsc(0.1))  # L must be positive
    with pytest.raises(PolarcubError):
        c.calcFrozenSet_degradingUpgrading(3, 8, 0.1, c.make_bsc(0.1), c.make_bsc(0.1))  # non-uniform input: not offered


def test_qary_small_cases_match_the_live_reference():
    """pc_tv_degrade_pe_qary against QaryMemorylessDistribution.minusTransform().degrade(L) / plusTransform().degrade(L) +
    errorProb() of the live reference (q = 2..5, QSC / QEC / random channels, L down to 1)."""
    from polarcub_b200 import construction as c
    d = np.load(os.path.join(GOLD, "tv_construct.npz"), allow_pickle=True)
    for nm in d["qnames"]:
        q, n, L = (int(v) for v in d[nm + "/params"])
        pe = c.tal_vardy_pe_qary(q, n, L, d[nm + "/table"], threads=2)
        np.testing.assert_array_equal(pe, d[nm + "/pe"], err_msg=str(nm))
        np.testing.assert_array_equal(c.tal_vardy_pe_qary(q, n, L, d[nm + "/table"], threads=1), pe)
        # QaryPolarEncoderDecoder.frozenSetFromTVAndPe: epsilon rule, and a fixed count (which keeps one index more: reference quirk)
        fs = c.calcFrozenSet_degradingUpgrading_qary(q, n, L, None, d[nm + "/table"], upperBoundOnErrorProbability=0.2)
        assert sorted(fs) == [int(i) for i in d[nm + "/frozen_eps"]], nm
        kk = max(0, (1 << n) // 2 - 1)
        fk = c.calcFrozenSet_degradingUpgrading_qary(q, n, L, None, d[nm + "/table"], numInfoIndices=kk)
        assert sorted(fk) == [int(i) for i in d[nm + "/frozen_k"]], nm
        assert (1 << n) - len(fk) == kk + 1


def test_qary_baseline_construction_matches_the_live_reference():
    """C3's code: q = 3, QSC(0.02), N = 2048, L = 100 -- hours of the reference's Python, committed as
    tests/golden/constructions/qsc_q3_p0.02_n11_L100_pe.npy; the native routine reproduces all 2048 values bit for bit."""
    from polarcub_b200 import construction as c
    pe = c.tal_vardy_pe_qary(3, 11, 100, c.make_qsc(3, 0.02))
    np.testing.assert_array_equal(pe, c.load_pe("qsc_q3_p0.02_n11_L100_pe.npy"))


def test_directory_cache_is_the_reference_layout(tmp_path):
    """calcTVAndPe_degradingUpgrading writes / reads `DegradingUpgrading_L=<L>_tv.npy` and `_pe.npy` under directory_name
    (QaryMemorylessDistribution.py:936-947, :987-990), so a cache written here is what the reference would load."""
    from polarcub_b200 import construction as c
    d = np.load(os.path.join(GOLD, "tv_construct.npz"), allow_pickle=True)
    nm = "q3_qsc0.1_n4_L16"
    q, n, L = (int(v) for v in d[nm + "/params"])
    root = str(tmp_path) + "/cons/"
    tv, pe = c.calcTVAndPe_degradingUpgrading(n, L, None, d[nm + "/table"], root)
    np.testing.assert_array_equal(pe, d[nm + "/pe"])
    assert np.all(tv == 0.0)
    assert sorted(os.listdir(root)) == ["DegradingUpgrading_L=16_pe.npy", "DegradingUpgrading_L=16_tv.npy"]
    np.save(root + "DegradingUpgrading_L=16_pe.npy", np.asarray(pe) + 1.0)  # a second call must LOAD, not recompute
    _, pe2 = c.calcTVAndPe_degradingUpgrading(n, L, None, d[nm + "/table"], root)
    np.testing.assert_array_equal(pe2, np.asarray(pe) + 1.0)

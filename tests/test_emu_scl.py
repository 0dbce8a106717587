"""CPU check of the SC-list KERNEL SOURCES (polarcub_b200/csrc/scl_path.cu) run through tests/emu (a coroutine emulation of
warps: test infrastructure, never a product path) against the oracle: same bar as tests/test_gpu_scl.py -- returned
information, ProbResult, the final list in order with float64-equal metrics and the genie metric."""
import os
import sys

import numpy as np
import pytest

import oracle

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu import run_sclp  # noqa: E402


def _bec_order(n, eps=0.5):
    z = [eps]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return np.argsort(-np.array(z), kind="stable")


def _pattern(n, how, rate, rng):
    N = 1 << n
    if how == "bec":
        return set(int(i) for i in _bec_order(n)[:N - int(rate * N)])
    if how == "random":
        return set(int(i) for i in rng.permutation(N)[:N - int(rate * N)])
    m = np.zeros(N, dtype=bool)
    pos = 0
    while pos < N:
        run = int(rng.choice([1, 2, 3, 4, 8, 16, 31, 32, 64, 128, 256, 512, 1024]))
        if rng.random() > rate:
            m[pos:pos + run] = True
        pos += run
    if m.all():
        m[-1] = False
    return set(np.nonzero(m)[0].tolist())


def _frames(n, fm, B, rng, sigma, fv_random):
    N = 1 << n
    k = int(N - fm.sum())
    info = rng.integers(0, 2, size=(B, k))
    fv = rng.integers(0, 2, size=(B, N - k)) if fv_random else np.zeros((B, N - k), dtype=np.int64)
    u = np.zeros((B, N), dtype=np.int64)
    u[:, fm == 0] = info
    u[:, fm == 1] = fv
    cw = np.stack([oracle.polar_transform_qudits(2, u[b]) for b in range(B)])
    y = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
    return info, fv, y


def _xy(y, sigma):
    l0, l1 = -(y - 1) ** 2 / (2 * sigma ** 2), -(y + 1) ** 2 / (2 * sigma ** 2)
    m = np.maximum(l0, l1)
    return np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=-1)


def _check(n, L, fm, xy, fv, info, out):
    N = 1 << n
    for b in range(info.shape[0]):
        oi, opr, ols, olinfo, olprob, oap = oracle.list_decode(2, N, L, fm, xy[b], fv[b], info[b], want_list=True)
        np.testing.assert_array_equal(out["info"][b], oi, err_msg="frame %d" % b)
        assert int(out["prob_result"][b]) == opr, b
        if "list_size" in out:
            assert int(out["list_size"][b]) == ols, b
            np.testing.assert_array_equal(out["list_info"][b][:ols], olinfo[:ols], err_msg="frame %d" % b)
            assert np.array_equal(out["list_prob"][b][:ols], olprob[:ols]), b
            assert float(out["actual_prob"][b]) == oap, b


@pytest.mark.parametrize("n,L,how,rate,B,env", [
    (6, 3, "blocks", 0.5, 10, {}), (7, 2, "random", 0.75, 18, {}), (8, 16, "blocks", 0.8, 5, {}), (9, 8, "bec", 0.9, 9, {}),
    (10, 32, "blocks", 0.5, 3, {}), (11, 4, "blocks", 0.25, 9, {}), (11, 8, "random", 0.5, 5, {"PC_SCLP_LSM": "2"}),
    (5, 8, "blocks", 1.0, 9, {}), (4, 32, "random", 0.5, 5, {}), (8, 8, "bec", 0.5, 9, {"PC_SCLP_NOFUSE": "1"}),
    (9, 8, "blocks", 0.6, 6, {"PC_SCLP_LSM": "6", "PC_SCLP_RGL": "9"}), (10, 8, "bec", 0.5, 5, {"PC_SCLP_LSM": "1", "PC_SCLP_RGL": "1"}),
    (3, 8, "bec", 0.5, 9, {}), (1, 8, "bec", 0.5, 9, {}), (2, 4, "bec", 0.5, 17, {}), (7, 1, "bec", 0.5, 40, {}),
    # the 12-warp build (operands of the next step prefetched into registers); the default is the 16-warp build
    (10, 8, "bec", 0.5, 5, {"PC_SCLP_WARPS_PER_SM": "12"}), (9, 8, "blocks", 0.6, 6, {"PC_SCLP_WARPS_PER_SM": "12", "PC_SCLP_STAGES": "0"}),
])
def test_emulated_kernels_vs_oracle(n, L, how, rate, B, env, monkeypatch):
    for kk, v in env.items():
        monkeypatch.setenv(kk, v)
    N = 1 << n
    rng = np.random.default_rng(12000 + 101 * n + L)
    fs = _pattern(n, how, rate, rng)
    fm = np.zeros(N, dtype=np.uint8)
    fm[list(fs)] = 1
    sigma = 0.9
    info, fv, y = _frames(n, fm, B, rng, sigma, fv_random=n <= 8)
    xy = _xy(y, sigma)
    out = run_sclp.list_decode(n, L, fm, fv if n <= 8 else None, info, xy=xy)
    _check(n, L, fm, xy, fv, info, out)
    # without the list outputs the selected word and the classification must not change
    out2 = run_sclp.list_decode(n, L, fm, fv if n <= 8 else None, info, xy=xy, want_list=False)
    np.testing.assert_array_equal(out2["info"], out["info"])
    np.testing.assert_array_equal(out2["prob_result"], out["prob_result"])


def test_emulated_symbol_input():
    """Channel symbols + table (the fused makeQaryMemorylessVectorDistribution adapter): same results as the table rows
    expanded to probability pairs; 64-level quantised BI-AWGN."""
    n, L, B = 8, 8, 9
    N = 1 << n
    rng = np.random.default_rng(5)
    fm = np.zeros(N, dtype=np.uint8)
    fm[_bec_order(n)[:N // 2]] = 1
    sigma = 0.8
    info, fv, y = _frames(n, fm, B, rng, sigma, fv_random=False)
    Y = 64
    edges = np.linspace(-3.0, 3.0, Y - 1)
    centers = np.concatenate([[edges[0] - 0.05], (edges[:-1] + edges[1:]) / 2, [edges[-1] + 0.05]])
    table = _xy(centers, sigma)
    ys = np.searchsorted(edges, y).astype(np.uint8)
    xy = table[ys]
    out = run_sclp.list_decode(n, L, fm, None, info, y=ys, table=table)
    _check(n, L, fm, xy, fv, info, out)

"""The LOG DOMAIN (`use_log=True`) of the q-ary SC decoder and of listDecode on the GPU, through the C-ABI
(pc_qsc_decode_logprobs, pc_qsc_decode_symbols_log, pc_scl_decode_logprobs), against

  * golden vectors produced by the live reference with use_log=True (tests/golden/qlog.npz, oracle/gen_golden_log.py), and
  * the C oracle's log domain on fresh seeded inputs.

Tolerance (stated, BASELINE.json allows 1e-5 relative on decisions / 1e-4 on intermediate values): the log branches call
exp / log1p, where the device's math library and the host libm differ in the last bit now and then.  Decisions (decoded
symbols, list members and order, ProbResult) must be IDENTICAL on these tie-free inputs; float64 log metrics within 1e-11.
"""
import math
import os

import numpy as np
import pytest
import torch

import oracle
import polarcub_b200 as pcb

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qlog.npz")
TOL = 1e-11


def _names(kind):
    g = np.load(GOLD)
    return [str(n) for n in g["names"] if str(g[str(n) + "/kind"]) == kind]


def _fs(fm):
    return set(int(i) for i in np.nonzero(fm)[0])


@pytest.mark.parametrize("name", _names("sc"))
def test_sc_log_vs_reference_goldens(name):
    g = np.load(GOLD)
    q, n = int(g[name + "/q"]), int(g[name + "/n"])
    ed = pcb.QaryPolarEncoderDecoder(q, 1 << n, _fs(g[name + "/frozen"]), 1, use_log=True)
    info = ed.decode_batch(g[name + "/xyl"])
    assert np.array_equal(info, g[name + "/info"])
    # the reference's own call shape: one frame, a VectorDistribution-like object
    class VD:  # noqa: E306
        def __init__(self, p):
            self.probs = p
        def __len__(self):  # noqa: E301
            return self.probs.shape[0]
    one = ed.decode(VD(np.full((1 << n, q), -math.log(q))), VD(g[name + "/xyl"][0]))
    assert np.array_equal(one, g[name + "/info"][0])


@pytest.mark.parametrize("name", _names("list"))
def test_list_log_vs_reference_goldens(name):
    g = np.load(GOLD)
    q, n, L = int(g[name + "/q"]), int(g[name + "/n"]), int(g[name + "/L"])
    ed = pcb.QaryPolarEncoderDecoder(q, 1 << n, _fs(g[name + "/frozen"]), 1, use_log=True)
    info, res, lst = ed.listDecode_batch(g[name + "/xyl"], g[name + "/fv"], L, g[name + "/ainfo"], return_list=True)
    assert np.array_equal(info, g[name + "/info"])
    assert np.array_equal(res, g[name + "/pr"])
    assert np.array_equal(lst["list_size"], g[name + "/lsize"])
    for f in range(info.shape[0]):
        ls = int(g[name + "/lsize"][f])
        assert np.array_equal(lst["list_info"][f][:ls], g[name + "/linfo"][f][:ls]), (name, f)
        assert np.allclose(lst["list_prob"][f][:ls], g[name + "/lprob"][f][:ls], rtol=0, atol=TOL), (name, f)
        assert abs(lst["actual_prob"][f] - g[name + "/aprob"][f]) <= TOL * max(1.0, abs(g[name + "/aprob"][f])), (name, f)
    # the reference's entry point on one frame
    class VD:  # noqa: E306
        def __init__(self, p):
            self.probs = p
        def __len__(self):  # noqa: E301
            return self.probs.shape[0]
    k = ed.k
    one, pr = ed.listDecode(VD(g[name + "/xyl"][0]), g[name + "/fv"][0], L, np.zeros((k, 0), dtype=np.int64),
                            np.zeros(0, dtype=np.int64), actualInformation=g[name + "/ainfo"][0])
    assert np.array_equal(one, g[name + "/info"][0]) and pr.value == int(g[name + "/pr"][0])


def _jqsc_log(q, cw, rng, p=0.1):
    B, N = cw.shape
    y = np.where(rng.random((B, N)) > p, cw, (cw + rng.integers(1, q, (B, N))) % q)
    xy = np.where(np.arange(q)[None, None, :] == y[:, :, None], 1.0 - p, p / (q - 1)) * (1.0 + 0.2 * rng.random((B, N, q)))
    return np.log(xy)


@pytest.mark.parametrize("q,n,B", [(2, 10, 96), (3, 11, 64), (5, 6, 130), (7, 5, 40), (4, 8, 33)])
def test_sc_log_vs_oracle(q, n, B):
    """Fresh seeded frames at sizes up to C3's (q = 3, N = 2048): decoded symbols identical to the oracle's log domain.
    (A polarisation-ordered frozen set: with a random one most information symbols are coin flips whose marginals differ by
    rounding noise, and a last-bit difference of exp() legitimately flips them.)"""
    N = 1 << n
    rng = np.random.default_rng(400 + 10 * q + n)
    fm = np.zeros(N, dtype=np.uint8)
    fm[np.argsort(-np.array(_bec_z(n)), kind="stable")[:N // 2]] = 1
    lin = pcb.QaryPolarEncoderDecoder(q, N, _fs(fm), 1)
    cw = lin.encode_batch(rng.integers(0, q, (B, lin.k)))
    xyl = _jqsc_log(q, cw, rng)
    ed = pcb.QaryPolarEncoderDecoder(q, N, _fs(fm), 1, use_log=True)
    got = ed.decode_batch(xyl)
    xv = np.full((N, q), -math.log(q))
    for f in range(0, B, max(1, B // 24)):
        _, want = oracle.q_decode(q, N, fm, xv, xyl[f], use_log=True)
        assert np.array_equal(got[f], want), (q, n, f)


def test_sc_log_symbol_input_equals_logprob_input():
    """decode_symbols_batch with use_log=True (makeQaryMemorylessVectorDistribution(..., use_log=True): math.log of the table,
    -inf for 0) against decode_batch on the expanded log-probabilities, incl. a table with zero entries."""
    q, n, B = 3, 9, 70
    N = 1 << n
    rng = np.random.default_rng(77)
    fm = np.zeros(N, dtype=np.uint8)
    fm[np.argsort(-np.array(_bec_z(n)), kind="stable")[:N // 2]] = 1
    table = np.array([[0.9, 0.1, 0.0], [0.05, 0.9, 0.05], [0.0, 0.1, 0.9], [1 / 3, 1 / 3, 1 / 3]])
    lin = pcb.QaryPolarEncoderDecoder(q, N, _fs(fm), 1)
    cw = lin.encode_batch(rng.integers(0, q, (B, lin.k)))
    # channel: the sent symbol w.p. 0.85, a neighbour or the erasure-like output 3 otherwise (consistent with the table's zeros)
    r = rng.random((B, N))
    y = np.where(r < 0.85, cw, np.where(r < 0.93, 3, 1)).astype(np.uint8)
    ed = pcb.QaryPolarEncoderDecoder(q, N, _fs(fm), 1, use_log=True)
    with np.errstate(divide="ignore"):
        ltab = np.array([[math.log(v) if v != 0 else -math.inf for v in row] for row in table])
    a = ed.decode_symbols_batch(y, table)
    b = ed.decode_batch(ltab[y])
    assert np.array_equal(a, b)
    xv = np.full((N, q), -math.log(q))
    for f in range(0, B, 9):
        _, want = oracle.q_decode(q, N, fm, xv, ltab[y[f]], use_log=True)
        assert np.array_equal(a[f], want)


@pytest.mark.parametrize("q,n,L,B", [(2, 8, 8, 40), (2, 10, 4, 12), (3, 6, 4, 40), (3, 8, 8, 10), (5, 5, 4, 20)])
def test_list_log_vs_oracle(q, n, L, B):
    N = 1 << n
    rng = np.random.default_rng(900 + 10 * q + n + L)
    fm = np.zeros(N, dtype=np.uint8)
    order = np.argsort(-np.array(_bec_z(n)), kind="stable")
    fm[order[:N // 2]] = 1
    lin = pcb.QaryPolarEncoderDecoder(q, N, _fs(fm), 1)
    k = lin.k
    info = rng.integers(0, q, (B, k))
    fv = rng.integers(0, q, (B, N - k))
    u = np.zeros((B, N), dtype=np.int64)
    u[:, fm == 0] = info
    u[:, fm == 1] = fv
    cw = np.stack([_transform(q, row) for row in u])
    if q == 2:
        sigma = 0.9
        yv = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
        xyl = np.stack([-(yv - 1) ** 2 / (2 * sigma ** 2), -(yv + 1) ** 2 / (2 * sigma ** 2)], axis=-1)
    else:
        xyl = _jqsc_log(q, cw, rng, p=0.12)
    ed = pcb.QaryPolarEncoderDecoder(q, N, _fs(fm), 1, use_log=True)
    got, res, lst = ed.listDecode_batch(xyl, fv, L, info, return_list=True)
    for f in range(B):
        winfo, wpr, wls, wlinfo, wlprob, wap = oracle.list_decode(q, N, L, fm, xyl[f], fv[f], info[f], want_list=True,
                                                                  use_log=True)
        assert np.array_equal(got[f], winfo) and int(res[f]) == wpr and int(lst["list_size"][f]) == wls, (q, n, L, f)
        assert np.array_equal(lst["list_info"][f][:wls], wlinfo[:wls]), (q, n, L, f)
        assert np.allclose(lst["list_prob"][f][:wls], wlprob[:wls], rtol=0, atol=TOL), (q, n, L, f)
        assert abs(lst["actual_prob"][f] - wap) <= TOL * max(1.0, abs(wap)), (q, n, L, f)


def _bec_z(n, eps=0.5):
    z = [eps]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return z


def _transform(q, u):
    """x with polarTransformOfQudits(q, x) == u: the encoder's butterfly (QaryPolarEncoderDecoder.py:397-399)."""
    u = np.asarray(u, dtype=np.int64)
    if u.shape[0] == 1:
        return u
    h = u.shape[0] // 2
    m, p = _transform(q, u[:h]), _transform(q, u[h:])
    out = np.empty_like(u)
    out[0::2] = (m + p) % q
    out[1::2] = (-p) % q
    return out

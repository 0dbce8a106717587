"""GPU parity: binary encode + SC decode (through the C-ABI) vs the reference goldens and the CPU oracle.

Bar: BIT-EXACT (the kernels reproduce the reference's float64 arithmetic, so there is no tolerance).
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import polarcub_b200
    return polarcub_b200


def _bec_order(n, eps=0.5):
    z = [eps]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return np.argsort(-np.array(z), kind="stable")


def test_golden_binary(golden_dir):
    pcb = _setup()
    g = np.load(os.path.join(golden_dir, "sc_binary.npz"))
    checked = 0
    for nm in [str(s) for s in g["names"]]:
        n = int(g[nm + "/n"])
        N = 1 << n
        fs = set(np.nonzero(g[nm + "/frozen"])[0].tolist())
        ed = pcb.BinaryPolarEncoderDecoder(N, fs, int(g[nm + "/seed"]))
        np.testing.assert_array_equal(ed.randomlyGeneratedNumbers, g[nm + "/r"])
        info, cw, xy = g[nm + "/info"], g[nm + "/cw"], g[nm + "/xy"]
        xp = g[nm + "/xprobs"]  # two cases carry a non-uniform a-priori distribution (data-dependent frozen bits)
        np.testing.assert_array_equal(ed.encode_batch(info, xp), cw, err_msg=nm)
        dcw, dinfo = ed.decode_batch(xy, xp)
        np.testing.assert_array_equal(dcw, g[nm + "/dec_cw"], err_msg=nm)
        np.testing.assert_array_equal(dinfo, g[nm + "/dec_info"], err_msg=nm)
        checked += 1
    assert checked >= 79


def test_reference_style_single_frame_api(golden_dir):
    pcb = _setup()
    from polarcub_b200.VectorDistributions.BinaryMemorylessVectorDistribution import BinaryMemorylessVectorDistribution
    g = np.load(os.path.join(golden_dir, "sc_binary.npz"))
    nm = "n3_survey_seed-1"
    ed = pcb.BinaryPolarEncoderDecoder(8, {0, 1, 2, 4}, -1)
    assert ed.k == 4
    x = BinaryMemorylessVectorDistribution(8)
    x.probs[:] = 0.5
    cw = ed.encode(x, [1, 0, 1, 1])
    assert cw.dtype == np.int64
    np.testing.assert_array_equal(cw, [1, 1, 0, 0, 1, 1, 0, 1])
    np.testing.assert_array_equal(pcb.polarTransformOfBits(cw), [1, 1, 1, 1, 1, 0, 1, 1])
    xy = BinaryMemorylessVectorDistribution(8)
    xy.probs[:] = g[nm + "/xy"][0]
    dcw, dinfo = ed.decode(x, xy)
    assert dcw.dtype == np.int64 and dinfo.dtype == np.int64
    np.testing.assert_array_equal(dcw, g[nm + "/dec_cw"][0])
    np.testing.assert_array_equal(dinfo, g[nm + "/dec_info"][0])
    # a non-uniform prior takes the lock-step a-priori tree (data-dependent frozen bits), still on the GPU
    import oracle
    x.probs[:, 0] = 0.7
    x.probs[:, 1] = 0.3
    dcw, dinfo = ed.decode(x, xy)
    ocw, oinfo = oracle.bin_decode(8, ed.frozenMask, ed.randomlyGeneratedNumbers, x.probs, xy.probs)
    np.testing.assert_array_equal(dcw, ocw)
    np.testing.assert_array_equal(dinfo, oinfo)


@pytest.mark.parametrize("n,seed", [(4, 1), (7, 3), (8, -1), (10, 1), (11, 5)])
def test_nonuniform_prior_vs_oracle(n, seed):
    """Honda-Yamamoto style shaping: frozen bits follow `0 iff P(u_i = 0 | past) >= r_i` on the a-priori tree
    (BinaryPolarEncoderDecoder.py:258-262); encode, decode and both genie captures against the oracle (pinned on the
    reference's own non-uniform goldens in tests/test_oracle_golden.py)."""
    import oracle
    pcb = _setup()
    N = 1 << n
    rng = np.random.default_rng(4000 + n)
    k = N // 2
    fs = set(int(i) for i in _bec_order(n)[:N - k])
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, seed)
    p1 = 0.3
    xp = np.tile(np.array([1 - p1, p1]), (N, 1))
    if n == 7:
        xp = np.stack([1 - 0.5 * rng.random(N), 0.5 * rng.random(N)], axis=1)  # position-dependent prior
    B = 48
    info = rng.integers(0, 2, size=(B, k))
    cw = ed.encode_batch(info, xp)
    ocw = oracle.bin_encode_batch(N, ed.frozenMask, ed.randomlyGeneratedNumbers, xp, info)
    np.testing.assert_array_equal(cw, ocw)
    pch = 0.06
    y = cw ^ (rng.random((B, N)) < pch)
    # joint probabilities P(x, y) under the prior
    xy = np.where((y[..., None] == np.arange(2)), 1 - pch, pch) * xp[None]
    dcw, dinfo = ed.decode_batch(xy, xp)
    for b in range(B):
        o1, o2 = oracle.bin_decode(N, ed.frozenMask, ed.randomlyGeneratedNumbers, xp, xy[b])
        np.testing.assert_array_equal(dcw[b], o1, err_msg="frame %d" % b)
        np.testing.assert_array_equal(dinfo[b], o2, err_msg="frame %d" % b)
    # genie runs with the prior: captured marginals of the a-priori (encode) and a-posteriori (decode) trees
    seeds = [11, 12, 13]
    enc, TV, H = ed.genie_encode_batch(xp, seeds)
    ones = np.ones(N, dtype=np.uint8)
    for t, sd in enumerate(seeds):
        r = oracle.common_randomness(N, sd)
        ecw, emarg = oracle.bin_encode(N, ones, r, xp, np.zeros(0, dtype=np.int64), want_marg=True)
        np.testing.assert_array_equal(enc[t], ecw)
        assert np.array_equal(TV[t], np.abs(emarg[:, 0] - emarg[:, 1]))
    ych = enc ^ (rng.random(enc.shape) < pch)
    xyg = np.where((ych[..., None] == np.arange(2)), 1 - pch, pch) * xp[None]
    dec, Pe, Hd, marg = ed.genie_decode_batch(xp, xyg, seeds, True, return_marginals=True)
    for t, sd in enumerate(seeds):
        r = oracle.common_randomness(N, sd)
        gcw, _, gm = oracle.bin_decode(N, ones, r, xp, xyg[t], want_marg=True)
        np.testing.assert_array_equal(dec[t], gcw)
        assert np.array_equal(marg[t], gm)
        assert np.array_equal(Pe[t], np.minimum(gm[:, 0], gm[:, 1]))


@pytest.mark.parametrize("n,kind", [(5, "bsc"), (6, "bec"), (9, "bsc"), (10, "bsc"), (10, "awgn"), (10, "bec_lossy"),
                                    (11, "awgn"), (12, "bsc")])
def test_random_frames_vs_oracle(n, kind):
    """Seeded synthetic frames at sizes the oracle finishes in seconds; decisions must be identical."""
    pcb = _setup()
    N = 1 << n
    rng = np.random.default_rng(1000 + n)
    k = N // 2
    fs = set(int(i) for i in _bec_order(n)[:N - k])
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, 1)
    B = 300 if n <= 10 else 70  # not multiples of 32: exercises the ragged tail
    info = rng.integers(0, 2, size=(B, k))
    cw = ed.encode_batch(info)
    fm, r = ed.frozenMask, ed.randomlyGeneratedNumbers
    xp = np.full((N, 2), 0.5)
    np.testing.assert_array_equal(cw[:16], oracle.bin_encode_batch(N, fm, r, xp, info[:16]))
    if kind == "bsc":
        p = 0.11
        tab = np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])
        y = (cw ^ (rng.random((B, N)) < p)).astype(np.uint8)
    elif kind == "bec":
        p = 0.45
        tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
        y = np.where(rng.random((B, N)) < p, 2, cw).astype(np.uint8)
    elif kind == "bec_lossy":
        p = 0.3
        tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
        y = np.where(rng.random((B, N)) < p, 2, cw ^ (rng.random((B, N)) < 0.03)).astype(np.uint8)
    else:
        tab = None
    if tab is not None:
        xy = tab[y]
        dcw_s, dinfo_s = ed.decode_symbols_batch(y, tab)
    else:
        sigma = 0.8
        yv = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
        l0, l1 = -(yv - 1) ** 2 / (2 * sigma ** 2), -(yv + 1) ** 2 / (2 * sigma ** 2)
        m = np.maximum(l0, l1)
        xy = np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=-1)
    dcw, dinfo = ed.decode_batch(xy)
    ocw, oinfo = oracle.bin_decode_batch(N, fm, r, xp, xy)
    np.testing.assert_array_equal(dcw, ocw)
    np.testing.assert_array_equal(dinfo, oinfo)
    if tab is not None:  # the symbol-table entry point is the same decoder fused with the table lookup
        np.testing.assert_array_equal(dcw_s, ocw)
        np.testing.assert_array_equal(dinfo_s, oinfo)
    # size-independent properties: re-encoding the decoded information reproduces the decoded codeword,
    # and the inverse transform of a codeword returns u with the frozen values in place
    np.testing.assert_array_equal(ed.encode_batch(dinfo), dcw)
    u = np.array(pcb.polarTransformOfBits(cw[0]))
    np.testing.assert_array_equal(u[fm == 0], info[0])
    np.testing.assert_array_equal(u[fm == 1], ed.frozenValues[fm == 1])


def test_edge_cases():
    pcb = _setup()
    # empty batch, N=1, all frozen, nothing frozen
    ed = pcb.BinaryPolarEncoderDecoder(16, {0, 1, 2}, 1)
    cw, info = ed.decode_batch(np.zeros((0, 16, 2)))
    assert cw.shape == (0, 16) and info.shape == (0, 13)
    assert ed.encode_batch(np.zeros((0, 13), dtype=np.int64)).shape == (0, 16)
    for fs in (set(), {0}):
        e1 = pcb.BinaryPolarEncoderDecoder(1, fs, 3)
        xy = np.array([[[0.2, 0.7]], [[0.7, 0.2]], [[0.3, 0.3]], [[0.0, 0.0]]])
        c, i = e1.decode_batch(xy)
        oc, oi = oracle.bin_decode_batch(1, e1.frozenMask, e1.randomlyGeneratedNumbers, np.full((1, 2), 0.5), xy)
        np.testing.assert_array_equal(c, oc)
        np.testing.assert_array_equal(i, oi)
    rng = np.random.default_rng(5)
    for fs in (set(range(64)), set()):
        e = pcb.BinaryPolarEncoderDecoder(64, fs, 7)
        xy = rng.random((40, 64, 2))
        c, i = e.decode_batch(xy)
        oc, oi = oracle.bin_decode_batch(64, e.frozenMask, e.randomlyGeneratedNumbers, np.full((64, 2), 0.5), xy)
        np.testing.assert_array_equal(c, oc)
        np.testing.assert_array_equal(i, oi)

"""GPU parity: the frame-per-CTA streamed SC decoder (sc_stream.cu, the large-block path) vs the CPU oracle.

PC_SC_STREAM=1 forces the streamed decoder for block lengths the frame-per-lane decoder also covers, so it is checked
at sizes the oracle finishes in seconds (N = 64 .. 2^17) with several frozen-set shapes; the BASELINE N = 2^20 BEC
configuration is checked on two frames against the oracle and through size-independent properties.
Bar: BIT-EXACT decisions (codeword and information bits).
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _z(n, eps):
    z = [eps]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return np.array(z)


def _frozen(n, k, how, rng):
    N = 1 << n
    if how == "bec":
        return set(int(i) for i in np.argsort(-_z(n, 0.5), kind="stable")[:N - k])
    if how == "random":
        return set(int(i) for i in rng.permutation(N)[:N - k])
    if how == "blocks":  # long frozen and long free runs: rate-0 nodes of many sizes
        m = np.zeros(N, dtype=bool)
        pos = 0
        while pos < N:
            run = int(rng.choice([1, 2, 3, 8, 32, 33, 64, 100, 256]))
            if rng.random() < 0.5:
                m[pos:pos + run] = True
            pos += run
        return set(np.nonzero(m)[0].tolist())
    raise ValueError(how)


def _channel(kind, cw, rng):
    B, N = cw.shape
    if kind == "bsc":
        p = 0.11
        tab = np.array([[0.5 * (1 - p), 0.5 * p], [0.5 * p, 0.5 * (1 - p)]])
        return tab, (cw ^ (rng.random((B, N)) < p)).astype(np.uint8)
    if kind == "bec":
        p = 0.35
        tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
        return tab, np.where(rng.random((B, N)) < p, 2, cw).astype(np.uint8)
    if kind == "bec_clean":  # few erasures: most all-information sub-trees see hard knowledge only
        p = 0.04
        tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
        return tab, np.where(rng.random((B, N)) < p, 2, cw).astype(np.uint8)
    if kind == "bec_soft":  # hard, erased AND soft outputs: rate-1 shortcut without the byte-state upper stages
        tab = np.array([[0.30, 0.0], [0.0, 0.30], [0.05, 0.05], [0.12, 0.03], [0.03, 0.12]])
        r = rng.random((B, N))
        y = np.where(r < 0.6, cw, np.where(r < 0.7, 2, np.where(r < 0.97, 3 + cw, 4 - cw)))
        return tab, y.astype(np.uint8)
    if kind == "bec_lossy":  # contradictions: (0,0) states appear
        p = 0.3
        tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
        return tab, np.where(rng.random((B, N)) < p, 2, cw ^ (rng.random((B, N)) < 0.03)).astype(np.uint8)
    raise ValueError(kind)


@pytest.mark.parametrize("n,how,kind,seed", [(6, "bec", "bsc", 1), (7, "random", "bec", -1), (8, "blocks", "bsc", 5),
                                             (9, "bec", "bec_lossy", 1), (10, "bec", "bsc", 1), (11, "blocks", "bec", 2),
                                             (12, "random", "bsc", 1), (13, "bec", "awgn", 1), (14, "blocks", "bsc", 9)])
def test_stream_decoder_vs_oracle(n, how, kind, seed, monkeypatch):
    monkeypatch.setenv("PC_SC_STREAM", "1")
    import polarcub_b200 as pcb
    N = 1 << n
    rng = np.random.default_rng(4000 + 31 * n)
    fs = _frozen(n, N // 2, how, rng)
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, seed)
    B = 37 if n <= 11 else 9
    info = rng.integers(0, 2, size=(B, ed.k))
    cw = ed.encode_batch(info)
    fm, r = ed.frozenMask, ed.randomlyGeneratedNumbers
    xp = np.full((N, 2), 0.5)
    if kind == "awgn":
        sigma = 0.8
        yv = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
        l0, l1 = -(yv - 1) ** 2 / (2 * sigma ** 2), -(yv + 1) ** 2 / (2 * sigma ** 2)
        m = np.maximum(l0, l1)
        xy = np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=-1)
        tab = None
    else:
        tab, y = _channel(kind, cw, rng)
        xy = tab[y]
    ocw, oinfo = oracle.bin_decode_batch(N, fm, r, xp, xy)
    dcw, dinfo = ed.decode_batch(xy)
    np.testing.assert_array_equal(dcw, ocw)
    np.testing.assert_array_equal(dinfo, oinfo)
    if tab is not None:
        dcw_s, dinfo_s = ed.decode_symbols_batch(y, tab)
        np.testing.assert_array_equal(dcw_s, ocw)
        np.testing.assert_array_equal(dinfo_s, oinfo)
    # the streamed and the frame-per-lane decoders agree
    monkeypatch.setenv("PC_SC_STREAM", "0")
    dcw2, dinfo2 = ed.decode_batch(xy)
    np.testing.assert_array_equal(dcw2, dcw)
    np.testing.assert_array_equal(dinfo2, dinfo)


@pytest.mark.parametrize("n,frames", [(17, 3), (20, 2)])
def test_large_block_bec(n, frames):
    """BASELINE config 4 shape: BEC(0.1), R = 0.8, frozen set from the closed-form BEC recursion."""
    import polarcub_b200 as pcb
    N = 1 << n
    K = int(0.8 * N)
    order = np.argsort(_z(n, 0.1), kind="stable")  # ascending Z: best indices first
    fs = set(int(i) for i in order[K:])
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, 1)
    assert ed.k == K
    rng = np.random.default_rng(77 + n)
    info = rng.integers(0, 2, size=(frames, K))
    cw = ed.encode_batch(info)
    p = 0.1
    tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])  # makeBEC, BinaryMemorylessDistribution.py:493-499
    y = np.where(rng.random((frames, N)) < p, 2, cw).astype(np.uint8)
    dcw, dinfo = ed.decode_symbols_batch(y, tab)
    # size-independent properties: the decoded codeword is the encoding of the decoded information, and it agrees with
    # every unerased channel output (the BEC never contradicts a correct decision)
    np.testing.assert_array_equal(ed.encode_batch(dinfo), dcw)
    known = y != 2
    assert np.array_equal(dcw[known], cw[known])
    # the oracle on the same frames
    xp = np.full((N, 2), 0.5)
    for f in range(frames):
        ocw, oinfo = oracle.bin_decode(N, ed.frozenMask, ed.randomlyGeneratedNumbers, xp, tab[y[f]])
        np.testing.assert_array_equal(dcw[f], ocw)
        np.testing.assert_array_equal(dinfo[f], oinfo)


@pytest.mark.parametrize("n,how,kind,seed,B", [(11, "bec", "bsc", 1, 70), (12, "blocks", "bec_lossy", 5, 45), (13, "random", "bsc", -1, 33),
                                               (15, "blocks", "bec", 2, 40), (17, "bec", "bsc", 1, 6),
                                               # erasure channels take the rate-1 shortcut of the sub-block kernel
                                               (12, "bec", "bec", 1, 64), (13, "bec", "bec_lossy", -1, 40), (14, "random", "bec", 3, 33),
                                               (12, "bec", "bec_clean", 1, 96), (13, "blocks", "bec_clean", 2, 64),
                                               (12, "bec", "bec_soft", 1, 64), (13, "random", "bec_soft", 4, 40),
                                               (16, "bec", "bec_lossy", 2, 36)])
def test_hybrid_decoder_vs_oracle(n, how, kind, seed, B, monkeypatch):
    """The hybrid large-block decoder (element-parallel upper stages through HBM + frame-per-lane 1024-leaf sub-blocks),
    forced for block lengths the other decoders cover too: bit-exact against the oracle and against the other decoders."""
    monkeypatch.setenv("PC_SC_HYBRID", "1")
    import polarcub_b200 as pcb
    N = 1 << n
    rng = np.random.default_rng(9000 + 31 * n)
    fs = _frozen(n, N // 2, how, rng)
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, seed)
    info = rng.integers(0, 2, size=(B, ed.k))
    cw = ed.encode_batch(info)
    tab, y = _channel(kind, cw, rng)
    dcw, dinfo = ed.decode_symbols_batch(y, tab)
    xp = np.full((N, 2), 0.5)
    nchk = B if n <= 13 else 4
    ocw, oinfo = oracle.bin_decode_batch(N, ed.frozenMask, ed.randomlyGeneratedNumbers, xp, tab[y[:nchk]])
    np.testing.assert_array_equal(dcw[:nchk], ocw)
    np.testing.assert_array_equal(dinfo[:nchk], oinfo)
    monkeypatch.setenv("PC_SC_R1", "0")  # the same walk without the rate-1 shortcut
    dcw3, dinfo3 = ed.decode_symbols_batch(y, tab)
    np.testing.assert_array_equal(dcw3, dcw)
    np.testing.assert_array_equal(dinfo3, dinfo)
    monkeypatch.delenv("PC_SC_R1")
    monkeypatch.setenv("PC_SC_HYBRID", "0")
    dcw2, dinfo2 = ed.decode_symbols_batch(y, tab)
    np.testing.assert_array_equal(dcw2, dcw)
    np.testing.assert_array_equal(dinfo2, dinfo)


def test_hybrid_large_block_bec_batch(monkeypatch):
    """N = 2^20, BEC(0.1), R = 0.8 (BASELINE config 4) through the hybrid decoder (the default for batches of thousands of
    frames; forced here): size-independent properties on the whole batch, the oracle on two frames, and agreement with
    the streamed decoder."""
    import os
    import polarcub_b200 as pcb
    monkeypatch.setenv("PC_SC_HYBRID", "1")
    n, frames = 20, 512
    N = 1 << n
    K = int(0.8 * N)
    order = np.argsort(_z(n, 0.1), kind="stable")
    fs = set(int(i) for i in order[K:])
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, 1)
    rng = np.random.default_rng(4242)
    info = rng.integers(0, 2, size=(frames, K), dtype=np.int8)
    cw = ed.encode_batch(info)
    p = 0.1
    tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
    y = np.where(rng.random((frames, N)) < p, 2, cw).astype(np.uint8)
    dcw, dinfo = ed.decode_symbols_batch(y, tab)
    np.testing.assert_array_equal(ed.encode_batch(dinfo), dcw)
    known = y != 2
    assert np.array_equal(dcw[known], cw[known])
    xp = np.full((N, 2), 0.5)
    for f in (0, frames - 1):
        ocw, oinfo = oracle.bin_decode(N, ed.frozenMask, ed.randomlyGeneratedNumbers, xp, tab[y[f]])
        np.testing.assert_array_equal(dcw[f], ocw)
        np.testing.assert_array_equal(dinfo[f], oinfo)
    monkeypatch.setenv("PC_SC_HYBRID", "0")
    scw, sinfo = ed.decode_symbols_batch(y[:8], tab)
    np.testing.assert_array_equal(scw, dcw[:8])
    np.testing.assert_array_equal(sinfo, dinfo[:8])


@pytest.mark.parametrize("kind", ["bec", "bsc", "bec_soft"])
def test_hybrid_decoder_multi_chunk(kind, monkeypatch):
    """Batches larger than the hybrid decoder's frame cap are walked chunk by chunk (PC_SC_HYBRID_FRAMES=64 here): the result
    does not depend on the chunking, for the byte-state, the float64 + rate-1 and the plain float64 walks."""
    import polarcub_b200 as pcb
    n, B = 13, 150
    N = 1 << n
    rng = np.random.default_rng(515)
    fs = _frozen(n, N // 2, "bec", rng)
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, 1)
    info = rng.integers(0, 2, size=(B, ed.k))
    cw = ed.encode_batch(info)
    tab, y = _channel(kind, cw, rng)
    monkeypatch.setenv("PC_SC_HYBRID", "1")
    ref_cw, ref_info = ed.decode_symbols_batch(y, tab)
    monkeypatch.setenv("PC_SC_HYBRID_FRAMES", "64")
    dcw, dinfo = ed.decode_symbols_batch(y, tab)
    np.testing.assert_array_equal(dcw, ref_cw)
    np.testing.assert_array_equal(dinfo, ref_info)
    monkeypatch.delenv("PC_SC_HYBRID_FRAMES")
    monkeypatch.setenv("PC_SC_HYBRID", "0")
    scw, sinfo = ed.decode_symbols_batch(y, tab)
    np.testing.assert_array_equal(scw, ref_cw)
    np.testing.assert_array_equal(sinfo, ref_info)


def test_large_block_bec_64_frames_vs_oracle(monkeypatch):
    """BASELINE config 4 (N = 2^20, R = 0.8, BEC(0.1)) through the HYBRID decoder on one-byte state codes (the path bench.py
    times; forced here because it engages on its own only from 6144 frames): 64 frames, every codeword and information word
    against the oracle."""
    import polarcub_b200 as pcb
    from concurrent.futures import ThreadPoolExecutor
    monkeypatch.setenv("PC_SC_HYBRID", "1")
    n, frames = 20, 64
    N = 1 << n
    K = int(0.8 * N)
    order = np.argsort(_z(n, 0.1), kind="stable")
    fs = set(int(i) for i in order[K:])
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, 1)
    rng = np.random.default_rng(2020)
    info = rng.integers(0, 2, size=(frames, K))
    cw = ed.encode_batch(info)
    p = 0.1
    tab = np.array([[0.5 * (1 - p), 0.0], [0.0, 0.5 * (1 - p)], [0.5 * p, 0.5 * p]])
    y = np.where(rng.random((frames, N)) < p, 2, cw).astype(np.uint8)
    dcw, dinfo = ed.decode_symbols_batch(y, tab)
    xp = np.full((N, 2), 0.5)
    oracle.lib()

    def one(f):
        return oracle.bin_decode(N, ed.frozenMask, ed.randomlyGeneratedNumbers, xp, tab[y[f]])

    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(one, range(frames)))
    for f, (ocw, oinfo) in enumerate(res):
        np.testing.assert_array_equal(dcw[f], ocw, err_msg="frame %d" % f)
        np.testing.assert_array_equal(dinfo[f], oinfo, err_msg="frame %d" % f)


@pytest.mark.parametrize("kind", ["bsc", "bec"])
def test_hybrid_random_frozen_set_exact_workspace(kind, monkeypatch):
    """Forced hybrid decoder at N = 2^17 with a RANDOM frozen set (sub-blocks of every rate: the sub-block scratch is sized
    from all of them) in a workspace of exactly pc_sc_workspace_bytes_symbols, against the oracle."""
    import polarcub_b200 as pcb
    monkeypatch.setenv("PC_SC_HYBRID", "1")
    n, B = 17, 24
    N = 1 << n
    rng = np.random.default_rng(1717)
    fm = np.zeros(N, dtype=np.uint8)
    # rates from ~0 to ~1 across the 128 sub-blocks of 1024 leaves, the last ones almost all information
    for sb in range(N // 1024):
        rate = sb / (N // 1024 - 1)
        fm[sb * 1024:(sb + 1) * 1024] = rng.random(1024) >= rate
    fm[-1024:] = 0
    fs = set(np.nonzero(fm)[0].tolist())
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, 3)
    info = rng.integers(0, 2, size=(B, ed.k))
    cw = ed.encode_batch(info)
    tab, y = _channel(kind, cw, rng)
    dcw, dinfo = ed.decode_symbols_batch(y, tab)
    xp = np.full((N, 2), 0.5)
    for f in range(B):
        ocw, oinfo = oracle.bin_decode(N, ed.frozenMask, ed.randomlyGeneratedNumbers, xp, tab[y[f]])
        np.testing.assert_array_equal(dcw[f], ocw, err_msg="frame %d" % f)
        np.testing.assert_array_equal(dinfo[f], oinfo, err_msg="frame %d" % f)

"""GPU parity: q-ary encode + SC decode vs the reference goldens and the CPU oracle (bit-exact)."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _bec_order(n, eps=0.5):
    z = [eps]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    return np.argsort(-np.array(z), kind="stable")


def test_golden_qary(golden_dir):
    import polarcub_b200 as pcb
    g = np.load(os.path.join(golden_dir, "sc_qary.npz"))
    checked = 0
    for nm in [str(s) for s in g["names"]]:
        q, n = int(g[nm + "/q"]), int(g[nm + "/n"])
        N = 1 << n
        fs = set(np.nonzero(g[nm + "/frozen"])[0].tolist())
        ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
        info, cw, xy = g[nm + "/info"], g[nm + "/cw"], g[nm + "/xy"]
        np.testing.assert_array_equal(ed.encode_batch(info), cw, err_msg=nm)
        dcw, dinfo = ed.decode_batch(xy, return_codeword=True)
        np.testing.assert_array_equal(dinfo, g[nm + "/dec_info"], err_msg=nm)
        np.testing.assert_array_equal(dcw[0], g[nm + "/dec_cw0"], err_msg=nm)
        checked += 1
    assert checked >= 60


@pytest.mark.parametrize("q,n", [(3, 9), (3, 11), (2, 8), (5, 7), (4, 6), (7, 6), (8, 5)])
def test_random_frames_vs_oracle(q, n):
    import polarcub_b200 as pcb
    N = 1 << n
    rng = np.random.default_rng(77 + q * 100 + n)
    k = N // 2
    fs = set(int(i) for i in _bec_order(n)[:N - k])
    ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
    B = 70 if n <= 9 else 40
    info = rng.integers(0, q, size=(B, k))
    cw = ed.encode_batch(info)
    xp = np.full((N, q), 1.0 / q)
    for f in range(4):
        np.testing.assert_array_equal(cw[f], oracle.q_encode(q, N, ed.frozenMask, xp, info[f]))
    p = 0.06
    tab = np.array([[1.0 - p if x == y else p / (q - 1) for x in range(q)] for y in range(q)])
    err = rng.random((B, N)) < p
    y = np.where(err, (cw + rng.integers(1, q, size=(B, N))) % q, cw)
    xy = tab[y]
    dcw, dinfo = ed.decode_batch(xy, return_codeword=True)
    ocw, oinfo = oracle.q_decode_batch(q, N, ed.frozenMask, xp, xy)
    np.testing.assert_array_equal(dinfo, oinfo)
    np.testing.assert_array_equal(dcw, ocw)
    # reference-style single-frame call returns only the information (QaryPolarEncoderDecoder.py:116)
    from polarcub_b200.VectorDistributions.QaryMemorylessVectorDistribution import QaryMemorylessVectorDistribution
    xv, xyv = QaryMemorylessVectorDistribution(q, N), QaryMemorylessVectorDistribution(q, N)
    xv.probs[:] = xp
    xyv.probs[:] = xy[0]
    one = ed.decode(xv, xyv)
    assert one.dtype == np.int64 and one.shape == (k,)
    np.testing.assert_array_equal(one, oinfo[0])
    np.testing.assert_array_equal(ed.encode(xv, info[0]), cw[0])


def test_qary_symbols_entry_point_matches_probs():
    """pc_qsc_decode_symbols (channel table lookup fused into the ingest) == pc_qsc_decode_probs on tab[y]."""
    import polarcub_b200 as pcb
    q, n = 3, 9
    N = 1 << n
    rng = np.random.default_rng(31)
    z = [0.5]
    for _ in range(n):
        z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
    fs = set(int(i) for i in np.argsort(-np.array(z), kind="stable")[:N // 2])
    ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
    p = 0.05
    tab = np.full((q, q), p / (q - 1))
    np.fill_diagonal(tab, 1.0 - p)
    info = rng.integers(0, q, size=(300, ed.k))
    cw = ed.encode_batch(info)
    err = rng.random(cw.shape) < p
    y = np.where(err, (cw + rng.integers(1, q, size=cw.shape)) % q, cw).astype(np.uint8)
    c1, i1 = ed.decode_symbols_batch(y, tab, return_codeword=True)
    c2, i2 = ed.decode_batch(tab[y], return_codeword=True)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_array_equal(c1, c2)
    # erasure-like extra output symbol (Y = q + 1 rows)
    tab2 = np.vstack([tab, np.full((1, q), 1.0 / q)])
    y2 = np.where(rng.random(cw.shape) < 0.1, q, y).astype(np.uint8)
    np.testing.assert_array_equal(ed.decode_symbols_batch(y2, tab2), ed.decode_batch(tab2[y2]))


@pytest.mark.parametrize("q,n,how", [(2, 3, "random"), (3, 3, "random"), (3, 4, "half"), (4, 6, "random"), (3, 11, "bec"),
                                     (2, 10, "bec"), (3, 7, "tail")])
def test_qary_symbol_lookup_variant(q, n, how, monkeypatch):
    """The lookup-table variant of the symbol entry point (level n-1 never stored) against the expanding ingest
    (PC_QSC_LUT=0), the probability entry point and the oracle; out-of-range symbols behave like the all-ones row."""
    import polarcub_b200 as pcb
    N = 1 << n
    rng = np.random.default_rng(77 + 13 * n + q)
    if how == "random":
        fs = set(int(i) for i in rng.permutation(N)[:N // 2])
    elif how == "half":  # the whole second half frozen: level n-1 is never in its g phase
        fs = set(range(N // 2, N)) | {0}
    elif how == "tail":  # only the last quarter carries information
        fs = set(range(3 * N // 4))
    else:
        z = [0.5]
        for _ in range(n):
            z = [v for zz in z for v in (2 * zz - zz * zz, zz * zz)]
        fs = set(int(i) for i in np.argsort(-np.array(z), kind="stable")[:N // 2])
    ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
    p = 0.08
    tab = np.full((q, q), p / (q - 1))
    np.fill_diagonal(tab, 1.0 - p)
    B = 70
    info = rng.integers(0, q, size=(B, ed.k))
    cw = ed.encode_batch(info)
    err = rng.random(cw.shape) < p
    y = np.where(err, (cw + rng.integers(1, q, size=cw.shape)) % q, cw).astype(np.uint8)
    y[0, : min(N, 5)] = 200  # out of range
    c1, i1 = ed.decode_symbols_batch(y, tab, return_codeword=True)
    monkeypatch.setenv("PC_QSC_LUT", "0")
    c0, i0 = ed.decode_symbols_batch(y, tab, return_codeword=True)
    monkeypatch.delenv("PC_QSC_LUT")
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(c1, c0)
    xy = tab[np.minimum(y[1:], q - 1)]
    c2, i2 = ed.decode_batch(xy, return_codeword=True)
    np.testing.assert_array_equal(i1[1:], i2)
    np.testing.assert_array_equal(c1[1:], c2)
    _, oinfo = oracle.q_decode_batch(q, N, ed.frozenMask, np.full((N, q), 1.0 / q), xy[:16])
    np.testing.assert_array_equal(i1[1:17], oinfo)


@pytest.mark.parametrize("q", [2, 3, 5, 8])
def test_fused_sweeps_every_depth_vs_oracle(q):
    """The fused two-level sweeps of qsc_decode_kernel (sc_qary.cu: fused_q) against the oracle on every block length 2 .. 512 with
    random frozen sets of several rates: descents of every length start at every level (g or f first), end at every stop level
    (rate-0 sub-trees) and cross the shared-memory / global-scratch boundary of each q; probability and symbol entry points."""
    import polarcub_b200 as pcb
    rng = np.random.default_rng(4242 + q)
    p = 0.07
    tab = np.full((q, q), p / (q - 1))
    np.fill_diagonal(tab, 1.0 - p)
    for n in range(1, 10):
        N = 1 << n
        for rate in (0.25, 0.5, 0.8):
            k = max(1, int(rate * N))
            fs = set(int(i) for i in rng.permutation(N)[:N - k])
            ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
            B = 40
            info = rng.integers(0, q, size=(B, ed.k))
            cw = ed.encode_batch(info)
            err = rng.random(cw.shape) < p
            y = np.where(err, (cw + rng.integers(1, q, size=cw.shape)) % q, cw).astype(np.uint8)
            xp = np.full((N, q), 1.0 / q)
            ocw, oinfo = oracle.q_decode_batch(q, N, ed.frozenMask, xp, tab[y])
            c1, i1 = ed.decode_batch(tab[y], return_codeword=True)
            np.testing.assert_array_equal(i1, oinfo, err_msg="probs n=%d rate=%s" % (n, rate))
            np.testing.assert_array_equal(c1, ocw, err_msg="probs n=%d rate=%s" % (n, rate))
            c2, i2 = ed.decode_symbols_batch(y, tab, return_codeword=True)
            np.testing.assert_array_equal(i2, oinfo, err_msg="symbols n=%d rate=%s" % (n, rate))
            np.testing.assert_array_equal(c2, ocw, err_msg="symbols n=%d rate=%s" % (n, rate))

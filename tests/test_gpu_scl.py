"""GPU parity: SC-list decoding (pc_scl_decode_probs through the C-ABI) vs the reference goldens and the oracle.

Bar: returned information and ProbResult identical; the final list identical IN ORDER with float64-equal
normalised metrics and genie metric (inputs are continuous-valued, so metrics are tie-free; see the oracle
header for the candidate-order convention).
"""
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


def test_golden_scl(golden_dir):
    import polarcub_b200 as pcb
    g = np.load(os.path.join(golden_dir, "scl.npz"))
    bad = []
    for nm in [str(s) for s in g["names"]]:
        q, n, L = int(g[nm + "/q"]), int(g[nm + "/n"]), int(g[nm + "/L"])
        N = 1 << n
        fs = set(np.nonzero(g[nm + "/frozen"])[0].tolist())
        ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
        info, res, lst = ed.listDecode_batch(g[nm + "/xy"], g[nm + "/fv"], L, g[nm + "/ainfo"], return_list=True)
        for f in range(info.shape[0]):
            ls = int(g[nm + "/lsize"][f])
            ok = (np.array_equal(info[f], g[nm + "/info"][f]) and int(res[f]) == int(g[nm + "/pr"][f])
                  and int(lst["list_size"][f]) == ls
                  and np.array_equal(lst["list_info"][f][:ls], g[nm + "/linfo"][f][:ls])
                  and np.array_equal(lst["list_prob"][f][:ls], g[nm + "/lprob"][f][:ls])
                  and float(lst["actual_prob"][f]) == float(g[nm + "/aprob"][f]))
            if not ok:
                bad.append((nm, f))
    assert not bad, bad


def test_reference_style_call(golden_dir):
    import polarcub_b200 as pcb
    from polarcub_b200.VectorDistributions.QaryMemorylessVectorDistribution import QaryMemorylessVectorDistribution
    g = np.load(os.path.join(golden_dir, "scl.npz"))
    nm = "q2_n6_L8_awgn"
    N = 64
    fs = set(np.nonzero(g[nm + "/frozen"])[0].tolist())
    ed = pcb.QaryPolarEncoderDecoder(2, N, fs, 1)
    for f in range(3):
        vd = QaryMemorylessVectorDistribution(2, N)
        vd.probs[:] = g[nm + "/xy"][f]
        info, pr = ed.listDecode(vd, g[nm + "/fv"][f], 8, np.zeros((ed.k, 0), dtype=np.int64), np.zeros(0, dtype=np.int64),
                                 actualInformation=g[nm + "/ainfo"][f])
        assert isinstance(pr, pcb.ProbResult) and pr.value == int(g[nm + "/pr"][f])
        assert info.dtype == np.int64
        np.testing.assert_array_equal(info, g[nm + "/info"][f])


@pytest.mark.parametrize("q,n,L,B", [(2, 8, 8, 70), (2, 10, 8, 40), (2, 12, 8, 33), (2, 9, 32, 20), (2, 7, 1, 64),
                                     (3, 7, 8, 40), (3, 9, 4, 20), (5, 5, 4, 20)])
def test_random_frames_vs_oracle(q, n, L, B):
    import polarcub_b200 as pcb
    N = 1 << n
    rng = np.random.default_rng(900 + 17 * n + q + L)
    k = N // 2
    fs = set(int(i) for i in _bec_order(n)[:N - k])
    ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
    fm = ed.frozenMask
    info = rng.integers(0, q, size=(B, k))
    fv = rng.integers(0, q, size=(B, N - k)) if q > 2 else np.zeros((B, N - k), dtype=np.int64)
    u = np.zeros((B, N), dtype=np.int64)
    u[:, fm == 0] = info
    u[:, fm == 1] = fv
    cw = np.stack([oracle.polar_transform_qudits(q, u[b]) for b in range(B)])  # encode == inverse map (involution)
    if q == 2:
        sigma = 0.75
        y = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
        l0, l1 = -(y - 1) ** 2 / (2 * sigma ** 2), -(y + 1) ** 2 / (2 * sigma ** 2)
        m = np.maximum(l0, l1)
        xy = np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=-1)
    else:
        p = 0.08
        err = rng.random((B, N)) < p
        yv = np.where(err, (cw + rng.integers(1, q, size=(B, N))) % q, cw)
        xy = np.where(np.arange(q)[None, None, :] == yv[..., None], 1.0 - p, p / (q - 1)) * (1.0 + 0.2 * rng.random((B, N, q)))
    ginfo, gres, lst = ed.listDecode_batch(xy, fv, L, info, return_list=True)
    for b in range(B):
        oi, opr, ols, olinfo, olprob, oap = oracle.list_decode(q, N, L, fm, xy[b], fv[b], info[b], want_list=True)
        np.testing.assert_array_equal(ginfo[b], oi, err_msg="frame %d" % b)
        assert int(gres[b]) == opr, b
        assert int(lst["list_size"][b]) == ols, b
        np.testing.assert_array_equal(lst["list_info"][b][:ols], olinfo[:ols], err_msg="frame %d" % b)
        assert np.array_equal(lst["list_prob"][b][:ols], olprob[:ols]), b
        assert float(lst["actual_prob"][b]) == oap, b
    # without the list outputs the selected word and the classification must not change
    ginfo2, gres2 = ed.listDecode_batch(xy, fv, L, info)
    np.testing.assert_array_equal(ginfo2, ginfo)
    np.testing.assert_array_equal(gres2, gres)


def _pattern(n, how, rate, rng):
    N = 1 << n
    if how == "bec":
        return set(int(i) for i in _bec_order(n)[:N - int(rate * N)])
    if how == "random":
        return set(int(i) for i in rng.permutation(N)[:N - int(rate * N)])
    m = np.zeros(N, dtype=bool)  # long frozen / free runs: fast nodes of every kind and size, also very large ones
    pos = 0
    while pos < N:
        run = int(rng.choice([1, 2, 3, 4, 8, 16, 31, 32, 64, 128, 256, 512, 1024]))
        if rng.random() > rate:
            m[pos:pos + run] = True
        pos += run
    if m.all():
        m[-1] = False
    return set(np.nonzero(m)[0].tolist())


@pytest.mark.parametrize("n,L,how,rate,B", [(6, 3, "blocks", 0.5, 24), (7, 2, "random", 0.75, 24), (8, 16, "blocks", 0.8, 20),
                                            (9, 8, "bec", 0.9, 16), (10, 32, "blocks", 0.5, 8), (11, 4, "blocks", 0.25, 12),
                                            (11, 8, "random", 0.5, 8), (12, 16, "bec", 0.75, 6), (13, 8, "blocks", 0.6, 5),
                                            (13, 2, "bec", 0.95, 5), (5, 8, "blocks", 1.0, 16), (4, 32, "random", 0.5, 16)])
def test_binary_scl_shapes_vs_oracle(n, L, how, rate, B):
    """The frame-per-warp list decoder on irregular frozen patterns (fast nodes of every kind up to 1024 symbols, all-information
    codes, tiny blocks), list sizes 2 .. 32 and block lengths up to 2^13: everything the oracle reports must match."""
    import polarcub_b200 as pcb
    N = 1 << n
    rng = np.random.default_rng(12000 + 101 * n + L)
    fs = _pattern(n, how, rate, rng)
    ed = pcb.QaryPolarEncoderDecoder(2, N, fs, 1)
    fm, k = ed.frozenMask, ed.k
    info = rng.integers(0, 2, size=(B, k))
    fv = np.zeros((B, N - k), dtype=np.int64)
    if n <= 8:
        fv = rng.integers(0, 2, size=(B, N - k))  # non-zero frozen values too
    u = np.zeros((B, N), dtype=np.int64)
    u[:, fm == 0] = info
    u[:, fm == 1] = fv
    cw = np.stack([oracle.polar_transform_qudits(2, u[b]) for b in range(B)])
    sigma = 0.9
    y = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
    l0, l1 = -(y - 1) ** 2 / (2 * sigma ** 2), -(y + 1) ** 2 / (2 * sigma ** 2)
    m = np.maximum(l0, l1)
    xy = np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=-1)
    ginfo, gres, lst = ed.listDecode_batch(xy, fv, L, info, return_list=True)
    for b in range(B):
        oi, opr, ols, olinfo, olprob, oap = oracle.list_decode(2, N, L, fm, xy[b], fv[b], info[b], want_list=True)
        np.testing.assert_array_equal(ginfo[b], oi, err_msg="frame %d" % b)
        assert int(gres[b]) == opr, b
        assert int(lst["list_size"][b]) == ols, b
        np.testing.assert_array_equal(lst["list_info"][b][:ols], olinfo[:ols], err_msg="frame %d" % b)
        assert np.array_equal(lst["list_prob"][b][:ols], olprob[:ols]), b
        assert float(lst["actual_prob"][b]) == oap, b
    ginfo2, gres2 = ed.listDecode_batch(xy, fv, L, info)
    np.testing.assert_array_equal(ginfo2, ginfo)
    np.testing.assert_array_equal(gres2, gres)


def test_ir_vs_live_reference_golden(golden_dir):
    """QaryPolarEncoderDecoder.ir end to end (QaryPolarEncoderDecoder.py:841-858: syndrome, frozen values, listDecode with
    genie selection): keys and ProbResult identical to the live reference (oracle/gen_golden_ir.py), q = 2 and 3, L = 1..8."""
    import polarcub_b200 as pcb
    g = np.load(os.path.join(golden_dir, "ir.npz"))
    for nm in [str(s) for s in g["names"]]:
        q, n, L, p = int(g[nm + "/q"]), int(g[nm + "/n"]), int(g[nm + "/L"]), float(g[nm + "/p"])
        N = 1 << n
        fs = set(np.nonzero(g[nm + "/frozen"])[0].tolist())
        ed = pcb.QaryPolarEncoderDecoder(q, N, fs, 1)
        for f in range(g[nm + "/a"].shape[0]):
            jit = g[nm + "/jit"][f]

            def make_xy(bv, jit=jit):
                return np.where(np.arange(q)[None, :] == np.asarray(bv)[:, None], 1.0 - p, p / (q - 1)) * jit

            a_key, b_key, pr = ed.ir(np.copy(g[nm + "/a"][f]), np.copy(g[nm + "/b"][f]), make_xy, list_size=L, check_size=0)
            np.testing.assert_array_equal(a_key, g[nm + "/a_key"][f], err_msg="%s frame %d" % (nm, f))
            np.testing.assert_array_equal(b_key, g[nm + "/b_key"][f], err_msg="%s frame %d" % (nm, f))
            assert pr.value == int(g[nm + "/pr"][f]), (nm, f)


def _awgn_table(sigma, Y):
    import math
    ymax = 1.0 + 4.0 * sigma
    step = 2 * ymax / Y
    edges = -ymax + step * np.arange(Y + 1)
    edges[0], edges[-1] = -np.inf, np.inf
    cdf = np.vectorize(lambda x: 0.5 * (1.0 + math.erf(x / math.sqrt(2.0))))
    tab = np.stack([0.5 * (cdf((edges[1:] - 1) / sigma) - cdf((edges[:-1] - 1) / sigma)),
                    0.5 * (cdf((edges[1:] + 1) / sigma) - cdf((edges[:-1] + 1) / sigma))], axis=-1)
    return ymax, step, tab


@pytest.mark.parametrize("n,L,Y,B", [(10, 8, 256, 24), (8, 4, 16, 40), (6, 32, 64, 12), (12, 8, 256, 9)])
def test_symbol_input_vs_oracle(n, L, Y, B):
    """pc_scl_decode_symbols: channel output symbols + the channel table (makeQaryMemorylessVectorDistribution fused into the
    decoder, QaryMemorylessDistribution.py:757-776) must give what the oracle gives on the table rows; quantised BI-AWGN, where
    equal rows make equal metrics possible, so only the returned word and ProbResult are compared with the list outputs off,
    and the whole list where the oracle's list is tie-free."""
    import polarcub_b200 as pcb
    N = 1 << n
    rng = np.random.default_rng(4000 + n + L)
    k = N // 2
    fs = set(int(i) for i in _bec_order(n)[:N - k])
    ed = pcb.QaryPolarEncoderDecoder(2, N, fs, 1)
    fm = ed.frozenMask
    info = rng.integers(0, 2, size=(B, k))
    fv = np.zeros((B, N - k), dtype=np.int64)
    u = np.zeros((B, N), dtype=np.int64)
    u[:, fm == 0] = info
    cw = np.stack([oracle.polar_transform_qudits(2, u[b]) for b in range(B)])
    sigma = 0.8
    ymax, step, tab = _awgn_table(sigma, Y)
    y = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
    ys = np.clip(np.floor((y + ymax) / step), 0, Y - 1).astype(np.uint8)
    ginfo, gres, lst = ed.listDecode_symbols_batch(ys, tab, fv, L, info, return_list=True)
    xy = tab[ys]
    for b in range(B):
        oi, opr, ols, olinfo, olprob, oap = oracle.list_decode(2, N, L, fm, xy[b], fv[b], info[b], want_list=True)
        if len(set(olprob[:ols].tolist())) < ols:
            continue  # exact ties in the oracle's list: the candidate order is numpy-implementation-defined (oracle header)
        np.testing.assert_array_equal(ginfo[b], oi, err_msg="frame %d" % b)
        assert int(gres[b]) == opr, b
        assert int(lst["list_size"][b]) == ols, b
        np.testing.assert_array_equal(lst["list_info"][b][:ols], olinfo[:ols], err_msg="frame %d" % b)
        assert np.array_equal(lst["list_prob"][b][:ols], olprob[:ols]), b
        assert float(lst["actual_prob"][b]) == oap, b
    if Y < 256:  # a symbol outside the table is refused (host arrays are checked; device tensors are read as row Y-1)
        with pytest.raises(pcb.PolarcubError):
            ed.listDecode_symbols_batch(np.full((1, N), Y, dtype=np.uint8), tab, fv[:1], L, info[:1])


def test_c2_code_tal_vardy_frozen_set():
    """BASELINE config 2 itself: N=4096, K=2048 on the committed Tal-Vardy frozen set (the reference's degrade pass over the
    400-bin BI-AWGN, tests/golden/constructions), L=8, Eb/N0 = 2 dB and 1 dB: word, ProbResult and the whole final list against
    the oracle, through both channel inputs (float64 pairs and 256-level symbols)."""
    import math
    import polarcub_b200 as pcb
    from polarcub_b200.construction import frozen_set_from_pe, load_pe
    n, N, K, L, B = 12, 4096, 2048, 8, 10
    fs = frozen_set_from_pe(load_pe("biawgn_ebn02.0_n12_L100_pe.npy"), K)
    ed = pcb.QaryPolarEncoderDecoder(2, N, fs, 1)
    fm = ed.frozenMask
    rng = np.random.default_rng(20480)
    for ebn0 in (2.0, 1.0):
        sigma = math.sqrt(1.0 / (2.0 * 0.5 * 10.0 ** (ebn0 / 10.0)))
        info = rng.integers(0, 2, size=(B, K))
        fv = np.zeros((B, N - K), dtype=np.int64)
        u = np.zeros((B, N), dtype=np.int64)
        u[:, fm == 0] = info
        cw = np.stack([oracle.polar_transform_qudits(2, u[b]) for b in range(B)])
        y = (1.0 - 2.0 * cw) + sigma * rng.standard_normal((B, N))
        l0, l1 = -(y - 1) ** 2 / (2 * sigma ** 2), -(y + 1) ** 2 / (2 * sigma ** 2)
        m = np.maximum(l0, l1)
        xy = np.stack([np.exp(l0 - m), np.exp(l1 - m)], axis=-1)
        ginfo, gres, lst = ed.listDecode_batch(xy, fv, L, info, return_list=True)
        for b in range(B):
            oi, opr, ols, olinfo, olprob, oap = oracle.list_decode(2, N, L, fm, xy[b], fv[b], info[b], want_list=True)
            np.testing.assert_array_equal(ginfo[b], oi, err_msg="frame %d" % b)
            assert int(gres[b]) == opr and int(lst["list_size"][b]) == ols, b
            np.testing.assert_array_equal(lst["list_info"][b][:ols], olinfo[:ols], err_msg="frame %d" % b)
            assert np.array_equal(lst["list_prob"][b][:ols], olprob[:ols]) and float(lst["actual_prob"][b]) == oap, b
        ymax, step, tab = _awgn_table(sigma, 256)
        ys = np.clip(np.floor((y + ymax) / step), 0, 255).astype(np.uint8)
        sinfo, sres = ed.listDecode_symbols_batch(ys, tab, fv, L, info)
        for b in range(B):
            oi, opr, ols, olinfo, olprob, oap = oracle.list_decode(2, N, L, fm, tab[ys[b]], fv[b], info[b], want_list=True)
            if len(set(olprob[:ols].tolist())) < ols:
                continue
            np.testing.assert_array_equal(sinfo[b], oi, err_msg="symbols, frame %d" % b)
            assert int(sres[b]) == opr, b


def test_list_decode_without_actual_information():
    """listDecode without actualInformation (the reference crashes in that form, QaryPolarEncoderDecoder.py:569): the first
    list entry that passes the check matrix is returned, None as ProbResult."""
    import polarcub_b200 as pcb
    n, N, L = 6, 64, 8
    rng = np.random.default_rng(6)
    fs = set(int(i) for i in _bec_order(n)[:N // 2])
    ed = pcb.QaryPolarEncoderDecoder(2, N, fs, 1)
    fm, k = ed.frozenMask, ed.k
    info = rng.integers(0, 2, size=k)
    u = np.zeros(N, dtype=np.int64)
    u[fm == 0] = info
    cw = oracle.polar_transform_qudits(2, u)
    y = (1.0 - 2.0 * cw) + 0.6 * rng.standard_normal(N)
    l0, l1 = -(y - 1) ** 2 / 0.72, -(y + 1) ** 2 / 0.72
    xy = np.stack([np.exp(l0 - np.maximum(l0, l1)), np.exp(l1 - np.maximum(l0, l1))], axis=-1)
    cm = rng.integers(0, 2, size=(k, 8))
    out, pr = ed.listDecode(xy, np.zeros(N - k, dtype=np.int64), L, cm, info @ cm % 2)
    assert pr is None
    np.testing.assert_array_equal(out, info)

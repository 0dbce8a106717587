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

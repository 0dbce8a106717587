"""csrc/channel.cu through the C-ABI: guard bands in and out BIT-EXACT against the reference's Guardbands.py (through the host
mirror polarcub_b200/Guardbands.py, itself pinned on live-reference goldens in tests/test_host_logic.py and, randomised, in
tests/test_live_reference_host.py); the noise generators statistically (the reference draws from CPython's
Mersenne Twister symbol by symbol) and for the properties the design promises: determinism and batch-split invariance."""
import math

import numpy as np
import pytest
import torch

from polarcub_b200 import Guardbands, channels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,n0,xi,ones", [(5, 2, 0.1, 0), (8, 2, 0.1, 0), (8, 3, 0.25, 2), (6, 1, 0.0, 1), (4, 4, 0.1, 0),
                                           (3, 5, 0.1, 1), (10, 3, 0.1, 0), (7, 2, 0.5, 3)])
def test_add_guard_bands_bit_exact(n, n0, xi, ones):
    rng = np.random.default_rng(n * 100 + n0)
    B, N = 37, 1 << n
    enc = rng.integers(0, 2, (B, N)).astype(np.uint8)
    out = channels.add_guard_bands(torch.from_numpy(enc).to(DEV), n, n0, xi, ones).cpu().numpy()
    for f in range(B):
        want = Guardbands.addDeletionGuardBands([int(v) for v in enc[f]], n, n0, xi, ones)
        assert out.shape[1] == len(want) == channels.guard_band_length(n, n0, xi, ones)
        assert np.array_equal(out[f], np.asarray(want, dtype=np.uint8)), (n, n0, xi, ones, f)


@pytest.mark.parametrize("n,n0,xi,ones,delta", [(5, 2, 0.1, 0, 0.1), (8, 2, 0.1, 0, 0.1), (8, 3, 0.25, 1, 0.05), (6, 1, 0.1, 0, 0.3),
                                                 (4, 4, 0.1, 0, 0.2), (9, 2, 0.1, 2, 0.02)])
def test_remove_guard_bands_bit_exact(n, n0, xi, ones, delta):
    rng = np.random.default_rng(n * 10 + n0)
    B, N = 61, 1 << n
    enc = rng.integers(0, 2, (B, N)).astype(np.uint8)
    enc[0] = 0                      # an all-zero codeword: every sub-word trims to nothing
    enc[1, : N // 2] = 0            # one half empty
    guarded = channels.add_guard_bands(torch.from_numpy(enc).to(DEV), n, n0, xi, ones)
    recv, rlen = channels.deletion_channel(guarded, delta, seed=5, frame0=100)
    maxlen = (1 << min(n, n0)) + 2 * ones + 6
    sb, sl, ov = channels.remove_guard_bands(recv, rlen, n, n0, maxlen)
    recv, rlen, sb, sl = recv.cpu().numpy(), rlen.cpu().numpy(), sb.cpu().numpy(), sl.cpu().numpy()
    any_long = False
    for f in range(B):
        word = [int(v) for v in recv[f][: rlen[f]]]
        want = Guardbands.removeDeletionGuardBands(word, n, n0)
        assert len(want) == sl.shape[1]
        for t, w in enumerate(want):
            assert sl[f, t] == len(w), (f, t)
            any_long |= len(w) > maxlen
            assert np.array_equal(sb[f, t, : min(len(w), maxlen)], np.asarray(w[:maxlen], dtype=np.uint8)), (f, t)
            assert not sb[f, t, min(len(w), maxlen):].any()
    assert ov == any_long


def test_deletion_channel_properties():
    B, ln, p = 4096, 300, 0.1
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.integers(0, 2, (B, ln)).astype(np.uint8)).to(DEV)
    out, olen = channels.deletion_channel(x, p, seed=42, frame0=0)
    out2, olen2 = channels.deletion_channel(x, p, seed=42, frame0=0)
    assert torch.equal(out, out2) and torch.equal(olen, olen2)            # deterministic
    a, la = channels.deletion_channel(x[: B // 2].contiguous(), p, 42, 0)    # batch-split invariance
    b, lb = channels.deletion_channel(x[B // 2:].contiguous(), p, 42, B // 2)
    assert torch.equal(torch.cat([a, b]), out) and torch.equal(torch.cat([la, lb]), olen)
    o3, _ = channels.deletion_channel(x, p, seed=43)
    assert not torch.equal(o3, out)
    lens = olen.cpu().numpy()
    kept = lens.sum() / (B * ln)
    assert abs(kept - (1 - p)) < 5 * math.sqrt(p * (1 - p) / (B * ln))
    # every received word is a subsequence of its input, and the padding is zero
    xo, oo = x.cpu().numpy(), out.cpu().numpy()
    for f in range(0, B, 97):
        it = iter(xo[f])
        assert all(any(v == w for w in it) for v in oo[f][: lens[f]])
        assert not oo[f][lens[f]:].any()
    # p = 0 keeps everything, p = 1 deletes everything
    k0, l0 = channels.deletion_channel(x, 0.0, 1)
    assert torch.equal(k0, x) and int(l0.min()) == ln
    _, l1 = channels.deletion_channel(x, 1.0, 1)
    assert int(l1.max()) == 0


@pytest.mark.parametrize("name,joint", [("bsc", [[0.89, 0.11], [0.11, 0.89]]),
                                         ("bec", [[0.9, 0.0], [0.0, 0.9], [0.1, 0.1]]),
                                         ("qsc3", [[0.98 if x == y else 0.01 for x in range(3)] for y in range(3)])])
def test_dmc_statistics_and_invariance(name, joint):
    cond = channels.conditional_table(joint)  # [X][Y]
    X, Y = cond.shape
    B, N = 2048, 512
    rng = np.random.default_rng(7)
    xs = rng.integers(0, X, (B, N)).astype(np.uint8)
    x = torch.from_numpy(xs).to(DEV)
    y = channels.simulate_dmc(x, cond, seed=11, frame0=7)
    assert torch.equal(y, channels.simulate_dmc(x, cond, seed=11, frame0=7))
    h = B // 2
    two = torch.cat([channels.simulate_dmc(x[:h].contiguous(), cond, 11, 7), channels.simulate_dmc(x[h:].contiguous(), cond, 11, 7 + h)])
    assert torch.equal(two, y)
    ys = y.cpu().numpy()
    for xv in range(X):
        sel = xs == xv
        cnt = sel.sum()
        for yv in range(Y):
            emp = (ys[sel] == yv).mean()
            pth = cond[xv, yv]
            assert abs(emp - pth) <= 5 * math.sqrt(max(pth * (1 - pth), 1e-12) / cnt) + (0 if pth > 0 else 0), (name, xv, yv, emp, pth)
            if pth == 0:
                assert emp == 0
    if X == 2:  # packed binary input gives the same symbols
        packed = torch.from_numpy(np.packbits(xs, axis=1, bitorder="little").view(np.int32).copy()).to(DEV)
        assert torch.equal(channels.simulate_dmc(packed, cond, 11, 7, packed_bits=N), y)


def test_biawgn_statistics_and_quantiser():
    B, N, sigma, Y = 1024, 1024, 0.8, 256
    rng = np.random.default_rng(5)
    xs = rng.integers(0, 2, (B, N)).astype(np.uint8)
    x = torch.from_numpy(xs).to(DEV)
    ymax = 1.0 + 4.0 * sigma
    yq, yr = channels.simulate_biawgn(x, sigma, seed=3, frame0=0, levels=Y, ymax=ymax, want_real=True)
    yq2, yr2 = channels.simulate_biawgn(x, sigma, seed=3, frame0=0, levels=Y, ymax=ymax, want_real=True)
    assert torch.equal(yq, yq2) and torch.equal(yr, yr2)
    noise = (yr.cpu().numpy() - (1.0 - 2.0 * xs)) / sigma
    n = noise.size
    assert abs(noise.mean()) < 5 / math.sqrt(n) and abs(noise.var() - 1.0) < 5 * math.sqrt(2.0 / n)
    assert abs((noise ** 3).mean()) < 5 * math.sqrt(15.0 / n) and abs((noise ** 4).mean() - 3.0) < 5 * math.sqrt(96.0 / n)
    assert abs(np.corrcoef(noise[:, :-1].ravel(), noise[:, 1:].ravel())[0, 1]) < 5 / math.sqrt(n)
    step = 2 * ymax / Y
    want = np.clip(np.floor((yr.cpu().numpy() + ymax) / step), 0, Y - 1).astype(np.uint8)
    assert np.array_equal(yq.cpu().numpy(), want)
    tab = channels.biawgn_table(sigma, Y, ymax)
    assert tab.shape == (Y, 2) and abs(tab.sum() - 1.0) < 1e-12
    emp = np.bincount(yq.cpu().numpy()[xs == 0], minlength=Y) / (xs == 0).sum()
    assert np.abs(emp - 2 * tab[:, 0]).max() < 6 * math.sqrt(0.02 / (xs == 0).sum())


@pytest.mark.parametrize("bits,Y", [(1, 2), (2, 3), (2, 4), (4, 11)])
def test_pack_unpack_symbols_and_packed_host_decode(bits, Y):
    """pc_pack_symbols / pc_unpack_symbols round trip and layout (symbol i in bits [i*bits, (i+1)*bits) of the little-endian
    stream), and the packed host-buffer SC pipeline against the unpacked one."""
    rng = np.random.default_rng(bits * 10 + Y)
    B, N = 300, 1024
    ys = rng.integers(0, Y, (B, N)).astype(np.uint8)
    y = torch.from_numpy(ys).to(DEV)
    p = channels.pack_symbols(y, bits)
    assert p.shape == (B, N * bits // 8)
    want = np.zeros((B, N * bits // 8), dtype=np.uint8)
    per = 8 // bits
    for j in range(per):
        want |= (ys[:, j::per] << (j * bits)).astype(np.uint8)
    assert np.array_equal(p.cpu().numpy(), want)
    assert torch.equal(channels.unpack_symbols(p, bits), y)
    if Y <= 3:  # SC decode through pinned host buffers: packed == unpacked
        import polarcub_b200 as pcb
        from polarcub_b200 import engine
        n = 10
        fm = np.zeros(N, dtype=np.uint8)
        fm[rng.permutation(N)[: N // 2]] = 1
        plan = engine.Plan(2, n, fm, np.zeros(N, dtype=np.uint8), device=DEV)
        tab = np.array([[0.445, 0.055], [0.055, 0.445], [0.05, 0.05]])[:Y]
        outs = []
        for pb, src in ((0, y), (bits, p)):
            yh = src.cpu().pin_memory()
            cw = torch.empty((B, plan.Nw), dtype=torch.int32).pin_memory()
            info = torch.empty((B, plan.Kw), dtype=torch.int32).pin_memory()
            engine.sc_decode_symbols_host(plan, yh, tab, cw, info, chunk=64, packed_bits=pb)
            torch.cuda.synchronize()
            outs.append((cw.clone(), info.clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert pcb is not None

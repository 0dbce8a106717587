"""GPU parity: the polar encode butterfly (pc_encode_bits / pc_polar_transform_bits) -- the warp-per-frame register kernel
(2^10 <= N <= 2^15), the block kernel (2^16 <= N <= 2^20) and the generic frame-per-CTA kernel -- against the oracle's encoder (pinned on the reference goldens) and
against each other.  Bar: bit-exact."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _cases(cs):
    return [c if len(c) == 4 else c + (67,) for c in cs]


@pytest.mark.parametrize("n,rate,seed,B", _cases([(10, 0.5, 1), (10, 0.03, -1), (11, 0.9, 3), (12, 0.5, 1), (13, 0.31, 7), (14, 0.77, 1),
                                         (15, 0.5, 2), (10, 1.0, 1), (12, 0.0, 1), (9, 0.5, 1), (16, 0.5, 1),
                                         (16, 0.9, -1, 2500), (17, 0.5, 1, 37), (18, 0.8, 2, 21), (19, 0.25, 1, 11), (20, 0.8, 1, 5),
                                         (20, 1.0, 1, 3), (18, 0.0, -1, 9)]))
def test_encode_vs_oracle_and_cta_kernel(n, rate, seed, B, monkeypatch):
    import torch
    import polarcub_b200 as pcb
    from polarcub_b200 import engine
    N = 1 << n
    rng = np.random.default_rng(300 + n)
    K = int(rate * N)
    fs = set(int(i) for i in rng.permutation(N)[:N - K])
    ed = pcb.BinaryPolarEncoderDecoder(N, fs, seed)
    info = rng.integers(0, 2, size=(B, ed.k))
    cw = ed.encode_batch(info)
    ref = oracle.bin_encode_batch(N, ed.frozenMask, ed.randomlyGeneratedNumbers, np.full((N, 2), 0.5), info)
    np.testing.assert_array_equal(cw, ref)
    monkeypatch.setenv("PC_ENCODE_CTA", "1")
    np.testing.assert_array_equal(ed.encode_batch(info), ref)
    monkeypatch.delenv("PC_ENCODE_CTA")
    # polarTransformOfBits: x -> u is the same involution
    packed = torch.from_numpy(engine.pack_bits(cw).view(np.int32)).cuda().contiguous()
    u = engine.unpack_bits(engine.polar_transform_bits(n, packed).cpu().numpy(), N)
    uu = np.zeros((B, N), dtype=np.uint8)
    uu[:, ed.frozenMask == 0] = info
    uu[:, ed.frozenMask == 1] = ed.frozenValues[ed.frozenMask == 1]
    np.testing.assert_array_equal(u, uu)
    np.testing.assert_array_equal(u[0], oracle.polar_transform_bits(cw[0]))

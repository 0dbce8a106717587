"""Channel simulation and guard-band plumbing on the device: the steps either side of the decoders in the reference's
Monte-Carlo loops (test2.py:29-65, test3.py:35-70, Guardbands.py:4-93, VectorDistributions/BinaryTrellis.py:441-461), as
batched calls into csrc/channel.cu.  Noise is drawn by a counter-based generator keyed by (seed, global frame index,
position): a frame's noise does not depend on the batch split or the number of ranks."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import PolarcubError


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise PolarcubError("device tensors expected (there is no CPU path)")


def conditional_table(joint):
    """P(y | x) rows [X][Y] from a reference-style table probs[y][x] (joint or conditional-per-row, e.g. makeQSC / makeBSC):
    probXGivenY(x, y) = probs[y][x] / sum_y' probs[y'][x] (QaryMemorylessDistribution.py / test3.py:45)."""
    j = np.asarray(joint, dtype=np.float64)
    col = j.sum(axis=0)
    return np.ascontiguousarray((j / col[None, :]).T)


def simulate_dmc(x, cond, seed, frame0=0, packed_bits=None):
    """x uint8 [B, N] input symbols (device) -- or, binary, packed_bits=N with x int32 [B, ceil(N/32)] -- through the discrete
    memoryless channel cond [X][Y] = P(y | x): uint8 [B, N] output symbols."""
    _need_cuda(x)
    cond = np.ascontiguousarray(cond, dtype=np.float64)
    X, Y = cond.shape
    B = x.shape[0]
    N = int(packed_bits) if packed_bits else x.shape[1]
    y = torch.empty((B, N), dtype=torch.uint8, device=x.device)
    ws = torch.empty((X * Y,), dtype=torch.float64, device=x.device)
    a = (None, _ptr(x)) if packed_bits else (_ptr(x), None)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().pc_channel_simulate_dmc(a[0], a[1], B, N, X, Y, cond.ctypes.data_as(ctypes.c_void_p),
                                                      ctypes.c_uint64(seed), int(frame0), _ptr(y), _ptr(ws), ws.numel() * 8,
                                                      _stream(x.device)), "pc_channel_simulate_dmc")
    return y


def simulate_biawgn(x, sigma, seed, frame0=0, levels=0, ymax=0.0, want_real=False, packed_bits=None):
    """BPSK over AWGN: y = (1 - 2 x) + sigma * N(0, 1).  Returns (quantised uint8 [B, N] or None, real float64 [B, N] or None);
    levels > 0 quantises to `levels` uniform bins on [-ymax, ymax] (the symbol input of the decoders)."""
    _need_cuda(x)
    B = x.shape[0]
    N = int(packed_bits) if packed_bits else x.shape[1]
    yq = torch.empty((B, N), dtype=torch.uint8, device=x.device) if levels else None
    yr = torch.empty((B, N), dtype=torch.float64, device=x.device) if (want_real or not levels) else None
    a = (None, _ptr(x)) if packed_bits else (_ptr(x), None)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().pc_channel_simulate_biawgn(a[0], a[1], B, N, float(sigma), ctypes.c_uint64(seed), int(frame0),
                                                         int(levels), float(ymax), _ptr(yq), _ptr(yr), _stream(x.device)),
                   "pc_channel_simulate_biawgn")
    return yq, yr


def biawgn_table(sigma, levels, ymax):
    """The [levels][2] joint table of the quantised BI-AWGN (uniform input): P(x) * P(y in bin | x), the outer bins open."""
    from scipy.stats import norm
    step = 2.0 * ymax / levels
    edges = -ymax + step * np.arange(levels + 1)
    edges[0], edges[-1] = -np.inf, np.inf
    return np.stack([0.5 * (norm.cdf((edges[1:] - 1) / sigma) - norm.cdf((edges[:-1] - 1) / sigma)),
                     0.5 * (norm.cdf((edges[1:] + 1) / sigma) - norm.cdf((edges[:-1] + 1) / sigma))], axis=-1)


def guard_band_length(n, n0, xi, ones=0):
    r = _lib.lib().pc_guard_band_length(int(n), int(n0), float(xi), int(ones))
    if r < 0:
        raise PolarcubError("bad guard-band parameters")
    return r


def add_guard_bands(encoded, n, n0, xi, ones=0):
    """Guardbands.addDeletionGuardBands on a batch: encoded uint8 [B, 2^n] (device) -> uint8 [B, guard_band_length]."""
    _need_cuda(encoded)
    assert encoded.dtype == torch.uint8 and encoded.is_contiguous() and encoded.shape[1] == (1 << n)
    B = encoded.shape[0]
    total = guard_band_length(n, n0, xi, ones)
    out = torch.empty((B, total), dtype=torch.uint8, device=encoded.device)
    ws = torch.empty((1 << max(n - n0, 0),), dtype=torch.int32, device=encoded.device)
    with torch.cuda.device(encoded.device):
        _lib.check(_lib.lib().pc_add_guard_bands(_ptr(encoded), B, int(n), int(n0), float(xi), int(ones), _ptr(out), _ptr(ws),
                                                 ws.numel() * 4, _stream(encoded.device)), "pc_add_guard_bands")
    return out


def deletion_channel(words, deletion_prob, seed, frame0=0):
    """deletionChannelSimulation on a batch: words uint8 [B, len] -> (received uint8 [B, len] zero padded, lengths int32 [B])."""
    _need_cuda(words)
    assert words.dtype == torch.uint8 and words.is_contiguous()
    B, ln = words.shape
    out = torch.empty_like(words)
    olen = torch.full((B,), ln if ln == 0 else 0, dtype=torch.int32, device=words.device)
    with torch.cuda.device(words.device):
        _lib.check(_lib.lib().pc_deletion_channel(_ptr(words), B, ln, float(deletion_prob), ctypes.c_uint64(seed), int(frame0),
                                                  _ptr(out), _ptr(olen), _stream(words.device)), "pc_deletion_channel")
    return out, olen


def remove_guard_bands(received, lengths, n, n0, maxlen):
    """Guardbands.removeDeletionGuardBands on a batch: received uint8 [B, stride], lengths int32 [B] (or None) ->
    (sub_bits uint8 [B, 2^(n-n0), maxlen], sub_len int32 [B, 2^(n-n0)], overflow bool): the inputs of the trellis decoder."""
    _need_cuda(received)
    assert received.dtype == torch.uint8 and received.is_contiguous()
    B, stride = received.shape
    T = 1 << max(n - n0, 0)
    sb = torch.empty((B, T, maxlen), dtype=torch.uint8, device=received.device)
    sl = torch.empty((B, T), dtype=torch.int32, device=received.device)
    ov = torch.zeros((1,), dtype=torch.int32, device=received.device)
    with torch.cuda.device(received.device):
        _lib.check(_lib.lib().pc_remove_guard_bands(_ptr(received), _ptr(lengths), B, stride, int(n), int(n0), int(maxlen), _ptr(sb),
                                                    _ptr(sl), _ptr(ov), _stream(received.device)), "pc_remove_guard_bands")
    return sb, sl, bool(int(ov.item()))


def symbol_bits(Y):
    """Bits per channel symbol of the packed host format for an alphabet of Y output symbols (1, 2 or 4; None above 16)."""
    return 1 if Y <= 2 else 2 if Y <= 4 else 4 if Y <= 16 else None


def pack_symbols(y, bits):
    """uint8 symbols [B, N] (device) -> packed uint8 [B, N * bits / 8] (device), `bits` in {1, 2, 4} per symbol."""
    _need_cuda(y)
    assert y.dtype == torch.uint8 and y.is_contiguous() and (y.shape[-1] * bits) % 32 == 0
    out = torch.empty(y.shape[:-1] + (y.shape[-1] * bits // 8,), dtype=torch.uint8, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().pc_pack_symbols(_ptr(y), int(bits), y.numel(), _ptr(out), _stream(y.device)), "pc_pack_symbols")
    return out


def unpack_symbols(packed, bits, out=None):
    """packed uint8 [B, N * bits / 8] (device) -> uint8 symbols [B, N]."""
    _need_cuda(packed)
    assert packed.dtype == torch.uint8 and packed.is_contiguous()
    n = packed.shape[-1] * 8 // bits
    if out is None:
        out = torch.empty(packed.shape[:-1] + (n,), dtype=torch.uint8, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(_lib.lib().pc_unpack_symbols(_ptr(packed), int(bits), packed.numel() * 8 // bits, _ptr(out), _stream(packed.device)),
                   "pc_unpack_symbols")
    return out

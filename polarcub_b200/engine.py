"""Batched array entry points over the C-ABI (torch tensors own the device memory, nothing else).

All functions launch on torch's current CUDA stream and return device tensors; packed bit tensors are
int32 views of the uint32 words described in include/polarcub_b200.h (bit i -> word i//32, bit i%32).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

INPUT_SYMBOLS = 0
INPUT_PROBS = 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_plan_device(fn):
    """Runs an entry point whose first argument is a Plan with the plan's device current: the plan's tables live there, and
    `_stream()` / `Plan.workspace()` then agree on that device's current stream whatever device the caller had selected."""
    import functools

    @functools.wraps(fn)
    def wrapper(plan, *a, **kw):
        with torch.cuda.device(plan.device):
            return fn(plan, *a, **kw)
    return wrapper


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def pack_bits(bits):
    """[..., n] array of 0/1 -> uint32 [..., ceil(n/32)] (LSB first)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    n = bits.shape[-1]
    W = (n + 31) // 32
    pad = W * 32 - n
    if pad:
        bits = np.concatenate([bits, np.zeros(bits.shape[:-1] + (pad,), dtype=np.uint8)], axis=-1)
    by = np.packbits(bits, axis=-1, bitorder="little")
    return np.ascontiguousarray(by).view("<u4").reshape(bits.shape[:-1] + (W,))


def unpack_bits(words, n):
    """uint32 [..., W] -> uint8 [..., n]."""
    words = np.ascontiguousarray(words).view(np.uint32)
    by = words.view(np.uint8).reshape(words.shape[:-1] + (words.shape[-1] * 4,))
    return np.unpackbits(by, axis=-1, bitorder="little")[..., :n]


class Plan:
    """Immutable (q, N, frozen set, frozen values) plan bound to one CUDA device (pc_plan_create)."""

    def __init__(self, q, n, frozen_mask, frozen_vals=None, device=None):
        if not torch.cuda.is_available():
            raise _lib.PolarcubError("polarcub_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.q, self.n, self.N = int(q), int(n), 1 << int(n)
        fm = np.ascontiguousarray(frozen_mask, dtype=np.uint8)
        assert fm.shape == (self.N,)
        fv = np.zeros(self.N, dtype=np.uint8) if frozen_vals is None else np.ascontiguousarray(frozen_vals, dtype=np.uint8)
        assert fv.shape == (self.N,)
        self.frozen_mask, self.frozen_vals = fm, fv
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = ctypes.c_void_p(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pc_plan_create(self.q, self.n, fm.ctypes.data_as(ctypes.c_void_p),
                                                 fv.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self._h)),
                       "pc_plan_create")
        self.k = _lib.lib().pc_plan_k(self._h)
        self.Nw = (self.N + 31) // 32
        self.Kw = (self.k + 31) // 32
        self._ws = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().pc_plan_destroy(self._h)
            self._h = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def schedule_len(self):
        return _lib.lib().pc_plan_schedule_len(self._h)

    def workspace(self, nbytes):
        """Device scratch for one call, private to the CUDA stream the call is launched on (calls on different streams may
        overlap, see `host_pipeline`)."""
        if self._ws is None:
            self._ws = {}
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            self._ws[key] = None
            ws = self._ws[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return ws


@_on_plan_device
def encode_bits(plan, info_packed):
    """info_packed int32 [B, Kw] (device) -> codewords int32 [B, Nw] (BinaryPolarEncoderDecoder.encode, uniform prior)."""
    B = info_packed.shape[0]
    assert info_packed.is_cuda and info_packed.dtype == torch.int32 and info_packed.is_contiguous()
    assert info_packed.shape[1] == plan.Kw
    cw = torch.empty((B, plan.Nw), dtype=torch.int32, device=info_packed.device)
    _lib.check(_lib.lib().pc_encode_bits(plan._h, _ptr(info_packed), _ptr(cw), B, _stream()), "pc_encode_bits")
    return cw


def polar_transform_bits(n, cw_packed):
    """x -> u (polarTransformOfBits) on packed words."""
    B = cw_packed.shape[0]
    assert cw_packed.is_cuda and cw_packed.dtype == torch.int32 and cw_packed.is_contiguous()
    u = torch.empty_like(cw_packed)
    _lib.check(_lib.lib().pc_polar_transform_bits(int(n), _ptr(cw_packed), _ptr(u), B, _stream()),
               "pc_polar_transform_bits")
    return u


@_on_plan_device
def sc_decode_probs(plan, xy, out=None):
    """xy float64 [B, N, 2] (device) -> (cw_packed int32 [B, Nw], info_packed int32 [B, Kw])."""
    assert xy.is_cuda and xy.dtype == torch.float64 and xy.is_contiguous() and xy.shape[1:] == (plan.N, 2)
    B = xy.shape[0]
    cw, info = out if out is not None else (torch.empty((B, plan.Nw), dtype=torch.int32, device=xy.device),
                                            torch.empty((B, max(plan.Kw, 1)), dtype=torch.int32, device=xy.device))
    need = _lib.lib().pc_sc_workspace_bytes(plan._h, B, INPUT_PROBS)
    ws = plan.workspace(need)
    _lib.check(_lib.lib().pc_sc_decode_probs(plan._h, _ptr(xy), B, _ptr(cw), _ptr(info), _ptr(ws), ws.numel(),
                                             _stream()), "pc_sc_decode_probs")
    return cw, info[:, :plan.Kw]


@_on_plan_device
def sc_decode_symbols(plan, y, table, out=None):
    """y uint8 [B, N] channel output symbols (device), table float64 [Y, 2] joint probabilities (host)."""
    assert y.is_cuda and y.dtype == torch.uint8 and y.is_contiguous() and y.shape[1] == plan.N
    table = np.ascontiguousarray(table, dtype=np.float64)
    assert table.ndim == 2 and table.shape[1] == 2 and 1 <= table.shape[0] <= 16
    B = y.shape[0]
    cw, info = out if out is not None else (torch.empty((B, plan.Nw), dtype=torch.int32, device=y.device),
                                            torch.empty((B, max(plan.Kw, 1)), dtype=torch.int32, device=y.device))
    need = _lib.lib().pc_sc_workspace_bytes_symbols(plan._h, B, table.ctypes.data_as(ctypes.c_void_p), table.shape[0])
    ws = plan.workspace(need)
    _lib.check(_lib.lib().pc_sc_decode_symbols(plan._h, _ptr(y), B, table.ctypes.data_as(ctypes.c_void_p),
                                               table.shape[0], _ptr(cw), _ptr(info), _ptr(ws), ws.numel(), _stream()),
               "pc_sc_decode_symbols")
    return cw, info[:, :plan.Kw]


@_on_plan_device
def qsc_encode(plan, info):
    """info uint8 [B, k] (device) -> codeword symbols uint8 [B, N] (QaryPolarEncoderDecoder.encode)."""
    assert info.is_cuda and info.dtype == torch.uint8 and info.is_contiguous() and info.shape[1] == plan.k
    B = info.shape[0]
    cw = torch.empty((B, plan.N), dtype=torch.uint8, device=info.device)
    _lib.check(_lib.lib().pc_qsc_encode(plan._h, _ptr(info), _ptr(cw), B, _stream()), "pc_qsc_encode")
    return cw


@_on_plan_device
def qsc_decode_probs(plan, xy, use_log=False):
    """xy float64 [B, N, q] (device) -> (cw uint8 [B, N], info uint8 [B, k]).  use_log: xy holds natural logarithms
    (QaryPolarEncoderDecoder(..., use_log=True))."""
    assert xy.is_cuda and xy.dtype == torch.float64 and xy.is_contiguous() and xy.shape[1:] == (plan.N, plan.q)
    B = xy.shape[0]
    cw = torch.empty((B, plan.N), dtype=torch.uint8, device=xy.device)
    info = torch.empty((B, max(plan.k, 1)), dtype=torch.uint8, device=xy.device)
    need = _lib.lib().pc_qsc_workspace_bytes(plan._h, B)
    ws = plan.workspace(need)
    fn = _lib.lib().pc_qsc_decode_logprobs if use_log else _lib.lib().pc_qsc_decode_probs
    _lib.check(fn(plan._h, _ptr(xy), B, _ptr(cw), _ptr(info), _ptr(ws), ws.numel(), _stream()),
               "pc_qsc_decode_logprobs" if use_log else "pc_qsc_decode_probs")
    return cw, info[:, :plan.k]


@_on_plan_device
def qsc_decode_symbols(plan, y, table, out=None, use_log=False):
    """y uint8 [B, N] channel output symbols (device), table float64 [Y, q] = QaryMemorylessDistribution.probs (host)
    -> (cw uint8 [B, N], info uint8 [B, k]).  use_log: `table` holds natural logarithms (-inf for 0)."""
    assert y.is_cuda and y.dtype == torch.uint8 and y.is_contiguous() and y.shape[1] == plan.N
    table = np.ascontiguousarray(table, dtype=np.float64)
    assert table.ndim == 2 and table.shape[1] == plan.q and 1 <= table.shape[0] <= 16
    B = y.shape[0]
    cw, info = out if out is not None else (torch.empty((B, plan.N), dtype=torch.uint8, device=y.device),
                                            torch.empty((B, max(plan.k, 1)), dtype=torch.uint8, device=y.device))
    ws = plan.workspace(_lib.lib().pc_qsc_workspace_bytes(plan._h, B))
    fn = _lib.lib().pc_qsc_decode_symbols_log if use_log else _lib.lib().pc_qsc_decode_symbols
    _lib.check(fn(plan._h, _ptr(y), B, table.ctypes.data_as(ctypes.c_void_p), table.shape[0],
                  _ptr(cw), _ptr(info), _ptr(ws), ws.numel(), _stream()), "pc_qsc_decode_symbols")
    return cw, info[:, :plan.k]


@_on_plan_device
def scl_decode_probs(plan, L, xy, frozen_values, actual_info, want_list=False, want_list_info=False, use_log=False):
    """SC-list decoding (QaryPolarEncoderDecoder.listDecode with genie selection) of a batch.

    xy float64 [B, N, q]; frozen_values uint8 [B, N-k]; actual_info uint8 [B, k] (all on the device).
    Returns dict(info uint8 [B,k], prob_result int32 [B]) plus, with want_list, list_size int32 [B],
    list_prob float64 [B,L], actual_prob float64 [B] and (want_list_info) list_info uint8 [B,L,k].
    use_log: xy, list_prob and actual_prob are natural logarithms (QaryPolarEncoderDecoder(..., use_log=True))."""
    assert xy.is_cuda and xy.dtype == torch.float64 and xy.is_contiguous() and xy.shape[1:] == (plan.N, plan.q)
    B = xy.shape[0]
    dev = xy.device
    nf = plan.N - plan.k
    assert frozen_values.dtype == torch.uint8 and frozen_values.shape == (B, nf) and frozen_values.is_contiguous()
    assert actual_info.dtype == torch.uint8 and actual_info.shape == (B, plan.k) and actual_info.is_contiguous()
    info = torch.empty((B, max(plan.k, 1)), dtype=torch.uint8, device=dev)
    res = torch.empty((B,), dtype=torch.int32, device=dev)
    ls = lp = ap = li = None
    if want_list or want_list_info:
        want_list = True
        ls = torch.empty((B,), dtype=torch.int32, device=dev)
        lp = torch.empty((B, L), dtype=torch.float64, device=dev)
        ap = torch.empty((B,), dtype=torch.float64, device=dev)
        if want_list_info:
            li = torch.empty((B, L, max(plan.k, 1)), dtype=torch.uint8, device=dev)
    wsf = _lib.lib().pc_scl_workspace_bytes_log if use_log else _lib.lib().pc_scl_workspace_bytes
    need = wsf(plan._h, int(L), B, 1 if want_list else 0)
    ws = plan.workspace(need)
    fn = _lib.lib().pc_scl_decode_logprobs if use_log else _lib.lib().pc_scl_decode_probs
    _lib.check(fn(plan._h, int(L), _ptr(xy), _ptr(frozen_values), _ptr(actual_info), B,
                  _ptr(info), _ptr(res), _ptr(ls), _ptr(lp), _ptr(ap), _ptr(li), _ptr(ws),
                  ws.numel(), _stream()), "pc_scl_decode_logprobs" if use_log else "pc_scl_decode_probs")
    out = {"info": info[:, :plan.k], "prob_result": res}
    if want_list:
        out.update(list_size=ls, list_prob=lp, actual_prob=ap)
        if want_list_info:
            out["list_info"] = li[:, :, :plan.k]
    return out


DEVICE_PIPELINE = True  # batches larger than one resident wave: wave-sized calls rotated over the three pipeline streams


@_on_plan_device
def _pipelined(plan, B, wave):
    """True when a device-resident batch should go out as wave-sized calls over the pipeline streams (not from inside one)."""
    return bool(DEVICE_PIPELINE and wave > 0 and B > wave and
                torch.cuda.current_stream(plan.device) not in _pipe_streams(plan.device))
# (measured on the SC decoders too: binary SC N=1024 loses 5 % -- its ingest / egress transposes compete with the DRAM-bound
# decode kernel -- and q-ary SC gains 1 %: only the list decoder, whose kernels have long tails, is pipelined.)


@_on_plan_device
def scl_decode_packed(plan, L, actual_info_packed, xy=None, y=None, table=None, frozen_packed=None, want_list=False,
                      want_list_info=False, out=None):
    """Binary SC-list decoding on bit-packed buffers (pc_scl_decode_packed).  Channel input: xy float64 [B, N, 2], or y uint8
    [B, N] output symbols + table float64 [Y, 2] (host).  actual_info_packed int32 [B, Kw]; frozen_packed int32 [B, ceil((N-k)/32)]
    or None (all-zero frozen values).  Returns dict(info_packed int32 [B, Kw], prob_result int32 [B]) plus the list outputs.

    A batch larger than one resident wave of the decode kernel goes out as wave-sized calls on three CUDA streams (the caller's
    stream waits for all of them): the preparation / selection kernels of one chunk and the tail of its decode kernel overlap
    the decode kernel of the next chunk, as in the host-buffer pipeline (host_pipeline)."""
    assert plan.q == 2
    B = actual_info_packed.shape[0]
    wave = scl_wave_frames(plan, L)
    if _pipelined(plan, B, wave):
        dev = actual_info_packed.device
        if out is not None:
            info, res = out
        else:
            info = torch.empty((B, max(plan.Kw, 1)), dtype=torch.int32, device=dev)
            res = torch.empty((B,), dtype=torch.int32, device=dev)
        ls = lp = ap = li = None
        if want_list or want_list_info:
            want_list = True
            ls = torch.empty((B,), dtype=torch.int32, device=dev)
            lp = torch.empty((B, L), dtype=torch.float64, device=dev)
            ap = torch.empty((B,), dtype=torch.float64, device=dev)
            if want_list_info:
                li = torch.empty((B, L, max(plan.Kw, 1)), dtype=torch.int32, device=dev)

        def body(lo, hi, slot):
            _scl_decode_packed_one(plan, L, actual_info_packed[lo:hi], None if xy is None else xy[lo:hi],
                                   None if y is None else y[lo:hi], table, None if frozen_packed is None else frozen_packed[lo:hi],
                                   want_list, info[lo:hi], res[lo:hi], None if ls is None else ls[lo:hi],
                                   None if lp is None else lp[lo:hi], None if ap is None else ap[lo:hi],
                                   None if li is None else li[lo:hi])

        host_pipeline(plan, B, wave, body)
        o = {"info_packed": info[:, :plan.Kw], "prob_result": res}
        if want_list:
            o.update(list_size=ls, list_prob=lp, actual_prob=ap)
            if want_list_info:
                o["list_info_packed"] = li[:, :, :plan.Kw]
        return o
    return _scl_decode_packed_single(plan, L, actual_info_packed, xy, y, table, frozen_packed, want_list, want_list_info, out)


@_on_plan_device
def _scl_decode_packed_one(plan, L, ai, xy, y, table, fvp, want_list, info, res, ls, lp, ap, li):
    """One pc_scl_decode_packed call on views of the caller's buffers, on the current stream."""
    B = ai.shape[0]
    tptr, Y = ctypes.c_void_p(0), 0
    if y is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
        tptr, Y = table.ctypes.data_as(ctypes.c_void_p), table.shape[0]
    need = _lib.lib().pc_scl_workspace_bytes_packed(plan._h, int(L), B, 1 if want_list else 0)
    ws = plan.workspace(need)
    _lib.check(_lib.lib().pc_scl_decode_packed(plan._h, int(L), _ptr(xy), _ptr(y), tptr, Y, _ptr(fvp), _ptr(ai), B, _ptr(info),
                                               _ptr(res), _ptr(ls), _ptr(lp), _ptr(ap), _ptr(li), _ptr(ws), ws.numel(), _stream()),
               "pc_scl_decode_packed")


@_on_plan_device
def _scl_decode_packed_single(plan, L, actual_info_packed, xy, y, table, frozen_packed, want_list, want_list_info, out):
    B = actual_info_packed.shape[0]
    dev = actual_info_packed.device
    assert actual_info_packed.is_cuda and actual_info_packed.dtype == torch.int32 and actual_info_packed.is_contiguous()
    assert actual_info_packed.shape == (B, max(plan.Kw, 1)) or actual_info_packed.shape == (B, plan.Kw)
    tptr, Y = ctypes.c_void_p(0), 0
    if xy is not None:
        assert y is None and xy.is_cuda and xy.dtype == torch.float64 and xy.is_contiguous() and xy.shape == (B, plan.N, 2)
    else:
        assert y.is_cuda and y.dtype == torch.uint8 and y.is_contiguous() and y.shape == (B, plan.N)
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.ndim == 2 and table.shape[1] == 2 and 1 <= table.shape[0] <= 256
        tptr, Y = table.ctypes.data_as(ctypes.c_void_p), table.shape[0]
    nfw = (plan.N - plan.k + 31) // 32
    if frozen_packed is not None:
        assert frozen_packed.is_cuda and frozen_packed.dtype == torch.int32 and frozen_packed.is_contiguous()
        assert frozen_packed.shape == (B, nfw)
    if out is not None:
        info, res = out
    else:
        info = torch.empty((B, max(plan.Kw, 1)), dtype=torch.int32, device=dev)
        res = torch.empty((B,), dtype=torch.int32, device=dev)
    ls = lp = ap = li = None
    if want_list or want_list_info:
        want_list = True
        ls = torch.empty((B,), dtype=torch.int32, device=dev)
        lp = torch.empty((B, L), dtype=torch.float64, device=dev)
        ap = torch.empty((B,), dtype=torch.float64, device=dev)
        if want_list_info:
            li = torch.empty((B, L, max(plan.Kw, 1)), dtype=torch.int32, device=dev)
    need = _lib.lib().pc_scl_workspace_bytes_packed(plan._h, int(L), B, 1 if want_list else 0)
    ws = plan.workspace(need)
    _lib.check(_lib.lib().pc_scl_decode_packed(plan._h, int(L), _ptr(xy), _ptr(y), tptr, Y, _ptr(frozen_packed),
                                               _ptr(actual_info_packed), B, _ptr(info), _ptr(res), _ptr(ls), _ptr(lp), _ptr(ap),
                                               _ptr(li), _ptr(ws), ws.numel(), _stream()), "pc_scl_decode_packed")
    o = {"info_packed": info[:, :plan.Kw], "prob_result": res}
    if want_list:
        o.update(list_size=ls, list_prob=lp, actual_prob=ap)
        if want_list_info:
            o["list_info_packed"] = li[:, :, :plan.Kw]
    return o


@_on_plan_device
def scl_decode_symbols_host(plan, L, y_host, table, ai_host, info_host, res_host, fv_host=None, chunk=None):
    """pc_scl_decode_symbols over a batch in pinned host memory: y_host uint8 [B, N] channel symbols, ai_host int32 [B, Kw]
    packed actual information, fv_host int32 [B, ceil((N-k)/32)] or None -> info_host int32 [B, Kw], res_host int32 [B]."""
    for t, nm in ((y_host, "y_host"), (ai_host, "ai_host"), (info_host, "info_host"), (res_host, "res_host")):
        _pinned(t, nm)
    if fv_host is not None:
        _pinned(fv_host, "fv_host")
    B = y_host.shape[0]
    nfw = (plan.N - plan.k + 31) // 32
    chunk = chunk or default_host_chunk(B, plan.N, scl_wave_frames(plan, L))
    sl = _Slots(plan, "sclsym")

    def body(lo, hi, slot):
        m = hi - lo
        y = sl.get(slot, "y", (chunk, plan.N), torch.uint8)[:m]
        ai = sl.get(slot, "ai", (chunk, max(plan.Kw, 1)), torch.int32)[:m]
        info = sl.get(slot, "info", (chunk, max(plan.Kw, 1)), torch.int32)[:m]
        res = sl.get(slot, "res", (chunk,), torch.int32)[:m]
        y.copy_(y_host[lo:hi], non_blocking=True)
        ai.copy_(ai_host[lo:hi], non_blocking=True)
        fv = None
        if fv_host is not None:
            fv = sl.get(slot, "fv", (chunk, nfw), torch.int32)[:m]
            fv.copy_(fv_host[lo:hi], non_blocking=True)
        scl_decode_packed(plan, L, ai, y=y, table=table, frozen_packed=fv, out=(info, res))
        info_host[lo:hi].copy_(info[:, :plan.Kw], non_blocking=True)
        res_host[lo:hi].copy_(res, non_blocking=True)

    host_pipeline(plan, B, chunk, body)


@_on_plan_device
def scl_decode_packed_host(plan, L, xy_host, ai_host, info_host, res_host, fv_host=None, chunk=None):
    """pc_scl_decode_packed over float64 probability pairs in pinned host memory (xy_host [B, N, 2]); packed side buffers as
    scl_decode_symbols_host."""
    for t, nm in ((xy_host, "xy_host"), (ai_host, "ai_host"), (info_host, "info_host"), (res_host, "res_host")):
        _pinned(t, nm)
    B = xy_host.shape[0]
    nfw = (plan.N - plan.k + 31) // 32
    chunk = chunk or default_host_chunk(B, plan.N * 16, scl_wave_frames(plan, L))
    sl = _Slots(plan, "sclpk")

    def body(lo, hi, slot):
        m = hi - lo
        xy = sl.get(slot, "xy", (chunk, plan.N, 2), torch.float64)[:m]
        ai = sl.get(slot, "ai", (chunk, max(plan.Kw, 1)), torch.int32)[:m]
        info = sl.get(slot, "info", (chunk, max(plan.Kw, 1)), torch.int32)[:m]
        res = sl.get(slot, "res", (chunk,), torch.int32)[:m]
        xy.copy_(xy_host[lo:hi], non_blocking=True)
        ai.copy_(ai_host[lo:hi], non_blocking=True)
        fv = None
        if fv_host is not None:
            fv = sl.get(slot, "fv", (chunk, nfw), torch.int32)[:m]
            fv.copy_(fv_host[lo:hi], non_blocking=True)
        scl_decode_packed(plan, L, ai, xy=xy, frozen_packed=fv, out=(info, res))
        info_host[lo:hi].copy_(info[:, :plan.Kw], non_blocking=True)
        res_host[lo:hi].copy_(res, non_blocking=True)

    host_pipeline(plan, B, chunk, body)


@_on_plan_device
def trellis_decode(plan, n0, deletion_prob, ones, sub_bits, sub_len, want_collapse=False):
    """Deletion-channel SC decoding (BinaryPolarEncoderDecoder.decode over a CollectionOfBinaryTrellises).

    sub_bits uint8 [B, T, maxlen], sub_len int32 [B, T] (device): the trimmed sub-words of each received word.
    Returns (cw_packed int32 [B, Nw], info_packed int32 [B, Kw]) and, with want_collapse, the first collapsed
    unnormalised vector float64 [B, T, 2]."""
    assert sub_bits.is_cuda and sub_bits.dtype == torch.uint8 and sub_bits.is_contiguous() and sub_bits.dim() == 3
    assert sub_len.is_cuda and sub_len.dtype == torch.int32 and sub_len.is_contiguous()
    B, T, maxlen = sub_bits.shape
    assert sub_len.shape == (B, T) and T == plan.N >> n0
    dev = sub_bits.device
    cw = torch.empty((B, plan.Nw), dtype=torch.int32, device=dev)
    info = torch.empty((B, max(plan.Kw, 1)), dtype=torch.int32, device=dev)
    col = torch.zeros((B, T, 2), dtype=torch.float64, device=dev) if want_collapse else None
    need = _lib.lib().pc_trellis_workspace_bytes(plan._h, int(n0), int(maxlen), B)
    ws = plan.workspace(need)
    _lib.check(_lib.lib().pc_trellis_decode(plan._h, int(n0), float(deletion_prob), int(ones), _ptr(sub_bits), _ptr(sub_len),
                                            int(maxlen), B, _ptr(cw), _ptr(info), _ptr(col), _ptr(ws), ws.numel(), _stream()),
               "pc_trellis_decode")
    return (cw, info[:, :plan.Kw], col) if want_collapse else (cw, info[:, :plan.Kw])


@_on_plan_device
def sc_genie_probs(plan, xy, u_packed):
    """Genie pass over memoryless inputs: xy float64 [B, N, 2], u_packed int32 [B, Nw] (the known u bits, device)
    -> (cw_packed int32 [B, Nw], marg float64 [B, N, 2])."""
    assert xy.is_cuda and xy.dtype == torch.float64 and xy.is_contiguous() and xy.shape[1:] == (plan.N, 2)
    B = xy.shape[0]
    assert u_packed.is_cuda and u_packed.dtype == torch.int32 and u_packed.is_contiguous() and u_packed.shape == (B, plan.Nw)
    cw = torch.empty((B, plan.Nw), dtype=torch.int32, device=xy.device)
    marg = torch.empty((B, plan.N, 2), dtype=torch.float64, device=xy.device)
    ws = plan.workspace(_lib.lib().pc_sc_genie_workspace_bytes(plan._h, B))
    _lib.check(_lib.lib().pc_sc_genie_probs(plan._h, _ptr(xy), _ptr(u_packed), B, _ptr(cw), _ptr(marg), _ptr(ws), ws.numel(),
                                            _stream()), "pc_sc_genie_probs")
    return cw, marg


def _rnd_arg(rnd, rows_per_frame, B, N):
    """randomlyGeneratedNumbers as a device float64 tensor + row stride: [N] shared (stride 0) or one row per frame."""
    assert rnd.is_cuda and rnd.dtype == torch.float64
    if rnd.dim() == 1:
        assert rnd.shape == (N,)
        return rnd.contiguous(), 0
    assert rnd.shape == (B, N)
    r = rnd.repeat_interleave(rows_per_frame, dim=0) if rows_per_frame > 1 else rnd
    return r.contiguous(), N


@_on_plan_device
def sc_decode_probs_prior(plan, xy, x, rnd, want_marg=False):
    """SC decoding under a non-uniform a-priori distribution (two trees in lock step, pc_sc_decode_probs_prior).
    xy float64 [B, N, 2]; x float64 [B, N, 2] or [N, 2]; rnd float64 [N] or [B, N] (all device).
    Returns (cw_packed int32 [B, Nw], info_packed int32 [B, Kw]) (+ marg_xy, marg_x float64 [B, N, 2] with want_marg)."""
    assert xy.is_cuda and xy.dtype == torch.float64 and xy.shape[1:] == (plan.N, 2)
    B = xy.shape[0]
    xx = x.expand(B, plan.N, 2) if x.dim() == 2 else x
    assert xx.shape == xy.shape and xx.dtype == torch.float64
    pairs = torch.stack([xy, xx], dim=1).reshape(2 * B, plan.N, 2).contiguous()
    r, stride = _rnd_arg(rnd, 2, B, plan.N)
    cw = torch.empty((2 * B, plan.Nw), dtype=torch.int32, device=xy.device)
    info = torch.zeros((2 * B, max(plan.Kw, 1)), dtype=torch.int32, device=xy.device)
    marg = torch.empty((2 * B, plan.N, 2), dtype=torch.float64, device=xy.device) if want_marg else None
    ws = plan.workspace(_lib.lib().pc_sc_genie_workspace_bytes(plan._h, 2 * B))
    _lib.check(_lib.lib().pc_sc_decode_probs_prior(plan._h, _ptr(pairs), _ptr(r), stride, 2 * B, _ptr(cw), _ptr(info), _ptr(marg),
                                                   _ptr(ws), ws.numel(), _stream()), "pc_sc_decode_probs_prior")
    out = (cw[0::2].contiguous(), info[0::2, :plan.Kw].contiguous())
    return out + (marg[0::2], marg[1::2]) if want_marg else out


@_on_plan_device
def sc_encode_prior(plan, x, u_packed, rnd, want_marg=False):
    """Encoding under a non-uniform a-priori distribution (pc_sc_encode_prior).  x float64 [B, N, 2]; u_packed int32
    [B, Nw] the information bits at their u positions; rnd float64 [N] or [B, N].  Returns cw_packed (+ marg [B, N, 2])."""
    assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.shape[1:] == (plan.N, 2)
    B = x.shape[0]
    assert u_packed.is_cuda and u_packed.dtype == torch.int32 and u_packed.is_contiguous() and u_packed.shape == (B, plan.Nw)
    r, stride = _rnd_arg(rnd, 1, B, plan.N)
    cw = torch.empty((B, plan.Nw), dtype=torch.int32, device=x.device)
    marg = torch.empty((B, plan.N, 2), dtype=torch.float64, device=x.device) if want_marg else None
    ws = plan.workspace(_lib.lib().pc_sc_genie_workspace_bytes(plan._h, B))
    _lib.check(_lib.lib().pc_sc_encode_prior(plan._h, _ptr(x), _ptr(u_packed), _ptr(r), stride, B, _ptr(cw), _ptr(marg), _ptr(ws),
                                             ws.numel(), _stream()), "pc_sc_encode_prior")
    return (cw, marg) if want_marg else cw


@_on_plan_device
def trellis_genie(plan, n0, deletion_prob, ones, sub_bits, sub_len, u_packed):
    """Genie pass over trellis collections (see trellis_decode / sc_genie_probs)."""
    assert sub_bits.is_cuda and sub_bits.dtype == torch.uint8 and sub_bits.is_contiguous() and sub_bits.dim() == 3
    assert sub_len.is_cuda and sub_len.dtype == torch.int32 and sub_len.is_contiguous()
    B, T, maxlen = sub_bits.shape
    assert sub_len.shape == (B, T) and T == plan.N >> n0
    assert u_packed.is_cuda and u_packed.dtype == torch.int32 and u_packed.is_contiguous() and u_packed.shape == (B, plan.Nw)
    dev = sub_bits.device
    cw = torch.empty((B, plan.Nw), dtype=torch.int32, device=dev)
    marg = torch.empty((B, plan.N, 2), dtype=torch.float64, device=dev)
    ws = plan.workspace(_lib.lib().pc_trellis_workspace_bytes(plan._h, int(n0), int(maxlen), B))
    _lib.check(_lib.lib().pc_trellis_genie(plan._h, int(n0), float(deletion_prob), int(ones), _ptr(sub_bits), _ptr(sub_len),
                                           int(maxlen), _ptr(u_packed), B, _ptr(cw), _ptr(marg), _ptr(ws), ws.numel(), _stream()),
               "pc_trellis_genie")
    return cw, marg


# ---- host-resident batches: chunked, copies overlapped with decoding ---------------------------------------------------
_PIPE_STREAMS = {}


PIPE_SLOTS = int(os.environ.get("PC_PIPE_SLOTS", "3"))


def _pipe_streams(device):
    key = torch.device(device).index
    if key not in _PIPE_STREAMS:
        _PIPE_STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(PIPE_SLOTS)]
    return _PIPE_STREAMS[key]


@_on_plan_device
def host_pipeline(plan, B, chunk, body):
    """Runs body(lo, hi, slot) for consecutive chunks [lo, hi) of a batch of B frames, rotating over PIPE_SLOTS (3) CUDA
    streams / staging slots: the H2D copies of one chunk overlap the decode kernel of the previous one and the D2H copies
    of the one before.  body must enqueue everything (copies from / to PINNED host tensors with non_blocking=True and the
    decode call) on the current stream and use per-slot device buffers.  Returns after enqueueing; the caller's
    stream waits for both."""
    cur = torch.cuda.current_stream(plan.device)
    streams = _pipe_streams(plan.device)
    for s in streams:
        s.wait_stream(cur)
    for j, lo in enumerate(range(0, B, chunk)):
        with torch.cuda.stream(streams[j % len(streams)]):
            body(lo, min(B, lo + chunk), j % len(streams))
    for s in streams:
        cur.wait_stream(s)


class _Slots:
    """Per-slot device staging buffers, cached on the plan."""

    def __init__(self, plan, tag):
        self.plan, self.tag = plan, tag
        if not hasattr(plan, "_slots"):
            plan._slots = {}

    def get(self, slot, name, shape, dtype):
        key = (self.tag, slot, name)
        t = self.plan._slots.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self.plan._slots[key] = torch.empty(shape, dtype=dtype, device=self.plan.device)
        return t


def _pinned(t, what):
    assert (not t.is_cuda) and t.is_pinned() and t.is_contiguous(), what + " must be a contiguous pinned host tensor"


def default_host_chunk(B, bytes_per_frame, wave=0, target_bytes=128 << 20):
    """Chunk size for host_pipeline: one resident wave of the decode kernel when known (every SM busy, nothing queued
    behind it), else ~128 MiB of input; at least 4 chunks when the batch allows it."""
    if wave > 0:
        return max(1, min(B, wave))
    c = max(1024, int(target_bytes // max(1, bytes_per_frame)))
    c = min(c, max(1024, (B + 3) // 4))
    return max(1, min(B, (c + 31) // 32 * 32))


@_on_plan_device
def sc_wave_frames(plan):
    return int(_lib.lib().pc_sc_wave_frames(plan._h))


@_on_plan_device
def scl_wave_frames(plan, L):
    return int(_lib.lib().pc_scl_wave_frames(plan._h, int(L)))


@_on_plan_device
def sc_decode_symbols_host(plan, y_host, table, cw_host, info_host, chunk=None, packed_bits=0):
    """pc_sc_decode_symbols over a batch in pinned host memory: y_host uint8 [B, N] -> cw_host int32 [B, Nw],
    info_host int32 [B, Kw] (pinned), copies overlapped with decoding.  packed_bits in {1, 2, 4}: y_host holds the symbols
    packed (uint8 [B, N * packed_bits / 8], channels.pack_symbols layout) -- 8 / 4 / 2 times less PCIe traffic for the small
    alphabets of BSC / BEC-type channels; they are unpacked on the device (pc_unpack_symbols) in front of the decoder."""
    _pinned(y_host, "y_host"), _pinned(cw_host, "cw_host"), _pinned(info_host, "info_host")
    B = y_host.shape[0]
    ycols = y_host.shape[1]
    assert ycols == (plan.N * packed_bits // 8 if packed_bits else plan.N)

    def stage(slots, slot, rows, lo, hi):
        """H2D copy of frames [lo, hi) into the slot's buffers; returns the uint8 [m, N] symbols on the device."""
        m = hi - lo
        if not packed_bits:
            y = slots.get(slot, "y", (rows, plan.N), torch.uint8)[:m]
            y.copy_(y_host[lo:hi], non_blocking=True)
            return y
        yp = slots.get(slot, "yp", (rows, ycols), torch.uint8)[:m]
        yp.copy_(y_host[lo:hi], non_blocking=True)
        y = slots.get(slot, "y", (rows, plan.N), torch.uint8)[:m]
        _lib.check(_lib.lib().pc_unpack_symbols(_ptr(yp), int(packed_bits), m * plan.N, _ptr(y), _stream()), "pc_unpack_symbols")
        return y

    if chunk is None and plan.n > 16 and B >= 6144:
        # large blocks (hybrid decoder): the workspace -- 5.4 MB per 2^20 frame, 1.9 MB over erasure channels -- is the
        # binding resource, so every batch is decoded on the caller's stream with its workspace; the H2D copy of the next
        # batch (256 KiB - 1 MiB of symbols per 2^20 frame: PCIe time comparable to the decode) runs on a side stream into the
        # other of two staging slots, and the D2H copies of the previous batch's results (228 KB per frame) on a third stream
        cb = B if B <= 16384 else 16384
        sl = _Slots(plan, "scsym_big")
        cur = torch.cuda.current_stream(plan.device)
        side, back = _pipe_streams(plan.device)[0], _pipe_streams(plan.device)[1]
        side.wait_stream(cur)
        back.wait_stream(cur)
        free = [None, None]     # the decode that read staging slot s has finished
        drained = [None, None]  # the D2H copies out of output slot s have finished
        for j, lo in enumerate(range(0, B, cb)):
            hi = min(B, lo + cb)
            s2 = j % 2
            with torch.cuda.stream(side):
                if free[s2] is not None:
                    side.wait_event(free[s2])  # the decode that read this slot two batches ago
                y = stage(sl, s2, cb, lo, hi)
                copied = torch.cuda.Event()
                copied.record(side)
            cur.wait_event(copied)
            if drained[s2] is not None:
                cur.wait_event(drained[s2])
            cw = sl.get(s2, "cw", (cb, plan.Nw), torch.int32)[:hi - lo]
            info = sl.get(s2, "info", (cb, max(plan.Kw, 1)), torch.int32)[:hi - lo]
            sc_decode_symbols(plan, y, table, out=(cw, info))
            free[s2] = torch.cuda.Event()
            free[s2].record(cur)
            with torch.cuda.stream(back):  # results go home while the next batch decodes
                back.wait_event(free[s2])
                cw_host[lo:hi].copy_(cw, non_blocking=True)
                info_host[lo:hi].copy_(info[:, :plan.Kw], non_blocking=True)
                drained[s2] = torch.cuda.Event()
                drained[s2].record(back)
        cur.wait_stream(back)
        cur.wait_stream(side)
        return
    chunk = chunk or default_host_chunk(B, plan.N, sc_wave_frames(plan))
    sl = _Slots(plan, "scsym")

    def body(lo, hi, slot):
        m = hi - lo
        cw = sl.get(slot, "cw", (chunk, plan.Nw), torch.int32)[:m]
        info = sl.get(slot, "info", (chunk, max(plan.Kw, 1)), torch.int32)[:m]
        y = stage(sl, slot, chunk, lo, hi)
        sc_decode_symbols(plan, y, table, out=(cw, info))
        cw_host[lo:hi].copy_(cw, non_blocking=True)
        info_host[lo:hi].copy_(info[:, :plan.Kw], non_blocking=True)

    host_pipeline(plan, B, chunk, body)


@_on_plan_device
def sc_decode_probs_host(plan, xy_host, cw_host, info_host, chunk=None):
    """pc_sc_decode_probs over a batch in pinned host memory: xy_host float64 [B, N, 2]."""
    _pinned(xy_host, "xy_host"), _pinned(cw_host, "cw_host"), _pinned(info_host, "info_host")
    B = xy_host.shape[0]
    chunk = chunk or default_host_chunk(B, plan.N * 16, sc_wave_frames(plan))
    sl = _Slots(plan, "scprob")

    def body(lo, hi, slot):
        m = hi - lo
        xy = sl.get(slot, "xy", (chunk, plan.N, 2), torch.float64)[:m]
        cw = sl.get(slot, "cw", (chunk, plan.Nw), torch.int32)[:m]
        info = sl.get(slot, "info", (chunk, max(plan.Kw, 1)), torch.int32)[:m]
        xy.copy_(xy_host[lo:hi], non_blocking=True)
        sc_decode_probs(plan, xy, out=(cw, info))
        cw_host[lo:hi].copy_(cw, non_blocking=True)
        info_host[lo:hi].copy_(info[:, :plan.Kw], non_blocking=True)

    host_pipeline(plan, B, chunk, body)


@_on_plan_device
def qsc_decode_probs_host(plan, xy_host, info_host, cw_host=None, chunk=None):
    """pc_qsc_decode_probs over a batch in pinned host memory: xy_host float64 [B, N, q] -> info_host uint8 [B, k]."""
    _pinned(xy_host, "xy_host"), _pinned(info_host, "info_host")
    B = xy_host.shape[0]
    chunk = chunk or default_host_chunk(B, plan.N * plan.q * 8, int(_lib.lib().pc_qsc_wave_frames(plan._h)))
    sl = _Slots(plan, "qsc")

    def body(lo, hi, slot):
        m = hi - lo
        xy = sl.get(slot, "xy", (chunk, plan.N, plan.q), torch.float64)[:m]
        xy.copy_(xy_host[lo:hi], non_blocking=True)
        cw, info = qsc_decode_probs(plan, xy)
        info_host[lo:hi].copy_(info, non_blocking=True)
        if cw_host is not None:
            cw_host[lo:hi].copy_(cw, non_blocking=True)

    host_pipeline(plan, B, chunk, body)


@_on_plan_device
def qsc_decode_symbols_host(plan, y_host, table, info_host, cw_host=None, chunk=None):
    """pc_qsc_decode_symbols over a batch in pinned host memory: y_host uint8 [B, N] -> info_host uint8 [B, k]."""
    _pinned(y_host, "y_host"), _pinned(info_host, "info_host")
    B = y_host.shape[0]
    chunk = chunk or default_host_chunk(B, plan.N, int(_lib.lib().pc_qsc_wave_frames(plan._h)))
    sl = _Slots(plan, "qscsym")

    def body(lo, hi, slot):
        m = hi - lo
        y = sl.get(slot, "y", (chunk, plan.N), torch.uint8)[:m]
        cw = sl.get(slot, "cw", (chunk, plan.N), torch.uint8)[:m]
        info = sl.get(slot, "info", (chunk, max(plan.k, 1)), torch.uint8)[:m]
        y.copy_(y_host[lo:hi], non_blocking=True)
        qsc_decode_symbols(plan, y, table, out=(cw, info))
        info_host[lo:hi].copy_(info[:, :plan.k], non_blocking=True)
        if cw_host is not None:
            cw_host[lo:hi].copy_(cw, non_blocking=True)

    host_pipeline(plan, B, chunk, body)


@_on_plan_device
def scl_decode_probs_host(plan, L, xy_host, fv_host, ai_host, info_host, res_host, chunk=None):
    """pc_scl_decode_probs over a batch in pinned host memory: xy_host float64 [B, N, q], fv_host uint8 [B, N-k],
    ai_host uint8 [B, k] -> info_host uint8 [B, k], res_host int32 [B]."""
    for t, nm in ((xy_host, "xy_host"), (fv_host, "fv_host"), (ai_host, "ai_host"), (info_host, "info_host"), (res_host, "res_host")):
        _pinned(t, nm)
    B = xy_host.shape[0]
    nf = plan.N - plan.k
    chunk = chunk or default_host_chunk(B, plan.N * plan.q * 8, scl_wave_frames(plan, L))
    sl = _Slots(plan, "scl")

    def body(lo, hi, slot):
        m = hi - lo
        xy = sl.get(slot, "xy", (chunk, plan.N, plan.q), torch.float64)[:m]
        fv = sl.get(slot, "fv", (chunk, nf), torch.uint8)[:m]
        ai = sl.get(slot, "ai", (chunk, plan.k), torch.uint8)[:m]
        xy.copy_(xy_host[lo:hi], non_blocking=True)
        fv.copy_(fv_host[lo:hi], non_blocking=True)
        ai.copy_(ai_host[lo:hi], non_blocking=True)
        o = scl_decode_probs(plan, L, xy, fv, ai)
        info_host[lo:hi].copy_(o["info"], non_blocking=True)
        res_host[lo:hi].copy_(o["prob_result"], non_blocking=True)

    host_pipeline(plan, B, chunk, body)


def kernel_launch_count():
    return int(_lib.lib().pc_kernel_launch_count())


def count_errors(a_packed, b_packed, nbits, out=None):
    """Accumulate {frames, frame errors, bit errors} (int64[3], device) over two packed-bit tensors."""
    assert a_packed.shape == b_packed.shape and a_packed.is_cuda and b_packed.is_cuda
    assert a_packed.is_contiguous() and b_packed.is_contiguous()
    if out is None:
        out = torch.zeros(3, dtype=torch.int64, device=a_packed.device)
    _lib.check(_lib.lib().pc_count_errors(_ptr(a_packed), _ptr(b_packed), a_packed.shape[0], int(nbits), _ptr(out),
                                          _stream()), "pc_count_errors")
    return out


def profile_enable(on=True):
    _lib.check(_lib.lib().pc_profile_enable(1 if on else 0), "pc_profile_enable")


def profile_read():
    ms, cnt = ctypes.c_double(0), ctypes.c_ulonglong(0)
    _lib.check(_lib.lib().pc_profile_read(ctypes.byref(ms), ctypes.byref(cnt)), "pc_profile_read")
    return ms.value, cnt.value

"""ctypes driver of tests/emu/libpc_emu.so (the SC-list kernel sources emulated on the CPU; test infrastructure)."""
import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_build.build())
        P = ctypes.c_void_p
        L.emu_sclp_decode.restype = ctypes.c_int
        L.emu_sclp_decode.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, P, ctypes.c_int, P, P, ctypes.c_int64, P, P, P, P, P, P]
        L.emu_last_error.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def list_decode(n, L, frozen_mask, fv, ainfo, xy=None, y=None, table=None, want_list=True, sms=2):
    """Returns dict(info, prob_result[, list_size, list_prob, actual_prob, list_info]) like engine.scl_decode_probs."""
    N = 1 << n
    fm = np.ascontiguousarray(frozen_mask, dtype=np.uint8)
    k = int(N - fm.sum())
    ainfo = np.ascontiguousarray(ainfo, dtype=np.uint8)
    B = ainfo.shape[0]
    fvb = None if fv is None else np.ascontiguousarray(fv, dtype=np.uint8)
    info = np.zeros((B, max(k, 1)), dtype=np.uint8)
    res = np.zeros(B, dtype=np.int32)
    ls = lp = ap = li = None
    if want_list:
        ls = np.zeros(B, dtype=np.int32)
        lp = np.zeros((B, L), dtype=np.float64)
        ap = np.zeros(B, dtype=np.float64)
        li = np.zeros((B, L, max(k, 1)), dtype=np.uint8)
    if xy is not None:
        xy = np.ascontiguousarray(xy, dtype=np.float64)
    else:
        y = np.ascontiguousarray(y, dtype=np.uint8)
        table = np.ascontiguousarray(table, dtype=np.float64)
    lb = lib()
    lb.emu_set_sms(int(sms))
    rc = lb.emu_sclp_decode(n, L, _p(fm), _p(xy), _p(y), _p(table), 0 if table is None else table.shape[0], _p(fvb), _p(ainfo), B,
                            _p(info), _p(res), _p(ls), _p(lp), _p(ap), _p(li))
    if rc != 0:
        raise RuntimeError("emu_sclp_decode failed: %d %s" % (rc, lb.emu_last_error().decode()))
    out = {"info": info[:, :k], "prob_result": res}
    if want_list:
        out.update(list_size=ls, list_prob=lp, actual_prob=ap, list_info=li[:, :, :k])
    return out

"""CPU oracle for the polar encode / SC / SCL hot path -- TEST INFRASTRUCTURE ONLY.

`oracle/` is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import it.  polarcub_b200/ never does.

The arithmetic lives in oracle/polar_oracle*.c (a restatement of the reference's float64
probability-domain recursion, cited line by line there); this module is the ctypes loader plus
numpy-typed wrappers.  Parity is PINNED against golden vectors generated from the live reference
(oracle/gen_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile oracle/libpolar_oracle.so with gcc (Makefile in this directory)."""
    so = os.path.join(_HERE, "libpolar_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = (not os.path.isfile(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libpolar_oracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpolar_oracle.so")
        if not os.path.isfile(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def frozen_mask(N, frozenSet):
    m = np.zeros(N, dtype=np.uint8)
    if len(frozenSet):
        m[np.fromiter(frozenSet, dtype=np.int64)] = 1
    return m


def common_randomness(N, seed):
    """randomlyGeneratedNumbers, BinaryPolarEncoderDecoder.py:33-44 (CPython MT19937 via the stdlib)."""
    import random
    if seed == -1:
        return np.ones(N, dtype=np.float64)
    rng = random.Random()
    rng.seed(seed)
    return np.array([rng.random() for _ in range(N)], dtype=np.float64)


# ---------------------------------------------------------------------------------------------------
def bin_encode(N, fmask, r, xprobs, info, want_marg=False):
    fmask, r, xprobs, info = _c(fmask, np.uint8), _c(r, np.float64), _c(xprobs, np.float64), _c(info, np.int64)
    cw = np.empty(N, dtype=np.int64)
    marg = np.empty((N, 2), dtype=np.float64) if want_marg else None
    rc = lib().po_bin_encode(ctypes.c_int(N), _p(fmask, _u8p), _p(r, _f64p), _p(xprobs, _f64p), _p(info, _i64p),
                             _p(cw, _i64p), _p(marg, _f64p))
    assert rc == 0, rc
    return (cw, marg) if want_marg else cw


def bin_decode(N, fmask, r, xprobs, xyprobs, want_marg=False, want_lvl1=False):
    fmask, r = _c(fmask, np.uint8), _c(r, np.float64)
    xprobs, xyprobs = _c(xprobs, np.float64), _c(xyprobs, np.float64)
    k = int(N - fmask.sum())
    cw = np.empty(N, dtype=np.int64)
    info = np.full(k, -1, dtype=np.int64)
    marg = np.empty((N, 2), dtype=np.float64) if want_marg else None
    l1m = np.empty((N // 2, 2), dtype=np.float64) if want_lvl1 and N > 1 else None
    l1p = np.empty((N // 2, 2), dtype=np.float64) if want_lvl1 and N > 1 else None
    rc = lib().po_bin_decode(ctypes.c_int(N), _p(fmask, _u8p), _p(r, _f64p), _p(xprobs, _f64p), _p(xyprobs, _f64p),
                             _p(cw, _i64p), _p(info, _i64p), _p(marg, _f64p), _p(l1m, _f64p), _p(l1p, _f64p))
    assert rc == 0, rc
    out = [cw, info]
    if want_marg:
        out.append(marg)
    if want_lvl1:
        out += [l1m, l1p]
    return tuple(out)


def bin_decode_batch(N, fmask, r, xprobs, xyprobs):
    fmask, r = _c(fmask, np.uint8), _c(r, np.float64)
    xprobs, xyprobs = _c(xprobs, np.float64), _c(xyprobs, np.float64)
    B = xyprobs.shape[0]
    k = int(N - fmask.sum())
    cw = np.empty((B, N), dtype=np.int64)
    info = np.full((B, k), -1, dtype=np.int64)
    rc = lib().po_bin_decode_batch(ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(k), _p(fmask, _u8p), _p(r, _f64p),
                                   _p(xprobs, _f64p), _p(xyprobs, _f64p), _p(cw, _i64p), _p(info, _i64p))
    assert rc == 0, rc
    return cw, info


def bin_encode_batch(N, fmask, r, xprobs, info):
    fmask, r, xprobs, info = _c(fmask, np.uint8), _c(r, np.float64), _c(xprobs, np.float64), _c(info, np.int64)
    B, k = info.shape
    cw = np.empty((B, N), dtype=np.int64)
    rc = lib().po_bin_encode_batch(ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(k), _p(fmask, _u8p), _p(r, _f64p),
                                   _p(xprobs, _f64p), _p(info, _i64p), _p(cw, _i64p))
    assert rc == 0, rc
    return cw


def polar_transform_bits(x):
    x = _c(x, np.int64)
    u = np.empty_like(x)
    rc = lib().po_polar_transform_bits(ctypes.c_int(x.shape[0]), _p(x, _i64p), _p(u, _i64p))
    assert rc == 0, rc
    return u


# ---------------------------------------------------------------------------------------------------
def q_encode(q, N, fmask, xprobs, info):
    fmask, xprobs, info = _c(fmask, np.uint8), _c(xprobs, np.float64), _c(info, np.int64)
    cw = np.empty(N, dtype=np.int64)
    rc = lib().po_q_encode(ctypes.c_int(q), ctypes.c_int(N), _p(fmask, _u8p), _p(xprobs, _f64p), _p(info, _i64p),
                           _p(cw, _i64p))
    assert rc == 0, rc
    return cw


def q_decode(q, N, fmask, xprobs, xyprobs, want_marg=False, want_lvl1=False, use_log=False):
    """use_log=True: xprobs / xyprobs are natural logarithms (QaryPolarEncoderDecoder(..., use_log=True))."""
    fmask, xprobs, xyprobs = _c(fmask, np.uint8), _c(xprobs, np.float64), _c(xyprobs, np.float64)
    k = int(N - fmask.sum())
    cw = np.empty(N, dtype=np.int64)
    info = np.full(k, -1, dtype=np.int64)
    marg = np.empty((N, q), dtype=np.float64) if want_marg else None
    l1m = np.empty((N // 2, q), dtype=np.float64) if want_lvl1 and N > 1 else None
    l1p = np.empty((N // 2, q), dtype=np.float64) if want_lvl1 and N > 1 else None
    fn = lib().po_q_decode_log if use_log else lib().po_q_decode
    rc = fn(ctypes.c_int(q), ctypes.c_int(N), _p(fmask, _u8p), _p(xprobs, _f64p), _p(xyprobs, _f64p),
            _p(cw, _i64p), _p(info, _i64p), _p(marg, _f64p), _p(l1m, _f64p), _p(l1p, _f64p))
    assert rc == 0, rc
    out = [cw, info]
    if want_marg:
        out.append(marg)
    if want_lvl1:
        out += [l1m, l1p]
    return tuple(out)


def q_decode_batch(q, N, fmask, xprobs, xyprobs):
    fmask, xprobs, xyprobs = _c(fmask, np.uint8), _c(xprobs, np.float64), _c(xyprobs, np.float64)
    B = xyprobs.shape[0]
    k = int(N - fmask.sum())
    cw = np.empty((B, N), dtype=np.int64)
    info = np.full((B, k), -1, dtype=np.int64)
    rc = lib().po_q_decode_batch(ctypes.c_int(B), ctypes.c_int(q), ctypes.c_int(N), ctypes.c_int(k), _p(fmask, _u8p),
                                 _p(xprobs, _f64p), _p(xyprobs, _f64p), _p(cw, _i64p), _p(info, _i64p))
    assert rc == 0, rc
    return cw, info


def polar_transform_qudits(q, x):
    x = _c(x, np.int64)
    u = np.empty_like(x)
    rc = lib().po_polar_transform_qudits(ctypes.c_int(q), ctypes.c_int(x.shape[0]), _p(x, _i64p), _p(u, _i64p))
    assert rc == 0, rc
    return u


# ---------------------------------------------------------------------------------------------------
def list_decode(q, N, L, fmask, xyprobs, frozen_values, actual_info, want_list=False, use_log=False):
    """QaryPolarEncoderDecoder.listDecode with actualInformation (genie selection); use_log=True: log-domain inputs and metrics.
    Returns (information[k], ProbResult value) and, with want_list, (list_size, list_info[L,k], list_prob[L], actual_prob)."""
    fmask, xyprobs = _c(fmask, np.uint8), _c(xyprobs, np.float64)
    fv, ai = _c(frozen_values, np.int64), _c(actual_info, np.int64)
    k = int(N - fmask.sum())
    info = np.empty(k, dtype=np.int64)
    pr = ctypes.c_int(-1)
    ls = ctypes.c_int(0)
    linfo = np.full((L, k), -1, dtype=np.int64)
    lprob = np.zeros(L, dtype=np.float64)
    ap = ctypes.c_double(0)
    fn = lib().po_list_decode_log if use_log else lib().po_list_decode
    rc = fn(ctypes.c_int(q), ctypes.c_int(N), ctypes.c_int(L), _p(fmask, _u8p), _p(xyprobs, _f64p),
            _p(fv, _i64p), _p(ai, _i64p), _p(info, _i64p), ctypes.byref(pr), ctypes.byref(ls),
            _p(linfo, _i64p), _p(lprob, _f64p), ctypes.byref(ap))
    assert rc == 0, rc
    if want_list:
        return info, pr.value, ls.value, linfo, lprob, ap.value
    return info, pr.value


def list_decode_batch(q, N, L, fmask, xyprobs, frozen_values, actual_info):
    fmask, xyprobs = _c(fmask, np.uint8), _c(xyprobs, np.float64)
    fv, ai = _c(frozen_values, np.int64), _c(actual_info, np.int64)
    B = xyprobs.shape[0]
    k = int(N - fmask.sum())
    info = np.empty((B, k), dtype=np.int64)
    pr = np.empty(B, dtype=np.int32)
    rc = lib().po_list_decode_batch(ctypes.c_int(B), ctypes.c_int(q), ctypes.c_int(N), ctypes.c_int(L), _p(fmask, _u8p),
                                    _p(xyprobs, _f64p), _p(fv, _i64p), _p(ai, _i64p), _p(info, _i64p), _p(pr, _i32p))
    assert rc == 0, rc
    return info, pr


# ---------------------------------------------------------------------------------------------------
# deletion channel: guard bands + trellis decoding (polar_oracle_trellis.c)
def add_guard_bands(encoded, n, n0, xi, ones=0):
    """Guardbands.addDeletionGuardBands (Guardbands.py:4-44) -> uint8 array."""
    enc = _c(encoded, np.uint8)
    assert enc.shape == (1 << n,)
    out = np.zeros(8 * (1 << n) + 2 * ones * (1 << n) + 64, dtype=np.uint8)
    m = lib().po_add_guard_bands(_p(enc, _u8p), ctypes.c_int(n), ctypes.c_int(n0), ctypes.c_double(xi), ctypes.c_int(ones),
                                 _p(out, _u8p))
    return out[:m].copy()


def remove_guard_bands(received, n, n0, maxlen):
    """Guardbands.removeDeletionGuardBands (Guardbands.py:47-63) -> (sub_bits uint8 [T, maxlen], sub_len int32 [T], overflow)."""
    rw = _c(received, np.uint8)
    T = 1 << (n - n0)
    sub_bits = np.zeros((T, maxlen), dtype=np.uint8)
    sub_len = np.zeros(T, dtype=np.int32)
    ov = lib().po_remove_guard_bands(_p(rw, _u8p), ctypes.c_int(rw.shape[0]), ctypes.c_int(n), ctypes.c_int(n0),
                                     _p(sub_bits, _u8p), _p(sub_len, _i32p), ctypes.c_int(maxlen))
    return sub_bits, sub_len, bool(ov)


def trellis_decode(n, n0, fmask, r, sub_bits, sub_len, deletion_prob, ones=0, want_collapse=False):
    """BinaryPolarEncoderDecoder.decode over a CollectionOfBinaryTrellises built from the trimmed sub-words (uniform prior).
    Returns (codeword int64[N], information int64[k]) and, with want_collapse, the collapsed unnormalised vectors
    [2^n0, T, 2] in visiting order."""
    fmask, r = _c(fmask, np.uint8), _c(r, np.float64)
    sub_bits, sub_len = _c(sub_bits, np.uint8), _c(sub_len, np.int32)
    N, T = 1 << n, 1 << (n - n0)
    assert sub_bits.shape[0] == T and sub_len.shape == (T,)
    k = int(N - fmask.sum())
    cw = np.empty(N, dtype=np.int64)
    info = np.full(max(k, 1), -1, dtype=np.int64)
    col = np.zeros((1 << n0, T, 2), dtype=np.float64) if want_collapse else None
    rc = lib().po_trellis_decode(ctypes.c_int(n), ctypes.c_int(n0), _p(fmask, _u8p), _p(r, _f64p), _p(sub_bits, _u8p),
                                 _p(sub_len, _i32p), ctypes.c_int(sub_bits.shape[1]), ctypes.c_double(deletion_prob),
                                 ctypes.c_int(ones), _p(cw, _i64p), _p(info, _i64p), _p(col, _f64p))
    assert rc == 0, rc
    return (cw, info[:k], col) if want_collapse else (cw, info[:k])

"""ctypes binding of polarcub_b200/libpolarcub_b200.so (the C-ABI declared in include/polarcub_b200.h).

There is NO CPU fallback: if the CUDA library is missing or fails to load, importing this module's
`lib()` raises.  Build it with `python polarcub_b200/build.py` (or `__graft_entry__.build()`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpolarcub_b200.so")
_LIB = None

c_void_p, c_int, c_int64, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t

# name -> (restype, argtypes); every symbol declared in include/polarcub_b200.h
SIGNATURES = {
    "pc_version": (c_int, []),
    "pc_last_error": (ctypes.c_char_p, []),
    "pc_kernel_launch_count": (ctypes.c_ulonglong, []),
    "pc_plan_create": (c_int, [c_int, c_int, c_void_p, c_void_p, ctypes.POINTER(c_void_p)]),
    "pc_plan_destroy": (None, [c_void_p]),
    "pc_plan_k": (c_int, [c_void_p]),
    "pc_plan_length": (c_int, [c_void_p]),
    "pc_plan_schedule_len": (c_int, [c_void_p]),
    "pc_encode_bits": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pc_polar_transform_bits": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "pc_sc_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int]),
    "pc_sc_workspace_bytes_symbols": (c_size_t, [c_void_p, c_int64, c_void_p, c_int]),
    "pc_sc_wave_frames": (c_int64, [c_void_p]),
    "pc_scl_wave_frames": (c_int64, [c_void_p, c_int]),
    "pc_sc_decode_probs": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_sc_decode_symbols": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "pc_qsc_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pc_qsc_wave_frames": (c_int64, [c_void_p]),
    "pc_qsc_workspace_bytes": (c_size_t, [c_void_p, c_int64]),
    "pc_qsc_decode_symbols": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_qsc_decode_probs": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_qsc_decode_logprobs": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_qsc_decode_symbols_log": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                          c_void_p]),
    "pc_scl_workspace_bytes_log": (c_size_t, [c_void_p, c_int, c_int64, c_int]),
    "pc_scl_decode_logprobs": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_scl_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int64, c_int]),
    "pc_scl_decode_probs": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_scl_workspace_bytes_packed": (c_size_t, [c_void_p, c_int, c_int64, c_int]),
    "pc_scl_decode_packed": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_scl_decode_symbols": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "pc_sc_genie_workspace_bytes": (c_size_t, [c_void_p, c_int64]),
    "pc_sc_genie_probs": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_sc_decode_probs_prior": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_size_t, c_void_p]),
    "pc_sc_encode_prior": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "pc_trellis_genie": (c_int, [c_void_p, c_int, ctypes.c_double, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_trellis_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int64]),
    "pc_trellis_decode": (c_int, [c_void_p, c_int, ctypes.c_double, c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_tv_degrade_pe": (c_int, [c_int, c_int, c_void_p, c_int, c_void_p, c_int]),
    "pc_tv_degrade_pe_qary": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int]),
    "pc_channel_simulate_dmc": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, ctypes.c_uint64, c_int64, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "pc_channel_simulate_biawgn": (c_int, [c_void_p, c_void_p, c_int64, c_int, ctypes.c_double, ctypes.c_uint64, c_int64, c_int,
                                           ctypes.c_double, c_void_p, c_void_p, c_void_p]),
    "pc_unpack_symbols": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "pc_pack_symbols": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "pc_guard_band_length": (c_int, [c_int, c_int, ctypes.c_double, c_int]),
    "pc_add_guard_bands": (c_int, [c_void_p, c_int64, c_int, c_int, ctypes.c_double, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pc_deletion_channel": (c_int, [c_void_p, c_int64, c_int, ctypes.c_double, ctypes.c_uint64, c_int64, c_void_p, c_void_p, c_void_p]),
    "pc_remove_guard_bands": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pc_count_errors": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "pc_profile_enable": (c_int, [c_int]),
    "pc_profile_read": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong)]),
}


class PolarcubError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.isfile(SO_PATH):
            raise PolarcubError(
                "polarcub_b200: CUDA library %s not built (run `python polarcub_b200/build.py`); "
                "there is no CPU fallback" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc, what=""):
    if rc != 0:
        msg = lib().pc_last_error()
        raise PolarcubError("%s failed with status %d: %s" % (what or "polarcub_b200 call", rc,
                                                              msg.decode() if msg else ""))

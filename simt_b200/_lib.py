"""ctypes binding of libsimt_b200.so (the C ABI declared in include/simt_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call
fails, the caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIMT_B200_LIB") or os.path.join(_HERE, "libsimt_b200.so")  # env override: A/B builds
_lib = None

# name -> (restype, argtypes); mirrors include/simt_b200.h one to one
SIGNATURES = {
    "simt_b200_abi_version": (c_int, []),
    "simt_b200_strerror": (c_char_p, [c_int]),
    "simt_b200_profile_enable": (None, [c_int]),
    "simt_b200_profile_read": (c_int, [ctypes.POINTER(c_double), ctypes.POINTER(c_longlong)]),
    "simt_head_workspace_bytes": (c_size_t, [c_int] * 7),
    "simt_head_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                              c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "simt_head_fwdbwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                 c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "simt_head_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                              c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "simt_head_scale": (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "simt_head_set_tuning": (None, [c_int, c_int, c_int, c_int]),
    "simt_nll2d_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "simt_nll2d_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "simt_confusion": (c_int, [c_void_p, c_int, c_void_p, c_int, c_longlong, c_void_p, c_int, c_int, c_void_p,
                               c_void_p, c_void_p]),
    "simt_class_hist": (c_int, [c_void_p, c_int, c_longlong, c_int, c_void_p, c_void_p]),
    "simt_label_map": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p]),
    "simt_hist_set_tuning": (None, [c_int, c_int, c_int]),
    "simt_t_regularizers": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "simt_xchg_bytes": (c_size_t, [c_int]),
    "simt_xchg_create": (c_int, [c_size_t, ctypes.POINTER(c_void_p), c_char_p]),
    "simt_xchg_open": (c_int, [c_char_p, ctypes.POINTER(c_void_p)]),
    "simt_xchg_close": (c_int, [c_void_p]),
    "simt_xchg_destroy": (c_int, [c_void_p]),
    "simt_xchg_set_timeout": (None, [c_longlong]),
    "simt_debug_resize_tables": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "simt_head_step_sharded": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                       c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_int, c_int, ctypes.POINTER(c_void_p), c_void_p, c_int, c_void_p]),
    "simt_head_finish_sharded": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_int, c_int, ctypes.POINTER(c_void_p), c_void_p]),
    "simt_head_step": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "simt_placeholder_fwdbwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "simt_w_fit": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_double, c_double,
                           c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "simt_anchor_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "simt_pseudo_labels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                   c_void_p, c_void_p, c_void_p]),
    "simt_eval_argmax": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p]),
    "simt_bilinear_gather": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                     c_void_p]),
}
OPTIONAL = set()


def load():
    """Load the library once; raises RuntimeError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the simt_b200 CUDA kernels are not built. "
            "Run `python -m simt_b200.build` (needs nvcc); there is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if name in OPTIONAL:
                continue
            raise RuntimeError(f"{LIB_PATH} does not export {name}; rebuild with `python -m simt_b200.build --force`")
        fn.restype = res
        fn.argtypes = args
    if lib.simt_b200_abi_version() != 1:
        raise RuntimeError("libsimt_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().simt_b200_strerror(code)
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else code} (code {code})")

"""ctypes binding of libmoondream_b200.so (the C-ABI declared in include/moondream_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``moondream_b200.build``.  There is
no CPU fallback: if the shared object is missing, importing this module's ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_longlong, c_void_p, c_float

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmoondream_b200.so")

_lib = None

# name -> (restype, [argtypes])
_SIGNATURES = {
    "md_last_error": (c_char_p, []),
    "md_abi_version": (c_int, []),
    "md_launch_count": (c_longlong, []),
    "md_reset_launch_count": (None, []),
    "md_linear_bf16": (
        c_int,
        [c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
         c_longlong, c_int, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p],
    ),
    "md_linear_small_batch_splits": (c_int, [c_int, c_int]),
    "md_linear_small_batch_workspace_bytes": (c_longlong, [c_int, c_int, c_int]),
    "md_linear_small_batch_bf16": (
        c_int,
        [c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
         c_longlong, c_void_p, c_longlong, c_void_p, c_void_p],
    ),
}


class NativeError(RuntimeError):
    pass


def exported_symbols():
    """Every symbol include/moondream_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is deliberately no CPU fallback)"
            )
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().md_last_error()
        raise NativeError(f"{what}: {msg.decode() if msg else 'error'}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def current_stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

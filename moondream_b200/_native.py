"""ctypes binding of libmoondream_b200.so (the C-ABI declared in include/moondream_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``moondream_b200.build``.  There is
no CPU fallback: if the shared object is missing, importing this module's ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_longlong, c_void_p, c_uint

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmoondream_b200.so")

_lib = None

class md_kv(ctypes.Structure):
    _fields_ = [("pool", c_void_p), ("n_pages", c_int), ("block_tables", c_void_p), ("max_blocks", c_int),
                ("n_layers", c_int), ("n_kv_heads", c_int)]


class md_dims(ctypes.Structure):
    _fields_ = [(n, c_int) for n in (
        "vis_dim", "vis_ff", "vis_layers", "vis_heads", "crop", "patch", "patch_k", "grid", "margin",
        "proj_inner", "txt_dim", "txt_ff", "txt_layers", "txt_heads", "vocab", "max_context",
        "prefix_len", "reg_inner", "coord_feat", "coord_out", "size_feat", "size_out", "txt_fused", "txt_kv_heads")]


_P = c_void_p
_LL = c_longlong
_KV = ctypes.POINTER(md_kv)
_DIMS = ctypes.POINTER(md_dims)

# name -> (restype, [argtypes]); must list every function include/moondream_b200.h declares
_SIGNATURES = {
    "md_last_error": (c_char_p, []),
    "md_abi_version": (c_int, []),
    "md_launch_count": (c_longlong, []),
    "md_reset_launch_count": (None, []),
    "md_profile_linear": (None, [c_int]),
    "md_debug_force_cta_group": (None, [c_int]),
    "md_profile_linear_read": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(c_longlong)]),
    "md_linear_bf16": (c_int, [_P, _LL, _P, _LL, c_int, c_int, c_int, c_int, _P, _P, _LL, c_int, _P, _LL,
                               c_int, c_int, c_int, _P]),
    "md_linear_small_batch_splits": (c_int, [c_int, c_int]),
    "md_linear_small_batch_workspace_bytes": (_LL, [c_int, c_int, c_int]),
    "md_linear_small_batch_bf16": (c_int, [_P, _LL, _P, _LL, c_int, c_int, c_int, c_int, _P, _P, _LL, _P,
                                           _LL, _P, _P]),
    "md_dequantize_weights": (c_int, [c_int, _P, _P, _P, c_int, c_int, _P, _LL, _P]),
    "md_linear_small_batch_quant": (c_int, [c_int, _P, _LL, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _LL, _P,
                                            _LL, _P, _P]),
    "md_resample_u8": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, _P]),
    "md_extract_windows_u8": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "md_patchify_u8": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "md_stitch_pool_concat_bf16": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "md_layernorm_bf16": (c_int, [_P, _LL, _P, _P, _P, _LL, c_int, c_int, _P]),
    "md_vit_attention_bf16": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "md_rope_kv_write_bf16": (c_int, [_P, c_int, c_int, _P, _P, c_int, _P, _P, _KV, c_int, _P]),
    "md_prefill_attention_bf16": (c_int, [_P, c_int, c_int, _P, _P, c_int, c_int, c_int, _KV, c_int, _P, _P]),
    "md_debug_attention_impl": (None, [c_int]),
    "md_debug_set_pdl": (None, [c_int]),
    "md_debug_skip_decode_kernels": (None, [c_int]),
    "md_debug_gemm": (None, [c_int]),
    "md_debug_gemm_sm_cap": (None, [c_int]),
    "md_debug_timeline": (c_int, [c_void_p, c_void_p, c_uint]),
    "md_decode_attention_bf16": (c_int, [_P, c_int, _P, c_int, _KV, c_int, _P, _P]),
    "md_model_num_weights": (c_int, [_DIMS]),
    "md_model_create": (c_int, [_DIMS, ctypes.POINTER(c_void_p), c_int, _P, _P, ctypes.POINTER(c_void_p)]),
    "md_model_destroy": (None, [_P]),
    "md_model_set_quantized_block": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "md_vision_encode_workspace_bytes": (_LL, [_P, c_int]),
    "md_vision_encode": (c_int, [_P, _P, c_int, _P, _P, _P]),
    "md_vision_project_workspace_bytes": (_LL, [_P, c_int]),
    "md_vision_project": (c_int, [_P, _P, _P, _P, c_int, _P, c_int, _P, _P]),
    "md_vision_project_stitched": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, _P]),
    "md_embed_tokens": (c_int, [_P, _P, _LL, c_int, _P, _LL, _P]),
    "md_text_prefill_workspace_bytes": (_LL, [_P, c_int]),
    "md_text_prefill": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, _KV, _P, _P]),
    "md_text_prefill_lora_workspace_bytes": (_LL, [_P, c_int, c_int]),
    "md_text_prefill_lora": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, _KV, ctypes.POINTER(c_void_p), c_int,
                                     _P, _P]),
    "md_text_decode_workspace_bytes": (_LL, [_P, c_int]),
    "md_text_decode_step": (c_int, [_P, _P, _P, c_int, _KV, _P, _P, _P]),
    "md_lm_head_workspace_bytes": (_LL, [_P, c_int]),
    "md_lm_head_argmax": (c_int, [_P, _P, _LL, c_int, c_int, c_int, c_int, _P, _LL, _P, _P, _P, _P, _P]),
    "md_sample_top_p": (c_int, [_P, c_int, c_int, ctypes.c_float, ctypes.c_float, _P, _P, _P, _P, c_int, _P, _LL,
                                c_int, _P]),
    "md_embed_tokens_select": (c_int, [_P, _P, _LL, c_int, c_int, _P, _LL, _P, _LL, _P]),
    "md_store_column_f32": (c_int, [_P, c_int, _P, _LL, _P, c_int, _P]),
    "md_decode_advance": (c_int, [_P, _P, _P, _P, _P, _LL, c_int, c_int, _P, _P]),
    "md_gather_rows_bf16": (c_int, [_P, _LL, _P, c_int, c_int, _P, _LL, _P]),
    "md_safetensors_open": (c_int, [c_char_p, ctypes.POINTER(c_void_p)]),
    "md_safetensors_count": (c_int, [_P]),
    "md_safetensors_info": (c_int, [_P, c_int, ctypes.POINTER(c_char_p), ctypes.POINTER(c_char_p), ctypes.POINTER(c_int),
                                    ctypes.POINTER(c_longlong), ctypes.POINTER(c_longlong)]),
    "md_safetensors_find": (c_int, [_P, c_char_p]),
    "md_safetensors_read": (c_int, [_P, c_int, _P, _LL, c_int, _P]),
    "md_safetensors_close": (None, [_P]),
    "md_region_workspace_bytes": (_LL, [_P, c_int]),
    "md_region_decode": (c_int, [_P, c_int, _P, _LL, c_int, _P, _P, _P]),
    "md_region_encode": (c_int, [_P, c_int, _P, c_int, _P, _LL, _P, _P]),
    "md_region_bins_to_values": (c_int, [c_int, _P, c_int, c_int, _P, _P]),
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

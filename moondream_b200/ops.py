"""Tensor-level wrappers over the C-ABI (torch is used only for device memory and streams)."""
from __future__ import annotations

import torch

from . import _native as N

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2


def _req(t: torch.Tensor, dtype=torch.bfloat16):
    if not t.is_cuda:
        raise N.NativeError("moondream_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise N.NativeError(f"expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise N.NativeError("innermost dimension must be contiguous")


def linear(x, w, bias=None, epilogue=EPI_BIAS, residual=None, res_mod=0, out=None, remap=(0, 0, 0)):
    """y = epilogue(x @ w.T + bias); x [M,K], w [N,K] bf16.  See md_linear_bf16."""
    _req(x), _req(w)
    M, K = x.shape
    Nout = w.shape[0]
    if out is None:
        out = torch.empty((M, Nout), device=x.device, dtype=torch.bfloat16)
    ldr = residual.stride(0) if residual is not None else 0
    rc = N.lib().md_linear_bf16(
        N.ptr(x), x.stride(0), N.ptr(w), w.stride(0), M, Nout, K, epilogue, N.ptr(bias),
        N.ptr(residual), ldr, res_mod, N.ptr(out), out.stride(0), remap[0], remap[1], remap[2],
        N.current_stream())
    N.check(rc, "md_linear_bf16")
    return out


def linear_small_batch(x, w, bias=None, epilogue=EPI_BIAS, residual=None, out=None, workspace=None):
    """Decode-sized batch: every CTA streams one (weight-row tile, K split) once; split-K partials are reduced
    in a fixed order by the epilogue kernel."""
    _req(x), _req(w)
    B, K = x.shape
    Nout = w.shape[0]
    if out is None:
        out = torch.empty((B, Nout), device=x.device, dtype=torch.bfloat16)
    need = N.lib().md_linear_small_batch_workspace_bytes(Nout, B, K)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need // 4, device=x.device, dtype=torch.float32)
    ldr = residual.stride(0) if residual is not None else 0
    rc = N.lib().md_linear_small_batch_bf16(
        N.ptr(x), x.stride(0), N.ptr(w), w.stride(0), B, Nout, K, epilogue, N.ptr(bias),
        N.ptr(residual), ldr, N.ptr(out), out.stride(0), N.ptr(workspace), N.current_stream())
    N.check(rc, "md_linear_small_batch_bf16")
    return out


def dequantize_weights(bits, wq, scale, zero, n_out, K, out=None):
    """bf16 [n_out, K] = bf16(bf16(q - zero) * scale) from stream-layout packed weights (md_dequantize_weights)."""
    _req(wq, torch.uint8), _req(scale, torch.float32), _req(zero, torch.float32)
    if out is None:
        out = torch.empty((n_out, K), device=wq.device, dtype=torch.bfloat16)
    N.check(N.lib().md_dequantize_weights(bits, N.ptr(wq), N.ptr(scale), N.ptr(zero), n_out, K, N.ptr(out), out.stride(0),
                                          N.current_stream()), "md_dequantize_weights")
    return out


def linear_small_batch_quant(bits, x, wq, scale, zero, n_out, bias=None, epilogue=EPI_BIAS, residual=None, out=None):
    """linear_small_batch with int4 (group 128) / int8 weights in the stream layout (md_linear_small_batch_quant)."""
    _req(x), _req(wq, torch.uint8), _req(scale, torch.float32), _req(zero, torch.float32)
    B, K = x.shape
    if out is None:
        out = torch.empty((B, n_out), device=x.device, dtype=torch.bfloat16)
    need = N.lib().md_linear_small_batch_workspace_bytes(n_out, B, K)
    workspace = torch.empty(need // 4, device=x.device, dtype=torch.float32)
    ldr = residual.stride(0) if residual is not None else 0
    N.check(N.lib().md_linear_small_batch_quant(
        bits, N.ptr(x), x.stride(0), N.ptr(wq), N.ptr(scale), N.ptr(zero), B, n_out, K, epilogue, N.ptr(bias),
        N.ptr(residual), ldr, N.ptr(out), out.stride(0), N.ptr(workspace), N.current_stream()),
        "md_linear_small_batch_quant")
    return out

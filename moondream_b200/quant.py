"""Weight-only int8 for the decode-time weight streams (BASELINE.json config 5, SURVEY.md §8c "int8" and §8f rank 3).

Groundwork only: the format, the quantiser and the dequantised weights the oracle runs on.  The reference has no
int8 path (only the torchao int4 `QuantizedLinear`, layers.py:38-110), so — as SURVEY.md §8c prescribes — parity for
int8 means: *the bf16 path run on the dequantised weights*.  The scheme is chosen so that a kernel can meet that
definition exactly:

  * symmetric, per output feature:  scale[n] = bf16(max_k |w[n, k]| / 127),  q[n, k] = clamp(rne(w[n, k] / scale[n]), -127, 127)
  * dequantised weight:             w'[n, k] = bf16(q[n, k] * scale[n])            (one rounding, per element)

A weight-stream kernel moves q (1 byte per weight: half the HBM traffic of bf16) with TMA, rewrites the tile in shared
memory as w' (int8 -> fp32 -> * scale -> bf16, the row's scale is a scalar) and issues the same bf16 MMAs as today, so
its output equals the bf16 kernels' output on w' bit for bit.  Not wired into the engine yet.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch

from .config import MoondreamConfig


def quantize_weight_int8(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """w [out, in] (any float dtype) -> (q int8 [out, in], scale bf16 [out])."""
    if w.dim() != 2:
        raise ValueError("quantize_weight_int8 expects a [out, in] matrix")
    wf = w.detach().to(torch.float32)
    amax = wf.abs().amax(dim=1)
    scale = (amax / 127.0).to(torch.bfloat16)
    # rows of zeros (or denormal scales): any non-zero scale reproduces the zeros
    scale = torch.where(scale.float() > 0, scale, torch.ones_like(scale))
    q = torch.round(wf / scale.float().unsqueeze(1)).clamp_(-127, 127).to(torch.int8)
    return q, scale


def dequantize_weight_int8(q: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """The bf16 operand a kernel reconstructs in shared memory: bf16(q * scale), one rounding per element."""
    return (q.to(torch.float32) * scale.to(torch.float32).unsqueeze(1)).to(torch.bfloat16)


def decode_stream_keys(cfg: MoondreamConfig) -> Iterable[str]:
    """The matrices a decode step streams from HBM (2.63 GB per step for the 2B): the decoder blocks'
    four Linear layers and the LM head.  Vision weights, embeddings and the region head stay bf16."""
    for i in range(cfg.text.n_layers):
        for name in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
            yield f"text.blocks.{i}.{name}.weight"
    yield "text.lm_head.weight"


def quantize_decoder_int8(cfg: MoondreamConfig, sd: Dict[str, torch.Tensor]):
    """-> (packed, dequantised): `packed[key] = (q, scale)` for every decode-stream matrix; `dequantised` is a full
    state dict (shared tensors for everything untouched) in which those matrices hold bf16(q * scale) — the weights
    the oracle, and the bf16 engine, must be run on to define int8 parity."""
    packed: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
    deq = dict(sd)
    for key in decode_stream_keys(cfg):
        q, scale = quantize_weight_int8(sd[key])
        packed[key] = (q, scale)
        deq[key] = dequantize_weight_int8(q, scale)
    return packed, deq


def stream_bytes(cfg: MoondreamConfig, int8: bool) -> int:
    """Weight bytes one decode step reads (the HBM-roofline numerator of SURVEY.md §8d), bf16 or int8 + scales."""
    t = cfg.text
    per_block = t.dim * 3 * t.dim + t.dim * t.dim + 2 * t.dim * t.ff_dim
    weights = t.n_layers * per_block + t.dim * t.vocab_size
    if not int8:
        return 2 * weights
    rows = t.n_layers * (3 * t.dim + t.dim + t.ff_dim + t.dim) + t.vocab_size
    return weights + 2 * rows

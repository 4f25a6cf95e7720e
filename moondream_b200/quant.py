"""Weight-only quantised decoder weights: the reference's int4 group-128 ``QuantizedLinear``
checkpoints (moondream/torch/layers.py:38-110, selected by ``TextConfig.group_size``) and int8 (BASELINE.json config 5).

The reference dequantises a ``QuantizedLinear`` once (``dequantize_tensor``, layers.py:38-44):

    W_r[:step] = high nibbles of packed, W_r[step:] = low nibbles          (packed: uint8 [out*in/256, 128])
    W = bf16( bf16(W_r - zero_point) * scale ).reshape(out, in)            (scale, zero_point: [out*in/128, 1])

i.e. group g = 128 consecutive input features of output row g // (in/128); byte (g, j) carries row r = g // (in/128)
in its high nibble and row r + out/2 in its low nibble.  (It then hands W to torchao's tinygemm, which re-quantises it;
torchao is not in this image, so — as SURVEY.md §8c prescribes — parity for quantised weights is *the bf16 path on W*.)

The CUDA stream (csrc/gemm_quant.cu) wants k-contiguous bytes, so at load time a checkpoint is re-laid out, values
untouched, into the *stream layout*:

    int4: wq uint8 [out, in/2], low nibble = feature 2j, high nibble = 2j+1;   scale, zero fp32 [out, in/128]
    int8: wq uint8 [out, in] (signed bytes);  scale fp32 [out, in/128] (one value per row, repeated), zero = 0

and the kernel rebuilds exactly W (two roundings, as above) tile by tile in shared memory.  The decoder blocks are
streamed fused — W1 = [qkv ; fc1] (rows), W2 = [proj | fc2] (input features) — and so are the packed tensors.
The LM head stays bf16 (the reference builds it as ``nn.Linear``, text.py:209); a quantised region head is
dequantised at load time (it is a few MB).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from .config import MoondreamConfig

GROUP = 128


# ---------------------------------------------------------------------------------------------------------------
# formats
# ---------------------------------------------------------------------------------------------------------------
def dequantize(values: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor) -> torch.Tensor:
    """values [out, in] (integers in any dtype), scale / zero [out, in/128] -> bf16 [out, in], with the reference's
    two roundings (layers.py:42-43: ``W_r.sub_(zero).mul_(scale)`` on a bf16 tensor, fp32 parameters)."""
    out_f, in_f = values.shape
    v = values.to(torch.float32).view(out_f, in_f // GROUP, GROUP)
    t = (v - zero.to(torch.float32).unsqueeze(-1)).to(torch.bfloat16)
    w = (t.to(torch.float32) * scale.to(torch.float32).unsqueeze(-1)).to(torch.bfloat16)
    return w.view(out_f, in_f)


def quantize_weight_int4(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Asymmetric 4-bit, group 128 along the input features: w [out, in] -> (nibbles uint8 [out, in] in 0..15,
    scale fp32 [out, in/128], zero fp32 [out, in/128]) with w ~ (nibble - zero) * scale."""
    if w.dim() != 2 or w.shape[1] % GROUP:
        raise ValueError("quantize_weight_int4 expects [out, in] with in % 128 == 0")
    out_f, in_f = w.shape
    g = w.detach().to(torch.float32).view(out_f, in_f // GROUP, GROUP)
    lo, hi = g.amin(dim=-1), g.amax(dim=-1)
    scale = ((hi - lo) / 15.0).to(torch.bfloat16).to(torch.float32)      # bf16-valued, like a bf16 checkpoint
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    zero = torch.round(-lo / scale).clamp_(0, 15)                          # integer zero points
    q = torch.round(g / scale.unsqueeze(-1) + zero.unsqueeze(-1)).clamp_(0, 15).to(torch.uint8)
    return q.view(out_f, in_f), scale, zero


def pack_reference_int4(nibbles: torch.Tensor) -> torch.Tensor:
    """[out, in] nibbles -> the reference checkpoint tensor ``packed`` uint8 [out*in/256, 128] (layers.py:58-63)."""
    out_f, in_f = nibbles.shape
    if out_f % 2 or in_f % GROUP:
        raise ValueError("the reference layout needs an even number of rows and in % 128 == 0")
    flat = nibbles.reshape(-1, GROUP)                                       # one row per group, row-major
    step = flat.shape[0] // 2
    return ((flat[:step] << 4) | flat[step:]).to(torch.uint8)


def unpack_reference_int4(packed: torch.Tensor, out_features: int, in_features: int) -> torch.Tensor:
    """The index part of ``dequantize_tensor`` (layers.py:39-41): packed [out*in/256, 128] -> nibbles [out, in]."""
    if tuple(packed.shape) != (out_features * in_features // (2 * GROUP), GROUP):
        raise ValueError(f"packed has shape {tuple(packed.shape)}, expected {(out_features * in_features // 256, 128)}")
    p = packed.to(torch.uint8)
    return torch.cat([(p & 0xF0) >> 4, p & 0x0F], 0).reshape(out_features, in_features)


def to_stream_int4(nibbles: torch.Tensor) -> torch.Tensor:
    """[out, in] nibbles -> stream bytes [out, in/2]: low nibble = even input feature, high nibble = odd."""
    return (nibbles[:, 0::2] | (nibbles[:, 1::2] << 4)).to(torch.uint8).contiguous()


def quantize_weight_int8(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Symmetric, per output feature: w [out, in] -> (q int8 [out, in], scale bf16 [out]);
    scale[n] = bf16(max_k |w[n, k]| / 127), q = clamp(rne(w / scale), -127, 127)."""
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
    """bf16(q * scale), one rounding per element (q is exact in bf16, so this equals `dequantize` with zero = 0)."""
    return (q.to(torch.float32) * scale.to(torch.float32).unsqueeze(1)).to(torch.bfloat16)


# ---------------------------------------------------------------------------------------------------------------
# the decoder's quantised state
# ---------------------------------------------------------------------------------------------------------------
BLOCK_LINEARS = ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")


@dataclass
class QuantLinear:
    bits: int
    values: torch.Tensor        # uint8 [out, in]: nibbles 0..15 (int4) or signed bytes viewed as uint8 (int8)
    scale: torch.Tensor         # fp32 [out, in/128]
    zero: torch.Tensor          # fp32 [out, in/128]

    def dequantized(self) -> torch.Tensor:
        v = self.values.view(torch.int8) if self.bits == 8 else self.values
        return dequantize(v, self.scale, self.zero)

    def stream_bytes(self) -> torch.Tensor:
        return to_stream_int4(self.values) if self.bits == 4 else self.values.contiguous()


@dataclass
class QuantizedText:
    """Per decoder block the four Linear layers in quantised form (host tensors)."""
    bits: int
    blocks: List[Dict[str, QuantLinear]]

    def fused(self, i: int):
        """Stream tensors of block i in the fused decode layout: (w1q, w1_scale, w1_zero, w2q, w2_scale, w2_zero),
        W1 = [qkv ; fc1] along rows, W2 = [proj | fc2] along input features."""
        b = self.blocks[i]
        w1q = torch.cat([b["attn.qkv"].stream_bytes(), b["mlp.fc1"].stream_bytes()], 0)
        w1s = torch.cat([b["attn.qkv"].scale, b["mlp.fc1"].scale], 0)
        w1z = torch.cat([b["attn.qkv"].zero, b["mlp.fc1"].zero], 0)
        w2q = torch.cat([b["attn.proj"].stream_bytes(), b["mlp.fc2"].stream_bytes()], 1)
        w2s = torch.cat([b["attn.proj"].scale, b["mlp.fc2"].scale], 1)
        w2z = torch.cat([b["attn.proj"].zero, b["mlp.fc2"].zero], 1)
        return tuple(t.contiguous() for t in (w1q, w1s.float(), w1z.float(), w2q, w2s.float(), w2z.float()))

    def nbytes(self) -> int:
        return sum(q.stream_bytes().numel() + 8 * q.scale.numel() for b in self.blocks for q in b.values())


def _block_shapes(cfg: MoondreamConfig) -> Dict[str, Tuple[int, int]]:
    t = cfg.text
    qkv = t.dim + 2 * t.n_kv_heads * t.head_dim
    return {"attn.qkv": (qkv, t.dim), "attn.proj": (t.dim, t.dim), "mlp.fc1": (t.ff_dim, t.dim),
            "mlp.fc2": (t.dim, t.ff_dim)}


def check_quantizable(cfg: MoondreamConfig) -> None:
    t = cfg.text
    if t.dim % GROUP or t.ff_dim % GROUP:
        raise ValueError("quantised decoder weights need text.dim and text.ff_dim to be multiples of the group size 128")


def quantize_decoder(cfg: MoondreamConfig, sd: Dict[str, torch.Tensor], bits: int):
    """Quantise the decoder blocks of a bf16 state dict -> (QuantizedText, dequantised state dict).  The second value
    shares every untouched tensor with `sd` and holds W (see the module docstring) for the block matrices: the weights
    the oracle, and the bf16 engine, must be run on to define parity for the quantised engine."""
    if bits not in (4, 8):
        raise ValueError("bits must be 4 or 8")
    check_quantizable(cfg)
    blocks: List[Dict[str, QuantLinear]] = []
    deq = dict(sd)
    for i in range(cfg.text.n_layers):
        blk: Dict[str, QuantLinear] = {}
        for name in BLOCK_LINEARS:
            key = f"text.blocks.{i}.{name}.weight"
            w = sd[key]
            if bits == 4:
                q, s, z = quantize_weight_int4(w)
                ql = QuantLinear(4, q, s, z)
            else:
                q8, s8 = quantize_weight_int8(w)
                groups = w.shape[1] // GROUP
                ql = QuantLinear(8, q8.view(torch.uint8), s8.float().unsqueeze(1).repeat(1, groups).contiguous(),
                                 torch.zeros(w.shape[0], groups))
            blk[name] = ql
            deq[key] = ql.dequantized()
        blocks.append(blk)
    return QuantizedText(bits, blocks), deq


def reference_checkpoint_entries(cfg: MoondreamConfig, qt: QuantizedText) -> Dict[str, torch.Tensor]:
    """The state-dict entries a reference int4 checkpoint holds for the decoder blocks
    (``…weight.packed / .scale / .zero_point``, layers.py:58-76) — used by tests and tools to write one."""
    if qt.bits != 4:
        raise ValueError("the reference checkpoint format is int4")
    out: Dict[str, torch.Tensor] = {}
    for i, blk in enumerate(qt.blocks):
        for name, ql in blk.items():
            p = f"text.blocks.{i}.{name}.weight."
            out[p + "packed"] = pack_reference_int4(ql.values)
            out[p + "scale"] = ql.scale.reshape(-1, 1).clone()
            out[p + "zero_point"] = ql.zero.reshape(-1, 1).clone()
    return out


def is_quantized_checkpoint(sd: Dict[str, torch.Tensor]) -> bool:
    return any(k.endswith(".weight.packed") for k in sd)


def from_reference_checkpoint(cfg: MoondreamConfig, sd: Dict[str, torch.Tensor]):
    """A state dict in the reference's int4 format -> (QuantizedText, state dict without the packed entries in which
    a quantised region head, if any, is replaced by its dequantised bf16 weights)."""
    check_quantizable(cfg)
    shapes = _block_shapes(cfg)
    blocks: List[Dict[str, QuantLinear]] = []
    rest = {k: v for k, v in sd.items() if ".weight." not in k}

    def take(prefix: str, out_f: int, in_f: int) -> QuantLinear:
        packed, scale, zero = sd[prefix + "packed"], sd[prefix + "scale"], sd[prefix + "zero_point"]
        nib = unpack_reference_int4(packed, out_f, in_f)
        groups = in_f // GROUP
        return QuantLinear(4, nib, scale.to(torch.float32).reshape(out_f, groups).contiguous(),
                           zero.to(torch.float32).reshape(out_f, groups).contiguous())

    for i in range(cfg.text.n_layers):
        blk = {}
        for name in BLOCK_LINEARS:
            out_f, in_f = shapes[name]
            blk[name] = take(f"text.blocks.{i}.{name}.weight.", out_f, in_f)
        blocks.append(blk)
    for k in list(sd):                                   # anything else that is packed (region head): dequantise now
        if k.endswith(".weight.packed") and not k.startswith("text.blocks."):
            base = k[: -len("packed")]
            bias = sd[base[: -len("weight.")] + "bias"]
            out_f = bias.numel()
            in_f = sd[k].numel() * 2 // out_f
            rest[base[:-1]] = take(base, out_f, in_f).dequantized()
    return QuantizedText(4, blocks), rest


# ---------------------------------------------------------------------------------------------------------------
# accounting
# ---------------------------------------------------------------------------------------------------------------
def decode_stream_keys(cfg: MoondreamConfig) -> Iterable[str]:
    """The matrices a decode step streams from HBM (2.63 GB per step for the 2B): the decoder blocks'
    four Linear layers and the LM head."""
    for i in range(cfg.text.n_layers):
        for name in BLOCK_LINEARS:
            yield f"text.blocks.{i}.{name}.weight"
    yield "text.lm_head.weight"


def quantize_decoder_int8(cfg: MoondreamConfig, sd: Dict[str, torch.Tensor]):
    """Round-1 helper kept for its tests: per-matrix (q, scale) incl. the LM head, and the dequantised state dict."""
    packed: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
    deq = dict(sd)
    for key in decode_stream_keys(cfg):
        q, scale = quantize_weight_int8(sd[key])
        packed[key] = (q, scale)
        deq[key] = dequantize_weight_int8(q, scale)
    return packed, deq


def stream_bytes(cfg: MoondreamConfig, int8: bool = False, bits: Optional[int] = None) -> int:
    """Weight bytes one decode step reads (the HBM-roofline numerator of SURVEY.md §8d).  `int8=True` is the round-1
    estimate (everything incl. the LM head at one byte); `bits` in (4, 8) is what the engine does: blocks quantised
    with fp32 scale + zero per group of 128, LM head bf16."""
    t = cfg.text
    qkv = t.dim + 2 * t.n_kv_heads * t.head_dim
    per_block = t.dim * qkv + t.dim * t.dim + 2 * t.dim * t.ff_dim
    head = t.dim * t.vocab_size
    if bits in (4, 8):
        return t.n_layers * (per_block * bits // 8 + per_block // GROUP * 8) + 2 * head
    weights = t.n_layers * per_block + head
    if not int8:
        return 2 * weights
    rows = t.n_layers * (qkv + t.dim + t.ff_dim + t.dim) + t.vocab_size
    return weights + 2 * rows

"""Seeded synthetic checkpoints and inputs (there is no network, hence no real weights).

The tensors use the canonical ``state_dict`` layout of the reference (``vision.*``, ``text.*``,
``region.*`` — see SURVEY.md §2.4 and moondream/torch/weights.py:123-131), so the same dict feeds
the reference / oracle and this engine.  Generation is tensor-by-tensor from one CPU generator in
a fixed key order, which makes it reproducible on any box with the same torch build
(tests/golden/synth_hashes.json pins it).

Recipe (chosen so activations stay O(1) through 27 + 24 pre-LN blocks and greedy decoding never
stops early):  W ~ N(0, g / sqrt(fan_in)) with g = 1 for expanding layers and 0.35 for the layers
that write into the residual stream; biases ~ N(0, 0.02); LayerNorm weight ~ 1 + N(0, 0.05);
``lm_head.bias[eos_id] = -3e4`` so exactly ``max_tokens`` steps run.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterable, List, Tuple

import numpy as np
import torch

from .config import MoondreamConfig


def state_dict_spec(cfg: MoondreamConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every persistent tensor, in canonical order.
    kind: 'w' expanding linear weight, 'wr' residual-writing linear weight, 'b' bias,
    'lnw' / 'lnb' LayerNorm, 'emb' embeddings, 'pos' positional, 'feat' fourier frequencies."""
    v, t, r = cfg.vision, cfg.text, cfg.region
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    add = spec.append
    add(("vision.pos_emb", (1, v.tokens_per_crop, v.enc_dim), "pos"))
    add(("vision.patch_emb.weight", (v.enc_dim, v.patch_dim), "w"))
    add(("vision.patch_emb.bias", (v.enc_dim,), "b"))
    for i in range(v.enc_n_layers):
        p = f"vision.blocks.{i}."
        add((p + "ln1.weight", (v.enc_dim,), "lnw"))
        add((p + "ln1.bias", (v.enc_dim,), "lnb"))
        add((p + "attn.qkv.weight", (3 * v.enc_dim, v.enc_dim), "w"))
        add((p + "attn.qkv.bias", (3 * v.enc_dim,), "b"))
        add((p + "attn.proj.weight", (v.enc_dim, v.enc_dim), "wr"))
        add((p + "attn.proj.bias", (v.enc_dim,), "b"))
        add((p + "ln2.weight", (v.enc_dim,), "lnw"))
        add((p + "ln2.bias", (v.enc_dim,), "lnb"))
        add((p + "mlp.fc1.weight", (v.enc_ff_dim, v.enc_dim), "w"))
        add((p + "mlp.fc1.bias", (v.enc_ff_dim,), "b"))
        add((p + "mlp.fc2.weight", (v.enc_dim, v.enc_ff_dim), "wr"))
        add((p + "mlp.fc2.bias", (v.enc_dim,), "b"))
    add(("vision.post_ln.weight", (v.enc_dim,), "lnw"))
    add(("vision.post_ln.bias", (v.enc_dim,), "lnb"))
    add(("vision.proj_mlp.fc1.weight", (v.proj_inner_dim, 2 * v.enc_dim), "w"))
    add(("vision.proj_mlp.fc1.bias", (v.proj_inner_dim,), "b"))
    add(("vision.proj_mlp.fc2.weight", (v.proj_out_dim, v.proj_inner_dim), "w"))
    add(("vision.proj_mlp.fc2.bias", (v.proj_out_dim,), "b"))
    qkv_dim = int(t.dim * (1 + 2 * t.n_kv_heads / t.n_heads))
    add(("text.wte", (t.vocab_size, t.dim), "emb"))
    for i in range(t.n_layers):
        p = f"text.blocks.{i}."
        add((p + "ln.weight", (t.dim,), "lnw"))
        add((p + "ln.bias", (t.dim,), "lnb"))
        add((p + "attn.qkv.weight", (qkv_dim, t.dim), "w"))
        add((p + "attn.qkv.bias", (qkv_dim,), "b"))
        add((p + "attn.proj.weight", (t.dim, t.dim), "wr"))
        add((p + "attn.proj.bias", (t.dim,), "b"))
        add((p + "mlp.fc1.weight", (t.ff_dim, t.dim), "w"))
        add((p + "mlp.fc1.bias", (t.ff_dim,), "b"))
        add((p + "mlp.fc2.weight", (t.dim, t.ff_dim), "wr"))
        add((p + "mlp.fc2.bias", (t.dim,), "b"))
    add(("text.post_ln.weight", (t.dim,), "lnw"))
    add(("text.post_ln.bias", (t.dim,), "lnb"))
    add(("text.lm_head.weight", (t.vocab_size, t.dim), "head"))
    add(("text.lm_head.bias", (t.vocab_size,), "b"))
    add(("region.coord_features", (1, r.coord_feat_dim // 2), "feat"))
    add(("region.size_features", (2, r.size_feat_dim // 2), "feat"))
    add(("region.coord_encoder.weight", (r.dim, r.coord_feat_dim), "w"))
    add(("region.coord_encoder.bias", (r.dim,), "b"))
    add(("region.coord_decoder.fc1.weight", (r.inner_dim, r.dim), "w"))
    add(("region.coord_decoder.fc1.bias", (r.inner_dim,), "b"))
    add(("region.coord_decoder.fc2.weight", (r.coord_out_dim, r.inner_dim), "w"))
    add(("region.coord_decoder.fc2.bias", (r.coord_out_dim,), "b"))
    add(("region.size_encoder.weight", (r.dim, r.size_feat_dim), "w"))
    add(("region.size_encoder.bias", (r.dim,), "b"))
    add(("region.size_decoder.fc1.weight", (r.inner_dim, r.dim), "w"))
    add(("region.size_decoder.fc1.bias", (r.inner_dim,), "b"))
    add(("region.size_decoder.fc2.weight", (r.size_out_dim, r.inner_dim), "w"))
    add(("region.size_decoder.fc2.bias", (r.size_out_dim,), "b"))
    return spec


def param_count(cfg: MoondreamConfig) -> int:
    return sum(int(np.prod(s)) for _, s, _ in state_dict_spec(cfg))


def synthetic_state_dict(cfg: MoondreamConfig, seed: int = 0, head_gain: float = 4.0,
                         head_peak: float = 0.0) -> Dict[str, torch.Tensor]:
    """bf16 CPU tensors in the canonical layout.  `head_peak` > 0 adds the wide-margin component of
    `peaked_lm_head` to the LM head (see there); 0 keeps the plain Gaussian head."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec(cfg):
        x = torch.empty(shape, dtype=torch.float32).normal_(0.0, 1.0, generator=gen)
        if kind in ("w", "wr", "head"):
            gain = {"w": 1.0, "wr": 0.35, "head": head_gain}[kind]
            x.mul_(gain / shape[-1] ** 0.5)
        elif kind == "b":
            x.mul_(0.02)
        elif kind == "lnw":
            x.mul_(0.05).add_(1.0)
        elif kind == "lnb":
            x.mul_(0.02)
        elif kind == "emb":
            x.mul_(0.5)
        elif kind == "pos":
            x.mul_(0.1)
        elif kind == "feat":
            x.mul_(3.0)
        sd[key] = x.to(torch.bfloat16)
    sd["text.lm_head.bias"][cfg.tokenizer.eos_id] = -3.0e4
    if head_peak:
        sd["text.lm_head.weight"] = peaked_lm_head(sd, head_peak, seed)
    return sd


def peaked_lm_head(sd: Dict[str, torch.Tensor], peak: float, seed: int = 0) -> torch.Tensor:
    """Wide-margin LM head ("tied lm_head"): the Gaussian head plus `peak` times a
    row-permuted, row-normalised copy of the token embedding, head[perm[t]] += peak * wte[t] / |wte[t]|.
    A plain Gaussian head gives top-1/top-2 margins of 1..10 bf16 ulps (about half of all greedy decisions are ties
    up to bf16 rounding, so strict token equality between two bf16 implementations cannot hold); the tied
    component makes the token that follows t under the fixed permutation stand out by tens of ulps.  Around
    peak = 2..3 the sequence still depends on the image and the context (the Gaussian part overrides the chain at
    some steps, and those steps are the remaining near-ties); from peak = 6 on every margin is > 40 ulps and the
    sequence is a function of the prompt's last token only (a strict end-to-end plumbing check)."""
    wte = sd["text.wte"].float()
    perm = torch.randperm(wte.shape[0], generator=torch.Generator(device="cpu").manual_seed(1_000_003 + seed))
    tied = torch.empty_like(wte)
    tied[perm] = wte / wte.norm(dim=1, keepdim=True)
    return (sd["text.lm_head.weight"].float() + peak * tied).to(torch.bfloat16)


def special_token_bias(sd: Dict[str, torch.Tensor], cfg: MoondreamConfig, answer: float, coord: float,
                       ground: float) -> torch.Tensor:
    """lm_head.bias with the control tokens of the grounded chain of thought lifted (answer_id, coord_id,
    start_ground_points_id / end_ground_id): a random network never emits them on its own, and the reasoning tests need
    sequences that interleave coordinates, close grounding spans and terminate (moondream.py:363-432)."""
    tk = cfg.tokenizer
    b = sd["text.lm_head.bias"].clone().float()
    b[tk.answer_id] += answer
    b[tk.coord_id] += coord
    b[tk.start_ground_points_id] += ground
    b[tk.end_ground_id] += ground
    return b.to(torch.bfloat16)


def synthetic_lora(cfg: MoondreamConfig, rank: int = 8, seed: int = 0, gain: float = 0.5) -> Dict[str, torch.Tensor]:
    """A seeded LoRA variant in the flat key layout the reference's `variant_state_dict` produces after its renames
    (lora.py:64-79): text.blocks.{i}.attn.{qkv,proj}.{A,B} and text.blocks.{i}.mlp.{fc1,fc2}.{A,B}, A [rank, in],
    B [out, rank], bf16.  `gain` scales B so that the adapters move the logits visibly."""
    gen = torch.Generator(device="cpu").manual_seed(7_000_003 + seed)
    t = cfg.text
    qkv = int(t.dim * (1 + 2 * t.n_kv_heads / t.n_heads))
    shapes = {"attn.qkv": (t.dim, qkv), "attn.proj": (t.dim, t.dim), "mlp.fc1": (t.dim, t.ff_dim), "mlp.fc2": (t.ff_dim, t.dim)}
    out: Dict[str, torch.Tensor] = {}
    for i in range(t.n_layers):
        for name, (fin, fout) in shapes.items():
            a = torch.empty((rank, fin), dtype=torch.float32).normal_(0.0, 1.0, generator=gen) / fin ** 0.5
            b = torch.empty((fout, rank), dtype=torch.float32).normal_(0.0, 1.0, generator=gen) * (gain / rank ** 0.5)
            out[f"text.blocks.{i}.{name}.A"] = a.to(torch.bfloat16)
            out[f"text.blocks.{i}.{name}.B"] = b.to(torch.bfloat16)
    return out


def nest_lora(flat: Dict[str, torch.Tensor]) -> dict:
    """lora.py:43-52 `nest`: flat dotted keys -> nested dicts."""
    tree: dict = {}
    for k, v in flat.items():
        parts = k.split(".")
        d = tree
        for p_ in parts[:-1]:
            d = d.setdefault(p_, {})
        d[parts[-1]] = v
    return tree


def tensor_hash(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()[:16]


def state_dict_fingerprint(sd: Dict[str, torch.Tensor], keys: Iterable[str]) -> Dict[str, str]:
    return {k: tensor_hash(sd[k]) for k in keys}


FINGERPRINT_KEYS = ("vision.patch_emb.weight", "vision.blocks.26.mlp.fc2.weight", "text.wte",
                    "text.blocks.0.attn.qkv.weight", "text.lm_head.bias", "region.size_features")


def synthetic_image(index: int, height: int = 378, width: int = 378, seed: int = 0) -> np.ndarray:
    """uint8 HxWx3, `np.random.default_rng(seed + index)`."""
    return np.random.default_rng(seed + index).integers(0, 256, (height, width, 3), dtype=np.uint8)


def synthetic_prompt(index: int, length: int = 32, vocab_size: int = 51200, seed: int = 0) -> List[int]:
    """`length` token ids uniform in [10, vocab) — avoids the special ids 0-9 (config.py:46-53)."""
    rng = np.random.default_rng(1_000_003 * (seed + 1) + index)
    return [int(v) for v in rng.integers(10, vocab_size, length)]

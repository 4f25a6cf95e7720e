"""Model configuration — same names, fields and defaults as the reference's
``moondream/torch/config.py:5-94`` (TextConfig / VisionConfig / RegionConfig / TokenizerConfig /
MoondreamConfig with ``from_dict`` / ``to_dict``) so existing JSON configs load unchanged.

Additions over the reference: named presets (``moondream_2b``, ``moondream_0_5b``, ``tiny``),
derived properties the kernels need, and ``validate()`` which states the shapes the sm_100a
kernels support instead of failing deep inside a launch.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Dict, List, Optional


@dataclass(frozen=True)
class TextConfig:
    dim: int = 2048
    ff_dim: int = 8192
    n_layers: int = 24
    vocab_size: int = 51200
    max_context: int = 2048
    n_heads: int = 32
    n_kv_heads: int = 32
    prefix_attn: int = 730
    group_size: Optional[int] = None

    @property
    def head_dim(self) -> int:
        return self.dim // self.n_heads


@dataclass(frozen=True)
class VisionConfig:
    enc_dim: int = 1152
    enc_patch_size: int = 14
    enc_n_layers: int = 27
    enc_ff_dim: int = 4304
    enc_n_heads: int = 16
    proj_out_dim: int = 2048
    crop_size: int = 378
    in_channels: int = 3
    max_crops: int = 12
    overlap_margin: int = 4
    proj_inner_dim: int = 8192

    @property
    def grid(self) -> int:
        """Patches per crop side (27).  The reference reuses ``enc_n_layers`` for this
        (moondream.py:216-217, vision.py:85) — a coincidence this code does not rely on."""
        return self.crop_size // self.enc_patch_size

    @property
    def tokens_per_crop(self) -> int:
        return self.grid * self.grid

    @property
    def patch_dim(self) -> int:
        return self.enc_patch_size * self.enc_patch_size * self.in_channels

    @property
    def head_dim(self) -> int:
        return self.enc_dim // self.enc_n_heads


@dataclass(frozen=True)
class RegionConfig:
    dim: int = 2048
    coord_feat_dim: int = 256
    coord_out_dim: int = 1024
    size_feat_dim: int = 512
    size_out_dim: int = 2048
    inner_dim: int = 8192
    group_size: Optional[int] = None


def _default_templates() -> Dict[str, Optional[Dict[str, List[int]]]]:
    return {
        "caption": {
            "short": [1, 32708, 2, 12492, 3],
            "normal": [1, 32708, 2, 6382, 3],
            "long": [1, 32708, 2, 4059, 3],
        },
        "query": {"prefix": [1, 15381, 2], "suffix": [3]},
        "detect": {"prefix": [1, 7235, 476, 2], "suffix": [3]},
        "point": {"prefix": [1, 2581, 2], "suffix": [3]},
    }


@dataclass(frozen=True)
class TokenizerConfig:
    bos_id: int = 0
    eos_id: int = 0
    answer_id: int = 3
    thinking_id: int = 4
    coord_id: int = 5
    size_id: int = 6
    start_ground_points_id: int = 7
    end_ground_id: int = 9
    templates: Dict[str, Optional[Dict[str, List[int]]]] = field(default_factory=_default_templates)


@dataclass(frozen=True)
class MoondreamConfig:
    text: TextConfig = TextConfig()
    vision: VisionConfig = VisionConfig()
    region: RegionConfig = RegionConfig()
    tokenizer: TokenizerConfig = TokenizerConfig()

    @classmethod
    def from_dict(cls, config_dict: dict) -> "MoondreamConfig":
        return cls(
            text=TextConfig(**config_dict.get("text", {})),
            vision=VisionConfig(**config_dict.get("vision", {})),
            region=RegionConfig(**config_dict.get("region", {})),
            tokenizer=TokenizerConfig(**config_dict.get("tokenizer", {})),
        )

    def to_dict(self) -> dict:
        return {
            "text": dict(self.text.__dict__),
            "vision": dict(self.vision.__dict__),
            "region": dict(self.region.__dict__),
            "tokenizer": dict(self.tokenizer.__dict__),
        }

    # ------------------------------------------------------------------ additions
    def validate(self) -> None:
        """Raise ValueError for shapes the sm_100a kernels do not implement."""
        t, v, r = self.text, self.vision, self.region
        if t.dim % t.n_heads or t.head_dim != 64:
            raise ValueError("text head_dim must be 64 (partial RoPE over 32 dims, rope.py:20-48)")
        if t.n_kv_heads <= 0 or t.n_heads % t.n_kv_heads:
            raise ValueError("text.n_heads must be a multiple of text.n_kv_heads (grouped-query attention, "
                             "text.py:49); the 0.5B JSON omits n_kv_heads and the reference default of 32 does not "
                             "divide its 16 heads: set it explicitly")
        if t.ff_dim % (2 * t.n_heads):
            raise ValueError("text.ff_dim must be a multiple of 2 * n_heads (the fused decode attention finishes an "
                             "equal, even slice of the fc1 features per head)")
        if v.enc_dim % v.enc_n_heads or v.head_dim != 72:
            raise ValueError("vision head_dim must be 72")
        if v.crop_size % v.enc_patch_size:
            raise ValueError("crop_size must be a multiple of enc_patch_size")
        if t.prefix_attn != v.tokens_per_crop + 1:
            raise ValueError("prefix_attn must equal 1 + tokens per crop (moondream.py:143-145)")
        if v.proj_out_dim != t.dim or r.dim != t.dim:
            raise ValueError("projection / region width must equal the text width")
        for gs in (t.group_size, r.group_size):
            if gs is not None and gs != 128:
                raise ValueError("QuantizedLinear checkpoints use group_size 128 (layers.py:54)")
        if t.group_size is not None and (t.dim % 128 or t.ff_dim % 128):
            raise ValueError("int4 decoder weights need text.dim and text.ff_dim to be multiples of 128")
        for name, val in (("text.dim", t.dim), ("text.ff_dim", t.ff_dim), ("vision.enc_dim", v.enc_dim),
                          ("vision.proj_inner_dim", v.proj_inner_dim), ("region.inner_dim", r.inner_dim),
                          ("text.vocab_size", t.vocab_size)):
            if val % 8:
                raise ValueError(f"{name} must be a multiple of 8")
        if r.size_out_dim % 2:
            raise ValueError("region.size_out_dim must be even (decode_size views it as (2, -1))")


def moondream_2b() -> MoondreamConfig:
    """Moondream-2B (the Python defaults of the reference, config.py:5-41)."""
    return MoondreamConfig()


def moondream_0_5b() -> MoondreamConfig:
    """Moondream-0.5B (moondream/config/config_md05.json) with ``n_kv_heads`` set explicitly:
    the JSON omits it and the reference default of 32 does not divide 16 heads."""
    return MoondreamConfig(
        text=TextConfig(dim=1024, ff_dim=4096, n_layers=24, n_heads=16, n_kv_heads=16),
        vision=VisionConfig(enc_dim=720, enc_ff_dim=2690, enc_n_heads=10, proj_out_dim=1024),
        region=RegionConfig(dim=1024),
    )


def tiny() -> MoondreamConfig:
    """A few-million-parameter model with the same topology (27 ViT blocks, 729 tokens per crop,
    730-token prefix) for fast CPU oracle runs and GPU parity tests."""
    templates = {
        "caption": {"short": [1, 708, 2, 492, 3], "normal": [1, 708, 2, 382, 3], "long": [1, 708, 2, 59, 3]},
        "query": {"prefix": [1, 381, 2], "suffix": [3]},
        "detect": {"prefix": [1, 235, 476, 2], "suffix": [3]},
        "point": {"prefix": [1, 581, 2], "suffix": [3]},
    }
    return MoondreamConfig(
        text=TextConfig(dim=128, ff_dim=512, n_layers=4, vocab_size=2048, n_heads=2, n_kv_heads=2),
        vision=VisionConfig(enc_dim=144, enc_ff_dim=304, enc_n_heads=2, proj_out_dim=128,
                            proj_inner_dim=256),
        region=RegionConfig(dim=128, inner_dim=512),
        tokenizer=TokenizerConfig(templates=templates),
    )


def tiny_gqa() -> MoondreamConfig:
    """`tiny` with grouped-query attention in the decoder: 4 query heads share 2 KV heads (text.py:36-38,49)."""
    base = tiny()
    return MoondreamConfig(
        text=TextConfig(dim=256, ff_dim=512, n_layers=4, vocab_size=2048, n_heads=4, n_kv_heads=2),
        vision=dataclasses.replace(base.vision, proj_out_dim=256),
        region=RegionConfig(dim=256, inner_dim=512),
        tokenizer=base.tokenizer,
    )


PRESETS = {"moondream-2b": moondream_2b, "moondream-0.5b": moondream_0_5b, "tiny": tiny, "tiny-gqa": tiny_gqa}


def preset(name: str) -> MoondreamConfig:
    try:
        return PRESETS[name]()
    except KeyError:
        raise ValueError(f"unknown preset {name!r}; choose from {sorted(PRESETS)}") from None


def replace(cfg, **kw):
    return dataclasses.replace(cfg, **kw)

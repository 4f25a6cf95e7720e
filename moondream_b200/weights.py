"""Checkpoint loading — accepts what the reference's ``moondream/torch/weights.py`` accepts:

  * safetensors or ``.pt`` files (``load_weights_into_model`` :156-171);
  * canonical keys (``vision.blocks.0.attn.proj.bias`` …), optionally ``model.``-prefixed (:123-131);
  * legacy HF keys (``vision_encoder.encoder.model.visual.*``, ``text_model.transformer.h.N.mixer.Wqkv``,
    ``region_model.*``) with ``._orig_mod`` stripped and the two ``*_features`` tensors transposed
    (:36-117).

Everything is normalised to the canonical flat dict of bf16 tensors that ``MoondreamModel``
uploads; the one-time device re-layout (K padding for TMA alignment) happens in the engine.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable

import torch

from .config import MoondreamConfig
from .synth import state_dict_spec

_LEGACY_FIXED = {
    "vision_encoder.encoder.model.visual.patch_embed.linear.weight": "vision.patch_emb.weight",
    "vision_encoder.encoder.model.visual.patch_embed.linear.bias": "vision.patch_emb.bias",
    "vision_encoder.encoder.model.visual.pos_embed": "vision.pos_emb",
    "vision_encoder.encoder.model.visual.norm.weight": "vision.post_ln.weight",
    "vision_encoder.encoder.model.visual.norm.bias": "vision.post_ln.bias",
    "vision_encoder.projection.mlp.fc1.weight": "vision.proj_mlp.fc1.weight",
    "vision_encoder.projection.mlp.fc1.bias": "vision.proj_mlp.fc1.bias",
    "vision_encoder.projection.mlp.fc2.weight": "vision.proj_mlp.fc2.weight",
    "vision_encoder.projection.mlp.fc2.bias": "vision.proj_mlp.fc2.bias",
    "text_model.transformer.embd.wte.weight": "text.wte",
    "text_model.lm_head.ln.weight": "text.post_ln.weight",
    "text_model.lm_head.ln.bias": "text.post_ln.bias",
    "text_model.lm_head.linear.weight": "text.lm_head.weight",
    "text_model.lm_head.linear.bias": "text.lm_head.bias",
    "region_model.coordinate_encoder.weight": "region.coord_encoder.weight",
    "region_model.coordinate_encoder.bias": "region.coord_encoder.bias",
    "region_model.coordinate_decoder.fc1.weight": "region.coord_decoder.fc1.weight",
    "region_model.coordinate_decoder.fc1.bias": "region.coord_decoder.fc1.bias",
    "region_model.coordinate_decoder.fc2.weight": "region.coord_decoder.fc2.weight",
    "region_model.coordinate_decoder.fc2.bias": "region.coord_decoder.fc2.bias",
    "region_model.size_encoder.weight": "region.size_encoder.weight",
    "region_model.size_encoder.bias": "region.size_encoder.bias",
    "region_model.size_decoder.fc1.weight": "region.size_decoder.fc1.weight",
    "region_model.size_decoder.fc1.bias": "region.size_decoder.fc1.bias",
    "region_model.size_decoder.fc2.weight": "region.size_decoder.fc2.weight",
    "region_model.size_decoder.fc2.bias": "region.size_decoder.fc2.bias",
}
_LEGACY_TRANSPOSED = {
    "region_model.coordinate_features.weight": "region.coord_features",
    "region_model.size_features.weight": "region.size_features",
}
_VIS_BLOCK = {"norm1": "ln1", "norm2": "ln2", "attn.qkv": "attn.qkv", "attn.proj": "attn.proj",
              "mlp.fc1": "mlp.fc1", "mlp.fc2": "mlp.fc2"}
_TXT_BLOCK = {"ln": "ln", "mixer.Wqkv": "attn.qkv", "mixer.out_proj": "attn.proj",
              "mlp.fc1": "mlp.fc1", "mlp.fc2": "mlp.fc2"}


def legacy_key_map(cfg: MoondreamConfig) -> Dict[str, str]:
    """legacy HF key -> canonical key (tensors in _LEGACY_TRANSPOSED also need ``.T``)."""
    m = dict(_LEGACY_FIXED)
    for i in range(cfg.vision.enc_n_layers):
        for old, new in _VIS_BLOCK.items():
            for leaf in ("weight", "bias"):
                m[f"vision_encoder.encoder.model.visual.blocks.{i}.{old}.{leaf}"] = \
                    f"vision.blocks.{i}.{new}.{leaf}"
    for i in range(cfg.text.n_layers):
        for old, new in _TXT_BLOCK.items():
            for leaf in ("weight", "bias"):
                m[f"text_model.transformer.h.{i}.{old}.{leaf}"] = f"text.blocks.{i}.{new}.{leaf}"
    return m


def _is_canonical(keys: Iterable[str]) -> bool:
    keys = set(keys)
    return "vision.blocks.0.attn.proj.bias" in keys or "model.vision.blocks.0.attn.proj.bias" in keys


def normalize_state_dict(cfg: MoondreamConfig, keys: Iterable[str],
                         get: Callable[[str], torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Map either accepted key layout to canonical bf16 tensors and check shapes."""
    keys = list(keys)
    out: Dict[str, torch.Tensor] = {}
    if _is_canonical(keys):
        for k in keys:
            out[k.replace("model.", "")] = get(k)
    else:
        stripped = {k.replace("._orig_mod", ""): k for k in keys}
        lm = legacy_key_map(cfg)
        for old, new in lm.items():
            if old not in stripped:
                raise KeyError(f"checkpoint is missing {old}")
            out[new] = get(stripped[old])
        for old, new in _LEGACY_TRANSPOSED.items():
            if old not in stripped:
                raise KeyError(f"checkpoint is missing {old}")
            out[new] = get(stripped[old]).T
    result: Dict[str, torch.Tensor] = {}
    for key, shape, _ in state_dict_spec(cfg):
        if key + ".packed" in out:
            # a QuantizedLinear of the reference's int4 checkpoints (layers.py:58-76): packed uint8 + fp32 scale /
            # zero_point per group of 128 travel as they are; moondream_b200.quant re-lays them out for the engine
            result[key + ".packed"] = out[key + ".packed"].to(torch.uint8).contiguous()
            result[key + ".scale"] = out[key + ".scale"].to(torch.float32).contiguous()
            result[key + ".zero_point"] = out[key + ".zero_point"].to(torch.float32).contiguous()
            continue
        if key not in out:
            raise KeyError(f"checkpoint is missing {key}")
        t = out[key].to(torch.bfloat16).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{key}: expected shape {tuple(shape)}, checkpoint has {tuple(t.shape)}")
        result[key] = t
    return result


_ST_DTYPES = {"BF16": torch.bfloat16, "F16": torch.float16, "F32": torch.float32, "F64": torch.float64,
              "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8,
              "BOOL": torch.bool}


class NativeSafetensors:
    """The library's own safetensors reader (csrc/loader.cu: mmap + header parse): tensors are copied from the file
    mapping straight into the torch tensor `get` allocates — on `device` (e.g. "cuda": one H2D per tensor, no host
    copy) or on the CPU."""

    def __init__(self, path: str, device="cpu"):
        import ctypes

        from . import _native as N

        self._N, self._lib = N, N.lib()
        self.device = torch.device(device)
        h = ctypes.c_void_p()
        N.check(self._lib.md_safetensors_open(path.encode(), ctypes.byref(h)), "md_safetensors_open")
        self._h = h
        self._index: Dict[str, tuple] = {}
        for i in range(self._lib.md_safetensors_count(h)):
            name, dtype = ctypes.c_char_p(), ctypes.c_char_p()
            ndim, nbytes = ctypes.c_int(), ctypes.c_longlong()
            shape = (ctypes.c_longlong * 8)()
            N.check(self._lib.md_safetensors_info(h, i, ctypes.byref(name), ctypes.byref(dtype), ctypes.byref(ndim),
                                                  shape, ctypes.byref(nbytes)), "md_safetensors_info")
            self._index[name.value.decode()] = (i, dtype.value.decode(), tuple(shape[: ndim.value]), nbytes.value)

    def keys(self):
        return list(self._index)

    def get_tensor(self, key: str) -> torch.Tensor:
        i, dtype, shape, nbytes = self._index[key]
        if dtype not in _ST_DTYPES:
            raise ValueError(f"{key}: unsupported safetensors dtype {dtype}")
        t = torch.empty(shape, dtype=_ST_DTYPES[dtype], device=self.device)
        on_dev = self.device.type == "cuda"
        stream = self._N.current_stream() if on_dev else None
        self._N.check(self._lib.md_safetensors_read(self._h, i, self._N.ptr(t) if t.numel() else None, nbytes, int(on_dev),
                                                    stream), "md_safetensors_read") if t.numel() else None
        return t

    def close(self):
        if self._h is not None:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)        # the async copies read the mapping
            self._lib.md_safetensors_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def load_state_dict_from_file(weights_file: str, cfg: MoondreamConfig, device="cpu") -> Dict[str, torch.Tensor]:
    """safetensors: through the native reader (tensors land on `device` directly); .pt: torch.load."""
    if weights_file.endswith(".safetensors"):
        with NativeSafetensors(weights_file, device) as st:
            return normalize_state_dict(cfg, st.keys(), st.get_tensor)
    tensors = torch.load(weights_file, map_location="cpu", weights_only=True)
    return normalize_state_dict(cfg, list(tensors.keys()), lambda k: tensors[k])


def load_weights_into_model(weights_file: str, model) -> None:
    """Drop-in for ``moondream.torch.weights.load_weights_into_model`` (weights.py:156)."""
    device = model.device if torch.cuda.is_available() else "cpu"
    model.load_state_dict(load_state_dict_from_file(weights_file, model.config, device))

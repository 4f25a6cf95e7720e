"""Batched inference engine over the C-ABI (host side, Python; device side, hand-written CUDA).

Responsibilities kept in Python: weight preparation/upload, KV page bookkeeping, batch assembly,
the decode loop driver (one CUDA graph per batch size, replayed per token, no per-token host sync).
All arithmetic happens inside libmoondream_b200.so.  There is no CPU / eager fallback.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from . import quant as Q
from .config import MoondreamConfig
from .image_crops import crop_tiling, overlap_crop_image
from .synth import state_dict_spec

PAGE = 64


def _on_device(fn):
    """Run a public Engine method with the engine's GPU as the current device, so allocations, the current
    stream and every native launch land on it even when the caller's current device is another GPU."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)

    return wrapped


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _lru_get(cache: dict, key):
    """dict as an LRU (insertion order = age): a hit moves the entry to the young end"""
    hit = cache.pop(key, None)
    if hit is not None:
        cache[key] = hit
    return hit


def _lru_put(cache: dict, key, value, cap: int):
    cache.pop(key, None)
    cache[key] = value
    while len(cache) > cap:
        cache.pop(next(iter(cache)))


def pixel_lut() -> torch.Tensor:
    """The reference's pixel normalisation (vision.py:33-40) applied to every uint8 value with the
    same torch CPU ops, so the device path is bit-exact: uint8 -> bf16, /255, -0.5, /0.5."""
    v = torch.arange(256, dtype=torch.uint8)
    return v.to(dtype=torch.bfloat16).div_(255.0).sub_(0.5).div_(0.5)


def rope_table(head_dim: int, max_context: int, theta: float = 10000.0) -> torch.Tensor:
    """precompute_freqs_cis (rope.py:6-17) as text.py:215-219 calls it -> f32 [ctx, hd/4, 2]."""
    dim = head_dim // 2
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.arange(max_context, dtype=torch.float32).unsqueeze(1) * freqs.unsqueeze(0)
    unit = torch.exp(1j * ang)
    return torch.stack([unit.real, unit.imag], dim=-1).contiguous()


_BLOCK_WEIGHT_SUFFIXES = ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")


def _is_block_weight(key: str) -> bool:
    return key.startswith("text.blocks.") and key.endswith(_BLOCK_WEIGHT_SUFFIXES)


def prepare_weights(cfg: MoondreamConfig, sd: Dict[str, torch.Tensor],
                    quantized_blocks: bool = False) -> Tuple[List[Optional[torch.Tensor]], int, int]:
    """One-time re-layout so every GEMM operand is TMA-addressable (16-byte row pitch):
    patch_emb K 588 -> 592 and the ViT MLP width to a multiple of 8 (0.5B: 2690 -> 2696) with zero
    padding, which leaves the arithmetic unchanged (gelu(0) = 0 meets zero fc2 columns)."""
    v = cfg.vision
    patch_k = _round_up(v.patch_dim, 8)
    vis_ff = _round_up(v.enc_ff_dim, 8)
    out: List[Optional[torch.Tensor]] = []
    for key, shape, _ in state_dict_spec(cfg):
        if quantized_blocks and _is_block_weight(key):
            out.append(None)                 # lives in packed form (quant.QuantizedText); bf16 scratch on the device
            continue
        if key not in sd:
            raise KeyError(f"state dict is missing {key}")
        t = sd[key].to(torch.bfloat16)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{key}: expected {tuple(shape)}, got {tuple(t.shape)}")
        if key == "vision.patch_emb.weight" and patch_k != v.patch_dim:
            t = torch.nn.functional.pad(t, (0, patch_k - v.patch_dim))
        elif vis_ff != v.enc_ff_dim and key.startswith("vision.blocks."):
            if key.endswith("mlp.fc1.weight"):
                t = torch.nn.functional.pad(t, (0, 0, 0, vis_ff - v.enc_ff_dim))
            elif key.endswith("mlp.fc1.bias"):
                t = torch.nn.functional.pad(t, (0, vis_ff - v.enc_ff_dim))
            elif key.endswith("mlp.fc2.weight"):
                t = torch.nn.functional.pad(t, (0, vis_ff - v.enc_ff_dim))
        out.append(t.contiguous())
    return out, patch_k, vis_ff


def upload_weights(cfg: MoondreamConfig, prepared: List[Optional[torch.Tensor]], device,
                   quantized_blocks: bool = False) -> Tuple[List[torch.Tensor], list]:
    """Upload in canonical order.  The decoder blocks use the fused decode layout the C runtime checks
    (md_dims.txt_fused): W1 = [qkv.weight ; fc1.weight], b1 = [qkv.bias ; fc1.bias] and
    W2 = [proj.weight | fc2.weight]; the canonical entries become views of those buffers, so prefill
    (separate GEMMs) and decode (one weight stream per pair) share the same memory."""
    keys = [k for k, _, _ in state_dict_spec(cfg)]
    idx = {k: i for i, k in enumerate(keys)}
    dev: List[Optional[torch.Tensor]] = [None] * len(keys)
    owners = []
    D = cfg.text.dim
    FF = cfg.text.ff_dim
    Q = D + 2 * cfg.text.n_kv_heads * cfg.text.head_dim          # rows of qkv.weight (text.py:36-38)
    scratch = None
    if quantized_blocks:
        # packed decoder blocks (quant.py): every block's bf16 pointers alias ONE scratch pair that prefill rebuilds
        # block by block (md_model_set_quantized_block)
        scratch = (torch.zeros((Q + FF, D), dtype=torch.bfloat16, device=device),
                   torch.zeros((D, D + FF), dtype=torch.bfloat16, device=device))
        owners += list(scratch)
    for i in range(cfg.text.n_layers):
        p = f"text.blocks.{i}."
        b1 = torch.cat([prepared[idx[p + "attn.qkv.bias"]], prepared[idx[p + "mlp.fc1.bias"]]], 0).to(device)
        if scratch is not None:
            w1, w2 = scratch
            owners += [b1]
        else:
            w1 = torch.cat([prepared[idx[p + "attn.qkv.weight"]], prepared[idx[p + "mlp.fc1.weight"]]], 0).to(device)
            w2 = torch.cat([prepared[idx[p + "attn.proj.weight"]], prepared[idx[p + "mlp.fc2.weight"]]], 1).to(device)
            owners += [w1, b1, w2]
        dev[idx[p + "attn.qkv.weight"]] = w1[:Q]
        dev[idx[p + "mlp.fc1.weight"]] = w1[Q:]
        dev[idx[p + "attn.qkv.bias"]] = b1[:Q]
        dev[idx[p + "mlp.fc1.bias"]] = b1[Q:]
        dev[idx[p + "attn.proj.weight"]] = w2[:, :D]
        dev[idx[p + "mlp.fc2.weight"]] = w2[:, D:]
    for i, t in enumerate(prepared):
        if dev[i] is None:
            dev[i] = t.to(device)
    return dev, owners  # type: ignore[return-value]


class PagePool:
    """KV pages: bf16 [layers, n_pages, 2, kv_heads, 64, 64]; a free list hands out page ids."""

    def __init__(self, cfg: MoondreamConfig, n_pages: int, device):
        t = cfg.text
        self.n_pages = n_pages
        # zeros, not empty: masked / not-yet-written slots still flow through P.V as 0 * v and must be finite
        self.pool = torch.zeros((t.n_layers, n_pages, 2, t.n_kv_heads, PAGE, 64), dtype=torch.bfloat16,
                                device=device)
        self._free = list(range(n_pages - 1, -1, -1))

    def alloc(self, n: int) -> List[int]:
        if n > len(self._free):
            raise N.NativeError(f"KV pool exhausted: need {n} pages, {len(self._free)} free "
                                f"(raise kv_pages when constructing the model)")
        return [self._free.pop() for _ in range(n)]

    def release(self, pages: Sequence[int]):
        self._free.extend(pages)

    @property
    def free_pages(self) -> int:
        return len(self._free)


@dataclass
class PrefixKV:
    """Device-resident KV prefix of one encoded image: `pos` tokens spread over `pages`."""
    pos: int
    pages: List[int]
    pool: PagePool = field(repr=False, default=None)
    _released: bool = field(default=False, repr=False)

    def release(self):
        if not self._released and self.pool is not None:
            self.pool.release(self.pages)
            self._released = True

    def __del__(self):  # pages go back to the pool when the handle dies
        try:
            self.release()
        except Exception:
            pass


@dataclass(frozen=True)
class DecodeMode:
    """What one captured decode step does; also the key of the CUDA-graph cache."""
    forced: bool            # teacher forcing: feed forced[:, step + 1] instead of the prediction
    temperature: float      # 0 = greedy argmax; > 0 = on-device top-p sampling
    top_p: float
    reasoning: bool         # coord_id tokens are fed as region-encoded coordinates (moondream.py:381-391)
    mask_id: int            # ids excluded by the LM head at decode steps (-1 = none)
    mask_id2: int
    eos_id: int             # sets finished[b]

    @property
    def sampled(self) -> bool:
        return self.temperature > 0.0


@dataclass
class GenerationResult:
    tokens: torch.Tensor      # int32 [B, max_tokens + 1] on host: greedy prediction at every step
    margins: Optional[torch.Tensor]
    steps: int


class LoraVariant:
    """A LoRA variant on the device (the reference's `variant_state_dict` tree, lora.py:55-79): per decoder block the
    (A, B) pairs of attn.qkv / attn.proj / mlp.fc1 / mlp.fc2, bf16, plus the host pointer table the C-ABI takes."""

    ORDER = (("attn", "qkv"), ("attn", "proj"), ("mlp", "fc1"), ("mlp", "fc2"))

    def __init__(self, cfg: MoondreamConfig, tree: dict, device):
        t = cfg.text
        blocks = tree["text"]["blocks"] if "text" in tree else tree["blocks"]
        self.tensors: List[torch.Tensor] = []
        ranks = set()
        qkv_rows = t.dim + 2 * t.n_kv_heads * t.head_dim
        want = {("attn", "qkv"): (t.dim, qkv_rows), ("attn", "proj"): (t.dim, t.dim),
                ("mlp", "fc1"): (t.dim, t.ff_dim), ("mlp", "fc2"): (t.ff_dim, t.dim)}
        for i in range(t.n_layers):
            blk = blocks[str(i)]
            for group, name in self.ORDER:
                ab = blk[group][name]
                a = ab["A"].to(device=device, dtype=torch.bfloat16).contiguous()
                b = ab["B"].to(device=device, dtype=torch.bfloat16).contiguous()
                fin, fout = want[(group, name)]
                if a.dim() != 2 or b.dim() != 2 or a.shape[1] != fin or b.shape[0] != fout or b.shape[1] != a.shape[0]:
                    raise ValueError(f"variant block {i} {group}.{name}: A {tuple(a.shape)} / B {tuple(b.shape)} do not fit "
                                     f"a Linear {fin} -> {fout}")
                ranks.add(int(a.shape[0]))
                self.tensors += [a, b]
        if len(ranks) != 1 or next(iter(ranks)) % 8:
            raise ValueError(f"the adapters must share one rank that is a multiple of 8, got {sorted(ranks)}")
        self.rank = next(iter(ranks))
        self.table = (ctypes.c_void_p * len(self.tensors))(*[x.data_ptr() for x in self.tensors])


class Engine:
    _MAX_DECODE_STATES = 4
    _MAX_DECODE_GRAPHS = 8      # captured decode steps per batch size: one per DecodeMode (temperature and top_p are
                                # kernel arguments baked into the capture, so callers that vary them would otherwise
                                # grow the cache, and the graphs' private memory, without bound)

    def __init__(self, cfg: MoondreamConfig, state_dict: Dict[str, torch.Tensor], device="cuda",
                 kv_pages: Optional[int] = None, max_batch: int = 32, quantize: Optional[str] = None):
        """quantize: None (bf16 decoder weights; a state dict in the reference's int4 QuantizedLinear format is
        detected by its ``.weight.packed`` keys), or "int4" / "int8": quantise the decoder blocks of a bf16 state
        dict at load time (quant.quantize_decoder) and stream them packed."""
        cfg.validate()
        if quantize not in (None, "int4", "int8"):
            raise ValueError('quantize must be None, "int4" or "int8"')
        self._quantize = quantize
        if not torch.cuda.is_available():
            raise N.NativeError("moondream_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise N.NativeError(f"moondream_b200 runs on CUDA devices only, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = N.lib()
        with torch.cuda.device(self.device):
            self._init(cfg, state_dict, kv_pages, max_batch)

    def _init(self, cfg, state_dict, kv_pages, max_batch):
        import os as _os0
        if _os0.environ.get("MD_PDL") is not None:          # A/B switches for profiling / validation runs
            self.lib.md_debug_set_pdl(int(_os0.environ["MD_PDL"]))
        if _os0.environ.get("MD_DEBUG_GEMM") is not None:
            self.lib.md_debug_gemm(int(_os0.environ["MD_DEBUG_GEMM"]))
        if _os0.environ.get("MD_ATTENTION_IMPL") is not None:
            self.lib.md_debug_attention_impl(int(_os0.environ["MD_ATTENTION_IMPL"]))
        self.quantized: Optional[Q.QuantizedText] = None
        if Q.is_quantized_checkpoint(state_dict):
            self.quantized, state_dict = Q.from_reference_checkpoint(cfg, state_dict)
        elif self._quantize is not None:
            self.quantized, _ = Q.quantize_decoder(cfg, state_dict, 4 if self._quantize == "int4" else 8)
        qb = self.quantized is not None
        prepared, self.patch_k, self.vis_ff = prepare_weights(cfg, state_dict, quantized_blocks=qb)
        self.weights, self._owners = upload_weights(cfg, prepared, self.device, quantized_blocks=qb)   # keeps device memory alive
        del prepared
        self.lut = pixel_lut().to(self.device)
        self.rope = rope_table(cfg.text.head_dim, cfg.text.max_context).to(self.device)
        v, t, r = cfg.vision, cfg.text, cfg.region
        self.dims = N.md_dims(
            vis_dim=v.enc_dim, vis_ff=self.vis_ff, vis_layers=v.enc_n_layers, vis_heads=v.enc_n_heads,
            crop=v.crop_size, patch=v.enc_patch_size, patch_k=self.patch_k, grid=v.grid,
            margin=v.overlap_margin, proj_inner=v.proj_inner_dim, txt_dim=t.dim, txt_ff=t.ff_dim,
            txt_layers=t.n_layers, txt_heads=t.n_heads, vocab=t.vocab_size, max_context=t.max_context,
            prefix_len=t.prefix_attn, reg_inner=r.inner_dim, coord_feat=r.coord_feat_dim,
            coord_out=r.coord_out_dim, size_feat=r.size_feat_dim, size_out=r.size_out_dim, txt_fused=1,
            txt_kv_heads=t.n_kv_heads)
        n = self.lib.md_model_num_weights(ctypes.byref(self.dims))
        assert n == len(self.weights), (n, len(self.weights))
        arr = (ctypes.c_void_p * n)(*[w.data_ptr() for w in self.weights])
        handle = ctypes.c_void_p()
        N.check(self.lib.md_model_create(ctypes.byref(self.dims), arr, n, N.ptr(self.lut),
                                         N.ptr(self.rope), ctypes.byref(handle)), "md_model_create")
        self.model = handle
        if self.quantized is not None:
            for i in range(t.n_layers):
                dev = [x.to(self.device) for x in self.quantized.fused(i)]
                self._owners += dev
                N.check(self.lib.md_model_set_quantized_block(self.model, i, self.quantized.bits, *[N.ptr(x) for x in dev]),
                        "md_model_set_quantized_block")
        self.max_blocks = t.max_context // PAGE
        if kv_pages is None:
            kv_pages = max_batch * self.max_blocks
        self.pages = PagePool(cfg, kv_pages, self.device)
        self._ws: Optional[torch.Tensor] = None
        self._decode_state: Dict[int, dict] = {}
        self._stage: Optional[torch.Tensor] = None
        self._raw_stage: Optional[torch.Tensor] = None
        self._stage_evt = None
        self._coeffs: Dict[Tuple[int, int], tuple] = {}
        self.preprocess = "device"          # stage_images: resize / crop on the GPU (bit-exact with PIL); "host" = PIL
        import concurrent.futures
        import os as _os
        n_cpu = len(_os.sched_getaffinity(0)) if hasattr(_os, "sched_getaffinity") else (_os.cpu_count() or 1)
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(8, n_cpu)))

    def __del__(self):
        try:
            if getattr(self, "model", None):
                self.lib.md_model_destroy(self.model)
                self.model = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def _kv(self, block_tables: torch.Tensor) -> N.md_kv:
        return N.md_kv(pool=self.pages.pool.data_ptr(), n_pages=self.pages.n_pages,
                       block_tables=block_tables.data_ptr(), max_blocks=block_tables.shape[1],
                       n_layers=self.cfg.text.n_layers, n_kv_heads=self.cfg.text.n_kv_heads)

    def _i32(self, values) -> torch.Tensor:
        return torch.tensor(values, dtype=torch.int32).to(self.device, non_blocking=True)

    def replace_weight(self, key: str, tensor: torch.Tensor):
        """Overwrite one canonical tensor in place on the device (same shape; the decoder's fused [qkv;fc1] /
        [proj|fc2] buffers are views, so they follow).  Tests use it to try another LM head on a loaded model."""
        keys = [k for k, _, _ in state_dict_spec(self.cfg)]
        w = self.weights[keys.index(key)]
        if tuple(w.shape) != tuple(tensor.shape):
            raise ValueError(f"{key}: expected {tuple(w.shape)}, got {tuple(tensor.shape)}")
        with torch.cuda.device(self.device):
            w.copy_(tensor.to(torch.bfloat16))
            torch.cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ vision
    @_on_device
    def vision_encode(self, crops_u8: torch.Tensor) -> torch.Tensor:
        """_vis_enc: uint8 NHWC crops on device -> bf16 [n_crops * 729, enc_dim]."""
        v = self.cfg.vision
        if crops_u8.dtype != torch.uint8 or crops_u8.dim() != 4 or crops_u8.shape[1:] != (v.crop_size, v.crop_size, 3):
            raise ValueError(f"crops must be uint8 [n, {v.crop_size}, {v.crop_size}, 3]")
        crops_u8 = crops_u8.contiguous()
        n = crops_u8.shape[0]
        feats = torch.empty((n * v.tokens_per_crop, v.enc_dim), dtype=torch.bfloat16, device=self.device)
        ws = self._workspace(self.lib.md_vision_encode_workspace_bytes(self.model, n))
        N.check(self.lib.md_vision_encode(self.model, N.ptr(crops_u8), n, N.ptr(feats), N.ptr(ws),
                                          N.current_stream()), "md_vision_encode")
        return feats

    @_on_device
    def vision_project(self, feats: torch.Tensor, crop_offsets: Sequence[int],
                       tilings: Sequence[Tuple[int, int]], embeds: torch.Tensor, rows_per_image: int = 0):
        """reconstruct_from_crops + _vis_proj for all images; fills embeds rows 1..729 of each image."""
        n = len(tilings)
        offs = self._i32(list(crop_offsets))
        til = self._i32([list(x) for x in tilings])
        ws = self._workspace(self.lib.md_vision_project_workspace_bytes(self.model, n))
        N.check(self.lib.md_vision_project(self.model, N.ptr(feats), N.ptr(offs), N.ptr(til), n,
                                           N.ptr(embeds), rows_per_image, N.ptr(ws), N.current_stream()),
                "md_vision_project")

    # ------------------------------------------------------------------ text
    @_on_device
    def embed(self, ids: torch.Tensor, out: torch.Tensor, id_stride: int = 1, n: Optional[int] = None,
              ldo: Optional[int] = None):
        n = ids.numel() if n is None else n
        N.check(self.lib.md_embed_tokens(self.model, N.ptr(ids), id_stride, n, N.ptr(out),
                                         out.stride(0) if ldo is None else ldo, N.current_stream()),
                "md_embed_tokens")

    def load_lora(self, variant) -> "LoraVariant":
        """dict tree (or the flat dotted-key state dict of a variant file) -> LoraVariant on this engine's device."""
        if isinstance(variant, LoraVariant):
            return variant
        if variant and not any(isinstance(v, dict) for v in variant.values()):
            from .synth import nest_lora

            variant = nest_lora(variant)
        return LoraVariant(self.cfg, variant, self.device)

    @_on_device
    def prefill(self, x: torch.Tensor, q_offsets: Sequence[int], start_pos: Sequence[int],
                block_tables: torch.Tensor, prefix_len: int = -1, lora: Optional["LoraVariant"] = None):
        """_prefill over a ragged batch, in place on x [total_tokens, dim].  prefix_len: -1 = the model's 730-token
        bidirectional image prefix (moondream.py:143-145), 0 = pure causal (text-only query, :565-574)."""
        T = x.shape[0]
        n_seqs = len(start_pos)
        assert q_offsets[-1] == T and len(q_offsets) == n_seqs + 1
        max_q = max(q_offsets[i + 1] - q_offsets[i] for i in range(n_seqs))
        qo, sp = self._i32(list(q_offsets)), self._i32(list(start_pos))
        kv = self._kv(block_tables)
        if lora is not None:
            ws = self._workspace(self.lib.md_text_prefill_lora_workspace_bytes(self.model, T, lora.rank))
            N.check(self.lib.md_text_prefill_lora(self.model, N.ptr(x), T, N.ptr(qo), N.ptr(sp), n_seqs, max_q, prefix_len,
                                                  ctypes.byref(kv), lora.table, lora.rank, N.ptr(ws), N.current_stream()),
                    "md_text_prefill_lora")
            return
        ws = self._workspace(self.lib.md_text_prefill_workspace_bytes(self.model, T))
        N.check(self.lib.md_text_prefill(self.model, N.ptr(x), T, N.ptr(qo), N.ptr(sp), n_seqs, max_q, prefix_len,
                                         ctypes.byref(kv), N.ptr(ws), N.current_stream()), "md_text_prefill")

    @_on_device
    def lm_head(self, hidden: torch.Tensor, out_ids: torch.Tensor, out_stride: int, mask_id: int = -1,
                out_index: Optional[torch.Tensor] = None, margins: Optional[torch.Tensor] = None,
                logits: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None,
                out_offset: int = 0, prenormed: bool = False, mask_id2: int = -1):
        B = hidden.shape[0]
        if ws is None:
            ws = self._workspace(self.lib.md_lm_head_workspace_bytes(self.model, B))
        ids_ptr = ctypes.c_void_p(out_ids.data_ptr() + 4 * out_offset)
        mar_ptr = None if margins is None else ctypes.c_void_p(margins.data_ptr() + 4 * out_offset)
        N.check(self.lib.md_lm_head_argmax(self.model, N.ptr(hidden), hidden.stride(0), int(prenormed), B, mask_id, mask_id2,
                                           ids_ptr,
                                           out_stride, N.ptr(out_index), mar_ptr, N.ptr(logits), N.ptr(ws),
                                           N.current_stream()), "md_lm_head_argmax")

    # ------------------------------------------------------------------ image encoding
    @_on_device
    def encode_crops(self, crops_u8: torch.Tensor, crop_offsets: Sequence[int],
                     tilings: Sequence[Tuple[int, int]], return_hidden: bool = False,
                     lora: Optional["LoraVariant"] = None):
        """encode_image (moondream.py:230-268) for a batch whose crops are already on the device:
        ViT -> stitch/pool/project -> [BOS; image] prefill at positions 0..729 into fresh KV pages."""
        t = self.cfg.text
        n_img = len(tilings)
        feats = self.vision_encode(crops_u8)
        embeds = torch.empty((n_img * t.prefix_attn, t.dim), dtype=torch.bfloat16, device=self.device)
        self.vision_project(feats, crop_offsets, tilings, embeds)
        img_emb = embeds.view(n_img, t.prefix_attn, t.dim)[:, 1:].clone() if return_hidden else None
        bos = torch.full((n_img,), self.cfg.tokenizer.bos_id, dtype=torch.int32, device=self.device)
        self.embed(bos, embeds, ldo=t.prefix_attn * t.dim)
        n_prefix_pages = math.ceil(t.prefix_attn / PAGE)
        prefixes = [PrefixKV(t.prefix_attn, self.pages.alloc(n_prefix_pages), self.pages) for _ in range(n_img)]
        bt = torch.zeros((n_img, self.max_blocks), dtype=torch.int32)
        for i, p in enumerate(prefixes):
            bt[i, : len(p.pages)] = torch.tensor(p.pages, dtype=torch.int32)
        bt = bt.to(self.device)
        self.prefill(embeds, [i * t.prefix_attn for i in range(n_img + 1)], [0] * n_img, bt, lora=lora)
        if return_hidden:
            return prefixes, feats, img_emb, embeds
        return prefixes

    @_on_device
    def encode_images(self, images: Sequence[np.ndarray], return_hidden: bool = False,
                      lora: Optional["LoraVariant"] = None):
        """Host uint8 HxWx3 images -> crops (PIL Lanczos, image_crops.py:58-167) -> H2D -> encode_crops."""
        # crops are written straight into persistent pinned staging memory by a small thread pool
        # (PIL's resize and numpy's copies release the GIL)
        dev, offsets, tilings = self.stage_images(images)
        return self.encode_crops(dev, offsets, tilings, return_hidden=return_hidden, lora=lora)

    @_on_device
    def encode_crops_with_prompt(self, crops_u8: torch.Tensor, crop_offsets: Sequence[int],
                                 tilings: Sequence[Tuple[int, int]], prompt: Sequence[Sequence[int]]):
        """encode_image + the prompt prefill of caption()/query() in ONE decoder pass per batch:
        rows [BOS | 729 image tokens | prompt] at positions 0..729+Tp under the same prefix-LM mask
        (moondream.py:138-146, :254-257, :308-310).  Needs equal prompt lengths.  Returns the prefixes
        (pos = 730 + Tp) and the last-token hidden states [n_img, dim] for the LM head."""
        t = self.cfg.text
        n_img = len(tilings)
        Tp = len(prompt[0])
        assert all(len(p) == Tp for p in prompt) and len(prompt) == n_img
        rows = t.prefix_attn + Tp
        feats = self.vision_encode(crops_u8)
        embeds = torch.empty((n_img * rows, t.dim), dtype=torch.bfloat16, device=self.device)
        self.vision_project(feats, crop_offsets, tilings, embeds, rows_per_image=rows)
        bos = torch.full((n_img,), self.cfg.tokenizer.bos_id, dtype=torch.int32, device=self.device)
        self.embed(bos, embeds, ldo=rows * t.dim)
        view = embeds.view(n_img, rows, t.dim)
        ptoks = self._i32([tok for p in prompt for tok in p])
        pemb = torch.empty((n_img * Tp, t.dim), dtype=torch.bfloat16, device=self.device)
        self.embed(ptoks, pemb)
        view[:, t.prefix_attn:].copy_(pemb.view(n_img, Tp, t.dim))
        n_pages = math.ceil(rows / PAGE)
        prefixes = [PrefixKV(rows, self.pages.alloc(n_pages), self.pages) for _ in range(n_img)]
        bt = torch.zeros((n_img, self.max_blocks), dtype=torch.int32)
        for i, p in enumerate(prefixes):
            bt[i, : len(p.pages)] = torch.tensor(p.pages, dtype=torch.int32)
        bt = bt.to(self.device)
        self.prefill(embeds, [i * rows for i in range(n_img + 1)], [0] * n_img, bt)
        hidden_last = view[:, rows - 1].contiguous()
        return prefixes, hidden_last

    @_on_device
    def caption_from_crops(self, crops_u8: torch.Tensor, crop_offsets: Sequence[int],
                           tilings: Sequence[Tuple[int, int]], prompts: Sequence[Sequence[int]], max_tokens: int,
                           to_host: bool = True, stop_on_eos: bool = True, **gen_kw) -> "GenerationResult":
        """encode + generation for a batch; one fused prefill pass when the prompts have equal length.
        `gen_kw`: temperature / top_p / seed of `generate`."""
        if len({len(p) for p in prompts}) == 1:
            prefixes, hidden_last = self.encode_crops_with_prompt(crops_u8, crop_offsets, tilings, prompts)
            return self.generate(prefixes, prompts, max_tokens, consume=True, stop_on_eos=stop_on_eos, to_host=to_host,
                                 prefilled_hidden=hidden_last, **gen_kw)
        prefixes = self.encode_crops(crops_u8, crop_offsets, tilings)
        return self.generate(prefixes, prompts, max_tokens, consume=True, stop_on_eos=stop_on_eos, to_host=to_host,
                             **gen_kw)

    @_on_device
    def stage_images(self, images: Sequence[np.ndarray], preprocess: Optional[str] = None):
        """host uint8 images -> crops on the device; returns (crops, offsets, tilings).
        preprocess "device" (default): images that need a real resize are uploaded raw and resized / cropped by the
        CUDA kernels of csrc/preprocess.cu (bit-exact with PIL's Lanczos, tests/test_resample.py); images that already
        have the crop geometry (378 x 378) are plain copies and take the host staging path.
        preprocess "host": the reference's way for every image — PIL on host threads into pinned staging, then H2D."""
        v = self.cfg.vision
        mode = preprocess or self.preprocess
        if mode not in ("device", "host"):
            raise ValueError("preprocess must be 'device' or 'host'")
        kw = dict(overlap_margin=v.overlap_margin, max_crops=v.max_crops,
                  base_size=(v.crop_size, v.crop_size), patch_size=v.enc_patch_size)
        tilings = [crop_tiling(im.shape, **kw) for im in images]
        offsets = [0]
        for th, tw in tilings:
            offsets.append(offsets[-1] + th * tw + 1)
        n = offsets[-1]
        window = v.crop_size - 2 * v.overlap_margin * v.enc_patch_size
        margin2 = 2 * v.overlap_margin * v.enc_patch_size

        def needs_resize(i):
            th, tw = tilings[i]
            h, w = images[i].shape[:2]
            return (h, w) != (v.crop_size, v.crop_size) or (th * window + margin2, tw * window + margin2) != (h, w)

        if self._stage_evt is not None:
            self._stage_evt.synchronize()       # the previous call's async H2D copies have left the pinned buffers
        on_device = [i for i in range(len(images)) if mode == "device" and needs_resize(i)]
        on_host = [i for i in range(len(images)) if i not in set(on_device)]
        if self._stage is None or self._stage.shape[0] < n:
            self._stage = torch.empty((max(n, 64), v.crop_size, v.crop_size, 3), dtype=torch.uint8).pin_memory()
        stage_np = self._stage.numpy()

        def work(i):
            overlap_crop_image(images[i], out=stage_np[offsets[i]: offsets[i + 1]], **kw)

        if len(on_host) > 1:
            list(self._pool.map(work, on_host))
        elif on_host:
            work(on_host[0])
        crops = self._stage[:n].to(self.device, non_blocking=True)
        if on_device:
            self._preprocess_on_device(images, on_device, tilings, offsets, crops)
        self._stage_evt = torch.cuda.Event()
        self._stage_evt.record()
        return crops, offsets, tilings

    def _coeff_tables(self, in_size: int, out_size: int):
        """Pillow's fixed-point Lanczos tables for one (source, target) length on the device (cached)."""
        from .resample import lanczos_coeffs

        key = (in_size, out_size)
        hit = self._coeffs.pop(key, None)
        if hit is None:
            bounds, kk = lanczos_coeffs(in_size, out_size)
            hit = (torch.from_numpy(bounds).to(self.device), torch.from_numpy(kk).to(self.device), int(kk.shape[1]))
            while len(self._coeffs) >= 64:
                self._coeffs.pop(next(iter(self._coeffs)))
        self._coeffs[key] = hit
        return hit

    def _resample(self, src: torch.Tensor, axis: int, out_size: int, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
        """one pass of PIL's resize on a uint8 [H, W, 3] device image (axis 1 = width, 0 = height)"""
        h, w = int(src.shape[0]), int(src.shape[1])
        bounds, kk, ksize = self._coeff_tables(w if axis == 1 else h, out_size)
        shape = (h, out_size, 3) if axis == 1 else (out_size, w, 3)
        if dst is None:
            dst = torch.empty(shape, dtype=torch.uint8, device=self.device)
        N.check(self.lib.md_resample_u8(N.ptr(src), h, w, axis, N.ptr(bounds), N.ptr(kk), ksize, out_size, N.ptr(dst),
                                        N.current_stream()), "md_resample_u8")
        return dst

    def resize_lanczos(self, image: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
        """PIL.Image.resize((out_w, out_h), LANCZOS) of a uint8 [H, W, 3] device image, bit for bit: horizontal pass
        first, then vertical (libImaging/Resample.c ImagingResampleInner); a pass whose length does not change is
        skipped, as in Pillow."""
        cur = image.contiguous()
        if int(cur.shape[1]) != out_w:
            cur = self._resample(cur, 1, out_w)
        if int(cur.shape[0]) != out_h:
            cur = self._resample(cur, 0, out_h)
        return cur

    def _preprocess_on_device(self, images, which, tilings, offsets, crops: torch.Tensor):
        """overlap_crop_image (image_crops.py:58-167) for images[which] on the device, written into `crops`."""
        v = self.cfg.vision
        S = v.crop_size
        window = S - 2 * v.overlap_margin * v.enc_patch_size
        margin2 = 2 * v.overlap_margin * v.enc_patch_size
        total = sum(int(images[i].size) for i in which)
        if self._raw_stage is None or self._raw_stage.numel() < total:
            self._raw_stage = torch.empty(max(total, 1 << 22), dtype=torch.uint8).pin_memory()
        raw_np = self._raw_stage.numpy()
        at = 0
        spans = []
        for i in which:
            im = np.ascontiguousarray(images[i])
            raw_np[at: at + im.size] = im.reshape(-1)
            spans.append((at, im.shape[0], im.shape[1]))
            at += im.size
        raw = self._raw_stage[:total].to(self.device, non_blocking=True)
        for i, (lo, h, w) in zip(which, spans):
            th, tw = tilings[i]
            img = raw[lo: lo + h * w * 3].view(h, w, 3)
            o = offsets[i]
            g = self.resize_lanczos(img, S, S)                                   # global crop (image_crops.py:146-149)
            crops[o].copy_(g)
            canvas = self.resize_lanczos(img, th * window + margin2, tw * window + margin2)
            N.check(self.lib.md_extract_windows_u8(N.ptr(canvas), int(canvas.shape[0]), int(canvas.shape[1]), th, tw,
                                                   window, S, N.ptr(crops[o + 1]), N.current_stream()),
                    "md_extract_windows_u8")

    @_on_device
    def prefix_kv_tensors(self, prefix: PrefixKV) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Materialise (k, v) [1, heads, pos, 64] per layer like EncodedImage.caches (moondream.py:56-59)."""
        pg = torch.tensor(prefix.pages, dtype=torch.long, device=self.device)
        blk = self.pages.pool[:, pg]                       # [L, P, 2, KVH, 64, 64]
        L, P, _, H, _, _ = blk.shape
        kv = blk.permute(0, 2, 3, 1, 4, 5).reshape(L, 2, H, P * PAGE, 64)[:, :, :, : prefix.pos]
        return [(kv[i, 0].unsqueeze(0).clone(), kv[i, 1].unsqueeze(0).clone()) for i in range(L)]

    # ------------------------------------------------------------------ generation
    def _sequence_tables(self, prefixes: Sequence[PrefixKV], total_len: int, consume: bool):
        """Block tables for sequences that continue `prefixes`.  Full prefix pages are shared; the
        partially filled last prefix page is copied (copy-on-write) unless `consume` hands the
        prefix's pages over to the sequence.  All-or-nothing: the pages the whole batch needs are counted
        before anything is taken, and a failure part-way gives back what was taken and leaves the prefixes'
        ownership untouched."""
        n_blocks = math.ceil(total_len / PAGE)
        if n_blocks > self.max_blocks:
            raise ValueError(f"sequence of {total_len} tokens exceeds max_context {self.cfg.text.max_context}")
        need = sum(max(0, n_blocks - (len(p.pages) if consume else p.pos // PAGE)) for p in prefixes)
        if need > self.pages.free_pages:
            raise N.NativeError(f"KV pool exhausted: this batch needs {need} more pages, {self.pages.free_pages} free "
                                f"(raise kv_pages when constructing the model, or release EncodedImages)")
        bt = torch.zeros((len(prefixes), self.max_blocks), dtype=torch.int32)
        owned: List[List[int]] = []
        taken: List[int] = []
        consumed: List[PrefixKV] = []
        copies_src, copies_dst = [], []
        try:
            for i, p in enumerate(prefixes):
                full = p.pos // PAGE
                if consume:
                    if p._released:
                        raise ValueError("this encoded image's KV pages were already handed to a sequence")
                    pages = list(p.pages)
                    fresh = self.pages.alloc(max(0, n_blocks - len(pages)))
                    taken += fresh
                    p._released = True                      # ownership moves to the sequence
                    consumed.append(p)
                    own = pages + fresh
                    table = own
                else:
                    fresh = self.pages.alloc(max(0, n_blocks - full))
                    taken += fresh
                    if p.pos % PAGE and fresh:
                        copies_src.append(p.pages[full])
                        copies_dst.append(fresh[0])
                    own = fresh
                    table = list(p.pages[:full]) + fresh
                bt[i, : len(table)] = torch.tensor(table, dtype=torch.int32)
                owned.append(own)
        except Exception:
            self.pages.release(taken)
            for p in consumed:
                p._released = False
            raise
        if copies_src:
            src = torch.tensor(copies_src, dtype=torch.long, device=self.device)
            dst = torch.tensor(copies_dst, dtype=torch.long, device=self.device)
            self.pages.pool[:, dst] = self.pages.pool[:, src]
        return bt.to(self.device), owned

    def _decode_buffers(self, B: int) -> dict:
        """Per-batch-size decode state (buffers, workspaces, captured graphs).  The per-step outputs use a fixed
        row stride of max_context + 1 slots, so one set of buffers and graphs serves every max_tokens; at most
        `_MAX_DECODE_STATES` batch sizes stay cached (least recently used goes first)."""
        st = self._decode_state.pop(B, None)
        if st is None:
            t = self.cfg.text
            dev = self.device
            S = t.max_context + 1
            st = {
                "x": torch.empty((B, t.dim), dtype=torch.bfloat16, device=dev),
                "normed": torch.empty((B, t.dim), dtype=torch.bfloat16, device=dev),
                "pos": torch.zeros(B, dtype=torch.int32, device=dev),
                "cur": torch.zeros(B, dtype=torch.int32, device=dev),
                "step": torch.zeros(1, dtype=torch.int32, device=dev),
                "seed": torch.zeros(1, dtype=torch.int64, device=dev),
                "preds": torch.zeros((B, S), dtype=torch.int32, device=dev),
                "forced": torch.zeros((B, S), dtype=torch.int32, device=dev),
                "margins": torch.zeros((B, S), dtype=torch.float32, device=dev),
                "finished": torch.zeros(B, dtype=torch.int32, device=dev),
                "bt": torch.zeros((B, self.max_blocks), dtype=torch.int32, device=dev),
                "ws": torch.empty(int(self.lib.md_text_decode_workspace_bytes(self.model, B)) +
                                  int(self.lib.md_lm_head_workspace_bytes(self.model, B)),
                                  dtype=torch.uint8, device=dev),
                "graphs": {},
                "S": S,
            }
            while len(self._decode_state) >= self._MAX_DECODE_STATES:
                self._decode_state.pop(next(iter(self._decode_state)))
        self._decode_state[B] = st          # most recently used last
        return st

    def _sampling_buffers(self, st: dict, B: int):
        if "logits" not in st:
            V = self.cfg.text.vocab_size
            st["logits"] = torch.empty((B, V), dtype=torch.bfloat16, device=self.device)
            st["probs"] = torch.empty((B, V), dtype=torch.bfloat16, device=self.device)

    def _reasoning_buffers(self, st: dict, B: int):
        if "coords" not in st:
            dev = self.device
            st["coords"] = torch.zeros((B, st["S"]), dtype=torch.float32, device=dev)
            st["coord_bins"] = torch.zeros((B,), dtype=torch.int32, device=dev)
            st["coord_vals"] = torch.zeros((B, 1), dtype=torch.float32, device=dev)
            st["coord_emb"] = torch.empty((B, self.cfg.text.dim), dtype=torch.bfloat16, device=dev)
            st["region_ws"] = torch.empty(int(self.lib.md_region_workspace_bytes(self.model, B)), dtype=torch.uint8,
                                          device=dev)

    def sample_tokens(self, logits: torch.Tensor, temperature: float, top_p: float, out_ids: torch.Tensor,
                      out_stride: int, out_offset: int = 0, step: Optional[torch.Tensor] = None,
                      seed: Optional[torch.Tensor] = None, uniforms: Optional[torch.Tensor] = None,
                      scratch: Optional[torch.Tensor] = None, keep_probs: bool = False):
        """softmax(logits / T) -> _apply_top_p -> one draw per row, on the device (moondream.py:270-278, 312-318)."""
        B, V = logits.shape
        if scratch is None:
            scratch = torch.empty((B, V), dtype=torch.bfloat16, device=self.device)
        N.check(self.lib.md_sample_top_p(N.ptr(logits), B, V, float(temperature), float(top_p), N.ptr(seed), N.ptr(step),
                                         N.ptr(uniforms), N.ptr(scratch), int(keep_probs), N.ptr(out_ids), out_stride,
                                         out_offset, N.current_stream()), "md_sample_top_p")
        return scratch

    def _decode_step_launch(self, st: dict, B: int, mode: "DecodeMode"):
        """One decode step for the whole batch, graph-capturable: [coordinate interleave ->] embed(cur) -> decoder
        blocks -> lm_head (+ masks) -> argmax [-> top-p sample] -> bookkeeping.  No host involvement."""
        lib, s = self.lib, N.current_stream()
        S = st["S"]
        kv = self._kv(st["bt"])
        tk = self.cfg.tokenizer
        if mode.reasoning:
            # _generate_reasoning (moondream.py:381-391): a coord_id token is fed as encode_coordinate(argmax of
            # decode_coordinate(last hidden)) instead of its embedding.  Computed for every row (50 MB of region
            # weights, ~1 % of a step), used by the rows whose current token is coord_id.
            N.check(lib.md_region_decode(self.model, 0, N.ptr(st["x"]), st["x"].stride(0), B, N.ptr(st["coord_bins"]),
                                         N.ptr(st["region_ws"]), s), "md_region_decode")
            N.check(lib.md_region_bins_to_values(0, N.ptr(st["coord_bins"]), B, self.cfg.region.coord_out_dim,
                                                 N.ptr(st["coord_vals"]), s), "md_region_bins_to_values")
            N.check(lib.md_store_column_f32(N.ptr(st["coord_vals"]), B, N.ptr(st["coords"]), S, N.ptr(st["step"]), 0, s),
                    "md_store_column_f32")
            N.check(lib.md_region_encode(self.model, 0, N.ptr(st["coord_vals"]), B, N.ptr(st["coord_emb"]),
                                         st["coord_emb"].stride(0), N.ptr(st["region_ws"]), s), "md_region_encode")
            N.check(lib.md_embed_tokens_select(self.model, N.ptr(st["cur"]), 1, B, tk.coord_id, N.ptr(st["coord_emb"]),
                                               st["coord_emb"].stride(0), N.ptr(st["x"]), st["x"].stride(0), s),
                    "md_embed_tokens_select")
        else:
            self.embed(st["cur"], st["x"])
        N.check(lib.md_text_decode_step(self.model, N.ptr(st["x"]), N.ptr(st["pos"]), B, ctypes.byref(kv),
                                        N.ptr(st["normed"]), N.ptr(st["ws"]), s), "md_text_decode_step")
        off = int(lib.md_text_decode_workspace_bytes(self.model, B))
        self.lm_head(st["normed"], st["preds"], S, mask_id=mode.mask_id, mask_id2=mode.mask_id2, out_index=st["step"],
                     margins=st["margins"], logits=st["logits"] if mode.sampled else None, ws=st["ws"][off:],
                     out_offset=1, prenormed=True)
        if mode.sampled:       # the sampled token replaces the argmax in the slot the bookkeeping reads
            self.sample_tokens(st["logits"], mode.temperature, mode.top_p, st["preds"], S, out_offset=1, step=st["step"],
                               seed=st["seed"], scratch=st["probs"])
        N.check(lib.md_decode_advance(N.ptr(st["cur"]), N.ptr(st["pos"]), N.ptr(st["step"]), N.ptr(st["preds"]),
                                      N.ptr(st["forced"]) if mode.forced else None, S, B,
                                      mode.eos_id, N.ptr(st["finished"]), s), "md_decode_advance")

    def decode_mode(self, forced: bool = False, temperature: float = 0.0, top_p: float = 1.0,
                    reasoning: bool = False) -> "DecodeMode":
        tk = self.cfg.tokenizer
        if reasoning:      # moondream.py:344 eos = answer_id; :395-396 mask eos_id and size_id
            return DecodeMode(forced, float(temperature), float(top_p), True, tk.eos_id, tk.size_id, tk.answer_id)
        return DecodeMode(forced, float(temperature), float(top_p), False, tk.answer_id, -1, tk.eos_id)   # :517

    def _prefill_phase(self, st: dict, prompts, start_pos: Sequence[int], prompt_embeds, prefix_len: int,
                       lora: Optional["LoraVariant"] = None):
        """_prefill_prompt (moondream.py:280-321) for the batch: ragged prompt prefill at `start_pos`, last rows ->
        st["x"].  Returns the prompt lengths."""
        t = self.cfg.text
        B = len(start_pos)
        lens = [len(p) for p in prompts]
        q_off = [0]
        for n in lens:
            q_off.append(q_off[-1] + n)
        if prompt_embeds is None:
            flat = self._i32([tok for p in prompts for tok in p])
            x = torch.empty((q_off[-1], t.dim), dtype=torch.bfloat16, device=self.device)
            self.embed(flat, x)
        else:
            x = prompt_embeds
        self.prefill(x, q_off, list(start_pos), st["bt"], prefix_len=prefix_len, lora=lora)
        last = self._i32([q_off[i + 1] - 1 for i in range(B)])
        N.check(self.lib.md_gather_rows_bf16(N.ptr(x), x.stride(0), N.ptr(last), B, t.dim, N.ptr(st["x"]),
                                             st["x"].stride(0), N.current_stream()), "md_gather_rows_bf16")
        return lens

    def _decode_phase(self, st: dict, B: int, pos0: Sequence[int], max_tokens: int, mode: "DecodeMode",
                      forced=None, use_graph: bool = True, stop_on_eos: bool = True, chunk: int = 0,
                      seed: Optional[int] = None):
        """First token from st["x"] (the prefill's last hidden rows), then `max_tokens` decode steps.  A generator:
        yields (lo, hi) after every `chunk` steps (0 = only once, at the end) so a caller can read preds[:, lo:hi]
        while the loop is still running (streaming); the final yield covers the remainder."""
        S, n_out = st["S"], max_tokens + 1
        if n_out > S:
            raise ValueError(f"max_tokens {max_tokens} exceeds max_context {self.cfg.text.max_context}")
        st["step"].zero_()
        st["finished"].zero_()
        if mode.sampled:
            self._sampling_buffers(st, B)
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())     # torch's global RNG, like the reference's multinomial
            st["seed"].fill_(seed)
        if mode.reasoning:
            self._reasoning_buffers(st, B)
        # first token: no ids are masked at the prefill (moondream.py:312-318)
        self.lm_head(st["x"], st["preds"], S, mask_id=-1, out_index=None, margins=st["margins"],
                     logits=st["logits"] if mode.sampled else None)
        if mode.sampled:
            self.sample_tokens(st["logits"], mode.temperature, mode.top_p, st["preds"], S, out_offset=0, step=None,
                               seed=st["seed"], scratch=st["probs"])
        if mode.forced:
            f = torch.zeros((B, n_out), dtype=torch.int32)
            for i, row in enumerate(forced):
                f[i, : min(len(row), n_out)] = torch.tensor(list(row)[:n_out], dtype=torch.int32)
            st["forced"][:, :n_out].copy_(f.to(self.device))
            st["cur"].copy_(st["forced"][:, 0])
        else:
            st["cur"].copy_(st["preds"][:, 0])
        st["pos"].copy_(self._i32(list(pos0)))
        graph = _lru_get(st["graphs"], mode) if use_graph else None
        if use_graph and graph is None and max_tokens > 0:
            self._decode_step_launch(st, B, mode)           # warm-up (also validates)
            torch.cuda.synchronize()
            # rewind the state the warm-up step advanced
            st["step"].zero_()
            st["pos"].sub_(1)
            st["cur"].copy_(st["forced"][:, 0] if mode.forced else st["preds"][:, 0])
            st["finished"].zero_()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._decode_step_launch(st, B, mode)
            # capture does not execute; state is still at step 0.  An evicted graph is idle: the device was
            # synchronised after the warm-up step above and nothing has been replayed since
            _lru_put(st["graphs"], mode, graph, self._MAX_DECODE_GRAPHS)
        lo = 0
        steps = 0
        for s in range(max_tokens):
            if graph is not None:
                graph.replay()
            else:
                self._decode_step_launch(st, B, mode)
            steps += 1
            if chunk and steps % chunk == 0:
                yield lo, steps                             # slots [lo, steps) are final once this step is queued
                lo = steps
            if stop_on_eos and not mode.forced and (s % 16 == 15) and bool(st["finished"].all().item()):
                break
        st["steps_run"] = steps
        yield lo, steps + 1

    @_on_device
    def generate(self, prefixes: Sequence[PrefixKV], prompts: Sequence[Sequence[int]], max_tokens: int,
                 forced: Optional[Sequence[Sequence[int]]] = None, consume: bool = False,
                 use_graph: bool = True, stop_on_eos: bool = True,
                 prompt_embeds: Optional[torch.Tensor] = None, to_host: bool = True,
                 prefilled_hidden: Optional[torch.Tensor] = None,
                 sampler: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 temperature: float = 0.0, top_p: float = 1.0, seed: Optional[int] = None,
                 prefix_len: int = -1, lora: Optional["LoraVariant"] = None) -> GenerationResult:
        """`_generate_answer` (moondream.py:434-539) for a batch: ragged prompt prefill, first
        token from the LM head, then `max_tokens` decode steps (the reference also runs the step after
        the last emitted token), all inside one CUDA graph per step with no per-token host sync.
        temperature 0: greedy, returns the argmax at every step; callers cut at eos.
        temperature > 0: the reference's softmax / top-p / multinomial on the device (md_sample_top_p), still inside
        the graph; the returned tokens are the sampled ones.
        `sampler` (moondream_b200.sampling.HostSampler) is the host-side restatement of the same arithmetic with torch
        CPU ops (bit-equal to the oracle, one synchronisation per token) kept for parity tests.
        prefix_len 0 = pure causal mask (text-only query, moondream.py:565-574)."""
        B = len(prefixes)
        assert len(prompts) == B
        n_out = max_tokens + 1              # slots the caller gets back: the first token + one per decode step
        lens = [len(p) for p in prompts]
        if prefilled_hidden is not None:
            lens = [0] * B              # prefixes already include the prompt; prefilled_hidden = last-token hidden states
        total = [prefixes[i].pos + lens[i] + max_tokens + 1 for i in range(B)]
        bt, owned = self._sequence_tables(prefixes, max(total), consume)
        try:
            st = self._decode_buffers(B)
            S = st["S"]                     # row stride of preds / forced / margins (fixed, >= n_out)
            st["bt"].copy_(bt)
            if prefilled_hidden is not None:
                st["x"].copy_(prefilled_hidden)
            else:
                self._prefill_phase(st, prompts, [p.pos for p in prefixes], prompt_embeds, prefix_len, lora)
            pos0 = [prefixes[i].pos + lens[i] for i in range(B)]
            if lora is not None:
                if forced is not None or sampler is not None:
                    raise ValueError("generate: `forced` / `sampler` are not combined with a LoRA variant")
                return self._generate_lora(st, pos0, B, max_tokens, lora, temperature, top_p, seed, stop_on_eos, to_host,
                                           prefix_len)
            if sampler is not None:
                if forced is not None:
                    raise ValueError("generate: `forced` and `sampler` are mutually exclusive")
                return self._generate_sampled(st, pos0, B, S, max_tokens, sampler, stop_on_eos, to_host)
            mode = self.decode_mode(forced is not None, temperature, top_p)
            for _ in self._decode_phase(st, B, pos0, max_tokens, mode, forced, use_graph, stop_on_eos, seed=seed):
                pass
            steps = st["steps_run"]
            if to_host:
                tokens = st["preds"][:, :n_out].to("cpu")          # the one device->host read of the call
                margins = st["margins"][:, :n_out].to("cpu")
            else:
                tokens = st["preds"][:, :n_out].clone(memory_format=torch.contiguous_format)
                margins = st["margins"][:, :n_out].clone(memory_format=torch.contiguous_format)
        finally:
            for pages in owned:
                self.pages.release(pages)
        return GenerationResult(tokens, margins, steps)

    def generate_stream(self, prefixes: Sequence[PrefixKV], prompts: Sequence[Sequence[int]], max_tokens: int,
                        chunk: int = 8, temperature: float = 0.0, top_p: float = 1.0, seed: Optional[int] = None,
                        prompt_embeds: Optional[torch.Tensor] = None, prefix_len: int = -1, eos_id: Optional[int] = None):
        """Streaming form of `generate` (the generator of moondream.py:470-537): yields int32 [B, k] host tensors of
        newly decoded tokens every `chunk` graph replays while later steps are still being queued, and stops once every
        sequence has produced `eos_id`.  Closing the generator early (a consumer's `break`) releases the pages.
        A generator cannot hold a `with torch.cuda.device(...)` across its yields without changing the CALLER's current
        device while it is suspended, so every resumed segment enters and leaves the engine's device by itself."""
        B = len(prefixes)
        eos = self.cfg.tokenizer.eos_id if eos_id is None else eos_id
        lens = [len(p) for p in prompts]
        total = [prefixes[i].pos + lens[i] + max_tokens + 1 for i in range(B)]
        with torch.cuda.device(self.device):
            bt, owned = self._sequence_tables(prefixes, max(total), consume=False)
        try:
            with torch.cuda.device(self.device):
                st = self._decode_buffers(B)
                st["bt"].copy_(bt)
                self._prefill_phase(st, prompts, [p.pos for p in prefixes], prompt_embeds, prefix_len)
                pos0 = [prefixes[i].pos + lens[i] for i in range(B)]
                mode = self.decode_mode(False, temperature, top_p)
                phase = self._decode_phase(st, B, pos0, max_tokens, mode, None, True, False, chunk=chunk, seed=seed)
            done = torch.zeros(B, dtype=torch.bool)
            while True:
                with torch.cuda.device(self.device):
                    span = next(phase, None)                    # queues the next `chunk` graph replays
                    if span is None:
                        break
                    lo, hi = span[0], min(span[1], max_tokens)
                    part = st["preds"][:, lo:hi].to("cpu") if hi > lo else None    # waits for the steps queued so far only
                if part is None:
                    continue
                yield part
                done |= (part == eos).any(dim=1)
                if bool(done.all()):
                    break
        finally:
            for pages in owned:
                self.pages.release(pages)

    @_on_device
    def generate_reasoning(self, prefixes: Sequence[PrefixKV], prompts: Sequence[Sequence[int]],
                           answer_prompt: Sequence[int], max_tokens: int, temperature: float = 0.0, top_p: float = 1.0,
                           seed: Optional[int] = None, prompt_embeds: Optional[torch.Tensor] = None,
                           prefix_len: int = -1):
        """query(reasoning=True) (moondream.py:576-596 over _generate_reasoning :323-432): phase 1 decodes the chain
        of thought until answer_id with eos_id / size_id masked, feeding region-encoded coordinates for coord_id
        tokens (on the device, inside the graph); phase 2 prefills `answer_prompt` at each sequence's own position
        and decodes the answer (_generate_answer).  Returns per sequence (reasoning_tokens, coords, answer_tokens):
        coords[j] is the coordinate decoded for reasoning token j when that token is coord_id."""
        t, tk = self.cfg.text, self.cfg.tokenizer
        B = len(prefixes)
        lens = [len(p) for p in prompts]
        total = max(prefixes[i].pos + lens[i] for i in range(B)) + 2 * (max_tokens + 1) + len(answer_prompt)
        total = min(total, t.max_context)
        budget = total - max(prefixes[i].pos + lens[i] for i in range(B)) - len(answer_prompt) - 2
        if budget < 2:
            raise ValueError("the prompt leaves no room for generation within max_context")
        max_tokens = min(max_tokens, budget // 2)
        bt, owned = self._sequence_tables(prefixes, total, consume=False)
        try:
            st = self._decode_buffers(B)
            st["bt"].copy_(bt)
            self._prefill_phase(st, prompts, [p.pos for p in prefixes], prompt_embeds, prefix_len)
            pos0 = [prefixes[i].pos + lens[i] for i in range(B)]
            mode = self.decode_mode(False, temperature, top_p, reasoning=True)
            for _ in self._decode_phase(st, B, pos0, max_tokens, mode, None, True, True, seed=seed):
                pass
            toks = st["preds"][:, : max_tokens + 1].to("cpu")
            coords = st["coords"][:, : max_tokens + 1].to("cpu")
            out_r, out_c, pos1 = [], [], []
            for b in range(B):
                row = toks[b].tolist()
                n = next((j for j, v in enumerate(row[:max_tokens]) if v == tk.answer_id), max_tokens)
                out_r.append(row[:n])
                out_c.append(coords[b, :n].tolist())
                pos1.append(pos0[b] + n)          # every emitted token went through the decoder (moondream.py:398)
            # phase 2: _generate_answer with prompt = answer_prompt at each sequence's position
            self._prefill_phase(st, [list(answer_prompt)] * B, pos1, None, prefix_len)
            pos2 = [p + len(answer_prompt) for p in pos1]
            mode2 = self.decode_mode(False, temperature, top_p)
            for _ in self._decode_phase(st, B, pos2, max_tokens, mode2, None, True, True,
                                        seed=None if seed is None else seed + 1):
                pass
            ans = st["preds"][:, : max_tokens + 1].to("cpu")
            out_a = []
            for b in range(B):
                row = ans[b].tolist()[:max_tokens]
                n = next((j for j, v in enumerate(row) if v == tk.eos_id), len(row))
                out_a.append(row[:n])
        finally:
            for pages in owned:
                self.pages.release(pages)
        return [(out_r[b], out_c[b], out_a[b]) for b in range(B)]

    def _generate_lora(self, st: dict, pos0: Sequence[int], B: int, max_tokens: int, lora: "LoraVariant",
                       temperature: float, top_p: float, seed: Optional[int], stop_on_eos: bool, to_host: bool,
                       prefix_len: int) -> GenerationResult:
        """Decode loop under a LoRA variant (settings["variant"]): each step is md_text_prefill_lora over one row per
        sequence (the fused weight-stream step has no adapter slots).  Eager launches, positions known on the host;
        the token choice (argmax or on-device top-p sampling) and the bookkeeping are the usual kernels."""
        tk = self.cfg.tokenizer
        S, n_out = st["S"], max_tokens + 1
        sampled = temperature > 0
        st["step"].zero_()
        st["finished"].zero_()
        if sampled:
            self._sampling_buffers(st, B)
            st["seed"].fill_(int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else seed)
        self.lm_head(st["x"], st["preds"], S, mask_id=-1, out_index=None, margins=st["margins"],
                     logits=st["logits"] if sampled else None)
        if sampled:
            self.sample_tokens(st["logits"], temperature, top_p, st["preds"], S, out_offset=0, step=None, seed=st["seed"],
                               scratch=st["probs"])
        st["cur"].copy_(st["preds"][:, 0])
        st["pos"].copy_(self._i32(list(pos0)))
        rows = list(range(B + 1))
        steps = 0
        for s in range(max_tokens):
            self.embed(st["cur"], st["x"])
            self.prefill(st["x"], rows, [p + s for p in pos0], st["bt"], prefix_len=prefix_len, lora=lora)
            self.lm_head(st["x"], st["preds"], S, mask_id=tk.answer_id, out_index=st["step"], margins=st["margins"],
                         logits=st["logits"] if sampled else None, out_offset=1)
            if sampled:
                self.sample_tokens(st["logits"], temperature, top_p, st["preds"], S, out_offset=1, step=st["step"],
                                   seed=st["seed"], scratch=st["probs"])
            N.check(self.lib.md_decode_advance(N.ptr(st["cur"]), N.ptr(st["pos"]), N.ptr(st["step"]), N.ptr(st["preds"]),
                                               None, S, B, tk.eos_id, N.ptr(st["finished"]), N.current_stream()),
                    "md_decode_advance")
            steps += 1
            if stop_on_eos and (s % 16 == 15) and bool(st["finished"].all().item()):
                break
        if to_host:
            return GenerationResult(st["preds"][:, :n_out].to("cpu"), st["margins"][:, :n_out].to("cpu"), steps)
        return GenerationResult(st["preds"][:, :n_out].clone(memory_format=torch.contiguous_format),
                                st["margins"][:, :n_out].clone(memory_format=torch.contiguous_format), steps)

    def _generate_sampled(self, st: dict, pos0: Sequence[int], B: int, S: int,
                          max_tokens: int, sampler: Callable[[torch.Tensor], torch.Tensor], stop_on_eos: bool,
                          to_host: bool) -> GenerationResult:
        """Decode loop of `generate` with the next token chosen on the HOST by `sampler` from each step's logits (the
        torch-CPU restatement of the reference's sampler; parity tests).  Uses the teacher-forcing slots: step s
        consumes forced[:, s], its sampled successor is written to forced[:, s + 1] before the bookkeeping kernel
        advances.  (The caller holds and releases the sequences' pages.)"""
        tk = self.cfg.tokenizer
        lib = self.lib
        self._sampling_buffers(st, B)
        logits = st["logits"]
        st["step"].zero_()
        st["finished"].zero_()
        self.lm_head(st["x"], st["preds"], S, mask_id=-1, out_index=None, margins=st["margins"], logits=logits)
        tok = sampler(logits).to(torch.int32)
        done = tok.cpu() == tk.eos_id
        st["forced"].zero_()
        st["forced"][:, 0].copy_(tok)
        st["cur"].copy_(st["forced"][:, 0])
        st["pos"].copy_(self._i32(list(pos0)))
        kv = self._kv(st["bt"])
        off = int(lib.md_text_decode_workspace_bytes(self.model, B))
        steps = 0
        for s in range(max_tokens):
            if stop_on_eos and bool(done.all()):
                break
            stream = N.current_stream()
            self.embed(st["cur"], st["x"])
            N.check(lib.md_text_decode_step(self.model, N.ptr(st["x"]), N.ptr(st["pos"]), B, ctypes.byref(kv),
                                            N.ptr(st["normed"]), N.ptr(st["ws"]), stream), "md_text_decode_step")
            self.lm_head(st["normed"], st["preds"], S, mask_id=tk.answer_id, out_index=st["step"],
                         margins=st["margins"], logits=logits, ws=st["ws"][off:], out_offset=1, prenormed=True)
            tok = sampler(logits).to(torch.int32)                 # synchronises: device logits -> host -> token ids
            done |= tok.cpu() == tk.eos_id
            st["forced"][:, s + 1].copy_(tok)
            N.check(lib.md_decode_advance(N.ptr(st["cur"]), N.ptr(st["pos"]), N.ptr(st["step"]), N.ptr(st["preds"]),
                                          N.ptr(st["forced"]), S, B, tk.eos_id, N.ptr(st["finished"]), stream),
                    "md_decode_advance")
            steps += 1
        n_out = max_tokens + 1
        if to_host:
            return GenerationResult(st["forced"][:, :n_out].to("cpu"), st["margins"][:, :n_out].to("cpu"), steps)
        return GenerationResult(st["forced"][:, :n_out].clone(memory_format=torch.contiguous_format),
                                st["margins"][:, :n_out].clone(memory_format=torch.contiguous_format), steps)

    # ------------------------------------------------------------------ region head
    @_on_device
    def region_encode(self, which: int, values: torch.Tensor) -> torch.Tensor:
        """encode_coordinate / encode_size (region.py:32-43, 60-71): values fp32 [B, 1|2] -> bf16 [B, dim]."""
        values = values.to(self.device, dtype=torch.float32).contiguous()
        B = values.shape[0]
        out = torch.empty((B, self.cfg.text.dim), dtype=torch.bfloat16, device=self.device)
        ws = self._workspace(self.lib.md_region_workspace_bytes(self.model, B))
        N.check(self.lib.md_region_encode(self.model, which, N.ptr(values), B, N.ptr(out), out.stride(0),
                                          N.ptr(ws), N.current_stream()), "md_region_encode")
        return out

    @_on_device
    def region_decode(self, which: int, hidden: torch.Tensor) -> torch.Tensor:
        """argmax bins of decode_coordinate / decode_size (region.py:46-57, 74-93): int32 [B] / [B, 2]."""
        B = hidden.shape[0]
        bins = torch.empty((B,) if which == 0 else (B, 2), dtype=torch.int32, device=self.device)
        ws = self._workspace(self.lib.md_region_workspace_bytes(self.model, B))
        N.check(self.lib.md_region_decode(self.model, which, N.ptr(hidden), hidden.stride(0), B, N.ptr(bins),
                                          N.ptr(ws), N.current_stream()), "md_region_decode")
        return bins

    def _bins_to_values(self, which: int, bins: torch.Tensor) -> torch.Tensor:
        out = torch.empty(bins.shape, dtype=torch.float32, device=self.device)
        N.check(self.lib.md_region_bins_to_values(which, N.ptr(bins), bins.numel(), self.cfg.region.coord_out_dim,
                                                  N.ptr(out), N.current_stream()), "md_region_bins_to_values")
        return out

    def _points_buffers(self, st: dict, B: int):
        if "pt_h" not in st:
            dev, t = self.device, self.cfg.text
            st["pt_h"] = torch.empty((B, t.dim), dtype=torch.bfloat16, device=dev)      # hidden state entering an object
            st["pt_e"] = torch.empty((B, t.dim), dtype=torch.bfloat16, device=dev)
            st["pt_xb"] = torch.zeros((B,), dtype=torch.int32, device=dev)
            st["pt_yb"] = torch.zeros((B,), dtype=torch.int32, device=dev)
            st["pt_sb"] = torch.zeros((B, 2), dtype=torch.int32, device=dev)
            st["pt_xv"] = torch.zeros((B, 1), dtype=torch.float32, device=dev)
            st["pt_yv"] = torch.zeros((B, 1), dtype=torch.float32, device=dev)
            st["pt_sv"] = torch.zeros((B, 2), dtype=torch.float32, device=dev)
            st["pt_nxt"] = torch.zeros((B,), dtype=torch.int32, device=dev)
            st["pt_region_ws"] = torch.empty(int(self.lib.md_region_workspace_bytes(self.model, B)), dtype=torch.uint8,
                                             device=dev)
            st["pt_graphs"] = {}

    def _points_object_launch(self, st: dict, B: int, include_size: bool):
        """One object of `_generate_points` (moondream.py:653-733) for the whole batch, graph-capturable:
        x = decode_coordinate(h); h = decoder(encode_coordinate(x)); y likewise; [size likewise;] next token = lm_head(h)."""
        lib, s = self.lib, N.current_stream()
        r = self.cfg.region
        kv = self._kv(st["bt"])
        H, E, rws = st["pt_h"], st["pt_e"], st["pt_region_ws"]

        def dec(which, hid, bins):
            N.check(lib.md_region_decode(self.model, which, N.ptr(hid), hid.stride(0), B, N.ptr(bins), N.ptr(rws), s),
                    "md_region_decode")

        def b2v(which, bins, out):
            N.check(lib.md_region_bins_to_values(which, N.ptr(bins), bins.numel(), r.coord_out_dim, N.ptr(out), s),
                    "md_region_bins_to_values")

        def enc(which, vals, out):
            N.check(lib.md_region_encode(self.model, which, N.ptr(vals), B, N.ptr(out), out.stride(0), N.ptr(rws), s),
                    "md_region_encode")

        def step(x):                       # in place: x becomes the hidden state after this token
            N.check(lib.md_text_decode_step(self.model, N.ptr(x), N.ptr(st["pos"]), B, ctypes.byref(kv), None,
                                            N.ptr(st["ws"]), s), "md_text_decode_step")
            st["pos"].add_(1)

        dec(0, H, st["pt_xb"]); b2v(0, st["pt_xb"], st["pt_xv"]); enc(0, st["pt_xv"], E); step(E)
        dec(0, E, st["pt_yb"]); b2v(0, st["pt_yb"], st["pt_yv"]); enc(0, st["pt_yv"], H)
        if include_size:
            step(H)
            dec(1, H, st["pt_sb"]); b2v(1, st["pt_sb"], st["pt_sv"]); enc(1, st["pt_sv"], E); step(E)
            H.copy_(E)
        else:
            step(H)
        off = int(lib.md_text_decode_workspace_bytes(self.model, B))
        self.lm_head(H, st["pt_nxt"], 1, ws=st["ws"][off:])

    @_on_device
    def generate_points(self, prefixes: Sequence[PrefixKV], prompts: Sequence[Sequence[int]],
                        include_size: bool, max_objects: int, lora: Optional["LoraVariant"] = None) -> List[List[dict]]:
        """detect / point (moondream.py:735-829 -> _generate_points :653-733) for a batch in lock-step:
        every sequence walks x -> y -> (size) -> next-token together; finished ones are masked on the
        host.  One host sync per object instead of the reference's ~5 .item() calls."""
        t, tk = self.cfg.text, self.cfg.tokenizer
        B = len(prefixes)
        steps_per_obj = 3 if include_size else 2
        lens = [len(p) for p in prompts]
        total = max(prefixes[i].pos + lens[i] + steps_per_obj * max_objects + 1 for i in range(B))
        bt, owned = self._sequence_tables(prefixes, total, consume=False)
        results: List[List[dict]] = [[] for _ in range(B)]
        try:
            q_off = [0]
            for n in lens:
                q_off.append(q_off[-1] + n)
            flat = self._i32([tok for p in prompts for tok in p])
            x = torch.empty((q_off[-1], t.dim), dtype=torch.bfloat16, device=self.device)
            self.embed(flat, x)
            self.prefill(x, q_off, [p.pos for p in prefixes], bt, lora=lora)
            last = self._i32([q_off[i + 1] - 1 for i in range(B)])
            hidden = torch.empty((B, t.dim), dtype=torch.bfloat16, device=self.device)
            N.check(self.lib.md_gather_rows_bf16(N.ptr(x), x.stride(0), N.ptr(last), B, t.dim, N.ptr(hidden),
                                                 hidden.stride(0), N.current_stream()), "md_gather_rows_bf16")
            nxt = torch.empty((B,), dtype=torch.int32, device=self.device)
            self.lm_head(hidden, nxt, 1)
            pos = self._i32([prefixes[i].pos + lens[i] for i in range(B)])
            host_pos = [prefixes[i].pos + lens[i] for i in range(B)]
            active = [True] * B
            tok_host = nxt.tolist()
            for b in range(B):
                if tok_host[b] == tk.eos_id:
                    active[b] = False
            n_obj = 0
            if lora is None:
                # one CUDA graph per object: region decode / encode, the 2 or 3 decode steps between them and the LM head
                # (~300 launches) replay without host involvement; one host sync per object reads the results
                st = self._decode_buffers(B)
                self._points_buffers(st, B)
                st["bt"].zero_()
                st["bt"][:, : bt.shape[1]].copy_(bt)
                st["pos"].copy_(pos)
                st["pt_h"].copy_(hidden)
                graph = st["pt_graphs"].get(include_size)
                if graph is None and any(active) and max_objects > 0:
                    keep = st["pt_h"].clone()
                    self._points_object_launch(st, B, include_size)       # warm-up (also validates)
                    torch.cuda.synchronize()
                    st["pos"].sub_(steps_per_obj)                          # rewind what the warm-up advanced
                    st["pt_h"].copy_(keep)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        self._points_object_launch(st, B, include_size)
                    st["pt_graphs"][include_size] = graph
                while any(active) and n_obj < max_objects:
                    graph.replay()
                    xh, yh, tok_host = st["pt_xv"].view(B).tolist(), st["pt_yv"].view(B).tolist(), st["pt_nxt"].tolist()
                    xbh, ybh = st["pt_xb"].tolist(), st["pt_yb"].tolist()
                    if include_size:
                        sh, sbh = st["pt_sv"].tolist(), st["pt_sb"].tolist()
                    for b in range(B):
                        if not active[b]:
                            continue
                        if include_size:
                            w, h = sh[b]
                            results[b].append({"x_min": xh[b] - w / 2, "y_min": yh[b] - h / 2,
                                               "x_max": xh[b] + w / 2, "y_max": yh[b] + h / 2,
                                               "bins": [xbh[b], ybh[b], sbh[b][0], sbh[b][1]]})
                        else:
                            results[b].append({"x": xh[b], "y": yh[b], "bins": [xbh[b], ybh[b]]})
                        if tok_host[b] == tk.eos_id:
                            active[b] = False
                    n_obj += 1
                return results

            # a LoRA variant: the adapter-aware decoder over one row per sequence, launched eagerly
            def step(emb):
                self.prefill(emb, list(range(B + 1)), list(host_pos), bt, lora=lora)
                pos.add_(1)
                for b_ in range(B):
                    host_pos[b_] += 1
                return emb

            while any(active) and n_obj < max_objects:
                xb = self.region_decode(0, hidden)
                xv = self._bins_to_values(0, xb)
                hidden = step(self.region_encode(0, xv.view(B, 1)))
                yb = self.region_decode(0, hidden)
                yv = self._bins_to_values(0, yb)
                emb = self.region_encode(0, yv.view(B, 1))
                if include_size:
                    hidden = step(emb)
                    sb = self.region_decode(1, hidden)
                    sv = self._bins_to_values(1, sb)
                    emb = self.region_encode(1, sv)
                hidden = step(emb)
                self.lm_head(hidden, nxt, 1)
                # one sync per object
                xh, yh, tok_host = xv.tolist(), yv.tolist(), nxt.tolist()
                xbh, ybh = xb.tolist(), yb.tolist()
                if include_size:
                    sh, sbh = sv.tolist(), sb.tolist()
                for b in range(B):
                    if not active[b]:
                        continue
                    if include_size:
                        w, h = sh[b]
                        results[b].append({"x_min": xh[b] - w / 2, "y_min": yh[b] - h / 2,
                                           "x_max": xh[b] + w / 2, "y_max": yh[b] + h / 2,
                                           "bins": [xbh[b], ybh[b], sbh[b][0], sbh[b][1]]})
                    else:
                        results[b].append({"x": xh[b], "y": yh[b], "bins": [xbh[b], ybh[b]]})
                    if tok_host[b] == tk.eos_id:
                        active[b] = False
                n_obj += 1
        finally:
            for pages in owned:
                self.pages.release(pages)
        return results

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


def prepare_weights(cfg: MoondreamConfig, sd: Dict[str, torch.Tensor]) -> Tuple[List[torch.Tensor], int, int]:
    """One-time re-layout so every GEMM operand is TMA-addressable (16-byte row pitch):
    patch_emb K 588 -> 592 and the ViT MLP width to a multiple of 8 (0.5B: 2690 -> 2696) with zero
    padding, which leaves the arithmetic unchanged (gelu(0) = 0 meets zero fc2 columns)."""
    v = cfg.vision
    patch_k = _round_up(v.patch_dim, 8)
    vis_ff = _round_up(v.enc_ff_dim, 8)
    out: List[torch.Tensor] = []
    for key, shape, _ in state_dict_spec(cfg):
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


def upload_weights(cfg: MoondreamConfig, prepared: List[torch.Tensor], device) -> Tuple[List[torch.Tensor], list]:
    """Upload in canonical order.  The decoder blocks use the fused decode layout the C runtime checks
    (md_dims.txt_fused): W1 = [qkv.weight ; fc1.weight], b1 = [qkv.bias ; fc1.bias] and
    W2 = [proj.weight | fc2.weight]; the canonical entries become views of those buffers, so prefill
    (separate GEMMs) and decode (one weight stream per pair) share the same memory."""
    keys = [k for k, _, _ in state_dict_spec(cfg)]
    idx = {k: i for i, k in enumerate(keys)}
    dev: List[Optional[torch.Tensor]] = [None] * len(keys)
    owners = []
    D = cfg.text.dim
    for i in range(cfg.text.n_layers):
        p = f"text.blocks.{i}."
        w1 = torch.cat([prepared[idx[p + "attn.qkv.weight"]], prepared[idx[p + "mlp.fc1.weight"]]], 0).to(device)
        b1 = torch.cat([prepared[idx[p + "attn.qkv.bias"]], prepared[idx[p + "mlp.fc1.bias"]]], 0).to(device)
        w2 = torch.cat([prepared[idx[p + "attn.proj.weight"]], prepared[idx[p + "mlp.fc2.weight"]]], 1).to(device)
        owners += [w1, b1, w2]
        dev[idx[p + "attn.qkv.weight"]] = w1[: 3 * D]
        dev[idx[p + "mlp.fc1.weight"]] = w1[3 * D:]
        dev[idx[p + "attn.qkv.bias"]] = b1[: 3 * D]
        dev[idx[p + "mlp.fc1.bias"]] = b1[3 * D:]
        dev[idx[p + "attn.proj.weight"]] = w2[:, :D]
        dev[idx[p + "mlp.fc2.weight"]] = w2[:, D:]
    for i, t in enumerate(prepared):
        if dev[i] is None:
            dev[i] = t.to(device)
    return dev, owners  # type: ignore[return-value]


class PagePool:
    """KV pages: bf16 [layers, n_pages, 2, heads, 64, 64]; a free list hands out page ids."""

    def __init__(self, cfg: MoondreamConfig, n_pages: int, device):
        t = cfg.text
        self.n_pages = n_pages
        # zeros, not empty: masked / not-yet-written slots still flow through P.V as 0 * v and must be finite
        self.pool = torch.zeros((t.n_layers, n_pages, 2, t.n_heads, PAGE, 64), dtype=torch.bfloat16,
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


@dataclass
class GenerationResult:
    tokens: torch.Tensor      # int32 [B, max_tokens + 1] on host: greedy prediction at every step
    margins: Optional[torch.Tensor]
    steps: int


class Engine:
    _MAX_DECODE_STATES = 4

    def __init__(self, cfg: MoondreamConfig, state_dict: Dict[str, torch.Tensor], device="cuda",
                 kv_pages: Optional[int] = None, max_batch: int = 32):
        cfg.validate()
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
        if _os0.environ.get("MD_PDL") is not None:          # A/B switch for profiling runs
            self.lib.md_debug_set_pdl(int(_os0.environ["MD_PDL"]))
        prepared, self.patch_k, self.vis_ff = prepare_weights(cfg, state_dict)
        self.weights, self._owners = upload_weights(cfg, prepared, self.device)   # keeps device memory alive
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
            coord_out=r.coord_out_dim, size_feat=r.size_feat_dim, size_out=r.size_out_dim, txt_fused=1)
        n = self.lib.md_model_num_weights(ctypes.byref(self.dims))
        assert n == len(self.weights), (n, len(self.weights))
        arr = (ctypes.c_void_p * n)(*[w.data_ptr() for w in self.weights])
        handle = ctypes.c_void_p()
        N.check(self.lib.md_model_create(ctypes.byref(self.dims), arr, n, N.ptr(self.lut),
                                         N.ptr(self.rope), ctypes.byref(handle)), "md_model_create")
        self.model = handle
        self.max_blocks = t.max_context // PAGE
        if kv_pages is None:
            kv_pages = max_batch * self.max_blocks
        self.pages = PagePool(cfg, kv_pages, self.device)
        self._ws: Optional[torch.Tensor] = None
        self._decode_state: Dict[int, dict] = {}
        self._stage: Optional[torch.Tensor] = None
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
                       n_layers=self.cfg.text.n_layers)

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

    @_on_device
    def prefill(self, x: torch.Tensor, q_offsets: Sequence[int], start_pos: Sequence[int],
                block_tables: torch.Tensor):
        """_prefill over a ragged batch, in place on x [total_tokens, dim]."""
        T = x.shape[0]
        n_seqs = len(start_pos)
        assert q_offsets[-1] == T and len(q_offsets) == n_seqs + 1
        max_q = max(q_offsets[i + 1] - q_offsets[i] for i in range(n_seqs))
        qo, sp = self._i32(list(q_offsets)), self._i32(list(start_pos))
        ws = self._workspace(self.lib.md_text_prefill_workspace_bytes(self.model, T))
        kv = self._kv(block_tables)
        N.check(self.lib.md_text_prefill(self.model, N.ptr(x), T, N.ptr(qo), N.ptr(sp), n_seqs, max_q,
                                         ctypes.byref(kv), N.ptr(ws), N.current_stream()), "md_text_prefill")

    @_on_device
    def lm_head(self, hidden: torch.Tensor, out_ids: torch.Tensor, out_stride: int, mask_id: int = -1,
                out_index: Optional[torch.Tensor] = None, margins: Optional[torch.Tensor] = None,
                logits: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None,
                out_offset: int = 0, prenormed: bool = False):
        B = hidden.shape[0]
        if ws is None:
            ws = self._workspace(self.lib.md_lm_head_workspace_bytes(self.model, B))
        ids_ptr = ctypes.c_void_p(out_ids.data_ptr() + 4 * out_offset)
        mar_ptr = None if margins is None else ctypes.c_void_p(margins.data_ptr() + 4 * out_offset)
        N.check(self.lib.md_lm_head_argmax(self.model, N.ptr(hidden), hidden.stride(0), int(prenormed), B, mask_id, ids_ptr,
                                           out_stride, N.ptr(out_index), mar_ptr, N.ptr(logits), N.ptr(ws),
                                           N.current_stream()), "md_lm_head_argmax")

    # ------------------------------------------------------------------ image encoding
    @_on_device
    def encode_crops(self, crops_u8: torch.Tensor, crop_offsets: Sequence[int],
                     tilings: Sequence[Tuple[int, int]], return_hidden: bool = False):
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
        self.prefill(embeds, [i * t.prefix_attn for i in range(n_img + 1)], [0] * n_img, bt)
        if return_hidden:
            return prefixes, feats, img_emb, embeds
        return prefixes

    @_on_device
    def encode_images(self, images: Sequence[np.ndarray], return_hidden: bool = False):
        """Host uint8 HxWx3 images -> crops (PIL Lanczos, image_crops.py:58-167) -> H2D -> encode_crops."""
        # crops are written straight into persistent pinned staging memory by a small thread pool
        # (PIL's resize and numpy's copies release the GIL)
        dev, offsets, tilings = self.stage_images(images)
        return self.encode_crops(dev, offsets, tilings, return_hidden=return_hidden)

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
                           to_host: bool = True, stop_on_eos: bool = True) -> "GenerationResult":
        """encode + greedy generation for a batch; one fused prefill pass when the prompts have equal length."""
        if len({len(p) for p in prompts}) == 1:
            prefixes, hidden_last = self.encode_crops_with_prompt(crops_u8, crop_offsets, tilings, prompts)
            return self.generate(prefixes, prompts, max_tokens, consume=True, stop_on_eos=stop_on_eos, to_host=to_host,
                                 prefilled_hidden=hidden_last)
        prefixes = self.encode_crops(crops_u8, crop_offsets, tilings)
        return self.generate(prefixes, prompts, max_tokens, consume=True, stop_on_eos=stop_on_eos, to_host=to_host)

    @_on_device
    def stage_images(self, images: Sequence[np.ndarray]):
        """host uint8 images -> crops in pinned staging memory -> device; returns (crops, offsets, tilings)"""
        v = self.cfg.vision
        kw = dict(overlap_margin=v.overlap_margin, max_crops=v.max_crops,
                  base_size=(v.crop_size, v.crop_size), patch_size=v.enc_patch_size)
        tilings = [crop_tiling(im.shape, **kw) for im in images]
        offsets = [0]
        for th, tw in tilings:
            offsets.append(offsets[-1] + th * tw + 1)
        n = offsets[-1]
        if self._stage is None or self._stage.shape[0] < n:
            self._stage = torch.empty((max(n, 64), v.crop_size, v.crop_size, 3), dtype=torch.uint8).pin_memory()
        stage_np = self._stage.numpy()

        def work(i):
            overlap_crop_image(images[i], out=stage_np[offsets[i]: offsets[i + 1]], **kw)

        if len(images) > 1:
            list(self._pool.map(work, range(len(images))))
        else:
            work(0)
        return self._stage[:n].to(self.device, non_blocking=True), offsets, tilings

    @_on_device
    def prefix_kv_tensors(self, prefix: PrefixKV) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Materialise (k, v) [1, heads, pos, 64] per layer like EncodedImage.caches (moondream.py:56-59)."""
        pg = torch.tensor(prefix.pages, dtype=torch.long, device=self.device)
        blk = self.pages.pool[:, pg]                       # [L, P, 2, H, 64, 64]
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

    def _decode_step_launch(self, st: dict, B: int, S: int, use_forced: bool, mask_id: int):
        """embed(cur) -> 24 blocks -> lm_head + argmax -> bookkeeping; graph-capturable."""
        lib, s = self.lib, N.current_stream()
        kv = self._kv(st["bt"])
        self.embed(st["cur"], st["x"])
        N.check(lib.md_text_decode_step(self.model, N.ptr(st["x"]), N.ptr(st["pos"]), B, ctypes.byref(kv),
                                        N.ptr(st["normed"]), N.ptr(st["ws"]), s), "md_text_decode_step")
        off = int(lib.md_text_decode_workspace_bytes(self.model, B))
        self.lm_head(st["normed"], st["preds"], S, mask_id=mask_id, out_index=st["step"], margins=st["margins"],
                     ws=st["ws"][off:], out_offset=1, prenormed=True)
        N.check(lib.md_decode_advance(N.ptr(st["cur"]), N.ptr(st["pos"]), N.ptr(st["step"]), N.ptr(st["preds"]),
                                      N.ptr(st["forced"]) if use_forced else None, S, B,
                                      self.cfg.tokenizer.eos_id, N.ptr(st["finished"]), s), "md_decode_advance")

    @_on_device
    def generate(self, prefixes: Sequence[PrefixKV], prompts: Sequence[Sequence[int]], max_tokens: int,
                 forced: Optional[Sequence[Sequence[int]]] = None, consume: bool = False,
                 use_graph: bool = True, stop_on_eos: bool = True,
                 prompt_embeds: Optional[torch.Tensor] = None, to_host: bool = True,
                 prefilled_hidden: Optional[torch.Tensor] = None,
                 sampler: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> GenerationResult:
        """`_generate_answer` (moondream.py:434-539) for a batch: ragged prompt prefill, first
        token from the LM head, then `max_tokens` decode steps (the reference also runs the step after
        the last emitted token).  Greedy by default: returns the argmax at every step; callers cut at eos.
        `sampler` (moondream_b200.sampling.HostSampler) switches to the reference's temperature / top-p
        sampling (:312-318, :524-530): every step hands its bf16 logits [B, vocab] to the callable, which
        returns the tokens to feed next; the returned tokens are then the sampled ones.  That path
        synchronises once per token and does not use the CUDA graph."""
        t, tk = self.cfg.text, self.cfg.tokenizer
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
            if n_out > S:
                raise ValueError(f"max_tokens {max_tokens} exceeds max_context {t.max_context}")
            st["bt"].copy_(bt)
            # ---- prompt prefill (moondream.py:280-321) ----
            if prefilled_hidden is not None:
                st["x"].copy_(prefilled_hidden)
            else:
                q_off = [0]
                for n in lens:
                    q_off.append(q_off[-1] + n)
                if prompt_embeds is None:
                    flat = self._i32([tok for p in prompts for tok in p])
                    x = torch.empty((q_off[-1], t.dim), dtype=torch.bfloat16, device=self.device)
                    self.embed(flat, x)
                else:
                    x = prompt_embeds
                self.prefill(x, q_off, [p.pos for p in prefixes], st["bt"])
                last = self._i32([q_off[i + 1] - 1 for i in range(B)])
                N.check(self.lib.md_gather_rows_bf16(N.ptr(x), x.stride(0), N.ptr(last), B, t.dim, N.ptr(st["x"]),
                                                     st["x"].stride(0), N.current_stream()), "md_gather_rows_bf16")
            st["step"].zero_()
            st["finished"].zero_()
            if sampler is not None:
                if forced is not None:
                    raise ValueError("generate: `forced` and `sampler` are mutually exclusive")
                return self._generate_sampled(st, prefixes, lens, B, S, max_tokens, sampler, stop_on_eos, to_host)
            self.lm_head(st["x"], st["preds"], S, mask_id=-1, out_index=None, margins=st["margins"])
            use_forced = forced is not None
            if use_forced:
                f = torch.zeros((B, n_out), dtype=torch.int32)
                for i, row in enumerate(forced):
                    f[i, : min(len(row), n_out)] = torch.tensor(list(row)[:n_out], dtype=torch.int32)
                st["forced"][:, :n_out].copy_(f.to(self.device))
                st["cur"].copy_(st["forced"][:, 0])
            else:
                st["cur"].copy_(st["preds"][:, 0])
            st["pos"].copy_(self._i32([prefixes[i].pos + lens[i] for i in range(B)]))
            # ---- decode loop (moondream.py:481-530), one graph replay per token ----
            gkey = (use_forced, tk.answer_id)
            graph = st["graphs"].get(gkey) if use_graph else None
            if use_graph and graph is None and max_tokens > 0:
                self._decode_step_launch(st, B, S, use_forced, tk.answer_id)   # warm-up (also validates)
                torch.cuda.synchronize()
                # rewind the state the warm-up step advanced
                st["step"].zero_()
                st["pos"].sub_(1)
                st["cur"].copy_(st["forced"][:, 0] if use_forced else st["preds"][:, 0])
                st["finished"].zero_()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._decode_step_launch(st, B, S, use_forced, tk.answer_id)
                st["graphs"][gkey] = graph
                # capture does not execute; state is still at step 0
            steps = 0
            for s in range(max_tokens):
                if graph is not None:
                    graph.replay()
                else:
                    self._decode_step_launch(st, B, S, use_forced, tk.answer_id)
                steps += 1
                if stop_on_eos and not use_forced and (s % 16 == 15) and bool(st["finished"].all().item()):
                    break
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

    def _generate_sampled(self, st: dict, prefixes: Sequence[PrefixKV], lens: Sequence[int], B: int, S: int,
                          max_tokens: int, sampler: Callable[[torch.Tensor], torch.Tensor], stop_on_eos: bool,
                          to_host: bool) -> GenerationResult:
        """Decode loop of `generate` with the next token chosen by `sampler` from each step's logits.  Uses the
        teacher-forcing slots: step s consumes forced[:, s], its sampled successor is written to forced[:, s + 1]
        before the bookkeeping kernel advances.  (The caller holds and releases the sequences' pages.)"""
        tk = self.cfg.tokenizer
        lib = self.lib
        logits = st.get("logits")
        if logits is None:
            logits = st["logits"] = torch.empty((B, self.cfg.text.vocab_size), dtype=torch.bfloat16, device=self.device)
        self.lm_head(st["x"], st["preds"], S, mask_id=-1, out_index=None, margins=st["margins"], logits=logits)
        tok = sampler(logits).to(torch.int32)
        done = tok.cpu() == tk.eos_id
        st["forced"].zero_()
        st["forced"][:, 0].copy_(tok)
        st["cur"].copy_(st["forced"][:, 0])
        st["pos"].copy_(self._i32([prefixes[i].pos + lens[i] for i in range(B)]))
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
        N.check(self.lib.md_region_bins_to_values(which, N.ptr(bins), bins.numel(), N.ptr(out),
                                                  N.current_stream()), "md_region_bins_to_values")
        return out

    @_on_device
    def generate_points(self, prefixes: Sequence[PrefixKV], prompts: Sequence[Sequence[int]],
                        include_size: bool, max_objects: int) -> List[List[dict]]:
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
            self.prefill(x, q_off, [p.pos for p in prefixes], bt)
            last = self._i32([q_off[i + 1] - 1 for i in range(B)])
            hidden = torch.empty((B, t.dim), dtype=torch.bfloat16, device=self.device)
            N.check(self.lib.md_gather_rows_bf16(N.ptr(x), x.stride(0), N.ptr(last), B, t.dim, N.ptr(hidden),
                                                 hidden.stride(0), N.current_stream()), "md_gather_rows_bf16")
            nxt = torch.empty((B,), dtype=torch.int32, device=self.device)
            self.lm_head(hidden, nxt, 1)
            pos = self._i32([prefixes[i].pos + lens[i] for i in range(B)])
            ws = torch.empty(int(self.lib.md_text_decode_workspace_bytes(self.model, B)), dtype=torch.uint8,
                             device=self.device)
            kv = self._kv(bt)

            def step(emb):
                N.check(self.lib.md_text_decode_step(self.model, N.ptr(emb), N.ptr(pos), B, ctypes.byref(kv),
                                                     None, N.ptr(ws), N.current_stream()), "md_text_decode_step")
                pos.add_(1)
                return emb

            active = [True] * B
            tok_host = nxt.tolist()
            for b in range(B):
                if tok_host[b] == tk.eos_id:
                    active[b] = False
            n_obj = 0
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

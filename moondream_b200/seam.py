"""The reference's own swap seam, with its tensor signatures, on top of the native engine.

``MoondreamModel`` in the reference routes all heavy arithmetic through four bound methods
(moondream/torch/moondream.py:168-192) that ``compile()`` rebinds (:194-204):

    _vis_enc(x)                                  bf16 [n_crops, 3, 378, 378]  -> [n_crops, 729, enc_dim]
    _vis_proj(g, r)                              [729, enc_dim], [h, w, enc_dim] -> [729, text_dim]
    _prefill(x, attn_mask, pos_ids, lora)        [1, T, dim] -> hidden [1, T, dim]      (writes the KV cache)
    _decode_one_tok(x, attn_mask, pos_ids, lora) [1, 1, dim] -> (logits [1, vocab], hidden [1, 1, dim])

``SeamAdapter`` exposes exactly these (batch 1, CUDA tensors in and out), so a maintainer can rebind the reference's
attributes to it and every caller above the seam — ``encode_image``, ``_prefill_prompt``,
the generators, ``_generate_points`` — keeps working unchanged.  Like the reference's model object, an adapter holds
ONE mutable KV cache (one sequence's pages) and is not re-entrant.  The batched product path does not go through here.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _native as N
from .engine import PAGE, Engine


class SeamAdapter:
    def __init__(self, engine: Engine):
        self.eng = engine
        self._pages = None
        self._bt: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ KV cache of the one sequence
    def _table(self) -> torch.Tensor:
        if self._bt is None:
            e = self.eng
            self._pages = e.pages.alloc(e.max_blocks)
            self._bt = torch.tensor([self._pages], dtype=torch.int32, device=e.device)
        return self._bt

    def release(self):
        if self._pages is not None:
            self.eng.pages.release(self._pages)
            self._pages, self._bt = None, None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _setup_caches(self):
        """moondream.py:148-166: the reference re-creates zeroed caches; stale entries here are never attended to
        (keys beyond the current position are masked by construction), so keeping the pages is equivalent."""
        return None

    # ------------------------------------------------------------------ the four seams
    def _vis_enc(self, x: torch.Tensor) -> torch.Tensor:
        """vision_encoder (vision.py:64-74).  `x` is prepare_crops' output (vision.py:25-41), i.e. pixel values through
        the reference's bf16 normalisation chain; the chain is injective on 0..255, so the uint8 pixels are recovered
        exactly and fed to the native path (which applies the same chain through a 256-entry table)."""
        e = self.eng
        v = e.cfg.vision
        if x.dim() != 4 or x.shape[1:] != (3, v.crop_size, v.crop_size):
            raise ValueError(f"_vis_enc expects [n, 3, {v.crop_size}, {v.crop_size}]")
        with torch.cuda.device(e.device):
            u8 = (x.to(e.device, torch.float32) * 127.5 + 127.5).round_().clamp_(0, 255).to(torch.uint8)
            feats = e.vision_encode(u8.permute(0, 2, 3, 1).contiguous())
        return feats.view(x.shape[0], v.tokens_per_crop, v.enc_dim)

    def _vis_proj(self, g: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
        """vision_projection (vision.py:77-89): global features [729, D] + reconstructed map [h, w, D]."""
        e = self.eng
        v, t = e.cfg.vision, e.cfg.text
        g = g.to(e.device, torch.bfloat16).contiguous()
        r = r.to(e.device, torch.bfloat16).contiguous()
        if g.shape != (v.tokens_per_crop, v.enc_dim) or r.dim() != 3 or r.shape[2] != v.enc_dim:
            raise ValueError("_vis_proj expects g [729, enc_dim] and r [h, w, enc_dim]")
        with torch.cuda.device(e.device):
            out = torch.empty((v.tokens_per_crop, t.dim), dtype=torch.bfloat16, device=e.device)
            ws = e._workspace(e.lib.md_vision_project_workspace_bytes(e.model, 1))
            N.check(e.lib.md_vision_project_stitched(e.model, N.ptr(g), N.ptr(r), r.shape[0], r.shape[1], N.ptr(out),
                                                     N.ptr(ws), N.current_stream()), "md_vision_project_stitched")
        return out

    def _positions(self, pos_ids: torch.Tensor, T: int) -> int:
        p0 = int(pos_ids.flatten()[0].item())
        if pos_ids.numel() != T:
            raise ValueError("pos_ids must hold one position per input row")
        return p0

    def _prefill(self, x: torch.Tensor, attn_mask: torch.Tensor, pos_ids: torch.Tensor, lora=None) -> torch.Tensor:
        """text_decoder (text.py:128-160) over T rows at consecutive positions pos_ids; K/V go to this adapter's
        pages.  The mask the reference passes is either the prefix-LM one (moondream.py:138-146) or the plain causal
        one of a text-only query (:571-574); they differ only for rows inside the 730-token prefix, where the first
        row of the prefix-LM mask also sees later positions — that is how the two are told apart."""
        if lora is not None:
            raise NotImplementedError("LoRA variants are not supported")
        e = self.eng
        t = e.cfg.text
        if x.dim() != 3 or x.shape[0] != 1 or x.shape[2] != t.dim:
            raise ValueError("_prefill expects [1, T, dim]")
        T = x.shape[1]
        p0 = self._positions(pos_ids, T)
        prefix_len = -1
        # the two masks can only differ for a row inside the prefix looking at a LATER position that is still inside
        # the prefix: test the first row against the last such column (rows that reach past the prefix, e.g. a fused
        # [BOS | image | prompt] pass, are classified by their in-prefix part)
        last_in_prefix = min(p0 + T - 1, t.prefix_attn - 1)
        if p0 < t.prefix_attn and last_in_prefix > p0:
            bidirectional = bool(attn_mask[0, 0, 0, last_in_prefix].item()) if attn_mask is not None else True
            prefix_len = -1 if bidirectional else 0
        with torch.cuda.device(e.device):
            h = x[0].to(e.device, torch.bfloat16).contiguous().clone()
            e.prefill(h, [0, T], [p0], self._table(), prefix_len=prefix_len)
        return h.unsqueeze(0)

    def _decode_one_tok(self, x: torch.Tensor, attn_mask: torch.Tensor, pos_ids: torch.Tensor,
                        lora=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """text_decoder at T = 1 + lm_head (moondream.py:183-192): returns (logits [1, vocab], hidden [1, 1, dim])."""
        if lora is not None:
            raise NotImplementedError("LoRA variants are not supported")
        e = self.eng
        t = e.cfg.text
        p0 = self._positions(pos_ids, 1)
        with torch.cuda.device(e.device):
            h = x.reshape(1, t.dim).to(e.device, torch.bfloat16).contiguous().clone()
            pos = torch.tensor([p0], dtype=torch.int32, device=e.device)
            normed = torch.empty((1, t.dim), dtype=torch.bfloat16, device=e.device)
            ws = e._workspace(e.lib.md_text_decode_workspace_bytes(e.model, 1))
            kv = e._kv(self._table())
            N.check(e.lib.md_text_decode_step(e.model, N.ptr(h), N.ptr(pos), 1, ctypes.byref(kv), N.ptr(normed),
                                              N.ptr(ws), N.current_stream()), "md_text_decode_step")
            logits = torch.empty((1, t.vocab_size), dtype=torch.bfloat16, device=e.device)
            ids = torch.empty((1,), dtype=torch.int32, device=e.device)
            e.lm_head(normed, ids, 1, logits=logits, prenormed=True)
        return logits, h.view(1, 1, t.dim)

    # ------------------------------------------------------------------ what the reference calls around the seams
    def load_kv_prefix(self, caches) -> None:
        """load_encoded_image (moondream.py:620-623): copy per-layer (k, v) [1, kv_heads, n, 64] into the pages."""
        e = self.eng
        bt = self._table()
        pool = e.pages.pool
        for layer, (k, v) in enumerate(caches):
            n = k.shape[2]
            for lo in range(0, n, PAGE):
                hi = min(n, lo + PAGE)
                page = int(self._pages[lo // PAGE])
                pool[layer, page, 0, :, : hi - lo] = k[0, :, lo:hi].to(e.device, torch.bfloat16)
                pool[layer, page, 1, :, : hi - lo] = v[0, :, lo:hi].to(e.device, torch.bfloat16)

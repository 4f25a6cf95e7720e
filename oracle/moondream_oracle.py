"""TEST INFRASTRUCTURE — CPU restatement ("port") of the reference algorithm for the hot path.

Plain torch CPU ops over a flat dict of tensors in the canonical state_dict layout.  With
``dtype=torch.bfloat16`` every op rounds where the reference rounds (the reference hard-wires bf16:
vision.py:36, weights.py:32), which makes this port bit-identical to the unmodified reference on
the same torch build — ``tests/test_oracle.py`` and ``tests/test_oracle_r2.py`` check exactly that in the build
container (tiny presets, and the Moondream-2B / 0.5B architectures, the 2B with the bench's weights) and ``tests/golden/*.json``
carries the reference's outputs to the GPU box.  With ``dtype=torch.float32`` it is the "truth" used for error budgeting.

Parity pinning: PINNED.  The reference's own tests hold no model-path vectors (only tests/test_image_crops.py, re-hosted
in tests/test_image_crops.py), so the pins are (a) bit-equality with the unmodified reference run in the same process
(KV caches, prefill logits and hidden states, tokens, boxes, grounding) and (b) the committed golden fixtures generated
from the reference by oracle/make_golden.py, make_golden_r2.py and make_golden_quant.py.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# image crops (moondream/torch/image_crops.py)
# --------------------------------------------------------------------------------------------
def select_tiling(height: int, width: int, crop_size: int, max_crops: int) -> Tuple[int, int]:
    """image_crops.py:17-50."""
    if height <= crop_size or width <= crop_size:
        return 1, 1
    need_h, need_w = math.ceil(height / crop_size), math.ceil(width / crop_size)
    if need_h * need_w > max_crops:
        shrink = math.sqrt(max_crops / (need_h * need_w))
        return max(1, math.floor(need_h * shrink)), max(1, math.floor(need_w * shrink))
    th = max(math.floor(math.sqrt(max_crops * height / width)), need_h)
    tw = max(math.floor(math.sqrt(max_crops * width / height)), need_w)
    if th * tw > max_crops:
        if tw > th:
            tw = math.floor(max_crops / th)
        else:
            th = math.floor(max_crops / tw)
    return max(1, th), max(1, tw)


def overlap_crops(image: np.ndarray, overlap_margin: int, max_crops: int, base: int = 378,
                  patch: int = 14) -> Tuple[np.ndarray, Tuple[int, int]]:
    """image_crops.py:58-167, PIL-Lanczos branch (:138-150; pyvips is absent in this image).
    Returns (uint8 [1 + th*tw, base, base, C], (th, tw))."""
    h0, w0 = image.shape[:2]
    margin_px = patch * overlap_margin
    window = (base // patch - 2 * overlap_margin) * patch
    th, tw = select_tiling(h0 - 2 * margin_px, w0 - 2 * margin_px, window, max_crops)
    out = np.zeros((1 + th * tw, base, base, image.shape[2]), dtype=np.uint8)
    pil = Image.fromarray(image)
    tgt_h, tgt_w = th * window + 2 * margin_px, tw * window + 2 * margin_px
    resized = np.asarray(pil.resize((int(tgt_w), int(tgt_h)), resample=Image.Resampling.LANCZOS))
    out[0] = np.asarray(pil.resize((base, base), resample=Image.Resampling.LANCZOS))
    for ty in range(th):
        for tx in range(tw):
            y0, x0 = ty * window, tx * window
            piece = resized[y0:min(y0 + base, resized.shape[0]), x0:min(x0 + base, resized.shape[1])]
            out[1 + ty * tw + tx, :piece.shape[0], :piece.shape[1]] = piece
    return out, (th, tw)


def stitch_crops(local: Tensor, tiling: Tuple[int, int], margin: int) -> Tensor:
    """reconstruct_from_crops with patch_size=1 (image_crops.py:170-231, called at
    moondream.py:221-226): local [th*tw, g, g, D] -> [(g-2m)*th+2m, (g-2m)*tw+2m, D]."""
    th, tw = tiling
    g = local.shape[1]
    inner = g - 2 * margin
    canvas = torch.zeros((inner * th + 2 * margin, inner * tw + 2 * margin, local.shape[3]),
                         dtype=local.dtype, device=local.device)
    for idx in range(local.shape[0]):
        ty, tx = divmod(idx, tw)
        ys, ye = (0 if ty == 0 else margin), (g if ty == th - 1 else g - margin)
        xs, xe = (0 if tx == 0 else margin), (g if tx == tw - 1 else g - margin)
        canvas[ty * inner + ys: ty * inner + ye, tx * inner + xs: tx * inner + xe] = \
            local[idx, ys:ye, xs:xe]
    return canvas


# --------------------------------------------------------------------------------------------
# primitives (moondream/torch/layers.py, rope.py)
# --------------------------------------------------------------------------------------------
def dequantize_tensor(W_q: Tensor, scale: Tensor, zero: Tensor, orig_shape, dtype=torch.bfloat16) -> Tensor:
    """layers.py:38-44 (`dequantize_tensor`, the unpacking of an int4 group-128 `QuantizedLinear`): rows [0, step) of
    the result are the high nibbles of `W_q`, rows [step, 2 step) the low nibbles (one row per group of 128 input
    features); `W_r.sub_(zero).mul_(scale)` runs on a `dtype` tensor with fp32 [groups, 1] parameters, so each step is
    evaluated in fp32 and rounded to `dtype`."""
    step = W_q.shape[0]
    W_r = torch.empty([2 * step, W_q.shape[1]], dtype=dtype)
    W_r[:step] = ((W_q & 0b11110000) >> 4).to(dtype)
    W_r[step:] = (W_q & 0b00001111).to(dtype)
    W_r = (W_r.to(torch.float32) - zero.to(torch.float32)).to(dtype)
    W_r = (W_r.to(torch.float32) * scale.to(torch.float32)).to(dtype)
    return W_r.reshape(orig_shape)


def dequantized_state_dict(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """A state dict holding reference-format int4 entries (`X.weight.packed / .scale / .zero_point`, layers.py:58-76)
    -> the bf16 state dict `QuantizedLinear.unpack` builds from it (layers.py:79-99; the torchao re-quantisation that
    follows, :102, needs a library this image does not have: SURVEY.md section 8c defines parity on these weights)."""
    out = {k: v for k, v in sd.items() if ".weight." not in k}
    for k, v in sd.items():
        if k.endswith(".weight.packed"):
            base = k[: -len(".packed")]
            out_f = sd[base[: -len("weight")] + "bias"].numel()
            in_f = v.numel() * 2 // out_f
            out[base] = dequantize_tensor(v, sd[base + ".scale"], sd[base + ".zero_point"], (out_f, in_f))
    return out


def _lin(x: Tensor, w: Dict[str, Tensor], prefix: str) -> Tensor:
    """layers.py:34-35 / nn.Linear."""
    return F.linear(x, w[prefix + ".weight"], w[prefix + ".bias"])


def _ln(x: Tensor, w: Dict[str, Tensor], prefix: str) -> Tensor:
    """layers.py:118-119 (eps 1e-5)."""
    b = w[prefix + ".bias"]
    return F.layer_norm(x, b.shape, w[prefix + ".weight"], b)


def _adapter(x: Tensor, ab: Dict[str, Tensor]) -> Tensor:
    """the LoRA side path F.linear(F.linear(x, A), B) (text.py:31-32,54-56, layers.py:133,141)"""
    return F.linear(F.linear(x, ab["A"]), ab["B"])


def _mlp(x: Tensor, w: Dict[str, Tensor], prefix: str, lora: Optional[dict] = None) -> Tensor:
    """layers.py:129-146: fc2(gelu_tanh(fc1 x)), each Linear optionally followed by `+ B(A x)` of a LoRA variant."""
    h = _lin(x, w, prefix + ".fc1")
    if lora is not None:
        h = h + _adapter(x, lora["fc1"])
    h = F.gelu(h, approximate="tanh")
    y = _lin(h, w, prefix + ".fc2")
    if lora is not None:
        y = y + _adapter(h, lora["fc2"])
    return y


def rope_table(head_dim: int, max_context: int, theta: float = 10000.0) -> Tensor:
    """rope.py:6-17 as called from text.py:215-219: dim = head_dim // 2 -> [ctx, head_dim/4, 2] f32."""
    dim = head_dim // 2
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    angles = torch.arange(max_context, dtype=torch.float32).unsqueeze(1) * freqs.unsqueeze(0)
    unit = torch.exp(1j * angles)
    return torch.stack([unit.real, unit.imag], dim=-1)


def apply_rope(x: Tensor, table: Tensor, pos: Tensor, rot_dim: int = 32) -> Tensor:
    """rope.py:20-48, interleave=False: split-half input, interleaved output, fp32 math by
    promotion against the fp32 table, cast back to x.dtype.  x: [1, H, T, hd]."""
    rot, keep = x[..., :rot_dim], x[..., rot_dim:]
    half = rot_dim // 2
    re, im = rot[..., :half], rot[..., half:]
    cos = table[..., 0][pos, :].unsqueeze(0).unsqueeze(0)
    sin = table[..., 1][pos, :].unsqueeze(0).unsqueeze(0)
    out_re = re * cos - im * sin
    out_im = re * sin + im * cos
    mixed = torch.stack((out_re, out_im), dim=-1).flatten(-2)
    return torch.cat([mixed.to(x.dtype), keep], dim=-1)


# --------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------
@dataclass
class Encoded:
    """moondream.py:56-59 EncodedImage: pos + per-layer (k, v) [1, KVH, pos, hd]."""
    pos: int
    caches: List[Tuple[Tensor, Tensor]]


@dataclass
class Generation:
    tokens: List[int]            # emitted token ids (first one comes from the prompt prefill)
    margins: List[float]         # top1 - top2 logit margin at each emission (after masking)
    predicted: List[int]         # argmax at each step (== tokens unless teacher-forced)
    last_hidden: Optional[Tensor] = None
    margin_ulps: Optional[List[float]] = None   # the margins in bf16 ulps of the top logit


class OracleModel:
    def __init__(self, cfg, weights: Dict[str, Tensor], dtype: torch.dtype = torch.bfloat16, device="cpu"):
        """device="cuda" runs the same eager torch ops on a GPU: the reference's torch-CUDA path used as the
        speed comparator (tools/torch_cuda_comparator.py); parity work always uses the CPU instance."""
        self.cfg = cfg
        self.dtype = dtype
        self.device = torch.device(device)
        self.w = {k: v.to(device=self.device, dtype=dtype) for k, v in weights.items()}
        t = cfg.text
        self.head_dim = t.dim // t.n_heads
        self.rope = rope_table(self.head_dim, t.max_context).to(self.device)
        # moondream.py:138-146: causal mask with a bidirectional [prefix x prefix] block
        mask = torch.tril(torch.ones(1, 1, t.max_context, t.max_context, dtype=torch.bool))
        prefix = 1 + (cfg.vision.crop_size // cfg.vision.enc_patch_size) ** 2
        mask[..., :prefix, :prefix] = 1
        self.attn_mask = mask.to(self.device)
        self.lora: Optional[dict] = None       # a LoRA variant tree (lora.py:55-79 `variant_state_dict`), applied by text_decoder
        self.reset_cache()

    # ---- KV cache (moondream.py:62-78) ----
    def reset_cache(self):
        t = self.cfg.text
        shape = (1, t.n_kv_heads, t.max_context, self.head_dim)
        self.k_cache = [torch.zeros(shape, dtype=self.dtype, device=self.device) for _ in range(t.n_layers)]
        self.v_cache = [torch.zeros(shape, dtype=self.dtype, device=self.device) for _ in range(t.n_layers)]

    def load_encoded(self, enc: Encoded):
        """moondream.py:620-623."""
        for i, (k, v) in enumerate(enc.caches):
            self.k_cache[i][:, :, : k.size(2), :] = k
            self.v_cache[i][:, :, : v.size(2), :] = v

    # ---- vision (moondream/torch/vision.py) ----
    def prepare_crops(self, image: np.ndarray) -> Tuple[Tensor, Tuple[int, int]]:
        """vision.py:25-41: uint8 NHWC -> NCHW in `dtype`, in-place /255, -0.5, /0.5."""
        v = self.cfg.vision
        crops, tiling = overlap_crops(image, v.overlap_margin, v.max_crops, v.crop_size, v.enc_patch_size)
        x = torch.from_numpy(np.transpose(crops, (0, 3, 1, 2))).to(device=self.device, dtype=self.dtype)
        x = x.div_(255.0).sub_(0.5).div_(0.5)
        return x, tiling

    def vision_encoder(self, crops: Tensor) -> Tensor:
        """vision.py:44-74: patchify (c, py, px feature order) -> patch_emb + pos_emb ->
        27 x [x += attn(ln1 x); x += mlp(ln2 x)] -> post_ln.  crops [B,3,S,S] -> [B, T, D]."""
        v, w = self.cfg.vision, self.w
        B, Cc, Hh, Ww = crops.shape
        p = v.enc_patch_size
        x = crops.reshape(B, Cc, Hh // p, p, Ww // p, p).permute(0, 2, 4, 1, 3, 5)
        x = x.reshape(B, (Hh // p) * (Ww // p), Cc * p * p)
        x = _lin(x, w, "vision.patch_emb")
        x = x + w["vision.pos_emb"]
        nh = v.enc_n_heads
        for i in range(v.enc_n_layers):
            pre = f"vision.blocks.{i}"
            h = _ln(x, w, pre + ".ln1")
            # layers.py:155-166
            bsz, T, D = h.shape
            q, k, val = [t.view(bsz, T, nh, D // nh).transpose(1, 2)
                         for t in _lin(h, w, pre + ".attn.qkv").chunk(3, dim=-1)]
            a = F.scaled_dot_product_attention(q, k, val).transpose(1, 2).reshape(bsz, T, D)
            x = x + _lin(a, w, pre + ".attn.proj")
            x = x + _mlp(_ln(x, w, pre + ".ln2"), w, pre + ".mlp")
        return _ln(x, w, "vision.post_ln")

    def vision_projection(self, global_feat: Tensor, stitched: Tensor) -> Tensor:
        """vision.py:77-89: adaptive-avg-pool the stitched map to grid x grid, concat with the global
        crop's features, 2-layer GELU MLP."""
        v = self.cfg.vision
        g = v.crop_size // v.enc_patch_size
        pooled = F.adaptive_avg_pool2d(stitched.permute(2, 0, 1), output_size=(g, g))
        pooled = pooled.permute(1, 2, 0).reshape(g * g, v.enc_dim)
        return _mlp(torch.cat([global_feat, pooled], dim=-1), self.w, "vision.proj_mlp")

    def run_vision(self, image: np.ndarray) -> Tensor:
        """moondream.py:206-228."""
        v = self.cfg.vision
        crops, tiling = self.prepare_crops(image)
        feats = self.vision_encoder(crops)
        g = v.crop_size // v.enc_patch_size
        stitched = stitch_crops(feats[1:].view(-1, g, g, v.enc_dim), tiling, v.overlap_margin)
        return self.vision_projection(feats[0], stitched)

    # ---- text (moondream/torch/text.py) ----
    def embed(self, ids: Tensor) -> Tensor:
        """text.py:12-13."""
        return F.embedding(ids, self.w["text.wte"])

    def text_decoder(self, x: Tensor, mask: Tensor, pos_ids: Tensor) -> Tensor:
        """text.py:128-160 (+ attn :16-60): parallel-residual blocks with ONE LayerNorm; the KV
        cache is scattered at pos_ids and attention runs over all max_context slots under `mask`."""
        t, w = self.cfg.text, self.w
        nh, nkv, hd = t.n_heads, t.n_kv_heads, self.head_dim
        for i in range(t.n_layers):
            pre = f"text.blocks.{i}"
            h = _ln(x, w, pre + ".ln")
            bsz, T, D = h.shape
            lora = self.lora["text"]["blocks"][str(i)] if self.lora is not None else None
            qkv = _lin(h, w, pre + ".attn.qkv")
            if lora is not None:
                qkv = qkv + _adapter(h, lora["attn"]["qkv"])                       # text.py:31-32 (in place there)
            q, k, v = qkv.split([nh * hd, nkv * hd, nkv * hd], dim=-1)
            q = q.view(bsz, T, nh, hd).transpose(1, 2)
            k = k.view(bsz, T, nkv, hd).transpose(1, 2)
            v = v.view(bsz, T, nkv, hd).transpose(1, 2)
            q = apply_rope(q, self.rope, pos_ids)
            k = apply_rope(k, self.rope, pos_ids)
            self.k_cache[i][:, :, pos_ids, :] = k
            self.v_cache[i][:, :, pos_ids, :] = v
            a = F.scaled_dot_product_attention(q, self.k_cache[i], self.v_cache[i], attn_mask=mask,
                                               enable_gqa=nh != nkv)
            a = a.transpose(1, 2).reshape(bsz, T, D)
            l_attn = _lin(a, w, pre + ".attn.proj")
            if lora is not None:
                l_attn = l_attn + _adapter(h, lora["attn"]["proj"])                # text.py:53-56: fed the block INPUT
            x = x + l_attn + _mlp(h, w, pre + ".mlp", lora["mlp"] if lora is not None else None)
        return x

    def lm_head(self, hidden: Tensor) -> Tensor:
        """text.py:163-167: last token -> post_ln -> Linear."""
        return _lin(_ln(hidden[:, -1, :], self.w, "text.post_ln"), self.w, "text.lm_head")

    # ---- region head (moondream/torch/region.py) ----
    def _fourier(self, x: Tensor, key: str) -> Tensor:
        """region.py:12-29."""
        f = 2 * math.pi * x @ self.w[key]
        return torch.cat([f.cos(), f.sin()], dim=-1)

    def encode_coordinate(self, coord: Tensor) -> Tensor:
        """region.py:32-43."""
        return _lin(self._fourier(coord, "region.coord_features"), self.w, "region.coord_encoder")

    def decode_coordinate(self, hidden: Tensor) -> Tensor:
        """region.py:46-57."""
        return _mlp(hidden, self.w, "region.coord_decoder")

    def encode_size(self, size: Tensor) -> Tensor:
        """region.py:60-71."""
        return _lin(self._fourier(size, "region.size_features"), self.w, "region.size_encoder")

    def decode_size(self, hidden: Tensor) -> Tensor:
        """region.py:74-93."""
        return _mlp(hidden, self.w, "region.size_decoder").view(2, -1)

    # ---- API-level flows (moondream/torch/moondream.py) ----
    def encode_image(self, image: np.ndarray, return_embeds: bool = False):
        """moondream.py:230-268: vision -> [BOS; image] prefill at positions 0..729 -> KV snapshot."""
        with torch.no_grad():
            img_emb = self.run_vision(image)
            bos = self.embed(torch.tensor([[self.cfg.tokenizer.bos_id]], device=self.device))
            x = torch.cat([bos, img_emb[None]], dim=1)
            n = x.size(1)
            hidden = self.text_decoder(x, self.attn_mask[:, :, 0:n, :], torch.arange(n, dtype=torch.long, device=self.device))
            enc = Encoded(n, [(self.k_cache[i][:, :, :n, :].clone(), self.v_cache[i][:, :, :n, :].clone())
                              for i in range(self.cfg.text.n_layers)])
        if return_embeds:
            return enc, img_emb, hidden
        return enc

    def spatial_prompt_embeds(self, prompt: Sequence[int], spatial_refs) -> Tensor:
        """Prompt embedding with the coord / size placeholder tokens replaced by region encodings
        (moondream.py:293-301 over region.py:96-136 `encode_spatial_refs`): a point contributes (x, y), a box its
        centre (x_c, y_c) and its (width, height); coordinates are encoded one scalar at a time, sizes as pairs."""
        tk = self.cfg.tokenizer
        coords, sizes = [], []
        for ref in spatial_refs:
            if len(ref) == 2:
                coords += [ref[0], ref[1]]
            else:
                coords += [(ref[0] + ref[2]) / 2, (ref[1] + ref[3]) / 2]
                sizes.append([ref[2] - ref[0], ref[3] - ref[1]])
        ids = torch.tensor([list(prompt)], device=self.device)
        with torch.no_grad():
            x = self.embed(ids)
            c = torch.tensor(coords, device=self.device, dtype=self.dtype).view(-1, 1)
            x[ids == tk.coord_id] = self.encode_coordinate(c)
            if sizes:
                x[ids == tk.size_id] = self.encode_size(torch.tensor(sizes, device=self.device, dtype=self.dtype))
        return x

    def causal_mask(self) -> Tensor:
        """The mask query() builds for a text-only question (moondream.py:571-574): plain lower-triangular."""
        if getattr(self, "_causal", None) is None:
            n = self.cfg.text.max_context
            self._causal = torch.tril(torch.ones(1, 1, n, n, dtype=torch.bool)).to(self.device)
        return self._causal

    def prefill_prompt(self, prompt: Sequence[int], pos: int, embeds: Optional[Tensor] = None, causal: bool = False):
        """moondream.py:280-321 with temperature == 0: returns (logits, hidden, next_token, pos).
        `causal`: use the text-only query's mask instead of the prefix-LM one (:305-308)."""
        with torch.no_grad():
            x = self.embed(torch.tensor([list(prompt)], device=self.device)) if embeds is None else embeds
            T = x.size(1)
            mask = self.causal_mask() if causal else self.attn_mask
            hidden = self.text_decoder(x, mask[:, :, pos:pos + T, :],
                                       torch.arange(pos, pos + T, dtype=torch.long, device=self.device))
            logits = self.lm_head(hidden)
            nxt = torch.argmax(logits, dim=-1).unsqueeze(1)
        return logits, hidden, nxt, pos + T

    def decode_one(self, emb: Tensor, pos: int):
        """moondream.py:183-192 + the mask/pos bookkeeping of the generator (:472-474, :514)."""
        mask = torch.zeros(1, 1, self.cfg.text.max_context, dtype=torch.bool, device=self.device)
        mask[:, :, : pos + 1] = 1
        with torch.no_grad():
            hidden = self.text_decoder(emb, mask, torch.tensor([pos], dtype=torch.long, device=self.device))
            logits = self.lm_head(hidden)
        return logits, hidden

    @staticmethod
    def _margin(logits: Tensor) -> float:
        top = torch.topk(logits.float().flatten(), 2).values
        return float(top[0] - top[1])

    @staticmethod
    def _margin_ulps(logits: Tensor) -> float:
        """top1 - top2 in units of the bf16 spacing at |top1| (8 significand bits)."""
        top = torch.topk(logits.float().flatten(), 2).values
        mag = max(abs(float(top[0])), 1e-30)
        ulp = 2.0 ** (math.floor(math.log2(mag)) - 7)
        return float(top[0] - top[1]) / ulp

    @staticmethod
    def next_token(logits: Tensor, temperature: float, top_p: float) -> int:
        """moondream.py:312-318 / :524-530: argmax at temperature 0, else softmax(logits / T) -> `_apply_top_p`
        (:270-278: sort descending, drop tokens whose preceding mass exceeds top_p, renormalise, scatter back)
        -> torch.multinomial on the global RNG, all in the logits' dtype."""
        if temperature == 0:
            return int(torch.argmax(logits, dim=-1).item())
        p = torch.softmax(logits / temperature, dim=-1)
        srt, idx = torch.sort(p, dim=-1, descending=True)
        csum = torch.cumsum(srt, dim=-1)
        srt[csum - srt > top_p] = 0.0
        srt.div_(srt.sum(dim=-1, keepdim=True))
        kept = torch.zeros_like(p)
        kept.scatter_(dim=-1, index=idx, src=srt)
        return int(torch.multinomial(kept, num_samples=1).item())

    def generate(self, enc: Optional[Encoded], prompt: Sequence[int], max_tokens: int,
                 forced: Optional[Sequence[int]] = None, temperature: float = 0.0,
                 top_p: float = 0.3, spatial_refs=None, pos: Optional[int] = None) -> Generation:
        """``_generate_answer`` (moondream.py:434-539) reduced to token ids (greedy unless temperature > 0,
        in which case `predicted` holds the sampled tokens and the margins still describe the argmax): prompt prefill,
        then one decoder step per emitted token; ``answer_id`` is masked from the 2nd token on
        (:517), stop on eos (:481-483).  Note the reference runs the decoder once more after the
        last emitted token (its result is discarded); that trailing step is reproduced.
        `forced`: teacher forcing — feed these tokens instead of the argmax (records the argmax)."""
        tk = self.cfg.tokenizer
        causal = False
        if enc is None and pos is None:        # text-only query (moondream.py:565-574): fresh caches, position 0, causal mask
            self.reset_cache()
            start, causal = 0, True
        elif enc is None:                      # continue in the current cache (the answer phase after reasoning)
            start = pos
            causal = getattr(self, "_text_only", False)
        else:
            self.load_encoded(enc)
            start = enc.pos
        self._text_only = causal
        embeds = self.spatial_prompt_embeds(prompt, spatial_refs) if spatial_refs else None
        logits, hidden, nxt, pos = self.prefill_prompt(prompt, start, embeds, causal=causal)
        out = Generation([], [], [], margin_ulps=[])
        n = 0
        pred = int(nxt.item()) if temperature == 0 else self.next_token(logits, temperature, top_p)
        margin = self._margin(logits)
        ulps = self._margin_ulps(logits)
        while True:
            if n >= max_tokens:
                break
            tok = pred if forced is None else int(forced[n])
            if forced is None and tok == tk.eos_id:
                break
            out.tokens.append(tok)
            out.predicted.append(pred)
            out.margins.append(margin)
            out.margin_ulps.append(ulps)
            logits, hidden = self.decode_one(self.embed(torch.tensor([[tok]], device=self.device)), pos)
            logits[:, tk.answer_id] = float("-inf")
            pos += 1
            pred = self.next_token(logits, temperature, top_p)
            margin = self._margin(logits)
            ulps = self._margin_ulps(logits)
            n += 1
        out.last_hidden = hidden
        return out

    def generate_reasoning(self, enc: Optional[Encoded], prompt: Sequence[int], max_tokens: int,
                           temperature: float = 0.0, top_p: float = 0.3, spatial_refs=None) -> dict:
        """``_generate_reasoning`` (moondream.py:323-432) reduced to ids: the chain of thought ends at answer_id;
        eos_id and size_id are masked from the 2nd token on (:395-396); a coord_id token is fed to the decoder as
        encode_coordinate(argmax(decode_coordinate(last hidden)) / bins) (:381-391) and its value recorded.
        Returns {"pos", "tokens", "coords" (one per token, None unless coord_id), "margin_ulps" (token decisions),
        "coord_ulps" (coordinate decisions)}.  enc None = text-only (fresh caches, causal mask)."""
        tk = self.cfg.tokenizer
        causal = enc is None
        if causal:
            self.reset_cache()
            start = 0
        else:
            self.load_encoded(enc)
            start = enc.pos
        self._text_only = causal
        embeds = self.spatial_prompt_embeds(prompt, spatial_refs) if spatial_refs else None
        logits, hidden, _, pos = self.prefill_prompt(prompt, start, embeds, causal=causal)
        hidden = hidden[:, -1:, :]
        nxt = self.next_token(logits, temperature, top_p)
        ulps = self._margin_ulps(logits)
        out = {"tokens": [], "coords": [], "margin_ulps": [], "coord_ulps": []}
        n = 0
        with torch.no_grad():
            while nxt != tk.answer_id and n < max_tokens:
                out["tokens"].append(nxt)
                out["margin_ulps"].append(ulps)
                if nxt == tk.coord_id:
                    cl = self.decode_coordinate(hidden)
                    coord = torch.argmax(cl, dim=-1) / cl.size(-1)
                    out["coords"].append(float(coord.item()))
                    out["coord_ulps"].append(self._margin_ulps(cl))
                    emb = self.encode_coordinate(coord.to(dtype=cl.dtype)).unsqueeze(0)
                else:
                    out["coords"].append(None)
                    out["coord_ulps"].append(None)
                    emb = self.embed(torch.tensor([[nxt]], device=self.device))
                logits, hidden = self.decode_one(emb, pos)
                logits[:, tk.eos_id] = float("-inf")
                logits[:, tk.size_id] = float("-inf")
                pos += 1
                nxt = self.next_token(logits, temperature, top_p)
                ulps = self._margin_ulps(logits)
                n += 1
        out["pos"] = pos
        out["end_margin_ulps"] = ulps          # the decision that ended the chain (or would have continued it)
        return out

    def generate_points(self, enc: Encoded, prompt: Sequence[int], include_size: bool,
                        max_objects: int) -> List[dict]:
        """detect / point (moondream.py:735-829) -> _generate_points (:653-733)."""
        tk = self.cfg.tokenizer
        self.load_encoded(enc)
        _, hidden, nxt, pos = self.prefill_prompt(prompt, enc.pos)
        hidden = hidden[:, -1:, :]
        objs: List[dict] = []
        with torch.no_grad():
            while int(nxt.item()) != tk.eos_id and len(objs) < max_objects:
                xl = self.decode_coordinate(hidden)
                xc = torch.argmax(xl, dim=-1) / xl.size(-1)
                emb = self.encode_coordinate(xc.to(dtype=xl.dtype)).unsqueeze(0)
                _, hidden = self.decode_one(emb, pos)
                pos += 1
                yl = self.decode_coordinate(hidden)
                yc = torch.argmax(yl, dim=-1) / yl.size(-1)
                emb = self.encode_coordinate(yc.to(dtype=yl.dtype)).unsqueeze(0)
                if include_size:
                    _, hidden = self.decode_one(emb, pos)
                    pos += 1
                    sl = self.decode_size(hidden)
                    wb, hb = torch.argmax(sl[0], dim=-1), torch.argmax(sl[1], dim=-1)
                    wv = torch.pow(2.0, (wb.float() / 1023.0) * 10.0 - 10.0)
                    hv = torch.pow(2.0, (hb.float() / 1023.0) * 10.0 - 10.0)
                    emb = self.encode_size(torch.tensor([wv, hv], dtype=sl.dtype)).unsqueeze(0).unsqueeze(0)
                    objs.append({"x_min": xc.item() - wv.item() / 2, "y_min": yc.item() - hv.item() / 2,
                                 "x_max": xc.item() + wv.item() / 2, "y_max": yc.item() + hv.item() / 2,
                                 "bins": [int(torch.argmax(xl)), int(torch.argmax(yl)), int(wb), int(hb)],
                                 "ulps": [self._margin_ulps(xl), self._margin_ulps(yl),
                                          self._margin_ulps(sl[0]), self._margin_ulps(sl[1])]})
                else:
                    objs.append({"x": xc.item(), "y": yc.item(),
                                 "bins": [int(torch.argmax(xl)), int(torch.argmax(yl))],
                                 "ulps": [self._margin_ulps(xl), self._margin_ulps(yl)]})
                logits, hidden = self.decode_one(emb, pos)
                pos += 1
                nxt = torch.argmax(logits, dim=-1)
                objs[-1]["ulps"].append(self._margin_ulps(logits))     # the continue/stop decision
        return objs

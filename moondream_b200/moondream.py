"""Drop-in for ``moondream.torch.moondream.MoondreamModel`` (reference moondream/torch/moondream.py)
on the path BASELINE.json names: ``encode_image`` / ``caption`` / ``query`` / ``detect`` / ``point``
with the reference's signatures, settings keys, result dict shapes and exceptions — plus the batched
calls the reference cannot offer (it is strictly batch-1: moondream.py:66,369,482,669):
``encode_images`` / ``caption_batch`` / ``query_batch`` / ``detect_batch`` / ``point_batch``.

All arithmetic runs in libmoondream_b200.so; this module keeps what the reference also keeps in
Python: crop geometry (PIL), tokenizer, prompt templates, result shaping, streaming detokenisation.

Deliberate differences from the reference (documented in DESIGN.md):
  * ``settings`` without "variant" is accepted by ``encode_image`` (the reference raises KeyError at
    moondream.py:240-243); LoRA variants (lora.py) are read from the reference's local cache layout or a given
    file — a variant that is not cached raises instead of being downloaded (this build never uses the network).
  * temperature > 0 (the reference's default 0.5 / top_p 0.3) follows the reference's softmax / top-p arithmetic
    (moondream.py:270-278, 312-318) ON THE DEVICE inside the decode graph (csrc/sampling.cu); the draw is an inverse
    CDF over the kept probabilities from a Philox stream seeded from torch's global RNG (the reference's
    torch.multinomial consumes that RNG differently, so tokens agree in distribution, not draw by draw).
    `settings["host_sampler"] = True` selects the host restatement with torch CPU ops (parity tests).
  * ``EncodedImage`` holds KV *pages* (shared, copy-on-write for the partial page) instead of
    cloned tensors; ``.caches`` materialises the reference's per-layer (k, v) view on demand.
"""
from __future__ import annotations

from typing import Any, Dict, List, Literal, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .config import MoondreamConfig
from .engine import Engine, PrefixKV

try:
    from PIL import Image
except Exception:  # pragma: no cover
    Image = None

DEFAULT_MAX_TOKENS = 768
DEFAULT_TEMPERATURE = 0.5
DEFAULT_TOP_P = 0.3
DEFAULT_MAX_OBJECTS = 50

SpatialRefs = List[Union[Tuple[float, float], Tuple[float, float, float, float]]]


class EncodedImage:
    """Counterpart of moondream.py:56-59 (``pos`` + per-layer KV).  Immutable; owns its pages."""

    def __init__(self, prefix: PrefixKV, engine: Engine):
        self._prefix = prefix
        self._engine = engine

    @property
    def pos(self) -> int:
        return self._prefix.pos

    @property
    def caches(self) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        return self._engine.prefix_kv_tensors(self._prefix)


def _as_array(image) -> np.ndarray:
    if Image is not None and isinstance(image, Image.Image):
        return np.array(image.convert("RGB"))
    if isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 3 and image.dtype == np.uint8:
        return image
    raise ValueError("image must be a PIL Image or EncodedImage")


class MoondreamModel:
    def __init__(self, config: MoondreamConfig, dtype=torch.bfloat16, setup_caches: bool = True,
                 tokenizer=None, device: str = "cuda", max_batch: int = 32, kv_pages: Optional[int] = None,
                 quantize: Optional[str] = None):
        """quantize: None, "int4" or "int8" — quantise the decoder blocks of a bf16 checkpoint at load time.  A
        checkpoint already in the reference's int4 QuantizedLinear format (layers.py:47-110) is detected by its keys."""
        if dtype != torch.bfloat16:
            raise ValueError("the path is bf16 end to end, like the reference (vision.py:36, weights.py:32)")
        self.config = config
        self._device = torch.device(device)
        self._tokenizer = tokenizer
        self._max_batch = max_batch
        self._kv_pages = kv_pages
        self._quantize = quantize
        self._variants: Dict[str, Any] = {}
        self._engine: Optional[Engine] = None

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self):
        return self._device

    @property
    def tokenizer(self):
        if self._tokenizer is None:
            from tokenizers import Tokenizer  # the reference's tokenizer (moondream.py:89)

            self._tokenizer = Tokenizer.from_pretrained("moondream/starmie-v1")
        return self._tokenizer

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            raise RuntimeError("weights are not loaded: call load_state_dict() or "
                               "moondream_b200.weights.load_weights_into_model(path, model)")
        return self._engine

    # the reference's four swap seams (moondream.py:168-192), batch-1 tensor signatures, on the native engine
    @property
    def seam(self):
        if getattr(self, "_seam", None) is None:
            from .seam import SeamAdapter

            self._seam = SeamAdapter(self.engine)
        return self._seam

    def _vis_enc(self, x: torch.Tensor):
        return self.seam._vis_enc(x)

    def _vis_proj(self, g: torch.Tensor, r: torch.Tensor):
        return self.seam._vis_proj(g, r)

    def _prefill(self, x: torch.Tensor, attn_mask: torch.Tensor, pos_ids: torch.Tensor, lora=None):
        return self.seam._prefill(x, attn_mask, pos_ids, lora)

    def _decode_one_tok(self, x: torch.Tensor, attn_mask: torch.Tensor, pos_ids: torch.Tensor, lora=None):
        return self.seam._decode_one_tok(x, attn_mask, pos_ids, lora)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = False):
        sd = {k.replace("model.", "", 1) if k.startswith("model.") else k: v for k, v in state_dict.items()}
        self._engine = Engine(self.config, sd, device=self._device, kv_pages=self._kv_pages,
                              max_batch=self._max_batch, quantize=self._quantize)
        self._seam = None
        return [], []

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def compile(self):
        """The reference swaps in torch.compile'd seam methods here (moondream.py:194-204); this
        engine is already native, the decode step is CUDA-graph captured on first use."""
        return None

    def _setup_caches(self):
        return None

    def load_encoded_image(self, encoded_image: "EncodedImage"):
        """No-op: sequences reference the prefix pages directly (moondream.py:620-623 copies 48 tensors)."""
        return None

    # ------------------------------------------------------------------ settings
    @staticmethod
    def _text_settings(settings: Optional[dict]):
        """(max_tokens, sampling) from the reference's TextSamplingSettings (moondream.py:443-454).  `sampling` is the
        keyword set Engine.generate takes: empty for temperature 0 (greedy), temperature / top_p for on-device
        sampling, or a HostSampler when settings["host_sampler"] asks for the torch-CPU restatement."""
        temperature = settings.get("temperature", DEFAULT_TEMPERATURE) if settings else DEFAULT_TEMPERATURE
        top_p = settings.get("top_p", DEFAULT_TOP_P) if settings else DEFAULT_TOP_P
        max_tokens = settings.get("max_tokens", DEFAULT_MAX_TOKENS) if settings else DEFAULT_MAX_TOKENS
        if temperature < 0:
            raise ValueError("temperature must be >= 0")
        sampling: Dict[str, Any] = {}
        if temperature != 0:
            if settings and settings.get("host_sampler"):
                from .sampling import HostSampler
                sampling["sampler"] = HostSampler(temperature, top_p)
            else:
                sampling.update(temperature=float(temperature), top_p=float(top_p))
                if settings and settings.get("seed") is not None:
                    sampling["seed"] = int(settings["seed"])
        return max_tokens, sampling

    # ------------------------------------------------------------------ LoRA variants
    def _lora(self, settings: Optional[dict]):
        """settings["variant"] -> LoraVariant on the device (lora.py:55-79 `variant_state_dict`): a variant id is looked
        up in the reference's cache layout ($HF_HUB_CACHE | $HF_HOME/hub | ~/.cache/huggingface/hub)/md_variants/<id>/
        final.pt (lora.py:11-29); a path to such a file, a state dict or a loaded LoraVariant are accepted too.  The
        reference downloads a missing variant from api.moondream.ai (lora.py:31-40); this build never touches the
        network and raises instead."""
        import os

        from .engine import LoraVariant

        v = settings.get("variant") if settings else None
        if v is None:
            return None
        if isinstance(v, LoraVariant):
            return v
        if isinstance(v, dict):
            return self.engine.load_lora(v)
        key = str(v)
        if key not in self._variants:
            path = key
            if not os.path.isfile(path):
                hub = os.environ.get("HF_HUB_CACHE")
                if hub is None:
                    home = os.environ.get("HF_HOME")
                    hub = os.path.join(home, "hub") if home is not None else os.path.expanduser("~/.cache/huggingface/hub")
                path = os.path.join(hub, "md_variants", key, "final.pt")
            if not os.path.isfile(path):
                raise RuntimeError(f"variant {key!r} is not in the local cache ({path}); the reference would download it "
                                   f"(lora.py:31-40), this build is offline")
            flat = torch.load(path, map_location="cpu", weights_only=True)
            renamed = {}
            for k, t in flat.items():                      # lora.py:64-76
                for old, new in (("text_model.transformer.h", "text.blocks"), (".mixer", ".attn"), (".out_proj", ".proj"),
                                 (".Wqkv", ".qkv"), (".parametrizations.weight.0", "")):
                    if old in k:
                        k = k.replace(old, new)
                renamed[k] = t
            self._variants[key] = self.engine.load_lora(renamed)
        return self._variants[key]

    # ------------------------------------------------------------------ image encoding
    def encode_images(self, images: Sequence[Any], settings: Optional[dict] = None) -> List[EncodedImage]:
        lora = self._lora(settings)
        todo = [(i, _as_array(im)) for i, im in enumerate(images) if not isinstance(im, EncodedImage)]
        out: List[Optional[EncodedImage]] = [im if isinstance(im, EncodedImage) else None for im in images]
        for lo in range(0, len(todo), self._max_batch):
            chunk = todo[lo: lo + self._max_batch]
            prefixes = self.engine.encode_images([a for _, a in chunk], lora=lora)
            for (i, _), p in zip(chunk, prefixes):
                out[i] = EncodedImage(p, self.engine)
        return out  # type: ignore[return-value]

    def encode_image(self, image, settings: Optional[dict] = None) -> EncodedImage:
        """moondream.py:230-268; under settings["variant"] the image prefill runs with the adapters (:240-257)."""
        if isinstance(image, EncodedImage):
            return image
        return self.encode_images([image], settings)[0]

    # ------------------------------------------------------------------ text generation
    def _run(self, encoded: Sequence[EncodedImage], prompts: Sequence[Sequence[int]], max_tokens: int,
             eos_id: Optional[int] = None, prompt_embeds=None, sampling: Optional[dict] = None,
             prefix_len: int = -1, lora=None) -> List[List[int]]:
        eos = self.config.tokenizer.eos_id if eos_id is None else eos_id
        res = self.engine.generate([e._prefix for e in encoded], prompts, max_tokens,
                                   prompt_embeds=prompt_embeds, prefix_len=prefix_len, lora=lora, **(sampling or {}))
        toks = res.tokens.tolist()
        out = []
        for row in toks:
            seq = []
            for t in row[:max_tokens]:
                if t == eos:
                    break
                seq.append(t)
            out.append(seq)
        return out

    def _cut(self, rows: List[List[int]], max_tokens: int) -> List[List[int]]:
        eos = self.config.tokenizer.eos_id
        out = []
        for row in rows:
            seq = []
            for t in row[:max_tokens]:
                if t == eos:
                    break
                seq.append(t)
            out.append(seq)
        return out

    def _run_images(self, images: Sequence[Any], prompts: Sequence[Sequence[int]], max_tokens: int,
                    sampling: Optional[dict] = None, settings: Optional[dict] = None) -> List[List[int]]:
        """Batched generation straight from raw images: when no image is pre-encoded the image prefix and the
        prompt are prefilled in one decoder pass (engine.caption_from_crops); otherwise (and when sampling)
        the two-step path."""
        lora = self._lora(settings)
        if sampling or lora is not None or any(isinstance(im, EncodedImage) for im in images):
            rows: List[List[int]] = []
            for lo in range(0, len(images), self._max_batch):
                rows += self._run(self.encode_images(images[lo: lo + self._max_batch], settings),
                                  prompts[lo: lo + self._max_batch], max_tokens, sampling=sampling, lora=lora)
            return rows
        out: List[List[int]] = []
        for lo in range(0, len(images), self._max_batch):
            arrs = [_as_array(im) for im in images[lo: lo + self._max_batch]]
            dev, offs, til = self.engine.stage_images(arrs)
            res = self.engine.caption_from_crops(dev, offs, til, prompts[lo: lo + self._max_batch], max_tokens)
            out += self._cut(res.tokens.tolist(), max_tokens)
        return out

    def _stream_text(self, tokens: Sequence[int]):
        """Streaming detokenisation with the reference's flush rules (moondream.py:476-537)."""
        dec = _StreamDecoder(self.tokenizer)
        for tok in tokens:
            out = dec.push(tok)
            if out:
                yield out
        out = dec.flush()
        if out:
            yield out

    def caption_batch(self, images: Sequence[Any], length: str = "normal",
                      settings: Optional[dict] = None) -> List[Dict[str, str]]:
        tpl = self.config.tokenizer.templates["caption"]
        if tpl is None:
            raise NotImplementedError("Model does not support captioning.")
        if length not in tpl:
            raise ValueError(f"Model does not support caption length '{length}'.")
        max_tokens, sampling = self._text_settings(settings)
        toks = self._run_images(images, [tpl[length]] * len(images), max_tokens, sampling, settings)
        return [{"caption": "".join(self._stream_text(t))} for t in toks]

    def caption(self, image, length: Literal["normal", "short", "long"] = "normal", stream: bool = False,
                settings: Optional[dict] = None):
        tpl = self.config.tokenizer.templates["caption"]
        if tpl is None:
            raise NotImplementedError("Model does not support captioning.")
        if length not in tpl:
            raise ValueError(f"Model does not support caption length '{length}'.")
        max_tokens, sampling = self._text_settings(settings)
        enc = self.encode_image(image, settings)
        lora = self._lora(settings)
        if stream and lora is None:
            return {"caption": self._stream_generate(enc, tpl[length], max_tokens, sampling)}
        toks = self._run([enc], [tpl[length]], max_tokens, sampling=sampling, lora=lora)[0]
        if stream:
            return {"caption": self._stream_text(toks)}
        return {"caption": "".join(self._stream_text(toks))}

    def _query_prompt(self, question: str, spatial_refs: Optional[SpatialRefs], with_bos: bool,
                      reasoning: bool = False) -> List[int]:
        tk = self.config.tokenizer
        toks = ([tk.bos_id] if with_bos else []) + list(tk.templates["query"]["prefix"])
        if spatial_refs:
            for ref in spatial_refs:
                toks += [tk.coord_id, tk.coord_id] if len(ref) == 2 else [tk.coord_id, tk.coord_id, tk.size_id]
        toks += self.tokenizer.encode(question).ids + list(tk.templates["query"]["suffix"])
        if reasoning:
            toks += [tk.thinking_id]                   # moondream.py:586-587
        else:
            # the reference appends the suffix a second time when reasoning is off (moondream.py:597-598)
            toks += list(tk.templates["query"]["suffix"])
        return toks

    def query_batch(self, images: Sequence[Any], questions: Sequence[str],
                    settings: Optional[dict] = None) -> List[Dict[str, str]]:
        if self.config.tokenizer.templates["query"] is None:
            raise NotImplementedError("Model does not support querying.")
        max_tokens, sampling = self._text_settings(settings)
        prompts = [self._query_prompt(q, None, False) for q in questions]
        toks = self._run_images(images, prompts, max_tokens, sampling, settings)
        return [{"answer": "".join(self._stream_text(t))} for t in toks]

    def query(self, image=None, question: str = None, reasoning: bool = False,
              spatial_refs: Optional[SpatialRefs] = None, stream: bool = False,
              settings: Optional[dict] = None):
        """moondream.py:541-618.  image=None is the text-only query (BOS + prompt at position 0 under a pure causal
        mask, :565-574); reasoning=True decodes the grounded chain of thought first (:576-596)."""
        if self.config.tokenizer.templates["query"] is None:
            raise NotImplementedError("Model does not support querying.")
        if question is None:
            raise ValueError("question must be provided.")
        if spatial_refs and image is None:
            raise ValueError("spatial_refs can only be used with an image.")
        max_tokens, sampling = self._text_settings(settings)
        if image is not None:
            enc = self.encode_image(image, settings)
            prefix_len = -1
        else:
            enc = EncodedImage(PrefixKV(0, [], self.engine.pages), self.engine)      # nothing cached: position 0
            prefix_len = 0
        prompt = self._query_prompt(question, spatial_refs, with_bos=image is None, reasoning=reasoning)
        embeds = self._prompt_embeds_with_refs(prompt, spatial_refs) if spatial_refs else None
        lora = self._lora(settings)
        if reasoning:
            if "sampler" in sampling:
                raise NotImplementedError("reasoning runs on the device; the host sampler is a parity-test path")
            if lora is not None:
                raise NotImplementedError("reasoning under a LoRA variant is not implemented (the variant path decodes "
                                          "eagerly; the reasoning loop is a captured graph)")
            tk = self.config.tokenizer
            r_toks, coords, a_toks = self.engine.generate_reasoning(
                [enc._prefix], [prompt], tk.templates["query"]["suffix"], max_tokens, prompt_embeds=embeds,
                prefix_len=prefix_len, **sampling)[0]
            text, grounding = self._reasoning_result(r_toks, coords)
            answer = self._stream_text(a_toks) if stream else "".join(self._stream_text(a_toks))
            return {"reasoning": {"text": text, "grounding": grounding}, "answer": answer}
        if stream and lora is None:
            return {"answer": self._stream_generate(enc, prompt, max_tokens, sampling, embeds, prefix_len)}
        toks = self._run([enc], [prompt], max_tokens, prompt_embeds=embeds, sampling=sampling, prefix_len=prefix_len,
                         lora=lora)[0]
        if stream:
            return {"answer": self._stream_text(toks)}
        return {"answer": "".join(self._stream_text(toks))}

    def _reasoning_result(self, tokens: Sequence[int], coords: Sequence[float]):
        """Text + grounding of _generate_reasoning (moondream.py:363-432): a new chunk starts at every
        start_ground_points / end_ground token; a chunk with >= 2 coordinates grounds its text span."""
        tk = self.config.tokenizer
        text_chunks: List[List[int]] = [[]]
        ground_chunks: List[List[float]] = [[]]
        for tok, c in zip(tokens, coords):
            if tok == tk.start_ground_points_id or tok == tk.end_ground_id:
                text_chunks.append([])
                ground_chunks.append([])
            text_chunks[-1].append(tok)
            if tok == tk.coord_id:
                ground_chunks[-1].append(float(c))
        texts = [self.tokenizer.decode(ch) for ch in text_chunks]
        grounding, start = [], 0
        for txt, g in zip(texts, ground_chunks):
            if len(g) > 1:
                pts = [(g[i], g[i + 1]) for i in range(0, len(g) - (len(g) % 2), 2)]
                grounding.append({"start_idx": start, "end_idx": start + len(txt), "points": pts})
            start += len(txt)
        return "".join(texts), grounding

    def _stream_generate(self, enc: EncodedImage, prompt: Sequence[int], max_tokens: int, sampling: dict,
                         prompt_embeds=None, prefix_len: int = -1, chunk: int = 8):
        """The streaming generator of moondream.py:470-537: text chunks are yielded while the decode loop is still
        running (every `chunk` graph replays), and closing the generator stops the loop and frees the pages."""
        if "sampler" in sampling:                 # host restatement: decode first, then detokenise
            yield from self._stream_text(self._run([enc], [prompt], max_tokens, prompt_embeds=prompt_embeds,
                                                   sampling=sampling, prefix_len=prefix_len)[0])
            return
        eos = self.config.tokenizer.eos_id
        dec = _StreamDecoder(self.tokenizer)
        gen = self.engine.generate_stream([enc._prefix], [prompt], max_tokens, chunk=chunk, prompt_embeds=prompt_embeds,
                                          prefix_len=prefix_len, **sampling)
        try:
            for part in gen:
                for tok in part[0].tolist():
                    if tok == eos:
                        out = dec.flush()
                        if out:
                            yield out
                        return
                    out = dec.push(tok)
                    if out:
                        yield out
            out = dec.flush()
            if out:
                yield out
        finally:
            gen.close()

    def _prompt_embeds_with_refs(self, prompt: List[int], spatial_refs: SpatialRefs) -> torch.Tensor:
        """Substitute region encodings for coord/size placeholder tokens (moondream.py:293-301,
        region.py:96-136)."""
        eng, tk = self.engine, self.config.tokenizer
        ids = torch.tensor(prompt, dtype=torch.int32, device=self.device)
        x = torch.empty((len(prompt), self.config.text.dim), dtype=torch.bfloat16, device=self.device)
        eng.embed(ids, x)
        coords, sizes = [], []
        for ref in spatial_refs:
            if len(ref) == 2:
                coords += [ref[0], ref[1]]
            else:
                coords += [(ref[0] + ref[2]) / 2, (ref[1] + ref[3]) / 2]
                sizes.append([ref[2] - ref[0], ref[3] - ref[1]])
        crow = [i for i, t in enumerate(prompt) if t == tk.coord_id]
        srow = [i for i, t in enumerate(prompt) if t == tk.size_id]
        if coords:
            x[torch.tensor(crow, device=self.device)] = eng.region_encode(0, torch.tensor(coords, dtype=torch.float32).view(-1, 1))
        if sizes:
            x[torch.tensor(srow, device=self.device)] = eng.region_encode(1, torch.tensor(sizes, dtype=torch.float32))
        return x

    # ------------------------------------------------------------------ detect / point
    def _object_prompts(self, kind: str, objects: Sequence[str]) -> List[List[int]]:
        tpl = self.config.tokenizer.templates[kind]
        return [list(tpl["prefix"]) + self.tokenizer.encode(" " + o).ids + list(tpl["suffix"]) for o in objects]

    def _points_batch(self, kind: str, images: Sequence[Any], objects: Sequence[str], settings: Optional[dict],
                      include_size: bool) -> List[List[dict]]:
        """detect / point for any number of images: at most `max_batch` sequences walk the region head in lock-step
        at a time (like `_run_images`), and a chunk's image prefixes go back to the pool before the next one is encoded."""
        if len(objects) != len(images):
            raise ValueError("one object name per image")
        max_objects = settings.get("max_objects", DEFAULT_MAX_OBJECTS) if settings else DEFAULT_MAX_OBJECTS
        lora = self._lora(settings)
        out: List[List[dict]] = []
        for lo in range(0, len(images), self._max_batch):
            enc = self.encode_images(images[lo: lo + self._max_batch], settings)
            out += self.engine.generate_points([e._prefix for e in enc],
                                               self._object_prompts(kind, objects[lo: lo + self._max_batch]),
                                               include_size=include_size, max_objects=max_objects, lora=lora)
        return out

    def detect_batch(self, images: Sequence[Any], objects: Sequence[str],
                     settings: Optional[dict] = None) -> List[Dict[str, list]]:
        if self.config.tokenizer.templates["detect"] is None:
            raise NotImplementedError("Model does not support object detection.")
        res = self._points_batch("detect", images, objects, settings, include_size=True)
        return [{"objects": [{k: o[k] for k in ("x_min", "y_min", "x_max", "y_max")} for o in r]} for r in res]

    def detect(self, image, object: str, settings: Optional[dict] = None):
        return self.detect_batch([self.encode_image(image, settings)], [object], settings)[0]

    def point_batch(self, images: Sequence[Any], objects: Sequence[str],
                    settings: Optional[dict] = None) -> List[Dict[str, list]]:
        if self.config.tokenizer.templates["point"] is None:
            raise NotImplementedError("Model does not support pointing.")
        res = self._points_batch("point", images, objects, settings, include_size=False)
        return [{"points": [{"x": o["x"], "y": o["y"]} for o in r]} for r in res]

    def point(self, image, object: str, settings: Optional[dict] = None):
        return self.point_batch([self.encode_image(image, settings)], [object], settings)[0]

    def detect_gaze(self, image, eye=None, face=None, unstable_settings: Dict[str, Any] = {}):
        raise NotImplementedError("detect_gaze (moondream.py:831-973) is out of the hot path's scope "
                                  "(SURVEY.md §8a lists detect/point; gaze is an application of point)")


class _StreamDecoder:
    """Token-at-a-time detokeniser with the reference's flush rules (moondream.py:476-510, 531-537): text is released
    after a newline (cache reset), after a CJK character, or up to the last space."""

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer
        self.cache: List[int] = []
        self.print_len = 0

    def push(self, tok: int) -> str:
        self.cache.append(tok)
        text = self.tokenizer.decode(self.cache)
        if text.endswith("\n"):
            chunk = text[self.print_len:]
            self.cache, self.print_len = [], 0
            return chunk
        if len(text) > 0 and _is_cjk_char(ord(text[-1])):
            chunk = text[self.print_len:]
            self.print_len += len(chunk)
            return chunk
        sp = text.rfind(" ", self.print_len)
        if sp >= self.print_len:
            chunk = text[self.print_len: sp + 1]
            self.print_len += len(chunk)
            return chunk
        return ""

    def flush(self) -> str:
        if not self.cache:
            return ""
        chunk = self.tokenizer.decode(self.cache)[self.print_len:]
        self.cache, self.print_len = [], 0
        return chunk


def _is_cjk_char(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF) or (0x3400 <= cp <= 0x4DBF) or (0x2F800 <= cp <= 0x2FA1F)

"""Drop-in for ``moondream.torch.moondream.MoondreamModel`` (reference moondream/torch/moondream.py)
on the path BASELINE.json names: ``encode_image`` / ``caption`` / ``query`` / ``detect`` / ``point``
with the reference's signatures, settings keys, result dict shapes and exceptions — plus the batched
calls the reference cannot offer (it is strictly batch-1: moondream.py:66,369,482,669):
``encode_images`` / ``caption_batch`` / ``query_batch`` / ``detect_batch`` / ``point_batch``.

All arithmetic runs in libmoondream_b200.so; this module keeps what the reference also keeps in
Python: crop geometry (PIL), tokenizer, prompt templates, result shaping, streaming detokenisation.

Deliberate differences from the reference (documented in DESIGN.md):
  * ``settings`` without "variant" is accepted by ``encode_image`` (the reference raises KeyError at
    moondream.py:240-243); LoRA variants (lora.py) need the network and are rejected with
    NotImplementedError when requested.
  * temperature > 0 follows the reference's softmax / top-p / multinomial arithmetic (moondream.py:270-278,
    312-318) on the host from each step's logits (moondream_b200/sampling.py): functionally the reference's
    default behaviour, but one synchronisation per token; the CUDA-graph path is the greedy one.
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
                 tokenizer=None, device: str = "cuda", max_batch: int = 32, kv_pages: Optional[int] = None):
        if dtype != torch.bfloat16:
            raise ValueError("the path is bf16 end to end, like the reference (vision.py:36, weights.py:32)")
        self.config = config
        self._device = torch.device(device)
        self._tokenizer = tokenizer
        self._max_batch = max_batch
        self._kv_pages = kv_pages
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

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = False):
        sd = {k.replace("model.", "", 1) if k.startswith("model.") else k: v for k, v in state_dict.items()}
        self._engine = Engine(self.config, sd, device=self._device, kv_pages=self._kv_pages,
                              max_batch=self._max_batch)
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
        """(max_tokens, sampler) from the reference's TextSamplingSettings (moondream.py:443-454): sampler is None
        for temperature 0 (greedy, CUDA-graph decode), else the host nucleus sampler."""
        temperature = settings.get("temperature", DEFAULT_TEMPERATURE) if settings else DEFAULT_TEMPERATURE
        top_p = settings.get("top_p", DEFAULT_TOP_P) if settings else DEFAULT_TOP_P
        if settings and settings.get("variant") is not None:
            raise NotImplementedError("LoRA variants are downloaded from the network by the reference "
                                      "(lora.py:23-40) and are not supported offline")
        max_tokens = settings.get("max_tokens", DEFAULT_MAX_TOKENS) if settings else DEFAULT_MAX_TOKENS
        sampler = None
        if temperature != 0:
            from .sampling import HostSampler
            sampler = HostSampler(temperature, top_p)          # ValueError for a negative temperature
        return max_tokens, sampler

    # ------------------------------------------------------------------ image encoding
    def encode_images(self, images: Sequence[Any]) -> List[EncodedImage]:
        todo = [(i, _as_array(im)) for i, im in enumerate(images) if not isinstance(im, EncodedImage)]
        out: List[Optional[EncodedImage]] = [im if isinstance(im, EncodedImage) else None for im in images]
        for lo in range(0, len(todo), self._max_batch):
            chunk = todo[lo: lo + self._max_batch]
            prefixes = self.engine.encode_images([a for _, a in chunk])
            for (i, _), p in zip(chunk, prefixes):
                out[i] = EncodedImage(p, self.engine)
        return out  # type: ignore[return-value]

    def encode_image(self, image, settings: Optional[dict] = None) -> EncodedImage:
        if isinstance(image, EncodedImage):
            return image
        if settings and settings.get("variant") is not None:
            raise NotImplementedError("LoRA variants are not supported offline")
        return self.encode_images([image])[0]

    # ------------------------------------------------------------------ text generation
    def _run(self, encoded: Sequence[EncodedImage], prompts: Sequence[Sequence[int]], max_tokens: int,
             eos_id: Optional[int] = None, prompt_embeds=None, sampler=None) -> List[List[int]]:
        eos = self.config.tokenizer.eos_id if eos_id is None else eos_id
        res = self.engine.generate([e._prefix for e in encoded], prompts, max_tokens,
                                   prompt_embeds=prompt_embeds, sampler=sampler)
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
                    sampler=None) -> List[List[int]]:
        """Batched generation straight from raw images: when no image is pre-encoded the image prefix and the
        prompt are prefilled in one decoder pass (engine.caption_from_crops); otherwise (and when sampling)
        the two-step path."""
        if sampler is not None or any(isinstance(im, EncodedImage) for im in images):
            rows: List[List[int]] = []
            for lo in range(0, len(images), self._max_batch):
                rows += self._run(self.encode_images(images[lo: lo + self._max_batch]),
                                  prompts[lo: lo + self._max_batch], max_tokens, sampler=sampler)
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
        cache: List[int] = []
        print_len = 0
        for tok in tokens:
            cache.append(tok)
            text = self.tokenizer.decode(cache)
            if text.endswith("\n"):
                chunk = text[print_len:]
                cache, print_len = [], 0
                if chunk:
                    yield chunk
            elif len(text) > 0 and _is_cjk_char(ord(text[-1])):
                chunk = text[print_len:]
                print_len += len(chunk)
                if chunk:
                    yield chunk
            else:
                sp = text.rfind(" ", print_len)
                if sp >= print_len:
                    chunk = text[print_len: sp + 1]
                    print_len += len(chunk)
                    if chunk:
                        yield chunk
        if cache:
            chunk = self.tokenizer.decode(cache)[print_len:]
            if chunk:
                yield chunk

    def caption_batch(self, images: Sequence[Any], length: str = "normal",
                      settings: Optional[dict] = None) -> List[Dict[str, str]]:
        tpl = self.config.tokenizer.templates["caption"]
        if tpl is None:
            raise NotImplementedError("Model does not support captioning.")
        if length not in tpl:
            raise ValueError(f"Model does not support caption length '{length}'.")
        max_tokens, sampler = self._text_settings(settings)
        toks = self._run_images(images, [tpl[length]] * len(images), max_tokens, sampler)
        return [{"caption": "".join(self._stream_text(t))} for t in toks]

    def caption(self, image, length: Literal["normal", "short", "long"] = "normal", stream: bool = False,
                settings: Optional[dict] = None):
        tpl = self.config.tokenizer.templates["caption"]
        if tpl is None:
            raise NotImplementedError("Model does not support captioning.")
        if length not in tpl:
            raise ValueError(f"Model does not support caption length '{length}'.")
        max_tokens, sampler = self._text_settings(settings)
        enc = self.encode_image(image, settings)
        toks = self._run([enc], [tpl[length]], max_tokens, sampler=sampler)[0]
        if stream:
            return {"caption": self._stream_text(toks)}
        return {"caption": "".join(self._stream_text(toks))}

    def _query_prompt(self, question: str, spatial_refs: Optional[SpatialRefs], with_bos: bool) -> List[int]:
        tk = self.config.tokenizer
        toks = ([tk.bos_id] if with_bos else []) + list(tk.templates["query"]["prefix"])
        if spatial_refs:
            for ref in spatial_refs:
                toks += [tk.coord_id, tk.coord_id] if len(ref) == 2 else [tk.coord_id, tk.coord_id, tk.size_id]
        toks += self.tokenizer.encode(question).ids + list(tk.templates["query"]["suffix"])
        # the reference appends the suffix a second time when reasoning is off (moondream.py:586-604)
        toks += list(tk.templates["query"]["suffix"])
        return toks

    def query_batch(self, images: Sequence[Any], questions: Sequence[str],
                    settings: Optional[dict] = None) -> List[Dict[str, str]]:
        if self.config.tokenizer.templates["query"] is None:
            raise NotImplementedError("Model does not support querying.")
        max_tokens, sampler = self._text_settings(settings)
        prompts = [self._query_prompt(q, None, False) for q in questions]
        toks = self._run_images(images, prompts, max_tokens, sampler)
        return [{"answer": "".join(self._stream_text(t))} for t in toks]

    def query(self, image=None, question: str = None, reasoning: bool = False,
              spatial_refs: Optional[SpatialRefs] = None, stream: bool = False,
              settings: Optional[dict] = None):
        if self.config.tokenizer.templates["query"] is None:
            raise NotImplementedError("Model does not support querying.")
        if question is None:
            raise ValueError("question must be provided.")
        if spatial_refs and image is None:
            raise ValueError("spatial_refs can only be used with an image.")
        if reasoning:
            raise NotImplementedError("reasoning=True (grounded chain of thought, moondream.py:323-432) "
                                      "is listed as 'next' in DESIGN.md")
        if image is None:
            raise NotImplementedError("text-only query (pure causal mask, moondream.py:565-574) is "
                                      "listed as 'next' in DESIGN.md")
        max_tokens, sampler = self._text_settings(settings)
        enc = self.encode_image(image, settings)
        prompt = self._query_prompt(question, spatial_refs, False)
        embeds = None
        if spatial_refs:
            embeds = self._prompt_embeds_with_refs(prompt, spatial_refs)
        toks = self._run([enc], [prompt], max_tokens, prompt_embeds=embeds, sampler=sampler)[0]
        if stream:
            return {"answer": self._stream_text(toks)}
        return {"answer": "".join(self._stream_text(toks))}

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

    def detect_batch(self, images: Sequence[Any], objects: Sequence[str],
                     settings: Optional[dict] = None) -> List[Dict[str, list]]:
        if self.config.tokenizer.templates["detect"] is None:
            raise NotImplementedError("Model does not support object detection.")
        max_objects = settings.get("max_objects", DEFAULT_MAX_OBJECTS) if settings else DEFAULT_MAX_OBJECTS
        enc = self.encode_images(images)
        res = self.engine.generate_points([e._prefix for e in enc], self._object_prompts("detect", objects),
                                          include_size=True, max_objects=max_objects)
        return [{"objects": [{k: o[k] for k in ("x_min", "y_min", "x_max", "y_max")} for o in r]} for r in res]

    def detect(self, image, object: str, settings: Optional[dict] = None):
        return self.detect_batch([self.encode_image(image, None)], [object], settings)[0]

    def point_batch(self, images: Sequence[Any], objects: Sequence[str],
                    settings: Optional[dict] = None) -> List[Dict[str, list]]:
        if self.config.tokenizer.templates["point"] is None:
            raise NotImplementedError("Model does not support pointing.")
        max_objects = settings.get("max_objects", DEFAULT_MAX_OBJECTS) if settings else DEFAULT_MAX_OBJECTS
        enc = self.encode_images(images)
        res = self.engine.generate_points([e._prefix for e in enc], self._object_prompts("point", objects),
                                          include_size=False, max_objects=max_objects)
        return [{"points": [{"x": o["x"], "y": o["y"]} for o in r]} for r in res]

    def point(self, image, object: str, settings: Optional[dict] = None):
        return self.point_batch([self.encode_image(image, None)], [object], settings)[0]

    def detect_gaze(self, image, eye=None, face=None, unstable_settings: Dict[str, Any] = {}):
        raise NotImplementedError("detect_gaze (moondream.py:831-973) is out of the hot path's scope "
                                  "(SURVEY.md §8a lists detect/point; gaze is an application of point)")


def _is_cjk_char(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF) or (0x3400 <= cp <= 0x4DBF) or (0x2F800 <= cp <= 0x2FA1F)

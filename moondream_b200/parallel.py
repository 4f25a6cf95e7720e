"""Multi-GPU: independent replicas, images sharded contiguously, ONE collective.

Every image (its crops, 730-token prefix, KV pages and decode loop) is independent of every other
image, so the path shards with no data-path collective; the only exchange is the all-gather of the
finished token ids (int32 [B/G, max_tokens + 1] per rank) over NCCL/NVLink.  The reference has no
distributed code at all.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `n_items` for `rank` (first n % world ranks get one extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of items (e.g. images weighted by crop count, 2..13)
    to ranks so ViT work balances when image sizes differ.  Returns item indices per rank."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    for r in range(world):
        out[r].sort()
    return out


def gather_tokens(local_tokens: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather per-rank token matrices [n_r, T] (n_r = counts[r]) into [sum(n_r), T] on every rank.
    Ragged counts are padded to the maximum for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    assert len(counts) == world
    n_max, T = max(counts), local_tokens.shape[1]
    pad = torch.zeros((n_max, T), dtype=local_tokens.dtype, device=local_tokens.device)
    pad[: local_tokens.shape[0]] = local_tokens
    out = torch.empty((world * n_max, T), dtype=local_tokens.dtype, device=local_tokens.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * n_max: r * n_max + counts[r]] for r in range(world)], dim=0)


# --------------------------------------------------------------------------------------------------
# Product entry points: one process per GPU, every rank holds a replica of the model and receives the SAME request
# batch; each rank runs its shard of the images through the local engine and the results are all-gathered, so every
# rank returns the full answer in request order.  (bench.py's N > 1 runs go through ShardedEngine.)
# --------------------------------------------------------------------------------------------------
class ShardedEngine:
    """Shards a request batch over the ranks of `group` by crop count (LPT) and gathers token ids / boxes.

    `engine` needs `stage_images(images) -> (crops, offsets, tilings)`, `caption_from_crops(...)`,
    `encode_images(images)`, `generate_points(...)` and `cfg` — i.e. moondream_b200.engine.Engine (the gloo tests
    on CPU pass a stub with the same methods)."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.dist = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.dist else 0
        self.world = dist.get_world_size(group) if self.dist else 1

    # ---- planning (pure host logic, identical on every rank)
    def plan(self, images) -> List[List[int]]:
        from .image_crops import crop_tiling

        v = self.engine.cfg.vision
        kw = dict(overlap_margin=v.overlap_margin, max_crops=v.max_crops, base_size=(v.crop_size, v.crop_size),
                  patch_size=v.enc_patch_size)
        costs = []
        for im in images:
            th, tw = crop_tiling(im.shape, **kw)
            costs.append(th * tw + 1)
        return shard_by_cost(costs, self.world)

    def _to_request_order(self, gathered: torch.Tensor, parts: List[List[int]]) -> torch.Tensor:
        order = [i for p in parts for i in p]                 # gathered row j belongs to request order[j]
        inv = torch.empty(len(order), dtype=torch.long)
        inv[torch.tensor(order, dtype=torch.long)] = torch.arange(len(order))
        return gathered.index_select(0, inv.to(gathered.device))

    def _gather(self, local: torch.Tensor, parts: List[List[int]]) -> torch.Tensor:
        if self.world == 1:
            return local
        return self._to_request_order(gather_tokens(local, [len(p) for p in parts], self.group), parts)

    # ---- generation
    def stage(self, images):
        """Host images of this rank's shard -> device crops (pinned staging + H2D)."""
        parts = self.plan(images)
        mine = parts[self.rank]
        staged = self.engine.stage_images([images[i] for i in mine]) if mine else None
        return parts, staged

    def caption_tokens(self, images, prompts, max_tokens: int, staged=None, to_host: bool = True, **gen_kw) -> torch.Tensor:
        """Token ids int32 [len(images), max_tokens + 1] in request order, identical on every rank: local
        ViT + prefill + decode of this rank's shard, then ONE all-gather over NCCL/NVLink."""
        parts, st = staged if staged is not None else self.stage(images)
        mine = parts[self.rank]
        dev = self.engine.device
        if mine:
            crops, offs, til = st
            res = self.engine.caption_from_crops(crops, offs, til, [prompts[i] for i in mine], max_tokens,
                                                 to_host=False, **gen_kw)
            local = res.tokens
        else:
            local = torch.zeros((0, max_tokens + 1), dtype=torch.int32, device=dev)
        out = self._gather(local, parts)
        return out.to("cpu") if to_host else out

    def detect_boxes(self, images, prompts, max_objects: int, include_size: bool = True):
        """detect / point for the request batch: returns (values float32 [N, max_objects, 4 or 2], counts int32 [N])
        in request order on every rank (boxes as x_min, y_min, x_max, y_max; points as x, y)."""
        parts = self.plan(images)
        mine = parts[self.rank]
        width = 4 if include_size else 2
        dev = self.engine.device
        vals = torch.zeros((len(mine), max_objects, width), dtype=torch.float32)
        cnt = torch.zeros((len(mine), 1), dtype=torch.float32)
        if mine:
            enc = self.engine.encode_images([images[i] for i in mine])
            res = self.engine.generate_points(enc, [prompts[i] for i in mine], include_size, max_objects)
            keys = ("x_min", "y_min", "x_max", "y_max") if include_size else ("x", "y")
            for j, objs in enumerate(res):
                cnt[j, 0] = len(objs)
                for k, o in enumerate(objs[:max_objects]):
                    vals[j, k] = torch.tensor([o[q] for q in keys])
        flat = torch.cat([vals.view(len(mine), -1), cnt], dim=1).to(dev)
        full = self._gather(flat, parts).to("cpu")
        n = full.shape[0]
        return full[:, :-1].view(n, max_objects, width), full[:, -1].to(torch.int32)


class ShardedModel:
    """`MoondreamModel`'s batched calls across the GPUs of one box: construct one per rank around that rank's model
    (after torch.distributed.init_process_group); every rank passes the same request and gets the same answer."""

    def __init__(self, model, group=None):
        self.model = model
        self.sharded = ShardedEngine(model.engine, group)

    @staticmethod
    def _no_variant(settings):
        # the sharded calls go through the fused encode + prompt prefill, which has no adapter slots; silently
        # answering without the adapters would be wrong, so say so (MoondreamModel's own batch calls accept variants)
        if settings and settings.get("variant") is not None:
            raise NotImplementedError('settings["variant"] (LoRA) is not supported by the sharded batch calls')

    def _generate(self, images, prompts, settings):
        from .moondream import _as_array

        self._no_variant(settings)
        max_tokens, sampling = self.model._text_settings(settings)
        if "sampler" in sampling:
            raise NotImplementedError("the host sampler is a single-process parity path")
        arrs = [_as_array(im) for im in images]
        toks = self.sharded.caption_tokens(arrs, prompts, max_tokens, **sampling)
        return self.model._cut(toks.tolist(), max_tokens)

    def caption_batch(self, images, length: str = "normal", settings=None):
        tpl = self.model.config.tokenizer.templates["caption"]
        if tpl is None:
            raise NotImplementedError("Model does not support captioning.")
        if length not in tpl:
            raise ValueError(f"Model does not support caption length '{length}'.")
        rows = self._generate(images, [tpl[length]] * len(images), settings)
        return [{"caption": "".join(self.model._stream_text(t))} for t in rows]

    def query_batch(self, images, questions, settings=None):
        if self.model.config.tokenizer.templates["query"] is None:
            raise NotImplementedError("Model does not support querying.")
        prompts = [self.model._query_prompt(q, None, False) for q in questions]
        rows = self._generate(images, prompts, settings)
        return [{"answer": "".join(self.model._stream_text(t))} for t in rows]

    def _points(self, kind, images, objects, settings, include_size):
        from .moondream import DEFAULT_MAX_OBJECTS, _as_array

        if self.model.config.tokenizer.templates[kind] is None:
            raise NotImplementedError(f"Model does not support {kind}.")
        self._no_variant(settings)
        max_objects = settings.get("max_objects", DEFAULT_MAX_OBJECTS) if settings else DEFAULT_MAX_OBJECTS
        vals, cnt = self.sharded.detect_boxes([_as_array(im) for im in images], self.model._object_prompts(kind, objects),
                                              max_objects, include_size)
        return vals, cnt

    def detect_batch(self, images, objects, settings=None):
        vals, cnt = self._points("detect", images, objects, settings, True)
        return [{"objects": [dict(zip(("x_min", "y_min", "x_max", "y_max"), map(float, vals[i, k]))) for k in range(int(cnt[i]))]}
                for i in range(vals.shape[0])]

    def point_batch(self, images, objects, settings=None):
        vals, cnt = self._points("point", images, objects, settings, False)
        return [{"points": [dict(zip(("x", "y"), map(float, vals[i, k]))) for k in range(int(cnt[i]))]}
                for i in range(vals.shape[0])]

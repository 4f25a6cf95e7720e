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

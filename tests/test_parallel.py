"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch and all-gather token ids."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moondream_b200.parallel import gather_tokens, shard_by_cost, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_by_cost_balances_crops():
    costs = [13, 2, 2, 10, 9, 2, 13, 5]
    parts = shard_by_cost(costs, 2)
    assert sorted(i for p in parts for i in p) == list(range(len(costs)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 2


def _worker(rank, world, port, n_items, T):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]
        lo, hi = shard_range(n_items, rank, world)
        # each "image" i produces the token row [i, i+1, ...] on the rank that owns it
        local = (torch.arange(lo, hi, dtype=torch.int32).view(-1, 1) + torch.arange(T, dtype=torch.int32).view(1, -1))
        full = gather_tokens(local, counts)
        want = torch.arange(n_items, dtype=torch.int32).view(-1, 1) + torch.arange(T, dtype=torch.int32).view(1, -1)
        assert torch.equal(full, want), (rank, full, want)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("n_items", [8, 7])
def test_two_rank_gloo_allgather(n_items):
    mp.spawn(_worker, args=(2, _free_port(), n_items, 5), nprocs=2, join=True)


@pytest.mark.parametrize("world,n_items", [(4, 10), (3, 2)])
def test_wider_worlds_with_ragged_and_empty_shards(world, n_items):
    """The 4- and 8-GPU runs use the same code as the 2-GPU one; cover a remainder (10 items on 4 ranks) and ranks
    that own nothing (2 items on 3 ranks) on CPU."""
    mp.spawn(_worker, args=(world, _free_port(), n_items, 3), nprocs=world, join=True)

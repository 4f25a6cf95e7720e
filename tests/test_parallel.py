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


class _StubEngine:
    """CPU stand-in with the methods ShardedEngine needs: a sequence's "tokens" encode which image and prompt it saw."""

    def __init__(self):
        from moondream_b200 import config as C

        self.cfg = C.tiny()
        self.device = torch.device("cpu")

    def stage_images(self, images):
        return images, None, None

    def caption_from_crops(self, crops, offs, til, prompts, max_tokens, to_host=False, **kw):
        rows = [[int(im[0, 0, 0]), int(im.shape[0]), len(p)] + [0] * (max_tokens - 2) for im, p in zip(crops, prompts)]
        return type("R", (), {"tokens": torch.tensor(rows, dtype=torch.int32)})()

    def encode_images(self, images):
        return images

    def generate_points(self, enc, prompts, include_size, max_objects):
        return [[{"x_min": float(im[0, 0, 0]), "y_min": float(k), "x_max": 1.0, "y_max": 2.0} for k in range(int(im[0, 0, 0]) % 3)]
                for im in enc]


def _sharded_worker(rank, world, port):
    import numpy as np

    from moondream_b200.parallel import ShardedEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [(378, 378), (800, 600), (378, 378), (1200, 900), (500, 700), (300, 200), (378, 378)]
        images = []
        for i, (h, w) in enumerate(sizes):
            im = np.zeros((h, w, 3), dtype=np.uint8)
            im[0, 0, 0] = 10 + i
            images.append(im)
        prompts = [[1] * (i + 1) for i in range(len(images))]
        se = ShardedEngine(_StubEngine())
        parts = se.plan(images)
        assert sorted(i for p in parts for i in p) == list(range(len(images)))
        toks = se.caption_tokens(images, prompts, 4)
        want = torch.tensor([[10 + i, h, i + 1, 0, 0] for i, (h, w) in enumerate(sizes)], dtype=torch.int32)
        assert torch.equal(toks, want), (rank, toks, want)          # request order on every rank
        vals, cnt = se.detect_boxes(images, prompts, 3)
        assert cnt.tolist() == [(10 + i) % 3 for i in range(len(images))]
        for i in range(len(images)):
            for k in range(int(cnt[i])):
                assert vals[i, k].tolist() == [10.0 + i, float(k), 1.0, 2.0]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_engine_returns_request_order_on_every_rank(world):
    """ShardedEngine (the N > 1 product path, also what bench.py drives): LPT plan by crop count, local generation,
    one all-gather, rows back in request order — with ragged shards."""
    mp.spawn(_sharded_worker, args=(world, _free_port()), nprocs=world, join=True)


def test_sharded_model_rejects_what_it_would_silently_ignore():
    """settings["variant"] has no slot in the fused sharded path; answering without the adapters would be silently wrong"""
    from moondream_b200 import config as C
    from moondream_b200.parallel import ShardedModel

    cfg = C.tiny()
    model = type("M", (), {"engine": _StubEngine(), "config": cfg})()
    sm = ShardedModel(model)
    for call in (lambda: sm.caption_batch([], "short", settings={"variant": "v1"}),
                 lambda: sm.query_batch([], [], settings={"variant": "v1"}),
                 lambda: sm.detect_batch([], [], settings={"variant": "v1"}),
                 lambda: sm.point_batch([], [], settings={"variant": "v1"})):
        with pytest.raises(NotImplementedError):
            call()


def _sharded_model_worker(rank, world, port):
    import numpy as np

    from moondream_b200 import config as C
    from moondream_b200.moondream import MoondreamModel
    from moondream_b200.parallel import ShardedModel
    from oracle.reference_shim import StubTokenizer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = C.tiny()
        model = MoondreamModel(cfg, tokenizer=StubTokenizer(cfg.text.vocab_size))
        model._engine = _StubEngine()
        sm = ShardedModel(model)
        sizes = [(378, 378), (800, 600), (378, 378), (1200, 900), (500, 700)]
        images = []
        for i, (h, w) in enumerate(sizes):
            im = np.zeros((h, w, 3), dtype=np.uint8)
            im[0, 0, 0] = 10 + i
            images.append(im)
        # the stub's "tokens" are [image id, image height, prompt length, 0, ...]; the stub tokenizer prints ids
        caps = sm.caption_batch(images, "short", settings={"temperature": 0, "max_tokens": 4})
        n_cap = len(cfg.tokenizer.templates["caption"]["short"])
        assert [c["caption"] for c in caps] == [f"{10 + i} {h} {n_cap} " for i, (h, w) in enumerate(sizes)]     # 0 = eos: cut
        qs = ["7", "7 8", "7 8 9", "7", "7 8"]
        ans = sm.query_batch(images, qs, settings={"temperature": 0, "max_tokens": 4})
        tq = cfg.tokenizer.templates["query"]
        assert [a["answer"] for a in ans] == [f"{10 + i} {h} {len(tq['prefix']) + len(q.split()) + 2 * len(tq['suffix'])} "
                                              for (i, (h, w)), q in zip(enumerate(sizes), qs)]
        det = sm.detect_batch(images, ["17"] * 5, settings={"max_objects": 3})
        assert [len(d["objects"]) for d in det] == [(10 + i) % 3 for i in range(5)]
        assert det[1]["objects"][0] == {"x_min": 11.0, "y_min": 0.0, "x_max": 1.0, "y_max": 2.0}
        with pytest.raises(NotImplementedError):
            sm.caption_batch(images, "short", settings={"temperature": 0.5, "host_sampler": True})
        with pytest.raises(ValueError):
            sm.caption_batch(images, "epic")
    finally:
        dist.destroy_process_group()


def test_sharded_model_api_on_two_ranks():
    """ShardedModel (MoondreamModel's batched calls over the ranks of a box) on 2 gloo ranks with a stub engine: every
    rank returns the full answer in request order, shaped like the single-process API."""
    mp.spawn(_sharded_model_worker, args=(2, _free_port()), nprocs=2, join=True)

"""KV page bookkeeping (host side of row a19, SURVEY.md §8: EncodedImage owns pages instead of cloned tensors)."""
import gc

import pytest
import torch

from moondream_b200 import _native as N, config as C
from moondream_b200.engine import PAGE, PagePool, PrefixKV


def test_pool_layout_alloc_release_and_exhaustion():
    cfg = C.tiny()
    pool = PagePool(cfg, 6, "cpu")
    t = cfg.text
    assert tuple(pool.pool.shape) == (t.n_layers, 6, 2, t.n_heads, PAGE, 64) and pool.pool.dtype == torch.bfloat16
    assert float(pool.pool.abs().max()) == 0.0                      # unwritten slots must be finite (0 * v in P.V)
    a = pool.alloc(4)
    assert len(set(a)) == 4 and all(0 <= p < 6 for p in a) and pool.free_pages == 2
    with pytest.raises(N.NativeError):
        pool.alloc(3)
    assert pool.free_pages == 2                                     # a failed request takes nothing
    pool.release(a[:2])
    b = pool.alloc(4)
    assert set(b).isdisjoint(a[2:]) and pool.free_pages == 0


def test_prefix_returns_its_pages_once():
    pool = PagePool(C.tiny(), 4, "cpu")
    pre = PrefixKV(730, pool.alloc(3), pool)
    assert pool.free_pages == 1
    pre.release()
    pre.release()                                                   # idempotent
    assert pool.free_pages == 4
    PrefixKV(730, pool.alloc(2), pool)                              # dropped handle: pages come back with it
    gc.collect()
    assert pool.free_pages == 4


def _engine_shell(n_pages: int):
    """An Engine with only the attributes `_sequence_tables` touches, on the CPU (the method is pure bookkeeping
    plus one tensor copy, so it can be exercised without the library or a GPU)."""
    from moondream_b200.engine import Engine

    cfg = C.tiny()
    eng = Engine.__new__(Engine)
    eng.cfg, eng.device = cfg, torch.device("cpu")
    eng.pages = PagePool(cfg, n_pages, "cpu")
    eng.max_blocks = cfg.text.max_context // PAGE
    return eng


def test_sequence_tables_share_full_pages_and_copy_the_partial_one():
    eng = _engine_shell(20)
    pre = PrefixKV(730, eng.pages.alloc(12), eng.pages)             # 11 full pages + 26 tokens of a 12th
    eng.pages.pool[:, pre.pages[11]] = 3.0                          # recognisable content of the partial page
    bt, owned = eng._sequence_tables([pre, pre], 730 + 40, consume=False)    # 770 tokens -> 13 blocks each
    assert eng.pages.free_pages == 20 - 12 - 4 and not pre._released
    for i in range(2):
        row = bt[i].tolist()
        assert row[:11] == pre.pages[:11]                           # full prefix pages are shared, not copied
        assert row[11:13] == owned[i] and row[13:] == [0] * (eng.max_blocks - 13)
        assert float(eng.pages.pool[:, owned[i][0]].min()) == 3.0   # copy-on-write of the partially filled page
        assert float(eng.pages.pool[:, owned[i][1]].abs().max()) == 0.0
    assert set(owned[0]).isdisjoint(owned[1]) and set(owned[0] + owned[1]).isdisjoint(pre.pages)
    for own in owned:
        eng.pages.release(own)
    pre.release()
    assert eng.pages.free_pages == 20


def test_sequence_tables_take_nothing_when_the_pool_cannot_serve_the_batch():
    """an exhausted pool part-way through a batch must not leak the pages taken for earlier sequences
    nor orphan consumed prefixes."""
    eng = _engine_shell(27)
    pres = [PrefixKV(730, eng.pages.alloc(12), eng.pages) for _ in range(2)]     # 3 pages left
    for consume in (False, True):
        with pytest.raises(N.NativeError):
            eng._sequence_tables(pres, 730 + 200, consume=consume)               # 15 blocks each: 2 x 4 (or 2 x 3) > 3
        assert eng.pages.free_pages == 3 and not any(p._released for p in pres)
    bt, owned = eng._sequence_tables(pres[:1], 730 + 200, consume=True)          # one sequence fits: 3 fresh pages
    assert pres[0]._released and owned[0][:12] == pres[0].pages and eng.pages.free_pages == 0
    assert bt[0, :15].tolist() == owned[0]
    # a prefix already handed over cannot be consumed again; what the call took before noticing goes back
    eng.pages.release(owned[0][12:])
    with pytest.raises(ValueError):
        eng._sequence_tables([pres[1], pres[0]], 730 + 30, consume=True)
    assert eng.pages.free_pages == 3 and not pres[1]._released and pres[0]._released
    with pytest.raises(ValueError):
        eng._sequence_tables(pres[1:], eng.cfg.text.max_context + 1, consume=False)


def test_lru_helpers_bound_the_graph_cache():
    from moondream_b200.engine import DecodeMode, Engine, _lru_get, _lru_put

    cache = {}
    modes = [DecodeMode(False, 0.1 * i, 0.3, False, 1, -1, 0) for i in range(Engine._MAX_DECODE_GRAPHS + 3)]
    for i, m in enumerate(modes[: Engine._MAX_DECODE_GRAPHS]):
        _lru_put(cache, m, i, Engine._MAX_DECODE_GRAPHS)
    assert _lru_get(cache, modes[0]) == 0                           # a hit makes the oldest entry the youngest
    for i, m in enumerate(modes[Engine._MAX_DECODE_GRAPHS:]):
        _lru_put(cache, m, 100 + i, Engine._MAX_DECODE_GRAPHS)
    assert len(cache) == Engine._MAX_DECODE_GRAPHS
    assert modes[0] in cache and modes[1] not in cache and modes[2] not in cache and modes[3] not in cache
    assert _lru_get(cache, modes[1]) is None and list(cache)[-1] == modes[-1]
    assert modes[0] == DecodeMode(False, 0.0, 0.3, False, 1, -1, 0) and hash(modes[0]) == hash(DecodeMode(False, 0.0, 0.3, False, 1, -1, 0))

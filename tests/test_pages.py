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

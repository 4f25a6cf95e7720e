"""Streaming detokenisation (reference moondream.py:476-537, row a21 of SURVEY.md §8): the chunks must be exactly the
ones the reference's generator yields.  tests/golden/streaming.json was produced by the unmodified reference (its own
sampled tokens, a tokenizer whose pieces hit the newline / CJK / last-space rules); no GPU is involved."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


class _Pieces:
    def __init__(self, pieces):
        self.pieces = pieces

    def decode(self, ids):
        return "".join(self.pieces[int(i) % len(self.pieces)] for i in ids)


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(HERE, "golden", "streaming.json")))


def _stream(tokenizer, tokens):
    from moondream_b200.moondream import MoondreamModel

    holder = type("Holder", (), {"tokenizer": tokenizer})()
    return list(MoondreamModel._stream_text(holder, tokens))


def test_chunks_equal_the_reference_generator(gold):
    tok = _Pieces(gold["pieces"])
    assert len(gold["cases"]) >= 4
    for case in gold["cases"]:
        got = _stream(tok, case["tokens"])
        assert got == case["chunks"], (case["seed"], got[:8], case["chunks"][:8])
        assert "".join(got) == tok.decode(case["tokens"])


def test_flush_rules():
    tok = _Pieces([" ab", "cd", "\n", "漢", " ", "."])
    # nothing is printed before a space proves the word is complete; the tail is flushed at the end
    assert _stream(tok, [0, 1]) == [" ", "abcd"]
    # a newline flushes everything and restarts the cache
    assert _stream(tok, [0, 2, 1]) == [" ", "ab\n", "cd"]
    # a CJK character is printable immediately
    assert _stream(tok, [1, 3, 3]) == ["cd漢", "漢"]
    assert _stream(tok, []) == []


def test_generate_stream_leaves_the_device_context_before_every_yield(monkeypatch):
    """Engine.generate_stream on a CPU shell (decode loop stubbed): chunks arrive in order, the loop stops once every
    sequence has produced eos, the pages go back when the consumer stops early, and the generator is never suspended
    inside `torch.cuda.device(...)` (which would leave the caller's current device switched between chunks)."""
    import contextlib

    import torch

    from moondream_b200 import config as C
    from moondream_b200.engine import Engine, PagePool, PrefixKV, PAGE

    depth = {"now": 0, "entered": 0}

    @contextlib.contextmanager
    def fake_device(_dev):
        depth["now"] += 1
        depth["entered"] += 1
        try:
            yield
        finally:
            depth["now"] -= 1

    monkeypatch.setattr(torch.cuda, "device", fake_device)
    cfg = C.tiny()
    eng = Engine.__new__(Engine)
    eng.cfg, eng.device = cfg, torch.device("cpu")
    eng.pages = PagePool(cfg, 32, "cpu")
    eng.max_blocks = cfg.text.max_context // PAGE
    eos = cfg.tokenizer.eos_id
    rows = torch.tensor([[5, 6, 7, 8, 9, 10, 11, eos, 1, 1, 1, 1, 1, 1, 1, 1, 1],
                         [3, eos, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2]], dtype=torch.int32)
    st = {"preds": torch.zeros((2, 64), dtype=torch.int32), "bt": torch.zeros((2, eng.max_blocks), dtype=torch.int32)}
    queued = []

    def decode_phase(st_, B, pos0, max_tokens, mode, forced, use_graph, stop_on_eos, chunk=0, seed=None):
        assert depth["now"] == 1 and pos0 == [733, 733] and not stop_on_eos
        lo = 0
        for s in range(max_tokens):
            st_["preds"][:, s] = rows[:, s]                         # "decode" one step
            if (s + 1) % chunk == 0:
                queued.append(s + 1)
                yield lo, s + 1
                lo = s + 1
        yield lo, max_tokens + 1

    eng._decode_buffers = lambda B: st
    eng._prefill_phase = lambda *a, **k: None
    eng._decode_phase = decode_phase
    pre = [PrefixKV(730, eng.pages.alloc(12), eng.pages) for _ in range(2)]
    free0 = eng.pages.free_pages
    got = []
    for part in eng.generate_stream(pre, [[1, 2, 3]] * 2, 16, chunk=4):
        assert depth["now"] == 0                                    # the caller runs outside the engine's device context
        assert eng.pages.free_pages < free0                         # the sequences hold their pages while streaming
        got.append(part.clone())
    assert [tuple(p.shape) for p in got] == [(2, 4), (2, 4)] and queued == [4, 8]     # stopped after both rows hit eos
    assert torch.equal(torch.cat(got, 1), rows[:, :8]) and eng.pages.free_pages == free0
    # a consumer that stops early: closing the generator releases the pages
    gen = eng.generate_stream(pre, [[1, 2, 3]] * 2, 16, chunk=4)
    next(gen)
    assert eng.pages.free_pages < free0 and depth["now"] == 0
    gen.close()
    assert eng.pages.free_pages == free0 and depth["now"] == 0
    # no eos at all: the final partial span is delivered and cut at max_tokens
    rows[:] = 4
    got = [p.clone() for p in eng.generate_stream(pre, [[1, 2, 3]] * 2, 6, chunk=4)]
    assert [tuple(p.shape) for p in got] == [(2, 4), (2, 2)]

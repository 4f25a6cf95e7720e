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


def test_generate_reasoning_host_bookkeeping(monkeypatch):
    """Engine.generate_reasoning on a CPU shell (decode loops stubbed): the chain of thought is cut at answer_id, the
    decoded coordinates stay aligned with their tokens, phase 2 prefills the answer prompt at each sequence's OWN
    position (every emitted reasoning token went through the decoder, moondream.py:398), the answer is cut at eos, the
    token budget is halved when the context is short, and the pages are released."""
    import contextlib

    import torch

    from moondream_b200 import config as C
    from moondream_b200.engine import Engine, PagePool, PrefixKV, PAGE

    monkeypatch.setattr(torch.cuda, "device", lambda _d: contextlib.nullcontext())
    cfg = C.tiny()
    tk = cfg.tokenizer
    eng = Engine.__new__(Engine)
    eng.cfg, eng.device = cfg, torch.device("cpu")
    eng.pages = PagePool(cfg, 80, "cpu")
    eng.max_blocks = cfg.text.max_context // PAGE
    st = {"preds": torch.zeros((2, 4096), dtype=torch.int32), "coords": torch.zeros((2, 4096)),
          "bt": torch.zeros((2, eng.max_blocks), dtype=torch.int32)}
    reasoning = [[50, tk.coord_id, tk.coord_id, 51, tk.answer_id, 9, 9, 9, 9],      # 4 reasoning tokens, then answer_id
                 [60, 61, 62, 63, 64, 65, 66, 67, 68]]                               # never says answer_id: runs to max_tokens
    coords = [[0.0, 0.25, 0.75, 0.0, 0.0, 0, 0, 0, 0], [0.0] * 9]
    answers = [[70, 71, tk.eos_id, 5, 5, 5, 5, 5, 5], [80, 81, 82, 83, 84, 85, 86, 87, 88]]
    calls = {"prefill": [], "phase": []}

    def prefill_phase(st_, prompts, start_pos, prompt_embeds, prefix_len, lora=None):
        calls["prefill"].append(([list(p) for p in prompts], list(start_pos), prefix_len))

    def decode_phase(st_, B, pos0, max_tokens, mode, forced, use_graph, stop_on_eos, chunk=0, seed=None):
        calls["phase"].append((list(pos0), max_tokens, mode.reasoning, mode.eos_id, mode.mask_id, mode.mask_id2, seed))
        rows = reasoning if mode.reasoning else answers
        st_["preds"][:, :9] = torch.tensor(rows, dtype=torch.int32)
        if mode.reasoning:
            st_["coords"][:, :9] = torch.tensor(coords)
        yield 0, max_tokens + 1

    eng._decode_buffers = lambda B: st
    eng._prefill_phase = prefill_phase
    eng._decode_phase = decode_phase
    pre = [PrefixKV(730, eng.pages.alloc(12), eng.pages) for _ in range(2)]
    free0 = eng.pages.free_pages
    prompts = [[1, 2, 3], [1, 2, 3, 4, 5]]
    out = eng.generate_reasoning(pre, prompts, [3], 8, temperature=0.5, top_p=0.3, seed=11)
    assert eng.pages.free_pages == free0
    assert out[0] == ([50, tk.coord_id, tk.coord_id, 51], [0.0, 0.25, 0.75, 0.0], [70, 71])
    assert out[1] == ([60, 61, 62, 63, 64, 65, 66, 67], [0.0] * 8, [80, 81, 82, 83, 84, 85, 86, 87])
    assert calls["prefill"][0] == (prompts, [730, 730], -1)
    assert calls["prefill"][1] == ([[3], [3]], [733 + 4, 735 + 8], -1)             # each sequence's own position
    # phase 1: reasoning mode, stops at answer_id, eos and size masked (moondream.py:344,395-396); phase 2: the plain answer mode
    assert calls["phase"][0] == ([733, 735], 8, True, tk.answer_id, tk.eos_id, tk.size_id, 11)
    assert calls["phase"][1] == ([733 + 4 + 1, 735 + 8 + 1], 8, False, tk.eos_id, tk.answer_id, -1, 12)
    # a context that cannot hold 2 x (max_tokens + 1): the budget is split between reasoning and answer
    calls["phase"].clear()
    deep = [PrefixKV(cfg.text.max_context - 30, eng.pages.alloc(eng.max_blocks), eng.pages)]
    st["preds"] = torch.zeros((1, 4096), dtype=torch.int32)
    st["coords"] = torch.zeros((1, 4096))
    reasoning[:] = [reasoning[1]]
    coords[:] = [coords[1]]
    answers[:] = [answers[1]]
    eng.generate_reasoning(deep, [[1, 2, 3]], [3], 100)
    assert calls["phase"][0][1] == calls["phase"][1][1] == (30 - 3 - 1 - 2) // 2
    import pytest

    with pytest.raises(ValueError):
        eng.generate_reasoning([PrefixKV(cfg.text.max_context - 4, [], eng.pages)], [[1, 2]], [3], 10)

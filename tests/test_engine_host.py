"""Host logic of moondream_b200.engine.Engine exercised WITHOUT a GPU: an `Engine` shell (no library, no device) whose
device-side steps are replaced by stand-ins, so the Python bookkeeping around them — streaming spans and device
contexts, reasoning cuts and positions, the lock-step detect / point loop, page accounting — is tested on the CPU.
(The same paths run for real, against the oracle and the reference's fixtures, in tests/test_features_gpu.py.)"""


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


def test_generate_points_lockstep_loop_on_a_cpu_shell(monkeypatch):
    """Engine.generate_points (detect / point for a batch in lock-step, moondream.py:653-733) with the device work
    replaced by stand-ins: a sequence stops contributing objects once its next token is eos (also when that happens at
    the prompt), the others go on up to max_objects, boxes are centre +- size / 2, and the pages are released."""
    import contextlib

    import torch

    from moondream_b200 import _native as N, config as C
    from moondream_b200.engine import Engine, PagePool, PrefixKV, PAGE

    monkeypatch.setattr(torch.cuda, "device", lambda _d: contextlib.nullcontext())
    monkeypatch.setattr(N, "current_stream", lambda: None)
    cfg = C.tiny()
    eos = cfg.tokenizer.eos_id
    eng = Engine.__new__(Engine)
    eng.cfg, eng.device = cfg, torch.device("cpu")
    eng.pages = PagePool(cfg, 64, "cpu")
    eng.max_blocks = cfg.text.max_context // PAGE
    eng.lib = type("L", (), {"md_gather_rows_bf16": lambda self, *a: 0})()
    B = 4
    # per object: next token of each sequence after it; sequence 3 says eos right after the prompt, 1 after its first object
    next_tokens = [[7, eos, 7, 7], [7, 7, eos, 7], [7, 7, 7, 7]]
    first_tokens = [7, 7, 7, eos]
    st = {"bt": torch.zeros((B, eng.max_blocks), dtype=torch.int32), "pos": torch.zeros(B, dtype=torch.int32),
          "pt_h": torch.zeros((B, cfg.text.dim), dtype=torch.bfloat16),
          "pt_xv": torch.zeros((B, 1)), "pt_yv": torch.zeros((B, 1)), "pt_sv": torch.zeros((B, 2)),
          "pt_xb": torch.zeros(B, dtype=torch.int32), "pt_yb": torch.zeros(B, dtype=torch.int32),
          "pt_sb": torch.zeros((B, 2), dtype=torch.int32), "pt_nxt": torch.zeros(B, dtype=torch.int32)}
    replays = []

    class Graph:
        def replay(self):
            k = len(replays)
            replays.append(st["pos"].tolist())
            for b in range(B):
                st["pt_xv"][b, 0], st["pt_yv"][b, 0] = 0.5 + 0.01 * b, 0.25 + 0.1 * k
                st["pt_sv"][b] = torch.tensor([0.2, 0.1])
                st["pt_xb"][b], st["pt_yb"][b] = 100 + b, 200 + k
                st["pt_sb"][b] = torch.tensor([300, 400])
            st["pt_nxt"][:] = torch.tensor(next_tokens[k], dtype=torch.int32)

    st["pt_graphs"] = {True: Graph(), False: Graph()}
    eng._decode_buffers = lambda b: st
    eng._points_buffers = lambda st_, b: None
    eng.embed = lambda ids, out, **k: None
    prefills = []
    eng.prefill = lambda x, q_off, start, bt, **k: prefills.append((int(x.shape[0]), list(q_off), list(start)))

    def lm_head(hidden, out_ids, stride, **k):
        out_ids[:] = torch.tensor(first_tokens, dtype=torch.int32)

    eng.lm_head = lm_head
    pre = [PrefixKV(730, eng.pages.alloc(12), eng.pages) for _ in range(B)]
    free0 = eng.pages.free_pages
    prompts = [[1, 2, 3, 4]] * 2 + [[1, 2, 3, 4, 5, 6]] * 2
    res = eng.generate_points(pre, prompts, include_size=True, max_objects=3)
    assert eng.pages.free_pages == free0 and prefills == [(20, [0, 4, 8, 14, 20], [730] * 4)]
    assert [len(r) for r in res] == [3, 1, 2, 0] and len(replays) == 3
    assert replays[0] == [734, 734, 736, 736]                       # positions after the ragged prompt prefill
    box = res[2][1]                                                 # sequence 2, its second object
    assert box["bins"] == [102, 201, 300, 400]
    for key, want in (("x_min", 0.52 - 0.1), ("x_max", 0.52 + 0.1), ("y_min", 0.35 - 0.05), ("y_max", 0.35 + 0.05)):
        assert abs(box[key] - want) < 1e-6
    # points: same loop without the size step; max_objects caps a sequence that never says eos
    replays.clear()
    pts = eng.generate_points(pre, prompts, include_size=False, max_objects=2)
    assert [len(r) for r in pts] == [2, 1, 2, 0] and len(replays) == 2 and set(pts[0][0]) == {"x", "y", "bins"}
    # every sequence done at the prompt: no object step runs at all
    replays.clear()
    first_tokens[:] = [eos] * B
    assert eng.generate_points(pre, prompts, include_size=True, max_objects=5) == [[], [], [], []] and not replays
    assert eng.pages.free_pages == free0


def test_generate_returns_every_page_when_the_decode_loop_fails(monkeypatch):
    """Engine.generate: whatever happens after the block tables were built — here the decode loop raises — the pages of
    the batch go back to the pool exactly once, including consumed prefix pages (whose handles stay marked as handed over);
    option conflicts are rejected; with `prefilled_hidden` the prompt is not prefilled again."""
    import contextlib

    import pytest
    import torch

    from moondream_b200 import config as C
    from moondream_b200.engine import Engine, PagePool, PrefixKV, PAGE

    monkeypatch.setattr(torch.cuda, "device", lambda _d: contextlib.nullcontext())
    cfg = C.tiny()
    eng = Engine.__new__(Engine)
    eng.cfg, eng.device = cfg, torch.device("cpu")
    eng.pages = PagePool(cfg, 40, "cpu")
    eng.max_blocks = cfg.text.max_context // PAGE
    st = {"S": cfg.text.max_context + 1, "bt": torch.zeros((2, eng.max_blocks), dtype=torch.int32),
          "x": torch.zeros((2, cfg.text.dim), dtype=torch.bfloat16),
          "preds": torch.arange(2 * 64, dtype=torch.int32).view(2, 64), "margins": torch.zeros((2, 64))}
    eng._decode_buffers = lambda B: st
    prefills = []
    eng._prefill_phase = lambda st_, prompts, start, emb, prefix_len, lora=None: prefills.append(list(start))
    boom = {"on": True}

    def decode_phase(st_, B, pos0, max_tokens, mode, forced, use_graph, stop_on_eos, seed=None):
        if boom["on"]:
            raise RuntimeError("device fault")
        st_["steps_run"] = max_tokens
        st_["pos0"] = list(pos0)
        yield 0, max_tokens + 1

    eng._decode_phase = decode_phase
    total = eng.pages.free_pages
    for consume in (False, True):
        pre = [PrefixKV(730, eng.pages.alloc(12), eng.pages) for _ in range(2)]
        with pytest.raises(RuntimeError, match="device fault"):
            eng.generate(pre, [[1, 2, 3]] * 2, 100, consume=consume)
        if consume:
            assert all(p._released for p in pre) and eng.pages.free_pages == total       # handed over, then returned
        else:
            assert eng.pages.free_pages == total - 24 and not any(p._released for p in pre)
        for p in pre:
            p.release()                                                                   # a no-op for consumed prefixes
        assert eng.pages.free_pages == total and sorted(eng.pages._free) == list(range(40))   # nothing returned twice
    boom["on"] = False
    pre = [PrefixKV(762, eng.pages.alloc(12), eng.pages) for _ in range(2)]                # prefixes that include the prompt
    prefills.clear()
    res = eng.generate(pre, [[1, 2, 3]] * 2, 6, consume=True, prefilled_hidden=torch.ones((2, cfg.text.dim), dtype=torch.bfloat16))
    assert prefills == [] and st["pos0"] == [762, 762] and float(st["x"].float().min()) == 1.0
    assert tuple(res.tokens.shape) == (2, 7) and res.tokens[1].tolist() == list(range(64, 71)) and res.steps == 6
    assert eng.pages.free_pages == total
    pre = [PrefixKV(730, eng.pages.alloc(12), eng.pages) for _ in range(2)]
    with pytest.raises(ValueError, match="mutually exclusive"):
        eng.generate(pre, [[1]] * 2, 4, forced=[[1] * 5] * 2, sampler=lambda logits: logits)
    assert eng.pages.free_pages == total - 24


def test_seam_prefill_tells_the_prefix_lm_mask_from_the_causal_one(monkeypatch):
    """SeamAdapter._prefill receives the reference's mask tensor and must hand the native prefill the matching rule:
    prefix-LM (moondream.py:138-146: the first 730 positions see each other) or plain causal (text-only query,
    :571-574).  Host logic only: the engine is a recorder."""
    import contextlib

    import torch

    from moondream_b200 import config as C
    from moondream_b200.seam import SeamAdapter

    monkeypatch.setattr(torch.cuda, "device", lambda _d: contextlib.nullcontext())
    cfg = C.tiny()
    t = cfg.text
    calls = []
    eng = type("E", (), {"cfg": cfg, "device": torch.device("cpu"),
                         "prefill": lambda self, h, q_off, start, bt, prefix_len=-1: calls.append((list(q_off), list(start), prefix_len))})()
    seam = SeamAdapter(eng)
    seam._table = lambda: None
    ctx = t.max_context
    prefix_lm = torch.tril(torch.ones(1, 1, ctx, ctx, dtype=torch.bool))
    prefix_lm[..., : t.prefix_attn, : t.prefix_attn] = True                        # moondream.py:143-145
    causal = torch.tril(torch.ones(1, 1, ctx, ctx, dtype=torch.bool))              # moondream.py:571-574

    def run(mask, pos, T):
        calls.clear()
        x = torch.zeros((1, T, t.dim), dtype=torch.bfloat16)
        out = seam._prefill(x, mask[:, :, pos: pos + T, :], torch.arange(pos, pos + T), None)
        assert tuple(out.shape) == (1, T, t.dim) and calls[0][:2] == ([0, T], [pos])
        return calls[0][2]

    assert run(prefix_lm, 0, t.prefix_attn) == -1          # encode_image: [BOS | image] under the prefix-LM mask
    assert run(causal, 0, 9) == 0                          # text-only query: BOS + prompt at position 0, causal
    assert run(prefix_lm, t.prefix_attn, 12) == -1         # a prompt after the image prefix: the masks coincide there
    assert run(causal, 40, 5) == 0
    assert run(prefix_lm, 0, t.prefix_attn + 32) == -1     # one fused pass over [BOS | image | prompt]
    assert run(causal, 0, t.prefix_attn + 32) == 0
    assert run(prefix_lm, 700, 100) == -1 and run(causal, 700, 100) == 0           # starts inside the prefix, ends past it
    assert run(prefix_lm, t.prefix_attn - 1, 4) == -1      # only the prefix's last row: nothing to tell apart, default rule

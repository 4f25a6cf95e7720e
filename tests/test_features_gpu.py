"""GPU: the round-2 rows of SURVEY.md section 8 against the oracle and the reference-generated fixtures —
grounded reasoning on the device (a22, moondream.py:323-432), text-only query (:565-574), grouped-query decoder
(a14, text.py:36-38,49), on-device top-p sampling inside the decode graph (a26 / f2, :270-278), streaming that yields
while decoding (a21, :470-537).

Tolerances: token ids / kept sets exact, except at decisions where the ORACLE's own margin is below 4.5 bf16 ulps
(see tests/test_model_parity_gpu.py); decoded coordinates exact (bin / 1024) under the same rule; probabilities of
the kept set within one bf16 ulp."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
NEAR_TIE_ULPS = 4.5


def _gold(name):
    return json.load(open(os.path.join(HERE, "golden", name)))


def _model(cfg, sd, max_batch=4):
    from moondream_b200.moondream import MoondreamModel
    from oracle.reference_shim import StubTokenizer

    m = MoondreamModel(cfg, tokenizer=StubTokenizer(cfg.text.vocab_size), max_batch=max_batch)
    m.load_state_dict(sd)
    return m


def _ids(text):
    return [int(t) for t in text.split()]


def _agree(got, want, ulps, what):
    """exact, or first difference at an oracle near-tie (after which the sequences may legitimately diverge)"""
    for i, (a, b) in enumerate(zip(got, want)):
        if a != b:
            assert ulps[i] < NEAR_TIE_ULPS, (what, i, a, b, ulps[i])
            return i
    return None


@pytest.fixture(scope="module")
def tiny():
    from moondream_b200 import config as C, synth

    cfg = C.tiny()
    return cfg, synth.synthetic_state_dict(cfg, 0)


# ------------------------------------------------------------------------------------------------ reasoning
def test_reasoning_matches_the_reference_golden(tiny):
    from moondream_b200 import synth

    cfg, sd = tiny
    gold = _gold("tiny_reasoning.json")
    sd = dict(sd)
    sd["text.lm_head.bias"] = synth.special_token_bias(sd, cfg, *gold["bias"])
    model = _model(cfg, sd)
    tk = cfg.tokenizer
    strict = 0
    for c in gold["cases"]:
        img = synth.synthetic_image(c["image_index"], c["height"], c["width"])
        out = model.query(img, c["question"], reasoning=True, settings={"temperature": 0, "max_tokens": c["max_tokens"]})
        assert set(out) == {"reasoning", "answer"} and set(out["reasoning"]) == {"text", "grounding"}
        got = _ids(out["reasoning"]["text"])
        want = c["reasoning_tokens"]
        div = _agree(got, want, c["margin_ulps"], "reasoning tokens")
        # coordinates: every coord token before the first divergence carries the reference's value
        coord_ok = True
        if div is None and len(got) == len(want):
            if len(want) == c["max_tokens"] or c["end_margin_ulps"] >= NEAR_TIE_ULPS:
                gpts = [p for g in out["reasoning"]["grounding"] for p in g["points"]]
                wpts = [p for g in c["grounding"] for p in g["points"]]
                near = [u for u in c["coord_ulps"] if u is not None and u < NEAR_TIE_ULPS]
                if not near:
                    assert len(gpts) == len(wpts)
                    for gp, wp in zip(gpts, wpts):
                        assert abs(gp[0] - wp[0]) < 1e-6 and abs(gp[1] - wp[1]) < 1e-6, (gp, wp)
                    assert [(g["start_idx"], g["end_idx"]) for g in out["reasoning"]["grounding"]] == \
                           [(g["start_idx"], g["end_idx"]) for g in c["grounding"]]
                    adiv = _agree(_ids(out["answer"]), c["answer_tokens"], c["answer_margin_ulps"], "answer tokens")
                    strict += int(adiv is None and out["reasoning"]["text"] == c["reasoning_text"])
                else:
                    coord_ok = False
    assert strict >= 2, f"only {strict}/{len(gold['cases'])} reasoning cases reproduced the reference strictly"


def test_reasoning_batch_and_engine_level_outputs(tiny):
    """generate_reasoning for a batch of 3 (lock-step, each row ends at its own answer_id) against the oracle."""
    from moondream_b200 import synth
    from moondream_b200.engine import Engine
    from oracle.moondream_oracle import OracleModel

    cfg, sd = tiny
    gold = _gold("tiny_reasoning.json")
    sd = dict(sd)
    sd["text.lm_head.bias"] = synth.special_token_bias(sd, cfg, *gold["bias"])
    eng = Engine(cfg, sd, max_batch=4)
    orc = OracleModel(cfg, sd)
    cases = gold["cases"][:3]
    imgs = [synth.synthetic_image(c["image_index"], c["height"], c["width"]) for c in cases]
    res = eng.generate_reasoning(eng.encode_images(imgs), [c["prompt"] for c in cases],
                                 cfg.tokenizer.templates["query"]["suffix"], 24)
    for (r_toks, coords, a_toks), c in zip(res, cases):
        div = _agree(r_toks, c["reasoning_tokens"], c["margin_ulps"], "batched reasoning")
        if div is None and len(r_toks) == len(c["reasoning_tokens"]):
            for t, got, want, u in zip(r_toks, coords, c["coords"], c["coord_ulps"]):
                if t == cfg.tokenizer.coord_id and u >= NEAR_TIE_ULPS:
                    assert abs(got - want) < 1e-6, (got, want)
    del eng
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ text-only
def test_text_only_query_matches_the_reference_golden(tiny):
    cfg, sd = tiny
    gold = _gold("tiny_text_only.json")
    model = _model(cfg, sd)
    exact = 0
    for c in gold["cases"]:
        assert model._query_prompt(c["question"], None, with_bos=True) == c["prompt"]
        out = model.query(None, c["question"], settings={"temperature": 0, "max_tokens": c["max_tokens"]})
        got = _ids(out["answer"])
        exact += int(_agree(got, c["tokens"], c["margin_ulps"], "text-only") is None and got == c["tokens"])
    assert exact >= 2, exact
    chunks = list(model.query(None, gold["cases"][0]["question"], stream=True,
                              settings={"temperature": 0, "max_tokens": 16})["answer"])
    assert "".join(chunks) == model.query(None, gold["cases"][0]["question"],
                                          settings={"temperature": 0, "max_tokens": 16})["answer"]


# ------------------------------------------------------------------------------------------------ GQA
def test_grouped_query_decoder_against_reference_golden_and_oracle():
    from moondream_b200 import config as C, synth
    from moondream_b200.engine import Engine
    from oracle.moondream_oracle import OracleModel

    cfg = C.tiny_gqa()
    sd = synth.synthetic_state_dict(cfg, 0)
    eng = Engine(cfg, sd, max_batch=4)
    orc = OracleModel(cfg, sd)
    gold = _gold("tiny_gqa.json")
    imgs = [synth.synthetic_image(c["image_index"], c["height"], c["width"]) for c in gold["cases"]]
    prefixes, feats, img_emb, hidden = eng.encode_images(imgs, return_hidden=True)
    for i, c in enumerate(gold["cases"]):
        o_enc = orc.encode_image(imgs[i])
        kv = eng.prefix_kv_tensors(prefixes[i])
        assert tuple(kv[0][0].shape) == (1, 2, 730, 64)
        for (k, v), (ok, ov) in zip(kv, o_enc.caches):
            rk = ((k.float().cpu() - ok.float()).norm() / ok.float().norm()).item()
            rv = ((v.float().cpu() - ov.float()).norm() / ov.float().norm()).item()
            assert rk < 3e-2 and rv < 3e-2, (rk, rv)
    res = eng.generate(prefixes, [c["prompt"] for c in gold["cases"]], 16)
    for i, c in enumerate(gold["cases"]):
        _agree(res.tokens[i, :16].tolist(), c["tokens"], c["margin_ulps"], f"gqa image {i}")
    # teacher-forced: every decode step (non-fused decode path) against the oracle at clear margins
    forced = [synth.synthetic_prompt(70 + i, 25, cfg.text.vocab_size) for i in range(2)]
    tf = eng.generate(eng.encode_images(imgs), [c["prompt"] for c in gold["cases"]], 24, forced=forced)
    checked = agree = 0
    for i, c in enumerate(gold["cases"]):
        o = orc.generate(orc.encode_image(imgs[i]), c["prompt"], 24, forced=forced[i])
        for s in range(24):
            if o.margin_ulps[s] >= NEAR_TIE_ULPS:
                checked += 1
                agree += int(tf.tokens[i, s].item() == o.predicted[s])
    assert checked > 20 and agree == checked, (agree, checked)
    del eng
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ sampling
def _oracle_next_probs(logits_bf16_cpu, temperature, top_p):
    from moondream_b200.sampling import apply_top_p

    return apply_top_p(torch.softmax(logits_bf16_cpu / temperature, dim=-1), top_p)


def _boundary_is_clear(logits_bf16_cpu, temperature, top_p):
    """the reference's own keep / drop decision at the edge of the nucleus is at least 2 bf16 ulps of top_p away from
    flipping (its cumulative sums are rounded to bf16, so closer calls depend on summation order)"""
    p = torch.softmax(logits_bf16_cpu / temperature, dim=-1)
    srt, _ = torch.sort(p, dim=-1, descending=True)
    before = (torch.cumsum(srt, dim=-1) - srt).float()
    tp = torch.tensor(top_p, dtype=torch.bfloat16).float()
    ulp = 2.0 ** (np.floor(np.log2(max(float(tp), 1e-30))) - 7)
    return bool(((before - tp).abs() >= 2 * ulp).all(dim=-1).all())


def _nucleus_agrees(p_soft, kept_dev, kept_orc):
    """Same nucleus up to ties at its edge: the reference's kept set is a prefix of torch.sort's order, which is not
    stable for equal probabilities (bf16 probabilities tie by the dozen), so WHICH of the tokens that share the
    smallest kept probability survive is an artefact of that sort; this kernel keeps the lowest indices.  Required:
    equal counts, and a symmetric difference confined to tokens within one bf16 ulp of the smallest kept probability."""
    pmin = p_soft[kept_orc].min()
    band = (p_soft - pmin).abs() <= pmin * 2.0 ** -7
    diff = kept_dev != kept_orc
    return int(kept_dev.sum()) == int(kept_orc.sum()) and bool((~diff | band).all())


def test_device_top_p_kept_set_matches_apply_top_p(tiny):
    """md_sample_top_p's `next_probs` against the reference's _apply_top_p: golden fixture (tiny vocab) and random
    51200-wide rows against the restatement the fixture pins; the draw is the inverse CDF of those probabilities."""
    from moondream_b200.engine import Engine

    cfg, sd = tiny
    eng = Engine(cfg, sd, max_batch=4)
    gold = _gold("top_p.json")
    for c in gold["cases"]:
        logits = torch.tensor(c["logits"]).to(torch.bfloat16).unsqueeze(0)
        out = torch.zeros((1, 1), dtype=torch.int32, device="cuda")
        u = torch.tensor([0.5], dtype=torch.float32, device="cuda")
        probs = eng.sample_tokens(logits.cuda(), c["temperature"], c["top_p"], out, 1, uniforms=u, keep_probs=True)
        probs = probs[0].float().cpu()
        if _boundary_is_clear(logits, c["temperature"], c["top_p"]):
            p_soft = torch.softmax(logits / c["temperature"], dim=-1)[0].float()
            want = torch.zeros_like(p_soft)
            want[c["kept_ids"]] = torch.tensor(c["kept_probs"])
            assert _nucleus_agrees(p_soft, probs > 0, want > 0), (c["temperature"], c["top_p"])
            got_sorted = probs[probs > 0].sort(descending=True).values
            want_sorted = want[want > 0].sort(descending=True).values
            assert ((got_sorted - want_sorted).abs() <= want_sorted * 2.0 ** -6).all()
    checked = 0
    g = torch.Generator().manual_seed(0)
    for temp, top_p, scale in ((0.5, 0.3, 3.0), (1.0, 0.9, 2.0), (1.5, 0.95, 1.0), (0.7, 0.5, 4.0), (1.0, 1.0, 1.0)):
        logits = (torch.randn(8, 51200, generator=g) * scale).to(torch.bfloat16)
        logits[:, 3] = float("-inf")
        u_host = torch.rand(8, generator=g)
        out = torch.zeros((8, 1), dtype=torch.int32, device="cuda")
        probs = eng.sample_tokens(logits.cuda(), temp, top_p, out, 1, uniforms=u_host.cuda(), keep_probs=True)
        probs = probs.float().cpu()
        toks = out.flatten().cpu()
        for b in range(8):
            row = logits[b: b + 1]
            if not _boundary_is_clear(row, temp, top_p):
                continue
            p_soft = torch.softmax(row / temp, dim=-1)[0].float()
            want = _oracle_next_probs(row, temp, top_p)[0].float()
            assert _nucleus_agrees(p_soft, probs[b] > 0, want > 0), (temp, top_p, b, int((probs[b] > 0).sum()), int((want > 0).sum()))
            got_sorted = probs[b][probs[b] > 0].sort(descending=True).values
            want_sorted = want[want > 0].sort(descending=True).values
            assert ((got_sorted - want_sorted).abs() <= want_sorted * 2.0 ** -6 + 1e-12).all()
            assert probs[b, 3] == 0
            # the draw: inverse CDF, in index order, of the probabilities the kernel itself reports, same uniform
            cdf = torch.cumsum(probs[b].double(), 0)
            target = float(u_host[b]) * float(cdf[-1])
            idx = min(int(torch.searchsorted(cdf, torch.tensor(target, dtype=torch.float64), right=True)), 51199)
            lo = float(cdf[idx - 1]) if idx > 0 else 0.0
            if min(abs(target - lo), abs(float(cdf[idx]) - target)) < 1e-4 * float(cdf[-1]):
                continue                                     # the uniform fell on a bin edge up to fp32 summation order
            assert int(toks[b]) == idx, (temp, top_p, b, int(toks[b]), idx)
            checked += 1
    assert checked >= 10, checked
    del eng
    torch.cuda.empty_cache()


def test_device_sampling_statistics_and_seeding(tiny):
    """Philox draws: frequencies follow the kept probabilities (chi-square), a seed reproduces itself, different
    steps / rows draw different numbers, and top_p -> 0 degenerates to the argmax."""
    from moondream_b200.engine import Engine

    cfg, sd = tiny
    eng = Engine(cfg, sd, max_batch=4)
    g = torch.Generator().manual_seed(1)
    logits = (torch.randn(1, 2048, generator=g) * 2.5).to(torch.bfloat16)
    p_soft = torch.softmax(logits, dim=-1)[0].float()
    ref = _oracle_next_probs(logits, 1.0, 0.8)[0].float()
    one = torch.zeros((1, 1), dtype=torch.int32, device="cuda")
    want = eng.sample_tokens(logits.cuda(), 1.0, 0.8, one, 1, uniforms=torch.tensor([0.5], device="cuda"),
                             keep_probs=True)[0].float().cpu()
    assert _nucleus_agrees(p_soft, want > 0, ref > 0)
    want = want / want.sum()
    R = 4096
    rows = logits.repeat(R, 1).cuda()
    out = torch.zeros((R, 4), dtype=torch.int32, device="cuda")
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    step = torch.tensor([2], dtype=torch.int32, device="cuda")
    eng.sample_tokens(rows, 1.0, 0.8, out, 4, out_offset=1, step=step, seed=seed)
    toks = out[:, 3].cpu().long()
    assert bool((want[toks] > 0).all()), "a token outside the nucleus was drawn"
    counts = torch.bincount(toks, minlength=2048).float()
    keep = want * R >= 5
    chi2 = float((((counts - want * R) ** 2) / (want * R))[keep].sum())
    dof = int(keep.sum()) - 1
    assert chi2 < dof + 6 * (2 * dof) ** 0.5, (chi2, dof)
    out2 = torch.zeros_like(out)
    eng.sample_tokens(rows, 1.0, 0.8, out2, 4, out_offset=1, step=step, seed=seed)
    assert torch.equal(out, out2)
    step.fill_(1)
    out3 = torch.zeros((R, 4), dtype=torch.int32, device="cuda")
    eng.sample_tokens(rows, 1.0, 0.8, out3, 4, out_offset=1, step=step, seed=seed)
    assert not torch.equal(out3[:, 2], out[:, 3])
    out4 = torch.zeros((R, 1), dtype=torch.int32, device="cuda")
    eng.sample_tokens(rows, 1.0, 1e-6, out4, 1, seed=seed)
    assert bool((out4.flatten().cpu() == int(torch.argmax(logits[0].float()))).all())
    del eng
    torch.cuda.empty_cache()


def test_sampled_generation_in_the_graph(tiny):
    """caption with the reference's default settings (temperature 0.5, top_p 0.3) runs inside the CUDA graph: tokens
    stay in the nucleus of the ORACLE's distribution when the oracle is teacher-forced along them; seeds reproduce."""
    from moondream_b200 import synth
    from moondream_b200.engine import Engine
    from oracle.moondream_oracle import OracleModel

    cfg, sd = tiny
    eng = Engine(cfg, sd, max_batch=4)
    orc = OracleModel(cfg, sd)
    imgs = [synth.synthetic_image(i, 378, 378) for i in range(3)]
    prompts = [synth.synthetic_prompt(i, 6, cfg.text.vocab_size) for i in range(3)]
    a = eng.generate(eng.encode_images(imgs), prompts, 12, temperature=1.0, top_p=0.9, seed=5, stop_on_eos=False).tokens
    b = eng.generate(eng.encode_images(imgs), prompts, 12, temperature=1.0, top_p=0.9, seed=5, stop_on_eos=False).tokens
    c = eng.generate(eng.encode_images(imgs), prompts, 12, temperature=1.0, top_p=0.9, seed=6, stop_on_eos=False).tokens
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert len({tuple(a[i].tolist()) for i in range(3)}) == 3
    inside = total = 0
    for i in range(3):
        enc = orc.encode_image(imgs[i])
        orc.load_encoded(enc)
        logits, _, _, pos = orc.prefill_prompt(prompts[i], enc.pos)
        for s in range(12):
            if s > 0:
                logits[:, cfg.tokenizer.answer_id] = float("-inf")
            kept = _oracle_next_probs(logits, 1.0, 0.9)[0]
            tok = int(a[i, s])
            total += 1
            inside += int(kept[tok] > 0)
            logits, _ = orc.decode_one(orc.embed(torch.tensor([[tok]])), pos)
            pos += 1
    assert inside >= total - 2, (inside, total)        # the engine's bf16 logits differ from the oracle's in the last ulp
    # temperature -> deterministic limit reproduces greedy
    greedy = eng.generate(eng.encode_images(imgs), prompts, 8, stop_on_eos=False)
    cold = eng.generate(eng.encode_images(imgs), prompts, 8, temperature=1.0, top_p=1e-6, seed=1, stop_on_eos=False)
    for i in range(3):
        for s in range(8):
            if int(greedy.tokens[i, s]) != int(cold.tokens[i, s]):
                assert greedy.margins[i, s].item() < 0.3          # an exact bf16 tie of the top two logits
                break
    del eng
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ streaming
def test_streaming_yields_while_decoding(tiny):
    from moondream_b200 import synth

    cfg, sd = tiny
    model = _model(cfg, sd)
    img = synth.synthetic_image(5, 378, 378)
    enc = model.encode_image(img)
    settings = {"temperature": 0, "max_tokens": 40}
    full = model.caption(enc, "normal", settings=settings)["caption"]
    gen = model.caption(enc, "normal", stream=True, settings=settings)["caption"]
    first = next(gen)                       # arrives after the first 8 graph replays, not after all 40
    state = model.engine._decode_state[1]
    assert int(state["step"].item()) <= 16, "the first chunk must not wait for the whole generation"
    rest = "".join(gen)
    assert first + rest == full
    # a consumer that stops early frees the sequence's pages
    free0 = model.engine.pages.free_pages
    gen = model.caption(enc, "normal", stream=True, settings=settings)["caption"]
    next(gen)
    gen.close()
    assert model.engine.pages.free_pages == free0
    # engine level, batch of 2: chunks concatenate to the non-streamed tokens
    imgs = [synth.synthetic_image(i, 378, 378) for i in range(2)]
    prompts = [synth.synthetic_prompt(i, 5, cfg.text.vocab_size) for i in range(2)]
    eng = model.engine
    parts = list(eng.generate_stream(eng.encode_images(imgs), prompts, 20, chunk=6))
    cat = torch.cat(parts, dim=1)
    ref = eng.generate(eng.encode_images(imgs), prompts, 20, stop_on_eos=False).tokens[:, :20]
    assert cat.shape[1] == 20 and torch.equal(cat, ref)


# ------------------------------------------------------------------------------------------------ the reference's seam
def test_seam_adapter_drives_the_reference_flow(tiny):
    """encode_image (moondream.py:206-268) and the greedy loop of _generate_answer (:434-539) written the way the
    REFERENCE writes them — prepare_crops, reconstruct_from_crops, masks and pos_ids as torch tensors — with the four
    seam methods bound to SeamAdapter, against the plain oracle."""
    from moondream_b200 import synth
    from moondream_b200.engine import Engine
    from moondream_b200.seam import SeamAdapter
    from oracle.moondream_oracle import OracleModel, stitch_crops

    cfg, sd = tiny
    eng = Engine(cfg, sd, max_batch=2)
    seam = SeamAdapter(eng)
    orc = OracleModel(cfg, sd)
    dev = eng.device
    v, t, tk = cfg.vision, cfg.text, cfg.tokenizer
    wte = sd["text.wte"].to(dev)
    attn_mask = orc.attn_mask.to(dev)
    for idx, (h, w) in enumerate([(500, 700), (378, 378)]):
        img = synth.synthetic_image(20 + idx, h, w)
        crops, tiling = orc.prepare_crops(img)                                   # vision.py:25-41 (host)
        feats = seam._vis_enc(crops.to(dev))                                     # moondream.py:211
        o_feats = orc.vision_encoder(crops)
        assert ((feats.float().cpu() - o_feats.float()).norm() / o_feats.float().norm()).item() < 3e-2
        g = v.crop_size // v.enc_patch_size
        recon = stitch_crops(feats[1:].view(-1, g, g, v.enc_dim), tiling, v.overlap_margin)   # image_crops.py:170-231
        img_emb = seam._vis_proj(feats[0], recon)                                # moondream.py:228
        o_enc, o_emb, _ = orc.encode_image(img, return_embeds=True)
        assert ((img_emb.float().cpu() - o_emb.float()).norm() / o_emb.float().norm()).item() < 3e-2
        bos = wte[torch.tensor([[tk.bos_id]], device=dev)]
        x = torch.cat([bos, img_emb[None]], dim=1)                               # moondream.py:250-254
        n = x.size(1)
        seam._prefill(x, attn_mask[:, :, 0:n, :], torch.arange(n, device=dev), None)
        prompt = synth.synthetic_prompt(idx, 5, t.vocab_size)
        pe = wte[torch.tensor([prompt], device=dev)]
        hidden = seam._prefill(pe, attn_mask[:, :, n:n + 5, :], torch.arange(n, n + 5, device=dev), None)
        assert hidden.shape == (1, 5, t.dim)
        gen = orc.generate(o_enc, prompt, 10)
        # first token: lm_head of the last prompt row == what _decode_one_tok's logits path computes; take it from a
        # zero-length trick instead: feed the oracle's first token and compare the following ones
        pos = n + 5
        mask = torch.zeros(1, 1, t.max_context, device=dev, dtype=torch.bool)
        mask[:, :, :pos] = 1
        tok = gen.tokens[0]
        got = [tok]
        for s in range(1, 8):
            mask[:, :, pos] = 1
            logits, hid = seam._decode_one_tok(wte[torch.tensor([[tok]], device=dev)], mask,
                                               torch.tensor([pos], device=dev), None)
            assert logits.shape == (1, t.vocab_size) and hid.shape == (1, 1, t.dim)
            logits[:, tk.answer_id] = float("-inf")
            tok = int(torch.argmax(logits, dim=-1).item())
            got.append(tok)
            pos += 1
        _agree(got, gen.tokens, gen.margin_ulps, f"seam image {idx}")
    # load_encoded_image through the adapter: the oracle's KV prefix copied into the pages gives the same continuation
    seam.load_kv_prefix([(k.to(dev), vv.to(dev)) for k, vv in o_enc.caches])
    hidden = seam._prefill(pe, attn_mask[:, :, n:n + 5, :], torch.arange(n, n + 5, device=dev), None)
    orc.load_encoded(o_enc)
    o_hidden = orc.prefill_prompt(prompt, o_enc.pos)[1]
    assert ((hidden.float().cpu() - o_hidden.float()).norm() / o_hidden.float().norm()).item() < 3e-2
    seam.release()
    del eng
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ preprocessing
def test_device_preprocessing_is_bit_exact_with_pil(tiny):
    """overlap_crop_image on the device (resize with Pillow's fixed-point Lanczos tables + window extraction) against
    the host path (PIL itself) and the crop hashes the REFERENCE produced (tests/golden/crops.json)."""
    import hashlib

    from moondream_b200 import synth
    from moondream_b200.engine import Engine
    from moondream_b200.image_crops import overlap_crop_image

    cfg, sd = tiny
    eng = Engine(cfg, sd, max_batch=4)
    gold = _gold("crops.json")["cases"]
    imgs = [synth.synthetic_image(c["image_index"], c["height"], c["width"]) for c in gold]
    for lo in range(0, len(imgs), 4):
        batch = imgs[lo: lo + 4]
        crops, offs, til = eng.stage_images(batch, preprocess="device")
        host, offs_h, til_h = eng.stage_images(batch, preprocess="host")
        assert offs == offs_h and til == til_h
        crops, host = crops.cpu().numpy().copy(), host.cpu().numpy()
        assert np.array_equal(crops, host)
        for j, c in enumerate(gold[lo: lo + 4]):
            mine = crops[offs[j]: offs[j + 1]]
            assert list(til[j]) == c["tiling"] and mine.shape[0] == c["n_crops"]
            assert hashlib.sha256(mine.tobytes()).hexdigest()[:16] == c["sha256"], (c["height"], c["width"])
    # single resize entry point against PIL
    from PIL import Image
    img = synth.synthetic_image(77, 480, 640)
    got = eng.resize_lanczos(torch.from_numpy(img).cuda(), 378, 378).cpu().numpy()
    want = np.asarray(Image.fromarray(img).resize((378, 378), resample=Image.Resampling.LANCZOS))
    assert np.array_equal(got, want)
    del eng
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ LoRA variants
def test_lora_variant_matches_the_reference_golden(tiny, tmp_path, monkeypatch):
    """settings["variant"] (lora.py:55-79; text.py:31-32,54-56; layers.py:131-143): a synthetic rank-8 variant placed in
    the reference's cache layout, through the public API, against the tokens / boxes the UNMODIFIED reference produced
    with the same file (tests/golden/tiny_lora.json) and the oracle's KV prefix."""
    from moondream_b200 import synth
    from oracle.moondream_oracle import OracleModel

    cfg, sd = tiny
    gold = _gold("tiny_lora.json")
    flat = synth.synthetic_lora(cfg, gold["rank"], gold["seed"])
    vdir = tmp_path / "md_variants" / "synthetic-r8"
    vdir.mkdir(parents=True)
    torch.save(flat, vdir / "final.pt")
    monkeypatch.setenv("HF_HUB_CACHE", str(tmp_path))
    model = _model(cfg, sd)
    orc = OracleModel(cfg, sd)
    orc.lora = synth.nest_lora(flat)
    settings = {"temperature": 0, "max_tokens": 12, "variant": "synthetic-r8"}
    base = {"temperature": 0, "max_tokens": 12}
    differs = 0
    for c in gold["cases"]:
        img = synth.synthetic_image(c["image_index"], c["height"], c["width"])
        enc = model.encode_image(img, settings)
        o_enc = orc.encode_image(img)
        for (k, v), (ok, ov) in zip(enc.caches, o_enc.caches):
            assert ((k.float().cpu() - ok.float()).norm() / ok.float().norm()).item() < 3e-2
            assert ((v.float().cpu() - ov.float()).norm() / ov.float().norm()).item() < 3e-2
        res = model.engine.generate([enc._prefix], [c["prompt"]], 12, lora=model._lora(settings))
        _agree(res.tokens[0, :12].tolist(), c["tokens"], c["margin_ulps"], "lora tokens")
        plain = model.engine.generate([model.encode_image(img)._prefix], [c["prompt"]], 12)
        differs += int(plain.tokens[0, :12].tolist() != res.tokens[0, :12].tolist())
        det = model.detect(enc, "17 23", settings={"max_objects": 2, "variant": "synthetic-r8"})["objects"]
        near = any(u < NEAR_TIE_ULPS for o in c["detect_ulps"] for u in o)
        if not near:
            assert len(det) == len(c["detect_boxes"])
            for got, want in zip(det, c["detect_boxes"]):
                for key in ("x_min", "y_min", "x_max", "y_max"):
                    assert abs(got[key] - want[key]) < 1e-5
    assert differs == len(gold["cases"]), "the adapters must change the output"
    # API level: caption under the variant, sampling under the variant, unknown variants are not downloaded
    out = model.caption(synth.synthetic_image(0, 378, 378), "short", settings=settings)
    assert isinstance(out["caption"], str) and out["caption"] != model.caption(synth.synthetic_image(0, 378, 378), "short", settings=base)["caption"]
    out = model.query(synth.synthetic_image(0, 378, 378), "11 12", settings={"temperature": 0.7, "max_tokens": 6, "variant": "synthetic-r8"})
    assert len(_ids(out["answer"])) <= 6
    with pytest.raises(RuntimeError):
        model.caption(synth.synthetic_image(0, 378, 378), "short", settings={"temperature": 0, "variant": "missing"})


def test_native_loader_reads_straight_into_device_memory(tiny, tmp_path):
    """load_weights_into_model through csrc/loader.cu: the checkpoint's tensors go from the file mapping to HBM
    (no host tensors), legacy HF key layout included; the loaded model generates the same tokens."""
    from safetensors.torch import save_file

    from moondream_b200 import synth
    from moondream_b200.weights import legacy_key_map, load_state_dict_from_file, load_weights_into_model

    cfg, sd = tiny
    path = str(tmp_path / "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    dev_sd = load_state_dict_from_file(path, cfg, "cuda")
    assert all(t.is_cuda for t in dev_sd.values())
    for k, v in sd.items():
        assert torch.equal(dev_sd[k].cpu(), v), k
    legacy = {old: sd[new] for old, new in legacy_key_map(cfg).items()}
    legacy["region_model.coordinate_features.weight"] = sd["region.coord_features"].T.contiguous()
    legacy["region_model.size_features.weight"] = sd["region.size_features"].T.contiguous()
    lpath = str(tmp_path / "legacy.safetensors")
    save_file({k: v.contiguous() for k, v in legacy.items()}, lpath)
    a, b = _model(cfg, sd), _model(cfg, sd)
    load_weights_into_model(path, a)
    load_weights_into_model(lpath, b)
    img = synth.synthetic_image(0, 378, 378)
    s = {"temperature": 0, "max_tokens": 8}
    ref = _model(cfg, sd).caption(img, "short", settings=s)["caption"]
    assert a.caption(img, "short", settings=s)["caption"] == ref == b.caption(img, "short", settings=s)["caption"]

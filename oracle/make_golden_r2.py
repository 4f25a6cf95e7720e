"""TEST INFRASTRUCTURE (build container only): round-2 golden fixtures from the UNMODIFIED reference.

    python -m oracle.make_golden_r2

  tests/golden/tiny_reasoning.json  query(reasoning=True) (moondream.py:576-596 over _generate_reasoning :323-432):
                                    reasoning text / grounding / answer the reference returns, greedy
  tests/golden/tiny_text_only.json  query(image=None) (moondream.py:565-574): pure causal mask, BOS + prompt at 0
  tests/golden/tiny_gqa.json        grouped-query decoder (n_heads 4, n_kv_heads 2; text.py:36-38,49): tokens + KV probe
  tests/golden/top_p.json           _apply_top_p (moondream.py:270-278) on model logits: kept ids and probabilities
  tests/golden/tiny_lora.json       settings["variant"]: a synthetic rank-8 LoRA through lora.py / text.py:31-56 / layers.py:131-143
The oracle restatement is asserted equal to the reference on the way and contributes the margins.
"""
from __future__ import annotations

import json
import os

import torch
from PIL import Image

from moondream_b200 import config as C, synth
from oracle import reference_shim as R
from oracle.moondream_oracle import OracleModel

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REASONING_BIAS = (8.0, 11.0, 9.5)     # synth.special_token_bias(answer, coord, ground)
REASONING_CASES = [(0, 378, 378, "11 12 13", 24), (1, 500, 700, "12 12 13", 24), (2, 300, 200, "13 12 13", 24),
                   (3, 378, 378, "31 7", 12)]


def _grounding_from(tokens, coords, tk, decode):
    """restates moondream.py:363-378,410-430 for the oracle's (tokens, coords)"""
    text_chunks, ground = [[]], [[]]
    for t, c in zip(tokens, coords):
        if t in (tk.start_ground_points_id, tk.end_ground_id):
            text_chunks.append([])
            ground.append([])
        text_chunks[-1].append(t)
        if t == tk.coord_id:
            ground[-1].append(c)
    texts = [decode(ch) for ch in text_chunks]
    out, start = [], 0
    for txt, g in zip(texts, ground):
        if len(g) > 1:
            out.append({"start_idx": start, "end_idx": start + len(txt),
                        "points": [[g[i], g[i + 1]] for i in range(0, len(g) - (len(g) % 2), 2)]})
        start += len(txt)
    return "".join(texts), out


def main(out_dir: str | None = None):
    """writes the five fixtures into out_dir (default tests/golden); the tests regenerate into a scratch directory"""
    assert R.reference_available(), "run in the build container (/root/reference)"
    out_dir = out_dir or OUT
    torch.manual_seed(0)
    cfg = C.tiny()
    tk = cfg.tokenizer
    sd = synth.synthetic_state_dict(cfg, 0)

    # ---------------- reasoning ----------------
    sd_r = dict(sd)
    sd_r["text.lm_head.bias"] = synth.special_token_bias(sd, cfg, *REASONING_BIAS)
    ref = R.load_reference_model(cfg, sd_r)
    orc = OracleModel(cfg, sd_r)
    stub = R.StubTokenizer(cfg.text.vocab_size)
    cases = []
    for idx, h, w, question, max_tokens in REASONING_CASES:
        img = synth.synthetic_image(idx, h, w)
        settings = {"temperature": 0, "max_tokens": max_tokens}
        with torch.inference_mode():
            enc = ref.encode_image(Image.fromarray(img))
        out = ref.query(enc, question, reasoning=True, settings=settings)
        prompt = tk.templates["query"]["prefix"] + stub.encode(question).ids + tk.templates["query"]["suffix"] + [tk.thinking_id]
        o_enc = orc.encode_image(img)
        r = orc.generate_reasoning(o_enc, prompt, max_tokens)
        text, grounding = _grounding_from(r["tokens"], r["coords"], tk, stub.decode)
        assert text == out["reasoning"]["text"], (text, out["reasoning"]["text"])
        want_g = [{"start_idx": g["start_idx"], "end_idx": g["end_idx"], "points": [list(p) for p in g["points"]]}
                  for g in out["reasoning"]["grounding"]]
        assert grounding == want_g, (grounding, want_g)
        ans = orc.generate(None, tk.templates["query"]["suffix"], max_tokens, pos=r["pos"])
        assert ans.tokens == R.tokens_from_text(out["answer"]), (ans.tokens, out["answer"])
        cases.append({"image_index": idx, "height": h, "width": w, "question": question, "max_tokens": max_tokens,
                      "prompt": prompt, "reasoning_text": out["reasoning"]["text"], "grounding": want_g,
                      "reasoning_tokens": r["tokens"], "coords": r["coords"], "margin_ulps": r["margin_ulps"],
                      "coord_ulps": r["coord_ulps"], "end_margin_ulps": r["end_margin_ulps"],
                      "answer": out["answer"], "answer_tokens": ans.tokens, "answer_margin_ulps": ans.margin_ulps})
        print("reasoning", idx, r["tokens"][:10], want_g[:1], ans.tokens[:4])
    json.dump({"generator": "oracle/make_golden_r2.py (unmodified reference query(reasoning=True), greedy; tiny preset, "
                            "lm_head bias lifted with synth.special_token_bias%s)" % (REASONING_BIAS,),
               "bias": list(REASONING_BIAS), "cases": cases},
              open(os.path.join(out_dir, "tiny_reasoning.json"), "w"), indent=1)

    # ---------------- text-only query ----------------
    ref = R.load_reference_model(cfg, sd)
    orc = OracleModel(cfg, sd)
    tcases = []
    for question, max_tokens in (("11 12 13", 16), ("7 8", 12), ("100 200 300 400 500 600 700", 20)):
        out = ref.query(None, question, settings={"temperature": 0, "max_tokens": max_tokens})
        prompt = [tk.bos_id] + tk.templates["query"]["prefix"] + stub.encode(question).ids + tk.templates["query"]["suffix"] * 2
        gen = orc.generate(None, prompt, max_tokens)
        assert gen.tokens == R.tokens_from_text(out["answer"]), (gen.tokens, out["answer"])
        tcases.append({"question": question, "max_tokens": max_tokens, "prompt": prompt, "tokens": gen.tokens,
                       "margin_ulps": gen.margin_ulps})
        print("text-only", question, gen.tokens[:6])
    # reasoning without an image
    sd_r2 = dict(sd)
    sd_r2["text.lm_head.bias"] = synth.special_token_bias(sd, cfg, *REASONING_BIAS)
    ref_r = R.load_reference_model(cfg, sd_r2)
    orc_r = OracleModel(cfg, sd_r2)
    out = ref_r.query(None, "11 12 13", reasoning=True, settings={"temperature": 0, "max_tokens": 16})
    prompt = [tk.bos_id] + tk.templates["query"]["prefix"] + [11, 12, 13] + tk.templates["query"]["suffix"] + [tk.thinking_id]
    r = orc_r.generate_reasoning(None, prompt, 16)
    text, grounding = _grounding_from(r["tokens"], r["coords"], tk, stub.decode)
    assert text == out["reasoning"]["text"]
    ans = orc_r.generate(None, tk.templates["query"]["suffix"], 16, pos=r["pos"])
    assert ans.tokens == R.tokens_from_text(out["answer"])
    json.dump({"generator": "oracle/make_golden_r2.py (unmodified reference query(image=None), greedy; tiny preset)",
               "cases": tcases,
               "reasoning": {"question": "11 12 13", "max_tokens": 16, "bias": list(REASONING_BIAS), "prompt": prompt,
                             "reasoning_tokens": r["tokens"], "coords": r["coords"], "margin_ulps": r["margin_ulps"],
                             "coord_ulps": r["coord_ulps"], "end_margin_ulps": r["end_margin_ulps"],
                             "answer_tokens": ans.tokens, "answer_margin_ulps": ans.margin_ulps}},
              open(os.path.join(out_dir, "tiny_text_only.json"), "w"), indent=1)

    # ---------------- grouped-query attention ----------------
    gcfg = C.tiny_gqa()
    gsd = synth.synthetic_state_dict(gcfg, 0)
    gref = R.load_reference_model(gcfg, gsd)
    gorc = OracleModel(gcfg, gsd)
    gcases = []
    for idx, h, w, plen, ntok in ((0, 378, 378, 5, 16), (1, 500, 700, 9, 16)):
        img = synth.synthetic_image(idx, h, w)
        prompt = synth.synthetic_prompt(idx, plen, gcfg.text.vocab_size)
        with torch.inference_mode():
            enc = gref.encode_image(Image.fromarray(img))
        gref.load_encoded_image(enc)
        tokens = R.tokens_from_text("".join(gref._generate_answer(torch.tensor([prompt]), enc.pos,
                                                                  {"temperature": 0, "max_tokens": ntok})))
        o_enc = gorc.encode_image(img)
        assert tuple(enc.caches[0][0].shape) == (1, 2, 730, 64)
        assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(enc.caches, o_enc.caches))
        gen = gorc.generate(o_enc, prompt, ntok)
        assert gen.tokens == tokens, (gen.tokens, tokens)
        gcases.append({"image_index": idx, "height": h, "width": w, "prompt": prompt, "tokens": tokens,
                       "margin_ulps": gen.margin_ulps,
                       "kv_abs_mean_first_last": [float(enc.caches[i][0].float().abs().mean()) for i in (0, 3)]})
        print("gqa", idx, tokens[:8])
    json.dump({"generator": "oracle/make_golden_r2.py (unmodified reference, config tiny-gqa: 4 query heads, 2 KV heads)",
               "cases": gcases}, open(os.path.join(out_dir, "tiny_gqa.json"), "w"), indent=1)

    # ---------------- LoRA variant (settings["variant"], lora.py) ----------------
    import tempfile

    flat = synth.synthetic_lora(cfg, rank=8, seed=0)
    tmp = tempfile.mkdtemp()
    os.environ["HF_HUB_CACHE"] = tmp
    os.makedirs(os.path.join(tmp, "md_variants", "synthetic-r8"))
    torch.save(flat, os.path.join(tmp, "md_variants", "synthetic-r8", "final.pt"))
    ref = R.load_reference_model(cfg, sd)
    orc = OracleModel(cfg, sd)
    orc.lora = synth.nest_lora(flat)
    lcases = []
    for idx, h, w, plen, ntok in ((0, 378, 378, 6, 12), (1, 500, 700, 9, 12)):
        img = synth.synthetic_image(idx, h, w)
        prompt = synth.synthetic_prompt(idx, plen, cfg.text.vocab_size)
        settings = {"temperature": 0, "max_tokens": ntok, "variant": "synthetic-r8"}
        with torch.inference_mode():
            enc = ref.encode_image(Image.fromarray(img), settings)
        ref.load_encoded_image(enc)
        tokens = R.tokens_from_text("".join(ref._generate_answer(torch.tensor([prompt]), enc.pos, settings)))
        o_enc = orc.encode_image(img)
        assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(enc.caches, o_enc.caches))
        gen = orc.generate(o_enc, prompt, ntok)
        assert gen.tokens == tokens, (gen.tokens, tokens)
        det = ref.detect(enc, "17 23", settings={"max_objects": 2, "variant": "synthetic-r8"})["objects"]
        dprompt = tk.templates["detect"]["prefix"] + [17, 23] + tk.templates["detect"]["suffix"]
        o_det = orc.generate_points(o_enc, dprompt, True, 2)
        assert [{k: o[k] for k in d} for o, d in zip(o_det, det)] == det and len(o_det) == len(det)
        lcases.append({"image_index": idx, "height": h, "width": w, "prompt": prompt, "tokens": tokens,
                       "margin_ulps": gen.margin_ulps, "detect_prompt": dprompt, "detect_boxes": det,
                       "detect_bins": [o["bins"] for o in o_det], "detect_ulps": [o["ulps"] for o in o_det],
                       "kv_abs_mean_first_last": [float(enc.caches[i][0].float().abs().mean()) for i in (0, cfg.text.n_layers - 1)]})
        print("lora", idx, tokens[:8], det[:1])
    orc.lora = None
    json.dump({"generator": "oracle/make_golden_r2.py (unmodified reference with settings['variant'] = a synthetic rank-8 "
                            "LoRA, synth.synthetic_lora(cfg, 8, 0), placed in the reference's variant cache layout)",
               "rank": 8, "seed": 0, "cases": lcases}, open(os.path.join(out_dir, "tiny_lora.json"), "w"), indent=1)

    # ---------------- _apply_top_p on model logits ----------------
    pcases = []
    img = synth.synthetic_image(0, 378, 378)
    o_enc = orc.encode_image(img)
    for seed, (temp, top_p) in enumerate(((0.5, 0.3), (1.0, 0.9), (2.0, 0.5), (1.5, 0.95), (0.25, 0.3), (1.0, 0.05))):
        prompt = synth.synthetic_prompt(40 + seed, 6, cfg.text.vocab_size)
        orc.load_encoded(o_enc)
        logits = orc.prefill_prompt(prompt, o_enc.pos)[0]
        probs = torch.softmax(logits / temp, dim=-1)
        kept = ref._apply_top_p(probs.clone(), top_p)
        nz = kept[0].nonzero().flatten()
        # how close the first dropped element is to staying (in bf16 ulps of top_p): ambiguity measure for the device test
        srt, _ = torch.sort(probs, dim=-1, descending=True)
        before = (torch.cumsum(srt, dim=-1) - srt)[0].float()
        n_keep = int((before <= torch.tensor(top_p, dtype=torch.bfloat16).float()).sum())
        pcases.append({"prompt": prompt, "temperature": temp, "top_p": top_p,
                       "logits": logits[0].float().tolist(), "kept_ids": nz.tolist(),
                       "kept_probs": kept[0, nz].float().tolist(), "n_kept_by_mass": n_keep,
                       "mass_before_last_kept": float(before[len(nz) - 1]),
                       "mass_before_first_dropped": float(before[len(nz)]) if len(nz) < before.numel() else None})
        print("top_p", temp, top_p, len(nz))
    json.dump({"generator": "oracle/make_golden_r2.py (MoondreamModel._apply_top_p of the unmodified reference on bf16 "
                            "prefill logits of the tiny preset)", "cases": pcases},
              open(os.path.join(out_dir, "top_p.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main()

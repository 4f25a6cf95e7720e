"""TEST INFRASTRUCTURE (build container only): generate tests/golden/* from the UNMODIFIED reference.

    python -m oracle.make_golden

Runs /root/reference (stub tokenizer, seeded synthetic weights, moondream_b200.synth) on the tiny
configuration and records what its public API returns: generated token ids (through
MoondreamModel._generate_answer, the call caption()/query() make, moondream.py:645), detect boxes and
points.  The oracle restatement is asserted bit-identical on the way (tokens, KV prefix, boxes) and
contributes what the reference API does not expose (top-1/top-2 margins, region bins).
Also pins the synthetic-weight recipe and the host crop geometry (hashes).
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import torch
from PIL import Image

from moondream_b200 import config as C, synth
from oracle import reference_shim as R
from oracle.moondream_oracle import OracleModel, overlap_crops

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [  # name, image index, H, W, prompt length, new tokens
    ("single_crop", 0, 378, 378, 5, 20),
    ("multi_crop_2x3", 1, 500, 700, 9, 20),
    ("multi_crop_4x2", 2, 800, 600, 32, 16),
    ("small_image", 3, 300, 200, 3, 12),
]


def main():
    assert R.reference_available(), "run in the build container (/root/reference)"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    cfg = C.tiny()
    sd = synth.synthetic_state_dict(cfg, 0)
    ref = R.load_reference_model(cfg, sd)
    orc = OracleModel(cfg, sd)
    cases = []
    for name, idx, h, w, plen, ntok in CASES:
        img = synth.synthetic_image(idx, h, w)
        prompt = synth.synthetic_prompt(idx, plen, cfg.text.vocab_size)
        with torch.inference_mode():
            enc = ref.encode_image(Image.fromarray(img))
        ref.load_encoded_image(enc)
        text = "".join(ref._generate_answer(torch.tensor([prompt]), enc.pos,
                                            {"temperature": 0, "max_tokens": ntok}))
        tokens = R.tokens_from_text(text)
        o_enc = orc.encode_image(img)
        assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(enc.caches, o_enc.caches))
        gen = orc.generate(o_enc, prompt, ntok)
        assert gen.tokens == tokens, (name, gen.tokens, tokens)
        tk = cfg.tokenizer
        det = ref.detect(enc, "17 23", settings={"max_objects": 3})["objects"]
        dprompt = tk.templates["detect"]["prefix"] + [17, 23] + tk.templates["detect"]["suffix"]
        o_det = orc.generate_points(o_enc, dprompt, True, 3)
        assert [{k: o[k] for k in d} for o, d in zip(o_det, det)] == det and len(o_det) == len(det)
        pts = ref.point(enc, "17 23", settings={"max_objects": 3})["points"]
        pprompt = tk.templates["point"]["prefix"] + [17, 23] + tk.templates["point"]["suffix"]
        o_pts = orc.generate_points(o_enc, pprompt, False, 3)
        assert [{"x": o["x"], "y": o["y"]} for o in o_pts] == pts
        kv_probe = [float(enc.caches[i][0].float().abs().mean()) for i in (0, cfg.text.n_layers - 1)]
        cases.append({"name": name, "image_index": idx, "height": h, "width": w, "prompt": prompt,
                      "tokens": tokens, "margins": gen.margins, "margin_ulps": gen.margin_ulps, "detect_prompt": dprompt,
                      "detect_boxes": det, "detect_bins": [o["bins"] for o in o_det], "detect_ulps": [o["ulps"] for o in o_det],
                      "point_prompt": pprompt, "points": pts, "point_bins": [o["bins"] for o in o_pts], "point_ulps": [o["ulps"] for o in o_pts],
                      "kv_abs_mean_first_last": kv_probe})
        print(name, tokens[:8], det[:1])
    json.dump({"generator": "oracle/make_golden.py (unmodified reference, tiny preset, seed 0)",
               "torch": torch.__version__, "cases": cases},
              open(os.path.join(OUT, "tiny_reference.json"), "w"), indent=1)

    # nucleus sampling (moondream.py:270-278, 312-318, 524-530): the reference's tokens under a fixed global seed
    SAMPLING = [("single_crop", 0, 378, 378, 5, 0.5, 0.3, 1), ("single_crop_hot", 0, 378, 378, 5, 1.0, 0.9, 2),
                ("multi_crop_hot", 1, 500, 700, 9, 2.0, 0.5, 3)]
    sampled = []
    for name, idx, h, w, plen, temp, top_p, seed in SAMPLING:
        img = synth.synthetic_image(idx, h, w)
        prompt = synth.synthetic_prompt(idx, plen, cfg.text.vocab_size)
        with torch.inference_mode():
            enc = ref.encode_image(Image.fromarray(img))
        ref.load_encoded_image(enc)
        torch.manual_seed(seed)
        text = "".join(ref._generate_answer(torch.tensor([prompt]), enc.pos,
                                            {"temperature": temp, "top_p": top_p, "max_tokens": 12}))
        tokens = R.tokens_from_text(text)
        torch.manual_seed(seed)
        gen = orc.generate(orc.encode_image(img), prompt, 12, temperature=temp, top_p=top_p)
        assert gen.tokens == tokens, (name, gen.tokens, tokens)
        sampled.append({"name": name, "image_index": idx, "height": h, "width": w, "prompt": prompt,
                        "temperature": temp, "top_p": top_p, "seed": seed, "tokens": tokens})
        print(name, temp, top_p, tokens[:8])
    json.dump({"generator": "oracle/make_golden.py (unmodified reference, tiny preset, torch.manual_seed(seed) before "
                            "_generate_answer)", "torch": torch.__version__, "cases": sampled},
              open(os.path.join(OUT, "tiny_sampling.json"), "w"), indent=1)

    # streaming detokenisation (moondream.py:476-537): the chunks the reference's generator yields for token
    # sequences it sampled itself, with a tokenizer whose pieces hit every flush rule
    stream_cases = []
    img = synth.synthetic_image(0, 378, 378)
    with torch.inference_mode():
        enc = ref.encode_image(Image.fromarray(img))
    stub = ref.tokenizer
    for seed, temp in ((11, 1.0), (12, 2.0), (13, 3.0), (14, 0.0)):
        prompt = synth.synthetic_prompt(seed, 4, cfg.text.vocab_size)
        settings = {"temperature": temp, "top_p": 0.95, "max_tokens": 40}
        ref.load_encoded_image(enc)
        torch.manual_seed(seed)
        tokens = R.tokens_from_text("".join(ref._generate_answer(torch.tensor([prompt]), enc.pos, settings)))
        ref.tokenizer = R.PieceTokenizer()
        try:
            ref.load_encoded_image(enc)
            torch.manual_seed(seed)
            chunks = list(ref._generate_answer(torch.tensor([prompt]), enc.pos, settings))
        finally:
            ref.tokenizer = stub
        assert "".join(chunks) == R.PieceTokenizer().decode(tokens)
        stream_cases.append({"seed": seed, "temperature": temp, "tokens": tokens, "chunks": chunks})
        print("stream", seed, len(tokens), len(chunks))
    json.dump({"generator": "oracle/make_golden.py (unmodified reference generator, oracle.reference_shim.PieceTokenizer)",
               "pieces": R.PieceTokenizer.PIECES, "cases": stream_cases},
              open(os.path.join(OUT, "streaming.json"), "w"), indent=1, ensure_ascii=True)

    # prompt assembly (moondream.py:541-604 query incl. spatial refs and the duplicated suffix, :625-651 caption,
    # :735-829 detect / point): the token ids the reference actually hands to its prefill
    seen = []
    orig_prefill = ref._prefill_prompt

    def recording_prefill(prompt_tokens, pos, *a, **k):
        seen.append((prompt_tokens.flatten().tolist(), int(pos)))
        return orig_prefill(prompt_tokens, pos, *a, **k)

    ref._prefill_prompt = recording_prefill
    prompt_cases = []
    try:
        greedy = {"temperature": 0, "max_tokens": 1}
        calls = [("caption_short", lambda: ref.caption(enc, "short", settings=greedy), {"length": "short"}),
                 ("caption_normal", lambda: ref.caption(enc, "normal", settings=greedy), {"length": "normal"}),
                 ("query", lambda: ref.query(enc, "11 12 13", settings=greedy), {"question": "11 12 13"}),
                 ("query_refs", lambda: ref.query(enc, "15 16", spatial_refs=[(0.25, 0.75), (0.1, 0.2, 0.5, 0.9)],
                                                  settings=greedy),
                  {"question": "15 16", "spatial_refs": [[0.25, 0.75], [0.1, 0.2, 0.5, 0.9]]}),
                 ("detect", lambda: ref.detect(enc, "17 23", settings={"max_objects": 1}), {"object": "17 23"}),
                 ("point", lambda: ref.point(enc, "17 23", settings={"max_objects": 1}), {"object": "17 23"})]
        for name, call, args in calls:
            seen.clear()
            call()
            assert len(seen) == 1, (name, len(seen))
            prompt_cases.append({"name": name, "args": args, "prompt": seen[0][0], "pos": seen[0][1]})
            print("prompt", name, seen[0][0])
    finally:
        ref._prefill_prompt = orig_prefill
    json.dump({"generator": "oracle/make_golden.py (prompt tokens the unmodified reference passes to _prefill_prompt; "
                            "StubTokenizer)", "cases": prompt_cases},
              open(os.path.join(OUT, "prompts.json"), "w"), indent=1)

    # query with spatial references (moondream.py:293-301, region.py:96-136): the reference's answer tokens, plus the
    # rows of the prompt embedding its region encoders produced (taken from the oracle after asserting bit equality
    # of the prefill logits with the reference's)
    tk = cfg.tokenizer
    ref_cases = []
    img = synth.synthetic_image(2, 500, 700)
    with torch.inference_mode():
        enc = ref.encode_image(Image.fromarray(img))
    o_enc = orc.encode_image(img)
    got = []
    orig_prefill = ref._prefill_prompt

    def capture(prompt_tokens, pos, *a, **k):
        out = orig_prefill(prompt_tokens, pos, *a, **k)
        got.append((prompt_tokens.flatten().tolist(), out[0].clone()))
        return out

    ref._prefill_prompt = capture
    try:
        for refs in ([(0.25, 0.75)], [(0.1, 0.2, 0.5, 0.9)], [(0.25, 0.75), (0.1, 0.2, 0.5, 0.9), (0.6, 0.6)]):
            got.clear()
            text = ref.query(enc, "15 16", spatial_refs=refs, settings={"temperature": 0, "max_tokens": 8})["answer"]
            prompt, ref_logits = got[0]
            emb = orc.spatial_prompt_embeds(prompt, refs)
            orc.load_encoded(o_enc)
            assert torch.equal(orc.prefill_prompt(prompt, o_enc.pos, emb)[0], ref_logits)
            gen = orc.generate(o_enc, prompt, 8, spatial_refs=refs)
            assert gen.tokens == R.tokens_from_text(text)
            rows = [i for i, t in enumerate(prompt) if t in (tk.coord_id, tk.size_id)]
            ref_cases.append({"spatial_refs": [list(r) for r in refs], "question": "15 16", "prompt": prompt,
                              "tokens": gen.tokens, "margin_ulps": gen.margin_ulps, "rows": rows,
                              "row_embeds": emb[0, rows].float().tolist()})
            print("refs", refs, gen.tokens[:4])
    finally:
        ref._prefill_prompt = orig_prefill
    json.dump({"generator": "oracle/make_golden.py (unmodified reference query(spatial_refs=...); embedding rows from the "
                            "oracle after bit-equality of the prefill logits)", "image": [2, 500, 700], "cases": ref_cases},
              open(os.path.join(OUT, "tiny_spatial_refs.json"), "w"), indent=1)

    hashes = {}
    for preset in ("tiny", "moondream-0.5b"):
        c = C.preset(preset)
        s = sd if preset == "tiny" else synth.synthetic_state_dict(c, 0)
        hashes[preset] = synth.state_dict_fingerprint(s, synth.FINGERPRINT_KEYS)
        del s
    json.dump(hashes, open(os.path.join(OUT, "synth_hashes.json"), "w"), indent=1)

    import sys
    sys.path.insert(0, R.REFERENCE_ROOT)
    from moondream.torch.image_crops import overlap_crop_image as ref_crop  # the reference itself
    crops = []
    for (h, w) in [(378, 378), (300, 200), (500, 700), (800, 600), (768, 1024), (1024, 768), (1080, 1920),
                   (756, 756), (50, 1000), (2000, 3000)]:
        img = synth.synthetic_image(h + w, h, w)
        out = ref_crop(img, overlap_margin=4, max_crops=12)
        mine, tiling = overlap_crops(img, 4, 12)
        assert tuple(out["tiling"]) == tiling and np.array_equal(out["crops"], mine)
        crops.append({"height": h, "width": w, "image_index": h + w, "tiling": list(out["tiling"]),
                      "n_crops": int(out["crops"].shape[0]),
                      "sha256": hashlib.sha256(out["crops"].tobytes()).hexdigest()[:16]})
    json.dump({"generator": "reference overlap_crop_image (PIL Lanczos branch)", "cases": crops},
              open(os.path.join(OUT, "crops.json"), "w"), indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()

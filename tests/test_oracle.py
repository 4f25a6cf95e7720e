"""The oracle (CPU restatement of the reference) against (a) the committed golden fixtures produced by
the unmodified reference and (b) the reference itself when /root/reference is present (build container)."""
import json
import os

import numpy as np
import pytest
import torch

from moondream_b200 import config as C, synth
from oracle import reference_shim as R
from oracle.moondream_oracle import OracleModel

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tiny():
    cfg = C.tiny()
    sd = synth.synthetic_state_dict(cfg, 0)
    return cfg, sd, OracleModel(cfg, sd)


def test_synthetic_weights_are_reproducible(tiny):
    cfg, sd, _ = tiny
    gold = json.load(open(os.path.join(GOLDEN, "synth_hashes.json")))
    assert synth.state_dict_fingerprint(sd, synth.FINGERPRINT_KEYS) == gold["tiny"]
    assert synth.param_count(C.moondream_2b()) == 1_927_237_104
    assert synth.param_count(C.moondream_0_5b()) == 631_889_382


def test_oracle_reproduces_reference_golden(tiny):
    cfg, sd, orc = tiny
    gold = json.load(open(os.path.join(GOLDEN, "tiny_reference.json")))
    for case in gold["cases"]:
        img = synth.synthetic_image(case["image_index"], case["height"], case["width"])
        enc = orc.encode_image(img)
        gen = orc.generate(enc, case["prompt"], len(case["tokens"]))
        assert gen.tokens == case["tokens"], case["name"]
        assert np.allclose(gen.margins, case["margins"])
        det = orc.generate_points(enc, case["detect_prompt"], True, 3)
        assert [d["bins"] for d in det] == case["detect_bins"]
        for d, want in zip(det, case["detect_boxes"]):
            assert all(d[k] == want[k] for k in want)
        pts = orc.generate_points(enc, case["point_prompt"], False, 3)
        assert [{"x": p["x"], "y": p["y"]} for p in pts] == case["points"]
        probe = [float(enc.caches[i][0].float().abs().mean()) for i in (0, cfg.text.n_layers - 1)]
        assert np.allclose(probe, case["kv_abs_mean_first_last"])


def test_teacher_forcing_is_consistent(tiny):
    cfg, sd, orc = tiny
    img = synth.synthetic_image(5, 378, 378)
    enc = orc.encode_image(img)
    prompt = synth.synthetic_prompt(5, 4, cfg.text.vocab_size)
    free = orc.generate(enc, prompt, 6)
    forced = orc.generate(enc, prompt, 6, forced=free.tokens)
    assert forced.predicted == free.tokens and np.allclose(forced.margins, free.margins)


def test_fp32_truth_is_close_to_bf16_port(tiny):
    cfg, sd, orc = tiny
    truth = OracleModel(cfg, sd, dtype=torch.float32)
    img = synth.synthetic_image(1, 378, 378)
    a = orc.vision_encoder(orc.prepare_crops(img)[0]).float()
    b = truth.vision_encoder(truth.prepare_crops(img)[0])
    assert ((a - b).norm() / b.norm()).item() < 5e-2


@pytest.mark.skipif(not R.reference_available(), reason="/root/reference only exists in the build container")
def test_oracle_is_bit_identical_to_reference(tiny):
    from PIL import Image

    cfg, sd, orc = tiny
    ref = R.load_reference_model(cfg, sd)
    img = synth.synthetic_image(11, 600, 450)
    with torch.inference_mode():
        enc = ref.encode_image(Image.fromarray(img))
    o_enc = orc.encode_image(img)
    for (k, v), (ok, ov) in zip(enc.caches, o_enc.caches):
        assert torch.equal(k, ok) and torch.equal(v, ov)
    text = ref.caption(enc, "short", settings={"temperature": 0, "max_tokens": 10})["caption"]
    gen = orc.generate(o_enc, cfg.tokenizer.templates["caption"]["short"], 10)
    assert R.tokens_from_text(text) == gen.tokens

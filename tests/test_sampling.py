"""Nucleus sampling (reference moondream.py:270-278, 312-318, 524-530): the oracle's restatement is pinned to the
unmodified reference by tests/golden/tiny_sampling.json (tokens the reference sampled under a fixed global seed); the
product's host sampler must then agree with the oracle bit for bit on the same logits and RNG state."""
import json
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def tiny():
    from moondream_b200 import config as C, synth
    from oracle.moondream_oracle import OracleModel

    cfg = C.tiny()
    sd = synth.synthetic_state_dict(cfg, 0)
    return cfg, OracleModel(cfg, sd)


def test_oracle_sampling_reproduces_the_reference(tiny):
    from moondream_b200 import synth

    cfg, orc = tiny
    gold = json.load(open(os.path.join(HERE, "golden", "tiny_sampling.json")))
    assert len(gold["cases"]) >= 3
    for c in gold["cases"]:
        img = synth.synthetic_image(c["image_index"], c["height"], c["width"])
        enc = orc.encode_image(img)
        torch.manual_seed(c["seed"])
        gen = orc.generate(enc, c["prompt"], len(c["tokens"]), temperature=c["temperature"], top_p=c["top_p"])
        assert gen.tokens == c["tokens"], (c["name"], gen.tokens, c["tokens"])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_host_sampler_matches_the_oracle(dtype):
    from moondream_b200.sampling import HostSampler, apply_top_p, sample_next
    from oracle.moondream_oracle import OracleModel

    for seed in range(12):
        g = torch.Generator().manual_seed(seed)
        logits = (torch.randn(1, 2048, generator=g) * 3).to(dtype)
        logits[0, 5] = float("-inf")                                  # a masked id (answer_id) must never be drawn
        for temp, top_p in ((0.5, 0.3), (1.0, 0.9), (2.0, 0.05), (0.0, 0.3)):
            torch.manual_seed(seed)
            want = OracleModel.next_token(logits.clone(), temp, top_p)
            torch.manual_seed(seed)
            got = int(sample_next(logits.clone(), temp, top_p).item())
            assert got == want and got != 5, (seed, temp, top_p, got, want)
            torch.manual_seed(seed)
            assert int(HostSampler(temp, top_p)(logits.clone()).item()) == want
    # batch rows are sampled independently of their order in the nucleus mask
    p = torch.softmax(torch.randn(4, 300, generator=torch.Generator().manual_seed(0)), dim=-1)
    kept = apply_top_p(p.clone(), 0.4)
    assert torch.allclose(kept.sum(-1), torch.ones(4)) and bool(((kept > 0).sum(-1) >= 1).all())
    top = p.argmax(-1)
    assert bool((kept[torch.arange(4), top] > 0).all())              # the most likely token always survives


def test_sampler_rejects_negative_temperature():
    from moondream_b200.sampling import HostSampler

    with pytest.raises(ValueError):
        HostSampler(-0.1, 0.3)

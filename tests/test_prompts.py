"""Prompt assembly of the API mirror (rows a22 / a25 of SURVEY.md §8; reference moondream.py:541-604 query with
spatial refs and the duplicated suffix, :625-651 caption, :735-829 detect / point) against the token ids the
unmodified reference handed to its own prefill (tests/golden/prompts.json).  Host logic only: no GPU."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def holder():
    from moondream_b200 import config as C
    from oracle.reference_shim import StubTokenizer

    cfg = C.tiny()
    return type("Holder", (), {"config": cfg, "tokenizer": StubTokenizer(cfg.text.vocab_size)})()


def test_prompts_equal_the_reference(holder):
    from moondream_b200.moondream import MoondreamModel as M

    gold = {c["name"]: c for c in json.load(open(os.path.join(HERE, "golden", "prompts.json")))["cases"]}
    tk = holder.config.tokenizer
    assert list(tk.templates["caption"]["short"]) == gold["caption_short"]["prompt"]
    assert list(tk.templates["caption"]["normal"]) == gold["caption_normal"]["prompt"]
    q = gold["query"]
    assert M._query_prompt(holder, q["args"]["question"], None, False) == q["prompt"]
    q = gold["query_refs"]
    refs = [tuple(r) for r in q["args"]["spatial_refs"]]
    assert M._query_prompt(holder, q["args"]["question"], refs, False) == q["prompt"]
    for kind in ("detect", "point"):
        g = gold[kind]
        assert M._object_prompts(holder, kind, [g["args"]["object"]]) == [g["prompt"]]
    # every prompt starts after the 730-position image prefix
    assert all(c["pos"] == 730 for c in gold.values())


class _PointsEngine:
    """CPU stand-in for Engine behind MoondreamModel.detect_batch / point_batch: an image's first byte names it."""

    def __init__(self):
        self.calls = []

    def encode_images(self, arrs, lora=None):
        from moondream_b200.engine import PrefixKV

        return [PrefixKV(730, [int(a[0, 0, 0])], None) for a in arrs]

    def generate_points(self, prefixes, prompts, include_size, max_objects, lora=None):
        assert len(prompts) == len(prefixes)
        self.calls.append((len(prefixes), include_size, max_objects))
        out = []
        for p, pr in zip(prefixes, prompts):
            i = p.pages[0]
            obj = ({"x_min": float(i), "y_min": float(len(pr)), "x_max": 1.0, "y_max": 2.0, "bins": [0, 0, 0, 0]}
                   if include_size else {"x": float(i), "y": float(len(pr)), "bins": [0, 0]})
            out.append([obj] * (i % 3))
        return out


def test_detect_and_point_batches_run_in_chunks_of_max_batch(holder):
    import numpy as np

    from moondream_b200.moondream import MoondreamModel

    model = MoondreamModel(holder.config, tokenizer=holder.tokenizer, max_batch=3)
    model._engine = eng = _PointsEngine()
    images = []
    for i in range(8):
        im = np.zeros((20, 20, 3), dtype=np.uint8)
        im[0, 0, 0] = i
        images.append(im)
    objects = [" ".join(str(40 + k) for k in range(1 + i % 4)) for i in range(8)]       # 1..4 tokens per object name
    tpl = holder.config.tokenizer.templates
    det = model.detect_batch(images, objects, settings={"max_objects": 5})
    assert eng.calls == [(3, True, 5), (3, True, 5), (2, True, 5)]
    for i, d in enumerate(det):
        n_prompt = len(tpl["detect"]["prefix"]) + 1 + i % 4 + len(tpl["detect"]["suffix"])
        assert d == {"objects": [{"x_min": float(i), "y_min": float(n_prompt), "x_max": 1.0, "y_max": 2.0}] * (i % 3)}
    eng.calls.clear()
    pts = model.point_batch(images[:4], objects[:4])
    assert eng.calls == [(3, False, 50), (1, False, 50)]                                # the reference's default max_objects
    for i, d in enumerate(pts):
        n_prompt = len(tpl["point"]["prefix"]) + 1 + i % 4 + len(tpl["point"]["suffix"])
        assert d == {"points": [{"x": float(i), "y": float(n_prompt)}] * (i % 3)}
    assert model.detect(images[2], "7") == det[2] | {"objects": [{**o, "y_min": float(len(tpl["detect"]["prefix"]) + 1 + len(tpl["detect"]["suffix"]))}
                                                                     for o in det[2]["objects"]]}
    with pytest.raises(ValueError):
        model.detect_batch(images[:2], objects[:1])


def test_reasoning_text_and_grounding_equal_the_reference(holder):
    """MoondreamModel._reasoning_result (the host half of query(reasoning=True): chunking at start_ground_points /
    end_ground, text spans, point pairs; moondream.py:363-432) on the reasoning tokens and coordinates of
    tests/golden/tiny_reasoning.json must give the text and grounding the UNMODIFIED reference returned for them."""
    from moondream_b200.moondream import MoondreamModel as M

    gold = json.load(open(os.path.join(HERE, "golden", "tiny_reasoning.json")))
    grounded = 0
    for c in gold["cases"]:
        text, grounding = M._reasoning_result(holder, c["reasoning_tokens"], c["coords"])
        assert text == c["reasoning_text"]
        assert [{"start_idx": g["start_idx"], "end_idx": g["end_idx"], "points": [list(p) for p in g["points"]]}
                for g in grounding] == c["grounding"]
        grounded += len(grounding)
        assert M._query_prompt(holder, c["question"], None, False, reasoning=True) == c["prompt"]
    assert grounded >= 1, "the fixture must contain at least one grounded span"


def test_text_settings_follow_the_reference_defaults():
    """TextSamplingSettings (moondream.py:443-454, :51-53): temperature 0.5, top_p 0.3, max_tokens 768."""
    from moondream_b200.moondream import MoondreamModel as M

    assert M._text_settings(None) == (768, {"temperature": 0.5, "top_p": 0.3})
    assert M._text_settings({"temperature": 0}) == (768, {})
    assert M._text_settings({"temperature": 1, "top_p": 0.9, "max_tokens": 5, "seed": 7}) == (5, {"temperature": 1.0, "top_p": 0.9, "seed": 7})
    mt, s = M._text_settings({"temperature": 0.7, "host_sampler": True})
    assert mt == 768 and set(s) == {"sampler"}
    with pytest.raises(ValueError):
        M._text_settings({"temperature": -1})

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

"""The drop-in `MoondreamModel` surface (reference moondream/torch/moondream.py) driven the way the reference's
callers drive it: encode_image / caption / query / detect / point, settings keys, result shapes, exceptions."""
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from moondream_b200 import config as C, synth
    from moondream_b200.moondream import MoondreamModel
    from oracle.moondream_oracle import OracleModel
    from oracle.reference_shim import StubTokenizer

    cfg = C.tiny()
    sd = synth.synthetic_state_dict(cfg, 0)
    model = MoondreamModel(cfg, tokenizer=StubTokenizer(cfg.text.vocab_size), max_batch=8)
    model.load_state_dict(sd)
    return cfg, model, OracleModel(cfg, sd)


def _ids(text):
    return [int(t) for t in text.split()]


def _agree(got, gen, what):
    for i, (a, b) in enumerate(zip(got, gen.tokens)):
        if a != b:
            assert gen.margin_ulps[i] < 4.5, (what, i, a, b, gen.margin_ulps[i])
            return


def test_caption_query_through_the_api(api):
    from moondream_b200 import synth
    from moondream_b200.moondream import EncodedImage

    cfg, model, orc = api
    img = synth.synthetic_image(3, 500, 700)
    enc = model.encode_image(Image.fromarray(img))
    assert isinstance(enc, EncodedImage) and enc.pos == 730 and model.encode_image(enc) is enc
    assert len(enc.caches) == cfg.text.n_layers and tuple(enc.caches[0][0].shape) == (1, cfg.text.n_heads, 730, 64)
    settings = {"temperature": 0, "max_tokens": 10}
    o_enc = orc.encode_image(img)
    out = model.caption(enc, "short", settings=settings)
    assert set(out) == {"caption"} and isinstance(out["caption"], str)
    _agree(_ids(out["caption"]), orc.generate(o_enc, cfg.tokenizer.templates["caption"]["short"], 10), "caption")
    chunks = list(model.caption(enc, "normal", stream=True, settings=settings)["caption"])
    assert "".join(chunks) == model.caption(enc, "normal", settings=settings)["caption"]
    # query: prefix + question + suffix + suffix (the reference's duplicated suffix, moondream.py:586-604)
    q = model.query(enc, "11 12 13", settings=settings)
    tk = cfg.tokenizer
    prompt = tk.templates["query"]["prefix"] + [11, 12, 13] + tk.templates["query"]["suffix"] * 2
    _agree(_ids(q["answer"]), orc.generate(o_enc, prompt, 10), "query")
    # batch calls agree with single calls
    many = model.caption_batch([img, synth.synthetic_image(4, 378, 378)], "short", settings=settings)
    assert len(many) == 2 and set(many[0]) == {"caption"}
    _agree(_ids(many[0]["caption"]), orc.generate(o_enc, cfg.tokenizer.templates["caption"]["short"], 10), "caption_batch")


def test_detect_point_through_the_api(api):
    from moondream_b200 import synth

    cfg, model, orc = api
    img = synth.synthetic_image(8, 800, 600)
    enc = model.encode_image(img)
    o_enc = orc.encode_image(img)
    tk = cfg.tokenizer
    det = model.detect(enc, "17 23", settings={"max_objects": 2})
    assert set(det) == {"objects"} and all(set(o) == {"x_min", "y_min", "x_max", "y_max"} for o in det["objects"])
    want = orc.generate_points(o_enc, tk.templates["detect"]["prefix"] + [17, 23] + tk.templates["detect"]["suffix"], True, 2)
    if want and min(want[0]["ulps"][:4]) >= 4.5:
        assert len(det["objects"]) >= 1
        for k in ("x_min", "y_min", "x_max", "y_max"):
            assert abs(det["objects"][0][k] - want[0][k]) < 1e-5
    pts = model.point(enc, "17 23", settings={"max_objects": 2})
    assert set(pts) == {"points"} and all(set(p) == {"x", "y"} for p in pts["points"])


def test_error_behaviour_matches_the_reference(api):
    from moondream_b200 import synth

    cfg, model, _ = api
    img = synth.synthetic_image(1, 378, 378)
    with pytest.raises(ValueError):
        model.encode_image("not an image")                       # moondream.py:237-238
    with pytest.raises(ValueError):
        model.query(img, None, settings={"temperature": 0})      # moondream.py:553-554
    with pytest.raises(ValueError):
        model.query(None, "x", spatial_refs=[(0.5, 0.5)], settings={"temperature": 0})   # :556-557
    with pytest.raises(ValueError):
        model.caption(img, "poetic", settings={"temperature": 0})  # :634-635
    with pytest.raises(ValueError):
        model.caption(img, "short", settings={"temperature": -1.0})
    with pytest.raises(RuntimeError):                                # a variant that is not cached is never downloaded
        model.caption(img, "short", settings={"temperature": 0, "variant": "no-such-variant"})


def test_spatial_refs_query_runs(api):
    from moondream_b200 import synth

    cfg, model, _ = api
    img = synth.synthetic_image(2, 378, 378)
    out = model.query(img, "15 16", spatial_refs=[(0.25, 0.75), (0.1, 0.2, 0.5, 0.9)], settings={"temperature": 0, "max_tokens": 4})
    assert isinstance(out["answer"], str) and len(_ids(out["answer"])) <= 4


def test_sampling_settings(api):
    """temperature / top_p (moondream.py:270-278, 312-318, 524-530): the default settings sample; the limit
    that makes sampling deterministic must reproduce the greedy tokens; a seed must reproduce itself."""
    from moondream_b200 import synth

    cfg, model, orc = api
    img = synth.synthetic_image(3, 378, 378)        # argmax margins of its first 5 caption tokens: 9..31 bf16 ulps
    enc = model.encode_image(img)
    tpl = cfg.tokenizer.templates["caption"]["short"]
    gen = orc.generate(orc.encode_image(img), tpl, 5)
    greedy = _ids(model.caption(enc, "short", settings={"temperature": 0, "max_tokens": 5})["caption"])
    _agree(greedy, gen, "greedy")
    # top_p -> 0 keeps only the most likely token (the one that crosses the threshold is kept)
    nucleus = _ids(model.caption(enc, "short", settings={"temperature": 1.0, "top_p": 1e-6, "max_tokens": 5})["caption"])
    _agree(nucleus, gen, "nucleus")
    assert nucleus == greedy, (nucleus, greedy)
    torch.manual_seed(7)
    a = _ids(model.caption(enc, "short", settings={"temperature": 1.5, "top_p": 0.95, "max_tokens": 8})["caption"])
    torch.manual_seed(7)
    b = _ids(model.caption(enc, "short", settings={"temperature": 1.5, "top_p": 0.95, "max_tokens": 8})["caption"])
    assert a == b and all(0 <= t < cfg.text.vocab_size for t in a), (a, b)
    out = model.caption(img, "short", settings={"max_tokens": 6})          # reference defaults: temperature 0.5, top_p 0.3
    assert isinstance(out["caption"], str)
    q = model.query(enc, "15 16", settings={"temperature": 0.7, "max_tokens": 5})
    assert isinstance(q["answer"], str) and len(_ids(q["answer"])) <= 5
    both = model.caption_batch([enc, img], "short", settings={"temperature": 0.7, "max_tokens": 4})
    assert len(both) == 2 and all(isinstance(o["caption"], str) for o in both)


def test_spatial_refs_match_the_reference_golden(api):
    """query(spatial_refs=...) against tests/golden/tiny_spatial_refs.json: the rows of the prompt embedding that carry
    the region encodings (placement and values) and the answer tokens."""
    import json
    import os

    from moondream_b200 import synth

    cfg, model, _ = api
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tiny_spatial_refs.json")))
    idx, h, w = gold["image"]
    enc = model.encode_image(synth.synthetic_image(idx, h, w))
    for c in gold["cases"]:
        refs = [tuple(r) for r in c["spatial_refs"]]
        assert model._query_prompt(c["question"], refs, False) == c["prompt"]
        emb = model._prompt_embeds_with_refs(c["prompt"], refs).float().cpu()
        want = torch.tensor(c["row_embeds"])
        got = emb[c["rows"]]
        assert (got - want).norm() / want.norm() < 2e-2, (refs, float((got - want).norm() / want.norm()))
        out = _ids(model.query(enc, c["question"], spatial_refs=refs, settings={"temperature": 0, "max_tokens": 8})["answer"])
        gen = type("G", (), {"tokens": c["tokens"], "margin_ulps": c["margin_ulps"]})()
        _agree(out, gen, "spatial refs")

"""Weight-only int8 groundwork (BASELINE.json config 5; SURVEY.md §8c: int8 parity = the bf16 path on the dequantised
weights).  CPU only: format invariants, error bound, and the drift the oracle sees on dequantised weights."""
import pytest
import torch

from moondream_b200 import config as C, quant, synth


def test_round_trip_bounds_and_edge_rows():
    g = torch.Generator().manual_seed(0)
    w = (torch.randn(96, 320, generator=g) * 0.05).to(torch.bfloat16)
    w[3] = 0                                          # an all-zero output feature
    w[5, 7] = 2.0                                     # an outlier sets that row's scale
    q, s = quant.quantize_weight_int8(w)
    assert q.dtype == torch.int8 and s.dtype == torch.bfloat16 and q.shape == w.shape and s.shape == (96,)
    assert int(q.abs().max()) <= 127 and bool((s.float() > 0).all())
    d = quant.dequantize_weight_int8(q, s)
    assert d.dtype == torch.bfloat16 and bool((d[3] == 0).all())
    # |w - w'| <= half a quantisation step + one bf16 rounding of the product (2^-9 relative)
    err = (w.float() - d.float()).abs()
    bound = 0.5 * s.float().unsqueeze(1) * (1 + 2 ** -7) + d.float().abs() * 2 ** -8
    assert bool((err <= bound).all()), float((err - bound).max())
    # every row uses (almost) its whole range: the row maximum quantises to +-127 or +-126 (bf16 scale rounding)
    rowmax = q.abs().amax(dim=1)
    assert bool((rowmax[torch.arange(96) != 3] >= 126).all())
    # idempotent: re-quantising the dequantised weights reproduces q and scale
    q2, s2 = quant.quantize_weight_int8(d)
    assert torch.equal(s2, s) and torch.equal(q2, q)
    with pytest.raises(ValueError):
        quant.quantize_weight_int8(torch.zeros(4))


def test_decoder_packing_and_stream_bytes():
    cfg = C.tiny()
    sd = synth.synthetic_state_dict(cfg, 0)
    packed, deq = quant.quantize_decoder_int8(cfg, sd)
    keys = list(quant.decode_stream_keys(cfg))
    assert set(packed) == set(keys) and len(keys) == 4 * cfg.text.n_layers + 1
    for k in sd:
        if k in packed:
            assert deq[k].shape == sd[k].shape and deq[k].dtype == torch.bfloat16 and not torch.equal(deq[k], sd[k])
        else:
            assert deq[k] is sd[k]                     # vision, embeddings, norms, biases, region head untouched
    big = C.preset("moondream-2b")
    bf16, i8 = quant.stream_bytes(big, False), quant.stream_bytes(big, True)
    assert abs(bf16 / 1e9 - 2.63) < 0.02               # 2.63 GB per decode step
    assert 0.50 < i8 / bf16 < 0.51


def test_oracle_drift_on_dequantised_weights():
    """The int8 oracle is the bf16 oracle on w'; report how far it moves from the unquantised model."""
    from oracle.moondream_oracle import OracleModel

    cfg = C.tiny()
    sd = synth.synthetic_state_dict(cfg, 0)
    _, deq = quant.quantize_decoder_int8(cfg, sd)
    a, b = OracleModel(cfg, sd), OracleModel(cfg, deq)
    img = synth.synthetic_image(3, 378, 378)
    prompt = synth.synthetic_prompt(3, 6, cfg.text.vocab_size)
    ea, eb = a.encode_image(img), b.encode_image(img)
    # the image prefix runs through the (quantised) decoder blocks: KV drifts, the vision tower does not
    k0a, k0b = ea.caches[0][0].float(), eb.caches[0][0].float()
    rel = float((k0a - k0b).norm() / k0a.norm())
    assert 0 < rel < 5e-2, rel
    la = a.prefill_prompt(prompt, ea.pos)[0].float()
    b.load_encoded(eb)
    lb = b.prefill_prompt(prompt, eb.pos)[0].float()
    assert float((la - lb).norm() / la.norm()) < 0.1

"""Weight-only quantised decoder weights on the GPU (reference layers.py:38-110).

Parity definition: the bf16 path on the DEQUANTISED weights W = bf16(bf16(q - zero) * scale).  The
kernels rebuild exactly W and keep the bf16 stream's split plan, so every comparison below is bit-exact:
  * md_dequantize_weights            == quant.dequantize (itself pinned to the reference's dequantize_tensor, CPU tests)
  * md_linear_small_batch_quant      == md_linear_small_batch_bf16 on W
  * Engine(quantize=...) / a reference-format int4 checkpoint == Engine on W: hidden states, KV, tokens, margins
and the oracle (CPU restatement of the reference) on W bounds the whole thing like every other model-level test."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _packed(bits, n_out, K, seed, awkward=False):
    from moondream_b200 import quant

    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(n_out, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    if bits == 4:
        nib, scale, zero = quant.quantize_weight_int4(w)
        if awkward:
            scale = scale * (1 + torch.rand(scale.shape, generator=g) * 1e-3)
            zero = zero + torch.rand(zero.shape, generator=g) - 0.5
        ql = quant.QuantLinear(4, nib, scale.float().contiguous(), zero.float().contiguous())
    else:
        q8, s8 = quant.quantize_weight_int8(w)
        groups = K // 128
        ql = quant.QuantLinear(8, q8.view(torch.uint8), s8.float().unsqueeze(1).repeat(1, groups).contiguous(),
                               torch.zeros(n_out, groups))
    return ql, ql.dequantized()


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("n_out,K,awkward", [(64, 128, False), (100, 256, True), (640, 1024, True), (2048, 2048, False)])
def test_dequantize_kernel_is_bit_exact(bits, n_out, K, awkward):
    from moondream_b200 import ops

    ql, want = _packed(bits, n_out, K, n_out + K + bits, awkward)
    dev = [t.cuda() for t in (ql.stream_bytes(), ql.scale, ql.zero)]
    out = torch.full((n_out, K + 64), 7.0, device="cuda", dtype=torch.bfloat16)      # row pitch wider than K
    ops.dequantize_weights(bits, *dev, n_out, K, out=out[:, :K])
    torch.cuda.synchronize()
    assert torch.equal(out[:, :K].cpu(), want)
    assert bool((out[:, K:] == 7.0).all())


STREAM_SHAPES = [  # (batch, n_out, K): 2B [qkv;fc1], fc2-like, 0.5B [qkv;fc1] at b128, ragged batch / rows, tiny
    (32, 14336, 2048), (32, 2048, 8192), (128, 7168, 1024), (70, 1000, 512), (5, 640, 128), (64, 4096, 256)]


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("batch,n_out,K", STREAM_SHAPES)
def test_quantised_stream_equals_bf16_stream_on_dequantised_weights(bits, batch, n_out, K):
    from moondream_b200 import ops

    ql, w = _packed(bits, n_out, K, batch + n_out + K + bits, awkward=(bits == 4))
    g = torch.Generator(device="cuda").manual_seed(batch + K)
    x = torch.randn(batch, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(n_out, device="cuda", generator=g).bfloat16()
    res = torch.randn(batch, n_out, device="cuda", generator=g).bfloat16()
    dev = [t.cuda() for t in (ql.stream_bytes(), ql.scale, ql.zero)]
    wd = w.cuda()
    for mode, r in ((0, None), (1, None), (2, res)):
        want = ops.linear_small_batch(x, wd, b, mode, r)
        got = ops.linear_small_batch_quant(bits, x, *dev, n_out, b, mode, r)
        torch.cuda.synchronize()
        assert torch.equal(got, want), f"mode {mode}: {(got.float() - want.float()).abs().max().item()}"
    # and the stream is a Linear: within one bf16 ulp of the fp32 product (accumulation order is the only freedom)
    ref = (x.float() @ wd.float().t() + b.float()).bfloat16().float()
    got = ops.linear_small_batch_quant(bits, x, *dev, n_out, b, 0, None).float()
    assert (got - ref).abs().max().item() <= ref.abs().max().item() * 2.0 ** -7


NEAR_TIE_ULPS = 4.5          # as in tests/test_model_parity_gpu.py


def _check_tokens(got, oracle_gen, what):
    """exact match, or first divergence at a step where the oracle itself is at a near-tie."""
    n = len(oracle_gen.tokens)
    for i in range(n):
        if got[i] != oracle_gen.tokens[i]:
            assert oracle_gen.margin_ulps[i] < NEAR_TIE_ULPS, \
                f"{what}: token {i} differs ({got[i]} vs {oracle_gen.tokens[i]}) at {oracle_gen.margin_ulps[i]:.1f} ulps"
            return i
    return n


@pytest.fixture(scope="module")
def tiny():
    from moondream_b200 import config as C, synth

    cfg = C.tiny()
    return cfg, synth.synthetic_state_dict(cfg, 0)


def _run(eng, cfg, n_tokens=20):
    from moondream_b200 import synth

    imgs = [synth.synthetic_image(40 + i, *hw) for i, hw in enumerate([(378, 378), (500, 700), (378, 378)])]
    prompts = [synth.synthetic_prompt(50 + i, 5 + 2 * i, cfg.text.vocab_size) for i in range(3)]
    prefixes, feats, img_emb, hidden = eng.encode_images(imgs, return_hidden=True)
    kv = [[(k.clone(), v.clone()) for k, v in eng.prefix_kv_tensors(p)] for p in prefixes]
    res = eng.generate(prefixes, prompts, max_tokens=n_tokens)
    torch.cuda.synchronize()
    return imgs, prompts, hidden.clone(), kv, res


@pytest.mark.parametrize("bits", [4, 8])
def test_quantised_engine_equals_bf16_engine_on_dequantised_weights(tiny, bits):
    from moondream_b200 import quant
    from moondream_b200.engine import Engine
    from oracle.moondream_oracle import OracleModel
    cfg, sd = tiny
    qt, deq = quant.quantize_decoder(cfg, sd, bits)
    eng_q = Engine(cfg, sd, max_batch=4, quantize="int4" if bits == 4 else "int8")
    assert eng_q.quantized is not None and eng_q.quantized.bits == bits
    eng_d = Engine(cfg, deq, max_batch=4)
    imgs, prompts, hid_q, kv_q, res_q = _run(eng_q, cfg)
    _, _, hid_d, kv_d, res_d = _run(eng_d, cfg)
    assert torch.equal(hid_q, hid_d)
    for a, b in zip(kv_q, kv_d):
        for (ka, va), (kb, vb) in zip(a, b):
            assert torch.equal(ka, kb) and torch.equal(va, vb)
    assert torch.equal(res_q.tokens, res_d.tokens)
    assert torch.equal(res_q.margins, res_d.margins)
    # the quantised model is a different model from the bf16 one (the comparison above is not vacuous) ...
    eng_b = Engine(cfg, sd, max_batch=4)
    _, _, hid_b, _, _ = _run(eng_b, cfg, n_tokens=1)
    assert not torch.equal(hid_b, hid_q)
    # ... and the oracle on the dequantised weights bounds it like any other model-level test
    orc = OracleModel(cfg, deq)
    exact = 0
    for i in range(3):
        o = orc.generate(orc.encode_image(imgs[i]), prompts[i], 20)
        exact += int(_check_tokens(res_q.tokens[i].tolist(), o, f"int{bits} image {i}") == len(o.tokens))
    assert exact >= 1


def test_reference_format_int4_checkpoint_loads_and_runs(tiny, tmp_path):
    """a state dict with `…weight.packed / .scale / .zero_point` entries (layers.py:58-76) goes through the loader
    and the engine and generates what the bf16 engine generates on the weights `QuantizedLinear.unpack` builds."""
    from moondream_b200 import quant, weights
    from moondream_b200.engine import Engine
    from oracle.moondream_oracle import dequantized_state_dict

    cfg, sd = tiny
    qt, deq = quant.quantize_decoder(cfg, sd, 4)
    ck = {k: v for k, v in sd.items() if deq[k] is sd[k]}
    ck.update(quant.reference_checkpoint_entries(cfg, qt))
    path = str(tmp_path / "int4.pt")
    torch.save(ck, path)
    loaded = weights.load_state_dict_from_file(path, cfg)
    assert quant.is_quantized_checkpoint(loaded)
    eng = Engine(cfg, loaded, max_batch=4)
    assert eng.quantized is not None and eng.quantized.bits == 4
    _, _, hid, _, res = _run(eng, cfg)
    ref_sd = dequantized_state_dict(ck)                       # the oracle's restatement of QuantizedLinear.unpack
    eng_d = Engine(cfg, {k: ref_sd[k] for k in sd}, max_batch=4)
    _, _, hid_d, _, res_d = _run(eng_d, cfg)
    assert torch.equal(hid, hid_d) and torch.equal(res.tokens, res_d.tokens)

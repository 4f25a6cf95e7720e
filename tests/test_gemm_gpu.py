"""Parity of the tcgen05 GEMM against a plain torch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b, mode, res=None, res_mod=0):
    y = x.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    y = y.bfloat16()
    if mode == 1:
        y = torch.nn.functional.gelu(y.float(), approximate="tanh").bfloat16()
    elif mode == 2:
        r = res
        if res_mod:
            idx = torch.arange(y.shape[0], device=y.device) % res_mod
            r = res[idx]
        y = (y.float() + r.float()).bfloat16()
    return y


def _close(got, ref, what):
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    # one bf16 ulp of the largest magnitude: accumulation order differs from cuBLAS, nothing else may
    assert err <= scale * 2.0 ** -7, f"{what}: max err {err} vs scale {scale}"
    frac = ((got - ref).abs() > 0).float().mean().item()
    assert frac < 0.05, f"{what}: {frac:.3f} of elements differ"


SHAPES = [
    (128, 256, 64), (128, 256, 128), (256, 512, 256), (1458, 1152, 592), (729, 3456, 1152),
    (1000, 4304, 1152), (517, 1152, 4304), (64, 64, 64), (200, 48, 72), (130, 2696, 720),
    (4096, 6144, 2048),
]


@pytest.fixture(params=[1, 2], ids=["cta1", "ctapair"])
def cta_group(request):
    """run every row-form case with single-CTA tiles and with cta_group::2 pairs (where BN = 256)"""
    from moondream_b200 import _native as N

    N.lib().md_debug_force_cta_group(request.param)
    yield request.param
    N.lib().md_debug_force_cta_group(0)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_linear_rowform(M, N, K, mode, cta_group):
    from moondream_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + mode)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16() if mode == 2 else None
    y = ops.linear(x, w, b, epilogue=mode, residual=res)
    torch.cuda.synchronize()
    _close(y, _ref(x, w, b, mode, res), f"linear {M}x{N}x{K} mode {mode}")


def test_linear_posemb_broadcast_and_remap(cta_group):
    from moondream_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    B, T, K, N = 3, 729, 592, 1152
    x = torch.randn(B * T, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    pos = torch.randn(T, N, device="cuda", generator=g).bfloat16()
    y = ops.linear(x, w, b, epilogue=2, residual=pos, res_mod=T)
    _close(y, _ref(x, w, b, 2, pos, T), "pos_emb broadcast")
    out = torch.zeros(B * (T + 1), N, device="cuda", dtype=torch.bfloat16)
    ops.linear(x, w, b, epilogue=0, out=out, remap=(T, T + 1, 1))
    ref = _ref(x, w, b, 0).view(B, T, N)
    got = out.view(B, T + 1, N)
    _close(got[:, 1:], ref, "remap rows")
    assert got[:, 0].abs().max().item() == 0


def test_linear_strided_views(cta_group):
    from moondream_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(9)
    big = torch.randn(300, 4096, device="cuda", generator=g).bfloat16()
    x = big[:, 1024:1024 + 2304]           # row pitch 4096, K = 2304
    w = (torch.randn(512, 2304, device="cuda", generator=g) / 48).bfloat16()
    y = ops.linear(x, w, None)
    _close(y, _ref(x, w, None, 0), "strided A")


SMALL = [(32, 6144, 2048), (32, 2048, 8192), (1, 2048, 2048), (5, 1024, 8192), (32, 51200, 2048),
         (64, 8192, 2048), (128, 3072, 1024), (7, 1032, 264), (200, 2048, 1024)]


@pytest.mark.parametrize("B,N,K", SMALL)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_linear_small_batch(B, N, K, mode):
    from moondream_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(B * 7 + N * 3 + K + mode)
    x = torch.randn(B, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(B, N, device="cuda", generator=g).bfloat16() if mode == 2 else None
    y = ops.linear_small_batch(x, w, b, epilogue=mode, residual=res)
    torch.cuda.synchronize()
    _close(y, _ref(x, w, b, mode, res), f"small-batch {B}x{N}x{K} mode {mode}")


@pytest.mark.parametrize("B,N,K", [(32, 6144, 2048), (5, 1024, 8192), (64, 8192, 2048), (17, 1032, 264)])
def test_linear_small_batch_forced_m128(B, N, K):
    """batches <= 64 run M = 64 MMAs by default (the other small-batch tests); bit 6 forces the M = 128 instantiation"""
    from moondream_b200 import _native as N_, ops

    g = torch.Generator(device="cuda").manual_seed(B + N + K)
    x = torch.randn(B, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    N_.lib().md_debug_gemm(64)
    try:
        y = ops.linear_small_batch(x, w, b, epilogue=0)
        torch.cuda.synchronize()
    finally:
        N_.lib().md_debug_gemm(0)
    _close(y, _ref(x, w, b, 0, None), f"small-batch M=128 {B}x{N}x{K}")

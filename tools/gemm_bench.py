"""Times the tcgen05 GEMM on the ViT / decoder shapes (CUDA events) and prints TFLOP/s / GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    out = []
    for (M, N, K, mode) in [(46656, 3456, 1152, 0), (46656, 1152, 1152, 2), (46656, 4304, 1152, 1),
                            (46656, 1152, 4304, 2), (23360, 6144, 2048, 0), (23360, 8192, 2048, 1),
                            (23360, 2048, 8192, 2), (23328, 8192, 2304, 1)]:
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        r = torch.randn(M, N, device="cuda").bfloat16()
        y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        from moondream_b200 import _native as N_
        res = {}
        for cg in (1, 2):
            N_.lib().md_debug_force_cta_group(cg)
            res[cg] = timeit(lambda: ops.linear(x, w, b, epilogue=mode, residual=r if mode == 2 else None, out=y))
        N_.lib().md_debug_force_cta_group(0)
        ms_t = timeit(lambda: torch.nn.functional.linear(x, w, b))
        fl = 2 * M * N * K / 1e9
        out.append({"M": M, "N": N, "K": K, "mode": mode, "cta1_tflops": fl / res[1], "ctapair_tflops": fl / res[2],
                    "torch_tflops": fl / ms_t})
        print(out[-1], flush=True)
    for (B, N, K) in [(32, 6144, 2048), (32, 2048, 2048), (32, 8192, 2048), (32, 2048, 8192),
                      (32, 51200, 2048), (128, 3072, 1024)]:
        x = torch.randn(B, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        y = torch.empty(B, N, device="cuda", dtype=torch.bfloat16)
        ws = torch.empty(16 * B * N, device="cuda", dtype=torch.float32)
        ms = timeit(lambda: ops.linear_small_batch(x, w, b, out=y, workspace=ws), iters=50)
        ms_t = timeit(lambda: torch.nn.functional.linear(x, w, b), iters=50)
        out.append({"B": B, "N": N, "K": K, "ms": ms, "gbs": N * K * 2 / ms / 1e6,
                    "torch_ms": ms_t, "torch_gbs": N * K * 2 / ms_t / 1e6})
        print(out[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()

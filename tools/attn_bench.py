"""A/B timing of the prefill attention kernels (tcgen05 single-pass persistent, two-pass, one item per CTA; legacy mma.sync) at the 2B bench shape."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import _native as N  # noqa: E402

lib = N.lib()
heads, layers, n_seqs, L = 32, 1, 32, 730
max_blocks = 32
n_pages = n_seqs * max_blocks
D = heads * 64
pool = torch.randn(layers, n_pages, 2, heads, 64, 64, device="cuda").bfloat16()
bt = torch.arange(n_seqs * max_blocks, dtype=torch.int32, device="cuda").view(n_seqs, max_blocks)
kv = N.md_kv(pool=pool.data_ptr(), n_pages=n_pages, block_tables=bt.data_ptr(), max_blocks=max_blocks, n_layers=layers)
T = n_seqs * L
q = torch.randn(T, D, device="cuda").bfloat16()
out = torch.empty_like(q)
qo = torch.arange(0, T + 1, L, dtype=torch.int32, device="cuda")
sp = torch.zeros(n_seqs, dtype=torch.int32, device="cuda")
flops = 4.0 * heads * n_seqs * L * L * 64
res = {}
for impl in (0, 1, 2, 4):
    lib.md_debug_attention_impl(impl)
    def run():
        N.check(lib.md_prefill_attention_bf16(N.ptr(q), heads, T, N.ptr(qo), N.ptr(sp), n_seqs, L, 730,
                                              ctypes.byref(kv), 0, N.ptr(out), N.current_stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        run()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    res[impl] = out.float().clone()
    print({"impl": {0: "tcgen05 single-pass softmax, persistent CTAs (default)", 1: "mma.sync", 2: "tcgen05 two-pass softmax", 4: "tcgen05 single-pass, one item per CTA"}[impl], "ms": ms, "tflops": flops / ms / 1e9}, flush=True)
print("rel diff vs mma.sync:", {i: ((res[i] - res[1]).norm() / res[1].norm()).item() for i in (0, 2, 4)})
lib.md_debug_attention_impl(0)
# occupancy probe
print("done")

# ---- ViT attention: 64 crops x 16 heads x 729 x 72 ----
n_crops, vh, seq = 64, 16, 729
Dv = vh * 72
qkv = torch.randn(n_crops * seq, 3 * Dv, device="cuda").bfloat16()
vout = torch.empty(n_crops * seq, Dv, device="cuda", dtype=torch.bfloat16)
vflops = 4.0 * vh * n_crops * seq * seq * 72
vres = {}
for impl in (0, 1, 2, 4):
    lib.md_debug_attention_impl(impl)
    def runv():
        N.check(lib.md_vit_attention_bf16(N.ptr(qkv), n_crops, seq, vh, N.ptr(vout), N.current_stream()))
    for _ in range(3):
        runv()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        runv()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    vres[impl] = vout.float().clone()
    print({"vit impl": {0: "tcgen05 single-pass softmax, persistent CTAs (default)", 1: "mma.sync", 2: "tcgen05 two-pass softmax", 4: "tcgen05 single-pass, one item per CTA"}[impl], "ms": ms, "tflops": vflops / ms / 1e9}, flush=True)
print("vit rel diff vs mma.sync:", {i: ((vres[i] - vres[1]).norm() / vres[1].norm()).item() for i in (0, 2, 4)})
lib.md_debug_attention_impl(0)

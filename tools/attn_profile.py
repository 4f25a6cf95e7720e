"""One flash-attention launch per implementation (after two warm-ups) for `ncu --set full` captures:
    ncu --set full --clock-control none --import-source on -k regex:fa_tc_ -f -o gpurun_out/fa python tools/attn_profile.py
runs prefill (32 seqs x 32 heads x 730 x 64) and ViT (64 crops x 16 heads x 729 x 72) with the default (single-pass softmax) and impl 2 (two-pass)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import _native as N  # noqa: E402

lib = N.lib()
impls = [int(a) for a in sys.argv[1:]] or [0, 2]
heads, n_seqs, L, max_blocks = 32, 32, 730, 32
n_pages = n_seqs * max_blocks
D = heads * 64
pool = torch.randn(1, n_pages, 2, heads, 64, 64, device="cuda").bfloat16()
bt = torch.arange(n_seqs * max_blocks, dtype=torch.int32, device="cuda").view(n_seqs, max_blocks)
kv = N.md_kv(pool=pool.data_ptr(), n_pages=n_pages, block_tables=bt.data_ptr(), max_blocks=max_blocks, n_layers=1)
T = n_seqs * L
q = torch.randn(T, D, device="cuda").bfloat16()
out = torch.empty_like(q)
qo = torch.arange(0, T + 1, L, dtype=torch.int32, device="cuda")
sp = torch.zeros(n_seqs, dtype=torch.int32, device="cuda")
n_crops, vh, seq = 64, 16, 729
Dv = vh * 72
qkv = torch.randn(n_crops * seq, 3 * Dv, device="cuda").bfloat16()
vout = torch.empty(n_crops * seq, Dv, device="cuda", dtype=torch.bfloat16)
for impl in impls:
    lib.md_debug_attention_impl(impl)
    for _ in range(3):
        N.check(lib.md_prefill_attention_bf16(N.ptr(q), heads, T, N.ptr(qo), N.ptr(sp), n_seqs, L, 730,
                                              ctypes.byref(kv), 0, N.ptr(out), N.current_stream()))
        N.check(lib.md_vit_attention_bf16(N.ptr(qkv), n_crops, seq, vh, N.ptr(vout), N.current_stream()))
    torch.cuda.synchronize()
lib.md_debug_attention_impl(0)
print("done")

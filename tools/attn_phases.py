"""Where a softmax thread of the prefill flash-attention kernel spends its clocks: one launch of the instrumented
instantiation (md_debug_attention_impl(3)) with the debug timeline installed; per-CTA clock sums of thread 64 (first
softmax thread), averaged over the CTAs.  32 sequences x 32 heads x 730 x 64."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import _native as N  # noqa: E402

lib = N.lib()
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


def run():
    N.check(lib.md_prefill_attention_bf16(N.ptr(q), heads, T, N.ptr(qo), N.ptr(sp), n_seqs, L, 730, ctypes.byref(kv), 0,
                                          N.ptr(out), N.current_stream()))


impl = 3                                                   # the default kernel with phase clocks
lib.md_debug_attention_impl(impl)
for _ in range(3):
    run()
torch.cuda.synchronize()
CAP = 1 << 16
rec = torch.zeros((CAP, 6), dtype=torch.int64, device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
assert lib.md_debug_timeline(rec.data_ptr(), cnt.data_ptr(), CAP) == 0
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
run()
e.record()
torch.cuda.synchronize()
assert lib.md_debug_timeline(None, None, 0) == 0
lib.md_debug_attention_impl(0)
n = int(cnt.item())
r = rec[:n].cpu().numpy().astype(np.int64)
kind = (r[:, 0] >> 60) & 0xF
a, b = r[kind == 5], r[kind == 6]
tiles = b[:, 3].astype(np.float64)
names = ["wait S (s_full)", "TMEM load of 128 scores + release S", "mask + row max", "wait previous P V (+ rescale)",
         "exponentials + pack + P stores"]
res = {"launch_ms": s.elapsed_time(e), "ctas": int(len(a)), "tiles_per_cta": float(tiles.mean()),
       "clocks_per_tile": {nm: float((a[:, 1 + i] / tiles).mean()) for i, nm in enumerate(names)}}
res["clocks_per_tile"]["fence + arrive"] = float((b[:, 1] / tiles).mean())
res["clocks_per_tile"]["whole softmax loop"] = float((b[:, 2] / tiles).mean())
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/attn_phases.json", "w"), indent=1)

"""Per-CTA %globaltimer timeline of the CUDA-graph decode step (md_debug_timeline).

CUDA events cannot see inside a graph replay and ncu serialises launches (no PDL overlap), so this is the
ground truth for where a decode step's time goes: for every launch of the four per-layer kernels it prints
when the first CTA started, when the dependency wait released, and when the last CTA finished.

    python tools/decode_timeline.py [--batch 32] [--out gpurun_out/decode_timeline.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, synth  # noqa: E402
from moondream_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--out", default="gpurun_out/decode_timeline.json")
ap.add_argument("--gemm-debug", type=int, default=0, help="md_debug_gemm flags (timing experiments)")
ap.add_argument("--brief", action="store_true")
args = ap.parse_args()

B = args.batch
cfg = C.preset("moondream-2b")
sd = synth.synthetic_state_dict(cfg, 0)
images = [synth.synthetic_image(i, 378, 378) for i in range(B)]
prompts = [synth.synthetic_prompt(i, 32, cfg.text.vocab_size) for i in range(B)]
eng = Engine(cfg, sd, max_batch=B)
eng.lib.md_debug_gemm(args.gemm_debug)
pre = eng.encode_images(images)
eng.generate(pre, prompts, 8, stop_on_eos=False, to_host=False)          # capture
for _ in range(3):                                                       # settle clocks before measuring
    pre = eng.encode_images(images)
    eng.generate(pre, prompts, 64, stop_on_eos=False, to_host=False)
pre = eng.encode_images(images)
torch.cuda.synchronize()

CAP = 1 << 20
rec = torch.zeros((CAP, 6), dtype=torch.int64, device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
assert eng.lib.md_debug_timeline(rec.data_ptr(), cnt.data_ptr(), CAP) == 0
eng.generate(pre, prompts, args.steps, stop_on_eos=False, to_host=False)
torch.cuda.synchronize()
assert eng.lib.md_debug_timeline(None, None, 0) == 0
n = int(cnt.item())
print(f"{n} records", flush=True)
r = rec[: min(n, CAP)].cpu().numpy().astype(np.int64)
tag = (r[:, 0] >> 32) & 0xFFFFFFFF
kind = tag >> 28
mode = (tag >> 24) & 0xF
rows = tag & 0xFFFFFF

D, FF = cfg.text.dim, cfg.text.ff_dim
sel = {
    "gemm1": (kind == 1) & (mode == 3) & (rows == 3 * D + FF),
    "gemm2": (kind == 1) & (mode == 3) & (rows == D),
    "attn": (kind == 2) & ((tag & 1) == 1),
    "epi": kind == 3,
}
launches = []
for name, m in sel.items():
    x = r[m]
    if len(x) == 0:
        continue
    x = x[np.argsort(x[:, 1])]
    if name == "attn":                                        # fixed grid; may run in more than one wave
        cuts = np.arange(B * cfg.text.n_heads, len(x), B * cfg.text.n_heads)
    else:
        cuts = np.where(np.diff(x[:, 1]) > 8000)[0] + 1      # launches of one kind are >= 30 us apart
    for g in np.split(x, cuts):
        launches.append({
            "kind": name, "ctas": int(len(g)),
            "entry_first": int(g[:, 1].min()), "entry_last": int(g[:, 1].max()),
            "wait_first": int(g[:, 2].min()), "wait_last": int(g[:, 2].max()),
            "mid0_first": int(g[:, 3].min()), "mid0_last": int(g[:, 3].max()),
            "mid1_first": int(g[:, 4].min()) if name != "epi" else 0, "mid1_last": int(g[:, 4].max()),
            "exit_first": int(g[:, 5].min()), "exit_last": int(g[:, 5].max()),
            "cta_busy_mean": float((g[:, 5] - g[:, 2]).mean()),
        })
launches.sort(key=lambda d: d["wait_first"])
# the last decode step = the last 4 * n_layers launches (the lm_head GEMM has other row counts)
per_step = 4 * cfg.text.n_layers
last = launches[-per_step:]
t0 = last[0]["entry_first"]
print(f"last decode step: {len(last)} launches, span {(last[-1]['exit_last'] - last[0]['wait_first']) / 1000:.1f} us "
      f"(+ lm_head / argmax / advance outside)")
print("times in us relative to the step's first entry; wait = dependency (griddepcontrol.wait) released")
print(f"{'kind':6s} {'ctas':>5s} {'entry':>8s} {'wait0':>8s} {'wait1':>8s} {'mid0a':>8s} {'mid0b':>8s} {'mid1b':>8s} {'exit0':>8s} {'exit1':>8s} {'busy':>7s}")
for d in ([] if args.brief else last[4 * 10: 4 * 12]):
    f = lambda k: (d[k] - t0) / 1000.0  # noqa: E731
    print(f"{d['kind']:6s} {d['ctas']:5d} {f('entry_first'):8.1f} {f('wait_first'):8.1f} {f('wait_last'):8.1f} {f('mid0_first'):8.1f} "
          f"{f('mid0_last'):8.1f} {f('mid1_last'):8.1f} {f('exit_first'):8.1f} {f('exit_last'):8.1f} {d['cta_busy_mean'] / 1000:7.1f}")

g1 = [d["wait_first"] for d in last if d["kind"] == "gemm1"]
print(f"layer period {np.mean(np.diff(g1)) / 1000:.2f} us")
# averages over the steady-state launches of the last step
summary = {}
by_kind = {k: [d for d in last if d["kind"] == k] for k in sel}
order = ["gemm1", "attn", "gemm2", "epi"]
seq = [d for d in last]
for i, d in enumerate(seq):
    prev_exit = seq[i - 1]["exit_last"] if i else None
    d["gap_after_prev_exit"] = None if prev_exit is None else d["wait_last"] - prev_exit
for k in order:
    ds = by_kind[k]
    if not ds:
        continue
    summary[k] = {
        "launches": len(ds),
        "ctas": ds[0]["ctas"],
        "wait_to_exit_us": float(np.mean([d["exit_last"] - d["wait_first"] for d in ds])) / 1000,
        "wait_spread_us": float(np.mean([d["wait_last"] - d["wait_first"] for d in ds])) / 1000,
        "exit_spread_us": float(np.mean([d["exit_last"] - d["exit_first"] for d in ds])) / 1000,
        "entry_before_wait_us": float(np.mean([d["wait_first"] - d["entry_first"] for d in ds])) / 1000,
        "first_data_after_wait_us": float(np.mean([d["mid0_last"] - d["wait_last"] for d in ds])) / 1000,
        "acc_ready_before_exit_us": float(np.mean([d["exit_last"] - d["mid1_last"] for d in ds])) / 1000,
        "wait_last_after_prev_exit_us": float(np.mean([d["gap_after_prev_exit"] for d in ds if d["gap_after_prev_exit"] is not None])) / 1000,
        "cta_busy_mean_us": float(np.mean([d["cta_busy_mean"] for d in ds])) / 1000,
    }
    print(k, json.dumps(summary[k]) if not args.brief else
          {q: round(summary[k][q], 2) for q in ("wait_to_exit_us", "cta_busy_mean_us", "wait_last_after_prev_exit_us")})
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump({"batch": B, "summary": summary, "last_step": last}, open(args.out, "w"), indent=1)

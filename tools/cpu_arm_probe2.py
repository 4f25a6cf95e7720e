"""Follow-up to tools/cpu_arm_probe.py: the oracle's 2B decode runs at ~0.04 s/token inside bench.py's GPU arm and at
~0.25 s/token in a process that never touched CUDA, on the same cores.  One process: time 8 decode steps, profile
them, initialise CUDA (no engine, no kernels of this repo), time again, then allocate pinned memory and time again."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (NUMA placement + OpenMP pool size exactly as the CPU legs have them)
import torch  # noqa: E402

from moondream_b200 import config as C, synth  # noqa: E402
from oracle.moondream_oracle import OracleModel  # noqa: E402

torch.set_num_threads(bench.cpu_threads())
cfg = C.preset("moondream-2b")
sd = synth.synthetic_state_dict(cfg, 0)
orc = OracleModel(cfg, sd)
img = synth.synthetic_image(0, 378, 378)
prompt = synth.synthetic_prompt(0, 32, cfg.text.vocab_size)
enc = orc.encode_image(img)
orc.load_encoded(enc)
_, _, nxt, pos = orc.prefill_prompt(prompt, enc.pos)
state = {"tok": int(nxt.item()), "pos": pos}


def decode(n):
    t0 = time.perf_counter()
    for _ in range(n):
        logits, _ = orc.decode_one(orc.embed(torch.tensor([[state["tok"]]])), state["pos"])
        state["pos"] += 1
        state["tok"] = int(torch.argmax(logits, dim=-1).item())
    return (time.perf_counter() - t0) / n


out = {"placement": bench.CPU_PLACEMENT, "threads": torch.get_num_threads()}
out["before_cuda_s_per_token"] = [round(decode(8), 4) for _ in range(2)]
try:
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU]) as prof:
        decode(2)
    rows = sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:8]
    out["top_ops_before"] = [(e.key, round(e.self_cpu_time_total / 1e3, 1), e.count) for e in rows]
except Exception as e:  # noqa: BLE001
    out["profile_error"] = repr(e)[:200]
t0 = time.perf_counter()
torch.cuda.init()
x = torch.zeros(1, device="cuda")
torch.cuda.synchronize()
out["cuda_init_s"] = round(time.perf_counter() - t0, 2)
out["after_cuda_init_s_per_token"] = [round(decode(8), 4) for _ in range(2)]
pinned = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
out["after_pinned_alloc_s_per_token"] = [round(decode(8), 4) for _ in range(2)]
try:
    with profile(activities=[ProfilerActivity.CPU]) as prof:
        decode(2)
    rows = sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:8]
    out["top_ops_after"] = [(e.key, round(e.self_cpu_time_total / 1e3, 1), e.count) for e in rows]
except Exception as e:  # noqa: BLE001
    out["profile_error2"] = repr(e)[:200]
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_arm_probe2.json"), "w"), indent=1)

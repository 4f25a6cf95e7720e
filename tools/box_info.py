"""Prints what the GPU box's host looks like (CPU quota, threads) and times a small oracle run."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads", torch.get_num_threads())
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try:
        print(f, open(f).read().strip())
    except Exception as e:
        print(f, "n/a")
os.system("grep -m1 'model name' /proc/cpuinfo; grep -m1 -o -E 'avx512_bf16|amx_bf16' /proc/cpuinfo | sort -u; free -g | head -2")
from moondream_b200 import config as C, synth  # noqa: E402
from oracle.moondream_oracle import OracleModel  # noqa: E402

cfg = C.tiny()
sd = synth.synthetic_state_dict(cfg, 0)
img = synth.synthetic_image(0, 500, 700)
for nt in (torch.get_num_threads(), 8, 1):
    torch.set_num_threads(nt)
    o = OracleModel(cfg, sd)
    t = time.time(); enc = o.encode_image(img); t1 = time.time() - t
    t = time.time(); o.generate(enc, [11, 12, 13], 8); t2 = time.time() - t
    print(f"threads {nt}: oracle encode {t1:.2f}s, 8 tokens {t2:.2f}s", flush=True)

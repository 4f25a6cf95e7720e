"""Why is the oracle's CPU decode 5x faster inside bench.py's GPU arm than in a pure-CPU process?  Checks which ISA oneDNN
dispatches for the bf16 matvec of a decode step in (a) a plain process, (b) a process that asked the kernel for the AMX
tile-data permission before importing torch, (c) a process that initialised CUDA first."""
import ctypes
import os
import subprocess
import sys
import time

CHILD = r'''
import ctypes, os, sys, time
mode = sys.argv[1]
if mode == "amx":
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.syscall(158, 0x1023, 18)          # arch_prctl(ARCH_REQ_XCOMP_PERM, XFEATURE_XTILEDATA)
    print("arch_prctl rc", rc, "errno", ctypes.get_errno(), flush=True)
import torch
if mode == "cuda":
    torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
torch.set_num_threads(int(os.environ.get("OMP_NUM_THREADS", "16")))
w = [torch.randn(8192, 2048).bfloat16() for _ in range(24)]
x = torch.randn(1, 2048).bfloat16()
for _ in range(2):
    for m in w: torch.nn.functional.linear(x, m)
t0 = time.perf_counter()
for _ in range(5):
    for m in w: torch.nn.functional.linear(x, m)
dt = (time.perf_counter() - t0) / 5
print(mode, "24 x linear(1x2048 @ 8192x2048 bf16):", round(dt * 1e3, 2), "ms ->", round(24 * 8192 * 2048 * 2 / dt / 1e9, 1), "GB/s", flush=True)
'''
flags = open("/proc/cpuinfo").read()
print("cpu flags: amx_bf16" if "amx_bf16" in flags else "cpu flags: no amx_bf16", "| avx512_bf16" if "avx512_bf16" in flags else "| no avx512_bf16")
n = str(len(os.sched_getaffinity(0)))
try:
    q = open("/sys/fs/cgroup/cpu.max").read().split()
    if q[0] != "max":
        n = str(max(1, int(int(q[0]) / int(q[1]))))
except Exception:
    pass
for mode in ("plain", "amx", "cuda", "plain"):
    env = dict(os.environ, OMP_NUM_THREADS=n, ONEDNN_VERBOSE="1" if mode != "cuda" else "0")
    r = subprocess.run([sys.executable, "-c", CHILD, mode], capture_output=True, text=True, env=env, timeout=300)
    lines = (r.stdout + r.stderr).splitlines()
    info = [ln for ln in lines if "isa:" in ln or "info,cpu" in ln][:3]
    prim = [ln for ln in lines if ",exec," in ln and ("matmul" in ln or "inner_product" in ln)][:2]
    tail = [ln for ln in lines if ln.startswith(mode) or "arch_prctl" in ln]
    print(f"== {mode} (OMP {n})"); [print("  ", ln[:200]) for ln in info + prim + tail]

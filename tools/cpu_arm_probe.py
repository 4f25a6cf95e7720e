"""Why does the CPU arm's decode run at 0.05 s/token in one process and 0.38 s/token in another on the same box
(measurement defect 9)?  Runs the oracle's 2B decode in fresh subprocesses under different OpenMP
settings and prints seconds per token, encode seconds and the cgroup's CFS throttling counters around each run.

    python tools/cpu_arm_probe.py            # parent: caches the synthetic weights in /dev/shm, runs the matrix
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = "/dev/shm/md_probe_sd.pt"


def usable_cpus():
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def cpu_stat():
    try:
        d = dict(l.split() for l in open("/sys/fs/cgroup/cpu.stat").read().strip().splitlines())
        return {k: int(d[k]) for k in ("nr_periods", "nr_throttled", "throttled_usec", "usage_usec") if k in d}
    except Exception:
        return {}


def child():
    sys.path.insert(0, ROOT)
    placement = "unpinned"
    if os.environ.get("PROBE_PIN"):
        import bench                       # pins to one NUMA node at import (bench.pin_to_numa_node)

        placement = bench.CPU_PLACEMENT
    import torch

    from moondream_b200 import config as C, synth
    from oracle.moondream_oracle import OracleModel

    threads = int(os.environ["PROBE_THREADS"])
    torch.set_num_threads(threads)
    cfg = C.preset("moondream-2b")
    sd = torch.load(CACHE, mmap=True)
    if os.environ.get("PROBE_COPY"):       # private copies, first-touched by this process (like bench.py's own weights)
        sd = {k: v.clone() for k, v in sd.items()}
    orc = OracleModel(cfg, sd)
    img = synth.synthetic_image(0, 378, 378)
    prompt = synth.synthetic_prompt(0, 32, cfg.text.vocab_size)
    out = []
    for rep in range(2):
        s0 = cpu_stat()
        t0 = time.perf_counter()
        enc = orc.encode_image(img)
        t1 = time.perf_counter()
        orc.load_encoded(enc)
        _, _, nxt, pos = orc.prefill_prompt(prompt, enc.pos)
        tok = int(nxt.item())
        t2 = time.perf_counter()
        for _ in range(16):
            logits, _ = orc.decode_one(orc.embed(torch.tensor([[tok]])), pos)
            pos += 1
            tok = int(torch.argmax(logits, dim=-1).item())
        t3 = time.perf_counter()
        s1 = cpu_stat()
        out.append({"encode_s": round(t1 - t0, 3), "decode_s_per_token": round((t3 - t2) / 16, 4),
                    "throttled_periods": s1.get("nr_throttled", 0) - s0.get("nr_throttled", 0),
                    "throttled_ms": (s1.get("throttled_usec", 0) - s0.get("throttled_usec", 0)) / 1e3,
                    "cpu_s_used": (s1.get("usage_usec", 0) - s0.get("usage_usec", 0)) / 1e6,
                    "wall_s": round(t3 - t0, 2)})
    print("PROBE " + json.dumps({"threads": threads, "torch_threads": torch.get_num_threads(), "placement": placement,
                                 "runs": out}))


def main():
    if os.environ.get("PROBE_CHILD"):
        return child()
    sys.path.insert(0, ROOT)
    import torch

    from moondream_b200 import config as C, synth

    n = usable_cpus()
    try:
        cpu_max = open("/sys/fs/cgroup/cpu.max").read().strip()
    except OSError:
        cpu_max = "n/a"
    print(f"affinity {len(os.sched_getaffinity(0))} cpus, cgroup cpu.max {cpu_max!r}, usable {n}", flush=True)
    if not os.path.exists(CACHE):
        torch.save(synth.synthetic_state_dict(C.preset("moondream-2b"), 0), CACHE)
    cases = [
        ("private weights, pinned to one NUMA node", {"OMP_NUM_THREADS": str(n), "PROBE_PIN": "1", "PROBE_COPY": "1"}, n),
        ("private weights, unpinned", {"OMP_NUM_THREADS": str(n), "PROBE_COPY": "1"}, n),
        ("private weights, pinned (second process)", {"OMP_NUM_THREADS": str(n), "PROBE_PIN": "1", "PROBE_COPY": "1"}, n),
        ("private weights, unpinned (second process)", {"OMP_NUM_THREADS": str(n), "PROBE_COPY": "1"}, n),
        ("omp=usable (round-1 reference arm)", {"OMP_NUM_THREADS": str(n), "MKL_NUM_THREADS": str(n)}, n),
        ("omp=usable, passive wait", {"OMP_NUM_THREADS": str(n), "OMP_WAIT_POLICY": "passive"}, n),
        ("omp=usable, GOMP_SPINCOUNT=0", {"OMP_NUM_THREADS": str(n), "GOMP_SPINCOUNT": "0"}, n),
        ("omp=usable-2", {"OMP_NUM_THREADS": str(max(1, n - 2))}, max(1, n - 2)),
        ("omp=usable-2, passive wait", {"OMP_NUM_THREADS": str(max(1, n - 2)), "OMP_WAIT_POLICY": "passive"}, max(1, n - 2)),
        ("omp=usable/2", {"OMP_NUM_THREADS": str(max(1, n // 2))}, max(1, n // 2)),
        ("omp unset, set_num_threads(usable)", {}, n),
    ]
    results = []
    if os.environ.get("PROBE_ONLY"):
        cases = cases[: int(os.environ["PROBE_ONLY"])]
    for name, env, threads in cases:
        e = dict(os.environ)
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OMP_WAIT_POLICY", "GOMP_SPINCOUNT", "KMP_BLOCKTIME",
                  "PROBE_PIN", "PROBE_COPY"):
            e.pop(k, None)
        e.update(env)
        e["PROBE_CHILD"] = "1"
        e["PROBE_THREADS"] = str(threads)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=e, capture_output=True, text=True,
                               timeout=240)
            line = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            rec = json.loads(line[-1][6:]) if line else {"error": (r.stderr or r.stdout)[-400:]}
        except subprocess.TimeoutExpired:
            rec = {"error": "timeout"}
        rec["case"] = name
        results.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "cpu_arm_probe.json"), "w"), indent=1)
    try:
        os.remove(CACHE)
    except OSError:
        pass


if __name__ == "__main__":
    main()

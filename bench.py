#!/usr/bin/env python
"""bench.py — images/sec for encode + 64-token greedy caption, Moondream-2B, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--model moondream-2b]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input per GPU: 32 images of
378x378 (2 crops each) -> ViT -> stitch/pool/project -> [BOS; image] prefill (730 tokens) ->
32-token prompt prefill -> 64 greedy decode steps (CUDA graph) -> token ids.  Weights are seeded
synthetic tensors of the real architecture (no network, no checkpoint).

`value`   : whole-job images/s, inputs (uint8 crops, prompt ids) resident in HBM, CUDA-event timed.
`e2e`     : the same metric through the public host-buffer call (host uint8 images -> crop -> pinned
            H2D -> engine -> D2H token ids), wall clock around a synchronised region.
`roofline`: the dominant kernel (tcgen05 row-form GEMM) timed live with CUDA events on its stream.
`cpu_baseline` / `--impl reference`: the oracle port of the reference (bf16 torch CPU, the reference's
            own arithmetic) on the box's host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)



def _usable_cpus_early() -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


# Size the OpenMP / oneDNN pools for the CPUs this container may really use BEFORE torch creates them: the
# host shows 128 cores under a 16-CPU cgroup quota, and an oversubscribed pool makes the CPU arm several times
# slower (decode 0.32 vs 0.05 s/token measured).  torchrun's own OMP_NUM_THREADS=1 default is overridden for rank 0
# of the reference arm only.
if "--impl" in sys.argv and "reference" in sys.argv:
    os.environ["OMP_NUM_THREADS"] = str(_usable_cpus_early())
    os.environ["MKL_NUM_THREADS"] = os.environ["OMP_NUM_THREADS"]
else:
    os.environ.setdefault("OMP_NUM_THREADS", str(_usable_cpus_early()))

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH_PER_GPU = 32
PROMPT_LEN = 32
NEW_TOKENS = 64
IMAGE_HW = (378, 378)
METRIC = "images/sec (encode+64-tok greedy caption) Moondream-2B b32"


def usable_cpus() -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (sustained bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (mx or 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic():
    """Mean DRAM read+write bytes per launch of the dominant GEMM class from the committed `ncu --set full`
    capture (profiles/r01_ncu_full_summary.json); None when absent."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")
    try:
        return float(json.load(open(path))["gemm_mean_traffic_bytes"])
    except Exception:
        return None


def flops_per_image(cfg, n_crops: int) -> float:
    """Algorithmic FLOPs: 2*M*K*N per GEMM, 4*H*Tq*Tk*hd per attention."""
    v, t = cfg.vision, cfg.text
    T = v.tokens_per_crop
    vit = 2 * T * v.patch_dim * v.enc_dim + v.enc_n_layers * (
        2 * T * v.enc_dim * (4 * v.enc_dim + 2 * v.enc_ff_dim) + 4 * v.enc_n_heads * T * T * v.head_dim)
    proj = 2 * T * (2 * v.enc_dim * v.proj_inner_dim + v.proj_inner_dim * t.dim)
    per_tok = t.n_layers * 2 * (t.dim * 3 * t.dim + t.dim * t.dim + 2 * t.dim * t.ff_dim)
    pre = t.prefix_attn
    img_prefill = pre * per_tok + t.n_layers * 4 * t.n_heads * pre * pre * t.head_dim
    prompt = PROMPT_LEN * per_tok + t.n_layers * 4 * t.n_heads * PROMPT_LEN * (pre + PROMPT_LEN) * t.head_dim
    decode = NEW_TOKENS * (per_tok + 2 * t.dim * t.vocab_size)
    return n_crops * vit + proj + img_prefill + prompt + decode + 2 * t.dim * t.vocab_size


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# --------------------------------------------------------------------------------------------
def cpu_sample(cfg, sd, threads: int, decode_tokens: int = 16):
    """One image through the oracle: full encode + 32-token prompt prefill + `decode_tokens` decode
    steps, decode extrapolated to 64 tokens (stated in `sample`).  Returns (images/s, detail)."""
    from moondream_b200 import synth
    from oracle.moondream_oracle import OracleModel

    torch.set_num_threads(threads)
    orc = getattr(cpu_sample, "_model", None)
    if orc is None:
        orc = OracleModel(cfg, sd)
        cpu_sample._model = orc
    img = synth.synthetic_image(0, *IMAGE_HW)
    prompt = synth.synthetic_prompt(0, PROMPT_LEN, cfg.text.vocab_size)
    t0 = time.perf_counter()
    enc = orc.encode_image(img)
    t1 = time.perf_counter()
    orc.load_encoded(enc)
    _, _, nxt, pos = orc.prefill_prompt(prompt, enc.pos)
    t2 = time.perf_counter()
    tok = int(nxt.item())
    for _ in range(decode_tokens):
        logits, _ = orc.decode_one(orc.embed(torch.tensor([[tok]])), pos)
        pos += 1
        tok = int(torch.argmax(logits, dim=-1).item())
    t3 = time.perf_counter()
    per_image = (t1 - t0) + (t2 - t1) + (t3 - t2) * (NEW_TOKENS / decode_tokens)
    detail = {"encode_s": round(t1 - t0, 3), "prompt_prefill_s": round(t2 - t1, 3),
              "decode_s_per_token": round((t3 - t2) / decode_tokens, 4)}
    return 1.0 / per_image, detail


def run_reference(args, cfg, sd):
    threads = usable_cpus()
    for _ in range(args.warmup):
        cpu_sample(cfg, sd, threads)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, detail = cpu_sample(cfg, sd, threads)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = (f"1 image per step: full encode_image + {PROMPT_LEN}-token prompt prefill + 16 of {NEW_TOKENS} "
              f"decode steps, decode time scaled x4; oracle port of the reference (bf16 torch CPU); {detail}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model} caption: {IMAGE_HW[0]}x{IMAGE_HW[1]} image (2 crops), "
                               f"{PROMPT_LEN}-token prompt, {NEW_TOKENS} greedy tokens; sequential batch-1 "
                               f"(the reference cannot batch, moondream.py:66)"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# --------------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------------
def run_main(args, cfg, sd, rank, world, local_rank):
    import torch.distributed as dist

    from moondream_b200 import synth
    from moondream_b200 import _native as N
    from moondream_b200.engine import Engine
    from moondream_b200.image_crops import overlap_crop_image

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    eng = Engine(cfg, sd, device=dev, max_batch=B)
    lib = eng.lib
    v = cfg.vision
    # synthetic inputs; rank r owns images [r*B, (r+1)*B)
    images = [synth.synthetic_image(rank * B + i, *IMAGE_HW) for i in range(B)]
    prompts = [synth.synthetic_prompt(rank * B + i, PROMPT_LEN, cfg.text.vocab_size) for i in range(B)]
    crops, offsets, tilings = [], [0], []
    for im in images:
        oc = overlap_crop_image(im, overlap_margin=v.overlap_margin, max_crops=v.max_crops)
        crops.append(oc["crops"]); tilings.append(oc["tiling"]); offsets.append(offsets[-1] + oc["crops"].shape[0])
    crops_host = torch.from_numpy(np.concatenate(crops, 0)).pin_memory()
    crops_dev = crops_host.to(dev)
    n_crops = crops_dev.shape[0]
    gathered = torch.empty((world * B, NEW_TOKENS + 1), dtype=torch.int32, device=dev) if world > 1 else None

    def step_resident():
        # ViT -> stitch/pool/project -> [BOS; image; prompt] prefill (one pass) -> first token -> decode loop
        res = eng.caption_from_crops(crops_dev, offsets, tilings, prompts, NEW_TOKENS, to_host=False, stop_on_eos=False)
        if world > 1:   # the one collective of the path: finished token ids over NVLink
            dist.all_gather_into_tensor(gathered, res.tokens)
        return res

    def step_e2e():
        dev, offs, til = eng.stage_images(images)      # host crop (PIL) -> pinned staging -> H2D
        return eng.caption_from_crops(dev, offs, til, prompts, NEW_TOKENS, to_host=True, stop_on_eos=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step_resident()
    barrier()
    lib.md_reset_launch_count()
    lib.md_profile_linear(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = step_resident()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = int(lib.md_launch_count())
    g_ms, g_fl, g_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    N.check(lib.md_profile_linear_read(ctypes.byref(g_ms), ctypes.byref(g_fl), ctypes.byref(g_n)), "profile")
    lib.md_profile_linear(0)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B * args.steps / (ms_max / 1000.0)

    # e2e: host buffers in, token ids out, wall clock around a synchronised region
    if args.no_e2e:
        out = step_e2e() if False else None
    else:
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(0 if args.no_e2e else args.steps):
        out = step_e2e()
    barrier()
    e2e_s = max(time.perf_counter() - t0, 1e-9)
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t_e.item())
    h2d = int(crops_host.numel() + B * PROMPT_LEN * 4 + 3 * 4 * (B + 1) + B * eng.max_blocks * 4)
    d2h = int(out.tokens.numel() * 4 + out.margins.numel() * 4) if out is not None else 0

    if rank != 0:
        return
    peaks = measured_peaks()
    achieved = (g_fl.value / 1e12) / (g_ms.value / 1e3) if g_ms.value > 0 else 0.0
    total_flops = flops_per_image(cfg, n_crops // B) * B * args.steps
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model} caption b{B}/GPU: {B} images {IMAGE_HW[0]}x{IMAGE_HW[1]} (2 crops each, "
                               f"{n_crops} crops), {PROMPT_LEN}-token prompts, {NEW_TOKENS} greedy tokens "
                               f"(+ the reference's trailing decode step)",
                   "global_batch": world * B, "parallelism": f"dp{world} (images sharded, weights replicated)",
                   "l2": "per-step working set (weights 3.9 GB + activations) exceeds the 126 MB L2",
                   "params": synth.param_count(cfg)},
        "clocks": clocks,
        "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05 row-form GEMM: ViT, projection, prefill)",
                     "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops"] if peaks["bf16_tflops"] else None, "traffic": ncu_traffic(),
                     "traffic_note": "bytes per launch, mean of 4 ViT GEMM launches (ncu --set full, profiles/r01_ncu_full_summary.json); "
                                     "algorithmic bytes of the same launches: 438 / 325 / 519 / 642 MB",
                     "peak_source": peaks["source"], "launches": int(g_n.value),
                     "share_of_step": (g_ms.value / ms) if ms > 0 else None,
                     "step_model_tflops": total_flops / 1e12 / (ms_max / 1e3)},
    }
    try:   # the reference's eager torch-CUDA path on the same GPU model (separate run, tools/torch_cuda_comparator.py)
        comp = json.load(open(os.path.join(ROOT, "profiles", "r01_torch_cuda_comparator.json")))
        line["comparators"] = {"reference_torch_cuda_eager_images_per_s": comp["images_per_s"],
                               "source": "profiles/r01_torch_cuda_comparator.json (oracle port of the reference on cuda, "
                                         "batch-1 sequential, measured in a separate run on this pool's B200)"}
    except Exception:
        pass
    if world == 1 and not args.no_cpu_baseline:
        threads = usable_cpus()
        val, detail = cpu_sample(cfg, sd, threads)
        line["cpu_baseline"] = {
            "value": val, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"1 image: full encode_image + {PROMPT_LEN}-token prompt prefill + 16 of {NEW_TOKENS} decode "
                      f"steps (decode time x4); oracle port of the reference (bf16 torch CPU, sequential batch-1); {detail}"}
    emit_line(line)


_JSON_OUT = None


def protect_stdout():
    """stdout carries exactly one JSON line: keep a private handle on it and point fd 1 at stderr, so banners that
    libraries write straight to fd 1 (NCCL prints its version there whatever NCCL_DEBUG_FILE says) cannot reach it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global NEW_TOKENS
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="moondream-2b")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only (ncu launch lists)")
    ap.add_argument("--new-tokens", type=int, default=NEW_TOKENS,
                    help="profiling runs only: anything but 64 is not the BASELINE.json metric")
    args = ap.parse_args()
    NEW_TOKENS = args.new_tokens

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0                                      # rank 0 alone runs the CPU arm

    from moondream_b200 import config as C, synth

    cfg = C.preset(args.model)
    sd = synth.synthetic_state_dict(cfg, 0)
    if args.impl == "reference":
        run_reference(args, cfg, sd)
        return 0
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_main(args, cfg, sd, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Experiment, NOT part of the product path; run on a B200 this round
(profiles/r02_overlap_probe.json: +4 % at best, dropped):
does batch i's decode overlap with batch i+1's ViT + prefill when the two run on different streams?

Two Engine instances (own KV pools, own workspaces, same weights uploaded twice) alternate batches, each on its own
CUDA stream, so that at any time one is in its tensor-bound encode phase and the other in its HBM-bound decode
phase.  `--sm-cap N` caps the persistent row-form GEMM's grid (md_debug_gemm_sm_cap) so its CTAs do not hold every
SM for the length of a kernel.  Prints images/s for the serial baseline and for each cap.

    python tools/overlap_probe.py [--batches 6] [--caps 0,128,112,96,80]
"""
import argparse
import os
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, synth  # noqa: E402
from moondream_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--batches", type=int, default=6)
ap.add_argument("--caps", default="0,128,112,96,80")
ap.add_argument("--new-tokens", type=int, default=64)
args = ap.parse_args()

B = args.batch
cfg = C.preset("moondream-2b")
sd = synth.synthetic_state_dict(cfg, 0)
images = [synth.synthetic_image(i, 378, 378) for i in range(B)]
prompts = [synth.synthetic_prompt(i, 32, cfg.text.vocab_size) for i in range(B)]
engines = [Engine(cfg, sd, max_batch=B) for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
lib = engines[0].lib


def one_batch(eng):
    pre = eng.encode_images(images)
    eng.generate(pre, prompts, args.new_tokens, consume=True, stop_on_eos=False, to_host=False)


# warm-up (graph capture happens on each engine's own stream)
for eng, st in zip(engines, streams):
    with torch.cuda.stream(st):
        one_batch(eng)
        one_batch(eng)
torch.cuda.synchronize()


def serial(n):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        one_batch(engines[0])
    e.record()
    torch.cuda.synchronize()
    return n * B / (s.elapsed_time(e) / 1e3)


def overlapped(n):
    """Each engine is driven by its own host thread on its own stream; the second starts half a period late."""
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    start.record()
    gate = threading.Event()

    def worker(k):
        with torch.cuda.stream(streams[k]):
            streams[k].wait_event(start)
            if k == 1:
                gate.wait()
            for i in range(n // 2):
                if k == 0 and i == 0:
                    pre = engines[0].encode_images(images)           # engine 1 starts once this encode is queued
                    gate.set()
                    engines[0].generate(pre, prompts, args.new_tokens, consume=True, stop_on_eos=False, to_host=False)
                else:
                    one_batch(engines[k])
            ends[k].record()

    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    ms = max(start.elapsed_time(e) for e in ends)
    return (n // 2) * 2 * B / (ms / 1e3)


print(f"serial, one engine: {serial(args.batches):.1f} images/s", flush=True)
for cap in [int(c) for c in args.caps.split(",")]:
    lib.md_debug_gemm_sm_cap(cap)
    print(f"two engines / two streams, GEMM cap {cap or 'none'}: {overlapped(args.batches):.1f} images/s", flush=True)
lib.md_debug_gemm_sm_cap(0)

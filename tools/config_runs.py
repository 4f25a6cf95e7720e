"""Runs the other BASELINE.json configs at full size (sanity + CUDA-event timing) -> gpurun_out/config_runs.json.
  C3  Moondream-2B, 32 images 756x756 (tiling 3x3 = 10 crops each, 320 crops), caption 64 tokens
  C4' Moondream-2B, 32 images/GPU detect() with the region head, max_objects 8 (the per-GPU share of b256 over 8 GPUs)
  C5' Moondream-0.5B, batch 128, 378x378, 256-token decode: bf16, and the decoder blocks streamed as int8 / int4 group 128
      (the reference's only quantisation is the int4 QuantizedLinear, layers.py:38-110; parity = bf16 on the dequantised weights)
  C2  Moondream-2B, batch 1: encode_image + query (32-token prompt, 64 tokens) latency
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, quant, synth  # noqa: E402
from moondream_b200.engine import Engine  # noqa: E402

out = {}


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        r = fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps, r


cfg = C.preset("moondream-2b")
sd = synth.synthetic_state_dict(cfg, 0)
eng = Engine(cfg, sd, max_batch=32)
B = 32
prompts = [synth.synthetic_prompt(i, 32, cfg.text.vocab_size) for i in range(B)]
imgs = [synth.synthetic_image(i, 756, 756) for i in range(B)]
dev, offs, til = eng.stage_images(imgs)
ms, res = timed(lambda: eng.caption_from_crops(dev, offs, til, prompts, 64, to_host=False, stop_on_eos=False))
out["C3_2B_b32_756x756_10crops_caption64"] = {"ms_per_batch": ms, "images_per_s": B / ms * 1e3, "crops": int(dev.shape[0]),
                                             "tiling": list(til[0])}
print(out, flush=True)

tk = cfg.tokenizer
imgs = [synth.synthetic_image(100 + i, 378, 378) for i in range(B)]
dprompts = [tk.templates["detect"]["prefix"] + synth.synthetic_prompt(i, 27, cfg.text.vocab_size) + tk.templates["detect"]["suffix"]
            for i in range(B)]


def detect():
    pre = eng.encode_images(imgs)
    return eng.generate_points(pre, dprompts, include_size=True, max_objects=8)


ms, objs = timed(detect)
out["C4_2B_b32_detect_max8"] = {"ms_per_batch": ms, "images_per_s": B / ms * 1e3,
                                "objects_per_image_mean": sum(len(o) for o in objs) / B}
print(out["C4_2B_b32_detect_max8"], flush=True)
# C2: one image, one query
img1 = [synth.synthetic_image(7, 378, 378)]
p1 = [synth.synthetic_prompt(7, 32, cfg.text.vocab_size)]


def single():
    d1, o1, t1 = eng.stage_images(img1)
    return eng.caption_from_crops(d1, o1, t1, p1, 64, to_host=True, stop_on_eos=False)


ms, _ = timed(single, reps=5)
out["C2_2B_b1_encode_plus_query64"] = {"ms": ms, "images_per_s": 1e3 / ms}
print(out["C2_2B_b1_encode_plus_query64"], flush=True)
del eng
torch.cuda.empty_cache()

cfg = C.preset("moondream-0.5b")
sd = synth.synthetic_state_dict(cfg, 0)
B = 128
eng = Engine(cfg, sd, max_batch=B)
imgs = [synth.synthetic_image(i, 378, 378) for i in range(B)]
prompts = [synth.synthetic_prompt(i, 32, cfg.text.vocab_size) for i in range(B)]
dev, offs, til = eng.stage_images(imgs)
ms, res = timed(lambda: eng.caption_from_crops(dev, offs, til, prompts, 256, to_host=False, stop_on_eos=False), reps=1)
ms8, _ = timed(lambda: eng.caption_from_crops(dev, offs, til, prompts, 8, to_host=False, stop_on_eos=False), reps=1)
step_ms = (ms - ms8) / 248.0
kv_bytes = sum(2 * cfg.text.n_layers * cfg.text.n_heads * 64 * 2 * (762 + s) for s in range(8, 256)) / 248.0 * B
w_bytes = 2 * (cfg.text.n_layers * (4 * cfg.text.dim ** 2 + 2 * cfg.text.dim * cfg.text.ff_dim) + cfg.text.dim * cfg.text.vocab_size)
out["C5_0.5B_b128_decode256_bf16"] = {"ms_per_batch": ms, "images_per_s": B / ms * 1e3, "decode_ms_per_step": step_ms,
                                       "tokens_per_s": B / step_ms * 1e3,
                                       "algorithmic_GB_per_step": (kv_bytes + w_bytes) / 1e9,
                                       "achieved_GBps": (kv_bytes + w_bytes) / 1e6 / step_ms}
print(out["C5_0.5B_b128_decode256_bf16"], flush=True)
tok_bf16 = res.tokens.clone() if hasattr(res, "tokens") else None
for mode, bits in (("int8", 8), ("int4", 4)):
    del eng
    torch.cuda.empty_cache()
    eng = Engine(cfg, sd, max_batch=B, quantize=mode)
    ms, res = timed(lambda: eng.caption_from_crops(dev, offs, til, prompts, 256, to_host=False, stop_on_eos=False), reps=1)
    ms8, _ = timed(lambda: eng.caption_from_crops(dev, offs, til, prompts, 8, to_host=False, stop_on_eos=False), reps=1)
    step_ms = (ms - ms8) / 248.0
    wq_bytes = quant.stream_bytes(cfg, bits=bits)
    out[f"C5_0.5B_b128_decode256_{mode}"] = {
        "ms_per_batch": ms, "images_per_s": B / ms * 1e3, "decode_ms_per_step": step_ms, "tokens_per_s": B / step_ms * 1e3,
        "weight_GB_per_step": wq_bytes / 1e9, "algorithmic_GB_per_step": (kv_bytes + wq_bytes) / 1e9,
        "achieved_GBps": (kv_bytes + wq_bytes) / 1e6 / step_ms,
        "decoder_block_bytes_resident": eng.quantized.nbytes()}
    print(out[f"C5_0.5B_b128_decode256_{mode}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/config_runs.json", "w"), indent=1)

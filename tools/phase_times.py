"""CUDA-event timing of each phase of one bench step (2B, b32) -> gpurun_out/phase_times.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, synth  # noqa: E402
from moondream_b200.engine import Engine  # noqa: E402
from moondream_b200.image_crops import overlap_crop_image  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "moondream-2b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = C.preset(model)
sd = synth.synthetic_state_dict(cfg, 0)
eng = Engine(cfg, sd, max_batch=B)
eng.lib.md_debug_gemm(int(os.environ.get("MD_DEBUG_GEMM", "0")))   # timing experiments (A/B of plans)
eng.lib.md_debug_attention_impl(int(os.environ.get("MD_ATTENTION_IMPL", "0")))
images = [synth.synthetic_image(i, 378, 378) for i in range(B)]
prompts = [synth.synthetic_prompt(i, 32, cfg.text.vocab_size) for i in range(B)]
crops, offsets, tilings = [], [0], []
for im in images:
    oc = overlap_crop_image(im, overlap_margin=4, max_crops=12)
    crops.append(oc["crops"]); tilings.append(oc["tiling"]); offsets.append(offsets[-1] + oc["crops"].shape[0])
crops_dev = torch.from_numpy(np.concatenate(crops, 0)).cuda()


class T:
    def __init__(self):
        self.ev = []

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append((name, e))

    def report(self):
        torch.cuda.synchronize()
        return {self.ev[i + 1][0]: self.ev[i][1].elapsed_time(self.ev[i + 1][1]) for i in range(len(self.ev) - 1)}


t = cfg.text
res = {}
for it in range(3):
    tm = T()
    tm.mark("start")
    feats = eng.vision_encode(crops_dev)
    tm.mark("vit_encode")
    embeds = torch.empty((B * t.prefix_attn, t.dim), dtype=torch.bfloat16, device="cuda")
    eng.vision_project(feats, offsets, tilings, embeds)
    tm.mark("stitch_pool_project")
    res = tm.report()
    tm2 = T()
    tm2.mark("start")
    prefixes = eng.encode_crops(crops_dev, offsets, tilings)
    tm2.mark("encode_total(vit+proj+image_prefill)")
    out = eng.generate(prefixes, prompts, 64, consume=True, stop_on_eos=False, to_host=False)
    tm2.mark("generate_total(prompt_prefill+65 lm_head+64 decode)")
    out1 = eng.generate(eng.encode_crops(crops_dev, offsets, tilings), prompts, 1, consume=True, stop_on_eos=False, to_host=False)
    tm2.mark("encode+generate(1 token)")
    res.update(tm2.report())
res["decode_ms_per_step_est"] = (res["generate_total(prompt_prefill+65 lm_head+64 decode)"] -
                                 (res["encode+generate(1 token)"] - res["encode_total(vit+proj+image_prefill)"])) / 63.0
res["image_prefill_est"] = res["encode_total(vit+proj+image_prefill)"] - res["vit_encode"] - res["stitch_pool_project"]
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/phase_times.json", "w"), indent=1)

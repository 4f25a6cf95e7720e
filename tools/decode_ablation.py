"""Time the CUDA-graph decode step with individual kernels skipped (md_debug_skip_decode_kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, synth  # noqa: E402
from moondream_b200.engine import Engine  # noqa: E402

B = 32
cfg = C.preset("moondream-2b")
sd = synth.synthetic_state_dict(cfg, 0)
images = [synth.synthetic_image(i, 378, 378) for i in range(B)]
prompts = [synth.synthetic_prompt(i, 32, cfg.text.vocab_size) for i in range(B)]
names = {0: "full", 1: "-gemm1", 2: "-epi1", 4: "-attn", 8: "-gemm2", 16: "-epi2", 31: "-all five (LN/embed/lm_head/argmax/advance only)",
         27: "attention only + rest", 4 | 2 | 16: "gemms only + rest"}
base = None
for mask, name in names.items():
    eng = Engine(cfg, sd, max_batch=B)          # fresh engine: the graph is captured with this mask
    eng.lib.md_debug_skip_decode_kernels(mask)
    pre = eng.encode_images(images)
    eng.generate(pre, prompts, 8, stop_on_eos=False, to_host=False)      # capture + warm
    times = []
    for n_tok in (8, 72):
        pre = eng.encode_images(images)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.generate(pre, prompts, n_tok, stop_on_eos=False, to_host=False)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    per_step = (times[1] - times[0]) / 64.0
    base = per_step if base is None else base
    print(f"{name:55s} {per_step*1000:8.1f} us/step   delta {1000*(base-per_step):7.1f} us", flush=True)
    eng.lib.md_debug_skip_decode_kernels(0)
    del eng
    torch.cuda.empty_cache()

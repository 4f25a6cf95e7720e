"""Short decode-only run for profilers: 2B, b32, encode once, then N eager (non-graph) decode steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, synth  # noqa: E402
from moondream_b200.engine import Engine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = C.preset("moondream-2b")
eng = Engine(cfg, synth.synthetic_state_dict(cfg, 0), max_batch=B)
images = [synth.synthetic_image(i, 378, 378) for i in range(B)]
prompts = [synth.synthetic_prompt(i, 32, cfg.text.vocab_size) for i in range(B)]
res = eng.generate(eng.encode_images(images), prompts, steps, use_graph=False, stop_on_eos=False)
torch.cuda.synchronize()
print("tokens", res.tokens[0, : steps + 1].tolist())

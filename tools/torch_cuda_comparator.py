"""The reference's torch path on the SAME B200 (eager torch-CUDA ops, batch-1, sequential): the comparator of
BASELINE.json's ">= 8x the reference torch-CUDA path" target.  Uses the oracle port of the
reference (bit-identical arithmetic to /root/reference, which cannot travel to the GPU box) on device="cuda".
Writes gpurun_out/torch_cuda_comparator.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, synth  # noqa: E402
from oracle.moondream_oracle import OracleModel  # noqa: E402

cfg = C.preset("moondream-2b")
sd = synth.synthetic_state_dict(cfg, 0)
orc = OracleModel(cfg, sd, device="cuda")
n_img, n_tok = 4, 64


def one(i):
    img = synth.synthetic_image(i, 378, 378)
    prompt = synth.synthetic_prompt(i, 32, cfg.text.vocab_size)
    enc = orc.encode_image(img)
    return orc.generate(enc, prompt, n_tok).tokens


one(0)
torch.cuda.synchronize()
t0 = time.perf_counter()
toks = [one(i) for i in range(n_img)]
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res = {"impl": "reference arithmetic (oracle port), eager torch-CUDA ops, batch-1 sequential, per-token .item() sync like "
               "moondream.py:482", "images": n_img, "tokens_per_image": n_tok, "seconds": dt, "images_per_s": n_img / dt,
       "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/torch_cuda_comparator.json", "w"), indent=1)

"""A few launches of the patch-embedding kernel and of the quantised weight stream (2B shapes) for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, ops, quant, synth  # noqa: E402
from moondream_b200.engine import Engine  # noqa: E402

cfg = C.tiny()
eng = Engine(cfg, synth.synthetic_state_dict(cfg, 0), max_batch=2, kv_pages=64)
crops = torch.randint(0, 256, (64, 378, 378, 3), dtype=torch.uint8).cuda()
for _ in range(2):
    eng.vision_encode(crops)                     # patch_embed_kernel over 64 crops (N = 144 here; the gather is the same)
torch.cuda.synchronize()
for bits in (4, 8):
    g = torch.Generator().manual_seed(bits)
    w = (torch.randn(14336, 2048, generator=g) / 45).to(torch.bfloat16)
    if bits == 4:
        nib, s, z = quant.quantize_weight_int4(w)
        ql = quant.QuantLinear(4, nib, s, z)
    else:
        q8, s8 = quant.quantize_weight_int8(w)
        ql = quant.QuantLinear(8, q8.view(torch.uint8), s8.float().unsqueeze(1).repeat(1, 16).contiguous(), torch.zeros(14336, 16))
    dev = [t.cuda() for t in (ql.stream_bytes(), ql.scale, ql.zero)]
    x = torch.randn(32, 2048, device="cuda").bfloat16()
    for _ in range(2):
        ops.linear_small_batch_quant(bits, x, *dev, 14336)
    ops.dequantize_weights(bits, *dev, 14336, 2048)
torch.cuda.synchronize()
print("done")

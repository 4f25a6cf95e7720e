"""TEST INFRASTRUCTURE (build container only): golden vectors for the int4 group-128 dequantisation from the UNMODIFIED
reference function `moondream.torch.layers.dequantize_tensor` (layers.py:38-44).

    python -m oracle.make_golden_quant      ->  tests/golden/int4_dequant.json

Each case: a seeded weight matrix quantised by moondream_b200.quant.quantize_weight_int4 (integer zero points, bf16-valued
scales) or with deliberately awkward parameters (fractional zero points, full-precision fp32 scales), packed in the
reference's checkpoint layout, dequantised by the reference.  Stored: the seed / recipe (inputs are regenerated), the
sha256 of the reference's bf16 bytes and the first 16 values.  The oracle restatement is asserted equal on the way.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import torch

from moondream_b200 import quant
from oracle import reference_shim as R
from oracle.moondream_oracle import dequantize_tensor as oracle_dequantize

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "int4_dequant.json")
CASES = [  # (seed, out, in, awkward)
    (0, 16, 256, False), (1, 64, 128, False), (2, 32, 512, True), (3, 8, 1024, True), (4, 256, 256, False)]


def make_case(seed: int, out_f: int, in_f: int, awkward: bool):
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(out_f, in_f, generator=g) * 0.04).to(torch.bfloat16)
    nib, scale, zero = quant.quantize_weight_int4(w)
    if awkward:       # parameters a bf16 / integer recipe never produces: both roundings of the formula become visible
        scale = scale * (1 + torch.rand(scale.shape, generator=g) * 1e-3)
        zero = zero + torch.rand(zero.shape, generator=g) - 0.5
    return nib, scale.float().contiguous(), zero.float().contiguous()


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().view(torch.int16).numpy().tobytes()).hexdigest()


def main():
    assert R.reference_available(), "run in the build container (/root/reference)"
    sys.path.insert(0, R.REFERENCE_ROOT)
    from moondream.torch.layers import dequantize_tensor as ref_dequantize

    cases = []
    for seed, out_f, in_f, awkward in CASES:
        nib, scale, zero = make_case(seed, out_f, in_f, awkward)
        packed = quant.pack_reference_int4(nib)
        ref = ref_dequantize(packed.clone(), scale.reshape(-1, 1), zero.reshape(-1, 1), (out_f, in_f), torch.bfloat16)
        orc = oracle_dequantize(packed, scale.reshape(-1, 1), zero.reshape(-1, 1), (out_f, in_f))
        assert torch.equal(ref, orc), "oracle restatement differs from the reference"
        assert torch.equal(ref, quant.dequantize(nib, scale, zero)), "product-side dequantize differs from the reference"
        cases.append({"seed": seed, "out": out_f, "in": in_f, "awkward": awkward, "sha256": sha(ref),
                      "first16": ref.flatten()[:16].view(torch.int16).tolist()})
    json.dump({"source": "moondream.torch.layers.dequantize_tensor (layers.py:38-44), unmodified", "cases": cases},
              open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

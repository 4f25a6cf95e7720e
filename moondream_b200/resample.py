"""Lanczos resize with Pillow's exact 8-bit arithmetic, for the on-device image preprocessing.

The reference resizes on the host with ``PIL.Image.resize(..., LANCZOS)`` (moondream/torch/image_crops.py:138-150).
Pillow's 8-bit resampler (src/libImaging/Resample.c) is integer arithmetic: per output coordinate a window of
double-precision Lanczos-3 weights (support 3 x max(scale, 1), antialiased) is normalised and converted to 22-bit fixed
point; a horizontal pass and then a vertical pass accumulate uint8 x int32 products from 2^21 and shift right by 22 with
a clamp to 0..255, with a uint8 intermediate image between the passes.  This module computes the SAME coefficient
tables on the host (same expressions in the same order, in double), so the CUDA kernels that apply them
(csrc/preprocess.cu) reproduce Pillow bit for bit — tests/test_resample.py pins that against PIL itself.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients in 22-bit fixed point
LANCZOS_SUPPORT = 3.0


def _lanczos(x: float) -> float:
    """lanczos_filter / sinc_filter of Resample.c (truncated sinc, a = 3)."""
    if not (-3.0 <= x < 3.0):
        return 0.0

    def sinc(v: float) -> float:
        if v == 0.0:
            return 1.0
        v = v * math.pi
        return math.sin(v) / v

    return sinc(x) * sinc(x / 3)


@lru_cache(maxsize=256)
def lanczos_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """precompute_coeffs + normalize_coeffs_8bpc for the full-image box (in0 = 0, in1 = in_size):
    bounds int32 [out_size, 2] = (first source index, tap count), coeffs int32 [out_size, ksize]."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            # normalize_coeffs_8bpc: round half away from zero through a C cast (truncation)
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """one resampling pass along `axis` (0 vertical, 1 horizontal) in Pillow's fixed point (numpy restatement)"""
    src = np.moveaxis(img, axis, 0).astype(np.int64)          # [in, other, C]
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for i in range(bounds.shape[0]):
        lo, n = int(bounds[i, 0]), int(bounds[i, 1])
        acc = np.tensordot(kk[i, :n].astype(np.int64), src[lo: lo + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_lanczos_numpy(image: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """ImagingResample for uint8 HWC images: horizontal pass (if the width changes), then vertical pass."""
    h, w = image.shape[:2]
    if (h, w) == (out_h, out_w):
        return image.copy()
    cur = image
    if w != out_w:
        # Pillow resamples only the source rows the vertical pass will read; the rows it skips are never used
        cur = _pass(cur, *lanczos_coeffs(w, out_w), axis=1)
    if h != out_h:
        cur = _pass(cur, *lanczos_coeffs(h, out_h), axis=0)
    return cur

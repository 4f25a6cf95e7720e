"""Host-side overlap-and-resize cropping — API-compatible with the reference's
``moondream/torch/image_crops.py`` (``select_tiling`` :17-50, ``overlap_crop_image`` :58-167,
``reconstruct_from_crops`` :170-231) so ``tests/test_image_crops.py`` of the reference passes
unchanged against this module.

The device path never calls ``reconstruct_from_crops``: stitching + pooling of crop *features*
runs in one CUDA kernel (csrc/elementwise.cu: stitch_pool_concat).  The function is kept for
callers that stitch pixel crops on the host.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple, TypedDict

import numpy as np
import torch

try:  # the reference prefers libvips when present (image_crops.py:8-14); this image has PIL only
    import pyvips  # type: ignore

    HAS_VIPS = True
except Exception:  # pragma: no cover - depends on the box
    from PIL import Image

    HAS_VIPS = False


class OverlapCropOutput(TypedDict):
    crops: np.ndarray
    tiling: Tuple[int, int]


def select_tiling(height: int, width: int, crop_size: int, max_crops: int) -> Tuple[int, int]:
    """Tile grid (rows, cols) covering a height x width image with at most `max_crops` windows."""
    if min(height, width) <= crop_size:
        return (1, 1)
    rows_needed = math.ceil(height / crop_size)
    cols_needed = math.ceil(width / crop_size)
    if rows_needed * cols_needed > max_crops:
        scale = math.sqrt(max_crops / (rows_needed * cols_needed))
        return (max(1, math.floor(rows_needed * scale)), max(1, math.floor(cols_needed * scale)))
    rows = max(math.floor(math.sqrt(max_crops * height / width)), rows_needed)
    cols = max(math.floor(math.sqrt(max_crops * width / height)), cols_needed)
    if rows * cols > max_crops:
        if cols > rows:
            cols = math.floor(max_crops / rows)
        else:
            rows = math.floor(max_crops / cols)
    return (max(1, rows), max(1, cols))


def _resize(image: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    if image.shape[0] == out_h and image.shape[1] == out_w:
        return image        # PIL (Image.resize) and vips return an unchanged copy for a same-size resize
    if HAS_VIPS:  # pragma: no cover
        vimg = pyvips.Image.new_from_array(image)
        return vimg.resize(out_w / image.shape[1], vscale=out_h / image.shape[0]).numpy()
    return np.asarray(Image.fromarray(image).resize((int(out_w), int(out_h)),
                                                     resample=Image.Resampling.LANCZOS))


def crop_tiling(image_shape, overlap_margin: int, max_crops: int, base_size: Tuple[int, int] = (378, 378),
                patch_size: int = 14) -> Tuple[int, int]:
    """The (rows, cols) grid `overlap_crop_image` will use for an image of this shape."""
    margin_px = patch_size * overlap_margin
    window = (base_size[0] // patch_size - 2 * overlap_margin) * patch_size
    return select_tiling(image_shape[0] - 2 * margin_px, image_shape[1] - 2 * margin_px, window, max_crops)


def overlap_crop_image(image: np.ndarray, overlap_margin: int, max_crops: int,
                       base_size: Tuple[int, int] = (378, 378), patch_size: int = 14,
                       out: np.ndarray = None) -> OverlapCropOutput:
    """Global crop (index 0) + rows*cols overlapping local crops, all `base_size`, uint8 HWC.
    `out` (optional, beyond the reference's signature): a preallocated uint8
    [rows*cols + 1, H, W, C] array (e.g. a slice of pinned staging memory) to write into."""
    margin_px = patch_size * overlap_margin
    window = (base_size[0] // patch_size - 2 * overlap_margin) * patch_size
    rows, cols = crop_tiling(image.shape, overlap_margin, max_crops, base_size, patch_size)
    shape = (rows * cols + 1, base_size[0], base_size[1], image.shape[2])
    if out is None:
        crops = np.zeros(shape, dtype=np.uint8)
    else:
        if tuple(out.shape) != shape or out.dtype != np.uint8:
            raise ValueError(f"out must be uint8 {shape}")
        crops = out
    canvas = _resize(image, rows * window + 2 * margin_px, cols * window + 2 * margin_px)
    crops[0] = _resize(image, base_size[0], base_size[1])
    for r in range(rows):
        for c in range(cols):
            tile = canvas[r * window: r * window + base_size[0], c * window: c * window + base_size[1]]
            dst = crops[1 + r * cols + c]
            if out is not None and (tile.shape[0] < base_size[0] or tile.shape[1] < base_size[1]):
                dst[...] = 0                      # partially covered tile: clear the staging slot first
            dst[: tile.shape[0], : tile.shape[1]] = tile
    return {"crops": crops, "tiling": (rows, cols)}


def reconstruct_from_crops(crops: Sequence[torch.Tensor], tiling: Tuple[int, int],
                           overlap_margin: int, patch_size: int = 14) -> torch.Tensor:
    """Stitch local crops (H, W, C each) back into one image, keeping outer margins only."""
    rows, cols = tiling
    ch, cw = crops[0].shape[:2]
    m = overlap_margin * patch_size
    out = torch.zeros(((ch - 2 * m) * rows + 2 * m, (cw - 2 * m) * cols + 2 * m, crops[0].shape[2]),
                      device=crops[0].device, dtype=crops[0].dtype)
    for idx, crop in enumerate(crops):
        r, c = divmod(idx, cols)
        y0, y1 = (0 if r == 0 else m), (ch if r == rows - 1 else ch - m)
        x0, x1 = (0 if c == 0 else m), (cw if c == cols - 1 else cw - m)
        oy, ox = r * (ch - 2 * m), c * (cw - 2 * m)
        out[oy + y0: oy + y1, ox + x0: ox + x1] = crop[y0:y1, x0:x1]
    return out

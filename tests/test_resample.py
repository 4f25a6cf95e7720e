"""CPU: moondream_b200/resample.py restates Pillow's 8-bit Lanczos resize (libImaging/Resample.c) — coefficient tables
and the two fixed-point passes — and must equal PIL.Image.resize(..., LANCZOS) bit for bit; the CUDA kernels apply the
same tables (tests/test_features_gpu.py::test_device_preprocessing_is_bit_exact_with_pil)."""
import numpy as np
import pytest
from PIL import Image

from moondream_b200.resample import PRECISION_BITS, lanczos_coeffs, resize_lanczos_numpy


@pytest.mark.parametrize("src,dst", [((500, 700), (644, 910)), ((800, 600), (1176, 644)), ((300, 200), (378, 378)),
                                     ((1080, 1920), (910, 1176)), ((50, 1000), (378, 1442)), ((378, 500), (378, 378)),
                                     ((379, 378), (378, 378)), ((17, 19), (378, 378)), ((378, 378), (378, 378))])
def test_numpy_restatement_equals_pil(src, dst):
    img = np.random.default_rng(src[0] * 7 + src[1]).integers(0, 256, (*src, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((dst[1], dst[0]), resample=Image.Resampling.LANCZOS))
    assert np.array_equal(resize_lanczos_numpy(img, *dst), want)


def test_smooth_image_and_saturation():
    yy, xx = np.mgrid[0:300, 0:400]
    img = np.stack([yy * 255 // 299, xx * 255 // 399, (yy // 50 + xx // 50) % 2 * 255], -1).astype(np.uint8)   # ramps + checkerboard
    for dst in ((644, 910), (150, 123), (378, 378)):
        want = np.asarray(Image.fromarray(img).resize((dst[1], dst[0]), resample=Image.Resampling.LANCZOS))
        assert np.array_equal(resize_lanczos_numpy(img, *dst), want)


def test_coefficient_tables():
    for in_size, out_size in ((700, 910), (1920, 1176), (200, 378), (3000, 378)):
        bounds, kk = lanczos_coeffs(in_size, out_size)
        scale = max(in_size / out_size, 1.0)
        assert kk.shape == (out_size, int(np.ceil(3.0 * scale)) * 2 + 1) and bounds.shape == (out_size, 2)
        assert bounds[:, 0].min() >= 0 and (bounds[:, 0] + bounds[:, 1]).max() <= in_size
        sums = np.array([kk[i, : bounds[i, 1]].sum() for i in range(out_size)])
        assert np.abs(sums - (1 << PRECISION_BITS)).max() <= kk.shape[1]        # weights sum to 1 up to rounding
        assert (kk[np.arange(kk.shape[1])[None, :] >= bounds[:, 1:2]] == 0).all()

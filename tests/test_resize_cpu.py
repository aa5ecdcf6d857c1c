"""The resampling oracle (oracle/resize.py) against PIL itself -- the implementation the reference's drivers call
(infer/infer_omgsr_s.py:76-85) -- and the library's host-side table builder against the oracle. Bit-exact."""
import numpy as np
import pytest
from PIL import Image

from oracle import resize as oresize

CASES = [
    # H, W -> h, w, filter           (the two steps of the reference's pre-processing chain, plus odd sizes)
    (32, 48, 128, 192, oresize.BICUBIC),     # x4 upscale
    (67, 45, 268, 180, oresize.BICUBIC),
    (133, 90, 128, 88, oresize.LANCZOS),     # crop-to-multiple-of-8 step
    (64, 64, 61, 37, oresize.LANCZOS),
    (50, 70, 23, 31, oresize.BICUBIC),       # downscale (process_size smaller than the image)
    (40, 40, 40, 64, oresize.BICUBIC),       # one axis unchanged
]
PIL_FILTER = {oresize.BICUBIC: Image.BICUBIC, oresize.LANCZOS: Image.LANCZOS}


@pytest.mark.parametrize("H,W,h,w,flt", CASES)
def test_oracle_resize_matches_pil(H, W, h, w, flt):
    a = np.random.default_rng(H * 1000 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a, "RGB").resize((w, h), PIL_FILTER[flt]))
    got = oresize.resize_u8(a, (h, w), flt)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_default_filter_is_bicubic():
    """the reference calls img.resize(size) with no filter for the upscale step: Pillow's default must be BICUBIC"""
    a = np.random.default_rng(7).integers(0, 256, (16, 16, 3), dtype=np.uint8)
    im = Image.fromarray(a, "RGB")
    assert np.array_equal(np.asarray(im.resize((64, 64))), np.asarray(im.resize((64, 64), Image.BICUBIC)))


@pytest.mark.parametrize("n_in,n_out,flt", [(48, 192, 0), (90, 88, 1), (70, 31, 0), (512, 2048, 0), (1023, 1016, 1)])
def test_library_tables_match_oracle(n_in, n_out, flt):
    """omgsr_resample_coeffs is host code: no GPU needed"""
    from omgsr_b200 import capi
    b, k = capi.resample_coeffs(n_in, n_out, flt)
    ob, ok = oresize.coeffs(n_in, n_out, flt)
    assert np.array_equal(b, ob) and np.array_equal(k, ok)

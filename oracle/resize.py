"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the image resampling the reference's drivers
do through PIL (`Image.resize`, infer/infer_omgsr_s.py:76-85: default filter BICUBIC for the x`upscale` step, LANCZOS
for the crop-to-multiple-of-8 step).

The arithmetic lives in a third-party dependency, Pillow (requirements.txt pins no version; this image has 12.2.0):
src/libImaging/Resample.c -- `precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`,
`ImagingResampleVertical_8bpc`, filters `bicubic_filter` (a = -0.5, support 2) and `lanczos_filter` (support 3).
Restated here in numpy (double-precision tables, 22-bit fixed-point taps, int32 accumulation, clip8).
Pinned: tests/test_resize_cpu.py compares it bit for bit with PIL itself on seeded images.
"""
import math

import numpy as np

BICUBIC, LANCZOS = 0, 1
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def coeffs(in_size, out_size, flt):
    """-> (bounds int32 [out, 2] = (first tap, count), taps int32 [out, ksize]) as Pillow builds them."""
    f, fsupport = (_lanczos, 3.0) if flt == LANCZOS else (_bicubic, 2.0)
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [f((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        bounds[xx] = (xmin, xmax)
        for x, v in enumerate(w):
            s = v * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + s) if v < 0 else int(0.5 + s)
    return bounds, kk


def _pass(img, out_size, axis, flt):
    b, k = coeffs(img.shape[axis], out_size, flt)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for o in range(out_size):
        first, n = int(b[o, 0]), int(b[o, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(k[o, :n].astype(np.int64), src[first:first + n], axes=([0], [0]))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def resize_u8(img, size_hw, flt=BICUBIC):
    """img: uint8 [H, W, 3] -> uint8 [h, w, 3]; horizontal pass first, then vertical, as ImagingResampleInner."""
    h, w = size_hw
    out = img
    if w != img.shape[1]:
        out = _pass(out, w, 1, flt)
    if h != img.shape[0]:
        out = _pass(out, h, 0, flt)
    return out.copy() if out is img else out

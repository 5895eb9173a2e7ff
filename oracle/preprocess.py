"""CPU oracle for the RGB image processor.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates what the reference's SigLipImageProcessor.preprocess does to one frame
(/root/reference/llava/model/multimodal_encoder/siglip_encoder.py:34-66: convert_to_rgb, PIL BICUBIC
resize to 384x384 (squash), rescale 1/255, normalize mean = std = 0.5, channels first), with the resize
written out as Pillow's ImagingResample algorithm for 8-bit images: separable convolution, horizontal pass
then vertical pass, support scaled by the down-sampling factor, coefficients normalised in double and
converted to 22-bit fixed point, each pass rounded and clipped to uint8.  Pillow is a third-party
dependency of the reference (not vendored); PINNED by comparing against Pillow itself in
tests/test_preprocess_cpu.py.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc.  Returns (xmin [out], count [out], kk [out, ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx], cnt[xx] = lo, n
    return xmin, cnt, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """one resampling pass along `axis` of a uint8 [H,W,C] image"""
    in_size = img.shape[axis]
    xmin, cnt, kk = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)       # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt[xx]):
            acc += src[xmin[xx] + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int = 384, out_w: int = 384) -> np.ndarray:
    """uint8 [H,W,3] -> uint8 [out_h,out_w,3], Pillow Image.resize(..., BICUBIC) semantics."""
    tmp = _pass(np.ascontiguousarray(img), out_w, 1) if img.shape[1] != out_w else img   # horizontal first
    return _pass(tmp, out_h, 0) if img.shape[0] != out_h else tmp


def preprocess(img: np.ndarray, size: int = 384) -> np.ndarray:
    """uint8 [H,W,3] -> float32 [3,size,size]: resize, * 1/255, (x - 0.5) / 0.5 (all float32)."""
    r = resize_bicubic_u8(img, size, size)
    a = r.astype(np.float32) * np.float32(1.0 / 255.0)
    a = (a - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(a.transpose(2, 0, 1))

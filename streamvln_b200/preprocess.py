"""GPU image processor: uint8 frame -> [3,384,384] float32, bit-exact with the reference's
SigLipImageProcessor (llava/model/multimodal_encoder/siglip_encoder.py:34-66; PIL BICUBIC squash, 1/255,
mean = std = 0.5) through svln_preprocess_rgb.  The tap tables follow Pillow's precompute_coeffs /
normalize_coeffs_8bpc (libImaging/Resample.c; Pillow is a dependency of the reference, not vendored)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr

PRECISION_BITS = 32 - 8 - 2


def _cubic(x: float) -> float:
    a = -0.5
    x = -x if x < 0.0 else x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_taps(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    scale = in_size / out_size
    fscale = scale if scale > 1.0 else 1.0
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    lo_a, n_a, kk = np.zeros(out_size, np.int32), np.zeros(out_size, np.int32), np.zeros((out_size, ksize), np.int32)
    inv = 1.0 / fscale
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        w = [_cubic((x + lo - center + 0.5) * inv) for x in range(hi - lo)]
        tot = 0.0
        for v in w:
            tot += v
        for x, v in enumerate(w):
            v = v / tot if tot != 0.0 else v
            kk[o, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        lo_a[o], n_a[o] = lo, hi - lo
    return lo_a, n_a, kk


class GpuImageProcessor:
    def __init__(self, device, size: int = 384):
        self.lib = _lib.load()
        self.device, self.size = torch.device(device), size
        self._taps: Dict[int, tuple] = {}
        self._ws = None

    def _tables(self, in_size: int):
        t = self._taps.get(in_size)
        if t is None:
            lo, n, kk = bicubic_taps(in_size, self.size)
            t = tuple(torch.from_numpy(x).to(self.device) for x in (lo, n, kk)) + (kk.shape[1],)
            self._taps[in_size] = t
        return t

    def __call__(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """frames_u8 [n,H,W,3] uint8 on the device -> [n,3,size,size] float32."""
        n, H, W, _ = frames_u8.shape
        xl, xn, xk, xks = self._tables(W)
        yl, yn, yk, yks = self._tables(H)
        need = self.lib.svln_preprocess_workspace_bytes(n, H, self.size)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(n, 3, self.size, self.size, dtype=torch.float32, device=self.device)
        check(self.lib.svln_preprocess_rgb(ptr(frames_u8), n, H, W, self.size, self.size, ptr(xl), ptr(xn), ptr(xk), xks,
                                           ptr(yl), ptr(yn), ptr(yk), yks, ptr(self._ws), self._ws.numel(), ptr(out),
                                           cur_stream()))
        return out


class GpuDepthProcessor:
    """Habitat [0,1] depth [n,H,W] (device float32) -> metres on the 384x384 grid through svln_preprocess_depth, bit-exact
    with the reference's host path (streamvln_eval.py:126-135,254-256,276; host restatement: synth.preprocess_depth)."""

    def __init__(self, device, size: int = 384, min_depth: float = 0.0, max_depth: float = 10.0):
        self.lib = _lib.load()
        self.device, self.size, self.min_depth, self.max_depth = torch.device(device), size, min_depth, max_depth

    def __call__(self, depth01: torch.Tensor) -> torch.Tensor:
        d = depth01.to(self.device, torch.float32).contiguous()
        d = d.reshape(-1, d.shape[-2], d.shape[-1]) if d.dim() != 3 else d
        n, H, W = d.shape
        out = torch.empty(n, self.size, self.size, dtype=torch.float32, device=self.device)
        check(self.lib.svln_preprocess_depth(ptr(d), n, H, W, self.size, self.size, self.min_depth, self.max_depth, ptr(out),
                                             cur_stream()))
        return out

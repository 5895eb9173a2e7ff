"""Slow-memory voxel token pruning — host side over the C-ABI (include/svln.h §1).

The reference plumbs `depths / poses / intrinsics` through `generate -> encode_rgbd`
(streamvln/model/stream_video_vln.py:353-361,102) and never uses them; this is the stage they
were meant for (spec: DESIGN.md §3).  All compute is in libsvln.so; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import INVALID_DROP, INVALID_KEEP, check, cur_stream, ptr

__all__ = ["VoxelPruneConfig", "VoxelPruner", "PruneResult", "INVALID_KEEP", "INVALID_DROP"]


@dataclass
class VoxelPruneConfig:
    grid: int = 14               # pooled tokens per side (14 -> 196 tokens; 27 -> 729)
    dim: int = 3584              # token feature width
    voxel_size: float = 0.2      # metres
    max_depth: float = 10.0      # Habitat depth range 0-10 m (config/vln_r2r.yaml:24-35)
    period: int = 0              # frames per temporal period; <=0 => one global period
    frame_threshold: float = 0.0 # frames keeping < threshold*grid^2 tokens are dropped whole
    invalid_policy: int = INVALID_KEEP
    height: int = 384
    width: int = 384

    @property
    def min_keep(self) -> int:
        return int(math.ceil(self.frame_threshold * self.grid * self.grid)) if self.frame_threshold > 0 else 0


@dataclass
class PruneResult:
    key: torch.Tensor              # [S,T,G,G] int64
    mask: torch.Tensor             # [S*T*G*G] uint8
    idx: torch.Tensor              # [N] int32, first n_keep entries valid (global token index)
    stream_keep_off: torch.Tensor  # [S+1] int32
    n_keep_dev: torch.Tensor       # [1] int32 (device)
    feat: torch.Tensor | None      # [N, D] bf16, first n_keep rows valid

    def n_keep(self) -> int:
        return int(self.n_keep_dev.item())  # host sync: only when the caller needs the count


class VoxelPruner:
    """Owns the scratch and output buffers for pruning up to (max_streams x max_frames) frames."""

    def __init__(self, cfg: VoxelPruneConfig, max_streams: int, max_frames: int, device="cuda"):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VoxelPruner needs a CUDA device (no CPU fallback)")
        self.max_streams, self.max_frames = max_streams, max_frames
        G = cfg.grid
        n = max_streams * max_frames * G * G
        # worst case for the workspace: every frame its own segment vs one global period
        ws = max(self.lib.svln_prune_workspace_bytes(n, max_streams * max_frames, max_frames * G * G),
                 self.lib.svln_prune_workspace_bytes(n, max_streams, max_frames * G * G))
        self.workspace = torch.empty(ws, dtype=torch.uint8, device=self.device)
        self.key = torch.empty(n, dtype=torch.int64, device=self.device)
        self.mask = torch.empty(n, dtype=torch.uint8, device=self.device)
        self.idx = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        self.keep_off = torch.empty(max_streams + 1, dtype=torch.int32, device=self.device)
        self.n_keep = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.feat_out = None

    def prune(self, depth: torch.Tensor, pose: torch.Tensor, intr: torch.Tensor,
              feat: torch.Tensor | None = None, feat_out: torch.Tensor | None = None) -> PruneResult:
        """depth [S,T,H,W] f32, pose/intr [S,T,4,4] f32, feat [S,T,G*G,D] bf16 (all on device)."""
        cfg = self.cfg
        S, T, H, W = depth.shape
        if S > self.max_streams or T > self.max_frames:
            raise ValueError(f"shape ({S},{T}) exceeds pruner capacity ({self.max_streams},{self.max_frames})")
        for name, x, dt in (("depth", depth, torch.float32), ("pose", pose, torch.float32),
                            ("intr", intr, torch.float32)):
            if x.dtype != dt or not x.is_cuda or not x.is_contiguous():
                raise ValueError(f"{name} must be a contiguous {dt} CUDA tensor (geometry stays fp32: "
                                 "the reference's bf16 cast at streamvln_eval.py:328-330 loses ~6 cm at 10 m)")
        G = cfg.grid
        n = S * T * G * G
        if feat is not None:
            if feat.dtype != torch.bfloat16 or not feat.is_contiguous() or feat.numel() != n * cfg.dim:
                raise ValueError("feat must be contiguous bf16 [S,T,G*G,D]")
            if feat_out is None:
                if self.feat_out is None or self.feat_out.shape[0] < n:
                    self.feat_out = torch.empty(max(n, 1), cfg.dim, dtype=torch.bfloat16, device=self.device)
                feat_out = self.feat_out
        check(self.lib.svln_voxel_prune(
            ptr(depth), ptr(pose), ptr(intr), ptr(feat), S, T, H, W, G, cfg.dim, cfg.voxel_size,
            cfg.max_depth, cfg.period, cfg.min_keep, cfg.invalid_policy, ptr(self.workspace),
            self.workspace.numel(), ptr(self.key), ptr(self.mask), ptr(self.idx), ptr(self.keep_off),
            ptr(self.n_keep), ptr(feat_out) if feat is not None else None, cur_stream()))
        return PruneResult(self.key[:n].view(S, T, G, G), self.mask[:n], self.idx, self.keep_off[:S + 1],
                           self.n_keep, feat_out if feat is not None else None)

    # -- stage-level entry points (used by the incremental slow-memory update and the tests)
    def voxel_keys(self, depth, pose, intr) -> torch.Tensor:
        F = depth.shape[0]
        H, W = depth.shape[-2:]
        G = self.cfg.grid
        out = torch.empty(F, G, G, dtype=torch.int64, device=self.device)
        check(self.lib.svln_voxel_keys(ptr(depth), ptr(pose), ptr(intr), F, H, W, G, self.cfg.voxel_size,
                                       self.cfg.max_depth, ptr(out), cur_stream()))
        return out

    def keep_latest(self, key: torch.Tensor, t: torch.Tensor, seg_off: torch.Tensor, max_seg_tokens: int,
                    invalid_policy: int | None = None) -> torch.Tensor:
        n = key.numel()
        mask = torch.empty(n, dtype=torch.uint8, device=self.device)
        need = self.lib.svln_prune_workspace_bytes(n, seg_off.numel() - 1, max_seg_tokens)
        if self.workspace.numel() < need:
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        pol = self.cfg.invalid_policy if invalid_policy is None else invalid_policy
        check(self.lib.svln_keep_latest(ptr(key), ptr(t), n, ptr(seg_off), seg_off.numel() - 1, max_seg_tokens,
                                        pol, ptr(self.workspace), self.workspace.numel(), ptr(mask), cur_stream()))
        return mask

    def compact(self, mask: torch.Tensor, range_off: torch.Tensor | None = None):
        n = mask.numel()
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        n_keep = torch.zeros(1, dtype=torch.int32, device=self.device)
        r_out = None
        n_ranges = 0
        if range_off is not None:
            n_ranges = range_off.numel() - 1
            r_out = torch.empty(n_ranges + 1, dtype=torch.int32, device=self.device)
        check(self.lib.svln_compact_mask(ptr(mask), n, ptr(self.workspace), self.workspace.numel(), ptr(idx),
                                         ptr(n_keep), ptr(range_off), n_ranges, ptr(r_out), cur_stream()))
        return idx, n_keep, r_out

    def gather(self, src: torch.Tensor, idx: torch.Tensor, n_rows_dev: torch.Tensor, max_rows: int,
               out: torch.Tensor | None = None) -> torch.Tensor:
        row_bytes = src.shape[-1] * src.element_size()
        if out is None:
            out = torch.empty(max(max_rows, 1), src.shape[-1], dtype=src.dtype, device=self.device)
        check(self.lib.svln_gather_rows(ptr(src), row_bytes, ptr(idx), ptr(n_rows_dev), max_rows, ptr(out),
                                        cur_stream()))
        return out


def voxel_prune_host(cfg: VoxelPruneConfig, depth, pose, intr, feat_u16, staging: torch.Tensor | None = None):
    """Host-buffer entry (numpy in, numpy out) through svln_voxel_prune_host: the H2D / D2H copies are
    inside the call.  depth [S,T,H,W] f32, pose/intr [S,T,4,4] f32, feat_u16 [S,T,G*G,D] uint16."""
    import numpy as np

    lib = _lib.load()
    S, T, H, W = depth.shape
    G, D = cfg.grid, cfg.dim
    n = S * T * G * G
    need = lib.svln_voxel_prune_host_staging_bytes(S, T, H, W, G, D)
    if staging is None or staging.numel() < need:
        staging = torch.empty(need, dtype=torch.uint8, device="cuda")
    depth = np.ascontiguousarray(depth, np.float32)
    pose = np.ascontiguousarray(pose, np.float32)
    intr = np.ascontiguousarray(intr, np.float32)
    mask = np.empty(n, np.uint8)
    idx = np.empty(max(n, 1), np.int32)
    off = np.empty(S + 1, np.int32)
    n_keep = np.zeros(1, np.int32)
    feat_out = None
    if feat_u16 is not None:
        feat_u16 = np.ascontiguousarray(feat_u16, np.uint16)
        feat_out = np.empty((max(n, 1), D), np.uint16)

    def hp(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    check(lib.svln_voxel_prune_host(hp(depth), hp(pose), hp(intr), hp(feat_u16), S, T, H, W, G, D,
                                    cfg.voxel_size, cfg.max_depth, cfg.period, cfg.min_keep, cfg.invalid_policy,
                                    ptr(staging), staging.numel(), hp(mask), hp(idx), hp(off), hp(n_keep),
                                    hp(feat_out), cur_stream()))
    k = int(n_keep[0])
    return dict(mask=mask, idx=idx[:k], stream_keep_off=off, n_keep=k,
                feat=None if feat_out is None else feat_out[:k], staging=staging)

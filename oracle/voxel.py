"""CPU oracle for slow-memory voxel token pruning.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under streamvln_b200/ does.

PARITY UNPINNED.  The released reference has no voxel pruning (streamvln/streamvln_agent.py:171,
"Currently the model does not use voxel pooling"; streamvln/model/stream_video_vln.py:102-142
accepts depths/poses/intrinsics and ignores them) and no golden vectors for it.  The normative
statement of the spec is oracle/voxel_oracle.c; this module holds

  * `c_*`     : ctypes wrappers of that C restatement (fast; used for full-size CPU baselines);
  * `np_*`    : an independent numpy restatement of the same spec, used to cross-check the C code
                (two implementations, two algorithms: sort/unique here, qsort there, a hash table
                on the GPU);
  * `torch_reference_idiom_prune`: the same result written the way the reference's PyTorch code
                would write it (torch.unique / scatter-max / boolean index), which is what
                bench.py times as the "reference PyTorch CPU path" for this stage.

Input conventions follow streamvln/streamvln_eval.py:126-185,254-283: depth in metres resampled to
the 384x384 processor grid, 4x4 intrinsics rescaled to that grid, 4x4 camera->episodic pose.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libvoxel_oracle.so"
_lib = None

KEY_INVALID = np.int64(-1)
INVALID_KEEP, INVALID_DROP = 0, 1


def build(force: bool = False) -> Path:
    src = _HERE / "voxel_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.oracle_voxel_prune.restype = C.c_int64
        _lib.oracle_compact_gather.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- C wrappers
def c_voxel_keys(depth, pose, intr, G, voxel_size, max_depth=10.0):
    depth = np.ascontiguousarray(depth, np.float32)
    F, H, W = depth.shape
    pose = np.ascontiguousarray(pose, np.float32).reshape(F, 16)
    intr = np.ascontiguousarray(intr, np.float32).reshape(F, 16)
    out = np.empty((F, G, G), np.int64)
    _load().oracle_voxel_keys(_p(depth), _p(pose), _p(intr), C.c_int64(F), H, W, G,
                              C.c_float(voxel_size), C.c_float(max_depth), _p(out))
    return out


def c_keep_latest(key, t, seg_off, invalid_policy=INVALID_KEEP):
    key = np.ascontiguousarray(key, np.int64).ravel()
    t = np.ascontiguousarray(t, np.int32).ravel()
    seg_off = np.ascontiguousarray(seg_off, np.int64)
    mask = np.zeros(key.shape[0], np.uint8)
    _load().oracle_keep_latest(_p(key), _p(t), C.c_int64(key.shape[0]), _p(seg_off),
                               C.c_int64(seg_off.shape[0] - 1), int(invalid_policy), _p(mask))
    return mask


def c_voxel_prune(depth, pose, intr, feat_u16, G, voxel_size, max_depth=10.0, period=0, min_keep=0,
                  invalid_policy=INVALID_KEEP):
    """depth [S,T,H,W]; pose/intr [S,T,4,4]; feat_u16 [S,T,G*G,D] uint16 view of bf16 (or None).
    Returns dict(key, mask, idx, stream_keep_off, n_keep, feat)."""
    depth = np.ascontiguousarray(depth, np.float32)
    S, T, H, W = depth.shape
    pose = np.ascontiguousarray(pose, np.float32).reshape(S * T, 16)
    intr = np.ascontiguousarray(intr, np.float32).reshape(S * T, 16)
    N = S * T * G * G
    D = 0 if feat_u16 is None else feat_u16.shape[-1]
    if feat_u16 is not None:
        feat_u16 = np.ascontiguousarray(feat_u16, np.uint16)
    key = np.empty((S, T, G, G), np.int64)
    mask = np.zeros(N, np.uint8)
    idx = np.empty(max(N, 1), np.int32)
    off = np.zeros(S + 1, np.int32)
    feat_out = None if feat_u16 is None else np.empty((max(N, 1), D), np.uint16)
    n_keep = _load().oracle_voxel_prune(
        _p(depth), _p(pose), _p(intr), _p(feat_u16), S, T, H, W, G, D, C.c_float(voxel_size),
        C.c_float(max_depth), int(period), int(min_keep), int(invalid_policy), _p(key), _p(mask),
        _p(idx), _p(off), _p(feat_out))
    return dict(key=key, mask=mask, idx=idx[:n_keep].copy(), stream_keep_off=off, n_keep=int(n_keep),
                feat=None if feat_out is None else feat_out[:n_keep].copy())


# --------------------------------------------------------------------- numpy restatement
def np_voxel_keys(depth, pose, intr, G, voxel_size, max_depth=10.0):
    """fp32, fixed op order (every intermediate is a float32 array, so no FMA contraction)."""
    f32 = np.float32
    depth = np.asarray(depth, f32)
    F, H, W = depth.shape
    P = np.asarray(pose, f32).reshape(F, 4, 4)
    K = np.asarray(intr, f32).reshape(F, 4, 4)
    sw, sh = f32(W) / f32(G), f32(H) / f32(G)
    j = np.arange(G, dtype=f32)
    u = (j + f32(0.5)) * sw                      # [G]
    v = (j + f32(0.5)) * sh
    px = np.clip(np.floor(u).astype(np.int64), 0, W - 1)
    py = np.clip(np.floor(v).astype(np.int64), 0, H - 1)
    d = depth[:, py][:, :, px]                   # [F,G,G]  (row i -> v, col j -> u)
    fx, cx = K[:, 0, 0][:, None, None], K[:, 0, 2][:, None, None]
    fy, cy = K[:, 1, 1][:, None, None], K[:, 1, 2][:, None, None]
    uu = np.broadcast_to(u[None, None, :], d.shape)
    vv = np.broadcast_to(v[None, :, None], d.shape)
    with np.errstate(all="ignore"):
        valid = (d > 0) & (d < f32(max_depth)) & np.isfinite(d)
        ds = np.where(valid, d, f32(1.0))
        xc = ds * ((uu - cx) / fx)
        yc = ds * ((vv - cy) / fy)
        zc = ds
        out = []
        for r in range(3):
            a = P[:, r, 0][:, None, None] * xc
            b = P[:, r, 1][:, None, None] * yc
            c = P[:, r, 2][:, None, None] * zc
            w = ((a + b) + c) + P[:, r, 3][:, None, None]
            q = np.floor(w / f32(voxel_size))
            q = np.fmin(np.fmax(q, f32(-1048576.0)), f32(1048575.0))
            out.append(q.astype(np.int64) + 1048576)
    key = (out[0] << 42) | (out[1] << 21) | out[2]
    return np.where(valid, key, KEY_INVALID).astype(np.int64)


def np_keep_latest(key, t, seg_off, invalid_policy=INVALID_KEEP):
    key = np.asarray(key, np.int64).ravel()
    t = np.asarray(t, np.int64).ravel()
    mask = np.zeros(key.shape[0], np.uint8)
    for g in range(len(seg_off) - 1):
        a, b = int(seg_off[g]), int(seg_off[g + 1])
        k, tt = key[a:b], t[a:b]
        inv = k < 0
        mask[a:b][inv] = 1 if invalid_policy == INVALID_KEEP else 0
        loc = np.nonzero(~inv)[0]
        if loc.size == 0:
            continue
        # sort by (key asc, t desc, index asc); first of each key run wins
        order = np.lexsort((loc, -tt[loc], k[loc]))
        ks = k[loc][order]
        first = np.ones(ks.shape[0], bool)
        first[1:] = ks[1:] != ks[:-1]
        mask[a + loc[order][first]] = 1
    return mask


def np_voxel_prune(depth, pose, intr, feat_u16, G, voxel_size, max_depth=10.0, period=0, min_keep=0,
                   invalid_policy=INVALID_KEEP):
    depth = np.asarray(depth, np.float32)
    S, T, H, W = depth.shape
    tpf = G * G
    key = np_voxel_keys(depth.reshape(S * T, H, W), np.asarray(pose).reshape(S * T, 4, 4),
                        np.asarray(intr).reshape(S * T, 4, 4), G, voxel_size, max_depth)
    per = T if (period <= 0 or period > T) else period
    seg_off = []
    for s in range(S):
        for p0 in range(0, T, per):
            seg_off.append((s * T + p0) * tpf)
    seg_off.append(S * T * tpf)
    t = np.tile(np.repeat(np.arange(T, dtype=np.int32), tpf), S)
    mask = np_keep_latest(key, t, np.asarray(seg_off, np.int64), invalid_policy)
    if min_keep > 0:
        m = mask.reshape(S * T, tpf)
        m[m.sum(1) < min_keep] = 0
        mask = m.reshape(-1)
    idx = np.nonzero(mask)[0].astype(np.int32)
    off = np.zeros(S + 1, np.int32)
    off[1:] = np.cumsum(mask.reshape(S, -1).sum(1)) if S else 0
    feat = None if feat_u16 is None else np.asarray(feat_u16).reshape(S * T * tpf, -1)[idx]
    return dict(key=key.reshape(S, T, G, G), mask=mask, idx=idx, stream_keep_off=off, n_keep=int(idx.size),
                feat=feat)


# -------------------------------------------------- reference-idiom PyTorch (CPU baseline)
def torch_reference_idiom_prune(depth, pose, intr, feat, G, voxel_size, max_depth=10.0):
    """One stream, global period, invalid tokens kept: the pruning stage written in the reference's
    own PyTorch idiom (batched tensor ops, torch.unique, scatter-reduce 'amax', boolean indexing —
    the style of stream_video_vln.py:102-142,182-238).  depth [T,H,W], pose/intr [T,4,4] float32
    tensors, feat [T,G*G,D] bf16 tensor.  Returns (mask [T*G*G] bool, kept feats [n_keep, D])."""
    import torch

    T, H, W = depth.shape
    dev = depth.device
    sw, sh = torch.tensor(W / G, dtype=torch.float32), torch.tensor(H / G, dtype=torch.float32)
    j = torch.arange(G, dtype=torch.float32, device=dev)
    u = (j + 0.5) * sw
    v = (j + 0.5) * sh
    px = u.floor().long().clamp(0, W - 1)
    py = v.floor().long().clamp(0, H - 1)
    d = depth[:, py][:, :, px]
    valid = (d > 0) & (d < max_depth) & torch.isfinite(d)
    ds = torch.where(valid, d, torch.ones_like(d))
    fx, cx, fy, cy = (intr[:, 0, 0], intr[:, 0, 2], intr[:, 1, 1], intr[:, 1, 2])
    xc = ds * ((u[None, None, :] - cx[:, None, None]) / fx[:, None, None])
    yc = ds * ((v[None, :, None] - cy[:, None, None]) / fy[:, None, None])
    comps = []
    for r in range(3):
        w = ((pose[:, r, 0, None, None] * xc + pose[:, r, 1, None, None] * yc) + pose[:, r, 2, None, None] * ds) \
            + pose[:, r, 3, None, None]
        q = torch.floor(w / voxel_size).clamp(-1048576.0, 1048575.0).long() + 1048576
        comps.append(q)
    key = ((comps[0] << 42) | (comps[1] << 21) | comps[2]).reshape(-1)
    validf = valid.reshape(-1)
    n = key.numel()
    tpf = G * G
    tok = torch.arange(n, device=dev)
    t = tok // tpf
    local = tok % tpf
    rank = t * tpf + (tpf - 1 - local)           # latest frame, then lowest in-frame index
    uniq, inv = torch.unique(key[validf], return_inverse=True)
    best = torch.full((uniq.numel(),), -1, dtype=torch.long, device=dev)
    best = best.scatter_reduce(0, inv, rank[validf], reduce="amax")
    mask = torch.ones(n, dtype=torch.bool, device=dev)
    mask[validf] = best[inv] == rank[validf]
    return mask, feat.reshape(n, -1)[mask]

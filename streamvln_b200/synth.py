"""Synthetic R2R/RxR-shaped RGB-D streams (no dataset, no simulator is available offline).

Geometry conventions are the reference's Habitat eval loop (streamvln/streamvln_eval.py):
  * intrinsics: fx = (W/2)/tan(hfov/2), cx = (W-1)/2, cy = (H-1)/2 for the 640x480 sensor
    (:137-152), rows 0/1 rescaled to the 384x384 processor grid (:154-168);
  * pose: xyz_yaw_to_tf_matrix([x, -y, cam_h], yaw) @ axis_align (:170-185, :260-262, :282),
    i.e. OpenCV camera (x right, y down, z forward) -> episodic frame (x forward, y left, z up);
  * depth: metres on the 384x384 grid (:126-135, :254-256); actions 0.25 m / 15 deg
    (config/vln_r2r.yaml:34-35); camera height 1.25 m (streamvln_agent.py:297).
"""
from __future__ import annotations

import math

import numpy as np

SENSOR_W, SENSOR_H, HFOV_DEG = 640, 480, 79.0
GRID = 384
CAM_HEIGHT = 1.25
FORWARD_M, TURN_DEG = 0.25, 15.0

AXIS_ALIGN = np.array([[0, 0, 1, 0], [-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float64)


def intrinsic_matrix(width=SENSOR_W, height=SENSOR_H, hfov_deg=HFOV_DEG, target=GRID) -> np.ndarray:
    fx = (width / 2.0) / np.tan(np.deg2rad(hfov_deg / 2.0))
    K = np.array([[fx, 0, (width - 1.0) / 2.0, 0], [0, fx, (height - 1.0) / 2.0, 0],
                  [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    K[0] /= width / target
    K[1] /= height / target
    return K.astype(np.float32)


def pose_matrix(x: float, y: float, yaw: float, cam_h: float = CAM_HEIGHT) -> np.ndarray:
    c, s = math.cos(yaw), math.sin(yaw)
    tf = np.array([[c, -s, 0, x], [s, c, 0, -y], [0, 0, 1, cam_h], [0, 0, 0, 1]], np.float64)
    return (tf @ AXIS_ALIGN).astype(np.float32)


def preprocess_depth(depth01: np.ndarray, min_depth: float = 0.0, max_depth: float = 10.0, size: int = GRID) -> np.ndarray:
    """Habitat depth in [0,1] -> metres on the size x size processor grid, as the reference's eval loop does it
    (streamvln_eval.py:254-256 scale to metres then millimetres, :276 uint16 truncation, :126-135 PIL NEAREST squash to
    384x384 and / 1000; `filter_depth(blur_type=None)` of depth-camera-filtering is not reproduced).  Host reference of
    `svln_preprocess_depth`; pinned by tests/golden/ref_geometry.npz."""
    d = np.asarray(depth01, np.float32)
    d = d.reshape(d.shape[:2])
    mm = ((d * np.float32(max_depth - min_depth) + np.float32(min_depth)) * np.float32(1000)).astype(np.uint16)
    h, w = mm.shape
    iy = ((np.arange(size) + 0.5) * (h / size)).astype(np.int64)     # Pillow NEAREST: (int)((i + 0.5) * scale)
    ix = ((np.arange(size) + 0.5) * (w / size)).astype(np.int64)
    return (mm[iy][:, ix] / 1000).astype(np.float32)


def action_walk(n_steps: int, rng: np.random.Generator):
    """Seeded R2R-style walk: p(forward)=.6, p(left)=p(right)=.2.  Returns actions and poses."""
    acts = rng.choice(np.array([1, 2, 3]), size=n_steps, p=[0.6, 0.2, 0.2])
    x = y = yaw = 0.0
    poses = np.empty((n_steps, 4, 4), np.float32)
    for i, a in enumerate(acts):
        poses[i] = pose_matrix(x, y, yaw)
        if a == 1:
            x += FORWARD_M * math.cos(yaw)
            y -= FORWARD_M * math.sin(yaw)  # habitat GPS: west negative (eval loop flips y)
        elif a == 2:
            yaw += math.radians(TURN_DEG)
        else:
            yaw -= math.radians(TURN_DEG)
    return acts.astype(np.int32), poses


def _bilinear_up(noise: np.ndarray, size: int) -> np.ndarray:
    n = noise.shape[-1]
    pos = (np.arange(size, dtype=np.float64) + 0.5) * n / size - 0.5
    i0 = np.clip(np.floor(pos).astype(np.int64), 0, n - 1)
    i1 = np.clip(i0 + 1, 0, n - 1)
    w = np.clip(pos - i0, 0.0, 1.0)
    rows = noise[..., i0, :] * (1 - w)[:, None] + noise[..., i1, :] * w[:, None]
    return rows[..., :, i0] * (1 - w) + rows[..., :, i1] * w


def depth_frames(n_frames: int, rng: np.random.Generator, size: int = GRID, lo=0.5, hi=5.0,
                 zero_frac=0.02) -> np.ndarray:
    """Smooth depth fields U(lo,hi) m (bilinear-upsampled 12x12 noise) with zero_frac invalid pixels."""
    noise = rng.uniform(lo, hi, size=(n_frames, 12, 12))
    d = np.ascontiguousarray(_bilinear_up(noise, size), dtype=np.float32)
    holes = rng.random(size=d.shape) < zero_frac
    d[holes] = 0.0
    return d


def make_stream(n_frames: int = 32, G: int = 14, D: int = 3584, seed: int = 0, with_feat: bool = True):
    """One synthetic stream: dict(depth [T,384,384] f32, pose [T,4,4] f32, intr [T,4,4] f32,
    actions [T] i32, feat [T,G*G,D] bf16-as-uint16 or None)."""
    rng = np.random.default_rng(seed)
    actions, poses = action_walk(n_frames, rng)
    depth = depth_frames(n_frames, rng)
    intr = np.broadcast_to(intrinsic_matrix(), (n_frames, 4, 4)).copy()
    feat = None
    if with_feat:
        f = rng.standard_normal(size=(n_frames, G * G, D), dtype=np.float32)
        feat = f32_to_bf16_u16(f)
    return dict(depth=depth, pose=poses, intr=intr, actions=actions, feat=feat)


def make_streams(n_streams: int, n_frames: int = 32, G: int = 14, D: int = 3584, seed: int = 0,
                 with_feat: bool = True):
    """S streams stacked: depth [S,T,384,384], pose/intr [S,T,4,4], feat [S,T,G*G,D] uint16."""
    parts = [make_stream(n_frames, G, D, seed + 1000 * s, with_feat) for s in range(n_streams)]
    out = {k: np.stack([p[k] for p in parts]) for k in ("depth", "pose", "intr", "actions")}
    out["feat"] = np.stack([p["feat"] for p in parts]) if with_feat else None
    return out


def f32_to_bf16_u16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit patterns (uint16), same rounding as torch."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (rounded >> np.uint32(16)).astype(np.uint16)


def bf16_u16_to_f32(x: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(x, np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)

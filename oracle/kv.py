"""CPU oracle for the fast-streaming KV arena ops.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`np_rope_hf` restates HF transformers' Qwen2 `apply_rotary_pos_emb` in bf16 (transformers==4.45.1 is
pinned by the reference's requirements.txt:140 and is not vendored; call site
streamvln/model/stream_video_vln.py:340) — every elementwise product and the sum are rounded to
bf16, as eager torch does.  `np_append` restates `DynamicCache.update` (= torch.cat along the sequence).

`np_kv_compact` / `np_rerotate` are PARITY UNPINNED: the reference has no window eviction, no in-place
compaction, no memory splice and no RoPE remap (it drops the whole cache every
num_frames steps, stream_video_vln.py:473-479).  They are the normative spec.
"""
from __future__ import annotations

import numpy as np


def f32_to_bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    return (r >> np.uint32(16)).astype(np.uint16)


def bf16_to_f32(x):
    return (np.ascontiguousarray(x, np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def rbf(x):
    """round an fp32 array to bf16 precision, staying in fp32"""
    return bf16_to_f32(f32_to_bf16(x))


def np_rope_hf(x_u16, cos_u16, sin_u16):
    """x [..., 128] bf16 bits, cos/sin [..., 128] bf16 bits (HF layout: halves duplicated).
    out = bf16( bf16(x*cos) + bf16(rotate_half(x)*sin) ), rotate_half(x) = cat(-x[64:], x[:64])."""
    x = bf16_to_f32(x_u16)
    c = bf16_to_f32(cos_u16)
    s = bf16_to_f32(sin_u16)
    half = x.shape[-1] // 2
    rot = np.concatenate([-x[..., half:], x[..., :half]], axis=-1)
    return f32_to_bf16(rbf(x * c) + rbf(rot * s))


def np_append_rope(arena_u16, layer, qkv_u16, n_q_head, tok_env, tok_pos, cos_u16, sin_u16):
    """arena [E, Lyr, 2, Hkv, max_len, 128] uint16 (modified in place); qkv [n, (Hq+2Hkv)*128].
    Returns rotated q [n, Hq, 128]."""
    E, Lyr, _, Hkv, max_len, hd = arena_u16.shape
    n = qkv_u16.shape[0]
    x = qkv_u16.reshape(n, n_q_head + 2 * Hkv, hd)
    q, k, v = x[:, :n_q_head], x[:, n_q_head:n_q_head + Hkv], x[:, n_q_head + Hkv:]
    cos = cos_u16[tok_pos][:, None, :]
    sin = sin_u16[tok_pos][:, None, :]
    q_rot = np_rope_hf(q, np.broadcast_to(cos, q.shape), np.broadcast_to(sin, q.shape))
    k_rot = np_rope_hf(k, np.broadcast_to(cos, k.shape), np.broadcast_to(sin, k.shape))
    for i in range(n):
        arena_u16[tok_env[i], layer, 0, :, tok_pos[i], :] = k_rot[i]
        arena_u16[tok_env[i], layer, 1, :, tok_pos[i], :] = v[i]
    return q_rot


def _rotate(plane_f32, m, cos_delta, sin_delta, sgn):
    """rotate rows of plane [n,128] (fp32 values of bf16) by table rows m[n]; sgn=+1 forward."""
    lo, hi = plane_f32[:, :64], plane_f32[:, 64:]
    c = cos_delta[m]
    s = sin_delta[m] * np.float32(sgn)
    nl = lo * c - hi * s
    nh = hi * c + lo * s
    return np.concatenate([nl, nh], axis=1).astype(np.float32)


def np_kv_compact(arena_u16, env, first_row, new_len, src_row, rerotate, cos_delta, sin_delta):
    """In-place compaction of stream `env` over all layers: row d <- row src_row[d] for
    d in [first_row, new_len); K rows rotated by (d - src) positions when rerotate."""
    E, Lyr, _, Hkv, max_len, hd = arena_u16.shape
    d = np.arange(first_row, new_len)
    s = np.asarray(src_row[first_row:new_len], np.int64)
    assert np.all(s >= d) and np.all(np.diff(s) > 0)
    for layer in range(Lyr):
        for kv in range(2):
            for h in range(Hkv):
                plane = arena_u16[env, layer, kv, h]
                rows = plane[s].copy()
                if rerotate and kv == 0:
                    m = (s - d)
                    moved = m > 0
                    if moved.any():
                        f = bf16_to_f32(rows[moved])
                        rows[moved] = f32_to_bf16(_rotate(f, m[moved], cos_delta, sin_delta, -1.0))
                plane[d] = rows


def np_rerotate(arena_u16, env, row0, delta_pos, cos_delta, sin_delta):
    E, Lyr, _, Hkv, max_len, hd = arena_u16.shape
    delta_pos = np.asarray(delta_pos, np.int64)
    rows = row0 + np.arange(delta_pos.shape[0])
    nz = delta_pos != 0
    if not nz.any():
        return
    for layer in range(Lyr):
        for h in range(Hkv):
            plane = arena_u16[env, layer, 0, h]
            f = bf16_to_f32(plane[rows[nz]])
            m = np.abs(delta_pos[nz])
            out = np.empty_like(f)
            for sgn in (1.0, -1.0):
                sel = (delta_pos[nz] > 0) if sgn > 0 else (delta_pos[nz] < 0)
                if sel.any():
                    out[sel] = _rotate(f[sel], m[sel], cos_delta, sin_delta, sgn)
            plane[rows[nz]] = f32_to_bf16(out)

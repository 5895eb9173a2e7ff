"""Per-kernel roofline measurements at the batched shapes of BASELINE.json configs[2]/[4] (64 concurrent
episodes), CUDA events on the launching stream, L2 flushed before every timed call.  Used by bench.py
(`kernels` key) and scripts; algorithmic bytes / FLOPs per unit are SURVEY.md §8d's figures.
"""
from __future__ import annotations

import time
from typing import Callable, Dict

import torch

KV_BYTES_PER_TOKEN = 2 * 28 * 4 * 128 * 2  # 57 344 B: K+V, 28 layers, 4 kv heads, 128 dims, bf16


_FLUSH = None


def _time(fn: Callable[[], None], iters: int = 10, warmup: int = 3) -> float:
    """average milliseconds per call: one CUDA-event pair per call on the current stream, with the 126 MB L2 flushed
    (a 256 MB buffer rewritten) before every timed call, so no call finds its inputs cached by the previous one"""
    import os

    global _FLUSH
    if os.environ.get("SVLN_MICRO_ITERS"):  # profiling runs: launch each kernel only a couple of times
        iters, warmup = int(os.environ["SVLN_MICRO_ITERS"]), 1
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    pairs = []
    for _ in range(iters):
        _FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in pairs) / iters


def _row(name, bound, work, ms, peak, unit, note=""):
    ach = work / (ms / 1e3) / (1e9 if unit == "GB/s" else 1e12)
    return {"kernel": name, "bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit,
            "frac": round(ach / peak, 4), "ms": round(ms, 4), "note": note}


def bench_prune(device, pk, S=64, T=32, G=14, D=3584, voxel=0.2) -> Dict:
    import numpy as np

    from . import synth
    from .voxel import INVALID_DROP, VoxelPruneConfig, VoxelPruner

    st = synth.make_streams(S, T, G, 8, seed=0, with_feat=False)
    cfg = VoxelPruneConfig(grid=G, dim=D, voxel_size=voxel, invalid_policy=INVALID_DROP)
    pr = VoxelPruner(cfg, S, T, device)
    depth = torch.from_numpy(st["depth"]).to(device)
    pose = torch.from_numpy(st["pose"]).to(device)
    intr = torch.from_numpy(st["intr"]).to(device)
    feat = torch.randn(S, T, G * G, D, device=device, dtype=torch.bfloat16)
    res = pr.prune(depth, pose, intr, feat)
    n_in, n_keep = S * T * G * G, res.n_keep()
    ms_all = _time(lambda: pr.prune(depth, pose, intr, feat))
    ms_gather = _time(lambda: pr.gather(feat.view(-1, D), res.idx, res.n_keep_dev, n_in, res.feat))
    ms_index = _time(lambda: pr.prune(depth, pose, intr, None))
    bytes_all = n_in * 25 + n_keep * (4 * D + 4)
    rows = [_row("svln_voxel_prune (keys+dedup+compact+gather)", "hbm", bytes_all, ms_all, pk["hbm_gbs"], "GB/s",
                 f"{S} streams x {T} frames x {G * G} tok x {D}; kept {n_keep}/{n_in}; {n_in / ms_all / 1e6:.3f} Gtok/s"),
            _row("gather_rows_kernel", "hbm", n_keep * 4 * D, ms_gather, pk["hbm_gbs"], "GB/s", "read+write of kept rows"),
            _row("keys+dedup+compact (index stages)", "hbm", n_in * 25, ms_index, pk["hbm_gbs"], "GB/s",
                 "latency/atomics-bound: 25 B per candidate token")]
    return {"rows": rows, "gtok_per_s": round(n_in / ms_all / 1e6, 4), "n_in": n_in, "n_keep": n_keep}


def bench_kv_attn(device, pk, n_env=64, L=3000, n_new=217) -> Dict:
    from .kv_arena import KvArena

    Hq, Hkv, hd, layers = 28, 4, 128, 28
    ar = KvArena(n_env, layers, Hkv, hd, 4096, 1e6, device)
    for e in range(n_env):  # finite random contents
        ar.buf[e].normal_(0, 1)
    rows = []
    # fused RoPE + append: one layer, n_env x n_new tokens
    n_tok = n_env * n_new
    qkv = torch.randn(n_tok, (Hq + 2 * Hkv) * hd, device=device, dtype=torch.bfloat16)
    tok_env = torch.arange(n_env, device=device, dtype=torch.int32).repeat_interleave(n_new)
    tok_pos = (L - n_new + torch.arange(n_new, device=device, dtype=torch.int32)).repeat(n_env)
    q_out = torch.empty(n_tok, Hq, hd, device=device, dtype=torch.bfloat16)
    ms = _time(lambda: ar.append_rope(3, qkv, Hq, tok_env, tok_pos, q_out))
    rows.append(_row("kv_append_rope_kernel", "hbm", 2 * n_tok * (Hq + 2 * Hkv) * hd * 2, ms, pk["hbm_gbs"], "GB/s",
                     f"{n_env} streams x {n_new} tokens, one layer; read qkv + write q/K/V"))
    del qkv, q_out
    # the same kernel at the window-start shape (9 views x 196 tokens + text per stream): 0.9 GB per launch
    n_ws = 1900
    ws_env = max(1, n_env // 4)
    n_tok = ws_env * n_ws
    qkv = torch.randn(n_tok, (Hq + 2 * Hkv) * hd, device=device, dtype=torch.bfloat16)
    tok_env_w = torch.arange(ws_env, device=device, dtype=torch.int32).repeat_interleave(n_ws)
    tok_pos_w = torch.arange(n_ws, device=device, dtype=torch.int32).repeat(ws_env)
    q_out = torch.empty(n_tok, Hq, hd, device=device, dtype=torch.bfloat16)
    ms = _time(lambda: ar.append_rope(3, qkv, Hq, tok_env_w, tok_pos_w, q_out))
    rows.append(_row("kv_append_rope_kernel, window start", "hbm", 2 * n_tok * (Hq + 2 * Hkv) * hd * 2, ms, pk["hbm_gbs"], "GB/s",
                     f"{ws_env} streams x {n_ws} tokens, one layer; read qkv + write q/K/V"))
    del qkv, q_out
    # eviction + compaction with re-rotation: drop one 221-row turn and a third of a 1568-row memory region
    P, M, turn = 60, 1568, 221
    keep_mem = torch.arange(M)[torch.arange(M) % 3 != 0]
    src = torch.cat([torch.arange(P), P + keep_mem, torch.arange(P + M + turn, L)]).to(torch.int32)
    new_len = src.numel()
    first = int((src != torch.arange(new_len)).nonzero()[0])
    src_all = torch.zeros(n_env, 4096, dtype=torch.int32)
    src_all[:, :new_len] = src
    src_all = src_all.to(device)
    job_env = torch.arange(n_env, device=device, dtype=torch.int32)
    first_t = torch.full((n_env,), first, device=device, dtype=torch.int32)
    newlen_t = torch.full((n_env,), new_len, device=device, dtype=torch.int32)
    moved = n_env * (new_len - first)
    ms = _time(lambda: ar.compact(job_env, first_t, newlen_t, src_all, True), iters=5)
    rows.append(_row("kv_compact_kernel (+K re-rotation)", "hbm", 2 * KV_BYTES_PER_TOKEN * moved, ms, pk["hbm_gbs"], "GB/s",
                     f"{n_env} streams, {new_len - first} moved rows each, all 28 layers, in place"))
    # split-KV decode attention: one layer, every stream one token over L keys
    q = torch.randn(n_env, Hq, hd, device=device, dtype=torch.bfloat16)
    out = torch.empty(n_env, Hq * hd, device=device, dtype=torch.bfloat16)
    ones = torch.ones(n_env, device=device, dtype=torch.int32)
    kv_len = ones * L
    ar.reserve_workspace(n_env, Hq)
    layer_i = [0]

    def dec():
        layer_i[0] = (layer_i[0] + 1) % layers  # rotate layers: 14 MB per stream-layer, 0.9 GB per launch
        ar.attn_decode(layer_i[0], q, Hq, job_env, job_env, ones, kv_len, 1, L, hd ** -0.5, out)
    ms = _time(dec, iters=28)
    rows.append(_row("attn_decode_kernel (split-KV)", "hbm", n_env * L * Hkv * 2 * hd * 2, ms, pk["hbm_gbs"], "GB/s",
                     f"{n_env} streams x 1 token over {L} keys, one layer (K+V read once, GQA-shared)"))
    one = torch.ones(1, device=device, dtype=torch.int32)
    ms = _time(lambda: ar.attn_decode(5, q[:1], Hq, one * 0, one * 0, one, one * L, 1, L, hd ** -0.5, out[:1]), iters=20)
    rows.append(_row("attn_decode_kernel, 1 stream", "hbm", L * Hkv * 2 * hd * 2, ms, pk["hbm_gbs"], "GB/s",
                     "single episode: 3 MB per launch, launch-latency-bound"))
    # tcgen05 prefill attention
    def prefill(n_seq, nq, kvl, name):
        qq = torch.randn(n_seq * nq, Hq, hd, device=device, dtype=torch.bfloat16)
        oo = torch.empty(n_seq * nq, Hq * hd, device=device, dtype=torch.bfloat16)
        se, qs = list(range(n_seq)), [i * nq for i in range(n_seq)]
        ms = _time(lambda: ar.attn_prefill(7, qq, Hq, se, qs, [nq] * n_seq, [kvl] * n_seq, hd ** -0.5, oo), iters=5)
        flops = n_seq * Hq * 4 * hd * nq * ((kvl - nq) + nq / 2)
        rows.append(_row("attn_prefill_kernel (tcgen05) " + name, "tensor", flops, ms, pk["bf16_tflops"], "TFLOP/s",
                         f"{n_seq} streams x n_q={nq} over kv_len={kvl}, one layer"))
    prefill(n_env, n_new, L, "steady turn x64")
    prefill(16, 1900, 1900, "window start x16")
    prefill(1, n_new, L, "steady turn x1")
    prefill(1, 1900, 1900, "window start x1")
    return {"rows": rows}


def run_all(device, pk) -> Dict:
    out = {"peak_source": pk["source"]}
    t = time.time()
    try:
        out["prune"] = bench_prune(device, pk)
        torch.cuda.empty_cache()
        out["kv_attn"] = bench_kv_attn(device, pk)
    except Exception as e:  # a microbench failure must not lose the headline line
        out["error"] = repr(e)
    out["seconds"] = round(time.time() - t, 1)
    return out

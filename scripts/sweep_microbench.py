"""BASELINE.json configs[3] (long-horizon RxR-shaped streams) and configs[4] (microbench sweep): pruning and
KV-kernel throughput vs tokens/frame, voxel size, stream count, stream length and window size, with the
reference-idiom PyTorch CPU path timed beside a subset.  Writes a markdown table (stdout) for profiles/."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

import bench
from streamvln_b200 import synth
from streamvln_b200.kv_arena import KvArena
from streamvln_b200.microbench import KV_BYTES_PER_TOKEN, _time
from streamvln_b200.voxel import INVALID_DROP, VoxelPruneConfig, VoxelPruner

dev = torch.device("cuda:0")
pk = bench.peaks()
D = 3584
quick = "--quick" in sys.argv


def prune_case(S, T, G, voxel, cpu=False):
    st = synth.make_streams(S, T, G, 8, seed=0, with_feat=False)
    cfg = VoxelPruneConfig(grid=G, dim=D, voxel_size=voxel, invalid_policy=INVALID_DROP)
    pr = VoxelPruner(cfg, S, T, dev)
    depth, pose, intr = (torch.from_numpy(st[k]).to(dev) for k in ("depth", "pose", "intr"))
    feat = torch.randn(S, T, G * G, D, device=dev, dtype=torch.bfloat16)
    res = pr.prune(depth, pose, intr, feat)
    n_in, n_keep = S * T * G * G, res.n_keep()
    ms = _time(lambda: pr.prune(depth, pose, intr, feat), iters=5, warmup=2)
    by = n_in * 25 + n_keep * (4 * D + 4)
    row = dict(S=S, T=T, tok=G * G, voxel=voxel, n_in=n_in, keep=round(n_keep / n_in, 3), ms=round(ms, 4),
               gbs=round(by / ms / 1e6, 1), frac=round(by / ms / 1e6 / pk["hbm_gbs"], 3), gtok=round(n_in / ms / 1e6, 4))
    if cpu:
        from oracle import voxel as ov

        f = feat[0].cpu()
        d0, p0, i0 = (torch.from_numpy(st[k][0]) for k in ("depth", "pose", "intr"))
        ov.torch_reference_idiom_prune(d0, p0, i0, f, G, voxel)
        t0 = time.perf_counter()
        ov.torch_reference_idiom_prune(d0, p0, i0, f, G, voxel)
        row["cpu_ms_1stream"] = round((time.perf_counter() - t0) * 1e3, 2)
    del feat, pr
    torch.cuda.empty_cache()
    return row


def kv_case(n_env, window_turns, ar):
    """evict one 221-row turn + a third of the memory region out of a window of `window_turns` turns"""
    P, M, turn = 60, 1568, 221
    L = P + M + window_turns * turn
    keep_mem = torch.arange(M)[torch.arange(M) % 3 != 0]
    src = torch.cat([torch.arange(P), P + keep_mem, torch.arange(P + M + turn, L)]).to(torch.int32)
    new_len = src.numel()
    first = int((src != torch.arange(new_len)).nonzero()[0])
    src_all = torch.zeros(n_env, ar.max_len, dtype=torch.int32)
    src_all[:, :new_len] = src
    src_all = src_all.to(dev)
    je = torch.arange(n_env, device=dev, dtype=torch.int32)
    ft = torch.full((n_env,), first, device=dev, dtype=torch.int32)
    nl = torch.full((n_env,), new_len, device=dev, dtype=torch.int32)
    ms = _time(lambda: ar.compact(je, ft, nl, src_all, True), iters=5, warmup=2)
    by = 2 * KV_BYTES_PER_TOKEN * n_env * (new_len - first)
    # decode attention over the resulting context, one layer
    q = torch.randn(n_env, 28, 128, device=dev, dtype=torch.bfloat16)
    out = torch.empty(n_env, 28 * 128, device=dev, dtype=torch.bfloat16)
    ones = torch.ones(n_env, device=dev, dtype=torch.int32)
    li = [0]

    def dec():
        li[0] = (li[0] + 1) % 28
        ar.attn_decode(li[0], q, 28, je, je, ones, ones * new_len, 1, new_len, 128 ** -0.5, out)
    ms_d = _time(dec, iters=28, warmup=3)
    by_d = n_env * new_len * 4 * 2 * 128 * 2
    return dict(streams=n_env, window_turns=window_turns, L=L, moved_rows=new_len - first, compact_ms=round(ms, 4),
                compact_gbs=round(by / ms / 1e6, 1), compact_frac=round(by / ms / 1e6 / pk["hbm_gbs"], 3),
                decode_ms=round(ms_d, 4), decode_gbs=round(by_d / ms_d / 1e6, 1), decode_frac=round(by_d / ms_d / 1e6 / pk["hbm_gbs"], 3))


def table(rows):
    keys = list(rows[0].keys())
    for r in rows:
        for k in r:
            if k not in keys:
                keys.append(k)
    print("| " + " | ".join(keys) + " |")
    print("|" + "---|" * len(keys))
    for r in rows:
        print("| " + " | ".join(str(r.get(k, "")) for k in keys) + " |")
    print()


print(f"# pruning / KV microbench sweep (B200, peaks: {pk})\n")
rows = []
grids = [14, 20, 24, 27]
voxels = [0.1, 0.2, 0.25, 0.5]
for S in (1, 64):
    for G in grids:
        for vx in (voxels if (G == 14 or not quick) else [0.2]):
            rows.append(prune_case(S, 32, G, vx, cpu=(S == 1 and vx == 0.2)))
print("## configs[4]: pruning, T = 32 frames\n")
table(rows)
rows = []
for S, T, G in ((1, 400, 14), (1, 400, 27), (16, 400, 14), (8, 400, 27)):
    rows.append(prune_case(S, T, G, 0.2, cpu=(S == 1)))
print("## configs[3]: long-horizon streams (T = 400 frames, one global dedup period -> global hash table path)\n")
table(rows)
rows = []
for n_env in (1, 64):
    ar = KvArena(n_env, 28, 4, 128, 4096 + 1536, 1e6, dev)
    for e in range(n_env):
        ar.buf[e].normal_(0, 1)
    for w in (4, 8, 16):
        rows.append(kv_case(n_env, w, ar))
    del ar
    torch.cuda.empty_cache()
print("## configs[4]: KV window update (evict a turn + prune a third of memory) and decode attention\n")
table(rows)

"""Per-CTA fixed cost of the tcgen05 prefill kernel: time vs number of key tiles at a fixed grid (64 x 256 queries)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.kv_arena import KvArena
from streamvln_b200.microbench import _time
dev = torch.device("cuda:0")
Hq, Hkv, hd = 28, 4, 128
ar = KvArena(64, 2, Hkv, hd, 4096, 1e6, dev)
ar.buf.normal_(0, 1)
n_seq, nq = 64, 256
qq = torch.randn(n_seq * nq, Hq, hd, device=dev, dtype=torch.bfloat16)
oo = torch.empty(n_seq * nq, Hq * hd, device=dev, dtype=torch.bfloat16)
se, qs = list(range(n_seq)), [i * nq for i in range(n_seq)]
pts = []
for kvl in (384, 512, 1024, 2048, 3072, 4000):
    ms = _time(lambda: ar.attn_prefill(1, qq, Hq, se, qs, [nq] * n_seq, [kvl] * n_seq, hd ** -0.5, oo), iters=5)
    tiles = (kvl + 127) // 128
    ctas = n_seq * Hkv * ((nq * 7 + 127) // 128)
    per_cta_us = ms * 1e3 / (ctas / 148)
    pts.append((tiles, per_cta_us))
    print(f"kv={kvl:5d} tiles={tiles:3d}: {ms:.4f} ms, {ctas} CTAs, {per_cta_us:.2f} us per CTA slot", flush=True)
(t0, y0), (t1, y1) = pts[1], pts[-2]
b = (y1 - y0) / (t1 - t0)
print(f"fit: {b:.3f} us per key tile, fixed {y0 - b * t0:.2f} us per CTA")

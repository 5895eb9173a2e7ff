"""Packed-feed stage timeline of the fused decode stack (CTA 0, middle layer): consumer wait / done and producer issue stamps."""
import os, sys
os.environ["SVLN_STACK_TRACE"] = "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ctypes as C
import numpy as np
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk

n_ctx = 3000
dev = torch.device("cuda:0")
cfg = Qwen2Config()
w = Qwen2Weights.random(cfg, dev, seed=0)
eng = Qwen2Engine(cfg, w, n_env=1, device=dev)
emb = (torch.randn(n_ctx, cfg.hidden_size) * 0.05).bfloat16().to(dev)
tok = eng.forward(emb, [Chunk(0, 0, n_ctx, n_ctx)]).argmax(-1)
runner = eng._decode_runner([0], False)
for _ in range(5):
    tok = runner.step(tok)
torch.cuda.synchronize()
L = cfg.num_hidden_layers
base = (L + 1) * 6 * 8 + 8 + 160 * 16
buf = np.zeros(base + 4096, dtype=np.int64)
eng.lib.svln_decode_stack_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
cons = buf[base:base + 2000].reshape(5, 100, 4)
prod = buf[base + 2048:base + 2048 + 800].reshape(4, 100, 2)
t0 = cons[0, 0, 0]
for ph, nm in ((0, "qkv"), (2, "o"), (3, "gate_up"), (4, "down")):
    pp = {0: 0, 2: 1, 3: 2, 4: 3}[ph]
    print(f"--- {nm}: stage | consumer wait-begin, wait-end (waited), done | producer ready-to-issue, issued (waited for empty)")
    for st in range(100):
        if cons[ph, st, 0] == 0:
            break
        if st < 24 or st % 8 == 0:
            c0, c1, c2 = cons[ph, st, :3] - t0
            p0, p1 = prod[pp, st] - t0
            print(f"  {st:3d} | {c0:8d} {c1:8d} ({c1 - c0:6d}) {c2:8d} | {p0:8d} {p1:8d} ({p1 - p0:6d})")

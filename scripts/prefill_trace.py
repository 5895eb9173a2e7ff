import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
dev = torch.device("cuda:0")
trace = torch.zeros(4 * 32 * 8, dtype=torch.int64, device=dev)
os.environ["SVLN_PF_TRACE"] = str(trace.data_ptr())
from streamvln_b200.kv_arena import KvArena
Hq, Hkv, hd = 28, 4, 128
ar = KvArena(64, 2, Hkv, hd, 4096, 1e6, dev)
ar.buf.normal_(0, 1)
n_seq, nq, kvl = 64, 256, 3072
qq = torch.randn(n_seq * nq, Hq, hd, device=dev, dtype=torch.bfloat16)
oo = torch.empty(n_seq * nq, Hq * hd, device=dev, dtype=torch.bfloat16)
se, qs = list(range(n_seq)), [i * nq for i in range(n_seq)]
for _ in range(3):
    ar.attn_prefill(1, qq, Hq, se, qs, [nq] * n_seq, [kvl] * n_seq, hd ** -0.5, oo)
torch.cuda.synchronize()
t = trace.cpu().view(4, 32, 8)
base = int(t[t > 0].min())
names = {0: "softmax half0: 0 start,1 S_FULL,2 ld+max done,3 exchange+rescale decided,4 exp done,5 P stored+arrive",
         1: "softmax half1", 2: "MMA: 0 qk start,1 K_FULL,2 qk issued,3 P_FULL,4 V_FULL,5 pv issued", 3: "TMA: 0 start,1 K_EMPTY,2 V_EMPTY"}
for role in range(4):
    print(names[role])
    for j in range(2, 14):
        print(f"  tile {j:2d}: " + " ".join(f"{int(x) - base:7d}" if x > 0 else "      -" for x in t[role, j]))

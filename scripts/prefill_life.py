"""Lifetimes of every CTA of the tcgen05 prefill kernel by SM (needs a library built with
SVLN_NVCC_EXTRA=-DSVLN_PF_INSTRUMENT): where do the ~8 us per CTA outside the key loop go?"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
dev = torch.device("cuda:0")
Hq, Hkv, hd = 28, 4, 128
n_seq, nq, kvl = 64, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_cta = n_seq * Hkv * ((nq * 7 + 127) // 128)
trace = torch.zeros(4 * 32 * 8 + n_cta * 4, dtype=torch.int64, device=dev)
os.environ["SVLN_PF_TRACE"] = str(trace.data_ptr())
from streamvln_b200.kv_arena import KvArena
ar = KvArena(64, 2, Hkv, hd, 4096, 1e6, dev)
ar.buf.normal_(0, 1)
qq = torch.randn(n_seq * nq, Hq, hd, device=dev, dtype=torch.bfloat16)
oo = torch.empty(n_seq * nq, Hq * hd, device=dev, dtype=torch.bfloat16)
se, qs = list(range(n_seq)), [i * nq for i in range(n_seq)]
for _ in range(3):
    ar.attn_prefill(1, qq, Hq, se, qs, [nq] * n_seq, [kvl] * n_seq, hd ** -0.5, oo)
torch.cuda.synchronize()
life = trace[4 * 32 * 8:].cpu().view(n_cta, 4).numpy()
import numpy as np
life = life[life[:, 1] > 0]
t0 = life[:, 1].min()
dur = life[:, 3] - life[:, 1]
pro = life[:, 2] - life[:, 1]
print(f"kv={kvl}: {len(life)} CTAs, kernel span {(life[:, 3].max() - t0) / 1e3:.1f} us; CTA lifetime mean {dur.mean() / 1e3:.2f} us "
      f"(prologue to Q stored {pro.mean() / 1e3:.2f} us)")
gaps = []
for sm in np.unique(life[:, 0]):
    rows = life[life[:, 0] == sm]
    rows = rows[np.argsort(rows[:, 1])]
    gaps += list((rows[1:, 1] - rows[:-1, 3]) / 1e3)
gaps = np.array(gaps)
print(f"gap between a CTA's exit and the next CTA's entry on the same SM: mean {gaps.mean():.2f} us, median {np.median(gaps):.2f}, "
      f"p90 {np.percentile(gaps, 90):.2f}; CTAs per SM {len(life) / len(np.unique(life[:, 0])):.1f}")

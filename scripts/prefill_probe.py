import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.kv_arena import KvArena
from streamvln_b200.microbench import _time
dev = torch.device("cuda:0")
Hq, Hkv, hd = 28, 4, 128
ar = KvArena(64, 2, Hkv, hd, 4096, 1e6, dev)
ar.buf.normal_(0, 1)
def run(n_seq, nq, kvl):
    qq = torch.randn(n_seq * nq, Hq, hd, device=dev, dtype=torch.bfloat16)
    oo = torch.empty(n_seq * nq, Hq * hd, device=dev, dtype=torch.bfloat16)
    se, qs = list(range(n_seq)), [i * nq for i in range(n_seq)]
    ms = _time(lambda: ar.attn_prefill(1, qq, Hq, se, qs, [nq] * n_seq, [kvl] * n_seq, hd ** -0.5, oo), iters=5)
    flops = n_seq * Hq * 4 * hd * nq * ((kvl - nq) + nq / 2)
    print(f"debug={os.environ.get('SVLN_PF_DEBUG','0')} n_seq={n_seq} nq={nq} kv={kvl}: {ms:.4f} ms {flops/ms/1e9:.1f} TF/s", flush=True)
run(64, 217, 3000)
run(64, 256, 3072)
run(16, 1900, 1900)

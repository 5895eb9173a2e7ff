import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200 import _lib
from streamvln_b200.microbench import _time
lib = _lib.load()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
st = _lib.cur_stream()
for N, K in ((4608, 3584), (3584, 3584), (2 * 18944, 3584), (3584, 18944), (152064, 3584)):
    # 28 distinct weight copies so every call streams from HBM, not L2
    n_copy = 28 if N * K * 2 < 300e6 else 2
    Ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(n_copy)]
    x = torch.randn(1, K, device=dev).bfloat16()
    out = torch.empty(1, N, dtype=torch.bfloat16, device=dev)
    i = [0]
    def mine():
        i[0] = (i[0] + 1) % n_copy
        lib.svln_gemv(Ws[i[0]].data_ptr(), x.data_ptr(), None, out.data_ptr(), 1, N, K, 0, st)
    def cublas():
        i[0] = (i[0] + 1) % n_copy
        torch.matmul(x, Ws[i[0]].t(), out=out)
    for name, fn in (("svln_gemv", mine), ("cuBLAS", cublas)):
        ms = _time(fn, iters=56, warmup=5)
        print(f"N={N:6d} K={K:5d} {name:10s} {ms*1e3:8.2f} us  {N*K*2/ms/1e6:8.1f} GB/s", flush=True)
    del Ws
    torch.cuda.empty_cache()

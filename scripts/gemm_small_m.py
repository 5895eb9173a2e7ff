"""cuBLAS (torch) linears at small M on the four Qwen2-7B layer shapes vs svln_gemv at M <= 4: time and weight-streaming
GB/s (run on the GPU box).  Decides whether extending svln_gemv beyond M = 4 would pay for 5-8 streams per GPU."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200 import _lib
from streamvln_b200.microbench import _time
lib = _lib.load()
dev = torch.device("cuda:0")
shapes = {"qkv": (4608, 3584), "o": (3584, 3584), "gate_up": (37888, 3584), "down": (3584, 18944), "lm_head": (152064, 3584)}
for name, (N, K) in shapes.items():
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    line = f"{name:8s} {N}x{K} ({N * K * 2 / 1e6:7.1f} MB):"
    for M in (1, 2, 3, 4, 8):
        x = torch.randn(M, K, device=dev).bfloat16()
        ms = _time(lambda: x @ W.t(), iters=10)
        line += f"  cuBLAS M={M}: {ms * 1e3:6.1f} us ({N * K * 2 / ms / 1e6:5.0f} GB/s)"
        if M <= 4:
            out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            ms2 = _time(lambda: _lib.check(lib.svln_gemv(W.data_ptr(), x.data_ptr(), None, out.data_ptr(), M, N, K, 0, _lib.cur_stream())), iters=10)
            line += f" [svln_gemv {ms2 * 1e3:6.1f} us ({N * K * 2 / ms2 / 1e6:5.0f} GB/s)]"
    print(line, flush=True)
    del W

"""Steady-turn prefill linears (M = 217) on the four Qwen2-7B layer shapes: torch's default BLAS route vs cuBLASLt, cold L2.
The floor is streaming the weights once (bytes / measured HBM peak)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.microbench import _time
dev = torch.device("cuda:0")
shapes = {"qkv(bias)": (4608, 3584, True), "o": (3584, 3584, False), "gate_up": (37888, 3584, False), "down": (3584, 18944, False)}
for M in (217, 1900):
    for lib in ("default", "cublaslt", "cublas"):
        torch.backends.cuda.preferred_blas_library(lib)
        line = f"M={M:4d} {lib:9s}:"
        tot = 0.0
        for name, (N, K, bias) in shapes.items():
            W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
            b = torch.randn(N, device=dev).bfloat16() if bias else None
            x = torch.randn(M, K, device=dev).bfloat16()
            ms = _time(lambda: torch.nn.functional.linear(x, W, b), iters=20)
            tot += ms
            line += f"  {name} {ms * 1e3:6.1f} us (floor {N * K * 2 / 6540e9 * 1e6:5.1f})"
            del W
        print(line + f"  | layer sum {tot * 1e3:6.1f} us", flush=True)

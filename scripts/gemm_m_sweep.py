"""cuBLAS time of the four Qwen2-7B layer linears against the row count M around the steady-turn chunk (217 tokens): does padding pay?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.microbench import _time
dev = torch.device("cuda:0")
shapes = {"qkv(bias)": (4608, 3584, True), "o": (3584, 3584, False), "gate_up": (37888, 3584, False), "down": (3584, 18944, False)}
Ws = {n: ((torch.randn(N, K, device=dev) * 0.02).bfloat16(), torch.randn(N, device=dev).bfloat16() if b else None) for n, (N, K, b) in shapes.items()}
for M in (128, 192, 208, 217, 224, 240, 256, 288, 320):
    line = f"M={M:4d}:"
    tot = 0.0
    for name, (N, K, _) in shapes.items():
        W, b = Ws[name]
        x = torch.randn(M, K, device=dev).bfloat16()
        ms = _time(lambda: torch.nn.functional.linear(x, W, b), iters=20)
        tot += ms
        line += f"  {name} {ms * 1e3:6.1f}"
    print(line + f"  | layer sum {tot * 1e3:6.1f} us", flush=True)

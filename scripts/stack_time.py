"""Time per fused decode step at 7B shapes (run on the GPU box): ms/step, algorithmic GB/s, fraction of the measured peak.
usage: [SVLN_STACK_*=...] python scripts/stack_time.py [n_ctx] [steps] [tag] [n_streams] [stack_max_streams]"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk

n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tag = sys.argv[3] if len(sys.argv) > 3 else ""
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda:0")
cfg = Qwen2Config()
w = Qwen2Weights.random(cfg, dev, seed=0)
eng = Qwen2Engine(cfg, w, n_env=B, device=dev)
if len(sys.argv) > 5:
    eng.STACK_MAX_STREAMS = int(sys.argv[5])     # 1: two streams take the launch-per-op path
emb = (torch.randn(n_ctx, cfg.hidden_size, generator=torch.Generator().manual_seed(0)) * 0.05).bfloat16().to(dev)
tok = torch.cat([eng.forward(emb, [Chunk(b, 0, n_ctx, n_ctx)]).argmax(-1).reshape(1) for b in range(B)])
runner = eng._decode_runner(list(range(B)), True)
toks = []
for _ in range(6):
    tok = runner.step(tok)
    toks.append(tok.tolist())
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tok = runner.step(tok)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
runner.finish()
pk = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6540.2
nbytes = eng.stack_step_bytes(n_ctx + 10) + (B - 1) * (eng.stack_step_bytes(n_ctx + 10) - eng.stack_step_bytes(0))   # weights once, every stream's K/V
env = {k: v for k, v in os.environ.items() if k.startswith("SVLN_")}
print(f"{tag:24s} B={B} {best:.4f} ms/step  {nbytes / best / 1e6:.0f} GB/s  frac {nbytes / best / 1e6 / pk:.4f}  first tokens {toks[:4]}  env {env}", flush=True)

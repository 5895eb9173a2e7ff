"""Fused decode stack: packed feed (default) against the row-major LDGSTS feed on the same weights and cache - logits of a few
decode steps must agree to fp32 summation-order noise.  usage: python scripts/stack_feed_check.py [layers]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk

dev = torch.device("cuda:0")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
worst = 0.0
for name, cfg in (("tiny", Qwen2Config.tiny(vocab_size=512, layers=3)), (f"7b-dims x{layers}", Qwen2Config(num_hidden_layers=layers))):
    w = Qwen2Weights.random(cfg, dev, seed=0)
    for n_ctx in (3, 64, 700):
        outs = []
        for feed in ("1", "0"):
            os.environ["SVLN_STACK_FEED"] = feed
            eng = Qwen2Engine(cfg, w, n_env=1, device=dev)
            eng.keep_decode_logits = True
            emb = (torch.randn(n_ctx, cfg.hidden_size, generator=torch.Generator().manual_seed(1)) * 0.05).bfloat16().to(dev)
            tok = eng.forward(emb, [Chunk(0, 0, n_ctx, n_ctx)]).argmax(-1)
            runner = eng._decode_runner([0], False)
            lg = []
            for _ in range(3):
                tok = runner.step(tok)
                lg.append(runner.last_logits().float().clone())
            runner.finish()
            outs.append(torch.stack(lg))
            del eng
        a, b = outs
        rel = float((a - b).norm() / b.norm())
        worst = max(worst, rel)
        print(f"{name:14s} ctx {n_ctx:4d}: packed vs row-major feed rel {rel:.3e}  argmax equal {bool((a.argmax(-1) == b.argmax(-1)).all())}", flush=True)
os.environ.pop("SVLN_STACK_FEED", None)
print("OK" if worst < 1e-2 else "MISMATCH", worst)

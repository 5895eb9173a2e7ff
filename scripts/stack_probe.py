"""Fused decode stack vs the launch-per-op decode path: same tokens, time per step (run on the GPU box).
usage: python scripts/stack_probe.py [tiny|7b] [n_ctx]"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk

which = sys.argv[1] if len(sys.argv) > 1 else "7b"
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
dev = torch.device("cuda:0")
cfg = (Qwen2Config.tiny(vocab_size=512, layers=3) if which == "tiny" else
       Qwen2Config(num_hidden_layers=2, vocab_size=8192) if which == "mid" else Qwen2Config())
w = Qwen2Weights.random(cfg, dev, seed=0, lm_head_std=0.3 if which == "tiny" else None)
n_env = 2
eng = Qwen2Engine(cfg, w, n_env=n_env, device=dev)
eng.keep_decode_logits = True
g = torch.Generator().manual_seed(0)
emb = (torch.randn(n_ctx, cfg.hidden_size, generator=g) * 0.05).bfloat16().to(dev)

def run(use_stack, B, use_graph, steps=6):
    eng.use_stack = use_stack
    eng._graphs.clear()
    for e in range(n_env):
        eng.arena.reset_env(e)
    chunks = [Chunk(e, e * n_ctx if False else 0, n_ctx, n_ctx) for e in range(B)]
    # same prompt on every stream (packed twice)
    h = torch.cat([emb] * B)
    chunks = [Chunk(e, e * n_ctx, n_ctx, n_ctx) for e in range(B)]
    logits = eng.forward(h, chunks)
    tok = logits.argmax(-1)
    runner = eng._decode_runner(list(range(B)), use_graph)
    toks = []
    lg = []
    for _ in range(steps):
        tok = runner.step(tok)
        toks.append(tok.cpu().tolist())
        if use_stack and min(toks[-1]) < 0:
            import ctypes as C
            st = runner.graph.state if runner.graph is not None else runner.state
            cta, cnt = C.c_int(0), C.c_int(0)
            rc = eng.lib.svln_decode_stack_status(C.byref(st.stack), C.byref(cta), C.byref(cnt))
            raise SystemExit(f"stack aborted: barrier {rc}, cta {cta.value}, count {cnt.value}")
        st = runner.graph.state if runner.graph is not None else runner.state
        if use_stack and getattr(st, "stack_logits", None) is not None:
            lg.append(st.stack_logits.clone())
    runner.finish()
    # timing
    runner = eng._decode_runner(list(range(B)), use_graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        tok = runner.step(tok)
    e1.record()
    torch.cuda.synchronize()
    runner.finish()
    return toks, e0.elapsed_time(e1) / n, lg

for B in ((2,) if os.environ.get("ONLY_B2") else (1, 2)):
    for use_graph in ((False,) if os.environ.get("ONLY_B2") else (False, True)):
        t_ref, ms_ref, _ = run(False, B, use_graph)
        t_new, ms_new, lg = run(True, B, use_graph)
        same = sum(a == b for a, b in zip(t_ref, t_new))
        print(f"B={B} graph={use_graph}: unfused {ms_ref:.3f} ms/step, stack {ms_new:.3f} ms/step, "
              f"tokens equal {same}/{len(t_ref)}  ref={t_ref[:3]} new={t_new[:3]}", flush=True)
bytes_w = sum(sum(v.numel() * 2 for v in L.values()) for L in w.layers) + w.lm_head.numel() * 2
print(f"weights per step: {bytes_w/1e9:.2f} GB -> HBM floor {bytes_w/6540.2e9*1e3:.3f} ms at 6540 GB/s")

"""Run-to-run determinism of the fused decode stack (run on the GPU box): the same state decoded R times must give
bit-identical logits.  usage: python scripts/stack_determinism.py [n_streams] [repeats] [layers]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
R = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = int(sys.argv[3]) if len(sys.argv) > 3 else 28
dev = torch.device("cuda:0")
cfg = Qwen2Config(num_hidden_layers=L)
w = Qwen2Weights.random(cfg, dev, seed=0, lm_head_std=0.05, norm_std=0.25)
eng = Qwen2Engine(cfg, w, n_env=B, device=dev)
eng.keep_decode_logits = True
import os
ctx = [int(x) for x in os.environ.get("DET_CTX", "2970,1733").split(",")][:B]
USE_GRAPH = os.environ.get("DET_GRAPH", "1") == "1"
embs = [(torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5 + i)) * 0.05).bfloat16().to(dev) for i, n in enumerate(ctx)]
first = torch.cat([eng.forward(e.clone(), [Chunk(b, 0, e.shape[0], e.shape[0])]).argmax(-1).reshape(1) for b, e in enumerate(embs)])
ref = None
bad = 0
prev = None
worst, steps_hit, streams_hit = 0.0, [], [False] * B


def decode_once():
    for b, n in enumerate(ctx):
        eng.arena.lens[b] = n
    runner = eng._decode_runner(list(range(B)), USE_GRAPH)
    tok, logs, toks = first.clone(), [], []
    for _ in range(8):
        fed = tok
        tok = runner.step(tok)
        logs.append(runner.last_logits())
        toks.append(tok.clone())
        tok = FORCED[len(toks) - 1] if FORCED is not None else tok      # teacher forcing: every run feeds the same tokens
    runner.finish()
    return torch.stack(logs), torch.stack(toks)


FORCED = None
eng.STACK_MAX_STREAMS = 0 if B > 1 else 1          # launch-per-op reference for B > 1 (B = 1: the stack itself)
eng._graphs.clear()
po = decode_once()
FORCED = [t.clone() for t in po[1]]
eng.STACK_MAX_STREAMS = 2
eng._graphs.clear()
for rep in range(R):
    got = decode_once()
    rel = [(float((got[0][s_] - po[0][s_]).abs().max() / po[0][s_].abs().max())) for s_ in range(8)]
    if os.environ.get("DET_VERBOSE"):
        print(f"rep {rep}: vs launch-per-op per step {[round(x, 4) for x in rel]}  same as previous rep: {prev is not None and torch.equal(prev[0], got[0])}", flush=True)
    if prev is not None and not torch.equal(prev[0], got[0]):
        bad += 1
        d = (prev[0] - got[0]).abs()
        first_step = int((d.amax(dim=(1, 2)) > 0).nonzero()[0])
        worst = max(worst, float(d.max()))
        steps_hit.append(first_step)
        streams_hit = [a or b for a, b in zip(streams_hit, (d.amax(dim=(0, 2)) > 0).tolist())]
    prev = got
    continue
    for b, n in enumerate(ctx):
        eng.arena.lens[b] = n
    runner = eng._decode_runner(list(range(B)), True)
    tok, logs, toks = first.clone(), [], []
    for _ in range(8):
        tok = runner.step(tok)
        logs.append(runner.last_logits())
        toks.append(tok.clone())
    runner.finish()
    got = (torch.stack(logs), torch.stack(toks))
    if ref is None:
        ref = got
    else:
        same = torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
        if not same:
            bad += 1
            d = (ref[0] - got[0]).abs()
            step = int((d.amax(dim=(1, 2)) > 0).nonzero()[0])
            print(f"rep {rep}: logits differ from step {step}, max |d| {float(d.max()):.4g}, streams {(d.amax(dim=(0, 2)) > 0).tolist()}, tokens equal {torch.equal(ref[1], got[1])}", flush=True)
print(f"B={B} layers={L} ctx={ctx} graph={USE_GRAPH} env={ {k: v for k, v in os.environ.items() if k.startswith('SVLN_')} }: {bad} of {R - 1} repeats differ from the one before, worst |d| {worst:.4g}, first differing steps {steps_hit}, streams {streams_hit}", flush=True)

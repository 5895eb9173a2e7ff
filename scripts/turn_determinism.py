"""Run-to-run determinism of a single stream's turns at 7B shapes (run on the GPU box): window-start prefill, steady
217-token prefills over the growing cache (split-KV tcgen05 attention + combine), 4 fused decode steps after each.
The logits of every prefill and every decode step must be bit-identical over R repeats."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk

R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
cfg = Qwen2Config()
w = Qwen2Weights.random(cfg, dev, seed=0, lm_head_std=0.05, norm_std=0.25)
eng = Qwen2Engine(cfg, w, n_env=1, device=dev)
eng.keep_decode_logits = True
g = torch.Generator().manual_seed(3)
chunks = [(torch.randn(n, cfg.hidden_size, generator=g) * 0.05).bfloat16().to(dev) for n in (1900, 217, 217, 217, 217)]


def episode():
    eng.arena.reset_env(0)
    out, kv = [], 0
    for e in chunks:
        kv += e.shape[0]
        lg = eng.forward(e.clone(), [Chunk(0, 0, e.shape[0], kv)])
        out.append(lg.clone())
        tok = lg.argmax(-1)
        runner = eng._decode_runner([0], True)
        for _ in range(4):
            tok = runner.step(tok)
            out.append(runner.last_logits())
        runner.finish()
        kv += 4
    return torch.cat([o.reshape(-1, cfg.vocab_size) for o in out])


ref = episode()            # first pass: graphs are captured on the way
bad = 0
for rep in range(R):
    got = episode()
    if not torch.equal(ref, got):
        bad += 1
        d = (ref - got).abs().amax(dim=1)
        print(f"rep {rep}: rows that differ {(d > 0).nonzero().flatten().tolist()[:12]}, worst |d| {float(d.max()):.4g}", flush=True)
    ref = got
print(f"single stream, {len(chunks)} turns x (prefill + 4 decode steps): {bad} of {R} repeats differ from the one before", flush=True)

"""Per-phase cycle breakdown of the fused decode stack (CTA 0), 7B shapes.  SVLN_STACK_TRACE=1 is set here."""
import os, sys
os.environ["SVLN_STACK_TRACE"] = "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ctypes as C
import numpy as np
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk

which = sys.argv[1] if len(sys.argv) > 1 else "7b"
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
dev = torch.device("cuda:0")
cfg = Qwen2Config.tiny(vocab_size=512, layers=3) if which == "tiny" else Qwen2Config()
w = Qwen2Weights.random(cfg, dev, seed=0)
eng = Qwen2Engine(cfg, w, n_env=1, device=dev)
emb = (torch.randn(n_ctx, cfg.hidden_size) * 0.05).bfloat16().to(dev)
tok = eng.forward(emb, [Chunk(0, 0, n_ctx, n_ctx)]).argmax(-1)
runner = eng._decode_runner([0], False)
for _ in range(5):
    tok = runner.step(tok)
torch.cuda.synchronize()
L = cfg.num_hidden_layers
buf = np.zeros((L + 1) * 6 * 8 + 8 + 160 * 16, dtype=np.int64)
n = eng.lib.svln_decode_stack_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
nt = (L + 1) * 6 * 8
ad = buf[nt:nt + 8].astype(np.float64)
gt = buf[nt + 8:].reshape(160, 16).astype(np.float64)
t = buf[:nt].reshape(L + 1, 6, 8).astype(np.float64)
mid = slice(2, L - 2)
# points: 0 phase entry (rowmap + first weight loads issued next), 1 after the grid barrier, 3 x ready (GEMV main
# loop starts), 4 GEMV done, 5 epilogue done.  attention (phase 1): 0 start, 5 end.
print("cycles, CTA 0, mean over layers 2..L-3")
print("phase      rowmap+issue  barrier wait  preamble      gemv  epilogue")
tot = 0
for ph, nm in ((0, "qkv"), (2, "o"), (3, "gate_up"), (4, "down")):
    iss = (t[mid, ph, 2] - t[mid, ph, 0]).mean()
    bar = (t[mid, ph, 1] - t[mid, ph, 2]).mean()
    pre = (t[mid, ph, 3] - t[mid, ph, 1]).mean()
    gemv = (t[mid, ph, 4] - t[mid, ph, 3]).mean()
    epi = (t[mid, ph, 5] - t[mid, ph, 4]).mean()
    print(f"{nm:8s} {iss:13.0f} {bar:13.0f} {pre:9.0f} {gemv:9.0f} {epi:9.0f}")
    tot += iss + bar + pre + gemv + epi
qkv_to_attn = (t[mid, 1, 0] - t[mid, 0, 5]).mean()
attn = (t[mid, 1, 5] - t[mid, 1, 0]).mean()
print(f"barrier qkv->attn {qkv_to_attn:.0f}; attention work (CTA 0) {attn:.0f}: item {(t[mid,1,1]-t[mid,1,0]).mean():.0f}, "
      f"fence+sync {(t[mid,1,2]-t[mid,1,1]).mean():.0f}, atomic {(t[mid,1,3]-t[mid,1,2]).mean():.0f}, "
      f"combine/rest {(t[mid,1,5]-t[mid,1,3]).mean():.0f}")
tot += qkv_to_attn + attn
layer = (t[3:L - 1, 0, 0] - t[2:L - 2, 0, 0]).mean()
print(f"sum {tot:.0f}; layer period {layer:.0f} cycles; whole step {(t[L,0,5]-t[0,0,0]):.0f} cycles")
pn = t[L, 5, :5]
print("norm preamble (last one traced, CTA 0 thread 0): loads+sumsq %.0f, sync %.0f, scale+store %.0f, sync %.0f" % tuple(pn[i + 1] - pn[i] for i in range(4)))
print("attention item (last layer, CTA 0): q loads issued %.0f, first tile landed %.0f, tiles done %.0f, merge staged %.0f, partial written %.0f" %
      tuple(ad[i] - ad[0] for i in range(1, 6)))

G = int((gt[:, 0] > 0).sum())
gt = gt[:G]
t0 = gt[:, 0].min()
names = {0: "qkv: barrier passed", 1: "qkv: gemv done", 2: "qkv: epilogue done", 3: "attn: barrier passed", 5: "attn: items done",
         6: "o: barrier passed", 7: "o: gemv done", 8: "o: epilogue done", 9: "gu: barrier passed", 10: "gu: gemv done",
         11: "gu: epilogue done", 12: "down: barrier passed", 13: "down: gemv done", 14: "down: epilogue done"}
print(f"all {G} CTAs, layer {L // 2}, ns since the first CTA passed the layer's first barrier: min / mean / max")
for k in sorted(names):
    v = gt[:, k] - t0
    print(f"  {names[k]:22s} {v.min():8.0f} {v.mean():8.0f} {v.max():8.0f}   (argmax CTA {int(v.argmax())})")

# ---- per-CTA streaming time of the gate/up phase (slots 9 -> 10) and its persistence across steps
def gu_durations():
    n = eng.lib.svln_decode_stack_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
    g = buf[nt + 8:].reshape(160, 16).astype(np.float64)[:G]
    return g[:, 10] - g[:, 9], g[:, 13] - g[:, 12]
runs = []
for _ in range(3):
    for _ in range(2):
        tok = runner.step(tok)
    torch.cuda.synchronize()
    runs.append(gu_durations())
gu = np.stack([r[0] for r in runs]); dn = np.stack([r[1] for r in runs])
print("gate/up main-loop ns per CTA: mean %.0f  min %.0f  max %.0f  std %.0f" % (gu.mean(), gu.min(), gu.max(), gu.std()))
print("correlation of per-CTA gate/up times between steps: %.3f %.3f; gate/up vs down: %.3f" %
      (np.corrcoef(gu[0], gu[1])[0, 1], np.corrcoef(gu[1], gu[2])[0, 1], np.corrcoef(gu.mean(0), dn.mean(0))[0, 1]))
order = np.argsort(gu.mean(0))
print("slowest CTAs (mean over 3 steps):", [(int(c), int(gu.mean(0)[c])) for c in order[-8:]])
print("fastest CTAs:", [(int(c), int(gu.mean(0)[c])) for c in order[:8]])
rel = gu.mean(0) / gu.mean()
print("relative time quantiles 5/25/50/75/95 %%: " + " ".join("%.3f" % q for q in np.quantile(rel, [0.05, 0.25, 0.5, 0.75, 0.95])))

# ---- attention phase per CTA (layer L // 2): time from its barrier pass to the end of its work, by split index
n = eng.lib.svln_decode_stack_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
g = buf[nt + 8:].reshape(160, 16).astype(np.float64)[:G]
dur = g[:, 5] - g[:, 3]
passed = g[:, 3] - g[:, 3].min()
print("attention phase per CTA: ns from barrier pass to work done (CTA = kv head * 32 + split for CTA < 128)")
for kvh in range(4):
    row = dur[kvh * 32:(kvh + 1) * 32]
    print(f"  kv head {kvh}: " + " ".join(f"{int(v / 100):3d}" for v in row) + "   (x100 ns)")
print("  CTAs 128..147 (no item): " + " ".join(f"{int(v / 100)}" for v in dur[128:]))
print("barrier pass skew per CTA (x100 ns after the first): " + " ".join(f"{int(v / 100)}" for v in passed))

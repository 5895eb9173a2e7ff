"""Where does one 7B streaming turn go?  CUDA-event + wall-clock breakdown (run on the GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
from streamvln_b200.qwen2 import Chunk

dev = torch.device("cuda:0")
m = StreamVLNForCausalLM(StreamVLNConfig(), n_env=1, device=dev, seed=0)
eng = m.engine
frame = torch.randn(1, 3, 384, 384, device=dev)
D = 3584

def ev():
    return torch.cuda.Event(enable_timing=True)

def timeit(name, fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0, e1 = ev(), ev(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:32s} gpu {e0.elapsed_time(e1) / n:8.3f} ms   wall {wall:8.3f} ms", flush=True)

timeit("vision (SigLIP+proj+pool) V=1", lambda: m.vision(frame))
ids = [11] * 20 + [-200]
feats = m.vision(frame)
timeit("splice 216 rows", lambda: m.splice(ids, feats, None))
emb = m.splice(ids, feats, None)
eng.arena.lens[0] = 1000
def prefill():
    eng.arena.lens[0] = 1000
    eng.forward(emb.clone(), [Chunk(0, 0, emb.shape[0], 1000 + emb.shape[0])])
timeit("prefill forward 216 tok @1000", prefill)
tok = torch.tensor([5], device=dev)
for use_graph in (False, True):
    runner = eng._decode_runner([0], use_graph)
    timeit(f"decode step graph={use_graph}", lambda: runner.step(tok), n=10)
x = torch.randn(216, D, device=dev, dtype=torch.bfloat16)
L = eng.w.layers[0]
timeit("qkv gemm 216", lambda: torch.addmm(L["bqkv"], x, L["wqkv"].t()))
timeit("gate_up gemm 216", lambda: x @ L["wgu"].t())
g = x @ L["wgu"].t()
a = eng._swiglu(g)
timeit("down gemm 216", lambda: a @ L["wd"].t())
timeit("lm_head 1 row", lambda: x[:1] @ eng.w.lm_head.t())
timeit("rmsnorm 216", lambda: eng._rmsnorm(x, L["ln1"]))
timeit("swiglu 216", lambda: eng._swiglu(g))
x1 = x[:1].contiguous()
timeit("qkv gemv", lambda: torch.addmm(L["bqkv"], x1, L["wqkv"].t()))
timeit("gate_up gemv", lambda: x1 @ L["wgu"].t())

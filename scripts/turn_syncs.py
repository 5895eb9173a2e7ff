"""Host synchronisations inside one slowfast model turn (run on the GPU box): every CUDA runtime call that can block the
host, in order, between `generate` entry and the D2H read of the token ids (torch.profiler / CUPTI runtime events)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from streamvln_b200 import synth

mode = sys.argv[1] if len(sys.argv) > 1 else "slowfast"
dev = torch.device("cuda:0")
model, agent = bench.build_agent(dev, mode)
frames = np.random.default_rng(0).integers(0, 255, size=(16, 480, 640, 3), dtype=np.uint8)
st = synth.make_stream(256, 14, 8, seed=0, with_feat=False)
geom = dict(depth=st["depth"], pose=st["pose"], intr=st["intr"]) if mode == "slowfast" else None
agent.reset_memory()
bench.run_turns(agent, 30, frames, None, False, geom)     # window full: every turn slides
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    bench.run_turns(agent, 1, frames, None, False, geom)
    torch.cuda.synchronize()
names = ("cudaStreamSynchronize", "cudaDeviceSynchronize", "cudaEventSynchronize", "cudaMemcpy", "cudaMemcpyAsync", "cudaStreamWaitEvent")
evs = sorted((e for e in prof.events() if any(e.name == n for n in names)), key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start if evs else 0
print(f"mode {mode}: blocking-capable runtime calls of one model turn (+3 buffered frames), us since the first:")
agg = {}
for e in evs:
    agg.setdefault(e.name, [0, 0.0, 0.0])
    a = agg[e.name]
    a[0] += 1
    a[1] += e.cpu_time_total
    a[2] = max(a[2], e.cpu_time_total)
for k, (n, tot, mx) in agg.items():
    print(f"  {k:24s} x{n:4d}  host time total {tot:9.1f} us  max {mx:9.1f} us")
long_ = [(e.name, round(e.time_range.start - t0), round(e.cpu_time_total)) for e in evs if e.cpu_time_total > 200]
print("calls that held the host for > 200 us (name, at us, for us):", long_)

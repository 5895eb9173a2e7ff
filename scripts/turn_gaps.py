"""How much of a steady turn is GPU-idle?  Wall clock per turn vs summed kernel time (torch.profiler / CUPTI)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bench

dev = torch.device("cuda:0")
model, agent = bench.build_agent(dev, "reference")
rng = np.random.default_rng(0)
frames = rng.integers(0, 255, size=(16, 480, 640, 3), dtype=np.uint8)
for n in (25, 25, 9):
    agent.reset_memory()
    bench.run_turns(agent, n, frames, None, False)
agent.reset_memory()
bench.run_turns(agent, 2, frames, None, False)
torch.cuda.synchronize()
t0 = time.perf_counter()
bench.run_turns(agent, 5, frames, None, False)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    bench.run_turns(agent, 1, frames, None, False)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in evs if hasattr(e, "device_time")) / 1e3
print(f"steady turn: wall {wall:.2f} ms; kernels+memcpy in one profiled turn {tot:.2f} ms over {len(evs)} device events")
agg = {}
for e in evs:
    agg.setdefault(e.name[:60], [0, 0.0])
    agg[e.name[:60]][0] += 1
    agg[e.name[:60]][1] += e.device_time / 1e3
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {t:8.3f} ms  x{n:4d}  {k}")
# span of device activity
starts = [e.time_range.start for e in evs]; ends = [e.time_range.end for e in evs]
print(f"device span {(max(ends) - min(starts)) / 1e3:.2f} ms")

"""cProfile of the host side of steady turns (resident frames): where does the GPU-idle time between turns go?"""
import cProfile, pstats, sys, io
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
pr = cProfile.Profile()
pr.enable()
bench.run_turns(agent, 5, frames, None, False)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])

"""Where does a PACKED turn of 64 streams go?  Wall time per packed turn vs the sum of CUDA kernel time (torch.profiler),
plus host-side stopwatches around the stages of BatchedEvaluator.step_all / generate_batch (run on the GPU box)."""
import sys, time
from pathlib import Path
from types import SimpleNamespace
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
from streamvln_b200.scheduler import BatchedEvaluator
from streamvln_b200.synth import intrinsic_matrix
from streamvln_b200.tokenizer import StubQwenTokenizer

E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
cfg = StreamVLNConfig(mode="reference")
m = StreamVLNForCausalLM(cfg, n_env=E, device=dev, seed=0)
args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, max_new_tokens=5)
ev = BatchedEvaluator(m, lambda: StubQwenTokenizer(cfg.llm.vocab_size), E, args, {"camera_intrinsic": intrinsic_matrix()})
frames = np.random.default_rng(0).integers(0, 255, size=(16, 480, 640, 3), dtype=np.uint8)
live = list(range(E))
instr = {i: "walk down the hallway and stop at the door" for i in live}

def turn():
    for k in range(4):
        s = ev.agents[0].step_id
        ev.step_all(live, {i: frames[(s + i) % 16] for i in live}, instr, k == 0)

for _ in range(12):      # windows start at turn 8: prime both shapes
    turn()
torch.cuda.synchronize()
# steady turns only (turns 12..15 of the second window)
t0 = time.perf_counter()
for _ in range(3):
    turn()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    turn()
    torch.cuda.synchronize()
ka = prof.key_averages()
cuda_total = sum(getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0)) for e in ka) / 1e3
print(f"{E} streams: wall {wall * 1e3:.1f} ms per packed steady turn; CUDA kernel time {cuda_total:.1f} ms -> GPU idle {wall * 1e3 - cuda_total:.1f} ms")
rows = sorted(ka, key=lambda e: -getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0)))[:14]
for e in rows:
    print(f"  {getattr(e, 'self_device_time_total', getattr(e, 'self_cuda_time_total', 0)) / 1e3:8.2f} ms  x{e.count:5d}  {e.key[:90]}")
cpu = sorted(ka, key=lambda e: -e.self_cpu_time_total)[:10]
print("top host-side self times:")
for e in cpu:
    print(f"  {e.self_cpu_time_total / 1e3:8.2f} ms  x{e.count:5d}  {e.key[:90]}")

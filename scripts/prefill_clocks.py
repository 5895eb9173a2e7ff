"""SM clock and power while the prefill kernel runs back to back (plain vs SVLN_PF_PERSIST=1): is the kernel running
into the power cap?"""
import os, sys, time, threading
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, pynvml
from streamvln_b200.kv_arena import KvArena
dev = torch.device("cuda:0")
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
Hq, Hkv, hd = 28, 4, 128
ar = KvArena(64, 2, Hkv, hd, 4096, 1e6, dev)
ar.buf.normal_(0, 1)
n_seq, nq, kvl = 64, 256, 3072
qq = torch.randn(n_seq * nq, Hq, hd, device=dev, dtype=torch.bfloat16)
oo = torch.empty(n_seq * nq, Hq * hd, device=dev, dtype=torch.bfloat16)
se, qs = list(range(n_seq)), [i * nq for i in range(n_seq)]
samples, stop = [], threading.Event()
def poll():
    while not stop.is_set():
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
        stop.wait(0.02)
for _ in range(3):
    ar.attn_prefill(1, qq, Hq, se, qs, [nq] * n_seq, [kvl] * n_seq, hd ** -0.5, oo)
torch.cuda.synchronize()
th = threading.Thread(target=poll); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 1500
e0.record()
for _ in range(n):
    ar.attn_prefill(1, qq, Hq, se, qs, [nq] * n_seq, [kvl] * n_seq, hd ** -0.5, oo)
e1.record()
torch.cuda.synchronize()
stop.set(); th.join()
ms = e0.elapsed_time(e1) / n
flops = n_seq * Hq * 4 * hd * nq * ((kvl - nq) + nq / 2)
s = samples[len(samples) // 4:]
clk = sorted(x[0] for x in s)[len(s) // 2]
pw = sorted(x[1] for x in s)[len(s) // 2]
cap = sum(1 for x in s if x[2] & 0x4) / len(s)
print(f"persist={os.environ.get('SVLN_PF_PERSIST', '0')}: {ms:.4f} ms/launch back to back, {flops / ms / 1e9:.0f} TFLOP/s; "
      f"median SM clock {clk} MHz, power {pw:.0f} W (limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3:.0f} W), "
      f"sw_power_cap active in {100 * cap:.0f} % of {len(s)} samples")

"""Per-kernel microbench rows only (run on the GPU box): python scripts/micro_only.py [filter-substring]"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from streamvln_b200 import microbench
dev = torch.device("cuda:0")
pk = bench.peaks()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
rows = microbench.bench_kv_attn(dev, pk)["rows"]
if "prune" in flt or not flt:
    rows = microbench.bench_prune(dev, pk)["rows"] + rows
for r in rows:
    if flt in r["kernel"] or not flt or flt == "prune":
        print(f"{r['kernel'][:58]:58s} {r['achieved']:9.2f} {r['unit']:8s} frac {r['frac']:.4f}  {r['ms']*1e3:9.1f} us")

"""Launch each hot-path kernel a few times at its batched benchmark shape (for `ncu --set full -k regex:...`).
Same shapes as streamvln_b200/microbench.py; prints the CUDA-event numbers so the plain run is self-checking."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from streamvln_b200 import microbench

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench

dev = torch.device("cuda:0")
pk = bench.peaks()
out = microbench.run_all(dev, pk)
print(json.dumps(out))

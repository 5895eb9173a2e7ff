"""svln_swiglu at the packed-prefill shape: time, GB/s and bit-exactness against torch's bf16 silu * up (GPU box)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
from streamvln_b200 import _lib
from streamvln_b200.microbench import _time
lib = _lib.load()
dev = torch.device("cuda:0")
for rows in (217, 13888, 121600):
    gu = (torch.randn(rows, 2 * 18944, device=dev) * 2).bfloat16()
    out = torch.empty(rows, 18944, dtype=torch.bfloat16, device=dev)
    fn = lambda: _lib.check(lib.svln_swiglu(gu.data_ptr(), rows, 18944, out.data_ptr(), _lib.cur_stream()))
    ms = _time(fn, iters=5)
    want = F.silu(gu[:, :18944]) * gu[:, 18944:]
    nbytes = rows * 18944 * 2 * 3
    print(f"rows {rows:7d}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s  bit-exact vs torch: {bool(torch.equal(out, want))}")

"""attn_prefill3 on the single-episode steady shape (217 new tokens over ~3000 keys) against the split-KV count."""
import sys, os, subprocess
from pathlib import Path
if len(sys.argv) > 1:
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import torch
    from streamvln_b200.kv_arena import KvArena
    from streamvln_b200.microbench import _time
    dev = torch.device("cuda:0")
    Hq, Hkv, hd = 28, 4, 128
    ar = KvArena(2, 2, Hkv, hd, 4096, 1e6, dev)
    ar.buf.normal_(0, 1)
    for nq, kvl in ((217, 2200), (217, 3000), (217, 3400)):
        qq = torch.randn(nq, Hq, hd, device=dev, dtype=torch.bfloat16)
        oo = torch.empty(nq, Hq * hd, device=dev, dtype=torch.bfloat16)
        ms = _time(lambda: ar.attn_prefill(1, qq, Hq, [0], [0], [nq], [kvl], hd ** -0.5, oo), iters=20)
        print(f"SVLN_PF_SPLITS={os.environ.get('SVLN_PF_SPLITS', 'default')} nq={nq} kv={kvl}: {ms * 1e3:.1f} us", flush=True)
else:
    for s in ("1", "2", "3", "4"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, SVLN_PF_SPLITS=s))

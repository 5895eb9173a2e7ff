"""Debug: run ONE fused decode step that stops after SVLN_STACK_STOP phases and dump the residual buffers (hb0, hb1) of the workspace.
usage: SVLN_STACK_FEED=0|1 SVLN_STACK_STOP=n python scripts/stack_stop_dump.py out.pt"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights, Chunk
dev = torch.device("cuda:0")
cfg = Qwen2Config(num_hidden_layers=2)
w = Qwen2Weights.random(cfg, dev, seed=0)
eng = Qwen2Engine(cfg, w, n_env=1, device=dev)
emb = (torch.randn(64, cfg.hidden_size, generator=torch.Generator().manual_seed(1)) * 0.05).bfloat16().to(dev)
tok = eng.forward(emb, [Chunk(0, 0, 64, 64)]).argmax(-1)
runner = eng._decode_runner([0], False)
runner.step(tok)
torch.cuda.synchronize()
st = runner.state
base = (st.stack_ws.data_ptr() + 255) // 256 * 256 - st.stack_ws.data_ptr()
D = cfg.hidden_size
hb = st.stack_ws[base:base + 4 * D].view(torch.bfloat16).float().cpu()
I = cfg.intermediate_size
off_act = base + 2 * D * 2 + 3 * D * 2      # hbuf, qrot, attn_out, o precede act (decode_stack.cu::stack_layout, hidden == n_q_head * 128)
act = st.stack_ws[off_act:off_act + 2 * I].view(torch.bfloat16).float().cpu()
torch.save({"hb0": hb[:D], "hb1": hb[D:2 * D], "act": act}, sys.argv[1])
print(os.environ.get("SVLN_STACK_FEED"), os.environ.get("SVLN_STACK_STOP"), "hb0", float(hb[:D].norm()), "hb1", float(hb[D:].norm()))

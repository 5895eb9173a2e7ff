import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from oracle import glue
from oracle.slowfast import SlowFastOracle
from streamvln_b200 import synth
from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
from test_gpu_model import _hf_from
IMG, MEM = glue.IMAGE_TOKEN_INDEX, glue.MEMORY_TOKEN_INDEX
cuda = torch.device("cuda:0")

def run(vision_graph, decode_graph, prefill_graph):
    cfg = StreamVLNConfig.tiny(vocab_size=512, mode="slowfast", layers=3)
    cfg.window_turns, cfg.memory_cap, cfg.invalid_policy, cfg.voxel_size = 3, 1568, 0, 0.25
    m = StreamVLNForCausalLM(cfg, n_env=1, device=cuda, seed=0, lm_head_std=0.3)
    m.vision.use_graph = vision_graph
    m.engine.use_graphs = decode_graph
    m.engine.use_prefill_graph = prefill_graph
    lm = _hf_from(m.weights)
    o = SlowFastOracle(lm, 14, 3, 0.25, 10.0, 0, 1568)
    st = synth.make_stream(36, 14, 8, seed=5, with_feat=False)
    frames = torch.randn(9, 3, 384, 384, generator=torch.Generator().manual_seed(3)) * 0.5
    feats = m.vision(frames.to(cuda)).float().cpu()
    out_ids = None; past = None
    for k in range(4):
        t = 4 * k
        ids = [7, 8, 9, 30, MEM, 31, 32, IMG, 10, 11] if k == 0 else out_ids + [20, 21, IMG, 22, 23]
        depth, pose, intr = st["depth"][t], st["pose"][t], st["intr"][t]
        r = m.generate(inputs=torch.tensor([ids]), images=frames[k][None, None], depths=torch.from_numpy(depth)[None, None],
                       poses=torch.from_numpy(pose)[None, None], intrinsics=torch.from_numpy(intr)[None, None], env_id=0,
                       time_ids=[[t]], frame_ids=[k], max_new_tokens=4, past_key_values=past)
        past = r.past_key_values
        seq_o = o.generate(list(ids), feats[k:k + 1], depth, pose, intr, t, 4)
        out_ids = seq_o[0].tolist()
        L = min(m.engine.arena.lens[0], o.cache.get_seq_length())
        errs = []
        for layer in range(3):
            kp = m.engine.arena.plane(0, layer, 0)[:, :L].float().cpu()
            ko = o.cache.layers[layer].keys[0][:, :L]
            e = (kp - ko).abs().amax(dim=(0, 2))
            bad = (e > 0.05 * ko.abs().max()).nonzero().flatten().tolist()
            errs.append((round(float(e.max()), 4), bad[:8], len(bad)))
        fc = (m.feat_cache[0][k].float().cpu() - feats[k]).abs().max().item()
        print(f"  turn {k}: got {r.sequences.cpu().tolist()} want {seq_o.tolist()} lens {m.engine.arena.lens[0]} {o.cache.get_seq_length()} featdiff {fc:.4f} Kerr {errs}")

for cfgs in ((False, False, False), (True, False, False), (False, True, False), (False, False, True), (True, True, True)):
    print("vision_graph, decode_graph, prefill_graph =", cfgs)
    run(*cfgs)

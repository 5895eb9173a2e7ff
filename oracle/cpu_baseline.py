"""The reference's PyTorch path timed on host CPU cores.  TEST / BENCH INFRASTRUCTURE ONLY.

Used by bench.py's `cpu_baseline` leg and by `bench.py --impl reference` (the one other place allowed to
execute oracle/).  It times the reference-semantics turn (oracle/glue.py: reference glue restated over
stock transformers Qwen2ForCausalLM + a stock transformers SigLIP tower, both random-init, bf16 like the
reference's `torch_dtype=torch.bfloat16`, streamvln_agent.py:282-289) at LLaVA-Video-7B-Qwen2 shapes.

A full 7B turn costs tens of seconds of CPU time, so one "step" is a BOUNDED SAMPLE: the turn is run on a
model with `sample_layers` of the 28 decoder layers (identical per-layer shapes, distinct weights) and the
per-layer time is scaled to 28 layers; lm_head (over every prefill row, as transformers 4.45.1 does),
embedding, SigLIP (all 26 layers) and the glue are timed in full.  The sample is stated in the JSON line.
"""
from __future__ import annotations

import os
import time
from typing import Dict

import torch

from . import glue


def build_reference_cpu(sample_layers: int = 4, full_layers: int = 28, dtype=torch.bfloat16, with_vision: bool = True):
    from transformers import Qwen2Config, Qwen2ForCausalLM

    torch.manual_seed(0)
    cfg = Qwen2Config(hidden_size=3584, intermediate_size=18944, num_hidden_layers=sample_layers, num_attention_heads=28,
                      num_key_value_heads=4, vocab_size=152064, max_position_embeddings=32768, rope_theta=1e6,
                      rms_norm_eps=1e-6, tie_word_embeddings=False)
    cfg.rope_parameters = {"rope_type": "default", "rope_theta": 1e6}
    with torch.device("meta"):
        lm = Qwen2ForCausalLM(cfg)
    lm = lm.to_empty(device="cpu").to(dtype)
    for n, p in lm.named_parameters():
        if p.dim() >= 2:
            p.data.uniform_(-0.035, 0.035)
        elif "norm" in n:
            p.data.fill_(1.0)
        else:
            p.data.uniform_(-0.035, 0.035)
    # rotary buffers are created on meta: rebuild them
    lm.model.rotary_emb.inv_freq = 1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    lm.eval()
    vis = None
    if with_vision:
        from transformers import SiglipVisionConfig, SiglipVisionModel

        vcfg = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=26, num_attention_heads=16,
                                  image_size=384, patch_size=14)
        with torch.device("meta"):
            vis = SiglipVisionModel(vcfg)
        vis = vis.to_empty(device="cpu").to(dtype)
        for n, p in vis.named_parameters():
            if p.dim() >= 2:
                p.data.uniform_(-0.035, 0.035)
            elif "norm" in n and n.endswith("weight"):
                p.data.fill_(1.0)
            else:
                p.data.zero_()
        vis.vision_model.head = torch.nn.Identity()
        vis.vision_model.embeddings.position_ids = torch.arange(729)[None]  # buffer was created on meta
        vis.eval()
    proj = (torch.empty(3584, 1152, dtype=dtype).uniform_(-0.035, 0.035), torch.zeros(3584, dtype=dtype),
            torch.empty(3584, 3584, dtype=dtype).uniform_(-0.035, 0.035), torch.zeros(3584, dtype=dtype))
    return lm, vis, proj, full_layers / sample_layers


@torch.no_grad()
def time_reference_turn(lm, vis, proj, layer_scale: float, cached_len: int = 1000, n_text: int = 20, n_gen: int = 5,
                        env_steps_per_turn: int = 4) -> Dict[str, float]:
    """One steady-state turn in reference semantics: SigLIP on the new frame, projector + 2x2 pool, splice,
    prefill of (1 + n_text + 196) new tokens over `cached_len` cached rows, lm_head on every row, then
    n_gen - 1 greedy decode tokens with torch.cat KV appends (DynamicCache)."""
    from transformers import DynamicCache

    dtype = next(lm.parameters()).dtype
    D = 3584
    o = glue.ReferenceOracle(lm, num_history=8)
    # pre-existing window context (not timed)
    cache = DynamicCache(config=lm.config)
    ctx = torch.randn(1, cached_len, D).to(dtype) * 0.02
    lm(inputs_embeds=ctx, past_key_values=cache, use_cache=True)
    o.curr_t[0], o.cache[0] = 1, {"inputs_embeds": ctx}
    t_vis = 0.0
    frame = torch.randn(1, 3, 384, 384).to(dtype)
    t0 = time.perf_counter()
    if vis is not None:
        tower = vis(frame, output_hidden_states=True).hidden_states[-1]
    else:
        tower = torch.randn(1, 729, 1152).to(dtype)
    feats, _ = glue.encode_rgbd(tower, proj, 8, [0])
    t_vis = time.perf_counter() - t0
    ids = [11] + list(range(100, 100 + n_text - 1)) + [glue.IMAGE_TOKEN_INDEX]
    # time lm_head separately so only the decoder layers are scaled
    t1 = time.perf_counter()
    seq, cache, _ = o.generate(ids, feats, [[0]], 0, cache, n_gen)
    t_llm = time.perf_counter() - t1
    n_new = len(ids) - 1 + 196
    h = torch.randn(1, n_new, D).to(dtype)
    t2 = time.perf_counter()
    lm.lm_head(h).float()
    for _ in range(n_gen - 1):
        lm.lm_head(h[:, :1]).float()
    t_head = time.perf_counter() - t2
    t_layers = max(t_llm - t_head, 0.0)
    t_full = t_vis + t_head + t_layers * layer_scale
    return dict(turn_s=t_full, vision_s=t_vis, lm_head_s=t_head, layers_sample_s=t_layers, layer_scale=layer_scale,
                steps_per_s=env_steps_per_turn / t_full)


def cores() -> int:
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()

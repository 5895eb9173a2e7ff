"""The reference's own PyTorch path, timed.  TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

Used by bench.py's `cpu_baseline` leg and `bench.py --impl reference` (host CPU cores), and by the informative
`hf_gpu` leg (the same stock code on the B200: HF Qwen2 + flash-attn 2 / sdpa + DynamicCache, which is what the reference
actually runs in production, streamvln_eval.py:523-527).  Nothing here touches streamvln_b200's kernels.

What is executed is the reference-semantics streaming turn of oracle/glue.py (`ReferenceOracle`: the reference's glue
restated over STOCK transformers modules) at LLaVA-Video-7B-Qwen2 shapes, random-init, bf16 like the reference's
`torch_dtype=torch.bfloat16` (streamvln_agent.py:282-289):
  * window-start turn: SigLIP (26 layers) on 9 views (8 re-encoded history frames + the current one, streamvln_agent.py:
    222-232), projector + 2x2 bilinear pool, splice of [system + instruction + 1568 memory + 196 frame tokens], prefill
    of ~1.83k rows from an empty cache, lm_head on EVERY row + .float() (transformers 4.45.1), 4 greedy decode tokens
    with torch.cat KV appends (DynamicCache);
  * steady turn: SigLIP on 1 frame, prefill of the fed id + ~20 text ids + 196 frame rows over the carried cache, the
    O(L) embed-cache cat (stream_video_vln.py:396-401), decode as above.
All 28 decoder layers are real and executed - no per-layer extrapolation (round 1 timed 4 layers and scaled by 7).
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional

import torch

from . import glue

D_LLM, D_VIS = 3584, 1152
N_SYS_INSTR = 60          # system block + task prompt + instruction, as token ids
N_TURN_TEXT = 20          # "<|im_start|>user\nyou can see <image>.<|im_end|>\n<|im_start|>assistant\n<|im_end|>\n" ~ 20 ids


def cores() -> int:
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def build_reference(device="cpu", layers: int = 28, dtype=torch.bfloat16, attn_impl: Optional[str] = None):
    """Stock transformers Qwen2ForCausalLM + SiglipVisionModel + the mlp2x_gelu projector weights, random-init."""
    from transformers import Qwen2Config, Qwen2ForCausalLM, SiglipVisionConfig, SiglipVisionModel

    device = torch.device(device)
    torch.manual_seed(0)
    cfg = Qwen2Config(hidden_size=D_LLM, intermediate_size=18944, num_hidden_layers=layers, num_attention_heads=28,
                      num_key_value_heads=4, vocab_size=152064, max_position_embeddings=32768, rope_theta=1e6,
                      rms_norm_eps=1e-6, tie_word_embeddings=False)
    cfg.rope_parameters = {"rope_type": "default", "rope_theta": 1e6}
    if attn_impl is not None:
        cfg._attn_implementation = attn_impl
    with torch.device("meta"):
        lm = Qwen2ForCausalLM(cfg).to(dtype)       # dtype first: to_empty then allocates bf16, not fp32 + a conversion pass
    lm = lm.to_empty(device=device)

    def init(n, p):
        if p.dim() >= 2 and p.numel() > (1 << 22) and p.device.type == "cpu":
            # bf16 uniform_ on CPU draws ~40 M values/s: a 4 M-value random block is tiled instead (a memcpy)
            blk = torch.empty(1 << 22, dtype=p.dtype).uniform_(-0.035, 0.035)
            flat = p.data.view(-1)
            for o in range(0, flat.numel(), blk.numel()):
                m = min(blk.numel(), flat.numel() - o)
                flat[o:o + m].copy_(blk[:m])
        elif p.dim() >= 2:
            p.data.uniform_(-0.035, 0.035)
        elif "norm" in n and n.endswith("weight"):
            p.data.fill_(1.0)
        else:
            p.data.uniform_(-0.035, 0.035)

    first = dict(lm.model.layers[0].named_parameters())
    for n, p in lm.named_parameters():
        if n.startswith("model.layers.") and not n.startswith("model.layers.0."):
            continue
        init(n, p)
    for li in range(1, layers):            # distinct memory per layer (copies of layer 0: a memcpy, not 7 G random draws)
        for n, p in lm.model.layers[li].named_parameters():
            p.data.copy_(first[n].data)
    lm.model.rotary_emb.inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))).to(device)
    if hasattr(lm.model.rotary_emb, "original_inv_freq"):
        lm.model.rotary_emb.original_inv_freq = lm.model.rotary_emb.inv_freq
    lm.eval()
    vcfg = SiglipVisionConfig(hidden_size=D_VIS, intermediate_size=4304, num_hidden_layers=26, num_attention_heads=16,
                              image_size=384, patch_size=14)
    vcfg._attn_implementation = "eager"    # the reference's SigLIP attention is eager matmul + softmax (siglip_encoder.py:177-239)
    with torch.device("meta"):
        vis = SiglipVisionModel(vcfg).to(dtype)
    vis = vis.to_empty(device=device)
    for n, p in vis.named_parameters():
        if p.dim() >= 2:
            p.data.uniform_(-0.035, 0.035)
        elif "norm" in n and n.endswith("weight"):
            p.data.fill_(1.0)
        else:
            p.data.zero_()
    vis.vision_model.head = torch.nn.Identity()
    vis.vision_model.embeddings.position_ids = torch.arange(729, device=device)[None]  # buffer was created on meta
    vis.eval()
    proj = (torch.empty(D_LLM, D_VIS, dtype=dtype, device=device).uniform_(-0.035, 0.035), torch.zeros(D_LLM, dtype=dtype, device=device),
            torch.empty(D_LLM, D_LLM, dtype=dtype, device=device).uniform_(-0.035, 0.035), torch.zeros(D_LLM, dtype=dtype, device=device))
    return lm, vis, proj


class ReferenceStream:
    """One episode streamed in reference semantics, turn by turn, on whatever device the modules live on."""

    def __init__(self, lm, vis, proj, n_gen: int = 5, turns_per_window: int = 8):
        self.lm, self.vis, self.proj, self.n_gen, self.tpw = lm, vis, proj, n_gen, turns_per_window
        self.device = next(lm.parameters()).device
        self.dtype = next(lm.parameters()).dtype
        self.o = glue.ReferenceOracle(lm, num_history=8)
        self.turn = 0
        self.window = 1           # window 0 of an episode has no memory; the steady state is windows >= 1
        self.past = None
        self.out_ids: Optional[List[int]] = None

    @torch.no_grad()
    def run_turn(self) -> Dict[str, float]:
        """Executes the next turn of the stream; returns its wall time (device-synchronised) and what it was."""
        dev, dt = self.device, self.dtype
        start = self.turn % self.tpw == 0
        if start:                                            # streamvln_agent.py:192-198: drop ids, cache, embeds
            self.o.reset_for_env(0)
            self.past, self.out_ids = None, None
        n_views = 9 if start else 1
        frames = torch.randn(n_views, 3, 384, 384, device=dev).to(dt)
        _sync(dev)
        t0 = time.perf_counter()
        tower = self.vis(frames, output_hidden_states=True).hidden_states[-1]                   # [V,729,1152]
        feats = glue.pool_2d(glue.mlp2x_gelu(tower, *self.proj), 27, 2, "bilinear")             # [V,196,3584]
        if start:
            ids = list(range(100, 100 + N_SYS_INSTR - 12)) + [glue.MEMORY_TOKEN_INDEX] + list(range(300, 310)) + \
                [glue.IMAGE_TOKEN_INDEX] + list(range(400, 400 + 8))
            time_ids = [[32 * self.window]]
        else:
            ids = list(self.out_ids) + list(range(500, 500 + N_TURN_TEXT - 8)) + [glue.IMAGE_TOKEN_INDEX] + list(range(600, 607))
            time_ids = [[32 * self.window + 4 * (self.turn % self.tpw)]]
            time_ids[0][0] = 32 * self.window
        seq, self.past, _ = self.o.generate(ids, feats, time_ids, 0, self.past, self.n_gen)
        self.out_ids = seq[0].tolist()                       # the D2H read of the ids (what the agent decodes)
        _sync(dev)
        dt_s = time.perf_counter() - t0
        kind = "window_start" if start else "steady"
        cached = self.past.get_seq_length()
        self.turn += 1
        if self.turn % self.tpw == 0:
            self.window += 1
        return dict(seconds=dt_s, kind=kind, cached_rows=cached)

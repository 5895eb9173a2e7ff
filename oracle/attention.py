"""CPU oracle for the per-step attention.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates what HF Qwen2 attention computes after RoPE (transformers==4.45.1, requirements.txt:140;
selected with attn_implementation="flash_attention_2" at streamvln/streamvln_eval.py:526):
causal softmax(q k^T / sqrt(d)) v with GQA (repeat_kv), fp32 softmax.  Evaluated here in float64
so it is a tolerance reference for both CUDA kernels (bf16 in/out; max rel err 1e-2 per north_star).
"""
from __future__ import annotations

import numpy as np
import torch


def attention_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_len: int, scale: float) -> torch.Tensor:
    """q [n_q, Hq, d]; k, v [Hkv, >=kv_len, d] (cache rows); query i sits at position kv_len-n_q+i.
    Returns [n_q, Hq*d] float64."""
    n_q, Hq, d = q.shape
    Hkv = k.shape[0]
    g = Hq // Hkv
    qd = q.double()
    kd = k[:, :kv_len].double().repeat_interleave(g, dim=0)  # [Hq, L, d]
    vd = v[:, :kv_len].double().repeat_interleave(g, dim=0)
    s = torch.einsum("qhd,hkd->hqk", qd, kd) * scale
    pos_q = torch.arange(kv_len - n_q, kv_len)[:, None]
    pos_k = torch.arange(kv_len)[None, :]
    s = s.masked_fill((pos_k > pos_q)[None], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("hqk,hkd->qhd", p, vd)
    return o.reshape(n_q, Hq * d)

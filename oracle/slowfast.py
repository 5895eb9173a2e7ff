"""CPU oracle for mode="slowfast".  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: the released reference has no voxel-pruned memory and no KV-retaining slide
(/root/reference/streamvln/model/stream_video_vln.py:473-479 drops the whole window;
/root/reference/streamvln/streamvln_agent.py:171,192-198), so this file is the normative CPU statement of
the semantics DESIGN.md §4 defines: stock transformers Qwen2ForCausalLM + a DynamicCache whose rows are
dropped / reordered / re-rotated exactly as the spec says, voxel keys and keep-latest from
oracle/voxel.py (C).  Positions are absolute cache rows, as in the reference (position_ids forced to None,
stream_video_vln.py:460).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import glue
from . import voxel as ov

IMAGE, MEMORY = glue.IMAGE_TOKEN_INDEX, glue.MEMORY_TOKEN_INDEX


class SlowFastOracle:
    def __init__(self, hf_model, tokens_per_side: int = 14, window_turns: int = 8, voxel_size: float = 0.2,
                 max_depth: float = 10.0, invalid_policy: int = ov.INVALID_KEEP, memory_cap: int = 1568,
                 rope_theta: float = 1e6, head_dim: int = 128, round_rotated_k_to_bf16: bool = True):
        from transformers import DynamicCache

        self.lm = hf_model.eval()
        self.G = tokens_per_side
        self.tpf = tokens_per_side ** 2
        self.window_turns, self.voxel_size, self.max_depth = window_turns, voxel_size, max_depth
        self.invalid_policy, self.memory_cap = invalid_policy, memory_cap
        self.cache = DynamicCache(config=self.lm.config)
        self.prefix_len = 0
        self.mem_keys = np.empty(0, np.int64)
        self.mem_t = np.empty(0, np.int32)
        self.turns: List[dict] = []
        self.prev_n_gen = 0
        self.inv_freq = 1.0 / (rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.float64) / head_dim))
        self.round_k = round_rotated_k_to_bf16
        self.first_logits: List[torch.Tensor] = []

    def _embed(self, ids: Sequence[int], feats: torch.Tensor) -> torch.Tensor:
        table = self.lm.get_input_embeddings().weight
        return glue.splice(ids, table, feats.to(table.dtype), None)

    @torch.no_grad()
    def generate(self, ids: List[int], feats: torch.Tensor, depth, pose, intr, t: int, max_new_tokens: int,
                 force_tokens: Optional[Sequence[int]] = None):
        """ids: the caller's input ids (first turn: full prompt with one <memory>; later: previous output ids +
        the new turn).  feats [1,196,D] pooled features of the current frame; depth [384,384], pose/intr [4,4]."""
        keys = ov.c_voxel_keys(np.asarray(depth, np.float32)[None], np.asarray(pose, np.float32)[None],
                               np.asarray(intr, np.float32)[None], self.G, self.voxel_size, self.max_depth).reshape(-1)
        L = self.cache.get_seq_length()
        if not self.turns and self.prefix_len == 0:
            self.prefix_len = ids.index(MEMORY)
            ids = ids[: self.prefix_len] + ids[self.prefix_len + 1:]
            turn_ids, tail_ids, turn_start = ids[self.prefix_len:], ids, self.prefix_len
        else:
            n_prev = self.prev_n_gen
            tail_ids, turn_ids = ids[max(n_prev - 1, 0):], ids[n_prev:]
            turn_start = L + (1 if n_prev > 0 else 0)
            if n_prev > 0:
                self.turns[-1]["n_rows"] += 1
        img_off = turn_ids.index(IMAGE)
        self.turns.append(dict(row_start=turn_start, n_rows=len(turn_ids) - 1 + self.tpf, img_off=img_off, t=t, keys=keys))
        emb = self._embed(tail_ids, feats)
        pos = torch.arange(L, L + emb.shape[0])[None]
        out = self.lm(inputs_embeds=emb[None], past_key_values=self.cache, position_ids=pos, use_cache=True)
        logits = out.logits[:, -1].float()
        self.first_logits.append(logits[0].clone())
        self.step_logits = [logits[0].clone()]
        # teacher forcing (parity tests): follow the implementation under test through near-ties
        pick = (lambda lg, i: lg.argmax(-1)) if force_tokens is None else (lambda lg, i: torch.tensor([int(force_tokens[i])]))
        tok = pick(logits, 0)
        gen = [tok]
        for i in range(1, max_new_tokens if force_tokens is None else len(force_tokens)):
            p = torch.tensor([[self.cache.get_seq_length()]])
            out = self.lm(input_ids=tok[None], past_key_values=self.cache, position_ids=p, use_cache=True)
            lg = out.logits[:, -1].float()
            self.step_logits.append(lg[0].clone())
            tok = pick(lg, i)
            gen.append(tok)
        n_gen = len(gen)
        self.prev_n_gen = n_gen
        self.turns[-1]["n_rows"] += n_gen - 1
        if len(self.turns) > self.window_turns:
            self._evict()
        return torch.stack(gen, dim=1)

    def _evict(self):
        turn = self.turns.pop(0)
        L = self.cache.get_seq_length()
        P, M = self.prefix_len, self.mem_keys.shape[0]
        assert turn["row_start"] == P + M
        cand_keys = np.concatenate([self.mem_keys, turn["keys"]])
        cand_t = np.concatenate([self.mem_t, np.full(self.tpf, turn["t"], np.int32)])
        mask = ov.c_keep_latest(cand_keys, cand_t, np.array([0, cand_keys.shape[0]], np.int64), self.invalid_policy)
        kept = np.nonzero(mask)[0]
        if kept.shape[0] > self.memory_cap:
            kept = kept[kept.shape[0] - self.memory_cap:]
        src_mem = np.where(kept < M, P + kept, turn["row_start"] + turn["img_off"] + (kept - M))
        src = np.concatenate([np.arange(P), src_mem, np.arange(turn["row_start"] + turn["n_rows"], L)]).astype(np.int64)
        new_len = src.shape[0]
        delta = torch.from_numpy(np.arange(new_len) - src)            # new position - old position (<= 0)
        ang = delta.double()[:, None] * self.inv_freq[None, :]          # [new_len, 64]
        cos, sin = ang.cos(), ang.sin()
        idx = torch.from_numpy(src)
        for layer in self.cache.layers:
            k = layer.keys[:, :, idx].double()
            lo, hi = k[..., :64], k[..., 64:]
            k_new = torch.cat([lo * cos - hi * sin, hi * cos + lo * sin], dim=-1)
            moved = (delta != 0)[None, None, :, None]
            k_new = torch.where(moved, k_new, k)
            k_new = k_new.to(layer.keys.dtype)
            if self.round_k:
                k_new = torch.where(moved, k_new.to(torch.bfloat16).to(layer.keys.dtype), k_new)
            layer.keys = k_new.contiguous()
            layer.values = layer.values[:, :, idx].contiguous()
        self.mem_keys, self.mem_t = cand_keys[kept], cand_t[kept]
        shift = L - new_len
        for t_ in self.turns:
            t_["row_start"] -= shift

    # introspection used by the parity tests
    def memory_set(self):
        return self.mem_keys.copy(), self.mem_t.copy()

"""SlowFast streaming context (mode="slowfast"): voxel-pruned slow memory spliced into a live,
turn-granular sliding KV window.  The released reference has no counterpart (it drops the whole cache
every num_frames steps, streamvln/model/stream_video_vln.py:473-479, and re-prefills
[system + instruction + 1568 memory tokens + frame]); this is the north-star semantics, specified in
DESIGN.md §4 and restated on CPU in oracle/slowfast.py.

Cache layout of one stream (row index == RoPE position):

    [ prefix P | slow memory M (pruned, grows/shrinks) | turn_0 | turn_1 | ... | turn_{w-1} ]

  * prefix  : the first prompt up to the <memory> placeholder (system + task + instruction), never moves;
  * turn    : "<user> you can see <image>. <assistant>" + 196 image rows + the generated ids;
  * after a turn, if more than `window_turns` turns are live, the OLDEST turn is evicted: its text rows are
    dropped, its 196 image rows compete with the memory rows by voxel (keep-latest-per-voxel,
    svln_keep_latest), the survivors are compacted in place right after the prefix (svln_kv_compact) and
    every moved K row is re-rotated by its position shift.  No token is ever re-encoded or re-prefilled.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch

from ._lib import check, cur_stream, ptr
from .voxel import VoxelPruneConfig, VoxelPruner

IMAGE_TOKEN_INDEX = -200
MEMORY_TOKEN_INDEX = -300


@dataclass
class Turn:
    row_start: int
    n_rows: int
    img_off: int            # offset of the first image row inside the turn
    t: int                  # frame time of the turn's image
    keys: torch.Tensor      # [tokens_per_frame] int64 voxel keys (device)


class SlowFastState:
    def __init__(self, model, env: int):
        self.m, self.env = model, env
        cfg = model.config
        G = cfg.vision.pooled_per_side
        self.tpf = G * G
        self.pruner = VoxelPruner(VoxelPruneConfig(grid=G, dim=cfg.llm.hidden_size, voxel_size=cfg.voxel_size,
                                                   max_depth=cfg.max_depth, invalid_policy=cfg.invalid_policy),
                                  1, 1, model.device)
        self.reset()

    def reset(self):
        dev = self.m.device
        self.prefix_len = 0
        self.mem_keys = torch.empty(0, dtype=torch.int64, device=dev)
        self.mem_t = torch.empty(0, dtype=torch.int32, device=dev)
        self.turns: List[Turn] = []
        self.prev_n_gen = 0
        self.evictions = 0
        self.moved_rows = 0

    # ------------------------------------------------------------------ one turn
    def prepare_turn(self, ids: List[int], r: Dict[str, Any]) -> torch.Tensor:
        m = self.m
        arena = m.engine.arena
        images = r["images"]
        assert images is not None and images.shape[1] == 1, "slowfast mode takes exactly the current frame"
        depth, pose, intr = r.get("depths"), r.get("poses"), r.get("intrinsics")
        if depth is None or pose is None or intr is None:
            raise ValueError("slowfast mode needs depths / poses / intrinsics (fp32) for the slow-memory voxel pruning")
        dev = m.device
        depth = depth.reshape(1, depth.shape[-2], depth.shape[-1]).to(dev, torch.float32).contiguous()
        pose = pose.reshape(1, 4, 4).to(dev, torch.float32).contiguous()
        intr = intr.reshape(1, 4, 4).to(dev, torch.float32).contiguous()
        keys = self.pruner.voxel_keys(depth, pose, intr).reshape(-1)
        m.engine.launches += 1
        time_ids = r.get("time_ids")
        t = int(time_ids[0][-1]) if time_ids and len(time_ids[0]) else len(self.turns)
        frame_ids = r.get("frame_ids")
        feats = m._encode_views(images[0], self.env, frame_ids)     # [1,196,D]
        if not self.turns and self.prefix_len == 0:
            # first turn of the episode: the <memory> placeholder marks the end of the immutable prefix
            if ids.count(MEMORY_TOKEN_INDEX) != 1:
                raise ValueError("the first slowfast prompt must contain exactly one <memory> placeholder")
            self.prefix_len = ids.index(MEMORY_TOKEN_INDEX)
            if any(i < 0 for i in ids[: self.prefix_len]):
                raise ValueError("<image> must come after <memory> in the first prompt")
            ids = ids[: self.prefix_len] + ids[self.prefix_len + 1:]   # memory is empty at step 0
            arena.reset_env(self.env)
            turn_ids = ids[self.prefix_len:]
            tail_ids = ids
            turn_start = self.prefix_len
        else:
            if MEMORY_TOKEN_INDEX in ids:
                raise ValueError("<memory> only appears in the first prompt of a slowfast episode")
            # ids = previous output ids + new turn; all but the last output id are already cached
            n_prev = self.prev_n_gen
            tail_ids = ids[max(n_prev - 1, 0):]
            turn_ids = ids[n_prev:]
            turn_start = arena.lens[self.env] + (1 if n_prev > 0 else 0)
            if n_prev > 0:
                self.turns[-1].n_rows += 1          # the last generated token is fed now and belongs to its turn
        if turn_ids.count(IMAGE_TOKEN_INDEX) != 1:
            raise ValueError("a slowfast turn carries exactly one <image>")
        img_off = turn_ids.index(IMAGE_TOKEN_INDEX)
        n_rows = len(turn_ids) - 1 + self.tpf
        self.turns.append(Turn(turn_start, n_rows, img_off, t, keys))
        return m.splice(tail_ids, feats, None)

    def finish_turn(self, n_gen: int) -> None:
        self.prev_n_gen = n_gen
        self.turns[-1].n_rows += n_gen - 1          # decode fed n_gen-1 generated ids into the cache
        if len(self.turns) > self.m.config.window_turns:
            self.evict_oldest()

    # ------------------------------------------------------------------ eviction
    def evict_oldest(self) -> None:
        m = self.m
        arena = m.engine.arena
        dev = m.device
        cfg = m.config
        turn = self.turns.pop(0)
        L = arena.lens[self.env]
        P, M = self.prefix_len, int(self.mem_keys.numel())
        assert turn.row_start == P + M, "the oldest turn must sit right after the memory region"
        cand_keys = torch.cat([self.mem_keys, turn.keys])
        cand_t = torch.cat([self.mem_t, torch.full((self.tpf,), turn.t, dtype=torch.int32, device=dev)])
        n = M + self.tpf
        seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
        mask = self.pruner.keep_latest(cand_keys, cand_t, seg, n, cfg.invalid_policy)
        m.engine.launches += 1
        keep = mask.bool()
        kept = torch.nonzero(keep).squeeze(1)       # ascending: old memory first (older t), then the new frame
        if kept.numel() > cfg.memory_cap:           # bounded memory: forget the oldest survivors
            kept = kept[kept.numel() - cfg.memory_cap:]
        kept_h = kept.cpu()                          # one small D2H per eviction (host bookkeeping)
        src_mem = torch.where(kept_h < M, P + kept_h, turn.row_start + turn.img_off + (kept_h - M))
        rest = torch.arange(turn.row_start + turn.n_rows, L)
        src = torch.cat([torch.arange(P), src_mem, rest]).to(torch.int32)
        new_len = int(src.numel())
        moved = (src != torch.arange(new_len, dtype=torch.int32)).nonzero()
        first = int(moved[0]) if moved.numel() else new_len
        src_row = torch.zeros(1, arena.max_len, dtype=torch.int32)
        src_row[0, :new_len] = src
        job = torch.tensor([[self.env], [first], [new_len]], dtype=torch.int32).to(dev)
        arena.compact(job[0], job[1], job[2], src_row.to(dev), rerotate=True)
        m.engine.launches += 1
        self.mem_keys = cand_keys.index_select(0, kept)
        self.mem_t = cand_t.index_select(0, kept)
        shift = L - new_len
        for t_ in self.turns:
            t_.row_start -= shift
        arena.lens[self.env] = new_len
        self.evictions += 1
        self.moved_rows += new_len - first

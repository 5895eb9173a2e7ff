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

Everything data-dependent stays on the device: the evictions of ALL streams of a packed turn are one
svln_sf_gather_candidates + one svln_keep_latest (one dedup segment per stream) + one svln_sf_evict_plan + one
svln_kv_compact (n_jobs = streams).  The new cache lengths / memory sizes travel to pinned host memory asynchronously and
are read only when the next turn needs them - after the turn's own D2H read of the token ids - so there is no host
synchronisation between `generate` entry and that read.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import torch

from ._lib import SF_JOB_WORDS, check, cur_stream, ptr
from .voxel import VoxelPruneConfig, VoxelPruner

IMAGE_TOKEN_INDEX = -200
MEMORY_TOKEN_INDEX = -300


@dataclass
class Turn:
    row_start: int
    n_rows: int
    img_off: int            # offset of the first image row inside the turn
    t: int                  # frame time of the turn's image
    slot: int               # ring slot of the turn's voxel keys (SlowFastBatch.turn_key)


class SlowFastState:
    """Host bookkeeping of one stream (rows, turns); the data lives in SlowFastBatch / the KV arena."""

    def __init__(self, batch: "SlowFastBatch", env: int):
        self.b, self.env = batch, env
        self._reset_fields()

    def _reset_fields(self):
        self.prefix_len = 0
        self.mem_n = 0
        self.turns: List[Turn] = []
        self.prev_n_gen = 0
        self.evictions = 0
        self.moved_rows = 0
        self.next_slot = 0

    def reset(self):
        self.b.settle()
        self._reset_fields()

    # -- views used by the tests / stats (settle a pending eviction first)
    @property
    def mem_keys(self) -> torch.Tensor:
        self.b.settle()
        return self.b.mem_key[self.env, : self.mem_n]

    @property
    def mem_t(self) -> torch.Tensor:
        self.b.settle()
        return self.b.mem_t[self.env, : self.mem_n]

    # ------------------------------------------------------------------ host side of one turn
    def plan_turn(self, ids: List[int], t: int):
        """Returns (tail_ids, turn): the ids whose rows this turn appends, and the new Turn record."""
        arena = self.b.m.engine.arena
        tpf = self.b.tpf
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
        n_rows = len(turn_ids) - 1 + tpf
        turn = Turn(turn_start, n_rows, img_off, t, self.next_slot)
        self.next_slot = (self.next_slot + 1) % self.b.n_slots
        self.turns.append(turn)
        return tail_ids, turn

    def finish_turn(self, n_gen: int) -> bool:
        """Returns True when the stream now holds more than window_turns turns (its oldest must be evicted)."""
        self.prev_n_gen = n_gen
        self.turns[-1].n_rows += n_gen - 1          # decode fed n_gen-1 generated ids into the cache
        return len(self.turns) > self.b.m.config.window_turns


class SlowFastBatch:
    """Device state of every stream's slow memory + the batched turn preparation and eviction."""

    def __init__(self, model, n_env: int):
        self.m = model
        cfg = model.config
        dev = model.device
        G = cfg.vision.pooled_per_side
        self.tpf = G * G
        self.cap = int(cfg.memory_cap)
        self.n_slots = int(cfg.window_turns) + 1
        self.n_env = n_env
        self.pruner = VoxelPruner(VoxelPruneConfig(grid=G, dim=cfg.llm.hidden_size, voxel_size=cfg.voxel_size,
                                                   max_depth=cfg.max_depth, invalid_policy=cfg.invalid_policy),
                                  n_env, 1, dev)
        self.mem_key = torch.zeros(n_env, self.cap, dtype=torch.int64, device=dev)
        self.mem_t = torch.zeros(n_env, self.cap, dtype=torch.int32, device=dev)
        self.turn_key = torch.zeros(n_env, self.n_slots, self.tpf, dtype=torch.int64, device=dev)
        seg = self.cap + self.tpf
        self.cand_key = torch.empty(n_env * seg, dtype=torch.int64, device=dev)
        self.cand_t = torch.empty(n_env * seg, dtype=torch.int32, device=dev)
        max_len = model.engine.arena.max_len
        self.src_row = torch.zeros(n_env, max_len, dtype=torch.int32, device=dev)
        self.plan = torch.zeros(4, n_env, dtype=torch.int32, device=dev)
        # one packed upload per eviction: [seg_off int64 x (J+1) | job words int32 x J*SF_JOB_WORDS]
        self._words = (n_env + 1) * 2 + n_env * SF_JOB_WORDS
        self.desc_host = torch.zeros(self._words, dtype=torch.int32).pin_memory()
        self.desc_dev = torch.zeros(self._words, dtype=torch.int32, device=dev)
        self.plan_host = torch.zeros(4, n_env, dtype=torch.int32).pin_memory()
        self.desc_free = None        # event: the previous upload of desc_host has been consumed
        self.geom_host: Optional[torch.Tensor] = None
        self.states = [SlowFastState(self, e) for e in range(n_env)]
        self._pending = None         # (event, [(state, L, job index)], n_jobs)

    # ------------------------------------------------------------------ turn preparation (all requests of a packed turn)
    def prepare_turns(self, requests: Sequence[Dict[str, Any]], ids_list: Sequence[List[int]]) -> List[torch.Tensor]:
        m = self.m
        dev = m.device
        R = len(requests)
        H = W = None
        plans = []
        for r, ids in zip(requests, ids_list):
            env = r.get("env_id", 0) or 0
            images = r["images"]
            if images is None or images.shape[1] != 1:
                raise ValueError("slowfast mode takes exactly the current frame")
            if r.get("depths") is None or r.get("poses") is None or r.get("intrinsics") is None:
                raise ValueError("slowfast mode needs depths / poses / intrinsics (fp32) for the slow-memory voxel pruning")
            time_ids = r.get("time_ids")
            st = self.states[env]
            t = int(time_ids[0][-1]) if time_ids and len(time_ids[0]) else len(st.turns)
            tail_ids, turn = st.plan_turn(ids, t)
            plans.append((env, tail_ids, turn))
            H, W = r["depths"].shape[-2:]
        # geometry of all streams: one pinned staging buffer, one H2D, one svln_voxel_keys launch
        per = H * W + 32
        if all(r["depths"].is_cuda for r in requests):
            depth = torch.stack([r["depths"].reshape(H, W) for r in requests]).to(dev, torch.float32).contiguous()
            pose = torch.stack([r["poses"].reshape(4, 4) for r in requests]).to(dev, torch.float32).contiguous()
            intr = torch.stack([r["intrinsics"].reshape(4, 4) for r in requests]).to(dev, torch.float32).contiguous()
        else:
            if self.geom_host is None or self.geom_host.numel() < self.n_env * per:
                self.geom_host = torch.zeros(self.n_env * per, dtype=torch.float32).pin_memory()
            gh = self.geom_host[: R * per].view(R, per)
            if getattr(self, "_geom_free", None) is not None:
                self._geom_free.synchronize()
            for i, r in enumerate(requests):
                gh[i, : H * W].copy_(r["depths"].reshape(-1).float())
                gh[i, H * W: H * W + 16].copy_(r["poses"].reshape(-1).float())
                gh[i, H * W + 16: H * W + 32].copy_(r["intrinsics"].reshape(-1).float())
            gd = gh.to(dev, non_blocking=True)
            self._geom_free = torch.cuda.Event()
            self._geom_free.record()
            depth = gd[:, : H * W].reshape(R, H, W).contiguous()
            pose = gd[:, H * W: H * W + 16].contiguous()
            intr = gd[:, H * W + 16: H * W + 32].contiguous()
        keys = self.pruner.voxel_keys(depth, pose, intr).reshape(R, self.tpf)
        m.engine.launches += 1
        envs_t = torch.tensor([p[0] for p in plans], dtype=torch.int64)
        slots_t = torch.tensor([p[2].slot for p in plans], dtype=torch.int64)
        self.turn_key[envs_t.to(dev, non_blocking=True), slots_t.to(dev, non_blocking=True)] = keys
        # vision: every stream's current frame through SigLIP in as few batches as possible
        frames = torch.cat([r["images"][0].to(dev) for r in requests])
        feats = torch.cat([m.vision(frames[b0:b0 + 16]) for b0 in range(0, R, 16)]) if R > 1 else m.vision(frames)
        return [m.splice(tail_ids, feats[i:i + 1], None) for i, (_, tail_ids, _) in enumerate(plans)]

    # ------------------------------------------------------------------ eviction (all streams that slide this turn)
    def evict(self, envs: Sequence[int]) -> None:
        if not envs:
            return
        m = self.m
        arena = m.engine.arena
        dev = m.device
        J = len(envs)
        if self.desc_free is not None:
            self.desc_free.synchronize()     # the previous upload has left the pinned buffer (long ago)
        seg_off = self.desc_host[: (self.n_env + 1) * 2].view(torch.int64)
        jobs = self.desc_host[(self.n_env + 1) * 2:].view(self.n_env, SF_JOB_WORDS)
        off = 0
        meta = []
        for j, env in enumerate(envs):
            st = self.states[env]
            turn = st.turns.pop(0)
            L = arena.lens[env]
            P, M = st.prefix_len, st.mem_n
            assert turn.row_start == P + M, "the oldest turn must sit right after the memory region"
            seg_off[j] = off
            jobs[j] = torch.tensor([env, P, M, turn.row_start, turn.img_off, turn.n_rows, L, turn.t, turn.slot, off], dtype=torch.int32)
            off += M + self.tpf
            meta.append((st, L, j))
            st.evictions += 1
        seg_off[J] = off
        self.desc_dev.copy_(self.desc_host, non_blocking=True)
        self.desc_free = torch.cuda.Event()
        self.desc_free.record()
        seg_dev = self.desc_dev[: (self.n_env + 1) * 2].view(torch.int64)[: J + 1]
        jobs_dev = self.desc_dev[(self.n_env + 1) * 2:]
        lib = m.lib
        check(lib.svln_sf_gather_candidates(ptr(jobs_dev), J, ptr(self.mem_key), ptr(self.mem_t), self.cap, ptr(self.turn_key),
                                            self.n_slots, self.tpf, ptr(self.cand_key), ptr(self.cand_t), cur_stream()))
        mask = self.pruner.keep_latest(self.cand_key[:off], self.cand_t[:off], seg_dev, self.cap + self.tpf, m.config.invalid_policy)
        plan = self.plan.view(-1)[: 4 * J].view(4, J)
        check(lib.svln_sf_evict_plan(ptr(jobs_dev), J, ptr(mask), ptr(self.cand_key), ptr(self.cand_t), self.cap, self.tpf,
                                     arena.max_len, ptr(self.mem_key), ptr(self.mem_t), ptr(self.src_row), ptr(plan), cur_stream()))
        arena.compact(jobs_dev.view(self.n_env, SF_JOB_WORDS)[:J, 0].contiguous(), plan[0], plan[1], self.src_row, rerotate=True)
        m.engine.launches += 4
        host = self.plan_host.view(-1)[: 4 * J].view(4, J)
        host.copy_(plan, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending = (ev, meta, host)
        arena.lens.pending = self.settle

    def settle(self) -> None:
        """Book the outcome of the last eviction on the host (new lengths, memory sizes, turn row shifts)."""
        if self._pending is None:
            return
        (ev, meta, host), self._pending = self._pending, None
        arena = self.m.engine.arena
        arena.lens.pending = None
        ev.synchronize()
        for st, L, j in meta:
            first, new_len, mem_n = int(host[0, j]), int(host[1, j]), int(host[2, j])
            shift = L - new_len
            for t_ in st.turns:
                t_.row_start -= shift
            st.mem_n = mem_n
            list.__setitem__(arena.lens, st.env, new_len)
            st.moved_rows += new_len - first

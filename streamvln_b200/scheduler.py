"""Concurrent-episode streaming and episode sharding across GPUs.

The reference evaluates one episode at a time per rank and shards episodes across ranks with
`episodes[rank::world]` (streamvln/streamvln_eval.py:218-219), one model replica per rank (:533), no
collective inside the loop and an `all_gather` of the per-rank metric vectors at the end (:553-565).
Here each rank additionally runs E episodes concurrently: on a model step the E turns are packed into ONE
forward (`StreamVLNForCausalLM.generate_batch`), so the weight-bound decode GEMVs become GEMMs shared by
all streams.  No collective touches the hot path; `gather_results` is the single NCCL/gloo exchange.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch


def shard_episodes(episodes: Sequence, rank: int, world: int) -> List:
    """streamvln_eval.py:218-219: rank r takes episodes[r::world]."""
    return list(episodes[rank::world])


@dataclass
class EpisodeResult:
    episode_id: int
    steps: int
    turns: int
    action_hash: int        # 63-bit hash of the action sequence (parity across runs / ranks)
    seconds: float = 0.0

    def as_row(self) -> List[float]:
        return [float(self.episode_id), float(self.steps), float(self.turns), float(self.action_hash >> 32),
                float(self.action_hash & 0xFFFFFFFF), self.seconds]

    @staticmethod
    def from_row(r) -> "EpisodeResult":
        return EpisodeResult(int(r[0]), int(r[1]), int(r[2]), (int(r[3]) << 32) | int(r[4]), float(r[5]))


def hash_actions(actions: Sequence[int]) -> int:
    h = hashlib.sha256(bytes(int(a) & 0xFF for a in actions)).digest()
    return int.from_bytes(h[:8], "little") >> 1


def gather_results(local: Sequence[EpisodeResult], device=None) -> List[EpisodeResult]:
    """The final metric gather (reference: all_gather of counts, then of padded vectors, streamvln_eval.py:553-565).
    Works with any initialised torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sorted(local, key=lambda r: r.episode_id)
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    mx = max(int(c.item()) for c in counts)
    rows = torch.zeros(max(mx, 1), 6, dtype=torch.float64, device=dev)
    if local:
        rows[: len(local)] = torch.tensor([r.as_row() for r in local], dtype=torch.float64, device=dev)
    allrows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(allrows, rows)
    out = []
    for c, rr in zip(counts, allrows):
        out += [EpisodeResult.from_row(r) for r in rr[: int(c.item())].cpu().tolist()]
    return sorted(out, key=lambda r: r.episode_id)


class BatchedEvaluator:
    """E concurrent streams on one model replica.  Each stream keeps the reference agent's state machine
    (`VLNEvaluator.prepare/finish`); model steps of all live streams are packed into one `generate_batch`."""

    def __init__(self, model, make_tokenizer: Callable[[], object], n_streams: int, args, sim_sensors_config: Dict):
        from .agent import VLNEvaluator

        self.model = model
        model.reset(n_streams)
        self.agents = [VLNEvaluator(sim_sensors_config, model=model, tokenizer=make_tokenizer(), args=args)
                       for _ in range(n_streams)]
        self.n = n_streams
        self._pool = None
        self._pin, self._pin_free = None, None

    def reset_stream(self, i: int):
        self.agents[i]._init_state()
        self.model.reset_episode(i)

    def step_all(self, live: Sequence[int], frames: Dict[int, np.ndarray], instructions: Dict[int, str], run_model: bool,
                 geom: Optional[Dict[int, dict]] = None, preprocessed: Optional[Dict[int, torch.Tensor]] = None):
        """One env step for every live stream.  Returns {stream: action list} on a model step, else {}."""
        reqs, who = [], []
        if run_model and frames and preprocessed is None and self.agents[0].gpu_preprocess:
            # host uint8 frames of ALL live streams: one pinned staging buffer, one H2D, one launch pair of the GPU image
            # processor (per stream it was a copy + two launches each; bit-exact either way)
            arrs = [np.asarray(frames[i])[..., :3] for i in live]
            if len({a.shape for a in arrs}) == 1:
                n, shape = len(arrs), arrs[0].shape
                if self._pin is None or self._pin.shape[0] < n or tuple(self._pin.shape[1:]) != shape:
                    self._pin = torch.empty((max(n, self.n),) + shape, dtype=torch.uint8).pin_memory()
                    self._pin_free = None
                if self._pin_free is not None:
                    self._pin_free.synchronize()          # the previous step's H2D has left the staging buffer
                for k, a in enumerate(arrs):
                    self._pin[k].copy_(torch.from_numpy(np.ascontiguousarray(a)))
                dev = self._pin[:n].to(self.model.device, non_blocking=True)
                self._pin_free = torch.cuda.Event()
                self._pin_free.record()
                out = self.agents[0].gpu_proc(dev)
                self.model.engine.launches += 2
                preprocessed = {i: out[k] for k, i in enumerate(live)}
        if run_model and frames and preprocessed is None and len(live) > 1 and not self.agents[0].gpu_preprocess:
            # host uint8 frames: run the PIL image processor of all streams on a thread pool (PIL drops the GIL)
            from concurrent.futures import ThreadPoolExecutor
            from PIL import Image
            import os

            if self._pool is None:
                self._pool = ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4))
            proc = self.agents[0].image_processor
            futs = {i: self._pool.submit(lambda f: proc.preprocess(images=Image.fromarray(f).convert("RGB"),
                                                                   return_tensors="pt")["pixel_values"][0], frames[i]) for i in live}
            preprocessed = {i: f.result() for i, f in futs.items()}
        for i in live:
            g = geom[i] if geom else {}
            req = self.agents[i].prepare(i, frames.get(i) if frames else None, instructions[i], run_model, g.get("depth"),
                                         g.get("pose"), g.get("intrinsic"), None if preprocessed is None else preprocessed.get(i))
            if req is not None:
                reqs.append(req)
                who.append(i)
        out = {}
        if reqs:
            results = self.model.generate_batch(reqs)
            for i, r in zip(who, results):
                out[i] = self.agents[i].finish(r)[0]
        for i in live:
            self.agents[i].step_id += 1
        return out


def run_episodes(ev: BatchedEvaluator, episodes: Sequence[dict], frame_source: Callable[[int, int], np.ndarray],
                 num_future_steps: int = 4, resident_source: Optional[Callable[[int, int], torch.Tensor]] = None,
                 geom_source: Optional[Callable[[int, int], dict]] = None) -> List[EpisodeResult]:
    """Stream `episodes` (dicts with id, length, instruction) through the E slots of `ev`, refilling a slot as soon
    as its episode ends (the reference runs them one after another per rank, streamvln_eval.py:218-350).
    frame_source(episode_id, step) -> uint8 HxWx3 host frame; resident_source(episode_id, step) -> preprocessed
    [3,384,384] device tensor (frames already in HBM; frame_source is then unused); geom_source(episode_id, step) ->
    dict(depth, pose, intrinsic) for the slowfast mode."""
    queue = list(episodes)
    slots: Dict[int, dict] = {}
    done: List[EpisodeResult] = []
    for i in range(ev.n):
        if queue:
            slots[i] = dict(ep=queue.pop(0), step=0, actions=[], turns=0)
            ev.reset_stream(i)
    while slots:
        # streams advance in lock-step groups by phase so that model steps coincide
        by_phase: Dict[bool, List[int]] = {True: [], False: []}
        for i, s in slots.items():
            by_phase[s["step"] % num_future_steps == 0].append(i)
        for run_model in (True, False):
            live = by_phase[run_model]
            if not live:
                continue
            geom = None if geom_source is None else {i: geom_source(slots[i]["ep"]["id"], slots[i]["step"]) for i in live}
            if resident_source is not None:
                pre = {i: resident_source(slots[i]["ep"]["id"], slots[i]["step"]) for i in live} if run_model else None
                acts = ev.step_all(live, None, {i: slots[i]["ep"]["instruction"] for i in live}, run_model, geom, pre)
            else:
                frames = {i: frame_source(slots[i]["ep"]["id"], slots[i]["step"]) for i in live}
                acts = ev.step_all(live, frames, {i: slots[i]["ep"]["instruction"] for i in live}, run_model, geom)
            for i in live:
                s = slots[i]
                if i in acts:
                    s["actions"] += acts[i]
                    s["turns"] += 1
                s["step"] += 1
        for i in list(slots):
            s = slots[i]
            if s["step"] >= s["ep"]["length"]:
                done.append(EpisodeResult(s["ep"]["id"], s["step"], s["turns"], hash_actions(s["actions"])))
                if queue:
                    slots[i] = dict(ep=queue.pop(0), step=0, actions=[], turns=0)
                    ev.reset_stream(i)
                else:
                    del slots[i]
    return done

#!/usr/bin/env python
"""bench.py — streaming nav steps/s of the StreamVLN SlowFast streaming-context hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|hf_gpu] [--blocks a,b,...]

Headline workload (BASELINE.json configs[1]): StreamVLN LLaVA-Video-7B-Qwen2 single-episode streaming, num_history=8,
num_frames=32, num_future_steps=4, bf16, random-init weights (Qwen2-7B + SigLIP-so400m shapes), synthetic 480x640 RGB
frames, n_gen fixed to 5 (random weights never emit EOS).  One bench STEP = one 32-frame WINDOW of the stream = 8 model
turns (1 window-start turn that re-prefills ~1.9k tokens + 7 steady turns of ~217 tokens; 1 model call + 3 buffered
frames per turn, the cadence of http_realworld_server.py:94-102) = 32 env steps; `value` = env steps per second.  A step
is a whole window so that every timed step carries the workload's own 1-in-8 mix of expensive turns, whatever K is.
N>1: every rank streams its own episode (episodes shard with no collective on the hot path) - weak scaling.

One JSON line (task contract): value (frames resident in HBM), e2e (through VLNEvaluator.step with host uint8 frames:
pinned H2D + GPU image processor + turn + D2H of the ids), roofline (the dominant kernel - the fused decode stack - timed
live with CUDA events), roofline_prefill, clocks, gpu_launches, plus these blocks (all measured in the same process):
  slowfast  - the north_star semantics (voxel-pruned memory + sliding KV window) on the same single-episode stream;
  batched   - BASELINE.json configs[2]: 64 concurrent R2R-shaped episodes (lengths U{40..120}, seed 0) sharded
              episodes[rank::world] (streamvln_eval.py:218-219), strong scaling, results gathered over NCCL
              (scheduler.gather_results = streamvln_eval.py:553-565) with per-episode action hashes;
  kernels   - per-kernel rooflines at the batched shapes, with flash-attn 2 timed next to the tcgen05 prefill kernel;
  cpu_baseline - the reference-semantics turn on the host cores (oracle port, all 28 layers executed);
  hf_gpu    - informative: stock transformers Qwen2 + flash-attn 2 / sdpa + DynamicCache on the same B200.
`--impl reference` is the driver's reference arm: the same CPU port, same metric / step definition.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "streaming_nav_steps_per_sec"
UNIT = "env_steps/s"
ENV_STEPS_PER_TURN = 4
TURNS_PER_WINDOW = 8          # num_frames / num_future_steps = 32 / 4
ENV_STEPS_PER_STEP = ENV_STEPS_PER_TURN * TURNS_PER_WINDOW
N_GEN = 5
PRIME_TURNS = 25
INSTRUCTION = ("walk down the hallway past the kitchen turn left at the sofa go through the second door on the right "
               "and wait by the window next to the wooden table in the corner of the bedroom then stop near the bed "
               "facing the large mirror on the wall")  # 48 words ~ the 40 fixed token ids of SURVEY §8d
STEP_DEF = ("1 bench step = 1 window of the stream = 8 model turns (1 window-start re-prefill + 7 steady) = 32 env steps; "
            "value counts env steps")


def shared_config(mode: str, world: int):
    """`config` of the JSON line - identical for `--impl ours` and `--impl reference` (the driver compares the two arms)."""
    return {"workload": "configs[1]: StreamVLN LLaVA-Video-7B-Qwen2 single-episode streaming step per GPU, num_history=8, "
                        "num_frames=32, num_future_steps=4, n_gen=5, bf16, random-init weights",
            "mode": mode, "step_definition": STEP_DEF, "env_steps_per_step": ENV_STEPS_PER_STEP, "turns_per_step": TURNS_PER_WINDOW,
            "episodes_per_gpu": 1, "parallelism": f"episode-sharded x{world}"}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def stack_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one decode_stack_kernel launch, from the newest `ncu --set full`
    summary committed under profiles/ (profiles/decode_stack_traffic.json names the capture); None if absent."""
    p = ROOT / "profiles" / "decode_stack_traffic.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text())
    return int(d["dram_bytes_per_launch"]), d.get("source")


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a
    thread of this process (nvidia_ml_py): spawning `nvidia-smi -lms` at the start of the timed region and its
    periodic driver queries cost the first timed turns 3-12 ms each (measured); nvidia-smi is the fallback.  The timed
    region is several seconds long, so a sample every 0.25 s is plenty and keeps the driver lock out of the way."""

    _BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.25):
        self.idx, self.period = gpu_index, period_s
        self.sm, self.mx, self.reasons = [], [], set()
        self.thread = self.proc = self.nvml = self.handle = None
        self.stop_flag = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps indices: go through the PCI bus id of the torch device
            import torch

            props = torch.cuda.get_device_properties(gpu_index)
            bus = props.pci_bus_id if hasattr(props, "pci_bus_id") else None
            self.handle = (pynvml.nvmlDeviceGetHandleByIndex(gpu_index) if bus is None else None)
            if self.handle is None:
                n = pynvml.nvmlDeviceGetCount()
                for i in range(n):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        self.handle = h
                        break
                if self.handle is None:
                    self.handle = pynvml.nvmlDeviceGetHandleByIndex(min(gpu_index, n - 1))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.reasons |= {name for bit, name in self._BITS.items() if mask & bit}
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        cmd = ["nvidia-smi", f"--id={self.idx}",
               "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
               "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
               "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"]
        try:
            self.rows = []
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


def window_aligned(warmup: int) -> int:
    """(kept for the CPU tests) turns rounded up to whole windows."""
    return (max(warmup, 1) + TURNS_PER_WINDOW - 1) // TURNS_PER_WINDOW * TURNS_PER_WINDOW


# ------------------------------------------------------------------------------------------ ours
def build_agent(device, mode: str, share=None):
    from streamvln_b200.agent import VLNEvaluator
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
    from streamvln_b200.synth import intrinsic_matrix
    from streamvln_b200.tokenizer import StubQwenTokenizer

    cfg = StreamVLNConfig(mode=mode)  # Qwen2-7B + SigLIP-so400m shapes
    if share is None:
        model = StreamVLNForCausalLM(cfg, n_env=1, device=device, seed=0)
    else:   # a second semantic mode over the SAME weights (no second 15 GB copy)
        model = StreamVLNForCausalLM(cfg, weights=share.weights, vision=share.vision, n_env=1, device=device)
    args = SimpleNamespace(num_frames=32, num_future_steps=ENV_STEPS_PER_TURN, num_history=8, max_new_tokens=N_GEN)
    agent = VLNEvaluator({"camera_intrinsic": intrinsic_matrix()}, model=model, tokenizer=StubQwenTokenizer(cfg.llm.vocab_size), args=args)
    return model, agent


LAST_ACTIONS = []      # actions of the most recent run_turns call (hashed by main_ours: ranks must agree)


def run_turns(agent, n_turns: int, frames_host, frames_dev, resident: bool, geom=None):
    """n_turns model turns = 4*n_turns env steps through VLNEvaluator.step (the call a user makes)."""
    h2d = d2h = 0
    del LAST_ACTIONS[:]
    verbose = os.environ.get("SVLN_BENCH_VERBOSE") == "1"
    marks = []
    for ti in range(n_turns):
        if verbose and ti % TURNS_PER_WINDOW == 0:
            import torch
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((agent.step_id, ev, time.perf_counter()))
        for k in range(ENV_STEPS_PER_TURN):
            s = agent.step_id
            i = s % len(frames_host)
            kw = {}
            if geom is not None:
                kw = dict(depth=geom["depth"][s % len(geom["depth"])], pose=geom["pose"][s % len(geom["pose"])],
                          intrinsic=geom["intr"][0])
            if resident:
                acts = agent.step(0, None, INSTRUCTION, run_model=(k == 0), preprocessed=frames_dev[i] if k == 0 else None, **kw)[0]
            else:
                acts = agent.step(0, frames_host[i], INSTRUCTION, run_model=(k == 0), **kw)[0]
            if k == 0:
                LAST_ACTIONS.extend(acts or [])
                h2d += getattr(agent, "h2d_bytes", 0)
                d2h += getattr(agent, "d2h_bytes", 0)
                if geom is not None and not resident:
                    h2d += geom["depth"][0].nbytes + 128
            agent.step_id += 1
    if verbose and len(marks) > 1:
        import torch
        torch.cuda.synchronize()
        print(("resident" if resident else "e2e") + " per-window ms (gpu/host): " +
              " ".join(f"{sid}:{a.elapsed_time(b):.0f}/{1e3 * (tb - ta):.0f}" for (sid, a, ta), (_, b, tb) in zip(marks[:-1], marks[1:])),
              file=sys.stderr, flush=True)
    return h2d, d2h


def timed(agent, model, n_turns, frames_host, frames_dev, resident, dist, device, geom=None):
    import torch

    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = model.engine.launches
    rng_on = os.environ.get("SVLN_NCU_RANGE") == ("resident" if resident else "e2e")
    if rng_on:  # `ncu --profile-from-start off`: profile exactly this timed region
        torch.cuda.profiler.start()
    e0.record()
    h2d, d2h = run_turns(agent, n_turns, frames_host, frames_dev, resident, geom)
    e1.record()
    torch.cuda.synchronize(device)
    if rng_on:
        torch.cuda.profiler.stop()
        os.environ["SVLN_NCU_RANGE"] = "done"
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, model.engine.launches - l0, h2d, d2h


def prime(agent, frames_host, frames_dev, geom=None):
    """Initialisation, like building the model: throw-away episodes so that every CUDA graph the steady state replays
    (vision, steady turn, window-start turns, decode) is captured AND replayed once before any warm-up or timed step.
    A graph is captured on the second sighting of a shape and its first replay pays the upload (~12-16 ms).  The first
    window start of an episode (step 32) and the later ones (64, 96, ...) are different shapes, hence two episodes of
    three windows and a third of one."""
    for n in (PRIME_TURNS, PRIME_TURNS, 9):
        agent.reset_memory()
        run_turns(agent, n, frames_host, frames_dev, True, geom)


def slowfast_block(model, a, frames_host, frames_dev, dist, device, rank, world):
    """The north_star semantics on the same single-episode stream: voxel keys -> keep-latest dedup -> device-side
    eviction plan -> in-place KV compaction with K re-rotation every turn once the window is full (no re-prefill)."""
    import torch
    from streamvln_b200 import synth

    m_sf, agent = build_agent(device, "slowfast", share=model)
    st = synth.make_stream(512, 14, 8, seed=rank, with_feat=False)
    geom = dict(depth=st["depth"], pose=st["pose"], intr=st["intr"])
    K, W = a.steps * TURNS_PER_WINDOW, max(a.warmup, 3) * TURNS_PER_WINDOW
    K = min(K, 64)                       # 64 turns = 256 env steps are plenty (every slowfast turn costs the same)
    agent.reset_memory()
    run_turns(agent, PRIME_TURNS, frames_host, frames_dev, True, geom)
    agent.reset_memory()
    run_turns(agent, max(W, 16), frames_host, frames_dev, True, geom)   # the window (8 turns) is full: every turn slides
    sf = m_sf.sf[0]
    ev0, mv0 = sf.evictions, sf.moved_rows
    ms, launches, _, _ = timed(agent, m_sf, K, frames_host, frames_dev, True, dist, device, geom)
    _ = m_sf.engine.arena.lens[0]        # settle the last eviction's bookkeeping (outside the timed region)
    ev1, mv1 = sf.evictions, sf.moved_rows
    ms_e2e, _, h2d, d2h = timed(agent, m_sf, K, frames_host, frames_dev, False, dist, device, geom)
    total = K * ENV_STEPS_PER_TURN * world
    out = {"mode": "slowfast", "parity": "self-specified (no reference counterpart; DESIGN.md section 4, oracle/slowfast.py)",
           "value": round(total / (ms / 1e3), 3), "unit": UNIT, "turns": K, "ms_per_turn": round(ms / K, 3),
           "e2e": {"value": round(total / (ms_e2e / 1e3), 3), "unit": UNIT, "ms_per_turn": round(ms_e2e / K, 3),
                   "h2d_bytes_per_turn": int(h2d / K), "d2h_bytes_per_turn": int(d2h / K)},
           "evictions_per_turn": round((ev1 - ev0) / K, 3), "moved_rows_per_eviction": round((mv1 - mv0) / max(ev1 - ev0, 1), 1),
           "memory_tokens": int(sf.mem_n), "cache_rows": int(m_sf.engine.arena.lens[0]), "gpu_launches": int(launches),
           "note": "window_turns=8, memory_cap=1568, voxel 0.2 m; every turn appends ~221 rows and evicts the oldest turn: "
                   "svln_voxel_keys + svln_sf_gather_candidates + svln_keep_latest + svln_sf_evict_plan + svln_kv_compact, "
                   "no host synchronisation before the turn's D2H read of the ids"}
    del agent, m_sf
    torch.cuda.empty_cache()
    return out


def batched_block(model, a, dist, device, rank, world):
    """BASELINE.json configs[2]: E concurrent synthetic R2R-shaped episodes (lengths U{40..120}, seed 0), sharded
    episodes[rank::world]; every model step packs the turns of all local streams into one forward; the job ends when the
    last episode ends; results are gathered with scheduler.gather_results over NCCL."""
    import numpy as np
    import torch
    from PIL import Image
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
    from streamvln_b200.scheduler import BatchedEvaluator, gather_results, run_episodes, shard_episodes
    from streamvln_b200.synth import intrinsic_matrix
    from streamvln_b200.tokenizer import StubQwenTokenizer

    E_total = a.episodes
    rng = np.random.default_rng(0)
    lengths = rng.integers(40, 121, size=E_total)
    episodes = [dict(id=i, length=int(lengths[i]), instruction=INSTRUCTION + f" then look for object {i}") for i in range(E_total)]
    mine = shard_episodes(episodes, rank, world)
    E = len(mine)
    cfg = StreamVLNConfig(mode="reference")
    m_b = StreamVLNForCausalLM(cfg, weights=model.weights, vision=model.vision, n_env=E, device=device)
    args = SimpleNamespace(num_frames=32, num_future_steps=ENV_STEPS_PER_TURN, num_history=8, max_new_tokens=N_GEN)
    ev = BatchedEvaluator(m_b, lambda: StubQwenTokenizer(cfg.llm.vocab_size), E, args, {"camera_intrinsic": intrinsic_matrix()})
    pool_host = np.random.default_rng(99).integers(0, 255, size=(16, 480, 640, 3), dtype=np.uint8)   # same on every rank
    proc = ev.agents[0].image_processor
    pool_dev = torch.stack([proc.preprocess(images=Image.fromarray(f), return_tensors="pt")["pixel_values"][0]
                            for f in pool_host]).to(device)
    pick = lambda ep, step: (ep * 7 + step) % 16     # a frame depends on (episode, step) only: the result cannot depend on the sharding
    host_src = lambda ep, step: pool_host[pick(ep, step)]
    dev_src = lambda ep, step: pool_dev[pick(ep, step)]
    # priming = the job itself, TWICE, untimed: as episodes end the pack shrinks; every new pack size captures its own
    # decode graph at first sight and its prefill graph at the second (measured: 13.0 s for the second pass, 8.4 s for the
    # third); the timed passes only replay
    for _ in range(2):
        run_episodes(ev, mine, host_src, resident_source=dev_src)

    def run(resident: bool):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = m_b.engine.launches
        e0.record()
        res = run_episodes(ev, mine, host_src, resident_source=dev_src if resident else None)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
        if dist is not None:
            dist.barrier()
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, m_b.engine.launches - l0, res

    ms, launches, res = run(True)
    ms_e2e, _, res_e2e = run(False)
    h2d_local = sum(r.turns for r in res_e2e) * (480 * 640 * 3)
    d2h_local = sum(r.turns for r in res_e2e) * N_GEN * 8
    allres = gather_results(res, device=device)          # the single collective: all_gather of counts + padded rows (NCCL)
    alle2e = gather_results(res_e2e, device=device)
    tot = torch.tensor([h2d_local, d2h_local], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(tot)
    steps_total = sum(r.steps for r in allres)
    digest = hashlib.sha256(",".join(f"{r.episode_id}:{r.action_hash:016x}" for r in allres).encode()).hexdigest()[:16]
    out = {"config": f"configs[2]: {E_total} concurrent synthetic R2R-shaped episodes (lengths U{{40..120}}, seed 0), sharded "
                     f"episodes[rank::{world}] ({E} per GPU), LLaVA-Video-7B-Qwen2 shapes, n_gen=5, bf16, reference mode",
           "scaling": "strong", "episodes": len(allres), "episodes_per_gpu": E, "env_steps_total": int(steps_total),
           "value": round(steps_total / (ms / 1e3), 2), "unit": UNIT, "seconds": round(ms / 1e3, 3),
           "e2e": {"value": round(steps_total / (ms_e2e / 1e3), 2), "unit": UNIT, "seconds": round(ms_e2e / 1e3, 3),
                   "h2d_bytes_total": int(tot[0].item()), "d2h_bytes_total": int(tot[1].item())},
           "gpu_launches": int(launches), "gather": "scheduler.gather_results over " + ("nccl" if dist is not None else "1 rank (no collective)"),
           "action_hash_digest": digest, "e2e_hashes_equal_resident": [r.action_hash for r in alle2e] == [r.action_hash for r in allres],
           "action_hashes": [f"{r.action_hash:016x}" for r in allres]}
    del ev, m_b
    torch.cuda.empty_cache()
    worlds = [int(x) for x in str(getattr(a, "emulate_world", "") or "").split(",") if x.strip() and int(x) > 1]
    if worlds and world == 1:
        # Multi-GPU parity check that needs one GPU: run every shard episodes[r::W] of a W-rank job here, one after the
        # other, with the SAME pack compositions a W-rank run has (same priming, so the same graphs replay).  A pack of
        # 64 and a pack of 64 / W give cuBLAS different GEMM shapes, so the N = 1 hashes above are NOT comparable with an
        # N = W run's; these are: `digest` must equal the `action_hash_digest` the W-rank run prints.
        out["emulated_world"] = {}
        for W in worlds:
            emu = []
            for r in range(W):
                shard = shard_episodes(episodes, r, W)
                m_s = StreamVLNForCausalLM(cfg, weights=model.weights, vision=model.vision, n_env=len(shard), device=device)
                ev_s = BatchedEvaluator(m_s, lambda: StubQwenTokenizer(cfg.llm.vocab_size), len(shard), args, {"camera_intrinsic": intrinsic_matrix()})
                for _ in range(2):
                    run_episodes(ev_s, shard, host_src, resident_source=dev_src)
                emu += run_episodes(ev_s, shard, host_src, resident_source=dev_src)
                del ev_s, m_s
                torch.cuda.empty_cache()
            emu.sort(key=lambda r: r.episode_id)
            out["emulated_world"][str(W)] = {
                "digest": hashlib.sha256(",".join(f"{r.episode_id}:{r.action_hash:016x}" for r in emu).encode()).hexdigest()[:16],
                "action_hashes": [f"{r.action_hash:016x}" for r in emu],
                "note": f"the {W} shards episodes[r::{W}] run one after another on this GPU with a {W}-rank job's pack compositions"}
    return out


def fa2_rows(device, pk):
    """flash-attn 2 (the reference's attention, streamvln_eval.py:526) at the three prefill shapes, one layer, next to
    the tcgen05 kernel's rows in `kernels` - a library yardstick, not a product path."""
    import torch
    from streamvln_b200.microbench import _row, _time

    rows = []
    try:
        from flash_attn import flash_attn_varlen_func
    except Exception as e:
        return [{"kernel": "flash_attn_varlen_func", "unavailable": repr(e)}]
    Hq, Hkv, hd = 28, 4, 128
    for n_seq, nq, kvl, name in ((64, 217, 3000, "steady turn x64"), (16, 1900, 1900, "window start x16"),
                                 (1, 217, 3000, "steady turn x1"), (1, 1900, 1900, "window start x1")):
        try:
            q = torch.randn(n_seq * nq, Hq, hd, device=device, dtype=torch.bfloat16)
            k = torch.randn(n_seq * kvl, Hkv, hd, device=device, dtype=torch.bfloat16)
            v = torch.randn(n_seq * kvl, Hkv, hd, device=device, dtype=torch.bfloat16)
            cq = torch.arange(0, (n_seq + 1) * nq, nq, device=device, dtype=torch.int32)
            ck = torch.arange(0, (n_seq + 1) * kvl, kvl, device=device, dtype=torch.int32)
            ms = _time(lambda: flash_attn_varlen_func(q, k, v, cq, ck, nq, kvl, softmax_scale=hd ** -0.5, causal=True), iters=5)
            flops = n_seq * Hq * 4 * hd * nq * ((kvl - nq) + nq / 2)
            rows.append(_row("flash_attn_varlen_func (FA2 2.8.3, library) " + name, "tensor", flops, ms, pk["bf16_tflops"], "TFLOP/s",
                             f"{n_seq} streams x n_q={nq} over kv_len={kvl}, one layer, causal (bottom-right aligned)"))
            del q, k, v
        except Exception as e:
            rows.append({"kernel": "flash_attn_varlen_func " + name, "unavailable": repr(e)[:200]})
    return rows


def hf_gpu_leg(device, windows: int = 2):
    """Informative: the reference's production path - stock transformers Qwen2 (28 layers) + SigLIP + DynamicCache with
    flash-attn 2 (streamvln_eval.py:523-527) and with sdpa - streamed on the same B200, same window definition."""
    import torch
    from oracle import cpu_baseline as cb

    out = {}
    for impl in ("flash_attention_2", "sdpa"):
        try:
            lm, vis, proj = cb.build_reference(device, 28, attn_impl=impl)
            s = cb.ReferenceStream(lm, vis, proj, n_gen=N_GEN)
            for _ in range(TURNS_PER_WINDOW):
                s.run_turn()
            t = [s.run_turn() for _ in range(windows * TURNS_PER_WINDOW)]
            sec = sum(x["seconds"] for x in t)
            steady = [x["seconds"] for x in t if x["kind"] == "steady"]
            start = [x["seconds"] for x in t if x["kind"] == "window_start"]
            out[impl] = {"value": round(windows * ENV_STEPS_PER_STEP / sec, 3), "unit": UNIT, "windows": windows,
                         "steady_turn_ms": round(1e3 * statistics.mean(steady), 2), "window_start_turn_ms": round(1e3 * statistics.mean(start), 2)}
            del lm, vis, proj, s
        except Exception as e:
            out[impl] = {"unavailable": repr(e)[:300]}
        torch.cuda.empty_cache()
    out["note"] = ("stock transformers " + __import__("transformers").__version__ + " Qwen2ForCausalLM + SiglipVisionModel (eager attention, as "
                   "the reference's tower), random-init 7B shapes, bf16, reference-semantics turn of oracle/glue.py (re-encodes 8 history frames "
                   "per window, lm_head on every prefill row, torch.cat KV appends), frames resident on the device; wall clock with "
                   "device synchronisation per turn")
    return out


def main_ours(a):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    from streamvln_b200 import build as svln_build

    svln_build.build()
    blocks = set(a.blocks.split(","))
    model, agent = build_agent(device, a.mode)
    # every rank streams the SAME synthetic episode (same frames, same random-init weights): the action sequences of
    # the timed regions must then be identical on every GPU - `stream_check` below
    rng = np.random.default_rng(1234)
    frames_host = rng.integers(0, 255, size=(64, 480, 640, 3), dtype=np.uint8)
    geom = None
    if a.mode == "slowfast":
        from streamvln_b200 import synth

        st = synth.make_stream(512, 14, 8, seed=rank, with_feat=False)
        geom = dict(depth=st["depth"], pose=st["pose"], intr=st["intr"])
    # frames resident in HBM for `value`: the image processor's output (fp32 [3,384,384]) staged on the device
    from PIL import Image

    proc = agent.image_processor
    frames_dev = torch.stack([proc.preprocess(images=Image.fromarray(f), return_tensors="pt")["pixel_values"][0]
                              for f in frames_host]).to(device)
    K, W = a.steps, a.warmup
    KT, WT = K * TURNS_PER_WINDOW, W * TURNS_PER_WINDOW
    prime(agent, frames_host, frames_dev, geom)
    # ---- device-resident arm
    agent.reset_memory()
    run_turns(agent, WT, frames_host, frames_dev, True, geom)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, _, _ = timed(agent, model, KT, frames_host, frames_dev, True, dist, device, geom)
    clocks = sampler.stop() if rank == 0 else None
    from streamvln_b200.scheduler import hash_actions

    hash_res, n_act = hash_actions(LAST_ACTIONS), len(LAST_ACTIONS)
    # ---- a few windows once more with CUDA events around the dominant kernels (events cannot be recorded inside a
    #      replayed CUDA graph, so this pass launches the kernels eagerly)
    model.engine.prof = {}
    n_prof = min(K, 4) * TURNS_PER_WINDOW
    ms_prof, _, _, _ = timed(agent, model, n_prof, frames_host, frames_dev, True, dist, device, geom)
    prof, model.engine.prof = model.engine.prof, None
    # ---- end-to-end arm (host uint8 frames through the public agent API)
    agent.reset_memory()
    run_turns(agent, WT, frames_host, frames_dev, False, geom)
    ms_e2e, _, h2d, d2h = timed(agent, model, KT, frames_host, frames_dev, False, dist, device, geom)
    hash_e2e = hash_actions(LAST_ACTIONS)
    hashes = torch.tensor([hash_res, hash_e2e], dtype=torch.int64, device=device)
    all_h = [hashes.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(all_h, hashes)
    stream_check = {"resident_action_hash": f"{hash_res:016x}", "e2e_action_hash": f"{hash_e2e:016x}", "actions": n_act,
                    "e2e_equals_resident": hash_e2e == hash_res,     # GPU image processor == PIL, same steps of the same episode
                    "ranks_agree": all(bool(torch.equal(h, hashes)) for h in all_h), "ranks": world,
                    "note": "every rank streams the same synthetic episode with the same random-init weights"}
    steps_total = K * ENV_STEPS_PER_STEP * world
    value = steps_total / (ms / 1e3)
    e2e = steps_total / (ms_e2e / 1e3)
    pk = peaks()
    roof = None
    ev = prof.get("decode_stack", [])
    if ev:
        t_ms = sum(s.elapsed_time(e) for s, e, _ in ev)
        nbytes = sum(w for _, _, w in ev)
        ach = nbytes / (t_ms / 1e3) / 1e9
        traffic, traffic_src = stack_traffic()
        roof = {"kernel": "decode_stack_kernel (persistent fused decode step: tensor-path weight-streaming GEMVs, RoPE/append, "
                          "split-KV attention, SwiGLU, lm_head, argmax)",
                "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(ach / pk["hbm_gbs"], 4), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": int(nbytes / len(ev)), "launches": len(ev),
                "avg_launch_us": round(1e3 * t_ms / len(ev), 1), "share_of_step": round(t_ms / ms_prof, 4),
                "peak_source": pk["source"],
                "note": "algorithmic bytes = all decoder weights + lm_head once per token + the stream's K/V cache "
                        "(57 344 B x cached length); timed in a second pass (eager launches, CUDA events around each launch: events "
                        "cannot sit inside a CUDA graph)"}
    roof_prefill = None
    ev = prof.get("attn_prefill", [])
    if ev:
        t_ms = sum(s.elapsed_time(e) for s, e, _ in ev)
        flops = sum(w for _, _, w in ev)
        ach = flops / (t_ms / 1e3) / 1e12
        roof_prefill = {"kernel": "attn_prefill3_kernel (tcgen05 TS-form, TMEM, TMA-fed)", "bound": "tensor", "achieved": round(ach, 2),
                        "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["bf16_tflops_sustained"], 4),
                        "traffic": None, "launches": len(ev), "avg_launch_us": round(1e3 * t_ms / len(ev), 2),
                        "share_of_step": round(t_ms / ms_prof, 4),
                        "note": "single-episode turns (latency-bound grids); the kernel's roofline is claimed at the 64-stream shape, see `kernels`"}
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": shared_config(a.mode, world),
        "config_notes": {"l2": "working set (15 GB of weights per token) exceeds the 126 MB L2; no flush needed (microbench rows flush L2 per call)",
                         "parity": "reference mode: pinned glue + HF-checked arithmetic (tests/); slowfast block: self-specified"},
        "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d / K) * world, "d2h_bytes_per_step": int(d2h / K) * world,
                "ms_per_step": round(ms_e2e / K, 3)},
        "gpu_launches": int(launches),
        "stream_check": stream_check,
        "roofline": roof,
        "roofline_prefill": roof_prefill,
        "clocks": clocks,
    }
    del agent
    if "slowfast" in blocks and a.mode == "reference":
        try:
            out["slowfast"] = slowfast_block(model, a, frames_host, frames_dev, dist, device, rank, world)
        except Exception as e:
            if dist is not None:
                raise
            out["slowfast"] = {"error": repr(e)[:300]}
    if "batched" in blocks:
        try:
            out["batched"] = batched_block(model, a, dist, device, rank, world)
        except Exception as e:
            if dist is not None:
                raise
            out["batched"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1:
        del model
        torch.cuda.empty_cache()
        if "kernels" in blocks:
            from streamvln_b200 import microbench

            out["kernels"] = microbench.run_all(device, pk)
            out["kernels"]["flash_attn_2"] = fa2_rows(device, pk)
            torch.cuda.empty_cache()
        if "hf" in blocks:
            out["hf_gpu"] = hf_gpu_leg(device)
        if "cpu" in blocks:
            out["cpu_baseline"] = cpu_baseline_leg()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------- reference (CPU port)
def _window_seconds(sample, last):
    """A window's time from the turn a step sampled: 7 steady turns + 1 window-start turn, using the most recent
    measurement of the other kind."""
    if sample["kind"] == "steady":
        return 7 * sample["seconds"] + last["window_start"]
    return 7 * last["steady"] + sample["seconds"]


def cpu_baseline_leg():
    """Bounded sample on the host cores: one window-start turn + two steady turns of the CPU port, all 28 layers."""
    from oracle import cpu_baseline as cb

    cores = cb.cores()
    lm, vis, proj = cb.build_reference("cpu", 28)
    s = cb.ReferenceStream(lm, vis, proj, n_gen=N_GEN)
    t = [s.run_turn() for _ in range(3)]
    win = 7 * min(t[1]["seconds"], t[2]["seconds"]) + t[0]["seconds"]
    return {"value": round(ENV_STEPS_PER_STEP / win, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 window-start turn ({t[0]['seconds']:.2f} s: SigLIP on 9 views + ~1.83k-row prefill + 4 decode tokens) and 2 steady "
                      f"turns ({t[1]['seconds']:.2f}, {t[2]['seconds']:.2f} s) of the reference-semantics turn (stock transformers SigLIP + Qwen2, "
                      f"7B shapes, ALL 28 layers executed, bf16, DynamicCache, lm_head on every prefill row); window = 7 x best steady + start"}


def main_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb

    cores = cb.cores()
    lm, vis, proj = cb.build_reference("cpu", 28)
    s = cb.ReferenceStream(lm, vis, proj, n_gen=N_GEN)
    last = {}
    first = s.run_turn()                       # the stream's first window-start turn (also warms thread pools / oneDNN)
    last["window_start"] = first["seconds"]
    last["steady"] = s.run_turn()["seconds"]
    for _ in range(max(a.warmup - 2, 0)):
        r = s.run_turn()
        last[r["kind"]] = r["seconds"]
    secs, kinds = [], []
    for _ in range(a.steps):
        r = s.run_turn()
        secs.append(_window_seconds(r, last))
        last[r["kind"]] = r["seconds"]
        kinds.append(r["kind"])
    total = sum(secs)
    value = a.steps * ENV_STEPS_PER_STEP / total
    sample = (f"each step executes ONE turn of a continuously streamed reference-semantics episode on the host cores (stock transformers "
              f"SigLIP 26 layers + Qwen2 7B shapes, ALL 28 decoder layers, bf16, carried DynamicCache with torch.cat appends, lm_head on every "
              f"prefill row) and prices its window as 7 steady turns + 1 window-start turn with the latest measurement of the other kind; "
              f"sampled {kinds.count('steady')} steady + {kinds.count('window_start')} window-start turns; last steady {last['steady']:.2f} s, "
              f"window-start {last['window_start']:.2f} s")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(1e3 * total / a.steps, 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": shared_config("reference", int(os.environ.get("WORLD_SIZE", 1))),
           "config_notes": {"arm": "CPU path of the reference (oracle port) on the box's host cores"},
           "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main_hf_gpu(a):
    import torch

    if int(os.environ.get("RANK", 0)) != 0:
        return
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    r = hf_gpu_leg(device, windows=max(a.steps, 1))
    best = max((v["value"] for v in r.values() if isinstance(v, dict) and "value" in v), default=None)
    print(json.dumps({"impl": "hf_gpu", "metric": METRIC, "value": best, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": 1,
                      "higher_is_better": True, "dtype": "bf16", "data": "synthetic", "config": {"step_definition": STEP_DEF}, "hf_gpu": r}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16, help="timed steps; one step = one 32-frame window = 8 model turns")
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "hf_gpu"])
    ap.add_argument("--mode", default="reference", choices=["reference", "slowfast"])
    ap.add_argument("--episodes", type=int, default=64, help="total concurrent episodes of the batched block (configs[2])")
    ap.add_argument("--blocks", default="slowfast,batched,kernels,hf,cpu",
                    help="extra blocks measured after the headline arms (kernels / hf / cpu only at N=1)")
    ap.add_argument("--emulate-world", default="",
                    help="1-GPU runs: W or a comma list (2,4,8): also run the batched block's shards of a W-rank job one after another (multi-GPU parity: "
                         "the digest must equal the W-rank run's action_hash_digest)")
    ap.add_argument("--no-micro", action="store_true", help="skip the per-kernel microbenchmarks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    b = [x for x in a.blocks.split(",") if x]
    if a.no_micro:
        b = [x for x in b if x != "kernels"]
    if a.no_cpu:
        b = [x for x in b if x != "cpu"]
    a.blocks = ",".join(b)
    if a.impl == "reference":
        main_reference(a)
    elif a.impl == "hf_gpu":
        main_hf_gpu(a)
    else:
        main_ours(a)


if __name__ == "__main__":
    main()

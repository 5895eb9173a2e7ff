#!/usr/bin/env python
"""bench.py — streaming nav steps/s of the StreamVLN SlowFast streaming-context hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode reference|slowfast]

Workload at N=1 (BASELINE.json configs[1]): StreamVLN LLaVA-Video-7B-Qwen2 single-episode streaming,
num_history=8, num_frames=32, num_future_steps=4, bf16, random-init weights (Qwen2-7B + SigLIP-so400m
shapes), synthetic 480x640 RGB frames, n_gen fixed to 5 (random weights never emit EOS).  One bench
"step" = one model TURN = num_future_steps (4) env steps (1 model call + 3 buffered frames, the cadence
of http_realworld_server.py:94-102); `value` = env steps per second.  N>1: every rank streams its own
episode (episodes shard with no collective on the hot path; NCCL only for the final gather) — weak scaling.

Every 8th turn starts a new 32-frame window and re-prefills ~1.9k tokens (37 ms against 21 ms for a steady turn), so
the warm-up is rounded up to whole windows: the K timed turns start with a window-start turn and contain one per 8.

Prints ONE JSON line (see the task contract): value (frames resident in HBM), e2e (through
VLNEvaluator.step with host uint8 frames: pinned H2D + GPU image processor + turn + D2H of the token ids),
roofline (the dominant kernel of the turn - the fused decode stack - timed live with CUDA events; roofline_prefill:
the tcgen05 prefill attention), cpu_baseline, kernels (per-kernel rooflines at their batched shapes), clocks.
"""
from __future__ import annotations

import argparse
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
N_GEN = 5
PRIME_TURNS = 25
# dram__bytes_read.sum + dram__bytes_write.sum of one decode_stack_kernel launch at the bench's shapes, from the
# `ncu --set full` capture summarised under profiles/ (None until a capture exists)
STACK_NCU_TRAFFIC_BYTES = 14_221_525_000   # profiles/r1_decode_stack_ncu_full.md: 14.2037 GB read + 17.87 MB written
INSTRUCTION = ("walk down the hallway past the kitchen turn left at the sofa go through the second door on the right "
               "and wait by the window next to the wooden table in the corner of the bedroom then stop near the bed "
               "facing the large mirror on the wall")  # 48 words ~ the 40 fixed token ids of SURVEY §8d


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a
    thread of this process (nvidia_ml_py): spawning `nvidia-smi -lms` at the start of the timed region and its
    periodic driver queries cost the first timed turns 3-12 ms each (measured); nvidia-smi is the fallback."""

    _BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.05):
        self.idx, self.period = gpu_index, period_s
        self.sm, self.mx, self.reasons = [], [], set()
        self.thread = self.proc = self.nvml = self.handle = None
        self.stop_flag = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps indices: go through the PCI bus id of the torch device
            import torch

            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
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


TURNS_PER_WINDOW = 8   # num_frames / num_future_steps = 32 / 4: every 8th turn re-prefills the window (reference mode)


def window_aligned(warmup: int) -> int:
    """Warm-up turns rounded up to whole windows, so that the timed region always STARTS with a window-start turn:
    K timed turns then contain ceil(K / 8) of the expensive turns - the workload's own 1-in-8 mix when K is a multiple
    of 8, a pessimistic mix for smaller K, never an all-steady-turn region (3 steady turns alone read 198 env steps/s)."""
    return (max(warmup, 3) + TURNS_PER_WINDOW - 1) // TURNS_PER_WINDOW * TURNS_PER_WINDOW


# ------------------------------------------------------------------------------------------ ours
def build_agent(device, mode: str):
    import torch
    from streamvln_b200.agent import VLNEvaluator
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
    from streamvln_b200.synth import intrinsic_matrix
    from streamvln_b200.tokenizer import StubQwenTokenizer

    cfg = StreamVLNConfig(mode=mode)  # Qwen2-7B + SigLIP-so400m shapes
    model = StreamVLNForCausalLM(cfg, n_env=1, device=device, seed=0)
    args = SimpleNamespace(num_frames=32, num_future_steps=ENV_STEPS_PER_TURN, num_history=8, max_new_tokens=N_GEN)
    agent = VLNEvaluator({"camera_intrinsic": intrinsic_matrix()}, model=model, tokenizer=StubQwenTokenizer(cfg.llm.vocab_size), args=args)
    return model, agent


def run_turns(agent, n_turns: int, frames_host, frames_dev, resident: bool, geom=None):
    """n_turns model turns = 4*n_turns env steps through VLNEvaluator.step (the call a user makes)."""
    h2d = d2h = 0
    verbose = os.environ.get("SVLN_BENCH_VERBOSE") == "1"
    marks = []
    for _ in range(n_turns):
        if verbose:
            import torch
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((agent.step_id, ev))
        for k in range(ENV_STEPS_PER_TURN):
            s = agent.step_id
            i = s % len(frames_host)
            kw = {}
            if geom is not None:
                kw = dict(depth=geom["depth"][s % len(geom["depth"])], pose=geom["pose"][s % len(geom["pose"])],
                          intrinsic=geom["intr"][0])
            if resident:
                agent.step(0, None, INSTRUCTION, run_model=(k == 0), preprocessed=frames_dev[i] if k == 0 else None, **kw)
            else:
                agent.step(0, frames_host[i], INSTRUCTION, run_model=(k == 0), **kw)
            if k == 0:
                h2d += getattr(agent, "h2d_bytes", 0)
                d2h += getattr(agent, "d2h_bytes", 0)
            agent.step_id += 1
    if verbose and len(marks) > 1:
        import torch
        torch.cuda.synchronize()
        print(("resident" if resident else "e2e") + " per-turn ms: " +
              " ".join(f"{sid}:{a.elapsed_time(b):.1f}" for (sid, a), (_, b) in zip(marks[:-1], marks[1:])), file=sys.stderr, flush=True)
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


def main_batched(a):
    """BASELINE.json configs[2]: E concurrent synthetic episodes sharded across the GPUs of one box (E/N per GPU,
    total fixed -> strong scaling); every model step packs the turns of all local streams into one forward."""
    import numpy as np
    import torch
    from streamvln_b200 import build as svln_build
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
    from streamvln_b200.scheduler import BatchedEvaluator
    from streamvln_b200.synth import intrinsic_matrix
    from streamvln_b200.tokenizer import StubQwenTokenizer

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    svln_build.build()
    E = a.episodes // world
    cfg = StreamVLNConfig(mode=a.mode)
    model = StreamVLNForCausalLM(cfg, n_env=E, device=device, seed=0)
    args = SimpleNamespace(num_frames=32, num_future_steps=ENV_STEPS_PER_TURN, num_history=8, max_new_tokens=N_GEN)
    ev = BatchedEvaluator(model, lambda: StubQwenTokenizer(cfg.llm.vocab_size), E, args, {"camera_intrinsic": intrinsic_matrix()})
    rng = np.random.default_rng(99 + rank)
    frames_host = rng.integers(0, 255, size=(16, 480, 640, 3), dtype=np.uint8)
    from PIL import Image

    proc = ev.agents[0].image_processor
    frames_dev = torch.stack([proc.preprocess(images=Image.fromarray(f), return_tensors="pt")["pixel_values"][0]
                              for f in frames_host]).to(device)
    live = list(range(E))
    instr = {i: INSTRUCTION for i in live}
    geom_src = None
    if a.mode == "slowfast":
        from streamvln_b200 import synth

        geom_src = [synth.make_stream(256, 14, 8, seed=1000 * rank + i, with_feat=False) for i in range(min(E, 8))]

    def turns(n, resident):
        for _ in range(n):
            for k in range(ENV_STEPS_PER_TURN):
                s = ev.agents[0].step_id
                geom = None
                if geom_src is not None:
                    geom = {i: dict(depth=geom_src[i % len(geom_src)]["depth"][s % 256], pose=geom_src[i % len(geom_src)]["pose"][s % 256],
                                    intrinsic=geom_src[i % len(geom_src)]["intr"][0]) for i in live}
                if resident:
                    ev.step_all(live, None, instr, k == 0, geom, {i: frames_dev[(s + i) % 16] for i in live} if k == 0 else None)
                else:
                    ev.step_all(live, {i: frames_host[(s + i) % 16] for i in live}, instr, k == 0, geom)

    def run(n, resident):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.engine.launches
        e0.record()
        turns(n, resident)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
        if dist is not None:
            dist.barrier()
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, model.engine.launches - l0

    K, W = a.steps, window_aligned(a.warmup)
    for n in (PRIME_TURNS, PRIME_TURNS, 9):   # graph priming: every shape captured and replayed once (see main_ours)
        for i in live:
            ev.reset_stream(i)
        turns(n, True)
    for i in live:
        ev.reset_stream(i)
    turns(W, True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = run(K, True)
    clocks = sampler.stop() if rank == 0 else None
    for i in live:
        ev.reset_stream(i)
    turns(W, False)   # whole windows: the timed e2e turns start at a window boundary too
    ms_e2e, _ = run(K, False)
    h2d_step = sum(int(getattr(ag, "h2d_bytes", 0)) for ag in ev.agents)   # bytes the agents copied for their last turn
    d2h_step = sum(int(getattr(ag, "d2h_bytes", 0)) for ag in ev.agents)
    total = E * world * K * ENV_STEPS_PER_TURN
    out = {"metric": METRIC, "value": round(total / (ms / 1e3), 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic",
           "config": {"workload": f"configs[2]: {a.episodes} concurrent synthetic R2R-shaped episodes sharded across {world} GPU(s) "
                                  f"({E} per GPU), LLaVA-Video-7B-Qwen2 shapes, num_history=8, n_gen=5, bf16, random-init weights",
                      "mode": a.mode, "env_steps_per_step": ENV_STEPS_PER_TURN, "episodes_per_gpu": E,
                      "step_definition": "1 bench step = 1 packed model turn of every stream = 4 env steps per stream"},
           "e2e": {"value": round(total / (ms_e2e / 1e3), 2), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 3),
                   "h2d_bytes_per_step": h2d_step * world, "d2h_bytes_per_step": d2h_step * world},
           "gpu_launches": int(launches), "clocks": clocks}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
    model, agent = build_agent(device, a.mode)
    rng = np.random.default_rng(1234 + rank)
    n_frames = 64
    frames_host = rng.integers(0, 255, size=(n_frames, 480, 640, 3), dtype=np.uint8)
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
    K, W = a.steps, window_aligned(a.warmup)
    # ---- priming (initialisation, like building the model): throw-away episodes so that every CUDA graph the
    #      steady state replays (vision, steady turn, window-start turns, decode) is captured AND replayed once
    #      before any warm-up or timed step.  A graph is captured on the second sighting of a shape and its first
    #      replay pays the upload (~12-16 ms).  The first window start of an episode (step 32) and the later ones
    #      (64, 96, ...) are different shapes, hence two episodes of three windows and a third of one.
    for n in (PRIME_TURNS, PRIME_TURNS, 9):
        agent.reset_memory()
        run_turns(agent, n, frames_host, frames_dev, True, geom)
    # ---- device-resident arm
    agent.reset_memory()
    run_turns(agent, W, frames_host, frames_dev, True, geom)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, _, _ = timed(agent, model, K, frames_host, frames_dev, True, dist, device, geom)
    clocks = sampler.stop() if rank == 0 else None
    # ---- same K turns once more with CUDA events around every prefill-attention launch (events cannot be
    #      recorded inside a replayed CUDA graph, so this pass launches the kernels eagerly)
    model.engine.prof = {}
    ms_prof, _, _, _ = timed(agent, model, K, frames_host, frames_dev, True, dist, device, geom)
    prof, model.engine.prof = model.engine.prof, None
    # ---- end-to-end arm (host uint8 frames through the public agent API)
    agent.reset_memory()
    run_turns(agent, W, frames_host, frames_dev, False, geom)
    ms_e2e, _, h2d, d2h = timed(agent, model, K, frames_host, frames_dev, False, dist, device, geom)
    steps_total = K * ENV_STEPS_PER_TURN * world
    value = steps_total / (ms / 1e3)
    e2e = steps_total / (ms_e2e / 1e3)
    pk = peaks()
    # ---- roofline of the dominant kernel of the turn: the fused decode stack (one persistent launch per generated
    #      token, HBM-bound weight streaming).  The tcgen05 prefill-attention kernel is reported next to it.
    roof = None
    ev = prof.get("decode_stack", [])
    if ev:
        t_ms = sum(s.elapsed_time(e) for s, e, _ in ev)
        nbytes = sum(w for _, _, w in ev)
        ach = nbytes / (t_ms / 1e3) / 1e9
        roof = {"kernel": "decode_stack_kernel (persistent fused decode step: tensor-path weight-streaming GEMVs, RoPE/append, "
                          "split-KV attention, SwiGLU, lm_head, argmax)",
                "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(ach / pk["hbm_gbs"], 4), "traffic": STACK_NCU_TRAFFIC_BYTES,
                "algorithmic_bytes_per_launch": int(nbytes / len(ev)), "launches": len(ev),
                "avg_launch_us": round(1e3 * t_ms / len(ev), 1), "share_of_step": round(t_ms / ms_prof, 4),
                "peak_source": pk["source"],
                "note": "algorithmic bytes = all decoder weights + lm_head once per token + the stream's K/V cache "
                        "(57 344 B x cached length); timed in a second pass of the same K turns with eager launches "
                        "(events cannot sit inside a CUDA graph); traffic = dram read+write bytes per launch from the ncu "
                        "--set full capture under profiles/r1_decode_stack_ncu_full.md"}
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
                        "note": "single-episode turn: ~217 new tokens x 7 GQA heads = 12 M-tiles x 4 kv heads = 48 CTAs on 148 SMs "
                                "(latency-bound; the kernel's roofline is claimed at the 64-stream shape, see `kernels`)"}
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: StreamVLN LLaVA-Video-7B-Qwen2 single-episode streaming step per GPU, num_history=8, "
                               "num_frames=32, num_future_steps=4, n_gen=5, bf16, random-init weights",
                   "mode": a.mode, "step_definition": "1 bench step = 1 model turn = 4 env steps; value counts env steps",
                   "env_steps_per_step": ENV_STEPS_PER_TURN, "episodes_per_gpu": 1, "parallelism": f"episode-sharded x{world}",
                   "l2": "working set (15 GB of weights per token) exceeds the 126 MB L2; no flush needed",
                   "turn_mix": "warm-up is rounded up to whole 8-turn windows: the timed turns start with a window-start "
                               "(1.9k-token re-prefill) turn and contain one per 8 turns"},
        "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d / K) * world, "d2h_bytes_per_step": int(d2h / K) * world,
                "ms_per_step": round(ms_e2e / K, 3)},
        "gpu_launches": int(launches),
        "roofline": roof,
        "roofline_prefill": roof_prefill,
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not a.no_micro:
        from streamvln_b200 import microbench

        del agent, model
        torch.cuda.empty_cache()
        out["kernels"] = microbench.run_all(device, pk)
    if rank == 0 and world == 1 and not a.no_cpu:
        out["cpu_baseline"] = cpu_baseline_leg(sample_layers=2, turns=1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------- reference
def cpu_baseline_leg(sample_layers: int, turns: int):
    from oracle import cpu_baseline as cb

    cores = cb.cores()
    lm, vis, proj, scale = cb.build_reference_cpu(sample_layers=sample_layers)
    cb.time_reference_turn(lm, vis, proj, scale, cached_len=200)  # warm-up (thread pools, oneDNN primitives)
    rs = [cb.time_reference_turn(lm, vis, proj, scale, cached_len=1000) for _ in range(turns)]
    best = min(rs, key=lambda r: r["turn_s"])
    return {"value": round(best["steps_per_s"], 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{turns} steady-state turn(s) (SigLIP 26 layers on 1 frame + 217-token prefill over a 1000-row cache + "
                      f"{N_GEN - 1} decode tokens, lm_head on every prefill row) with {sample_layers} of 28 decoder layers "
                      f"timed and scaled x{scale:.0f}; bf16 on CPU; turn = {best['turn_s']:.2f} s "
                      f"(vision {best['vision_s']:.2f}, lm_head {best['lm_head_s']:.2f}, layers sample {best['layers_sample_s']:.2f})"}


def main_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb

    cores = cb.cores()
    sample_layers = 4
    lm, vis, proj, scale = cb.build_reference_cpu(sample_layers=sample_layers)
    for _ in range(max(1, min(a.warmup, 3))):
        cb.time_reference_turn(lm, vis, proj, scale, cached_len=1000)
    t = []
    for _ in range(a.steps):
        t.append(cb.time_reference_turn(lm, vis, proj, scale, cached_len=1000)["turn_s"])
    total = sum(t)
    value = a.steps * ENV_STEPS_PER_TURN / total
    sample = (f"each step = 1 steady-state reference-semantics turn on CPU (stock transformers SigLIP 26 layers + Qwen2 at 7B "
              f"shapes, bf16, DynamicCache torch.cat appends, lm_head on every prefill row): {sample_layers} of 28 decoder "
              f"layers timed and scaled x{scale:.0f}, everything else timed in full")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(1e3 * total / a.steps, 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "configs[1]: StreamVLN LLaVA-Video-7B-Qwen2 single-episode streaming step, num_history=8, "
                                  "n_gen=5, bf16, random-init weights (CPU path of the reference)",
                      "env_steps_per_step": ENV_STEPS_PER_TURN},
           "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="reference", choices=["reference", "slowfast"])
    ap.add_argument("--episodes", type=int, default=1, help="total concurrent episodes (>1: configs[2], sharded across GPUs)")
    ap.add_argument("--no-micro", action="store_true", help="skip the per-kernel microbenchmarks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        main_reference(a)
    elif a.episodes > 1:
        main_batched(a)
    else:
        main_ours(a)


if __name__ == "__main__":
    main()

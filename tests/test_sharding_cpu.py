"""CPU tests of the N>1 host logic: episode sharding and the final metric gather over gloo (world_size 2),
the C-ABI export list, and host-side helpers."""
import os
import re
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def test_shard_episodes_matches_reference_rule():
    from streamvln_b200.scheduler import shard_episodes

    eps = list(range(11))
    parts = [shard_episodes(eps, r, 4) for r in range(4)]
    assert parts == [[0, 4, 8], [1, 5, 9], [2, 6, 10], [3, 7]]
    assert sorted(sum(parts, [])) == eps


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from streamvln_b200.scheduler import EpisodeResult, gather_results, hash_actions, shard_episodes

    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_episodes(list(range(7)), rank, world)
    local = [EpisodeResult(e, 40 + e, 10 + e, hash_actions([e % 4, 1, 2, 3] * (e + 1)), 0.5 * e) for e in mine]
    allr = gather_results(local)
    dist.barrier()
    q.put((rank, [(r.episode_id, r.steps, r.turns, r.action_hash, r.seconds) for r in allr]))
    dist.destroy_process_group()


def test_gather_results_gloo_world2():
    from streamvln_b200.scheduler import hash_actions

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [(e, 40 + e, 10 + e, hash_actions([e % 4, 1, 2, 3] * (e + 1)), 0.5 * e) for e in range(7)]
    assert got[0] == want and got[1] == want     # uneven shards (4 + 3 episodes), every rank sees all results


def test_cabi_exports_every_declared_symbol(svln_lib):
    """The shared library loads without a GPU and exports exactly what include/svln.h declares."""
    from streamvln_b200 import _lib

    header = (ROOT / "include" / "svln.h").read_text()
    declared = set(re.findall(r"\b(svln_[a-z0-9_]+)\s*\(", header))
    declared -= {"svln_stream_t"}
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(svln_lib, name), name
    assert svln_lib.svln_version() == 100
    # pure host-side size queries work without a device
    assert svln_lib.svln_kv_arena_bytes(1, 28, 4, 128, 4096) == 57344 * 4096
    assert svln_lib.svln_kv_arena_bytes(0, 28, 4, 128, 4096) == 0
    assert svln_lib.svln_prune_workspace_bytes(6272, 1, 6272) > 0
    assert svln_lib.svln_attn_workspace_bytes(64, 28, 128, 4096) > 0
    assert svln_lib.svln_voxel_prune_host_staging_bytes(1, 32, 384, 384, 14, 3584) > 32 * 196 * 3584 * 4


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    for f in (ROOT / "streamvln_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
    for f in (ROOT / "streamvln_b200" / "csrc").glob("*"):
        assert "oracle" not in f.read_text().replace("oracle/voxel_oracle.c", "").replace("CPU oracle", "").replace("oracle (", ""), f


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from streamvln_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(ImportError):
        _lib.load()


def test_stub_tokenizer_and_prompt_layout():
    """Prompt layout of SURVEY §3.1b with the offline tokenizer: system block only on a window's first turn, the
    empty assistant block is rendered closed, <image>/<memory> become -200/-300."""
    from streamvln_b200.tokenizer import StubQwenTokenizer

    tok = StubQwenTokenizer(152064)
    tok.add_tokens(["<image>"], special_tokens=True)
    tok.add_tokens(["<memory>"], special_tokens=True)
    ids = tok.apply_chat_template([{"role": "user", "content": "you can see <image>."}, {"role": "assistant", "content": ""}])
    im_s, im_e = tok.additional_special_tokens_ids
    assert ids[0] == im_s and ids.count(im_s) == 2 and ids.count(im_e) == 2 and ids[-1] == 198
    assert ids.count(tok.convert_tokens_to_ids("<image>")) == 1
    assert tok.batch_decode([[im_s, im_e, 198]])[0] == "<|im_start|><|im_end|>\n"
    assert tok.encode("walk to the door") == tok.encode("walk  to the\tdoor")


def test_synth_geometry_conventions():
    """Intrinsics and pose conventions of streamvln_eval.py:137-185."""
    import numpy as np
    from streamvln_b200 import synth

    K = synth.intrinsic_matrix()
    assert abs(K[0, 0] - 232.9) < 0.1 and abs(K[1, 1] - 310.6) < 0.1
    assert abs(K[0, 2] - 191.7) < 0.01 and abs(K[1, 2] - 191.6) < 0.01
    P = synth.pose_matrix(1.0, 2.0, 0.0)
    # camera z (forward) maps to episodic +x; camera x (right) maps to episodic -y; camera y (down) to -z
    assert np.allclose(P[:3, :3] @ np.array([0, 0, 1.0]), [1, 0, 0])
    assert np.allclose(P[:3, :3] @ np.array([1.0, 0, 0]), [0, -1, 0])
    assert np.allclose(P[:3, 3], [1.0, -2.0, 1.25])
    a, poses = synth.action_walk(50, np.random.default_rng(0))
    assert set(a.tolist()) <= {1, 2, 3} and poses.shape == (50, 4, 4)


def test_no_misencoded_ldgsts_in_the_library():
    """ptxas 12.9 has emitted (in a two-stream build of decode_stack.cu) LDGSTS instructions whose cache-policy
    descriptor operand is an odd uniform register (`desc[UR1]`): the B200 raises "Warp Illegal Instruction" on them.
    Guard the shipped library against that codegen fault (no GPU needed: cuobjdump reads the cubin)."""
    import re
    import shutil
    import subprocess
    from streamvln_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    bad = [l.strip() for l in sass.splitlines() if "LDGSTS" in l and re.search(r"desc\[UR\d*[13579]\]", l)]
    assert not bad, bad[:3]


def test_no_hoistable_coherent_loads_in_the_kernels():
    """`__ldcg` is a plain, non-volatile asm in the CUDA headers: nvcc may hoist it out of a loop or reuse an earlier result
    across a grid barrier.  It did, in the fused decode stack (a residual value written by the previous layer was read once,
    before the layer loop).  Data produced inside a launch is read with the asm-volatile ld_cg_* helpers of svln_common.cuh."""
    import re

    csrc = Path(__file__).resolve().parent.parent / "streamvln_b200" / "csrc"
    bad = []
    for f in sorted(list(csrc.glob("*.cu")) + list(csrc.glob("*.cuh"))):
        for n, line in enumerate(f.read_text().splitlines(), 1):
            code = line.split("//")[0]
            if re.search(r"\b__ldcg\s*\(", code):
                bad.append(f"{f.name}:{n}")
    assert not bad, bad


def test_decode_stack_abi_argument_checks():
    """svln_decode_stack_* validate their descriptor before touching the device (runs without a GPU): one stream only,
    head_dim 128, matching arena, non-null buffers; errors are negative errno-style codes with a message."""
    import ctypes as C
    from streamvln_b200 import _lib

    lib = _lib.load()
    d = _lib.DecodeStackStruct()
    d.n_layer, d.hidden, d.inter, d.n_q_head, d.n_kv_head, d.head_dim, d.vocab, d.n_stream = 28, 3584, 18944, 28, 4, 128, 152064, 1
    assert lib.svln_decode_stack_workspace_bytes(C.byref(d)) > 14 * 3584          # room for the activation vectors
    d.n_stream = 2
    assert lib.svln_decode_stack_workspace_bytes(C.byref(d)) > 2 * 14 * 3584      # two streams: packed feed only
    d.n_stream = 3
    assert lib.svln_decode_stack_workspace_bytes(C.byref(d)) == 0                 # at most two
    d.n_stream = 1
    arena = _lib.KvArenaStruct()
    assert lib.svln_decode_stack_step(C.byref(arena), C.byref(d), None) == -22    # null arena base
    assert b"svln_decode_stack_step" in lib.svln_last_error()
    arena.base, arena.n_env, arena.n_layer, arena.n_kv_head, arena.head_dim, arena.max_len = 4096, 1, 28, 4, 128, 4096
    d.n_stream = 3
    assert lib.svln_decode_stack_step(C.byref(arena), C.byref(d), None) == -34    # ERANGE: caller falls back
    d.n_stream, d.head_dim = 1, 64
    assert lib.svln_decode_stack_step(C.byref(arena), C.byref(d), None) == -22
    d.head_dim = 128
    assert lib.svln_decode_stack_step(C.byref(arena), C.byref(d), None) == -22    # null weight / token pointers
    # the packed weight stream is exactly the layer linears + lm_head, re-laid (no padding bytes), whatever the grid
    D, I, V, L = 3584, 18944, 152064, 28
    want = 2 * (L * ((28 + 2 * 4) * 128 * D + D * 28 * 128 + 2 * I * D + D * I) + V * D)
    assert lib.svln_decode_stack_packed_bytes(C.byref(d)) == want
    assert lib.svln_decode_stack_pack(C.byref(d), None, want, None) == -22        # null destination / weight table
    d.head_dim = 64
    assert lib.svln_decode_stack_packed_bytes(C.byref(d)) == 0
    d.head_dim = 128
    assert lib.svln_pool_bilinear(None, 1, 27, 14, 3584, None, None) == -22
    assert lib.svln_pool_bilinear(None, 0, 27, 14, 3584, None, None) == 0         # nothing to do
    assert lib.svln_pool_bilinear(None, 1, 27, 14, 3583, None, None) == -22       # dim must be a multiple of 8


def test_bench_window_alignment_and_peaks():
    """bench.py helpers that shape the measurement: one bench step is a whole 32-frame window (8 turns, one of them the
    window-start re-prefill), the CPU arm prices a window from the turn it sampled, and the roofline denominators come
    from MEASURED_PEAKS.json when the driver wrote it, else the documented fallback."""
    import importlib.util
    import sys

    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    sys.modules["bench_mod"] = bench
    spec.loader.exec_module(bench)
    assert [bench.window_aligned(w) for w in (0, 1, 3, 8, 9, 16, 17)] == [8, 8, 8, 8, 16, 16, 24]
    assert bench.ENV_STEPS_PER_STEP == 32 and "window" in bench.STEP_DEF
    last = {"steady": 1.0, "window_start": 9.0}
    assert bench._window_seconds({"kind": "steady", "seconds": 2.0}, last) == 7 * 2.0 + 9.0
    assert bench._window_seconds({"kind": "window_start", "seconds": 5.0}, last) == 7 * 1.0 + 5.0
    t, src = bench.stack_traffic()
    assert t is None or (t > 1e9 and src)
    pk = bench.peaks()
    assert pk["source"] in ("measured", "fallback") and pk["hbm_gbs"] > 1000 and pk["bf16_tflops"] >= pk["bf16_tflops_sustained"] > 100
    assert bench.ENV_STEPS_PER_TURN * bench.TURNS_PER_WINDOW == 32      # num_frames


def test_run_episodes_refills_slots_and_keeps_the_cadence():
    """scheduler.run_episodes (host logic, no GPU): E slots stream a longer list of episodes, a slot is refilled as soon
    as its episode ends, the model runs on every 4th step of each episode, every episode gets exactly its own frames and
    instruction, and results carry steps / turns / the action hash."""
    from streamvln_b200.scheduler import hash_actions, run_episodes

    class FakeEvaluator:
        def __init__(self, n):
            self.n, self.resets, self.model_steps, self.seen = n, [], [], {}

        def reset_stream(self, i):
            self.resets.append(i)

        def step_all(self, live, frames, instructions, run_model, geom=None, preprocessed=None):
            out = {}
            for i in live:
                ep = int(frames[i][0, 0, 0])                    # the frame source stamps the episode id
                self.seen.setdefault(ep, []).append((int(frames[i][0, 0, 1]), run_model, instructions[i]))
                if run_model:
                    out[i] = [1 + (ep + len(self.seen[ep])) % 3]
            if run_model:
                self.model_steps.append(sorted(live))
            return out

    eps = [dict(id=i, length=L, instruction=f"go {i}") for i, L in enumerate((9, 4, 13, 5, 8))]
    frame = lambda ep, step: np.full((2, 2, 3), 0, np.uint8) + np.array([ep, step, 0], np.uint8)
    import numpy as np
    ev = FakeEvaluator(2)
    res = run_episodes(ev, eps, frame)
    assert sorted(r.episode_id for r in res) == [0, 1, 2, 3, 4]
    by = {r.episode_id: r for r in res}
    for e in eps:
        steps = ev.seen[e["id"]]
        assert [s for s, _, _ in steps] == list(range(e["length"]))                     # its own frames, in order
        assert [rm for _, rm, _ in steps] == [s % 4 == 0 for s in range(e["length"])]     # model on every 4th step
        assert all(ins == e["instruction"] for _, _, ins in steps)
        assert by[e["id"]].steps == e["length"] and by[e["id"]].turns == (e["length"] + 3) // 4
    assert len(ev.resets) == 5 and ev.resets[:2] == [0, 1]                              # 2 initial fills + 3 refills
    # the hash is a pure function of the action list
    acts = [1 + (0 + k) % 3 for k in (1, 5, 9)]
    assert by[0].action_hash == hash_actions(acts)


def test_attention_item_waits_for_its_tile_on_the_full_ring():
    """attn_decode_body's FULL ring (the fused decode stack's 2-stage variant) refills a stage one iteration after it was
    consumed, so the newest cp.async group IS the tile about to be read: the loop must wait for every group there
    (`wait_group 1` let two-stream steps - splits of up to 4 tiles - read tiles still in flight; GPU regression test:
    tests/test_gpu_7b_shape.py::test_decode_stack_is_deterministic_at_full_depth)."""
    src = (Path(__file__).resolve().parent.parent / "streamvln_b200" / "csrc" / "attn_decode_body.cuh").read_text()
    assert "cp_async_wait<FULL ? 0 : PRO - 1>();" in src
    assert "cp_async_wait<PRO - 1>();" not in src

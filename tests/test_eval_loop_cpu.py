"""The Habitat evaluation driver (`streamvln_b200.habitat_eval`) PINNED to the reference's own loop.

tests/golden/ref_eval_loop.npz was produced by scripts/gen_golden_eval_loop.py, which imports the UNMODIFIED
streamvln/streamvln_eval.py (Habitat stubbed out) and runs `VLNEvaluator.eval_action` against tests/fake_env.py: seven
episodes in two scenes, a scripted recording model whose answers carry 0-5 actions, world sizes 1 and 2, and a resumed
run.  The same fake simulator and model drive the product's loop here; every model call (prompt ids, selected views,
time_ids, cache hand-over, geometry), the order of model resets, the simulator's action log, the result.json rows and the
returned metric vectors must be identical."""
import json
import os
import random
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_env  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


def _run(idx, env_num, out_dir, rank=0, num_history=8, config_path="config/vln_r2r.yaml"):
    from streamvln_b200.habitat_eval import VLNEvaluator
    from streamvln_b200.tokenizer import StubQwenTokenizer

    tok = StubQwenTokenizer(152064)
    model = fake_env.RecordingModel(tok)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=num_history, save_video=False, gpu_preprocess=False)
    random.seed(0)
    envs = []
    ev = VLNEvaluator(config_path=config_path, split="val_unseen", env_num=env_num, output_path=str(out_dir), model=model,
                      tokenizer=tok, epoch=0, args=args, env_factory=lambda: envs.append(fake_env.FakeEnv()) or envs[-1],
                      sim_sensors=fake_env.fake_sensors(), filter_depth=lambda d, blur_type=None: d, rank=rank)
    ret = ev.eval_action(idx)
    rows = [json.loads(l) for l in (Path(out_dir) / "result.json").read_text().splitlines()]
    return model, envs[0], rows, ret


def _check(z, prefix, model, env, rows, ret):
    g = lambda k: z[f"{prefix}_{k}"]
    calls = model.calls
    assert len(calls) == len(g("inputs_len"))
    for i, c in enumerate(calls):
        assert c["inputs"] == g("inputs")[i, : g("inputs_len")[i]].tolist(), (prefix, i)
        assert [v[0] for v in c["views"]] == g("view_step")[i, : g("n_views")[i]].tolist(), (prefix, i)
        assert c["views"][-1][1] == g("view_episode")[i]
        assert (len(c["time_ids"]), c["time_ids"][0], c["time_ids"][-1]) == (g("time_ids_len")[i], g("time_ids_first")[i], g("time_ids_last")[i])
        assert c["fresh"] == bool(g("fresh")[i]) and c["env_id"] == g("env_id")[i] and c["max_new_tokens"] == g("max_new_tokens")[i]
        # geometry: the reference hands bf16 tensors to the model (streamvln_eval.py:328-330), the product fp32
        bf = lambda a: torch.from_numpy(np.ascontiguousarray(a)).bfloat16().float().numpy()
        assert np.array_equal(bf(c["depths"][-1][::37, ::41]), g("depth_sample")[i]), (prefix, i, "depth")
        assert abs(float(bf(c["depths"]).astype(np.float64).sum()) - g("depth_sum")[i]) <= 1e-6 * abs(g("depth_sum")[i])
        assert np.allclose(c["poses"][-1], g("pose_last")[i], rtol=2.0 ** -8, atol=1e-6), (prefix, i, "pose")
        assert np.array_equal(bf(c["intrinsics"][-1]), g("intrinsic_last")[i]), (prefix, i, "intrinsic")
    resets = [(a, b) for a, b in model.resets if a != "reset"]
    assert resets == [tuple(r) for r in g("resets").tolist()]
    assert [(int(e), a) for e, a in env.log] == [tuple(r) for r in g("env_log").tolist()]
    assert [int(r["episode_id"]) for r in rows if "scene_id" in r] == g("row_episode").tolist()
    assert [r["steps"] for r in rows] == g("row_steps").tolist()
    assert [r["scene_id"].encode() for r in rows] == g("row_scene").tolist()
    assert [r["episode_instruction"].encode() for r in rows] == g("row_instruction").tolist()
    assert np.array_equal(np.array([[r["success"], r["spl"], r["os"], r["ne"]] for r in rows], np.float64), g("row_metrics"))
    assert np.array_equal(np.array([v.tolist() for v in ret[:4]], np.float64), g("ret")) and int(ret[4]) == int(g("ret_n"))
    assert env.closed


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "ref_eval_loop.npz")


def test_fixture_covers_the_loop(gold):
    z = gold
    assert z["w1r0_row_episode"].tolist() == [1, 2, 5, 7, 3, 4, 6]            # scenes sorted, dataset order inside a scene
    assert z["w1r0_row_steps"].tolist()[2] == 66 and z["w1r0_row_steps"].tolist()[4] == 70      # two windows crossed
    assert 9 in z["w1r0_n_views"].tolist()                                    # a window start with 8 history views
    assert set(z["w1r0_dtypes"].tolist()) == {b"torch.bfloat16"}
    assert z["w2r0_row_episode"].tolist() == [1, 5, 3, 6] and z["w2r1_row_episode"].tolist() == [2, 7, 4]


def test_eval_action_is_pinned_to_the_reference_loop(gold, tmp_path):
    _check(gold, "w1r0", *_run(0, 1, tmp_path))


@pytest.mark.parametrize("rank", [0, 1])
def test_episode_sharding_is_pinned(gold, tmp_path, rank):
    _check(gold, f"w2r{rank}", *_run(rank, 2, tmp_path, rank=rank))


def test_resume_from_result_json_is_pinned(gold, tmp_path):
    (tmp_path / "result.json").write_text("\n".join(l.decode() for l in gold["resume_prefill"].tolist()) + "\n")
    _check(gold, "resume", *_run(0, 1, tmp_path))


def test_history_slice_without_num_history_is_pinned(gold, tmp_path):
    """args.num_history = None: every num_future_steps-th past frame becomes a history view (streamvln_eval.py:314-315)."""
    assert gold["nohist_n_views"].max() == 17 and gold["w1r0_n_views"].max() == 9      # step 64: frames 0, 4, ..., 60 + the current one (8 + 1 with num_history = 8)
    _check(gold, "nohist", *_run(0, 1, tmp_path, num_history=None))


def test_objectnav_rows_are_pinned(gold, tmp_path):
    """'objectnav' in the config path: result rows (and the resume key) carry the episode's object category (:221)."""
    assert gold["objnav_row_instruction"].tolist()[0] == b"category1"
    _check(gold, "objnav", *_run(0, 1, tmp_path, config_path="config/objectnav_hm3d.yaml"))


def test_only_selected_frames_are_preprocessed(tmp_path):
    """The driver buffers raw observations and runs the image / depth processors on the views a model call selects."""
    from streamvln_b200 import habitat_eval

    n = {"img": 0}
    orig = fake_env.StampProcessor.preprocess

    def counting(self, images, return_tensors="pt"):
        n["img"] += 1
        return orig(self, images, return_tensors)

    fake_env.StampProcessor.preprocess = counting
    try:
        model, env, rows, _ = _run(0, 1, tmp_path)
    finally:
        fake_env.StampProcessor.preprocess = orig
    steps = sum(r["steps"] for r in rows)
    distinct = len({(c["views"][-1][1], v[0]) for c in model.calls for v in c["views"]})
    assert n["img"] == distinct < steps


def _worker(rank, world, out_dir, port, q):
    import torch.distributed as dist
    from streamvln_b200.habitat_eval import evaluate
    from streamvln_b200.tokenizer import StubQwenTokenizer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tok = StubQwenTokenizer(152064)
    model = fake_env.RecordingModel(tok)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, save_video=False, gpu_preprocess=False, output_path=out_dir)
    random.seed(0)
    res = evaluate(model, tok, args, env_factory=fake_env.FakeEnv, sim_sensors=fake_env.fake_sensors(),
                   filter_depth=lambda d, blur_type=None: d)
    q.put((rank, res))
    dist.destroy_process_group()


def test_evaluate_gathers_over_two_ranks(gold, tmp_path):
    """`evaluate` (streamvln_eval.py:536-581) with gloo, world size 2: the summary averages the metric vectors of both
    ranks, i.e. the reference's per-rank returns concatenated."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, str(tmp_path), port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=180) for _ in procs)
    [p.join(60) for p in procs]
    both = np.concatenate([gold["w2r0_ret"], gold["w2r1_ret"]], axis=1)
    want = dict(sucs_all=both[0].mean(), spls_all=both[1].mean(), oss_all=both[2].mean(), ones_all=both[3].mean(), length=both.shape[1])
    for r in (0, 1):
        assert got[r]["length"] == want["length"] == 7
        for k in ("sucs_all", "spls_all", "oss_all", "ones_all"):
            assert abs(got[r][k] - want[k]) < 1e-9, (k, got[r][k], want[k])
    last = (tmp_path / "result.json").read_text().splitlines()[-1]
    assert json.loads(last[last.index('{"sucs_all"'):])["length"] == 7


def _per_episode(model, env_logs):
    calls = {}
    for c in model.calls:
        calls.setdefault(c["views"][-1][1], []).append((c["inputs"], [v[0] for v in c["views"]], c["time_ids"], c["fresh"],
                                                        c["poses"].tobytes(), c["depths"][-1][::37, ::41].tobytes()))
    acts = {}
    for log in env_logs:
        for e, a in log:
            acts.setdefault(int(e), []).append(a)
    return calls, acts


@pytest.mark.parametrize("n_slots", [2, 3, 8])
def test_concurrent_evaluator_equals_the_sequential_loop(tmp_path, n_slots):
    """E simulators per rank with packed model turns: per episode the model calls (prompt ids, views, time_ids, cache
    hand-over, geometry), the actions and the result row are those of the one-episode-at-a-time loop."""
    from streamvln_b200.habitat_eval import ConcurrentEvaluator, VLNEvaluator
    from streamvln_b200.tokenizer import StubQwenTokenizer

    def make(kind, out):
        tok = StubQwenTokenizer(152064)
        model = fake_env.RecordingModel(tok, by_step=True)
        args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, save_video=False, gpu_preprocess=False)
        envs = []
        kw = dict(config_path="config/vln_r2r.yaml", split="val_unseen", env_num=1, output_path=str(out), model=model, tokenizer=tok,
                  epoch=0, args=args, env_factory=lambda: envs.append(fake_env.FakeEnv()) or envs[-1], sim_sensors=fake_env.fake_sensors(),
                  filter_depth=lambda d, blur_type=None: d, rank=0, episode_seed=7)
        ev = VLNEvaluator(**kw) if kind == "seq" else ConcurrentEvaluator(n_slots, **kw)
        ret = ev.eval_action(0)
        rows = [json.loads(l) for l in (Path(out) / "result.json").read_text().splitlines()]
        return model, envs, rows, ret

    ms, es, rs, rets = make("seq", tmp_path / "seq")
    mc, ec, rc, retc = make("conc", tmp_path / "conc")
    assert len(ec) == n_slots and all(e.closed for e in ec)
    cs, as_ = _per_episode(ms, [e.log for e in es])
    cc, ac = _per_episode(mc, [e.log for e in ec])
    assert cs == cc and as_ == ac
    key = lambda r: int(r["episode_id"])
    assert sorted(rs, key=key) == sorted(rc, key=key)
    assert sorted(zip(*[v.tolist() for v in rets[:4]])) == sorted(zip(*[v.tolist() for v in retc[:4]])) and int(rets[4]) == int(retc[4]) == 7
    assert max(mc.batches) == min(n_slots, 7) and sum(mc.batches) <= len(mc.calls)       # single requests go through generate()
    # every stream of a packed call is a different model slot
    assert all(0 <= c["env_id"] < n_slots for c in mc.calls)


def test_resume_after_a_finished_run(gold, tmp_path):
    """`evaluate` appends its summary without a newline (streamvln_eval.py:580-581): a later resume finds the next row on
    the same line.  Every episode of the finished run is recognised as done."""
    _run(0, 1, tmp_path)
    with open(tmp_path / "result.json", "a") as f:
        f.write(json.dumps({"sucs_all": 0.5, "spls_all": 0.1, "oss_all": 0.5, "ones_all": 1.0, "length": 7}))
    model, env, rows, ret = _run(0, 1, tmp_path)
    assert len(model.calls) == 0 and env.log == [] and int(ret[4]) == 7
    assert np.array_equal(np.array([v.tolist() for v in ret[:4]], np.float64), gold["w1r0_ret"])


def test_concurrent_evaluator_resumes(gold, tmp_path):
    """result.json already holds three finished episodes: the concurrent driver runs the other four and returns all seven."""
    from streamvln_b200.habitat_eval import ConcurrentEvaluator
    from streamvln_b200.tokenizer import StubQwenTokenizer

    (tmp_path / "result.json").write_text("\n".join(l.decode() for l in gold["resume_prefill"].tolist()) + "\n")
    tok = StubQwenTokenizer(152064)
    model = fake_env.RecordingModel(tok, by_step=True)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, save_video=False, gpu_preprocess=False)
    ev = ConcurrentEvaluator(3, config_path="config/vln_r2r.yaml", split="val_unseen", env_num=1, output_path=str(tmp_path), model=model,
                             tokenizer=tok, epoch=0, args=args, env_factory=fake_env.FakeEnv, sim_sensors=fake_env.fake_sensors(),
                             filter_depth=lambda d, blur_type=None: d, rank=0)
    ret = ev.eval_action(0)
    rows = [json.loads(l) for l in (tmp_path / "result.json").read_text().splitlines()]
    done = [json.loads(l.decode())["episode_id"] for l in gold["resume_prefill"].tolist()]
    assert int(ret[4]) == 7 and len(rows) == 7
    assert sorted(r["episode_id"] for r in rows[3:]) == sorted({"1", "2", "3", "4", "5", "6", "7"} - set(done))
    assert {c["views"][-1][1] for c in model.calls} == {int(e) for e in {"1", "2", "3", "4", "5", "6", "7"} - set(done)}

"""Generate tests/golden/ref_eval_loop.npz by running the REFERENCE's own Habitat evaluation loop.

`VLNEvaluator.eval_action` (streamvln/streamvln_eval.py:192-380) is plain Python around three things that are absent here:
the Habitat simulator, the model and the tokenizer.  The module is imported UNMODIFIED with `habitat*`, `quaternion`,
`omegaconf` and `depth_camera_filtering` stubbed in `sys.modules`; `habitat.Env` is `tests/fake_env.FakeEnv` (a
deterministic simulator), the model is `tests/fake_env.RecordingModel` (records every `generate(**kw)` call and answers
with a scripted action sequence of 0-5 actions), the tokenizer is the offline stand-in.  What is recorded: every model
call's `inputs`, selected views, `time_ids`, fresh/carried cache flag, samples of the depth / pose / intrinsic tensors,
the order of `reset_for_env` calls, the (episode, action) log of the simulator, and the `result.json` rows - for rank 0
of 1 and for ranks 0 and 1 of 2 (the `episodes[idx::env_num]` sharding), plus a resumed run (result.json pre-filled).
Runs only in the build container (needs /root/reference, read-only; nothing is copied from it).

    python scripts/gen_golden_eval_loop.py
"""
import contextlib
import io
import json
import random
import sys
import tempfile
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/streamvln")

import fake_env  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Cfg(dict):
    def update(self, *a, **k):   # habitat.task.measurements.update(...)
        pass


def install_stubs():
    sensors = fake_env.fake_sensors()
    cfg = SimpleNamespace(habitat=SimpleNamespace(
        simulator=SimpleNamespace(agents=SimpleNamespace(main_agent=SimpleNamespace(sim_sensors=sensors))),
        dataset=SimpleNamespace(split=None), task=SimpleNamespace(measurements=_Cfg())))
    _stub("quaternion", as_rotation_matrix=lambda q: np.eye(3))
    _stub("omegaconf", OmegaConf=SimpleNamespace(to_yaml=lambda c: ""))
    _stub("depth_camera_filtering", filter_depth=lambda d, blur_type=None: d)    # identity: the real filter is not installed
    hab = _stub("habitat", logger=None, Env=lambda config=None: fake_env.FakeEnv())
    hab.config = _stub("habitat.config", read_write=lambda c: contextlib.nullcontext())
    _stub("habitat.config.default", get_agent_config=lambda sim: None)
    kw = lambda **k: k
    _stub("habitat.config.default_structured_configs", CollisionsMeasurementConfig=kw, FogOfWarConfig=kw, TopDownMapMeasurementConfig=kw)
    _stub("habitat_extensions", measures=None)
    _stub("habitat_extensions.measures")
    _stub("habitat_baselines")
    _stub("habitat_baselines.config")
    _stub("habitat_baselines.config.default", get_config=lambda path: cfg)
    _stub("habitat.utils")
    _stub("habitat.utils.visualizations", maps=None)
    _stub("habitat.utils.visualizations.utils", images_to_video=None, observations_to_image=None)
    q = types.ModuleType("llava.model.multimodal_resampler.qformer")
    q.Qformer = type("Qformer", (), {})
    sys.modules["llava.model.multimodal_resampler.qformer"] = q


def run_reference(idx: int, env_num: int, out_dir: str, seed: int = 0, num_history=8, config_path="config/vln_r2r.yaml"):
    import streamvln_eval as ref                       # the reference module, unmodified
    from streamvln_b200.tokenizer import StubQwenTokenizer

    tok = StubQwenTokenizer(152064)
    model = fake_env.RecordingModel(tok)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=num_history, save_video=False)
    random.seed(seed)
    envs = []
    ref.Env = lambda config=None: envs.append(fake_env.FakeEnv()) or envs[-1]
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        ev = ref.VLNEvaluator(config_path=config_path, split="val_unseen", env_num=env_num, output_path=out_dir,
                              model=model, tokenizer=tok, epoch=0, args=args)
        ev.device = torch.device("cpu")                  # the reference hard-codes cuda (streamvln_eval.py:53)
        sucs, spls, oss, ones, n = ev.eval_action(idx)
    rows = [json.loads(l) for l in (Path(out_dir) / "result.json").read_text().splitlines()]
    return dict(model=model, env=envs[0], rows=rows, sucs=sucs.tolist(), spls=spls.tolist(), oss=oss.tolist(), ones=ones.tolist(), n=int(n))


def pack(run, prefix):
    m, calls = run["model"], run["model"].calls
    pad = lambda rows, fill: np.array([list(r) + [fill] * (max(map(len, rows)) - len(r)) for r in rows], np.int64)
    out = {
        "inputs": pad([c["inputs"] for c in calls], -1), "inputs_len": np.array([len(c["inputs"]) for c in calls]),
        "view_step": pad([[v[0] for v in c["views"]] for c in calls], -1), "n_views": np.array([len(c["views"]) for c in calls]),
        "view_episode": np.array([c["views"][-1][1] for c in calls]),
        "time_ids_len": np.array([len(c["time_ids"]) for c in calls]), "time_ids_first": np.array([c["time_ids"][0] for c in calls]),
        "time_ids_last": np.array([c["time_ids"][-1] for c in calls]),
        "fresh": np.array([c["fresh"] for c in calls]), "env_id": np.array([c["env_id"] for c in calls]),
        "max_new_tokens": np.array([c["max_new_tokens"] for c in calls]),
        # geometry of the newest view of every call (what the loop computed for that step), bf16-rounded as the reference
        # passes it to the model (streamvln_eval.py:328-330)
        "depth_sample": np.stack([c["depths"][-1][::37, ::41] for c in calls]).astype(np.float32),
        "depth_sum": np.array([float(c["depths"].astype(np.float64).sum()) for c in calls]),
        "pose_last": np.stack([c["poses"][-1] for c in calls]).astype(np.float32),
        "intrinsic_last": np.stack([c["intrinsics"][-1] for c in calls]).astype(np.float32),
        "resets": np.array([(a, b) for a, b in m.resets if a != "reset"], np.int64).reshape(-1, 2),
        "env_log": np.array([(int(e), a) for e, a in run["env"].log], np.int64),
        "row_episode": np.array([int(r["episode_id"]) for r in run["rows"]]), "row_steps": np.array([r["steps"] for r in run["rows"]]),
        "row_scene": np.array([r["scene_id"].encode() for r in run["rows"]]),
        "row_metrics": np.array([[r["success"], r["spl"], r["os"], r["ne"]] for r in run["rows"]], np.float64),
        "row_instruction": np.array([r["episode_instruction"].encode() for r in run["rows"]]),
        "ret": np.array([run["sucs"], run["spls"], run["oss"], run["ones"]], np.float64), "ret_n": np.int64(run["n"]),
        "dtypes": np.array([d.encode() for d in calls[0]["dtypes"]]),
    }
    return {f"{prefix}_{k}": v for k, v in out.items()}


if __name__ == "__main__":
    install_stubs()
    OUT.mkdir(parents=True, exist_ok=True)
    blob = {}
    with tempfile.TemporaryDirectory() as d:
        r = run_reference(0, 1, d)
        blob.update(pack(r, "w1r0"))
        print("world 1:", len(r["model"].calls), "model calls,", len(r["rows"]), "episodes, steps", [x["steps"] for x in r["rows"]])
        first_rows = (Path(d) / "result.json").read_text().splitlines()
    for rank in (0, 1):
        with tempfile.TemporaryDirectory() as d:
            r = run_reference(rank, 2, d)
            blob.update(pack(r, f"w2r{rank}"))
            print(f"world 2 rank {rank}:", len(r["model"].calls), "model calls, episodes", [x["episode_id"] for x in r["rows"]])
    with tempfile.TemporaryDirectory() as d:                 # resume: the first three finished episodes are already on disk
        (Path(d) / "result.json").write_text("\n".join(first_rows[:3]) + "\n")
        r = run_reference(0, 1, d)
        blob.update(pack(r, "resume"))
        blob["resume_prefill"] = np.array([l.encode() for l in first_rows[:3]])
        print("resumed:", len(r["model"].calls), "model calls, rows on disk", len(r["rows"]), "returned n", r["n"])
    with tempfile.TemporaryDirectory() as d:                 # num_history=None: the history slice steps by num_future_steps (:314-315)
        r = run_reference(0, 1, d, num_history=None)
        blob.update(pack(r, "nohist"))
        print("num_history=None:", len(r["model"].calls), "model calls, views per call up to", max(len(c["views"]) for c in r["model"].calls))
    with tempfile.TemporaryDirectory() as d:                 # an objectnav config: result rows carry the object category (:221)
        r = run_reference(0, 1, d, config_path="config/objectnav_hm3d.yaml")
        blob.update(pack(r, "objnav"))
        print("objectnav:", [x["episode_instruction"] for x in r["rows"]][:3])
    np.savez_compressed(OUT / "ref_eval_loop.npz", **blob)
    print("wrote", OUT / "ref_eval_loop.npz", (OUT / "ref_eval_loop.npz").stat().st_size, "bytes")

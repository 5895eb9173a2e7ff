"""Generate tests/golden/ref_agent_*.npz / ref_tail_rule.npz / ref_geometry.npz by running the REFERENCE's own code.

Runs only in the build container (needs /root/reference, read-only; nothing is copied from it).  Pins three pieces of
the hot path that are plain Python in the reference and therefore runnable here:

  1. `VLNEvaluator.step` (streamvln/streamvln_agent.py:169-258), driven for 104 env steps with a fake model that
     records every `generate(**kwargs)` call and returns fixed `.sequences`.  `quaternion`, `omegaconf` and
     `depth_camera_filtering` (absent here, unused by `step`) are stubbed in `sys.modules`; the image processor is a
     stub that stamps each frame with its step id so that the views the agent selects can be read back from `images`.
  2. `StreamVLNForCausalLM.prepare_inputs_for_generation` (streamvln/model/stream_video_vln.py:409-471), a pure function
     of its arguments, called unbound for the three `cache_position` cases (window start, carried cache + new
     embeds, one-token decode).
  3. `get_intrinsic_matrix`, `preprocess_instrinsic`, `get_axis_align_matrix`, `xyz_yaw_to_tf_matrix` and
     `preprocess_depth_image` (streamvln/streamvln_eval.py:126-185), AST-extracted from the Habitat evaluator (whose
     module imports habitat) and called on the sensor constants of config/vln_r2r.yaml and seeded poses / depths.

    python scripts/gen_golden_agent.py
"""
import ast
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/streamvln")

for name in ("quaternion", "omegaconf", "depth_camera_filtering"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["omegaconf"].OmegaConf = object
sys.modules["depth_camera_filtering"].filter_depth = lambda *a, **k: a[0]
stub = types.ModuleType("llava.model.multimodal_resampler.qformer")
stub.Qformer = type("Qformer", (), {})
sys.modules["llava.model.multimodal_resampler.qformer"] = stub

INSTRUCTION = "walk past the sofa and stop at the door"
N_STEPS = 104           # windows start at 32, 64 and 96: three memory re-prefills with different history strides


class StampProcessor:
    """image_processor stand-in: the 'pixel_values' of a frame is a [3,2,2] tensor holding the frame's first pixel."""
    crop_size = {"height": 384, "width": 384}

    def preprocess(self, images, return_tensors="pt"):
        v = float(np.asarray(images)[0, 0, 0])
        return {"pixel_values": [torch.full((3, 2, 2), v)]}


class RecordingModel:
    def __init__(self, eos_id):
        self.calls, self.resets, self.eos_id = [], [], eos_id
        self.tower = SimpleNamespace(image_processor=StampProcessor())
        self.n_cached = 0

    def get_vision_tower(self):
        return self.tower

    def reset_for_env(self, idx):
        self.resets.append((len(self.calls), idx))

    def generate(self, **kw):
        t = len(self.calls)
        seq = [(t * 7 + j * 3) % 50 + 10 for j in range(4)] + [self.eos_id]
        views = kw["images"][0, :, 0, 0, 0].float().round().to(torch.int64).tolist()
        self.calls.append(dict(
            inputs=kw["inputs"][0].tolist(), views=views, time_ids=list(kw["time_ids"][0]), env_id=kw["env_id"],
            fresh=kw["past_key_values"] is None, max_new_tokens=kw["max_new_tokens"], task_ids=kw["task_ids"],
            n_images=kw["images"].shape[1], depth_shape=list(kw["depths"].shape), pose_shape=list(kw["poses"].shape),
            intr_shape=list(kw["intrinsics"].shape), dtypes=[str(kw[k].dtype) for k in ("images", "depths", "poses", "intrinsics")],
            flags=[kw["do_sample"], kw["num_beams"], kw["use_cache"], kw["return_dict_in_generate"]]))
        return SimpleNamespace(sequences=torch.tensor([seq]), past_key_values=("kv", t))


def agent_fixture():
    import streamvln_agent as ref_agent                      # the reference module, unmodified
    from streamvln_b200.tokenizer import StubQwenTokenizer   # offline stand-in for the Qwen2 tokenizer (no files here)

    tok = StubQwenTokenizer(152064)
    model = RecordingModel(tok.im_end)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8)
    K = np.array([[192.0, 0, 191.42857143, 0], [0, 192.0, 191.42857143, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        ev = ref_agent.VLNEvaluator({"rgb_height": 1.25, "camera_intrinsic": K}, model=model, tokenizer=tok, args=args)
        ev.device = torch.device("cpu")                      # the reference hard-codes cuda:0 (streamvln_agent.py:40)
        actions, texts = [], []
        for s in range(N_STEPS):
            rgb = np.full((480, 640, 3), s % 256, np.uint8)
            a, _, txt = ev.step(0, rgb, INSTRUCTION, run_model=(s % 4 == 0))
            actions.append([-1] if a is None else a)
            texts.append("" if txt is None else txt)
            ev.step_id += 1                                  # the caller increments it (http_realworld_server.py:112)
    n = len(model.calls)
    pad = lambda rows, fill: np.array([r + [fill] * (max(map(len, rows)) - len(r)) for r in rows], np.int64)
    np.savez_compressed(
        OUT / "ref_agent_104steps.npz",
        instruction=np.bytes_(INSTRUCTION.encode()), n_steps=np.int64(N_STEPS),
        inputs=pad([c["inputs"] for c in model.calls], -1), inputs_len=np.array([len(c["inputs"]) for c in model.calls]),
        views=pad([c["views"] for c in model.calls], -1), views_len=np.array([len(c["views"]) for c in model.calls]),
        time_ids=pad([c["time_ids"] for c in model.calls], -1), time_ids_len=np.array([len(c["time_ids"]) for c in model.calls]),
        fresh=np.array([c["fresh"] for c in model.calls]), max_new_tokens=np.array([c["max_new_tokens"] for c in model.calls]),
        resets=np.array(model.resets, np.int64), actions=pad(actions, -1), actions_len=np.array([len(a) for a in actions]),
        texts=np.array([t.encode() for t in texts]), call_step=np.arange(0, N_STEPS, 4),
        dtypes=np.array([d.encode() for d in model.calls[0]["dtypes"]]),
        depth_shape=np.array(model.calls[8]["depth_shape"]), pose_shape=np.array(model.calls[8]["pose_shape"]))
    print(f"agent: {n} model calls, resets at calls {[r[0] for r in model.resets]}, window-start views {model.calls[8]['views']}, "
          f"{model.calls[16]['views']}, {model.calls[24]['views']}")


def tail_rule_fixture():
    from model.stream_video_vln import StreamVLNForCausalLM as Ref

    D = 4
    cases = []

    def call(name, n_embeds, past_len, n_ids, first):
        embeds = torch.arange(n_embeds * D, dtype=torch.float32).reshape(1, n_embeds, D)
        ids = torch.arange(100, 100 + n_ids)[None]
        # HF 4.45.1 _get_initial_cache_position: arange(L_full)[past_len:] on the first forward of a generate call, then
        # one position per decoded token
        cache_position = torch.arange(past_len, n_embeds) if first else torch.tensor([n_embeds + n_ids - 1])
        out = Ref.prepare_inputs_for_generation(None, ids, past_key_values=object(), attention_mask=None, inputs_embeds=embeds,
                                                cache_position=cache_position, position_ids=None, use_cache=True)
        e = out["inputs_embeds"]
        cases.append(dict(name=name, n_embeds=n_embeds, past_len=past_len, n_ids=n_ids, first=first,
                          uses_embeds=e is not None, first_row=-1 if e is None else int(e[0, 0, 0]) // D,
                          n_rows=0 if e is None else e.shape[1], ids_out=[] if out["input_ids"] is None else out["input_ids"][0].tolist(),
                          position_ids_none=out["position_ids"] is None))

    call("window_start", 1768, 0, 0, True)            # cache_position[0] == 0 -> every embedded token
    call("carried_cache", 1768 + 221, 1768 + 4, 0, True)   # len(cache_position) > 1 -> embeds[:, -len(cache_position):]
    call("one_uncached", 300, 299, 0, True)           # a single un-cached embed row: falls to the input_ids branch
    call("decode", 1768, 1768, 3, False)              # one position -> last input id
    np.savez_compressed(OUT / "ref_tail_rule.npz", **{k: np.array([c[k] if not isinstance(c[k], list) else (c[k] + [-1])[0] for c in cases])
                                                       for k in ("n_embeds", "past_len", "n_ids", "first", "uses_embeds", "first_row", "n_rows",
                                                                 "ids_out", "position_ids_none")},
                        names=np.array([c["name"].encode() for c in cases]))
    for c in cases:
        print("tail rule:", c)


def geometry_fixture():
    src = Path(REF + "/streamvln/streamvln_eval.py").read_text()
    tree = ast.parse(src)
    want = {"get_intrinsic_matrix", "preprocess_instrinsic", "get_axis_align_matrix", "xyz_yaw_to_tf_matrix", "preprocess_depth_image"}
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "VLNEvaluator")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {f.name for f in fns} == want
    mod = ast.Module(body=fns, type_ignores=[])
    import copy
    from PIL import Image
    from transformers.image_utils import to_numpy_array
    ns = dict(np=np, torch=torch, copy=copy, Image=Image, to_numpy_array=to_numpy_array)
    exec(compile(mod, "streamvln_eval.py[extract]", "exec"), ns)
    me = SimpleNamespace(image_processor=SimpleNamespace(crop_size={"height": 384, "width": 384}))
    sensor = SimpleNamespace(width=640, height=480, hfov=79)          # config/vln_r2r.yaml:22-30
    K = ns["get_intrinsic_matrix"](me, sensor)
    K384 = ns["preprocess_instrinsic"](me, K, (640, 480), (384, 384))  # ori_size = (w, h), target (w, h): :262,283
    align = ns["get_axis_align_matrix"](me).numpy()
    rng = np.random.default_rng(0)
    xyz = rng.uniform(-5, 5, size=(6, 3))
    yaw = rng.uniform(-np.pi, np.pi, size=6)
    tf = np.stack([ns["xyz_yaw_to_tf_matrix"](me, xyz[i], yaw[i]) for i in range(6)])
    pose = np.stack([tf[i] @ align for i in range(6)])                 # streamvln_eval.py:282
    depth01 = rng.uniform(0, 1, size=(480, 640)).astype(np.float32)
    depth_mm = (depth01 * 10.0 * 1000).astype(np.uint16)               # :255-258 (min_depth 0, max_depth 10 m)
    d384, size = ns["preprocess_depth_image"](me, Image.fromarray(depth_mm), do_depth_scale=True)
    np.savez_compressed(OUT / "ref_geometry.npz", K=K, K384=K384, axis_align=align, xyz=xyz, yaw=yaw, tf=tf, pose=pose,
                        depth01=depth01, depth384=np.asarray(d384, np.float64), depth_size=np.array(size))
    print("geometry: fx", K[0, 0], "K384 row0", K384[0], "depth384", np.asarray(d384).shape, np.asarray(d384).dtype)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    agent_fixture()
    tail_rule_fixture()
    geometry_fixture()

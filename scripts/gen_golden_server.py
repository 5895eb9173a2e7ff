"""Generate tests/golden/ref_server.npz by driving the REFERENCE's own Flask handler
(streamvln/http_realworld_server.py:61-129 `eval_vln`), unmodified, with a stubbed `flask` module (not installed here),
a fake `request` per call and a recording evaluator.  Pins the wire-level control flow of the `/eval_vln` endpoint:
JPEG decode + channel flip, reset handling, 4 agent steps per request with the model on every 4th agent step, the STOP
latch, the response.  Runs only in the build container (/root/reference is read-only; nothing is copied from it).

    python scripts/gen_golden_server.py
"""
import io
import json
import sys
import types
import zlib
from pathlib import Path
from types import SimpleNamespace

import numpy as np

REF = "/root/reference"
ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/streamvln")

for name in ("quaternion", "omegaconf", "depth_camera_filtering"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["omegaconf"].OmegaConf = object
sys.modules["depth_camera_filtering"].filter_depth = lambda *a, **k: a[0]
q = types.ModuleType("llava.model.multimodal_resampler.qformer")
q.Qformer = type("Qformer", (), {})
sys.modules["llava.model.multimodal_resampler.qformer"] = q
flask = types.ModuleType("flask")


class _Flask:
    def __init__(self, *a, **k):
        pass

    def route(self, *a, **k):
        return lambda f: f


flask.Flask, flask.request, flask.jsonify = _Flask, SimpleNamespace(), (lambda d: d)
sys.modules["flask"] = flask

# scripted model outputs per model call: (action list, decoded text); a STOP in call 5 latches `terminate`
SCRIPT = [[1, 1, 2, 1], [3, 1, 1, 1], [1, 2, 2, 1], [1, 1, 1, 1], [2, 1, 3, 1], [1, 0], [1, 1, 1, 1], [3, 3, 1, 1], [1, 1, 1, 2], [0]]
RESETS = {0, 9}            # requests that carry reset = true
N_REQ = 14


def frame(i):
    rng = np.random.default_rng(100 + i)
    base = rng.integers(0, 255, size=(6, 8, 3), dtype=np.uint8)
    return np.kron(base, np.ones((40, 40, 1), np.uint8))     # 240 x 320 blocky image: survives JPEG recognisably


def jpeg(rgb):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(rgb).save(b, format="jpeg")
    return b.getvalue()


class RecordingEvaluator:
    def __init__(self):
        self.step_id = 0
        self.calls, self.resets, self.n_model = [], [], 0

    def reset_memory(self):
        self.resets.append(len(self.calls))
        self.step_id = 0

    def eval_action(self, idx, image, depth, camera_pose, instruction, run_model=False):
        self.calls.append((int(idx), int(zlib.crc32(np.ascontiguousarray(image).tobytes())), tuple(image.shape), bool(run_model),
                           tuple(depth.shape), instruction))
        if not run_model:
            return None, 0, None
        a = SCRIPT[self.n_model % len(SCRIPT)]
        self.n_model += 1
        return list(a), 0.01, "".join("STOP" if x == 0 else "↑←→"[x - 1] for x in a)


def main():
    import contextlib
    import http_realworld_server as srv          # the reference module, unmodified

    srv.annotate_image = lambda *a, **k: None     # the PNG dump needs a TTF font and writes under ./runs...
    srv.os.makedirs = lambda *a, **k: None
    ev = RecordingEvaluator()
    srv.evaluator = ev
    answers, jpegs = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(N_REQ):
            jb = jpeg(frame(i))
            jpegs.append(np.frombuffer(jb, np.uint8))
            srv.request.files = {"image": SimpleNamespace(stream=io.BytesIO(jb))}
            srv.request.form = {"json": json.dumps({"reset": i in RESETS})}
            out = srv.eval_vln()
            answers.append([int(a) for a in out["action"]])
    pad = lambda rows, fill: np.array([list(r) + [fill] * (max(map(len, rows)) - len(r)) for r in rows], np.int64)
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(
        OUT / "ref_server.npz", n_req=np.int64(N_REQ), resets=np.array(sorted(RESETS)), script=pad(SCRIPT, -1),
        script_len=np.array([len(a) for a in SCRIPT]), answers=pad(answers, -1), answers_len=np.array([len(a) for a in answers]),
        call_idx=np.array([c[0] for c in ev.calls]), call_crc=np.array([c[1] for c in ev.calls], np.int64),
        call_shape=np.array([c[2] for c in ev.calls]), call_run=np.array([c[3] for c in ev.calls]),
        depth_shape=np.array([c[4] for c in ev.calls]), instruction=np.bytes_(ev.calls[0][5].encode()),
        reset_at_call=np.array(ev.resets), jpeg=np.concatenate(jpegs), jpeg_len=np.array([len(j) for j in jpegs]))
    print("answers", answers)
    print("evaluator calls", len(ev.calls), "model calls", ev.n_model, "resets at call", ev.resets)


if __name__ == "__main__":
    main()

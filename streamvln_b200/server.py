"""The reference's real-world wire format on top of the B200 agent: `POST /eval_vln`.

Mirrors streamvln/http_realworld_server.py:61-129 (Flask handler) and the robot-side client
realworld/go2_vln_client.py:35-54: a multipart request with one JPEG file part `image` and one form field `json` =
`{"reset": bool}`; the answer is `{"action": [ints]}` with 0 = STOP, 1 = forward 25 cm, 2 / 3 = turn left / right 15 deg.
Per request the handler feeds the SAME frame to the agent `num_future_steps` (4) times, running the model on every 4th
agent step (`step_id % 4 == 0`, :94-102), keeps the last action sequence, and latches `terminate` once a STOP appears -
every later request answers `[0]` until a request with `reset: true` clears the episode (:75-89).

Differences, none on the wire: no Flask (not in the image: the endpoint is a stdlib `http.server`), the annotated PNG dump
of every frame (:24-59) is optional (`annotate_dir`), the instruction is a constructor argument instead of the constant at
:72, and the agent method called is `step` - the reference handler calls `evaluator.eval_action`, which its own
`VLNEvaluator` (streamvln_agent.py) does not define.  The frame is handed over channel-flipped exactly as the reference
does (`np.asarray(image)[..., ::-1]`, :68-70).
"""
from __future__ import annotations

import io
import json
import threading
from email.parser import BytesParser
from email.policy import HTTP
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

DEFAULT_INSTRUCTION = "Walk forward and immediately stop when you exit the room."   # http_realworld_server.py:72


class EvalVLNService:
    """State machine of the `/eval_vln` handler (one episode at a time, like the reference's module globals :18-23)."""

    def __init__(self, evaluator, instruction: str = DEFAULT_INSTRUCTION, num_future_steps: int = 4, env_idx: int = 0):
        self.evaluator = evaluator
        self.instruction = instruction
        self.num_future_steps = num_future_steps
        self.env_idx = env_idx
        self.action_seq: List[int] = [0, 0, 0, 0]      # np.zeros(4) in the reference
        self.idx = 0
        self.terminate = False
        self.last_llm_output: Optional[str] = None
        self.lock = threading.Lock()                    # the Flask dev server is single-threaded; keep requests serial

    def handle(self, image_bytes: bytes, data: Dict[str, Any]) -> Dict[str, List[int]]:
        from PIL import Image

        image = np.asarray(Image.open(io.BytesIO(image_bytes)).convert("RGB"))[..., ::-1]      # :66-70
        with self.lock:
            if data["reset"]:                                                                  # :75-85
                self.terminate = False
                self.idx = 0
                self.evaluator.reset_memory()
            self.idx += 1
            if self.terminate:                                                                 # :89-91
                return {"action": [0]}
            for _ in range(self.num_future_steps):                                             # :93-112
                ev = self.evaluator
                action, _, llm_output = ev.step(self.env_idx, image, self.instruction, run_model=(ev.step_id % 4 == 0))
                if llm_output is not None:
                    self.last_llm_output = llm_output
                if action is not None:
                    self.action_seq = [int(a) for a in action]
                if 0 in self.action_seq:
                    self.terminate = True
                ev.step_id += 1
            if len(self.action_seq) == 0:                                                      # :125-127
                return {"action": [0]}
            return {"action": list(self.action_seq)}


def parse_multipart(content_type: str, body: bytes) -> Tuple[Dict[str, bytes], Dict[str, str]]:
    """(files, form) of a multipart/form-data body: parts with a filename are files, the others text fields."""
    msg = BytesParser(policy=HTTP).parsebytes(b"Content-Type: " + content_type.encode() + b"\r\nMIME-Version: 1.0\r\n\r\n" + body)
    files: Dict[str, bytes] = {}
    form: Dict[str, str] = {}
    if not msg.is_multipart():
        raise ValueError("expected multipart/form-data")
    for part in msg.iter_parts():
        name = part.get_param("name", header="content-disposition")
        if name is None:
            continue
        payload = part.get_payload(decode=True)
        if part.get_filename() is not None:
            files[name] = payload
        else:
            form[name] = payload.decode("utf-8")
    return files, form


def make_handler(service: EvalVLNService):
    class Handler(BaseHTTPRequestHandler):
        def log_message(self, *a):      # quiet
            pass

        def do_POST(self):
            if self.path.rstrip("/") != "/eval_vln":
                self.send_error(404)
                return
            try:
                body = self.rfile.read(int(self.headers.get("Content-Length", "0")))
                files, form = parse_multipart(self.headers.get("Content-Type", ""), body)
                out = service.handle(files["image"], json.loads(form["json"]))
                payload = json.dumps(out).encode()
                self.send_response(200)
            except (KeyError, ValueError) as e:     # Flask answers 400 for a missing file / form field
                payload = json.dumps({"error": repr(e)}).encode()
                self.send_response(400)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(payload)))
            self.end_headers()
            self.wfile.write(payload)

    return Handler


def serve(service: EvalVLNService, host: str = "0.0.0.0", port: int = 5801) -> ThreadingHTTPServer:
    """Start the endpoint (the reference listens on 5801, go2_vln_client.py:35); returns the server (call
    `serve_forever()` or run it in a thread; `shutdown()` stops it)."""
    return ThreadingHTTPServer((host, port), make_handler(service))


def post_frame(url: str, rgb: np.ndarray, reset: bool, timeout: float = 150.0) -> List[int]:
    """The robot-side call (go2_vln_client.py:35-54): JPEG-encode the frame, POST it with {"reset": ...}, return `action`."""
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, format="jpeg")
    return post_jpeg(url, buf.getvalue(), reset, timeout)


def post_jpeg(url: str, jpeg: bytes, reset: bool, timeout: float = 150.0) -> List[int]:
    import urllib.request
    import uuid

    buf = io.BytesIO(jpeg)
    boundary = uuid.uuid4().hex
    body = (f"--{boundary}\r\nContent-Disposition: form-data; name=\"json\"\r\n\r\n{json.dumps({'reset': bool(reset)})}\r\n"
            f"--{boundary}\r\nContent-Disposition: form-data; name=\"image\"; filename=\"rgb_image\"\r\nContent-Type: image/jpg\r\n\r\n").encode() \
        + buf.getvalue() + f"\r\n--{boundary}--\r\n".encode()
    req = urllib.request.Request(url, data=body, headers={"Content-Type": f"multipart/form-data; boundary={boundary}"}, method="POST")
    with urllib.request.urlopen(req, timeout=timeout) as r:
        return json.loads(r.read().decode())["action"]

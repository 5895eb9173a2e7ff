"""`/eval_vln` end to end on the GPU: JPEG frames over HTTP (127.0.0.1) -> EvalVLNService -> the product agent
(`streamvln_b200.agent.VLNEvaluator` on a tiny StreamVLN) -> actions, against the same agent stepped directly with the
frames the handler decodes (http_realworld_server.py:61-129 cadence: 4 agent steps per request, model on every 4th)."""
import io
import threading
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_eval_vln_endpoint_drives_the_product_agent(svln_lib, cuda):
    from PIL import Image
    from streamvln_b200.agent import VLNEvaluator
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
    from streamvln_b200.server import EvalVLNService, post_frame, serve
    from streamvln_b200.synth import intrinsic_matrix
    from streamvln_b200.tokenizer import StubQwenTokenizer

    cfg = StreamVLNConfig.tiny(vocab_size=512, layers=2)
    m = StreamVLNForCausalLM(cfg, n_env=1, device=cuda, seed=3, lm_head_std=0.3)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, max_new_tokens=4)
    mk = lambda: VLNEvaluator({"camera_intrinsic": intrinsic_matrix()}, model=m, tokenizer=StubQwenTokenizer(512), args=args)
    rng = np.random.default_rng(1)
    frames = [np.kron(rng.integers(0, 255, size=(6, 8, 3), dtype=np.uint8), np.ones((40, 40, 1), np.uint8)) for _ in range(10)]
    svc = EvalVLNService(mk(), instruction="go to the kitchen")
    srv = serve(svc, "127.0.0.1", 0)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        url = f"http://127.0.0.1:{srv.server_address[1]}/eval_vln"
        got = [post_frame(url, f, reset=(i == 0), timeout=60) for i, f in enumerate(frames)]
    finally:
        srv.shutdown()
        srv.server_close()
    assert all(a and all(x in (0, 1, 2, 3) for x in a) for a in got)
    # the same episode stepped directly: what the handler decodes from the JPEG (channel-flipped) fed 4 times per request
    ev = mk()
    ev.reset_memory()
    want, last, stopped = [], [0, 0, 0, 0], False
    for f in frames:
        if stopped:
            want.append([0])
            continue
        b = io.BytesIO()
        Image.fromarray(f).save(b, format="jpeg")
        img = np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGB"))[..., ::-1]
        for _ in range(4):
            a, _, _ = ev.step(0, img, "go to the kitchen", run_model=(ev.step_id % 4 == 0))
            last = last if a is None else a
            stopped |= 0 in last
            ev.step_id += 1
        want.append(list(last))
    assert got == want
    assert svc.evaluator.step_id == ev.step_id

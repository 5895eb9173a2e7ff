"""The `/eval_vln` wire format pinned to the REFERENCE'S OWN Flask handler: the fixture
tests/golden/ref_server.npz was produced by scripts/gen_golden_server.py, which imports streamvln/http_realworld_server.py
unmodified (flask stubbed) and drives `eval_vln()` for 14 requests with a recording evaluator.  Checked here: the service's
state machine in process, and the same over a real HTTP round trip on 127.0.0.1 with multipart bodies as the robot client
builds them (realworld/go2_vln_client.py:35-54)."""
import threading
import zlib
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden" / "ref_server.npz"


class RecordingAgent:
    """Duck type of VLNEvaluator as the handler uses it: step / step_id / reset_memory."""

    def __init__(self, script):
        self.script, self.step_id, self.calls, self.resets, self.n_model = script, 0, [], [], 0

    def reset_memory(self):
        self.resets.append(len(self.calls))
        self.step_id = 0

    def step(self, idx, rgb, instruction_text="", run_model=False):
        self.calls.append((int(idx), int(zlib.crc32(np.ascontiguousarray(rgb).tobytes())), tuple(rgb.shape), bool(run_model), instruction_text))
        if not run_model:
            return None, 0, None
        a = self.script[self.n_model % len(self.script)]
        self.n_model += 1
        return list(a), 0.01, "".join("STOP" if x == 0 else "↑←→"[x - 1] for x in a)


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    off = np.concatenate([[0], np.cumsum(z["jpeg_len"])])
    return dict(z=z, jpegs=[z["jpeg"][off[i]:off[i + 1]].tobytes() for i in range(int(z["n_req"]))],
                script=[z["script"][i, :z["script_len"][i]].tolist() for i in range(z["script"].shape[0])],
                answers=[z["answers"][i, :z["answers_len"][i]].tolist() for i in range(int(z["n_req"]))],
                resets=set(z["resets"].tolist()))


def _check_calls(agent, z):
    assert [c[0] for c in agent.calls] == z["call_idx"].tolist()
    assert [c[1] for c in agent.calls] == z["call_crc"].tolist()          # same decoded, channel-flipped pixels
    assert [list(c[2]) for c in agent.calls] == z["call_shape"].tolist()
    assert [c[3] for c in agent.calls] == z["call_run"].tolist()           # model on every 4th agent step
    assert agent.resets == z["reset_at_call"].tolist()
    assert all(c[4] == bytes(z["instruction"]).decode() for c in agent.calls)


def test_service_state_machine_matches_reference_handler(gold):
    from streamvln_b200.server import EvalVLNService

    agent = RecordingAgent(gold["script"])
    svc = EvalVLNService(agent)
    got = [svc.handle(gold["jpegs"][i], {"reset": i in gold["resets"]})["action"] for i in range(len(gold["jpegs"]))]
    assert got == gold["answers"]
    _check_calls(agent, gold["z"])


def test_http_round_trip_matches_reference_handler(gold):
    from streamvln_b200.server import EvalVLNService, post_jpeg, serve

    agent = RecordingAgent(gold["script"])
    srv = serve(EvalVLNService(agent), "127.0.0.1", 0)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        url = f"http://127.0.0.1:{srv.server_address[1]}/eval_vln"
        got = [post_jpeg(url, gold["jpegs"][i], i in gold["resets"], timeout=30) for i in range(len(gold["jpegs"]))]
    finally:
        srv.shutdown()
        srv.server_close()
    assert got == gold["answers"]
    _check_calls(agent, gold["z"])


def test_bad_requests_are_rejected(gold):
    import urllib.error
    import urllib.request
    from streamvln_b200.server import EvalVLNService, serve

    srv = serve(EvalVLNService(RecordingAgent(gold["script"])), "127.0.0.1", 0)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        base = f"http://127.0.0.1:{srv.server_address[1]}"
        for path, body, ctype, code in (("/eval_vln", b"x", "text/plain", 400), ("/nope", b"", "text/plain", 404)):
            req = urllib.request.Request(base + path, data=body, headers={"Content-Type": ctype}, method="POST")
            with pytest.raises(urllib.error.HTTPError) as e:
                urllib.request.urlopen(req, timeout=10)
            assert e.value.code == code
    finally:
        srv.shutdown()
        srv.server_close()

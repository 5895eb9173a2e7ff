"""The Habitat evaluation driver (`streamvln_b200.habitat_eval`) on the GPU with a real (tiny) StreamVLN: the loop with the
CUDA image / depth processors and pinned H2D must walk the fake simulator exactly like the same loop with the host (PIL /
numpy) processors - the kernels are bit-exact with the host paths, so the action logs are identical - and the slowfast
mode must run the same driver."""
import json
import random
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_env  # noqa: E402

pytestmark = pytest.mark.gpu


class NoStopEnv(fake_env.FakeEnv):
    """A random-init model says STOP within a few turns; walking forward instead lets every episode run to its step cap
    (up to 70 steps: two window boundaries)."""

    def step(self, action):
        return super().step(action if action != 0 else 1)


def _walk(cuda, out_dir, gpu_preprocess, mode="reference", concurrent=1):
    from streamvln_b200.habitat_eval import evaluate
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
    from streamvln_b200.tokenizer import StubQwenTokenizer

    cfg = StreamVLNConfig.tiny(vocab_size=512, layers=2, mode=mode)
    m = StreamVLNForCausalLM(cfg, n_env=max(1, concurrent), device=cuda, seed=3, lm_head_std=0.3)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, max_new_tokens=4, save_video=False,
                           gpu_preprocess=gpu_preprocess, output_path=str(out_dir), concurrent_episodes=concurrent)
    envs = []
    random.seed(0)
    res = evaluate(m, StubQwenTokenizer(512), args, env_factory=lambda: envs.append(NoStopEnv()) or envs[-1],
                   sim_sensors=fake_env.fake_sensors(), filter_depth=lambda d, blur_type=None: d)
    rows = [json.loads(l) for l in (Path(out_dir) / "result.json").read_text().splitlines()[:-1]]
    return res, rows, [x for e in envs for x in e.log]


def test_driver_gpu_processors_equal_host_processors(svln_lib, cuda, tmp_path):
    a = _walk(cuda, tmp_path / "gpu", True)
    b = _walk(cuda, tmp_path / "host", False)
    assert a[2] == b[2] and a[1] == b[1] and a[0] == b[0]
    assert a[0]["length"] == 7 and len(a[1]) == 7 and all(r["steps"] >= 1 for r in a[1])
    assert sorted(r["steps"] for r in a[1]) == [5, 9, 21, 33, 37, 66, 70]     # the step caps: two window boundaries crossed


def test_driver_runs_the_slowfast_mode(svln_lib, cuda, tmp_path):
    res, rows, log = _walk(cuda, tmp_path / "sf", True, mode="slowfast")
    assert res["length"] == 7 and len(log) == sum(r["steps"] for r in rows)


@pytest.mark.parametrize("mode", ["reference", "slowfast"])
def test_driver_with_concurrent_simulators(svln_lib, cuda, tmp_path, mode):
    """Three simulators on one model replica, model turns packed through generate_batch: every episode runs to its step cap
    and the job reports the same episodes as the sequential driver (token-level equality of packed and single streams is
    bounded in tests/test_gpu_batched.py)."""
    res, rows, _ = _walk(cuda, tmp_path / "conc", True, mode=mode, concurrent=3)
    assert res["length"] == 7 and sorted(r["steps"] for r in rows) == [5, 9, 21, 33, 37, 66, 70]
    assert sorted(int(r["episode_id"]) for r in rows) == [1, 2, 3, 4, 5, 6, 7]

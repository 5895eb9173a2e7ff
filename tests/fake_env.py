"""A deterministic stand-in for `habitat.Env` (test infrastructure; Habitat is not installed here).

Implements exactly the surface the reference's Habitat evaluation loop touches (streamvln/streamvln_eval.py:192-380):
`episodes`, `current_episode`, `reset()`, `step(action)`, `episode_over`, `get_metrics()`, `sim.get_agent_state()`,
`close()`; observations carry `rgb` uint8 [480,640,3], `depth` float32 [480,640,1] in [0,1], `gps` [x, y], `compass`
[yaw].  Motion follows config/vln_r2r.yaml:34-35 (0.25 m forward, 15 degree turns).  Used by
scripts/gen_golden_eval_loop.py (to drive the REFERENCE's own `eval_action`) and by tests/test_eval_loop_cpu.py (to drive
`streamvln_b200.habitat_eval`), so both loops see the same simulator.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List

import numpy as np

H, W = 480, 640


def make_episodes() -> List[SimpleNamespace]:
    """Seven episodes in two scenes; `max_steps` bounds episodes whose fake model never says STOP."""
    eps = []
    spec = [("sceneB", 3, "go down the hall and stop by the plant", 70), ("sceneA", 1, "walk past the sofa", 37),
            ("sceneA", 2, "turn left at the stairs", 9), ("sceneB", 4, "enter the kitchen", 33),
            ("sceneA", 5, "leave the room", 66), ("sceneB", 6, "find the red chair", 21), ("sceneA", 7, "stop at the door", 5)]
    for scene, eid, text, cap in spec:
        eps.append(SimpleNamespace(scene_id=f"data/scene_datasets/mp3d/{scene}/{scene}.glb", episode_id=str(eid),
                                   instruction=SimpleNamespace(instruction_text=text), object_category=f"category{eid}", max_steps=cap))
    return eps


class FakeSim:
    def __init__(self, env):
        self.env = env

    def get_agent_state(self):
        e = self.env
        # Habitat world frame: y up; the evaluator only uses position[1] (height above the start) and the rotation handle
        return SimpleNamespace(position=np.array([e.x, 0.1 + 0.01 * e.n_steps, e.y], np.float32), rotation=("quat", e.yaw))


class FakeEnv:
    def __init__(self):
        self.episodes = make_episodes()
        self.current_episode = None
        self.sim = FakeSim(self)
        self.closed = False
        self.log = []                 # (episode_id, action) in call order
        self._over = True
        self.x = self.y = self.yaw = 0.0
        self.n_steps = 0

    # ---- observations ---------------------------------------------------------------------------------------------
    def _obs(self):
        k = int(self.current_episode.episode_id) * 1000 + self.n_steps
        rgb = np.empty((H, W, 3), np.uint8)
        rgb[..., 0] = self.n_steps % 256                       # the frame's step id, read back by the stamp processor
        rgb[..., 1] = int(self.current_episode.episode_id)
        rgb[..., 2] = (np.arange(W, dtype=np.int64)[None, :] + k) % 251
        yy, xx = np.mgrid[0:H, 0:W]
        depth = (((xx * 7 + yy * 13 + k * 31) % 1000) / 1000.0).astype(np.float32)[..., None]   # [0,1), every pixel distinct-ish
        return dict(rgb=rgb, depth=depth, gps=np.array([self.x, self.y], np.float32), compass=np.array([self.yaw], np.float32))

    def reset(self):
        assert self.current_episode is not None
        self.x = 0.5 * int(self.current_episode.episode_id)
        self.y = -0.25 * int(self.current_episode.episode_id)
        self.yaw = 0.1 * int(self.current_episode.episode_id)
        self.n_steps = 0
        self._over = False
        return self._obs()

    def step(self, action: int):
        assert not self._over
        self.log.append((self.current_episode.episode_id, int(action)))
        if action == 1:
            self.x += 0.25 * math.cos(self.yaw)
            self.y += 0.25 * math.sin(self.yaw)
        elif action == 2:
            self.yaw += math.radians(15.0)
        elif action == 3:
            self.yaw -= math.radians(15.0)
        self.n_steps += 1
        if action == 0 or self.n_steps >= self.current_episode.max_steps:
            self._over = True
        return self._obs()

    @property
    def episode_over(self) -> bool:
        return self._over

    def get_metrics(self):
        e = int(self.current_episode.episode_id)
        d = float(np.float32(abs(self.x) + abs(self.y)))
        return dict(success=float(self.n_steps % 2), spl=float(np.float32(0.125 * (e % 5))), oracle_success=float(e % 2),
                    distance_to_goal=d, top_down_map=None)

    def close(self):
        self.closed = True


def fake_sensors():
    """The sim_sensors node of config/vln_r2r.yaml:22-33 (640x480, hfov 79, camera at 1.25 m, depth 0..10 m)."""
    rgb = SimpleNamespace(width=640, height=480, hfov=79, position=[0, 1.25, 0])
    depth = SimpleNamespace(width=640, height=480, hfov=79, position=[0, 1.25, 0], min_depth=0.0, max_depth=10.0)
    return SimpleNamespace(rgb_sensor=rgb, depth_sensor=depth)


class StampProcessor:
    """image_processor stand-in: 'pixel_values' of a frame = [3,2,2] tensor holding (step id, episode id, pixel (0,0,2))."""
    crop_size = {"height": 384, "width": 384}

    def preprocess(self, images, return_tensors="pt"):
        import torch

        a = np.asarray(images)
        v = torch.tensor([float(a[0, 0, 0]), float(a[0, 0, 1]), float(a[0, 0, 2])])
        return {"pixel_values": [v[:, None, None].expand(3, 2, 2).clone()]}


class RecordingModel:
    """Fake model: records every `generate(**kw)` and `reset_for_env`, answers with a scripted token sequence whose decoded
    text yields 0..5 actions (so the loop's action queue, its empty-answer STOP rule and odd cadences are all exercised)."""

    def __init__(self, tokenizer, by_step: bool = False):
        self.tok = tokenizer
        self.by_step = by_step            # answers depend on (episode, step) only: independent of the order of the calls
        self.calls, self.resets, self.batches = [], [], []
        self.tower = SimpleNamespace(image_processor=StampProcessor())
        self.model = SimpleNamespace(num_history=None)
        # StubQwenTokenizer decodes an ordinary id i as the action glyph i % 4 (STOP, forward, left, right)
        self._ids = {a: [1000 + a] for a in range(4)}
        import torch

        self.device = torch.device("cpu")                       # read by the product's evaluator, not by the reference's
        self.config = SimpleNamespace(mode="reference")

    def get_vision_tower(self):
        return self.tower

    def eval(self):
        return self

    def reset(self, n):
        self.resets.append(("reset", n))

    def reset_for_env(self, idx):
        self.resets.append((len(self.calls), idx))

    def reset_episode(self, idx):          # the product agent's per-episode reset; recorded like reset_for_env
        self.resets.append((len(self.calls), idx))

    def script(self, t: int, views) -> List[int]:
        """Actions of call t: usually 4 moves, sometimes 3 / 5 / none (never none in the long episodes 3 and 5, which run
        to their step caps across two window boundaries), STOP after step 12 in even episodes."""
        ep, step = int(views[-1][1]), int(views[-1][0])
        if ep % 2 == 0 and step >= 12:
            return [1, 0]
        r = (t * 5 + ep) % 11
        if r == 3 and ep not in (3, 5):
            return []                                           # no action words in the answer -> the loop stops ([0])
        n = 3 if r == 7 else 5 if r == 9 else 4
        if ep == 5:
            n = 4                                               # model steps land on the window boundaries 32 and 64
        return [((t + j + ep) % 3) + 1 for j in range(n)]

    def generate_batch(self, requests):
        self.batches.append(len(requests))
        return [self.generate(**r) for r in requests]

    def generate(self, **kw):
        import torch

        t = len(self.calls)
        img = kw["images"][0].float()
        views = [[int(round(float(img[v, c, 0, 0]))) for c in range(3)] for v in range(img.shape[0])]
        acts = self.script(views[-1][0] if self.by_step else t, views)
        seq = [i for a in acts for i in self._ids[a]] + [self.tok.im_end]
        self.calls.append(dict(
            inputs=kw["inputs"][0].tolist(), views=views, time_ids=list(kw["time_ids"][0]), env_id=kw["env_id"],
            fresh=kw["past_key_values"] is None, max_new_tokens=kw["max_new_tokens"],
            depths=kw["depths"][0].float().numpy().copy(), poses=kw["poses"][0].float().numpy().copy(),
            intrinsics=kw["intrinsics"][0].float().numpy().copy(),
            dtypes=[str(kw[k].dtype) for k in ("images", "depths", "poses", "intrinsics")]))
        return SimpleNamespace(sequences=torch.tensor([seq]), past_key_values=("kv", t))

"""Habitat evaluation driver on top of the B200 agent: the reference's `VLNEvaluator.eval_action` / `evaluate`
(streamvln/streamvln_eval.py:192-380, 536-581).

Same loop, same results: episodes grouped by scene, `episodes[idx::env_num]` per rank (:218-219), resume from
`result.json` (:205-216), one observation per env step turned into (rgb, depth in metres on the 384 x 384 grid, camera ->
episodic pose, scaled intrinsics) (:247-283), a model turn whenever the action queue is empty (:289-340), the empty-answer
STOP rule (:341-342), the window reset every `num_frames` steps (:346-351), one `result.json` row per episode (:366-379)
and the final all_gather of the four metric vectors (:553-565).  Pinned to the reference's own loop by
tests/golden/ref_eval_loop.npz (scripts/gen_golden_eval_loop.py drives the unmodified reference module against
tests/fake_env.py; tests/test_eval_loop_cpu.py drives this module against the same fake simulator).

What differs, all outside the result:
  * frames are preprocessed lazily - only the views a model call selects ever reach the image / depth processors (the
    reference runs PIL bicubic + the depth path on every env step and uses one in four);
  * with the model on a GPU both processors are the CUDA kernels behind `svln_preprocess_rgb` / `svln_preprocess_depth`
    (bit-exact with the PIL / numpy paths), fed through pinned memory;
  * the tokenizer is prepared once (the reference deep-copies it on every model call, :399);
  * `frame_ids` / `episode_id` accompany every request so window starts reuse the SigLIP features of frames the model
    has already seen; depth / pose / intrinsics stay fp32 (the reference casts them to bf16 and then ignores them);
  * the simulator is injected (`env_factory`): Habitat itself is imported only when no factory is given.
With a `mode="slowfast"` model the loop keeps one live context per episode (memory sentence from the first turn, no window
reset, no history views), exactly as `streamvln_b200.agent` does for the real-world agent.
Video / top-down-map dumps (`save_video`) are not reproduced.
"""
from __future__ import annotations

import copy
import json
import os
import random
from types import SimpleNamespace
from typing import Any, Callable, List, Optional

import numpy as np
import torch

from .agent import DEFAULT_MEMORY_TOKEN, DEFAULT_VIDEO_TOKEN, VLNEvaluator as _Agent
from .synth import AXIS_ALIGN, intrinsic_matrix, preprocess_depth


def camera_to_episodic(xyz, yaw) -> np.ndarray:
    """streamvln_eval.py:175-185 (`xyz_yaw_to_tf_matrix`): rotation about z by `yaw`, translation xyz.  cos / sin are taken
    in the dtype the simulator delivers the compass in (float32 from Habitat), as the reference does."""
    x, y, z = xyz
    return np.array([[np.cos(yaw), -np.sin(yaw), 0, x], [np.sin(yaw), np.cos(yaw), 0, y], [0, 0, 1, z], [0, 0, 0, 1]])


def _load_filter_depth():
    try:
        from depth_camera_filtering import filter_depth     # requirements.txt:26 (depth-camera-filtering@39d6e2f)

        return filter_depth, "depth_camera_filtering.filter_depth"
    except ImportError:
        return (lambda d, blur_type=None: d), "identity (depth_camera_filtering is not installed)"


def _read_result_rows(path: str):
    """Episode rows of a result.json (one JSON object per line).  `evaluate` appends its summary WITHOUT a newline
    (streamvln_eval.py:580-581), so the first row of a resumed run shares a line with it: every object on a line is read,
    summary objects (no `scene_id`) are skipped."""
    rows, dec = [], json.JSONDecoder()
    if not os.path.exists(path):
        return rows
    with open(path) as f:
        for line in f:
            i, line = 0, line.strip()
            while i < len(line):
                obj, j = dec.raw_decode(line, i)
                i = j
                while i < len(line) and line[i].isspace():
                    i += 1
                if "scene_id" in obj:
                    rows.append(obj)
    return rows


class VLNEvaluator(_Agent):
    """Drop-in for the Habitat evaluator class of streamvln/streamvln_eval.py:44-380 (same constructor keywords; the
    simulator and its sensor config can be injected)."""

    def __init__(self, config_path: Optional[str] = None, split: str = "val_seen", env_num: int = 8, output_path: str = None,
                 model: Any = None, tokenizer: Any = None, epoch: int = 0, args=None, *,
                 env_factory: Optional[Callable[[], Any]] = None, sim_sensors=None,
                 filter_depth: Optional[Callable] = None, rank: Optional[int] = None, episode_seed=None):
        self.episode_seed, self._rng = episode_seed, None
        self.split, self.env_num, self.output_path, self.epoch, self.config_path = split, env_num, output_path, epoch, config_path or ""
        if env_factory is None:
            env_factory, sim_sensors = self._habitat_factory(config_path, split)
        self.env_factory, self.sim_sensors = env_factory, sim_sensors
        rgb, dep = sim_sensors.rgb_sensor, sim_sensors.depth_sensor
        self._camera_height = rgb.position[1]
        self._min_depth, self._max_depth = dep.min_depth, dep.max_depth
        self._sensor_wh = (rgb.width, rgb.height)
        if not hasattr(args, "gpu_preprocess"):
            args = SimpleNamespace(**vars(args), gpu_preprocess=getattr(getattr(model, "device", None), "type", "cpu") == "cuda")
        size = model.get_vision_tower().image_processor.crop_size["height"]
        K = intrinsic_matrix(rgb.width, rgb.height, rgb.hfov, size)           # :137-168 (get_intrinsic_matrix + preprocess_instrinsic)
        super().__init__({"rgb_height": self._camera_height, "camera_intrinsic": K}, model=model, tokenizer=tokenizer, args=args)
        self.save_video = getattr(args, "save_video", False)
        self.device = getattr(model, "device", None) or torch.device("cpu")
        self.filter_depth, self.depth_filter_name = (filter_depth, "injected") if filter_depth is not None else _load_filter_depth()
        self._rank = rank
        self.gpu_depth = None
        if self.gpu_preprocess:
            from .preprocess import GpuDepthProcessor

            self.gpu_depth = GpuDepthProcessor(self.device, size, self._min_depth, self._max_depth)

    # ---- pieces of the loop -----------------------------------------------------------------------------------------
    @staticmethod
    def _habitat_factory(config_path: str, split: str):
        """The real simulator (:60-90, 186-189).  Habitat is not installed in the build image; this is the only place that
        imports it."""
        import habitat
        from habitat_baselines.config.default import get_config as get_habitat_config

        config = get_habitat_config(config_path)
        with habitat.config.read_write(config):
            config.habitat.dataset.split = split
        return (lambda: habitat.Env(config=config)), config.habitat.simulator.agents.main_agent.sim_sensors

    def _conjunction(self) -> str:
        return (self._rng or random).choice(self.conjunctions)              # :424 (the reference draws from the global RNG)

    def _get_rank(self) -> int:
        if self._rank is not None:
            return self._rank
        import torch.distributed as dist

        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    def _image(self, rgb: np.ndarray) -> torch.Tensor:
        if self.gpu_preprocess:
            u8 = torch.from_numpy(np.ascontiguousarray(np.asarray(rgb)[..., :3]))
            if self._pin is None or self._pin.shape != u8.shape:
                self._pin = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
            self._pin.copy_(u8)
            out = self.gpu_proc(self._pin.to(self.device, non_blocking=True)[None])[0]
            torch.cuda.current_stream().synchronize()       # the pinned staging buffer is reused by the next frame
            return out
        from PIL import Image

        return self.image_processor.preprocess(images=Image.fromarray(rgb).convert("RGB"), return_tensors="pt")["pixel_values"][0]

    def _depth(self, depth01: np.ndarray) -> torch.Tensor:
        """:254-256, 276, 126-135: filter, [0,1] -> metres -> millimetres (uint16) -> NEAREST to the processor grid -> metres."""
        d = self.filter_depth(depth01.reshape(depth01.shape[:2]), blur_type=None)
        if self.gpu_depth is not None:
            return self.gpu_depth(torch.from_numpy(np.ascontiguousarray(d, np.float32)))[0]
        return torch.from_numpy(preprocess_depth(d, self._min_depth, self._max_depth, self.image_processor.crop_size["height"])).float()

    # ---- the loop ---------------------------------------------------------------------------------------------------
    def eval_action(self, idx: int):
        env = self.env_factory()
        scene_episode_dict = {}
        for episode in env.episodes:
            scene_episode_dict.setdefault(episode.scene_id, []).append(episode)
        sucs, spls, oss, ones = [], [], [], []
        done_res = []
        result_path = os.path.join(self.output_path, "result.json")
        for res in _read_result_rows(result_path):                          # resume (:205-216)
            done_res.append([res["scene_id"], res["episode_id"], res["episode_instruction"]])
            if self._get_rank() == 0:
                sucs.append(res["success"]); spls.append(res["spl"]); oss.append(res["os"]); ones.append(res["ne"])
        os.makedirs(self.output_path, exist_ok=True)
        self._finished: List[tuple] = []
        for scene in sorted(scene_episode_dict.keys()):
            episodes = scene_episode_dict[scene]
            scene_id = scene.split("/")[-2]
            for episode in episodes[idx::self.env_num]:                     # :218-219
                episode_instruction = episode.instruction.instruction_text if "objectnav" not in self.config_path else episode.object_category
                episode_id = episode.episode_id
                if [scene_id, episode_id, episode_instruction] in done_res:
                    continue
                metrics, step_id = self._run_episode(env, episode, idx)
                self._finished.append((scene_id, episode_id))
                sucs.append(metrics["success"]); spls.append(metrics["spl"])
                oss.append(metrics["oracle_success"]); ones.append(metrics["distance_to_goal"])
                result = {"scene_id": scene_id, "episode_id": episode_id, "success": metrics["success"], "spl": metrics["spl"],
                          "os": metrics["oracle_success"], "ne": metrics["distance_to_goal"], "steps": step_id,
                          "episode_instruction": episode_instruction}
                with open(result_path, "a") as f:
                    f.write(json.dumps(result) + "\n")
        env.close()
        dev = self.device
        return (torch.tensor(sucs).to(dev), torch.tensor(spls).to(dev), torch.tensor(oss).to(dev), torch.tensor(ones).to(dev),
                torch.tensor(len(sucs)).to(dev))

    def _run_episode(self, env, episode, idx: int):
        """One episode, one model call at a time (the reference's schedule)."""
        gen = self._episode(env, episode, idx)
        req = next(gen)
        try:
            while True:
                req = gen.send(self.model.generate(**req))
        except StopIteration as done:
            return done.value

    def _episode(self, env, episode, idx: int):
        """One episode (:225-364) as a generator: yields the kwargs of each `model.generate` call and is sent its output, so
        that several episodes can be in flight on one model replica (`ConcurrentEvaluator`).  Raw observations are buffered;
        a frame is preprocessed the first time a model call selects it."""
        self._init_state()
        if self.episode_seed is not None:          # per-episode prompt RNG: the schedule of other episodes cannot shift it
            self._rng = random.Random(f"{self.episode_seed}/{episode.scene_id}/{episode.episode_id}")
        self.model.reset_episode(idx) if hasattr(self.model, "reset_episode") else self.model.reset_for_env(idx)
        env.current_episode = episode
        observations = env.reset()
        initial_height = env.sim.get_agent_state().position[1]
        raw_rgb, raw_depth, done_rgb, done_depth = [], [], {}, {}
        pose_list, time_ids, action_seq = [], [], []
        past_key_values, output_ids = None, None
        step_id = 0
        K = torch.from_numpy(np.asarray(self.intrinsic_matrix)).float()
        align = AXIS_ALIGN
        while not env.episode_over:
            time_ids.append(step_id)
            raw_rgb.append(observations["rgb"])
            raw_depth.append(observations["depth"])
            x, y = observations["gps"]
            camera_yaw = observations["compass"][0]
            height = env.sim.get_agent_state().position[1] - initial_height
            camera_position = np.array([x, -y, self._camera_height + height])   # Habitat GPS makes west negative (:262)
            pose_list.append(camera_to_episodic(camera_position, camera_yaw) @ align)
            if len(action_seq) == 0:
                if output_ids is None:
                    sources = copy.deepcopy(self.conversation)
                    if step_id != 0 or self.slowfast:
                        sources[0]["value"] += f" These are your historical observations {DEFAULT_MEMORY_TOKEN}."
                    sources[0]["value"] = sources[0]["value"].replace(DEFAULT_VIDEO_TOKEN + "\n", "")
                    sources[0]["value"] = sources[0]["value"].replace("<instruction>.", episode.instruction.instruction_text)
                    add_system = True
                else:
                    sources = [{"from": "human", "value": ""}, {"from": "gpt", "value": ""}]
                    add_system = False
                input_ids, _ = self.preprocess_qwen([sources], self.tokenizer, True, add_system=add_system)
                if output_ids is not None:
                    input_ids = torch.cat([output_ids.cpu(), input_ids], dim=1)
                sel = [step_id]
                if not self.slowfast and step_id != 0 and step_id % self.num_frames == 0:           # :312-321
                    if self.num_history is None:
                        history_ids = slice(0, time_ids[0], self.num_future_steps)
                    else:
                        history_ids = slice(0, time_ids[0], (time_ids[0] // self.num_history))
                    sel = list(range(step_id + 1))[history_ids] + sel
                for i in sel:
                    if i not in done_rgb:
                        done_rgb[i] = self._image(raw_rgb[i])
                        done_depth[i] = self._depth(raw_depth[i])
                images = torch.stack([done_rgb[i] for i in sel]).unsqueeze(0)
                depths = torch.stack([done_depth[i] for i in sel]).unsqueeze(0)
                poses = torch.from_numpy(np.stack([pose_list[i] for i in sel])).float().unsqueeze(0)
                intrinsics = torch.stack([K for _ in sel]).unsqueeze(0)
                if not images.is_cuda and self.device.type == "cuda":
                    images = images.pin_memory().to(self.device, non_blocking=True)
                outputs = yield dict(inputs=input_ids, images=images, depths=depths, poses=poses, intrinsics=intrinsics,
                                     env_id=idx, time_ids=[list(time_ids)], task_type=[0], frame_ids=list(sel),
                                     episode_id=self.episode_id, do_sample=False, num_beams=1,
                                     max_new_tokens=self.max_new_tokens, use_cache=True, return_dict_in_generate=True,
                                     past_key_values=past_key_values)
                self.past_key_values = None
                action_seq, _ = self.finish(outputs)                          # keeps ids, decodes, parses; [] -> [0] (:341-342)
                output_ids, past_key_values = self.output_ids, outputs.past_key_values
            action = action_seq.pop(0)
            observations = env.step(action)
            step_id += 1
            if not self.slowfast and step_id % self.num_frames == 0:          # :346-351 (the slowfast mode slides its window instead)
                self.model.reset_for_env(idx)
                output_ids, past_key_values, time_ids = None, None, []
        return env.get_metrics(), step_id


class ConcurrentEvaluator:
    """`eval_action` with E simulators per rank: the rank's episodes (`episodes[idx::env_num]` of every scene, in the
    reference's order) are dealt to E slots, each slot walks its episode with its own simulator and agent state, and the
    model turns that are due at the same time go through ONE packed `model.generate_batch` - on a B200 64 concurrent
    streams cost ~3x less GPU time per env step than one (profiles/README.md, `batched`).  Per episode the calls, the
    actions and the result row are those of the sequential loop (with `episode_seed` the prompt RNG is per episode, so the
    prompts match too); result.json rows appear in completion order."""

    def __init__(self, n_slots: int, **evaluator_kw):
        self.n_slots = n_slots
        evaluator_kw.setdefault("episode_seed", 0)
        self.slots = [VLNEvaluator(**evaluator_kw) for _ in range(n_slots)]
        self.model = evaluator_kw["model"]
        self.batch_sizes: List[int] = []

    def eval_action(self, idx: int):
        lead = self.slots[0]
        envs = [ev.env_factory() for ev in self.slots]
        scene_episode_dict = {}
        for episode in envs[0].episodes:
            scene_episode_dict.setdefault(episode.scene_id, []).append(episode)
        sucs, spls, oss, ones, done_res = [], [], [], [], []
        result_path = os.path.join(lead.output_path, "result.json")
        for res in _read_result_rows(result_path):
            done_res.append([res["scene_id"], res["episode_id"], res["episode_instruction"]])
            if lead._get_rank() == 0:
                sucs.append(res["success"]); spls.append(res["spl"]); oss.append(res["os"]); ones.append(res["ne"])
        os.makedirs(lead.output_path, exist_ok=True)
        queue = []
        for scene in sorted(scene_episode_dict.keys()):
            for episode in scene_episode_dict[scene][idx::lead.env_num]:
                text = episode.instruction.instruction_text if "objectnav" not in lead.config_path else episode.object_category
                if [scene.split("/")[-2], episode.episode_id, text] not in done_res:
                    queue.append((scene.split("/")[-2], episode, text))
        live = {}                                  # slot -> (generator, pending request, scene_id, episode, text)

        def start(j):
            while queue:
                scene_id, episode, text = queue.pop(0)
                gen = self.slots[j]._episode(envs[j], episode, j)
                try:
                    live[j] = (gen, next(gen), scene_id, episode, text)
                    return
                except StopIteration as done:          # an episode that is over before its first model turn
                    finish(j, scene_id, episode, text, done.value)

        def finish(j, scene_id, episode, text, value):
            metrics, step_id = value
            sucs.append(metrics["success"]); spls.append(metrics["spl"])
            oss.append(metrics["oracle_success"]); ones.append(metrics["distance_to_goal"])
            row = {"scene_id": scene_id, "episode_id": episode.episode_id, "success": metrics["success"], "spl": metrics["spl"],
                   "os": metrics["oracle_success"], "ne": metrics["distance_to_goal"], "steps": step_id, "episode_instruction": text}
            with open(result_path, "a") as f:
                f.write(json.dumps(row) + "\n")

        for j in range(self.n_slots):
            start(j)
        while live:
            order = sorted(live)
            reqs = [live[j][1] for j in order]
            outs = self.model.generate_batch(reqs) if len(reqs) > 1 or not hasattr(self.model, "generate") \
                else [self.model.generate(**reqs[0])]
            self.batch_sizes.append(len(reqs))
            for j, out in zip(order, outs):
                gen, _, scene_id, episode, text = live[j]
                try:
                    live[j] = (gen, gen.send(out), scene_id, episode, text)
                except StopIteration as done:
                    del live[j]
                    finish(j, scene_id, episode, text, done.value)
                    start(j)
        for env in envs:
            env.close()
        dev = lead.device
        return (torch.tensor(sucs).to(dev), torch.tensor(spls).to(dev), torch.tensor(oss).to(dev), torch.tensor(ones).to(dev),
                torch.tensor(len(sucs)).to(dev))


def evaluate(model, tokenizer, args, **evaluator_kw):
    """streamvln_eval.py:536-581: one evaluator per rank, `eval_action(rank)`, all_gather of the episode counts and of the
    four metric vectors, the averaged summary appended to result.json by rank 0.  Returns the summary dict."""
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    world_size, rank = (dist.get_world_size(), dist.get_rank()) if on else (1, 0)
    model.eval()
    n_concurrent = int(getattr(args, "concurrent_episodes", 1) or 1)      # > 1: E simulators per rank, packed model turns
    kw = dict(config_path=getattr(args, "habitat_config_path", None), split=getattr(args, "eval_split", "val_unseen"),
              env_num=world_size, output_path=args.output_path, model=model, tokenizer=tokenizer, epoch=0, args=args, **evaluator_kw)
    if n_concurrent > 1:
        model.reset(n_concurrent)
        evaluator = ConcurrentEvaluator(n_concurrent, **kw)
    else:
        model.reset(world_size)
        evaluator = VLNEvaluator(**kw)
    sucs, spls, oss, ones, ep_num = evaluator.eval_action(rank)
    vecs = [v.to(torch.float64) for v in (sucs, spls, oss, ones)]
    if on:
        ep_num_all = [torch.zeros_like(ep_num) for _ in range(world_size)]
        dist.all_gather(ep_num_all, ep_num)
        mx = max(int(n) for n in ep_num_all)
        gathered = []
        dist.barrier()
        for v in vecs:                                       # padded rows: NCCL/gloo all_gather wants equal shapes
            pad = torch.zeros(max(mx, 1), dtype=v.dtype, device=v.device)
            pad[: v.numel()] = v
            parts = [torch.zeros_like(pad) for _ in range(world_size)]
            dist.all_gather(parts, pad)
            gathered.append(torch.cat([p[: int(n)] for p, n in zip(parts, ep_num_all)]))
        dist.barrier()
        vecs = gathered
    n = vecs[0].numel()
    result_all = {"sucs_all": (vecs[0].sum() / n).item(), "spls_all": (vecs[1].sum() / n).item(),
                  "oss_all": (vecs[2].sum() / n).item(), "ones_all": (vecs[3].sum() / n).item(), "length": n}
    if rank == 0:
        with open(os.path.join(args.output_path, "result.json"), "a") as f:
            f.write(json.dumps(result_all))
    return result_all

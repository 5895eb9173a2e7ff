"""Simulator-free streaming agent with the reference's `VLNEvaluator.step` API.

Mirrors streamvln/streamvln_agent.py:31-258 (class `VLNEvaluator`: frame buffers, window reset every
`num_frames` steps, history sampling `slice(0, time_ids[0], time_ids[0] // num_history)`, prompt build
with the overridden Qwen chat template, greedy generate, regex action parse) on top of
`streamvln_b200.model.StreamVLNForCausalLM`.  Differences, all outside the result:
  * depth / pose / intrinsics stay fp32 on their way to the model (the reference casts them to bf16,
    streamvln_agent.py:238-240, and then ignores them);
  * frames travel through pinned host memory with one async H2D per turn;
  * `frame_ids` is passed so a window start does not re-encode its history frames;
  * real depth / pose can be supplied (`depth=`, `pose=`) instead of the faked zeros / identity
    (streamvln_agent.py:172-174), which is what the slow-memory voxel pruning consumes.
"""
from __future__ import annotations

import copy
import itertools
import re
import time
from collections import OrderedDict
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .model import IMAGE_TOKEN_INDEX, MEMORY_TOKEN_INDEX, StreamVLNForCausalLM

DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_MEMORY_TOKEN = "<memory>"
DEFAULT_VIDEO_TOKEN = "<video>"
_EPISODE_NONCE = itertools.count(1)   # one per episode and process: tags the model's per-frame feature cache


class VLNEvaluator:
    def __init__(self, sim_sensors_config: Dict, model: StreamVLNForCausalLM = None, tokenizer: Any = None, args=None):
        self.args = args
        self.device = model.device
        self.sim_sensors_config = sim_sensors_config
        self.intrinsic_matrix = np.asarray(sim_sensors_config["camera_intrinsic"])
        self.image_processor = model.get_vision_tower().image_processor
        self.model = model
        self.tokenizer = tokenizer
        self.tokenizer.add_tokens(["<image>"], special_tokens=True)
        self.tokenizer.add_tokens(["<memory>"], special_tokens=True)
        prompt = ("<video>\nYou are an autonomous navigation assistant. Your task is to <instruction>. Devise an action "
                  "sequence to follow the instruction using the four actions: TURN LEFT (←) or TURN RIGHT (→) by 15 "
                  "degrees, MOVE FORWARD (↑) by 25 centimeters, or STOP.")
        self.conversation = [{"from": "human", "value": prompt}, {"from": "gpt", "value": ""}]
        self.actions2idx = OrderedDict({"STOP": [0], "↑": [1], "←": [2], "→": [3]})
        self.conjunctions = ["you can see ", "in front of you is ", "there is ", "you can spot ",
                             "you are toward the ", "ahead of you is ", "in your sight is "]
        self.use_memory_tokens = True
        self.num_frames = args.num_frames
        self.num_future_steps = args.num_future_steps
        self.num_history = args.num_history
        self.max_new_tokens = getattr(args, "max_new_tokens", 10000)  # streamvln_agent.py:244
        self.slowfast = model.config.mode == "slowfast"
        # the image processor runs on the GPU (bit-exact with the PIL path); args.gpu_preprocess=False keeps PIL
        self.gpu_preprocess = getattr(args, "gpu_preprocess", True)
        if self.gpu_preprocess:
            from .preprocess import GpuImageProcessor

            self.gpu_proc = GpuImageProcessor(self.device, self.image_processor.crop_size["height"])
        self._pin = None
        self._regex = re.compile("|".join(re.escape(a) for a in self.actions2idx))
        self._init_state()

    def _init_state(self):
        self.rgb_list: List[torch.Tensor] = []
        self.depth_list: List[torch.Tensor] = []
        self.pose_list: List[torch.Tensor] = []
        self.intrinsic_list: List[torch.Tensor] = []
        self.time_ids: List[int] = []
        self.action_seq: List[int] = []
        self.output_ids = None
        self.past_key_values = None
        self.step_id = 0
        self.last_image = None
        self.episode_id = next(_EPISODE_NONCE)

    def reset_memory(self):
        """streamvln_agent.py:87-99 (the reference hard-codes env 0; this resets the stream the agent last stepped)."""
        self._init_state()
        env = getattr(self, "_env", 0)
        self.model.reset_episode(env) if hasattr(self.model, "reset_episode") else self.model.reset_for_env(env)

    def _conjunction(self) -> str:
        """The phrase in front of the turn's <image> token.  The real-world agent fixes it (benchmark config: conjunctions[0]);
        the Habitat evaluator draws it with random.choice (streamvln_eval.py:424) and overrides this."""
        return self.conjunctions[0]

    def parse_actions(self, output: str) -> List[int]:
        """streamvln_agent.py:101-107"""
        matches = self._regex.findall(output)
        return list(itertools.chain.from_iterable(self.actions2idx[m] for m in matches))

    def preprocess_qwen(self, sources, tokenizer, has_image: bool = False, max_len=2048,
                        system_message: str = "You are a helpful assistant.", add_system: bool = False):
        """streamvln_agent.py:109-167: every message (including the empty assistant one) is rendered with the
        template '<|im_start|>role\\ncontent<|im_end|>\\n'; <image>/<memory> ids become -200/-300."""
        roles = {"human": "user", "gpt": "assistant"}
        image_token_index = tokenizer.convert_tokens_to_ids("<image>")
        memory_token_index = tokenizer.convert_tokens_to_ids("<memory>")
        tokenizer.chat_template = ("{% for message in messages %}{{'<|im_start|>' + message['role'] + '\n' + message['content'] "
                                   "+ '<|im_end|>' + '\n'}}{% endfor %}{% if add_generation_prompt %}{{ '<|im_start|>assistant\n' }}{% endif %}")
        # the steady-turn prompt is the same text every turn: tokenise a given conversation once (0.2 ms of host time per
        # turn sits between the previous turn's ids and this turn's first launch); the in-place edit of the caller's
        # messages is kept
        memo = self.__dict__.setdefault("_prompt_memo", {})
        picks = [self._conjunction() for _ in sources]
        key = (id(tokenizer), add_system, system_message, tuple(picks),
               tuple(tuple((m.get("from", m.get("role")), m.get("value", m.get("content"))) for m in src) for src in sources))
        hit = memo.get(key)
        conversations, input_ids = [], []
        for source, pick in zip(sources, picks):
            prompt = pick + DEFAULT_IMAGE_TOKEN
            if len(source[0]["value"]) != 0:
                source[0]["value"] += f" {prompt}."
            else:
                source[0]["value"] = f"{prompt}."
            if hit is not None:
                continue
            if roles[source[0]["from"]] != roles["human"]:
                source = source[1:]
            input_id: List[int] = []
            if add_system:
                input_id += tokenizer.apply_chat_template([{"role": "system", "content": system_message}])
            for conv in source:
                role = roles.get(conv.get("role", conv.get("from")), conv.get("role", conv.get("from")))
                content = conv.get("content", conv.get("value"))
                conversations.append(content)
                input_id += tokenizer.apply_chat_template([{"role": role, "content": content}])
            input_id = [IMAGE_TOKEN_INDEX if t == image_token_index else MEMORY_TOKEN_INDEX if t == memory_token_index else t
                        for t in input_id]
            input_ids.append(input_id)
        if hit is not None:
            return hit[0].clone(), list(hit[1])
        out = torch.tensor(input_ids, dtype=torch.long)
        if len(memo) < 256:
            memo[key] = (out.clone(), list(conversations))
        return out, conversations

    # ------------------------------------------------------------------------------------------
    def prepare(self, idx, rgb, instruction_text: str = "", run_model: bool = False, depth=None, pose=None, intrinsic=None,
                preprocessed: Optional[torch.Tensor] = None):
        """First half of `step` (streamvln_agent.py:169-240): buffer the frame, handle the window boundary, and on
        a model step build the kwargs of `model.generate`.  Returns None on a non-model step."""
        self._env = idx
        # Step 0: observations (faked like the reference unless the caller supplies real geometry)
        camera_pose = np.eye(4) if pose is None else np.asarray(pose)
        depth = np.zeros((self.image_processor.crop_size["height"], self.image_processor.crop_size["width"]), np.float32) \
            if depth is None else np.asarray(depth, np.float32)
        intrinsic = torch.from_numpy(np.asarray(self.intrinsic_matrix if intrinsic is None else intrinsic)).float()
        # Step 1: preprocess the image only on model steps; other steps repeat the last one (:177-184)
        if run_model:
            self.h2d_bytes = 0
            if preprocessed is not None:
                image = preprocessed
            elif self.gpu_preprocess:
                # uint8 frame -> pinned staging -> device; PIL-exact bicubic + normalise run as CUDA kernels
                u8 = torch.from_numpy(np.ascontiguousarray(np.asarray(rgb)[..., :3]))
                if self._pin is None or self._pin.shape != u8.shape:
                    self._pin = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
                self._pin.copy_(u8)
                image = self.gpu_proc(self._pin.to(self.device, non_blocking=True)[None])[0]
                self.h2d_bytes = u8.numel()
                self.model.engine.launches += 2
            else:
                from PIL import Image

                image = self.image_processor.preprocess(images=Image.fromarray(rgb).convert("RGB"), return_tensors="pt")["pixel_values"][0]
            self.last_image = image
        else:
            image = self.last_image
        self.time_ids.append(self.step_id)
        self.rgb_list.append(image)
        self.depth_list.append(torch.from_numpy(depth).float())
        self.pose_list.append(torch.from_numpy(camera_pose).float())
        self.intrinsic_list.append(intrinsic)
        # Step 2: window boundary on a non-model step (:191-199); the slowfast mode slides instead of resetting
        if not run_model:
            if not self.slowfast and self.use_memory_tokens and (self.step_id + 1) % self.num_frames == 0:
                self.model.reset_for_env(idx)
                self.output_ids = None
                self.past_key_values = None
                self.time_ids = []
            return None
        # Step 3: prompt (:201-217)
        if self.output_ids is None:
            sources = copy.deepcopy(self.conversation)
            sources[0]["value"] = sources[0]["value"].replace(
                " Where should you go next to stay on track?",
                " Please devise an action sequence to follow the instruction which may include turning left or right by a "
                "certain degree, moving forward by a certain distance or stopping once the task is complete.")
            if self.step_id != 0 or self.slowfast:
                sources[0]["value"] += f" You have visited these areas {DEFAULT_MEMORY_TOKEN}."
            sources[0]["value"] = sources[0]["value"].replace(DEFAULT_VIDEO_TOKEN + "\n", "")
            sources[0]["value"] = sources[0]["value"].replace("<instruction>.", instruction_text)
            add_system = True
        else:
            sources = [{"from": "human", "value": ""}, {"from": "gpt", "value": ""}]
            add_system = False
        input_ids, _ = self.preprocess_qwen([sources], self.tokenizer, True, add_system=add_system)
        if self.output_ids is not None:
            input_ids = torch.cat([self.output_ids.cpu(), input_ids], dim=1)
        sel = [len(self.rgb_list) - 1]
        if self.use_memory_tokens and not self.slowfast:
            if self.step_id != 0 and self.step_id % self.num_frames == 0:  # (:222-232)
                if self.num_history is None:
                    history_ids = slice(0, self.time_ids[0], self.num_future_steps)
                else:
                    history_ids = slice(0, self.time_ids[0], (self.time_ids[0] // self.num_history))
                sel = list(range(len(self.rgb_list)))[history_ids] + sel
        images = torch.stack([self.rgb_list[i] for i in sel]).unsqueeze(0)
        depths = torch.stack([self.depth_list[i] for i in sel]).unsqueeze(0)
        poses = torch.stack([self.pose_list[i] for i in sel]).unsqueeze(0)
        intrinsics = torch.stack([self.intrinsic_list[i] for i in sel]).unsqueeze(0)
        # frames repeated on non-model steps are the same image: key the feature cache by the model-step frame
        frame_ids = [self._source_frame(i) for i in sel]
        # Step 4: generate (:242-246).  One pinned-memory H2D per turn for the frames.
        if not images.is_cuda:
            self.h2d_bytes = images.numel() * images.element_size()
            if self.device.type == "cuda":   # (a CPU `device` only occurs in the host-logic tests, with a recording model)
                images = images.pin_memory().to(self.device, non_blocking=True)
        self.h2d_bytes = getattr(self, "h2d_bytes", 0) + input_ids.numel() * 8
        return dict(inputs=input_ids, images=images, depths=depths, poses=poses, intrinsics=intrinsics, env_id=idx,
                    time_ids=[list(self.time_ids)], task_ids=[0], frame_ids=frame_ids, episode_id=self.episode_id, do_sample=False, num_beams=1,
                    max_new_tokens=self.max_new_tokens, use_cache=True, return_dict_in_generate=True,
                    past_key_values=self.past_key_values)

    def finish(self, outputs):
        """Second half of `step` (:247-258): keep ids + cache handle, decode, parse the actions."""
        self.output_ids = outputs.sequences
        ids_host = self.output_ids.cpu()  # the D2H read of the turn's result
        self.d2h_bytes = ids_host.numel() * 8
        if ids_host.numel() and int(ids_host.min()) < 0:
            # svln_decode_stack_step reports token -1 when one of its grid barriers timed out (include/svln.h)
            raise RuntimeError("libsvln: the fused decode stack aborted (grid barrier timeout); "
                               "svln_decode_stack_status() has the barrier index")
        llm_outputs = self.tokenizer.batch_decode(ids_host, skip_special_tokens=False)[0].strip()
        self.past_key_values = outputs.past_key_values
        action_seq = self.parse_actions(llm_outputs)
        if len(action_seq) == 0:
            action_seq = [0]
        return action_seq, llm_outputs

    def step(self, idx, rgb, instruction_text: str = "", run_model: bool = False, depth=None, pose=None, intrinsic=None,
             preprocessed: Optional[torch.Tensor] = None):
        """streamvln_agent.py:169-258.  Returns (action_seq | None, generate_time, llm_outputs | None).
        `preprocessed` (optional [3,384,384] tensor, host or device) bypasses the PIL image processor — used
        when frames are already resident in HBM."""
        req = self.prepare(idx, rgb, instruction_text, run_model, depth, pose, intrinsic, preprocessed)
        if req is None:
            return None, 0, None
        tt = time.time()
        outputs = self.model.generate(**req)
        action_seq, llm_outputs = self.finish(outputs)
        return action_seq, time.time() - tt, llm_outputs

    def _source_frame(self, i: int) -> int:
        """Index of the model-step frame whose preprocessed image frame i repeats (identity lookup)."""
        img = self.rgb_list[i]
        j = i
        while j > 0 and self.rgb_list[j - 1] is img:
            j -= 1
        return j

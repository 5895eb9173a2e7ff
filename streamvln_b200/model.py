"""`StreamVLNForCausalLM`-compatible model object over the B200 kernels.

Drop-in for streamvln/model/stream_video_vln.py at its three call sites (streamvln_agent.py:244,
streamvln_eval.py:332, streamvln_dagger.py:307): `generate(**kw)` with the same kwargs, returning an
object with `.sequences` (new token ids only) and `.past_key_values` (opaque handle); `reset(env_num)`,
`reset_for_env(env_idx)`; `.model.num_history`; `.get_vision_tower().image_processor`; `.eval()`,
`.requires_grad_(False)`, `.to(device)`.

Two semantic modes:
  * mode="reference": the released code's behaviour — the KV cache grows inside a window and is dropped
    with the window; memory = the first `num_history` views pooled to 196 tokens each
    (stream_video_vln.py:102-142).  Token stream, positions and greedy outputs match the CPU oracle.
  * mode="slowfast": the north-star semantics — voxel-pruned slow memory spliced into a live,
    turn-granular sliding KV window with in-place compaction and RoPE re-rotation.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import check, cur_stream, ptr
from .qwen2 import Chunk, Qwen2Config, Qwen2Engine, Qwen2Weights
from .vision import VisionConfig, VisionFrontEnd
from .voxel import INVALID_KEEP, VoxelPruneConfig, VoxelPruner

IGNORE_INDEX = -100          # streamvln/utils/utils.py:8-16
IMAGE_TOKEN_INDEX = -200
MEMORY_TOKEN_INDEX = -300


@dataclass
class StreamVLNConfig:
    llm: Qwen2Config = field(default_factory=Qwen2Config.qwen2_7b)
    vision: VisionConfig = field(default_factory=VisionConfig)
    num_history: Optional[int] = 8
    eos_token_id: Optional[int] = None       # <|im_end|> with a real tokenizer; None = run to max_new_tokens
    tokenizer_model_max_length: Optional[int] = None
    mode: str = "reference"                  # "reference" | "slowfast"
    memory: str = "sampled"                  # reference mode, what fills <memory> at a window start: "sampled" = the reference's
                                             # num_history uniformly sampled frames; "voxel" = every past model-step frame,
                                             # voxel-pruned (keep-latest-per-voxel, stable compaction, row gather: svln_voxel_prune)
    # slowfast only
    window_turns: int = 8                    # num_frames / num_future_steps
    voxel_size: float = 0.2
    max_depth: float = 10.0
    invalid_policy: int = INVALID_KEEP
    memory_cap: int = 1568                   # max memory tokens kept (oldest dropped first)
    feat_cache_frames: Optional[int] = None  # per-stream LRU bound of the pooled-feature cache; None = from the byte budget
    feat_cache_bytes: int = 8 << 30          # ... shared by all streams (1.4 MB per frame at 7B: 1 stream 1024 frames,
                                             # 64 streams 89 frames each = 356 env steps; a miss only costs a re-encode)

    @staticmethod
    def tiny(vocab_size: int = 512, mode: str = "reference", layers: int = 2, max_pos: int = 4096) -> "StreamVLNConfig":
        llm = Qwen2Config.tiny(vocab_size, layers, max_pos)
        return StreamVLNConfig(llm=llm, vision=VisionConfig.tiny(llm.hidden_size), mode=mode)


def uncached_tail(total_len: int, past_len: int):
    """Rows of the stream's embed cache that one `generate` call runs through the LLM: (first_row, n_rows).
    The reference passes the FULL embed cache to HF generate together with the carried KV cache; HF 4.45.1 sets
    cache_position = arange(total_len)[past_len:] and the reference's prepare_inputs_for_generation override keeps
    `inputs_embeds[:, -len(cache_position):]` (all of them when cache_position[0] == 0)
    (streamvln/model/stream_video_vln.py:449-453; pinned by tests/golden/ref_tail_rule.npz)."""
    n = total_len - past_len
    if n <= 0:
        raise ValueError("nothing to run: the KV cache already covers every embedded token")
    return past_len, n


class KvHandle:
    """Opaque `past_key_values`: the arena owns the data; the handle only names the stream.  The length is read
    from the arena on demand (in slowfast mode a device-side eviction may still owe it to the host)."""

    def __init__(self, env: int, arena):
        self.env, self._arena, self._length = env, arena, None

    @property
    def length(self) -> int:
        if self._length is None:
            self._length = self._arena.lens[self.env]
        return self._length

    def get_seq_length(self) -> int:
        return self.length

    def __repr__(self):
        return f"KvHandle(env={self.env})"


class GenerateOutput(SimpleNamespace):
    pass


class _Inner:
    """`.model` of the reference object: only `num_history` is touched by callers (streamvln_eval.py:531)."""

    def __init__(self, num_history):
        self.num_history = num_history


class StreamVLNForCausalLM:
    def __init__(self, config: StreamVLNConfig, weights: Optional[Qwen2Weights] = None, vision: Optional[VisionFrontEnd] = None,
                 n_env: int = 1, device="cuda", seed: int = 0, lm_head_std: Optional[float] = None):
        self.config = config
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("streamvln_b200 needs a CUDA device (sm_100a kernels; no CPU fallback)")
        self.lib = _lib.load()
        self.model = _Inner(config.num_history)
        self.weights = weights or Qwen2Weights.random(config.llm, self.device, seed=seed, lm_head_std=lm_head_std)
        self.vision = vision or VisionFrontEnd(config.vision, self.device, seed=seed + 1)
        self.engine = Qwen2Engine(config.llm, self.weights, n_env=n_env, device=self.device)
        self.tokens_per_frame = config.vision.pooled_per_side ** 2
        self.reset(n_env)

    # ------------------------------------------------------------------ checkpoints
    @classmethod
    def from_pretrained(cls, path, device="cuda", n_env: int = 1, mode: str = "reference", **overrides) -> "StreamVLNForCausalLM":
        """Load a StreamVLN / LLaVA-Video-Qwen2 checkpoint directory (HF layout: config.json + *.safetensors with the
        names `model.*`, `model.vision_tower.vision_tower.*`, `model.mm_projector.*`, `lm_head.*`), the counterpart of
        `StreamVLNForCausalLM.from_pretrained(model_path, torch_dtype=bfloat16, ...)` at streamvln_agent.py:281-293 /
        streamvln_eval.py:521-533.  `overrides` set StreamVLNConfig fields (window_turns, voxel_size, ...)."""
        import json
        from pathlib import Path

        from safetensors import safe_open

        path = Path(path)
        hc = json.loads((path / "config.json").read_text())
        llm = Qwen2Config(hidden_size=hc["hidden_size"], num_hidden_layers=hc["num_hidden_layers"],
                          num_attention_heads=hc["num_attention_heads"], num_key_value_heads=hc.get("num_key_value_heads", hc["num_attention_heads"]),
                          head_dim=hc.get("head_dim") or hc["hidden_size"] // hc["num_attention_heads"], intermediate_size=hc["intermediate_size"],
                          vocab_size=hc["vocab_size"], rope_theta=float(hc.get("rope_theta", 1e6)), rms_norm_eps=float(hc.get("rms_norm_eps", 1e-6)),
                          max_position_embeddings=int(hc.get("tokenizer_model_max_length") or hc.get("svln_arena_rows") or 4096))
        vc = hc.get("svln_vision") or {}
        vision_cfg = VisionConfig(out_dim=llm.hidden_size, pool_mode=hc.get("mm_spatial_pool_mode", "bilinear"),
                                  pool_stride=int(hc.get("mm_spatial_pool_stride", 2)), **vc)
        eos = hc.get("eos_token_id")
        cfg = StreamVLNConfig(llm=llm, vision=vision_cfg, num_history=hc.get("num_history", 8), mode=mode,
                              eos_token_id=eos[0] if isinstance(eos, list) else eos)
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown StreamVLNConfig field {k!r}")
            setattr(cfg, k, v)
        sd: Dict[str, torch.Tensor] = {}
        shards = sorted(path.glob("*.safetensors"))
        if not shards:
            raise FileNotFoundError(f"no *.safetensors under {path}")
        for f in shards:
            with safe_open(str(f), framework="pt", device="cpu") as st:
                for k in st.keys():
                    sd[k] = st.get_tensor(k)
        weights = Qwen2Weights.from_state_dict(llm, sd, device)
        vision = VisionFrontEnd(vision_cfg, device)
        vision.load_hf_state_dict(sd)
        return cls(cfg, weights=weights, vision=vision, n_env=n_env, device=device)

    def save_pretrained(self, path) -> None:
        """Write config.json + model.safetensors in the layout `from_pretrained` (and HF) reads."""
        import json
        from pathlib import Path

        from safetensors.torch import save_file

        path = Path(path)
        path.mkdir(parents=True, exist_ok=True)
        c, v = self.config.llm, self.config.vision
        hc = dict(architectures=["StreamVLNForCausalLM"], model_type="llava_qwen", hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers,
                  num_attention_heads=c.num_attention_heads, num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim,
                  intermediate_size=c.intermediate_size, vocab_size=c.vocab_size, rope_theta=c.rope_theta, rms_norm_eps=c.rms_norm_eps,
                  svln_arena_rows=c.max_position_embeddings, num_history=self.model.num_history, eos_token_id=self.config.eos_token_id,
                  mm_spatial_pool_mode=v.pool_mode, mm_spatial_pool_stride=v.pool_stride, mm_projector_type="mlp2x_gelu",
                  svln_vision=dict(hidden_size=v.hidden_size, intermediate_size=v.intermediate_size, num_layers=v.num_layers, num_heads=v.num_heads,
                                   image_size=v.image_size, patch_size=v.patch_size, layer_norm_eps=v.layer_norm_eps))
        (path / "config.json").write_text(json.dumps(hc, indent=1))
        sd = {k: t.cpu().contiguous() for k, t in self.weights.to_hf_state_dict().items()}
        sd.update({k: t.cpu().contiguous() for k, t in self.vision.to_hf_state_dict().items()})
        save_file(sd, str(path / "model.safetensors"))

    # ------------------------------------------------------------------ reference object surface
    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.vision

    def eval(self):
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    def to(self, *args, **kwargs):
        return self

    def reset(self, env_num: int) -> None:
        """stream_video_vln.py:473-475 (without the `[dict()] * n` aliasing)."""
        if env_num > self.engine.arena.n_env:
            self.engine = Qwen2Engine(self.config.llm, self.weights, n_env=env_num, device=self.device)
        self.curr_t = [0] * env_num
        self.cache: List[Dict[str, Any]] = [dict() for _ in range(env_num)]
        # per-frame pooled features of the CURRENT episode of each stream (an LRU; a miss only costs a re-encode)
        self.feat_cache: List["OrderedDict[int, torch.Tensor]"] = [OrderedDict() for _ in range(env_num)]
        self.feat_episode: List[Any] = [None] * env_num
        # memory = "voxel": depth / pose / intrinsics of every model-step frame (device fp32), same keys as feat_cache
        self.geom_cache: List[Dict[int, tuple]] = [dict() for _ in range(env_num)]
        self.last_memory: List[Optional[torch.Tensor]] = [None] * env_num     # introspection (tests): rows spliced at <memory>
        self._mem_pruner = None
        for e in range(env_num):
            self.engine.arena.reset_env(e)
        if self.config.mode == "slowfast":
            from .slowfast import SlowFastBatch

            self.sfb = SlowFastBatch(self, max(env_num, self.engine.arena.n_env))
            self.sf = self.sfb.states

    def reset_for_env(self, env_idx: int) -> None:
        """stream_video_vln.py:477-479: forget the stream's embeds (the KV goes with `past_key_values=None`).
        The reference calls this at every window boundary AND at every episode start (streamvln_eval.py:225,347), so
        it cannot drop the per-frame feature cache (a window start re-uses its history frames); stale features
        are excluded by the `episode_id` every request carries (see `_cache_for`)."""
        self.curr_t[env_idx] = 0
        self.cache[env_idx] = dict()
        if self.config.mode == "slowfast":
            self.sf[env_idx].reset()
            self.engine.arena.reset_env(env_idx)

    def reset_episode(self, env_idx: int) -> None:
        """New episode on this stream: also drop the per-frame feature cache."""
        self.reset_for_env(env_idx)
        self.feat_cache[env_idx] = OrderedDict()
        self.geom_cache[env_idx] = dict()
        self.feat_episode[env_idx] = None
        self.engine.arena.reset_env(env_idx)

    def _cache_for(self, env_id: int, episode_id) -> "OrderedDict[int, torch.Tensor]":
        """Feature cache of stream env_id for the episode `episode_id` (a nonce the agent draws per episode): frame
        ids restart at 0 with every episode, so a cache that belongs to another episode is dropped, whoever forgot
        to call reset_episode."""
        if self.feat_episode[env_id] != episode_id:
            self.feat_cache[env_id] = OrderedDict()
            self.geom_cache[env_id] = dict()
            self.feat_episode[env_id] = episode_id
        return self.feat_cache[env_id]

    def _cache_limit(self) -> int:
        if self.config.feat_cache_frames is not None:
            return self.config.feat_cache_frames
        per_frame = self.tokens_per_frame * self.config.llm.hidden_size * 2
        return max(16, min(1024, self.config.feat_cache_bytes // (per_frame * max(len(self.feat_cache), 1))))

    def _cache_put(self, cache, frame_id: int, feat: torch.Tensor) -> None:
        cache[frame_id] = feat
        limit = self._cache_limit()
        while len(cache) > limit:
            old, _ = cache.popitem(last=False)
            for env, c in enumerate(self.feat_cache):
                if c is cache:
                    self.geom_cache[env].pop(old, None)

    # ------------------------------------------------------------------ voxel-pruned slow memory (memory = "voxel")
    def _remember_geometry(self, env_id: int, frame_ids, depths, poses, intrinsics) -> None:
        g = self.geom_cache[env_id]
        for i, f in enumerate(frame_ids):
            if f not in g:
                g[f] = (depths[0, i].to(self.device, torch.float32).contiguous(), poses[0, i].to(self.device, torch.float32).reshape(4, 4).contiguous(),
                        intrinsics[0, i].to(self.device, torch.float32).reshape(4, 4).contiguous())

    def _voxel_memory(self, env_id: int, before_frame: int) -> Optional[torch.Tensor]:
        """north_star subsystem (1) as the reference's plumbing intended it (depths / poses / intrinsics reach encode_rgbd,
        stream_video_vln.py:102, and are ignored): the <memory> block of a window start = the pooled tokens of EVERY past
        model-step frame, back-projected to voxels, keep-latest-per-voxel, stably compacted and gathered by
        svln_voxel_prune (one batched call), newest `memory_cap` survivors.  Returns [1, n, D] or None."""
        cache, geom = self.feat_cache[env_id], self.geom_cache[env_id]
        frames = sorted(f for f in cache if f in geom and f < before_frame)
        if not frames:
            return None
        T = len(frames)
        cfg = self.config
        G = cfg.vision.pooled_per_side
        if self._mem_pruner is None or self._mem_pruner.max_frames < T:
            self._mem_pruner = VoxelPruner(VoxelPruneConfig(grid=G, dim=cfg.llm.hidden_size, voxel_size=cfg.voxel_size, max_depth=cfg.max_depth,
                                                            invalid_policy=cfg.invalid_policy), 1, max(2 * T, 32), self.device)
        depth = torch.stack([geom[f][0] for f in frames])[None]
        pose = torch.stack([geom[f][1] for f in frames])[None]
        intr = torch.stack([geom[f][2] for f in frames])[None]
        feat = torch.stack([cache[f] for f in frames])[None].contiguous()
        res = self._mem_pruner.prune(depth, pose, intr, feat)
        self.engine.launches += 7
        n = res.n_keep()                 # the one host read of a window start: the embed length depends on it
        lo = max(0, n - cfg.memory_cap)
        return res.feat[lo:n].clone()[None]

    # ------------------------------------------------------------------ encode_rgbd
    @torch.no_grad()
    def encode_rgbd(self, images: torch.Tensor, depths, poses, intrinsics, time_ids=None, task_ids=None,
                    env_id: int = 0, frame_ids: Optional[Sequence[int]] = None, episode_id=None):
        """stream_video_vln.py:102-142 for batch size 1: returns (image_features [V',196,D],
        memory_features [1, nh*196, D] or None).  `frame_ids` (optional, one global frame index per view)
        enables the per-frame pooled-feature cache."""
        assert images.shape[0] == 1, "one stream per call (the reference call sites use batch size 1)"
        V = images.shape[1]
        feats = self._encode_views(images[0], env_id, frame_ids, episode_id)
        voxel = self.config.memory == "voxel" and self.config.mode == "reference"
        if voxel:
            if frame_ids is None or depths is None or poses is None or intrinsics is None:
                raise ValueError('memory="voxel" needs frame_ids and fp32 depths / poses / intrinsics with every request')
            self._remember_geometry(env_id, frame_ids, depths, poses, intrinsics)
        start_idx = 0
        if time_ids is not None and time_ids[0] is not None and len(time_ids[0]):
            start_idx = time_ids[0][0]
        if V != 1 and start_idx != 0:
            nh = self.model.num_history
            if voxel:
                memory = self._voxel_memory(env_id, before_frame=frame_ids[-1])
                self.last_memory[env_id] = memory
                return feats[nh:], memory
            memory = feats[:nh].reshape(1, -1, feats.shape[-1])
            return feats[nh:], memory
        return feats, None

    def _encode_views(self, views: torch.Tensor, env_id: int, frame_ids, episode_id=None) -> torch.Tensor:
        if frame_ids is None or self.config.mode == "slowfast":   # slowfast never looks at an old frame again
            return self.vision(views.to(self.device))
        cache = self._cache_for(env_id, episode_id)
        missing = [i for i, f in enumerate(frame_ids) if f not in cache]
        fresh = {}
        if missing:
            new = self.vision(views[missing].to(self.device))
            for j, i in enumerate(missing):
                fresh[frame_ids[i]] = new[j]
        out = []
        for f in frame_ids:
            if f in fresh:
                out.append(fresh[f])
            else:
                cache.move_to_end(f)
                out.append(cache[f])
        for f, ft in fresh.items():
            self._cache_put(cache, f, ft)
        return torch.stack(out)

    def _prefetch_features(self, requests: Sequence[Dict[str, Any]], max_batch: int = 16) -> None:
        """Encode, in as few SigLIP batches as possible, every frame of this batch of turns that is not in the
        per-stream feature cache yet (requests that carry `frame_ids`)."""
        todo = []
        if self.config.mode == "slowfast":
            return
        for r in requests:
            fids, images = r.get("frame_ids"), r.get("images")
            if fids is None or images is None:
                continue
            env = r.get("env_id", 0) or 0
            cache = self._cache_for(env, r.get("episode_id"))
            for i, f in enumerate(fids):
                if f not in cache and all(not (e == env and g == f) for e, g, _ in todo):
                    todo.append((env, f, images[0, i]))
        if len(todo) <= 1:
            return  # a single frame goes through the (CUDA-graphed) per-request path
        for b0 in range(0, len(todo), max_batch):
            part = todo[b0:b0 + max_batch]
            feats = self.vision(torch.stack([x[2].to(self.device) for x in part]))
            for (env, f, _), ft in zip(part, feats):
                self._cache_put(self.feat_cache[env], f, ft)

    # ------------------------------------------------------------------ embed splice
    def splice(self, input_ids: Sequence[int], image_features: torch.Tensor, memory_features: Optional[torch.Tensor]) -> torch.Tensor:
        """prepare_inputs_labels_for_multimodal (stream_video_vln.py:144-291), inference path, one sample:
        every IMAGE_TOKEN_INDEX is replaced by the next frame's 196 rows, every MEMORY_TOKEN_INDEX by the
        next memory block, in order of appearance.  One gather kernel instead of the Python cat loop."""
        tpf = image_features.shape[1] if image_features.numel() else self.tokens_per_frame
        D = self.config.llm.hidden_size
        blocks = [image_features.reshape(-1, D)]
        mem_base = blocks[0].shape[0]
        mem_rows = 0
        if memory_features is not None:
            blocks.append(memory_features.reshape(-1, D))
            mem_rows = blocks[1].shape[0]
        feat = torch.cat(blocks) if len(blocks) > 1 else blocks[0]
        n_img = image_features.shape[0]
        n_mem = 0 if memory_features is None else memory_features.shape[0]
        row_src: List[int] = []
        img_i = mem_i = 0
        for tid in input_ids:
            if tid == IMAGE_TOKEN_INDEX:
                if img_i >= n_img:
                    raise IndexError("more <image> placeholders than encoded frames")
                row_src.extend(range(-(img_i * tpf) - 1, -((img_i + 1) * tpf) - 1, -1))
                img_i += 1
            elif tid == MEMORY_TOKEN_INDEX:
                if mem_i >= n_mem:
                    raise IndexError("more <memory> placeholders than memory blocks")
                per = mem_rows // n_mem
                row_src.extend(range(-(mem_base + mem_i * per) - 1, -(mem_base + (mem_i + 1) * per) - 1, -1))
                mem_i += 1
            elif tid < 0:
                raise NotImplementedError(f"unknown special token {tid}")  # stream_video_vln.py:230
            else:
                row_src.append(tid)
        if self.config.tokenizer_model_max_length is not None:
            row_src = row_src[: self.config.tokenizer_model_max_length]
        n = len(row_src)
        src = torch.tensor(row_src, dtype=torch.int32).to(self.device, non_blocking=True)
        out = torch.empty(n, D, dtype=torch.bfloat16, device=self.device)
        check(self.lib.svln_splice_rows(ptr(self.weights.embed), ptr(feat.contiguous()), ptr(src), n, D * 2, ptr(out), cur_stream()))
        self.engine.launches += 1
        return out

    # ------------------------------------------------------------------ generate
    @torch.no_grad()
    def generate(self, inputs: Optional[torch.Tensor] = None, images: Optional[torch.Tensor] = None,
                 image_sizes=None, depths: Optional[torch.Tensor] = None, poses: Optional[torch.Tensor] = None,
                 intrinsics: Optional[torch.Tensor] = None, task_ids=None, **kwargs) -> GenerateOutput:
        """stream_video_vln.py:353-407 + HF greedy `generate` (do_sample=False, num_beams=1)."""
        return self.generate_batch([dict(inputs=inputs, images=images, depths=depths, poses=poses,
                                         intrinsics=intrinsics, **kwargs)])[0]

    @torch.no_grad()
    def generate_batch(self, requests: Sequence[Dict[str, Any]]) -> List[GenerateOutput]:
        """Several streams' turns in one packed forward (one request per env_id; same kwargs as `generate`).
        This is what lets 64 concurrent episodes share every GEMM."""
        tails, envs, max_new = [], [], []
        slowfast = self.config.mode == "slowfast"
        ids_list = []
        for r in requests:
            if "inputs_embeds" in r:
                raise NotImplementedError("`inputs_embeds` is not supported")  # stream_video_vln.py:369-370
            if r.get("do_sample", False) or r.get("num_beams", 1) != 1:
                raise NotImplementedError("only greedy decoding (do_sample=False, num_beams=1) is supported")
            ids = r["inputs"]
            ids_list.append(ids.reshape(-1).tolist() if torch.is_tensor(ids) else list(ids))
            envs.append(r.get("env_id", 0) or 0)
            max_new.append(int(r.get("max_new_tokens", 20)))
        if len(set(envs)) != len(envs):
            raise ValueError("one request per env_id in a batch")
        if slowfast:
            tails = self.sfb.prepare_turns(requests, ids_list)
        else:
            self._prefetch_features(requests)
            for r, ids, env in zip(requests, ids_list, envs):
                time_ids = r.get("time_ids")
                if r.get("images") is not None:
                    img_f, mem_f = self.encode_rgbd(r["images"], r.get("depths"), r.get("poses"), r.get("intrinsics"),
                                                    time_ids, r.get("task_type"), env, r.get("frame_ids"), r.get("episode_id"))
                    emb = self.splice(ids, img_f, mem_f)
                else:
                    emb = self.weights.embed.index_select(0, torch.tensor(ids, dtype=torch.int64).to(self.device))
                # per-env embed cache (stream_video_vln.py:396-401), kept as chunks: only a re-prefill reads it
                c = self.cache[env]
                if self.curr_t[env] == 0:
                    c["chunks"], c["len"] = [emb], emb.shape[0]
                else:
                    c["chunks"].append(emb)
                    c["len"] += emb.shape[0]
                self.curr_t[env] += 1
                past = r.get("past_key_values")
                past_len = self.engine.arena.lens[env] if past is not None else 0
                if past is None:
                    self.engine.arena.reset_env(env)
                first, n_tail = uncached_tail(c["len"], past_len)
                tails.append(emb[emb.shape[0] - n_tail:] if n_tail <= emb.shape[0] else torch.cat(c["chunks"])[first:])
        n_new = max(max_new)
        seqs = self.engine.generate(tails, envs, n_new, self.config.eos_token_id)
        outs, evict = [], []
        for r, env, s, mn in zip(requests, envs, seqs, max_new):
            if s.shape[0] > mn:  # a shorter budget than the batch maximum: un-feed the surplus tokens
                self.engine.arena.lens[env] -= s.shape[0] - mn
                s = s[:mn]
            if slowfast and self.sf[env].finish_turn(int(s.shape[0])):
                evict.append(env)
            outs.append(GenerateOutput(sequences=s.unsqueeze(0), past_key_values=KvHandle(env, self.engine.arena)))
        if evict:
            self.sfb.evict(evict)   # all streams that slide this turn: one batched, device-side pass
        return outs

"""Vision front-end: SigLIP tower -> mlp2x_gelu projector -> 2x2 bilinear pool (27x27 -> 14x14).

This is the step BEFORE the hot path (SURVEY.md §8f-1 "next"), kept as plain torch library calls
(cuBLAS GEMMs, SDPA): llava/model/multimodal_encoder/siglip_encoder.py:538-589 (SigLIP-so400m with the
last encoder layer dropped and the head replaced by Identity, hidden_states[-1] = output of layer 26,
no post-layernorm), llava/model/multimodal_projector/builder.py:32-65 (mlp2x_gelu) and
streamvln/model/stream_video_vln.py:53-73 (get_2dPool, mode "bilinear", ceil(27/2) = 14).
What this module adds is a per-frame cache of pooled features, so a window start never re-encodes
its num_history frames.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class VisionConfig:
    hidden_size: int = 1152
    intermediate_size: int = 4304
    num_layers: int = 26          # 27 in the checkpoint, last one dropped (siglip_encoder.py:570)
    num_heads: int = 16
    image_size: int = 384
    patch_size: int = 14
    layer_norm_eps: float = 1e-6
    out_dim: int = 3584           # LLM hidden size (projector output)
    pool_stride: int = 2
    pool_mode: str = "bilinear"   # mm_spatial_pool_mode (streamvln/args.py:41)

    @property
    def patches_per_side(self) -> int:
        return self.image_size // self.patch_size

    @property
    def pooled_per_side(self) -> int:
        n = self.patches_per_side
        return math.ceil(n / self.pool_stride) if self.pool_mode == "bilinear" else n // self.pool_stride

    @staticmethod
    def tiny(out_dim: int = 256) -> "VisionConfig":
        return VisionConfig(hidden_size=64, intermediate_size=128, num_layers=2, num_heads=2, out_dim=out_dim)


class SigLipImageProcessor:
    """384x384 bicubic squash, rescale 1/255, normalise mean=std=0.5 (siglip_encoder.py:34-66)."""

    def __init__(self, size=(384, 384)):
        self.size = size
        self.crop_size = {"height": size[0], "width": size[1]}
        self.image_mean = (0.5, 0.5, 0.5)
        self.image_std = (0.5, 0.5, 0.5)

    def preprocess(self, images, return_tensors="pt"):
        from PIL import Image

        if isinstance(images, Image.Image):
            images = [images]
        out = []
        for im in images:
            if not isinstance(im, Image.Image):
                im = Image.fromarray(np.asarray(im))
            im = im.convert("RGB").resize((self.size[1], self.size[0]), resample=Image.BICUBIC)
            a = np.asarray(im, dtype=np.float32) * (1.0 / 255.0)
            a = (a - 0.5) / 0.5
            out.append(torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))))
        return {"pixel_values": out if return_tensors != "pt" else torch.stack(out)}


class VisionFrontEnd(torch.nn.Module):
    def __init__(self, cfg: VisionConfig, device="cuda", dtype=torch.bfloat16, seed: int = 0):
        super().__init__()
        self.cfg = cfg
        self.image_processor = SigLipImageProcessor((cfg.image_size, cfg.image_size))
        g = torch.Generator().manual_seed(seed)
        D, I = cfg.hidden_size, cfg.intermediate_size
        n_pos = cfg.patches_per_side ** 2

        def p(*shape, s=0.02):
            return torch.nn.Parameter((torch.randn(*shape, generator=g) * s).to(dtype), requires_grad=False)

        self.patch_w = p(D, 3 * cfg.patch_size * cfg.patch_size)
        self.patch_b = p(D)
        self.pos = p(n_pos, D)
        self.layers = torch.nn.ParameterList()
        self._layer_names = ("ln1_w", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2")
        for _ in range(cfg.num_layers):
            ps = [torch.nn.Parameter(torch.ones(D, dtype=dtype), requires_grad=False), p(D, s=0.0), p(3 * D, D), p(3 * D),
                  p(D, D), p(D), torch.nn.Parameter(torch.ones(D, dtype=dtype), requires_grad=False), p(D, s=0.0),
                  p(I, D), p(I), p(D, I), p(D)]
            self.layers.extend(ps)
        self.proj_w1, self.proj_b1 = p(cfg.out_dim, D), p(cfg.out_dim)
        self.proj_w2, self.proj_b2 = p(cfg.out_dim, cfg.out_dim), p(cfg.out_dim)
        self.to(device)
        self.use_graph = os.environ.get("SVLN_NO_GRAPH", "0") != "1"
        self._graphs: Dict[int, dict] = {}
        self._seen: Dict[int, int] = {}

    @torch.no_grad()
    def load_hf_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "model.vision_tower.vision_tower.",
                           projector_prefix: str = "model.mm_projector.") -> None:
        """Load the tower + projector from a StreamVLN / LLaVA checkpoint's state dict (HF SigLIP names:
        `{prefix}vision_model.encoder.layers.N.self_attn.q_proj.weight`, ...; projector `{projector_prefix}0.weight`,
        `{projector_prefix}2.weight` for mlp2x_gelu).  q/k/v are fused row-wise; the patch-embedding conv becomes a
        [D, 3*14*14] matrix (same (c, ky, kx) order as the conv weight).  Layers beyond `num_layers` (the dropped
        27th) and the pooling head are ignored, as in siglip_encoder.py:570-571."""
        dt, dev = self.patch_w.dtype, self.patch_w.device

        def t(name):
            return sd[prefix + "vision_model." + name].to(device=dev, dtype=dt)

        self.patch_w.copy_(t("embeddings.patch_embedding.weight").reshape(self.cfg.hidden_size, -1))
        self.patch_b.copy_(t("embeddings.patch_embedding.bias"))
        self.pos.copy_(t("embeddings.position_embedding.weight"))
        for i in range(self.cfg.num_layers):
            L, e = self._layer(i), f"encoder.layers.{i}."
            L["ln1_w"].copy_(t(e + "layer_norm1.weight")); L["ln1_b"].copy_(t(e + "layer_norm1.bias"))
            L["wqkv"].copy_(torch.cat([t(e + f"self_attn.{n}_proj.weight") for n in "qkv"]))
            L["bqkv"].copy_(torch.cat([t(e + f"self_attn.{n}_proj.bias") for n in "qkv"]))
            L["wo"].copy_(t(e + "self_attn.out_proj.weight")); L["bo"].copy_(t(e + "self_attn.out_proj.bias"))
            L["ln2_w"].copy_(t(e + "layer_norm2.weight")); L["ln2_b"].copy_(t(e + "layer_norm2.bias"))
            L["w1"].copy_(t(e + "mlp.fc1.weight")); L["b1"].copy_(t(e + "mlp.fc1.bias"))
            L["w2"].copy_(t(e + "mlp.fc2.weight")); L["b2"].copy_(t(e + "mlp.fc2.bias"))
        if projector_prefix is not None and projector_prefix + "0.weight" in sd:
            self.proj_w1.copy_(sd[projector_prefix + "0.weight"].to(dev, dt)); self.proj_b1.copy_(sd[projector_prefix + "0.bias"].to(dev, dt))
            self.proj_w2.copy_(sd[projector_prefix + "2.weight"].to(dev, dt)); self.proj_b2.copy_(sd[projector_prefix + "2.bias"].to(dev, dt))
        self._graphs.clear()

    def to_hf_state_dict(self, prefix: str = "model.vision_tower.vision_tower.", projector_prefix: str = "model.mm_projector.") -> Dict[str, torch.Tensor]:
        """Inverse of load_hf_state_dict (checkpoint round trips in the tests / save_pretrained)."""
        cfg = self.cfg
        D, ps = cfg.hidden_size, cfg.patch_size
        sd = {"embeddings.patch_embedding.weight": self.patch_w.reshape(D, 3, ps, ps), "embeddings.patch_embedding.bias": self.patch_b,
              "embeddings.position_embedding.weight": self.pos}
        for i in range(cfg.num_layers):
            L, e = self._layer(i), f"encoder.layers.{i}."
            for j, n in enumerate("qkv"):
                sd[e + f"self_attn.{n}_proj.weight"] = L["wqkv"][j * D:(j + 1) * D]
                sd[e + f"self_attn.{n}_proj.bias"] = L["bqkv"][j * D:(j + 1) * D]
            sd.update({e + "layer_norm1.weight": L["ln1_w"], e + "layer_norm1.bias": L["ln1_b"], e + "self_attn.out_proj.weight": L["wo"],
                       e + "self_attn.out_proj.bias": L["bo"], e + "layer_norm2.weight": L["ln2_w"], e + "layer_norm2.bias": L["ln2_b"],
                       e + "mlp.fc1.weight": L["w1"], e + "mlp.fc1.bias": L["b1"], e + "mlp.fc2.weight": L["w2"], e + "mlp.fc2.bias": L["b2"]})
        out = {prefix + "vision_model." + k: v.detach().contiguous() for k, v in sd.items()}
        out.update({projector_prefix + "0.weight": self.proj_w1.detach(), projector_prefix + "0.bias": self.proj_b1.detach(),
                    projector_prefix + "2.weight": self.proj_w2.detach(), projector_prefix + "2.bias": self.proj_b2.detach()})
        return out

    # attribute access used by the reference's call sites (streamvln_agent.py:43)
    @property
    def num_patches_per_side(self) -> int:
        return self.cfg.patches_per_side

    def _layer(self, i: int) -> Dict[str, torch.Tensor]:
        n = len(self._layer_names)
        return {k: self.layers[i * n + j] for j, k in enumerate(self._layer_names)}

    @torch.no_grad()
    def encode_patches(self, images: torch.Tensor) -> torch.Tensor:
        """images [F,3,384,384] -> [F,729,hidden] (hidden_states[-1] of the truncated tower)."""
        cfg = self.cfg
        Fn = images.shape[0]
        ps = cfg.patch_size
        n = cfg.patches_per_side
        x = images[:, :, : n * ps, : n * ps].to(self.patch_w.dtype)
        x = x.unfold(2, ps, ps).unfold(3, ps, ps)                      # [F,3,n,n,ps,ps]
        x = x.permute(0, 2, 3, 1, 4, 5).reshape(Fn, n * n, 3 * ps * ps)
        x = F.linear(x, self.patch_w, self.patch_b) + self.pos
        H = cfg.num_heads
        hd = cfg.hidden_size // H
        for i in range(cfg.num_layers):
            L = self._layer(i)
            y = F.layer_norm(x, (cfg.hidden_size,), L["ln1_w"], L["ln1_b"], cfg.layer_norm_eps)
            qkv = F.linear(y, L["wqkv"], L["bqkv"]).view(Fn, n * n, 3, H, hd).permute(2, 0, 3, 1, 4)
            a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
            x = x + F.linear(a.transpose(1, 2).reshape(Fn, n * n, cfg.hidden_size), L["wo"], L["bo"])
            y = F.layer_norm(x, (cfg.hidden_size,), L["ln2_w"], L["ln2_b"], cfg.layer_norm_eps)
            x = x + F.linear(F.gelu(F.linear(y, L["w1"], L["b1"]), approximate="tanh"), L["w2"], L["b2"])
        return x

    @torch.no_grad()
    def project_pool(self, patches: torch.Tensor) -> torch.Tensor:
        """[F,729,hidden] -> mlp2x_gelu -> get_2dPool -> [F,196,out_dim] (stream_video_vln.py:138-139)."""
        cfg = self.cfg
        x = F.linear(F.gelu(F.linear(patches, self.proj_w1, self.proj_b1)), self.proj_w2, self.proj_b2)
        return pool_2d(x, cfg.patches_per_side, cfg.pool_stride, cfg.pool_mode)

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """[V,3,384,384] -> [V,196,out_dim].  The single-frame call of every turn (~300 small launches for the
        26-layer tower) is replayed from a CUDA graph after it has been seen twice."""
        V = images.shape[0]
        if not (self.use_graph and images.is_cuda and V <= 2):
            return self.project_pool(self.encode_patches(images))
        g = self._graphs.get(V)
        if g is not None:
            g["inp"].copy_(images)
            g["graph"].replay()
            return g["out"].clone()
        out = self.project_pool(self.encode_patches(images))
        self._seen[V] = self._seen.get(V, 0) + 1
        if self._seen[V] >= 2:
            inp = torch.zeros(images.shape, dtype=self.patch_w.dtype, device=images.device)
            inp.copy_(images)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                gout = self.project_pool(self.encode_patches(inp))
            self._graphs[V] = dict(inp=inp, graph=graph, out=gout)
        return out


def pool_2d(x: torch.Tensor, side: int, stride: int = 2, mode: str = "bilinear") -> torch.Tensor:
    """get_2dPool (stream_video_vln.py:53-73): [F, side*side, D] -> [F, side'^2, D]."""
    Fn, _, D = x.shape
    if mode == "bilinear" and x.is_cuda and x.dtype == torch.bfloat16 and D % 8 == 0:
        # svln_pool_bilinear: ATen's bilinear arithmetic on the token-major layout (torch's NCHW kernel takes 2 ms
        # per frame at D = 3584, more than the rest of the vision front end)
        from . import _lib

        lib = _lib.load()
        side_out = math.ceil(side / stride)
        x = x.contiguous()
        out = torch.empty(Fn, side_out * side_out, D, dtype=x.dtype, device=x.device)
        _lib.check(lib.svln_pool_bilinear(_lib.ptr(x), Fn, side, side_out, D, _lib.ptr(out), _lib.cur_stream()))
        return out
    x = x.view(Fn, side, side, D).permute(0, 3, 1, 2).contiguous()
    if mode == "average":
        x = F.avg_pool2d(x, stride)
    elif mode == "max":
        x = F.max_pool2d(x, stride)
    elif mode == "bilinear":
        x = F.interpolate(x, size=[math.ceil(side / stride), math.ceil(side / stride)], mode="bilinear")
    else:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")
    return x.permute(0, 2, 3, 1).reshape(Fn, -1, D)

"""Generate tests/golden/ref_glue_*.npz by running the REFERENCE's own code on CPU.

Runs only in the build container (needs /root/reference, read-only; nothing is copied from it).
Imports streamvln/model/stream_video_vln.py under the two shims SURVEY.md §8c describes
(transformers 5.5 here vs 4.45.1 pinned):
  1. a stub module for llava.model.multimodal_resampler.qformer (its imports are gone in 5.x);
  2. a Qwen2Config subclass that ignores `config.rope_scaling = None` (stream_video_vln.py:40).
The SigLIP tower is replaced by a module returning seeded features (the tower is a third-party
checkpoint that cannot be downloaded here), so the fixtures pin exactly the reference's glue:
encode_rgbd (:102-142), get_2dPool (:53-73), prepare_inputs_labels_for_multimodal (:144-291).

    python scripts/gen_golden_reference.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REF + "/streamvln")
stub = types.ModuleType("llava.model.multimodal_resampler.qformer")
stub.Qformer = type("Qformer", (), {})
sys.modules["llava.model.multimodal_resampler.qformer"] = stub

from transformers import Qwen2Config  # noqa: E402
from model.stream_video_vln import StreamVLNForCausalLM  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
IMAGE, MEMORY = -200, -300


class Cfg(Qwen2Config):
    model_type = "llava_qwen"

    def __setattr__(self, k, v):
        if k == "rope_scaling" and v is None:
            return
        super().__setattr__(k, v)


class FakeTower(torch.nn.Module):
    """Stands in for SigLipVisionTower: returns stored [V,729,C] features for any input."""

    num_patches_per_side = 27

    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def forward(self, images):
        assert images.shape[0] == self.feats.shape[0]
        return self.feats


def build(C=16, D=32, vocab=64, num_history=2, pool="bilinear"):
    cfg = Cfg(hidden_size=D, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1,
              vocab_size=vocab, max_position_embeddings=4096)
    cfg.mm_vision_tower = "google/siglip-so400m-patch14-384"
    cfg.mm_projector_type = "mlp2x_gelu"
    cfg.mm_hidden_size = C
    cfg.mm_spatial_pool_mode = pool
    cfg.num_history = num_history
    torch.manual_seed(0)
    m = StreamVLNForCausalLM(cfg).float().eval()
    m.model.num_history = num_history
    return m


def case(name, m, V, ids, time_ids, C=16):
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    tower_out = torch.randn(V, 729, C, generator=g)
    tower = FakeTower(tower_out)
    m.get_model().get_vision_tower = lambda: tower
    m.get_vision_tower = lambda: tower
    images = torch.zeros(1, V, 3, 4, 4)  # content is irrelevant to the fake tower
    input_ids = torch.tensor([ids], dtype=torch.long)
    with torch.no_grad():
        img_f, mem_f = m.encode_rgbd(images, None, None, None, time_ids, None)
        out = m.prepare_inputs_labels_for_multimodal(input_ids, None, None, None, None, images, None, None, None, None,
                                                     time_ids, None)
    _, position_ids, attention_mask, _, embeds, _ = out
    assert position_ids is None and attention_mask is None
    proj = m.get_model().mm_projector
    np.savez_compressed(
        OUT / f"ref_glue_{name}.npz",
        tower_out=tower_out.numpy(), input_ids=np.asarray(ids, np.int64), time_ids=np.asarray(time_ids[0], np.int64),
        num_history=np.int64(m.model.num_history), pool_mode=np.bytes_(m.config.mm_spatial_pool_mode),
        proj_w1=proj[0].weight.detach().numpy(), proj_b1=proj[0].bias.detach().numpy(),
        proj_w2=proj[2].weight.detach().numpy(), proj_b2=proj[2].bias.detach().numpy(),
        embed=m.get_model().embed_tokens.weight.detach().numpy(),
        image_features=img_f[0].numpy(),  # list over the batch (B = 1) of [V',196,D]
        memory_features=np.zeros(0, np.float32) if mem_f[0] is None else mem_f[0].numpy(),
        inputs_embeds=embeds.numpy())
    print(name, "embeds", tuple(embeds.shape))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    m = build()
    # a normal turn: text + one <image> (streamvln_agent.py:212-217); V = 1, window started at step 0
    case("turn", m, 1, [5, 9, IMAGE, 7, 11, 3], [[0, 1, 2, 3, 4]])
    # first turn of an episode (time_ids[0] == 0)
    case("first", m, 1, [1, 2, 3, IMAGE, 4], [[0]])
    # window start: num_history views of memory + the current frame; <memory> before <image>
    case("window_start", m, 3, [5, 9, 21, MEMORY, 13, IMAGE, 7, 3], [[32]])
    # edge: several views while the window starts at 0 -> every view is an <image> candidate
    case("multi_view_no_memory", m, 2, [8, IMAGE, 2, IMAGE], [[0]])
    for pool in ("average", "max"):
        m2 = build(pool=pool)
        case(f"pool_{pool}", m2, 1, [4, IMAGE, 6], [[0]])


if __name__ == "__main__":
    main()

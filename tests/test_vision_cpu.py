"""CPU: the vision front-end (SigLIP tower as used by the reference: last encoder layer dropped, no post-layernorm,
no pooling head; mlp2x_gelu projector; 2x2 bilinear pool) against transformers' own SiglipVisionModel with the same
weights, loaded through the HF state-dict names a StreamVLN checkpoint uses."""
import torch

from oracle import glue
from streamvln_b200.vision import VisionConfig, VisionFrontEnd


def test_tower_matches_transformers_siglip():
    from transformers import SiglipVisionConfig, SiglipVisionModel

    torch.manual_seed(0)
    hcfg = SiglipVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                              image_size=384, patch_size=14)
    hf = SiglipVisionModel(hcfg).eval()
    # the reference deletes the last encoder layer and replaces the head (siglip_encoder.py:570-571)
    del hf.vision_model.encoder.layers[-1:]
    hf.vision_model.head = torch.nn.Identity()
    cfg = VisionConfig(hidden_size=64, intermediate_size=128, num_layers=3, num_heads=4, out_dim=96)
    v = VisionFrontEnd(cfg, device="cpu", dtype=torch.float32, seed=1)
    sd = {"model.vision_tower.vision_tower." + k: t for k, t in hf.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    sd.update({"model.mm_projector.0.weight": torch.randn(96, 64, generator=g) * 0.1, "model.mm_projector.0.bias": torch.randn(96, generator=g) * 0.1,
               "model.mm_projector.2.weight": torch.randn(96, 96, generator=g) * 0.1, "model.mm_projector.2.bias": torch.randn(96, generator=g) * 0.1})
    v.load_hf_state_dict(sd)
    x = torch.randn(2, 3, 384, 384, generator=g)
    with torch.no_grad():
        want = hf(x, output_hidden_states=True).hidden_states[-1]     # siglip_encoder.py:585-586
        got = v.encode_patches(x)
    assert got.shape == (2, 729, 64)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), (got - want).abs().max()
    # projector + pool == the glue oracle (itself pinned to the reference's encode_rgbd by the golden fixtures)
    proj = tuple(sd[f"model.mm_projector.{k}"] for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    want_feat, _ = glue.encode_rgbd(want, proj, 8, [0])
    assert torch.allclose(v.project_pool(got), want_feat, atol=2e-5, rtol=1e-4)
    assert v(x).shape == (2, 196, 96)


def test_vision_state_dict_round_trip():
    """to_hf_state_dict is the inverse of load_hf_state_dict (what save_pretrained / from_pretrained rely on)."""
    cfg = VisionConfig(hidden_size=64, intermediate_size=128, num_layers=2, num_heads=4, out_dim=96)
    a = VisionFrontEnd(cfg, device="cpu", dtype=torch.float32, seed=3)
    b = VisionFrontEnd(cfg, device="cpu", dtype=torch.float32, seed=4)
    b.load_hf_state_dict(a.to_hf_state_dict())
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))
    x = torch.randn(1, 3, 384, 384, generator=torch.Generator().manual_seed(0))
    assert torch.equal(a(x), b(x))

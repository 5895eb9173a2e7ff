"""GPU parity of the vision front-end at SigLIP-so400m dims: `VisionFrontEnd` (bf16, cuBLAS
GEMMs + SDPA + the hand-written bilinear pool, CUDA-graphed) against transformers' own SiglipVisionModel in fp32 on the
same GPU, with the same weights loaded through the HF state-dict names a StreamVLN checkpoint uses.  The reference tower
(llava/model/multimodal_encoder/siglip_encoder.py:177-239 eager attention, :570-571 last layer dropped + Identity head,
:576-589 hidden_states[-1]) runs in bf16; the yardstick for "bf16 within tolerance" is therefore the same HF model in
bf16: the product's error against fp32 must not exceed 1.25 x HF-bf16's own (+1e-3)."""
import pytest
import torch

from oracle import glue

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def _rms(a, b):
    return ((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()


def test_vision_front_end_matches_transformers_siglip_so400m(svln_lib, cuda):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    from streamvln_b200.vision import VisionConfig, VisionFrontEnd

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    n_layers = 6                                    # so400m width/heads/MLP/patching, 6 of the 26 layers (error grows with depth)
    hcfg = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=n_layers + 1, num_attention_heads=16,
                              image_size=384, patch_size=14)
    hcfg._attn_implementation = "eager"
    with torch.device(cuda):
        hf = SiglipVisionModel(hcfg).eval()
    for n, p in hf.named_parameters():              # HF's default init leaves the biases / norms trivial: randomise them
        if p.dim() == 1:
            p.data = (1.0 if "layer_norm" in n and n.endswith("weight") else 0.0) + 0.1 * torch.randn_like(p)
    del hf.vision_model.encoder.layers[-1:]         # siglip_encoder.py:570-571
    hf.vision_model.head = torch.nn.Identity()
    cfg = VisionConfig(num_layers=n_layers)
    v = VisionFrontEnd(cfg, device=cuda, seed=1)
    sd = {"model.vision_tower.vision_tower." + k: t for k, t in hf.state_dict().items()}
    g = torch.Generator(device=cuda).manual_seed(2)
    D = cfg.out_dim
    sd.update({"model.mm_projector.0.weight": torch.randn(D, 1152, generator=g, device=cuda) * 0.03,
               "model.mm_projector.0.bias": torch.randn(D, generator=g, device=cuda) * 0.1,
               "model.mm_projector.2.weight": torch.randn(D, D, generator=g, device=cuda) * 0.02,
               "model.mm_projector.2.bias": torch.randn(D, generator=g, device=cuda) * 0.1})
    v.load_hf_state_dict(sd)
    x = torch.randn(3, 3, 384, 384, generator=g, device=cuda) * 0.6
    proj32 = tuple(sd[f"model.mm_projector.{k}"].float() for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
    with torch.no_grad():
        tower32 = hf(x, output_hidden_states=True).hidden_states[-1]                # siglip_encoder.py:585-586
        want32, _ = glue.encode_rgbd(tower32, proj32, 8, [0])                        # projector + 2x2 bilinear pool, fp32
        hf16 = hf.to(torch.bfloat16)
        tower16 = hf16(x.bfloat16(), output_hidden_states=True).hidden_states[-1]
        want16, _ = glue.encode_rgbd(tower16, tuple(p.bfloat16() for p in proj32), 8, [0])
        got_tower = v.encode_patches(x.bfloat16())
        got = v(x)
        again = v(x)
        one = [v(x[:1]).clone() for _ in range(3)]                                   # V = 1: captured, then replayed from the CUDA graph
    assert got.shape == (3, 196, D) and got.dtype == torch.bfloat16 and got_tower.shape == (3, 729, 1152)
    assert torch.equal(got, again)
    assert torch.equal(one[0], one[1]) and torch.equal(one[1], one[2]) and _rel(one[2].float(), got[:1].float()) <= 1e-2
    e_t, h_t = _rms(got_tower.float(), tower32), _rms(tower16.float(), tower32)
    e_f, h_f = _rms(got.float(), want32), _rms(want16.float(), want32)
    m_f, hm_f = _rel(got.float(), want32), _rel(want16.float(), want32)
    print(f"\n  tower rms rel err: product {e_t:.2e}, HF bf16 {h_t:.2e}; pooled features rms {e_f:.2e} vs {h_f:.2e}, max-norm {m_f:.2e} vs {hm_f:.2e}")
    assert e_t <= 1.25 * h_t + 1e-3 and e_f <= 1.25 * h_f + 1e-3 and m_f <= 1.25 * hm_f + 2e-3
    assert e_f <= 2e-2

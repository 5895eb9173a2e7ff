"""CPU tests for oracle/kv.py: the bf16 RoPE restatement against transformers' own apply_rotary_pos_emb, and
the compaction / re-rotation spec's algebraic properties."""
import numpy as np
import torch

from oracle import kv as okv
from streamvln_b200.kv_arena import delta_rope_tables, hf_rope_tables

HD = 128


def _bits(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def test_rope_matches_transformers_bf16():
    from transformers.models.qwen2.modeling_qwen2 import apply_rotary_pos_emb

    cos, sin = hf_rope_tables(64, HD, 1e6)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, 4, 64, HD, generator=g).bfloat16()
    k = torch.randn(1, 2, 64, HD, generator=g).bfloat16()
    qh, kh = apply_rotary_pos_emb(q, k, cos[None], sin[None])
    for x, want in ((q, qh), (k, kh)):
        n = x.shape[1]
        got = okv.np_rope_hf(_bits(x[0]), np.broadcast_to(_bits(cos)[None], (n, 64, HD)), np.broadcast_to(_bits(sin)[None], (n, 64, HD)))
        assert np.array_equal(got, _bits(want[0]))


def test_rope_tables_match_hf_module():
    from transformers import Qwen2Config
    from transformers.models.qwen2.modeling_qwen2 import Qwen2RotaryEmbedding

    cfg = Qwen2Config(hidden_size=512, num_attention_heads=4, head_dim=HD, rope_theta=1e6, max_position_embeddings=512)
    cfg.rope_parameters = {"rope_type": "default", "rope_theta": 1e6}
    rot = Qwen2RotaryEmbedding(cfg)
    x = torch.zeros(1, 300, 512, dtype=torch.bfloat16)
    cos, sin = rot(x, torch.arange(300)[None])
    c2, s2 = hf_rope_tables(300, HD, 1e6)
    assert torch.equal(cos[0], c2) and torch.equal(sin[0], s2)


def test_compact_spec_properties():
    rng = np.random.default_rng(0)
    arena = okv.f32_to_bf16(rng.standard_normal((1, 2, 2, 2, 64, HD)).astype(np.float32))
    orig = arena.copy()
    cd, sd = [t.numpy() for t in delta_rope_tables(64, HD, 1e6)]
    src = np.zeros(64, np.int32)
    keep = np.array([0, 1, 2, 5, 6, 9, 20, 21, 40])
    src[: keep.size] = keep
    okv.np_kv_compact(arena, 0, 0, keep.size, src, True, cd, sd)
    # V rows are moved verbatim; unmoved K rows (src == dst) are untouched
    assert np.array_equal(arena[0, :, 1, :, : keep.size], orig[0, :, 1][:, :, keep])
    assert np.array_equal(arena[0, :, 0, :, :3], orig[0, :, 0, :, :3])
    # re-rotation preserves the norm of every (lo, hi) pair up to bf16 rounding
    k_new = okv.bf16_to_f32(arena[0, 0, 0, 0, 3: keep.size])
    k_old = okv.bf16_to_f32(orig[0, 0, 0, 0][keep[3:]])
    n_new = np.hypot(k_new[:, :64], k_new[:, 64:])
    n_old = np.hypot(k_old[:, :64], k_old[:, 64:])
    assert np.allclose(n_new, n_old, rtol=0, atol=2 ** -6)
    # rotate forward then back by the same shift: within two bf16 roundings
    back = arena.copy()
    okv.np_rerotate(back, 0, 3, (keep[3:] - np.arange(3, keep.size)), cd, sd)
    err = np.abs(okv.bf16_to_f32(back[0, 0, 0, 0, 3: keep.size]) - k_old).max()
    assert err <= 4 * 2 ** -8 * np.abs(k_old).max()

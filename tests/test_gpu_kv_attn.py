"""GPU parity: KV arena kernels (bit-exact vs oracle/kv.py) and attention kernels (tolerance vs
oracle/attention.py; north_star: max rel err 1e-2), all through the C-ABI."""
import numpy as np
import pytest
import torch

from oracle import attention as oa
from oracle import kv as okv

pytestmark = pytest.mark.gpu

HD = 128


def _bits(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def _from_bits(a, dev):
    return torch.from_numpy(a.view(np.int16)).view(torch.bfloat16).to(dev)


def _arena(cuda, n_env=2, n_layer=2, n_kv=2, max_len=512):
    from streamvln_b200.kv_arena import KvArena

    return KvArena(n_env, n_layer, n_kv, HD, max_len, 1e6, cuda)


def test_append_rope_bit_exact(svln_lib, cuda):
    ar = _arena(cuda, n_env=3, n_layer=2, n_kv=2, max_len=256)
    Hq = 6
    g = torch.Generator().manual_seed(0)
    n = 37
    qkv = (torch.randn(n, (Hq + 4) * HD, generator=g) * 2).bfloat16()
    tok_env = torch.randint(0, 3, (n,), generator=g, dtype=torch.int32)
    # unique (env, pos) pairs
    tok_pos = torch.randperm(200, generator=g)[:n].to(torch.int32)
    want_arena = np.zeros((3, 2, 2, 2, 256, HD), np.uint16)
    q_want = okv.np_append_rope(want_arena, 1, _bits(qkv), Hq, tok_env.numpy(), tok_pos.numpy(),
                                _bits(ar.cos.cpu()), _bits(ar.sin.cpu()))
    q_out = torch.empty(n, Hq, HD, dtype=torch.bfloat16, device=cuda)
    ar.append_rope(1, qkv.to(cuda), Hq, tok_env.to(cuda), tok_pos.to(cuda), q_out)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(q_out), q_want)
    assert np.array_equal(_bits(ar.buf), want_arena)


def test_rope_matches_hf_apply_rotary(cuda):
    """The oracle's bf16 RoPE is HF's apply_rotary_pos_emb evaluated by eager torch in bf16."""
    from transformers.models.qwen2.modeling_qwen2 import apply_rotary_pos_emb
    from streamvln_b200.kv_arena import hf_rope_tables

    cos, sin = hf_rope_tables(64, HD, 1e6)
    q = torch.randn(1, 4, 64, HD).bfloat16()
    k = torch.randn(1, 2, 64, HD).bfloat16()
    qh, kh = apply_rotary_pos_emb(q, k, cos[None], sin[None])
    got = okv.np_rope_hf(_bits(q[0]), np.broadcast_to(_bits(cos)[None], (4, 64, HD)), np.broadcast_to(_bits(sin)[None], (4, 64, HD)))
    assert np.array_equal(got, _bits(qh[0]))


@pytest.mark.parametrize("rerotate", [False, True])
def test_compact_bit_exact(svln_lib, cuda, rerotate):
    ar = _arena(cuda, n_env=3, n_layer=2, n_kv=2, max_len=700)
    g = torch.Generator().manual_seed(1)
    ar.buf.copy_(torch.randn(ar.buf.shape, generator=g).bfloat16())
    base = _bits(ar.buf).copy()
    rng = np.random.default_rng(0)
    jobs = []
    src_all = np.zeros((2, 700), np.int32)
    for j, (env, old_len, first) in enumerate([(2, 650, 40), (0, 300, 0)]):
        keep = np.sort(rng.choice(np.arange(first, old_len), size=(old_len - first) * 2 // 3, replace=False))
        new_len = first + keep.size
        src_all[j, first:new_len] = keep
        jobs.append((env, first, new_len))
        okv.np_kv_compact(base, env, first, new_len, src_all[j], rerotate, ar.cos_delta.cpu().numpy(),
                          ar.sin_delta.cpu().numpy())
    ar.compact(torch.tensor([j[0] for j in jobs], dtype=torch.int32, device=cuda),
               torch.tensor([j[1] for j in jobs], dtype=torch.int32, device=cuda),
               torch.tensor([j[2] for j in jobs], dtype=torch.int32, device=cuda),
               torch.from_numpy(src_all).to(cuda), rerotate)
    torch.cuda.synchronize()
    got = _bits(ar.buf)
    for env, first, new_len in jobs:  # rows beyond new_len are dead and unspecified
        assert np.array_equal(got[env][..., :new_len, :], base[env][..., :new_len, :])
    assert np.array_equal(got[1], base[1])  # untouched stream


def test_rerotate_bit_exact_and_inverse(svln_lib, cuda):
    ar = _arena(cuda, n_env=2, n_layer=2, n_kv=2, max_len=300)
    g = torch.Generator().manual_seed(2)
    ar.buf.copy_(torch.randn(ar.buf.shape, generator=g).bfloat16())
    base = _bits(ar.buf).copy()
    delta = np.random.default_rng(1).integers(-120, 120, size=150).astype(np.int32)
    okv.np_rerotate(base, 1, 20, delta, ar.cos_delta.cpu().numpy(), ar.sin_delta.cpu().numpy())
    ar.rerotate(1, 20, torch.from_numpy(delta).to(cuda))
    torch.cuda.synchronize()
    assert np.array_equal(_bits(ar.buf), base)
    # property: rotating back returns K to within bf16 rounding (2 roundings) of the original
    orig = ar.buf.clone()
    ar.rerotate(1, 20, torch.from_numpy(delta).to(cuda))
    ar.rerotate(1, 20, torch.from_numpy(-delta).to(cuda))
    err = (ar.buf.float() - orig.float()).abs().max().item()
    assert err <= 4 * 2 ** -8 * orig.float().abs().max().item()


def test_rerotate_equals_fresh_rope(svln_lib, cuda):
    """K roped at position p, then moved to p-150 by the compaction kernel (re-rotation by -150), must equal
    K roped directly at p-150 within bf16 rounding."""
    ar = _arena(cuda, n_env=1, n_layer=1, n_kv=2, max_len=400)
    Hq = 2
    g = torch.Generator().manual_seed(3)
    n = 100
    qkv = torch.randn(n, (Hq + 4) * HD, generator=g).bfloat16().to(cuda)
    env = torch.zeros(n, dtype=torch.int32, device=cuda)
    q_out = torch.empty(n, Hq, HD, dtype=torch.bfloat16, device=cuda)
    ar.append_rope(0, qkv, Hq, env, torch.arange(200, 200 + n, dtype=torch.int32, device=cuda), q_out)
    src = torch.zeros(1, 400, dtype=torch.int32, device=cuda)
    src[0, 50:150] = torch.arange(200, 300, dtype=torch.int32, device=cuda)
    ar.compact(torch.zeros(1, dtype=torch.int32, device=cuda), torch.tensor([50], dtype=torch.int32, device=cuda),
               torch.tensor([150], dtype=torch.int32, device=cuda), src, True)
    moved_k = ar.buf[0, 0, 0, :, 50:150].float().clone()
    moved_v = ar.buf[0, 0, 1, :, 50:150].clone()
    ar.append_rope(0, qkv, Hq, env, torch.arange(50, 50 + n, dtype=torch.int32, device=cuda), q_out)
    fresh_k = ar.buf[0, 0, 0, :, 50:150].float()
    assert torch.equal(moved_v, ar.buf[0, 0, 1, :, 50:150])  # V moves unchanged
    tol = 3 * 2 ** -8 * fresh_k.abs().max().item()
    assert (moved_k - fresh_k).abs().max().item() <= tol


def _fill_cache(ar, Hq, lens, seed):
    """Write random rotated K/V for every env through the append kernel; returns nothing."""
    g = torch.Generator().manual_seed(seed)
    for layer in range(ar.n_layer):
        for e, L in enumerate(lens):
            qkv = torch.randn(L, (Hq + 2 * ar.n_kv_head) * HD, generator=g).bfloat16().to(ar.device)
            q_out = torch.empty(L, Hq, HD, dtype=torch.bfloat16, device=ar.device)
            ar.append_rope(layer, qkv, Hq, torch.full((L,), e, dtype=torch.int32, device=ar.device),
                           torch.arange(L, dtype=torch.int32, device=ar.device), q_out)


def _check_attn(out, q, ar, layer, seqs, scale, Hq, tol=1e-2):
    worst = 0.0
    for env, qs, nq, L in seqs:
        want = oa.attention_ref(q[qs:qs + nq].cpu().float(), ar.plane(env, layer, 0).cpu().float(),
                                ar.plane(env, layer, 1).cpu().float(), L, scale)
        got = out[qs:qs + nq].cpu().double()
        err = (got - want).abs().max().item() / want.abs().max().item()
        worst = max(worst, err)
    assert worst <= tol, worst
    return worst


@pytest.mark.parametrize("Hq,Hkv,lens,nqs", [
    (28, 4, [333], [1]),                       # Qwen2-7B GQA 7:1, single stream decode
    (28, 4, [1, 64, 65, 700, 3000], [1, 1, 1, 1, 1]),
    (4, 2, [130, 257], [3, 8]),                # short chunks, several tokens per stream
    (8, 8, [200], [2]),                        # MHA (g = 1)
    (16, 1, [511], [1]),                       # g = 16 fills the MMA tile
])
def test_decode_attention(svln_lib, cuda, Hq, Hkv, lens, nqs):
    from streamvln_b200.kv_arena import KvArena

    ar = KvArena(len(lens), 2, Hkv, HD, 3072, 1e6, cuda)
    _fill_cache(ar, Hq, lens, seed=4)
    g = torch.Generator().manual_seed(5)
    n_tok = sum(nqs)
    q = torch.randn(n_tok, Hq, HD, generator=g).bfloat16().to(cuda)
    qs = np.concatenate([[0], np.cumsum(nqs)[:-1]]).astype(np.int32)
    out = torch.zeros(n_tok, Hq * HD, dtype=torch.bfloat16, device=cuda)
    scale = HD ** -0.5
    for layer in (0, 1):
        ar.attn_decode(layer, q, Hq, torch.arange(len(lens), dtype=torch.int32, device=cuda),
                       torch.from_numpy(qs).to(cuda), torch.tensor(nqs, dtype=torch.int32, device=cuda),
                       torch.tensor(lens, dtype=torch.int32, device=cuda), max(nqs), max(lens), scale, out)
        torch.cuda.synchronize()
        _check_attn(out, q, ar, layer, [(e, int(qs[e]), nqs[e], lens[e]) for e in range(len(lens))], scale, Hq)


@pytest.mark.parametrize("Hq,Hkv,lens,nqs", [
    (28, 4, [216], [216]),                     # first turn: pure causal prefill
    (28, 4, [2000], [217]),                    # steady-state turn against a long prefix
    (28, 4, [1800], [1800]),                   # window start
    (4, 2, [100, 300, 129], [1, 128, 129]),    # ragged batch incl. a single-row chunk
    (8, 8, [640], [77]),
    (28, 4, [900, 1500, 400], [300, 257, 400]),   # 216 (M tile, kv head) items: more than one per SM, ragged key ranges
])
def test_prefill_attention_tcgen05(svln_lib, cuda, Hq, Hkv, lens, nqs):
    from streamvln_b200.kv_arena import KvArena

    ar = KvArena(len(lens), 2, Hkv, HD, 2048, 1e6, cuda)
    _fill_cache(ar, Hq, lens, seed=6)
    g = torch.Generator().manual_seed(7)
    n_tok = sum(nqs)
    q = torch.randn(n_tok, Hq, HD, generator=g).bfloat16().to(cuda)
    qs = [int(x) for x in np.concatenate([[0], np.cumsum(nqs)[:-1]])]
    out = torch.zeros(n_tok, Hq * HD, dtype=torch.bfloat16, device=cuda)
    scale = HD ** -0.5
    ar.attn_prefill(1, q, Hq, list(range(len(lens))), qs, nqs, lens, scale, out)
    torch.cuda.synchronize()
    _check_attn(out, q, ar, 1, [(e, qs[e], nqs[e], lens[e]) for e in range(len(lens))], scale, Hq)


@pytest.mark.parametrize("var,val", [("SVLN_PF_PERSIST", "2"), ("SVLN_PF_SPLITS", "3"), ("SVLN_PF_CLUSTER", "1")])
def test_prefill_attention_opt_in_variants(svln_lib, cuda, var, val):
    """The measured-but-not-default variants of the prefill kernel (persistent CTAs with a double-buffered Q,
    split-KV with a combine pass, 2-CTA cluster with multicast K/V) stay correct: the same parity cases in a child
    process with the switch set (the library reads these switches once)."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, **{var: val})
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k", "test_prefill_attention_tcgen05"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_prefill_equals_decode_path(svln_lib, cuda):
    """The two kernels implement the same function: compare them on a mid-size chunk."""
    from streamvln_b200.kv_arena import KvArena

    Hq, Hkv, L, nq = 28, 4, 900, 40
    ar = KvArena(1, 1, Hkv, HD, 1024, 1e6, cuda)
    _fill_cache(ar, Hq, [L], seed=8)
    q = torch.randn(nq, Hq, HD, generator=torch.Generator().manual_seed(9)).bfloat16().to(cuda)
    o1 = torch.zeros(nq, Hq * HD, dtype=torch.bfloat16, device=cuda)
    o2 = torch.zeros_like(o1)
    scale = HD ** -0.5
    ar.attn_prefill(0, q, Hq, [0], [0], [nq], [L], scale, o1)
    one = torch.ones(1, dtype=torch.int32, device=cuda)
    ar.attn_decode(0, q, Hq, one * 0, one * 0, one * nq, one * L, nq, L, scale, o2)
    torch.cuda.synchronize()
    assert (o1.float() - o2.float()).abs().max().item() <= 2e-2 * o2.float().abs().max().item()


def test_fused_ops(svln_lib, cuda):
    from streamvln_b200 import _lib

    lib = svln_lib
    st = _lib.cur_stream()
    g = torch.Generator().manual_seed(10)
    for n, dim in ((5, 3584), (3, 256), (2, 8192)):
        x = (torch.randn(n, dim, generator=g) * 3).bfloat16().to(cuda)
        r = torch.randn(n, dim, generator=g).bfloat16().to(cuda)
        w = (1 + 0.1 * torch.randn(dim, generator=g)).bfloat16().to(cuda)
        out = torch.empty_like(x)
        _lib.check(lib.svln_rmsnorm(x.data_ptr(), w.data_ptr(), n, dim, 1e-6, out.data_ptr(), st))

        def hf_norm(h):
            v = h.float().pow(2).mean(-1, keepdim=True)
            return w * (h.float() * torch.rsqrt(v + 1e-6)).to(torch.bfloat16)

        want = hf_norm(x)
        assert (out.float() - want.float()).abs().max().item() <= 2 ** -7 * want.float().abs().max().item()
        x2 = x.clone()
        _lib.check(lib.svln_add_rmsnorm(x2.data_ptr(), r.data_ptr(), w.data_ptr(), n, dim, 1e-6, out.data_ptr(), st))
        assert torch.equal(x2, x + r)
        want = hf_norm(x + r)
        assert (out.float() - want.float()).abs().max().item() <= 2 ** -7 * want.float().abs().max().item()
    gu = torch.randn(7, 2 * 18944, generator=g).bfloat16().to(cuda)
    out = torch.empty(7, 18944, dtype=torch.bfloat16, device=cuda)
    _lib.check(lib.svln_swiglu(gu.data_ptr(), 7, 18944, out.data_ptr(), st))
    want = torch.nn.functional.silu(gu[:, :18944]) * gu[:, 18944:]
    assert (out.float() - want.float()).abs().max().item() <= 2 ** -7 * want.float().abs().max().item()
    logits = torch.randn(3, 152064, generator=g).to(cuda)
    logits[1, 77] = logits[1, 99999] = 50.0  # tie -> lowest index
    idx = torch.empty(3, dtype=torch.int64, device=cuda)
    _lib.check(lib.svln_argmax(logits.data_ptr(), 1, 3, 152064, idx.data_ptr(), st))
    assert torch.equal(idx, logits.argmax(-1)) and idx[1].item() == 77
    lb = logits.bfloat16()
    _lib.check(lib.svln_argmax(lb.data_ptr(), 0, 3, 152064, idx.data_ptr(), st))
    assert torch.equal(idx, lb.float().argmax(-1))


def test_gemv_matches_linear(svln_lib, cuda):
    """svln_gemv == F.linear in bf16 (fp32 accumulation, single rounding) within bf16 tolerance, for the Qwen2-7B
    layer shapes and M = 1..4."""
    from streamvln_b200 import _lib

    lib = svln_lib
    g = torch.Generator().manual_seed(11)
    for N, K in ((4608, 3584), (3584, 3584), (2 * 18944, 3584), (3584, 18944), (1000, 256)):
        W = (torch.randn(N, K, generator=g) * 0.02).bfloat16().to(cuda)
        bias = (torch.randn(N, generator=g) * 0.02).bfloat16().to(cuda)
        for M in (1, 2, 3, 4):
            x = torch.randn(M, K, generator=g).bfloat16().to(cuda)
            for b in (None, bias):
                for f32 in (0, 1):
                    out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=cuda)
                    _lib.check(lib.svln_gemv(W.data_ptr(), x.data_ptr(), None if b is None else b.data_ptr(), out.data_ptr(),
                                             M, N, K, f32, _lib.cur_stream()))
                    want = x.double() @ W.double().t() + (0 if b is None else b.double())
                    err = (out.double() - want).abs().max().item() / want.abs().max().item()
                    assert err <= 2 ** -8, (N, K, M, f32, err)
                    if f32:   # fp32 logits carry bf16-rounded values (HF's lm_head is a bf16 Linear upcast with .float())
                        assert torch.equal(out, out.bfloat16().float())
                        assert torch.equal(out.bfloat16(), prev_bf16) or \
                            (out.double() - prev_bf16.double()).abs().max().item() <= 2 ** -7 * want.abs().max().item()
                    else:
                        prev_bf16 = out.clone()
    assert lib.svln_gemv(W.data_ptr(), x.data_ptr(), None, out.data_ptr(), 5, N, K, 0, None) == -34

"""GPU parity of the whole reference-semantics turn: product (`StreamVLNForCausalLM.generate`,
`VLNEvaluator.step`) vs the CPU oracle (reference glue restated + stock transformers Qwen2ForCausalLM).
Tolerance (north_star): logits max rel err <= 1e-2, identical greedy tokens / actions."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import glue

pytestmark = pytest.mark.gpu

TOL = 1e-2


def _hf_from(weights, dtype=torch.float32):
    from transformers import Qwen2Config, Qwen2ForCausalLM

    c = weights.cfg
    cfg = Qwen2Config(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size, num_hidden_layers=c.num_hidden_layers,
                      num_attention_heads=c.num_attention_heads, num_key_value_heads=c.num_key_value_heads,
                      head_dim=c.head_dim, vocab_size=c.vocab_size, max_position_embeddings=c.max_position_embeddings,
                      rope_theta=c.rope_theta, rms_norm_eps=c.rms_norm_eps, tie_word_embeddings=False)
    cfg.rope_parameters = {"rope_type": "default", "rope_theta": c.rope_theta}
    lm = Qwen2ForCausalLM(cfg)
    missing, unexpected = lm.load_state_dict({k: v.to(dtype) for k, v in weights.to_hf_state_dict().items()}, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    return lm.to(dtype).eval()


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / b.double().abs().max().item()


@pytest.fixture(scope="module")
def tiny(svln_lib, cuda):
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM

    cfg = StreamVLNConfig.tiny(vocab_size=512, layers=3)
    m = StreamVLNForCausalLM(cfg, n_env=2, device=cuda, seed=0, lm_head_std=0.3)
    return m, _hf_from(m.weights)


def test_engine_prefill_and_decode_logits(tiny, cuda):
    from streamvln_b200.qwen2 import Chunk

    m, lm = tiny
    eng = m.engine
    g = torch.Generator().manual_seed(0)
    D = m.config.llm.hidden_size
    emb = (torch.randn(300, D, generator=g) * 0.05).bfloat16()
    for env in (0, 1):
        eng.arena.reset_env(env)
    # two streams prefilled together (ragged: 300 and 77 tokens), then a second chunk on stream 0
    logits = eng.forward(torch.cat([emb, emb[:77]]).to(cuda), [Chunk(0, 0, 300, 300), Chunk(1, 300, 77, 77)])
    from transformers import DynamicCache

    c0, c1 = DynamicCache(config=lm.config), DynamicCache(config=lm.config)
    with torch.no_grad():
        w0 = lm(inputs_embeds=emb[None].float(), past_key_values=c0, use_cache=True).logits[0, -1]
        w1 = lm(inputs_embeds=emb[None, :77].float(), past_key_values=c1, use_cache=True).logits[0, -1]
    assert _rel(logits[0].cpu(), w0) <= TOL and _rel(logits[1].cpu(), w1) <= TOL
    logits = eng.forward(emb[100:140].clone().to(cuda), [Chunk(0, 0, 40, 340)])
    with torch.no_grad():
        w0 = lm(inputs_embeds=emb[None, 100:140].float(), past_key_values=c0, use_cache=True).logits[0, -1]
    assert _rel(logits[0].cpu(), w0) <= TOL
    # decode: both streams step together, with and without the CUDA graph
    tok = torch.tensor([5, 9], device=cuda)
    for use_graph in (False, True):
        runner = eng._decode_runner([0, 1], use_graph)
        t = tok
        for _ in range(3):
            t = runner.step(t)
        runner.finish()
        with torch.no_grad():
            tt = [tok[0:1].cpu(), tok[1:2].cpu()]
            for _ in range(3):
                tt[0] = lm(input_ids=tt[0][None], past_key_values=c0, use_cache=True).logits[0, -1:].argmax(-1)
                tt[1] = lm(input_ids=tt[1][None], past_key_values=c1, use_cache=True).logits[0, -1:].argmax(-1)
        assert t.cpu().tolist() == [int(tt[0]), int(tt[1])]
        assert eng.arena.lens[0] == c0.get_seq_length() and eng.arena.lens[1] == c1.get_seq_length()


def test_fused_decode_stack_matches_hf(tiny, cuda):
    """svln_decode_stack_step (one persistent launch per token) vs stock transformers, eager and CUDA-graphed:
    fp32 logits of every step within the tolerance, tokens greedy under the oracle's logits."""
    from transformers import DynamicCache
    from streamvln_b200.qwen2 import Chunk

    m, lm = tiny
    eng = m.engine
    assert eng.use_stack, "the fused decode stack must be the default single-stream decode path"
    g = torch.Generator().manual_seed(3)
    emb = (torch.randn(150, m.config.llm.hidden_size, generator=g) * 0.05).bfloat16()
    eng.keep_decode_logits = True
    eng._graphs.clear()
    try:
        eng.arena.reset_env(0)
        logits = eng.forward(emb.to(cuda), [Chunk(0, 0, 150, 150)])
        c0 = DynamicCache(config=lm.config)
        with torch.no_grad():
            want = lm(inputs_embeds=emb[None].float(), past_key_values=c0, use_cache=True).logits[0, -1]
        assert _rel(logits[0].cpu(), want) <= TOL
        tok = logits.argmax(-1)
        for use_graph in (False, True):
            runner = eng._decode_runner([0], use_graph)
            state = runner.graph.state if runner.graph is not None else runner.state
            for _ in range(4):
                fed = tok.clone()
                tok = runner.step(tok)
                assert state.stack is not None, "decode did not go through svln_decode_stack_step"
                with torch.no_grad():
                    want = lm(input_ids=fed.cpu()[None], past_key_values=c0, use_cache=True).logits[0, -1]
                got = state.stack_logits[0].cpu()
                assert _rel(got, want) <= TOL
                assert int(tok) == int(got.argmax())                       # the kernel's own argmax
                top2 = want.topk(2).values
                assert int(tok) == int(want.argmax()) or float(top2[0] - want[int(tok)]) <= 2e-2 * float(want.abs().max())
            runner.finish()
            assert eng.arena.lens[0] == c0.get_seq_length()
    finally:
        eng.keep_decode_logits = False
        eng._graphs.clear()


@pytest.mark.parametrize("n_ctx", [1, 63, 64, 65, 1000])
def test_fused_decode_stack_context_lengths(tiny, cuda, n_ctx):
    """Edge cases of the in-kernel split-KV attention: a cache shorter than one split, exactly on / around a tile
    boundary, and many tiles; logits of every step within the tolerance of stock transformers."""
    from transformers import DynamicCache
    from streamvln_b200.qwen2 import Chunk

    m, lm = tiny
    eng = m.engine
    g = torch.Generator().manual_seed(100 + n_ctx)
    emb = (torch.randn(n_ctx, m.config.llm.hidden_size, generator=g) * 0.05).bfloat16()
    eng.keep_decode_logits = True
    eng._graphs.clear()
    try:
        eng.arena.reset_env(0)
        tok = eng.forward(emb.clone().to(cuda), [Chunk(0, 0, n_ctx, n_ctx)]).argmax(-1)
        c0 = DynamicCache(config=lm.config)
        with torch.no_grad():
            lm(inputs_embeds=emb[None].float(), past_key_values=c0, use_cache=True)
        runner = eng._decode_runner([0], True)
        for _ in range(3):
            fed = tok.clone()
            tok = runner.step(tok)
            with torch.no_grad():
                want = lm(input_ids=fed.cpu()[None], past_key_values=c0, use_cache=True).logits[0, -1]
            assert int(tok) >= 0, "a grid barrier of the fused stack timed out"
            assert _rel(runner.graph.state.stack_logits[0].cpu(), want) <= TOL
        runner.finish()
        assert eng.arena.lens[0] == c0.get_seq_length() == n_ctx + 3
    finally:
        eng.keep_decode_logits = False
        eng._graphs.clear()


def test_fused_decode_stack_agrees_with_launch_per_op_path(tiny, cuda):
    """Same state, same fed tokens: every token the unfused kernel sequence picks is greedy (up to a near-tie
    margin: accumulation orders differ) under the fused stack's fp32 logits."""
    from streamvln_b200.qwen2 import Chunk

    m, _ = tiny
    eng = m.engine
    g = torch.Generator().manual_seed(4)
    emb = (torch.randn(700, m.config.llm.hidden_size, generator=g) * 0.05).bfloat16().to(cuda)
    try:
        eng.use_stack = False
        eng._graphs.clear()
        eng.arena.reset_env(1)
        first = eng.forward(emb.clone(), [Chunk(1, 0, 700, 700)]).argmax(-1)   # forward consumes h
        runner = eng._decode_runner([1], True)
        tok, ref = first, []
        for _ in range(6):
            tok = runner.step(tok)
            ref.append(int(tok))
        runner.finish()

        eng.use_stack, eng.keep_decode_logits = True, True
        eng._graphs.clear()
        eng.arena.reset_env(1)
        assert int(eng.forward(emb.clone(), [Chunk(1, 0, 700, 700)]).argmax(-1)) == int(first)
        runner = eng._decode_runner([1], True)
        fed, exact = first, 0
        for want in ref:
            got = runner.step(fed)
            lg = runner.graph.state.stack_logits[0]
            assert float(lg.max() - lg[want]) <= 2e-2 * float(lg.abs().max()), (want, int(got))
            exact += int(got) == want
            fed = torch.tensor([want], device=cuda)
        runner.finish()
        assert exact >= len(ref) - 2
    finally:
        eng.use_stack, eng.keep_decode_logits = True, False
        eng._graphs.clear()


def test_generate_matches_oracle_over_a_window_boundary(tiny, cuda):
    m, lm = tiny
    m.reset(2)
    o = glue.ReferenceOracle(lm, num_history=m.model.num_history)
    g = torch.Generator().manual_seed(1)
    frames = (torch.randn(10, 3, 384, 384, generator=g) * 0.5)
    feats = m.vision(frames.to(cuda))                      # [10,196,D] bf16, shared by both sides
    feats_cpu = feats.float().cpu()
    IMG, MEM = glue.IMAGE_TOKEN_INDEX, glue.MEMORY_TOKEN_INDEX
    n_gen = 5
    past_p = past_o = None
    out_ids = None
    worst = 0.0
    # three ordinary turns (V=1), then a window start with num_history=8 memory views + current frame
    plan = [([7, 8, 9, IMG, 10, 11], [0], [[0]]), (None, [1], [[0, 1, 2, 3, 4]]), (None, [2], [[0, 1, 2, 3, 4, 5, 6, 7, 8]])]
    for ids, views, time_ids in plan:
        if ids is None:
            ids = out_ids + [20, IMG, 21, 22]
        r = m.generate(inputs=torch.tensor([ids]), images=frames[views][None], env_id=1, time_ids=time_ids,
                       frame_ids=views, do_sample=False, num_beams=1, max_new_tokens=n_gen, use_cache=True,
                       return_dict_in_generate=True, past_key_values=past_p)
        got = r.sequences.cpu()[0].tolist()
        seq_o, past_o, logits_o = o.generate(ids, feats_cpu[views], time_ids, 0, past_o, n_gen, force_tokens=got)
        past_p = r.past_key_values
        assert r.sequences.shape == (1, n_gen)
        ok, n_exact, gap = glue.greedy_consistent(o.step_logits, got)
        assert ok and n_exact >= n_gen - 1, (got, gap)
        assert past_p.get_seq_length() == past_o.get_seq_length()
        out_ids = got
    # window boundary: caller drops the cache (streamvln_agent.py:192-198)
    m.reset_for_env(1)
    o.reset_for_env(0)
    ids = [7, 8, 9, 30, MEM, 31, IMG, 10, 11]
    views = [0, 1, 2, 3, 4, 5, 6, 7, 9]
    r = m.generate(inputs=torch.tensor([ids]), images=frames[views][None], env_id=1, time_ids=[[32]], frame_ids=views,
                   max_new_tokens=n_gen, past_key_values=None)
    got = r.sequences.cpu()[0].tolist()
    seq_o, past_o, logits_o = o.generate(ids, feats_cpu[views], [[32]], 0, None, n_gen, force_tokens=got)
    ok, n_exact, gap = glue.greedy_consistent(o.step_logits, got)
    assert ok and n_exact >= n_gen - 1, (got, gap)
    assert r.past_key_values.get_seq_length() == past_o.get_seq_length() == len(ids) - 2 + 9 * 196 + n_gen - 1


def test_first_step_logits_tolerance(tiny, cuda):
    """Logits of a window-start prefill (1.7k tokens through the tcgen05 kernel) within 1e-2 of the oracle."""
    from streamvln_b200.qwen2 import Chunk

    m, lm = tiny
    m.reset(2)
    feats = (torch.randn(9, 196, m.config.llm.hidden_size, generator=torch.Generator().manual_seed(2)) * 0.05).bfloat16()
    ids = [7, 8, 9, 30, glue.MEMORY_TOKEN_INDEX, 31, glue.IMAGE_TOKEN_INDEX, 10, 11]
    emb = m.splice(ids, feats[8:].to(cuda), feats[:8].reshape(1, -1, feats.shape[-1]).to(cuda))
    want_emb = glue.splice(ids, m.weights.embed.float().cpu(), feats[8:].float(), feats[:8].reshape(1, -1, feats.shape[-1]).float())
    assert torch.equal(emb.float().cpu(), want_emb)       # the splice kernel is a pure gather: bit-exact
    logits = m.engine.forward(emb, [Chunk(0, 0, emb.shape[0], emb.shape[0])])
    with torch.no_grad():
        want = lm(inputs_embeds=want_emb[None]).logits[0, -1]
    assert _rel(logits[0].cpu(), want) <= TOL
    assert int(logits[0].argmax()) == int(want.argmax())


def test_agent_actions_match_oracle_agent(tiny, cuda):
    """40 env steps (model turn every 4th frame, window reset at 32) through VLNEvaluator.step."""
    from streamvln_b200.agent import VLNEvaluator
    from streamvln_b200.synth import intrinsic_matrix
    from streamvln_b200.tokenizer import StubQwenTokenizer

    m, lm = tiny
    m.reset(1)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, max_new_tokens=4)
    ev = VLNEvaluator({"camera_intrinsic": intrinsic_matrix()}, model=m, tokenizer=StubQwenTokenizer(512), args=args)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 255, size=(40, 480, 640, 3), dtype=np.uint8)
    got, ids = [], []
    for s in range(40):
        a, _, txt = ev.step(0, frames[s], "walk past the sofa and stop at the door", run_model=(s % 4 == 0))
        got.append(a)
        if a is not None:
            ids.append(ev.output_ids.cpu()[0].tolist())
        ev.step_id += 1
    # the oracle agent replays the episode teacher-forced with the product's ids (same history through near-ties):
    # identical actions, and every product token greedy under the oracle's logits
    o = glue.ReferenceOracle(lm, num_history=8)
    oa = glue.OracleAgent(o, StubQwenTokenizer(512), lambda i: m.feat_cache[0][i].float().cpu(), 32, 4, 8, 4)
    want, turn, exact = [], 0, 0
    for s in range(40):
        run = s % 4 == 0
        want.append(oa.step(0, "walk past the sofa and stop at the door", run_model=run, force_tokens=ids[turn] if run else None))
        if run:
            ok, n_exact, gap = glue.greedy_consistent(oa.step_logits_log[-1], ids[turn])
            assert ok, (turn, ids[turn], gap)
            exact += n_exact
            turn += 1
        oa.step_id += 1
    assert got == want
    assert sum(a is not None for a in got) == 10 and exact >= 35


def test_gpu_image_processor_matches_pil(svln_lib, cuda):
    """svln_preprocess_rgb == the reference's PIL image processor, bit for bit."""
    from PIL import Image
    from streamvln_b200.preprocess import GpuImageProcessor
    from streamvln_b200.vision import SigLipImageProcessor

    rng = np.random.default_rng(0)
    proc, gpu = SigLipImageProcessor(), GpuImageProcessor(cuda)
    for h, w in ((480, 640), (720, 1280), (384, 384), (97, 131)):
        frames = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
        frames[0, : h // 2] = np.linspace(0, 255, w)[None, :, None].astype(np.uint8)
        got = gpu(torch.from_numpy(frames).to(cuda)).cpu()
        for i in range(2):
            want = proc.preprocess(images=Image.fromarray(frames[i]), return_tensors="pt")["pixel_values"][0]
            assert torch.equal(got[i], want), (h, w, i)


def test_pool_bilinear_kernel_matches_reference_pool(svln_lib, cuda):
    """svln_pool_bilinear is BIT-EXACT against glue.pool_bilinear_spec (the stated fp32 arithmetic), and within one bf16
    ulp on at most 2e-4 of the elements of the oracle's get_2dPool restatement (CPU F.interpolate, pinned against the
    reference's own code by tests/golden/ref_glue_*.npz) - ATen's fused-multiply-add pattern is not reproducible.
    Shapes: the SigLIP grid and the sweep grids."""
    from streamvln_b200.vision import pool_2d

    g = torch.Generator().manual_seed(5)
    for side, dim, frames in ((27, 3584, 2), (27, 256, 9), (24, 512, 1), (20, 64, 3), (14, 128, 2), (5, 8, 1)):
        x = (torch.randn(frames, side * side, dim, generator=g) * 3).bfloat16()
        want = glue.pool_2d(x, side, 2, "bilinear")                     # bf16 in, fp32 opmath, bf16 out
        got = pool_2d(x.to(cuda), side, 2, "bilinear").cpu()
        assert got.shape == want.shape == (frames, ((side + 1) // 2) ** 2, dim)
        assert torch.equal(got, glue.pool_bilinear_spec(x, side, 2)), (side, dim)   # bit-exact against the stated arithmetic
        diff = (got.float() - want.float()).abs()
        assert float((diff > 0).float().mean()) <= 2e-4, (side, dim, float((diff > 0).float().mean()))
        bound = torch.maximum(want.float().abs(), got.float().abs()) * 2.0 ** -7   # one ulp of the larger binade
        # + an absolute floor for results that are tiny by cancellation (fp32 rounding of operands of magnitude ~10)
        assert bool((diff <= bound * 1.001 + 1e-5).all()), (side, dim, float(diff.max()))


def test_gpu_depth_processor_matches_reference_host_path(svln_lib, cuda):
    """svln_preprocess_depth == the Habitat evaluator's host depth path, bit for bit: against the golden fixture produced
    by the reference's own preprocess_depth_image (tests/golden/ref_geometry.npz) and against synth.preprocess_depth on
    other shapes."""
    from pathlib import Path
    from streamvln_b200 import synth
    from streamvln_b200.preprocess import GpuDepthProcessor

    z = np.load(Path(__file__).resolve().parent / "golden" / "ref_geometry.npz")
    proc = GpuDepthProcessor(cuda)
    got = proc(torch.from_numpy(z["depth01"]).to(cuda))[0].cpu().numpy()
    assert np.array_equal(got, z["depth384"].astype(np.float32))
    rng = np.random.default_rng(3)
    for h, w in ((480, 640), (256, 256), (97, 131)):
        d = rng.uniform(0, 1, size=(2, h, w)).astype(np.float32)
        d[0, : h // 3] = 0.0
        got = proc(torch.from_numpy(d).to(cuda)).cpu().numpy()
        for i in range(2):
            assert np.array_equal(got[i], synth.preprocess_depth(d[i])), (h, w, i)


def test_from_pretrained_round_trip(tiny, cuda, tmp_path):
    """save_pretrained -> from_pretrained (HF layout: config.json + model.safetensors with the reference's parameter names,
    streamvln_agent.py:281-293) reproduces the model bit for bit: identical weights, identical greedy ids on a turn."""
    from streamvln_b200.model import StreamVLNForCausalLM

    m, _ = tiny
    m.save_pretrained(tmp_path / "ckpt")
    m2 = StreamVLNForCausalLM.from_pretrained(tmp_path / "ckpt", device=cuda, n_env=1)
    assert m2.config.llm == m.config.llm and m2.config.vision == m.config.vision
    for a, b in zip(m.weights.layers, m2.weights.layers):
        assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(m.weights.lm_head, m2.weights.lm_head) and torch.equal(m.weights.final_norm, m2.weights.final_norm)
    assert all(torch.equal(p, q) for p, q in zip(m.vision.parameters(), m2.vision.parameters()))
    frames = torch.randn(1, 1, 3, 384, 384, generator=torch.Generator().manual_seed(9)) * 0.5
    ids = torch.tensor([[7, 8, glue.IMAGE_TOKEN_INDEX, 10, 11]])
    m.reset(2)
    r1 = m.generate(inputs=ids, images=frames, env_id=0, time_ids=[[0]], max_new_tokens=4, past_key_values=None)
    r2 = m2.generate(inputs=ids, images=frames, env_id=0, time_ids=[[0]], max_new_tokens=4, past_key_values=None)
    assert r1.sequences.tolist() == r2.sequences.tolist()


def test_voxel_memory_window_start_matches_oracle(svln_lib, cuda):
    """memory="voxel": at a window start the <memory> block is every past model-step frame's pooled tokens, pruned by
    svln_voxel_prune (keys -> keep-latest -> compaction -> row gather).  72 env steps = two window starts (steps 32, 64)
    over 8 and 16 past frames.  The spliced rows are BIT-EXACT against the C restatement on the same features and
    geometry; the turn's ids are greedy under the oracle's logits with that block; ordinary turns are untouched."""
    from streamvln_b200 import synth
    from streamvln_b200.agent import VLNEvaluator
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM
    from streamvln_b200.tokenizer import StubQwenTokenizer

    cfg = StreamVLNConfig.tiny(vocab_size=512, layers=3)
    cfg.memory, cfg.memory_cap = "voxel", 1568
    m = StreamVLNForCausalLM(cfg, n_env=1, device=cuda, seed=0, lm_head_std=0.3)
    lm = _hf_from(m.weights)
    m.reset(1)
    n_steps = 72
    st = synth.make_stream(n_steps, with_feat=False, seed=4)
    args = SimpleNamespace(num_frames=32, num_future_steps=4, num_history=8, max_new_tokens=4)
    ev = VLNEvaluator({"camera_intrinsic": synth.intrinsic_matrix()}, model=m, tokenizer=StubQwenTokenizer(512), args=args)
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 255, size=(n_steps, 480, 640, 3), dtype=np.uint8)
    got, ids, mems = [], [], {}
    for s in range(n_steps):
        a, _, _ = ev.step(0, frames[s], "leave the kitchen and wait by the stairs", run_model=(s % 4 == 0),
                          depth=st["depth"][s], pose=st["pose"][s], intrinsic=st["intr"][s])
        got.append(a)
        if a is not None:
            ids.append(ev.output_ids.cpu()[0].tolist())
            if s in (32, 64):
                mems[s] = m.last_memory[0].cpu()
        ev.step_id += 1
    G = cfg.vision.pooled_per_side

    def memory_fn(before_frame):
        past = [f for f in range(0, before_frame, 4)]
        feats = torch.stack([m.feat_cache[0][f].cpu() for f in past])
        return glue.voxel_memory(feats, st["depth"][past], st["pose"][past], st["intr"][past], G, cfg.voxel_size, cfg.max_depth,
                                 cfg.invalid_policy, cfg.memory_cap)

    for s in (32, 64):
        want = memory_fn(s)
        assert mems[s].shape == want.shape and torch.equal(mems[s].view(torch.int16), want.view(torch.int16)), s
    n8, n16 = mems[32].shape[1], mems[64].shape[1]
    assert 0 < n8 < min(8 * G * G, cfg.memory_cap) and n16 == cfg.memory_cap          # the second start hits the cap: the oldest survivors go
    print(f"\n  voxel memory: {n8} of {8 * G * G} tokens kept at step 32, {n16} (the cap) of {16 * G * G} at step 64")
    o = glue.ReferenceOracle(lm, num_history=8)
    oa = glue.OracleAgent(o, StubQwenTokenizer(512), lambda i: m.feat_cache[0][i].float().cpu(), 32, 4, 8, 4)
    oa.memory_fn = lambda f: memory_fn(f).float()
    want, turn = [], 0
    for s in range(n_steps):
        run = s % 4 == 0
        want.append(oa.step(0, "leave the kitchen and wait by the stairs", run_model=run, force_tokens=ids[turn] if run else None))
        if run:
            ok, _, gap = glue.greedy_consistent(oa.step_logits_log[-1], ids[turn])
            assert ok, (turn, ids[turn], gap)
            turn += 1
        oa.step_id += 1
    assert got == want
    # the mode needs geometry with every request
    m.reset(1)
    ev2 = VLNEvaluator({"camera_intrinsic": synth.intrinsic_matrix()}, model=m, tokenizer=StubQwenTokenizer(512), args=args)
    req = ev2.prepare(0, frames[0], "go", run_model=True)
    req["depths"] = None
    with pytest.raises(ValueError):
        m.generate(**req)

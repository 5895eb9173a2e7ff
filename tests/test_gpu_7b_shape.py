"""GPU parity AT THE BENCHMARKED SHAPE (BASELINE.json configs[1]): hidden 3584 / inter 18944 / 28 q : 4 kv heads /
vocab 152064, 2 decoder layers, RANDOM RMSNorm weights and biases.  What is checked is the LLM arithmetic the product
substitutes for HF `Qwen2ForCausalLM.forward` + flash-attn 2 + `DynamicCache` + the greedy loop
(streamvln/model/stream_video_vln.py:340,402; streamvln_eval.py:526): the tcgen05 prefill path, the fused decode stack
(`svln_decode_stack_step`, 148-CTA row maps at 7B sizes), the `svln_gemv` path (2-4 streams), the packed cuBLAS path
(>4 streams), the EOS stop with its arena-length roll-back, and mixed `max_new_tokens` in one packed batch.

Checkers, both on the same GPU, same weights:
  * `lm32` - stock transformers Qwen2ForCausalLM in fp32 (TF32 off): the ground truth.
  * `lm16` - the same model in bf16 with sdpa attention - the reference's own arithmetic (torch_dtype=bfloat16,
    streamvln_eval.py:523-527) and therefore the yardstick for what "bf16 within tolerance" can mean.

MEASURED (B200, scripts/parity_probe_7b.py, profiles/r2_parity_7b_shape.md): with a 152 064-way logit row the error
metric max|d| / max|ref| is a maximum over 152k rounding-noise samples, and stock HF bf16 itself sits at 1.5e-2 .. 2.1e-2
of fp32 (per element above half the maximum: 2.3e-2 .. 3.5e-2).  The product measures the same (1.5e-2 .. 2.1e-2 /
2.4e-2 .. 3.7e-2) and is within 0.9e-2 .. 1.2e-2 of HF bf16.  north_star's "max rel err 1e-2" is met by neither at this
vocabulary size (it is met at vocab 512, tests/test_gpu_model.py); what `_check_logits` ENFORCES here is therefore
  (a) product-vs-fp32 <= max(1e-2, 1.2 x HF-bf16-vs-fp32 + 2e-3), and <= 2.5e-2 absolutely   (whole row, max-norm);
  (b) per element, for every logit with |ref| >= 0.5 max|ref| (the ones that can win the argmax): relative error
      <= max(1e-2, 1.5 x HF-bf16's + 3e-3), and <= 5e-2 absolutely (a maximum over ~4k samples fluctuates more);
  (c) product-vs-HF-bf16 <= 1.5e-2 (whole row);
  (d) the product's greedy token is HF bf16's, or within two bf16 ulps of its top logit (`_greedy_ok`).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-2
FLOOR = 0.5      # per-element relative check applies to logits with |want| >= FLOOR * max|want|
STATS = []       # (what, max|d|/max|ref|, worst per-element rel above the floor): printed by the last test


def _hf(weights, dtype, device):
    from transformers import Qwen2Config, Qwen2ForCausalLM

    c = weights.cfg
    cfg = Qwen2Config(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size, num_hidden_layers=c.num_hidden_layers,
                      num_attention_heads=c.num_attention_heads, num_key_value_heads=c.num_key_value_heads,
                      head_dim=c.head_dim, vocab_size=c.vocab_size, max_position_embeddings=c.max_position_embeddings,
                      rope_theta=c.rope_theta, rms_norm_eps=c.rms_norm_eps, tie_word_embeddings=False)
    cfg.rope_parameters = {"rope_type": "default", "rope_theta": c.rope_theta}
    with torch.device(device):
        lm = Qwen2ForCausalLM(cfg)
    # `.to(bf16)` would also round the rotary inv_freq buffer (from_pretrained(torch_dtype=bf16) keeps it fp32): keep it
    rot = lm.model.rotary_emb
    keep = {n: getattr(rot, n).clone() for n in ("inv_freq", "original_inv_freq") if hasattr(rot, n)}
    lm = lm.to(dtype)
    for n, v in keep.items():
        setattr(rot, n, v)
    missing, unexpected = lm.load_state_dict({k: v.to(dtype) for k, v in weights.to_hf_state_dict(device).items()}, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    return lm.eval()


def _errs(a: torch.Tensor, ref: torch.Tensor):
    """(max|a - ref| / max|ref|, worst per-element relative error over |ref| >= FLOOR max|ref|)"""
    a, ref = a.double(), ref.double()
    scale = ref.abs().max()
    err = (a - ref).abs()
    big = ref.abs() >= FLOOR * scale
    return (err.max() / scale).item(), (err[big] / ref.abs()[big]).max().item()


def _check_logits(got: torch.Tensor, want32: torch.Tensor, want16: torch.Tensor, what=""):
    p_all, p_big = _errs(got, want32)
    h_all, h_big = _errs(want16.float(), want32)
    ph_all, _ = _errs(got, want16.float())
    STATS.append((what, p_all, p_big, h_all, h_big, ph_all))
    assert p_all <= max(TOL, 1.2 * h_all + 2e-3) and p_all <= 2.5e-2, (what, "whole row vs fp32", p_all, "HF bf16:", h_all)
    assert p_big <= max(TOL, 1.5 * h_big + 3e-3) and p_big <= 5e-2, (what, "per element above the floor", p_big, "HF bf16:", h_big)
    assert ph_all <= 1.5e-2, (what, "vs HF bf16", ph_all)


def _greedy_ok(tok: int, want16: torch.Tensor) -> bool:
    """tok is the bf16 reference's pick, or within two bf16 ulps of its top logit (each side rounds its logits to bf16
    and accumulates in a different order, so a one-ulp difference per logit is the noise floor of the comparison)."""
    import math

    top = float(want16.max())
    ulp = 2.0 ** (math.floor(math.log2(abs(top))) - 7) if top != 0 else 0.0
    return int(want16.argmax()) == tok or top - float(want16[tok]) <= 2 * ulp * 1.001


@pytest.fixture(scope="module")
def big(svln_lib, cuda):
    from streamvln_b200.qwen2 import Qwen2Config, Qwen2Engine, Qwen2Weights

    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = Qwen2Config(num_hidden_layers=2)          # every other dimension is Qwen2-7B's
    assert (cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.vocab_size) == \
        (3584, 18944, 28, 4, 152064)
    w = Qwen2Weights.random(cfg, cuda, seed=11, lm_head_std=0.05, norm_std=0.25)
    assert float((w.layers[0]["ln1"].float() - 1).abs().max()) > 0.3 and float(w.layers[1]["bqkv"].float().abs().max()) > 0
    eng = Qwen2Engine(cfg, w, n_env=6, device=cuda)
    lm32 = _hf(w, torch.float32, cuda)
    lm16 = _hf(w, torch.bfloat16, cuda)
    return eng, lm32, lm16


def _emb(n, D, seed, cuda):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, D, generator=g) * 0.05).bfloat16().to(cuda)


def test_prefill_logits_at_7b_dims(big, cuda):
    """Window-start-sized prefill (1.9k tokens) then a steady-turn chunk (217 tokens over the cache), two streams packed
    raggedly: last-row logits vs fp32 HF, per element."""
    from transformers import DynamicCache
    from streamvln_b200.qwen2 import Chunk

    eng, lm32, lm16 = big
    D = eng.cfg.hidden_size
    e0, e1 = _emb(1900, D, 1, cuda), _emb(217, D, 2, cuda)
    for env in range(2):
        eng.arena.reset_env(env)
    logits = eng.forward(torch.cat([e0, e0[:333]]), [Chunk(0, 0, 1900, 1900), Chunk(1, 1900, 333, 333)])
    c0, c1, h0, h1 = (DynamicCache(config=lm32.config) for _ in range(4))
    with torch.no_grad():
        w0 = lm32(inputs_embeds=e0[None].float(), past_key_values=c0, use_cache=True).logits[0, -1]
        w1 = lm32(inputs_embeds=e0[None, :333].float(), past_key_values=c1, use_cache=True).logits[0, -1]
        b0 = lm16(inputs_embeds=e0[None], past_key_values=h0, use_cache=True).logits[0, -1]
        b1 = lm16(inputs_embeds=e0[None, :333], past_key_values=h1, use_cache=True).logits[0, -1]
    _check_logits(logits[0], w0, b0, "prefill 1900")
    _check_logits(logits[1], w1, b1, "prefill 333 (packed)")
    assert _greedy_ok(int(logits[0].argmax()), b0.float()) and _greedy_ok(int(logits[1].argmax()), b1.float())
    logits = eng.forward(e1.clone(), [Chunk(0, 0, 217, 2117)])
    with torch.no_grad():
        w0 = lm32(inputs_embeds=e1[None].float(), past_key_values=c0, use_cache=True).logits[0, -1]
        b0 = lm16(inputs_embeds=e1[None], past_key_values=h0, use_cache=True).logits[0, -1]
    _check_logits(logits[0], w0, b0, "chunk 217 over 1900")
    assert _greedy_ok(int(logits[0].argmax()), b0.float())


@pytest.mark.parametrize("kv_len", [1, 63, 64, 65, 1000, 3000, 4094])
def test_decode_stack_step_at_7b_dims(big, cuda, kv_len):
    """svln_decode_stack_step with the 148-CTA row maps of the real shape: cache lengths below one split, around the
    64-key tile boundary, the bench's lengths and the last usable arena rows.  Every step's logits vs fp32 HF (per
    element) and the kernel's token vs bf16 HF."""
    from transformers import DynamicCache
    from streamvln_b200.qwen2 import Chunk

    eng, lm32, lm16 = big
    assert eng.use_stack
    n_steps = 3
    n_ctx = min(kv_len, eng.arena.max_len - 1 - n_steps)     # 4094 -> 4092 prefilled, decode reaches row 4094
    emb = _emb(n_ctx, eng.cfg.hidden_size, 100 + kv_len, cuda)
    eng.keep_decode_logits = True
    eng._graphs.clear()
    try:
        eng.arena.reset_env(0)
        tok = eng.forward(emb.clone(), [Chunk(0, 0, n_ctx, n_ctx)]).argmax(-1)
        c32, c16 = DynamicCache(config=lm32.config), DynamicCache(config=lm16.config)
        with torch.no_grad():
            lm32(inputs_embeds=emb[None].float(), past_key_values=c32, use_cache=True)
            lm16(inputs_embeds=emb[None], past_key_values=c16, use_cache=True)
        runner = eng._decode_runner([0], True)
        for step in range(n_steps):
            fed = tok.clone()
            tok = runner.step(tok)
            assert int(tok) >= 0, "a grid barrier of the fused stack timed out"
            with torch.no_grad():
                want = lm32(input_ids=fed[None], past_key_values=c32, use_cache=True).logits[0, -1]
                want16 = lm16(input_ids=fed[None], past_key_values=c16, use_cache=True).logits[0, -1].float()
            got = runner.graph.state.stack_logits[0]
            _check_logits(got, want, want16, f"decode stack, kv {n_ctx + step + 1}")
            assert torch.equal(got, got.bfloat16().float()), "stack logits must be bf16-rounded (HF lm_head is a bf16 Linear)"
            assert int(tok) == int(got.argmax())
            assert _greedy_ok(int(tok), want16), (int(tok), int(want16.argmax()))
        runner.finish()
        assert eng.arena.lens[0] == c32.get_seq_length() == n_ctx + n_steps
    finally:
        eng.keep_decode_logits = False
        eng._graphs.clear()


@pytest.mark.parametrize("kv_lens", [(65, 3000), (1000, 1000), (1, 4090)])
def test_decode_stack_two_streams_at_7b_dims(big, cuda, kv_lens):
    """n_stream = 2 through svln_decode_stack_step (packed weight feed): two streams with different cache lengths share
    one weight pass; each stream's logits vs its own fp32 / bf16 HF run, step by step."""
    from transformers import DynamicCache
    from streamvln_b200.qwen2 import Chunk

    eng, lm32, lm16 = big
    assert eng.use_stack and eng.STACK_MAX_STREAMS >= 2
    n_steps = 3
    eng.keep_decode_logits = True
    eng._graphs.clear()
    try:
        toks, caches = [], []
        for b, n_ctx in enumerate(kv_lens):
            env = 2 + b
            emb = _emb(n_ctx, eng.cfg.hidden_size, 300 + n_ctx + b, cuda)
            eng.arena.reset_env(env)
            toks.append(eng.forward(emb.clone(), [Chunk(env, 0, n_ctx, n_ctx)]).argmax(-1).reshape(1))
            c32, c16 = DynamicCache(config=lm32.config), DynamicCache(config=lm16.config)
            with torch.no_grad():
                lm32(inputs_embeds=emb[None].float(), past_key_values=c32, use_cache=True)
                lm16(inputs_embeds=emb[None], past_key_values=c16, use_cache=True)
            caches.append((c32, c16))
        tok = torch.cat(toks)
        runner = eng._decode_runner([2, 3], True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=cuda)
        for step in range(n_steps):
            fed = tok.clone()
            flush.fill_(step)                         # K/V and weights come from HBM, as in a 28-layer step
            tok = runner.step(tok)
            assert runner.graph.state.stack is not None, "two streams must take the fused stack"
            assert int(tok.min()) >= 0, "a grid barrier of the fused stack timed out"
            for b, (c32, c16) in enumerate(caches):
                with torch.no_grad():
                    want = lm32(input_ids=fed[b].reshape(1, 1), past_key_values=c32, use_cache=True).logits[0, -1]
                    want16 = lm16(input_ids=fed[b].reshape(1, 1), past_key_values=c16, use_cache=True).logits[0, -1].float()
                got = runner.graph.state.stack_logits[b]
                _check_logits(got, want, want16, f"decode stack x2, stream {b}, kv {kv_lens[b] + step + 1}")
                assert torch.equal(got, got.bfloat16().float())
                assert int(tok[b]) == int(got.argmax())
                assert _greedy_ok(int(tok[b]), want16), (b, int(tok[b]), int(want16.argmax()))
        runner.finish()
        assert [eng.arena.lens[2], eng.arena.lens[3]] == [n + n_steps for n in kv_lens]
    finally:
        eng.keep_decode_logits = False
        eng._graphs.clear()


def test_decode_stack_is_deterministic_at_full_depth(svln_lib, cuda):
    """All 28 layers at the 7B shape, two streams (18 attention splits of up to 4 key tiles): the same state decoded four
    times gives bit-identical logits, and the stack stays close to the launch-per-op path.  Regression test: the in-kernel
    attention item consumed splits of more than two tiles before their last tiles had landed (`cp.async.wait_group 1` on
    a ring that refills one iteration late) - invisible with K/V resident in L2 (2 layers), every step with K/V in HBM."""
    from streamvln_b200.qwen2 import Chunk, Qwen2Config, Qwen2Engine, Qwen2Weights

    cfg = Qwen2Config()
    w = Qwen2Weights.random(cfg, cuda, seed=0, lm_head_std=0.05, norm_std=0.25)
    eng = Qwen2Engine(cfg, w, n_env=2, device=cuda)
    eng.keep_decode_logits = True
    ctx = [2970, 1733]
    first = []
    for b, n in enumerate(ctx):
        e = _emb(n, cfg.hidden_size, 5 + b, cuda)
        first.append(eng.forward(e, [Chunk(b, 0, n, n)]).argmax(-1).reshape(1))
    first = torch.cat(first)
    forced = None

    def decode(n_stack):
        eng.STACK_MAX_STREAMS = n_stack
        eng._graphs.clear()
        for b, n in enumerate(ctx):
            eng.arena.lens[b] = n
        runner = eng._decode_runner([0, 1], True)
        tok, logs, toks = first.clone(), [], []
        for i in range(6):
            tok = runner.step(tok)
            assert int(tok.min()) >= 0
            logs.append(runner.last_logits())
            toks.append(tok.clone())
            if forced is not None:
                tok = forced[i]                     # every run feeds the same tokens
        runner.finish()
        assert (runner.graph.state.stack is not None) == (n_stack >= 2)
        return torch.stack(logs), toks

    ref, forced = decode(0)                          # launch-per-op path (svln_gemv / cuBLAS + the small kernels)
    runs = [decode(2)[0] for _ in range(4)]
    assert all(torch.equal(runs[0], r) for r in runs[1:]), "the fused two-stream step is not deterministic"
    rel = float((runs[0] - ref).abs().max() / ref.abs().max())
    assert rel <= 8e-2, rel                          # 28 layers of bf16 rounding in two different accumulation orders
    del eng, w
    torch.cuda.empty_cache()


def test_bf16_tie_goes_to_the_lowest_index_on_every_decode_path(big, cuda):
    """Two lm_head rows whose fp32 logits differ by far less than a bf16 ulp: HF (bf16 lm_head, .float(), argmax) sees
    a tie and takes the LOWER index although the higher row's unrounded sum is larger.  The fused stack (1 stream),
    svln_gemv (the lm_head of 1 and 2 streams on the launch-per-op path) and cuBLAS (5 streams) must all agree."""
    from streamvln_b200.qwen2 import Chunk

    eng, lm32, lm16 = big
    D = eng.cfg.hidden_size
    emb = _emb(70, D, 7, cuda)
    n_env = 5
    lm_head = eng.w.lm_head
    saved = None
    try:
        for env in range(n_env):
            eng.arena.reset_env(env)
        eng.forward(torch.cat([emb] * n_env), [Chunk(e, e * 70, 70, 70) for e in range(n_env)])
        with torch.no_grad():
            o = lm32(inputs_embeds=emb[None].float(), use_cache=True)
            fed = torch.tensor([[123]], device=cuda)
            o = lm32(input_ids=fed, past_key_values=o.past_key_values, use_cache=True, output_hidden_states=True)
        want = o.logits[0, -1]
        x = o.hidden_states[-1][0, -1]                  # after the final norm: the lm_head input
        lo = int(want.argmax())
        hi = lo + 1000 if lo + 1000 < lm_head.shape[0] else lo - 1000
        lo, hi = min(lo, hi), max(lo, hi)
        # row `hi` := the winning row nudged by one bf16 ulp in ONE weight so that its fp32 logit is slightly LARGER
        src = int(want.argmax())
        other = hi if src == lo else lo
        saved = (other, lm_head[other].clone(), lm32.lm_head.weight[other].clone())
        row = lm_head[src].clone()
        lm_head[other] = row
        lm32.lm_head.weight.data[other] = row.float()
        top = float(want[src])
        # a weight whose product with x is ~2^-6 of the logit: one bf16 ulp of it moves the fp32 logit by ~2^-14 (relative),
        # far above fp32 accumulation noise (2^-24) and far below a bf16 ulp of the logit (2^-8)
        k = int(((x.abs() * row.float().abs()) - 2.0 ** -6 * abs(top)).abs().argmin())
        ulp_w = 2.0 ** (torch.floor(torch.log2(row[k].float().abs().clamp_min(1e-30))).item() - 7)
        sign = 1.0 if (float(x[k]) > 0) == (other == hi) else -1.0   # make the HIGHER index the (unrounded) larger one
        lm_head[other, k] = (row[k].float() + sign * ulp_w).bfloat16()
        lm32.lm_head.weight.data[other, k] = lm_head[other, k].float()
        with torch.no_grad():
            w2 = lm32.lm_head(x)
        assert float(w2[hi]) > float(w2[lo]) and abs(float(w2[hi] - w2[lo])) < 2.0 ** -12 * abs(top), "nudge is not a near-tie"
        assert float(w2[hi].bfloat16()) == float(w2[lo].bfloat16()) == float(w2.bfloat16().max())
        eng.repack_decode_weights()          # the fused step reads the packed copy of lm_head
        picks = {}
        for name, n, stack in (("stack", 1, True), ("stack2", 2, True), ("gemv", 1, False), ("gemv2", 2, False), ("cublas", 5, False)):
            eng.use_stack = stack
            eng._graphs.clear()
            for env in range(n_env):
                eng.arena.lens[env] = 70
            runner = eng._decode_runner(list(range(n)), False)
            t = runner.step(torch.full((n,), 123, device=cuda, dtype=torch.int64))
            runner.finish()
            picks[name] = t.tolist()
        assert all(v == [lo] * len(v) for v in picks.values()), (picks, lo, hi)
    finally:
        eng.use_stack = True
        eng._graphs.clear()
        if saved is not None:
            lm_head[saved[0]] = saved[1]
            lm32.lm_head.weight.data[saved[0]] = saved[2]
            eng.repack_decode_weights()


def _alone(eng, emb, env, n_new, eos=None):
    eng.arena.reset_env(env)
    return eng.generate([emb.clone()], [env], n_new, eos)[0].tolist()


def test_eos_stop_and_arena_rollback(big, cuda):
    """HF semantics of the greedy loop with EOS (GenerationMixin._sample, transformers 4.45.1): the sequence ends WITH
    the EOS token and the cache holds L_embeds + n_gen - 1 rows.  EOS as the very first token, mid-sequence, and never."""
    eng, lm32, lm16 = big
    emb = _emb(150, eng.cfg.hidden_size, 21, cuda)
    free = _alone(eng, emb, 0, 8)
    assert len(free) == 8 and eng.arena.lens[0] == 150 + 8 - 1
    # EOS = first generated token: one token out, nothing fed back
    got = _alone(eng, emb, 0, 8, eos=free[0])
    assert got == free[:1] and eng.arena.lens[0] == 150
    # mid-sequence (first occurrence of that id decides)
    ks = [i for i in range(1, 8) if free[i] not in free[:i]]
    assert ks, f"degenerate random sequence {free}: pick another seed"
    k = ks[-1]
    got = _alone(eng, emb, 0, 8, eos=free[k])
    assert got == free[:k + 1] and eng.arena.lens[0] == 150 + (k + 1) - 1
    # an id the model never emits: runs to max_new_tokens
    never = next(t for t in range(1000) if t not in free)
    got = _alone(eng, emb, 0, 8, eos=never)
    assert got == free and eng.arena.lens[0] == 150 + 8 - 1
    # the decode continues correctly from a rolled-back cache: feed one more chunk and compare with fp32 HF
    from transformers import DynamicCache
    from streamvln_b200.qwen2 import Chunk

    got = _alone(eng, emb, 0, 8, eos=free[k])
    nxt = _emb(30, eng.cfg.hidden_size, 22, cuda)
    fed_ids = torch.tensor([got], device=cuda)
    tail = torch.cat([eng.w.embed[fed_ids[0, -1:]], nxt])          # the last generated id is fed with the next turn
    logits = eng.forward(tail.clone(), [Chunk(0, 0, 31, eng.arena.lens[0] + 31)])
    c, c16 = DynamicCache(config=lm32.config), DynamicCache(config=lm16.config)
    with torch.no_grad():
        lm32(inputs_embeds=emb[None].float(), past_key_values=c, use_cache=True)
        lm16(inputs_embeds=emb[None], past_key_values=c16, use_cache=True)
        if fed_ids.shape[1] > 1:
            lm32(input_ids=fed_ids[:, :-1], past_key_values=c, use_cache=True)
            lm16(input_ids=fed_ids[:, :-1], past_key_values=c16, use_cache=True)
        want = lm32(inputs_embeds=tail[None].float(), past_key_values=c, use_cache=True).logits[0, -1]
        want16 = lm16(inputs_embeds=tail[None], past_key_values=c16, use_cache=True).logits[0, -1]
    assert c.get_seq_length() == eng.arena.lens[0]
    _check_logits(logits[0], want, want16, "chunk after an EOS roll-back")


def test_packed_batch_with_mixed_budgets_and_eos(big, cuda):
    """Three streams in one packed generate: different prompt lengths, one stops on EOS early, and (through
    StreamVLNForCausalLM.generate_batch's budget rule) different max_new_tokens.  Each stream's ids and cache length
    equal the stream run alone."""
    eng, lm32, lm16 = big
    D = eng.cfg.hidden_size
    embs = [_emb(n, D, 30 + i, cuda) for i, n in enumerate((90, 217, 40))]
    alone = [_alone(eng, e, 3, 6) for e in embs]
    eos = alone[1][2]
    if eos in alone[1][:2]:
        pytest.skip("degenerate random sequence")
    want = []
    for a in alone:
        want.append(a[:a.index(eos) + 1] if eos in a else a)
    for env in range(3):
        eng.arena.reset_env(env)
    got = eng.generate([e.clone() for e in embs], [0, 1, 2], 6, eos)
    for b in range(3):
        g = got[b].tolist()
        # packed (svln_gemv, M = 3) and alone (fused stack) differ only in accumulation order: identical ids expected
        assert g == want[b], (b, g, want[b])
        assert eng.arena.lens[b] == embs[b].shape[0] + len(g) - 1


def test_generate_batch_budget_rule_and_clamp(svln_lib, cuda):
    """model.generate_batch: a stream with a smaller max_new_tokens than the batch maximum gets its surplus ids un-fed;
    a budget beyond the arena is clamped (HF would run into max_position_embeddings)."""
    from streamvln_b200.model import StreamVLNConfig, StreamVLNForCausalLM

    cfg = StreamVLNConfig.tiny(vocab_size=512, layers=2, max_pos=512)
    m = StreamVLNForCausalLM(cfg, n_env=3, device=cuda, seed=5, lm_head_std=0.3)
    ids = [[7, 8, 9, 10], [11, 12, 13, 14, 15, 16], [17, 18]]
    reqs = [dict(inputs=torch.tensor([i]), images=None, env_id=e, max_new_tokens=mn, past_key_values=None)
            for e, (i, mn) in enumerate(zip(ids, (2, 7, 4)))]
    outs = m.generate_batch(reqs)
    for e, (o, mn) in enumerate(zip(outs, (2, 7, 4))):
        assert o.sequences.shape == (1, mn)
        assert m.engine.arena.lens[e] == o.past_key_values.get_seq_length() == len(ids[e]) + mn - 1
    solo = []
    for e, (i, mn) in enumerate(zip(ids, (2, 7, 4))):
        m.reset_for_env(e)
        solo.append(m.generate(inputs=torch.tensor([i]), images=None, env_id=e, max_new_tokens=mn, past_key_values=None))
    for o, s in zip(outs, solo):
        assert o.sequences.tolist() == s.sequences.tolist()
    m.reset_for_env(0)
    with pytest.warns(UserWarning, match="clamped"):
        o = m.generate(inputs=torch.tensor([ids[0]]), images=None, env_id=0, max_new_tokens=10000, past_key_values=None)
    assert o.sequences.shape[1] == 512 - len(ids[0]) and m.engine.arena.lens[0] == 511


def test_zz_print_logit_error_stats():
    """Not a check: prints what the tolerance checks above measured (pytest -s), for DESIGN.md."""
    print("\n  case | product vs fp32: row, per-element>=floor | HF bf16 vs fp32: row, per-element | product vs HF bf16: row")
    for what, p_all, p_big, h_all, h_big, ph_all in STATS:
        print(f"  {what:32s} {p_all:.2e} {p_big:.2e} | {h_all:.2e} {h_big:.2e} | {ph_all:.2e}")


def _feeds_agree(cfg, ctxs, cuda):
    import os
    from streamvln_b200.qwen2 import Chunk, Qwen2Engine, Qwen2Weights

    w = Qwen2Weights.random(cfg, cuda, seed=0)
    old = os.environ.get("SVLN_STACK_FEED")
    try:
        for n_ctx in ctxs:
            outs = []
            for feed in ("1", "0"):
                os.environ["SVLN_STACK_FEED"] = feed
                eng = Qwen2Engine(cfg, w, n_env=1, device=cuda)
                eng.keep_decode_logits = True
                emb = (torch.randn(n_ctx, cfg.hidden_size, generator=torch.Generator().manual_seed(1)) * 0.05).bfloat16().to(cuda)
                tok = eng.forward(emb, [Chunk(0, 0, n_ctx, n_ctx)]).argmax(-1)
                runner = eng._decode_runner([0], False)
                lg = []
                for _ in range(3):
                    tok = runner.step(tok)
                    lg.append(runner.last_logits().float().clone())
                runner.finish()
                outs.append(torch.stack(lg))
                del eng
            a, b = outs
            rel = float((a - b).norm() / b.norm())
            assert rel <= 1e-2, (cfg.hidden_size, n_ctx, rel)
            assert bool((a.argmax(-1) == b.argmax(-1)).all()), (cfg.hidden_size, n_ctx)
    finally:
        if old is None:
            os.environ.pop("SVLN_STACK_FEED", None)
        else:
            os.environ["SVLN_STACK_FEED"] = old
        del w
        torch.cuda.empty_cache()


def test_decode_stack_feeds_agree(svln_lib, cuda):
    """The two weight feeds of svln_decode_stack_step - the packed stream (bulk copies by a producer thread; default) and
    the row-major LDGSTS feed (descriptor without `packed`) - are the same arithmetic up to the summation order of the wide
    phases: logits of three decode steps agree to 1e-2 (each is within the HF tolerance on its own) and pick the same ids.
    7B widths, 2 layers, contexts 3 / 64 / 700."""
    from streamvln_b200.qwen2 import Qwen2Config

    _feeds_agree(Qwen2Config(num_hidden_layers=2), (3, 64, 700), cuda)


def test_decode_stack_feeds_agree_other_widths(svln_lib, cuda):
    """The packed layout is generic in the model dims: widths that give a CTA more than 32 QKV rows (two row groups in
    the narrow phases), 8 kv heads (4 query heads per group), an inter size and a vocabulary that do not divide by the grid
    (hidden 4096, 32:8 heads, inter 14336, vocab 128256; 2 layers)."""
    from streamvln_b200.qwen2 import Qwen2Config

    _feeds_agree(Qwen2Config(num_hidden_layers=2, hidden_size=4096, intermediate_size=14336, num_attention_heads=32,
                             num_key_value_heads=8, vocab_size=128256), (65, 1500), cuda)

"""Error budget at the benchmarked shape (run on the GPU box): product vs HF fp32, HF bf16 vs HF fp32, product vs HF bf16."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch
from transformers import DynamicCache
from streamvln_b200.qwen2 import Chunk, Qwen2Config, Qwen2Engine, Qwen2Weights
from test_gpu_7b_shape import _hf, _emb

cuda = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = Qwen2Config(num_hidden_layers=layers)
for norm_std, lm_std in ((0.25, 0.05), (0.0, 0.02)):
    w = Qwen2Weights.random(cfg, cuda, seed=11, lm_head_std=lm_std, norm_std=norm_std)
    eng = Qwen2Engine(cfg, w, n_env=2, device=cuda)
    eng.keep_decode_logits = True
    lm32, lm16 = _hf(w, torch.float32, cuda), _hf(w, torch.bfloat16, cuda)

    def rel(a, b):
        a, b = a.double(), b.double()
        return ((a - b).abs().max() / b.abs().max()).item()

    def rel_big(a, b, floor=0.5):
        a, b = a.double(), b.double()
        m = b.abs() >= floor * b.abs().max()
        return ((a - b).abs()[m] / b.abs()[m]).max().item(), int(m.sum())

    print(f"--- layers {layers} norm_std {norm_std} lm_head_std {lm_std}")
    for n_ctx in (300, 1900):
        emb = _emb(n_ctx, cfg.hidden_size, 1, cuda)
        eng.arena.reset_env(0)
        got = eng.forward(emb.clone(), [Chunk(0, 0, n_ctx, n_ctx)])[0]
        c32, c16 = DynamicCache(config=lm32.config), DynamicCache(config=lm16.config)
        with torch.no_grad():
            w32 = lm32(inputs_embeds=emb[None].float(), past_key_values=c32, use_cache=True).logits[0, -1]
            w16 = lm16(inputs_embeds=emb[None], past_key_values=c16, use_cache=True).logits[0, -1].float()
        print(f"prefill {n_ctx}: prod-vs-fp32 {rel(got, w32):.2e} (big {rel_big(got, w32)[0]:.2e})  hf16-vs-fp32 {rel(w16, w32):.2e} "
              f"(big {rel_big(w16, w32)[0]:.2e})  prod-vs-hf16 {rel(got, w16):.2e} (big {rel_big(got, w16)[0]:.2e}, n {rel_big(got, w16)[1]}) "
              f"argmax {int(got.argmax())} {int(w32.argmax())} {int(w16.argmax())}")
        tok = got.argmax(-1)[None]
        runner = eng._decode_runner([0], False)
        for step in range(3):
            fed = tok.clone()
            tok = runner.step(tok)
            with torch.no_grad():
                w32 = lm32(input_ids=fed[None], past_key_values=c32, use_cache=True).logits[0, -1]
                w16 = lm16(input_ids=fed[None], past_key_values=c16, use_cache=True).logits[0, -1].float()
            got = runner.state.stack_logits[0]
            print(f"  decode {step}: prod-vs-fp32 {rel(got, w32):.2e} (big {rel_big(got, w32)[0]:.2e})  hf16-vs-fp32 {rel(w16, w32):.2e} "
                  f"(big {rel_big(w16, w32)[0]:.2e})  prod-vs-hf16 {rel(got, w16):.2e} (big {rel_big(got, w16)[0]:.2e}) "
                  f"argmax {int(got.argmax())} {int(w32.argmax())} {int(w16.argmax())}")
        runner.finish()
    del eng, lm32, lm16, w
    torch.cuda.empty_cache()

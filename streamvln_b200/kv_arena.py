"""Fast-streaming KV arena — host side over the C-ABI (include/svln.h §2, §3).

Replaces HF `DynamicCache` (torch.cat per layer per forward; reached from
streamvln/model/stream_video_vln.py:340) with one pre-allocated arena
[env][layer][K|V][kv_head][max_len][128] in HBM, and the reference's drop-everything window policy
(stream_video_vln.py:473-479, streamvln_agent.py:192-198) with append / evict+compact / re-rotate
kernels.  All compute is in libsvln.so; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib
from ._lib import KvArenaStruct, check, cur_stream, ptr


def hf_rope_tables(max_pos: int, head_dim: int, theta: float):
    """cos/sin exactly as HF Qwen2RotaryEmbedding builds them (fp32 outer product, cat(freqs, freqs),
    cos()/sin(), cast to bf16) — computed on CPU so that the CPU oracle sees identical tables."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(max_pos, dtype=torch.float32)
    freqs = (inv_freq[:, None] @ pos[None, :]).transpose(0, 1)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)


def delta_rope_tables(n_delta: int, head_dim: int, theta: float):
    """fp32 cos/sin of m * inv_freq[i] for position shifts m in [0, n_delta) (computed in fp64)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float64) / head_dim))
    ang = torch.arange(n_delta, dtype=torch.float64)[:, None] * inv_freq[None, :]
    return ang.cos().float(), ang.sin().float()


class _Lens(list):
    """Host-side cache lengths.  A device-side eviction (slowfast.py) owes the host its data-dependent new lengths
    until the turn's D2H read has happened; `pending` (a callable) settles that debt on the first access."""

    pending = None

    def _settle(self):
        fn, self.pending = self.pending, None
        if fn is not None:
            fn()

    def __getitem__(self, i):
        self._settle()
        return list.__getitem__(self, i)

    def __setitem__(self, i, v):
        self._settle()
        list.__setitem__(self, i, v)

    def __iter__(self):
        self._settle()
        return list.__iter__(self)


class KvArena:
    def __init__(self, n_env: int, n_layer: int, n_kv_head: int, head_dim: int, max_len: int,
                 rope_theta: float = 1e6, device="cuda"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("KvArena needs a CUDA device (no CPU fallback)")
        if head_dim != 128:
            raise ValueError("libsvln kernels are specialised for head_dim 128 (Qwen2-7B)")
        self.n_env, self.n_layer, self.n_kv_head, self.head_dim, self.max_len = n_env, n_layer, n_kv_head, head_dim, max_len
        # zero-initialised: masked-out rows must hold finite values (0 * NaN would poison P.V)
        self.buf = torch.zeros(n_env, n_layer, 2, n_kv_head, max_len, head_dim, dtype=torch.bfloat16,
                               device=self.device)
        self.struct = KvArenaStruct(self.buf.data_ptr(), n_env, n_layer, n_kv_head, head_dim, max_len, 0)
        self.ref = C.byref(self.struct)
        cos, sin = hf_rope_tables(max_len, head_dim, rope_theta)
        self.cos, self.sin = cos.to(self.device).contiguous(), sin.to(self.device).contiguous()
        cd, sd = delta_rope_tables(max_len, head_dim, rope_theta)
        self.cos_delta, self.sin_delta = cd.to(self.device).contiguous(), sd.to(self.device).contiguous()
        self.lens = _Lens([0] * n_env)  # host-side cache lengths (rows in use per stream)
        # attention scratch (split-KV partials), allocated ONCE: captured CUDA graphs hold its address
        self._attn_ws_rows = max(256, 4 * n_env)
        self._attn_ws = torch.empty(self.lib.svln_attn_workspace_bytes(self._attn_ws_rows, 32, head_dim, max_len),
                                    dtype=torch.uint8, device=self.device)

    # ---------------------------------------------------------------- append
    def append_rope(self, layer: int, qkv: torch.Tensor, n_q_head: int, tok_env: torch.Tensor,
                    tok_pos: torch.Tensor, q_out: torch.Tensor) -> None:
        n_tok = qkv.shape[0]
        check(self.lib.svln_kv_append_rope(self.ref, layer, ptr(qkv), n_q_head, ptr(tok_env), ptr(tok_pos), n_tok,
                                           ptr(self.cos), ptr(self.sin), self.max_len, ptr(q_out), cur_stream()))

    # ---------------------------------------------------------------- evict + compact
    def compact(self, job_env: torch.Tensor, first_row: torch.Tensor, new_len: torch.Tensor,
                src_row: torch.Tensor, rerotate: bool = True) -> None:
        """job_env/first_row/new_len [J] int32, src_row [J, max_len] int32 (device)."""
        n_jobs = job_env.numel()
        check(self.lib.svln_kv_compact(self.ref, ptr(job_env), ptr(first_row), ptr(new_len), ptr(src_row), n_jobs,
                                       1 if rerotate else 0, ptr(self.cos_delta), ptr(self.sin_delta),
                                       self.max_len, cur_stream()))

    def rerotate(self, env: int, row0: int, delta_pos: torch.Tensor) -> None:
        check(self.lib.svln_kv_rerotate(self.ref, env, row0, delta_pos.numel(), ptr(delta_pos), ptr(self.cos_delta),
                                        ptr(self.sin_delta), self.max_len, cur_stream()))

    def reset_env(self, env: int) -> None:
        """Window reset in reference semantics = forget the rows (stream_video_vln.py:477-479)."""
        self.lens[env] = 0

    # ---------------------------------------------------------------- attention
    def _workspace(self, n_rows: int, n_q_head: int) -> torch.Tensor:
        return self._attn_ws  # the C side checks the size it needs against workspace_bytes and fails loudly

    def reserve_workspace(self, n_rows: int, n_q_head: int) -> None:
        self._workspace(n_rows, n_q_head)

    def attn_decode(self, layer: int, q: torch.Tensor, n_q_head: int, seq_env: torch.Tensor, q_start: torch.Tensor,
                    n_q: torch.Tensor, kv_len: torch.Tensor, max_n_q: int, max_kv_len: int, scale: float,
                    out: torch.Tensor) -> None:
        ws = self._workspace(q.shape[0], n_q_head)
        check(self.lib.svln_attn_decode(self.ref, layer, ptr(q), n_q_head, ptr(seq_env), ptr(q_start), ptr(n_q),
                                        ptr(kv_len), seq_env.numel(), max_n_q, max_kv_len, scale, ptr(ws), ws.numel(),
                                        ptr(out), cur_stream()))

    def attn_prefill(self, layer: int, q: torch.Tensor, n_q_head: int, seq_env: Sequence[int], q_start: Sequence[int],
                     n_q: Sequence[int], kv_len: Sequence[int], scale: float, out: torch.Tensor) -> None:
        n = len(seq_env)
        arr = (C.c_int32 * n)
        ws = self._workspace(0, n_q_head)
        check(self.lib.svln_attn_prefill(self.ref, layer, ptr(q), n_q_head, arr(*seq_env), arr(*q_start), arr(*n_q),
                                         arr(*kv_len), n, scale, ptr(ws), ws.numel(), ptr(out), cur_stream()))

    def attn_prefill_dev(self, layer: int, q: torch.Tensor, n_q_head: int, desc: torch.Tensor, n_q_host: Sequence[int],
                         scale: float, out: torch.Tensor) -> None:
        """desc: device int32 [4, n_seq] = (env, q_start, n_q, kv_len) rows; graph-replayable."""
        n = len(n_q_host)
        ws = self._workspace(0, n_q_head)
        check(self.lib.svln_attn_prefill_dev(self.ref, layer, ptr(q), n_q_head, ptr(desc[0]), ptr(desc[1]), ptr(desc[2]),
                                             ptr(desc[3]), (C.c_int32 * n)(*n_q_host), n, scale, ptr(ws), ws.numel(), ptr(out),
                                             cur_stream()))

    # ---------------------------------------------------------------- views (tests, debugging)
    def plane(self, env: int, layer: int, kv: int) -> torch.Tensor:
        return self.buf[env, layer, kv]  # [n_kv_head, max_len, 128]

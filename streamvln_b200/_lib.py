"""ctypes binding of libsvln.so (the C-ABI in include/svln.h).

There is no CPU fallback: if the shared library is missing or fails to load, importing a product
module that needs it raises.  The library is built in-tree by `streamvln_b200.build`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsvln.so"

SVLN_OK = 0
INVALID_KEEP = 0
INVALID_DROP = 1
KEY_INVALID = -1
SF_JOB_WORDS = 10   # SVLN_SF_JOB_WORDS

_ERRNAMES = {-22: "EINVAL", -12: "ENOMEM", -34: "ERANGE", -5: "EIO(CUDA)"}


class SvlnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsvln error {code} ({_ERRNAMES.get(code, '?')}): {msg}")
        self.code = code


class KvArenaStruct(C.Structure):
    """Mirror of `struct svln_kv_arena`."""

    _fields_ = [
        ("base", C.c_void_p),
        ("n_env", C.c_int32),
        ("n_layer", C.c_int32),
        ("n_kv_head", C.c_int32),
        ("head_dim", C.c_int32),
        ("max_len", C.c_int32),
        ("reserved", C.c_int32),
    ]


class DecodeStackStruct(C.Structure):
    """Mirror of `struct svln_decode_stack`."""

    _fields_ = [
        ("n_layer", C.c_int32), ("hidden", C.c_int32), ("inter", C.c_int32), ("n_q_head", C.c_int32),
        ("n_kv_head", C.c_int32), ("head_dim", C.c_int32), ("vocab", C.c_int32), ("n_stream", C.c_int32),
        ("rms_eps", C.c_float), ("attn_scale", C.c_float),
        ("layer_ptrs", C.c_void_p), ("embed", C.c_void_p), ("final_norm", C.c_void_p), ("lm_head", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("tok_in", C.c_void_p), ("tok_out", C.c_void_p), ("logits", C.c_void_p),
        ("seq_env", C.c_void_p), ("kv_len", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("packed", C.c_void_p), ("packed_bytes", C.c_size_t),
    ]


_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_arena_p = C.POINTER(KvArenaStruct)

# name -> (restype, argtypes); must list every symbol include/svln.h declares
PROTOTYPES = {
    "svln_version": (_i, []),
    "svln_last_error": (C.c_char_p, []),
    "svln_sm_count": (_i, []),
    "svln_voxel_keys": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _f, _f, _vp, _vp]),
    "svln_prune_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "svln_keep_latest": (_i, [_vp, _vp, _i64, _vp, _i64, _i64, _i, _vp, _sz, _vp, _vp]),
    "svln_frame_threshold": (_i, [_vp, _i64, _i, _i, _vp]),
    "svln_compact_mask": (_i, [_vp, _i64, _vp, _sz, _vp, _vp, _vp, _i64, _vp, _vp]),
    "svln_gather_rows": (_i, [_vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "svln_voxel_prune": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _i, _i, _i,
                              _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "svln_voxel_prune_host_staging_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "svln_voxel_prune_host": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _i, _i, _i,
                                   _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "svln_preprocess_workspace_bytes": (_sz, [_i, _i, _i]),
    "svln_preprocess_rgb": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "svln_preprocess_depth": (_i, [_vp, _i, _i, _i, _i, _i, _f, _f, _vp, _vp]),
    "svln_splice_rows": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "svln_kv_arena_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "svln_kv_append_rope": (_i, [_arena_p, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "svln_kv_compact": (_i, [_arena_p, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "svln_kv_rerotate": (_i, [_arena_p, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "svln_sf_gather_candidates": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "svln_sf_evict_plan": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "svln_attn_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "svln_attn_decode": (_i, [_arena_p, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _sz,
                              _vp, _vp]),
    "svln_attn_prefill": (_i, [_arena_p, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _f, _vp, _sz, _vp,
                               _vp]),
    "svln_attn_prefill_dev": (_i, [_arena_p, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _sz, _vp, _vp]),
    "svln_rmsnorm": (_i, [_vp, _vp, _i64, _i, _f, _vp, _vp]),
    "svln_add_rmsnorm": (_i, [_vp, _vp, _vp, _i64, _i, _f, _vp, _vp]),
    "svln_swiglu": (_i, [_vp, _i64, _i, _vp, _vp]),
    "svln_gemv": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "svln_argmax": (_i, [_vp, _i, _i64, _i64, _vp, _vp]),
    "svln_pool_bilinear": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "svln_decode_stack_workspace_bytes": (C.c_size_t, [C.POINTER(DecodeStackStruct)]),
    "svln_decode_stack_packed_bytes": (C.c_size_t, [C.POINTER(DecodeStackStruct)]),
    "svln_decode_stack_pack": (_i, [C.POINTER(DecodeStackStruct), _vp, _sz, _vp]),
    "svln_decode_stack_status": (_i, [C.POINTER(DecodeStackStruct), _vp, _vp]),
    "svln_decode_stack_trace": (_i, [_vp, _i]),
    "svln_decode_stack_step": (_i, [C.POINTER(KvArenaStruct), C.POINTER(DecodeStackStruct), _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libsvln.so and bind every prototype.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m streamvln_b200.build` "
            "(streamvln_b200 has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != SVLN_OK:
        raise SvlnError(rc, load().svln_last_error().decode(errors="replace"))


def ptr(t) -> int:
    """Device/host address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream

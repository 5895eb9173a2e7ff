// Fast-streaming KV arena: fused RoPE + append, window eviction with in-place compaction and
// memory-token splice, RoPE position remap.  Bandwidth-bound kernels (128-bit accesses).
// Replaces HF DynamicCache.update (torch.cat per layer per forward; reached from
// stream_video_vln.py:340) and the drop-everything window policy (stream_video_vln.py:473-479).
#include "svln_common.cuh"

namespace svln {

__device__ __forceinline__ int64_t plane_elems(const svln_kv_arena& a) {
    return (int64_t)a.max_len * a.head_dim;
}
// element offset of (env, layer, kv, head) plane
__device__ __forceinline__ int64_t plane_offset(const svln_kv_arena& a, int env, int layer, int kv, int head) {
    return ((((int64_t)env * a.n_layer + layer) * 2 + kv) * a.n_kv_head + head) * plane_elems(a);
}

// ------------------------------------------------------------------ fused RoPE + append
// 8 threads per 128-wide head vector; thread j owns dims [8j,8j+8) and [64+8j,64+8j+8).
// HF rounding points (apply_rotary_pos_emb in bf16): out = bf16(bf16(x*cos) + bf16(rot(x)*sin)).
// Packed bf16x2 arithmetic: a product of two bf16 values is exact in fp32, so HMUL2.BF16 (one rounding) equals HF's
// bf16(x * cos); the sum of two bf16 values rounded once equals the fp32 sum rounded to bf16 (no double-rounding case
// exists for 8-bit significands).  The _rn intrinsics forbid contraction into an FMA, which would drop HF's
// intermediate rounding.  6 packed instructions per 4 outputs instead of ~40 scalar ones: the kernel was issue-bound
// (ncu round 1: DRAM 45 % while SM throughput 73 %).
__device__ __forceinline__ void rope8(const int4& lo, const int4& hi, const int4& cos_lo, const int4& cos_hi,
                                      const int4& sin_lo, const int4& sin_hi, int4& out_lo, int4& out_hi) {
    const __nv_bfloat162* l = reinterpret_cast<const __nv_bfloat162*>(&lo);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&hi);
    const __nv_bfloat162* cl = reinterpret_cast<const __nv_bfloat162*>(&cos_lo);
    const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&cos_hi);
    const __nv_bfloat162* sl = reinterpret_cast<const __nv_bfloat162*>(&sin_lo);
    const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&sin_hi);
    __nv_bfloat162* ol = reinterpret_cast<__nv_bfloat162*>(&out_lo);
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&out_hi);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        // lo: x_lo*cos_lo + (-x_hi)*sin_lo ; hi: x_hi*cos_hi + x_lo*sin_hi   (rotate_half(x) = (-x_hi, x_lo))
        ol[w] = __hsub2_rn(__hmul2_rn(l[w], cl[w]), __hmul2_rn(h[w], sl[w]));
        oh[w] = __hadd2_rn(__hmul2_rn(h[w], ch[w]), __hmul2_rn(l[w], sh[w]));
    }
}

__global__ void __launch_bounds__(256, 8)
kv_append_rope_kernel(svln_kv_arena arena, int layer, const int4* __restrict__ qkv, int n_q_head,
                      const int32_t* __restrict__ tok_env, const int32_t* __restrict__ tok_pos, int n_tok,
                      const int4* __restrict__ cos_t, const int4* __restrict__ sin_t, int n_pos,
                      int4* __restrict__ q_out) {
    const int heads = n_q_head + 2 * arena.n_kv_head;
    const int64_t vec = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int j = threadIdx.x & 7;
    if (vec >= (int64_t)n_tok * heads) return;
    const int tok = (int)(vec / heads), h = (int)(vec - (int64_t)tok * heads);
    const int env = tok_env[tok], pos = tok_pos[tok];
    if (env < 0 || env >= arena.n_env || pos < 0 || pos >= arena.max_len || pos >= n_pos) return;
    const int4* src = qkv + ((int64_t)tok * heads + h) * 16;
    const int4 lo = ld_stream(src + j), hi = ld_stream(src + 8 + j);   // the GEMM output is read exactly once
    int4* dst;
    bool rotate = true;
    if (h < n_q_head) {
        dst = q_out + ((int64_t)tok * n_q_head + h) * 16;
    } else if (h < n_q_head + arena.n_kv_head) {
        dst = reinterpret_cast<int4*>(reinterpret_cast<__nv_bfloat16*>(arena.base) +
                                      plane_offset(arena, env, layer, 0, h - n_q_head)) + (int64_t)pos * 16;
    } else {
        dst = reinterpret_cast<int4*>(reinterpret_cast<__nv_bfloat16*>(arena.base) +
                                      plane_offset(arena, env, layer, 1, h - n_q_head - arena.n_kv_head)) +
              (int64_t)pos * 16;
        rotate = false;
    }
    if (rotate) {
        const int4* c = cos_t + (int64_t)pos * 16;
        const int4* s = sin_t + (int64_t)pos * 16;
        int4 olo, ohi;
        rope8(lo, hi, __ldg(c + j), __ldg(c + 8 + j), __ldg(s + j), __ldg(s + 8 + j), olo, ohi);
        dst[j] = olo;
        dst[8 + j] = ohi;
    } else {
        dst[j] = lo;
        dst[8 + j] = hi;
    }
}

// ------------------------------------------------------- eviction + in-place compaction
// One CTA walks one (job, layer, K|V, head) plane front to back in chunks of CR rows.  Because
// src_row is strictly increasing with src >= dst, a chunk's destination rows are never the
// source of a later chunk, and the barrier orders this chunk's loads before its stores.
// K rows are re-rotated by (dst - src) positions while they are in registers.
constexpr int CR = 128;           // rows per chunk: 128 rows x 256 B = 32 KB in flight per CTA, 3 CTAs per SM
constexpr int CQ = CR * 16 / 256; // int4 per thread per chunk (16)

// rotate an 8-element (lo, hi) pair of fp32-expanded bf16 chunks by angle table row m:
// lo' = lo*cos - hi*sin*sgn ; hi' = hi*cos + lo*sin*sgn, sgn=+1 rotates forward (pos increases)
__device__ __forceinline__ void rerotate8(int4& lo, int4& hi, const float* __restrict__ cosr,
                                          const float* __restrict__ sinr, int chunk, float sgn) {
    uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
    uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t rl = 0, rh = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int sft = e * 16;
            const int d = chunk * 8 + w * 2 + e;  // pair index 0..63
            const float c = __ldg(cosr + d);
            const float s = __fmul_rn(__ldg(sinr + d), sgn);
            const float xl = bf16_bits_to_f32((l[w] >> sft) & 0xFFFFu);
            const float xh = bf16_bits_to_f32((h[w] >> sft) & 0xFFFFu);
            const float nl = __fsub_rn(__fmul_rn(xl, c), __fmul_rn(xh, s));
            const float nh = __fadd_rn(__fmul_rn(xh, c), __fmul_rn(xl, s));
            rl |= f32_to_bf16_bits(nl) << sft;
            rh |= f32_to_bf16_bits(nh) << sft;
        }
        l[w] = rl;
        h[w] = rh;
    }
}

__global__ void __launch_bounds__(256, 3)
kv_compact_kernel(svln_kv_arena arena, const int32_t* __restrict__ job_env,
                  const int32_t* __restrict__ first_row, const int32_t* __restrict__ new_len,
                  const int32_t* __restrict__ src_row, int rerotate, const float* __restrict__ cos_delta,
                  const float* __restrict__ sin_delta, int n_delta) {
    const int job = blockIdx.y;
    const int planes_per_env = arena.n_layer * 2 * arena.n_kv_head;
    const int p = blockIdx.x;  // plane within env
    if (p >= planes_per_env) return;
    const int env = job_env[job];
    const int first = first_row[job], nlen = new_len[job];
    if (env < 0 || env >= arena.n_env || first >= nlen) return;
    const bool is_k = ((p / arena.n_kv_head) & 1) == 0;
    int4* plane = reinterpret_cast<int4*>(reinterpret_cast<__nv_bfloat16*>(arena.base) +
                                          ((int64_t)env * planes_per_env + p) * plane_elems(arena));
    const int32_t* src = src_row + (int64_t)job * arena.max_len;
    const int col = threadIdx.x & 15;          // 16-byte chunk within the row
    const int row_in_pass = threadIdx.x >> 4;  // 16 rows per pass of 256 threads
    const bool do_rot = rerotate && is_k;
    for (int d0 = first; d0 < nlen; d0 += CR) {
        int4 v[CQ];
        int32_t s_row[CQ];
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
            const int d = d0 + q * 16 + row_in_pass;
            s_row[q] = d < nlen ? __ldg(src + d) : -1;
            if (s_row[q] >= 0) v[q] = ld_cg(plane + (int64_t)s_row[q] * 16 + col);
        }
        if (do_rot) {
#pragma unroll
            for (int q = 0; q < CQ; ++q) {
                const int d = d0 + q * 16 + row_in_pass;
                // partner chunk (col ^ 8) lives in lane ^ 8 of the same warp
                int4 other;
                other.x = __shfl_xor_sync(0xffffffffu, v[q].x, 8);
                other.y = __shfl_xor_sync(0xffffffffu, v[q].y, 8);
                other.z = __shfl_xor_sync(0xffffffffu, v[q].z, 8);
                other.w = __shfl_xor_sync(0xffffffffu, v[q].w, 8);
                const int m = s_row[q] - d;  // >= 0: row moves to a smaller position
                if (s_row[q] >= 0 && m > 0 && m < n_delta) {
                    const float* cr = cos_delta + (int64_t)m * 64;
                    const float* sr = sin_delta + (int64_t)m * 64;
                    if (col < 8) {
                        int4 lo = v[q], hi = other;
                        rerotate8(lo, hi, cr, sr, col, -1.0f);
                        v[q] = lo;
                    } else {
                        int4 lo = other, hi = v[q];
                        rerotate8(lo, hi, cr, sr, col - 8, -1.0f);
                        v[q] = hi;
                    }
                }
            }
        }
        __syncthreads();  // every load of this chunk has been performed before any store below
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
            const int d = d0 + q * 16 + row_in_pass;
            if (s_row[q] >= 0 && (s_row[q] != d)) plane[(int64_t)d * 16 + col] = v[q];
        }
        // no second barrier: the next chunk's sources are all >= d0 + CR, beyond these stores
    }
}

__global__ void __launch_bounds__(256)
kv_rerotate_kernel(svln_kv_arena arena, int env, int row0, int n_rows, const int32_t* __restrict__ delta_pos,
                   const float* __restrict__ cos_delta, const float* __restrict__ sin_delta, int n_delta) {
    // blockIdx.y: K plane (layer, head); blockIdx.x: chunks of 32 rows (8 threads per row)
    const int layer = blockIdx.y / arena.n_kv_head, head = blockIdx.y % arena.n_kv_head;
    const int r = blockIdx.x * 32 + (threadIdx.x >> 3);
    const int j = threadIdx.x & 7;
    if (r >= n_rows) return;
    const int delta = delta_pos[r];
    const int m = delta < 0 ? -delta : delta;
    if (m == 0 || m >= n_delta) return;
    int4* row = reinterpret_cast<int4*>(reinterpret_cast<__nv_bfloat16*>(arena.base) +
                                        plane_offset(arena, env, layer, 0, head)) + (int64_t)(row0 + r) * 16;
    int4 lo = row[j], hi = row[8 + j];
    rerotate8(lo, hi, cos_delta + (int64_t)m * 64, sin_delta + (int64_t)m * 64, j, delta < 0 ? -1.0f : 1.0f);
    row[j] = lo;
    row[8 + j] = hi;
}

static int check_arena(const svln_kv_arena* a, const char* who) {
    SVLN_REQUIRE(a && a->base, SVLN_ERR_INVALID, "%s: null arena", who);
    SVLN_REQUIRE(a->head_dim == 128, SVLN_ERR_INVALID, "%s: head_dim %d != 128", who, a->head_dim);
    SVLN_REQUIRE(a->n_env > 0 && a->n_layer > 0 && a->n_kv_head > 0 && a->max_len > 0, SVLN_ERR_INVALID,
                 "%s: bad arena shape", who);
    SVLN_REQUIRE(((uintptr_t)a->base & 15) == 0, SVLN_ERR_INVALID, "%s: arena base not 16-byte aligned", who);
    return SVLN_OK;
}

}  // namespace svln

using namespace svln;

extern "C" {

size_t svln_kv_arena_bytes(int n_env, int n_layer, int n_kv_head, int head_dim, int max_len) {
    if (n_env <= 0 || n_layer <= 0 || n_kv_head <= 0 || head_dim <= 0 || max_len <= 0) return 0;
    return (size_t)n_env * n_layer * 2 * n_kv_head * max_len * head_dim * 2;
}

int svln_kv_append_rope(const svln_kv_arena* arena, int layer, const void* qkv, int n_q_head,
                        const int32_t* tok_env, const int32_t* tok_pos, int n_tok, const void* cos,
                        const void* sin, int n_pos, void* q_out, svln_stream_t stream) {
    int rc = check_arena(arena, "svln_kv_append_rope");
    if (rc) return rc;
    SVLN_REQUIRE(layer >= 0 && layer < arena->n_layer, SVLN_ERR_INVALID, "svln_kv_append_rope: layer %d out of range", layer);
    SVLN_REQUIRE(n_q_head > 0 && n_tok >= 0 && n_pos > 0, SVLN_ERR_INVALID, "svln_kv_append_rope: bad sizes");
    if (n_tok == 0) return SVLN_OK;
    SVLN_REQUIRE(qkv && tok_env && tok_pos && cos && sin && q_out, SVLN_ERR_INVALID, "svln_kv_append_rope: null pointer");
    const int64_t vecs = (int64_t)n_tok * (n_q_head + 2 * arena->n_kv_head);
    const int64_t blocks = (vecs * 8 + 255) / 256;
    kv_append_rope_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        *arena, layer, (const int4*)qkv, n_q_head, tok_env, tok_pos, n_tok, (const int4*)cos, (const int4*)sin,
        n_pos, (int4*)q_out);
    return check_launch("kv_append_rope_kernel");
}

int svln_kv_compact(const svln_kv_arena* arena, const int32_t* job_env, const int32_t* first_row,
                    const int32_t* new_len, const int32_t* src_row, int n_jobs, int rerotate,
                    const float* cos_delta, const float* sin_delta, int n_delta, svln_stream_t stream) {
    int rc = check_arena(arena, "svln_kv_compact");
    if (rc) return rc;
    SVLN_REQUIRE(n_jobs >= 0, SVLN_ERR_INVALID, "svln_kv_compact: negative n_jobs");
    if (n_jobs == 0) return SVLN_OK;
    SVLN_REQUIRE(job_env && first_row && new_len && src_row, SVLN_ERR_INVALID, "svln_kv_compact: null pointer");
    SVLN_REQUIRE(!rerotate || (cos_delta && sin_delta && n_delta > 0), SVLN_ERR_INVALID,
                 "svln_kv_compact: rerotate needs delta tables");
    SVLN_REQUIRE(n_jobs <= 65535, SVLN_ERR_RANGE, "svln_kv_compact: too many jobs");
    dim3 grid(arena->n_layer * 2 * arena->n_kv_head, n_jobs);
    kv_compact_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*arena, job_env, first_row, new_len, src_row,
                                                              rerotate, cos_delta, sin_delta, n_delta);
    return check_launch("kv_compact_kernel");
}

int svln_kv_rerotate(const svln_kv_arena* arena, int env, int row0, int n_rows, const int32_t* delta_pos,
                     const float* cos_delta, const float* sin_delta, int n_delta, svln_stream_t stream) {
    int rc = check_arena(arena, "svln_kv_rerotate");
    if (rc) return rc;
    SVLN_REQUIRE(env >= 0 && env < arena->n_env, SVLN_ERR_INVALID, "svln_kv_rerotate: env %d out of range", env);
    SVLN_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= arena->max_len, SVLN_ERR_INVALID,
                 "svln_kv_rerotate: rows [%d,%d) outside arena", row0, row0 + n_rows);
    if (n_rows == 0) return SVLN_OK;
    SVLN_REQUIRE(delta_pos && cos_delta && sin_delta && n_delta > 0, SVLN_ERR_INVALID, "svln_kv_rerotate: null pointer");
    dim3 grid((n_rows + 31) / 32, arena->n_layer * arena->n_kv_head);
    kv_rerotate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*arena, env, row0, n_rows, delta_pos, cos_delta,
                                                               sin_delta, n_delta);
    return check_launch("kv_rerotate_kernel");
}

}  // extern "C"

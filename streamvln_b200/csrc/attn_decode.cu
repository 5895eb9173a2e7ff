// Split-KV attention for the action-token decode (and any short chunk): an HBM-streaming kernel.
// Replaces flash-attn 2's decode path for HF Qwen2 attention (streamvln_eval.py:526; reached
// from stream_video_vln.py:340).  GQA-packed: the g = n_q_head / n_kv_head query heads that
// share one KV head (and up to 16/g query tokens) form the 16 rows of one m16n8k16 tile, so
// every K/V byte is read from HBM once per (stream, kv head) and used by all of them.
// K/V tiles stream through a 3-stage cp.async ring with a 128-byte XOR swizzle; the scalar
// fp32 pipe could not keep up with HBM for g = 7 (3.5 FMA per byte), so the dot products run
// on the legacy tensor path (HMMA) — the tcgen05 path is reserved for the dense prefill.
#include <stdlib.h>

#include "attn_decode_body.cuh"

namespace svln {

__global__ void __launch_bounds__(DWARPS * 32)
attn_decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    attn_decode_body<0>(p, blockIdx.x, blockIdx.y, blockIdx.z, smem, threadIdx.x);
}

// combine the split partials of one row of one (row group, kv head); 128 threads = 128 dims.
// Split weights are computed once, cooperatively, into shared memory; the accumulation loop then
// issues independent loads (no dependent global reads), so it pipelines.
constexpr int MAX_SPLITS = 64;

__global__ void __launch_bounds__(HD)
attn_decode_combine_kernel(const DecodeParams p) {
    __shared__ float s_w[MAX_SPLITS];
    __shared__ float s_red[4];
    const int kvh = blockIdx.y, r = blockIdx.z;
    const int b = blockIdx.x / p.groups_per_seq, grp = blockIdx.x % p.groups_per_seq;
    const int g = p.n_q_head / p.arena.n_kv_head;
    const int nq = p.n_q[b], qs = p.q_start[b];
    const int tok0 = grp * p.tok_per_group;
    if (tok0 >= nq) return;
    const int ntok = min(p.tok_per_group, nq - tok0);
    if (r >= ntok * g) return;
    const int64_t slot0 = ((int64_t)blockIdx.x * p.arena.n_kv_head + kvh) * p.n_splits;
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
    // n_splits <= MAX_SPLITS <= 128 threads: thread s owns split s
    const float ms = d < p.n_splits ? p.part_m[(slot0 + d) * DROWS + r] : -INFINITY;
    float m = warp_max(ms);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    const float w = (d < p.n_splits && ms != -INFINITY) ? exp2f((ms - m) * p.scale_log2) : 0.f;
    if (d < p.n_splits) s_w[d] = w;
    float l = d < p.n_splits ? p.part_l[(slot0 + d) * DROWS + r] * w : 0.f;
    l = warp_sum(l);
    __syncthreads();
    if (lane == 0) s_red[warp] = l;
    __syncthreads();
    l = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    float acc = 0.f;
    const float* po = p.part_o + (slot0 * DROWS + r) * HD + d;
#pragma unroll 4
    for (int s = 0; s < p.n_splits; ++s) {
        const float ws = s_w[s];
        if (ws != 0.f) acc += po[(int64_t)s * DROWS * HD] * ws;  // empty splits never wrote part_o
    }
    const int tq = tok0 + r / g, hq = kvh * g + r % g;
    const float val = l > 0.f ? acc / l : 0.f;
    p.out[((int64_t)(qs + tq) * p.n_q_head + hq) * HD + d] = __float2bfloat16_rn(val);
}

}  // namespace svln

using namespace svln;

extern "C" {

size_t svln_attn_workspace_bytes(int n_q_rows_max, int n_q_head, int head_dim, int max_len) {
    if (n_q_rows_max <= 0 || n_q_head <= 0 || head_dim != 128 || max_len <= 0) return 0;
    // decode partial slots: (groups x kv heads x splits) <= groups*Hkv + 4*SMs (see svln_attn_decode);
    // groups <= rows, Hkv <= n_q_head
    const size_t slots = (size_t)n_q_rows_max * n_q_head + 4 * 160 + 64;
    const size_t decode = slots * DROWS * (HD + 2) * sizeof(float) + 1024;
    // split-KV prefill partials: up to 2 x SM-count slots of [128 rows][128 dims + 2] fp32
    const size_t prefill = (size_t)2 * 160 * 128 * (HD + 2) * sizeof(float);
    return decode > prefill ? decode : prefill;
}

int svln_attn_decode(const svln_kv_arena* arena, int layer, const void* q, int n_q_head, const int32_t* seq_env,
                     const int32_t* q_start, const int32_t* n_q, const int32_t* kv_len, int n_seq, int max_n_q,
                     int max_kv_len, float scale, void* workspace, size_t workspace_bytes, void* out,
                     svln_stream_t stream) {
    SVLN_REQUIRE(arena && arena->base && arena->head_dim == 128, SVLN_ERR_INVALID, "svln_attn_decode: bad arena");
    SVLN_REQUIRE(layer >= 0 && layer < arena->n_layer, SVLN_ERR_INVALID, "svln_attn_decode: layer out of range");
    SVLN_REQUIRE(n_q_head > 0 && n_q_head % arena->n_kv_head == 0, SVLN_ERR_INVALID,
                 "svln_attn_decode: n_q_head %d not a multiple of n_kv_head %d", n_q_head, arena->n_kv_head);
    const int g = n_q_head / arena->n_kv_head;
    SVLN_REQUIRE(g <= DROWS, SVLN_ERR_RANGE, "svln_attn_decode: GQA group %d > 16", g);
    SVLN_REQUIRE(n_seq >= 0 && max_n_q >= 1 && max_kv_len >= 1 && max_kv_len <= arena->max_len, SVLN_ERR_INVALID,
                 "svln_attn_decode: bad sizes n_seq=%d max_n_q=%d max_kv_len=%d", n_seq, max_n_q, max_kv_len);
    if (n_seq == 0) return SVLN_OK;
    SVLN_REQUIRE(q && seq_env && q_start && n_q && kv_len && out && workspace, SVLN_ERR_INVALID,
                 "svln_attn_decode: null pointer");
    DecodeParams p;
    p.trace = nullptr;
    p.arena = *arena;
    p.layer = layer;
    p.q = (const __nv_bfloat16*)q;
    p.n_q_head = n_q_head;
    p.seq_env = seq_env; p.q_start = q_start; p.n_q = n_q; p.kv_len = kv_len;
    p.tok_per_group = DROWS / g;
    p.groups_per_seq = (max_n_q + p.tok_per_group - 1) / p.tok_per_group;
    const int64_t groups = (int64_t)n_seq * p.groups_per_seq;
    const int64_t base_ctas = groups * arena->n_kv_head;
    const int tiles = (max_kv_len + DT - 1) / DT;
    // aim for ~2 CTAs per SM (96 KB smem each), but give every CTA at least 4 tiles (256 keys, 128 KB of
    // K+V) so the per-CTA prologue / merge / partial write is amortised; at most MAX_SPLITS splits
    int64_t want = (2ll * sm_count() + base_ctas - 1) / base_ctas;
    if (want < 1) want = 1;
    if (want > MAX_SPLITS) want = MAX_SPLITS;
    int tiles_per_split = (int)((tiles + want - 1) / want);
    if (tiles_per_split < 4) tiles_per_split = tiles < 4 ? tiles : 4;
    p.n_splits = (tiles + tiles_per_split - 1) / tiles_per_split;
    // With >= 1.5 CTAs per SM before any split (64 streams x 4 kv heads = 256) one split per (stream, kv head) is best:
    // it writes the final rows itself, so the partials and the combine launch disappear (measured at 64 x 3000 keys,
    // L2 flushed: 1 / 2 / 3 / 4 / 6 splits + combine = 88.4 / 89.0 / 91.6 / 92.6 / 95.1 us; 1 split without combine: see
    // profiles/).  SVLN_AD_SPLITS overrides (experiments).
    static const int splits_env = getenv("SVLN_AD_SPLITS") ? atoi(getenv("SVLN_AD_SPLITS")) : 0;
    const int force = splits_env > 0 ? splits_env : (2 * base_ctas >= 3ll * sm_count() ? 1 : 0);
    if (force > 0) {
        tiles_per_split = (tiles + force - 1) / force;
        p.n_splits = (tiles + tiles_per_split - 1) / tiles_per_split;
    }
    p.split_keys = tiles_per_split * DT;
    p.scale_log2 = scale * 1.4426950408889634f;
    const size_t slots = (size_t)base_ctas * p.n_splits;
    const size_t need = slots * DROWS * (HD + 2) * sizeof(float);
    SVLN_REQUIRE(workspace_bytes >= need, SVLN_ERR_NOMEM, "svln_attn_decode: workspace %zu < %zu bytes",
                 workspace_bytes, need);
    SVLN_REQUIRE(groups <= 65535, SVLN_ERR_RANGE, "svln_attn_decode: too many row groups");
    p.part_o = (float*)workspace;
    p.part_m = p.part_o + slots * DROWS * HD;
    p.part_l = p.part_m + slots * DROWS;
    p.out = (__nv_bfloat16*)out;
    static unsigned long long attr_set = 0;
    if (first_use_on_device(&attr_set)) {
        SVLN_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM));
    }
    dim3 grid(p.n_splits, arena->n_kv_head, (unsigned)groups);
    attn_decode_kernel<<<grid, DWARPS * 32, DEC_SMEM, (cudaStream_t)stream>>>(p);
    if (p.n_splits > 1) {
        dim3 cgrid((unsigned)groups, arena->n_kv_head, DROWS);
        attn_decode_combine_kernel<<<cgrid, HD, 0, (cudaStream_t)stream>>>(p);
    }
    return check_launch("attn_decode kernels");
}

}  // extern "C"

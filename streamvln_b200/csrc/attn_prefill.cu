// Dense prefill attention on 5th-gen tensor cores: TMA-fed tcgen05.mma with TMEM accumulators.
// Replaces flash-attn 2 (mma.sync) for the new-token prefill of HF Qwen2 attention
// (streamvln_eval.py:526; reached from stream_video_vln.py:340): per turn ~216 new tokens, at a
// window start ~1.8-2k, against [pruned memory | window KV | new tokens] in the KV arena.
//
// One CTA = one 128-row M tile of one (sequence, kv head).  Rows are the flattened
// (query token, head-in-GQA-group) pairs R = tq * g + hg, so the g query heads that share a KV
// head reuse every K/V tile (GQA packing) and causal limits are monotone in R.
//
//   warps 0-7 : softmax.  Warps w and w+4 share TMEM lanes 32(w%4)..+31 (= tile rows) and take 64 columns
//               each, so every SM sub-partition runs two softmax warps (MUFU / FMA latency of one hides
//               behind the other) and a row's max / sum needs one 2-thread exchange through shared
//               memory (64-thread named barrier), no shuffles.  Loads S with tcgen05.ld, writes P (bf16)
//               to swizzled smem, rescales O in TMEM lazily (only when the running max grew by > 2^8).
//   warp 8    : one elected thread issues tcgen05.mma: S = Q K^T (SS, K-major x K-major) into a
//               double-buffered S in TMEM, and O += P V (SS, K-major x MN-major).
//   warp 9    : one elected thread issues TMA loads of K and V tiles (128 keys x 128 dims, two
//               64-column SWIZZLE_128B panels each) into 2-stage rings.
// TMEM: S0 cols [0,128), S1 [128,256), O [256,384) of a 512-column allocation.
#include <stdlib.h>

#include "tcgen05_helpers.cuh"

namespace svln {

constexpr int PF_BM = 128;   // rows per CTA tile
constexpr int PF_BN = 128;   // keys per KV tile
constexpr int PF_HD = 128;   // head_dim
constexpr int PF_PANEL = 128 * 128;            // bytes of one [128 rows][64 elem] SW128 panel
constexpr int PF_TILE = 2 * PF_PANEL;          // 32 KB: a [128][128] bf16 tile = two panels
constexpr int PF_SOFTMAX_WARPS = 8;  // two threads per tile row (64 columns each): 2 warps per SM sub-partition
constexpr int PF_THREADS = (PF_SOFTMAX_WARPS + 2) * 32;
constexpr int PF_MAX_SEQ = 64;

// shared memory map (dynamic, 1024-aligned)
constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + PF_TILE;           // 2 stages
constexpr int SM_V = SM_K + 2 * PF_TILE;       // 2 stages
constexpr int SM_P = SM_V + 2 * PF_TILE;       // 2 buffers: softmax(j+1) writes P while PV(j) still reads P(j)
constexpr int SM_BAR = SM_P + 2 * PF_TILE;     // barriers + tmem pointer
constexpr int SM_XCH = SM_BAR + 256;           // float [2 tiles][2 halves][128 rows] row-max exchange (reused for the row sums)
constexpr int PF_SMEM = SM_XCH + 2 * 2 * 128 * 4;  // 231 680 B of the 232 448 B a CTA may own: the base must be 1024-aligned

enum { BAR_K_FULL = 0, BAR_K_EMPTY = 2, BAR_V_FULL = 4, BAR_V_EMPTY = 6, BAR_S_FULL = 8, BAR_S_EMPTY = 10,
       BAR_P_FULL = 12, BAR_PV_DONE = 14, BAR_COUNT = 16 };

struct PrefillParams {
    const __nv_bfloat16* q;   // [n_tok, Hq, 128]
    __nv_bfloat16* out;       // [n_tok, Hq*128]
    int n_q_head, n_kv_head, n_layer, layer, max_len;
    int n_seq;
    int debug;  // SVLN_PF_DEBUG bit flags for bottleneck experiments (0 in production): 1 = no MUFU, 2 = no P store
    float scale_log2;
    // sequence descriptors: device arrays when d_env != nullptr (CUDA-graph replayable: only the grid
    // shape, i.e. n_q, is baked into the launch), else the by-value copies below
    const int32_t *d_env, *d_q_start, *d_n_q, *d_kv_len;
    int32_t env[PF_MAX_SEQ], q_start[PF_MAX_SEQ], n_q[PF_MAX_SEQ], kv_len[PF_MAX_SEQ];
    int32_t tile_start[PF_MAX_SEQ + 1];  // prefix sum of M tiles per sequence (from the host-known n_q)
};

// byte offset of 16-byte chunk c (0..15) of row r inside a K-major [128][128] SW128 tile (two panels)
__device__ __forceinline__ uint32_t kmajor_off(int r, int c) {
    return (uint32_t)((c >> 3) * PF_PANEL + r * 128 + (((c & 7) ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(PF_THREADS, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap kv_map, const __grid_constant__ PrefillParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = s2u(smem);
    if ((sbase & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need a 1024-byte aligned base (no static smem in this kernel)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM_BAR + BAR_COUNT * 8);
    const uint32_t bar0 = sbase + SM_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kvh = blockIdx.y;
    const int g = p.n_q_head / p.n_kv_head;

    // locate (sequence, M tile)
    // heaviest tiles first: within a sequence later M tiles see more keys (causal), so walk the grid backwards
    const int gtile = (int)gridDim.x - 1 - (int)blockIdx.x;
    int b = 0;
    while (b + 1 < p.n_seq && gtile >= p.tile_start[b + 1]) ++b;
    const int mtile = gtile - p.tile_start[b];
    const bool dev = p.d_env != nullptr;
    const int nq = dev ? p.d_n_q[b] : p.n_q[b];
    const int len = dev ? p.d_kv_len[b] : p.kv_len[b];
    const int env = dev ? p.d_env[b] : p.env[b];
    const int qs = dev ? p.d_q_start[b] : p.q_start[b];
    const int rows_total = nq * g;
    const int row0 = mtile * PF_BM;
    const int rows_here = min(PF_BM, rows_total - row0);
    const int tq_last = (row0 + rows_here - 1) / g;
    const int kv_hi = len - nq + tq_last + 1;          // keys [0, kv_hi) are visible to some row
    const int n_tiles = (kv_hi + PF_BN - 1) / PF_BN;
    const int plane_k = ((env * p.n_layer + p.layer) * 2 + 0) * p.n_kv_head + kvh;
    const int plane_v = plane_k + p.n_kv_head;

    if (tid == 0) {
        mbar_init(BAR(BAR_K_FULL), 1); mbar_init(BAR(BAR_K_FULL + 1), 1);
        mbar_init(BAR(BAR_K_EMPTY), 1); mbar_init(BAR(BAR_K_EMPTY + 1), 1);
        mbar_init(BAR(BAR_V_FULL), 1); mbar_init(BAR(BAR_V_FULL + 1), 1);
        mbar_init(BAR(BAR_V_EMPTY), 1); mbar_init(BAR(BAR_V_EMPTY + 1), 1);
        mbar_init(BAR(BAR_S_FULL), 1); mbar_init(BAR(BAR_S_FULL + 1), 1);
        // one arrival per softmax warp (elected lane after __syncwarp), not one per thread: 512 shared-memory
        // atomics per tile would serialise on the barrier word
        mbar_init(BAR(BAR_S_EMPTY), PF_SOFTMAX_WARPS); mbar_init(BAR(BAR_S_EMPTY + 1), PF_SOFTMAX_WARPS);
        mbar_init(BAR(BAR_P_FULL), PF_SOFTMAX_WARPS); mbar_init(BAR(BAR_P_FULL + 1), PF_SOFTMAX_WARPS);
        mbar_init(BAR(BAR_PV_DONE), 1); mbar_init(BAR(BAR_PV_DONE + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM allocation: 512 columns, by one full warp (which also frees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Q tile -> swizzled smem (generic proxy), rows past the end are zero.  Thread mapping = softmax mapping:
    // row 32*(warp%4)+lane, column half warp/4 (eight 16-byte chunks).
    if (warp < PF_SOFTMAX_WARPS) {
        const int r = (warp & 3) * 32 + lane, half = warp >> 2;
        const int R = row0 + r;
        int4 v[8];
        if (r < rows_here) {
            const int tq = R / g, hg = R - tq * g;
            const int4* src = reinterpret_cast<const int4*>(p.q + ((int64_t)(qs + tq) * p.n_q_head + kvh * g + hg) * PF_HD) + half * 8;
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __ldg(src + c);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) sts128(sbase + SM_Q + kmajor_off(r, half * 8 + c), v[c]);
    }
    fence_async_smem();  // make the generic-proxy smem writes visible to the async proxy (MMA)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t TM_S0 = tmem_base, TM_O = tmem_base + 256;

    if (warp == PF_SOFTMAX_WARPS + 1) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&kv_map) : "memory");
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j & 1, u = j >> 1;
                const int key0 = j * PF_BN;
                mbar_wait(BAR(BAR_K_EMPTY + s), (u & 1) ^ 1);
                mbar_expect_tx(BAR(BAR_K_FULL + s), PF_TILE);
                tma_load_3d(sbase + SM_K + s * PF_TILE, &kv_map, BAR(BAR_K_FULL + s), 0, key0, plane_k);
                tma_load_3d(sbase + SM_K + s * PF_TILE + PF_PANEL, &kv_map, BAR(BAR_K_FULL + s), 64, key0, plane_k);
                mbar_wait(BAR(BAR_V_EMPTY + s), (u & 1) ^ 1);
                mbar_expect_tx(BAR(BAR_V_FULL + s), PF_TILE);
                tma_load_3d(sbase + SM_V + s * PF_TILE, &kv_map, BAR(BAR_V_FULL + s), 0, key0, plane_v);
                tma_load_3d(sbase + SM_V + s * PF_TILE + PF_PANEL, &kv_map, BAR(BAR_V_FULL + s), 64, key0, plane_v);
            }
        }
    } else if (warp == PF_SOFTMAX_WARPS) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t IDESC_QK = umma_idesc(0);  // A, B K-major
            constexpr uint32_t IDESC_PV = umma_idesc(1);  // B (V) MN-major
            auto issue_qk = [&](int j) {
                const int s = j & 1;
                const uint32_t kb = sbase + SM_K + s * PF_TILE;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {  // 16 dims per step; 4 steps per 64-dim panel
                    const uint32_t off = (ks >> 2) * PF_PANEL + (ks & 3) * 32;
                    tc_mma(TM_S0 + s * 128, umma_desc(sbase + SM_Q + off, 16, 1024), umma_desc(kb + off, 16, 1024),
                           IDESC_QK, ks > 0);
                }
            };
            mbar_wait(BAR(BAR_K_FULL), 0);
            tc_fence_after();
            issue_qk(0);
            tc_commit(BAR(BAR_K_EMPTY));
            tc_commit(BAR(BAR_S_FULL));
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) {  // S(j+1) overlaps softmax(j)
                    const int s = (j + 1) & 1, u = (j + 1) >> 1;
                    mbar_wait(BAR(BAR_K_FULL + s), u & 1);
                    mbar_wait(BAR(BAR_S_EMPTY + s), (u & 1) ^ 1);
                    tc_fence_after();
                    issue_qk(j + 1);
                    tc_commit(BAR(BAR_K_EMPTY + s));
                    tc_commit(BAR(BAR_S_FULL + s));
                }
                const int s = j & 1, u = j >> 1;
                mbar_wait(BAR(BAR_P_FULL + s), u & 1);
                mbar_wait(BAR(BAR_V_FULL + s), u & 1);
                tc_fence_after();
                const uint32_t vb = sbase + SM_V + s * PF_TILE;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {  // 16 keys per step
                    // A = P: K-major (keys contiguous), 4 steps per 64-key panel
                    const uint32_t aoff = (ks >> 2) * PF_PANEL + (ks & 3) * 32;
                    // B = V: MN-major; 16 keys = two 8-row swizzle atoms = 2048 B; dims panels 16 KB apart
                    tc_mma(TM_O, umma_desc(sbase + SM_P + s * PF_TILE + aoff, 16, 1024), umma_desc(vb + ks * 2048, PF_PANEL, 1024),
                           IDESC_PV, (j > 0 || ks > 0));
                }
                tc_commit(BAR(BAR_V_EMPTY + s));
                tc_commit(BAR(BAR_PV_DONE + s));
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warps
        const int r = (warp & 3) * 32 + lane;    // tile row == TMEM lane
        const int half = warp >> 2;              // columns [64*half, 64*half + 64)
        const int R = row0 + r;
        const bool row_valid = r < rows_here;
        const int tq = row_valid ? R / g : 0;
        const int hg = row_valid ? R - tq * g : 0;
        const int lim = row_valid ? len - nq + tq + 1 : 0;   // keys < lim are visible
        const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 64);
        const uint32_t xch = sbase + SM_XCH;                   // float [tile parity][half][row], shared-space address
        const int pair_bar = 1 + (warp & 3);                   // named barrier shared by warps w and w+4
        float m_used = -INFINITY;   // running reference max in scaled (log2) units; identical in both halves
        float l_sum = 0.f;          // this thread's half of the row sum
        for (int j = 0; j < n_tiles; ++j) {
            const int s = j & 1, u = j >> 1;
            mbar_wait(BAR(BAR_S_FULL + s), u & 1);
            tc_fence_after();
            uint32_t sv[2][32];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld32(TM_S0 + s * 128 + lane_addr + c * 32, sv[c]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_S_EMPTY + s));
            // mask + row max on the raw scores (scale > 0 commutes with max)
            const int key0 = j * PF_BN + half * 64;
            float mx = -INFINITY;
            if (key0 + 64 > lim) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (key0 + c * 32 + e >= lim) sv[c][e] = 0xFF800000u;  // -inf
                        mx = fmaxf(mx, __uint_as_float(sv[c][e]));
                    }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(sv[c][e]));
            }
            sts32f(xch + ((s * 2 + half) * 128 + r) * 4, mx);
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            mx = fmaxf(mx, lds32f(xch + ((s * 2 + (half ^ 1)) * 128 + r) * 4)) * p.scale_log2;
            // lazy rescale: only move the reference max when it grew by more than 8 (factor 256)
            float alpha = 1.f;
            bool rescale = false;
            if (mx > m_used + 8.0f || (m_used == -INFINITY && mx != -INFINITY)) {
                alpha = m_used == -INFINITY ? 0.f : exp2f(m_used - mx);
                m_used = mx;
                rescale = j > 0;
            }
            const float neg_m = m_used == -INFINITY ? 0.f : -m_used;
            // P buffer s is free once PV(j-2) has completed (long ago in steady state: PV(j-1) is what may still run)
            if (j > 1) mbar_wait(BAR(BAR_PV_DONE + s), (u - 1) & 1);
            l_sum *= alpha;
            if (__any_sync(0xffffffffu, rescale)) {
                // rare (lazy rescale): O must be quiescent, i.e. PV(j-1) done; PV(j) cannot start before our P(j)
                mbar_wait(BAR(BAR_PV_DONE + (s ^ 1)), ((j - 1) >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t ov[32];
                    tmem_ld32(TM_O + lane_addr + c * 32, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                    tmem_st32(TM_O + lane_addr + c * 32, ov);
                }
                tmem_wait_st();
            }
            // P = exp2(S * scale - m): one FFMA + one MUFU per element; row sum in fp32; bf16 into swizzled smem
            float psum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // 8 keys = one 16-byte chunk
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float p0 = fmaf(__uint_as_float(sv[c][q * 8 + 2 * e]), p.scale_log2, neg_m);
                        float p1 = fmaf(__uint_as_float(sv[c][q * 8 + 2 * e + 1]), p.scale_log2, neg_m);
                        if (!(p.debug & 1)) { p0 = ex2_approx(p0); p1 = ex2_approx(p1); }
                        psum += p0 + p1;
                        w[e] = pack2_bf16(p0, p1);
                    }
                    if (!(p.debug & 2))
                        sts128(sbase + SM_P + s * PF_TILE + kmajor_off(r, half * 8 + c * 4 + q), make_int4(w[0], w[1], w[2], w[3]));
                }
            }
            l_sum += psum;
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_P_FULL + s));
        }
        // epilogue: O / l -> bf16 -> global (this thread: 64 dims of its row)
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // partner has read the last row max: reuse the block
        sts32f(xch + (half * 128 + r) * 4, l_sum);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float l_tot = l_sum + lds32f(xch + ((half ^ 1) * 128 + r) * 4);
        mbar_wait(BAR(BAR_PV_DONE + ((n_tiles - 1) & 1)), ((n_tiles - 1) >> 1) & 1);
        tc_fence_after();
        const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
        __nv_bfloat16* orow = p.out + ((int64_t)(qs + tq) * p.n_q_head + kvh * g + hg) * PF_HD + half * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld32(TM_O + lane_addr + c * 32, ov);
            tmem_wait_ld();
            if (row_valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = pack2_bf16(__uint_as_float(ov[q * 8 + 2 * e]) * inv_l, __uint_as_float(ov[q * 8 + 2 * e + 1]) * inv_l);
                    *reinterpret_cast<int4*>(orow + c * 32 + q * 8) = make_int4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------ host side
struct MapCacheEntry {
    void* base; int n_env, n_layer, n_kv_head, max_len, box_rows; CUtensorMap map; bool valid;
};
static MapCacheEntry g_map_cache[8];

// 3-D tensor map (128 dims, max_len rows, planes) over the whole arena, box = 64 dims x box_rows keys, SWIZZLE_128B
int get_kv_map(const svln_kv_arena* a, int box_rows, CUtensorMap* out) {
    for (auto& e : g_map_cache)
        if (e.valid && e.base == a->base && e.n_env == a->n_env && e.n_layer == a->n_layer &&
            e.n_kv_head == a->n_kv_head && e.max_len == a->max_len && e.box_rows == box_rows) {
            *out = e.map;
            return SVLN_OK;
        }
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SVLN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        SVLN_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, SVLN_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    }
    const cuuint64_t planes = (cuuint64_t)a->n_env * a->n_layer * 2 * a->n_kv_head;
    cuuint64_t gdim[3] = {(cuuint64_t)PF_HD, (cuuint64_t)a->max_len, planes};
    cuuint64_t gstride[2] = {(cuuint64_t)PF_HD * 2, (cuuint64_t)a->max_len * PF_HD * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estride[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->base, gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SVLN_REQUIRE(r == CUDA_SUCCESS, SVLN_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    static int next = 0;
    MapCacheEntry& e = g_map_cache[next++ % 8];
    e = {a->base, a->n_env, a->n_layer, a->n_kv_head, a->max_len, box_rows, m, true};
    *out = m;
    return SVLN_OK;
}

static int prefill_launch(const svln_kv_arena* arena, int layer, const void* q, int n_q_head, const int32_t* env_h,
                          const int32_t* q_start_h, const int32_t* n_q_h, const int32_t* kv_len_h, const int32_t* env_d,
                          const int32_t* q_start_d, const int32_t* n_q_d, const int32_t* kv_len_d, int n_seq, float scale,
                          void* out, svln_stream_t stream, const char* who, void* workspace, size_t workspace_bytes) {
    SVLN_REQUIRE(arena && arena->base && arena->head_dim == 128, SVLN_ERR_INVALID, "%s: bad arena", who);
    SVLN_REQUIRE(layer >= 0 && layer < arena->n_layer, SVLN_ERR_INVALID, "%s: layer out of range", who);
    SVLN_REQUIRE(n_q_head > 0 && n_q_head % arena->n_kv_head == 0, SVLN_ERR_INVALID,
                 "%s: n_q_head %d not a multiple of n_kv_head %d", who, n_q_head, arena->n_kv_head);
    SVLN_REQUIRE(n_seq >= 0, SVLN_ERR_INVALID, "%s: negative n_seq", who);
    if (n_seq == 0) return SVLN_OK;
    SVLN_REQUIRE(q && out && n_q_h, SVLN_ERR_INVALID, "%s: null pointer", who);
    SVLN_REQUIRE(((uintptr_t)arena->base & 15) == 0, SVLN_ERR_INVALID, "%s: arena base must be 16-byte aligned", who);
    const bool dev = env_d != nullptr;
    SVLN_REQUIRE(dev ? (q_start_d && n_q_d && kv_len_d) : (env_h && q_start_h && kv_len_h), SVLN_ERR_INVALID,
                 "%s: null descriptor array", who);
    static const int version = getenv("SVLN_PF_V") ? atoi(getenv("SVLN_PF_V")) : 5;  // 3: SS 128 rows, 4: SS 256 rows, 5: TS 128 rows (default)
    if (version >= 5)  // Q and P in tensor memory (TS-form MMAs): shared memory only carries K and V
        return prefill3_launch(arena, layer, q, n_q_head, env_h, q_start_h, n_q_h, kv_len_h, env_d, q_start_d, n_q_d, kv_len_d,
                               n_seq, scale, out, stream, who, workspace, workspace_bytes);
    if (version == 4)  // 256 rows per CTA (two Q tiles share every K/V tile): half the L2->SM bytes per FLOP
        return prefill2_launch(arena, layer, q, n_q_head, env_h, q_start_h, n_q_h, kv_len_h, env_d, q_start_d, n_q_d, kv_len_d,
                               n_seq, scale, out, stream, who);
    CUtensorMap map;
    int rc = get_kv_map(arena, PF_BN, &map);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        SVLN_CUDA(cudaFuncSetAttribute(attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
        attr_set = true;
    }
    const int g = n_q_head / arena->n_kv_head;
    for (int s0 = 0; s0 < n_seq; s0 += PF_MAX_SEQ) {
        PrefillParams p;
        p.q = (const __nv_bfloat16*)q;
        p.out = (__nv_bfloat16*)out;
        p.n_q_head = n_q_head; p.n_kv_head = arena->n_kv_head; p.n_layer = arena->n_layer; p.layer = layer;
        p.max_len = arena->max_len;
        p.scale_log2 = scale * 1.4426950408889634f;
        static const int dbg = getenv("SVLN_PF_DEBUG") ? atoi(getenv("SVLN_PF_DEBUG")) : 0;
        p.debug = dbg;
        const int ns = n_seq - s0 < PF_MAX_SEQ ? n_seq - s0 : PF_MAX_SEQ;
        p.n_seq = ns;
        p.d_env = dev ? env_d + s0 : nullptr;
        p.d_q_start = dev ? q_start_d + s0 : nullptr;
        p.d_n_q = dev ? n_q_d + s0 : nullptr;
        p.d_kv_len = dev ? kv_len_d + s0 : nullptr;
        int tiles = 0;
        for (int i = 0; i < ns; ++i) {
            const int b = s0 + i;
            SVLN_REQUIRE(n_q_h[b] >= 1 && n_q_h[b] <= arena->max_len, SVLN_ERR_INVALID, "%s: bad n_q=%d", who, n_q_h[b]);
            if (!dev) {
                SVLN_REQUIRE(env_h[b] >= 0 && env_h[b] < arena->n_env, SVLN_ERR_INVALID, "%s: env %d out of range", who, env_h[b]);
                SVLN_REQUIRE(kv_len_h[b] >= n_q_h[b] && kv_len_h[b] <= arena->max_len, SVLN_ERR_INVALID,
                             "%s: bad lengths n_q=%d kv_len=%d", who, n_q_h[b], kv_len_h[b]);
                p.env[i] = env_h[b]; p.q_start[i] = q_start_h[b]; p.kv_len[i] = kv_len_h[b];
            } else {
                p.env[i] = p.q_start[i] = p.kv_len[i] = 0;
            }
            p.n_q[i] = n_q_h[b];
            p.tile_start[i] = tiles;
            tiles += (n_q_h[b] * g + PF_BM - 1) / PF_BM;
        }
        p.tile_start[ns] = tiles;
        for (int i = ns + 1; i <= PF_MAX_SEQ; ++i) p.tile_start[i] = tiles;
        for (int i = ns; i < PF_MAX_SEQ; ++i) p.env[i] = p.q_start[i] = p.n_q[i] = p.kv_len[i] = 0;
        dim3 grid(tiles, arena->n_kv_head);
        attn_prefill_kernel<<<grid, PF_THREADS, PF_SMEM, (cudaStream_t)stream>>>(map, p);
        rc = check_launch("attn_prefill_kernel");
        if (rc) return rc;
    }
    return SVLN_OK;
}

}  // namespace svln

using namespace svln;

extern "C" {

int svln_attn_prefill(const svln_kv_arena* arena, int layer, const void* q, int n_q_head,
                      const int32_t* seq_env_host, const int32_t* q_start_host, const int32_t* n_q_host,
                      const int32_t* kv_len_host, int n_seq, float scale, void* workspace, size_t workspace_bytes,
                      void* out, svln_stream_t stream) {
    return prefill_launch(arena, layer, q, n_q_head, seq_env_host, q_start_host, n_q_host, kv_len_host, nullptr, nullptr,
                          nullptr, nullptr, n_seq, scale, out, stream, "svln_attn_prefill", workspace, workspace_bytes);
}

int svln_attn_prefill_dev(const svln_kv_arena* arena, int layer, const void* q, int n_q_head, const int32_t* seq_env,
                          const int32_t* q_start, const int32_t* n_q, const int32_t* kv_len, const int32_t* n_q_host,
                          int n_seq, float scale, void* workspace, size_t workspace_bytes, void* out, svln_stream_t stream) {
    SVLN_REQUIRE(seq_env, SVLN_ERR_INVALID, "svln_attn_prefill_dev: null seq_env");
    return prefill_launch(arena, layer, q, n_q_head, nullptr, nullptr, n_q_host, nullptr, seq_env, q_start, n_q, kv_len,
                          n_seq, scale, out, stream, "svln_attn_prefill_dev", workspace, workspace_bytes);
}

}  // extern "C"

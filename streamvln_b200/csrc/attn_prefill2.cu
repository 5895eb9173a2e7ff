// Dense prefill attention, 256 query rows per CTA: two 128-row Q tiles share every K/V tile.
//
// Why: the 128-row kernel (attn_prefill.cu) streams 64 KB of K+V through L2->SM per 8.4 MFLOP tile
// (128 FLOP/B) and saturates the L2->SM path at ~6.9 TB/s (measured: 893 TFLOP/s even with the softmax
// removed, profiles/README.md).  Here one K/V tile of 64 keys (32 KB) feeds 256 rows: 256 FLOP/B.
//
//   warps 0-3 : softmax of Q tile 0 (thread = tile row = TMEM lane; a full row of 64 scores per thread,
//               so no cross-thread reduction at all);  warps 4-7 : softmax of Q tile 1.
//   warp 8    : one thread issues tcgen05.mma: S_t = Q_t K^T (M128 N64 K16 x8) for both tiles, then
//               O_t += P_t V (M128 N128 K16 x4).  S and P are double-buffered per tile, so QK(j+1)
//               runs under softmax(j) and softmax(j+1) never waits for PV(j).
//   warp 9    : one thread issues TMA loads of K and V (64 keys x 128 dims, two SWIZZLE_128B panels each)
//               into a 3-stage ring.
// TMEM (512 cols): S[t][b] 64 cols each at 0/64/128/192, O_0 at 256, O_1 at 384.
// smem: Q 2x32 KB | K 3x16 KB | V 3x16 KB | P 2x2x16 KB = 224 KB + barriers.
#include <stdlib.h>

#include "tcgen05_helpers.cuh"

namespace svln {

constexpr int P2_BM = 128, P2_NQT = 2, P2_ROWS = P2_BM * P2_NQT, P2_BN = 64, P2_HD = 128;
constexpr int P2_STAGES = 3;
constexpr int P2_QTILE = 128 * 256;        // 32 KB: [2 panels][128 rows][128 B]
constexpr int P2_QPANEL = 128 * 128;       // 16 KB
constexpr int P2_KVTILE = P2_BN * 256;     // 16 KB: [2 panels][64 rows][128 B]
constexpr int P2_KVPANEL = P2_BN * 128;    // 8 KB
constexpr int P2_PTILE = 128 * 128;        // 16 KB: [128 rows][64 keys] one panel
constexpr int P2_SOFTMAX_WARPS = 8;
constexpr int P2_THREADS = (P2_SOFTMAX_WARPS + 2) * 32;
constexpr int P2_MAX_SEQ = 64;

constexpr int S2_Q = 0;
constexpr int S2_K = S2_Q + P2_NQT * P2_QTILE;
constexpr int S2_V = S2_K + P2_STAGES * P2_KVTILE;
constexpr int S2_P = S2_V + P2_STAGES * P2_KVTILE;      // [tile][buf]
constexpr int S2_BAR = S2_P + 4 * P2_PTILE;
constexpr int P2_SMEM = S2_BAR + 512;

enum { B2_K_FULL = 0, B2_K_EMPTY = 3, B2_V_FULL = 6, B2_V_EMPTY = 9, B2_S_FULL = 12, B2_S_EMPTY = 16, B2_P_FULL = 20,
       B2_PV_DONE = 24, B2_COUNT = 28 };   // S_*/P_*/PV_DONE indexed [tile * 2 + buf]

struct Prefill2Params {
    const __nv_bfloat16* q;
    __nv_bfloat16* out;
    int n_q_head, n_kv_head, n_layer, layer, max_len;
    int n_seq;
    float scale_log2;
    const int32_t *d_env, *d_q_start, *d_n_q, *d_kv_len;  // device descriptors (graph-replayable) or nullptr
    int32_t env[P2_MAX_SEQ], q_start[P2_MAX_SEQ], n_q[P2_MAX_SEQ], kv_len[P2_MAX_SEQ];
    int32_t tile_start[P2_MAX_SEQ + 1];                    // prefix sum of 256-row CTA tiles per sequence
    long long* trace;                                      // SVLN_PF_TRACE: per-phase clock64 stamps of CTA (0,0), else nullptr
};

// trace layout: [role 0..3][tile 0..31][slot 0..7]; roles: 0 softmax warp 0, 1 softmax warp 4, 2 MMA thread, 3 TMA thread
#define P2_TRACE(role, tile, slot)                                                                     \
    do {                                                                                               \
        if (p.trace && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && (tile) < 32 && ((role) < 2 || lane == 0)) \
            p.trace[((role) * 32 + (tile)) * 8 + (slot)] = clock64();                                  \
    } while (0)

__global__ void __launch_bounds__(P2_THREADS, 1)
attn_prefill2_kernel(const __grid_constant__ CUtensorMap kv_map, const __grid_constant__ Prefill2Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = s2u(smem);
    if ((sbase & 1023u) != 0) __trap();
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + S2_BAR + B2_COUNT * 8);
    const uint32_t bar0 = sbase + S2_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kvh = blockIdx.y;
    const int g = p.n_q_head / p.n_kv_head;
    // heaviest CTAs first (causal: later rows see more keys)
    const int gtile = (int)gridDim.x - 1 - (int)blockIdx.x;
    int b = 0;
    while (b + 1 < p.n_seq && gtile >= p.tile_start[b + 1]) ++b;
    const int ctile = gtile - p.tile_start[b];
    const bool dev = p.d_env != nullptr;
    const int nq = dev ? p.d_n_q[b] : p.n_q[b];
    const int len = dev ? p.d_kv_len[b] : p.kv_len[b];
    const int env = dev ? p.d_env[b] : p.env[b];
    const int qs = dev ? p.d_q_start[b] : p.q_start[b];
    const int rows_total = nq * g;
    const int row0 = ctile * P2_ROWS;
    const int rows_here = min(P2_ROWS, rows_total - row0);
    const int tq_last = (row0 + rows_here - 1) / g;
    const int kv_hi = len - nq + tq_last + 1;
    const int n_tiles = (kv_hi + P2_BN - 1) / P2_BN;
    const int plane_k = ((env * p.n_layer + p.layer) * 2 + 0) * p.n_kv_head + kvh;
    const int plane_v = plane_k + p.n_kv_head;

    if (tid == 0) {
        for (int i = 0; i < 12; ++i) mbar_init(BAR(i), 1);                               // K/V full/empty
        for (int i = 0; i < 4; ++i) {
            mbar_init(BAR(B2_S_FULL + i), 1);
            mbar_init(BAR(B2_S_EMPTY + i), 4);      // one elected arrival per softmax warp of the tile
            mbar_init(BAR(B2_P_FULL + i), 4);
            mbar_init(BAR(B2_PV_DONE + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Q tiles -> swizzled smem; thread (tile t = warp/4, row r) copies its own 256-byte row
    if (warp < P2_SOFTMAX_WARPS) {
        const int t = warp >> 2, r = (warp & 3) * 32 + lane;
        const int R = row0 + t * P2_BM + r;
        const uint32_t qbase = sbase + S2_Q + t * P2_QTILE;
        if (R < rows_total) {
            const int tq = R / g, hg = R - tq * g;
            const int4* src = reinterpret_cast<const int4*>(p.q + ((int64_t)(qs + tq) * p.n_q_head + kvh * g + hg) * P2_HD);
#pragma unroll
            for (int c0 = 0; c0 < 16; c0 += 8) {
                int4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __ldg(src + c0 + c);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    sts128(qbase + (c0 >> 3) * P2_QPANEL + r * 128 + ((c ^ (r & 7)) << 4), v[c]);
            }
        } else {
            const int4 z = make_int4(0, 0, 0, 0);
#pragma unroll
            for (int c = 0; c < 16; ++c) sts128(qbase + (c >> 3) * P2_QPANEL + r * 128 + (((c & 7) ^ (r & 7)) << 4), z);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    auto TM_S = [&](int t, int bf) { return tmem_base + (uint32_t)((t * 2 + bf) * 64); };
    auto TM_O = [&](int t) { return tmem_base + 256u + (uint32_t)(t * 128); };

    if (warp == P2_SOFTMAX_WARPS + 1) {
        // ------------------------------------------------------------ TMA producer
        // the whole warp runs the (uniform) control flow; one elected lane issues
        if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&kv_map) : "memory");
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % P2_STAGES, u = j / P2_STAGES;
            const int key0 = j * P2_BN;
            mbar_wait(BAR(B2_K_EMPTY + st), (u & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(BAR(B2_K_FULL + st), P2_KVTILE);
                tma_load_3d(sbase + S2_K + st * P2_KVTILE, &kv_map, BAR(B2_K_FULL + st), 0, key0, plane_k);
                tma_load_3d(sbase + S2_K + st * P2_KVTILE + P2_KVPANEL, &kv_map, BAR(B2_K_FULL + st), 64, key0, plane_k);
            }
            mbar_wait(BAR(B2_V_EMPTY + st), (u & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(BAR(B2_V_FULL + st), P2_KVTILE);
                tma_load_3d(sbase + S2_V + st * P2_KVTILE, &kv_map, BAR(B2_V_FULL + st), 0, key0, plane_v);
                tma_load_3d(sbase + S2_V + st * P2_KVTILE + P2_KVPANEL, &kv_map, BAR(B2_V_FULL + st), 64, key0, plane_v);
            }
            __syncwarp();
        }
    } else if (warp == P2_SOFTMAX_WARPS) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp executes this warp-uniform control flow and only the tcgen05 instructions sit under an
        // elected-lane predicate: descriptors then live in uniform registers.  (Issuing from `if (lane == 0)` made
        // ptxas wrap every UTCHMMA in an R2UR + waterfall loop: ~80 cycles per MMA, the measured bottleneck.)
        constexpr uint32_t IDESC_QK = umma_idesc(0, 64);    // S tile: N = 64 keys
        constexpr uint32_t IDESC_PV = umma_idesc(1, 128);   // O tile: N = 128 dims, V MN-major
        auto issue_qk = [&](int j) {
            const int st = j % P2_STAGES, bf = j & 1, ub = j >> 1;
            P2_TRACE(2, j, 0);
            mbar_wait(BAR(B2_K_FULL + st), (j / P2_STAGES) & 1);
            P2_TRACE(2, j, 1);
            const uint32_t kb = sbase + S2_K + st * P2_KVTILE;
#pragma unroll
            for (int t = 0; t < P2_NQT; ++t) {
                mbar_wait(BAR(B2_S_EMPTY + t * 2 + bf), (ub & 1) ^ 1);
                tc_fence_after();
                const uint32_t qb = sbase + S2_Q + t * P2_QTILE;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        tc_mma(TM_S(t, bf), umma_desc(qb + (ks >> 2) * P2_QPANEL + (ks & 3) * 32, 16, 1024),
                               umma_desc(kb + (ks >> 2) * P2_KVPANEL + (ks & 3) * 32, 16, 1024), IDESC_QK, ks > 0);
                    tc_commit(BAR(B2_S_FULL + t * 2 + bf));
                    if (t == P2_NQT - 1) tc_commit(BAR(B2_K_EMPTY + st));
                }
                __syncwarp();
            }
            P2_TRACE(2, j, 2);
        };
        issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) issue_qk(j + 1);
            const int st = j % P2_STAGES, bf = j & 1, ub = j >> 1;
            mbar_wait(BAR(B2_V_FULL + st), (j / P2_STAGES) & 1);
            P2_TRACE(2, j, 3);
            const uint32_t vb = sbase + S2_V + st * P2_KVTILE;
#pragma unroll
            for (int t = 0; t < P2_NQT; ++t) {
                mbar_wait(BAR(B2_P_FULL + t * 2 + bf), ub & 1);
                P2_TRACE(2, j, 4 + t);
                tc_fence_after();
                const uint32_t pb = sbase + S2_P + (t * 2 + bf) * P2_PTILE;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)   // 16 keys per step: P k-advance 32 B; V k-advance 16 rows x 128 B
                        tc_mma(TM_O(t), umma_desc(pb + ks * 32, 16, 1024), umma_desc(vb + ks * 2048, P2_KVPANEL, 1024), IDESC_PV,
                               (j > 0 || ks > 0));
                    tc_commit(BAR(B2_PV_DONE + t * 2 + bf));
                    if (t == P2_NQT - 1) tc_commit(BAR(B2_V_EMPTY + st));
                }
                __syncwarp();
            }
            P2_TRACE(2, j, 6);
        }
    } else {
        // ------------------------------------------------------------ softmax warps
        const int t = warp >> 2;
        const int r = (warp & 3) * 32 + lane;
        const int R = row0 + t * P2_BM + r;
        const bool row_valid = R < rows_total;
        const int tq = row_valid ? R / g : 0;
        const int hg = row_valid ? R - tq * g : 0;
        const int lim = row_valid ? len - nq + tq + 1 : 0;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        float m_used = -INFINITY, l_sum = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int bf = j & 1, ub = j >> 1;
            const int trole = (lane == 0 && (warp & 3) == 0) ? t : 99;
            if (trole < 2) P2_TRACE(trole, j, 0);
            mbar_wait(BAR(B2_S_FULL + t * 2 + bf), ub & 1);
            if (trole < 2) P2_TRACE(trole, j, 1);
            tc_fence_after();
            uint32_t sv[2][32];
            tmem_ld32(TM_S(t, bf) + lane_addr, sv[0]);
            tmem_ld32(TM_S(t, bf) + lane_addr + 32, sv[1]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B2_S_EMPTY + t * 2 + bf));
            if (trole < 2) P2_TRACE(trole, j, 2);
            const int key0 = j * P2_BN;
            float mx = -INFINITY;
            if (key0 + P2_BN > lim) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (key0 + c * 32 + e >= lim) sv[c][e] = 0xFF800000u;
                        mx = fmaxf(mx, __uint_as_float(sv[c][e]));
                    }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(sv[c][e]));
            }
            mx *= p.scale_log2;
            float alpha = 1.f;
            bool rescale = false;
            if (mx > m_used + 8.0f || (m_used == -INFINITY && mx != -INFINITY)) {   // lazy rescale (factor 2^8)
                alpha = m_used == -INFINITY ? 0.f : exp2f(m_used - mx);
                m_used = mx;
                rescale = j > 0;
            }
            const float neg_m = m_used == -INFINITY ? 0.f : -m_used;
            if (trole < 2) P2_TRACE(trole, j, 3);
            if (j > 1) mbar_wait(BAR(B2_PV_DONE + t * 2 + bf), (ub - 1) & 1);          // P[t][bf] free: PV_t(j-2) done
            if (trole < 2) P2_TRACE(trole, j, 4);
            l_sum *= alpha;
            if (__any_sync(0xffffffffu, rescale)) {
                mbar_wait(BAR(B2_PV_DONE + t * 2 + (bf ^ 1)), ((j - 1) >> 1) & 1);      // O_t quiescent: PV_t(j-1) done
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t ov[32];
                    tmem_ld32(TM_O(t) + lane_addr + c * 32, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                    tmem_st32(TM_O(t) + lane_addr + c * 32, ov);
                }
                tmem_wait_st();
                tc_fence_before();
            }
            const uint32_t pb = sbase + S2_P + (t * 2 + bf) * P2_PTILE + r * 128;
            float psum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c][q * 8 + 2 * e]), p.scale_log2, neg_m));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c][q * 8 + 2 * e + 1]), p.scale_log2, neg_m));
                        psum += p0 + p1;
                        w[e] = pack2_bf16(p0, p1);
                    }
                    sts128(pb + (((c * 4 + q) ^ (r & 7)) << 4), make_int4(w[0], w[1], w[2], w[3]));
                }
            l_sum += psum;
            if (trole < 2) P2_TRACE(trole, j, 5);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B2_P_FULL + t * 2 + bf));
            if (trole < 2) P2_TRACE(trole, j, 6);
        }
        // epilogue
        mbar_wait(BAR(B2_PV_DONE + t * 2 + ((n_tiles - 1) & 1)), ((n_tiles - 1) >> 1) & 1);
        tc_fence_after();
        const float inv_l = l_sum > 0.f ? 1.0f / l_sum : 0.f;
        __nv_bfloat16* orow = p.out + ((int64_t)(qs + tq) * p.n_q_head + kvh * g + hg) * P2_HD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t ov[32];
            tmem_ld32(TM_O(t) + lane_addr + c * 32, ov);
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

int prefill2_launch(const svln_kv_arena* arena, int layer, const void* q, int n_q_head, const int32_t* env_h,
                    const int32_t* q_start_h, const int32_t* n_q_h, const int32_t* kv_len_h, const int32_t* env_d,
                    const int32_t* q_start_d, const int32_t* n_q_d, const int32_t* kv_len_d, int n_seq, float scale,
                    void* out, svln_stream_t stream, const char* who) {
    const bool dev = env_d != nullptr;
    CUtensorMap map;
    int rc = get_kv_map(arena, P2_BN, &map);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        SVLN_CUDA(cudaFuncSetAttribute(attn_prefill2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM));
        attr_set = true;
    }
    const int g = n_q_head / arena->n_kv_head;
    for (int s0 = 0; s0 < n_seq; s0 += P2_MAX_SEQ) {
        Prefill2Params p;
        p.q = (const __nv_bfloat16*)q;
        p.out = (__nv_bfloat16*)out;
        p.n_q_head = n_q_head; p.n_kv_head = arena->n_kv_head; p.n_layer = arena->n_layer; p.layer = layer;
        p.max_len = arena->max_len;
        p.scale_log2 = scale * 1.4426950408889634f;
        p.trace = getenv("SVLN_PF_TRACE") ? (long long*)strtoull(getenv("SVLN_PF_TRACE"), nullptr, 0) : nullptr;
        const int ns = n_seq - s0 < P2_MAX_SEQ ? n_seq - s0 : P2_MAX_SEQ;
        p.n_seq = ns;
        p.d_env = dev ? env_d + s0 : nullptr;
        p.d_q_start = dev ? q_start_d + s0 : nullptr;
        p.d_n_q = dev ? n_q_d + s0 : nullptr;
        p.d_kv_len = dev ? kv_len_d + s0 : nullptr;
        int tiles = 0;
        for (int i = 0; i < P2_MAX_SEQ; ++i) {
            if (i < ns) {
                const int bb = s0 + i;
                SVLN_REQUIRE(n_q_h[bb] >= 1 && n_q_h[bb] <= arena->max_len, SVLN_ERR_INVALID, "%s: bad n_q=%d", who, n_q_h[bb]);
                if (!dev) {
                    SVLN_REQUIRE(env_h[bb] >= 0 && env_h[bb] < arena->n_env, SVLN_ERR_INVALID, "%s: env %d out of range", who, env_h[bb]);
                    SVLN_REQUIRE(kv_len_h[bb] >= n_q_h[bb] && kv_len_h[bb] <= arena->max_len, SVLN_ERR_INVALID,
                                 "%s: bad lengths n_q=%d kv_len=%d", who, n_q_h[bb], kv_len_h[bb]);
                }
                p.env[i] = dev ? 0 : env_h[bb]; p.q_start[i] = dev ? 0 : q_start_h[bb]; p.kv_len[i] = dev ? 0 : kv_len_h[bb];
                p.n_q[i] = n_q_h[bb];
                p.tile_start[i] = tiles;
                tiles += (n_q_h[bb] * g + P2_ROWS - 1) / P2_ROWS;
            } else {
                p.env[i] = p.q_start[i] = p.n_q[i] = p.kv_len[i] = 0;
                p.tile_start[i] = tiles;
            }
        }
        p.tile_start[P2_MAX_SEQ] = tiles;
        dim3 grid(tiles, arena->n_kv_head);
        attn_prefill2_kernel<<<grid, P2_THREADS, P2_SMEM, (cudaStream_t)stream>>>(map, p);
        rc = check_launch("attn_prefill2_kernel");
        if (rc) return rc;
    }
    return SVLN_OK;
}

}  // namespace svln

// Dense prefill attention, operands A in tensor memory (tcgen05.mma "TS" form): Q and P never touch smem.
//
// Why: the SS-form kernels of round 1 (both operands from shared memory; removed) were bound by SHARED-MEMORY bandwidth, not by the
// tensor pipe, MUFU or L2 (profiles/README.md): per 128-key tile the tensor core re-reads Q (32 KB) and P
// (32 KB) from smem besides K and V, the softmax writes P (32 KB) and TMA writes K/V (64 KB) - 224 KB per
// 1024 MMA cycles against ~128 B/clk.  Here Q is stored to TMEM once (tcgen05.st) and P overwrites the
// front of its own S buffer in TMEM as packed bf16, so smem only carries K and V (TMA in, tensor core out).
//
//   warps 0-7 : softmax; warps w and w+4 share TMEM lanes (tile rows) 32(w%4)..+31 and take 64 score columns
//               each; P(j) is written back with tcgen05.st into columns [0,64) of S buffer j&1.
//   warp 8    : MMA issue (warp-uniform control flow, elected lane): S = Q K^T  [d=S buf, a=Q tmem, b=K smem],
//               O += P V [d=O, a=P tmem, b=V smem, MN-major].  In-order tensor pipe: PV(j) precedes QK(j+2),
//               which is what makes the S/P aliasing safe.
//   warp 9    : TMA producer, 3-stage rings of K and V tiles (128 keys x 128 dims, two SWIZZLE_128B panels).
// TMEM (512 cols): S/P buf 0 [0,128), buf 1 [128,256), O [256,384), Q [384,448).
#include <stdlib.h>

#include "tcgen05_helpers.cuh"

namespace svln {

constexpr int P3_BM = 128, P3_BN = 128, P3_HD = 128;
constexpr int P3_STAGES = 3;
constexpr int P3_PANEL = 128 * 128;
constexpr int P3_TILE = 2 * P3_PANEL;           // 32 KB
constexpr int P3_SOFTMAX_WARPS = 8;
constexpr int P3_THREADS = (P3_SOFTMAX_WARPS + 2) * 32;
constexpr int P3_MAX_SEQ = 64;

constexpr int S3_K = 0;
constexpr int S3_V = S3_K + P3_STAGES * P3_TILE;
constexpr int S3_BAR = S3_V + P3_STAGES * P3_TILE;   // 192 KB
constexpr int S3_XCH = S3_BAR + 256;                 // float [2][2][128] row-max exchange
constexpr int P3_SMEM = S3_XCH + 2 * 2 * 128 * 4;

enum { B3_K_FULL = 0, B3_K_EMPTY = 3, B3_V_FULL = 6, B3_V_EMPTY = 9, B3_S_FULL = 12, B3_P_FULL = 14, B3_PV_DONE = 16,
       B3_Q_FULL = 18, B3_COUNT = 19 };

struct Prefill3Params {
    const __nv_bfloat16* q;
    __nv_bfloat16* out;
    int n_q_head, n_kv_head, n_layer, layer, max_len;
    int n_seq;
    float scale_log2;
    const int32_t *d_env, *d_q_start, *d_n_q, *d_kv_len;
    int32_t env[P3_MAX_SEQ], q_start[P3_MAX_SEQ], n_q[P3_MAX_SEQ], kv_len[P3_MAX_SEQ];
    int32_t tile_start[P3_MAX_SEQ + 1];
    // split-KV (small grids): gridDim.z CTAs share one M tile's key range; partials go to the workspace and
    // attn_prefill3_combine_kernel merges them.  n_splits == 1 writes the final output directly.
    int n_splits;
    float* part_o;   // [slot][128 rows][128 dims] fp32, unnormalised
    float* part_ml;  // [slot][128 rows][2] (reference max in log2 units, row sum)
    int debug;         // SVLN_PF_DEBUG bit 0: skip the MUFU.EX2 (bottleneck experiments only; results are wrong)
    long long* trace;  // SVLN_PF_TRACE: clock64 stamps of one CTA (scripts/prefill_trace.py), else nullptr
};

// Instrumentation is compiled out unless the library is built with -DSVLN_PF_INSTRUMENT (it costs ~8 % of the
// kernel's throughput even when disabled at run time).
// trace layout: [role 0..3][tile 0..31][slot 0..7]; roles: 0 softmax warp 0, 1 softmax warp 4, 2 MMA warp, 3 TMA warp
#ifndef SVLN_PF_INSTRUMENT
#define P3_TRACE(role, tile, slot) do { } while (0)
#define P3_NO_MUFU false
#else
#define P3_NO_MUFU ((p.debug & 1) != 0)
#define P3_TRACE(role, tile, slot)                                                                           \
    do {                                                                                                     \
        if (p.trace && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 32 && lane == 0) \
            p.trace[((role) * 32 + (tile)) * 8 + (slot)] = clock64();                                        \
    } while (0)
#endif
// every CTA's lifetime (instrumented builds): [cta][4] = smid, globaltimer at entry, after the Q store, at exit
#ifdef SVLN_PF_INSTRUMENT
#define P3_LIFE(slot_)                                                                                       \
    do {                                                                                                     \
        if (p.trace && threadIdx.x == 0) {                                                                   \
            unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                    \
            unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));                                   \
            long long* d_ = p.trace + 4 * 32 * 8 + ((int64_t)(blockIdx.y * gridDim.x + blockIdx.x)) * 4;     \
            d_[0] = sm_; d_[slot_] = (long long)t_;                                                          \
        }                                                                                                    \
    } while (0)
#else
#define P3_LIFE(slot_) do { } while (0)
#endif

// sequence that owns global M tile `gtile`: last b with tile_start[b] <= gtile.  Binary search: the linear scan was a
// chain of up to 64 dependent constant-bank loads at the very start of every CTA (~1 us of the prologue).
__device__ __forceinline__ int p3_find_seq(const Prefill3Params& p, int gtile) {
    int lo = 0, hi = p.n_seq - 1;
#pragma unroll 1
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.tile_start[mid] <= gtile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}

// CLUSTER = true: two CTAs (adjacent M tiles of the same sequence / kv head) form a cluster; each loads ONE 64-dim
// panel of every K and V tile and TMA-multicasts it into both CTAs' shared memory, so every K/V byte leaves L2
// once per 256 query rows instead of once per 128 (the 128-row kernel sits on the ~6.9 TB/s L2->SM ceiling).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

template <bool CLUSTER>
__global__ void __launch_bounds__(P3_THREADS, 1)
attn_prefill3_kernel(const __grid_constant__ CUtensorMap kv_map, const __grid_constant__ Prefill3Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = s2u(smem);
    if ((sbase & 1023u) != 0) __trap();
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + S3_BAR + B3_COUNT * 8);
    const uint32_t bar0 = sbase + S3_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    P3_LIFE(1);
    const int kvh = blockIdx.y;
    const int g = p.n_q_head / p.n_kv_head;
    const int gtile = (int)gridDim.x - 1 - (int)blockIdx.x;   // heaviest (latest rows) first
    const int b = p3_find_seq(p, gtile);
    const int mtile = gtile - p.tile_start[b];
    const bool dev = p.d_env != nullptr;
    const int nq = dev ? p.d_n_q[b] : p.n_q[b];
    const int len = dev ? p.d_kv_len[b] : p.kv_len[b];
    const int env = dev ? p.d_env[b] : p.env[b];
    const int qs = dev ? p.d_q_start[b] : p.q_start[b];
    const int rows_total = nq * g;
    const int row0 = mtile * P3_BM;
    const int rows_here = min(P3_BM, rows_total - row0);
    // CLUSTER: both CTAs of a pair walk the key range of the pair's LAST valid row (tile counts are padded to even
    // per sequence, so the partner is mtile ^ 1 of the same sequence; a padding tile has no valid rows)
    const int pair_rows_end = CLUSTER ? min(rows_total, ((mtile | 1) + 1) * P3_BM) : row0 + rows_here;
    const int tq_last = (pair_rows_end - 1) / g;
    const int kv_hi = len - nq + tq_last + 1;
    const uint32_t crank = CLUSTER ? cluster_ctarank() : 0u;
    const int n_tiles_all = (kv_hi + P3_BN - 1) / P3_BN;
    // this CTA's share of the key tiles (contiguous, balanced); may be empty for the last splits of short rows
    const int split = blockIdx.z;
    const int tile_lo = (int)(((int64_t)n_tiles_all * split) / p.n_splits);
    const int tile_hi = (int)(((int64_t)n_tiles_all * (split + 1)) / p.n_splits);
    const int n_tiles = tile_hi - tile_lo;
    const int64_t slot = ((int64_t)gtile * gridDim.y + kvh) * p.n_splits + split;
    const int plane_k = ((env * p.n_layer + p.layer) * 2 + 0) * p.n_kv_head + kvh;
    const int plane_v = plane_k + p.n_kv_head;

    // Prologue, overlapped: the Q rows are requested first (registers only), then barriers and tensor memory are set
    // up while they fly; after the one CTA barrier the TMA warp starts streaming K/V at once, the softmax warps move
    // Q into TMEM and signal Q_FULL, and only the MMA warp waits for it.  (Measured before: 4.1 us from CTA entry to
    // "Q stored" with these steps in sequence, of the 8.4 us a CTA spends outside its key loop; tcgen05.alloc itself
    // is 666 cycles.  Requesting the first K/V tiles before the CTA barrier as well changed nothing.)
    uint32_t qv[32];
    if (warp < P3_SOFTMAX_WARPS) {
        const int r = (warp & 3) * 32 + lane, half = warp >> 2;
        const int R = row0 + r;
        if (r < rows_here) {
            const int tq = R / g, hg = R - tq * g;
            const int4* src = reinterpret_cast<const int4*>(p.q + ((int64_t)(qs + tq) * p.n_q_head + kvh * g + hg) * P3_HD) + half * 8;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int4 v = __ldg(src + c);
                qv[c * 4 + 0] = v.x; qv[c * 4 + 1] = v.y; qv[c * 4 + 2] = v.z; qv[c * 4 + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) qv[c] = 0u;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < 12; ++i) mbar_init(BAR(i), 1);
        if (CLUSTER)   // a stage is free only when BOTH CTAs' tensor cores are done with it
            for (int i = 0; i < 3; ++i) { mbar_init(BAR(B3_K_EMPTY + i), 2); mbar_init(BAR(B3_V_EMPTY + i), 2); }
        mbar_init(BAR(B3_S_FULL), 1); mbar_init(BAR(B3_S_FULL + 1), 1);
        mbar_init(BAR(B3_P_FULL), P3_SOFTMAX_WARPS); mbar_init(BAR(B3_P_FULL + 1), P3_SOFTMAX_WARPS);
        mbar_init(BAR(B3_PV_DONE), 1); mbar_init(BAR(B3_PV_DONE + 1), 1);
        mbar_init(BAR(B3_Q_FULL), P3_SOFTMAX_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == P3_SOFTMAX_WARPS) {   // the MMA warp allocates: the softmax warps have loads in flight
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CLUSTER) cluster_sync_all();   // the partner's barriers are initialised before anything is multicast into them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t TM_S0 = tmem_base, TM_O = tmem_base + 256, TM_Q = tmem_base + 384;
    P3_LIFE(2);

    // Q tile -> TMEM as the A operand: lane = row, 32-bit column c = dims (2c, 2c+1).  Softmax mapping:
    // row 32*(warp%4)+lane, dims [64*half, 64*half+64) = columns [32*half, 32*half+32).
    if (warp < P3_SOFTMAX_WARPS) {
        const int half = warp >> 2;
        tmem_st32(TM_Q + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 32), qv);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B3_Q_FULL));
    }

    if (warp == P3_SOFTMAX_WARPS + 1) {
        // ------------------------------------------------------------ TMA producer (uniform warp, elected lane)
        if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&kv_map) : "memory");
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % P3_STAGES, u = j / P3_STAGES;
            const int key0 = (tile_lo + j) * P3_BN;
            P3_TRACE(3, j, 0);
            mbar_wait(BAR(B3_K_EMPTY + st), (u & 1) ^ 1);
            P3_TRACE(3, j, 1);
            if (elect_one()) {
                mbar_expect_tx(BAR(B3_K_FULL + st), P3_TILE);   // both panels land here: one from each CTA of the pair
                if (CLUSTER) {
                    tma_load_3d_mc(sbase + S3_K + st * P3_TILE + crank * P3_PANEL, &kv_map, BAR(B3_K_FULL + st), 64 * crank, key0, plane_k, 3);
                } else {
                    tma_load_3d(sbase + S3_K + st * P3_TILE, &kv_map, BAR(B3_K_FULL + st), 0, key0, plane_k);
                    tma_load_3d(sbase + S3_K + st * P3_TILE + P3_PANEL, &kv_map, BAR(B3_K_FULL + st), 64, key0, plane_k);
                }
            }
            mbar_wait(BAR(B3_V_EMPTY + st), (u & 1) ^ 1);
            P3_TRACE(3, j, 2);
            if (elect_one()) {
                mbar_expect_tx(BAR(B3_V_FULL + st), P3_TILE);
                if (CLUSTER) {
                    tma_load_3d_mc(sbase + S3_V + st * P3_TILE + crank * P3_PANEL, &kv_map, BAR(B3_V_FULL + st), 64 * crank, key0, plane_v, 3);
                } else {
                    tma_load_3d(sbase + S3_V + st * P3_TILE, &kv_map, BAR(B3_V_FULL + st), 0, key0, plane_v);
                    tma_load_3d(sbase + S3_V + st * P3_TILE + P3_PANEL, &kv_map, BAR(B3_V_FULL + st), 64, key0, plane_v);
                }
            }
            __syncwarp();
        }
    } else if (warp == P3_SOFTMAX_WARPS) {
        // ------------------------------------------------------------ MMA issuer (uniform warp, elected lane)
        constexpr uint32_t IDESC_QK = umma_idesc(0, 128);
        constexpr uint32_t IDESC_PV = umma_idesc(1, 128);
        auto issue_qk = [&](int j) {
            const int st = j % P3_STAGES, bf = j & 1;
            P3_TRACE(2, j, 0);
            mbar_wait(BAR(B3_K_FULL + st), (j / P3_STAGES) & 1);
            P3_TRACE(2, j, 1);
            tc_fence_after();
            const uint32_t kb = sbase + S3_K + st * P3_TILE;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)   // 16 dims per step = 8 TMEM columns of Q; K: 4 steps per 64-dim panel
                    tc_mma_ts(TM_S0 + bf * 128, TM_Q + ks * 8, umma_desc(kb + (ks >> 2) * P3_PANEL + (ks & 3) * 32, 16, 1024),
                              IDESC_QK, ks > 0);
                tc_commit(BAR(B3_S_FULL + bf));
                if (CLUSTER) tc_commit_mc(BAR(B3_K_EMPTY + st), 3); else tc_commit(BAR(B3_K_EMPTY + st));
            }
            __syncwarp();
            P3_TRACE(2, j, 2);
        };
        // S buffer bf is rewritten by QK(j+2) only after PV(j) (which read P(j) from it) was ISSUED: the tensor
        // pipe runs in order, so no extra barrier is needed; PV(j) itself waits for P_FULL(j), i.e. for every
        // softmax thread to be done with S(j).
        mbar_wait(BAR(B3_Q_FULL), 0);   // Q is in tensor memory
        tc_fence_after();
        if (n_tiles > 0) issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) issue_qk(j + 1);
            const int st = j % P3_STAGES, bf = j & 1, ub = j >> 1;
            mbar_wait(BAR(B3_P_FULL + bf), ub & 1);
            P3_TRACE(2, j, 3);
            mbar_wait(BAR(B3_V_FULL + st), (j / P3_STAGES) & 1);
            P3_TRACE(2, j, 4);
            tc_fence_after();
            const uint32_t vb = sbase + S3_V + st * P3_TILE;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)   // 16 keys per step = 8 TMEM columns of P; V: 16 rows x 128 B per step
                    tc_mma_ts(TM_O, TM_S0 + bf * 128 + ks * 8, umma_desc(vb + ks * 2048, P3_PANEL, 1024), IDESC_PV,
                              (j > 0 || ks > 0));
                if (CLUSTER) tc_commit_mc(BAR(B3_V_EMPTY + st), 3); else tc_commit(BAR(B3_V_EMPTY + st));
                tc_commit(BAR(B3_PV_DONE + bf));
            }
            __syncwarp();
            P3_TRACE(2, j, 5);
        }
    } else {
        // ------------------------------------------------------------ softmax warps
        const int r = (warp & 3) * 32 + lane;
        const int half = warp >> 2;
        const int R = row0 + r;
        const bool row_valid = r < rows_here;
        const int tq = row_valid ? R / g : 0;
        const int hg = row_valid ? R - tq * g : 0;
        const int lim = row_valid ? len - nq + tq + 1 : 0;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t xch = sbase + S3_XCH;
        const int pair_bar = 1 + (warp & 3);
        float m_used = -INFINITY, l_sum = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int bf = j & 1, ub = j >> 1;
            const int trole = (warp & 3) == 0 ? half : 99;
            (void)trole;
            if (trole < 2) P3_TRACE(trole, j, 0);
            mbar_wait(BAR(B3_S_FULL + bf), ub & 1);
            if (trole < 2) P3_TRACE(trole, j, 1);
            tc_fence_after();
#ifdef SVLN_PF_INSTRUMENT
            if (p.debug & 2) {   // experiment: no softmax at all (wrong results) - the pure MMA + TMA pipeline rate
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B3_P_FULL + bf));
                continue;
            }
#endif
            uint32_t sv[2][32];
            const uint32_t sbuf = TM_S0 + bf * 128 + lane_base;
            tmem_ld32(sbuf + half * 64, sv[0]);
            tmem_ld32(sbuf + half * 64 + 32, sv[1]);
            tmem_wait_ld();
            const int key0 = (tile_lo + j) * P3_BN + half * 64;
            float mx = -INFINITY;
            if (key0 + 64 > lim) {
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
            if (trole < 2) P3_TRACE(trole, j, 2);
            sts32f(xch + ((bf * 2 + half) * 128 + r) * 4, mx);
            // pair barrier: (a) row max exchange, (b) both halves have their S(j) in registers, so P(j) may now
            // overwrite columns [0,64) of this S buffer
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            mx = fmaxf(mx, lds32f(xch + ((bf * 2 + (half ^ 1)) * 128 + r) * 4)) * p.scale_log2;
            float alpha = 1.f;
            bool rescale = false;
            if (mx > m_used + 8.0f || (m_used == -INFINITY && mx != -INFINITY)) {   // lazy rescale (factor 2^8)
                alpha = m_used == -INFINITY ? 0.f : exp2f(m_used - mx);
                m_used = mx;
                rescale = j > 0;
            }
            const float neg_m = m_used == -INFINITY ? 0.f : -m_used;
            l_sum *= alpha;
            if (trole < 2) P3_TRACE(trole, j, 3);
            if (__any_sync(0xffffffffu, rescale)) {
                mbar_wait(BAR(B3_PV_DONE + (bf ^ 1)), ((j - 1) >> 1) & 1);   // O quiescent: PV(j-1) done
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t ov[32];
                    tmem_ld32(TM_O + lane_base + half * 64 + c * 32, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                    tmem_st32(TM_O + lane_base + half * 64 + c * 32, ov);
                }
            }
            // P = exp2(S*scale - m) -> packed bf16x2 -> TMEM columns [32*half, 32*half+32) of this S buffer
            uint32_t pv[32];
            float psum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0 = fmaf(__uint_as_float(sv[c][2 * e]), p.scale_log2, neg_m);
                    float p1 = fmaf(__uint_as_float(sv[c][2 * e + 1]), p.scale_log2, neg_m);
                    if (!P3_NO_MUFU) { p0 = ex2_approx(p0); p1 = ex2_approx(p1); }
                    psum += p0 + p1;
                    pv[c * 16 + e] = pack2_bf16(p0, p1);
                }
            l_sum += psum;
            if (trole < 2) P3_TRACE(trole, j, 4);
            tmem_st32(TM_S0 + bf * 128 + lane_base + half * 32, pv);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B3_P_FULL + bf));
            if (trole < 2) P3_TRACE(trole, j, 5);
        }
        // epilogue
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        sts32f(xch + (half * 128 + r) * 4, l_sum);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float l_tot = l_sum + lds32f(xch + ((half ^ 1) * 128 + r) * 4);
        if (n_tiles > 0) {
            mbar_wait(BAR(B3_PV_DONE + ((n_tiles - 1) & 1)), ((n_tiles - 1) >> 1) & 1);
            tc_fence_after();
        }
        if (p.n_splits == 1) {
            const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
            __nv_bfloat16* orow = p.out + ((int64_t)(qs + tq) * p.n_q_head + kvh * g + hg) * P3_HD + half * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t ov[32];
                tmem_ld32(TM_O + lane_base + half * 64 + c * 32, ov);
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
        } else {
            // partial: unnormalised O, reference max and row sum of this key range (empty range: m = -inf, l = 0)
            float* po = p.part_o + (slot * P3_BM + r) * P3_HD + half * 64;
            if (half == 0) {
                p.part_ml[(slot * P3_BM + r) * 2 + 0] = n_tiles > 0 ? m_used : -INFINITY;
                p.part_ml[(slot * P3_BM + r) * 2 + 1] = n_tiles > 0 ? l_tot : 0.f;
            }
            if (n_tiles > 0) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t ov[32];
                    tmem_ld32(TM_O + lane_base + half * 64 + c * 32, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<int4*>(po + c * 32 + q * 4) = make_int4(ov[q * 4], ov[q * 4 + 1], ov[q * 4 + 2], ov[q * 4 + 3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    P3_LIFE(3);
    if (CLUSTER) cluster_sync_all();   // neither CTA may exit while the partner can still multicast / commit into it
    if (warp == P3_SOFTMAX_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- persistent variant ---------------------------------------------------------------------------------------
// The kernel above costs 0.97 us per key tile plus a fixed ~6.3 us per CTA (launch gap, setup, pipeline fill, O
// read-out): a fifth of the steady-turn shape (24 key tiles per CTA) and half of the window-start shape (7.5 on
// average).  Here one CTA per SM walks a list of (M tile, kv head) items and the tensor pipe never drains between
// them: the TMA warp streams the next item's K/V tiles as ring slots free up; Q is double-buffered in tensor memory
// (the 64 spare columns), stored one item AHEAD by the softmax warps, so the MMA warp issues the next item's first
// QK right after the current item's last one - before the current item's last softmax, PV and O read-out - and
// keeps one global tile counter, so the S/P double buffer and the ring phases simply continue.  What orders the
// items: Q_FULL[2] (count = softmax warps); O of item i is read out by the softmax warps before they produce
// P(0) of item i+1, which PV(0) of item i+1 (the first write to O) waits for.
// MEASURED: not a win - 845 vs 964 TFLOP/s for the plain kernel at 64 x 256 over 3072 keys (1.09 us per tile and
// still 7.5 us per item), so it is opt-in (SVLN_PF_PERSIST) and kept for the next round's work on the item
// boundary.  (A first persistent version without the Q double buffer cost 7.6 us per item; a 16-softmax-warp
// version, 32 scores per thread, did not change the 0.97 us per tile: with the softmax removed the MMA + TMA
// pipeline alone needs 0.79 us per tile.)
constexpr int B3P_Q_FULL = B3_COUNT;          // [2], count = softmax warps, one phase per use
constexpr int B3P_COUNT = B3_COUNT + 2;
constexpr int S3P_XCH2 = P3_SMEM;             // float [2][128] row-sum exchange of the read-out (separate from the row-max one)
constexpr int P3P_SMEM = S3P_XCH2 + 2 * 128 * 4;

struct P3Item {
    int kvh, nq, len, qs, row0, rows_here, n_tiles, plane_k, plane_v;
};
__device__ __forceinline__ P3Item p3_item(const Prefill3Params& p, int it, int total_tiles) {
    P3Item I;
    const int gt = total_tiles - 1 - it / p.n_kv_head;     // heaviest (latest rows) first
    I.kvh = it % p.n_kv_head;
    const int b = p3_find_seq(p, gt);
    const int mtile = gt - p.tile_start[b];
    const bool dev = p.d_env != nullptr;
    I.nq = dev ? p.d_n_q[b] : p.n_q[b];
    I.len = dev ? p.d_kv_len[b] : p.kv_len[b];
    const int env = dev ? p.d_env[b] : p.env[b];
    I.qs = dev ? p.d_q_start[b] : p.q_start[b];
    const int g = p.n_q_head / p.n_kv_head;
    const int rows_total = I.nq * g;
    I.row0 = mtile * P3_BM;
    I.rows_here = min(P3_BM, rows_total - I.row0);
    const int tq_last = (I.row0 + I.rows_here - 1) / g;
    const int kv_hi = I.len - I.nq + tq_last + 1;
    I.n_tiles = max(0, (kv_hi + P3_BN - 1) / P3_BN);
    I.plane_k = ((env * p.n_layer + p.layer) * 2 + 0) * p.n_kv_head + I.kvh;
    I.plane_v = I.plane_k + p.n_kv_head;
    return I;
}

__global__ void __launch_bounds__(P3_THREADS, 1)
attn_prefill3p_kernel(const __grid_constant__ CUtensorMap kv_map, const __grid_constant__ Prefill3Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = s2u(smem);
    if ((sbase & 1023u) != 0) __trap();
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + S3_BAR + B3P_COUNT * 8);
    const uint32_t bar0 = sbase + S3_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = p.n_q_head / p.n_kv_head;
    const int total_tiles = p.tile_start[P3_MAX_SEQ];
    const int n_items = total_tiles * p.n_kv_head;
    const int stride = gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < 12; ++i) mbar_init(BAR(i), 1);
        mbar_init(BAR(B3_S_FULL), 1); mbar_init(BAR(B3_S_FULL + 1), 1);
        mbar_init(BAR(B3_P_FULL), P3_SOFTMAX_WARPS); mbar_init(BAR(B3_P_FULL + 1), P3_SOFTMAX_WARPS);
        mbar_init(BAR(B3_PV_DONE), 1); mbar_init(BAR(B3_PV_DONE + 1), 1);
        mbar_init(BAR(B3P_Q_FULL), P3_SOFTMAX_WARPS); mbar_init(BAR(B3P_Q_FULL + 1), P3_SOFTMAX_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == P3_SOFTMAX_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s2u(tmem_ptr_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t TM_S0 = tmem_base, TM_O = tmem_base + 256, TM_Q = tmem_base + 384;   // Q buffers at +384 and +448

    if (warp == P3_SOFTMAX_WARPS + 1) {
        // ------------------------------------------------------------ TMA producer: runs ahead across items
        if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&kv_map) : "memory");
        uint32_t T = 0;
        for (int it = blockIdx.x; it < n_items; it += stride) {
            const P3Item I = p3_item(p, it, total_tiles);
            for (int j = 0; j < I.n_tiles; ++j, ++T) {
                const int st = T % P3_STAGES;
                const uint32_t u = T / P3_STAGES;
                const int key0 = j * P3_BN;
                mbar_wait(BAR(B3_K_EMPTY + st), (u & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(BAR(B3_K_FULL + st), P3_TILE);
                    tma_load_3d(sbase + S3_K + st * P3_TILE, &kv_map, BAR(B3_K_FULL + st), 0, key0, I.plane_k);
                    tma_load_3d(sbase + S3_K + st * P3_TILE + P3_PANEL, &kv_map, BAR(B3_K_FULL + st), 64, key0, I.plane_k);
                }
                mbar_wait(BAR(B3_V_EMPTY + st), (u & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(BAR(B3_V_FULL + st), P3_TILE);
                    tma_load_3d(sbase + S3_V + st * P3_TILE, &kv_map, BAR(B3_V_FULL + st), 0, key0, I.plane_v);
                    tma_load_3d(sbase + S3_V + st * P3_TILE + P3_PANEL, &kv_map, BAR(B3_V_FULL + st), 64, key0, I.plane_v);
                }
                __syncwarp();
            }
        }
    } else if (warp == P3_SOFTMAX_WARPS) {
        // ------------------------------------------------------------ MMA issuer (uniform warp, elected lane)
        constexpr uint32_t IDESC_QK = umma_idesc(0, 128);
        constexpr uint32_t IDESC_PV = umma_idesc(1, 128);
        auto issue_qk = [&](uint32_t T, uint32_t qbuf) {
            const int st = T % P3_STAGES, bf = T & 1;
            mbar_wait(BAR(B3_K_FULL + st), (T / P3_STAGES) & 1);
            tc_fence_after();
            const uint32_t kb = sbase + S3_K + st * P3_TILE;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tc_mma_ts(TM_S0 + bf * 128, TM_Q + qbuf * 64 + ks * 8,
                              umma_desc(kb + (ks >> 2) * P3_PANEL + (ks & 3) * 32, 16, 1024), IDESC_QK, ks > 0);
                tc_commit(BAR(B3_S_FULL + bf));
                tc_commit(BAR(B3_K_EMPTY + st));
            }
            __syncwarp();
        };
        // first QK of item `idx` (its tile counter starts at T): its Q must be in tensor memory
        auto first_qk = [&](uint32_t idx, uint32_t T) {
            mbar_wait(BAR(B3P_Q_FULL + (idx & 1)), (idx >> 1) & 1);
            tc_fence_after();
            issue_qk(T, idx & 1);
        };
        uint32_t T0 = 0, idx = 0;
        bool first_issued = false;
        for (int it = blockIdx.x; it < n_items; it += stride, ++idx) {
            const P3Item I = p3_item(p, it, total_tiles);
            const int it_next = it + stride;
            int next_tiles = 0;
            if (it_next < n_items) next_tiles = p3_item(p, it_next, total_tiles).n_tiles;
            if (I.n_tiles > 0 && !first_issued) first_qk(idx, T0);
            first_issued = false;
            for (int j = 0; j < I.n_tiles; ++j) {
                const uint32_t T = T0 + j;
                if (j + 1 < I.n_tiles) {
                    issue_qk(T + 1, idx & 1);
                } else if (next_tiles > 0) {      // keep the tensor pipe fed across the item boundary
                    first_qk(idx + 1, T + 1);
                    first_issued = true;
                }
                const int st = T % P3_STAGES, bf = T & 1;
                mbar_wait(BAR(B3_P_FULL + bf), (T >> 1) & 1);
                mbar_wait(BAR(B3_V_FULL + st), (T / P3_STAGES) & 1);
                tc_fence_after();
                const uint32_t vb = sbase + S3_V + st * P3_TILE;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        tc_mma_ts(TM_O, TM_S0 + bf * 128 + ks * 8, umma_desc(vb + ks * 2048, P3_PANEL, 1024), IDESC_PV,
                                  (j > 0 || ks > 0));
                    tc_commit(BAR(B3_V_EMPTY + st));
                    tc_commit(BAR(B3_PV_DONE + bf));
                }
                __syncwarp();
            }
            T0 += I.n_tiles;
        }
    } else {
        // ------------------------------------------------------------ softmax warps
        const int r = (warp & 3) * 32 + lane;
        const int half = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t xch = sbase + S3_XCH, xch2 = sbase + S3P_XCH2;
        const int pair_bar = 1 + (warp & 3);
        auto load_q = [&](const P3Item& I, uint32_t (&qv)[32]) {
            const int R = I.row0 + r;
            if (r < I.rows_here) {
                const int tq = R / g, hg = R - tq * g;
                const int4* src = reinterpret_cast<const int4*>(p.q + ((int64_t)(I.qs + tq) * p.n_q_head + I.kvh * g + hg) * P3_HD) + half * 8;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int4 v = __ldg(src + c);
                    qv[c * 4 + 0] = v.x; qv[c * 4 + 1] = v.y; qv[c * 4 + 2] = v.z; qv[c * 4 + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) qv[c] = 0u;
            }
        };
        auto store_q = [&](uint32_t idx, const uint32_t (&qv)[32]) {
            tmem_st32(TM_Q + (idx & 1) * 64 + lane_base + (uint32_t)(half * 32), qv);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B3P_Q_FULL + (idx & 1)));
        };
        P3Item I;
        {
            uint32_t qv[32];
            if ((int)blockIdx.x < n_items) {
                I = p3_item(p, blockIdx.x, total_tiles);
                load_q(I, qv);
                store_q(0, qv);
            }
        }
        uint32_t T0 = 0, idx = 0;
        for (int it = blockIdx.x; it < n_items; it += stride, ++idx) {
            // the NEXT item's Q rows are requested now and stored into the other Q buffer after this item's first
            // tile: every QK of item idx-1 (the buffer's previous user) completed before this item started
            const int it_next = it + stride;
            const bool has_next = it_next < n_items;
            P3Item In = I;
            uint32_t qn[32];
            if (has_next) {
                In = p3_item(p, it_next, total_tiles);
                load_q(In, qn);
            }
            bool q_stored = !has_next;
            const int R = I.row0 + r;
            const bool row_valid = r < I.rows_here;
            const int tq = row_valid ? R / g : 0;
            const int hg = row_valid ? R - tq * g : 0;
            const int lim = row_valid ? I.len - I.nq + tq + 1 : 0;
            const int n_tiles = I.n_tiles;
            float m_used = -INFINITY, l_sum = 0.f;
            for (int j = 0; j < n_tiles; ++j) {
                const uint32_t T = T0 + j;
                const int bf = T & 1;
                mbar_wait(BAR(B3_S_FULL + bf), (T >> 1) & 1);
                tc_fence_after();
                uint32_t sv[2][32];
                const uint32_t sbuf = TM_S0 + bf * 128 + lane_base;
                tmem_ld32(sbuf + half * 64, sv[0]);
                tmem_ld32(sbuf + half * 64 + 32, sv[1]);
                tmem_wait_ld();
                const int key0 = j * P3_BN + half * 64;
                float mx = -INFINITY;
                if (key0 + 64 > lim) {
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
                sts32f(xch + ((bf * 2 + half) * 128 + r) * 4, mx);
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                mx = fmaxf(mx, lds32f(xch + ((bf * 2 + (half ^ 1)) * 128 + r) * 4)) * p.scale_log2;
                float alpha = 1.f;
                bool rescale = false;
                if (mx > m_used + 8.0f || (m_used == -INFINITY && mx != -INFINITY)) {   // lazy rescale (factor 2^8)
                    alpha = m_used == -INFINITY ? 0.f : exp2f(m_used - mx);
                    m_used = mx;
                    rescale = j > 0;
                }
                const float neg_m = m_used == -INFINITY ? 0.f : -m_used;
                l_sum *= alpha;
                if (__any_sync(0xffffffffu, rescale)) {
                    mbar_wait(BAR(B3_PV_DONE + (bf ^ 1)), ((T - 1) >> 1) & 1);   // O quiescent: PV(T-1) done
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t ov[32];
                        tmem_ld32(TM_O + lane_base + half * 64 + c * 32, ov);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
                        tmem_st32(TM_O + lane_base + half * 64 + c * 32, ov);
                    }
                }
                uint32_t pv[32];
                float psum = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c][2 * e]), p.scale_log2, neg_m));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c][2 * e + 1]), p.scale_log2, neg_m));
                        psum += p0 + p1;
                        pv[c * 16 + e] = pack2_bf16(p0, p1);
                    }
                l_sum += psum;
                tmem_st32(TM_S0 + bf * 128 + lane_base + half * 32, pv);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B3_P_FULL + bf));
                if (!q_stored) {   // after the first tile: off the critical path of this item's pipeline fill
                    store_q(idx + 1, qn);
                    q_stored = true;
                }
            }
            if (!q_stored) store_q(idx + 1, qn);   // an item without tiles
            // read-out of this item
            sts32f(xch2 + (half * 128 + r) * 4, l_sum);
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            const float l_tot = l_sum + lds32f(xch2 + ((half ^ 1) * 128 + r) * 4);
            if (n_tiles > 0) {
                const uint32_t Tl = T0 + n_tiles - 1;
                mbar_wait(BAR(B3_PV_DONE + (Tl & 1)), (Tl >> 1) & 1);
                tc_fence_after();
            }
            const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
            __nv_bfloat16* orow = p.out + ((int64_t)(I.qs + tq) * p.n_q_head + I.kvh * g + hg) * P3_HD + half * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t ov[32];
                tmem_ld32(TM_O + lane_base + half * 64 + c * 32, ov);
                tmem_wait_ld();
                if (row_valid && n_tiles > 0) {
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
            // the row-sum slots are reused by the next item's read-out only after both halves have read them
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            T0 += n_tiles;
            I = In;
        }
    }
    __syncthreads();
    if (warp == P3_SOFTMAX_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// merge the split-KV partials of one (M tile, kv head): one warp per row, lanes over dims (4 floats each).  The grid's
// z dimension cuts the tile into 32-row slabs and every warp issues the loads of its 4 rows x all splits before it uses
// any of them: the first version walked 16 rows per warp one dependent L2 round trip after the other and took ~30 us,
// which is why split-KV measured SLOWER than the 48-CTA grid in round 1 (53 vs 40 us per layer).
constexpr int P3C_ROWS = 32, P3C_MAX_SPLITS = 4;
__global__ void __launch_bounds__(256)
attn_prefill3_combine_kernel(const __grid_constant__ Prefill3Params p) {
    const int gtile = blockIdx.x, kvh = blockIdx.y;
    const int g = p.n_q_head / p.n_kv_head;
    const int b = p3_find_seq(p, gtile);
    const int mtile = gtile - p.tile_start[b];
    const bool dev = p.d_env != nullptr;
    const int nq = dev ? p.d_n_q[b] : p.n_q[b];
    const int qs = dev ? p.d_q_start[b] : p.q_start[b];
    const int rows_total = nq * g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t slot0 = ((int64_t)gtile * gridDim.y + kvh) * p.n_splits;
    constexpr int RPW = P3C_ROWS / 8;
    float ms[RPW][P3C_MAX_SPLITS], ls[RPW][P3C_MAX_SPLITS];
    float4 o[RPW][P3C_MAX_SPLITS];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int r = blockIdx.z * P3C_ROWS + i * 8 + warp;
#pragma unroll
        for (int s_ = 0; s_ < P3C_MAX_SPLITS; ++s_) {
            const bool ok = s_ < p.n_splits && mtile * P3_BM + r < rows_total;
            const int64_t at = (slot0 + (ok ? s_ : 0)) * P3_BM + r;
            ms[i][s_] = ok ? p.part_ml[at * 2] : -INFINITY;
            ls[i][s_] = ok ? p.part_ml[at * 2 + 1] : 0.f;
            o[i][s_] = ok ? *reinterpret_cast<const float4*>(p.part_o + at * P3_HD + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int r = blockIdx.z * P3C_ROWS + i * 8 + warp;
        const int R = mtile * P3_BM + r;
        if (R >= rows_total) continue;
        float m = -INFINITY;
#pragma unroll
        for (int s_ = 0; s_ < P3C_MAX_SPLITS; ++s_) m = fmaxf(m, ms[i][s_]);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float l = 0.f;
#pragma unroll
        for (int s_ = 0; s_ < P3C_MAX_SPLITS; ++s_) {        // fixed split order: deterministic
            if (ms[i][s_] == -INFINITY) continue;
            const float w = exp2f(ms[i][s_] - m);
            l += ls[i][s_] * w;
            acc.x += o[i][s_].x * w; acc.y += o[i][s_].y * w; acc.z += o[i][s_].z * w; acc.w += o[i][s_].w * w;
        }
        const float inv_l = l > 0.f ? 1.0f / l : 0.f;
        const int tq = R / g, hg = R - tq * g;
        uint2 o2;
        o2.x = pack2_bf16(acc.x * inv_l, acc.y * inv_l);
        o2.y = pack2_bf16(acc.z * inv_l, acc.w * inv_l);
        *reinterpret_cast<uint2*>(p.out + ((int64_t)(qs + tq) * p.n_q_head + kvh * g + hg) * P3_HD + lane * 4) = o2;
    }
}

size_t prefill3_workspace_bytes(int64_t n_slots) { return (size_t)n_slots * P3_BM * (P3_HD + 2) * sizeof(float); }

int prefill3_launch(const svln_kv_arena* arena, int layer, const void* q, int n_q_head, const int32_t* env_h,
                    const int32_t* q_start_h, const int32_t* n_q_h, const int32_t* kv_len_h, const int32_t* env_d,
                    const int32_t* q_start_d, const int32_t* n_q_d, const int32_t* kv_len_d, int n_seq, float scale,
                    void* out, svln_stream_t stream, const char* who, void* workspace, size_t workspace_bytes) {
    const bool dev = env_d != nullptr;
    CUtensorMap map;
    int rc = get_kv_map(arena, P3_BN, &map);
    if (rc) return rc;
    static unsigned long long attr_set = 0;
    if (first_use_on_device(&attr_set)) {
        SVLN_CUDA(cudaFuncSetAttribute(attn_prefill3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM));
        SVLN_CUDA(cudaFuncSetAttribute(attn_prefill3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM));
        SVLN_CUDA(cudaFuncSetAttribute(attn_prefill3p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P3P_SMEM));
    }
    // measured on B200: the multicast pair is ~5 % SLOWER than the plain kernel at the 64-stream shapes (L2 already
    // serves the repeated K/V tiles; the pair adds cluster barriers and lock-step) - opt-in via SVLN_PF_CLUSTER=1
    static const int cluster_mode = getenv("SVLN_PF_CLUSTER") ? atoi(getenv("SVLN_PF_CLUSTER")) : 0;
    const int g = n_q_head / arena->n_kv_head;
    for (int s0 = 0; s0 < n_seq; s0 += P3_MAX_SEQ) {
        Prefill3Params p;
        p.q = (const __nv_bfloat16*)q;
        p.out = (__nv_bfloat16*)out;
        p.n_q_head = n_q_head; p.n_kv_head = arena->n_kv_head; p.n_layer = arena->n_layer; p.layer = layer;
        p.max_len = arena->max_len;
        p.scale_log2 = scale * 1.4426950408889634f;
        static const int dbg_env = getenv("SVLN_PF_DEBUG") ? atoi(getenv("SVLN_PF_DEBUG")) : 0;
        p.debug = dbg_env;
#ifdef SVLN_PF_INSTRUMENT   // build with SVLN_NVCC_EXTRA=-DSVLN_PF_INSTRUMENT: a device pointer in the environment is never
                            // dereferenced by a production build
        static long long* const trace_env = getenv("SVLN_PF_TRACE") ? (long long*)strtoull(getenv("SVLN_PF_TRACE"), nullptr, 0) : nullptr;
        p.trace = trace_env;
#else
        p.trace = nullptr;
#endif
        const int ns = n_seq - s0 < P3_MAX_SEQ ? n_seq - s0 : P3_MAX_SEQ;
        p.n_seq = ns;
        p.d_env = dev ? env_d + s0 : nullptr;
        p.d_q_start = dev ? q_start_d + s0 : nullptr;
        p.d_n_q = dev ? n_q_d + s0 : nullptr;
        p.d_kv_len = dev ? kv_len_d + s0 : nullptr;
        // multicast pairs pay off once there are enough M tiles to fill the GPU; tiny launches keep the plain kernel
        int raw_tiles = 0;
        for (int i = 0; i < ns; ++i) raw_tiles += (n_q_h[s0 + i] * g + P3_BM - 1) / P3_BM;
        const bool use_cluster = cluster_mode != 0 && raw_tiles * arena->n_kv_head >= sm_count();
        int tiles = 0;
        for (int i = 0; i < P3_MAX_SEQ; ++i) {
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
                int nt = (n_q_h[bb] * g + P3_BM - 1) / P3_BM;
                if (use_cluster) nt = (nt + 1) & ~1;   // pairs never straddle sequences
                tiles += nt;
            } else {
                p.env[i] = p.q_start[i] = p.n_q[i] = p.kv_len[i] = 0;
                p.tile_start[i] = tiles;
            }
        }
        p.tile_start[P3_MAX_SEQ] = tiles;
        // split the key range when the grid cannot fill the GPU (single-episode turns: 12 M tiles x 4 kv heads)
        const int ctas = tiles * arena->n_kv_head;
        int n_splits = 1;
        // measured on B200 (single-episode steady turn, 217 q x 7 heads over 3000 keys = 48 CTAs, L2 flushed): 1 / 2 / 3 / 4
        // splits = 40.0 / 33.8 / 29.6 / 29.7 us per layer (round 2, after the combine kernel stopped serialising its L2
        // round trips); SVLN_PF_SPLITS=1 turns it off
        static const int max_splits = getenv("SVLN_PF_SPLITS") ? atoi(getenv("SVLN_PF_SPLITS")) : 3;
        if (!use_cluster && workspace && ctas * 2 <= sm_count() && max_splits > 1) {
            n_splits = sm_count() / ctas;   // one wave: never more CTAs than SMs (each CTA owns a whole SM)
            if (n_splits > max_splits) n_splits = max_splits;
            if (n_splits > P3C_MAX_SPLITS) n_splits = P3C_MAX_SPLITS;
            while (n_splits > 1 && prefill3_workspace_bytes((int64_t)ctas * n_splits) > workspace_bytes) --n_splits;
        }
        p.n_splits = n_splits;
        p.part_o = (float*)workspace;
        p.part_ml = p.part_o + (size_t)ctas * n_splits * P3_BM * P3_HD;
        dim3 grid(tiles, arena->n_kv_head, n_splits);
        // persistent variant: pays off as soon as some SM would run more than one CTA of the plain kernel
        // (SVLN_PF_PERSIST: 0 never, 1 when the grid exceeds the SM count, 2 always - tests)
        static const int persist = getenv("SVLN_PF_PERSIST") ? atoi(getenv("SVLN_PF_PERSIST")) : 0;
        if (persist && !use_cluster && n_splits == 1 && (ctas > sm_count() || persist == 2)) {
            const int pgrid = ctas < sm_count() ? ctas : sm_count();
            attn_prefill3p_kernel<<<pgrid, P3_THREADS, P3P_SMEM, (cudaStream_t)stream>>>(map, p);
        } else if (use_cluster) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid;
            cfg.blockDim = dim3(P3_THREADS);
            cfg.dynamicSmemBytes = P3_SMEM;
            cfg.stream = (cudaStream_t)stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            SVLN_CUDA(cudaLaunchKernelEx(&cfg, attn_prefill3_kernel<true>, map, p));
        } else {
            attn_prefill3_kernel<false><<<grid, P3_THREADS, P3_SMEM, (cudaStream_t)stream>>>(map, p);
        }
        if (n_splits > 1) {
            dim3 cgrid(tiles, arena->n_kv_head, P3_BM / P3C_ROWS);
            attn_prefill3_combine_kernel<<<cgrid, 256, 0, (cudaStream_t)stream>>>(p);
        }
        rc = check_launch("attn_prefill3_kernel");
        if (rc) return rc;
    }
    return SVLN_OK;
}

}  // namespace svln

// Split-KV decode attention work item shared by attn_decode_kernel (attn_decode.cu) and the fused decode
// stack (decode_stack.cu).  See attn_decode.cu for the design notes.
#pragma once
#include "svln_common.cuh"

namespace svln {

constexpr int DT = 64;        // keys per smem tile
constexpr int DSTAGES = 3;    // cp.async ring depth
constexpr int DWARPS = 4;     // warps per CTA; each owns 16 keys of a tile
constexpr int DROWS = 16;     // rows of the MMA tile
constexpr int HD = 128;       // head_dim
constexpr int TILE_BYTES = DT * HD * 2;  // 16 KB
constexpr int DEC_SMEM = DSTAGES * 2 * TILE_BYTES;  // 96 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    return f32_to_bf16_bits(lo) | (f32_to_bf16_bits(hi) << 16);
}

struct DecodeParams {
    svln_kv_arena arena;
    int layer;
    const __nv_bfloat16* q;   // [n_tok, Hq, 128]
    int n_q_head;
    const int32_t* seq_env;
    const int32_t* q_start;
    const int32_t* n_q;
    const int32_t* kv_len;
    int groups_per_seq;       // ceil(max_n_q / tok_per_group)
    int tok_per_group;        // 16 / g
    int split_keys;           // keys per split (multiple of DT)
    int n_splits;
    float scale_log2;         // softmax scale * log2(e)
    float* part_o;            // [slots][16][128]
    float* part_m;            // [slots][16]  (raw score units)
    float* part_l;            // [slots][16]
    __nv_bfloat16* out;       // [n_tok, Hq*128]
    long long* trace;         // debug: clock64 stamps of CTA 0 / thread 0 (8 slots), else nullptr
};
#define AD_TRACE(slot_) do { if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[slot_] = clock64(); } while (0)

// swizzled byte offset of 16-byte chunk `chunk` (0..15) of row `row` inside a [DT][128] bf16 tile
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return (uint32_t)(row * 256 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void load_tile_async(uint32_t smem_tile, const __nv_bfloat16* plane, int key0,
                                                int key_end, int tid) {
    // 64 rows x 16 chunks = 1024 x 16 B, 128 threads -> 8 each; rows past key_end are zero-filled by
    // clamping to a valid row (their scores are masked anyway)
#pragma unroll
    for (int i = 0; i < (DT * 16) / (DWARPS * 32); ++i) {
        const int e = i * (DWARPS * 32) + tid;
        const int row = e >> 4, chunk = e & 15;
        int key = key0 + row;
        key = key < key_end ? key : key_end - 1;
        cp_async16(smem_tile + tile_off(row, chunk), plane + (int64_t)key * HD + chunk * 8);
    }
}

template <int BAR_ID>
__device__ __forceinline__ void group_sync() {
    if (BAR_ID == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(DWARPS * 32) : "memory");
}
// (q may have been written earlier in the same kernel by another CTA - fused decode stack - so it is read
// through L2 with ld.global.cg (volatile asm), never through the non-coherent path)

// One (split, kv head, row group) work item, executed by DWARPS*32 threads that share barrier BAR_ID
// (0 = the whole CTA, i.e. __syncthreads; other ids let a 4-warp group of a larger CTA run it).
// MODE 0: the whole item.  MODE 1: only issue the first DSTAGES - 1 K/V tiles (cp.async) and return - the fused
// decode stack runs this BEFORE the grid barrier that publishes the step's new K/V row, to hide HBM and TLB latency.
// MODE 2: the rest of the item after a MODE 1 call; the tile holding the newest key is loaded again first.
// STAGES: depth of the K/V tile ring (a K and a V tile per stage).  FULL: the prologue requests STAGES tiles instead of
// STAGES - 1 and a stage is refilled one iteration after it was consumed - the fused decode stack's packed-feed variant
// runs the item with 2 stages (64 KB) and still has both tiles of its <= 128-key splits in flight before the barrier.
template <int BAR_ID, int MODE = 0, int STAGES = DSTAGES, bool FULL = false>
__device__ __forceinline__ void attn_decode_body(const DecodeParams& p, int split, int kvh, int z,
                                                 unsigned char* smem, int tid) {
    const int b = z / p.groups_per_seq, grp = z % p.groups_per_seq;
    AD_TRACE(0);
    const int g = p.n_q_head / p.arena.n_kv_head;
    const int nq = p.n_q[b], len = p.kv_len[b], env = p.seq_env[b], qs = p.q_start[b];
    const int tok0 = grp * p.tok_per_group;
    if (tok0 >= nq) return;
    const int ntok = min(p.tok_per_group, nq - tok0);
    const int rows_valid = ntok * g;
    const int kv_hi = len - nq + tok0 + ntok;  // one past the last key any row of this group may see
    const int k_begin = split * p.split_keys;
    const int k_end = min(k_begin + p.split_keys, kv_hi);
    const int lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t slot = ((int64_t)z * p.arena.n_kv_head + kvh) * p.n_splits + split;

    if (MODE == 1 && k_begin >= k_end) return;
    if (k_begin >= k_end) {  // empty split: neutral partial
        if (p.n_splits == 1) {
            for (int r = 0; r < rows_valid; ++r)
                p.out[((int64_t)(qs + tok0 + r / g) * p.n_q_head + kvh * g + r % g) * HD + tid] = __float2bfloat16_rn(0.f);
            return;
        }
        if (tid < DROWS) {
            p.part_m[slot * DROWS + tid] = -INFINITY;
            p.part_l[slot * DROWS + tid] = 0.f;
        }
        return;
    }
    const int64_t plane_elems = (int64_t)p.arena.max_len * HD;
    const __nv_bfloat16* kplane = reinterpret_cast<const __nv_bfloat16*>(p.arena.base) +
        ((((int64_t)env * p.arena.n_layer + p.layer) * 2 + 0) * p.arena.n_kv_head + kvh) * plane_elems;
    const __nv_bfloat16* vplane = kplane + (int64_t)p.arena.n_kv_head * plane_elems;

    const uint32_t smem_base = smem_u32(smem);
    const int n_tiles = (k_end - k_begin + DT - 1) / DT;
    // prologue
    constexpr int PRO = FULL ? STAGES : STAGES - 1;   // tiles requested ahead of the first one consumed
    if (MODE != 2) {
#pragma unroll
        for (int s = 0; s < PRO; ++s) {
            if (s < n_tiles) {
                load_tile_async(smem_base + s * 2 * TILE_BYTES, kplane, k_begin + s * DT, k_end, tid);
                load_tile_async(smem_base + s * 2 * TILE_BYTES + TILE_BYTES, vplane, k_begin + s * DT, k_end, tid);
            }
            cp_async_commit();
        }
        if (MODE == 1) return;
    } else {
        // the prologue tiles were requested before the newest key (kv_hi - 1) was published: fetch its tile again
        // (waited for below, after the Q loads have been issued: one L2 round trip instead of two)
        const int t_new = (kv_hi - 1 - k_begin) / DT;
        if (k_end == kv_hi && t_new < PRO) {
            load_tile_async(smem_base + t_new * 2 * TILE_BYTES, kplane, k_begin + t_new * DT, k_end, tid);
            load_tile_async(smem_base + t_new * 2 * TILE_BYTES + TILE_BYTES, vplane, k_begin + t_new * DT, k_end, tid);
        }
        cp_async_commit();
    }

    // Q fragments: rows gid and gid+8; row r -> token tok0 + r / g, head kvh*g + r % g
    uint32_t qa[8][4];
    int row_lim[2];  // exclusive key limit of each row (causal)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = gid + h * 8;
        const bool valid = r < rows_valid;
        const int tq = tok0 + (valid ? r / g : 0);
        const int hq = kvh * g + (valid ? r % g : 0);
        row_lim[h] = valid ? min(len - nq + tq + 1, k_end) : 0;
        const uint32_t* qrow = reinterpret_cast<const uint32_t*>(p.q + ((int64_t)(qs + tq) * p.n_q_head + hq) * HD);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            qa[ks][h] = valid ? ld_cg_u32(qrow + ks * 8 + tig) : 0u;          // cols ks*16 + 2*tig
            qa[ks][h + 2] = valid ? ld_cg_u32(qrow + ks * 8 + 4 + tig) : 0u;  // cols ks*16 + 8 + 2*tig
        }
    }

    if (MODE == 2) cp_async_wait<0>();
    AD_TRACE(1);
    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    for (int it = 0; it < n_tiles; ++it) {
        // tile `it` must have landed.  Rolling ring (PRO = STAGES - 1): the one group that may stay in flight is the tile
        // requested in the previous iteration (it + STAGES - 2).  FULL ring: a stage is refilled one iteration after it was
        // consumed, so the newest group IS tile `it` - nothing may stay in flight (waiting with <PRO - 1> consumed splits of
        // more than STAGES tiles before their last tiles had arrived: two-stream steps, 18 splits of up to 4 tiles).
        cp_async_wait<FULL ? 0 : PRO - 1>();
        group_sync<BAR_ID>();
        if (it == 0) AD_TRACE(2);
        {   // refill the stage consumed in the previous iteration
            const int nt = it + STAGES - 1;
            if (nt < n_tiles && (!FULL || it >= 1)) {
                const int s = nt % STAGES;
                load_tile_async(smem_base + s * 2 * TILE_BYTES, kplane, k_begin + nt * DT, k_end, tid);
                load_tile_async(smem_base + s * 2 * TILE_BYTES + TILE_BYTES, vplane, k_begin + nt * DT, k_end, tid);
            }
            cp_async_commit();
        }
        const uint32_t ktile = smem_base + (it % STAGES) * 2 * TILE_BYTES;
        const uint32_t vtile = ktile + TILE_BYTES;
        const int key_base = k_begin + it * DT + warp * 16;
        if (key_base >= k_end) continue;  // warp-uniform: this warp's 16 keys are all out of range

        // S = Q K^T for 16 keys (two n-tiles of 8)
        float s[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            // matrices: m0 = keys 0-7 x dims lo8, m1 = keys 0-7 x dims hi8, m2 = keys 8-15 x lo8, m3 = keys 8-15 x hi8
            const int mi = lane >> 3;
            const int row = warp * 16 + (mi >> 1) * 8 + (lane & 7);
            const int chunk = ks * 2 + (mi & 1);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(ktile + tile_off(row, chunk), b0, b1, b2, b3);
            mma_bf16(s[0], qa[ks], b0, b1);
            mma_bf16(s[1], qa[ks], b2, b3);
        }
        // causal / range mask, online softmax (rows gid -> s[n][0..1], gid+8 -> s[n][2..3])
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = key_base + n * 8 + 2 * tig + (e & 1);
                const int h = e >> 1;
                if (key >= row_lim[h]) s[n][e] = -INFINITY;
                mx[h] = fmaxf(mx[h], s[n][e]);
            }
        float alpha[2], m_use[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);
            m_use[h] = m_new == -INFINITY ? 0.f : m_new;
            alpha[h] = m_run[h] == -INFINITY ? 0.f : exp2f((m_run[h] - m_use[h]) * p.scale_log2);
            m_run[h] = m_new;
        }
        uint32_t pa[4];
        float psum[2] = {0.f, 0.f};
        {
            float pv[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int h = e >> 1;
                    const float pe = exp2f((s[n][e] - m_use[h]) * p.scale_log2);  // exp2(-inf) = 0
                    // P is rounded to bf16 for the PV product (as flash-attn does); the row sum uses
                    // the same rounded values so that normalisation is consistent
                    pv[n][e] = round_bf16(pe);
                    psum[h] += pv[n][e];
                }
            pa[0] = pack_bf16(pv[0][0], pv[0][1]);
            pa[1] = pack_bf16(pv[0][2], pv[0][3]);
            pa[2] = pack_bf16(pv[1][0], pv[1][1]);
            pa[3] = pack_bf16(pv[1][2], pv[1][3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * alpha[h] + psum[h];
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            o[n][0] *= alpha[0]; o[n][1] *= alpha[0];
            o[n][2] *= alpha[1]; o[n][3] *= alpha[1];
        }
        // O += P V : B fragments via ldmatrix.trans, two 8-dim n-tiles per instruction
#pragma unroll
        for (int dn = 0; dn < 16; dn += 2) {
            const int mi = lane >> 3;
            const int row = warp * 16 + (mi & 1) * 8 + (lane & 7);
            const int chunk = dn + (mi >> 1);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(vtile + tile_off(row, chunk), b0, b1, b2, b3);
            mma_bf16(o[dn], pa, b0, b1);
            mma_bf16(o[dn + 1], pa, b2, b3);
        }
    }
    AD_TRACE(3);
    cp_async_wait<0>();
    group_sync<BAR_ID>();

    // ---- merge the four warps (each saw a disjoint key subset) through shared memory
    float* s_o = reinterpret_cast<float*>(smem);                  // [4][16][128] fp32 = 32 KB
    float* s_m = s_o + DWARPS * DROWS * HD;                       // [4][16]
    float* s_l = s_m + DWARPS * DROWS;                            // [4][16]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
        if (tig == 0) {
            s_m[warp * DROWS + gid + h * 8] = m_run[h];
            s_l[warp * DROWS + gid + h * 8] = l_run[h];
        }
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        float* base = s_o + (warp * DROWS) * HD + n * 8 + 2 * tig;
        *reinterpret_cast<float2*>(base + gid * HD) = make_float2(o[n][0], o[n][1]);
        *reinterpret_cast<float2*>(base + (gid + 8) * HD) = make_float2(o[n][2], o[n][3]);
    }
    group_sync<BAR_ID>();
    AD_TRACE(4);
    // per-(warp,row) rescale factors once, then 128 threads = 128 dims with every row independent
    float* s_f = s_l + DWARPS * DROWS;                            // [4][16]
    if (tid < DWARPS * DROWS) {
        const int w = tid >> 4, r = tid & 15;
        float m = -INFINITY;
#pragma unroll
        for (int ww = 0; ww < DWARPS; ++ww) m = fmaxf(m, s_m[ww * DROWS + r]);
        const float mw = s_m[w * DROWS + r];
        s_f[tid] = mw == -INFINITY ? 0.f : exp2f((mw - m) * p.scale_log2);
        if (w == 0 && r < rows_valid) p.part_m[slot * DROWS + r] = m;
    }
    group_sync<BAR_ID>();
    float* s_lrow = s_f + DWARPS * DROWS;                         // [16]
    if (tid < rows_valid) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < DWARPS; ++w) l += s_l[w * DROWS + tid] * s_f[w * DROWS + tid];
        if (p.n_splits == 1) s_lrow[tid] = l;
        else p.part_l[slot * DROWS + tid] = l;
    }
    if (p.n_splits == 1) {
        // a single split owns the whole key range: normalise and store the final rows here (same arithmetic as
        // attn_decode_combine_kernel with one split of weight 1), no partials, no second kernel
        group_sync<BAR_ID>();
#pragma unroll
        for (int r = 0; r < DROWS; ++r) {
            if (r < rows_valid) {
                float acc = 0.f;
#pragma unroll
                for (int w = 0; w < DWARPS; ++w) acc += s_o[(w * DROWS + r) * HD + tid] * s_f[w * DROWS + r];
                const float l = s_lrow[r];
                const int tq = tok0 + r / g, hq = kvh * g + r % g;
                p.out[((int64_t)(qs + tq) * p.n_q_head + hq) * HD + tid] = __float2bfloat16_rn(l > 0.f ? acc / l : 0.f);
            }
        }
        return;
    }
    float* po = p.part_o + (slot * DROWS) * HD + tid;
#pragma unroll
    for (int r = 0; r < DROWS; ++r) {
        if (r < rows_valid) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < DWARPS; ++w) acc += s_o[(w * DROWS + r) * HD + tid] * s_f[w * DROWS + r];
            po[r * HD] = acc;
        }
    }
    AD_TRACE(5);
}


}  // namespace svln

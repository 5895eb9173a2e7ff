// Split-KV attention for the action-token decode (and any short chunk): an HBM-streaming kernel.
// Replaces flash-attn 2's decode path for HF Qwen2 attention (streamvln_eval.py:526; reached
// from stream_video_vln.py:340).  GQA-packed: the g = n_q_head / n_kv_head query heads that
// share one KV head (and up to 16/g query tokens) form the 16 rows of one m16n8k16 tile, so
// every K/V byte is read from HBM once per (stream, kv head) and used by all of them.
// K/V tiles stream through a 3-stage cp.async ring with a 128-byte XOR swizzle; the scalar
// fp32 pipe could not keep up with HBM for g = 7 (3.5 FMA per byte), so the dot products run
// on the legacy tensor path (HMMA) — the tcgen05 path is reserved for the dense prefill.
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
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
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
};

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

__global__ void __launch_bounds__(DWARPS * 32)
attn_decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int split = blockIdx.x, kvh = blockIdx.y;
    const int b = blockIdx.z / p.groups_per_seq, grp = blockIdx.z % p.groups_per_seq;
    const int g = p.n_q_head / p.arena.n_kv_head;
    const int nq = p.n_q[b], len = p.kv_len[b], env = p.seq_env[b], qs = p.q_start[b];
    const int tok0 = grp * p.tok_per_group;
    if (tok0 >= nq) return;
    const int ntok = min(p.tok_per_group, nq - tok0);
    const int rows_valid = ntok * g;
    const int kv_hi = len - nq + tok0 + ntok;  // one past the last key any row of this group may see
    const int k_begin = split * p.split_keys;
    const int k_end = min(k_begin + p.split_keys, kv_hi);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t slot = ((int64_t)blockIdx.z * p.arena.n_kv_head + kvh) * p.n_splits + split;

    if (k_begin >= k_end) {  // empty split: neutral partial
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
#pragma unroll
    for (int s = 0; s < DSTAGES - 1; ++s) {
        if (s < n_tiles) {
            load_tile_async(smem_base + s * 2 * TILE_BYTES, kplane, k_begin + s * DT, k_end, tid);
            load_tile_async(smem_base + s * 2 * TILE_BYTES + TILE_BYTES, vplane, k_begin + s * DT, k_end, tid);
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
            qa[ks][h] = valid ? __ldg(qrow + ks * 8 + tig) : 0u;          // cols ks*16 + 2*tig
            qa[ks][h + 2] = valid ? __ldg(qrow + ks * 8 + 4 + tig) : 0u;  // cols ks*16 + 8 + 2*tig
        }
    }

    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    for (int it = 0; it < n_tiles; ++it) {
        cp_async_wait<DSTAGES - 2>();
        __syncthreads();
        {   // refill the stage consumed in the previous iteration
            const int nt = it + DSTAGES - 1;
            if (nt < n_tiles) {
                const int s = nt % DSTAGES;
                load_tile_async(smem_base + s * 2 * TILE_BYTES, kplane, k_begin + nt * DT, k_end, tid);
                load_tile_async(smem_base + s * 2 * TILE_BYTES + TILE_BYTES, vplane, k_begin + nt * DT, k_end, tid);
            }
            cp_async_commit();
        }
        const uint32_t ktile = smem_base + (it % DSTAGES) * 2 * TILE_BYTES;
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
    cp_async_wait<0>();
    __syncthreads();

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
    __syncthreads();
    // 128 threads: thread t handles dim t of every row
    for (int r = 0; r < rows_valid; ++r) {
        float m = -INFINITY;
#pragma unroll
        for (int w = 0; w < DWARPS; ++w) m = fmaxf(m, s_m[w * DROWS + r]);
        float acc = 0.f, l = 0.f;
#pragma unroll
        for (int w = 0; w < DWARPS; ++w) {
            const float mw = s_m[w * DROWS + r];
            const float f = mw == -INFINITY ? 0.f : exp2f((mw - m) * p.scale_log2);
            acc += s_o[(w * DROWS + r) * HD + tid] * f;
            l += s_l[w * DROWS + r] * f;
        }
        p.part_o[(slot * DROWS + r) * HD + tid] = acc;
        if (tid == 0) {
            p.part_m[slot * DROWS + r] = m;
            p.part_l[slot * DROWS + r] = l;
        }
    }
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
    static bool attr_set = false;
    if (!attr_set) {
        SVLN_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM));
        attr_set = true;
    }
    dim3 grid(p.n_splits, arena->n_kv_head, (unsigned)groups);
    attn_decode_kernel<<<grid, DWARPS * 32, DEC_SMEM, (cudaStream_t)stream>>>(p);
    dim3 cgrid((unsigned)groups, arena->n_kv_head, DROWS);
    attn_decode_combine_kernel<<<cgrid, HD, 0, (cudaStream_t)stream>>>(p);
    return check_launch("attn_decode kernels");
}

}  // extern "C"

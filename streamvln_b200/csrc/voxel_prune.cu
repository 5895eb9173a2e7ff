// Slow-memory voxel token pruning: voxel keys, keep-latest dedup, stable compaction, gather.
// Integer / indexing kernels, HBM-bound; bit-exact against oracle/voxel_oracle.c.
// Spec: DESIGN.md §3 (the reference plumbs depths/poses/intrinsics but never uses them:
// stream_video_vln.py:102-142).
#include "svln_common.cuh"

namespace svln {

// ------------------------------------------------------------------------------ voxel keys
// fp32 with a fixed operation order and no FMA contraction (explicit _rn intrinsics), so the
// CPU oracle (compiled with -ffp-contract=off) produces the same bits.
__device__ __forceinline__ int32_t voxel_coord(float x, float voxel_size) {
    float q = floorf(__fdiv_rn(x, voxel_size));
    q = fminf(fmaxf(q, -1048576.0f), 1048575.0f);
    return (int32_t)q;
}

struct PruneSetup {               // optional by-products of the keys kernel (t_out == nullptr: keys only)
    int32_t* t_out;
    int64_t* seg_off;
    int64_t* range_off;
    unsigned* ticket;
    int64_t n_segments, n_streams, stream_tokens, seg_tokens, n_tokens;
    int pps, n_frames;
};

__global__ void __launch_bounds__(256)
voxel_keys_kernel(const float* __restrict__ depth, const float* __restrict__ pose,
                  const float* __restrict__ intr, int H, int W, int G, float voxel_size,
                  float max_depth, int64_t* __restrict__ key_out, PruneSetup su) {
    const int64_t f = blockIdx.x;
    if (su.t_out) {
        // the fused pipeline's bookkeeping rides along (it was three more launches): frame times, the (stream, period)
        // dedup segment offsets, the per-stream range offsets and the compaction's arrival counter
        const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // grid threads >= n_segments + 1
        if (g <= su.n_segments) {
            const int64_t s = g / su.pps, p = g % su.pps;
            const int64_t o = s * su.stream_tokens + p * su.seg_tokens;
            const int64_t lim = (s + 1) * su.stream_tokens;
            su.seg_off[g] = o > lim ? lim : o;
        }
        if (g <= su.n_streams) su.range_off[g] = min(g * su.stream_tokens, su.n_tokens);
        if (g == 0) *su.ticket = 0;
    }
    const float* P = pose + f * 16;
    const float* K = intr + f * 16;
    const float fx = K[0], cx = K[2], fy = K[5], cy = K[6];
    const float r00 = P[0], r01 = P[1], r02 = P[2], tx = P[3];
    const float r10 = P[4], r11 = P[5], r12 = P[6], ty = P[7];
    const float r20 = P[8], r21 = P[9], r22 = P[10], tz = P[11];
    const float sw = __fdiv_rn((float)W, (float)G);
    const float sh = __fdiv_rn((float)H, (float)G);
    const float* dframe = depth + f * (int64_t)H * W;
    const int n_tok = G * G;
    for (int tok = threadIdx.x; tok < n_tok; tok += blockDim.x) {
        const int i = tok / G, j = tok - i * G;
        const float u = __fmul_rn(__fadd_rn((float)j, 0.5f), sw);
        const float v = __fmul_rn(__fadd_rn((float)i, 0.5f), sh);
        int px = (int)floorf(u), py = (int)floorf(v);
        px = min(max(px, 0), W - 1);
        py = min(max(py, 0), H - 1);
        const float d = __ldg(dframe + (int64_t)py * W + px);
        int64_t key = SVLN_KEY_INVALID;
        // d > 0 is false for NaN; isfinite rejects +inf
        if (d > 0.0f && d < max_depth && isfinite(d)) {
            const float xc = __fmul_rn(d, __fdiv_rn(__fsub_rn(u, cx), fx));
            const float yc = __fmul_rn(d, __fdiv_rn(__fsub_rn(v, cy), fy));
            const float zc = d;
            const float xw = __fadd_rn(
                __fadd_rn(__fadd_rn(__fmul_rn(r00, xc), __fmul_rn(r01, yc)), __fmul_rn(r02, zc)),
                tx);
            const float yw = __fadd_rn(
                __fadd_rn(__fadd_rn(__fmul_rn(r10, xc), __fmul_rn(r11, yc)), __fmul_rn(r12, zc)),
                ty);
            const float zw = __fadd_rn(
                __fadd_rn(__fadd_rn(__fmul_rn(r20, xc), __fmul_rn(r21, yc)), __fmul_rn(r22, zc)),
                tz);
            const int64_t vx = (int64_t)voxel_coord(xw, voxel_size) + 1048576;
            const int64_t vy = (int64_t)voxel_coord(yw, voxel_size) + 1048576;
            const int64_t vz = (int64_t)voxel_coord(zw, voxel_size) + 1048576;
            key = (vx << 42) | (vy << 21) | vz;
        }
        key_out[f * n_tok + tok] = key;
        if (su.t_out) su.t_out[f * n_tok + tok] = (int32_t)(f % su.n_frames);
    }
}

// --------------------------------------------------------------------- keep-latest dedup
struct __align__(16) Slot {
    unsigned long long key;   // 0xFFFF... = empty (== SVLN_KEY_INVALID, never inserted)
    unsigned long long best;  // min over ((~t) << 32 | local index): latest frame, lowest index
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ int64_t find_segment(const int64_t* __restrict__ seg_off,
                                                int64_t n_segments, int64_t uniform, int64_t n) {
    if (uniform > 0) return n / uniform;
    int64_t lo = 0, hi = n_segments;  // seg_off[lo] <= n < seg_off[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= n) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ unsigned long long pack_rank(int32_t t, int64_t local) {
    return ((unsigned long long)(0xFFFFFFFFu - (uint32_t)t) << 32) | (unsigned long long)(uint32_t)local;
}

__global__ void __launch_bounds__(256)
dedup_insert_kernel(const int64_t* __restrict__ key, const int32_t* __restrict__ t, int64_t n_tokens,
                    const int64_t* __restrict__ seg_off, int64_t n_segments, int64_t uniform,
                    Slot* __restrict__ table, uint32_t cap_mask) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_tokens) return;
    const int64_t k = key[n];
    if (k < 0) return;
    const int64_t g = find_segment(seg_off, n_segments, uniform, n);
    const int64_t seg_start = uniform > 0 ? g * uniform : __ldg(seg_off + g);
    const unsigned long long rank = pack_rank(t[n], n - seg_start);
    Slot* seg_table = table + g * ((int64_t)cap_mask + 1);
    uint32_t slot = (uint32_t)mix64((uint64_t)k) & cap_mask;
    while (true) {
        const unsigned long long prev =
            atomicCAS(&seg_table[slot].key, 0xFFFFFFFFFFFFFFFFull, (unsigned long long)k);
        if (prev == 0xFFFFFFFFFFFFFFFFull || prev == (unsigned long long)k) {
            atomicMin(&seg_table[slot].best, rank);
            return;
        }
        slot = (slot + 1) & cap_mask;
    }
}

__global__ void __launch_bounds__(256)
dedup_resolve_kernel(const int64_t* __restrict__ key, const int32_t* __restrict__ t, int64_t n_tokens,
                     const int64_t* __restrict__ seg_off, int64_t n_segments, int64_t uniform,
                     const Slot* __restrict__ table, uint32_t cap_mask, int invalid_policy,
                     uint8_t* __restrict__ mask) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_tokens) return;
    const int64_t k = key[n];
    if (k < 0) {
        mask[n] = invalid_policy == SVLN_INVALID_KEEP ? 1 : 0;
        return;
    }
    const int64_t g = find_segment(seg_off, n_segments, uniform, n);
    const int64_t seg_start = uniform > 0 ? g * uniform : __ldg(seg_off + g);
    const unsigned long long rank = pack_rank(t[n], n - seg_start);
    const Slot* seg_table = table + g * ((int64_t)cap_mask + 1);
    uint32_t slot = (uint32_t)mix64((uint64_t)k) & cap_mask;
    while (true) {
        const Slot s = seg_table[slot];
        if (s.key == (unsigned long long)k) {
            mask[n] = s.best == rank ? 1 : 0;
            return;
        }
        slot = (slot + 1) & cap_mask;
    }
}

// Shared-memory staged variant: one CTA owns one segment whose table fits in shared memory
// (<= 8192 tokens -> 16384 slots x (8 B key + 4 B best) = 192 KB).  No global table traffic and
// no memset.  `best` holds the segment-local index of the current winner; candidates are
// compared by (t desc, index asc) with a CAS loop, so any t values are handled.
constexpr int SMEM_DEDUP_MAX_TOKENS = 8192;
constexpr int SMEM_DEDUP_THREADS = 1024;
constexpr size_t SMEM_DEDUP_MAX_BYTES = 16384 * 12;

__global__ void __launch_bounds__(SMEM_DEDUP_THREADS, 1)
dedup_smem_kernel(const int64_t* __restrict__ key, const int32_t* __restrict__ t,
                  const int64_t* __restrict__ seg_off, int64_t uniform, int64_t n_tokens,
                  uint32_t cap_mask, int invalid_policy, uint8_t* __restrict__ mask) {
    // gridDim.y CTAs share a segment by KEY: CTA y owns the voxels whose hash falls in its partition (every CTA scans the
    // whole segment, inserts and resolves only its own keys), so the winner of a voxel is still decided inside one CTA
    // and the result does not depend on the partition count.  64 segments used to mean 64 busy SMs of 148.
    const uint32_t n_part = gridDim.y, my_part = blockIdx.y;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t cap = cap_mask + 1;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem_raw);
    uint32_t* sbest = reinterpret_cast<uint32_t*>(skey + cap);
    const int64_t g = blockIdx.x;
    int64_t seg_start, seg_end;
    if (uniform > 0) {
        seg_start = g * uniform;
        seg_end = min(seg_start + uniform, n_tokens);
    } else {
        seg_start = seg_off[g];
        seg_end = seg_off[g + 1];
    }
    const int seg_n = (int)(seg_end - seg_start);
    const int32_t* tt = t + seg_start;
    for (uint32_t s = threadIdx.x; s < cap; s += blockDim.x) {
        skey[s] = 0xFFFFFFFFFFFFFFFFull;
        sbest[s] = 0xFFFFFFFFu;
    }
    // Frame-time range of the segment.  When it fits 19 bits (always, for step counters) the winner of a voxel is ONE
    // shared-memory atomicMin on the packed word ((t_max - t) << 13 | index): largest t first, lowest index on ties.
    // Round 1 kept only the index and re-read t[current best] from global memory inside a CAS loop - a dependent L2
    // round trip per contended insert, which is what made this stage latency-bound.  Wider ranges take that path still.
    __shared__ int s_tmin, s_tmax;
    if (threadIdx.x == 0) { s_tmin = 0x7FFFFFFF; s_tmax = (int)0x80000000; }
    __syncthreads();
    {
        int lo = 0x7FFFFFFF, hi = (int)0x80000000;
        for (int i = threadIdx.x; i < seg_n; i += blockDim.x) {
            const int v = tt[i];
            lo = min(lo, v);
            hi = max(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0 && seg_n > 0) { atomicMin(&s_tmin, lo); atomicMax(&s_tmax, hi); }
    }
    __syncthreads();
    const int tmax = s_tmax;
    const bool packed = seg_n <= 8192 && (int64_t)tmax - (int64_t)s_tmin < (1 << 19);
    for (int i = threadIdx.x; i < seg_n; i += blockDim.x) {
        const int64_t k = key[seg_start + i];
        if (k < 0) continue;
        const uint64_t hsh = mix64((uint64_t)k);
        if ((uint32_t)(hsh >> 40) % n_part != my_part) continue;
        const int32_t ti = tt[i];
        uint32_t slot = (uint32_t)hsh & cap_mask;
        while (true) {
            const unsigned long long prev =
                atomicCAS(&skey[slot], 0xFFFFFFFFFFFFFFFFull, (unsigned long long)k);
            if (prev == 0xFFFFFFFFFFFFFFFFull || prev == (unsigned long long)k) break;
            slot = (slot + 1) & cap_mask;
        }
        if (packed) {
            atomicMin(&sbest[slot], ((uint32_t)(tmax - ti) << 13) | (uint32_t)i);
            continue;
        }
        uint32_t cur = sbest[slot];
        while (true) {
            bool better = cur == 0xFFFFFFFFu;
            if (!better) {
                const int32_t tc = tt[cur];
                better = ti > tc || (ti == tc && (uint32_t)i < cur);
            }
            if (!better) break;
            const uint32_t seen = atomicCAS(&sbest[slot], cur, (uint32_t)i);
            if (seen == cur) break;
            cur = seen;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < seg_n; i += blockDim.x) {
        const int64_t k = key[seg_start + i];
        uint8_t m;
        if (k < 0) {
            if (my_part != 0) continue;
            m = invalid_policy == SVLN_INVALID_KEEP ? 1 : 0;
        } else {
            const uint64_t hsh = mix64((uint64_t)k);
            if ((uint32_t)(hsh >> 40) % n_part != my_part) continue;
            uint32_t slot = (uint32_t)hsh & cap_mask;
            while (skey[slot] != (unsigned long long)k) slot = (slot + 1) & cap_mask;
            const uint32_t best = sbest[slot];
            m = (packed ? (best & 0x1FFFu) : best) == (uint32_t)i ? 1 : 0;
        }
        mask[seg_start + i] = m;
    }
}

// ------------------------------------------------------------------------ frame threshold
__global__ void __launch_bounds__(128)
frame_threshold_kernel(uint8_t* __restrict__ mask, int tokens_per_frame, int min_keep) {
    __shared__ int s_cnt;
    uint8_t* m = mask + (int64_t)blockIdx.x * tokens_per_frame;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < tokens_per_frame; i += blockDim.x) c += m[i] != 0;
    c = (int)warp_sum((float)c);  // exact for counts < 2^24
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (s_cnt < min_keep)
        for (int i = threadIdx.x; i < tokens_per_frame; i += blockDim.x) m[i] = 0;
}

// ------------------------------------------------------------------- stable compaction
constexpr int CB = 1024;  // tokens per compaction block (256 threads x 4 mask bytes)

__device__ __forceinline__ uint32_t load_mask4(const uint8_t* __restrict__ mask, int64_t base,
                                               int64_t n_tokens) {
    // returns a 4-bit pattern of kept tokens base..base+3
    uint32_t bits = 0;
    if (base + 3 < n_tokens && ((reinterpret_cast<uintptr_t>(mask + base) & 3) == 0)) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(mask + base);
        bits = ((w & 0xFFu) ? 1u : 0u) | ((w & 0xFF00u) ? 2u : 0u) | ((w & 0xFF0000u) ? 4u : 0u) |
               ((w & 0xFF000000u) ? 8u : 0u);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (base + q < n_tokens && mask[base + q]) bits |= 1u << q;
    }
    return bits;
}

// exclusive scan of block counts in place by one CTA of NT threads; blk_cnt[n_blocks] receives the total
template <int NT>
__device__ __forceinline__ void scan_block_counts(int32_t* blk_cnt, int64_t n_blocks, int32_t* __restrict__ n_keep_out) {
    __shared__ int s_sw[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NT / 32;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += NT) {
        const int64_t b = b0 + threadIdx.x;
        const int v = b < n_blocks ? (int)ld_cg_u32(blk_cnt + b) : 0;   // written by other CTAs of this launch
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_sw[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = lane < NW ? s_sw[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_sw[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp ? s_sw[warp - 1] : 0) + inc - v;
        if (b < n_blocks) blk_cnt[b] = excl;
        __syncthreads();
        if (threadIdx.x == NT - 1) s_carry = carry + s_sw[NW - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        blk_cnt[n_blocks] = s_carry;
        if (n_keep_out) *n_keep_out = s_carry;
    }
}

// per-block survivor counts; the LAST block to arrive (ticket counter, zeroed by the caller's previous launch) scans them,
// so counting and scanning are one launch
__global__ void __launch_bounds__(256)
compact_count_scan_kernel(const uint8_t* __restrict__ mask, int64_t n_tokens, int32_t* __restrict__ blk_cnt,
                          unsigned* __restrict__ ticket, int32_t* __restrict__ n_keep_out) {
    __shared__ int s_w[8];
    __shared__ bool s_last;
    const int64_t base = (int64_t)blockIdx.x * CB + threadIdx.x * 4;
    int c = __popc(load_mask4(mask, base, n_tokens));
    c = (int)warp_sum((float)c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_w[w];
        __stcg(blk_cnt + blockIdx.x, tot);
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    scan_block_counts<256>(blk_cnt, (int64_t)gridDim.x, n_keep_out);
    if (threadIdx.x == 0) *ticket = 0;      // ready for the next call on this workspace
}

// n_tokens == 0: nothing to count, the outputs are zeros
__global__ void __launch_bounds__(32)
compact_empty_kernel(int32_t* __restrict__ blk_cnt, int32_t* __restrict__ n_keep_out) {
    if (threadIdx.x == 0) {
        blk_cnt[0] = 0;
        if (n_keep_out) *n_keep_out = 0;
    }
}

__global__ void __launch_bounds__(256)
compact_write_kernel(const uint8_t* __restrict__ mask, int64_t n_tokens,
                     const int32_t* __restrict__ blk_off, int32_t* __restrict__ idx_out, int n_write_blocks,
                     const int64_t* __restrict__ range_off, int64_t n_offsets, int32_t* __restrict__ range_out) {
    __shared__ int s_w[8];
    if ((int)blockIdx.x >= n_write_blocks) {
        // trailing CTAs: survivors in front of each range boundary (per-stream keep offsets), one WARP per boundary (a
        // thread per boundary walked up to 1023 mask bytes one dependent load at a time: 27 us for 65 boundaries)
        const int64_t r = (int64_t)(blockIdx.x - n_write_blocks) * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (r >= n_offsets) return;
        const int lane = threadIdx.x & 31;
        int64_t off = range_off[r];
        off = off < 0 ? 0 : (off > n_tokens ? n_tokens : off);
        const int64_t b = off / CB;
        int c = 0;
        for (int64_t i = b * CB + lane; i < off; i += 32) c += mask[i] != 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) range_out[r] = blk_off[b] + c;
        return;
    }
    const int64_t base = (int64_t)blockIdx.x * CB + threadIdx.x * 4;
    const uint32_t bits = load_mask4(mask, base, n_tokens);
    const int c = __popc(bits);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) woff += (w < warp) ? s_w[w] : 0;
    int pos = blk_off[blockIdx.x] + woff + inc - c;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (bits & (1u << q)) idx_out[pos++] = (int32_t)(base + q);
}

// --------------------------------------------------------------------------- row gather
// One warp per surviving row; every lane keeps GU 128-bit loads in flight before storing.
constexpr int GU = 8;

__global__ void __launch_bounds__(256)
gather_rows_kernel(const int4* __restrict__ src, int row_i4, const int32_t* __restrict__ idx,
                   const int32_t* __restrict__ n_rows_dev, int4* __restrict__ dst) {
    const int n_rows = *n_rows_dev;
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps_total) {
        const int4* s = src + (int64_t)__ldg(idx + r) * row_i4;
        int4* d = dst + (int64_t)r * row_i4;
        for (int c0 = 0; c0 < row_i4; c0 += 32 * GU) {
            int4 v[GU];
#pragma unroll
            for (int q = 0; q < GU; ++q) {
                const int c = c0 + q * 32 + lane;
                if (c < row_i4) v[q] = ld_stream(s + c);
            }
#pragma unroll
            for (int q = 0; q < GU; ++q) {
                const int c = c0 + q * 32 + lane;
                if (c < row_i4) st_stream(d + c, v[q]);
            }
        }
    }
}

// splice: row_src >= 0 -> embedding row; < 0 -> feature row (-row_src-1)
__global__ void __launch_bounds__(256)
splice_rows_kernel(const int4* __restrict__ embed, const int4* __restrict__ feat,
                   const int32_t* __restrict__ row_src, int64_t n_rows, int row_i4,
                   int4* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps_total) {
        const int32_t sidx = __ldg(row_src + r);
        const int4* s = sidx >= 0 ? embed + (int64_t)sidx * row_i4 : feat + (int64_t)(-sidx - 1) * row_i4;
        int4* d = dst + r * row_i4;
        for (int c0 = 0; c0 < row_i4; c0 += 32 * GU) {
            int4 v[GU];
#pragma unroll
            for (int q = 0; q < GU; ++q) {
                const int c = c0 + q * 32 + lane;
                if (c < row_i4) v[q] = ld_stream(s + c);
            }
#pragma unroll
            for (int q = 0; q < GU; ++q) {
                const int c = c0 + q * 32 + lane;
                if (c < row_i4) st_stream(d + c, v[q]);
            }
        }
    }
}

// ------------------------------------------------------------------------------ host side
static uint32_t table_cap(int64_t seg_tokens) {
    uint32_t cap = 64;
    while ((int64_t)cap < 2 * seg_tokens) cap <<= 1;
    return cap;
}

struct PruneWs {
    size_t table_bytes, scan_bytes, head_bytes, t_off, seg_off_off, range_off_off, total;
};
static PruneWs prune_ws(int64_t n_tokens, int64_t n_segments, int64_t seg_tokens) {
    PruneWs w;
    // segments that fit the shared-memory table need no global table, but callers may still
    // route through the global path (ragged segment lists), so always reserve it.
    w.table_bytes = (size_t)n_segments * table_cap(seg_tokens) * sizeof(Slot);
    w.scan_bytes = ((size_t)(n_tokens + CB - 1) / CB + 2) * sizeof(int32_t);
    w.head_bytes = align_up(w.table_bytes > w.scan_bytes ? w.table_bytes : w.scan_bytes, 256);
    w.t_off = w.head_bytes;
    w.seg_off_off = w.t_off + align_up((size_t)n_tokens * 4, 256);
    w.range_off_off = w.seg_off_off + align_up((size_t)(n_segments + 2) * 8, 256);
    w.total = w.range_off_off + align_up((size_t)(n_segments + 2) * 8, 256);
    return w;
}

static int keep_latest_impl(const int64_t* key, const int32_t* t, int64_t n_tokens,
                            const int64_t* seg_off, int64_t n_segments, int64_t max_seg_tokens,
                            int64_t uniform, int invalid_policy, void* workspace,
                            size_t workspace_bytes, uint8_t* mask_out, cudaStream_t st) {
    if (n_tokens == 0) return SVLN_OK;
    const uint32_t cap = table_cap(max_seg_tokens);
    if (max_seg_tokens <= SMEM_DEDUP_MAX_TOKENS) {
        const size_t smem = (size_t)cap * 12;
        static unsigned long long attr_set = 0;
        if (first_use_on_device(&attr_set)) {
            SVLN_CUDA(cudaFuncSetAttribute(dedup_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)SMEM_DEDUP_MAX_BYTES));
        }
        // one CTA per SM (192 KB of shared memory each): spread a segment over as many CTAs as there are idle SMs
        int n_part = (int)(sm_count() / (n_segments > 0 ? n_segments : 1));
        n_part = n_part < 1 ? 1 : (n_part > 8 ? 8 : n_part);
        dedup_smem_kernel<<<dim3((unsigned)n_segments, (unsigned)n_part), SMEM_DEDUP_THREADS, smem, st>>>(
            key, t, seg_off, uniform, n_tokens, cap - 1, invalid_policy, mask_out);
        return check_launch("dedup_smem_kernel");
    }
    const size_t table_bytes = (size_t)n_segments * cap * sizeof(Slot);
    SVLN_REQUIRE(workspace && workspace_bytes >= table_bytes, SVLN_ERR_NOMEM,
                 "svln_keep_latest: workspace %zu < table %zu bytes", workspace_bytes, table_bytes);
    SVLN_CUDA(cudaMemsetAsync(workspace, 0xFF, table_bytes, st));
    const unsigned grid = (unsigned)((n_tokens + 255) / 256);
    dedup_insert_kernel<<<grid, 256, 0, st>>>(key, t, n_tokens, seg_off, n_segments, uniform,
                                              (Slot*)workspace, cap - 1);
    dedup_resolve_kernel<<<grid, 256, 0, st>>>(key, t, n_tokens, seg_off, n_segments, uniform,
                                               (const Slot*)workspace, cap - 1, invalid_policy,
                                               mask_out);
    return check_launch("dedup kernels");
}

static int compact_impl(const uint8_t* mask, int64_t n_tokens, void* workspace, size_t workspace_bytes,
                        int32_t* idx_out, int32_t* n_keep_out, const int64_t* range_off,
                        int64_t n_ranges, int32_t* range_keep_off_out, cudaStream_t st, bool ticket_zeroed = false) {
    const int64_t n_blocks = (n_tokens + CB - 1) / CB;
    const size_t need = (size_t)(n_blocks + 2) * sizeof(int32_t);
    SVLN_REQUIRE(workspace && workspace_bytes >= need, SVLN_ERR_NOMEM,
                 "svln_compact_mask: workspace %zu < %zu bytes", workspace_bytes, need);
    int32_t* blk = (int32_t*)workspace;
    unsigned* ticket = (unsigned*)(blk + n_blocks + 1);
    if (n_blocks == 0) {
        compact_empty_kernel<<<1, 32, 0, st>>>(blk, n_keep_out);
    } else {
        if (!ticket_zeroed) SVLN_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
        compact_count_scan_kernel<<<(unsigned)n_blocks, 256, 0, st>>>(mask, n_tokens, blk, ticket, n_keep_out);
    }
    const bool ranges = n_ranges > 0 && range_off && range_keep_off_out;
    const int64_t n_off = ranges ? n_ranges + 1 : 0;
    const int64_t extra = (n_off + 7) / 8;          // one warp per range boundary
    if (n_blocks + extra > 0) {
        compact_write_kernel<<<(unsigned)(n_blocks + extra), 256, 0, st>>>(mask, n_tokens, blk, idx_out, (int)n_blocks,
                                                                          range_off, n_off, range_keep_off_out);
    }
    return check_launch("compact kernels");
}

static int gather_impl(const void* src, int64_t row_bytes, const int32_t* idx, const int32_t* n_rows_dev,
                       int64_t max_rows, void* dst, cudaStream_t st) {
    if (max_rows <= 0) return SVLN_OK;
    const int64_t warps_needed = max_rows;
    int64_t blocks = (warps_needed + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;  // 8 CTAs x 8 warps = full occupancy per SM
    if (blocks > cap) blocks = cap;
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>((const int4*)src, (int)(row_bytes / 16), idx,
                                                         n_rows_dev, (int4*)dst);
    return check_launch("gather_rows_kernel");
}

}  // namespace svln

using namespace svln;

extern "C" {

int svln_voxel_keys(const float* depth, const float* pose, const float* intr, int64_t n_frames, int H,
                    int W, int G, float voxel_size, float max_depth, int64_t* key_out,
                    svln_stream_t stream) {
    SVLN_REQUIRE(depth && pose && intr && key_out, SVLN_ERR_INVALID, "svln_voxel_keys: null pointer");
    SVLN_REQUIRE(n_frames >= 0 && H > 0 && W > 0 && G > 0 && G <= 64, SVLN_ERR_INVALID,
                 "svln_voxel_keys: bad shape n_frames=%lld H=%d W=%d G=%d", (long long)n_frames, H, W, G);
    SVLN_REQUIRE(voxel_size > 0.f, SVLN_ERR_INVALID, "svln_voxel_keys: voxel_size must be > 0");
    if (n_frames == 0) return SVLN_OK;
    SVLN_REQUIRE(n_frames <= 0x7FFFFFFF, SVLN_ERR_RANGE, "svln_voxel_keys: too many frames");
    voxel_keys_kernel<<<(unsigned)n_frames, 256, 0, (cudaStream_t)stream>>>(
        depth, pose, intr, H, W, G, voxel_size, max_depth, key_out, PruneSetup{});
    return check_launch("voxel_keys_kernel");
}

size_t svln_prune_workspace_bytes(int64_t n_tokens, int64_t n_segments, int64_t seg_tokens) {
    if (n_tokens < 0 || n_segments < 0 || seg_tokens < 0) return 0;
    return prune_ws(n_tokens, n_segments, seg_tokens).total;
}

int svln_keep_latest(const int64_t* key, const int32_t* t, int64_t n_tokens, const int64_t* seg_off,
                     int64_t n_segments, int64_t max_seg_tokens, int invalid_policy, void* workspace,
                     size_t workspace_bytes, uint8_t* mask_out, svln_stream_t stream) {
    SVLN_REQUIRE(n_tokens >= 0 && n_segments >= 0 && max_seg_tokens >= 0, SVLN_ERR_INVALID,
                 "svln_keep_latest: negative size");
    if (n_tokens == 0) return SVLN_OK;
    SVLN_REQUIRE(key && t && seg_off && mask_out, SVLN_ERR_INVALID, "svln_keep_latest: null pointer");
    SVLN_REQUIRE(n_segments >= 1, SVLN_ERR_INVALID, "svln_keep_latest: need >= 1 segment");
    SVLN_REQUIRE(max_seg_tokens < (1ll << 30), SVLN_ERR_RANGE, "svln_keep_latest: segment too large");
    SVLN_REQUIRE(invalid_policy == SVLN_INVALID_KEEP || invalid_policy == SVLN_INVALID_DROP,
                 SVLN_ERR_INVALID, "svln_keep_latest: bad invalid_policy %d", invalid_policy);
    return keep_latest_impl(key, t, n_tokens, seg_off, n_segments, max_seg_tokens, 0, invalid_policy,
                            workspace, workspace_bytes, mask_out, (cudaStream_t)stream);
}

int svln_frame_threshold(uint8_t* mask, int64_t n_frames, int tokens_per_frame, int min_keep,
                         svln_stream_t stream) {
    SVLN_REQUIRE(n_frames >= 0 && tokens_per_frame > 0, SVLN_ERR_INVALID, "svln_frame_threshold: bad shape");
    if (n_frames == 0 || min_keep <= 0) return SVLN_OK;
    SVLN_REQUIRE(mask, SVLN_ERR_INVALID, "svln_frame_threshold: null mask");
    frame_threshold_kernel<<<(unsigned)n_frames, 128, 0, (cudaStream_t)stream>>>(mask, tokens_per_frame,
                                                                                  min_keep);
    return check_launch("frame_threshold_kernel");
}

int svln_compact_mask(const uint8_t* mask, int64_t n_tokens, void* workspace, size_t workspace_bytes,
                      int32_t* idx_out, int32_t* n_keep_out, const int64_t* range_off, int64_t n_ranges,
                      int32_t* range_keep_off_out, svln_stream_t stream) {
    SVLN_REQUIRE(n_tokens >= 0 && n_tokens < (1ll << 31), SVLN_ERR_RANGE, "svln_compact_mask: n_tokens out of range");
    SVLN_REQUIRE((mask && idx_out) || n_tokens == 0, SVLN_ERR_INVALID, "svln_compact_mask: null pointer");
    SVLN_REQUIRE(n_keep_out, SVLN_ERR_INVALID, "svln_compact_mask: null n_keep_out");
    return compact_impl(mask, n_tokens, workspace, workspace_bytes, idx_out, n_keep_out, range_off,
                        n_ranges, range_keep_off_out, (cudaStream_t)stream);
}

int svln_gather_rows(const void* src, int64_t row_bytes, const int32_t* idx, const int32_t* n_rows_dev,
                     int64_t max_rows, void* dst, svln_stream_t stream) {
    SVLN_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, SVLN_ERR_INVALID,
                 "svln_gather_rows: row_bytes %lld not a positive multiple of 16", (long long)row_bytes);
    SVLN_REQUIRE(max_rows >= 0, SVLN_ERR_INVALID, "svln_gather_rows: negative max_rows");
    if (max_rows == 0) return SVLN_OK;
    SVLN_REQUIRE(src && idx && n_rows_dev && dst, SVLN_ERR_INVALID, "svln_gather_rows: null pointer");
    SVLN_REQUIRE(((uintptr_t)src | (uintptr_t)dst) % 16 == 0, SVLN_ERR_INVALID, "svln_gather_rows: unaligned buffer");
    return gather_impl(src, row_bytes, idx, n_rows_dev, max_rows, dst, (cudaStream_t)stream);
}

int svln_splice_rows(const void* embed_table, const void* feat, const int32_t* row_src, int64_t n_rows,
                     int64_t row_bytes, void* dst, svln_stream_t stream) {
    SVLN_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, SVLN_ERR_INVALID, "svln_splice_rows: bad row_bytes");
    if (n_rows <= 0) return n_rows == 0 ? SVLN_OK : SVLN_ERR_INVALID;
    SVLN_REQUIRE(row_src && dst && (embed_table || feat), SVLN_ERR_INVALID, "svln_splice_rows: null pointer");
    int64_t blocks = (n_rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    splice_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const int4*)embed_table, (const int4*)feat, row_src, n_rows, (int)(row_bytes / 16), (int4*)dst);
    return check_launch("splice_rows_kernel");
}

int svln_voxel_prune(const float* depth, const float* pose, const float* intr, const void* feat,
                     int n_streams, int n_frames, int H, int W, int G, int D, float voxel_size,
                     float max_depth, int period, int min_keep, int invalid_policy, void* workspace,
                     size_t workspace_bytes, int64_t* key_out, uint8_t* mask_out, int32_t* idx_out,
                     int32_t* stream_keep_off_out, int32_t* n_keep_out, void* feat_out,
                     svln_stream_t stream) {
    SVLN_REQUIRE(n_streams >= 0 && n_frames >= 0 && G > 0 && G <= 64 && H > 0 && W > 0, SVLN_ERR_INVALID,
                 "svln_voxel_prune: bad shape S=%d T=%d G=%d H=%d W=%d", n_streams, n_frames, G, H, W);
    SVLN_REQUIRE(D > 0 && D % 8 == 0, SVLN_ERR_INVALID, "svln_voxel_prune: D=%d must be a multiple of 8", D);
    SVLN_REQUIRE(invalid_policy == SVLN_INVALID_KEEP || invalid_policy == SVLN_INVALID_DROP,
                 SVLN_ERR_INVALID, "svln_voxel_prune: bad invalid_policy %d", invalid_policy);
    SVLN_REQUIRE(voxel_size > 0.f, SVLN_ERR_INVALID, "svln_voxel_prune: voxel_size must be > 0");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t tpf = (int64_t)G * G;
    const int64_t n_tokens = (int64_t)n_streams * n_frames * tpf;
    SVLN_REQUIRE(n_tokens < (1ll << 31), SVLN_ERR_RANGE, "svln_voxel_prune: too many tokens");
    SVLN_REQUIRE(n_keep_out, SVLN_ERR_INVALID, "svln_voxel_prune: null n_keep_out");
    if (n_tokens == 0) {
        SVLN_CUDA(cudaMemsetAsync(n_keep_out, 0, 4, st));
        if (stream_keep_off_out)
            SVLN_CUDA(cudaMemsetAsync(stream_keep_off_out, 0, 4 * (size_t)(n_streams + 1), st));
        return SVLN_OK;
    }
    SVLN_REQUIRE(depth && pose && intr && key_out && mask_out && idx_out, SVLN_ERR_INVALID,
                 "svln_voxel_prune: null pointer");
    const int per = (period <= 0 || period > n_frames) ? n_frames : period;
    const int pps = (n_frames + per - 1) / per;  // periods per stream (last may be short)
    const int64_t n_segments = (int64_t)n_streams * pps;
    const int64_t seg_tokens = (int64_t)per * tpf;
    const int64_t stream_tokens = (int64_t)n_frames * tpf;
    const PruneWs w = prune_ws(n_tokens, n_segments, seg_tokens);
    SVLN_REQUIRE(workspace && workspace_bytes >= w.total, SVLN_ERR_NOMEM,
                 "svln_voxel_prune: workspace %zu < %zu bytes", workspace_bytes, w.total);
    char* ws = (char*)workspace;
    int32_t* t_arr = (int32_t*)(ws + w.t_off);
    int64_t* seg_off = (int64_t*)(ws + w.seg_off_off);
    int64_t* range_off = (int64_t*)(ws + w.range_off_off);

    // keys + frame times + segment / range offsets + the compaction's arrival counter: one launch
    const int64_t n_blocks = (n_tokens + CB - 1) / CB;
    PruneSetup su;
    su.t_out = t_arr; su.seg_off = seg_off; su.range_off = range_off;
    su.ticket = (unsigned*)((int32_t*)workspace + n_blocks + 1);
    su.n_segments = n_segments; su.n_streams = n_streams; su.stream_tokens = stream_tokens; su.seg_tokens = seg_tokens;
    su.n_tokens = n_tokens; su.pps = pps; su.n_frames = n_frames;
    SVLN_REQUIRE((int64_t)n_streams * n_frames <= 0x7FFFFFFF, SVLN_ERR_RANGE, "svln_voxel_prune: too many frames");
    voxel_keys_kernel<<<(unsigned)((int64_t)n_streams * n_frames), 256, 0, st>>>(depth, pose, intr, H, W, G, voxel_size,
                                                                                 max_depth, key_out, su);
    int rc = check_launch("voxel_keys_kernel");
    if (rc) return rc;
    // uniform fast path (no binary search) when every segment has the same length
    const int64_t uniform = (n_frames % per == 0) ? seg_tokens : 0;
    rc = keep_latest_impl(key_out, t_arr, n_tokens, seg_off, n_segments, seg_tokens, uniform,
                          invalid_policy, workspace, w.head_bytes, mask_out, st);
    if (rc) return rc;
    rc = svln_frame_threshold(mask_out, (int64_t)n_streams * n_frames, (int)tpf, min_keep, stream);
    if (rc) return rc;
    // the global-table dedup path memsets the head of the workspace (the ticket lives there): re-zero it in that case
    rc = compact_impl(mask_out, n_tokens, workspace, w.head_bytes, idx_out, n_keep_out, range_off,
                      stream_keep_off_out ? n_streams : 0, stream_keep_off_out, st,
                      /*ticket_zeroed=*/seg_tokens <= SMEM_DEDUP_MAX_TOKENS);
    if (rc) return rc;
    if (feat && feat_out) rc = gather_impl(feat, (int64_t)D * 2, idx_out, n_keep_out, n_tokens, feat_out, st);
    return rc;
}

static size_t a256(size_t x) { return align_up(x, 256); }

struct HostStaging {
    size_t depth, pose, intr, feat, key, mask, idx, keep_off, n_keep, feat_out, ws, total, ws_bytes;
};
static HostStaging host_staging(int S, int T, int H, int W, int G, int D) {
    HostStaging h;
    const size_t frames = (size_t)S * T, n = frames * G * G;
    size_t o = 0;
    h.depth = o; o += a256(frames * H * W * 4);
    h.pose = o; o += a256(frames * 64);
    h.intr = o; o += a256(frames * 64);
    h.feat = o; o += a256(n * D * 2);
    h.key = o; o += a256(n * 8);
    h.mask = o; o += a256(n);
    h.idx = o; o += a256(n * 4);
    h.keep_off = o; o += a256((size_t)(S + 1) * 4);
    h.n_keep = o; o += 256;
    h.feat_out = o; o += a256(n * D * 2);
    h.ws = o;
    // worst case for the workspace: one global period per stream
    h.ws_bytes = prune_ws((int64_t)n, (int64_t)S * T, (int64_t)T * G * G).total;
    o += a256(h.ws_bytes);
    h.total = o;
    return h;
}

size_t svln_voxel_prune_host_staging_bytes(int n_streams, int n_frames, int H, int W, int G, int D) {
    if (n_streams < 0 || n_frames < 0 || H <= 0 || W <= 0 || G <= 0 || D <= 0) return 0;
    return host_staging(n_streams, n_frames, H, W, G, D).total;
}

int svln_voxel_prune_host(const float* depth_host, const float* pose_host, const float* intr_host,
                          const void* feat_host, int n_streams, int n_frames, int H, int W, int G, int D,
                          float voxel_size, float max_depth, int period, int min_keep, int invalid_policy,
                          void* dev_staging, size_t dev_staging_bytes, uint8_t* mask_out_host,
                          int32_t* idx_out_host, int32_t* stream_keep_off_out_host,
                          int32_t* n_keep_out_host, void* feat_out_host, svln_stream_t stream) {
    SVLN_REQUIRE(n_streams >= 0 && n_frames >= 0 && H > 0 && W > 0 && G > 0 && D > 0, SVLN_ERR_INVALID,
                 "svln_voxel_prune_host: bad shape");
    SVLN_REQUIRE(n_keep_out_host, SVLN_ERR_INVALID, "svln_voxel_prune_host: null n_keep_out_host");
    const HostStaging h = host_staging(n_streams, n_frames, H, W, G, D);
    SVLN_REQUIRE(dev_staging && dev_staging_bytes >= h.total, SVLN_ERR_NOMEM,
                 "svln_voxel_prune_host: staging %zu < %zu bytes", dev_staging_bytes, h.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* d = (char*)dev_staging;
    const size_t frames = (size_t)n_streams * n_frames, n = frames * G * G;
    if (n == 0) {
        *n_keep_out_host = 0;
        if (stream_keep_off_out_host)
            for (int s = 0; s <= n_streams; ++s) stream_keep_off_out_host[s] = 0;
        return SVLN_OK;
    }
    SVLN_REQUIRE(depth_host && pose_host && intr_host, SVLN_ERR_INVALID, "svln_voxel_prune_host: null input");
    SVLN_CUDA(cudaMemcpyAsync(d + h.depth, depth_host, frames * H * W * 4, cudaMemcpyHostToDevice, st));
    SVLN_CUDA(cudaMemcpyAsync(d + h.pose, pose_host, frames * 64, cudaMemcpyHostToDevice, st));
    SVLN_CUDA(cudaMemcpyAsync(d + h.intr, intr_host, frames * 64, cudaMemcpyHostToDevice, st));
    if (feat_host) SVLN_CUDA(cudaMemcpyAsync(d + h.feat, feat_host, n * D * 2, cudaMemcpyHostToDevice, st));
    const bool want_feat = feat_host && feat_out_host;
    int rc = svln_voxel_prune((const float*)(d + h.depth), (const float*)(d + h.pose),
                              (const float*)(d + h.intr), want_feat ? d + h.feat : nullptr, n_streams,
                              n_frames, H, W, G, D, voxel_size, max_depth, period, min_keep, invalid_policy,
                              d + h.ws, h.ws_bytes, (int64_t*)(d + h.key), (uint8_t*)(d + h.mask),
                              (int32_t*)(d + h.idx), (int32_t*)(d + h.keep_off), (int32_t*)(d + h.n_keep),
                              want_feat ? d + h.feat_out : nullptr, stream);
    if (rc) return rc;
    SVLN_CUDA(cudaMemcpyAsync(n_keep_out_host, d + h.n_keep, 4, cudaMemcpyDeviceToHost, st));
    if (mask_out_host) SVLN_CUDA(cudaMemcpyAsync(mask_out_host, d + h.mask, n, cudaMemcpyDeviceToHost, st));
    if (stream_keep_off_out_host)
        SVLN_CUDA(cudaMemcpyAsync(stream_keep_off_out_host, d + h.keep_off, (size_t)(n_streams + 1) * 4,
                                  cudaMemcpyDeviceToHost, st));
    SVLN_CUDA(cudaStreamSynchronize(st));  // n_keep is needed to size the variable-length copies
    const size_t n_keep = (size_t)*n_keep_out_host;
    if (idx_out_host && n_keep)
        SVLN_CUDA(cudaMemcpyAsync(idx_out_host, d + h.idx, n_keep * 4, cudaMemcpyDeviceToHost, st));
    if (want_feat && n_keep)
        SVLN_CUDA(cudaMemcpyAsync(feat_out_host, d + h.feat_out, n_keep * D * 2, cudaMemcpyDeviceToHost, st));
    SVLN_CUDA(cudaStreamSynchronize(st));
    return SVLN_OK;
}

}  // extern "C"

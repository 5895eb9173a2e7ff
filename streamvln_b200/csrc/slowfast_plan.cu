// Device-side planning of the slow-memory eviction, batched over every stream that slides its window this turn.
//
// No reference counterpart: the released StreamVLN drops the whole KV cache every num_frames steps
// (streamvln/model/stream_video_vln.py:473-479, streamvln/streamvln_agent.py:192-198).  This is north_star subsystem
// (2) "window eviction, in-place compaction with memory-token splice, RoPE position remap" glued to subsystem (1)
// "keep-latest-per-voxel dedup", to the spec in DESIGN.md section 4 (its normative CPU statement is the `_evict` routine
// of the slowfast checker under the repo's test infrastructure).
//
// Per stream (job): candidates = [ M slow-memory tokens | the 196 image tokens of the evicted turn ]; svln_keep_latest
// marks the survivors (one dedup segment per job); then ONE CTA per job turns that mask into
//   * the source-row map of svln_kv_compact:  [prefix rows | survivors, oldest first, at most mem_cap (oldest dropped) |
//     every row behind the evicted turn], first moved row, new cache length;
//   * the stream's new memory key / time arrays (stable order).
// Nothing returns to the host in between: round 1 did nonzero() + .cpu() + a Python-built map per stream per turn.
// HBM-bound integer work on a few KB per job; what matters is the launch count (3 per turn instead of ~6 per stream).
#include "svln_common.cuh"

namespace svln {

constexpr int SFJ = SVLN_SF_JOB_WORDS;
constexpr int PLAN_THREADS = 1024;

__global__ void __launch_bounds__(256)
sf_gather_candidates_kernel(const int32_t* __restrict__ jobs, const int64_t* __restrict__ mem_key,
                            const int32_t* __restrict__ mem_t, int mem_cap, const int64_t* __restrict__ turn_key,
                            int n_slots, int tpf, int64_t* __restrict__ cand_key, int32_t* __restrict__ cand_t) {
    const int32_t* J = jobs + (int64_t)blockIdx.y * SFJ;
    const int env = J[0], M = J[2], t = J[7], slot = J[8];
    const int64_t off = J[9];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M + tpf) return;
    if (i < M) {
        cand_key[off + i] = mem_key[(int64_t)env * mem_cap + i];
        cand_t[off + i] = mem_t[(int64_t)env * mem_cap + i];
    } else {
        cand_key[off + i] = turn_key[((int64_t)env * n_slots + slot) * tpf + (i - M)];
        cand_t[off + i] = t;
    }
}

__global__ void __launch_bounds__(PLAN_THREADS)
sf_evict_plan_kernel(const int32_t* __restrict__ jobs, int n_jobs, const uint8_t* __restrict__ mask,
                     const int64_t* __restrict__ cand_key, const int32_t* __restrict__ cand_t, int mem_cap, int tpf,
                     int max_len, int64_t* __restrict__ mem_key, int32_t* __restrict__ mem_t,
                     int32_t* __restrict__ src_row, int32_t* __restrict__ plan_out) {
    __shared__ int s_warp[32];
    __shared__ int s_total, s_first;
    const int job = blockIdx.x;
    const int32_t* J = jobs + (int64_t)job * SFJ;
    const int env = J[0], P = J[1], M = J[2], rs = J[3], io = J[4], nr = J[5], L = J[6];
    const int64_t off = J[9];
    const int n = M + tpf;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int32_t* src = src_row + (int64_t)job * max_len;

    // pass 1: how many candidates survive the dedup
    int cnt = 0;
    for (int i = tid; i < n; i += PLAN_THREADS) cnt += mask[off + i] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_warp[warp] = cnt;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_total = v;
    }
    __syncthreads();
    const int K = s_total;
    const int drop = K > mem_cap ? K - mem_cap : 0;   // bounded memory: the OLDEST survivors go first
    const int Kc = K - drop;
    const int rest = max(L - (rs + nr), 0);
    const int new_len = P + Kc + rest;
    if (tid == 0) s_first = new_len;
    __syncthreads();

    // pass 2: stable ranks (ballot + popcount per warp, shuffle scan over the 32 warp totals), 1024 candidates a time
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += PLAN_THREADS) {
        const int i = i0 + tid;
        const bool keep = i < n && mask[off + i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int before = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const int v = s_warp[lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            s_warp[lane] = incl - v;                  // exclusive offset of each warp
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        if (keep) {
            const int r = base + s_warp[warp] + before;
            if (r >= drop) {
                const int q = r - drop;
                const int d = P + q;
                const int s = i < M ? P + i : rs + io + (i - M);
                src[d] = s;
                mem_key[(int64_t)env * mem_cap + q] = cand_key[off + i];
                mem_t[(int64_t)env * mem_cap + q] = cand_t[off + i];
                if (s != d) atomicMin(&s_first, d);
            }
        }
        base += s_total;
        __syncthreads();
    }
    for (int d = tid; d < P; d += PLAN_THREADS) src[d] = d;
    for (int r = tid; r < rest; r += PLAN_THREADS) {
        const int d = P + Kc + r, s = rs + nr + r;
        src[d] = s;
        if (s != d) atomicMin(&s_first, d);
    }
    __syncthreads();
    if (tid == 0) {
        plan_out[0 * n_jobs + job] = s_first;
        plan_out[1 * n_jobs + job] = new_len;
        plan_out[2 * n_jobs + job] = Kc;
        plan_out[3 * n_jobs + job] = K;
    }
}

}  // namespace svln

using namespace svln;

extern "C" {

int svln_sf_gather_candidates(const int32_t* jobs, int n_jobs, const int64_t* mem_key, const int32_t* mem_t,
                              int mem_cap, const int64_t* turn_key, int n_slots, int tokens_per_frame,
                              int64_t* cand_key, int32_t* cand_t, svln_stream_t stream) {
    SVLN_REQUIRE(n_jobs >= 0 && mem_cap > 0 && n_slots > 0 && tokens_per_frame > 0, SVLN_ERR_INVALID,
                 "svln_sf_gather_candidates: bad sizes");
    if (n_jobs == 0) return SVLN_OK;
    SVLN_REQUIRE(jobs && mem_key && mem_t && turn_key && cand_key && cand_t, SVLN_ERR_INVALID,
                 "svln_sf_gather_candidates: null pointer");
    SVLN_REQUIRE(n_jobs <= 65535, SVLN_ERR_RANGE, "svln_sf_gather_candidates: too many jobs");
    dim3 grid((mem_cap + tokens_per_frame + 255) / 256, n_jobs);
    sf_gather_candidates_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(jobs, mem_key, mem_t, mem_cap, turn_key, n_slots,
                                                                       tokens_per_frame, cand_key, cand_t);
    return check_launch("sf_gather_candidates_kernel");
}

int svln_sf_evict_plan(const int32_t* jobs, int n_jobs, const uint8_t* mask, const int64_t* cand_key,
                       const int32_t* cand_t, int mem_cap, int tokens_per_frame, int max_len, int64_t* mem_key,
                       int32_t* mem_t, int32_t* src_row, int32_t* plan_out, svln_stream_t stream) {
    SVLN_REQUIRE(n_jobs >= 0 && mem_cap > 0 && tokens_per_frame > 0 && max_len > 0, SVLN_ERR_INVALID,
                 "svln_sf_evict_plan: bad sizes");
    if (n_jobs == 0) return SVLN_OK;
    SVLN_REQUIRE(jobs && mask && cand_key && cand_t && mem_key && mem_t && src_row && plan_out, SVLN_ERR_INVALID,
                 "svln_sf_evict_plan: null pointer");
    sf_evict_plan_kernel<<<n_jobs, PLAN_THREADS, 0, (cudaStream_t)stream>>>(jobs, n_jobs, mask, cand_key, cand_t, mem_cap,
                                                                            tokens_per_frame, max_len, mem_key, mem_t,
                                                                            src_row, plan_out);
    return check_launch("sf_evict_plan_kernel");
}

}  // extern "C"

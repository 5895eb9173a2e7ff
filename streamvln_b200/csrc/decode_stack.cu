// Fused decode stack: ONE persistent kernel runs a whole greedy decode step (embedding row -> every Qwen2
// layer -> final norm -> lm_head -> argmax) of one stream.
//
// What it replaces: HF's per-token forward inside GenerationMixin's greedy loop (transformers 4.45.1; reached
// from StreamVLNForCausalLM.generate, stream_video_vln.py:340 / :402).  At batch 1 that step is nothing but
// streaming the 15.2 GB of bf16 weights once per token; the launch-per-op
// form is 10 kernels per layer (cuBLAS GEMV + small fused ops), ~290 launches per token even inside a CUDA graph.
//
// Design (B200: 148 SMs, 6.5 TB/s measured HBM copy peak, 126 MB L2):
//   * grid = one CTA per SM (cooperative launch), 8 warps; phases separated by a grid barrier (one L2 atomic
//     + spin per CTA).  5 barriers per layer: QKV | attention | O | gate/up | down.
//   * every linear is a weight-streaming GEMV on the legacy tensor path: W rows are the 16-row A operand of
//     mma.m16n8k16 and the activation vector is a column of the B operand.  The k index inside an MMA is a free
//     permutation as long as A and B agree, so a lane copies 16 contiguous bytes of a W row into a slot of shared
//     memory that only it reads back (cp.async: no barrier, no unpack/FMA work) and pairs it with the matching
//     16 bytes of x.  Warps interleave over k (step S -> warp S % 8): a CTA sweeps its rows front to back.
//     Up to 4 row blocks advance through k together (independent MMAs), 96 KB of copies in flight per SM.
//   * rows are dealt to CTAs through a per-phase row map: QKV rows come in RoPE pairs (d, d+64) so the epilogue
//     applies bias + RoPE and writes Q / the new K,V rows of the arena directly; gate and up rows come in pairs
//     so the epilogue applies SwiGLU; the residual adds and RMSNorms are recomputed by every CTA in the preamble
//     of the phase that needs them (7 KB from L2) instead of being a phase of their own.
//   * the barrier is split: a CTA ARRIVES (fence + atomic) as soon as its outputs are stored, then builds the next
//     row map and issues the next phase's first weight copies (or the attention item's K/V tiles), and only then
//     WAITS - so HBM latency overlaps the barrier and the preamble.
//   * attention is the split-KV work item of attn_decode.cu (attn_decode_body.cuh) run by warps 0-3; the CTA
//     that completes the last split of a (stream, kv head) merges the partials (fixed order: deterministic).
// Measured dead ends, kept out of the code: a register ring of LDG.128 (the six per-warp scoreboards serialise
// it: 1.3 TB/s); a 20-deep ring (30 % slower than 8-12 steps); bulk or per-line L2 prefetch of the next phase
// and page-touching one phase ahead (no gain: the main loops are HBM-bound, gate/up streams at 6.9 TB/s);
// fencing after the weight prologue (the fence waits for the copies: +3-6 us per barrier); 512-byte coalesced
// requests instead of 8 rows x 64 B (35 % slower in DRAM); no L2 evict_first hint on the weight copies (12 % slower);
// an 8-warp attention item with the O projection's weights streaming behind it (5 % slower); prefetch.global.L2
// of the attention tiles one phase early (6 % slower: every use of the L2 prefetch instructions cost more than it hid);
// a barrier whose last arriver publishes a generation number to eight separate flag lines that the CTAs poll instead
// of the arrival counter (3-6 % slower: one more L2 round trip per barrier than it saves in poll contention); dealing
// rows round-robin (CTA c owns rows c, c + G, ...: the grid sweeps one contiguous band) instead of in contiguous
// slices (no difference); 12 warps per CTA with a 16-slot ring (5 % slower: 168 registers per thread, more spills);
// round 2: cp.async.bulk.prefetch.L2 of the bytes right behind the ring at every phase start (64 / 160 / 320 KB per CTA:
// 4 / 6 / 9 % slower) and of the CTA's O slice + 0 / 256 / 512 KB of its gate/up slice from the idle warps during the
// attention phase (1.5 / 3 / 8 % slower) - profiles/r2_decode_stack_experiments.md.
// Round 2, the default feed (FEED = 1, descriptor field `packed`): the weights are re-laid by svln_decode_stack_pack as
// one contiguous stream per CTA in consumption order and copied into a 4 x 32 KB shared-memory ring by ONE producer
// thread per SM (cp.async.bulk + mbarriers) that runs ahead across phase boundaries; the kernel launches three
// warpgroups, the producer's gives its registers back and the compute warps run with 240 (setmaxnreg); residual adds sit
// in the O / down epilogues, x of those two projections arrives by a bulk copy, everything a phase needs that does
// not depend on the step (norm weights, bias, RoPE factors) is fetched before its barrier is waited for; the grid
// barrier is a release reduction + relaxed polls; nothing is called (calls cost caller-saved spills): no stack frame.
// 2.64 ms per token at 7B (0.82-0.83 of the measured HBM peak)
// against 3.3-3.4 for the LDGSTS feed (FEED = 0, kept: descriptor without `packed`).  See the "packed weight stream"
// block below and profiles/r2_decode_stack_experiments.md.
// Rounding points are those of the unfused path (fused_ops.cu, kv_window.cu rope8, gemv.cu), i.e. HF's bf16
// modules; accumulation order differs from cuBLAS, which the parity tolerance (1e-2 on logits) covers.
#include <stdlib.h>

#include <vector>

#include "attn_decode_body.cuh"

namespace svln {

constexpr int MK_WARPS = 8;
constexpr int MK_THREADS = MK_WARPS * 32;
constexpr int MK_CAP = 304;          // rows per group (multiple of 16; gate/up pairs of one CTA at 7B: 2 x 128)
// The kernel is written for NB <= 2 streams, but only NB = 1 is instantiated: in the NB = 2 build ptxas 12.9 selects,
// for some of the weight copies, an LDGSTS form ([R+UR0], desc[UR1]) that does not reference the cache-policy
// descriptor registers it has just loaded and raises "Warp Illegal Instruction" on the B200 (pinned down with
// cuda-gdb; laundering the address, a 64-bit shared operand and creating the policy locally did not avoid it).
constexpr int MK_MAX_B = 2;          // 2 streams only with the packed feed (the LDGSTS feed's two-stream build hits the ptxas fault below)
constexpr int MK_MAX_SPLITS = 32;
constexpr int MK_XPAD = 32;          // elements between the x rows of two streams: 64 B, spreads the banks
constexpr int MK_MISC = 4096;        // bytes of misc shared memory (reductions, flags, combine weights)
constexpr unsigned MK_SPIN_LIMIT = 1u << 22;
// CTA-wide barrier of the MK_THREADS compute threads: a named barrier, so that the packed-stream variant can run a
// ninth (producer) warp that never joins it.  With exactly MK_THREADS threads in the CTA it is __syncthreads().
#define MK_SYNC() asm volatile("bar.sync 2, %0;" ::"n"(MK_THREADS) : "memory")

struct StackParams {
    svln_kv_arena arena;
    int n_layer, D, I, Hq, Hkv, V, G;
    float eps, scale_log2;
    const long long* layer_ptrs;  // [n_layer][8]: wqkv, bqkv, wo, wgu, wd, ln1, ln2, unused
    const __nv_bfloat16 *embed, *final_norm, *lm_head, *cos_t, *sin_t;
    const int64_t* tok_in;
    int64_t* tok_out;
    float* logits;                // [B][V] or nullptr
    const int32_t *seq_env, *kv_len;
    __nv_bfloat16 *hbuf, *qrot, *attn_out, *o, *act, *pending;
    float *part_o, *part_m, *part_l, *best_v;
    int* best_i;
    unsigned* sync;               // [0] grid barrier, [16 + b*Hkv + kvh] split arrival counters (zeroed per launch)
    int n_splits;
    int xs_stride;                // elements between the x rows of two streams in shared memory
    int off_part, off_rowmap, off_ring, off_misc;
    int dbg;                      // SVLN_STACK_DBG (debug) bit 0: skip preambles, 1: skip GEMV main loop, 2: skip epilogues
    int stop_after;               // SVLN_STACK_STOP (debug): return after this many phases, 0 = run everything
    long long* trace;             // SVLN_STACK_TRACE: clock64 stamps of CTA 0 [layer][phase 0..5][point 0..7], else nullptr
    const unsigned char* packed;  // packed weight stream (svln_decode_stack_pack) or nullptr
};

__device__ __forceinline__ unsigned long long mk_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// every CTA, middle layer only: [cta][16] nanosecond stamps (globaltimer is common to all SMs) after the trace block
#define MK_GT(li, slot_)                                                                              \
    do {                                                                                              \
        if (p.trace && (li) == p.n_layer / 2 && threadIdx.x == 0)                                      \
            p.trace[(int64_t)(p.n_layer + 1) * 6 * 8 + 8 + blockIdx.x * 16 + (slot_)] = (long long)mk_globaltimer(); \
    } while (0)
#define MK_TRACE(li, ph, pt)                                                                      \
    do {                                                                                          \
        if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[((li) * 6 + (ph)) * 8 + (pt)] = clock64(); \
    } while (0)

__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// weight stream: read-only path, no L1 allocation, first to leave L2 (every byte is used once per step)
__device__ __forceinline__ int4 ld_weight(const int4* p, uint64_t pol) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int4 lds128u(uint32_t addr) {
    int4 r;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
// 16-byte async copy global -> shared through L2 only; src_bytes = 0 zero-fills the destination.  The evict_first
// hint matters: without it the weight stream pushes the activations, K/V partials and barrier lines out of L2
// and the whole step gets 12 % slower (measured).
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void* src, int src_bytes, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;"
                 ::"r"(dst), "l"(src), "r"(src_bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ float ld_cg_f32(const float* p) {
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// Grid barrier in two halves.  grid_arrive publishes this CTA's writes (fence + one atomic); grid_wait spins until
// every CTA has arrived.  Work that needs nothing from the other CTAs - building the next row map, issuing the
// next phase's first weight loads - goes BETWEEN the two, and must come AFTER grid_arrive: a __threadfence() also
// waits for the thread's outstanding cp.async loads, so fencing after the weight prologue added 3-6 us of HBM
// latency to every barrier (measured).  The wait side needs no fence: ld.acquire + the CTA barrier order the
// later reads.  `epoch` counts arrivals expected so far; the counter is zeroed by a memset node before every launch.
// A CTA that waits absurdly long (a bug, or a grid that is not co-resident) raises the abort flag sync[1] = barrier
// index instead of hanging the GPU: every later wait falls through, the kernel ends with garbage, and CTA 0 reports
// token -1 so the caller fails loudly (sync[2], sync[3] = CTA, count seen).
__device__ __forceinline__ void grid_arrive(unsigned* sync, unsigned& epoch, int G) {
    MK_SYNC();
    epoch += (unsigned)G;
    if (threadIdx.x == 0) {
        // release at gpu scope (cumulative over the CTA barrier above): this CTA's stores are performed at L2 before the
        // count moves.  __threadfence() here is MEMBAR.SC.GPU + CCTL.IVALL - a sequentially consistent fence that also
        // throws the SM's L1 (spills, stack, read-only tables) away at every barrier.
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sync) : "memory");
    }
}
__device__ __forceinline__ void grid_wait(unsigned* sync, unsigned epoch, int G) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        for (;;) {
            unsigned v;
            // relaxed poll: ld.acquire is LDG + CCTL.IVALL, an L1 flush per poll.  Nothing produced inside the launch is
            // ever read through L1 (ld_cg_* helpers, cp.async.cg, bulk copies; tests/test_sharding_cpu.py guards it), so
            // the data the count stands for is read from L2, where the producers' release put it.
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sync) : "memory");
            if (v >= epoch) break;
            if ((++spins & 1023u) == 0) {
                unsigned ab;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(sync + 1) : "memory");
                if (ab != 0) break;
                if (spins > MK_SPIN_LIMIT) {
                    if (atomicCAS(sync + 1, 0u, epoch / (unsigned)G) == 0u) { sync[2] = blockIdx.x; sync[3] = v; }
                    break;
                }
            }
        }
    }
    MK_SYNC();
}

// ---- weight-streaming GEMV core -------------------------------------------------------------------------------
// Partial sums of rows rowmap[0..nrows) of W (row-major [*, K]) against the NB vectors in xs.  Rows are taken in
// nblk = ceil(nrows / 16) balanced blocks (the 16 A rows of an MMA); warp w covers k32-steps S = w, w + 8, ...;
// partials go to part[(w * MK_CAP + r) * NB + n].
// The W fragments travel global -> shared with cp.async into slots private to each lane (a register ring is
// serialised by the six per-warp scoreboards: a wait on the oldest load also waits for the newest one sharing its
// scoreboard; measured 1.3 TB/s).  cp.async groups give a true rolling pipeline, and no barrier is needed because
// every lane reads back only what it copied itself.  gemv_start fills the ring - it needs nothing but the weights,
// so the caller runs it BEFORE waiting for the previous phase (grid barrier) and building x; gemv_finish streams
// the rest.
constexpr int MK_RING_SLOTS = 24;                  // 4 KB ring slots (one lane-private int4 per thread each): 96 KB
                                                   // (measured, time relative to the unfused path: 16 slots 0.98, 24: 0.92-0.95, 32: 1.03)
constexpr int MK_MAX_SB = 4;
struct GemvCursor {
    int it, st, sup;                  // next step to issue: flat index, step within the super-block, super-block
    const int4* p[MK_MAX_SB][2];      // where the lane's next 16 B of rows g, g + 8 of each block come from
    int sz[MK_MAX_SB][2];             // 16, or 0 for a row that does not exist (the copy zero-fills its slot)
};
struct GemvGeom {
    const int4* Wv;
    const int* rowmap;      // [nrows] rows of this group, then (at MK_CAP) the balanced block starts [nblk + 1]
    int K8, nrows, nblk, nsup, nst, total, g;
    uint32_t slot0;
};
constexpr int MK_STEP = MK_WARPS * 4;              // int4 between two consecutive steps of one warp
constexpr uint32_t MK_SLOT = MK_THREADS * 16;      // bytes between two ring slots of one lane

// A super-block is SB consecutive row blocks (of <= 16 rows, balanced) that advance through k together: one
// iteration = one k32-step of SB blocks = 2 SB cp.async per lane, 1 x fragment, 2 SB independent MMAs.  The ring
// holds DEPTH = 12 / SB iterations.  (With one block per iteration the loop was bound by the per-warp dependent
// latency - wait, LDS, two chained MMAs: ~340 cycles per KB-sized step whatever the phase - not by HBM.)
template <int SB>
__device__ __forceinline__ GemvGeom gemv_geom(const __nv_bfloat16* W, int K, const int* rowmap, int nrows,
                                              unsigned char* ring) {
    GemvGeom G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsteps = K >> 5;
    G.g = lane >> 2;
    G.nst = (nsteps - warp + MK_WARPS - 1) / MK_WARPS;
    G.nrows = nrows;
    G.nblk = (nrows + 15) >> 4;
    G.nsup = (G.nblk + SB - 1) / SB;
    G.total = G.nsup * G.nst;
    G.K8 = K >> 3;
    G.Wv = reinterpret_cast<const int4*>(W) + warp * 4 + (lane & 3);
    G.rowmap = rowmap;
    G.slot0 = smem_u32(ring) + threadIdx.x * 16;
    return G;
}
template <int SB>
__device__ __forceinline__ void gemv_set_super(GemvCursor& c, const GemvGeom& G) {
#pragma unroll
    for (int j = 0; j < SB; ++j) {
        const int blk = c.sup * SB + j;
        c.p[j][0] = c.p[j][1] = G.Wv;
        c.sz[j][0] = c.sz[j][1] = 0;
        if (blk < G.nblk) {
            const int start = G.rowmap[MK_CAP + blk], size = G.rowmap[MK_CAP + blk + 1] - start;
            if (G.g < size) { c.p[j][0] = G.Wv + (int64_t)G.rowmap[start + G.g] * G.K8; c.sz[j][0] = 16; }
            if (G.g + 8 < size) { c.p[j][1] = G.Wv + (int64_t)G.rowmap[start + G.g + 8] * G.K8; c.sz[j][1] = 16; }
        }
    }
}
template <int SB>
__device__ __forceinline__ void gemv_issue(GemvCursor& c, const GemvGeom& G, int q, uint64_t pol) {
    if (c.it < G.total) {
        const int nsb = min(SB, G.nblk - c.sup * SB);
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            if (j < nsb) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    cp_async16_hint(G.slot0 + ((q * SB + j) * 2 + h) * MK_SLOT, c.p[j][h], c.sz[j][h], pol);
                    c.p[j][h] += MK_STEP;
                }
            }
        }
        ++c.it;
        if (++c.st == G.nst) {
            c.st = 0;
            if (++c.sup < G.nsup) gemv_set_super<SB>(c, G);
        }
    }
    cp_async_commit();   // always: keeps the group count in step with q
}

template <int SB>
__device__ __noinline__ GemvCursor gemv_start(const __nv_bfloat16* __restrict__ W, int K, const int* rowmap, int nrows,
                                              unsigned char* ring, uint64_t pol) {
    constexpr int DEPTH = MK_RING_SLOTS / (2 * SB);
    const GemvGeom G = gemv_geom<SB>(W, K, rowmap, nrows, ring);
    GemvCursor c;
    c.it = c.st = c.sup = 0;
#pragma unroll
    for (int j = 0; j < MK_MAX_SB; ++j) {
        c.p[j][0] = c.p[j][1] = nullptr;
        c.sz[j][0] = c.sz[j][1] = 0;
    }
    if (G.total > 0) gemv_set_super<SB>(c, G);
#pragma unroll
    for (int q = 0; q < DEPTH; ++q) gemv_issue<SB>(c, G, q, pol);
    return c;
}

template <int NB, int SB>
__device__ __noinline__ void gemv_finish(GemvCursor c, const __nv_bfloat16* __restrict__ W, int K, const int* rowmap,
                                         int nrows, const __nv_bfloat16* xs, int xs_stride, float* part,
                                         unsigned char* ring, uint64_t pol) {
    constexpr int DEPTH = MK_RING_SLOTS / (2 * SB);
    const GemvGeom G = gemv_geom<SB>(W, K, rowmap, nrows, ring);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t xv = smem_u32(xs + (g < NB ? g : 0) * xs_stride) + (warp * 4 + t) * 16;
    float acc[SB][2][4];
#pragma unroll
    for (int j = 0; j < SB; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][0][e] = acc[j][1][e] = 0.f;
    int c_st = 0, c_sup = 0;
    for (int it = 0; it < G.total; it += DEPTH) {
#pragma unroll
        for (int q = 0; q < DEPTH; ++q) {
            cp_async_wait<DEPTH - 1>();
            if (it + q < G.total) {
                const int nsb = min(SB, G.nblk - c_sup * SB);
                const int4 x = lds128u(xv + c_st * (MK_STEP * 16));
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    if (j < nsb) {
                        const int4 w0 = lds128u(G.slot0 + ((q * SB + j) * 2) * MK_SLOT);
                        const int4 w1 = lds128u(G.slot0 + ((q * SB + j) * 2 + 1) * MK_SLOT);
                        const uint32_t a0[4] = {(uint32_t)w0.x, (uint32_t)w1.x, (uint32_t)w0.y, (uint32_t)w1.y};
                        const uint32_t a1[4] = {(uint32_t)w0.z, (uint32_t)w1.z, (uint32_t)w0.w, (uint32_t)w1.w};
                        mma_bf16(acc[j][0], a0, (uint32_t)x.x, (uint32_t)x.y);
                        mma_bf16(acc[j][1], a1, (uint32_t)x.z, (uint32_t)x.w);
                    }
                }
                if (++c_st == G.nst) {
#pragma unroll
                    for (int j = 0; j < SB; ++j) {
                        if (j < nsb) {
                            const int blk = c_sup * SB + j;
                            const int start = rowmap[MK_CAP + blk], size = rowmap[MK_CAP + blk + 1] - start;
                            float* dst = part + ((int64_t)warp * MK_CAP + start + g) * NB;
                            if (g < size) {
                                if (2 * t < NB) dst[2 * t] = acc[j][0][0] + acc[j][1][0];
                                if (2 * t + 1 < NB) dst[2 * t + 1] = acc[j][0][1] + acc[j][1][1];
                            }
                            if (g + 8 < size) {
                                if (2 * t < NB) dst[8 * NB + 2 * t] = acc[j][0][2] + acc[j][1][2];
                                if (2 * t + 1 < NB) dst[8 * NB + 2 * t + 1] = acc[j][0][3] + acc[j][1][3];
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[j][0][e] = acc[j][1][e] = 0.f;
                        }
                    }
                    c_st = 0;
                    ++c_sup;
                }
            }
            gemv_issue<SB>(c, G, q, pol);
        }
    }
    cp_async_wait<0>();
}

// One GEMV phase of this CTA: `n_units` units of RPU rows each (row = rowfn(unit, which)); epi(unit, n, sums[RPU])
// consumes the finished dot products.  `mid` runs exactly once, after the first weights are in flight and before
// they are consumed: it waits for the previous phase (grid barrier) and builds x in shared memory.
template <int NB, int RPU, class RowFn, class MidFn, class EpiFn>
__device__ __forceinline__ void gemv_phase(const StackParams& p, unsigned char* smem, const __nv_bfloat16* W, int K,
                                           int n_units, RowFn rowfn, MidFn mid, EpiFn epi, uint64_t pol,
                                           int tr_li, int tr_ph) {
    const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(smem);
    float* part = reinterpret_cast<float*>(smem + p.off_part);
    int* rowmap = reinterpret_cast<int*>(smem + p.off_rowmap);
    unsigned char* ring = smem + p.off_ring;
    constexpr int UPG = MK_CAP / RPU;
    const int n_active = min(MK_WARPS, K >> 5);
    const int tid = threadIdx.x;
    MK_TRACE(tr_li, tr_ph, 0);
    if (n_units <= 0) { mid(); return; }
    for (int u0 = 0; u0 < n_units; u0 += UPG) {
        const int nu = min(UPG, n_units - u0);
        for (int j = tid; j < nu * RPU; j += MK_THREADS) rowmap[j] = rowfn(u0 + j / RPU, j % RPU);
        {
            const int nr = nu * RPU, nb = (nr + 15) >> 4;
            if (tid <= nb) rowmap[MK_CAP + tid] = (tid * nr) / nb;   // balanced blocks of <= 16 rows
        }
        MK_SYNC();
        const int nrows = nu * RPU, nblk = (nrows + 15) >> 4;
        const int sb = nblk >= 4 ? 4 : (nblk >= 2 ? 2 : 1);     // blocks advancing together (ILP) vs ring depth
        GemvCursor cur;
        if (sb == 4) cur = gemv_start<4>(W, K, rowmap, nrows, ring, pol);
        else if (sb == 2) cur = gemv_start<2>(W, K, rowmap, nrows, ring, pol);
        else cur = gemv_start<1>(W, K, rowmap, nrows, ring, pol);
        MK_TRACE(tr_li, tr_ph, 2);
        if (u0 == 0) {
            mid();
        }
        MK_TRACE(tr_li, tr_ph, 3);
        if (p.dbg & 2) cp_async_wait<0>();
        else if (sb == 4) gemv_finish<NB, 4>(cur, W, K, rowmap, nrows, xs, p.xs_stride, part, ring, pol);
        else if (sb == 2) gemv_finish<NB, 2>(cur, W, K, rowmap, nrows, xs, p.xs_stride, part, ring, pol);
        else gemv_finish<NB, 1>(cur, W, K, rowmap, nrows, xs, p.xs_stride, part, ring, pol);
        MK_SYNC();
        MK_TRACE(tr_li, tr_ph, 4);
        MK_GT(tr_li, tr_ph * 3 + 1);
        for (int i = tid; i < nu * NB && !(p.dbg & 4); i += MK_THREADS) {
            const int j = i / NB, n = i % NB;
            float s[RPU];
#pragma unroll
            for (int q = 0; q < RPU; ++q) {
                float a = 0.f;
                for (int w = 0; w < n_active; ++w) a += part[((int64_t)w * MK_CAP + j * RPU + q) * NB + n];
                s[q] = a;
            }
            epi(u0 + j, n, s);
        }
        MK_SYNC();
    }
    MK_TRACE(tr_li, tr_ph, 5);
    MK_GT(tr_li, tr_ph * 3 + 2);
}

// The RMSNorm weights of a phase do not depend on this step: every thread fetches its share BEFORE the grid barrier is
// waited for (they come from HBM - 14 GB of weights have passed through L2 since their last use).
constexpr int MK_NORM_Q = 2;  // D <= 256 * 2 * 8 = 4096
struct NormW { int4 w[MK_NORM_Q]; };
__device__ __forceinline__ NormW norm_w_load(const StackParams& p, const __nv_bfloat16* lnw) {
    NormW r;
    const int D8 = p.D >> 3;
#pragma unroll
    for (int q = 0; q < MK_NORM_Q; ++q) {
        const int c = q * MK_THREADS + threadIdx.x;
        r.w[q] = c < D8 ? __ldg(reinterpret_cast<const int4*>(lnw) + c) : make_int4(0, 0, 0, 0);
    }
    return r;
}

// x = RMSNorm(src) * w for every stream into shared memory.  src is the residual stream [NB][D] (summed by the
// epilogues of the O and down projections), or the embedding rows of the step's tokens (layer 0).
template <int NB>
__device__ __forceinline__ void preamble_norm(const StackParams& p, unsigned char* smem, const __nv_bfloat16* src,
                                           bool src_embed, const NormW& lw) {
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);
    float* red = reinterpret_cast<float*>(smem + p.off_misc);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D8 = p.D >> 3;
    constexpr int MAXQ = MK_NORM_Q;
    int4 v[NB][MAXQ];
    long long* pt = (p.trace && blockIdx.x == 0 && tid == 0 && !src_embed) ? p.trace + ((int64_t)p.n_layer * 6 + 5) * 8 : nullptr;
    if (pt) pt[0] = clock64();
#pragma unroll
    for (int n = 0; n < NB; ++n) {
        const int4* s;
        if (src_embed) {
            int64_t tok = p.tok_in[n];
            tok = tok < 0 ? 0 : (tok >= p.V ? p.V - 1 : tok);
            s = reinterpret_cast<const int4*>(p.embed + tok * p.D);
        } else {
            s = reinterpret_cast<const int4*>(src + (int64_t)n * p.D);
        }
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int c = q * MK_THREADS + tid;
            if (c < D8) v[n][q] = src_embed ? __ldg(s + c) : ld_cg_v4(s + c);
        }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int c = q * MK_THREADS + tid;
            if (c < D8) {
                const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[n][q]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a0 = bf16_bits_to_f32(u[k] & 0xFFFFu), a1 = bf16_bits_to_f32(u[k] >> 16);
                    ss += a0 * a0 + a1 * a1;
                }
            }
        }
        ss = warp_sum(ss);
        if (lane == 0) red[n * MK_WARPS + warp] = ss;
    }
    if (pt) pt[1] = clock64();
    MK_SYNC();
    if (pt) pt[2] = clock64();
#pragma unroll
    for (int n = 0; n < NB; ++n) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < MK_WARPS; ++i) tot += red[n * MK_WARPS + i];
        const float rstd = rsqrtf(tot * (1.0f / (float)p.D) + p.eps);
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int c = q * MK_THREADS + tid;
            if (c < D8) {
                const uint32_t* wu = reinterpret_cast<const uint32_t*>(&lw.w[q]);
                const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[n][q]);
                int4 o;
                uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float n0 = round_bf16(bf16_bits_to_f32(u[k] & 0xFFFFu) * rstd);
                    const float n1 = round_bf16(bf16_bits_to_f32(u[k] >> 16) * rstd);
                    const float y0 = bf16_bits_to_f32(wu[k] & 0xFFFFu) * n0;
                    const float y1 = bf16_bits_to_f32(wu[k] >> 16) * n1;
                    ou[k] = f32_to_bf16_bits(y0) | (f32_to_bf16_bits(y1) << 16);
                }
                reinterpret_cast<int4*>(xs + (int64_t)n * p.xs_stride)[c] = o;
            }
        }
    }
    if (pt) pt[3] = clock64();
    MK_SYNC();
    if (pt) pt[4] = clock64();
}

// x = src[NB][K] (written earlier in this launch by other CTAs) into shared memory, 8 loads in flight per thread
template <int NB>
__device__ __forceinline__ void preamble_copy(const StackParams& p, unsigned char* smem, const __nv_bfloat16* src, int K) {
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);
    const int K8 = K >> 3, total = NB * K8;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    for (int base = 0; base < total; base += MK_THREADS * 8) {
        int4 tmp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = base + u * MK_THREADS + threadIdx.x;
            if (c < total) tmp[u] = ld_cg_v4(s4 + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = base + u * MK_THREADS + threadIdx.x;
            if (c < total) {
                const int n = c / K8, cc = c - n * K8;
                reinterpret_cast<int4*>(xs + (int64_t)n * p.xs_stride)[cc] = tmp[u];
            }
        }
    }
    MK_SYNC();
}

// merge the split partials of (stream b, kv head kvh): warp <-> query row, lane <-> 4 dims; same arithmetic and
// split order as attn_decode_combine_kernel, with all loads of a row in flight together
__device__ __forceinline__ void combine_splits(const StackParams& p, int b, int kvh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = p.Hq / p.Hkv;
    const int64_t slot0 = ((int64_t)b * p.Hkv + kvh) * p.n_splits;
    for (int r = warp; r < g; r += MK_WARPS) {
        const float ms = lane < p.n_splits ? ld_cg_f(p.part_m + (slot0 + lane) * DROWS + r) : -INFINITY;
        const float m = warp_max(ms);
        const float w = (lane < p.n_splits && ms != -INFINITY) ? exp2f((ms - m) * p.scale_log2) : 0.f;
        float l = lane < p.n_splits ? ld_cg_f(p.part_l + (slot0 + lane) * DROWS + r) * w : 0.f;
        l = warp_sum(l);
        const float4* po = reinterpret_cast<const float4*>(p.part_o + (slot0 * DROWS + r) * HD) + lane;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < p.n_splits; s0 += 16) {
            float4 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (s0 + u < p.n_splits) v[u] = ld_cg_f4(po + (int64_t)(s0 + u) * (DROWS * HD / 4));
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const float ws = __shfl_sync(0xffffffffu, w, (s0 + u) & 31);
                if (s0 + u < p.n_splits && ws != 0.f) {   // empty splits hold stale data: skipped
                    acc.x += v[u].x * ws; acc.y += v[u].y * ws; acc.z += v[u].z * ws; acc.w += v[u].w * ws;
                }
            }
        }
        const float inv_ok = l > 0.f ? 1.f : 0.f;
        const float o0 = inv_ok != 0.f ? acc.x / l : 0.f, o1 = inv_ok != 0.f ? acc.y / l : 0.f;
        const float o2 = inv_ok != 0.f ? acc.z / l : 0.f, o3 = inv_ok != 0.f ? acc.w / l : 0.f;
        uint2 packed;
        packed.x = pack_bf16(o0, o1);
        packed.y = pack_bf16(o2, o3);
        *reinterpret_cast<uint2*>(p.attn_out + ((int64_t)b * p.Hq + kvh * g + r) * HD + lane * 4) = packed;
    }
}

template <int MODE, int STAGES = DSTAGES>
__device__ __forceinline__ void attn_item(const DecodeParams& dp, int split, int kvh, int b, unsigned char* smem, int tid) {
    attn_decode_body<1, MODE, STAGES, STAGES != DSTAGES>(dp, split, kvh, b, smem, tid);
}


// ---- packed weight stream (FEED = 1) ------------------------------------------------------------------------------
// svln_decode_stack_pack re-lays the weights of the step once: every CTA's share of every linear, in the order the CTA
// consumes it, as ONE contiguous byte stream per CTA (layer by layer: QKV, O, gate/up, down; then lm_head).  One
// thread of a third warpgroup walks that stream with 1-D bulk copies (cp.async.bulk, L2 evict-first) into a ring of
// TK_SLOTS x 32 KB stages guarded by full/empty mbarriers.  It knows nothing of phases or grid barriers, so it keeps streaming
// while the compute warps sit in an epilogue, a grid barrier, a preamble or the attention phase: every phase starts
// on a full ring (128 KB per SM, 19 MB chip-wide: three quarters of a CTA's O slice are resident before the attention
// phase ends), and the 3-6 k cycles of LDGSTS issue per phase (profiles/r2_decode_stack_experiments.md) are gone.  Measured
// first with scripts/tma_stream_probe.cu: one thread per SM sustains 6.1-6.3 TB/s with 16-32 KB copies (2-D tensor
// boxes of fewer than ~128 rows do not: 100-200 ns per box).
// Layout of one stage: [k32-step][row of the group][64 B]: lane (g, t) of the warp that owns a k32-step reads the
// 16 bytes t of rows 16 b + g and 16 b + g + 8 - 512 contiguous bytes per warp load, conflict-free without a swizzle.
// Rows of a CTA are taken in balanced groups of <= 32 rows (QKV, O, down: the 8 warps interleave over the k32-steps
// of a stage, as in the LDGSTS variant) or <= 256 rows (gate/up, lm_head: one k32-step fills a stage, warp w owns row
// blocks w and w + 8 and sweeps every k-step itself: no cross-warp sum).
#ifndef TK_XCOPY
#define TK_XCOPY 3            // bit 0: x of the O projection by bulk copy, bit 1: x of the down projection
#endif
constexpr int TK_SLOT_BYTES = 32768;
constexpr int TK_SLOTS = 4;
constexpr int TK_ATTN_STAGES = 2;                  // K/V tile stages of the in-kernel attention item: 64 KB
constexpr int TK_ATTN_SMEM = TK_ATTN_STAGES * 2 * TILE_BYTES;
constexpr int TK_WIDE_ROWS = 256;
constexpr int TK_NARROW_ROWS = 32;
constexpr int TK_BAR_OFF = 256;                    // full[TK_SLOTS], empty[TK_SLOTS] inside the misc block
constexpr int TK_MISC = 512;                       // the misc block of the packed-feed variant (MK_MISC otherwise); the final
                                                   // argmax reduction borrows the partial-sum buffer instead

struct TkDims { int n_layer, D, I, Hq, Hkv, V, G; };
struct TkPhase { int n_total, rpu, K, wide; };
// balanced contiguous ranges; 32-bit arithmetic (n * G < 2^31 is checked on the host: 64-bit divisions are subroutine
// calls on the GPU and these sit between a grid barrier and the first weight byte)
__host__ __device__ inline int tk_lo(int n, int c, int G) { return (int)((unsigned)n * (unsigned)c / (unsigned)G); }
// the c with tk_lo(n, c, G) <= u < tk_lo(n, c + 1, G)
__host__ __device__ inline int tk_owner(int u, int n, int G) {
    return (int)((((unsigned)u + 1u) * (unsigned)G + (unsigned)n - 1u) / (unsigned)n) - 1;
}
__host__ __device__ inline int tk_groups(int n_units, int rpu, int wide) {
    const int cap = (wide ? TK_WIDE_ROWS : TK_NARROW_ROWS) / rpu;
    return (n_units + cap - 1) / cap;
}
__host__ __device__ inline int tk_ks(int rows, int wide) {     // k32-steps per stage
    const int fit = TK_SLOT_BYTES / (rows * 64);
    (void)wide;
    return fit < 1 ? 1 : fit;
}
// phases: 0 QKV (units = RoPE pairs), 1 O, 2 gate/up (units = gate/up pairs), 3 down, 4 lm_head
__host__ __device__ inline TkPhase tk_phase(const TkDims& d, int ph) {
    switch (ph) {
        case 0: return {(d.Hq + 2 * d.Hkv) * 64, 2, d.D, 0};
        case 1: return {d.D, 1, d.Hq * 128, 0};
        case 2: return {d.I, 2, d.D, 1};
        case 3: return {d.D, 1, d.I, 0};
        default: return {d.V, 1, d.D, 1};
    }
}
// bytes of one layer that belong to CTAs 0..c-1
__host__ __device__ inline long long tk_layer_before(const TkDims& d, int c) {
    long long b = 0;
    for (int ph = 0; ph < 4; ++ph) {
        const TkPhase P = tk_phase(d, ph);
        b += (long long)tk_lo(P.n_total, c, d.G) * P.rpu * P.K * 2;
    }
    return b;
}
__host__ __device__ inline long long tk_cta_base(const TkDims& d, int c) {
    return (long long)d.n_layer * tk_layer_before(d, c) + (long long)tk_lo(d.V, c, d.G) * d.D * 2;
}

struct TkFeed { uint32_t ring, full, empty; unsigned slot, par; unsigned* sync; const int* gate; uint32_t xbar; };

__device__ __forceinline__ void tk_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tk_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tk_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tk_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// a wait that cannot hang the GPU: after ~2^20 polls (seconds) it raises the abort flag of the grid barrier.
// SLEEP_NS > 0 (the producer thread): back off between polls - a thread that spins on try_wait competes with the
// compute warps for the SM's MIO queue (the attention item got 45 % slower with a spinning producer).
template <int SLEEP_NS = 0>
__device__ __forceinline__ void tk_mbar_wait(uint32_t bar, uint32_t parity, unsigned* sync) {
    unsigned polls = 0;
    while (!tk_mbar_try(bar, parity)) {
        if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
        if ((++polls & 4095u) == 0) {
            unsigned ab;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(sync + 1) : "memory");
            if (ab != 0) return;
            if (polls > (1u << 20)) { atomicCAS(sync + 1, 0u, 0x7FFFu); return; }
        }
    }
}

// the producer thread: walks the CTA's packed stream stage by stage, in the order gemv_phase_tk consumes it
__device__ __forceinline__ void tk_producer(const StackParams& p, TkFeed f, uint64_t pol) {
    const TkDims d = {p.n_layer, p.D, p.I, p.Hq, p.Hkv, p.V, p.G};
    const int c = blockIdx.x;
    const unsigned char* src = p.packed + tk_cta_base(d, c);
    unsigned slot = 0, par = 1, filled = 0;      // par: parity of the empty barrier's phase to wait for (first lap: none)
    int phases_left = p.stop_after > 0 ? p.stop_after : 0x7FFFFFFF;
    int cur_li = 0;
    auto run = [&](int ph) {
        const TkPhase P = tk_phase(d, ph);
        const int u0 = tk_lo(P.n_total, c, d.G), nu = tk_lo(P.n_total, c + 1, d.G) - u0;
        if (nu <= 0) return;
        const int ng = tk_groups(nu, P.rpu, P.wide), KT = P.K >> 5;
        for (int gi = 0; gi < ng; ++gi) {
            const int rows = (tk_lo(nu, gi + 1, ng) - tk_lo(nu, gi, ng)) * P.rpu;
            const int KS = tk_ks(rows, P.wide);
            for (int s0 = 0; s0 < KT; s0 += KS) {
                const uint32_t bytes = (uint32_t)(min(KS, KT - s0) * rows * 64);
                long long* tkt = nullptr;
                if (p.trace && c == 0 && cur_li == p.n_layer / 2 && ph < 4 && gi == 0 && s0 / KS < 100)
                    tkt = p.trace + (int64_t)(p.n_layer + 1) * 6 * 8 + 8 + 160 * 16 + 2048 + (ph * 100 + s0 / KS) * 2;
                if (tkt) tkt[0] = clock64();
                if (filled >= TK_SLOTS) tk_mbar_wait<0>(f.empty + slot * 8, par, f.sync);
                else ++filled;
                if (tkt) tkt[1] = clock64();
                tk_mbar_expect(f.full + slot * 8, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(f.ring + slot * TK_SLOT_BYTES), "l"(src), "r"(bytes), "r"(f.full + slot * 8), "l"(pol) : "memory");
                src += bytes;
                if (++slot == TK_SLOTS) { slot = 0; par ^= 1u; }
            }
        }
    };
    for (int li = 0; li < p.n_layer; ++li) {
        cur_li = li;
        run(0); if (--phases_left == 0) return;
        if (--phases_left == 0) return;                    // attention: no weights
        if (p.dbg & 8) {                                   // experiment: hold the O slice back until this CTA's attention work is done
            const volatile int* gate = reinterpret_cast<const volatile int*>(f.gate);
            unsigned polls = 0;
            while (*gate < li + 1 && ++polls < (1u << 22)) __nanosleep(64);
        }
        run(1); if (--phases_left == 0) return;
        run(2); if (--phases_left == 0) return;
        run(3); if (--phases_left == 0) return;
    }
    run(4);
}

// One GEMV phase of this CTA fed from the packed stream; same contract as gemv_phase (mid: wait for the previous phase
// and build x; epi(unit, n, sums[RPU])).
template <int NB, int RPU, bool WIDE, class MidFn, class EpiFn>
__device__ __forceinline__ void gemv_phase_tk(const StackParams& p, unsigned char* smem, TkFeed& f, int K, int n_units,
                                              MidFn mid, EpiFn epi, int tr_li, int tr_ph) {
    const __nv_bfloat16* xs = reinterpret_cast<const __nv_bfloat16*>(smem);
    float* part = reinterpret_cast<float*>(smem + p.off_part);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int KW = WIDE ? 1 : MK_WARPS;               // warps that share a row
    // geometry of the first group before the grid barrier is waited for (integer divisions off the critical path)
    const int ng = n_units > 0 ? tk_groups(n_units, RPU, WIDE) : 0, KT = K >> 5;
    const int blk0 = WIDE ? warp : 0, blk1 = WIDE ? warp + MK_WARPS : 1;      // this warp's two row blocks
    int gu0 = 0, nu = ng > 0 ? tk_lo(n_units, 1, ng) : 0;
    int rows = nu * RPU, KS = rows > 0 ? tk_ks(rows, WIDE) : 1;
    const uint32_t xv = smem_u32(xs + (g < NB ? g : 0) * p.xs_stride) + t * 16;
    MK_TRACE(tr_li, tr_ph, 0);
    MK_TRACE(tr_li, tr_ph, 2);
    mid();
    MK_TRACE(tr_li, tr_ph, 3);
    for (int gi = 0; gi < ng; ++gi) {
        // rows past the group's end re-read its last row (their sums are never used)
        const uint32_t r00 = (uint32_t)min(blk0 * 16 + g, rows - 1) * 64, r01 = (uint32_t)min(blk0 * 16 + g + 8, rows - 1) * 64;
        const uint32_t r10 = (uint32_t)min(blk1 * 16 + g, rows - 1) * 64, r11 = (uint32_t)min(blk1 * 16 + g + 8, rows - 1) * 64;
        const bool has0 = blk0 * 16 < rows, has1 = blk1 * 16 < rows;
        const uint32_t step_bytes = (uint32_t)rows * 64;
        float acc[2][2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][0][e] = acc[j][1][e] = 0.f;
        for (int s0 = 0; s0 < KT; s0 += KS) {
            const int ksn = min(KS, KT - s0);
            long long* tkt = nullptr;
            if (p.trace && blockIdx.x == 0 && tid == 0 && tr_li == p.n_layer / 2 && gi == 0 && s0 / KS < 100)
                tkt = p.trace + (int64_t)(p.n_layer + 1) * 6 * 8 + 8 + 160 * 16 + (tr_ph * 100 + s0 / KS) * 4;
            if (tkt) tkt[0] = clock64();
            tk_mbar_wait(f.full + f.slot * 8, f.par, f.sync);
            if (tkt) tkt[1] = clock64();
            const uint32_t base = f.ring + f.slot * TK_SLOT_BYTES + t * 16;
            if (!(p.dbg & 2) && has0) {
                // narrow: warp w owns the k32-steps whose index in the whole sweep is w mod 8 (fixed summation order)
#pragma unroll 2
                for (int s = WIDE ? 0 : ((warp - s0) & (MK_WARPS - 1)); s < ksn; s += KW) {
                    const int4 x = lds128u(xv + (s0 + s) * 64);
                    const uint32_t sb = base + (uint32_t)s * step_bytes;
                    {
                        const int4 w0 = lds128u(sb + r00), w1 = lds128u(sb + r01);
                        const uint32_t a0[4] = {(uint32_t)w0.x, (uint32_t)w1.x, (uint32_t)w0.y, (uint32_t)w1.y};
                        const uint32_t a1[4] = {(uint32_t)w0.z, (uint32_t)w1.z, (uint32_t)w0.w, (uint32_t)w1.w};
                        mma_bf16(acc[0][0], a0, (uint32_t)x.x, (uint32_t)x.y);
                        mma_bf16(acc[0][1], a1, (uint32_t)x.z, (uint32_t)x.w);
                    }
                    if (has1) {
                        const int4 w0 = lds128u(sb + r10), w1 = lds128u(sb + r11);
                        const uint32_t a0[4] = {(uint32_t)w0.x, (uint32_t)w1.x, (uint32_t)w0.y, (uint32_t)w1.y};
                        const uint32_t a1[4] = {(uint32_t)w0.z, (uint32_t)w1.z, (uint32_t)w0.w, (uint32_t)w1.w};
                        mma_bf16(acc[1][0], a0, (uint32_t)x.x, (uint32_t)x.y);
                        mma_bf16(acc[1][1], a1, (uint32_t)x.z, (uint32_t)x.w);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tk_mbar_arrive(f.empty + f.slot * 8);
            if (tkt) tkt[2] = clock64();
            if (++f.slot == TK_SLOTS) { f.slot = 0; f.par ^= 1u; }
        }
        // partial sums: part[(kw * TK_WIDE_ROWS + row) * NB + n], kw = warp (narrow) or 0 (wide)
        {
            float* pw = part + (int64_t)(WIDE ? 0 : warp) * TK_WIDE_ROWS * NB;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r0 = (j == 0 ? blk0 : blk1) * 16 + g;
                if (WIDE && !(j == 0 ? has0 : has1)) continue;
                if (r0 < rows) {
                    if (2 * t < NB) pw[r0 * NB + 2 * t] = acc[j][0][0] + acc[j][1][0];
                    if (2 * t + 1 < NB) pw[r0 * NB + 2 * t + 1] = acc[j][0][1] + acc[j][1][1];
                }
                if (r0 + 8 < rows) {
                    if (2 * t < NB) pw[(r0 + 8) * NB + 2 * t] = acc[j][0][2] + acc[j][1][2];
                    if (2 * t + 1 < NB) pw[(r0 + 8) * NB + 2 * t + 1] = acc[j][0][3] + acc[j][1][3];
                }
            }
        }
        MK_SYNC();
        MK_TRACE(tr_li, tr_ph, 4);
        MK_GT(tr_li, tr_ph * 3 + 1);
        for (int i = tid; i < nu * NB && !(p.dbg & 4); i += MK_THREADS) {
            const int j = i / NB, n = i % NB;
            float s[RPU];
#pragma unroll
            for (int q = 0; q < RPU; ++q) {
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < KW; ++w) a += part[((int64_t)w * TK_WIDE_ROWS + j * RPU + q) * NB + n];
                s[q] = a;
            }
            epi(gu0 + j, n, s);
        }
        MK_SYNC();
        if (gi + 1 < ng) {                                 // next group's geometry
            gu0 = tk_lo(n_units, gi + 1, ng);
            nu = tk_lo(n_units, gi + 2, ng) - gu0;
            rows = nu * RPU;
            KS = tk_ks(rows, WIDE);
        }
    }
    MK_TRACE(tr_li, tr_ph, 5);
    MK_GT(tr_li, tr_ph * 3 + 2);
}

// x = src[K] (written earlier in this launch by other CTAs) into shared memory with ONE bulk copy issued by thread 0
// (packed-feed variant, one stream): a single L2 round trip instead of two rounds of 8 LDG + 8 STS per thread.
// x = src[K] (written earlier in this launch by other CTAs) into shared memory with ONE bulk copy issued by thread 0
// (packed-feed variant, one stream): a single L2 round trip instead of two rounds of 8 LDG + 8 STS per thread.
__device__ __forceinline__ void preamble_copy_tk(const StackParams& p, unsigned char* smem, const TkFeed& f, uint32_t bar, uint32_t parity,
                                                 const __nv_bfloat16* src, int K) {
    if (threadIdx.x == 0) {
        // src was written through the generic proxy by other SMs (ordered before us by their fence + the grid barrier's
        // acquire); the bulk copy reads through the async proxy: a proxy fence in between
        asm volatile("fence.proxy.async.global;" ::: "memory");
        tk_mbar_expect(bar, (uint32_t)K * 2);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(src), "r"((uint32_t)K * 2), "r"(bar) : "memory");
    }
    tk_mbar_wait(bar, parity, f.sync);      // every thread waits itself: the bytes are then visible to it
}

template <int NB, int RPU, bool WIDE, int FEED, class RowFn, class MidFn, class EpiFn>
__device__ __forceinline__ void run_gemv(const StackParams& p, unsigned char* smem, TkFeed& feed, const __nv_bfloat16* W,
                                         int K, int n_units, RowFn rowfn, MidFn mid, EpiFn epi, uint64_t pol, int tr_li,
                                         int tr_ph) {
    if constexpr (FEED == 1) gemv_phase_tk<NB, RPU, WIDE>(p, smem, feed, K, n_units, mid, epi, tr_li, tr_ph);
    else gemv_phase<NB, RPU>(p, smem, W, K, n_units, rowfn, mid, epi, pol, tr_li, tr_ph);
}

// FEED = 1 launches three warpgroups: two of compute warps and one that holds the producer thread.  The kernel is
// compiled for 168 registers (384 threads); the producer's warpgroup hands its registers back (setmaxnreg.dec) and the
// compute warpgroups grow to 232 (setmaxnreg.inc), which takes most spills out of the main loops and the attention item.
constexpr int TK_EXTRA_THREADS = 128;
template <int NB, int FEED>
__global__ void __launch_bounds__(MK_THREADS + FEED * TK_EXTRA_THREADS, 1)
decode_stack_kernel(const __grid_constant__ StackParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int c = blockIdx.x, G = p.G;
    const int D = p.D, I = p.I, Hq = p.Hq, Hkv = p.Hkv;
    const uint64_t pol = make_evict_first_policy();
    unsigned epoch = 0;
    int phases_left = p.stop_after > 0 ? p.stop_after : 0x7FFFFFFF;
#define MK_PHASE_DONE() do { if (--phases_left == 0) return; } while (0)
    int* s_misc_i = reinterpret_cast<int*>(smem + p.off_misc);   // floats 0..31: norm partials; ints 32..47: flags
    int* s_flag = s_misc_i + 32;
    int* s_qstart = s_misc_i + 36;  // [NB]
    int* s_nq = s_misc_i + 40;      // [NB]
    if (tid < NB) { s_qstart[tid] = tid; s_nq[tid] = 1; }
    if (tid == 0) s_misc_i[44] = 0;
    TkFeed feed = {0, 0, 0, 0, 0, p.sync, s_misc_i + 44, 0};
    if constexpr (FEED == 1) {
        feed.ring = smem_u32(smem + p.off_ring);
        feed.full = smem_u32(smem + p.off_misc + TK_BAR_OFF);
        feed.empty = feed.full + TK_SLOTS * 8;
        feed.xbar = feed.empty + TK_SLOTS * 8;
        if (tid == 0) {
            tk_mbar_init(feed.xbar, 1);          // x of the O projection, one use per layer: parity = layer & 1
            tk_mbar_init(feed.xbar + 8, 1);      // x of the down projection
            for (int s = 0; s < TK_SLOTS; ++s) { tk_mbar_init(feed.full + s * 8, 1); tk_mbar_init(feed.empty + s * 8, MK_WARPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();                                 // all twelve warps, once
        if (warp >= MK_WARPS) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
            if (tid == MK_THREADS) tk_producer(p, feed, pol);
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    }

    // this CTA's share of each phase (balanced contiguous ranges)
    const int n_pairs_all = (Hq + 2 * Hkv) * 64;
    const int qkv_p0 = (int)((int64_t)n_pairs_all * c / G), qkv_p1 = (int)((int64_t)n_pairs_all * (c + 1) / G);
    const int d_r0 = (int)((int64_t)D * c / G), d_r1 = (int)((int64_t)D * (c + 1) / G);
    const int i_r0 = (int)((int64_t)I * c / G), i_r1 = (int)((int64_t)I * (c + 1) / G);
    const int v_r0 = (int)((int64_t)p.V * c / G), v_r1 = (int)((int64_t)p.V * (c + 1) / G);
    auto qkv_row = [&](int u, int which) { const int pr = qkv_p0 + u; return (pr >> 6) * 128 + (pr & 63) + which * 64; };
    auto d_row = [&](int u, int) { return d_r0 + u; };
    auto gu_row = [&](int u, int which) { return i_r0 + u + which * I; };
    auto v_row = [&](int u, int) { return v_r0 + u; };

    int pos[NB], env[NB];
    int max_len = 1;
#pragma unroll
    for (int n = 0; n < NB; ++n) {
        const int len = p.kv_len[n];
        pos[n] = len - 1;
        env[n] = p.seq_env[n];
        max_len = max(max_len, len);
    }
    DecodeParams dp;
    dp.arena = p.arena;
    dp.q = p.qrot;
    dp.n_q_head = Hq;
    dp.seq_env = p.seq_env;
    dp.q_start = s_qstart;
    dp.n_q = s_nq;
    dp.kv_len = p.kv_len;
    dp.groups_per_seq = 1;
    dp.tok_per_group = DROWS / (Hq / Hkv);
    dp.n_splits = p.n_splits;
    dp.split_keys = (((max_len + p.n_splits - 1) / p.n_splits + DT - 1) / DT) * DT;
    dp.scale_log2 = p.scale_log2;
    dp.part_o = p.part_o;
    dp.part_m = p.part_m;
    dp.part_l = p.part_l;
    dp.out = p.attn_out;
    dp.trace = p.trace ? p.trace + (int64_t)(p.n_layer + 1) * 6 * 8 : nullptr;   // 8 stamps of the last layer's item
    const int n_items = NB * Hkv * p.n_splits;
    __nv_bfloat16* hb0 = p.hbuf;
    __nv_bfloat16* hb1 = p.hbuf + (int64_t)NB * D;
    const int64_t plane_sz = (int64_t)p.arena.max_len * HD;
    __nv_bfloat16* arena_base = reinterpret_cast<__nv_bfloat16*>(p.arena.base);

    for (int li = 0; li < p.n_layer; ++li) {
        const long long* L = p.layer_ptrs + (int64_t)li * 8;
        const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(__ldg(L + 0));
        const __nv_bfloat16* bqkv = reinterpret_cast<const __nv_bfloat16*>(__ldg(L + 1));
        const __nv_bfloat16* wo = reinterpret_cast<const __nv_bfloat16*>(__ldg(L + 2));
        const __nv_bfloat16* wgu = reinterpret_cast<const __nv_bfloat16*>(__ldg(L + 3));
        const __nv_bfloat16* wd = reinterpret_cast<const __nv_bfloat16*>(__ldg(L + 4));
        const __nv_bfloat16* ln1 = reinterpret_cast<const __nv_bfloat16*>(__ldg(L + 5));
        const __nv_bfloat16* ln2 = reinterpret_cast<const __nv_bfloat16*>(__ldg(L + 6));

        // ---- QKV: hn = h (+ pending MLP output); x = RMSNorm(hn); q,k,v = x W^T + b; RoPE; append
        // bias and RoPE factors of this thread's epilogue unit are fetched before the barrier is waited for (they do not
        // depend on this step): the epilogue then needs no load at all
        const bool q_pre = NB == 1 && qkv_p1 - qkv_p0 <= 16;
        float q_b0 = 0.f, q_b1 = 0.f, q_c0 = 0.f, q_s0 = 0.f, q_c1 = 0.f, q_s1 = 0.f;
        run_gemv<NB, 2, false, FEED>(p, smem, feed, wqkv, D, qkv_p1 - qkv_p0, qkv_row,
            [&] {
                if (q_pre && tid < qkv_p1 - qkv_p0) {
                    const int pr = qkv_p0 + tid, h = pr >> 6, d = pr & 63, row = h * 128 + d;
                    if (bqkv) { q_b0 = __bfloat162float(bqkv[row]); q_b1 = __bfloat162float(bqkv[row + 64]); }
                    if (h < Hq + Hkv) {
                        const __nv_bfloat16* cr = p.cos_t + (int64_t)pos[0] * HD;
                        const __nv_bfloat16* sr = p.sin_t + (int64_t)pos[0] * HD;
                        q_c0 = __bfloat162float(cr[d]); q_s0 = __bfloat162float(sr[d]);
                        q_c1 = __bfloat162float(cr[d + 64]); q_s1 = __bfloat162float(sr[d + 64]);
                    }
                }
                const NormW lw = norm_w_load(p, ln1);
                if (li > 0) grid_wait(p.sync, epoch, G);     // previous layer's down projection
                MK_TRACE(li, 0, 1);
                MK_GT(li, 0);
                if (!(p.dbg & 1)) preamble_norm<NB>(p, smem, hb0, li == 0, lw);
            },
            [&](int u, int n, const float (&s)[2]) {
                const int pr = qkv_p0 + u, h = pr >> 6, d = pr & 63;
                const int row = h * 128 + d;
                const float xl = round_bf16(s[0] + (q_pre ? q_b0 : (bqkv ? __bfloat162float(bqkv[row]) : 0.f)));
                const float xh = round_bf16(s[1] + (q_pre ? q_b1 : (bqkv ? __bfloat162float(bqkv[row + 64]) : 0.f)));
                __nv_bfloat16 lo, hi;
                if (h < Hq + Hkv) {
                    const __nv_bfloat16* cr = p.cos_t + (int64_t)pos[n] * HD;
                    const __nv_bfloat16* sr = p.sin_t + (int64_t)pos[n] * HD;
                    const float a = round_bf16(__fmul_rn(xl, q_pre ? q_c0 : __bfloat162float(cr[d])));
                    const float b = round_bf16(__fmul_rn(-xh, q_pre ? q_s0 : __bfloat162float(sr[d])));
                    const float cc = round_bf16(__fmul_rn(xh, q_pre ? q_c1 : __bfloat162float(cr[d + 64])));
                    const float dd = round_bf16(__fmul_rn(xl, q_pre ? q_s1 : __bfloat162float(sr[d + 64])));
                    lo = __float2bfloat16_rn(__fadd_rn(a, b));
                    hi = __float2bfloat16_rn(__fadd_rn(cc, dd));
                } else {
                    lo = __float2bfloat16_rn(xl);
                    hi = __float2bfloat16_rn(xh);
                }
                __nv_bfloat16* dst;
                if (h < Hq) {
                    dst = p.qrot + ((int64_t)n * Hq + h) * HD;
                } else {
                    const int kv = h < Hq + Hkv ? 0 : 1;
                    const int head = h - Hq - kv * Hkv;
                    dst = arena_base + ((((int64_t)env[n] * p.arena.n_layer + li) * 2 + kv) * Hkv + head) * plane_sz +
                          (int64_t)pos[n] * HD;
                }
                dst[d] = lo;
                dst[d + 64] = hi;
            },
            pol, li, 0);
        MK_PHASE_DONE();

        // ---- attention: split-KV work items on warps 0-3; the CTA that completes the last split of a
        //      (stream, kv head) merges the partials.  The first item's K/V tiles are requested before the barrier
        //      (they predate this step except for the newest row, which is fetched again afterwards).
        dp.layer = li;
        grid_arrive(p.sync, epoch, G);
        if (c < n_items && warp < DWARPS)
            attn_item<1, FEED ? TK_ATTN_STAGES : DSTAGES>(dp, c % p.n_splits, (c / p.n_splits) % Hkv, c / (p.n_splits * Hkv), smem, tid);
        grid_wait(p.sync, epoch, G);
        MK_TRACE(li, 1, 0);
        MK_GT(li, 3);
        for (int item = c; item < n_items; item += G) {
            const int split = item % p.n_splits, kvh = (item / p.n_splits) % Hkv, b = item / (p.n_splits * Hkv);
            if (warp < DWARPS) {
                if (item == c) attn_item<2, FEED ? TK_ATTN_STAGES : DSTAGES>(dp, split, kvh, b, smem, tid);
                else attn_item<0, FEED ? TK_ATTN_STAGES : DSTAGES>(dp, split, kvh, b, smem, tid);
                MK_TRACE(li, 1, 1);
            }
            MK_SYNC();
            MK_TRACE(li, 1, 2);
            if (tid == 0) {
                // release (cumulative over the CTA barrier): the item's partials are at L2 before the count moves; the CTA
                // that sees the last count reads them with ld.global.cg (combine_splits)
                unsigned old;
                asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.sync + 16 + b * Hkv + kvh) : "memory");
                *s_flag = ((old + 1) % (unsigned)p.n_splits) == 0;
            }
            MK_SYNC();
            MK_TRACE(li, 1, 3);
            if (*s_flag) combine_splits(p, b, kvh);
            MK_SYNC();
        }
        MK_TRACE(li, 1, 5);
        MK_GT(li, 5);
        if (FEED == 1 && tid == 0) *reinterpret_cast<volatile int*>(s_misc_i + 44) = li + 1;   // opens the producer's O gate (dbg bit 3)
        MK_PHASE_DONE();
        grid_arrive(p.sync, epoch, G);

        // ---- O projection
        // the epilogue adds the projection to the residual stream: hb1 = resid_in + o (HF: hidden = residual + attn).
        // resid_in = hb0, or the token's embedding row in layer 0; this thread's element is fetched before the barrier.
        const bool r_pre = NB == 1 && d_r1 - d_r0 <= 32;
        float r_val = 0.f;
        auto resid_in = [&](int n, int row) {
            if (li > 0) return ld_cg_bf16(hb0 + (int64_t)n * D + row);
            int64_t tok = p.tok_in[n];
            tok = tok < 0 ? 0 : (tok >= p.V ? p.V - 1 : tok);
            return __bfloat162float(p.embed[tok * D + row]);
        };
        run_gemv<NB, 1, false, FEED>(p, smem, feed, wo, Hq * HD, d_r1 - d_r0, d_row,
            [&] {
                grid_wait(p.sync, epoch, G);
                MK_TRACE(li, 2, 1);
                MK_GT(li, 6);
                if constexpr (FEED == 1 && NB == 1 && (TK_XCOPY & 1)) preamble_copy_tk(p, smem, feed, feed.xbar, (uint32_t)(li & 1), p.attn_out, Hq * HD);
                else preamble_copy<NB>(p, smem, p.attn_out, Hq * HD);
                // requested here, used in the epilogue: the load flies under the main loop
                if (r_pre && tid < d_r1 - d_r0) r_val = resid_in(0, d_r0 + tid);
            },
            [&](int u, int n, const float (&s)[1]) {
                const float r = r_pre ? r_val : resid_in(n, d_r0 + u);
                hb1[(int64_t)n * D + d_r0 + u] = __float2bfloat16_rn(r + round_bf16(s[0]));
            },
            pol, li, 2);
        MK_PHASE_DONE();
        grid_arrive(p.sync, epoch, G);

        // ---- gate/up: h2 = hn + o; x = RMSNorm(h2); act = silu(x Wg^T) * (x Wu^T)
        run_gemv<NB, 2, true, FEED>(p, smem, feed, wgu, D, i_r1 - i_r0, gu_row,
            [&] {
                const NormW lw = norm_w_load(p, ln2);
                grid_wait(p.sync, epoch, G);
                MK_TRACE(li, 3, 1);
                MK_GT(li, 9);
                preamble_norm<NB>(p, smem, hb1, false, lw);
            },
            [&](int u, int n, const float (&s)[2]) {
                const float gv = round_bf16(s[0]), uv = round_bf16(s[1]);
                const float a = round_bf16(gv / (1.0f + expf(-gv)));
                p.act[(int64_t)n * I + i_r0 + u] = __float2bfloat16_rn(a * uv);
            },
            pol, li, 3);
        MK_PHASE_DONE();
        grid_arrive(p.sync, epoch, G);

        // ---- down projection; the epilogue closes the layer's residual stream: hb0 = hb1 + mlp (HF: residual + mlp)
        float m_val = 0.f;
        run_gemv<NB, 1, false, FEED>(p, smem, feed, wd, I, d_r1 - d_r0, d_row,
            [&] {
                grid_wait(p.sync, epoch, G);
                MK_TRACE(li, 4, 1);
                MK_GT(li, 12);
                if constexpr (FEED == 1 && NB == 1 && (TK_XCOPY & 2)) preamble_copy_tk(p, smem, feed, feed.xbar + 8, (uint32_t)(li & 1), p.act, I);
                else preamble_copy<NB>(p, smem, p.act, I);
                if (r_pre && tid < d_r1 - d_r0) m_val = ld_cg_bf16(hb1 + d_r0 + tid);
            },
            [&](int u, int n, const float (&s)[1]) {
                const float r = r_pre ? m_val : ld_cg_bf16(hb1 + (int64_t)n * D + d_r0 + u);
                hb0[(int64_t)n * D + d_r0 + u] = __float2bfloat16_rn(r + round_bf16(s[0]));
            },
            pol, li, 4);
        MK_PHASE_DONE();
        grid_arrive(p.sync, epoch, G);
    }

    // ---- final norm, lm_head, greedy argmax (ties -> lowest index, as torch.argmax on these logits)
    float best = -INFINITY;
    int best_i = 0x7FFFFFFF;
    run_gemv<NB, 1, true, FEED>(p, smem, feed, p.lm_head, D, v_r1 - v_r0, v_row,
        [&] {
            const NormW lw = norm_w_load(p, p.final_norm);
            grid_wait(p.sync, epoch, G);
            preamble_norm<NB>(p, smem, hb0, false, lw);
        },
        [&](int u, int n, const float (&s)[1]) {
            // HF's lm_head is a bf16 Linear upcast with .float(): the greedy pick (and its lowest-index tie-break)
            // is over bf16-rounded logits, as on the svln_gemv and cuBLAS paths
            const int row = v_r0 + u;
            const float lg = round_bf16(s[0]);
            if (p.logits) p.logits[(int64_t)n * p.V + row] = lg;
            if (lg > best) { best = lg; best_i = row; }
        },
        pol, p.n_layer, 0);
    {
        // threads with tid % NB == n saw stream n only (MK_THREADS is a multiple of NB)
        float* red_v = FEED ? reinterpret_cast<float*>(smem + p.off_part) : reinterpret_cast<float*>(smem + p.off_misc) + 64;
        int* red_i = reinterpret_cast<int*>(red_v + MK_THREADS);
        red_v[tid] = best;
        red_i[tid] = best_i;
        MK_SYNC();
        if (tid < NB) {
            float bv = -INFINITY;
            int bi = 0x7FFFFFFF;
            for (int j = tid; j < MK_THREADS; j += NB) {
                const float v = red_v[j];
                const int i = red_i[j];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            p.best_v[c * NB + tid] = bv;
            p.best_i[c * NB + tid] = bi;
        }
    }
    grid_arrive(p.sync, epoch, G);
    grid_wait(p.sync, epoch, G);
    if (c == 0 && tid < NB) {
        float bv = -INFINITY;
        int bi = 0x7FFFFFFF;
        for (int j = 0; j < G; ++j) {
            const float v = ld_cg_f(p.best_v + j * NB + tid);
            const int i = (int)ld_cg_u32(p.best_i + j * NB + tid);
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        const unsigned aborted = ld_cg_u32(p.sync + 1);
        p.tok_out[tid] = aborted ? -1 : (bi == 0x7FFFFFFF ? 0 : bi);
    }
}

// one thread per 16 bytes of a weight matrix: where does it go in the packed stream?
__global__ void __launch_bounds__(256)
tk_pack_kernel(const int4* __restrict__ W, int N, int K, int ph, int li, const TkDims d, unsigned char* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int K8 = K >> 3;
    if (idx >= (long long)N * K8) return;
    const int r = (int)(idx / K8), kc = (int)(idx % K8);
    const TkPhase P = tk_phase(d, ph);
    int unit = r, which = 0;
    if (ph == 0) {
        const int h = r >> 7, dd = r & 127;
        which = dd >> 6;
        unit = h * 64 + (dd & 63);
    } else if (ph == 2) {
        which = r >= d.I ? 1 : 0;
        unit = which ? r - d.I : r;
    }
    const int c = tk_owner(unit, P.n_total, d.G);
    const int u0 = tk_lo(P.n_total, c, d.G), nu = tk_lo(P.n_total, c + 1, d.G) - u0;
    const int ul = unit - u0;
    const int ng = tk_groups(nu, P.rpu, P.wide);
    const int gi = tk_owner(ul, nu, ng);
    const int g0 = tk_lo(nu, gi, ng), g1 = tk_lo(nu, gi + 1, ng);
    const int rows = (g1 - g0) * P.rpu, rg = (ul - g0) * P.rpu + which;
    const int KS = tk_ks(rows, P.wide);
    const int s = kc >> 2, t = kc & 3, stage = s / KS, s_in = s % KS;
    const long long layer_bytes = tk_layer_before(d, c + 1) - tk_layer_before(d, c);
    long long off = tk_cta_base(d, c);
    if (ph < 4) {
        off += (long long)li * layer_bytes;
        for (int q = 0; q < ph; ++q) {
            const TkPhase Q = tk_phase(d, q);
            off += (long long)(tk_lo(Q.n_total, c + 1, d.G) - tk_lo(Q.n_total, c, d.G)) * Q.rpu * Q.K * 2;
        }
    } else {
        off += (long long)d.n_layer * layer_bytes;
    }
    off += (long long)g0 * P.rpu * K * 2 + (long long)stage * KS * rows * 64 + ((long long)s_in * rows + rg) * 64 + t * 16;
    *reinterpret_cast<int4*>(out + off) = __ldg(W + idx);
}

static size_t ws_align(size_t x) { return (x + 255) & ~(size_t)255; }

struct StackLayout {
    size_t hbuf, qrot, attn_out, o, act, pending, part_o, part_m, part_l, best_v, best_i, sync, total;
};

static StackLayout stack_layout(const svln_decode_stack* d, int G, int n_splits) {
    StackLayout l;
    const size_t B = d->n_stream, D = d->hidden, I = d->inter, Hq = d->n_q_head, Hkv = d->n_kv_head;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += ws_align(bytes); return at; };
    l.hbuf = take(2 * B * D * 2);
    l.qrot = take(B * Hq * HD * 2);
    l.attn_out = take(B * Hq * HD * 2);
    l.o = take(B * D * 2);
    l.act = take(B * I * 2);
    l.pending = take(B * D * 2);
    const size_t slots = B * Hkv * n_splits;
    l.part_o = take(slots * DROWS * HD * 4);
    l.part_m = take(slots * DROWS * 4);
    l.part_l = take(slots * DROWS * 4);
    l.best_v = take((size_t)G * B * 4);
    l.best_i = take((size_t)G * B * 4);
    l.sync = take(64 * 4);
    l.total = off;
    return l;
}

static int stack_grid() {
    static const int g_env = getenv("SVLN_STACK_CTAS") ? atoi(getenv("SVLN_STACK_CTAS")) : 0;
    const int sms = sm_count();
    return g_env > 0 && g_env <= sms ? g_env : sms;
}

static int stack_splits(int G, int B, int Hkv) {
    static const int s_env = getenv("SVLN_STACK_SPLITS") ? atoi(getenv("SVLN_STACK_SPLITS")) : 0;
    if (s_env >= 1 && s_env <= MK_MAX_SPLITS && s_env * B * Hkv <= G) return s_env;
    int s = G / (B * Hkv);
    if (s < 1) s = 1;
    if (s > MK_MAX_SPLITS) s = MK_MAX_SPLITS;
    return s;
}

static int stack_check(const svln_kv_arena* arena, const svln_decode_stack* d, const char* who) {
    SVLN_REQUIRE(arena && arena->base && d, SVLN_ERR_INVALID, "%s: null arena / descriptor", who);
    SVLN_REQUIRE(arena->head_dim == HD && d->head_dim == HD, SVLN_ERR_INVALID, "%s: head_dim must be 128", who);
    SVLN_REQUIRE(d->n_stream >= 1 && d->n_stream <= MK_MAX_B, SVLN_ERR_RANGE, "%s: n_stream %d outside [1,%d]", who,
                 d->n_stream, MK_MAX_B);
    SVLN_REQUIRE(d->n_layer == arena->n_layer && d->n_kv_head == arena->n_kv_head, SVLN_ERR_INVALID,
                 "%s: descriptor does not match the arena", who);
    SVLN_REQUIRE(d->n_q_head > 0 && d->n_kv_head > 0 && d->n_q_head % d->n_kv_head == 0 &&
                 d->n_q_head / d->n_kv_head <= DROWS, SVLN_ERR_INVALID, "%s: bad head counts", who);
    SVLN_REQUIRE(d->hidden > 0 && d->hidden % 32 == 0 && d->hidden <= 4096 && d->inter > 0 && d->inter % 32 == 0 &&
                 (d->n_q_head * HD) % 32 == 0 && d->vocab > 0, SVLN_ERR_INVALID,
                 "%s: hidden/inter must be multiples of 32 (hidden <= 4096)", who);
    return SVLN_OK;
}

}  // namespace svln

using namespace svln;

extern "C" {

size_t svln_decode_stack_workspace_bytes(const svln_decode_stack* d) {
    if (!d || d->n_stream < 1 || d->n_stream > MK_MAX_B || d->n_kv_head < 1) return 0;
    const int G = stack_grid();
    return stack_layout(d, G, stack_splits(G, d->n_stream, d->n_kv_head)).total;
}

/* debug: abort state of the last launch on this workspace: 0 = ok, else 1-based index of the grid barrier that timed
 * out; cta / count = who gave up and the arrival count it saw.  Synchronous (cudaMemcpy). */
int svln_decode_stack_status(const svln_decode_stack* d, int* cta, int* count) {
    if (!d || !d->workspace) return -1;
    const int G = stack_grid();
    const StackLayout l = stack_layout(d, G, stack_splits(G, d->n_stream, d->n_kv_head));
    unsigned host[4] = {0, 0, 0, 0};
    if (cudaMemcpy(host, (unsigned char*)d->workspace + l.sync, sizeof(host), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cta) *cta = (int)host[2];
    if (count) *count = (int)host[3];
    return (int)host[1];
}

/* debug: copies the clock64 trace of the last launch (SVLN_STACK_TRACE=1) to host; returns entries copied */
static long long* g_trace_dev = nullptr;
static int g_trace_n = 0;
int svln_decode_stack_trace(long long* host, int max_entries) {
    if (!g_trace_dev || !host) return 0;
    const int n = max_entries < g_trace_n ? max_entries : g_trace_n;
    if (cudaMemcpy(host, g_trace_dev, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return n;
}

static TkDims tk_dims(const svln_decode_stack* d, int G) {
    return TkDims{d->n_layer, d->hidden, d->inter, d->n_q_head, d->n_kv_head, d->vocab, G};
}

// The packed layout depends on the grid and on the model dims: remember what every packed buffer of this process was
// packed for, so that a step on a stale or foreign buffer fails loudly instead of streaming the wrong bytes.
struct PackedRecord { const void* ptr; TkDims dims; };
static PackedRecord g_packed[32];
static int g_packed_n = 0;
static bool same_dims(const TkDims& a, const TkDims& b) {
    return a.n_layer == b.n_layer && a.D == b.D && a.I == b.I && a.Hq == b.Hq && a.Hkv == b.Hkv && a.V == b.V && a.G == b.G;
}
static void packed_remember(const void* ptr, const TkDims& dims) {
    for (int i = 0; i < g_packed_n; ++i)
        if (g_packed[i].ptr == ptr) { g_packed[i].dims = dims; return; }
    g_packed[g_packed_n < 32 ? g_packed_n++ : 31] = PackedRecord{ptr, dims};
}
static bool packed_matches(const void* ptr, const TkDims& dims) {
    for (int i = 0; i < g_packed_n; ++i)
        if (g_packed[i].ptr == ptr) return same_dims(g_packed[i].dims, dims);
    return false;
}

size_t svln_decode_stack_packed_bytes(const svln_decode_stack* d) {
    if (!d || d->n_layer < 1 || d->hidden < 32 || d->inter < 32 || d->n_q_head < 1 || d->n_kv_head < 1 || d->vocab < 1 ||
        d->head_dim != HD) return 0;
    const int G = stack_grid();
    return (size_t)tk_cta_base(tk_dims(d, G), G);
}

int svln_decode_stack_pack(const svln_decode_stack* d, void* packed, size_t packed_bytes, svln_stream_t stream) {
    SVLN_REQUIRE(d && packed && d->layer_ptrs && d->lm_head, SVLN_ERR_INVALID, "svln_decode_stack_pack: null pointer");
    SVLN_REQUIRE(d->head_dim == HD && d->hidden % 32 == 0 && d->inter % 32 == 0 && d->hidden <= 4096, SVLN_ERR_INVALID,
                 "svln_decode_stack_pack: hidden/inter must be multiples of 32 (hidden <= 4096), head_dim 128");
    SVLN_REQUIRE((long long)d->vocab * stack_grid() < (1ll << 31) && (long long)d->inter * stack_grid() < (1ll << 31), SVLN_ERR_RANGE,
                 "svln_decode_stack_pack: vocab / inter too large for the 32-bit row partition");
    const size_t need = svln_decode_stack_packed_bytes(d);
    SVLN_REQUIRE(need > 0 && packed_bytes >= need, SVLN_ERR_NOMEM, "svln_decode_stack_pack: %zu < %zu bytes", packed_bytes, need);
    SVLN_REQUIRE(((uintptr_t)packed & 127) == 0, SVLN_ERR_INVALID, "svln_decode_stack_pack: destination not 128-byte aligned");
    const int G = stack_grid();
    const TkDims dims = tk_dims(d, G);
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<long long> ptrs((size_t)d->n_layer * 8);
    SVLN_CUDA(cudaMemcpyAsync(ptrs.data(), d->layer_ptrs, ptrs.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    SVLN_CUDA(cudaStreamSynchronize(st));
    auto pack = [&](const void* W, int ph, int li) {
        const TkPhase P = tk_phase(dims, ph);
        const long long n16 = (long long)P.n_total * P.rpu * (P.K >> 3);
        tk_pack_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((const int4*)W, P.n_total * P.rpu, P.K, ph, li, dims,
                                                                     (unsigned char*)packed);
    };
    for (int li = 0; li < d->n_layer; ++li) {
        const long long* L = ptrs.data() + (size_t)li * 8;
        SVLN_REQUIRE(L[0] && L[2] && L[3] && L[4], SVLN_ERR_INVALID, "svln_decode_stack_pack: null weight pointer in layer %d", li);
        pack((const void*)L[0], 0, li);
        pack((const void*)L[2], 1, li);
        pack((const void*)L[3], 2, li);
        pack((const void*)L[4], 3, li);
    }
    pack(d->lm_head, 4, 0);
    packed_remember(packed, dims);
    return check_launch("tk_pack_kernel");
}

int svln_decode_stack_step(const svln_kv_arena* arena, const svln_decode_stack* d, svln_stream_t stream) {
    int rc = stack_check(arena, d, "svln_decode_stack_step");
    if (rc) return rc;
    SVLN_REQUIRE(d->layer_ptrs && d->embed && d->final_norm && d->lm_head && d->rope_cos && d->rope_sin && d->tok_in &&
                 d->tok_out && d->seq_env && d->kv_len && d->workspace, SVLN_ERR_INVALID,
                 "svln_decode_stack_step: null pointer");
    const int G = stack_grid();
    const int B = d->n_stream;
    StackParams p;
    p.arena = *arena;
    p.n_layer = d->n_layer; p.D = d->hidden; p.I = d->inter; p.Hq = d->n_q_head; p.Hkv = d->n_kv_head; p.V = d->vocab;
    p.G = G;
    p.eps = d->rms_eps;
    p.scale_log2 = d->attn_scale * 1.4426950408889634f;
    p.layer_ptrs = (const long long*)d->layer_ptrs;
    p.embed = (const __nv_bfloat16*)d->embed;
    p.final_norm = (const __nv_bfloat16*)d->final_norm;
    p.lm_head = (const __nv_bfloat16*)d->lm_head;
    p.cos_t = (const __nv_bfloat16*)d->rope_cos;
    p.sin_t = (const __nv_bfloat16*)d->rope_sin;
    p.tok_in = d->tok_in; p.tok_out = d->tok_out; p.logits = d->logits;
    p.seq_env = d->seq_env; p.kv_len = d->kv_len;
    p.n_splits = stack_splits(G, B, p.Hkv);
    const StackLayout l = stack_layout(d, G, p.n_splits);
    SVLN_REQUIRE(d->workspace_bytes >= l.total, SVLN_ERR_NOMEM, "svln_decode_stack_step: workspace %zu < %zu bytes",
                 d->workspace_bytes, l.total);
    SVLN_REQUIRE(((uintptr_t)d->workspace & 255) == 0, SVLN_ERR_INVALID, "svln_decode_stack_step: workspace not 256-B aligned");
    unsigned char* ws = (unsigned char*)d->workspace;
    p.hbuf = (__nv_bfloat16*)(ws + l.hbuf); p.qrot = (__nv_bfloat16*)(ws + l.qrot);
    p.attn_out = (__nv_bfloat16*)(ws + l.attn_out); p.o = (__nv_bfloat16*)(ws + l.o);
    p.act = (__nv_bfloat16*)(ws + l.act); p.pending = (__nv_bfloat16*)(ws + l.pending);
    p.part_o = (float*)(ws + l.part_o); p.part_m = (float*)(ws + l.part_m); p.part_l = (float*)(ws + l.part_l);
    p.best_v = (float*)(ws + l.best_v); p.best_i = (int*)(ws + l.best_i); p.sync = (unsigned*)(ws + l.sync);
    const int kmax = p.I > p.D ? (p.I > p.Hq * HD ? p.I : p.Hq * HD) : (p.D > p.Hq * HD ? p.D : p.Hq * HD);
    p.xs_stride = kmax + MK_XPAD;
    p.off_part = (int)align_up((size_t)B * p.xs_stride * 2, 128);
    p.off_rowmap = p.off_part + MK_WARPS * MK_CAP * B * 4;
    p.off_ring = (int)align_up((size_t)p.off_rowmap + (MK_CAP + MK_CAP / 16 + 2) * 4, 128);
    size_t misc = (size_t)p.off_ring + (size_t)MK_RING_SLOTS * MK_THREADS * 16;
    if (misc < (size_t)DEC_SMEM) misc = DEC_SMEM;
    static const int feed_env = getenv("SVLN_STACK_FEED") ? atoi(getenv("SVLN_STACK_FEED")) : 1;
    const bool packed_feed = d->packed != nullptr && feed_env != 0;
    p.packed = packed_feed ? (const unsigned char*)d->packed : nullptr;
    if (packed_feed) {
        // x and the partial sums share the attention item's 96 KB; the stage ring sits behind it, so the producer
        // may fill it whatever the compute warps are doing
        SVLN_REQUIRE(d->packed_bytes >= svln_decode_stack_packed_bytes(d), SVLN_ERR_NOMEM,
                     "svln_decode_stack_step: packed stream %zu < %zu bytes", d->packed_bytes, svln_decode_stack_packed_bytes(d));
        SVLN_REQUIRE(packed_matches(d->packed, tk_dims(d, G)), SVLN_ERR_INVALID,
                     "svln_decode_stack_step: `packed` was not written by svln_decode_stack_pack for these dims and a grid of %d CTAs", G);
        size_t ring = align_up((size_t)p.off_part + (size_t)MK_WARPS * TK_WIDE_ROWS * B * 4, 1024);
        if (ring < (size_t)TK_ATTN_SMEM) ring = TK_ATTN_SMEM;
        p.off_ring = (int)ring;
        misc = ring + (size_t)TK_SLOTS * TK_SLOT_BYTES;
    }
    p.off_misc = (int)misc;
    static const int pad_env = getenv("SVLN_STACK_SMEM_PAD") ? atoi(getenv("SVLN_STACK_SMEM_PAD")) : 0;   // experiment: shrink L1
    const size_t smem = misc + (packed_feed ? TK_MISC : MK_MISC) + (size_t)pad_env;
    SVLN_REQUIRE(smem <= 227 * 1024, SVLN_ERR_RANGE, "svln_decode_stack_step: %zu bytes of shared memory needed", smem);
    static const int stop_env = getenv("SVLN_STACK_STOP") ? atoi(getenv("SVLN_STACK_STOP")) : 0;
    p.stop_after = stop_env;
    static const int dbg_env = getenv("SVLN_STACK_DBG") ? atoi(getenv("SVLN_STACK_DBG")) : 0;
    p.dbg = dbg_env;
    static long long* trace_buf = nullptr;
    if (getenv("SVLN_STACK_TRACE") && !trace_buf) {
        SVLN_CUDA(cudaMalloc(&trace_buf, ((size_t)(p.n_layer + 1) * 6 * 8 + 8 + 160 * 16 + 4096) * sizeof(long long)));
        SVLN_CUDA(cudaMemset(trace_buf, 0, ((size_t)(p.n_layer + 1) * 6 * 8 + 8 + 160 * 16 + 4096) * sizeof(long long)));
    }
    p.trace = trace_buf;
    g_trace_dev = trace_buf;
    g_trace_n = (p.n_layer + 1) * 6 * 8 + 8 + 160 * 16 + 4096;
    static const int coop = getenv("SVLN_STACK_COOP") ? atoi(getenv("SVLN_STACK_COOP")) : 1;

    cudaStream_t st = (cudaStream_t)stream;
    SVLN_CUDA(cudaMemsetAsync(p.sync, 0, 64 * 4, st));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(MK_THREADS + (packed_feed ? TK_EXTRA_THREADS : 0));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = coop ? 1 : 0;
    static size_t smem_set[3][64] = {{0}};  // per device: function attributes do not carry over to another GPU
    int dev_i = 0;
    SVLN_CUDA(cudaGetDevice(&dev_i));
    dev_i = dev_i < 0 || dev_i >= 64 ? 0 : dev_i;
    SVLN_REQUIRE(B == 1 || packed_feed, SVLN_ERR_RANGE, "svln_decode_stack_step: n_stream %d needs the packed weight stream", B);
    const int variant = packed_feed ? B : 0;               // 0: <1, 0>, 1: <1, 1>, 2: <2, 1>
    if (smem > smem_set[variant][dev_i]) {
        if (variant == 2) SVLN_CUDA(cudaFuncSetAttribute(decode_stack_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else if (variant == 1) SVLN_CUDA(cudaFuncSetAttribute(decode_stack_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else SVLN_CUDA(cudaFuncSetAttribute(decode_stack_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[variant][dev_i] = smem;
    }
    if (variant == 2) SVLN_CUDA(cudaLaunchKernelEx(&cfg, decode_stack_kernel<2, 1>, p));
    else if (variant == 1) SVLN_CUDA(cudaLaunchKernelEx(&cfg, decode_stack_kernel<1, 1>, p));
    else SVLN_CUDA(cudaLaunchKernelEx(&cfg, decode_stack_kernel<1, 0>, p));
    return check_launch("decode_stack_kernel");
}

}  // extern "C"

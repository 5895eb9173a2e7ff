// Weight-streaming GEMV for the action-token decode of a single (or a few) stream(s): y[m,n] = sum_k x[m,k] W[n,k]
// (+ bias[n]), bf16 in, fp32 accumulate, bf16 or fp32 out — the arithmetic of the Qwen2 linears that HF runs
// through cuBLAS (reached from stream_video_vln.py:340).  At M = 1 the decode step is nothing but streaming the
// 15.2 GB of weights once per token (SURVEY §7 "decode at batch 1 is weight-bound"); cuBLAS's GEMV kernels reach
// 2.4-5.5 TB/s on the four layer shapes, this kernel keeps every warp's 128-bit loads of one or more whole W rows
// in flight and reads x through L1.  Used only for M <= 4; larger batches are GEMMs and stay cuBLAS.
#include "svln_common.cuh"

namespace svln {

constexpr int GV_THREADS = 256;
constexpr int GV_UNROLL = 8;

template <int M, int ROWS, bool F32_OUT>
__global__ void __launch_bounds__(GV_THREADS)
gemv_kernel(const int4* __restrict__ W, const int4* __restrict__ x, const __nv_bfloat16* __restrict__ bias,
            void* __restrict__ out, int N, int K8) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * GV_THREADS + threadIdx.x) >> 5;
    const int n0 = warp_global * ROWS;
    if (n0 >= N) return;
    float acc[ROWS][M];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
    const int4* wrow[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) wrow[r] = W + (int64_t)min(n0 + r, N - 1) * K8;
    for (int c0 = lane; c0 < K8; c0 += 32 * GV_UNROLL) {
        int4 wv[GV_UNROLL][ROWS];
#pragma unroll
        for (int u = 0; u < GV_UNROLL; ++u) {
            const int c = c0 + u * 32;
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                if (c < K8) wv[u][r] = ld_stream(wrow[r] + c);
        }
#pragma unroll
        for (int u = 0; u < GV_UNROLL; ++u) {
            const int c = c0 + u * 32;
            if (c < K8) {
                float xf[M][8];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const int4 xv = __ldg(x + (int64_t)m * K8 + c);
                    const uint32_t* xu = reinterpret_cast<const uint32_t*>(&xv);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        xf[m][2 * k] = bf16_bits_to_f32(xu[k] & 0xFFFFu);
                        xf[m][2 * k + 1] = bf16_bits_to_f32(xu[k] >> 16);
                    }
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint32_t* wu = reinterpret_cast<const uint32_t*>(&wv[u][r]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float w0 = bf16_bits_to_f32(wu[k] & 0xFFFFu), w1 = bf16_bits_to_f32(wu[k] >> 16);
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            acc[r][m] = fmaf(w0, xf[m][2 * k], acc[r][m]);
                            acc[r][m] = fmaf(w1, xf[m][2 * k + 1], acc[r][m]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) acc[r][m] = warp_sum(acc[r][m]);
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int n = n0 + r;
            if (n >= N) break;
            const float b = bias ? __bfloat162float(bias[n]) : 0.f;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                // HF adds the bias inside the GEMM epilogue in fp32 and rounds once
                const float v = acc[r][m] + b;
                // fp32 output (lm_head logits) carries bf16-ROUNDED values: HF's lm_head is a bf16 Linear whose
                // result is upcast with .float(), so greedy argmax and its lowest-index
                // tie-break see bf16 logits on every path (this kernel, the fused stack, cuBLAS for M > 4)
                if (F32_OUT) reinterpret_cast<float*>(out)[(int64_t)m * N + n] = round_bf16(v);
                else reinterpret_cast<__nv_bfloat16*>(out)[(int64_t)m * N + n] = __float2bfloat16_rn(v);
            }
        }
    }
}

template <int M, int ROWS>
static void launch_gemv(const void* W, const void* x, const void* bias, void* out, int N, int K, int f32_out, cudaStream_t st) {
    const int warps = (N + ROWS - 1) / ROWS;
    const int blocks = (warps * 32 + GV_THREADS - 1) / GV_THREADS;
    if (f32_out)
        gemv_kernel<M, ROWS, true><<<blocks, GV_THREADS, 0, st>>>((const int4*)W, (const int4*)x, (const __nv_bfloat16*)bias, out, N, K / 8);
    else
        gemv_kernel<M, ROWS, false><<<blocks, GV_THREADS, 0, st>>>((const int4*)W, (const int4*)x, (const __nv_bfloat16*)bias, out, N, K / 8);
}

}  // namespace svln

using namespace svln;

extern "C" {

int svln_gemv(const void* W, const void* x, const void* bias, void* out, int M, int N, int K, int out_is_f32,
              svln_stream_t stream) {
    SVLN_REQUIRE(M >= 1 && M <= 4, SVLN_ERR_RANGE, "svln_gemv: M=%d outside [1,4] (larger batches are GEMMs)", M);
    SVLN_REQUIRE(N > 0 && K > 0 && K % 8 == 0, SVLN_ERR_INVALID, "svln_gemv: bad shape N=%d K=%d", N, K);
    SVLN_REQUIRE(W && x && out, SVLN_ERR_INVALID, "svln_gemv: null pointer");
    SVLN_REQUIRE((((uintptr_t)W | (uintptr_t)x) & 15) == 0, SVLN_ERR_INVALID, "svln_gemv: unaligned operand");
    cudaStream_t st = (cudaStream_t)stream;
    // one W row per warp unless that would leave fewer than ~4 CTAs per SM worth of warps
    const bool wide = (int64_t)N >= 16 * 8 * (int64_t)sm_count();
    switch (M) {
        case 1: wide ? launch_gemv<1, 2>(W, x, bias, out, N, K, out_is_f32, st) : launch_gemv<1, 1>(W, x, bias, out, N, K, out_is_f32, st); break;
        case 2: wide ? launch_gemv<2, 2>(W, x, bias, out, N, K, out_is_f32, st) : launch_gemv<2, 1>(W, x, bias, out, N, K, out_is_f32, st); break;
        case 3: wide ? launch_gemv<3, 2>(W, x, bias, out, N, K, out_is_f32, st) : launch_gemv<3, 1>(W, x, bias, out, N, K, out_is_f32, st); break;
        default: wide ? launch_gemv<4, 2>(W, x, bias, out, N, K, out_is_f32, st) : launch_gemv<4, 1>(W, x, bias, out, N, K, out_is_f32, st); break;
    }
    return check_launch("gemv_kernel");
}

}  // extern "C"

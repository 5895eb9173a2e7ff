// Small fused ops around the Qwen2 linears (which stay cuBLAS): RMSNorm (+ residual add),
// SwiGLU product, greedy argmax.  Rounding points follow HF transformers' bf16 modules
// (Qwen2RMSNorm, Qwen2MLP; reached from stream_video_vln.py:340) so that logits stay within the
// stated tolerance of the CPU oracle.
#include "svln_common.cuh"

namespace svln {

constexpr int NORM_THREADS = 256;
constexpr int NORM_MAX_I4 = 4;  // dim <= 256*4*8 = 8192

template <bool ADD>
__global__ void __launch_bounds__(NORM_THREADS)
rmsnorm_kernel(int4* __restrict__ x, const int4* __restrict__ r, const int4* __restrict__ w, int dim_i4,
               float inv_dim, float eps, int4* __restrict__ out) {
    __shared__ float s_part[NORM_THREADS / 32];
    const int64_t row = blockIdx.x;
    int4* xr = x + row * dim_i4;
    const int4* rr = ADD ? r + row * dim_i4 : nullptr;
    int4 v[NORM_MAX_I4];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < NORM_MAX_I4; ++q) {
        const int c = q * NORM_THREADS + threadIdx.x;
        if (c < dim_i4) {
            v[q] = xr[c];
            uint32_t* u = reinterpret_cast<uint32_t*>(&v[q]);
            if (ADD) {
                const int4 rv = rr[c];
                const uint32_t* ru = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a0 = bf16_bits_to_f32(u[k] & 0xFFFFu) + bf16_bits_to_f32(ru[k] & 0xFFFFu);
                    const float a1 = bf16_bits_to_f32(u[k] >> 16) + bf16_bits_to_f32(ru[k] >> 16);
                    u[k] = f32_to_bf16_bits(a0) | (f32_to_bf16_bits(a1) << 16);
                }
                xr[c] = v[q];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a0 = bf16_bits_to_f32(u[k] & 0xFFFFu), a1 = bf16_bits_to_f32(u[k] >> 16);
                ss += a0 * a0 + a1 * a1;
            }
        }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_THREADS / 32; ++i) tot += s_part[i];
    const float rstd = rsqrtf(tot * inv_dim + eps);
    int4* orow = out + row * dim_i4;
#pragma unroll
    for (int q = 0; q < NORM_MAX_I4; ++q) {
        const int c = q * NORM_THREADS + threadIdx.x;
        if (c < dim_i4) {
            const int4 wv = __ldg(w + c);
            const uint32_t* wu = reinterpret_cast<const uint32_t*>(&wv);
            const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[q]);
            int4 o;
            uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // HF: weight * (x_fp32 * rstd).to(bf16), product rounded to bf16 again
                const float n0 = round_bf16(bf16_bits_to_f32(u[k] & 0xFFFFu) * rstd);
                const float n1 = round_bf16(bf16_bits_to_f32(u[k] >> 16) * rstd);
                const float y0 = bf16_bits_to_f32(wu[k] & 0xFFFFu) * n0;
                const float y1 = bf16_bits_to_f32(wu[k] >> 16) * n1;
                ou[k] = f32_to_bf16_bits(y0) | (f32_to_bf16_bits(y1) << 16);
            }
            orow[c] = o;
        }
    }
}

// 4 independent (gate, up) chunk pairs per thread, every load issued before the first use: silu is a dependent chain of
// ~30 instructions per element (precise expf and IEEE division, to round exactly like torch's bf16 silu), so with one
// chunk per thread the kernel was latency/issue-bound at 4 TB/s (13.5 ms of a 64-stream packed turn, round 2 profile).
constexpr int SW_UNROLL = 4;
__global__ void __launch_bounds__(256)
swiglu_kernel(const int4* __restrict__ gate_up, int64_t n_rows, int inter_i4, int4* __restrict__ out) {
    const int64_t row = blockIdx.x;       // rows on x: a packed window-start prefill has > 65535 of them
    const int4* grow = gate_up + row * 2 * inter_i4;
    int4* orow = out + row * inter_i4;
    const int c0 = (blockIdx.y * SW_UNROLL) * 256 + threadIdx.x;
    int4 g[SW_UNROLL], u[SW_UNROLL];
#pragma unroll
    for (int i = 0; i < SW_UNROLL; ++i) {
        const int c = c0 + i * 256;
        if (c < inter_i4) {
            g[i] = ld_stream(grow + c);
            u[i] = ld_stream(grow + inter_i4 + c);
        }
    }
#pragma unroll
    for (int i = 0; i < SW_UNROLL; ++i) {
        const int c = c0 + i * 256;
        if (c < inter_i4) {
            const uint32_t* gu = reinterpret_cast<const uint32_t*>(&g[i]);
            const uint32_t* uu = reinterpret_cast<const uint32_t*>(&u[i]);
            int4 o;
            uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t res = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float gv = bf16_bits_to_f32((gu[k] >> (16 * h)) & 0xFFFFu);
                    const float uv = bf16_bits_to_f32((uu[k] >> (16 * h)) & 0xFFFFu);
                    const float act = round_bf16(gv / (1.0f + expf(-gv)));  // silu in fp32, rounded like torch
                    res |= f32_to_bf16_bits(act * uv) << (16 * h);
                }
                ou[k] = res;
            }
            st_stream(orow + c, o);
        }
    }
}

template <bool F32>
__global__ void __launch_bounds__(1024)
argmax_kernel(const void* __restrict__ logits, int64_t vocab, int64_t* __restrict__ out_idx) {
    __shared__ float s_v[32];
    __shared__ int64_t s_i[32];
    const int64_t row = blockIdx.x;
    float best = -INFINITY;
    int64_t bi = vocab;  // sentinel: larger than any index
    for (int64_t i = threadIdx.x; i < vocab; i += blockDim.x) {
        float v;
        if (F32) v = reinterpret_cast<const float*>(logits)[row * vocab + i];
        else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(logits)[row * vocab + i]);
        // NaN compares false everywhere: torch.argmax treats NaN as max; random-init logits have none
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = s_v[threadIdx.x];
        bi = s_i[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) out_idx[row] = bi >= vocab ? 0 : bi;
    }
}

}  // namespace svln

using namespace svln;

extern "C" {

static int norm_args_ok(const void* x, const void* w, int64_t n_rows, int dim, void* out, const char* who) {
    SVLN_REQUIRE(n_rows >= 0 && dim > 0 && dim % 8 == 0 && dim <= NORM_THREADS * NORM_MAX_I4 * 8,
                 SVLN_ERR_INVALID, "%s: dim %d must be a multiple of 8 and <= %d", who, dim,
                 NORM_THREADS * NORM_MAX_I4 * 8);
    SVLN_REQUIRE(n_rows == 0 || (x && w && out), SVLN_ERR_INVALID, "%s: null pointer", who);
    SVLN_REQUIRE(n_rows <= 0x7FFFFFFF, SVLN_ERR_RANGE, "%s: too many rows", who);
    return SVLN_OK;
}

int svln_rmsnorm(const void* x, const void* weight, int64_t n_rows, int dim, float eps, void* out,
                 svln_stream_t stream) {
    int rc = norm_args_ok(x, weight, n_rows, dim, out, "svln_rmsnorm");
    if (rc || n_rows == 0) return rc;
    rmsnorm_kernel<false><<<(unsigned)n_rows, NORM_THREADS, 0, (cudaStream_t)stream>>>(
        (int4*)x, nullptr, (const int4*)weight, dim / 8, 1.0f / (float)dim, eps, (int4*)out);
    return check_launch("rmsnorm_kernel");
}

int svln_add_rmsnorm(void* x, const void* r, const void* weight, int64_t n_rows, int dim, float eps, void* out,
                     svln_stream_t stream) {
    int rc = norm_args_ok(x, weight, n_rows, dim, out, "svln_add_rmsnorm");
    if (rc || n_rows == 0) return rc;
    SVLN_REQUIRE(r, SVLN_ERR_INVALID, "svln_add_rmsnorm: null residual");
    rmsnorm_kernel<true><<<(unsigned)n_rows, NORM_THREADS, 0, (cudaStream_t)stream>>>(
        (int4*)x, (const int4*)r, (const int4*)weight, dim / 8, 1.0f / (float)dim, eps, (int4*)out);
    return check_launch("add_rmsnorm_kernel");
}

int svln_swiglu(const void* gate_up, int64_t n_rows, int inter, void* out, svln_stream_t stream) {
    SVLN_REQUIRE(n_rows >= 0 && inter > 0 && inter % 8 == 0, SVLN_ERR_INVALID, "svln_swiglu: bad shape");
    if (n_rows == 0) return SVLN_OK;
    SVLN_REQUIRE(gate_up && out, SVLN_ERR_INVALID, "svln_swiglu: null pointer");
    SVLN_REQUIRE(n_rows <= 0x7FFFFFFF, SVLN_ERR_RANGE, "svln_swiglu: %lld rows exceed the grid", (long long)n_rows);
    const int inter_i4 = inter / 8;
    dim3 grid((unsigned)n_rows, (unsigned)((inter_i4 + SW_UNROLL * 256 - 1) / (SW_UNROLL * 256)));
    swiglu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const int4*)gate_up, n_rows, inter_i4, (int4*)out);
    return check_launch("swiglu_kernel");
}

int svln_argmax(const void* logits, int is_f32, int64_t n_rows, int64_t vocab, int64_t* out_idx,
                svln_stream_t stream) {
    SVLN_REQUIRE(n_rows >= 0 && vocab > 0, SVLN_ERR_INVALID, "svln_argmax: bad shape");
    if (n_rows == 0) return SVLN_OK;
    SVLN_REQUIRE(logits && out_idx, SVLN_ERR_INVALID, "svln_argmax: null pointer");
    if (is_f32)
        argmax_kernel<true><<<(unsigned)n_rows, 1024, 0, (cudaStream_t)stream>>>(logits, vocab, out_idx);
    else
        argmax_kernel<false><<<(unsigned)n_rows, 1024, 0, (cudaStream_t)stream>>>(logits, vocab, out_idx);
    return check_launch("argmax_kernel");
}

}  // extern "C"

// Bilinear 2-D token pooling of projected vision features: [F, gin*gin, D] -> [F, gout*gout, D] bf16.
// Replaces StreamVLNForCausalLM.get_2dPool (streamvln/model/stream_video_vln.py:53-73, mode "bilinear":
// view F x gin x gin x D -> permute to NCHW -> F.interpolate(size=ceil(gin/2), mode="bilinear") -> permute back).
// The permutes only exist to feed an NCHW op; with D = 3584 channels torch's upsample_bilinear2d kernel takes
// 2.06 ms per frame on a B200 (ncu launch list), 70 % of the whole vision front end.  Here a thread owns 8
// channels of one output token and reads the four source tokens with 128-bit loads: 5.2 MB in, 1.4 MB out.
// Arithmetic follows ATen's (align_corners = False, fp32 opmath, scale = gin / gout as a float division,
// src = max(0, fma(scale, dst + 0.5, -0.5)), lambda1 = src - floor(src)):
// out = l0h * (l0w * x00 + l1w * x01) + l1h * (l0w * x10 + l1w * x11), rounded to bf16 once.  Against CPU
// F.interpolate (what the parity tests compare with) this differs by one bf16 ulp on ~5e-5 of the elements (ATen's exact contraction
// pattern is not reproducible from outside); the test bounds both the rate and the size of those flips.
#include "svln_common.cuh"

namespace svln {

__device__ __forceinline__ void pool_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
    float src = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);   // ATen's CPU build contracts scale * (dst + 0.5) - 0.5
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
    l0 = __fsub_rn(1.f, l1);
}

__global__ void __launch_bounds__(256)
pool_bilinear_kernel(const int4* __restrict__ feat, int gin, int gout, int d8, float scale, int4* __restrict__ out) {
    const int tok = blockIdx.x;                  // output token of frame blockIdx.y
    const int oy = tok / gout, ox = tok - oy * gout;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    pool_src(oy, scale, gin, y0, y1, ly0, ly1);
    pool_src(ox, scale, gin, x0, x1, lx0, lx1);
    const int64_t frame = (int64_t)blockIdx.y * gin * gin;
    const int4* r00 = feat + (frame + y0 * gin + x0) * d8;
    const int4* r01 = feat + (frame + y0 * gin + x1) * d8;
    const int4* r10 = feat + (frame + y1 * gin + x0) * d8;
    const int4* r11 = feat + (frame + y1 * gin + x1) * d8;
    int4* o = out + ((int64_t)blockIdx.y * gout * gout + tok) * d8;
    for (int c = threadIdx.x; c < d8; c += blockDim.x) {
        const int4 a = __ldg(r00 + c), b = __ldg(r01 + c), cc = __ldg(r10 + c), d = __ldg(r11 + c);
        const uint32_t* ua = reinterpret_cast<const uint32_t*>(&a);
        const uint32_t* ub = reinterpret_cast<const uint32_t*>(&b);
        const uint32_t* uc = reinterpret_cast<const uint32_t*>(&cc);
        const uint32_t* ud = reinterpret_cast<const uint32_t*>(&d);
        int4 res;
        uint32_t* ur = reinterpret_cast<uint32_t*>(&res);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t packed = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int sft = 16 * h;
                const float v00 = bf16_bits_to_f32((ua[k] >> sft) & 0xFFFFu), v01 = bf16_bits_to_f32((ub[k] >> sft) & 0xFFFFu);
                const float v10 = bf16_bits_to_f32((uc[k] >> sft) & 0xFFFFu), v11 = bf16_bits_to_f32((ud[k] >> sft) & 0xFFFFu);
                const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
                const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
                const float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
                packed |= f32_to_bf16_bits(v) << sft;
            }
            ur[k] = packed;
        }
        o[c] = res;
    }
}

}  // namespace svln

using namespace svln;

extern "C" {

int svln_pool_bilinear(const void* feat, int n_frames, int grid_in, int grid_out, int dim, void* out,
                       svln_stream_t stream) {
    SVLN_REQUIRE(n_frames >= 0 && grid_in > 0 && grid_out > 0 && dim > 0 && dim % 8 == 0, SVLN_ERR_INVALID,
                 "svln_pool_bilinear: bad shape frames=%d grid %d -> %d dim=%d", n_frames, grid_in, grid_out, dim);
    if (n_frames == 0) return SVLN_OK;
    SVLN_REQUIRE(feat && out, SVLN_ERR_INVALID, "svln_pool_bilinear: null pointer");
    SVLN_REQUIRE(n_frames <= 65535, SVLN_ERR_RANGE, "svln_pool_bilinear: too many frames");
    const float scale = (float)grid_in / (float)grid_out;
    dim3 grid(grid_out * grid_out, n_frames);
    pool_bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const int4*)feat, grid_in, grid_out, dim / 8, scale, (int4*)out);
    return check_launch("pool_bilinear_kernel");
}

}  // extern "C"

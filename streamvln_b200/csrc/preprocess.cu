// RGB frame preprocessing on the GPU: the reference's SigLipImageProcessor
// (llava/model/multimodal_encoder/siglip_encoder.py:34-66: PIL BICUBIC squash to 384x384, rescale 1/255,
// normalise mean = std = 0.5, channels first) restated as two integer resampling passes that reproduce
// Pillow's 8-bit ImagingResample bit for bit (22-bit fixed-point coefficients computed on the host exactly as
// Pillow does, horizontal pass then vertical pass, each rounded and clipped to uint8) plus the float32
// normalisation.  Replaces ~5 ms of host PIL/numpy work per model turn and halves the H2D bytes (uint8 in).
#include "svln_common.cuh"

namespace svln {

constexpr int PP_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int acc) {
    const int v = acc >> PP_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: in [n,H,W,3] u8 -> tmp [n,H,OW,3] u8
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ in, int H, int W, int OW, const int32_t* __restrict__ xmin,
                  const int32_t* __restrict__ cnt, const int32_t* __restrict__ kk, int ksize, uint8_t* __restrict__ tmp,
                  int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*H*OW
    if (i >= total) return;
    const int xo = (int)(i % OW);
    const int64_t row = i / OW;  // n*H + y
    const uint8_t* src = in + (row * W + __ldg(xmin + xo)) * 3;
    const int32_t* k = kk + (int64_t)xo * ksize;
    const int c = __ldg(cnt + xo);
    int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < c; ++x) {
        const int w = __ldg(k + x);
        a0 += (int)src[x * 3 + 0] * w;
        a1 += (int)src[x * 3 + 1] * w;
        a2 += (int)src[x * 3 + 2] * w;
    }
    uint8_t* d = tmp + i * 3;
    d[0] = (uint8_t)clip8(a0); d[1] = (uint8_t)clip8(a1); d[2] = (uint8_t)clip8(a2);
}

// vertical pass + normalisation: tmp [n,H,OW,3] u8 -> out [n,3,OH,OW] f32
__global__ void __launch_bounds__(256)
resample_v_norm_kernel(const uint8_t* __restrict__ tmp, int H, int OW, int OH, const int32_t* __restrict__ ymin,
                       const int32_t* __restrict__ cnt, const int32_t* __restrict__ kk, int ksize, float* __restrict__ out,
                       int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*OH*OW
    if (i >= total) return;
    const int xo = (int)(i % OW);
    const int yo = (int)((i / OW) % OH);
    const int64_t n = i / ((int64_t)OW * OH);
    const uint8_t* src = tmp + ((n * H + __ldg(ymin + yo)) * OW + xo) * 3;
    const int32_t* k = kk + (int64_t)yo * ksize;
    const int c = __ldg(cnt + yo);
    int a0 = 1 << (PP_PRECISION_BITS - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < c; ++y) {
        const int w = __ldg(k + y);
        const uint8_t* p = src + (int64_t)y * OW * 3;
        a0 += (int)p[0] * w; a1 += (int)p[1] * w; a2 += (int)p[2] * w;
    }
    const int64_t plane = (int64_t)OH * OW;
    float* o = out + n * 3 * plane + (int64_t)yo * OW + xo;
    // float32(u8) * float32(1/255), then (x - 0.5) / 0.5: the host processor's float32 arithmetic
    const float s = 0.003921568859368563f;
    o[0] = __fmul_rn(__fsub_rn(__fmul_rn((float)clip8(a0), s), 0.5f), 2.0f);
    o[plane] = __fmul_rn(__fsub_rn(__fmul_rn((float)clip8(a1), s), 0.5f), 2.0f);
    o[2 * plane] = __fmul_rn(__fsub_rn(__fmul_rn((float)clip8(a2), s), 0.5f), 2.0f);
}

// Depth frame: Habitat's [0,1] float depth -> metres on the out x out processor grid, as the reference's eval loop does
// on the host (streamvln_eval.py:254-256 scale to metres and millimetres, :276 uint16 truncation, :126-135 Pillow NEAREST
// squash + / 1000).  Pillow's nearest source index is (int)((i + 0.5) * in / out) in double arithmetic; the uint16 value
// is divided by 1000 in double and rounded to float32 once, like numpy's uint16 / int -> float64 followed by .float().
__global__ void __launch_bounds__(256)
depth_nearest_kernel(const float* __restrict__ in, int H, int W, int OH, int OW, float span, float min_depth,
                     float* __restrict__ out, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*OH*OW
    if (i >= total) return;
    const int xo = (int)(i % OW);
    const int yo = (int)((i / OW) % OH);
    const int64_t n = i / ((int64_t)OW * OH);
    const int ys = (int)(((double)yo + 0.5) * ((double)H / (double)OH));
    const int xs = (int)(((double)xo + 0.5) * ((double)W / (double)OW));
    const float d = in[(n * H + ys) * W + xs];
    const float mm = __fmul_rn(__fadd_rn(__fmul_rn(d, span), min_depth), 1000.0f);   // float32, as numpy computes it
    const unsigned short q = (unsigned short)(int)mm;                                 // .astype(np.uint16): truncation (depth >= 0)
    out[i] = (float)((double)q / 1000.0);
}

}  // namespace svln

using namespace svln;

extern "C" {

size_t svln_preprocess_workspace_bytes(int n, int H, int out_w) {
    if (n <= 0 || H <= 0 || out_w <= 0) return 0;
    return (size_t)n * H * out_w * 3;
}

int svln_preprocess_rgb(const uint8_t* rgb, int n, int H, int W, int out_h, int out_w, const int32_t* xmin,
                        const int32_t* xcnt, const int32_t* xkk, int xksize, const int32_t* ymin, const int32_t* ycnt,
                        const int32_t* ykk, int yksize, void* workspace, size_t workspace_bytes, float* out,
                        svln_stream_t stream) {
    SVLN_REQUIRE(n >= 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && xksize > 0 && yksize > 0, SVLN_ERR_INVALID,
                 "svln_preprocess_rgb: bad shape");
    if (n == 0) return SVLN_OK;
    SVLN_REQUIRE(rgb && xmin && xcnt && xkk && ymin && ycnt && ykk && out, SVLN_ERR_INVALID, "svln_preprocess_rgb: null pointer");
    const size_t need = (size_t)n * H * out_w * 3;
    SVLN_REQUIRE(workspace && workspace_bytes >= need, SVLN_ERR_NOMEM, "svln_preprocess_rgb: workspace %zu < %zu bytes",
                 workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t t1 = (int64_t)n * H * out_w, t2 = (int64_t)n * out_h * out_w;
    resample_h_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, st>>>(rgb, H, W, out_w, xmin, xcnt, xkk, xksize,
                                                                    (uint8_t*)workspace, t1);
    resample_v_norm_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>((const uint8_t*)workspace, H, out_w, out_h, ymin, ycnt,
                                                                         ykk, yksize, out, t2);
    return check_launch("preprocess kernels");
}

int svln_preprocess_depth(const float* depth01, int n, int H, int W, int out_h, int out_w, float min_depth,
                          float max_depth, float* out, svln_stream_t stream) {
    SVLN_REQUIRE(n >= 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && max_depth > min_depth && min_depth >= 0.f, SVLN_ERR_INVALID,
                 "svln_preprocess_depth: bad arguments");
    if (n == 0) return SVLN_OK;
    SVLN_REQUIRE(depth01 && out, SVLN_ERR_INVALID, "svln_preprocess_depth: null pointer");
    SVLN_REQUIRE(max_depth * 1000.f < 65536.f, SVLN_ERR_RANGE, "svln_preprocess_depth: max_depth %.1f m overflows the uint16 millimetres", max_depth);
    const int64_t total = (int64_t)n * out_h * out_w;
    depth_nearest_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(depth01, H, W, out_h, out_w,
                                                                                           max_depth - min_depth, min_depth, out, total);
    return check_launch("depth_nearest_kernel");
}

}  // extern "C"

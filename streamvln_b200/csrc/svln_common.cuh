// Shared helpers for libsvln (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/svln.h"

namespace svln {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();
// true the first time it is called for `*mask` on the CURRENT device: function attributes (opt-in shared memory) are
// per device, so a per-process `static bool` would leave every device after the first at the 48 KB default
bool first_use_on_device(unsigned long long* mask);

#define SVLN_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            svln::set_error(__VA_ARGS__);  \
            return (code);                 \
        }                                  \
    } while (0)

#define SVLN_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            svln::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
            return SVLN_ERR_CUDA;                                                    \
        }                                                                            \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// 128-bit streaming loads/stores that do not pollute L1 (rows are touched once).
__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w));
}
// plain (coherent) 128-bit load for data written earlier by other CTAs of the same kernel
__device__ __forceinline__ int4 ld_cg(const int4* p) {
    int4 r;
    asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float bf16_bits_to_f32(uint32_t b16) {
    return __uint_as_float(b16 << 16);
}
// round-to-nearest-even fp32 -> bf16 bits (matches torch / numpy-ml_dtypes rounding, NaN kept)
__device__ __forceinline__ uint32_t f32_to_bf16_bits(float f) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float round_bf16(float f) {
    return bf16_bits_to_f32(f32_to_bf16_bits(f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}


// Loads of data that ANOTHER CTA of the same launch produced (read through L2, never L1).  asm volatile: the compiler
// may neither hoist them out of a loop nor reuse an earlier result - `__ldcg` is a plain (non-volatile) asm in the CUDA
// headers, and nvcc moved __ldcg(hb0 + row) of the fused decode stack out of its layer loop (profiles/
// r2_decode_stack_experiments.md): garbage that came and went with unrelated code changes.
__device__ __forceinline__ int4 ld_cg_v4(const int4* p) {
    int4 r;
    asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ld_cg_f(const float* p) {
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ld_cg_bf16(const __nv_bfloat16* p) {
    unsigned short r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p) : "memory");
    return __uint_as_float((uint32_t)r << 16);
}

}  // namespace svln

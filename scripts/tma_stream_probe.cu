// Probe (run on the GPU box): how fast can ONE producer thread per SM stream a weight matrix from HBM into a shared-memory
// mbarrier ring with (a) 2-D tensor TMA boxes, (b) 1-D bulk copies of contiguous chunks (a pre-tiled weight layout)?
// The consumer only waits for the stage and releases it, so the number is the copy path's ceiling for the decode stack.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_probe scripts/tma_stream_probe.cu -lcuda && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_spin(uint64_t* b, uint32_t parity) {      // non-blocking test_wait poll
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
constexpr int MAX_STAGES = 32;

// mode 0: 2-D tensor boxes (box_cols x box_rows); mode 1: 1-D bulk, contiguous chunks of stage_bytes
__global__ void __launch_bounds__(64)
stream_kernel(const __grid_constant__ CUtensorMap map, const unsigned char* W, int mode, int rows_per_cta, int K, int box_rows,
              int box_cols, int stage_bytes, int n_stages, int passes, unsigned long long bytes_per_cta, unsigned* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * rows_per_cta;
    const int n_rb = rows_per_cta / box_rows, n_kb = K / box_cols;
    const long long n_items = mode == 0 ? (long long)n_rb * n_kb : (long long)(bytes_per_cta / stage_bytes);
    if (warp == 0 && lane == 0) {
        long long it = 0;
        for (int p = 0; p < passes; ++p)
            for (long long i = 0; i < n_items; ++i, ++it) {
                const int s = (int)(it % n_stages);
                const uint32_t ph = (uint32_t)((it / n_stages) & 1);
                if (it >= n_stages) mbar_wait(&empty[s], ph ^ 1);
                mbar_expect(&full[s], (uint32_t)stage_bytes);
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
                if (mode == 0) {
                    const int rb = (int)(i / n_kb), kb = (int)(i % n_kb);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(dst), "l"(&map), "r"(kb * box_cols), "r"(row0 + rb * box_rows), "r"(smem_u32(&full[s])) : "memory");
                } else {
                    const unsigned char* src = W + (size_t)blockIdx.x * bytes_per_cta + (size_t)i * stage_bytes;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(src), "r"(stage_bytes), "r"(smem_u32(&full[s])) : "memory");
                }
            }
    } else if (warp == 1) {
        unsigned acc = 0;
        const long long total = n_items * passes;
        for (long long it = 0; it < total; ++it) {
            const int s = (int)(it % n_stages);
            mbar_wait(&full[s], (uint32_t)((it / n_stages) & 1));
            acc += *(volatile unsigned*)(smem + (size_t)s * stage_bytes + lane * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

// mode 2: stage = nbox swizzle-128B boxes (64 cols x R rows) consecutive in k; 4 consumer warps LDS.128 the whole stage
__global__ void __launch_bounds__(160)
boxes_kernel(const __grid_constant__ CUtensorMap map, int rows_per_cta, int K, int R, int nbox, int n_stages, int passes, unsigned* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    const int box_bytes = R * 128, stage_bytes = box_bytes * nbox;
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * rows_per_cta;
    const int n_rb = rows_per_cta / R, n_ks = K / (64 * nbox);
    const long long n_items = (long long)n_rb * n_ks, total = n_items * passes;
    if (warp == 4) {
        if (lane == 0) {
            long long it = 0;
            for (int p = 0; p < passes; ++p)
                for (long long i = 0; i < n_items; ++i, ++it) {
                    const int s = (int)(it % n_stages);
                    if (it >= n_stages) mbar_wait(&empty[s], (uint32_t)(((it / n_stages) & 1) ^ 1));
                    mbar_expect(&full[s], (uint32_t)stage_bytes);
                    const int rb = (int)(i / n_ks), ks = (int)(i % n_ks);
                    for (int b = 0; b < nbox; ++b) {
                        const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes + (size_t)b * box_bytes);
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                     ::"r"(dst), "l"(&map), "r"((ks * nbox + b) * 64), "r"(row0 + rb * R), "r"(smem_u32(&full[s])) : "memory");
                    }
                }
        }
    } else {
        unsigned acc = 0;
        for (long long it = 0; it < total; ++it) {
            const int s = (int)(it % n_stages);
            mbar_wait(&full[s], (uint32_t)((it / n_stages) & 1));
            const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
            for (int o = (warp * 32 + lane) * 16; o < stage_bytes; o += 128 * 16) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + o));
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

// mode 3: the decode stack's packed feed: 1-D copies of stage_bytes into n_stages slots of slot_bytes, 8 consumer warps that
// each read 1/8 of the stage and arrive on the slot's empty barrier (count 8); one producer thread in a ninth warp
template <int NS, int HINT>
__global__ void __launch_bounds__(288)
feed_kernel(const unsigned char* W, unsigned bytes_per_cta, int stage_bytes, int slot_bytes, int passes, unsigned* sink, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[NS], empty[NS];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_items = bytes_per_cta / stage_bytes;
    const long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0) {
            unsigned it = 0;
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            const unsigned char* base = W + (size_t)blockIdx.x * bytes_per_cta;
            for (int p = 0; p < passes; ++p) {
                const unsigned char* src = base;
                for (unsigned i = 0; i < n_items; ++i, ++it, src += stage_bytes) {
                    const unsigned s = it % NS;
                    if (it >= NS) mbar_wait(&empty[s], ((it / NS) & 1u) ^ 1u);
                    mbar_expect(&full[s], (uint32_t)stage_bytes);
                    if (HINT)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                     ::"r"(smem_u32(smem + s * slot_bytes)), "l"(src), "r"(stage_bytes), "r"(smem_u32(&full[s])), "l"(pol) : "memory");
                    else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(smem + s * slot_bytes)), "l"(src), "r"(stage_bytes), "r"(smem_u32(&full[s])) : "memory");
                }
            }
        }
    } else {
        unsigned acc = 0;
        const unsigned total = n_items * passes;
        for (unsigned it = 0; it < total; ++it) {
            const unsigned s = it % NS;
            mbar_wait(&full[s], (it / NS) & 1u);
            const uint32_t base = smem_u32(smem + s * slot_bytes);
            for (int o = (warp * 32 + lane) * 16; o < stage_bytes; o += 256 * 16) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + o));
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
        if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
    }
}

template <int NS, int HINT>
static void run_feed(const unsigned char* W, size_t bytes, int stage, int passes, unsigned* sink, long long* cyc, unsigned char* flush,
                     cudaEvent_t e0, cudaEvent_t e1) {
    CK(cudaFuncSetAttribute(feed_kernel<NS, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const unsigned per_cta = (unsigned)((bytes / 148) / stage) * stage;
    const size_t smem = (size_t)stage * NS;
    if (smem > 220 * 1024) return;
    float best = 1e9f; long long bc = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemsetAsync(flush, rep, 256u << 20));
        CK(cudaEventRecord(e0));
        feed_kernel<NS, HINT><<<148, 288, smem>>>(W, per_cta, stage, stage, passes, sink, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        long long hc; CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
        if (ms < best) { best = ms; bc = hc; }
    }
    CK(cudaGetLastError());
    printf("feed%s: stage %5d B x%2d slots (%3zu KB), %d passes  %7.1f us  %6.0f GB/s  %5.1f B/clk/SM, %5.0f cycles per stage (CTA 0)\n", HINT ? "[evict_first]" : "", stage, NS, smem >> 10, passes,
           best * 1e3, (double)per_cta * 148 * passes / best / 1e6, (double)per_cta * passes / bc, (double)bc / ((double)per_cta * passes / stage));
    fflush(stdout);
}

// control: plain 128-bit loads, 8 in flight per thread
__global__ void __launch_bounds__(512) ldg_kernel(const uint4* W, size_t n_vec, unsigned* sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n_vec; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(W + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    const int N = 37888, K = 3584;                       // gate/up of Qwen2-7B, 271.6 MB
    const size_t bytes = (size_t)N * K * 2;
    unsigned char* W; unsigned* sink;
    CK(cudaMalloc(&W, bytes)); CK(cudaMemset(W, 1, bytes)); CK(cudaMalloc(&sink, 4));
    unsigned char* flush; CK(cudaMalloc(&flush, 256u << 20));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int ctas : {148 * 2, 148 * 4}) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemsetAsync(flush, rep, 256u << 20));
            CK(cudaEventRecord(e0));
            ldg_kernel<<<ctas, 512>>>((const uint4*)W, bytes / 16, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("control: ld.global.nc.v4 x8, %d CTAs x 512      %7.1f us  %6.0f GB/s\n", ctas, best * 1e3, bytes / best / 1e6); fflush(stdout);
    }
    struct Cfg { int mode, box_cols, box_rows, swz, stages; const char* name; int ctas_per_sm; int passes; };
    const Cfg cfgs[] = {
        {0, 64, 128, 1, 8, "2-D 64x128 swizzle128 (16 KB) x8", 1, 1},
        {0, 256, 64, 0, 6, "2-D 256x64 no swizzle (32 KB) x6", 1, 1},
        {0, 256, 64, 0, 6, "2-D 256x64 (32 KB) x6, 4 passes", 1, 4},
        {0, 256, 64, 0, 3, "2-D 256x64 (32 KB) x3, 2 CTAs/SM", 2, 1},
        {0, 256, 32, 0, 3, "2-D 256x32 (16 KB) x3, 4 CTAs/SM", 4, 1},
        {1, 0, 0, 0, 6, "1-D contiguous 32 KB x6", 1, 1},
        {1, 0, 0, 0, 6, "1-D contiguous 32 KB x6, 4 passes", 1, 4},
        {1, 0, 0, 0, 3, "1-D contiguous 32 KB x3, 2 CTAs/SM", 2, 1},
        {1, 0, 0, 0, 3, "1-D contiguous 16 KB x3, 4 CTAs/SM", 4, 1},
    };
    const int onedim_bytes[] = {32768, 32768, 32768, 16384};
    int od = 0;
    for (const Cfg& c : cfgs) {
        CUtensorMap map = {};
        int stage_bytes;
        if (c.mode == 0) {
            cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
            cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
            cuuint32_t box[2] = {(cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows};
            cuuint32_t es[2] = {1, 1};
            CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", c.name, (int)r); continue; }
            stage_bytes = c.box_cols * c.box_rows * 2;
        } else {
            stage_bytes = onedim_bytes[od++];
        }
        const int rows_per_cta = N / (148 * c.ctas_per_sm);
        const unsigned long long per_cta = (unsigned long long)rows_per_cta * K * 2;
        const size_t smem = (size_t)stage_bytes * c.stages;
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemsetAsync(flush, rep, 256u << 20));
            CK(cudaEventRecord(e0));
            stream_kernel<<<148 * c.ctas_per_sm, 64, smem>>>(map, W, c.mode, rows_per_cta, K, c.box_rows ? c.box_rows : 1, c.box_cols ? c.box_cols : 1,
                                             stage_bytes, c.stages, c.passes, per_cta, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-40s %7.1f us  %6.0f GB/s\n", c.name, best * 1e3, bytes * c.passes / best / 1e6); fflush(stdout);
    }
    CK(cudaFuncSetAttribute(boxes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct BCfg { int N, K, R, nbox, stages, passes; const char* name; };
    const BCfg bc[] = {
        {37888, 3584, 256, 1, 6, 1, "gate/up 256 rows x 64 (32 KB) x6"},
        {37888, 3584, 256, 1, 6, 4, "gate/up 256 rows x 64 (32 KB) x6, 4 passes"},
        {37888, 3584, 128, 2, 6, 4, "gate/up 128 rows x 2 boxes (32 KB) x6, 4 passes"},
        {37888, 3584, 64, 4, 6, 4, "gate/up 64 rows x 4 boxes (32 KB) x6, 4 passes"},
        {3552, 18944, 24, 8, 8, 1, "down 24 rows x 8 boxes (24 KB) x8"},
        {3552, 18944, 24, 8, 8, 8, "down 24 rows x 8 boxes (24 KB) x8, 8 passes"},
        {3552, 18944, 24, 4, 16, 8, "down 24 rows x 4 boxes (12 KB) x16, 8 passes"},
        {4736, 3584, 32, 8, 6, 16, "qkv 32 rows x 8 boxes (32 KB) x6, 16 passes"},
        {4736, 3584, 16, 8, 12, 16, "qkv 16 rows x 8 boxes (16 KB) x12, 16 passes"},
    };
    for (const BCfg& c : bc) {
        CUtensorMap map = {};
        cuuint64_t gdim[2] = {(cuuint64_t)c.K, (cuuint64_t)c.N};
        cuuint64_t gstride[1] = {(cuuint64_t)c.K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)c.R};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-50s encode failed %d\n", c.name, (int)r); continue; }
        const size_t smem = (size_t)c.R * 128 * c.nbox * c.stages;
        const size_t mbytes = (size_t)c.N * c.K * 2;
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemsetAsync(flush, rep, 256u << 20));
            CK(cudaEventRecord(e0));
            boxes_kernel<<<148, 160, smem>>>(map, c.N / 148, c.K, c.R, c.nbox, c.stages, c.passes, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-50s %7.1f us  %6.0f GB/s\n", c.name, best * 1e3, mbytes * c.passes / best / 1e6); fflush(stdout);
    }
    long long* cyc; CK(cudaMalloc(&cyc, 8));
    // source alignment: stages of 24960 / 32000 B keep every second copy 64-byte (not 128-byte) aligned
    for (int stage : {24576, 24960, 25088, 32000, 32768}) {
        run_feed<4, 1>(W, bytes, stage, 4, sink, cyc, flush, e0, e1);
    }
    return 0;
}

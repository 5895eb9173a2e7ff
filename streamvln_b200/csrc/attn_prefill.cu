// Dense prefill attention on 5th-gen tensor cores - C-ABI entry points, argument checks and the TMA tensor-map cache.
// Replaces flash-attn 2 (mma.sync) for the new-token prefill of HF Qwen2 attention (streamvln_eval.py:526; reached from
// stream_video_vln.py:340): per turn ~216 new tokens, at a window start ~1.8-2k, against [pruned memory | window KV |
// new tokens] in the KV arena.  The kernel (TS-form tcgen05.mma, Q and P in tensor memory, K/V by TMA) is in
// attn_prefill3.cu; its two superseded predecessors (SS form with 128 / 256 rows per CTA, round 1) are gone - their
// measurements are in DESIGN.md section 5.
#include <stdlib.h>

#include "tcgen05_helpers.cuh"

namespace svln {

constexpr int PF_HD = 128;   // head_dim

// ------------------------------------------------------------------------------ host side
struct MapCacheEntry {
    void* base; int n_env, n_layer, n_kv_head, max_len, box_rows; CUtensorMap map; bool valid;
};
static MapCacheEntry g_map_cache[8];

// 3-D tensor map (128 dims, max_len rows, planes) over the whole arena, box = 64 dims x box_rows keys, SWIZZLE_128B
int get_kv_map(const svln_kv_arena* a, int box_rows, CUtensorMap* out) {
    for (auto& e : g_map_cache)
        if (e.valid && e.base == a->base && e.n_env == a->n_env && e.n_layer == a->n_layer &&
            e.n_kv_head == a->n_kv_head && e.max_len == a->max_len && e.box_rows == box_rows) {
            *out = e.map;
            return SVLN_OK;
        }
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SVLN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        SVLN_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, SVLN_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    }
    const cuuint64_t planes = (cuuint64_t)a->n_env * a->n_layer * 2 * a->n_kv_head;
    cuuint64_t gdim[3] = {(cuuint64_t)PF_HD, (cuuint64_t)a->max_len, planes};
    cuuint64_t gstride[2] = {(cuuint64_t)PF_HD * 2, (cuuint64_t)a->max_len * PF_HD * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estride[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->base, gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SVLN_REQUIRE(r == CUDA_SUCCESS, SVLN_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    static int next = 0;
    MapCacheEntry& e = g_map_cache[next++ % 8];
    e = {a->base, a->n_env, a->n_layer, a->n_kv_head, a->max_len, box_rows, m, true};
    *out = m;
    return SVLN_OK;
}

static int prefill_launch(const svln_kv_arena* arena, int layer, const void* q, int n_q_head, const int32_t* env_h,
                          const int32_t* q_start_h, const int32_t* n_q_h, const int32_t* kv_len_h, const int32_t* env_d,
                          const int32_t* q_start_d, const int32_t* n_q_d, const int32_t* kv_len_d, int n_seq, float scale,
                          void* out, svln_stream_t stream, const char* who, void* workspace, size_t workspace_bytes) {
    SVLN_REQUIRE(arena && arena->base && arena->head_dim == 128, SVLN_ERR_INVALID, "%s: bad arena", who);
    SVLN_REQUIRE(layer >= 0 && layer < arena->n_layer, SVLN_ERR_INVALID, "%s: layer out of range", who);
    SVLN_REQUIRE(n_q_head > 0 && n_q_head % arena->n_kv_head == 0, SVLN_ERR_INVALID,
                 "%s: n_q_head %d not a multiple of n_kv_head %d", who, n_q_head, arena->n_kv_head);
    SVLN_REQUIRE(n_seq >= 0, SVLN_ERR_INVALID, "%s: negative n_seq", who);
    if (n_seq == 0) return SVLN_OK;
    SVLN_REQUIRE(q && out && n_q_h, SVLN_ERR_INVALID, "%s: null pointer", who);
    SVLN_REQUIRE(((uintptr_t)arena->base & 15) == 0, SVLN_ERR_INVALID, "%s: arena base must be 16-byte aligned", who);
    const bool dev = env_d != nullptr;
    SVLN_REQUIRE(dev ? (q_start_d && n_q_d && kv_len_d) : (env_h && q_start_h && kv_len_h), SVLN_ERR_INVALID,
                 "%s: null descriptor array", who);
    return prefill3_launch(arena, layer, q, n_q_head, env_h, q_start_h, n_q_h, kv_len_h, env_d, q_start_d, n_q_d, kv_len_d,
                           n_seq, scale, out, stream, who, workspace, workspace_bytes);
}

}  // namespace svln

using namespace svln;

extern "C" {

int svln_attn_prefill(const svln_kv_arena* arena, int layer, const void* q, int n_q_head,
                      const int32_t* seq_env_host, const int32_t* q_start_host, const int32_t* n_q_host,
                      const int32_t* kv_len_host, int n_seq, float scale, void* workspace, size_t workspace_bytes,
                      void* out, svln_stream_t stream) {
    return prefill_launch(arena, layer, q, n_q_head, seq_env_host, q_start_host, n_q_host, kv_len_host, nullptr, nullptr,
                          nullptr, nullptr, n_seq, scale, out, stream, "svln_attn_prefill", workspace, workspace_bytes);
}

int svln_attn_prefill_dev(const svln_kv_arena* arena, int layer, const void* q, int n_q_head, const int32_t* seq_env,
                          const int32_t* q_start, const int32_t* n_q, const int32_t* kv_len, const int32_t* n_q_host,
                          int n_seq, float scale, void* workspace, size_t workspace_bytes, void* out, svln_stream_t stream) {
    SVLN_REQUIRE(seq_env, SVLN_ERR_INVALID, "svln_attn_prefill_dev: null seq_env");
    return prefill_launch(arena, layer, q, n_q_head, nullptr, nullptr, n_q_host, nullptr, seq_env, q_start, n_q, kv_len,
                          n_seq, scale, out, stream, "svln_attn_prefill_dev", workspace, workspace_bytes);
}

}  // extern "C"

/*
 * svln.h — C-ABI of libsvln (streamvln_b200), the B200-native SlowFast streaming-context hot path.
 *
 * The reference (kason-huang/StreamVLN) has no FFI: its hot path is Python
 * (streamvln/model/stream_video_vln.py, streamvln/streamvln_agent.py) over HF transformers
 * and flash-attn.  This header is the boundary a maintainer binds from that Python with
 * ctypes.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host (or the function
 *     ends in _host: those take HOST buffers and do the copies themselves);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises unless documented;
 *   - return value: 0 = SVLN_OK, negative errno-style code otherwise; never throws;
 *     svln_last_error() returns a thread-local message for the last failure;
 *   - no entry point allocates device memory: scratch comes in as `workspace`;
 *   - bf16 buffers are passed as void* (no torch / cuda types in signatures).
 */
#ifndef SVLN_H_
#define SVLN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVLN_OK            0
#define SVLN_ERR_INVALID (-22) /* EINVAL: bad argument                   */
#define SVLN_ERR_NOMEM   (-12) /* ENOMEM: workspace too small            */
#define SVLN_ERR_RANGE   (-34) /* ERANGE: size beyond a kernel limit     */
#define SVLN_ERR_CUDA     (-5) /* EIO:    CUDA launch / runtime failure  */

#define SVLN_INVALID_KEEP 0 /* tokens with invalid depth are always kept (agent feeds depth=0,
                               streamvln_agent.py:172-174)                                   */
#define SVLN_INVALID_DROP 1 /* tokens with invalid depth are dropped                         */

#define SVLN_KEY_INVALID ((int64_t)-1)

typedef void* svln_stream_t;

int svln_version(void);
const char* svln_last_error(void);
/* SM count of the current device (grid sizing is done inside the library). */
int svln_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * (1) Slow-memory voxel token pruning.
 * The released reference accepts depths/poses/intrinsics and ignores them
 * (stream_video_vln.py:102-142; streamvln_agent.py:171 "does not use voxel pooling").
 * These calls implement the stage those arguments were plumbed for, to the spec in
 * DESIGN.md §3 (input conventions: streamvln_eval.py:126-185,254-283).
 * ---------------------------------------------------------------------------------------- */

/* Back-project the centre pixel of every pooled-token cell and quantise to a voxel key.
 *   depth [n_frames,H,W] f32 metres; pose [n_frames,16] f32 row-major cam->episodic;
 *   intr [n_frames,16] f32 row-major (fx=[0], cx=[2], fy=[5], cy=[6]);
 *   key_out [n_frames,G,G] i64: 3x21-bit packed voxel (+2^20 offset) or SVLN_KEY_INVALID. */
int svln_voxel_keys(const float* depth, const float* pose, const float* intr,
                    int64_t n_frames, int H, int W, int G, float voxel_size, float max_depth,
                    int64_t* key_out, svln_stream_t stream);

/* Scratch needed by svln_keep_latest / svln_compact_mask for at most n_tokens tokens split in
 * at most n_segments dedup segments of at most seg_tokens tokens each. */
size_t svln_prune_workspace_bytes(int64_t n_tokens, int64_t n_segments, int64_t seg_tokens);

/* Keep-latest-per-voxel dedup.  Tokens [seg_off[g], seg_off[g+1]) form dedup segment g
 * (one stream x one temporal period).  Within a segment, for every key the token with the
 * largest t survives; ties (same frame) -> lowest index.  Invalid keys follow
 * `invalid_policy`.  mask_out[n] = 1 keeps token n.  seg_off is a device array [n_segments+1];
 * max_seg_tokens bounds seg_off[g+1]-seg_off[g] (sizes the hash table; host-known). */
int svln_keep_latest(const int64_t* key, const int32_t* t, int64_t n_tokens,
                     const int64_t* seg_off, int64_t n_segments, int64_t max_seg_tokens,
                     int invalid_policy, void* workspace, size_t workspace_bytes,
                     uint8_t* mask_out, svln_stream_t stream);

/* Frame threshold: frames (runs of tokens_per_frame tokens) keeping fewer than min_keep tokens
 * are dropped whole (mask zeroed).  In place. */
int svln_frame_threshold(uint8_t* mask, int64_t n_frames, int tokens_per_frame, int min_keep,
                         svln_stream_t stream);

/* Stable compaction of a mask: idx_out[0..n_keep) = ascending indices n with mask[n]!=0.
 * n_keep_out: device i32[1].  Optional range_off (device i64 [n_ranges+1], ascending token
 * offsets, e.g. stream boundaries) -> range_keep_off_out (device i32 [n_ranges+1]) = number of
 * kept tokens before each offset.  Pass n_ranges=0 / NULLs to skip. */
int svln_compact_mask(const uint8_t* mask, int64_t n_tokens, void* workspace,
                      size_t workspace_bytes, int32_t* idx_out, int32_t* n_keep_out,
                      const int64_t* range_off, int64_t n_ranges, int32_t* range_keep_off_out,
                      svln_stream_t stream);

/* Gather surviving rows: dst[r,:] = src[idx[r],:] for r < *n_rows_dev (device i32, so the
 * data-dependent count never visits the host).  row_bytes must be a multiple of 16;
 * max_rows bounds *n_rows_dev (sizes the grid). */
int svln_gather_rows(const void* src, int64_t row_bytes, const int32_t* idx,
                     const int32_t* n_rows_dev, int64_t max_rows, void* dst,
                     svln_stream_t stream);

/* Whole pruning pass over S regular streams of T frames with G*G tokens each
 * (keys -> dedup per (stream, period) -> frame threshold -> compaction -> gather).
 *   feat [S,T,G*G,D] bf16; period<=0 => one global period; min_keep=0 => no threshold.
 *   Outputs: key_out [S,T,G,G] i64, mask_out [S*T*G*G] u8, idx_out [<=N] i32 (global token
 *   index), stream_keep_off_out [S+1] i32, n_keep_out [1] i32, feat_out [<=N, D] bf16. */
int svln_voxel_prune(const float* depth, const float* pose, const float* intr, const void* feat,
                     int n_streams, int n_frames, int H, int W, int G, int D, float voxel_size,
                     float max_depth, int period, int min_keep, int invalid_policy,
                     void* workspace, size_t workspace_bytes, int64_t* key_out,
                     uint8_t* mask_out, int32_t* idx_out, int32_t* stream_keep_off_out,
                     int32_t* n_keep_out, void* feat_out, svln_stream_t stream);

/* Same pass with HOST buffers (what a ctypes caller holding numpy arrays uses): copies
 * depth/pose/intr/feat to caller-provided device staging, runs svln_voxel_prune, copies
 * mask/idx/stream_keep_off/n_keep and the gathered rows back, synchronises the stream.
 * dev_staging must hold svln_voxel_prune_host_staging_bytes(...) bytes. */
size_t svln_voxel_prune_host_staging_bytes(int n_streams, int n_frames, int H, int W, int G,
                                           int D);
int svln_voxel_prune_host(const float* depth_host, const float* pose_host,
                          const float* intr_host, const void* feat_host, int n_streams,
                          int n_frames, int H, int W, int G, int D, float voxel_size,
                          float max_depth, int period, int min_keep, int invalid_policy,
                          void* dev_staging, size_t dev_staging_bytes, uint8_t* mask_out_host,
                          int32_t* idx_out_host, int32_t* stream_keep_off_out_host,
                          int32_t* n_keep_out_host, void* feat_out_host, svln_stream_t stream);

/* RGB frame preprocessing, replacing the host-side SigLipImageProcessor.preprocess
 * (llava/model/multimodal_encoder/siglip_encoder.py:34-66, called at streamvln_agent.py:179): Pillow's 8-bit
 * BICUBIC resample (horizontal then vertical pass, 22-bit fixed-point taps, uint8 rounding per pass) + rescale
 * 1/255 + normalise (0.5, 0.5), bit-exact.  rgb [n,H,W,3] u8 -> out [n,3,out_h,out_w] f32.  The tap tables
 * (xmin/xcnt [out_w], xkk [out_w,xksize]; ymin/ycnt [out_h], ykk [out_h,yksize]; i32, device) are Pillow's
 * precompute_coeffs + normalize_coeffs_8bpc, computed once on the host (streamvln_b200/preprocess.py). */
size_t svln_preprocess_workspace_bytes(int n, int H, int out_w);
int svln_preprocess_rgb(const uint8_t* rgb, int n, int H, int W, int out_h, int out_w, const int32_t* xmin,
                        const int32_t* xcnt, const int32_t* xkk, int xksize, const int32_t* ymin,
                        const int32_t* ycnt, const int32_t* ykk, int yksize, void* workspace,
                        size_t workspace_bytes, float* out, svln_stream_t stream);

/* Depth frame preprocessing, replacing the host path of the Habitat evaluator (streamvln_eval.py:254-256 depth in [0,1]
 * -> metres -> millimetres, :276 uint16 truncation, :126-135 Pillow NEAREST squash to the processor grid and / 1000;
 * depth-camera-filtering's filter_depth(blur_type=None) is not reproduced).  depth01 [n,H,W] f32 -> out [n,out_h,out_w]
 * f32 metres, bit-exact with that host arithmetic (tests/golden/ref_geometry.npz). */
int svln_preprocess_depth(const float* depth01, int n, int H, int W, int out_h, int out_w, float min_depth,
                          float max_depth, float* out, svln_stream_t stream);

/* Device-side embed splice, replacing the Python loop of torch.where/.tolist()/torch.cat in
 * prepare_inputs_labels_for_multimodal (stream_video_vln.py:182-238).  row_src[r] >= 0 selects
 * row row_src[r] of `embed_table` (token embedding); row_src[r] < 0 selects row (-row_src[r]-1)
 * of `feat` (image / memory feature rows). */
int svln_splice_rows(const void* embed_table, const void* feat, const int32_t* row_src,
                     int64_t n_rows, int64_t row_bytes, void* dst, svln_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2) Fast-streaming KV arena.  Replaces HF DynamicCache.update (= torch.cat per layer per
 * forward, reached via stream_video_vln.py:340) and the reset-everything window policy
 * (stream_video_vln.py:473-479, streamvln_agent.py:192-198).
 * Layout (bf16): [env][layer][2 (K,V)][kv_head][max_len][head_dim]; head_dim must be 128.
 * The memory is owned by the caller (one allocation made at start-up).
 * ---------------------------------------------------------------------------------------- */
typedef struct svln_kv_arena {
    void* base;
    int32_t n_env;
    int32_t n_layer;
    int32_t n_kv_head;
    int32_t head_dim;
    int32_t max_len;
    int32_t reserved;
} svln_kv_arena;

size_t svln_kv_arena_bytes(int n_env, int n_layer, int n_kv_head, int head_dim, int max_len);

/* Fused RoPE + KV append for one layer (HF Qwen2 apply_rotary_pos_emb + DynamicCache.update).
 *   qkv [n_tok, (n_q_head + 2*n_kv_head)*head_dim] bf16 (q heads, then k heads, then v heads);
 *   tok_env/tok_pos [n_tok] i32: destination stream and cache row (= RoPE position);
 *   cos/sin [n_pos, head_dim] bf16, HF layout (halves duplicated);
 *   q_out [n_tok, n_q_head, head_dim] bf16 rotated queries. */
int svln_kv_append_rope(const svln_kv_arena* arena, int layer, const void* qkv, int n_q_head,
                        const int32_t* tok_env, const int32_t* tok_pos, int n_tok,
                        const void* cos, const void* sin, int n_pos, void* q_out,
                        svln_stream_t stream);

/* Window eviction + in-place compaction (+ fused RoPE re-rotation of K) over ALL layers.
 * For each job j (one stream): cache rows d in [first_row[j], new_len[j]) receive old row
 * src_row[j*max_len + d] (src >= d, strictly increasing in d); rows below first_row are
 * untouched.  With rerotate!=0 every moved K row is rotated by (d - src) positions using the
 * fp32 tables cos_delta/sin_delta [n_delta, head_dim/2] indexed by (src - d) (rotation by
 * -(src-d)); V rows are moved unchanged.  Dropped rows simply disappear: this is both the
 * eviction of an old turn and the splice of its surviving tokens onto the memory region. */
int svln_kv_compact(const svln_kv_arena* arena, const int32_t* job_env,
                    const int32_t* first_row, const int32_t* new_len, const int32_t* src_row,
                    int n_jobs, int rerotate, const float* cos_delta, const float* sin_delta,
                    int n_delta, svln_stream_t stream);

/* RoPE position remap without moving rows: K row r of stream `env` (all layers) is rotated by
 * delta_pos[r - row0] positions, r in [row0, row0+n_rows).  |delta| < n_delta. */
int svln_kv_rerotate(const svln_kv_arena* arena, int env, int row0, int n_rows,
                     const int32_t* delta_pos, const float* cos_delta, const float* sin_delta,
                     int n_delta, svln_stream_t stream);

/* Batched, device-side planning of the slow-memory eviction: every stream that slides its window this turn is one
 * job; nothing visits the host between the dedup and the compaction.  No reference counterpart (the released code
 * drops the whole cache per window, stream_video_vln.py:473-479): this is the glue between (1) and (2), to the spec in
 * DESIGN.md section 4 (oracle/slowfast.py::_evict).  Per job j, jobs[j*SVLN_SF_JOB_WORDS + w] (device i32) holds
 *   w=0 stream (arena env)      w=1 prefix rows P           w=2 memory tokens M (before)
 *   w=3 first row of the evicted turn (== P + M)            w=4 offset of its image rows inside the turn
 *   w=5 rows of the evicted turn w=6 cache length L          w=7 frame time of the evicted turn
 *   w=8 ring slot of the turn's voxel keys                   w=9 offset of the job's segment in cand_key / cand_t
 * State: mem_key [n_env, mem_cap] i64, mem_t [n_env, mem_cap] i32 (slow-memory voxel keys / frame times, oldest
 * first), turn_key [n_env, n_slots, tokens_per_frame] i64 (keys of the live turns' image tokens).
 *   svln_sf_gather_candidates: cand[off .. off+M+tpf) = [memory | evicted turn] of every job (the dedup segments that
 *     svln_keep_latest takes);
 *   svln_sf_evict_plan (after svln_keep_latest -> mask): survivors in stable order, at most mem_cap (oldest dropped),
 *     become the new memory; src_row [n_jobs, max_len] = the source map of svln_kv_compact
 *     ([0,P) identity | survivors | every row behind the evicted turn); plan_out [4, n_jobs] i32 = first moved row,
 *     new cache length, new memory size, survivors before the cap - rows 0 and 1 are svln_kv_compact's
 *     first_row / new_len arguments as they stand. */
#define SVLN_SF_JOB_WORDS 10
int svln_sf_gather_candidates(const int32_t* jobs, int n_jobs, const int64_t* mem_key,
                              const int32_t* mem_t, int mem_cap, const int64_t* turn_key, int n_slots,
                              int tokens_per_frame, int64_t* cand_key, int32_t* cand_t,
                              svln_stream_t stream);
int svln_sf_evict_plan(const int32_t* jobs, int n_jobs, const uint8_t* mask, const int64_t* cand_key,
                       const int32_t* cand_t, int mem_cap, int tokens_per_frame, int max_len,
                       int64_t* mem_key, int32_t* mem_t, int32_t* src_row, int32_t* plan_out,
                       svln_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (3) Per-step attention over [pruned memory | window KV | new tokens]; replaces flash-attn 2
 * (selected at streamvln_eval.py:526) for HF Qwen2 attention (reached via
 * stream_video_vln.py:340).  GQA; bf16 in/out, fp32 softmax; causal.
 * A "sequence" b is one stream's chunk of new tokens: q rows [q_start[b], q_start[b]+n_q[b])
 * hold positions kv_len[b]-n_q[b] ... kv_len[b]-1 of stream seq_env[b], whose cache already
 * contains the new K/V (svln_kv_append_rope runs first).
 * ---------------------------------------------------------------------------------------- */

size_t svln_attn_workspace_bytes(int n_q_rows_max, int n_q_head, int head_dim, int max_len);

/* Split-KV decode / short-chunk attention (HBM-streaming; seq_env/q_start/n_q/kv_len are DEVICE
 * arrays so a CUDA graph can replay it; max_n_q / max_kv_len are host bounds used for the grid). */
int svln_attn_decode(const svln_kv_arena* arena, int layer, const void* q, int n_q_head,
                     const int32_t* seq_env, const int32_t* q_start, const int32_t* n_q,
                     const int32_t* kv_len, int n_seq, int max_n_q, int max_kv_len, float scale,
                     void* workspace, size_t workspace_bytes, void* out, svln_stream_t stream);

/* Dense prefill attention on tcgen05 tensor cores with TMEM accumulators (any n_q[b] >= 1).  `workspace`
 * (svln_attn_workspace_bytes; may be NULL) lets small grids split the key range across CTAs (split-KV) and
 * merge the partials, so a single-episode turn (48 CTAs) still fills the 148 SMs. */
int svln_attn_prefill(const svln_kv_arena* arena, int layer, const void* q, int n_q_head,
                      const int32_t* seq_env_host, const int32_t* q_start_host,
                      const int32_t* n_q_host, const int32_t* kv_len_host, int n_seq,
                      float scale, void* workspace, size_t workspace_bytes, void* out,
                      svln_stream_t stream);

/* Same kernel with the sequence descriptors in DEVICE memory (seq_env / q_start / n_q / kv_len, i32 [n_seq]),
 * so that a captured CUDA graph can be replayed with new cache lengths; only n_q (host copy, sizes the grid)
 * is baked into the launch and must equal the device n_q. */
int svln_attn_prefill_dev(const svln_kv_arena* arena, int layer, const void* q, int n_q_head,
                          const int32_t* seq_env, const int32_t* q_start, const int32_t* n_q,
                          const int32_t* kv_len, const int32_t* n_q_host, int n_seq, float scale,
                          void* workspace, size_t workspace_bytes, void* out, svln_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Small fused ops around the Qwen2 linears (which stay cuBLAS): HF Qwen2RMSNorm and the
 * SwiGLU product, with HF's bf16 rounding points.
 * ---------------------------------------------------------------------------------------- */
int svln_rmsnorm(const void* x, const void* weight, int64_t n_rows, int dim, float eps,
                 void* out, svln_stream_t stream);
/* x <- x + r (bf16), out <- rmsnorm(x) * weight : residual add fused with the next norm */
int svln_add_rmsnorm(void* x, const void* r, const void* weight, int64_t n_rows, int dim,
                     float eps, void* out, svln_stream_t stream);
/* gate_up [n_rows, 2*inter] bf16 (gate then up) -> out [n_rows, inter] = silu(gate)*up */
int svln_swiglu(const void* gate_up, int64_t n_rows, int inter, void* out,
                svln_stream_t stream);
/* Weight-streaming GEMV for the decode of 1..4 streams: out[m,n] = sum_k x[m,k] * W[n,k] (+ bias[n]); W [N,K] and
 * x [M,K] bf16 row-major, fp32 accumulation, out [M,N] bf16 (or fp32 when out_is_f32, for the lm_head logits).
 * The Qwen2 linears at M = 1 are pure weight streaming; larger M are GEMMs and stay cuBLAS. */
int svln_gemv(const void* W, const void* x, const void* bias, void* out, int M, int N, int K,
              int out_is_f32, svln_stream_t stream);
/* argmax over fp32 view of bf16/f32 logits rows; ties -> lowest index (torch.argmax). */
int svln_argmax(const void* logits, int is_f32, int64_t n_rows, int64_t vocab,
                int64_t* out_idx, svln_stream_t stream);

/* Bilinear token pooling of the projected vision features: feat [n_frames, grid_in^2, dim] bf16 ->
 * out [n_frames, grid_out^2, dim] bf16.  Replaces StreamVLNForCausalLM.get_2dPool with mm_spatial_pool_mode
 * "bilinear" (streamvln/model/stream_video_vln.py:53-73: permute to NCHW, F.interpolate(size=grid_out, bilinear,
 * align_corners=False), permute back; grid_out = ceil(grid_in / 2) = 14 for the 27 x 27 SigLIP grid) with ATen's
 * arithmetic in fp32. */
int svln_pool_bilinear(const void* feat, int n_frames, int grid_in, int grid_out, int dim, void* out,
                       svln_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * Fused decode stack: one persistent launch = one greedy decode step of ONE or TWO streams (n_stream = 1, 2;
 * two streams share one pass over the weights and need the packed weight stream)
 * (embedding row -> n_layer x [RMSNorm, QKV + bias + RoPE + append to the arena, split-KV attention, O,
 *  RMSNorm, SwiGLU MLP] -> final norm -> lm_head -> argmax).  Replaces the per-token forward of HF's greedy
 * loop (transformers 4.45.1 GenerationMixin._sample -> Qwen2ForCausalLM.forward; reached from
 * streamvln/model/stream_video_vln.py:340 and :402) for the action-token decode, which at batch 1 is pure
 * weight streaming.  Same rounding points as the launch-per-op path (svln_add_rmsnorm, svln_gemv,
 * svln_kv_append_rope, svln_attn_decode, svln_swiglu, svln_argmax).
 *   kv_len[b] INCLUDES the token fed in this step (its position is kv_len[b] - 1; the caller bumps the
 *   length before the call, as for svln_attn_decode).  All pointers are device pointers.
 *   layer_ptrs: int64 [n_layer][8] = {w_qkv [ (Hq+2Hkv)*128, hidden ], b_qkv (or 0), w_o [hidden, Hq*128],
 *   w_gate_up [2*inter, hidden] (gate rows then up rows), w_down [hidden, inter], ln1, ln2, 0} (bf16).
 * ---------------------------------------------------------------------------------------- */
typedef struct svln_decode_stack {
    int32_t n_layer, hidden, inter, n_q_head, n_kv_head, head_dim, vocab;
    int32_t n_stream;            /* 1 or 2 (2 only with `packed`) */
    float rms_eps, attn_scale;
    const void* layer_ptrs;
    const void* embed;           /* [vocab, hidden] bf16 */
    const void* final_norm;      /* [hidden] bf16 */
    const void* lm_head;         /* [vocab, hidden] bf16 */
    const void* rope_cos;        /* [>= max position + 1, 128] bf16 (HF cos/sin tables) */
    const void* rope_sin;
    const int64_t* tok_in;       /* [n_stream] token fed in this step */
    int64_t* tok_out;            /* [n_stream] greedy next token */
    float* logits;               /* [n_stream, vocab] fp32 or NULL */
    const int32_t* seq_env;      /* [n_stream] arena stream of each row */
    const int32_t* kv_len;       /* [n_stream] */
    void* workspace;             /* >= svln_decode_stack_workspace_bytes(), 256-byte aligned */
    size_t workspace_bytes;
    const void* packed;          /* packed weight stream written by svln_decode_stack_pack, or NULL: the step then reads
                                  * the row-major weights behind layer_ptrs / lm_head (slower; SVLN_STACK_FEED=0 forces it) */
    size_t packed_bytes;
} svln_decode_stack;
size_t svln_decode_stack_workspace_bytes(const svln_decode_stack* desc);
/* The decode step streams every weight byte exactly once per token.  svln_decode_stack_pack writes a second copy of the
 * layer linears and lm_head (same bytes, `svln_decode_stack_packed_bytes` of them: 15.3 GB for Qwen2-7B) laid out as one
 * contiguous stream per CTA in consumption order, which a single producer thread per SM feeds to shared memory with
 * bulk copies (cp.async.bulk + mbarriers) across phase boundaries.  The layout depends on the SM count of the current
 * device; pack on the device that runs the step (the library remembers what each packed buffer of the process was packed
 * for; a step on a buffer it does not know, or packed for other dims / another grid, returns SVLN_ERR_INVALID).
 * `packed` 128-byte aligned.  Reads layer_ptrs (synchronises `stream`). */
size_t svln_decode_stack_packed_bytes(const svln_decode_stack* desc);
int svln_decode_stack_pack(const svln_decode_stack* desc, void* packed, size_t packed_bytes, svln_stream_t stream);
/* SVLN_ERR_RANGE when n_stream > 2, or 2 without `packed` (the caller falls back to the launch-per-op path; larger
 * batches are GEMMs). */
int svln_decode_stack_step(const svln_kv_arena* arena, const svln_decode_stack* desc, svln_stream_t stream);
/* A grid barrier that times out (a bug, or a grid that could not be made co-resident) does not hang the GPU: the
 * kernel finishes and reports token -1 for every stream.  This call (synchronous) returns 0 when the last launch on
 * desc->workspace completed normally, else the 1-based index of the barrier that gave up (cta / count: which CTA,
 * and the arrival count it saw); -1 on a bad descriptor. */
int svln_decode_stack_status(const svln_decode_stack* desc, int* cta, int* count);
/* Debug aid: with SVLN_STACK_TRACE=1 in the environment CTA 0 stamps clock64 at [layer][phase 0..5][point 0..3]
 * (phase start, GEMV start, GEMV end, before the grid barrier); copies up to max_entries stamps of the last
 * launch to `host` (after a device synchronise by the caller) and returns how many. */
int svln_decode_stack_trace(long long* host, int max_entries);

#ifdef __cplusplus
}
#endif
#endif /* SVLN_H_ */

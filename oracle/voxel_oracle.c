/*
 * oracle/voxel_oracle.c — CPU oracle for the slow-memory voxel token pruning.  TEST INFRASTRUCTURE
 * ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg; the product path (streamvln_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the released reference contains no voxel pruning to restate
 * (/root/reference/streamvln/streamvln_agent.py:171 "Currently the model does not use voxel pooling";
 * /root/reference/streamvln/model/stream_video_vln.py:102-142 accepts depths/poses/intrinsics and
 * never reads them), and it holds no golden vector for it.  This file is therefore the normative
 * statement of the spec in DESIGN.md §3 / SURVEY.md §8c-V; input conventions follow
 * /root/reference/streamvln/streamvln_eval.py:126-185 (depth in metres on the 384x384 grid, 4x4
 * intrinsics rescaled to 384x384, 4x4 camera->episodic pose).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile); fp32 with a
 * fixed operation order and no FMA contraction so the GPU kernel can match bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KEY_INVALID ((int64_t)-1)

static int32_t voxel_coord(float x, float voxel_size) {
    float q = floorf(x / voxel_size);
    q = fminf(fmaxf(q, -1048576.0f), 1048575.0f);
    return (int32_t)q;
}

/* depth [n_frames,H,W], pose/intr [n_frames,16] row-major, key_out [n_frames,G,G] */
void oracle_voxel_keys(const float* depth, const float* pose, const float* intr, int64_t n_frames, int H,
                       int W, int G, float voxel_size, float max_depth, int64_t* key_out) {
    const float sw = (float)W / (float)G, sh = (float)H / (float)G;
    for (int64_t f = 0; f < n_frames; ++f) {
        const float* P = pose + f * 16;
        const float* K = intr + f * 16;
        const float fx = K[0], cx = K[2], fy = K[5], cy = K[6];
        for (int i = 0; i < G; ++i)
            for (int j = 0; j < G; ++j) {
                const float u = ((float)j + 0.5f) * sw;
                const float v = ((float)i + 0.5f) * sh;
                int px = (int)floorf(u), py = (int)floorf(v);
                px = px < 0 ? 0 : (px > W - 1 ? W - 1 : px);
                py = py < 0 ? 0 : (py > H - 1 ? H - 1 : py);
                const float d = depth[(f * H + py) * W + px];
                int64_t key = KEY_INVALID;
                if (d > 0.0f && d < max_depth && isfinite(d)) {
                    const float xc = d * ((u - cx) / fx);
                    const float yc = d * ((v - cy) / fy);
                    const float zc = d;
                    const float xw = ((P[0] * xc + P[1] * yc) + P[2] * zc) + P[3];
                    const float yw = ((P[4] * xc + P[5] * yc) + P[6] * zc) + P[7];
                    const float zw = ((P[8] * xc + P[9] * yc) + P[10] * zc) + P[11];
                    const int64_t vx = (int64_t)voxel_coord(xw, voxel_size) + 1048576;
                    const int64_t vy = (int64_t)voxel_coord(yw, voxel_size) + 1048576;
                    const int64_t vz = (int64_t)voxel_coord(zw, voxel_size) + 1048576;
                    key = (vx << 42) | (vy << 21) | vz;
                }
                key_out[(f * G + i) * G + j] = key;
            }
    }
}

typedef struct { int64_t key; int32_t t; int32_t local; } Ent;

static int ent_cmp(const void* a, const void* b) {
    const Ent* x = (const Ent*)a; const Ent* y = (const Ent*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->t != y->t) return x->t > y->t ? -1 : 1;      /* latest frame first */
    return x->local < y->local ? -1 : (x->local > y->local);  /* then lowest index */
}

/* keep-latest per key inside each segment [seg_off[g], seg_off[g+1]); invalid_policy 0 keep / 1 drop.
 * Deliberately a different algorithm from the GPU's hash table (sort + first-of-run). */
void oracle_keep_latest(const int64_t* key, const int32_t* t, int64_t n_tokens, const int64_t* seg_off,
                        int64_t n_segments, int invalid_policy, uint8_t* mask_out) {
    (void)n_tokens;
    for (int64_t g = 0; g < n_segments; ++g) {
        const int64_t a = seg_off[g], b = seg_off[g + 1];
        const int64_t n = b - a;
        if (n <= 0) continue;
        Ent* e = (Ent*)malloc(sizeof(Ent) * (size_t)n);
        int64_t m = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (key[a + i] < 0) {
                mask_out[a + i] = invalid_policy == 0 ? 1 : 0;
            } else {
                mask_out[a + i] = 0;
                e[m].key = key[a + i]; e[m].t = t[a + i]; e[m].local = (int32_t)i; ++m;
            }
        }
        qsort(e, (size_t)m, sizeof(Ent), ent_cmp);
        for (int64_t i = 0; i < m; ++i)
            if (i == 0 || e[i].key != e[i - 1].key) mask_out[a + e[i].local] = 1;
        free(e);
    }
}

void oracle_frame_threshold(uint8_t* mask, int64_t n_frames, int tokens_per_frame, int min_keep) {
    for (int64_t f = 0; f < n_frames; ++f) {
        int c = 0;
        for (int i = 0; i < tokens_per_frame; ++i) c += mask[f * tokens_per_frame + i] != 0;
        if (c < min_keep) memset(mask + f * tokens_per_frame, 0, (size_t)tokens_per_frame);
    }
}

/* stable compaction + gather; returns n_keep.  feat/feat_out may be NULL. range_off [n_ranges+1] */
int64_t oracle_compact_gather(const uint8_t* mask, int64_t n_tokens, const uint16_t* feat, int64_t D,
                              int32_t* idx_out, uint16_t* feat_out, const int64_t* range_off,
                              int64_t n_ranges, int32_t* range_keep_off_out) {
    int64_t k = 0, r = 0;
    for (int64_t n = 0; n <= n_tokens; ++n) {
        while (range_off && r <= n_ranges && range_off[r] == n) range_keep_off_out[r++] = (int32_t)k;
        if (n == n_tokens) break;
        if (mask[n]) {
            idx_out[k] = (int32_t)n;
            if (feat && feat_out) memcpy(feat_out + k * D, feat + n * D, (size_t)D * 2);
            ++k;
        }
    }
    return k;
}

/* whole pass over S regular streams (mirrors svln_voxel_prune) */
int64_t oracle_voxel_prune(const float* depth, const float* pose, const float* intr, const uint16_t* feat,
                           int n_streams, int n_frames, int H, int W, int G, int D, float voxel_size,
                           float max_depth, int period, int min_keep, int invalid_policy, int64_t* key_out,
                           uint8_t* mask_out, int32_t* idx_out, int32_t* stream_keep_off_out,
                           uint16_t* feat_out) {
    const int64_t tpf = (int64_t)G * G;
    const int64_t n_tokens = (int64_t)n_streams * n_frames * tpf;
    const int per = (period <= 0 || period > n_frames) ? n_frames : period;
    const int pps = n_frames ? (n_frames + per - 1) / per : 0;
    const int64_t n_segments = (int64_t)n_streams * pps;
    oracle_voxel_keys(depth, pose, intr, (int64_t)n_streams * n_frames, H, W, G, voxel_size, max_depth, key_out);
    int32_t* t = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_tokens + 1));
    for (int64_t n = 0; n < n_tokens; ++n) t[n] = (int32_t)((n / tpf) % n_frames);
    int64_t* seg_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_segments + 1));
    for (int64_t g = 0; g <= n_segments; ++g) {
        const int64_t s = pps ? g / pps : 0, p = pps ? g % pps : 0;
        int64_t o = s * n_frames * tpf + p * per * tpf;
        const int64_t lim = (s + 1) * n_frames * tpf;
        seg_off[g] = o > lim ? lim : o;
    }
    oracle_keep_latest(key_out, t, n_tokens, seg_off, n_segments, invalid_policy, mask_out);
    if (min_keep > 0) oracle_frame_threshold(mask_out, (int64_t)n_streams * n_frames, (int)tpf, min_keep);
    int64_t* range_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_streams + 1));
    for (int s = 0; s <= n_streams; ++s) range_off[s] = (int64_t)s * n_frames * tpf;
    const int64_t k = oracle_compact_gather(mask_out, n_tokens, feat, D, idx_out, feat_out, range_off,
                                            n_streams, stream_keep_off_out);
    free(t); free(seg_off); free(range_off);
    return k;
}

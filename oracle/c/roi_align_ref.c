/*
 * CPU oracle for PyramidROIAlign -- plain C restatement (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product path never does.
 *
 * PARITY UNPINNED: the reference's arithmetic for this stage lives in TensorFlow 1.x
 * (tf.image.crop_and_resize, tf.log, tf.round -- not vendored, not installable here) and the
 * reference has no tests or golden vectors.  This file restates
 *   /root/reference/evaluate_models/modified_dense_model.py:313-315  (log2_graph)
 *   /root/reference/evaluate_models/modified_dense_model.py:351-363  (level assignment)
 *   /root/reference/evaluate_models/modified_dense_model.py:366-393  (per-level crop_and_resize)
 *   /root/reference/evaluate_models/modified_dense_model.py:395-416  (concat + re-sort)
 * together with TF's published CropAndResize CPU semantics (one bilinear sample per bin,
 * end-points inclusive, floor/ceil taps, whole sample = extrapolation value when out of range).
 *
 * Build with -ffp-contract=off: every fp32 operation is individually rounded, matching the
 * numpy oracle and the CUDA kernel (which uses __fmul_rn/__fadd_rn/__fsub_rn).
 *
 * Work is sharded over boxes with OpenMP, mirroring how TF's CPU kernel shards boxes over its
 * intra-op thread pool.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* correctly rounded fp32 log: fp64 log rounded once (see oracle/roi_align.py header) */
static float log_f32(float x) { return (float)log((double)x); }

/* x86 cvttss2si semantics: NaN / inf / out of range -> INT_MIN */
static int32_t f32_to_i32_x86(float r) {
    if (!(r >= -2147483648.0f && r < 2147483648.0f)) return INT32_MIN;
    return (int32_t)r;
}

/* modified_dense_model.py:351-363 */
int32_t oracle_fpn_level_one(const float *box, int img_h, int img_w) {
    float h = box[2] - box[0];
    float w = box[3] - box[1];
    float image_area = (float)((double)img_h * (double)img_w);
    float denom = 224.0f / sqrtf(image_area);
    float s = sqrtf(h * w);
    float q = s / denom;
    float r = log_f32(q) / log_f32(2.0f);
    float rr = rintf(r);                      /* tf.round: half to even (default FE mode) */
    int64_t lv = 4 + (int64_t)f32_to_i32_x86(rr);
    if (lv < 2) lv = 2;
    if (lv > 5) lv = 5;
    return (int32_t)lv;
}

void oracle_fpn_levels_f32(const float *boxes, int64_t n, int img_h, int img_w, int32_t *levels) {
    for (int64_t i = 0; i < n; ++i) levels[i] = oracle_fpn_level_one(boxes + 4 * i, img_h, img_w);
}

/* One box of tf.image.crop_and_resize(bilinear, extrapolation 0) into out[ph*pw*C]. */
static void crop_one(const float *fm, int H, int W, int C, const float *box, int ph, int pw,
                     float *out) {
    const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
    const float hs = (ph > 1) ? (y2 - y1) * (float)(H - 1) / (float)(ph - 1) : 0.0f;
    const float ws = (pw > 1) ? (x2 - x1) * (float)(W - 1) / (float)(pw - 1) : 0.0f;
    for (int y = 0; y < ph; ++y) {
        const float in_y = (ph > 1) ? y1 * (float)(H - 1) + (float)y * hs
                                    : 0.5f * (y1 + y2) * (float)(H - 1);
        float *orow = out + (size_t)y * pw * C;
        if (!(in_y >= 0.0f && in_y <= (float)(H - 1))) {   /* NaN treated as out of range */
            memset(orow, 0, sizeof(float) * (size_t)pw * C);
            continue;
        }
        const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        const float ly = in_y - floorf(in_y);
        for (int x = 0; x < pw; ++x) {
            const float in_x = (pw > 1) ? x1 * (float)(W - 1) + (float)x * ws
                                        : 0.5f * (x1 + x2) * (float)(W - 1);
            float *o = orow + (size_t)x * C;
            if (!(in_x >= 0.0f && in_x <= (float)(W - 1))) {
                memset(o, 0, sizeof(float) * (size_t)C);
                continue;
            }
            const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
            const float lx = in_x - floorf(in_x);
            const float *tl = fm + ((size_t)top * W + left) * C;
            const float *tr = fm + ((size_t)top * W + right) * C;
            const float *bl = fm + ((size_t)bot * W + left) * C;
            const float *br = fm + ((size_t)bot * W + right) * C;
            for (int c = 0; c < C; ++c) {
                const float t = tl[c] + (tr[c] - tl[c]) * lx;
                const float b = bl[c] + (br[c] - bl[c]) * lx;
                o[c] = t + (b - t) * ly;
            }
        }
    }
}

/*
 * Literal form: level assignment, one crop_and_resize call per level into a level-grouped
 * buffer, then the re-sort gather back to (image, box) order.  `scratch` must hold
 * B*N*ph*pw*C floats (the concat of the four per-level results).  out is [B*N, ph, pw, C].
 * Returns 0, or -1 when N > 100000 (the reference's sort key `batch*100000 + box` collides).
 */
int oracle_pyramid_roi_align_literal_f32(const float *boxes, const float *const fmaps[4],
                                         const int fm_h[4], const int fm_w[4], int B, int N, int C,
                                         int ph, int pw, int img_h, int img_w, float *scratch,
                                         float *out, int32_t *levels) {
    if (N > 100000) return -1;
    const int64_t R = (int64_t)B * N;
    const size_t row = (size_t)ph * pw * C;
    int32_t *lv = levels ? levels : (int32_t *)malloc(sizeof(int32_t) * (size_t)(R ? R : 1));
    oracle_fpn_levels_f32(boxes, R, img_h, img_w, lv);
    /* box_to_level: position in the concat for every (image, box), grouped by level */
    int64_t *src_of = (int64_t *)malloc(sizeof(int64_t) * (size_t)(R ? R : 1));
    int64_t *members = (int64_t *)malloc(sizeof(int64_t) * (size_t)(R ? R : 1));
    int64_t pos = 0;
    for (int l = 2; l <= 5; ++l) {
        const int64_t start = pos;
        for (int64_t i = 0; i < R; ++i)
            if (lv[i] == l) { members[pos] = i; src_of[i] = pos; ++pos; }
        const int li = l - 2;
        const int H = fm_h[li], W = fm_w[li];
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t k = start; k < pos; ++k) {
            const int64_t i = members[k];
            const int b = (int)(i / N);
            crop_one(fmaps[li] + (size_t)b * H * W * C, H, W, C, boxes + 4 * i, ph, pw,
                     scratch + (size_t)k * row);
        }
    }
    /* sort by batch*100000+box == original order; gather */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < R; ++i)
        memcpy(out + (size_t)i * row, scratch + (size_t)src_of[i] * row, sizeof(float) * row);
    free(src_of);
    free(members);
    if (!levels) free(lv);
    return 0;
}

/* Direct form (single pass, results written in final order). */
int oracle_pyramid_roi_align_f32(const float *boxes, const float *const fmaps[4], const int fm_h[4],
                                 const int fm_w[4], int B, int N, int C, int ph, int pw, int img_h,
                                 int img_w, float *out, int32_t *levels) {
    const int64_t R = (int64_t)B * N;
    const size_t row = (size_t)ph * pw * C;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < R; ++i) {
        const int32_t l = oracle_fpn_level_one(boxes + 4 * i, img_h, img_w);
        if (levels) levels[i] = l;
        const int li = l - 2;
        const int H = fm_h[li], W = fm_w[li];
        const int b = (int)(i / N);
        crop_one(fmaps[li] + (size_t)b * H * W * C, H, W, C, boxes + 4 * i, ph, pw,
                 out + (size_t)i * row);
    }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

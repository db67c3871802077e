// Caption post-processing that follows the decoder in the dense-captioning evaluation path
// (SURVEY.md section 8f rank 1): per image, greedy non-maximum suppression of the captioned RoIs by
// caption score and the top-K cut.  Replaces refine_generations
// (/root/reference/evaluate_models/test_score_dense_captions.py:245-283) + non_max_suppression
// (/root/reference/evaluate_models/utils.py:69-104) with that copy's overlap measure
// (utils.py:30-48: the Dice coefficient 2*I/(A+B), not IoU).
//
// One CTA per image: bitonic sort of (score, index) in shared memory (descending score; equal scores:
// larger index first = a stable ascending argsort reversed), then the sequential greedy scan with the
// suppression of each picked box done by the whole CTA.  fp32 arithmetic op for op as numpy's.
#include "common.cuh"

namespace dcap {

constexpr int kNmsThreads = 256;

__global__ void __launch_bounds__(kNmsThreads) refine_generations_kernel(const float *__restrict__ boxes,
                                                                         const float *__restrict__ scores, int n_boxes,
                                                                         int n_pad, float thr, int max_keep,
                                                                         int32_t *__restrict__ keep, int32_t *__restrict__ n_keep) {
    extern __shared__ unsigned char sm_raw[];
    float *s_score = reinterpret_cast<float *>(sm_raw);           // [n_pad] sorted scores
    int *s_idx = reinterpret_cast<int *>(s_score + n_pad);        // [n_pad] box index (-1 = padding)
    float4 *s_box = reinterpret_cast<float4 *>(s_idx + n_pad);    // [n_pad] boxes in sorted order
    float *s_area = reinterpret_cast<float *>(s_box + n_pad);     // [n_pad]
    unsigned char *s_dead = reinterpret_cast<unsigned char *>(s_area + n_pad);   // [n_pad]
    const int img = blockIdx.x, tid = threadIdx.x;
    const float *sc = scores + (long long)img * n_boxes;
    const float4 *bx = reinterpret_cast<const float4 *>(boxes) + (long long)img * n_boxes;
    for (int i = tid; i < n_pad; i += kNmsThreads) {
        s_score[i] = i < n_boxes ? sc[i] : -INFINITY;
        s_idx[i] = i < n_boxes ? i : -1;
    }
    __syncthreads();
    // bitonic sort, order: score descending, then index descending; padding (-1) last.  NaN scores sort last too.
    auto before = [](float sa, int ia, float sb, int ib) {
        if (ia < 0 || ib < 0) return ib < 0 && ia >= 0;
        const bool na = sa != sa, nb = sb != sb;
        if (na || nb) return nb && !na;
        return sa > sb || (sa == sb && ia > ib);
    };
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pad; i += kNmsThreads) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const float sa = s_score[i], sb = s_score[l];
                    const int ia = s_idx[i], ib = s_idx[l];
                    const bool swap = up ? before(sb, ib, sa, ia) : before(sa, ia, sb, ib);
                    if (swap) { s_score[i] = sb; s_score[l] = sa; s_idx[i] = ib; s_idx[l] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n_pad; i += kNmsThreads) {
        const int b = s_idx[i];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b >= 0) v = bx[b];
        s_box[i] = v;                                               // (y1, x1, y2, x2)
        s_area[i] = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
        s_dead[i] = b < 0;
    }
    __syncthreads();
    // greedy scan over ALL boxes (the reference suppresses first and cuts to the best max_keep afterwards);
    // the survivors' sorted positions are collected in s_kept (re-using the dead flags' neighbour: s_area is
    // no longer needed once a box has been visited, but a separate list keeps this simple)
    int *s_kept = reinterpret_cast<int *>(s_dead + n_pad);          // [n_pad]
    int cnt = 0;                                                    // replicated in every thread (uniform)
    for (int i = 0; i < n_boxes; ++i) {
        if (s_dead[i]) continue;                                    // CTA-uniform (shared memory, after a barrier)
        const float4 p = s_box[i];
        const float pa = s_area[i];
        for (int j = i + 1 + tid; j < n_boxes; j += kNmsThreads) {
            if (s_dead[j]) continue;
            const float4 q = s_box[j];
            const float y1 = fmaxf(p.x, q.x), y2 = fminf(p.z, q.z), x1 = fmaxf(p.y, q.y), x2 = fminf(p.w, q.w);
            const float inter = __fmul_rn(fmaxf(__fsub_rn(x2, x1), 0.f), fmaxf(__fsub_rn(y2, y1), 0.f));
            const float ov = __fdiv_rn(__fmul_rn(2.f, inter), __fadd_rn(pa, s_area[j]));
            if (ov > thr) s_dead[j] = 1;                             // NaN (0/0) compares false: kept, as in numpy
        }
        if (tid == 0) s_kept[cnt] = i;
        ++cnt;
        __syncthreads();
    }
    __syncthreads();
    // top max_keep by score: np.argsort(scores[keep])[::-1] on the (descending) survivor list reverses every run
    // of equal scores, so survivor q of a run [a, b) lands at a + (b - 1 - q)
    for (int q = tid; q < cnt; q += kNmsThreads) {
        const float v = s_score[s_kept[q]];
        int a = q, b = q + 1;
        while (a > 0 && s_score[s_kept[a - 1]] == v) --a;
        while (b < cnt && s_score[s_kept[b]] == v) ++b;
        const int dst = a + (b - 1 - q);
        if (dst < max_keep) keep[(long long)img * max_keep + dst] = s_idx[s_kept[q]];
    }
    const int kept = cnt < max_keep ? cnt : max_keep;
    for (int i = kept + tid; i < max_keep; i += kNmsThreads) keep[(long long)img * max_keep + i] = -1;
    if (tid == 0) n_keep[img] = kept;
}

}  // namespace dcap

using namespace dcap;

extern "C" int dc_refine_generations(const float *boxes, const float *scores, int n_images, int n_boxes,
                                     float nms_threshold, int max_keep, int32_t *keep, int32_t *n_keep, void *stream) {
    DC_REQUIRE(n_images >= 0 && n_boxes >= 0 && max_keep >= 1, "bad n_images / n_boxes / max_keep");
    if (n_images == 0) return DC_OK;
    DC_REQUIRE(keep && n_keep && (n_boxes == 0 || (boxes && scores)), "null pointer argument");
    DC_REQUIRE(((uintptr_t)boxes & 15) == 0, "boxes must be 16-byte aligned");
    int n_pad = 4;
    while (n_pad < n_boxes) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * (4 + 4 + 16 + 4 + 1 + 4) + 16;
    DC_REQUIRE(smem <= 200 * 1024, "n_boxes=%d per image exceeds the shared-memory NMS capacity (%d)", n_boxes, 200 * 1024 / 33);
    static std::atomic<unsigned long long> attr_set{0};
    DC_CHECK_CUDA(once_per_device(attr_set, [] {
        return cudaFuncSetAttribute(refine_generations_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }));
    refine_generations_kernel<<<n_images, kNmsThreads, smem, (cudaStream_t)stream>>>(boxes, scores, n_boxes, n_pad, nms_threshold,
                                                                                    max_keep, keep, n_keep);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

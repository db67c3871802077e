// End-to-end per-RoI captioning: PyramidROIAlign -> RoI head -> greedy inject-LSTM decoding in one
// call.  This is the path the reference walks per image in its evaluation loops
// (generate_features -> model.predict: evaluate_models/eval_text_generation_model.py:138-148,
// evaluate_models/generate_one_roi_features.py:69-76), with the RoI features never leaving HBM:
// the bf16 decoder consumes the ROIAlign output directly in bf16 (half the feature traffic).
#include "decoder.cuh"

using namespace dcap;

extern "C" int dc_caption_rois(DcDecoder *dec, const float *boxes, const float *const fmaps[4],
                               const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                               int img_h, int img_w, int32_t *tokens, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    Decoder &D = dec->impl;
    const long long R = (long long)n_images * n_boxes;
    if (int rc = D.check_ready((int)R)) return rc;
    DC_REQUIRE(D.cfg.arch == DC_ARCH_V1, "dc_caption_rois needs a v1 decoder");
    DC_REQUIRE(R < (1ll << 31), "too many RoIs in one call");
    if (R == 0) return DC_OK;
    DC_REQUIRE(tokens, "null pointer argument");
    if (int rc = D.reserve((int)R)) return rc;
    void *feat = nullptr;
    if (int rc = D.roi_feature_buffer((int)R, &feat)) return rc;
    const int p = D.cfg.pool, C = D.cfg.channels;
    if (D.cfg.dtype == DC_DTYPE_BF16) {
        if (int rc = dc_pyramid_roi_align_bf16out(boxes, fmaps, fm_h, fm_w, n_images, n_boxes, C, p, p, img_h,
                                                  img_w, (uint16_t *)feat, nullptr, stream)) return rc;
        return D.greedy(feat, DC_FEATS_ROI_BF16, (int)R, tokens, nullptr, (cudaStream_t)stream);
    }
    if (int rc = dc_pyramid_roi_align_f32(boxes, fmaps, fm_h, fm_w, n_images, n_boxes, C, p, p, img_h, img_w,
                                          (float *)feat, nullptr, stream)) return rc;
    return D.greedy(feat, DC_FEATS_ROI_F32, (int)R, tokens, nullptr, (cudaStream_t)stream);
}

// Host-buffer form: all pointers are HOST memory (pinned for asynchronous copies).  Images are
// uploaded one at a time on a copy stream while the previous image is aligned and decoded on the
// compute stream (PCIe H2D of image i+1 overlaps ROIAlign + decode of image i); only the token
// ids travel back.
extern "C" int dc_caption_rois_host(DcDecoder *dec, const float *boxes, const float *const fmaps[4],
                                    const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                    int img_h, int img_w, int32_t *tokens) {
    DC_REQUIRE(dec, "null decoder");
    Decoder &D = dec->impl;
    const long long R = (long long)n_images * n_boxes;
    if (int rc = D.check_ready((int)R)) return rc;
    DC_REQUIRE(D.cfg.arch == DC_ARCH_V1, "dc_caption_rois_host needs a v1 decoder");
    if (R == 0) return DC_OK;
    DC_REQUIRE(boxes && fmaps && fm_h && fm_w && tokens, "null pointer argument");
    const int P = D.cfg.padding, C = D.cfg.channels;
    if (int rc = D.reserve(n_boxes)) return rc;

    cudaStream_t sa = nullptr, sb = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    float *d_fm[2][4] = {{nullptr}};
    float *d_boxes = nullptr;
    int32_t *d_tok = nullptr;
    size_t fm_bytes[4];
    for (int l = 0; l < 4; ++l) fm_bytes[l] = sizeof(float) * (size_t)fm_h[l] * fm_w[l] * C;
    int rc = DC_OK;
#define PIPE_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess && rc == DC_OK)                                               \
            rc = set_error(DC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)
    PIPE_TRY(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
    PIPE_TRY(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
    for (int i = 0; i < 2 && rc == DC_OK; ++i) {
        PIPE_TRY(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming));
        PIPE_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        for (int l = 0; l < 4; ++l) PIPE_TRY(cudaMallocAsync(&d_fm[i][l], fm_bytes[l], sa));
    }
    if (rc == DC_OK) {
        PIPE_TRY(cudaMallocAsync(&d_boxes, sizeof(float) * 4 * (size_t)R, sa));
        PIPE_TRY(cudaMallocAsync(&d_tok, sizeof(int32_t) * (size_t)R * P, sa));
        PIPE_TRY(cudaMemcpyAsync(d_boxes, boxes, sizeof(float) * 4 * (size_t)R, cudaMemcpyHostToDevice, sa));
    }
    for (int img = 0; img < n_images && rc == DC_OK; ++img) {
        const int s = img & 1;
        if (img >= 2) PIPE_TRY(cudaStreamWaitEvent(sa, done[s], 0));
        for (int l = 0; l < 4; ++l)
            PIPE_TRY(cudaMemcpyAsync(d_fm[s][l], fmaps[l] + (size_t)img * (fm_bytes[l] / 4), fm_bytes[l],
                                     cudaMemcpyHostToDevice, sa));
        PIPE_TRY(cudaEventRecord(up[s], sa));
        PIPE_TRY(cudaStreamWaitEvent(sb, up[s], 0));
        if (rc != DC_OK) break;
        const float *maps[4] = {d_fm[s][0], d_fm[s][1], d_fm[s][2], d_fm[s][3]};
        int krc = dc_caption_rois(dec, d_boxes + (size_t)img * n_boxes * 4, maps, fm_h, fm_w, 1, n_boxes, img_h,
                                  img_w, d_tok + (size_t)img * n_boxes * P, sb);
        if (krc != DC_OK) { rc = krc; break; }
        PIPE_TRY(cudaEventRecord(done[s], sb));
    }
    if (rc == DC_OK)
        PIPE_TRY(cudaMemcpyAsync(tokens, d_tok, sizeof(int32_t) * (size_t)R * P, cudaMemcpyDeviceToHost, sb));
    if (sa) cudaStreamSynchronize(sa);
    if (sb) {
        cudaError_t e = cudaStreamSynchronize(sb);
        if (e != cudaSuccess && rc == DC_OK)
            rc = set_error(DC_ERR_CUDA, "caption pipeline failed: %s", cudaGetErrorString(e));
    }
    for (int i = 0; i < 2; ++i) {
        for (int l = 0; l < 4; ++l) if (d_fm[i][l]) cudaFreeAsync(d_fm[i][l], sb ? sb : 0);
        if (up[i]) cudaEventDestroy(up[i]);
        if (done[i]) cudaEventDestroy(done[i]);
    }
    if (d_boxes) cudaFreeAsync(d_boxes, sb ? sb : 0);
    if (d_tok) cudaFreeAsync(d_tok, sb ? sb : 0);
    if (sb) cudaStreamSynchronize(sb);
    if (sa) cudaStreamDestroy(sa);
    if (sb) cudaStreamDestroy(sb);
#undef PIPE_TRY
    return rc;
}

// End-to-end per-RoI captioning: PyramidROIAlign -> RoI head -> greedy inject-LSTM decoding in one
// call.  This is the path the reference walks per image in its evaluation loops
// (generate_features -> model.predict: evaluate_models/eval_text_generation_model.py:138-148,
// evaluate_models/generate_one_roi_features.py:69-76), with the RoI features never leaving HBM:
// the bf16 decoder consumes the ROIAlign output directly in bf16 (half the feature traffic).
#include "decoder.cuh"

#include <string.h>

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

// Host-buffer form: all pointers are HOST memory (pinned for asynchronous copies).  Images are uploaded one at a
// time on a copy stream while the previous image is aligned and decoded on the compute stream (PCIe H2D of image
// i+1 overlaps ROIAlign + decode of image i); only the token ids travel back.
//
// The pipeline state (two streams, events, two pyramid slots, per-call box / token staging) lives in the decoder
// handle and is reused by every call.  submit() only enqueues; wait() blocks until the OLDEST outstanding call's
// tokens are on the host (they land in pinned staging memory and wait() copies them into the caller's array).  With two calls outstanding the upload of call k+1 runs under the decode tail of call k
// (the last image of a call cannot be decoded before its pyramid has landed, so a single blocking call always
// exposes one image's decode: 14.6 ms against the 12.9 ms copy floor at 8 x 1000 RoIs).
// These entry points take no caller stream: they order against their own streams only.  Work the caller has
// enqueued on the handle from another stream (dc_adam_step ...) must be complete before calling them.
namespace dcap {

struct HostPipe {
    cudaStream_t copy = nullptr, compute = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};   // pyramid slot uploaded / consumed
    cudaEvent_t finished[2] = {nullptr, nullptr};                           // call slot: tokens are on the host
    float *fm[2][4] = {{nullptr}};
    size_t fm_cap[4] = {0, 0, 0, 0};
    float *boxes[2] = {nullptr, nullptr};
    int32_t *tok[2] = {nullptr, nullptr};
    // pinned staging of a call's token ids: a device -> host copy straight into the caller's (usually pageable) array
    // would make cudaMemcpyAsync block until the whole call has run, and the "two calls in flight" would never overlap
    float *box_host[2] = {nullptr, nullptr};              // the (small) box array is staged too: callers pass pageable numpy arrays
    int32_t *tok_host[2] = {nullptr, nullptr};
    int32_t *tok_user[2] = {nullptr, nullptr};
    size_t tok_bytes[2] = {0, 0};
    size_t box_cap[2] = {0, 0}, tok_cap[2] = {0, 0};
    long long images = 0;        // pyramid-slot uses so far (slot = images & 1)
    long long submitted = 0, waited = 0;
    bool used[2] = {false, false};
    ~HostPipe() {
        if (compute) cudaStreamSynchronize(compute);
        if (copy) cudaStreamSynchronize(copy);
        for (int i = 0; i < 2; ++i) {
            for (int l = 0; l < 4; ++l) if (fm[i][l]) cudaFree(fm[i][l]);
            if (boxes[i]) cudaFree(boxes[i]);
            if (tok[i]) cudaFree(tok[i]);
            if (tok_host[i]) cudaFreeHost(tok_host[i]);
            if (box_host[i]) cudaFreeHost(box_host[i]);
            for (cudaEvent_t e : {up[i], done[i], finished[i]}) if (e) cudaEventDestroy(e);
        }
        if (copy) cudaStreamDestroy(copy);
        if (compute) cudaStreamDestroy(compute);
    }
};

void free_host_pipe(HostPipe *p) { delete p; }

static int pipe_grow(void **ptr, size_t *cap, size_t bytes) {
    if (*cap >= bytes) return DC_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    DC_CHECK_CUDA(cudaMalloc(ptr, bytes));
    *cap = bytes;
    return DC_OK;
}

}  // namespace dcap

extern "C" int dc_caption_rois_host_wait(DcDecoder *dec) {
    DC_REQUIRE(dec, "null decoder");
    HostPipe *hp = dec->impl.host_pipe;
    if (!hp || hp->waited == hp->submitted) return DC_OK;
    const int slot = (int)(hp->waited & 1);
    hp->waited++;
    const cudaError_t e = cudaEventSynchronize(hp->finished[slot]);
    if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "caption pipeline failed: %s", cudaGetErrorString(e));
    if (hp->tok_user[slot] && hp->tok_bytes[slot]) memcpy(hp->tok_user[slot], hp->tok_host[slot], hp->tok_bytes[slot]);
    hp->tok_user[slot] = nullptr;
    return DC_OK;
}

extern "C" int dc_caption_rois_host_submit(DcDecoder *dec, const float *boxes, const float *const fmaps[4],
                                           const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                           int img_h, int img_w, int32_t *tokens) {
    DC_REQUIRE(dec, "null decoder");
    Decoder &D = dec->impl;
    const long long R = (long long)n_images * n_boxes;
    DC_REQUIRE(n_images >= 0 && n_boxes >= 0 && R < (1ll << 31), "too many RoIs in one call");
    if (int rc = D.check_ready((int)R)) return rc;
    DC_REQUIRE(D.cfg.arch == DC_ARCH_V1, "dc_caption_rois_host needs a v1 decoder");
    if (R == 0) return DC_OK;
    DC_REQUIRE(boxes && fmaps && fm_h && fm_w && tokens, "null pointer argument");
    const int P = D.cfg.padding, C = D.cfg.channels;
    if (!D.host_pipe) D.host_pipe = new HostPipe();
    HostPipe &hp = *D.host_pipe;
    if (hp.submitted - hp.waited >= 2)                      // two call slots: retire the oldest first
        if (int rc = dc_caption_rois_host_wait(dec)) return rc;
    if (!hp.copy) {
        DC_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.copy, cudaStreamNonBlocking));
        DC_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.compute, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            DC_CHECK_CUDA(cudaEventCreateWithFlags(&hp.up[i], cudaEventDisableTiming));
            DC_CHECK_CUDA(cudaEventCreateWithFlags(&hp.done[i], cudaEventDisableTiming));
            DC_CHECK_CUDA(cudaEventCreateWithFlags(&hp.finished[i], cudaEventDisableTiming));
        }
    }
    size_t fm_bytes[4];
    bool regrow = false;
    for (int l = 0; l < 4; ++l) {
        DC_REQUIRE(fm_h[l] >= 1 && fm_w[l] >= 1 && fmaps[l], "feature map %d is empty / null", l);
        fm_bytes[l] = sizeof(float) * (size_t)fm_h[l] * fm_w[l] * C;
        regrow = regrow || fm_bytes[l] > hp.fm_cap[l];
    }
    const int cs = (int)(hp.submitted & 1);
    const bool grow_call = hp.box_cap[cs] < sizeof(float) * 4 * (size_t)R || hp.tok_cap[cs] < sizeof(int32_t) * (size_t)R * P;
    if (regrow || grow_call || D.cap < n_boxes) {
        // (re)allocation and the decoder workspace are not stream ordered: drain the pipeline first (first call only,
        // or when a later call brings larger shapes)
        while (hp.waited < hp.submitted)
            if (int rc = dc_caption_rois_host_wait(dec)) return rc;
        DC_CHECK_CUDA(cudaStreamSynchronize(hp.compute));
        DC_CHECK_CUDA(cudaStreamSynchronize(hp.copy));
        for (int l = 0; l < 4; ++l)
            if (fm_bytes[l] > hp.fm_cap[l]) {
                size_t cap = 0;
                for (int i = 0; i < 2; ++i) {
                    cap = hp.fm_cap[l];
                    if (int rc = pipe_grow((void **)&hp.fm[i][l], &cap, fm_bytes[l])) return rc;
                }
                hp.fm_cap[l] = cap;
            }
        if (hp.box_cap[cs] < sizeof(float) * 4 * (size_t)R) {
            if (hp.box_host[cs]) cudaFreeHost(hp.box_host[cs]);
            hp.box_host[cs] = nullptr;
            DC_CHECK_CUDA(cudaHostAlloc((void **)&hp.box_host[cs], sizeof(float) * 4 * (size_t)R, cudaHostAllocDefault));
        }
        if (int rc = pipe_grow((void **)&hp.boxes[cs], &hp.box_cap[cs], sizeof(float) * 4 * (size_t)R)) return rc;
        if (hp.tok_cap[cs] < sizeof(int32_t) * (size_t)R * P) {
            if (hp.tok_host[cs]) cudaFreeHost(hp.tok_host[cs]);
            hp.tok_host[cs] = nullptr;
            DC_CHECK_CUDA(cudaHostAlloc((void **)&hp.tok_host[cs], sizeof(int32_t) * (size_t)R * P, cudaHostAllocDefault));
        }
        if (int rc = pipe_grow((void **)&hp.tok[cs], &hp.tok_cap[cs], sizeof(int32_t) * (size_t)R * P)) return rc;
        if (int rc = D.reserve(n_boxes)) return rc;
        void *unused = nullptr;
        if (int rc = D.roi_feature_buffer(n_boxes, &unused)) return rc;
    }
    memcpy(hp.box_host[cs], boxes, sizeof(float) * 4 * (size_t)R);       // slot cs is free: its previous call has been waited for
    DC_CHECK_CUDA(cudaMemcpyAsync(hp.boxes[cs], hp.box_host[cs], sizeof(float) * 4 * (size_t)R, cudaMemcpyHostToDevice, hp.copy));
    for (int img = 0; img < n_images; ++img) {
        const int s = (int)(hp.images & 1);
        if (hp.used[s]) DC_CHECK_CUDA(cudaStreamWaitEvent(hp.copy, hp.done[s], 0));   // the slot's previous image is decoded
        for (int l = 0; l < 4; ++l)
            DC_CHECK_CUDA(cudaMemcpyAsync(hp.fm[s][l], fmaps[l] + (size_t)img * (fm_bytes[l] / 4), fm_bytes[l],
                                          cudaMemcpyHostToDevice, hp.copy));
        DC_CHECK_CUDA(cudaEventRecord(hp.up[s], hp.copy));
        DC_CHECK_CUDA(cudaStreamWaitEvent(hp.compute, hp.up[s], 0));
        const float *maps[4] = {hp.fm[s][0], hp.fm[s][1], hp.fm[s][2], hp.fm[s][3]};
        if (int rc = dc_caption_rois(dec, hp.boxes[cs] + (size_t)img * n_boxes * 4, maps, fm_h, fm_w, 1, n_boxes, img_h,
                                     img_w, hp.tok[cs] + (size_t)img * n_boxes * P, hp.compute)) return rc;
        DC_CHECK_CUDA(cudaEventRecord(hp.done[s], hp.compute));
        hp.used[s] = true;
        hp.images++;
    }
    hp.tok_user[cs] = tokens;
    hp.tok_bytes[cs] = sizeof(int32_t) * (size_t)R * P;
    DC_CHECK_CUDA(cudaMemcpyAsync(hp.tok_host[cs], hp.tok[cs], hp.tok_bytes[cs], cudaMemcpyDeviceToHost, hp.compute));
    DC_CHECK_CUDA(cudaEventRecord(hp.finished[cs], hp.compute));
    hp.submitted++;
    return DC_OK;
}

// blocking form: submit + wait for everything outstanding
extern "C" int dc_caption_rois_host(DcDecoder *dec, const float *boxes, const float *const fmaps[4],
                                    const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                    int img_h, int img_w, int32_t *tokens) {
    if (int rc = dc_caption_rois_host_submit(dec, boxes, fmaps, fm_h, fm_w, n_images, n_boxes, img_h, img_w, tokens)) return rc;
    DC_REQUIRE(dec, "null decoder");
    HostPipe *hp = dec->impl.host_pipe;
    while (hp && hp->waited < hp->submitted)
        if (int rc = dc_caption_rois_host_wait(dec)) return rc;
    return DC_OK;
}

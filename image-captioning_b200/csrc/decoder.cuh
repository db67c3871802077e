// Decoder handle (see decoder.cu).
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "decoder_kernels.cuh"
#include "gemm.cuh"

namespace dcap {

constexpr int kDense = 1024;      // Dense(1024, relu) of the word model (text_generation_model.py:143)

struct Weight {
    std::string name;
    std::vector<int64_t> shape;
    int64_t numel = 0;
    float *dev = nullptr;         // fp32, Keras layout; points into Decoder::arena
    int64_t offset = 0;           // floats from the start of the arena
    bool trainable = true;        // frozen: embeddings (trainable=False) and BatchNorm moving statistics
    bool is_set = false;
};

struct Workspace {
    float *F = nullptr, *a1 = nullptr, *g1f = nullptr, *d1f = nullptr;
    float *xh1 = nullptr, *xh1b = nullptr, *xh2 = nullptr, *xh2b = nullptr;
    float *c1 = nullptr, *c1b = nullptr, *c2 = nullptr, *c2b = nullptr;
    float *gates = nullptr, *d = nullptr, *h2 = nullptr, *logits = nullptr, *cand_p = nullptr;
    int32_t *tok = nullptr, *newtok = nullptr, *parent = nullptr, *cand_idx = nullptr;
    int32_t *hist_a = nullptr, *hist_b = nullptr;
    double *score_a = nullptr, *score_b = nullptr;
};

struct Bf16State;                 // decoder_bf16.cu
struct HostPipe;                  // pipeline.cu: streams / events / staging of dc_caption_rois_host_*
void free_host_pipe(HostPipe *p);
int accumulate_log(float *scores, const float *maxprob, int rows, bool first, cudaStream_t s);   // decoder.cu

struct Decoder {
    DcDecoderConfig cfg{};
    int device = 0;
    std::vector<Weight> weights;
    std::vector<void *> owned, ws_owned;
    bool finalized = false;
    bool emb_dirty = true;        // the bf16 embedding table must be rebuilt at the next refresh
    int cap = 0, rep_cap = 0;
    Workspace ws;
    float *bn_scale[2] = {nullptr, nullptr}, *bn_shift[2] = {nullptr, nullptr};
    float *w1cat = nullptr, *w2cat = nullptr;
    float *rep_g1f = nullptr, *rep_d1f = nullptr;
    Bf16State *bf = nullptr;
    HostPipe *host_pipe = nullptr;
    void *roi_buf = nullptr;
    // CUDA graphs of the bf16 greedy loop, keyed by (feats, kind, B, tokens); captured on the second
    // call with the same key and replayed afterwards (removes ~100 launches of host overhead per call)
    typedef std::tuple<const void *, int, int, void *> GraphKey;
    std::map<GraphKey, cudaGraphExec_t> graphs;
    std::map<GraphKey, int> graph_calls;
    cudaStream_t graph_stream = nullptr;
    cudaEvent_t graph_ev_in = nullptr, graph_ev_out = nullptr;
    bool use_graphs = true;
    size_t roi_buf_bytes = 0;
    // One flat fp32 arena for all weights: trainable tensors first (declaration order), frozen ones
    // after, so that gradients / optimiser state / the NCCL all-reduce are single contiguous ranges.
    float *arena = nullptr;
    int64_t n_train = 0, n_total = 0;
    float *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr, *adam_vhat = nullptr;   // [n_train], lazily allocated

    ~Decoder();
    void declare(const std::string &name, std::vector<int64_t> shape);
    void declare_all();
    int layout_arena();
    Weight *find(const std::string &name);
    const float *W(const char *name);
    int dev_alloc(void **p, size_t bytes, std::vector<void *> &list);
    int finalize(cudaStream_t s);
    int refresh_derived(cudaStream_t s);
    int reserve(int rows);
    int ensure_rep(int R);
    int roi_feature_buffer(int R, void **out);
    int check_ready(int B);

    int linear_f32(const float *x, int ldx, int M, const float *Wk, int K, int N, const float *bias,
                   const float *addend, int ld_addend, const float *scale, const float *shift, bool relu,
                   float *out, int ldo, cudaStream_t s);
    int head(const void *feats, int kind, int B, float *out, cudaStream_t s);
    int v1_hoist(int B, cudaStream_t s);
    int v1_reset_state(int R, cudaStream_t s);
    int v1_step(int R, const float *g1f, const float *d1f, cudaStream_t s);
    int greedy(const void *feats, int kind, int B, int32_t *tokens, float *probs, cudaStream_t s, float *scores = nullptr);
    int beam(const void *feats, int kind, int B, int k, int32_t *tokens, double *scores, cudaStream_t s);
    int v2_reset(int B, cudaStream_t s);
    int v2_word_step(int B, cudaStream_t s);
    int v2_output(int B, cudaStream_t s);
    int v2_head_into_xin(const void *feats, int kind, int B, cudaStream_t s);
    int v2_predict(const void *feats, int kind, const int32_t *words, int B, int L, float *probs, cudaStream_t s);
    int v2_greedy(const void *feats, int kind, int B, int32_t *tokens, float *probs, cudaStream_t s, const int32_t *start = nullptr,
                  float *scores = nullptr);

    // bf16 / tcgen05 path (decoder_bf16.cu)
    int refresh_bf16(bool fresh, cudaStream_t s);
    int reserve_bf16(size_t R);
    int head_bf16(const void *feats, int kind, int B, float *out, cudaStream_t s);
    int v1_hoist_bf16(int B, cudaStream_t s, bool blocked32 = false);   // blocked32: the layout greedy_loop.cu reads
    int v1_hoist_merged_bf16(int B, bool fresh_fb, cudaStream_t s);
    static bool hoist_merged();
    int reset_state_bf16(int R, cudaStream_t s);
    int v1_step_bf16(int R, const float *g1f, const float *d1f, cudaStream_t s);
    int greedy_bf16(const void *feats, int kind, int B, int32_t *tokens, cudaStream_t s, float *scores = nullptr);
    bool greedy_loop_folds() const;                    // ... with the head-feature terms folded into its contractions (no hoisted fp32 terms)
    bool greedy_loop_ok(int B) const;                      // greedy_loop.cu: the whole loop as one persistent kernel
    int greedy_loop_bf16(int B, int32_t *tokens, float *scores, cudaStream_t s);
    int greedy_bf16_graphed(const void *feats, int kind, int B, int32_t *tokens, cudaStream_t s);
    void drop_graphs();
    int v2_begin_bf16(const void *feats, int kind, int B, cudaStream_t s);
    int v2_word_step_bf16(int B, bool gather, cudaStream_t s);
    int v2_image_step_bf16(int B, cudaStream_t s);
    int v2_predict_bf16(const void *feats, int kind, const int32_t *words, int B, int L, float *probs, cudaStream_t s);
    int v2_greedy_bf16(const void *feats, int kind, int B, int32_t *tokens, float *probs, cudaStream_t s, const int32_t *start,
                       float *scores);
    int beam_bf16(const void *feats, int kind, int B, int k, int32_t *tokens, double *scores, cudaStream_t s);
    void free_bf16();

    // training step (train.cu; bf16 v1 decoder only)
    int train_forward(const void *feats, int kind, int B, const int32_t *gt, const int32_t *targets, cudaStream_t s);
    int teacher_forced_probs(const void *feats, int kind, int B, const int32_t *gt, float *probs, cudaStream_t s);
    const DcTrainOptions *train_opts = nullptr;        // valid during train_step only
    bool train_logits_bf16 = false;                    // train_step: the vocabulary GEMM stores bf16 logits into the dlogits buffer
    int train_step(const void *feats, int kind, int B, const int32_t *gt, const int32_t *targets, float inv_count,
                   float *loss, cudaStream_t s);
    int train_step_v2(const void *feats, int kind, int B, const int32_t *words, int L, const int32_t *targets, float inv_count,
                      float *loss, cudaStream_t s);
    int adam_step(float lr, float beta1, float beta2, float eps, int amsgrad, long long t, float grad_scale,
                  cudaStream_t s);
    int adam_step_range(float lr, float beta1, float beta2, float eps, int amsgrad, long long t, float grad_scale,
                        long long offset, long long numel, bool refresh, cudaStream_t s);
    int params_updated(cudaStream_t s);
    int ensure_grads();
    int refresh_train_weights(cudaStream_t s);
    void invalidate_train_copy();
    int grad_bucket(int i, int64_t *offset, int64_t *numel);
    int wait_grad_bucket(int i, cudaStream_t waiter);
    void free_train();
};

}  // namespace dcap

struct DcDecoder { dcap::Decoder impl; };

// Decoder handle: weights in Keras layout, derived device buffers, and the orchestration of the
// RoI head + inject-LSTM decoder (greedy / beam / v2) on one CUDA stream.  C ABI in include/dcap.h.
//
// Reference semantics (paths relative to /root/reference):
//   head            dense_img_cap_separate_models/text_generation_model.py:249-262
//   word model      text_generation_model.py:130-156   (incremental form)
//   greedy          text_generation_model.py:192-232
//   v2 inject       text_generation_model_v2.py:140-166, evaluate_models/test_score_dense_captions.py:216-225
//   beam            image captioning/test.py:23-64
#include "decoder.cuh"

#include <string.h>
#include <algorithm>

namespace dcap {

static const float kBnEps = 1e-3f;       // keras BatchNormalization default epsilon

// ------------------------------------------------------------------------------------------
// weight table
// ------------------------------------------------------------------------------------------
void Decoder::declare(const std::string &name, std::vector<int64_t> shape) {
    Weight w;
    w.name = name;
    w.shape = shape;
    w.numel = 1;
    for (auto d : shape) w.numel *= d;
    weights.push_back(w);
}

Weight *Decoder::find(const std::string &name) {
    for (auto &w : weights)
        if (w.name == name) return &w;
    return nullptr;
}

const float *Decoder::W(const char *name) {
    Weight *w = find(name);
    return w ? w->dev : nullptr;
}

void Decoder::declare_all() {
    const int64_t F = cfg.feat, E = cfg.embed, V = cfg.vocab, U = cfg.units, p = cfg.pool, C = cfg.channels;
    declare("mrcnn_class_conv1/kernel", {p, p, C, F});
    declare("mrcnn_class_conv1/bias", {F});
    for (const char *bn : {"mrcnn_class_bn1", "mrcnn_class_bn2"}) {
        if (std::string(bn) == "mrcnn_class_bn2") {
            declare("mrcnn_class_conv2/kernel", {1, 1, F, F});
            declare("mrcnn_class_conv2/bias", {F});
        }
        declare(std::string(bn) + "/gamma", {F});
        declare(std::string(bn) + "/beta", {F});
        declare(std::string(bn) + "/moving_mean", {F});
        declare(std::string(bn) + "/moving_variance", {F});
    }
    declare("imgcap_embedding_layer/embeddings", {V, E});
    if (cfg.arch == DC_ARCH_V1) {
        declare("imgcap_lstm1/kernel", {E + F, 4 * U});
        declare("imgcap_lstm1/recurrent_kernel", {U, 4 * U});
        declare("imgcap_lstm1/bias", {4 * U});
        declare("imgcap_lstm2/kernel", {U, 4 * U});
        declare("imgcap_lstm2/recurrent_kernel", {U, 4 * U});
        declare("imgcap_lstm2/bias", {4 * U});
        declare("imgcap_lstm_d1/kernel", {U + F, kDense});
        declare("imgcap_lstm_d1/bias", {kDense});
        declare("imgcap_lstm_d2/kernel", {kDense, V});
        declare("imgcap_lstm_d2/bias", {V});
    } else {
        const int64_t Wu = cfg.word_units;
        declare("lstm_1/kernel", {E, 4 * Wu});
        declare("lstm_1/recurrent_kernel", {Wu, 4 * Wu});
        declare("lstm_1/bias", {4 * Wu});
        declare("imgcap_lstm/kernel", {F + Wu, 4 * U});
        declare("imgcap_lstm/recurrent_kernel", {U, 4 * U});
        declare("imgcap_lstm/bias", {4 * U});
        declare("imgcap_d1/kernel", {U, V});
        declare("imgcap_d1/bias", {V});
    }
}

// arena offsets: trainable tensors first, each 256-byte aligned (sizes here are multiples anyway)
int Decoder::layout_arena() {
    int64_t off = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (auto &w : weights) {
            const bool frozen = w.name.find("/moving_") != std::string::npos || w.name.find("/embeddings") != std::string::npos;
            w.trainable = !frozen;
            if ((pass == 0) != w.trainable) continue;
            w.offset = off;
            off += (w.numel + 63) / 64 * 64;
        }
        if (pass == 0) n_train = off;
    }
    n_total = off;
    DC_CHECK_CUDA(cudaMalloc((void **)&arena, sizeof(float) * (size_t)n_total));
    DC_CHECK_CUDA(cudaMemset(arena, 0, sizeof(float) * (size_t)n_total));
    for (auto &w : weights) w.dev = arena + w.offset;
    return DC_OK;
}

Decoder::~Decoder() {
    if (host_pipe) free_host_pipe(host_pipe);
    drop_graphs();
    free_train();
    if (graph_stream) cudaStreamDestroy(graph_stream);
    if (graph_ev_in) cudaEventDestroy(graph_ev_in);
    if (graph_ev_out) cudaEventDestroy(graph_ev_out);
    free_bf16();
    if (roi_buf) cudaFree(roi_buf);
    if (arena) cudaFree(arena);
    for (float *p : {grads, adam_m, adam_v, adam_vhat})
        if (p) cudaFree(p);
    for (void *p : owned) cudaFree(p);
    for (void *p : ws_owned) cudaFree(p);
}

int Decoder::dev_alloc(void **p, size_t bytes, std::vector<void *> &list) {
    DC_CHECK_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    list.push_back(*p);
    return DC_OK;
}

// ------------------------------------------------------------------------------------------
// finalize: derived buffers
// ------------------------------------------------------------------------------------------
struct BnFoldArgs {
    const float *gamma[2], *beta[2], *mean[2], *var[2];
    float *scale[2], *shift[2];
};
__global__ void bn_fold_kernel(BnFoldArgs a, float eps, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y;
    if (i >= n) return;
    // tf.nn.batch_normalization: inv = rsqrt(var + eps) * gamma; x*inv + (beta - mean*inv)
    const float inv = (1.0f / sqrtf(a.var[l][i] + eps)) * a.gamma[l][i];
    a.scale[l][i] = inv;
    a.shift[l][i] = a.beta[l][i] - a.mean[l][i] * inv;
}

int Decoder::finalize(cudaStream_t s) {
    drop_graphs();
    invalidate_train_copy();                  // weights were set from the host: the bf16 mirror is stale
    for (auto &w : weights)
        if (!w.is_set)
            return set_error(DC_ERR_STATE, "weight '%s' has not been set", w.name.c_str());
    if (int rc = refresh_derived(s)) return rc;
    DC_CHECK_CUDA(cudaStreamSynchronize(s));
    finalized = true;
    return DC_OK;
}

// (Re)builds every buffer derived from the fp32 master weights: folded BatchNorm, stacked fp32
// operands, bf16 K-major copies.  Buffers are allocated on the first call and re-used afterwards
// (their addresses are baked into captured CUDA graphs), so the optimiser can call this every step.
int Decoder::refresh_derived(cudaStream_t s) {
    const bool fresh = owned.empty();
    const int F = cfg.feat, E = cfg.embed, U = cfg.units;
    for (int i = 0; i < 2; ++i) {
        if (fresh) {
            if (int rc = dev_alloc((void **)&bn_scale[i], sizeof(float) * F, owned)) return rc;
            if (int rc = dev_alloc((void **)&bn_shift[i], sizeof(float) * F, owned)) return rc;
        }
    }
    BnFoldArgs bn;                                          // both BatchNorm layers of the head in one launch (blockIdx.y)
    for (int i = 0; i < 2; ++i) {
        const std::string name = i == 0 ? "mrcnn_class_bn1" : "mrcnn_class_bn2";
        bn.gamma[i] = W((name + "/gamma").c_str()); bn.beta[i] = W((name + "/beta").c_str());
        bn.mean[i] = W((name + "/moving_mean").c_str()); bn.var[i] = W((name + "/moving_variance").c_str());
        bn.scale[i] = bn_scale[i]; bn.shift[i] = bn_shift[i];
    }
    bn_fold_kernel<<<dim3(ceil_div(F, 256), 2), 256, 0, s>>>(bn, kBnEps, F);
    DC_CHECK_LAUNCH();
    if (cfg.dtype == DC_DTYPE_BF16) return refresh_bf16(fresh, s);
    if (cfg.arch == DC_ARCH_V1) {
        // stacked operands: [W1[:E] ; U1]  and  [W2 ; U2]
        if (fresh) {
            if (int rc = dev_alloc((void **)&w1cat, sizeof(float) * (size_t)(E + U) * 4 * U, owned)) return rc;
            if (int rc = dev_alloc((void **)&w2cat, sizeof(float) * (size_t)(2 * U) * 4 * U, owned)) return rc;
        }
        DC_CHECK_CUDA(cudaMemcpyAsync(w1cat, W("imgcap_lstm1/kernel"), sizeof(float) * (size_t)E * 4 * U,
                                      cudaMemcpyDeviceToDevice, s));
        DC_CHECK_CUDA(cudaMemcpyAsync(w1cat + (size_t)E * 4 * U, W("imgcap_lstm1/recurrent_kernel"),
                                      sizeof(float) * (size_t)U * 4 * U, cudaMemcpyDeviceToDevice, s));
        DC_CHECK_CUDA(cudaMemcpyAsync(w2cat, W("imgcap_lstm2/kernel"), sizeof(float) * (size_t)U * 4 * U,
                                      cudaMemcpyDeviceToDevice, s));
        DC_CHECK_CUDA(cudaMemcpyAsync(w2cat + (size_t)U * 4 * U, W("imgcap_lstm2/recurrent_kernel"),
                                      sizeof(float) * (size_t)U * 4 * U, cudaMemcpyDeviceToDevice, s));
    } else {
        const int Wu = cfg.word_units;
        if (fresh)
            if (int rc = dev_alloc((void **)&w1cat, sizeof(float) * (size_t)(E + Wu) * 4 * Wu, owned)) return rc;
        DC_CHECK_CUDA(cudaMemcpyAsync(w1cat, W("lstm_1/kernel"), sizeof(float) * (size_t)E * 4 * Wu,
                                      cudaMemcpyDeviceToDevice, s));
        DC_CHECK_CUDA(cudaMemcpyAsync(w1cat + (size_t)E * 4 * Wu, W("lstm_1/recurrent_kernel"),
                                      sizeof(float) * (size_t)Wu * 4 * Wu, cudaMemcpyDeviceToDevice, s));
    }
    return DC_OK;
}

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
int Decoder::reserve(int rows) {
    if (rows <= cap) return DC_OK;
    drop_graphs();
    for (void *p : ws_owned) cudaFree(p);
    ws_owned.clear();
    cap = 0;
    rep_cap = 0;
    rep_g1f = rep_d1f = nullptr;
    const size_t R = (size_t)((rows + 127) / 128 * 128);
    const size_t F = cfg.feat, E = cfg.embed, U = cfg.units, V = cfg.vocab;
    const size_t Wu = cfg.arch == DC_ARCH_V1 ? U : cfg.word_units;
    auto A = [&](float **p, size_t n) { return dev_alloc((void **)p, sizeof(float) * n, ws_owned); };
    int rc = 0;
    rc |= A(&ws.F, R * F); rc |= A(&ws.a1, R * F);
    rc |= A(&ws.g1f, R * 4 * U); rc |= A(&ws.d1f, R * kDense);
    rc |= A(&ws.xh1, R * (E + Wu)); rc |= A(&ws.xh1b, R * (E + Wu));
    rc |= A(&ws.xh2, R * (cfg.arch == DC_ARCH_V1 ? 2 * U : F + Wu));
    rc |= A(&ws.xh2b, R * (cfg.arch == DC_ARCH_V1 ? 2 * U : F + Wu));
    rc |= A(&ws.c1, R * Wu); rc |= A(&ws.c1b, R * Wu);
    rc |= A(&ws.c2, R * U); rc |= A(&ws.c2b, R * U);
    rc |= A(&ws.gates, R * 4 * Wu); rc |= A(&ws.d, R * kDense); rc |= A(&ws.h2, R * U);
    rc |= A(&ws.logits, R * V);
    rc |= A(&ws.cand_p, R * kMaxBeam);
    rc |= dev_alloc((void **)&ws.tok, sizeof(int32_t) * R, ws_owned);
    rc |= dev_alloc((void **)&ws.newtok, sizeof(int32_t) * R, ws_owned);
    rc |= dev_alloc((void **)&ws.parent, sizeof(int32_t) * R, ws_owned);
    rc |= dev_alloc((void **)&ws.cand_idx, sizeof(int32_t) * R * kMaxBeam, ws_owned);
    rc |= dev_alloc((void **)&ws.hist_a, sizeof(int32_t) * R * cfg.padding, ws_owned);
    rc |= dev_alloc((void **)&ws.hist_b, sizeof(int32_t) * R * cfg.padding, ws_owned);
    rc |= dev_alloc((void **)&ws.score_a, sizeof(double) * R, ws_owned);
    rc |= dev_alloc((void **)&ws.score_b, sizeof(double) * R, ws_owned);
    if (rc) return rc;
    if (cfg.dtype == DC_DTYPE_BF16)
        if (int r2 = reserve_bf16(R)) return r2;
    cap = (int)R;
    return DC_OK;
}

// ------------------------------------------------------------------------------------------
// building blocks
// ------------------------------------------------------------------------------------------
int Decoder::linear_f32(const float *x, int ldx, int M, const float *Wk, int K, int N, const float *bias,
                        const float *addend, int ld_addend, const float *scale, const float *shift,
                        bool relu, float *out, int ldo, cudaStream_t s) {
    SgemmArgs g;
    g.A = x; g.lda = ldx; g.B = Wk; g.ldb = N; g.C = out; g.ldc = ldo;
    g.M = M; g.N = N; g.K = K;
    g.bias = bias; g.addend = addend; g.ld_addend = ld_addend; g.scale = scale; g.shift = shift;
    g.relu = relu ? 1 : 0;
    return sgemm(g, false, false, s);
}

__global__ void repeat_rows_kernel(const float *__restrict__ src, int width, int k, int rows,
                                   float *__restrict__ dst) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    const float *sp = src + (long long)(r / k) * width;
    float *dp = dst + (long long)r * width;
    for (int c = threadIdx.x; c < width; c += blockDim.x) dp[c] = sp[c];
}

// head: [B, p*p*C] -> [B, F]
int Decoder::head(const void *feats, int kind, int B, float *out, cudaStream_t s) {
    const int F = cfg.feat, Kin = cfg.pool * cfg.pool * cfg.channels;
    if (kind == DC_FEATS_HEAD_F32) {
        if (out != feats)
            DC_CHECK_CUDA(cudaMemcpyAsync(out, feats, sizeof(float) * (size_t)B * F, cudaMemcpyDeviceToDevice, s));
        return DC_OK;
    }
    if (cfg.dtype == DC_DTYPE_BF16) return head_bf16(feats, kind, B, out, s);
    DC_REQUIRE(kind == DC_FEATS_ROI_F32, "fp32 decoder expects fp32 RoI features (feats_kind=%d)", kind);
    if (int rc = linear_f32((const float *)feats, Kin, B, W("mrcnn_class_conv1/kernel"), Kin, F,
                            W("mrcnn_class_conv1/bias"), nullptr, 0, bn_scale[0], bn_shift[0], true,
                            ws.a1, F, s)) return rc;
    return linear_f32(ws.a1, F, B, W("mrcnn_class_conv2/kernel"), F, F, W("mrcnn_class_conv2/bias"),
                      nullptr, 0, bn_scale[1], bn_shift[1], true, out, F, s);
}

// per-RoI constant terms of the v1 word model: g1f = f*W1[E:] + b1, d1f = f*Wd1[U:] + bd1
int Decoder::v1_hoist(int B, cudaStream_t s) {
    const int F = cfg.feat, E = cfg.embed, U = cfg.units;
    if (cfg.dtype == DC_DTYPE_BF16) return v1_hoist_bf16(B, s);
    if (int rc = linear_f32(ws.F, F, B, W("imgcap_lstm1/kernel") + (size_t)E * 4 * U, F, 4 * U,
                            W("imgcap_lstm1/bias"), nullptr, 0, nullptr, nullptr, false, ws.g1f, 4 * U, s))
        return rc;
    return linear_f32(ws.F, F, B, W("imgcap_lstm_d1/kernel") + (size_t)U * kDense, F, kDense,
                      W("imgcap_lstm_d1/bias"), nullptr, 0, nullptr, nullptr, false, ws.d1f, kDense, s);
}

int Decoder::v1_reset_state(int R, cudaStream_t s) {
    const size_t E = cfg.embed, U = cfg.units;
    DC_CHECK_CUDA(cudaMemsetAsync(ws.xh1, 0, sizeof(float) * R * (E + U), s));
    DC_CHECK_CUDA(cudaMemsetAsync(ws.xh2, 0, sizeof(float) * R * 2 * U, s));
    // whole 32-row blocks: the persistent greedy loop keeps c in a blocked-32 layout (greedy_loop.cu), where row m's state is
    // spread over its block; the workspace is reserved in multiples of 128 rows
    const size_t Rc = ((size_t)R + 31) / 32 * 32;
    DC_CHECK_CUDA(cudaMemsetAsync(ws.c1, 0, sizeof(float) * Rc * U, s));
    DC_CHECK_CUDA(cudaMemsetAsync(ws.c2, 0, sizeof(float) * Rc * U, s));
    if (cfg.dtype == DC_DTYPE_BF16) return reset_state_bf16(R, s);
    return DC_OK;
}

// One incremental step of the v1 word model on R rows: consumes ws.tok, leaves logits in
// ws.logits.  g1f / d1f rows are indexed by row (already replicated per beam if needed).
int Decoder::v1_step(int R, const float *g1f, const float *d1f, cudaStream_t s) {
    const int E = cfg.embed, U = cfg.units, V = cfg.vocab;
    if (cfg.dtype == DC_DTYPE_BF16) return v1_step_bf16(R, g1f, d1f, s);
    if (int rc = embed_gather(W("imgcap_embedding_layer/embeddings"), ws.tok, R, E, V, ws.xh1, E + U, false, s)) return rc;
    if (int rc = linear_f32(ws.xh1, E + U, R, w1cat, E + U, 4 * U, nullptr, g1f, 4 * U, nullptr, nullptr,
                            false, ws.gates, 4 * U, s)) return rc;
    CellArgs c1;
    c1.gates = ws.gates; c1.ld_gates = 4 * U; c1.tok = ws.tok; c1.c_in = ws.c1; c1.c_out = ws.c1;
    c1.h_f32_a = ws.xh1 + E; c1.ld_a = E + U; c1.h_f32_b = ws.xh2; c1.ld_b = 2 * U; c1.rows = R; c1.U = U;
    if (int rc = lstm_cell(c1, s)) return rc;
    if (int rc = linear_f32(ws.xh2, 2 * U, R, w2cat, 2 * U, 4 * U, W("imgcap_lstm2/bias"), nullptr, 0,
                            nullptr, nullptr, false, ws.gates, 4 * U, s)) return rc;
    CellArgs c2;
    c2.gates = ws.gates; c2.ld_gates = 4 * U; c2.tok = ws.tok; c2.c_in = ws.c2; c2.c_out = ws.c2;
    c2.h_f32_a = ws.xh2 + U; c2.ld_a = 2 * U; c2.rows = R; c2.U = U;
    if (int rc = lstm_cell(c2, s)) return rc;
    if (int rc = linear_f32(ws.xh2 + U, 2 * U, R, W("imgcap_lstm_d1/kernel"), U, kDense, nullptr, d1f, kDense,
                            nullptr, nullptr, true, ws.d, kDense, s)) return rc;
    return linear_f32(ws.d, kDense, R, W("imgcap_lstm_d2/kernel"), kDense, V, W("imgcap_lstm_d2/bias"),
                      nullptr, 0, nullptr, nullptr, false, ws.logits, V, s);
}

int Decoder::check_ready(int B) {
    if (!finalized) return set_error(DC_ERR_STATE, "decoder used before dc_decoder_finalize");
    DC_REQUIRE(B >= 0, "negative batch");
    return DC_OK;
}

// scores[r] (+)= log(maxprob[r]): the caption score of refine_generations (sum over steps of log max p)
__global__ void accumulate_log_kernel(float *__restrict__ scores, const float *__restrict__ maxprob, int rows, int first) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) scores[r] = (first ? 0.f : scores[r]) + logf(maxprob[r]);
}

int accumulate_log(float *scores, const float *maxprob, int rows, bool first, cudaStream_t s) {
    accumulate_log_kernel<<<ceil_div(rows, 256), 256, 0, s>>>(scores, maxprob, rows, first ? 1 : 0);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

int Decoder::greedy(const void *feats, int kind, int B, int32_t *tokens, float *probs, cudaStream_t s, float *scores) {
    if (int rc = check_ready(B)) return rc;
    DC_REQUIRE(cfg.arch == DC_ARCH_V1, "dc_decoder_greedy needs a v1 decoder");
    if (B == 0) return DC_OK;
    DC_REQUIRE(feats && tokens, "null pointer argument");
    if (int rc = reserve(B)) return rc;
    const int P = cfg.padding, V = cfg.vocab;
    if (cfg.dtype == DC_DTYPE_BF16 && !probs && !scores) return greedy_bf16_graphed(feats, kind, B, tokens, s);
    if (cfg.dtype == DC_DTYPE_BF16 && !probs) return greedy_bf16(feats, kind, B, tokens, s, scores);
    if (int rc = head(feats, kind, B, ws.F, s)) return rc;
    if (int rc = v1_hoist(B, s)) return rc;
    if (int rc = v1_reset_state(B, s)) return rc;
    if (int rc = fill_i32(ws.tok, B, 1, s)) return rc;                 // <start> = 1
    for (int t = 0; t < P; ++t) {
        if (int rc = v1_step(B, ws.g1f, ws.d1f, s)) return rc;
        if (int rc = softmax_argmax(ws.logits, V, B, V, probs ? probs + (size_t)t * V : nullptr,
                                    (long long)P * V, tokens + t, P, ws.tok, scores ? ws.cand_p : nullptr, s)) return rc;
        if (scores)
            if (int rc = accumulate_log(scores, ws.cand_p, B, t == 0, s)) return rc;
    }
    return DC_OK;
}

int Decoder::beam(const void *feats, int kind, int B, int k, int32_t *tokens, double *scores, cudaStream_t s) {
    if (int rc = check_ready(B)) return rc;
    DC_REQUIRE(cfg.arch == DC_ARCH_V1, "dc_decoder_beam needs a v1 decoder");
    DC_REQUIRE(k >= 1 && k <= kMaxBeam && k <= cfg.vocab, "beam width %d outside [1,%d]", k, kMaxBeam);
    if (B == 0) return DC_OK;
    DC_REQUIRE(feats && tokens && scores, "null pointer argument");
    DC_REQUIRE((long long)B * k < (1ll << 31), "too many beam rows in one call");
    if (cfg.dtype == DC_DTYPE_BF16) return beam_bf16(feats, kind, B, k, tokens, scores, s);
    const int R = B * k;
    if (int rc = reserve(R)) return rc;
    const int P = cfg.padding, V = cfg.vocab, E = cfg.embed, U = cfg.units;
    if (int rc = head(feats, kind, B, ws.F, s)) return rc;
    if (int rc = v1_hoist(B, s)) return rc;
    // replicate the per-RoI constant terms per beam
    if (int rc = ensure_rep(R)) return rc;
    repeat_rows_kernel<<<R, 128, 0, s>>>(ws.g1f, 4 * U, k, R, rep_g1f);
    repeat_rows_kernel<<<R, 128, 0, s>>>(ws.d1f, kDense, k, R, rep_d1f);
    DC_CHECK_LAUNCH();
    if (int rc = v1_reset_state(R, s)) return rc;
    if (int rc = fill_i32(ws.tok, R, 1, s)) return rc;
    DC_CHECK_CUDA(cudaMemsetAsync(ws.score_a, 0, sizeof(double) * R, s));
    DC_CHECK_CUDA(cudaMemsetAsync(ws.hist_a, 0, sizeof(int32_t) * (size_t)R * P, s));
    if (int rc = set_token_column(ws.hist_a, R, P, 0, ws.tok, s)) return rc;
    int32_t *hist = ws.hist_a, *hist_n = ws.hist_b;
    double *sc = ws.score_a, *sc_n = ws.score_b;
    for (int t = 0; t + 1 < P; ++t) {
        if (int rc = v1_step(R, rep_g1f, rep_d1f, s)) return rc;
        if (int rc = topk_softmax(ws.logits, V, R, V, k, ws.cand_idx, ws.cand_p, s)) return rc;
        if (int rc = beam_select(B, k, t == 0 ? 1 : k, ws.cand_idx, ws.cand_p, sc, sc_n, ws.parent, ws.newtok, s)) return rc;
        // children inherit the parent's post-step state and history
        int rc = 0;
        rc |= beam_gather(R, k, U, ws.parent, ws.xh1 + E, E + U, ws.xh1b + E, E + U, 4, s);
        rc |= beam_gather(R, k, 2 * U, ws.parent, ws.xh2, 2 * U, ws.xh2b, 2 * U, 4, s);
        rc |= beam_gather(R, k, U, ws.parent, ws.c1, U, ws.c1b, U, 4, s);
        rc |= beam_gather(R, k, U, ws.parent, ws.c2, U, ws.c2b, U, 4, s);
        rc |= beam_gather(R, k, P, ws.parent, hist, P, hist_n, P, 4, s);
        if (rc) return rc;
        std::swap(ws.xh1, ws.xh1b); std::swap(ws.xh2, ws.xh2b);
        std::swap(ws.c1, ws.c1b); std::swap(ws.c2, ws.c2b);
        std::swap(hist, hist_n); std::swap(sc, sc_n);
        if (int r2 = set_token_column(hist, R, P, t + 1, ws.newtok, s)) return r2;
        DC_CHECK_CUDA(cudaMemcpyAsync(ws.tok, ws.newtok, sizeof(int32_t) * R, cudaMemcpyDeviceToDevice, s));
    }
    DC_CHECK_CUDA(cudaMemcpyAsync(tokens, hist, sizeof(int32_t) * (size_t)R * P, cudaMemcpyDeviceToDevice, s));
    DC_CHECK_CUDA(cudaMemcpyAsync(scores, sc, sizeof(double) * R, cudaMemcpyDeviceToDevice, s));
    return DC_OK;
}

// RoI feature buffer for the fused ROIAlign -> decoder pipeline ([R, pool*pool*channels], bf16 or fp32)
int Decoder::roi_feature_buffer(int R, void **out) {
    const size_t elems = (size_t)R * cfg.pool * cfg.pool * cfg.channels;
    const size_t bytes = elems * (cfg.dtype == DC_DTYPE_BF16 ? 2 : 4);
    if (bytes > roi_buf_bytes) {
        if (roi_buf) cudaFree(roi_buf);
        roi_buf = nullptr; roi_buf_bytes = 0;
        DC_CHECK_CUDA(cudaMalloc(&roi_buf, bytes));
        roi_buf_bytes = bytes;
    }
    *out = roi_buf;
    return DC_OK;
}

int Decoder::ensure_rep(int R) {
    if (R <= rep_cap) return DC_OK;
    const size_t U = cfg.units;
    // growth: the previous (smaller) buffers are released here, not kept on the workspace list until the next reserve()
    for (float **p : {&rep_g1f, &rep_d1f}) {
        if (!*p) continue;
        for (auto it = ws_owned.begin(); it != ws_owned.end(); ++it)
            if (*it == *p) { ws_owned.erase(it); break; }
        cudaFree(*p);
        *p = nullptr;
    }
    rep_cap = 0;
    if (int rc = dev_alloc((void **)&rep_g1f, sizeof(float) * (size_t)R * 4 * U, ws_owned)) return rc;
    if (int rc = dev_alloc((void **)&rep_d1f, sizeof(float) * (size_t)R * kDense, ws_owned)) return rc;
    rep_cap = R;
    return DC_OK;
}

// ------------------------------------------------------------------------------------------
// v2 inject
// ------------------------------------------------------------------------------------------
int Decoder::v2_reset(int B, cudaStream_t s) {
    const size_t E = cfg.embed, Wu = cfg.word_units, F = cfg.feat;
    DC_CHECK_CUDA(cudaMemsetAsync(ws.xh1, 0, sizeof(float) * B * (E + Wu), s));
    DC_CHECK_CUDA(cudaMemsetAsync(ws.c1, 0, sizeof(float) * B * Wu, s));
    // xin = [head feature | word vector]; the word-vector half starts at zero (all-masked prefix)
    DC_CHECK_CUDA(cudaMemset2DAsync(ws.xh2 + F, sizeof(float) * (F + Wu), 0, sizeof(float) * Wu, B, s));
    return DC_OK;
}

// consume ws.tok with the word LSTM (masked), state in xh1[:,E:] / c1; h mirrored into xin[:,F:]
int Decoder::v2_word_step(int B, cudaStream_t s) {
    const int E = cfg.embed, Wu = cfg.word_units, F = cfg.feat, V = cfg.vocab;
    if (int rc = embed_gather(W("imgcap_embedding_layer/embeddings"), ws.tok, B, E, V, ws.xh1, E + Wu, false, s)) return rc;
    if (int rc = linear_f32(ws.xh1, E + Wu, B, w1cat, E + Wu, 4 * Wu, W("lstm_1/bias"), nullptr, 0, nullptr,
                            nullptr, false, ws.gates, 4 * Wu, s)) return rc;
    CellArgs c;
    c.gates = ws.gates; c.ld_gates = 4 * Wu; c.tok = ws.tok; c.c_in = ws.c1; c.c_out = ws.c1;
    c.h_f32_a = ws.xh1 + E; c.ld_a = E + Wu; c.h_f32_b = ws.xh2 + F; c.ld_b = F + Wu; c.rows = B; c.U = Wu;
    return lstm_cell(c, s);
}

// [head ; wv] -> LSTM(units) one step from the zero state -> Dense(V) logits in ws.logits
int Decoder::v2_output(int B, cudaStream_t s) {
    const int Wu = cfg.word_units, F = cfg.feat, U = cfg.units, V = cfg.vocab;
    if (int rc = linear_f32(ws.xh2, F + Wu, B, W("imgcap_lstm/kernel"), F + Wu, 4 * U, W("imgcap_lstm/bias"),
                            nullptr, 0, nullptr, nullptr, false, ws.gates, 4 * U, s)) return rc;
    CellArgs c;
    c.gates = ws.gates; c.ld_gates = 4 * U; c.c_in = nullptr; c.c_out = ws.c2;
    c.h_f32_a = ws.h2; c.ld_a = U; c.rows = B; c.U = U;
    if (int rc = lstm_cell(c, s)) return rc;
    return linear_f32(ws.h2, U, B, W("imgcap_d1/kernel"), U, V, W("imgcap_d1/bias"), nullptr, 0, nullptr,
                      nullptr, false, ws.logits, V, s);
}

int Decoder::v2_head_into_xin(const void *feats, int kind, int B, cudaStream_t s) {
    const size_t F = cfg.feat, Wu = cfg.word_units;
    if (int rc = head(feats, kind, B, ws.F, s)) return rc;
    DC_CHECK_CUDA(cudaMemcpy2DAsync(ws.xh2, sizeof(float) * (F + Wu), ws.F, sizeof(float) * F,
                                    sizeof(float) * F, B, cudaMemcpyDeviceToDevice, s));
    return DC_OK;
}

int Decoder::v2_predict(const void *feats, int kind, const int32_t *words, int B, int L, float *probs,
                        cudaStream_t s) {
    if (int rc = check_ready(B)) return rc;
    DC_REQUIRE(cfg.arch == DC_ARCH_V2_INJECT, "dc_decoder_v2_predict needs a v2 inject decoder");
    DC_REQUIRE(L >= 0, "negative sequence length");
    if (B == 0) return DC_OK;
    DC_REQUIRE(feats && probs && (words || L == 0), "null pointer argument");
    if (int rc = reserve(B)) return rc;
    if (cfg.dtype == DC_DTYPE_BF16) return v2_predict_bf16(feats, kind, words, B, L, probs, s);
    if (int rc = v2_reset(B, s)) return rc;
    if (int rc = v2_head_into_xin(feats, kind, B, s)) return rc;
    for (int t = 0; t < L; ++t) {
        if (int rc = token_column(words, B, L, t, ws.tok, s)) return rc;
        if (int rc = v2_word_step(B, s)) return rc;
    }
    if (int rc = v2_output(B, s)) return rc;
    return softmax_argmax(ws.logits, cfg.vocab, B, cfg.vocab, probs, cfg.vocab, nullptr, 0, nullptr, nullptr, s);
}

int Decoder::v2_greedy(const void *feats, int kind, int B, int32_t *tokens, float *probs, cudaStream_t s, const int32_t *start,
                       float *scores) {
    if (int rc = check_ready(B)) return rc;
    DC_REQUIRE(cfg.arch == DC_ARCH_V2_INJECT, "dc_decoder_v2_greedy needs a v2 inject decoder");
    if (B == 0) return DC_OK;
    DC_REQUIRE(feats && tokens, "null pointer argument");
    if (int rc = reserve(B)) return rc;
    if (cfg.dtype == DC_DTYPE_BF16) return v2_greedy_bf16(feats, kind, B, tokens, probs, s, start, scores);
    const int P = cfg.padding, V = cfg.vocab;
    if (int rc = v2_reset(B, s)) return rc;
    if (int rc = v2_head_into_xin(feats, kind, B, s)) return rc;
    if (start) {                                                 // eval_text_generation_model_v2.py:176-186: prev = [gt[0]]
        DC_CHECK_CUDA(cudaMemcpyAsync(ws.tok, start, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, s));
    } else if (int rc = fill_i32(ws.tok, B, 0, s)) return rc;   // argmax(zeros(V)) = 0 -> masked
    for (int t = 0; t + 1 < P; ++t) {
        // the padded window never truncates: the sequence has at most P-1 ids,
        // so consuming one new id per step equals re-running the LSTM over the whole prefix
        if (int rc = v2_word_step(B, s)) return rc;
        if (int rc = v2_output(B, s)) return rc;
        if (int rc = softmax_argmax(ws.logits, V, B, V, probs ? probs + (size_t)t * V : nullptr,
                                    (long long)(P - 1) * V, tokens + t, P - 1, ws.tok, scores ? ws.cand_p : nullptr, s)) return rc;
        if (scores)
            if (int rc = accumulate_log(scores, ws.cand_p, B, t == 0, s)) return rc;
    }
    return DC_OK;
}

}  // namespace dcap

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace dcap;


extern "C" int dc_decoder_create(const DcDecoderConfig *cfg, DcDecoder **out) {
    DC_REQUIRE(cfg && out, "null pointer argument");
    DC_REQUIRE(cfg->arch == DC_ARCH_V1 || cfg->arch == DC_ARCH_V2_INJECT, "unknown arch %d", cfg->arch);
    DC_REQUIRE(cfg->dtype == DC_DTYPE_F32 || cfg->dtype == DC_DTYPE_BF16, "unknown dtype %d", cfg->dtype);
    DC_REQUIRE(cfg->vocab > 1 && cfg->embed > 0 && cfg->feat > 0 && cfg->units > 0 && cfg->pool > 0 &&
               cfg->channels > 0 && cfg->padding > 0, "non-positive decoder dimension");
    DC_REQUIRE(cfg->arch == DC_ARCH_V1 || cfg->word_units > 0, "v2 needs word_units > 0");
    if (cfg->dtype == DC_DTYPE_BF16) {
        DC_REQUIRE(cfg->units % 64 == 0 && cfg->feat % 64 == 0 && (cfg->pool * cfg->pool * cfg->channels) % 64 == 0,
                   "bf16 path needs units, feat and pool*pool*channels to be multiples of 64");
        DC_REQUIRE(cfg->arch == DC_ARCH_V1 || cfg->word_units % 64 == 0, "bf16 v2 path needs word_units %% 64 == 0");
    }
    int dev = 0;
    DC_CHECK_CUDA(cudaGetDevice(&dev));
    DcDecoder *d = new DcDecoder();
    d->impl.cfg = *cfg;
    d->impl.device = dev;
    d->impl.declare_all();
    if (int rc = d->impl.layout_arena()) { delete d; return rc; }
    *out = d;
    return DC_OK;
}

extern "C" int dc_decoder_destroy(DcDecoder *dec) {
    delete dec;
    return DC_OK;
}

extern "C" int dc_decoder_weight_count(const DcDecoder *dec) {
    return dec ? (int)dec->impl.weights.size() : 0;
}
extern "C" const char *dc_decoder_weight_name(const DcDecoder *dec, int i) {
    if (!dec || i < 0 || i >= (int)dec->impl.weights.size()) return nullptr;
    return dec->impl.weights[i].name.c_str();
}
extern "C" int64_t dc_decoder_weight_numel(const DcDecoder *dec, int i) {
    if (!dec || i < 0 || i >= (int)dec->impl.weights.size()) return -1;
    return dec->impl.weights[i].numel;
}

extern "C" int dc_decoder_set_weight(DcDecoder *dec, const char *name, const float *host, int64_t numel) {
    DC_REQUIRE(dec && name && host, "null pointer argument");
    Weight *w = dec->impl.find(name);
    DC_REQUIRE(w != nullptr, "unknown weight '%s'", name);
    DC_REQUIRE(w->numel == numel, "weight '%s' expects %lld values, got %lld", name, (long long)w->numel,
               (long long)numel);
    DC_CHECK_CUDA(cudaMemcpy(w->dev, host, sizeof(float) * (size_t)numel, cudaMemcpyHostToDevice));
    w->is_set = true;
    if (w->name.find("/embeddings") != std::string::npos) dec->impl.emb_dirty = true;
    dec->impl.finalized = false;
    return DC_OK;
}

extern "C" int dc_decoder_get_weight(DcDecoder *dec, const char *name, float *host, int64_t numel) {
    DC_REQUIRE(dec && name && host, "null pointer argument");
    Weight *w = dec->impl.find(name);
    DC_REQUIRE(w != nullptr, "unknown weight '%s'", name);
    DC_REQUIRE(w->numel == numel, "weight '%s' holds %lld values, got %lld", name, (long long)w->numel,
               (long long)numel);
    if (!w->is_set) return set_error(DC_ERR_STATE, "weight '%s' has not been set", name);
    DC_CHECK_CUDA(cudaMemcpy(host, w->dev, sizeof(float) * (size_t)numel, cudaMemcpyDeviceToHost));
    return DC_OK;
}

extern "C" int dc_decoder_finalize(DcDecoder *dec, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.finalize((cudaStream_t)stream);
}

extern "C" int dc_head_forward(DcDecoder *dec, const void *feats, int kind, int B, float *out, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    if (int rc = dec->impl.check_ready(B)) return rc;
    if (B == 0) return DC_OK;
    DC_REQUIRE(feats && out, "null pointer argument");
    if (int rc = dec->impl.reserve(B)) return rc;
    return dec->impl.head(feats, kind, B, out, (cudaStream_t)stream);
}

extern "C" int dc_decoder_greedy(DcDecoder *dec, const void *feats, int kind, int B, int32_t *tokens,
                                 float *probs, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.greedy(feats, kind, B, tokens, probs, (cudaStream_t)stream);
}

extern "C" int dc_decoder_greedy_scored(DcDecoder *dec, const void *feats, int kind, int B, int32_t *tokens,
                                        float *scores, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    DC_REQUIRE(scores || B == 0, "null pointer argument");
    return dec->impl.greedy(feats, kind, B, tokens, nullptr, (cudaStream_t)stream, scores);
}

extern "C" int dc_decoder_beam(DcDecoder *dec, const void *feats, int kind, int B, int k, int32_t *tokens,
                               double *scores, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.beam(feats, kind, B, k, tokens, scores, (cudaStream_t)stream);
}

extern "C" int dc_decoder_v2_predict(DcDecoder *dec, const void *feats, int kind, const int32_t *words,
                                     int B, int L, float *probs, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.v2_predict(feats, kind, words, B, L, probs, (cudaStream_t)stream);
}

extern "C" int dc_decoder_v2_greedy(DcDecoder *dec, const void *feats, int kind, int B, int32_t *tokens,
                                    float *probs, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.v2_greedy(feats, kind, B, tokens, probs, (cudaStream_t)stream);
}

extern "C" int dc_decoder_v2_greedy_from(DcDecoder *dec, const void *feats, int kind, int B, const int32_t *start,
                                         int32_t *tokens, float *probs, float *scores, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.v2_greedy(feats, kind, B, tokens, probs, (cudaStream_t)stream, start, scores);
}

extern "C" int dc_decoder_greedy_host(DcDecoder *dec, const float *feats, int kind, int B, int32_t *tokens,
                                      float *probs) {
    DC_REQUIRE(dec, "null decoder");
    if (int rc = dec->impl.check_ready(B)) return rc;
    if (B == 0) return DC_OK;
    DC_REQUIRE(feats && tokens, "null pointer argument");
    DC_REQUIRE(kind == DC_FEATS_ROI_F32 || kind == DC_FEATS_HEAD_F32, "host features must be fp32");
    const DcDecoderConfig &c = dec->impl.cfg;
    const size_t in_elems = (size_t)B * (kind == DC_FEATS_HEAD_F32 ? c.feat : c.pool * c.pool * c.channels);
    const size_t P = c.padding, V = c.vocab;
    cudaStream_t s;
    DC_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    float *d_in = nullptr, *d_probs = nullptr;
    int32_t *d_tok = nullptr;
    int rc = DC_OK;
    cudaError_t e = cudaMallocAsync((void **)&d_in, sizeof(float) * in_elems, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_tok, sizeof(int32_t) * B * P, s);
    if (e == cudaSuccess && probs) e = cudaMallocAsync((void **)&d_probs, sizeof(float) * B * P * V, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, feats, sizeof(float) * in_elems, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) rc = set_error(DC_ERR_CUDA, "greedy_host staging failed: %s", cudaGetErrorString(e));
    if (rc == DC_OK) rc = dec->impl.greedy(d_in, kind, B, d_tok, d_probs, s);
    if (rc == DC_OK) {
        e = cudaMemcpyAsync(tokens, d_tok, sizeof(int32_t) * B * P, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && probs)
            e = cudaMemcpyAsync(probs, d_probs, sizeof(float) * B * P * V, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = set_error(DC_ERR_CUDA, "greedy_host failed: %s", cudaGetErrorString(e));
    }
    if (d_in) cudaFreeAsync(d_in, s);
    if (d_tok) cudaFreeAsync(d_tok, s);
    if (d_probs) cudaFreeAsync(d_probs, s);
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc;
}

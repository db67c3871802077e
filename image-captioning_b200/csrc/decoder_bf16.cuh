// State of the bf16 / tcgen05 decoder path (decoder_bf16.cu) and of the training step (train.cu).
#pragma once
#include "decoder.cuh"
#include "gemm_tc.cuh"

namespace dcap {

struct TrainState;                // train.cu

struct Bf16State {
    // weights (bf16, K-major)
    __nv_bfloat16 *w_head1 = nullptr, *w_head2 = nullptr;
    __nv_bfloat16 *w1cat = nullptr, *w1f = nullptr, *w2cat = nullptr, *wd1h = nullptr, *wd1f = nullptr, *wd2 = nullptr;
    float *b1_i = nullptr, *b2_i = nullptr;          // gate-interleaved LSTM biases
    float *bias_hoist = nullptr;                     // [4U + 1024] = [b1 (gate-interleaved) | bd1]: bias of the merged hoist GEMM (w1f and wd1f are adjacent)
    float *hoist_all = nullptr;                      // [R, 4U + 1024] fp32, blocked-32: [f W1f + b1 | f Wd1f + bd1] for greedy_loop.cu
    __nv_bfloat16 *emb = nullptr;                    // [V, Epad] bf16 embedding table (zero padded)
    int Epad = 0;
    // activations
    __nv_bfloat16 *roi = nullptr, *a1 = nullptr, *Fb = nullptr, *d = nullptr;
    __nv_bfloat16 *X1[2] = {nullptr, nullptr}, *X2[2] = {nullptr, nullptr};
    float *partial = nullptr;
    float *topk_partial = nullptr;                   // beam search: [rows, slots, 2 + 2k]
    size_t topk_cap = 0;
    int parity = 0;
    unsigned int *loop_cnt = nullptr;                // greedy_loop.cu: dependency counters of the persistent decoding kernel
    size_t loop_cnt_bytes = 0;
    unsigned int *loop_err_host = nullptr;           // pinned: the kernel's watchdog word of the latest finished launch (checked at the next call)
    ~Bf16State() { if (loop_err_host) cudaFreeHost(loop_err_host); }
    // backward-pass operand copies: a bf16 mirror of the whole trainable arena (same offsets; written by the
    // optimiser kernel itself, or by one cast after set_weights); the Keras [in, out] tensors inside it are
    // the K-major B operands of dX = dY * W^T (N = in, K = out).  Pointers set by refresh_train_weights().
    __nv_bfloat16 *arena_k = nullptr;
    bool arena_k_valid = false;
    __nv_bfloat16 *wd2_k = nullptr;      // [1024, V]   imgcap_lstm_d2/kernel
    __nv_bfloat16 *wd1h_k = nullptr;     // [U, 1024]   imgcap_lstm_d1/kernel[:U]
    __nv_bfloat16 *wd1f_k = nullptr;     // [F, 1024]   imgcap_lstm_d1/kernel[U:]
    __nv_bfloat16 *w2cat_k = nullptr;    // [2U, 4U]    [imgcap_lstm2/kernel ; recurrent_kernel]
    __nv_bfloat16 *u1_k = nullptr;       // [U, 4U]     imgcap_lstm1/recurrent_kernel
    __nv_bfloat16 *w1f_k = nullptr;      // [F, 4U]     imgcap_lstm1/kernel[E:]
    __nv_bfloat16 *wc2_k = nullptr;      // [F, F]      mrcnn_class_conv2/kernel
    // v2 inject model (text_generation_model_v2.py:140-166) on the tensor-core path
    __nv_bfloat16 *v2_w1cat = nullptr, *v2_wimg = nullptr, *v2_wd = nullptr;   // [4Wu, Epad+Wu], [4U, F+Wu] gate-interleaved; [V, U]
    float *v2_bw = nullptr, *v2_bimg = nullptr;                                 // gate-interleaved biases
    __nv_bfloat16 *v2_xin = nullptr, *v2_hb = nullptr;                          // [R, F+Wu] = [head | word vector], [R, U]
    float *v2_czero = nullptr;                                                  // [R, U] zeros: the image LSTM starts from the zero state
    TrainState *train = nullptr;
};

// dst[n, k_off + k] = bf16(src[(row_off + k) * n_src + colmap(n)]): a Keras [in, out] kernel re-laid as a K-major [N, K] TMA operand
int build_kmajor(const float *src, int n_src, int row_off, int K, int N, int interleave_units, __nv_bfloat16 *dst,
                 long long ld_dst, int k_off, cudaStream_t s);

inline TcOperand tc_op(const __nv_bfloat16 *p, long long ld, bool mn_major = false) {
    TcOperand o;
    o.ptr = p; o.ld = ld; o.mn_major = mn_major;
    return o;
}

}  // namespace dcap

// Element-wise / reduction kernels of the decoder (see decoder_kernels.cu).
#pragma once
#include "common.cuh"

namespace dcap {

constexpr int kMaxBeam = 8;

struct CellArgs {
    const float *gates = nullptr; int ld_gates = 0;     // [rows, 4U] pre-activations (i|f|c|o)
    const float *gates2 = nullptr; int ld_gates2 = 0;   // optional second addend
    const int32_t *tok = nullptr;                       // consumed token per row (0 = masked) or null
    const float *c_in = nullptr;                        // [rows, U] (null = zero state)
    float *c_out = nullptr;                             // [rows, U]
    float *h_f32_a = nullptr, *h_f32_b = nullptr;       // destinations of h (row stride ld_a / ld_b)
    __nv_bfloat16 *h_bf16_a = nullptr, *h_bf16_b = nullptr;
    int ld_a = 0, ld_b = 0;
    float *save_act = nullptr;                          // [rows, 5U] (i,f,g,o,tanh c) for backward
    int rows = 0, U = 0;
};

int embed_gather(const float *emb, const int32_t *tok, int rows, int E, int V, void *out, int ld,
                 bool bf16, cudaStream_t s);
int lstm_cell(const CellArgs &a, cudaStream_t s);
int softmax_argmax(const float *logits, int ld, int rows, int V, float *probs, long long ld_probs,
                   int32_t *tok_out, int tok_stride, int32_t *tok_cur, float *maxprob, cudaStream_t s);
int topk_softmax(const float *logits, int ld, int rows, int V, int k, int32_t *idx_out, float *p_out,
                 cudaStream_t s);
int beam_select(int n_roi, int k, int n_active, const int32_t *cand_idx, const float *cand_p,
                const double *score_in, double *score_out, int32_t *parent, int32_t *new_tok,
                cudaStream_t s, bool compact = false);
int beam_gather(int rows, int k, int width, const int32_t *parent, const void *src, int ld_src,
                void *dst, int ld_dst, int elem_bytes, cudaStream_t s);
int fill_i32(int32_t *p, long long n, int32_t v, cudaStream_t s);
int set_token_column(int32_t *tokens, int rows, int stride, int col, const int32_t *src, cudaStream_t s);
int token_column(const int32_t *words, int rows, int L, int col, int32_t *tok, cudaStream_t s);
int f32_to_bf16(const float *src, void *dst, long long n, cudaStream_t s);
int transpose_f32(const float *in, int rows, int cols, void *out, int ld_out, bool bf16, cudaStream_t s);

}  // namespace dcap

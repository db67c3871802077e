// bf16 tcgen05 GEMM (see gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace dcap {

constexpr int kEpiStore = 0;
constexpr int kEpiArgmax = 1;       // max + first arg-max
constexpr int kEpiArgmaxSum = 2;    // ... + sum exp(v - max) (softmax probability of the arg-max)
constexpr int kEpiCell = 3;         // fused Keras LSTM cell on gate-interleaved columns

struct TcOperand {
    const __nv_bfloat16 *ptr = nullptr;   // [rows, K] row-major, K contiguous
    long long ld = 0;                     // elements; multiple of 8
};

struct TcEpilogue {
    const float *bias = nullptr;          // [N]
    const float *addend = nullptr;        // [M, ld_addend] fp32
    long long ld_addend = 0;
    const float *scale = nullptr;         // [N] (with shift): frozen BatchNorm
    const float *shift = nullptr;
    int relu = 0;
    float *out_f32 = nullptr; long long ld_f32 = 0;
    __nv_bfloat16 *out_bf16 = nullptr; long long ld_bf16 = 0;
    float *partial = nullptr;             // arg-max epilogues: [M, slots] float4 {max, argmax bits, sumexp, -}
    // kEpiCell: column n = 4*unit + gate (i,f,g,o); z = acc + addend + bias
    float *cell_c = nullptr;              // [M, cell_units] fp32, updated in place
    int cell_units = 0;
    const int32_t *cell_tok = nullptr;    // consumed token per row (0 = masked: carry h, c) or null
    const __nv_bfloat16 *cell_h_prev = nullptr; long long ld_h_prev = 0;   // previous h (for masked rows)
    __nv_bfloat16 *cell_h_a = nullptr; long long ld_h_a = 0;               // destinations of the new h
    __nv_bfloat16 *cell_h_b = nullptr; long long ld_h_b = 0;
};

// D = epilogue(A * B^T): A [M,K], B [N,K], both bf16 K-major.
int gemm_bf16_tc(const TcOperand &A, const TcOperand &B, const TcEpilogue &ep, int M, int N, int K, int epi,
                 cudaStream_t stream);
int gemm_tc_argmax_tiles(int N);
int argmax_merge(const float *partial, int rows, int tiles, int32_t *tok_out, int tok_stride, int32_t *tok_cur,
                 float *maxprob, cudaStream_t s, const __nv_bfloat16 *emb = nullptr, int emb_ld = 0,
                 __nv_bfloat16 *x_out = nullptr, long long ld_x = 0);

}  // namespace dcap

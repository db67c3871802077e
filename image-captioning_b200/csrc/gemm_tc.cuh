// bf16 tcgen05 GEMM (see gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace dcap {

constexpr int kEpiStore = 0;
constexpr int kEpiArgmax = 1;       // max + first arg-max
constexpr int kEpiArgmaxSum = 2;    // ... + sum exp(v - max) (softmax probability of the arg-max)
constexpr int kEpiCell = 3;         // fused Keras LSTM cell on gate-interleaved columns
constexpr int kEpiTopK = 4;         // per row and column region: max, sum exp, the k best (value, index) pairs (beam search)
constexpr int kTopKMax = 8;
// internal variants of kEpiStore chosen by gemm_bf16_tc: the output leaves through shared memory + TMA bulk stores
constexpr int kEpiStoreTmaF32 = 5;
constexpr int kEpiStoreTmaB16 = 6;
// internal variant of kEpiCell: 128-wide tiles, the tile's addend and cell state are staged in shared memory by TMA
constexpr int kEpiCellTma = 7;

struct TcOperand {
    // K-major (default): [rows, K] row-major, K contiguous, ld = row stride.
    // MN-major: the operand as it lies in memory when the GEMM contracts over ROWS, i.e.
    // [K, rows] row-major (rows contiguous), ld = stride between consecutive k.  This is what the
    // weight-gradient GEMMs X^T * dY consume: no activation is ever transposed in memory.
    const __nv_bfloat16 *ptr = nullptr;
    long long ld = 0;                     // elements; multiple of 8
    bool mn_major = false;
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
    // kEpiStore extras (training)
    int addend_mod = 0;                   // > 0: addend row = m % addend_mod (per-RoI term broadcast over time)
    int addend_div = 0;                   // > 0: addend row = m / addend_div (per-RoI term shared by the beams of a RoI)
    int deint_units = 0;                  // > 0: fp32 output column 4u+g is written to column g*units+u
    int addend_blocked32 = 0;             // the addend lies in the blocked-32 layout [row / 32][col / 4][row % 32][4] (row = m, m % mod or m / div;
                                          // ld_addend = its row length): lane = row loads of neighbouring rows coalesce
    int blocked32 = 0;                    // fp32 output in the blocked-32 layout [m / 32][n / 4][m % 32][4] (greedy_loop.cu; rows padded to 32)
    const __nv_bfloat16 *mask_src = nullptr; long long ld_mask = 0;   // v = mask_src[m,n] > 0 ? v : 0 (ReLU backward)
    int atomic = 0;                       // fp32 output is accumulated with red.global.add (implied by split-K)
    float *partial = nullptr;             // arg-max epilogues: [M, slots] float4 {max, argmax bits, sumexp, -}
                                          // top-k epilogue: [M, slots, 2 + 2*topk] {max, sumexp, (value, index bits) x topk}
    int topk = 0;
    // kEpiCell: column n = 4*unit + gate (i,f,g,o); z = acc + addend + bias
    float *cell_c = nullptr;              // [M, cell_units] fp32, updated in place
    int cell_units = 0;
    const int32_t *cell_tok = nullptr;    // consumed token per row (0 = masked: carry h, c) or null
    const __nv_bfloat16 *cell_h_prev = nullptr; long long ld_h_prev = 0;   // previous h (for masked rows)
    __nv_bfloat16 *cell_h_a = nullptr; long long ld_h_a = 0;               // destinations of the new h
    __nv_bfloat16 *cell_h_b = nullptr; long long ld_h_b = 0;
    float *cell_c_out = nullptr;          // null: c updated in place; else new c written here (training: time-major)
    __nv_bfloat16 *cell_gates_out = nullptr; long long ld_gates_out = 0;   // optional [M, 4*units] bf16 post-activation (i,f,g,o)
};

// D = epilogue(A * B^T): A [M,K], B [N,K], both bf16 K-major.
// split_k: 0 = choose automatically (only kEpiStore with ep.atomic may split), 1 = never split.
int gemm_bf16_tc(const TcOperand &A, const TcOperand &B, const TcEpilogue &ep, int M, int N, int K, int epi,
                 cudaStream_t stream, int split_k = 1);
int gemm_tc_argmax_tiles(int N);
// merges the kEpiTopK partials: per row the k largest softmax probabilities in ASCENDING order (ties: the larger
// index ranks higher, as a stable ascending argsort followed by [-k:]) -> idx_out / p_out [rows, k]
int topk_merge(const float *partial, int rows, int slots, int k, int32_t *idx_out, float *p_out, cudaStream_t s);
int argmax_merge(const float *partial, int rows, int tiles, int32_t *tok_out, int tok_stride, int32_t *tok_cur,
                 float *maxprob, cudaStream_t s, const __nv_bfloat16 *emb = nullptr, int emb_ld = 0,
                 __nv_bfloat16 *x_out = nullptr, long long ld_x = 0);

}  // namespace dcap

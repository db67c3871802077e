// GEMM entry points shared by the decoder (fp32 SIMT path and bf16 tcgen05 path).
#pragma once
#include "common.cuh"

namespace dcap {

struct SgemmArgs {
    const float *A = nullptr; int lda = 0;
    const float *B = nullptr; int ldb = 0;
    float *C = nullptr; int ldc = 0;
    int M = 0, N = 0, K = 0;
    const float *bias = nullptr;        // [N]
    const float *addend = nullptr;      // [M, ld_addend]
    int ld_addend = 0;
    const float *scale = nullptr;       // [N] frozen-BN scale (with shift)
    const float *shift = nullptr;
    int relu = 0;
    int accumulate = 0;                 // C += ...
};

// C = epilogue(op(A) * op(B)); see gemm_simt.cu.
int sgemm(const SgemmArgs &g, bool transA, bool transB, cudaStream_t stream);

}  // namespace dcap

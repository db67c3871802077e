// fp32 SIMT GEMM (FFMA) -- the "fp32 mode" of the decoder: bit-level behaviour of an fp32 model
// (greedy token ids must equal the fp32 CPU oracle), so no tensor-core rounding of the inputs.
// The bf16 performance path is gemm_tc.cu (tcgen05/TMEM/TMA).
//
//   C[M,N] = epilogue( op(A)[M,K] * op(B)[K,N] )
//   op(A) = A (row-major [M,K], lda) or A^T (A stored [K,M], lda)
//   op(B) = B (row-major [K,N], ldb) or B^T (B stored [N,K], ldb)
//   epilogue: v = acc (+ C_old if accumulate) (+ bias[n]) (+ addend[m,n]);
//             v = v*scale[n] + shift[n] (frozen BatchNorm); v = relu(v) -- each optional.
//
// 128x128x16 tiles, 256 threads, 8x8 micro-tile, double-buffered shared memory.
#include "common.cuh"
#include "gemm.cuh"

namespace dcap {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NT = 256;

// Loads a [rows=BK][cols=BMN] tile slice into registers: element (k, mn).
// `trans` == false: source is [MN, K] row-major (K contiguous)  -> element (mn, k) at src[mn*ld + k]
// `trans` == true : source is [K, MN] row-major (MN contiguous) -> element (k, mn) at src[k*ld + mn]
template <bool kContigK>
__device__ __forceinline__ void load_tile(const float *__restrict__ src, int ld, int mn0, int k0,
                                          int MN, int K, int tid, float (&reg)[8]) {
    if constexpr (kContigK) {
        // 128 rows (mn) x 16 k: thread -> row = tid/2 (0..127), k-half = (tid%2)*8
        const int r = tid >> 1, kk = (tid & 1) * 8;
        const int mn = mn0 + r;
        const float *p = src + (long long)mn * ld + k0 + kk;
        const bool row_ok = mn < MN;
        if (row_ok && k0 + kk + 8 <= K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
            reg[0] = a.x; reg[1] = a.y; reg[2] = a.z; reg[3] = a.w;
            reg[4] = b.x; reg[5] = b.y; reg[6] = b.z; reg[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) reg[i] = (row_ok && k0 + kk + i < K) ? __ldg(p + i) : 0.f;
        }
    } else {
        // 16 k x 128 mn: thread -> k = tid/16 (0..15), mn-chunk = (tid%16)*8
        const int kk = tid >> 4, c = (tid & 15) * 8;
        const int k = k0 + kk;
        const float *p = src + (long long)k * ld + mn0 + c;
        const bool k_ok = k < K;
        if (k_ok && mn0 + c + 8 <= MN && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
            reg[0] = a.x; reg[1] = a.y; reg[2] = a.z; reg[3] = a.w;
            reg[4] = b.x; reg[5] = b.y; reg[6] = b.z; reg[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) reg[i] = (k_ok && mn0 + c + i < MN) ? __ldg(p + i) : 0.f;
        }
    }
}

template <bool kContigK>
__device__ __forceinline__ void store_tile(float (*s)[BM + 4], int tid, const float (&reg)[8]) {
    if constexpr (kContigK) {
        const int r = tid >> 1, kk = (tid & 1) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) s[kk + i][r] = reg[i];
    } else {
        const int kk = tid >> 4, c = (tid & 15) * 8;
        *reinterpret_cast<float4 *>(&s[kk][c]) = make_float4(reg[0], reg[1], reg[2], reg[3]);
        *reinterpret_cast<float4 *>(&s[kk][c + 4]) = make_float4(reg[4], reg[5], reg[6], reg[7]);
    }
}

template <bool kTransA, bool kTransB>
__global__ void __launch_bounds__(NT) sgemm_kernel(const SgemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads; micro-tile rows ty*8.., cols tx*8..
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    const int nk = (g.K + BK - 1) / BK;
    // A tile: element (k, m).  Not transposed: A[m,k] (K contiguous).  Transposed: A stored [K,M].
    load_tile<!kTransA>(g.A, g.lda, m0, 0, g.M, g.K, tid, ra);
    // B tile: element (k, n).  Not transposed: B[k,n] (N contiguous).  Transposed: B stored [N,K].
    load_tile<kTransB>(g.B, g.ldb, n0, 0, g.N, g.K, tid, rb);
    store_tile<!kTransA>(As[0], tid, ra);
    store_tile<kTransB>(Bs[0], tid, rb);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_tile<!kTransA>(g.A, g.lda, m0, (kt + 1) * BK, g.M, g.K, tid, ra);
            load_tile<kTransB>(g.B, g.ldb, n0, (kt + 1) * BK, g.N, g.K, tid, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile<!kTransA>(As[cur ^ 1], tid, ra);
            store_tile<kTransB>(Bs[cur ^ 1], tid, rb);
        }
        __syncthreads();
    }
    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            float *c = g.C + (long long)m * g.ldc + n;
            if (g.accumulate) v += *c;
            if (g.bias) v += __ldg(g.bias + n);
            if (g.addend) v += __ldg(g.addend + (long long)m * g.ld_addend + n);
            if (g.scale) v = v * __ldg(g.scale + n) + __ldg(g.shift + n);
            if (g.relu) v = fmaxf(v, 0.f);
            *c = v;
        }
    }
}

int sgemm(const SgemmArgs &g, bool transA, bool transB, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0) return DC_OK;
    DC_REQUIRE(g.K > 0 && g.A && g.B && g.C, "sgemm: bad arguments");
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM));
    if (!transA && !transB) sgemm_kernel<false, false><<<grid, NT, 0, stream>>>(g);
    else if (!transA && transB) sgemm_kernel<false, true><<<grid, NT, 0, stream>>>(g);
    else if (transA && !transB) sgemm_kernel<true, false><<<grid, NT, 0, stream>>>(g);
    else sgemm_kernel<true, true><<<grid, NT, 0, stream>>>(g);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

}  // namespace dcap

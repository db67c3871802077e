// Error reporting and device introspection for libdcap.so.
#include "common.cuh"
#include "gemm.cuh"
#include "gemm_tc.cuh"
#include <string.h>
#include <stdlib.h>

namespace dcap {

char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

static thread_local bool tls_pdl_scope = false;

bool pdl_enabled() {
    static const bool on = !(getenv("DCAP_PDL") && atoi(getenv("DCAP_PDL")) == 0);
    return on && tls_pdl_scope;
}

PdlScope::PdlScope() : prev(tls_pdl_scope) { tls_pdl_scope = true; }
PdlScope::~PdlScope() { tls_pdl_scope = prev; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace dcap

extern "C" const char *dc_last_error(void) { return dcap::err_buf(); }

extern "C" int dc_device_info(int *sm_count, int *compute_capability) {
    int dev = 0;
    DC_CHECK_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    DC_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DC_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (compute_capability) *compute_capability = major * 10 + minor;
    // keep stream-ordered allocations cached between calls (host-buffer entry points)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return DC_OK;
}

using namespace dcap;

extern "C" int dc_gemm_f32(const float *A, int64_t lda, int trans_a, const float *B, int64_t ldb, int trans_b,
                           int M, int N, int K, const float *bias, const float *addend, int64_t ld_addend,
                           int relu, int accumulate, float *C, int64_t ldc, void *stream) {
    SgemmArgs g;
    g.A = A; g.lda = (int)lda; g.B = B; g.ldb = (int)ldb; g.C = C; g.ldc = (int)ldc;
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.addend = addend; g.ld_addend = (int)ld_addend;
    g.relu = relu; g.accumulate = accumulate;
    return sgemm(g, trans_a != 0, trans_b != 0, (cudaStream_t)stream);
}

extern "C" int dc_gemm_bf16(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M, int N, int K,
                            const float *bias, const float *addend, int64_t ld_addend, int relu,
                            float *out_f32, int64_t ld_f32, uint16_t *out_bf16, int64_t ld_bf16, void *stream) {
    TcOperand a, b;
    a.ptr = reinterpret_cast<const __nv_bfloat16 *>(A); a.ld = lda;
    b.ptr = reinterpret_cast<const __nv_bfloat16 *>(Bt); b.ld = ldb;
    TcEpilogue ep;
    ep.bias = bias; ep.addend = addend; ep.ld_addend = ld_addend; ep.relu = relu;
    ep.out_f32 = out_f32; ep.ld_f32 = ld_f32;
    ep.out_bf16 = reinterpret_cast<__nv_bfloat16 *>(out_bf16); ep.ld_bf16 = ld_bf16;
    return gemm_bf16_tc(a, b, ep, M, N, K, kEpiStore, (cudaStream_t)stream);
}

extern "C" int dc_gemm_bf16_ex(const uint16_t *A, int64_t lda, int a_mn, const uint16_t *B, int64_t ldb, int b_mn,
                               int M, int N, int K, const float *bias, const float *addend, int64_t ld_addend,
                               int addend_mod, int relu, const uint16_t *mask_src, int64_t ld_mask,
                               int deint_units, int atomic, int split_k, float *out_f32, int64_t ld_f32,
                               uint16_t *out_bf16, int64_t ld_bf16, void *stream) {
    TcOperand a, b;
    a.ptr = reinterpret_cast<const __nv_bfloat16 *>(A); a.ld = lda; a.mn_major = a_mn != 0;
    b.ptr = reinterpret_cast<const __nv_bfloat16 *>(B); b.ld = ldb; b.mn_major = b_mn != 0;
    TcEpilogue ep;
    ep.bias = bias; ep.addend = addend; ep.ld_addend = ld_addend; ep.addend_mod = addend_mod; ep.relu = relu;
    ep.mask_src = reinterpret_cast<const __nv_bfloat16 *>(mask_src); ep.ld_mask = ld_mask;
    ep.deint_units = deint_units; ep.atomic = atomic;
    ep.out_f32 = out_f32; ep.ld_f32 = ld_f32;
    ep.out_bf16 = reinterpret_cast<__nv_bfloat16 *>(out_bf16); ep.ld_bf16 = ld_bf16;
    return gemm_bf16_tc(a, b, ep, M, N, K, kEpiStore, (cudaStream_t)stream, split_k);
}

extern "C" int dc_gemm_bf16_argmax(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M, int N,
                                   int K, const float *bias, int32_t *tokens, float *maxprob, void *stream) {
    if (M <= 0) return DC_OK;
    DC_REQUIRE(tokens && bias, "null pointer argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int tiles = gemm_tc_argmax_tiles(N);
    float *partial = nullptr;
    DC_CHECK_CUDA(cudaMallocAsync((void **)&partial, sizeof(float) * 4 * (size_t)M * tiles, s));
    TcOperand a, b;
    a.ptr = reinterpret_cast<const __nv_bfloat16 *>(A); a.ld = lda;
    b.ptr = reinterpret_cast<const __nv_bfloat16 *>(Bt); b.ld = ldb;
    TcEpilogue ep;
    ep.bias = bias; ep.partial = partial;
    int rc = gemm_bf16_tc(a, b, ep, M, N, K, maxprob ? kEpiArgmaxSum : kEpiArgmax, s);
    if (rc == DC_OK) rc = argmax_merge(partial, M, tiles, tokens, 1, nullptr, maxprob, s);
    cudaFreeAsync(partial, s);
    return rc;
}

extern "C" int dc_gemm_bf16_topk(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M, int N, int K,
                                 const float *bias, int k, int32_t *idx, float *prob, void *stream) {
    if (M <= 0) return DC_OK;
    DC_REQUIRE(idx && prob && bias, "null pointer argument");
    DC_REQUIRE(k >= 1 && k <= kTopKMax && k <= N, "k=%d outside [1,%d]", k, kTopKMax);
    cudaStream_t s = (cudaStream_t)stream;
    const int slots = gemm_tc_argmax_tiles(N);
    float *partial = nullptr;
    DC_CHECK_CUDA(cudaMallocAsync((void **)&partial, sizeof(float) * (2 + 2 * k) * (size_t)M * slots, s));
    TcOperand a, b;
    a.ptr = reinterpret_cast<const __nv_bfloat16 *>(A); a.ld = lda;
    b.ptr = reinterpret_cast<const __nv_bfloat16 *>(Bt); b.ld = ldb;
    TcEpilogue ep;
    ep.bias = bias; ep.partial = partial; ep.topk = k;
    int rc = gemm_bf16_tc(a, b, ep, M, N, K, kEpiTopK, s);
    if (rc == DC_OK) rc = topk_merge(partial, M, slots, k, idx, prob, s);
    cudaFreeAsync(partial, s);
    return rc;
}

extern "C" int dc_gemm_bf16_lstm_cell(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M,
                                      int units, int K, const float *addend, int64_t ld_addend,
                                      const float *bias, const int32_t *tok, float *c,
                                      const uint16_t *h_prev, int64_t ld_h_prev, uint16_t *h_out,
                                      int64_t ld_h_out, uint16_t *h_out2, int64_t ld_h_out2, void *stream) {
    TcOperand a, b;
    a.ptr = reinterpret_cast<const __nv_bfloat16 *>(A); a.ld = lda;
    b.ptr = reinterpret_cast<const __nv_bfloat16 *>(Bt); b.ld = ldb;
    TcEpilogue ep;
    ep.addend = addend; ep.ld_addend = ld_addend; ep.bias = bias;
    ep.cell_c = c; ep.cell_units = units; ep.cell_tok = tok;
    ep.cell_h_prev = reinterpret_cast<const __nv_bfloat16 *>(h_prev); ep.ld_h_prev = ld_h_prev;
    ep.cell_h_a = reinterpret_cast<__nv_bfloat16 *>(h_out); ep.ld_h_a = ld_h_out;
    ep.cell_h_b = reinterpret_cast<__nv_bfloat16 *>(h_out2); ep.ld_h_b = ld_h_out2;
    return gemm_bf16_tc(a, b, ep, M, 4 * units, K, kEpiCell, (cudaStream_t)stream);
}

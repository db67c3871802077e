// bf16 / tcgen05 decoder path (placeholder until gemm_tc.cu lands).
#include "decoder.cuh"

namespace dcap {
struct Bf16State {};
static int unsupported() { return set_error(DC_ERR_UNSUPPORTED, "bf16 decoder path is not built yet"); }
int Decoder::finalize_bf16(cudaStream_t) { return unsupported(); }
int Decoder::reserve_bf16(size_t) { return unsupported(); }
int Decoder::head_bf16(const void *, int, int, float *, cudaStream_t) { return unsupported(); }
int Decoder::v1_hoist_bf16(int, cudaStream_t) { return unsupported(); }
int Decoder::reset_state_bf16(int, cudaStream_t) { return unsupported(); }
int Decoder::v1_step_bf16(int, const float *, const float *, cudaStream_t) { return unsupported(); }
int Decoder::greedy_bf16(const void *, int, int, int32_t *, cudaStream_t) { return unsupported(); }
int Decoder::beam_gather_bf16(int, int, cudaStream_t) { return unsupported(); }
}  // namespace dcap

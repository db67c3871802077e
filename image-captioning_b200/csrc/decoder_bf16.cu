// bf16 decoder path: every dense contraction runs on the tcgen05 GEMM (gemm_tc.cu) with bf16
// operands and fp32 accumulation; LSTM cell state c, the hoisted per-RoI terms and all reductions
// stay fp32.  Per decoding step (v1 word model, text_generation_model.py:130-156):
//
//   embed gather -> [gates1 GEMM + fused cell] -> [gates2 GEMM + fused cell]
//                -> [dense1 GEMM + hoisted term + ReLU] -> [vocab GEMM + fused arg-max] -> merge
//
// Weights are re-laid once (finalize) as K-major bf16 [N, K]; LSTM kernels additionally with
// gate-interleaved rows (row 4u+g) so that one epilogue thread sees the four gates of a unit.
// h lives only as bf16 inside the next GEMM's operand buffers ([emb|h1] and [h1|h2], ping-pong
// per step because the GEMM that consumes h_{t-1} also produces h_t).
#include "decoder.cuh"
#include "gemm_tc.cuh"
#include "decoder_bf16.cuh"

#include <stdlib.h>

namespace dcap {

void Decoder::free_bf16() {
    delete bf;
    bf = nullptr;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// dst[n, k_off + k] = bf16(src[(row_off + k) * n_src + colmap(n)]), colmap = gate interleave (n = 4u+g ->
// g*units + u) or identity: the Keras [in, out] kernel re-laid as the K-major [N, K] TMA operand.
// 32 x 32 tiles through shared memory: reads run along n (whole sectors), writes along k.
__global__ void __launch_bounds__(256) build_kmajor_kernel(const float *__restrict__ src, int n_src, int row_off, int K, int N,
                                                           int interleave_units, __nv_bfloat16 *__restrict__ dst,
                                                           long long ld_dst, int k_off) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    const int n = n0 + threadIdx.x;
    const int col = interleave_units ? (n & 3) * interleave_units + (n >> 2) : n;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int k = k0 + i;
        tile[i][threadIdx.x] = (k < K && n < N) ? src[(long long)(row_off + k) * n_src + col] : 0.f;
    }
    __syncthreads();
    const int k = k0 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int nn = n0 + i;
        if (k < K && nn < N) dst[(long long)nn * ld_dst + k_off + k] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
}

// dst[r, c] = bf16(src[r, c]) for c < cols; dst rows are ld_dst wide (padding pre-zeroed)
__global__ void pad_rows_bf16_kernel(const float *__restrict__ src, int rows, int cols,
                                     __nv_bfloat16 *__restrict__ dst, int ld_dst) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * cols) return;
    const int r = (int)(idx / cols), c = (int)(idx - (long long)r * cols);
    dst[(long long)r * ld_dst + c] = __float2bfloat16_rn(src[idx]);
}

__global__ void interleave_bias_kernel(const float *__restrict__ src, int units, float *__restrict__ dst) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 4 * units) dst[n] = src[(n & 3) * units + (n >> 2)];
}

// The same re-layout for up to kMaxKmajorJobs tensors in ONE launch (the optimiser re-derives every operand copy after
// each step: ten launches of 3 - 20 us each were a chain of launch gaps, 0.12 ms of a 2.6 ms step at 8 ranks).
// Blocks are numbered job after job; a block finds its job by walking the (short) table of first-block indices.
constexpr int kMaxKmajorJobs = 12;
struct KmajorJob {
    const float *src; __nv_bfloat16 *dst; long long ld_dst;
    int n_src, row_off, K, N, interleave_units, k_off, tiles_k, first_block;
    int pair_ok;                 // destination rows start on 4-byte boundaries: bf16 pairs can be stored as one word
};
struct KmajorBatch {
    KmajorJob job[kMaxKmajorJobs];
    int n = 0, blocks = 0;
    void add(const float *src, int n_src, int row_off, int K, int N, int interleave_units, __nv_bfloat16 *dst, long long ld_dst,
             int k_off) {
        KmajorJob &j = job[n++];
        j.src = src; j.dst = dst; j.ld_dst = ld_dst; j.n_src = n_src; j.row_off = row_off; j.K = K; j.N = N;
        j.interleave_units = interleave_units; j.k_off = k_off; j.tiles_k = ceil_div(K, 64); j.first_block = blocks;
        j.pair_ok = (ld_dst % 2 == 0 && k_off % 2 == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) ? 1 : 0;
        blocks += j.tiles_k * ceil_div(N, 32);
    }
};

// 64 (k) x 32 (n) tiles: reads run along n (whole 32-byte sectors, also through the gate interleave), writes along k as
// bf16 pairs (128 bytes per warp instruction; the 32 x 32 tiles of build_kmajor_kernel write 64)
__global__ void __launch_bounds__(256) build_kmajor_batch_kernel(const __grid_constant__ KmajorBatch batch) {
    __shared__ float tile[64][33];
    int ji = 0;
    while (ji + 1 < batch.n && (int)blockIdx.x >= batch.job[ji + 1].first_block) ++ji;
    const KmajorJob &j = batch.job[ji];
    const int local = blockIdx.x - j.first_block;
    const int k0 = (local % j.tiles_k) * 64, n0 = (local / j.tiles_k) * 32;
    const int n = n0 + threadIdx.x;
    const int col = j.interleave_units ? (n & 3) * j.interleave_units + (n >> 2) : n;
#pragma unroll
    for (int i = threadIdx.y; i < 64; i += 8) {
        const int k = k0 + i;
        tile[i][threadIdx.x] = (k < j.K && n < j.N) ? j.src[(long long)(j.row_off + k) * j.n_src + col] : 0.f;
    }
    __syncthreads();
    const int k = k0 + 2 * threadIdx.x;
#pragma unroll
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int nn = n0 + i;
        if (nn >= j.N || k >= j.K) continue;
        __nv_bfloat16 *d = j.dst + (long long)nn * j.ld_dst + j.k_off + k;
        if (k + 1 < j.K && j.pair_ok)
            *reinterpret_cast<__nv_bfloat162 *>(d) = __floats2bfloat162_rn(tile[2 * threadIdx.x][i], tile[2 * threadIdx.x + 1][i]);
        else {
            d[0] = __float2bfloat16_rn(tile[2 * threadIdx.x][i]);
            if (k + 1 < j.K) d[1] = __float2bfloat16_rn(tile[2 * threadIdx.x + 1][i]);
        }
    }
}

static int launch_kmajor_batch(const KmajorBatch &batch, cudaStream_t s) {
    if (batch.blocks == 0) return DC_OK;
    build_kmajor_batch_kernel<<<batch.blocks, dim3(32, 8), 0, s>>>(batch);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// gate-interleaved LSTM biases and the dense1 bias of the merged hoist GEMM, one launch
__global__ void derive_biases_kernel(const float *__restrict__ b1, const float *__restrict__ b2, const float *__restrict__ bd1,
                                     int units, int dense, float *__restrict__ b1_i, float *__restrict__ b2_i,
                                     float *__restrict__ bd1_dst) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 4 * units) {
        const int src = (n & 3) * units + (n >> 2);
        b1_i[n] = b1[src];
        b2_i[n] = b2[src];
    }
    if (n < dense) bd1_dst[n] = bd1[n];
}

int build_kmajor(const float *src, int n_src, int row_off, int K, int N, int interleave_units,
                        __nv_bfloat16 *dst, long long ld_dst, int k_off, cudaStream_t s) {
    const dim3 grid(ceil_div(K, 32), ceil_div(N, 32)), block(32, 8);
    build_kmajor_kernel<<<grid, block, 0, s>>>(src, n_src, row_off, K, N, interleave_units, dst, ld_dst, k_off);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

int Decoder::refresh_bf16(bool fresh, cudaStream_t s) {
    if (!bf) bf = new Bf16State();
    Bf16State &b = *bf;
    const int F = cfg.feat, E = cfg.embed, U = cfg.units, V = cfg.vocab;
    const int Kin = cfg.pool * cfg.pool * cfg.channels;
    b.Epad = round_up(E, 64);
    auto A16 = [&](__nv_bfloat16 **p, size_t n) { return dev_alloc((void **)p, 2 * n, owned); };
    int rc = 0;
    if (cfg.arch == DC_ARCH_V2_INJECT) {
        // word LSTM(Wu) on [emb | h], image LSTM(U) on [head | word vector] (one step from the zero state, so its
        // recurrent kernel never contributes), Dense(V)
        const int Wu = cfg.word_units, Kw = b.Epad + Wu;
        if (fresh) {
            rc |= A16(&b.w_head1, (size_t)F * Kin); rc |= A16(&b.w_head2, (size_t)F * F);
            rc |= A16(&b.v2_w1cat, (size_t)4 * Wu * Kw); rc |= A16(&b.v2_wimg, (size_t)4 * U * (F + Wu));
            rc |= A16(&b.v2_wd, (size_t)V * U); rc |= A16(&b.emb, (size_t)V * b.Epad);
            rc |= dev_alloc((void **)&b.v2_bw, sizeof(float) * 4 * Wu, owned);
            rc |= dev_alloc((void **)&b.v2_bimg, sizeof(float) * 4 * U, owned);
            if (rc) return rc;
            DC_CHECK_CUDA(cudaMemsetAsync(b.v2_w1cat, 0, 2 * (size_t)4 * Wu * Kw, s));
        }
        if (emb_dirty) {
            emb_dirty = false;
            DC_CHECK_CUDA(cudaMemsetAsync(b.emb, 0, 2 * (size_t)V * b.Epad, s));
            pad_rows_bf16_kernel<<<(unsigned)ceil_div<long long>((long long)V * E, 256), 256, 0, s>>>(
                W("imgcap_embedding_layer/embeddings"), V, E, b.emb, b.Epad);
            DC_CHECK_LAUNCH();
        }
        rc |= build_kmajor(W("mrcnn_class_conv1/kernel"), F, 0, Kin, F, 0, b.w_head1, Kin, 0, s);
        rc |= build_kmajor(W("mrcnn_class_conv2/kernel"), F, 0, F, F, 0, b.w_head2, F, 0, s);
        rc |= build_kmajor(W("lstm_1/kernel"), 4 * Wu, 0, E, 4 * Wu, Wu, b.v2_w1cat, Kw, 0, s);
        rc |= build_kmajor(W("lstm_1/recurrent_kernel"), 4 * Wu, 0, Wu, 4 * Wu, Wu, b.v2_w1cat, Kw, b.Epad, s);
        rc |= build_kmajor(W("imgcap_lstm/kernel"), 4 * U, 0, F + Wu, 4 * U, U, b.v2_wimg, F + Wu, 0, s);
        rc |= build_kmajor(W("imgcap_d1/kernel"), V, 0, U, V, 0, b.v2_wd, U, 0, s);
        if (rc) return rc;
        interleave_bias_kernel<<<ceil_div(4 * Wu, 256), 256, 0, s>>>(W("lstm_1/bias"), Wu, b.v2_bw);
        interleave_bias_kernel<<<ceil_div(4 * U, 256), 256, 0, s>>>(W("imgcap_lstm/bias"), U, b.v2_bimg);
        DC_CHECK_LAUNCH();
        return DC_OK;
    }
    const int K1 = b.Epad + U;
    if (fresh) {
        rc |= A16(&b.w_head1, (size_t)F * Kin); rc |= A16(&b.w_head2, (size_t)F * F);
        rc |= A16(&b.w1cat, (size_t)4 * U * K1);
        rc |= A16(&b.w1f, ((size_t)4 * U + kDense) * F);      // [W1f ; Wd1f] adjacent: one merged hoist GEMM (v1_hoist_merged_bf16)
        b.wd1f = rc ? nullptr : b.w1f + (size_t)4 * U * F;
        rc |= A16(&b.w2cat, (size_t)4 * U * 2 * U);
        rc |= A16(&b.wd1h, (size_t)kDense * U);
        rc |= A16(&b.wd2, (size_t)V * kDense);
        rc |= A16(&b.emb, (size_t)V * b.Epad);
        rc |= dev_alloc((void **)&b.bias_hoist, sizeof(float) * (4 * U + kDense), owned);
        b.b1_i = b.bias_hoist;                                  // [b1 (gate-interleaved) | bd1]
        rc |= dev_alloc((void **)&b.b2_i, sizeof(float) * 4 * U, owned);
        if (rc) return rc;
        DC_CHECK_CUDA(cudaMemsetAsync(b.w1cat, 0, 2 * (size_t)4 * U * K1, s));      // zero the E..Epad padding
    }
    if (emb_dirty) {
        // the embedding table is frozen (trainable=False): rebuilt only when it is set
        emb_dirty = false;
        DC_CHECK_CUDA(cudaMemsetAsync(b.emb, 0, 2 * (size_t)V * b.Epad, s));
        pad_rows_bf16_kernel<<<(unsigned)ceil_div<long long>((long long)V * E, 256), 256, 0, s>>>(
            W("imgcap_embedding_layer/embeddings"), V, E, b.emb, b.Epad);
        DC_CHECK_LAUNCH();
    }
    KmajorBatch kb;
    kb.add(W("mrcnn_class_conv1/kernel"), F, 0, Kin, F, 0, b.w_head1, Kin, 0);
    kb.add(W("mrcnn_class_conv2/kernel"), F, 0, F, F, 0, b.w_head2, F, 0);
    kb.add(W("imgcap_lstm1/kernel"), 4 * U, 0, E, 4 * U, U, b.w1cat, K1, 0);
    kb.add(W("imgcap_lstm1/recurrent_kernel"), 4 * U, 0, U, 4 * U, U, b.w1cat, K1, b.Epad);
    kb.add(W("imgcap_lstm1/kernel"), 4 * U, E, F, 4 * U, U, b.w1f, F, 0);
    kb.add(W("imgcap_lstm2/kernel"), 4 * U, 0, U, 4 * U, U, b.w2cat, 2 * U, 0);
    kb.add(W("imgcap_lstm2/recurrent_kernel"), 4 * U, 0, U, 4 * U, U, b.w2cat, 2 * U, U);
    kb.add(W("imgcap_lstm_d1/kernel"), kDense, 0, U, kDense, 0, b.wd1h, U, 0);
    kb.add(W("imgcap_lstm_d1/kernel"), kDense, U, F, kDense, 0, b.wd1f, F, 0);
    kb.add(W("imgcap_lstm_d2/kernel"), V, 0, kDense, V, 0, b.wd2, kDense, 0);
    if (int rc2 = launch_kmajor_batch(kb, s)) return rc2;
    const int nb = 4 * U > kDense ? 4 * U : kDense;
    derive_biases_kernel<<<ceil_div(nb, 256), 256, 0, s>>>(W("imgcap_lstm1/bias"), W("imgcap_lstm2/bias"), W("imgcap_lstm_d1/bias"), U,
                                                           kDense, b.b1_i, b.b2_i, b.bias_hoist + 4 * U);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

int Decoder::reserve_bf16(size_t R) {
    Bf16State &b = *bf;
    b.topk_partial = nullptr; b.topk_cap = 0;            // lived in the workspace that was just released
    const size_t F = cfg.feat, U = cfg.units;
    const size_t Kin = (size_t)cfg.pool * cfg.pool * cfg.channels;
    const size_t K1 = b.Epad + U;
    auto A16 = [&](__nv_bfloat16 **p, size_t n) { return dev_alloc((void **)p, 2 * n, ws_owned); };
    int rc = 0;
    if (cfg.arch == DC_ARCH_V2_INJECT) {
        const size_t Wu = cfg.word_units;
        rc |= A16(&b.roi, R * Kin); rc |= A16(&b.a1, R * F); rc |= A16(&b.Fb, R * F);
        for (int i = 0; i < 2; ++i) rc |= A16(&b.X1[i], R * (b.Epad + Wu));
        rc |= A16(&b.v2_xin, R * (F + Wu)); rc |= A16(&b.v2_hb, R * U);
        rc |= dev_alloc((void **)&b.v2_czero, sizeof(float) * R * U, ws_owned);
        rc |= dev_alloc((void **)&b.partial, sizeof(float) * 4 * R * gemm_tc_argmax_tiles(cfg.vocab), ws_owned);
        if (rc) return rc;
        DC_CHECK_CUDA(cudaMemset(b.v2_czero, 0, sizeof(float) * R * U));
        return DC_OK;
    }
    rc |= A16(&b.roi, R * Kin); rc |= A16(&b.a1, R * F); rc |= A16(&b.Fb, R * F); rc |= A16(&b.d, R * kDense);
    for (int i = 0; i < 2; ++i) { rc |= A16(&b.X1[i], R * K1); rc |= A16(&b.X2[i], R * 2 * U); }
    rc |= dev_alloc((void **)&b.partial, sizeof(float) * 4 * R * gemm_tc_argmax_tiles(cfg.vocab), ws_owned);
    rc |= dev_alloc((void **)&b.hoist_all, sizeof(float) * R * (4 * U + kDense), ws_owned);
    b.loop_cnt_bytes = sizeof(unsigned) * (16 * ((R + 255) / 256) + 4);     // greedy_loop.cu: 8 counters per 128 rows + error word
    rc |= dev_alloc((void **)&b.loop_cnt, b.loop_cnt_bytes, ws_owned);
    return rc;
}

static TcOperand op(const __nv_bfloat16 *p, long long ld) {
    TcOperand o;
    o.ptr = p; o.ld = ld;
    return o;
}

int Decoder::head_bf16(const void *feats, int kind, int B, float *out, cudaStream_t s) {
    Bf16State &b = *bf;
    const int F = cfg.feat, Kin = cfg.pool * cfg.pool * cfg.channels;
    const __nv_bfloat16 *x = nullptr;
    if (kind == DC_FEATS_ROI_BF16) {
        x = reinterpret_cast<const __nv_bfloat16 *>(feats);
    } else {
        DC_REQUIRE(kind == DC_FEATS_ROI_F32, "unknown feats_kind %d", kind);
        if (int rc = f32_to_bf16((const float *)feats, b.roi, (long long)B * Kin, s)) return rc;
        x = b.roi;
    }
    TcEpilogue e1;
    e1.bias = W("mrcnn_class_conv1/bias"); e1.scale = bn_scale[0]; e1.shift = bn_shift[0]; e1.relu = 1;
    e1.out_bf16 = b.a1; e1.ld_bf16 = F;
    if (int rc = gemm_bf16_tc(op(x, Kin), op(b.w_head1, Kin), e1, B, F, Kin, kEpiStore, s)) return rc;
    TcEpilogue e2;
    e2.bias = W("mrcnn_class_conv2/bias"); e2.scale = bn_scale[1]; e2.shift = bn_shift[1]; e2.relu = 1;
    e2.out_bf16 = b.Fb; e2.ld_bf16 = F;
    e2.out_f32 = out; e2.ld_f32 = F;                               // out == nullptr: only the bf16 copy (one bulk tensor store)
    return gemm_bf16_tc(op(b.a1, F), op(b.w_head2, F), e2, B, F, F, kEpiStore, s);
}

int Decoder::v1_hoist_bf16(int B, cudaStream_t s, bool blocked32) {
    Bf16State &b = *bf;
    const int F = cfg.feat, U = cfg.units;
    // ws.F (fp32) may have come from the caller (DC_FEATS_HEAD_F32): refresh the bf16 operand copy
    if (int rc = f32_to_bf16(ws.F, b.Fb, (long long)B * F, s)) return rc;
    TcEpilogue e1;
    e1.bias = b.b1_i; e1.out_f32 = ws.g1f; e1.ld_f32 = 4 * U;                  // gate-interleaved columns
    e1.blocked32 = blocked32 ? 1 : 0;
    if (int rc = gemm_bf16_tc(op(b.Fb, F), op(b.w1f, F), e1, B, 4 * U, F, kEpiStore, s)) return rc;
    TcEpilogue e2;
    e2.bias = W("imgcap_lstm_d1/bias"); e2.out_f32 = ws.d1f; e2.ld_f32 = kDense;
    e2.blocked32 = blocked32 ? 1 : 0;
    return gemm_bf16_tc(op(b.Fb, F), op(b.wd1f, F), e2, B, kDense, F, kEpiStore, s);
}

// [f W1f + b1 | f Wd1f + bd1] in ONE GEMM (N = 4U + 1024; the two weight blocks are adjacent), written in the blocked-32
// layout of greedy_loop.cu.  fresh_fb: bf->Fb already holds bf16(ws.F) (head_bf16 ran in this call).
int Decoder::v1_hoist_merged_bf16(int B, bool fresh_fb, cudaStream_t s) {
    Bf16State &b = *bf;
    const int F = cfg.feat, U = cfg.units, N = 4 * U + kDense;
    if (!fresh_fb)
        if (int rc = f32_to_bf16(ws.F, b.Fb, (long long)B * F, s)) return rc;
    TcEpilogue e;
    e.bias = b.bias_hoist; e.out_f32 = b.hoist_all; e.ld_f32 = N; e.blocked32 = 1;
    return gemm_bf16_tc(op(b.Fb, F), op(b.w1f, F), e, B, N, F, kEpiStore, s);
}

int Decoder::reset_state_bf16(int R, cudaStream_t s) {
    Bf16State &b = *bf;
    const size_t U = cfg.units, K1 = b.Epad + U;
    for (int i = 0; i < 2; ++i) {
        DC_CHECK_CUDA(cudaMemsetAsync(b.X1[i], 0, 2 * (size_t)R * K1, s));
        DC_CHECK_CUDA(cudaMemsetAsync(b.X2[i], 0, 2 * (size_t)R * 2 * U, s));
    }
    b.parity = 0;
    return DC_OK;
}

// the three GEMMs up to the Dense(1024) activations; leaves d (bf16) ready for the vocab GEMM
// addend_div > 0: rows are (RoI, beam) pairs and the per-RoI terms g1f / d1f are indexed by row / addend_div
// blocked_ld > 0: g1f / d1f are column blocks of one blocked-32 array of that row length (v1_hoist_merged_bf16)
static int step_core(Decoder &D, int R, const float *g1f, const float *d1f, bool gather, cudaStream_t s, int addend_div = 0,
                     int blocked_ld = 0) {
    Bf16State &b = *D.bf;
    const DcDecoderConfig &cfg = D.cfg;
    const int E = cfg.embed, U = cfg.units, V = cfg.vocab, K1 = b.Epad + U, p = b.parity;
    // gather == false: the embedding rows were already placed in X1[p] by the previous step's merge kernel
    if (gather)
        if (int rc = embed_gather(D.W("imgcap_embedding_layer/embeddings"), D.ws.tok, R, E, V, b.X1[p], K1, true, s)) return rc;
    TcEpilogue c1;
    c1.addend = g1f; c1.ld_addend = blocked_ld > 0 ? blocked_ld : 4 * U; c1.addend_blocked32 = blocked_ld > 0; c1.addend_div = addend_div; c1.cell_c = D.ws.c1; c1.cell_units = U; c1.cell_tok = D.ws.tok;
    c1.cell_h_prev = b.X1[p] + b.Epad; c1.ld_h_prev = K1;
    c1.cell_h_a = b.X1[p ^ 1] + b.Epad; c1.ld_h_a = K1;
    c1.cell_h_b = b.X2[p]; c1.ld_h_b = 2 * U;
    if (int rc = gemm_bf16_tc(op(b.X1[p], K1), op(b.w1cat, K1), c1, R, 4 * U, K1, kEpiCell, s)) return rc;
    TcEpilogue c2;
    c2.bias = b.b2_i; c2.cell_c = D.ws.c2; c2.cell_units = U; c2.cell_tok = D.ws.tok;
    c2.cell_h_prev = b.X2[p] + U; c2.ld_h_prev = 2 * U;
    c2.cell_h_a = b.X2[p ^ 1] + U; c2.ld_h_a = 2 * U;
    if (int rc = gemm_bf16_tc(op(b.X2[p], 2 * U), op(b.w2cat, 2 * U), c2, R, 4 * U, 2 * U, kEpiCell, s)) return rc;
    TcEpilogue d1;
    d1.addend = d1f; d1.ld_addend = blocked_ld > 0 ? blocked_ld : kDense; d1.addend_blocked32 = blocked_ld > 0; d1.addend_div = addend_div; d1.relu = 1; d1.out_bf16 = b.d; d1.ld_bf16 = kDense;
    if (int rc = gemm_bf16_tc(op(b.X2[p ^ 1] + U, 2 * U), op(b.wd1h, U), d1, R, kDense, U, kEpiStore, s)) return rc;
    b.parity ^= 1;
    return DC_OK;
}

// generic step (predict surface / beam): logits are materialised in ws.logits (fp32)
int Decoder::v1_step_bf16(int R, const float *g1f, const float *d1f, cudaStream_t s) {
    if (int rc = step_core(*this, R, g1f, d1f, true, s)) return rc;
    TcEpilogue e;
    e.bias = W("imgcap_lstm_d2/bias"); e.out_f32 = ws.logits; e.ld_f32 = cfg.vocab;
    return gemm_bf16_tc(op(bf->d, kDense), op(bf->wd2, kDense), e, R, cfg.vocab, kDense, kEpiStore, s);
}

// greedy fast path: tokens only, the [B,V] logits never exist
// (Round 2 tried decoding the batch as two lanes -- half-batches on two streams, every GEMM limited to half of the
// SMs so that the two chains run side by side -- to hide launch gaps and tile-wave tails: 4.09 ms against 4.04 ms per
// 8000 RoIs, no gain; inside the CUDA graph the kernels already run back to back, and their inefficiency is in the
// main loop / epilogue, not between launches.  Removed.)
int Decoder::greedy_bf16(const void *feats, int kind, int B, int32_t *tokens, cudaStream_t s, float *scores) {
    const int P = cfg.padding, V = cfg.vocab;
    const bool loop = greedy_loop_ok(B);
    // the loop kernel's hoist GEMM reads the bf16 head output only: the fp32 copy is not produced at all then
    const bool head_bf16_only = loop && !greedy_loop_folds() && hoist_merged() && kind != DC_FEATS_HEAD_F32;
    if (head_bf16_only) { if (int rc = head_bf16(feats, kind, B, nullptr, s)) return rc; }
    else if (int rc = head(feats, kind, B, ws.F, s)) return rc;
    if (loop && greedy_loop_folds()) {
        // the loop kernel contracts over the head features itself: only their bf16 copy is needed (ws.F may have come
        // from the caller: DC_FEATS_HEAD_F32), not the hoisted fp32 terms
        if (int rc = f32_to_bf16(ws.F, bf->Fb, (long long)B * cfg.feat, s)) return rc;
    } else if (loop) {
        // hoisted per-RoI terms, one GEMM, in the blocked-32 layout the loop kernel's epilogues read with coalesced accesses
        if (hoist_merged()) { if (int rc = v1_hoist_merged_bf16(B, head_bf16_only, s)) return rc; }
        else if (int rc = v1_hoist_bf16(B, s, true)) return rc;
    } else if (int rc = v1_hoist(B, s)) return rc;
    if (int rc = v1_reset_state(B, s)) return rc;
    if (int rc = fill_i32(ws.tok, B, 1, s)) return rc;
    // the whole loop as one persistent kernel (greedy_loop.cu); DCAP_GREEDY_LOOP=0 keeps the launch-per-GEMM form below
    if (loop) return greedy_loop_bf16(B, tokens, scores, s);
    const int slots = gemm_tc_argmax_tiles(V);
    for (int t = 0; t < P; ++t) {
        if (int rc = step_core(*this, B, ws.g1f, ws.d1f, t == 0, s)) return rc;
        TcEpilogue e;
        e.bias = W("imgcap_lstm_d2/bias"); e.partial = bf->partial;
        // caption scores need the softmax probability of the arg-max: the epilogue also sums exp(v - max)
        if (int rc = gemm_bf16_tc(op(bf->d, kDense), op(bf->wd2, kDense), e, B, V, kDense, scores ? kEpiArgmaxSum : kEpiArgmax, s)) return rc;
        // token of this step + its embedding row, written into the operand buffer of step t+1
        const bool more = t + 1 < P;
        if (int rc = argmax_merge(bf->partial, B, slots, tokens + t, P, ws.tok, scores ? ws.cand_p : nullptr, s,
                                  more ? bf->emb : nullptr, bf->Epad, more ? bf->X1[bf->parity] : nullptr,
                                  bf->Epad + cfg.units)) return rc;
        if (scores)
            if (int rc = accumulate_log(scores, ws.cand_p, B, t == 0, s)) return rc;
    }
    return DC_OK;
}

void Decoder::drop_graphs() {
    for (auto &kv : graphs) cudaGraphExecDestroy(kv.second);
    graphs.clear();
    graph_calls.clear();
}

// Replays (or, on the second call with the same arguments, captures) the whole greedy sequence as
// one CUDA graph.  Capture runs on a private non-blocking stream (the caller's stream may be the
// legacy default stream, which cannot be captured); events order it after / before the caller's
// stream.
int Decoder::greedy_bf16_graphed(const void *feats, int kind, int B, int32_t *tokens, cudaStream_t s) {
    static const bool env_off = getenv("DCAP_NO_GRAPHS") != nullptr;
    if (!use_graphs || env_off) return greedy_bf16(feats, kind, B, tokens, s);
    // the graph writes the ids into the handle's own buffer (hist_a) so that the key does not depend
    // on the caller's output pointer; a small device-to-device copy delivers them
    int32_t *internal = ws.hist_a;
    const GraphKey key(feats, kind, B, (void *)internal);
    auto it = graphs.find(key);
    if (it == graphs.end()) {
        if (graph_calls.size() >= 64) graph_calls.clear();          // callers that pass a fresh feats pointer every call never capture: bound the map
        if (++graph_calls[key] < 2) return greedy_bf16(feats, kind, B, tokens, s);   // first call: eager
        if (!graph_stream) {
            DC_CHECK_CUDA(cudaStreamCreateWithFlags(&graph_stream, cudaStreamNonBlocking));
            DC_CHECK_CUDA(cudaEventCreateWithFlags(&graph_ev_in, cudaEventDisableTiming));
            DC_CHECK_CUDA(cudaEventCreateWithFlags(&graph_ev_out, cudaEventDisableTiming));
        }
        if (graphs.size() >= 16) drop_graphs();
        cudaGraph_t g = nullptr;
        DC_CHECK_CUDA(cudaStreamBeginCapture(graph_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = greedy_bf16(feats, kind, B, internal, graph_stream);
        const cudaError_t e = cudaStreamEndCapture(graph_stream, &g);
        if (rc != DC_OK || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            use_graphs = false;                                     // fall back to eager launches for good
            return greedy_bf16(feats, kind, B, tokens, s);
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&exec, g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) {
            cudaGetLastError();
            use_graphs = false;
            return greedy_bf16(feats, kind, B, tokens, s);
        }
        it = graphs.emplace(key, exec).first;
    }
    DC_CHECK_CUDA(cudaEventRecord(graph_ev_in, s));
    DC_CHECK_CUDA(cudaStreamWaitEvent(graph_stream, graph_ev_in, 0));
    DC_CHECK_CUDA(cudaGraphLaunch(it->second, graph_stream));
    DC_CHECK_CUDA(cudaMemcpyAsync(tokens, internal, sizeof(int32_t) * (size_t)B * cfg.padding,
                                  cudaMemcpyDeviceToDevice, graph_stream));
    DC_CHECK_CUDA(cudaEventRecord(graph_ev_out, graph_stream));
    DC_CHECK_CUDA(cudaStreamWaitEvent(s, graph_ev_out, 0));
    return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// beam search (gen_captions, image captioning/test.py:23-64) on the tensor-core path
// ------------------------------------------------------------------------------------------------
// Children inherit the parent's post-step state: one CTA per beam row copies h1, h2 (bf16, inside the
// GEMM operand buffers), c1, c2 (fp32) and the token history from row (b, parent[b, j]), then appends
// the new token.  dst buffers are the "other parity" operand buffers / spare state buffers.
struct BeamPermute {
    const int32_t *parent, *new_tok;
    const __nv_bfloat16 *h1_src, *h2_src; __nv_bfloat16 *h1_dst, *h2_dst;
    long long ld1, ld2;
    const float *c1_src, *c2_src; float *c1_dst, *c2_dst;
    const int32_t *hist_src; int32_t *hist_dst, *tok;
    int k, U, P, col;
    int compact;          // first step: the state rows are per RoI (row b), not per beam (row b*k + parent)
    const uint4 *emb; int e8;     // bf16 embedding table [V, 8 * e8]: the new token's row goes into the next step's [emb | h1] operand
    uint4 *x1_dst;                // row base of that operand (ld1 elements per row); null: the next step gathers itself
};

__global__ void __launch_bounds__(128) beam_permute_kernel(const BeamPermute a, int rows) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    const long long sr = a.compact ? (long long)(r / a.k) : (long long)(r / a.k) * a.k + a.parent[r];
    const long long hr = a.compact ? (long long)r : sr;               // history rows always exist per beam
    const int u8 = a.U >> 3, u4 = a.U >> 2;
    const uint4 *h1s = reinterpret_cast<const uint4 *>(a.h1_src + sr * a.ld1), *h2s = reinterpret_cast<const uint4 *>(a.h2_src + sr * a.ld2);
    uint4 *h1d = reinterpret_cast<uint4 *>(a.h1_dst + (long long)r * a.ld1), *h2d = reinterpret_cast<uint4 *>(a.h2_dst + (long long)r * a.ld2);
    const float4 *c1s = reinterpret_cast<const float4 *>(a.c1_src + sr * a.U), *c2s = reinterpret_cast<const float4 *>(a.c2_src + sr * a.U);
    float4 *c1d = reinterpret_cast<float4 *>(a.c1_dst + (long long)r * a.U), *c2d = reinterpret_cast<float4 *>(a.c2_dst + (long long)r * a.U);
    for (int i = threadIdx.x; i < u8; i += blockDim.x) { h1d[i] = h1s[i]; h2d[i] = h2s[i]; }
    for (int i = threadIdx.x; i < u4; i += blockDim.x) { c1d[i] = c1s[i]; c2d[i] = c2s[i]; }
    const int nt = a.new_tok[r];
    for (int i = threadIdx.x; i < a.P; i += blockDim.x)
        a.hist_dst[(long long)r * a.P + i] = (i == a.col) ? nt : a.hist_src[hr * a.P + i];
    if (threadIdx.x == 0) a.tok[r] = nt;
    if (a.x1_dst)
        for (int i = threadIdx.x; i < a.e8; i += blockDim.x)
            a.x1_dst[(long long)r * (a.ld1 >> 3) + i] = __ldg(a.emb + (long long)nt * a.e8 + i);
}

int Decoder::beam_bf16(const void *feats, int kind, int B, int k, int32_t *tokens, double *scores, cudaStream_t s) {
    const int R = B * k;
    if (int rc = reserve(R)) return rc;
    Bf16State &b = *bf;
    const int P = cfg.padding, V = cfg.vocab, U = cfg.units, K1 = b.Epad + U;
    const int slots = gemm_tc_argmax_tiles(V);
    const size_t part_floats = (size_t)R * slots * (2 + 2 * k);
    if (part_floats > b.topk_cap) {                      // top-k partials: [R, slots, 2 + 2k] floats
        if (int rc = dev_alloc((void **)&b.topk_partial, sizeof(float) * part_floats, ws_owned)) return rc;
        b.topk_cap = part_floats;
    }
    // head + hoisted per-RoI terms on B rows; the step GEMMs read them with addend row = beam row / k
    if (int rc = head(feats, kind, B, ws.F, s)) return rc;
    // per-RoI terms: one GEMM into the blocked-32 layout (the beams of a RoI read row m / k: neighbouring lanes then hit
    // the same 512-byte line instead of 11 different ones); DCAP_BEAM_BLOCKED=0 keeps the row-major pair.  Measured on
    // cfg4 (100 k RoIs, width 3): 147-150 ms either way -- at 100 k rows per chunk the step is bound by the main loops of
    // its GEMMs under the power cap, not by these epilogues; kept because it saves a launch and 1.2 GB of strided reads
    static const bool blocked_env = !(getenv("DCAP_BEAM_BLOCKED") && atoi(getenv("DCAP_BEAM_BLOCKED")) == 0);
    const int blocked_ld = blocked_env ? 4 * U + kDense : 0;
    const float *g1f = ws.g1f, *d1f = ws.d1f;
    if (blocked_ld) {
        if (int rc = v1_hoist_merged_bf16(B, kind != DC_FEATS_HEAD_F32, s)) return rc;
        g1f = b.hoist_all; d1f = b.hoist_all + (size_t)U * 32 * 4;
    } else if (int rc = v1_hoist(B, s)) return rc;
    if (int rc = v1_reset_state(R, s)) return rc;
    if (int rc = fill_i32(ws.tok, R, 1, s)) return rc;                  // <start> = 1
    DC_CHECK_CUDA(cudaMemsetAsync(ws.score_a, 0, sizeof(double) * R, s));
    DC_CHECK_CUDA(cudaMemsetAsync(ws.hist_a, 0, sizeof(int32_t) * (size_t)R * P, s));
    if (int rc = set_token_column(ws.hist_a, R, P, 0, ws.tok, s)) return rc;
    int32_t *hist = ws.hist_a, *hist_n = ws.hist_b;
    double *sc = ws.score_a, *sc_n = ws.score_b;
    for (int t = 0; t + 1 < P; ++t) {
        // step 0: every beam of a RoI is the same (<start>, zero state), so it runs ONCE per RoI (B rows, per-RoI
        // addends read directly) and the permutation below fans the state out to the k beams
        const bool first = t == 0;
        const int rows = first ? B : R;
        // from the second step on the permutation kernel has already placed the new tokens' embedding rows
        if (int rc = step_core(*this, rows, g1f, d1f, first, s, first ? 0 : k, blocked_ld)) return rc;
        TcEpilogue e;
        e.bias = W("imgcap_lstm_d2/bias"); e.partial = b.topk_partial; e.topk = k;
        if (int rc = gemm_bf16_tc(op(b.d, kDense), op(b.wd2, kDense), e, rows, V, kDense, kEpiTopK, s)) return rc;
        if (int rc = topk_merge(b.topk_partial, rows, slots, k, ws.cand_idx, ws.cand_p, s)) return rc;
        if (int rc = beam_select(B, k, first ? 1 : k, ws.cand_idx, ws.cand_p, sc, sc_n, ws.parent, ws.newtok, s, first)) return rc;
        // after step_core the live state sits in X1[parity] (h1) / X2[parity] (h2): permute it into the other
        // parity's buffers and make those current
        const int p = b.parity;
        BeamPermute a;
        a.parent = ws.parent; a.new_tok = ws.newtok;
        a.h1_src = b.X1[p] + b.Epad; a.h1_dst = b.X1[p ^ 1] + b.Epad; a.ld1 = K1;
        a.h2_src = b.X2[p] + U; a.h2_dst = b.X2[p ^ 1] + U; a.ld2 = 2 * U;
        a.c1_src = ws.c1; a.c1_dst = ws.c1b; a.c2_src = ws.c2; a.c2_dst = ws.c2b;
        a.hist_src = hist; a.hist_dst = hist_n; a.tok = ws.tok;
        a.k = k; a.U = U; a.P = P; a.col = t + 1; a.compact = first ? 1 : 0;
        a.emb = reinterpret_cast<const uint4 *>(b.emb); a.e8 = b.Epad / 8;
        a.x1_dst = t + 2 < P ? reinterpret_cast<uint4 *>(b.X1[p ^ 1]) : nullptr;
        beam_permute_kernel<<<R, 128, 0, s>>>(a, R);
        DC_CHECK_LAUNCH();
        b.parity ^= 1;
        std::swap(ws.c1, ws.c1b); std::swap(ws.c2, ws.c2b);
        std::swap(hist, hist_n); std::swap(sc, sc_n);
    }
    DC_CHECK_CUDA(cudaMemcpyAsync(tokens, hist, sizeof(int32_t) * (size_t)R * P, cudaMemcpyDeviceToDevice, s));
    DC_CHECK_CUDA(cudaMemcpyAsync(scores, sc, sizeof(double) * R, cudaMemcpyDeviceToDevice, s));
    return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// v2 inject model (text_generation_model_v2.py:140-166) on the tensor-core path
// ------------------------------------------------------------------------------------------------
int Decoder::v2_begin_bf16(const void *feats, int kind, int B, cudaStream_t s) {
    Bf16State &b = *bf;
    const size_t F = cfg.feat, Wu = cfg.word_units, Kw = b.Epad + Wu;
    if (int rc = head(feats, kind, B, ws.F, s)) return rc;
    // xin = [head feature | word vector]: the word-vector half starts at zero (all-masked prefix)
    DC_CHECK_CUDA(cudaMemsetAsync(b.v2_xin, 0, 2 * (size_t)B * (F + Wu), s));
    pad_rows_bf16_kernel<<<(unsigned)ceil_div<long long>((long long)B * F, 256), 256, 0, s>>>(ws.F, B, (int)F, b.v2_xin, (int)(F + Wu));
    DC_CHECK_LAUNCH();
    for (int i = 0; i < 2; ++i) DC_CHECK_CUDA(cudaMemsetAsync(b.X1[i], 0, 2 * (size_t)B * Kw, s));
    DC_CHECK_CUDA(cudaMemsetAsync(ws.c1, 0, sizeof(float) * (size_t)B * Wu, s));
    b.parity = 0;
    return DC_OK;
}

// consume ws.tok with the word LSTM (masked): state in X1[parity][:, Epad:] / ws.c1; h mirrored into xin[:, F:]
int Decoder::v2_word_step_bf16(int B, bool gather, cudaStream_t s) {
    Bf16State &b = *bf;
    const int E = cfg.embed, V = cfg.vocab, F = cfg.feat, Wu = cfg.word_units, Kw = b.Epad + Wu, p = b.parity;
    if (gather)
        if (int rc = embed_gather(W("imgcap_embedding_layer/embeddings"), ws.tok, B, E, V, b.X1[p], Kw, true, s)) return rc;
    TcEpilogue c;
    c.bias = b.v2_bw; c.cell_c = ws.c1; c.cell_units = Wu; c.cell_tok = ws.tok;
    c.cell_h_prev = b.X1[p] + b.Epad; c.ld_h_prev = Kw;
    c.cell_h_a = b.X1[p ^ 1] + b.Epad; c.ld_h_a = Kw;
    c.cell_h_b = b.v2_xin + F; c.ld_h_b = F + Wu;
    if (int rc = gemm_bf16_tc(op(b.X1[p], Kw), op(b.v2_w1cat, Kw), c, B, 4 * Wu, Kw, kEpiCell, s)) return rc;
    b.parity ^= 1;
    return DC_OK;
}

// [head ; wv] -> LSTM(units), one step from the zero state -> h (bf16) ready for the Dense(V) GEMM
int Decoder::v2_image_step_bf16(int B, cudaStream_t s) {
    Bf16State &b = *bf;
    const int F = cfg.feat, Wu = cfg.word_units, U = cfg.units;
    TcEpilogue c;
    c.bias = b.v2_bimg; c.cell_c = b.v2_czero; c.cell_c_out = ws.c2; c.cell_units = U;
    c.cell_h_a = b.v2_hb; c.ld_h_a = U;
    return gemm_bf16_tc(op(b.v2_xin, F + Wu), op(b.v2_wimg, F + Wu), c, B, 4 * U, F + Wu, kEpiCell, s);
}

int Decoder::v2_predict_bf16(const void *feats, int kind, const int32_t *words, int B, int L, float *probs, cudaStream_t s) {
    Bf16State &b = *bf;
    const int U = cfg.units, V = cfg.vocab;
    if (int rc = v2_begin_bf16(feats, kind, B, s)) return rc;
    for (int t = 0; t < L; ++t) {
        if (int rc = token_column(words, B, L, t, ws.tok, s)) return rc;
        if (int rc = v2_word_step_bf16(B, true, s)) return rc;
    }
    if (int rc = v2_image_step_bf16(B, s)) return rc;
    TcEpilogue e;
    e.bias = W("imgcap_d1/bias"); e.out_f32 = ws.logits; e.ld_f32 = V;
    if (int rc = gemm_bf16_tc(op(b.v2_hb, U), op(b.v2_wd, U), e, B, V, U, kEpiStore, s)) return rc;
    return softmax_argmax(ws.logits, V, B, V, probs, V, nullptr, 0, nullptr, nullptr, s);
}

int Decoder::v2_greedy_bf16(const void *feats, int kind, int B, int32_t *tokens, float *probs, cudaStream_t s,
                            const int32_t *start, float *scores) {
    Bf16State &b = *bf;
    const int P = cfg.padding, U = cfg.units, V = cfg.vocab, Wu = cfg.word_units;
    if (int rc = v2_begin_bf16(feats, kind, B, s)) return rc;
    if (start) DC_CHECK_CUDA(cudaMemcpyAsync(ws.tok, start, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, s));
    else if (int rc = fill_i32(ws.tok, B, 0, s)) return rc;           // argmax(zeros(V)) = 0 -> masked
    const int slots = gemm_tc_argmax_tiles(V);
    for (int t = 0; t + 1 < P; ++t) {
        // the pre-padded window never truncates inside the reference loop (at most P-1 ids), so consuming one
        // new id per step equals re-running the word LSTM over the whole prefix
        if (int rc = v2_word_step_bf16(B, probs != nullptr || t == 0, s)) return rc;
        if (int rc = v2_image_step_bf16(B, s)) return rc;
        if (probs) {
            TcEpilogue e;
            e.bias = W("imgcap_d1/bias"); e.out_f32 = ws.logits; e.ld_f32 = V;
            if (int rc = gemm_bf16_tc(op(b.v2_hb, U), op(b.v2_wd, U), e, B, V, U, kEpiStore, s)) return rc;
            if (int rc = softmax_argmax(ws.logits, V, B, V, probs + (size_t)t * V, (long long)(P - 1) * V, tokens + t, P - 1,
                                        ws.tok, scores ? ws.cand_p : nullptr, s)) return rc;
        } else {
            TcEpilogue e;
            e.bias = W("imgcap_d1/bias"); e.partial = b.partial;
            if (int rc = gemm_bf16_tc(op(b.v2_hb, U), op(b.v2_wd, U), e, B, V, U, scores ? kEpiArgmaxSum : kEpiArgmax, s)) return rc;
            const bool more = t + 2 < P;
            if (int rc = argmax_merge(b.partial, B, slots, tokens + t, P - 1, ws.tok, scores ? ws.cand_p : nullptr, s,
                                      more ? b.emb : nullptr, b.Epad, more ? b.X1[b.parity] : nullptr, b.Epad + Wu)) return rc;
        }
        if (scores)
            if (int rc = accumulate_log(scores, ws.cand_p, B, t == 0, s)) return rc;
    }
    return DC_OK;
}

}  // namespace dcap

// Training step of the v1 "inject" caption model (bf16 tensor-core path, fp32 master weights).
//
// Reference (paths relative to /root/reference/dense_img_cap_separate_models):
//   training graph   text_generation_model.py:159-189 (build_roi_caption_model_training), :264-277
//   loss             text_generation_model.py:286-294 (roi_caption_loss -> K.categorical_crossentropy)
//   optimiser        text_generation_model.py:425     (keras.optimizers.Adam(amsgrad=True))
//   data generator   text_generation_model.py:332-371 (targets = shift-left(gt) ++ [0])
//
// The reference evaluates the word model once per prefix position (O(P^2) LSTM steps); the
// causal, post-padded, masked structure makes that equal to ONE teacher-forced masked scan over
// gt[:, 0..P-1] (oracle/decoder.py asserts the two forms identical), which is what runs here.
//
// Layout: every per-(time, RoI) activation is TIME-MAJOR, row = t*B + b.  The two recurrent GEMMs
// (LSTM1, LSTM2: fused cell epilogues) run per step; everything that does not depend on the
// recurrence is ONE GEMM over all P*B rows: dense1, the vocabulary projection, their data
// gradients, and every weight gradient (X^T * dY with MN-major UMMA operands straight from the
// saved activations, split-K + red.global.add).  The [P*B, V] one-hot targets are never built: the
// fused softmax / cross-entropy kernel takes integer targets and overwrites the logits' role with
// dlogits in bf16.
//
// recurrent_dropout=0.2 of the reference's two LSTMs (text_generation_model.py:141-142) is available through
// dc_decoder_train_step_ex (DcTrainOptions): per-gate, time-invariant masks from a counter-based Philox generator --
// see "recurrent dropout" below.  It is OFF by default: parity with the oracle is defined on the deterministic graph,
// and TF's own RNG stream cannot be reproduced (the oracle draws the identical Philox masks instead).
#include "decoder.cuh"
#include "decoder_bf16.cuh"
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <math.h>
#include <type_traits>
#include <stdlib.h>

namespace dcap {

struct TrainState {
    int B = 0, T = 0;                       // capacity the buffers were sized for
    std::vector<void *> owned;
    int32_t *tok_tm = nullptr, *tgt_tm = nullptr;                  // [T, B]
    __nv_bfloat16 *X1 = nullptr, *X2 = nullptr;                    // [(T+1), B, K1] / [(T+1), B, 2U]
    float *c1 = nullptr, *c2 = nullptr;                            // [(T+1), B, U]
    __nv_bfloat16 *gates1 = nullptr, *gates2 = nullptr;            // [T, B, 4U] bf16 post-activation (i,f,g,o), gate-interleaved
    __nv_bfloat16 *d_all = nullptr;                                // [T*B, 1024] relu(dense1)
    float *logits = nullptr;                                       // [T*B, V] fp32, allocated on first use by teacher_forced_probs
    bool logits_in_dz = false;                                     // the last forward stored bf16 logits in dz
    __nv_bfloat16 *dz = nullptr;                                   // [T*B, V]   dlogits
    __nv_bfloat16 *dd = nullptr;                                   // [T*B, 1024]
    float *dh2d = nullptr;                                         // [T*B, U]   dense-path gradient of h2_t
    __nv_bfloat16 *dz1_all = nullptr, *dz2_all = nullptr;          // [T*B, 4U]  Keras block order i|f|c|o
    float *ddsum = nullptr, *dz1sum = nullptr;                     // [B, 1024] / [B, 4U] sums over time
    __nv_bfloat16 *ddsum_b = nullptr, *dz1sum_b = nullptr;
    float *dxh2 = nullptr, *dh1p = nullptr;                        // [T, B, 2U] (kept per step: the two layers' chains run concurrently) / [B, U]
    float *carry1 = nullptr, *carry2 = nullptr, *dc1 = nullptr, *dc2 = nullptr;   // [B, U]
    float *dF = nullptr, *da1 = nullptr;                           // [B, F]
    __nv_bfloat16 *dzh2 = nullptr, *dzh1 = nullptr;                // [B, F] head pre-activation gradients
    __nv_bfloat16 *x0 = nullptr;                                   // [B, Kin] bf16 copy of fp32 RoI features
    float *rowloss = nullptr;                                      // [T*B]
    const __nv_bfloat16 *x0_used = nullptr;                        // bf16 RoI features the last forward consumed
    cudaEvent_t bucket_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // gradient bucket i is complete on the step's stream
    // wavefront over the two LSTM layers: layer 1 of step t+1 and layer 2 of step t are independent (teacher forcing),
    // so each layer's chain of short kernels runs on its own stream, linked by one event per time step
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> step_ev;                              // [T] chain A -> chain B hand-over per time step
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    // recurrent dropout (DcTrainOptions::recurrent_dropout > 0): K-stacked operands, see dropout_* below
    int do_B = 0, do_T = 0;
    std::vector<void *> do_owned;
    __nv_bfloat16 *do_X1 = nullptr, *do_X2 = nullptr;               // [(T+1), B, Epad+4U], [(T+1), B, 5U]
    __nv_bfloat16 *do_H1 = nullptr, *do_H2 = nullptr;               // [(T+1), B, U] plain h (slot 0 = zero state)
    __nv_bfloat16 *do_w1 = nullptr, *do_w2 = nullptr;               // forward: [4U, Epad+4U], [4U, 5U] gate-interleaved rows
    __nv_bfloat16 *do_u1k = nullptr, *do_w2k = nullptr;             // backward: [4U, 4U], [5U, 4U] block rows, Keras columns
    float *do_dx = nullptr;                                         // [B, 5U] per-gate data gradients before the masked fold
    // v2 inject model (train_step_v2): word LSTM over L steps, one image-LSTM step, Dense(V)
    int v2_B = 0, v2_L = 0;
    std::vector<void *> v2_owned;
    int32_t *v2_tok_tm = nullptr, *v2_tgt_dummy = nullptr, *v2_ones = nullptr;   // [L, B], [L, B], [B]
    __nv_bfloat16 *v2_X = nullptr;                                  // [(L+1), B, Epad+Wu]
    float *v2_c = nullptr;                                          // [(L+1), B, Wu]
    __nv_bfloat16 *v2_gates_w = nullptr, *v2_gates_i = nullptr;     // [L, B, 4Wu], [B, 4U]
    float *v2_c_img = nullptr, *v2_logits = nullptr;                // [B, U], [B, V]
    __nv_bfloat16 *v2_dz = nullptr, *v2_dzi = nullptr, *v2_dzw = nullptr;   // [B, V], [B, 4U], [L, B, 4Wu]
    float *v2_dh_img = nullptr, *v2_dxin = nullptr, *v2_dh_prev = nullptr;  // [B, U], [B, F+Wu], [B, Wu]
    float *v2_carry = nullptr, *v2_dc = nullptr, *v2_carry_i = nullptr, *v2_dc_i = nullptr;
    float *v2_rowloss = nullptr;
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

void Decoder::free_train() {
    if (bf && bf->train) {
        for (void *p : bf->train->owned) cudaFree(p);
        for (void *p : bf->train->v2_owned) cudaFree(p);
        for (void *p : bf->train->do_owned) cudaFree(p);
        for (cudaEvent_t e : bf->train->bucket_ev)
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : bf->train->step_ev)
            if (e) cudaEventDestroy(e);
        if (bf->train->fork_ev) cudaEventDestroy(bf->train->fork_ev);
        if (bf->train->join_ev) cudaEventDestroy(bf->train->join_ev);
        if (bf->train->side) cudaStreamDestroy(bf->train->side);
        delete bf->train;
        bf->train = nullptr;
    }
}

int Decoder::ensure_grads() {
    if (grads) return DC_OK;
    const size_t bytes = sizeof(float) * (size_t)n_train;
    for (float **p : {&grads, &adam_m, &adam_v, &adam_vhat}) {
        DC_CHECK_CUDA(cudaMalloc((void **)p, bytes));
        DC_CHECK_CUDA(cudaMemset(*p, 0, bytes));
    }
    return DC_OK;
}

// bf16 mirror of the trainable arena: the Keras-layout tensors inside it are the K-major B operands of the
// data-gradient GEMMs.  After an optimiser step the mirror is already current (adam_amsgrad_kernel writes it).
int Decoder::refresh_train_weights(cudaStream_t s) {
    Bf16State &b = *bf;
    const size_t E = cfg.embed, U = cfg.units;
    if (!b.arena_k)
        if (int rc = dev_alloc((void **)&b.arena_k, 2 * (size_t)n_train, owned)) return rc;
    if (!b.arena_k_valid) {
        if (int rc = f32_to_bf16(arena, b.arena_k, n_train, s)) return rc;
        b.arena_k_valid = true;
    }
    auto K = [&](const char *name) { return b.arena_k + find(name)->offset; };
    // lstm2 kernel and recurrent_kernel are adjacent in the arena: [W2 ; U2] is one range
    DC_REQUIRE(W("imgcap_lstm2/recurrent_kernel") == W("imgcap_lstm2/kernel") + U * 4 * U, "arena layout");
    b.wd2_k = K("imgcap_lstm_d2/kernel");
    b.wd1h_k = K("imgcap_lstm_d1/kernel");
    b.wd1f_k = K("imgcap_lstm_d1/kernel") + U * kDense;
    b.w2cat_k = K("imgcap_lstm2/kernel");
    b.u1_k = K("imgcap_lstm1/recurrent_kernel");
    b.w1f_k = K("imgcap_lstm1/kernel") + E * 4 * U;
    b.wc2_k = K("mrcnn_class_conv2/kernel");
    return DC_OK;
}

void Decoder::invalidate_train_copy() {
    if (bf) bf->arena_k_valid = false;
}

static int train_reserve(Decoder &D, int B, int T) {
    Bf16State &b = *D.bf;
    if (b.train && b.train->B >= B && b.train->T >= T) return DC_OK;
    D.free_train();
    b.train = new TrainState();
    TrainState &t = *b.train;
    const DcDecoderConfig &c = D.cfg;
    const size_t U = c.units, F = c.feat, V = c.vocab, K1 = b.Epad + U, Kin = (size_t)c.pool * c.pool * c.channels;
    const size_t Bp = round_up(B, 128), R = (size_t)T * Bp;
    auto A = [&](void **p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
        if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "training workspace allocation failed: %s", cudaGetErrorString(e));
        t.owned.push_back(*p);
        return DC_OK;
    };
    int rc = 0;
    rc |= A((void **)&t.tok_tm, 4 * R); rc |= A((void **)&t.tgt_tm, 4 * R);
    rc |= A((void **)&t.X1, 2 * (R + Bp) * K1); rc |= A((void **)&t.X2, 2 * (R + Bp) * 2 * U);
    rc |= A((void **)&t.c1, 4 * (R + Bp) * U); rc |= A((void **)&t.c2, 4 * (R + Bp) * U);
    rc |= A((void **)&t.gates1, 2 * R * 4 * U); rc |= A((void **)&t.gates2, 2 * R * 4 * U);
    rc |= A((void **)&t.d_all, 2 * R * kDense);
    rc |= A((void **)&t.dz, 2 * R * V);                               // fp32 logits: only the predict surface needs them (lazy)
    rc |= A((void **)&t.dd, 2 * R * kDense); rc |= A((void **)&t.dh2d, 4 * R * U);
    rc |= A((void **)&t.dz1_all, 2 * R * 4 * U); rc |= A((void **)&t.dz2_all, 2 * R * 4 * U);
    rc |= A((void **)&t.ddsum, 4 * Bp * kDense); rc |= A((void **)&t.dz1sum, 4 * Bp * 4 * U);
    rc |= A((void **)&t.ddsum_b, 2 * Bp * kDense); rc |= A((void **)&t.dz1sum_b, 2 * Bp * 4 * U);
    rc |= A((void **)&t.dxh2, 4 * R * 2 * U); rc |= A((void **)&t.dh1p, 4 * Bp * U);
    rc |= A((void **)&t.carry1, 4 * Bp * U); rc |= A((void **)&t.carry2, 4 * Bp * U);
    rc |= A((void **)&t.dc1, 4 * Bp * U); rc |= A((void **)&t.dc2, 4 * Bp * U);
    rc |= A((void **)&t.dF, 4 * Bp * F); rc |= A((void **)&t.da1, 4 * Bp * F);
    rc |= A((void **)&t.dzh2, 2 * Bp * F); rc |= A((void **)&t.dzh1, 2 * Bp * F);
    rc |= A((void **)&t.x0, 2 * Bp * Kin);
    rc |= A((void **)&t.rowloss, 4 * R);
    if (rc) { D.free_train(); return rc; }
    bool ev_ok = cudaStreamCreateWithFlags(&t.side, cudaStreamNonBlocking) == cudaSuccess;
    ev_ok = ev_ok && cudaEventCreateWithFlags(&t.fork_ev, cudaEventDisableTiming) == cudaSuccess;
    ev_ok = ev_ok && cudaEventCreateWithFlags(&t.join_ev, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && ev_ok; ++i) ev_ok = cudaEventCreateWithFlags(&t.bucket_ev[i], cudaEventDisableTiming) == cudaSuccess;
    t.step_ev.assign(T, nullptr);
    for (int i = 0; i < T && ev_ok; ++i) ev_ok = cudaEventCreateWithFlags(&t.step_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ev_ok) {
        D.free_train();
        return set_error(DC_ERR_CUDA, "stream / event creation failed");
    }
    t.B = B; t.T = T;
    return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// gt [B, T] (ids as the data generator yields them) -> time-major token / target columns.
// targets == nullptr: targets = shift-left(gt) ++ [0]  (text_generation_model.py:352-358)
__global__ void time_major_tokens_kernel(const int32_t *__restrict__ gt, const int32_t *__restrict__ targets, int B,
                                         int T, int V, int32_t *__restrict__ tok_tm, int32_t *__restrict__ tgt_tm) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * T) return;
    const int t = idx / B, b = idx - t * B;
    int w = gt[(long long)b * T + t];
    w = w < 0 ? 0 : (w >= V ? V - 1 : w);
    tok_tm[idx] = w;
    int y = targets ? targets[(long long)b * T + t] : (t + 1 < T ? gt[(long long)b * T + t + 1] : 0);
    tgt_tm[idx] = y >= V ? V - 1 : y;                     // negative target = position ignored
}

// X[r, 0:Epad] = emb_bf16[tok[r], 0:Epad]  (frozen embedding, zero padded to Epad)
__global__ void gather_embedding_rows_kernel(const uint4 *__restrict__ emb, int emb_ld8, const int32_t *__restrict__ tok,
                                             long long rows, uint4 *__restrict__ X, long long ld8) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const uint4 *src = emb + (long long)tok[r] * emb_ld8;
    uint4 *dst = X + r * ld8;
    for (int j = lane; j < emb_ld8; j += 32) dst[j] = __ldg(src + j);
}

// Fused softmax + Keras categorical_crossentropy (forward value and gradient w.r.t. the logits) on
// integer targets; one CTA per (time, RoI) row.
//   p = softmax(z); loss_row = -log(clip(p_y, 1e-7, 1-1e-7)); dz = (p - onehot(y)) * inv_count, and
//   zero when the clip is active (the clip has zero gradient) or the position is ignored (y < 0).
template <typename TIn>
__device__ __forceinline__ void load8(const TIn *p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

// TIn = float: logits as the vocabulary GEMM accumulated them (predict surface, v2).  TIn = bf16: the training step's
// vocabulary GEMM stores its logits ONCE, as bf16, straight into the dlogits buffer, and this kernel turns them into
// dlogits IN PLACE (dz == logits): the fp32 [T*B, V] logits (2.6 GB written + read per cfg3 step) never exist.
// V % 8 == 0.  In place is safe: every thread re-reads exactly the elements it then overwrites, after the row
// statistics (first pass) are complete.
template <typename TIn>
__global__ void __launch_bounds__(256) softmax_xent_kernel(const TIn *logits, long long ld, int V,
                                                           const int32_t *__restrict__ tgt, float inv_count,
                                                           __nv_bfloat16 *dz, long long ld_dz,
                                                           float *__restrict__ rowloss) {
    const long long r = blockIdx.x;
    const TIn *z = logits + r * ld;
    __nv_bfloat16 *g = dz + r * ld_dz;
    const int y = tgt[r];
    __shared__ float red[2][8];
    float mx = -INFINITY, sum = 0.f;
    for (int j = threadIdx.x * 8; j < V; j += 256 * 8) {
        float v[8];
        load8<TIn>(z + j, v);
        float m8 = v[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) m8 = fmaxf(m8, v[i]);
        if (m8 > mx) { sum *= __expf(mx - m8); mx = m8; }
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += __expf(v[i] - mx);
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
        const float nm = fmaxf(mx, om);
        sum = (mx > -INFINITY ? sum * __expf(mx - nm) : 0.f) + (om > -INFINITY ? os * __expf(om - nm) : 0.f);
        mx = nm;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = mx; red[1][warp] = sum; }
    __syncthreads();
    float gm = red[0][0];
    for (int i = 1; i < 8; ++i) gm = fmaxf(gm, red[0][i]);
    float gs = 0.f;
    for (int i = 0; i < 8; ++i) gs += red[0][i] > -INFINITY ? red[1][i] * __expf(red[0][i] - gm) : 0.f;
    const float inv = 1.0f / gs;
    bool live = y >= 0;
    float py = 0.f;
    if (live) py = __expf((float)z[y] - gm) * inv;
    __syncthreads();                                       // z[y] is read by every thread before anyone overwrites it (in place)
    if (live) {
        if (threadIdx.x == 0) rowloss[r] = -logf(fminf(fmaxf(py, 1e-7f), 1.0f - 1e-7f));
        live = py > 1e-7f && py < 1.0f - 1e-7f;
    } else if (threadIdx.x == 0) {
        rowloss[r] = 0.f;
    }
    const float sc = live ? inv_count : 0.f;
    for (int j = threadIdx.x * 8; j < V; j += 256 * 8) {
        float v[8];
        load8<TIn>(z + j, v);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float p0 = __expf(v[2 * i] - gm) * inv, p1 = __expf(v[2 * i + 1] - gm) * inv;
            if (y == j + 2 * i) p0 -= 1.0f;
            if (y == j + 2 * i + 1) p1 -= 1.0f;
            __nv_bfloat162 pk = __floats2bfloat162_rn(p0 * sc, p1 * sc);
            w[i] = *reinterpret_cast<uint32_t *>(&pk);
        }
        *reinterpret_cast<uint4 *>(g + j) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- fused form for the training step: softmax + cross-entropy + dlogits IN PLACE + vocabulary-bias gradient ----
// The one-CTA-per-row kernel above reads every row twice through L1/L2 with ~32 KB of loads in flight per SM and two
// exponentials per element (3.9 TB/s), and the bias gradient then re-reads all 1.3 GB of dlogits (colsum_kernel:
// 0.28 ms per cfg3 step).  Here two persistent CTAs per SM stream blocks of G rows through two-slot shared-memory rings
// filled by 1-D bulk copies (cp.async.bulk, one mbarrier per slot: no register stands behind a byte in flight);
// thread t owns the same 8-column vectors (t + 512 k) of EVERY row, so the column sums of dlogits stay in its registers
// across all the rows the CTA sees and leave with V atomics per CTA at the end.  Three passes over a slot, one
// exponential per element: (A) row maxima; (B) e = exp(z - max), row sums, e written back into the slot as bf16;
// (C) dlogits = (e / sum - onehot) * scale to global memory, column sums.  The target's own probability (loss, clip
// decision) is taken from the fp32 e of the thread that owns its column, not from the bf16 copy.
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kXentThreads = 512, kXentWarps = kXentThreads / 32;

// one MUFU.EX2 (exp2f() without fast-math brackets it with a range test and two scalings; flushing the results that
// would be denormal is exactly right for a softmax numerator)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NV, int G>      // NV 8-column vectors per thread (V <= NV * 4096), G rows per ring slot
__global__ void __launch_bounds__(kXentThreads, 2) softmax_xent_colsum_kernel(__nv_bfloat16 *z, long long ld, int V, long long R,
                                                                               const int32_t *__restrict__ tgt, float inv_count,
                                                                               float *__restrict__ rowloss,
                                                                               float *__restrict__ bias_grad) {
    extern __shared__ __align__(128) unsigned char xent_smem[];
    __shared__ uint64_t full[2];
    __shared__ float red_m[kXentWarps][G], red_s[kXentWarps][G];
    __shared__ float st_ey[G];
    __shared__ int st_y[G];
    constexpr float kLog2e = 1.4426950408889634f;
    // the dispatcher picks NV from V's range, so only the last vector(s) of a thread can lie beyond V
    constexpr int kNoCheck = NV == 6 ? 4 : NV - 1;
    const uint32_t rowbytes = (uint32_t)V * 2;
    const long long nblk = (R + G - 1) / G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long blk, int b) {
        const long long r0 = blk * G;
        const int rows = (int)(R - r0 < G ? R - r0 : G);
        mbar_expect_tx(&full[b], (uint32_t)rows * rowbytes);
        for (int g = 0; g < rows; ++g)
            bulk_load_1d(xent_smem + ((size_t)b * G + g) * rowbytes, z + (r0 + g) * ld, rowbytes, &full[b]);
    };
    float acc[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    long long blk = blockIdx.x;
    if (tid == 0 && blk < nblk) issue(blk, 0);
    for (int it = 0; blk < nblk; blk += gridDim.x, ++it) {
        const int b = it & 1;
        // slot b^1 was released by the barrier that closed the previous iteration
        if (tid == 0 && blk + gridDim.x < nblk) issue(blk + gridDim.x, b ^ 1);
        const long long r0 = blk * G;
        const int rows = (int)(R - r0 < G ? R - r0 : G);
        if (tid < rows) st_y[tid] = tgt[r0 + tid];
        mbar_wait(&full[b], (uint32_t)(it >> 1) & 1u);
        unsigned char *buf = xent_smem + (size_t)b * G * rowbytes;
        // FULL = all G rows of the slot are live (every block but a ragged last one): no per-row predicates
        auto passes = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        // ---- (A) row maxima, on the packed bf16 pairs as they lie in the slot (max is exact in any format)
        __nv_bfloat162 mx2[G];
#pragma unroll
        for (int g = 0; g < G; ++g) mx2[g] = __float2bfloat162_rn(-INFINITY);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int j = (tid + k * kXentThreads) * 8;
            if (k < kNoCheck || j < V) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (FULL || g < rows) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(buf + (size_t)g * rowbytes + (size_t)j * 2);
                        const __nv_bfloat162 a = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&q.x), *reinterpret_cast<const __nv_bfloat162 *>(&q.y));
                        const __nv_bfloat162 c = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&q.z), *reinterpret_cast<const __nv_bfloat162 *>(&q.w));
                        mx2[g] = __hmax2(mx2[g], __hmax2(a, c));
                    }
                }
            }
        }
        float m[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            m[g] = fmaxf(__low2float(mx2[g]), __high2float(mx2[g]));
            for (int o = 16; o > 0; o >>= 1) m[g] = fmaxf(m[g], __shfl_xor_sync(0xffffffffu, m[g], o));
            if (lane == 0) red_m[warp][g] = m[g];
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < G; ++g) {                       // 16 warp partials: one load per lane and a 4-step shuffle tree
            float gm = red_m[lane & (kXentWarps - 1)][g];
#pragma unroll
            for (int o = kXentWarps / 2; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
            m[g] = gm * kLog2e;
            // the thread that owns the target's column keeps its fp32 exponential: loss and clip decision do not see the
            // bf16 copy (it reads the logit before its own pass-B store overwrites it)
            const int y = st_y[g];
            if ((FULL || g < rows) && y >= 0 && ((y >> 3) & (kXentThreads - 1)) == tid)
                st_ey[g] = ex2_ftz(fmaf(__bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(buf + (size_t)g * rowbytes)[y]), kLog2e, -m[g]));
        }
        // ---- (B) e = exp(z - max) once per element, row sums; e goes back into the slot as bf16
        float sum[G];
#pragma unroll
        for (int g = 0; g < G; ++g) sum[g] = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int j = (tid + k * kXentThreads) * 8;
            if (k < kNoCheck || j < V) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (FULL || g < rows) {
                        __nv_bfloat16 *row = reinterpret_cast<__nv_bfloat16 *>(buf + (size_t)g * rowbytes) + j;
                        float v[8];
                        load8<__nv_bfloat16>(row, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { v[i] = ex2_ftz(fmaf(v[i], kLog2e, -m[g])); sum[g] += v[i]; }
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __nv_bfloat162 pk = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            w[i] = *reinterpret_cast<uint32_t *>(&pk);
                        }
                        *reinterpret_cast<uint4 *>(row) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            for (int o = 16; o > 0; o >>= 1) sum[g] += __shfl_xor_sync(0xffffffffu, sum[g], o);
            if (lane == 0) red_s[warp][g] = sum[g];
        }
        __syncthreads();
        // ---- (C) dlogits to global memory, column sums in registers
        float scale[G];                                    // inv * (gradient scale, or 0 when the clip is active / no target)
        float sub[G];                                      // what the target's column loses: that scale without the 1 / sum
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float gs = red_s[lane & (kXentWarps - 1)][g];
#pragma unroll
            for (int o = kXentWarps / 2; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
            const float inv = 1.0f / gs;
            bool live = (FULL || g < rows) && st_y[g] >= 0;
            float loss = 0.f;
            if (live) {
                const float py = st_ey[g] * inv;
                loss = -logf(fminf(fmaxf(py, 1e-7f), 1.0f - 1e-7f));
                live = py > 1e-7f && py < 1.0f - 1e-7f;
            }
            if (tid == g && (FULL || g < rows)) rowloss[r0 + g] = loss;
            sub[g] = live ? inv_count : 0.f;
            scale[g] = inv * sub[g];
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int j = (tid + k * kXentThreads) * 8;
            if (k < kNoCheck || j < V) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (FULL || g < rows) {
                        float v[8];
                        load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16 *>(buf + (size_t)g * rowbytes) + j, v);
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { v[i] *= scale[g]; acc[k][i] += v[i]; }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __nv_bfloat162 pk = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            w[i] = *reinterpret_cast<uint32_t *>(&pk);
                        }
                        *reinterpret_cast<uint4 *>(z + (r0 + g) * ld + j) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
        // the "- onehot" term: the owner of the target's column rewrites that one element behind its own vector store
        // (program order within the thread) and takes it out of the column sum
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int y = st_y[g];
            if ((FULL || g < rows) && y >= 0 && ((y >> 3) & (kXentThreads - 1)) == tid && sub[g] != 0.f) {
                const float e = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(buf + (size_t)g * rowbytes)[y]);
                z[(r0 + g) * ld + y] = __float2bfloat16_rn(e * scale[g] - sub[g]);
                atomicAdd(bias_grad + y, -sub[g]);
            }
        }
        };
        if (rows == G) passes(std::true_type{});
        else passes(std::false_type{});
        // pass B wrote into the slot through the generic proxy and the next bulk copy into it goes through the async proxy:
        // order the two before the barrier that releases the slot
        fence_async_smem();
        __syncthreads();                                   // every read of slot b (and of st_*) is done: it may be refilled
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int j = (tid + k * kXentThreads) * 8;
        if (j < V) {
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(bias_grad + j + i, acc[k][i]);
        }
    }
}

template <int NV, int G>
static int launch_xent_colsum(__nv_bfloat16 *z, long long ld, int V, long long R, const int32_t *tgt, float inv_count,
                              float *rowloss, float *bias_grad, cudaStream_t s) {
    const size_t smem = 2 * (size_t)G * V * 2;
    static bool attr_set = false;
    if (!attr_set) {
        DC_CHECK_CUDA(cudaFuncSetAttribute(softmax_xent_colsum_kernel<NV, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    const long long nblk = (R + G - 1) / G, slots = 2ll * sm_count();
    const unsigned grid = (unsigned)(nblk < slots ? nblk : slots);
    softmax_xent_colsum_kernel<NV, G><<<grid, kXentThreads, smem, s>>>(z, ld, V, R, tgt, inv_count, rowloss, bias_grad);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// returns false when the shape is outside the fused kernel's range (V > 24576, unaligned rows): the caller then runs
// softmax_xent_kernel + colsum
static bool xent_colsum_supported(const void *z, long long ld, int V) {
    return V % 8 == 0 && V <= 24576 && (ld * 2) % 16 == 0 && ((uintptr_t)z & 15) == 0;
}

// ring slots of at most 48 KB so that two CTAs fit on an SM
static int softmax_xent_colsum(__nv_bfloat16 *z, long long ld, int V, long long R, const int32_t *tgt, float inv_count,
                               float *rowloss, float *bias_grad, cudaStream_t s) {
    if (R <= 0) return DC_OK;
    if (V <= 4096) return launch_xent_colsum<1, 2>(z, ld, V, R, tgt, inv_count, rowloss, bias_grad, s);
    if (V <= 8192) return launch_xent_colsum<2, 2>(z, ld, V, R, tgt, inv_count, rowloss, bias_grad, s);
    if (V <= 12288) return launch_xent_colsum<3, 2>(z, ld, V, R, tgt, inv_count, rowloss, bias_grad, s);
    if (V <= 16384) return launch_xent_colsum<4, 1>(z, ld, V, R, tgt, inv_count, rowloss, bias_grad, s);
    return launch_xent_colsum<6, 1>(z, ld, V, R, tgt, inv_count, rowloss, bias_grad, s);
}

// loss = sum(rowloss) * inv_count, deterministic (one CTA, fixed order)
__global__ void __launch_bounds__(1024) reduce_loss_kernel(const float *__restrict__ rowloss, long long n, float inv_count,
                                                           float *__restrict__ loss) {
    __shared__ double red[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) acc += (double)rowloss[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = red[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) *loss = (float)(acc * (double)inv_count);
    }
}

// out[c] += sum_r src[r, c]; grid (column strips of 256, row chunks); src bf16 or fp32
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T *__restrict__ src, long long rows, int cols, long long ld,
                                                     int rows_per_block, float *__restrict__ out) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) {
        if constexpr (sizeof(T) == 2) acc += __bfloat162float(src[r * ld + c]);
        else acc += src[r * ld + c];
    }
    atomicAdd(out + c, acc);
}

template <typename T>
static int colsum(const T *src, long long rows, int cols, long long ld, float *out, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return DC_OK;
    const int rpb = 256;
    dim3 grid(ceil_div(cols, 256), (unsigned)ceil_div<long long>(rows, rpb));
    colsum_kernel<T><<<grid, 256, 0, s>>>(src, rows, cols, ld, rpb, out);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// sum over time of a time-major bf16 tensor: out[b, n] = sum_t src[(t*B + b), n]
__global__ void time_sum_kernel(const __nv_bfloat16 *__restrict__ src, int B, int T, int N, float *__restrict__ out_f32,
                                __nv_bfloat16 *__restrict__ out_bf16) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * N) return;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += __bfloat162float(src[(long long)t * B * N + idx]);
    out_f32[idx] = acc;
    out_bf16[idx] = __float2bfloat16_rn(acc);
}

__global__ void f32_to_bf16_rows_kernel(const float *__restrict__ src, long long n, __nv_bfloat16 *__restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// Backward of one Keras LSTM cell step for 4 consecutive units of one row.
//   gates      [B, 4U] post-activation (i, f, g, o), gate-interleaved (column 4u+gate), saved by the forward epilogue
//   c_prev/c_new  the cell state before / after the step
//   dh = dh_a (+ dh_b) + carry;  masked row (tok == 0): dz = 0, carry <- dh, dc unchanged
//   else: dz (Keras block order i|f|c|o, bf16), dc <- dc_prev, carry <- 0, dzsum += dz (optional)
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(int B, int U, const __nv_bfloat16 *__restrict__ gates,
                                                            const float *__restrict__ c_prev, const float *__restrict__ c_new,
                                                            const int32_t *__restrict__ tok, const float *__restrict__ dh_a,
                                                            int ld_a, const float *__restrict__ dh_b, int ld_b,
                                                            float *__restrict__ carry, float *__restrict__ dc,
                                                            __nv_bfloat16 *__restrict__ dz, float *__restrict__ dzsum) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int uq = U >> 2;
    pdl_wait();                                           // launched with PDL between the BPTT GEMMs
    pdl_launch_dependents();
    if (idx >= (long long)B * uq) return;
    const int b = (int)(idx / uq), u0 = (int)(idx - (long long)b * uq) * 4;
    const long long so = (long long)b * U + u0;
    float4 dh = *reinterpret_cast<const float4 *>(dh_a + (long long)b * ld_a + u0);
    if (dh_b) {
        const float4 t = *reinterpret_cast<const float4 *>(dh_b + (long long)b * ld_b + u0);
        dh.x += t.x; dh.y += t.y; dh.z += t.z; dh.w += t.w;
    }
    {
        const float4 t = *reinterpret_cast<const float4 *>(carry + so);
        dh.x += t.x; dh.y += t.y; dh.z += t.z; dh.w += t.w;
    }
    __nv_bfloat16 *dzr = dz + (long long)b * 4 * U + u0;
    if (tok[b] == 0) {
        *reinterpret_cast<float4 *>(carry + so) = dh;
        const uint2 zero = make_uint2(0u, 0u);
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint2 *>(dzr + (long long)g * U) = zero;
        return;
    }
    *reinterpret_cast<float4 *>(carry + so) = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 dcv = *reinterpret_cast<const float4 *>(dc + so);
    const float4 cpv = *reinterpret_cast<const float4 *>(c_prev + so);
    const float4 cnv = *reinterpret_cast<const float4 *>(c_new + so);
    const float dhs[4] = {dh.x, dh.y, dh.z, dh.w}, dcs[4] = {dcv.x, dcv.y, dcv.z, dcv.w};
    const float cps[4] = {cpv.x, cpv.y, cpv.z, cpv.w}, cns[4] = {cnv.x, cnv.y, cnv.z, cnv.w};
    float dzi[4], dzf[4], dzg[4], dzo[4], dcp[4];
    const uint4 *gp = reinterpret_cast<const uint4 *>(gates + (long long)b * 4 * U + 4 * u0);      // 4 units x 4 gates bf16
    const uint4 gq[2] = {gp[0], gp[1]};
    const uint32_t gw[8] = {gq[0].x, gq[0].y, gq[0].z, gq[0].w, gq[1].x, gq[1].y, gq[1].z, gq[1].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float ig = __uint_as_float(gw[2 * j] << 16), fg = __uint_as_float(gw[2 * j] & 0xffff0000u);
        const float gg = __uint_as_float(gw[2 * j + 1] << 16), og = __uint_as_float(gw[2 * j + 1] & 0xffff0000u);
        const float tc = tanhf(cns[j]);
        const float dcn = dcs[j] + dhs[j] * og * (1.f - tc * tc);
        // hard_sigmoid'(z) = 0.2 inside the linear range, i.e. where the activation is strictly in (0, 1)
        dzi[j] = (ig > 0.f && ig < 1.f) ? 0.2f * dcn * gg : 0.f;
        dzf[j] = (fg > 0.f && fg < 1.f) ? 0.2f * dcn * cps[j] : 0.f;
        dzg[j] = dcn * ig * (1.f - gg * gg);
        dzo[j] = (og > 0.f && og < 1.f) ? 0.2f * dhs[j] * tc : 0.f;
        dcp[j] = dcn * fg;
    }
    *reinterpret_cast<float4 *>(dc + so) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    const float *blocks[4] = {dzi, dzf, dzg, dzo};
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(blocks[g][0], blocks[g][1]), hi = __floats2bfloat162_rn(blocks[g][2], blocks[g][3]);
        *reinterpret_cast<uint2 *>(dzr + (long long)g * U) = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
        if (dzsum) {
            float4 *sp = reinterpret_cast<float4 *>(dzsum + (long long)b * 4 * U + (long long)g * U + u0);
            float4 a = *sp;
            a.x += blocks[g][0]; a.y += blocks[g][1]; a.z += blocks[g][2]; a.w += blocks[g][3];
            *sp = a;
        }
    }
}

// Backward through relu(BatchNorm_frozen(z)) of the RoI head (modified_dense_model.py:52-62: BN runs
// with training=False, gamma/beta stay trainable): y = z*s + (beta - mean*s), s = gamma*rsqrt(var+eps).
//   dy = dx where act > 0;  dz = dy*s (bf16, the conv data/weight gradient operand)
//   dbeta += sum dy;  dgamma += sum dy*(z-mean)*rsqrt(var+eps) = sum dy*(act-beta)/gamma;  dbias += sum dz
// block (32, 8): a 32-column strip x rows_per_block rows.
__global__ void __launch_bounds__(256) bn_relu_bwd_kernel(const float *__restrict__ dx, const __nv_bfloat16 *__restrict__ act,
                                                          int B, int N, const float *__restrict__ scale,
                                                          const float *__restrict__ gamma, const float *__restrict__ beta,
                                                          int rows_per_block, __nv_bfloat16 *__restrict__ dz,
                                                          float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                          float *__restrict__ dbias) {
    const int n = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(r0 + rows_per_block, B);
    __shared__ float red[3][8][33];
    float sg = 0.f, sb = 0.f, sz = 0.f;
    if (n < N) {
        const float s = scale[n], ga = gamma[n], be = beta[n];
        const float inv_g = ga != 0.f ? 1.0f / ga : 0.f;
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            const long long o = (long long)r * N + n;
            const float a = __bfloat162float(act[o]);
            const float dy = a > 0.f ? dx[o] : 0.f;
            const float v = dy * s;
            dz[o] = __float2bfloat16_rn(v);
            sb += dy; sg += dy * (a - be) * inv_g; sz += v;
        }
    }
    red[0][threadIdx.y][threadIdx.x] = sg; red[1][threadIdx.y][threadIdx.x] = sb; red[2][threadIdx.y][threadIdx.x] = sz;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float a = 0.f, b2 = 0.f, c = 0.f;
        for (int i = 0; i < 8; ++i) { a += red[0][i][threadIdx.x]; b2 += red[1][i][threadIdx.x]; c += red[2][i][threadIdx.x]; }
        atomicAdd(dgamma + n, a); atomicAdd(dbeta + n, b2); atomicAdd(dbias + n, c);
    }
}

// keras.optimizers.Adam.get_updates (amsgrad optional), SURVEY.md A11:
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g^2;  vhat = max(vhat, v);  p -= lr_t * m / (sqrt(vhat) + eps)
// with lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed on the host in double; g is scaled by grad_scale first
// (1/world after a sum all-reduce).
__global__ void adam_amsgrad_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                    float *__restrict__ v, float *__restrict__ vhat, long long n, float lr_t, float b1,
                                    float b2, float eps, int amsgrad, float grad_scale, __nv_bfloat16 *__restrict__ p_b16) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;                                   // n is a multiple of 64
    const float4 g4 = *reinterpret_cast<const float4 *>(g + i);
    float4 m4 = *reinterpret_cast<float4 *>(m + i), v4 = *reinterpret_cast<float4 *>(v + i);
    float4 h4 = *reinterpret_cast<float4 *>(vhat + i), p4 = *reinterpret_cast<float4 *>(p + i);
    const float gs[4] = {g4.x * grad_scale, g4.y * grad_scale, g4.z * grad_scale, g4.w * grad_scale};
    float ms[4] = {m4.x, m4.y, m4.z, m4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w}, hs[4] = {h4.x, h4.y, h4.z, h4.w};
    float ps[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ms[j] = __fadd_rn(__fmul_rn(b1, ms[j]), __fmul_rn(1.f - b1, gs[j]));
        vs[j] = __fadd_rn(__fmul_rn(b2, vs[j]), __fmul_rn(1.f - b2, __fmul_rn(gs[j], gs[j])));
        float den;
        if (amsgrad) { hs[j] = fmaxf(hs[j], vs[j]); den = __fadd_rn(sqrtf(hs[j]), eps); }
        else den = __fadd_rn(sqrtf(vs[j]), eps);
        ps[j] = __fsub_rn(ps[j], __fdiv_rn(__fmul_rn(lr_t, ms[j]), den));
    }
    *reinterpret_cast<float4 *>(m + i) = make_float4(ms[0], ms[1], ms[2], ms[3]);
    *reinterpret_cast<float4 *>(v + i) = make_float4(vs[0], vs[1], vs[2], vs[3]);
    if (amsgrad) *reinterpret_cast<float4 *>(vhat + i) = make_float4(hs[0], hs[1], hs[2], hs[3]);
    *reinterpret_cast<float4 *>(p + i) = make_float4(ps[0], ps[1], ps[2], ps[3]);
    if (p_b16) {                                          // bf16 mirror used by the next step's data-gradient GEMMs
        __nv_bfloat162 lo = __floats2bfloat162_rn(ps[0], ps[1]), hi = __floats2bfloat162_rn(ps[2], ps[3]);
        *reinterpret_cast<uint2 *>(p_b16 + i) = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
    }
}

// ------------------------------------------------------------------------------------------------
// recurrent dropout (KL.LSTM(..., recurrent_dropout=0.2), text_generation_model.py:141-142; a10)
//
// Keras draws four masks per LSTM call (one per gate, constant over time) and multiplies h_{t-1} by mask g before the
// recurrent product of gate g.  Four differently masked copies of h would break the "one GEMM per LSTM" stacking, so the
// masked copies are STACKED ALONG K instead: A = [x | h*m_i | h*m_f | h*m_c | h*m_o] against a weight whose recurrent
// part is block-diagonal over the gates (row 4u+g only sees block g).  The fused cell epilogue, the gate-interleaved
// layout and the tcgen05 kernels are untouched; the recurrent K grows 4x (only in this mode), and three small kernels
// do the rest: expand (h -> 4 masked copies), fold (sum_g m_g * dX_g) and the block-structured weight builders.
// Masks: Philox-4x32-10, key = seed, counter = (global row, unit, layer, step) -> four words = four gates; stateless,
// identical for any sharding of the batch; the oracle draws the same ones (oracle/decoder.py: philox_masks).
// ------------------------------------------------------------------------------------------------
struct DropoutCtx {
    float rate, scale;            // scale = 1/(1-rate)
    unsigned k0, k1, step, thr;   // keep iff word >= thr = rate * 2^32
    long long row_offset;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ void dropout_masks4(const DropoutCtx &d, int layer, long long row, int unit, float (&m)[4]) {
    const unsigned long long gr = (unsigned long long)(row + d.row_offset);
    const uint4 w = philox4x32_10(make_uint4((unsigned)gr, (unsigned)unit, (unsigned)(gr >> 32) | ((unsigned)layer << 16), d.step), d.k0, d.k1);
    m[0] = w.x >= d.thr ? d.scale : 0.f; m[1] = w.y >= d.thr ? d.scale : 0.f;
    m[2] = w.z >= d.thr ? d.scale : 0.f; m[3] = w.w >= d.thr ? d.scale : 0.f;
}

// out[b, g*U + u] = bf16(h[b, u] * m_g(b, u))
__global__ void dropout_expand_kernel(const __nv_bfloat16 *__restrict__ h, long long ld_h, int B, int U, int layer, DropoutCtx d,
                                      __nv_bfloat16 *__restrict__ out, long long ld_out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * U) return;
    const int b = (int)(idx / U), u = (int)(idx - (long long)b * U);
    float m[4];
    dropout_masks4(d, layer, b, u, m);
    const float v = __bfloat162float(h[(long long)b * ld_h + u]);
#pragma unroll
    for (int g = 0; g < 4; ++g) out[(long long)b * ld_out + (long long)g * U + u] = __float2bfloat16_rn(v * m[g]);
}

// out[b, u] = sum_g m_g(b, u) * dx[b, g*U + u]; optional straight copy of `ncopy` leading columns of src_copy
__global__ void dropout_fold_kernel(const float *__restrict__ dx, long long ld_dx, int B, int U, int layer, DropoutCtx d,
                                    float *__restrict__ out, long long ld_out, const float *__restrict__ src_copy, int ncopy,
                                    float *__restrict__ dst_copy) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * U) return;
    const int b = (int)(idx / U), u = (int)(idx - (long long)b * U);
    float m[4];
    dropout_masks4(d, layer, b, u, m);
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) acc += m[g] * dx[(long long)b * ld_dx + (long long)g * U + u];
    out[(long long)b * ld_out + u] = acc;
    if (src_copy && u < ncopy) dst_copy[(long long)b * ld_out + u] = src_copy[(long long)b * ld_dx + u];
}

// forward operand: dst[(4u+g), k_off + g*U + k] = bf16(R[k, g*U + u])   (R = Keras recurrent kernel [U, 4U]; rest pre-zeroed)
__global__ void dropout_build_fwd_kernel(const float *__restrict__ R, int U, __nv_bfloat16 *__restrict__ dst, long long ld, int k_off) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)U * 4 * U) return;
    const int k = (int)(idx / (4 * U)), col = (int)(idx - (long long)k * 4 * U), g = col / U, u = col - g * U;
    dst[(long long)(4 * u + g) * ld + k_off + (long long)g * U + k] = __float2bfloat16_rn(R[idx]);
}

// backward operand (B of dX = dz * W^T, [N = in, K = out] K-major): dst[(row_off + g*U + k), g*U + u] = bf16(R[k, g*U + u])
__global__ void dropout_build_bwd_kernel(const float *__restrict__ R, int U, __nv_bfloat16 *__restrict__ dst, int row_off) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)U * 4 * U) return;
    const int k = (int)(idx / (4 * U)), col = (int)(idx - (long long)k * 4 * U), g = col / U;
    dst[(long long)(row_off + g * U + k) * 4 * U + col] = __float2bfloat16_rn(R[idx]);
}

static DropoutCtx dropout_ctx(const DcTrainOptions *o) {
    DropoutCtx d;
    d.rate = o->recurrent_dropout; d.scale = 1.0f / (1.0f - d.rate);
    d.k0 = (unsigned)(o->dropout_seed & 0xffffffffull); d.k1 = (unsigned)(o->dropout_seed >> 32);
    d.step = (unsigned)((unsigned long long)o->dropout_step & 0xffffffffull);
    d.thr = (unsigned)((double)d.rate * 4294967296.0);
    d.row_offset = o->row_offset;
    return d;
}

static int dropout_reserve(Decoder &D, int B, int T) {
    Bf16State &b = *D.bf;
    TrainState &t = *b.train;
    if (t.do_B >= B && t.do_T >= T) return DC_OK;
    for (void *p : t.do_owned) cudaFree(p);
    t.do_owned.clear();
    const size_t U = D.cfg.units, K1d = b.Epad + 4 * U, Bp = round_up(B, 128), R = (size_t)(T + 1) * Bp;
    auto A = [&](void **p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "dropout workspace allocation failed: %s", cudaGetErrorString(e));
        t.do_owned.push_back(*p);
        return DC_OK;
    };
    int rc = 0;
    rc |= A((void **)&t.do_X1, 2 * R * K1d); rc |= A((void **)&t.do_X2, 2 * R * 5 * U);
    rc |= A((void **)&t.do_H1, 2 * R * U); rc |= A((void **)&t.do_H2, 2 * R * U);
    rc |= A((void **)&t.do_w1, 2 * 4 * U * K1d); rc |= A((void **)&t.do_w2, 2 * 4 * U * 5 * U);
    rc |= A((void **)&t.do_u1k, 2 * 4 * U * 4 * U); rc |= A((void **)&t.do_w2k, 2 * 5 * U * 4 * U);
    rc |= A((void **)&t.do_dx, 4 * Bp * 5 * U);
    if (rc) { t.do_B = t.do_T = 0; return rc; }
    t.do_B = B; t.do_T = T;
    return DC_OK;
}

// the block-structured weights follow the master weights: rebuilt at every step of this mode (16 MB of writes)
static int dropout_build_weights(Decoder &D, cudaStream_t s) {
    Bf16State &b = *D.bf;
    TrainState &t = *b.train;
    const int U = D.cfg.units, E = D.cfg.embed, K1d = b.Epad + 4 * U;
    DC_CHECK_CUDA(cudaMemsetAsync(t.do_w1, 0, 2 * (size_t)4 * U * K1d, s));
    DC_CHECK_CUDA(cudaMemsetAsync(t.do_w2, 0, 2 * (size_t)4 * U * 5 * U, s));
    DC_CHECK_CUDA(cudaMemsetAsync(t.do_u1k, 0, 2 * (size_t)4 * U * 4 * U, s));
    DC_CHECK_CUDA(cudaMemsetAsync(t.do_w2k, 0, 2 * (size_t)5 * U * 4 * U, s));
    if (int rc = build_kmajor(D.W("imgcap_lstm1/kernel"), 4 * U, 0, E, 4 * U, U, t.do_w1, K1d, 0, s)) return rc;
    if (int rc = build_kmajor(D.W("imgcap_lstm2/kernel"), 4 * U, 0, U, 4 * U, U, t.do_w2, 5 * U, 0, s)) return rc;
    const unsigned grid = (unsigned)ceil_div<long long>((long long)U * 4 * U, 256);
    dropout_build_fwd_kernel<<<grid, 256, 0, s>>>(D.W("imgcap_lstm1/recurrent_kernel"), U, t.do_w1, K1d, b.Epad);
    dropout_build_fwd_kernel<<<grid, 256, 0, s>>>(D.W("imgcap_lstm2/recurrent_kernel"), U, t.do_w2, 5 * U, U);
    dropout_build_bwd_kernel<<<grid, 256, 0, s>>>(D.W("imgcap_lstm1/recurrent_kernel"), U, t.do_u1k, 0);
    dropout_build_bwd_kernel<<<grid, 256, 0, s>>>(D.W("imgcap_lstm2/recurrent_kernel"), U, t.do_w2k, U);
    DC_CHECK_LAUNCH();
    // rows [0, U) of the layer-2 backward operand = the input kernel W2 [U, 4U] as it lies in the bf16 arena mirror
    DC_CHECK_CUDA(cudaMemcpyAsync(t.do_w2k, b.w2cat_k, 2 * (size_t)U * 4 * U, cudaMemcpyDeviceToDevice, s));
    return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------------
// Teacher-forced forward up to the logits [T*B, V] (time-major) with every activation the backward needs.
int Decoder::train_forward(const void *feats, int kind, int B, const int32_t *gt, const int32_t *targets, cudaStream_t s) {
    const bool want_bf16_logits = train_logits_bf16;                   // set by train_step only: predict() keeps fp32 logits
    PdlScope pdl;                                          // chained GEMMs overlap their prologues with the predecessor's tail
    if (int rc = check_ready(B)) return rc;
    DC_REQUIRE(cfg.arch == DC_ARCH_V1 && cfg.dtype == DC_DTYPE_BF16, "the training graph is served by the bf16 v1 decoder");
    DC_REQUIRE(cfg.vocab % 8 == 0, "training needs VOCABULARY_SIZE %% 8 == 0 (vector gradient stores)");
    DC_REQUIRE(B > 0 && feats && gt, "null pointer argument / empty batch");
    if (int rc = reserve(B)) return rc;
    Bf16State &b = *bf;
    const int T = cfg.padding, U = cfg.units, V = cfg.vocab;
    const int K1 = b.Epad + U, Kin = cfg.pool * cfg.pool * cfg.channels;
    if (int rc = train_reserve(*this, B, T)) return rc;
    TrainState &t = *b.train;
    const long long R = (long long)T * B;
    const bool train_head = kind != DC_FEATS_HEAD_F32;

    // ---------------- forward ----------------
    time_major_tokens_kernel<<<ceil_div((int)R, 256), 256, 0, s>>>(gt, targets, B, T, V, t.tok_tm, t.tgt_tm);
    DC_CHECK_LAUNCH();
    const __nv_bfloat16 *x0 = nullptr;
    if (kind == DC_FEATS_ROI_BF16) {
        x0 = reinterpret_cast<const __nv_bfloat16 *>(feats);
    } else if (kind == DC_FEATS_ROI_F32) {
        if (int rc = f32_to_bf16((const float *)feats, t.x0, (long long)B * Kin, s)) return rc;
        x0 = t.x0;
    }
    if (int rc = head(train_head ? (const void *)x0 : feats, train_head ? DC_FEATS_ROI_BF16 : kind, B, ws.F, s)) return rc;
    if (int rc = v1_hoist_bf16(B, s)) return rc;                       // ws.g1f (gate-interleaved, + b1), ws.d1f (+ bd1)

    // recurrent dropout: K-stacked operands (see dropout_* above); off in parity mode and for predict()
    const bool dropout = train_opts && train_opts->recurrent_dropout > 0.f;
    DropoutCtx dctx{};
    if (dropout) {
        if (int rc = dropout_reserve(*this, B, T)) return rc;
        if (int rc = refresh_train_weights(s)) return rc;                 // w2cat_k feeds the layer-2 backward operand
        if (int rc = dropout_build_weights(*this, s)) return rc;
        dctx = dropout_ctx(train_opts);
    }
    const int K1x = dropout ? b.Epad + 4 * U : K1, K2x = dropout ? 5 * U : 2 * U;
    __nv_bfloat16 *X1b = dropout ? t.do_X1 : t.X1, *X2b = dropout ? t.do_X2 : t.X2;
    const __nv_bfloat16 *W1op = dropout ? t.do_w1 : b.w1cat, *W2op = dropout ? t.do_w2 : b.w2cat;
    // slot 0 of the state buffers = zero state; embedding rows of every (t, b) gathered at once
    DC_CHECK_CUDA(cudaMemsetAsync(X1b, 0, 2 * (size_t)B * K1x, s));
    DC_CHECK_CUDA(cudaMemsetAsync(X2b, 0, 2 * (size_t)B * K2x, s));
    if (dropout) {
        DC_CHECK_CUDA(cudaMemsetAsync(t.do_H1, 0, 2 * (size_t)B * U, s));
        DC_CHECK_CUDA(cudaMemsetAsync(t.do_H2, 0, 2 * (size_t)B * U, s));
    }
    DC_CHECK_CUDA(cudaMemsetAsync(t.c1, 0, 4 * (size_t)B * U, s));
    DC_CHECK_CUDA(cudaMemsetAsync(t.c2, 0, 4 * (size_t)B * U, s));
    gather_embedding_rows_kernel<<<(unsigned)ceil_div<long long>(R * 32, 256), 256, 0, s>>>(
        reinterpret_cast<const uint4 *>(b.emb), b.Epad / 8, t.tok_tm, R, reinterpret_cast<uint4 *>(X1b), K1x / 8);
    DC_CHECK_LAUNCH();
    // Wavefront: LSTM1 of step t+1 needs only LSTM1 of step t (the tokens are given), LSTM2 of step t needs LSTM1 of
    // step t and LSTM2 of step t-1.  Chain A (LSTM1) runs on `s`, chain B (LSTM2) on the side stream one step behind;
    // their persistent GEMMs interleave at CTA granularity and fill each other's tails (a single chain of these
    // 1.7-wave launches leaves most SMs idle most of the time).
    static const bool env_no_wave = getenv("DCAP_NO_WAVEFRONT") != nullptr;
    const bool no_wave = env_no_wave || dropout;
    cudaStream_t s2 = no_wave ? s : t.side;
    const unsigned ex_grid = (unsigned)ceil_div<long long>((long long)B * U, 256);
    for (int st = 0; st < T; ++st) {
        __nv_bfloat16 *x1 = X1b + (size_t)st * B * K1x, *x1n = x1 + (size_t)B * K1x;
        __nv_bfloat16 *x2 = X2b + (size_t)st * B * K2x, *x2n = x2 + (size_t)B * K2x;
        __nv_bfloat16 *h1p = dropout ? t.do_H1 + (size_t)st * B * U : x1 + b.Epad, *h1n = dropout ? h1p + (size_t)B * U : x1n + b.Epad;
        __nv_bfloat16 *h2p = dropout ? t.do_H2 + (size_t)st * B * U : x2 + U, *h2n = dropout ? h2p + (size_t)B * U : x2n + U;
        const int ldh1 = dropout ? U : K1, ldh2 = dropout ? U : 2 * U;
        const int32_t *tok = t.tok_tm + (size_t)st * B;
        TcEpilogue c1;
        c1.addend = ws.g1f; c1.ld_addend = 4 * U; c1.cell_units = U; c1.cell_tok = tok;
        c1.cell_c = t.c1 + (size_t)st * B * U; c1.cell_c_out = t.c1 + (size_t)(st + 1) * B * U;
        c1.cell_h_prev = h1p; c1.ld_h_prev = ldh1;
        c1.cell_h_a = h1n; c1.ld_h_a = ldh1;
        c1.cell_h_b = x2; c1.ld_h_b = K2x;
        c1.cell_gates_out = t.gates1 + (size_t)st * B * 4 * U; c1.ld_gates_out = 4 * U;
        if (int rc = gemm_bf16_tc(tc_op(x1, K1x), tc_op(W1op, K1x), c1, B, 4 * U, K1x, kEpiCell, s)) return rc;
        if (dropout) {
            dropout_expand_kernel<<<ex_grid, 256, 0, s>>>(h1n, U, B, U, 1, dctx, x1n + b.Epad, K1x);
            DC_CHECK_LAUNCH();
        }
        if (!no_wave) {
            DC_CHECK_CUDA(cudaEventRecord(t.step_ev[st], s));
            DC_CHECK_CUDA(cudaStreamWaitEvent(s2, t.step_ev[st], 0));
        }
        TcEpilogue c2;
        c2.bias = b.b2_i; c2.cell_units = U; c2.cell_tok = tok;
        c2.cell_c = t.c2 + (size_t)st * B * U; c2.cell_c_out = t.c2 + (size_t)(st + 1) * B * U;
        c2.cell_h_prev = h2p; c2.ld_h_prev = ldh2;
        c2.cell_h_a = h2n; c2.ld_h_a = ldh2;
        c2.cell_gates_out = t.gates2 + (size_t)st * B * 4 * U; c2.ld_gates_out = 4 * U;
        if (int rc = gemm_bf16_tc(tc_op(x2, K2x), tc_op(W2op, K2x), c2, B, 4 * U, K2x, kEpiCell, s2)) return rc;
        if (dropout) {
            dropout_expand_kernel<<<ex_grid, 256, 0, s2>>>(h2n, U, B, U, 2, dctx, x2n + U, K2x);
            DC_CHECK_LAUNCH();
        }
    }
    if (!no_wave) {
        DC_CHECK_CUDA(cudaEventRecord(t.join_ev, s2));
        DC_CHECK_CUDA(cudaStreamWaitEvent(s, t.join_ev, 0));
    }
    // h2_t of row (t, b): slot t+1 of the state buffer -- one strided view over all T*B rows
    const __nv_bfloat16 *h2_all = dropout ? t.do_H2 + (size_t)B * U : t.X2 + (size_t)B * 2 * U + U;
    const int ld_h2 = dropout ? U : 2 * U;
    {
        TcEpilogue e;
        e.addend = ws.d1f; e.ld_addend = kDense; e.addend_mod = B; e.relu = 1; e.out_bf16 = t.d_all; e.ld_bf16 = kDense;
        if (int rc = gemm_bf16_tc(tc_op(h2_all, ld_h2), tc_op(b.wd1h, U), e, (int)R, kDense, U, kEpiStore, s)) return rc;
        TcEpilogue v;
        v.bias = W("imgcap_lstm_d2/bias");
        static const bool f32_logits_env = getenv("DCAP_TRAIN_F32_LOGITS") != nullptr;
        t.logits_in_dz = want_bf16_logits && !f32_logits_env;
        if (t.logits_in_dz) {
            v.out_bf16 = t.dz; v.ld_bf16 = V;                              // turned into dlogits in place by softmax_xent_kernel
        } else {
            if (!t.logits) {
                const size_t Bp = round_up(t.B, 128);
                const cudaError_t e = cudaMalloc((void **)&t.logits, 4 * (size_t)t.T * Bp * V);
                if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "fp32 logits allocation failed: %s", cudaGetErrorString(e));
                t.owned.push_back(t.logits);
            }
            v.out_f32 = t.logits; v.ld_f32 = V;
        }
        if (int rc = gemm_bf16_tc(tc_op(t.d_all, kDense), tc_op(b.wd2, kDense), v, (int)R, V, kDense, kEpiStore, s)) return rc;
    }
    t.x0_used = x0;
    return DC_OK;
}

int Decoder::train_step(const void *feats, int kind, int B, const int32_t *gt, const int32_t *targets, float inv_count,
                        float *loss, cudaStream_t s) {
    DC_REQUIRE(loss, "null loss pointer");
    PdlScope pdl;
    train_logits_bf16 = true;
    const int frc = train_forward(feats, kind, B, gt, targets, s);
    train_logits_bf16 = false;
    if (frc) return frc;
    if (int rc = ensure_grads()) return rc;
    Bf16State &b = *bf;
    if (int rc = refresh_train_weights(s)) return rc;
    const int T = cfg.padding, U = cfg.units, F = cfg.feat, E = cfg.embed, V = cfg.vocab;
    const int K1 = b.Epad + U, Kin = cfg.pool * cfg.pool * cfg.channels;
    TrainState &t = *b.train;
    const long long R = (long long)T * B;
    if (inv_count <= 0.f) inv_count = 1.0f / (float)R;
    const bool train_head = kind != DC_FEATS_HEAD_F32;
    const __nv_bfloat16 *x0 = t.x0_used;
    const bool dropout = train_opts && train_opts->recurrent_dropout > 0.f;
    const DropoutCtx dctx = dropout ? dropout_ctx(train_opts) : DropoutCtx{};
    const __nv_bfloat16 *h2_all = dropout ? t.do_H2 + (size_t)B * U : t.X2 + (size_t)B * 2 * U + U;
    const int ld_h2 = dropout ? U : 2 * U;
    const int K1x = dropout ? b.Epad + 4 * U : K1, K2x = dropout ? 5 * U : 2 * U;
    const __nv_bfloat16 *X1b = dropout ? t.do_X1 : t.X1, *X2b = dropout ? t.do_X2 : t.X2;
    auto G = [&](const char *name) { return grads + find(name)->offset; };
    DC_CHECK_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)n_train, s));

    // softmax + cross-entropy + dlogits: one CTA per row keeps ~10 rows in flight per SM, which is what makes
    // this kernel run at HBM speed (a variant that kept rows in registers and accumulated the vocabulary-bias
    // gradient in place was latency-bound on its per-row barrier and measured 0.7 ms SLOWER per step)
    // The fused form (softmax_xent_colsum_kernel: shared-memory ring fed by bulk copies, persistent CTAs, the
    // vocabulary-bias gradient accumulated in registers on the way) replaces this kernel AND colsum(dz) below whenever
    // the logits sit in dz; DCAP_XENT_FUSED=0 keeps the two-kernel form (A/B, tests).
    const char *xf = getenv("DCAP_XENT_FUSED");
    const bool fused_xent = t.logits_in_dz && !(xf && atoi(xf) == 0) && xent_colsum_supported(t.dz, V, V);
    if (fused_xent) {
        if (int rc = softmax_xent_colsum(t.dz, V, V, R, t.tgt_tm, inv_count, t.rowloss, G("imgcap_lstm_d2/bias"), s)) return rc;
    } else {
        if (t.logits_in_dz)
            softmax_xent_kernel<__nv_bfloat16><<<(unsigned)R, 256, 0, s>>>(t.dz, V, V, t.tgt_tm, inv_count, t.dz, V, t.rowloss);
        else
            softmax_xent_kernel<float><<<(unsigned)R, 256, 0, s>>>(t.logits, V, V, t.tgt_tm, inv_count, t.dz, V, t.rowloss);
        DC_CHECK_LAUNCH();
    }
    reduce_loss_kernel<<<1, 1024, 0, s>>>(t.rowloss, R, inv_count, loss);
    DC_CHECK_LAUNCH();

    // ---------------- backward: time-batched part ----------------
    auto wgrad = [&](const __nv_bfloat16 *X, long long ldx, int M, const __nv_bfloat16 *dY, long long ldy, int N,
                     long long rows, float *out, long long ldo) {
        TcEpilogue e;
        e.out_f32 = out; e.ld_f32 = ldo; e.atomic = 1;
        return gemm_bf16_tc(tc_op(X, ldx, true), tc_op(dY, ldy, true), e, M, N, (int)rows, kEpiStore, s, 0);
    };
    // dense2: dWd2 = d^T dz, dbd2 = colsum(dz), dd = (dz Wd2^T) * [d > 0]
    if (int rc = wgrad(t.d_all, kDense, kDense, t.dz, V, V, R, G("imgcap_lstm_d2/kernel"), V)) return rc;
    if (!fused_xent)
        if (int rc = colsum(t.dz, R, V, V, G("imgcap_lstm_d2/bias"), s)) return rc;
    DC_CHECK_CUDA(cudaEventRecord(t.bucket_ev[0], s));                 // bucket 0: vocabulary projection
    {
        TcEpilogue e;
        e.mask_src = t.d_all; e.ld_mask = kDense; e.out_bf16 = t.dd; e.ld_bf16 = kDense;
        if (int rc = gemm_bf16_tc(tc_op(t.dz, V), tc_op(b.wd2_k, V), e, (int)R, kDense, V, kEpiStore, s)) return rc;
    }
    time_sum_kernel<<<(unsigned)ceil_div<long long>((long long)B * kDense, 256), 256, 0, s>>>(t.dd, B, T, kDense, t.ddsum,
                                                                                             t.ddsum_b);
    DC_CHECK_LAUNCH();
    // dense1: rows [0,U) of the kernel see h2_t, rows [U, U+F) the (time-constant) RoI feature
    if (int rc = wgrad(h2_all, ld_h2, U, t.dd, kDense, kDense, R, G("imgcap_lstm_d1/kernel"), kDense)) return rc;
    if (int rc = wgrad(b.Fb, F, F, t.ddsum_b, kDense, kDense, B, G("imgcap_lstm_d1/kernel") + (size_t)U * kDense, kDense)) return rc;
    if (int rc = colsum(t.ddsum, B, kDense, kDense, G("imgcap_lstm_d1/bias"), s)) return rc;
    DC_CHECK_CUDA(cudaEventRecord(t.bucket_ev[1], s));                 // bucket 1: dense1
    {
        TcEpilogue e;
        e.out_f32 = t.dh2d; e.ld_f32 = U;
        if (int rc = gemm_bf16_tc(tc_op(t.dd, kDense), tc_op(b.wd1h_k, kDense), e, (int)R, U, kDense, kEpiStore, s)) return rc;
    }

    // ---------------- backward: BPTT ----------------
    // Layer 2's backward chain never needs layer 1 (gradients only flow downwards), so it runs on the side stream
    // and hands dL/dh1_t (the first U columns of its per-step data gradient, kept for every t) to layer 1's chain
    // on `s` through one event per step: the two chains of short kernels overlap.
    for (float *p : {t.carry1, t.carry2, t.dc1, t.dc2})
        DC_CHECK_CUDA(cudaMemsetAsync(p, 0, 4 * (size_t)B * U, s));
    DC_CHECK_CUDA(cudaMemsetAsync(t.dz1sum, 0, 4 * (size_t)B * 4 * U, s));
    static const bool env_no_wave = getenv("DCAP_NO_WAVEFRONT") != nullptr;
    const bool no_wave = env_no_wave || dropout;
    cudaStream_t s2 = no_wave ? s : t.side;
    const unsigned fold_grid = (unsigned)ceil_div<long long>((long long)B * U, 256);
    if (!no_wave) {
        DC_CHECK_CUDA(cudaEventRecord(t.fork_ev, s));                  // dh2d, zeroed carries are ready
        DC_CHECK_CUDA(cudaStreamWaitEvent(s2, t.fork_ev, 0));
    }
    const unsigned cb_grid = (unsigned)ceil_div<long long>((long long)B * (U / 4), 256);
    for (int st = T - 1; st >= 0; --st) {                              // chain B: layer 2
        const int32_t *tok = t.tok_tm + (size_t)st * B;
        const bool last = st == T - 1;
        __nv_bfloat16 *dz2 = t.dz2_all + (size_t)st * B * 4 * U;
        float *dx = t.dxh2 + (size_t)st * B * 2 * U;                   // [dh1_t | dh2_{t-1}] of this step
        DC_CHECK_CUDA(launch_pdl(lstm_cell_bwd_kernel, dim3(cb_grid), dim3(256), 0, s2, B, U,
                                 (const __nv_bfloat16 *)(t.gates2 + (size_t)st * B * 4 * U), (const float *)(t.c2 + (size_t)st * B * U),
                                 (const float *)(t.c2 + (size_t)(st + 1) * B * U), tok, (const float *)(t.dh2d + (size_t)st * B * U), U,
                                 (const float *)(last ? nullptr : dx + (size_t)B * 2 * U + U), 2 * U, t.carry2, t.dc2, dz2,
                                 (float *)nullptr));
        if (!dropout) {   // [dh1_t | dh2_{t-1}] = dz2 [W2 ; U2]^T
            TcEpilogue e;
            e.out_f32 = dx; e.ld_f32 = 2 * U;
            if (int rc = gemm_bf16_tc(tc_op(dz2, 4 * U), tc_op(b.w2cat_k, 4 * U), e, B, 2 * U, 4 * U, kEpiStore, s2)) return rc;
        } else {          // [dh1_t | d(h2*m_i) | d(h2*m_f) | d(h2*m_c) | d(h2*m_o)], then dh2_{t-1} = sum_g m_g * d(h2*m_g)
            TcEpilogue e;
            e.out_f32 = t.do_dx; e.ld_f32 = 5 * U;
            if (int rc = gemm_bf16_tc(tc_op(dz2, 4 * U), tc_op(t.do_w2k, 4 * U), e, B, 5 * U, 4 * U, kEpiStore, s2)) return rc;
            dropout_fold_kernel<<<fold_grid, 256, 0, s2>>>(t.do_dx + U, 5 * U, B, U, 2, dctx, dx + U, 2 * U, t.do_dx, U, dx);
            DC_CHECK_LAUNCH();
        }
        if (!no_wave) DC_CHECK_CUDA(cudaEventRecord(t.step_ev[st], s2));
    }
    for (int st = T - 1; st >= 0; --st) {                              // chain A: layer 1
        const int32_t *tok = t.tok_tm + (size_t)st * B;
        const bool last = st == T - 1;
        __nv_bfloat16 *dz1 = t.dz1_all + (size_t)st * B * 4 * U;
        const float *dx = t.dxh2 + (size_t)st * B * 2 * U;
        if (!no_wave) DC_CHECK_CUDA(cudaStreamWaitEvent(s, t.step_ev[st], 0));
        DC_CHECK_CUDA(launch_pdl(lstm_cell_bwd_kernel, dim3(cb_grid), dim3(256), 0, s, B, U,
                                 (const __nv_bfloat16 *)(t.gates1 + (size_t)st * B * 4 * U), (const float *)(t.c1 + (size_t)st * B * U),
                                 (const float *)(t.c1 + (size_t)(st + 1) * B * U), tok, dx, 2 * U,
                                 (const float *)(last ? nullptr : t.dh1p), U, t.carry1, t.dc1, dz1, t.dz1sum));
        if (st > 0 && !dropout) {   // dh1_{t-1} = dz1 U1^T
            TcEpilogue e;
            e.out_f32 = t.dh1p; e.ld_f32 = U;
            if (int rc = gemm_bf16_tc(tc_op(dz1, 4 * U), tc_op(b.u1_k, 4 * U), e, B, U, 4 * U, kEpiStore, s)) return rc;
        } else if (st > 0) {
            TcEpilogue e;
            e.out_f32 = t.do_dx; e.ld_f32 = 4 * U;
            if (int rc = gemm_bf16_tc(tc_op(dz1, 4 * U), tc_op(t.do_u1k, 4 * U), e, B, 4 * U, 4 * U, kEpiStore, s)) return rc;
            dropout_fold_kernel<<<fold_grid, 256, 0, s>>>(t.do_dx, 4 * U, B, U, 1, dctx, t.dh1p, U, nullptr, 0, nullptr);
            DC_CHECK_LAUNCH();
        }
    }
    f32_to_bf16_rows_kernel<<<(unsigned)ceil_div<long long>((long long)B * 4 * U, 256), 256, 0, s>>>(t.dz1sum, (long long)B * 4 * U,
                                                                                                    t.dz1sum_b);
    DC_CHECK_LAUNCH();

    // ---------------- backward: LSTM weight gradients (one GEMM over all T*B rows each) ----------------
    // [dW2 ; dU2] = [h1_t | h2_{t-1}]^T dz2  (kernel and recurrent_kernel gradients are adjacent)
    if (!dropout) {
        if (int rc = wgrad(t.X2, 2 * U, 2 * U, t.dz2_all, 4 * U, 4 * U, R, G("imgcap_lstm2/kernel"), 4 * U)) return rc;
    } else {              // dW2 = h1^T dz2; dU2[:, gate g] = (h2*m_g)^T dz2[:, gate g]
        if (int rc = wgrad(X2b, K2x, U, t.dz2_all, 4 * U, 4 * U, R, G("imgcap_lstm2/kernel"), 4 * U)) return rc;
        for (int g = 0; g < 4; ++g)
            if (int rc = wgrad(X2b + U + (size_t)g * U, K2x, U, t.dz2_all + (size_t)g * U, 4 * U, U, R,
                               G("imgcap_lstm2/recurrent_kernel") + (size_t)g * U, 4 * U)) return rc;
    }
    if (int rc = colsum(t.dz2_all, R, 4 * U, 4 * U, G("imgcap_lstm2/bias"), s)) return rc;
    // dW1[:E] = emb_t^T dz1 ; dU1 = h1_{t-1}^T dz1 ; dW1[E:] = f^T sum_t dz1
    if (int rc = wgrad(X1b, K1x, E, t.dz1_all, 4 * U, 4 * U, R, G("imgcap_lstm1/kernel"), 4 * U)) return rc;
    if (!dropout) {
        if (int rc = wgrad(t.X1 + b.Epad, K1, U, t.dz1_all, 4 * U, 4 * U, R, G("imgcap_lstm1/recurrent_kernel"), 4 * U)) return rc;
    } else {
        for (int g = 0; g < 4; ++g)
            if (int rc = wgrad(X1b + b.Epad + (size_t)g * U, K1x, U, t.dz1_all + (size_t)g * U, 4 * U, U, R,
                               G("imgcap_lstm1/recurrent_kernel") + (size_t)g * U, 4 * U)) return rc;
    }
    if (int rc = wgrad(b.Fb, F, F, t.dz1sum_b, 4 * U, 4 * U, B, G("imgcap_lstm1/kernel") + (size_t)E * 4 * U, 4 * U)) return rc;
    if (int rc = colsum(t.dz1sum, B, 4 * U, 4 * U, G("imgcap_lstm1/bias"), s)) return rc;
    DC_CHECK_CUDA(cudaEventRecord(t.bucket_ev[2], s));                 // bucket 2: both LSTMs

    if (!train_head) {
        DC_REQUIRE(!(train_opts && train_opts->d_feats), "d_feats needs RoI-feature input (the head must be part of the graph)");
        DC_CHECK_CUDA(cudaEventRecord(t.bucket_ev[3], s));             // head gradients stay zero
        return DC_OK;
    }
    // ---------------- backward: RoI head ----------------
    {   // dF = (sum_t dd) Wd1[U:]^T + (sum_t dz1) W1[E:]^T
        TcEpilogue e;
        e.out_f32 = t.dF; e.ld_f32 = F;
        if (int rc = gemm_bf16_tc(tc_op(t.ddsum_b, kDense), tc_op(b.wd1f_k, kDense), e, B, F, kDense, kEpiStore, s)) return rc;
        TcEpilogue e2;
        e2.addend = t.dF; e2.ld_addend = F; e2.out_f32 = t.dF; e2.ld_f32 = F;
        if (int rc = gemm_bf16_tc(tc_op(t.dz1sum_b, 4 * U), tc_op(b.w1f_k, 4 * U), e2, B, F, 4 * U, kEpiStore, s)) return rc;
    }
    const dim3 bn_grid(ceil_div(F, 32), ceil_div(B, 256)), bn_block(32, 8);
    bn_relu_bwd_kernel<<<bn_grid, bn_block, 0, s>>>(t.dF, b.Fb, B, F, bn_scale[1], W("mrcnn_class_bn2/gamma"),
                                                    W("mrcnn_class_bn2/beta"), 256, t.dzh2, G("mrcnn_class_bn2/gamma"),
                                                    G("mrcnn_class_bn2/beta"), G("mrcnn_class_conv2/bias"));
    DC_CHECK_LAUNCH();
    if (int rc = wgrad(b.a1, F, F, t.dzh2, F, F, B, G("mrcnn_class_conv2/kernel"), F)) return rc;
    {
        TcEpilogue e;
        e.out_f32 = t.da1; e.ld_f32 = F;
        if (int rc = gemm_bf16_tc(tc_op(t.dzh2, F), tc_op(b.wc2_k, F), e, B, F, F, kEpiStore, s)) return rc;
    }
    bn_relu_bwd_kernel<<<bn_grid, bn_block, 0, s>>>(t.da1, b.a1, B, F, bn_scale[0], W("mrcnn_class_bn1/gamma"),
                                                    W("mrcnn_class_bn1/beta"), 256, t.dzh1, G("mrcnn_class_bn1/gamma"),
                                                    G("mrcnn_class_bn1/beta"), G("mrcnn_class_conv1/bias"));
    DC_CHECK_LAUNCH();
    if (int rc = wgrad(x0, Kin, Kin, t.dzh1, F, F, B, G("mrcnn_class_conv1/kernel"), F)) return rc;
    DC_CHECK_CUDA(cudaEventRecord(t.bucket_ev[3], s));                 // bucket 3: RoI head
    if (train_opts && train_opts->d_feats) {
        // joint model (dense_img_cap/dense_model.py:738-755): dL/dX = dz1 K1^T, the conv7x7(valid) read every one
        // of the pool*pool*C inputs exactly once.  The Keras-layout kernel [Kin, F] is the K-major B operand.
        TcEpilogue e;
        e.out_f32 = train_opts->d_feats; e.ld_f32 = Kin;
        const __nv_bfloat16 *k1 = b.arena_k + find("mrcnn_class_conv1/kernel")->offset;
        if (int rc = gemm_bf16_tc(tc_op(t.dzh1, F), tc_op(k1, F), e, B, Kin, F, kEpiStore, s)) return rc;
    }
    return DC_OK;
}

// Gradient buckets in the order the backward pass completes them (reverse layer order); each is ONE
// contiguous range of the flat gradient buffer because the arena keeps declaration order.
int Decoder::grad_bucket(int i, int64_t *offset, int64_t *numel) {
    DC_REQUIRE(i >= 0 && i < 4 && offset && numel, "gradient bucket index outside [0,4)");
    DC_REQUIRE(cfg.arch == DC_ARCH_V1, "gradient buckets are defined for the v1 model");
    auto off = [&](const char *n) { return find(n)->offset; };
    const int64_t head0 = off("mrcnn_class_conv1/kernel"), lstm0 = off("imgcap_lstm1/kernel"),
                  d1 = off("imgcap_lstm_d1/kernel"), d2 = off("imgcap_lstm_d2/kernel");
    DC_REQUIRE(head0 == 0 && head0 < lstm0 && lstm0 < d1 && d1 < d2 && d2 < n_train, "unexpected arena order");
    const int64_t lo[4] = {d2, d1, lstm0, head0}, hi[4] = {n_train, d2, d1, lstm0};
    *offset = lo[i]; *numel = hi[i] - lo[i];
    return DC_OK;
}

int Decoder::wait_grad_bucket(int i, cudaStream_t waiter) {
    DC_REQUIRE(i >= 0 && i < 4, "gradient bucket index outside [0,4)");
    DC_REQUIRE(bf && bf->train && bf->train->bucket_ev[i], "no training step has run on this handle");
    DC_CHECK_CUDA(cudaStreamWaitEvent(waiter, bf->train->bucket_ev[i], 0));
    return DC_OK;
}

// probs[b, t, :] = softmax(logits[t*B + b, :]): the Keras output layout [B, P, V] of the training graph
__global__ void __launch_bounds__(256) softmax_rows_to_bpv_kernel(const float *__restrict__ logits, int B, int T, int V,
                                                                  float *__restrict__ probs) {
    const int r = blockIdx.x, t = r / B, b = r - t * B;
    const float *z = logits + (long long)r * V;
    float *p = probs + ((long long)b * T + t) * V;
    __shared__ float red[8];
    __shared__ float bc;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < V; j += 256) mx = fmaxf(mx, z[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) { float m = red[0]; for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]); bc = m; }
    __syncthreads();
    mx = bc;
    float sum = 0.f;
    for (int j = threadIdx.x; j < V; j += 256) sum += expf(z[j] - mx);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { float a = 0.f; for (int i = 0; i < 8; ++i) a += red[i]; bc = a; }
    __syncthreads();
    const float inv = 1.0f / bc;
    for (int j = threadIdx.x; j < V; j += 256) p[j] = expf(z[j] - mx) * inv;
}

int Decoder::teacher_forced_probs(const void *feats, int kind, int B, const int32_t *gt, float *probs, cudaStream_t s) {
    DC_REQUIRE(probs, "null pointer argument");
    if (int rc = train_forward(feats, kind, B, gt, nullptr, s)) return rc;
    softmax_rows_to_bpv_kernel<<<B * cfg.padding, 256, 0, s>>>(bf->train->logits, B, cfg.padding, cfg.vocab, probs);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// v2 inject model: one training step (text_generation_model_v2.py:140-166 build_model(inject=True), :263-267
// compile(Adam(amsgrad=True), categorical_crossentropy), :312 fit_generator on batches of 1024 (prefix, next word)
// pairs).  Trainable: lstm_1 (word LSTM 1024 over the embedded, pre-padded, masked prefix), imgcap_lstm (one
// step from the zero state on [head feature ; word vector]) and imgcap_d1; the RoI head (trainable=False) and
// the embedding are frozen -- their gradient slots stay zero, which Adam leaves untouched.
// ------------------------------------------------------------------------------------------------
static int train_reserve_v2(Decoder &D, int B, int L) {
    Bf16State &b = *D.bf;
    if (!b.train) b.train = new TrainState();
    TrainState &t = *b.train;
    if (t.v2_B >= B && t.v2_L >= L) return DC_OK;
    for (void *p : t.v2_owned) cudaFree(p);
    t.v2_owned.clear();
    const DcDecoderConfig &c = D.cfg;
    const size_t U = c.units, F = c.feat, V = c.vocab, Wu = c.word_units, Kw = b.Epad + Wu;
    const size_t Bp = round_up(B, 128), R = (size_t)L * Bp;
    auto A = [&](void **p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
        if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "training workspace allocation failed: %s", cudaGetErrorString(e));
        t.v2_owned.push_back(*p);
        return DC_OK;
    };
    int rc = 0;
    rc |= A((void **)&t.v2_tok_tm, 4 * R); rc |= A((void **)&t.v2_tgt_dummy, 4 * R); rc |= A((void **)&t.v2_ones, 4 * Bp);
    rc |= A((void **)&t.v2_X, 2 * (R + Bp) * Kw); rc |= A((void **)&t.v2_c, 4 * (R + Bp) * Wu);
    rc |= A((void **)&t.v2_gates_w, 2 * R * 4 * Wu); rc |= A((void **)&t.v2_gates_i, 2 * Bp * 4 * U);
    rc |= A((void **)&t.v2_c_img, 4 * Bp * U); rc |= A((void **)&t.v2_logits, 4 * Bp * V);
    rc |= A((void **)&t.v2_dz, 2 * Bp * V); rc |= A((void **)&t.v2_dzi, 2 * Bp * 4 * U); rc |= A((void **)&t.v2_dzw, 2 * R * 4 * Wu);
    rc |= A((void **)&t.v2_dh_img, 4 * Bp * U); rc |= A((void **)&t.v2_dxin, 4 * Bp * (F + Wu)); rc |= A((void **)&t.v2_dh_prev, 4 * Bp * Wu);
    rc |= A((void **)&t.v2_carry, 4 * Bp * Wu); rc |= A((void **)&t.v2_dc, 4 * Bp * Wu);
    rc |= A((void **)&t.v2_carry_i, 4 * Bp * U); rc |= A((void **)&t.v2_dc_i, 4 * Bp * U);
    rc |= A((void **)&t.v2_rowloss, 4 * Bp);
    if (rc) { t.v2_B = t.v2_L = 0; return rc; }
    t.v2_B = B; t.v2_L = L;
    return DC_OK;
}

int Decoder::train_step_v2(const void *feats, int kind, int B, const int32_t *words, int L, const int32_t *targets,
                           float inv_count, float *loss, cudaStream_t s) {
    PdlScope pdl;
    if (int rc = check_ready(B)) return rc;
    DC_REQUIRE(cfg.arch == DC_ARCH_V2_INJECT && cfg.dtype == DC_DTYPE_BF16, "v2 training is served by the bf16 v2 inject decoder");
    DC_REQUIRE(cfg.vocab % 8 == 0, "training needs VOCABULARY_SIZE %% 8 == 0 (vector gradient stores)");
    DC_REQUIRE(B > 0 && L > 0 && feats && words && targets && loss, "null pointer argument / empty batch");
    if (int rc = reserve(B)) return rc;
    if (int rc = ensure_grads()) return rc;
    Bf16State &b = *bf;
    if (int rc = train_reserve_v2(*this, B, L)) return rc;
    TrainState &t = *b.train;
    const int U = cfg.units, F = cfg.feat, V = cfg.vocab, Wu = cfg.word_units, E = cfg.embed, Kw = b.Epad + Wu;
    const long long R = (long long)L * B;
    if (inv_count <= 0.f) inv_count = 1.0f / (float)B;
    // bf16 mirror of the trainable arena: Keras [in, out] tensors = K-major B operands of the data-gradient GEMMs
    if (!b.arena_k)
        if (int rc = dev_alloc((void **)&b.arena_k, 2 * (size_t)n_train, owned)) return rc;
    if (!b.arena_k_valid) {
        if (int rc = f32_to_bf16(arena, b.arena_k, n_train, s)) return rc;
        b.arena_k_valid = true;
    }
    auto Kk = [&](const char *name) { return b.arena_k + find(name)->offset; };
    auto G = [&](const char *name) { return grads + find(name)->offset; };
    auto wgrad = [&](const __nv_bfloat16 *X, long long ldx, int M, const __nv_bfloat16 *dY, long long ldy, int N, long long rows,
                     float *out, long long ldo) {
        TcEpilogue e;
        e.out_f32 = out; e.ld_f32 = ldo; e.atomic = 1;
        return gemm_bf16_tc(tc_op(X, ldx, true), tc_op(dY, ldy, true), e, M, N, (int)rows, kEpiStore, s, 0);
    };

    // ---------------- forward ----------------
    if (int rc = v2_begin_bf16(feats, kind, B, s)) return rc;          // head -> xin[:, :F]; xin[:, F:] = 0
    time_major_tokens_kernel<<<ceil_div((int)R, 256), 256, 0, s>>>(words, nullptr, B, L, V, t.v2_tok_tm, t.v2_tgt_dummy);
    DC_CHECK_LAUNCH();
    if (int rc = fill_i32(t.v2_ones, B, 1, s)) return rc;
    DC_CHECK_CUDA(cudaMemsetAsync(t.v2_X, 0, 2 * (size_t)B * Kw, s));
    DC_CHECK_CUDA(cudaMemsetAsync(t.v2_c, 0, 4 * (size_t)B * Wu, s));
    gather_embedding_rows_kernel<<<(unsigned)ceil_div<long long>(R * 32, 256), 256, 0, s>>>(
        reinterpret_cast<const uint4 *>(b.emb), b.Epad / 8, t.v2_tok_tm, R, reinterpret_cast<uint4 *>(t.v2_X), Kw / 8);
    DC_CHECK_LAUNCH();
    for (int st = 0; st < L; ++st) {
        __nv_bfloat16 *x = t.v2_X + (size_t)st * B * Kw, *xn = x + (size_t)B * Kw;
        TcEpilogue c;
        c.bias = b.v2_bw; c.cell_units = Wu; c.cell_tok = t.v2_tok_tm + (size_t)st * B;
        c.cell_c = t.v2_c + (size_t)st * B * Wu; c.cell_c_out = t.v2_c + (size_t)(st + 1) * B * Wu;
        c.cell_h_prev = x + b.Epad; c.ld_h_prev = Kw;
        c.cell_h_a = xn + b.Epad; c.ld_h_a = Kw;
        c.cell_h_b = b.v2_xin + F; c.ld_h_b = F + Wu;                  // the last step leaves the word vector in xin
        c.cell_gates_out = t.v2_gates_w + (size_t)st * B * 4 * Wu; c.ld_gates_out = 4 * Wu;
        if (int rc = gemm_bf16_tc(tc_op(x, Kw), tc_op(b.v2_w1cat, Kw), c, B, 4 * Wu, Kw, kEpiCell, s)) return rc;
    }
    {
        TcEpilogue c;
        c.bias = b.v2_bimg; c.cell_c = b.v2_czero; c.cell_c_out = t.v2_c_img; c.cell_units = U;
        c.cell_h_a = b.v2_hb; c.ld_h_a = U;
        c.cell_gates_out = t.v2_gates_i; c.ld_gates_out = 4 * U;
        if (int rc = gemm_bf16_tc(tc_op(b.v2_xin, F + Wu), tc_op(b.v2_wimg, F + Wu), c, B, 4 * U, F + Wu, kEpiCell, s)) return rc;
        TcEpilogue e;
        e.bias = W("imgcap_d1/bias"); e.out_f32 = t.v2_logits; e.ld_f32 = V;
        if (int rc = gemm_bf16_tc(tc_op(b.v2_hb, U), tc_op(b.v2_wd, U), e, B, V, U, kEpiStore, s)) return rc;
    }
    // ---------------- loss ----------------
    DC_CHECK_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)n_train, s));
    softmax_xent_kernel<float><<<(unsigned)B, 256, 0, s>>>(t.v2_logits, V, V, targets, inv_count, t.v2_dz, V, t.v2_rowloss);
    DC_CHECK_LAUNCH();
    reduce_loss_kernel<<<1, 1024, 0, s>>>(t.v2_rowloss, B, inv_count, loss);
    DC_CHECK_LAUNCH();
    // ---------------- backward ----------------
    if (int rc = wgrad(b.v2_hb, U, U, t.v2_dz, V, V, B, G("imgcap_d1/kernel"), V)) return rc;
    if (int rc = colsum(t.v2_dz, B, V, V, G("imgcap_d1/bias"), s)) return rc;
    {
        TcEpilogue e;
        e.out_f32 = t.v2_dh_img; e.ld_f32 = U;
        if (int rc = gemm_bf16_tc(tc_op(t.v2_dz, V), tc_op(Kk("imgcap_d1/kernel"), V), e, B, U, V, kEpiStore, s)) return rc;
    }
    for (float *p : {t.v2_carry_i, t.v2_dc_i}) DC_CHECK_CUDA(cudaMemsetAsync(p, 0, 4 * (size_t)B * U, s));
    for (float *p : {t.v2_carry, t.v2_dc}) DC_CHECK_CUDA(cudaMemsetAsync(p, 0, 4 * (size_t)B * Wu, s));
    DC_CHECK_CUDA(launch_pdl(lstm_cell_bwd_kernel, dim3((unsigned)ceil_div<long long>((long long)B * (U / 4), 256)), dim3(256), 0, s, B, U,
                             (const __nv_bfloat16 *)t.v2_gates_i, (const float *)b.v2_czero, (const float *)t.v2_c_img,
                             (const int32_t *)t.v2_ones, (const float *)t.v2_dh_img, U, (const float *)nullptr, 0, t.v2_carry_i,
                             t.v2_dc_i, t.v2_dzi, (float *)nullptr));
    if (int rc = wgrad(b.v2_xin, F + Wu, F + Wu, t.v2_dzi, 4 * U, 4 * U, B, G("imgcap_lstm/kernel"), 4 * U)) return rc;
    if (int rc = colsum(t.v2_dzi, B, 4 * U, 4 * U, G("imgcap_lstm/bias"), s)) return rc;
    {   // d[head ; word vector] = dz_img K^T; the recurrent kernel saw the zero state: no gradient
        TcEpilogue e;
        e.out_f32 = t.v2_dxin; e.ld_f32 = F + Wu;
        if (int rc = gemm_bf16_tc(tc_op(t.v2_dzi, 4 * U), tc_op(Kk("imgcap_lstm/kernel"), 4 * U), e, B, F + Wu, 4 * U, kEpiStore, s)) return rc;
    }
    const unsigned cb_grid = (unsigned)ceil_div<long long>((long long)B * (Wu / 4), 256);
    for (int st = L - 1; st >= 0; --st) {
        const bool last = st == L - 1;
        __nv_bfloat16 *dzw = t.v2_dzw + (size_t)st * B * 4 * Wu;
        DC_CHECK_CUDA(launch_pdl(lstm_cell_bwd_kernel, dim3(cb_grid), dim3(256), 0, s, B, Wu,
                                 (const __nv_bfloat16 *)(t.v2_gates_w + (size_t)st * B * 4 * Wu), (const float *)(t.v2_c + (size_t)st * B * Wu),
                                 (const float *)(t.v2_c + (size_t)(st + 1) * B * Wu), (const int32_t *)(t.v2_tok_tm + (size_t)st * B),
                                 (const float *)(last ? t.v2_dxin + F : t.v2_dh_prev), last ? F + Wu : Wu, (const float *)nullptr, 0,
                                 t.v2_carry, t.v2_dc, dzw, (float *)nullptr));
        if (st > 0) {
            TcEpilogue e;
            e.out_f32 = t.v2_dh_prev; e.ld_f32 = Wu;
            if (int rc = gemm_bf16_tc(tc_op(dzw, 4 * Wu), tc_op(Kk("lstm_1/recurrent_kernel"), 4 * Wu), e, B, Wu, 4 * Wu, kEpiStore, s)) return rc;
        }
    }
    if (int rc = wgrad(t.v2_X, Kw, E, t.v2_dzw, 4 * Wu, 4 * Wu, R, G("lstm_1/kernel"), 4 * Wu)) return rc;
    if (int rc = wgrad(t.v2_X + b.Epad, Kw, Wu, t.v2_dzw, 4 * Wu, 4 * Wu, R, G("lstm_1/recurrent_kernel"), 4 * Wu)) return rc;
    if (int rc = colsum(t.v2_dzw, R, 4 * Wu, 4 * Wu, G("lstm_1/bias"), s)) return rc;
    return DC_OK;
}

// The update on [offset, offset + numel) of the flat buffers; refresh: re-derive the operand copies afterwards (the
// whole-buffer step); a sharded optimiser (parallel.py, ZeRO-1 style) steps its own range without it, exchanges the
// updated ranges and calls params_updated().
int Decoder::adam_step_range(float lr, float beta1, float beta2, float eps, int amsgrad, long long t, float grad_scale,
                             long long offset, long long numel, bool refresh, cudaStream_t s) {
    DC_REQUIRE(grads, "dc_adam_step before any dc_decoder_train_step");
    DC_REQUIRE(t >= 1, "iteration count starts at 1");
    DC_REQUIRE(offset >= 0 && numel >= 0 && offset + numel <= n_train && offset % 4 == 0 && numel % 4 == 0,
               "optimiser range [%lld, %lld) outside the %lld trainable parameters or not a multiple of 4", offset, offset + numel,
               (long long)n_train);
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
    if (numel > 0) {
        adam_amsgrad_kernel<<<(unsigned)ceil_div<long long>(numel / 4, 256), 256, 0, s>>>(
            arena + offset, grads + offset, adam_m + offset, adam_v + offset, adam_vhat + offset, numel, (float)lr_t, beta1, beta2,
            eps, amsgrad, grad_scale, (bf && bf->arena_k) ? bf->arena_k + offset : nullptr);
        DC_CHECK_LAUNCH();
    }
    if (!refresh) {
        if (bf) bf->arena_k_valid = false;                 // only this range of the bf16 mirror is current
        return DC_OK;
    }
    if (bf && bf->arena_k) bf->arena_k_valid = true;
    // folded BN + bf16 operand copies (forward and backward layouts) are rebuilt IN PLACE, so captured
    // inference graphs stay valid
    return refresh_derived(s);
}

int Decoder::adam_step(float lr, float beta1, float beta2, float eps, int amsgrad, long long t, float grad_scale,
                       cudaStream_t s) {
    return adam_step_range(lr, beta1, beta2, eps, amsgrad, t, grad_scale, 0, n_train, true, s);
}

// The flat parameter buffer was written from outside (all-gather of the ranks' updated ranges): bf16 mirror of the arena
// and every derived operand copy again.
int Decoder::params_updated(cudaStream_t s) {
    if (bf) {
        bf->arena_k_valid = false;
        if (bf->arena_k)
            if (int rc = refresh_train_weights(s)) return rc;
    }
    return refresh_derived(s);
}

}  // namespace dcap

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace dcap;

extern "C" int dc_decoder_train_step(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *gt,
                                     const int32_t *targets, float inv_count, float *loss, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.train_step(feats, feats_kind, B, gt, targets, inv_count, loss, (cudaStream_t)stream);
}

extern "C" int dc_decoder_train_step_ex(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *gt,
                                        const int32_t *targets, float inv_count, float *loss, const DcTrainOptions *opts,
                                        void *stream) {
    DC_REQUIRE(dec, "null decoder");
    if (opts) {
        DC_REQUIRE(opts->recurrent_dropout >= 0.f && opts->recurrent_dropout < 1.f, "recurrent_dropout must be in [0, 1)");
    }
    dec->impl.train_opts = opts;
    const int rc = dec->impl.train_step(feats, feats_kind, B, gt, targets, inv_count, loss, (cudaStream_t)stream);
    dec->impl.train_opts = nullptr;
    return rc;
}

extern "C" int dc_decoder_v2_train_step(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *words, int L,
                                       const int32_t *targets, float inv_count, float *loss, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.train_step_v2(feats, feats_kind, B, words, L, targets, inv_count, loss, (cudaStream_t)stream);
}

extern "C" int dc_decoder_teacher_forced(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *gt,
                                        float *probs, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.teacher_forced_probs(feats, feats_kind, B, gt, probs, (cudaStream_t)stream);
}

extern "C" int dc_adam_step(DcDecoder *dec, float lr, float beta1, float beta2, float epsilon, int amsgrad,
                            int64_t iteration, float grad_scale, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    if (int rc = dec->impl.check_ready(0)) return rc;
    return dec->impl.adam_step(lr, beta1, beta2, epsilon, amsgrad, iteration, grad_scale, (cudaStream_t)stream);
}

extern "C" int dc_adam_step_range(DcDecoder *dec, float lr, float beta1, float beta2, float epsilon, int amsgrad,
                                  int64_t iteration, float grad_scale, int64_t offset, int64_t numel, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    if (int rc = dec->impl.check_ready(0)) return rc;
    return dec->impl.adam_step_range(lr, beta1, beta2, epsilon, amsgrad, iteration, grad_scale, offset, numel, false,
                                     (cudaStream_t)stream);
}

extern "C" int dc_decoder_params_updated(DcDecoder *dec, void *stream) {
    DC_REQUIRE(dec, "null decoder");
    if (int rc = dec->impl.check_ready(0)) return rc;
    return dec->impl.params_updated((cudaStream_t)stream);
}

extern "C" int dc_decoder_grad_buffer(DcDecoder *dec, float **ptr, int64_t *numel) {
    DC_REQUIRE(dec && ptr && numel, "null pointer argument");
    if (int rc = dec->impl.ensure_grads()) return rc;
    *ptr = dec->impl.grads;
    *numel = dec->impl.n_train;
    return DC_OK;
}

extern "C" int dc_decoder_grad_bucket(DcDecoder *dec, int index, int64_t *offset, int64_t *numel) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.grad_bucket(index, offset, numel);
}

extern "C" int dc_decoder_wait_grad_bucket(DcDecoder *dec, int index, void *waiting_stream) {
    DC_REQUIRE(dec, "null decoder");
    return dec->impl.wait_grad_bucket(index, (cudaStream_t)waiting_stream);
}

extern "C" int dc_decoder_param_buffer(DcDecoder *dec, float **ptr, int64_t *numel) {
    DC_REQUIRE(dec && ptr && numel, "null pointer argument");
    *ptr = dec->impl.arena;
    *numel = dec->impl.n_train;
    return DC_OK;
}

extern "C" int64_t dc_decoder_weight_offset(const DcDecoder *dec, int index) {
    if (!dec || index < 0 || index >= (int)dec->impl.weights.size()) return -1;
    const Weight &w = dec->impl.weights[index];
    return w.trainable ? w.offset : -1;
}

extern "C" int dc_decoder_get_grad(DcDecoder *dec, const char *name, float *host, int64_t numel) {
    DC_REQUIRE(dec && name && host, "null pointer argument");
    Weight *w = dec->impl.find(name);
    DC_REQUIRE(w != nullptr, "unknown weight '%s'", name);
    DC_REQUIRE(w->trainable, "weight '%s' is frozen (no gradient)", name);
    DC_REQUIRE(w->numel == numel, "weight '%s' holds %lld values, got %lld", name, (long long)w->numel, (long long)numel);
    DC_REQUIRE(dec->impl.grads, "no gradients yet: call dc_decoder_train_step first");
    DC_CHECK_CUDA(cudaMemcpy(host, dec->impl.grads + w->offset, sizeof(float) * (size_t)numel, cudaMemcpyDeviceToHost));
    return DC_OK;
}

// Element-wise / reduction kernels of the inject-LSTM decoder (fp32 state; shared by the fp32
// and bf16 GEMM paths).  Semantics follow the Keras layers used by
// /root/reference/dense_img_cap_separate_models/text_generation_model.py:130-156 (word model),
// :192-232 (greedy feedback) and /root/reference/image captioning/test.py:23-64 (beam search).
#include "decoder_kernels.cuh"

namespace dcap {

// ---------------------------------------------------------------------------------------------
// Embedding(mask_zero=True): row gather of the frozen table into the LSTM1 operand buffer.
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void embed_gather_kernel(const float *__restrict__ emb, const int32_t *__restrict__ tok,
                                    int rows, int E, int V, OutT *__restrict__ out, int ld) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    int t = tok[r];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    const float *src = emb + (long long)t * E;
    OutT *dst = out + (long long)r * ld;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        if constexpr (sizeof(OutT) == 2) dst[e] = __float2bfloat16_rn(__ldg(src + e));
        else dst[e] = __ldg(src + e);
    }
}

int embed_gather(const float *emb, const int32_t *tok, int rows, int E, int V, void *out, int ld,
                 bool bf16, cudaStream_t s) {
    if (rows <= 0) return DC_OK;
    if (bf16) embed_gather_kernel<__nv_bfloat16><<<rows, 128, 0, s>>>(emb, tok, rows, E, V, (__nv_bfloat16 *)out, ld);
    else embed_gather_kernel<float><<<rows, 128, 0, s>>>(emb, tok, rows, E, V, (float *)out, ld);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Keras LSTMCell (gate blocks i|f|c|o, hard_sigmoid recurrent activation, tanh) with the K.rnn
// mask rule: a row whose consumed token is 0 keeps (h, c).  h is written (fp32 and/or bf16) into
// the operand buffers of the GEMMs that consume it next.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float hard_sigmoid(float x) {
    const float y = __fadd_rn(__fmul_rn(0.2f, x), 0.5f);
    return fminf(fmaxf(y, 0.f), 1.f);
}

__global__ void lstm_cell_kernel(const CellArgs a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.rows * a.U) return;
    const int r = (int)(idx / a.U), u = (int)(idx - (long long)r * a.U);
    if (a.tok && a.tok[r] == 0) return;                       // masked timestep: carry state
    const float *g = a.gates + (long long)r * a.ld_gates;
    float zi = g[u], zf = g[a.U + u], zg = g[2 * a.U + u], zo = g[3 * a.U + u];
    if (a.gates2) {                                           // second addend (hoisted terms)
        const float *g2 = a.gates2 + (long long)r * a.ld_gates2;
        zi += g2[u]; zf += g2[a.U + u]; zg += g2[2 * a.U + u]; zo += g2[3 * a.U + u];
    }
    const float i = hard_sigmoid(zi), f = hard_sigmoid(zf), gg = tanhf(zg), o = hard_sigmoid(zo);
    const float c_old = a.c_in ? a.c_in[(long long)r * a.U + u] : 0.f;
    const float c = __fadd_rn(__fmul_rn(f, c_old), __fmul_rn(i, gg));
    const float tc = tanhf(c);
    const float h = __fmul_rn(o, tc);
    a.c_out[(long long)r * a.U + u] = c;
    if (a.h_f32_a) a.h_f32_a[(long long)r * a.ld_a + u] = h;
    if (a.h_f32_b) a.h_f32_b[(long long)r * a.ld_b + u] = h;
    if (a.h_bf16_a) a.h_bf16_a[(long long)r * a.ld_a + u] = __float2bfloat16_rn(h);
    if (a.h_bf16_b) a.h_bf16_b[(long long)r * a.ld_b + u] = __float2bfloat16_rn(h);
    if (a.save_act) {                                         // training: keep activations
        float *s = a.save_act + (long long)r * 5 * a.U;
        s[u] = i; s[a.U + u] = f; s[2 * a.U + u] = gg; s[3 * a.U + u] = o; s[4 * a.U + u] = tc;
    }
}

int lstm_cell(const CellArgs &a, cudaStream_t s) {
    if (a.rows <= 0) return DC_OK;
    const long long n = (long long)a.rows * a.U;
    lstm_cell_kernel<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, s>>>(a);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Row softmax + argmax (first index wins ties, as tf.argmax) over the vocabulary.
// One CTA per row.  probs (optional) are written at probs + row*ld_probs.
// ---------------------------------------------------------------------------------------------
struct MaxIdx { float v; int i; };

__device__ __forceinline__ MaxIdx better_first(MaxIdx a, MaxIdx b) {   // larger value, then smaller index
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads)
softmax_argmax_kernel(const float *__restrict__ logits, int ld, int V, float *__restrict__ probs,
                      long long ld_probs, int32_t *__restrict__ tok_out, int tok_stride,
                      int32_t *__restrict__ tok_cur, float *__restrict__ maxprob) {
    __shared__ MaxIdx s_mi[kThreads / 32];
    __shared__ float s_sum[kThreads / 32];
    const int r = blockIdx.x, tid = threadIdx.x;
    const float *z = logits + (long long)r * ld;
    MaxIdx m = {-INFINITY, 0x7fffffff};
    for (int v = tid; v < V; v += kThreads) m = better_first(m, MaxIdx{z[v], v});
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        MaxIdx o = {__shfl_xor_sync(0xffffffffu, m.v, d), __shfl_xor_sync(0xffffffffu, m.i, d)};
        m = better_first(m, o);
    }
    if ((tid & 31) == 0) s_mi[tid >> 5] = m;
    __syncthreads();
    m = s_mi[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) m = better_first(m, s_mi[w]);
    float sum = 0.f;
    for (int v = tid; v < V; v += kThreads) sum += expf(z[v] - m.v);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((tid & 31) == 0) s_sum[tid >> 5] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) sum += s_sum[w];
    if (probs) {
        float *p = probs + (long long)r * ld_probs;
        for (int v = tid; v < V; v += kThreads) p[v] = expf(z[v] - m.v) / sum;
    }
    if (tid == 0) {
        if (tok_out) tok_out[(long long)r * tok_stride] = m.i;
        if (tok_cur) tok_cur[r] = m.i;
        if (maxprob) maxprob[r] = 1.0f / sum;                 // softmax value of the arg-max
    }
}

int softmax_argmax(const float *logits, int ld, int rows, int V, float *probs, long long ld_probs,
                   int32_t *tok_out, int tok_stride, int32_t *tok_cur, float *maxprob,
                   cudaStream_t s) {
    if (rows <= 0) return DC_OK;
    softmax_argmax_kernel<256><<<rows, 256, 0, s>>>(logits, ld, V, probs, ld_probs, tok_out,
                                                     tok_stride, tok_cur, maxprob);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Beam search pieces (image captioning/test.py:23-64).
//   topk_softmax: per row the k largest softmax probabilities in ASCENDING order, ties broken
//   like a stable ascending argsort followed by [-k:] (larger index ranks higher).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lex_less(float av, int ai, float bv, int bi) {
    return av < bv || (av == bv && ai < bi);
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads)
topk_softmax_kernel(const float *__restrict__ logits, int ld, int V, int k,
                    int32_t *__restrict__ idx_out, float *__restrict__ p_out) {
    __shared__ float s_v[kThreads / 32];
    __shared__ int s_i[kThreads / 32];
    __shared__ float s_red[kThreads / 32];
    __shared__ float s_selv[kMaxBeam];
    __shared__ int s_seli[kMaxBeam];
    const int r = blockIdx.x, tid = threadIdx.x;
    const float *z = logits + (long long)r * ld;
    float bound_v = INFINITY;
    int bound_i = 0x7fffffff;
    for (int round = 0; round < k; ++round) {
        float bv = -INFINITY;
        int bi = -1;
        for (int v = tid; v < V; v += kThreads) {
            const float x = z[v];
            if (lex_less(x, v, bound_v, bound_i) && (bi < 0 || lex_less(bv, bi, x, v))) { bv = x; bi = v; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (oi >= 0 && (bi < 0 || lex_less(bv, bi, ov, oi))) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_v[tid >> 5] = bv; s_i[tid >> 5] = bi; }
        __syncthreads();
        bv = s_v[0]; bi = s_i[0];
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w)
            if (s_i[w] >= 0 && (bi < 0 || lex_less(bv, bi, s_v[w], s_i[w]))) { bv = s_v[w]; bi = s_i[w]; }
        if (tid == 0) { s_selv[round] = bv; s_seli[round] = bi; }
        bound_v = bv; bound_i = bi;
        __syncthreads();
    }
    // softmax denominator
    const float mx = s_selv[0];
    float sum = 0.f;
    for (int v = tid; v < V; v += kThreads) sum += expf(z[v] - mx);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((tid & 31) == 0) s_red[tid >> 5] = sum;
    __syncthreads();
    if (tid < k) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) tot += s_red[w];
        const int src = k - 1 - tid;                        // ascending order
        idx_out[(long long)r * k + tid] = s_seli[src];
        p_out[(long long)r * k + tid] = expf(s_selv[src] - mx) / tot;
    }
}

int topk_softmax(const float *logits, int ld, int rows, int V, int k, int32_t *idx_out,
                 float *p_out, cudaStream_t s) {
    if (rows <= 0) return DC_OK;
    DC_REQUIRE(k >= 1 && k <= kMaxBeam && k <= V, "beam width %d outside [1,%d]", k, kMaxBeam);
    topk_softmax_kernel<256><<<rows, 256, 0, s>>>(logits, ld, V, k, idx_out, p_out);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// One thread per RoI: pool the children of the active beams in generation order, stable
// ascending sort by score (double accumulation of fp32 probabilities), keep the last k.
// compact != 0 (first step computed once per RoI): candidate row of RoI b is b, not b*k
__global__ void beam_select_kernel(int n_roi, int k, int n_active, int compact, const int32_t *__restrict__ cand_idx,
                                   const float *__restrict__ cand_p, const double *__restrict__ score_in,
                                   double *__restrict__ score_out, int32_t *__restrict__ parent,
                                   int32_t *__restrict__ new_tok) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_roi) return;
    double sc[kMaxBeam * kMaxBeam];
    int par[kMaxBeam * kMaxBeam], tk[kMaxBeam * kMaxBeam];
    int n = 0;
    for (int j = 0; j < n_active; ++j)
        for (int c = 0; c < k; ++c) {
            const long long row = (long long)b * k + j;
            const long long crow = compact ? (long long)b : row;
            sc[n] = score_in[row] + (double)cand_p[crow * k + c];
            par[n] = j;
            tk[n] = cand_idx[crow * k + c];
            ++n;
        }
    // stable insertion sort, ascending
    for (int i = 1; i < n; ++i) {
        const double s = sc[i]; const int p = par[i], t = tk[i];
        int j = i - 1;
        while (j >= 0 && sc[j] > s) { sc[j + 1] = sc[j]; par[j + 1] = par[j]; tk[j + 1] = tk[j]; --j; }
        sc[j + 1] = s; par[j + 1] = p; tk[j + 1] = t;
    }
    for (int j = 0; j < k; ++j) {
        const int src = n - k + j;
        score_out[(long long)b * k + j] = sc[src];
        parent[(long long)b * k + j] = par[src];
        new_tok[(long long)b * k + j] = tk[src];
    }
}

int beam_select(int n_roi, int k, int n_active, const int32_t *cand_idx, const float *cand_p,
                const double *score_in, double *score_out, int32_t *parent, int32_t *new_tok,
                cudaStream_t s, bool compact) {
    if (n_roi <= 0) return DC_OK;
    beam_select_kernel<<<ceil_div(n_roi, 128), 128, 0, s>>>(n_roi, k, n_active, compact ? 1 : 0, cand_idx, cand_p,
                                                           score_in, score_out, parent, new_tok);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// dst row (b, j) <- src row (b, parent[b, j]) for a [n_roi*k, width] fp32/bf16/int32 array.
template <typename T>
__global__ void beam_gather_kernel(int rows, int k, int width, const int32_t *__restrict__ parent,
                                   const T *__restrict__ src, int ld_src, T *__restrict__ dst, int ld_dst) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    const int b = r / k;
    const long long sr = (long long)b * k + parent[r];
    for (int c = threadIdx.x; c < width; c += blockDim.x)
        dst[(long long)r * ld_dst + c] = src[sr * ld_src + c];
}

int beam_gather(int rows, int k, int width, const int32_t *parent, const void *src, int ld_src,
                void *dst, int ld_dst, int elem_bytes, cudaStream_t s) {
    if (rows <= 0 || width <= 0) return DC_OK;
    if (elem_bytes == 4)
        beam_gather_kernel<float><<<rows, 128, 0, s>>>(rows, k, width, parent, (const float *)src, ld_src, (float *)dst, ld_dst);
    else if (elem_bytes == 2)
        beam_gather_kernel<uint16_t><<<rows, 128, 0, s>>>(rows, k, width, parent, (const uint16_t *)src, ld_src, (uint16_t *)dst, ld_dst);
    else
        return set_error(DC_ERR_INVALID, "beam_gather: unsupported element size %d", elem_bytes);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__global__ void fill_i32_kernel(int32_t *p, long long n, int32_t v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
int fill_i32(int32_t *p, long long n, int32_t v, cudaStream_t s) {
    if (n <= 0) return DC_OK;
    fill_i32_kernel<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, s>>>(p, n, v);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

__global__ void set_token_column_kernel(int32_t *tokens, int rows, int stride, int col,
                                        const int32_t *__restrict__ src) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) tokens[(long long)r * stride + col] = src[r];
}
int set_token_column(int32_t *tokens, int rows, int stride, int col, const int32_t *src, cudaStream_t s) {
    if (rows <= 0) return DC_OK;
    set_token_column_kernel<<<ceil_div(rows, 256), 256, 0, s>>>(tokens, rows, stride, col, src);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// tok[r] = words[r, col]: one time step of the [rows, L] word-id matrix (v2 predict)
__global__ void token_column_kernel(const int32_t *__restrict__ words, int rows, int L, int col, int32_t *__restrict__ tok) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) tok[r] = words[(long long)r * L + col];
}
int token_column(const int32_t *words, int rows, int L, int col, int32_t *tok, cudaStream_t s) {
    if (rows <= 0) return DC_OK;
    token_column_kernel<<<ceil_div(rows, 256), 256, 0, s>>>(words, rows, L, col, tok);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

template <typename T>
__global__ void f32_to_kernel(const float *__restrict__ src, T *__restrict__ dst, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if constexpr (sizeof(T) == 2) dst[i] = __float2bfloat16_rn(src[i]);
        else dst[i] = src[i];
    }
}
int f32_to_bf16(const float *src, void *dst, long long n, cudaStream_t s) {
    if (n <= 0) return DC_OK;
    f32_to_kernel<__nv_bfloat16><<<(unsigned)ceil_div<long long>(n, 256), 256, 0, s>>>(src, (__nv_bfloat16 *)dst, n);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

// out[c, r] = in[r, c]  (weights re-layout at set_weights time; not on the hot path)
template <typename T>
__global__ void transpose_to_kernel(const float *__restrict__ in, int rows, int cols, T *__restrict__ out,
                                    int ld_out) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(long long)r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) {
            if constexpr (sizeof(T) == 2) out[(long long)c * ld_out + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
            else out[(long long)c * ld_out + r] = tile[threadIdx.x][i];
        }
    }
}
int transpose_f32(const float *in, int rows, int cols, void *out, int ld_out, bool bf16, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return DC_OK;
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
    if (bf16) transpose_to_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(in, rows, cols, (__nv_bfloat16 *)out, ld_out);
    else transpose_to_kernel<float><<<grid, block, 0, s>>>(in, rows, cols, (float *)out, ld_out);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

}  // namespace dcap

// The whole greedy decoding loop of the v1 word model (text_generation_model.py:130-156, 192-232) as ONE
// persistent tcgen05 kernel.
//
// Per decoding step the bf16 path runs four dependent GEMMs (decoder_bf16.cu): LSTM1 gates + cell, LSTM2 gates +
// cell, Dense(1024)+ReLU, vocabulary projection + arg-max, and a merge that picks the token and fetches its
// embedding row.  As separate launches the three small GEMMs (26 % of the FLOPs) take half of the step: each has
// <= 4 tile waves, so its ~4 us prologue, its last epilogue and its partial last wave are not amortised.  Every one
// of these dependencies is PER ROW BLOCK, though: rows [256 rb, 256 rb + 256) of step t need nothing from any other
// row block.  This kernel therefore walks ONE global list of work items
//
//     item = (step t, stage s, row block rb, column block cb)
//
// ordered as a diagonal wavefront over (step, row block) (decode_item) and dealt round-robin over the CTA pairs
// (cluster of 2, tcgen05 cta_group::2, 256 x 256 tiles, the main loop of gemm_bf16_tc2_kernel).  An item waits for
// the items it depends on through monotone counters in global memory (one per stage and 128-row block); since every
// dependency points to an EARLIER item of the list and every pair works through its items in list order, the earliest
// unfinished item can always run: no deadlock as long as all pairs are resident (one CTA per SM; the host asks the
// occupancy API how many clusters fit).  The tail of one stage overlaps the head of the next, tile waves never drain,
// and the 75 launches of a 15-step loop become one.
//
//   stage 0  z1 = [emb | h1] . [W1e ; U1]^T + (f . W1f + b1)      -> Keras LSTM cell -> h1 (bf16, into X1' and X2), c1
//   stage 1  z2 = [h1 | h2] . [W2 ; U2]^T + b2                    -> Keras LSTM cell -> h2 (bf16, into X2'), c2
//   stage 2  d  = relu(h2 . Wd1h^T + (f . Wd1f + bd1))            -> bf16
//   stage 3  per row and 128-column region: {max, first arg-max, sum exp, step tag} of d . Wd2^T + bd2 -> partial[region][row]
//   stage 4  merge item (no GEMM): per row the token id, optional caption score, and the token's embedding row copied
//            into the next step's [emb | h1] operand; it polls the step tags of the partials (flag in data).
//
// Warp roles per CTA (10 warps): warp 0 = TMA producer (dependency wait + operand loads), warp 1 = MMA issuer in the
// leader CTA / PUBLISHER in CTA 1 (one gpu-scope fence + counter increments per item once the pair's 16 epilogue warps
// have arrived on a cluster-scope mbarrier), warps 2..9 = epilogue (TMEM lane quarter x column half).
//
// Memory-model notes.  Operands written by other SMs' epilogues (generic proxy) are read by TMA (async proxy): the
// publisher fences at gpu scope before it bumps a counter, the TMA producer acquires the counter and issues
// fence.proxy.async.global before its loads.  Mutable data read by epilogue threads (c, tokens, previous h, partials)
// is loaded with ld.global.cg (L2, the coherence point), and only after the producer has seen the dependency (a
// shared-memory barrier ring from the producer to the epilogue warps).  Every wait carries a watchdog: after ~2 s it
// records a code in an error word that all waits poll, so a protocol bug or a second spinning kernel on the same GPU
// drains the grid (with garbage results, reported under DCAP_LOOP_DEBUG) instead of hanging the device.
//
// Measured without gain and not kept: L2 prefetch of the next item's hoisted-term / cell-state lines (bulk prefetch from
// the producer warp: slower; prefetch.global.L2 from the epilogue warps, row-major and blocked-32: 2.95-3.05 ms either
// way), 16 epilogue warps of 64 columns (96 registers: spills, 3.83 vs 3.30 ms), two chunks of addend look-ahead
// (spills), bias values of the vocabulary / LSTM2 epilogues staged in shared memory before the accumulator wait (the
// vocabulary tile's hold time 2.2 -> 1.85 us, the kernel 2.94 ms either way), a SEVENTH operand-ring slot in the
// shared memory the missing output staging leaves free (2.96 vs 2.93 ms at 8000 RoIs, 1.155 vs 1.105 at 2500: the main
// loop's ~0.39 us per k-block -- 0.27 at the tensor peak -- is not a bytes-in-flight x latency limit, and the L1 the
// epilogues' global loads live in shrinks).  DESIGN.md section 4 has the numbers.
#include "decoder.cuh"
#include "decoder_bf16.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace dcap {

enum { kMapX1a = 0, kMapX1b, kMapX2a, kMapX2b, kMapH2a, kMapH2b, kMapD, kMapW1, kMapW2, kMapWd1, kMapWd2, kMapF, kMapW1f, kMapWd1f, kLoopMaps };
struct LoopMaps { CUtensorMap m[kLoopMaps]; };

struct LoopParams {
    int R, P, V, U, Epad, K1;
    int tiles_m;                       // row blocks of 256
    int tiles_n[5], num_kb[5], first[6];   // per stage (4 = merge, one item, no GEMM): column tiles, k-blocks, first item of the stage inside a slot ([5] = items per slot)
    int skew[5], total;                // item order: slot v holds stage s of virtual row block v - skew[s]; total = number of items
    int map_a[4][2], map_b[4];         // tensor-map index of the A operand by step parity, and of B
    int kb_main[4], map_b2[4];         // folded feature term: k-blocks >= kb_main[s] come from (Fb, map_b2[s]) -- [x | f] . [W ; Wf]^T in one accumulator
    const float *b1, *bd1;             // fold: biases of LSTM1 (gate-interleaved) and Dense(1024) (without fold they sit inside g1f / d1f)
    int fold, pfence, lookahead, defer;  // knobs (see greedy_loop_bf16)
    const float *g1f, *d1f, *b2, *bias_v;   // g1f / d1f: column blocks of ONE blocked-32 array of row length hoist_ld4 float4
    int g1f_ld4, d1f_ld4;
    float *c1, *c2;
    __nv_bfloat16 *X1[2], *X2[2], *d;
    float4 *partial;                   // [slots][R] {max, arg-max bits, sum exp, -}
    int slots;
    const __nv_bfloat16 *emb;
    int32_t *tok, *tokens;
    float *scores;                     // kSum only
    unsigned int *cnt;                 // [s * n128 + 128-row block], s = 0..4: epilogue warps that have finished that block of a stage-s item; then the error word at 8 * n128
    int n128;
    int writer_proxy_fence;            // 1: epilogue warps also run fence.proxy.async before they publish (belt and braces; measured)
    unsigned long long *trace;         // debugging: [pairs][trace_items][12] globaltimer marks of the leader CTA (DCAP_LOOP_TRACE)
    int trace_items;
};

constexpr int kLoopThreads = kThreads;
constexpr int kRing = 8;                                 // item-done / dependency-seen barrier rings
constexpr long long kWatchdogCycles = 4000000000ll;      // ~2 s

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// Watchdog: a wait that lasts ~2 s records its code in the error word; every wait polls that word and gives up once
// it is set, so the whole grid drains (with garbage results) instead of hanging the device.  The host checks the
// word after the call when DCAP_LOOP_DEBUG is set (tools/loop_check.py) -- in production it never fires.
__device__ __forceinline__ bool loop_aborted(const unsigned *err) { return *reinterpret_cast<const volatile unsigned *>(err) != 0; }
// counter >= target (acquire)
__device__ __forceinline__ void wait_count(const unsigned *p, unsigned target, unsigned *err, unsigned code) {
    if (ld_acquire_u32(p) >= target) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (ld_acquire_u32(p) < target) {
        __nanosleep(40);
        if ((spins++ & 63u) == 0) {
            if (loop_aborted(err)) return;
            if (clock64() - t0 > kWatchdogCycles) { atomicCAS(err, 0u, code | (target << 8)); return; }
        }
    }
}
__device__ __forceinline__ void mbar_wait_acq_cluster_wd(uint64_t *bar, uint32_t parity, unsigned *err, unsigned code) {
    if (mbar_try_wait_acq_cluster(bar, parity)) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait_acq_cluster(bar, parity)) {
        if ((spins++ & 255u) == 0) {
            if (loop_aborted(err)) return;
            if (clock64() - t0 > kWatchdogCycles) { atomicCAS(err, 0u, code); return; }
        }
    }
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity, unsigned *err, unsigned code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((spins++ & 255u) == 0) {
            if (loop_aborted(err)) return;
            if (clock64() - t0 > kWatchdogCycles) { atomicCAS(err, 0u, code); return; }
        }
    }
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
#define LOOP_TRACE(slot)                                                                                      \
    do {                                                                                                      \
        if (p.trace && rank == 0) {                                                                           \
            const int n_ = (item - pair) / num_pairs;                                                         \
            if (n_ < p.trace_items) p.trace[((long long)pair * p.trace_items + n_) * 12 + (slot)] = globaltimer_ns(); \
        }                                                                                                     \
    } while (0)

// Item order: a diagonal wavefront over (step, row block).  Slot v of the list holds, one after the other, the column
// tiles of stage 0 for virtual row block u = v, of stage 1 for u = v - skew[1], of stage 2 for u = v - skew[2], of
// the vocabulary stage for u = v - skew[3] and the merge item for u = v - skew[4]; u = t * tiles_m + rb.  So at any moment different row blocks are in
// different stages: every pair sees a fine mixture of long-epilogue (LSTM cells) and long-main-loop (vocabulary) tiles
// -- the epilogues hide under the next tile's MMAs --, and every dependency was listed skew slots (several tile times)
// earlier, so it is normally met when its consumer comes up.  skew[4] < tiles_m keeps "dependencies point backwards":
// stage 0 of (t + 1, rb) sits tiles_m - skew[4] slots after the merge of (t, rb).
struct LoopItem { int t, s, rb, cb; bool live; };
__device__ __forceinline__ LoopItem decode_item(const LoopParams &p, int item) {
    LoopItem it;
    const int ips = p.first[5];
    const int v = item / ips;
    const int j = item - v * ips;
    it.s = (j >= p.first[1]) + (j >= p.first[2]) + (j >= p.first[3]) + (j >= p.first[4]);
    it.cb = j - p.first[it.s];
    const int u = v - p.skew[it.s];
    it.live = u >= 0 && u < p.P * p.tiles_m;
    it.t = u / p.tiles_m;
    it.rb = u - it.t * p.tiles_m;
    return it;
}

// ---- epilogue bodies: one warp = 32 rows (lane = row) x 128 columns of the CTA's 128 x 256 accumulator half ----

// "Blocked-32" layout of the fp32 arrays that the epilogues touch with lane = row (hoisted per-RoI terms, cell state):
// [row / 32][column / 4][row % 32][4 floats], so that the 32 lanes of a warp (32 consecutive rows, same column
// quad) read or write 512 CONTIGUOUS bytes.  Row-major, a lane = row access touches 32 different lines per
// instruction and the load/store unit serialises them: 16 bytes per cycle and SM, which made the LSTM1 tile's
// epilogue LSU-bound (16 k LSU cycles = 8.6 us of its 11.4 us, measured).  ld4 = row length in float4.
__device__ __forceinline__ const float4 *blk32(const float *base, long long m, int ld4) {
    return reinterpret_cast<const float4 *>(base) + ((m >> 5) * ld4) * 32 + (m & 31);
}
__device__ __forceinline__ float4 *blk32(float *base, long long m, int ld4) {
    return reinterpret_cast<float4 *>(base) + ((m >> 5) * ld4) * 32 + (m & 31);
}

// Deferred publication (big batches): a warp hands its PREVIOUS item to the publisher here, right after the wait for
// the current item's accumulator and before the current item's first store -- the previous item's stores have
// drained by then, so the release does not stall the warp (~1.4 us per item when it directly follows the stores).
__device__ __forceinline__ void publish_pending(uint32_t pend) {
    if (pend) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive_release_cluster(pend);
    }
}

// Keras LSTM cell on gate-interleaved columns (column 4u+g): z = acc + addend / bias; hard-sigmoid gates, tanh
// candidate, masked rows (consumed token 0) carry (h, c).  c fp32 in place (blocked-32), h bf16 into one or two
// operand buffers (row-major: they are TMA operands).  add_blk / c_blk = blk32(...) of this lane's row.
template <bool kAdd>
__device__ __forceinline__ void loop_cell(uint32_t taddr, int n0, bool valid, const float4 *add_blk, const float *bias,
                                          float4 *c_blk, bool masked, const __nv_bfloat16 *h_prev_row,
                                          __nv_bfloat16 *h_a_row, __nv_bfloat16 *h_b_row, uint64_t *full_bar,
                                          uint32_t full_phase, unsigned *err, uint32_t pend) {
    // The cell state of all four chunks is requested before the accumulator is waited for (it does not depend on
    // it): inside the TMEM-holding part only the addend / bias rows are still fetched, one chunk ahead.
    float4 a_buf[2][8], c_all[8];
    uint4 h_buf[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    auto load_operands = [&](float4 (&a)[8], uint4 &hp, int nb) {
        if constexpr (kAdd) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __ldg(add_blk + ((nb >> 2) + j) * 32);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __ldg(reinterpret_cast<const float4 *>(bias + nb + 4 * j));
        }
        if (masked) hp = __ldcg(reinterpret_cast<const uint4 *>(h_prev_row + (nb >> 2)));
    };
#pragma unroll
    for (int j = 0; j < 8; ++j) c_all[j] = __ldcg(c_blk + ((n0 >> 4) + j) * 32);
    load_operands(a_buf[0], h_buf[0], n0);
    mbar_wait_wd(full_bar, full_phase, err, 0x30u);
    tc_fence_after();
    publish_pending(pend);
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
        const int nb = n0 + c0, b = (c0 >> 5) & 1;
        float4 a_cur[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a_cur[j] = a_buf[b][j];
        const float4 c_lo = c_all[c0 >> 4], c_hi = c_all[(c0 >> 4) + 1];
        const float c_old[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
        const uint32_t hw[4] = {h_buf[b].x, h_buf[b].y, h_buf[b].z, h_buf[b].w};
        if (c0 + 32 < 128) load_operands(a_buf[b ^ 1], h_buf[b ^ 1], nb + 32);
        float v[32];
        tmem_ld32(taddr + c0, v);
        float c_new[8], h_new[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float zi = v[4 * j] + a_cur[j].x, zf = v[4 * j + 1] + a_cur[j].y;
            const float zg = v[4 * j + 2] + a_cur[j].z, zo = v[4 * j + 3] + a_cur[j].w;
            const float ig = hard_sigmoid_tc(zi), fg = hard_sigmoid_tc(zf);
            const float gg = tanh_fast(zg), og = hard_sigmoid_tc(zo);
            c_new[j] = __fadd_rn(__fmul_rn(fg, c_old[j]), __fmul_rn(ig, gg));
            h_new[j] = __fmul_rn(og, tanh_fast(c_new[j]));
            if (masked) {                                            // K.rnn mask: carry (h, c)
                c_new[j] = c_old[j];
                h_new[j] = __uint_as_float((j & 1) ? (hw[j >> 1] & 0xffff0000u) : (hw[j >> 1] << 16));
            }
        }
        if (valid) {
            const int u0 = nb >> 2;
            c_blk[(u0 >> 2) * 32] = make_float4(c_new[0], c_new[1], c_new[2], c_new[3]);
            c_blk[((u0 >> 2) + 1) * 32] = make_float4(c_new[4], c_new[5], c_new[6], c_new[7]);
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 t2 = __floats2bfloat162_rn(h_new[2 * j], h_new[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t *>(&t2);
            }
            const uint4 hv = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4 *>(h_a_row + u0) = hv;
            if (h_b_row) *reinterpret_cast<uint4 *>(h_b_row + u0) = hv;
        }
    }
}

// d = relu(acc + addend) -> bf16.  blocked: the addend is blk32(...) of this lane's row; else a row-major row
// (the bias vector in fold mode)
template <bool kBlocked>
__device__ __forceinline__ void loop_dense(uint32_t taddr, int n0, bool valid, const float4 *add4, __nv_bfloat16 *out_row,
                                           uint64_t *full_bar, uint32_t full_phase, unsigned *err, uint32_t pend) {
    float4 a_buf[2][8];
    auto load_operands = [&](float4 (&a)[8], int nb) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __ldg(add4 + ((nb >> 2) + j) * (kBlocked ? 32 : 1));
    };
    load_operands(a_buf[0], n0);
    load_operands(a_buf[1], n0 + 32);
    mbar_wait_wd(full_bar, full_phase, err, 0x31u);
    tc_fence_after();
    publish_pending(pend);
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
        const int nb = n0 + c0;
        float4 a_cur[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a_cur[j] = a_buf[(c0 >> 5) & 1][j];
        if (c0 + 64 < 128) load_operands(a_buf[(c0 >> 5) & 1], nb + 64);
        float v[32];
        tmem_ld32(taddr + c0, v);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x0 = fmaxf(v[4 * j] + a_cur[j].x, 0.f), x1 = fmaxf(v[4 * j + 1] + a_cur[j].y, 0.f);
            const float x2 = fmaxf(v[4 * j + 2] + a_cur[j].z, 0.f), x3 = fmaxf(v[4 * j + 3] + a_cur[j].w, 0.f);
            __nv_bfloat162 t0 = __floats2bfloat162_rn(x0, x1), t1 = __floats2bfloat162_rn(x2, x3);
            pk[2 * j] = *reinterpret_cast<uint32_t *>(&t0);
            pk[2 * j + 1] = *reinterpret_cast<uint32_t *>(&t1);
        }
        if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                reinterpret_cast<uint4 *>(out_row + nb)[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
    }
}

// 32 lanes x 64 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// per row: max, first arg-max (, sum exp(v - max)) of acc + bias over this warp's 128 columns.  64 columns per TMEM
// load; the arg-max is a TREE of (value, index) pairs -- a serial "if (x > best)" scan is a 32-long dependent chain
// per chunk that two warps per scheduler cannot hide.  In a pair the higher index wins only on strict >, so the
// first index of the maximum survives every level, as np.argmax.
template <bool kSum>
__device__ __forceinline__ float4 loop_argmax(uint32_t taddr, int n0, int N, const float *bias, uint64_t *full_bar,
                                              uint32_t full_phase, unsigned *err, uint32_t pend, int tag,
                                              unsigned long long *tr) {
    mbar_wait_wd(full_bar, full_phase, err, 0x32u);
    tc_fence_after();
    publish_pending(pend);
    if (tr) tr[10] = globaltimer_ns();
    float best = -INFINITY, sum = 0.f;
    int best_i = 0x7fffffff;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 64) {
        const int nb = n0 + c0;
        if (nb >= N) break;                                          // warp-uniform
        float v[64];
        tmem_ld64(taddr + c0, v);
        if (tr && c0 == 0) tr[11] = globaltimer_ns();
        if (nb + 64 <= N) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + nb + 4 * j));
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = (nb + j < N) ? v[j] + __ldg(bias + min(nb + j, N - 1)) : -INFINITY;
        }
        float m[32];
        int ix[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const bool hi = v[2 * j + 1] > v[2 * j];
            m[j] = hi ? v[2 * j + 1] : v[2 * j];
            ix[j] = hi ? 2 * j + 1 : 2 * j;
        }
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
#pragma unroll
            for (int j = 0; j < w; ++j) {
                const bool hi = m[2 * j + 1] > m[2 * j];
                m[j] = hi ? m[2 * j + 1] : m[2 * j];
                ix[j] = hi ? ix[2 * j + 1] : ix[2 * j];
            }
        }
        const float cmax = m[0];
        if constexpr (kSum) {
            if (cmax > best) sum *= __expf(best - cmax);
        }
        if (cmax > best) { best = cmax; best_i = nb + ix[0]; }
        if constexpr (kSum) {
#pragma unroll
            for (int j = 0; j < 64; ++j) sum += __expf(v[j] - best);
        }
    }
    return make_float4(best, __int_as_float(best_i), sum, __int_as_float(tag));
}

// Merge of the vocabulary stage's partials for 16 rows (one warp): lanes 2r / 2r+1 scan the lower / upper half of
// row r's column regions (8 loads in flight each) and combine; ties keep the smaller column index, as np.argmax.
// Then the token ids, the caption score (kSum) and the tokens' embedding rows, copied as bf16 into the next step's
// [emb | h1] operand (Embedding lookup of the greedy feedback, text_generation_model.py:147,222-225).
template <bool kSum>
__device__ __forceinline__ void loop_merge(const LoopParams &p, int m_base, int t, int lane, unsigned *err) {
    const int r = lane >> 1, h = lane & 1;
    const int m = m_base + r;
    const bool valid = m < p.R;
    const long long mr = valid ? m : (long long)(p.R - 1);
    const int half_slots = (p.slots + 1) >> 1;
    const int s_lo = h * half_slots, s_hi = min(p.slots, s_lo + half_slots);
    float best = -INFINITY, sum = 0.f;
    int bi = 0x7fffffff;
    const float4 *pp = p.partial + mr;
    const int tag = t + 1;
    for (int s0 = s_lo; s0 < s_hi; s0 += 8) {
        float4 q[8];
        // Flag in data: a partial is ONE 16-byte store {max, arg-max, sum exp, step tag}, so the entry itself says whether
        // this step's value has arrived (the buffer is zeroed before every launch) -- the vocabulary tiles, two thirds
        // of all items, publish nothing through a fence.
        const long long t0 = clock64();
        for (unsigned spins = 0;; ++spins) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                q[i] = __ldcg(pp + (long long)min(s0 + i, s_hi - 1) * p.R);
                ok = ok && __float_as_int(q[i].w) == tag;
            }
            if (ok) break;
            __nanosleep(100);
            if ((spins & 63u) == 63u) {
                if (loop_aborted(err)) break;
                if (clock64() - t0 > kWatchdogCycles) { atomicCAS(err, 0u, 0x45u); break; }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (s0 + i >= s_hi) break;
            const int idx = __float_as_int(q[i].y);
            if constexpr (kSum) {
                if (q[i].x > best) { sum = sum * __expf(best - q[i].x) + q[i].z; best = q[i].x; bi = idx; }
                else if (q[i].x == best) { sum += q[i].z; bi = min(bi, idx); }
                else sum += q[i].z * __expf(q[i].x - best);
            } else {
                if (q[i].x > best) { best = q[i].x; bi = idx; }       // regions ascend in column order: strict > keeps the first index
            }
        }
    }
    {
        const float ob = __shfl_xor_sync(0xffffffffu, best, 1);
        const float os = __shfl_xor_sync(0xffffffffu, sum, 1);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, 1);
        // a = lower half (smaller indices), b = upper half
        const float ab = h ? ob : best, bb = h ? best : ob, as = h ? os : sum, bs = h ? sum : os;
        const int ai = h ? oi : bi, bi2 = h ? bi : oi;
        if (bb > ab) { best = bb; bi = bi2; if constexpr (kSum) sum = bs + as * __expf(ab - bb); }
        else { best = ab; bi = ai; if constexpr (kSum) sum = as + (bb > -INFINITY ? bs * __expf(bb - ab) : 0.f); }
    }
    if ((unsigned)bi >= (unsigned)p.V) bi = 0;                        // all-NaN row
    if (valid && h == 0) {
        p.tokens[(long long)m * p.P + t] = bi;
        p.tok[m] = bi;
        if constexpr (kSum) p.scores[m] = (t ? __ldcg(p.scores + m) : 0.f) + logf(1.0f / sum);
    }
    if (t + 1 < p.P) {
        // the 16 x e8 16-byte pieces are dealt over the lanes, 8 loads in flight each (a row-by-row copy serialises
        // load -> store -> load on possible aliasing: measured ~100 us per 32 rows)
        const uint4 *__restrict__ emb4 = reinterpret_cast<const uint4 *>(p.emb);
        uint4 *__restrict__ xn4 = reinterpret_cast<uint4 *>(p.X1[(t + 1) & 1]);
        const int e8 = p.Epad >> 3, k8 = p.K1 >> 3;
        const int iters = e8 >> 1;                                    // 16 * e8 pieces / 32 lanes
        for (int k0 = 0; k0 < iters; k0 += 8) {
            uint4 v8[8];
            long long dsti[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int q = lane + 32 * min(k0 + i, iters - 1);
                const int rr = q / e8, j = q - rr * e8;
                const int ti = __shfl_sync(0xffffffffu, bi, 2 * rr);
                v8[i] = __ldg(emb4 + (long long)ti * e8 + j);
                dsti[i] = (m_base + rr < p.R && k0 + i < iters) ? (long long)(m_base + rr) * k8 + j : -1;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (dsti[i] >= 0) xn4[dsti[i]] = v8[i];
        }
    }
    // the embedding rows are read by TMA (async proxy) in the next step
    if (p.writer_proxy_fence) fence_proxy_async_all();
}

template <bool kSum>
__global__ void __launch_bounds__(kLoopThreads, 1)
greedy_loop_kernel(const __grid_constant__ LoopMaps maps, const LoopParams p) {
    using S = TcSmem2;
    constexpr int kStages = S::kStages;
    constexpr int kBlockN = 256;
    constexpr uint32_t kTmemCols = 2 * kBlockN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + kStages * S::kStageA;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + S::kBarOff);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *tmem_full = empty_bar + kStages;
    uint64_t *tmem_empty = tmem_full + 2;
    uint64_t *done_bar = tmem_empty + 2;                           // [kRing] (CTA 1's copy is used) the 16 epilogue warps of the pair have finished an item -> publisher
    uint64_t *dep_bar = done_bar + kRing;                          // [kRing] the producer has seen a stage-0/1 item's dependency -> epilogue warps
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(dep_bar + kRing);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int total = p.total;
    const int n128 = p.n128;
    unsigned *err = p.cnt + 8 * n128;
    unsigned *cnt_stage = p.cnt;                    // + s * n128 + rb128

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kLoopMaps; ++i) tma_prefetch_desc(&maps.m[i]);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * kEpiWarps); }
        for (int i = 0; i < kRing; ++i) { mbar_init(&done_bar[i], 2 * kEpiWarps); mbar_init(&dep_bar[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_2sm(tmem_ptr, kTmemCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        // The dependency of the NEXT item is polled (acquire loads) in the producer's idle time -- the spins on a full
        // operand ring -- so that an item whose dependency is met (the normal case) starts loading with no L2 round trip
        // in front of it: with six ring slots the producer is only ~2 us ahead of the MMAs, and ~0.5 us per item spent
        // on the counter showed up as MMA stalls (28 us per step, measured).
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int n_dep = 0;                                             // stage-0/1 items so far (dep_bar ring position)
            auto next_gemm = [&](int item) {                           // next live GEMM item of this pair (>= item), or total
                for (; item < total; item += num_pairs) {
                    const LoopItem it = decode_item(p, item);
                    if (it.live && it.s != 4) break;
                }
                return item;
            };
            auto dependency = [&](const LoopItem &it, const unsigned *&cnt, unsigned &target) {
                const int rb128 = it.rb * 2 + (int)rank;
                if (it.s == 0) { cnt = cnt_stage + 4 * n128 + rb128; target = (unsigned)it.t; }
                else { cnt = cnt_stage + (it.s - 1) * n128 + rb128; target = (unsigned)p.tiles_n[it.s - 1] * (it.t + 1); }
            };
            int item = next_gemm(pair);
            bool ready = false;
            while (item < total) {
                const LoopItem it = decode_item(p, item);
                const int m0 = it.rb * 256 + (int)rank * 128;
                LOOP_TRACE(0);
                // operands of this CTA's 128 rows written by earlier items
                if (!ready) {
                    const unsigned *cnt; unsigned target;
                    dependency(it, cnt, target);
                    if (target) wait_count(cnt, target, err, 0x10u + it.s);
                }
                // the epilogue warps of stage-0/1 items read state published by the same items: tell them it is there
                // (cta-scope release on top of the gpu-scope acquire)
                if (it.s <= 1) { mbar_arrive(&dep_bar[n_dep & (kRing - 1)]); ++n_dep; }
                LOOP_TRACE(1);
                if (p.pfence == 1) fence_proxy_async_all();
                else if (p.pfence == 2) fence_proxy_async_global();
                LOOP_TRACE(2);
                const int nxt = next_gemm(item + num_pairs);
                const unsigned *ncnt = nullptr; unsigned ntarget = 0;
                if (nxt < total) { const LoopItem ni = decode_item(p, nxt); dependency(ni, ncnt, ntarget); }
                bool nready = ntarget == 0;
                const CUtensorMap *ma = &maps.m[p.map_a[it.s][it.t & 1]], *mb = &maps.m[p.map_b[it.s]];
                const CUtensorMap *ma2 = &maps.m[kMapF], *mb2 = &maps.m[p.map_b2[it.s]];
                const int nb0 = it.cb * kBlockN + (int)rank * 128;       // this CTA's half of the B tile
                const int num_kb = p.num_kb[it.s], kb_main = p.kb_main[it.s];
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (!mbar_try_wait(&empty_bar[stage], phase ^ 1)) {
                        const long long t0 = clock64();
                        unsigned spins = 0;
                        do {
                            if (!nready && p.lookahead) nready = ld_acquire_u32(ncnt) >= ntarget;
                            if ((spins++ & 255u) == 0) {
                                if (loop_aborted(err)) break;
                                if (clock64() - t0 > kWatchdogCycles) { atomicCAS(err, 0u, 0x20u); break; }
                            }
                        } while (!mbar_try_wait(&empty_bar[stage], phase ^ 1));
                    }
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (S::kStageA + S::kStageB));
                    const uint32_t bar = mapa_u32(&full_bar[stage], 0);
                    const bool main = kb < kb_main;
                    const int kc = (main ? kb : kb - kb_main) * kBlockK;
                    tma_load_2d_2sm(main ? ma : ma2, bar, smem_a + stage * S::kStageA, kc, m0);
                    tma_load_2d_2sm(main ? mb : mb2, bar, smem_b + stage * S::kStageB, kc, nb0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                ready = nready;
                item = nxt;
            }
        }
        __syncwarp();
    } else if (warp == 1 && rank == 1) {
        // ===================== publisher (this warp is idle in CTA 1: the leader issues the MMAs) =====================
        // An item's results become visible to the other SMs here: once the pair's 16 epilogue warps have arrived on the
        // item's barrier (cluster-scope release / acquire), ONE gpu-scope fence and one counter increment per 128-row
        // block publish them all (cumulativity: the pattern of a grid barrier).  The epilogue warps never wait for a
        // fence themselves.
        if (lane == 0) {
            int n_done = 0;
            for (int item = pair; item < total; item += num_pairs) {
                const LoopItem it = decode_item(p, item);
                if (!it.live) continue;
                mbar_wait_acq_cluster_wd(&done_bar[n_done & (kRing - 1)], (n_done / kRing) & 1, err, 0x50u);
                ++n_done;
                __threadfence();
                atomicAdd(cnt_stage + it.s * n128 + it.rb * 2, 1u);
                atomicAdd(cnt_stage + it.s * n128 + it.rb * 2 + 1, 1u);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = make_idesc_bf16(2 * kBlockM, kBlockN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = pair; item < total; item += num_pairs) {
                const LoopItem it = decode_item(p, item);
                if (!it.live || it.s == 4) continue;
                const int num_kb = p.num_kb[it.s];
                mbar_wait_wd(&tmem_empty[acc], acc_phase ^ 1, err, 0x21u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kBlockN;
                LOOP_TRACE(3);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_wd(&full_bar[stage], phase, err, 0x22u);
                    tc_fence_after();
                    if (kb == 0) LOOP_TRACE(4);
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * S::kStageA), 16);
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * S::kStageB), 16);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, kb > 0 || k != 0);
                    umma_commit_2sm(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tmem_full[acc]);
                LOOP_TRACE(5);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9, both CTAs: own 128 x 256 accumulator half) =====================
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                              // which 128 of the tile's 256 columns
        int acc = 0;
        uint32_t acc_phase = 0;
        int n_done = 0, n_dep = 0;                                     // ring positions: items finished, stage-0/1 items seen
        uint32_t pend = 0;                                             // deferred publication: the previous item's done barrier (cluster address) or 0
        for (int item = pair; item < total; item += num_pairs) {
            const LoopItem it = decode_item(p, item);
            if (!it.live) continue;
            const int rb128 = it.rb * 2 + (int)rank;
            const int par = it.t & 1;
            if (warp == 2 && lane == 0) LOOP_TRACE(8);
            if (it.s == 4) {
                // ---- merge item: token, caption score and next embedding row of this CTA's 128 rows; 16 rows per warp ----
                publish_pending(pend);                                 // never block on a counter with an item unpublished
                pend = 0;
                loop_merge<kSum>(p, it.rb * 256 + (int)rank * 128 + (warp - 2) * 16, it.t, lane, err);
                __syncwarp();
                if (lane == 0) mbar_arrive_release_cluster(mapa_u32(&done_bar[n_done & (kRing - 1)], 1));
                ++n_done;
                if (warp == 2 && lane == 0) { LOOP_TRACE(6); LOOP_TRACE(7); }
                continue;
            }
            const int m_base = it.rb * 256 + (int)rank * 128 + quarter * 32;
            const int m = m_base + lane;
            const bool valid = m < p.R;
            const long long mr = valid ? m : (long long)(p.R - 1);
            const int n0 = it.cb * kBlockN + half * 128;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kBlockN + half * 128;
            if (it.s <= 1) {
                // This warp reads state written by other SMs (c, consumed token, previous h), and it requests that state
                // BEFORE it waits for the accumulator -- i.e. possibly before this CTA's producer has seen the dependency:
                // it waits until the producer, which acquired the dependency counter, says so (without a wait small batches,
                // whose dependencies are only just met, decode wrong tokens: measured).
                mbar_wait_wd(&dep_bar[n_dep & (kRing - 1)], (n_dep / kRing) & 1, err, 0x40u);
                ++n_dep;
                const bool masked = __ldcg(p.tok + mr) == 0;
                const int u4 = p.U >> 2;
                if (it.s == 0 && p.fold)
                    loop_cell<false>(taddr, n0, valid, nullptr, p.b1, blk32(p.c1, mr, u4), masked,
                                     p.X1[par] + mr * p.K1 + p.Epad, p.X1[par ^ 1] + mr * p.K1 + p.Epad,
                                     p.X2[par] + mr * (2ll * p.U), &tmem_full[acc], acc_phase, err, pend);
                else if (it.s == 0)
                    loop_cell<true>(taddr, n0, valid, blk32(p.g1f, mr, p.g1f_ld4), nullptr, blk32(p.c1, mr, u4), masked,
                                    p.X1[par] + mr * p.K1 + p.Epad, p.X1[par ^ 1] + mr * p.K1 + p.Epad,
                                    p.X2[par] + mr * (2ll * p.U), &tmem_full[acc], acc_phase, err, pend);
                else
                    loop_cell<false>(taddr, n0, valid, nullptr, p.b2, blk32(p.c2, mr, u4), masked,
                                     p.X2[par] + mr * (2ll * p.U) + p.U, p.X2[par ^ 1] + mr * (2ll * p.U) + p.U,
                                     nullptr, &tmem_full[acc], acc_phase, err, pend);
            } else if (it.s == 2) {
                // fold: the addend is the bias row (same for every lane: broadcast loads)
                if (p.fold) loop_dense<false>(taddr, n0, valid, reinterpret_cast<const float4 *>(p.bd1), p.d + mr * kDense, &tmem_full[acc], acc_phase, err, pend);
                else loop_dense<true>(taddr, n0, valid, blk32(p.d1f, mr, p.d1f_ld4), p.d + mr * kDense, &tmem_full[acc], acc_phase, err, pend);
            } else {
                const float4 r4 = loop_argmax<kSum>(taddr, n0, p.V, p.bias_v, &tmem_full[acc], acc_phase, err, pend, it.t + 1,
                                                    (p.trace && rank == 0 && warp == 2 && lane == 0 && (item - pair) / num_pairs < p.trace_items)
                                                        ? p.trace + ((long long)pair * p.trace_items + (item - pair) / num_pairs) * 12 : nullptr);
                if (valid) p.partial[(long long)(it.cb * 2 + half) * p.R + m] = r4;
            }
            pend = 0;                                                  // published inside the body
            if (warp == 2 && lane == 0) LOOP_TRACE(9);
            // accumulator buffer drained: hand it back to the MMA issuer (the leader's barrier)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (warp == 2 && lane == 0) LOOP_TRACE(6);
            // this warp's rows of the tile are written: hand them to the publisher -- now, or (big batches) after the wait
            // for the next item's accumulator
            if (p.writer_proxy_fence) fence_proxy_async_all();
            __syncwarp();
            pend = mapa_u32(&done_bar[n_done & (kRing - 1)], 1);
            // deferral is safe only if this pair's NEXT list item is live: then the item that waits is one stride
            // (num_pairs items) behind the unpublished one, closer than any dependency reaches (host: p.defer), so no
            // chain of dependencies can lead from the unpublished item to an item some pair is blocked on
            bool defer = p.defer && it.s != 3 && item + num_pairs < total;
            if (defer) defer = decode_item(p, item + num_pairs).live;
            if (it.s == 3) {
                // nothing to release: the partials carry their own flags (loop_merge); the arrival only keeps the
                // publisher's ring in step
                if (lane == 0) mbar_arrive_relaxed_cluster(pend);
                pend = 0;
            } else if (!defer) {
                if (lane == 0) mbar_arrive_release_cluster(pend);
                pend = 0;
            }
            ++n_done;
            if (warp == 2 && lane == 0) LOOP_TRACE(7);
        }
        publish_pending(pend);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, kTmemCols);
    }
}

// read at every call (not cached): tools/loop_check.py flips it inside one process for the A/B
static bool loop_env_on() {
    const char *e = getenv("DCAP_GREEDY_LOOP");
    return !(e && atoi(e) == 0);
}

// Measured at F = 1024 (8000 RoIs): folding costs 16 more k-blocks in two stages (+36 % tensor work per step) and is
// slower than re-reading the hoisted fp32 terms (3.56 vs 3.38 ms); off unless DCAP_LOOP_FOLD=1.
bool Decoder::greedy_loop_folds() const {
    const char *e = getenv("DCAP_LOOP_FOLD");
    return e && atoi(e) != 0 && cfg.feat % kBlockK == 0;
}

// DCAP_HOIST_MERGED=0: the two hoisted terms from two GEMMs into two blocked-32 arrays (debugging aid)
bool Decoder::hoist_merged() {
    const char *e = getenv("DCAP_HOIST_MERGED");
    return !(e && atoi(e) == 0);
}

// Can the kernel run here at all?  One probe per device: shared memory opt-in + at least one resident cluster of 2.
static bool loop_launchable() {
    static std::atomic<unsigned long long> probed{0}, usable{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(probed.load(std::memory_order_acquire) & bit)) {
        bool ok = true;
        for (int v = 0; v < 2 && ok; ++v) {
            auto kern = v ? greedy_loop_kernel<true> : greedy_loop_kernel<false>;
            ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem2::kBaseBytes) == cudaSuccess &&
                 cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0) == cudaSuccess;
            if (!ok) break;
            cudaLaunchConfig_t c = {};
            c.blockDim = dim3(kLoopThreads); c.gridDim = dim3(2 * (sm_count() / 2 > 0 ? sm_count() / 2 : 1));
            c.dynamicSmemBytes = TcSmem2::kBaseBytes;
            cudaLaunchAttribute a[1];
            a[0].id = cudaLaunchAttributeClusterDimension;
            a[0].val.clusterDim.x = 2; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
            c.attrs = a; c.numAttrs = 1;
            int fit = 0;
            ok = cudaOccupancyMaxActiveClusters(&fit, kern, &c) == cudaSuccess && fit >= 1;
        }
        if (!ok) cudaGetLastError();
        if (ok) usable.fetch_or(bit, std::memory_order_release);
        probed.fetch_or(bit, std::memory_order_release);
    }
    return (usable.load(std::memory_order_acquire) & bit) != 0;
}

// Used at every batch size: measured (ms per call, loop vs launch per GEMM) 0.71 vs 0.75 (1 RoI), 0.69 vs 0.76 (37),
// 0.79 vs 0.94 (300), 0.95 vs 1.09 (1000), 0.89 vs 1.26 (1300), 1.11 vs 1.70 (2500), 1.56 vs 2.10 (4000), 2.93 vs 3.73 (8000).
bool Decoder::greedy_loop_ok(int B) const {
    (void)B;
    if (!loop_env_on() || !bf || cfg.arch != DC_ARCH_V1) return false;
    const int U = cfg.units;
    return U % 64 == 0 && (4 * U) % 256 == 0 && bf->Epad % 64 == 0 && cfg.vocab >= 256 && cfg.padding >= 1 && loop_launchable();
}

// Steps 0..P-1 of the greedy loop after head / hoisted terms / state reset; ws.tok holds <start>.
int Decoder::greedy_loop_bf16(int B, int32_t *tokens, float *scores, cudaStream_t s) {
    Bf16State &b = *bf;
    const int U = cfg.units, V = cfg.vocab, P = cfg.padding, K1 = b.Epad + U;
    LoopParams p = {};
    p.R = B; p.P = P; p.V = V; p.U = U; p.Epad = b.Epad; p.K1 = K1;
    p.tiles_m = ceil_div(B, 256);
    p.n128 = 2 * p.tiles_m;
    const int Ns[4] = {4 * U, 4 * U, kDense, V}, Ks[4] = {K1, 2 * U, U, kDense};
    int first = 0;
    for (int i = 0; i < 4; ++i) {
        p.tiles_n[i] = ceil_div(Ns[i], 256);
        p.num_kb[i] = ceil_div(Ks[i], kBlockK);
        p.first[i] = first;
        first += p.tiles_n[i];
    }
    p.tiles_n[4] = 1; p.num_kb[4] = 0; p.first[4] = first; first += 1;   // the merge item
    p.first[5] = first;                                    // items per slot
    // wavefront skews (slots): each must cover its producer stage's latency (tile + epilogue + publish, 11-15 us;
    // a slot of ~60 items is ~4.5 us on 74 pairs), and skew[4] < tiles_m (see decode_item)
    // Stage gaps: 5 / 7 / 6 / 5 slots when there are enough row blocks (>= 29), else the row blocks of a step are spread
    // evenly over the five stages (gap = tiles_m / 5), so that even a few row blocks sit in DIFFERENT stages at any time
    // instead of queueing up behind one another in the same one.  DCAP_LOOP_SKEW=n caps the last skew.
    static const int skew_env = getenv("DCAP_LOOP_SKEW") ? atoi(getenv("DCAP_LOOP_SKEW")) : 23;
    static const int full[5] = {0, 5, 12, 18, 23};
    for (int i = 0; i < 5; ++i) {
        int sk = i * p.tiles_m / 5;
        const int cap = skew_env >= 23 ? full[i] + (i == 4 ? skew_env - 23 : 0) : full[i] * skew_env / 23;
        p.skew[i] = sk < cap ? sk : cap;
    }
    const int sk4 = p.skew[4];
    const long long total_ll = ((long long)P * p.tiles_m + sk4) * first;
    DC_REQUIRE(total_ll < (1ll << 31), "greedy loop: too many work items");
    p.total = (int)total_ll;
    p.slots = 2 * p.tiles_n[3];
    // counters (+ error word), zeroed before every launch
    const size_t cnt_bytes = sizeof(unsigned) * (8 * (size_t)p.n128 + 4);
    DC_REQUIRE(b.loop_cnt && cnt_bytes <= b.loop_cnt_bytes, "greedy loop: counter buffer not reserved");
    // A watchdog hit of an EARLIER launch on this handle (its error word is copied to pinned memory behind every launch)
    // is reported now: that call returned garbage tokens.
    if (!b.loop_err_host) {
        DC_CHECK_CUDA(cudaHostAlloc((void **)&b.loop_err_host, sizeof(unsigned), cudaHostAllocDefault));
        *b.loop_err_host = 0;
    }
    if (const unsigned word = *reinterpret_cast<volatile unsigned *>(b.loop_err_host)) {
        *b.loop_err_host = 0;
        return set_error(DC_ERR_CUDA, "an earlier greedy-loop launch on this decoder hit its watchdog (code 0x%x): its tokens are invalid "
                                      "(another persistent kernel on the same GPU, or a protocol fault)", word);
    }
    DC_CHECK_CUDA(cudaMemsetAsync(b.loop_cnt, 0, cnt_bytes, s));
    // the partials carry step tags (flag in data, see loop_merge): no stale tag of an earlier launch may survive
    DC_CHECK_CUDA(cudaMemsetAsync(b.partial, 0, sizeof(float4) * (size_t)p.slots * B, s));
    p.cnt = b.loop_cnt;
    LoopMaps maps;
    int rc = 0;
    rc |= make_tmap_bf16(&maps.m[kMapX1a], b.X1[0], B, K1, K1, 128);
    rc |= make_tmap_bf16(&maps.m[kMapX1b], b.X1[1], B, K1, K1, 128);
    rc |= make_tmap_bf16(&maps.m[kMapX2a], b.X2[0], B, 2 * U, 2 * U, 128);
    rc |= make_tmap_bf16(&maps.m[kMapX2b], b.X2[1], B, 2 * U, 2 * U, 128);
    rc |= make_tmap_bf16(&maps.m[kMapH2a], b.X2[0] + U, B, U, 2 * U, 128);
    rc |= make_tmap_bf16(&maps.m[kMapH2b], b.X2[1] + U, B, U, 2 * U, 128);
    rc |= make_tmap_bf16(&maps.m[kMapD], b.d, B, kDense, kDense, 128);
    rc |= make_tmap_bf16(&maps.m[kMapW1], b.w1cat, 4 * U, K1, K1, 128);
    rc |= make_tmap_bf16(&maps.m[kMapW2], b.w2cat, 4 * U, 2 * U, 2 * U, 128);
    rc |= make_tmap_bf16(&maps.m[kMapWd1], b.wd1h, kDense, U, U, 128);
    rc |= make_tmap_bf16(&maps.m[kMapWd2], b.wd2, V, kDense, kDense, 128);
    const int F = cfg.feat;
    rc |= make_tmap_bf16(&maps.m[kMapF], b.Fb, B, F, F, 128);
    rc |= make_tmap_bf16(&maps.m[kMapW1f], b.w1f, 4 * U, F, F, 128);
    rc |= make_tmap_bf16(&maps.m[kMapWd1f], b.wd1f, kDense, F, F, 128);
    if (rc) return rc;
    // Optional (DCAP_LOOP_FOLD=1): feature terms folded into the contraction, [emb | h1 | f] . [W1e ; U1 ; W1f]^T and
    // [h2 | f] . [Wd1h ; Wd1f]^T in ONE accumulator, the feature k-blocks coming from the bf16 head output -- instead of
    // re-reading the hoisted fp32 terms (97 MB per step at 8000 RoIs) in the epilogues.
    p.fold = greedy_loop_folds() ? 1 : 0;
    for (int i = 0; i < 4; ++i) { p.kb_main[i] = p.num_kb[i]; p.map_b2[i] = kMapW1f; }
    if (p.fold) {
        p.num_kb[0] += F / kBlockK; p.map_b2[0] = kMapW1f;
        p.num_kb[2] += F / kBlockK; p.map_b2[2] = kMapWd1f;
    }
    p.b1 = b.b1_i; p.bd1 = W("imgcap_lstm_d1/bias");
    static const int look_env = getenv("DCAP_LOOP_LOOKAHEAD") ? atoi(getenv("DCAP_LOOP_LOOKAHEAD")) : 1;
    p.lookahead = look_env;
    static const int pfence_env = getenv("DCAP_LOOP_PFENCE") ? atoi(getenv("DCAP_LOOP_PFENCE")) : 2;
    p.pfence = pfence_env;

    // step t (parity t & 1): LSTM1 reads X1[par]; LSTM2 reads X2[par]; Dense(1024) reads the h2 half of X2[par ^ 1]
    p.map_a[0][0] = kMapX1a; p.map_a[0][1] = kMapX1b; p.map_b[0] = kMapW1;
    p.map_a[1][0] = kMapX2a; p.map_a[1][1] = kMapX2b; p.map_b[1] = kMapW2;
    p.map_a[2][0] = kMapH2b; p.map_a[2][1] = kMapH2a; p.map_b[2] = kMapWd1;
    p.map_a[3][0] = kMapD;   p.map_a[3][1] = kMapD;   p.map_b[3] = kMapWd2;
    // [g1f | d1f] = the merged hoist GEMM's output (v1_hoist_merged_bf16): d1f starts at column 4U = float4 column U
    if (hoist_merged()) {
        p.g1f_ld4 = p.d1f_ld4 = (4 * U + kDense) / 4;
        p.g1f = b.hoist_all; p.d1f = b.hoist_all + (size_t)U * 32 * 4;
    } else {
        p.g1f_ld4 = U; p.d1f_ld4 = kDense / 4;
        p.g1f = ws.g1f; p.d1f = ws.d1f;
    }
    p.b2 = b.b2_i; p.bias_v = W("imgcap_lstm_d2/bias");
    p.c1 = ws.c1; p.c2 = ws.c2;
    p.X1[0] = b.X1[0]; p.X1[1] = b.X1[1]; p.X2[0] = b.X2[0]; p.X2[1] = b.X2[1]; p.d = b.d;
    p.partial = reinterpret_cast<float4 *>(b.partial);
    p.emb = b.emb; p.tok = ws.tok; p.tokens = tokens; p.scores = scores;
    static const int wpf = getenv("DCAP_LOOP_WPF") ? atoi(getenv("DCAP_LOOP_WPF")) : 0;
    p.writer_proxy_fence = wpf;
    // <start> embedding rows of step 0
    if (int rc2 = embed_gather(W("imgcap_embedding_layer/embeddings"), ws.tok, B, cfg.embed, V, b.X1[0], K1, true, s)) return rc2;

    using S = TcSmem2;
    auto kern = scores ? greedy_loop_kernel<true> : greedy_loop_kernel<false>;
    static std::atomic<unsigned long long> attr_set[2];
    DC_CHECK_CUDA(once_per_device(attr_set[scores ? 1 : 0], [&] {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBaseBytes);
        return e != cudaSuccess ? e : cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
    }));
    cudaLaunchConfig_t cfgl = {};
    cfgl.blockDim = dim3(kLoopThreads);
    cfgl.dynamicSmemBytes = S::kBaseBytes; cfgl.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfgl.attrs = attr; cfgl.numAttrs = 1;
    // every pair must be resident at once (the items wait for each other): ask how many clusters fit
    int pairs = sm_count() / 2;
    cfgl.gridDim = dim3(2 * pairs);
    int fit = 0;
    if (cudaOccupancyMaxActiveClusters(&fit, kern, &cfgl) == cudaSuccess && fit > 0 && fit < pairs) pairs = fit;
    else cudaGetLastError();
    static const int pairs_env = getenv("DCAP_LOOP_PAIRS") ? atoi(getenv("DCAP_LOOP_PAIRS")) : 0;
    if (pairs_env > 0 && pairs_env < pairs) pairs = pairs_env;
    const int total = p.total;
    if (pairs > total) pairs = total;
    {
        // deferred publication (see publish_pending) needs every dependency to reach further back than one stride of a
        // pair through the item list: (smallest stage gap - 1) slots + 1 item > pairs
        int gap = p.skew[1];
        for (int i = 2; i <= 4; ++i) gap = gap < p.skew[i] - p.skew[i - 1] ? gap : p.skew[i] - p.skew[i - 1];
        gap = gap < p.tiles_m - p.skew[4] ? gap : p.tiles_m - p.skew[4];
        static const int defer_env = getenv("DCAP_LOOP_DEFER") ? atoi(getenv("DCAP_LOOP_DEFER")) : 1;
        p.defer = (defer_env && (gap - 1) * first + 1 > pairs) ? 1 : 0;
    }
    cfgl.gridDim = dim3(2 * pairs);
    const char *trace_path = getenv("DCAP_LOOP_TRACE");
    static unsigned long long *trace_buf = nullptr;
    const int trace_items = ceil_div(total, pairs);
    const size_t trace_bytes = sizeof(unsigned long long) * 12 * (size_t)trace_items * pairs;
    if (trace_path) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusNone) {
            if (trace_buf) cudaFree(trace_buf);
            DC_CHECK_CUDA(cudaMalloc((void **)&trace_buf, trace_bytes));
            DC_CHECK_CUDA(cudaMemsetAsync(trace_buf, 0, trace_bytes, s));
            p.trace = trace_buf; p.trace_items = trace_items;
        }
    }
    DC_CHECK_CUDA(cudaLaunchKernelEx(&cfgl, kern, maps, p));
    DC_CHECK_CUDA(cudaMemcpyAsync(b.loop_err_host, b.loop_cnt + 8 * p.n128, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    b.parity = P & 1;
    if (p.trace) {
        // debugging aid: dump the marks as int32 header {pairs, items per pair, steps, items per slot, first[1..4], tiles_m, skew[1..4], 0, 0, 0} + uint64 data
        std::vector<unsigned long long> host(trace_bytes / sizeof(unsigned long long));
        DC_CHECK_CUDA(cudaStreamSynchronize(s));
        DC_CHECK_CUDA(cudaMemcpy(host.data(), trace_buf, trace_bytes, cudaMemcpyDeviceToHost));
        if (FILE *fp = fopen(trace_path, "wb")) {
            const int hdr[16] = {pairs, trace_items, P, first, p.first[1], p.first[2], p.first[3], p.first[4],
                                 p.tiles_m, p.skew[1], p.skew[2], p.skew[3], p.skew[4], 0, 0, 0};
            fwrite(hdr, sizeof(int), 16, fp);
            fwrite(host.data(), 1, trace_bytes, fp);
            fclose(fp);
        }
    }
    if (getenv("DCAP_LOOP_DEBUG")) {
        // debugging aid (never inside graph capture): drain the stream and report the watchdog word
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusNone) {
            unsigned word = 0;
            DC_CHECK_CUDA(cudaStreamSynchronize(s));
            DC_CHECK_CUDA(cudaMemcpy(&word, b.loop_cnt + 8 * p.n128, sizeof(word), cudaMemcpyDeviceToHost));
            if (word) return set_error(DC_ERR_CUDA, "greedy loop kernel: watchdog code 0x%x (pairs %d, items %d)", word, pairs, total);
        }
    }
    return DC_OK;
}

}  // namespace dcap

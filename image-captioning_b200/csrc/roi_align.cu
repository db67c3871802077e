// PyramidROIAlign for sm_100a: FPN level assignment + bilinear crop-and-resize + final ordering
// in one pass.  Replaces PyramidROIAlign.call
// (/root/reference/evaluate_models/modified_dense_model.py:343-416).
//
// HBM-bound gather.  Layout facts the kernel is built on:
//   * maps are NHWC, so one bilinear tap is `channels` contiguous floats (1 KB at C=256):
//     a tap is read by C/4 consecutive lanes with one 128-bit LDG each -> fully coalesced;
//   * one RoI's output row (pool_h*pool_w*C floats, 50 176 B at 7x7x256) is contiguous and is
//     written once, in its final (image, box) slot, with 128-bit streaming stores -- the
//     reference's concat + top_k re-sort + gather (an extra read+write of the whole output)
//     does not exist here;
//   * a small first pass (one CTA per image) turns every box into a sampling record and
//     counting-sorts the image's RoIs by (level, Morton tile of the box centre); the streaming
//     pass walks that order, so overlapping boxes are processed together and every touched
//     pixel crosses HBM about once while its re-uses hit the 126 MB L2.
//
// Arithmetic is bit-identical to the CPU oracle: every fp32 op is individually rounded
// (__fmul_rn/__fadd_rn/__fsub_rn are never contracted into FMA) and tf.log is the correctly
// rounded fp32 log (fp64 log rounded once).
#include "common.cuh"
#include <math.h>
#include <limits.h>
#include <stdlib.h>

namespace dcap {

constexpr int kMaxPool = 32;

// modified_dense_model.py:351-363.  NaN / +-inf / out-of-range follow x86 cvttss2si (INT_MIN),
// which is what the reference's TF CPU cast produces; CUDA's own cvt would give 0 for NaN.
__device__ __forceinline__ int fpn_level_dev(float y1, float x1, float y2, float x2, float denom) {
    const float h = __fsub_rn(y2, y1);
    const float w = __fsub_rn(x2, x1);
    const float s = __fsqrt_rn(__fmul_rn(h, w));
    const float q = __fdiv_rn(s, denom);
    const float l = (float)log((double)q);                 // correctly rounded fp32 log
    const float r = __fdiv_rn(l, 0.693147180559945309417f);
    const float rr = rintf(r);                             // half to even
    const int iv = (rr >= -2147483648.0f && rr < 2147483648.0f) ? (int)rr : INT_MIN;
    long long lv = 4ll + (long long)iv;
    lv = lv < 2 ? 2 : lv;
    lv = lv > 5 ? 5 : lv;
    return (int)lv;
}

__global__ void fpn_levels_kernel(const float *__restrict__ boxes, long long n, float denom,
                                  int32_t *__restrict__ levels) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = __ldg(reinterpret_cast<const float4 *>(boxes) + i);
    levels[i] = fpn_level_dev(b.x, b.y, b.z, b.w, denom);
}

// ---------------------------------------------------------------------------------------------
// Pass 1 (one CTA per image): locality order + per-RoI sampling records, stored IN SORTED ORDER.
//
// Record of sorted position s (16-byte entries, `1 + ph + pw` of them):
//   [0]        {map base pointer of the RoI's image at its level (64 bit), RoI index, level}
//   [1+y]      {top*W*c4, bottom*W*c4, bits(y_lerp), in_range}   for each of the ph sample rows
//   [1+ph+x]   {left*c4,  right*c4,    bits(x_lerp), in_range}   for each of the pw sample columns
// (offsets in float4 units) so that the streaming pass needs no per-RoI arithmetic, no shared
// memory and no block barrier.
//
// Order: RoIs of an image are counting-sorted by (level, Morton code of the box centre on a
// 16x16 grid).  Overlapping boxes are then processed close together in time, every feature-map
// pixel is fetched from HBM about once and re-used out of L2 (the natural order re-read 1.75x
// the compulsory bytes at the cfg2 size).  Results are still written to the RoI's own slot.
// ---------------------------------------------------------------------------------------------
struct RoiPrepParams {
    const float *boxes;
    const float *fm[4];
    int fm_h[4];
    int fm_w[4];
    int n_boxes;          // per image
    int c4;               // channels / 4
    int ph, pw;
    float denom;          // 224 / sqrt(img_h*img_w), fp32
    int4 *records;        // [total][1 + ph + pw], sorted order
    int4 *scratch;        // [total] {key, rank, level, -}
    int32_t *levels;      // optional
};

constexpr int kBuckets = 1024;      // 4 levels x 256 Morton tiles
constexpr int kPrepThreads = 1024;

__device__ __forceinline__ unsigned morton4(unsigned v) {   // spread 4 bits: abcd -> 0a0b0c0d
    v &= 0xF;
    v = (v | (v << 2)) & 0x33;
    v = (v | (v << 1)) & 0x55;
    return v;
}

// kSmemCache: the image's boxes and their (key, rank, level) stay in dynamic shared memory between the two
// phases (32 B per box) instead of a global scratch round trip per record entry -- the second phase was
// latency-bound on it (18 -> ~7 us per launch at 1000 boxes per image).
template <bool kSmemCache>
__global__ void __launch_bounds__(kPrepThreads) roi_prepare_kernel(const RoiPrepParams p) {
    __shared__ int s_hist[kBuckets];
    __shared__ int s_warp_sum[kPrepThreads / 32];
    extern __shared__ float4 s_dyn[];                      // [n_boxes] boxes, then [n_boxes] int4 {key, rank, level, -}
    float4 *s_box = s_dyn;
    int4 *s_kr = reinterpret_cast<int4 *>(s_dyn + (kSmemCache ? p.n_boxes : 0));
    const int tid = threadIdx.x;
    const long long img = blockIdx.x;
    const int rec_len = 1 + p.ph + p.pw;
    for (int i = tid; i < kBuckets; i += kPrepThreads) s_hist[i] = 0;
    __syncthreads();

    // phase 1: level (the only fp64 work), locality key, rank inside the bucket
    for (int i = tid; i < p.n_boxes; i += kPrepThreads) {
        const long long roi = img * p.n_boxes + i;
        const float4 box = __ldg(reinterpret_cast<const float4 *>(p.boxes) + roi);
        const float y1 = box.x, x1 = box.y, y2 = box.z, x2 = box.w;
        const int lv = fpn_level_dev(y1, x1, y2, x2, p.denom);
        if (p.levels) p.levels[roi] = lv;
        float cy = 0.5f * (y1 + y2), cx = 0.5f * (x1 + x2);
        cy = (cy >= 0.f && cy <= 1.f) ? cy : 0.f;       // also maps NaN to 0
        cx = (cx >= 0.f && cx <= 1.f) ? cx : 0.f;
        const unsigned ty = min(15u, (unsigned)(cy * 16.f)), tx = min(15u, (unsigned)(cx * 16.f));
        const int key = (lv - 2) * 256 + (int)((morton4(ty) << 1) | morton4(tx));
        const int rank = atomicAdd(&s_hist[key], 1);
        if constexpr (kSmemCache) { s_box[i] = box; s_kr[i] = make_int4(key, rank, lv, 0); }
        else p.scratch[roi] = make_int4(key, rank, lv, 0);
    }
    __syncthreads();
    // exclusive scan of the 1024 bucket counts (one bucket per thread)
    {
        const int v = s_hist[tid];
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((tid & 31) >= d) inc += t;
        }
        if ((tid & 31) == 31) s_warp_sum[tid >> 5] = inc;
        __syncthreads();
        if (tid < 32) {
            int w = s_warp_sum[tid];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, d);
                if (tid >= d) w += t;
            }
            s_warp_sum[tid] = w;
        }
        __syncthreads();
        const int warp_off = (tid >> 5) ? s_warp_sum[(tid >> 5) - 1] : 0;
        s_hist[tid] = warp_off + inc - v;
    }
    __syncthreads();
    // phase 2: sampling records at the sorted slots.  One thread per record ENTRY, so that a warp
    // writes 32 consecutive 16-byte entries (coalesced) instead of 32 scattered records.
    const int entries = p.n_boxes * rec_len;
    for (int idx = tid; idx < entries; idx += kPrepThreads) {
        const int i = idx / rec_len;
        const int e = idx - i * rec_len;
        const long long roi = img * p.n_boxes + i;
        const int4 kr = kSmemCache ? s_kr[i] : p.scratch[roi];
        const float4 box = kSmemCache ? s_box[i] : __ldg(reinterpret_cast<const float4 *>(p.boxes) + roi);
        const int lv = kr.z, li = lv - 2;
        int H, W;
        const float *base;
        switch (li) {
            case 0: H = p.fm_h[0]; W = p.fm_w[0]; base = p.fm[0]; break;
            case 1: H = p.fm_h[1]; W = p.fm_w[1]; base = p.fm[1]; break;
            case 2: H = p.fm_h[2]; W = p.fm_w[2]; base = p.fm[2]; break;
            default: H = p.fm_h[3]; W = p.fm_w[3]; base = p.fm[3]; break;
        }
        int4 *rec = p.records + (img * p.n_boxes + s_hist[kr.x] + kr.y) * rec_len;
        int4 v;
        if (e == 0) {
            base += img * (long long)H * W * p.c4 * 4;
            const unsigned long long bp = (unsigned long long)base;
            v = make_int4((int)(bp & 0xffffffffull), (int)(bp >> 32), (int)roi, lv);
        } else {
            // tf.image.crop_and_resize sample coordinate (one sample per bin, end points inclusive)
            const bool is_y = e <= p.ph;
            const int j = is_y ? e - 1 : e - 1 - p.ph;
            const int n = is_y ? p.ph : p.pw;
            const float a1 = is_y ? box.x : box.y, a2 = is_y ? box.z : box.w;
            const float Dm1 = (float)((is_y ? H : W) - 1);
            const int stride = is_y ? W * p.c4 : p.c4;
            float in;
            if (n > 1) {
                const float sc = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), Dm1), (float)(n - 1));
                in = __fadd_rn(__fmul_rn(a1, Dm1), __fmul_rn((float)j, sc));
            } else {
                in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), Dm1);
            }
            const bool ok = (in >= 0.0f) && (in <= Dm1);
            const float fl = floorf(in);
            v = make_int4(ok ? (int)fl * stride : 0, ok ? (int)ceilf(in) * stride : 0,
                          __float_as_int(__fsub_rn(in, fl)), ok ? 1 : 0);
        }
        rec[e] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 2: streaming gather.  One CTA walks sorted RoIs (grid-stride); warp w owns sample rows
// w, w+W, ...  For a row the warp loads the y entry once and then, per 32-quad channel part,
// streams the pw bins in groups of kUnroll: 4 x LDG.128 taps per bin (each 512 B contiguous per
// warp), two-level lerp, one streaming STG.128.  4*kUnroll independent 128-bit loads are in
// flight per lane; there is no shared memory and no block barrier.
// ---------------------------------------------------------------------------------------------
struct RoiStreamParams {
    const int4 *records;
    int c4;
    int parts;            // ceil(c4 / 32)
    int ph, pw;
    void *out;
    int total;            // RoIs
};

__device__ __forceinline__ float lerp_rn(float a, float b, float t) {
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

__device__ __forceinline__ float4 bilerp4(const float4 &tl, const float4 &tr, const float4 &bl,
                                          const float4 &br, float lx, float ly) {
    float4 o;
    o.x = lerp_rn(lerp_rn(tl.x, tr.x, lx), lerp_rn(bl.x, br.x, lx), ly);
    o.y = lerp_rn(lerp_rn(tl.y, tr.y, lx), lerp_rn(bl.y, br.y, lx), ly);
    o.z = lerp_rn(lerp_rn(tl.z, tr.z, lx), lerp_rn(bl.z, br.z, lx), ly);
    o.w = lerp_rn(lerp_rn(tl.w, tr.w, lx), lerp_rn(bl.w, br.w, lx), ly);
    return o;
}

template <bool kBf16>
__device__ __forceinline__ void store_out(void *out, long long idx4, const float4 &v) {
    if constexpr (kBf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t *>(&lo);
        u.y = *reinterpret_cast<uint32_t *>(&hi);
        __stcs(reinterpret_cast<uint2 *>(out) + idx4, u);
    } else {
        __stcs(reinterpret_cast<float4 *>(out) + idx4, v);
    }
}

template <int kUnroll, int kMaxReg, bool kBf16, bool kDedupe = false>
__global__ void __maxnreg__(kMaxReg)
roi_align_stream_kernel(const RoiStreamParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int rec_len = 1 + p.ph + p.pw;
    const int bins = p.ph * p.pw;
    pdl_wait();                       // launched with programmatic dependent launch right behind roi_prepare_kernel

    for (int spos = blockIdx.x; spos < p.total; spos += gridDim.x) {
        const int4 *rec = p.records + (long long)spos * rec_len;
        const int4 hd = __ldg(rec);
        const float4 *fm = reinterpret_cast<const float4 *>(
            ((unsigned long long)(unsigned)hd.x) | ((unsigned long long)(unsigned)hd.y << 32));
        const long long out_roi = (long long)hd.z * bins * p.c4;
        for (int by = warp; by < p.ph; by += nwarp) {
            const int4 ye = __ldg(rec + 1 + by);
            const float ly = __int_as_float(ye.z);
            const float4 *row_t = fm + ye.x;
            const float4 *row_b = fm + ye.y;
            const long long out_row = out_roi + (long long)by * p.pw * p.c4;
            for (int part = 0; part < p.parts; ++part) {
                const int c = part * 32 + lane;
                const bool act = c < p.c4;
                for (int bx0 = 0; bx0 < p.pw; bx0 += kUnroll) {
                    float4 tl[kUnroll], tr[kUnroll], bl[kUnroll], br[kUnroll];
                    float lx[kUnroll];
                    bool ok[kUnroll];
                    int pl = -1, pr = -1;                    // tap offsets of the previous bin of this group (warp-uniform)
#pragma unroll
                    for (int k = 0; k < kUnroll; ++k) {
                        ok[k] = false;
                        if (bx0 + k < p.pw) {
                            const int4 xe = __ldg(rec + 1 + p.ph + bx0 + k);
                            ok[k] = act && ye.w && xe.w;
                            lx[k] = __int_as_float(xe.z);
                            if (ok[k]) {
                                // neighbouring samples less than two pixels apart share a tap column: re-use the registers
                                // (kDedupe: the L2 -> SM tap traffic, not DRAM, is what bounds this kernel)
                                if (kDedupe && k > 0 && ok[k - 1] && xe.x == pr) { tl[k] = tr[k - 1]; bl[k] = br[k - 1]; }
                                else if (kDedupe && k > 0 && ok[k - 1] && xe.x == pl) { tl[k] = tl[k - 1]; bl[k] = bl[k - 1]; }
                                else { tl[k] = __ldg(row_t + xe.x + c); bl[k] = __ldg(row_b + xe.x + c); }
                                if (kDedupe && xe.y == xe.x) { tr[k] = tl[k]; br[k] = bl[k]; }
                                else if (kDedupe && k > 0 && ok[k - 1] && xe.y == pr) { tr[k] = tr[k - 1]; br[k] = br[k - 1]; }
                                else { tr[k] = __ldg(row_t + xe.y + c); br[k] = __ldg(row_b + xe.y + c); }
                                pl = xe.x; pr = xe.y;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kUnroll; ++k) {
                        if (act && bx0 + k < p.pw) {
                            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ok[k]) o = bilerp4(tl[k], tr[k], bl[k], br[k], lx[k], ly);
                            store_out<kBf16>(p.out, out_row + (long long)(bx0 + k) * p.c4 + c, o);
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Round-2 forward path: shared-memory ring fed by bulk async copies (cp.async.bulk -> SASS UBLKCP).
//
// The register-gather kernel above is latency-bound (ncu: long_scoreboard 56 %, DRAM 59 % busy, 43 % of the
// warp slots): every lane waits on its own 1 KB-granular taps.  Here the gather is decoupled from the math:
//
//   * roi_order_kernel (one CTA per image) only computes levels and the locality order -- 8 bytes per RoI.
//     The sampling records of the first design (240 B per RoI, written and re-read through L2) are gone: the
//     producer warp derives them in registers, one sample per lane.
//   * roi_align_ring_kernel, persistent CTAs = 1 producer warp + N consumer warps around a ring of ROW SLOTS in
//     shared memory.  For one RoI the taps live in <= 2*ph distinct map rows, and in every such row in the SAME
//     <= 2*pw distinct pixels (NHWC: one pixel = `channels` contiguous floats).  The producer de-duplicates both
//     (a sample row's bottom row is usually the next sample row's top row; neighbouring samples share pixels),
//     merges adjacent pixels into runs and issues ONE bulk copy per (row, run) into the row's slot; completion
//     is counted on the slot's mbarrier (expect_tx).  A row slot is released by the consumers as soon as the last
//     sample row that reads it is done, so the producer runs ahead by the whole ring (K x 14 KB in flight per CTA
//     -- far more than the 28 warps x 8 LDG.128 the gather kernel could keep outstanding) and across RoI
//     boundaries.  Per RoI the L2 -> SM traffic drops from 196 taps to the ~144 distinct pixels.
//   * consumers: warp = output bin column; per sample row they wait for its (top, bottom) slots, lerp out of
//     shared memory with 128-bit LDS (lane = 4 channels, conflict-free) and write with streaming 128-bit stores.
//     A tiny descriptor ring (slots, lerp weights, release flags) carries the producer's decisions.
// Arithmetic is unchanged (same individually rounded ops), so results stay bit-identical to the oracle.
// ---------------------------------------------------------------------------------------------
struct RoiOrderParams {
    const float *boxes;
    int n_boxes;          // per image
    float denom;
    int2 *order;          // [total] sorted: {roi index, level}
    int2 *scratch;        // [total] {key, rank} when the image does not fit the shared-memory cache
    int32_t *levels;      // optional
};

template <bool kSmemCache>
__global__ void __launch_bounds__(kPrepThreads) roi_order_kernel(const RoiOrderParams p) {
    __shared__ int s_hist[kBuckets];
    __shared__ int s_warp_sum[kPrepThreads / 32];
    extern __shared__ int2 s_kr[];                         // [n_boxes] {key | level << 16, rank}
    const int tid = threadIdx.x;
    const long long img = blockIdx.x;
    for (int i = tid; i < kBuckets; i += kPrepThreads) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < p.n_boxes; i += kPrepThreads) {
        const long long roi = img * p.n_boxes + i;
        const float4 box = __ldg(reinterpret_cast<const float4 *>(p.boxes) + roi);
        const float y1 = box.x, x1 = box.y, y2 = box.z, x2 = box.w;
        const int lv = fpn_level_dev(y1, x1, y2, x2, p.denom);
        if (p.levels) p.levels[roi] = lv;
        float cy = 0.5f * (y1 + y2), cx = 0.5f * (x1 + x2);
        cy = (cy >= 0.f && cy <= 1.f) ? cy : 0.f;       // also maps NaN to 0
        cx = (cx >= 0.f && cx <= 1.f) ? cx : 0.f;
        const unsigned ty = min(15u, (unsigned)(cy * 16.f)), tx = min(15u, (unsigned)(cx * 16.f));
        const int key = (lv - 2) * 256 + (int)((morton4(ty) << 1) | morton4(tx));
        const int rank = atomicAdd(&s_hist[key], 1);
        const int2 kr = make_int2(key | (lv << 16), rank);
        if constexpr (kSmemCache) s_kr[i] = kr; else p.scratch[roi] = kr;
    }
    __syncthreads();
    {
        const int v = s_hist[tid];
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((tid & 31) >= d) inc += t;
        }
        if ((tid & 31) == 31) s_warp_sum[tid >> 5] = inc;
        __syncthreads();
        if (tid < 32) {
            int w = s_warp_sum[tid];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, d);
                if (tid >= d) w += t;
            }
            s_warp_sum[tid] = w;
        }
        __syncthreads();
        const int warp_off = (tid >> 5) ? s_warp_sum[(tid >> 5) - 1] : 0;
        s_hist[tid] = warp_off + inc - v;
    }
    __syncthreads();
    for (int i = tid; i < p.n_boxes; i += kPrepThreads) {
        const long long roi = img * p.n_boxes + i;
        const int2 kr = kSmemCache ? s_kr[i] : p.scratch[roi];
        p.order[img * p.n_boxes + s_hist[kr.x & 0xffff] + kr.y] = make_int2((int)roi, kr.x >> 16);
    }
    pdl_launch_dependents();
}

namespace ring {

constexpr int kDescDepth = 4;
constexpr int kMaxSamples = 16;          // per axis: one producer lane per sample
constexpr int kMaxSlots = 16;

struct Desc {                             // producer -> consumers, one per RoI in flight
    int roi, pad0, pad1, pad2;
    int4 y[kMaxSamples];                  // {top slot, bottom slot, bits(y lerp), flags}
    int4 x[kMaxSamples];                  // {byte offset of the left tap inside a row slot, of the right tap, bits(x lerp), in range}
};
// y flags: bit0 in range, bit1 release top slot after this sample row, bit2 release bottom slot,
//          bit3 / bit4 mbarrier phase parity of the top / bottom slot's current use

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// mode 0: every lane polls; mode 1: lane 0 polls, the warp re-converges behind it
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity, int lane, int one_lane) {
    if (one_lane) {
        if (lane == 0) mbar_wait(bar, parity);
        __syncwarp();
    } else {
        mbar_wait(bar, parity);
    }
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16); completion = complete_tx on `bar`
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace ring

struct RoiRingParams {
    const float *boxes;
    const int2 *order;
    const float *fm[4];
    int fm_h[4];
    int fm_w[4];
    int n_boxes;          // per image
    int c4;               // channels / 4
    int ph, pw;
    void *out;
    int total;
    int slots;            // K row slots in the ring
    unsigned slot_bytes;  // 2*pw pixels
    int diag;             // DCAP_ROI_DIAG (perf triage only, wrong results): 1 = no tap reads / math, 2 = no copies, 4 = no stores
    int sync_mode;        // DCAP_ROI_SYNC bit0: one lane polls the mbarriers, bit1: wait only for rows that are new at this sample
    unsigned long long *prof;   // DCAP_ROI_PROF: 8 cycle counters (producer 0..3, consumer warp 4..7), or null
};

template <bool kBf16>
__global__ void __launch_bounds__(512, 1) roi_align_ring_kernel(const RoiRingParams p) {
    using namespace ring;
    extern __shared__ __align__(128) unsigned char ring_smem[];
    __shared__ int s_row[2 * kMaxSamples];                // producer scratch: map row of every row position of the RoI
    __shared__ int4 s_run[2 * kMaxSamples];               // producer scratch: pixel runs {byte offset in slot, first pixel, bytes}
    const int K = p.slots;
    unsigned char *slots = ring_smem;
    Desc *desc = reinterpret_cast<Desc *>(ring_smem + (size_t)K * p.slot_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(desc + kDescDepth);
    uint64_t *empty = full + K;
    uint64_t *dfull = empty + K;
    uint64_t *dempty = dfull + kDescDepth;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncons = (blockDim.x >> 5) - 1;
    const uint32_t px_bytes = (uint32_t)p.c4 * 16u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < K; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, ncons); }
        for (int i = 0; i < kDescDepth; ++i) { mbar_init(dfull + i, 1); mbar_init(dempty + i, ncons); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                                            // the order array comes from roi_order_kernel

    if (warp == 0) {
        // ------------------------------- producer -------------------------------
        const bool is_y = lane < 16;
        const int j = lane & 15;
        const int n = is_y ? p.ph : p.pw;
        unsigned head = 0;                                 // row positions issued so far (warp-uniform)
        int it = 0;
        int spos = blockIdx.x;
        int2 ord = make_int2(0, 2);
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        if (spos < p.total) {
            ord = __ldg(p.order + spos);
            box = __ldg(reinterpret_cast<const float4 *>(p.boxes) + ord.x);
        }
        unsigned long long pc[4] = {0, 0, 0, 0};
        long long tk = clock64();
        auto tick = [&](int i) { if (p.prof) { const long long now = clock64(); pc[i] += (unsigned long long)(now - tk); tk = now; } };
        for (; spos < p.total; spos += gridDim.x, ++it) {
            const int2 cur = ord;
            const float4 cbox = box;
            if (spos + (long long)gridDim.x < p.total) {   // next RoI's box: in flight while this one is issued
                ord = __ldg(p.order + spos + gridDim.x);
                box = __ldg(reinterpret_cast<const float4 *>(p.boxes) + ord.x);
            }
            const int roi = cur.x;
            int H, W;
            const float *base;
            switch (cur.y - 2) {
                case 0: H = p.fm_h[0]; W = p.fm_w[0]; base = p.fm[0]; break;
                case 1: H = p.fm_h[1]; W = p.fm_w[1]; base = p.fm[1]; break;
                case 2: H = p.fm_h[2]; W = p.fm_w[2]; base = p.fm[2]; break;
                default: H = p.fm_h[3]; W = p.fm_w[3]; base = p.fm[3]; break;
            }
            base += (long long)(roi / p.n_boxes) * H * W * p.c4 * 4;
            // tf.image.crop_and_resize sample coordinate of this lane (one sample per bin, end points inclusive)
            const float a1 = is_y ? cbox.x : cbox.y, a2 = is_y ? cbox.z : cbox.w;
            const float Dm1 = (float)((is_y ? H : W) - 1);
            float in;
            if (n > 1) {
                const float sc = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), Dm1), (float)(n - 1));
                in = __fadd_rn(__fmul_rn(a1, Dm1), __fmul_rn((float)j, sc));
            } else {
                in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), Dm1);
            }
            const bool ok = (j < n) && (in >= 0.0f) && (in <= Dm1);
            const float fl = floorf(in);
            const int lo = ok ? (int)fl : 0, hi = ok ? (int)ceilf(in) : 0;
            const float frac = __fsub_rn(in, fl);

            // distinct rows (y half) / pixels (x half), in sample order; a tap already loaded for the previous
            // sample is re-used.  All lanes of a half walk the same broadcast values -> uniform results.
            int posLo = 0, posHi = 0, cnt = 0;
            bool newLo = false, newHi = false, runStart = false;
            {
                int pLo = 0, pHi = 0, pPosLo = 0, pPosHi = 0, lastNew = INT_MIN;
                bool pOk = false;
                const int ns = p.ph > p.pw ? p.ph : p.pw;
                for (int s = 0; s < ns; ++s) {
                    const int sLo = __shfl_sync(0xffffffffu, lo, s, 16), sHi = __shfl_sync(0xffffffffu, hi, s, 16);
                    const bool sOk = __shfl_sync(0xffffffffu, (int)ok, s, 16) != 0;
                    int qLo = 0, qHi = 0;
                    bool nLo = false, nHi = false, rs = false;
                    if (sOk) {
                        if (pOk && sLo == pLo) qLo = pPosLo;
                        else if (pOk && sLo == pHi) qLo = pPosHi;
                        else { qLo = cnt++; nLo = true; rs = (sLo != lastNew + 1) || lastNew == INT_MIN; lastNew = sLo; }
                        if (sHi == sLo) qHi = qLo;
                        else if (pOk && sHi == pLo) qHi = pPosLo;
                        else if (pOk && sHi == pHi) qHi = pPosHi;
                        else {
                            qHi = cnt++; nHi = true;
                            if (!nLo) rs = (sHi != lastNew + 1) || lastNew == INT_MIN;
                            lastNew = sHi;
                        }
                    }
                    pOk = sOk; pLo = sLo; pHi = sHi; pPosLo = qLo; pPosHi = qHi;
                    if (j == s) { posLo = qLo; posHi = qHi; newLo = nLo; newHi = nHi; runStart = rs && (nLo || nHi); }
                }
            }
            const int ny = __shfl_sync(0xffffffffu, cnt, 0), nq = __shfl_sync(0xffffffffu, cnt, 16);
            // a row / pixel position dies after the last sample that reads it: the next sample either re-uses it or
            // nobody does (re-use only ever looks one sample back)
            const int nxOk = __shfl_down_sync(0xffffffffu, (int)ok, 1, 16);
            const int nxLo = __shfl_down_sync(0xffffffffu, posLo, 1, 16), nxHi = __shfl_down_sync(0xffffffffu, posHi, 1, 16);
            const bool next_ok = (j < 15) && nxOk;
            const bool relLo = ok && !(next_ok && (posLo == nxLo || posLo == nxHi));
            const bool relHi = ok && posHi != posLo && !(next_ok && (posHi == nxLo || posHi == nxHi));
            // pixel runs (x half): one bulk copy per run and row
            const int runQ = newLo ? posLo : posHi, runPx = newLo ? lo : hi;
            const unsigned start_mask = __reduce_or_sync(0xffffffffu, (!is_y && runStart) ? (1u << runQ) : 0u);
            const unsigned above = (runQ + 1 < 32) ? (start_mask >> (runQ + 1)) : 0u;
            const int runLen = above ? __ffs(above) : nq - runQ;

            // publish the descriptor
            const int di = it % kDescDepth;
            tick(0);
            mbar_wait_warp(dempty + di, (((unsigned)it / kDescDepth) & 1u) ^ 1u, lane, p.sync_mode & 1);
            tick(1);
            Desc *d = desc + di;
            {
                const unsigned PLo = head + (unsigned)posLo, PHi = head + (unsigned)posHi;
                if (is_y) {
                    if (j < p.ph)
                        d->y[j] = make_int4((int)(PLo % K), (int)(PHi % K), __float_as_int(frac),
                                            (ok ? 1 : 0) | (relLo ? 2 : 0) | (relHi ? 4 : 0) | ((int)((PLo / K) & 1u) << 3) |
                                                ((int)((PHi / K) & 1u) << 4) | (newLo ? 32 : 0) | (newHi ? 64 : 0));
                } else if (j < p.pw) {
                    d->x[j] = make_int4(posLo * (int)px_bytes, posHi * (int)px_bytes, __float_as_int(frac), ok ? 1 : 0);
                }
                if (lane == 0) d->roi = roi;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(dfull + di);

            // issue the rows: lane r owns row position r.  Every lane polls its own slot's empty barrier (the slot's
            // previous use was issued K positions ago, so at most one phase is outstanding per slot as long as a batch
            // holds <= K rows), arms the full barrier and fires one bulk copy per pixel run -- no lane waits for another.
            __syncwarp();                                  // previous RoI's table reads are done
            if (is_y) {
                if (newLo) s_row[posLo] = lo;
                if (newHi) s_row[posHi] = hi;
            } else if (runStart) {
                const int ri = __popc(start_mask & ((1u << runQ) - 1u));
                s_run[ri] = make_int4(runQ * (int)px_bytes, runPx, runLen * (int)px_bytes, 0);
            }
            __syncwarp();
            const int nruns = __popc(start_mask);
            tick(2);
            for (int b0 = 0; b0 < ny; b0 += K) {
                const int r = b0 + lane;
                const bool mine = lane < K && r < ny;
                const unsigned P = head + (unsigned)r;
                const int slot = (int)(P % K);
                const uint32_t par = ((P / K) & 1u) ^ 1u;
                const float *src_row = mine ? base + (long long)s_row[r] * W * p.c4 * 4 : nullptr;
                unsigned char *dst_row = slots + (size_t)slot * p.slot_bytes;
                unsigned pending = __ballot_sync(0xffffffffu, mine);
                bool todo = mine;
                while (pending) {
                    bool fired = false;
                    if (todo && mbar_try_wait(empty + slot, par)) {
                        mbar_expect_tx(full + slot, (p.diag & 2) ? 0u : (uint32_t)nq * px_bytes);
                        for (int i = 0; i < ((p.diag & 2) ? 0 : nruns); ++i) {
                            const int4 rn = s_run[i];
                            bulk_load(dst_row + rn.x, src_row + (long long)rn.y * p.c4 * 4, (uint32_t)rn.z, full + slot);
                        }
                        fired = true;
                        todo = false;
                    }
                    pending &= ~__ballot_sync(0xffffffffu, fired);
                }
            }
            head += (unsigned)ny;
            tick(3);
        }
        if (p.prof && lane == 0)
            for (int i = 0; i < 4; ++i) atomicAdd(p.prof + i, pc[i]);
    } else {
        // ------------------------------- consumers -------------------------------
        const int cw = warp - 1;
        const int bins = p.ph * p.pw;
        int it = 0;
        unsigned long long pc[4] = {0, 0, 0, 0};
        long long tk = clock64();
        auto tick = [&](int i) { if (p.prof) { const long long now = clock64(); pc[i] += (unsigned long long)(now - tk); tk = now; } };
        const int one = p.sync_mode & 1, only_new = p.sync_mode & 2;
        for (int spos = blockIdx.x; spos < p.total; spos += gridDim.x, ++it) {
            const int di = it % kDescDepth;
            mbar_wait_warp(dfull + di, ((unsigned)it / kDescDepth) & 1u, lane, one);
            tick(0);
            const Desc *d = desc + di;
            const long long out_roi = (long long)d->roi * bins * p.c4;
            for (int by = 0; by < p.ph; ++by) {
                const int4 ye = d->y[by];
                const bool yok = (ye.w & 1) != 0;
                if (yok) {
                    if (!only_new || (ye.w & 32)) mbar_wait_warp(full + ye.x, (unsigned)(ye.w >> 3) & 1u, lane, one);
                    if (!only_new || (ye.w & 64)) mbar_wait_warp(full + ye.y, (unsigned)(ye.w >> 4) & 1u, lane, one);
                }
                tick(1);
                const unsigned char *top = slots + (size_t)ye.x * p.slot_bytes;
                const unsigned char *bot = slots + (size_t)ye.y * p.slot_bytes;
                const float ly = __int_as_float(ye.z);
                for (int bx = cw; bx < p.pw; bx += ncons) {
                    const int4 xe = d->x[bx];
                    const bool ok = yok && xe.w && !(p.diag & 1);
                    const float lx = __int_as_float(xe.z);
                    const long long o = out_roi + ((long long)by * p.pw + bx) * p.c4;
                    if (p.diag & 4) continue;
                    for (int c0 = 0; c0 < p.c4; c0 += 64) {
                        const int ca = c0 + lane, cb = c0 + 32 + lane;
                        const bool acta = ca < p.c4, actb = cb < p.c4;
                        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                        if (ok) {
                            float4 tl, tr, bl, br, tl2, tr2, bl2, br2;
                            if (acta) {
                                tl = *reinterpret_cast<const float4 *>(top + xe.x + ca * 16);
                                tr = *reinterpret_cast<const float4 *>(top + xe.y + ca * 16);
                                bl = *reinterpret_cast<const float4 *>(bot + xe.x + ca * 16);
                                br = *reinterpret_cast<const float4 *>(bot + xe.y + ca * 16);
                            }
                            if (actb) {
                                tl2 = *reinterpret_cast<const float4 *>(top + xe.x + cb * 16);
                                tr2 = *reinterpret_cast<const float4 *>(top + xe.y + cb * 16);
                                bl2 = *reinterpret_cast<const float4 *>(bot + xe.x + cb * 16);
                                br2 = *reinterpret_cast<const float4 *>(bot + xe.y + cb * 16);
                            }
                            if (acta) va = bilerp4(tl, tr, bl, br, lx, ly);
                            if (actb) vb = bilerp4(tl2, tr2, bl2, br2, lx, ly);
                        }
                        if (acta) store_out<kBf16>(p.out, o + ca, va);
                        if (actb) store_out<kBf16>(p.out, o + cb, vb);
                    }
                }
                __syncwarp();
                tick(2);
                if (lane == 0 && yok) {
                    if (ye.w & 2) mbar_arrive(empty + ye.x);
                    if (ye.w & 4) mbar_arrive(empty + ye.y);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty + di);
            tick(3);
        }
        if (p.prof && lane == 0 && cw == 0)
            for (int i = 0; i < 4; ++i) atomicAdd(p.prof + 4 + i, pc[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// Ring kernel, second form.  Profiling the first one with clock64 marks (DCAP_ROI_PROF) showed its single producer
// warp as the bottleneck: ~4000 dependent-instruction cycles per RoI for the sample arithmetic and the serial
// de-duplication, ~800 to publish the descriptor, ~3400 in the issue loop -- 7700 cycles per RoI against a budget of
// ~4500, while the consumers sat in their row waits.  Here everything that depends only on the BOX moves into a
// massively parallel pre-pass (one warp per RoI, latency irrelevant) that writes a 1 KB "ring record" per RoI:
// sample weights, row / pixel positions, release flags, pixel runs, map rows.  The producer warp only loads the
// record (prefetched one RoI ahead) and issues the bulk copies; the consumer warps read the record themselves
// (L2 / L1 resident) and track the ring head on their own, so the shared-memory descriptor ring and its two
// barrier sets are gone.  Slot arithmetic is division free (the head is kept as slot + phase).
// ---------------------------------------------------------------------------------------------
namespace ring2 {
constexpr int kRec = 64;                                   // int4 entries per RoI record (1 KB)
constexpr int kHdr = 0, kHdr2 = 1, kY = 2, kX = 18, kRun = 34, kRow = 50;   // rows: 32 ints = 8 entries
}  // namespace ring2

struct RoiRecParams {
    const float *boxes;
    const int2 *order;
    const float *fm[4];
    int fm_h[4];
    int fm_w[4];
    int n_boxes, c4, ph, pw, total;
    int4 *records;
};

__global__ void __launch_bounds__(256) roi_ring_records_kernel(const RoiRecParams p) {
    using namespace ring2;
    const int lane = threadIdx.x & 31;
    const long long spos = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    pdl_wait();                                            // the order array comes from roi_order_kernel
    pdl_launch_dependents();
    if (spos >= p.total) return;
    const int2 cur = __ldg(p.order + spos);
    const float4 cbox = __ldg(reinterpret_cast<const float4 *>(p.boxes) + cur.x);
    const int roi = cur.x;
    int H, W;
    const float *base;
    switch (cur.y - 2) {
        case 0: H = p.fm_h[0]; W = p.fm_w[0]; base = p.fm[0]; break;
        case 1: H = p.fm_h[1]; W = p.fm_w[1]; base = p.fm[1]; break;
        case 2: H = p.fm_h[2]; W = p.fm_w[2]; base = p.fm[2]; break;
        default: H = p.fm_h[3]; W = p.fm_w[3]; base = p.fm[3]; break;
    }
    base += (long long)(roi / p.n_boxes) * H * W * p.c4 * 4;
    const bool is_y = lane < 16;
    const int j = lane & 15;
    const int n = is_y ? p.ph : p.pw;
    const unsigned px_bytes = (unsigned)p.c4 * 16u;
    // tf.image.crop_and_resize sample coordinate of this lane (one sample per bin, end points inclusive)
    const float a1 = is_y ? cbox.x : cbox.y, a2 = is_y ? cbox.z : cbox.w;
    const float Dm1 = (float)((is_y ? H : W) - 1);
    float in;
    if (n > 1) {
        const float sc = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), Dm1), (float)(n - 1));
        in = __fadd_rn(__fmul_rn(a1, Dm1), __fmul_rn((float)j, sc));
    } else {
        in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), Dm1);
    }
    const bool ok = (j < n) && (in >= 0.0f) && (in <= Dm1);
    const float fl = floorf(in);
    const int lo = ok ? (int)fl : 0, hi = ok ? (int)ceilf(in) : 0;
    const float frac = __fsub_rn(in, fl);
    // distinct rows (y half) / pixels (x half) in sample order; a tap already loaded for the previous sample is re-used
    int posLo = 0, posHi = 0, cnt = 0;
    bool newLo = false, newHi = false, runStart = false;
    {
        int pLo = 0, pHi = 0, pPosLo = 0, pPosHi = 0, lastNew = INT_MIN;
        bool pOk = false;
        const int ns = p.ph > p.pw ? p.ph : p.pw;
        for (int s = 0; s < ns; ++s) {
            const int sLo = __shfl_sync(0xffffffffu, lo, s, 16), sHi = __shfl_sync(0xffffffffu, hi, s, 16);
            const bool sOk = __shfl_sync(0xffffffffu, (int)ok, s, 16) != 0;
            int qLo = 0, qHi = 0;
            bool nLo = false, nHi = false, rs = false;
            if (sOk) {
                if (pOk && sLo == pLo) qLo = pPosLo;
                else if (pOk && sLo == pHi) qLo = pPosHi;
                else { qLo = cnt++; nLo = true; rs = (sLo != lastNew + 1) || lastNew == INT_MIN; lastNew = sLo; }
                if (sHi == sLo) qHi = qLo;
                else if (pOk && sHi == pLo) qHi = pPosLo;
                else if (pOk && sHi == pHi) qHi = pPosHi;
                else {
                    qHi = cnt++; nHi = true;
                    if (!nLo) rs = (sHi != lastNew + 1) || lastNew == INT_MIN;
                    lastNew = sHi;
                }
            }
            pOk = sOk; pLo = sLo; pHi = sHi; pPosLo = qLo; pPosHi = qHi;
            if (j == s) { posLo = qLo; posHi = qHi; newLo = nLo; newHi = nHi; runStart = rs && (nLo || nHi); }
        }
    }
    const int ny = __shfl_sync(0xffffffffu, cnt, 0), nq = __shfl_sync(0xffffffffu, cnt, 16);
    // a row / pixel position dies after the last sample that reads it (re-use only ever looks one sample back)
    const int nxOk = __shfl_down_sync(0xffffffffu, (int)ok, 1, 16);
    const int nxLo = __shfl_down_sync(0xffffffffu, posLo, 1, 16), nxHi = __shfl_down_sync(0xffffffffu, posHi, 1, 16);
    const bool next_ok = (j < 15) && nxOk;
    const bool relLo = ok && !(next_ok && (posLo == nxLo || posLo == nxHi));
    const bool relHi = ok && posHi != posLo && !(next_ok && (posHi == nxLo || posHi == nxHi));
    const int runQ = newLo ? posLo : posHi, runPx = newLo ? lo : hi;
    const unsigned start_mask = __reduce_or_sync(0xffffffffu, (!is_y && runStart) ? (1u << runQ) : 0u);
    const unsigned above = (runQ + 1 < 32) ? (start_mask >> (runQ + 1)) : 0u;
    const int runLen = above ? __ffs(above) : nq - runQ;
    const int nruns = __popc(start_mask);

    int4 *rec = p.records + spos * kRec;
    if (lane == 0) {
        const unsigned long long bp = (unsigned long long)base;
        rec[kHdr] = make_int4((int)(bp & 0xffffffffull), (int)(bp >> 32), roi, ny | (nq << 8) | (nruns << 16));
        rec[kHdr2] = make_int4(W * p.c4, 0, 0, 0);
    }
    if (is_y) {
        rec[kY + j] = make_int4(posLo, posHi, __float_as_int(frac),
                                (ok ? 1 : 0) | (relLo ? 2 : 0) | (relHi ? 4 : 0) | (newLo ? 32 : 0) | (newHi ? 64 : 0));
        int *rows = reinterpret_cast<int *>(rec + kRow);
        if (newLo) rows[posLo] = lo;
        if (newHi) rows[posHi] = hi;
    } else {
        rec[kX + j] = make_int4(posLo * (int)px_bytes, posHi * (int)px_bytes, __float_as_int(frac), ok ? 1 : 0);
        if (runStart) rec[kRun + __popc(start_mask & ((1u << runQ) - 1u))] = make_int4(runQ * (int)px_bytes, runPx * p.c4, runLen * (int)px_bytes, 0);
    }
}

struct RoiRing2Params {
    const int4 *records;
    int c4, ph, pw, total;
    void *out;
    int slots;            // K row slots in the ring
    unsigned slot_bytes;  // 2*pw pixels
    int diag;
    int nprod;            // producer warps (1 or 2): warp q issues the row positions r with r % nprod == q
    int groups;           // consumer warp groups: group g computes the sample rows by with by % groups == g
    int only_new;         // wait only for the rows that are new at this sample (the others were waited for one sample earlier)
    unsigned long long *prof;
};

template <bool kBf16>
__global__ void __launch_bounds__(512, 1) roi_align_ring2_kernel(const RoiRing2Params p) {
    using namespace ring;
    using namespace ring2;
    extern __shared__ __align__(128) unsigned char ring_smem[];
    __shared__ int4 s_run[2][2 * kMaxSamples];             // producer scratch (per producer warp): this RoI's pixel runs
    const int K = p.slots;
    unsigned char *slots = ring_smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring_smem + (size_t)K * p.slot_bytes);
    uint64_t *empty = full + K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nprod = p.nprod;
    const int ncons = (blockDim.x >> 5) - nprod;
    const uint32_t px_bytes = (uint32_t)p.c4 * 16u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < K; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, ncons); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                                            // the records come from roi_ring_records_kernel
    int headSlot = 0, headPhase = 0;                       // ring position of the current RoI's first row (all warps track it)
    unsigned long long pc[4] = {0, 0, 0, 0};
    long long tk = clock64();
    auto tick = [&](int i) { if (p.prof) { const long long now = clock64(); pc[i] += (unsigned long long)(now - tk); tk = now; } };

    if (warp < nprod) {
        // ------------------------------- producers -------------------------------
        int4 *my_runs = s_run[warp];
        int spos = blockIdx.x;
        int4 hdr = make_int4(0, 0, 0, 0), hdr2 = hdr, myrun = hdr;
        int myrow = 0;
        if (spos < p.total) {
            const int4 *rec = p.records + (long long)spos * kRec;
            hdr = __ldg(rec + kHdr); hdr2 = __ldg(rec + kHdr2);
            myrun = __ldg(rec + kRun + (lane & 15));
            myrow = __ldg(reinterpret_cast<const int *>(rec + kRow) + lane);
        }
        for (; spos < p.total; spos += gridDim.x) {
            const int4 chdr = hdr, chdr2 = hdr2, crun = myrun;
            const int crow = myrow;
            if (spos + (long long)gridDim.x < p.total) {   // next RoI's record: in flight while this one is issued
                const int4 *rec = p.records + ((long long)spos + gridDim.x) * kRec;
                hdr = __ldg(rec + kHdr); hdr2 = __ldg(rec + kHdr2);
                myrun = __ldg(rec + kRun + (lane & 15));
                myrow = __ldg(reinterpret_cast<const int *>(rec + kRow) + lane);
            }
            const int ny = chdr.w & 255, nq = (chdr.w >> 8) & 255, nruns = (chdr.w >> 16) & 255;
            const float4 *base = reinterpret_cast<const float4 *>(((unsigned long long)(unsigned)chdr.x) | ((unsigned long long)(unsigned)chdr.y << 32));
            __syncwarp();                                  // the previous RoI's run table is no longer read
            if (lane < 16) my_runs[lane] = crun;
            __syncwarp();
            tick(0);
            for (int b0 = 0; b0 < ny; b0 += K) {
                const int r = b0 + lane * nprod + warp;            // this warp's positions: r % nprod == warp
                const bool mine = lane * nprod + warp < K && r < ny;
                int slot = headSlot + r, wraps = 0;
                while (slot >= K) { slot -= K; ++wraps; }
                const uint32_t par = (uint32_t)((headPhase + wraps) & 1) ^ 1u;
                const int rowidx = __shfl_sync(0xffffffffu, crow, r & 31);
                const float4 *src_row = base + (long long)rowidx * chdr2.x;
                unsigned char *dst_row = slots + (size_t)slot * p.slot_bytes;
                unsigned pending = __ballot_sync(0xffffffffu, mine);
                bool todo = mine;
                while (pending) {
                    bool fired = false;
                    if (todo && mbar_try_wait(empty + slot, par)) {
                        mbar_expect_tx(full + slot, (p.diag & 2) ? 0u : (uint32_t)nq * px_bytes);
                        for (int i = 0; i < ((p.diag & 2) ? 0 : nruns); ++i) {
                            const int4 rn = my_runs[i];
                            bulk_load(dst_row + rn.x, src_row + rn.y, (uint32_t)rn.z, full + slot);
                        }
                        fired = true;
                        todo = false;
                    }
                    pending &= ~__ballot_sync(0xffffffffu, fired);
                }
            }
            headSlot += ny;
            while (headSlot >= K) { headSlot -= K; headPhase ^= 1; }
            tick(1);
        }
        if (p.prof && lane == 0 && warp == 0)
            for (int i = 0; i < 2; ++i) atomicAdd(p.prof + i, pc[i]);
    } else {
        // ------------------------------- consumers -------------------------------
        const int cw_all = warp - nprod;
        const int per_group = ncons / p.groups;                // warps per group
        const int grp = cw_all / per_group, cw = cw_all - grp * per_group;
        const int bins = p.ph * p.pw;
        int spos = blockIdx.x;
        int4 hdr = make_int4(0, 0, 0, 0), ylane = hdr;
        if (spos < p.total) {
            const int4 *rec = p.records + (long long)spos * kRec;
            hdr = __ldg(rec + kHdr);
            ylane = __ldg(rec + kY + (lane & 15));
        }
        for (; spos < p.total; spos += gridDim.x) {
            const int4 chdr = hdr, cy = ylane;
            const int4 *rec = p.records + (long long)spos * kRec;
            if (spos + (long long)gridDim.x < p.total) {
                const int4 *nrec = p.records + ((long long)spos + gridDim.x) * kRec;
                hdr = __ldg(nrec + kHdr);
                ylane = __ldg(nrec + kY + (lane & 15));
            }
            const int ny = chdr.w & 255;
            const long long out_roi = (long long)chdr.z * bins * p.c4;
            // ring slot / phase of row position `lane` of this RoI
            int slotOf = headSlot + lane, wraps = 0;
            while (slotOf >= K) { slotOf -= K; ++wraps; }
            const int parOf = (headPhase + wraps) & 1;
            const bool single = p.pw <= per_group;             // one bin column per warp: its x entry is loaded once per RoI
            const int4 xe0 = (single && cw < p.pw) ? __ldg(rec + kX + cw) : make_int4(0, 0, 0, 0);
            tick(0);
            for (int by = 0; by < p.ph; ++by) {
                const int yPosLo = __shfl_sync(0xffffffffu, cy.x, by), yPosHi = __shfl_sync(0xffffffffu, cy.y, by);
                const float ly = __int_as_float(__shfl_sync(0xffffffffu, cy.z, by));
                const int yflags = __shfl_sync(0xffffffffu, cy.w, by);
                const bool yok = (yflags & 1) != 0;
                const int sTop = __shfl_sync(0xffffffffu, slotOf, yPosLo), sBot = __shfl_sync(0xffffffffu, slotOf, yPosHi);
                const int pTop = __shfl_sync(0xffffffffu, parOf, yPosLo), pBot = __shfl_sync(0xffffffffu, parOf, yPosHi);
                if (yok) {
                    if (!p.only_new || (yflags & 32)) mbar_wait(full + sTop, (uint32_t)pTop);
                    if (!p.only_new || (yflags & 64)) mbar_wait(full + sBot, (uint32_t)pBot);
                }
                tick(1);
                const bool mine_row = (by % p.groups) == grp;      // the other group(s) compute this sample row
                const unsigned char *top = slots + (size_t)sTop * p.slot_bytes;
                const unsigned char *bot = slots + (size_t)sBot * p.slot_bytes;
                for (int bx = mine_row ? cw : p.pw; bx < p.pw; bx += per_group) {
                    const int4 xe = single ? xe0 : __ldg(rec + kX + bx);
                    const bool ok = yok && xe.w && !(p.diag & 1);
                    const float lx = __int_as_float(xe.z);
                    const long long o = out_roi + ((long long)by * p.pw + bx) * p.c4;
                    if (p.diag & 4) continue;
                    for (int c0 = 0; c0 < p.c4; c0 += 64) {
                        const int ca = c0 + lane, cb = c0 + 32 + lane;
                        const bool acta = ca < p.c4, actb = cb < p.c4;
                        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                        if (ok) {
                            float4 tl, tr, bl, br, tl2, tr2, bl2, br2;
                            if (acta) {
                                tl = *reinterpret_cast<const float4 *>(top + xe.x + ca * 16);
                                tr = *reinterpret_cast<const float4 *>(top + xe.y + ca * 16);
                                bl = *reinterpret_cast<const float4 *>(bot + xe.x + ca * 16);
                                br = *reinterpret_cast<const float4 *>(bot + xe.y + ca * 16);
                            }
                            if (actb) {
                                tl2 = *reinterpret_cast<const float4 *>(top + xe.x + cb * 16);
                                tr2 = *reinterpret_cast<const float4 *>(top + xe.y + cb * 16);
                                bl2 = *reinterpret_cast<const float4 *>(bot + xe.x + cb * 16);
                                br2 = *reinterpret_cast<const float4 *>(bot + xe.y + cb * 16);
                            }
                            if (acta) va = bilerp4(tl, tr, bl, br, lx, ly);
                            if (actb) vb = bilerp4(tl2, tr2, bl2, br2, lx, ly);
                        }
                        if (acta) store_out<kBf16>(p.out, o + ca, va);
                        if (actb) store_out<kBf16>(p.out, o + cb, vb);
                    }
                }
                __syncwarp();
                tick(2);
                if (lane == 0 && yok) {
                    if (yflags & 2) mbar_arrive(empty + sTop);
                    if (yflags & 4) mbar_arrive(empty + sBot);
                }
            }
            headSlot += ny;
            while (headSlot >= K) { headSlot -= K; headPhase ^= 1; }
            tick(3);
        }
        if (p.prof && lane == 0 && cw_all == 0)
            for (int i = 0; i < 4; ++i) atomicAdd(p.prof + 4 + i, pc[i]);
    }
}

// Sampling records of the register-gather kernel in sorted order, one thread per record ENTRY over the whole grid
// (round 1 built them with one CTA per image: 8 of 148 SMs busy for ~30 us at the benchmark shape).
__global__ void __launch_bounds__(256) roi_gather_records_kernel(const RoiRecParams p) {
    const int rec_len = 1 + p.ph + p.pw;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    if (idx >= (long long)p.total * rec_len) return;
    const long long spos = idx / rec_len;
    const int e = (int)(idx - spos * rec_len);
    const int2 cur = __ldg(p.order + spos);
    const float4 box = __ldg(reinterpret_cast<const float4 *>(p.boxes) + cur.x);
    int H, W;
    const float *base;
    switch (cur.y - 2) {
        case 0: H = p.fm_h[0]; W = p.fm_w[0]; base = p.fm[0]; break;
        case 1: H = p.fm_h[1]; W = p.fm_w[1]; base = p.fm[1]; break;
        case 2: H = p.fm_h[2]; W = p.fm_w[2]; base = p.fm[2]; break;
        default: H = p.fm_h[3]; W = p.fm_w[3]; base = p.fm[3]; break;
    }
    int4 v;
    if (e == 0) {
        base += (long long)(cur.x / p.n_boxes) * H * W * p.c4 * 4;
        const unsigned long long bp = (unsigned long long)base;
        v = make_int4((int)(bp & 0xffffffffull), (int)(bp >> 32), cur.x, cur.y);
    } else {
        const bool is_y = e <= p.ph;
        const int j = is_y ? e - 1 : e - 1 - p.ph;
        const int n = is_y ? p.ph : p.pw;
        const float a1 = is_y ? box.x : box.y, a2 = is_y ? box.z : box.w;
        const float Dm1 = (float)((is_y ? H : W) - 1);
        const int stride = is_y ? W * p.c4 : p.c4;
        float in;
        if (n > 1) {
            const float sc = __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), Dm1), (float)(n - 1));
            in = __fadd_rn(__fmul_rn(a1, Dm1), __fmul_rn((float)j, sc));
        } else {
            in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a1, a2)), Dm1);
        }
        const bool ok = (in >= 0.0f) && (in <= Dm1);
        const float fl = floorf(in);
        v = make_int4(ok ? (int)fl * stride : 0, ok ? (int)ceilf(in) * stride : 0, __float_as_int(__fsub_rn(in, fl)), ok ? 1 : 0);
    }
    p.records[idx] = v;
}

// ---------------------------------------------------------------------------------------------
// Backward: gradient w.r.t. the feature maps, the mirror image of the gather.
// TF's CropAndResizeGradImage per in-range sample: dtop = (1-ly)*g, dbottom = ly*g,
// d[top,left] += (1-lx)*dtop, d[top,right] += lx*dtop, likewise for the bottom row; boxes get no gradient
// (tf.stop_gradient, modified_dense_model.py:379-380).  Same records / locality order / warp mapping as the
// forward kernel; the scatter uses 128-bit vector reductions (red.global.add.v4.f32), so the taps that
// overlapping RoIs share are combined in L2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add4(float4 *p, float s, const float4 &g) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(s * g.x), "f"(s * g.y), "f"(s * g.z), "f"(s * g.w)
                 : "memory");
}

__global__ void __launch_bounds__(224) roi_align_backward_kernel(const RoiStreamParams p, const float4 *__restrict__ grad_out) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int rec_len = 1 + p.ph + p.pw;
    const int bins = p.ph * p.pw;
    pdl_wait();
    for (int spos = blockIdx.x; spos < p.total; spos += gridDim.x) {
        const int4 *rec = p.records + (long long)spos * rec_len;
        const int4 hd = __ldg(rec);
        float4 *fm = reinterpret_cast<float4 *>(((unsigned long long)(unsigned)hd.x) | ((unsigned long long)(unsigned)hd.y << 32));
        const long long g_roi = (long long)hd.z * bins * p.c4;
        for (int by = warp; by < p.ph; by += nwarp) {
            const int4 ye = __ldg(rec + 1 + by);
            if (!ye.w) continue;                                       // whole sample row out of range: no gradient
            const float ly = __int_as_float(ye.z);
            float4 *row_t = fm + ye.x, *row_b = fm + ye.y;
            for (int bx = 0; bx < p.pw; ++bx) {
                const int4 xe = __ldg(rec + 1 + p.ph + bx);
                if (!xe.w) continue;
                const float lx = __int_as_float(xe.z);
                const float wt = __fsub_rn(1.f, ly), wl = __fsub_rn(1.f, lx);
                for (int c = lane; c < p.c4; c += 32) {
                    const float4 g = __ldg(grad_out + g_roi + ((long long)by * p.pw + bx) * p.c4 + c);
                    red_add4(row_t + xe.x + c, __fmul_rn(wl, wt), g);
                    red_add4(row_t + xe.y + c, __fmul_rn(lx, wt), g);
                    red_add4(row_b + xe.x + c, __fmul_rn(wl, ly), g);
                    red_add4(row_b + xe.y + c, __fmul_rn(lx, ly), g);
                }
            }
        }
    }
}

static int validate(const float *boxes, const float *const fmaps[4], const int fm_h[4],
                    const int fm_w[4], int n_images, int n_boxes, int channels, int pool_h,
                    int pool_w, int img_h, int img_w, const void *out) {
    DC_REQUIRE(n_images >= 0 && n_boxes >= 0, "negative n_images/n_boxes");
    DC_REQUIRE(n_boxes <= 100000,
               "n_boxes=%d > 100000: the reference's sort key batch*100000+box collides "
               "(modified_dense_model.py:408)", n_boxes);
    DC_REQUIRE(channels > 0 && channels % 4 == 0, "channels=%d must be a positive multiple of 4",
               channels);
    DC_REQUIRE(pool_h >= 1 && pool_h <= kMaxPool && pool_w >= 1 && pool_w <= kMaxPool,
               "pool shape %dx%d outside [1,%d]", pool_h, pool_w, kMaxPool);
    DC_REQUIRE(img_h > 0 && img_w > 0, "image shape must be positive");
    if ((long long)n_images * n_boxes == 0) return DC_OK;
    DC_REQUIRE(boxes && fmaps && fm_h && fm_w && out, "null pointer argument");
    DC_REQUIRE(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "boxes/out must be 16-byte aligned");
    for (int l = 0; l < 4; ++l) {
        DC_REQUIRE(fmaps[l] != nullptr, "feature map %d is null", l);
        DC_REQUIRE(((uintptr_t)fmaps[l] & 15) == 0, "feature map %d not 16-byte aligned", l);
        DC_REQUIRE(fm_h[l] >= 1 && fm_w[l] >= 1, "feature map %d has empty shape", l);
        DC_REQUIRE((long long)fm_h[l] * fm_w[l] < (1ll << 30), "feature map %d too large", l);
        DC_REQUIRE((long long)fm_h[l] * fm_w[l] * (channels / 4) < (1ll << 31),
                   "feature map %d: H*W*C/4 must fit in int32 (record offsets)", l);
    }
    return DC_OK;
}

static float level_denominator(int img_h, int img_w) {
    const float image_area = (float)((double)img_h * (double)img_w);
    return 224.0f / sqrtf(image_area);
}

template <bool kBf16>
static int launch(const float *boxes, const float *const fmaps[4], const int fm_h[4],
                  const int fm_w[4], int n_images, int n_boxes, int channels, int pool_h,
                  int pool_w, int img_h, int img_w, void *out, int32_t *levels,
                  cudaStream_t stream) {
    int rc = validate(boxes, fmaps, fm_h, fm_w, n_images, n_boxes, channels, pool_h, pool_w, img_h,
                      img_w, out);
    if (rc != DC_OK) return rc;
    const long long total = (long long)n_images * n_boxes;
    if (total == 0) return DC_OK;
    DC_REQUIRE(total < (1ll << 31), "n_images*n_boxes must fit in int32");
    const int rec_len = 1 + pool_h + pool_w;

    static const int path = getenv("DCAP_ROI_PATH") ? atoi(getenv("DCAP_ROI_PATH")) : 2;
    // ---- paths 2 / 3 (round 2): locality order (one small CTA per image) -> wide record pre-pass -> gather or ring ----
    {
        const unsigned px_b = (unsigned)channels * 4u, slot_b = 2u * (unsigned)pool_w * px_b;
        const size_t fixed = 2 * ring::kMaxSlots * sizeof(uint64_t) + 128;
        const bool ring_ok = pool_h <= ring::kMaxSamples && pool_w <= ring::kMaxSamples && 4 * (size_t)slot_b + fixed <= 222 * 1024;
        if (path == 2 || (path == 3 && ring_ok)) {
            const bool use_ring = path == 3;
            const size_t ord_bytes = sizeof(int2) * (size_t)total;
            const size_t rec_bytes2 = use_ring ? sizeof(int4) * (size_t)total * ring2::kRec : sizeof(int4) * (size_t)total * rec_len;
            char *ws = nullptr;
            DC_CHECK_CUDA(cudaMallocAsync((void **)&ws, 2 * ord_bytes + rec_bytes2, stream));
            RoiOrderParams op;
            op.boxes = boxes; op.n_boxes = n_boxes; op.denom = level_denominator(img_h, img_w);
            op.order = reinterpret_cast<int2 *>(ws);
            op.scratch = reinterpret_cast<int2 *>(ws + ord_bytes);
            op.levels = levels;
            const size_t ord_smem = (size_t)n_boxes * sizeof(int2);
            cudaError_t e;
            if (ord_smem <= 160 * 1024) {
                static std::atomic<unsigned long long> attr_set{0};
                e = once_per_device(attr_set, [] {
                    return cudaFuncSetAttribute(roi_order_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
                });
                if (e == cudaSuccess) {
                    roi_order_kernel<true><<<n_images, kPrepThreads, ord_smem, stream>>>(op);
                    e = cudaGetLastError();
                }
            } else {
                roi_order_kernel<false><<<n_images, kPrepThreads, 0, stream>>>(op);
                e = cudaGetLastError();
            }
            PdlScope pdl;                 // each kernel's set-up overlaps its predecessor's tail
            RoiRecParams rp;
            rp.boxes = boxes; rp.order = op.order;
            for (int l = 0; l < 4; ++l) { rp.fm[l] = fmaps[l]; rp.fm_h[l] = fm_h[l]; rp.fm_w[l] = fm_w[l]; }
            rp.n_boxes = n_boxes; rp.c4 = channels / 4; rp.ph = pool_h; rp.pw = pool_w; rp.total = (int)total;
            rp.records = reinterpret_cast<int4 *>(ws + 2 * ord_bytes);
            if (e == cudaSuccess && use_ring) {
                e = launch_pdl(roi_ring_records_kernel, dim3((unsigned)ceil_div<long long>(total * 32, 256)), dim3(256), 0, stream, rp);
            } else if (e == cudaSuccess) {
                e = launch_pdl(roi_gather_records_kernel, dim3((unsigned)ceil_div<long long>(total * rec_len, 256)), dim3(256), 0, stream, rp);
            }
            if (e == cudaSuccess && use_ring) {
                static const int env_ctas = getenv("DCAP_ROI_CTAS") ? atoi(getenv("DCAP_ROI_CTAS")) : 2;
                static const int env_slots = getenv("DCAP_ROI_RING") ? atoi(getenv("DCAP_ROI_RING")) : 0;
                static const int env_warps = getenv("DCAP_ROI_WARPS") ? atoi(getenv("DCAP_ROI_WARPS")) : 0;
                static const int env_diag = getenv("DCAP_ROI_DIAG") ? atoi(getenv("DCAP_ROI_DIAG")) : 0;
                static const bool env_prof = getenv("DCAP_ROI_PROF") != nullptr;
                int ctas = env_ctas < 1 ? 1 : env_ctas, slots = 0;
                for (; ctas >= 1; --ctas) {     // K row slots: what fits next to `ctas` resident CTAs per SM
                    const size_t per_cta = (size_t)(224 * 1024) / ctas - 1024 - 512;
                    slots = per_cta > fixed ? (int)((per_cta - fixed) / slot_b) : 0;
                    if (slots >= 4) break;
                }
                if (ctas < 1) ctas = 1;
                if (env_slots >= 4 && env_slots <= slots) slots = env_slots;
                if (slots > ring::kMaxSlots) slots = ring::kMaxSlots;
                static const int env_prod = getenv("DCAP_ROI_PROD") ? atoi(getenv("DCAP_ROI_PROD")) : 1;
                static const int env_groups = getenv("DCAP_ROI_GROUPS") ? atoi(getenv("DCAP_ROI_GROUPS")) : 1;
                static const int env_only_new = getenv("DCAP_ROI_ONLY_NEW") ? atoi(getenv("DCAP_ROI_ONLY_NEW")) : 0;
                const int nprod = env_prod >= 2 ? 2 : 1;
                int groups = env_groups >= 1 && env_groups <= 4 ? env_groups : 1;
                int ncons = env_warps > 0 ? env_warps : (pool_w < 8 ? pool_w : 8) * groups;
                if (ncons > 16 - nprod) ncons = 16 - nprod;
                if (ncons % groups) groups = 1;
                const size_t smem = (size_t)slots * slot_b + fixed;
                static std::atomic<unsigned long long> attr_set2{0};
                e = once_per_device(attr_set2, [] {
                    return cudaFuncSetAttribute(roi_align_ring2_kernel<kBf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
                });
                RoiRing2Params kp;
                kp.records = rp.records; kp.c4 = channels / 4; kp.ph = pool_h; kp.pw = pool_w; kp.total = (int)total;
                kp.out = out; kp.slots = slots; kp.slot_bytes = slot_b; kp.diag = env_diag; kp.prof = nullptr;
                kp.nprod = nprod; kp.groups = groups; kp.only_new = env_only_new;
                const long long mg = (long long)sm_count() * ctas;
                static unsigned long long *prof_buf = nullptr;
                static int prof_calls = 0;
                if (env_prof) {
                    if (!prof_buf) { cudaMalloc((void **)&prof_buf, 64); cudaMemset(prof_buf, 0, 64); }
                    kp.prof = prof_buf;
                    if (++prof_calls % 16 == 0) {
                        unsigned long long h[8];
                        cudaMemcpy(h, prof_buf, 64, cudaMemcpyDeviceToHost);
                        cudaMemset(prof_buf, 0, 64);
                        const double nc = 16.0 * (double)(total < mg ? total : mg);
                        fprintf(stderr, "[roi ring2 prof] per CTA: producer load+table %.0f, issue %.0f | consumer record %.0f, row wait %.0f, "
                                "compute+store %.0f, release %.0f cycles\n", h[0] / nc, h[1] / nc, h[4] / nc, h[5] / nc, h[6] / nc, h[7] / nc);
                    }
                }
                if (e == cudaSuccess)
                    e = launch_pdl(roi_align_ring2_kernel<kBf16>, dim3((unsigned)(total < mg ? total : mg)), dim3((ncons + nprod) * 32), smem, stream, kp);
            } else if (e == cudaSuccess) {
                RoiStreamParams sp;
                sp.records = rp.records; sp.c4 = channels / 4; sp.parts = (sp.c4 + 31) / 32; sp.ph = pool_h; sp.pw = pool_w;
                sp.out = out; sp.total = (int)total;
                const int warps = pool_h < 7 ? pool_h : 7;
                static const int g_ctas = getenv("DCAP_ROI_CTAS") ? atoi(getenv("DCAP_ROI_CTAS")) : 4;
                static const int g_var = getenv("DCAP_ROI_VARIANT") ? atoi(getenv("DCAP_ROI_VARIANT")) : 0;
                const long long mg = (long long)sm_count() * g_ctas;
                const dim3 grid((unsigned)(total < mg ? total : mg)), block(warps * 32);
                switch (g_var) {
                    case 1: e = launch_pdl(roi_align_stream_kernel<2, 72, kBf16, true>, grid, block, 0, stream, sp); break;
                    case 2: e = launch_pdl(roi_align_stream_kernel<4, 128, kBf16, true>, grid, block, 0, stream, sp); break;
                    case 3: e = launch_pdl(roi_align_stream_kernel<4, 128, kBf16, false>, grid, block, 0, stream, sp); break;
                    case 4: e = launch_pdl(roi_align_stream_kernel<7, 168, kBf16, true>, grid, block, 0, stream, sp); break;
                    default: e = launch_pdl(roi_align_stream_kernel<2, 72, kBf16>, grid, block, 0, stream, sp); break;
                }
            }
            if (e == cudaSuccess) e = cudaGetLastError();
            cudaFreeAsync(ws, stream);
            if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "roi align launch failed: %s", cudaGetErrorString(e));
            return DC_OK;
        }
    }
    // ---- ring path, first form (kept for A/B: DCAP_ROI_PATH=1) ----
    const unsigned px_bytes = (unsigned)channels * 4u;
    const unsigned slot_bytes = 2u * (unsigned)pool_w * px_bytes;
    const size_t ring_fixed = ring::kDescDepth * sizeof(ring::Desc) + (2 * ring::kMaxSlots + 2 * ring::kDescDepth) * sizeof(uint64_t) + 128;
    if (path == 1 && pool_h <= ring::kMaxSamples && pool_w <= ring::kMaxSamples && 4 * (size_t)slot_bytes + ring_fixed <= 223 * 1024) {
        static const int env_ctas = getenv("DCAP_ROI_CTAS") ? atoi(getenv("DCAP_ROI_CTAS")) : 2;
        static const int env_slots = getenv("DCAP_ROI_RING") ? atoi(getenv("DCAP_ROI_RING")) : 0;
        static const int env_warps = getenv("DCAP_ROI_WARPS") ? atoi(getenv("DCAP_ROI_WARPS")) : 0;
        int ctas = env_ctas < 1 ? 1 : env_ctas;
        // K row slots: what fits next to `ctas` resident CTAs per SM (227 KB usable, 1 KB reserved per CTA)
        int slots = 0;
        for (; ctas >= 1; --ctas) {
            const size_t per_cta = (size_t)(224 * 1024) / ctas - 1024;   // static smem + the 1 KB the driver reserves per CTA
            slots = per_cta > ring_fixed ? (int)((per_cta - ring_fixed) / slot_bytes) : 0;
            if (slots >= 4) break;
        }
        if (ctas < 1) ctas = 1;
        if (env_slots >= 4 && env_slots <= slots) slots = env_slots;
        if (slots > ring::kMaxSlots) slots = ring::kMaxSlots;
        int ncons = env_warps > 0 ? env_warps : (pool_w < 8 ? pool_w : 8);
        if (ncons > 15) ncons = 15;
        const size_t smem = (size_t)slots * slot_bytes + ring_fixed;

        const size_t ord_bytes = sizeof(int2) * (size_t)total;
        char *ws = nullptr;
        DC_CHECK_CUDA(cudaMallocAsync((void **)&ws, 2 * ord_bytes, stream));
        RoiOrderParams op;
        op.boxes = boxes; op.n_boxes = n_boxes; op.denom = level_denominator(img_h, img_w);
        op.order = reinterpret_cast<int2 *>(ws);
        op.scratch = reinterpret_cast<int2 *>(ws + ord_bytes);
        op.levels = levels;
        const size_t ord_smem = (size_t)n_boxes * sizeof(int2);
        cudaError_t e;
        if (ord_smem <= 160 * 1024) {
            static std::atomic<unsigned long long> attr_set{0};
            e = once_per_device(attr_set, [] {
                return cudaFuncSetAttribute(roi_order_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            });
            if (e == cudaSuccess) {
                roi_order_kernel<true><<<n_images, kPrepThreads, ord_smem, stream>>>(op);
                e = cudaGetLastError();
            }
        } else {
            roi_order_kernel<false><<<n_images, kPrepThreads, 0, stream>>>(op);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
            static std::atomic<unsigned long long> attr_set2{0};
            e = once_per_device(attr_set2, [] {
                return cudaFuncSetAttribute(roi_align_ring_kernel<kBf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);   // + 640 B static
            });
        }
        if (e == cudaSuccess) {
            PdlScope pdl;                 // the ring kernel's barrier set-up overlaps the order kernel
            const long long mg = (long long)sm_count() * ctas;
            RoiRingParams rp;
            rp.boxes = boxes; rp.order = op.order;
            for (int l = 0; l < 4; ++l) { rp.fm[l] = fmaps[l]; rp.fm_h[l] = fm_h[l]; rp.fm_w[l] = fm_w[l]; }
            rp.n_boxes = n_boxes; rp.c4 = channels / 4; rp.ph = pool_h; rp.pw = pool_w;
            rp.out = out; rp.total = (int)total; rp.slots = slots; rp.slot_bytes = slot_bytes;
            static const int env_diag = getenv("DCAP_ROI_DIAG") ? atoi(getenv("DCAP_ROI_DIAG")) : 0;
            static const int env_sync = getenv("DCAP_ROI_SYNC") ? atoi(getenv("DCAP_ROI_SYNC")) : 0;
            static const bool env_prof = getenv("DCAP_ROI_PROF") != nullptr;
            rp.diag = env_diag; rp.sync_mode = env_sync; rp.prof = nullptr;
            static unsigned long long *prof_buf = nullptr;
            static int prof_calls = 0;
            if (env_prof) {
                if (!prof_buf) { cudaMalloc((void **)&prof_buf, 64); cudaMemset(prof_buf, 0, 64); }
                rp.prof = prof_buf;
                if (++prof_calls % 16 == 0) {              // mean cycles per CTA role and call, over the last 16 calls
                    unsigned long long h[8];
                    cudaMemcpy(h, prof_buf, 64, cudaMemcpyDeviceToHost);
                    cudaMemset(prof_buf, 0, 64);
                    const double nc = 16.0 * (double)(total < mg ? total : mg);
                    fprintf(stderr, "[roi ring prof] per CTA: producer compute %.0f, desc wait %.0f, publish %.0f, issue %.0f | consumer "
                            "desc wait %.0f, row wait %.0f, compute+store %.0f, release %.0f cycles\n", h[0] / nc, h[1] / nc, h[2] / nc,
                            h[3] / nc, h[4] / nc, h[5] / nc, h[6] / nc, h[7] / nc);
                }
            }
            e = launch_pdl(roi_align_ring_kernel<kBf16>, dim3((unsigned)(total < mg ? total : mg)), dim3((ncons + 1) * 32), smem,
                           stream, rp);
            if (e == cudaSuccess) e = cudaGetLastError();
        }
        cudaFreeAsync(ws, stream);
        if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "roi align launch failed: %s", cudaGetErrorString(e));
        return DC_OK;
    }

    // ---- register-gather path (round 1; also the fallback for pools > 16 or very wide pixels) ----
    // stream-ordered workspace: sorted records + scratch (cached by the default pool)
    const size_t rec_bytes = sizeof(int4) * (size_t)total * rec_len;
    const size_t scr_bytes = sizeof(int4) * (size_t)total;
    char *ws = nullptr;
    DC_CHECK_CUDA(cudaMallocAsync((void **)&ws, rec_bytes + scr_bytes, stream));

    RoiPrepParams pp;
    pp.boxes = boxes;
    for (int l = 0; l < 4; ++l) { pp.fm[l] = fmaps[l]; pp.fm_h[l] = fm_h[l]; pp.fm_w[l] = fm_w[l]; }
    pp.n_boxes = n_boxes;
    pp.c4 = channels / 4;
    pp.ph = pool_h; pp.pw = pool_w;
    pp.denom = level_denominator(img_h, img_w);
    pp.records = reinterpret_cast<int4 *>(ws);
    pp.scratch = reinterpret_cast<int4 *>(ws + rec_bytes);
    pp.levels = levels;
    const size_t prep_smem = (size_t)n_boxes * 32;
    if (prep_smem <= 160 * 1024) {
        static std::atomic<unsigned long long> attr_set{0};
        DC_CHECK_CUDA(once_per_device(attr_set, [] {
            return cudaFuncSetAttribute(roi_prepare_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        }));
        roi_prepare_kernel<true><<<n_images, kPrepThreads, prep_smem, stream>>>(pp);
    } else {
        roi_prepare_kernel<false><<<n_images, kPrepThreads, 0, stream>>>(pp);
    }
    cudaError_t e = cudaGetLastError();

    if (e == cudaSuccess) {
        PdlScope pdl;                 // the stream kernel's launch overlaps the tail of the prepare kernel
        RoiStreamParams sp;
        sp.records = pp.records;
        sp.c4 = channels / 4;
        sp.parts = (sp.c4 + 31) / 32;
        sp.ph = pool_h; sp.pw = pool_w;
        sp.out = out;
        sp.total = (int)total;
        const int warps = pool_h < 7 ? pool_h : 7;          // one warp per sample row
        // a whole number of waves: SM count x resident CTAs per SM
        const long long max_grid = (long long)sm_count() * 8;
        const int grid = (int)(total < max_grid ? total : max_grid);
        static const int variant = getenv("DCAP_ROI_VARIANT") ? atoi(getenv("DCAP_ROI_VARIANT")) : 5;
        static const int ctas = getenv("DCAP_ROI_CTAS") ? atoi(getenv("DCAP_ROI_CTAS")) : 4;
        const long long mg = (long long)sm_count() * ctas;
        const int g2 = (int)(total < mg ? total : mg);
        switch (variant) {
            case 1: e = launch_pdl(roi_align_stream_kernel<4, 128, kBf16>, dim3(g2), dim3(warps * 32), 0, stream, sp); break;
            case 2: e = launch_pdl(roi_align_stream_kernel<2, 64, kBf16>, dim3(g2), dim3(warps * 32), 0, stream, sp); break;
            case 3: e = launch_pdl(roi_align_stream_kernel<2, 80, kBf16>, dim3(g2), dim3(warps * 32), 0, stream, sp); break;
            case 4: e = launch_pdl(roi_align_stream_kernel<1, 40, kBf16>, dim3(g2), dim3(warps * 32), 0, stream, sp); break;
            case 5: e = launch_pdl(roi_align_stream_kernel<2, 72, kBf16>, dim3(g2), dim3(warps * 32), 0, stream, sp); break;
            case 6: e = launch_pdl(roi_align_stream_kernel<1, 48, kBf16>, dim3(g2), dim3(warps * 32), 0, stream, sp); break;
            default: e = launch_pdl(roi_align_stream_kernel<4, 96, kBf16>, dim3(g2), dim3(warps * 32), 0, stream, sp); break;
        }
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    cudaFreeAsync(ws, stream);
    if (e != cudaSuccess)
        return set_error(DC_ERR_CUDA, "roi align launch failed: %s", cudaGetErrorString(e));
    return DC_OK;
}

}  // namespace dcap

using namespace dcap;

extern "C" int dc_fpn_levels_f32(const float *boxes, int64_t n_boxes, int img_h, int img_w,
                                 int32_t *levels, void *stream) {
    DC_REQUIRE(n_boxes >= 0, "negative n_boxes");
    DC_REQUIRE(img_h > 0 && img_w > 0, "image shape must be positive");
    if (n_boxes == 0) return DC_OK;
    DC_REQUIRE(boxes && levels, "null pointer argument");
    DC_REQUIRE(((uintptr_t)boxes & 15) == 0, "boxes must be 16-byte aligned");
    const int threads = 128;
    const long long blocks = ceil_div<long long>(n_boxes, threads);
    fpn_levels_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        boxes, n_boxes, level_denominator(img_h, img_w), levels);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

extern "C" int dc_pyramid_roi_align_f32(const float *boxes, const float *const fmaps[4],
                                        const int fm_h[4], const int fm_w[4], int n_images,
                                        int n_boxes, int channels, int pool_h, int pool_w,
                                        int img_h, int img_w, float *out, int32_t *levels,
                                        void *stream) {
    return launch<false>(boxes, fmaps, fm_h, fm_w, n_images, n_boxes, channels, pool_h, pool_w,
                         img_h, img_w, out, levels, (cudaStream_t)stream);
}

extern "C" int dc_pyramid_roi_align_bf16out(const float *boxes, const float *const fmaps[4],
                                            const int fm_h[4], const int fm_w[4], int n_images,
                                            int n_boxes, int channels, int pool_h, int pool_w,
                                            int img_h, int img_w, uint16_t *out, int32_t *levels,
                                            void *stream) {
    DC_REQUIRE(channels % 8 == 0 || (long long)n_images * n_boxes == 0,
               "bf16 output needs channels %% 8 == 0");
    return launch<true>(boxes, fmaps, fm_h, fm_w, n_images, n_boxes, channels, pool_h, pool_w,
                        img_h, img_w, out, levels, (cudaStream_t)stream);
}

extern "C" int dc_pyramid_roi_align_backward_f32(const float *boxes, const float *grad_out, float *const d_fmaps[4],
                                                 const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                                 int channels, int pool_h, int pool_w, int img_h, int img_w, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const float *maps[4] = {d_fmaps ? d_fmaps[0] : nullptr, d_fmaps ? d_fmaps[1] : nullptr, d_fmaps ? d_fmaps[2] : nullptr,
                            d_fmaps ? d_fmaps[3] : nullptr};
    int rc = validate(boxes, d_fmaps ? maps : nullptr, fm_h, fm_w, n_images, n_boxes, channels, pool_h, pool_w, img_h, img_w, grad_out);
    if (rc != DC_OK) return rc;
    const long long total = (long long)n_images * n_boxes;
    if (total == 0) return DC_OK;
    DC_REQUIRE(total < (1ll << 31), "n_images*n_boxes must fit in int32");
    const int rec_len = 1 + pool_h + pool_w;
    const size_t rec_bytes = sizeof(int4) * (size_t)total * rec_len;
    char *ws = nullptr;
    DC_CHECK_CUDA(cudaMallocAsync((void **)&ws, rec_bytes + sizeof(int4) * (size_t)total, stream));
    RoiPrepParams pp;
    pp.boxes = boxes;
    for (int l = 0; l < 4; ++l) { pp.fm[l] = maps[l]; pp.fm_h[l] = fm_h[l]; pp.fm_w[l] = fm_w[l]; }
    pp.n_boxes = n_boxes; pp.c4 = channels / 4; pp.ph = pool_h; pp.pw = pool_w;
    pp.denom = level_denominator(img_h, img_w);
    pp.records = reinterpret_cast<int4 *>(ws);
    pp.scratch = reinterpret_cast<int4 *>(ws + rec_bytes);
    pp.levels = nullptr;
    const size_t prep_smem = (size_t)n_boxes * 32;
    if (prep_smem <= 160 * 1024) {
        static std::atomic<unsigned long long> attr_set{0};
        DC_CHECK_CUDA(once_per_device(attr_set, [] {
            return cudaFuncSetAttribute(roi_prepare_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        }));
        roi_prepare_kernel<true><<<n_images, kPrepThreads, prep_smem, stream>>>(pp);
    } else {
        roi_prepare_kernel<false><<<n_images, kPrepThreads, 0, stream>>>(pp);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        RoiStreamParams sp;
        sp.records = pp.records; sp.c4 = channels / 4; sp.parts = (sp.c4 + 31) / 32; sp.ph = pool_h; sp.pw = pool_w;
        sp.out = nullptr; sp.total = (int)total;
        const int warps = pool_h < 7 ? pool_h : 7;
        const long long mg = (long long)sm_count() * 8;
        roi_align_backward_kernel<<<(int)(total < mg ? total : mg), warps * 32, 0, stream>>>(sp, reinterpret_cast<const float4 *>(grad_out));
        e = cudaGetLastError();
    }
    cudaFreeAsync(ws, stream);
    if (e != cudaSuccess) return set_error(DC_ERR_CUDA, "roi align backward launch failed: %s", cudaGetErrorString(e));
    return DC_OK;
}

// Host-buffer form: image-by-image pipeline.  Stream A uploads image i's four maps, stream B runs
// the kernel on image i's boxes and downloads its output slice, so the H2D of image i+1 overlaps
// the kernel + D2H of image i (PCIe is full duplex).  Device staging comes from the stream-ordered
// pool (cudaMallocAsync), double-buffered per image.
extern "C" int dc_pyramid_roi_align_host_f32(const float *boxes, const float *const fmaps[4],
                                             const int fm_h[4], const int fm_w[4], int n_images,
                                             int n_boxes, int channels, int pool_h, int pool_w,
                                             int img_h, int img_w, float *out, int32_t *levels) {
    // validate() checks device-style alignment; host buffers only need float alignment, so the
    // checks that matter here are the shape ones.
    DC_REQUIRE(n_images >= 0 && n_boxes >= 0 && n_boxes <= 100000, "bad n_images/n_boxes");
    DC_REQUIRE(channels > 0 && channels % 4 == 0, "channels=%d must be a positive multiple of 4",
               channels);
    DC_REQUIRE(pool_h >= 1 && pool_h <= kMaxPool && pool_w >= 1 && pool_w <= kMaxPool,
               "pool shape %dx%d outside [1,%d]", pool_h, pool_w, kMaxPool);
    DC_REQUIRE(img_h > 0 && img_w > 0, "image shape must be positive");
    if ((long long)n_images * n_boxes == 0) return DC_OK;
    DC_REQUIRE(boxes && fmaps && fm_h && fm_w && out, "null pointer argument");

    cudaStream_t sa = nullptr, sb = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    float *d_fm[2][4] = {{nullptr}};
    float *d_out[2] = {nullptr, nullptr};
    float *d_boxes = nullptr;
    int32_t *d_levels = nullptr;
    size_t fm_bytes[4];
    for (int l = 0; l < 4; ++l)
        fm_bytes[l] = sizeof(float) * (size_t)fm_h[l] * fm_w[l] * channels;
    const size_t out_bytes = sizeof(float) * (size_t)n_boxes * pool_h * pool_w * channels;
    int rc = DC_OK;
#define HOST_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess && rc == DC_OK)                                               \
            rc = set_error(DC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)
    HOST_TRY(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
    HOST_TRY(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
    for (int i = 0; i < 2 && rc == DC_OK; ++i) {
        HOST_TRY(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming));
        HOST_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        for (int l = 0; l < 4; ++l) HOST_TRY(cudaMallocAsync(&d_fm[i][l], fm_bytes[l], sa));
        HOST_TRY(cudaMallocAsync(&d_out[i], out_bytes, sa));
    }
    if (rc == DC_OK) {
        HOST_TRY(cudaMallocAsync(&d_boxes, sizeof(float) * 4 * (size_t)n_images * n_boxes, sa));
        if (levels)
            HOST_TRY(cudaMallocAsync(&d_levels, sizeof(int32_t) * (size_t)n_images * n_boxes, sa));
        HOST_TRY(cudaMemcpyAsync(d_boxes, boxes, sizeof(float) * 4 * (size_t)n_images * n_boxes,
                                 cudaMemcpyHostToDevice, sa));
    }
    for (int img = 0; img < n_images && rc == DC_OK; ++img) {
        const int s = img & 1;
        if (img >= 2) HOST_TRY(cudaStreamWaitEvent(sa, done[s], 0));   // slot free again
        for (int l = 0; l < 4; ++l)
            HOST_TRY(cudaMemcpyAsync(d_fm[s][l], fmaps[l] + (size_t)img * (fm_bytes[l] / 4),
                                     fm_bytes[l], cudaMemcpyHostToDevice, sa));
        HOST_TRY(cudaEventRecord(up[s], sa));
        HOST_TRY(cudaStreamWaitEvent(sb, up[s], 0));
        if (rc != DC_OK) break;
        const float *maps[4] = {d_fm[s][0], d_fm[s][1], d_fm[s][2], d_fm[s][3]};
        int krc = dc_pyramid_roi_align_f32(d_boxes + (size_t)img * n_boxes * 4, maps, fm_h, fm_w, 1,
                                           n_boxes, channels, pool_h, pool_w, img_h, img_w,
                                           d_out[s], d_levels ? d_levels + (size_t)img * n_boxes
                                                              : nullptr, sb);
        if (krc != DC_OK) { rc = krc; break; }
        HOST_TRY(cudaMemcpyAsync(out + (size_t)img * (out_bytes / 4), d_out[s], out_bytes,
                                 cudaMemcpyDeviceToHost, sb));
        HOST_TRY(cudaEventRecord(done[s], sb));
    }
    if (rc == DC_OK && levels) {
        HOST_TRY(cudaStreamWaitEvent(sb, up[(n_images - 1) & 1], 0));
        HOST_TRY(cudaMemcpyAsync(levels, d_levels, sizeof(int32_t) * (size_t)n_images * n_boxes,
                                 cudaMemcpyDeviceToHost, sb));
    }
    if (sa) cudaStreamSynchronize(sa);
    if (sb) {
        cudaError_t e = cudaStreamSynchronize(sb);
        if (e != cudaSuccess && rc == DC_OK)
            rc = set_error(DC_ERR_CUDA, "roi align host pipeline failed: %s", cudaGetErrorString(e));
    }
    for (int i = 0; i < 2; ++i) {
        for (int l = 0; l < 4; ++l) if (d_fm[i][l]) cudaFreeAsync(d_fm[i][l], sb ? sb : 0);
        if (d_out[i]) cudaFreeAsync(d_out[i], sb ? sb : 0);
        if (up[i]) cudaEventDestroy(up[i]);
        if (done[i]) cudaEventDestroy(done[i]);
    }
    if (d_boxes) cudaFreeAsync(d_boxes, sb ? sb : 0);
    if (d_levels) cudaFreeAsync(d_levels, sb ? sb : 0);
    if (sb) cudaStreamSynchronize(sb);
    if (sa) cudaStreamDestroy(sa);
    if (sb) cudaStreamDestroy(sb);
#undef HOST_TRY
    return rc;
}

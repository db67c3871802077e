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
// Pass 1: locality order + per-RoI sampling records, stored IN SORTED ORDER (roi_order_kernel +
// roi_gather_records_kernel, further down).
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
constexpr int kBuckets = 1024;      // 4 levels x 256 Morton tiles
constexpr int kPrepThreads = 1024;

__device__ __forceinline__ unsigned morton4(unsigned v) {   // spread 4 bits: abcd -> 0a0b0c0d
    v &= 0xF;
    v = (v | (v << 2)) & 0x33;
    v = (v | (v << 1)) & 0x55;
    return v;
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

template <int kUnroll, int kMaxReg, bool kBf16>
__global__ void __maxnreg__(kMaxReg)
roi_align_stream_kernel(const RoiStreamParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int rec_len = 1 + p.ph + p.pw;
    const int bins = p.ph * p.pw;
    pdl_wait();                       // launched with programmatic dependent launch right behind roi_gather_records_kernel

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
#pragma unroll
                    for (int k = 0; k < kUnroll; ++k) {
                        ok[k] = false;
                        if (bx0 + k < p.pw) {
                            const int4 xe = __ldg(rec + 1 + p.ph + bx0 + k);
                            ok[k] = act && ye.w && xe.w;
                            lx[k] = __int_as_float(xe.z);
                            if (ok[k]) {
                                tl[k] = __ldg(row_t + xe.x + c);
                                tr[k] = __ldg(row_t + xe.y + c);
                                bl[k] = __ldg(row_b + xe.x + c);
                                br[k] = __ldg(row_b + xe.y + c);
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
// Round-2 prologue.  The round-1 pass 1 above (one CTA per image doing levels, sort AND 240 B of sampling records per
// RoI) kept 8 of 148 SMs busy for ~18 us and cost ~30 us of the 0.18 ms launch pair.  It is split in two:
//   * roi_order_kernel (one CTA per image): levels + locality order only -- 8 bytes per RoI;
//   * roi_gather_records_kernel: the sampling records in sorted order, ONE THREAD PER RECORD ENTRY over the whole grid.
// Both are chained to the streaming kernel with programmatic dependent launch.
//
// Tried in round 2 and removed (measured on B200, cfg2 = 8 x 1000 RoIs, fp32 output, 863 MB algorithmic; the
// register-gather kernel with this prologue: 0.169-0.171 ms = 0.77-0.78 of the measured copy peak):
//   * a producer/consumer kernel around a shared-memory ring of map ROWS fed by cp.async.bulk (UBLKCP), one producer warp
//     de-duplicating rows / pixels and merging pixel runs, 7 consumer warps lerping out of shared memory: bit-exact,
//     0.243 ms.  clock64 marks showed the single producer warp at ~7700 dependent cycles per RoI;
//   * the same with everything box-only (samples, de-duplication, runs, release flags) precomputed by a one-warp-per-RoI
//     pre-pass into 1 KB "ring records", no descriptor ring, division-free slot arithmetic: 0.208 ms; with all memory
//     traffic switched off the mbarrier skeleton alone still took 0.100 ms (per row slot: empty-wait + expect_tx + issue on
//     the producer, two full-waits + release on seven consumer warps), a second producer warp or alternating consumer
//     groups made it slower (0.28-0.32 ms).  A RoI's taps span ~144 KB of map rows, so shared memory holds less than two
//     RoIs: the ring has to be row-granular, and at ~12 rows x 8000 RoIs the per-row handshakes cost more than the
//     ~25 % of L2 -> SM tap traffic the de-duplication saves;
//   * re-using tap registers between neighbouring bins inside the gather kernel (the branch-free loads become
//     conditional): 0.287 ms;
//   * prefetch.global.L2 of the taps of the CTA's next RoI from inside the gather kernel: 0.221-0.228 ms (more requests,
//     not fewer stalls: the memory system's request throughput on 1 KB granules is the bound, not latency);
//   * the 4 (or 2, or 8) CTAs resident on one SM working on ADJACENT positions of the locality order at the same time, so
//     that overlapping boxes meet in that SM's L1: 0.177 / 0.170 ms (4 / 2 adjacent, 4 CTAs per SM) and 0.193 ms (8 per
//     SM) against 0.169 ms -- neighbours in the order then no longer spread over many SMs' request queues.
// What bounds the gather kernel is the L2 -> SM path: ~1.6 GB of taps per launch in ~0.14 ms = 11-12 TB/s, with DRAM at
// the compulsory bytes (profiles/): close to what the part's L2 delivers to tcgen05 operand loads as well (section 9).
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

struct RoiRecParams {
    const float *boxes;
    const int2 *order;
    const float *fm[4];
    int fm_h[4];
    int fm_w[4];
    int n_boxes, c4, ph, pw, total;
    int4 *records;
};

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

// Prologue shared by the forward and the backward pass: levels + locality order, then the sorted sampling records.
// Leaves the records in a stream-ordered workspace (*ws, to be released with cudaFreeAsync by the caller).
static int sorted_records(const float *boxes, const float *const maps[4], const int fm_h[4], const int fm_w[4], int n_images,
                          int n_boxes, int channels, int pool_h, int pool_w, int img_h, int img_w, int32_t *levels,
                          cudaStream_t stream, char **ws, int4 **records) {
    const long long total = (long long)n_images * n_boxes;
    const int rec_len = 1 + pool_h + pool_w;
    const size_t ord_bytes = sizeof(int2) * (size_t)total;
    DC_CHECK_CUDA(cudaMallocAsync((void **)ws, 2 * ord_bytes + sizeof(int4) * (size_t)total * rec_len, stream));
    RoiOrderParams op;
    op.boxes = boxes; op.n_boxes = n_boxes; op.denom = level_denominator(img_h, img_w);
    op.order = reinterpret_cast<int2 *>(*ws);
    op.scratch = reinterpret_cast<int2 *>(*ws + ord_bytes);
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
        PdlScope pdl;                 // the record kernel's launch overlaps the order kernel's tail
        RoiRecParams rp;
        rp.boxes = boxes; rp.order = op.order;
        for (int l = 0; l < 4; ++l) { rp.fm[l] = maps[l]; rp.fm_h[l] = fm_h[l]; rp.fm_w[l] = fm_w[l]; }
        rp.n_boxes = n_boxes; rp.c4 = channels / 4; rp.ph = pool_h; rp.pw = pool_w; rp.total = (int)total;
        rp.records = reinterpret_cast<int4 *>(*ws + 2 * ord_bytes);
        *records = rp.records;
        e = launch_pdl(roi_gather_records_kernel, dim3((unsigned)ceil_div<long long>(total * rec_len, 256)), dim3(256), 0, stream, rp);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        cudaFreeAsync(*ws, stream);
        *ws = nullptr;
        return set_error(DC_ERR_CUDA, "roi align prologue failed: %s", cudaGetErrorString(e));
    }
    return DC_OK;
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
    char *ws = nullptr;
    int4 *records = nullptr;
    if (int prc = sorted_records(boxes, fmaps, fm_h, fm_w, n_images, n_boxes, channels, pool_h, pool_w, img_h, img_w, levels,
                                 stream, &ws, &records)) return prc;
    cudaError_t e;
    {
        PdlScope pdl;                 // the stream kernel's launch overlaps the tail of the record kernel
        RoiStreamParams sp;
        sp.records = records;
        sp.c4 = channels / 4;
        sp.parts = (sp.c4 + 31) / 32;
        sp.ph = pool_h; sp.pw = pool_w;
        sp.out = out;
        sp.total = (int)total;
        const int warps = pool_h < 7 ? pool_h : 7;          // one warp per sample row
        // a whole number of waves: SM count x 4 resident CTAs per SM (measured on cfg2: 3 -> 0.193, 4 -> 0.171, 5 -> 0.219,
        // 6 -> 0.199, 8 -> 0.177 ms; 4 loads x 2 bins in flight per lane at 72 registers beat 4 bins at 128 and 1 bin at 48)
        const long long mg = (long long)sm_count() * 4;
        e = launch_pdl(roi_align_stream_kernel<2, 72, kBf16>, dim3((unsigned)(total < mg ? total : mg)), dim3(warps * 32), 0, stream, sp);
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
    char *ws = nullptr;
    int4 *records = nullptr;
    if (int prc = sorted_records(boxes, maps, fm_h, fm_w, n_images, n_boxes, channels, pool_h, pool_w, img_h, img_w, nullptr, stream,
                                 &ws, &records)) return prc;
    RoiStreamParams sp;
    sp.records = records; sp.c4 = channels / 4; sp.parts = (sp.c4 + 31) / 32; sp.ph = pool_h; sp.pw = pool_w;
    sp.out = nullptr; sp.total = (int)total;
    const int warps = pool_h < 7 ? pool_h : 7;
    const long long mg = (long long)sm_count() * 8;
    cudaError_t e;
    {
        PdlScope pdl;
        e = launch_pdl(roi_align_backward_kernel, dim3((unsigned)(total < mg ? total : mg)), dim3(warps * 32), 0, stream, sp,
                       reinterpret_cast<const float4 *>(grad_out));
        if (e == cudaSuccess) e = cudaGetLastError();
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

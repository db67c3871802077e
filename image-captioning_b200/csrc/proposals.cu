// Box front-end of the RoI path: the RPN proposal filter that produces the
// boxes PyramidROIAlign consumes when use_generated_rois is set, and the GT-box normalisation used otherwise.
// Replaces ProposalLayer.call (/root/reference/dense_img_cap_separate_models/modified_dense_model.py:247-303:
// tf.nn.top_k(6000) -> gather -> apply_box_deltas_graph (:179-200) -> clip_boxes_graph (:203-218) ->
// / [h,w,h,w] -> tf.image.non_max_suppression -> zero pad) and the Lambda at :1523-1526.
//
// Three kernels (select once; masks + scan per band of candidates), all latency- or issue-bound integer and fp32
// work (no tensor cores):
//
//   1. proposal_select_kernel: one CLUSTER of 8 CTAs per image.  Every CTA reads its eighth of the image's
//      foreground scores ONCE into shared memory as order-preserving 32-bit keys; an 8-bit x 4-pass radix
//      select finds the pre_nms_limit-th largest key with the per-pass histograms summed over the cluster in
//      CTA 0's shared memory (DSMEM atomics).  Each CTA then compacts ITS candidates (ties at the threshold:
//      lowest anchor index first, as top_k), deals them evenly over the cluster (DSMEM stores) and every CTA
//      bitonic-sorts its share as 64-bit (key, ~anchor) words; the
//      eight sorted lists are merged by ranking -- a candidate's final position is its own index plus the number
//      of larger words in the other seven lists, found by interleaved binary searches over DSMEM -- and each
//      CTA refines / clips / normalises its own candidates' boxes and writes them to their score-ordered slots.
//   2. proposal_iou_mask_kernel: 64x256 tiles of the upper triangle of the IoU > threshold relation as bit masks.
//   3. proposal_nms_scan_kernel: one CTA per image walks the boxes in score order 64 at a time (the 64-step
//      dependency chain runs in registers of one warp, the survivors' mask rows are OR-ed by the whole CTA),
//      stops at proposal_count, gathers the survivors and zero-pads.
//   2 and 3 alternate over bands of 16 x 64 candidates; an image that has all its proposals skips the later bands.
//
// fp32 arithmetic op for op as the TF graph (explicit _rn intrinsics: no FMA contraction); tf.exp is the
// correctly rounded fp32 exponential (fp64 exp rounded once), as in oracle/proposals.py.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dcap {

constexpr int kSelThreads = 1024;
constexpr int kSelCluster = 8;
constexpr int kMaxPreNms = 8192;
constexpr int kSelMaxChunk = 48 * 1024;          // keys cached in shared memory per CTA (192 KB) -> 393216 anchors per image
constexpr int kPendCap = 16384;                  // keys kept for the 2nd-4th radix passes (64 KB behind the key cache)
constexpr int kScanThreads = 1024;               // the OR phase wants many independent fetches in flight
constexpr int kBandBlocks = 16;                  // NMS runs in bands of 16 x 64 candidates (mask launch + scan launch per band),
constexpr int kMaxBands = 4;                     // the last band taking whatever is left

// Descending-score order as ascending-key order reversed: larger score <-> larger key.  -0 == +0; NaN sorts last.
__device__ __forceinline__ uint32_t score_key(float s) {
    if (s != s) return 0u;
    const uint32_t u = __float_as_uint(s + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// std::min / std::max as TF's C++ kernels (and Eigen's scalar min/max behind tf.minimum / tf.maximum) evaluate them:
// the FIRST argument is returned when the comparison is false, so a NaN first argument propagates and equal
// operands (+-0) return the first.  fminf/fmaxf would drop the NaN.
__device__ __forceinline__ float std_min(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return a < b ? b : a; }

static_assert(kSelThreads == 4 * 256, "the digit histograms are zeroed one word per thread");
constexpr int kHistCopies = 8;                  // replicated per-CTA histograms: RPN scores share few top bytes

struct SelSmem {
    unsigned long long sort[kMaxPreNms / kSelCluster];   // this CTA's share of the candidates as (key << 32) | ~anchor, sorted descending
    unsigned int hist[4][256];                // cluster-wide histogram of each digit (CTA 0's copy), all zeroed up front
    unsigned int lhist[kHistCopies][256];     // this CTA's histogram, one copy per warp & 7
    unsigned int n_gt[kSelCluster], n_eq[kSelCluster];   // per-CTA counts, replicated in every CTA
    unsigned int lsum[256];                   // this CTA's histogram summed over the copies (kept until the digit is chosen)
    unsigned int warp_cnt[kSelThreads / 32];
    unsigned int sel_bin, sel_above, ctr, loc_gt, loc_eq, pend_ctr;
};

template <bool kCache>
__global__ void __cluster_dims__(kSelCluster, 1, 1) __launch_bounds__(kSelThreads)
proposal_select_kernel(const float *__restrict__ rpn_probs, const float4 *__restrict__ rpn_bbox,
                       const float4 *__restrict__ anchors, int n_anchors, int k_eff, float4 std_dev, float img_h,
                       float img_w, float4 *__restrict__ ws_boxes, int32_t *__restrict__ ws_index, int pend_cap) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    SelSmem &sm = *reinterpret_cast<SelSmem *>(sm_raw);
    uint32_t *s_key = reinterpret_cast<uint32_t *>(sm_raw + sizeof(SelSmem));
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int img = blockIdx.x / kSelCluster;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SelSmem *sm0 = cluster.map_shared_rank(&sm, 0);            // holds the cluster-wide digit histograms

    const int chunk = ceil_div(n_anchors, kSelCluster);
    const int lo = min(rank * chunk, n_anchors), n_local = min(lo + chunk, n_anchors) - lo;
    // keys still in the running after the first digit (those sharing the chosen top byte), compacted so that the
    // remaining three passes do not rescan the whole slice; behind the key cache, pend_cap entries (0 = none)
    uint32_t *s_pend = s_key + (kCache ? ((chunk + 3) & ~3) : 0);
    bool use_pend = false;
    int n_scan = n_local;
    const float *score = rpn_probs + ((long long)img * n_anchors + lo) * 2 + 1;      // foreground column of [A, 2]
    auto key_at = [&](int i) -> uint32_t { return kCache ? s_key[i] : score_key(__ldg(score + 2 * (long long)i)); };
    if (kCache) {
        // [A,2] rows are 8 bytes: read whole float2 rows (coalesced) and keep the foreground half
        const float2 *rows = reinterpret_cast<const float2 *>(rpn_probs) + (long long)img * n_anchors + lo;
        for (int i = tid; i < n_local; i += kSelThreads) s_key[i] = score_key(__ldg(rows + i).y);
    }

    // ---- radix select: the k_eff-th largest key ------------------------------------------------------------
    uint32_t prefix = 0, mask = 0;
    unsigned int remaining = (unsigned int)k_eff;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        unsigned int *hist0 = sm0->hist[pass];
        for (int i = tid; i < kHistCopies * 256; i += kSelThreads) (&sm.lhist[0][0])[i] = 0;
        if (pass == 0) {
            if (rank == 0) (&sm.hist[0][0])[tid] = 0;       // 4 x 256 words, one per thread
            cluster.sync();                                 // zeroed before any remote add; also orders s_key writes
        } else {
            __syncthreads();
        }
        for (int base = 0; base < n_scan; base += kSelThreads) {
            const int i = base + tid;
            uint32_t key = 0;
            bool p = i < n_scan;
            if (p) { key = use_pend ? s_pend[i] : key_at(i); p = (key & mask) == prefix; }
            const unsigned int bal = __ballot_sync(0xffffffffu, p);
            if (p) {
                const unsigned int bin = (key >> shift) & 255u;
                const unsigned int peers = __match_any_sync(bal, bin);
                if (lane == __ffs(peers) - 1) atomicAdd(&sm.lhist[warp & (kHistCopies - 1)][bin], (unsigned int)__popc(peers));
            }
        }
        __syncthreads();
        if (tid < 256) {
            unsigned int c = 0;
#pragma unroll
            for (int r = 0; r < kHistCopies; ++r) c += sm.lhist[r][tid];
            sm.lsum[tid] = c;
            if (c) atomicAdd(hist0 + tid, c);                                           // DSMEM
        }
        cluster.sync();
        if (warp == 0) {
            // lane l owns bins [8l, 8l+8); walk from the top bin down
            unsigned int h[8], mine = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist0[lane * 8 + j]; mine += h[j]; }
            unsigned int above = 0;                           // keys in the bins of higher lanes (uniform shuffle loop)
            for (int l = 31; l >= 0; --l) {
                const unsigned int v = __shfl_sync(0xffffffffu, mine, l);
                if (l > lane) above += v;
            }
            if (above < remaining && remaining <= above + mine) {
                unsigned int a = above;
                for (int j = 7; j >= 0; --j) {
                    if (remaining <= a + h[j]) { sm.sel_bin = lane * 8 + j; sm.sel_above = a; break; }
                    a += h[j];
                }
            }
        }
        __syncthreads();
        const unsigned int sel_bin = sm.sel_bin;
        remaining -= sm.sel_above;
        prefix |= sel_bin << shift;
        mask |= 255u << shift;
        if (warp == 1) {
            // this CTA's own keys above the chosen digit are selected for good; the last pass also yields its ties
            unsigned int above = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) above += (unsigned int)(lane * 8 + j) > sel_bin ? sm.lsum[lane * 8 + j] : 0u;
            above = __reduce_add_sync(0xffffffffu, above);
            if (lane == 0) {
                sm.loc_gt = (pass == 0 ? 0u : sm.loc_gt) + above;
                sm.loc_eq = sm.lsum[sel_bin];
                sm.pend_ctr = 0;
            }
        }
        __syncthreads();
        if (pass == 0 && (int)sm.lsum[sel_bin] <= pend_cap) {
            // compact the keys that share the chosen top byte (order is irrelevant for the histograms)
            for (int base = 0; base < n_local; base += kSelThreads) {
                const int i = base + tid;
                uint32_t key = 0;
                if (i < n_local) key = key_at(i);
                const bool p = i < n_local && (key & mask) == prefix;
                const unsigned int bal = __ballot_sync(0xffffffffu, p);
                if (p) {
                    unsigned int pos = 0;
                    const int leader = __ffs(bal) - 1;
                    if (lane == leader) pos = atomicAdd(&sm.pend_ctr, (unsigned int)__popc(bal));
                    pos = __shfl_sync(bal, pos, leader) + __popc(bal & ((1u << lane) - 1u));
                    s_pend[pos] = key;
                }
            }
            use_pend = true;
            n_scan = (int)sm.lsum[sel_bin];
            __syncthreads();
        }
    }
    const uint32_t thr_key = prefix;                         // keys > thr_key are all taken; `remaining` ties are needed

    // ---- counts per CTA (accumulated from the histograms above) -> every CTA knows every CTA's counts ----------
    if (tid == 0) sm.ctr = 0;
    if (tid < kSelCluster) {
        SelSmem *peer = cluster.map_shared_rank(&sm, tid);
        peer->n_gt[rank] = sm.loc_gt;
        peer->n_eq[rank] = sm.loc_eq;
    }
    cluster.sync();
    // Candidates are re-dealt evenly over the cluster before sorting (RPN scores cluster spatially, so one CTA's
    // slice may hold most of them): CTA r's candidates take the virtual slots [base, base + count) of a k_eff-long
    // array, and virtual slot v lives in CTA v / per at index v % per.
    unsigned int eq_before = 0, base_pos = 0, my_gt = 0, quota = 0;
#pragma unroll
    for (int r = 0; r < kSelCluster; ++r) {
        const unsigned int q = min(sm.n_eq[r], remaining > eq_before ? remaining - eq_before : 0u);
        if (r == rank) { my_gt = sm.n_gt[r]; quota = q; }
        if (r < rank) base_pos += sm.n_gt[r] + q;
        eq_before += sm.n_eq[r];
    }
    const unsigned int per = (unsigned int)ceil_div(k_eff, kSelCluster);
    unsigned int n_sel[kSelCluster];                           // candidates per CTA after the deal
#pragma unroll
    for (int r = 0; r < kSelCluster; ++r)
        n_sel[r] = (unsigned int)max(0, min((int)per, k_eff - r * (int)per));
    const unsigned int n_mine = (unsigned int)max(0, min((int)per, k_eff - rank * (int)per));
    auto deal = [&](unsigned int v, unsigned long long comp) {
        cluster.map_shared_rank(&sm, v / per)->sort[v % per] = comp;                    // DSMEM store
    };
    __syncthreads();

    // ---- compaction: keys above the threshold in any order (they are sorted next); ties: the first `quota` in
    // anchor order -------------------------------------------------------------------------------------------
    unsigned int eq_run = 0;                                   // ties seen so far in this CTA (uniform)
    for (int base = 0; base < n_local; base += kSelThreads) {
        const int i = base + tid;
        uint32_t key = 0;
        if (i < n_local) key = key_at(i);
        const bool gt = i < n_local && key > thr_key;
        const bool eq = i < n_local && key == thr_key && eq_run < quota;
        const unsigned long long comp = ((unsigned long long)key << 32) | (0xffffffffu - (uint32_t)(lo + i));
        const unsigned int bal_gt = __ballot_sync(0xffffffffu, gt);
        if (gt) {
            unsigned int pos = 0;
            const int leader = __ffs(bal_gt) - 1;
            if (lane == leader) pos = atomicAdd(&sm.ctr, (unsigned int)__popc(bal_gt));
            pos = __shfl_sync(bal_gt, pos, leader) + __popc(bal_gt & ((1u << lane) - 1u));
            deal(base_pos + pos, comp);
        }
        if (eq_run < quota && __syncthreads_or(eq)) {          // uniform; ties are rare: one barrier tells that none is here
            const unsigned int bal_eq = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) sm.warp_cnt[warp] = __popc(bal_eq);
            __syncthreads();
            unsigned int before = 0, total = 0;
            for (int wi = 0; wi < kSelThreads / 32; ++wi) {
                const unsigned int c = sm.warp_cnt[wi];
                before += wi < warp ? c : 0u;
                total += c;
            }
            const unsigned int r = eq_run + before + __popc(bal_eq & ((1u << lane) - 1u));
            if (eq && r < quota) deal(base_pos + my_gt + r, comp);
            eq_run += total;
            __syncthreads();
        }
    }
    cluster.sync();                                            // every dealt candidate has landed
    // ---- local sort, descending by (key, ~anchor) = score descending, anchor ascending --------------------------
    // every real word is > 0 (~anchor >= 1), so zero padding sorts last; the padded length is a power of two
    auto pad_len = [](unsigned int c) { int p = 2; while (p < (int)c) p <<= 1; return p; };
    const int n_pad = pad_len(n_mine);
    for (int i = (int)n_mine + tid; i < n_pad; i += kSelThreads) sm.sort[i] = 0ull;
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n_pad >> 1); t += kSelThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));       // index with bit j clear
                const int l = i | j;
                const unsigned long long a = sm.sort[i], b = sm.sort[l];
                const bool desc = (i & k) == 0;
                if (desc ? a < b : a > b) { sm.sort[i] = b; sm.sort[l] = a; }
            }
            __syncthreads();
        }
    }
    cluster.sync();                                            // every CTA's sorted list is visible cluster-wide

    // ---- merge by ranking: position = own index + number of larger words in each of the other seven lists ------
    // (all words are distinct: the anchor index is part of them); the seven binary searches run interleaved so
    // that their DSMEM round trips overlap.  Then refine / clip / normalise the box (modified_dense_model.py:
    // 179-218, :287) and write it to its final, score-ordered slot.
    const unsigned long long *lists[kSelCluster];
    int pads[kSelCluster], max_pad = 0;
#pragma unroll
    for (int r = 0; r < kSelCluster; ++r) {
        lists[r] = cluster.map_shared_rank(&sm, r)->sort;
        pads[r] = (r == rank || n_sel[r] == 0) ? 0 : pad_len(n_sel[r]);
        max_pad = max(max_pad, pads[r]);
    }
    for (int e = tid; e < (int)n_mine; e += kSelThreads) {
        const unsigned long long comp = sm.sort[e];
        int pos[kSelCluster];
#pragma unroll
        for (int r = 0; r < kSelCluster; ++r) pos[r] = 0;
        for (int step = max_pad >> 1; step > 0; step >>= 1) {
            unsigned long long v[kSelCluster];
#pragma unroll
            for (int r = 0; r < kSelCluster; ++r) v[r] = step < pads[r] ? lists[r][pos[r] + step - 1] : 0ull;
#pragma unroll
            for (int r = 0; r < kSelCluster; ++r) pos[r] += v[r] > comp ? step : 0;
        }
        int final_pos = e;
#pragma unroll
        for (int r = 0; r < kSelCluster; ++r)
            if (pads[r]) final_pos += pos[r] + (lists[r][pos[r]] > comp ? 1 : 0);

        const uint32_t a_idx = 0xffffffffu - (uint32_t)comp;
        const float4 a = __ldg(anchors + a_idx);               // (y1, x1, y2, x2) pixels
        float4 d = __ldg(rpn_bbox + (long long)img * n_anchors + a_idx);
        d.x = __fmul_rn(d.x, std_dev.x); d.y = __fmul_rn(d.y, std_dev.y);
        d.z = __fmul_rn(d.z, std_dev.z); d.w = __fmul_rn(d.w, std_dev.w);
        float height = __fsub_rn(a.z, a.x), width = __fsub_rn(a.w, a.y);
        float cy = __fadd_rn(a.x, __fmul_rn(0.5f, height)), cx = __fadd_rn(a.y, __fmul_rn(0.5f, width));
        cy = __fadd_rn(cy, __fmul_rn(d.x, height));
        cx = __fadd_rn(cx, __fmul_rn(d.y, width));
        height = __fmul_rn(height, (float)exp((double)d.z));
        width = __fmul_rn(width, (float)exp((double)d.w));
        float y1 = __fsub_rn(cy, __fmul_rn(0.5f, height)), x1 = __fsub_rn(cx, __fmul_rn(0.5f, width));
        float y2 = __fadd_rn(y1, height), x2 = __fadd_rn(x1, width);
        y1 = std_max(std_min(y1, img_h), 0.f); x1 = std_max(std_min(x1, img_w), 0.f);
        y2 = std_max(std_min(y2, img_h), 0.f); x2 = std_max(std_min(x2, img_w), 0.f);
        ws_boxes[(long long)img * k_eff + final_pos] = make_float4(__fdiv_rn(y1, img_h), __fdiv_rn(x1, img_w),
                                                                   __fdiv_rn(y2, img_h), __fdiv_rn(x2, img_w));
        ws_index[(long long)img * k_eff + final_pos] = (int32_t)a_idx;
    }
    cluster.sync();                                            // no CTA leaves while its list may still be read
}

// tf.image.non_max_suppression's IOU() (non_max_suppression_op.cc): corners min/max-normalised by the caller.
struct NmsBox { float ymin, xmin, ymax, xmax, area; };

__device__ __forceinline__ NmsBox nms_box(float4 b) {
    NmsBox r;
    r.ymin = std_min(b.x, b.z); r.ymax = std_max(b.x, b.z);
    r.xmin = std_min(b.y, b.w); r.xmax = std_max(b.y, b.w);
    r.area = __fmul_rn(__fsub_rn(r.ymax, r.ymin), __fsub_rn(r.xmax, r.xmin));
    return r;
}

constexpr int kMaskColBlocks = 4;                               // a CTA covers 64 rows x (4 x 64) columns

// mask[img][i][cb] bit t: box cb*64+t (later than i in score order) has IoU(i, .) > thr.  Only the upper triangle
// (cb >= i / 64) is written and read.
//
// Inner loop diet (this kernel is issue-bound: 18 M box pairs per image):
//  * fminf/fmaxf instead of the std::min/std::max selects: they differ only when a NaN is involved, and a box with
//    a NaN corner has area NaN or 0 under either rule, i.e. it never suppresses and is never suppressed;
//  * `inter / denom > thr` is decided without the division whenever inter is clear of thr * denom by more than the
//    rounding of both sides (1e-6 relative); only the remaining sliver takes the IEEE division, so the decision is
//    bit-identical to the fp32 quotient's.
// The rows are processed in BANDS of row blocks (rb0 .. rb0 + gridDim.y), each band followed by the scan over the
// same rows: once an image has its proposal_count survivors (`done`), the later bands of that image return at once
// -- the scan stops early on real RPN outputs, and the mask rows past that point would never be read.
__global__ void __launch_bounds__(64 * kMaskColBlocks) proposal_iou_mask_kernel(const float4 *__restrict__ ws_boxes, int n, int n_blk,
                                                                                float thr, unsigned long long *__restrict__ mask,
                                                                                int rb0, const int32_t *__restrict__ done) {
    const int rb = rb0 + blockIdx.y, img = blockIdx.z, t = threadIdx.x & 63, sub = threadIdx.x >> 6;
    const int cb0 = blockIdx.x * kMaskColBlocks;
    if (cb0 + kMaskColBlocks - 1 < rb) return;                  // whole tile below the diagonal
    if (done && done[img]) return;                              // this image already has all its proposals
    __shared__ float4 s_box[kMaskColBlocks][64];                // (ymin, xmin, ymax, xmax)
    __shared__ float s_area[kMaskColBlocks][64];
    const float4 *boxes = ws_boxes + (long long)img * n;
    const int cb = cb0 + sub;
    const int cj = cb * 64 + t;
    const NmsBox c = nms_box(cj < n ? __ldg(boxes + cj) : make_float4(0.f, 0.f, 0.f, 0.f));
    // columns that can never be suppressed (past the end, empty or NaN boxes: IOU() returns 0 for them and the
    // threshold is >= 0) become an empty box at +inf: intersection exactly 0 with everything
    const bool valid = cj < n && c.area > 0.f;
    s_box[sub][t] = valid ? make_float4(c.ymin, c.xmin, c.ymax, c.xmax) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    s_area[sub][t] = valid ? c.area : 0.f;
    __syncthreads();
    const int i = rb * 64 + t;
    if (i >= n || cb < rb || cb >= n_blk) return;
    const NmsBox r = nms_box(__ldg(boxes + i));
    unsigned long long bits = 0, unsure = 0;
    if (r.area > 0.f) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const float4 q = s_box[sub][j];
            const float ih = fmaxf(__fsub_rn(fminf(r.ymax, q.z), fmaxf(r.ymin, q.x)), 0.f);
            const float iw = fmaxf(__fsub_rn(fminf(r.xmax, q.w), fmaxf(r.xmin, q.y)), 0.f);
            const float inter = __fmul_rn(ih, iw);
            const float denom = __fsub_rn(__fadd_rn(r.area, s_area[sub][j]), inter);
            const float p = __fmul_rn(thr, denom);
            const bool over = inter > __fmaf_rn(p, 1.000001f, 1e-30f);
            const bool under = inter < __fmul_rn(p, 0.999999f) || inter == 0.f;
            if (over) bits |= 1ull << j;
            if (!over && !under) unsure |= 1ull << j;
        }
        while (unsure) {                                         // the sliver around the threshold: IEEE division
            const int j = __ffsll((long long)unsure) - 1;
            unsure &= unsure - 1;
            const float4 q = s_box[sub][j];
            const float ih = fmaxf(__fsub_rn(fminf(r.ymax, q.z), fmaxf(r.ymin, q.x)), 0.f);
            const float iw = fmaxf(__fsub_rn(fminf(r.xmax, q.w), fmaxf(r.xmin, q.y)), 0.f);
            const float inter = __fmul_rn(ih, iw);
            const float denom = __fsub_rn(__fadd_rn(r.area, s_area[sub][j]), inter);
            if (__fdiv_rn(inter, denom) > thr) bits |= 1ull << j;
        }
        if (cb == rb) bits &= ~((2ull << t) - 1ull);            // only boxes later in score order
    }
    mask[((long long)img * n + i) * n_blk + cb] = bits;
}

constexpr int kScanBatch = 8;                                   // mask words in flight per thread in the OR phase

// Scan state carried from one band to the next (global memory, per image).
struct ScanState {
    unsigned long long *removed;   // [n_images, n_blk]
    int32_t *keep;                 // [n_images, proposal_count] sorted positions of the survivors
    int32_t *count;                // [n_images]
    int32_t *done;                 // [n_images] 1 once the output has been written
};

__global__ void __launch_bounds__(kScanThreads) proposal_nms_scan_kernel(const float4 *__restrict__ ws_boxes,
                                                                         const int32_t *__restrict__ ws_index,
                                                                         const unsigned long long *__restrict__ mask, int n,
                                                                         int n_blk, int proposal_count, int blk0, int blk1,
                                                                         ScanState st, float4 *__restrict__ proposals,
                                                                         int32_t *__restrict__ n_valid,
                                                                         int32_t *__restrict__ anchor_index) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long *s_removed = reinterpret_cast<unsigned long long *>(sm_raw);    // [n_blk]
    int *s_keep = reinterpret_cast<int *>(s_removed + n_blk);                            // [proposal_count]
    __shared__ unsigned long long s_diag[2][64];
    __shared__ unsigned long long s_kept_bits;
    __shared__ int s_count;
    const int img = blockIdx.x, tid = threadIdx.x;
    const bool first = blk0 == 0, last = blk1 >= n_blk;
    if (!first && st.done[img]) return;                        // finished in an earlier band (uniform)
    const unsigned long long *m = mask + (long long)img * n * n_blk;
    int count0 = first ? 0 : st.count[img];                    // survivors before this block (replicated, uniform)
    for (int c = tid; c < n_blk; c += kScanThreads) s_removed[c] = first ? 0ull : st.removed[(long long)img * n_blk + c];
    if (!first)
        for (int j = tid; j < count0; j += kScanThreads) s_keep[j] = st.keep[(long long)img * proposal_count + j];
    if (tid == 0) s_count = count0;
    if (tid < 64) s_diag[blk0 & 1][tid] = blk0 * 64 + tid < n ? __ldg(m + (long long)(blk0 * 64 + tid) * n_blk + blk0) : 0ull;
    __syncthreads();
    for (int blk = blk0; blk < blk1; ++blk) {
        const int i0 = blk * 64, rows = min(64, n - i0);
        // the next block's diagonal words do not depend on this block's outcome: fetch them under the serial chain
        // (within the band: the next band's mask rows do not exist yet)
        unsigned long long next_diag = 0ull;
        if (tid < 64 && blk + 1 < blk1 && i0 + 64 + tid < n) next_diag = __ldg(m + (long long)(i0 + 64 + tid) * n_blk + blk + 1);
        if (tid == 0) {
            unsigned long long r = s_removed[blk], kept = 0ull;
            int count = count0;
            const unsigned long long *diag = s_diag[blk & 1];
            // eight candidates per round: their diagonal words are fetched together (independent shared-memory loads),
            // so the dependency chain between candidates is register arithmetic only
            for (int t0 = 0; t0 < rows && count < proposal_count; t0 += 8) {
                unsigned long long d[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) d[q] = diag[t0 + q];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int t = t0 + q;
                    if (t < rows && !((r >> t) & 1ull) && count < proposal_count) {
                        kept |= 1ull << t;
                        ++count;
                        r |= d[q];
                    }
                }
            }
            s_kept_bits = kept;
            s_count = count;
        }
        if (tid < 64) s_diag[(blk + 1) & 1][tid] = next_diag;
        __syncthreads();
        const unsigned long long kept = s_kept_bits;
        const int count = s_count;
        if (tid < 64 && ((kept >> tid) & 1ull)) s_keep[count0 + __popcll(kept & ((1ull << tid) - 1ull))] = i0 + tid;
        if (count >= proposal_count) { count0 = count; break; }  // uniform
        __syncthreads();                                       // s_keep of this block is read below
        // OR the survivors' mask rows into the removed set of the later blocks: one (row, column) word per thread
        // and round, kScanBatch independent fetches in flight per thread
        {
            const int n_kept = count - count0, n_cols = n_blk - blk - 1, total = n_kept * n_cols;
            for (int base = tid; base < total; base += kScanThreads * kScanBatch) {
                unsigned long long w[kScanBatch];
                int col[kScanBatch];
#pragma unroll
                for (int q = 0; q < kScanBatch; ++q) {
                    const int idx = base + q * kScanThreads;
                    w[q] = 0ull;
                    col[q] = 0;
                    if (idx < total) {
                        const int row = s_keep[count0 + idx / n_cols];
                        col[q] = blk + 1 + idx % n_cols;
                        w[q] = __ldg(m + (long long)row * n_blk + col[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < kScanBatch; ++q)
                    if (w[q]) atomicOr(&s_removed[col[q]], w[q]);
            }
        }
        count0 = count;
        __syncthreads();
    }
    __syncthreads();
    const int count = count0;
    if (count < proposal_count && !last) {                     // carry the state into the next band
        for (int c = tid; c < n_blk; c += kScanThreads) st.removed[(long long)img * n_blk + c] = s_removed[c];
        for (int j = tid; j < count; j += kScanThreads) st.keep[(long long)img * proposal_count + j] = s_keep[j];
        if (tid == 0) { st.count[img] = count; st.done[img] = 0; }
        return;
    }
    for (int j = tid; j < proposal_count; j += kScanThreads) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int a = -1;
        if (j < count) {
            const int i = s_keep[j];
            v = ws_boxes[(long long)img * n + i];
            a = ws_index[(long long)img * n + i];
        }
        proposals[(long long)img * proposal_count + j] = v;
        if (anchor_index) anchor_index[(long long)img * proposal_count + j] = a;
    }
    if (tid == 0) {
        if (n_valid) n_valid[img] = count;
        st.done[img] = 1;
    }
}

__global__ void normalize_boxes_kernel(const float4 *__restrict__ in, long long n, float h, float w, float4 *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = in[i];
    out[i] = make_float4(__fdiv_rn(b.x, h), __fdiv_rn(b.y, w), __fdiv_rn(b.z, h), __fdiv_rn(b.w, w));
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace dcap

using namespace dcap;

extern "C" size_t dc_proposal_workspace_bytes(int n_images, int n_anchors, int pre_nms_limit, int proposal_count) {
    if (n_images <= 0 || n_anchors <= 0 || pre_nms_limit <= 0 || proposal_count <= 0) return 0;
    const size_t k = (size_t)(pre_nms_limit < n_anchors ? pre_nms_limit : n_anchors);
    const size_t n_blk = (k + 63) / 64;
    return align256((size_t)n_images * k * sizeof(float4)) + align256((size_t)n_images * k * sizeof(int32_t)) +
           align256((size_t)n_images * k * n_blk * sizeof(unsigned long long)) +
           align256((size_t)n_images * n_blk * sizeof(unsigned long long)) +                // scan state: removed set,
           align256((size_t)n_images * (size_t)proposal_count * sizeof(int32_t)) +          // survivors so far,
           2 * align256((size_t)n_images * sizeof(int32_t));                                // count, done
}

extern "C" int dc_proposal_layer(const float *rpn_probs, const float *rpn_bbox, const float *anchors, int n_images,
                                 int n_anchors, const float *bbox_std_dev, float image_h, float image_w, int pre_nms_limit,
                                 int proposal_count, float nms_threshold, float *proposals, int32_t *n_valid,
                                 int32_t *anchor_index, void *workspace, size_t workspace_bytes, void *stream) {
    DC_REQUIRE(n_images >= 0 && n_anchors >= 1 && proposal_count >= 1 && pre_nms_limit >= 1, "bad n_images / n_anchors / counts");
    if (n_images == 0) return DC_OK;
    DC_REQUIRE(rpn_probs && rpn_bbox && anchors && bbox_std_dev && proposals && workspace, "null pointer argument");
    DC_REQUIRE((((uintptr_t)rpn_probs & 7) | ((uintptr_t)rpn_bbox & 15) | ((uintptr_t)anchors & 15) | ((uintptr_t)proposals & 15) |
                ((uintptr_t)workspace & 15)) == 0, "rpn_probs must be 8-byte, rpn_bbox / anchors / proposals / workspace 16-byte aligned");
    DC_REQUIRE(image_h > 0.f && image_w > 0.f, "image size must be positive");
    DC_REQUIRE(nms_threshold >= 0.f && nms_threshold <= 1.f, "nms_threshold must lie in [0, 1] (as tf.image.non_max_suppression requires)");
    const int k = pre_nms_limit < n_anchors ? pre_nms_limit : n_anchors;
    DC_REQUIRE(k <= kMaxPreNms, "pre_nms_limit=%d exceeds %d", k, kMaxPreNms);
    DC_REQUIRE(workspace_bytes >= dc_proposal_workspace_bytes(n_images, n_anchors, pre_nms_limit, proposal_count),
               "workspace too small: %zu < %zu bytes", workspace_bytes,
               dc_proposal_workspace_bytes(n_images, n_anchors, pre_nms_limit, proposal_count));
    const int n_blk = (k + 63) / 64;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    float4 *ws_boxes = reinterpret_cast<float4 *>(ws);
    ws += align256((size_t)n_images * k * sizeof(float4));
    int32_t *ws_index = reinterpret_cast<int32_t *>(ws);
    ws += align256((size_t)n_images * k * sizeof(int32_t));
    unsigned long long *ws_mask = reinterpret_cast<unsigned long long *>(ws);
    ws += align256((size_t)n_images * k * n_blk * sizeof(unsigned long long));
    ScanState st;
    st.removed = reinterpret_cast<unsigned long long *>(ws);
    ws += align256((size_t)n_images * n_blk * sizeof(unsigned long long));
    st.keep = reinterpret_cast<int32_t *>(ws);
    ws += align256((size_t)n_images * (size_t)proposal_count * sizeof(int32_t));
    st.count = reinterpret_cast<int32_t *>(ws);
    ws += align256((size_t)n_images * sizeof(int32_t));
    st.done = reinterpret_cast<int32_t *>(ws);
    cudaStream_t s = (cudaStream_t)stream;

    const int chunk = (n_anchors + kSelCluster - 1) / kSelCluster;
    const bool cache = chunk <= kSelMaxChunk;
    const size_t key_bytes = cache ? (size_t)((chunk + 3) & ~3) * sizeof(uint32_t) : 0;
    const size_t smem_max = sizeof(SelSmem) + (size_t)kSelMaxChunk * sizeof(uint32_t);           // what the kernels are opted into
    const int pend_cap = (int)std::min<size_t>(kPendCap, (smem_max - sizeof(SelSmem) - key_bytes) / sizeof(uint32_t));
    const size_t smem = sizeof(SelSmem) + key_bytes + (size_t)pend_cap * sizeof(uint32_t);
    static std::atomic<unsigned long long> attr_set{0};
    DC_CHECK_CUDA(once_per_device(attr_set, [] {
        const cudaError_t e = cudaFuncSetAttribute(proposal_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(sizeof(SelSmem) + (size_t)kSelMaxChunk * sizeof(uint32_t)));
        return e != cudaSuccess ? e : cudaFuncSetAttribute(proposal_select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           (int)(sizeof(SelSmem) + (size_t)kSelMaxChunk * sizeof(uint32_t)));
    }));
    const float4 std_dev = make_float4(bbox_std_dev[0], bbox_std_dev[1], bbox_std_dev[2], bbox_std_dev[3]);
    if (cache)
        proposal_select_kernel<true><<<n_images * kSelCluster, kSelThreads, smem, s>>>(
            rpn_probs, reinterpret_cast<const float4 *>(rpn_bbox), reinterpret_cast<const float4 *>(anchors), n_anchors, k, std_dev,
            image_h, image_w, ws_boxes, ws_index, pend_cap);
    else
        proposal_select_kernel<false><<<n_images * kSelCluster, kSelThreads, smem, s>>>(
            rpn_probs, reinterpret_cast<const float4 *>(rpn_bbox), reinterpret_cast<const float4 *>(anchors), n_anchors, k, std_dev,
            image_h, image_w, ws_boxes, ws_index, pend_cap);
    DC_CHECK_LAUNCH();
    const size_t scan_smem = (size_t)n_blk * sizeof(unsigned long long) + (size_t)proposal_count * sizeof(int);
    DC_REQUIRE(scan_smem <= 48 * 1024, "proposal_count=%d too large", proposal_count);
    // bands of kBandBlocks row blocks: IoU masks of the band, then the scan over it; later bands of an image that
    // already has its proposal_count survivors return immediately
    for (int b0 = 0, band = 0; b0 < n_blk; ++band) {
        const int b1 = (band == kMaxBands - 1 || b0 + kBandBlocks >= n_blk) ? n_blk : b0 + kBandBlocks;
        proposal_iou_mask_kernel<<<dim3((n_blk + kMaskColBlocks - 1) / kMaskColBlocks, b1 - b0, n_images), 64 * kMaskColBlocks, 0, s>>>(
            ws_boxes, k, n_blk, nms_threshold, ws_mask, b0, b0 == 0 ? nullptr : st.done);
        DC_CHECK_LAUNCH();
        proposal_nms_scan_kernel<<<n_images, kScanThreads, scan_smem, s>>>(ws_boxes, ws_index, ws_mask, k, n_blk, proposal_count, b0, b1, st,
                                                                          reinterpret_cast<float4 *>(proposals), n_valid, anchor_index);
        DC_CHECK_LAUNCH();
        b0 = b1;
    }
    return DC_OK;
}

extern "C" int dc_normalize_boxes(const float *boxes, int64_t n_boxes, float image_h, float image_w, float *out, void *stream) {
    DC_REQUIRE(n_boxes >= 0 && image_h > 0.f && image_w > 0.f, "bad n_boxes / image size");
    if (n_boxes == 0) return DC_OK;
    DC_REQUIRE(boxes && out && (((uintptr_t)boxes | (uintptr_t)out) & 15) == 0, "boxes / out must be non-null and 16-byte aligned");
    normalize_boxes_kernel<<<(unsigned int)((n_boxes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(boxes), n_boxes, image_h, image_w, reinterpret_cast<float4 *>(out));
    DC_CHECK_LAUNCH();
    return DC_OK;
}

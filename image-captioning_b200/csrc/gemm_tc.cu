// bf16 GEMM on the 5th-generation tensor cores (sm_100a): tcgen05.mma issued by one thread,
// operands staged in shared memory by TMA (128-byte swizzle), fp32 accumulators in TMEM
// (double buffered, so the epilogue of tile i overlaps the MMAs of tile i+1), persistent over
// tiles with one CTA per SM.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )        A, B bf16, K contiguous ("K-major") in both
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (each owns the 32 TMEM lanes 32*(warp%4)..+31, i.e. 32 rows of the tile).
//
// Epilogues
//   kEpiStore : v = acc (+bias[n]) (+addend[m,n]); v = v*scale[n]+shift[n]; relu; store fp32 and/or
//               bf16 (the bf16 copy is the next GEMM's A operand)
//   kEpiArgmax: v = acc + bias[n]; per row and per N-tile (max, first arg-max, sum exp(v-max))
//               -> partial[m, n_tile]; the [M,V] logits are never written (greedy decoding)
//   kEpiTopK  : like kEpiArgmax but keeps the k best (value, index) pairs per row and tile
#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <map>
#include <mutex>

namespace dcap {


// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------

struct TcGeom {
    int M, N, K;
    int a_mn, b_mn;          // operand majorness
    int splits, kb_per;      // split-K: k-blocks [s*kb_per, min((s+1)*kb_per, num_kb)) per work unit
    int split_major;         // work-unit order: 1 = unit -> (split, tile) with CTAs that run together sharing a K range
    int tma_out;             // store epilogue output path: 0 = per-thread stores, 1 = fp32 via TMA, 2 = bf16 via TMA
};

constexpr int kCellAddBytes = 128 * 128 * 4;      // fp32 addend of a 128 x 128 tile: 4 boxes of [128 rows x 32 cols]
constexpr int kCellCBytes = 128 * 32 * 4;         // fp32 cell state of the tile's 32 units

template <int kBlockN, int kEpi = kEpiStore>
struct TcSmem {
    static constexpr bool kCellTma = kEpi == kEpiCellTma;
    static constexpr int kStageA = kBlockM * kBlockK * 2;
    static constexpr int kStageB = kBlockN * kBlockK * 2;
    static constexpr int kStages = (kBlockN == 256 || kCellTma) ? 4 : 6;
    static constexpr int kBarOff = kStages * (kStageA + kStageB);
    static constexpr int kOutOff = kBarOff + 1024;                       // barriers live in their own 1 KB
    static constexpr int kBaseBytes = kOutOff + 1024 /*align*/ + (kCellTma ? kCellAddBytes + kCellCBytes : 0);
    static constexpr int kBytes = kBaseBytes + (kCellTma ? 0 : kEpiWarps * kOutStage);   // + output staging (TMA stores)
};


__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Addend of epilogue row `ri`: a float4 pointer and the stride (in float4) between consecutive column quads --
// row-major (stride 1) or blocked-32 (stride 32, see TcEpilogue::addend_blocked32).
struct AddendRow {
    const float4 *p;
    int stride;
    __device__ __forceinline__ float4 quad(int n) const { return __ldg(p + (long long)(n >> 2) * stride); }
    __device__ __forceinline__ float at(int n) const { return __ldg(reinterpret_cast<const float *>(p + (long long)(n >> 2) * stride) + (n & 3)); }
};
__device__ __forceinline__ AddendRow addend_row(const TcEpilogue &ep, long long mr) {
    AddendRow a;
    a.p = nullptr; a.stride = 1;
    if (!ep.addend) return a;
    const long long ri = ep.addend_mod > 0 ? mr % ep.addend_mod : (ep.addend_div > 0 ? mr / ep.addend_div : mr);
    if (ep.addend_blocked32) {
        a.p = reinterpret_cast<const float4 *>(ep.addend) + ((ri >> 5) * (ep.ld_addend >> 2)) * 32 + (ri & 31);
        a.stride = 32;
    } else {
        a.p = reinterpret_cast<const float4 *>(ep.addend + ri * ep.ld_addend);
    }
    return a;
}

// Top-k epilogue body for one (warp, column region): per row (lane) the running max / sum exp and the K best
// (value, index) pairs in descending lexicographic order (value, then index: among equal values the LARGER
// index ranks higher, as np.argsort(p)[-k:] of a stable sort does).  Lane = row, so a data-dependent
// insertion branch would be taken by some lane at almost every element; instead every element goes through a
// branch-free K-slot insertion network (K compares + 2K selects).  Two independent lists (even / odd columns)
// halve the dependent chain; they are merged once at the end of the region.
template <int K>
__device__ __forceinline__ void topk_insert_ge(float (&tv)[K], int (&ti)[K], float x, int n) {
    bool c[K];
#pragma unroll
    for (int q = 0; q < K; ++q) c[q] = x >= tv[q];                  // NaN (masked column) never enters; ties: the later column wins
#pragma unroll
    for (int q = K - 1; q > 0; --q) {
        tv[q] = c[q - 1] ? tv[q - 1] : (c[q] ? x : tv[q]);
        ti[q] = c[q - 1] ? ti[q - 1] : (c[q] ? n : ti[q]);
    }
    tv[0] = c[0] ? x : tv[0];
    ti[0] = c[0] ? n : ti[0];
}

template <int kCols, int K>
__device__ __forceinline__ void topk_region(const TcEpilogue &ep, uint32_t taddr, int lane, int m_base, int n_base, int M,
                                            int N, int part_slot, int part_slots) {
    float ta[K], tb[K];
    int ia[K], ib[K];
#pragma unroll
    for (int q = 0; q < K; ++q) { ta[q] = tb[q] = -INFINITY; ia[q] = ib[q] = -1; }
    float best = -INFINITY, sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kCols; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        const int nb = n_base + c0;
        if (nb < N) {
            float cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = nb + j;
                const float x = (n < N) ? v[j] + __ldg(ep.bias + n) : -INFINITY;
                v[j] = x;
                cmax = fmaxf(cmax, x);
            }
            if (cmax > best) { sum *= __expf(best - cmax); best = cmax; }
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                sum += __expf(v[j] - best) + __expf(v[j + 1] - best);
                topk_insert_ge<K>(ta, ia, nb + j < N ? v[j] : __int_as_float(0x7fc00000), nb + j);
                topk_insert_ge<K>(tb, ib, nb + j + 1 < N ? v[j + 1] : __int_as_float(0x7fc00000), nb + j + 1);
            }
        }
    }
    // merge the odd-column list into the even-column one (full lexicographic compare: indices interleave)
#pragma unroll
    for (int p = 0; p < K; ++p) {
        float iv = tb[p];
        int ii = ib[p];
#pragma unroll
        for (int q = 0; q < K; ++q) {
            const bool up = ii >= 0 && (iv > ta[q] || (iv == ta[q] && ii > ia[q]));
            const float t1 = ta[q]; const int t2 = ia[q];
            ta[q] = up ? iv : t1; ia[q] = up ? ii : t2;
            iv = up ? t1 : iv; ii = up ? t2 : ii;
        }
    }
    const int m = m_base + lane;
    if (m < M) {
        const int stride = 2 + 2 * ep.topk;
        float *dst = ep.partial + ((long long)m * part_slots + part_slot) * stride;
        dst[0] = best; dst[1] = sum;
#pragma unroll
        for (int q = 0; q < K; ++q)
            if (q < ep.topk) { dst[2 + 2 * q] = ta[q]; dst[3 + 2 * q] = __int_as_float(ia[q]); }
    }
}

// Epilogue of one (warp, column-half) region: rows = the warp's 32 TMEM lanes (lane i <-> row i),
// columns [n_base, n_base + kCols) of the tile, pulled 32 columns at a time with tcgen05.ld.
// Every variant is thread-per-row: a lane's global accesses are whole 32-byte sectors of its own
// row, and the operands of the FIRST chunk are requested before the wait on the accumulator
// barrier, those of chunk c+1 before chunk c is computed.
template <int kCols, int kEpi>
__device__ __forceinline__ void epilogue_region(const TcEpilogue &ep, uint32_t taddr, int lane, int m_base, int n_base,
                                                int M, int N, int part_slot, int part_slots, bool atomic,
                                                uint64_t *full_bar, uint32_t full_phase, uint8_t *stage,
                                                const CUtensorMap *map_o, int tma_out) {
    auto wait_acc = [&]() {
        mbar_wait(full_bar, full_phase);
        tc_fence_after();
    };
    if constexpr (kEpi == kEpiArgmax || kEpi == kEpiArgmaxSum) {
        wait_acc();
        // per-row statistics over this region: max, first arg-max (, sum exp(v - max))
        float best = -INFINITY, sum = 0.f;
        int best_i = 0x7fffffff;
#pragma unroll 1
        for (int c0 = 0; c0 < kCols; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
            const int nb = n_base + c0;
            if (nb < N) {
                float cmax = -INFINITY;
                int ci = 0x7fffffff;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = nb + j;
                    const float x = (n < N) ? v[j] + __ldg(ep.bias + n) : -INFINITY;
                    v[j] = x;
                    if (x > cmax) { cmax = x; ci = n; }            // strict > keeps the first index
                }
                if constexpr (kEpi == kEpiArgmaxSum) {
                    if (cmax > best) sum *= __expf(best - cmax);
                }
                if (cmax > best) { best = cmax; best_i = ci; }
                if constexpr (kEpi == kEpiArgmaxSum) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum += __expf(v[j] - best);
                }
            }
        }
        const int m = m_base + lane;
        if (m < M) {
            float4 *dst = reinterpret_cast<float4 *>(ep.partial) + (long long)m * part_slots + part_slot;
            *dst = make_float4(best, __int_as_float(best_i), sum, 0.f);
        }
    } else if constexpr (kEpi == kEpiTopK) {
        wait_acc();
        const int k = ep.topk;
        if (k <= 1) topk_region<kCols, 1>(ep, taddr, lane, m_base, n_base, M, N, part_slot, part_slots);
        else if (k == 2) topk_region<kCols, 2>(ep, taddr, lane, m_base, n_base, M, N, part_slot, part_slots);
        else if (k == 3) topk_region<kCols, 3>(ep, taddr, lane, m_base, n_base, M, N, part_slot, part_slots);
        else if (k == 4) topk_region<kCols, 4>(ep, taddr, lane, m_base, n_base, M, N, part_slot, part_slots);
        else topk_region<kCols, kTopKMax>(ep, taddr, lane, m_base, n_base, M, N, part_slot, part_slots);
    } else if constexpr (kEpi == kEpiStore || kEpi == kEpiStoreTmaF32 || kEpi == kEpiStoreTmaB16) {
        const int m = m_base + lane;
        const bool valid = m < M;
        const long long mr = valid ? m : (long long)(M - 1);
        const AddendRow add_row = addend_row(ep, mr);
        const __nv_bfloat16 *mask_row = ep.mask_src ? ep.mask_src + mr * ep.ld_mask : nullptr;
        const bool bf16_vec8 = ep.out_bf16 && ((ep.ld_bf16 & 7) == 0) && ((reinterpret_cast<uintptr_t>(ep.out_bf16) & 15) == 0);
        // Output path.  tma_out != 0: the warp's 32 x 32 fp32 (or 32 x 64 bf16) sub-tile is staged in a
        // 128-byte-swizzled shared-memory buffer and leaves the SM as ONE bulk tensor store (or fp32
        // reduce-add for accumulating outputs): full 128-byte lines per row instead of 32 scattered
        // 16-byte pieces per store instruction, and the bounds are clipped by the TMA unit.
        constexpr bool tma_f32 = kEpi == kEpiStoreTmaF32, tma_b16 = kEpi == kEpiStoreTmaB16;
        uint8_t *my_row = stage + lane * 128;
        float4 a_nxt[8];
        uint4 k_nxt[4];
        auto load_operands = [&](int nb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a_nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (add_row.p && nb + 4 * j + 4 <= N)
                    a_nxt[j] = add_row.quad(nb + 4 * j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                k_nxt[j] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);     // bf16 1.0: keep
                if (mask_row && nb + 8 * j + 8 <= N)
                    k_nxt[j] = __ldg(reinterpret_cast<const uint4 *>(mask_row + nb + 8 * j));
            }
        };
        if (n_base < N) load_operands(n_base);
        wait_acc();
#pragma unroll 1
        for (int c0 = 0; c0 < kCols; c0 += 32) {
            const int nb = n_base + c0;
            if (nb >= N) break;                                      // warp-uniform
            float4 a_cur[8];
            uint32_t k_cur[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) a_cur[j] = a_nxt[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                k_cur[4 * j] = k_nxt[j].x; k_cur[4 * j + 1] = k_nxt[j].y;
                k_cur[4 * j + 2] = k_nxt[j].z; k_cur[4 * j + 3] = k_nxt[j].w;
            }
            if (c0 + 32 < kCols && nb + 32 < N) load_operands(nb + 32);
            float v[32];
            tmem_ld32(taddr + c0, v);
            const bool full = nb + 32 <= N;
            if (full) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float xs[4] = {v[4 * j] + a_cur[j].x, v[4 * j + 1] + a_cur[j].y, v[4 * j + 2] + a_cur[j].z,
                                   v[4 * j + 3] + a_cur[j].w};
                    const int n = nb + 4 * j;
                    if (ep.bias) {
                        const float4 t = __ldg(reinterpret_cast<const float4 *>(ep.bias + n));
                        xs[0] += t.x; xs[1] += t.y; xs[2] += t.z; xs[3] += t.w;
                    }
                    if (ep.scale) {
                        const float4 sc = __ldg(reinterpret_cast<const float4 *>(ep.scale + n));
                        const float4 sh = __ldg(reinterpret_cast<const float4 *>(ep.shift + n));
                        xs[0] = xs[0] * sc.x + sh.x; xs[1] = xs[1] * sc.y + sh.y;
                        xs[2] = xs[2] * sc.z + sh.z; xs[3] = xs[3] * sc.w + sh.w;
                    }
                    if (ep.relu) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) xs[q] = fmaxf(xs[q], 0.f);
                    }
                    if (mask_row) {
                        // ReLU backward: the forward activation (bf16, >= 0) is positive iff it passed
                        const uint32_t w0 = k_cur[2 * j], w1 = k_cur[2 * j + 1];
                        if ((w0 & 0x7fffu) == 0 || (w0 & 0x8000u)) xs[0] = 0.f;
                        if ((w0 & 0x7fff0000u) == 0 || (w0 & 0x80000000u)) xs[1] = 0.f;
                        if ((w1 & 0x7fffu) == 0 || (w1 & 0x8000u)) xs[2] = 0.f;
                        if ((w1 & 0x7fff0000u) == 0 || (w1 & 0x80000000u)) xs[3] = 0.f;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[4 * j + q] = xs[q];
                }
            } else {
                // ragged last chunk of the last N tile: element-wise (never with deint); columns >= N become 0
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = nb + j;
                    float x = 0.f;
                    if (n < N) {
                        x = v[j];
                        if (add_row.p) x += add_row.at(n);
                        if (ep.bias) x += __ldg(ep.bias + n);
                        if (ep.scale) x = x * __ldg(ep.scale + n) + __ldg(ep.shift + n);
                        if (ep.relu) x = fmaxf(x, 0.f);
                        if (mask_row && !(__bfloat162float(mask_row[n]) > 0.f)) x = 0.f;
                    }
                    v[j] = x;
                }
            }
            // ---- fp32 output ----
            if constexpr (tma_f32) {
                if (lane == 0) tma_store_wait_read();                // previous bulk store has drained the buffer
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(my_row + ((j ^ (lane & 7)) << 4)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (atomic) tma_reduce_add_2d(map_o, stage, nb, m_base);
                    else tma_store_2d(map_o, stage, nb, m_base);
                    tma_store_commit();
                }
            } else if (!tma_b16 && ep.out_f32 && valid) {
                if (!full) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nb + j;
                        if (n >= N) continue;
                        if (atomic) atomicAdd(ep.out_f32 + (long long)m * ep.ld_f32 + n, v[j]);
                        else ep.out_f32[(long long)m * ep.ld_f32 + n] = v[j];
                    }
                } else if (ep.deint_units > 0) {
                    // columns 4u+g of this chunk -> 8 consecutive units of each of the 4 gate blocks
                    const int u0 = nb >> 2;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float *dst = ep.out_f32 + (long long)m * ep.ld_f32 + (long long)g * ep.deint_units + u0;
                        if (atomic) {
                            red_add_v4(dst, v[g], v[4 + g], v[8 + g], v[12 + g]);
                            red_add_v4(dst + 4, v[16 + g], v[20 + g], v[24 + g], v[28 + g]);
                        } else {
                            reinterpret_cast<float4 *>(dst)[0] = make_float4(v[g], v[4 + g], v[8 + g], v[12 + g]);
                            reinterpret_cast<float4 *>(dst)[1] = make_float4(v[16 + g], v[20 + g], v[24 + g], v[28 + g]);
                        }
                    }
                } else if (ep.blocked32) {
                    // 32 consecutive rows x one column quad are 512 contiguous bytes: the lanes' stores coalesce
                    float4 *dst = reinterpret_cast<float4 *>(ep.out_f32) + ((long long)(m >> 5) * (ep.ld_f32 >> 2) + (nb >> 2)) * 32 + (m & 31);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j * 32] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                    float *dst = ep.out_f32 + (long long)m * ep.ld_f32 + nb;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (atomic) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        else reinterpret_cast<float4 *>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            // ---- bf16 output ----
            if (tma_b16 || (!tma_f32 && ep.out_bf16 && valid)) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                    pk[j] = *reinterpret_cast<uint32_t *>(&t);
                }
                if constexpr (tma_b16) {
                    // two 32-column chunks share one 64-column (128-byte) staging row
                    const int half = (c0 >> 5) & 1;
                    if (half == 0) {
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4 *>(my_row + (((half * 4 + j) ^ (lane & 7)) << 4)) =
                            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    if (half == 1 || nb + 32 >= N) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(map_o, stage, nb - half * 32, m_base);
                            tma_store_commit();
                        }
                    }
                } else if (!full) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < N) ep.out_bf16[(long long)m * ep.ld_bf16 + nb + j] = __float2bfloat16_rn(v[j]);
                } else if constexpr (!tma_b16) {
                    __nv_bfloat16 *dst = ep.out_bf16 + (long long)m * ep.ld_bf16 + nb;
                    if (bf16_vec8) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            reinterpret_cast<uint4 *>(dst)[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<uint2 *>(dst)[j] = make_uint2(pk[2 * j], pk[2 * j + 1]);
                    }
                }
            }
        }
    } else {
        // kEpiCell: columns are gate-interleaved, so the 32 columns of a chunk are the (i,f,g,o)
        // pre-activations of 8 consecutive units of this lane's row.  Per chunk a lane reads 128 B of
        // addend + 32 B of c, writes 32 B of c and 16 B of h (whole sectors).
        const int m = m_base + lane;
        const bool valid = m < M;
        const long long mr = valid ? m : (long long)(M - 1);
        const AddendRow add_row = addend_row(ep, mr);
        const float *c_row = ep.cell_c + mr * ep.cell_units;
        float *c_dst = (ep.cell_c_out ? ep.cell_c_out : ep.cell_c) + mr * ep.cell_units;
        const bool masked = ep.cell_tok && __ldg(ep.cell_tok + mr) == 0;
        float4 a_nxt[8], c_nxt[2];
        uint4 h_nxt = make_uint4(0, 0, 0, 0);
        auto load_operands = [&](int nb) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                a_nxt[j] = add_row.p ? add_row.quad(nb + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            c_nxt[0] = *reinterpret_cast<const float4 *>(c_row + (nb >> 2));
            c_nxt[1] = *reinterpret_cast<const float4 *>(c_row + (nb >> 2) + 4);
            if (masked) h_nxt = *reinterpret_cast<const uint4 *>(ep.cell_h_prev + mr * ep.ld_h_prev + (nb >> 2));
        };
        if (n_base < N) load_operands(n_base);
        wait_acc();
#pragma unroll 1
        for (int c0 = 0; c0 < kCols; c0 += 32) {
            const int nb = n_base + c0;
            if (nb >= N) break;                                      // warp-uniform; N % 32 == 0 here
            float4 a_cur[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a_cur[j] = a_nxt[j];
            const float c_old[8] = {c_nxt[0].x, c_nxt[0].y, c_nxt[0].z, c_nxt[0].w,
                                    c_nxt[1].x, c_nxt[1].y, c_nxt[1].z, c_nxt[1].w};
            const uint32_t hw[4] = {h_nxt.x, h_nxt.y, h_nxt.z, h_nxt.w};
            if (c0 + 32 < kCols && nb + 32 < N) load_operands(nb + 32);
            float v[32];
            tmem_ld32(taddr + c0, v);
            const int u0 = nb >> 2;
            float c_new[8], h_new[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 z = make_float4(v[4 * j] + a_cur[j].x, v[4 * j + 1] + a_cur[j].y, v[4 * j + 2] + a_cur[j].z,
                                       v[4 * j + 3] + a_cur[j].w);
                if (ep.bias) {
                    const float4 t = __ldg(reinterpret_cast<const float4 *>(ep.bias + nb + 4 * j));
                    z.x += t.x; z.y += t.y; z.z += t.z; z.w += t.w;
                }
                const float ig = hard_sigmoid_tc(z.x), fg = hard_sigmoid_tc(z.y);
                const float gg = tanh_fast(z.z), og = hard_sigmoid_tc(z.w);
                // saved (bf16) copies: an unsaturated hard-sigmoid (< 1) must not round up to 1, the backward
                // pass reads "strictly inside (0,1)" as "derivative 0.2"
                v[4 * j] = ig < 1.f ? fminf(ig, 0.99609375f) : 1.f; v[4 * j + 1] = fg < 1.f ? fminf(fg, 0.99609375f) : 1.f;
                v[4 * j + 2] = gg; v[4 * j + 3] = og < 1.f ? fminf(og, 0.99609375f) : 1.f;
                c_new[j] = __fadd_rn(__fmul_rn(fg, c_old[j]), __fmul_rn(ig, gg));
                h_new[j] = __fmul_rn(og, tanh_fast(c_new[j]));
                if (masked) {                                        // K.rnn mask: carry (h, c)
                    c_new[j] = c_old[j];
                    h_new[j] = __uint_as_float((j & 1) ? (hw[j >> 1] & 0xffff0000u) : (hw[j >> 1] << 16));
                }
            }
            if (valid) {
                reinterpret_cast<float4 *>(c_dst + u0)[0] = make_float4(c_new[0], c_new[1], c_new[2], c_new[3]);
                reinterpret_cast<float4 *>(c_dst + u0)[1] = make_float4(c_new[4], c_new[5], c_new[6], c_new[7]);
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat162 t = __floats2bfloat162_rn(h_new[2 * j], h_new[2 * j + 1]);
                    pk[j] = *reinterpret_cast<uint32_t *>(&t);
                }
                const uint4 hv = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (ep.cell_h_a) *reinterpret_cast<uint4 *>(ep.cell_h_a + (long long)m * ep.ld_h_a + u0) = hv;
                if (ep.cell_h_b) *reinterpret_cast<uint4 *>(ep.cell_h_b + (long long)m * ep.ld_h_b + u0) = hv;
                if (ep.cell_gates_out) {                             // saved for the backward pass (bf16: 64 B per row and chunk)
                    uint4 *g = reinterpret_cast<uint4 *>(ep.cell_gates_out + (long long)m * ep.ld_gates_out + nb);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t w[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
                            w[q] = *reinterpret_cast<uint32_t *>(&t);
                        }
                        g[j] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }
}

// Fused Keras LSTM cell epilogue, shared-memory operand variant (kEpiCellTma, 128-wide tiles): the tile's fp32
// addend [128 x 128] and cell state [128 x 32 units] were staged by the TMA producer (128-byte swizzle) while the
// tile's MMAs ran, so the only global LOADS left on the epilogue's critical path are the rare masked rows'
// previous h.  Lane i owns row i of the warp's 32 rows; 16-byte piece j of a 128-byte staged row sits at
// piece (j ^ (row & 7)) -- conflict-free for 8 consecutive lanes.
template <int kCols>
__device__ __forceinline__ void epilogue_cell_tma(const TcEpilogue &ep, uint32_t taddr, int lane, int m_base, int n_base,
                                                  int col_in_tile, int row_in_tile, int M, int N, uint64_t *full_bar,
                                                  uint32_t full_phase, uint64_t *add_full, uint32_t add_phase,
                                                  const uint8_t *smem_add, const uint8_t *smem_c) {
    const int m = m_base + lane;
    const bool valid = m < M;
    const long long mr = valid ? m : (long long)(M - 1);
    float *c_dst = (ep.cell_c_out ? ep.cell_c_out : ep.cell_c) + mr * ep.cell_units;
    const bool masked = ep.cell_tok && __ldg(ep.cell_tok + mr) == 0;
    const bool has_add = ep.addend != nullptr;
    const uint32_t sw = (uint32_t)row_in_tile & 7u;
    const uint8_t *crow = smem_c + row_in_tile * 128;
    mbar_wait(add_full, add_phase);                                  // staged operands have landed
    mbar_wait(full_bar, full_phase);                                 // accumulator complete
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < kCols; c0 += 32) {
        const int nb = n_base + c0;
        if (nb >= N) break;                                          // warp-uniform; N % 32 == 0 here
        const int ct = col_in_tile + c0;                             // column inside the 128-wide tile
        const uint8_t *arow = smem_add + (ct >> 5) * (kCellAddBytes / 4) + row_in_tile * 128;
        float4 a_cur[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            a_cur[j] = has_add ? *reinterpret_cast<const float4 *>(arow + (((uint32_t)j ^ sw) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t cp = (uint32_t)ct >> 4;                       // first 16-byte piece of the chunk's 8 units
        const float4 c_lo = *reinterpret_cast<const float4 *>(crow + ((cp ^ sw) << 4));
        const float4 c_hi = *reinterpret_cast<const float4 *>(crow + (((cp + 1) ^ sw) << 4));
        const float c_old[8] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w, c_hi.x, c_hi.y, c_hi.z, c_hi.w};
        uint4 hq = make_uint4(0, 0, 0, 0);
        if (masked) hq = *reinterpret_cast<const uint4 *>(ep.cell_h_prev + mr * ep.ld_h_prev + (nb >> 2));
        const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
        float v[32];
        tmem_ld32(taddr + c0, v);
        const int u0 = nb >> 2;
        float c_new[8], h_new[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 z = make_float4(v[4 * j] + a_cur[j].x, v[4 * j + 1] + a_cur[j].y, v[4 * j + 2] + a_cur[j].z,
                                   v[4 * j + 3] + a_cur[j].w);
            if (ep.bias) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(ep.bias + nb + 4 * j));
                z.x += t.x; z.y += t.y; z.z += t.z; z.w += t.w;
            }
            const float ig = hard_sigmoid_tc(z.x), fg = hard_sigmoid_tc(z.y);
            const float gg = tanh_fast(z.z), og = hard_sigmoid_tc(z.w);
            v[4 * j] = ig < 1.f ? fminf(ig, 0.99609375f) : 1.f; v[4 * j + 1] = fg < 1.f ? fminf(fg, 0.99609375f) : 1.f;
            v[4 * j + 2] = gg; v[4 * j + 3] = og < 1.f ? fminf(og, 0.99609375f) : 1.f;
            c_new[j] = __fadd_rn(__fmul_rn(fg, c_old[j]), __fmul_rn(ig, gg));
            h_new[j] = __fmul_rn(og, tanh_fast(c_new[j]));
            if (masked) {                                            // K.rnn mask: carry (h, c)
                c_new[j] = c_old[j];
                h_new[j] = __uint_as_float((j & 1) ? (hw[j >> 1] & 0xffff0000u) : (hw[j >> 1] << 16));
            }
        }
        if (valid) {
            reinterpret_cast<float4 *>(c_dst + u0)[0] = make_float4(c_new[0], c_new[1], c_new[2], c_new[3]);
            reinterpret_cast<float4 *>(c_dst + u0)[1] = make_float4(c_new[4], c_new[5], c_new[6], c_new[7]);
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 t = __floats2bfloat162_rn(h_new[2 * j], h_new[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t *>(&t);
            }
            const uint4 hv = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (ep.cell_h_a) *reinterpret_cast<uint4 *>(ep.cell_h_a + (long long)m * ep.ld_h_a + u0) = hv;
            if (ep.cell_h_b) *reinterpret_cast<uint4 *>(ep.cell_h_b + (long long)m * ep.ld_h_b + u0) = hv;
            if (ep.cell_gates_out) {
                uint4 *g = reinterpret_cast<uint4 *>(ep.cell_gates_out + (long long)m * ep.ld_gates_out + nb);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
                        w[q] = *reinterpret_cast<uint32_t *>(&t);
                    }
                    g[j] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
}

template <int kBlockN, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_c,
                    const TcEpilogue ep, const TcGeom g) {
    using S = TcSmem<kBlockN, kEpi>;
    constexpr int kStages = S::kStages;
    constexpr bool kCellTma = S::kCellTma;
    constexpr uint32_t kTmemCols = 2 * kBlockN;                  // two accumulator buffers (power of 2)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + kStages * S::kStageA;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + S::kBarOff);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *tmem_full = empty_bar + kStages;
    uint64_t *tmem_empty = tmem_full + 2;
    uint64_t *add_full = tmem_empty + 2;                           // cell-TMA variant: addend + c tile landed / consumed
    uint64_t *add_empty = add_full + 1;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(add_empty + 1);
    uint8_t *smem_add = smem + S::kOutOff;                         // cell-TMA variant only (no output staging there)
    uint8_t *smem_c = smem_add + kCellAddBytes;

    const int M = g.M, N = g.N, K = g.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + kBlockN - 1) / kBlockN;
    const int tiles_m = (M + kBlockM - 1) / kBlockM;
    const int num_tiles = tiles_m * tiles_n;
    const int num_units = num_tiles * g.splits;          // unit = split * num_tiles + tile: CTAs running together share a K range (L2 reuse of the operand slabs)
    const int num_kb = (K + kBlockK - 1) / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (g.tma_out) tma_prefetch_desc(&map_o);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
        mbar_init(add_full, 1); mbar_init(add_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // everything above overlapped the previous kernel's tail (programmatic dependent launch); from here on
    // its results are read
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] uint32_t add_phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                const int split = g.split_major ? unit / num_tiles : unit % g.splits;
                const int tile = g.split_major ? unit - split * num_tiles : unit / g.splits;
                const int m0 = (tile / tiles_n) * kBlockM, n0 = (tile % tiles_n) * kBlockN;
                const int kb0 = split * g.kb_per, kb1 = min(kb0 + g.kb_per, num_kb);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], S::kStageA + S::kStageB);
                    uint8_t *da = smem_a + stage * S::kStageA, *db = smem_b + stage * S::kStageB;
                    if (!g.a_mn) {
                        tma_load_2d(&map_a, &full_bar[stage], da, kb * kBlockK, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockM / 64; ++c)
                            tma_load_2d(&map_a, &full_bar[stage], da + c * kSlab, m0 + 64 * c, kb * kBlockK);
                    }
                    if (!g.b_mn) {
                        tma_load_2d(&map_b, &full_bar[stage], db, kb * kBlockK, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockN / 64; ++c)
                            tma_load_2d(&map_b, &full_bar[stage], db + c * kSlab, n0 + 64 * c, kb * kBlockK);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if constexpr (kCellTma) {
                    // the tile's fp32 addend (4 boxes of 32 columns) and cell state (its 32 units), issued after the
                    // operand loads so that waiting for the previous tile's epilogue does not starve the MMA
                    mbar_wait(add_empty, add_phase ^ 1);
                    mbar_expect_tx(add_full, (ep.addend ? kCellAddBytes : 0) + kCellCBytes);
                    if (ep.addend) {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            tma_load_2d(&map_o, add_full, smem_add + b * (kCellAddBytes / 4), n0 + 32 * b, m0);
                    }
                    tma_load_2d(&map_c, add_full, smem_c, n0 >> 2, m0);
                    add_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(kBlockM, kBlockN, g.a_mn, g.b_mn);
            // per-operand descriptor constants: leading byte offset and start-address step per 16 k
            const uint32_t a_lbo = g.a_mn ? kSlab : 16, b_lbo = g.b_mn ? kSlab : 16;
            const uint32_t a_step = g.a_mn ? (16 * 128) >> 4 : 32 >> 4, b_step = g.b_mn ? (16 * 128) >> 4 : 32 >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                const int split = g.split_major ? unit / num_tiles : unit % g.splits;
                const int kb0 = split * g.kb_per, kb1 = min(kb0 + g.kb_per, num_kb);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);             // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kBlockN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);                // TMA bytes landed
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * S::kStageA), a_lbo);
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * S::kStageB), b_lbo);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16(d_tmem, a_desc + a_step * k, b_desc + b_step * k, idesc, (kb > kb0) || k != 0);
                    umma_commit(&empty_bar[stage]);                    // frees the smem slot when MMAs finish
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);                          // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int e = warp - 2;
        const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = e >> 2;                                       // which half of the tile's columns
        constexpr int kCols = kBlockN / 2;
        const bool atomic = ep.atomic != 0 || g.splits > 1;
        int acc = 0;
        uint32_t acc_phase = 0;
        [[maybe_unused]] uint32_t add_phase = 0;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            const int tile = g.split_major ? unit % num_tiles : unit / g.splits;
            const int tile_n = tile % tiles_n;
            const int m0 = (tile / tiles_n) * kBlockM, n0 = tile_n * kBlockN;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kBlockN + half * kCols;
            if constexpr (kCellTma) {
                epilogue_cell_tma<kCols>(ep, taddr, lane, m0 + quarter * 32, n0 + half * kCols, half * kCols, quarter * 32 + lane,
                                         M, N, &tmem_full[acc], acc_phase, add_full, add_phase, smem_add, smem_c);
                __syncwarp();
                if (lane == 0) mbar_arrive(add_empty);                 // this warp no longer reads the staged tile
                add_phase ^= 1;
            } else
            epilogue_region<kCols, kEpi>(ep, taddr, lane, m0 + quarter * 32, n0 + half * kCols, M, N,
                                         tile_n * 2 + half, tiles_n * 2, atomic, &tmem_full[acc], acc_phase,
                                         smem + S::kOutOff + e * kOutStage, &map_o, g.tma_out);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (g.tma_out && lane == 0) tma_store_wait_all();            // bulk stores issued by this lane are complete
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 2-D bf16 row-major [outer, inner] with leading dimension ld (elements); box = [box_outer, 64].
// K-major operand: inner = K, outer = operand rows, box_outer = tile rows.
// MN-major operand: inner = operand rows, outer = K, box_outer = 64 (one slab).
// Encoding a map costs a driver call (~microseconds); the decoder re-uses a handful of
// (pointer, shape) combinations every step, so encoded maps are memoised.
struct TmapKey {
    const void *ptr; long long outer, inner, ld; int box_outer, box_inner, esize;
    bool operator<(const TmapKey &o) const {
        if (ptr != o.ptr) return ptr < o.ptr;
        if (outer != o.outer) return outer < o.outer;
        if (inner != o.inner) return inner < o.inner;
        if (ld != o.ld) return ld < o.ld;
        if (box_outer != o.box_outer) return box_outer < o.box_outer;
        if (box_inner != o.box_inner) return box_inner < o.box_inner;
        return esize < o.esize;
    }
};

// esize 2 = bf16, 4 = fp32; box = [box_outer, box_inner] elements with box_inner * esize == 128 bytes (one swizzle row)
static int make_tmap(CUtensorMap *map, const void *ptr, long long outer, long long inner, long long ld, int box_outer,
                     int box_inner, int esize) {
    static std::map<TmapKey, CUtensorMap> cache;
    static std::mutex mu;
    const TmapKey key{ptr, outer, inner, ld, box_outer, box_inner, esize};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *map = it->second; return DC_OK; }
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(DC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    DC_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * esize) % 16 == 0, "TMA tensor must be 16-byte aligned with a row stride that is a multiple of 16 bytes");
    cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * esize};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(DC_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 4096) cache.clear();
        cache[key] = *map;
    }
    return DC_OK;
}

int make_tmap_bf16(CUtensorMap *map, const void *ptr, long long outer, long long inner, long long ld, int box_outer) {
    return make_tmap(map, ptr, outer, inner, ld, box_outer, kBlockK, 2);
}

template <int kBlockN, int kEpi>
static int launch_tc(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mo, const TcEpilogue &ep,
                     const TcGeom &g, cudaStream_t stream, const CUtensorMap *mc = nullptr) {
    using S = TcSmem<kBlockN, kEpi>;
    static std::atomic<unsigned long long> attr_set{0};
    auto kern = gemm_bf16_tc_kernel<kBlockN, kEpi>;
    DC_CHECK_CUDA(once_per_device(attr_set, [&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes); }));
    const int units = ceil_div(g.M, kBlockM) * ceil_div(g.N, kBlockN) * g.splits;
    const int grid = units < sm_count() ? units : sm_count();
    DC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), (size_t)(g.tma_out ? S::kBytes : S::kBaseBytes), stream, ma, mb, mo,
                             mc ? *mc : ma, ep, g));
    return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant (tcgen05 cta_group::2): a CTA PAIR (cluster of 2 = the two SMs of a TPC) works on one
// 256 x 256 tile.  Each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256
// columns) per k-block -- 32 KB instead of 48 KB per stage, so the shared-memory read traffic per MMA
// and per SM is 2/3 and the TMA fill traffic 2/3 of the single-CTA kernel, which is what bounds that
// kernel's main loop.  The leader CTA (cluster rank 0) issues one
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) that reads both CTAs' shared memory and writes both
// CTAs' tensor memory; each CTA runs the epilogue of its own 128 x 256 accumulator half with the same
// epilogue code as the single-CTA kernel.
//
// Barriers: TMA loads of BOTH CTAs complete on the LEADER's full barrier (armed by the leader's producer
// with the bytes of both); tcgen05.commit multicasts to both CTAs' empty / tmem-full barriers; the
// epilogue warps of both CTAs arrive on the leader's tmem-empty barrier.  K-major or MN-major operands, split-K
// with reduce-add epilogues as in the single-CTA kernel.
// ------------------------------------------------------------------------------------------------

template <int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_o, const TcEpilogue ep, const TcGeom g) {
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
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int M = g.M, N = g.N, K = g.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int tiles_n = (N + kBlockN - 1) / kBlockN;
    const int tiles_m = (M + 2 * kBlockM - 1) / (2 * kBlockM);
    const int num_tiles = tiles_m * tiles_n;
    const int num_units = num_tiles * g.splits;          // unit = split * num_tiles + tile (split-K: CTAs running together share a K range)
    const int num_kb = (K + kBlockK - 1) / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (g.tma_out) tma_prefetch_desc(&map_o);
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc_2sm(tmem_ptr, kTmemCols);
    tc_fence_before();
    cluster_sync_all();                                                // both CTAs: barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = pair; unit < num_units; unit += num_pairs) {
                const int split = unit / num_tiles, tile = unit - split * num_tiles;
                const int m0 = (tile / tiles_n) * (2 * kBlockM) + (int)rank * kBlockM;
                const int nb0 = (tile % tiles_n) * kBlockN + (int)rank * 128;       // this CTA's half of the B tile
                const int kb0 = split * g.kb_per, kb1 = min(kb0 + g.kb_per, num_kb);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (S::kStageA + S::kStageB));   // bytes of both CTAs
                    const uint32_t bar = mapa_u32(&full_bar[stage], 0);
                    uint8_t *da = smem_a + stage * S::kStageA, *db = smem_b + stage * S::kStageB;
                    if (!g.a_mn) {
                        tma_load_2d_2sm(&map_a, bar, da, kb * kBlockK, m0);
                    } else {                                           // MN-major: two [64 k x 64 m] slabs
                        tma_load_2d_2sm(&map_a, bar, da, m0, kb * kBlockK);
                        tma_load_2d_2sm(&map_a, bar, da + kSlab, m0 + 64, kb * kBlockK);
                    }
                    if (!g.b_mn) {
                        tma_load_2d_2sm(&map_b, bar, db, kb * kBlockK, nb0);
                    } else {
                        tma_load_2d_2sm(&map_b, bar, db, nb0, kb * kBlockK);
                        tma_load_2d_2sm(&map_b, bar, db + kSlab, nb0 + 64, kb * kBlockK);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = make_idesc_bf16(2 * kBlockM, kBlockN, g.a_mn, g.b_mn);
            const uint32_t a_lbo = g.a_mn ? kSlab : 16, b_lbo = g.b_mn ? kSlab : 16;
            const uint32_t a_step = g.a_mn ? (16 * 128) >> 4 : 32 >> 4, b_step = g.b_mn ? (16 * 128) >> 4 : 32 >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int unit = pair; unit < num_units; unit += num_pairs) {
                const int split = unit / num_tiles;
                const int kb0 = split * g.kb_per, kb1 = min(kb0 + g.kb_per, num_kb);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);             // both CTAs' epilogues drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kBlockN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);                // both CTAs' TMA bytes landed
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * S::kStageA), a_lbo);
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * S::kStageB), b_lbo);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16_2sm(d_tmem, a_desc + a_step * k, b_desc + b_step * k, idesc, kb > kb0 || k != 0);
                    umma_commit_2sm(&empty_bar[stage]);                // frees the slot in both CTAs
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tmem_full[acc]);                      // accumulator complete, both CTAs
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9, both CTAs: own 128 x 256 half) =====================
        const int e = warp - 2;
        const int quarter = warp & 3;
        const int half = e >> 2;
        constexpr int kCols = kBlockN / 2;
        const bool atomic = ep.atomic != 0 || g.splits > 1;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = pair; unit < num_units; unit += num_pairs) {
            const int tile = unit % num_tiles;
            const int tile_n = tile % tiles_n;
            const int m0 = (tile / tiles_n) * (2 * kBlockM) + (int)rank * kBlockM, n0 = tile_n * kBlockN;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kBlockN + half * kCols;
            epilogue_region<kCols, kEpi>(ep, taddr, lane, m0 + quarter * 32, n0 + half * kCols, M, N,
                                         tile_n * 2 + half, tiles_n * 2, atomic, &tmem_full[acc], acc_phase,
                                         smem + S::kOutOff + e * kOutStage, &map_o, g.tma_out);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));    // the leader's barrier
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (g.tma_out && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();                                                // nobody touches the peer's smem / TMEM any more
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, kTmemCols);
    }
}

template <int kEpi>
static int launch_tc2(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mo, const TcEpilogue &ep,
                      const TcGeom &g, cudaStream_t stream) {
    using S = TcSmem2;
    static std::atomic<unsigned long long> attr_set{0};
    auto kern = gemm_bf16_tc2_kernel<kEpi>;
    DC_CHECK_CUDA(once_per_device(attr_set, [&] {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
        return e != cudaSuccess ? e : cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
    }));
    const int tiles = ceil_div(g.M, 2 * kBlockM) * ceil_div(g.N, 256) * g.splits;
    const int pairs = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = g.tma_out ? S::kBytes : S::kBaseBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    DC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mo, ep, g));
    return DC_OK;
}

int gemm_bf16_tc(const TcOperand &A, const TcOperand &B, const TcEpilogue &ep, int M, int N, int K, int epi,
                 cudaStream_t stream, int split_k) {
    if (M <= 0 || N <= 0) return DC_OK;
    DC_REQUIRE(K > 0 && A.ptr && B.ptr, "gemm_bf16_tc: bad arguments");
    const int sms = sm_count();
    const int num_kb = ceil_div(K, kBlockK);
    const int tiles_m = ceil_div(M, kBlockM);
    // tile width: 256 columns unless that leaves most SMs without a tile and 128 fills more of them
    bool wide = N > 128;
    if (wide && epi != kEpiArgmax && epi != kEpiArgmaxSum && epi != kEpiTopK) {
        const int t256 = tiles_m * ceil_div(N, 256), t128 = tiles_m * ceil_div(N, 128);
        if (t256 * 4 < sms * 3 && t128 > t256) wide = false;
    }
    const int tiles = tiles_m * ceil_div(N, wide ? 256 : 128);
    TcGeom g;
    g.M = M; g.N = N; g.K = K; g.a_mn = A.mn_major ? 1 : 0; g.b_mn = B.mn_major ? 1 : 0;
    g.splits = 1;
    g.tma_out = 0;
    static const int split_major_env = getenv("DCAP_SPLIT_MAJOR") ? atoi(getenv("DCAP_SPLIT_MAJOR")) : 1;
    g.split_major = split_major_env;
    // CTA-pair variant (cta_group::2, 256 x 256 tiles).  Measured (same box, A/B): greedy decoder 3.92 -> 3.70 ms
    // per 8000 RoIs, training step 9.15 -> 8.85 ms, beam search 229 -> 215 ms.
    // Launches of fewer than ~sms/8 pair tiles (the 512- / 1024-row step and data-gradient GEMMs of a training step
    // split over 8 / 4 ranks, the launch-per-GEMM decoder on a few hundred RoIs) run faster as 128 x 128 single-CTA
    // tiles: four times as many SMs get a tile and the launch is latency-, not throughput-bound.  Measured on one
    // GPU (gpurun_out/r2f_smallm*.log, forward+backward of the cfg3 step): 512 rows 1.96 -> 1.78 ms, 1024 rows
    // 2.70 -> 2.62 ms, 2048 / 4096 rows unchanged (a threshold of 33 and more is SLOWER there: 4.28 -> 4.38 ms).
    // (read per call so that the tests can hold both kernel families to the same small and ragged shapes)
    const char *two_cta_str = getenv("DCAP_2CTA");
    const int two_cta_env = two_cta_str ? atoi(two_cta_str) : -1;                            // 0 = off, n = minimum number of pair tiles
    // (accumulating weight-gradient GEMMs are split along K into ~2 waves of work units whatever their tile count: they
    // keep the pair kernel -- with the rule applied to them the 4096-row step measured 0.1 ms slower)
    const bool splittable = epi == kEpiStore && ep.atomic && ep.out_f32 && !ep.out_bf16;
    const int two_cta_min = two_cta_env >= 0 ? two_cta_env : (splittable ? 1 : (sms + 7) / 8);
    const int tiles2 = ceil_div(M, 256) * ceil_div(N, 256);
    const bool two_cta = two_cta_min != 0 && N >= 256 && tiles2 >= two_cta_min;
    if (epi == kEpiStore && ep.atomic && ep.out_f32 && !ep.out_bf16) {
        int want = split_k;
        if (want <= 0) {                      // fill ~2 waves of CTAs (CTA pairs), keep >= 8 k-blocks per unit
            const int slots = two_cta ? sms / 2 : sms, t = two_cta ? tiles2 : tiles;
            want = (2 * slots + t - 1) / t;
            const int cap = num_kb / 8 > 0 ? num_kb / 8 : 1;
            if (want > cap) want = cap;
        }
        if (want > num_kb) want = num_kb;
        if (want < 1) want = 1;
        g.splits = want;
    }
    g.kb_per = ceil_div(num_kb, g.splits);
    g.splits = ceil_div(num_kb, g.kb_per);        // no empty work unit
    CUtensorMap ma, mb;
    if (A.mn_major) { if (int rc = make_tmap_bf16(&ma, A.ptr, K, M, A.ld, 64)) return rc; }
    else            { if (int rc = make_tmap_bf16(&ma, A.ptr, M, K, A.ld, kBlockM)) return rc; }
    if (B.mn_major) { if (int rc = make_tmap_bf16(&mb, B.ptr, K, N, B.ld, 64)) return rc; }
    else            { if (int rc = make_tmap_bf16(&mb, B.ptr, N, K, B.ld, wide ? 256 : 128)) return rc; }
    CUtensorMap mb2 = mb;                         // K-major B of the pair kernel: each CTA loads 128 of the tile's 256 rows
    if (two_cta && !B.mn_major)
        if (int rc = make_tmap_bf16(&mb2, B.ptr, N, K, B.ld, 128)) return rc;
    if (epi == kEpiStore) {
        DC_REQUIRE(ep.out_f32 || ep.out_bf16, "gemm_bf16_tc: no output");
        DC_REQUIRE(!ep.out_f32 || (((uintptr_t)ep.out_f32 & 15) == 0 && ep.ld_f32 % 4 == 0), "fp32 output alignment");
        DC_REQUIRE(!ep.out_bf16 || (((uintptr_t)ep.out_bf16 & 7) == 0 && ep.ld_bf16 % 4 == 0), "bf16 output alignment");
        DC_REQUIRE(!ep.addend || (((uintptr_t)ep.addend & 15) == 0 && ep.ld_addend % 4 == 0), "addend alignment");
        DC_REQUIRE(!ep.mask_src || ((((uintptr_t)ep.mask_src & 15) == 0) && ep.ld_mask % 8 == 0 && N % 32 == 0),
                   "mask source must be 16-byte aligned, ld %% 8 == 0, N %% 32 == 0");
        DC_REQUIRE(ep.deint_units == 0 || (ep.out_f32 && !ep.out_bf16 && N == 4 * ep.deint_units && ep.deint_units % 8 == 0),
                   "de-interleaved store needs fp32 output with N == 4*units, units %% 8 == 0");
        // bulk tensor stores through shared memory for ONE plain output (fp32 preferred); a second output,
        // de-interleaved columns and unaligned leading dimensions keep the per-thread stores
        // Which outputs take the TMA path (bit mask, tunable for experiments): 1 = plain fp32, 2 = plain bf16,
        // 4 = outputs whose epilogue also reads a per-row addend / mask, 8 = accumulating (reduce-add) fp32.
        // Default 11: epilogues with per-row operand loads keep the smaller shared-memory footprint (the
        // staging buffers cost 32 KB of L1, which those loads need more than they gain from bulk stores).
        static const int tma_mask = getenv("DCAP_TMA_STORE_MASK") ? atoi(getenv("DCAP_TMA_STORE_MASK")) : 11;
        CUtensorMap mo = ma;
        const bool row_ops = ep.addend || ep.mask_src;
        const bool accum = ep.atomic || g.splits > 1;
        DC_REQUIRE(!ep.blocked32 || (ep.out_f32 && !ep.out_bf16 && !accum && N % 32 == 0 && ep.ld_f32 == N && ep.deint_units == 0),
                   "blocked-32 output needs a plain fp32 output with N %% 32 == 0, ld == N");
        // (with BOTH outputs requested the bulk-store variants would write only the fp32 one: they keep the per-thread stores)
        bool want = ep.deint_units == 0 && !ep.blocked32 && !(ep.out_f32 && ep.out_bf16) && (!row_ops || (tma_mask & 4)) && (!accum || (tma_mask & 8));
        if (want && ep.out_f32 && (accum || (tma_mask & 1))) {
            if (int rc = make_tmap(&mo, ep.out_f32, M, N, ep.ld_f32, 32, 32, 4)) return rc;
            g.tma_out = 1;
        } else if (want && !ep.out_f32 && (tma_mask & 2) && ep.ld_bf16 % 8 == 0 && ((uintptr_t)ep.out_bf16 & 15) == 0) {
            if (int rc = make_tmap(&mo, ep.out_bf16, M, N, ep.ld_bf16, 32, 64, 2)) return rc;
            g.tma_out = 2;
        }
        if (two_cta) {
            if (g.tma_out == 1) return launch_tc2<kEpiStoreTmaF32>(ma, mb2, mo, ep, g, stream);
            if (g.tma_out == 2) return launch_tc2<kEpiStoreTmaB16>(ma, mb2, mo, ep, g, stream);
            return launch_tc2<kEpiStore>(ma, mb2, mo, ep, g, stream);
        }
        if (g.tma_out == 1)
            return wide ? launch_tc<256, kEpiStoreTmaF32>(ma, mb, mo, ep, g, stream) : launch_tc<128, kEpiStoreTmaF32>(ma, mb, mo, ep, g, stream);
        if (g.tma_out == 2)
            return wide ? launch_tc<256, kEpiStoreTmaB16>(ma, mb, mo, ep, g, stream) : launch_tc<128, kEpiStoreTmaB16>(ma, mb, mo, ep, g, stream);
        return wide ? launch_tc<256, kEpiStore>(ma, mb, mo, ep, g, stream) : launch_tc<128, kEpiStore>(ma, mb, mo, ep, g, stream);
    }
    if (epi == kEpiCell) {
        DC_REQUIRE(ep.cell_c && ep.cell_units * 4 == N && N % 32 == 0, "cell epilogue: N must be 4*units, units %% 8 == 0");
        DC_REQUIRE(!ep.cell_tok || ep.cell_h_prev, "cell epilogue: masking needs the previous h");
        DC_REQUIRE(!ep.addend || (((uintptr_t)ep.addend & 15) == 0 && ep.ld_addend % 4 == 0), "addend alignment");
        DC_REQUIRE((!ep.cell_h_a || (ep.ld_h_a % 8 == 0 && ((uintptr_t)ep.cell_h_a & 15) == 0)) &&
                   (!ep.cell_h_b || (ep.ld_h_b % 8 == 0 && ((uintptr_t)ep.cell_h_b & 15) == 0)) &&
                   (!ep.cell_h_prev || (ep.ld_h_prev % 8 == 0 && ((uintptr_t)ep.cell_h_prev & 15) == 0)),
                   "cell epilogue: h buffers must be 16-byte aligned with ld %% 8 == 0");
        DC_REQUIRE(!ep.cell_gates_out || (((uintptr_t)ep.cell_gates_out & 15) == 0 && ep.ld_gates_out % 8 == 0),
                   "cell epilogue: gate buffer alignment");
        // shared-memory operand variant: 128-wide tiles, addend + cell state staged by TMA (see epilogue_cell_tma)
        static const bool cell_tma_off = getenv("DCAP_CELL_TMA") && atoi(getenv("DCAP_CELL_TMA")) == 0;
        // (measured: pays when there is an fp32 addend to stage -- 44 -> 39 us at 8000 x 2048 x 832; without one the
        // 128 x 256 tiles win because they read the B operand from shared memory half as often)
        const bool cell_tma = !cell_tma_off && ep.addend && !ep.addend_blocked32 && !g.a_mn && !g.b_mn && N % 128 == 0 && ep.addend_mod == 0 && ep.addend_div == 0 &&
                              ((uintptr_t)ep.cell_c & 15) == 0 && ep.cell_units % 4 == 0;
        if (two_cta && !cell_tma) return launch_tc2<kEpiCell>(ma, mb2, ma, ep, g, stream);
        if (cell_tma) {
            CUtensorMap mb128, madd = ma, mc;
            if (int rc = make_tmap_bf16(&mb128, B.ptr, N, K, B.ld, 128)) return rc;
            if (ep.addend)
                if (int rc = make_tmap(&madd, ep.addend, M, N, ep.ld_addend, 128, 32, 4)) return rc;
            if (int rc = make_tmap(&mc, ep.cell_c, M, ep.cell_units, ep.cell_units, 128, 32, 4)) return rc;
            return launch_tc<128, kEpiCellTma>(ma, mb128, madd, ep, g, stream, &mc);
        }
        return wide ? launch_tc<256, kEpiCell>(ma, mb, ma, ep, g, stream) : launch_tc<128, kEpiCell>(ma, mb, ma, ep, g, stream);
    }
    if (epi == kEpiTopK) {
        DC_REQUIRE(ep.partial && ep.bias && ep.topk >= 1 && ep.topk <= kTopKMax, "top-k epilogue needs bias, partial buffer, 1 <= k <= %d", kTopKMax);
        if (two_cta) return launch_tc2<kEpiTopK>(ma, mb2, ma, ep, g, stream);
        return wide ? launch_tc<256, kEpiTopK>(ma, mb, ma, ep, g, stream) : launch_tc<128, kEpiTopK>(ma, mb, ma, ep, g, stream);
    }
    DC_REQUIRE((epi == kEpiArgmax || epi == kEpiArgmaxSum) && ep.partial && ep.bias,
               "gemm_bf16_tc: arg-max epilogue needs bias and partial buffer");
    if (two_cta) return epi == kEpiArgmax ? launch_tc2<kEpiArgmax>(ma, mb2, ma, ep, g, stream)
                                          : launch_tc2<kEpiArgmaxSum>(ma, mb2, ma, ep, g, stream);
    if (epi == kEpiArgmax)
        return wide ? launch_tc<256, kEpiArgmax>(ma, mb, ma, ep, g, stream) : launch_tc<128, kEpiArgmax>(ma, mb, ma, ep, g, stream);
    return wide ? launch_tc<256, kEpiArgmaxSum>(ma, mb, ma, ep, g, stream)
                : launch_tc<128, kEpiArgmaxSum>(ma, mb, ma, ep, g, stream);
}

int gemm_tc_argmax_tiles(int N) { return 2 * ceil_div(N, N > 128 ? 256 : 128); }

// Merge the per-tile partials of the arg-max epilogue, one warp per row: token = first arg-max
// over tiles, optional softmax probability of that token, and (optionally) the embedding row of the
// token copied as bf16 into the next step's gate-GEMM operand (Embedding lookup of the greedy
// feedback, text_generation_model.py:147,222-225).
__global__ void __launch_bounds__(256) argmax_merge_kernel(const float4 *__restrict__ partial, int rows, int tiles,
                                                           int32_t *__restrict__ tok_out, int tok_stride,
                                                           int32_t *__restrict__ tok_cur, float *__restrict__ maxprob,
                                                           const uint4 *__restrict__ emb, int emb_ld8,
                                                           uint4 *__restrict__ x_out, long long ld_x8) {
    const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    if (r >= rows) return;
    float best = -INFINITY, sum = 0.f;
    int bi = 0x7fffffff;
    for (int t = lane; t < tiles; t += 32) {
        const float4 p = __ldg(partial + (long long)r * tiles + t);
        const int idx = __float_as_int(p.y);
        if (p.x > best) {
            sum = sum * __expf(best - p.x) + p.z;
            best = p.x; bi = idx;
        } else if (p.x == best) {
            sum += p.z; bi = min(bi, idx);
        } else {
            sum += p.z * __expf(p.x - best);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const float os = __shfl_xor_sync(0xffffffffu, sum, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best) {
            sum = sum * __expf(best - ob) + os;
            best = ob; bi = oi;
        } else if (ob == best) {
            sum += os; bi = min(bi, oi);                 // ties: the smaller column index wins
        } else {
            sum += os * __expf(ob - best);
        }
    }
    if (lane == 0) {
        if (tok_out) tok_out[(long long)r * tok_stride] = bi;
        if (tok_cur) tok_cur[r] = bi;
        if (maxprob) maxprob[r] = 1.0f / sum;
    }
    if (emb) {
        const uint4 *src = emb + (long long)bi * emb_ld8;
        uint4 *dst = x_out + (long long)r * ld_x8;
        for (int j = lane; j < emb_ld8; j += 32) dst[j] = __ldg(src + j);
    }
}

int argmax_merge(const float *partial, int rows, int tiles, int32_t *tok_out, int tok_stride, int32_t *tok_cur,
                 float *maxprob, cudaStream_t s, const __nv_bfloat16 *emb, int emb_ld, __nv_bfloat16 *x_out,
                 long long ld_x) {
    if (rows <= 0) return DC_OK;
    DC_REQUIRE(!emb || (emb_ld % 8 == 0 && ld_x % 8 == 0 && x_out), "argmax_merge: embedding rows must be 16-byte multiples");
    DC_CHECK_CUDA(launch_pdl(argmax_merge_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, s,
                             reinterpret_cast<const float4 *>(partial), rows, tiles, tok_out, tok_stride, tok_cur, maxprob,
                             reinterpret_cast<const uint4 *>(emb), emb_ld / 8, reinterpret_cast<uint4 *>(x_out),
                             (long long)(ld_x / 8)));
    return DC_OK;
}

// Merge of the kEpiTopK partials, one warp per row.  Each lane first reduces its slots (lane, lane+32, ...)
// to a sorted list of k, then k rounds of a warp-wide lexicographic arg-max pop the global order.
__global__ void __launch_bounds__(256) topk_merge_kernel(const float *__restrict__ partial, int rows, int slots, int k,
                                                         int32_t *__restrict__ idx_out, float *__restrict__ p_out) {
    const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int stride = 2 + 2 * k;
    float tv[kTopKMax];
    int ti[kTopKMax];
#pragma unroll
    for (int q = 0; q < kTopKMax; ++q) { tv[q] = -INFINITY; ti[q] = -1; }
    float best = -INFINITY, sum = 0.f;
    for (int t = lane; t < slots; t += 32) {
        const float *p = partial + ((long long)r * slots + t) * stride;
        const float mx = p[0], sm = p[1];
        if (mx > best) { sum = sum * __expf(best - mx) + sm; best = mx; }
        else if (mx > -INFINITY) sum += sm * __expf(mx - best);
        for (int c = 0; c < k; ++c) {
            float cv = p[2 + 2 * c];
            int ci = __float_as_int(p[3 + 2 * c]);
            if (ci < 0) break;
#pragma unroll
            for (int q = 0; q < kTopKMax; ++q) {
                if (cv > tv[q] || (cv == tv[q] && ci > ti[q])) {
                    const float t1 = tv[q]; const int t2 = ti[q];
                    tv[q] = cv; ti[q] = ci; cv = t1; ci = t2;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
        const float nm = fmaxf(best, ob);
        sum = (best > -INFINITY ? sum * __expf(best - nm) : 0.f) + (ob > -INFINITY ? os * __expf(ob - nm) : 0.f);
        best = nm;
    }
    float keep_v = 0.f;
    int keep_i = 0;
    for (int round = 0; round < k; ++round) {
        float hv = tv[0];
        int hi = ti[0], hl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, hv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, hi, o), ol = __shfl_xor_sync(0xffffffffu, hl, o);
            if (ov > hv || (ov == hv && oi > hi)) { hv = ov; hi = oi; hl = ol; }
        }
        if (lane == hl) {                                   // pop the winner's head
#pragma unroll
            for (int q = 0; q + 1 < kTopKMax; ++q) { tv[q] = tv[q + 1]; ti[q] = ti[q + 1]; }
            tv[kTopKMax - 1] = -INFINITY; ti[kTopKMax - 1] = -1;
        }
        if (lane == k - 1 - round) { keep_v = hv; keep_i = hi; }      // ascending output order
    }
    if (lane < k) {
        idx_out[(long long)r * k + lane] = keep_i;
        p_out[(long long)r * k + lane] = __expf(keep_v - best) / sum;
    }
}

int topk_merge(const float *partial, int rows, int slots, int k, int32_t *idx_out, float *p_out, cudaStream_t s) {
    if (rows <= 0) return DC_OK;
    DC_REQUIRE(k >= 1 && k <= kTopKMax, "beam width %d outside [1,%d]", k, kTopKMax);
    topk_merge_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(partial, rows, slots, k, idx_out, p_out);
    DC_CHECK_LAUNCH();
    return DC_OK;
}

}  // namespace dcap

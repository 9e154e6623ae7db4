// jg_scan.cuh -- single-pass decoupled-lookback exclusive scans (int64 result).
//
// Replaces numpy.cumsum at jagged.py:46 (counts2offsets) and every later use of it (new offsets after a mask
// :577, pair/cross counts :1075,:1095,:1164,:1311, argmax counts).
//
// Two kernels share the look-back protocol of jg_common.cuh (one 64-bit descriptor per tile: 2-bit state +
// 62-bit value, so flag and value travel in one atomic store):
//   * scan_win_kernel<T, XF, ITEMS>   -- round 2, every plain scan: windowed count / fill CTAs on published aggregates
//     (jg_lag.cuh), no tile waits.
//   * scan_tma_kernel<T, XF, ITEMS>   -- round 1's look-back, kept for the combinatorial counts.  The tile (128 threads
//     x 24 items = 3072 items, 24.6 KB of int64) is PARKED IN SHARED MEMORY by one TMA bulk copy, 9 CTAs resident per
//     SM: on B200 every tile of a look-back scan waits ~6 us for the slowest of its predecessors
//     (profiles/r1_lookback_study.md) and only parked bytes -- not registers -- can cover that wait.  XF turns the
//     staged data into items: the items themselves (counts) or combinatorial counts formed from staged offsets.
//   * scan_lookback_kernel<In>        -- functor input held in registers (validity flags, tiny tile-total scans);
//     tiles take their index from an atomic ticket.
// Both use 32-bit in-warp shuffle scans when every item of the tile is < 2^19 (the common case), else 64-bit.
//
// Algorithmic traffic: itemsize*n read + 8*(n+1) written (16n bytes for int64 counts).
#pragma once

#include <stdlib.h>

#include "jg_common.cuh"
#include "jg_lag.cuh"
#include "jg_tile.cuh"

namespace jg {

constexpr int kScanThreads = 256;
constexpr int kScanRows = 8;
constexpr int kScanTile = kScanThreads * kScanRows * 2;   // 4096 items

// ---- input functors: load2(i, n, a, b) yields items i and i+1 (0 beyond n) ---------------------------------
template <typename T>
struct ScanInArray {
    const T *p;
    bool aligned;   // p is aligned to 2*sizeof(T)
    __device__ __forceinline__ void load2(int64_t i, int64_t n, int64_t &a, int64_t &b) const {
        a = 0;
        b = 0;
        if (i + 1 < n) {
            if (aligned) {
                if constexpr (sizeof(T) == 8) {
                    longlong2 v = __ldcs(reinterpret_cast<const longlong2 *>(p + i));
                    a = v.x;
                    b = v.y;
                } else if constexpr (sizeof(T) == 4) {
                    int2 v = __ldcs(reinterpret_cast<const int2 *>(p + i));
                    a = static_cast<T>(v.x);
                    b = static_cast<T>(v.y);
                } else {
                    a = static_cast<int64_t>(p[i]);
                    b = static_cast<int64_t>(p[i + 1]);
                }
            } else {
                a = static_cast<int64_t>(p[i]);
                b = static_cast<int64_t>(p[i + 1]);
            }
        } else if (i < n) {
            a = static_cast<int64_t>(p[i]);
        }
    }
};

// combinatorics: item = f(count_a[, count_b])
__device__ __forceinline__ int64_t comb_count(int kind, int k, int64_t na, int64_t nb) {
    switch (kind) {
        case JG_COMB_PAIRS: return (na * (na + 1)) >> 1;
        case JG_COMB_DISTINCTS: return (na * (na - 1)) / 2;
        case JG_COMB_CROSS: return na * nb;
        default: {   // JG_COMB_CHOOSE, k = 2..5 (jagged.py:1152-1159)
            if (k == 2) return na * (na - 1) / 2;
            if (k == 3) return na * (na - 1) * (na - 2) / 6;
            if (k == 4) return na * (na - 1) * (na - 2) * (na - 3) / 24;
            return na * (na - 1) * (na - 2) * (na - 3) * (na - 4) / 120;
        }
    }
}

struct ScanInComb {
    const int64_t *oa;
    const int64_t *ob;
    int kind, k;
    __device__ __forceinline__ int64_t one(int64_t i) const {
        int64_t na = oa[i + 1] - oa[i];
        int64_t nb = ob ? (ob[i + 1] - ob[i]) : 0;
        return comb_count(kind, k, na, nb);
    }
    __device__ __forceinline__ void load2(int64_t i, int64_t n, int64_t &a, int64_t &b) const {
        a = (i < n) ? one(i) : 0;
        b = (i + 1 < n) ? one(i + 1) : 0;
    }
};

template <typename T>
__device__ __forceinline__ void scan_rows(const int64_t (&v)[kScanRows][2], T (&excl)[kScanRows], T &warp_total,
                                          int lane) {
    T incl[kScanRows];
#pragma unroll
    for (int r = 0; r < kScanRows; ++r) incl[r] = static_cast<T>(v[r][0]) + static_cast<T>(v[r][1]);
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
#pragma unroll
        for (int r = 0; r < kScanRows; ++r) {
            T o = __shfl_up_sync(0xffffffffu, incl[r], d);
            if (lane >= d) incl[r] += o;
        }
    }
    T run = 0;
#pragma unroll
    for (int r = 0; r < kScanRows; ++r) {
        T rowtot = __shfl_sync(0xffffffffu, incl[r], 31);
        excl[r] = run + incl[r] - (static_cast<T>(v[r][0]) + static_cast<T>(v[r][1]));
        run += rowtot;
    }
    warp_total = run;
}

template <class In>
__global__ void __launch_bounds__(kScanThreads)
scan_lookback_kernel(In in, int64_t n, int64_t *__restrict__ out, int64_t *__restrict__ total,
                     int32_t *__restrict__ status, LookbackWs *__restrict__ ws, uint32_t ntiles, bool out_aligned) {
    __shared__ uint32_t s_tile;
    __shared__ int64_t s_warp[kScanThreads / kWarp + 1];
    __shared__ int64_t s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&ws->ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t warp_item0 = static_cast<int64_t>(tile) * kScanTile + static_cast<int64_t>(warp) * (kScanRows * 64);

    int64_t v[kScanRows][2];
#pragma unroll
    for (int r = 0; r < kScanRows; ++r) in.load2(warp_item0 + r * 64 + lane * 2, n, v[r][0], v[r][1]);

    bool neg = false, big = false;
#pragma unroll
    for (int r = 0; r < kScanRows; ++r) {
        neg |= (v[r][0] < 0) | (v[r][1] < 0);
        big |= (static_cast<uint64_t>(v[r][0]) >= (1ull << 19)) | (static_cast<uint64_t>(v[r][1]) >= (1ull << 19));
    }
    if (neg && status) atomicOr(status, JG_ST_NEGATIVE_COUNT);
    const bool wide = __syncthreads_or(big);

    int64_t excl[kScanRows];
    int64_t warp_total;
    if (!wide) {
        int32_t e32[kScanRows];
        int32_t wt;
        scan_rows<int32_t>(v, e32, wt, lane);
#pragma unroll
        for (int r = 0; r < kScanRows; ++r) excl[r] = e32[r];
        warp_total = wt;
    } else {
        scan_rows<int64_t>(v, excl, warp_total, lane);
    }
    if (lane == 0) s_warp[warp] = warp_total;
    __syncthreads();

    if (warp == 0) {
        constexpr int NW = kScanThreads / kWarp;
        int64_t w = lane < NW ? s_warp[lane] : 0;
        int64_t wi = warp_incl_scan_add(w, lane);
        int64_t block_total = __shfl_sync(0xffffffffu, wi, NW - 1);
        int64_t tile_base = lookback_exclusive_warp(ws->desc, tile, block_total);
        if (lane < NW) s_warp[lane] = wi - w;
        if (lane == 0) {
            s_base = tile_base;
            if (tile == ntiles - 1) {
                out[n] = tile_base + block_total;
                if (total) *total = tile_base + block_total;
            }
        }
    }
    __syncthreads();
    const int64_t base = s_base + s_warp[warp];

#pragma unroll
    for (int r = 0; r < kScanRows; ++r) {
        const int64_t i = warp_item0 + r * 64 + lane * 2;
        const int64_t e0 = base + excl[r];
        const int64_t e1 = e0 + v[r][0];
        if (i + 1 < n) {
            if (out_aligned) {
                __stcs(reinterpret_cast<longlong2 *>(out + i), make_longlong2(e0, e1));
            } else {
                out[i] = e0;
                out[i + 1] = e1;
            }
        } else if (i < n) {
            out[i] = e0;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Array-input variant with the tile PARKED IN SHARED MEMORY: the tile (128 x 24 = 3072 items) is fetched with one TMA
// bulk copy (UBLKCP), the 128 threads keep only a handful of registers, so 9 CTAs are resident per SM.  That
// residency is what a look-back scan needs on B200: every tile waits ~6 us for the slowest of its predecessors
// (profiles/r1_lookback_study.md), and only parked tiles -- not registers -- can cover that wait.
// Measured 4.0 TB/s (vs 2.9 TB/s for the register-resident kernel above at 3 CTAs/SM).  The tile index is
// blockIdx.x: a tile only waits for tiles with smaller block indices, which a 1-D grid dispatches first.
// ------------------------------------------------------------------------------------------------------------
constexpr int kScanTmaThreads = 128;
constexpr int kScanTmaItems = 24;                                   // per thread (24.6 KB tile -> 9 CTAs/SM)
constexpr int kScanTmaTile = kScanTmaThreads * kScanTmaItems;      // 3072 items

// what a staged tile means: the items themselves, or offsets whose differences feed a combinatorial count
template <bool COUNT_NONEMPTY>
struct XfItems {
    static constexpr int EXTRA = 0;      // staged entries beyond the tile's items
    static constexpr int ARRAYS = 1;
    static constexpr bool NONEMPTY = COUNT_NONEMPTY;
    static constexpr bool SCATTER = false;
    int kind, k;
    int64_t *tile_nonempty = nullptr;    // by-product (COUNT_NONEMPTY): non-zero items per 512 consecutive items (event tile)
    int64_t n_event_tiles = 0;
    template <typename T>
    __device__ __forceinline__ int64_t item(const T *a, const T *, int p) const { return static_cast<int64_t>(a[p]); }
};
using XfIdentity = XfItems<false>;
using XfIdentityTiles = XfItems<true>;
// item = (best[i] >= 0): the 0 / 1 counts of a dense argmin / argmax result (jg_compact_nonneg).  SCATTER: the pass
// that writes the offsets also drops every non-negative best[i] at its slot -- the staged tile is still on chip, so the
// separate scatter pass (16N bytes read again) disappears.
struct XfNonNeg {
    static constexpr int EXTRA = 0;
    static constexpr int ARRAYS = 1;
    static constexpr bool NONEMPTY = false;
    static constexpr bool SCATTER = true;
    int kind, k;
    int64_t *out_index;
    template <typename T>
    __device__ __forceinline__ int64_t item(const T *a, const T *, int p) const { return a[p] >= 0 ? 1 : 0; }
    template <typename T>
    __device__ __forceinline__ void scatter(const T *a, int p, int64_t slot) const { out_index[slot] = a[p]; }
};
template <int NARR>
struct XfComb {
    static constexpr int EXTRA = 1;
    static constexpr int ARRAYS = NARR;
    static constexpr bool NONEMPTY = false;
    static constexpr bool SCATTER = false;
    int kind, k;
    template <typename T>
    __device__ __forceinline__ int64_t item(const T *a, const T *b, int p) const {
        const int64_t na = static_cast<int64_t>(a[p + 1]) - static_cast<int64_t>(a[p]);
        const int64_t nb = NARR == 2 ? static_cast<int64_t>(b[p + 1]) - static_cast<int64_t>(b[p]) : 0;
        return comb_count(kind, k, na, nb);
    }
};

template <typename T, typename S, class XF, int ITEMS, bool FULL>     // S = in-tile accumulator (int32 / int64); FULL: no bounds tests
__device__ __forceinline__ void scan_tma_rows(const XF &xf, const T *sa, const T *sb, int valid, int wbase, int lane,
                                              int64_t run, int64_t *__restrict__ out, int64_t i0, int64_t n,
                                              bool out_aligned) {
    constexpr int ROWS = ITEMS / 2;
    S carry = 0;
#pragma unroll 4
    for (int r = 0; r < ROWS; ++r) {
        const int p = wbase + r * 64 + lane * 2;
        const S x = (FULL || p < valid) ? static_cast<S>(xf.item(sa, sb, p)) : S(0);
        const S y = (FULL || p + 1 < valid) ? static_cast<S>(xf.item(sa, sb, p + 1)) : S(0);
        const S s2 = x + y;
        S incl = s2;
#pragma unroll
        for (int d = 1; d < kWarp; d <<= 1) {
            const S o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const int64_t e0 = run + static_cast<int64_t>(carry + incl - s2);
        const int64_t i = i0 + p;
        if (FULL || i + 1 < n) {
            if (out_aligned) __stcs(reinterpret_cast<longlong2 *>(out + i), make_longlong2(e0, e0 + static_cast<int64_t>(x)));
            else {
                out[i] = e0;
                out[i + 1] = e0 + static_cast<int64_t>(x);
            }
        } else if (i < n) {
            out[i] = e0;
        }
        if constexpr (XF::SCATTER) {
            if ((FULL || p < valid) && x != S(0)) xf.scatter(sa, p, e0);
            if ((FULL || p + 1 < valid) && y != S(0)) xf.scatter(sa, p + 1, e0 + static_cast<int64_t>(x));
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

template <typename T, class XF, int ITEMS>
__global__ void __launch_bounds__(kScanTmaThreads)
scan_tma_kernel(XF xf, const T *__restrict__ in_a, const T *__restrict__ in_b, int64_t n, int64_t *__restrict__ out,
                int64_t *__restrict__ total, int32_t *__restrict__ status, LookbackWs *__restrict__ ws, uint32_t ntiles,
                bool out_aligned) {
    constexpr int NW = kScanTmaThreads / kWarp;
    constexpr int TILE = kScanTmaThreads * ITEMS;
    __shared__ __align__(16) unsigned char s_raw_a[(TILE + XF::EXTRA) * sizeof(T) + 32];
    __shared__ __align__(16) unsigned char s_raw_b[XF::ARRAYS == 2 ? (TILE + XF::EXTRA) * sizeof(T) + 32 : 16];
    __shared__ uint64_t s_bar[2];
    __shared__ int64_t s_warp[NW + 1];
    __shared__ int64_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const int64_t i0 = static_cast<int64_t>(tile) * TILE;
    const int64_t left = n - i0;
    const int valid = left < TILE ? static_cast<int>(left < 0 ? 0 : left) : TILE;

    constexpr int kSubs = XF::NONEMPTY ? TILE / 512 : 1;
    __shared__ int s_ne[kSubs];
    if (XF::NONEMPTY && tid < kSubs) s_ne[tid] = 0;          // (published by the barrier inside stage_init)
    stage_init(s_bar, 2);
    const T *sa = reinterpret_cast<const T *>(s_raw_a);
    const T *sb = reinterpret_cast<const T *>(s_raw_b);
    if (valid > 0) {
        const uint32_t bytes = static_cast<uint32_t>((valid + XF::EXTRA) * sizeof(T));
        const uint32_t sha = stage_issue<sizeof(T)>(s_raw_a, reinterpret_cast<const unsigned char *>(in_a + i0), bytes, &s_bar[0]);
        sa = reinterpret_cast<const T *>(s_raw_a + sha);
        if (XF::ARRAYS == 2) {
            const uint32_t shb = stage_issue<sizeof(T)>(s_raw_b, reinterpret_cast<const unsigned char *>(in_b + i0), bytes, &s_bar[1]);
            sb = reinterpret_cast<const T *>(s_raw_b + shb);
            mbar_wait(&s_bar[1], 0);
        }
        stage_wait(&s_bar[0], 0);
    }
    // phase A: warp sums + validity flags (+ non-empty events per 512-event tile: what a dense argmin / argmax needs
    // to place its results, jg_segargminmax_dense -- a by-product of a pass that reads every count anyway)
    constexpr int SUBS = TILE / 512;
    static_assert(!XF::NONEMPTY || (TILE % 512 == 0 && (ITEMS * kWarp) % 64 == 0), "event tiles must tile the scan tile");
    constexpr bool count_ne = XF::NONEMPTY;
    const int wbase = warp * (ITEMS * kWarp);
    int64_t sum = 0;
    bool neg = false, big = false;
    const bool full = valid == TILE;
#pragma unroll
    for (int r = 0; r < ITEMS / 2; ++r) {
        const int p = wbase + r * 64 + lane * 2;
        const int64_t x = (full || p < valid) ? xf.item(sa, sb, p) : 0;
        const int64_t y = (full || p + 1 < valid) ? xf.item(sa, sb, p + 1) : 0;
        neg |= (x < 0) | (y < 0);
        big |= (static_cast<uint64_t>(x) >= (1ull << 19)) | (static_cast<uint64_t>(y) >= (1ull << 19));
        sum += x + y;
    }
    if (neg && status) atomicOr(status, JG_ST_NEGATIVE_COUNT);
    sum = warp_reduce_add(sum);
    if (lane == 0) s_warp[warp] = sum;
    const bool wide = __syncthreads_or(big);
    if (warp == 0) {
        int64_t w = lane < NW ? s_warp[lane] : 0;
        int64_t wi = warp_incl_scan_add(w, lane);
        const int64_t block_total = __shfl_sync(0xffffffffu, wi, NW - 1);
        const int64_t tile_base = lookback_exclusive_warp<1>(ws->desc, tile, block_total);
        if (lane < NW) s_warp[lane] = wi - w;
        if (lane == 0) {
            s_base = tile_base;
            if (tile == ntiles - 1) {
                out[n] = tile_base + block_total;
                if (total) *total = tile_base + block_total;
            }
        }
    }
    if constexpr (XF::NONEMPTY) {
        // OFF the critical path (the tile's aggregate is already published): warps 1.. count while warp 0 looks
        // back, warp 0 counts afterwards.  A warp's 64-item rows fall into at most two event tiles.
        int ne_lo = 0, ne_hi = 0;
        const int sub_lo = wbase >> 9;
#pragma unroll 4
        for (int r = 0; r < ITEMS / 2; ++r) {
            const int p = wbase + r * 64 + lane * 2;
            const int c = ((full || p < valid) ? (xf.item(sa, sb, p) > 0) : 0) + ((full || p + 1 < valid) ? (xf.item(sa, sb, p + 1) > 0) : 0);
            if (((wbase + r * 64) >> 9) == sub_lo) ne_lo += c;
            else ne_hi += c;
        }
        ne_lo = warp_reduce_add(ne_lo);
        ne_hi = warp_reduce_add(ne_hi);
        if (lane == 0) {
            if (ne_lo) atomicAdd(&s_ne[sub_lo], ne_lo);
            if (ne_hi) atomicAdd(&s_ne[sub_lo + 1], ne_hi);
        }
    }
    __syncthreads();
    if constexpr (XF::NONEMPTY) {
        if (count_ne && tid < SUBS) {
            const int64_t et = static_cast<int64_t>(tile) * SUBS + tid;
            if (et < xf.n_event_tiles) xf.tile_nonempty[et] = s_ne[tid];
        }
    }
    const int64_t run = s_base + s_warp[warp];
    if (valid == TILE) {
        if (wide) scan_tma_rows<T, int64_t, XF, ITEMS, true>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
        else scan_tma_rows<T, int32_t, XF, ITEMS, true>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
    } else {
        if (wide) scan_tma_rows<T, int64_t, XF, ITEMS, false>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
        else scan_tma_rows<T, int32_t, XF, ITEMS, false>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
    }
}

// ------------------------------------------------------------------------------------------------------------
// scan_tma_kernel without the look-back wait, on the windowed count / fill schedule of jg_lag.cuh.  A COUNT CTA sums
// one tile straight from global memory (coalesced loads that leave the tile in L2) and publishes the aggregate; a
// FILL CTA, a window of tickets later, stages the same tile with one TMA bulk copy (an L2 hit), takes its base from
// the published aggregates -- one L2 round trip, overlapped with the copy -- and writes the offsets.
// ------------------------------------------------------------------------------------------------------------
constexpr int kScanWindow = 512;         // tiles per window: 12 MB of int64 counts between a tile's two reads

template <typename T, class XF, int ITEMS>
__global__ void __launch_bounds__(kScanTmaThreads)
scan_win_kernel(XF xf, const T *__restrict__ in_a, const T *__restrict__ in_b, int64_t n, int64_t *__restrict__ out,
                int64_t *__restrict__ total, int32_t *__restrict__ status, LagDesc lag, const LagSchedule sched,
                bool out_aligned) {
    constexpr int NW = kScanTmaThreads / kWarp;
    constexpr int TILE = kScanTmaThreads * ITEMS;
    __shared__ __align__(16) unsigned char s_raw_a[(TILE + XF::EXTRA) * sizeof(T) + 32];
    __shared__ __align__(16) unsigned char s_raw_b[XF::ARRAYS == 2 ? (TILE + XF::EXTRA) * sizeof(T) + 32 : 16];
    __shared__ uint64_t s_bar[2];
    __shared__ int64_t s_warp[NW + 1];
    __shared__ int64_t s_base;
    __shared__ uint32_t s_ticket;
    constexpr int kSubs = XF::NONEMPTY ? TILE / 512 : 1;
    __shared__ int s_ne[kSubs];
    constexpr int SUBS = TILE / 512;
    static_assert(!XF::NONEMPTY || (TILE % 512 == 0 && (ITEMS * kWarp) % 64 == 0), "event tiles must tile the scan tile");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ntiles = sched.ntiles;
    uint32_t tile, tile_end;
    const int role = lag_role(lag_take_ticket(lag, &s_ticket), sched, tile, tile_end);
    if (role < 0) return;

    const int64_t i0 = static_cast<int64_t>(tile) * TILE;
    const int64_t left = n - i0;
    const int valid = left < TILE ? static_cast<int>(left < 0 ? 0 : left) : TILE;
    if (role == 0) {
        // ---------------- count: the tile's aggregate straight from global memory (the whole CTA: a lone warp has
        // too few loads in flight for 24 KB, measured) ----------------------------------------------------------
        const T *ga = in_a + i0;
        const T *gb = XF::ARRAYS == 2 ? in_b + i0 : nullptr;
        int64_t sum = 0;
        if (valid == TILE) {
#pragma unroll 8
            for (int j = 0; j < ITEMS; ++j) sum += xf.item(ga, gb, j * kScanTmaThreads + tid);
        } else {
#pragma unroll 1
            for (int p = tid; p < valid; p += kScanTmaThreads) sum += xf.item(ga, gb, p);
        }
        sum = warp_reduce_add(sum);
        if (lane == 0) s_warp[warp] = sum;
        __syncthreads();
        if (tid == 0) {
            int64_t t = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += s_warp[w];
            lag_publish(lag, tile, t);
        }
        return;
    }

    // ---------------- fill ------------------------------------------------------------------------------------
    if (XF::NONEMPTY && tid < kSubs) s_ne[tid] = 0;
    stage_init(s_bar, 2);
    const T *sa = reinterpret_cast<const T *>(s_raw_a);
    const T *sb = reinterpret_cast<const T *>(s_raw_b);
    if (valid > 0) {
        const uint32_t bytes = static_cast<uint32_t>((valid + XF::EXTRA) * sizeof(T));
        const uint32_t sha = stage_issue<sizeof(T)>(s_raw_a, reinterpret_cast<const unsigned char *>(in_a + i0), bytes, &s_bar[0]);
        sa = reinterpret_cast<const T *>(s_raw_a + sha);
        if (XF::ARRAYS == 2) {
            const uint32_t shb = stage_issue<sizeof(T)>(s_raw_b, reinterpret_cast<const unsigned char *>(in_b + i0), bytes, &s_bar[1]);
            sb = reinterpret_cast<const T *>(s_raw_b + shb);
        }
    }
    if (warp == 0) {          // the base: aggregates published a window ago (one L2 round trip while the copy lands)
        const int64_t tile_base = lag_prefix(lag, tile);
        if (lane == 0) s_base = tile_base;
    }
    if (valid > 0) {
        if (XF::ARRAYS == 2) mbar_wait(&s_bar[1], 0);
        stage_wait(&s_bar[0], 0);
    } else {
        __syncthreads();
    }
    const int wbase = warp * (ITEMS * kWarp);
    const bool full = valid == TILE;
    int64_t sum = 0;
    bool neg = false, big = false;
#pragma unroll
    for (int r = 0; r < ITEMS / 2; ++r) {
        const int p = wbase + r * 64 + lane * 2;
        const int64_t x = (full || p < valid) ? xf.item(sa, sb, p) : 0;
        const int64_t y = (full || p + 1 < valid) ? xf.item(sa, sb, p + 1) : 0;
        neg |= (x < 0) | (y < 0);
        big |= (static_cast<uint64_t>(x) >= (1ull << 19)) | (static_cast<uint64_t>(y) >= (1ull << 19));
        sum += x + y;
    }
    if (neg && status) atomicOr(status, JG_ST_NEGATIVE_COUNT);
    sum = warp_reduce_add(sum);
    if (lane == 0) s_warp[warp] = sum;
    if constexpr (XF::NONEMPTY) {
        int ne_lo = 0, ne_hi = 0;
        const int sub_lo = wbase >> 9;
#pragma unroll 4
        for (int r = 0; r < ITEMS / 2; ++r) {
            const int p = wbase + r * 64 + lane * 2;
            const int c = ((full || p < valid) ? (xf.item(sa, sb, p) > 0) : 0) + ((full || p + 1 < valid) ? (xf.item(sa, sb, p + 1) > 0) : 0);
            if (((wbase + r * 64) >> 9) == sub_lo) ne_lo += c;
            else ne_hi += c;
        }
        ne_lo = warp_reduce_add(ne_lo);
        ne_hi = warp_reduce_add(ne_hi);
        if (lane == 0) {
            if (ne_lo) atomicAdd(&s_ne[sub_lo], ne_lo);
            if (ne_hi) atomicAdd(&s_ne[sub_lo + 1], ne_hi);
        }
    }
    const bool wide = __syncthreads_or(big);
    if constexpr (XF::NONEMPTY) {
        if (tid < SUBS) {
            const int64_t et = static_cast<int64_t>(tile) * SUBS + tid;
            if (et < xf.n_event_tiles) xf.tile_nonempty[et] = s_ne[tid];
        }
    }
    int64_t run = s_base, tile_sum = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int64_t part = s_warp[w];
        if (w < warp) run += part;
        tile_sum += part;
    }
    if (tile == ntiles - 1 && tid == 0) {
        out[n] = s_base + tile_sum;
        if (total) *total = s_base + tile_sum;
    }
    if (full) {
        if (wide) scan_tma_rows<T, int64_t, XF, ITEMS, true>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
        else scan_tma_rows<T, int32_t, XF, ITEMS, true>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
    } else {
        if (wide) scan_tma_rows<T, int64_t, XF, ITEMS, false>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
        else scan_tma_rows<T, int32_t, XF, ITEMS, false>(xf, sa, sb, valid, wbase, lane, run, out, i0, n, out_aligned);
    }
}

template <typename T, class XF, int ITEMS>
int launch_scan_win(const XF &xf, const T *in_a, const T *in_b, int64_t n, int64_t *out, int64_t *total,
                    int32_t *status, void *ws, size_t ws_bytes, cudaStream_t stream) {
    constexpr int TILE = kScanTmaThreads * ITEMS;
    const int64_t ntiles = n <= 0 ? 1 : ceil_div(n, TILE);
    const size_t need = lag_ws_bytes(ntiles);
    JG_REQUIRE(ws != nullptr && ws_bytes >= need, "scan: workspace too small (%zu < %zu)", ws_bytes, need);
    JG_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "scan: workspace must be 16-byte aligned");
    JG_REQUIRE(ntiles < (1ll << 30), "scan: too many tiles");
    cudaError_t e = cudaMemsetAsync(ws, 0, need, stream);
    if (e != cudaSuccess) {
        set_error("scan: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return -1;
    }
    static bool carveout_set = false;
    if (!carveout_set) {
        cudaFuncSetAttribute(scan_win_kernel<T, XF, ITEMS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        carveout_set = true;
    }
    const LagSchedule sched = lag_schedule(ntiles, kScanWindow, 1);
    const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    LagDesc lag = lag_carve(ws, ntiles);
    // CTA order: blockIdx.x by default -- a fill CTA then only waits for count CTAs with smaller block indices, which a
    // 1-D grid dispatches first (the same assumption CUB-style single-pass scans made before tickets, and round 1's
    // scan_tma_kernel).  JG_TICKET=1 deals the tiles by an execution-order atomic instead: no assumption at all, for
    // +0.7 us (one atomic round trip) at the head of every CTA, 0.427 -> 0.446 ms on 125 M counts.
    static int ticket_env = -1;
    if (ticket_env < 0) {
        const char *c = getenv("JG_TICKET");
        ticket_env = c ? atoi(c) : 0;
    }
    if (ticket_env == 0) lag.ticket = nullptr;
    scan_win_kernel<T, XF, ITEMS><<<static_cast<unsigned>(lag_grid(sched)), kScanTmaThreads, 0, stream>>>(
        xf, in_a, in_b, n, out, total, status, lag, sched, out_aligned);
    return check_launch("scan_win_kernel");
}

template <typename T, class XF, int ITEMS>
int launch_scan_tma(const XF &xf, const T *in_a, const T *in_b, int64_t n, int64_t *out, int64_t *total,
                    int32_t *status, void *ws, size_t ws_bytes, cudaStream_t stream) {
    // plain scans: the windowed count / fill kernel; combinatorial counts (two offsets per item, a division-free but
    // longer transform) keep the look-back -- their count CTAs cost more than the wait they avoid (measured)
    if (XF::EXTRA == 0)
        return launch_scan_win<T, XF, ITEMS>(xf, in_a, in_b, n, out, total, status, ws, ws_bytes, stream);
    constexpr int TILE = kScanTmaThreads * ITEMS;
    const int64_t ntiles = n <= 0 ? 1 : ceil_div(n, TILE);
    const size_t need = sizeof(LookbackWs) + static_cast<size_t>(ntiles) * 8;
    JG_REQUIRE(ws != nullptr && ws_bytes >= need, "scan: workspace too small (%zu < %zu)", ws_bytes, need);
    JG_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "scan: workspace must be 16-byte aligned");
    JG_REQUIRE(ntiles < (1ll << 31), "scan: too many tiles");
    cudaError_t e = cudaMemsetAsync(ws, 0, need, stream);
    if (e != cudaSuccess) {
        set_error("scan: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return -1;
    }
    static bool carveout_set = false;
    if (!carveout_set) {
        cudaFuncSetAttribute(scan_tma_kernel<T, XF, ITEMS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        carveout_set = true;
    }
    const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    scan_tma_kernel<T, XF, ITEMS><<<static_cast<unsigned>(ntiles), kScanTmaThreads, 0, stream>>>(
        xf, in_a, in_b, n, out, total, status, reinterpret_cast<LookbackWs *>(ws), static_cast<uint32_t>(ntiles),
        out_aligned);
    return check_launch("scan_tma_kernel");
}

template <typename T>
int launch_scan_array(const T *in, int64_t n, int64_t *out, int64_t *total, int32_t *status, void *ws,
                      size_t ws_bytes, cudaStream_t stream, int64_t *tile_nonempty = nullptr) {
    if (tile_nonempty == nullptr)
        return launch_scan_tma<T, XfIdentity, kScanTmaItems>(XfIdentity{0, 0}, in, static_cast<const T *>(nullptr), n, out,
                                                            total, status, ws, ws_bytes, stream);
    XfIdentityTiles xf{0, 0};
    xf.tile_nonempty = tile_nonempty;
    xf.n_event_tiles = n <= 0 ? 0 : (n + 511) / 512;
    return launch_scan_tma<T, XfIdentityTiles, kScanTmaItems>(xf, in, static_cast<const T *>(nullptr), n, out, total,
                                                             status, ws, ws_bytes, stream);
}

// Host launcher.  `out` has n+1 entries: out[i] = sum of items before i, out[n] = grand total.
template <class In>
int launch_scan(const In &in, int64_t n, int64_t *out, int64_t *total, int32_t *status, void *ws, size_t ws_bytes,
                cudaStream_t stream) {
    const int64_t ntiles = n <= 0 ? 1 : ceil_div(n, kScanTile);
    const size_t need = sizeof(LookbackWs) + static_cast<size_t>(ntiles) * 8;
    JG_REQUIRE(ws != nullptr && ws_bytes >= need, "scan: workspace too small (%zu < %zu)", ws_bytes, need);
    JG_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "scan: workspace must be 16-byte aligned");
    JG_REQUIRE(ntiles < (1ll << 31), "scan: too many tiles");
    cudaError_t e = cudaMemsetAsync(ws, 0, need, stream);
    if (e != cudaSuccess) {
        set_error("scan: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return -1;
    }
    const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    scan_lookback_kernel<In><<<static_cast<unsigned>(ntiles), kScanThreads, 0, stream>>>(
        in, n, out, total, status, reinterpret_cast<LookbackWs *>(ws), static_cast<uint32_t>(ntiles), out_aligned);
    return check_launch("scan_lookback_kernel");
}

}  // namespace jg

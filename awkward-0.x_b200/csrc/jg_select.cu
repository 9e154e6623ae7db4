// jg_select.cu -- jagged __getitem__: boolean-mask compaction (jagged.py:562-589), integer gather
// (jagged.py:541-560), plain take (jagged.py:1294) and the dense form of argmin/argmax results.
#include <stdlib.h>

#include "jg_scan.cuh"
#include <cstring>
#include <type_traits>

#include "jg_compact.cuh"
#include "jg_tile.cuh"

namespace jg {

// ------------------------------------------------------------------------------------------------------------
// a[jagged_bool]   (jagged.py:562-589)
//
// Three launches, none of which waits on another CTA.  Single-pass forms were measured in both rounds and lost: the
// decoupled look-back (every tile waits ~6 us for the slowest of its predecessors, profiles/r1_lookback_study.md), a
// persistent kernel that counts LAG tiles ahead and fills from L2, and a one-launch "count CTAs a window ahead of fill
// CTAs" schedule (profiles/r2_lag_study.md: 2.0 - 2.7 ms against 1.62 ms, and a straggling count stalls a whole
// window).  The column is therefore read twice (the second read is what keeps a[a > c] at 62 % of its ALGORITHMIC
// bytes while the two kernels run at 86 % of the HBM peak on the bytes they actually move):
//   1. mask_count_kernel : one WARP per event tile popcounts the tile's mask run (M bytes read) -> tile totals
//   2. the look-back scan of jg_scan.cuh over the N/512 tile totals -> tile bases + grand total K (tiny)
//   3. mask_fill_kernel  : the tile's content run and mask run are staged with TMA bulk copies (UBLKCP); the
//      compaction is ELEMENT-parallel (a warp walks 32 consecutive positions: ballot + popc give each selected
//      element its slot, so the global stores of a warp are consecutive); the per-event new offsets are read
//      back from the on-chip per-position prefix.
// Algorithmic traffic: v*M + M + 8(N+1) read, v*K + 8(N+1) written (the mask is read twice: +M actual).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int popcount_bytes16(int4 v) {
    // bytes are 0/1 booleans (anything non-zero counts)
    return __popc(__vcmpne4(v.x, 0) & 0x01010101u) + __popc(__vcmpne4(v.y, 0) & 0x01010101u) +
           __popc(__vcmpne4(v.z, 0) & 0x01010101u) + __popc(__vcmpne4(v.w, 0) & 0x01010101u);
}

template <class SEL>
__device__ __forceinline__ int count16(const SEL &sel, int4 v) {
    using E = typename SEL::Elem;
    if constexpr (sizeof(E) == 1) {
        return popcount_bytes16(v);
    } else {
        union { int4 raw; E e[16 / sizeof(E)]; } q;
        q.raw = v;
        int c = 0;
#pragma unroll
        for (int k = 0; k < static_cast<int>(16 / sizeof(E)); ++k) c += sel.test(q.e[k]);
        return c;
    }
}

constexpr int kCountWarps = 8;     // event tiles per CTA of the counting kernel

constexpr int64_t kCountLongRun = 1 << 14;      // a run at least this long is counted by the whole CTA

// selected items of m[0, len) seen by threads first, first + step, ...: unaligned head (< 16 bytes), 16-byte body (U
// loads in flight per thread), tail
template <class SEL, int U>
__device__ __forceinline__ int64_t count_run(const SEL &sel, const typename SEL::Elem *m, int64_t len, int first, int step) {
    using E = typename SEL::Elem;
    constexpr int PER = 16 / sizeof(E);
    int64_t cnt = 0;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(m);
    int64_t head = static_cast<int64_t>(((16 - (addr & 15)) & 15) / sizeof(E));
    if (head > len) head = len;
    if (first < head) cnt += sel.test(m[first]);
    const int64_t nvec = (len - head) / PER;
    const int4 *mv = reinterpret_cast<const int4 *>(m + head);
#pragma unroll U
    for (int64_t i = first; i < nvec; i += step) cnt += count16(sel, __ldcs(mv + i));
    const int64_t tail0 = head + nvec * PER;
    if (tail0 + first < len) cnt += sel.test(m[tail0 + first]);
    return cnt;
}

// OVERSIZE: only the tiles whose run exceeds `km` positions are counted (the single-pass selection counts every other
// tile out of its own staged copy) and the totals are PUBLISHED (jg_lag.cuh) instead of stored.
// One warp per event tile; a tile with a LONG run (an event of tens of thousands of items among short ones) is left to
// the whole CTA afterwards: one warp keeps 1 KB in flight, and a kernel ends when its slowest warp does (70 000
// float32 items: ~280 us for a warp, ~20 us for the CTA).
template <class SEL, bool OVERSIZE = false>
__global__ void __launch_bounds__(kCountWarps * kWarp, 8)
mask_count_kernel(const SEL sel, const typename SEL::Elem *__restrict__ mask, int64_t mask_shift,
                  const int64_t *__restrict__ offsets, int64_t n, int64_t ntiles, int64_t *__restrict__ tile_totals,
                  const LagDesc lag = LagDesc{}, int64_t km = 0) {
    using E = typename SEL::Elem;
    __shared__ int s_nlong;
    __shared__ int s_long[kCountWarps];
    __shared__ int64_t s_part[kCountWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_nlong = 0;
    __syncthreads();
    const int64_t tile = static_cast<int64_t>(blockIdx.x) * kCountWarps + warp;
    if (tile < ntiles) {
        const int64_t ev0 = tile * kTileEvents;
        const int64_t ev1 = min(n, ev0 + kTileEvents);
        const int64_t e0 = __ldg(offsets + ev0), e1 = __ldg(offsets + ev1);
        const int64_t len = e1 - e0;
        if (!(OVERSIZE && len <= km)) {
            if (len >= kCountLongRun) {
                if (lane == 0) s_long[atomicAdd(&s_nlong, 1)] = warp;
            } else {
                int64_t cnt = count_run<SEL, 2>(sel, mask + e0 + mask_shift, len, lane, kWarp);
                cnt = warp_reduce_add(cnt);
                if (lane == 0) {
                    if (OVERSIZE) lag_publish(lag, static_cast<uint32_t>(tile), cnt);
                    else tile_totals[tile] = cnt;
                }
            }
        }
    }
    __syncthreads();
    const int nlong = s_nlong;
#pragma unroll 1
    for (int q = 0; q < nlong; ++q) {
        const int64_t t = static_cast<int64_t>(blockIdx.x) * kCountWarps + s_long[q];
        const int64_t ev0 = t * kTileEvents;
        const int64_t ev1 = min(n, ev0 + kTileEvents);
        const int64_t e0 = __ldg(offsets + ev0), e1 = __ldg(offsets + ev1);
        int64_t cnt = count_run<SEL, 4>(sel, mask + e0 + mask_shift, e1 - e0, threadIdx.x, kCountWarps * kWarp);
        cnt = warp_reduce_add(cnt);
        if (lane == 0) s_part[warp] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t total = 0;
#pragma unroll
            for (int w = 0; w < kCountWarps; ++w) total += s_part[w];
            if (OVERSIZE) lag_publish(lag, static_cast<uint32_t>(t), total);
            else tile_totals[t] = total;
        }
        __syncthreads();
    }
}

// SAME: the selector looks at the content column itself (a[a > c]): one staged copy serves both
// SINGLE: ONE pass over the data -- the tile's total comes out of the counting pass over its own staged copy, is
// published (jg_lag.cuh: hierarchical aggregates, no chain of prefixes) and the tile's base is the sum of what its
// predecessors published; a tile waits only until the tiles before it (dispatched earlier, so resident or done) have
// counted theirs.  `tile_bases` is not read; the grand total goes to *total_out.  (Tiles too long to stage were
// counted and published by mask_count_kernel<OVERSIZE> before the launch: a CTA reading a long run alone would hold
// up every tile behind it.)
template <typename V, int KM, class SEL, bool SAME, bool SINGLE = false>     // KM = positions a tile may stage
__global__ void __launch_bounds__(kTileThreads, sizeof(V) == 8 ? 5 : (SAME ? 7 : 6))
mask_fill_kernel(const SEL sel, const V *__restrict__ content, const typename SEL::Elem *__restrict__ mask,
                 int64_t mask_shift, const int64_t *__restrict__ offsets, int64_t n, V *__restrict__ out,
                 int64_t *__restrict__ new_offsets, int64_t *__restrict__ new_counts,
                 const int64_t *__restrict__ tile_bases, uint32_t ntiles, const LagDesc lag = LagDesc{},
                 int64_t *__restrict__ total_out = nullptr) {
    using E = typename SEL::Elem;
    static_assert(!SAME || sizeof(E) == sizeof(V), "SAME needs the selector to read the content's own items");
    constexpr int NW = kTileThreads / kWarp;
    __shared__ int64_t s_off[kTileEvents + 1];
    // (+ one row of padding behind the selector's run: the rows of the compaction are whole)
    __shared__ __align__(16) unsigned char s_content[KM * sizeof(V) + 32 + (SAME ? 32 * sizeof(V) : 0)];
    __shared__ __align__(16) unsigned char s_mask[SAME ? 16 : KM * sizeof(E) + 32 + 32 * sizeof(E)];
    __shared__ uint16_t s_pfx[KM + 2];
    __shared__ uint64_t s_bar[2];
    __shared__ int32_t s_scan[NW + 1];
    __shared__ int64_t s_base;

    const uint32_t tile_index = blockIdx.x;
    EventTile tile;
    if (threadIdx.x == 0) {
        // start both TMA copies before the cooperative offsets load (see segreduce_kernel)
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        int64_t ev0;
        int nev;
        EventTile::range(n, tile_index, ev0, nev);
        const int64_t a = __ldg(offsets + ev0), b = __ldg(offsets + ev0 + nev);
        if (b > a && b - a <= KM) {
            if (!SAME)
                stage_bulk(s_mask, reinterpret_cast<const unsigned char *>(mask + a + mask_shift),
                           static_cast<uint32_t>((b - a) * sizeof(E)), &s_bar[1]);
            stage_bulk(s_content, reinterpret_cast<const unsigned char *>(content + a),
                       static_cast<uint32_t>((b - a) * sizeof(V)), &s_bar[0]);
        }
    }
    tile.load(s_off, offsets, n, tile_index);
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const int64_t len = e1 - e0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t base = 0;
    if constexpr (!SINGLE) {
        base = __ldg(tile_bases + tile_index);
        if (tile_index == ntiles - 1 && threadIdx.x == 0) new_offsets[n] = __ldg(tile_bases + ntiles);
    }

    if (len <= KM) {
        // ---------------- staged tile (rows of 32 positions, jg_compact.cuh) ------------------------------------
        const int L = static_cast<int>(len);
        const int rows_total = (L + 31) >> 5;
        const V *cv = nullptr;
        const E *mv = nullptr;
        if (L > 0) {
            const uint32_t cs = stage_edges<sizeof(V)>(s_content, reinterpret_cast<const unsigned char *>(content + e0),
                                                       static_cast<uint32_t>(L * sizeof(V)));
            cv = reinterpret_cast<const V *>(s_content + cs);
            E *pad;
            if (SAME) {
                pad = reinterpret_cast<E *>(s_content + cs);
            } else {
                const uint32_t ms = stage_edges<sizeof(E)>(s_mask, reinterpret_cast<const unsigned char *>(mask + e0 + mask_shift),
                                                           static_cast<uint32_t>(L * sizeof(E)));
                pad = reinterpret_cast<E *>(s_mask + ms);
            }
            mv = pad;
            if (warp == 1) rows_pad<SEL>(pad, L, lane);
            mbar_wait(&s_bar[SAME ? 0 : 1], 0);
            __syncthreads();
        }
        // warp w owns rows [w * rows_per_warp, (w + 1) * rows_per_warp)
        const int rows_per_warp = (rows_total + NW - 1) / NW;
        const int r_begin = min(rows_total, warp * rows_per_warp);
        const int r_end = min(rows_total, r_begin + rows_per_warp);
        const int cnt = rows_count(sel, mv, r_begin, r_end, lane);
        if (lane == 0) s_scan[warp] = cnt;
        __syncthreads();
        // every warp adds up the warps before it itself (one REDUX): no scan by warp 0, no second barrier
        uint32_t run = __reduce_add_sync(0xffffffffu, lane < warp ? s_scan[lane] : 0);
        if constexpr (SINGLE) {
            if (warp == 0) {
                const int64_t tile_total = __reduce_add_sync(0xffffffffu, lane < NW ? s_scan[lane] : 0);
                if (lane == 0) lag_publish(lag, tile_index, tile_total);
                const int64_t b = lag_prefix(lag, tile_index);
                if (lane == 0) {
                    s_base = b;
                    if (tile_index == ntiles - 1) {
                        new_offsets[n] = b + tile_total;
                        *total_out = b + tile_total;
                    }
                }
            }
            __syncthreads();
            base = s_base;
        }
        if (!SAME && L > 0) mbar_wait(&s_bar[0], 0);      // content has landed (waited late: it is only needed now)
        run = rows_compact<V, SEL, SAME>(sel, mv, cv, s_pfx, r_begin, r_end, lane, run, out + base);
        // the rank one past the last row (the end of the last event when L is a multiple of 32; the last warp's chunk
        // always ends the tile, also when it is empty)
        if (warp == NW - 1 && lane == 0) s_pfx[rows_total << 5] = static_cast<uint16_t>(run);
        __syncthreads();
        // per-event new offsets -- and counts, which every later step asks for (jagged.py:383-388), for one more
        // shared-memory read instead of a 16N-byte pass of their own
#pragma unroll 1
        for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
            const int first = s_pfx[s_off[ev] - e0];
            new_offsets[tile.event0 + ev] = base + first;
            if (new_counts) new_counts[tile.event0 + ev] = static_cast<int>(s_pfx[s_off[ev + 1] - e0]) - first;
        }
        return;
    }

    // ---------------- oversize tile (long events): the run in chunks of KM positions ---------------------------
    // Every chunk is staged and compacted exactly like a staged tile (one bulk copy, the same two passes over whole
    // rows) with a running base; an event takes its new offset from the chunk that holds its first position.  (Round 1
    // read every chunk twice from global memory between four barriers; a kernel ends when its slowest CTA does, and
    // 1000 tiles with an event of 5 k - 70 k elements among 150 M short events cost a[a > 20] 0.35 ms of 2.05,
    // profiles/r2_exp32.log.)
    if constexpr (SINGLE) {       // (counted and published before the launch)
        if (warp == 0) {
            const int64_t b = lag_prefix(lag, tile_index);
            if (lane == 0) s_base = b;
        }
        __syncthreads();
        base = s_base;
    }
    const int evA = threadIdx.x, evB = threadIdx.x + kTileThreads;
    const int64_t stA = evA < tile.nev ? s_off[evA] - e0 : -1;
    const int64_t stB = evB < tile.nev ? s_off[evB] - e0 : -1;
    __syncthreads();              // every start is in its thread's registers: the offsets tile becomes the NEW offsets
    int64_t *s_new = s_off;
    int64_t run = 0;
    uint32_t phase = 0;           // (neither barrier has been used: thread 0 only starts a copy for a run that fits)
#pragma unroll 1
    for (int64_t cs = 0; cs < len; cs += KM) {
        const int L = static_cast<int>(min(static_cast<int64_t>(KM), len - cs));
        const V *gc = content + e0 + cs;
        const E *gm = mask + e0 + mask_shift + cs;
        if (threadIdx.x == 0) {
            fence_proxy_async_smem();          // (the previous chunk's reads of the stage, ended by the barrier above)
            if (!SAME) stage_bulk(s_mask, reinterpret_cast<const unsigned char *>(gm), static_cast<uint32_t>(L * sizeof(E)), &s_bar[1]);
            stage_bulk(s_content, reinterpret_cast<const unsigned char *>(gc), static_cast<uint32_t>(L * sizeof(V)), &s_bar[0]);
        }
        const uint32_t cshift = stage_edges<sizeof(V)>(s_content, reinterpret_cast<const unsigned char *>(gc), static_cast<uint32_t>(L * sizeof(V)));
        const V *cv = reinterpret_cast<const V *>(s_content + cshift);
        E *pad;
        if (SAME) {
            pad = reinterpret_cast<E *>(s_content + cshift);
        } else {
            const uint32_t ms = stage_edges<sizeof(E)>(s_mask, reinterpret_cast<const unsigned char *>(gm), static_cast<uint32_t>(L * sizeof(E)));
            pad = reinterpret_cast<E *>(s_mask + ms);
        }
        const E *mv = pad;
        if (warp == 1) rows_pad<SEL>(pad, L, lane);
        mbar_wait(&s_bar[0], phase);
        if (!SAME) mbar_wait(&s_bar[1], phase);
        phase ^= 1u;
        __syncthreads();
        const int rows_total = (L + 31) >> 5;
        const int rows_per_warp = (rows_total + NW - 1) / NW;
        const int r_begin = min(rows_total, warp * rows_per_warp);
        const int r_end = min(rows_total, r_begin + rows_per_warp);
        const int cnt = rows_count(sel, mv, r_begin, r_end, lane);
        if (lane == 0) s_scan[warp] = cnt;
        __syncthreads();
        uint32_t wrun = __reduce_add_sync(0xffffffffu, lane < warp ? s_scan[lane] : 0);
        wrun = rows_compact<V, SEL, SAME>(sel, mv, cv, s_pfx, r_begin, r_end, lane, wrun, out + base + run);
        if (warp == NW - 1 && lane == 0) s_pfx[rows_total << 5] = static_cast<uint16_t>(wrun);
        __syncthreads();
        if (stA >= cs && stA < cs + L) s_new[evA] = run + s_pfx[stA - cs];
        if (stB >= cs && stB < cs + L) s_new[evB] = run + s_pfx[stB - cs];
        run += s_pfx[L];
        __syncthreads();      // the stage, s_pfx and s_scan are rewritten by the next chunk
    }
    if (stA >= len) s_new[evA] = run;       // empty events at the end of the tile
    if (stB >= len) s_new[evB] = run;
    if (threadIdx.x == 0) {
        s_new[tile.nev] = run;
        if (SINGLE && tile_index == ntiles - 1) {
            new_offsets[n] = base + run;
            *total_out = base + run;
        }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ev = threadIdx.x + h * kTileThreads;
        if (ev < tile.nev) {
            new_offsets[tile.event0 + ev] = base + s_new[ev];
            if (new_counts) new_counts[tile.event0 + ev] = s_new[ev + 1] - s_new[ev];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The fifth single-pass form (phase = 4, JG_SELECT_RANGES=1 for the Python host): one CTA owns a RANGE of kRangeTiles
// consecutive event tiles.  Phase A: one warp per tile counts its run straight from global memory (a streaming read
// that leaves the range in L2), the range total is published (jg_lag.cuh) and the range's base is the sum of what the
// ranges before it published -- ONE wait per kRangeTiles tiles, for CTAs that were dispatched earlier and publish at
// the very start of their life.  Phase B: the tiles of the range one after the other, each staged (an L2 hit a few
// microseconds after phase A) and compacted like the chunks of an oversize tile, with the tile's base known on chip.
// The column is read from HBM once; a CTA lives ~8 tiles long and holds one tile's shared memory.
// ------------------------------------------------------------------------------------------------------------
constexpr int kRangeTiles = 8;

template <typename V, int KM, class SEL, bool SAME>
__global__ void __launch_bounds__(kTileThreads, sizeof(V) == 8 ? 5 : (SAME ? 7 : 6))
mask_range_kernel(const SEL sel, const V *__restrict__ content, const typename SEL::Elem *__restrict__ mask,
                  int64_t mask_shift, const int64_t *__restrict__ offsets, int64_t n, V *__restrict__ out,
                  int64_t *__restrict__ new_offsets, int64_t *__restrict__ new_counts, uint32_t ntiles,
                  const LagDesc lag, int64_t *__restrict__ total_out) {
    using E = typename SEL::Elem;
    static_assert(!SAME || sizeof(E) == sizeof(V), "SAME needs the selector to read the content's own items");
    static_assert(kRangeTiles <= kTileThreads / kWarp, "one warp per tile in the counting phase");
    constexpr int NW = kTileThreads / kWarp;
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) unsigned char s_content[KM * sizeof(V) + 32 + (SAME ? 32 * sizeof(V) : 0)];
    __shared__ __align__(16) unsigned char s_mask[SAME ? 16 : KM * sizeof(E) + 32 + 32 * sizeof(E)];
    __shared__ uint16_t s_pfx[KM + 2];
    __shared__ uint64_t s_bar[2];
    __shared__ int32_t s_scan[NW + 1];
    __shared__ int64_t s_bound[kRangeTiles + 1];
    __shared__ int64_t s_tilebase[kRangeTiles + 1];
    __shared__ int64_t s_base;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t range = blockIdx.x;
    const uint32_t t0 = range * kRangeTiles;
    const int nt = static_cast<int>(min(static_cast<uint32_t>(kRangeTiles), ntiles - t0));
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    if (static_cast<int>(threadIdx.x) <= nt) {
        const int64_t ev = min(n, static_cast<int64_t>(t0 + threadIdx.x) * kTileEvents);
        s_bound[threadIdx.x] = __ldg(offsets + ev);
    }
    __syncthreads();
    // ---- phase A: one warp per tile
    if (warp < nt) {
        const int64_t a = s_bound[warp], b = s_bound[warp + 1];
        int64_t cnt = count_run<SEL, 4>(sel, mask + a + mask_shift, b - a, lane, kWarp);
        cnt = warp_reduce_add(cnt);
        if (lane == 0) s_tilebase[warp + 1] = cnt;
    }
    __syncthreads();
    if (warp == 0) {
        int64_t total = 0;
        if (lane == 0) {
            s_tilebase[0] = 0;
            for (int j = 1; j <= nt; ++j) {
                total += s_tilebase[j];
                s_tilebase[j] = total;              // -> exclusive bases of the tiles, [nt] = the range's total
            }
            lag_publish(lag, range, total);
        }
        const int64_t rb = lag_prefix(lag, range);
        if (lane == 0) {
            s_base = rb;
            if (t0 + nt == ntiles) {
                new_offsets[n] = rb + total;
                *total_out = rb + total;
            }
        }
    }
    __syncthreads();
    const int64_t range_base = s_base;
    uint32_t phase = 0;
    // ---- phase B: the tiles one after the other, each compacted chunk by chunk (one chunk unless events are long)
#pragma unroll 1
    for (int j = 0; j < nt; ++j) {
        EventTile tile;
        tile.load(s_off, offsets, n, t0 + j);
        const int64_t base = range_base + s_tilebase[j];
        const int64_t e0 = tile.first_element(), len = tile.last_element() - e0;
        const int evA = threadIdx.x, evB = threadIdx.x + kTileThreads;
        const int64_t stA = evA < tile.nev ? s_off[evA] - e0 : -1;
        const int64_t stB = evB < tile.nev ? s_off[evB] - e0 : -1;
        __syncthreads();          // every start is in its thread's registers: the offsets tile becomes the NEW offsets
        int64_t *s_new = s_off;
        int64_t run = 0;
#pragma unroll 1
        for (int64_t cs = 0; cs < len; cs += KM) {
            const int L = static_cast<int>(min(static_cast<int64_t>(KM), len - cs));
            const V *gc = content + e0 + cs;
            const E *gm = mask + e0 + mask_shift + cs;
            if (threadIdx.x == 0) {
                fence_proxy_async_smem();
                if (!SAME) stage_bulk(s_mask, reinterpret_cast<const unsigned char *>(gm), static_cast<uint32_t>(L * sizeof(E)), &s_bar[1]);
                stage_bulk(s_content, reinterpret_cast<const unsigned char *>(gc), static_cast<uint32_t>(L * sizeof(V)), &s_bar[0]);
            }
            const uint32_t cshift = stage_edges<sizeof(V)>(s_content, reinterpret_cast<const unsigned char *>(gc), static_cast<uint32_t>(L * sizeof(V)));
            const V *cv = reinterpret_cast<const V *>(s_content + cshift);
            E *pad;
            if (SAME) {
                pad = reinterpret_cast<E *>(s_content + cshift);
            } else {
                const uint32_t ms = stage_edges<sizeof(E)>(s_mask, reinterpret_cast<const unsigned char *>(gm), static_cast<uint32_t>(L * sizeof(E)));
                pad = reinterpret_cast<E *>(s_mask + ms);
            }
            const E *mv = pad;
            if (warp == 1) rows_pad<SEL>(pad, L, lane);
            mbar_wait(&s_bar[0], phase);
            if (!SAME) mbar_wait(&s_bar[1], phase);
            phase ^= 1u;
            __syncthreads();
            const int rows_total = (L + 31) >> 5;
            const int rows_per_warp = (rows_total + NW - 1) / NW;
            const int r_begin = min(rows_total, warp * rows_per_warp);
            const int r_end = min(rows_total, r_begin + rows_per_warp);
            const int cnt = rows_count(sel, mv, r_begin, r_end, lane);
            if (lane == 0) s_scan[warp] = cnt;
            __syncthreads();
            uint32_t wrun = __reduce_add_sync(0xffffffffu, lane < warp ? s_scan[lane] : 0);
            wrun = rows_compact<V, SEL, SAME>(sel, mv, cv, s_pfx, r_begin, r_end, lane, wrun, out + base + run);
            if (warp == NW - 1 && lane == 0) s_pfx[rows_total << 5] = static_cast<uint16_t>(wrun);
            __syncthreads();
            if (stA >= cs && stA < cs + L) s_new[evA] = run + s_pfx[stA - cs];
            if (stB >= cs && stB < cs + L) s_new[evB] = run + s_pfx[stB - cs];
            run += s_pfx[L];
            __syncthreads();      // the stage, s_pfx and s_scan are rewritten by the next chunk
        }
        if (stA >= len) s_new[evA] = run;       // empty events at the end of the tile (all of them when it is empty)
        if (stB >= len) s_new[evB] = run;
        if (threadIdx.x == 0) s_new[tile.nev] = run;
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ev = threadIdx.x + h * kTileThreads;
            if (ev < tile.nev) {
                new_offsets[tile.event0 + ev] = base + s_new[ev];
                if (new_counts) new_counts[tile.event0 + ev] = s_new[ev + 1] - s_new[ev];
            }
        }
        __syncthreads();          // the offsets tile is reloaded for the next tile of the range
    }
}

// counts of two layouts must agree (jagged.py:911-912)
__global__ void __launch_bounds__(256)
same_counts_kernel(const int64_t *__restrict__ oa, const int64_t *__restrict__ ob, int64_t n,
                   int32_t *__restrict__ status) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    bool bad = false;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        bad |= (oa[i + 1] - oa[i]) != (ob[i + 1] - ob[i]);
    if (bad) atomicOr(status, JG_ST_COUNTS_MISMATCH);
}

// ------------------------------------------------------------------------------------------------------------
// a[jagged_int]   (jagged.py:541-560)
// ------------------------------------------------------------------------------------------------------------
template <typename V, typename I>
__global__ void __launch_bounds__(kTileThreads)
index_gather_kernel(const V *__restrict__ content, const int64_t *__restrict__ offsets, const I *__restrict__ index,
                    const int64_t *__restrict__ index_offsets, int64_t n, V *__restrict__ out,
                    int32_t *__restrict__ status) {
    constexpr int KM = 8192;         // (tiles of index runs up to a mean of ~14 per event fit one owner table)
    __shared__ int64_t s_ioff[kTileEvents + 1];
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) uint16_t s_own[KM + 128];
    EventTile itile, tile;
    // both offsets tiles in ONE memory round trip; the index run and the start of every event's content are asked
    // for at the same time, so the two dependent loads below (index, then content[start + index]) find them in L2
    itile.load_nosync<sizeof(I)>(s_ioff, index_offsets, n, blockIdx.x, index);
    tile.load_nosync<sizeof(V)>(s_off, offsets, n, blockIdx.x, content);
    __syncthreads();
    const int64_t io0 = __ldg(index_offsets);
    const int64_t k0 = itile.first_element(), k1 = itile.last_element();
    bool oob = false;
    // element-parallel over the index run of this tile, in chunks of one owner table (a run longer than the table --
    // index events of more than ~14 entries on average -- was a binary search per entry in round 1: a[full_index] cost
    // 1.5x as much per element at a mean of 10 as at 5): coalesced reads of index, coalesced writes of out;
    // 4 independent (index -> content) load chains in flight per thread
#pragma unroll 1
    for (int64_t cs = k0; cs < k1; cs += KM) {
        const int64_t ce = min(k1, cs + KM);
        if (k1 - k0 <= KM) build_owner_table(s_own, itile, static_cast<int>(k1 - k0));
        else build_owner_range(s_own, itile, cs - k0, static_cast<int>(ce - cs));
#pragma unroll 1
        for (int64_t kb = cs + threadIdx.x; kb < ce; kb += 4 * kTileThreads) {
            int64_t ix[4], src[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t k = kb + u * kTileThreads;
                if (k < ce) ix[u] = static_cast<int64_t>(__ldcs(index + k));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t k = kb + u * kTileThreads;
                src[u] = -1;
                if (k < ce) {
                    const int ev = static_cast<int>(s_own[k - cs]);
                    const int64_t start = s_off[ev], count = s_off[ev + 1] - start;
                    int64_t i = ix[u];
                    if (i < 0) i += count;                                  // jagged.py:552-553
                    if (i < 0 || i >= count) oob = true;                    // jagged.py:555-556
                    else src[u] = start + i;
                }
            }
            V val[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (src[u] >= 0) val[u] = __ldg(content + src[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t k = kb + u * kTileThreads;
                if (src[u] >= 0) __stcs(out + (k - io0), val[u]);
            }
        }
        if (k1 - k0 > KM) __syncthreads();               // the table is rewritten by the next chunk
    }
    if (oob) atomicOr(status, JG_ST_INDEX_OOB);
}

template <typename V>
__global__ void __launch_bounds__(256)
take_kernel(const V *__restrict__ content, const int64_t *__restrict__ index, int64_t n, V *__restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
        int64_t ix[4];
        V v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < n) ix[u] = __ldcs(index + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < n) v[u] = __ldg(content + ix[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < n) __stcs(out + i + u * stride, v[u]);
    }
}

// content[index[k]] = values[k] (or one scalar): the store side of __setitem__ (jagged.py:789-838) and of the
// scatter steps of parents2startsstops (jagged.py:73-93).  Negative indexes are skipped when `skip_negative`.
template <typename V>
__global__ void __launch_bounds__(256)
put_kernel(V *__restrict__ content, const int64_t *__restrict__ index, int64_t n, const V *__restrict__ values,
           V scalar, int skip_negative) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int64_t at = __ldcs(index + k);
        if (skip_negative && at < 0) continue;
        content[at] = values ? __ldcs(values + k) : scalar;
    }
}

static inline int flat_grid(int64_t work, int threads) {
    int64_t blocks = ceil_div(work, threads);
    int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

template <typename V>
static int launch_put(void *content, const int64_t *index, int64_t n, const void *values, uint64_t scalar_bits,
                      int skip_negative, cudaStream_t s) {
    V scalar;
    memcpy(&scalar, &scalar_bits, sizeof(V));
    put_kernel<V><<<flat_grid(n, 256), 256, 0, s>>>(static_cast<V *>(content), index, n,
                                                   static_cast<const V *>(values), scalar, skip_negative);
    return check_launch("put_kernel");
}

// out[pos[i]] = content[i] for mask[i] != 0   (flat boolean selection, jagged.py:599-605 applied to starts/stops)
template <typename V>
__global__ void __launch_bounds__(256)
flat_scatter_kernel(const V *__restrict__ content, const uint8_t *__restrict__ mask, const int64_t *__restrict__ pos,
                    int64_t n, V *__restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        if (mask[i]) out[__ldg(pos + i)] = content[i];
}


// workspace layout shared by the "tile totals -> tile bases" helpers: [totals: nt+1][bases: nt+1][scan ws]
struct TileScanWs {
    int64_t *totals;
    int64_t *bases;
    void *scan_ws;
    size_t scan_ws_bytes;
};

static int carve_tile_ws(void *ws, size_t ws_bytes, int64_t ntiles, TileScanWs &o, const char *who) {
    const size_t arr = (static_cast<size_t>(ntiles) + 2) * 8;
    const size_t arr_al = (arr + 255) & ~static_cast<size_t>(255);
    const size_t scan_need = sizeof(LookbackWs) + static_cast<size_t>(ceil_div(ntiles, kScanTile) + 1) * 8;
    if (ws == nullptr || ws_bytes < 2 * arr_al + scan_need + 256) {
        set_error("%s: workspace too small (%zu < %zu)", who, ws_bytes, 2 * arr_al + scan_need + 256);
        return -2;
    }
    char *p = static_cast<char *>(ws);
    o.totals = reinterpret_cast<int64_t *>(p);
    o.bases = reinterpret_cast<int64_t *>(p + arr_al);
    o.scan_ws = p + 2 * arr_al;
    o.scan_ws_bytes = ws_bytes - 2 * arr_al;
    return 0;
}

// phase 0: the whole selection; 1: count + scan only (the total K is then in *total, one read-back away, while the
// compaction has not been queued yet); 2: the compaction only (after a phase-1 call with the same arguments and
// workspace).  Splitting lets the host read K while the compaction runs: the next operation is queued behind it
// without the device ever going idle.
template <typename V, int KM, class SEL, bool SAME>
static int launch_mask_select(const SEL &sel, const void *content, const typename SEL::Elem *mask, int64_t mask_shift,
                              const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                              int64_t *total, void *ws, size_t ws_bytes, cudaStream_t s, int phase) {
    const int64_t ntiles = tile_grid(n);
    TileScanWs t;
    int rc = carve_tile_ws(ws, ws_bytes, ntiles, t, "jg_mask_select");
    if (rc) return rc;
    if (phase != 2 && phase != 3) {
        mask_count_kernel<SEL><<<static_cast<unsigned>(ceil_div(ntiles, kCountWarps)), kCountWarps * kWarp, 0, s>>>(
            sel, mask, mask_shift, offsets, n, ntiles, t.totals);
        rc = check_launch("mask_count_kernel");
        if (rc) return rc;
        ScanInArray<int64_t> in{t.totals, true};
        rc = launch_scan(in, ntiles, t.bases, total, nullptr, t.scan_ws, t.scan_ws_bytes, s);
        if (rc || phase == 1) return rc;
    }
    if (phase == 3) {
        // single pass: descriptors zeroed, the (usually absent) oversize tiles counted and published, then ONE sweep
        const size_t need = lag_ws_bytes(ntiles);
        if (need > 2 * ((static_cast<size_t>(ntiles) + 2) * 8 + 255)) {      // (never: 8.3 bytes per tile against 16)
            set_error("jg_mask_select: workspace too small for the single-pass descriptors");
            return -2;
        }
        if (cudaMemsetAsync(t.totals, 0, need, s) != cudaSuccess) {
            set_error("jg_mask_select: cudaMemsetAsync failed");
            return -3;
        }
        const LagDesc lag = lag_carve(t.totals, ntiles);
        mask_count_kernel<SEL, true><<<static_cast<unsigned>(ceil_div(ntiles, kCountWarps)), kCountWarps * kWarp, 0, s>>>(
            sel, mask, mask_shift, offsets, n, ntiles, nullptr, lag, KM);
        rc = check_launch("mask_count_kernel<oversize>");
        if (rc) return rc;
        mask_fill_kernel<V, KM, SEL, SAME, true><<<static_cast<unsigned>(ntiles), kTileThreads, 0, s>>>(
            sel, static_cast<const V *>(content), mask, mask_shift, offsets, n, static_cast<V *>(out), new_offsets,
            new_counts, nullptr, static_cast<uint32_t>(ntiles), lag, total);
        return check_launch("mask_fill_kernel<single>");
    }
    mask_fill_kernel<V, KM, SEL, SAME><<<static_cast<unsigned>(ntiles), kTileThreads, 0, s>>>(
        sel, static_cast<const V *>(content), mask, mask_shift, offsets, n, static_cast<V *>(out), new_offsets,
        new_counts, t.bases, static_cast<uint32_t>(ntiles));
    return check_launch("mask_fill_kernel");
}

// phase 4: ranges of kRangeTiles tiles per CTA (mask_range_kernel), ONE launch
template <typename V, int KM, class SEL, bool SAME>
static int launch_mask_ranges(const SEL &sel, const void *content, const typename SEL::Elem *mask, int64_t mask_shift,
                              const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                              int64_t *total, void *ws, size_t ws_bytes, cudaStream_t s) {
    const int64_t ntiles = tile_grid(n);
    const int64_t nranges = ceil_div(ntiles, kRangeTiles);
    const size_t need = lag_ws_bytes(nranges);
    JG_REQUIRE(ws != nullptr && ws_bytes >= need + 256, "jg_mask_select: workspace too small for the range descriptors");
    if (cudaMemsetAsync(ws, 0, need, s) != cudaSuccess) {
        set_error("jg_mask_select: cudaMemsetAsync failed");
        return -3;
    }
    const LagDesc lag = lag_carve(ws, nranges);
    mask_range_kernel<V, KM, SEL, SAME><<<static_cast<unsigned>(nranges), kTileThreads, 0, s>>>(
        sel, static_cast<const V *>(content), mask, mask_shift, offsets, n, static_cast<V *>(out), new_offsets,
        new_counts, static_cast<uint32_t>(ntiles), lag, total);
    return check_launch("mask_range_kernel");
}

// staged positions per tile: content + selector + 16-bit prefix must leave room for ~6 CTAs per SM
template <class SEL, bool SAME>
static int dispatch_mask_select(const SEL &sel, const void *content, int itemsize, const typename SEL::Elem *mask,
                                int64_t mask_shift, const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                                int64_t *new_counts, int64_t *total, void *ws, size_t ws_bytes, cudaStream_t s, int phase) {
    constexpr int E = sizeof(typename SEL::Elem);
    if constexpr (SAME) {
        if constexpr (E == 4) {
            if (phase == 4) return launch_mask_ranges<uint32_t, 4096, SEL, true>(sel, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s);
            return launch_mask_select<uint32_t, 4096, SEL, true>(sel, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
        }
        else return launch_mask_select<uint64_t, 3584, SEL, true>(sel, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
    } else {
        constexpr int KM8 = E == 1 ? 3584 : (E == 4 ? 3072 : 2304);   // (f64 by another f64 column: the host keeps the mask path)
        constexpr int KM4 = E == 8 ? 3072 : 4096;
        if constexpr (E == 1) {          // narrow contents only through a byte mask (keeps the number of kernels down)
            if (itemsize == 1) return launch_mask_select<uint8_t, KM4, SEL, false>(sel, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
            if (itemsize == 2) return launch_mask_select<uint16_t, KM4, SEL, false>(sel, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
        }
        switch (itemsize) {
            case 4: return launch_mask_select<uint32_t, KM4, SEL, false>(sel, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
            case 8: return launch_mask_select<uint64_t, KM8, SEL, false>(sel, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
        }
        set_error("jg_mask_select: unsupported itemsize %d", itemsize);
        return -2;
    }
}

template <typename K, bool STRICT>
static int compare_select(const void *content, int itemsize, const void *key, int op, double scalar, int64_t key_shift,
                          const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                          int64_t *total, void *ws, size_t ws_bytes, cudaStream_t s, int phase) {
    using SEL = FloatCompare<K, STRICT>;
    SEL sel;
    const K c = static_cast<K>(scalar);                 // the eager comparison converts the scalar the same way
    const bool flip = (op == JG_OP_LT || op == JG_OP_LE);
    sel.flip = flip ? (static_cast<typename SEL::Bits>(1) << (8 * sizeof(K) - 1)) : 0;
    sel.c = flip ? -c : c;
    const K *k = static_cast<const K *>(key);
    if (key == content && key_shift == 0 && itemsize == static_cast<int>(sizeof(K)))
        return dispatch_mask_select<SEL, true>(sel, content, itemsize, k, 0, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
    JG_REQUIRE(itemsize == 4 || itemsize == 8, "jg_compare_select: a column other than the compared one must have 4- or 8-byte items");
    return dispatch_mask_select<SEL, false>(sel, content, itemsize, k, key_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase);
}

template <typename V>
static int launch_index_gather(const void *content, const int64_t *offsets, const void *index, int index_dtype,
                               const int64_t *index_offsets, int64_t n, void *out, int32_t *status, cudaStream_t s) {
    if (index_dtype == JG_I64)
        index_gather_kernel<V, int64_t><<<tile_grid(n), kTileThreads, 0, s>>>(
            static_cast<const V *>(content), offsets, static_cast<const int64_t *>(index), index_offsets, n,
            static_cast<V *>(out), status);
    else if (index_dtype == JG_I32)
        index_gather_kernel<V, int32_t><<<tile_grid(n), kTileThreads, 0, s>>>(
            static_cast<const V *>(content), offsets, static_cast<const int32_t *>(index), index_offsets, n,
            static_cast<V *>(out), status);
    else {
        set_error("jg_index_gather: unsupported index dtype %d", index_dtype);
        return -2;
    }
    return check_launch("index_gather_kernel");
}

}  // namespace jg

using namespace jg;

extern "C" {

int jg_mask_select_phase(const void *content, int itemsize, const uint8_t *mask, int64_t mask_shift,
                         const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                         int64_t *total, void *ws, size_t ws_bytes, int phase, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && new_offsets, "jg_mask_select: bad arguments");
    JG_REQUIRE(phase >= 0 && phase <= 3, "jg_mask_select: phase must be 0 .. 3");
    return dispatch_mask_select<ByteMask, false>(ByteMask{}, content, itemsize, mask, mask_shift, offsets, n, out,
                                                 new_offsets, new_counts, total, ws, ws_bytes, as_stream(stream), phase);
}

int jg_mask_select(const void *content, int itemsize, const uint8_t *mask, int64_t mask_shift,
                   const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                   int64_t *total, void *ws, size_t ws_bytes, void *stream) {
    return jg_mask_select_phase(content, itemsize, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, ws,
                                ws_bytes, 0, stream);
}

int jg_compare_select_phase(const void *content, int itemsize, const void *key, int key_dtype, int op, double scalar,
                            int64_t key_shift, const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                            int64_t *new_counts, int64_t *total, void *ws, size_t ws_bytes, int phase, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && new_offsets && key, "jg_compare_select: bad arguments");
    JG_REQUIRE(op == JG_OP_GT || op == JG_OP_GE || op == JG_OP_LT || op == JG_OP_LE,
               "jg_compare_select: operator %d is not an ordering comparison", op);
    JG_REQUIRE(phase >= 0 && phase <= 4, "jg_compare_select: phase must be 0 .. 4");
    cudaStream_t s = as_stream(stream);
    const bool strict = (op == JG_OP_GT || op == JG_OP_LT);
#define JG_CS(K, ST) compare_select<K, ST>(content, itemsize, key, op, scalar, key_shift, offsets, n, out, new_offsets, new_counts, total, ws, ws_bytes, s, phase)
    if (key_dtype == JG_F32) return strict ? JG_CS(float, true) : JG_CS(float, false);
    if (key_dtype == JG_F64) return strict ? JG_CS(double, true) : JG_CS(double, false);
#undef JG_CS
    set_error("jg_compare_select: key dtype %d is not float32 / float64", key_dtype);
    return -2;
}

int jg_compare_select(const void *content, int itemsize, const void *key, int key_dtype, int op, double scalar,
                      int64_t key_shift, const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                      int64_t *new_counts, int64_t *total, void *ws, size_t ws_bytes, void *stream) {
    return jg_compare_select_phase(content, itemsize, key, key_dtype, op, scalar, key_shift, offsets, n, out, new_offsets,
                                   new_counts, total, ws, ws_bytes, 0, stream);
}

int jg_put(void *content, int itemsize, const int64_t *index, int64_t n, const void *values, uint64_t scalar_bits,
           int skip_negative, void *stream) {
    if (n <= 0) return 0;
    JG_REQUIRE(content && index, "jg_put: bad arguments");
    cudaStream_t s = as_stream(stream);
    switch (itemsize) {
        case 1: return launch_put<uint8_t>(content, index, n, values, scalar_bits, skip_negative, s);
        case 2: return launch_put<uint16_t>(content, index, n, values, scalar_bits, skip_negative, s);
        case 4: return launch_put<uint32_t>(content, index, n, values, scalar_bits, skip_negative, s);
        case 8: return launch_put<uint64_t>(content, index, n, values, scalar_bits, skip_negative, s);
    }
    set_error("jg_put: unsupported itemsize %d", itemsize);
    return -2;
}

int jg_check_same_counts(const int64_t *offsets_a, const int64_t *offsets_b, int64_t n, int32_t *status,
                         void *stream) {
    if (n <= 0) return 0;
    same_counts_kernel<<<flat_grid(n, 256), 256, 0, as_stream(stream)>>>(offsets_a, offsets_b, n, status);
    return check_launch("same_counts_kernel");
}

int jg_index_gather(const void *content, int itemsize, const int64_t *offsets, const void *index, int index_dtype,
                    const int64_t *index_offsets, int64_t n, void *out, int32_t *status, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    switch (itemsize) {
        case 1: return launch_index_gather<uint8_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
        case 2: return launch_index_gather<uint16_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
        case 4: return launch_index_gather<uint32_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
        case 8: return launch_index_gather<uint64_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
    }
    set_error("jg_index_gather: unsupported itemsize %d", itemsize);
    return -2;
}

int jg_take(const void *content, int itemsize, const int64_t *index, int64_t n_index, void *out, void *stream) {
    if (n_index <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const int g = flat_grid(n_index, 256);
    switch (itemsize) {
        case 1: take_kernel<uint8_t><<<g, 256, 0, s>>>(static_cast<const uint8_t *>(content), index, n_index, static_cast<uint8_t *>(out)); break;
        case 2: take_kernel<uint16_t><<<g, 256, 0, s>>>(static_cast<const uint16_t *>(content), index, n_index, static_cast<uint16_t *>(out)); break;
        case 4: take_kernel<uint32_t><<<g, 256, 0, s>>>(static_cast<const uint32_t *>(content), index, n_index, static_cast<uint32_t *>(out)); break;
        case 8: take_kernel<uint64_t><<<g, 256, 0, s>>>(static_cast<const uint64_t *>(content), index, n_index, static_cast<uint64_t *>(out)); break;
        default:
            set_error("jg_take: unsupported itemsize %d", itemsize);
            return -2;
    }
    return check_launch("take_kernel");
}

int jg_flat_compact(const void *content, int itemsize, const uint8_t *mask, int64_t n, void *out, int64_t *positions,
                    int64_t *total, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && positions != nullptr, "jg_flat_compact: bad arguments");
    cudaStream_t s = as_stream(stream);
    int rc = launch_scan_array(mask, n, positions, total, nullptr, ws, ws_bytes, s);
    if (rc || n == 0 || content == nullptr) return rc;
    const int g = flat_grid(n, 256);
    switch (itemsize) {
        case 1: flat_scatter_kernel<uint8_t><<<g, 256, 0, s>>>(static_cast<const uint8_t *>(content), mask, positions, n, static_cast<uint8_t *>(out)); break;
        case 2: flat_scatter_kernel<uint16_t><<<g, 256, 0, s>>>(static_cast<const uint16_t *>(content), mask, positions, n, static_cast<uint16_t *>(out)); break;
        case 4: flat_scatter_kernel<uint32_t><<<g, 256, 0, s>>>(static_cast<const uint32_t *>(content), mask, positions, n, static_cast<uint32_t *>(out)); break;
        case 8: flat_scatter_kernel<uint64_t><<<g, 256, 0, s>>>(static_cast<const uint64_t *>(content), mask, positions, n, static_cast<uint64_t *>(out)); break;
        default:
            set_error("jg_flat_compact: unsupported itemsize %d", itemsize);
            return -2;
    }
    return check_launch("flat_scatter_kernel");
}

int jg_compact_nonneg(const int64_t *best, int64_t n, int64_t *out_offsets, int64_t *out_index, int64_t *total,
                      void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = as_stream(stream);
    XfNonNeg xf{0, 0, out_index};
    return launch_scan_tma<int64_t, XfNonNeg, kScanTmaItems>(xf, best, static_cast<const int64_t *>(nullptr), n, out_offsets, total,
                                                            nullptr, ws, ws_bytes, s);
}

}  // extern "C"

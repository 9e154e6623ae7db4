// jg_reduce.cu -- segmented reductions: the replacement for numpy ufunc.reduceat (jagged.py:1552-1563)
// behind JaggedArray.sum/prod/min/max/any/all (base.py:193-215) and argmin/argmax (jagged.py:1567-1601).
//
// One CTA owns an event tile (512 events, jg_tile.cuh).  Per CTA, chosen on the device from the tile's own
// element count (no host synchronisation, so the choice follows the count distribution tile by tile):
//   * SHORT events (tile content fits the shared-memory stage): the content run of the whole tile is staged
//     with one TMA bulk copy (UBLKCP) and each thread reduces whole events sequentially out of shared memory
//     -- left to right, exactly the reference's order, so min/max/argmax ties and -0.0 behave identically;
//   * LONG events (tile content exceeds the stage): warp-per-event -- the 32 lanes stride the event with
//     coalesced loads and combine with shuffles (events >= kBlockEventMin elements are reduced by the whole
//     CTA instead: block-per-event).
// Semantics: NaN -> identity, empty -> identity, first-occurrence argmax,
// "later element wins" on equal min/max (numpy's minimum/maximum loops), sum of a lone -0.0 stays -0.0.
//
// Algorithmic traffic: v*M + 8(N+1) read, v_out*N written.
#include "jg_reduce.cuh"
#include "jg_scan.cuh"
#include "jg_tile.cuh"

namespace jg {

// (launch bounds = the residency shared memory allows: the long-event paths below must not cost the staged path a CTA
// through register allocation, profiles/r1_owner_table_and_launch_bounds.md)
template <int OP, typename T, typename A, int STAGE_BYTES>
__global__ void __launch_bounds__(kTileThreads, STAGE_BYTES > 16384 ? 6 : 8)
segreduce_kernel(const T *__restrict__ content, const int64_t *__restrict__ offsets, int64_t n,
                 A *__restrict__ out) {
    __shared__ int64_t s_off[kTileEvents + 1];
    // (a stage of more than 32 KB is dynamic shared memory: the static limit is 48 KB per CTA)
    __shared__ __align__(16) unsigned char s_stage_static[STAGE_BYTES > 32768 ? 16 : STAGE_BYTES + 32];
    extern __shared__ __align__(16) unsigned char s_stage_dynamic[];
    unsigned char *const s_stage = STAGE_BYTES > 32768 ? s_stage_dynamic : s_stage_static;
    __shared__ uint64_t s_bar;
    __shared__ Partial<A> s_part[kTileThreads / kWarp];
    using R = Red<OP, T, A>;

    EventTile tile;
    if (threadIdx.x == 0) {
        // thread 0 reads the tile's two boundary offsets and starts the TMA copy of the content run straight
        // away, so that it overlaps the cooperative load of the offsets tile below
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        int64_t ev0;
        int nev;
        EventTile::range(n, blockIdx.x, ev0, nev);
        const int64_t a = __ldg(offsets + ev0), b = __ldg(offsets + ev0 + nev);
        const int64_t nb = (b - a) * static_cast<int64_t>(sizeof(T));
        if (nb > 0 && nb <= STAGE_BYTES)
            stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + a), static_cast<uint32_t>(nb), &s_bar);
    }
    tile.load(s_off, offsets, n, blockIdx.x);
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const int64_t nbytes = (e1 - e0) * static_cast<int64_t>(sizeof(T));

    if (nbytes <= STAGE_BYTES) {
        // ---------------- SHORT: stage the tile's content run, thread-per-event out of shared memory --------
        if (nbytes > 0) {
            const uint32_t shift = stage_edges<sizeof(T)>(
                s_stage, reinterpret_cast<const unsigned char *>(content + e0), static_cast<uint32_t>(nbytes));
            stage_wait(&s_bar, 0);
            const T *sm = reinterpret_cast<const T *>(s_stage + shift);
            // Thread t reduces events t and t+256 TOGETHER, one element of each per step, predicated: the kernel is
            // instruction-issue bound (ncu: issue-active ~80 %), and this shares the loop overhead between two
            // events and lets a warp iterate max-over-64-events once instead of max-over-32 twice.  Each event is
            // still reduced strictly left to right (the reference's order).  Float sum / prod run without the
            // per-element NaN test; a NaN result (a NaN input, or inf - inf) sends that one event to the exact loop.
            constexpr bool FAST = IsFloat<T>::value && (OP == JG_SUM || OP == JG_PROD);
            const int evA = threadIdx.x, evB = threadIdx.x + kTileThreads;
            int aA = 0, lenA = 0, aB = 0, lenB = 0;
            if (evA < tile.nev) {
                aA = static_cast<int>(s_off[evA] - e0);
                lenA = static_cast<int>(s_off[evA + 1] - e0) - aA;
            }
            if (evB < tile.nev) {
                aB = static_cast<int>(s_off[evB] - e0);
                lenB = static_cast<int>(s_off[evB + 1] - e0) - aB;
            }
            const T *pA = sm + aA, *pB = sm + aB;
            A accA = R::identity(), accB = R::identity();
            if (lenA > 0) accA = FAST ? static_cast<A>(pA[0]) : R::lift(pA[0]);
            if (lenB > 0) accB = FAST ? static_cast<A>(pB[0]) : R::lift(pB[0]);
            const int len = max(lenA, lenB);
#pragma unroll 1
            for (int k = 1; k < len; k += 2) {
                if (k < lenA) accA = FAST ? R::combine(accA, static_cast<A>(pA[k])) : R::accumulate(accA, pA[k]);
                if (k < lenB) accB = FAST ? R::combine(accB, static_cast<A>(pB[k])) : R::accumulate(accB, pB[k]);
                if (k + 1 < lenA) accA = FAST ? R::combine(accA, static_cast<A>(pA[k + 1])) : R::accumulate(accA, pA[k + 1]);
                if (k + 1 < lenB) accB = FAST ? R::combine(accB, static_cast<A>(pB[k + 1])) : R::accumulate(accB, pB[k + 1]);
            }
            if (FAST) {
                if (accA != accA) {
                    accA = R::lift(pA[0]);
                    for (int k = 1; k < lenA; ++k) accA = R::combine(accA, R::lift(pA[k]));
                }
                if (accB != accB) {
                    accB = R::lift(pB[0]);
                    for (int k = 1; k < lenB; ++k) accB = R::combine(accB, R::lift(pB[k]));
                }
            }
            if (evA < tile.nev) out[tile.event0 + evA] = accA;
            if (evB < tile.nev) out[tile.event0 + evB] = accB;
        } else {
#pragma unroll 1
            for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) out[tile.event0 + ev] = R::identity();
        }
        return;
    }

    // ---------------- LONG: the tile's run does not fit the stage ----------------------------------------------
    // The events are taken in GROUPS of consecutive events whose run does fit (the largest such group, by a binary
    // search over the offsets tile): each group is staged with one bulk copy and reduced exactly like a short tile;
    // an event that exceeds the stage on its own is reduced by the whole CTA from global memory (16-byte loads, four in
    // flight per thread).  A tile of 511 short events and one of 70 000 elements is two staged groups and one long
    // event, ~25 us; round 1 reduced all 512 events one warp at a time from global memory (64 dependent round trips
    // per warp, ~250 us), and a kernel ends when its slowest CTA does: 1000 such tiles among 150 M events cost the
    // sum 0.22 ms of 1.12 (profiles/r2_exp32.log).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int64_t kCap = STAGE_BYTES / static_cast<int64_t>(sizeof(T));
    uint32_t phase = 0;
    int g0 = 0;
#pragma unroll 1
    while (g0 < tile.nev) {
        const int64_t start = s_off[g0];
        int lo = g0, hi = tile.nev;                      // largest g1 with off[g1] - off[g0] <= kCap
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_off[mid] - start <= kCap) lo = mid;
            else hi = mid - 1;
        }
        if (lo == g0) {
            // ---- one event longer than the stage: the whole CTA
            const int64_t a = start, b = s_off[g0 + 1];
            Partial<A> p;
            if constexpr (IsFloat<T>::value && (OP == JG_SUM || OP == JG_PROD)) {
                // float sums in chunks of 1024 elements per thread: a thread's running float32 sum over millions of
                // elements drops the low bits of everything it adds (0.8 % on a 2^31-element event); chunk totals keep
                // the order-of-summation error at the 1e-6 level for any length.  (Only where rounding exists: the
                // chunk loop costs the min / max instantiations 8 registers and a resident CTA.)
                p.has = false;
                p.val = R::identity();
                p.pos = -1;
                constexpr int64_t kChunk = static_cast<int64_t>(kTileThreads) * 1024;
                for (int64_t c = a; c < b; c += kChunk)
                    p = merge<OP, T, A>(p, strided_partial<OP, T, A>(content, c, min(b, c + kChunk), threadIdx.x, kTileThreads));
            } else {
                p = strided_partial<OP, T, A>(content, a, b, threadIdx.x, kTileThreads);
            }
            p = shuffle_merge<OP, T, A>(p);
            __syncthreads();
            if (lane == 0) s_part[warp] = p;
            __syncthreads();
            if (warp == 0) {
                Partial<A> q;
                if (lane < kTileThreads / kWarp) q = s_part[lane];
                else { q.has = false; q.val = R::identity(); q.pos = -1; }
                q = shuffle_merge<OP, T, A>(q);
                if (lane == 0) out[tile.event0 + g0] = q.has ? q.val : R::identity();
            }
            ++g0;
            continue;
        }
        // ---- events [g0, g1): staged, thread-per-event (two events per thread as in the short path)
        const int g1 = lo;
        const int64_t nb = (s_off[g1] - start) * static_cast<int64_t>(sizeof(T));
        __syncthreads();                                 // the stage is free (the previous group has been read)
        const T *sm = reinterpret_cast<const T *>(s_stage);
        if (nb > 0) {
            if (threadIdx.x == 0) {
                fence_proxy_async_smem();
                stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + start), static_cast<uint32_t>(nb), &s_bar);
            }
            const uint32_t shift = stage_edges<sizeof(T)>(
                s_stage, reinterpret_cast<const unsigned char *>(content + start), static_cast<uint32_t>(nb));
            stage_wait(&s_bar, phase);
            phase ^= 1u;
            sm = reinterpret_cast<const T *>(s_stage + shift);
        }
#pragma unroll 1
        for (int ev = g0 + threadIdx.x; ev < g1; ev += kTileThreads) {
            const int a = static_cast<int>(s_off[ev] - start), len = static_cast<int>(s_off[ev + 1] - start) - a;
            A acc = R::identity();
            if (len > 0) {
                acc = R::lift(sm[a]);
#pragma unroll 1
                for (int k = 1; k < len; ++k) acc = R::accumulate(acc, sm[a + k]);
            }
            out[tile.event0 + ev] = acc;
        }
        g0 = g1;
    }
}

// ------------------------------------------------------------------------------------------------------------
// argmin / argmax: best[i] = local index of the FIRST extreme non-NaN element, -1 if the event has none.
// ------------------------------------------------------------------------------------------------------------
template <bool ISMIN, typename T, int STAGE_BYTES>
__global__ void __launch_bounds__(kTileThreads, STAGE_BYTES > 16384 ? 6 : 8)
segarg_kernel(const T *__restrict__ content, const int64_t *__restrict__ offsets, int64_t n,
              int64_t *__restrict__ best, T *__restrict__ best_value, int64_t *__restrict__ tile_totals) {
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) unsigned char s_stage[STAGE_BYTES + 32];
    __shared__ uint64_t s_bar;
    __shared__ ArgPartial<T> s_part[kTileThreads / kWarp];
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;

    EventTile tile;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        int64_t ev0;
        int nev;
        EventTile::range(n, blockIdx.x, ev0, nev);
        const int64_t a = __ldg(offsets + ev0), b = __ldg(offsets + ev0 + nev);
        const int64_t nb = (b - a) * static_cast<int64_t>(sizeof(T));
        if (nb > 0 && nb <= STAGE_BYTES)
            stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + a), static_cast<uint32_t>(nb), &s_bar);
    }
    tile.load(s_off, offsets, n, blockIdx.x);
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const int64_t nbytes = (e1 - e0) * static_cast<int64_t>(sizeof(T));

    if (nbytes <= STAGE_BYTES) {
        const T *sm = nullptr;
        if (nbytes > 0) {
            const uint32_t shift = stage_edges<sizeof(T)>(
                s_stage, reinterpret_cast<const unsigned char *>(content + e0), static_cast<uint32_t>(nbytes));
            stage_wait(&s_bar, 0);
            sm = reinterpret_cast<const T *>(s_stage + shift);
        }
        int nvalid = 0;
#pragma unroll 1
        for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
            const int a = static_cast<int>(s_off[ev] - e0), b = static_cast<int>(s_off[ev + 1] - e0);
            int idx = -1;
            T bv = T(0);
#pragma unroll 1
            for (int j = a; j < b; ++j) {
                const T x = sm[j];
                // (x better than bv) is false for NaN; the first non-NaN value is always taken
                const bool take = (ISMIN ? (x < bv) : (x > bv)) || (idx < 0 && !is_nan(x));
                if (take) {
                    bv = x;
                    idx = j - a;
                }
            }
            best[tile.event0 + ev] = idx;
            if (best_value) best_value[tile.event0 + ev] = bv;
            nvalid += (idx >= 0);
        }
        if (tile_totals) {
            nvalid = warp_reduce_add(nvalid);
            if ((threadIdx.x & 31) == 0 && nvalid) atomicAdd(&s_valid, nvalid);
            __syncthreads();
            if (threadIdx.x == 0) tile_totals[blockIdx.x] = s_valid;
        }
        return;
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int ev = warp; ev < tile.nev; ev += kTileThreads / kWarp) {
        const int64_t a = s_off[ev], b = s_off[ev + 1];
        if (b - a >= kBlockEventMin) continue;
        ArgPartial<T> p = arg_shuffle<ISMIN, T>(arg_strided<ISMIN, T>(content, a, b, lane, kWarp));
        if (lane == 0) {
            best[tile.event0 + ev] = p.idx;
            if (best_value) best_value[tile.event0 + ev] = p.val;
            if (p.idx >= 0) atomicAdd(&s_valid, 1);
        }
    }
    for (int ev = 0; ev < tile.nev; ++ev) {
        const int64_t a = s_off[ev], b = s_off[ev + 1];
        if (b - a < kBlockEventMin) continue;
        ArgPartial<T> p = arg_shuffle<ISMIN, T>(arg_strided<ISMIN, T>(content, a, b, threadIdx.x, kTileThreads));
        __syncthreads();
        if (lane == 0) s_part[warp] = p;
        __syncthreads();
        if (warp == 0) {
            ArgPartial<T> q;
            if (lane < kTileThreads / kWarp) q = s_part[lane];
            else { q.idx = -1; q.val = T(0); }
            q = arg_shuffle<ISMIN, T>(q);
            if (lane == 0) {
                best[tile.event0 + ev] = q.idx;
                if (best_value) best_value[tile.event0 + ev] = q.val;
                if (q.idx >= 0) atomicAdd(&s_valid, 1);
            }
        }
    }
    if (tile_totals) {
        __syncthreads();
        if (threadIdx.x == 0) tile_totals[blockIdx.x] = s_valid;
    }
}

// ------------------------------------------------------------------------------------------------------------
// OPTIMISTIC one-pass dense argmin / argmax.  For almost every input "the event has a result" == "the event is
// not empty" (the exception: a non-empty event whose values are all NaN).  The result offsets then depend on the
// offsets only: count_nonempty_kernel (8N bytes) + a tiny scan give every tile its base, and segarg_direct_kernel
// writes out_offsets while its content run is still in flight and out_index straight from the arg-reduce -- no
// best[] round trip, no compaction pass.  An all-NaN non-empty event raises JG_ST_ARG_ALLNAN and the host redoes
// the operation with the exact three-kernel path above (jg_segargminmax + jg_compact_nonneg).
// Traffic: 8N + vM + 8(N+1) read, 8(N+1) + 8N+ written.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
count_nonempty_kernel(const int64_t *__restrict__ offsets, int64_t n, int64_t ntiles, int64_t *__restrict__ tile_totals) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = static_cast<int64_t>(blockIdx.x) * (kTileThreads / kWarp) + warp;
    if (tile >= ntiles) return;
    const int64_t ev0 = tile * kTileEvents;
    const int64_t ev1 = min(n, ev0 + kTileEvents);
    int cnt = 0;
    for (int64_t i = ev0 + lane; i < ev1; i += kWarp) cnt += (__ldg(offsets + i + 1) > __ldg(offsets + i));
    cnt = warp_reduce_add(cnt);
    if (lane == 0) tile_totals[tile] = cnt;
}

// A tile whose run does not fit the stage: groups of consecutive events that do are staged and arg-reduced by one thread
// per event, an event longer than the stage by the whole CTA from global memory (see the long path of segreduce_kernel).
// s_idx64[ev] = local index of the event's first extreme non-NaN item or -1; ends with a barrier.
template <bool ISMIN, typename T, int STAGE_BYTES>
__device__ __noinline__ void segarg_long_tile(const T *__restrict__ content, const int64_t *s_off, int nev,
                                              unsigned char *s_stage, uint64_t *s_bar, ArgPartial<T> *s_part,
                                              int64_t *s_idx64) {
    constexpr int NW = kTileThreads / kWarp;
    constexpr int64_t kCap = STAGE_BYTES / static_cast<int64_t>(sizeof(T));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t phase = 0;
    int g0 = 0;
#pragma unroll 1
    while (g0 < nev) {
        const int64_t start = s_off[g0];
        int lo = g0, hi = nev;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_off[mid] - start <= kCap) lo = mid;
            else hi = mid - 1;
        }
        if (lo == g0) {
            ArgPartial<T> p = arg_shuffle<ISMIN, T>(arg_strided<ISMIN, T, 2>(content, start, s_off[g0 + 1], threadIdx.x, kTileThreads));
            __syncthreads();
            if (lane == 0) s_part[warp] = p;
            __syncthreads();
            if (warp == 0) {
                ArgPartial<T> q;
                if (lane < NW) q = s_part[lane];
                else { q.idx = -1; q.val = T(0); }
                q = arg_shuffle<ISMIN, T>(q);
                if (lane == 0) s_idx64[g0] = q.idx;
            }
            ++g0;
            continue;
        }
        const int g1 = lo;
        const int64_t nb = (s_off[g1] - start) * static_cast<int64_t>(sizeof(T));
        __syncthreads();                             // the stage is free (the previous group has been read)
        const T *sm = reinterpret_cast<const T *>(s_stage);
        if (nb > 0) {
            if (threadIdx.x == 0) {
                fence_proxy_async_smem();
                stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + start), static_cast<uint32_t>(nb), s_bar);
            }
            const uint32_t shift = stage_edges<sizeof(T)>(
                s_stage, reinterpret_cast<const unsigned char *>(content + start), static_cast<uint32_t>(nb));
            stage_wait(s_bar, phase);
            phase ^= 1u;
            sm = reinterpret_cast<const T *>(s_stage + shift);
        }
#pragma unroll 1
        for (int ev = g0 + threadIdx.x; ev < g1; ev += kTileThreads) {
            const T *pe = sm + static_cast<int>(s_off[ev] - start);
            const int len = static_cast<int>(s_off[ev + 1] - s_off[ev]);
            int ibest = -1;
            T vbest = T(0);
#pragma unroll 1
            for (int k = 0; k < len; ++k) {
                const T x = pe[k];
                if (!is_nan(x) && (ibest < 0 || (ISMIN ? (x < vbest) : (x > vbest)))) {
                    vbest = x;
                    ibest = k;
                }
            }
            s_idx64[ev] = ibest;
        }
        g0 = g1;
    }
    __syncthreads();
}

template <bool ISMIN, typename T, int STAGE_BYTES>
__global__ void __launch_bounds__(kTileThreads, STAGE_BYTES > 16384 ? 5 : 8)
segarg_direct_kernel(const T *__restrict__ content, const int64_t *__restrict__ offsets, int64_t n,
                     const int64_t *__restrict__ tile_bases, int64_t *__restrict__ out_offsets,
                     int64_t *__restrict__ out_index, T *__restrict__ out_value, int32_t *__restrict__ status,
                     uint32_t ntiles) {
    constexpr int NW = kTileThreads / kWarp;
    __shared__ int64_t s_off[kTileEvents + 1];
    // (a stage of more than 32 KB is dynamic shared memory: the static limit is 48 KB per CTA)
    __shared__ __align__(16) unsigned char s_stage_static[STAGE_BYTES > 32768 ? 16 : STAGE_BYTES + 32];
    extern __shared__ __align__(16) unsigned char s_stage_dynamic[];
    unsigned char *const s_stage = STAGE_BYTES > 32768 ? s_stage_dynamic : s_stage_static;
    __shared__ uint64_t s_bar;
    __shared__ ArgPartial<T> s_part[NW];
    __shared__ int32_t s_scan[NW + 1];

    EventTile tile;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        int64_t ev0;
        int nev;
        EventTile::range(n, blockIdx.x, ev0, nev);
        const int64_t a = __ldg(offsets + ev0), b = __ldg(offsets + ev0 + nev);
        const int64_t nb = (b - a) * static_cast<int64_t>(sizeof(T));
        if (nb > 0 && nb <= STAGE_BYTES)
            stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + a), static_cast<uint32_t>(nb), &s_bar);
    }
    tile.load(s_off, offsets, n, blockIdx.x);
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const int64_t nbytes = (e1 - e0) * static_cast<int64_t>(sizeof(T));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // result offsets from the offsets alone (thread t owns events 2t, 2t+1): written while the content is in flight
    const int ev0 = threadIdx.x * 2;
    int64_t a[2] = {0, 0}, b[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 2; ++q)
        if (ev0 + q < tile.nev) {
            a[q] = s_off[ev0 + q];
            b[q] = s_off[ev0 + q + 1];
        }
    const int ne0 = b[0] > a[0], ne1 = b[1] > a[1];
    // events with a result before this thread's two: two ballots and popcounts inside the warp, one REDUX over the
    // totals of the warps before it (one barrier; a generic block scan costs three and ~45 instructions per warp)
    int32_t excl;
    {
        const unsigned v0 = __ballot_sync(0xffffffffu, ne0), v1 = __ballot_sync(0xffffffffu, ne1);
        const unsigned lt = (1u << lane) - 1u;
        if (lane == 0) s_scan[warp] = __popc(v0) + __popc(v1);
        __syncthreads();
        excl = __reduce_add_sync(0xffffffffu, lane < warp ? s_scan[lane] : 0) + __popc(v0 & lt) + __popc(v1 & lt);
    }
    const int64_t base = __ldg(tile_bases + blockIdx.x) + excl;
    {
        int64_t *dst = out_offsets + tile.event0 + ev0;
        if (ev0 + 1 < tile.nev && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            __stcs(reinterpret_cast<longlong2 *>(dst), make_longlong2(base, base + ne0));
        } else {
            if (ev0 < tile.nev) dst[0] = base;
            if (ev0 + 1 < tile.nev) dst[1] = base + ne0;
        }
        if (blockIdx.x == ntiles - 1 && threadIdx.x == 0) out_offsets[n] = __ldg(tile_bases + ntiles);
    }

    int64_t idx[2] = {-1, -1};
    T val[2] = {T(0), T(0)};          // the extreme value itself: the "leading object" a[a.argmax()] for free
    if (nbytes <= STAGE_BYTES) {
        const T *sm = nullptr;
        if (nbytes > 0) {
            const uint32_t shift = stage_edges<sizeof(T)>(
                s_stage, reinterpret_cast<const unsigned char *>(content + e0), static_cast<uint32_t>(nbytes));
            stage_wait(&s_bar, 0);
            sm = reinterpret_cast<const T *>(s_stage + shift);
        }
        // both events of the thread advance together, one element each per step (predicated); the running best
        // starts at the identity, so "strictly better" needs no NaN or first-element test.  An event that ends
        // without a result although it is not empty (only NaN, or only identity values) takes the exact loop.
        const T *pA = sm + static_cast<int>(a[0] - e0), *pB = sm + static_cast<int>(a[1] - e0);
        const int lenA = static_cast<int>(b[0] - a[0]), lenB = static_cast<int>(b[1] - a[1]);
        T bvA = ISMIN ? min_identity<T>() : max_identity<T>(), bvB = bvA;
        int iA = -1, iB = -1;
        const int len = max(lenA, lenB);
#pragma unroll 1
        for (int k = 0; k < len; ++k) {
            if (k < lenA) {
                const T x = pA[k];
                if (ISMIN ? (x < bvA) : (x > bvA)) {
                    bvA = x;
                    iA = k;
                }
            }
            if (k < lenB) {
                const T x = pB[k];
                if (ISMIN ? (x < bvB) : (x > bvB)) {
                    bvB = x;
                    iB = k;
                }
            }
        }
        if (lenA > 0 && iA < 0)
            for (int k = 0; k < lenA; ++k)
                if (!is_nan(pA[k])) {      // every non-NaN value equals the identity: the first one wins
                    iA = k;
                    break;
                }
        if (lenB > 0 && iB < 0)
            for (int k = 0; k < lenB; ++k)
                if (!is_nan(pB[k])) {
                    iB = k;
                    break;
                }
        idx[0] = iA;
        idx[1] = iB;
        val[0] = (iA >= 0) ? pA[iA] : T(0);
        val[1] = (iB >= 0) ? pB[iB] : T(0);
    } else {
        // the tile's run does not fit the stage (a separate function: inlined, its live ranges cost the staged path a
        // spilled register and 7 % of its speed)
        __shared__ int64_t s_idx64[kTileEvents];
        segarg_long_tile<ISMIN, T, STAGE_BYTES>(content, s_off, tile.nev, s_stage, &s_bar, s_part, s_idx64);
        // (results written here: joining the staged path's idx / val below costs that path a spilled register)
        bool missing = false;
        if (ne0) {
            const int64_t i = s_idx64[ev0];
            missing |= i < 0;
            out_index[base] = i;
            if (out_value) out_value[base] = i >= 0 ? __ldg(content + s_off[ev0] + i) : T(0);
        }
        if (ne1) {
            const int64_t i = s_idx64[ev0 + 1];
            missing |= i < 0;
            out_index[base + ne0] = i;
            if (out_value) out_value[base + ne0] = i >= 0 ? __ldg(content + s_off[ev0 + 1] + i) : T(0);
        }
        if (missing) atomicOr(status, JG_ST_ARG_ALLNAN);
        return;
    }
    bool mismatch = false;
    if (ne0) {
        mismatch |= idx[0] < 0;
        out_index[base] = idx[0];
        if (out_value) out_value[base] = val[0];
    }
    if (ne1) {
        mismatch |= idx[1] < 0;
        out_index[base + ne0] = idx[1];
        if (out_value) out_value[base + ne0] = val[1];
    }
    if (mismatch) atomicOr(status, JG_ST_ARG_ALLNAN);
}

template <typename T> struct StageBytes { static constexpr int value = sizeof(T) >= 8 ? 32768 : (sizeof(T) >= 4 ? 16384 : 8192); };

// JG_WIDE_STAGE (a hint OR-ed into the operation by a caller who knows the mean count is 7 - 14): 4-byte contents are
// staged 32 KB per tile instead of 16 KB, so that the 512 events of a tile still fit and such arrays stay on the
// thread-per-event kernels (5 - 6 resident CTAs per SM instead of 8, twice the bytes each).  Same results either way.
template <typename T> struct WideStage { static constexpr bool value = sizeof(T) == 4; };
// JG_WIDER_STAGE (mean 14 - 28): 64 KB per tile, dynamic shared memory, 3 resident CTAs per SM with the same bytes in flight
constexpr int kWiderStage = 65536;
template <class K>
static void allow_wider_stage(K kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWiderStage + 32);
}

template <int OP, typename T, typename A>
static int launch_reduce(const void *content, const int64_t *offsets, int64_t n, void *out, cudaStream_t s, int wide = 0) {
    if constexpr (WideStage<T>::value) {
        if (wide == 2) {
            static bool allowed = false;
            if (!allowed) {
                allow_wider_stage(segreduce_kernel<OP, T, A, kWiderStage>);
                allowed = true;
            }
            segreduce_kernel<OP, T, A, kWiderStage><<<tile_grid(n), kTileThreads, kWiderStage + 32, s>>>(
                static_cast<const T *>(content), offsets, n, static_cast<A *>(out));
            return check_launch("segreduce_kernel<wider>");
        }
        if (wide) {
            segreduce_kernel<OP, T, A, 32768><<<tile_grid(n), kTileThreads, 0, s>>>(
                static_cast<const T *>(content), offsets, n, static_cast<A *>(out));
            return check_launch("segreduce_kernel<wide>");
        }
    }
    segreduce_kernel<OP, T, A, StageBytes<T>::value><<<tile_grid(n), kTileThreads, 0, s>>>(
        static_cast<const T *>(content), offsets, n, static_cast<A *>(out));
    return check_launch("segreduce_kernel");
}

template <typename T, typename ASUM>
static int dispatch_reduce(int op, const void *content, const int64_t *offsets, int64_t n, void *out, cudaStream_t s) {
    const int wide = (op & JG_WIDER_STAGE) ? 2 : ((op & JG_WIDE_STAGE) ? 1 : 0);
    switch (op & ~(JG_WIDE_STAGE | JG_WIDER_STAGE)) {
        case JG_SUM: return launch_reduce<JG_SUM, T, ASUM>(content, offsets, n, out, s, wide);
        case JG_PROD: return launch_reduce<JG_PROD, T, ASUM>(content, offsets, n, out, s, wide);
        case JG_MIN: return launch_reduce<JG_MIN, T, T>(content, offsets, n, out, s, wide);
        case JG_MAX: return launch_reduce<JG_MAX, T, T>(content, offsets, n, out, s, wide);
        case JG_ANY: return launch_reduce<JG_ANY, T, uint8_t>(content, offsets, n, out, s, wide);
        case JG_ALL: return launch_reduce<JG_ALL, T, uint8_t>(content, offsets, n, out, s, wide);
    }
    set_error("jg_segreduce: unsupported op %d", op);
    return -2;
}

template <typename T>
static int dispatch_arg(int is_min, const void *content, const int64_t *offsets, int64_t n, int64_t *best,
                        void *best_value, cudaStream_t s) {
    int64_t *tile_totals = nullptr;   // (per-tile counts of events with a result: kept for a future exact dense path)
    if (is_min)
        segarg_kernel<true, T, StageBytes<T>::value><<<tile_grid(n), kTileThreads, 0, s>>>(
            static_cast<const T *>(content), offsets, n, best, static_cast<T *>(best_value), tile_totals);
    else
        segarg_kernel<false, T, StageBytes<T>::value><<<tile_grid(n), kTileThreads, 0, s>>>(
            static_cast<const T *>(content), offsets, n, best, static_cast<T *>(best_value), tile_totals);
    return check_launch("segarg_kernel");
}

template <typename T>
static int dispatch_arg_dense(int is_min, const void *content, const int64_t *offsets, int64_t n, int64_t *out_offsets,
                              int64_t *out_index, void *out_value, int64_t *total, int32_t *status, void *ws,
                              size_t ws_bytes, cudaStream_t s, const int64_t *tile_nonempty) {
    const int64_t ntiles = tile_grid(n);
    // workspace: [totals: nt+2][bases: nt+2][scan descriptors]
    const size_t arr = ((static_cast<size_t>(ntiles) + 2) * 8 + 255) & ~static_cast<size_t>(255);
    const size_t scan_need = sizeof(LookbackWs) + static_cast<size_t>(ceil_div(ntiles, kScanTile) + 1) * 8;
    JG_REQUIRE(ws != nullptr && ws_bytes >= 2 * arr + scan_need + 256,
               "jg_segargminmax_dense: workspace too small (%zu < %zu)", ws_bytes, 2 * arr + scan_need + 256);
    JG_REQUIRE(status != nullptr, "jg_segargminmax_dense: status word required");
    int64_t *totals = static_cast<int64_t *>(ws);
    int64_t *bases = reinterpret_cast<int64_t *>(static_cast<char *>(ws) + arr);
    void *scan_ws = static_cast<char *>(ws) + 2 * arr;
    int rc = 0;
    if (tile_nonempty == nullptr) {
        count_nonempty_kernel<<<static_cast<unsigned>(ceil_div(ntiles, kTileThreads / kWarp)), kTileThreads, 0, s>>>(
            offsets, n, ntiles, totals);
        rc = check_launch("count_nonempty_kernel");
        if (rc) return rc;
        tile_nonempty = totals;
    }
    rc = launch_scan_array(tile_nonempty, ntiles, bases, total, nullptr, scan_ws, ws_bytes - 2 * arr, s);
    if (rc) return rc;
    const int wide = (is_min & JG_WIDER_STAGE) ? 2 : ((is_min & JG_WIDE_STAGE) ? 1 : 0);
    is_min &= ~(JG_WIDE_STAGE | JG_WIDER_STAGE);
    if constexpr (WideStage<T>::value) {
        if (wide == 2) {
            static bool allowed = false;
            if (!allowed) {
                allow_wider_stage(segarg_direct_kernel<true, T, kWiderStage>);
                allow_wider_stage(segarg_direct_kernel<false, T, kWiderStage>);
                allowed = true;
            }
            if (is_min)
                segarg_direct_kernel<true, T, kWiderStage><<<static_cast<unsigned>(ntiles), kTileThreads, kWiderStage + 32, s>>>(
                    static_cast<const T *>(content), offsets, n, bases, out_offsets, out_index, static_cast<T *>(out_value),
                    status, static_cast<uint32_t>(ntiles));
            else
                segarg_direct_kernel<false, T, kWiderStage><<<static_cast<unsigned>(ntiles), kTileThreads, kWiderStage + 32, s>>>(
                    static_cast<const T *>(content), offsets, n, bases, out_offsets, out_index, static_cast<T *>(out_value),
                    status, static_cast<uint32_t>(ntiles));
            return check_launch("segarg_direct_kernel<wider>");
        }
        if (wide) {
            if (is_min)
                segarg_direct_kernel<true, T, 32768><<<static_cast<unsigned>(ntiles), kTileThreads, 0, s>>>(
                    static_cast<const T *>(content), offsets, n, bases, out_offsets, out_index, static_cast<T *>(out_value),
                    status, static_cast<uint32_t>(ntiles));
            else
                segarg_direct_kernel<false, T, 32768><<<static_cast<unsigned>(ntiles), kTileThreads, 0, s>>>(
                    static_cast<const T *>(content), offsets, n, bases, out_offsets, out_index, static_cast<T *>(out_value),
                    status, static_cast<uint32_t>(ntiles));
            return check_launch("segarg_direct_kernel<wide>");
        }
    }
    if (is_min)
        segarg_direct_kernel<true, T, StageBytes<T>::value><<<static_cast<unsigned>(ntiles), kTileThreads, 0, s>>>(
            static_cast<const T *>(content), offsets, n, bases, out_offsets, out_index, static_cast<T *>(out_value),
            status, static_cast<uint32_t>(ntiles));
    else
        segarg_direct_kernel<false, T, StageBytes<T>::value><<<static_cast<unsigned>(ntiles), kTileThreads, 0, s>>>(
            static_cast<const T *>(content), offsets, n, bases, out_offsets, out_index, static_cast<T *>(out_value),
            status, static_cast<uint32_t>(ntiles));
    return check_launch("segarg_direct_kernel");
}

}  // namespace jg

using namespace jg;

extern "C" {

int jg_segreduce(const void *content, int dtype, const int64_t *offsets, int64_t n, int op, void *out,
                 int64_t content_len, void *stream) {
    (void)content_len;
    JG_REQUIRE(n >= 0 && offsets && out, "jg_segreduce: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    switch (dtype) {
        case JG_F32: return dispatch_reduce<float, float>(op, content, offsets, n, out, s);
        case JG_F64: return dispatch_reduce<double, double>(op, content, offsets, n, out, s);
        case JG_I32: return dispatch_reduce<int32_t, int64_t>(op, content, offsets, n, out, s);
        case JG_I64: return dispatch_reduce<int64_t, int64_t>(op, content, offsets, n, out, s);
        case JG_U8: return dispatch_reduce<uint8_t, uint64_t>(op, content, offsets, n, out, s);
        case JG_I8: return dispatch_reduce<int8_t, int64_t>(op, content, offsets, n, out, s);
        case JG_I16: return dispatch_reduce<int16_t, int64_t>(op, content, offsets, n, out, s);
        case JG_U16: return dispatch_reduce<uint16_t, uint64_t>(op, content, offsets, n, out, s);
        case JG_U32: return dispatch_reduce<uint32_t, uint64_t>(op, content, offsets, n, out, s);
        case JG_U64: return dispatch_reduce<uint64_t, uint64_t>(op, content, offsets, n, out, s);
        case JG_B8:
            // bool: sum/prod -> int64, any/all -> bool; min/max -> float64 with +-inf identities (jagged.py:1530)
            switch (op) {
                case JG_SUM: return launch_reduce<JG_SUM, uint8_t, int64_t>(content, offsets, n, out, s);
                case JG_PROD: return launch_reduce<JG_PROD, uint8_t, int64_t>(content, offsets, n, out, s);
                case JG_MIN: return launch_reduce<JG_MIN, uint8_t, double>(content, offsets, n, out, s);
                case JG_MAX: return launch_reduce<JG_MAX, uint8_t, double>(content, offsets, n, out, s);
                case JG_ANY: return launch_reduce<JG_ANY, uint8_t, uint8_t>(content, offsets, n, out, s);
                case JG_ALL: return launch_reduce<JG_ALL, uint8_t, uint8_t>(content, offsets, n, out, s);
            }
            break;
    }
    set_error("jg_segreduce: unsupported dtype %d / op %d", dtype, op);
    return -2;
}

int jg_segargminmax(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min, int64_t *best,
                    void *best_value, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && best, "jg_segargminmax: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    switch (dtype) {
        case JG_F32: return dispatch_arg<float>(is_min, content, offsets, n, best, best_value, s);
        case JG_F64: return dispatch_arg<double>(is_min, content, offsets, n, best, best_value, s);
        case JG_I32: return dispatch_arg<int32_t>(is_min, content, offsets, n, best, best_value, s);
        case JG_I64: return dispatch_arg<int64_t>(is_min, content, offsets, n, best, best_value, s);
        case JG_U8:
        case JG_B8: return dispatch_arg<uint8_t>(is_min, content, offsets, n, best, best_value, s);
        case JG_I8: return dispatch_arg<int8_t>(is_min, content, offsets, n, best, best_value, s);
        case JG_I16: return dispatch_arg<int16_t>(is_min, content, offsets, n, best, best_value, s);
        case JG_U16: return dispatch_arg<uint16_t>(is_min, content, offsets, n, best, best_value, s);
        case JG_U32: return dispatch_arg<uint32_t>(is_min, content, offsets, n, best, best_value, s);
        case JG_U64: return dispatch_arg<uint64_t>(is_min, content, offsets, n, best, best_value, s);
    }
    set_error("jg_segargminmax: unsupported dtype %d", dtype);
    return -2;
}

int jg_segargminmax_dense_tiles(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min,
                                const int64_t *tile_nonempty, int64_t *out_offsets, int64_t *out_index, void *out_value,
                                int64_t *total, int32_t *status, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && out_offsets, "jg_segargminmax_dense: bad arguments");
    cudaStream_t s = as_stream(stream);
    switch (dtype) {
        case JG_F32: return dispatch_arg_dense<float>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_F64: return dispatch_arg_dense<double>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_I32: return dispatch_arg_dense<int32_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_I64: return dispatch_arg_dense<int64_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_U8:
        case JG_B8: return dispatch_arg_dense<uint8_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_I8: return dispatch_arg_dense<int8_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_I16: return dispatch_arg_dense<int16_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_U16: return dispatch_arg_dense<uint16_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_U32: return dispatch_arg_dense<uint32_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_U64: return dispatch_arg_dense<uint64_t>(is_min, content, offsets, n, out_offsets, out_index, out_value, total, status, ws, ws_bytes, s, tile_nonempty);
    }
    set_error("jg_segargminmax_dense: unsupported dtype %d", dtype);
    return -2;
}

int jg_segargminmax_dense(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min,
                          int64_t *out_offsets, int64_t *out_index, void *out_value, int64_t *total, int32_t *status,
                          void *ws, size_t ws_bytes, void *stream) {
    return jg_segargminmax_dense_tiles(content, dtype, offsets, n, is_min, nullptr, out_offsets, out_index, out_value, total,
                                       status, ws, ws_bytes, stream);
}

}  // extern "C"

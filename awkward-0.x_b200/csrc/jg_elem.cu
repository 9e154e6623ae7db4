// jg_elem.cu -- ELEMENT-tiled variants of the hot-path kernels, for count distributions whose events are long.
//
// The event tile of jg_tile.cuh (512 events per CTA, content run staged in 16 KB of shared memory) is sized for the
// short events of collider data (Poisson(5): 2560 elements per tile).  When the mean count exceeds ~7 every tile
// overflows the stage, and when a few events hold most of the content (a heavy tail, 1000 events of a million
// elements) a handful of CTAs do all the work.  The host therefore switches to the kernels of this file when
// content_elements > 7 * events (jagged.py:1552-1563 is the loop being replaced; north_star: "warp-per-segment and
// block-per-segment variants chosen by count distribution, 128-bit coalesced loads, TMA-staged tiles for long
// segments"):
//
//   * the unit of work is a tile of kElemStageBytes of CONTENT (4096 float32): the load is balanced whatever the
//     counts are, and every tile is one 16-byte-aligned TMA bulk copy;
//   * jg_elem_partition finds, for every tile, the first event that STARTS in it (one binary search per tile over the
//     resident offsets: a merge-path partition); a tile "owns" the events that start in it;
//   * inside a tile a segment (an owned event clipped to the tile, or the tail of an event that started earlier) is
//     reduced by ONE THREAD when it is short (< 64 elements, sequentially out of shared memory, the reference's order),
//     by ONE WARP when it is longer (conflict-free strided reads + shuffle merge on (value, position)), and by the
//     WHOLE CTA from 1024 elements on;
//   * an event that spans several tiles leaves one partial per tile (`carry`); a second tiny launch merges them in
//     tile order (one warp per spanning event, lanes stride the tiles), so a single event of 2^31 elements is spread
//     over half a million CTAs instead of one.
//
// Selection (a[mask], a[a > c]) is a flat stream compaction plus "rank at the event boundaries": tile totals, their
// scan, then the compaction of every staged tile, in which an owned event reads its new offset from the tile's on-chip
// per-position prefix (three launches that never wait on another CTA, as in jg_select.cu).  parents / localindex / broadcast build the owner table of a tile from the events that start
// in it (scatter + running maximum over 32-bit relative event numbers).
#include <stdlib.h>

#include <type_traits>

#include "jg_reduce.cuh"
#include "jg_compact.cuh"
#include "jg_scan.cuh"
#include "jg_tile.cuh"

namespace jg {

constexpr int kElemThreads = 256;
constexpr int kElemStageBytes = 16384;
constexpr int kElemWarpSeg = 64;        // segments at least this long: one warp
constexpr int kElemCtaSeg = 1024;       // segments at least this long: the whole CTA
constexpr int kElemWarps = kElemThreads / kWarp;
// the reductions run 128-thread CTAs: a tile's lifetime is dominated by the latency of its one bulk copy, so what counts
// is how many 16 KB tiles an SM has in flight (13 CTAs of 128 threads fit the shared memory, 8 CTAs of 256 the threads)
constexpr int kElemRedThreads = 128;
constexpr int kElemRedWarps = kElemRedThreads / kWarp;

template <typename T>
struct ElemTile {
    static constexpr int ET = kElemStageBytes / static_cast<int>(sizeof(T)) > 4096 ? 4096 : kElemStageBytes / static_cast<int>(sizeof(T));
};

static inline int elem_tile_items(int itemsize) {
    const int et = kElemStageBytes / itemsize;
    return et > 4096 ? 4096 : et;
}

// first_event[c] = smallest i in [0, n] with offsets[i] >= offsets[0] + c * et  (c = 0 .. ntiles - 1); first_event[ntiles] = n:
// tile c owns the events [first_event[c], first_event[c + 1]) -- those that start inside it; the last tile also owns
// the empty events at the very end.
__global__ void __launch_bounds__(256)
elem_partition_kernel(const int64_t *__restrict__ offsets, int64_t n, int64_t ntiles, int64_t et,
                      int64_t *__restrict__ first_event) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c > ntiles) return;
    if (c == ntiles) {
        first_event[c] = n;
        return;
    }
    const int64_t target = __ldg(offsets) + c * et;
    int64_t lo = 0, hi = n;                   // answer in [lo, hi]; offsets[n] >= target for every tile
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) < target) lo = mid + 1;
        else hi = mid;
    }
    first_event[c] = lo;
}

struct ElemRange {
    int64_t lo, hi;          // content positions of the tile
    int64_t ev_lo, ev_hi;    // owned events
    int64_t in_end;          // end of the segment that enters from the left (== lo when there is none)
};

__device__ __forceinline__ ElemRange elem_range(const int64_t *__restrict__ offsets, int64_t n,
                                                const int64_t *__restrict__ first_event, int64_t c, int64_t et) {
    ElemRange r;
    const int64_t base = __ldg(offsets), end = __ldg(offsets + n);
    r.lo = base + c * et;
    r.hi = r.lo + et < end ? r.lo + et : end;
    if (r.hi < r.lo) r.hi = r.lo;
    r.ev_lo = __ldg(first_event + c);
    r.ev_hi = __ldg(first_event + c + 1);
    const int64_t o = __ldg(offsets + r.ev_lo);           // ev_lo <= n
    r.in_end = o < r.hi ? o : r.hi;
    if (r.in_end < r.lo) r.in_end = r.lo;
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------------------
template <int OP, typename T, typename A>
__device__ __forceinline__ Partial<A> smem_strided_partial(const T *__restrict__ sm, int a, int b, int first, int step,
                                                           int64_t pos0) {
    using R = Red<OP, T, A>;
    Partial<A> p;
    p.has = false;
    p.val = R::identity();
    p.pos = -1;
#pragma unroll 1
    for (int j = a + first; j < b; j += step) {
        const A x = R::lift(sm[j]);
        if (!p.has) {
            p.val = x;
            p.pos = pos0 + j;
            p.has = true;
        } else if constexpr (OP == JG_MIN || OP == JG_MAX) {
            const bool keep = (OP == JG_MIN) ? (p.val < x) : (p.val > x);
            if (!keep) {
                p.val = x;
                p.pos = pos0 + j;
            }
        } else {
            p.val = R::combine(p.val, x);
        }
    }
    return p;
}

template <int OP, typename T, typename A>
__global__ void __launch_bounds__(kElemRedThreads)
segreduce_elem_kernel(const T *__restrict__ content, const int64_t *__restrict__ offsets, int64_t n,
                      const int64_t *__restrict__ first_event, A *__restrict__ out, A *__restrict__ carry) {
    constexpr int ET = ElemTile<T>::ET;
    using R = Red<OP, T, A>;
    __shared__ __align__(16) unsigned char s_stage[ET * sizeof(T) + 32];
    __shared__ uint64_t s_bar;
    __shared__ int64_t s_list[ET / kElemWarpSeg + 2];       // event numbers of the long segments (-1: the entering one)
    __shared__ int s_nlist;
    __shared__ Partial<A> s_part[kElemRedWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ElemRange r = elem_range(offsets, n, first_event, blockIdx.x, ET);
    const int L = static_cast<int>(r.hi - r.lo);
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        s_nlist = 0;
        if (L > 0) stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + r.lo), static_cast<uint32_t>(L * sizeof(T)), &s_bar);
    }
    __syncthreads();
    const T *sm = reinterpret_cast<const T *>(s_stage);
    // the offsets of the thread's first owned event are asked for BEFORE the wait for the tile, those of its next event
    // while the current one is reduced: a tile's life is one bulk copy (~1.5 us), and a dependent round trip to the
    // offsets behind it (one more per 128 owned events) was a third of it at mean counts of 10 - 40
    int64_t e = r.ev_lo + tid;
    int64_t a64 = 0, b64 = 0;
    if (e < r.ev_hi) {
        a64 = __ldg(offsets + e);
        b64 = __ldg(offsets + e + 1);
    }
    if (L > 0) {
        const uint32_t shift = stage_edges<sizeof(T)>(s_stage, reinterpret_cast<const unsigned char *>(content + r.lo),
                                                      static_cast<uint32_t>(L * sizeof(T)));
        sm = reinterpret_cast<const T *>(s_stage + shift);
        stage_wait(&s_bar, 0);
    }
    // the segment entering from the left belongs to an event that started in an earlier tile: its partial goes to carry[]
    const int in_len = static_cast<int>(r.in_end - r.lo);
    if (in_len >= kElemWarpSeg) {
        if (tid == 0) s_list[atomicAdd(&s_nlist, 1)] = -1;
    } else if (tid == 0) {
        A acc = R::identity();
        if (in_len > 0) {
            acc = R::lift(sm[0]);
            for (int k = 1; k < in_len; ++k) acc = R::accumulate(acc, sm[k]);
        }
        carry[blockIdx.x] = acc;
    }
    // owned events: short ones by their own thread, sequentially (the reference's order)
#pragma unroll 1
    while (e < r.ev_hi) {
        const int64_t en = e + kElemRedThreads;
        int64_t na = 0, nb = 0;
        if (en < r.ev_hi) {
            na = __ldg(offsets + en);
            nb = __ldg(offsets + en + 1);
        }
        const int a = static_cast<int>(a64 - r.lo);
        const int b = static_cast<int>((b64 < r.hi ? b64 : r.hi) - r.lo);
        if (b - a >= kElemWarpSeg) {
            s_list[atomicAdd(&s_nlist, 1)] = e;
        } else {
            A acc = R::identity();
            if (b > a) {
                acc = R::lift(sm[a]);
#pragma unroll 1
                for (int k = a + 1; k < b; ++k) acc = R::accumulate(acc, sm[k]);
            }
            out[e] = acc;
        }
        e = en;
        a64 = na;
        b64 = nb;
    }
    __syncthreads();
    // long segments: one warp each, or the whole CTA
    const int nlist = s_nlist;
#pragma unroll 1
    for (int q = 0; q < nlist; ++q) {
        const int64_t e = s_list[q];
        int a, b;
        if (e < 0) {
            a = 0;
            b = in_len;
        } else {
            const int64_t b64 = __ldg(offsets + e + 1);
            a = static_cast<int>(__ldg(offsets + e) - r.lo);
            b = static_cast<int>((b64 < r.hi ? b64 : r.hi) - r.lo);
        }
        if (b - a >= kElemCtaSeg) {
            Partial<A> p = shuffle_merge<OP, T, A>(smem_strided_partial<OP, T, A>(sm, a, b, tid, kElemRedThreads, r.lo));
            __syncthreads();
            if (lane == 0) s_part[warp] = p;
            __syncthreads();
            if (warp == 0) {
                Partial<A> w;
                if (lane < kElemRedWarps) w = s_part[lane];
                else { w.has = false; w.val = R::identity(); w.pos = -1; }
                w = shuffle_merge<OP, T, A>(w);
                if (lane == 0) {
                    const A v = w.has ? w.val : R::identity();
                    if (e < 0) carry[blockIdx.x] = v;
                    else out[e] = v;
                }
            }
        } else if ((q & (kElemRedWarps - 1)) == warp) {
            Partial<A> p = shuffle_merge<OP, T, A>(smem_strided_partial<OP, T, A>(sm, a, b, lane, kWarp, r.lo));
            if (lane == 0) {
                const A v = p.has ? p.val : R::identity();
                if (e < 0) carry[blockIdx.x] = v;
                else out[e] = v;
            }
        }
    }
}

// events that span several tiles: out[e] holds the part inside the tile where e starts, carry[t] the parts inside the
// following tiles, merged in tile order.  One THREAD per tile whose last owned event leaves it (the usual case at mean
// counts of 10 - 100: the event ends in the next tile or the one after: a warp per tile ran four dependent loads for
// one useful lane, 52 us of a 0.67 ms argmax on Poisson(40)); a span of more than kFixSerial tiles is handed to the
// whole warp (lanes stride the tiles, shuffle merge).
constexpr int kFixSerial = 8;

template <int OP, typename T, typename A>
__global__ void __launch_bounds__(kElemThreads)
segreduce_elem_fix_kernel(const int64_t *__restrict__ offsets, int64_t n, const int64_t *__restrict__ first_event,
                          int64_t ntiles, A *__restrict__ out, const A *__restrict__ carry) {
    constexpr int ET = ElemTile<T>::ET;
    using R = Red<OP, T, A>;
    const int lane = threadIdx.x & 31;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * kElemThreads + threadIdx.x;
    int64_t e = -1, t_end = -1;
    if (c < ntiles) {
        const int64_t ev_lo = __ldg(first_event + c), ev_hi = __ldg(first_event + c + 1);
        if (ev_hi > ev_lo) {
            const int64_t base = __ldg(offsets);
            const int64_t stop = __ldg(offsets + ev_hi);
            if (stop > base + (c + 1) * ET) {             // the last owned event leaves its tile
                e = ev_hi - 1;
                t_end = (stop - 1 - base) / ET;
            }
        }
    }
    const bool spans = e >= 0;
    if (spans && t_end - c <= kFixSerial) {
        Partial<A> p;
        p.has = true;
        p.val = out[e];
        p.pos = c;
        for (int64_t t = c + 1; t <= t_end; ++t) {
            Partial<A> q;
            q.has = true;
            q.val = carry[t];
            q.pos = t;
            p = merge<OP, T, A>(p, q);
        }
        out[e] = p.val;
    }
    unsigned todo = __ballot_sync(0xffffffffu, spans && t_end - c > kFixSerial);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t wc = __shfl_sync(0xffffffffu, c, src), we = __shfl_sync(0xffffffffu, e, src);
        const int64_t wt = __shfl_sync(0xffffffffu, t_end, src);
        Partial<A> p;
        p.has = false;
        p.val = R::identity();
        p.pos = -1;
        for (int64_t t = wc + 1 + lane; t <= wt; t += kWarp) {
            Partial<A> q;
            q.has = true;
            q.val = carry[t];
            q.pos = t;
            p = merge<OP, T, A>(p, q);
        }
        p = shuffle_merge<OP, T, A>(p);
        if (lane == 0) {
            Partial<A> first;
            first.has = true;
            first.val = out[we];
            first.pos = wc;
            out[we] = merge<OP, T, A>(first, p).val;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// argmin / argmax: best[e] = local index of the first extreme non-NaN element (-1: none), best_value[e] its value
// ------------------------------------------------------------------------------------------------------------
template <bool ISMIN, typename T>
__device__ __forceinline__ ArgPartial<T> smem_arg_strided(const T *__restrict__ sm, int a, int b, int first, int step,
                                                          int64_t pos0) {
    // (32-bit tile-relative index in the loop; the running best starts at the identity so that "strictly better" needs no
    // first-element test; a segment whose non-NaN values all equal the identity takes the exact loop)
    T best = ISMIN ? min_identity<T>() : max_identity<T>();
    int jbest = -1;
#pragma unroll 2
    for (int j = a + first; j < b; j += step) {
        const T x = sm[j];
        if (ISMIN ? (x < best) : (x > best)) {
            best = x;
            jbest = j;
        }
    }
    if (jbest < 0) {
#pragma unroll 1
        for (int j = a + first; j < b; j += step)
            if (!is_nan(sm[j])) {
                jbest = j;
                best = sm[j];
                break;
            }
    }
    ArgPartial<T> p;
    p.val = jbest < 0 ? T(0) : best;
    p.idx = jbest < 0 ? -1 : pos0 + jbest;          // absolute content position
    return p;
}

template <bool ISMIN, typename T>
__global__ void __launch_bounds__(kElemRedThreads)
segarg_elem_kernel(const T *__restrict__ content, const int64_t *__restrict__ offsets, int64_t n,
                   const int64_t *__restrict__ first_event, int64_t *__restrict__ best, T *__restrict__ best_value,
                   int64_t *__restrict__ carry_idx, T *__restrict__ carry_val) {
    constexpr int ET = ElemTile<T>::ET;
    __shared__ __align__(16) unsigned char s_stage[ET * sizeof(T) + 32];
    __shared__ uint64_t s_bar;
    __shared__ int64_t s_list[ET / kElemWarpSeg + 2];
    __shared__ int s_nlist;
    __shared__ ArgPartial<T> s_part[kElemRedWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ElemRange r = elem_range(offsets, n, first_event, blockIdx.x, ET);
    const int L = static_cast<int>(r.hi - r.lo);
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        s_nlist = 0;
        if (L > 0) stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + r.lo), static_cast<uint32_t>(L * sizeof(T)), &s_bar);
    }
    __syncthreads();
    const T *sm = reinterpret_cast<const T *>(s_stage);
    // (offsets of the thread's first owned event before the wait for the tile, of its next one while the current one is
    // reduced: see segreduce_elem_kernel)
    int64_t e = r.ev_lo + tid;
    int64_t a64 = 0, b64 = 0;
    if (e < r.ev_hi) {
        a64 = __ldg(offsets + e);
        b64 = __ldg(offsets + e + 1);
    }
    if (L > 0) {
        const uint32_t shift = stage_edges<sizeof(T)>(s_stage, reinterpret_cast<const unsigned char *>(content + r.lo),
                                                      static_cast<uint32_t>(L * sizeof(T)));
        sm = reinterpret_cast<const T *>(s_stage + shift);
        stage_wait(&s_bar, 0);
    }
    // a result is (absolute position, value); the event's start turns it into a local index
    auto emit_at = [&](int64_t ev, int64_t start, ArgPartial<T> p) {
        if (ev < 0) {
            carry_idx[blockIdx.x] = p.idx;
            carry_val[blockIdx.x] = p.val;
        } else {
            best[ev] = p.idx < 0 ? -1 : p.idx - start;
            best_value[ev] = p.val;
        }
    };
    auto emit = [&](int64_t ev, ArgPartial<T> p) { emit_at(ev, ev < 0 ? 0 : __ldg(offsets + ev), p); };
    const int in_len = static_cast<int>(r.in_end - r.lo);
    if (in_len >= kElemWarpSeg) {
        if (tid == 0) s_list[atomicAdd(&s_nlist, 1)] = -1;
    } else if (tid == 0) {
        emit(-1, smem_arg_strided<ISMIN, T>(sm, 0, in_len, 0, 1, r.lo));
    }
#pragma unroll 1
    while (e < r.ev_hi) {
        const int64_t en = e + kElemRedThreads;
        int64_t na = 0, nb = 0;
        if (en < r.ev_hi) {
            na = __ldg(offsets + en);
            nb = __ldg(offsets + en + 1);
        }
        const int a = static_cast<int>(a64 - r.lo);
        const int b = static_cast<int>((b64 < r.hi ? b64 : r.hi) - r.lo);
        if (b - a >= kElemWarpSeg) s_list[atomicAdd(&s_nlist, 1)] = e;
        else emit_at(e, a64, smem_arg_strided<ISMIN, T>(sm, a, b, 0, 1, r.lo));
        e = en;
        a64 = na;
        b64 = nb;
    }
    __syncthreads();
    const int nlist = s_nlist;
#pragma unroll 1
    for (int q = 0; q < nlist; ++q) {
        const int64_t e = s_list[q];
        int a, b;
        if (e < 0) {
            a = 0;
            b = in_len;
        } else {
            const int64_t b64 = __ldg(offsets + e + 1);
            a = static_cast<int>(__ldg(offsets + e) - r.lo);
            b = static_cast<int>((b64 < r.hi ? b64 : r.hi) - r.lo);
        }
        if (b - a >= kElemCtaSeg) {
            ArgPartial<T> p = arg_shuffle<ISMIN, T>(smem_arg_strided<ISMIN, T>(sm, a, b, tid, kElemRedThreads, r.lo));
            __syncthreads();
            if (lane == 0) s_part[warp] = p;
            __syncthreads();
            if (warp == 0) {
                ArgPartial<T> w;
                if (lane < kElemRedWarps) w = s_part[lane];
                else { w.idx = -1; w.val = T(0); }
                w = arg_shuffle<ISMIN, T>(w);
                if (lane == 0) emit(e, w);
            }
        } else if ((q & (kElemRedWarps - 1)) == warp) {
            ArgPartial<T> p = arg_shuffle<ISMIN, T>(smem_arg_strided<ISMIN, T>(sm, a, b, lane, kWarp, r.lo));
            if (lane == 0) emit(e, p);
        }
    }
}

template <bool ISMIN, typename T>
__global__ void __launch_bounds__(kElemThreads)
segarg_elem_fix_kernel(const int64_t *__restrict__ offsets, int64_t n, const int64_t *__restrict__ first_event,
                       int64_t ntiles, int64_t *__restrict__ best, T *__restrict__ best_value,
                       const int64_t *__restrict__ carry_idx, const T *__restrict__ carry_val) {
    // (one thread per tile, long spans by the whole warp: see segreduce_elem_fix_kernel)
    constexpr int ET = ElemTile<T>::ET;
    const int lane = threadIdx.x & 31;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * kElemThreads + threadIdx.x;
    int64_t e = -1, t_end = -1, start = 0;
    if (c < ntiles) {
        const int64_t ev_lo = __ldg(first_event + c), ev_hi = __ldg(first_event + c + 1);
        if (ev_hi > ev_lo) {
            const int64_t base = __ldg(offsets);
            const int64_t stop = __ldg(offsets + ev_hi);
            if (stop > base + (c + 1) * ET) {
                e = ev_hi - 1;
                start = __ldg(offsets + e);
                t_end = (stop - 1 - base) / ET;
            }
        }
    }
    const bool spans = e >= 0;
    if (spans && t_end - c <= kFixSerial) {
        ArgPartial<T> p;
        const int64_t local = best[e];
        p.idx = local < 0 ? -1 : local + start;
        p.val = best_value[e];
        for (int64_t t = c + 1; t <= t_end; ++t) {
            ArgPartial<T> q;
            q.idx = carry_idx[t];
            q.val = carry_val[t];
            p = arg_merge<ISMIN, T>(p, q);
        }
        best[e] = p.idx < 0 ? -1 : p.idx - start;
        best_value[e] = p.val;
    }
    unsigned todo = __ballot_sync(0xffffffffu, spans && t_end - c > kFixSerial);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t wc = __shfl_sync(0xffffffffu, c, src), we = __shfl_sync(0xffffffffu, e, src);
        const int64_t wt = __shfl_sync(0xffffffffu, t_end, src), ws = __shfl_sync(0xffffffffu, start, src);
        ArgPartial<T> p;
        p.idx = -1;
        p.val = T(0);
        for (int64_t t = wc + 1 + lane; t <= wt; t += kWarp) {
            ArgPartial<T> q;
            q.idx = carry_idx[t];
            q.val = carry_val[t];
            p = arg_merge<ISMIN, T>(p, q);
        }
        p = arg_shuffle<ISMIN, T>(p);
        if (lane == 0) {
            ArgPartial<T> first;
            const int64_t local = best[we];
            first.idx = local < 0 ? -1 : local + ws;
            first.val = best_value[we];
            const ArgPartial<T> w = arg_merge<ISMIN, T>(first, p);
            best[we] = w.idx < 0 ? -1 : w.idx - ws;
            best_value[we] = w.val;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// parents / localindex / broadcast
// ------------------------------------------------------------------------------------------------------------
enum { ELEM_PARENTS = 0, ELEM_LOCALINDEX = 1, ELEM_BROADCAST = 2 };

template <int MODE, typename V>
__global__ void __launch_bounds__(kElemThreads)
expand_elem_kernel(const int64_t *__restrict__ offsets, int64_t n, const int64_t *__restrict__ first_event,
                   V *__restrict__ out, const V *__restrict__ flat) {
    constexpr int ET = 4096;
    constexpr int PER = ET / kElemThreads;
    __shared__ __align__(16) int32_t s_own[ET];
    __shared__ int32_t s_warpmax[kElemWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ElemRange r = elem_range(offsets, n, first_event, blockIdx.x, ET);
    const int L = static_cast<int>(r.hi - r.lo);
    const int64_t o0 = __ldg(offsets);
    if (MODE == ELEM_PARENTS && blockIdx.x == 0)
        for (int64_t j = tid; j < o0; j += kElemThreads) out[j] = static_cast<V>(-1);      // unreachable content (jagged.py:52-54)
    // relative event numbers: 0 is the event that enters from the left (or the first owned one)
    const int64_t ev_first = (r.in_end > r.lo) ? r.ev_lo - 1 : r.ev_lo;
#pragma unroll
    for (int k = 0; k < PER; ++k) s_own[tid + k * kElemThreads] = 0;
    __syncthreads();
#pragma unroll 1
    for (int64_t e = r.ev_lo + tid; e < r.ev_hi; e += kElemThreads) {
        const int64_t a = __ldg(offsets + e), b = __ldg(offsets + e + 1);
        if (b > a) s_own[a - r.lo] = static_cast<int32_t>(e - ev_first);      // non-empty events start at distinct positions
    }
    __syncthreads();
    // running maximum.  Warp w owns the 512 positions [512 w, 512 (w + 1)) as 4 rows of 128 (a lane reads 4 consecutive
    // entries with one conflict-free 16-byte access): first the maximum of every warp's chunk, then the rows with the
    // carry of everything before them
    constexpr int kChunk = ET / kElemWarps, kRows = kChunk / 128;
    int4 rows[kRows];
    int32_t cm = 0;
#pragma unroll
    for (int q = 0; q < kRows; ++q) {
        rows[q] = *reinterpret_cast<const int4 *>(&s_own[warp * kChunk + q * 128 + lane * 4]);
        cm = max(max(cm, max(rows[q].x, rows[q].y)), max(rows[q].z, rows[q].w));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, d));
    if (lane == 0) s_warpmax[warp] = cm;
    __syncthreads();
    int32_t carry = 0;
#pragma unroll
    for (int w = 0; w < kElemWarps; ++w)
        if (w < warp) carry = max(carry, s_warpmax[w]);
#pragma unroll
    for (int q = 0; q < kRows; ++q) {
        int32_t v0 = rows[q].x, v1 = max(rows[q].y, v0), v2 = max(rows[q].z, v1), v3 = max(rows[q].w, v2);
        int32_t incl = v3;
#pragma unroll
        for (int d = 1; d < kWarp; d <<= 1) {
            const int32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl = max(incl, o);
        }
        int32_t before = __shfl_up_sync(0xffffffffu, incl, 1);
        before = lane == 0 ? carry : max(before, carry);
        *reinterpret_cast<int4 *>(&s_own[warp * kChunk + q * 128 + lane * 4]) =
            make_int4(max(v0, before), max(v1, before), max(v2, before), max(v3, before));
        carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
    }
    __syncthreads();
    V *dst = (MODE == ELEM_PARENTS) ? out + r.lo : out + (r.lo - o0);
#pragma unroll 4
    for (int p = tid; p < L; p += kElemThreads) {
        const int64_t ev = ev_first + s_own[p];
        if (MODE == ELEM_PARENTS) __stcs(dst + p, static_cast<V>(ev));
        else if (MODE == ELEM_LOCALINDEX) __stcs(dst + p, static_cast<V>(r.lo + p - __ldg(offsets + ev)));
        else __stcs(dst + p, __ldg(flat + ev));
    }
}

// ------------------------------------------------------------------------------------------------------------
// selection: flat compaction of element tiles on the lagged prefixes + rank at the event boundaries
// ------------------------------------------------------------------------------------------------------------
// (the selectors and the two passes over a staged run are shared with the event-tiled kernels: jg_compact.cuh)
using ElemByteMask = ByteMask;
template <typename K, bool STRICT>
using ElemCompare = FloatCompare<K, STRICT>;

template <class SEL>
__device__ __forceinline__ int elem_count16(const SEL &sel, int4 v) {
    using E = typename SEL::Elem;
    if constexpr (sizeof(E) == 1) {
        return __popc(__vcmpne4(v.x, 0) & 0x01010101u) + __popc(__vcmpne4(v.y, 0) & 0x01010101u) +
               __popc(__vcmpne4(v.z, 0) & 0x01010101u) + __popc(__vcmpne4(v.w, 0) & 0x01010101u);
    } else {
        union { int4 raw; E e[16 / sizeof(E)]; } q;
        q.raw = v;
        int c = 0;
#pragma unroll
        for (int k = 0; k < static_cast<int>(16 / sizeof(E)); ++k) c += sel.test(q.e[k]);
        return c;
    }
}

// count: one warp per tile (16-byte loads, unaligned head and tail item by item) -> tile_totals
template <class SEL>
__global__ void __launch_bounds__(kElemThreads)
select_elem_count_kernel(const SEL sel, const typename SEL::Elem *__restrict__ mask, int64_t mask_shift,
                         const int64_t *__restrict__ offsets, int64_t n, int64_t et, int64_t ntiles,
                         int64_t *__restrict__ tile_totals) {
    using E = typename SEL::Elem;
    constexpr int PERV = 16 / static_cast<int>(sizeof(E));
    const int lane = threadIdx.x & 31;
    const int64_t t = static_cast<int64_t>(blockIdx.x) * kElemWarps + (threadIdx.x >> 5);
    if (t >= ntiles) return;
    const int64_t o0 = __ldg(offsets), end = __ldg(offsets + n);
    const int64_t lo = o0 + t * et;
    const int64_t hi = lo + et < end ? lo + et : end;
    const int L = hi > lo ? static_cast<int>(hi - lo) : 0;
    const E *m = mask + lo + mask_shift;
    int cnt = 0;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(m);
    int head = static_cast<int>(((16 - (addr & 15)) & 15) / sizeof(E));
    if (head > L) head = L;
    if (lane < head) cnt += sel.test(m[lane]);
    const int nvec = (L - head) / PERV;
    const int4 *mv = reinterpret_cast<const int4 *>(m + head);
#pragma unroll 4
    for (int i = lane; i < nvec; i += kWarp) cnt += elem_count16(sel, __ldcs(mv + i));
    const int tail0 = head + nvec * PERV;
    if (tail0 + lane < L) cnt += sel.test(m[tail0 + lane]);
    cnt = warp_reduce_add(cnt);
    if (lane == 0) tile_totals[t] = cnt;
}

// fill: the tile's run staged by TMA, element-parallel ballot / popc compaction (a warp's stores are consecutive), then
// every owned event reads its new offset (and, when it ends inside the tile, its count) from the per-position prefix
template <typename V, class SEL, bool SAME>
__global__ void __launch_bounds__(kElemThreads, sizeof(V) == 8 ? 6 : 8)
select_elem_fill_kernel(const SEL sel, const V *__restrict__ content, const typename SEL::Elem *__restrict__ mask,
                        int64_t mask_shift, const int64_t *__restrict__ offsets, int64_t n,
                        const int64_t *__restrict__ first_event, V *__restrict__ out, int64_t *__restrict__ new_offsets,
                        int64_t *__restrict__ new_counts, const int64_t *__restrict__ tile_bases, int64_t ntiles) {
    using E = typename SEL::Elem;
    constexpr int ET = ElemTile<V>::ET;
    // (+ one row of padding behind the selector's run: the rows of the compaction are whole)
    __shared__ __align__(16) unsigned char s_content[ET * sizeof(V) + 32 + (SAME ? 32 * sizeof(V) : 0)];
    __shared__ __align__(16) unsigned char s_mask[SAME ? 16 : ET * sizeof(E) + 32 + 32 * sizeof(E)];
    __shared__ uint16_t s_pfx[ET + 2];
    __shared__ uint64_t s_bar[2];
    __shared__ int32_t s_scan[kElemWarps + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tile = blockIdx.x;
    const int64_t o0 = __ldg(offsets), end = __ldg(offsets + n);
    const int64_t lo = o0 + tile * ET;
    const int64_t hi = lo + ET < end ? lo + ET : end;
    const int L = hi > lo ? static_cast<int>(hi - lo) : 0;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        if (L > 0) {
            if (!SAME) stage_bulk(s_mask, reinterpret_cast<const unsigned char *>(mask + lo + mask_shift), static_cast<uint32_t>(L * sizeof(E)), &s_bar[1]);
            stage_bulk(s_content, reinterpret_cast<const unsigned char *>(content + lo), static_cast<uint32_t>(L * sizeof(V)), &s_bar[0]);
        }
    }
    const int64_t ev_lo = __ldg(first_event + tile), ev_hi = __ldg(first_event + tile + 1);
    const int64_t base = __ldg(tile_bases + tile);
    if (tile == ntiles - 1 && tid == 0) new_offsets[n] = __ldg(tile_bases + ntiles);
    // the offsets of the thread's first owned event are asked for now, while the tile is on its way (they are needed
    // after the compaction: a dependent round trip at the end of the CTA's life otherwise, holding its shared memory)
    int64_t ea = 0, eb = 0;
    if (ev_lo + tid < ev_hi) {
        ea = __ldg(offsets + ev_lo + tid);
        eb = __ldg(offsets + ev_lo + tid + 1);
    }
    __syncthreads();
    const V *cv = nullptr;
    const E *mv = nullptr;
    if (L > 0) {
        const uint32_t cs = stage_edges<sizeof(V)>(s_content, reinterpret_cast<const unsigned char *>(content + lo), static_cast<uint32_t>(L * sizeof(V)));
        cv = reinterpret_cast<const V *>(s_content + cs);
        E *pad;
        if (SAME) {
            pad = reinterpret_cast<E *>(s_content + cs);
        } else {
            const uint32_t ms = stage_edges<sizeof(E)>(s_mask, reinterpret_cast<const unsigned char *>(mask + lo + mask_shift), static_cast<uint32_t>(L * sizeof(E)));
            pad = reinterpret_cast<E *>(s_mask + ms);
        }
        mv = pad;
        if (warp == 1) rows_pad<SEL>(pad, L, lane);
        mbar_wait(&s_bar[SAME ? 0 : 1], 0);
    }
    __syncthreads();
    // rows of 32 positions (jg_compact.cuh): warp w owns rows [w * rows_per_warp, (w + 1) * rows_per_warp)
    const int rows_total = (L + 31) >> 5;
    const int rows_per_warp = (rows_total + kElemWarps - 1) / kElemWarps;
    const int r_begin = min(rows_total, warp * rows_per_warp);
    const int r_end = min(rows_total, r_begin + rows_per_warp);
    const int cnt = rows_count(sel, mv, r_begin, r_end, lane);
    if (lane == 0) s_scan[warp] = cnt;
    if (!SAME && L > 0) mbar_wait(&s_bar[0], 0);
    __syncthreads();
    uint32_t run = __reduce_add_sync(0xffffffffu, lane < warp ? s_scan[lane] : 0);
    run = rows_compact<V, SEL, SAME>(sel, mv, cv, s_pfx, r_begin, r_end, lane, run, out + base);
    if (warp == kElemWarps - 1 && lane == 0) s_pfx[rows_total << 5] = static_cast<uint16_t>(run);
    __syncthreads();
#pragma unroll 1
    for (int64_t e = ev_lo + tid; e < ev_hi; e += kElemThreads) {
        const int64_t en = e + kElemThreads;
        int64_t na = 0, nb = 0;
        if (en < ev_hi) {
            na = __ldg(offsets + en);
            nb = __ldg(offsets + en + 1);
        }
        const int64_t a = ea - lo, b = eb - lo;
        const int first = s_pfx[a];
        new_offsets[e] = base + first;
        if (new_counts && b <= L) new_counts[e] = static_cast<int>(s_pfx[b]) - first;
        ea = na;
        eb = nb;
    }
}

// counts of the events that leave their tile (at most one per tile): new_offsets is complete after the main kernel
__global__ void __launch_bounds__(256)
select_elem_fix_kernel(const int64_t *__restrict__ offsets, const int64_t *__restrict__ first_event, int64_t ntiles,
                       int64_t et, const int64_t *__restrict__ new_offsets, int64_t *__restrict__ new_counts) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= ntiles) return;
    const int64_t ev_lo = __ldg(first_event + c), ev_hi = __ldg(first_event + c + 1);
    if (ev_hi <= ev_lo) return;
    const int64_t e = ev_hi - 1;
    if (__ldg(offsets + e + 1) > __ldg(offsets) + (c + 1) * et) new_counts[e] = new_offsets[e + 1] - new_offsets[e];
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct ElemPlan {
    int64_t ntiles;
    int64_t et;
    int64_t *first_event;      // ntiles + 1
    char *rest;                // what follows in the workspace
    size_t rest_bytes;
};

static inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static int elem_plan(const int64_t *offsets, int64_t n, int64_t content_len, int itemsize, void *ws, size_t ws_bytes,
                     ElemPlan &p, cudaStream_t s, const char *who) {
    p.et = elem_tile_items(itemsize);
    p.ntiles = content_len <= 0 ? 1 : ceil_div(content_len, p.et);
    const size_t part = align256(static_cast<size_t>(p.ntiles + 1) * 8);
    if (ws == nullptr || ws_bytes < part) {
        set_error("%s: workspace too small (%zu < %zu)", who, ws_bytes, part);
        return -2;
    }
    p.first_event = static_cast<int64_t *>(ws);
    p.rest = static_cast<char *>(ws) + part;
    p.rest_bytes = ws_bytes - part;
    elem_partition_kernel<<<static_cast<unsigned>(ceil_div(p.ntiles + 1, 256)), 256, 0, s>>>(offsets, n, p.ntiles, p.et, p.first_event);
    return check_launch("elem_partition_kernel");
}

template <int OP, typename T, typename A>
static int launch_reduce_elem(const void *content, const int64_t *offsets, int64_t n, void *out, const ElemPlan &p, cudaStream_t s) {
    const size_t need = static_cast<size_t>(p.ntiles) * sizeof(A);
    JG_REQUIRE(p.rest_bytes >= need, "jg_segreduce_elem: workspace too small");
    A *carry = reinterpret_cast<A *>(p.rest);
    segreduce_elem_kernel<OP, T, A><<<static_cast<unsigned>(p.ntiles), kElemRedThreads, 0, s>>>(
        static_cast<const T *>(content), offsets, n, p.first_event, static_cast<A *>(out), carry);
    int rc = check_launch("segreduce_elem_kernel");
    if (rc) return rc;
    segreduce_elem_fix_kernel<OP, T, A><<<static_cast<unsigned>(ceil_div(p.ntiles, kElemThreads)), kElemThreads, 0, s>>>(
        offsets, n, p.first_event, p.ntiles, static_cast<A *>(out), carry);
    return check_launch("segreduce_elem_fix_kernel");
}

template <typename T, typename ASUM>
static int dispatch_reduce_elem(int op, const void *content, const int64_t *offsets, int64_t n, void *out, const ElemPlan &p, cudaStream_t s) {
    switch (op) {
        case JG_SUM: return launch_reduce_elem<JG_SUM, T, ASUM>(content, offsets, n, out, p, s);
        case JG_PROD: return launch_reduce_elem<JG_PROD, T, ASUM>(content, offsets, n, out, p, s);
        case JG_MIN: return launch_reduce_elem<JG_MIN, T, T>(content, offsets, n, out, p, s);
        case JG_MAX: return launch_reduce_elem<JG_MAX, T, T>(content, offsets, n, out, p, s);
        case JG_ANY: return launch_reduce_elem<JG_ANY, T, uint8_t>(content, offsets, n, out, p, s);
        case JG_ALL: return launch_reduce_elem<JG_ALL, T, uint8_t>(content, offsets, n, out, p, s);
    }
    set_error("jg_segreduce_elem: unsupported op %d", op);
    return -2;
}

template <typename T>
static int dispatch_arg_elem(int is_min, const void *content, const int64_t *offsets, int64_t n, int64_t *best, void *best_value,
                             const ElemPlan &p, cudaStream_t s) {
    const size_t idx_bytes = align256(static_cast<size_t>(p.ntiles) * 8);
    JG_REQUIRE(p.rest_bytes >= idx_bytes + static_cast<size_t>(p.ntiles) * sizeof(T), "jg_segargminmax_elem: workspace too small");
    int64_t *carry_idx = reinterpret_cast<int64_t *>(p.rest);
    T *carry_val = reinterpret_cast<T *>(p.rest + idx_bytes);
    const unsigned g = static_cast<unsigned>(p.ntiles), gf = static_cast<unsigned>(ceil_div(p.ntiles, kElemThreads));
    if (is_min) {
        segarg_elem_kernel<true, T><<<g, kElemRedThreads, 0, s>>>(static_cast<const T *>(content), offsets, n, p.first_event, best,
                                                              static_cast<T *>(best_value), carry_idx, carry_val);
        segarg_elem_fix_kernel<true, T><<<gf, kElemThreads, 0, s>>>(offsets, n, p.first_event, p.ntiles, best,
                                                                   static_cast<T *>(best_value), carry_idx, carry_val);
    } else {
        segarg_elem_kernel<false, T><<<g, kElemRedThreads, 0, s>>>(static_cast<const T *>(content), offsets, n, p.first_event, best,
                                                               static_cast<T *>(best_value), carry_idx, carry_val);
        segarg_elem_fix_kernel<false, T><<<gf, kElemThreads, 0, s>>>(offsets, n, p.first_event, p.ntiles, best,
                                                                    static_cast<T *>(best_value), carry_idx, carry_val);
    }
    return check_launch("segarg_elem_kernel");
}

template <typename V, class SEL, bool SAME>
static int launch_select_elem(const SEL &sel, const void *content, const typename SEL::Elem *mask, int64_t mask_shift,
                              const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                              int64_t *total, const ElemPlan &p, cudaStream_t s) {
    // three launches, none of which waits on another CTA (see jg_select.cu): tile totals, their scan, the compaction
    const size_t arr = align256(static_cast<size_t>(p.ntiles + 2) * 8);
    const size_t scan_need = lag_ws_bytes(ceil_div(p.ntiles, kScanTmaThreads * kScanTmaItems) + 1) + sizeof(LookbackWs) +
                             static_cast<size_t>(ceil_div(p.ntiles, kScanTmaThreads * kScanTmaItems) + 1) * 8;
    JG_REQUIRE(p.rest_bytes >= 2 * arr + scan_need + 256, "jg_select_elem: workspace too small (%zu < %zu)", p.rest_bytes,
               2 * arr + scan_need + 256);
    int64_t *totals = reinterpret_cast<int64_t *>(p.rest);
    int64_t *bases = reinterpret_cast<int64_t *>(p.rest + arr);
    select_elem_count_kernel<SEL><<<static_cast<unsigned>(ceil_div(p.ntiles, kElemWarps)), kElemThreads, 0, s>>>(
        sel, mask, mask_shift, offsets, n, p.et, p.ntiles, totals);
    int rc = check_launch("select_elem_count_kernel");
    if (rc) return rc;
    rc = launch_scan_array(totals, p.ntiles, bases, total, nullptr, p.rest + 2 * arr, p.rest_bytes - 2 * arr, s);
    if (rc) return rc;
    select_elem_fill_kernel<V, SEL, SAME><<<static_cast<unsigned>(p.ntiles), kElemThreads, 0, s>>>(
        sel, static_cast<const V *>(content), mask, mask_shift, offsets, n, p.first_event, static_cast<V *>(out), new_offsets,
        new_counts, bases, p.ntiles);
    rc = check_launch("select_elem_fill_kernel");
    if (rc || new_counts == nullptr) return rc;
    select_elem_fix_kernel<<<static_cast<unsigned>(ceil_div(p.ntiles, 256)), 256, 0, s>>>(offsets, p.first_event, p.ntiles, p.et,
                                                                                        new_offsets, new_counts);
    return check_launch("select_elem_fix_kernel");
}

template <typename K, bool STRICT>
static int compare_select_elem(const void *content, int itemsize, const void *key, int op, double scalar, int64_t key_shift,
                               const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts,
                               int64_t *total, const ElemPlan &p, cudaStream_t s) {
    using SEL = ElemCompare<K, STRICT>;
    SEL sel;
    const K c = static_cast<K>(scalar);
    const bool flip = (op == JG_OP_LT || op == JG_OP_LE);
    sel.flip = flip ? (static_cast<typename SEL::Bits>(1) << (8 * sizeof(K) - 1)) : 0;
    sel.c = flip ? -c : c;
    const K *k = static_cast<const K *>(key);
    using U = typename std::conditional<sizeof(K) == 4, uint32_t, uint64_t>::type;
    if (key == content && key_shift == 0 && itemsize == static_cast<int>(sizeof(K)))
        return launch_select_elem<U, SEL, true>(sel, content, k, 0, offsets, n, out, new_offsets, new_counts, total, p, s);
    JG_REQUIRE(itemsize == static_cast<int>(sizeof(K)), "jg_compare_select_elem: the selected column must have the key's item size");
    return launch_select_elem<U, SEL, false>(sel, content, k, key_shift, offsets, n, out, new_offsets, new_counts, total, p, s);
}

}  // namespace jg

using namespace jg;

extern "C" {

// scratch for the element-tiled calls on `content_len` items of `itemsize` bytes (jg_workspace_bytes is sized by events)
size_t jg_elem_workspace_bytes(int64_t content_len, int itemsize) {
    if (content_len < 0) content_len = 0;
    if (itemsize < 1) itemsize = 1;
    const size_t tiles = static_cast<size_t>(content_len / elem_tile_items(itemsize < 4 ? 4 : itemsize) + 2);
    // first_event + (carry index + carry value | tile totals + tile bases + the scratch of their scan)
    return align256((tiles + 1) * 8) + 2 * align256((tiles + 2) * 8) + lag_ws_bytes(static_cast<int64_t>(tiles)) + 4096;
}

int jg_segreduce_elem(const void *content, int dtype, const int64_t *offsets, int64_t n, int op, void *out,
                      int64_t content_len, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && out, "jg_segreduce_elem: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    ElemPlan p;
    int rc = elem_plan(offsets, n, content_len, dtype_size(dtype), ws, ws_bytes, p, s, "jg_segreduce_elem");
    if (rc) return rc;
    switch (dtype) {
        case JG_F32: return dispatch_reduce_elem<float, float>(op, content, offsets, n, out, p, s);
        case JG_F64: return dispatch_reduce_elem<double, double>(op, content, offsets, n, out, p, s);
        case JG_I32: return dispatch_reduce_elem<int32_t, int64_t>(op, content, offsets, n, out, p, s);
        case JG_I64: return dispatch_reduce_elem<int64_t, int64_t>(op, content, offsets, n, out, p, s);
    }
    set_error("jg_segreduce_elem: dtype %d takes the event-tiled path", dtype);
    return -2;
}

int jg_segargminmax_elem(const void *content, int dtype, const int64_t *offsets, int64_t n, int is_min, int64_t *best,
                         void *best_value, int64_t content_len, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && best && best_value, "jg_segargminmax_elem: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    ElemPlan p;
    int rc = elem_plan(offsets, n, content_len, dtype_size(dtype), ws, ws_bytes, p, s, "jg_segargminmax_elem");
    if (rc) return rc;
    switch (dtype) {
        case JG_F32: return dispatch_arg_elem<float>(is_min, content, offsets, n, best, best_value, p, s);
        case JG_F64: return dispatch_arg_elem<double>(is_min, content, offsets, n, best, best_value, p, s);
        case JG_I32: return dispatch_arg_elem<int32_t>(is_min, content, offsets, n, best, best_value, p, s);
        case JG_I64: return dispatch_arg_elem<int64_t>(is_min, content, offsets, n, best, best_value, p, s);
    }
    set_error("jg_segargminmax_elem: dtype %d takes the event-tiled path", dtype);
    return -2;
}

int jg_mask_select_elem(const void *content, int itemsize, const uint8_t *mask, int64_t mask_shift, const int64_t *offsets,
                        int64_t n, void *out, int64_t *new_offsets, int64_t *new_counts, int64_t *total,
                        int64_t content_len, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && new_offsets, "jg_mask_select_elem: bad arguments");
    JG_REQUIRE(itemsize == 4 || itemsize == 8, "jg_mask_select_elem: 4- or 8-byte items");
    cudaStream_t s = as_stream(stream);
    ElemPlan p;
    int rc = elem_plan(offsets, n, content_len, itemsize, ws, ws_bytes, p, s, "jg_mask_select_elem");
    if (rc) return rc;
    if (itemsize == 4)
        return launch_select_elem<uint32_t, ElemByteMask, false>(ElemByteMask{}, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, p, s);
    return launch_select_elem<uint64_t, ElemByteMask, false>(ElemByteMask{}, content, mask, mask_shift, offsets, n, out, new_offsets, new_counts, total, p, s);
}

int jg_compare_select_elem(const void *content, int itemsize, const void *key, int key_dtype, int op, double scalar,
                           int64_t key_shift, const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets,
                           int64_t *new_counts, int64_t *total, int64_t content_len, void *ws, size_t ws_bytes,
                           void *stream) {
    JG_REQUIRE(n >= 0 && offsets && new_offsets && key, "jg_compare_select_elem: bad arguments");
    JG_REQUIRE(op == JG_OP_GT || op == JG_OP_GE || op == JG_OP_LT || op == JG_OP_LE,
               "jg_compare_select_elem: operator %d is not an ordering comparison", op);
    cudaStream_t s = as_stream(stream);
    ElemPlan p;
    int rc = elem_plan(offsets, n, content_len, itemsize, ws, ws_bytes, p, s, "jg_compare_select_elem");
    if (rc) return rc;
    const bool strict = (op == JG_OP_GT || op == JG_OP_LT);
#define JG_CSE(K, ST) compare_select_elem<K, ST>(content, itemsize, key, op, scalar, key_shift, offsets, n, out, new_offsets, new_counts, total, p, s)
    if (key_dtype == JG_F32) return strict ? JG_CSE(float, true) : JG_CSE(float, false);
    if (key_dtype == JG_F64) return strict ? JG_CSE(double, true) : JG_CSE(double, false);
#undef JG_CSE
    set_error("jg_compare_select_elem: key dtype %d is not float32 / float64", key_dtype);
    return -2;
}

// mode: 0 parents (int64, absolute positions, -1 below offsets[0]), 1 localindex (int64), 2 broadcast of `flat`
int jg_expand_elem(int mode, const void *flat, int itemsize, const int64_t *offsets, int64_t n, void *out,
                   int64_t content_len, void *ws, size_t ws_bytes, void *stream) {
    if (n <= 0) return 0;
    JG_REQUIRE(offsets && out, "jg_expand_elem: bad arguments");
    cudaStream_t s = as_stream(stream);
    ElemPlan p;
    int rc = elem_plan(offsets, n, content_len, 4, ws, ws_bytes, p, s, "jg_expand_elem");      // 4096 positions per tile
    if (rc) return rc;
    const unsigned g = static_cast<unsigned>(p.ntiles);
    if (mode == ELEM_PARENTS)
        expand_elem_kernel<ELEM_PARENTS, int64_t><<<g, kElemThreads, 0, s>>>(offsets, n, p.first_event, static_cast<int64_t *>(out), nullptr);
    else if (mode == ELEM_LOCALINDEX)
        expand_elem_kernel<ELEM_LOCALINDEX, int64_t><<<g, kElemThreads, 0, s>>>(offsets, n, p.first_event, static_cast<int64_t *>(out), nullptr);
    else if (mode == ELEM_BROADCAST) {
        JG_REQUIRE(flat, "jg_expand_elem: broadcast needs the flat values");
        switch (itemsize) {
            case 1: expand_elem_kernel<ELEM_BROADCAST, uint8_t><<<g, kElemThreads, 0, s>>>(offsets, n, p.first_event, static_cast<uint8_t *>(out), static_cast<const uint8_t *>(flat)); break;
            case 2: expand_elem_kernel<ELEM_BROADCAST, uint16_t><<<g, kElemThreads, 0, s>>>(offsets, n, p.first_event, static_cast<uint16_t *>(out), static_cast<const uint16_t *>(flat)); break;
            case 4: expand_elem_kernel<ELEM_BROADCAST, uint32_t><<<g, kElemThreads, 0, s>>>(offsets, n, p.first_event, static_cast<uint32_t *>(out), static_cast<const uint32_t *>(flat)); break;
            case 8: expand_elem_kernel<ELEM_BROADCAST, uint64_t><<<g, kElemThreads, 0, s>>>(offsets, n, p.first_event, static_cast<uint64_t *>(out), static_cast<const uint64_t *>(flat)); break;
            default:
                set_error("jg_expand_elem: unsupported itemsize %d", itemsize);
                return -2;
        }
    } else {
        set_error("jg_expand_elem: unknown mode %d", mode);
        return -2;
    }
    return check_launch("expand_elem_kernel");
}

}  // extern "C"

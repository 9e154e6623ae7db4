// jg_tile.cuh -- the "event tile": a CTA owns kTileEvents consecutive events and keeps their offsets in
// shared memory.  Everything that walks content event by event (reductions, mask, gathers, broadcasts,
// combinatorics) is built on it, so content is always addressed through on-chip offsets.
#pragma once

#include "jg_common.cuh"

namespace jg {

constexpr int kTileThreads = 256;
constexpr int kTileEvents = 512;

struct EventTile {
    int64_t *off;      // shared: off[0..nev]
    int64_t event0;    // first event of this CTA
    int nev;           // number of events in this CTA (<= kTileEvents)

    // all threads of the CTA call this; `s_off` is a __shared__ int64_t[kTileEvents + 1]
    __device__ __forceinline__ void load(int64_t *s_off, const int64_t *__restrict__ offsets, int64_t n,
                                         int64_t tile_index) {
        load_nosync(s_off, offsets, n, tile_index);
        __syncthreads();
    }
    // The same without the barrier: a kernel that needs several offsets tiles (combinatorics: output, left, right;
    // a[jagged_int]: index and content layouts) issues all their loads back to back and synchronises ONCE -- the
    // tiles arrive in one memory round trip instead of one per tile (a CTA of these kernels lives a few
    // microseconds, most of it waiting for dependent loads).
    //   PF != 0: every thread also asks L2 for the line at `pf_base + offset * PF` of each offset it loaded (the
    //   start of each event's run in a PF-byte-item array): by the time the tile is on chip the run that the next,
    //   dependent stage reads is on its way -- the second round trip overlaps the first.
    template <int PF = 0>
    __device__ __forceinline__ void load_nosync(int64_t *s_off, const int64_t *__restrict__ offsets, int64_t n,
                                                int64_t tile_index, const void *pf_base = nullptr) {
        off = s_off;
        event0 = tile_index * kTileEvents;
        const int64_t left = n - event0;
        nev = left < kTileEvents ? static_cast<int>(left) : kTileEvents;
        // nev + 1 <= 513 entries, 256 threads: two predicated loads per thread + one straggler (a strided loop
        // here is unrolled by the compiler into 8/4/2/1 cascades that every warp walks through)
        const int i0 = threadIdx.x, i1 = threadIdx.x + kTileThreads;
        const int64_t *src = offsets + event0;
        int64_t v0 = 0, v1 = 0;
        if (i0 <= nev) v0 = __ldg(src + i0);
        if (i1 <= nev) v1 = __ldg(src + i1);
        if (i0 <= nev) s_off[i0] = v0;
        if (i1 <= nev) s_off[i1] = v1;
        if (threadIdx.x == 0 && nev == 2 * kTileThreads) s_off[nev] = __ldg(src + nev);
        if constexpr (PF != 0) {
            // only where an event is known to hold an item (its start is then a valid address): a lane sees its
            // neighbour's offset by shuffle, the last lane of a warp leaves its event to the neighbouring lines
            const char *b = static_cast<const char *>(pf_base);
            const int64_t n0 = __shfl_down_sync(0xffffffffu, v0, 1), n1 = __shfl_down_sync(0xffffffffu, v1, 1);
            const bool inner = (threadIdx.x & 31) != 31;
            if (inner && i0 < nev && n0 > v0) prefetch_l2(b + v0 * PF);
            if (inner && i1 < nev && n1 > v1) prefetch_l2(b + v1 * PF);
        }
    }
    // the tile's event range without touching memory (for the early bulk copy)
    __device__ static __forceinline__ void range(int64_t n, int64_t tile_index, int64_t &ev0, int &nev_) {
        ev0 = tile_index * kTileEvents;
        const int64_t left = n - ev0;
        nev_ = left < kTileEvents ? static_cast<int>(left) : kTileEvents;
    }
    __device__ __forceinline__ int64_t first_element() const { return off[0]; }
    __device__ __forceinline__ int64_t last_element() const { return off[nev]; }
    // event (local index) that contains element j, first_element() <= j < last_element()
    __device__ __forceinline__ int find_event(int64_t j) const {
        int lo = 0, hi = nev - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (off[mid + 1] <= j) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    }
};

static inline unsigned tile_grid(int64_t n) { return static_cast<unsigned>(n <= 0 ? 1 : ceil_div(n, kTileEvents)); }

}  // namespace jg

namespace jg {

// ---------------------------------------------------------------------------------------------------------
// Staging a contiguous run of global memory into shared memory with the TMA engine (cp.async.bulk, SASS
// UBLKCP): the 16-byte-aligned middle goes through ONE bulk copy issued by thread 0 and tracked by an
// mbarrier; the (< 16 byte) unaligned head and tail are copied by ordinary loads.  Shared byte k mirrors
// global byte (align_down(gaddr, 16) + k), so an element keeps its 16-byte phase and the returned `shift`
// locates the first requested byte.
//
//   smem   : shared buffer, 16-byte aligned, capacity >= nbytes + 32
//   g      : first requested global byte (element aligned)
//   nbytes : number of requested bytes (> 0)
//   bar    : one mbarrier per staged buffer, initialised with count 1 by thread 0 (see stage_init)
// All threads of the CTA call stage_issue(); they later call stage_wait() before reading.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_init(uint64_t *bar, int nbars) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < nbars; ++i) mbar_init(bar + i, 1);
        mbar_fence_init();
    }
    __syncthreads();
}

struct StageSplit {
    uint32_t shift;   // first requested byte within the 16-byte phase
    uint32_t head;    // unaligned leading bytes (< 16)
    uint32_t mid;     // 16-byte aligned body moved by the bulk copy
};

__device__ __forceinline__ StageSplit stage_split(const unsigned char *g, uint32_t nbytes) {
    const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
    const uintptr_t mid_begin = (ga + 15) & ~static_cast<uintptr_t>(15);
    const uintptr_t mid_end = (ga + nbytes) & ~static_cast<uintptr_t>(15);
    StageSplit s;
    s.shift = static_cast<uint32_t>(ga & 15);
    s.head = static_cast<uint32_t>(mid_begin > ga + nbytes ? nbytes : mid_begin - ga);
    s.mid = mid_end > mid_begin ? static_cast<uint32_t>(mid_end - mid_begin) : 0u;
    return s;
}

// ONE thread: arm the mbarrier and start the bulk copy of the aligned body.  Can be called before the rest of the
// CTA knows the run (e.g. by thread 0 straight after reading the tile's two boundary offsets), so that the copy
// overlaps the cooperative load of the offsets tile.
__device__ __forceinline__ void stage_bulk(unsigned char *smem, const unsigned char *g, uint32_t nbytes,
                                           uint64_t *bar) {
    const StageSplit s = stage_split(g, nbytes);
    mbar_expect_tx(bar, s.mid);      // mid == 0: nothing in flight, the phase completes immediately
    if (s.mid) bulk_g2s(smem + s.shift + s.head, g + s.head, s.mid, bar);
}

// ALL threads: copy the unaligned head and tail (< 16 bytes each) by ordinary loads; returns the shift.
template <int ELEM>
__device__ __forceinline__ uint32_t stage_edges(unsigned char *smem, const unsigned char *g, uint32_t nbytes) {
    const uint32_t shift = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g) & 15);
    if (threadIdx.x < 32) {      // at most 15 + 15 bytes: one warp is plenty, the others skip the address arithmetic
        const StageSplit s = stage_split(g, nbytes);
        const uint32_t b0 = threadIdx.x;
        if (b0 < s.head) smem[s.shift + b0] = g[b0];
        const uint32_t b1 = s.head + s.mid + threadIdx.x;
        if (b1 < nbytes) smem[s.shift + b1] = g[b1];
    }
    return shift;
}

template <int ELEM>
__device__ __forceinline__ uint32_t stage_issue(unsigned char *smem, const unsigned char *g, uint32_t nbytes,
                                                uint64_t *bar) {
    if (threadIdx.x == 0) stage_bulk(smem, g, nbytes, bar);
    return stage_edges<ELEM>(smem, g, nbytes);
}

__device__ __forceinline__ void stage_wait(uint64_t *bar, uint32_t phase) {
    mbar_wait(bar, phase);     // bulk bytes have landed (async proxy -> visible to waiting threads)
    __syncthreads();           // head / tail written by other threads
}

}  // namespace jg

namespace jg {

// ---------------------------------------------------------------------------------------------------------
// Owner table: for every position p of a tile's run [0, L) the local index of the event that owns it.  When all
// events of the tile are short it is written event-parallel (each thread fills its own event); otherwise it is built
// by a vectorised SCATTER (each non-empty event drops its index on its first position) followed by
// a running-MAX SCAN (rows of 128 positions per warp, 4 per lane, shuffles across lanes).  This is the on-chip
// form of offsets2parents (jagged.py:49-54, numpy.repeat) and feeds parents / localindex / broadcast / gathers /
// combinatorics without a per-element binary search.
//   own : __shared__ uint16_t[KM + 128], 16-byte aligned;  requires L <= KM and nev <= 65535
// All threads of the CTA call it; it ends with a __syncthreads().
// ---------------------------------------------------------------------------------------------------------
constexpr int kOwnerShortEvent = 24;     // events up to this long are written by their own thread

// positions [base, base + L) of the tile's run (relative to its first position): scatter + running maximum, free of
// divergence whatever the counts are.  A run longer than the table is swept in chunks of KM positions, one table each.
__device__ __forceinline__ void build_owner_range(uint16_t *own, const EventTile &tile, int64_t base, int L) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = kTileThreads / kWarp;
    const int padded = (L + 127) & ~127;
    // 1. clear (128-bit stores)
#pragma unroll 1
    for (int q = threadIdx.x; q < padded / 8; q += kTileThreads) reinterpret_cast<uint4 *>(own)[q] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // 2. scatter: non-empty events have distinct first positions
    const int64_t e0 = tile.off[0] + base;
#pragma unroll 1
    for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
        const int64_t a = tile.off[ev] - e0, b = tile.off[ev + 1] - e0;
        if (b > a && a >= 0 && a < L) own[a] = static_cast<uint16_t>(ev);
    }
    __syncthreads();
    // 3. running max: warp w owns a contiguous chunk of rows
    const int rows_total = padded >> 7;
    const int rows_per_warp = (rows_total + NW - 1) / NW;
    const int r_begin = warp * rows_per_warp;
    const int r_end = min(rows_total, r_begin + rows_per_warp);
    if (r_begin < r_end) {
        // owner of the chunk's first position (the event that started at or before it)
        uint32_t carry = 0;
        const int first = r_begin << 7;
        if ((first > 0 || base > 0) && first < L) carry = static_cast<uint32_t>(tile.find_event(e0 + first));
#pragma unroll 1
        for (int r = r_begin; r < r_end; ++r) {
            uint2 w = *reinterpret_cast<const uint2 *>(own + (r << 7) + lane * 4);
            uint32_t v0 = w.x & 0xffffu, v1 = w.x >> 16, v2 = w.y & 0xffffu, v3 = w.y >> 16;
            v1 = max(v1, v0);
            v2 = max(v2, v1);
            v3 = max(v3, v2);
            uint32_t incl = v3;
#pragma unroll
            for (int d = 1; d < kWarp; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl = max(incl, o);
            }
            uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1);
            before = lane == 0 ? carry : max(before, carry);
            v0 = max(v0, before);
            v1 = max(v1, before);
            v2 = max(v2, before);
            v3 = max(v3, before);
            w.x = v0 | (v1 << 16);
            w.y = v2 | (v3 << 16);
            *reinterpret_cast<uint2 *>(own + (r << 7) + lane * 4) = w;
            carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
        }
    }
    __syncthreads();
}

// The same table in three separable steps, for kernels that already walk the tile's events once (they scatter in that
// walk) and clear the table while their offsets tiles are still in flight -- and with NO binary search: the event that
// owns the first position of each warp's chunk registers itself in `carry` during the scatter.
//   lg: log2 of the positions one warp scans (a power of two >= 128, 8 << lg >= L: owner_chunk_lg)
__device__ __forceinline__ int owner_chunk_lg(int L) {
    int lg = 7;
    while ((kTileThreads / kWarp) << lg < L) ++lg;
    return lg;
}
__device__ __forceinline__ void owner_clear(uint16_t *own, int L) {
    const int padded = (L + 127) & ~127;
#pragma unroll 1
    for (int q = threadIdx.x; q < padded / 8; q += kTileThreads) reinterpret_cast<uint4 *>(own)[q] = make_uint4(0, 0, 0, 0);
}
// one event whose positions are [a, b) relative to the first position of the table (any thread; a barrier after the
// clear and before this, another one before owner_scan)
__device__ __forceinline__ void owner_scatter_event(uint16_t *own, uint16_t *carry, int ev, int64_t a, int64_t b, int L, int lg) {
    if (b > a && b > 0 && a < L) {
        if (a >= 0) own[a] = static_cast<uint16_t>(ev);
        const int lo = a <= 0 ? 0 : static_cast<int>((a + (int64_t(1) << lg) - 1) >> lg);
        const int hi = static_cast<int>((b < L ? b - 1 : int64_t(L) - 1) >> lg);
        for (int w = lo; w <= hi; ++w) carry[w] = static_cast<uint16_t>(ev);
    }
}
__device__ __forceinline__ void owner_scan(uint16_t *own, const uint16_t *carry_in, int L, int lg) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows_total = (L + 127) >> 7;
    const int r_begin = warp << (lg - 7);
    const int r_end = min(rows_total, (warp + 1) << (lg - 7));
    if (r_begin < r_end) {
        uint32_t carry = carry_in[warp];
#pragma unroll 1
        for (int r = r_begin; r < r_end; ++r) {
            uint2 w = *reinterpret_cast<const uint2 *>(own + (r << 7) + lane * 4);
            uint32_t v0 = w.x & 0xffffu, v1 = w.x >> 16, v2 = w.y & 0xffffu, v3 = w.y >> 16;
            v1 = max(v1, v0);
            v2 = max(v2, v1);
            v3 = max(v3, v2);
            uint32_t incl = v3;
#pragma unroll
            for (int d = 1; d < kWarp; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl = max(incl, o);
            }
            uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1);
            before = lane == 0 ? carry : max(before, carry);
            v0 = max(v0, before);
            v1 = max(v1, before);
            v2 = max(v2, before);
            v3 = max(v3, before);
            w.x = v0 | (v1 << 16);
            w.y = v2 | (v3 << 16);
            *reinterpret_cast<uint2 *>(own + (r << 7) + lane * 4) = w;
            carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
        }
    }
}

constexpr int kOwnerLongList = 64;       // longer events a tile may hold before the scatter + scan form takes over

__device__ __forceinline__ void build_owner_table(uint16_t *own, const EventTile &tile, int L) {
    // 0. usual case, every event of the tile is short: each thread writes its event's index over the event's own
    //    positions (a few 16-bit stores); no clearing, no scan, no binary search for the warp carries.
    // 1. a FEW longer events among short ones (a geometric tail: 87 % of the tiles of a mean-4 geometric distribution
    //    hold an event of more than 24 items, ~2 per tile): they are listed and written by one warp each, lanes over
    //    consecutive positions; the short ones as in 0.  (Round 1 sent such a tile to the scatter + scan form below:
    //    parents on geometric counts ran at 67 % against 103 % on Poisson(5).)
    // 2. many long events: scatter + running maximum, free of divergence whatever the counts are.
    __shared__ int s_nlong;
    __shared__ uint16_t s_long[kOwnerLongList];
    if (threadIdx.x == 0) s_nlong = 0;
    const int64_t e0s = tile.off[0];
    bool has_long = false;
#pragma unroll 1
    for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
        const int c = static_cast<int>(tile.off[ev + 1] - tile.off[ev]);
        if (c > kOwnerShortEvent) {
            has_long = true;
            continue;
        }
        uint16_t *o = own + (tile.off[ev] - e0s);
        const uint16_t me = static_cast<uint16_t>(ev);
        // four predicated stores per step: a warp walks its longest event, and a one-store loop spends
        // 7 instructions per position on it (28 % of tojagged's instructions, ncu source page)
#pragma unroll 1
        for (int k = 0; k < c; k += 4) {
            o[k] = me;
            if (k + 1 < c) o[k + 1] = me;
            if (k + 2 < c) o[k + 2] = me;
            if (k + 3 < c) o[k + 3] = me;
        }
    }
    if (!__syncthreads_or(has_long)) return;
    if (has_long) {
#pragma unroll 1
        for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads)
            if (tile.off[ev + 1] - tile.off[ev] > kOwnerShortEvent) {
                const int slot = atomicAdd(&s_nlong, 1);
                if (slot < kOwnerLongList) s_long[slot] = static_cast<uint16_t>(ev);
            }
    }
    __syncthreads();
    const int nlong = s_nlong;
    if (nlong > kOwnerLongList) {
        __syncthreads();                  // (everybody has read s_nlong before a later call may reset it)
        build_owner_range(own, tile, 0, L);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int q = warp; q < nlong; q += kTileThreads / kWarp) {
        const int ev = s_long[q];
        const int a = static_cast<int>(tile.off[ev] - e0s), b = static_cast<int>(tile.off[ev + 1] - e0s);
        const uint16_t me = static_cast<uint16_t>(ev);
#pragma unroll 1
        for (int p = a + lane; p < b; p += kWarp) own[p] = me;
    }
    __syncthreads();
}

}  // namespace jg

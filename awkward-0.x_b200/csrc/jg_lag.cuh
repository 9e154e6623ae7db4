// jg_lag.cuh -- tile prefixes without the look-back wait: "count ahead, consume one window later".
//
// The classic single-pass decoupled look-back makes every tile wait ~6 us on B200 for the slowest of its predecessors
// and for the prefix to ripple forward (profiles/r1_lookback_study.md), so such a kernel is only as fast as the bytes
// it can PARK on chip during the wait.  The two-pass alternative (count kernel, scan of the tile totals, fill kernel)
// never waits but reads its input from HBM twice.  The kernels built on this header do neither:
//
//   * ONE launch holds two kinds of CTA, dealt by a ticket counter in windows of W tiles:
//         C0 C1 F0 C2 F1 C3 F2 ... C(nw-1) F(nw-2) F(nw-1)
//     a COUNT CTA reduces one tile of the input to its aggregate (a streaming read that leaves the tile in the 126 MB
//     L2) and publishes it; a FILL CTA processes one tile for good -- it reads the tile again (an L2 hit: the window in
//     between is a few tens of MB), takes the tile's exclusive prefix and writes the results;
//   * the counts a fill CTA needs were issued a whole window of tickets earlier -- several microseconds -- so it
//     practically never waits, and because a CTA takes its ticket only once it is running, every CTA it could wait for
//     is resident or done: forward progress needs no assumption about dispatch order or co-residency;
//   * aggregates are published hierarchically (tile, 32 tiles, 1024 tiles, 32768 tiles; the thread that completes a
//     group carries its total one level up), so a prefix is the sum of at most 32 words per level: four independent
//     loads per lane of one warp, ONE L2 round trip, no chain of dependent polls and no prefix states.
//
// DRAM traffic is the algorithmic one (input once, output once).  The persistent form of the same idea -- one CTA
// alternating both phases -- was measured first and lost: a CTA's two dependent memory phases serialise
// (profiles/r2_lag_study.md).
#pragma once

#include "jg_common.cuh"

namespace jg {

struct LagDesc {
    uint32_t *ticket;  // execution-order counter of the launch's CTAs
    uint64_t *l0;      // per tile:              bit 63 = published, low 63 bits = aggregate
    uint64_t *l1;      // per 32 tiles:          low 16 bits = members added, high 48 bits = their sum
    uint64_t *l2;      // per 1024 tiles         (same format, members = 32 level-1 groups)
    uint64_t *l3;      // per 32768 tiles        (same format, members = 32 level-2 groups)
};

constexpr uint64_t kLagPublished = 1ull << 63;

static inline size_t lag_ws_bytes(int64_t ntiles) {
    const size_t t = static_cast<size_t>(ntiles < 1 ? 1 : ntiles);
    return 8 * (t + (t + 31) / 32 + (t + 1023) / 1024 + (t + 32767) / 32768) + 64 + 16;
}

// carve the descriptor arrays out of a (16-byte aligned) workspace; the caller zeroes lag_ws_bytes(ntiles) bytes
static inline LagDesc lag_carve(void *ws, int64_t ntiles) {
    const size_t t = static_cast<size_t>(ntiles < 1 ? 1 : ntiles);
    LagDesc d;
    d.ticket = static_cast<uint32_t *>(ws);
    d.l0 = static_cast<uint64_t *>(ws) + 2;
    d.l1 = d.l0 + t;
    d.l2 = d.l1 + (t + 31) / 32;
    d.l3 = d.l2 + (t + 1023) / 1024;
    return d;
}

// The launch.  A count CTA covers `cpt` consecutive tiles (one per warp), so a window of W tiles has cw = ceil(W / cpt)
// count CTAs and W fill CTAs; tickets are dealt as  C0 | C1 F0 | C2 F1 | ... | C(nw) F(nw-1)  (C(nw) is empty).
struct LagSchedule {
    uint32_t window, cpt, cw, ntiles;
};

static inline LagSchedule lag_schedule(int64_t ntiles, int64_t window, int cpt) {
    if (window > ntiles) window = ntiles;
    if (window < 1) window = 1;
    LagSchedule s;
    s.window = static_cast<uint32_t>(window);
    s.cpt = static_cast<uint32_t>(cpt);
    s.cw = static_cast<uint32_t>((window + cpt - 1) / cpt);
    s.ntiles = static_cast<uint32_t>(ntiles);
    return s;
}

static inline int64_t lag_grid(const LagSchedule &s) {
    const int64_t nw = (static_cast<int64_t>(s.ntiles) + s.window - 1) / s.window;
    return s.cw + nw * (static_cast<int64_t>(s.cw) + s.window);
}

#ifdef __CUDACC__

// ticket -> (role, tile).  role 0: count the tiles [tile, tile_end), one per warp; 1: fill `tile`; -1: nothing to do.
__device__ __forceinline__ int lag_role(uint32_t ticket, const LagSchedule &s, uint32_t &tile, uint32_t &tile_end) {
    uint32_t k, r;
    int role;
    if (ticket < s.cw) {
        k = 0;
        r = ticket;
        role = 0;
    } else {
        const uint32_t t = ticket - s.cw, period = s.cw + s.window;
        k = t / period;
        r = t - k * period;
        if (r < s.cw) {
            role = 0;
            k += 1;
        } else {
            role = 1;
            r -= s.cw;
        }
    }
    const uint64_t w0 = static_cast<uint64_t>(k) * s.window;
    uint64_t wend = w0 + s.window;
    if (wend > s.ntiles) wend = s.ntiles;
    const uint64_t t0 = w0 + (role == 0 ? static_cast<uint64_t>(r) * s.cpt : r);
    if (t0 >= wend) return -1;
    tile = static_cast<uint32_t>(t0);
    const uint64_t te = t0 + s.cpt;
    tile_end = static_cast<uint32_t>(te < wend ? te : wend);
    return role;
}

// all threads: this CTA's ticket (one atomic by thread 0, broadcast through shared memory)
__device__ __forceinline__ uint32_t lag_take_ticket(const LagDesc &d, uint32_t *s_ticket) {
    if (d.ticket == nullptr) return blockIdx.x;          // in-order dispatch of a 1-D grid (see launch_scan_win)
    if (threadIdx.x == 0) *s_ticket = atomicAdd(d.ticket, 1u);
    __syncthreads();
    return *s_ticket;
}

__device__ __forceinline__ uint64_t atom_add_relaxed_u64(uint64_t *p, uint64_t v) {
    uint64_t old;
    asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}

// ONE thread: publish the aggregate of `tile` (0 <= aggregate < 2^47)
__device__ __forceinline__ void lag_publish(const LagDesc &d, uint32_t tile, int64_t aggregate) {
    const uint64_t agg = static_cast<uint64_t>(aggregate);
    st_relaxed_u64(d.l0 + tile, kLagPublished | (agg & ~kLagPublished));
    uint64_t old = atom_add_relaxed_u64(d.l1 + (tile >> 5), (agg << 16) | 1ull);
    if ((old & 0xffffull) != 31ull) return;
    uint64_t sum = (old >> 16) + agg;                       // this thread completed its group of 32 tiles
    old = atom_add_relaxed_u64(d.l2 + (tile >> 10), (sum << 16) | 1ull);
    if ((old & 0xffffull) != 31ull) return;
    sum += old >> 16;
    atom_add_relaxed_u64(d.l3 + (tile >> 15), (sum << 16) | 1ull);
}

// ONE WARP (all 32 lanes): exclusive prefix of `tile` = sum of the aggregates of tiles 0 .. tile-1.
__device__ __forceinline__ int64_t lag_prefix(const LagDesc &d, uint32_t tile) {
    const int lane = threadIdx.x & 31;
    const uint32_t g = tile >> 5, h = tile >> 10, k = tile >> 15;
    const uint32_t n0 = tile & 31u, n1 = g & 31u, n2 = h & 31u;
    const uint64_t *p0 = d.l0 + (tile & ~31u) + lane;
    const uint64_t *p1 = d.l1 + (g & ~31u) + lane;
    const uint64_t *p2 = d.l2 + (h & ~31u) + lane;
    uint64_t sum;
    while (true) {
        const uint64_t w0 = static_cast<uint32_t>(lane) < n0 ? ld_relaxed_u64(p0) : kLagPublished;
        const uint64_t w1 = static_cast<uint32_t>(lane) < n1 ? ld_relaxed_u64(p1) : 32ull;
        const uint64_t w2 = static_cast<uint32_t>(lane) < n2 ? ld_relaxed_u64(p2) : 32ull;
        bool ok = (w0 >> 63) && ((w1 & 0xffffull) == 32ull) && ((w2 & 0xffffull) == 32ull);
        sum = (w0 & ~kLagPublished) + (w1 >> 16) + (w2 >> 16);
        for (uint32_t j = lane; j < k; j += 32) {
            const uint64_t w3 = ld_relaxed_u64(d.l3 + j);
            ok = ok && ((w3 & 0xffffull) == 32ull);
            sum += w3 >> 16;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        __nanosleep(64);
    }
    return static_cast<int64_t>(warp_reduce_add(sum));
}

#endif  // __CUDACC__

}  // namespace jg

// jg_lag.cuh -- "lagged" tile prefixes for persistent single-launch scans and compactions.
//
// The classic decoupled look-back makes every tile wait ~6 us on B200 for the slowest of its predecessors and for
// the prefix to ripple forward (profiles/r1_lookback_study.md), so a kernel is only as fast as the bytes it can PARK
// on chip during that wait.  Here the parking lot is the 126 MB L2 instead of shared memory:
//
//   * the kernel is PERSISTENT (grid = resident CTAs, tiles dealt round robin: tile = blockIdx.x + it * gridDim.x);
//   * in iteration `it` a CTA first runs PHASE A of the tile it will own LAG iterations later -- read the tile's
//     input, reduce it to one aggregate, publish it -- and then PHASE B of its current tile: fetch the input again
//     (an L2 hit: the whole grid's window of LAG * gridDim.x tiles is a few tens of MB), take the tile's exclusive
//     prefix from aggregates that were published a whole iteration ago, and write the results;
//   * aggregates are published hierarchically (tile, 32 tiles, 1024 tiles, 32768 tiles; the thread that completes a
//     group carries its total one level up), so a prefix is the sum of at most 32 words per level: four independent
//     loads per lane of one warp, ONE L2 round trip, no chain of dependent polls and no prefix states.
//
// DRAM traffic stays the algorithmic one (input once, output once); nothing waits unless a predecessor is more than
// a whole iteration late.  Forward progress does not depend on the order in which CTAs are dispatched: all CTAs
// of the grid are resident (the host sizes the grid from the occupancy query), and a CTA publishes phase A of iteration
// it + LAG before it can block in phase B of iteration it, so the smallest unfinished tile can always complete.
#pragma once

#include "jg_common.cuh"

namespace jg {

struct LagDesc {
    uint64_t *l0;      // per tile:              bit 63 = published, low 63 bits = aggregate
    uint64_t *l1;      // per 32 tiles:          low 16 bits = members added, high 48 bits = their sum
    uint64_t *l2;      // per 1024 tiles         (same format, members = 32 level-1 groups)
    uint64_t *l3;      // per 32768 tiles        (same format, members = 32 level-2 groups)
};

constexpr uint64_t kLagPublished = 1ull << 63;

static inline size_t lag_ws_bytes(int64_t ntiles) {
    const size_t t = static_cast<size_t>(ntiles < 1 ? 1 : ntiles);
    return 8 * (t + (t + 31) / 32 + (t + 1023) / 1024 + (t + 32767) / 32768) + 64;
}

// carve the descriptor arrays out of a (16-byte aligned) workspace; the caller zeroes lag_ws_bytes(ntiles) bytes
static inline LagDesc lag_carve(void *ws, int64_t ntiles) {
    const size_t t = static_cast<size_t>(ntiles < 1 ? 1 : ntiles);
    LagDesc d;
    d.l0 = static_cast<uint64_t *>(ws);
    d.l1 = d.l0 + t;
    d.l2 = d.l1 + (t + 31) / 32;
    d.l3 = d.l2 + (t + 1023) / 1024;
    return d;
}

#ifdef __CUDACC__

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

// generic-proxy accesses to shared memory (reads of the previous tile, head / tail bytes) must be ordered before the
// async proxy (TMA) overwrites the same buffer in the next iteration
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// L2 prefetch of a byte range (any alignment; one instruction per 128-byte line and caller thread)
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

#endif  // __CUDACC__

// resident grid of a persistent kernel: occupancy * SMs, at most `ntiles`
template <class Kernel>
static inline unsigned persistent_grid(Kernel kernel, int threads, int64_t ntiles) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0);
    if (occ < 1) occ = 1;
    int64_t g = static_cast<int64_t>(occ) * sm_count();
    if (g > ntiles) g = ntiles;
    if (g < 1) g = 1;
    return static_cast<unsigned>(g);
}

// Launch a persistent kernel whose CTAs wait on each other.  A cooperative launch makes the runtime guarantee what
// the protocol needs -- every CTA of the grid resident at the same time -- and fail with an error (instead of
// dead-locking) if the grid does not fit, e.g. when another kernel holds part of the device.
template <class... Params, class... Args>
static inline cudaError_t launch_resident(void (*kernel)(Params...), unsigned grid, unsigned threads, cudaStream_t stream,
                                          Args... args) {
    static_assert(sizeof...(Params) == sizeof...(Args), "argument count");
    auto pack = [&](Params... p) {
        void *ptrs[] = {static_cast<void *>(&p)...};
        return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kernel), dim3(grid), dim3(threads), ptrs, 0, stream);
    };
    return pack(args...);
}

}  // namespace jg

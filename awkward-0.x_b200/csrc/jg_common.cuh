// jg_common.cuh -- shared device/host helpers for libjagged_sm100a.so (B200, sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/jagged_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libjagged_sm100a is written for sm_100a (B200) only"
#endif

namespace jg {

// ------------------------------------------------------------------------------------------------
// host side: error string + launch checks
// ------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);            // jg_api.cu
int  check_launch(const char *what);             // jg_api.cu: cudaGetLastError -> 0 / -1
int  sm_count();                                 // cached cudaDevAttrMultiProcessorCount (per device)

#define JG_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            jg::set_error(__VA_ARGS__);  \
            return -2;                   \
        }                                \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// dtype dispatch
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int dtype_size(int dt) {
    switch (dt) {
        case JG_F32: case JG_I32: case JG_U32: return 4;
        case JG_F64: case JG_I64: case JG_U64: return 8;
        case JG_B8: case JG_U8: case JG_I8: return 1;
        case JG_I16: case JG_U16: return 2;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device: PTX helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

constexpr int kWarp = 32;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// streaming (evict-first) accessors for data that is touched exactly once
template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
    return __ldcs(p);
}
template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) {
    __stcs(p, v);
}

// a hint, no register result and nothing to wait for: start moving the line that holds *p towards L2
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    while (!mbar_try_wait(bar, phase)) {
    }
}
// generic-proxy accesses to shared memory (reads of the previous tile, head / tail bytes) must be ordered before the
// async proxy (TMA) overwrites the same buffer: called by the thread that starts the next bulk copy into a stage that
// has been read, after the barrier that ends those reads
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy; src, dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- warp / block primitives ----------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_incl_scan_add(T v, int lane) {
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_reduce_add(T v) {
#pragma unroll
    for (int d = kWarp / 2; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Block-wide exclusive scan of one value per thread (T = int32/int64).  `sh` holds one T per warp plus one.
// Returns the exclusive prefix for this thread and the block total through `total`.
template <typename T, int THREADS>
__device__ __forceinline__ T block_excl_scan(T v, T *sh, T &total) {
    constexpr int NW = THREADS / kWarp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = warp_incl_scan_add(v, lane);
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < NW ? sh[lane] : T(0);
        T wi = warp_incl_scan_add(w, lane);
        if (lane < NW) sh[lane] = wi - w;
        if (lane == NW - 1) sh[NW] = wi;
    }
    __syncthreads();
    T base = sh[warp];
    total = sh[NW];
    __syncthreads();
    return base + incl - v;
}

// ---- decoupled look-back tile state --------------------------------------------------------------------
// One 64-bit word per tile: top 2 bits = state, low 62 bits = value (all totals on this path are
// non-negative element counts < 2^62, so flag and value travel in a single atomic 64-bit store).
constexpr uint64_t kStateInvalid = 0ull;
constexpr uint64_t kStateAggregate = 1ull;
constexpr uint64_t kStatePrefix = 2ull;
constexpr uint64_t kValueMask = (1ull << 62) - 1;

struct LookbackWs {
    uint32_t ticket;      // dynamic tile counter of scan_lookback_kernel (tiles ordered by arrival => forward progress with
                          // any dispatch order); scan_tma_kernel takes blockIdx.x instead and relies on the in-order
                          // dispatch of a 1-D grid -- stated where it is launched (launch_scan_tma)
    uint32_t pad[3];
    uint64_t desc[1];     // ntiles entries follow
};

// Called by ONE thread of the tile after the tile aggregate is known.  Publishes it, walks back over the
// predecessors and returns the exclusive prefix of this tile; then publishes the inclusive prefix.
__device__ __forceinline__ int64_t lookback_exclusive(uint64_t *desc, uint32_t tile, int64_t aggregate) {
    const uint64_t agg = static_cast<uint64_t>(aggregate) & kValueMask;
    if (tile == 0) {
        st_relaxed_u64(&desc[0], (kStatePrefix << 62) | agg);
        return 0;
    }
    st_relaxed_u64(&desc[tile], (kStateAggregate << 62) | agg);
    uint64_t excl = 0;
    int64_t t = static_cast<int64_t>(tile) - 1;
    while (true) {
        uint64_t w = ld_relaxed_u64(&desc[t]);
        uint64_t st = w >> 62;
        if (st == kStateInvalid) continue;            // predecessor has not published yet: spin
        excl += w & kValueMask;
        if (st == kStatePrefix) break;
        --t;
    }
    st_relaxed_u64(&desc[tile], (kStatePrefix << 62) | ((excl + agg) & kValueMask));
    return static_cast<int64_t>(excl);
}

// Warp-cooperative variant.  Every round the 32 lanes fetch kLookDepth x 32 predecessor descriptors with
// independent loads (so one L2 round trip covers 256 tiles -- with ~900 tiles in flight on 148 SMs a
// 32-per-round walk costs ~28 dependent L2 latencies per tile and dominates short tiles), then consume
// them 32 at a time, nearest first.
template <int kLookDepth = 1>
__device__ __forceinline__ int64_t lookback_exclusive_warp(uint64_t *desc, uint32_t tile, int64_t aggregate) {
    const int lane = threadIdx.x & 31;
    const uint64_t agg = static_cast<uint64_t>(aggregate) & kValueMask;
    if (tile == 0) {
        if (lane == 0) st_relaxed_u64(&desc[0], (kStatePrefix << 62) | agg);
        return 0;
    }
    if (lane == 0) st_relaxed_u64(&desc[tile], (kStateAggregate << 62) | agg);
    uint64_t excl = 0;
    int64_t base = static_cast<int64_t>(tile) - 1;      // group g, lane l looks at tile base - 32*g - l
    bool done = false;
    while (!done) {
        uint64_t w[kLookDepth];
#pragma unroll
        for (int g = 0; g < kLookDepth; ++g) {
            const int64_t t = base - 32 * g - lane;
            w[g] = (t >= 0) ? ld_relaxed_u64(&desc[t]) : (kStatePrefix << 62);
        }
#pragma unroll
        for (int g = 0; g < kLookDepth; ++g) {
            if (done) break;
            const uint64_t st = w[g] >> 62;
            const unsigned invalid = __ballot_sync(0xffffffffu, st == kStateInvalid);
            const unsigned prefix = __ballot_sync(0xffffffffu, st == kStatePrefix);
            const int first_prefix = prefix ? (__ffs(prefix) - 1) : 32;
            const unsigned need = first_prefix >= 31 ? 0xffffffffu : ((2u << first_prefix) - 1u);
            if (invalid & need) {                       // a needed predecessor has not published: re-poll from here
                __nanosleep(32);                        // (back off: the predecessor's CTA shares the memory system)
                break;
            }
            const uint64_t contrib = (lane <= first_prefix) ? (w[g] & kValueMask) : 0ull;
            excl += warp_reduce_add(contrib);
            base -= 32;
            if (prefix) done = true;
        }
    }
    if (lane == 0) st_relaxed_u64(&desc[tile], (kStatePrefix << 62) | ((excl + agg) & kValueMask));
    return static_cast<int64_t>(excl);
}

#endif  // __CUDACC__

}  // namespace jg

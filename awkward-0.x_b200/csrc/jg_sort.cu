// jg_sort.cu -- per-event argsort (awkward0/array/jagged.py:1603-1631).
//
// The reference repeatedly peels "all elements equal to the event's current extreme" off every event (one
// segmented reduction + one mask select per distinct value) and emits their local indices in position order;
// NaN is never equal to anything and comes out last.  Logically that is a STABLE sort of every event by value
// (descending unless `ascending`), -0.0 == 0.0 and other ties in position order, NaN last in both directions.
// The comparator below is that total order on (value, position), so any sorting network gives the same answer.
//
//   argsort_kernel       : event tile (512 events per CTA, content run staged by TMA, owner table on chip);
//                          ELEMENT-parallel rank sort -- each element counts the elements of its own event that
//                          sort before it and writes its local index to that slot.  Events are a handful of
//                          elements long, so the O(c^2) comparisons are cheaper than any network and the kernel
//                          is one pass: v*M + 8(N+1) read, 8*M written.
//   events > kSortLong   : registered in a list and sorted one by one by a bitonic network over an index array
//                          (shared-memory phases for strides < 4096, one launch per larger stride).
#include "jg_tile.cuh"

namespace jg {

constexpr int kSortStage = 4096;     // positions of a tile that are staged on chip
constexpr int kSortLong = 2048;      // longer events go through the bitonic network
constexpr int kBitonicChunk = 4096;  // indices per CTA in the shared-memory phases
constexpr int kBitonicThreads = 256;

template <typename T> struct SortIsFloat { static constexpr bool value = false; };
template <> struct SortIsFloat<float> { static constexpr bool value = true; };
template <> struct SortIsFloat<double> { static constexpr bool value = true; };

// strict total order: does (a at position ia) come out before (b at position ib)?
template <typename T, bool ASC>
__device__ __forceinline__ bool sorts_before(T a, int64_t ia, T b, int64_t ib) {
    if constexpr (SortIsFloat<T>::value) {
        const bool na = a != a, nb = b != b;
        if (na | nb) return na == nb ? ia < ib : nb;      // NaN last (jagged.py:1618-1620), NaNs in position order
    }
    if (a == b) return ia < ib;                            // ties (and -0.0 == 0.0) keep their position order
    return ASC ? a < b : a > b;
}

template <typename T, bool ASC>
__global__ void __launch_bounds__(kTileThreads)
argsort_kernel(const T *__restrict__ content, const int64_t *__restrict__ offsets, int64_t n,
               int64_t *__restrict__ out, int32_t *__restrict__ status, int64_t *__restrict__ long_list,
               int64_t long_cap) {
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) unsigned char s_stage[kSortStage * sizeof(T) + 32];
    __shared__ __align__(16) uint16_t s_own[kSortStage + 128];
    __shared__ uint64_t s_bar;

    EventTile tile;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        int64_t ev0;
        int nev;
        EventTile::range(n, blockIdx.x, ev0, nev);
        const int64_t a = __ldg(offsets + ev0), b = __ldg(offsets + ev0 + nev);
        if (b > a && b - a <= kSortStage)
            stage_bulk(s_stage, reinterpret_cast<const unsigned char *>(content + a),
                       static_cast<uint32_t>((b - a) * sizeof(T)), &s_bar);
    }
    tile.load(s_off, offsets, n, blockIdx.x);
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const int64_t base0 = __ldg(offsets);
    const int64_t len = e1 - e0;
    if (len == 0) return;

    if (len <= kSortStage) {
        const int L = static_cast<int>(len);
        const uint32_t shift = stage_edges<sizeof(T)>(s_stage, reinterpret_cast<const unsigned char *>(content + e0),
                                                      static_cast<uint32_t>(L * sizeof(T)));
        build_owner_table(s_own, tile, L);
        stage_wait(&s_bar, 0);
        const T *sm = reinterpret_cast<const T *>(s_stage + shift);
        int64_t *dst = out + (e0 - base0);
#pragma unroll 1
        for (int p = threadIdx.x; p < L; p += kTileThreads) {
            const int ev = s_own[p];
            const int a = static_cast<int>(s_off[ev] - e0), b = static_cast<int>(s_off[ev + 1] - e0);
            if (b - a > kSortLong) continue;
            const T key = sm[p];
            int rank = 0;
#pragma unroll 1
            for (int j = a; j < b; ++j) rank += sorts_before<T, ASC>(sm[j], j, key, p) ? 1 : 0;
            dst[a + rank] = p - a;
        }
    } else {
        // the tile's run does not fit on chip (long events around): same rank sort straight from L1/L2
#pragma unroll 1
        for (int64_t p = e0 + threadIdx.x; p < e1; p += kTileThreads) {
            const int ev = tile.find_event(p);
            const int64_t a = s_off[ev], b = s_off[ev + 1];
            if (b - a > kSortLong) continue;
            const T key = __ldg(content + p);
            int rank = 0;
#pragma unroll 1
            for (int64_t j = a; j < b; ++j) rank += sorts_before<T, ASC>(__ldg(content + j), j, key, p) ? 1 : 0;
            out[a - base0 + rank] = p - a;
        }
    }
    // long events are left to the bitonic path: put them on the list
#pragma unroll 1
    for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
        if (s_off[ev + 1] - s_off[ev] > kSortLong) {
            const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long *>(long_list), 1ull);
            if (static_cast<int64_t>(slot) < long_cap) long_list[1 + slot] = tile.event0 + ev;
            atomicOr(status, JG_ST_SORT_LONG);
        }
    }
}

// ---- bitonic network over an index array x[0..P), P a power of two >= kBitonicChunk; indices >= c are padding --
template <typename T, bool ASC>
__device__ __forceinline__ bool index_before(const T *__restrict__ keys, uint32_t c, uint32_t i, uint32_t j) {
    if (i >= c || j >= c) return i < j;                   // padding sorts after every real element
    return sorts_before<T, ASC>(__ldg(keys + i), i, __ldg(keys + j), j);
}

// Shared-memory phases.  INIT: x[t] = t, then every stage k = 2 .. kBitonicChunk.  Otherwise: the strides
// j = kBitonicChunk/2 .. 1 of stage k (the larger strides of that stage were done by bitonic_global_kernel).
template <typename T, bool ASC, bool INIT>
__global__ void __launch_bounds__(kBitonicThreads)
bitonic_local_kernel(const T *__restrict__ keys, uint32_t c, uint32_t *__restrict__ x, uint32_t k_stage) {
    __shared__ uint32_t s[kBitonicChunk];
    const uint32_t base = blockIdx.x * static_cast<uint32_t>(kBitonicChunk);
    for (int t = threadIdx.x; t < kBitonicChunk; t += kBitonicThreads) s[t] = INIT ? base + t : x[base + t];
    __syncthreads();
    const uint32_t k_first = INIT ? 2u : k_stage, k_last = INIT ? static_cast<uint32_t>(kBitonicChunk) : k_stage;
    for (uint32_t k = k_first; k <= k_last && k != 0; k <<= 1) {
        uint32_t j = k >> 1;
        if (j > kBitonicChunk / 2) j = kBitonicChunk / 2;
        for (; j > 0; j >>= 1) {
            for (uint32_t q = threadIdx.x; q < kBitonicChunk / 2; q += kBitonicThreads) {
                const uint32_t t = ((q & ~(j - 1)) << 1) | (q & (j - 1));
                const uint32_t l = t | j;
                const bool up = ((base + t) & k) == 0;
                const uint32_t a = s[t], b = s[l];
                if (index_before<T, ASC>(keys, c, b, a) == up) {
                    s[t] = b;
                    s[l] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < kBitonicChunk; t += kBitonicThreads) x[base + t] = s[t];
}

// one stride j >= kBitonicChunk of stage k
template <typename T, bool ASC>
__global__ void __launch_bounds__(kBitonicThreads)
bitonic_global_kernel(const T *__restrict__ keys, uint32_t c, uint32_t *__restrict__ x, uint32_t half, uint32_t k,
                      uint32_t j) {
    for (uint32_t q = blockIdx.x * kBitonicThreads + threadIdx.x; q < half; q += gridDim.x * kBitonicThreads) {
        const uint32_t t = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        const uint32_t l = t | j;
        const bool up = (t & k) == 0;
        const uint32_t a = x[t], b = x[l];
        if (index_before<T, ASC>(keys, c, b, a) == up) {
            x[t] = b;
            x[l] = a;
        }
    }
}

__global__ void __launch_bounds__(256)
bitonic_emit_kernel(const uint32_t *__restrict__ x, uint32_t c, int64_t *__restrict__ out) {
    for (uint32_t r = blockIdx.x * 256u + threadIdx.x; r < c; r += gridDim.x * 256u) out[r] = x[r];
}

static inline uint32_t bitonic_size(int64_t count) {
    uint32_t p = kBitonicChunk;
    while (static_cast<int64_t>(p) < count) p <<= 1;
    return p;
}

template <typename T, bool ASC>
static int launch_argsort_long(const T *keys, int64_t count, int64_t *out, void *ws, size_t ws_bytes, cudaStream_t s) {
    JG_REQUIRE(count <= (1ll << 30), "jg_argsort_long: events longer than 2^30 elements are not supported");
    const uint32_t c = static_cast<uint32_t>(count), P = bitonic_size(count);
    JG_REQUIRE(ws && ws_bytes >= static_cast<size_t>(P) * sizeof(uint32_t), "jg_argsort_long: workspace too small");
    uint32_t *x = static_cast<uint32_t *>(ws);
    const unsigned chunks = P / kBitonicChunk;
    const uint32_t half = P / 2;
    unsigned gblocks = half / kBitonicThreads;
    const unsigned cap = static_cast<unsigned>(sm_count()) * 16u;
    if (gblocks > cap) gblocks = cap;
    bitonic_local_kernel<T, ASC, true><<<chunks, kBitonicThreads, 0, s>>>(keys, c, x, 0);
    for (uint32_t k = 2u * kBitonicChunk; k <= P && k != 0; k <<= 1) {
        for (uint32_t j = k >> 1; j >= static_cast<uint32_t>(kBitonicChunk); j >>= 1)
            bitonic_global_kernel<T, ASC><<<gblocks, kBitonicThreads, 0, s>>>(keys, c, x, half, k, j);
        bitonic_local_kernel<T, ASC, false><<<chunks, kBitonicThreads, 0, s>>>(keys, c, x, k);
    }
    unsigned eblocks = static_cast<unsigned>(ceil_div(count, 256));
    if (eblocks > cap) eblocks = cap;
    bitonic_emit_kernel<<<eblocks, 256, 0, s>>>(x, c, out);
    return check_launch("bitonic argsort");
}

template <typename T>
static int dispatch_argsort(const void *content, const int64_t *offsets, int64_t n, int ascending, int64_t *out,
                            int32_t *status, int64_t *long_list, int64_t long_cap, cudaStream_t s) {
    const T *c = static_cast<const T *>(content);
    if (ascending)
        argsort_kernel<T, true><<<tile_grid(n), kTileThreads, 0, s>>>(c, offsets, n, out, status, long_list, long_cap);
    else
        argsort_kernel<T, false><<<tile_grid(n), kTileThreads, 0, s>>>(c, offsets, n, out, status, long_list, long_cap);
    return check_launch("argsort_kernel");
}

template <typename T>
static int dispatch_argsort_long(const void *content, int64_t start, int64_t count, int ascending, int64_t *out,
                                 void *ws, size_t ws_bytes, cudaStream_t s) {
    const T *keys = static_cast<const T *>(content) + start;
    return ascending ? launch_argsort_long<T, true>(keys, count, out, ws, ws_bytes, s)
                     : launch_argsort_long<T, false>(keys, count, out, ws, ws_bytes, s);
}

}  // namespace jg

using namespace jg;

#define JG_SORT_DISPATCH(FN, ...)                                    \
    switch (dtype) {                                                 \
        case JG_F32: return FN<float>(__VA_ARGS__);                  \
        case JG_F64: return FN<double>(__VA_ARGS__);                 \
        case JG_I32: return FN<int32_t>(__VA_ARGS__);                \
        case JG_I64: return FN<int64_t>(__VA_ARGS__);                \
        case JG_U8:                                                  \
        case JG_B8: return FN<uint8_t>(__VA_ARGS__);                 \
        case JG_I8: return FN<int8_t>(__VA_ARGS__);                  \
        case JG_I16: return FN<int16_t>(__VA_ARGS__);                \
        case JG_U16: return FN<uint16_t>(__VA_ARGS__);               \
        case JG_U32: return FN<uint32_t>(__VA_ARGS__);               \
        case JG_U64: return FN<uint64_t>(__VA_ARGS__);               \
    }

extern "C" {

int jg_argsort(const void *content, int dtype, const int64_t *offsets, int64_t n, int ascending, int64_t *out,
               int32_t *status, int64_t *long_list, int64_t long_cap, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && status && long_list && long_cap >= 0, "jg_argsort: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = as_stream(stream);
    JG_SORT_DISPATCH(dispatch_argsort, content, offsets, n, ascending, out, status, long_list, long_cap, s)
    set_error("jg_argsort: unsupported dtype %d", dtype);
    return -2;
}

int64_t jg_argsort_long_workspace(int64_t count) {
    return count <= 0 ? 0 : static_cast<int64_t>(bitonic_size(count)) * static_cast<int64_t>(sizeof(uint32_t));
}

int jg_argsort_long(const void *content, int dtype, int64_t start, int64_t count, int ascending, int64_t *out,
                    void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(content && out && start >= 0 && count >= 0, "jg_argsort_long: bad arguments");
    if (count == 0) return 0;
    cudaStream_t s = as_stream(stream);
    JG_SORT_DISPATCH(dispatch_argsort_long, content, start, count, ascending, out, ws, ws_bytes, s)
    set_error("jg_argsort_long: unsupported dtype %d", dtype);
    return -2;
}

}  // extern "C"

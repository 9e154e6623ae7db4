// jg_sort.cu -- per-event argsort (awkward0/array/jagged.py:1603-1631).
//
// The reference repeatedly peels "all elements equal to the event's current extreme" off every event (one
// segmented reduction + one mask select per distinct value) and emits their local indices in position order;
// NaN is never equal to anything and comes out last.  Logically that is a STABLE sort of every event by value
// (descending unless `ascending`), -0.0 == 0.0 and other ties in position order, NaN last in both directions.
// The comparator below is that total order on (value, position), so any sorting network gives the same answer.
//
//   argsort_kernel       : event tile (512 events per CTA, content run staged by TMA and turned into
//                          order-preserving unsigned keys in place).  Events of up to 8 elements -- nearly all
//                          of them -- are ranked by their own thread in registers: 28 compares, fully unrolled,
//                          no loop and no divergence; each element's local index goes to the slot of its rank.
//                          Medium events (9 .. 2048) are listed and ranked by half a warp each.  One pass:
//                          v*M + 8(N+1) read, 8*M written.
//   events > kSortLong   : registered in a list and sorted one by one by a bitonic network over an index array
//                          (shared-memory phases for strides < 4096, one launch per larger stride).
#include "jg_tile.cuh"

namespace jg {

template <typename T> struct SortStage { static constexpr int value = sizeof(T) >= 8 ? 3072 : 4096; };   // positions a tile stages on chip
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

// Order-preserving key: an unsigned integer of the same width whose ascending order is the output order of the
// event (NaN -> all ones = last in both directions, -0.0 folded onto +0.0 so that it ties with it, descending =
// bitwise complement).  With keys the comparator of the rank sort is one unsigned compare.
template <int BYTES> struct KeyOf;
template <> struct KeyOf<1> { using type = uint8_t; };
template <> struct KeyOf<2> { using type = uint16_t; };
template <> struct KeyOf<4> { using type = uint32_t; };
template <> struct KeyOf<8> { using type = uint64_t; };

template <typename T, bool ASC>
__device__ __forceinline__ typename KeyOf<sizeof(T)>::type to_key(T x) {
    using U = typename KeyOf<sizeof(T)>::type;
    constexpr U kSign = static_cast<U>(U(1) << (8 * sizeof(T) - 1));
    U u;
    if constexpr (SortIsFloat<T>::value) {
        if (x != x) return static_cast<U>(~U(0));
        x += T(0);                                        // -0.0 + 0.0 = +0.0
        if constexpr (sizeof(T) == 4) u = __float_as_uint(x);
        else u = static_cast<U>(__double_as_longlong(x));
        u = (u & kSign) ? static_cast<U>(~u) : static_cast<U>(u | kSign);
    } else if constexpr (static_cast<T>(-1) < static_cast<T>(0)) {
        u = static_cast<U>(static_cast<U>(x) ^ kSign);    // signed integers: flip the sign bit
    } else {
        u = static_cast<U>(x);
    }
    return ASC ? u : static_cast<U>(~u);
}

constexpr int kSortRegs = 8;         // events up to this length are ranked entirely in registers

// Ranks of 8 keys held in registers, packed one hex digit per key.  rank_i = #(j < i: k_j <= k_i) + #(j > i:
// k_j < k_i); with T_i = #(j > i: k_i <= k_j) the second term is (7 - i) - T_i, so ONE test per pair i < j feeds
// both keys: digit j += 1, digit i -= 1.  The digits start at 7 - i; intermediate borrows between digits cancel
// because the final value of every digit is a rank in [0, 7].  Two instructions per pair (compare, predicated add).
static_assert(kSortRegs == 8, "rank_packed is written for 8 keys");
template <typename U>
__device__ __forceinline__ uint32_t rank_packed(const U (&k)[kSortRegs]) {
    uint32_t ranks = 0x01234567u;
#pragma unroll
    for (int i = 0; i < kSortRegs; ++i) {
#pragma unroll
        for (int j = i + 1; j < kSortRegs; ++j)
            if (k[i] <= k[j]) ranks += (1u << (4 * j)) - (1u << (4 * i));
    }
    return ranks;
}

template <typename T, bool ASC>
__global__ void __launch_bounds__(kTileThreads)
argsort_kernel(const T *__restrict__ content, const int64_t *__restrict__ offsets, int64_t n,
               int64_t *__restrict__ out, int32_t *__restrict__ status, int64_t *__restrict__ long_list,
               int64_t long_cap) {
    using U = typename KeyOf<sizeof(T)>::type;
    constexpr int kSortStage = SortStage<T>::value;
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) unsigned char s_stage[kSortStage * sizeof(T) + 32];
    __shared__ uint16_t s_med[kTileEvents];      // medium events (9 .. kSortLong elements) of the tile
    __shared__ uint16_t s_res[kSortStage];       // the tile's result (local indices < 2048), written out coalesced
    __shared__ uint64_t s_bar;
    __shared__ int s_cnt;

    EventTile tile;
    if (threadIdx.x == 0) {
        s_cnt = 0;
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
    bool has_long = false;

    if (len <= kSortStage) {
        const int L = static_cast<int>(len);
        const uint32_t shift = stage_edges<sizeof(T)>(s_stage, reinterpret_cast<const unsigned char *>(content + e0),
                                                      static_cast<uint32_t>(L * sizeof(T)));
        stage_wait(&s_bar, 0);
        U *sk = reinterpret_cast<U *>(s_stage + shift);
        // values -> keys, in place
#pragma unroll 1
        for (int q = threadIdx.x; q < L; q += kTileThreads) sk[q] = to_key<T, ASC>(reinterpret_cast<const T *>(sk)[q]);
        __syncthreads();
        // phase 1, event-parallel: short events are ranked in registers (no loop, no divergence); the elements of
        // medium events are put on a list
#pragma unroll 1
        for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
            const int a = static_cast<int>(s_off[ev] - e0);
            const int64_t c64 = s_off[ev + 1] - s_off[ev];
            if (c64 <= kSortRegs) {
                const int c = static_cast<int>(c64);
                U k[kSortRegs];
#pragma unroll
                for (int i = 0; i < kSortRegs; ++i) k[i] = (i < c) ? sk[a + i] : static_cast<U>(~U(0));
                const uint32_t ranks = rank_packed(k);
#pragma unroll
                for (int i = 0; i < kSortRegs; ++i)
                    if (i < c) s_res[a + ((ranks >> (4 * i)) & 0xFu)] = static_cast<uint16_t>(i);
            } else if (c64 <= kSortLong) {
                s_med[atomicAdd(&s_cnt, 1)] = static_cast<uint16_t>(ev);
            } else {
                has_long = true;
            }
        }
        __syncthreads();
        // phase 2, half a warp per medium event: every lane ranks one element against the whole event (the keys
        // are broadcast reads; events of 9 .. 16 elements, the usual case, take one pass of the 16 lanes)
        const int nmed = s_cnt;
        constexpr int kGroup = 16;
        const int gl = threadIdx.x & (kGroup - 1), group = threadIdx.x / kGroup;
#pragma unroll 1
        for (int m = group; m < nmed; m += kTileThreads / kGroup) {
            const int ev = s_med[m];
            const U *first = sk + (s_off[ev] - e0), *last = sk + (s_off[ev + 1] - e0);
#pragma unroll 1
            for (const U *mine = first + gl; mine < last; mine += kGroup) {
                const U key = *mine;
                int rank = 0;
#pragma unroll 1
                for (const U *q = first; q < last; ++q) {
                    const U kj = *q;
                    rank += ((kj < key) | ((kj == key) & (q < mine))) ? 1 : 0;
                }
                s_res[(first - sk) + rank] = static_cast<uint16_t>(mine - first);
            }
        }
        __syncthreads();
        // coalesced write-out (positions of long events hold garbage here; jg_argsort_long overwrites them)
        int64_t *dst = out + (e0 - base0);
#pragma unroll 2
        for (int q = threadIdx.x; q < L; q += kTileThreads) __stcs(dst + q, static_cast<int64_t>(s_res[q]));
    } else {
        // the tile's run does not fit on chip (long events around): rank sort straight from L1/L2
#pragma unroll 1
        for (int64_t p = e0 + threadIdx.x; p < e1; p += kTileThreads) {
            const int ev = tile.find_event(p);
            const int64_t a = s_off[ev], b = s_off[ev + 1];
            if (b - a > kSortLong) continue;
            const U key = to_key<T, ASC>(__ldg(content + p));
            int rank = 0;
#pragma unroll 1
            for (int64_t j = a; j < b; ++j) {
                const U kj = to_key<T, ASC>(__ldg(content + j));
                rank += ((kj < key) | ((kj == key) & (j < p))) ? 1 : 0;
            }
            out[a - base0 + rank] = p - a;
        }
#pragma unroll 1
        for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) has_long |= (s_off[ev + 1] - s_off[ev] > kSortLong);
    }
    // long events are left to the bitonic path: put them on the list
    if (has_long) {
#pragma unroll 1
        for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
            if (s_off[ev + 1] - s_off[ev] > kSortLong) {
                const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long *>(long_list), 1ull);
                if (static_cast<int64_t>(slot) < long_cap) long_list[1 + slot] = tile.event0 + ev;
                atomicOr(status, JG_ST_SORT_LONG);
            }
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

// jg_reduce.cuh -- reduction functors and partial-result merging shared by the event-tiled (jg_reduce.cu) and the
// element-tiled (jg_elem.cu) segmented reductions: NaN -> identity, "later element wins" on equal min / max,
// first occurrence for argmin / argmax (jagged.py:1459-1601).
#pragma once

#include <math_constants.h>

#include "jg_common.cuh"

namespace jg {

template <typename T> struct IsFloat { static constexpr bool value = false; };
template <> struct IsFloat<float> { static constexpr bool value = true; };
template <> struct IsFloat<double> { static constexpr bool value = true; };

template <typename T>
__device__ __forceinline__ bool is_nan(T x) {
    if constexpr (IsFloat<T>::value) return x != x;
    else return false;
}

template <typename A> __device__ __forceinline__ A max_identity();
template <> __device__ __forceinline__ float max_identity<float>() { return -CUDART_INF_F; }
template <> __device__ __forceinline__ double max_identity<double>() { return -CUDART_INF; }
template <> __device__ __forceinline__ int32_t max_identity<int32_t>() { return INT32_MIN; }
template <> __device__ __forceinline__ int64_t max_identity<int64_t>() { return INT64_MIN; }
template <> __device__ __forceinline__ uint8_t max_identity<uint8_t>() { return 0; }
template <> __device__ __forceinline__ int8_t max_identity<int8_t>() { return INT8_MIN; }
template <> __device__ __forceinline__ int16_t max_identity<int16_t>() { return INT16_MIN; }
template <> __device__ __forceinline__ uint16_t max_identity<uint16_t>() { return 0; }
template <> __device__ __forceinline__ uint32_t max_identity<uint32_t>() { return 0; }
template <> __device__ __forceinline__ uint64_t max_identity<uint64_t>() { return 0; }
template <typename A> __device__ __forceinline__ A min_identity();
template <> __device__ __forceinline__ float min_identity<float>() { return CUDART_INF_F; }
template <> __device__ __forceinline__ double min_identity<double>() { return CUDART_INF; }
template <> __device__ __forceinline__ int32_t min_identity<int32_t>() { return INT32_MAX; }
template <> __device__ __forceinline__ int64_t min_identity<int64_t>() { return INT64_MAX; }
template <> __device__ __forceinline__ uint8_t min_identity<uint8_t>() { return 255; }
template <> __device__ __forceinline__ int8_t min_identity<int8_t>() { return INT8_MAX; }
template <> __device__ __forceinline__ int16_t min_identity<int16_t>() { return INT16_MAX; }
template <> __device__ __forceinline__ uint16_t min_identity<uint16_t>() { return UINT16_MAX; }
template <> __device__ __forceinline__ uint32_t min_identity<uint32_t>() { return UINT32_MAX; }
template <> __device__ __forceinline__ uint64_t min_identity<uint64_t>() { return UINT64_MAX; }

// Reduction functor: A = accumulator/output type, T = content type.
template <int OP, typename T, typename A>
struct Red {
    __device__ static __forceinline__ A identity() {
        if constexpr (OP == JG_SUM) return A(0);
        else if constexpr (OP == JG_PROD) return A(1);
        else if constexpr (OP == JG_MIN) return min_identity<A>();
        else if constexpr (OP == JG_MAX) return max_identity<A>();
        else if constexpr (OP == JG_ANY) return A(0);
        else return A(1);
    }
    // content value -> accumulator domain (NaN -> identity, jagged.py:1524-1528; any/all: astype(bool))
    __device__ static __forceinline__ A lift(T x) {
        if (is_nan(x)) return identity();
        if constexpr (OP == JG_ANY || OP == JG_ALL) return A(x != T(0));
        else return static_cast<A>(x);
    }
    // acc (earlier elements) followed by the raw content value x: combine(acc, lift(x)) in fewer instructions.
    // For float min/max a single ordered comparison does both jobs: it is false for NaN (NaN is skipped, i.e.
    // replaced by the identity) and true on equality (the later element wins, like numpy's loops).
    __device__ static __forceinline__ A accumulate(A acc, T x) {
        if constexpr (OP == JG_MIN && IsFloat<T>::value && sizeof(A) == sizeof(T)) return (static_cast<A>(x) <= acc) ? static_cast<A>(x) : acc;
        else if constexpr (OP == JG_MAX && IsFloat<T>::value && sizeof(A) == sizeof(T)) return (static_cast<A>(x) >= acc) ? static_cast<A>(x) : acc;
        else return combine(acc, lift(x));
    }
    // acc (earlier elements) combined with x (a later element)
    __device__ static __forceinline__ A combine(A acc, A x) {
        if constexpr (OP == JG_SUM) return acc + x;
        else if constexpr (OP == JG_PROD) return acc * x;
        else if constexpr (OP == JG_MIN) return (acc < x) ? acc : x;      // equal -> later element
        else if constexpr (OP == JG_MAX) return (acc > x) ? acc : x;
        else if constexpr (OP == JG_ANY) return A(acc | x);
        else return A(acc & x);
    }
};

constexpr int kBlockEventMin = 1 << 11;   // events at least this long are reduced by the whole CTA (a warp alone keeps too few loads in flight)

// partial result of a lane/warp over a strided subset; `pos` = position of the element that produced a
// min/max value (needed only to reproduce "later wins on ties" exactly for +-0.0)
template <typename A>
struct Partial {
    A val;
    int64_t pos;
    bool has;
};

template <int OP, typename T, typename A>
__device__ __forceinline__ Partial<A> merge(Partial<A> a, Partial<A> b) {
    if (!b.has) return a;
    if (!a.has) return b;
    Partial<A> r;
    r.has = true;
    if constexpr (OP == JG_MIN || OP == JG_MAX) {
        const bool a_better = (OP == JG_MIN) ? (a.val < b.val) : (a.val > b.val);
        const bool b_better = (OP == JG_MIN) ? (b.val < a.val) : (b.val > a.val);
        const bool take_a = a_better || (!b_better && a.pos > b.pos);
        r.val = take_a ? a.val : b.val;
        r.pos = take_a ? a.pos : b.pos;
    } else {
        r.val = Red<OP, T, A>::combine(a.val, b.val);
        r.pos = 0;
    }
    return r;
}

template <int OP, typename T, typename A>
__device__ __forceinline__ Partial<A> shuffle_merge(Partial<A> p) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Partial<A> o;
        o.val = __shfl_xor_sync(0xffffffffu, p.val, d);
        o.pos = __shfl_xor_sync(0xffffffffu, p.pos, d);
        o.has = __shfl_xor_sync(0xffffffffu, static_cast<int>(p.has), d) != 0;
        p = merge<OP, T, A>(p, o);
    }
    return p;
}

template <int OP, typename T, typename A>
__device__ __forceinline__ Partial<A> strided_partial(const T *__restrict__ content, int64_t a, int64_t b, int first,
                                                      int step) {
    // four independent loads in flight per thread: a long event is read by few threads, and one load per iteration
    // makes every iteration a full memory round trip (a 37 k-element event took a warp 0.8 ms that way)
    constexpr int U = 4;
    Partial<A> p;
    p.has = false;
    p.val = Red<OP, T, A>::identity();
    p.pos = -1;
#pragma unroll 1
    for (int64_t j0 = a + first; j0 < b; j0 += static_cast<int64_t>(U) * step) {
        T raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = j0 + static_cast<int64_t>(u) * step;
            if (j < b) raw[u] = __ldcs(content + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = j0 + static_cast<int64_t>(u) * step;
            if (j >= b) break;
            const A x = Red<OP, T, A>::lift(raw[u]);
            if (!p.has) {
                p.val = x;
                p.pos = j;
                p.has = true;
            } else if constexpr (OP == JG_MIN || OP == JG_MAX) {
                const bool keep = (OP == JG_MIN) ? (p.val < x) : (p.val > x);
                if (!keep) {
                    p.val = x;
                    p.pos = j;
                }
            } else {
                p.val = Red<OP, T, A>::combine(p.val, x);
            }
        }
    }
    return p;
}

template <typename T>
struct ArgPartial {
    T val;
    int64_t idx;    // local index, -1 = none
};

template <bool ISMIN, typename T>
__device__ __forceinline__ ArgPartial<T> arg_merge(ArgPartial<T> a, ArgPartial<T> b) {
    if (b.idx < 0) return a;
    if (a.idx < 0) return b;
    const bool a_better = ISMIN ? (a.val < b.val) : (a.val > b.val);
    const bool b_better = ISMIN ? (b.val < a.val) : (b.val > a.val);
    if (a_better) return a;
    if (b_better) return b;
    return a.idx < b.idx ? a : b;     // equal values: first occurrence
}

template <bool ISMIN, typename T>
__device__ __forceinline__ ArgPartial<T> arg_shuffle(ArgPartial<T> p) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ArgPartial<T> o;
        o.val = __shfl_xor_sync(0xffffffffu, p.val, d);
        o.idx = __shfl_xor_sync(0xffffffffu, p.idx, d);
        p = arg_merge<ISMIN, T>(p, o);
    }
    return p;
}

template <bool ISMIN, typename T, int U = 4>      // U independent loads in flight per thread (see strided_partial)
__device__ __forceinline__ ArgPartial<T> arg_strided(const T *__restrict__ content, int64_t a, int64_t b, int first,
                                                     int step) {
    ArgPartial<T> p;
    p.idx = -1;
    p.val = T(0);
#pragma unroll 1
    for (int64_t j0 = a + first; j0 < b; j0 += static_cast<int64_t>(U) * step) {
        T raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = j0 + static_cast<int64_t>(u) * step;
            if (j < b) raw[u] = __ldcs(content + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = j0 + static_cast<int64_t>(u) * step;
            if (j >= b) break;
            const T x = raw[u];
            if (is_nan(x)) continue;
            if (p.idx < 0 || (ISMIN ? (x < p.val) : (x > p.val))) {
                p.val = x;
                p.idx = j - a;
            }
        }
    }
    return p;
}

}  // namespace jg

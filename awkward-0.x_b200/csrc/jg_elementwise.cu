// jg_elementwise.cu -- the ufunc side of JaggedArray.__array_ufunc__ (jagged.py:944-1051): elementwise binary
// and unary operators on dense contents, with the second operand being an array of the same length, a scalar,
// or a flat per-event array broadcast through the events (jagged.py:1003-1037) WITHOUT materialising the
// broadcast; plus whole-array reductions of flat device arrays (global totals).
//
// Operands are converted to the compute type C on load (numpy's promotion is decided on the host), the operator
// is a compile-time functor, the result is stored as C or as bool (comparisons / predicates).
#include <math_constants.h>

#include "jg_ops.cuh"
#include "jg_tile.cuh"

namespace jg {

// ---- loading any supported dtype as C ---------------------------------------------------------------------------
template <typename C>
__device__ __forceinline__ C load_as(const void *p, int dt, int64_t i) {
    switch (dt) {
        case JG_F32: return static_cast<C>(__ldg(static_cast<const float *>(p) + i));
        case JG_F64: return static_cast<C>(__ldg(static_cast<const double *>(p) + i));
        case JG_I32: return static_cast<C>(__ldg(static_cast<const int32_t *>(p) + i));
        case JG_I64: return static_cast<C>(__ldg(static_cast<const int64_t *>(p) + i));
        case JG_I8: return static_cast<C>(__ldg(static_cast<const int8_t *>(p) + i));
        case JG_I16: return static_cast<C>(__ldg(static_cast<const int16_t *>(p) + i));
        case JG_U16: return static_cast<C>(__ldg(static_cast<const uint16_t *>(p) + i));
        case JG_U32: return static_cast<C>(__ldg(static_cast<const uint32_t *>(p) + i));
        case JG_U64: return static_cast<C>(__ldg(static_cast<const uint64_t *>(p) + i));
        default: return static_cast<C>(__ldg(static_cast<const uint8_t *>(p) + i));   // JG_B8 / JG_U8
    }
}


// ---- kernels --------------------------------------------------------------------------------------------------------
template <class F, typename C>
__global__ void __launch_bounds__(256)
binary_flat_kernel(const void *__restrict__ lhs, int lhs_dt, const void *__restrict__ rhs, int rhs_dt, int rhs_kind,
                   C scalar, int swap, typename F::Out *__restrict__ out, int64_t len) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const C a = load_as<C>(lhs, lhs_dt, i);
        const C b = (rhs_kind == JG_RHS_SCALAR) ? scalar : load_as<C>(rhs, rhs_dt, i);
        __stcs(out + i, swap ? F::apply(b, a) : F::apply(a, b));
    }
}

// Vectorised fast path: both operands already have the compute type C (or rhs is a scalar) and every pointer is
// 16-byte aligned.  Each thread owns EPT = 16 consecutive elements: 128-bit loads, one 128-bit store per 16 bytes
// of output (a comparison writes its 16 bool results with ONE STG.128 instead of 16 byte stores).
template <typename T>
union Vec16 {
    int4 raw;
    T v[16 / sizeof(T)];
};

template <typename T, int N>
__device__ __forceinline__ void load_pack(const T *p, T (&dst)[N]) {
    constexpr int PER = 16 / sizeof(T);          // elements per 128-bit access
    static_assert(N % PER == 0, "pack size");
#pragma unroll
    for (int k = 0; k < N / PER; ++k) {
        Vec16<T> q;
        q.raw = __ldcs(reinterpret_cast<const int4 *>(p) + k);
#pragma unroll
        for (int e = 0; e < PER; ++e) dst[k * PER + e] = q.v[e];
    }
}

template <typename T, int N>
__device__ __forceinline__ void store_pack(T *p, const T (&src)[N]) {
    if constexpr (N * sizeof(T) == 8) {          // 8 one-byte results: one 64-bit store
        union { uint2 raw; T v[N]; } q;
#pragma unroll
        for (int e = 0; e < N; ++e) q.v[e] = src[e];
        __stcs(reinterpret_cast<uint2 *>(p), q.raw);
        return;
    }
    constexpr int PER = 16 / sizeof(T);
#pragma unroll
    for (int k = 0; k < N / PER; ++k) {
        Vec16<T> q;
#pragma unroll
        for (int e = 0; e < PER; ++e) q.v[e] = src[k * PER + e];
        __stcs(reinterpret_cast<int4 *>(p) + k, q.raw);
    }
}

template <class F, typename C>
__global__ void __launch_bounds__(256)
binary_vec_kernel(const C *__restrict__ lhs, const C *__restrict__ rhs, bool rhs_scalar, C scalar, int swap,
                  typename F::Out *__restrict__ out, int64_t nchunks) {
    using O = typename F::Out;
    constexpr int EPT = sizeof(C) == 8 ? 8 : 16;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < nchunks; c += stride) {
        C a[EPT], b[EPT];
        O r[EPT];
        load_pack<C, EPT>(lhs + c * EPT, a);
        if (rhs_scalar) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) b[e] = scalar;
        } else {
            load_pack<C, EPT>(rhs + c * EPT, b);
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) r[e] = swap ? F::apply(b[e], a[e]) : F::apply(a[e], b[e]);
        store_pack<O, EPT>(out + c * EPT, r);
    }
}

template <class F, typename C>
__global__ void __launch_bounds__(256)
unary_vec_kernel(const C *__restrict__ in, typename F::Out *__restrict__ out, int64_t nchunks) {
    using O = typename F::Out;
    constexpr int EPT = sizeof(C) == 8 ? 8 : 16;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < nchunks; c += stride) {
        C a[EPT];
        O r[EPT];
        load_pack<C, EPT>(in + c * EPT, a);
#pragma unroll
        for (int e = 0; e < EPT; ++e) r[e] = F::apply(a[e]);
        store_pack<O, EPT>(out + c * EPT, r);
    }
}

template <typename C> struct DtypeOf;
template <> struct DtypeOf<float> { static constexpr int value = JG_F32; };
template <> struct DtypeOf<double> { static constexpr int value = JG_F64; };
template <> struct DtypeOf<int32_t> { static constexpr int value = JG_I32; };
template <> struct DtypeOf<int64_t> { static constexpr int value = JG_I64; };
template <> struct DtypeOf<uint8_t> { static constexpr int value = JG_U8; };
template <> struct DtypeOf<uint64_t> { static constexpr int value = JG_U64; };

template <typename C>
static inline bool same_type(int dt) {
    return dt == DtypeOf<C>::value || (sizeof(C) == 1 && (dt == JG_B8 || dt == JG_U8));
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// rhs is flat[N]; lhs/out are dense contents starting at offsets[0].  The flat values of the tile's events and the
// owner of every position (scatter + max-scan table, jg_tile.cuh) are kept in shared memory: the broadcast of
// jagged.py:1003-1037 is never materialised.
template <class F, typename C, bool TYPED>      // TYPED: both operands already have the compute type (no per-element dtype switch)
__global__ void __launch_bounds__(kTileThreads, 8)
binary_parent_kernel(const void *__restrict__ lhs, int lhs_dt, const void *__restrict__ flat, int rhs_dt, int swap,
                     typename F::Out *__restrict__ out, const int64_t *__restrict__ offsets, int64_t n) {
    // (an owner table of 8192 positions: tiles of mean counts up to ~14 fit it; a longer run -- long events -- is swept
    // in chunks, one table each.  Round 1 found the owner of every element of such a tile by a binary search, with an
    // untyped load per operand: every tile of an array with a mean count above 8 took that path.)
    constexpr int KM = 8192;
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) uint16_t s_own[KM + 128];
    __shared__ C s_flat[kTileEvents];
    EventTile tile;
    tile.load(s_off, offsets, n, blockIdx.x);
    const int64_t o0 = __ldg(offsets);
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const int64_t len = e1 - e0;
    if (len <= 0) return;
#pragma unroll 1
    for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads)
        s_flat[ev] = TYPED ? __ldg(static_cast<const C *>(flat) + tile.event0 + ev) : load_as<C>(flat, rhs_dt, tile.event0 + ev);
#pragma unroll 1
    for (int64_t cs = 0; cs < len; cs += KM) {
        const int L = static_cast<int>(min(static_cast<int64_t>(KM), len - cs));
        if (len <= KM) build_owner_table(s_own, tile, L);         // (both end with a barrier; the first one publishes s_flat)
        else build_owner_range(s_own, tile, cs, L);
        const int64_t d0 = e0 - o0 + cs;
        // 4 independent loads in flight per thread (the sweep is otherwise one dependent load per iteration)
#pragma unroll 1
        for (int p = threadIdx.x; p < L; p += 4 * kTileThreads) {
            C a[4];
            int ev[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = p + u * kTileThreads;
                if (q < L) {
                    a[u] = TYPED ? __ldcs(static_cast<const C *>(lhs) + d0 + q) : load_as<C>(lhs, lhs_dt, d0 + q);
                    ev[u] = s_own[q];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = p + u * kTileThreads;
                if (q < L) {
                    const C b = s_flat[ev[u]];
                    __stcs(out + d0 + q, swap ? F::apply(b, a[u]) : F::apply(a[u], b));
                }
            }
        }
        if (len > KM) __syncthreads();                            // the table is rewritten by the next chunk
    }
}

static inline int ew_grid(int64_t work) {
    int64_t blocks = ceil_div(work, 256);
    int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

struct BinArgs {
    const void *lhs; int lhs_dt; const void *rhs; int rhs_dt; int rhs_kind; double sf; int64_t si; int swap;
    void *out; int64_t len; const int64_t *offsets; int64_t n; cudaStream_t s;
};

template <class F, typename C>
static int launch_binary(const BinArgs &a) {
    using O = typename F::Out;
    if (a.rhs_kind == JG_RHS_PARENT) {
        if (a.n <= 0) return 0;
        if (same_type<C>(a.lhs_dt) && same_type<C>(a.rhs_dt))
            binary_parent_kernel<F, C, true><<<tile_grid(a.n), kTileThreads, 0, a.s>>>(
                a.lhs, a.lhs_dt, a.rhs, a.rhs_dt, a.swap, static_cast<O *>(a.out), a.offsets, a.n);
        else
            binary_parent_kernel<F, C, false><<<tile_grid(a.n), kTileThreads, 0, a.s>>>(
                a.lhs, a.lhs_dt, a.rhs, a.rhs_dt, a.swap, static_cast<O *>(a.out), a.offsets, a.n);
        return check_launch("binary_parent_kernel");
    }
    if (a.len <= 0) return 0;
    C scalar;
    if constexpr (IsFp<C>::value) scalar = static_cast<C>(a.sf);
    else scalar = static_cast<C>(a.si);
    const bool rhs_scalar = a.rhs_kind == JG_RHS_SCALAR;
    int64_t done = 0;
    if (same_type<C>(a.lhs_dt) && (rhs_scalar || same_type<C>(a.rhs_dt)) && aligned16(a.lhs) && aligned16(a.out) &&
        (rhs_scalar || aligned16(a.rhs)) && a.len >= 16) {
        constexpr int EPT = sizeof(C) == 8 ? 8 : 16;
        const int64_t nchunks = a.len / EPT;
        binary_vec_kernel<F, C><<<ew_grid(nchunks), 256, 0, a.s>>>(static_cast<const C *>(a.lhs),
                                                                  static_cast<const C *>(a.rhs), rhs_scalar, scalar,
                                                                  a.swap, static_cast<O *>(a.out), nchunks);
        done = nchunks * EPT;
        if (done == a.len) return check_launch("binary_vec_kernel");
    }
    // scalar tail / unaligned / mixed-dtype operands
    const char *lp = static_cast<const char *>(a.lhs) + done * dtype_size(a.lhs_dt);
    const char *rp = rhs_scalar ? nullptr : static_cast<const char *>(a.rhs) + done * dtype_size(a.rhs_dt);
    binary_flat_kernel<F, C><<<ew_grid(a.len - done), 256, 0, a.s>>>(lp, a.lhs_dt, rp, a.rhs_dt, a.rhs_kind, scalar,
                                                                    a.swap, static_cast<O *>(a.out) + done,
                                                                    a.len - done);
    return check_launch("binary_flat_kernel");
}

template <typename C>
static int dispatch_binary_arith(int op, const BinArgs &a) {
    switch (op) {
        case JG_OP_ADD: return launch_binary<BinaryOp<JG_OP_ADD, C>, C>(a);
        case JG_OP_SUB: return launch_binary<BinaryOp<JG_OP_SUB, C>, C>(a);
        case JG_OP_MUL: return launch_binary<BinaryOp<JG_OP_MUL, C>, C>(a);
        case JG_OP_FLOORDIV: return launch_binary<BinaryOp<JG_OP_FLOORDIV, C>, C>(a);
        case JG_OP_MOD: return launch_binary<BinaryOp<JG_OP_MOD, C>, C>(a);
        case JG_OP_POW: return launch_binary<BinaryOp<JG_OP_POW, C>, C>(a);
        case JG_OP_MAXIMUM: return launch_binary<BinaryOp<JG_OP_MAXIMUM, C>, C>(a);
        case JG_OP_MINIMUM: return launch_binary<BinaryOp<JG_OP_MINIMUM, C>, C>(a);
    }
    if constexpr (IsFp<C>::value) {
        switch (op) {
            case JG_OP_DIV: return launch_binary<BinaryOp<JG_OP_DIV, C>, C>(a);
            case JG_OP_ARCTAN2: return launch_binary<BinaryOp<JG_OP_ARCTAN2, C>, C>(a);
            case JG_OP_HYPOT: return launch_binary<BinaryOp<JG_OP_HYPOT, C>, C>(a);
        }
    } else {
        switch (op) {
            case JG_OP_AND: return launch_binary<BinaryOp<JG_OP_AND, C>, C>(a);
            case JG_OP_OR: return launch_binary<BinaryOp<JG_OP_OR, C>, C>(a);
            case JG_OP_XOR: return launch_binary<BinaryOp<JG_OP_XOR, C>, C>(a);
        }
    }
    set_error("jg_binary: op %d not available for this compute dtype", op);
    return -2;
}

template <typename C>
static int dispatch_binary_compare(int op, const BinArgs &a) {
    switch (op) {
        case JG_OP_GT: return launch_binary<CompareOp<JG_OP_GT, C>, C>(a);
        case JG_OP_GE: return launch_binary<CompareOp<JG_OP_GE, C>, C>(a);
        case JG_OP_LT: return launch_binary<CompareOp<JG_OP_LT, C>, C>(a);
        case JG_OP_LE: return launch_binary<CompareOp<JG_OP_LE, C>, C>(a);
        case JG_OP_EQ: return launch_binary<CompareOp<JG_OP_EQ, C>, C>(a);
        case JG_OP_NE: return launch_binary<CompareOp<JG_OP_NE, C>, C>(a);
    }
    return -2;
}


template <class F, typename C>
__global__ void __launch_bounds__(256)
unary_kernel(const void *__restrict__ in, int in_dt, typename F::Out *__restrict__ out, int64_t len) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
        __stcs(out + i, F::apply(load_as<C>(in, in_dt, i)));
}

template <class F, typename C>
static int launch_unary(const void *in, int in_dt, void *out, int64_t len, cudaStream_t s) {
    if (len <= 0) return 0;
    using O = typename F::Out;
    int64_t done = 0;
    if (same_type<C>(in_dt) && aligned16(in) && aligned16(out) && len >= 16) {
        constexpr int EPT = sizeof(C) == 8 ? 8 : 16;
        const int64_t nchunks = len / EPT;
        unary_vec_kernel<F, C><<<ew_grid(nchunks), 256, 0, s>>>(static_cast<const C *>(in), static_cast<O *>(out), nchunks);
        done = nchunks * EPT;
        if (done == len) return check_launch("unary_vec_kernel");
    }
    unary_kernel<F, C><<<ew_grid(len - done), 256, 0, s>>>(static_cast<const char *>(in) + done * dtype_size(in_dt), in_dt,
                                                          static_cast<O *>(out) + done, len - done);
    return check_launch("unary_kernel");
}

template <typename C>
static int dispatch_unary(int op, const void *in, int in_dt, void *out, int out_dt, int64_t len, cudaStream_t s) {
#define JG_U(OPC) case OPC: return launch_unary<UnaryOp<OPC, C>, C>(in, in_dt, out, len, s)
    switch (op) {
        JG_U(JG_OP_NEG); JG_U(JG_OP_ABS); JG_U(JG_OP_SQUARE); JG_U(JG_OP_CAST); JG_U(JG_OP_SIGN);
    }
    if constexpr (IsFp<C>::value) {
        switch (op) {
            JG_U(JG_OP_SQRT); JG_U(JG_OP_EXP); JG_U(JG_OP_LOG); JG_U(JG_OP_SIN); JG_U(JG_OP_COS); JG_U(JG_OP_TAN);
            JG_U(JG_OP_SINH); JG_U(JG_OP_COSH); JG_U(JG_OP_TANH); JG_U(JG_OP_ARCSINH); JG_U(JG_OP_FLOOR);
            JG_U(JG_OP_CEIL); JG_U(JG_OP_RINT); JG_U(JG_OP_LOG10); JG_U(JG_OP_EXP2); JG_U(JG_OP_ARCSIN);
            JG_U(JG_OP_ARCCOS); JG_U(JG_OP_ARCTAN); JG_U(JG_OP_LOG2); JG_U(JG_OP_LOG1P); JG_U(JG_OP_EXPM1);
            JG_U(JG_OP_RECIPROCAL);
            case JG_OP_ISNAN: return launch_unary<PredicateOp<JG_OP_ISNAN, C>, C>(in, in_dt, out, len, s);
            case JG_OP_ISINF: return launch_unary<PredicateOp<JG_OP_ISINF, C>, C>(in, in_dt, out, len, s);
            case JG_OP_ISFINITE: return launch_unary<PredicateOp<JG_OP_ISFINITE, C>, C>(in, in_dt, out, len, s);
        }
    } else {
        if (op == JG_OP_NOT) {
            if (out_dt == JG_B8) return launch_unary<PredicateOp<JG_OP_NOT, C>, C>(in, in_dt, out, len, s);
            return launch_unary<UnaryOp<JG_OP_NOT, C>, C>(in, in_dt, out, len, s);
        }
    }
#undef JG_U
    set_error("jg_unary: op %d not available for this compute dtype", op);
    return -2;
}

// cast between any two supported dtypes
template <typename O>
__global__ void __launch_bounds__(256)
cast_kernel(const void *__restrict__ in, int in_dt, O *__restrict__ out, int64_t len, bool to_bool) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        if (to_bool) {
            const double v = load_as<double>(in, in_dt, i);
            out[i] = static_cast<O>(v != 0.0);     // NaN != 0 -> True, like numpy astype(bool)
        } else if (in_dt == JG_F32 || in_dt == JG_F64) {
            out[i] = static_cast<O>(load_as<double>(in, in_dt, i));
        } else if (in_dt == JG_U64) {
            out[i] = static_cast<O>(load_as<uint64_t>(in, in_dt, i));
        } else {
            out[i] = static_cast<O>(load_as<int64_t>(in, in_dt, i));
        }
    }
}

// ---- flat reductions (two stages: per-CTA partials in the workspace, then one CTA) ----------------------------------
template <int OP, typename A>
__device__ __forceinline__ A flat_identity() {
    if constexpr (OP == JG_SUM || OP == JG_COUNT_NONZERO) return A(0);
    else if constexpr (OP == JG_PROD) return A(1);
    else if constexpr (OP == JG_MIN) {
        if constexpr (IsFp<A>::value) return A(CUDART_INF);
        else if constexpr (sizeof(A) == 8) return static_cast<A>(INT64_MAX);
        else return static_cast<A>(INT32_MAX);
    } else if constexpr (OP == JG_MAX) {
        if constexpr (IsFp<A>::value) return A(-CUDART_INF);
        else if constexpr (sizeof(A) == 8) return static_cast<A>(INT64_MIN);
        else return static_cast<A>(INT32_MIN);
    } else if constexpr (OP == JG_ANY) return A(0);
    else return A(1);
}

template <int OP, typename A>
__device__ __forceinline__ A flat_combine(A a, A b) {
    if constexpr (OP == JG_SUM || OP == JG_COUNT_NONZERO) return a + b;
    else if constexpr (OP == JG_PROD) return a * b;
    else if constexpr (OP == JG_MIN) {
        if constexpr (IsFp<A>::value) return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));   // numpy: NaN propagates
        else return a < b ? a : b;
    } else if constexpr (OP == JG_MAX) {
        if constexpr (IsFp<A>::value) return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
        else return a > b ? a : b;
    } else if constexpr (OP == JG_ANY) return A(a | b);
    else return A(a & b);
}

template <int OP, typename A>
__device__ __forceinline__ A block_reduce(A v, A *sh) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = flat_combine<OP, A>(v, __shfl_xor_sync(0xffffffffu, v, d));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < 8 ? sh[lane] : flat_identity<OP, A>();
#pragma unroll
        for (int d = 4; d > 0; d >>= 1) v = flat_combine<OP, A>(v, __shfl_xor_sync(0xffffffffu, v, d));
    }
    return v;
}

template <int OP, typename A>
__global__ void __launch_bounds__(256)
flat_reduce_stage1(const void *__restrict__ in, int dt, int64_t len, A *__restrict__ partial) {
    __shared__ A sh[8];
    A acc = flat_identity<OP, A>();
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        A x = load_as<A>(in, dt, i);
        if constexpr (OP == JG_ANY || OP == JG_ALL || OP == JG_COUNT_NONZERO) x = A(load_as<double>(in, dt, i) != 0.0);
        acc = flat_combine<OP, A>(acc, x);
    }
    acc = block_reduce<OP, A>(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

template <int OP, typename A>
__global__ void __launch_bounds__(256)
flat_reduce_stage2(const A *__restrict__ partial, int nparts, A *__restrict__ out) {
    __shared__ A sh[8];
    A acc = flat_identity<OP, A>();
    for (int i = threadIdx.x; i < nparts; i += 256) acc = flat_combine<OP, A>(acc, partial[i]);
    acc = block_reduce<OP, A>(acc, sh);
    if (threadIdx.x == 0) *out = acc;
}

template <int OP, typename A>
static int launch_flat_reduce(const void *in, int dt, int64_t len, void *out, void *ws, size_t ws_bytes, cudaStream_t s) {
    int parts = ew_grid(len);
    if (parts > 1024) parts = 1024;
    JG_REQUIRE(ws && ws_bytes >= static_cast<size_t>(parts) * sizeof(A), "jg_flat_reduce: workspace too small");
    flat_reduce_stage1<OP, A><<<parts, 256, 0, s>>>(in, dt, len, static_cast<A *>(ws));
    flat_reduce_stage2<OP, A><<<1, 256, 0, s>>>(static_cast<const A *>(ws), parts, static_cast<A *>(out));
    return check_launch("flat_reduce");
}

template <typename A>
static int dispatch_flat_reduce(int op, const void *in, int dt, int64_t len, void *out, void *ws, size_t wsb, cudaStream_t s) {
    switch (op) {
        case JG_SUM: return launch_flat_reduce<JG_SUM, A>(in, dt, len, out, ws, wsb, s);
        case JG_PROD: return launch_flat_reduce<JG_PROD, A>(in, dt, len, out, ws, wsb, s);
        case JG_MIN: return launch_flat_reduce<JG_MIN, A>(in, dt, len, out, ws, wsb, s);
        case JG_MAX: return launch_flat_reduce<JG_MAX, A>(in, dt, len, out, ws, wsb, s);
    }
    return -2;
}

// Arrow boolean buffers are bit-packed, least-significant bit first (arrow.py:380-395 unpacks them with
// numpy.unpackbits + a byte-wise reversal): out[i] = bit (bit_offset + i) of `bits`.  8 outputs per thread.
__global__ void __launch_bounds__(256)
unpack_bits_kernel(const uint8_t *__restrict__ bits, int64_t bit_offset, int64_t n, uint8_t *__restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * 256 * 8;
    for (int64_t i0 = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) * 8; i0 < n; i0 += stride) {
        const int64_t b = bit_offset + i0;
        // two source bytes cover 8 consecutive bits at any phase
        const uint32_t lo = bits[b >> 3];
        const int64_t last = min(b + 7, bit_offset + n - 1);          // last bit this thread needs
        const uint32_t hi = ((last >> 3) != (b >> 3)) ? bits[last >> 3] : 0u;
        const uint32_t w = (lo | (hi << 8)) >> (b & 7);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k < n) out[i0 + k] = (w >> k) & 1u;
    }
}

// fillna (masked.py:401-406 via base.py:645-654): masked slots AND NaN take the fill value
template <typename T>
__global__ void __launch_bounds__(256)
fill_masked_kernel(const T *__restrict__ x, const uint8_t *__restrict__ mask, int masked_when, T value,
                   T *__restrict__ out, int64_t len) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < len; i += stride) {
        const T v = x[i];
        bool fill = mask != nullptr && ((mask[i] != 0) == (masked_when != 0));
        if constexpr (sizeof(T) >= 4 && (T(0.5) != T(0))) fill |= (v != v);
        out[i] = fill ? value : v;
    }
}

template <typename T>
static int launch_fill_masked(const void *x, const uint8_t *mask, int masked_when, T value, void *out, int64_t len,
                              cudaStream_t s) {
    fill_masked_kernel<T><<<ew_grid(len), 256, 0, s>>>(static_cast<const T *>(x), mask, masked_when, value,
                                                       static_cast<T *>(out), len);
    return check_launch("fill_masked_kernel");
}

}  // namespace jg

using namespace jg;

extern "C" {

int jg_unpack_bits(const uint8_t *bits, int64_t bit_offset, int64_t n, uint8_t *out, void *stream) {
    if (n <= 0) return 0;
    JG_REQUIRE(bits && out && bit_offset >= 0, "jg_unpack_bits: bad arguments");
    unpack_bits_kernel<<<ew_grid((n + 7) / 8), 256, 0, as_stream(stream)>>>(bits, bit_offset, n, out);
    return check_launch("unpack_bits_kernel");
}

int jg_fill_masked(const void *content, int dtype, const uint8_t *mask, int masked_when, double value_f64,
                   int64_t value_i64, void *out, int64_t len, void *stream) {
    if (len <= 0) return 0;
    JG_REQUIRE(content && out, "jg_fill_masked: bad arguments");
    cudaStream_t s = as_stream(stream);
    switch (dtype) {
        case JG_F32: return launch_fill_masked<float>(content, mask, masked_when, static_cast<float>(value_f64), out, len, s);
        case JG_F64: return launch_fill_masked<double>(content, mask, masked_when, value_f64, out, len, s);
        case JG_I32: return launch_fill_masked<int32_t>(content, mask, masked_when, static_cast<int32_t>(value_i64), out, len, s);
        case JG_I64: return launch_fill_masked<int64_t>(content, mask, masked_when, value_i64, out, len, s);
        case JG_U8: return launch_fill_masked<uint8_t>(content, mask, masked_when, static_cast<uint8_t>(value_i64), out, len, s);
        case JG_B8: return launch_fill_masked<uint8_t>(content, mask, masked_when, value_i64 != 0 ? 1 : 0, out, len, s);
        case JG_I8: return launch_fill_masked<int8_t>(content, mask, masked_when, static_cast<int8_t>(value_i64), out, len, s);
        case JG_I16: return launch_fill_masked<int16_t>(content, mask, masked_when, static_cast<int16_t>(value_i64), out, len, s);
        case JG_U16: return launch_fill_masked<uint16_t>(content, mask, masked_when, static_cast<uint16_t>(value_i64), out, len, s);
        case JG_U32: return launch_fill_masked<uint32_t>(content, mask, masked_when, static_cast<uint32_t>(value_i64), out, len, s);
        case JG_U64: return launch_fill_masked<uint64_t>(content, mask, masked_when, static_cast<uint64_t>(value_i64), out, len, s);
    }
    set_error("jg_fill_masked: unsupported dtype %d", dtype);
    return -2;
}

int jg_binary(int op, const void *lhs, int lhs_dtype, const void *rhs, int rhs_dtype, int rhs_kind, double scalar_f64,
              int64_t scalar_i64, int swap, int compute_dtype, void *out, int out_dtype, int64_t len,
              const int64_t *offsets, int64_t n, void *stream) {
    BinArgs a{lhs, lhs_dtype, rhs, rhs_dtype, rhs_kind, scalar_f64, scalar_i64, swap, out, len, offsets, n, as_stream(stream)};
    const bool compare = op >= JG_OP_GT && op <= JG_OP_NE;
    JG_REQUIRE(compare ? out_dtype == JG_B8 : out_dtype == compute_dtype,
               "jg_binary: out dtype %d does not match compute dtype %d for op %d", out_dtype, compute_dtype, op);
    int rc = -2;
    if (compare) {
        switch (compute_dtype) {
            case JG_F32: rc = dispatch_binary_compare<float>(op, a); break;
            case JG_F64: rc = dispatch_binary_compare<double>(op, a); break;
            case JG_I32: rc = dispatch_binary_compare<int32_t>(op, a); break;
            case JG_I64: rc = dispatch_binary_compare<int64_t>(op, a); break;
            case JG_U8:
            case JG_B8: rc = dispatch_binary_compare<uint8_t>(op, a); break;
            case JG_U64: rc = dispatch_binary_compare<uint64_t>(op, a); break;
        }
    } else {
        switch (compute_dtype) {
            case JG_F32: rc = dispatch_binary_arith<float>(op, a); break;
            case JG_F64: rc = dispatch_binary_arith<double>(op, a); break;
            case JG_I32: rc = dispatch_binary_arith<int32_t>(op, a); break;
            case JG_I64: rc = dispatch_binary_arith<int64_t>(op, a); break;
            case JG_U8: rc = dispatch_binary_arith<uint8_t>(op, a); break;
            case JG_B8:
                // numpy on bools: + is or, * is and, maximum/minimum are or/and; & | ^ are logical
                switch (op) {
                    case JG_OP_ADD: case JG_OP_OR: case JG_OP_MAXIMUM:
                        rc = launch_binary<BinaryOp<JG_OP_OR, uint8_t>, uint8_t>(a); break;
                    case JG_OP_MUL: case JG_OP_AND: case JG_OP_MINIMUM:
                        rc = launch_binary<BinaryOp<JG_OP_AND, uint8_t>, uint8_t>(a); break;
                    case JG_OP_XOR: rc = launch_binary<BinaryOp<JG_OP_XOR, uint8_t>, uint8_t>(a); break;
                }
                break;
        }
    }
    if (rc == -2 && jg_last_error()[0] == 0) set_error("jg_binary: unsupported op %d / compute dtype %d", op, compute_dtype);
    return rc;
}

int jg_unary(int op, const void *in, int in_dtype, int compute_dtype, void *out, int out_dtype, int64_t len,
             void *stream) {
    cudaStream_t s = as_stream(stream);
    if (op == JG_OP_CAST) {
        if (len <= 0) return 0;
        const int g = ew_grid(len);
        switch (out_dtype) {
            case JG_F32: cast_kernel<float><<<g, 256, 0, s>>>(in, in_dtype, static_cast<float *>(out), len, false); break;
            case JG_F64: cast_kernel<double><<<g, 256, 0, s>>>(in, in_dtype, static_cast<double *>(out), len, false); break;
            case JG_I32: cast_kernel<int32_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<int32_t *>(out), len, false); break;
            case JG_I64: cast_kernel<int64_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<int64_t *>(out), len, false); break;
            case JG_U8: cast_kernel<uint8_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<uint8_t *>(out), len, false); break;
            case JG_B8: cast_kernel<uint8_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<uint8_t *>(out), len, true); break;
            case JG_I8: cast_kernel<int8_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<int8_t *>(out), len, false); break;
            case JG_I16: cast_kernel<int16_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<int16_t *>(out), len, false); break;
            case JG_U16: cast_kernel<uint16_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<uint16_t *>(out), len, false); break;
            case JG_U32: cast_kernel<uint32_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<uint32_t *>(out), len, false); break;
            case JG_U64: cast_kernel<uint64_t><<<g, 256, 0, s>>>(in, in_dtype, static_cast<uint64_t *>(out), len, false); break;
            default: set_error("jg_unary: cast to dtype %d unsupported", out_dtype); return -2;
        }
        return check_launch("cast_kernel");
    }
    switch (compute_dtype) {
        case JG_F32: return dispatch_unary<float>(op, in, in_dtype, out, out_dtype, len, s);
        case JG_F64: return dispatch_unary<double>(op, in, in_dtype, out, out_dtype, len, s);
        case JG_I32: return dispatch_unary<int32_t>(op, in, in_dtype, out, out_dtype, len, s);
        case JG_I64: return dispatch_unary<int64_t>(op, in, in_dtype, out, out_dtype, len, s);
        case JG_U8:
        case JG_B8: return dispatch_unary<uint8_t>(op, in, in_dtype, out, out_dtype, len, s);
    }
    set_error("jg_unary: unsupported compute dtype %d", compute_dtype);
    return -2;
}

int jg_flat_reduce(const void *in, int dtype, int64_t len, int op, void *out, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = as_stream(stream);
    if (op == JG_ANY) return launch_flat_reduce<JG_ANY, uint8_t>(in, dtype, len, out, ws, ws_bytes, s);
    if (op == JG_ALL) return launch_flat_reduce<JG_ALL, uint8_t>(in, dtype, len, out, ws, ws_bytes, s);
    if (op == JG_COUNT_NONZERO) return launch_flat_reduce<JG_COUNT_NONZERO, int64_t>(in, dtype, len, out, ws, ws_bytes, s);
    int rc = -2;
    switch (dtype) {
        case JG_F32: rc = dispatch_flat_reduce<float>(op, in, dtype, len, out, ws, ws_bytes, s); break;
        case JG_F64: rc = dispatch_flat_reduce<double>(op, in, dtype, len, out, ws, ws_bytes, s); break;
        case JG_I32:
            // numpy: sum/prod of int32 -> int64; min/max keep int32
            if (op == JG_SUM || op == JG_PROD) rc = dispatch_flat_reduce<int64_t>(op, in, dtype, len, out, ws, ws_bytes, s);
            else rc = dispatch_flat_reduce<int32_t>(op, in, dtype, len, out, ws, ws_bytes, s);
            break;
        case JG_I64: rc = dispatch_flat_reduce<int64_t>(op, in, dtype, len, out, ws, ws_bytes, s); break;
        case JG_B8:
        case JG_U8:
            if (op == JG_SUM || op == JG_PROD) rc = dispatch_flat_reduce<int64_t>(op, in, dtype, len, out, ws, ws_bytes, s);
            break;
    }
    if (rc == -2) set_error("jg_flat_reduce: unsupported dtype %d / op %d", dtype, op);
    return rc;
}

}  // extern "C"

// jg_ops.cuh -- the elementwise operator functors (numpy ufunc semantics) shared by the eager kernels of
// jg_elementwise.cu and the fused expression evaluator of jg_eval.cu, so that both paths round identically.
#pragma once

#include <math_constants.h>

#include "jg_common.cuh"

namespace jg {

template <typename T> struct IsFp { static constexpr bool value = false; };
template <> struct IsFp<float> { static constexpr bool value = true; };
template <> struct IsFp<double> { static constexpr bool value = true; };

struct Bool8 { uint8_t v; };   // tag type: compute in "bool" (logical semantics)

// ---- binary functors ----------------------------------------------------------------------------------------------
template <int OP, typename C>
struct BinaryOp {
    using Out = C;
    __device__ static __forceinline__ C apply(C a, C b) {
        if constexpr (OP == JG_OP_ADD) return a + b;
        else if constexpr (OP == JG_OP_SUB) return a - b;
        else if constexpr (OP == JG_OP_MUL) return a * b;
        else if constexpr (OP == JG_OP_DIV) return a / b;
        else if constexpr (OP == JG_OP_MAXIMUM) {
            if constexpr (IsFp<C>::value) return (a != a) ? a : ((b != b) ? b : (a >= b ? a : b));
            else return a >= b ? a : b;
        } else if constexpr (OP == JG_OP_MINIMUM) {
            if constexpr (IsFp<C>::value) return (a != a) ? a : ((b != b) ? b : (a <= b ? a : b));
            else return a <= b ? a : b;
        } else if constexpr (OP == JG_OP_FLOORDIV) {
            if constexpr (IsFp<C>::value) {
                // numpy npy_divmod
                if (b == C(0)) return a / b;
                C mod = fmod(a, b);
                C div = (a - mod) / b;
                if (mod != C(0) && ((b < C(0)) != (mod < C(0)))) div -= C(1);
                if (div != C(0)) {
                    C fl = floor(div);
                    if (div - fl > C(0.5)) fl += C(1);
                    return fl;
                }
                return copysign(C(0), a / b);
            } else {
                if (b == C(0)) return C(0);
                if ((C(-1) < C(0)) && b == C(-1)) return C(0) - a;
                C q = a / b;
                if ((a % b != C(0)) && ((a < C(0)) != (b < C(0)))) q -= C(1);
                return q;
            }
        } else if constexpr (OP == JG_OP_MOD) {
            if constexpr (IsFp<C>::value) {
                if (b == C(0)) return fmod(a, b);
                C mod = fmod(a, b);
                if (mod != C(0)) {
                    if ((b < C(0)) != (mod < C(0))) mod += b;
                } else {
                    mod = copysign(C(0), b);
                }
                return mod;
            } else {
                if (b == C(0)) return C(0);
                if ((C(-1) < C(0)) && b == C(-1)) return C(0);
                C r = a % b;
                if (r != C(0) && ((r < C(0)) != (b < C(0)))) r += b;
                return r;
            }
        } else if constexpr (OP == JG_OP_POW) {
            if constexpr (IsFp<C>::value) return pow(a, b);
            else {
                C result = C(1), base = a;
                C e = b;
                if (e < C(0)) return C(0);
                while (e > C(0)) {
                    if (e & C(1)) result *= base;
                    base *= base;
                    e >>= 1;
                }
                return result;
            }
        } else if constexpr (OP == JG_OP_ARCTAN2) return atan2(a, b);
        else if constexpr (OP == JG_OP_HYPOT) return hypot(a, b);
        else if constexpr (OP == JG_OP_AND) return a & b;
        else if constexpr (OP == JG_OP_OR) return a | b;
        else return a ^ b;   // JG_OP_XOR
    }
};

template <int OP, typename C>
struct CompareOp {
    using Out = uint8_t;
    __device__ static __forceinline__ uint8_t apply(C a, C b) {
        if constexpr (OP == JG_OP_GT) return a > b;
        else if constexpr (OP == JG_OP_GE) return a >= b;
        else if constexpr (OP == JG_OP_LT) return a < b;
        else if constexpr (OP == JG_OP_LE) return a <= b;
        else if constexpr (OP == JG_OP_EQ) return a == b;
        else return a != b;
    }
};

// ---- unary ----------------------------------------------------------------------------------------------------------
template <int OP, typename C>
struct UnaryOp {
    using Out = C;
    __device__ static __forceinline__ C apply(C x) {
        if constexpr (OP == JG_OP_NEG) {
            if constexpr (IsFp<C>::value) return -x;      // numpy.negative(0.0) is -0.0
            else return C(0) - x;
        }
        else if constexpr (OP == JG_OP_ABS) {
            if constexpr (IsFp<C>::value) return fabs(x);
            else return x < C(0) ? C(0) - x : x;
        } else if constexpr (OP == JG_OP_SQUARE) return x * x;
        else if constexpr (OP == JG_OP_CAST) return x;
        else if constexpr (OP == JG_OP_SIGN) {
            if constexpr (IsFp<C>::value) return (x != x) ? x : C((x > C(0)) - (x < C(0)));
            else return C((x > C(0)) - (x < C(0)));
        } else if constexpr (OP == JG_OP_NOT) return ~x;
        else if constexpr (OP == JG_OP_SQRT) return sqrt(x);
        else if constexpr (OP == JG_OP_EXP) return exp(x);
        else if constexpr (OP == JG_OP_LOG) return log(x);
        else if constexpr (OP == JG_OP_SIN) return sin(x);
        else if constexpr (OP == JG_OP_COS) return cos(x);
        else if constexpr (OP == JG_OP_TAN) return tan(x);
        else if constexpr (OP == JG_OP_SINH) return sinh(x);
        else if constexpr (OP == JG_OP_COSH) return cosh(x);
        else if constexpr (OP == JG_OP_TANH) return tanh(x);
        else if constexpr (OP == JG_OP_ARCSINH) return asinh(x);
        else if constexpr (OP == JG_OP_FLOOR) return floor(x);
        else if constexpr (OP == JG_OP_CEIL) return ceil(x);
        else if constexpr (OP == JG_OP_RINT) return rint(x);
        else if constexpr (OP == JG_OP_LOG10) return log10(x);
        else if constexpr (OP == JG_OP_EXP2) return exp2(x);
        else if constexpr (OP == JG_OP_ARCSIN) return asin(x);
        else if constexpr (OP == JG_OP_ARCCOS) return acos(x);
        else if constexpr (OP == JG_OP_ARCTAN) return atan(x);
        else if constexpr (OP == JG_OP_LOG2) return log2(x);
        else if constexpr (OP == JG_OP_LOG1P) return log1p(x);
        else if constexpr (OP == JG_OP_EXPM1) return expm1(x);
        else return C(1) / x;   // JG_OP_RECIPROCAL
    }
};

template <int OP, typename C>
struct PredicateOp {
    using Out = uint8_t;
    __device__ static __forceinline__ uint8_t apply(C x) {
        if constexpr (OP == JG_OP_ISNAN) return x != x;
        else if constexpr (OP == JG_OP_ISINF) return isinf(x);
        else if constexpr (OP == JG_OP_ISFINITE) return isfinite(x);
        else return x == C(0);   // JG_OP_NOT on bool-like: logical not
    }
};

}  // namespace jg

// jg_compact.cuh -- what the selection kernels share (jg_select.cu: event tiles, jg_elem.cu: element tiles): the
// selectors and the two passes over a staged run, written for the instruction-issue limit these kernels run into
// (profiles/r2_sweep_kernels_ncu_full.csv: the compaction of a[a > c] was 83 % issue-active at 64 % of the DRAM peak).
//
// The staged run is walked in ROWS of 32 positions.  The caller pads the selector's run to whole rows with
// SEL::never(), an item that is never selected, so neither pass tests a bound; both loops are unrolled by hand (two
// rows per step and one odd row: nvcc's own unrolling adds 8/4/2/1 remainder cascades that every warp walks through);
// a store's address is one 32 x 32 -> 64 multiply-add on a base held in a register pair.  A row of the compaction
// costs 16 instructions (28 before), a row of the counting pass 7.
#pragma once

#include <cstring>
#include <type_traits>

#include "jg_common.cuh"

namespace jg {

// What decides whether a position is selected: a byte mask (a[mask]) or a comparison of a float column with a scalar
// evaluated on the fly (a[a > c]: the mask is never written nor read).  `x OP c` is brought to one of two forms on
// the host: flip negates both sides (x < c  ==  -x > -c, exact for every float, NaN stays unordered) and strict
// picks > or >= (a template parameter: the test is one XOR and one compare).
struct ByteMask {
    using Elem = uint8_t;
    __device__ __forceinline__ bool test(uint8_t v) const { return v != 0; }
    __device__ __forceinline__ int minus_one_if(uint8_t v) const { return v != 0 ? -1 : 0; }
    __device__ static __forceinline__ uint8_t never() { return 0; }       // an item that test() always rejects
};
template <typename K, bool STRICT>
struct FloatCompare {
    using Elem = K;
    using Bits = typename std::conditional<sizeof(K) == 4, uint32_t, uint64_t>::type;
    K c;             // already negated when flipping
    Bits flip;       // sign bit when the comparison is mirrored, else 0: one XOR per item, no select
    __device__ __forceinline__ K mirrored(K v) const {
        Bits b;
        memcpy(&b, &v, sizeof(K));
        b ^= flip;
        K y;
        memcpy(&y, &b, sizeof(K));
        return y;
    }
    __device__ __forceinline__ bool test(K v) const {
        const K y = mirrored(v);
        return STRICT ? (y > c) : (y >= c);
    }
    // -1 when selected, else 0 (PTX set): a counting loop subtracts it, two items per three-input add
    __device__ __forceinline__ int minus_one_if(K v) const {
        const K y = mirrored(v);
        int r;
        if constexpr (sizeof(K) == 4) {
            if (STRICT) asm("set.gt.s32.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(c));
            else asm("set.ge.s32.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(c));
        } else {
            if (STRICT) asm("set.gt.s32.f64 %0, %1, %2;" : "=r"(r) : "d"(y), "d"(c));
            else asm("set.ge.s32.f64 %0, %1, %2;" : "=r"(r) : "d"(y), "d"(c));
        }
        return r;
    }
    // NaN is unordered whatever the sign bit says: rejected by both comparisons, mirrored or not
    __device__ static __forceinline__ K never() {
        const Bits b = sizeof(K) == 4 ? static_cast<Bits>(0x7fc00000u) : static_cast<Bits>(0x7ff8000000000000ull);
        K y;
        memcpy(&y, &b, sizeof(K));
        return y;
    }
};

// out[slot] = v; `gbase` is a GLOBAL-space address kept opaque by the caller so that it stays in its register pair
template <typename V>
__device__ __forceinline__ void store_slot(uint64_t gbase, uint32_t slot, V v) {
    const uint64_t a = gbase + static_cast<uint64_t>(slot) * sizeof(V);
    if constexpr (sizeof(V) == 1) asm volatile("st.global.u8 [%0], %1;" ::"l"(a), "r"(static_cast<uint32_t>(v)));
    else if constexpr (sizeof(V) == 2) asm volatile("st.global.u16 [%0], %1;" ::"l"(a), "h"(static_cast<uint16_t>(v)));
    else if constexpr (sizeof(V) == 4) asm volatile("st.global.u32 [%0], %1;" ::"l"(a), "r"(static_cast<uint32_t>(v)));
    else asm volatile("st.global.u64 [%0], %1;" ::"l"(a), "l"(static_cast<uint64_t>(v)));
}

// what the compaction stores for a selected position: the staged selector item itself when it IS the content (SAME:
// one shared load per position), else the content item beside it
template <typename V, typename E, bool SAME>
__device__ __forceinline__ V selected_item(E m, const V *c) {
    if constexpr (SAME) {
        V v;
        memcpy(&v, &m, sizeof(V));
        return v;
    } else {
        return *c;
    }
}

// behind the selector's run of L items: whole rows (called by ONE warp; these bytes are written by no copy)
template <class SEL>
__device__ __forceinline__ void rows_pad(typename SEL::Elem *mv, int L, int lane) {
    if (L + lane < ((L + 31) & ~31)) mv[L + lane] = SEL::never();
}

// pass 1: how many positions of rows [r_begin, r_end) are selected (the whole warp gets the total: one REDUX)
template <class SEL>
__device__ __forceinline__ int rows_count(const SEL &sel, const typename SEL::Elem *mv, int r_begin, int r_end, int lane) {
    using E = typename SEL::Elem;
    const E *mp = mv + (r_begin << 5) + lane;
    int cnt = 0;
    int r = r_begin;
#pragma unroll 1
    for (; r + 2 <= r_end; r += 2) {
        const E m0 = mp[0], m1 = mp[32];
        cnt -= sel.minus_one_if(m0);
        cnt -= sel.minus_one_if(m1);
        mp += 64;
    }
    if (r < r_end) cnt -= sel.minus_one_if(mp[0]);
    return __reduce_add_sync(0xffffffffu, cnt);
}

// pass 2: rows [r_begin, r_end) compacted to out[run ...] (consecutive lanes -> consecutive slots), the rank of every
// position left in s_pfx; returns the rank behind the last row
template <typename V, class SEL, bool SAME>
__device__ __forceinline__ uint32_t rows_compact(const SEL &sel, const typename SEL::Elem *mv, const V *cv, uint16_t *s_pfx,
                                                 int r_begin, int r_end, int lane, uint32_t run, V *out) {
    using E = typename SEL::Elem;
    const uint32_t lt = (1u << lane) - 1u;
    uint64_t dst = static_cast<uint64_t>(__cvta_generic_to_global(out));
    asm volatile("" : "+l"(dst));
    const E *mp = mv + (r_begin << 5) + lane;
    const V *cp = cv + (r_begin << 5) + lane;
    uint16_t *pf = s_pfx + (r_begin << 5) + lane;
    int r = r_begin;
#pragma unroll 1
    for (; r + 2 <= r_end; r += 2) {
        const E m0 = mp[0], m1 = mp[32];
        const bool on0 = sel.test(m0), on1 = sel.test(m1);
        const unsigned b0 = __ballot_sync(0xffffffffu, on0), b1 = __ballot_sync(0xffffffffu, on1);
        const uint32_t slot0 = run + __popc(b0 & lt);
        run += __popc(b0);
        const uint32_t slot1 = run + __popc(b1 & lt);
        run += __popc(b1);
        pf[0] = static_cast<uint16_t>(slot0);
        pf[32] = static_cast<uint16_t>(slot1);
        if (on0) store_slot<V>(dst, slot0, selected_item<V, E, SAME>(m0, cp));
        if (on1) store_slot<V>(dst, slot1, selected_item<V, E, SAME>(m1, cp + 32));
        mp += 64;
        cp += 64;
        pf += 64;
    }
    if (r < r_end) {
        const E m0 = mp[0];
        const bool on0 = sel.test(m0);
        const unsigned b0 = __ballot_sync(0xffffffffu, on0);
        const uint32_t slot0 = run + __popc(b0 & lt);
        run += __popc(b0);
        pf[0] = static_cast<uint16_t>(slot0);
        if (on0) store_slot<V>(dst, slot0, selected_item<V, E, SAME>(m0, cp));
    }
    return run;
}

}  // namespace jg

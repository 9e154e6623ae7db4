// jg_compact.cuh -- the on-chip stream compaction shared by the event-tiled (jg_select.cu) and element-tiled
// (jg_elem.cu) selections: jagged.py:562-589 `content[mask]` for the staged run of one tile.
//
// The selection kernels are instruction-issue bound, not DRAM bound (ncu, profiles/r1_sweep_source_blocks.txt: the
// 32-positions-per-step ballot loop and its remainder cascades were 44 % of the warp instructions).  Here a warp step
// covers a ROW of 32 x EPL positions -- every lane owns EPL = 16 / sizeof(V) CONSECUTIVE positions that it fetches
// with ONE 16-byte shared-memory load (conflict free) -- so the per-step overhead (loop, prefix across lanes, running
// base) is paid once per 128 float32 positions instead of once per 32:
//   * the lane tests its EPL values -> a small bit mask, popcount = its number of selected positions;
//   * one shuffle scan over the lane counts gives every lane its first output slot; its selected values go to
//     consecutive slots (neighbouring lanes write neighbouring addresses: the same sectors as before);
//   * the per-position output rank (what the per-event epilogue reads back: new offsets and counts) is stored EPL
//     uint16 at a time.
// Positions are indexed from the 16-byte boundary below the first element (`lead` dummy positions in front, always
// unselected), so that the 16-byte loads are aligned whatever the alignment of the tile's run in global memory.
#pragma once

#include "jg_common.cuh"

namespace jg {

template <typename V>
struct RowCompact {
    static constexpr int EPL = 16 / static_cast<int>(sizeof(V));      // positions per lane and step
    static constexpr int ROW = kWarp * EPL;                           // positions per warp step
    static constexpr int kSlack = ROW;                                // a row may run this far past the staged run
};

// the lane's EPL values of row position q0 (16-byte aligned) and the bit mask of the selected ones
template <typename V, class SEL, bool SAME, bool LOADV>
__device__ __forceinline__ unsigned row_test(const SEL &sel, const V *__restrict__ cvb,
                                             const typename SEL::Elem *__restrict__ mv, int q0, int lead, int L,
                                             V (&val)[RowCompact<V>::EPL]) {
    using E = typename SEL::Elem;
    constexpr int EPL = RowCompact<V>::EPL;
    union {
        uint4 raw;
        V v[EPL];
    } w;
    if constexpr (SAME || LOADV) w.raw = *reinterpret_cast<const uint4 *>(cvb + q0);
    else w.raw = make_uint4(0, 0, 0, 0);
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        val[k] = w.v[k];
        const int p = q0 + k - lead;
        const bool valid = static_cast<unsigned>(p) < static_cast<unsigned>(L);
        bool on;
        if constexpr (SAME) {
            E e;
            memcpy(&e, &w.v[k], sizeof(E));
            on = valid && sel.test(e);
        } else {
            on = valid && sel.test(mv[valid ? p : 0]);
        }
        m |= on ? (1u << k) : 0u;
    }
    return m;
}

// ONE WARP: selected positions in its rows [row_begin, row_end)  (already summed over the lanes)
template <typename V, class SEL, bool SAME>
__device__ __forceinline__ int rows_count(const SEL &sel, const V *__restrict__ cvb,
                                          const typename SEL::Elem *__restrict__ mv, int lead, int L, int row_begin,
                                          int row_end) {
    constexpr int EPL = RowCompact<V>::EPL, ROW = RowCompact<V>::ROW;
    const int lane = threadIdx.x & 31;
    int cnt = 0;
#pragma unroll 1
    for (int r = row_begin; r < row_end; ++r) {
        V val[EPL];
        cnt += __popc(row_test<V, SEL, SAME, false>(sel, cvb, mv, r * ROW + lane * EPL, lead, L, val));
    }
    return warp_reduce_add(cnt);
}

// ONE WARP: compact its rows.  `run` = selected positions of the tile before the warp's first row; dst = out + tile base.
// s_pfx[q] receives the number of selected positions before q (q counted from the aligned boundary).
template <typename V, class SEL, bool SAME>
__device__ __forceinline__ void rows_compact(const SEL &sel, const V *__restrict__ cvb,
                                             const typename SEL::Elem *__restrict__ mv, int lead, int L, int row_begin,
                                             int row_end, int run, V *__restrict__ dst, uint16_t *__restrict__ s_pfx) {
    constexpr int EPL = RowCompact<V>::EPL, ROW = RowCompact<V>::ROW;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int r = row_begin; r < row_end; ++r) {
        const int q0 = r * ROW + lane * EPL;
        V val[EPL];
        const unsigned m = row_test<V, SEL, SAME, true>(sel, cvb, mv, q0, lead, L, val);
        const int c = __popc(m);
        const int incl = warp_incl_scan_add(c, lane);
        const int slot0 = run + incl - c;
        uint32_t rank[EPL];
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            rank[k] = static_cast<uint32_t>(slot0 + __popc(m & ((1u << k) - 1u)));
            if (m & (1u << k)) dst[rank[k]] = val[k];
        }
        if constexpr (EPL == 4) {
            *reinterpret_cast<uint2 *>(s_pfx + q0) = make_uint2(rank[0] | (rank[1] << 16), rank[2] | (rank[3] << 16));
        } else {
            *reinterpret_cast<uint32_t *>(s_pfx + q0) = rank[0] | (rank[1] << 16);
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

}  // namespace jg

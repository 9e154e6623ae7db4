// jg_offsets.cu -- offsets / counts / parents / localindex machinery (jagged.py:42-71, 352-443, 469-502).
#include "jg_scan.cuh"
#include "jg_tile.cuh"

namespace jg {

// ------------------------------------------------------------------------------------------------------------
// counts = offsets[1:] - offsets[:-1]           (jagged.py:383-388)   traffic 8(n+1) + 8n
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
offsets2counts_kernel(const int64_t *__restrict__ offsets, int64_t n, int64_t *__restrict__ counts, bool aligned) {
    // each thread produces two consecutive counts from three consecutive offsets
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 2;
    for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
        if (i + 1 < n) {
            int64_t a, b, c;
            if (aligned) {
                longlong2 v = __ldg(reinterpret_cast<const longlong2 *>(offsets + i));
                a = v.x;
                b = v.y;
            } else {
                a = offsets[i];
                b = offsets[i + 1];
            }
            c = __ldg(offsets + i + 2);
            if (aligned) {
                __stcs(reinterpret_cast<longlong2 *>(counts + i), make_longlong2(b - a, c - b));
            } else {
                counts[i] = b - a;
                counts[i + 1] = c - b;
            }
        } else {
            counts[i] = offsets[i + 1] - offsets[i];
        }
    }
}

__global__ void __launch_bounds__(256)
startsstops2counts_kernel(const int64_t *__restrict__ starts, const int64_t *__restrict__ stops, int64_t n,
                          int64_t *__restrict__ counts) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        counts[i] = stops[i] - starts[i];
}

// ------------------------------------------------------------------------------------------------------------
// _valid()   (jagged.py:120-127, 469-477)     traffic 8(n+1)
// info = {offsets[0], offsets[n], max count, #non-empty}
// ------------------------------------------------------------------------------------------------------------
// (Round 2 measured two vectorised forms -- 16-byte loads with the neighbour by shuffle, four or two pairs in flight:
// 0.240 and 0.249 ms on 125 M offsets against 0.228 ms for this plain grid-stride loop, which the compiler already
// unrolls with four loads in flight at full occupancy.)
__global__ void __launch_bounds__(256)
validate_offsets_kernel(const int64_t *__restrict__ offsets, int64_t n, int64_t content_len,
                        int64_t *__restrict__ info, int32_t *__restrict__ status) {
    int32_t st = 0;
    int64_t maxcount = 0, nonempty = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i <= n; i += stride) {
        const int64_t a = offsets[i];
        if (a < 0) st |= JG_ST_NEGATIVE_OFFSET;
        if (a > content_len) st |= JG_ST_BEYOND_CONTENT;
        if (i < n) {
            const int64_t c = offsets[i + 1] - a;
            if (c < 0) st |= JG_ST_NONMONOTONE;
            maxcount = max(maxcount, c);
            nonempty += (c > 0);
        }
    }
    // block reduce
    __shared__ int64_t s_max[8], s_cnt[8];
    __shared__ int32_t s_st[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        maxcount = max(maxcount, __shfl_xor_sync(0xffffffffu, maxcount, d));
        nonempty += __shfl_xor_sync(0xffffffffu, nonempty, d);
        st |= __shfl_xor_sync(0xffffffffu, st, d);
    }
    if (lane == 0) {
        s_max[warp] = maxcount;
        s_cnt[warp] = nonempty;
        s_st[warp] = st;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            maxcount = max(maxcount, s_max[w]);
            nonempty += s_cnt[w];
            st |= s_st[w];
        }
        if (st) atomicOr(status, st);
        atomicMax(reinterpret_cast<long long *>(info + 2), static_cast<long long>(maxcount));
        atomicAdd(reinterpret_cast<unsigned long long *>(info + 3), static_cast<unsigned long long>(nonempty));
        if (blockIdx.x == 0) {
            info[0] = offsets[0];
            info[1] = offsets[n];
        }
    }
}

// info = {min start over non-empty (or 0), max stop over non-empty (or 0), max count, #non-empty}
__global__ void __launch_bounds__(256)
validate_startsstops_kernel(const int64_t *__restrict__ starts, const int64_t *__restrict__ stops, int64_t n,
                            int64_t content_len, int64_t *__restrict__ info, int32_t *__restrict__ status) {
    int32_t st = 0;
    int64_t maxcount = 0, nonempty = 0, minstart = INT64_MAX, maxstop = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t a = starts[i], b = stops[i];
        if (a < 0 || b < 0) st |= JG_ST_NEGATIVE_OFFSET;
        if (b < a) st |= JG_ST_STOPS_LT_STARTS;
        if (b != a) {
            if (a >= content_len || b > content_len) st |= JG_ST_BEYOND_CONTENT;
            minstart = min(minstart, a);
            maxstop = max(maxstop, b);
            nonempty += 1;
            maxcount = max(maxcount, b - a);
        }
        if (i + 1 < n && starts[i + 1] != b) st |= JG_ST_NOT_OFFSETS;
    }
    if (st) atomicOr(status, st);
    if (nonempty) {
        atomicMin(reinterpret_cast<long long *>(info + 0), static_cast<long long>(minstart));
        atomicMax(reinterpret_cast<long long *>(info + 1), static_cast<long long>(maxstop));
        atomicMax(reinterpret_cast<long long *>(info + 2), static_cast<long long>(maxcount));
        atomicAdd(reinterpret_cast<unsigned long long *>(info + 3), static_cast<unsigned long long>(nonempty));
    }
}

__global__ void fill_i64_kernel(int64_t *p, int64_t n, int64_t v) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// ------------------------------------------------------------------------------------------------------------
// per-element expansion kernels built on the event tile (jg_tile.cuh):
//   parents[j] = i, localindex[j - o0] = j - offsets[i], broadcast out[j - o0] = flat[i]
// ------------------------------------------------------------------------------------------------------------
enum { EXPAND_PARENTS = 0, EXPAND_LOCALINDEX = 1, EXPAND_BROADCAST = 2 };

constexpr int kOwnPositions = 4096;      // positions of a tile the on-chip owner table covers

// (the expansions afford a table of 8192 positions -- 16 KB, still 8 CTAs per SM: tiles of mean counts up to ~14 fit it,
// so the host keeps such arrays on this kernel instead of the element-tiled one: parents on Poisson(10) 79 -> 9x %)
constexpr int kExpandPositions = 8192;

template <int MODE, typename V>
__global__ void __launch_bounds__(kTileThreads)
expand_kernel(const int64_t *__restrict__ offsets, int64_t n, V *__restrict__ out, const V *__restrict__ flat) {
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) uint16_t s_own[kExpandPositions + 128];
    __shared__ V s_flat[MODE == EXPAND_BROADCAST ? kTileEvents : 1];
    EventTile tile;
    // the per-event values do not depend on the offsets: their loads are in flight while the tile's offsets arrive
    V f0 = V(), f1 = V();
    if (MODE == EXPAND_BROADCAST) {
        int64_t ev0;
        int nev;
        EventTile::range(n, blockIdx.x, ev0, nev);
        if (static_cast<int>(threadIdx.x) < nev) f0 = __ldg(flat + ev0 + threadIdx.x);
        if (static_cast<int>(threadIdx.x) + kTileThreads < nev) f1 = __ldg(flat + ev0 + threadIdx.x + kTileThreads);
    }
    const int64_t o0 = __ldg(offsets);
    tile.load(s_off, offsets, n, blockIdx.x);
    if (MODE == EXPAND_PARENTS && blockIdx.x == 0) {
        // content below offsets[0] is unreachable: -1 (jagged.py:52-54)
        for (int64_t j = threadIdx.x; j < o0; j += kTileThreads) out[j] = static_cast<V>(-1);
    }
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    if (e1 - e0 <= kExpandPositions) {
        // short events: scatter + max-scan owner table in shared memory, then a coalesced element-parallel sweep
        const int L = static_cast<int>(e1 - e0);
        if (MODE == EXPAND_BROADCAST) {
            if (static_cast<int>(threadIdx.x) < tile.nev) s_flat[threadIdx.x] = f0;
            if (static_cast<int>(threadIdx.x) + kTileThreads < tile.nev) s_flat[threadIdx.x + kTileThreads] = f1;
        }
        build_owner_table(s_own, tile, L);
        V *dst = (MODE == EXPAND_PARENTS) ? out + e0 : out + (e0 - o0);
        if constexpr (MODE == EXPAND_BROADCAST) {
            // 16 bytes per thread and store: the sweep is issue-bound, not DRAM-bound, when every 1..4-byte item
            // costs its own store and loop step (tojagged(float32): 61 % of peak).  Head and tail of the run, up to
            // the first / from the last 16-byte boundary of the OUTPUT, go item by item.
            constexpr int VEC = 16 / static_cast<int>(sizeof(V));
            const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(dst) / sizeof(V)) & (VEC - 1));
            const int head = min(L, (VEC - mis) & (VEC - 1));
            const int ngroups = (L - head) / VEC;
            const int tail0 = head + ngroups * VEC;
            if (static_cast<int>(threadIdx.x) < head) dst[threadIdx.x] = s_flat[s_own[threadIdx.x]];
            if (tail0 + static_cast<int>(threadIdx.x) < L) dst[tail0 + threadIdx.x] = s_flat[s_own[tail0 + threadIdx.x]];
#pragma unroll 1
            for (int g = threadIdx.x; g < ngroups; g += kTileThreads) {
                const int p = head + g * VEC;
                union {
                    uint4 q;
                    V e[VEC];
                } w;
#pragma unroll
                for (int u = 0; u < VEC; ++u) w.e[u] = s_flat[s_own[p + u]];
                __stcs(reinterpret_cast<uint4 *>(dst + p), w.q);
            }
            return;
        }
#pragma unroll 2
        for (int p = threadIdx.x; p < L; p += kTileThreads) {
            const int ev = s_own[p];
            if (MODE == EXPAND_PARENTS) __stcs(dst + p, static_cast<V>(tile.event0 + ev));
            else if (MODE == EXPAND_LOCALINDEX) __stcs(dst + p, static_cast<V>(e0 + p - s_off[ev]));
            else __stcs(dst + p, s_flat[ev]);
        }
        return;
    }
    // long events: the run in chunks of kExpandPositions positions, one owner table each (scatter + running maximum, the
    // owner of a chunk's first position by one binary search per warp).  A binary search per ELEMENT, as in round 1,
    // is ~60 instructions per element: 1000 tiles with an event of 5 k - 70 k elements among 150 M short events made
    // up 40 % of the kernel's instructions (parents: 1.35 against 1.15 ms without them, profiles/r2_exp32.log).
    if (MODE == EXPAND_BROADCAST) {
        if (static_cast<int>(threadIdx.x) < tile.nev) s_flat[threadIdx.x] = f0;
        if (static_cast<int>(threadIdx.x) + kTileThreads < tile.nev) s_flat[threadIdx.x + kTileThreads] = f1;
    }
#pragma unroll 1
    for (int64_t cs = 0; cs < e1 - e0; cs += kExpandPositions) {
        const int L = static_cast<int>(min(static_cast<int64_t>(kExpandPositions), e1 - e0 - cs));
        build_owner_range(s_own, tile, cs, L);          // (ends with a barrier; s_flat is published by its first one)
        V *dst = (MODE == EXPAND_PARENTS) ? out + e0 + cs : out + (e0 + cs - o0);
#pragma unroll 2
        for (int p = threadIdx.x; p < L; p += kTileThreads) {
            const int ev = s_own[p];
            if (MODE == EXPAND_PARENTS) __stcs(dst + p, static_cast<V>(tile.event0 + ev));
            else if (MODE == EXPAND_LOCALINDEX) __stcs(dst + p, static_cast<V>(e0 + cs + p - s_off[ev]));
            else __stcs(dst + p, s_flat[ev]);
        }
        __syncthreads();                                 // the table is rewritten by the next chunk
    }
}

// general layout (jagged.py:56-71): out pre-filled with -1
__global__ void __launch_bounds__(256)
startsstops2parents_kernel(const int64_t *__restrict__ starts, const int64_t *__restrict__ stops, int64_t n,
                           int64_t *__restrict__ out, int64_t out_len) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t b = min(stops[i], out_len);
        // events may overlap (a[[2, 2]]): numpy's `out[content_indexes] = parents` (jagged.py:69) leaves the LAST
        // event's index, i.e. the largest one -- an atomic max instead of a racing plain store
        for (int64_t j = starts[i]; j < b; ++j) atomicMax(reinterpret_cast<long long *>(out) + j, static_cast<long long>(i));
    }
}

// Per-event block moves between a dense layout (offsets) and a general one (one start per event):
//   GATHER : compact()              out[offsets[i] + k] = content[other[i] + k]   (jagged.py:883-942)
//   SCATTER: concatenate(axis=1)    out[other[i] + k]   = content[offsets[i] + k] (jagged.py:1651-1733: the reference
//            builds one boolean mask per input from +1/-1 marks and a cumsum, then assigns through it)
// Element-parallel over the tile's dense run; the owner table turns a position into its event and a per-event
// shift (other[i] - offsets[i]) turns the dense address into the general one; 4 loads in flight per thread.
template <typename V, bool SCATTER>
__global__ void __launch_bounds__(kTileThreads)
segment_move_kernel(const V *__restrict__ content, const int64_t *__restrict__ offsets,
                    const int64_t *__restrict__ other, int64_t n, V *__restrict__ out) {
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ int64_t s_shift[kTileEvents];
    __shared__ __align__(16) uint16_t s_own[kOwnPositions + 128];
    EventTile tile;
    tile.load(s_off, offsets, n, blockIdx.x);
    {
        const int i0 = threadIdx.x, i1 = threadIdx.x + kTileThreads;
        if (i0 < tile.nev) s_shift[i0] = __ldg(other + tile.event0 + i0) - s_off[i0];
        if (i1 < tile.nev) s_shift[i1] = __ldg(other + tile.event0 + i1) - s_off[i1];
    }
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const bool table = (e1 - e0) <= kOwnPositions;
    if (table) build_owner_table(s_own, tile, static_cast<int>(e1 - e0));
    else __syncthreads();
#pragma unroll 1
    for (int64_t jb = e0 + threadIdx.x; jb < e1; jb += 4 * kTileThreads) {
        int64_t src[4], dst[4];
        V val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = jb + u * kTileThreads;
            src[u] = -1;
            if (j < e1) {
                const int ev = table ? static_cast<int>(s_own[j - e0]) : tile.find_event(j);
                const int64_t g = j + s_shift[ev];
                src[u] = SCATTER ? j : g;
                dst[u] = SCATTER ? g : j;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (src[u] >= 0) val[u] = __ldg(content + src[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (src[u] >= 0) out[dst[u]] = val[u];
    }
}

// pad(): out[dst[i] + k] = content[starts[i] + k] for k < counts[i], else 0 with invalid[...] = 1  (jagged.py:1851-1900:
// the reference builds an index from parents / localindex of the padded layout, gathers, and wraps the result in a
// MaskedArray whose mask is localindex >= counts).  Element-parallel over the padded (dense) layout.
template <typename V>
__global__ void __launch_bounds__(kTileThreads)
pad_kernel(const V *__restrict__ content, const int64_t *__restrict__ starts, const int64_t *__restrict__ counts,
           const int64_t *__restrict__ dst_offsets, int64_t n, V *__restrict__ out, uint8_t *__restrict__ invalid) {
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ int64_t s_shift[kTileEvents];
    __shared__ int64_t s_end[kTileEvents];       // first padded position of the event in the dense layout
    __shared__ __align__(16) uint16_t s_own[kOwnPositions + 128];
    EventTile tile;
    tile.load(s_off, dst_offsets, n, blockIdx.x);
#pragma unroll 1
    for (int ev = threadIdx.x; ev < tile.nev; ev += kTileThreads) {
        s_shift[ev] = __ldg(starts + tile.event0 + ev) - s_off[ev];
        s_end[ev] = s_off[ev] + __ldg(counts + tile.event0 + ev);
    }
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const bool table = (e1 - e0) <= kOwnPositions;
    if (table) build_owner_table(s_own, tile, static_cast<int>(e1 - e0));
    else __syncthreads();
#pragma unroll 1
    for (int64_t jb = e0 + threadIdx.x; jb < e1; jb += 4 * kTileThreads) {
        int64_t src[4];
        V val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = jb + u * kTileThreads;
            src[u] = -1;
            if (j < e1) {
                const int ev = table ? static_cast<int>(s_own[j - e0]) : tile.find_event(j);
                if (j < s_end[ev]) src[u] = j + s_shift[ev];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) val[u] = src[u] >= 0 ? __ldg(content + src[u]) : V(0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = jb + u * kTileThreads;
            if (j < e1) {
                out[j] = val[u];
                if (invalid) invalid[j] = src[u] < 0;
            }
        }
    }
}

template <bool SCATTER>
static int launch_segment_move(const void *content, int itemsize, const int64_t *offsets, const int64_t *other,
                               int64_t n, void *out, cudaStream_t s, const char *what) {
    const unsigned g = tile_grid(n);
    switch (itemsize) {
        case 1:
            segment_move_kernel<uint8_t, SCATTER><<<g, kTileThreads, 0, s>>>(static_cast<const uint8_t *>(content), offsets,
                                                                            other, n, static_cast<uint8_t *>(out));
            break;
        case 2:
            segment_move_kernel<uint16_t, SCATTER><<<g, kTileThreads, 0, s>>>(static_cast<const uint16_t *>(content), offsets,
                                                                             other, n, static_cast<uint16_t *>(out));
            break;
        case 4:
            segment_move_kernel<uint32_t, SCATTER><<<g, kTileThreads, 0, s>>>(static_cast<const uint32_t *>(content), offsets,
                                                                             other, n, static_cast<uint32_t *>(out));
            break;
        case 8:
            segment_move_kernel<uint64_t, SCATTER><<<g, kTileThreads, 0, s>>>(static_cast<const uint64_t *>(content), offsets,
                                                                             other, n, static_cast<uint64_t *>(out));
            break;
        default:
            set_error("%s: unsupported itemsize %d", what, itemsize);
            return -2;
    }
    return check_launch(what);
}

// concatenate(axis=1) of K dense inputs in ONE sweep over the OUTPUT layout (jagged.py:1651-1733).  A tile owns 512
// merged events; per event and input it keeps where the input's part ends in the output and the shift that turns an
// output position into the input's content position.  The sweep is element-parallel over the tile's output run:
// stores are fully coalesced, and so are the loads -- the parts an input contributes to consecutive events are
// consecutive in that input's content.  (One jg_segment_scatter per input instead writes runs of a few items
// between the gaps the other inputs fill later: partial sectors on every store.)
constexpr int kMergeMax = 4;
struct MergeInputs {
    const void *content[kMergeMax];
    const int64_t *offsets[kMergeMax];
};

template <typename V, int K>
__global__ void __launch_bounds__(kTileThreads)
segment_merge_kernel(const MergeInputs in, const int64_t *__restrict__ out_offsets, int64_t n, V *__restrict__ out) {
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ int64_t s_end[K - 1][kTileEvents];      // output position where the part of input q ends
    __shared__ int64_t s_shift[K][kTileEvents];        // content position in input q = output position + shift
    __shared__ __align__(16) uint16_t s_own[kOwnPositions + 128];
    EventTile tile;
    // the inputs' offsets of this thread's two events are requested before the tile's output offsets are waited for
    int64_t ia[2][K], ib[2][K];
    {
        int64_t ev0;
        int nev;
        EventTile::range(n, blockIdx.x, ev0, nev);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ev = threadIdx.x + h * kTileThreads;
#pragma unroll
            for (int q = 0; q < K; ++q) {
                ia[h][q] = ib[h][q] = 0;
                if (ev < nev) {
                    ia[h][q] = __ldg(in.offsets[q] + ev0 + ev);
                    ib[h][q] = __ldg(in.offsets[q] + ev0 + ev + 1);
                }
            }
        }
    }
    tile.load(s_off, out_offsets, n, blockIdx.x);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ev = threadIdx.x + h * kTileThreads;
        if (ev < tile.nev) {
            int64_t pos = s_off[ev];
#pragma unroll
            for (int q = 0; q < K; ++q) {
                s_shift[q][ev] = ia[h][q] - pos;
                pos += ib[h][q] - ia[h][q];
                if (q < K - 1) s_end[q][ev] = pos;
            }
        }
    }
    __syncthreads();
    // merged events are K times as long as their inputs': the tile is swept in K sub-ranges of events so that the
    // output run of each fits the owner table as a one-input tile does
    constexpr int kSub = (kTileEvents + K - 1) / K;
#pragma unroll 1
    for (int c0 = 0; c0 < tile.nev; c0 += kSub) {
        EventTile sub;
        sub.off = s_off + c0;
        sub.event0 = tile.event0 + c0;
        sub.nev = min(kSub, tile.nev - c0);
        const int64_t e0 = sub.first_element(), e1 = sub.last_element();
        const bool table = (e1 - e0) <= kOwnPositions;
        if (table) build_owner_table(s_own, sub, static_cast<int>(e1 - e0));
#pragma unroll 1
        for (int64_t jb = e0 + threadIdx.x; jb < e1; jb += 4 * kTileThreads) {
            const V *src[4];
            V val[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t j = jb + u * kTileThreads;
                src[u] = nullptr;
                if (j < e1) {
                    const int ev = c0 + (table ? static_cast<int>(s_own[j - e0]) : sub.find_event(j));
                    const V *c = static_cast<const V *>(in.content[K - 1]);
                    int64_t shift = s_shift[K - 1][ev];
#pragma unroll
                    for (int q = K - 2; q >= 0; --q) {
                        if (j < s_end[q][ev]) {
                            c = static_cast<const V *>(in.content[q]);
                            shift = s_shift[q][ev];
                        }
                    }
                    src[u] = c + (j + shift);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (src[u]) val[u] = __ldg(src[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (src[u]) out[jb + u * kTileThreads] = val[u];
        }
        __syncthreads();      // the owner table is rebuilt for the next sub-range
    }
}

template <typename V>
static int launch_segment_merge(const MergeInputs &in, int k, const int64_t *out_offsets, int64_t n, void *out,
                                cudaStream_t s) {
    const unsigned g = tile_grid(n);
    V *o = static_cast<V *>(out);
    if (k == 2) segment_merge_kernel<V, 2><<<g, kTileThreads, 0, s>>>(in, out_offsets, n, o);
    else if (k == 3) segment_merge_kernel<V, 3><<<g, kTileThreads, 0, s>>>(in, out_offsets, n, o);
    else segment_merge_kernel<V, 4><<<g, kTileThreads, 0, s>>>(in, out_offsets, n, o);
    return check_launch("jg_segment_merge");
}

// concatenate(axis=0): offsets of a later array shifted by the content that precedes it (jagged.py:1633-1649)
__global__ void __launch_bounds__(256)
offsets_rebase_kernel(const int64_t *__restrict__ offsets, int64_t count, int64_t delta, int64_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < count; i += gridDim.x * 256LL) out[i] = offsets[i] + delta;
}

static inline int grid_for(int64_t work, int threads, int per_thread = 1) {
    int64_t blocks = ceil_div(work, static_cast<int64_t>(threads) * per_thread);
    int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace jg

using namespace jg;

extern "C" {

int jg_counts2offsets(const void *counts, int counts_dtype, int64_t n, int64_t *offsets, int64_t *total,
                      int32_t *status, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && offsets != nullptr, "jg_counts2offsets: bad arguments");
    cudaStream_t s = as_stream(stream);
    switch (counts_dtype) {
        case JG_I64:
            return launch_scan_array(static_cast<const int64_t *>(counts), n, offsets, total, status, ws, ws_bytes, s);
        case JG_I32:
            return launch_scan_array(static_cast<const int32_t *>(counts), n, offsets, total, status, ws, ws_bytes, s);
        case JG_U8:
        case JG_B8: {
            ScanInArray<uint8_t> in{static_cast<const uint8_t *>(counts), false};
            return launch_scan(in, n, offsets, total, status, ws, ws_bytes, s);
        }
    }
    set_error("jg_counts2offsets: unsupported counts dtype %d", counts_dtype);
    return -2;
}

int jg_counts2offsets_tiles(const void *counts, int counts_dtype, int64_t n, int64_t *offsets, int64_t *tile_nonempty,
                            int64_t *total, int32_t *status, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && offsets != nullptr && tile_nonempty != nullptr, "jg_counts2offsets_tiles: bad arguments");
    static_assert(kTileEvents == 512, "the scan counts non-empty events per 512-event tile");
    cudaStream_t s = as_stream(stream);
    switch (counts_dtype) {
        case JG_I64:
            return launch_scan_array(static_cast<const int64_t *>(counts), n, offsets, total, status, ws, ws_bytes, s, tile_nonempty);
        case JG_I32:
            return launch_scan_array(static_cast<const int32_t *>(counts), n, offsets, total, status, ws, ws_bytes, s, tile_nonempty);
    }
    set_error("jg_counts2offsets_tiles: counts dtype %d (int32 / int64 only)", counts_dtype);
    return -2;
}

int jg_offsets2counts(const int64_t *offsets, int64_t n, int64_t *counts, void *stream) {
    if (n <= 0) return 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(offsets) | reinterpret_cast<uintptr_t>(counts)) & 15) == 0;
    offsets2counts_kernel<<<grid_for(n, 256, 2), 256, 0, as_stream(stream)>>>(offsets, n, counts, aligned);
    return check_launch("offsets2counts_kernel");
}

int jg_startsstops2counts(const int64_t *starts, const int64_t *stops, int64_t n, int64_t *counts, void *stream) {
    if (n <= 0) return 0;
    startsstops2counts_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(starts, stops, n, counts);
    return check_launch("startsstops2counts_kernel");
}

int jg_validate_offsets(const int64_t *offsets, int64_t n, int64_t content_len, int64_t *info, int32_t *status,
                        void *stream) {
    JG_REQUIRE(n >= 0 && info && status, "jg_validate_offsets: bad arguments");
    cudaStream_t s = as_stream(stream);
    cudaMemsetAsync(info, 0, 4 * sizeof(int64_t), s);
    validate_offsets_kernel<<<grid_for(n + 1, 256, 4), 256, 0, s>>>(offsets, n, content_len, info, status);
    return check_launch("validate_offsets_kernel");
}

int jg_validate_startsstops(const int64_t *starts, const int64_t *stops, int64_t n, int64_t content_len,
                            int64_t *info, int32_t *status, void *stream) {
    JG_REQUIRE(n >= 0 && info && status, "jg_validate_startsstops: bad arguments");
    cudaStream_t s = as_stream(stream);
    fill_i64_kernel<<<1, 32, 0, s>>>(info, 1, INT64_MAX);
    cudaMemsetAsync(info + 1, 0, 3 * sizeof(int64_t), s);
    if (n > 0)
        validate_startsstops_kernel<<<grid_for(n, 256, 2), 256, 0, s>>>(starts, stops, n, content_len, info, status);
    return check_launch("validate_startsstops_kernel");
}

int jg_offsets2parents(const int64_t *offsets, int64_t n, int64_t *parents, void *stream) {
    if (n <= 0) return 0;
    expand_kernel<EXPAND_PARENTS, int64_t>
        <<<tile_grid(n), kTileThreads, 0, as_stream(stream)>>>(offsets, n, parents, nullptr);
    return check_launch("expand_kernel<parents>");
}

int jg_localindex(const int64_t *offsets, int64_t n, int64_t *out, void *stream) {
    if (n <= 0) return 0;
    expand_kernel<EXPAND_LOCALINDEX, int64_t>
        <<<tile_grid(n), kTileThreads, 0, as_stream(stream)>>>(offsets, n, out, nullptr);
    return check_launch("expand_kernel<localindex>");
}

int jg_startsstops2parents(const int64_t *starts, const int64_t *stops, int64_t n, int64_t *out, int64_t out_len,
                           void *stream) {
    cudaStream_t s = as_stream(stream);
    if (out_len > 0) fill_i64_kernel<<<grid_for(out_len, 256), 256, 0, s>>>(out, out_len, -1);
    if (n > 0) startsstops2parents_kernel<<<grid_for(n, 256), 256, 0, s>>>(starts, stops, n, out, out_len);
    return check_launch("startsstops2parents_kernel");
}

int jg_compact_gather(const void *content, int itemsize, const int64_t *starts, const int64_t *dense_offsets,
                      int64_t n, void *out, void *stream) {
    if (n <= 0) return 0;
    JG_REQUIRE(starts && dense_offsets && out, "jg_compact_gather: bad arguments");
    return launch_segment_move<false>(content, itemsize, dense_offsets, starts, n, out, as_stream(stream),
                                      "jg_compact_gather");
}

int jg_segment_scatter(const void *content, int itemsize, const int64_t *offsets, const int64_t *dst_starts, int64_t n,
                       void *out, void *stream) {
    if (n <= 0) return 0;
    JG_REQUIRE(offsets && dst_starts && out, "jg_segment_scatter: bad arguments");
    return launch_segment_move<true>(content, itemsize, offsets, dst_starts, n, out, as_stream(stream),
                                     "jg_segment_scatter");
}

int jg_segment_merge(const void *const *contents, const int64_t *const *offsets, int k, int itemsize,
                     const int64_t *out_offsets, int64_t n, void *out, void *stream) {
    JG_REQUIRE(k >= 2 && k <= kMergeMax, "jg_segment_merge: between 2 and 4 inputs per call");
    if (n <= 0) return 0;
    JG_REQUIRE(contents && offsets && out_offsets && out, "jg_segment_merge: bad arguments");
    MergeInputs in;
    for (int q = 0; q < kMergeMax; ++q) {
        in.content[q] = q < k ? contents[q] : nullptr;
        in.offsets[q] = q < k ? offsets[q] : nullptr;
        JG_REQUIRE(q >= k || in.offsets[q] != nullptr, "jg_segment_merge: missing offsets");
    }
    cudaStream_t s = as_stream(stream);
    switch (itemsize) {
        case 1: return launch_segment_merge<uint8_t>(in, k, out_offsets, n, out, s);
        case 2: return launch_segment_merge<uint16_t>(in, k, out_offsets, n, out, s);
        case 4: return launch_segment_merge<uint32_t>(in, k, out_offsets, n, out, s);
        case 8: return launch_segment_merge<uint64_t>(in, k, out_offsets, n, out, s);
        default:
            set_error("jg_segment_merge: unsupported itemsize %d", itemsize);
            return -2;
    }
}

int jg_pad(const void *content, int itemsize, const int64_t *starts, const int64_t *counts, const int64_t *dst_offsets,
           int64_t n, void *out, uint8_t *invalid, void *stream) {
    if (n <= 0) return 0;
    JG_REQUIRE(starts && counts && dst_offsets && out, "jg_pad: bad arguments");
    cudaStream_t s = as_stream(stream);
    const unsigned g = tile_grid(n);
    switch (itemsize) {
        case 1:
            pad_kernel<uint8_t><<<g, kTileThreads, 0, s>>>(static_cast<const uint8_t *>(content), starts, counts,
                                                          dst_offsets, n, static_cast<uint8_t *>(out), invalid);
            break;
        case 2:
            pad_kernel<uint16_t><<<g, kTileThreads, 0, s>>>(static_cast<const uint16_t *>(content), starts, counts,
                                                           dst_offsets, n, static_cast<uint16_t *>(out), invalid);
            break;
        case 4:
            pad_kernel<uint32_t><<<g, kTileThreads, 0, s>>>(static_cast<const uint32_t *>(content), starts, counts,
                                                           dst_offsets, n, static_cast<uint32_t *>(out), invalid);
            break;
        case 8:
            pad_kernel<uint64_t><<<g, kTileThreads, 0, s>>>(static_cast<const uint64_t *>(content), starts, counts,
                                                           dst_offsets, n, static_cast<uint64_t *>(out), invalid);
            break;
        default:
            set_error("jg_pad: unsupported itemsize %d", itemsize);
            return -2;
    }
    return check_launch("pad_kernel");
}

int jg_offsets_rebase(const int64_t *offsets, int64_t count, int64_t delta, int64_t *out, void *stream) {
    if (count <= 0) return 0;
    JG_REQUIRE(offsets && out, "jg_offsets_rebase: bad arguments");
    offsets_rebase_kernel<<<grid_for(count, 256, 4), 256, 0, as_stream(stream)>>>(offsets, count, delta, out);
    return check_launch("offsets_rebase_kernel");
}

int jg_broadcast(const void *flat, int itemsize, const int64_t *offsets, int64_t n, void *out, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const unsigned g = tile_grid(n);
    switch (itemsize) {
        case 1:
            expand_kernel<EXPAND_BROADCAST, uint8_t>
                <<<g, kTileThreads, 0, s>>>(offsets, n, static_cast<uint8_t *>(out), static_cast<const uint8_t *>(flat));
            break;
        case 2:
            expand_kernel<EXPAND_BROADCAST, uint16_t><<<g, kTileThreads, 0, s>>>(
                offsets, n, static_cast<uint16_t *>(out), static_cast<const uint16_t *>(flat));
            break;
        case 4:
            expand_kernel<EXPAND_BROADCAST, uint32_t><<<g, kTileThreads, 0, s>>>(
                offsets, n, static_cast<uint32_t *>(out), static_cast<const uint32_t *>(flat));
            break;
        case 8:
            expand_kernel<EXPAND_BROADCAST, uint64_t><<<g, kTileThreads, 0, s>>>(
                offsets, n, static_cast<uint64_t *>(out), static_cast<const uint64_t *>(flat));
            break;
        default:
            set_error("jg_broadcast: unsupported itemsize %d", itemsize);
            return -2;
    }
    return check_launch("expand_kernel<broadcast>");
}

}  // extern "C"

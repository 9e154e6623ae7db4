// jg_comb.cu -- per-event combinatorics: pairs (i<=j), distincts (i<j), choose (k = 2..5, colexicographic) and
// cross (rectangular) index generation, jagged.py:1070-1126, 1128-1221, 1302-1325.
//
// Phase 1 (jg_comb_offsets): per-event output count computed on the fly inside the look-back scan -> new offsets
// and the grand total P.  Phase 2 (jg_comb_fill / jg_comb_gather2 / jg_comb_gather_multi): one CTA per event tile.
// Usual case: the pairs of the tile are first laid out EVENT-parallel in a shared-memory pair table (each thread
// walks its event's pairs in order with two counters), then the output run is swept element-parallel with
// coalesced stores, reading its pair from the table.  Tiles whose output run exceeds the table, or that hold an
// event with hundreds of pairs, take the general path: the output finds its event through the on-chip owner
// table (binary search for very long runs) and inverts the triangular / rectangular / combinatorial numbering
// with integer-exact arithmetic (a floating-point estimate corrected by integer comparisons; the reference's
// float64 closed forms are exact only up to n ~ 6000 for pairs, "1000 choose 3", "100 choose 4").
//
// Algorithmic traffic: 8(N+1) [x2 for cross] read, 8(N+1) + k*8*P written.
#include "jg_ops.cuh"
#include "jg_scan.cuh"
#include "jg_tile.cuh"

#include <cstddef>
#include <mutex>
#include <type_traits>

namespace jg {

// (i, j) of the k-th pair with i <= j in an event of n items; rows are i, row i holds n - i entries
__device__ __forceinline__ void unrank_pairs(int64_t k, int64_t n, int64_t &i, int64_t &j) {
    if (n <= 48) {
        // short events (the common case): walk the rows -- a handful of 32-bit subtractions beats the square root
        int m = static_cast<int>(k), row = static_cast<int>(n), r32 = 0;
        while (m >= row) {
            m -= row;
            --row;
            ++r32;
        }
        i = r32;
        j = r32 + m;
        return;
    }
    int64_t r;
    if (n < 1024) {       // b*b - 8k < 2^23: exact in float32, the estimate is within one row and corrected below
        const float b = 2.0f * static_cast<float>(n) + 1.0f;
        r = static_cast<int64_t>((b - sqrtf(b * b - 8.0f * static_cast<float>(k))) * 0.5f);
    } else {
        const double b = 2.0 * static_cast<double>(n) + 1.0;
        r = static_cast<int64_t>(floor((b - sqrt(b * b - 8.0 * static_cast<double>(k))) * 0.5));
    }
    if (r < 0) r = 0;
    if (r > n - 1) r = n - 1;
    // L(r) = r*n - r(r-1)/2 = first index of row r
    while (r + 1 < n && (r + 1) * n - ((r + 1) * r >> 1) <= k) ++r;
    while (r > 0 && r * n - (r * (r - 1) >> 1) > k) --r;
    i = r;
    j = r + (k - (r * n - (r * (r - 1) >> 1)));
}

// (i, j) of the k-th pair with i < j; row i holds n - 1 - i entries
__device__ __forceinline__ void unrank_distincts(int64_t k, int64_t n, int64_t &i, int64_t &j) {
    if (n <= 48) {
        int m = static_cast<int>(k), row = static_cast<int>(n) - 1, r32 = 0;
        while (m >= row) {
            m -= row;
            --row;
            ++r32;
        }
        i = r32;
        j = r32 + 1 + m;
        return;
    }
    int64_t r;
    if (n < 1024) {
        const float b = 2.0f * static_cast<float>(n) - 1.0f;
        r = static_cast<int64_t>((b - sqrtf(b * b - 8.0f * static_cast<float>(k))) * 0.5f);
    } else {
        const double b = 2.0 * static_cast<double>(n) - 1.0;
        r = static_cast<int64_t>(floor((b - sqrt(b * b - 8.0 * static_cast<double>(k))) * 0.5));
    }
    if (r < 0) r = 0;
    if (r > n - 2) r = n - 2;
    // L(r) = r*(n-1) - r(r-1)/2
    while (r + 1 < n - 1 && (r + 1) * (n - 1) - ((r + 1) * r >> 1) <= k) ++r;
    while (r > 0 && r * (n - 1) - (r * (r - 1) >> 1) > k) --r;
    i = r;
    j = r + 1 + (k - (r * (n - 1) - (r * (r - 1) >> 1)));
}

__device__ __forceinline__ int64_t binom(int64_t c, int k) {
    if (c < k) return 0;
    switch (k) {
        case 1: return c;
        case 2: return c * (c - 1) / 2;
        case 3: return c * (c - 1) * (c - 2) / 6;
        case 4: return c * (c - 1) * (c - 2) * (c - 3) / 24;
        default: return c * (c - 1) * (c - 2) * (c - 3) * (c - 4) / 120;
    }
}

// combinatorial number system (colex order): largest c with C(c, k) <= m
__device__ __forceinline__ int64_t largest_binom_le(int64_t m, int k, int64_t n) {
    double est;
    const double dm = static_cast<double>(m);
    switch (k) {
        case 2: est = 0.5 * (1.0 + sqrt(8.0 * dm + 1.0)); break;
        case 3: est = cbrt(6.0 * dm) + 1.0; break;
        case 4: est = sqrt(sqrt(24.0 * dm)) + 1.5; break;
        default: est = pow(120.0 * dm, 0.2) + 2.0; break;
    }
    int64_t c = static_cast<int64_t>(est);
    if (c < k - 1) c = k - 1;
    if (c > n - 1) c = n - 1;
    while (c + 1 <= n - 1 && binom(c + 1, k) <= m) ++c;
    while (c > k - 1 && binom(c, k) > m) --c;
    return c;
}

// C(c, q) for c <= 64 in 32-bit arithmetic (C(64, 5) = 7 624 512; the falling factorial 64*63*62*61*60 < 2^32).
// c < q gives 0: one of the factors is zero.
__device__ __forceinline__ uint32_t binom32(uint32_t c, int q) {
    switch (q) {
        case 1: return c;
        case 2: return c * (c - 1) / 2;
        case 3: return c * (c - 1) * (c - 2) / 6;
        case 4: return c * (c - 1) * (c - 2) * (c - 3) / 24;
        default: return c * (c - 1) * (c - 2) * (c - 3) * (c - 4) / 120;
    }
}

constexpr int kChooseSmall = 64;        // events up to this long un-rank k-combinations by a short downward walk

constexpr int kCombPositions = 4096;    // output positions one owner table of the direct path covers (longer runs: chunk by chunk)
constexpr int kOwnerPositions = 8192;   // ... and of the short-event kernels (two pair tables of 4096 entries alias its storage)
static_assert(2 * kCombPositions <= kOwnerPositions + 128, "pair tables must fit the owner table's storage");

struct CombIndex {
    int64_t v[5];
};

template <int KIND>
__device__ __forceinline__ CombIndex unrank(int64_t m, int64_t na, int64_t nb, int k) {
    CombIndex r;
    if constexpr (KIND == JG_COMB_PAIRS) {
        unrank_pairs(m, na, r.v[0], r.v[1]);
    } else if constexpr (KIND == JG_COMB_DISTINCTS) {
        unrank_distincts(m, na, r.v[0], r.v[1]);
    } else if constexpr (KIND == JG_COMB_CROSS) {
        if ((m | nb) < (1ll << 31)) {
            const uint32_t q = static_cast<uint32_t>(m) / static_cast<uint32_t>(nb);
            r.v[0] = q;
            r.v[1] = static_cast<uint32_t>(m) - q * static_cast<uint32_t>(nb);
        } else {
            r.v[0] = m / nb;
            r.v[1] = m - r.v[0] * nb;
        }
    } else {
        // columns are i1 < i2 < ... < ik; the largest index is found first
        if (na <= kChooseSmall) {
            // short events (every collider event): walk c down from the previous column's index -- at most n + k steps of
            // 32-bit integer arithmetic in total, where the closed-form estimates below cost a double-precision
            // sqrt / cbrt / pow per column (jagged.py:1131-1144 `root2/3/4`)
            uint32_t m32 = static_cast<uint32_t>(m), c = static_cast<uint32_t>(na);
            for (int q = k; q >= 2; --q) {
                --c;
                uint32_t bq = binom32(c, q);
                while (bq > m32) {
                    --c;
                    bq = binom32(c, q);
                }
                r.v[q - 1] = c;
                m32 -= bq;
            }
            r.v[0] = m32;
            return r;
        }
        for (int q = k; q >= 2; --q) {
            const int64_t c = largest_binom_le(m, q, na);
            r.v[q - 1] = c;
            m -= binom(c, q);
        }
        r.v[0] = m;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// Direct tables: every OUTPUT position of a tile finds its combination with a handful of instructions and no
// divergence -- the usual path of all four kernels below.
//   * owner table (jg_tile.cuh): output position -> event, built by scatter + running maximum over chunks of
//     kCombPositions positions (a tile of "choose 4" holds more outputs than one table: it is swept chunk by chunk);
//   * event info (16 bytes per event, built event-parallel): where the event's outputs and items start, relative to
//     the tile, and how to turn the rank r of an output inside its event into item indices:
//       pairs / distincts / choose(k): a look-up table indexed by (combinations of all shorter events) + r holds the
//         k indices in 6-bit fields -- events up to comb_lut_max_n items; the table is filled once per device by
//         running the exact un-ranking above on every (n, r), so both paths agree by construction.  Its head (events
//         up to comb_lut_stage_n items: every collider event) is copied to shared memory by each CTA while the
//         offsets tiles are in flight; the rest is read from global memory (with ~35 KB of shared memory per CTA the
//         L1 is too small to keep it: reading everything from there made the pair kernels 10 - 80 % slower);
//       cross: i = r / nb by one multiply-high with a per-event reciprocal, j = r - i * nb;
//   * longer events un-rank arithmetically (rare lanes); tiles whose runs do not fit the packed fields (2^20 outputs,
//     cross with 4096 or more right-hand items in one event) take the general path of each kernel.
// Round 2 (profiles/r2_comb_direct.md): the event-parallel pair table this replaces cost 2.9 warp instructions per
// pair at 18 of 32 lanes active (a warp walked the longest of its 32 events).
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kNoLut = 0xffffffffu;
constexpr int kCombLutShared = 1536;          // entries of the table a CTA keeps on chip
constexpr int kCombDirectMax = 1 << 20;      // outputs of a tile the packed event info can address
constexpr int kCrossMaxRight = 1 << 12;

__host__ __device__ constexpr int comb_lut_slot(int kind, int k) { return kind == JG_COMB_PAIRS ? 0 : (kind == JG_COMB_DISTINCTS ? 1 : k); }
__host__ __device__ constexpr int comb_lut_max_n(int kind, int k) {
    return kind != JG_COMB_CHOOSE ? 24 : (k == 2 ? 24 : (k == 3 ? 20 : (k == 4 ? 16 : 14)));
}
// events up to this long are un-ranked from the on-chip head of the table
__host__ __device__ constexpr int comb_lut_stage_n(int kind, int k) { return (kind == JG_COMB_CHOOSE && k == 5) ? 10 : 12; }
// number of combinations of all events shorter than n: the first table entry of the n-item events
__host__ __device__ inline uint32_t comb_lut_base(int kind, int k, uint32_t n) {
    if (kind == JG_COMB_PAIRS) return n == 0 ? 0u : (n - 1) * n * (n + 1) / 6;                 // sum m(m+1)/2
    if (kind == JG_COMB_DISTINCTS) return n < 3 ? 0u : n * (n - 1) * (n - 2) / 6;              // sum m(m-1)/2 = C(n, 3)
    // sum over m < n of C(m, k) = C(n, k + 1)
    if (n < static_cast<uint32_t>(k + 1)) return 0u;
    uint32_t c = 1;
    for (int q = 1; q <= k + 1; ++q) c = c * (n - (k + 1) + q) / q;      // exact at every step, < 2^32 for n <= 25
    return c;
}

template <int KIND>
__global__ void __launch_bounds__(256)
comb_lut_kernel(uint32_t *__restrict__ lut, int k) {
    const uint32_t n = blockIdx.x;
    const uint32_t base = comb_lut_base(KIND, k, n), cnt = comb_lut_base(KIND, k, n + 1) - base;
    const int ncols = KIND == JG_COMB_CHOOSE ? k : 2;
    for (uint32_t m = threadIdx.x; m < cnt; m += blockDim.x) {
        const CombIndex r = unrank<KIND>(m, n, 0, k);
        uint32_t code = 0;
        for (int c = 0; c < ncols; ++c) code |= static_cast<uint32_t>(r.v[c]) << (6 * c);
        lut[base + m] = code;
    }
}

// the table of (kind, k) on the current device, filled on first use
static const uint32_t *comb_lut(int kind, int k, cudaStream_t s) {
    if (kind == JG_COMB_CROSS) return nullptr;
    constexpr int kMaxDev = 64;
    static std::mutex mu;
    static uint32_t *tab[kMaxDev][6] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    const int slot = comb_lut_slot(kind, k);
    std::lock_guard<std::mutex> lock(mu);
    if (tab[dev][slot]) return tab[dev][slot];
    const int maxn = comb_lut_max_n(kind, k);
    const uint32_t entries = comb_lut_base(kind, k, static_cast<uint32_t>(maxn) + 1);
    uint32_t *t = nullptr;
    if (cudaMalloc(&t, static_cast<size_t>(entries) * 4) != cudaSuccess) return nullptr;
    switch (kind) {
        case JG_COMB_PAIRS: comb_lut_kernel<JG_COMB_PAIRS><<<maxn + 1, 256, 0, s>>>(t, k); break;
        case JG_COMB_DISTINCTS: comb_lut_kernel<JG_COMB_DISTINCTS><<<maxn + 1, 256, 0, s>>>(t, k); break;
        default: comb_lut_kernel<JG_COMB_CHOOSE><<<maxn + 1, 256, 0, s>>>(t, k); break;
    }
    // once per device and table: later calls may come on other streams
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        cudaFree(t);
        return nullptr;
    }
    tab[dev][slot] = t;
    return t;
}

template <int KIND>
__device__ __noinline__ CombIndex unrank_call(uint32_t m, uint32_t na, int k) {
    return unrank<KIND>(m, na, 0, k);
}

struct CombLut {
    const uint32_t *table;     // device memory
    uint32_t staged;           // leading entries every CTA copies to shared memory
};

static CombLut comb_lut_for(int kind, int k, cudaStream_t s) {
    CombLut l{comb_lut(kind, k, s), 0};
    if (l.table) l.staged = comb_lut_base(kind, k, static_cast<uint32_t>(comb_lut_stage_n(kind, k)) + 1);
    return l;
}

// all threads, before the offsets tiles are loaded: the copy's latency hides behind theirs (made visible by the
// barrier that follows those loads)
__device__ __forceinline__ void comb_stage_lut(uint32_t *__restrict__ s_lut, const CombLut &lut) {
    for (uint32_t i = threadIdx.x; i < lut.staged; i += kTileThreads) s_lut[i] = __ldg(lut.table + i);
}

// Shared memory of the direct path (one per kernel, next to the offsets tiles)
template <int KIND, bool VALUES = false>
struct CombShared {
    uint4 info[kTileEvents];                                    // one entry per event of the tile
    uint16_t own[kCombPositions + 128];                         // owner of every output position of the current chunk
    uint16_t right[VALUES ? kCombPositions : 1];                // value forms: own / right double as the pair table
    uint32_t lut[KIND == JG_COMB_CROSS ? 1 : kCombLutShared];   // head of the table
    uint16_t carry[kTileThreads / kWarp];                       // owner of the first position of each warp's part
    int flags;                                                  // 1: general path, 2: some event is not in the on-chip head
};

template <int KIND, bool VALUES = false>
struct CombDirect {
    CombShared<KIND, VALUES> &sh;
    CombLut lut;
    int k;

    // all threads, first thing in the kernel (nothing here depends on the offsets: it overlaps their load; the
    // barrier after that load covers it)
    __device__ __forceinline__ void prologue() const {
        if (threadIdx.x == 0) sh.flags = 0;
        owner_clear(sh.own, kCombPositions);
        if constexpr (KIND != JG_COMB_CROSS) {
            for (uint32_t i = threadIdx.x; i < lut.staged; i += kTileThreads) sh.lut[i] = __ldg(lut.table + i);
        }
    }

    // all threads, after the offsets tiles are on chip: ONE walk over the events builds their info, drops them on the
    // first chunk's owner table and registers the warp carries; then the running maximum.  Returns the flags
    // (bit 0: take the general path) after two barriers.
    __device__ __forceinline__ int setup(const EventTile &pt, const int64_t *__restrict__ s_oa,
                                         const int64_t *__restrict__ s_ob, int64_t total) const {
        int bad = (total < kCombDirectMax && (s_oa[pt.nev] - s_oa[0]) < (1ll << 32)) ? 0 : 1;
        if (KIND == JG_COMB_CROSS) bad |= (s_ob[pt.nev] - s_ob[0]) < (1ll << 32) ? 0 : 1;
        else bad |= lut.table != nullptr ? 0 : 1;               // (the table could not be allocated: un-rank everything)
        const int64_t p0 = pt.off[0], a0 = s_oa[0];
        const int stage_n = comb_lut_stage_n(KIND, k), maxn = comb_lut_max_n(KIND, k);
        const int L = static_cast<int>(min(total, static_cast<int64_t>(kCombPositions)));
        const int lg = owner_chunk_lg(L);
#pragma unroll 1
        for (int ev = threadIdx.x; ev < pt.nev; ev += kTileThreads) {
            uint4 w;
            const int64_t pa = pt.off[ev] - p0, pb = pt.off[ev + 1] - p0;
            w.x = static_cast<uint32_t>(pa);
            w.z = static_cast<uint32_t>(s_oa[ev] - a0);
            const int64_t na = s_oa[ev + 1] - s_oa[ev];
            if constexpr (KIND == JG_COMB_CROSS) {
                const int64_t nb = s_ob[ev + 1] - s_ob[ev];
                if (nb >= kCrossMaxRight) bad |= 1;
                const uint32_t nb32 = static_cast<uint32_t>(nb);
                w.x |= nb32 << 20;
                w.y = nb32 > 1 ? 0xffffffffu / nb32 + 1u : 0u;      // floor(2^32 / nb) + 1 (2^32 / nb for powers of two)
                w.w = static_cast<uint32_t>(s_ob[ev] - s_ob[0]);
            } else {
                w.y = na <= maxn ? comb_lut_base(KIND, k, static_cast<uint32_t>(na)) : kNoLut;
                w.w = static_cast<uint32_t>(na);
                if (na > stage_n && pb > pa) bad |= 2;
            }
            sh.info[ev] = w;
            owner_scatter_event(sh.own, sh.carry, ev, pa, pb, L, lg);
        }
        if (bad) atomicOr(&sh.flags, bad);
        __syncthreads();
        const int flags = sh.flags;
        if (flags & 1) return flags;
        owner_scan(sh.own, sh.carry, L, lg);
        __syncthreads();
        return flags;
    }

    // all threads: the owner table of the chunk that starts at output `cb` of the tile (cb > 0; the first chunk's
    // table comes out of setup)
    __device__ __forceinline__ void next_chunk(const EventTile &pt, int64_t cb, int L) const {
        __syncthreads();                                        // the previous chunk's table is still being read
        owner_clear(sh.own, L);
        __syncthreads();
        const int lg = owner_chunk_lg(L);
        const int64_t e0 = pt.off[0] + cb;
#pragma unroll 1
        for (int ev = threadIdx.x; ev < pt.nev; ev += kTileThreads)
            owner_scatter_event(sh.own, sh.carry, ev, pt.off[ev] - e0, pt.off[ev + 1] - e0, L, lg);
        __syncthreads();
        owner_scan(sh.own, sh.carry, L, lg);
        __syncthreads();
    }

    // local item indices of the output at tile-relative position qt, q-th of the current chunk (the info of its event
    // is returned in w).  SIMPLE: every event of the tile is in the on-chip head of the table -- no branch at all.
    template <bool SIMPLE, int NC>
    __device__ __forceinline__ void resolve(uint32_t qt, int q, uint32_t (&idx)[NC], uint4 &w) const {
        w = sh.info[sh.own[q]];
        if constexpr (KIND == JG_COMB_CROSS) {
            const uint32_t nb = w.x >> 20, r = qt - (w.x & 0xfffffu);
            const uint32_t i = nb == 1 ? r : __umulhi(r, w.y);        // r * nb < 2^32: the reciprocal is exact
            idx[0] = i;
            idx[1] = r - i * nb;
        } else {
            const uint32_t r = qt - w.x;
            if (SIMPLE || w.y != kNoLut) {
                const uint32_t at = w.y + r;
                const uint32_t code = (SIMPLE || at < lut.staged) ? sh.lut[at] : __ldg(lut.table + at);
#pragma unroll
                for (int c = 0; c < NC; ++c) idx[c] = c == NC - 1 ? code >> (6 * c) : (code >> (6 * c)) & 63u;
            } else {
                const CombIndex ci = unrank_call<KIND>(r, w.w, k);      // a call: the usual path stays lean
#pragma unroll
                for (int c = 0; c < NC; ++c) idx[c] = static_cast<uint32_t>(ci.v[c]);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// Pair table of the VALUE forms: the (left, right) item of every output position, as 16-bit positions relative to the
// tile's first item; the output-parallel sweep reads its pair with two 16-bit loads and keeps its coalesced stores.
// Tiles of short events (dimuon-like: a few pairs per event) build it EVENT-parallel -- a thread walks the pairs of
// its event in output order with two counters: 2.2 warp instructions per pair, still fewer than the direct tables need
// when an event's info serves only 2 - 4 outputs (measured, profiles/r2_comb_direct.md).  Longer runs or events are
// resolved through the direct tables chunk by chunk INTO the same pair table (comb_value_tables), so each value
// kernel has one sweep.
//   REL = true : entries are positions relative to the first item of the tile (a0 / b0) -> value gathers
//   REL = false: entries are the local indices i, j                                  -> arg forms
// ---------------------------------------------------------------------------------------------------------
constexpr int kPairTable = kCombPositions;   // entries of the left and of the right table
constexpr int kPairEventMax = 384;           // more outputs than this in one event: the direct tables instead
constexpr int kPairStride = kCombPositions + 128;    // the right-hand table follows the left-hand one (CombShared)

template <int KIND, bool REL, int STRIDE>
__device__ __forceinline__ bool build_pair_table(uint16_t *__restrict__ s_l, uint16_t *__restrict__ /* s_l + STRIDE */,
                                                 const EventTile &pt, const int64_t *__restrict__ s_oa,
                                                 const int64_t *__restrict__ s_ob) {
    const int64_t p0 = pt.off[0], total = pt.off[pt.nev] - p0;
    bool ok = total <= kPairTable;
    if (REL) {
        ok = ok && (s_oa[pt.nev] - s_oa[0]) < 65536;
        if (KIND == JG_COMB_CROSS) ok = ok && (s_ob[pt.nev] - s_ob[0]) < 65536;
    }
    for (int ev = threadIdx.x; ev < pt.nev; ev += kTileThreads) ok = ok && (pt.off[ev + 1] - pt.off[ev]) <= kPairEventMax;
    if (!__syncthreads_and(ok)) return false;
#pragma unroll 1
    for (int ev = threadIdx.x; ev < pt.nev; ev += kTileThreads) {
        const int cnt = static_cast<int>(pt.off[ev + 1] - pt.off[ev]);
        if (cnt == 0) continue;
        // one moving pointer (the right-hand entry at a fixed distance), the counters carry the event's relative start
        uint16_t *l = s_l + (pt.off[ev] - p0);
        uint16_t *const end = l + cnt;
        constexpr int rd = STRIDE;
        const int na = static_cast<int>(s_oa[ev + 1] - s_oa[ev]);
        const int ra = REL ? static_cast<int>(s_oa[ev] - s_oa[0]) : 0;
        if constexpr (KIND == JG_COMB_CROSS) {
            const int nb = static_cast<int>(s_ob[ev + 1] - s_ob[ev]);
            const int rb = REL ? static_cast<int>(s_ob[ev] - s_ob[0]) : 0;
            int i = ra, j = rb;
            const int jend = rb + nb;
#pragma unroll 1
            do {
                l[0] = static_cast<uint16_t>(i);
                l[rd] = static_cast<uint16_t>(j);
                ++l;
                if (++j == jend) {
                    j = rb;
                    ++i;
                }
            } while (l != end);
        } else {
            constexpr int kFirst = (KIND == JG_COMB_DISTINCTS) ? 1 : 0;      // j starts at i (pairs) or i + 1
            int i = ra, j = ra + kFirst;
            const int jend = ra + na;
#pragma unroll 1
            do {
                l[0] = static_cast<uint16_t>(i);
                l[rd] = static_cast<uint16_t>(j);
                ++l;
                if (++j == jend) {
                    ++i;
                    j = i + kFirst;
                }
            } while (l != end);
        }
    }
    __syncthreads();
    return true;
}

// The same table with both items in ONE 32-bit entry (left | right << 16), `shift` entries into `tab`: the short-event
// kernels read it with one load per output -- or one 64-bit load per TWO outputs, entry u = q + shift having the
// parity of its global position so that the two results go out in one vector store.  The walk keeps the entry
// itself as its counter: j + 1 is += 0x10000 and "row finished" one comparison with jend << 16.
template <int KIND, bool REL>
__device__ __forceinline__ bool build_pair_table_packed(uint32_t *__restrict__ tab, int shift, const EventTile &pt,
                                                        const int64_t *__restrict__ s_oa, const int64_t *__restrict__ s_ob) {
    const int64_t p0 = pt.off[0], total = pt.off[pt.nev] - p0;
    bool ok = total <= kPairTable;
    // (the arg forms hold local indices: an event of 65536 items has more than kPairEventMax pairs anyway)
    if (REL) {
        ok = ok && (s_oa[pt.nev] - s_oa[0]) < 65536;
        if (KIND == JG_COMB_CROSS) ok = ok && (s_ob[pt.nev] - s_ob[0]) < 65536;
    }
    for (int ev = threadIdx.x; ev < pt.nev; ev += kTileThreads) ok = ok && (pt.off[ev + 1] - pt.off[ev]) <= kPairEventMax;
    if (!__syncthreads_and(ok)) return false;
#pragma unroll 1
    for (int ev = threadIdx.x; ev < pt.nev; ev += kTileThreads) {
        const int cnt = static_cast<int>(pt.off[ev + 1] - pt.off[ev]);
        if (cnt == 0) continue;
        uint32_t *t = tab + shift + (pt.off[ev] - p0);
        uint32_t *const end = t + cnt;
        const uint32_t na = static_cast<uint32_t>(s_oa[ev + 1] - s_oa[ev]);
        const uint32_t ra = REL ? static_cast<uint32_t>(s_oa[ev] - s_oa[0]) : 0u;
        if constexpr (KIND == JG_COMB_CROSS) {
            const uint32_t nb = static_cast<uint32_t>(s_ob[ev + 1] - s_ob[ev]);
            const uint32_t rb = REL ? static_cast<uint32_t>(s_ob[ev] - s_ob[0]) : 0u;
            uint32_t e = ra | (rb << 16);
            const uint32_t lim = (rb + nb) << 16, back = (nb << 16) - 1u;      // next row: j back to rb, i + 1
#pragma unroll 1
            do {
                *t++ = e;
                e += 0x10000u;
                if (e >= lim) e -= back;
            } while (t != end);
        } else {
            constexpr uint32_t kFirst = (KIND == JG_COMB_DISTINCTS) ? 1u : 0u;      // j starts at i (pairs) or i + 1
            uint32_t i = ra, e = ra | ((ra + kFirst) << 16);
            const uint32_t lim = (ra + na) << 16;
#pragma unroll 1
            do {
                *t++ = e;
                e += 0x10000u;
                if (e >= lim) {
                    ++i;
                    e = i * 0x10001u + (kFirst << 16);
                }
            } while (t != end);
        }
    }
    __syncthreads();
    return true;
}

// two adjacent results in one streaming store
template <typename V>
__device__ __forceinline__ void st2_cs(V *p, V x, V y) {
    if constexpr (sizeof(V) == 1) __stcs(reinterpret_cast<uchar2 *>(p), make_uchar2(x, y));
    else if constexpr (sizeof(V) == 2) __stcs(reinterpret_cast<ushort2 *>(p), make_ushort2(x, y));
    else if constexpr (sizeof(V) == 4) __stcs(reinterpret_cast<uint2 *>(p), make_uint2(x, y));
    else __stcs(reinterpret_cast<ulonglong2 *>(p), make_ulonglong2(x, y));
}
template <typename V>
__device__ __forceinline__ bool aligned2(const V *p) { return (reinterpret_cast<uintptr_t>(p) & (2 * sizeof(V) - 1)) == 0; }

// The pair table(s) of a tile for the value kernels: `sweep(cb, L)` is called once per chunk of the tile's output run
// with sh.own[q] / sh.right[q] = the tile-relative left / right item of output cb + q.  false: nothing was done, the
// tile takes the caller's general path.  All threads.
template <int KIND, class Sweep>
__device__ __forceinline__ bool comb_value_tables(CombShared<KIND, true> &sh, const CombLut &lut, const EventTile &pt,
                                                  const int64_t *__restrict__ s_oa, const int64_t *__restrict__ s_ob,
                                                  Sweep sweep) {
    using Shared = CombShared<KIND, true>;
    static_assert(offsetof(Shared, right) - offsetof(Shared, own) == 2 * kPairStride, "the right-hand table must follow the left-hand one");
    const int64_t total = pt.off[pt.nev] - pt.off[0];
    if (build_pair_table<KIND, true, kPairStride>(sh.own, sh.right, pt, s_oa, s_ob)) {
        sweep(int64_t(0), static_cast<int>(total));
        return true;
    }
    // (16-bit item positions: a tile whose items do not fit takes the general path)
    const bool narrow = (s_oa[pt.nev] - s_oa[0]) < 65536 && (KIND != JG_COMB_CROSS || (s_ob[pt.nev] - s_ob[0]) < 65536);
    const CombDirect<KIND, true> direct{sh, lut, 2};
    const int flags = direct.setup(pt, s_oa, s_ob, narrow ? total : int64_t(kCombDirectMax));
    if (flags & 1) return false;
#pragma unroll 1
    for (int64_t cb = 0; cb < total; cb += kCombPositions) {
        const int L = static_cast<int>(min(static_cast<int64_t>(kCombPositions), total - cb));
        if (cb) direct.next_chunk(pt, cb, L);
        const uint32_t q0 = static_cast<uint32_t>(cb);
        // every look-up of the chunk is resolved ONCE, the left item over the owner entry it came from
        auto lookup = [&](auto simple) {
#pragma unroll 2
            for (int q = threadIdx.x; q < L; q += kTileThreads) {
                uint32_t idx[2];
                uint4 w;
                direct.template resolve<decltype(simple)::value>(q0 + q, q, idx, w);
                sh.own[q] = static_cast<uint16_t>(w.z + idx[0]);
                sh.right[q] = static_cast<uint16_t>(((KIND == JG_COMB_CROSS) ? w.w : w.z) + idx[1]);
            }
        };
        if (flags & 2) lookup(std::false_type{});
        else lookup(std::true_type{});
        __syncthreads();
        sweep(cb, L);
    }
    return true;
}

struct ColPtrs {
    int64_t *p[5];
};

template <int KIND>
__global__ void __launch_bounds__(kTileThreads)
comb_fill_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                 const int64_t *__restrict__ out_offsets, ColPtrs cols, int k, int absolute) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[kTileEvents + 1];
    EventTile pt, ta, tb;
    // the two or three offsets tiles in ONE memory round trip
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync(s_oa, offsets_a, n, blockIdx.x);
    if (KIND == JG_COMB_CROSS) tb.load_nosync(s_ob, offsets_b, n, blockIdx.x);
    __syncthreads();
    const int ncols = (KIND == JG_COMB_CHOOSE) ? k : 2;
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    __shared__ __align__(16) uint16_t s_own[kOwnerPositions + 128];
    if constexpr (KIND == JG_COMB_PAIRS || KIND == JG_COMB_CROSS) {      // (distincts: mostly 0-1 pairs, measured slower)
        // relative positions when absolute indices are wanted (the tile's first item is added back), else i, j
        uint32_t *tab = reinterpret_cast<uint32_t *>(s_own);
        const int shift = static_cast<int>(p0 & 1);
        const bool fast = absolute ? build_pair_table_packed<KIND, true>(tab, shift, pt, s_oa, s_ob)
                                   : build_pair_table_packed<KIND, false>(tab, shift, pt, s_oa, s_ob);
        if (fast) {
            const uint64_t add_l = absolute ? s_oa[0] : 0;
            const uint64_t add_r = absolute ? ((KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0]) : 0;
            const int last = static_cast<int>(p1 - p0) + shift;          // entries [shift, last)
            // two outputs per thread: one 64-bit table load, one 16-byte store per column
            unsigned long long *c0 = reinterpret_cast<unsigned long long *>(cols.p[0]) + (p0 - shift);
            unsigned long long *c1 = reinterpret_cast<unsigned long long *>(cols.p[1]) + (p0 - shift);
            const bool vec = aligned2(c0) && aligned2(c1);
#pragma unroll 2
            for (int u = 2 * threadIdx.x; u < last; u += 2 * kTileThreads) {
                const uint2 e = *reinterpret_cast<const uint2 *>(tab + u);
                const bool v0 = u >= shift, v1 = u + 1 < last;
                const unsigned long long l0 = add_l + (e.x & 0xffffu), r0 = add_r + (e.x >> 16);
                const unsigned long long l1 = add_l + (e.y & 0xffffu), r1 = add_r + (e.y >> 16);
                if (v0 && v1 && vec) {
                    st2_cs(c0 + u, l0, l1);
                    st2_cs(c1 + u, r0, r1);
                } else {
                    if (v0) {
                        __stcs(c0 + u, l0);
                        __stcs(c1 + u, r0);
                    }
                    if (v1) {
                        __stcs(c0 + u + 1, l1);
                        __stcs(c1 + u + 1, r1);
                    }
                }
            }
            return;
        }
    }
    const bool table = (p1 - p0) <= kOwnerPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(p1 - p0));
#pragma unroll 1
    for (int64_t p = p0 + threadIdx.x; p < p1; p += kTileThreads) {
        const int ev = table ? static_cast<int>(s_own[p - p0]) : pt.find_event(p);
        const int64_t na = s_oa[ev + 1] - s_oa[ev];
        const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
        CombIndex r = unrank<KIND>(p - s_po[ev], na, nb, k);
        const int64_t base_a = absolute ? s_oa[ev] : 0;
        const int64_t base_b = absolute ? ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) : 0;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (c < ncols) __stcs(cols.p[c] + p, r.v[c] + ((c == 1 && KIND == JG_COMB_CROSS) ? base_b : base_a));
        }
    }
}

// the same on the direct tables (choose, cross, and pairs / distincts of long events)
template <int KIND>
__global__ void __launch_bounds__(kTileThreads)
comb_fill_direct_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                 const int64_t *__restrict__ out_offsets, ColPtrs cols, int k, int absolute, const CombLut lut) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[KIND == JG_COMB_CROSS ? kTileEvents + 1 : 1];
    __shared__ __align__(16) CombShared<KIND> sh;
    uint16_t *const s_own = sh.own;
    const CombDirect<KIND> direct{sh, lut, k};
    direct.prologue();
    EventTile pt, ta, tb;
    // the two or three offsets tiles in ONE memory round trip
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync(s_oa, offsets_a, n, blockIdx.x);
    if (KIND == JG_COMB_CROSS) tb.load_nosync(s_ob, offsets_b, n, blockIdx.x);
    __syncthreads();
    const int ncols = (KIND == JG_COMB_CHOOSE) ? k : 2;
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    const int flags = direct.setup(pt, s_oa, s_ob, p1 - p0);
    if (!(flags & 1)) {
        const int64_t a0 = absolute ? s_oa[0] : 0, b0 = absolute ? ((KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0]) : 0;
        const uint32_t rel = absolute ? 0xffffffffu : 0u;        // local indices: the event's start is not added
#pragma unroll 1
        for (int64_t cb = 0; cb < p1 - p0; cb += kCombPositions) {
            const int L = static_cast<int>(min(static_cast<int64_t>(kCombPositions), p1 - p0 - cb));
            if (cb) direct.next_chunk(pt, cb, L);
            const uint32_t q0 = static_cast<uint32_t>(cb);
            int64_t *const c0 = cols.p[0] + p0 + cb, *const c1 = cols.p[1] + p0 + cb;
            auto sweep = [&](auto simple) {
#pragma unroll 2
                for (int q = threadIdx.x; q < L; q += kTileThreads) {
                    uint32_t idx[KIND == JG_COMB_CHOOSE ? 5 : 2];
                    uint4 w;
                    direct.template resolve<decltype(simple)::value>(q0 + q, q, idx, w);
                    const int64_t sa = a0 + (w.z & rel);
                    __stcs(c0 + q, sa + idx[0]);
                    __stcs(c1 + q, ((KIND == JG_COMB_CROSS) ? b0 + (w.w & rel) : sa) + idx[1]);
                    if constexpr (KIND == JG_COMB_CHOOSE) {
#pragma unroll
                        for (int c = 2; c < 5; ++c)
                            if (c < ncols) __stcs(cols.p[c] + p0 + cb + q, sa + idx[c]);
                    }
                }
            };
            if (flags & 2) sweep(std::false_type{});
            else sweep(std::true_type{});
        }
        return;
    }
    const bool table = (p1 - p0) <= kCombPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(p1 - p0));
#pragma unroll 1
    for (int64_t p = p0 + threadIdx.x; p < p1; p += kTileThreads) {
        const int ev = table ? static_cast<int>(s_own[p - p0]) : pt.find_event(p);
        const int64_t na = s_oa[ev + 1] - s_oa[ev];
        const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
        CombIndex r = unrank<KIND>(p - s_po[ev], na, nb, k);
        const int64_t base_a = absolute ? s_oa[ev] : 0;
        const int64_t base_b = absolute ? ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) : 0;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            if (c < ncols) __stcs(cols.p[c] + p, r.v[c] + ((c == 1 && KIND == JG_COMB_CROSS) ? base_b : base_a));
        }
    }
}

// fused value form: out_l[p] = content_a[start_a + i], out_r[p] = content_b[start_b + j]
template <int KIND, typename VA, typename VB>
__global__ void __launch_bounds__(kTileThreads)
comb_gather2_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                    const int64_t *__restrict__ out_offsets, const VA *__restrict__ content_a,
                    const VB *__restrict__ content_b, VA *__restrict__ out_l, VB *__restrict__ out_r) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[kTileEvents + 1];
    EventTile pt, ta, tb;
    // the two or three offsets tiles in ONE memory round trip
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync(s_oa, offsets_a, n, blockIdx.x);
    if (KIND == JG_COMB_CROSS) tb.load_nosync(s_ob, offsets_b, n, blockIdx.x);
    __syncthreads();
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    __shared__ __align__(16) uint16_t s_own[kOwnerPositions + 128];
    {
        // (staging the tile's content runs in shared memory by TMA was measured: slower -- the loads below hit L1)
        uint32_t *tab = reinterpret_cast<uint32_t *>(s_own);
        const int shift = static_cast<int>(p0 & 1);
        if (build_pair_table_packed<KIND, true>(tab, shift, pt, s_oa, s_ob)) {
            const VA *ca = content_a + s_oa[0];
            const VB *cb = content_b + ((KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0]);
            // two outputs per thread and step (one 64-bit table load, one vector store per column), two steps in flight
            VA *ol = out_l + (p0 - shift);
            VB *orr = out_r + (p0 - shift);
            const bool vec = aligned2(ol) && aligned2(orr);
            const int last = static_cast<int>(p1 - p0) + shift;          // entries [shift, last)
#pragma unroll 1
            for (int ub = 2 * threadIdx.x; ub < last; ub += 4 * kTileThreads) {
                VA va[2][2];
                VB vb[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int u = ub + h * 2 * kTileThreads;
                    if (u < last) {
                        const uint2 e = *reinterpret_cast<const uint2 *>(tab + u);
                        if (u >= shift) {
                            va[h][0] = __ldg(ca + (e.x & 0xffffu));
                            vb[h][0] = __ldg(cb + (e.x >> 16));
                        }
                        if (u + 1 < last) {
                            va[h][1] = __ldg(ca + (e.y & 0xffffu));
                            vb[h][1] = __ldg(cb + (e.y >> 16));
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int u = ub + h * 2 * kTileThreads;
                    if (u < last) {
                        const bool v0 = u >= shift, v1 = u + 1 < last;
                        if (v0 && v1 && vec) {
                            st2_cs(ol + u, va[h][0], va[h][1]);
                            st2_cs(orr + u, vb[h][0], vb[h][1]);
                        } else {
                            if (v0) {
                                __stcs(ol + u, va[h][0]);
                                __stcs(orr + u, vb[h][0]);
                            }
                            if (v1) {
                                __stcs(ol + u + 1, va[h][1]);
                                __stcs(orr + u + 1, vb[h][1]);
                            }
                        }
                    }
                }
            }
            return;
        }
    }
    const bool table = (p1 - p0) <= kOwnerPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(p1 - p0));
#pragma unroll 1
    for (int64_t pb = p0 + threadIdx.x; pb < p1; pb += 2 * kTileThreads) {
        int64_t ia[2], ib[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t p = pb + u * kTileThreads;
            if (p < p1) {
                const int ev = table ? static_cast<int>(s_own[p - p0]) : pt.find_event(p);
                const int64_t na = s_oa[ev + 1] - s_oa[ev];
                const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
                CombIndex r = unrank<KIND>(p - s_po[ev], na, nb, 2);
                ia[u] = s_oa[ev] + r.v[0];
                ib[u] = ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) + r.v[1];
            }
        }
        VA va[2];
        VB vb[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (pb + u * kTileThreads < p1) {
                va[u] = __ldg(content_a + ia[u]);
                vb[u] = __ldg(content_b + ib[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t p = pb + u * kTileThreads;
            if (p < p1) {
                __stcs(out_l + p, va[u]);
                __stcs(out_r + p, vb[u]);
            }
        }
    }
}

// the same for long events: pair tables chunk by chunk out of the direct tables
template <int KIND, typename VA, typename VB>
__global__ void __launch_bounds__(kTileThreads, 5)
comb_gather2_direct_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                    const int64_t *__restrict__ out_offsets, const VA *__restrict__ content_a,
                    const VB *__restrict__ content_b, VA *__restrict__ out_l, VB *__restrict__ out_r,
                    const CombLut lut) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[KIND == JG_COMB_CROSS ? kTileEvents + 1 : 1];
    __shared__ __align__(16) CombShared<KIND, true> sh;
    uint16_t *const s_own = sh.own;
    CombDirect<KIND, true>{sh, lut, 2}.prologue();
    EventTile pt, ta, tb;
    // the two or three offsets tiles in ONE memory round trip, and the items the pairs will gather already on their way
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync<sizeof(VA)>(s_oa, offsets_a, n, blockIdx.x, content_a);
    if (KIND == JG_COMB_CROSS) tb.load_nosync<sizeof(VB)>(s_ob, offsets_b, n, blockIdx.x, content_b);
    __syncthreads();
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    {
        // (staging the tile's content runs in shared memory by TMA was measured: slower -- the loads below hit L1)
        const VA *ca = content_a + s_oa[0];
        const VB *cb_ = content_b + ((KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0]);
        if (comb_value_tables<KIND>(sh, lut, pt, s_oa, s_ob, [&](int64_t cb, int L) {
                VA *ol = out_l + p0 + cb;
                VB *orr = out_r + p0 + cb;
#pragma unroll 1
                for (int qb = threadIdx.x; qb < L; qb += 4 * kTileThreads) {
                    VA va[4];
                    VB vb[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int q = qb + u * kTileThreads;
                        if (q < L) {
                            va[u] = __ldg(ca + sh.own[q]);
                            vb[u] = __ldg(cb_ + sh.right[q]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int q = qb + u * kTileThreads;
                        if (q < L) {
                            __stcs(ol + q, va[u]);
                            __stcs(orr + q, vb[u]);
                        }
                    }
                }
            }))
            return;
    }
    const bool table = (p1 - p0) <= kCombPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(p1 - p0));
#pragma unroll 1
    for (int64_t pb = p0 + threadIdx.x; pb < p1; pb += 2 * kTileThreads) {
        int64_t ia[2], ib[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t p = pb + u * kTileThreads;
            if (p < p1) {
                const int ev = table ? static_cast<int>(s_own[p - p0]) : pt.find_event(p);
                const int64_t na = s_oa[ev + 1] - s_oa[ev];
                const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
                CombIndex r = unrank<KIND>(p - s_po[ev], na, nb, 2);
                ia[u] = s_oa[ev] + r.v[0];
                ib[u] = ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) + r.v[1];
            }
        }
        VA va[2];
        VB vb[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (pb + u * kTileThreads < p1) {
                va[u] = __ldg(content_a + ia[u]);
                vb[u] = __ldg(content_b + ib[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int64_t p = pb + u * kTileThreads;
            if (p < p1) {
                __stcs(out_l + p, va[u]);
                __stcs(out_r + p, vb[u]);
            }
        }
    }
}

// Record form (Table content, jagged.py:1294 / 1350 with table.py:619-629 row views): every leaf column of the
// left and of the right collection is gathered in ONE sweep, so the owner lookup and the un-ranking are paid once
// per output pair instead of once per column.
struct MultiCols {
    const void *in_l[4];
    void *out_l[4];
    const void *in_r[4];
    void *out_r[4];
    int nl, nr;
};

template <int KIND, typename V>
__global__ void __launch_bounds__(kTileThreads)
comb_gather_multi_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                         const int64_t *__restrict__ out_offsets, MultiCols cols) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[kTileEvents + 1];
    __shared__ __align__(16) uint16_t s_own[kOwnerPositions + 128];
    EventTile pt, ta, tb;
    // the two or three offsets tiles in ONE memory round trip
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync(s_oa, offsets_a, n, blockIdx.x);
    if (KIND == JG_COMB_CROSS) tb.load_nosync(s_ob, offsets_b, n, blockIdx.x);
    __syncthreads();
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    {
        uint32_t *tab = reinterpret_cast<uint32_t *>(s_own);
        if (build_pair_table_packed<KIND, true>(tab, 0, pt, s_oa, s_ob)) {
            const int64_t a0 = s_oa[0], b0 = (KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0];
            const int total = static_cast<int>(p1 - p0);
#pragma unroll 1
            for (int q = threadIdx.x; q < total; q += kTileThreads) {
                const uint32_t e = tab[q];
                const int64_t ia = a0 + (e & 0xffffu), ib = b0 + (e >> 16);
                V vl[4], vr[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < cols.nl) vl[c] = __ldg(static_cast<const V *>(cols.in_l[c]) + ia);
                    if (c < cols.nr) vr[c] = __ldg(static_cast<const V *>(cols.in_r[c]) + ib);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < cols.nl) __stcs(static_cast<V *>(cols.out_l[c]) + p0 + q, vl[c]);
                    if (c < cols.nr) __stcs(static_cast<V *>(cols.out_r[c]) + p0 + q, vr[c]);
                }
            }
            return;
        }
    }
    const bool table = (p1 - p0) <= kOwnerPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(p1 - p0));
#pragma unroll 1
    for (int64_t p = p0 + threadIdx.x; p < p1; p += kTileThreads) {
        const int ev = table ? static_cast<int>(s_own[p - p0]) : pt.find_event(p);
        const int64_t na = s_oa[ev + 1] - s_oa[ev];
        const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
        CombIndex r = unrank<KIND>(p - s_po[ev], na, nb, 2);
        const int64_t ia = s_oa[ev] + r.v[0];
        const int64_t ib = ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) + r.v[1];
        V vl[4], vr[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < cols.nl) vl[c] = __ldg(static_cast<const V *>(cols.in_l[c]) + ia);
            if (c < cols.nr) vr[c] = __ldg(static_cast<const V *>(cols.in_r[c]) + ib);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < cols.nl) __stcs(static_cast<V *>(cols.out_l[c]) + p, vl[c]);
            if (c < cols.nr) __stcs(static_cast<V *>(cols.out_r[c]) + p, vr[c]);
        }
    }
}

template <int KIND, typename V>
__global__ void __launch_bounds__(kTileThreads, 5)
comb_gather_multi_direct_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                         const int64_t *__restrict__ out_offsets, MultiCols cols, const CombLut lut) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[KIND == JG_COMB_CROSS ? kTileEvents + 1 : 1];
    __shared__ __align__(16) CombShared<KIND, true> sh;
    uint16_t *const s_own = sh.own;
    CombDirect<KIND, true>{sh, lut, 2}.prologue();
    EventTile pt, ta, tb;
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync<sizeof(V)>(s_oa, offsets_a, n, blockIdx.x, cols.nl > 0 ? cols.in_l[0] : cols.in_r[0]);
    if (KIND == JG_COMB_CROSS) tb.load_nosync<sizeof(V)>(s_ob, offsets_b, n, blockIdx.x, cols.nr > 0 ? cols.in_r[0] : cols.in_l[0]);
    __syncthreads();
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    {
        const int64_t a0 = s_oa[0], b0 = (KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0];
        if (comb_value_tables<KIND>(sh, lut, pt, s_oa, s_ob, [&](int64_t cb, int L) {
#pragma unroll 1
                for (int q = threadIdx.x; q < L; q += kTileThreads) {
                    const int64_t ia = a0 + sh.own[q], ib = b0 + sh.right[q];
                    V vl[4], vr[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c < cols.nl) vl[c] = __ldg(static_cast<const V *>(cols.in_l[c]) + ia);
                        if (c < cols.nr) vr[c] = __ldg(static_cast<const V *>(cols.in_r[c]) + ib);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c < cols.nl) __stcs(static_cast<V *>(cols.out_l[c]) + p0 + cb + q, vl[c]);
                        if (c < cols.nr) __stcs(static_cast<V *>(cols.out_r[c]) + p0 + cb + q, vr[c]);
                    }
                }
            }))
            return;
    }
    const bool table = (p1 - p0) <= kCombPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(p1 - p0));
#pragma unroll 1
    for (int64_t p = p0 + threadIdx.x; p < p1; p += kTileThreads) {
        const int ev = table ? static_cast<int>(s_own[p - p0]) : pt.find_event(p);
        const int64_t na = s_oa[ev + 1] - s_oa[ev];
        const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
        CombIndex r = unrank<KIND>(p - s_po[ev], na, nb, 2);
        const int64_t ia = s_oa[ev] + r.v[0];
        const int64_t ib = ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) + r.v[1];
        V vl[4], vr[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < cols.nl) vl[c] = __ldg(static_cast<const V *>(cols.in_l[c]) + ia);
            if (c < cols.nr) vr[c] = __ldg(static_cast<const V *>(cols.in_r[c]) + ib);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < cols.nl) __stcs(static_cast<V *>(cols.out_l[c]) + p, vl[c]);
            if (c < cols.nr) __stcs(static_cast<V *>(cols.out_r[c]) + p, vr[c]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused per-combination expression (SURVEY 8d "pair mass fused from index generation"): the value form of
// pairs / distincts / cross followed by a chain of ufuncs, e.g.
//     p = mu.pairs();  sqrt(2 * p.i0.pt * p.i1.pt * (cosh(p.i0.eta - p.i1.eta) - cos(p.i0.phi - p.i1.phi)))
// is evaluated in ONE sweep: every output position finds its (left, right) item exactly like comb_gather2_kernel,
// loads the columns the expression names straight from the collections and runs a small program; the
// 2 x columns x P gathered temporaries and the P-sized intermediate of every ufunc are never written.
// The host (awkward0_b200/lazy.py) compiles the expression DAG into an ACCUMULATOR program, one 32-bit word per
// step:   op | operand kind << 8 | reverse << 12 | operand index << 16
// A step combines the accumulator (a register) with one operand -- a column of the left / right collection, a
// P-sized column, a constant or a spill slot -- using the SAME functors as the eager kernels (jg_ops.cuh: both
// paths round identically); unary steps act on the accumulator alone; JG_EV_LOAD / JG_EV_STORE move between the
// accumulator and a spill slot (a per-thread column of shared memory) when the expression is not a chain.  Every
// step is executed for U positions of the thread at once: the decode is amortised over U results and U
// independent dependency chains are in flight.
// ---------------------------------------------------------------------------------------------------------
// Positions are relative to the tile (a0 / b0: first item of the tile in the left / right collection, p0: first
// output of the tile): IX = uint32_t on the usual path (they come from the 16-bit pair table), int64_t otherwise.
template <typename T, int U, typename IX>
__device__ __forceinline__ void eval_program(const JgEvalProgram &pr, T *__restrict__ slots, int64_t a0, int64_t b0,
                                             int64_t p0, const IX (&ia)[U], const IX (&ib)[U], const IX (&ip)[U],
                                             T (&acc)[U]) {
    constexpr int S = kTileThreads;        // stride between the U copies of one slot
#define JG_EV_BIN(OPC)                                                                                      \
    case OPC:                                                                                               \
        _Pragma("unroll") for (int u = 0; u < U; ++u)                                                       \
            acc[u] = rev ? BinaryOp<OPC, T>::apply(b[u], acc[u]) : BinaryOp<OPC, T>::apply(acc[u], b[u]);   \
        break;
#define JG_EV_CMP(OPC)                                                                                      \
    case OPC:                                                                                               \
        _Pragma("unroll") for (int u = 0; u < U; ++u)                                                       \
            acc[u] = T(rev ? CompareOp<OPC, T>::apply(b[u], acc[u]) : CompareOp<OPC, T>::apply(acc[u], b[u])); \
        break;
#define JG_EV_UN(OPC)                                                                         \
    case OPC:                                                                                 \
        _Pragma("unroll") for (int u = 0; u < U; ++u) acc[u] = UnaryOp<OPC, T>::apply(acc[u]); \
        break;
#pragma unroll 1
    for (int pc = 0; pc < pr.nins; ++pc) {
        const uint32_t w = pr.ins[pc];
        const int op = w & 255, kind = (w >> 8) & 15, idx = w >> 16;
        const bool rev = (w >> 12) & 1;
        if (op == JG_EV_STORE) {
#pragma unroll
            for (int u = 0; u < U; ++u) slots[(idx * U + u) * S] = acc[u];
            continue;
        }
        T b[U];
        if (op < JG_OP_NEG) {              // binary / comparison / JG_EV_LOAD: fetch the operand
            switch (kind) {
                case JG_EV_LEFT: {
                    const T *col = static_cast<const T *>(pr.left[idx]) + a0;
#pragma unroll
                    for (int u = 0; u < U; ++u) b[u] = __ldg(col + ia[u]);
                    break;
                }
                case JG_EV_RIGHT: {
                    const T *col = static_cast<const T *>(pr.right[idx]) + b0;
#pragma unroll
                    for (int u = 0; u < U; ++u) b[u] = __ldg(col + ib[u]);
                    break;
                }
                case JG_EV_POS: {
                    const T *col = static_cast<const T *>(pr.pos[idx]) + p0;
#pragma unroll
                    for (int u = 0; u < U; ++u) b[u] = __ldcs(col + ip[u]);
                    break;
                }
                case JG_EV_CONST: {
                    const T c = static_cast<T>(pr.consts[idx]);
#pragma unroll
                    for (int u = 0; u < U; ++u) b[u] = c;
                    break;
                }
                default: {                 // JG_EV_SLOT
#pragma unroll
                    for (int u = 0; u < U; ++u) b[u] = slots[(idx * U + u) * S];
                    break;
                }
            }
        }
        if (op < JG_OP_NEG) {
            switch (op) {                  // operand steps
                case JG_EV_LOAD:
#pragma unroll
                    for (int u = 0; u < U; ++u) acc[u] = b[u];
                    break;
                JG_EV_BIN(JG_OP_ADD) JG_EV_BIN(JG_OP_SUB) JG_EV_BIN(JG_OP_MUL) JG_EV_BIN(JG_OP_DIV)
                JG_EV_BIN(JG_OP_FLOORDIV) JG_EV_BIN(JG_OP_MOD) JG_EV_BIN(JG_OP_POW) JG_EV_BIN(JG_OP_MAXIMUM)
                JG_EV_BIN(JG_OP_MINIMUM) JG_EV_BIN(JG_OP_ARCTAN2) JG_EV_BIN(JG_OP_HYPOT)
                JG_EV_CMP(JG_OP_GT) JG_EV_CMP(JG_OP_GE) JG_EV_CMP(JG_OP_LT) JG_EV_CMP(JG_OP_LE)
                JG_EV_CMP(JG_OP_EQ) JG_EV_CMP(JG_OP_NE)
                default: break;            // rejected on the host (eval_op_class)
            }
            continue;
        }
        switch (op) {                      // unary steps
            JG_EV_UN(JG_OP_NEG) JG_EV_UN(JG_OP_ABS) JG_EV_UN(JG_OP_SQRT) JG_EV_UN(JG_OP_EXP) JG_EV_UN(JG_OP_LOG)
            JG_EV_UN(JG_OP_SIN) JG_EV_UN(JG_OP_COS) JG_EV_UN(JG_OP_TAN) JG_EV_UN(JG_OP_SINH) JG_EV_UN(JG_OP_COSH)
            JG_EV_UN(JG_OP_TANH) JG_EV_UN(JG_OP_SQUARE) JG_EV_UN(JG_OP_CAST) JG_EV_UN(JG_OP_ARCSINH)
            JG_EV_UN(JG_OP_FLOOR) JG_EV_UN(JG_OP_CEIL) JG_EV_UN(JG_OP_RINT) JG_EV_UN(JG_OP_SIGN)
            JG_EV_UN(JG_OP_LOG10) JG_EV_UN(JG_OP_EXP2) JG_EV_UN(JG_OP_ARCSIN) JG_EV_UN(JG_OP_ARCCOS)
            JG_EV_UN(JG_OP_ARCTAN) JG_EV_UN(JG_OP_LOG2) JG_EV_UN(JG_OP_LOG1P) JG_EV_UN(JG_OP_EXPM1)
            JG_EV_UN(JG_OP_RECIPROCAL)
            default: break;      // rejected on the host (eval_op_class)
        }
    }
#undef JG_EV_BIN
#undef JG_EV_CMP
#undef JG_EV_UN
}

// 0: not supported, 1: takes an operand (binary, comparison, load), 2: unary, 3: store
static int eval_op_class(int op) {
    if (op == JG_EV_LOAD) return 1;
    if (op == JG_EV_STORE) return 3;
    if (op >= JG_OP_ADD && op <= JG_OP_HYPOT) return 1;
    if (op >= JG_OP_GT && op <= JG_OP_NE) return 1;
    switch (op) {
        case JG_OP_NEG: case JG_OP_ABS: case JG_OP_SQRT: case JG_OP_EXP: case JG_OP_LOG: case JG_OP_SIN:
        case JG_OP_COS: case JG_OP_TAN: case JG_OP_SINH: case JG_OP_COSH: case JG_OP_TANH: case JG_OP_SQUARE:
        case JG_OP_CAST: case JG_OP_ARCSINH: case JG_OP_FLOOR: case JG_OP_CEIL: case JG_OP_RINT: case JG_OP_SIGN:
        case JG_OP_LOG10: case JG_OP_EXP2: case JG_OP_ARCSIN: case JG_OP_ARCCOS: case JG_OP_ARCTAN: case JG_OP_LOG2:
        case JG_OP_LOG1P: case JG_OP_EXPM1: case JG_OP_RECIPROCAL:
            return 2;
    }
    return 0;
}

// The batches of one tile: a batch is U rows of `stride` consecutive positions and thread t owns position t of
// every row.  Full batches use stride = 256; the last batch shrinks the rows (to a multiple of 32) so that the
// warps that stay in it have all U positions valid and the others leave: a nearly empty tail costs what it holds,
// not a full batch.  `resolve(q, ia, ib)` gives the tile-relative positions of the two items of output q.
template <typename T, int U, typename IX, class Resolve>
__device__ __forceinline__ void eval_tile(const JgEvalProgram &pr, T *__restrict__ slots, int64_t a0, int64_t b0,
                                          int64_t p0, int64_t total, void *__restrict__ out, Resolve resolve) {
#pragma unroll 1
    for (int64_t base = 0; base < total; base += U * kTileThreads) {
        const int64_t left = total - base;
        const int stride = left >= U * kTileThreads ? kTileThreads
                                                    : static_cast<int>(((left + U - 1) / U + 31) & ~int64_t(31));
        if (static_cast<int>(threadIdx.x) >= stride) break;          // warp-uniform (stride is a multiple of 32)
        IX ia[U], ib[U], ip[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int64_t q = base + u * stride + threadIdx.x;
            if (q >= total) q = total - 1;            // spare lanes of the last row recompute the last position
            ip[u] = static_cast<IX>(q);
            resolve(q, ia[u], ib[u]);
        }
        T acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = T(0);
        eval_program<T, U, IX>(pr, slots, a0, b0, p0, ia, ib, ip, acc);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = base + u * stride + threadIdx.x;
            if (q < total) {
                if (pr.out_bool) __stcs(static_cast<uint8_t *>(out) + p0 + q, static_cast<uint8_t>(acc[u] != T(0)));
                else __stcs(static_cast<T *>(out) + p0 + q, acc[u]);
            }
        }
    }
}

template <int KIND, typename T, int U>       // U: positions per thread and step on the pair-table path
__global__ void __launch_bounds__(kTileThreads)
comb_eval_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                 const int64_t *__restrict__ out_offsets, const __grid_constant__ JgEvalProgram pr,
                 void *__restrict__ out) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[kTileEvents + 1];
    __shared__ __align__(16) uint16_t s_own[kOwnerPositions + 128];
    extern __shared__ __align__(16) unsigned char s_dyn[];        // spill slots: nslots x U x 256 values of T
    T *slots = reinterpret_cast<T *>(s_dyn) + threadIdx.x;
    EventTile pt, ta, tb;
    // the two or three offsets tiles in ONE memory round trip
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync(s_oa, offsets_a, n, blockIdx.x);
    if (KIND == JG_COMB_CROSS) tb.load_nosync(s_ob, offsets_b, n, blockIdx.x);
    __syncthreads();
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    const int64_t total = p1 - p0;
    if (total <= 0) return;
    const int64_t a0 = s_oa[0], b0 = (KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0];
    uint32_t *tab = reinterpret_cast<uint32_t *>(s_own);
    if (build_pair_table_packed<KIND, true>(tab, 0, pt, s_oa, s_ob)) {
        eval_tile<T, U, uint32_t>(pr, slots, a0, b0, p0, total, out, [&](int64_t q, uint32_t &ia, uint32_t &ib) {
            const uint32_t e = tab[q];
            ia = e & 0xffffu;
            ib = e >> 16;
        });
        return;
    }
    // long events: owner table (binary search for very long runs) + un-ranking, 64-bit positions
    const bool table = total <= kOwnerPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(total));
    eval_tile<T, 2, int64_t>(pr, slots, a0, b0, p0, total, out, [&](int64_t q, int64_t &ia, int64_t &ib) {
        const int ev = table ? static_cast<int>(s_own[q]) : pt.find_event(p0 + q);
        const int64_t na = s_oa[ev + 1] - s_oa[ev];
        const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
        CombIndex r = unrank<KIND>(p0 + q - s_po[ev], na, nb, 2);
        ia = s_oa[ev] - a0 + r.v[0];
        ib = ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) - b0 + r.v[1];
    });
}

template <int KIND, typename T, int U>       // U: positions per thread and step on the pair-table path
__global__ void __launch_bounds__(kTileThreads, 4)        // (64 registers)
comb_eval_direct_kernel(const int64_t *__restrict__ offsets_a, const int64_t *__restrict__ offsets_b, int64_t n,
                 const int64_t *__restrict__ out_offsets, const __grid_constant__ JgEvalProgram pr,
                 void *__restrict__ out, const CombLut lut) {
    __shared__ int64_t s_po[kTileEvents + 1];
    __shared__ int64_t s_oa[kTileEvents + 1];
    __shared__ int64_t s_ob[KIND == JG_COMB_CROSS ? kTileEvents + 1 : 1];
    __shared__ __align__(16) CombShared<KIND, true> sh;
    uint16_t *const s_own = sh.own;
    extern __shared__ __align__(16) unsigned char s_dyn[];        // spill slots: nslots x U x 256 values of T
    T *slots = reinterpret_cast<T *>(s_dyn) + threadIdx.x;
    CombDirect<KIND, true>{sh, lut, 2}.prologue();
    EventTile pt, ta, tb;
    pt.load_nosync(s_po, out_offsets, n, blockIdx.x);
    ta.load_nosync(s_oa, offsets_a, n, blockIdx.x);
    if (KIND == JG_COMB_CROSS) tb.load_nosync(s_ob, offsets_b, n, blockIdx.x);
    __syncthreads();
    const int64_t p0 = pt.first_element(), p1 = pt.last_element();
    const int64_t total = p1 - p0;
    if (total <= 0) return;
    const int64_t a0 = s_oa[0], b0 = (KIND == JG_COMB_CROSS) ? s_ob[0] : s_oa[0];
    // the U positions a thread evaluates together cost two 16-bit loads each: resolving the direct tables inside the
    // interpreter's loop took 122 registers (or spills at 64)
    if (comb_value_tables<KIND>(sh, lut, pt, s_oa, s_ob, [&](int64_t cb, int L) {
            eval_tile<T, U, uint32_t>(pr, slots, a0, b0, p0 + cb, L, out, [&](int64_t q, uint32_t &ia, uint32_t &ib) {
                ia = sh.own[q];
                ib = sh.right[q];
            });
        }))
        return;
    // runs that do not fit the packed event info: owner table (binary search for very long runs) + un-ranking, 64-bit positions
    const bool table = total <= kCombPositions;
    if (table) build_owner_table(s_own, pt, static_cast<int>(total));
    eval_tile<T, 2, int64_t>(pr, slots, a0, b0, p0, total, out, [&](int64_t q, int64_t &ia, int64_t &ib) {
        const int ev = table ? static_cast<int>(s_own[q]) : pt.find_event(p0 + q);
        const int64_t na = s_oa[ev + 1] - s_oa[ev];
        const int64_t nb = (KIND == JG_COMB_CROSS) ? (s_ob[ev + 1] - s_ob[ev]) : 0;
        CombIndex r = unrank<KIND>(p0 + q - s_po[ev], na, nb, 2);
        ia = s_oa[ev] - a0 + r.v[0];
        ib = ((KIND == JG_COMB_CROSS) ? s_ob[ev] : s_oa[ev]) - b0 + r.v[1];
    });
}

template <int KIND, typename T, int U>
static int launch_comb_eval(bool direct, const int64_t *oa, const int64_t *ob, int64_t n, const int64_t *po,
                            const JgEvalProgram &pr, void *out, cudaStream_t s) {
    const size_t dyn = static_cast<size_t>(pr.nslots) * U * kTileThreads * sizeof(T);
    static bool configured = false;          // per instantiation
    if (!configured) {
        const int most = static_cast<int>(JG_EVAL_MAX_SLOTS * U * kTileThreads * sizeof(T));   // 80 KB
        cudaFuncSetAttribute(comb_eval_kernel<KIND, T, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        cudaFuncSetAttribute(comb_eval_direct_kernel<KIND, T, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        configured = true;
    }
    if (direct) comb_eval_direct_kernel<KIND, T, U><<<tile_grid(n), kTileThreads, dyn, s>>>(oa, ob, n, po, pr, out, comb_lut_for(KIND, 2, s));
    else comb_eval_kernel<KIND, T, U><<<tile_grid(n), kTileThreads, dyn, s>>>(oa, ob, n, po, pr, out);
    return check_launch("comb_eval_kernel");
}

template <int KIND>
static int launch_comb_eval_dtype(bool direct, int dtype, const int64_t *oa, const int64_t *ob, int64_t n, const int64_t *po,
                                  const JgEvalProgram &pr, void *out, cudaStream_t s) {
    if (dtype == JG_F32) return launch_comb_eval<KIND, float, 8>(direct, oa, ob, n, po, pr, out, s);
    if (dtype == JG_F64) return launch_comb_eval<KIND, double, 4>(direct, oa, ob, n, po, pr, out, s);
    set_error("jg_comb_eval: compute dtype must be float32 or float64 (got %d)", dtype);
    return -2;
}

template <int KIND, typename V>
static void launch_gather_multi_v(bool direct, const int64_t *oa, const int64_t *ob, int64_t n, const int64_t *po,
                                  const MultiCols &mc, cudaStream_t s) {
    const unsigned g = tile_grid(n);
    if (direct) comb_gather_multi_direct_kernel<KIND, V><<<g, kTileThreads, 0, s>>>(oa, ob, n, po, mc, comb_lut_for(KIND, 2, s));
    else comb_gather_multi_kernel<KIND, V><<<g, kTileThreads, 0, s>>>(oa, ob, n, po, mc);
}

template <int KIND>
static int launch_gather_multi(bool direct, int itemsize, const int64_t *oa, const int64_t *ob, int64_t n, const int64_t *po,
                               const MultiCols &mc, cudaStream_t s) {
    switch (itemsize) {
        case 1: launch_gather_multi_v<KIND, uint8_t>(direct, oa, ob, n, po, mc, s); break;
        case 2: launch_gather_multi_v<KIND, uint16_t>(direct, oa, ob, n, po, mc, s); break;
        case 4: launch_gather_multi_v<KIND, uint32_t>(direct, oa, ob, n, po, mc, s); break;
        case 8: launch_gather_multi_v<KIND, uint64_t>(direct, oa, ob, n, po, mc, s); break;
        default: set_error("jg_comb_gather_multi: unsupported itemsize %d", itemsize); return -2;
    }
    return check_launch("comb_gather_multi_kernel");
}

template <int KIND, typename VA>
static int gather2_b(bool direct, int itemsize_b, const int64_t *oa, const int64_t *ob, int64_t n, const int64_t *po,
                     const void *ca, const void *cb, void *l, void *r, cudaStream_t s) {
    const unsigned g = tile_grid(n);
    const CombLut lut = direct ? comb_lut_for(KIND, 2, s) : CombLut{nullptr, 0};
#define JG_G2(VB)                                                                                                     \
    if (direct)                                                                                                       \
        comb_gather2_direct_kernel<KIND, VA, VB><<<g, kTileThreads, 0, s>>>(                                          \
            oa, ob, n, po, static_cast<const VA *>(ca), static_cast<const VB *>(cb), static_cast<VA *>(l),            \
            static_cast<VB *>(r), lut);                                                                               \
    else                                                                                                              \
        comb_gather2_kernel<KIND, VA, VB><<<g, kTileThreads, 0, s>>>(oa, ob, n, po, static_cast<const VA *>(ca),       \
                                                                     static_cast<const VB *>(cb), static_cast<VA *>(l), \
                                                                     static_cast<VB *>(r))
    switch (itemsize_b) {
        case 1: JG_G2(uint8_t); break;
        case 2: JG_G2(uint16_t); break;
        case 4: JG_G2(uint32_t); break;
        case 8: JG_G2(uint64_t); break;
        default: set_error("jg_comb_gather2: unsupported itemsize %d", itemsize_b); return -2;
    }
#undef JG_G2
    return check_launch("comb_gather2_kernel");
}

template <int KIND>
static int gather2_a(bool direct, int itemsize_a, int itemsize_b, const int64_t *oa, const int64_t *ob, int64_t n,
                     const int64_t *po, const void *ca, const void *cb, void *l, void *r, cudaStream_t s) {
    switch (itemsize_a) {
        case 1: return gather2_b<KIND, uint8_t>(direct, itemsize_b, oa, ob, n, po, ca, cb, l, r, s);
        case 2: return gather2_b<KIND, uint16_t>(direct, itemsize_b, oa, ob, n, po, ca, cb, l, r, s);
        case 4: return gather2_b<KIND, uint32_t>(direct, itemsize_b, oa, ob, n, po, ca, cb, l, r, s);
        case 8: return gather2_b<KIND, uint64_t>(direct, itemsize_b, oa, ob, n, po, ca, cb, l, r, s);
    }
    set_error("jg_comb_gather2: unsupported itemsize %d", itemsize_a);
    return -2;
}

template <int KIND>
static void launch_comb_fill(bool direct, const int64_t *oa, const int64_t *ob, int64_t n, const int64_t *po, const ColPtrs &cp,
                             int k, int absolute, cudaStream_t s) {
    const unsigned g = tile_grid(n);
    if (direct) comb_fill_direct_kernel<KIND><<<g, kTileThreads, 0, s>>>(oa, ob, n, po, cp, k, absolute, comb_lut_for(KIND, k, s));
    else comb_fill_kernel<KIND><<<g, kTileThreads, 0, s>>>(oa, ob, n, po, cp, k, absolute);
}

}  // namespace jg

using namespace jg;

extern "C" {

int jg_comb_offsets(int kind, int k, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n,
                    int64_t *out_offsets, int64_t *total, void *ws, size_t ws_bytes, void *stream) {
    JG_REQUIRE(kind >= JG_COMB_PAIRS && kind <= JG_COMB_CROSS, "jg_comb_offsets: bad kind %d", kind);
    JG_REQUIRE(kind != JG_COMB_CHOOSE || (k >= 2 && k <= 5), "jg_comb_offsets: choose needs 2 <= k <= 5");
    JG_REQUIRE(kind != JG_COMB_CROSS || offsets_b != nullptr, "jg_comb_offsets: cross needs offsets_b");
    // the offsets tile(s) are parked in shared memory by TMA and the per-event count is formed on the fly
    if (kind == JG_COMB_CROSS)
        return launch_scan_tma<int64_t, XfComb<2>, kScanTmaItems / 2>(XfComb<2>{kind, k}, offsets_a, offsets_b, n,
                                                                      out_offsets, total, nullptr, ws, ws_bytes,
                                                                      as_stream(stream));
    return launch_scan_tma<int64_t, XfComb<1>, kScanTmaItems>(XfComb<1>{kind, k}, offsets_a,
                                                              static_cast<const int64_t *>(nullptr), n, out_offsets,
                                                              total, nullptr, ws, ws_bytes, as_stream(stream));
}

int jg_comb_fill(int kind, int k, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n,
                 const int64_t *out_offsets, int64_t *const *cols, int absolute, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const bool long_events = (kind & JG_COMB_LONG) != 0;
    kind &= ~JG_COMB_LONG;
    const int ncols = kind == JG_COMB_CHOOSE ? k : 2;
    JG_REQUIRE(ncols >= 2 && ncols <= 5, "jg_comb_fill: bad column count %d", ncols);
    ColPtrs cp{};
    for (int c = 0; c < ncols; ++c) cp.p[c] = cols[c];
    // choose and cross always run on the direct tables (measured faster on short and on long events alike); pairs and
    // distincts of dimuon-like events keep the event-parallel pair table / owner table + walk
    switch (kind) {
        case JG_COMB_PAIRS: launch_comb_fill<JG_COMB_PAIRS>(long_events, offsets_a, offsets_b, n, out_offsets, cp, k, absolute, s); break;
        case JG_COMB_DISTINCTS: launch_comb_fill<JG_COMB_DISTINCTS>(long_events, offsets_a, offsets_b, n, out_offsets, cp, k, absolute, s); break;
        case JG_COMB_CHOOSE: launch_comb_fill<JG_COMB_CHOOSE>(true, offsets_a, offsets_b, n, out_offsets, cp, k, absolute, s); break;
        case JG_COMB_CROSS: launch_comb_fill<JG_COMB_CROSS>(true, offsets_a, offsets_b, n, out_offsets, cp, k, absolute, s); break;
        default:
            set_error("jg_comb_fill: bad kind %d", kind);
            return -2;
    }
    return check_launch("comb_fill_kernel");
}

int jg_comb_gather2(int kind, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n, const int64_t *out_offsets,
                    const void *content_a, int itemsize_a, const void *content_b, int itemsize_b, void *out_l,
                    void *out_r, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const bool direct = (kind & JG_COMB_LONG) != 0;
    kind &= ~JG_COMB_LONG;
    switch (kind) {
        case JG_COMB_PAIRS:
            return gather2_a<JG_COMB_PAIRS>(direct, itemsize_a, itemsize_b, offsets_a, offsets_b, n, out_offsets, content_a, content_b, out_l, out_r, s);
        case JG_COMB_DISTINCTS:
            return gather2_a<JG_COMB_DISTINCTS>(direct, itemsize_a, itemsize_b, offsets_a, offsets_b, n, out_offsets, content_a, content_b, out_l, out_r, s);
        case JG_COMB_CROSS:
            return gather2_a<JG_COMB_CROSS>(direct, itemsize_a, itemsize_b, offsets_a, offsets_b, n, out_offsets, content_a, content_b, out_l, out_r, s);
    }
    set_error("jg_comb_gather2: bad kind %d", kind);
    return -2;
}

int jg_comb_gather_multi(int kind, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n,
                         const int64_t *out_offsets, int itemsize, int n_left, const void *const *in_left,
                         void *const *out_left, int n_right, const void *const *in_right, void *const *out_right,
                         void *stream) {
    if (n <= 0) return 0;
    JG_REQUIRE(n_left >= 0 && n_left <= 4 && n_right >= 0 && n_right <= 4 && n_left + n_right > 0,
               "jg_comb_gather_multi: 0..4 columns per side");
    MultiCols mc{};
    mc.nl = n_left;
    mc.nr = n_right;
    for (int c = 0; c < n_left; ++c) {
        mc.in_l[c] = in_left[c];
        mc.out_l[c] = out_left[c];
    }
    for (int c = 0; c < n_right; ++c) {
        mc.in_r[c] = in_right[c];
        mc.out_r[c] = out_right[c];
    }
    cudaStream_t s = as_stream(stream);
    const bool direct = (kind & JG_COMB_LONG) != 0;
    kind &= ~JG_COMB_LONG;
    switch (kind) {
        case JG_COMB_PAIRS: return launch_gather_multi<JG_COMB_PAIRS>(direct, itemsize, offsets_a, offsets_b, n, out_offsets, mc, s);
        case JG_COMB_DISTINCTS: return launch_gather_multi<JG_COMB_DISTINCTS>(direct, itemsize, offsets_a, offsets_b, n, out_offsets, mc, s);
        case JG_COMB_CROSS: return launch_gather_multi<JG_COMB_CROSS>(direct, itemsize, offsets_a, offsets_b, n, out_offsets, mc, s);
    }
    set_error("jg_comb_gather_multi: bad kind %d", kind);
    return -2;
}

int jg_comb_eval(int kind, const int64_t *offsets_a, const int64_t *offsets_b, int64_t n, const int64_t *out_offsets,
                 int dtype, const JgEvalProgram *program, void *out, void *stream) {
    if (n <= 0) return 0;
    const bool direct = (kind & JG_COMB_LONG) != 0;
    kind &= ~JG_COMB_LONG;
    JG_REQUIRE(program != nullptr && out != nullptr && offsets_a != nullptr && out_offsets != nullptr,
               "jg_comb_eval: bad arguments");
    JG_REQUIRE(kind != JG_COMB_CROSS || offsets_b != nullptr, "jg_comb_eval: cross needs offsets_b");
    const JgEvalProgram &pr = *program;
    JG_REQUIRE(pr.nins > 0 && pr.nins <= JG_EVAL_MAX_INS, "jg_comb_eval: program length %d out of range", pr.nins);
    JG_REQUIRE(pr.nslots >= 0 && pr.nslots <= JG_EVAL_MAX_SLOTS, "jg_comb_eval: %d spill slots out of range", pr.nslots);
    for (int i = 0; i < pr.nins; ++i) {
        const uint32_t w = pr.ins[i];
        const int op = w & 255, kind = (w >> 8) & 15, idx = w >> 16;
        const int cls = eval_op_class(op);
        JG_REQUIRE(cls != 0, "jg_comb_eval: step %d: operator %d cannot be fused", i, op);
        if (cls == 3) JG_REQUIRE(idx < pr.nslots, "jg_comb_eval: step %d stores to slot %d of %d", i, idx, pr.nslots);
        if (cls != 1) continue;
        switch (kind) {
            case JG_EV_SLOT: JG_REQUIRE(idx < pr.nslots, "jg_comb_eval: step %d reads slot %d of %d", i, idx, pr.nslots); break;
            case JG_EV_LEFT: JG_REQUIRE(idx < JG_EVAL_MAX_COLS && pr.left[idx], "jg_comb_eval: step %d: bad left column %d", i, idx); break;
            case JG_EV_RIGHT: JG_REQUIRE(idx < JG_EVAL_MAX_COLS && pr.right[idx], "jg_comb_eval: step %d: bad right column %d", i, idx); break;
            case JG_EV_POS: JG_REQUIRE(idx < JG_EVAL_MAX_COLS && pr.pos[idx], "jg_comb_eval: step %d: bad positional column %d", i, idx); break;
            case JG_EV_CONST: JG_REQUIRE(idx < JG_EVAL_MAX_CONSTS, "jg_comb_eval: step %d: bad constant %d", i, idx); break;
            default: JG_REQUIRE(false, "jg_comb_eval: step %d: bad operand kind %d", i, kind);
        }
    }
    JG_REQUIRE(eval_op_class(pr.ins[0] & 255) == 1 && (pr.ins[0] & 255) == JG_EV_LOAD, "jg_comb_eval: a program starts with a load");
    cudaStream_t s = as_stream(stream);
    switch (kind) {
        case JG_COMB_PAIRS: return launch_comb_eval_dtype<JG_COMB_PAIRS>(direct, dtype, offsets_a, offsets_b, n, out_offsets, pr, out, s);
        case JG_COMB_DISTINCTS: return launch_comb_eval_dtype<JG_COMB_DISTINCTS>(direct, dtype, offsets_a, offsets_b, n, out_offsets, pr, out, s);
        case JG_COMB_CROSS: return launch_comb_eval_dtype<JG_COMB_CROSS>(direct, dtype, offsets_a, offsets_b, n, out_offsets, pr, out, s);
    }
    set_error("jg_comb_eval: bad kind %d", kind);
    return -2;
}

}  // extern "C"

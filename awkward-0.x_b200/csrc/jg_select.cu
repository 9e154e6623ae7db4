// jg_select.cu -- jagged __getitem__: boolean-mask compaction (jagged.py:562-589), integer gather
// (jagged.py:541-560), plain take (jagged.py:1294) and the dense form of argmin/argmax results.
#include "jg_scan.cuh"
#include "jg_tile.cuh"

namespace jg {

// ------------------------------------------------------------------------------------------------------------
// a[jagged_bool] in ONE pass: per-event popcount -> block scan -> decoupled look-back across tiles ->
// new offsets + compaction.  The tile's content run and mask run are staged with TMA bulk copies.
// Algorithmic traffic: v*M + M + 8(N+1) read, v*K + 8(N+1) written.
// ------------------------------------------------------------------------------------------------------------
template <typename V, int STAGE_BYTES>
__global__ void __launch_bounds__(kTileThreads)
mask_select_kernel(const V *__restrict__ content, const uint8_t *__restrict__ mask, int64_t mask_shift,
                   const int64_t *__restrict__ offsets, int64_t n, V *__restrict__ out,
                   int64_t *__restrict__ new_offsets, int64_t *__restrict__ total, LookbackWs *__restrict__ ws,
                   uint32_t ntiles) {
    constexpr int kMaskBytes = STAGE_BYTES / sizeof(V);
    __shared__ int64_t s_off[kTileEvents + 1];
    __shared__ __align__(16) unsigned char s_content[STAGE_BYTES + 32];
    __shared__ __align__(16) unsigned char s_mask[kMaskBytes + 32];
    __shared__ uint64_t s_bar[2];
    __shared__ int32_t s_scan[kTileThreads / kWarp + 1];
    __shared__ uint32_t s_tile;
    __shared__ int64_t s_base;

    if (threadIdx.x == 0) s_tile = atomicAdd(&ws->ticket, 1u);
    stage_init(s_bar, 2);
    const uint32_t tile_index = s_tile;
    EventTile tile;
    tile.load(s_off, offsets, n, tile_index);
    const int64_t e0 = tile.first_element(), e1 = tile.last_element();
    const int64_t len = e1 - e0;
    const bool staged = len > 0 && len <= kMaskBytes;

    const V *cv = content + e0;
    const uint8_t *mv = mask + e0 + mask_shift;
    if (staged) {
        const uint32_t cs = stage_issue<sizeof(V)>(s_content, reinterpret_cast<const unsigned char *>(cv),
                                                   static_cast<uint32_t>(len * sizeof(V)), &s_bar[0]);
        const uint32_t ms = stage_issue<1>(s_mask, mv, static_cast<uint32_t>(len), &s_bar[1]);
        mbar_wait(&s_bar[0], 0);
        mbar_wait(&s_bar[1], 0);
        __syncthreads();
        cv = reinterpret_cast<const V *>(s_content + cs);
        mv = s_mask + ms;
    }

    // thread t owns events 2t and 2t+1 (kTileEvents == 2 * kTileThreads)
    const int ev0 = threadIdx.x * 2;
    int64_t a[2], b[2];
    int32_t c[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ev = ev0 + q;
        a[q] = b[q] = 0;
        if (ev < tile.nev) {
            a[q] = s_off[ev] - e0;
            b[q] = s_off[ev + 1] - e0;
            int32_t cnt = 0;
            for (int64_t j = a[q]; j < b[q]; ++j) cnt += (mv[j] != 0);
            c[q] = cnt;
        }
    }
    int32_t block_total;
    const int32_t excl = block_excl_scan<int32_t, kTileThreads>(c[0] + c[1], s_scan, block_total);
    if (threadIdx.x < kWarp) {
        const int64_t base = lookback_exclusive_warp(ws->desc, tile_index, block_total);
        if (threadIdx.x == 0) {
            s_base = base;
            if (tile_index == ntiles - 1) {
                new_offsets[n] = base + block_total;
                if (total) *total = base + block_total;
            }
        }
    }
    __syncthreads();
    const int64_t base = s_base;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ev = ev0 + q;
        if (ev < tile.nev) {
            int64_t w = base + excl + (q ? c[0] : 0);
            new_offsets[tile.event0 + ev] = w;
            for (int64_t j = a[q]; j < b[q]; ++j) {
                if (mv[j] != 0) out[w++] = cv[j];
            }
        }
    }
}

// counts of two layouts must agree (jagged.py:911-912)
__global__ void __launch_bounds__(256)
same_counts_kernel(const int64_t *__restrict__ oa, const int64_t *__restrict__ ob, int64_t n,
                   int32_t *__restrict__ status) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    bool bad = false;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        bad |= (oa[i + 1] - oa[i]) != (ob[i + 1] - ob[i]);
    if (bad) atomicOr(status, JG_ST_COUNTS_MISMATCH);
}

// ------------------------------------------------------------------------------------------------------------
// a[jagged_int]   (jagged.py:541-560)
// ------------------------------------------------------------------------------------------------------------
template <typename V, typename I>
__global__ void __launch_bounds__(kTileThreads)
index_gather_kernel(const V *__restrict__ content, const int64_t *__restrict__ offsets, const I *__restrict__ index,
                    const int64_t *__restrict__ index_offsets, int64_t n, V *__restrict__ out,
                    int32_t *__restrict__ status) {
    __shared__ int64_t s_ioff[kTileEvents + 1];
    __shared__ int64_t s_off[kTileEvents + 1];
    EventTile itile, tile;
    itile.load(s_ioff, index_offsets, n, blockIdx.x);
    tile.load(s_off, offsets, n, blockIdx.x);
    const int64_t io0 = __ldg(index_offsets);
    const int64_t k0 = itile.first_element(), k1 = itile.last_element();
    bool oob = false;
    // element-parallel over the index run of this tile: coalesced reads of index, coalesced writes of out
    for (int64_t k = k0 + threadIdx.x; k < k1; k += kTileThreads) {
        const int ev = itile.find_event(k);
        const int64_t start = s_off[ev], count = s_off[ev + 1] - start;
        int64_t ix = static_cast<int64_t>(index[k]);
        if (ix < 0) ix += count;                                    // jagged.py:552-553
        if (ix < 0 || ix >= count) {                                // jagged.py:555-556
            oob = true;
            continue;
        }
        out[k - io0] = __ldg(content + start + ix);
    }
    if (oob) atomicOr(status, JG_ST_INDEX_OOB);
}

template <typename V>
__global__ void __launch_bounds__(256)
take_kernel(const V *__restrict__ content, const int64_t *__restrict__ index, int64_t n, V *__restrict__ out) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __ldg(content + __ldcs(index + i));
}

// out_index[out_offsets[i]] = best[i] for best[i] >= 0
__global__ void __launch_bounds__(256)
scatter_nonneg_kernel(const int64_t *__restrict__ best, int64_t n, const int64_t *__restrict__ out_offsets,
                      int64_t *__restrict__ out_index) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t b = best[i];
        if (b >= 0) out_index[out_offsets[i]] = b;
    }
}

static inline int flat_grid(int64_t work, int threads) {
    int64_t blocks = ceil_div(work, threads);
    int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

template <typename V, int STAGE>
static int launch_mask_select(const void *content, const uint8_t *mask, int64_t mask_shift, const int64_t *offsets,
                              int64_t n, void *out, int64_t *new_offsets, int64_t *total, void *ws, size_t ws_bytes,
                              cudaStream_t s) {
    const int64_t ntiles = tile_grid(n);
    const size_t need = sizeof(LookbackWs) + static_cast<size_t>(ntiles) * 8;
    JG_REQUIRE(ws != nullptr && ws_bytes >= need, "jg_mask_select: workspace too small (%zu < %zu)", ws_bytes, need);
    cudaError_t e = cudaMemsetAsync(ws, 0, need, s);
    if (e != cudaSuccess) {
        set_error("jg_mask_select: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return -1;
    }
    mask_select_kernel<V, STAGE><<<static_cast<unsigned>(ntiles), kTileThreads, 0, s>>>(
        static_cast<const V *>(content), mask, mask_shift, offsets, n, static_cast<V *>(out), new_offsets, total,
        reinterpret_cast<LookbackWs *>(ws), static_cast<uint32_t>(ntiles));
    return check_launch("mask_select_kernel");
}

template <typename V>
static int launch_index_gather(const void *content, const int64_t *offsets, const void *index, int index_dtype,
                               const int64_t *index_offsets, int64_t n, void *out, int32_t *status, cudaStream_t s) {
    if (index_dtype == JG_I64)
        index_gather_kernel<V, int64_t><<<tile_grid(n), kTileThreads, 0, s>>>(
            static_cast<const V *>(content), offsets, static_cast<const int64_t *>(index), index_offsets, n,
            static_cast<V *>(out), status);
    else if (index_dtype == JG_I32)
        index_gather_kernel<V, int32_t><<<tile_grid(n), kTileThreads, 0, s>>>(
            static_cast<const V *>(content), offsets, static_cast<const int32_t *>(index), index_offsets, n,
            static_cast<V *>(out), status);
    else {
        set_error("jg_index_gather: unsupported index dtype %d", index_dtype);
        return -2;
    }
    return check_launch("index_gather_kernel");
}

}  // namespace jg

using namespace jg;

extern "C" {

int jg_mask_select(const void *content, int itemsize, const uint8_t *mask, int64_t mask_shift,
                   const int64_t *offsets, int64_t n, void *out, int64_t *new_offsets, int64_t *total, void *ws,
                   size_t ws_bytes, void *stream) {
    JG_REQUIRE(n >= 0 && offsets && new_offsets, "jg_mask_select: bad arguments");
    cudaStream_t s = as_stream(stream);
    switch (itemsize) {
        case 1: return launch_mask_select<uint8_t, 8192>(content, mask, mask_shift, offsets, n, out, new_offsets, total, ws, ws_bytes, s);
        case 2: return launch_mask_select<uint16_t, 8192>(content, mask, mask_shift, offsets, n, out, new_offsets, total, ws, ws_bytes, s);
        case 4: return launch_mask_select<uint32_t, 16384>(content, mask, mask_shift, offsets, n, out, new_offsets, total, ws, ws_bytes, s);
        case 8: return launch_mask_select<uint64_t, 32768>(content, mask, mask_shift, offsets, n, out, new_offsets, total, ws, ws_bytes, s);
    }
    set_error("jg_mask_select: unsupported itemsize %d", itemsize);
    return -2;
}

int jg_check_same_counts(const int64_t *offsets_a, const int64_t *offsets_b, int64_t n, int32_t *status,
                         void *stream) {
    if (n <= 0) return 0;
    same_counts_kernel<<<flat_grid(n, 256), 256, 0, as_stream(stream)>>>(offsets_a, offsets_b, n, status);
    return check_launch("same_counts_kernel");
}

int jg_index_gather(const void *content, int itemsize, const int64_t *offsets, const void *index, int index_dtype,
                    const int64_t *index_offsets, int64_t n, void *out, int32_t *status, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    switch (itemsize) {
        case 1: return launch_index_gather<uint8_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
        case 2: return launch_index_gather<uint16_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
        case 4: return launch_index_gather<uint32_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
        case 8: return launch_index_gather<uint64_t>(content, offsets, index, index_dtype, index_offsets, n, out, status, s);
    }
    set_error("jg_index_gather: unsupported itemsize %d", itemsize);
    return -2;
}

int jg_take(const void *content, int itemsize, const int64_t *index, int64_t n_index, void *out, void *stream) {
    if (n_index <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const int g = flat_grid(n_index, 256);
    switch (itemsize) {
        case 1: take_kernel<uint8_t><<<g, 256, 0, s>>>(static_cast<const uint8_t *>(content), index, n_index, static_cast<uint8_t *>(out)); break;
        case 2: take_kernel<uint16_t><<<g, 256, 0, s>>>(static_cast<const uint16_t *>(content), index, n_index, static_cast<uint16_t *>(out)); break;
        case 4: take_kernel<uint32_t><<<g, 256, 0, s>>>(static_cast<const uint32_t *>(content), index, n_index, static_cast<uint32_t *>(out)); break;
        case 8: take_kernel<uint64_t><<<g, 256, 0, s>>>(static_cast<const uint64_t *>(content), index, n_index, static_cast<uint64_t *>(out)); break;
        default:
            set_error("jg_take: unsupported itemsize %d", itemsize);
            return -2;
    }
    return check_launch("take_kernel");
}

int jg_compact_nonneg(const int64_t *best, int64_t n, int64_t *out_offsets, int64_t *out_index, int64_t *total,
                      void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t s = as_stream(stream);
    ScanInNonNeg in{best};
    int rc = launch_scan(in, n, out_offsets, total, nullptr, ws, ws_bytes, s);
    if (rc) return rc;
    if (n > 0) {
        scatter_nonneg_kernel<<<flat_grid(n, 256), 256, 0, s>>>(best, n, out_offsets, out_index);
        return check_launch("scatter_nonneg_kernel");
    }
    return 0;
}

}  // extern "C"

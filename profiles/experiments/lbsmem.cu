// experiment: look-back scan with the tile parked in shared memory (TMA bulk load), low register count, high residency
#include <cstdio>
#include <cstdlib>
#include "jg_scan.cuh"
#include "jg_tile.cuh"
namespace jg { void set_error(const char*, ...) {} int check_launch(const char*) { return cudaGetLastError() != cudaSuccess; } int sm_count() { return 148; } }
using namespace jg;

template <int THREADS, int ITEMS, int MINB>   // ITEMS per thread (multiple of 2)
__global__ void __launch_bounds__(THREADS, MINB) scan_smem_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out, LookbackWs* ws) {
    constexpr int NW = THREADS / 32; constexpr int TILE = THREADS * ITEMS;
    __shared__ __align__(16) int64_t s_data[TILE];
    __shared__ uint64_t s_bar;
    __shared__ int64_t s_warp[NW + 1];
    __shared__ int64_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    if (tid == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
    __syncthreads();
    const int64_t i0 = (int64_t)tile * TILE;
    if (tid == 0) { mbar_expect_tx(&s_bar, TILE * 8); bulk_g2s(s_data, in + i0, TILE * 8, &s_bar); }
    mbar_wait(&s_bar, 0);
    // warp w owns ITEMS*32 consecutive items; rows of 64 (lane: 2 items)
    constexpr int ROWS = ITEMS / 2;
    const int wbase = warp * (ITEMS * 32);
    int64_t sum = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { longlong2 q = *reinterpret_cast<const longlong2*>(&s_data[wbase + r * 64 + lane * 2]); sum += q.x + q.y; }
    sum = warp_reduce_add(sum);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < NW ? s_warp[lane] : 0;
        int64_t wi = warp_incl_scan_add(w, lane);
        int64_t tot = __shfl_sync(0xffffffffu, wi, NW - 1);
        int64_t excl = lookback_exclusive_warp<1>(ws->desc, tile, tot);
        if (lane < NW) s_warp[lane] = wi - w;
        if (lane == 0) s_base = excl;
    }
    __syncthreads();
    int64_t run = s_base + s_warp[warp];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        longlong2 q = *reinterpret_cast<const longlong2*>(&s_data[wbase + r * 64 + lane * 2]);
        int64_t s2 = q.x + q.y;
        int64_t incl = warp_incl_scan_add(s2, lane);
        int64_t e0 = run + incl - s2;
        __stcs(reinterpret_cast<longlong2*>(out + i0 + wbase + r * 64 + lane * 2), make_longlong2(e0, e0 + q.x));
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

template <int THREADS, int ITEMS, int MINB>
void run(const char* name, int64_t* in, int64_t* out, void* ws) {
    constexpr int TILE = THREADS * ITEMS;
    int64_t n = 125000000LL / 8192 * 8192;
    int64_t ntiles = n / TILE;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_smem_kernel<THREADS, ITEMS, MINB>, THREADS, 0);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(ws, 0, 16 + ntiles * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        scan_smem_kernel<THREADS, ITEMS, MINB><<<ntiles, THREADS>>>(in, n, out, (LookbackWs*)ws);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // verify
    int64_t last; cudaMemcpy(&last, out + n - 1, 8, cudaMemcpyDeviceToHost);
    printf("%-16s occ %2d tile %5d: best %.3f ms (%.0f GB/s) last=%lld (expect %lld) err=%d\n", name, occ, TILE, best, n * 16 / best / 1e6, (long long)last, (long long)(n - 1), (int)cudaGetLastError());
}

__global__ void fill1(int64_t* p, int64_t n) { int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (i < n) p[i] = 1; }

int main() {
    int64_t n = 125000000LL / 8192 * 8192;
    int64_t *in, *out; void* ws;
    cudaMalloc(&in, n * 8); cudaMalloc(&out, n * 8); cudaMalloc(&ws, 16 + (n / 512 + 64) * 8);
    fill1<<<(n + 255) / 256, 256>>>(in, n);
    run<256, 16, 1>("256x16", in, out, ws);
    run<256, 8, 1>("256x8", in, out, ws);
    run<128, 16, 1>("128x16", in, out, ws);
    run<128, 8, 1>("128x8", in, out, ws);
    run<512, 8, 1>("512x8", in, out, ws);
    run<256, 4, 1>("256x4", in, out, ws);
    run<128, 32, 1>("128x32", in, out, ws);
    return 0;
}

// trace dispatch / ready / done times per tile for the look-back scan (experiment only)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "jg_scan.cuh"
namespace jg { void set_error(const char*, ...) {} int check_launch(const char*) { return cudaGetLastError() != cudaSuccess; } int sm_count() { return 148; } }
using namespace jg;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }

template <int THREADS, int ROWS>
__global__ void __launch_bounds__(THREADS) lb_kernel(const int64_t* in, int64_t n, int64_t* out, LookbackWs* ws, unsigned long long* trace) {
    constexpr int NW = THREADS / 32; constexpr int TILE = THREADS * ROWS * 2;
    __shared__ int64_t s_warp[NW + 1];
    __shared__ int64_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    unsigned long long t_disp = gtime();
    int64_t v[ROWS][2];
    const int64_t w0 = (int64_t)tile * TILE + warp * (ROWS * 64);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { longlong2 q = __ldcs((const longlong2*)(in + w0 + r * 64 + lane * 2)); v[r][0] = q.x; v[r][1] = q.y; }
    int64_t sum = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) sum += v[r][0] + v[r][1];
    sum = warp_reduce_add(sum);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < NW ? s_warp[lane] : 0;
        int64_t tot = warp_reduce_add(w);
        unsigned long long t_ready = gtime();
        int64_t excl = lookback_exclusive_warp<1>(ws->desc, tile, tot);
        unsigned long long t_done = gtime();
        if (lane == 0) { s_base = excl; trace[4 * tile] = t_disp; trace[4 * tile + 1] = t_ready; trace[4 * tile + 2] = t_done; trace[4 * tile + 3] = smid(); }
    }
    __syncthreads();
    int64_t base = s_base;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { __stcs((longlong2*)(out + w0 + r * 64 + lane * 2), make_longlong2(base + v[r][0], base + v[r][1])); }
}

int main() {
    constexpr int THREADS = 256, ROWS = 8, TILE = THREADS * ROWS * 2;
    int64_t n = 125000000LL / 4096 * 4096;
    int64_t ntiles = n / TILE;
    int64_t *in, *out; void* ws; unsigned long long* trace;
    cudaMalloc(&in, n * 8); cudaMalloc(&out, n * 8); cudaMalloc(&ws, 16 + ntiles * 8); cudaMalloc(&trace, ntiles * 32);
    cudaMemset(in, 0, n * 8);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(ws, 0, 16 + ntiles * 8);
        lb_kernel<THREADS, ROWS><<<ntiles, THREADS>>>(in, n, out, (LookbackWs*)ws, trace);
        cudaDeviceSynchronize();
    }
    std::vector<unsigned long long> h(ntiles * 4);
    cudaMemcpy(h.data(), trace, ntiles * 32, cudaMemcpyDeviceToHost);
    unsigned long long t0 = h[0];
    for (int64_t t = 0; t < ntiles; ++t) t0 = std::min(t0, h[4 * t]);
    // print a window in the middle
    printf("tile disp ready done sm (ns rel)\n");
    for (int64_t t = 15000; t < 15120; ++t)
        printf("%lld %llu %llu %llu %llu\n", (long long)t, h[4 * t] - t0, h[4 * t + 1] - t0, h[4 * t + 2] - t0, h[4 * t + 3]);
    // stats
    double load = 0, wait = 0; long long inv = 0; double late = 0;
    for (int64_t t = 1000; t < ntiles; ++t) {
        load += (double)(h[4 * t + 1] - h[4 * t]); wait += (double)(h[4 * t + 2] - h[4 * t + 1]);
        if (h[4 * t] < h[4 * (t - 1)]) inv++;
        // how late is the latest of the previous 32 tiles' ready vs my ready
        unsigned long long mx = 0; for (int k = 1; k <= 32; ++k) mx = std::max(mx, h[4 * (t - k) + 1]);
        if (mx > h[4 * t + 1]) late += (double)(mx - h[4 * t + 1]);
    }
    double m = ntiles - 1000;
    printf("avg load %.0f ns, avg lookback %.0f ns, dispatch inversions %lld, avg lateness of slowest of prev 32: %.0f ns, total span %.0f us\n", load / m, wait / m, inv, late / m, 0.0);
    return 0;
}

// stand-alone experiment: where does decoupled look-back time go?  (not part of the library)
#include <cstdio>
#include <cstdlib>
#include "jg_scan.cuh"
namespace jg { void set_error(const char*, ...) {} int check_launch(const char*) { return cudaGetLastError() != cudaSuccess; } int sm_count() { return 148; } }
using namespace jg;


// two-level look-back: tile descriptors + one descriptor per group of 32 tiles
__device__ __forceinline__ int64_t lookback2(uint64_t* desc, uint64_t* gdesc, uint32_t tile, int64_t aggregate, unsigned& rounds, unsigned& spins) {
    const int lane = threadIdx.x & 31;
    const uint64_t agg = (uint64_t)aggregate & kValueMask;
    const uint32_t g = tile >> 5, r = tile & 31;
    if (lane == 0) st_relaxed_u64(&desc[tile], ((tile == 0 ? kStatePrefix : kStateAggregate) << 62) | agg);
    uint64_t excl = 0;
    bool resolved = (tile == 0);
    // level 1: my own group, tiles g*32 .. tile-1 (lane l looks at tile-1-l)
    if (r > 0) {
        while (true) {
            const bool mine = lane < (int)r;
            uint64_t w = mine ? ld_relaxed_u64(&desc[tile - 1 - lane]) : 0ull;
            uint64_t st = mine ? (w >> 62) : kStateAggregate;
            unsigned invalid = __ballot_sync(0xffffffffu, st == kStateInvalid);
            unsigned prefix = __ballot_sync(0xffffffffu, mine && st == kStatePrefix);
            int fp = prefix ? (__ffs(prefix) - 1) : 32;
            unsigned need = fp >= 31 ? 0xffffffffu : ((2u << fp) - 1u);
            ++rounds;
            if (invalid & need) { ++spins; continue; }
            uint64_t c = (mine && lane <= fp) ? (w & kValueMask) : 0ull;
            excl += warp_reduce_add(c);
            if (prefix) resolved = true;
            break;
        }
    }
    if (r == 31 && lane == 0 && !resolved) st_relaxed_u64(&gdesc[g], (kStateAggregate << 62) | ((excl + agg) & kValueMask));
    // level 2: previous groups (lane l looks at group g-1-l)
    if (!resolved) {
        if (g == 0) resolved = true;
        int64_t base = (int64_t)g - 1;
        while (!resolved) {
            int64_t q = base - lane;
            uint64_t w = (q >= 0) ? ld_relaxed_u64(&gdesc[q]) : (kStatePrefix << 62);
            uint64_t st = w >> 62;
            unsigned invalid = __ballot_sync(0xffffffffu, st == kStateInvalid);
            unsigned prefix = __ballot_sync(0xffffffffu, st == kStatePrefix);
            int fp = prefix ? (__ffs(prefix) - 1) : 32;
            unsigned need = fp >= 31 ? 0xffffffffu : ((2u << fp) - 1u);
            ++rounds;
            if (invalid & need) { ++spins; continue; }
            uint64_t c = (lane <= fp) ? (w & kValueMask) : 0ull;
            excl += warp_reduce_add(c);
            if (prefix) break;
            base -= 32;
        }
    }
    if (lane == 0) {
        st_relaxed_u64(&desc[tile], (kStatePrefix << 62) | ((excl + agg) & kValueMask));
        if (r == 31) st_relaxed_u64(&gdesc[g], (kStatePrefix << 62) | ((excl + agg) & kValueMask));
    }
    return (int64_t)excl;
}

__device__ unsigned long long g_rounds, g_spins, g_clk_lb, g_clk_total, g_maxwalk;

template <int MODE, int THREADS, int ROWS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) lb_kernel(const int64_t* in, int64_t n, int64_t* out, LookbackWs* ws, int work) {
    constexpr int NW = THREADS / 32; constexpr int TILE = THREADS * ROWS * 2;
    __shared__ int64_t s_warp[NW + 1];
    __shared__ int64_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    long long t0 = clock64();
    // load 16 items per thread (blocked via striped rows)
    int64_t v[ROWS][2];
    const int64_t w0 = (int64_t)tile * TILE + warp * (ROWS * 64);
    _Pragma("unroll") for (int r = 0; r < ROWS; ++r) { longlong2 q = __ldcs((const longlong2*)(in + w0 + r * 64 + lane * 2)); v[r][0] = q.x; v[r][1] = q.y; }
    int64_t sum = 0;
    _Pragma("unroll") for (int r = 0; r < ROWS; ++r) sum += v[r][0] + v[r][1];
    sum = warp_reduce_add(sum);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < NW ? s_warp[lane] : 0;
        int64_t tot = warp_reduce_add(w);
        long long a = clock64();
        // instrumented lookback
        const uint64_t agg = (uint64_t)tot & kValueMask;
        uint64_t excl = 0; unsigned rounds = 0, spins = 0;
        if (MODE == 2) { excl = (uint64_t)lookback2(ws->desc, ws->desc + gridDim.x + 32, tile, tot, rounds, spins); }
        else if (tile == 0) { if (lane == 0) st_relaxed_u64(&ws->desc[0], (kStatePrefix << 62) | agg); }
        else {
            if (lane == 0) st_relaxed_u64(&ws->desc[tile], (kStateAggregate << 62) | agg);
            int64_t base = (int64_t)tile - 1;
            while (true) {
                int64_t t = base - lane;
                uint64_t wv = (t >= 0) ? ld_relaxed_u64(&ws->desc[t]) : (kStatePrefix << 62);
                uint64_t st = wv >> 62;
                unsigned invalid = __ballot_sync(0xffffffffu, st == kStateInvalid);
                unsigned prefix = __ballot_sync(0xffffffffu, st == kStatePrefix);
                int fp = prefix ? (__ffs(prefix) - 1) : 32;
                unsigned need = fp >= 31 ? 0xffffffffu : ((2u << fp) - 1u);
                ++rounds;
                if (invalid & need) { ++spins; if (MODE == 1) __nanosleep(100); continue; }
                uint64_t c = (lane <= fp) ? (wv & kValueMask) : 0ull;
                excl += warp_reduce_add(c);
                if (prefix) break;
                base -= 32;
            }
            if (lane == 0) st_relaxed_u64(&ws->desc[tile], (kStatePrefix << 62) | ((excl + agg) & kValueMask));
        }
        long long b = clock64();
        if (lane == 0) { s_base = (int64_t)excl; atomicAdd(&g_rounds, rounds); atomicAdd(&g_spins, spins); atomicAdd(&g_clk_lb, (unsigned long long)(b - a)); atomicMax(&g_maxwalk, (unsigned long long)rounds); }
    }
    __syncthreads();
    int64_t base = s_base;
    _Pragma("unroll") for (int r = 0; r < ROWS; ++r) { __stcs((longlong2*)(out + w0 + r * 64 + lane * 2), make_longlong2(base + v[r][0], base + v[r][1])); }
    long long t1 = clock64();
    if (tid == 0) atomicAdd(&g_clk_total, (unsigned long long)(t1 - t0));
}


template <int MODE, int THREADS, int ROWS, int MINB>
void run(const char* name, int64_t* in, int64_t* out, void* ws) {
    constexpr int TILE = THREADS * ROWS * 2;
    int64_t n = 125000000LL / 4096 * 4096;
    int64_t ntiles = n / TILE;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lb_kernel<MODE, THREADS, ROWS, MINB>, THREADS, 0);
    float best = 1e9; unsigned long long r, s, c, t, m;
    for (int rep = 0; rep < 4; ++rep) {
        unsigned long long z = 0;
        cudaMemcpyToSymbol(g_rounds, &z, 8); cudaMemcpyToSymbol(g_spins, &z, 8); cudaMemcpyToSymbol(g_clk_lb, &z, 8); cudaMemcpyToSymbol(g_clk_total, &z, 8); cudaMemcpyToSymbol(g_maxwalk, &z, 8);
        cudaMemset(ws, 0, 16 + (ntiles * 2 + 64) * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        lb_kernel<MODE, THREADS, ROWS, MINB><<<ntiles, THREADS>>>(in, n, out, (LookbackWs*)ws, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        cudaMemcpyFromSymbol(&r, g_rounds, 8); cudaMemcpyFromSymbol(&s, g_spins, 8); cudaMemcpyFromSymbol(&c, g_clk_lb, 8); cudaMemcpyFromSymbol(&t, g_clk_total, 8); cudaMemcpyFromSymbol(&m, g_maxwalk, 8);
    }
    printf("%-22s occ %2d: best %.3f ms (%.0f GB/s) tiles %lld rounds/tile %.2f spins/tile %.2f lb clk/tile %.0f total clk/tile %.0f err=%d\n", name, occ, best, n * 16 / best / 1e6, (long long)ntiles,
           (double)r / ntiles, (double)s / ntiles, (double)c / ntiles, (double)t / ntiles, (int)cudaGetLastError());
}

int main(int argc, char** argv) {
    int64_t n = 125000000LL / 4096 * 4096;
    int64_t *in, *out; void* ws;
    cudaMalloc(&in, n * 8); cudaMalloc(&out, n * 8); cudaMalloc(&ws, 16 + (n / 256 * 2 + 64) * 8);
    cudaMemset(in, 0, n * 8);
    run<0, 256, 8, 1>("256x8", in, out, ws);
    run<2, 256, 8, 1>("256x8 2level", in, out, ws);
    run<0, 256, 4, 6>("256x4 minb6", in, out, ws);
    run<2, 256, 4, 6>("256x4 minb6 2level", in, out, ws);
    run<0, 128, 8, 8>("128x8 minb8", in, out, ws);
    run<2, 128, 8, 8>("128x8 minb8 2level", in, out, ws);
    run<2, 128, 4, 16>("128x4 minb16 2level", in, out, ws);
    run<2, 128, 2, 16>("128x2 minb16 2level", in, out, ws);
    run<2, 256, 2, 8>("256x2 minb8 2level", in, out, ws);
    return 0;
}

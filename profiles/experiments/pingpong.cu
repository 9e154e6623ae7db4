// experiment: flag visibility latency between SMs, idle vs under bulk streaming load
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t ld_relaxed(const uint64_t* p) { uint64_t v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed(uint64_t* p, uint64_t v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t ld_acquire(const uint64_t* p) { uint64_t v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(uint64_t* p, uint64_t v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// blocks 0 and 1 ping-pong; other blocks stream memory if `load` != 0
template <int MODE>
__global__ void pp(uint64_t* flags, int iters, const int4* src, int4* dst, int64_t nvec, int load, unsigned long long* result) {
    if (blockIdx.x < 2) {
        if (threadIdx.x != 0) return;
        const int me = blockIdx.x;
        uint64_t* mine = flags + 32 * me; uint64_t* other = flags + 32 * (1 - me);
        unsigned long long t0 = gtime();
        for (int i = 1; i <= iters; ++i) {
            if (me == 0) {
                if (MODE == 0) st_relaxed(mine, i); else if (MODE == 1) atomicExch((unsigned long long*)mine, (unsigned long long)i); else st_release(mine, i);
                while (true) { uint64_t v = (MODE == 2) ? ld_acquire(other) : (MODE == 1 ? atomicAdd((unsigned long long*)other, 0ull) : ld_relaxed(other)); if (v >= (uint64_t)i) break; }
            } else {
                while (true) { uint64_t v = (MODE == 2) ? ld_acquire(other) : (MODE == 1 ? atomicAdd((unsigned long long*)other, 0ull) : ld_relaxed(other)); if (v >= (uint64_t)i) break; }
                if (MODE == 0) st_relaxed(mine, i); else if (MODE == 1) atomicExch((unsigned long long*)mine, (unsigned long long)i); else st_release(mine, i);
            }
        }
        unsigned long long t1 = gtime();
        if (me == 0) *result = t1 - t0;
        return;
    }
    if (!load) return;
    // streaming copy load until block 0 is done (bounded)
    int64_t stride = (int64_t)(gridDim.x - 2) * blockDim.x;
    for (int rep = 0; rep < load; ++rep)
        for (int64_t i = (int64_t)(blockIdx.x - 2) * blockDim.x + threadIdx.x; i < nvec; i += stride) dst[i] = src[i];
}

int main() {
    uint64_t* flags; unsigned long long* res; int4 *src, *dst; int64_t nvec = (1LL << 30) / 16;
    cudaMalloc(&flags, 1024); cudaMalloc(&res, 8); cudaMalloc(&src, nvec * 16); cudaMalloc(&dst, nvec * 16);
    cudaMemset(src, 1, nvec * 16);
    const int iters = 200;
    for (int load = 0; load < 2; ++load) for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(flags, 0, 1024);
        int grid = load ? 2 + 148 * 6 : 2;
        if (mode == 0) pp<0><<<grid, 256>>>(flags, iters, src, dst, nvec, load ? 3 : 0, res);
        if (mode == 1) pp<1><<<grid, 256>>>(flags, iters, src, dst, nvec, load ? 3 : 0, res);
        if (mode == 2) pp<2><<<grid, 256>>>(flags, iters, src, dst, nvec, load ? 3 : 0, res);
        cudaDeviceSynchronize();
        unsigned long long r; cudaMemcpy(&r, res, 8, cudaMemcpyDeviceToHost);
        printf("load=%d mode=%d (0 relaxed st/ld, 1 atomics, 2 release/acquire): round trip %.0f ns  err=%d\n", load, mode, (double)r / iters, (int)cudaGetLastError());
    }
    return 0;
}

// Does a concurrent HBM read stream slow random-address RED?  (explains k_scatter's 124 G RED/s vs 210 G/s alone)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>  // 0: RED only; 1: RED + streaming 16B load per RED (ld.cs); 2: RED + default-cached load; 3: loads only
__global__ void __launch_bounds__(256) k(unsigned long long* acc, uint32_t ncell, const uint4* stream, size_t nvec, int iters, uint32_t* sink) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    uint32_t s = mix((uint32_t)tid * 2654435761U + 7U), a = 0;
    for (int i = 0; i < iters; ++i) {
        uint4 v = make_uint4(0, 0, 0, 0);
        const size_t idx = ((size_t)i * nthr + tid) % nvec;
        if (MODE == 1 || MODE == 3) v = __ldcs(stream + idx);
        if (MODE == 2) v = stream[idx];
        s = s * 1664525U + 1013904223U + v.x;
        a += v.y;
        if (MODE != 3) {
            const uint32_t c = (uint32_t)(((uint64_t)mix(s) * ncell) >> 32);
            atomicAdd(acc + c, (unsigned long long)(s >> 8));
        }
    }
    if (a == 0x12345678u) sink[tid] = a;
}
template <int MODE> float run(unsigned long long* acc, uint32_t ncell, const uint4* st, size_t nvec, uint32_t* sink, int grid, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE><<<grid, 256>>>(acc, ncell, st, nvec, iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    return best;
}
int main() {
    const size_t stream_bytes = 3ull << 30; uint4* st; cudaMalloc(&st, stream_bytes); cudaMemset(st, 1, stream_bytes);
    unsigned long long* acc; cudaMalloc(&acc, 256 << 20); cudaMemset(acc, 0, 256 << 20);
    uint32_t* sink; cudaMalloc(&sink, 64 << 20);
    const int grid = 148 * 16, iters = 256; const double ops = (double)grid * 256 * iters;
    for (uint32_t mb : {10u, 30u, 50u}) {
        const uint32_t ncell = mb * (1u << 20) / 8; const size_t nvec = stream_bytes / 16;
        float a = run<0>(acc, ncell, st, nvec, sink, grid, iters), b = run<1>(acc, ncell, st, nvec, sink, grid, iters);
        float c = run<2>(acc, ncell, st, nvec, sink, grid, iters), d = run<3>(acc, ncell, st, nvec, sink, grid, iters);
        printf("footprint %2u MB: RED only %6.1f G/s | RED + ld.cs 16B %6.1f G/s | RED + ld 16B %6.1f G/s | loads only %6.1f G/s (%.0f GB/s)\n",
               mb, ops / a * 1e-6, ops / b * 1e-6, ops / c * 1e-6, ops / d * 1e-6, ops * 16 / d * 1e-6);
    }
    return 0;
}

// Scatter-throughput microbenchmarks for sm_100a: decides the binning kernel design.
// Measures random-address RED (global, L2-resident footprints), ATOMS (shared) and
// plain LDS/STS read-modify-write rates.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <typename T, int PAIR>
__global__ void __launch_bounds__(256) k_red_global(T* __restrict__ buf, uint32_t nelem, uint32_t plane, int iters) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid * 2654435761U + 12345U);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525U + 1013904223U;
        uint32_t a = (uint32_t)(((uint64_t)mix(s) * nelem) >> 32);
        atomicAdd(buf + a, (T)(s >> 20));
        if (PAIR) atomicAdd(buf + a + plane, (T)3);
    }
}

// float2 vector red (sm_90+)
__global__ void __launch_bounds__(256) k_red_v2f32(float* __restrict__ buf, uint32_t nelem2, int iters) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid * 2654435761U + 12345U);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525U + 1013904223U;
        uint32_t a = (uint32_t)(((uint64_t)mix(s) * nelem2) >> 32);
        float* p = buf + 2ull * a;
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(1.0f), "f"(0.5f) : "memory");
    }
}

template <typename T, int MODE>  // MODE 0 = atomicAdd, 1 = plain RMW, 2 = store only
__global__ void __launch_bounds__(512) k_smem(T* __restrict__ out, uint32_t nelem, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    for (uint32_t i = threadIdx.x; i < nelem; i += blockDim.x) tile[i] = 0;
    __syncthreads();
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid * 2654435761U + 777U);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525U + 1013904223U;
        uint32_t a = (uint32_t)(((uint64_t)mix(s) * nelem) >> 32);
        if (MODE == 0) atomicAdd(tile + a, (T)(s >> 20));
        else if (MODE == 1) { T v = tile[a]; tile[a] = v + (T)1; }
        else tile[a] = (T)s;
    }
    __syncthreads();
    T acc = 0;
    for (uint32_t i = threadIdx.x; i < nelem; i += blockDim.x) acc += tile[i];
    if (acc == (T)0x7fffffff) out[tid] = acc;
}

// baseline: just the address generation, to subtract ALU cost
__global__ void __launch_bounds__(256) k_alu_only(uint32_t* out, uint32_t nelem, int iters) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid * 2654435761U + 12345U), acc = 0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525U + 1013904223U;
        acc += (uint32_t)(((uint64_t)mix(s) * nelem) >> 32);
    }
    if (acc == 0x12345678U) out[tid] = acc;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms=%d l2=%d MB persistingL2Max=%d MB smemOptin=%zu clock=%d kHz\n", prop.name, sms,
           prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.sharedMemPerBlockOptin, prop.clockRate);
    const size_t maxbytes = 512ull << 20;
    void* buf; CK(cudaMalloc(&buf, maxbytes + (64 << 20))); CK(cudaMemset(buf, 0, maxbytes + (64 << 20)));
    const int iters = 256;
    const int grid = sms * 16, block = 256;
    const double nops = (double)grid * block * iters;

    { float ms = time_ms([&] { k_alu_only<<<grid, block>>>((uint32_t*)buf, 1u << 20, iters); });
      printf("alu_only                         : %8.3f ms  %8.1f Gop/s\n", ms, nops / ms * 1e-6); }

    size_t foots[] = {1ull << 20, 5ull << 19, 10ull << 20, 40ull << 20, 96ull << 20, 256ull << 20};
    for (size_t f : foots) {
        uint32_t n32 = (uint32_t)(f / 4), n64 = (uint32_t)(f / 8);
        float a = time_ms([&] { k_red_global<uint32_t, 0><<<grid, block>>>((uint32_t*)buf, n32, 0, iters); });
        float b = time_ms([&] { k_red_global<unsigned long long, 0><<<grid, block>>>((unsigned long long*)buf, n64, 0, iters); });
        float c = time_ms([&] { k_red_global<uint32_t, 1><<<grid, block>>>((uint32_t*)buf, n32 / 2, n32 / 2, iters); });
        float d = time_ms([&] { k_red_v2f32<<<grid, block>>>((float*)buf, n64, iters); });
        float e = time_ms([&] { k_red_global<float, 0><<<grid, block>>>((float*)buf, n32, 0, iters); });
        printf("global footprint %6.1f MB: red.u32 %7.1f  red.u64 %7.1f  red.u32 pair(events/s) %7.1f  red.v2.f32 %7.1f  red.f32 %7.1f  Gop/s\n",
               f / 1048576.0, nops / a * 1e-6, nops / b * 1e-6, nops / c * 1e-6, nops / d * 1e-6, nops / e * 1e-6);
    }

    // shared-memory variants: one CTA per SM (big tile) and 2 CTAs per SM (half tiles)
    CK(cudaFuncSetAttribute(k_smem<uint32_t, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_smem<uint32_t, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_smem<uint32_t, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_smem<unsigned long long, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_smem<unsigned long long, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_smem<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int nthr : {256, 512}) {
        for (size_t tile : {16u * 1024, 96u * 1024, 200u * 1024}) {
            const int sgrid = sms * 4; const int sit = 2048;
            const double sops = (double)sgrid * nthr * sit;
            float a = time_ms([&] { k_smem<uint32_t, 0><<<sgrid, nthr, tile>>>((uint32_t*)buf, (uint32_t)(tile / 4), sit); });
            float b = time_ms([&] { k_smem<uint32_t, 1><<<sgrid, nthr, tile>>>((uint32_t*)buf, (uint32_t)(tile / 4), sit); });
            float c = time_ms([&] { k_smem<uint32_t, 2><<<sgrid, nthr, tile>>>((uint32_t*)buf, (uint32_t)(tile / 4), sit); });
            float d = time_ms([&] { k_smem<unsigned long long, 0><<<sgrid, nthr, tile>>>((unsigned long long*)buf, (uint32_t)(tile / 8), sit); });
            float e = time_ms([&] { k_smem<unsigned long long, 1><<<sgrid, nthr, tile>>>((unsigned long long*)buf, (uint32_t)(tile / 8), sit); });
            float f = time_ms([&] { k_smem<float, 0><<<sgrid, nthr, tile>>>((float*)buf, (uint32_t)(tile / 4), sit); });
            printf("smem tile %3zu KB thr %3d: atoms.u32 %7.1f  rmw.u32 %7.1f  sts.u32 %7.1f  atoms.u64 %7.1f  rmw.u64 %7.1f  atoms.f32 %7.1f  Gop/s\n",
                   tile >> 10, nthr, sops / a * 1e-6, sops / b * 1e-6, sops / c * 1e-6, sops / d * 1e-6, sops / e * 1e-6, sops / f * 1e-6);
        }
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}

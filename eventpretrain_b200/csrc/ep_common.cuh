// Shared helpers for the sm_100a kernels of the EventPretrain input hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/eventpretrain_b200.h"

namespace ep {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// Every kernel launch of the library goes through EP_LAUNCH_CHECK(): it counts the launch (bench.py's
// gpu_launches) and surfaces launch errors as the entry point's return value.
extern unsigned long long g_launch_count;

#define EP_LAUNCH_CHECK()                                   \
    do {                                                    \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) return (int)_e;              \
        __atomic_fetch_add(&ep::g_launch_count, 1ull, __ATOMIC_RELAXED); \
    } while (0)

// Optional per-kernel timing (ep_profile_*): CUDA events recorded on the launching stream around a launch.
enum ProfileKind { kProfScatter = 0, kProfFinalize = 1, kProfOther = 2, kProfKinds = 3 };
bool profile_enabled();
void profile_begin(cudaStream_t st, int kind);
void profile_end(cudaStream_t st);

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// csrc/ep_patch_tma.cu: TMA form of the (c, ph, pw) patch gather; EP_EUNSUPPORTED = shape does not qualify.
int patchify_gather_tma(cudaStream_t st, const float* x, const int64_t* ids_keep, int batch, int channels, int height, int width,
                        int patch, int K, float* out);
// Opt-in (EP_PATCH_TMA=1): measured 66 us vs 61 us for the plain float4 kernel at (512,5,224,224) p=16 keep=49 — 64-byte
// box rows keep the TMA below the LSU path here, so the plain kernel stays the default.
inline bool patch_tma_disabled() {
    static const int on = [] { const char* e = getenv("EP_PATCH_TMA"); return (e && e[0] == '1') ? 1 : 0; }();
    return on == 0;
}

// csrc/ep_binning_tiled.cu: EvRep for the 4 B packed transport layout (transposed route + shared-memory sweep);
// EP_EUNSUPPORTED = layout / shape does not qualify.  Workspace bytes 0 = does not qualify.
size_t evrep_packed4_workspace_bytes(const ep_events_soa* ev, int height, int width);
int run_evrep_packed4(cudaStream_t st, const ep_events_soa* ev, int height, int width, double* out, void* ws, size_t ws_bytes,
                      unsigned int* bad);

// ---- streaming loads / stores (events and finished outputs are touched exactly once) ------------
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }

__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float2* p, float2 v) { __stcs(p, v); }

// ---- warp / block reductions --------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// ---- scalar load of a tagged dtype (generic, non-vectorised event layouts) -----------------------
__device__ __forceinline__ double load_as_double(const void* base, int dtype, int64_t i) {
    switch (dtype) {
        case EP_U8: return (double)static_cast<const uint8_t*>(base)[i];
        case EP_I8: return (double)static_cast<const int8_t*>(base)[i];
        case EP_U16: return (double)static_cast<const uint16_t*>(base)[i];
        case EP_I16: return (double)static_cast<const int16_t*>(base)[i];
        case EP_I32: return (double)static_cast<const int32_t*>(base)[i];
        case EP_U32: return (double)static_cast<const uint32_t*>(base)[i];
        case EP_I64: return (double)static_cast<const int64_t*>(base)[i];
        case EP_F32: return (double)static_cast<const float*>(base)[i];
        default: return static_cast<const double*>(base)[i];
    }
}

inline bool valid_dtype(int d) { return d >= EP_U8 && d <= EP_U32; }

}  // namespace ep

// Time surface: per polarity, exp(-(t_ref - t_last)/tau) of the most recent event at each pixel, 0 where none.
//
// PARITY UNPINNED: the reference has no time-surface routine (SURVEY.md F5 — the middle channel of its MEM frame is
// a zero placeholder, events_to_image.py:58-59, and its EvRep E_T channel is a different statistic, implemented in
// ep_evrep.cu).  north_star names time surfaces, so this is the build's own definition; the checker is the numpy
// self-oracle oracle/stage3_np.py:time_surface.
//   pass 1: atomicMax of an order-preserving 64-bit key of the timestamp per (sample, polarity, pixel)
//   pass 2: key -> exp decay against the sample's last row timestamp, streaming fp32 stores, keys re-zeroed
#include <math.h>

#include "ep_common.cuh"

namespace ep {
namespace {

struct TsArgs {
    const void* x; const void* y; const void* t; const void* p;
    int xy_dtype, t_dtype, p_dtype;
    double t_div;
    const int64_t* offsets;
    int B, H, W;
    int64_t begin, end;
    unsigned long long* keys;    // [B][2][HW]
    unsigned int* bad;
};

__device__ __forceinline__ unsigned long long f64_key(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

__device__ __forceinline__ double ts_time(const TsArgs& a, int64_t i) {
    const double v = load_as_double(a.t, a.t_dtype, i);
    return a.t_div != 1.0 ? v / a.t_div : v;
}

__global__ void __launch_bounds__(256) k_ts_scatter(TsArgs a) {
    const int64_t i = a.begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.end) return;
    int lo = 0, hi = a.B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.offsets[mid] <= i) lo = mid; else hi = mid; }
    const int64_t x = __double2ll_rz(load_as_double(a.x, a.xy_dtype, i)), y = __double2ll_rz(load_as_double(a.y, a.xy_dtype, i));
    const double p = load_as_double(a.p, a.p_dtype, i);
    if (x < 0 || x >= a.W || y < 0 || y >= a.H || !(p == 1.0 || p == 0.0 || p == -1.0)) { if (a.bad) atomicAdd(a.bad, 1u); return; }
    const int64_t HW = (int64_t)a.H * a.W;
    atomicMax(a.keys + ((int64_t)lo * 2 + (p == 1.0 ? 0 : 1)) * HW + y * a.W + x, f64_key(ts_time(a, i)));
}

__global__ void __launch_bounds__(256) k_ts_finish(TsArgs a, double tau, const double* t_ref_opt, float* __restrict__ out) {
    const int64_t HW = (int64_t)a.H * a.W;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)a.B * 2 * HW) return;
    const int b = (int)(idx / (2 * HW));
    const unsigned long long k = a.keys[idx];
    a.keys[idx] = 0ull;
    float r = 0.f;
    if (k) {
        const int64_t hi = a.offsets[b + 1];
        const double t_ref = t_ref_opt ? t_ref_opt[b] : ts_time(a, hi - 1);      // default: the sample's last row
        r = (float)exp(-(t_ref - key_f64(k)) / tau);
    }
    st_stream(out + idx, r);
}

}  // namespace
}  // namespace ep

extern "C" {

size_t ep_time_surface_workspace_bytes(int batch, int height, int width) {
    return batch > 0 && height > 0 && width > 0 ? ep::align_up(sizeof(unsigned long long) * (size_t)batch * 2 * height * width, 256) : 0;
}

int ep_time_surface(void* stream, const ep_events_soa* ev, int height, int width, double tau, const double* t_ref,
                    float* out, void* workspace, size_t workspace_bytes, unsigned int* bad_count) {
    using namespace ep;
    if (!ev || !out || !workspace || ev->batch <= 0 || height <= 0 || width <= 0 || !(tau > 0.0)) return EP_EINVAL;
    if (!ev->offsets || !ev->offsets_host || !(ev->t_div != 0.0)) return EP_EINVAL;
    if (ev->t_base || ev->xy_dtype == EP_U32 || ev->t_dtype == EP_U32) return EP_EUNSUPPORTED;   // transport layouts: ep_bin_events only
    if (!valid_dtype(ev->xy_dtype) || !valid_dtype(ev->t_dtype) || !valid_dtype(ev->p_dtype)) return EP_EINVAL;
    const size_t need = ep_time_surface_workspace_bytes(ev->batch, height, width);
    if (workspace_bytes < need) return EP_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TsArgs a{ev->x, ev->y, ev->t, ev->p, ev->xy_dtype, ev->t_dtype, ev->p_dtype, ev->t_div, ev->offsets, ev->batch, height, width,
             ev->offsets_host[0], ev->offsets_host[ev->batch], static_cast<unsigned long long*>(workspace), bad_count};
    cudaError_t ce = cudaMemsetAsync(workspace, 0, need, st);
    if (ce != cudaSuccess) return (int)ce;
    const int64_t n = a.end - a.begin;
    if (n > 0) {
        k_ts_scatter<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(a);
        EP_LAUNCH_CHECK();
    }
    const int64_t cells = (int64_t)ev->batch * 2 * height * width;
    k_ts_finish<<<(unsigned)ceil_div64(cells, 256), 256, 0, st>>>(a, tau, t_ref, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

}  // extern "C"

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
    int g0, g1;                  // samples [g0, g1) of this group; their key slots are keys[(b - g0) * 2 * HW ...]
    int64_t begin, end;          // event range of the group
    int64_t n_total;
    unsigned long long* keys;    // [group][2][HW], L2-resident
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

__device__ __forceinline__ int ts_owner(const TsArgs& a, int64_t i) {
    int lo = a.g0, hi = a.g1;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.offsets[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

__device__ __forceinline__ void ts_put(const TsArgs& a, int b, int64_t x, int64_t y, double p, double t, unsigned& nbad) {
    if (x < 0 || x >= a.W || y < 0 || y >= a.H || !(p == 1.0 || p == 0.0 || p == -1.0)) { ++nbad; return; }
    const int64_t HW = (int64_t)a.H * a.W;
    atomicMax(a.keys + ((int64_t)(b - a.g0) * 2 + (p == 1.0 ? 0 : 1)) * HW + y * a.W + x, f64_key(t));
}

// CANON: x,y u16 | t i64 or f64 | p u8, 16-byte aligned: 4 consecutive events per thread with vector loads
template <bool CANON>
__global__ void __launch_bounds__(256) k_ts_scatter(TsArgs a, int64_t n_quads) {
    unsigned nbad = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = a.begin / 4 * 4 + q * 4;
        if (CANON && i0 >= a.begin && i0 + 4 <= a.end && i0 + 4 <= a.n_total) {
            const uint2 xv = ld_stream(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(a.x) + i0));
            const uint2 yv = ld_stream(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(a.y) + i0));
            const uint32_t pv = ld_stream(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(a.p) + i0));
            const longlong2 t0 = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(a.t) + i0));
            const longlong2 t1 = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(a.t) + i0 + 2));
            const uint32_t xs[4] = {xv.x & 0xffffu, xv.x >> 16, xv.y & 0xffffu, xv.y >> 16};
            const uint32_t ys[4] = {yv.x & 0xffffu, yv.x >> 16, yv.y & 0xffffu, yv.y >> 16};
            const long long tr[4] = {t0.x, t0.y, t1.x, t1.y};
            int b = ts_owner(a, i0);
            int64_t b_end = a.offsets[b + 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                while (i0 + j >= b_end) { ++b; b_end = a.offsets[b + 1]; }
                double t = a.t_dtype == EP_I64 ? (double)tr[j] : __longlong_as_double(tr[j]);
                if (a.t_div != 1.0) t /= a.t_div;
                ts_put(a, b, xs[j], ys[j], (double)((pv >> (8 * j)) & 0xffu), t, nbad);
            }
        } else {
            for (int j = 0; j < 4; ++j) {
                const int64_t i = i0 + j;
                if (i < a.begin || i >= a.end) continue;
                ts_put(a, ts_owner(a, i), __double2ll_rz(load_as_double(a.x, a.xy_dtype, i)), __double2ll_rz(load_as_double(a.y, a.xy_dtype, i)),
                       load_as_double(a.p, a.p_dtype, i), ts_time(a, i), nbad);
            }
        }
    }
    if (nbad && a.bad) atomicAdd(a.bad, nbad);
}

// One CTA row per sample of the group (blockIdx.y): the reference time is fetched once per CTA, and each thread turns two
// adjacent keys (one 16-byte load) into two fp32 decays.  (The first version derived the sample from a flat 64-bit index and
// re-read / re-divided the reference stamp in every thread: ~1300 issue slots per warp, 0.3 TB/s.)
__global__ void __launch_bounds__(256) k_ts_finish(TsArgs a, double tau, const double* t_ref_opt, float* __restrict__ out) {
    __shared__ double s_ref;
    const int slot = blockIdx.y, b = a.g0 + slot;
    const int64_t cells = 2 * (int64_t)a.H * a.W;                              // even: pairs never straddle a sample
    if (threadIdx.x == 0) {
        const int64_t lo = a.offsets[b], hi = a.offsets[b + 1];
        s_ref = t_ref_opt ? t_ref_opt[b] : (hi > lo ? ts_time(a, hi - 1) : 0.0);   // default: the sample's last row
    }
    __syncthreads();
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (c >= cells) return;
    unsigned long long* kp = a.keys + (int64_t)slot * cells + c;
    const ulonglong2 k = *reinterpret_cast<const ulonglong2*>(kp);
    if (k.x | k.y) *reinterpret_cast<ulonglong2*>(kp) = make_ulonglong2(0ull, 0ull);
    const double t_ref = s_ref;
    float2 r;
    r.x = k.x ? (float)exp(-(t_ref - key_f64(k.x)) / tau) : 0.f;
    r.y = k.y ? (float)exp(-(t_ref - key_f64(k.y)) / tau) : 0.f;
    st_stream(reinterpret_cast<float2*>(out + (int64_t)b * cells + c), r);
}

constexpr size_t kTsGroupBytes = (size_t)64 << 20;      // key slots kept in flight: L2-resident, like the binning accumulators

}  // namespace
}  // namespace ep

extern "C" {

size_t ep_time_surface_workspace_bytes(int batch, int height, int width) {
    if (!(batch > 0 && height > 0 && width > 0)) return 0;
    const size_t per = sizeof(unsigned long long) * 2 * (size_t)height * width;
    size_t g = ep::kTsGroupBytes / per;
    if (g < 1) g = 1;
    if (g > (size_t)batch) g = (size_t)batch;
    return ep::align_up(per * g, 256);
}

int ep_time_surface(void* stream, const ep_events_soa* ev, int height, int width, double tau, const double* t_ref,
                    float* out, void* workspace, size_t workspace_bytes, unsigned int* bad_count) {
    using namespace ep;
    if (!ev || !out || !workspace || ev->batch <= 0 || height <= 0 || width <= 0 || !(tau > 0.0)) return EP_EINVAL;
    if (!ev->offsets || !ev->offsets_host || !(ev->t_div != 0.0)) return EP_EINVAL;
    if (ev->t_base || ev->xy_dtype == EP_U32 || ev->t_dtype == EP_U32) return EP_EUNSUPPORTED;   // transport layouts: ep_bin_events only
    if (!valid_dtype(ev->xy_dtype) || !valid_dtype(ev->t_dtype) || !valid_dtype(ev->p_dtype)) return EP_EINVAL;
    const size_t per = sizeof(unsigned long long) * 2 * (size_t)height * width;
    if (workspace_bytes < per) return EP_EWORKSPACE;
    int G = (int)(workspace_bytes / per);
    if ((size_t)G > kTsGroupBytes / per && kTsGroupBytes / per >= 1) G = (int)(kTsGroupBytes / per);
    if (G > ev->batch) G = ev->batch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool canon = ev->xy_dtype == EP_U16 && ev->p_dtype == EP_U8 && (ev->t_dtype == EP_I64 || ev->t_dtype == EP_F64) &&
                       aligned16(ev->x) && aligned16(ev->y) && aligned16(ev->t) && aligned16(ev->p);
    TsArgs a{ev->x, ev->y, ev->t, ev->p, ev->xy_dtype, ev->t_dtype, ev->p_dtype, ev->t_div, ev->offsets, ev->batch, height, width,
             0, 0, 0, 0, ev->offsets_host[ev->batch], static_cast<unsigned long long*>(workspace), bad_count};
    cudaError_t ce = cudaMemsetAsync(workspace, 0, per * (size_t)G, st);
    if (ce != cudaSuccess) return (int)ce;
    for (int g0 = 0; g0 < ev->batch; g0 += G) {
        a.g0 = g0;
        a.g1 = g0 + G < ev->batch ? g0 + G : ev->batch;
        a.begin = ev->offsets_host[a.g0];
        a.end = ev->offsets_host[a.g1];
        const int64_t n_quads = a.end > a.begin ? ceil_div64(a.end - a.begin / 4 * 4, 4) : 0;
        if (n_quads > 0) {
            const int64_t blocks = ceil_div64(n_quads, 256), cap = (int64_t)kNumSMs * 8;
            const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
            if (canon) k_ts_scatter<true><<<grid, 256, 0, st>>>(a, n_quads);
            else k_ts_scatter<false><<<grid, 256, 0, st>>>(a, n_quads);
            EP_LAUNCH_CHECK();
        }
        const dim3 fin_grid((unsigned)ceil_div64((int64_t)height * width, 256), (unsigned)(a.g1 - a.g0));      // 2 cells per thread
        k_ts_finish<<<fin_grid, 256, 0, st>>>(a, tau, t_ref, out);
        EP_LAUNCH_CHECK();
    }
    return EP_OK;
}

}  // extern "C"

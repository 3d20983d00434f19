// Per-sample normalisers, hot-pixel filter, frame-side diff-map target, ABI housekeeping (sm_100a).
#include <math.h>

#include "ep_common.cuh"

#include <mutex>
#include <vector>

namespace ep {

unsigned long long g_launch_count = 0;

namespace {
struct ProfRec { cudaEvent_t a, b; int kind; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
bool g_prof_on = false;
cudaEvent_t g_prof_open = nullptr;
int g_prof_open_kind = 0;
}  // namespace

bool profile_enabled() { return g_prof_on; }

void profile_begin(cudaStream_t st, int kind) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventCreate(&g_prof_open);
    cudaEventRecord(g_prof_open, st);
    g_prof_open_kind = kind;
}

void profile_end(cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_open) return;
    ProfRec r;
    r.a = g_prof_open; r.kind = g_prof_open_kind;
    cudaEventCreate(&r.b);
    cudaEventRecord(r.b, st);
    g_prof.push_back(r);
    g_prof_open = nullptr;
}

namespace {

constexpr int kStatBlocks = 64;   // partial-sum blocks per sample (fixed => deterministic reduction order)

// order-preserving float <-> uint key, so a plain unsigned atomicMax orders all finite floats
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---- normalisers   dataset/pretrain/pr_n_imagenet_dataset.py:142-143, ft_n_caltech101_dataset.py:93-98
__global__ void __launch_bounds__(256) k_plane_max(const float* __restrict__ img, int64_t HW, uint32_t* __restrict__ keys) {
    const int plane = blockIdx.y;
    const float* p = img + (int64_t)plane * HW;
    float m = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, p[i]);
    m = warp_reduce(m, [](float a, float b) { return fmaxf(a, b); });
    if ((threadIdx.x & 31) == 0) atomicMax(keys + plane, f2key(m));
}

__global__ void __launch_bounds__(256) k_normalise_apply(float* __restrict__ img, int C, int64_t HW, int mode,
                                                         const uint32_t* __restrict__ keys) {
    const int plane = blockIdx.y;           // b * C + c
    const int b = plane / C, c = plane % C;
    float* p = img + (int64_t)plane * HW;
    if (mode == EP_NORM_COUNT) {
        const float den = key2f(keys[plane]) + 1.0f;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x)
            p[i] = __fmul_rn(__fsub_rn(__fdiv_rn(p[i], den), 0.5f), 2.0f);
    } else {
        if (c == 1) return;   // [0::2] only
        const float mx = fmaxf(key2f(keys[b * C + 0]), key2f(keys[b * C + 2]));
        const float factor = (mode == EP_NORM_MEM_GUARD && mx == 0.0f) ? (float)(1.0 / 0.001) : __fdiv_rn(1.0f, mx);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x)
            p[i] = __fmul_rn(p[i], factor);
    }
}

// ---- remove_hot_pixel_mem   dataset/dataset_utils/events_to_image.py:65-75
__global__ void __launch_bounds__(256) k_hot_stats(const float* __restrict__ hist, int64_t HW, float divide_by,
                                                   double* __restrict__ partial) {
    __shared__ double s_sum[8], s_sq[8];
    const int b = blockIdx.y;
    const float* c0 = hist + (int64_t)b * 3 * HW;
    const float* c2 = c0 + 2 * HW;
    double s = 0.0, q = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * HW; i += (int64_t)gridDim.x * blockDim.x) {
        float v = i < HW ? c0[i] : c2[i - HW];
        if (divide_by != 1.0f) v = __fdiv_rn(v, divide_by);
        s += (double)v;
        q += (double)v * (double)v;
    }
    s = warp_reduce(s, [](double a, double c) { return a + c; });
    q = warp_reduce(q, [](double a, double c) { return a + c; });
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = s; s_sq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0;
        for (int w = 0; w < 8; ++w) { ts += s_sum[w]; tq += s_sq[w]; }
        partial[((int64_t)b * kStatBlocks + blockIdx.x) * 2] = ts;
        partial[((int64_t)b * kStatBlocks + blockIdx.x) * 2 + 1] = tq;
    }
}

__global__ void k_hot_threshold(const double* __restrict__ partial, int B, int64_t HW, float num_stds,
                                float* __restrict__ thr) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < kStatBlocks; ++i) { s += partial[((int64_t)b * kStatBlocks + i) * 2]; q += partial[((int64_t)b * kStatBlocks + i) * 2 + 1]; }
    const double m = (double)(2 * HW);
    const double mean = s / m;
    double var = (q - m * mean * mean) / (m - 1.0);
    if (var < 0.0) var = 0.0;
    thr[b] = __fadd_rn((float)mean, __fmul_rn(num_stds, (float)sqrt(var)));   // :69 in fp32
}

__global__ void __launch_bounds__(256) k_hot_apply(float* __restrict__ hist, int64_t HW, float divide_by,
                                                   const float* __restrict__ thr) {
    const int b = blockIdx.y;
    float* c0 = hist + (int64_t)b * 3 * HW;
    const float t = thr[b];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
        float a = c0[i], m = c0[HW + i], c = c0[2 * HW + i];
        if (divide_by != 1.0f) { a = __fdiv_rn(a, divide_by); m = __fdiv_rn(m, divide_by); c = __fdiv_rn(c, divide_by); }
        if (a > t || c > t) { a = 0.0f; c = 0.0f; }   // :70-73
        c0[i] = a; c0[HW + i] = m; c0[2 * HW + i] = c;
    }
}

// ---- frame-side difference map (self-defined formula; sign flip = view_augment.py:60-63)
// Samples ride blockIdx.y (the negate flag is per sample: no per-element division); VEC: 16-byte loads and stores.
template <bool VEC>
__global__ void __launch_bounds__(256) k_diffmap(const float* __restrict__ f0, const float* __restrict__ f1,
                                                 float* __restrict__ out, int64_t n_samples, int64_t per_sample, int mode,
                                                 float eps, const uint8_t* __restrict__ negate) {
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 4 : 1);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * (VEC ? 4 : 1);
    for (int64_t s = blockIdx.y; s < n_samples; s += gridDim.y) {
        const bool neg = negate && negate[s];
        const int64_t base = s * per_sample;
        for (int64_t i = first; i < per_sample; i += stride) {
            if (VEC) {
                const float4 a = ld_stream(reinterpret_cast<const float4*>(f0 + base + i));
                const float4 b = ld_stream(reinterpret_cast<const float4*>(f1 + base + i));
                float4 d;
                if (mode == 1) {
                    d.x = __fsub_rn(logf(b.x + eps), logf(a.x + eps)); d.y = __fsub_rn(logf(b.y + eps), logf(a.y + eps));
                    d.z = __fsub_rn(logf(b.z + eps), logf(a.z + eps)); d.w = __fsub_rn(logf(b.w + eps), logf(a.w + eps));
                } else {
                    d.x = __fsub_rn(b.x, a.x); d.y = __fsub_rn(b.y, a.y); d.z = __fsub_rn(b.z, a.z); d.w = __fsub_rn(b.w, a.w);
                }
                if (neg) { d.x = -d.x; d.y = -d.y; d.z = -d.z; d.w = -d.w; }
                st_stream(reinterpret_cast<float4*>(out + base + i), d);
            } else {
                const float a = ld_stream(f0 + base + i), b = ld_stream(f1 + base + i);
                float d = mode == 1 ? __fsub_rn(logf(b + eps), logf(a + eps)) : __fsub_rn(b, a);
                st_stream(out + base + i, neg ? -d : d);
            }
        }
    }
}

}  // namespace
}  // namespace ep


namespace ep {
namespace {
constexpr int kPlaneStatBlocks = 2 * kNumSMs;

// per-channel (sum, sum of squares, max) of a (B,C,H,W) f32 tensor: fixed partition and fixed reduction order, fp64 partials
__global__ void __launch_bounds__(256) k_plane_stats_partial(const float* __restrict__ x, int batch, int channels, int64_t hw,
                                                             double* __restrict__ part) {
    __shared__ double s1[256], s2[256], sm[256];
    const int c = blockIdx.y, tid = threadIdx.x;
    const int64_t per = (int64_t)batch * hw;
    double a1 = 0.0, a2 = 0.0, am = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * 256 + tid; i < per; i += (int64_t)gridDim.x * 256) {
        const int64_t b = i / hw, r = i - b * hw;
        const double v = (double)x[((int64_t)b * channels + c) * hw + r];
        a1 += v; a2 += v * v; am = fmax(am, v);
    }
    s1[tid] = a1; s2[tid] = a2; sm[tid] = am;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o]; sm[tid] = fmax(sm[tid], sm[tid + o]); }
        __syncthreads();
    }
    if (tid == 0) {
        double* q = part + ((size_t)c * gridDim.x + blockIdx.x) * 3;
        q[0] = s1[0]; q[1] = s2[0]; q[2] = sm[0];
    }
}

__global__ void k_plane_stats_final(const double* __restrict__ part, int n_blocks, double count, double* __restrict__ out) {
    const int c = blockIdx.x;
    if (threadIdx.x) return;
    double a1 = 0.0, a2 = 0.0, am = -INFINITY;
    for (int i = 0; i < n_blocks; ++i) {
        const double* q = part + ((size_t)c * n_blocks + i) * 3;
        a1 += q[0]; a2 += q[1]; am = fmax(am, q[2]);
    }
    out[c * 4 + 0] = count; out[c * 4 + 1] = a1; out[c * 4 + 2] = a2; out[c * 4 + 3] = am;
}
}  // namespace

// out (channels,4) f64 = (count, sum, sum of squares, max) per channel; ws: 3 * channels * kPlaneStatBlocks doubles
int plane_statistics(cudaStream_t st, const float* x, int batch, int channels, int64_t hw, double* out, void* ws, size_t ws_bytes) {
    if (!x || !out || batch <= 0 || channels <= 0 || hw <= 0) return EP_EINVAL;
    if (!ws || ws_bytes < sizeof(double) * 3 * (size_t)channels * kPlaneStatBlocks) return EP_EWORKSPACE;
    double* part = static_cast<double*>(ws);
    k_plane_stats_partial<<<dim3(kPlaneStatBlocks, channels), 256, 0, st>>>(x, batch, channels, hw, part);
    EP_LAUNCH_CHECK();
    k_plane_stats_final<<<channels, 32, 0, st>>>(part, kPlaneStatBlocks, (double)batch * (double)hw, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}
}  // namespace ep

extern "C" {

int ep_abi_version(void) { return EP_ABI_VERSION; }

unsigned long long ep_launch_count(void) { return __atomic_load_n(&ep::g_launch_count, __ATOMIC_RELAXED); }

int ep_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(ep::g_prof_mu);
    for (auto& r : ep::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    ep::g_prof.clear();
    ep::g_prof_on = on != 0;
    return EP_OK;
}

int ep_profile_read(ep_profile_stats* out) {
    if (!out) return EP_EINVAL;
    std::lock_guard<std::mutex> lk(ep::g_prof_mu);
    for (int k = 0; k < 3; ++k) { out->ms[k] = 0.0; out->launches[k] = 0; }
    for (auto& r : ep::g_prof) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) return (int)e;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        out->ms[r.kind] += ms;
        out->launches[r.kind] += 1;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    ep::g_prof.clear();
    return EP_OK;
}

const char* ep_status_string(int status) {
    switch (status) {
        case EP_OK: return "ok";
        case EP_EINVAL: return "invalid argument";
        case EP_EWORKSPACE: return "workspace too small";
        case EP_EUNSUPPORTED: return "unsupported shape";
        case EP_EALIGN: return "misaligned pointer";
        default: return status > 0 ? cudaGetErrorString(static_cast<cudaError_t>(status)) : "unknown status";
    }
}

size_t ep_normalise_workspace_bytes(int batch, int channels) {
    return batch > 0 && channels > 0 ? ep::align_up(sizeof(uint32_t) * (size_t)batch * channels, 256) : 0;
}

int ep_normalise(void* stream, float* img, int batch, int channels, int height, int width, int mode, void* workspace,
                 size_t workspace_bytes) {
    if (!img || batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || !workspace) return EP_EINVAL;
    if (mode != EP_NORM_COUNT && mode != EP_NORM_MEM && mode != EP_NORM_MEM_GUARD) return EP_EINVAL;
    if (mode != EP_NORM_COUNT && channels != 3) return EP_EINVAL;
    if (workspace_bytes < ep_normalise_workspace_bytes(batch, channels)) return EP_EWORKSPACE;
    if ((int64_t)batch * channels > 65535) return EP_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t* keys = static_cast<uint32_t*>(workspace);
    cudaError_t ce = cudaMemsetAsync(keys, 0, sizeof(uint32_t) * (size_t)batch * channels, st);
    if (ce != cudaSuccess) return (int)ce;
    const int64_t HW = (int64_t)height * width;
    unsigned bx = (unsigned)ep::ceil_div64(HW, 256 * 8);
    if (bx > 64) bx = 64;
    if (bx < 1) bx = 1;
    dim3 grid(bx, (unsigned)(batch * channels));
    ep::k_plane_max<<<grid, 256, 0, st>>>(img, HW, keys);
    EP_LAUNCH_CHECK();
    ep::k_normalise_apply<<<grid, 256, 0, st>>>(img, channels, HW, mode, keys);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

size_t ep_mem_hotpixel_workspace_bytes(int batch) {
    return batch > 0 ? ep::align_up(sizeof(double) * (size_t)batch * (2 * ep::kStatBlocks + 1), 256) : 0;
}

int ep_mem_hotpixel(void* stream, float* hist, int batch, int height, int width, float divide_by, float num_stds,
                    void* workspace, size_t workspace_bytes) {
    if (!hist || batch <= 0 || height <= 0 || width <= 0 || !workspace || !(divide_by != 0.0f)) return EP_EINVAL;
    if (workspace_bytes < ep_mem_hotpixel_workspace_bytes(batch)) return EP_EWORKSPACE;
    if (batch > 65535) return EP_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(workspace);
    float* thr = reinterpret_cast<float*>(partial + (size_t)batch * 2 * ep::kStatBlocks);
    const int64_t HW = (int64_t)height * width;
    dim3 g1(ep::kStatBlocks, (unsigned)batch);
    ep::k_hot_stats<<<g1, 256, 0, st>>>(hist, HW, divide_by, partial);
    EP_LAUNCH_CHECK();
    ep::k_hot_threshold<<<(batch + 127) / 128, 128, 0, st>>>(partial, batch, HW, num_stds, thr);
    EP_LAUNCH_CHECK();
    ep::k_hot_apply<<<g1, 256, 0, st>>>(hist, HW, divide_by, thr);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_diffmap_frames(void* stream, const float* f0, const float* f1, float* out, int64_t n, int64_t per_sample,
                      int mode, float eps, const uint8_t* negate) {
    if (!f0 || !f1 || !out || n <= 0 || per_sample <= 0 || (mode != 0 && mode != 1)) return EP_EINVAL;
    if (n % per_sample) return EP_EINVAL;
    const int64_t n_samples = n / per_sample;
    const bool vec = per_sample % 4 == 0 && ep::aligned16(f0) && ep::aligned16(f1) && ep::aligned16(out);
    int64_t bx = ep::ceil_div64(per_sample, 256 * (vec ? 4 : 1));
    if (bx > ep::kNumSMs * 16) bx = ep::kNumSMs * 16;
    const dim3 grid((unsigned)bx, (unsigned)(n_samples < 65535 ? n_samples : 65535));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec) ep::k_diffmap<true><<<grid, 256, 0, st>>>(f0, f1, out, n_samples, per_sample, mode, eps, negate);
    else ep::k_diffmap<false><<<grid, 256, 0, st>>>(f0, f1, out, n_samples, per_sample, mode, eps, negate);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

size_t ep_plane_statistics_workspace_bytes(int channels) { return sizeof(double) * 3 * (size_t)(channels > 0 ? channels : 1) * ep::kPlaneStatBlocks; }

int ep_plane_statistics(void* stream, const float* x, int batch, int channels, int height, int width, double* out, void* workspace,
                        size_t workspace_bytes) {
    return ep::plane_statistics(static_cast<cudaStream_t>(stream), x, batch, channels, (int64_t)height * width, out, workspace, workspace_bytes);
}

}  // extern "C"

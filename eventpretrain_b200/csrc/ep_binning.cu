// Stage 1 — raw events -> voxel grids / polarity count frames (sm_100a).
//
// Replaces events_to_voxel_grid (dataset/dataset_utils/events_to_voxel_grid.py:4-61),
// events_to_image_ecdp / _mem (dataset/dataset_utils/events_to_image.py:6-62) and the fused
// events_reshape scale (dataset/augmentation/events_augment.py:22-26) for a ragged batch.
//
// Design:
//   * One pass over the events (13 B/event in the canonical SoA layout, 16-byte vector loads,
//     streaming cache policy).  Each event issues ONE 64-bit integer RED for the voxel grid
//     (both temporal-bilinear weights packed in one word) and, if requested, one 32-bit RED for
//     the count frame.  Measured on B200 (profiles/r01_scatter_microbench.txt) random-address
//     RED throughput is ~210 Gop/s whatever the operand width as long as the target stays
//     L2-resident, so the number of REDs per event is the cost that matters.
//   * Packed accumulator per (interval k, pixel): [ C : 20 bits | A : 44 bits ], both signed,
//         C_k = sum of p over events with floor(ts) == k,    A_k = sum of p * round(d * 2^24)
//     with d = ts - floor(ts) the fp32 fraction the reference computes.  Integer adds commute, so
//     the result is independent of event order and bit-reproducible.  The finalize pass emits
//         voxel[k] = (C_k * 2^24 - A_k + A_{k-1}) * 2^-24      (one rounding, int64 -> fp32).
//   * Samples are processed in groups whose accumulator slots fit in L2 (126 MB on B200); the
//     finalize kernel reads the slots back from L2, writes the fp32 outputs with streaming stores
//     and re-zeroes the slots, so accumulator traffic never reaches HBM.
//   * Headroom: the adds are exact modulo 2^64, so only the FINAL per-cell values must fit:
//     |C| < 2^19 and |A| < 2^43, i.e. fewer than 524288 net same-polarity events on ONE pixel in ONE
//     temporal interval of ONE sample.  The finalize kernel reports cells with |C| >= 2^18 through
//     bad_count (bit 31), which is a sound detector for samples of fewer than 786432 events; see
//     DESIGN.md "limits" for larger samples.
#include <math.h>
#include <stdlib.h>

#include "ep_common.cuh"

namespace ep {
namespace {

constexpr int kQ = 24;                       // fractional bits of the temporal weight
constexpr int kABits = 44;                   // low field of the packed accumulator
constexpr int kThreads = 256;
constexpr int kEvPerThread = 4;
constexpr uint32_t kFlagLastPlane = 1u;      // an event landed in interval num_bins-1 with d > 0
constexpr uint32_t kFlagZeroPol = 2u;        // the sample has p == 0 events (count-frame neg class)

struct __align__(16) SampleMeta {
    double t0;        // first row's timestamp (events_to_voxel_grid.py:19)
    double dT;        // last - first, 1.0 when zero (:22-25)
    uint32_t flags;
    uint32_t pad[3];
};

struct BinArgs {
    const int64_t* offsets;   // device B+1, or nullptr for the single-sample AoS entry
    int64_t single_n;
    int64_t begin, end;       // event range of this group
    int64_t n_total;          // events allocated in the arrays (vector loads stay below it)
    int64_t start4;           // begin rounded down to a multiple of kEvPerThread
    int g0, g1;               // samples [g0, g1) of this group
    int H, W, num_bins, count_channels;
    double sx, sy;
    int scaled;
    SampleMeta* meta;
    unsigned long long* vox_acc;   // [slot][num_bins][HW]
    uint32_t* cnt_acc;             // [slot][3][HW]  classes: p==1, p==0, p==-1
    unsigned int* bad_count;
};

__device__ __forceinline__ int64_t off_at(const BinArgs& a, int b) {
    return a.offsets ? a.offsets[b] : (b == 0 ? 0 : a.single_n);
}

// polarity classes: 0 -> p == 1, 1 -> p == 0, 2 -> p == -1, 3 -> anything else (unsupported)
__device__ __forceinline__ int pol_class_i(int p) { return p == 1 ? 0 : (p == 0 ? 1 : (p == -1 ? 2 : 3)); }
__device__ __forceinline__ int pol_class_d(double p) { return p == 1.0 ? 0 : (p == 0.0 ? 1 : (p == -1.0 ? 2 : 3)); }

// ---- loaders: produce (xi, yi, t, class) for kEvPerThread consecutive events --------------------
template <typename TT>
struct Ev {
    int64_t x[kEvPerThread], y[kEvPerThread];
    TT t[kEvPerThread];
    int cls[kEvPerThread];
};

// canonical SoA: x,y u16; t i64 or f64; p u8.  i0 is a multiple of 4 and bases are 16B aligned.
template <bool T_IS_I64>
struct SoaCanonLoader {
    const uint16_t* x;
    const uint16_t* y;
    const void* t;
    const uint8_t* p;
    double t_div;
    typedef double time_t_;
    __device__ __forceinline__ void load(int64_t i0, int64_t hi, const BinArgs& a, Ev<double>& e) const {
        (void)hi;
        uint2 xv, yv;
        uint32_t pv;
        double tv[4];
        if (i0 + 4 <= a.n_total) {
            xv = ld_stream(reinterpret_cast<const uint2*>(x + i0));
            yv = ld_stream(reinterpret_cast<const uint2*>(y + i0));
            pv = ld_stream(reinterpret_cast<const uint32_t*>(p + i0));
            if (T_IS_I64) {
                const longlong2 a0 = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(t) + i0));
                const longlong2 a1 = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(t) + i0 + 2));
                tv[0] = (double)a0.x; tv[1] = (double)a0.y; tv[2] = (double)a1.x; tv[3] = (double)a1.y;
            } else {
                const double2 a0 = ld_stream(reinterpret_cast<const double2*>(static_cast<const double*>(t) + i0));
                const double2 a1 = ld_stream(reinterpret_cast<const double2*>(static_cast<const double*>(t) + i0 + 2));
                tv[0] = a0.x; tv[1] = a0.y; tv[2] = a1.x; tv[3] = a1.y;
            }
        } else {   // last, partial quad of the arrays: scalar loads, nothing read past the end
            uint32_t xs_[4] = {0, 0, 0, 0}, ys_[4] = {0, 0, 0, 0};
            pv = 0;
            for (int j = 0; j < 4; ++j) {
                tv[j] = 0.0;
                if (i0 + j < a.n_total) {
                    xs_[j] = x[i0 + j]; ys_[j] = y[i0 + j];
                    pv |= (uint32_t)p[i0 + j] << (8 * j);
                    tv[j] = T_IS_I64 ? (double)static_cast<const int64_t*>(t)[i0 + j] : static_cast<const double*>(t)[i0 + j];
                }
            }
            xv = make_uint2(xs_[0] | (xs_[1] << 16), xs_[2] | (xs_[3] << 16));
            yv = make_uint2(ys_[0] | (ys_[1] << 16), ys_[2] | (ys_[3] << 16));
        }
        const uint32_t xs[4] = {xv.x & 0xffffu, xv.x >> 16, xv.y & 0xffffu, xv.y >> 16};
        const uint32_t ys[4] = {yv.x & 0xffffu, yv.x >> 16, yv.y & 0xffffu, yv.y >> 16};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (a.scaled) {
                e.x[j] = __double2ll_rz(__dmul_rn((double)xs[j], a.sx));
                e.y[j] = __double2ll_rz(__dmul_rn((double)ys[j], a.sy));
            } else {
                e.x[j] = xs[j];
                e.y[j] = ys[j];
            }
            e.t[j] = (t_div != 1.0) ? tv[j] / t_div : tv[j];
            e.cls[j] = pol_class_i((int)((pv >> (8 * j)) & 0xffu));
        }
    }
    __device__ __forceinline__ double time_at(int64_t i) const {
        double v = T_IS_I64 ? (double)static_cast<const int64_t*>(t)[i] : static_cast<const double*>(t)[i];
        return (t_div != 1.0) ? v / t_div : v;
    }
};

// any tagged SoA layout, scalar loads.  TT = float reproduces torch's fp32 time arithmetic.
template <typename TT>
struct SoaGenericLoader {
    const void* x;
    const void* y;
    const void* t;
    const void* p;
    int xy_dtype, t_dtype, p_dtype;
    double t_div;
    typedef TT time_t_;
    __device__ __forceinline__ TT time_at(int64_t i) const {
        if (sizeof(TT) == 4) {
            float v = (t_dtype == EP_F32) ? static_cast<const float*>(t)[i] : (float)load_as_double(t, t_dtype, i);
            return (TT)((t_div != 1.0) ? v / (float)t_div : v);
        }
        double v = load_as_double(t, t_dtype, i);
        return (TT)((t_div != 1.0) ? v / t_div : v);
    }
    __device__ __forceinline__ void load(int64_t i0, int64_t hi, const BinArgs& a, Ev<TT>& e) const {
#pragma unroll
        for (int j = 0; j < kEvPerThread; ++j) {
            const int64_t i = i0 + j;
            if (i >= hi || i < a.begin) { e.cls[j] = 3; e.x[j] = 0; e.y[j] = 0; e.t[j] = 0; continue; }
            double xd = load_as_double(x, xy_dtype, i), yd = load_as_double(y, xy_dtype, i);
            if (a.scaled) {
                if (xy_dtype == EP_F32) {   // numpy keeps fp32 arrays in fp32 under `*= python_float`
                    xd = (double)__fmul_rn((float)xd, (float)a.sx);
                    yd = (double)__fmul_rn((float)yd, (float)a.sy);
                } else {
                    xd = __dmul_rn(xd, a.sx);
                    yd = __dmul_rn(yd, a.sy);
                }
            }
            e.x[j] = __double2ll_rz(xd);
            e.y[j] = __double2ll_rz(yd);
            e.t[j] = time_at(i);
            e.cls[j] = pol_class_d(load_as_double(p, p_dtype, i));
        }
    }
};

// the reference's own (N,4) x,y,t,p rows
template <typename ET>   // ET = double | float (element type AND time arithmetic type)
struct AosLoader {
    const ET* ev;
    typedef ET time_t_;
    __device__ __forceinline__ ET time_at(int64_t i) const { return ev[i * 4 + 2]; }
    __device__ __forceinline__ void load(int64_t i0, int64_t hi, const BinArgs& a, Ev<ET>& e) const {
#pragma unroll
        for (int j = 0; j < kEvPerThread; ++j) {
            const int64_t i = i0 + j;
            if (i >= hi || i < a.begin) { e.cls[j] = 3; e.x[j] = 0; e.y[j] = 0; e.t[j] = 0; continue; }
            ET xv, yv, tv, pv;
            if (sizeof(ET) == 8) {
                const double2 q0 = ld_stream(reinterpret_cast<const double2*>(ev + i * 4));
                const double2 q1 = ld_stream(reinterpret_cast<const double2*>(ev + i * 4 + 2));
                xv = (ET)q0.x; yv = (ET)q0.y; tv = (ET)q1.x; pv = (ET)q1.y;
            } else {
                const float4 q = ld_stream(reinterpret_cast<const float4*>(ev + i * 4));
                xv = (ET)q.x; yv = (ET)q.y; tv = (ET)q.z; pv = (ET)q.w;
            }
            if (a.scaled) {
                if (sizeof(ET) == 8) { xv = (ET)__dmul_rn((double)xv, a.sx); yv = (ET)__dmul_rn((double)yv, a.sy); }
                else { xv = (ET)__fmul_rn((float)xv, (float)a.sx); yv = (ET)__fmul_rn((float)yv, (float)a.sy); }
            }
            e.x[j] = (sizeof(ET) == 8) ? __double2ll_rz((double)xv) : __float2ll_rz((float)xv);
            e.y[j] = (sizeof(ET) == 8) ? __double2ll_rz((double)yv) : __float2ll_rz((float)yv);
            e.t[j] = tv;
            e.cls[j] = pol_class_d((double)pv);
        }
    }
};

// ---- per-sample metadata: first/last timestamps ---------------------------------------------------
template <class Loader>
__global__ void k_sample_meta(Loader ld, BinArgs a, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    typedef typename Loader::time_t_ TT;
    const int64_t lo = off_at(a, b), hi = off_at(a, b + 1);
    SampleMeta m;
    m.t0 = 0.0; m.dT = 1.0; m.flags = 0; m.pad[0] = m.pad[1] = m.pad[2] = 0;
    if (hi > lo) {
        const TT first = ld.time_at(lo), last = ld.time_at(hi - 1);
        TT d = last - first;
        if (d == (TT)0) d = (TT)1;
        m.t0 = (double)first;
        m.dT = (double)d;
    }
    a.meta[b] = m;
}

// ---- scatter ----------------------------------------------------------------------------------------
// publish "sample has p == 0 events" (selects the neg class of the count frame, events_to_image.py:13-16);
// the flag is read through L2 so that one RED per sample, not per thread, is the steady state
__device__ __forceinline__ void publish_zero_flag(const BinArgs& a, int b) {
    if (!(__ldcg(&a.meta[b].flags) & kFlagZeroPol)) atomicOr(&a.meta[b].flags, kFlagZeroPol);
}

template <class Loader>
__global__ void __launch_bounds__(kThreads) k_scatter(Loader ld, BinArgs a) {
    typedef typename Loader::time_t_ TT;
    const int64_t i0 = a.start4 + ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kEvPerThread;
    if (i0 >= a.end) return;

    // sample that owns the first in-range event of this thread: last b with off[b] <= i
    const int64_t ifirst = i0 < a.begin ? a.begin : i0;
    int lo = a.g0, hi = a.g1;   // invariant: off[lo] <= ifirst < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off_at(a, mid) <= ifirst) lo = mid; else hi = mid;
    }
    int b = lo;
    int64_t b_end = off_at(a, b + 1);

    Ev<TT> e;
    ld.load(i0, a.end, a, e);

    const int64_t HW = (int64_t)a.H * a.W;
    const TT nbm1 = (TT)(a.num_bins - 1);
    TT t0 = (TT)a.meta[b].t0, dT = (TT)a.meta[b].dT;
    int zero_b = -1;

#pragma unroll
    for (int j = 0; j < kEvPerThread; ++j) {
        const int64_t i = i0 + j;
        if (i < a.begin || i >= a.end) continue;
        if (i >= b_end) {
            if (zero_b >= 0) { publish_zero_flag(a, zero_b); zero_b = -1; }
            do { ++b; b_end = off_at(a, b + 1); } while (i >= b_end);
            t0 = (TT)a.meta[b].t0;
            dT = (TT)a.meta[b].dT;
        }
        const int cls = e.cls[j];
        const int64_t flat = e.x[j] + e.y[j] * (int64_t)a.W;
        if (cls == 3 || flat < 0 || flat >= HW) {
            if (a.bad_count) atomicAdd(a.bad_count, 1u);
            continue;
        }
        const int slot = b - a.g0;
        if (a.count_channels) {
            atomicAdd(a.cnt_acc + ((int64_t)slot * 3 + cls) * HW + flat, 1u);
            if (cls == 1) zero_b = b;
        }
        if (a.num_bins) {
            // events_to_voxel_grid.py:34-42 in the events' own dtype
            const TT ts = nbm1 * (e.t[j] - t0) / dT;
            const TT tis = floor(ts);
            if (tis >= (TT)0 && tis < (TT)a.num_bins) {
                int k = (int)tis;
                const float d = (float)(ts - tis);
                int r = __float2int_rn(d * 16777216.0f);
                if (k == a.num_bins - 1) {
                    if (r == 0 && a.num_bins >= 2) { k -= 1; r = 1 << kQ; }   // weight 1 on the last node
                    else if (!(__ldcg(&a.meta[b].flags) & kFlagLastPlane)) atomicOr(&a.meta[b].flags, kFlagLastPlane);
                }
                const long long sgn = (cls == 0) ? 1 : -1;
                const long long delta = sgn * ((1ll << kABits) + (long long)r);
                atomicAdd(a.vox_acc + ((int64_t)slot * a.num_bins + k) * HW + flat, (unsigned long long)delta);
            }
        }
    }
    if (zero_b >= 0) publish_zero_flag(a, zero_b);
}

// ---- finalize: packed accumulators -> fp32 outputs, slots re-zeroed ---------------------------------
// Bins are processed in register-resident chunks of kFinChunk planes so that all accumulator loads of a
// chunk are in flight together (the first version issued one dependent L2 round trip per bin and was
// latency-bound: profiles/r01_binning_v1_ncu.txt).
constexpr int kFinChunk = 8;

template <int VEC>
__global__ void __launch_bounds__(256) k_finalize_voxel(BinArgs a, float* __restrict__ out_voxel,
                                                        float* __restrict__ out_sum) {
    const int64_t HW = (int64_t)a.H * a.W;
    const int64_t per = HW / VEC;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = blockIdx.y;
    if (idx >= per) return;
    const int b = a.g0 + slot;
    const int64_t pix = idx * VEC;
    const bool last_used = (a.meta[b].flags & kFlagLastPlane) != 0;
    const int B = a.num_bins;
    const int n_planes = last_used ? B : B - 1;        // plane B-1 is untouched unless flagged
    long long a_prev[VEC];
    float sum[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { a_prev[v] = 0; sum[v] = 0.f; }
    unsigned long long* acc = a.vox_acc + (int64_t)slot * B * HW + pix;
    float* o = out_voxel + (int64_t)b * B * HW + pix;
    for (int k0 = 0; k0 < B; k0 += kFinChunk) {
        unsigned long long w[kFinChunk][VEC];
#pragma unroll
        for (int j = 0; j < kFinChunk; ++j) {
            const int k = k0 + j;
            if (k < n_planes) {
                if (VEC == 2) {
                    const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(acc + (int64_t)k * HW);
                    w[j][0] = q.x; w[j][VEC - 1] = q.y;
                } else {
                    w[j][0] = acc[(int64_t)k * HW];
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) w[j][v] = 0ull;
            }
        }
#pragma unroll
        for (int j = 0; j < kFinChunk; ++j) {
            const int k = k0 + j;
            if (k >= B) break;
            if (k < n_planes) {
                if (VEC == 2) *reinterpret_cast<ulonglong2*>(acc + (int64_t)k * HW) = make_ulonglong2(0ull, 0ull);
                else acc[(int64_t)k * HW] = 0ull;
            }
            float r[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const long long A = ((long long)(w[j][v] << (64 - kABits))) >> (64 - kABits);
                const long long C = ((long long)(w[j][v] - (unsigned long long)A)) >> kABits;
                if ((C >= (1ll << 18) || C < -(1ll << 18)) && a.bad_count) atomicOr(a.bad_count, 0x80000000u);
                const long long val = C * (1ll << kQ) - A + a_prev[v];
                a_prev[v] = A;
                r[v] = __ll2float_rn(val) * (1.0f / 16777216.0f);
                sum[v] += r[v];    // voxel.sum(dim=0): sequential fp32 over bins
            }
            if (VEC == 2) st_stream(reinterpret_cast<float2*>(o + (int64_t)k * HW), make_float2(r[0], r[VEC - 1]));
            else st_stream(o + (int64_t)k * HW, r[0]);
        }
    }
    if (out_sum) {
        float* s = out_sum + (int64_t)b * HW + pix;
        if (VEC == 2) st_stream(reinterpret_cast<float2*>(s), make_float2(sum[0], sum[VEC - 1]));
        else st_stream(s, sum[0]);
    }
}

__global__ void __launch_bounds__(256) k_finalize_count(BinArgs a, float* __restrict__ out_count) {
    const int64_t HW = (int64_t)a.H * a.W;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = blockIdx.y;
    if (pix >= HW) return;
    const int b = a.g0 + slot;
    uint32_t* c = a.cnt_acc + (int64_t)slot * 3 * HW + pix;
    const uint32_t pos = c[0], zer = c[HW], neg = c[2 * HW];
    c[0] = 0; c[HW] = 0; c[2 * HW] = 0;
    // events_to_image.py:13-16: neg = rows with p == 0, or p == -1 when the sample has none
    const uint32_t n = (a.meta[b].flags & kFlagZeroPol) ? zer : neg;
    float* o = out_count + (int64_t)b * a.count_channels * HW + pix;
    st_stream(o, (float)pos);
    if (a.count_channels == 3) { st_stream(o + HW, 0.0f); st_stream(o + 2 * HW, (float)n); }
    else st_stream(o + HW, (float)n);
}

// ---- host side ----------------------------------------------------------------------------------------
struct SlotLayout {
    size_t meta_bytes, vox_bytes, cnt_bytes, slot_bytes;
};

SlotLayout slot_layout(const ep_bin_params* p, int B) {
    SlotLayout s;
    const size_t HW = (size_t)p->height * p->width;
    s.meta_bytes = align_up(sizeof(SampleMeta) * (size_t)(B > 0 ? B : 1), 256);
    s.vox_bytes = align_up(p->num_bins > 0 ? HW * p->num_bins * sizeof(unsigned long long) : 0, 256);
    s.cnt_bytes = align_up(p->count_channels > 0 ? HW * 3 * sizeof(uint32_t) : 0, 256);
    s.slot_bytes = s.vox_bytes + s.cnt_bytes;
    return s;
}

size_t l2_group_budget() {
    // accumulator bytes kept in flight per group; RED throughput on B200 is flat up to ~40 MB of
    // target footprint and degrades beyond (profiles/r01_scatter_microbench.txt)
    const char* e = getenv("EP_L2_GROUP_MB");
    long mb = e ? atol(e) : 40;
    if (mb < 1) mb = 1;
    return (size_t)mb << 20;
}

int check_params(const ep_bin_params* p) {
    if (!p || p->height <= 0 || p->width <= 0) return EP_EINVAL;
    if (p->num_bins < 0 || (p->count_channels != 0 && p->count_channels != 2 && p->count_channels != 3)) return EP_EINVAL;
    if (p->num_bins == 0 && p->count_channels == 0) return EP_EINVAL;
    if (!(p->scale_x > 0.0) || !(p->scale_y > 0.0)) return EP_EINVAL;
    if ((int64_t)p->height * p->width > (1ll << 31)) return EP_EUNSUPPORTED;
    return EP_OK;
}

template <class Loader>
int run_binning(cudaStream_t st, Loader ld, const int64_t* off_dev, const int64_t* off_host, int64_t single_n,
                int B, const ep_bin_params* p, float* out_voxel, float* out_sum, float* out_count,
                void* ws, size_t ws_bytes, unsigned int* bad) {
    const SlotLayout L = slot_layout(p, B);
    if (ws_bytes < L.meta_bytes + L.slot_bytes) return EP_EWORKSPACE;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u)) return EP_EALIGN;
    if ((p->num_bins > 0 && !out_voxel) || (p->count_channels > 0 && !out_count)) return EP_EINVAL;
    if ((reinterpret_cast<uintptr_t>(out_voxel) & 7u) || (reinterpret_cast<uintptr_t>(out_sum) & 7u)) return EP_EALIGN;
    int G = (int)((ws_bytes - L.meta_bytes) / L.slot_bytes);
    const size_t budget = l2_group_budget();
    const int g_l2 = (int)(budget / L.slot_bytes) > 0 ? (int)(budget / L.slot_bytes) : 1;
    if (G > g_l2) G = g_l2;
    if (G > B) G = B;
    if (G > 65535) G = 65535;

    BinArgs a;
    a.offsets = off_dev; a.single_n = single_n;
    a.H = p->height; a.W = p->width; a.num_bins = p->num_bins; a.count_channels = p->count_channels;
    a.sx = p->scale_x; a.sy = p->scale_y; a.scaled = (p->scale_x != 1.0 || p->scale_y != 1.0);
    a.meta = reinterpret_cast<SampleMeta*>(ws);
    char* slots = static_cast<char*>(ws) + L.meta_bytes;
    a.bad_count = bad;
    a.begin = a.end = a.start4 = 0; a.g0 = a.g1 = 0;
    a.n_total = off_host ? off_host[B] : single_n;
    a.vox_acc = nullptr; a.cnt_acc = nullptr;

    profile_begin(st, kProfOther);
    k_sample_meta<Loader><<<(B + 127) / 128, 128, 0, st>>>(ld, a, B);
    profile_end(st);
    EP_LAUNCH_CHECK();
    cudaError_t ce = cudaMemsetAsync(slots, 0, (size_t)G * L.slot_bytes, st);
    if (ce != cudaSuccess) return (int)ce;

    const int64_t HW = (int64_t)p->height * p->width;
    for (int g0 = 0; g0 < B; g0 += G) {
        const int g1 = (g0 + G < B) ? g0 + G : B;
        a.g0 = g0; a.g1 = g1;
        a.begin = off_host ? off_host[g0] : 0;
        a.end = off_host ? off_host[g1] : single_n;
        a.start4 = a.begin / kEvPerThread * kEvPerThread;
        a.vox_acc = reinterpret_cast<unsigned long long*>(slots);
        a.cnt_acc = reinterpret_cast<uint32_t*>(slots + (size_t)G * L.vox_bytes);
        if (a.end > a.begin) {
            const int64_t nthreads = ceil_div64(a.end - a.start4, kEvPerThread);
            const int64_t nblocks = ceil_div64(nthreads, kThreads);
            if (nblocks > 0x7fffffffLL) return EP_EUNSUPPORTED;
            profile_begin(st, kProfScatter);
            k_scatter<Loader><<<(unsigned)nblocks, kThreads, 0, st>>>(ld, a);
            profile_end(st);
            EP_LAUNCH_CHECK();
        }
        if (p->num_bins > 0) {
            profile_begin(st, kProfFinalize);
            if (HW % 2 == 0) {
                dim3 grid((unsigned)ceil_div64(HW / 2, 256), (unsigned)(g1 - g0));
                k_finalize_voxel<2><<<grid, 256, 0, st>>>(a, out_voxel, out_sum);
            } else {
                dim3 grid((unsigned)ceil_div64(HW, 256), (unsigned)(g1 - g0));
                k_finalize_voxel<1><<<grid, 256, 0, st>>>(a, out_voxel, out_sum);
            }
            profile_end(st);
            EP_LAUNCH_CHECK();
        }
        if (p->count_channels > 0) {
            dim3 grid((unsigned)ceil_div64(HW, 256), (unsigned)(g1 - g0));
            profile_begin(st, kProfFinalize);
            k_finalize_count<<<grid, 256, 0, st>>>(a, out_count);
            profile_end(st);
            EP_LAUNCH_CHECK();
        }
    }
    return EP_OK;
}

}  // namespace
}  // namespace ep

extern "C" {

size_t ep_bin_events_workspace_bytes(const ep_bin_params* prm, int batch, size_t* min_bytes) {
    if (ep::check_params(prm) != EP_OK || batch <= 0) { if (min_bytes) *min_bytes = 0; return 0; }
    const ep::SlotLayout L = ep::slot_layout(prm, batch);
    if (min_bytes) *min_bytes = L.meta_bytes + L.slot_bytes;
    size_t g = ep::l2_group_budget() / L.slot_bytes;
    if (g < 1) g = 1;
    if (g > (size_t)batch) g = (size_t)batch;
    return L.meta_bytes + g * L.slot_bytes;
}

int ep_bin_events(void* stream, const ep_events_soa* ev, const ep_bin_params* prm, float* out_voxel,
                  float* out_voxel_sum, float* out_count, void* workspace, size_t workspace_bytes,
                  unsigned int* bad_count) {
    using namespace ep;
    int rc = check_params(prm);
    if (rc != EP_OK) return rc;
    if (!ev || ev->batch <= 0 || !ev->offsets || !ev->offsets_host) return EP_EINVAL;
    if (!valid_dtype(ev->xy_dtype) || !valid_dtype(ev->t_dtype) || !valid_dtype(ev->p_dtype)) return EP_EINVAL;
    if (!(ev->t_div != 0.0)) return EP_EINVAL;
    const int B = ev->batch;
    for (int b = 0; b < B; ++b) if (ev->offsets_host[b + 1] < ev->offsets_host[b]) return EP_EINVAL;
    if (ev->offsets_host[B] > ev->offsets_host[0] && (!ev->x || !ev->y || !ev->t || !ev->p)) return EP_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool canon = ev->xy_dtype == EP_U16 && ev->p_dtype == EP_U8 && !prm->time_f32 &&
                       (ev->t_dtype == EP_I64 || ev->t_dtype == EP_F64) && aligned16(ev->x) && aligned16(ev->y) &&
                       aligned16(ev->t) && aligned16(ev->p);
    if (canon) {
        if (ev->t_dtype == EP_I64) {
            SoaCanonLoader<true> ld{static_cast<const uint16_t*>(ev->x), static_cast<const uint16_t*>(ev->y), ev->t,
                                    static_cast<const uint8_t*>(ev->p), ev->t_div};
            return run_binning(st, ld, ev->offsets, ev->offsets_host, 0, B, prm, out_voxel, out_voxel_sum, out_count,
                               workspace, workspace_bytes, bad_count);
        }
        SoaCanonLoader<false> ld{static_cast<const uint16_t*>(ev->x), static_cast<const uint16_t*>(ev->y), ev->t,
                                 static_cast<const uint8_t*>(ev->p), ev->t_div};
        return run_binning(st, ld, ev->offsets, ev->offsets_host, 0, B, prm, out_voxel, out_voxel_sum, out_count,
                           workspace, workspace_bytes, bad_count);
    }
    if (prm->time_f32) {
        SoaGenericLoader<float> ld{ev->x, ev->y, ev->t, ev->p, ev->xy_dtype, ev->t_dtype, ev->p_dtype, ev->t_div};
        return run_binning(st, ld, ev->offsets, ev->offsets_host, 0, B, prm, out_voxel, out_voxel_sum, out_count,
                           workspace, workspace_bytes, bad_count);
    }
    SoaGenericLoader<double> ld{ev->x, ev->y, ev->t, ev->p, ev->xy_dtype, ev->t_dtype, ev->p_dtype, ev->t_div};
    return run_binning(st, ld, ev->offsets, ev->offsets_host, 0, B, prm, out_voxel, out_voxel_sum, out_count,
                       workspace, workspace_bytes, bad_count);
}

int ep_bin_events_aos(void* stream, const ep_events_aos* ev, const ep_bin_params* prm, float* out_voxel,
                      float* out_voxel_sum, float* out_count, void* workspace, size_t workspace_bytes,
                      unsigned int* bad_count) {
    using namespace ep;
    int rc = check_params(prm);
    if (rc != EP_OK) return rc;
    if (!ev || ev->n < 0 || (ev->n > 0 && !ev->events)) return EP_EINVAL;
    if (ev->dtype != EP_F64 && ev->dtype != EP_F32) return EP_EINVAL;
    if (!aligned16(ev->events)) return EP_EALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ev->dtype == EP_F64) {
        AosLoader<double> ld{static_cast<const double*>(ev->events)};
        return run_binning(st, ld, nullptr, nullptr, ev->n, 1, prm, out_voxel, out_voxel_sum, out_count, workspace,
                           workspace_bytes, bad_count);
    }
    AosLoader<float> ld{static_cast<const float*>(ev->events)};
    return run_binning(st, ld, nullptr, nullptr, ev->n, 1, prm, out_voxel, out_voxel_sum, out_count, workspace,
                       workspace_bytes, bad_count);
}

}  // extern "C"

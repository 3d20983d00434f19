// Shared pieces of the binning kernels (global-RED path and the tiled shared-memory path):
// constants, per-sample metadata, argument block, event loaders.
#pragma once
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
constexpr uint32_t kFlagIntTime = 4u;        // integer-tick stamps spanning < 2^32 ticks: fixed-point time arithmetic

struct __align__(16) SampleMeta {
    double t0;        // first row's timestamp (events_to_voxel_grid.py:19)
    double dT;        // last - first, 1.0 when zero (:22-25)
    double scale_raw; // (num_bins-1) / (dT * t_div): ts = (t_raw - t0_raw) * scale_raw on the canonical fast path
    double t0_raw;    // first row's raw stamp (before t_div) when the stamps are fp64
    int64_t t0_ticks; // first row's raw stamp when the stamps are int64 ticks
    uint32_t flags;
    // kFlagIntTime: rn(ts * 2^24) = (dt * tmul + thalf) >> tshift for 0 <= dt < 2^32 ticks (ticks_to_v)
    uint32_t tmul, tshift, thalf;
};

struct BinArgs {
    const int64_t* offsets;   // device B+1, or nullptr for the single-sample AoS entry
    int64_t single_n;
    int64_t begin, end;       // event range of this group
    int64_t n_total;          // events allocated in the arrays (vector loads stay below it)
    int64_t start4;           // begin rounded down to a multiple of kEvPerThread
    int64_t n_tiles;          // 1024-event tiles of the group (k_scatter grid-stride loop)
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
    TT t[kEvPerThread];           // timestamp value (raw stamp for the canonical loaders, see kFastTime)
    int64_t ti[kEvPerThread];     // raw int64 ticks (canonical int64 loader only)
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
    // Canonical fast path: the per-event time arithmetic is ts = (t_raw - t0_raw) * scale_raw (one exact integer
    // or fp64 subtraction, one fp64 multiply) instead of the reference's (t/div - t0/div) * (bins-1) / dT with two
    // fp64 divisions; the two differ by O(1e-13) in ts, far below the 2^-25 weight quantum.
    static constexpr bool kFastTime = true;
    static constexpr bool kTicks = T_IS_I64;
    static constexpr bool kPrefetch = true;      // load_raw / decode: the loads of the next tile fly during this tile's REDs
    static constexpr bool kBlocked = false;
    struct Raw { uint2 xv, yv; uint32_t pv; longlong2 a0, a1; };
    __device__ __forceinline__ void load(int64_t i0, int64_t hi, const BinArgs& a, Ev<double>& e) const {
        Raw r;
        load_raw(i0, a, r);
        decode(r, i0, hi, a, e);
    }
    __device__ __forceinline__ void load_raw(int64_t i0, const BinArgs& a, Raw& r) const {
        uint2& xv = r.xv; uint2& yv = r.yv; uint32_t& pv = r.pv; longlong2& a0 = r.a0; longlong2& a1 = r.a1;
        if (i0 + 4 <= a.n_total) {
            xv = ld_stream(reinterpret_cast<const uint2*>(x + i0));
            yv = ld_stream(reinterpret_cast<const uint2*>(y + i0));
            pv = ld_stream(reinterpret_cast<const uint32_t*>(p + i0));
            a0 = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(t) + i0));
            a1 = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(t) + i0 + 2));
        } else {   // last, partial quad of the arrays: scalar loads, nothing read past the end
            uint32_t xs_[4] = {0, 0, 0, 0}, ys_[4] = {0, 0, 0, 0};
            long long tv_[4] = {0, 0, 0, 0};
            pv = 0;
            for (int j = 0; j < 4; ++j) {
                if (i0 + j < a.n_total) {
                    xs_[j] = x[i0 + j]; ys_[j] = y[i0 + j];
                    pv |= (uint32_t)p[i0 + j] << (8 * j);
                    tv_[j] = static_cast<const int64_t*>(t)[i0 + j];
                }
            }
            xv = make_uint2(xs_[0] | (xs_[1] << 16), xs_[2] | (xs_[3] << 16));
            yv = make_uint2(ys_[0] | (ys_[1] << 16), ys_[2] | (ys_[3] << 16));
            a0 = make_longlong2(tv_[0], tv_[1]); a1 = make_longlong2(tv_[2], tv_[3]);
        }
    }
    __device__ __forceinline__ void decode(const Raw& r, int64_t i0, int64_t hi, const BinArgs& a, Ev<double>& e) const {
        (void)i0; (void)hi;
        const uint2 xv = r.xv, yv = r.yv;
        const uint32_t pv = r.pv;
        const longlong2 a0 = r.a0, a1 = r.a1;
        const long long raw[4] = {a0.x, a0.y, a1.x, a1.y};
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
            e.ti[j] = raw[j];
            e.t[j] = T_IS_I64 ? 0.0 : __longlong_as_double(raw[j]);
            e.cls[j] = pol_class_i((int)((pv >> (8 * j)) & 0xffu));
        }
    }
    // 32-bit fields for the lean scatter path (integer ticks only): x, y, polarity byte, raw ticks
    static constexpr bool kLean = T_IS_I64;
    __device__ __forceinline__ void decode_lean(const Raw& r, uint32_t (&xs)[4], uint32_t (&ys)[4], uint32_t (&pb)[4],
                                                int64_t (&ti)[4]) const {
        xs[0] = r.xv.x & 0xffffu; xs[1] = r.xv.x >> 16; xs[2] = r.xv.y & 0xffffu; xs[3] = r.xv.y >> 16;
        ys[0] = r.yv.x & 0xffffu; ys[1] = r.yv.x >> 16; ys[2] = r.yv.y & 0xffffu; ys[3] = r.yv.y >> 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = (r.pv >> (8 * j)) & 0xffu;
        ti[0] = r.a0.x; ti[1] = r.a0.y; ti[2] = r.a1.x; ti[3] = r.a1.y;
    }
    __device__ __forceinline__ double time_at(int64_t i) const {
        double v = T_IS_I64 ? (double)static_cast<const int64_t*>(t)[i] : static_cast<const double*>(t)[i];
        return (t_div != 1.0) ? v / t_div : v;
    }
    __device__ __forceinline__ double time_of(int64_t i, int, int64_t) const { return time_at(i); }
    __device__ __forceinline__ double raw_at(int64_t i) const {
        return T_IS_I64 ? (double)static_cast<const int64_t*>(t)[i] : static_cast<const double*>(t)[i];
    }
    __device__ __forceinline__ int64_t ticks_at(int64_t i, int64_t) const {
        return T_IS_I64 ? static_cast<const int64_t*>(t)[i] : 0;
    }
    __device__ __forceinline__ double div() const { return t_div; }
};

// compact transport layout (8 B/event): x,y u16; tp u32 = ticks relative to the sample's base in bits 0..30 and the
// polarity in bit 31; t_base[b] int64 ticks per sample.  Same arithmetic as the int64 canonical layout: the metadata
// uses (t_base + rel) / t_div like SoaCanonLoader<true>, and the per-event difference rel - rel_first is the same exact
// integer, so both layouts give bit-identical results.
struct SoaCompactLoader {
    const uint16_t* x;
    const uint16_t* y;
    const uint32_t* tp;
    const int64_t* t_base;
    double t_div;
    typedef double time_t_;
    static constexpr bool kFastTime = true;
    static constexpr bool kTicks = true;
    static constexpr bool kPrefetch = true;
    static constexpr bool kBlocked = false;
    struct Raw { uint2 xv, yv; uint4 tv; };
    __device__ __forceinline__ void load(int64_t i0, int64_t hi, const BinArgs& a, Ev<double>& e) const {
        Raw r;
        load_raw(i0, a, r);
        decode(r, i0, hi, a, e);
    }
    __device__ __forceinline__ void load_raw(int64_t i0, const BinArgs& a, Raw& r) const {
        uint2& xv = r.xv; uint2& yv = r.yv; uint4& tv = r.tv;
        if (i0 + 4 <= a.n_total) {
            xv = ld_stream(reinterpret_cast<const uint2*>(x + i0));
            yv = ld_stream(reinterpret_cast<const uint2*>(y + i0));
            tv = ld_stream(reinterpret_cast<const uint4*>(tp + i0));
        } else {
            uint32_t xs_[4] = {0, 0, 0, 0}, ys_[4] = {0, 0, 0, 0}, ts_[4] = {0, 0, 0, 0};
            for (int j = 0; j < 4; ++j)
                if (i0 + j < a.n_total) { xs_[j] = x[i0 + j]; ys_[j] = y[i0 + j]; ts_[j] = tp[i0 + j]; }
            xv = make_uint2(xs_[0] | (xs_[1] << 16), xs_[2] | (xs_[3] << 16));
            yv = make_uint2(ys_[0] | (ys_[1] << 16), ys_[2] | (ys_[3] << 16));
            tv = make_uint4(ts_[0], ts_[1], ts_[2], ts_[3]);
        }
    }
    __device__ __forceinline__ void decode(const Raw& r, int64_t i0, int64_t hi, const BinArgs& a, Ev<double>& e) const {
        (void)i0; (void)hi;
        const uint2 xv = r.xv, yv = r.yv;
        const uint4 tv = r.tv;
        const uint32_t raw[4] = {tv.x, tv.y, tv.z, tv.w};
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
            e.ti[j] = (int64_t)(raw[j] & 0x7fffffffu);
            e.t[j] = 0.0;
            e.cls[j] = (raw[j] >> 31) ? 0 : 1;           // polarity bit: 1 = positive, 0 = negative (the p == 0 class)
        }
    }
    static constexpr bool kLean = true;
    __device__ __forceinline__ void decode_lean(const Raw& r, uint32_t (&xs)[4], uint32_t (&ys)[4], uint32_t (&pb)[4],
                                                int64_t (&ti)[4]) const {
        xs[0] = r.xv.x & 0xffffu; xs[1] = r.xv.x >> 16; xs[2] = r.xv.y & 0xffffu; xs[3] = r.xv.y >> 16;
        ys[0] = r.yv.x & 0xffffu; ys[1] = r.yv.x >> 16; ys[2] = r.yv.y & 0xffffu; ys[3] = r.yv.y >> 16;
        const uint32_t raw[4] = {r.tv.x, r.tv.y, r.tv.z, r.tv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { pb[j] = raw[j] >> 31; ti[j] = (int64_t)(raw[j] & 0x7fffffffu); }
    }
    __device__ __forceinline__ double time_of(int64_t i, int b, int64_t) const {
        const double v = (double)(t_base[b] + (int64_t)(tp[i] & 0x7fffffffu));
        return (t_div != 1.0) ? v / t_div : v;
    }
    __device__ __forceinline__ double raw_at(int64_t) const { return 0.0; }
    __device__ __forceinline__ int64_t ticks_at(int64_t i, int64_t) const { return (int64_t)(tp[i] & 0x7fffffffu); }
    __device__ __forceinline__ double div() const { return t_div; }
};

// packed transport layouts: one uint32 word x | y << 11 | polarity << 22 | tick bits << 23 per event, the ticks counted
// from a base that changes every BLOCK events of the arrays: for the events of a sample that share the block of the
// sample's first event the base is the sample's own (t_base[b], int64), for every later block g = i / BLOCK it is
// t_base[b] + blk_base[g].
//   5 B/event (TICK_BYTE = true):  BLOCK = 1024, ticks = 17 bits = (word >> 23) << 8 | one extra byte per event
//   4 B/event (TICK_BYTE = false): BLOCK = 256,  ticks = 9 bits  = word >> 23
// Same integer time arithmetic as the other tick layouts once the base is added, so results are bit-identical.
template <bool TICK_BYTE>
struct SoaPackedLoader {
    const uint32_t* w;          // [N]
    const uint8_t* tl;          // [N] (5 B/event form only)
    const uint32_t* blk_base;   // [ceil(N / BLOCK)]
    const int64_t* t_base;      // [B]
    double t_div;
    typedef double time_t_;
    static constexpr int kBlockShift = TICK_BYTE ? 10 : 8;
    static constexpr bool kFastTime = true;
    static constexpr bool kTicks = true;
    static constexpr bool kPrefetch = true;
    static constexpr bool kBlocked = true;       // e.ti holds block-relative ticks: the scatter loop adds block_base()
    struct Raw { uint4 wv; uint32_t tv; };
    __device__ __forceinline__ void load(int64_t i0, int64_t hi, const BinArgs& a, Ev<double>& e) const {
        Raw r;
        load_raw(i0, a, r);
        decode(r, i0, hi, a, e);
    }
    __device__ __forceinline__ void load_raw(int64_t i0, const BinArgs& a, Raw& r) const {
        r.tv = 0;
        if (i0 + 4 <= a.n_total) {
            r.wv = ld_stream(reinterpret_cast<const uint4*>(w + i0));
            if (TICK_BYTE) r.tv = ld_stream(reinterpret_cast<const uint32_t*>(tl + i0));
        } else {
            uint32_t ws_[4] = {0, 0, 0, 0};
            for (int j = 0; j < 4; ++j)
                if (i0 + j < a.n_total) {
                    ws_[j] = w[i0 + j];
                    if (TICK_BYTE) r.tv |= (uint32_t)tl[i0 + j] << (8 * j);
                }
            r.wv = make_uint4(ws_[0], ws_[1], ws_[2], ws_[3]);
        }
    }
    __device__ __forceinline__ void decode(const Raw& r, int64_t i0, int64_t hi, const BinArgs& a, Ev<double>& e) const {
        (void)i0; (void)hi;
        const uint32_t ws[4] = {r.wv.x, r.wv.y, r.wv.z, r.wv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t xs = ws[j] & 0x7ffu, ys = (ws[j] >> 11) & 0x7ffu;
            if (a.scaled) {
                e.x[j] = __double2ll_rz(__dmul_rn((double)xs, a.sx));
                e.y[j] = __double2ll_rz(__dmul_rn((double)ys, a.sy));
            } else {
                e.x[j] = xs;
                e.y[j] = ys;
            }
            e.ti[j] = TICK_BYTE ? (int64_t)(((ws[j] >> 23) << 8) | ((r.tv >> (8 * j)) & 0xffu)) : (int64_t)(ws[j] >> 23);
            e.t[j] = 0.0;
            e.cls[j] = ((ws[j] >> 22) & 1u) ? 0 : 1;     // polarity bit: 1 = positive, 0 = negative (the p == 0 class)
        }
    }
    static constexpr bool kLean = true;
    __device__ __forceinline__ void decode_lean(const Raw& r, uint32_t (&xs)[4], uint32_t (&ys)[4], uint32_t (&pb)[4],
                                                int64_t (&ti)[4]) const {
        const uint32_t ws[4] = {r.wv.x, r.wv.y, r.wv.z, r.wv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            xs[j] = ws[j] & 0x7ffu; ys[j] = (ws[j] >> 11) & 0x7ffu; pb[j] = (ws[j] >> 22) & 1u;
            ti[j] = TICK_BYTE ? (int64_t)(((ws[j] >> 23) << 8) | ((r.tv >> (8 * j)) & 0xffu)) : (int64_t)(ws[j] >> 23);
        }
    }
    // ticks to add to an event's block-relative count: 0 inside the block where its sample starts
    __device__ __forceinline__ int64_t block_base(int64_t i, int64_t s_lo) const {
        return ((i >> kBlockShift) == (s_lo >> kBlockShift)) ? 0 : (int64_t)__ldg(blk_base + (i >> kBlockShift));
    }
    __device__ __forceinline__ int64_t ticks_at(int64_t i, int64_t s_lo) const {
        const int64_t rel = TICK_BYTE ? (int64_t)(((w[i] >> 23) << 8) | (uint32_t)tl[i]) : (int64_t)(w[i] >> 23);
        return block_base(i, s_lo) + rel;
    }
    __device__ __forceinline__ double time_of(int64_t i, int b, int64_t s_lo) const {
        const double v = (double)(t_base[b] + ticks_at(i, s_lo));
        return (t_div != 1.0) ? v / t_div : v;
    }
    __device__ __forceinline__ double raw_at(int64_t) const { return 0.0; }
    __device__ __forceinline__ double div() const { return t_div; }
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
    static constexpr bool kFastTime = false;
    static constexpr bool kTicks = false;
    static constexpr bool kPrefetch = false;
    static constexpr bool kBlocked = false;
    static constexpr bool kLean = false;
    struct Raw {};
    __device__ __forceinline__ void load_raw(int64_t, const BinArgs&, Raw&) const {}
    __device__ __forceinline__ void decode(const Raw&, int64_t i0, int64_t hi, const BinArgs& a, Ev<TT>& e) const { load(i0, hi, a, e); }
    __device__ __forceinline__ double raw_at(int64_t) const { return 0.0; }
    __device__ __forceinline__ int64_t ticks_at(int64_t, int64_t) const { return 0; }
    __device__ __forceinline__ double div() const { return 1.0; }
    __device__ __forceinline__ TT time_of(int64_t i, int, int64_t) const { return time_at(i); }
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
    static constexpr bool kFastTime = false;
    static constexpr bool kTicks = false;
    static constexpr bool kPrefetch = false;
    static constexpr bool kBlocked = false;
    static constexpr bool kLean = false;
    struct Raw {};
    __device__ __forceinline__ void load_raw(int64_t, const BinArgs&, Raw&) const {}
    __device__ __forceinline__ void decode(const Raw&, int64_t i0, int64_t hi, const BinArgs& a, Ev<ET>& e) const { load(i0, hi, a, e); }
    __device__ __forceinline__ double raw_at(int64_t) const { return 0.0; }
    __device__ __forceinline__ int64_t ticks_at(int64_t, int64_t) const { return 0; }
    __device__ __forceinline__ double div() const { return 1.0; }
    __device__ __forceinline__ ET time_at(int64_t i) const { return ev[i * 4 + 2]; }
    __device__ __forceinline__ ET time_of(int64_t i, int, int64_t) const { return time_at(i); }
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

// ---- temporal-bilinear weights of one event (events_to_voxel_grid.py:34-42, in the events' own dtype) ----
// Returns false when the event falls outside the bins (:44-45).  k is the interval index the packed
// weights are filed under and r = rn(d * 2^24) the right-node weight; an event exactly on the last node
// (ts == num_bins-1, d == 0) is filed under interval num_bins-2 with r = 2^24 so that the last interval
// plane stays untouched for normal (time-sorted) input; *last_plane reports the rare other case.
template <typename TT>
__device__ __forceinline__ bool quantise_ts(TT ts, int num_bins, int& k, int& r, bool& last_plane) {
    const TT tis = floor(ts);
    last_plane = false;
    if (!(tis >= (TT)0 && tis < (TT)num_bins)) return false;
    k = (int)tis;
    const float d = (float)(ts - tis);
    r = __float2int_rn(d * 16777216.0f);
    if (k == num_bins - 1) {
        if (r == 0 && num_bins >= 2) { k -= 1; r = 1 << kQ; }
        else last_plane = true;
    }
    return true;
}

// Integer-tick layouts: v = rn(ts * 2^24) in fixed point.  ts = (bins-1) * dt / dT with dt, dT integer ticks, so
//   v = (dt * tmul + thalf) >> tshift,   tmul = floor((bins-1) * 2^(24+tshift) / dT) < 2^32,  thalf = 2^(tshift-1):
// one 32x32->64 multiply-add and a funnel shift per event instead of the reference's fp64 chain.  tmul is short of the
// exact ratio by < 1, so v is short of the exact rn() by < dT / 2^tshift <= 1/2 weight quantum of 2^-24 (checked in k_sample_meta; and
// v(dT) is exactly (bins-1) << 24): against the fp64 expression this moves a weight by at most one quantum, 6e-8.
// Returns false when the event lies outside [0, bins) (events_to_voxel_grid.py:44-45).
__device__ __forceinline__ bool ticks_to_v(int64_t dt, uint32_t tmul, uint32_t tshift, uint32_t thalf, uint32_t v_end,
                                           uint32_t& v) {
    const uint64_t q = (uint64_t)(uint32_t)dt * tmul + thalf;
    const uint32_t hi = (uint32_t)(q >> 32);
    v = __funnelshift_r((uint32_t)q, hi, tshift);
    return (((uint32_t)((uint64_t)dt >> 32)) | (hi >> tshift)) == 0u && v < v_end;
}

// v -> (k, r) with the last-node fold of quantise_ts
__device__ __forceinline__ void split_v(uint32_t v, int num_bins, int& k, int& r, bool& last_plane) {
    k = (int)(v >> kQ);
    r = (int)(v & ((1u << kQ) - 1u));
    last_plane = false;
    if (k == num_bins - 1) {
        if (r == 0 && num_bins >= 2) { k -= 1; r = 1 << kQ; }
        else last_plane = true;
    }
}

// event j of a loaded quad -> (k, r); Loader::kFastTime selects the multiply form (canonical layout)
template <class Loader, typename TT>
__device__ __forceinline__ bool voxel_weights(const Ev<TT>& e, int j, const SampleMeta& m, int num_bins, int& k, int& r,
                                              bool& last_plane) {
    if (Loader::kFastTime && Loader::kTicks && (m.flags & kFlagIntTime)) {
        uint32_t v;
        last_plane = false;
        if (!ticks_to_v(e.ti[j] - m.t0_ticks, m.tmul, m.tshift, m.thalf, (uint32_t)num_bins << kQ, v)) return false;
        split_v(v, num_bins, k, r, last_plane);
        return true;
    }
    if (Loader::kFastTime) {
        const double dt = Loader::kTicks ? (double)(e.ti[j] - m.t0_ticks) : ((double)e.t[j] - m.t0_raw);
        return quantise_ts<double>(dt * m.scale_raw, num_bins, k, r, last_plane);
    }
    const TT ts = (TT)(num_bins - 1) * (e.t[j] - (TT)m.t0) / (TT)m.dT;
    return quantise_ts<TT>(ts, num_bins, k, r, last_plane);
}

// ---- per-sample metadata: first/last timestamps ---------------------------------------------------
template <class Loader>
__global__ void k_sample_meta(Loader ld, BinArgs a, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    typedef typename Loader::time_t_ TT;
    const int64_t lo = off_at(a, b), hi = off_at(a, b + 1);
    SampleMeta m;
    m.t0 = 0.0; m.dT = 1.0; m.flags = 0; m.tmul = 0; m.tshift = 0; m.thalf = 0; m.scale_raw = 0.0; m.t0_raw = 0.0;
    m.t0_ticks = 0;
    if (hi > lo) {
        const TT first = ld.time_of(lo, b, lo), last = ld.time_of(hi - 1, b, lo);
        TT d = last - first;
        if (d == (TT)0) d = (TT)1;
        m.t0 = (double)first;
        m.dT = (double)d;
        m.t0_raw = ld.raw_at(lo);
        m.t0_ticks = ld.ticks_at(lo, lo);
        m.scale_raw = (double)(a.num_bins - 1) / ((double)d * ld.div());
        if (Loader::kTicks && a.num_bins >= 1 && a.num_bins <= 64) {
            // deltaT == 0 (-> 1.0 s, :24-25) and last < first (unsorted rows) keep the fp64 expression
            const int64_t dticks = ld.ticks_at(hi - 1, lo) - m.t0_ticks;
            if (dticks > 0 && dticks < (1ll << 32)) {
                const uint64_t nbm1 = (uint64_t)(a.num_bins - 1);
                int s = 31;
                uint64_t mul = 0;
                for (; s >= 0; --s) {
                    mul = (nbm1 << (kQ + s)) / (uint64_t)dticks;
                    if (mul < (1ull << 32)) break;
                }
                if (s >= 0 && (uint64_t)dticks <= ((1ull << s) >> 1) + (s == 0)) {   // truncation of tmul costs < 1/2 quantum
                    m.tmul = (uint32_t)mul; m.tshift = (uint32_t)s; m.thalf = s ? (1u << (s - 1)) : 0u;
                    m.flags |= kFlagIntTime;
                }
            }
        }
    }
    a.meta[b] = m;
}


}  // namespace
}  // namespace ep

// Stage 1, second path (opt-in, method "banded") — route + banded shared-memory sweep (sm_100a).
//
// Same arithmetic and outputs as ep_binning.cu (events_to_voxel_grid.py:4-61, events_to_image.py:6-62), bit for bit,
// but no global atomics and no accumulator round trip through L2.  Measured on B200 it matches the global-RED kernels on
// evenly spread events and is ~30 % slower on edge-concentrated ones, which is why it is not the default.
//
//   route  (k_route)   one pass over the events of a sample group.  A CTA takes a chunk of 4096 consecutive events of one
//                      sample: pass A histograms them by band (shared-memory RED), a scan turns the histogram into
//                      cursors, pass B computes per event (cell, interval k, r = rn(d * 2^24), polarity) — integer
//                      fixed-point time arithmetic for tick stamps (ticks_to_v) — and claims a slot with one returning
//                      ATOMS; the chunk goes back to global memory sorted by band, coalesced, as 8-byte records
//                      {cell in band, r | k << 25 | negative << 31}, plus one (begin, end) entry per band.  The record
//                      buffer of a group is sized to stay L2-resident.
//   bands              are row-cyclic (band b owns rows b, b + nb, ...): dense image regions are dealt out over all bands.
//   sweep  (k_sweep)   persistent CTAs take (sample, band) tasks from an atomic counter and own the band's cells for a
//                      window of up to 6 voxel planes in shared memory: int32 Q24 planes + one count word per cell.  A
//                      warp takes the band's run of one chunk at a time (the next run's rows already in flight); per
//                      record: one returning ATOMS on the count word (n_pos | n_neg << 16) and two fire-and-forget
//                      ATOMS.ADD on the planes k and k + 1 (p * (2^24 - r), p * r).  The flush converts the planes to
//                      fp32 (one rounding from the exact integer), adds the fused voxel.sum(0) plane and the polarity
//                      count frame, and writes every output element exactly once with 16-byte streaming stores.
//   Routes run on the caller's stream and sweeps on a side stream, so the route of group g + 1 fills the SMs the sweep
//   of group g leaves idle.
//
// Exactness: the count word's ATOMS returns how many records the cell had before; the first 127 contributions of a cell
// go to its int32 plane words (127 * 2^24 < 2^31: they never wrap), later ones to a small 64-bit side table keyed by
// (cell, plane), so dense cells (hot pixels, strong edges) cost a few slow adds, not a second pass.  If the side table
// overflows the window is redone with each plane split into two int32 words (sums of w >> 12 and w & 0xfff: exact for
// any count).  Count fields are 16 bits; a band whose counts do not add up to the records it consumed has wrapped a
// field and is reported through bad_count bit 31.  Integer accumulation => order independent, bit-reproducible, and
// identical to the global-RED path.
#include <stdio.h>

#include "ep_binning_common.cuh"

namespace ep {
namespace {

constexpr int kMaxBands = 255;
constexpr uint32_t kKShift = 25;            // record value: r (25 bits) | k << 25 (5 bits) | negative << 31
constexpr uint32_t kKCountOnly = 31;        // interval code of events outside the time bins (count frame only)
constexpr int kAdmit = 127;                 // records per cell and window the int32 planes hold exactly
constexpr int kMaxWindowPlanes = 6;
constexpr int kBufSets = 2;                 // record buffers: the route of group g + 1 runs while group g is swept
constexpr int kSweepThreads = 512;
constexpr int kMaxTableChunks = 1024;       // chunk-table entries of one sample staged in shared memory at a time
constexpr int kSweepSmemBudget = 232448 - 7168;   // opt-in shared memory per CTA minus static use (side table, flags)

enum { kKindTicks64 = 0, kKindF64 = 1, kKindCompact = 2 };

// Optional per-phase cycle accounting (build with EP_PHASE_TIMING=1; tools only, never in the shipped library):
// thread 0 of every CTA adds the cycles between phase boundaries to g.dbg[slot].
#ifdef EP_PHASE_TIMING
#define EP_TICK_INIT() long long _t_last = clock64()
#define EP_TICK(slot)                                                                      \
    do {                                                                                   \
        if (threadIdx.x == 0) {                                                            \
            const long long _now = clock64();                                              \
            atomicAdd(g.dbg + (slot), (unsigned long long)(_now - _t_last));               \
            _t_last = _now;                                                                \
        }                                                                                  \
    } while (0)
#else
#define EP_TICK_INIT() do { } while (0)
#define EP_TICK(slot) do { } while (0)
#endif

struct __align__(16) ChunkInfo {
    int64_t c_lo;            // first event slot of the chunk (multiple of 4; may precede the sample's first event)
    int32_t b;               // owning sample
    uint16_t lo_off, hi_off; // the sample's events inside the chunk: [c_lo + lo_off, c_lo + hi_off)
    // copy of the owning sample's integer-time constants (SampleMeta), so the steady-state path needs this one load
    int64_t t0_ticks;
    uint32_t tmul, tshift;
    uint32_t thalf, flags;
    uint32_t pad[2];
};

struct RouteSrc {
    const uint16_t* x;
    const uint16_t* y;
    const void* t;           // int64 ticks | fp64 | uint32 (relative ticks | p << 31)
    const uint8_t* p;
};

struct BandArgs {
    BinArgs bin;                 // offsets, meta, geometry, bad_count (g0/g1 describe the group)
    // Bands are row-cyclic: band b owns the rows b, b + nb, b + 2 nb, ... (dense image regions — edges, blobs — are
    // dealt out over all bands, so the sweep tasks of a sample carry about the same number of records whatever the
    // spatial distribution).  A record addresses its cell inside the band: (row / nb) * W + column.
    int nb;                      // bands (>= 2)
    int cpb;                     // cell slots per band: ceil(H / nb) * W rounded up to a multiple of 4
    uint32_t w_magic, w_shift;   // n / W  == umulhi(n, w_magic) >> w_shift    for n < 2^31  (W >= 2)
    uint32_t nb_magic, nb_shift; // n / nb == umulhi(n, nb_magic) >> nb_shift
    int P;                       // voxel planes per window
    int chunk;                   // events per routed chunk
    const int32_t* chunk_first;  // [B+1] first chunk id of each sample (global numbering)
    const ChunkInfo* info;       // [total chunks]
    int chunk_begin;             // first chunk id of this group
    int gchunks;                 // row length of runs
    uint32_t* runs;              // [nb][gchunks]  begin | end << 16 of the band's records inside the chunk
    uint16_t* kk;                // [gchunks]      kmin | kmax << 8 of the chunk's in-bins records (255 | 0 when none)
    uint2* rec;                  // [gchunks][chunk]
    float* out_voxel; float* out_sum; float* out_count;
    unsigned int* task_counter;  // sweep: next task of this launch (zeroed before the launch)
    unsigned long long* dbg;     // phase cycle counters (EP_PHASE_TIMING builds), else unused
};

__device__ __forceinline__ int64_t sample_chunk_origin(const BinArgs& a, int b) {
    return off_at(a, b) / kEvPerThread * kEvPerThread;   // chunks start on the 4-event grid of the arrays
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---- chunk numbering: chunks of sample b = ceil((off[b+1] - align4(off[b])) / chunk) --------------------
__global__ void __launch_bounds__(1024) k_chunk_prefix(BinArgs a, int B, int chunk, int32_t* __restrict__ chunk_first) {
    __shared__ int s_warp[32];
    __shared__ int s_base, s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_base = 0; chunk_first[0] = 0; }
    __syncthreads();
    for (int st = 0; st < B; st += blockDim.x) {
        const int b = st + threadIdx.x;
        int v = 0;
        if (b < B) {
            const int64_t lo = sample_chunk_origin(a, b), hi = off_at(a, b + 1);
            v = hi > off_at(a, b) ? (int)ceil_div64(hi - lo, chunk) : 0;
        }
        const int incl = warp_incl_scan(v, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            const int ws = warp_incl_scan(w, lane);
            s_warp[lane] = ws - w;
            if (lane == 31) s_total = ws;
        }
        __syncthreads();
        if (b < B) chunk_first[b + 1] = s_base + s_warp[warp] + incl;
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_chunk_info(BinArgs a, int B, int chunk, const int32_t* __restrict__ chunk_first,
                                                    int n_chunks, ChunkInfo* __restrict__ info) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    int lo = 0, hi = B;                   // smallest j in (0, B] with chunk_first[j] > c; owner = j - 1 (never empty)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_first[mid] > c) hi = mid; else lo = mid;
    }
    const int b = lo;
    const int64_t s_lo = off_at(a, b), s_hi = off_at(a, b + 1);
    const int64_t c_lo = sample_chunk_origin(a, b) + (int64_t)(c - chunk_first[b]) * chunk;
    ChunkInfo ci;
    ci.c_lo = c_lo;
    ci.b = b;
    ci.lo_off = (uint16_t)(c_lo < s_lo ? s_lo - c_lo : 0);
    ci.hi_off = (uint16_t)(c_lo + chunk < s_hi ? chunk : s_hi - c_lo);
    const SampleMeta m = a.meta[b];
    ci.t0_ticks = m.t0_ticks; ci.tmul = m.tmul; ci.tshift = m.tshift; ci.thalf = m.thalf; ci.flags = m.flags;
    ci.pad[0] = ci.pad[1] = 0;
    info[c] = ci;
}

// ---- shared-memory access through 32-bit window addresses (computed once per kernel; keeps the per-event code free
// of address-space conversions) -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
    uint32_t r;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}

// ---- route ------------------------------------------------------------------------------------------------
struct RouteConst {
    uint32_t W, HW;
    uint32_t tmul, tshift, thalf, v_end, v_last;
    uint32_t w_magic, w_shift, nb, nb_magic, nb_shift, drop_band;
    int64_t t0_ticks;
    const SampleMeta* meta;      // the owning sample's metadata: fp64 time constants, read on the non-lean paths only
    int num_bins, count;
    bool int_time, scaled;
};

// flat cell index -> (band, cell inside the band) of the row-cyclic banding
__device__ __forceinline__ uint32_t band_of(const RouteConst& rc, uint32_t flat, uint32_t& local) {
    const uint32_t row = __umulhi(flat, rc.w_magic) >> rc.w_shift, col = flat - row * rc.W;
    const uint32_t q = __umulhi(row, rc.nb_magic) >> rc.nb_shift;
    local = q * rc.W + col;
    return row - q * rc.nb;
}

// One event -> record words, branch-free on the steady-state path.  v = rn(ts * 2^24) (interval k = v >> 24, right
// weight r = v & 0xffffff); the record is {flat cell index, r | k << 25 | negative << 31}, with an event exactly on the last
// node filed under the interval before it (k - 1, r = 2^24), which is the same integer contribution.  Returns the band
// the record goes to (rc.drop_band for slots that hold no valid event of this sample); `bad` is set for events the
// reference raises on.
// LEAN: integer-tick time arithmetic and unscaled coordinates (the uniform branches are hoisted to the CTA level).
template <int KIND, bool LEAN, bool TRACK>
__device__ __forceinline__ uint32_t route_event(const RouteConst& rc, const BinArgs& a, uint32_t xs, uint32_t ys, uint32_t pb, uint32_t tlo,
                                                uint32_t thi, bool live, uint32_t& flat, uint32_t& val, uint32_t& vmin,
                                                uint32_t& vmax, bool& bad) {
    if (!LEAN && rc.scaled) {   // events_reshape fused: fp64 multiply, truncation (events_augment.py:22-26)
        const int64_t xi = __double2ll_rz(__dmul_rn((double)xs, a.sx)), yi = __double2ll_rz(__dmul_rn((double)ys, a.sy));
        const int64_t f = xi + yi * (int64_t)rc.W;
        flat = (f < 0 || f >= (int64_t)rc.HW) ? 0xffffffffu : (uint32_t)f;
    } else {
        flat = ys * rc.W + xs;
    }
    const bool valid = live && flat < rc.HW && pb <= 1u;     // `live` is false for slots of a neighbouring sample (edge chunks)
    bad = live && !valid;
    uint32_t v = 0;
    bool in_bins;
    if (KIND != kKindF64 && (LEAN || rc.int_time)) {
        const int64_t t = (KIND == kKindCompact) ? (int64_t)tlo : (int64_t)(((uint64_t)thi << 32) | tlo);
        in_bins = ticks_to_v(t - rc.t0_ticks, rc.tmul, rc.tshift, rc.thalf, rc.v_end, v);
    } else {
        double dt;
        if (KIND == kKindF64) dt = __longlong_as_double((long long)(((uint64_t)thi << 32) | tlo)) - rc.meta->t0_raw;
        else if (KIND == kKindCompact) dt = (double)((int64_t)tlo - rc.t0_ticks);
        else dt = (double)((int64_t)(((uint64_t)thi << 32) | tlo) - rc.t0_ticks);
        const double ts = dt * rc.meta->scale_raw, tis = floor(ts);
        in_bins = tis >= 0.0 && tis < (double)rc.num_bins;
        if (in_bins)   // same fp32 fraction as quantise_ts; a fraction that rounds to 1.0 carries into k
            v = ((uint32_t)(int)tis << kQ) + (uint32_t)__float2int_rn((float)(ts - tis) * 16777216.0f);
    }
    if (TRACK && in_bins && valid) { vmin = min(vmin, v); vmax = max(vmax, v); }
    uint32_t kr = v + (v & 0xff000000u);                       // r | k << 25
    if (v == rc.v_last) kr -= 1u << kQ;                        // (k, 0) on the last node -> (k - 1, 2^24)
    if (!in_bins) kr = kKCountOnly << kKShift;
    val = kr | ((pb ^ 1u) << 31);                          // bit 31 set = negative polarity
    uint32_t local;
    const uint32_t band = band_of(rc, valid ? flat : 0u, local);
    flat = local;                                              // the record carries the cell inside its band
    return valid ? band : rc.drop_band;      // events outside the time bins keep a slot (count-only record), like in pass A
}

// Two passes over the chunk.  Small CTAs that keep nothing in registers across a barrier: pass A reads x, y, p and
// histograms the chunk by band (fire-and-forget RED), a scan turns the histogram into cursors, pass B reads the events
// again (x, y, p from L2/L1; the stamps were prefetched into L2 at CTA start), builds the records and claims their slots
// with one returning ATOMS on the band's cursor.  Low register count => 5-6 CTAs per SM whose phases drift apart and
// hide each other's barriers and load latencies.
template <int KIND>
__device__ __forceinline__ void load_quad_xyp(const RouteSrc& src, const BinArgs& a, int64_t i0, bool guard, uint2& xv, uint2& yv,
                                              uint32_t& pv) {
    if (!guard || i0 + 4 <= a.n_total) {
        xv = __ldg(reinterpret_cast<const uint2*>(src.x + i0));
        yv = __ldg(reinterpret_cast<const uint2*>(src.y + i0));
        pv = (KIND == kKindCompact) ? 0u : __ldg(reinterpret_cast<const uint32_t*>(src.p + i0));
    } else {
        uint32_t xs_[4] = {0, 0, 0, 0}, ys_[4] = {0, 0, 0, 0};
        pv = 0;
        for (int j = 0; j < 4; ++j)
            if (i0 + j < a.n_total) {
                xs_[j] = src.x[i0 + j]; ys_[j] = src.y[i0 + j];
                if (KIND != kKindCompact) pv |= (uint32_t)src.p[i0 + j] << (8 * j);
            }
        xv = make_uint2(xs_[0] | (xs_[1] << 16), xs_[2] | (xs_[3] << 16));
        yv = make_uint2(ys_[0] | (ys_[1] << 16), ys_[2] | (ys_[3] << 16));
    }
}

template <int KIND>
__device__ __forceinline__ void load_quad_t(const RouteSrc& src, const BinArgs& a, int64_t i0, bool guard, uint4& ta, uint4& tb) {
    if (!guard || i0 + 4 <= a.n_total) {
        if (KIND == kKindCompact) {
            ta = ld_stream(reinterpret_cast<const uint4*>(static_cast<const uint32_t*>(src.t) + i0));
        } else {
            ta = ld_stream(reinterpret_cast<const uint4*>(static_cast<const int64_t*>(src.t) + i0));
            tb = ld_stream(reinterpret_cast<const uint4*>(static_cast<const int64_t*>(src.t) + i0 + 2));
        }
    } else {
        uint32_t lo_[4] = {0, 0, 0, 0}, hi_[4] = {0, 0, 0, 0};
        for (int j = 0; j < 4; ++j)
            if (i0 + j < a.n_total) {
                if (KIND == kKindCompact) {
                    lo_[j] = static_cast<const uint32_t*>(src.t)[i0 + j];
                } else {
                    const uint64_t tv = (uint64_t)static_cast<const int64_t*>(src.t)[i0 + j];
                    lo_[j] = (uint32_t)tv; hi_[j] = (uint32_t)(tv >> 32);
                }
            }
        if (KIND == kKindCompact) ta = make_uint4(lo_[0], lo_[1], lo_[2], lo_[3]);
        else { ta = make_uint4(lo_[0], hi_[0], lo_[1], hi_[1]); tb = make_uint4(lo_[2], hi_[2], lo_[3], hi_[3]); }
    }
}

// pass A of one quad: band histogram
template <int KIND, int THREADS, bool EDGE, bool LEAN>
__device__ __forceinline__ void route2_count(const RouteConst& rc, const BinArgs& a, int h, uint32_t lo_off, uint32_t hi_off,
                                             uint2 xv, uint2 yv, uint32_t pv, uint32_t cnt_addr) {
    const uint32_t xs[4] = {xv.x & 0xffffu, xv.x >> 16, xv.y & 0xffffu, xv.y >> 16};
    const uint32_t ys[4] = {yv.x & 0xffffu, yv.x >> 16, yv.y & 0xffffu, yv.y >> 16};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t flat;
        if (!LEAN && rc.scaled) {
            const int64_t xi = __double2ll_rz(__dmul_rn((double)xs[j], a.sx)), yi = __double2ll_rz(__dmul_rn((double)ys[j], a.sy));
            const int64_t f = xi + yi * (int64_t)rc.W;
            flat = (f < 0 || f >= (int64_t)rc.HW) ? 0xffffffffu : (uint32_t)f;
        } else {
            flat = ys[j] * rc.W + xs[j];
        }
        bool keep = flat < rc.HW && ((pv >> (8 * j)) & 0xffu) <= 1u;
        if (EDGE) {
            const uint32_t e_off = (uint32_t)((h * THREADS + threadIdx.x) * 4 + j);
            keep = keep && (e_off - lo_off < hi_off - lo_off);
        }
        uint32_t local;
        const uint32_t bd = band_of(rc, keep ? flat : 0u, local);
        const uint32_t band = keep ? bd : rc.drop_band;
        reds_add(cnt_addr + band * 4u, 1u);
    }
}

// pass B of one quad: records into their slots
template <int KIND, int THREADS, bool EDGE, bool LEAN, bool TRACK>
__device__ __forceinline__ void route2_place(const RouteConst& rc, const BinArgs& a, int h, uint32_t lo_off, uint32_t hi_off,
                                             uint2 xv, uint2 yv, uint32_t pv, uint4 ta, uint4 tb, uint32_t cur_addr,
                                             uint32_t rec_addr, uint32_t& vmin, uint32_t& vmax, unsigned& nbad) {
    const uint32_t xs[4] = {xv.x & 0xffffu, xv.x >> 16, xv.y & 0xffffu, xv.y >> 16};
    const uint32_t ys[4] = {yv.x & 0xffffu, yv.x >> 16, yv.y & 0xffffu, yv.y >> 16};
    uint32_t tlo[4], thi[4], pb[4];
    if (KIND == kKindCompact) {
        const uint32_t raw[4] = {ta.x, ta.y, ta.z, ta.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { tlo[j] = raw[j] & 0x7fffffffu; thi[j] = 0; pb[j] = raw[j] >> 31; }
    } else {
        tlo[0] = ta.x; thi[0] = ta.y; tlo[1] = ta.z; thi[1] = ta.w;
        tlo[2] = tb.x; thi[2] = tb.y; tlo[3] = tb.z; thi[3] = tb.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = (pv >> (8 * j)) & 0xffu;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bool live = true, bad;
        if (EDGE) {
            const uint32_t e_off = (uint32_t)((h * THREADS + threadIdx.x) * 4 + j);
            live = e_off - lo_off < hi_off - lo_off;
        }
        uint32_t cell, val;
        const uint32_t band = route_event<KIND, LEAN, TRACK>(rc, a, xs[j], ys[j], pb[j], tlo[j], thi[j], live, cell, val, vmin, vmax, bad);
        nbad += bad ? 1u : 0u;
        const uint32_t dst = atoms_add(cur_addr + band * 4u, 1u);
        sts_v2(rec_addr + dst * 8u, cell, val);
    }
}

template <int KIND, int CHUNK, int THREADS, int MINB, bool TRACK>
__global__ void __launch_bounds__(THREADS, MINB) k_route(RouteSrc src, BandArgs g) {
    constexpr int QUADS = CHUNK / (THREADS * 4);
    constexpr int kWarps = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_cnt[kMaxBands + 1], s_cur[kMaxBands + 1];   // per band, plus the drop band at index nb
    __shared__ int s_kmin, s_kmax, s_total;
    const BinArgs& a = g.bin;
    const int chunk = (int)blockIdx.x;
    const ChunkInfo ci = g.info[g.chunk_begin + chunk];
    for (int i = threadIdx.x; i <= g.nb; i += THREADS) s_cnt[i] = 0;
    if (threadIdx.x == 0) { s_kmin = TRACK ? 255 : 0; s_kmax = TRACK ? 0 : 31; }
    const uint32_t lo_off = ci.lo_off, hi_off = ci.hi_off;
    const bool interior = lo_off == 0 && hi_off == CHUNK;
    // the chunk's stamps -> L2 while pass A runs (one 128-byte line per thread and step)
    {
        const int t_bytes = (KIND == kKindCompact) ? 4 : 8;
        const char* tp = static_cast<const char*>(src.t) + ci.c_lo * t_bytes;
        const int64_t lim = (a.n_total - ci.c_lo) * t_bytes;
        for (int64_t o = (int64_t)threadIdx.x * 128; o < (int64_t)CHUNK * t_bytes && o < lim; o += (int64_t)THREADS * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + o));
    }
    RouteConst rc;
    rc.W = (uint32_t)a.W; rc.HW = (uint32_t)(a.H * a.W);
    rc.tmul = ci.tmul; rc.tshift = ci.tshift; rc.thalf = ci.thalf;
    rc.v_end = (uint32_t)a.num_bins << kQ;
    rc.v_last = a.num_bins >= 2 ? (uint32_t)(a.num_bins - 1) << kQ : 0xffffffffu;
    rc.w_magic = g.w_magic; rc.w_shift = g.w_shift; rc.nb = (uint32_t)g.nb; rc.nb_magic = g.nb_magic; rc.nb_shift = g.nb_shift;
    rc.drop_band = (uint32_t)g.nb;
    rc.t0_ticks = ci.t0_ticks;
    rc.meta = a.meta + ci.b;
    rc.num_bins = a.num_bins; rc.count = a.count_channels;
    rc.int_time = (ci.flags & kFlagIntTime) != 0 && a.num_bins > 0;
    rc.scaled = a.scaled != 0;
    const bool lean = KIND != kKindF64 && rc.int_time && !rc.scaled;
    const uint32_t cnt_addr = smem_u32(s_cnt), cur_addr = smem_u32(s_cur), rec_addr = smem_u32(smem_raw);

    // ---- pass A
    {
        uint2 xv[QUADS], yv[QUADS];
        uint32_t pv[QUADS];
#pragma unroll
        for (int h = 0; h < QUADS; ++h)
            load_quad_xyp<KIND>(src, a, ci.c_lo + (int64_t)(h * THREADS + threadIdx.x) * 4, !interior, xv[h], yv[h], pv[h]);
        __syncthreads();                                           // histogram zeroed
#pragma unroll
        for (int h = 0; h < QUADS; ++h) {
            if (interior && lean) route2_count<KIND, THREADS, false, true>(rc, a, h, lo_off, hi_off, xv[h], yv[h], pv[h], cnt_addr);
            else if (interior) route2_count<KIND, THREADS, false, false>(rc, a, h, lo_off, hi_off, xv[h], yv[h], pv[h], cnt_addr);
            else route2_count<KIND, THREADS, true, false>(rc, a, h, lo_off, hi_off, xv[h], yv[h], pv[h], cnt_addr);
        }
    }
    __syncthreads();
    {
        // exclusive scan of the band counts, eight per lane, computed by every warp; warp w publishes the bands b with
        // b % kWarps == w (cursor + run entry)
        const int l = threadIdx.x & 31, wp = threadIdx.x >> 5;
        int c[8], tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i] = (8 * l + i <= g.nb) ? s_cnt[8 * l + i] : 0; tot += c[i]; }
        int run = warp_incl_scan(tot, l) - tot;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int bd = 8 * l + i;
            if (bd <= g.nb && (bd % kWarps) == wp) {
                s_cur[bd] = run;
                if (bd < g.nb) g.runs[(int64_t)bd * g.gchunks + chunk] = (uint32_t)run | ((uint32_t)(run + c[i]) << 16);
                else s_total = run;
            }
            run += c[i];
        }
    }
    __syncthreads();
    // ---- pass B
    uint32_t vmin = 0xffffffffu, vmax = 0;
    unsigned nbad = 0;
#pragma unroll 1
    for (int h0 = 0; h0 < QUADS; h0 += 2) {
        uint2 xv[2], yv[2];
        uint32_t pv[2];
        uint4 ta[2], tb[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (h0 + u < QUADS) {
                const int64_t i0 = ci.c_lo + (int64_t)((h0 + u) * THREADS + threadIdx.x) * 4;
                load_quad_xyp<KIND>(src, a, i0, !interior, xv[u], yv[u], pv[u]);
                load_quad_t<KIND>(src, a, i0, !interior, ta[u], tb[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (h0 + u < QUADS) {
                const int h = h0 + u;
                if (interior && lean) route2_place<KIND, THREADS, false, true, TRACK>(rc, a, h, lo_off, hi_off, xv[u], yv[u], pv[u], ta[u], tb[u], cur_addr, rec_addr, vmin, vmax, nbad);
                else if (interior) route2_place<KIND, THREADS, false, false, TRACK>(rc, a, h, lo_off, hi_off, xv[u], yv[u], pv[u], ta[u], tb[u], cur_addr, rec_addr, vmin, vmax, nbad);
                else route2_place<KIND, THREADS, true, false, TRACK>(rc, a, h, lo_off, hi_off, xv[u], yv[u], pv[u], ta[u], tb[u], cur_addr, rec_addr, vmin, vmax, nbad);
            }
        }
    }
    if (TRACK) {
        vmin = warp_reduce(vmin, [](uint32_t x, uint32_t y) { return min(x, y); });
        vmax = warp_reduce(vmax, [](uint32_t x, uint32_t y) { return max(x, y); });
        if ((threadIdx.x & 31) == 0 && vmin <= vmax) { atomicMin(&s_kmin, (int)(vmin >> kQ)); atomicMax(&s_kmax, (int)(vmax >> kQ)); }
    }
    if (nbad && a.bad_count) atomicAdd(a.bad_count, nbad);
    __syncthreads();
    if (threadIdx.x == 32) g.kk[chunk] = (uint16_t)((s_kmin & 0xff) | (s_kmax << 8));
    const int n_vec = (s_total + 1) >> 1;
    uint4* gv = reinterpret_cast<uint4*>(g.rec + (int64_t)chunk * CHUNK);
    for (int i = threadIdx.x; i < n_vec; i += THREADS) gv[i] = reinterpret_cast<const uint4*>(smem_raw)[i];
}

template <int KIND, int CHUNK, int THREADS, int MINB, bool TRACK>
cudaError_t launch_route(cudaStream_t st, const RouteSrc& src, const BandArgs& g, unsigned grid) {
    static bool configured = false;
    const size_t smem = (size_t)CHUNK * 8;
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(k_route<KIND, CHUNK, THREADS, MINB, TRACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
        cudaFuncSetAttribute(k_route<KIND, CHUNK, THREADS, MINB, TRACK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    k_route<KIND, CHUNK, THREADS, MINB, TRACK><<<grid, THREADS, smem, st>>>(src, g);
    return cudaSuccess;
}

// ---- sweep ------------------------------------------------------------------------------------------------
struct Window { int q0, q1; };   // voxel planes [q0, q1) held in shared memory

// Rare passes of a task, written for clarity rather than speed (the steady state is accumulate_fast below):
// mode 1: hi/lo split planes (exact for any count); mode 2: count word only (count frame of a multi-window task).
// (mode 0, int32 planes + count word without the dense-cell side table, is kept for reference and unused.)
template <int MODE, int THREADS>
__device__ __forceinline__ unsigned accumulate(const BandArgs& g, Window w, bool count_all, bool filter, int band,
                                               int64_t band_base, int ch0, int nch, uint32_t* s_run, uint16_t* s_kk,
                                               uint32_t* slots) {
    constexpr int kWarps = THREADS / 32;
    constexpr int U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cpb = (uint32_t)g.cpb;
    // slot addresses as functions of the record's flat cell index: count word [cpb] first, then the planes
    const uint32_t n_addr0 = smem_u32(slots) - (uint32_t)band_base * 4u;
    const uint32_t p_addr0 = smem_u32(slots + cpb) - ((uint32_t)band_base + (uint32_t)w.q0 * cpb) * 4u;   // int32 [P][cpb]
    const uint32_t w_addr0 = smem_u32(slots) - ((uint32_t)band_base + 2u * (uint32_t)w.q0 * cpb) * 4u;    // hi/lo int32 [Pw][2][cpb]
    unsigned matched = 0;
    for (int t0 = 0; t0 < nch; t0 += kMaxTableChunks) {
        const int nt = min(kMaxTableChunks, nch - t0);
        __syncthreads();                                    // previous tile of the table consumed / slots ready
        for (int c = threadIdx.x; c < nt; c += THREADS) {
            s_run[c] = g.runs[(int64_t)band * g.gchunks + ch0 + t0 + c];
            s_kk[c] = g.kk[ch0 + t0 + c];
        }
        __syncthreads();
        for (int c = warp; c < nt; c += kWarps) {
            if (filter) {      // chunks are time ordered: most lie outside a window of a multi-window task
                const int kmin = s_kk[c] & 0xff, kmax = s_kk[c] >> 8;
                if (kmax + 1 < w.q0 || kmin >= w.q1) continue;
            }
            const uint32_t run = s_run[c];
            const int lo = run & 0xffff, hi = run >> 16;
            const uint2* rp = g.rec + (int64_t)(ch0 + t0 + c) * g.chunk;
            for (int i0 = lo + lane; i0 < hi; i0 += 32 * U) {
                uint2 rv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = i0 + 32 * u;
                    if (i < hi) rv[u] = __ldcs(rp + i);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (i0 + 32 * u >= hi) continue;
                    const uint32_t flat = rv[u].x, v = rv[u].y;
                    const uint32_t k = (v >> kKShift) & 31u;
                    const bool pos = (v >> 31) == 0u;
                    const uint32_t r = v & 0x1ffffffu;
                    const bool left_ok = k >= (uint32_t)w.q0 && k < (uint32_t)w.q1;
                    const bool right_ok = k + 1 >= (uint32_t)w.q0 && k + 1 < (uint32_t)w.q1;
                    if (MODE == 2 || count_all || left_ok || right_ok) {
                        ++matched;
                        if (MODE != 1) reds_add(n_addr0 + flat * 4u, pos ? 1u : 0x10000u);
                    }
                    const uint32_t idx = k * cpb + flat;
                    if (MODE == 0) {
                        const uint32_t wl = (1u << kQ) - r;
                        const uint32_t addr = p_addr0 + idx * 4u;
                        if (left_ok) reds_add(addr, pos ? wl : 0u - wl);
                        if (right_ok) reds_add(addr + cpb * 4u, pos ? r : 0u - r);
                    } else if (MODE == 1) {
                        // plane j of the window as two int32 words per cell: [2j] sum of p * (w >> 12), [2j+1] sum of
                        // p * (w & 0xfff); w <= 2^24, so both are exact for more events than a count field can hold
                        const uint32_t wl = (1u << kQ) - r, m = pos ? 0u : ~0u;
                        const uint32_t addr = w_addr0 + (2u * idx - flat) * 4u;     // (2 * (k * cpb) + flat) words
                        if (left_ok) {
                            reds_add(addr, ((wl >> 12) ^ m) - m);
                            reds_add(addr + cpb * 4u, ((wl & 0xfffu) ^ m) - m);
                        }
                        if (right_ok) {
                            reds_add(addr + cpb * 8u, ((r >> 12) ^ m) - m);
                            reds_add(addr + cpb * 12u, ((r & 0xfffu) ^ m) - m);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    return matched;
}

// Steady-state accumulate of window w with int32 planes + count word (same result as accumulate<0>): one branch per
// record, every address a multiply-add from the record's flat cell index, the next chunk's rows already in flight
// while the current ones are applied.
// Side table for cells that receive more than kAdmit records in a window: the first kAdmit contributions of a cell go
// to its int32 plane words (which therefore never wrap), every later one is added here in 64 bits, keyed by
// (cell, plane).  Which records are "later" is decided by the value the count word's ATOMS returns, so it is exact
// under any interleaving.  Only dense cells (hot pixels, strong edges) ever come here.
constexpr int kSpill = 512;                 // slots (power of two)
struct SpillTable {
    unsigned int key[kSpill];               // cell * 8 + plane + 1, 0 = empty
    unsigned long long acc[kSpill];
    int state;                              // 0 clean, 1 in use, 2 overflowed (the window is redone exactly)
};

__device__ __noinline__ void spill_add(SpillTable* t, uint32_t cell, uint32_t plane, int32_t v) {
    const unsigned int key = cell * 8u + plane + 1u;
    unsigned int h = (key * 2654435761u) >> (32 - 9);
    if (t->state == 0) atomicMax(&t->state, 1);
    for (int probe = 0; probe < kSpill; ++probe) {
        const unsigned int prev = atomicCAS(&t->key[h], 0u, key);
        if (prev == 0u || prev == key) { atomicAdd(&t->acc[h], (unsigned long long)(long long)v); return; }
        h = (h + 1) & (kSpill - 1);
    }
    atomicMax(&t->state, 2);
}

__device__ __noinline__ long long spill_get(const SpillTable* t, uint32_t cell, uint32_t plane) {
    const unsigned int key = cell * 8u + plane + 1u;
    unsigned int h = (key * 2654435761u) >> (32 - 9);
    for (int probe = 0; probe < kSpill; ++probe) {
        const unsigned int k = t->key[h];
        if (k == key) return (long long)t->acc[h];
        if (k == 0u) return 0;
        h = (h + 1) & (kSpill - 1);
    }
    return 0;
}

// Steady-state accumulate of window w with int32 planes + count word (same result as accumulate<0>): one branch per
// record, every address a multiply-add from the record's cell index, the next chunk's rows already in flight
// while the current ones are applied.
struct FastAddr { uint32_t n0, p0, cpb, cpb4, q0, nfast, q1; SpillTable* spill; };

__device__ __forceinline__ bool admitted(uint32_t old_n) { return (old_n & 0xffffu) + (old_n >> 16) < (uint32_t)kAdmit; }

__device__ __noinline__ void apply_record_slow(uint32_t cell, uint32_t v, FastAddr fa, bool count_all, unsigned& skipped) {
    const uint32_t k = (v >> kKShift) & 31u, r = v & 0x1ffffffu, m = (uint32_t)((int32_t)v >> 31);
    const bool left_ok = k >= fa.q0 && k < fa.q1;
    const bool right_ok = k + 1 >= fa.q0 && k + 1 < fa.q1;
    if (!(count_all || left_ok || right_ok)) { ++skipped; return; }
    const uint32_t old_n = atoms_add(fa.n0 + cell * 4u, (m & 0xffffu) + 1u);
    const uint32_t addr = fa.p0 + ((k - fa.q0) * fa.cpb + cell) * 4u, wl = (1u << kQ) - r;
    if (admitted(old_n)) {
        if (left_ok) reds_add(addr, (wl ^ m) - m);
        if (right_ok) reds_add(addr + fa.cpb4, (r ^ m) - m);
    } else {
        if (left_ok) spill_add(fa.spill, cell, k - fa.q0, (int32_t)((wl ^ m) - m));
        if (right_ok) spill_add(fa.spill, cell, k + 1 - fa.q0, (int32_t)((r ^ m) - m));
    }
}

__device__ __noinline__ void spill_record(uint32_t cell, uint32_t v, FastAddr fa) {
    const uint32_t kw = ((v >> kKShift) & 31u) - fa.q0, r = v & 0x1ffffffu, m = (uint32_t)((int32_t)v >> 31);
    const uint32_t wl = (1u << kQ) - r;
    spill_add(fa.spill, cell, kw, (int32_t)((wl ^ m) - m));
    spill_add(fa.spill, cell, kw + 1u, (int32_t)((r ^ m) - m));
}

// step 1 of a record: bump the count word (returning ATOMS); records off the fast path are finished right here
__device__ __forceinline__ uint32_t apply_begin(uint32_t cell, uint32_t v, const FastAddr& fa, bool count_all, unsigned& skipped) {
    const uint32_t kw = ((v >> kKShift) & 31u) - fa.q0;
    if (kw < fa.nfast) {                                        // both planes k and k + 1 lie inside the window
        const uint32_t m = (uint32_t)((int32_t)v >> 31);        // 0 (positive) | ~0 (negative)
        return atoms_add(fa.n0 + cell * 4u, (m & 0xffffu) + 1u);
    }
    apply_record_slow(cell, v, fa, count_all, skipped);
    return 0xffffffffu;
}

// step 2: the two plane contributions, into the plane words while the cell is within kAdmit records, else the side table
__device__ __forceinline__ void apply_finish(uint32_t cell, uint32_t v, uint32_t old_n, const FastAddr& fa) {
    if (old_n == 0xffffffffu) return;
    if (admitted(old_n)) {
        const uint32_t kw = ((v >> kKShift) & 31u) - fa.q0, r = v & 0x1ffffffu, m = (uint32_t)((int32_t)v >> 31);
        const uint32_t addr = (kw + 1u) * fa.cpb4 + fa.n0 + cell * 4u, wl = (1u << kQ) - r;   // plane j sits (j + 1) * cpb words after the count word
        reds_add(addr, (wl ^ m) - m);
        reds_add(addr + fa.cpb4, (r ^ m) - m);
    } else {
        spill_record(cell, v, fa);
    }
}

template <int THREADS>
__device__ __forceinline__ unsigned accumulate_fast(const BandArgs& g, Window w, bool count_all, bool filter, int band,
                                                    int64_t band_base, int ch0, int nch, uint32_t* s_run, uint16_t* s_kk,
                                                    uint32_t* slots, SpillTable* spill) {
    constexpr int kWarps = THREADS / 32;
    constexpr int U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FastAddr fa;
    fa.spill = spill;
    fa.cpb = (uint32_t)g.cpb; fa.cpb4 = fa.cpb * 4u;
    fa.q0 = (uint32_t)w.q0; fa.q1 = (uint32_t)w.q1;
    fa.nfast = w.q1 - w.q0 >= 2 ? (uint32_t)(w.q1 - w.q0 - 1) : 0u;
    fa.n0 = smem_u32(slots) - (uint32_t)band_base * 4u;               // count word [cpb], then int32 planes [P][cpb]
    fa.p0 = smem_u32(slots + fa.cpb) - (uint32_t)band_base * 4u;
    asm volatile("" : "+r"(fa.n0), "+r"(fa.p0), "+r"(fa.cpb4), "+r"(fa.cpb), "+r"(fa.q0), "+r"(fa.nfast));   // keep them in registers (no rematerialisation)
    unsigned taken = 0, skipped = 0;
    for (int t0 = 0; t0 < nch; t0 += kMaxTableChunks) {
        const int nt = min(kMaxTableChunks, nch - t0);
        __syncthreads();                                    // previous tile of the table consumed / slots ready
        for (int c = threadIdx.x; c < nt; c += THREADS) {
            s_run[c] = g.runs[(int64_t)band * g.gchunks + ch0 + t0 + c];
            s_kk[c] = g.kk[ch0 + t0 + c];
        }
        __syncthreads();
        const uint2* rec0 = g.rec + (int64_t)(ch0 + t0) * g.chunk;
        // chunks are time ordered: most lie outside a window of a multi-window task
        auto next_chunk = [&](int c) {
            if (filter) {
                while (c < nt) {
                    const int kmin = s_kk[c] & 0xff, kmax = s_kk[c] >> 8;
                    if (!(kmax + 1 < w.q0 || kmin >= w.q1)) break;
                    c += kWarps;
                }
            }
            return c;
        };
        auto load_rows = [&](int c, int& lo, int& hi, uint2 (&rv)[U]) {
            const uint32_t run = s_run[c];
            lo = (int)(run & 0xffffu) + lane; hi = (int)(run >> 16);
            const uint2* rp = rec0 + (int64_t)c * g.chunk;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (lo + 32 * u < hi) rv[u] = __ldcs(rp + lo + 32 * u);
        };
        int c = next_chunk(warp), lo = 0, hi = 0;
        uint2 rv[U];
        if (c < nt) load_rows(c, lo, hi, rv);
        while (c < nt) {
            const int cn = next_chunk(c + kWarps);
            int lon = 0, hin = 0;
            uint2 rn[U];
            if (cn < nt) load_rows(cn, lon, hin, rn);
            if (lo < hi) taken += (unsigned)((hi - lo + 31) >> 5);
            uint32_t old_n[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (lo + 32 * u < hi) old_n[u] = apply_begin(rv[u].x, rv[u].y, fa, count_all, skipped);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (lo + 32 * u < hi) apply_finish(rv[u].x, rv[u].y, old_n[u], fa);
            if (lo + 32 * U < hi) {                          // longer run than the rows in flight
                const uint2* rp = rec0 + (int64_t)c * g.chunk;
                for (int i = lo + 32 * U; i < hi; i += 32) {
                    const uint2 q = __ldcs(rp + i);
                    apply_finish(q.x, q.y, apply_begin(q.x, q.y, fa, count_all, skipped), fa);
                }
            }
            c = cn; lo = lon; hi = hin;
#pragma unroll
            for (int u = 0; u < U; ++u) rv[u] = rn[u];
        }
    }
    __syncthreads();
    return taken - skipped;
}

__device__ __forceinline__ float q24_to_float(long long val) {
    // one rounding from the exact fixed-point value; the 32-bit conversion is the same rounding when it applies
    const int lo = (int)val;
    const float f = ((long long)lo == val) ? (float)lo : __ll2float_rn(val);
    return f * (1.0f / 16777216.0f);
}

// cell slot c of a band (c a multiple of VEC inside a row) -> offset in an (H, W) output plane: row-cyclic banding
__device__ __forceinline__ int64_t out_offset(const BandArgs& g, int band, int c) {
    const uint32_t q = __umulhi((uint32_t)c, g.w_magic) >> g.w_shift;
    return (int64_t)(q * (uint32_t)g.nb + (uint32_t)band) * g.bin.W + ((uint32_t)c - q * (uint32_t)g.bin.W);
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
    if (VEC == 4) st_stream(reinterpret_cast<float4*>(p), make_float4(v[0], v[VEC > 1 ? 1 : 0], v[VEC > 2 ? 2 : 0], v[VEC > 3 ? 3 : 0]));
    else st_stream(p, v[0]);
}

template <int VEC>
__device__ __forceinline__ void load_u32(const uint32_t* p, uint32_t (&v)[VEC]) {
    if (VEC == 4) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        v[0] = q.x; v[VEC > 1 ? 1 : 0] = q.y; v[VEC > 2 ? 2 : 0] = q.z; v[VEC > 3 ? 3 : 0] = q.w;
    } else {
        v[0] = *p;
    }
}

template <int VEC>
__device__ __forceinline__ void zero_u32(uint32_t* p) {
    if (VEC == 4) *reinterpret_cast<uint4*>(p) = make_uint4(0u, 0u, 0u, 0u);
    else *p = 0u;
}

// Fast flush of window w (int32 planes): voxel planes, and when the task has a single window also the sum plane and the
// count frame.  Slots are left zeroed.  Accumulates this thread's (hot, counted).
template <int VEC, int THREADS, int NP>
__device__ __forceinline__ void flush_fast_np(const BandArgs& g, int q0, bool single, bool count, int b, int band,
                                              int ncell, uint32_t* slots, const SpillTable* spill, unsigned& counted) {
    const BinArgs& a = g.bin;
    const int64_t HW = (int64_t)a.H * a.W;
    const int cpb = g.cpb;
    const bool spilled = spill->state != 0;
    uint32_t* sN = slots;
    uint32_t* sP = slots + cpb;
    float* ov = g.out_voxel + ((int64_t)b * a.num_bins + q0) * HW;
    float* os = (single && g.out_sum && NP > 0) ? g.out_sum + (int64_t)b * HW : nullptr;
    float* oc = (single && count) ? g.out_count + (int64_t)b * a.count_channels * HW : nullptr;
    const int64_t neg_off = (int64_t)(a.count_channels - 1) * HW;
    for (int c = threadIdx.x * VEC; c < ncell; c += THREADS * VEC) {
        uint32_t n[VEC], pw[NP > 0 ? NP : 1][VEC];
        load_u32<VEC>(sN + c, n);
#pragma unroll
        for (int j = 0; j < NP; ++j) load_u32<VEC>(sP + j * cpb + c, pw[j]);
        zero_u32<VEC>(sN + c);
#pragma unroll
        for (int j = 0; j < NP; ++j) zero_u32<VEC>(sP + j * cpb + c);
        unsigned mx = 0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const unsigned tot = (n[v] & 0xffffu) + (n[v] >> 16);
            mx = max(mx, tot);
            counted += tot;
        }
        const bool dense = spilled && mx > (unsigned)kAdmit;      // some cell here has contributions in the side table
        float sum[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) sum[v] = 0.f;
        const int64_t off = out_offset(g, band, c);
        float* o = ov + off;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            float r[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                r[v] = (float)(int32_t)pw[j][v] * (1.0f / 16777216.0f);
                if (dense && (n[v] & 0xffffu) + (n[v] >> 16) > (unsigned)kAdmit)
                    r[v] = q24_to_float((long long)(int32_t)pw[j][v] + spill_get(spill, (uint32_t)(c + v), (uint32_t)j));
                sum[v] += r[v];                   // voxel.sum(dim=0): sequential fp32 over bins
            }
            store_vec<VEC>(o, r);
            o += HW;
        }
        if (os) store_vec<VEC>(os + off, sum);
        if (oc) {
            float cp[VEC], cn[VEC], z[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { cp[v] = (float)(n[v] & 0xffffu); cn[v] = (float)(n[v] >> 16); z[v] = 0.f; }
            store_vec<VEC>(oc + off, cp);
            store_vec<VEC>(oc + neg_off + off, cn);
            if (a.count_channels == 3) store_vec<VEC>(oc + HW + off, z);
        }
    }
}

template <int VEC, int THREADS>
__device__ __forceinline__ void flush_fast(const BandArgs& g, Window w, bool single, bool count, int b, int band,
                                           int ncell, uint32_t* slots, const SpillTable* spill, unsigned& counted) {
    switch (w.q1 - w.q0) {
        case 0: flush_fast_np<VEC, THREADS, 0>(g, w.q0, single, count, b, band, ncell, slots, spill, counted); break;
        case 1: flush_fast_np<VEC, THREADS, 1>(g, w.q0, single, count, b, band, ncell, slots, spill, counted); break;
        case 2: flush_fast_np<VEC, THREADS, 2>(g, w.q0, single, count, b, band, ncell, slots, spill, counted); break;
        case 3: flush_fast_np<VEC, THREADS, 3>(g, w.q0, single, count, b, band, ncell, slots, spill, counted); break;
        case 4: flush_fast_np<VEC, THREADS, 4>(g, w.q0, single, count, b, band, ncell, slots, spill, counted); break;
        case 5: flush_fast_np<VEC, THREADS, 5>(g, w.q0, single, count, b, band, ncell, slots, spill, counted); break;
        default: flush_fast_np<VEC, THREADS, 6>(g, w.q0, single, count, b, band, ncell, slots, spill, counted); break;
    }
}

// flush of the count word only (dedicated count pass of multi-window tasks)
template <int VEC, int THREADS>
__device__ __forceinline__ void flush_count(const BandArgs& g, int b, int band, int ncell, uint32_t* slots,
                                            unsigned& counted) {
    const BinArgs& a = g.bin;
    const int64_t HW = (int64_t)a.H * a.W;
    for (int c = threadIdx.x * VEC; c < ncell; c += THREADS * VEC) {
        uint32_t n[VEC];
        load_u32<VEC>(slots + c, n);
        zero_u32<VEC>(slots + c);
        float cp[VEC], cn[VEC], z[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            counted += (n[v] & 0xffffu) + (n[v] >> 16);
            cp[v] = (float)(n[v] & 0xffffu); cn[v] = (float)(n[v] >> 16); z[v] = 0.f;
        }
        float* oc = g.out_count + (int64_t)b * a.count_channels * HW + out_offset(g, band, c);
        store_vec<VEC>(oc, cp);
        store_vec<VEC>(oc + (int64_t)(a.count_channels - 1) * HW, cn);
        if (a.count_channels == 3) store_vec<VEC>(oc + HW, z);
    }
}

// exact flush of hi/lo split planes [w.q0, w.q1): voxel planes only
template <int VEC, int THREADS>
__device__ __forceinline__ void flush_wide(const BandArgs& g, Window w, int b, int band, int ncell, uint32_t* slots) {
    const BinArgs& a = g.bin;
    const int64_t HW = (int64_t)a.H * a.W;
    const int cpb = g.cpb, np = w.q1 - w.q0;
    float* ov = g.out_voxel + ((int64_t)b * a.num_bins + w.q0) * HW;
    for (int c = threadIdx.x * VEC; c < ncell; c += THREADS * VEC) {
        const int64_t off = out_offset(g, band, c);
        for (int j = 0; j < np; ++j) {
            uint32_t hi[VEC], lo[VEC];
            load_u32<VEC>(slots + (2 * j) * cpb + c, hi);
            load_u32<VEC>(slots + (2 * j + 1) * cpb + c, lo);
            zero_u32<VEC>(slots + (2 * j) * cpb + c);
            zero_u32<VEC>(slots + (2 * j + 1) * cpb + c);
            float r[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) r[v] = q24_to_float((long long)(int32_t)hi[v] * 4096 + (long long)(int32_t)lo[v]);
            store_vec<VEC>(ov + (int64_t)j * HW + off, r);
        }
    }
}

// voxel.sum(0) from the planes this thread wrote itself (same cell mapping as the flushes)
template <int VEC, int THREADS>
__device__ __forceinline__ void sum_pass(const BandArgs& g, int b, int band, int ncell) {
    const BinArgs& a = g.bin;
    const int64_t HW = (int64_t)a.H * a.W;
    const float* ov = g.out_voxel + (int64_t)b * a.num_bins * HW;
    for (int c = threadIdx.x * VEC; c < ncell; c += THREADS * VEC) {
        const int64_t off = out_offset(g, band, c);
        float sum[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) sum[v] = 0.f;
        for (int q = 0; q < a.num_bins; ++q) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) sum[v] += __ldcg(ov + (int64_t)q * HW + off + v);
        }
        store_vec<VEC>(g.out_sum + (int64_t)b * HW + off, sum);
    }
}

// Persistent over the (sample, band) tasks of a group: grid = min(tasks, SMs), one CTA per SM.
template <int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_sweep(BandArgs g, int n_tasks, int slot_words) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* slots = reinterpret_cast<uint32_t*>(smem_raw);                 // [(P + 1) * cpb] (count word + P planes)
    uint32_t* s_run = slots + slot_words;                                    // [kMaxTableChunks]
    uint16_t* s_kk = reinterpret_cast<uint16_t*>(s_run + kMaxTableChunks);   // [kMaxTableChunks]
    __shared__ unsigned s_check[2];
    __shared__ int s_task;
    __shared__ SpillTable s_spill;
    const BinArgs& a = g.bin;
    const int bins = a.num_bins;
    const bool count = a.count_channels != 0;
    const int n_win = bins > 0 ? (bins + g.P - 1) / g.P : 0;
    const bool single = n_win <= 1;
    const int Pw = (g.P + 1) / 2;                                            // hi/lo split planes the same slots hold

    for (int i = threadIdx.x; i < slot_words; i += THREADS) slots[i] = 0u;
    for (int i = threadIdx.x; i < kSpill; i += THREADS) { s_spill.key[i] = 0u; s_spill.acc[i] = 0ull; }
    if (threadIdx.x < 2) s_check[threadIdx.x] = 0u;
    if (threadIdx.x == 0) s_spill.state = 0;
    EP_TICK_INIT();

    // tasks are handed out dynamically: uneven bands (events concentrated on edges) do not pin the launch to the
    // slowest CTA of a static split
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_task = (int)atomicAdd(g.task_counter, 1u);
        __syncthreads();
        const int task = s_task;
        if (task >= n_tasks) break;
        const int band = task % g.nb;
        const int b = a.g0 + task / g.nb;
        const int64_t band_base = 0;                                      // records address cells inside the band
        const int ncell = band < a.H ? ((a.H - band + g.nb - 1) / g.nb) * a.W : 0;   // rows band, band + nb, ... of the image
        const int ch0 = g.chunk_first[b] - g.chunk_begin, nch = g.chunk_first[b + 1] - g.chunk_first[b];
        bool any_hot = false;

        for (int wi = 0; wi < (n_win > 0 ? n_win : 1); ++wi) {
            Window w;
            w.q0 = wi * g.P;
            w.q1 = min(w.q0 + g.P, bins);
            const bool count_here = single && count;                  // the count word doubles as the count frame
            EP_TICK(5);
            unsigned matched = accumulate_fast<THREADS>(g, w, count_here, !single, band, band_base, ch0, nch, s_run, s_kk, slots, &s_spill);
            unsigned counted = 0;
            EP_TICK(6);
            const int spill_state = s_spill.state;                    // uniform: accumulate ended with a barrier
            flush_fast<VEC, THREADS>(g, w, single, count_here, b, band, ncell, slots, &s_spill, counted);
            matched = warp_reduce(matched, [](unsigned x, unsigned y) { return x + y; });
            counted = warp_reduce(counted, [](unsigned x, unsigned y) { return x + y; });
            if ((threadIdx.x & 31) == 0) {
                if (matched) atomicAdd(&s_check[0], matched);
                if (counted) atomicAdd(&s_check[1], counted);
            }
            __syncthreads();
            const int hot_any = spill_state == 2;                     // the side table overflowed: redo the window exactly
            if (spill_state != 0) {
                for (int i = threadIdx.x; i < kSpill; i += THREADS) { s_spill.key[i] = 0u; s_spill.acc[i] = 0ull; }
                if (threadIdx.x == 0) s_spill.state = 0;
                __syncthreads();
            }
            EP_TICK(7);
            if (threadIdx.x == 0) {
                if (s_check[0] != s_check[1] && a.bad_count) atomicOr(a.bad_count, 0x80000000u);   // a 16-bit count wrapped
                s_check[0] = 0u; s_check[1] = 0u;
            }
            if (hot_any) {
                // more dense (cell, plane) pairs than the side table holds: redo the window's planes with hi/lo split accumulators
                any_hot = true;
                for (int s0 = w.q0; s0 < w.q1; s0 += Pw) {
                    Window ws;
                    ws.q0 = s0;
                    ws.q1 = min(s0 + Pw, w.q1);
                    accumulate<1, THREADS>(g, ws, false, false, band, band_base, ch0, nch, s_run, s_kk, slots);
                    flush_wide<VEC, THREADS>(g, ws, b, band, ncell, slots);
                }
                EP_TICK(8);
#ifdef EP_PHASE_TIMING
                if (threadIdx.x == 0) atomicAdd(g.dbg + 9, 1ull);
#endif
            }
        }
        if (g.out_sum && bins > 0 && (!single || any_hot)) sum_pass<VEC, THREADS>(g, b, band, ncell);
        if (count && !single) {
            Window w;
            w.q0 = 0; w.q1 = 0;
            unsigned matched = accumulate<2, THREADS>(g, w, true, false, band, band_base, ch0, nch, s_run, s_kk, slots);
            unsigned counted = 0;
            flush_count<VEC, THREADS>(g, b, band, ncell, slots, counted);
            matched = warp_reduce(matched, [](unsigned x, unsigned y) { return x + y; });
            counted = warp_reduce(counted, [](unsigned x, unsigned y) { return x + y; });
            if ((threadIdx.x & 31) == 0) {
                if (matched) atomicAdd(&s_check[0], matched);
                if (counted) atomicAdd(&s_check[1], counted);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (s_check[0] != s_check[1] && a.bad_count) atomicOr(a.bad_count, 0x80000000u);
                s_check[0] = 0u; s_check[1] = 0u;
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
struct BandPlan {
    int nb, cpb, P, chunk;
    uint32_t w_magic, w_shift, nb_magic, nb_shift;
    int slot_words;
    size_t sweep_smem;
};


size_t banded_group_budget() {
    long mb = env_int("EP_BANDED_GROUP_MB", 72);
    if (mb < 1) mb = 1;
    return (size_t)mb << 20;
}

// exact n / d for n < 2^31, d >= 2: s = ceil(log2 d), magic = floor(2^(31+s) / d) + 1, q = umulhi(n, magic) >> (s - 1)
void make_div(uint32_t d, uint32_t* magic, uint32_t* shift) {
    int sh = 1;
    while ((1ll << sh) < (int64_t)d) ++sh;
    *magic = (uint32_t)(((1ull << (31 + sh)) / (uint64_t)d) + 1);
    *shift = (uint32_t)(sh - 1);
}

bool plan_bands(const ep_bin_params* p, int batch, int64_t n_events, BandPlan* bp) {
    const int64_t HW = (int64_t)p->height * p->width;
    const int H = p->height, W = p->width;
    if (HW >= (1ll << 31) || W >= 65536 || W < 2 || p->num_bins > 30) return false;
    const int bins = p->num_bins;
    const int n_win = bins > 0 ? (bins + kMaxWindowPlanes - 1) / kMaxWindowPlanes : 1;
    const int P = bins > 0 ? (bins + n_win - 1) / n_win : 0;
    const int words_per_cell = P + 1;
    const int table_bytes = kMaxTableChunks * 6;
    int cpb_max = (kSweepSmemBudget - table_bytes) / (4 * words_per_cell);
    const int cap = env_int("EP_BANDED_CPB", 0);
    if (cap > 0 && cap < cpb_max) cpb_max = cap;
    cpb_max = cpb_max / 4 * 4;
    const int rows_max = cpb_max / W;                 // rows of the image one band can hold
    if (rows_max < 1) return false;
    int nb_min = (H + rows_max - 1) / rows_max;
    if (nb_min < 2) nb_min = 2;
    if (nb_min > kMaxBands) return false;
    int nb = nb_min;
    if ((int64_t)nb * batch < kNumSMs) {
        // small batches: more, smaller bands so that the (sample, band) tasks cover the SMs
        while ((int64_t)nb * batch < kNumSMs && nb < kMaxBands && nb < H && (int64_t)((H + nb) / (nb + 1)) * W >= 1024) ++nb;
    } else {
        // a group of s samples gives s * nb sweep tasks for one CTA per SM: pick the band count whose groups fill whole
        // waves (e.g. 640x480, 5 bins: 37 bands make 8 samples exactly two waves of 148)
        const int64_t per_sample = n_events > 0 ? n_events / batch : 1;
        int s_max = (int)(banded_group_budget() / (size_t)(per_sample * 8 + 1));
        if (s_max < 1) s_max = 1;
        if (s_max > batch) s_max = batch;
        double best = -1.0;
        for (int cand = nb_min; cand <= kMaxBands && cand <= nb_min + nb_min / 2 + 4; ++cand) {
            double eff = 0.0;
            for (int sg = (s_max + 1) / 2; sg <= s_max; ++sg) {
                const int64_t tasks = (int64_t)sg * cand;
                const double e = (double)tasks / (double)(ceil_div64(tasks, kNumSMs) * kNumSMs);
                if (e > eff) eff = e;
            }
            const double score = eff - 0.004 * (cand - nb_min);
            if (score > best + 1e-9) { best = score; nb = cand; }
        }
    }
    const int force_nb = env_int("EP_BANDED_NB", 0);
    if (force_nb >= nb_min && force_nb <= kMaxBands) nb = force_nb;
    bp->nb = nb;
    bp->cpb = (((H + nb - 1) / nb) * W + 3) / 4 * 4;
    bp->P = P;
    make_div((uint32_t)W, &bp->w_magic, &bp->w_shift);
    make_div((uint32_t)nb, &bp->nb_magic, &bp->nb_shift);
    bp->slot_words = words_per_cell * bp->cpb;
    bp->sweep_smem = (size_t)bp->slot_words * 4 + table_bytes;
    bp->chunk = 4096;
    return true;
}

int64_t chunks_of(const int64_t* off, int b, int chunk) {
    const int64_t lo = off[b] / kEvPerThread * kEvPerThread;
    return off[b + 1] > off[b] ? ceil_div64(off[b + 1] - lo, chunk) : 0;
}

size_t bytes_per_chunk(const BandPlan& bp) { return (size_t)bp.chunk * 8 + (size_t)bp.nb * 4 + 2; }   // per buffer set

// Second stream + events for the route / sweep overlap, one set per host thread (calls are re-entrant across threads).
struct OverlapCtx {
    int device = -1;
    cudaStream_t side = nullptr;
    cudaEvent_t start = nullptr, routed[kBufSets] = {nullptr, nullptr}, swept[kBufSets] = {nullptr, nullptr};
    bool ok = false;
};
OverlapCtx* overlap_ctx() {
    static thread_local OverlapCtx ctx;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (ctx.ok && ctx.device == dev) return &ctx;
    if (ctx.ok) return nullptr;                       // one device per process (and thread); anything else runs unoverlapped
    ctx.device = dev;
    bool good = cudaStreamCreateWithFlags(&ctx.side, cudaStreamNonBlocking) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&ctx.start, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kBufSets; ++i) {
        good = good && cudaEventCreateWithFlags(&ctx.routed[i], cudaEventDisableTiming) == cudaSuccess;
        good = good && cudaEventCreateWithFlags(&ctx.swept[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ctx.ok = good;
    if (!good) cudaGetLastError();
    return good ? &ctx : nullptr;
}

struct BandLayout { size_t meta, chunk_first, counters, info, runs[kBufSets], kk[kBufSets], rec[kBufSets], total; };

BandLayout band_layout(int B, int64_t total_chunks, int64_t group_chunks, const BandPlan& bp) {
    BandLayout L;
    L.meta = 0;
    L.chunk_first = align_up(sizeof(SampleMeta) * (size_t)B, 256);
    L.counters = L.chunk_first + align_up(sizeof(int32_t) * (size_t)(B + 1), 256);      // one sweep task counter per group
    L.info = L.counters + align_up(sizeof(unsigned int) * (size_t)B, 256);
    size_t at = L.info + align_up(sizeof(ChunkInfo) * (size_t)total_chunks, 256);
    for (int i = 0; i < kBufSets; ++i) {
        L.runs[i] = at; at += align_up((size_t)group_chunks * bp.nb * 4, 256);
        L.kk[i] = at;   at += align_up((size_t)group_chunks * 2, 256);
        L.rec[i] = at;  at += align_up((size_t)group_chunks * bp.chunk * 8, 256);
    }
    L.total = at;
    return L;
}

// largest group (in chunks) over a greedy split into consecutive samples within the record budget
int64_t max_group_chunks(const int64_t* off, int B, const BandPlan& bp, size_t budget) {
    int64_t best = 0, cur = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t c = chunks_of(off, b, bp.chunk);
        if (cur > 0 && (size_t)(cur + c) * bytes_per_chunk(bp) > budget) cur = 0;
        cur += c;
        if (cur > best) best = cur;
    }
    return best;
}

int64_t total_chunks_of(const int64_t* off, int B, int chunk) {
    int64_t t = 0;
    for (int b = 0; b < B; ++b) t += chunks_of(off, b, chunk);
    return t;
}

int banded_kind(const ep_events_soa* ev, const ep_bin_params* p) {
    if (p->time_f32 || ev->xy_dtype != EP_U16 || !aligned16(ev->x) || !aligned16(ev->y) || !aligned16(ev->t)) return -1;
    if (ev->t_dtype == EP_U32) return (ev->p == nullptr && ev->t_base) ? kKindCompact : -1;
    if (ev->p_dtype != EP_U8 || !aligned16(ev->p)) return -1;
    if (ev->t_dtype == EP_I64) return kKindTicks64;
    if (ev->t_dtype == EP_F64) return kKindF64;
    return -1;
}

template <int KIND>
cudaError_t launch_route_kind(cudaStream_t st, const RouteSrc& src, const BandArgs& g, unsigned grid) {
    const bool track = g.bin.num_bins > g.P;      // multi-window sweeps filter chunks by their interval span
    return track ? launch_route<KIND, 4096, 256, 4, true>(st, src, g, grid) : launch_route<KIND, 4096, 256, 4, false>(st, src, g, grid);
}

template <int VEC, int THREADS>
cudaError_t launch_sweep(cudaStream_t st, const BandArgs& g, int n_tasks, const BandPlan& bp) {
    static size_t configured = 0;
    if (configured < bp.sweep_smem) {
        cudaError_t ce = cudaFuncSetAttribute(k_sweep<VEC, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.sweep_smem);
        if (ce != cudaSuccess) return ce;
        configured = bp.sweep_smem;
    }
    const int grid = n_tasks < kNumSMs ? n_tasks : kNumSMs;
    k_sweep<VEC, THREADS><<<(unsigned)grid, THREADS, bp.sweep_smem, st>>>(g, n_tasks, bp.slot_words);
    return cudaSuccess;
}

template <class MetaLoader>
int run_banded(cudaStream_t st, MetaLoader mld, int kind, const RouteSrc& src, const ep_events_soa* ev, const ep_bin_params* p,
               const BandPlan& bp, float* out_voxel, float* out_sum, float* out_count, void* ws, size_t ws_bytes,
               unsigned int* bad) {
    const int B = ev->batch;
    const int64_t* off = ev->offsets_host;
    const int64_t total_chunks = total_chunks_of(off, B, bp.chunk);
    if (total_chunks >= (1ll << 31)) return EP_EUNSUPPORTED;
    const size_t fixed = band_layout(B, total_chunks, 0, bp).total;
    const size_t bpc = bytes_per_chunk(bp);
    if (ws_bytes < fixed + kBufSets * bpc + 4096) return EP_EWORKSPACE;
    const size_t cap_chunks = (ws_bytes - fixed - 4096) / (kBufSets * bpc);
    size_t budget = banded_group_budget();
    if (budget > cap_chunks * bpc) budget = cap_chunks * bpc;
    for (int b = 0; b < B; ++b)
        if ((size_t)chunks_of(off, b, bp.chunk) > cap_chunks) return EP_EWORKSPACE;
    const int64_t gchunks = max_group_chunks(off, B, bp, budget);
    const BandLayout L = band_layout(B, total_chunks, gchunks, bp);
    if (L.total > ws_bytes) return EP_EWORKSPACE;
    char* w = static_cast<char*>(ws);

    BandArgs g;
    BinArgs& a = g.bin;
    a.offsets = ev->offsets; a.single_n = 0;
    a.H = p->height; a.W = p->width; a.num_bins = p->num_bins; a.count_channels = p->count_channels;
    a.sx = p->scale_x; a.sy = p->scale_y; a.scaled = (p->scale_x != 1.0 || p->scale_y != 1.0);
    a.meta = reinterpret_cast<SampleMeta*>(w + L.meta);
    a.bad_count = bad;
    a.vox_acc = nullptr; a.cnt_acc = nullptr;
    a.n_total = off[B];
    a.begin = off[0]; a.end = off[B]; a.start4 = 0; a.n_tiles = 0; a.g0 = 0; a.g1 = B;
    g.nb = bp.nb; g.cpb = bp.cpb; g.P = bp.P; g.chunk = bp.chunk;
    g.w_magic = bp.w_magic; g.w_shift = bp.w_shift; g.nb_magic = bp.nb_magic; g.nb_shift = bp.nb_shift;
    int32_t* chunk_first = reinterpret_cast<int32_t*>(w + L.chunk_first);
    ChunkInfo* info = reinterpret_cast<ChunkInfo*>(w + L.info);
    g.chunk_first = chunk_first;
    g.info = info;
    g.gchunks = (int)gchunks;
    g.runs = reinterpret_cast<uint32_t*>(w + L.runs[0]);
    g.kk = reinterpret_cast<uint16_t*>(w + L.kk[0]);
    g.rec = reinterpret_cast<uint2*>(w + L.rec[0]);
    g.out_voxel = out_voxel; g.out_sum = out_sum; g.out_count = out_count;
    g.dbg = nullptr;
    unsigned int* counters = reinterpret_cast<unsigned int*>(w + L.counters);
    g.task_counter = counters;
    {
        cudaError_t ce = cudaMemsetAsync(counters, 0, sizeof(unsigned int) * (size_t)B, st);
        if (ce != cudaSuccess) return (int)ce;
    }
#ifdef EP_PHASE_TIMING
    static unsigned long long* s_dbg = nullptr;
    if (!s_dbg) cudaMalloc(&s_dbg, 16 * sizeof(unsigned long long));
    cudaMemsetAsync(s_dbg, 0, 16 * sizeof(unsigned long long), st);
    g.dbg = s_dbg;
#endif

    profile_begin(st, kProfOther);
    k_sample_meta<MetaLoader><<<(B + 127) / 128, 128, 0, st>>>(mld, a, B);
    EP_LAUNCH_CHECK();
    k_chunk_prefix<<<1, 1024, 0, st>>>(a, B, bp.chunk, chunk_first);
    EP_LAUNCH_CHECK();
    if (total_chunks > 0) {
        k_chunk_info<<<(unsigned)ceil_div64(total_chunks, 256), 256, 0, st>>>(a, B, bp.chunk, chunk_first, (int)total_chunks, info);
        EP_LAUNCH_CHECK();
    }
    profile_end(st);

    const bool vec4 = p->width % 4 == 0 && !(reinterpret_cast<uintptr_t>(out_voxel) & 15u) &&
                      !(reinterpret_cast<uintptr_t>(out_sum) & 15u) && !(reinterpret_cast<uintptr_t>(out_count) & 15u);

    // Routes run on the caller's stream, sweeps on a side stream: the route of group g + 1 fills the SMs the sweep of
    // group g leaves idle (its tail, launch gaps), and vice versa.  Everything is ordered after the caller's earlier work
    // and joined back before returning; no host synchronisation.
    OverlapCtx* ov = env_int("EP_BANDED_OVERLAP", 1) ? overlap_ctx() : nullptr;
    cudaStream_t st_sweep = ov ? ov->side : st;
    if (ov) {
        cudaEventRecord(ov->start, st);
        cudaStreamWaitEvent(ov->side, ov->start, 0);
    }
    int n_groups = 0;

    int64_t chunk_begin = 0;
    int g0 = 0;
    while (g0 < B) {
        // largest group the record budget allows ...
        int g_max = g0;
        int64_t cur = 0;
        while (g_max < B) {
            const int64_t c = chunks_of(off, g_max, bp.chunk);
            if (cur > 0 && ((size_t)(cur + c) * bpc > budget || cur + c > gchunks)) break;
            cur += c;
            ++g_max;
        }
        // ... then trimmed so that its (sample, band) tasks fill whole waves of the resident sweep CTAs
        int g1 = g_max;
        double best = -1.0;
        for (int cand = g_max; cand > g0 && cand >= g0 + (g_max - g0 + 1) / 2; --cand) {
            const int64_t tasks = (int64_t)(cand - g0) * bp.nb;
            const double eff = (double)tasks / (double)(ceil_div64(tasks, kNumSMs) * kNumSMs);
            if (eff > best + 1e-9) { best = eff; g1 = cand; }
        }
        cur = 0;
        for (int b = g0; b < g1; ++b) cur += chunks_of(off, b, bp.chunk);
        a.g0 = g0; a.g1 = g1;
        a.begin = off[g0]; a.end = off[g1];
        g.chunk_begin = (int)chunk_begin;
        const int buf = ov ? (n_groups % kBufSets) : 0;
        g.task_counter = counters + n_groups;
        g.runs = reinterpret_cast<uint32_t*>(w + L.runs[buf]);
        g.kk = reinterpret_cast<uint16_t*>(w + L.kk[buf]);
        g.rec = reinterpret_cast<uint2*>(w + L.rec[buf]);
        if (ov && n_groups >= kBufSets) cudaStreamWaitEvent(st, ov->swept[buf], 0);     // the buffer set is free again
        if (cur > 0) {
            profile_begin(st, kProfScatter);
            cudaError_t ce;
            if (kind == kKindTicks64) ce = launch_route_kind<kKindTicks64>(st, src, g, (unsigned)cur);
            else if (kind == kKindF64) ce = launch_route_kind<kKindF64>(st, src, g, (unsigned)cur);
            else ce = launch_route_kind<kKindCompact>(st, src, g, (unsigned)cur);
            profile_end(st);
            if (ce != cudaSuccess) return (int)ce;
            EP_LAUNCH_CHECK();
        }
        if (ov) {
            cudaEventRecord(ov->routed[buf], st);
            cudaStreamWaitEvent(st_sweep, ov->routed[buf], 0);
        }
        profile_begin(st_sweep, kProfFinalize);
        const int n_tasks = (g1 - g0) * bp.nb;
        cudaError_t ce;
        ce = vec4 ? launch_sweep<4, kSweepThreads>(st_sweep, g, n_tasks, bp) : launch_sweep<1, kSweepThreads>(st_sweep, g, n_tasks, bp);
        profile_end(st_sweep);
        if (ov) cudaEventRecord(ov->swept[buf], st_sweep);
        if (ce != cudaSuccess) return (int)ce;
        EP_LAUNCH_CHECK();
        chunk_begin += cur;
        g0 = g1;
        ++n_groups;
    }
    if (ov)      // join: the caller's stream continues after the last sweeps
        for (int i = 0; i < kBufSets && i < n_groups; ++i) cudaStreamWaitEvent(st, ov->swept[(n_groups - 1 - i) % kBufSets], 0);
#ifdef EP_PHASE_TIMING
    if (getenv("EP_PRINT_TIMING")) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, g.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[9] = {"route: setup + zero", "route: load wait + rank", "route: scan", "route: place", "route: writeout",
                                "sweep: between windows", "sweep: accumulate", "sweep: flush", "sweep: exact redo"};
        double rt = 0, sw = 0;
        for (int i = 0; i < 5; ++i) rt += (double)h[i];
        for (int i = 5; i < 9; ++i) sw += (double)h[i];
        for (int i = 0; i < 9; ++i) fprintf(stderr, "  %-28s %12.3f Mcycles  %5.1f%%\n", names[i], h[i] * 1e-6, 100.0 * h[i] / (i < 5 ? rt : sw));
        fprintf(stderr, "  sweep: windows redone exactly: %llu\n", h[9]);
    }
#endif
    return EP_OK;
}

}  // namespace

size_t banded_workspace_bytes(const ep_events_soa* ev, const ep_bin_params* p) {
    BandPlan bp;
    if (!ev || !ev->offsets_host || ev->batch <= 0 || !plan_bands(p, ev->batch, ev->offsets_host[ev->batch] - ev->offsets_host[0], &bp)) return 0;
    const int64_t total = total_chunks_of(ev->offsets_host, ev->batch, bp.chunk);
    const int64_t gchunks = max_group_chunks(ev->offsets_host, ev->batch, bp, banded_group_budget());
    return band_layout(ev->batch, total, gchunks, bp).total + kBufSets * bytes_per_chunk(bp) + 8192;
}

// true when the banded path can take this call and the batch is large enough for it to pay
bool banded_worthwhile(const ep_events_soa* ev, const ep_bin_params* p) {
    BandPlan bp;
    if (banded_kind(ev, p) < 0 || !ev->offsets_host || !plan_bands(p, ev->batch, ev->offsets_host[ev->batch] - ev->offsets_host[0], &bp)) return false;
    const int64_t n = ev->offsets_host[ev->batch] - ev->offsets_host[0];
    return n >= (int64_t)env_int("EP_BANDED_MIN_EVENTS", 4000000) && (int64_t)ev->batch * bp.nb >= kNumSMs;
}

int run_banded_canon(cudaStream_t st, const ep_events_soa* ev, const ep_bin_params* p, float* out_voxel, float* out_sum,
                     float* out_count, void* ws, size_t ws_bytes, unsigned int* bad) {
    BandPlan bp;
    const int kind = banded_kind(ev, p);
    if (kind < 0 || !plan_bands(p, ev->batch, ev->offsets_host[ev->batch] - ev->offsets_host[0], &bp)) return EP_EUNSUPPORTED;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u)) return EP_EALIGN;
    if ((p->num_bins > 0 && !out_voxel) || (p->count_channels > 0 && !out_count)) return EP_EINVAL;
    RouteSrc src{static_cast<const uint16_t*>(ev->x), static_cast<const uint16_t*>(ev->y), ev->t,
                 static_cast<const uint8_t*>(ev->p)};
    if (kind == kKindCompact) {
        SoaCompactLoader ld{src.x, src.y, static_cast<const uint32_t*>(ev->t), ev->t_base, ev->t_div};
        return run_banded(st, ld, kind, src, ev, p, bp, out_voxel, out_sum, out_count, ws, ws_bytes, bad);
    }
    if (kind == kKindTicks64) {
        SoaCanonLoader<true> ld{src.x, src.y, ev->t, src.p, ev->t_div};
        return run_banded(st, ld, kind, src, ev, p, bp, out_voxel, out_sum, out_count, ws, ws_bytes, bad);
    }
    SoaCanonLoader<false> ld{src.x, src.y, ev->t, src.p, ev->t_div};
    return run_banded(st, ld, kind, src, ev, p, bp, out_voxel, out_sum, out_count, ws, ws_bytes, bad);
}

}  // namespace ep

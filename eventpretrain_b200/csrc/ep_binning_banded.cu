// Stage 1, alternative path — banded shared-memory sweep (sm_100a).  EXPERIMENTAL in round 1: correct and bit-identical
// to the global-RED path (tests/test_gpu_stage1.py::test_banded_*), selected only with EP_BIN_FORCE_BANDED because it is
// still slower on the benchmark workload (3.8 ms vs 2.9 ms per step; analysis in profiles/r01_banded_phase_timing.txt).
//
// Same arithmetic and outputs as ep_binning.cu (events_to_voxel_grid.py:4-61, events_to_image.py:6-62), but no global
// atomics and no accumulator round trip:
//
//   route  (k_route)   one pass over the events of a sample group: each CTA takes a 4096-event chunk, computes
//                      (cell, interval k, r = rn(d*2^24), polarity) per event, counting-sorts the chunk by spatial band
//                      in shared memory and writes it back as 6-byte records (u16 cell-in-band + u32 r|k<<25|p<<30)
//                      with a per-chunk band-offset row.  The record buffer of a group is sized to stay L2-resident.
//   sweep  (k_sweep)   persistent CTAs take (sample, band) tasks and own the band for every interval: per interval they
//                      pull the band's records, accumulate them with fire-and-forget shared-memory ATOMS.ADD.u32
//                      (measured ~8x the throughput of global RED on B200) into two words per cell,
//                          N = n_pos | n_neg << 16     (exact polarity counts)
//                          A = sum p*r  (mod 2^32)     (exact while the cell holds <= 127 events of the interval)
//                      then emit voxel[k] = (C_k*2^24 - A_k + A_{k-1}) * 2^-24 for their cells straight to the fp32 output
//                      (coalesced, streaming).  Cells with more than 127 events in an interval are detected from the
//                      exact count and recomputed by re-scanning the interval into a 64-bit hash table.
//
// Integer accumulation => the result is order-independent and bit-identical to the global-RED path.
// Unsorted input is handled (chunks are revisited for every interval they contain), just slower.
#include <stdio.h>

#include "ep_binning_common.cuh"

namespace ep {
namespace {

constexpr int kChunk = 4096;            // events per routed chunk
constexpr int kRouteThreads = 512;      // 2 quads (8 events) per thread
constexpr int kMaxBands = 64;
constexpr int kSweepThreads = 512;
constexpr int kCPT4 = 3;                // quads of cells per sweep thread (register-resident A_{k-1} and sum)
constexpr int kCPT = 4 * kCPT4;
constexpr int kMaxCellsPerBand = kCPT * kSweepThreads;   // 6144 cells -> 48 KB of tile, two CTAs per SM
constexpr int kTeam = 8;                // lanes that walk one chunk's run of records together
constexpr int kUnroll = 12;             // records in flight per lane: a typical run (4096 / 50 bands) is one round
constexpr int kAdmit = 127;             // events per (cell, interval) accumulated in the 32-bit A word
constexpr int kSpill = 256;             // spill-table slots per CTA (power of two)
constexpr int kRowStride = kMaxBands + 2;   // u16 entries per chunk row: offsets[0..NB], then kmin|kmax<<8
constexpr int kMaxTableChunks = 1024;   // chunk rows of one sample staged in shared memory at a time
constexpr uint32_t kKShift = 25, kPolShift = 30;
constexpr uint32_t kKCountOnly = 31;    // interval code of events outside the time bins (count frame only)

// Optional per-phase cycle accounting (build with -DEP_PHASE_TIMING; tools only, never in the shipped library).
#ifdef EP_PHASE_TIMING
#define EP_TICK(slot)                                                                      \
    do {                                                                                   \
        if (threadIdx.x == 0) {                                                            \
            const long long _now = clock64();                                              \
            atomicAdd(g.dbg + (slot), (unsigned long long)(_now - _t_last));               \
            _t_last = _now;                                                                \
        }                                                                                  \
    } while (0)
#define EP_TICK_INIT() long long _t_last = clock64()
#else
#define EP_TICK(slot) do { } while (0)
#define EP_TICK_INIT() do { } while (0)
#endif

struct BandArgs {
    BinArgs bin;                 // offsets, meta, geometry, bad_count (begin/end/g0/g1 describe the group)
    int nb;                      // bands
    int cpb;                     // cells per band
    uint32_t cpb_magic;          // flat / cpb == umulhi(flat, cpb_magic) >> cpb_shift  for flat < 2^31
    int cpb_shift;
    const int32_t* chunk_first;  // [B+1] first chunk id of each sample (global numbering)
    int chunk_begin;             // first chunk id of this group
    uint16_t* rows;              // [chunks_in_group][kRowStride]
    uint32_t* rec_val;           // [chunks_in_group][kChunk]
    uint16_t* rec_cell;          // [chunks_in_group][kChunk]
    float* out_voxel; float* out_sum; float* out_count;
    unsigned long long* dbg;     // phase cycle counters (EP_PHASE_TIMING builds), else unused
};

__device__ __forceinline__ int64_t sample_chunk_origin(const BinArgs& a, int b) {
    return off_at(a, b) / kEvPerThread * kEvPerThread;   // chunks start on the 4-event grid of the arrays
}

// ---- chunk numbering: chunks of sample b = ceil((off[b+1] - align4(off[b])) / kChunk) ------------------
__global__ void __launch_bounds__(1024) k_chunk_prefix(BinArgs a, int B, int32_t* __restrict__ chunk_first) {
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
            v = hi > off_at(a, b) ? (int)ceil_div64(hi - lo, kChunk) : 0;
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

// ---- route ------------------------------------------------------------------------------------------------
struct RouteInfo {          // per-CTA facts computed once by thread 0
    int b;
    int lean;
    int64_t c_lo, ev_lo, ev_hi;
    SampleMeta m;
};

template <class Loader>
__global__ void __launch_bounds__(kRouteThreads, 2) k_route(Loader ld, BandArgs g) {
    __shared__ __align__(16) uint32_t s_val[kChunk];
    __shared__ __align__(16) uint16_t s_cell[kChunk];
    __shared__ int s_cnt[kMaxBands], s_base[kMaxBands + 1], s_cur[kMaxBands];
    __shared__ int s_kmin, s_kmax;
    __shared__ RouteInfo s_info;
    const BinArgs& a = g.bin;
    if (threadIdx.x == 0) {
        const int chunk = g.chunk_begin + blockIdx.x;
        int lo = a.g0, hi = a.g1;          // owning sample: last b in [g0, g1) with chunk_first[b] <= chunk
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (g.chunk_first[mid] <= chunk) lo = mid; else hi = mid;
        }
        const int b = lo;
        const int64_t s_lo = off_at(a, b), s_hi = off_at(a, b + 1);
        const int64_t c_lo = sample_chunk_origin(a, b) + (int64_t)(chunk - g.chunk_first[b]) * kChunk;
        s_info.b = b;
        s_info.c_lo = c_lo;
        s_info.ev_lo = c_lo > s_lo ? c_lo : s_lo;
        s_info.ev_hi = (c_lo + kChunk < s_hi) ? c_lo + kChunk : s_hi;
        // interior chunk (all 4096 events belong to sample b, hence lie inside the arrays), unscaled coordinates:
        // lean 32-bit path without per-event range checks; everything else takes the general loader path
        s_info.lean = Loader::kFastTime && !a.scaled && c_lo >= s_lo && c_lo + kChunk <= s_hi && a.W < 65536;
        s_info.m = a.meta[b];
        s_kmin = 255; s_kmax = 0;
    }
    if (threadIdx.x < kMaxBands) { s_cnt[threadIdx.x] = 0; s_cur[threadIdx.x] = 0; }
    EP_TICK_INIT();
    __syncthreads();
    EP_TICK(0);

    const int64_t c_lo = s_info.c_lo, ev_lo = s_info.ev_lo, ev_hi = s_info.ev_hi;
    const int64_t HW = (int64_t)a.H * a.W;
    uint32_t val[2 * kEvPerThread];
    uint32_t cb[2 * kEvPerThread];        // cell-in-band | band << 16, 0xffffffff = dropped
    int kmin = 255, kmax = 0;
    if (s_info.lean) {
        const uint32_t W32 = (uint32_t)a.W, HW32 = (uint32_t)HW;
        const int nbm1 = a.num_bins - 1;
        const double nb_d = (double)a.num_bins, scale = s_info.m.scale_raw, t0_raw = s_info.m.t0_raw;
        const long long t0_ticks = s_info.m.t0_ticks;
        // all loads of both quads first, so their latency overlaps
        uint2 xv[2], yv[2];
        uint32_t pv[2];
        longlong2 tq[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i0 = c_lo + ((int64_t)h * kRouteThreads + threadIdx.x) * kEvPerThread;
            xv[h] = ld_stream(reinterpret_cast<const uint2*>(ld.x + i0));
            yv[h] = ld_stream(reinterpret_cast<const uint2*>(ld.y + i0));
            pv[h] = ld_stream(reinterpret_cast<const uint32_t*>(ld.p + i0));
            tq[h][0] = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(ld.t) + i0));
            tq[h][1] = ld_stream(reinterpret_cast<const longlong2*>(static_cast<const int64_t*>(ld.t) + i0 + 2));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long raw[4] = {tq[h][0].x, tq[h][0].y, tq[h][1].x, tq[h][1].y};
            const uint32_t xs[4] = {xv[h].x & 0xffffu, xv[h].x >> 16, xv[h].y & 0xffffu, xv[h].y >> 16};
            const uint32_t ys[4] = {yv[h].x & 0xffffu, yv[h].x >> 16, yv[h].y & 0xffffu, yv[h].y >> 16};
#pragma unroll
            for (int j = 0; j < kEvPerThread; ++j) {
                const int q = h * kEvPerThread + j;
                cb[q] = 0xffffffffu;
                const uint32_t flat = ys[j] * W32 + xs[j];
                const uint32_t pb = (pv[h] >> (8 * j)) & 0xffu;
                if (flat >= HW32 || pb > 1u) {
                    if (a.bad_count) atomicAdd(a.bad_count, 1u);
                    continue;
                }
                const double dt = Loader::kTicks ? (double)(raw[j] - t0_ticks) : (__longlong_as_double(raw[j]) - t0_raw);
                const double ts = dt * scale;
                int k = (int)kKCountOnly, r = 0;
                if (nbm1 >= 0 && ts >= 0.0 && ts < nb_d) {
                    k = __double2int_rd(ts);
                    r = __float2int_rn((float)(ts - (double)k) * 16777216.0f);
                    const bool on_last = (k == nbm1) & (r == 0) & (nbm1 > 0);     // exactly on the last node
                    k -= on_last;
                    r = on_last ? (1 << kQ) : r;
                    kmin = min(kmin, k); kmax = max(kmax, k);
                } else if (!a.count_channels) {
                    continue;                                  // contributes to nothing
                }
                const uint32_t bd = __umulhi(flat, g.cpb_magic) >> g.cpb_shift;
                cb[q] = (flat - bd * (uint32_t)g.cpb) | (bd << 16);
                val[q] = (uint32_t)r | ((uint32_t)k << kKShift) | (pb << kPolShift);
                atomicAdd(&s_cnt[bd], 1);
            }
        }
    } else {
        const SampleMeta m = s_info.m;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i0 = c_lo + ((int64_t)h * kRouteThreads + threadIdx.x) * kEvPerThread;
            Ev<double> e;
            if (i0 < ev_hi && i0 + kEvPerThread > ev_lo) ld.load(i0, ev_hi, a, e);
#pragma unroll
            for (int j = 0; j < kEvPerThread; ++j) {
                const int q = h * kEvPerThread + j;
                const int64_t i = i0 + j;
                cb[q] = 0xffffffffu;
                if (i < ev_lo || i >= ev_hi) continue;
                const int cls = e.cls[j];
                const int64_t flat64 = e.x[j] + e.y[j] * (int64_t)a.W;
                if (cls == 3 || flat64 < 0 || flat64 >= HW) {
                    if (a.bad_count) atomicAdd(a.bad_count, 1u);
                    continue;
                }
                int k, r;
                bool last_plane;
                const bool in_bins = a.num_bins > 0 && voxel_weights<Loader, double>(e, j, m, a.num_bins, k, r, last_plane);
                if (!in_bins) {
                    if (!a.count_channels) continue;     // contributes to nothing
                    k = kKCountOnly; r = 0;
                }
                const uint32_t flat = (uint32_t)flat64;
                const uint32_t bd = __umulhi(flat, g.cpb_magic) >> g.cpb_shift;
                cb[q] = (flat - bd * (uint32_t)g.cpb) | (bd << 16);
                val[q] = (uint32_t)r | ((uint32_t)k << kKShift) | ((cls == 0 ? 1u : 0u) << kPolShift);
                if (k != (int)kKCountOnly) { kmin = min(kmin, k); kmax = max(kmax, k); }
                atomicAdd(&s_cnt[bd], 1);
            }
        }
    }
    kmin = warp_reduce(kmin, [](int x, int y) { return min(x, y); });
    kmax = warp_reduce(kmax, [](int x, int y) { return max(x, y); });
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_kmin, kmin); atomicMax(&s_kmax, kmax); }
    __syncthreads();
    EP_TICK(1);
    if (threadIdx.x < 32) {      // exclusive scan of the (<= 64) band counts: two per lane
        const int l = threadIdx.x;
        const int c0 = s_cnt[2 * l], c1 = s_cnt[2 * l + 1];
        const int incl = warp_incl_scan(c0 + c1, l);
        s_base[2 * l] = incl - c0 - c1;
        s_base[2 * l + 1] = incl - c1;
        if (l == 31) s_base[kMaxBands] = incl;
    }
    __syncthreads();
    EP_TICK(2);
#pragma unroll
    for (int q = 0; q < 2 * kEvPerThread; ++q) {
        if (cb[q] == 0xffffffffu) continue;
        const int bd = cb[q] >> 16;
        const int pos = s_base[bd] + atomicAdd(&s_cur[bd], 1);
        s_val[pos] = val[q];
        s_cell[pos] = (uint16_t)cb[q];
    }
    __syncthreads();
    EP_TICK(3);
    // write the sorted chunk back: 4 records per thread per step, 16-byte / 8-byte vectors (the tail of the
    // last vector may carry stale staging data; the band offsets in the row delimit what is read)
    const int n_vec = (s_base[kMaxBands] + 3) >> 2;
    uint4* gv = reinterpret_cast<uint4*>(g.rec_val + (int64_t)blockIdx.x * kChunk);
    uint2* gc = reinterpret_cast<uint2*>(g.rec_cell + (int64_t)blockIdx.x * kChunk);
    for (int i = threadIdx.x; i < n_vec; i += kRouteThreads) {
        gv[i] = reinterpret_cast<const uint4*>(s_val)[i];
        gc[i] = reinterpret_cast<const uint2*>(s_cell)[i];
    }
    uint16_t* row = g.rows + (int64_t)blockIdx.x * kRowStride;
    if (threadIdx.x <= g.nb) row[threadIdx.x] = (uint16_t)s_base[threadIdx.x < g.nb ? threadIdx.x : kMaxBands];
    if (threadIdx.x == 0) row[kMaxBands + 1] = (uint16_t)((s_kmin & 0xff) | (s_kmax << 8));
    EP_TICK(4);
}

// ---- sweep ------------------------------------------------------------------------------------------------
struct SpillTable {
    unsigned int key[kSpill];              // cell + 1, 0 = empty
    unsigned long long acc[kSpill];        // sum p*r of the events beyond the admitted ones
};

__device__ __forceinline__ void spill_add(SpillTable* t, unsigned int cell, long long v, unsigned int* bad) {
    unsigned int h = (cell * 2654435761u) >> (32 - 8);
    for (int probe = 0; probe < kSpill; ++probe) {
        const unsigned int prev = atomicCAS(&t->key[h], 0u, cell + 1);
        if (prev == 0u || prev == cell + 1) { atomicAdd(&t->acc[h], (unsigned long long)v); return; }
        h = (h + 1) & (kSpill - 1);
    }
    if (bad) atomicOr(bad, 0x80000000u);   // more than kSpill hot cells in one band and interval
}

__device__ __forceinline__ long long spill_get(const SpillTable* t, unsigned int cell) {
    unsigned int h = (cell * 2654435761u) >> (32 - 8);
    for (int probe = 0; probe < kSpill; ++probe) {
        const unsigned int k = t->key[h];
        if (k == cell + 1) return (long long)t->acc[h];
        if (k == 0u) return 0;
        h = (h + 1) & (kSpill - 1);
    }
    return 0;
}

// Tile, structure of arrays over the band's cells:
//   N[cell]  = n_pos | n_neg << 16 of the current interval          (ATOMS, fire and forget)
//   A[cell]  = sum p*r mod 2^32 over the interval's events          (ATOMS, fire and forget)
//   CS[cell] = { carry = A of the previous interval (low 32 bits), running fp32 sum over bins }
// A is exact as a signed 32-bit value while the cell holds at most kAdmit (127) events of the interval
// (|A| <= 127 * 2^24 < 2^31).  Cells that received more ("hot") are detected at flush time from the exact count and
// recomputed: the interval's records are re-scanned and the hot cells' weights are summed in a small 64-bit hash
// table.  Carries that do not fit 32 bits live in two more such tables (prev / next).  Both paths are rare, so the
// common path has no returning atomics and no dependent chains.
// Counter fields are 16 bits: a pass whose per-cell counts do not add up to the number of records it consumed has
// wrapped a field (> 65535 events of one polarity on one pixel in one interval) and is reported through bad_count.
template <bool RESCAN>
__device__ __forceinline__ int scan_records(const BandArgs& g, int k, int ch0, int t0, int nt, int team, int tl,
                                            const uint32_t* s_run, const uint16_t* s_kk, uint32_t* sN, uint32_t* sA,
                                            SpillTable* hot, unsigned int* bad) {
    constexpr int kTeams = kSweepThreads / kTeam;
    int matched = 0;
    // one team of kTeam lanes per chunk; chunks are time-ordered, so the chunks relevant to interval k are (for sorted
    // input) consecutive and land on distinct teams: one round of loads per pass
    for (int c = team; c < nt; c += kTeams) {
        const uint32_t kk = s_kk[c];
        const int kmin = kk & 0xff, kmax = kk >> 8;
        // count-only records are not covered by [kmin, kmax]: every chunk is scanned in that pass
        if (k != (int)kKCountOnly && (k < kmin || k > kmax)) continue;
        const uint32_t run = s_run[c];
        const int lo = run & 0xffff, hi = run >> 16;
        const uint32_t* rv = g.rec_val + (int64_t)(ch0 + t0 + c) * kChunk;
        const uint16_t* rc = g.rec_cell + (int64_t)(ch0 + t0 + c) * kChunk;
        for (int i0 = lo + tl; i0 < hi; i0 += kUnroll * kTeam) {
            uint32_t v[kUnroll], cl[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int i = i0 + u * kTeam;
                v[u] = 0xffffffffu;            // interval field 31 + polarity bit: never equals a voxel pass's k ...
                cl[u] = 0;
                if (i < hi) { v[u] = __ldcs(rv + i); cl[u] = __ldcs(rc + i); }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                if (i0 + u * kTeam >= hi) continue;   // ... but the count-only pass has k == 31, so test the bound too
                if ((int)((v[u] >> kKShift) & 31u) != k) continue;
                const bool pos = (v[u] >> kPolShift) & 1u;
                const int r = (int)(v[u] & 0x1ffffffu);
                if (!RESCAN) {
                    ++matched;
                    atomicAdd(sN + cl[u], pos ? 1u : 0x10000u);
                    if (k != (int)kKCountOnly) atomicAdd(sA + cl[u], (uint32_t)(pos ? r : -r));
                } else {
                    const uint32_t n = sN[cl[u]];
                    if ((n & 0xffffu) + (n >> 16) > (uint32_t)kAdmit) spill_add(hot, cl[u], pos ? (long long)r : -(long long)r, bad);
                }
            }
        }
    }
    return matched;
}

__device__ __forceinline__ void spill_clear(SpillTable* t, int tid) {
    for (int i = tid; i < kSpill; i += kSweepThreads) { t->key[i] = 0; t->acc[i] = 0ull; }
}

__device__ __forceinline__ float q24_to_float(long long val) {
    // one rounding from the exact fixed-point value; the 32-bit conversion is the same rounding when it applies
    const int lo = (int)val;
    const float f = ((long long)lo == val) ? (float)lo : __ll2float_rn(val);
    return f * (1.0f / 16777216.0f);
}

// Persistent over the (sample, band) tasks of a group: grid = min(tasks, 2 x SMs).
template <bool COUNT>
__global__ void __launch_bounds__(kSweepThreads, 2) k_sweep(BandArgs g, int n_tasks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BinArgs& a = g.bin;
    uint32_t* sN = reinterpret_cast<uint32_t*>(smem_raw);                   // [cpb]
    uint32_t* sA = sN + g.cpb;                                              // [cpb]
    uint2* sCS = reinterpret_cast<uint2*>(sA + g.cpb);                      // [cpb] {carry, sum}
    uint2* sCnt = sCS + g.cpb;                                              // [cpb] {count_pos, count_neg} (COUNT only)
    SpillTable* spill = reinterpret_cast<SpillTable*>(sCnt + (COUNT ? g.cpb : 0));   // [3]: hot, carry prev, carry next
    uint32_t* s_run = reinterpret_cast<uint32_t*>(spill + 3);              // lo | hi << 16 per staged chunk
    uint16_t* s_kk = reinterpret_cast<uint16_t*>(s_run + kMaxTableChunks);  // kmin | kmax << 8 per staged chunk
    __shared__ int s_hot, s_carry[2];          // flags: hot cells in this pass; wide carries in tables prev / next
    __shared__ unsigned int s_check[2];        // records consumed / events counted in the tile, per pass

    const int tid = threadIdx.x;
    const int team = tid / kTeam, tl = tid % kTeam;
    const int64_t HW = (int64_t)a.H * a.W;
    const int B = a.num_bins;
    const int n_pass = B + (COUNT ? 1 : 0);
    SpillTable* hot = spill;

    EP_TICK_INIT();
    for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int band = task % g.nb;
        const int b = a.g0 + task / g.nb;
        const int64_t band_base = (int64_t)band * g.cpb;
        const int ncell = (int)((band_base + g.cpb <= HW) ? g.cpb : (HW > band_base ? HW - band_base : 0));
        const int ch0 = g.chunk_first[b] - g.chunk_begin, nch = g.chunk_first[b + 1] - g.chunk_first[b];
        // the sample's chunk table (this band's run offsets + interval span) is staged in shared memory once per
        // task when it fits (<= kMaxTableChunks chunks = 4.2 M events), else re-staged tile by tile in every pass
        const bool table_once = nch <= kMaxTableChunks;

        __syncthreads();                                   // previous task fully flushed
        EP_TICK(8);                                        // tail of the previous task (sum / count write-out + wait)
        for (int i = tid; i < g.cpb; i += kSweepThreads) {
            sN[i] = 0; sA[i] = 0; sCS[i] = make_uint2(0u, 0u);
            if (COUNT) sCnt[i] = make_uint2(0u, 0u);
        }
        for (int t = 0; t < 3; ++t) spill_clear(spill + t, tid);
        if (tid == 0) { s_hot = 0; s_carry[0] = 0; s_carry[1] = 0; s_check[0] = 0; s_check[1] = 0; }
        if (table_once) {
            for (int c = tid; c < nch; c += kSweepThreads) {
                const uint16_t* row = g.rows + (int64_t)(ch0 + c) * kRowStride;
                s_run[c] = (uint32_t)row[band] | ((uint32_t)row[band + 1] << 16);
                s_kk[c] = row[kMaxBands + 1];
            }
        }
        int prev = 0, next = 1;                            // roles of the two carry tables (spill[1 + prev], spill[1 + next])

        // intervals 0..B-1 produce voxel planes; the pseudo-interval kKCountOnly only feeds the count frame
        for (int pass = 0; pass < n_pass; ++pass) {
            const int k = pass < B ? pass : (int)kKCountOnly;
            int matched = 0;
            for (int t0 = 0; t0 < nch; t0 += kMaxTableChunks) {
                const int nt = min(kMaxTableChunks, nch - t0);
                __syncthreads();                           // tile zeroed / flushed, table visible
                EP_TICK(pass == 0 ? 5 : 7);                // 5: task set-up; 7: flush of the previous pass
                if (!table_once) {
                    for (int c = tid; c < nt; c += kSweepThreads) {
                        const uint16_t* row = g.rows + (int64_t)(ch0 + t0 + c) * kRowStride;
                        s_run[c] = (uint32_t)row[band] | ((uint32_t)row[band + 1] << 16);
                        s_kk[c] = row[kMaxBands + 1];
                    }
                    __syncthreads();
                }
                matched += scan_records<false>(g, k, ch0, t0, nt, team, tl, s_run, s_kk, sN, sA, hot, a.bad_count);
            }
            matched = warp_reduce(matched, [](int x, int y) { return x + y; });
            if ((tid & 31) == 0 && matched) atomicAdd(&s_check[0], (unsigned int)matched);
            __syncthreads();
            EP_TICK(6);                                    // record phase (incl. waiting for the slowest warp)

            // ---- flush: every cell of the band emits voxel[k]; hot cells are deferred to the exact path below
            const bool wide_prev = s_carry[prev] != 0;
            float* o = (pass < B) ? g.out_voxel + ((int64_t)b * B + k) * HW + band_base : nullptr;
            unsigned int counted = 0;
#pragma unroll 2
            for (int cell = tid; cell < ncell; cell += kSweepThreads) {
                const uint32_t n = sN[cell];
                if (pass >= B) {                           // count-only pseudo-interval
                    if (n) {
                        counted += (n & 0xffffu) + (n >> 16);
                        uint2 c = sCnt[cell]; c.x += n & 0xffffu; c.y += n >> 16; sCnt[cell] = c;
                        sN[cell] = 0;
                    }
                    continue;
                }
                const uint2 cs = sCS[cell];
                if (n == 0 && !wide_prev) {                // no event in this interval: voxel[k] = carry
                    if (cs.x == 0) { st_stream(o + cell, 0.0f); continue; }
                    const float r = (float)(int32_t)cs.x * (1.0f / 16777216.0f);
                    st_stream(o + cell, r);
                    sCS[cell] = make_uint2(0u, __float_as_uint(__uint_as_float(cs.y) + r));
                    continue;
                }
                const int np = (int)(n & 0xffffu), nn = (int)(n >> 16);
                counted += (unsigned int)(np + nn);
                if (np + nn > kAdmit) { s_hot = 1; continue; }       // exact path below; N and A stay for the re-scan
                if (COUNT) { uint2 c = sCnt[cell]; c.x += np; c.y += nn; sCnt[cell] = c; }
                const int32_t A = (int32_t)sA[cell];
                long long carry = (long long)(int32_t)cs.x;
                if (wide_prev) carry += spill_get(spill + 1 + prev, (unsigned int)cell);
                const float r = q24_to_float(((long long)(np - nn) << kQ) - (long long)A + carry);
                st_stream(o + cell, r);
                sCS[cell] = make_uint2((uint32_t)A, __float_as_uint(__uint_as_float(cs.y) + r));
                if (n) { sN[cell] = 0; sA[cell] = 0; }
            }
            counted = warp_reduce(counted, [](unsigned int x, unsigned int y) { return x + y; });
            if ((tid & 31) == 0 && counted) atomicAdd(&s_check[1], counted);
            __syncthreads();
            if (pass < B && s_hot) {
                // ---- exact path for hot cells: re-scan the interval's records, 64-bit sums in the hash table
                for (int t0 = 0; t0 < nch; t0 += kMaxTableChunks) {
                    const int nt = min(kMaxTableChunks, nch - t0);
                    if (!table_once) {
                        __syncthreads();
                        for (int c = tid; c < nt; c += kSweepThreads) {
                            const uint16_t* row = g.rows + (int64_t)(ch0 + t0 + c) * kRowStride;
                            s_run[c] = (uint32_t)row[band] | ((uint32_t)row[band + 1] << 16);
                            s_kk[c] = row[kMaxBands + 1];
                        }
                        __syncthreads();
                    }
                    scan_records<true>(g, k, ch0, t0, nt, team, tl, s_run, s_kk, sN, sA, hot, a.bad_count);
                }
                __syncthreads();
                for (int cell = tid; cell < ncell; cell += kSweepThreads) {
                    const uint32_t n = sN[cell];
                    const int np = (int)(n & 0xffffu), nn = (int)(n >> 16);
                    if (np + nn <= kAdmit) continue;
                    if (COUNT) { uint2 c = sCnt[cell]; c.x += np; c.y += nn; sCnt[cell] = c; }
                    const uint2 cs = sCS[cell];
                    const long long A = spill_get(hot, (unsigned int)cell);
                    long long carry = (long long)(int32_t)cs.x;
                    if (wide_prev) carry += spill_get(spill + 1 + prev, (unsigned int)cell);
                    const float r = q24_to_float(((long long)(np - nn) << kQ) - A + carry);
                    st_stream(o + cell, r);
                    const long long rest = A - (long long)(int32_t)(uint32_t)A;      // carry = low 32 bits + rest
                    if (rest != 0) { spill_add(spill + 1 + next, (unsigned int)cell, rest, a.bad_count); s_carry[next] = 1; }
                    sCS[cell] = make_uint2((uint32_t)A, __float_as_uint(__uint_as_float(cs.y) + r));
                    sN[cell] = 0; sA[cell] = 0;
                }
                __syncthreads();
                spill_clear(hot, tid);
                if (tid == 0) s_hot = 0;
            }
            if (pass < B && (wide_prev || s_carry[next])) {
                // rotate the carry tables: "next" becomes "prev" for the following interval, the old "prev" is cleared
                __syncthreads();
                if (wide_prev) spill_clear(spill + 1 + prev, tid);
                if (tid == 0) s_carry[prev] = 0;
                const int t = prev; prev = next; next = t;
            }
            if (tid == 0) {
                if (s_check[0] != s_check[1] && a.bad_count) atomicOr(a.bad_count, 0x80000000u);   // a 16-bit count wrapped
                s_check[0] = 0; s_check[1] = 0;
            }
            // the next pass starts with a __syncthreads() before the tile and the tables are touched again
        }
        __syncthreads();
        if (g.out_sum && B > 0) {
            float* sp = g.out_sum + (int64_t)b * HW + band_base;
            for (int cell = tid; cell < ncell; cell += kSweepThreads) st_stream(sp + cell, __uint_as_float(sCS[cell].y));
        }
        if (COUNT) {
            float* c0 = g.out_count + (int64_t)b * a.count_channels * HW + band_base;
            float* cn = c0 + (int64_t)(a.count_channels - 1) * HW;
            for (int cell = tid; cell < ncell; cell += kSweepThreads) {
                const uint2 c = sCnt[cell];
                st_stream(c0 + cell, (float)c.x);
                st_stream(cn + cell, (float)c.y);
                if (a.count_channels == 3) st_stream(c0 + HW + cell, 0.0f);
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
struct BandPlan {
    int nb, cpb, shift;
    uint32_t magic;
    size_t sweep_smem;
};

bool plan_bands(const ep_bin_params* p, BandPlan* bp) {
    const int64_t HW = (int64_t)p->height * p->width;
    int nb = (int)ceil_div64(HW, kMaxCellsPerBand);
    if (nb > kMaxBands) return false;
    int cpb = (int)ceil_div64(HW, nb);
    cpb = (cpb + 3) / 4 * 4;
    if (cpb > kMaxCellsPerBand || cpb > 65536) return false;
    bp->nb = (int)ceil_div64(HW, cpb);
    bp->cpb = cpb;
    // exact flat / cpb for flat < 2^31: s = ceil(log2 cpb), magic = floor(2^(31+s) / cpb) + 1, q = umulhi(n, magic) >> (s-1)
    int sh = 0;
    while ((1ll << sh) < cpb) ++sh;
    bp->magic = (uint32_t)(((1ull << (31 + sh)) / (uint64_t)cpb) + 1);
    bp->shift = sh - 1;
    bp->sweep_smem = (size_t)cpb * 16 + 3 * sizeof(SpillTable) + kMaxTableChunks * (4 + 2) + 64;     // + cpb * 8 with a count frame
    return true;
}

size_t banded_group_budget() {
    const char* e = getenv("EP_BANDED_GROUP_MB");
    long mb = e ? atol(e) : 72;
    if (mb < 1) mb = 1;
    return (size_t)mb << 20;
}

int64_t chunks_of(const int64_t* off, int b) {
    const int64_t lo = off[b] / kEvPerThread * kEvPerThread;
    return off[b + 1] > off[b] ? ceil_div64(off[b + 1] - lo, kChunk) : 0;
}

constexpr size_t kBytesPerChunk = (size_t)kChunk * 6 + kRowStride * 2;

struct BandLayout { size_t meta, chunk_first, rows, rec_val, rec_cell, total; int64_t max_group_chunks; };

// groups: consecutive samples while the record buffer stays within the L2 budget (at least one sample)
int64_t max_group_chunks(const int64_t* off, int B, size_t budget, size_t ws_cap_chunks) {
    int64_t best = 0, cur = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t c = chunks_of(off, b);
        if (cur > 0 && ((size_t)(cur + c) * kBytesPerChunk > budget || (size_t)(cur + c) > ws_cap_chunks)) cur = 0;
        cur += c;
        if (cur > best) best = cur;
    }
    return best;
}

BandLayout band_layout(int B, int64_t group_chunks) {
    BandLayout L;
    L.meta = 0;
    L.chunk_first = align_up(sizeof(SampleMeta) * (size_t)B, 256);
    L.rows = L.chunk_first + align_up(sizeof(int32_t) * (size_t)(B + 1), 256);
    L.rec_val = L.rows + align_up((size_t)group_chunks * kRowStride * 2, 256);
    L.rec_cell = L.rec_val + align_up((size_t)group_chunks * kChunk * 4, 256);
    L.total = L.rec_cell + align_up((size_t)group_chunks * kChunk * 2, 256);
    L.max_group_chunks = group_chunks;
    return L;
}

bool banded_eligible(const ep_events_soa* ev, const ep_bin_params* p) {
    if (p->time_f32 || p->num_bins > 30) return false;
    if (!(ev->xy_dtype == EP_U16 && ev->p_dtype == EP_U8 && (ev->t_dtype == EP_I64 || ev->t_dtype == EP_F64))) return false;
    if (!aligned16(ev->x) || !aligned16(ev->y) || !aligned16(ev->t) || !aligned16(ev->p)) return false;
    return true;
}

template <class Loader>
int run_banded(cudaStream_t st, Loader ld, const ep_events_soa* ev, const ep_bin_params* p, const BandPlan& bp,
               float* out_voxel, float* out_sum, float* out_count, void* ws, size_t ws_bytes, unsigned int* bad) {
    const int B = ev->batch;
    const int64_t* off = ev->offsets_host;
    // how many chunks fit the caller's workspace
    const size_t fixed = band_layout(B, 0).total;
    if (ws_bytes < fixed + kBytesPerChunk + 1024) return EP_EWORKSPACE;
    const size_t cap_chunks = (ws_bytes - fixed - 1024) / kBytesPerChunk;
    for (int b = 0; b < B; ++b)
        if ((size_t)chunks_of(off, b) > cap_chunks) return EP_EWORKSPACE;
    const size_t budget = banded_group_budget();
    const int64_t gchunks = max_group_chunks(off, B, budget, cap_chunks);
    const BandLayout L = band_layout(B, gchunks);
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
    a.begin = off[0]; a.end = off[B]; a.start4 = 0; a.g0 = 0; a.g1 = B;
    g.nb = bp.nb; g.cpb = bp.cpb; g.cpb_magic = bp.magic; g.cpb_shift = bp.shift;
    int32_t* chunk_first = reinterpret_cast<int32_t*>(w + L.chunk_first);
    g.chunk_first = chunk_first;
    g.rows = reinterpret_cast<uint16_t*>(w + L.rows);
    g.rec_val = reinterpret_cast<uint32_t*>(w + L.rec_val);
    g.rec_cell = reinterpret_cast<uint16_t*>(w + L.rec_cell);
    g.out_voxel = out_voxel; g.out_sum = out_sum; g.out_count = out_count;
    g.dbg = nullptr;
#ifdef EP_PHASE_TIMING
    static unsigned long long* s_dbg = nullptr;
    if (!s_dbg) cudaMalloc(&s_dbg, 16 * sizeof(unsigned long long));
    cudaMemsetAsync(s_dbg, 0, 16 * sizeof(unsigned long long), st);
    g.dbg = s_dbg;
#endif

    profile_begin(st, kProfOther);
    k_sample_meta<Loader><<<(B + 127) / 128, 128, 0, st>>>(ld, a, B);
    profile_end(st);
    EP_LAUNCH_CHECK();
    k_chunk_prefix<<<1, 1024, 0, st>>>(a, B, chunk_first);
    EP_LAUNCH_CHECK();

    const bool count = p->count_channels != 0;
    const size_t sweep_smem = bp.sweep_smem + (count ? (size_t)bp.cpb * 8 : 0);
    cudaError_t ce = count ? cudaFuncSetAttribute(k_sweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem)
                           : cudaFuncSetAttribute(k_sweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem);
    if (ce != cudaSuccess) return (int)ce;
    cudaFuncSetAttribute(k_route<Loader>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_sweep<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_sweep<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);

    int64_t chunk_begin = 0;
    int g0 = 0;
    const int slots = 2 * kNumSMs;                 // resident sweep CTAs
    while (g0 < B) {
        // largest group the record budget allows ...
        int g_max = g0;
        int64_t cur = 0;
        while (g_max < B) {
            const int64_t c = chunks_of(off, g_max);
            if (cur > 0 && ((size_t)(cur + c) * kBytesPerChunk > budget || cur + c > gchunks)) break;
            cur += c;
            ++g_max;
        }
        // ... then trimmed so that its (sample, band) tasks fill whole waves of the resident sweep CTAs: a group of 12
        // samples x 50 bands = 600 tasks on 296 slots runs 3 rounds (2.03 needed), 11 samples run 2
        int g1 = g_max;
        double best = -1.0;
        for (int cand = g_max; cand > g0 && cand >= g0 + (g_max - g0 + 1) / 2; --cand) {
            const int64_t tasks = (int64_t)(cand - g0) * bp.nb;
            const double eff = (double)tasks / (double)(ceil_div64(tasks, slots) * slots);
            if (eff > best + 1e-9) { best = eff; g1 = cand; }
        }
        cur = 0;
        for (int b = g0; b < g1; ++b) cur += chunks_of(off, b);
        a.g0 = g0; a.g1 = g1;
        a.begin = off[g0]; a.end = off[g1];
        g.chunk_begin = (int)chunk_begin;
        if (cur > 0) {
            profile_begin(st, kProfScatter);
            k_route<Loader><<<(unsigned)cur, kRouteThreads, 0, st>>>(ld, g);
            profile_end(st);
            EP_LAUNCH_CHECK();
        }
        profile_begin(st, kProfFinalize);
        const int n_tasks = (g1 - g0) * bp.nb;
        const int sweep_grid = n_tasks < 2 * kNumSMs ? n_tasks : 2 * kNumSMs;
        if (count) k_sweep<true><<<(unsigned)sweep_grid, kSweepThreads, sweep_smem, st>>>(g, n_tasks);
        else k_sweep<false><<<(unsigned)sweep_grid, kSweepThreads, sweep_smem, st>>>(g, n_tasks);
        profile_end(st);
        EP_LAUNCH_CHECK();
        chunk_begin += cur;
        g0 = g1;
    }
#ifdef EP_PHASE_TIMING
    if (getenv("EP_PRINT_TIMING")) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, g.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[9] = {"route: setup", "route: load+compute+count", "route: scan", "route: place", "route: writeout",
                                "sweep: task setup", "sweep: records", "sweep: flush", "sweep: task tail"};
        double rt = 0, sw = 0;
        for (int i = 0; i < 5; ++i) rt += (double)h[i];
        for (int i = 5; i < 9; ++i) sw += (double)h[i];
        for (int i = 0; i < 9; ++i) fprintf(stderr, "  %-28s %12.3f Mcycles  %5.1f%%\n", names[i], h[i] * 1e-6, 100.0 * h[i] / (i < 5 ? rt : sw));
    }
#endif
    return EP_OK;
}

}  // namespace

size_t banded_workspace_bytes(const ep_events_soa* ev, const ep_bin_params* p) {
    BandPlan bp;
    if (!ev || !ev->offsets_host || ev->batch <= 0 || !plan_bands(p, &bp)) return 0;
    const int64_t gchunks = max_group_chunks(ev->offsets_host, ev->batch, banded_group_budget(), (size_t)1 << 40);
    return band_layout(ev->batch, gchunks).total + kBytesPerChunk + 2048;
}

int run_banded_canon(cudaStream_t st, const ep_events_soa* ev, const ep_bin_params* p, float* out_voxel, float* out_sum,
                     float* out_count, void* ws, size_t ws_bytes, unsigned int* bad) {
    BandPlan bp;
    if (!banded_eligible(ev, p) || !plan_bands(p, &bp)) return EP_EUNSUPPORTED;
    if (ev->offsets_host[ev->batch] - ev->offsets_host[0] > 0x7fffffffLL * (int64_t)kChunk / 2) return EP_EUNSUPPORTED;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u)) return EP_EALIGN;
    if (ev->t_dtype == EP_I64) {
        SoaCanonLoader<true> ld{static_cast<const uint16_t*>(ev->x), static_cast<const uint16_t*>(ev->y), ev->t,
                                static_cast<const uint8_t*>(ev->p), ev->t_div};
        return run_banded(st, ld, ev, p, bp, out_voxel, out_sum, out_count, ws, ws_bytes, bad);
    }
    SoaCanonLoader<false> ld{static_cast<const uint16_t*>(ev->x), static_cast<const uint16_t*>(ev->y), ev->t,
                             static_cast<const uint8_t*>(ev->p), ev->t_div};
    return run_banded(st, ld, ev, p, bp, out_voxel, out_sum, out_count, ws, ws_bytes, bad);
}

}  // namespace ep

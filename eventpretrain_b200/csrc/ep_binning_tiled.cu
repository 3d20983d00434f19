// Stage 1, default fast path — finalize-free, shared-memory-privatised voxel binning (sm_100a).
//
// Replaces events_to_voxel_grid (dataset/dataset_utils/events_to_voxel_grid.py:4-61), the fused events_reshape scale
// (dataset/augmentation/events_augment.py:22-26; the reference's own pre-training order "rescale to 224x224, then bin",
// dataset/pretrain/pr_n_imagenet_dataset.py:85-87) and voxel.sum(0) (dataset/pretrain/pr_ef_imagenet_dataset.py:192-193)
// for a ragged batch in the 4 B/event packed transport layout.  Same integers as the global-RED path of ep_binning.cu
// (Q24 fixed-point weights, one rounding per output element), so the two are bit-identical; that path stays as the
// general one (any layout, count frames, fp64 stamps, grids whose rows do not fit a tile).
//
// Design:
//   * route (k_route): one CTA per 8192-event chunk of one sample.  Each event becomes a 4-byte record
//         [ polarity as a signed 2-bit field (+1 / -1):2 | cell-in-tile:14 | ticks relative to the chunk:16 ]
//     and the chunk's records are bucketed by row tile (a tile = `rows` full rows of the output grid, at most 12800 cells)
//     with one returning shared-memory atomic per event for the rank, staged in shared memory and written back coalesced.
//     Coordinates go through shared-memory look-up tables (fp64 scale and truncation of events_reshape, tile id, row base),
//     out-of-grid events land in a trash bucket that is counted into bad_count and never read.  No weights are computed
//     here: a record stays 4 bytes, so the route moves 4 B in + 4 B out per event.  The next task's event words are
//     requested before the current task's records are copied out.
//   * sweep (k_sweep): persistent CTAs, two per SM; a task is (sample, tile).  Two adjacent output planes of the tile
//     live in shared memory as int32 Q24 sums (100 KB).  The task walks the temporal intervals k = 0 .. bins-1: the events
//     of interval k (found through the per-chunk interval range the route wrote; contiguous for time-sorted streams, any
//     order is handled) add 2^24 - r to plane k and r to plane k + 1 with two shared-memory atomics, then plane k is
//     converted to fp32 (one rounding), written to the output with 16-byte stores, and its buffer is re-zeroed to become
//     plane k + 2.  voxel.sum(0) is formed at the last plane from the earlier planes read back out of L2 (sequential fp32
//     over bins, like the reference's sum; grids with many bins keep a running sum instead).
//     The runs (chunk, tile) of a phase are cut into items of at most 128 records that the warps draw from a shared
//     counter, with the next item's records in flight while the current one is accumulated.
//     Every record is read once; there are no global accumulators, no memset and no finalize pass.
//   * exactness under any distribution: the int32 plane words may wrap (more than 127 same-polarity events on one cell
//     within one interval pair).  Every atomic returns the old value; a wrap is detected when it happens and logged as a
//     +-2^32 correction in a small spill list that the flush applies in 64-bit arithmetic.  Integer adds commute, so the
//     result does not depend on the order of events, warps or CTAs: bit-reproducible.
//   * generality: chunks whose stamps span 2^16 ticks or more keep block-relative ticks in the record ("wide" format),
//     and samples without integer-time constants (deltaT == 0, last row before the first) use the fp64 expression; both
//     take an out-of-line slow path in the sweep.
#include "ep_binning_common.cuh"

namespace ep {

namespace {

constexpr int kChunkShift = 13;
constexpr int kChunk = 1 << kChunkShift;      // events per route task = 32 tick blocks of the 4 B layout
constexpr int kTickBlockShift = 8;            // 256-event tick blocks (SoaPackedLoader<false>)
constexpr int kMaxTiles = 64;
#ifndef EP_SWEEP_THREADS
#define EP_SWEEP_THREADS 512
#define EP_SWEEP_CTAS 2
#define EP_TILE_CELLS 12800
#endif
constexpr int kTileCells = EP_TILE_CELLS;     // cells of a plane tile: two live planes = 100 KB -> two sweep CTAs per SM
constexpr int kRouteThreads = 512;
constexpr int kRouteEv = kChunk / kRouteThreads;   // 16 events per thread, as 4 quads
constexpr int kSweepThreads = EP_SWEEP_THREADS;
constexpr int kSweepCtas = EP_SWEEP_CTAS;     // resident sweep CTAs per SM
constexpr int kTabCap = kSweepThreads;        // chunks of one sample whose run table is resident (one per thread)
constexpr int kItemCap = 384;                 // items (<= 128 records each, 8-byte descriptors) listed per round of a phase
#ifndef EP_ITEM_RPL
#define EP_ITEM_RPL 4
#endif
#ifndef EP_SWEEP_PREDICATED
#define EP_SWEEP_PREDICATED 0                 // measured: predicating the atomics of masked lanes off costs 0.953 -> 0.994 ms (branches)
#endif
#ifndef EP_ITEM_DEPTH
#define EP_ITEM_DEPTH 2
#endif
constexpr int kItemRPL = EP_ITEM_RPL;         // records per lane and item
constexpr int kItemDepth = EP_ITEM_DEPTH;     // items whose records are in flight ahead of the current one (1 or 2)
constexpr int kItemRecs = 32 * kItemRPL;
static_assert(kItemRPL >= 1 && kItemRPL <= 4, "the item descriptor holds records - 1 in 7 bits");
// measured on B200, sweep of the 256-sample batch: 64-record items 1.046 ms, 96 1.008, 128 0.953 (kept)
constexpr int kSpillCap = 128;
constexpr int kPlaneCellsMax = 54272;         // whole-plane kernels: 212 KB of int32 plane + 13 KB of tables and spill list
constexpr int kPlaneMaxTiles = 3;
#ifndef EP_PLANE_THREADS
#define EP_PLANE_THREADS 1024
#endif
constexpr int kPlaneStatSlot = (EP_PLANE_THREADS / 32) * 3;      // doubles per (sample, tile, channel): (sum, sum of squares, max) per warp of k_plane

constexpr uint32_t kChunkFast = 1u;           // integer-tick sample, narrow records, every v of the chunk fits 32 bits
constexpr uint32_t kChunkNarrow = 2u;         // records carry chunk-relative ticks (else: tick block + block-relative ticks)

struct __align__(16) ChunkMeta {
    int64_t cbase;        // ticks: smallest block base of the task minus the sample's first-row ticks
    uint32_t pos0;        // position (in the record array) of the task's first record
    uint8_t klo, khi;     // temporal intervals the task's events can fall in (conservative)
    uint8_t flags, pad;
};
static_assert(sizeof(ChunkMeta) == 16, "ChunkMeta layout");

struct __align__(16) TaskDesc {
    int64_t c0;           // first array position of the chunk
    int b;                // sample
    uint32_t lohi;        // slots [lo, hi) of the chunk that belong to the sample: lo | hi << 16
};

struct TiledArgs {
    const uint32_t* w;          // packed words
    const uint32_t* blk_base;   // tick offset per 256-event block
    const int64_t* offsets;     // device, B+1
    int64_t n_total;            // offsets_host[B]: vector loads stay below it
    int64_t rec_pos0;           // array position of rec[0] (multiple of kChunk)
    int B, H, W, num_bins;
    int NT, rows;               // row tiles per plane, rows per tile
    int rep_shift;              // log2 of the rank-counter replicas per tile in the route (fewer same-address atomics)
    int off_stride;             // NT + 2 bucket offsets per task
    double sx, sy;
    int scaled;
    int n_tasks;                // route tasks (sample, chunk)
    int task_lo, task_hi;       // route tasks of this launch (a group of samples; [0, n_tasks) for the whole batch)
    int sweep_lo, sweep_hi;     // sweep tasks (sample * NT + tile) of this launch
    int counter_idx;            // which of the sweep task counters this launch draws from
    int deferred_sum;           // sweep: voxel.sum(0) at the last plane from the planes read back (else a running sum per plane)
    int count_bad;              // route: out-of-grid events are added to bad_count (0 when another kernel already counted them)
    const unsigned int* run_if; // null, or a device word: every kernel of the path returns at once while it is 0 (the path
                                // then stands by as the fallback of the whole-plane kernels, run_plane_packed4)
    SampleMeta* meta;
    int* first_task;            // B+1
    TaskDesc* desc;             // n_tasks
    ChunkMeta* cmeta;           // n_tasks
    uint16_t* coff;             // n_tasks x off_stride: bucket starts relative to pos0; [NT] = first dropped, [NT+1] = records
    uint32_t* crel;             // n_tasks x 32: block base minus the task's smallest (wide records)
    uint32_t* rec;              // routed records, indexed by array position - rec_pos0
    unsigned int* counters;     // [0] sweep task counter
    unsigned int* bad_count;
    float* out_voxel;
    float* out_sum;
    double* stats_part;         // (B * NT) x (num_bins + 1) x warps x 3 partial statistics, or null
};

__device__ __forceinline__ void pdl_wait_t() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger_t() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- setup: route tasks per sample ---------------------------------------------------------------------------------
// task c of sample b covers the part of array chunk (offsets[b] >> 13) + c that belongs to the sample.
// k_tiled_setup: first_task[b] = tasks of the samples before b (one CTA: a scan over the batch);
// k_tiled_desc: one thread per task finds its sample by bisection and writes the descriptor.
__global__ void __launch_bounds__(1024) k_tiled_setup(TiledArgs a) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (a.run_if && *a.run_if == 0u) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < a.B; b0 += 1024) {
        const int b = b0 + tid;
        int n = 0;
        if (b < a.B) {
            const int64_t lo = a.offsets[b], hi = a.offsets[b + 1];
            if (hi > lo) n = (int)(((hi - 1) >> kChunkShift) - (lo >> kChunkShift)) + 1;
        }
        const int incl = warp_incl_scan(n, lane);
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int v = s_warp[lane];
            const int s = warp_incl_scan(v, lane);
            s_warp[lane] = s - v;
        }
        __syncthreads();
        const int first = s_carry + s_warp[wid] + incl - n;
        if (b < a.B) a.first_task[b] = first;
        __syncthreads();
        if (tid == 1023) s_carry = first + n;
        __syncthreads();
    }
    if (tid == 0) a.first_task[a.B] = s_carry;
}

__global__ void __launch_bounds__(256) k_tiled_desc(TiledArgs a) {
    const int task = blockIdx.x * blockDim.x + threadIdx.x;
    if (task >= a.n_tasks || (a.run_if && *a.run_if == 0u)) return;
    int lo_b = 0, hi_b = a.B;                          // largest b with first_task[b] <= task (empty samples share a value:
    while (hi_b - lo_b > 1) {                          // the last of them is the one that owns tasks)
        const int mid = (lo_b + hi_b) >> 1;
        if (a.first_task[mid] <= task) lo_b = mid; else hi_b = mid;
    }
    const int b = lo_b, c = task - a.first_task[b];
    const int64_t lo = a.offsets[b], hi = a.offsets[b + 1];
    const int64_t c0 = ((lo >> kChunkShift) + c) << kChunkShift;
    TaskDesc d;
    d.c0 = c0; d.b = b;
    d.lohi = (uint32_t)((lo > c0 ? lo : c0) - c0) | ((uint32_t)((hi < c0 + kChunk ? hi : c0 + kChunk) - c0) << 16);
    a.desc[task] = d;
}

// ---- route ---------------------------------------------------------------------------------------------------------
// dynamic shared memory: lut_y[2048] u32 | lut_x[2048] u32 | stage[2][8192] u32 | cnt[2][kMaxTiles + 2] | off[warps][kMaxTiles + 4]
constexpr size_t kRouteSmem = 2048 * 4 + 2048 * 4 + (size_t)2 * kChunk * 4 + 2 * (kMaxTiles + 2) * 4 +
                              (size_t)(kRouteThreads / 32) * (kMaxTiles + 4) * 4 + 64;

// polarity bit of the packed word -> signed 2-bit field of the record: 1 -> 01 (+1), 0 -> 11 (-1)
__device__ __forceinline__ uint32_t pol2(uint32_t word) { return 3u - ((word >> 21) & 2u); }

__device__ __forceinline__ void route_load(const TiledArgs& a, const TaskDesc& d, int tid, uint32_t (&wv)[kRouteEv]) {
    const int slot_lo = (int)(d.lohi & 0xffffu), slot_hi = (int)(d.lohi >> 16);
    if (slot_lo == 0 && slot_hi == kChunk && d.c0 + kChunk <= a.n_total) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 v = ld_stream(reinterpret_cast<const uint4*>(a.w + d.c0) + q * kRouteThreads + tid);
            wv[q * 4 + 0] = v.x; wv[q * 4 + 1] = v.y; wv[q * 4 + 2] = v.z; wv[q * 4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int slot0 = (q * kRouteThreads + tid) * 4;
            const int64_t pos = d.c0 + slot0;
#pragma unroll
            for (int e = 0; e < 4; ++e) wv[q * 4 + e] = 0;
            if (slot0 + 4 > slot_lo && slot0 < slot_hi) {
                if (pos + 4 <= a.n_total) {
                    const uint4 v = ld_stream(reinterpret_cast<const uint4*>(a.w + pos));
                    wv[q * 4 + 0] = v.x; wv[q * 4 + 1] = v.y; wv[q * 4 + 2] = v.z; wv[q * 4 + 3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (pos + e < a.n_total) wv[q * 4 + e] = a.w[pos + e];
                }
            }
        }
    }
}

// Tick-block bases of a task (one warp, lane = block of the chunk) -> s_rel[32], chunk metadata.  Split into the loads
// (issued one phase early, so that the global-memory latency hides behind the bucket ranking) and the arithmetic.
struct BasesRegs {
    int64_t lo_b, t0_ticks;
    uint32_t base, tmul, tshift, thalf, flags;
};

__device__ __forceinline__ void bases_load(const TiledArgs& a, const TaskDesc& d, int lane, BasesRegs& r) {
    const int slot_lo = (int)(d.lohi & 0xffffu), slot_hi = (int)(d.lohi >> 16);
    const bool used = (lane << kTickBlockShift) < slot_hi && ((lane + 1) << kTickBlockShift) > slot_lo;
    r.lo_b = __ldg(a.offsets + d.b);
    r.base = used ? __ldg(a.blk_base + (d.c0 >> kTickBlockShift) + lane) : 0u;
    const SampleMeta* m = a.meta + d.b;          // written by k_sample_meta of this call: plain loads
    r.t0_ticks = m->t0_ticks; r.tmul = m->tmul; r.tshift = m->tshift; r.thalf = m->thalf; r.flags = m->flags;
}

__device__ __forceinline__ void bases_finish(const TiledArgs& a, const TaskDesc& d, int task, int lane, const BasesRegs& r,
                                             uint32_t* s_rel, uint32_t* s_fmt) {
    const int slot_lo = (int)(d.lohi & 0xffffu), slot_hi = (int)(d.lohi >> 16);
    const int64_t blk_id = (d.c0 >> kTickBlockShift) + lane;
    const bool used = (lane << kTickBlockShift) < slot_hi && ((lane + 1) << kTickBlockShift) > slot_lo;
    const uint32_t base = (used && blk_id != (r.lo_b >> kTickBlockShift)) ? r.base : 0u;
    uint32_t mn = used ? base : 0xffffffffu, mx = used ? base : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const uint32_t rel = used ? base - mn : 0u;
    const bool narrow = (mx - mn) < 65536u - 512u;
    s_rel[lane] = narrow ? (rel << 16) : ((uint32_t)lane << 16);      // per-quad addend of the record's high half
    a.crel[(size_t)task * 32 + lane] = rel;
    if (lane == 0) {
        *s_fmt = narrow ? 16u : 21u;                                   // shift of the 9 tick bits inside the record
        const int64_t cbase = (int64_t)mn - r.t0_ticks;
        const int64_t dt_hi = cbase + (int64_t)(mx - mn) + 511;
        const bool int_time = (r.flags & kFlagIntTime) != 0;
        int klo = 0, khi = a.num_bins - 1;
        bool fast = false;
        if (int_time) {
            uint32_t v;
            const uint32_t v_end = (uint32_t)a.num_bins << kQ;
            if (cbase > 0 && ticks_to_v(cbase, r.tmul, r.tshift, r.thalf, v_end, v)) klo = (int)(v >> kQ);
            else if (cbase >= (1ll << 32)) klo = a.num_bins - 1;
            if (dt_hi >= 0 && ticks_to_v(dt_hi, r.tmul, r.tshift, r.thalf, v_end, v)) khi = (int)(v >> kQ);
            else if (dt_hi < 0) khi = 0;
            if (narrow && cbase >= 0 && dt_hi < (1ll << 32)) {
                // every v of the chunk fits 32 bits: the sweep's fast path then needs no range test beyond the interval's
                const uint64_t q = (uint64_t)dt_hi * r.tmul + r.thalf;
                fast = ((q >> 32) >> r.tshift) == 0;
            }
        }
        ChunkMeta cm;
        cm.cbase = cbase;
        cm.pos0 = (uint32_t)(d.c0 + slot_lo - a.rec_pos0);
        cm.klo = (uint8_t)klo; cm.khi = (uint8_t)khi;
        cm.flags = (uint8_t)((fast ? kChunkFast : 0u) | (narrow ? kChunkNarrow : 0u));
        cm.pad = 0;
        a.cmeta[task] = cm;
    }
}

// Two barriers per chunk: [rank the events into buckets] B1 [every warp scans the bucket sizes for itself; records to their
// slots of this chunk's stage buffer] B2 [coalesced copy-out].  Stage buffer, bucket counters and tick-base tables are double
// buffered, so the copy-out of chunk i runs under the ranking of chunk i + 1; the next chunk's descriptor is fetched at the
// top of the iteration, its tick-block bases and sample constants between the barriers (by warp 1, while warp 0 also writes
// the chunk's run table), and its event words as soon as this chunk's are consumed.
constexpr int kOffStride = kMaxTiles + 4;
// TR = transposed tiles for EvRep: the tile and the row base come from x (tiles are column ranges, cells run x-major inside
// a tile: the reference's lexsort order, events_to_image.py:104), the minor coordinate is y and must lie inside the grid
// on its own (numpy's index check), a.H / a.W are then the image's width / height.
// REP: every bucket has 2^rep_shift rank counters, picked by the lane (grids with few tiles, where most lanes of a warp would
// hit the same counter word).  (Tried and dropped: tile / row base by multiply-high instead of the y table, 0.728 -> 0.748 ms.)
template <bool TR, bool REP = false>
__global__ void __launch_bounds__(kRouteThreads, 2) k_route(TiledArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* lut_y = reinterpret_cast<uint32_t*>(smem_raw);                    // tile << 16 | (row in tile) * W
    uint32_t* lut_x = lut_y + 2048;                                             // scaled x
    uint32_t* stage0 = lut_x + 2048;                                            // 2 x kChunk records
    uint32_t* s_cnt0 = stage0 + 2 * kChunk;                                     // 2 x (kMaxTiles + 2): tile << 16 | events so far
    uint32_t* s_off0 = s_cnt0 + 2 * (kMaxTiles + 2);                            // per warp: bucket start - (tile << 16)
    __shared__ uint32_t s_rel[2][32];
    __shared__ uint32_t s_fmt[2];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (a.run_if && *a.run_if == 0u) return;
    const int NT = a.NT, NB = NT + 1;          // bucket NT = trash (coordinates outside the grid)
    // every bucket has 2^rs rank counters, picked by the lane: with few tiles most lanes of a warp would hit the same word
    const uint32_t rs = REP ? (uint32_t)a.rep_shift : 0u, rep = REP ? ((uint32_t)lane & ((1u << rs) - 1u)) : 0u;
    const int NBs = NB << rs;                   // sub-buckets (<= kMaxTiles + 1)
    uint32_t* s_off = s_off0 + wid * kOffStride;
    pdl_trigger_t();
    // events_reshape: x * (input_w / sensor_w) in fp64, truncated by the .long() of the binning call
    for (int i = tid; i < 2048; i += kRouteThreads) {
        const long long yy = a.scaled ? __double2ll_rz(__dmul_rn((double)i, a.sy)) : (long long)i;
        const long long xx = a.scaled ? __double2ll_rz(__dmul_rn((double)i, a.sx)) : (long long)i;
        uint32_t ly = (uint32_t)NT << 16;
        if (yy < a.H) {
            const uint32_t t = (uint32_t)(yy / a.rows);
            // cells of the last (possibly shorter) tile are kept at the END of the plane, so that one compare against the
            // full tile size catches every index that leaves the tile or the grid
            const uint32_t bias = ((int)t == NT - 1) ? (uint32_t)(a.rows * a.W - (a.H - (NT - 1) * a.rows) * a.W) : 0u;
            ly = (t << 16) | ((uint32_t)((yy % a.rows) * a.W) + bias);
        }
        lut_y[i] = ly;
        lut_x[i] = (uint32_t)(xx < 4095 ? xx : 4095);                 // anything >= W is redone exactly in the cold path
    }
    if (tid < 2 * (kMaxTiles + 2)) s_cnt0[tid] = (uint32_t)(tid % (kMaxTiles + 2)) << 16;
    pdl_wait_t();          // the workspace (headers, records) may still be read by the previous call's sweep

    int task = a.task_lo + blockIdx.x;
    if (task >= a.task_hi) return;
    TaskDesc d = a.desc[task];
    uint32_t wv[kRouteEv];
    route_load(a, d, tid, wv);
    if (wid == 1) {
        BasesRegs br;
        bases_load(a, d, lane, br);
        bases_finish(a, d, task, lane, br, s_rel[0], &s_fmt[0]);
    }
    int buf = 0;
    __syncthreads();

    const uint32_t tc = (uint32_t)(a.rows * a.W);                       // cells of a full tile
    const uint32_t last_bias = tc - (uint32_t)((a.H - (NT - 1) * a.rows) * a.W);   // the last tile's cells sit at the end of the plane
    for (;;) {
        const int slot_lo = (int)(d.lohi & 0xffffu), slot_hi = (int)(d.lohi >> 16);
        const bool full = slot_lo == 0 && slot_hi == kChunk;
        const int next = task + gridDim.x;
        const bool more = next < a.task_hi;
        TaskDesc dn = d;
        if (more) dn = a.desc[next];                                    // consumed after the ranking
        uint32_t* stage = stage0 + buf * kChunk;
        uint32_t* s_cnt = s_cnt0 + buf * (kMaxTiles + 2);
        const uint32_t fmt = s_fmt[buf];
        // ---- pass 1: record and bucket of every event of the thread (branch-free), then the ranks ----
        uint32_t rt[kRouteEv];      // tile, then tile << 16 | rank; 0xffffffff = not an event of this task
        bool fix = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t add = s_rel[buf][q * 8 + (tid >> 6)];      // block of slot s = s >> 8 = q * 8 + (tid >> 6)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t word = wv[q * 4 + e];
                const uint32_t fa = TR ? (word & 0x7ffu) : ((word >> 11) & 0x7ffu);      // major coordinate: tile, row base
                const uint32_t fb = TR ? ((word >> 11) & 0x7ffu) : (word & 0x7ffu);      // minor coordinate
                const uint32_t ly = lut_y[fa];
                const uint32_t xx = (!TR && a.scaled) ? lut_x[fb] : fb;
                const uint32_t cell = (ly & 0xffffu) + xx;
                if (TR) {
                    rt[q * 4 + e] = (((fb < (uint32_t)a.W) ? (ly >> 16) : (uint32_t)NT) << rs) + rep;
                } else {
                    // x + y * W has no bound on x alone in the reference (events_to_voxel_grid.py:46): an x >= W stays correct while
                    // the sum stays inside the tile; anything that leaves it is redone exactly below (rare)
                    fix |= cell >= tc;
                    rt[q * 4 + e] = ((ly >> 16) << rs) + rep;
                }
                wv[q * 4 + e] = pol2(word) + (cell << 2) + ((word >> 23) << fmt) + add;
            }
        }
        if (!TR && fix) {
            // cold: redo the thread's 16 events from the source words with the reference's own index arithmetic
#pragma unroll 1
            for (int i = 0; i < kRouteEv; ++i) {
                const int q = i >> 2, e = i & 3;
                const int64_t pos = d.c0 + (q * kRouteThreads + tid) * 4 + e;
                const uint32_t word = pos < a.n_total ? a.w[pos] : 0u;
                const uint32_t x0 = word & 0x7ffu, y0 = (word >> 11) & 0x7ffu;
                const int64_t xs = a.scaled ? __double2ll_rz(__dmul_rn((double)x0, a.sx)) : (int64_t)x0;
                const int64_t ys = a.scaled ? __double2ll_rz(__dmul_rn((double)y0, a.sy)) : (int64_t)y0;
                uint32_t tile = (uint32_t)NT, cell = 0;
                if (ys < a.H + 65536 && xs < (1ll << 40)) {
                    const int64_t flat = ys * a.W + xs;
                    if (flat < (int64_t)a.H * a.W) {
                        const uint32_t y2 = (uint32_t)(flat / a.W);
                        tile = y2 / (uint32_t)a.rows;
                        cell = (y2 % (uint32_t)a.rows) * (uint32_t)a.W + (uint32_t)(flat - (int64_t)y2 * a.W);
                        if ((int)tile == NT - 1) cell += last_bias;
                    }
                }
                const uint32_t addq = s_rel[buf][q * 8 + (tid >> 6)];
                const uint32_t rec = pol2(word) + (cell << 2) + ((word >> 23) << fmt) + addq;
                // (static indexing: the arrays stay in registers)
#pragma unroll
                for (int j = 0; j < kRouteEv; ++j) if (j == i) { wv[j] = rec; rt[j] = (tile << rs) + rep; }
            }
        }
        if (full) {
#pragma unroll
            for (int i = 0; i < kRouteEv; ++i) rt[i] = atomicAdd(&s_cnt[rt[i]], 1u);
        } else {
#pragma unroll
            for (int i = 0; i < kRouteEv; ++i) {
                const int sl = ((i >> 2) * kRouteThreads + tid) * 4 + (i & 3);
                rt[i] = (sl >= slot_lo && sl < slot_hi) ? atomicAdd(&s_cnt[rt[i]], 1u) : 0xffffffffu;
            }
        }
        __syncthreads();                                                // B1: the bucket sizes are final
        // ---- bucket starts: every warp scans the NB <= 65 bucket sizes for itself (no serial section, no barrier) ----
        {
            uint32_t c[3], sum = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) { const int t = lane * 3 + j; c[j] = (t < NBs) ? (s_cnt[t] & 0xffffu) : 0u; sum += c[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
            uint32_t run = incl - sum;
            uint16_t* co = a.coff + (size_t)task * a.off_stride;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int t = lane * 3 + j;
                if (t <= NBs) {
                    s_off[t] = run - ((uint32_t)t << 16);
                    if (wid == 0 && (t & ((1 << rs) - 1)) == 0) co[t >> rs] = (uint16_t)run;      // a tile's records start at its first sub-bucket
                }
                run += c[j];
            }
            if (wid == 0 && a.bad_count && a.count_bad) {
                // events outside the grid: the sub-buckets of the trash bucket
                const int t0 = NT << rs;
                uint32_t nbad = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) { const int t = lane * 3 + j; if (t >= t0 && t < NBs) nbad += c[j]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) nbad += __shfl_xor_sync(0xffffffffu, nbad, o);
                if (lane == 0 && nbad) atomicAdd(a.bad_count, nbad);
            }
        }
        BasesRegs br;
        if (wid == 1 && more) bases_load(a, dn, lane, br);
        // the other counter set was last read by the scans of the previous chunk (all warps are past them): reset for the next
        if (wid == 2 || wid == 3 || wid == 4) {
            const int t = tid - 64;
            if (t < kMaxTiles + 2) s_cnt0[(buf ^ 1) * (kMaxTiles + 2) + t] = (uint32_t)t << 16;
        }
        __syncwarp();
        // ---- pass 2: records to their bucket slots ----
        if (full) {
#pragma unroll
            for (int i = 0; i < kRouteEv; ++i) stage[s_off[rt[i] >> 16] + rt[i]] = wv[i];
        } else {
#pragma unroll
            for (int i = 0; i < kRouteEv; ++i)
                if (rt[i] != 0xffffffffu) stage[s_off[rt[i] >> 16] + rt[i]] = wv[i];
        }
        // ---- the next task's words are requested now and land while this task's records are copied out ----
        if (more) {
            route_load(a, dn, tid, wv);
            if (wid == 1) bases_finish(a, dn, next, lane, br, s_rel[buf ^ 1], &s_fmt[buf ^ 1]);
        }
        __syncthreads();                                                // B2: the stage buffer is complete
        // ---- coalesced copy-out (no barrier behind it: this buffer is written again two chunks later, past B1 of the next) ----
        const int n = slot_hi - slot_lo;
        const int64_t p0 = d.c0 + slot_lo - a.rec_pos0;
        uint32_t* dst = a.rec + p0;
        if ((p0 & 3) == 0) {
            const int n4 = n >> 2;
            for (int i = tid; i < n4; i += kRouteThreads)
                reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(stage)[i];
            for (int i = (n4 << 2) + tid; i < n; i += kRouteThreads) dst[i] = stage[i];
        } else {
            for (int i = tid; i < n; i += kRouteThreads) dst[i] = stage[i];
        }
        if (!more) break;
        task = next; d = dn; buf ^= 1;
    }
}

// ---- sweep ---------------------------------------------------------------------------------------------------------
// dynamic shared memory: planes[2][tile_cells] int32 | t_q[kTabCap] u64 | items[kItemCap] uint2 | spill[kSpillCap] int2 |
//                        t_pos[kTabCap] u32 | t_len[kTabCap] u16 | t_kr[kTabCap] u16
__host__ __device__ inline size_t sweep_smem_bytes(int tile_cells) {
    return (size_t)2 * tile_cells * 4 + (size_t)kTabCap * (8 + 4 + 2 + 2) + (size_t)kItemCap * 8 + kSpillCap * 8 + 16;
}

// a wrapped int32 plane word: add a +-2^32 correction for (cell, plane) to the spill list.  Entries start as {-1, 0}
// (reset per task); a slot's key is published after its first correction went in with an atomic, so concurrent wraps of
// the same word either find the entry or open a duplicate, and the flush sums every entry that matches.
__device__ __noinline__ void note_wrap(int old, int w, uint32_t key, int2* spill, int* n_spill, unsigned int* bad) {
    const uint32_t nw = (uint32_t)old + (uint32_t)w;
    if ((int)(((uint32_t)old ^ nw) & ((uint32_t)w ^ nw)) >= 0) return;
    const int d = w > 0 ? 1 : -1;
    int n = *reinterpret_cast<volatile int*>(n_spill);
    if (n > kSpillCap) n = kSpillCap;
    for (int i = 0; i < n; ++i)
        if (reinterpret_cast<volatile int2*>(spill)[i].x == (int)key) { atomicAdd(&spill[i].y, d); return; }
    const int i = atomicAdd(n_spill, 1);
    if (i < kSpillCap) {
        atomicAdd(&spill[i].y, d);
        __threadfence_block();
        reinterpret_cast<volatile int2*>(spill)[i].x = (int)key;
    } else if (bad) {
        atomicOr(bad, 0x80000000u);
    }
}

struct SweepCtx {
    uint32_t pl0s, pl1s;  // shared-memory addresses of plane k (left node of interval k) and plane k+1 (right node)
    int2* spill;
    int* n_spill;
    unsigned int* bad;
    uint32_t kbase;       // k << 24
    int k;
};

// returning shared-memory atomic add on a 32-bit shared address
__device__ __forceinline__ int atoms_add(uint32_t saddr, int w) {
    int old;
    asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(w) : "memory");
    return old;
}
#if EP_SWEEP_PREDICATED
// the same, predicated off (returning 0) where ok is false: lanes past the end of an item and records of another interval
// then take no bank of the shared-memory pipe
__device__ __forceinline__ int atoms_add_if(bool ok, uint32_t saddr, int w) {
    int old;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tmov.s32 %0, 0;\n\t@p atom.shared.add.s32 %0, [%1], %2;\n\t}"
                 : "=&r"(old) : "r"(saddr), "r"(w), "r"((int)ok) : "memory");
    return old;
}
#endif
__device__ __forceinline__ int sgn2(uint32_t rec) {              // signed 2-bit polarity field
    int r;
    asm("bfe.s32 %0, %1, 0, 2;" : "=r"(r) : "r"(rec));
    return r;
}
__device__ __forceinline__ bool near_wrap(int old) { return (uint32_t)old + 0x7f000000u >= 0xfe000000u; }

// general time arithmetic of one record (samples without the integer-tick constants, chunks whose dt leaves [0, 2^32)):
// the same expressions as voxel_weights() for the tick layouts
__device__ __forceinline__ bool v_general(int64_t dt, const SampleMeta& m, int num_bins, uint32_t& v) {
    if (m.flags & kFlagIntTime) return ticks_to_v(dt, m.tmul, m.tshift, m.thalf, (uint32_t)num_bins << kQ, v);
    const double ts = (double)dt * m.scale_raw;
    const double tis = floor(ts);
    if (!(tis >= 0.0 && tis < (double)num_bins)) return false;
    const float d = (float)(ts - tis);
    v = ((uint32_t)(int)tis << kQ) + (uint32_t)__float2int_rn(d * 16777216.0f);
    return true;
}

// slow path of one record: wide records and / or general time arithmetic
__device__ __noinline__ void sweep_record_slow(const ChunkMeta* cmp, const uint32_t* crel, int num_bins, SweepCtx c,
                                               const SampleMeta* mp, uint32_t r, bool has_right) {
    const SampleMeta m = *mp;
    const ChunkMeta cm = *cmp;
    int64_t dt = cm.cbase;
    if (cm.flags & kChunkNarrow) dt += (int64_t)(r >> 16);
    else dt += (int64_t)crel[(r >> 16) & 31u] + (int64_t)(r >> 21);
    uint32_t v = 0;
    if (!v_general(dt, m, num_bins, v)) return;
    const uint32_t u = v - c.kbase;
    if (u >= (1u << kQ)) return;
    const int sgn = sgn2(r);
    const uint32_t boff = r & 0xfffcu;
    const int wr = (int)u * sgn, wl = (sgn << kQ) - wr;
    const int old0 = atoms_add(c.pl0s + boff, wl);
    if (near_wrap(old0)) note_wrap(old0, wl, (boff >> 2) | ((uint32_t)c.k << 16), c.spill, c.n_spill, c.bad);
    if (has_right) {
        const int old1 = atoms_add(c.pl1s + boff, wr);
        if (near_wrap(old1)) note_wrap(old1, wr, (boff >> 2) | ((uint32_t)(c.k + 1) << 16), c.spill, c.n_spill, c.bad);
    }
}

// One plane of the tile: int32 Q24 words -> fp32 output (one rounding), running voxel.sum(0) in the output buffer itself
// (read back through L2, the same thread wrote it in the previous phase), plane words re-zeroed.  The loop is cut into
// groups of kFlushUnroll vectors whose shared-memory and L2 loads are all issued before the first use: a thread owns at
// most 7 vectors of a 12800-cell tile, and one dependent L2 round trip per vector was the longest stall of the sweep.
constexpr int kFlushUnroll = 4;
constexpr int kStatSlices = 32;                          // CTAs per channel of the statistics reduction
constexpr int kStatSlot = (kSweepThreads / 32) * 3;      // doubles per (task, channel): one (sum, sum of squares, max) per warp

// Per-thread partial statistics of the values a thread writes (fixed element order => reproducible)
struct StatAcc {
    float s1, s2, mx;
    __device__ __forceinline__ void init() { s1 = 0.f; s2 = 0.f; mx = -INFINITY; }
    __device__ __forceinline__ void add(float v) { s1 += v; s2 = fmaf(v, v, s2); mx = fmaxf(mx, v); }
};

// voxel.sum(0) is formed at the LAST plane only (EP_DEFERRED_SUM, default): the earlier planes of the same cells are read back
// from the output tensor (written a few phases ago by this very thread with L2-resident stores), added in the reference's
// order ((p0 + p1) + p2 ...) and the sum is written once.  The flushes of the other planes then consist of stores only —
// nothing to wait for — where the running-sum form paid an L2 round trip per plane.  Same fp32 additions: bit-identical.
// Chosen per call (TiledArgs::deferred_sum) when the planes waiting for the read-back fit L2; many-bin grids keep the
// running sum (DEF = false), whose traffic per plane is one tile in and one out whatever the number of bins.

template <bool VEC, bool FIRST, bool LAST, bool SUM, bool STATS, bool DEF>
__device__ __forceinline__ void flush_plane_t(int* pl, int ncell, int k, float* __restrict__ o, float* __restrict__ so, int64_t plane_stride,
                                              const int2* spill, int n_spill, int key0, StatAcc& sa, StatAcc& ss) {
    constexpr float kInv = 1.0f / 16777216.0f;
    constexpr bool kDeferred = DEF;
    constexpr bool kReadBack = kDeferred && SUM && LAST && !FIRST;        // this flush adds planes 0 .. k-1 from the output
    constexpr bool kRunning = !kDeferred && SUM;                         // the running-sum form
    if (VEC) {
        constexpr int kStep = kSweepThreads * 4;
        constexpr int kU = kReadBack ? 2 : kFlushUnroll;
        for (int i0 = threadIdx.x * 4; i0 < ncell; i0 += kStep * kU) {
            int4 q[kU];
            float4 s[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * kStep;
                q[u] = make_int4(0, 0, 0, 0);
                s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < ncell) {
                    q[u] = *reinterpret_cast<const int4*>(pl + i);
                    if (kRunning && !FIRST) s[u] = __ldcg(reinterpret_cast<const float4*>(so + i));
                }
            }
            if (kReadBack) {
                for (int j0 = 0; j0 < k; j0 += 4) {
                    float4 v[kU][4];
#pragma unroll
                    for (int u = 0; u < kU; ++u)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int i = i0 + u * kStep;
                            v[u][jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (i < ncell && j0 + jj < k) v[u][jj] = __ldcg(reinterpret_cast<const float4*>(o + (int64_t)(j0 + jj - k) * plane_stride + i));
                        }
#pragma unroll
                    for (int u = 0; u < kU; ++u)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (j0 + jj < k) { s[u].x += v[u][jj].x; s[u].y += v[u][jj].y; s[u].z += v[u][jj].z; s[u].w += v[u][jj].w; }
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * kStep;
                if (i < ncell) {
                    *reinterpret_cast<int4*>(pl + i) = make_int4(0, 0, 0, 0);
                    float4 f = make_float4(__int2float_rn(q[u].x) * kInv, __int2float_rn(q[u].y) * kInv,
                                           __int2float_rn(q[u].z) * kInv, __int2float_rn(q[u].w) * kInv);
                    if (n_spill) {
                        const int qi[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
                        float* fp = reinterpret_cast<float*>(&f);
                        for (int j = 0; j < 4; ++j) {
                            const int key = (key0 + i + j) | (k << 16);
                            long long hi = 0;
                            for (int t = 0; t < n_spill; ++t) if (spill[t].x == key) hi += spill[t].y;
                            if (hi) fp[j] = __ll2float_rn((hi << 32) + (long long)qi[j]) * kInv;
                        }
                    }
                    // planes that the last flush reads back stay in L2 (plain store); everything else streams out
                    if (kDeferred && SUM && !LAST) __stcg(reinterpret_cast<float4*>(o + i), f);
                    else st_stream(reinterpret_cast<float4*>(o + i), f);
                    if (STATS) { sa.add(f.x); sa.add(f.y); sa.add(f.z); sa.add(f.w); }
                    if (kRunning || (kDeferred && SUM && LAST)) {
                        float4 t = s[u];
                        t.x += f.x; t.y += f.y; t.z += f.z; t.w += f.w;      // voxel.sum(dim=0): sequential fp32 over bins
                        if (LAST) st_stream(reinterpret_cast<float4*>(so + i), t);
                        else __stcg(reinterpret_cast<float4*>(so + i), t);
                        if (STATS && LAST) { ss.add(t.x); ss.add(t.y); ss.add(t.z); ss.add(t.w); }
                    }
                }
            }
        }
    } else {
        for (int i0 = threadIdx.x; i0 < ncell; i0 += kSweepThreads * kFlushUnroll) {
            int q[kFlushUnroll];
            float s[kFlushUnroll];
#pragma unroll
            for (int u = 0; u < kFlushUnroll; ++u) {
                const int i = i0 + u * kSweepThreads;
                q[u] = 0; s[u] = 0.f;
                if (i < ncell) {
                    q[u] = pl[i];
                    if (kRunning && !FIRST) s[u] = __ldcg(so + i);
                    if (kReadBack)
                        for (int j = 0; j < k; ++j) s[u] += __ldcg(o + (int64_t)(j - k) * plane_stride + i);
                }
            }
#pragma unroll
            for (int u = 0; u < kFlushUnroll; ++u) {
                const int i = i0 + u * kSweepThreads;
                if (i < ncell) {
                    pl[i] = 0;
                    float f = __int2float_rn(q[u]) * kInv;
                    if (n_spill) {
                        const int key = (key0 + i) | (k << 16);
                        long long hi = 0;
                        for (int t = 0; t < n_spill; ++t) if (spill[t].x == key) hi += spill[t].y;
                        if (hi) f = __ll2float_rn((hi << 32) + (long long)q[u]) * kInv;
                    }
                    if (kDeferred && SUM && !LAST) __stcg(o + i, f);
                    else st_stream(o + i, f);
                    if (STATS) sa.add(f);
                    if (kRunning || (kDeferred && SUM && LAST)) {
                        const float t = s[u] + f;
                        if (LAST) st_stream(so + i, t); else __stcg(so + i, t);
                        if (STATS && LAST) ss.add(t);
                    }
                }
            }
        }
    }
}

// the same in fp64 (whole-plane kernels: a thread owns up to ~50 values of a plane, hot cells included)
struct StatAcc64 {
    double s1, s2;
    float mx;
    __device__ __forceinline__ void init() { s1 = 0.0; s2 = 0.0; mx = -INFINITY; }
    __device__ __forceinline__ void add(float v) { const double d = (double)v; s1 += d; s2 = fma(d, d, s2); mx = fmaxf(mx, v); }
};

// per-warp reduction of the per-thread statistics (fixed shuffle tree) -> part[warp][0..2] = (sum, sum of squares, max) as
// fp64; no barrier: every (task, channel, warp) slot is written exactly once and k_stats_reduce adds them in a fixed order
template <class Acc>
__device__ __forceinline__ void stat_reduce_store(Acc a, double* part) {
    double d1 = (double)a.s1, d2 = (double)a.s2;
    float mx = a.mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        double* q = part + (threadIdx.x >> 5) * 3;
        q[0] = d1; q[1] = d2; q[2] = (double)mx;
    }
}

// pl points at the tile's first cell inside the plane buffer; key0 = that cell's index in the buffer (spill keys).
// stats_part (or null): per-(task, channel, warp) partial statistics of the written values, channel num_bins = the sum plane.
template <bool VEC, bool DEF>
__device__ __forceinline__ void flush_plane_d(int* pl, int ncell, int k, int num_bins, float* o, float* so, int64_t plane_stride,
                                              const int2* spill, int n_spill, int key0, double* stats_part) {
    const bool first = k == 0, last = k == num_bins - 1;
    StatAcc sa, ss;
    sa.init(); ss.init();
#define EP_FLUSH(F, L, S, T) flush_plane_t<VEC, F, L, S, T, DEF>(pl, ncell, k, o, so, plane_stride, spill, n_spill, key0, sa, ss)
    if (stats_part) {
        if (!so) EP_FLUSH(false, false, false, true);
        else if (first && last) EP_FLUSH(true, true, true, true);
        else if (first) EP_FLUSH(true, false, true, true);
        else if (last) EP_FLUSH(false, true, true, true);
        else EP_FLUSH(false, false, true, true);
        stat_reduce_store(sa, stats_part + (size_t)k * kStatSlot);
        if (so && last) stat_reduce_store(ss, stats_part + (size_t)num_bins * kStatSlot);
    } else {
        if (!so) EP_FLUSH(false, false, false, false);
        else if (first && last) EP_FLUSH(true, true, true, false);
        else if (first) EP_FLUSH(true, false, true, false);
        else if (last) EP_FLUSH(false, true, true, false);
        else EP_FLUSH(false, false, true, false);
    }
#undef EP_FLUSH
}

struct ItemRegs {
    uint2 d;             // .x = first record, .y = chunk slot | (records - 1) << 9 | slow << 16
    uint32_t r[kItemRPL];
};

// The items (<= kItemRecs records of one run) of a phase are dealt to the warps round-robin (they cost the same but for run
// tails, and a shared counter would put an atomic and a shuffle on every item); the records of the next two items are in
// flight while the current one is accumulated.
// Fast path per record: v = (Q + ticks * tmul) >> tshift, two returning atomics, no branch (events of other intervals and
// lanes past the end add zero); the returned values of the item's atomics are screened together for words that came
// near the int32 range (rare -> exact check, spill list).
template <bool HAS_RIGHT>
__device__ __forceinline__ void sweep_items(const TiledArgs& a, const SweepCtx& c, const SampleMeta* mp, uint32_t tmul, uint32_t tshift,
                                            int task0, const unsigned long long* t_q, const uint2* items, int n_items,
                                            int wid, int lane) {
    constexpr int kWarps = kSweepThreads / 32;
    auto load = [&](int it, ItemRegs& g) {
        g.d = items[it];
        const uint32_t cnt = ((g.d.y >> 9) & 127u) + 1u;
        const uint32_t* p = a.rec + g.d.x + lane;
        // lanes past the end of the item get a harmless record of their own (cell = lane): its weights are forced to 0 below
#pragma unroll
        for (int j = 0; j < kItemRPL; ++j) g.r[j] = ((uint32_t)lane + 32u * j < cnt) ? ld_stream(p + 32 * j) : (uint32_t)lane << 2;
    };
    ItemRegs n1, n2;
    n1.d = n2.d = make_uint2(0u, 0u);
#pragma unroll
    for (int j = 0; j < kItemRPL; ++j) n1.r[j] = n2.r[j] = 0u;
    int it = wid;
    if (it < n_items) load(it, n1);
    if (kItemDepth > 1 && it + kWarps < n_items) load(it + kWarps, n2);
    for (; it < n_items; it += kWarps) {
        const ItemRegs cur = n1;
        if (kItemDepth > 1) {
            n1 = n2;
            if (it + 2 * kWarps < n_items) load(it + 2 * kWarps, n2);
        } else {
            if (it + kWarps < n_items) load(it + kWarps, n1);
        }
        const uint32_t slot = cur.d.y & 511u;
        const uint32_t cnt = ((cur.d.y >> 9) & 127u) + 1u;
        if (!(cur.d.y >> 16)) {
            const uint64_t Q = t_q[slot];                                    // cbase * tmul + thalf
            int o0[kItemRPL], o1[kItemRPL];
            uint32_t u[kItemRPL];
#pragma unroll
            for (int j = 0; j < kItemRPL; ++j) {
                const uint64_t q = (uint64_t)(cur.r[j] >> 16) * tmul + Q;
                const uint32_t v = __funnelshift_r((uint32_t)q, (uint32_t)(q >> 32), tshift);
                u[j] = v - c.kbase;
                const bool ok = (uint32_t)lane + 32u * j < cnt && u[j] < (1u << kQ);
                const int sgn = ok ? sgn2(cur.r[j]) : 0;
                const uint32_t boff = cur.r[j] & 0xfffcu;                    // cell * 4
                const int wr = (int)u[j] * sgn, wl = (sgn << kQ) - wr;
#if EP_SWEEP_PREDICATED
                o0[j] = atoms_add_if(ok, c.pl0s + boff, wl);
                o1[j] = HAS_RIGHT ? atoms_add_if(ok, c.pl1s + boff, wr) : 0;
#else
                o0[j] = atoms_add(c.pl0s + boff, wl);
                o1[j] = HAS_RIGHT ? atoms_add(c.pl1s + boff, wr) : 0;
#endif
            }
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < kItemRPL; ++j) {
                m = max(m, (uint32_t)o0[j] + 0x7f000000u);
                if (HAS_RIGHT) m = max(m, (uint32_t)o1[j] + 0x7f000000u);
            }
            if (m >= 0xfe000000u) {
#pragma unroll
                for (int j = 0; j < kItemRPL; ++j) {
                    const bool ok = (uint32_t)lane + 32u * j < cnt && u[j] < (1u << kQ);
                    const int sgn = ok ? sgn2(cur.r[j]) : 0;
                    const uint32_t cell = (cur.r[j] & 0xfffcu) >> 2;
                    const int wr = (int)u[j] * sgn, wl = (sgn << kQ) - wr;
                    if (near_wrap(o0[j])) note_wrap(o0[j], wl, cell | ((uint32_t)c.k << 16), c.spill, c.n_spill, c.bad);
                    if (HAS_RIGHT && near_wrap(o1[j])) note_wrap(o1[j], wr, cell | ((uint32_t)(c.k + 1) << 16), c.spill, c.n_spill, c.bad);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < kItemRPL; ++j)
                if ((uint32_t)lane + 32u * j < cnt)
                    sweep_record_slow(a.cmeta + task0 + slot, a.crel + (size_t)(task0 + slot) * 32, a.num_bins, c, mp, cur.r[j], HAS_RIGHT);
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <bool VEC, bool DEF>
__global__ void __launch_bounds__(kSweepThreads, kSweepCtas) k_sweep(TiledArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile_cells = a.rows * a.W;
    int* plane0 = reinterpret_cast<int*>(smem_raw);
    int* plane1 = plane0 + tile_cells;
    unsigned long long* t_q = reinterpret_cast<unsigned long long*>(plane1 + tile_cells);
    uint2* s_items = reinterpret_cast<uint2*>(t_q + kTabCap);
    int2* s_spill = reinterpret_cast<int2*>(s_items + kItemCap);
    uint32_t* t_pos = reinterpret_cast<uint32_t*>(s_spill + kSpillCap);
    uint16_t* t_len = reinterpret_cast<uint16_t*>(t_pos + kTabCap);
    uint16_t* t_kr = t_len + kTabCap;
    __shared__ int s_task, s_nitems, s_nspill;
    __shared__ SampleMeta s_meta;

    const int tid = threadIdx.x, lane = tid & 31;
    if (a.run_if && *a.run_if == 0u) return;
    pdl_trigger_t();
    for (int i = tid; i < 2 * tile_cells; i += kSweepThreads) plane0[i] = 0;
    if (tid < kSpillCap) s_spill[tid] = make_int2(-1, 0);
    if (tid == 0) { s_nspill = 0; s_nitems = 0; }
    pdl_wait_t();          // records and headers of the route
    const int n_sweep = a.sweep_hi;
    unsigned int* counter = a.counters + a.counter_idx;
    const int64_t HW = (int64_t)a.H * a.W;
    if (tid == 0) s_task = a.sweep_lo + (int)atomicAdd(counter, 1u);
    __syncthreads();

    for (;;) {
        const int task = s_task;
        if (task >= n_sweep) break;
        const int b = task / a.NT, tile = task - b * a.NT;
        const int first = a.first_task[b], nch = a.first_task[b + 1] - first;
        if (tid == 0) s_meta = a.meta[b];
        const int row0 = tile * a.rows;
        const int nrows = (a.H - row0 < a.rows) ? a.H - row0 : a.rows;
        const int ncell = nrows * a.W;
        const int cell0 = tile_cells - ncell;              // a shorter last tile sits at the end of the plane (k_route's lut_y)
        const bool resident = nch <= kTabCap;
        __syncthreads();       // s_task is read by everyone before the last phase overwrites it
        const uint32_t tmul = s_meta.tmul, tshift = s_meta.tshift, thalf = s_meta.thalf;

        for (int k = 0; k < a.num_bins; ++k) {
            SweepCtx c;
            int* pl0 = (k & 1) ? plane1 : plane0;
            c.pl0s = (uint32_t)__cvta_generic_to_shared(pl0);
            c.pl1s = (uint32_t)__cvta_generic_to_shared((k & 1) ? plane0 : plane1);
            c.spill = s_spill; c.n_spill = &s_nspill; c.bad = a.bad_count;
            c.kbase = (uint32_t)k << kQ; c.k = k;
            const bool has_right = k + 1 < a.num_bins;
            // last phase: the next task is drawn now, its number is looked at after the accumulation
            int next_task = 0;
            if (!has_right && tid == kSweepThreads - 1) next_task = a.sweep_lo + (int)atomicAdd(counter, 1u);
            for (int c_round = 0; c_round < nch; c_round += kTabCap) {
                const int ci = c_round + tid;
                if (!resident || k == 0) {
                    // run table of the round: one chunk per thread
                    if (ci < nch) {
                        const ChunkMeta cm = a.cmeta[first + ci];
                        const uint16_t* co = a.coff + (size_t)(first + ci) * a.off_stride + tile;
                        const uint32_t o0 = co[0], o1 = co[1];
                        t_pos[tid] = cm.pos0 + o0;
                        t_q[tid] = (unsigned long long)(uint32_t)cm.cbase * tmul + thalf;
                        t_len[tid] = (uint16_t)((o1 - o0) | ((cm.flags & kChunkFast) ? 0u : 0x8000u));
                        t_kr[tid] = (uint16_t)(cm.klo | (cm.khi << 8));
                    } else {
                        t_len[tid] = 0; t_kr[tid] = 0xff;       // klo = 255: never qualifies
                    }
                    __syncthreads();
                }
                // items of this phase: runs whose chunk can hold events of interval k, cut into pieces of 128 records
                int remaining = 0, cursor = 0;
                {
                    const uint32_t kr = t_kr[tid];
                    const int len = t_len[tid] & 0x7fff;
                    if ((int)(kr & 0xffu) <= k && k <= (int)(kr >> 8)) remaining = (len + kItemRecs - 1) / kItemRecs;
                }
                for (;;) {
                    // s_nitems is 0 here (reset behind the barrier that follows the accumulation)
                    if (remaining) {
                        const int base = atomicAdd(&s_nitems, remaining);
                        int take = kItemCap - base;
                        take = take < 0 ? 0 : (take > remaining ? remaining : take);
                        const uint32_t lenf = t_len[tid], len = lenf & 0x7fffu, pos = t_pos[tid];
                        for (int j = 0; j < take; ++j) {
                            const uint32_t start = (uint32_t)(cursor + j) * kItemRecs;
                            const uint32_t cnt = len - start < (uint32_t)kItemRecs ? len - start : (uint32_t)kItemRecs;
                            s_items[base + j] = make_uint2(pos + start, (uint32_t)tid | ((cnt - 1u) << 9) | ((lenf >> 15) << 16));
                        }
                        cursor += take; remaining -= take;
                    }
                    const int any_left = __syncthreads_or(remaining > 0);
                    const int n_items = s_nitems < kItemCap ? s_nitems : kItemCap;
                    if (has_right) sweep_items<true>(a, c, &s_meta, tmul, tshift, first + c_round, t_q, s_items, n_items, tid >> 5, lane);
                    else sweep_items<false>(a, c, &s_meta, tmul, tshift, first + c_round, t_q, s_items, n_items, tid >> 5, lane);
                    __syncthreads();
                    if (tid == 0) s_nitems = 0;
                    if (!any_left) break;
                    __syncthreads();       // rare: more items than the list holds
                }
                if (!resident) __syncthreads();       // the next round rewrites the run table and the item counter
            }
            // The records of the next phase are requested into L2 now and arrive under the flush: the runs of the chunks that
            // can hold events of interval k + 1 (those that also fed interval k were just read).
            if (resident && has_right) {
                const uint32_t kr = t_kr[tid];
                if ((int)(kr & 0xffu) == k + 1) {
                    const uint32_t* p = a.rec + t_pos[tid];
                    const int len = t_len[tid] & 0x7fff;
                    for (int i = 0; i < len; i += 32) prefetch_l2(p + i);
                }
            }
            // Last phase: publish the next task (drawn before the accumulation, so the atomic's latency is already paid).
            // (Requesting the next task's chunk headers into L2 here was measured: no gain, and the dependent first_task
            // loads make this warp late for the flush.)
            if (!has_right && tid == kSweepThreads - 1) s_task = next_task;
            // plane k is complete: fp32 out (+ running voxel.sum(0)), buffer re-zeroed -> plane k + 2
            const int n_spill = s_nspill < kSpillCap ? s_nspill : kSpillCap;
            float* o = a.out_voxel + ((int64_t)b * a.num_bins + k) * HW + (int64_t)row0 * a.W;
            float* so = a.out_sum ? a.out_sum + (int64_t)b * HW + (int64_t)row0 * a.W : nullptr;
            double* sp = a.stats_part ? a.stats_part + (size_t)task * (a.num_bins + 1) * kStatSlot : nullptr;
            flush_plane_d<VEC, DEF>(pl0 + cell0, ncell, k, a.num_bins, o, so, HW, s_spill, n_spill, cell0, sp);
            if (!has_right && n_spill) {
                // the spill list is per task
                __syncthreads();
                if (tid < kSpillCap) s_spill[tid] = make_int2(-1, 0);
                if (tid == 0) s_nspill = 0;
            }
            __syncthreads();
        }
    }
}

// rank-counter replicas per bucket of the route: as many (a power of two, 4 or 8) as keep (tiles + trash) * replicas + 1
// within the kMaxTiles + 2 counter words
static int route_rep_shift(int NT) {
    // measured on B200: 224x224 (4 tiles, 8 replicas) route 0.93 -> 0.76 ms per 255 M events; 640x480 (24 tiles, 2 replicas)
    // 0.69 -> 0.74 (the lanes already spread over 25 words; two more scan entries per lane cost more): replicate from 4 up only
    int rs = 0;
    while (rs < 3 && (((NT + 1) << (rs + 1)) + 1) <= kMaxTiles + 2) ++rs;
    return rs >= 2 ? rs : 0;
}

struct TiledPlan {
    int NT, rows, n_tasks;
    int64_t rec_pos0, n_rec;
    size_t off_meta, off_first, off_desc, off_cmeta, off_coff, off_crel, off_counters, off_stats, off_stats2, off_rec, total;
    bool plane_ok;              // the grid's plane fits one SM's shared memory in at most 3 row tiles: whole-plane kernels
    int plane_T, plane_rows;    // (run_plane_packed4)
    size_t off_plane, off_plane_bounds, off_plane_stats;
};

bool tiled_plan(const ep_events_soa* ev, const ep_bin_params* p, TiledPlan& pl) {
    if (ev->xy_dtype != EP_U32 || ev->t_dtype != 0 || ev->t != nullptr) return false;      // 4 B packed layout only
    if (p->count_channels != 0 || p->num_bins < 1 || p->num_bins > 64 || p->time_f32) return false;
    if (!ev->offsets_host || ev->batch <= 0) return false;
    const int H = p->height, W = p->width;
    if (W > kTileCells || W > 65535 || H > 65535) return false;
    int rows = kTileCells / W;
    if (rows > H) rows = H;
    int NT = (H + rows - 1) / rows;
    if (NT > kMaxTiles) return false;
    rows = (H + NT - 1) / NT;
    NT = (H + rows - 1) / rows;
    pl.NT = NT; pl.rows = rows;
    const int B = ev->batch;
    int64_t tasks = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t lo = ev->offsets_host[b], hi = ev->offsets_host[b + 1];
        if (hi > lo) tasks += ((hi - 1) >> kChunkShift) - (lo >> kChunkShift) + 1;
    }
    if (tasks > (1ll << 30) || (int64_t)B * NT > (1ll << 30)) return false;
    pl.n_tasks = (int)tasks;
    pl.rec_pos0 = (ev->offsets_host[0] >> kChunkShift) << kChunkShift;
    pl.n_rec = ev->offsets_host[B] - pl.rec_pos0;
    if (pl.n_rec >= (1ll << 32)) return false;                  // ChunkMeta::pos0 is 32 bits
    const size_t nt = (size_t)(tasks ? tasks : 1);
    size_t o = 0;
    pl.off_meta = o; o += align_up(sizeof(SampleMeta) * (size_t)B, 256);
    pl.off_first = o; o += align_up(sizeof(int) * ((size_t)B + 1), 256);
    pl.off_desc = o; o += align_up(sizeof(TaskDesc) * nt, 256);
    pl.off_cmeta = o; o += align_up(sizeof(ChunkMeta) * nt, 256);
    pl.off_coff = o; o += align_up(sizeof(uint16_t) * nt * (size_t)(NT + 2), 256);
    pl.off_crel = o; o += align_up(sizeof(uint32_t) * nt * 32, 256);
    pl.off_counters = o; o += 256;
    pl.off_stats = o; o += align_up(sizeof(double) * kStatSlot * (size_t)B * NT * (size_t)(p->num_bins + 1), 256);
    pl.off_stats2 = o; o += align_up(sizeof(double) * 3 * kStatSlices * (size_t)(p->num_bins + 1), 256);
    pl.off_rec = o; o += align_up(sizeof(uint32_t) * (size_t)(pl.n_rec > 0 ? pl.n_rec : 1), 256);
    // Whole-plane kernels: every event is read 2 x (row tiles) times, so more than one tile only pays where the output
    // outweighs the events (MVSEC-shaped batches: 346 x 260, 9 bins, ~100 k events): 2 tiles always, 3 when cells >= 2 x events.
    // (Measured and dropped: half-size tiles with two 512-thread CTAs per SM for output-heavy batches — 346 x 260 x 9 bins,
    // B = 512: 0.94 ms against 0.67 ms with two whole tiles; 224 x 224 with 512 threads at 64 registers: 0.83 against 0.75 ms.)
    pl.plane_ok = false; pl.plane_T = 1; pl.plane_rows = H;
    if (W <= kPlaneCellsMax && (int64_t)B * p->num_bins * kPlaneMaxTiles < (1ll << 30)) {
        int prow = kPlaneCellsMax / W;
        if (prow > H) prow = H;
        int T = (H + prow - 1) / prow;
        prow = (H + T - 1) / T;
        T = (H + prow - 1) / prow;
        const int64_t cells = (int64_t)B * p->num_bins * H * W, n_ev = ev->offsets_host[B] - ev->offsets_host[0];
        if (T <= 2 || (T == 3 && cells >= 2 * n_ev)) { pl.plane_ok = true; pl.plane_T = T; pl.plane_rows = prow; }
    }
    pl.off_plane = pl.off_plane_bounds = pl.off_plane_stats = 0;
    if (pl.plane_ok) {
        pl.off_plane = o; o += align_up(256 + sizeof(unsigned int) * (size_t)B * kPlaneMaxTiles, 256);      // counters | finished planes per (sample, tile)
        pl.off_plane_bounds = o; o += align_up(sizeof(int64_t) * (size_t)B * (size_t)(p->num_bins + 3), 256);
        pl.off_plane_stats = o; o += align_up(sizeof(double) * kPlaneStatSlot * (size_t)B * pl.plane_T * (size_t)(p->num_bins + 1), 256);
    }
    pl.total = o;
    return true;
}

}  // namespace

size_t tiled_workspace_bytes(const ep_events_soa* ev, const ep_bin_params* p) {
    TiledPlan pl;
    return tiled_plan(ev, p, pl) ? pl.total : 0;
}

// Channel statistics from the sweep's per-(task, channel, warp) partials, in a fixed order (bit-reproducible):
// k_stats_slices: CTA (channel c, slice s) adds the partials of its contiguous share of the tasks -> part2[c][s][0..2];
// k_stats_final: out[c] = (count, sum, sum of squares, max) from the kStatSlices slice sums.
// kWarps = warps of the kernel that wrote the partials; guard (or null): the kernels run only while (*guard != 0) == want —
// the whole-plane path and its stand-by route + sweep each bring their own reduction, one of the two runs.
__global__ void __launch_bounds__(256) k_stats_slices(const double* __restrict__ part, int n_tasks, int n_ch, double* __restrict__ part2,
                                                      int kWarps, const unsigned int* guard, int want) {
    __shared__ double s1[256], s2[256], sm[256];
    if (guard && ((*guard != 0u) != (want != 0))) return;
    const int kStatSlot = kWarps * 3;
    const int c = blockIdx.x, sl = blockIdx.y, tid = threadIdx.x;
    const int per = (n_tasks + kStatSlices - 1) / kStatSlices;
    const int t0 = sl * per, t1 = (t0 + per < n_tasks) ? t0 + per : n_tasks;
    double a1 = 0.0, a2 = 0.0, am = -INFINITY;
    for (int i = t0 * kWarps + tid; i < t1 * kWarps; i += 256) {
        const int task = i / kWarps, w = i - task * kWarps;
        const double* q = part + ((size_t)task * n_ch + c) * kStatSlot + w * 3;
        a1 += q[0]; a2 += q[1]; am = fmax(am, q[2]);
    }
    s1[tid] = a1; s2[tid] = a2; sm[tid] = am;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o]; sm[tid] = fmax(sm[tid], sm[tid + o]); }
        __syncthreads();
    }
    if (tid == 0) {
        double* q = part2 + ((size_t)c * kStatSlices + sl) * 3;
        q[0] = s1[0]; q[1] = s2[0]; q[2] = sm[0];
    }
}

__global__ void __launch_bounds__(32) k_stats_final(const double* __restrict__ part2, double count, double* __restrict__ out,
                                                    const unsigned int* guard, int want) {
    if (guard && ((*guard != 0u) != (want != 0))) return;
    const int c = blockIdx.x, lane = threadIdx.x;
    const double* q = part2 + ((size_t)c * kStatSlices + lane) * 3;
    double a1 = q[0], a2 = q[1], am = q[2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
    }
    if (lane == 0) { out[c * 4 + 0] = count; out[c * 4 + 1] = a1; out[c * 4 + 2] = a2; out[c * 4 + 3] = am; }
}

namespace {
int run_plane_packed4(cudaStream_t st, const ep_events_soa* ev, const ep_bin_params* p, float* out_voxel, float* out_sum,
                      void* ws, size_t ws_bytes, unsigned int* bad, double* out_stats);
}

// standby = device word of the whole-plane path (run_plane_packed4): the kernels below return at once while it is 0, the
// per-sample metadata is already in place and the out-of-grid events are already counted
static int run_tiled_packed4_impl(cudaStream_t st, const ep_events_soa* ev, const ep_bin_params* p, float* out_voxel, float* out_sum,
                                  void* ws, size_t ws_bytes, unsigned int* bad, double* out_stats, const unsigned int* standby) {
    TiledPlan pl;
    if (!tiled_plan(ev, p, pl)) return EP_EUNSUPPORTED;
    if (!ws || ws_bytes < pl.total) return EP_EUNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(ws) & 255u) return EP_EALIGN;
    if (!out_voxel) return EP_EINVAL;
    if (!aligned16(ev->x)) return EP_EALIGN;
    const int B = ev->batch;
    char* base = static_cast<char*>(ws);
    TiledArgs a = {};
    a.run_if = standby;
    a.count_bad = standby ? 0 : 1;
    a.w = static_cast<const uint32_t*>(ev->x);
    a.blk_base = static_cast<const uint32_t*>(ev->p);
    a.offsets = ev->offsets;
    a.n_total = ev->offsets_host[B];
    a.rec_pos0 = pl.rec_pos0;
    a.B = B; a.H = p->height; a.W = p->width; a.num_bins = p->num_bins;
    a.NT = pl.NT; a.rows = pl.rows; a.off_stride = pl.NT + 2;
    a.rep_shift = route_rep_shift(pl.NT);
    a.sx = p->scale_x; a.sy = p->scale_y; a.scaled = (p->scale_x != 1.0 || p->scale_y != 1.0);
    a.n_tasks = pl.n_tasks;
    a.meta = reinterpret_cast<SampleMeta*>(base + pl.off_meta);
    a.first_task = reinterpret_cast<int*>(base + pl.off_first);
    a.desc = reinterpret_cast<TaskDesc*>(base + pl.off_desc);
    a.cmeta = reinterpret_cast<ChunkMeta*>(base + pl.off_cmeta);
    a.coff = reinterpret_cast<uint16_t*>(base + pl.off_coff);
    a.crel = reinterpret_cast<uint32_t*>(base + pl.off_crel);
    a.counters = reinterpret_cast<unsigned int*>(base + pl.off_counters);
    a.rec = reinterpret_cast<uint32_t*>(base + pl.off_rec);
    a.bad_count = bad;
    a.out_voxel = out_voxel;
    a.out_sum = out_sum;
    a.stats_part = out_stats ? reinterpret_cast<double*>(base + pl.off_stats) : nullptr;

    // per-sample first / last stamps and integer-time constants: the same kernel as the global path
    SoaPackedLoader<false> ld{a.w, nullptr, a.blk_base, ev->t_base, ev->t_div};
    BinArgs ba = {};
    ba.offsets = ev->offsets; ba.num_bins = p->num_bins; ba.meta = a.meta;
    profile_begin(st, kProfOther);
    if (!standby) {
        k_sample_meta<SoaPackedLoader<false>><<<(B + 127) / 128, 128, 0, st>>>(ld, ba, B);
        EP_LAUNCH_CHECK();
    }
    k_tiled_setup<<<1, 1024, 0, st>>>(a);
    EP_LAUNCH_CHECK();
    if (a.n_tasks > 0) {
        k_tiled_desc<<<(a.n_tasks + 255) / 256, 256, 0, st>>>(a);
        EP_LAUNCH_CHECK();
    }
    cudaError_t ce = cudaMemsetAsync(a.counters, 0, 256, st);
    if (ce != cudaSuccess) return (int)ce;
    profile_end(st);

    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_route<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRouteSmem);
        cudaFuncSetAttribute(k_route<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRouteSmem);
        cudaFuncSetAttribute(k_sweep<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes(kTileCells));
        cudaFuncSetAttribute(k_sweep<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes(kTileCells));
        cudaFuncSetAttribute(k_sweep<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes(kTileCells));
        cudaFuncSetAttribute(k_sweep<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes(kTileCells));
        attr_done = true;
    }
    // (Measured and dropped: the batch cut into 4-32 sample groups with the routes on the caller's stream and the sweeps on a
    // side stream, one CTA of each kernel per SM, so that route(g + 1) runs under sweep(g): 1.74-2.1 ms against 1.65 ms in
    // plain order — both kernels live on the shared-memory pipe, and each loses its second CTA per SM.)
    a.task_lo = 0; a.task_hi = pl.n_tasks; a.sweep_lo = 0; a.sweep_hi = B * pl.NT; a.counter_idx = 0;
    {
        // planes waiting in L2 for the read-back: (bins - 1) tiles per resident sweep CTA.  EP_DEFERRED_SUM=0/1 overrides.
        const size_t waiting = (size_t)(p->num_bins > 1 ? p->num_bins - 1 : 0) * pl.rows * p->width * 4 * (size_t)(kSweepCtas * kNumSMs);
        static const int forced = [] { const char* e = getenv("EP_DEFERRED_SUM"); return e ? (e[0] != '0' ? 1 : 0) : -1; }();      // read once
        a.deferred_sum = forced >= 0 ? forced : (waiting <= ((size_t)80 << 20));
    }
    if (pl.n_tasks > 0) {
        const int grid = pl.n_tasks < 2 * kNumSMs ? pl.n_tasks : 2 * kNumSMs;
        profile_begin(st, kProfScatter);
        if (a.rep_shift) k_route<false, true><<<grid, kRouteThreads, kRouteSmem, st>>>(a);
        else k_route<false><<<grid, kRouteThreads, kRouteSmem, st>>>(a);
        profile_end(st);
        EP_LAUNCH_CHECK();
    }
    {
        const int64_t n_sweep = (int64_t)B * pl.NT;
        const int grid = n_sweep < kSweepCtas * kNumSMs ? (int)n_sweep : kSweepCtas * kNumSMs;
        const size_t smem = sweep_smem_bytes(pl.rows * p->width);
        const bool vec = (p->width % 4 == 0) && aligned16(out_voxel) && aligned16(out_sum);
        profile_begin(st, kProfFinalize);
        if (a.deferred_sum) {
            if (vec) k_sweep<true, true><<<grid, kSweepThreads, smem, st>>>(a);
            else k_sweep<false, true><<<grid, kSweepThreads, smem, st>>>(a);
        } else {
            if (vec) k_sweep<true, false><<<grid, kSweepThreads, smem, st>>>(a);
            else k_sweep<false, false><<<grid, kSweepThreads, smem, st>>>(a);
        }
        profile_end(st);
        EP_LAUNCH_CHECK();
    }
    if (out_stats) {
        const int n_ch = p->num_bins + 1;
        profile_begin(st, kProfOther);
        double* part2 = reinterpret_cast<double*>(base + pl.off_stats2);
        const int n_out = out_sum ? n_ch : p->num_bins;
        k_stats_slices<<<dim3((unsigned)n_out, kStatSlices), 256, 0, st>>>(a.stats_part, B * pl.NT, n_ch, part2, kSweepThreads / 32, standby, 1);
        EP_LAUNCH_CHECK();
        k_stats_final<<<n_out, 32, 0, st>>>(part2, (double)B * p->height * p->width, out_stats, standby, 1);
        profile_end(st);
        EP_LAUNCH_CHECK();
    }
    return EP_OK;
}

// =====================================================================================================================
// Whole-plane path: grids whose plane fits the shared memory of one SM (224 x 224 of the reference's own pre-training
// order — events_reshape to 224 x 224, then the voxel grid, dataset/pretrain/pr_n_imagenet_dataset.py:85-87 — 240 x 180,
// anything up to kPlaneCells cells).  No route pass and no routed records:
//   * k_plane_bounds: per sample, the array positions s_1 .. s_(bins-1) where the temporal intervals begin (bisection on the
//     stamps, 128 probes per round; s_0 = first event, s_bins = end).  For a time-sorted sample the slice [s_j, s_(j+1)) holds
//     exactly the events of interval j.
//   * k_plane: persistent, one CTA of 1024 threads per SM, a task = (sample, output plane k).  The CTA streams the two slices
//     that feed plane k — interval k-1 (right-node weights r) and interval k (left-node weights 2^24 - r) — straight from the
//     packed words (coordinates through the shared-memory tables of the fused events_reshape, integer time arithmetic as in
//     the sweep), adds them with one returning shared-memory atomic per event into the int32 Q24 plane (wraps go to the
//     spill list as in the sweep), converts the plane to fp32 and writes it once.  Every event is read twice (the second
//     time mostly from L2), every output element written once.  The CTA that finishes the last plane of a sample forms
//     voxel.sum(0) from the planes in L2 (sequential fp32 over bins, the reference's order).
//   * exactness for ANY input: every event a CTA reads is checked against the interval its slice stands for.  One mismatch
//     (unsorted rows) raises a device flag, and the route + sweep kernels, which take any order and stand by behind the
//     same stream (they return at once while the flag is 0), redo the whole batch.  Out-of-grid events are counted by
//     position (every slice is some plane's left slice exactly once), so bad_count is right either way.
// Same integers as the other two paths: bit-identical outputs.
// =====================================================================================================================
namespace {

#ifndef EP_PLANE_THREADS
#define EP_PLANE_THREADS 1024
#endif
constexpr int kPlaneThreads = EP_PLANE_THREADS;
constexpr int kPlaneCells = kPlaneCellsMax;
constexpr uint32_t kLutBad = 0x40000000u;     // row base of a y outside the grid: the cell test fails whatever x is

// events_reshape of one axis without a table: trunc(fl(i * s)) == umulhi(i, mul) for every coordinate i < 2048 of the packed
// layout, proven on the host by comparing all of them (plane_axis_mul).  fl() is the fp64 rounding of the product: for some
// scales a few i whose exact product is an integer land just below it (0.35 * 180 = 62.99999999999999); such an axis keeps
// its shared-memory table.
constexpr int kCoordPlain = 0;       // no scale: y * W + x
constexpr int kCoordLut = 1;         // both axes through the tables
constexpr int kCoordMulY = 2;        // y by multiply-high, x through its table
constexpr int kCoordMulX = 3;        // x by multiply-high, y through its table
constexpr int kCoordMul = 4;         // both by multiply-high

struct PlaneArgs {
    const uint32_t* w;
    const uint32_t* blk_base;
    const int64_t* offsets;
    int64_t n_total;
    int B, H, W, num_bins;
    int T, rows;                // row tiles per plane (1 = the whole plane in one CTA), rows per tile
    double sx, sy;
    int scaled;
    int coord_mode;             // kCoord*
    uint32_t mul_x, mul_y;      // multiply-high constants of the axes that have one
    const SampleMeta* meta;
    int64_t* bounds;            // B x (num_bins + 3): s_0 .. s_bins, the last row's stamp in ticks from the first, dt_lim
    unsigned int* counters;     // [0] task counter, [1] fallback flag (an event outside its slice's interval)
    unsigned int* done;         // B x T: finished planes per (sample, tile)
    unsigned int* bad_count;
    float* out_voxel;
    float* out_sum;
    double* stats_part;         // (B x T) x (num_bins + 1) x warps x 3 partial statistics of the written values, or null
};

// dynamic shared memory of k_plane: lut_x u16[2048] | lut_y u32[2048] | spill int2[kSpillCap] | plane int32[tile cells] | dump int32[32]
constexpr uint32_t kPlaneOffLutX = 0, kPlaneOffLutY = 4096, kPlaneOffSpill = 12288, kPlaneOffPlane = kPlaneOffSpill + kSpillCap * 8;
static_assert(kPlaneOffPlane % 16 == 0, "the plane is read with 16-byte loads");

__host__ __device__ inline size_t plane_smem_bytes(int tile_cells) {
    return kPlaneOffPlane + (size_t)((tile_cells + 3) & ~3) * 4 + 128;
}

// stamp of array position i of a sample that starts at lo, as ticks from the sample's first row
__device__ __forceinline__ int64_t plane_dt(const PlaneArgs& a, int64_t i, int64_t lo, int64_t t0_ticks) {
    const int64_t blk = i >> kTickBlockShift;
    const int64_t base = (blk == (lo >> kTickBlockShift)) ? 0 : (int64_t)__ldg(a.blk_base + blk);
    return base + (int64_t)(__ldg(a.w + i) >> 23) - t0_ticks;
}

__global__ void __launch_bounds__(128) k_plane_bounds(PlaneArgs a) {
    __shared__ int s_min;
    const int nb1 = a.num_bins > 1 ? a.num_bins - 1 : 1;
    const int b = blockIdx.x / nb1, j = blockIdx.x % nb1 + 1;       // boundary j = first position of interval j
    const int tid = threadIdx.x;
    const int64_t lo = a.offsets[b], hi = a.offsets[b + 1];
    int64_t* bd = a.bounds + (size_t)b * (a.num_bins + 3);
    const SampleMeta m = a.meta[b];
    if (j == 1 && tid == 0) {
        bd[0] = lo; bd[a.num_bins] = hi;
        bd[a.num_bins + 1] = hi > lo ? plane_dt(a, hi - 1, lo, m.t0_ticks) : 0;      // last row's stamp: dT in ticks
        // largest dt whose v = (dt * tmul + thalf) >> tshift still fits 32 bits (PlaneTime::dt_lim), -1 = the fast quads do not apply
        uint64_t dt_safe = 0xffffffffull;
        if (m.tmul) { const uint64_t q = ((((1ull << 32) << m.tshift) - m.thalf) - 1ull) / m.tmul; dt_safe = q < dt_safe ? q : dt_safe; }
        bd[a.num_bins + 2] = dt_safe >= 1023u ? (int64_t)(dt_safe - 1023u) : -1;
    }
    if (a.num_bins < 2) return;
    const uint32_t T = (uint32_t)j << kQ;
    // Every round probes 128 equidistant positions of [cl, ch) and keeps the gap in front of the first probe at or past the
    // boundary.  The probe positions depend on (cl, ch) only, and "at or past T2" implies "at or past T1" for T1 < T2, so
    // s_1 <= s_2 <= ... holds for any input: the slices always partition the sample.
    int64_t cl = lo, ch = hi;
    while (ch > cl) {
        const int64_t stride = (ch - cl + 127) / 128;
        if (tid == 0) s_min = 128;
        __syncthreads();
        const int64_t pos = cl + tid * stride;
        bool past = true;                                  // positions from ch on are past the boundary
        if (pos < ch) {
            const int64_t dt = plane_dt(a, pos, lo, m.t0_ticks);
            uint32_t v = 0;
            past = v_general(dt, m, a.num_bins, v) ? (v >= T) : (dt > 0);
        }
        if (past) atomicMin(&s_min, tid);
        __syncthreads();
        const int t = s_min;
        const int64_t pt = cl + t * stride;
        const int64_t nl = t == 0 ? cl : cl + (t - 1) * stride + 1;
        const int64_t nh = pt < ch ? pt : ch;
        __syncthreads();
        cl = nl; ch = nh;
    }
    if (tid == 0) bd[j] = cl;
}

struct PlaneCtx {
    uint32_t base_s;            // shared-memory address of the CTA's dynamic buffer: tables, spill list and plane sit at the fixed
                                // offsets kPlaneOff* behind it, so one register addresses them all (immediate offsets)
    int2* spill;
    int* n_spill;
    unsigned int* bad;
    uint32_t hw;                // cells of the grid
    uint32_t dump;              // index of the word behind the plane: out-of-grid events add there (never read)
    uint32_t W, mul_x, mul_y;
    // grids of 2 or 3 row tiles (MT): the CTA owns cells [tbase, tbase + tcells) of the flat index and drops the rest of the
    // slice's events into 32 dump words (one per lane: no same-address serialisation); tile 0 counts the out-of-grid events
    uint32_t tbase, tcells, dumpl;
    bool count_bad;
};

// Per-task constants of the time arithmetic
struct PlaneTime {
    uint32_t tmul, tshift, dT;  // FAST: v = (dt * tmul + thalf) >> tshift for 0 <= dt <= dT ticks (v(dT) = (bins - 1) << 24)
    uint32_t dt_lim;            // block bases (minus the first row's ticks) up to here keep dt + 511 below 2^32 and v below 2^32:
                                // v is then monotone in dt and "v inside the slice's interval" is the whole test
    uint64_t thalf;
};

// Running maxima of a slice: every event of the slice must lie in the slice's interval (u < 2^24) at a stamp inside the
// sample (dt <= dT); one event outside (unsorted rows) hands the batch to the route + sweep kernels.  ce = largest cell index
// (left slices: events outside the grid are counted in a cold path when it leaves the plane).
struct PlaneMax {
    uint32_t u, dt, ce;
};

// shared-memory accesses at (register + immediate offset): the offset goes into the instruction
template <uint32_t OFF>
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(saddr), "n"(OFF));
    return v;
}
template <uint32_t OFF>
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    uint32_t v;
    asm volatile("{\n\t.reg .u16 h;\n\tld.shared.u16 h, [%1+%2];\n\tcvt.u32.u16 %0, h;\n\t}" : "=r"(v) : "r"(saddr), "n"(OFF));
    return v;
}
template <uint32_t OFF>
__device__ __forceinline__ int atoms_add_at(uint32_t saddr, int w) {
    int old;
    asm volatile("atom.shared.add.s32 %0, [%1+%3], %2;" : "=r"(old) : "r"(saddr), "r"(w), "n"(OFF) : "memory");
    return old;
}

// cell of a packed word through the tables of the fused events_reshape: row base + column; anything >= hw is outside the grid
__device__ __forceinline__ uint32_t plane_cell(const PlaneCtx& c, uint32_t word) {
    const uint32_t ly = lds_u32<kPlaneOffLutY>(c.base_s + ((word >> 9) & 0x1ffcu));
    const uint32_t lx = lds_u16<kPlaneOffLutX>(c.base_s + ((word << 1) & 0xffeu));
    return ly + lx;
}

// cells of the four words of a quad.  A table look-up costs ~3.5 shared-memory wavefronts for random coordinates (bank
// conflicts), as much as the atomic itself: with both axes on tables the look-ups, not the atomics, bound the kernel; an
// axis whose scale has a proven multiply-high form costs one IMAD.HI instead.
template <int CM>
__device__ __forceinline__ void plane_cells(const PlaneCtx& c, const uint32_t (&ws)[4], uint32_t (&ce)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t word = ws[e];
        if (CM == kCoordPlain) {
            ce[e] = ((word >> 11) & 0x7ffu) * c.W + (word & 0x7ffu);
        } else {
            const uint32_t row = (CM == kCoordMulY || CM == kCoordMul) ? __umulhi((word >> 11) & 0x7ffu, c.mul_y) * c.W
                                                                       : lds_u32<kPlaneOffLutY>(c.base_s + ((word >> 9) & 0x1ffcu));
            const uint32_t col = (CM == kCoordMulX || CM == kCoordMul) ? __umulhi(word & 0x7ffu, c.mul_x)
                                                                       : lds_u16<kPlaneOffLutX>(c.base_s + ((word << 1) & 0xffeu));
            ce[e] = row + col;
        }
    }
}

// One quad (four packed words of one 16-byte load) of one slice.  LEFT: the slice is interval kexp = k and feeds the plane
// with the left-node weights 2^24 - r, else it is interval kexp = k - 1 and feeds it with r.  EDGE: the quad lies in a tick
// block that straddles an end of the slice, event e belongs to the slice when (rel0 + e) < len (unsigned).  FAST:
// integer-tick sample whose dT leaves room for the wrap-around test: dtb = block base - first-row ticks (mod 2^32), a stamp
// before the first row wraps to more than dT like one past the last row.  An event outside its interval adds a
// meaningless weight — the batch is redone anyway.
template <bool FAST, int CM, bool LEFT, bool EDGE, bool MT>
__device__ __forceinline__ void plane_quad(const PlaneCtx& c, const SampleMeta& m, const PlaneTime& tm, int num_bins, uint4 wq,
                                           int64_t dtb, uint32_t rel0, uint32_t len, uint32_t kbase, PlaneMax& mx, uint32_t& mismatch) {
    const uint32_t ws[4] = {wq.x, wq.y, wq.z, wq.w};
    int old[4], val[4];
    uint32_t cell[4], cev[4];
    plane_cells<CM>(c, ws, cev);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t word = ws[e];
        const bool in = !EDGE || rel0 + (uint32_t)e < len;
        uint32_t u;
        if (FAST) {
            const uint32_t dt = (uint32_t)dtb + (word >> 23);
            const uint64_t q = (uint64_t)dt * tm.tmul + tm.thalf;
            u = __funnelshift_r((uint32_t)q, (uint32_t)(q >> 32), tm.tshift) - kbase;
            if (!EDGE) mx.u = max(mx.u, u);                  // (dt cannot leave the arithmetic's range: block base <= dt_lim)
            else if (in) { mx.dt = max(mx.dt, dt); mx.u = max(mx.u, u); }
        } else {
            uint32_t v = 0;
            const bool ok = v_general(dtb + (int64_t)(word >> 23), m, num_bins, v);
            u = v - kbase;
            if (in && !(ok && u < (1u << kQ))) { mismatch = 1u; u = 0u; }
        }
        const uint32_t ce = cev[e];
        if (LEFT && in) mx.ce = max(mx.ce, ce);
        const int wgt = LEFT ? (int)((1u << kQ) - u) : (int)u;
        val[e] = ((word >> 22) & 1u) ? wgt : -wgt;
        if (EDGE && !in) val[e] = 0;
        if (MT) { const uint32_t cl = ce - c.tbase; cell[e] = cl < c.tcells ? cl : c.dumpl; }
        else cell[e] = min(ce, c.dump);
        old[e] = atoms_add_at<kPlaneOffPlane>(c.base_s + cell[e] * 4u, val[e]);
    }
    uint32_t near = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) near = max(near, (uint32_t)old[e] + 0x7f000000u);
    if (near >= 0xfe000000u) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (cell[e] < c.tcells && near_wrap(old[e])) note_wrap(old[e], val[e], cell[e], c.spill, c.n_spill, c.bad);
    }
}

// the words of a slice are read by two CTAs (planes k and k + 1) at about the same time: EP_PLANE_LDG = 1 loads them with the
// default cache policy instead of evict-first, so that the second reader finds them in L2
#ifndef EP_PLANE_LDG
#define EP_PLANE_LDG 0
#endif
__device__ __forceinline__ uint4 plane_ld(const uint4* p) { return EP_PLANE_LDG ? __ldg(p) : ld_stream(p); }

#ifndef EP_PLANE_PIPE
#define EP_PLANE_PIPE 2        // 0 = plain loop, 1 = the block after next is requested into L2, 2 = next block's words in registers
#endif

// The slice [s0, s1) of the sample that starts at lo: interval kexp (kbase = kexp << 24).  A warp takes one 256-event tick
// block per step (two quads per lane, one block base for the warp).  Blocks that lie inside the slice and whose stamps cannot
// leave the range of the 32-bit time arithmetic (block base <= dt_lim) take the quads without per-event range tests; the
// first and the last block of a slice (they may straddle its ends), the block where the sample starts (its base is the
// sample's own, the first row's ticks are subtracted mod 2^32) and blocks past dt_lim (unsorted input) take the checked quads.
template <bool FAST, int CM, bool LEFT, bool MT>
__device__ __forceinline__ void plane_slice(const PlaneArgs& a, const PlaneCtx& c, const SampleMeta& m, const PlaneTime& tm, int64_t lo,
                                            int64_t s0, int64_t s1, uint32_t kbase, uint32_t& mismatch, uint32_t& nbad) {
    constexpr int kWarps = kPlaneThreads / 32;
    constexpr int kBlk = 1 << kTickBlockShift;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t blk0 = s0 >> kTickBlockShift;
    const uint32_t nblk = (uint32_t)(((s1 - 1) >> kTickBlockShift) - blk0) + 1u;
    const uint32_t first_rel = (uint32_t)((lo >> kTickBlockShift) - blk0);       // the block where the sample starts has base 0
    const bool head = (s0 & (kBlk - 1)) != 0, tail = (s1 & (kBlk - 1)) != 0;
    const uint32_t* wp = a.w + (blk0 << kTickBlockShift) + lane * 4;
    const uint32_t* bp = a.blk_base + blk0;
    const uint32_t t0 = (uint32_t)m.t0_ticks;
    const uint32_t len = (uint32_t)(s1 - s0);
    PlaneMax mx;
    mx.u = 0u; mx.dt = 0u; mx.ce = 0u;
#if EP_PLANE_PIPE == 2
    uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
    uint32_t nb = 0u;
    auto fetch = [&](uint32_t i) {
        const uint4* p = reinterpret_cast<const uint4*>(wp + ((size_t)i << kTickBlockShift));
        const bool edge = (i == 0u && head) || (i == nblk - 1u && tail);
        if (!edge) { n0 = plane_ld(p); n1 = plane_ld(p + 32); }
        nb = (i == first_rel) ? 0u : __ldg(bp + i);
    };
    if ((uint32_t)wid < nblk) fetch((uint32_t)wid);
#endif
    for (uint32_t i = (uint32_t)wid; i < nblk; i += kWarps) {
        const uint4* p = reinterpret_cast<const uint4*>(wp + ((size_t)i << kTickBlockShift));
        bool edge = (i == 0u && head) || (i == nblk - 1u && tail);
#if EP_PLANE_PIPE == 2
        const uint4 w0 = n0, w1 = n1;
        const uint32_t base = nb;
        if (i + kWarps < nblk) fetch(i + kWarps);
#else
        const uint32_t base = (i == first_rel) ? 0u : __ldg(bp + i);
#if EP_PLANE_PIPE == 1
        if (i + 2 * kWarps < nblk) {
            prefetch_l2(p + 2 * kWarps * (kBlk / 4));
            prefetch_l2(p + 2 * kWarps * (kBlk / 4) + 32);
        }
#endif
#endif
        const int64_t dtb = FAST ? (int64_t)(base - t0) : (int64_t)base - m.t0_ticks;
        if (FAST) edge = edge || i == first_rel || (base - t0) > tm.dt_lim;
        if (!edge) {
#if EP_PLANE_PIPE != 2
            const uint4 w0 = plane_ld(p), w1 = plane_ld(p + 32);
#endif
            plane_quad<FAST, CM, LEFT, false, MT>(c, m, tm, a.num_bins, w0, dtb, 0u, len, kbase, mx, mismatch);
            plane_quad<FAST, CM, LEFT, false, MT>(c, m, tm, a.num_bins, w1, dtb, 0u, len, kbase, mx, mismatch);
        } else {
            // checked block: only the events inside [s0, s1) are loaded and added, every event is tested against the sample's time range
            if (FAST && base > 0xfffffdffu) { mismatch = 1u; continue; }       // (a block 2^32 ticks past the first row: unsorted input)
            const int64_t bs = (blk0 + (int64_t)i) << kTickBlockShift;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int64_t pos = bs + h * 128 + lane * 4;
                uint32_t t4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int e = 0; e < 4; ++e) if (pos + e >= s0 && pos + e < s1) t4[e] = a.w[pos + e];
                plane_quad<FAST, CM, LEFT, true, MT>(c, m, tm, a.num_bins, make_uint4(t4[0], t4[1], t4[2], t4[3]), dtb,
                                                     (uint32_t)(pos - s0), len, kbase, mx, mismatch);
            }
        }
        if (LEFT && mx.ce >= c.hw && c.count_bad) {
            // cold: events outside the grid are counted where their slice is the left one (every position exactly once)
            const int64_t bs = (blk0 + (int64_t)i) << kTickBlockShift;
            for (int h = 0; h < 2; ++h)
                for (int e = 0; e < 4; ++e) {
                    const int64_t pos = bs + h * 128 + lane * 4 + e;
                    if (pos >= s0 && pos < s1 && plane_cell(c, a.w[pos]) >= c.hw) ++nbad;
                }
            mx.ce = 0u;
        }
    }
    if (FAST && (mx.u >= (1u << kQ) || mx.dt > tm.dT)) mismatch = 1u;
}

// plane k of a sample: right-node weights of interval k - 1 = [a0, mid), left-node weights of interval k = [mid, a1)
template <bool FAST, int CM, bool MT>
__device__ __forceinline__ void plane_accumulate(const PlaneArgs& a, const PlaneCtx& c, const SampleMeta& m, const PlaneTime& tm, int64_t lo,
                                                 int64_t a0, int64_t mid, int64_t a1, uint32_t k, uint32_t& mismatch, uint32_t& nbad) {
    // Slice j is read by the CTAs of planes j (left) and j + 1 (right), which are drawn one after the other and run side by
    // side: even planes take their left slice first, odd planes their right one, so both read it at the same time and the
    // second reader finds it in L2.
    if (k & 1u) {
        if (mid > a0) plane_slice<FAST, CM, false, MT>(a, c, m, tm, lo, a0, mid, (k - 1u) << kQ, mismatch, nbad);
        if (a1 > mid) plane_slice<FAST, CM, true, MT>(a, c, m, tm, lo, mid, a1, k << kQ, mismatch, nbad);
    } else {
        if (a1 > mid) plane_slice<FAST, CM, true, MT>(a, c, m, tm, lo, mid, a1, k << kQ, mismatch, nbad);
        if (mid > a0) plane_slice<FAST, CM, false, MT>(a, c, m, tm, lo, a0, mid, (k - 1u) << kQ, mismatch, nbad);
    }
}

template <bool VEC>
__global__ void __launch_bounds__(kPlaneThreads, 1) k_plane(PlaneArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int HW = a.H * a.W, tile_full = a.rows * a.W, HWp = (tile_full + 3) & ~3;      // HWp: words of the plane buffer
    int* plane = reinterpret_cast<int*>(smem_raw + kPlaneOffPlane);
    uint16_t* lut_x = reinterpret_cast<uint16_t*>(smem_raw + kPlaneOffLutX);
    uint32_t* lut_y = reinterpret_cast<uint32_t*>(smem_raw + kPlaneOffLutY);
    int2* s_spill = reinterpret_cast<int2*>(smem_raw + kPlaneOffSpill);
    __shared__ int s_task[2], s_nspill, s_last;

    const int tid = threadIdx.x;
    const int per_sample = a.num_bins * a.T, n_tasks = a.B * per_sample;
    if (tid == 0) { s_task[0] = (int)atomicAdd(&a.counters[0], 1u); s_nspill = 0; }
    for (int i = tid; i < HWp + 32; i += kPlaneThreads) plane[i] = 0;
    if (tid < kSpillCap) s_spill[tid] = make_int2(-1, 0);
    for (int i = tid; i < 2048; i += kPlaneThreads) {
        const long long yy = a.scaled ? __double2ll_rz(__dmul_rn((double)i, a.sy)) : (long long)i;
        const long long xx = a.scaled ? __double2ll_rz(__dmul_rn((double)i, a.sx)) : (long long)i;
        lut_y[i] = (yy < a.H + 65536 && yy * a.W < (long long)HW) ? (uint32_t)(yy * a.W) : kLutBad;
        lut_x[i] = (uint16_t)(xx < 65535 ? xx : 65535);            // x + y * W has no bound on x alone (events_to_voxel_grid.py:46)
    }
    __syncthreads();
    PlaneCtx c;
    // (laundered through an empty asm: otherwise the address and the dump index are recomputed for every quad, 6 of ~92 instructions)
    c.base_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
    asm volatile("" : "+r"(c.base_s));
    c.spill = s_spill; c.n_spill = &s_nspill; c.bad = a.bad_count;
    c.hw = (uint32_t)HW; c.dump = (uint32_t)HWp; c.W = (uint32_t)a.W; c.mul_x = a.mul_x; c.mul_y = a.mul_y;
    c.dumpl = (uint32_t)HWp + (uint32_t)(tid & 31);
    asm volatile("" : "+r"(c.dump));
    uint32_t mismatch = 0, nbad = 0;
    int cur = 0;
    for (;;) {
        const int task = s_task[cur];
        if (task >= n_tasks) break;
        if (tid == 0) s_task[cur ^ 1] = (int)atomicAdd(&a.counters[0], 1u);       // read behind this task's barriers
        // the planes (and tiles) of a sample run side by side (L2 reuse of its events)
        const int b = task / per_sample, k = (task - b * per_sample) / a.T, t = task - b * per_sample - k * a.T;
        const int tbase = t * tile_full, ncell = (HW - tbase < tile_full) ? HW - tbase : tile_full;
        c.tbase = (uint32_t)tbase; c.tcells = (uint32_t)ncell; c.count_bad = (t == 0);
        const int64_t* bd = a.bounds + (size_t)b * (a.num_bins + 3);
        const int64_t lo = bd[0];
        const int64_t a0 = bd[k > 0 ? k - 1 : 0], mid = bd[k], a1 = bd[k + 1];
        const SampleMeta m = a.meta[b];
        if (mid < a0 || a1 < mid || a1 - a0 >= (1ll << 32)) {
            mismatch = 1u;                                         // (never: the boundaries are monotone by construction)
        } else if (a1 > a0) {
            PlaneTime tm;
            tm.tmul = m.tmul; tm.tshift = m.tshift; tm.thalf = m.thalf;
            const int64_t dT = bd[a.num_bins + 1];                                // > 0 and < 2^32 for an integer-time sample
            tm.dT = (uint32_t)dT;
            const int64_t dt_lim = bd[a.num_bins + 2];
            tm.dt_lim = (uint32_t)dt_lim;
            const bool fast = (m.flags & kFlagIntTime) && dT > 0 && dT < (1ll << 32) - 1024 && dt_lim >= 0;
            if (fast && a.T == 1) {
#define EP_PLANE_ACC(CM, MT) plane_accumulate<true, CM, MT>(a, c, m, tm, lo, a0, mid, a1, (uint32_t)k, mismatch, nbad)
                switch (a.coord_mode) {
                    case kCoordPlain: EP_PLANE_ACC(kCoordPlain, false); break;
                    case kCoordMulY: EP_PLANE_ACC(kCoordMulY, false); break;
                    case kCoordMulX: EP_PLANE_ACC(kCoordMulX, false); break;
                    case kCoordMul: EP_PLANE_ACC(kCoordMul, false); break;
                    default: EP_PLANE_ACC(kCoordLut, false); break;
                }
            } else if (fast) {
                switch (a.coord_mode) {
                    case kCoordPlain: EP_PLANE_ACC(kCoordPlain, true); break;
                    case kCoordMul: EP_PLANE_ACC(kCoordMul, true); break;
                    default: EP_PLANE_ACC(kCoordLut, true); break;
                }
#undef EP_PLANE_ACC
            } else {
                plane_accumulate<false, kCoordLut, true>(a, c, m, tm, lo, a0, mid, a1, (uint32_t)k, mismatch, nbad);
            }
        }
        __syncthreads();
        // ---- plane k complete: fp32 out (one rounding), plane re-zeroed ----
        constexpr float kInv = 1.0f / 16777216.0f;
        const int n_spill = s_nspill < kSpillCap ? s_nspill : kSpillCap;
        float* o = a.out_voxel + ((int64_t)b * a.num_bins + k) * HW + tbase;
        const bool keep = a.out_sum != nullptr;                    // the planes are read back for voxel.sum(0): leave them in L2
        const bool stats = a.stats_part != nullptr;
        StatAcc64 sa, ss;
        sa.init(); ss.init();
        if (VEC) {
            for (int i = tid * 4; i < ncell; i += kPlaneThreads * 4) {
                const int4 q = *reinterpret_cast<const int4*>(plane + i);
                *reinterpret_cast<int4*>(plane + i) = make_int4(0, 0, 0, 0);
                float4 f = make_float4(__int2float_rn(q.x) * kInv, __int2float_rn(q.y) * kInv, __int2float_rn(q.z) * kInv,
                                       __int2float_rn(q.w) * kInv);
                if (n_spill) {
                    const int qi[4] = {q.x, q.y, q.z, q.w};
                    float* fp = reinterpret_cast<float*>(&f);
                    for (int j = 0; j < 4; ++j) {
                        long long hi = 0;
                        for (int t = 0; t < n_spill; ++t) if (s_spill[t].x == i + j) hi += s_spill[t].y;
                        if (hi) fp[j] = __ll2float_rn((hi << 32) + (long long)qi[j]) * kInv;
                    }
                }
                if (keep) __stcg(reinterpret_cast<float4*>(o + i), f);
                else st_stream(reinterpret_cast<float4*>(o + i), f);
                if (stats) { sa.add(f.x); sa.add(f.y); sa.add(f.z); sa.add(f.w); }
            }
        } else {
            for (int i = tid; i < ncell; i += kPlaneThreads) {
                const int q = plane[i];
                plane[i] = 0;
                float f = __int2float_rn(q) * kInv;
                if (n_spill) {
                    long long hi = 0;
                    for (int t = 0; t < n_spill; ++t) if (s_spill[t].x == i) hi += s_spill[t].y;
                    if (hi) f = __ll2float_rn((hi << 32) + (long long)q) * kInv;
                }
                if (keep) __stcg(o + i, f);
                else st_stream(o + i, f);
                if (stats) sa.add(f);
            }
        }
        // statistics of the written values as a by-product (fixed element order per thread, fixed shuffle tree per warp, one
        // slot per (sample, tile, channel, warp) => bit-reproducible), reduced by k_stats_slices / k_stats_final
        double* sp = stats ? a.stats_part + ((size_t)(b * a.T + t) * (a.num_bins + 1)) * kPlaneStatSlot : nullptr;
        if (stats) stat_reduce_store(sa, sp + (size_t)k * kPlaneStatSlot);
        if (keep) {
            // ---- the CTA that completes a sample's last plane forms voxel.sum(0) ----
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned int old = atomicAdd(&a.done[b * a.T + t], 1u);
                s_last = (old + 1u == (unsigned int)a.num_bins) ? 1 : 0;
                if (s_last) __threadfence();
            }
            __syncthreads();
            if (s_last) {
                const float* p0 = a.out_voxel + (int64_t)b * a.num_bins * HW + tbase;
                float* so = a.out_sum + (int64_t)b * HW + tbase;
                if (VEC) {
                    for (int i = tid * 4; i < ncell; i += kPlaneThreads * 4) {
                        float4 s = __ldcg(reinterpret_cast<const float4*>(p0 + i));
                        for (int j0 = 1; j0 < a.num_bins; j0 += 4) {
                            float4 v[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                v[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (j0 + jj < a.num_bins) v[jj] = __ldcg(reinterpret_cast<const float4*>(p0 + (int64_t)(j0 + jj) * HW + i));
                            }
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                if (j0 + jj < a.num_bins) { s.x += v[jj].x; s.y += v[jj].y; s.z += v[jj].z; s.w += v[jj].w; }
                        }
                        st_stream(reinterpret_cast<float4*>(so + i), s);
                        if (stats) { ss.add(s.x); ss.add(s.y); ss.add(s.z); ss.add(s.w); }
                    }
                } else {
                    for (int i = tid; i < ncell; i += kPlaneThreads) {
                        float s = __ldcg(p0 + i);
                        for (int j = 1; j < a.num_bins; ++j) s += __ldcg(p0 + (int64_t)j * HW + i);      // sequential fp32 over bins
                        st_stream(so + i, s);
                        if (stats) ss.add(s);
                    }
                }
                if (stats) stat_reduce_store(ss, sp + (size_t)a.num_bins * kPlaneStatSlot);
            }
        }
        __syncthreads();
        if (n_spill) {
            if (tid < kSpillCap) s_spill[tid] = make_int2(-1, 0);
            if (tid == 0) s_nspill = 0;
            __syncthreads();
        }
        cur ^= 1;
    }
    if (mismatch) atomicOr(&a.counters[1], 1u);
    if (nbad && a.bad_count) atomicAdd(a.bad_count, nbad);
}

// Host: multiply-high form of one axis of events_reshape, proven against the reference's expression (fp64 product, truncation)
// on all 2048 coordinates of the packed layout; false = the axis keeps its table.
static bool plane_axis_mul(double s, uint32_t& mul_out) {
    if (!(s > 0.0) || !(s < 1.0)) return false;
    const uint64_t mul = (uint64_t)ceil(ldexp(s, 32));
    if (mul == 0 || mul >= (1ull << 32)) return false;
    for (uint32_t i = 0; i < 2048; ++i) {
        const volatile double prod = (double)i * s;
        if ((int64_t)prod != (int64_t)(((uint64_t)i * mul) >> 32)) return false;
    }
    mul_out = (uint32_t)mul;
    return true;
}

int run_plane_packed4(cudaStream_t st, const ep_events_soa* ev, const ep_bin_params* p, float* out_voxel, float* out_sum,
                      void* ws, size_t ws_bytes, unsigned int* bad, double* out_stats) {
    TiledPlan pl;
    if (!tiled_plan(ev, p, pl) || !pl.plane_ok) return EP_EUNSUPPORTED;
    if (!ws || ws_bytes < pl.total) return EP_EUNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(ws) & 255u) return EP_EALIGN;
    if (!out_voxel) return EP_EINVAL;
    if (!aligned16(ev->x)) return EP_EALIGN;
    const int B = ev->batch;
    char* base = static_cast<char*>(ws);
    PlaneArgs a = {};
    a.w = static_cast<const uint32_t*>(ev->x);
    a.blk_base = static_cast<const uint32_t*>(ev->p);
    a.offsets = ev->offsets;
    a.n_total = ev->offsets_host[B];
    a.B = B; a.H = p->height; a.W = p->width; a.num_bins = p->num_bins;
    a.T = pl.plane_T; a.rows = pl.plane_rows;
    a.sx = p->scale_x; a.sy = p->scale_y; a.scaled = (p->scale_x != 1.0 || p->scale_y != 1.0);
    SampleMeta* meta = reinterpret_cast<SampleMeta*>(base + pl.off_meta);
    a.meta = meta;
    a.counters = reinterpret_cast<unsigned int*>(base + pl.off_plane);
    a.done = a.counters + 64;
    a.bounds = reinterpret_cast<int64_t*>(base + pl.off_plane_bounds);
    a.bad_count = bad;
    a.out_voxel = out_voxel;
    a.out_sum = out_sum;
    a.stats_part = out_stats ? reinterpret_cast<double*>(base + pl.off_plane_stats) : nullptr;

    SoaPackedLoader<false> ld{a.w, nullptr, a.blk_base, ev->t_base, ev->t_div};
    BinArgs ba = {};
    ba.offsets = ev->offsets; ba.num_bins = p->num_bins; ba.meta = meta;
    profile_begin(st, kProfOther);
    k_sample_meta<SoaPackedLoader<false>><<<(B + 127) / 128, 128, 0, st>>>(ld, ba, B);
    EP_LAUNCH_CHECK();
    cudaError_t ce = cudaMemsetAsync(a.counters, 0, 256 + sizeof(unsigned int) * (size_t)B * a.T, st);
    if (ce != cudaSuccess) return (int)ce;
    k_plane_bounds<<<B * (p->num_bins > 1 ? p->num_bins - 1 : 1), 128, 0, st>>>(a);
    EP_LAUNCH_CHECK();
    profile_end(st);

    a.coord_mode = kCoordPlain;
    if (a.scaled) {
        // (per call: 4096 fp64 products on the host, while the two kernels above start; EP_PLANE_LUT=1 keeps both tables)
        static const bool no_mul = [] { const char* e = getenv("EP_PLANE_LUT"); return e && e[0] == '1'; }();
        const bool mx = !no_mul && plane_axis_mul(a.sx, a.mul_x), my = !no_mul && plane_axis_mul(a.sy, a.mul_y);
        a.coord_mode = mx ? (my ? kCoordMul : kCoordMulX) : (my ? kCoordMulY : kCoordLut);
    }
    const int hw = p->height * p->width, tile_cells = a.rows * a.W;
    const size_t smem = plane_smem_bytes(tile_cells);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_plane<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_smem_bytes(kPlaneCells));
        cudaFuncSetAttribute(k_plane<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_smem_bytes(kPlaneCells));
        attr_done = true;
    }
    const int64_t n_tasks = (int64_t)B * p->num_bins * a.T;
    const int grid = n_tasks < kNumSMs ? (int)n_tasks : kNumSMs;
    const bool vec = (hw % 4 == 0) && (a.T == 1 || tile_cells % 4 == 0) && aligned16(out_voxel) && aligned16(out_sum);
    profile_begin(st, kProfFinalize);
    if (vec) k_plane<true><<<grid, kPlaneThreads, smem, st>>>(a);
    else k_plane<false><<<grid, kPlaneThreads, smem, st>>>(a);
    profile_end(st);
    EP_LAUNCH_CHECK();
    if (out_stats) {
        // (runs only if the planes stand: the stand-by path brings its own reduction)
        const int n_ch = p->num_bins + 1, n_out = out_sum ? n_ch : p->num_bins;
        double* part2 = reinterpret_cast<double*>(base + pl.off_stats2);
        profile_begin(st, kProfOther);
        k_stats_slices<<<dim3((unsigned)n_out, kStatSlices), 256, 0, st>>>(a.stats_part, B * a.T, n_ch, part2, kPlaneStatSlot / 3, a.counters + 1, 0);
        EP_LAUNCH_CHECK();
        k_stats_final<<<n_out, 32, 0, st>>>(part2, (double)B * p->height * p->width, out_stats, a.counters + 1, 0);
        profile_end(st);
        EP_LAUNCH_CHECK();
    }
    // the route + sweep kernels stand by: they run only if k_plane met an event outside its slice's interval
    return run_tiled_packed4_impl(st, ev, p, out_voxel, out_sum, ws, ws_bytes, bad, out_stats, a.counters + 1);
}

}  // namespace

// Returns EP_EUNSUPPORTED when the layout / shape / workspace does not qualify (the caller then takes the global path).
// Grids whose plane fits one SM's shared memory take the whole-plane kernels, everything else route + sweep; EP_BIN_FORCE_TILED / EP_BIN_FORCE_PLANE pin one.
int run_tiled_packed4(cudaStream_t st, const ep_events_soa* ev, const ep_bin_params* p, float* out_voxel, float* out_sum,
                      void* ws, size_t ws_bytes, unsigned int* bad, double* out_stats) {
    if (!(p->flags & EP_BIN_FORCE_TILED)) {
        const int rc = run_plane_packed4(st, ev, p, out_voxel, out_sum, ws, ws_bytes, bad, out_stats);
        if (rc != EP_EUNSUPPORTED || (p->flags & EP_BIN_FORCE_PLANE)) return rc;
    }
    return run_tiled_packed4_impl(st, ev, p, out_voxel, out_sum, ws, ws_bytes, bad, out_stats, nullptr);
}

// =====================================================================================================================
// EvRep over the routed records   events_to_EvRep, dataset/dataset_utils/events_to_image.py:77-125
//
// The reference sorts the events with np.lexsort((t, y, x)) and accumulates, per pixel and in that order, the differences
// of consecutive sorted stamps into float32 sums (np.add.at); the first event of a pixel is differenced against the last
// event of the previous non-empty pixel in x-major order.  Here, for a ragged batch in the 4 B packed transport layout:
//   * k_route<true>: the same route as the voxel path with transposed tiles — a tile is a range of columns, its cells run
//     x-major, so the concatenation (sample, tile, cell) IS the lexsort order;
//   * k_evrep_sweep (persistent, one CTA of 1024 threads per SM, task = (sample, tile); the stamps of a whole tile live in
//     shared memory — an earlier version placed them in an L2-resident scratch: the scattered 4-byte stores alone cost 0.30 of
//     its 1.12 ms on the C4 shape):
//       A  count per cell with one shared-memory atomic per record (events | positives << 16);
//       -  block scan of the counts -> segment starts; windows = consecutive cell ranges whose stamps fit the 188 KB
//          shared-memory window (one window for ordinary tiles, several for dense ones);
//       B  per window: every record's stamp (ticks relative to the sample's first row, u32) goes to its pixel's segment of
//          the window (one returning shared-memory atomic for the slot, one shared-memory store): the segments lie back to
//          back in exactly the reference's sorted order;
//       -  the tile publishes the last stamp of its last non-empty pixel (decoupled look-back: the first non-empty pixel of a
//          tile needs it from the nearest non-empty tile before; tasks are drawn in order, so the chain always advances);
//       C  a lane per cell: the short segment is sorted in place, then numpy's accumulation is replayed exactly (fp32
//          accumulators updated as (float)((double)acc + d), fp64 statistics of :117-120) — the stamp a pixel's first event is
//          differenced against is simply the element in front of its segment: bit-exact;
//       D  E_C, E_I, E_T leave as row pieces of the (3, H, W) float64 output (E_T through an L2-resident transpose buffer).
// Limits reported through bad_count: bit 31 = more than 65535 events on one pixel of one sample, bit 30 = a stamp more than
// 2^32 ticks away from (or before) the sample's first row, bit 29 = more than 47000 events on one pixel, or a tile whose
// events need more than 16 windows (752 k events on ~10 columns).
// =====================================================================================================================
namespace {

constexpr int kEvThreads = 1024;              // one CTA per SM: the stamps of a whole tile live in its shared memory
constexpr int kEvWarps = kEvThreads / 32;
constexpr int kEvTileCells = 4400;            // 8 B per cell (count word + cursor)
constexpr int kEvTab = 256;                   // chunks per run-table round
constexpr int kEvWin = 47000;                 // stamps of the shared-memory window (188 KB): a tile of the C4 shape holds ~34 k
constexpr int kEvMaxWin = 16;                 // windows per tile (dense tiles are taken in several cell ranges)

struct EvRepArgs {
    TiledArgs t;                // t.H = image width (major), t.W = image height (minor), t.rows = columns per tile
    const int64_t* t_base;      // B, or null
    double t_div;
    double* out;                // (B, 3, Himg, Wimg)
    double* et_scratch;         // per CTA: E_T of the tile's cells (L2-resident), read back transposed for the output rows
    double t_rcp;               // 1 / t_div, correctly rounded
    long long* lb_val;          // per task: last stamp of the last non-empty pixel
    int* lb_flag;               // per task: 0 = not yet, 1 = empty tile, 2 = value valid
};

__host__ __device__ inline size_t evrep_smem_bytes(int tile_cells) {
    return (size_t)tile_cells * 8 + (size_t)kEvWin * 4 + (size_t)kEvTab * 16 + 64;
}

__device__ __forceinline__ void sort_u32(uint32_t* s, int n) {
    if (n <= 48) {
        for (int i = 1; i < n; ++i) {
            const uint32_t v = s[i];
            int j = i - 1;
            while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; }
            s[j + 1] = v;
        }
        return;
    }
    for (int root0 = n / 2 - 1; root0 >= 0; --root0) {              // heapsort for hot pixels
        int root = root0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && s[child] < s[child + 1]) ++child;
            if (s[root] >= s[child]) break;
            const uint32_t tmp = s[root]; s[root] = s[child]; s[child] = tmp;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; --end) {
        const uint32_t tmp = s[0]; s[0] = s[end]; s[end] = tmp;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && s[child] < s[child + 1]) ++child;
            if (s[root] >= s[child]) break;
            const uint32_t t2 = s[root]; s[root] = s[child]; s[child] = t2;
            root = child;
        }
    }
}

// stamp of a record in ticks relative to the sample's first row
__device__ __forceinline__ bool rec_ticks(uint32_t r, long long cbase, bool narrow, const uint32_t* crel, uint32_t& out) {
    long long dt = cbase;
    if (narrow) dt += (long long)(r >> 16);
    else dt += (long long)crel[(r >> 16) & 31u] + (long long)(r >> 21);
    out = (uint32_t)dt;
    return dt >= 0 && dt < (1ll << 32);
}

__global__ void __launch_bounds__(kEvThreads, 1) k_evrep_sweep(EvRepArgs e) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TiledArgs& a = e.t;
    const int tile_cells = a.rows * a.W;
    long long* t_cb = reinterpret_cast<long long*>(smem_raw);
    uint32_t* s_cp = reinterpret_cast<uint32_t*>(t_cb + kEvTab);                 // events | positives << 16
    uint32_t* s_st = s_cp + tile_cells;                                          // segment start, then cursor / end
    uint32_t* s_win = s_st + tile_cells;                                         // the stamps of the window's cells, segment by segment
    double* s_et = e.et_scratch + (size_t)blockIdx.x * kEvTileCells;             // E_T per cell (global, stays in L2)
    uint32_t* t_pos = s_win + kEvWin;
    uint16_t* t_len = reinterpret_cast<uint16_t*>(t_pos + kEvTab);
    uint16_t* t_nar = t_len + kEvTab;
    __shared__ int s_task, s_warp[kEvWarps], s_last_cell, s_nwin, s_winb[kEvMaxWin + 1], s_wlast_cell;
    __shared__ unsigned int s_wlastmax, s_prevmax, s_bad;
    __shared__ int s_prev_valid;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int Himg = a.W, Wimg = a.H;                       // the route ran on the transposed image
    const int64_t HW = (int64_t)Himg * Wimg;
    const int n_sweep = a.B * a.NT;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_task = (int)atomicAdd(a.counters, 1u); s_bad = 0u; s_last_cell = -1; s_prev_valid = 0; s_prevmax = 0u; }
        for (int i = tid; i < tile_cells; i += kEvThreads) s_cp[i] = 0u;
        __syncthreads();
        const int task = s_task;
        if (task >= n_sweep) break;
        const int b = task / a.NT, tile = task - b * a.NT;
        const int first = a.first_task[b], nch = a.first_task[b + 1] - first;
        const int col0 = tile * a.rows;
        const int ncols = (Wimg - col0 < a.rows) ? Wimg - col0 : a.rows;
        const int ncell = ncols * Himg;
        const int cell0 = tile_cells - ncell;               // a narrower last tile sits at the end (k_route's lut_y)

        // ---- the passes over the tile's runs share this walker: a warp takes one run at a time ----
        auto walk = [&](auto&& body) {
            for (int c_round = 0; c_round < nch; c_round += kEvTab) {
                const int ci = c_round + tid;
                if (tid < kEvTab) {
                    if (ci < nch) {
                        const ChunkMeta cm = a.cmeta[first + ci];
                        const uint16_t* co = a.coff + (size_t)(first + ci) * a.off_stride + tile;
                        const uint32_t o0 = co[0], o1 = co[1];
                        t_pos[tid] = cm.pos0 + o0;
                        t_len[tid] = (uint16_t)(o1 - o0);
                        t_nar[tid] = (uint16_t)((cm.flags & kChunkNarrow) ? 1 : 0);
                        t_cb[tid] = cm.cbase;
                    } else {
                        t_len[tid] = 0;
                    }
                }
                __syncthreads();
                const int n_run = (nch - c_round < kEvTab) ? nch - c_round : kEvTab;
                for (int r = wid; r < n_run; r += kEvWarps) {
                    const int len = t_len[r];
                    const uint32_t* p = a.rec + t_pos[r];
                    const long long cb = t_cb[r];
                    const bool nar = t_nar[r] != 0;
                    const uint32_t* crel = a.crel + (size_t)(first + c_round + r) * 32;
                    // a run holds ~130 records at this tile size: all of its loads are issued before the first use
                    for (int i0 = 0; i0 < len; i0 += 256) {
                        uint32_t rv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int i = i0 + 32 * j + lane;
                            rv[j] = (i < len) ? __ldcg(p + i) : 0xffffffffu;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (i0 + 32 * j + lane < len) body(rv[j], cb, nar, crel);
                    }
                }
                __syncthreads();
            }
        };

        // ---- A: events and positive events per cell ----
        walk([&](uint32_t r, long long, bool, const uint32_t*) {
            const uint32_t cell = (r >> 2) & 0x3fffu;
            const uint32_t old = atomicAdd(&s_cp[cell], 1u + ((r & 2u) ? 0u : 0x10000u));
            if ((old & 0xffffu) == 0xffffu) atomicOr(&s_bad, 0x80000000u);
        });

        // ---- exclusive scan of the counts in cell (= x-major) order; last non-empty cell ----
        {
            const int per = (tile_cells + kEvThreads - 1) / kEvThreads;
            const int c_lo = tid * per, c_hi = (c_lo + per < tile_cells) ? c_lo + per : tile_cells;
            int sum = 0, lastc = -1;
            for (int c = c_lo; c < c_hi; ++c) {
                const int n = (int)(s_cp[c] & 0xffffu);
                if (n) lastc = c;
                sum += n;
            }
            const int incl = warp_incl_scan(sum, lane);
            if (lane == 31) s_warp[wid] = incl;
            if (lastc >= 0) atomicMax(&s_last_cell, lastc);
            __syncthreads();
            int wbase = 0;
            for (int w = 0; w < wid; ++w) wbase += s_warp[w];
            int run = wbase + incl - sum;
            for (int c = c_lo; c < c_hi; ++c) { const int n = (int)(s_cp[c] & 0xffffu); s_st[c] = (uint32_t)run; run += n; }
            __syncthreads();
        }
        const int last_cell = s_last_cell;
        // ---- windows: consecutive cell ranges whose stamps fit the shared-memory window (one for ordinary tiles) ----
        if (tid == 0) {
            int nw = 0, c = 0;
            s_winb[0] = 0;
            const int total = (last_cell >= 0) ? (int)s_st[last_cell] + (int)(s_cp[last_cell] & 0xffffu) : 0;
            while (c < tile_cells && nw < kEvMaxWin) {
                const int base = (int)s_st[c];
                if (total - base <= kEvWin) { c = tile_cells; }
                else {
                    // largest c2 with begin(c2) - base <= kEvWin: cells [c, c2) fit
                    int lo = c, hi = tile_cells;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int)s_st[mid] - base <= kEvWin) lo = mid; else hi = mid; }
                    if (lo == c) { atomicOr(&s_bad, 0x20000000u); c = tile_cells; }        // one pixel alone overflows the window
                    else c = lo;
                }
                s_winb[++nw] = c;
            }
            if (c < tile_cells) { atomicOr(&s_bad, 0x20000000u); s_winb[nw] = tile_cells; }
            s_nwin = nw;
        }
        __syncthreads();
        const int n_win = s_nwin;
        const bool tile_ok = (s_bad & 0x20000000u) == 0u;

        const SampleMeta* mp = a.meta + b;
        const long long abs0 = (e.t_base ? e.t_base[b] : 0) + mp->t0_ticks;
        const double tdiv = e.t_div, trcp = e.t_rcp;
        // stamp value = ticks / t_div, correctly rounded like the reference's fp64 division: with y = RN(1 / t_div), q = RN(x y)
        // lies within one ulp of x / t_div, and RN(q + (x - t_div q) y) is then the correctly rounded quotient (Markstein)
        auto stamp = [&](uint32_t tk) -> double {
            const double v = (double)(abs0 + (long long)tk);
            if (tdiv == 1.0) return v;
            const double q = __dmul_rn(v, trcp);
            const double r = __fma_rn(-q, tdiv, v);
            return __fma_rn(r, trcp, q);
        };

        for (int w = 0; w < n_win; ++w) {
            const int wc_lo = s_winb[w], wc_hi = s_winb[w + 1];
            const int base = (wc_lo < tile_cells) ? (int)s_st[wc_lo] : 0;        // (cursor not advanced yet for this window's cells)
            if (tid == 0) {
                int lc = wc_hi - 1;
                while (lc >= wc_lo && (s_cp[lc] & 0xffffu) == 0u) --lc;
                s_wlast_cell = lc >= wc_lo ? lc : -1;
                s_wlastmax = 0u;
            }
            __syncthreads();
            const int wlast = s_wlast_cell;
            // ---- B: every stamp of the window's cells to its pixel's segment, in shared memory ----
            if (tile_ok) {
                const bool all = n_win == 1;
                walk([&](uint32_t r, long long cb, bool nar, const uint32_t* crel) {
                    const uint32_t cell = (r >> 2) & 0x3fffu;
                    if (!all && ((int)cell < wc_lo || (int)cell >= wc_hi)) return;
                    uint32_t tk;
                    if (!rec_ticks(r, cb, nar, crel, tk)) { atomicOr(&s_bad, 0x40000000u); tk = 0u; }
                    const uint32_t slot = atomicAdd(&s_st[cell], 1u) - (uint32_t)base;
                    s_win[slot] = tk;
                    if ((int)cell == wlast) atomicMax(&s_wlastmax, tk);
                });
            }
            // (walk ends with a barrier: s_st[c] is now the END of cell c's segment for the window's cells)
            // ---- the tile's last stamp for the tiles behind (after the last window); errors ----
            if (tid == 0 && w == n_win - 1) {
                if (last_cell >= 0 && tile_ok) {
                    e.lb_val[task] = (long long)s_wlastmax;
                    __threadfence();
                    *reinterpret_cast<volatile int*>(e.lb_flag + task) = 2;
                } else {
                    __threadfence();
                    *reinterpret_cast<volatile int*>(e.lb_flag + task) = 1;
                }
                if (s_bad && a.bad_count) atomicOr(a.bad_count, s_bad);
            }

            // ---- C1: every pixel's segment sorted in place (a lane per cell, 32 consecutive cells per warp step) ----
            const int n_groups = (wc_hi - wc_lo + 31) / 32;
            for (int g = wid; g < n_groups; g += kEvWarps) {
                const int c = wc_lo + g * 32 + lane;
                if (c < wc_hi && tile_ok) {
                    const int n = (int)(s_cp[c] & 0xffffu);
                    if (n > 1) sort_u32(s_win + ((int)s_st[c] - n - base), n);
                }
            }
            __syncthreads();
            // ---- C2: replay.  The segments lie back to back in the reference's order, so the stamp a pixel's first event is
            //          differenced against is simply the element in front of its segment ----
            for (int g = wid; g < n_groups; g += kEvWarps) {
                const int c = wc_lo + g * 32 + lane;
                if (c >= wc_hi) continue;
                const int n = tile_ok ? (int)(s_cp[c] & 0xffffu) : 0;
                float tsum = 0.f, tsq = 0.f;
                if (n > 0) {
                    const int beg = (int)s_st[c] - n - base;
                    const uint32_t* mine = s_win + beg;
                    double pv;
                    if (beg > 0) pv = stamp(s_win[beg - 1]);                     // last stamp of the previous non-empty pixel
                    else if (s_prev_valid) pv = stamp(s_prevmax);                // ... which lies in the window before
                    else {
                        pv = stamp(mine[0]);                                     // np.diff(prepend=sorted[0]), :110
                        // first non-empty pixel of the tile: the nearest non-empty tile before it (decoupled look-back)
                        for (int tj = task - 1; tj >= b * a.NT; --tj) {
                            int f;
                            while ((f = *reinterpret_cast<volatile int*>(e.lb_flag + tj)) == 0) __nanosleep(64);
                            if (f == 2) {
                                __threadfence();
                                pv = stamp((uint32_t)*reinterpret_cast<volatile long long*>(e.lb_val + tj));
                                break;
                            }
                        }
                    }
                    for (int k = 0; k < n; ++k) {
                        const double t = stamp(mine[k]);
                        const double d = __dsub_rn(t, pv);
                        pv = t;
                        tsum = (float)__dadd_rn((double)tsum, d);                        // np.add.at into float32, :113
                        tsq = (float)__dadd_rn((double)tsq, __dmul_rn(d, d));            // :114
                    }
                }
                const double cnt = (double)(n < 1 ? 1 : n);                              // :117
                const double mean = __ddiv_rn((double)tsum, cnt);                         // :118
                double v = __dsub_rn(__ddiv_rn((double)tsq, cnt), __dmul_rn(mean, mean)); // :119
                if (!(v > 0.0)) v = (v != v) ? v : 0.0;
                double et = sqrt(v);
                if (et > 1000.0) et = 1000.0;                                             // :120
                s_et[c] = et;
            }
            __syncthreads();
            if (tid == 0 && s_wlast_cell >= 0) { s_prevmax = s_wlastmax; s_prev_valid = 1; }
            __syncthreads();
        }
        if (n_win == 0) __syncthreads();

        // ---- D: (3, H, W) float64 rows: pieces of ncols consecutive x per image row ----
        double* o = e.out + (int64_t)b * 3 * HW + col0;
        for (int i = tid; i < ncell; i += kEvThreads) {
            const int y = i / ncols, xl = i - y * ncols;
            const int c = cell0 + xl * Himg + y;
            const uint32_t cp = s_cp[c];
            const int n = (int)(cp & 0xffffu), pos = (int)(cp >> 16);
            double* q = o + (int64_t)y * Wimg + xl;
            q[0] = (double)n;
            q[HW] = (double)(2 * pos - n);
            q[2 * HW] = __ldcg(s_et + c);
        }
    }
}

}  // namespace

// layout of the workspace of the tiled EvRep
struct EvRepPlan {
    TiledPlan t;
    size_t off_lbval, off_lbflag, off_et, total;
};

static bool evrep_plan(const ep_events_soa* ev, int height, int width, EvRepPlan& pl) {
    if (ev->xy_dtype != EP_U32 || ev->t_dtype != 0 || ev->t != nullptr) return false;      // 4 B packed layout only
    if (!ev->offsets_host || ev->batch <= 0) return false;
    if (height > kEvTileCells || height > 2048 || width > 2048) return false;
    int cols = kEvTileCells / height;
    if (cols > width) cols = width;
    int NT = (width + cols - 1) / cols;
    if (NT > kMaxTiles) return false;
    cols = (width + NT - 1) / NT;
    NT = (width + cols - 1) / cols;
    const int B = ev->batch;
    int64_t tasks = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t lo = ev->offsets_host[b], hi = ev->offsets_host[b + 1];
        if (hi > lo) tasks += ((hi - 1) >> kChunkShift) - (lo >> kChunkShift) + 1;
    }
    if (tasks > (1ll << 30) || (int64_t)B * NT > (1ll << 30)) return false;
    TiledPlan& t = pl.t;
    t.NT = NT; t.rows = cols; t.n_tasks = (int)tasks;
    t.rec_pos0 = (ev->offsets_host[0] >> kChunkShift) << kChunkShift;
    t.n_rec = ev->offsets_host[B] - t.rec_pos0;
    if (t.n_rec >= (1ll << 32)) return false;
    const size_t nt = (size_t)(tasks ? tasks : 1);
    const size_t nrec = (size_t)(t.n_rec > 0 ? t.n_rec : 1);
    size_t o = 0;
    t.off_meta = o; o += align_up(sizeof(SampleMeta) * (size_t)B, 256);
    t.off_first = o; o += align_up(sizeof(int) * ((size_t)B + 1), 256);
    t.off_desc = o; o += align_up(sizeof(TaskDesc) * nt, 256);
    t.off_cmeta = o; o += align_up(sizeof(ChunkMeta) * nt, 256);
    t.off_coff = o; o += align_up(sizeof(uint16_t) * nt * (size_t)(NT + 2), 256);
    t.off_crel = o; o += align_up(sizeof(uint32_t) * nt * 32, 256);
    t.off_counters = o; o += 256;
    pl.off_lbflag = o; o += align_up(sizeof(int) * (size_t)B * NT, 256);          // zeroed together with the counters
    pl.off_lbval = o; o += align_up(sizeof(long long) * (size_t)B * NT, 256);
    pl.off_et = o; o += align_up(sizeof(double) * (size_t)kEvTileCells * 2 * 256, 256);      // up to 512 resident CTAs
    t.off_stats = t.off_stats2 = o;
    t.off_rec = o; o += align_up(sizeof(uint32_t) * nrec, 256);
    t.total = pl.total = o;
    return true;
}

size_t evrep_packed4_workspace_bytes(const ep_events_soa* ev, int height, int width) {
    EvRepPlan pl;
    return evrep_plan(ev, height, width, pl) ? pl.total : 0;
}

int run_evrep_packed4(cudaStream_t st, const ep_events_soa* ev, int height, int width, double* out, void* ws, size_t ws_bytes,
                      unsigned int* bad) {
    EvRepPlan pl;
    if (!evrep_plan(ev, height, width, pl)) return EP_EUNSUPPORTED;
    if (!ws || ws_bytes < pl.total) return EP_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(ws) & 255u) return EP_EALIGN;
    if (!out || !aligned16(ev->x)) return EP_EALIGN;
    const int B = ev->batch;
    char* base = static_cast<char*>(ws);
    EvRepArgs e = {};
    TiledArgs& a = e.t;
    a.count_bad = 1; a.run_if = nullptr;
    a.w = static_cast<const uint32_t*>(ev->x);
    a.blk_base = static_cast<const uint32_t*>(ev->p);
    a.offsets = ev->offsets;
    a.n_total = ev->offsets_host[B];
    a.rec_pos0 = pl.t.rec_pos0;
    a.B = B; a.H = width; a.W = height; a.num_bins = 2;              // transposed: tiles are column ranges
    a.NT = pl.t.NT; a.rows = pl.t.rows; a.off_stride = pl.t.NT + 2;
    a.rep_shift = route_rep_shift(pl.t.NT);
    a.sx = a.sy = 1.0; a.scaled = 0;
    a.n_tasks = pl.t.n_tasks;
    a.task_lo = 0; a.task_hi = pl.t.n_tasks; a.sweep_lo = 0; a.sweep_hi = B * pl.t.NT; a.counter_idx = 0; a.deferred_sum = 0;
    a.meta = reinterpret_cast<SampleMeta*>(base + pl.t.off_meta);
    a.first_task = reinterpret_cast<int*>(base + pl.t.off_first);
    a.desc = reinterpret_cast<TaskDesc*>(base + pl.t.off_desc);
    a.cmeta = reinterpret_cast<ChunkMeta*>(base + pl.t.off_cmeta);
    a.coff = reinterpret_cast<uint16_t*>(base + pl.t.off_coff);
    a.crel = reinterpret_cast<uint32_t*>(base + pl.t.off_crel);
    a.counters = reinterpret_cast<unsigned int*>(base + pl.t.off_counters);
    a.rec = reinterpret_cast<uint32_t*>(base + pl.t.off_rec);
    a.bad_count = bad;
    a.out_voxel = nullptr; a.out_sum = nullptr; a.stats_part = nullptr;
    e.t_base = ev->t_base; e.t_div = ev->t_div; e.t_rcp = 1.0 / ev->t_div; e.out = out;
    e.et_scratch = reinterpret_cast<double*>(base + pl.off_et);
    e.lb_val = reinterpret_cast<long long*>(base + pl.off_lbval);
    e.lb_flag = reinterpret_cast<int*>(base + pl.off_lbflag);

    SoaPackedLoader<false> ld{a.w, nullptr, a.blk_base, ev->t_base, ev->t_div};
    BinArgs ba = {};
    ba.offsets = ev->offsets; ba.num_bins = 2; ba.meta = a.meta;
    k_sample_meta<SoaPackedLoader<false>><<<(B + 127) / 128, 128, 0, st>>>(ld, ba, B);
    EP_LAUNCH_CHECK();
    k_tiled_setup<<<1, 1024, 0, st>>>(a);
    EP_LAUNCH_CHECK();
    if (a.n_tasks > 0) {
        k_tiled_desc<<<(a.n_tasks + 255) / 256, 256, 0, st>>>(a);
        EP_LAUNCH_CHECK();
    }
    cudaError_t ce = cudaMemsetAsync(a.counters, 0, pl.off_lbval - pl.t.off_counters, st);      // task counter + look-back flags
    if (ce != cudaSuccess) return (int)ce;
    static bool attr_done = false;
    static int sweep_ctas_per_sm = 0;
    const size_t smem = evrep_smem_bytes(kEvTileCells);
    if (!attr_done) {
        cudaFuncSetAttribute(k_route<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRouteSmem);
        cudaFuncSetAttribute(k_route<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRouteSmem);
        cudaFuncSetAttribute(k_evrep_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sweep_ctas_per_sm, k_evrep_sweep, kEvThreads, smem);
        attr_done = true;
    }
    if (sweep_ctas_per_sm < 1) return EP_EUNSUPPORTED;
    if (pl.t.n_tasks > 0) {
        const int grid = pl.t.n_tasks < 2 * kNumSMs ? pl.t.n_tasks : 2 * kNumSMs;
        if (a.rep_shift) k_route<true, true><<<grid, kRouteThreads, kRouteSmem, st>>>(a);
        else k_route<true><<<grid, kRouteThreads, kRouteSmem, st>>>(a);
        EP_LAUNCH_CHECK();
    }
    {
        // every CTA must be resident: the look-back of a tile waits for tiles drawn earlier
        int dev = 0, sms = kNumSMs;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t n_sweep = (int64_t)B * pl.t.NT;
        int64_t cap = (int64_t)sms * sweep_ctas_per_sm;
        if (cap > 512) cap = 512;                            // E_T scratch slots
        const int grid = (int)(n_sweep < cap ? n_sweep : cap);
        k_evrep_sweep<<<grid, kEvThreads, smem, st>>>(e);
        EP_LAUNCH_CHECK();
    }
    return EP_OK;
}

}  // namespace ep

extern "C" int ep_reshape_axis_multiplier_host(double scale, uint32_t* multiplier) {
    uint32_t m = 0;
    const bool ok = ep::plane_axis_mul(scale, m);
    if (multiplier) *multiplier = ok ? m : 0u;
    return ok ? 1 : 0;
}

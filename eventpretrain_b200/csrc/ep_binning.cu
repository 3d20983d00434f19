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
#include "ep_binning_common.cuh"

namespace ep {

// finalize-free shared-memory path (ep_binning_tiled.cu): 4 B/event packed layout, voxel grid (+ sum plane)
int run_tiled_packed4(cudaStream_t st, const ep_events_soa* ev, const ep_bin_params* p, float* out_voxel, float* out_sum,
                      void* ws, size_t ws_bytes, unsigned int* bad, double* out_stats);
// one-pass channel statistics of a (B,C,H,W) tensor (ep_misc.cu): the paths that do not produce them as a by-product
int plane_statistics(cudaStream_t st, const float* x, int batch, int channels, int64_t hw, double* out, void* ws, size_t ws_bytes);
size_t tiled_workspace_bytes(const ep_events_soa* ev, const ep_bin_params* p);

namespace {

// ---- scatter ----------------------------------------------------------------------------------------
// publish "sample has p == 0 events" (selects the neg class of the count frame, events_to_image.py:13-16);
// the flag is read through L2 so that one RED per sample, not per thread, is the steady state
__device__ __forceinline__ void publish_zero_flag(const BinArgs& a, int b) {
    if (!(__ldcg(&a.meta[b].flags) & kFlagZeroPol)) atomicOr(&a.meta[b].flags, kFlagZeroPol);
}

constexpr int kOffCache = 256;     // offsets of the group's samples cached in shared memory (groups are a few samples)

// per-sample constants of the lean path, cached in shared memory next to the offsets
struct LeanMeta {
    int64_t t0_ticks;
    uint32_t tmul, tshift, thalf, flags;
};

// Lean path of one thread's 4 consecutive events: every event of the warp's 128 belongs to sample b, lies inside the
// group, carries integer ticks (kFlagIntTime) and unscaled coordinates.  32-bit index arithmetic, no per-event
// boundary tests, one slot base pointer per quad.  Same REDs as the general path.
template <class Loader>
__device__ __forceinline__ void scatter_quad_lean(const Loader& ld, const BinArgs& a, const typename Loader::Raw& raw, int64_t i0,
                                                  int b, int64_t s_lo, const LeanMeta& lm, unsigned& nbad, bool& saw_zero) {
    uint32_t xs[4], ys[4], pb[4];
    int64_t ti[4];
    ld.decode_lean(raw, xs, ys, pb, ti);
    int64_t rebase = -lm.t0_ticks;
    if constexpr (Loader::kBlocked) rebase += ld.block_base(i0, s_lo);      // a quad never straddles a block (i0 and the block size are multiples of 4)
    const uint32_t W = (uint32_t)a.W, HW = (uint32_t)(a.H * a.W);
    const uint32_t v_end = (uint32_t)a.num_bins << kQ;
    const uint32_t v_last = a.num_bins >= 2 ? (uint32_t)(a.num_bins - 1) << kQ : 0xffffffffu;
    const int slot = b - a.g0;
    unsigned long long* acc = a.vox_acc + (int64_t)slot * a.num_bins * HW;
    uint32_t* cnt = a.cnt_acc + (int64_t)slot * 3 * HW;
#pragma unroll
    for (int j = 0; j < kEvPerThread; ++j) {
        const uint32_t flat = ys[j] * W + xs[j];
        if (flat >= HW || pb[j] > 1u) { ++nbad; continue; }
        if (a.count_channels) {
            atomicAdd(cnt + (pb[j] ? 0u : HW) + flat, 1u);          // class planes: p == 1, p == 0
            saw_zero |= pb[j] == 0u;
        }
        uint32_t v;
        if (a.num_bins && ticks_to_v(ti[j] + rebase, lm.tmul, lm.tshift, lm.thalf, v_end, v)) {
            const bool on_last = v == v_last;                     // exactly on the last node: file under the interval before
            const uint32_t k = (v >> kQ) - (on_last ? 1u : 0u);
            const uint32_t r = on_last ? (1u << kQ) : (v & ((1u << kQ) - 1u));
            if (k == (uint32_t)(a.num_bins - 1) && !(__ldcg(&a.meta[b].flags) & kFlagLastPlane)) atomicOr(&a.meta[b].flags, kFlagLastPlane);
            const long long w = (1ll << kABits) + (long long)r;
            atomicAdd(acc + (k * HW + flat), (unsigned long long)(pb[j] ? w : -w));
        }
    }
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start
// its CTAs as soon as every CTA of the previous kernel has called pdl_trigger() (or exited); pdl_wait() then blocks until
// that previous grid has completed and its memory operations are visible.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class Loader>
__device__ __forceinline__ void scatter_tiles(const Loader& ld, const BinArgs& a, int64_t first_tile, int64_t tile_stride) {
    typedef typename Loader::time_t_ TT;
    pdl_trigger();
    // the owner lookup below runs once per tile and thread: keep the group's offsets (and lean constants) on chip
    __shared__ int64_t s_off[kOffCache + 1];
    __shared__ LeanMeta s_lm[Loader::kLean ? kOffCache : 1];
    const bool cached = a.offsets != nullptr && a.g1 - a.g0 <= kOffCache;
    if (cached) {
        for (int i = threadIdx.x; i <= a.g1 - a.g0; i += kThreads) s_off[i] = a.offsets[a.g0 + i];
        if constexpr (Loader::kLean)
            for (int i = threadIdx.x; i < a.g1 - a.g0; i += kThreads) {
                const SampleMeta m = a.meta[a.g0 + i];
                LeanMeta lm;
                lm.t0_ticks = m.t0_ticks; lm.tmul = m.tmul; lm.tshift = m.tshift; lm.thalf = m.thalf; lm.flags = m.flags;
                s_lm[i] = lm;
            }
        __syncthreads();
    }
    auto off_of = [&](int b) -> int64_t { return cached ? s_off[b - a.g0] : off_at(a, b); };
    const bool lean_ok = Loader::kLean && cached && !a.scaled && a.W < 65536 && (int64_t)a.H * a.W * (a.num_bins > 0 ? a.num_bins : 1) < (1ll << 32);
    unsigned nbad = 0;
    // persistent stride loop over 1024-event tiles; the next tile's loads are issued before this tile's REDs
    typename Loader::Raw raw_next;
    if (Loader::kPrefetch && first_tile < a.n_tiles) {
        const int64_t i0 = a.start4 + (first_tile * kThreads + threadIdx.x) * kEvPerThread;
        if (i0 < a.end) ld.load_raw(i0, a, raw_next);
    }
    // Programmatic dependent launch: everything above (offset cache, first tile's event loads) ran while the previous
    // group's finalize was still draining; the accumulators may only be touched once that grid has completed.
    pdl_wait();
  for (int64_t tile = first_tile; tile < a.n_tiles; tile += tile_stride) {
    const int64_t i0 = a.start4 + (tile * kThreads + threadIdx.x) * kEvPerThread;
    typename Loader::Raw raw = raw_next;
    if (Loader::kPrefetch && tile + tile_stride < a.n_tiles) {
        const int64_t in = a.start4 + ((tile + tile_stride) * kThreads + threadIdx.x) * kEvPerThread;
        if (in < a.end) ld.load_raw(in, a, raw_next);
    }
    if constexpr (Loader::kLean) {
        // warp-uniform test: all 128 events of the warp inside the group and inside one sample with integer-tick constants
        const int64_t wf = a.start4 + (tile * kThreads + (threadIdx.x & ~31)) * kEvPerThread, wl = wf + 32 * kEvPerThread - 1;
        if (lean_ok && wf >= a.begin && wl < a.end) {
            int lo = a.g0, hi = a.g1;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid - a.g0] <= wf) lo = mid; else hi = mid;
            }
            const LeanMeta lm = s_lm[lo - a.g0];
            if (wl < s_off[lo + 1 - a.g0] && (lm.flags & kFlagIntTime)) {
                bool saw_zero = false;
                scatter_quad_lean<Loader>(ld, a, raw, i0, lo, s_off[lo - a.g0], lm, nbad, saw_zero);
                if (saw_zero) publish_zero_flag(a, lo);
                continue;
            }
        }
    }
    if (i0 >= a.end) continue;

    // sample that owns the first in-range event of this thread: last b with off[b] <= i
    const int64_t ifirst = i0 < a.begin ? a.begin : i0;
    int lo = a.g0, hi = a.g1;   // invariant: off[lo] <= ifirst < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off_of(mid) <= ifirst) lo = mid; else hi = mid;
    }
    int b = lo;
    int64_t b_end = off_of(b + 1);

    Ev<TT> e;
    ld.decode(raw, i0, a.end, a, e);

    const int64_t HW = (int64_t)a.H * a.W;
    SampleMeta m = a.meta[b];
    int zero_b = -1;

#pragma unroll
    for (int j = 0; j < kEvPerThread; ++j) {
        const int64_t i = i0 + j;
        if (i < a.begin || i >= a.end) continue;
        if (i >= b_end) {
            if (zero_b >= 0) { publish_zero_flag(a, zero_b); zero_b = -1; }
            do { ++b; b_end = off_of(b + 1); } while (i >= b_end);
            m = a.meta[b];
        }
        if constexpr (Loader::kBlocked) e.ti[j] += ld.block_base(i, off_of(b));
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
            int k, r;
            bool last_plane;
            if (voxel_weights<Loader, TT>(e, j, m, a.num_bins, k, r, last_plane)) {
                if (last_plane && !(__ldcg(&a.meta[b].flags) & kFlagLastPlane)) atomicOr(&a.meta[b].flags, kFlagLastPlane);
                const long long sgn = (cls == 0) ? 1 : -1;
                const long long delta = sgn * ((1ll << kABits) + (long long)r);
                atomicAdd(a.vox_acc + ((int64_t)slot * a.num_bins + k) * HW + flat, (unsigned long long)delta);
            }
        }
    }
    if (zero_b >= 0) publish_zero_flag(a, zero_b);
  }
    if (nbad && a.bad_count) atomicAdd(a.bad_count, nbad);
}

// ---- finalize: packed accumulators -> fp32 outputs, slots re-zeroed ---------------------------------
// Bins are processed in register-resident chunks of kFinChunk planes so that all accumulator loads of a
// chunk are in flight together (the first version issued one dependent L2 round trip per bin and was
// latency-bound: profiles/r01_binning_v1_ncu.txt).
template <int VEC, int kFinChunk>
__device__ __forceinline__ void finalize_voxel_block(const BinArgs& a, float* __restrict__ out_voxel,
                                                     float* __restrict__ out_sum, int slot, int64_t blk) {
    const int64_t HW = (int64_t)a.H * a.W;
    const int64_t per = HW / VEC;
    const int64_t idx = blk * kThreads + threadIdx.x;
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
    unsigned long long* acc = a.vox_acc + (int64_t)slot * B * HW + pix;      // walks plane by plane
    float* o = out_voxel + (int64_t)b * B * HW + pix;
    uint32_t ovf = 0;
    for (int k0 = 0; k0 < B; k0 += kFinChunk) {
        unsigned long long w[kFinChunk][VEC];
        {
            const unsigned long long* src = acc;
#pragma unroll
            for (int j = 0; j < kFinChunk; ++j, src += HW) {
                if (k0 + j < n_planes) {
                    if (VEC == 2) {
                        const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(src);
                        w[j][0] = q.x; w[j][VEC - 1] = q.y;
                    } else {
                        w[j][0] = *src;
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) w[j][v] = 0ull;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kFinChunk; ++j, acc += HW, o += HW) {
            if (k0 + j >= B) break;
            if ((w[j][0] | w[j][VEC - 1]) != 0ull) {   // sparse grids: most words are still zero (planes past n_planes read as zero)
                if (VEC == 2) *reinterpret_cast<ulonglong2*>(acc) = make_ulonglong2(0ull, 0ull);
                else *acc = 0ull;
            }
            float r[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                // w = C * 2^44 + A with A the signed low 44 bits: the low words of w and A coincide, so C comes from the high words alone
                const uint32_t w_lo = (uint32_t)w[j][v], w_hi = (uint32_t)(w[j][v] >> 32);
                const int32_t a_hi = (int32_t)(w_hi << (64 - kABits)) >> (64 - kABits);
                const int32_t C = (int32_t)(w_hi - (uint32_t)a_hi) >> (kABits - 32);
                ovf |= (uint32_t)(C + (1 << 18)) >> 19;                          // |C| >= 2^18: reported once per thread below
                const long long A = (long long)(((unsigned long long)(uint32_t)a_hi << 32) | w_lo);
                const long long val = ((long long)C << kQ) - A + a_prev[v];
                a_prev[v] = A;
                r[v] = __ll2float_rn(val) * (1.0f / 16777216.0f);
                sum[v] += r[v];    // voxel.sum(dim=0): sequential fp32 over bins
            }
            if (VEC == 2) st_stream(reinterpret_cast<float2*>(o), make_float2(r[0], r[VEC - 1]));
            else st_stream(o, r[0]);
        }
    }
    if (ovf && a.bad_count) atomicOr(a.bad_count, 0x80000000u);
    if (out_sum) {
        float* s = out_sum + (int64_t)b * HW + pix;
        if (VEC == 2) st_stream(reinterpret_cast<float2*>(s), make_float2(sum[0], sum[VEC - 1]));
        else st_stream(s, sum[0]);
    }
}

__device__ __forceinline__ void finalize_count_block(const BinArgs& a, float* __restrict__ out_count, int slot,
                                                     int64_t blk) {
    const int64_t HW = (int64_t)a.H * a.W;
    const int64_t pix = blk * kThreads + threadIdx.x;
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

// ---- kernels ------------------------------------------------------------------------------------------------
// (An experiment that ran the finalize of group g-1 as interleaved CTAs of the scatter launch of group g, with
// ping-pong accumulator slots, did not help: 3.12 ms vs 3.05 ms per step — both halves are limited by the same L2,
// see DESIGN.md §3.  The two stay separate launches.)
#ifndef EP_SCATTER_MINB
#define EP_SCATTER_MINB 4      // 64 registers
#endif
template <class Loader>
__global__ void __launch_bounds__(kThreads, EP_SCATTER_MINB) k_scatter(Loader ld, BinArgs a) {
    scatter_tiles<Loader>(ld, a, blockIdx.x, gridDim.x);
}

// CHUNK = planes in flight per thread: 8 at 4 CTAs/SM, 4 (grids of up to 5 bins) at 5 CTAs/SM.  (Choosing the residency
// per launch for the fullest last wave made no difference: the kernel is bound by L2 / HBM bandwidth, not by its tail.)
template <int VEC, int CHUNK>
__global__ void __launch_bounds__(kThreads, CHUNK == 4 ? 5 : 4) k_finalize_voxel(BinArgs a, float* __restrict__ out_voxel,
                                                                              float* __restrict__ out_sum) {
    pdl_trigger();
    pdl_wait();
    // (A persistent variant — one contiguous run of pixels per CTA, no partial last wave — measured slower: 1.23 vs 0.98 ms.)
    finalize_voxel_block<VEC, CHUNK>(a, out_voxel, out_sum, blockIdx.y, blockIdx.x);
}

__global__ void __launch_bounds__(kThreads) k_finalize_count(BinArgs a, float* __restrict__ out_count) {
    pdl_trigger();
    pdl_wait();
    finalize_count_block(a, out_count, blockIdx.y, blockIdx.x);
}

// ---- host side ----------------------------------------------------------------------------------------
struct SlotLayout {
    size_t meta_bytes, vox_bytes, cnt_bytes, slot_bytes;
    size_t touched_bytes;      // of a slot, in a normal run: the last interval's plane is neither written nor read (time-sorted input)
};

SlotLayout slot_layout(const ep_bin_params* p, int B) {
    SlotLayout s;
    const size_t HW = (size_t)p->height * p->width;
    s.meta_bytes = align_up(sizeof(SampleMeta) * (size_t)(B > 0 ? B : 1), 256);
    s.vox_bytes = align_up(p->num_bins > 0 ? HW * p->num_bins * sizeof(unsigned long long) : 0, 256);
    s.cnt_bytes = align_up(p->count_channels > 0 ? HW * 3 * sizeof(uint32_t) : 0, 256);
    s.slot_bytes = s.vox_bytes + s.cnt_bytes;
    s.touched_bytes = s.slot_bytes - (p->num_bins >= 2 ? HW * sizeof(unsigned long long) : 0);
    return s;
}

size_t l2_group_budget() {
    // accumulator bytes a group touches, kept L2-resident: measured on B200 with the bench workload (5 samples = 49 MB -> 97.2,
    // 6 = 59 MB -> 99.1, 7 = 69 MB -> 93.1, 8 = 79 MB -> 86.5 Gev/s; 64 rather than 60 lets two 15-bin 640x440 samples
    // (63 MB) share a group: C4 0.81-0.83 -> 0.77-0.83 ms, C5 1.33 -> 1.28 ms)
    // (read per call, ~0.1 us: tests/test_gpu_stage1.py::test_group_chain_orderings changes it between calls to force one-sample groups)
    const char* e = getenv("EP_L2_GROUP_MB");
    long mb = e ? atol(e) : 64;
    if (mb < 1) mb = 1;
    return (size_t)mb << 20;
}

// EP_PDL=0 turns programmatic dependent launches off (plain stream order).
bool pdl_enabled() {
    static const int on = [] { const char* e = getenv("EP_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
    return on != 0;
}

template <class... KArgs, class... Args>
cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, cudaStream_t st, bool after_kernel, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (after_kernel && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

bool persist_l2_enabled() {
    static const bool on = [] { const char* e = getenv("EP_PERSIST_L2"); return e ? atoi(e) != 0 : false; }();
    return on;
}

int scatter_ctas_per_sm() {
    static const int v = [] { const char* e = getenv("EP_SCATTER_CTAS_PER_SM"); const int n = e ? atoi(e) : 8; return n < 1 ? 1 : n; }();
    return v;
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
    const int g_l2 = (int)(budget / L.touched_bytes) > 0 ? (int)(budget / L.touched_bytes) : 1;
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
    a.begin = a.end = a.start4 = 0; a.g0 = a.g1 = 0; a.n_tiles = 0;
    a.n_total = off_host ? off_host[B] : single_n;
    a.vox_acc = nullptr; a.cnt_acc = nullptr;

    profile_begin(st, kProfOther);
    k_sample_meta<Loader><<<(B + 127) / 128, 128, 0, st>>>(ld, a, B);
    profile_end(st);
    EP_LAUNCH_CHECK();
    cudaError_t ce = cudaMemsetAsync(slots, 0, (size_t)G * L.slot_bytes, st);
    if (ce != cudaSuccess) return (int)ce;

    // Pin the accumulator slots in the persisting part of L2 for the duration of the call (B200: 126 MB L2, up to
    // 79 MB persisting): the event stream (13 B/event, read once) and the fp32 outputs (written once) otherwise
    // compete with them for residency, and a RED that misses L2 costs a DRAM round trip.
    bool window_set = false;
    cudaStreamAttrValue old_attr;
    if (persist_l2_enabled()) {
        int dev = 0, max_persist = 0, max_window = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        size_t want = (size_t)G * L.slot_bytes;
        if (max_persist > 0 && max_window > 0) {
            if (want > (size_t)max_persist) want = (size_t)max_persist;
            if (want > (size_t)max_window) want = (size_t)max_window;
            size_t cur_limit = 0;
            cudaDeviceGetLimit(&cur_limit, cudaLimitPersistingL2CacheSize);
            if (cur_limit < want) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamGetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &old_attr);
            cudaStreamAttrValue attr;
            attr.accessPolicyWindow.base_ptr = slots;
            attr.accessPolicyWindow.num_bytes = want;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            window_set = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
            cudaGetLastError();
        }
    }

    const int64_t HW = (int64_t)p->height * p->width;
    bool chained = false;      // the previous operation on the stream is one of this call's kernels (not the memset, not a timing event)
    for (int g0 = 0; g0 < B; g0 += G) {
        const int g1 = (g0 + G < B) ? g0 + G : B;
        a.g0 = g0; a.g1 = g1;
        a.begin = off_host ? off_host[g0] : 0;
        a.end = off_host ? off_host[g1] : single_n;
        a.start4 = a.begin / kEvPerThread * kEvPerThread;
        a.vox_acc = reinterpret_cast<unsigned long long*>(slots);
        a.cnt_acc = reinterpret_cast<uint32_t*>(slots + (size_t)G * L.vox_bytes);
        if (a.end > a.begin) {
            a.n_tiles = ceil_div64(ceil_div64(a.end - a.start4, kEvPerThread), kThreads);
            const int64_t max_grid = (int64_t)kNumSMs * scatter_ctas_per_sm();
            const unsigned grid = (unsigned)(a.n_tiles < max_grid ? a.n_tiles : max_grid);
            profile_begin(st, kProfScatter);
            launch_chain(k_scatter<Loader>, dim3(grid), st, chained, (size_t)0, ld, a);
            profile_end(st);
            EP_LAUNCH_CHECK();
            chained = !profile_enabled();
        }
        if (p->num_bins > 0) {
            profile_begin(st, kProfFinalize);
            const int vec = (HW % 2 == 0) ? 2 : 1;
            dim3 grid((unsigned)ceil_div64(HW / vec, kThreads), (unsigned)(g1 - g0));
            if (p->num_bins <= 5) {
                if (vec == 2) launch_chain(k_finalize_voxel<2, 4>, grid, st, chained, (size_t)0, a, out_voxel, out_sum);
                else launch_chain(k_finalize_voxel<1, 4>, grid, st, chained, (size_t)0, a, out_voxel, out_sum);
            } else {
                if (vec == 2) launch_chain(k_finalize_voxel<2, 8>, grid, st, chained, (size_t)0, a, out_voxel, out_sum);
                else launch_chain(k_finalize_voxel<1, 8>, grid, st, chained, (size_t)0, a, out_voxel, out_sum);
            }
            profile_end(st);
            EP_LAUNCH_CHECK();
            chained = !profile_enabled();
        }
        if (p->count_channels > 0) {
            dim3 grid((unsigned)ceil_div64(HW, kThreads), (unsigned)(g1 - g0));
            profile_begin(st, kProfFinalize);
            launch_chain(k_finalize_count, grid, st, chained, (size_t)0, a, out_count);
            profile_end(st);
            EP_LAUNCH_CHECK();
            chained = !profile_enabled();
        }
    }
    if (window_set) cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &old_attr);
    return EP_OK;
}

}  // namespace
}  // namespace ep

extern "C" {

size_t ep_bin_events_workspace_bytes(const ep_bin_params* prm, int batch, size_t* min_bytes) {
    if (ep::check_params(prm) != EP_OK || batch <= 0) { if (min_bytes) *min_bytes = 0; return 0; }
    const ep::SlotLayout L = ep::slot_layout(prm, batch);
    if (min_bytes) *min_bytes = L.meta_bytes + L.slot_bytes;
    size_t g = ep::l2_group_budget() / L.touched_bytes;
    if (g < 1) g = 1;
    if (g > (size_t)batch) g = (size_t)batch;
    return L.meta_bytes + g * L.slot_bytes;
}

size_t ep_bin_events_workspace_bytes_for(const ep_events_soa* ev, const ep_bin_params* prm) {
    if (!ev || ep::check_params(prm) != EP_OK || ev->batch <= 0) return 0;
    size_t need = ep_bin_events_workspace_bytes(prm, ev->batch, nullptr);
    if (!(prm->flags & EP_BIN_FORCE_GLOBAL) && ev->offsets_host) {
        const size_t t = ep::tiled_workspace_bytes(ev, prm);
        if (t > need) need = t;
    }
    return need;
}

int ep_bin_events(void* stream, const ep_events_soa* ev, const ep_bin_params* prm, float* out_voxel,
                  float* out_voxel_sum, float* out_count, void* workspace, size_t workspace_bytes,
                  unsigned int* bad_count) {
    return ep_bin_events_stats(stream, ev, prm, out_voxel, out_voxel_sum, out_count, workspace, workspace_bytes, bad_count, nullptr);
}

static int bin_events_any(void* stream, const ep_events_soa* ev, const ep_bin_params* prm, float* out_voxel,
                          float* out_voxel_sum, float* out_count, void* workspace, size_t workspace_bytes,
                          unsigned int* bad_count, double* out_stats, bool* stats_done);

int ep_bin_events_stats(void* stream, const ep_events_soa* ev, const ep_bin_params* prm, float* out_voxel,
                        float* out_voxel_sum, float* out_count, void* workspace, size_t workspace_bytes,
                        unsigned int* bad_count, double* out_stats) {
    bool done = false;
    int rc = bin_events_any(stream, ev, prm, out_voxel, out_voxel_sum, out_count, workspace, workspace_bytes, bad_count, out_stats, &done);
    if (rc != EP_OK || !out_stats || done) return rc;
    if (!prm->num_bins || !out_voxel) return EP_EINVAL;
    // kernels without the fused side output: one extra native pass over the finished planes
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t hw = (int64_t)prm->height * prm->width;
    rc = ep::plane_statistics(st, out_voxel, ev->batch, prm->num_bins, hw, out_stats, workspace, workspace_bytes);
    if (rc == EP_OK && out_voxel_sum)
        rc = ep::plane_statistics(st, out_voxel_sum, ev->batch, 1, hw, out_stats + 4 * prm->num_bins, workspace, workspace_bytes);
    return rc;
}

static int bin_events_any(void* stream, const ep_events_soa* ev, const ep_bin_params* prm, float* out_voxel,
                          float* out_voxel_sum, float* out_count, void* workspace, size_t workspace_bytes,
                          unsigned int* bad_count, double* out_stats, bool* stats_done) {
    using namespace ep;
    int rc = check_params(prm);
    if (rc != EP_OK) return rc;
    if (!ev || ev->batch <= 0 || !ev->offsets || !ev->offsets_host) return EP_EINVAL;
    const bool packed = ev->xy_dtype == EP_U32;       // packed transport layouts: their own pointer rules, checked below
    if (!valid_dtype(ev->xy_dtype)) return EP_EINVAL;
    if (!packed && (!valid_dtype(ev->t_dtype) || (!valid_dtype(ev->p_dtype) && ev->t_dtype != EP_U32))) return EP_EINVAL;
    if (!(ev->t_div != 0.0)) return EP_EINVAL;
    const int B = ev->batch;
    for (int b = 0; b < B; ++b) if (ev->offsets_host[b + 1] < ev->offsets_host[b]) return EP_EINVAL;
    const int64_t n_events = ev->offsets_host[B] - ev->offsets_host[0];
    if (n_events > 0 && !ev->x) return EP_EINVAL;
    if (n_events > 0 && !packed && (!ev->y || !ev->t || (!ev->p && ev->t_dtype != EP_U32))) return EP_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ev->xy_dtype == EP_U32) {
        // packed transport layouts: x -> uint32 words, p -> per-block tick offsets, t -> low tick bytes (5 B/event) or NULL (4 B/event)
        const bool five = ev->t_dtype == EP_U8;
        if ((!five && (ev->t_dtype != 0 || ev->t)) || (five && !ev->t && n_events > 0)) return EP_EINVAL;
        if (ev->y != nullptr || (!ev->p && n_events > 0) || ev->p_dtype != EP_U32 || !ev->t_base || prm->time_f32) return EP_EINVAL;
        if (!aligned16(ev->x) || !aligned16(ev->t)) return EP_EALIGN;
        if (!five && !(prm->flags & EP_BIN_FORCE_GLOBAL)) {
            // default: route + two-plane shared-memory sweep, no global accumulators (ep_binning_tiled.cu); shapes, workspaces or
            // outputs it does not take (count frames, > 64 row tiles) fall through to the global-RED kernels
            rc = run_tiled_packed4(st, ev, prm, out_voxel, out_voxel_sum, workspace, workspace_bytes, bad_count, out_stats);
            if (rc == EP_OK) *stats_done = true;
            if (rc != EP_EUNSUPPORTED || (prm->flags & (EP_BIN_FORCE_TILED | EP_BIN_FORCE_PLANE))) return rc;
        } else if (prm->flags & (EP_BIN_FORCE_TILED | EP_BIN_FORCE_PLANE)) {
            return EP_EUNSUPPORTED;
        }
        if (five) {
            SoaPackedLoader<true> ld{static_cast<const uint32_t*>(ev->x), static_cast<const uint8_t*>(ev->t),
                                     static_cast<const uint32_t*>(ev->p), ev->t_base, ev->t_div};
            return run_binning(st, ld, ev->offsets, ev->offsets_host, 0, B, prm, out_voxel, out_voxel_sum, out_count, workspace,
                               workspace_bytes, bad_count);
        }
        SoaPackedLoader<false> ld{static_cast<const uint32_t*>(ev->x), nullptr, static_cast<const uint32_t*>(ev->p), ev->t_base, ev->t_div};
        return run_binning(st, ld, ev->offsets, ev->offsets_host, 0, B, prm, out_voxel, out_voxel_sum, out_count, workspace,
                           workspace_bytes, bad_count);
    }
    if (prm->flags & (EP_BIN_FORCE_TILED | EP_BIN_FORCE_PLANE)) return EP_EUNSUPPORTED;
    const bool compact = ev->t_dtype == EP_U32;
    if (compact) {
        // compact transport layout: u16 x,y + u32 (relative ticks | polarity << 31) + per-sample int64 base
        if (ev->xy_dtype != EP_U16 || ev->p != nullptr || !ev->t_base || prm->time_f32) return EP_EINVAL;
        if (!aligned16(ev->x) || !aligned16(ev->y) || !aligned16(ev->t)) return EP_EALIGN;
    }
    const bool canon = !compact && ev->xy_dtype == EP_U16 && ev->p_dtype == EP_U8 && !prm->time_f32 &&
                       (ev->t_dtype == EP_I64 || ev->t_dtype == EP_F64) && aligned16(ev->x) && aligned16(ev->y) &&
                       aligned16(ev->t) && aligned16(ev->p);
    if (compact) {
        SoaCompactLoader ld{static_cast<const uint16_t*>(ev->x), static_cast<const uint16_t*>(ev->y),
                            static_cast<const uint32_t*>(ev->t), ev->t_base, ev->t_div};
        return run_binning(st, ld, ev->offsets, ev->offsets_host, 0, B, prm, out_voxel, out_voxel_sum, out_count, workspace,
                           workspace_bytes, bad_count);
    }
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

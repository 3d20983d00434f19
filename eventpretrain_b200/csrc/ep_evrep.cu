// EvRep (events_to_EvRep, dataset/dataset_utils/events_to_image.py:77-125) on sm_100a — the
// reference's pinned stand-in for a "time surface" channel.
//
// The reference sorts all events with np.lexsort((t, y, x)) and accumulates, per pixel and in that
// order, the differences between consecutive sorted timestamps (the first event of a pixel is
// differenced against the last event of the previous non-empty pixel in x-major order).  Here:
//   1. histogram per pixel in x-major order q = x*H + y        (E_C and E_I share one 64-bit word: one RED per event)
//   2. per-sample exclusive scan over q                        (chunk sums -> scan of the sums -> chunk scans)
//   3. counting-sort scatter of the timestamps into per-pixel segments (cursors start at the segment starts)
//   4. each pixel's (short) segment is sorted by t in place    (one thread per pixel)
//   5. one thread per pixel replays numpy's accumulation exactly: fp32 accumulators updated as
//      (float)((double)acc + d) and the fp64 statistics of :117-120 — bit-exact with the reference.
#include <math.h>

#include "ep_common.cuh"

namespace ep {
namespace {

struct EvDesc {
    const void* x; const void* y; const void* t; const void* p;
    int xy_dtype, t_dtype, p_dtype;
    double t_div;
    const int64_t* offsets;
    int B;
};

struct RepArgs {
    EvDesc ev;
    int H, W;
    int64_t begin, end;
    unsigned long long* cp;   // [B][HW]  indexed by q = x*H + y: count * 2^32 + net polarity (signed sum, two's complement)
    int32_t* start;    // [B][HW]  exclusive scan of the counts (relative to offsets[b])
    int32_t* cursor;   // [B][HW]  next free slot of the pixel's segment (initialised to start by the scan)
    int32_t* sums;     // [B][n_chunks]  scan scratch
    int n_chunks;
    double* sorted_t;  // [n_total] indexed by absolute event slot - offsets[0]
    unsigned int* bad;
};

__device__ __forceinline__ int find_sample(const int64_t* off, int B, int64_t i) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int pol_of(unsigned long long w) { return (int)(uint32_t)w; }
__device__ __forceinline__ int cnt_of(unsigned long long w) { return (int)((long long)(w - (unsigned long long)(long long)pol_of(w)) >> 32); }

// returns q (x-major pixel index) or -1 when the reference would raise IndexError
__device__ __forceinline__ int64_t pixel_q(const RepArgs& a, int64_t i) {
    int64_t x = __double2ll_rz(load_as_double(a.ev.x, a.ev.xy_dtype, i));
    int64_t y = __double2ll_rz(load_as_double(a.ev.y, a.ev.xy_dtype, i));
    if (x < 0) x += a.W;     // numpy fancy indexing wraps negatives
    if (y < 0) y += a.H;
    if (x < 0 || x >= a.W || y < 0 || y >= a.H) return -1;
    return x * a.H + y;
}

__global__ void __launch_bounds__(256) k_evrep_hist(RepArgs a) {
    const int64_t i = a.begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.end) return;
    const int b = find_sample(a.ev.offsets, a.ev.B, i);
    const int64_t q = pixel_q(a, i);
    const double p = load_as_double(a.ev.p, a.ev.p_dtype, i);
    if (q < 0 || !(p == 1.0 || p == 0.0 || p == -1.0)) { if (a.bad) atomicAdd(a.bad, 1u); return; }
    const int64_t HW = (int64_t)a.H * a.W;
    const long long delta = (1ll << 32) + (p == 1.0 ? 1 : -1);      // E_C += 1, E_I += +-1 (:97,100-101) in one RED
    atomicAdd(a.cp + b * HW + q, (unsigned long long)delta);
}

// ---- exclusive scan of the per-pixel counts, per sample, over chunks of 4096 pixels (1024 threads x 4 consecutive) ----
constexpr int kScanThreads = 1024, kScanPer = 4, kScanChunk = kScanThreads * kScanPer;

// block-wide exclusive scan of one value per thread; returns the thread's prefix, total in `total`
__device__ __forceinline__ int block_excl_scan(int v, int& total) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
    const int r = s_warp[warp] + incl - v;
    total = s_total;
    __syncthreads();            // s_warp / s_total may be reused by the caller's next round
    return r;
}

__global__ void __launch_bounds__(kScanThreads) k_evrep_chunk_sums(RepArgs a) {
    const int b = blockIdx.y, c = blockIdx.x;
    const int64_t HW = (int64_t)a.H * a.W;
    const unsigned long long* w = a.cp + b * HW;
    const int64_t i0 = (int64_t)c * kScanChunk + threadIdx.x * kScanPer;
    int v = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j)
        if (i0 + j < HW) v += cnt_of(w[i0 + j]);
    int total;
    block_excl_scan(v, total);
    if (threadIdx.x == 0) a.sums[b * a.n_chunks + c] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_evrep_scan_sums(RepArgs a) {
    __shared__ int s_base;
    int32_t* s = a.sums + blockIdx.x * a.n_chunks;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int st = 0; st < a.n_chunks; st += kScanThreads) {
        const int i = st + threadIdx.x;
        const int v = i < a.n_chunks ? s[i] : 0;
        int total;
        const int pre = block_excl_scan(v, total);
        if (i < a.n_chunks) s[i] = s_base + pre;
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) k_evrep_chunk_scan(RepArgs a) {
    const int b = blockIdx.y, c = blockIdx.x;
    const int64_t HW = (int64_t)a.H * a.W;
    const unsigned long long* w = a.cp + b * HW;
    const int64_t i0 = (int64_t)c * kScanChunk + threadIdx.x * kScanPer;
    int n[kScanPer], v = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        n[j] = i0 + j < HW ? cnt_of(w[i0 + j]) : 0;
        v += n[j];
    }
    int total;
    int run = a.sums[b * a.n_chunks + c] + block_excl_scan(v, total);
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        if (i0 + j < HW) {
            a.start[b * HW + i0 + j] = run;
            a.cursor[b * HW + i0 + j] = run;
        }
        run += n[j];
    }
}

__global__ void __launch_bounds__(256) k_evrep_scatter(RepArgs a) {
    const int64_t i = a.begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.end) return;
    const int b = find_sample(a.ev.offsets, a.ev.B, i);
    const int64_t q = pixel_q(a, i);
    const double p = load_as_double(a.ev.p, a.ev.p_dtype, i);
    if (q < 0 || !(p == 1.0 || p == 0.0 || p == -1.0)) return;
    const int64_t HW = (int64_t)a.H * a.W;
    double t = load_as_double(a.ev.t, a.ev.t_dtype, i);
    if (a.ev.t_div != 1.0) t = t / a.ev.t_div;
    const int slot = atomicAdd(a.cursor + b * HW + q, 1);
    a.sorted_t[a.ev.offsets[b] - a.begin + slot] = t;
}

__device__ void sort_segment(double* s, int n) {
    if (n <= 32) {
        for (int i = 1; i < n; ++i) {
            const double v = s[i];
            int j = i - 1;
            while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; }
            s[j + 1] = v;
        }
        return;
    }
    // heapsort for hot pixels
    for (int root0 = n / 2 - 1; root0 >= 0; --root0) {
        int root = root0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && s[child] < s[child + 1]) ++child;
            if (s[root] >= s[child]) break;
            const double tmp = s[root]; s[root] = s[child]; s[child] = tmp;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; --end) {
        const double tmp = s[0]; s[0] = s[end]; s[end] = tmp;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && s[child] < s[child + 1]) ++child;
            if (s[root] >= s[child]) break;
            const double t2 = s[root]; s[root] = s[child]; s[child] = t2;
            root = child;
        }
    }
}

__global__ void __launch_bounds__(256) k_evrep_sort(RepArgs a) {
    const int64_t HW = (int64_t)a.H * a.W;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // pixel of sample blockIdx.y, x-major
    if (q >= HW) return;
    const int b = blockIdx.y;
    const int64_t idx = b * HW + q;
    const int n = cnt_of(a.cp[idx]);
    if (n > 1) sort_segment(a.sorted_t + (a.ev.offsets[b] - a.begin) + a.start[idx], n);
}

__global__ void __launch_bounds__(256) k_evrep_finish(RepArgs a, double* __restrict__ out) {
    const int64_t HW = (int64_t)a.H * a.W;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // y*W + x of sample blockIdx.y (output order)
    if (pix >= HW) return;
    const int b = blockIdx.y;
    const int y = (int)((uint32_t)pix / (uint32_t)a.W), x = (int)((uint32_t)pix - (uint32_t)y * (uint32_t)a.W);
    const int64_t qi = b * HW + (int64_t)x * a.H + y;
    const unsigned long long cw = a.cp[qi];
    const int n = cnt_of(cw);
    float tsum = 0.f, tsq = 0.f;
    if (n > 0) {
        const int64_t lo = (a.ev.offsets[b] - a.begin) + a.start[qi];
        const int64_t sample_lo = a.ev.offsets[b] - a.begin;
        double prev = lo > sample_lo ? a.sorted_t[lo - 1] : a.sorted_t[lo];   // np.diff(prepend=sorted[0]), :110
        for (int k = 0; k < n; ++k) {
            const double t = a.sorted_t[lo + k];
            const double d = __dsub_rn(t, prev);
            prev = t;
            tsum = (float)__dadd_rn((double)tsum, d);                        // np.add.at into float32, :113
            tsq = (float)__dadd_rn((double)tsq, __dmul_rn(d, d));            // :114
        }
    }
    const double c = (double)(n < 1 ? 1 : n);                                // :117
    const double mean = __ddiv_rn((double)tsum, c);                           // :118
    double v = __dsub_rn(__ddiv_rn((double)tsq, c), __dmul_rn(mean, mean));   // :119
    if (!(v > 0.0)) v = (v != v) ? v : 0.0;
    double et = sqrt(v);
    if (et > 1000.0) et = 1000.0;                                             // :120
    double* o = out + (int64_t)b * 3 * HW + pix;
    o[0] = (double)n;
    o[HW] = (double)pol_of(cw);
    o[2 * HW] = et;
}

struct RepLayout { size_t cp, start, cursor, sums, sorted, total; int n_chunks; };

RepLayout rep_layout(int B, int H, int W, int64_t n_total) {
    RepLayout L;
    const size_t plane = align_up(sizeof(int32_t) * (size_t)B * H * W, 256);
    L.n_chunks = (int)ceil_div64((int64_t)H * W, kScanChunk);
    L.cp = 0; L.start = 2 * plane; L.cursor = 3 * plane; L.sums = 4 * plane;
    L.sorted = L.sums + align_up(sizeof(int32_t) * (size_t)B * L.n_chunks, 256);
    L.total = L.sorted + align_up(sizeof(double) * (size_t)(n_total > 0 ? n_total : 1), 256);
    return L;
}

}  // namespace
}  // namespace ep

extern "C" {

size_t ep_evrep_workspace_bytes(int batch, int height, int width, int64_t n_total) {
    if (batch <= 0 || height <= 0 || width <= 0 || n_total < 0) return 0;
    return ep::rep_layout(batch, height, width, n_total).total;
}

size_t ep_evrep_workspace_bytes_for(const ep_events_soa* ev, int height, int width) {
    if (!ev || !ev->offsets_host || ev->batch <= 0 || height <= 0 || width <= 0) return 0;
    const size_t tiled = ep::evrep_packed4_workspace_bytes(ev, height, width);
    if (tiled) return tiled;
    return ep_evrep_workspace_bytes(ev->batch, height, width, ev->offsets_host[ev->batch] - ev->offsets_host[0]);
}

int ep_evrep(void* stream, const ep_events_soa* ev, int height, int width, double* out, void* workspace,
             size_t workspace_bytes, unsigned int* bad_count) {
    using namespace ep;
    if (!ev || !out || !workspace || ev->batch <= 0 || height <= 0 || width <= 0) return EP_EINVAL;
    if (!ev->offsets || !ev->offsets_host) return EP_EINVAL;
    if (ev->xy_dtype == EP_U32 && ev->t_dtype == 0 && ev->t == nullptr)      // 4 B packed transport layout: route + shared-memory sweep
        return run_evrep_packed4(static_cast<cudaStream_t>(stream), ev, height, width, out, workspace, workspace_bytes, bad_count);
    if (ev->t_base || ev->xy_dtype == EP_U32 || ev->t_dtype == EP_U32) return EP_EUNSUPPORTED;   // the other transport layouts: ep_bin_events only
    if (!valid_dtype(ev->xy_dtype) || !valid_dtype(ev->t_dtype) || !valid_dtype(ev->p_dtype) || !(ev->t_div != 0.0))
        return EP_EINVAL;
    const int B = ev->batch;
    if (B > 65535) return EP_EUNSUPPORTED;                       // samples ride blockIdx.y in the scan
    const int64_t begin = ev->offsets_host[0], end = ev->offsets_host[B];
    if (end < begin) return EP_EINVAL;
    for (int b = 0; b < B; ++b)
        if (ev->offsets_host[b + 1] - ev->offsets_host[b] > 0x7fffffffLL || ev->offsets_host[b + 1] < ev->offsets_host[b])
            return EP_EINVAL;
    const RepLayout L = rep_layout(B, height, width, end - begin);
    if (workspace_bytes < L.total) return EP_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255u) return EP_EALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    RepArgs a;
    a.ev = EvDesc{ev->x, ev->y, ev->t, ev->p, ev->xy_dtype, ev->t_dtype, ev->p_dtype, ev->t_div, ev->offsets, B};
    a.H = height; a.W = width; a.begin = begin; a.end = end;
    a.cp = reinterpret_cast<unsigned long long*>(ws + L.cp);
    a.start = reinterpret_cast<int32_t*>(ws + L.start);
    a.cursor = reinterpret_cast<int32_t*>(ws + L.cursor);
    a.sums = reinterpret_cast<int32_t*>(ws + L.sums);
    a.n_chunks = L.n_chunks;
    a.sorted_t = reinterpret_cast<double*>(ws + L.sorted);
    a.bad = bad_count;
    cudaError_t ce = cudaMemsetAsync(ws, 0, L.start, st);      // the count / polarity words; start and cursor are written by the scan
    if (ce != cudaSuccess) return (int)ce;
    const int64_t n = end - begin;
    if ((int64_t)height * width > 0x7fffffffLL) return EP_EUNSUPPORTED;
    const dim3 cell_grid((unsigned)ceil_div64((int64_t)height * width, 256), (unsigned)B);      // one thread per pixel, samples on y
    if (n > 0) {
        k_evrep_hist<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(a);
        EP_LAUNCH_CHECK();
    }
    const dim3 scan_grid((unsigned)L.n_chunks, (unsigned)B);
    k_evrep_chunk_sums<<<scan_grid, kScanThreads, 0, st>>>(a);
    EP_LAUNCH_CHECK();
    k_evrep_scan_sums<<<B, kScanThreads, 0, st>>>(a);
    EP_LAUNCH_CHECK();
    k_evrep_chunk_scan<<<scan_grid, kScanThreads, 0, st>>>(a);
    EP_LAUNCH_CHECK();
    if (n > 0) {
        k_evrep_scatter<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(a);
        EP_LAUNCH_CHECK();
        k_evrep_sort<<<cell_grid, 256, 0, st>>>(a);
        EP_LAUNCH_CHECK();
    }
    k_evrep_finish<<<cell_grid, 256, 0, st>>>(a, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

}  // extern "C"

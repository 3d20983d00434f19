// SURVEY.md §8 row f2 — the host side of the ragged collate, as native multi-threaded code (no device work).
//
// The reference bins every sample on the CPU inside its DataLoader workers (e.g. pr_n_imagenet_dataset.py:85-87) and ships
// dense tensors.  Here the workers ship raw events; the collate step turns the per-sample (N,4) x,y,t,p arrays of the
// reference's event format into the canonical SoA batch, and the canonical batch into the packed transport layouts of
// include/eventpretrain_b200.h (ep_events_soa.t_base) that cross PCIe at 4-5 B/event.  Both steps stream tens of bytes per
// event; numpy does them at ~0.1 Gevents/s, a GPU bins 13 Gevents/s end to end, so they are threaded C++ here.
#include <atomic>
#include <cmath>
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#include <immintrin.h>
#define EP_HOST_AVX2 1
#endif
#include <thread>
#include <vector>

#include "ep_common.cuh"

// The streaming loops below are compiled twice on x86-64 (AVX2 and baseline) and picked at load time by the CPU.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define EP_HOST_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define EP_HOST_CLONES
#endif

namespace ep {
namespace {

int pick_threads(int threads, int64_t work_items) {
    int n = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    if ((int64_t)n > work_items) n = work_items > 0 ? (int)work_items : 1;
    return n;
}

// run fn(item) for item in [0, n_items) on `threads` threads, dynamically scheduled in runs of `grain`
template <class Fn>
void parallel_for(int64_t n_items, int64_t grain, int threads, Fn fn) {
    if (n_items <= 0) return;
    const int nt = pick_threads(threads, (n_items + grain - 1) / grain);
    if (nt == 1) {
        for (int64_t i = 0; i < n_items; ++i) fn(i);
        return;
    }
    std::atomic<int64_t> next(0);
    auto body = [&]() {
        for (;;) {
            const int64_t lo = next.fetch_add(grain, std::memory_order_relaxed);
            if (lo >= n_items) break;
            const int64_t hi = lo + grain < n_items ? lo + grain : n_items;
            for (int64_t i = lo; i < hi; ++i) fn(i);
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) pool.emplace_back(body);
    body();
    for (auto& th : pool) th.join();
}

// last sample b with offsets[b] <= i < offsets[b + 1] (empty samples skipped); offsets[0] == 0, i < offsets[B]
int owner_of(const int64_t* off, int B, int64_t i) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

#define EP_DEFINE_COLLATE(NAME, T) \
EP_HOST_CLONES bool NAME(const T* s, int64_t n, double t_scale, uint16_t* x, uint16_t* y, int64_t* t, uint8_t* p) { \
    bool ok = true; \
    for (int64_t i = 0; i < n; ++i) { \
        const double fx = (double)s[4 * i], fy = (double)s[4 * i + 1], ft = (double)s[4 * i + 2], fp = (double)s[4 * i + 3]; \
        const bool in_range = fx >= 0.0 && fx <= 65535.0 && fy >= 0.0 && fy <= 65535.0; \
        const uint16_t ux = in_range ? (uint16_t)fx : 0, uy = in_range ? (uint16_t)fy : 0; \
        ok &= in_range && (double)ux == fx && (double)uy == fy && (fp == 0.0 || fp == 1.0); \
 \
        const double v = ft * t_scale; \
        const double ticks = (v + 6755399441055744.0) - 6755399441055744.0; \
        ok &= std::fabs(v) < 2251799813685248.0; \
        x[i] = ux; y[i] = uy; p[i] = (uint8_t)(fp != 0.0); \
        t[i] = (int64_t)ticks; \
    } \
    return ok; \
}
EP_DEFINE_COLLATE(collate_f64, double)
EP_DEFINE_COLLATE(collate_f32, float)
#undef EP_DEFINE_COLLATE

#ifdef EP_HOST_AVX2
// float64 rows, four events per step: 4x4 transpose of the (x, y, t, p) rows, the same checks and the same rint trick as the
// scalar loop (the integer stamp is read off the mantissa of v + 1.5 * 2^52, exact for |v| < 2^51); tail in scalar code.
__attribute__((target("avx2"))) bool collate_f64_avx2(const double* s, int64_t n, double t_scale, uint16_t* x, uint16_t* y, int64_t* t,
                                                      uint8_t* p) {
    const __m256d zero = _mm256_setzero_pd(), one = _mm256_set1_pd(1.0), top = _mm256_set1_pd(65535.0);
    const __m256d scale = _mm256_set1_pd(t_scale), magic = _mm256_set1_pd(6755399441055744.0), lim = _mm256_set1_pd(2251799813685248.0);
    const __m256d absmask = _mm256_castsi256_pd(_mm256_set1_epi64x(0x7fffffffffffffffLL));
    const __m256i magic_bits = _mm256_castpd_si256(magic);
    __m256d ok = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const __m256d r0 = _mm256_loadu_pd(s + 4 * i), r1 = _mm256_loadu_pd(s + 4 * i + 4);
        const __m256d r2 = _mm256_loadu_pd(s + 4 * i + 8), r3 = _mm256_loadu_pd(s + 4 * i + 12);
        const __m256d a = _mm256_unpacklo_pd(r0, r1), b = _mm256_unpackhi_pd(r0, r1);      // [x0 x1 t0 t1], [y0 y1 p0 p1]
        const __m256d c = _mm256_unpacklo_pd(r2, r3), d = _mm256_unpackhi_pd(r2, r3);
        const __m256d X = _mm256_permute2f128_pd(a, c, 0x20), T = _mm256_permute2f128_pd(a, c, 0x31);
        const __m256d Y = _mm256_permute2f128_pd(b, d, 0x20), P = _mm256_permute2f128_pd(b, d, 0x31);
        const __m128i xi = _mm256_cvttpd_epi32(X), yi = _mm256_cvttpd_epi32(Y);
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_cmp_pd(X, zero, _CMP_GE_OQ), _mm256_cmp_pd(X, top, _CMP_LE_OQ)));
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_cmp_pd(Y, zero, _CMP_GE_OQ), _mm256_cmp_pd(Y, top, _CMP_LE_OQ)));
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(xi), X, _CMP_EQ_OQ), _mm256_cmp_pd(_mm256_cvtepi32_pd(yi), Y, _CMP_EQ_OQ)));
        ok = _mm256_and_pd(ok, _mm256_or_pd(_mm256_cmp_pd(P, zero, _CMP_EQ_OQ), _mm256_cmp_pd(P, one, _CMP_EQ_OQ)));
        const __m256d v = _mm256_mul_pd(T, scale);
        ok = _mm256_and_pd(ok, _mm256_cmp_pd(_mm256_and_pd(v, absmask), lim, _CMP_LT_OQ));
        const __m256i ticks = _mm256_sub_epi64(_mm256_castpd_si256(_mm256_add_pd(v, magic)), magic_bits);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(t + i), ticks);
        _mm_storel_epi64(reinterpret_cast<__m128i*>(x + i), _mm_packus_epi32(xi, xi));
        _mm_storel_epi64(reinterpret_cast<__m128i*>(y + i), _mm_packus_epi32(yi, yi));
        const __m128i pi = _mm256_cvttpd_epi32(P);
        const __m128i p16 = _mm_packus_epi32(pi, pi);
        const int p4 = _mm_cvtsi128_si32(_mm_packus_epi16(p16, p16));
        __builtin_memcpy(p + i, &p4, 4);
    }
    bool all = _mm256_movemask_pd(ok) == 0xf;
    if (i < n) all &= collate_f64(s + 4 * i, n - i, t_scale, x + i, y + i, t + i, p + i);
    return all;
}
#endif

// rint() above is the 1.5 * 2^52 trick (round-to-nearest-even, exact for |v| < 2^51): no libm call in the loop.

// Rows of one sample -> what the 4 B packed word needs: integer ticks and x | y << 11 | polarity << 22 (the collate rules of
// EP_DEFINE_COLLATE above, with the packed layout's 11-bit coordinates).  Scalar form; the float64 form below takes four rows per step.
template <typename T>
EP_HOST_CLONES bool rows_to_partial(const T* s, int64_t n, double t_scale, int64_t* ticks, uint32_t* part) {
    bool ok = true;
    for (int64_t i = 0; i < n; ++i) {
        const double fx = (double)s[4 * i], fy = (double)s[4 * i + 1], ft = (double)s[4 * i + 2], fp = (double)s[4 * i + 3];
        const bool in_range = fx >= 0.0 && fx <= 2047.0 && fy >= 0.0 && fy <= 2047.0;
        const uint32_t ux = in_range ? (uint32_t)fx : 0u, uy = in_range ? (uint32_t)fy : 0u;
        ok &= in_range && (double)ux == fx && (double)uy == fy && (fp == 0.0 || fp == 1.0);
        const double v = ft * t_scale;
        const double tk = (v + 6755399441055744.0) - 6755399441055744.0;
        ok &= std::fabs(v) < 2251799813685248.0;
        ticks[i] = (int64_t)tk;
        part[i] = ux | (uy << 11) | ((uint32_t)(fp != 0.0) << 22);
    }
    return ok;
}

#ifdef EP_HOST_AVX2
__attribute__((target("avx2"))) bool rows_to_partial_f64_avx2(const double* s, int64_t n, double t_scale, int64_t* ticks, uint32_t* part) {
    const __m256d zero = _mm256_setzero_pd(), one = _mm256_set1_pd(1.0), top = _mm256_set1_pd(2047.0);
    const __m256d scale = _mm256_set1_pd(t_scale), magic = _mm256_set1_pd(6755399441055744.0), lim = _mm256_set1_pd(2251799813685248.0);
    const __m256d absmask = _mm256_castsi256_pd(_mm256_set1_epi64x(0x7fffffffffffffffLL));
    const __m256i magic_bits = _mm256_castpd_si256(magic);
    __m256d ok = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));
    int64_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const __m256d r0 = _mm256_loadu_pd(s + 4 * i), r1 = _mm256_loadu_pd(s + 4 * i + 4);
        const __m256d r2 = _mm256_loadu_pd(s + 4 * i + 8), r3 = _mm256_loadu_pd(s + 4 * i + 12);
        const __m256d a = _mm256_unpacklo_pd(r0, r1), b = _mm256_unpackhi_pd(r0, r1);      // [x0 x1 t0 t1], [y0 y1 p0 p1]
        const __m256d c = _mm256_unpacklo_pd(r2, r3), d = _mm256_unpackhi_pd(r2, r3);
        const __m256d X = _mm256_permute2f128_pd(a, c, 0x20), T = _mm256_permute2f128_pd(a, c, 0x31);
        const __m256d Y = _mm256_permute2f128_pd(b, d, 0x20), P = _mm256_permute2f128_pd(b, d, 0x31);
        const __m128i xi = _mm256_cvttpd_epi32(X), yi = _mm256_cvttpd_epi32(Y), pi = _mm256_cvttpd_epi32(P);
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_cmp_pd(X, zero, _CMP_GE_OQ), _mm256_cmp_pd(X, top, _CMP_LE_OQ)));
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_cmp_pd(Y, zero, _CMP_GE_OQ), _mm256_cmp_pd(Y, top, _CMP_LE_OQ)));
        ok = _mm256_and_pd(ok, _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtepi32_pd(xi), X, _CMP_EQ_OQ), _mm256_cmp_pd(_mm256_cvtepi32_pd(yi), Y, _CMP_EQ_OQ)));
        ok = _mm256_and_pd(ok, _mm256_or_pd(_mm256_cmp_pd(P, zero, _CMP_EQ_OQ), _mm256_cmp_pd(P, one, _CMP_EQ_OQ)));
        const __m256d v = _mm256_mul_pd(T, scale);
        ok = _mm256_and_pd(ok, _mm256_cmp_pd(_mm256_and_pd(v, absmask), lim, _CMP_LT_OQ));
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(ticks + i), _mm256_sub_epi64(_mm256_castpd_si256(_mm256_add_pd(v, magic)), magic_bits));
        const __m128i word = _mm_or_si128(_mm_and_si128(xi, _mm_set1_epi32(0x7ff)),
                                          _mm_or_si128(_mm_slli_epi32(_mm_and_si128(yi, _mm_set1_epi32(0x7ff)), 11),
                                                       _mm_slli_epi32(_mm_and_si128(pi, _mm_set1_epi32(1)), 22)));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(part + i), word);
    }
    bool all = _mm256_movemask_pd(ok) == 0xf;
    if (i < n) all &= rows_to_partial<double>(s + 4 * i, n - i, t_scale, ticks + i, part + i);
    return all;
}
#endif


EP_HOST_CLONES int64_t min_run(const int64_t* t, int64_t i0, int64_t i1) {
    int64_t m = t[i0];
    for (int64_t i = i0 + 1; i < i1; ++i) m = t[i] < m ? t[i] : m;
    return m;
}

// one branch-free run of events of one sample inside a block; returns the OR of all range violations
EP_HOST_CLONES uint64_t pack_run5(const uint16_t* x, const uint16_t* y, const int64_t* t, const uint8_t* p, int64_t lo, int64_t hi,
                                  int64_t sub, uint32_t* w, uint8_t* tick_low) {
    uint64_t viol = 0;
    for (int64_t i = lo; i < hi; ++i) {
        const uint64_t ticks = (uint64_t)(t[i] - sub);
        const uint32_t xi = x[i], yi = y[i], pi = p[i], tk = (uint32_t)ticks;
        viol |= (ticks >> 17) | (uint64_t)((xi | yi) >> 11) | (uint64_t)(pi >> 1);
        w[i] = xi | (yi << 11) | (pi << 22) | ((tk >> 8) << 23);
        tick_low[i] = (uint8_t)(tk & 0xffu);
    }
    return viol;
}

EP_HOST_CLONES uint64_t pack_run4(const uint16_t* x, const uint16_t* y, const int64_t* t, const uint8_t* p, int64_t lo, int64_t hi,
                                  int64_t sub, uint32_t* w) {
    uint64_t viol = 0;
    for (int64_t i = lo; i < hi; ++i) {
        const uint64_t ticks = (uint64_t)(t[i] - sub);
        const uint32_t xi = x[i], yi = y[i], pi = p[i], tk = (uint32_t)ticks;
        viol |= (ticks >> 9) | (uint64_t)((xi | yi) >> 11) | (uint64_t)(pi >> 1);
        w[i] = xi | (yi << 11) | (pi << 22) | (tk << 23);
    }
    return viol;
}

EP_HOST_CLONES uint64_t pack_run8(const int64_t* t, const uint8_t* p, int64_t lo, int64_t hi, int64_t base, uint32_t* w) {
    uint64_t viol = 0;
    for (int64_t i = lo; i < hi; ++i) {
        const uint64_t rel = (uint64_t)(t[i] - base);
        const uint32_t pi = p[i];
        viol |= (rel >> 31) | (uint64_t)(pi >> 1);
        w[i] = (uint32_t)rel | (pi << 31);
    }
    return viol;
}

}  // namespace
}  // namespace ep

extern "C" {

int ep_collate_aos_host(const void* const* samples, const int64_t* counts, int batch, int dtype, double t_scale, uint16_t* x,
                        uint16_t* y, int64_t* t, uint8_t* p, int64_t* offsets, int threads) {
    using namespace ep;
    if (!samples || !counts || !offsets || batch <= 0) return EP_EINVAL;
    if (dtype != EP_F64 && dtype != EP_F32) return EP_EINVAL;
    if (!(t_scale > 0.0)) return EP_EINVAL;
    offsets[0] = 0;
    for (int b = 0; b < batch; ++b) {
        if (counts[b] < 0 || (counts[b] > 0 && !samples[b])) return EP_EINVAL;
        offsets[b + 1] = offsets[b] + counts[b];
    }
    if (offsets[batch] == 0) return EP_OK;                      // a batch of empty samples: nothing to write
    if (!x || !y || !t || !p) return EP_EINVAL;
    // work items = (sample, 64K-event piece) so that a few long samples still spread over all threads
    constexpr int64_t kPiece = 1 << 16;
    std::vector<int64_t> first_piece((size_t)batch + 1, 0);
    for (int b = 0; b < batch; ++b) first_piece[b + 1] = first_piece[b] + (counts[b] + kPiece - 1) / kPiece;
    std::atomic<int> bad(0);
    parallel_for(first_piece[batch], 1, threads, [&](int64_t item) {
        int lo = 0, hi = batch;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (first_piece[mid] <= item) lo = mid; else hi = mid;
        }
        const int b = lo;
        const int64_t i0 = (item - first_piece[b]) * kPiece;
        const int64_t n = counts[b] - i0 < kPiece ? counts[b] - i0 : kPiece;
        const int64_t o = offsets[b] + i0;
#ifdef EP_HOST_AVX2
        static const bool have_avx2 = __builtin_cpu_supports("avx2");
        if (dtype == EP_F64 && have_avx2) {
            if (!collate_f64_avx2(static_cast<const double*>(samples[b]) + 4 * i0, n, t_scale, x + o, y + o, t + o, p + o))
                bad.store(1, std::memory_order_relaxed);
            return;
        }
#endif
        const bool ok = dtype == EP_F64
            ? collate_f64(static_cast<const double*>(samples[b]) + 4 * i0, n, t_scale, x + o, y + o, t + o, p + o)
            : collate_f32(static_cast<const float*>(samples[b]) + 4 * i0, n, t_scale, x + o, y + o, t + o, p + o);
        if (!ok) bad.store(1, std::memory_order_relaxed);
    });
    return bad.load() ? EP_EUNSUPPORTED : EP_OK;
}

int ep_pack_transport_host(const uint16_t* x, const uint16_t* y, const int64_t* t, const uint8_t* p, const int64_t* offsets,
                           int batch, int nbytes, uint32_t* w, uint8_t* tick_low, uint32_t* blk_base, int64_t* t_base,
                           int threads) {
    using namespace ep;
    if (!offsets || !t_base || batch <= 0) return EP_EINVAL;
    if (nbytes != 4 && nbytes != 5 && nbytes != 8) return EP_EINVAL;
    if (nbytes != 8 ? offsets[0] != 0 : offsets[0] < 0) return EP_EINVAL;      // block offsets are tied to array positions
    for (int b = 0; b < batch; ++b)
        if (offsets[b + 1] < offsets[b]) return EP_EINVAL;
    if (offsets[batch] == offsets[0]) {                          // a batch of empty samples: only the bases exist
        for (int b = 0; b < batch; ++b) t_base[b] = 0;
        return EP_OK;
    }
    if (!t || !p || !w) return EP_EINVAL;
    if (nbytes != 8 && (!x || !y || !blk_base)) return EP_EINVAL;
    if (nbytes == 5 && !tick_low) return EP_EINVAL;
    const int64_t n = offsets[batch];
    const int64_t K = nbytes == 5 ? 1024 : 256;
    const int tick_bits = nbytes == 5 ? 17 : 9;
    const int64_t n_blocks = (n + K - 1) / K;
    // per-sample base = smallest stamp (0 for an empty sample); partial minima per (sample, piece), then combined
    constexpr int64_t kPiece = 1 << 16;
    std::vector<int64_t> first_piece((size_t)batch + 1, 0);
    for (int b = 0; b < batch; ++b) first_piece[b + 1] = first_piece[b] + (offsets[b + 1] - offsets[b] + kPiece - 1) / kPiece;
    std::vector<int64_t> part((size_t)first_piece[batch], 0);
    parallel_for(first_piece[batch], 1, threads, [&](int64_t item) {
        int lo = 0, hi = batch;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (first_piece[mid] <= item) lo = mid; else hi = mid;
        }
        const int64_t i0 = offsets[lo] + (item - first_piece[lo]) * kPiece;
        const int64_t i1 = i0 + kPiece < offsets[lo + 1] ? i0 + kPiece : offsets[lo + 1];
        part[(size_t)item] = min_run(t, i0, i1);
    });
    for (int b = 0; b < batch; ++b) {
        int64_t m = 0;
        for (int64_t q = first_piece[b]; q < first_piece[b + 1]; ++q) m = (q == first_piece[b] || part[(size_t)q] < m) ? part[(size_t)q] : m;
        t_base[b] = m;
    }
    std::atomic<int> bad(0);
    if (nbytes == 8) {      // compact layout: x, y stay as they are; one word = ticks since the sample's base | polarity << 31
        parallel_for(first_piece[batch], 1, threads, [&](int64_t item) {
            int lo = 0, hi = batch;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (first_piece[mid] <= item) lo = mid; else hi = mid;
            }
            const int64_t i0 = offsets[lo] + (item - first_piece[lo]) * kPiece;
            const int64_t i1 = i0 + kPiece < offsets[lo + 1] ? i0 + kPiece : offsets[lo + 1];
            if (pack_run8(t, p, i0, i1, t_base[lo], w)) bad.store(1, std::memory_order_relaxed);
        });
        return bad.load() ? EP_EUNSUPPORTED : EP_OK;
    }
    parallel_for(n_blocks, 64, threads, [&](int64_t g) {
        const int64_t i0 = g * K, i1 = i0 + K < n ? i0 + K : n;
        int b = owner_of(offsets, batch, i0);
        while (offsets[b + 1] <= i0) ++b;      // (owner_of lands on the last sample starting at or before i0: never empty, kept as a guard)
        // tick offset of the block: smallest relative stamp among the events of the sample that owns the block's first slot
        const int64_t own_end = offsets[b + 1] < i1 ? offsets[b + 1] : i1;
        int64_t m = min_run(t, i0, own_end) - t_base[b];
        bool ok = m >= 0 && m < ((int64_t)1 << 32);
        const int64_t add_own = ok ? m : 0;
        blk_base[g] = (uint32_t)add_own;
        // one branch-free run per sample inside the block (a sample that starts here counts from its own base)
        uint64_t viol = 0;
        for (int64_t lo = i0; lo < i1;) {
            while (lo >= offsets[b + 1]) ++b;
            const int64_t hi = offsets[b + 1] < i1 ? offsets[b + 1] : i1;
            const int64_t sub = t_base[b] + (offsets[b] / K == g ? 0 : add_own);
            viol |= nbytes == 5 ? pack_run5(x, y, t, p, lo, hi, sub, w, tick_low) : pack_run4(x, y, t, p, lo, hi, sub, w);
            lo = hi;
        }
        if (!ok || viol) bad.store(1, std::memory_order_relaxed);
    });
    return bad.load() ? EP_EUNSUPPORTED : EP_OK;
}


int ep_collate_transport4_host(const void* const* samples, const int64_t* counts, int batch, int dtype, double t_scale, uint32_t* w,
                               uint32_t* blk_base, int64_t* t_base, int64_t* offsets, int threads) {
    using namespace ep;
    if (!samples || !counts || !offsets || !t_base || batch <= 0) return EP_EINVAL;
    if (dtype != EP_F64 && dtype != EP_F32) return EP_EINVAL;
    if (!(t_scale > 0.0)) return EP_EINVAL;
    offsets[0] = 0;
    for (int b = 0; b < batch; ++b) {
        if (counts[b] < 0 || (counts[b] > 0 && !samples[b])) return EP_EINVAL;
        offsets[b + 1] = offsets[b] + counts[b];
    }
    const int64_t n = offsets[batch];
    // the sample's base is its first row's stamp: the smallest one for a time-sorted sample (an earlier stamp further down
    // shows up as a negative tick below and sends the batch to the two-step path, which takes the true minimum)
    std::atomic<int> bad(0);
    for (int b = 0; b < batch; ++b) {
        t_base[b] = 0;
        if (counts[b] > 0) {
            const double ft = dtype == EP_F64 ? static_cast<const double*>(samples[b])[2] : (double)static_cast<const float*>(samples[b])[2];
            const double v = ft * t_scale;
            if (!(std::fabs(v) < 2251799813685248.0)) return EP_EUNSUPPORTED;
            t_base[b] = (int64_t)((v + 6755399441055744.0) - 6755399441055744.0);
        }
    }
    if (n == 0) return EP_OK;
    if (!w || !blk_base) return EP_EINVAL;
    constexpr int64_t K = 256;
    const int64_t n_blocks = (n + K - 1) / K;
#ifdef EP_HOST_AVX2
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
#endif
    parallel_for(n_blocks, 64, threads, [&](int64_t g) {
        const int64_t i0 = g * K, i1 = i0 + K < n ? i0 + K : n;
        int64_t ticks[K];
        uint32_t part[K];
        int b0 = owner_of(offsets, batch, i0);
        while (offsets[b0 + 1] <= i0) ++b0;
        bool ok = true;
        // rows of every sample inside the block -> ticks and partial words (one pass over the 32 B/event rows)
        int b = b0;
        for (int64_t lo = i0; lo < i1;) {
            while (lo >= offsets[b + 1]) ++b;
            const int64_t hi = offsets[b + 1] < i1 ? offsets[b + 1] : i1;
            const int64_t r0 = lo - offsets[b], cnt = hi - lo;
#ifdef EP_HOST_AVX2
            if (dtype == EP_F64 && have_avx2)
                ok &= rows_to_partial_f64_avx2(static_cast<const double*>(samples[b]) + 4 * r0, cnt, t_scale, ticks + (lo - i0), part + (lo - i0));
            else
#endif
            ok &= dtype == EP_F64 ? rows_to_partial<double>(static_cast<const double*>(samples[b]) + 4 * r0, cnt, t_scale, ticks + (lo - i0), part + (lo - i0))
                                  : rows_to_partial<float>(static_cast<const float*>(samples[b]) + 4 * r0, cnt, t_scale, ticks + (lo - i0), part + (lo - i0));
            lo = hi;
        }
        // tick offset of the block: smallest relative stamp among the events of the sample that owns the block's first slot
        const int64_t own_end = offsets[b0 + 1] < i1 ? offsets[b0 + 1] : i1;
        int64_t m = ticks[0];
        for (int64_t i = 1; i < own_end - i0; ++i) m = ticks[i] < m ? ticks[i] : m;
        m -= t_base[b0];
        ok &= m >= 0 && m < ((int64_t)1 << 32);
        const int64_t add_own = ok ? m : 0;
        blk_base[g] = (uint32_t)add_own;
        uint64_t viol = 0;
        b = b0;
        for (int64_t lo = i0; lo < i1;) {
            while (lo >= offsets[b + 1]) ++b;
            const int64_t hi = offsets[b + 1] < i1 ? offsets[b + 1] : i1;
            const int64_t sub = t_base[b] + (offsets[b] / K == g ? 0 : add_own);
            for (int64_t i = lo; i < hi; ++i) {
                const uint64_t rel = (uint64_t)(ticks[i - i0] - sub);
                viol |= rel >> 9;
                w[i] = part[i - i0] | ((uint32_t)rel << 23);
            }
            lo = hi;
        }
        if (!ok || viol) bad.store(1, std::memory_order_relaxed);
    });
    return bad.load() ? EP_EUNSUPPORTED : EP_OK;
}

}  // extern "C"

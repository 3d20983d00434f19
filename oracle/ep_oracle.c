/*
 * ep_oracle.c — CPU restatement of EventPretrain's event -> dense-tensor routines.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under eventpretrain_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker (or as the timed CPU port), never as the product path.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the unmodified reference executed in the build
 * container (tests/golden/make_golden.py -> tests/golden/stage1_events.npz; checked by
 * tests/test_oracle_golden.py, bit-exact for every function below).
 *
 * Each function cites the reference file:line (relative to the EventPretrain root) it follows.
 * The arithmetic (dtype of every intermediate, order of every floating-point accumulation)
 * is that of the reference's torch / numpy CPU ops:
 *   - Tensor.index_add_ on a 1-D fp32 tensor accumulates sequentially in index order;
 *   - torch.bincount counts into int64 and .float() rounds to nearest;
 *   - np.add.at accumulates sequentially in the array's dtype.
 *
 * Return codes: 0 ok; EP_ORACLE_EINDEX (-2) where the reference raises IndexError /
 * RuntimeError for an out-of-range flat index; EP_ORACLE_EINVAL (-1) for bad arguments
 * (n == 0 raises IndexError in the reference at events[0, 2]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EP_ORACLE_EINVAL (-1)
#define EP_ORACLE_EINDEX (-2)

/* ------------------------------------------------------------------------------------------
 * events_to_voxel_grid  (dataset/dataset_utils/events_to_voxel_grid.py:4-61)
 * events: (n,4) AoS, columns x,y,t,p.  fp64 events -> time arithmetic in fp64; fp32 events ->
 * time arithmetic in fp32 (torch keeps the tensor dtype).  Weights are fp32 in both cases.
 * Two passes, exactly like the two index_add_ calls: all left contributions in event order
 * (:44-49), then all right contributions (:51-57).
 * ---------------------------------------------------------------------------------------- */
#define VOXEL_IMPL(NAME, T, FLOOR)                                                              \
    int NAME(const T* ev, int64_t n, int num_bins, int H, int W, float* out) {                  \
        if (n <= 0 || num_bins <= 0 || H <= 0 || W <= 0) return EP_ORACLE_EINVAL;               \
        const int64_t plane = (int64_t)H * W, total = plane * num_bins;                         \
        memset(out, 0, sizeof(float) * (size_t)total);                                          \
        const T first = ev[2], last = ev[(n - 1) * 4 + 2];                /* :19-20 */          \
        T deltaT = last - first;                                          /* :22 */             \
        if (deltaT == 0) deltaT = (T)1.0;                                 /* :24-25 */          \
        const T scale = (T)(num_bins - 1);                                                      \
        /* the reference materialises these per-event vectors too (:32-42) */                   \
        int64_t* il = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);                            \
        float* vl = (float*)malloc(sizeof(float) * (size_t)n);                                  \
        float* vr = (float*)malloc(sizeof(float) * (size_t)n);                                  \
        unsigned char* ok = (unsigned char*)malloc((size_t)n);                                  \
        if (!il || !vl || !vr || !ok) { free(il); free(vl); free(vr); free(ok); return EP_ORACLE_EINVAL; } \
        for (int64_t i = 0; i < n; ++i) {                                                       \
            const T* e = ev + i * 4;                                                            \
            const int64_t xs = (int64_t)e[0], ys = (int64_t)e[1];         /* :32-33 trunc */    \
            const T ts = scale * (e[2] - first) / deltaT;                 /* :34 */             \
            float ps = (float)e[3];                                       /* :35 */             \
            if (ps == 0.0f) ps = -1.0f;                                   /* :36 */             \
            const T tis = FLOOR(ts);                                      /* :38 */             \
            const float dts = (float)(ts - tis);                          /* :40-42 .float() */ \
            vl[i] = ps * (1.0f - dts);                                    /* :41 */             \
            vr[i] = ps * dts;                                             /* :42 */             \
            const int in_left = (tis < (T)num_bins) && (tis >= 0);        /* :44-45 */          \
            const int in_right = ((tis + 1) < (T)num_bins) && (tis >= 0); /* :51-52 */          \
            ok[i] = (unsigned char)(in_left | (in_right << 1));                                 \
            il[i] = in_left ? xs + ys * W + (int64_t)tis * plane : 0;     /* :47-48 */          \
        }                                                                                       \
        int rc = 0;                                                                             \
        for (int64_t i = 0; i < n && rc == 0; ++i) {      /* first index_add_, event order */   \
            if (!(ok[i] & 1)) continue;                                                         \
            if (il[i] < 0 || il[i] >= total) rc = EP_ORACLE_EINDEX;       /* index_add_ raises */\
            else out[il[i]] += vl[i];                                     /* fp32, in order */  \
        }                                                                                       \
        for (int64_t i = 0; i < n && rc == 0; ++i) {      /* second index_add_ */               \
            if (!(ok[i] & 2)) continue;                                                         \
            const int64_t idx = il[i] + plane;                            /* :55-56 */          \
            if (idx < 0 || idx >= total) rc = EP_ORACLE_EINDEX;                                 \
            else out[idx] += vr[i];                                                             \
        }                                                                                       \
        free(il); free(vl); free(vr); free(ok);                                                 \
        return rc;                                                                              \
    }

VOXEL_IMPL(oracle_voxel_grid_f64, double, floor)
VOXEL_IMPL(oracle_voxel_grid_f32, float, floorf)

/* ------------------------------------------------------------------------------------------
 * events_to_image_ecdp / events_to_image_mem  (dataset/dataset_utils/events_to_image.py:6-62)
 * channels == 2 -> [pos, neg]; channels == 3 -> [pos, 0, neg] (the zero "tss" plane, :58-59).
 * pos = rows with p == 1; neg = rows with p == 0, or p == -1 when no row has p == 0 (:13-16).
 * ---------------------------------------------------------------------------------------- */
#define COUNT_IMPL(NAME, T)                                                                     \
    int NAME(const T* ev, int64_t n, int H, int W, int channels, float* out) {                  \
        if (n < 0 || H <= 0 || W <= 0 || (channels != 2 && channels != 3)) return EP_ORACLE_EINVAL; \
        const int64_t plane = (int64_t)H * W;                                                   \
        int64_t* cnt = (int64_t*)calloc((size_t)(2 * plane), sizeof(int64_t));                  \
        if (!cnt) return EP_ORACLE_EINVAL;                                                      \
        int any_zero = 0;                                                                       \
        for (int64_t i = 0; i < n; ++i) any_zero |= (ev[i * 4 + 3] == (T)0);                    \
        const T negval = any_zero ? (T)0 : (T)-1;                                               \
        int rc = 0;                                                                             \
        for (int64_t i = 0; i < n && rc == 0; ++i) {                                            \
            const T* e = ev + i * 4;                                                            \
            int ch;                                                                             \
            if (e[3] == (T)1) ch = 0; else if (e[3] == negval) ch = 1; else continue;           \
            const int64_t idx = (int64_t)e[0] + (int64_t)e[1] * W;        /* :24-27 */          \
            if (idx < 0 || idx >= plane) rc = EP_ORACLE_EINDEX;           /* bincount/reshape raise */ \
            else cnt[ch * plane + idx] += 1;                                                    \
        }                                                                                       \
        if (rc == 0) {                                                                          \
            float* pos = out;                                                                   \
            float* neg = out + (channels - 1) * plane;                                          \
            if (channels == 3) memset(out + plane, 0, sizeof(float) * (size_t)plane);           \
            for (int64_t j = 0; j < plane; ++j) { pos[j] = (float)cnt[j]; neg[j] = (float)cnt[plane + j]; } \
        }                                                                                       \
        free(cnt);                                                                              \
        return rc;                                                                              \
    }

COUNT_IMPL(oracle_count_frame_f64, double)
COUNT_IMPL(oracle_count_frame_f32, float)

/* ------------------------------------------------------------------------------------------
 * remove_hot_pixel_mem  (dataset/dataset_utils/events_to_image.py:65-75)
 * hist: (3,H,W) fp32, modified in place.  threshold = mean + num_stds * std over channels 0 and
 * 2 jointly, std unbiased (torch.std default).  torch reduces fp32 means/variances with a
 * double accumulator on CPU; the comparison is done in fp32.
 * ---------------------------------------------------------------------------------------- */
void oracle_remove_hot_pixel_mem(float* hist, int H, int W, float num_stds) {
    const int64_t plane = (int64_t)H * W, m = 2 * plane;
    const float* c0 = hist;
    const float* c2 = hist + 2 * plane;
    double s = 0.0;
    for (int64_t j = 0; j < plane; ++j) s += c0[j];
    for (int64_t j = 0; j < plane; ++j) s += c2[j];
    const double mean = s / (double)m;
    double ss = 0.0;
    for (int64_t j = 0; j < plane; ++j) { double d = c0[j] - mean; ss += d * d; }
    for (int64_t j = 0; j < plane; ++j) { double d = c2[j] - mean; ss += d * d; }
    const float meanf = (float)mean;
    const float stdf = (float)sqrt(ss / (double)(m - 1));
    const float thr = meanf + num_stds * stdf;                           /* :69 */
    float* w0 = hist;
    float* w2 = hist + 2 * plane;
    for (int64_t j = 0; j < plane; ++j) {
        if (w0[j] > thr || w2[j] > thr) { w0[j] = 0.0f; w2[j] = 0.0f; }   /* :70-73 */
    }
}

/* per-sample normalisers applied by the datasets right after binning:
 *   count frame: x / (amax_hw(x) + 1), then (x - 0.5) * 2
 *     (dataset/pretrain/pr_n_imagenet_dataset.py:142-143, finetune_cls/ft_n_caltech101_dataset.py:93-95)
 *   MEM frame: channels 0,2 scaled by 1 / max(ch0, ch2); MVSEC guards a zero max with 1/0.001
 *     (finetune_cls/ft_n_caltech101_dataset.py:96-98, finetune_flow/ft_mvsec_dataset.py:244-249) */
void oracle_count_normalise(float* img, int C, int H, int W) {
    const int64_t plane = (int64_t)H * W;
    for (int c = 0; c < C; ++c) {
        float* p = img + c * plane;
        float mx = p[0];
        for (int64_t j = 1; j < plane; ++j) if (p[j] > mx) mx = p[j];
        const float den = mx + 1.0f;
        for (int64_t j = 0; j < plane; ++j) p[j] = (p[j] / den - 0.5f) * 2.0f;
    }
}

void oracle_mem_normalise(float* img, int H, int W, int guard_zero) {
    const int64_t plane = (int64_t)H * W;
    float* c0 = img;
    float* c2 = img + 2 * plane;
    float mx = c0[0];
    for (int64_t j = 0; j < plane; ++j) { if (c0[j] > mx) mx = c0[j]; if (c2[j] > mx) mx = c2[j]; }
    /* 1.0 / tensor is an fp32 division; the MVSEC guard is the Python double 1.0 / 0.001 times fp32 */
    const float factor = (guard_zero && mx == 0.0f) ? (float)(1.0 / 0.001) : 1.0f / mx;
    for (int64_t j = 0; j < plane; ++j) { c0[j] = c0[j] * factor; c2[j] = c2[j] * factor; }
}

/* ------------------------------------------------------------------------------------------
 * events_to_EvRep  (dataset/dataset_utils/events_to_image.py:77-125)
 * xs, ys int16 (callers cast, e.g. finetune_cls/ft_n_caltech101_dataset.py:79-80); ts, ps fp64.
 * out: (3,H,W) fp64 = [E_C, E_I, E_T].  E_T accumulators are fp32 arrays (:94-95) but np.add.at with
 * fp64 operands adds in fp64 and rounds the sum back to fp32 at every step (:113-114; verified
 * against numpy 2.3.5), accumulated in lexsort((t, y, x)) order (:104); the statistics after that
 * are fp64 because float32 / int32 promotes to float64 in numpy (:117-120).
 * Negative coordinates wrap like numpy fancy indexing; anything else out of range is an
 * IndexError there.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int16_t x, y; double t; int64_t i; } evkey_t;

static int evkey_cmp(const void* a, const void* b) {
    const evkey_t* p = (const evkey_t*)a;
    const evkey_t* q = (const evkey_t*)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    if (p->t != q->t) return p->t < q->t ? -1 : 1;
    return p->i < q->i ? -1 : (p->i > q->i);   /* lexsort is stable */
}

int oracle_evrep(const int16_t* xs, const int16_t* ys, const double* ts, const double* ps,
                 int64_t n, int W, int H, double* out) {
    if (n <= 0 || H <= 0 || W <= 0) return EP_ORACLE_EINVAL;   /* sorted_timestamps[0] raises on n == 0 */
    const int64_t plane = (int64_t)H * W;
    int32_t* ec = (int32_t*)calloc((size_t)plane, sizeof(int32_t));
    int32_t* ei = (int32_t*)calloc((size_t)plane, sizeof(int32_t));
    float* tsum = (float*)calloc((size_t)plane, sizeof(float));
    float* tsq = (float*)calloc((size_t)plane, sizeof(float));
    evkey_t* keys = (evkey_t*)malloc(sizeof(evkey_t) * (size_t)n);
    int rc = 0;
    for (int64_t i = 0; i < n && rc == 0; ++i) {
        int64_t x = xs[i], y = ys[i];
        if (x < 0) x += W;
        if (y < 0) y += H;
        if (x < 0 || x >= W || y < 0 || y >= H) { rc = EP_ORACLE_EINDEX; break; }
        const double p = (ps[i] == 0.0) ? -1.0 : ps[i];                  /* :97 */
        ec[y * W + x] += 1;                                               /* :100 */
        ei[y * W + x] += (int32_t)p;                                      /* :101 */
        keys[i].x = xs[i]; keys[i].y = ys[i]; keys[i].t = ts[i]; keys[i].i = i;
    }
    if (rc == 0) {
        qsort(keys, (size_t)n, sizeof(evkey_t), evkey_cmp);               /* :104 */
        double prev = keys[0].t;                                          /* prepend=sorted[0], :110 */
        for (int64_t k = 0; k < n; ++k) {
            int64_t x = keys[k].x, y = keys[k].y;
            if (x < 0) x += W;
            if (y < 0) y += H;
            const double d = keys[k].t - prev;
            prev = keys[k].t;
            tsum[y * W + x] = (float)((double)tsum[y * W + x] + d);      /* :113 */
            tsq[y * W + x] = (float)((double)tsq[y * W + x] + d * d);    /* :114 */
        }
        for (int64_t j = 0; j < plane; ++j) {
            const int32_t c = ec[j] < 1 ? 1 : ec[j];                      /* :117 */
            const double mean = (double)tsum[j] / (double)c;              /* :118 */
            double v = (double)tsq[j] / (double)c - mean * mean;          /* :119 */
            if (!(v > 0.0)) v = (v != v) ? v : 0.0;                       /* np.maximum(v, 0) propagates NaN */
            double et = sqrt(v);
            if (et > 1000.0) et = 1000.0;                                 /* :120 */
            out[j] = (double)ec[j];
            out[plane + j] = (double)ei[j];
            out[2 * plane + j] = et;
        }
    }
    free(keys); free(tsq); free(tsum); free(ei); free(ec);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Batched drivers used as the timed CPU port (bench.py cpu_baseline / --impl reference):
 * one sample per worker-thread task, which is how the reference runs these routines (one sample
 * per DataLoader worker, main_pretrain.py:12,236-243).  offsets: (B+1) exclusive scan of counts.
 * Plain pthreads (libgomp is not in the image); samples are claimed from an atomic counter.
 * ---------------------------------------------------------------------------------------- */
#include <pthread.h>

typedef struct {
    const double* ev; const int64_t* offsets; int B, num_bins, H, W, channels; float* out;
    int next; int rc; int kind;   /* kind 0 = voxel, 1 = count frame */
} batch_job_t;

static void* batch_worker(void* arg) {
    batch_job_t* j = (batch_job_t*)arg;
    for (;;) {
        const int b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->B) break;
        const double* e = j->ev + j->offsets[b] * 4;
        const int64_t n = j->offsets[b + 1] - j->offsets[b];
        int r;
        if (j->kind == 0)
            r = oracle_voxel_grid_f64(e, n, j->num_bins, j->H, j->W, j->out + (int64_t)b * j->num_bins * j->H * j->W);
        else
            r = oracle_count_frame_f64(e, n, j->H, j->W, j->channels, j->out + (int64_t)b * j->channels * j->H * j->W);
        if (r != 0) __atomic_store_n(&j->rc, r, __ATOMIC_RELAXED);
    }
    return NULL;
}

static int run_batch(batch_job_t* job, int num_threads) {
    if (num_threads < 1) num_threads = 1;
    if (num_threads > 256) num_threads = 256;
    pthread_t th[256];
    int started = 0;
    for (int i = 1; i < num_threads; ++i)
        if (pthread_create(&th[started], NULL, batch_worker, job) == 0) ++started;
    batch_worker(job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    return job->rc;
}

int oracle_voxel_grid_batch_f64(const double* ev, const int64_t* offsets, int B, int num_bins,
                                int H, int W, float* out, int num_threads) {
    batch_job_t job = {ev, offsets, B, num_bins, H, W, 0, out, 0, 0, 0};
    return run_batch(&job, num_threads);
}

int oracle_count_frame_batch_f64(const double* ev, const int64_t* offsets, int B, int H, int W,
                                 int channels, float* out, int num_threads) {
    batch_job_t job = {ev, offsets, B, 0, H, W, channels, out, 0, 0, 1};
    return run_batch(&job, num_threads);
}

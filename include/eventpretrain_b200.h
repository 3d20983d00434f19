/*
 * eventpretrain_b200.h — C ABI of the B200-native EventPretrain input hot path.
 *
 * The reference (BIT-Vision/EventPretrain) is pure Python: it has no plugin / operator / FFI layer
 * (SURVEY.md F3).  Its boundary for this path is a set of Python call signatures; this header is
 * the C ABI underneath the drop-in Python functions in eventpretrain_b200/ that keep those
 * signatures.  Every entry point cites the reference function (file:line relative to the
 * EventPretrain root) whose computation it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, POD structs only; no torch / C++ types.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous,
 *     never synchronise the device, never allocate; the caller owns every buffer.  All data
 *     pointers are DEVICE pointers on the current device unless the name ends in `_host`.
 *   - Outputs are fully overwritten.  Scratch is passed explicitly; query its size with the
 *     matching *_workspace_bytes().
 *   - Return: EP_OK (0); a negative EP_E* code for argument errors; a positive cudaError_t value
 *     if a launch failed.  ep_status_string() renders either.
 *   - Thread-safe and re-entrant (no global state); one process per GPU.
 *   - Out-of-range events are never clamped or silently dropped (the reference raises IndexError /
 *     RuntimeError for them): kernels skip them and add to the optional device counter
 *     `bad_count`, which the Python shim turns into IndexError when checking is enabled.
 */
#ifndef EVENTPRETRAIN_B200_H
#define EVENTPRETRAIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EP_ABI_VERSION 1

#if defined(__GNUC__)
#define EP_API __attribute__((visibility("default")))
#else
#define EP_API
#endif

#define EP_OK 0
#define EP_EINVAL (-1)       /* bad argument (null pointer, non-positive size, unsupported dtype tag) */
#define EP_EWORKSPACE (-2)   /* workspace smaller than *_workspace_bytes() minimum */
#define EP_EUNSUPPORTED (-3) /* shape outside what the kernels were built for */
#define EP_EALIGN (-4)       /* pointer not aligned as documented */

/* dtype tags for ragged event buffers */
enum ep_dtype {
    EP_U8 = 1, EP_I8 = 2, EP_U16 = 3, EP_I16 = 4, EP_I32 = 5, EP_I64 = 6, EP_F32 = 7, EP_F64 = 8, EP_U32 = 9
};

EP_API int ep_abi_version(void);
EP_API const char* ep_status_string(int status);

/* Instrumentation (bench.py): kernels launched by this library since load, and optional per-kernel
 * device timing of the binning path (CUDA events recorded on the launching stream around each launch;
 * kind 0 = scatter, 1 = finalize, 2 = other).  ep_profile_read synchronises on the recorded events,
 * sums them and clears the log.  Profiling state is process-global and off by default. */
typedef struct ep_profile_stats {
    double ms[3];
    int launches[3];
} ep_profile_stats;
EP_API unsigned long long ep_launch_count(void);
EP_API int ep_profile_enable(int on);
EP_API int ep_profile_read(ep_profile_stats* out);

/* ---------------------------------------------------------------------------------------------
 * Stage 1 — raw events -> dense tensors
 * ------------------------------------------------------------------------------------------- */

/* Ragged structure-of-arrays batch of event streams.  Sample b owns events
 * [offsets[b], offsets[b+1]).  The canonical (fast, vectorised) layout is x,y = EP_U16,
 * t = EP_I64 or EP_F64, p = EP_U8 with 16-byte aligned base pointers; any other combination of
 * the tags above takes a scalar-load kernel.  Timestamp value used by the arithmetic is
 * (double)t / t_div when t_div != 1 (e.g. int64 microseconds with t_div = 1e6 reproduce the
 * reference's `t / 1e6` seconds, dataset/pretrain/pr_n_imagenet_dataset.py:52-54), else (double)t.
 * Polarity: 1 = positive; 0 or -1 = negative (events_to_voxel_grid.py:35-36, events_to_image.py:13-16).
 */
typedef struct ep_events_soa {
    const void* x;
    const void* y;
    const void* t;
    const void* p;
    int xy_dtype;                /* EP_U16 | EP_I16 | EP_I32 | EP_F32 | EP_F64 (fractional coords truncate toward 0); EP_U32 = packed layout */
    int t_dtype;                 /* EP_I64 | EP_F64 | EP_F32; EP_U32 = compact layout; EP_U8 / 0 = packed layouts */
    int p_dtype;                 /* EP_U8 | EP_I8 | EP_F32 | EP_F64 */
    int batch;                   /* B */
    double t_div;
    const int64_t* offsets;      /* device, B+1 entries, offsets[0] may be > 0 */
    const int64_t* offsets_host; /* host copy of the same B+1 entries (required: sizes the launches) */
    const int64_t* t_base;       /* device, B entries, or NULL.  Transport layouts (ep_bin_events; the 4 B form also ep_evrep), bit-identical
                                    in results to the int64 canonical layout; timestamp value = (t_base[b] + ticks) / t_div:
                                    - compact, 8 B/event: t_dtype = EP_U32 holds ticks relative to t_base[b] in bits
                                      0..30 and the polarity in bit 31, p = NULL;
                                    - packed, 5 or 4 B/event: xy_dtype = EP_U32, x = one word per event
                                      x | y << 11 | polarity << 22 | tick bits << 23, y = NULL; p_dtype = EP_U32, p = one
                                      tick offset per BLOCK events of the arrays.  `ticks` of event i of sample b count
                                      from t_base[b] when i lies in the block of the sample's first event, else from
                                      t_base[b] + p[i / BLOCK].  5 B form: t_dtype = EP_U8, t = (ticks & 0xff) per event,
                                      word bits 23..31 = ticks >> 8 (17-bit ticks), BLOCK = 1024.  4 B form: t_dtype = 0,
                                      t = NULL, word bits 23..31 = ticks (9 bits), BLOCK = 256.  x, y < 2048; x and t
                                      16-byte aligned. */
} ep_events_soa;

/* Array-of-structures single sample: the reference's own (N,4) x,y,t,p array
 * (events_to_voxel_grid.py:8, events_to_image.py:10), fp64 or fp32, resident on the device. */
typedef struct ep_events_aos {
    const void* events;          /* (n,4) row-major */
    int dtype;                   /* EP_F64 | EP_F32 (fp32 input => the reference's fp32 time arithmetic) */
    int64_t n;
} ep_events_aos;

typedef struct ep_bin_params {
    int height, width;           /* output grid (size[0], size[1] of the reference calls) */
    int num_bins;                /* voxel bins (args.num_bins); 0 = no voxel grid */
    int count_channels;          /* 0 = none, 2 = events_to_image_ecdp [pos,neg], 3 = events_to_image_mem [pos,0,neg] */
    double scale_x, scale_y;     /* events_reshape fused (dataset/augmentation/events_augment.py:22-26):
                                    x*scale_x, y*scale_y in fp64, then truncation; 1.0 = none */
    int time_f32;                /* 1 = do the time arithmetic in fp32 (what torch does for float32 event arrays) */
    int flags;                   /* 0 = choose; EP_BIN_FORCE_GLOBAL / EP_BIN_FORCE_TILED / EP_BIN_FORCE_PLANE pin the kernel family */
} ep_bin_params;
#define EP_BIN_FORCE_GLOBAL 1    /* packed-u64 global RED + finalize (any layout, any size) */
#define EP_BIN_FORCE_TILED 4     /* route + two-plane shared-memory sweep, no global accumulators (4 B packed layout, voxel grid
                                    and sum plane only; EP_EUNSUPPORTED otherwise).  The default for what it takes. */
#define EP_BIN_FORCE_PLANE 8     /* whole-plane kernels: grids of at most 54272 cells (224 x 224, 240 x 180), one output plane per
                                    CTA in shared memory, no route pass (4 B packed layout; EP_EUNSUPPORTED otherwise).  The
                                    default for what it takes; unsorted samples are redone by the tiled kernels in the same call. */

/* Scratch for ep_bin_events*: per-sample accumulator slots.  Returns the recommended size (enough
 * slots to keep one group of samples L2-resident); any size >= the minimum (one slot + per-sample
 * metadata), returned through *min_bytes when non-NULL, works. */
EP_API size_t ep_bin_events_workspace_bytes(const ep_bin_params* prm, int batch, size_t* min_bytes);
/* Same, knowing the batch (reads offsets_host only): also covers the routed-record buffer of the tiled
 * path, which is taken when the layout is the 4 B packed one and the workspace is large enough. */
EP_API size_t ep_bin_events_workspace_bytes_for(const ep_events_soa* ev, const ep_bin_params* prm);

/* Host-only query (no GPU): does one axis of the fused events_reshape (dataset/augmentation/events_augment.py:22-26:
 * x * scale in fp64, then truncation) have a proven multiply-high form, trunc(fl(i * scale)) == (i * multiplier) >> 32 for every
 * coordinate i < 2048 of the packed layouts?  Returns 1 and the multiplier when it does — the whole-plane kernels then
 * compute that axis with one multiply instead of a shared-memory table — 0 when some coordinate differs (e.g. 0.35 * 180 =
 * 62.99999999999999 in fp64) or scale >= 1; the kernels decide per call with this very function. */
EP_API int ep_reshape_axis_multiplier_host(double scale, uint32_t* multiplier);

/* Replaces, for a whole ragged batch in one call:
 *   events_to_voxel_grid(args, events, size)   dataset/dataset_utils/events_to_voxel_grid.py:4-61
 *   events_to_image_ecdp / events_to_image_mem dataset/dataset_utils/events_to_image.py:6-62
 *   events_reshape (fused scale)               dataset/augmentation/events_augment.py:22-26
 *   voxel.sum(dim=0)[None] (event-side diff proxy) dataset/pretrain/pr_ef_imagenet_dataset.py:192-193
 * out_voxel (B,num_bins,H,W) f32; out_voxel_sum (B,1,H,W) f32 or NULL; out_count (B,count_channels,H,W)
 * f32 (exact integers) or NULL.  Voxel weights are accumulated in 64-bit fixed point (Q24), so the
 * result is independent of event order and bit-reproducible run to run. */
EP_API int ep_bin_events(void* stream, const ep_events_soa* ev, const ep_bin_params* prm,
                  float* out_voxel, float* out_voxel_sum, float* out_count,
                  void* workspace, size_t workspace_bytes, unsigned int* bad_count);

/* ep_bin_events plus the batch statistics that feed the path's one collective (north_star: "small all-reduce of
 * dataset-level normalisation statistics"; the reference's analogue is misc.all_reduce_mean, utils/misc.py:406-414):
 * out_stats (num_bins + 1, 4) f64 = per voxel channel, then for the sum plane when out_voxel_sum is given,
 * (element count, sum, sum of squares, max) over the batch.  On the tiled path they are by-products of the pass that
 * writes the planes (block-reduced in a fixed order: reproducible); other paths add one native pass.  NULL = none. */
EP_API int ep_bin_events_stats(void* stream, const ep_events_soa* ev, const ep_bin_params* prm,
                        float* out_voxel, float* out_voxel_sum, float* out_count,
                        void* workspace, size_t workspace_bytes, unsigned int* bad_count, double* out_stats);

/* The same statistics for any (B,C,H,W) f32 tensor: out (C,4) f64, one pass, fixed reduction order. */
EP_API size_t ep_plane_statistics_workspace_bytes(int channels);
EP_API int ep_plane_statistics(void* stream, const float* x, int batch, int channels, int height, int width, double* out,
                        void* workspace, size_t workspace_bytes);

/* Same, for one (N,4) array exactly as the reference functions receive it. */
EP_API int ep_bin_events_aos(void* stream, const ep_events_aos* ev, const ep_bin_params* prm,
                      float* out_voxel, float* out_voxel_sum, float* out_count,
                      void* workspace, size_t workspace_bytes, unsigned int* bad_count);

/* Per-sample normalisers applied by the datasets right after binning, in place on (B,C,H,W) f32:
 *   EP_NORM_COUNT: x/(amax_hw(x)+1) then (x-0.5)*2    dataset/pretrain/pr_n_imagenet_dataset.py:142-143
 *   EP_NORM_MEM:   channels 0,2 *= 1/max(ch0,ch2)     dataset/finetune_cls/ft_n_caltech101_dataset.py:96-98
 *   EP_NORM_MEM_GUARD: same, factor 1/0.001 when max == 0   dataset/finetune_flow/ft_mvsec_dataset.py:244-249
 * workspace: ep_normalise_workspace_bytes(). */
#define EP_NORM_COUNT 1
#define EP_NORM_MEM 2
#define EP_NORM_MEM_GUARD 3
EP_API size_t ep_normalise_workspace_bytes(int batch, int channels);
EP_API int ep_normalise(void* stream, float* img, int batch, int channels, int height, int width, int mode,
                 void* workspace, size_t workspace_bytes);

/* remove_hot_pixel_mem(hist, num_stds)   dataset/dataset_utils/events_to_image.py:65-75
 * In place on (B,3,H,W) f32: threshold = mean + num_stds*std (unbiased) over channels 0,2 jointly;
 * pixels where either exceeds it get both zeroed.  Optional `scale` multiplies the frame first
 * (the callers' `/ 255`, ft_n_caltech101_dataset.py:75, passed as 255 -> divide; 1 = none).  workspace: ep_mem_hotpixel_workspace_bytes(). */
EP_API size_t ep_mem_hotpixel_workspace_bytes(int batch);
EP_API int ep_mem_hotpixel(void* stream, float* hist, int batch, int height, int width, float divide_by,
                    float num_stds, void* workspace, size_t workspace_bytes);

/* events_to_EvRep(xs, ys, ts, ps, resolution=(W,H))   dataset/dataset_utils/events_to_image.py:77-125
 * out (B,3,H,W) f64 = [E_C, E_I, E_T] like the reference (callers cast to f32).  Bit-exact: events are
 * counting-sorted by pixel in the reference's lexsort order and each pixel's deltas are accumulated
 * sequentially with numpy's add.at rounding. */
EP_API size_t ep_evrep_workspace_bytes(int batch, int height, int width, int64_t n_total);
/* Same, knowing the batch: the 4 B packed transport layout takes the routed path (events bucketed into column tiles, per-tile
 * counting sort and replay in shared memory; timestamp value = (t_base[b] + ticks) / t_div), whose workspace also holds the
 * routed records.  bad_count then also reports: bit 31 = more than 65535 events on one pixel of one sample, bit 30 = a stamp
 * before, or 2^32 ticks or more after, the sample's first row, bit 29 = more than 47000 events on one pixel or more than
 * 752k events on one column tile (the per-tile shared-memory windows). */
EP_API size_t ep_evrep_workspace_bytes_for(const ep_events_soa* ev, int height, int width);
EP_API int ep_evrep(void* stream, const ep_events_soa* ev, int height, int width, double* out,
             void* workspace, size_t workspace_bytes, unsigned int* bad_count);

/* Time surface (PARITY UNPINNED: the reference has no such routine; self-oracle in oracle/stage3_np.py).
 * out (B,2,H,W) f32: channel 0 positive, 1 negative; exp(-(t_ref - t_last)/tau) of the latest event per pixel, 0 where
 * none.  t_ref: device array of B doubles, or NULL = each sample's last-row timestamp. */
EP_API size_t ep_time_surface_workspace_bytes(int batch, int height, int width);
EP_API int ep_time_surface(void* stream, const ep_events_soa* ev, int height, int width, double tau, const double* t_ref,
                    float* out, void* workspace, size_t workspace_bytes, unsigned int* bad_count);

/* View augmentation, fused: crop -> resize -> horizontal flip -> time flip.
 *   evg_augment / frame_augment   dataset/augmentation/view_augment.py:9-89
 * The stochastic choices stay on the host (the reference draws them from the global numpy RNG, re-seeded with the
 * same seed for a voxel grid and its sub_frame so that crops coincide, pr_ef_imagenet_dataset.py:187-206) and are
 * passed per sample.  in (B,C,H,W) f32 -> out (B,C,out_h,out_w) f32.  Resize = F.interpolate(align_corners=None). */
typedef struct ep_view_params {
    int crop_x, crop_y, crop_w, crop_h;   /* view_crop box (:26-28); full frame when no crop was drawn */
    int hflip;                            /* view_horizontal_flip (:41-47) */
    int time_flip;                        /* evg_time_flip: reverse the channel (bin) axis (:49-58) */
    int negate;                           /* evg_time_flip's sign change for 5/6 bins, or frame_time_flip (:60-63) */
    int reserved;
} ep_view_params;
#define EP_RESIZE_NEAREST 0
#define EP_RESIZE_BILINEAR 1
#define EP_RESIZE_BICUBIC 2
EP_API int ep_view_augment(void* stream, const float* in, int batch, int channels, int height, int width,
                    const ep_view_params* params /* device, B entries */, int out_h, int out_w, int mode, float* out);

/* ---------------------------------------------------------------------------------------------
 * Stage 2 — difference-map target
 * ------------------------------------------------------------------------------------------- */

/* Frame-side target T = g(f1) - g(f0), g = identity (mode 0) or log(.+eps) (mode 1), negated when
 * negate[b] != 0 (time reversal flips the sign: dataset/augmentation/view_augment.py:60-63,86-87).
 * The reference loads pre-computed sub_frame files (dataset/pretrain/pr_ef_imagenet_dataset.py:167-173)
 * and ships no generator, so this formula is the build's own (DESIGN.md "parity unpinned" list).
 * f0,f1,out: (B,1,H,W) f32 — n = B*H*W elements, per_sample = H*W; negate: B bytes or NULL. */
EP_API int ep_diffmap_frames(void* stream, const float* f0, const float* f1, float* out, int64_t n,
                      int64_t per_sample, int mode, float eps, const uint8_t* negate);

/* Target half of PrHubModel.reconstruct_loss   model/pretrain/pr_hub_model.py:125-131
 * + frame2emb                                   utils/reshape.py:15-22
 * frame (B,C,H,W) f32 -> out (B,(H/p)*(W/p), p*p*C) f32 in (ph,pw,c) element order; when norm_pix != 0
 * each patch is normalised (x-mean)/sqrt(var_unbiased + eps). */
EP_API int ep_patchify_normpix(void* stream, const float* frame, int batch, int channels, int height,
                        int width, int patch, int norm_pix, float eps, float* out);

/* Loss tail   model/pretrain/pr_hub_model.py:133-139
 * fused with the target build: per-patch mean((pred - target)^2) -> patch_loss (B,L) f32.  The final
 * (mask*loss).sum()/mask.sum() is two tiny reductions left to the caller. */
EP_API int ep_target_patch_loss(void* stream, const float* frame, const float* pred, int batch, int channels,
                         int height, int width, int patch, int norm_pix, float eps, float* patch_loss);
/* Same, computed only where the loss keeps it (pr_hub_model.py:139: `(mask * loss).sum() / mask.sum()`): mask (B,L) f32,
 * 1 = removed patch = counted.  Patches with mask == 0 are neither read nor normalised; their patch_loss entry is 0, so the
 * caller's (mask * patch_loss).sum() / mask.sum() is unchanged while 1 - mask_ratio of the target work disappears. */
EP_API int ep_target_patch_loss_masked(void* stream, const float* frame, const float* pred, const float* mask, int batch,
                                int channels, int height, int width, int patch, int norm_pix, float eps, float* patch_loss);

/* ---------------------------------------------------------------------------------------------
 * Stage 3 — masking, patchify, visible-token gather
 * ------------------------------------------------------------------------------------------- */

/* random_masking(x) with the noise supplied by the caller
 *   model/backbone/vit.py:66-105 (= convvit.py:85-124, swin.py:113-152)
 * noise (B,L) f32 -> ids_keep (B,len_keep) i64, mask (B,L) f32 (1 = removed), ids_restore (B,L) i64.
 * Ranks are those of a stable ascending argsort.  L <= 1024. */
EP_API int ep_mask_from_noise(void* stream, const float* noise, int batch, int L, int len_keep,
                       int64_t* ids_keep, float* mask, int64_t* ids_restore);

/* density noise   model/backbone/vit.py:80-83, swin.py:127-130
 * x (B,C,H,W) f32 -> out (B,(H/p)*(W/p)) f32 = sign * AvgPool2d(p,p)(abs(sum_c x)); sign = +1 "density",
 * -1 "anti-density".  Accumulation order matches the reference's CPU ops (bit-exact). */
EP_API int ep_patch_density(void* stream, const float* x, int batch, int channels, int height, int width,
                     int patch, float sign, float* out);

/* visible-token gather   model/backbone/vit.py:113-115, convvit.py:137-139,149-151,157-159
 * out[b,k,:] = tokens[b, ids_keep[b,k], :] + pos_embed[ids_keep[b,k], :]   (pos_embed may be NULL).
 * tokens (B,L,D) f32, pos_embed (L,D) f32, ids_keep (B,K) i64, out (B,K,D) f32.  D % 4 == 0. */
EP_API int ep_gather_tokens(void* stream, const float* tokens, const float* pos_embed, const int64_t* ids_keep,
                     int batch, int L, int K, int D, float* out);

/* patchify + gather on the raw input, so PatchEmbed runs on visible patches only (the per-patch ops
 * of vit.py:110-115 commute with the gather).  x (B,C,H,W) f32 -> out (B,K,C*p*p) f32.
 * order 0: (c,ph,pw) = Conv2d(k=s=p) weight order (model/sub_module/vit_block.py:44-68);
 * order 1: (ph,pw,c) = frame2emb order (utils/reshape.py:19).  ids_keep NULL = all patches in order. */
#define EP_ORDER_CPQ 0
#define EP_ORDER_PQC 1
EP_API int ep_patchify_gather(void* stream, const float* x, const int64_t* ids_keep, int batch, int channels,
                       int height, int width, int patch, int K, int order, float* out);

/* ConvViT block masks   model/backbone/convvit.py:129-130,142-143 (+ the `1 - mask` of :133,146)
 * mask (B,grid*grid) f32 -> out (B,1,grid*rep,grid*rep) f32, each cell repeated rep x rep;
 * invert != 0 writes 1 - mask (what ConvBlock receives, model/sub_module/conv_block.py:41-44). */
EP_API int ep_block_mask_expand(void* stream, const float* mask, int batch, int grid, int rep, int invert,
                         float* out);

/* Swin apply_mask   model/backbone/swin.py:154-179
 * Uses mask row 0 only (batch-shared, :158).  mask_row (Mh*Mw) f32 (nonzero = removed);
 * x (B,N,C) f32 with N = (Mh*r)*(Mw*r).  Writes vis_mask (N) u8, coords (n_vis,2) i64 (h,w),
 * x_vis (B,n_vis,C) f32 in row-major token order, and *n_vis_out (device int).  n_vis is data
 * dependent; the caller sizes x_vis/coords for n_vis_max = number of zeros in the mask row times r*r
 * (known on the host from len_keep) and passes it.  C % 4 == 0. */
EP_API int ep_swin_apply_mask(void* stream, const float* x, const float* mask_row, int batch, int Mh, int Mw,
                       int rep, int C, int n_vis_max, float* x_vis, int64_t* coords, uint8_t* vis_mask,
                       int* n_vis_out);

/* Swin stage outputs, the consumers of the visibility mask   model/backbone/swin.py:212-238
 * ep_swin_scatter_dense: x (B,n_vis,C) f32 tokens of the visible cells, coords (n_vis,2) i64 (h,w) (batch-shared, from
 *   ep_swin_apply_mask or a PatchMerging stage) -> out (B,C,G,G) f32, zero where no token sits: the
 *   `_emb = zeros(B,G*G,C); _emb[:, h*G+w, :] = x; reshape; permute(0,3,1,2)` of :221-225 in one pass.
 * ep_gather_tokens_nchw: feat (B,D,L) f32 (a conv output (B,D,Gh,Gw) as it lies in memory) and the per-sample
 *   ids_keep (B,K) i64 -> out (B,K,D) = feat.flatten(2).permute(0,2,1) gathered along the tokens (:226-228). */
EP_API int ep_swin_scatter_dense(void* stream, const float* x, const int64_t* coords, int batch, int n_vis, int grid, int C,
                          float* out);
EP_API int ep_gather_tokens_nchw(void* stream, const float* feat, const int64_t* ids_keep, int batch, int L, int K, int D,
                          float* out);

/* ConvViT feature fusion   model/backbone/convvit.py:137-140, 151-154, 166-167
 * out[b,k,:] = (feat1[b,:,ids[b,k]] + feat2[b,:,ids[b,k]]) + emb3[b,k,:]: the two stage decoders' outputs (B,D,L) as the
 * convolutions leave them, gathered by ids_keep (B,K) and added to the transformer stage's tokens emb3 (B,K,D; NULL = omit)
 * in one launch — no flatten/permute copies, no separate gathers and adds.  Same association as the reference's sum. */
EP_API int ep_gather_sum_nchw(void* stream, const float* feat1, const float* feat2, const float* emb3, const int64_t* ids_keep,
                       int batch, int L, int K, int D, float* out);

/* decoder un-shuffle   model/pretrain/pr_rec_decoder.py:56-62
 * out[b,l,:] = (ids_restore[b,l] < K ? emb[b, ids_restore[b,l], :] : mask_token[:]) + pos_embed[l,:]. */
EP_API int ep_unshuffle_tokens(void* stream, const float* emb, const float* mask_token, const float* pos_embed,
                        const int64_t* ids_restore, int batch, int L, int K, int D, float* out);

/* Backward passes of the token permutations — the reference trains through these index ops (pr_trainer.py:26-36), so the
 * drop-ins carry a gradient (torch.autograd.Function wrappers in eventpretrain_b200/masking.py).  D % 4 == 0.
 * ep_scatter_add_tokens: out (B,L,D) = 0; out[b, ids[b,k], :] += grad[b,k,:]; ids (B,K) i64, or (K) shared by the batch when
 *   ids_shared != 0.  Backward of ep_gather_tokens (vit.py:113-115), of GroupingModule.group / merge (swin_block.py:446-464;
 *   padded slots repeat token 0, hence the accumulation) and of ep_swin_apply_mask (swin.py:154-179).
 * ep_sum_over_batch: out[n] = sum_b in[b,n], fixed order (pos_embed / mask_token gradients).  n % 4 == 0.
 * ep_unshuffle_tokens_bwd: backward of ep_unshuffle_tokens (pr_rec_decoder.py:56-62): grad_emb (B,K,D), grad_mask_token (D);
 *   scratch: B*D floats.  The pos_embed gradient is ep_sum_over_batch of grad.
 * ep_scatter_add_tokens_nchw: out (B,D,L) = 0; out[b,d,ids[b,k]] += grad[b,k,d]: backward of ep_gather_tokens_nchw
 *   (swin.py:226-228); the backward of ep_swin_scatter_dense is ep_gather_tokens_nchw itself with the cells as indices. */
EP_API int ep_scatter_add_tokens(void* stream, const float* grad, const int64_t* ids, int ids_shared, int batch, int L, int K,
                          int D, float* out);
EP_API int ep_sum_over_batch(void* stream, const float* in, int batch, int64_t n, float* out);
EP_API int ep_unshuffle_tokens_bwd(void* stream, const float* grad, const int64_t* ids_restore, int batch, int L, int K, int D,
                            float* grad_emb, float* grad_mask_token, float* scratch);
EP_API int ep_scatter_add_tokens_nchw(void* stream, const float* grad, const int64_t* ids_keep, int batch, int L, int K, int D,
                               float* out);

/* Swin sparse-token grouping, HOST function, no device work:
 *   knapsack / group_windows   model/sub_module/swin_block.py:280-352 (called from GroupingModule._prepare_grouping :387-431)
 * Packs windows holding num_ele_win[i] visible tokens (1 <= . <= group_size) into groups of at most group_size tokens,
 * greedily, one 0/1 knapsack per group, with the reference's table, back-tracking and tie behaviour: identical groups.
 * Outputs: num_ele_group[g] tokens of group g; the windows of group g are grouped_idx[group_first[g] .. group_first[g+1])
 * in increasing order; *n_groups groups (<= n_win).  Arrays sized n_win (group_first: n_win + 1). */
EP_API int ep_swin_group_windows_host(int group_size, const int* num_ele_win, int n_win, int* num_ele_group, int* group_first,
                                      int* grouped_idx, int* n_groups);

/* Attention tables of a window grouping, HOST function   model/sub_module/swin_block.py:368-385 (+ :424-431)
 * group_id (n_groups, group_size) i64 window id per slot (-1 = padding), coords (n_groups, group_size, 2) i64 (h,w) ->
 *   attn_mask (n_groups, gs, gs) f32: 0 where two slots share a window id (compared as float32, like the reference) and are
 *     not both padding, else -100;
 *   rel_pos_idx (n_groups, gs, gs) i64 = (dh + window-1) * (2*window-1) + (dw + window-1), zeroed where attn_mask != 0 when
 *     mask_rel != 0 (the grouping mode's `rel_pos_idx * rel_pos_mask`). */
EP_API int ep_swin_group_tables_host(const int64_t* group_id, const int64_t* coords, int n_groups, int group_size, int window,
                                     int mask_rel, float* attn_mask, int64_t* rel_pos_idx);

/* Ragged collate, HOST functions, multi-threaded (threads <= 0: all hardware threads), no device work.
 * The reference's DataLoader workers hand over per-sample (N,4) arrays with columns x, y, t, p (float64, or float32 for
 * DDD17 / DVS128-Gesture: pr_n_imagenet_dataset.py:52-54, ft_ddd17_dataset.py:96-97); the collate step replaces the
 * per-sample CPU binning there (e.g. pr_n_imagenet_dataset.py:85-87).
 *
 * ep_collate_aos_host: B sample pointers + counts -> canonical SoA batch (x, y uint16; t int64 ticks = rint(t * t_scale);
 *   p uint8) and offsets[B + 1].  dtype = EP_F64 or EP_F32.  EP_EUNSUPPORTED when a coordinate is not an integer in
 *   [0, 65535] or a polarity is not 0 / 1 (such batches keep the generic layout); the outputs are then unspecified.
 * ep_pack_transport_host: canonical SoA batch (offsets[0] == 0) -> packed transport layout of ep_events_soa.t_base:
 *   nbytes = 5: w[n] + tick_low[n], blk_base[ceil(n / 1024)], t_base[B];  nbytes = 4: w[n], blk_base[ceil(n / 256)], t_base[B]
 *   (tick_low may be NULL);  nbytes = 8: compact layout, w[i] = ticks since t_base[b] | polarity << 31 at the events' own
 *   array positions (x, y, tick_low, blk_base unused and may be NULL; offsets[0] >= 0).  EP_EUNSUPPORTED when the batch does not fit the layout (x or y >= 2048, polarity > 1, a
 *   block spanning more ticks than its field, stamps far out of order): the caller takes the next wider layout. */
EP_API int ep_collate_aos_host(const void* const* samples, const int64_t* counts, int batch, int dtype, double t_scale,
                               uint16_t* x, uint16_t* y, int64_t* t, uint8_t* p, int64_t* offsets, int threads);
EP_API int ep_pack_transport_host(const uint16_t* x, const uint16_t* y, const int64_t* t, const uint8_t* p,
                                  const int64_t* offsets, int batch, int nbytes, uint32_t* w, uint8_t* tick_low,
                                  uint32_t* blk_base, int64_t* t_base, int threads);

/* The two steps above in one pass for the 4 B packed layout: per-sample (N,4) x,y,t,p rows -> packed words, per-block tick
 * offsets, per-sample bases and offsets, reading the 32 B/event rows once and writing 4 B/event (the two-step form moves 62 B
 * of host memory per event, this one 36).  The sample's base is taken from its first row, so it covers time-sorted samples
 * whose coordinates are < 2048, polarity in {0,1} and whose 256-event blocks span < 512 ticks: same arrays, bit for bit, as
 * ep_collate_aos_host + ep_pack_transport_host(nbytes = 4).  EP_EUNSUPPORTED for anything else (take the two-step form). */
EP_API int ep_collate_transport4_host(const void* const* samples, const int64_t* counts, int batch, int dtype, double t_scale,
                               uint32_t* w, uint32_t* blk_base, int64_t* t_base, int64_t* offsets, int threads);

#ifdef __cplusplus
}
#endif
#endif /* EVENTPRETRAIN_B200_H */

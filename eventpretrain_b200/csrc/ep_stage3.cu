// Stage 3 — masking, patchify, visible-token gather; and the patchified diff-map target (sm_100a).
//
// Every kernel here is a permutation / small reduction over fp32 tensors: HBM-bound, no dense
// contraction, so no tensor cores.  Rows are moved with 16-byte vector accesses by one warp per
// row; outputs that are consumed once downstream are written with streaming stores.
#include "ep_common.cuh"

namespace ep {
namespace {

// ---------------------------------------------------------------------------------------------------
// random_masking core   model/backbone/vit.py:91-105
// One CTA per sample.  rank_i = #{ j : n_j < n_i or (n_j == n_i and j < i) }  (stable ascending
// argsort); ids_restore[i] = rank_i; ids_keep[rank_i] = i when rank_i < len_keep; mask = rank >= keep.
// The all-pairs count reads noise from shared memory as warp-wide broadcasts: L^2 = 38k compares
// for L = 196, far cheaper than two device-wide sorts plus four gather launches.
// ---------------------------------------------------------------------------------------------------
__global__ void k_mask_from_noise(const float* __restrict__ noise, int L, int len_keep,
                                  int64_t* __restrict__ ids_keep, float* __restrict__ mask,
                                  int64_t* __restrict__ ids_restore) {
    extern __shared__ float s_noise[];
    const int b = blockIdx.x;
    const float* n = noise + (int64_t)b * L;
    for (int i = threadIdx.x; i < L; i += blockDim.x) s_noise[i] = n[i];
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float v = s_noise[i];
        int rank = 0;
        for (int j = 0; j < L; ++j) {
            const float w = s_noise[j];
            rank += (w < v) || (w == v && j < i);
        }
        ids_restore[(int64_t)b * L + i] = rank;
        mask[(int64_t)b * L + i] = rank >= len_keep ? 1.0f : 0.0f;
        if (rank < len_keep) ids_keep[(int64_t)b * len_keep + rank] = i;
    }
}

// ---------------------------------------------------------------------------------------------------
// density noise   model/backbone/vit.py:80-83
// One thread per patch.  fp32 accumulation in the reference's order: channels in order, then the
// p x p window row-major into a zero-initialised accumulator, then one division by p*p.
// ---------------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(128) k_patch_density(const float* __restrict__ x, int B, int C, int H, int W,
                                                       int p, float sign, float* __restrict__ out) {
    const int gh = H / p, gw = W / p, L = gh * gw;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * L) return;
    const int b = (int)(idx / L), l = (int)(idx % L);
    const int py = l / gw, px = l % gw;
    const int64_t HW = (int64_t)H * W;
    const float* base = x + (int64_t)b * C * HW + (int64_t)(py * p) * W + px * p;
    float acc = 0.0f;
    for (int r = 0; r < p; ++r) {
        const float* row = base + (int64_t)r * W;
        if (VEC4) {
            for (int q = 0; q < p; q += 4) {
                float4 s = __ldg(reinterpret_cast<const float4*>(row + q));
                for (int c = 1; c < C; ++c) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(row + c * HW + q));
                    s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y);
                    s.z = __fadd_rn(s.z, v.z); s.w = __fadd_rn(s.w, v.w);
                }
                acc = __fadd_rn(acc, fabsf(s.x)); acc = __fadd_rn(acc, fabsf(s.y));
                acc = __fadd_rn(acc, fabsf(s.z)); acc = __fadd_rn(acc, fabsf(s.w));
            }
        } else {
            for (int q = 0; q < p; ++q) {
                float s = __ldg(row + q);
                for (int c = 1; c < C; ++c) s = __fadd_rn(s, __ldg(row + c * HW + q));
                acc = __fadd_rn(acc, fabsf(s));
            }
        }
    }
    out[idx] = sign * __fdiv_rn(acc, (float)(p * p));
}

// ---------------------------------------------------------------------------------------------------
// row movers: one warp per destination row of D floats (D % 4 == 0), 16-byte accesses
// ---------------------------------------------------------------------------------------------------
// visible-token gather   model/backbone/vit.py:113-115
__global__ void __launch_bounds__(256) k_gather_tokens(const float* __restrict__ tokens,
                                                       const float* __restrict__ pos,
                                                       const int64_t* __restrict__ ids, int64_t rows, int L,
                                                       int K, int D, float* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = row / K;
    int64_t src = ids[row];
    if (src < 0 || src >= L) src = 0;   // torch.gather raises; ids come from ep_mask_from_noise
    const float4* s = reinterpret_cast<const float4*>(tokens + (b * L + src) * D);
    const float4* pp = pos ? reinterpret_cast<const float4*>(pos + src * D) : nullptr;
    float4* o = reinterpret_cast<float4*>(out + row * D);
    for (int i = lane; i < D / 4; i += 32) {
        float4 v = ld_stream(s + i);
        if (pp) {
            const float4 q = __ldg(pp + i);
            v.x = __fadd_rn(v.x, q.x); v.y = __fadd_rn(v.y, q.y); v.z = __fadd_rn(v.z, q.z); v.w = __fadd_rn(v.w, q.w);
        }
        o[i] = v;
    }
}

// decoder un-shuffle   model/pretrain/pr_rec_decoder.py:56-62
__global__ void __launch_bounds__(256) k_unshuffle_tokens(const float* __restrict__ emb,
                                                          const float* __restrict__ mask_token,
                                                          const float* __restrict__ pos,
                                                          const int64_t* __restrict__ ids_restore, int64_t rows,
                                                          int L, int K, int D, float* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = row / L, l = row % L;
    const int64_t src = ids_restore[row];
    const float4* s = (src >= 0 && src < K) ? reinterpret_cast<const float4*>(emb + (b * K + src) * D)
                                            : reinterpret_cast<const float4*>(mask_token);
    const float4* pp = pos ? reinterpret_cast<const float4*>(pos + l * D) : nullptr;
    float4* o = reinterpret_cast<float4*>(out + row * D);
    for (int i = lane; i < D / 4; i += 32) {
        float4 v = __ldg(s + i);
        if (pp) {
            const float4 q = __ldg(pp + i);
            v.x = __fadd_rn(v.x, q.x); v.y = __fadd_rn(v.y, q.y); v.z = __fadd_rn(v.z, q.z); v.w = __fadd_rn(v.w, q.w);
        }
        o[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// patchify (+ gather)   utils/reshape.py:15-22 / Conv2d(k=s=p) operand order
// One CTA (128 threads) per destination patch; threads walk the destination row so stores are
// fully coalesced; sources are p-float row segments of one 16x16xC box (L1-resident).
// ---------------------------------------------------------------------------------------------------
template <int ORDER, bool VEC4>
__global__ void __launch_bounds__(128) k_patchify_gather(const float* __restrict__ x,
                                                         const int64_t* __restrict__ ids, int C, int H, int W,
                                                         int p, int K, float* __restrict__ out) {
    const int gw = W / p, L = (H / p) * gw;
    const int64_t row = blockIdx.x;            // b * K + k
    const int64_t b = row / K;
    int64_t l = ids ? ids[row] : row % K;
    if (l < 0 || l >= L) l = 0;
    const int py = (int)(l / gw), px = (int)(l % gw);
    const int64_t HW = (int64_t)H * W;
    const float* base = x + b * C * HW + (int64_t)(py * p) * W + px * p;
    const int n = C * p * p;
    float* o = out + row * n;
    if (ORDER == EP_ORDER_CPQ) {
        if (VEC4) {
            for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
                const int e = i * 4, c = e / (p * p), r = (e / p) % p, q = e % p;
                st_stream(reinterpret_cast<float4*>(o) + i,
                          __ldg(reinterpret_cast<const float4*>(base + c * HW + (int64_t)r * W + q)));
            }
        } else {
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                const int c = e / (p * p), r = (e / p) % p, q = e % p;
                st_stream(o + e, __ldg(base + c * HW + (int64_t)r * W + q));
            }
        }
    } else {   // (ph, pw, c)
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int c = e % C, q = (e / C) % p, r = e / (C * p);
            st_stream(o + e, __ldg(base + c * HW + (int64_t)r * W + q));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// target patchify + norm_pix (+ fused per-patch MSE)   model/pretrain/pr_hub_model.py:125-139
// One warp per patch: the patch is staged in this warp's slice of shared memory in (ph,pw,c)
// order, mean and unbiased variance are two shuffle reductions, then either the normalised patch is
// written (coalesced) or it is compared with `pred` and only the per-patch loss is written.
// ---------------------------------------------------------------------------------------------------
template <bool LOSS>
__global__ void __launch_bounds__(256) k_patch_target(const float* __restrict__ frame,
                                                      const float* __restrict__ pred, int64_t patches, int C,
                                                      int H, int W, int p, int norm_pix, float eps,
                                                      const float* __restrict__ mask, float* __restrict__ out) {
    extern __shared__ float s_patch[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= patches) return;
    if (LOSS && mask && mask[row] == 0.0f) {            // a patch the loss discards (mask * loss, pr_hub_model.py:139): not even read
        if (lane == 0) out[row] = 0.0f;
        return;
    }
    const int gw = W / p, L = (H / p) * gw;
    const int64_t b = row / L;
    const int l = (int)(row % L);
    const int py = l / gw, px = l % gw;
    const int64_t HW = (int64_t)H * W;
    const float* base = frame + b * C * HW + (int64_t)(py * p) * W + px * p;
    const int n = C * p * p;
    float* sp = s_patch + (int64_t)warp * n;
    // the three reductions run in fp64 (a few dozen adds per lane): mean, variance and loss then carry one fp32 rounding each,
    // which keeps the per-patch loss within 1e-5 of the reference's whatever its summation order
    double sum = 0.0;
    for (int e = lane; e < n; e += 32) {
        const int c = e % C, q = (e / C) % p, r = e / (C * p);
        const float v = ld_stream(base + c * HW + (int64_t)r * W + q);
        sp[e] = v;
        sum += (double)v;
    }
    float mean = 0.f, sd = 1.f;
    if (norm_pix) {
        sum = warp_reduce(sum, [](double a, double c) { return a + c; });
        mean = (float)(sum / (double)n);
        double ss = 0.0;
        __syncwarp();
        for (int e = lane; e < n; e += 32) { const float d = sp[e] - mean; ss += (double)d * (double)d; }
        ss = warp_reduce(ss, [](double a, double c) { return a + c; });
        const float var = (float)(ss / (double)(n - 1));          // torch.var default: unbiased
        sd = sqrtf(var + eps);                          // (var + 1e-6) ** .5
    }
    __syncwarp();
    if (LOSS) {
        const float* pr = pred + row * n;
        double acc = 0.0;
        for (int e = lane; e < n; e += 32) {
            const float t = norm_pix ? (sp[e] - mean) / sd : sp[e];
            const float d = ld_stream(pr + e) - t;
            acc += (double)d * (double)d;
        }
        acc = warp_reduce(acc, [](double a, double c) { return a + c; });
        if (lane == 0) out[row] = (float)(acc / (double)n);
    } else {
        float* o = out + row * n;
        for (int e = lane; e < n; e += 32) st_stream(o + e, norm_pix ? (sp[e] - mean) / sd : sp[e]);
    }
}

// Single-channel frames (the reference's sub_frame target, pr_ef_imagenet_dataset.py:167-173) with p in {8,16,32}: (ph,pw,c)
// order is then plain row-major inside the patch, so a warp moves its patch as 16-byte vectors held in registers — no
// staging, no index divisions; mean / unbiased variance are the same two-pass reductions as above.
template <int P, bool LOSS>
__global__ void __launch_bounds__(256) k_patch_target_c1(const float* __restrict__ frame, const float* __restrict__ pred,
                                                         int64_t patches, int H, int W, int norm_pix, float eps,
                                                         const float* __restrict__ mask, float* __restrict__ out) {
    constexpr int P4 = P / 4, NV = P * P4, V = (NV + 31) / 32, n = P * P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= patches) return;
    if (LOSS && mask && mask[row] == 0.0f) {            // a patch the loss discards: not even read
        if (lane == 0) out[row] = 0.0f;
        return;
    }
    const int gw = W / P, L = (H / P) * gw;
    const int64_t b = row / L;
    const int l = (int)(row % L);
    const float* base = frame + b * (int64_t)H * W + (int64_t)((l / gw) * P) * W + (l % gw) * P;
    float4 v[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int f = lane + 32 * j;
        v[j] = (NV % 32 == 0 || f < NV) ? ld_stream(reinterpret_cast<const float4*>(base + (int64_t)(f / P4) * W) + (f % P4))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 pr[V];
    if (LOSS) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int f = lane + 32 * j;
            pr[j] = (NV % 32 == 0 || f < NV) ? ld_stream(reinterpret_cast<const float4*>(pred + row * n) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float mean = 0.f, sd = 1.f;
    if (norm_pix) {
        // fp64 reductions, as in k_patch_target
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < V; ++j) sum += ((double)v[j].x + (double)v[j].y) + ((double)v[j].z + (double)v[j].w);
        sum = warp_reduce(sum, [](double a, double c) { return a + c; });
        mean = (float)(sum / (double)n);
        double ss = 0.0;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (NV % 32 == 0 || lane + 32 * j < NV) {
                const double d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
                ss += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
        }
        ss = warp_reduce(ss, [](double a, double c) { return a + c; });
        sd = sqrtf((float)(ss / (double)(n - 1)) + eps);          // torch.var default: unbiased; (var + 1e-6) ** .5
#pragma unroll
        for (int j = 0; j < V; ++j)
            v[j] = make_float4((v[j].x - mean) / sd, (v[j].y - mean) / sd, (v[j].z - mean) / sd, (v[j].w - mean) / sd);
    }
    if (LOSS) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (NV % 32 == 0 || lane + 32 * j < NV) {
                const double d0 = pr[j].x - v[j].x, d1 = pr[j].y - v[j].y, d2 = pr[j].z - v[j].z, d3 = pr[j].w - v[j].w;
                acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
        }
        acc = warp_reduce(acc, [](double a, double c) { return a + c; });
        if (lane == 0) out[row] = (float)(acc / (double)n);
    } else {
        float4* o = reinterpret_cast<float4*>(out + row * n);
#pragma unroll
        for (int j = 0; j < V; ++j)
            if (NV % 32 == 0 || lane + 32 * j < NV) st_stream(o + lane + 32 * j, v[j]);
    }
}

// ---------------------------------------------------------------------------------------------------
// ConvViT block masks   model/backbone/convvit.py:129-130,142-143
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_block_mask_expand(const float* __restrict__ mask, int64_t total, int grid,
                                                           int rep, int invert, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int side = grid * rep;
    const int X = (int)(i % side), Y = (int)((i / side) % side);
    const int64_t b = i / ((int64_t)side * side);
    const float m = __ldg(mask + b * grid * grid + (Y / rep) * grid + X / rep);
    out[i] = invert ? 1.0f - m : m;
}

// ---------------------------------------------------------------------------------------------------
// Swin apply_mask   model/backbone/swin.py:154-179
// Kernel 1 (one CTA): expand mask row 0 to the token grid, block-wide exclusive scan of the
// visibility bits (warp-shuffle scans + one shared-memory pass), write vis_mask, coords, n_vis.
// Kernel 2: one warp per (b, visible token) row copy.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_swin_scan(const float* __restrict__ mask_row, int Mh, int Mw, int rep,
                                                    int n_vis_max, int64_t* __restrict__ coords,
                                                    uint8_t* __restrict__ vis_mask, int* __restrict__ n_vis_out) {
    __shared__ int s_warp[32];
    __shared__ int s_base, s_total;
    const int Wt = Mw * rep, N = Mh * rep * Wt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < N; start += blockDim.x) {
        const int i = start + threadIdx.x;
        int vis = 0, h = 0, w = 0;
        if (i < N) {
            h = i / Wt; w = i % Wt;
            vis = (__ldg(mask_row + (h / rep) * Mw + w / rep) == 0.0f);   // mask.bool(): nonzero = removed
            vis_mask[i] = (uint8_t)vis;
        }
        const int incl = warp_incl_scan(vis, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = s_warp[lane];               // blockDim.x == 1024: all 32 entries are live
            const int s = warp_incl_scan(v, lane);
            s_warp[lane] = s - v;                     // exclusive offset of each warp
            if (lane == 31) s_total = s;
        }
        __syncthreads();
        const int pos = s_base + s_warp[warp] + incl - vis;
        if (vis && pos < n_vis_max) { coords[2 * (int64_t)pos] = h; coords[2 * (int64_t)pos + 1] = w; }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_vis_out = s_base;
}

__global__ void __launch_bounds__(256) k_swin_gather(const float* __restrict__ x, const int64_t* __restrict__ coords,
                                                     const int* __restrict__ n_vis, int B, int N, int Wt, int C,
                                                     int n_vis_max, float* __restrict__ x_vis) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nv = min(*n_vis, n_vis_max);
    if (row >= (int64_t)B * nv) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = row / nv, j = row % nv;
    const int64_t src = coords[2 * j] * Wt + coords[2 * j + 1];
    const float4* s = reinterpret_cast<const float4*>(x + (b * N + src) * C);
    float4* o = reinterpret_cast<float4*>(x_vis + (b * nv + j) * C);
    for (int i = lane; i < C / 4; i += 32) o[i] = ld_stream(s + i);
}

// ---- Swin stage outputs back to dense grids, and the per-sample re-gather (model/backbone/swin.py:212-238) ----
// dense[b, c, cell] = x[b, n, c] for the visible token n that sits on cell = h * G + w, 0 elsewhere:
//   _emb = zeros(B, G*G, C); _emb[:, coords[0,:,0]*G + coords[0,:,1], :] = x; _emb.reshape(B,G,G,C).permute(0,3,1,2)   (swin.py:221-225)
// One CTA per (sample, 32-channel slab, 32-cell strip): the inverse map cell -> token is built per CTA from the coordinates
// (batch-shared, a few KB), tokens are read as rows, transposed through shared memory and written as channel rows.
__global__ void __launch_bounds__(256) k_swin_scatter_dense(const float* __restrict__ x, const int64_t* __restrict__ coords, int n_vis,
                                                            int G, int C, float* __restrict__ out) {
    __shared__ int s_tok[32];
    __shared__ float s_tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, cell0 = blockIdx.x * 32;
    const int cells = G * G;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < 32) s_tok[threadIdx.x] = -1;
    __syncthreads();
    for (int n = threadIdx.x; n < n_vis; n += 256) {
        const int cell = (int)(coords[2 * n] * G + coords[2 * n + 1]);
        if (cell >= cell0 && cell < cell0 + 32) s_tok[cell - cell0] = n;       // last writer wins like the indexed assignment; coords are unique
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {            // r = cell of the strip, tx = channel of the slab: coalesced token rows
        const int tok = s_tok[r];
        s_tile[r][tx] = (tok >= 0 && c0 + tx < C) ? x[((int64_t)b * n_vis + tok) * C + c0 + tx] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {            // r = channel of the slab, tx = cell: coalesced channel rows
        if (c0 + r < C && cell0 + tx < cells) out[((int64_t)b * C + c0 + r) * cells + cell0 + tx] = s_tile[tx][r];
    }
}

// out[b, k, d] = feat[b, d, ids[b, k]]: stage_output_decode(...).flatten(2).permute(0,2,1) gathered by ids_keep   (swin.py:226-228)
__global__ void __launch_bounds__(256) k_gather_tokens_nchw(const float* __restrict__ feat, const int64_t* __restrict__ ids, int L, int K,
                                                            int D, float* __restrict__ out) {
    const int b = blockIdx.y, k = blockIdx.x;
    int64_t src = ids[(int64_t)b * K + k];
    if (src < 0 || src >= L) src = 0;
    for (int d = threadIdx.x; d < D; d += 256) out[((int64_t)b * K + k) * D + d] = feat[((int64_t)b * D + d) * L + src];
}

// ConvViT feature fusion (convvit.py:137-140, 151-154, 166-167): the two stage decoders' outputs, still (B,D,14,14) as the
// convolutions leave them, gathered by ids_keep and added to the transformer stage's tokens in one pass:
//   out[b,k,:] = (feat1[b,:,ids[b,k]] + feat2[b,:,ids[b,k]]) + emb3[b,k,:]        (the reference's left-to-right sum)
// without the two flatten(2).permute(0,2,1) copies, the two gathers and the two adds.
__global__ void __launch_bounds__(256) k_gather_sum_nchw(const float* __restrict__ f1, const float* __restrict__ f2,
                                                         const float* __restrict__ e3, const int64_t* __restrict__ ids, int L, int K, int D,
                                                         float* __restrict__ out) {
    const int b = blockIdx.y, k = blockIdx.x;
    int64_t src = ids[(int64_t)b * K + k];
    if (src < 0 || src >= L) src = 0;
    for (int d = threadIdx.x; d < D; d += 256) {
        const int64_t fi = ((int64_t)b * D + d) * L + src, oi = ((int64_t)b * K + k) * D + d;
        float v = __fadd_rn(f1[fi], f2[fi]);
        if (e3) v = __fadd_rn(v, e3[oi]);
        out[oi] = v;
    }
}

}  // namespace
}  // namespace ep

extern "C" {

int ep_mask_from_noise(void* stream, const float* noise, int batch, int L, int len_keep, int64_t* ids_keep,
                       float* mask, int64_t* ids_restore) {
    if (!noise || !mask || !ids_restore || batch <= 0 || L <= 0 || len_keep < 0 || len_keep > L) return EP_EINVAL;
    if (len_keep > 0 && !ids_keep) return EP_EINVAL;
    if (L > 4096) return EP_EUNSUPPORTED;
    int threads = (L + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    ep::k_mask_from_noise<<<batch, threads, L * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        noise, L, len_keep, ids_keep, mask, ids_restore);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_patch_density(void* stream, const float* x, int batch, int channels, int height, int width, int patch,
                     float sign, float* out) {
    if (!x || !out || batch <= 0 || channels <= 0 || patch <= 0 || height < patch || width < patch) return EP_EINVAL;
    const int64_t total = (int64_t)batch * (height / patch) * (width / patch);
    const unsigned blocks = (unsigned)ep::ceil_div64(total, 128);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (patch % 4 == 0) && (width % 4 == 0) && ep::aligned16(x);
    if (vec) ep::k_patch_density<true><<<blocks, 128, 0, st>>>(x, batch, channels, height, width, patch, sign, out);
    else ep::k_patch_density<false><<<blocks, 128, 0, st>>>(x, batch, channels, height, width, patch, sign, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_gather_tokens(void* stream, const float* tokens, const float* pos_embed, const int64_t* ids_keep, int batch,
                     int L, int K, int D, float* out) {
    if (!tokens || !ids_keep || !out || batch <= 0 || L <= 0 || K <= 0 || D <= 0) return EP_EINVAL;
    if (D % 4) return EP_EUNSUPPORTED;
    if (!ep::aligned16(tokens) || !ep::aligned16(out) || (pos_embed && !ep::aligned16(pos_embed))) return EP_EALIGN;
    const int64_t rows = (int64_t)batch * K;
    ep::k_gather_tokens<<<(unsigned)ep::ceil_div64(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        tokens, pos_embed, ids_keep, rows, L, K, D, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_unshuffle_tokens(void* stream, const float* emb, const float* mask_token, const float* pos_embed,
                        const int64_t* ids_restore, int batch, int L, int K, int D, float* out) {
    if (!emb || !mask_token || !ids_restore || !out || batch <= 0 || L <= 0 || K < 0 || K > L || D <= 0) return EP_EINVAL;
    if (D % 4) return EP_EUNSUPPORTED;
    if (!ep::aligned16(emb) || !ep::aligned16(out) || !ep::aligned16(mask_token) ||
        (pos_embed && !ep::aligned16(pos_embed)))
        return EP_EALIGN;
    const int64_t rows = (int64_t)batch * L;
    ep::k_unshuffle_tokens<<<(unsigned)ep::ceil_div64(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        emb, mask_token, pos_embed, ids_restore, rows, L, K, D, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_patchify_gather(void* stream, const float* x, const int64_t* ids_keep, int batch, int channels, int height,
                       int width, int patch, int K, int order, float* out) {
    if (!x || !out || batch <= 0 || channels <= 0 || patch <= 0 || height < patch || width < patch || K <= 0)
        return EP_EINVAL;
    if (order != EP_ORDER_CPQ && order != EP_ORDER_PQC) return EP_EINVAL;
    if (height % patch || width % patch) return EP_EUNSUPPORTED;
    if (!ids_keep && K != (height / patch) * (width / patch)) return EP_EINVAL;
    const int64_t rows = (int64_t)batch * K;
    if (rows > 0x7fffffffLL) return EP_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (patch % 4 == 0) && (width % 4 == 0) && ep::aligned16(x) && ep::aligned16(out);
    if (order == EP_ORDER_CPQ && !ep::patch_tma_disabled()) {
        // TMA form (csrc/ep_patch_tma.cu): box load of the (p, p, C) patch, bulk store of the finished row
        const int rc = ep::patchify_gather_tma(st, x, ids_keep, batch, channels, height, width, patch, K, out);
        if (rc == EP_OK) { EP_LAUNCH_CHECK(); return EP_OK; }
    }
    if (order == EP_ORDER_CPQ) {
        if (vec) ep::k_patchify_gather<EP_ORDER_CPQ, true><<<(unsigned)rows, 128, 0, st>>>(x, ids_keep, channels, height, width, patch, K, out);
        else ep::k_patchify_gather<EP_ORDER_CPQ, false><<<(unsigned)rows, 128, 0, st>>>(x, ids_keep, channels, height, width, patch, K, out);
    } else {
        ep::k_patchify_gather<EP_ORDER_PQC, false><<<(unsigned)rows, 128, 0, st>>>(x, ids_keep, channels, height, width, patch, K, out);
    }
    EP_LAUNCH_CHECK();
    return EP_OK;
}

static int launch_patch_target(void* stream, bool loss, const float* frame, const float* pred, int batch,
                               int channels, int height, int width, int patch, int norm_pix, float eps, const float* mask, float* out) {
    if (!frame || !out || (loss && !pred) || batch <= 0 || channels <= 0 || patch <= 0) return EP_EINVAL;
    if (height < patch || width < patch || height % patch || width % patch) return EP_EUNSUPPORTED;
    const int n = channels * patch * patch;
    if (n < 2 || n > 6144) return EP_EUNSUPPORTED;
    const int warps = 8;
    const size_t smem = (size_t)warps * n * sizeof(float);
    const int64_t patches = (int64_t)batch * (height / patch) * (width / patch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)ep::ceil_div64(patches, warps);
    if (channels == 1 && (patch == 8 || patch == 16 || patch == 32) && width % 4 == 0 && ep::aligned16(frame) && ep::aligned16(out) &&
        (!loss || ep::aligned16(pred))) {
#define EP_TARGET_C1(P)                                                                                                            \
    if (loss) ep::k_patch_target_c1<P, true><<<blocks, warps * 32, 0, st>>>(frame, pred, patches, height, width, norm_pix, eps, mask, out); \
    else ep::k_patch_target_c1<P, false><<<blocks, warps * 32, 0, st>>>(frame, pred, patches, height, width, norm_pix, eps, mask, out)
        if (patch == 8) { EP_TARGET_C1(8); } else if (patch == 16) { EP_TARGET_C1(16); } else { EP_TARGET_C1(32); }
#undef EP_TARGET_C1
        EP_LAUNCH_CHECK();
        return EP_OK;
    }
    if (loss) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(ep::k_patch_target<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ep::k_patch_target<true><<<blocks, warps * 32, smem, st>>>(frame, pred, patches, channels, height, width, patch, norm_pix, eps, mask, out);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(ep::k_patch_target<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ep::k_patch_target<false><<<blocks, warps * 32, smem, st>>>(frame, pred, patches, channels, height, width, patch, norm_pix, eps, mask, out);
    }
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_patchify_normpix(void* stream, const float* frame, int batch, int channels, int height, int width, int patch,
                        int norm_pix, float eps, float* out) {
    return launch_patch_target(stream, false, frame, nullptr, batch, channels, height, width, patch, norm_pix, eps, nullptr, out);
}

int ep_target_patch_loss(void* stream, const float* frame, const float* pred, int batch, int channels, int height,
                         int width, int patch, int norm_pix, float eps, float* patch_loss) {
    return launch_patch_target(stream, true, frame, pred, batch, channels, height, width, patch, norm_pix, eps, nullptr, patch_loss);
}

int ep_target_patch_loss_masked(void* stream, const float* frame, const float* pred, const float* mask, int batch, int channels,
                                int height, int width, int patch, int norm_pix, float eps, float* patch_loss) {
    if (!mask) return EP_EINVAL;
    return launch_patch_target(stream, true, frame, pred, batch, channels, height, width, patch, norm_pix, eps, mask, patch_loss);
}

int ep_block_mask_expand(void* stream, const float* mask, int batch, int grid, int rep, int invert, float* out) {
    if (!mask || !out || batch <= 0 || grid <= 0 || rep <= 0) return EP_EINVAL;
    const int64_t total = (int64_t)batch * grid * rep * grid * rep;
    ep::k_block_mask_expand<<<(unsigned)ep::ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        mask, total, grid, rep, invert, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_swin_apply_mask(void* stream, const float* x, const float* mask_row, int batch, int Mh, int Mw, int rep, int C,
                       int n_vis_max, float* x_vis, int64_t* coords, uint8_t* vis_mask, int* n_vis_out) {
    if (!x || !mask_row || !x_vis || !coords || !vis_mask || !n_vis_out) return EP_EINVAL;
    if (batch <= 0 || Mh <= 0 || Mw <= 0 || rep <= 0 || C <= 0 || n_vis_max < 0) return EP_EINVAL;
    if (C % 4) return EP_EUNSUPPORTED;
    if (!ep::aligned16(x) || !ep::aligned16(x_vis)) return EP_EALIGN;
    const int Wt = Mw * rep, N = Mh * rep * Wt;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ep::k_swin_scan<<<1, 1024, 0, st>>>(mask_row, Mh, Mw, rep, n_vis_max, coords, vis_mask, n_vis_out);
    EP_LAUNCH_CHECK();
    if (n_vis_max > 0) {
        const int64_t rows = (int64_t)batch * n_vis_max;
        ep::k_swin_gather<<<(unsigned)ep::ceil_div64(rows, 8), 256, 0, st>>>(x, coords, n_vis_out, batch, N, Wt, C, n_vis_max, x_vis);
        EP_LAUNCH_CHECK();
    }
    return EP_OK;
}

int ep_swin_scatter_dense(void* stream, const float* x, const int64_t* coords, int batch, int n_vis, int grid, int C, float* out) {
    if (!x || !coords || !out || batch <= 0 || n_vis < 0 || grid <= 0 || C <= 0) return EP_EINVAL;
    if (batch > 65535) return EP_EUNSUPPORTED;
    const dim3 g((unsigned)((grid * grid + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)batch);
    ep::k_swin_scatter_dense<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, coords, n_vis, grid, C, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_gather_tokens_nchw(void* stream, const float* feat, const int64_t* ids_keep, int batch, int L, int K, int D, float* out) {
    if (!feat || !ids_keep || !out || batch <= 0 || L <= 0 || K <= 0 || D <= 0) return EP_EINVAL;
    if (batch > 65535) return EP_EUNSUPPORTED;
    ep::k_gather_tokens_nchw<<<dim3((unsigned)K, (unsigned)batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat, ids_keep, L, K, D, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_gather_sum_nchw(void* stream, const float* feat1, const float* feat2, const float* emb3, const int64_t* ids_keep, int batch,
                       int L, int K, int D, float* out) {
    if (!feat1 || !feat2 || !ids_keep || !out || batch <= 0 || L <= 0 || K <= 0 || D <= 0) return EP_EINVAL;
    if (batch > 65535) return EP_EUNSUPPORTED;
    ep::k_gather_sum_nchw<<<dim3((unsigned)K, (unsigned)batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat1, feat2, emb3, ids_keep, L, K, D, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

}  // extern "C"

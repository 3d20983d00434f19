// View augmentation: crop -> resize -> horizontal flip -> time flip, fused in one pass.
//
// Replaces the tensor half of evg_augment / frame_augment (dataset/augmentation/view_augment.py:9-89):
//   view_crop   (:9-33)   the crop box is drawn on the host (global numpy RNG) and passed per sample
//   view_resize (:35-39)  F.interpolate(mode, align_corners=None): nearest | bilinear | bicubic (A = -0.75)
//   view_horizontal_flip (:41-47), evg_time_flip (:49-58: reverse the bin axis, negate for 5/6 bins),
//   frame_time_flip (:60-63: negate)
#include "ep_common.cuh"

namespace ep {
namespace {

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ float fetch(const float* __restrict__ p, int W, int y, int x) { return __ldg(p + (int64_t)y * W + x); }

// One CTA per (sample, output channel, tile of kRows output rows); threads run along the output row.  Everything that
// depends only on the column (source column, taps, horizontal weights) is computed once per thread and reused down the
// tile; the sample and channel come from the block index, so no thread does an integer division.  (The first version
// decoded a flat 64-bit element index per thread — four 64-bit divisions for one load and one store — and ran at 10 %
// of the HBM roofline.)
constexpr int kRows = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k_view_augment(const float* __restrict__ in, int C, int H, int W,
                                                      const ep_view_params* __restrict__ prm, int OH, int OW, int n_row_tiles,
                                                      float* __restrict__ out) {
    const int tile = blockIdx.x % n_row_tiles;
    const int64_t bc = blockIdx.x / n_row_tiles;
    const int c_out = (int)(bc % C);
    const int64_t b = bc / C;
    const ep_view_params v = prm[b];
    const int c = v.time_flip ? C - 1 - c_out : c_out;              // torch.flip(dims=[0])
    const float* src = in + ((b * C + c) * (int64_t)H + v.crop_y) * W + v.crop_x;
    const int ch = v.crop_h, cw = v.crop_w;
    const float sh = (float)ch / (float)OH, sw = (float)cw / (float)OW;   // align_corners=False scales
    const int oy0 = tile * kRows, oy1 = min(oy0 + kRows, OH);
    float* orow = out + (bc * OH + oy0) * (int64_t)OW;
    for (int ox_out = threadIdx.x; ox_out < OW; ox_out += blockDim.x) {
        const int ox = v.hflip ? OW - 1 - ox_out : ox_out;          // flip is applied after the resize
        if (MODE == EP_RESIZE_NEAREST) {
            const int ix = min((int)floorf(ox * sw), cw - 1);
            float r[kRows];
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                const int oy = oy0 + k;
                if (oy < oy1) r[k] = fetch(src, W, min((int)floorf(oy * sh), ch - 1), ix);
            }
#pragma unroll
            for (int k = 0; k < kRows; ++k)
                if (oy0 + k < oy1) orow[(int64_t)k * OW + ox_out] = v.negate ? -r[k] : r[k];
        } else if (MODE == EP_RESIZE_BILINEAR) {
            const float fx = fmaxf(sw * (ox + 0.5f) - 0.5f, 0.f);
            const int x0 = (int)fx, x1 = x0 + (x0 < cw - 1);
            const float lx = fx - x0;
#pragma unroll 4
            for (int oy = oy0; oy < oy1; ++oy) {
                const float fy = fmaxf(sh * (oy + 0.5f) - 0.5f, 0.f);
                const int y0 = (int)fy, y1 = y0 + (y0 < ch - 1);
                const float ly = fy - y0;
                const float top = (1.f - lx) * fetch(src, W, y0, x0) + lx * fetch(src, W, y0, x1);
                const float bot = (1.f - lx) * fetch(src, W, y1, x0) + lx * fetch(src, W, y1, x1);
                const float r = (1.f - ly) * top + ly * bot;
                orow[(int64_t)(oy - oy0) * OW + ox_out] = v.negate ? -r : r;
            }
        } else {
            const float A = -0.75f;
            const float fx = sw * (ox + 0.5f) - 0.5f;
            const float xf = floorf(fx), tx = fx - xf;
            const float wx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
            int xx[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) xx[j] = min(max((int)xf - 1 + j, 0), cw - 1);
            for (int oy = oy0; oy < oy1; ++oy) {
                const float fy = sh * (oy + 0.5f) - 0.5f;
                const float yf = floorf(fy), ty = fy - yf;
                const float wy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
                float r = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int yy = min(max((int)yf - 1 + i, 0), ch - 1);
                    float row = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) row += wx[j] * fetch(src, W, yy, xx[j]);
                    r += wy[i] * row;
                }
                orow[(int64_t)(oy - oy0) * OW + ox_out] = v.negate ? -r : r;
            }
        }
    }
}

// Bilinear resize with the source rows of the tile staged in shared memory.  The direct kernel above issues four scalar
// loads per output element and is bound by the load/store unit (42 % of the HBM peak on the 346x260 -> 224x224 copy of a
// 512 x 9 plane batch); here the rows a tile of R output rows needs — one contiguous span of the input tensor, whole rows
// from the first tap row to the last — arrive as aligned 16-byte loads, and the four taps of an output element are
// shared-memory reads.  Same expressions in the same order as the direct kernel: bit-identical.
constexpr int kStageRows = 16;        // output rows per CTA (the host halves it until the source rows fit the staging buffer)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(256) k_view_bilinear_staged(const float* __restrict__ in, int C, int H, int W, int64_t n_in,
                                                              const ep_view_params* __restrict__ prm, int OH, int OW, int R,
                                                              int n_row_tiles, int cap, float* __restrict__ out) {
    extern __shared__ __align__(16) float s_src[];
    __shared__ int s_r0[kStageRows], s_r1[kStageRows];       // per output row of the tile: offsets of its two tap rows in s_src
    __shared__ float s_ly[kStageRows];
    const int tile = blockIdx.x % n_row_tiles;
    const int64_t bc = blockIdx.x / n_row_tiles;
    const int c_out = (int)(bc % C);
    const int64_t b = bc / C;
    const ep_view_params v = prm[b];
    const int c = v.time_flip ? C - 1 - c_out : c_out;
    const int ch = v.crop_h, cw = v.crop_w;
    const float sh = (float)ch / (float)OH, sw = (float)cw / (float)OW;
    const int oy0 = tile * R, nrow = min(R, OH - oy0);
    // tap rows of the tile: y0 of its first output row .. y1 of its last
    const int y_lo = (int)fmaxf(sh * (oy0 + 0.5f) - 0.5f, 0.f);
    const int y_last0 = (int)fmaxf(sh * (oy0 + nrow - 1 + 0.5f) - 0.5f, 0.f);
    const int y_hi = y_last0 + (y_last0 < ch - 1);
    const int64_t g0 = ((b * C + c) * (int64_t)H + v.crop_y + y_lo) * W + v.crop_x;      // first needed element
    const int64_t a0 = g0 & ~(int64_t)3;                                                  // aligned start of the staged span
    const int off = (int)(g0 - a0);
    const int span = off + (y_hi - y_lo) * W + cw;
    const bool staged = span <= cap;
    if (staged) {
        const float* gsrc = in + a0;
        for (int i = threadIdx.x * 4; i < span; i += 256 * 4) {
            if (a0 + i + 4 <= n_in) cp_async16(s_src + i, gsrc + i);
            else for (int j = 0; j < 4; ++j) if (a0 + i + j < n_in) s_src[i + j] = __ldg(gsrc + i + j);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (threadIdx.x < nrow) {
        const float fy = fmaxf(sh * (oy0 + (int)threadIdx.x + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)fy, y1 = y0 + (y0 < ch - 1);
        s_ly[threadIdx.x] = fy - y0;
        s_r0[threadIdx.x] = staged ? off + (y0 - y_lo) * W : y0 * W;
        s_r1[threadIdx.x] = staged ? off + (y1 - y_lo) * W : y1 * W;
    }
    if (staged) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float* src = in + ((b * C + c) * (int64_t)H + v.crop_y) * W + v.crop_x;
    for (int ox_out = threadIdx.x; ox_out < OW; ox_out += 256) {
        const int ox = v.hflip ? OW - 1 - ox_out : ox_out;
        const float fx = fmaxf(sw * (ox + 0.5f) - 0.5f, 0.f);
        const int x0 = (int)fx, x1 = x0 + (x0 < cw - 1);
        const float lx = fx - x0;
        float* o = out + (bc * OH + oy0) * (int64_t)OW + ox_out;
#pragma unroll 4
        for (int k = 0; k < nrow; ++k, o += OW) {
            const int r0 = s_r0[k], r1 = s_r1[k];
            const float ly = s_ly[k];
            float t0, t1, b0, b1;
            if (staged) { t0 = s_src[r0 + x0]; t1 = s_src[r0 + x1]; b0 = s_src[r1 + x0]; b1 = s_src[r1 + x1]; }
            else { t0 = __ldg(src + r0 + x0); t1 = __ldg(src + r0 + x1); b0 = __ldg(src + r1 + x0); b1 = __ldg(src + r1 + x1); }
            const float top = (1.f - lx) * t0 + lx * t1;
            const float bot = (1.f - lx) * b0 + lx * b1;
            const float r = (1.f - ly) * top + ly * bot;
            *o = v.negate ? -r : r;
        }
    }
}

}  // namespace
}  // namespace ep

extern "C" int ep_view_augment(void* stream, const float* in, int batch, int channels, int height, int width,
                               const ep_view_params* params, int out_h, int out_w, int mode, float* out) {
    if (!in || !out || !params || batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || out_h <= 0 || out_w <= 0)
        return EP_EINVAL;
    if (mode != EP_RESIZE_NEAREST && mode != EP_RESIZE_BILINEAR && mode != EP_RESIZE_BICUBIC) return EP_EINVAL;
    const int n_row_tiles = (out_h + ep::kRows - 1) / ep::kRows;
    const int64_t blocks = (int64_t)batch * channels * n_row_tiles;
    if (blocks > 0x7fffffffLL) return EP_EUNSUPPORTED;
    const int threads = out_w >= 256 ? 256 : (out_w + 31) / 32 * 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == EP_RESIZE_NEAREST)
        ep::k_view_augment<EP_RESIZE_NEAREST><<<(unsigned)blocks, threads, 0, st>>>(in, channels, height, width, params, out_h, out_w, n_row_tiles, out);
    else if (mode == EP_RESIZE_BILINEAR && ep::aligned16(in)) {
        // staged form: R output rows per CTA, as many as keep the worst-case span (crop = the whole frame) within 48 KB
        const int64_t cap_max = (48 * 1024 - 512) / 4;      // (the kernel also holds 192 B of static shared memory)
        int R = ep::kStageRows;
        auto span_of = [&](int r) { return (((int64_t)r * height + out_h - 1) / out_h + 2) * width + 4; };
        while (R > 1 && span_of(R) > cap_max) R >>= 1;
        const int cap = (int)(span_of(R) < cap_max ? span_of(R) : cap_max);       // wider spans take the direct loads inside the kernel
        const int tiles = (out_h + R - 1) / R;
        const int64_t nb = (int64_t)batch * channels * tiles;
        if (nb > 0x7fffffffLL) return EP_EUNSUPPORTED;
        ep::k_view_bilinear_staged<<<(unsigned)nb, 256, (size_t)cap * 4 + 16, st>>>(in, channels, height, width,
                                                                                    (int64_t)batch * channels * height * width, params, out_h,
                                                                                    out_w, R, tiles, cap, out);
    } else if (mode == EP_RESIZE_BILINEAR)
        ep::k_view_augment<EP_RESIZE_BILINEAR><<<(unsigned)blocks, threads, 0, st>>>(in, channels, height, width, params, out_h, out_w, n_row_tiles, out);
    else
        ep::k_view_augment<EP_RESIZE_BICUBIC><<<(unsigned)blocks, threads, 0, st>>>(in, channels, height, width, params, out_h, out_w, n_row_tiles, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

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
    else if (mode == EP_RESIZE_BILINEAR)
        ep::k_view_augment<EP_RESIZE_BILINEAR><<<(unsigned)blocks, threads, 0, st>>>(in, channels, height, width, params, out_h, out_w, n_row_tiles, out);
    else
        ep::k_view_augment<EP_RESIZE_BICUBIC><<<(unsigned)blocks, threads, 0, st>>>(in, channels, height, width, params, out_h, out_w, n_row_tiles, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

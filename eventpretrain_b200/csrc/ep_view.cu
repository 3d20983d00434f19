// View augmentation: crop -> resize -> horizontal flip -> time flip, fused in one pass.
//
// Replaces the tensor half of evg_augment / frame_augment (dataset/augmentation/view_augment.py:9-89):
//   view_crop   (:9-33)   the crop box is drawn on the host (global numpy RNG) and passed per sample
//   view_resize (:35-39)  F.interpolate(mode, align_corners=None): nearest | bilinear | bicubic (A = -0.75)
//   view_horizontal_flip (:41-47), evg_time_flip (:49-58: reverse the bin axis, negate for 5/6 bins),
//   frame_time_flip (:60-63: negate)
// One thread per output element; the source box of a sample is small enough to live in L1/L2.
#include "ep_common.cuh"

namespace ep {
namespace {

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ float fetch(const float* __restrict__ p, int W, int y, int x) { return __ldg(p + (int64_t)y * W + x); }

__global__ void __launch_bounds__(256) k_view_augment(const float* __restrict__ in, int C, int H, int W,
                                                      const ep_view_params* __restrict__ prm, int OH, int OW, int mode,
                                                      int64_t total, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox_out = (int)(idx % OW);
    const int oy = (int)((idx / OW) % OH);
    const int c_out = (int)((idx / ((int64_t)OW * OH)) % C);
    const int64_t b = idx / ((int64_t)OW * OH * C);
    const ep_view_params v = prm[b];
    const int ox = v.hflip ? OW - 1 - ox_out : ox_out;              // flip is applied after the resize
    const int c = v.time_flip ? C - 1 - c_out : c_out;              // torch.flip(dims=[0])
    const float* src = in + ((b * C + c) * (int64_t)H + v.crop_y) * W + v.crop_x;
    const int ch = v.crop_h, cw = v.crop_w;
    const float sh = (float)ch / (float)OH, sw = (float)cw / (float)OW;   // align_corners=False scales
    float r;
    if (mode == EP_RESIZE_NEAREST) {
        const int iy = min((int)floorf(oy * sh), ch - 1), ix = min((int)floorf(ox * sw), cw - 1);
        r = fetch(src, W, iy, ix);
    } else if (mode == EP_RESIZE_BILINEAR) {
        const float fy = fmaxf(sh * (oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sw * (ox + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < ch - 1), x1 = x0 + (x0 < cw - 1);
        const float ly = fy - y0, lx = fx - x0;
        const float top = (1.f - lx) * fetch(src, W, y0, x0) + lx * fetch(src, W, y0, x1);
        const float bot = (1.f - lx) * fetch(src, W, y1, x0) + lx * fetch(src, W, y1, x1);
        r = (1.f - ly) * top + ly * bot;
    } else {
        const float A = -0.75f;
        const float fy = sh * (oy + 0.5f) - 0.5f, fx = sw * (ox + 0.5f) - 0.5f;
        const float yf = floorf(fy), xf = floorf(fx);
        const float ty = fy - yf, tx = fx - xf;
        const float wy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
        const float wx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
        r = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int yy = min(max((int)yf - 1 + i, 0), ch - 1);
            float row = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = min(max((int)xf - 1 + j, 0), cw - 1);
                row += wx[j] * fetch(src, W, yy, xx);
            }
            r += wy[i] * row;
        }
    }
    out[idx] = v.negate ? -r : r;
}

}  // namespace
}  // namespace ep

extern "C" int ep_view_augment(void* stream, const float* in, int batch, int channels, int height, int width,
                               const ep_view_params* params, int out_h, int out_w, int mode, float* out) {
    if (!in || !out || !params || batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || out_h <= 0 || out_w <= 0)
        return EP_EINVAL;
    if (mode != EP_RESIZE_NEAREST && mode != EP_RESIZE_BILINEAR && mode != EP_RESIZE_BICUBIC) return EP_EINVAL;
    const int64_t total = (int64_t)batch * channels * out_h * out_w;
    const int64_t blocks = ep::ceil_div64(total, 256);
    if (blocks > 0x7fffffffLL) return EP_EUNSUPPORTED;
    ep::k_view_augment<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, channels, height, width, params,
                                                                                        out_h, out_w, mode, total, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

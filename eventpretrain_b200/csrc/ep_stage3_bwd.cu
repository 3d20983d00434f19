// Stage 3, backward passes of the token permutations (sm_100a) — what torch.autograd would derive for the index ops of
//   model/backbone/vit.py:113-115 (gather by ids_keep, + pos_embed), model/pretrain/pr_rec_decoder.py:56-62 (un-shuffle),
//   model/backbone/swin.py:154-179, :221-228 (apply_mask, scatter to the dense grid, gather by ids_keep),
//   model/sub_module/swin_block.py:446-464 (GroupingModule.group / merge: index_select along the tokens).
// The reference trains through these ops (pr_trainer.py:26-36), so the drop-ins carry a gradient: the Python side wraps
// each forward kernel in a torch.autograd.Function whose backward is one of the kernels below.  All of them are row
// permutations / row sums (HBM-bound, no contraction); duplicates in the indices (GroupingModule pads groups with token
// 0) are accumulated with fp32 atomics, like torch's own index_add.
#include "ep_common.cuh"

namespace ep {
namespace {

__device__ __forceinline__ void red_add4(float* p, float4 v) {
    // sm_90+: one vector reduction instead of four scalar ones
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// grad_tokens[b, ids[b or 0, k], :] += g[b, k, :]     (one warp per row; grad_tokens zeroed by the caller)
__global__ void __launch_bounds__(256) k_scatter_add_rows(const float* __restrict__ g, const int64_t* __restrict__ ids, int ids_shared,
                                                          int64_t rows, int L, int K, int D, float* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = row / K, k = row - b * K;
    int64_t dst = ids[ids_shared ? k : row];
    if (dst < 0 || dst >= L) dst = 0;                 // the forward kernel clamps the same way
    const float4* s = reinterpret_cast<const float4*>(g + row * D);
    float* o = out + (b * L + dst) * D;
    for (int i = lane; i < D / 4; i += 32) red_add4(o + 4 * i, ld_stream(s + i));
}

// out[n] = sum_b in[b, n] in the fixed order b = 0 .. B-1 (pos_embed / mask_token gradients: reproducible)
__global__ void __launch_bounds__(256) k_sum_over_batch(const float* __restrict__ in, int B, int64_t N4, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
        const float4 v = ld_stream(reinterpret_cast<const float4*>(in) + (int64_t)b * N4 + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
}

// un-shuffle backward: row (b, l) of g goes to grad_emb[b, ids_restore[b, l], :] when that is a kept token, else into the
// per-sample mask-token partial part[b, :] (fp32 atomics within a sample; the batch is then summed in a fixed order).
__global__ void __launch_bounds__(256) k_unshuffle_bwd(const float* __restrict__ g, const int64_t* __restrict__ ids_restore, int64_t rows,
                                                       int L, int K, int D, float* __restrict__ g_emb, float* __restrict__ part) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = row / L;
    const int64_t src = ids_restore[row];
    const float4* s = reinterpret_cast<const float4*>(g + row * D);
    float* o = (src >= 0 && src < K) ? g_emb + (b * K + src) * D : part + b * D;
    for (int i = lane; i < D / 4; i += 32) red_add4(o + 4 * i, ld_stream(s + i));
}

// gather_tokens_nchw backward: grad_feat[b, d, ids[b, k]] += g[b, k, d]   (grad_feat zeroed by the caller)
__global__ void __launch_bounds__(256) k_scatter_add_nchw(const float* __restrict__ g, const int64_t* __restrict__ ids, int L, int K, int D,
                                                          float* __restrict__ out) {
    const int b = blockIdx.y, k = blockIdx.x;
    int64_t dst = ids[(int64_t)b * K + k];
    if (dst < 0 || dst >= L) dst = 0;
    for (int d = threadIdx.x; d < D; d += 256) atomicAdd(out + ((int64_t)b * D + d) * L + dst, g[((int64_t)b * K + k) * D + d]);
}

}  // namespace
}  // namespace ep

extern "C" {

int ep_scatter_add_tokens(void* stream, const float* grad, const int64_t* ids, int ids_shared, int batch, int L, int K, int D,
                          float* out) {
    if (!grad || !ids || !out || batch <= 0 || L <= 0 || K <= 0 || D <= 0) return EP_EINVAL;
    if (D % 4) return EP_EUNSUPPORTED;
    if (!ep::aligned16(grad) || !ep::aligned16(out)) return EP_EALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t ce = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)batch * L * D, st);
    if (ce != cudaSuccess) return (int)ce;
    const int64_t rows = (int64_t)batch * K;
    ep::k_scatter_add_rows<<<(unsigned)ep::ceil_div64(rows, 8), 256, 0, st>>>(grad, ids, ids_shared, rows, L, K, D, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_sum_over_batch(void* stream, const float* in, int batch, int64_t n, float* out) {
    if (!in || !out || batch <= 0 || n <= 0) return EP_EINVAL;
    if (n % 4) return EP_EUNSUPPORTED;
    if (!ep::aligned16(in) || !ep::aligned16(out)) return EP_EALIGN;
    ep::k_sum_over_batch<<<(unsigned)ep::ceil_div64(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, batch, n / 4, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_unshuffle_tokens_bwd(void* stream, const float* grad, const int64_t* ids_restore, int batch, int L, int K, int D,
                            float* grad_emb, float* grad_mask_token, float* scratch) {
    if (!grad || !ids_restore || !grad_emb || !grad_mask_token || !scratch || batch <= 0 || L <= 0 || K < 0 || K > L || D <= 0)
        return EP_EINVAL;
    if (D % 4) return EP_EUNSUPPORTED;
    if (!ep::aligned16(grad) || !ep::aligned16(grad_emb) || !ep::aligned16(grad_mask_token) || !ep::aligned16(scratch)) return EP_EALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t ce = cudaMemsetAsync(grad_emb, 0, sizeof(float) * (size_t)batch * (K > 0 ? K : 1) * D, st);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)batch * D, st);
    if (ce != cudaSuccess) return (int)ce;
    const int64_t rows = (int64_t)batch * L;
    ep::k_unshuffle_bwd<<<(unsigned)ep::ceil_div64(rows, 8), 256, 0, st>>>(grad, ids_restore, rows, L, K, D, grad_emb, scratch);
    EP_LAUNCH_CHECK();
    ep::k_sum_over_batch<<<(unsigned)ep::ceil_div64(D / 4, 256), 256, 0, st>>>(scratch, batch, D / 4, grad_mask_token);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

int ep_scatter_add_tokens_nchw(void* stream, const float* grad, const int64_t* ids_keep, int batch, int L, int K, int D, float* out) {
    if (!grad || !ids_keep || !out || batch <= 0 || L <= 0 || K <= 0 || D <= 0) return EP_EINVAL;
    if (batch > 65535) return EP_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t ce = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)batch * D * L, st);
    if (ce != cudaSuccess) return (int)ce;
    ep::k_scatter_add_nchw<<<dim3((unsigned)K, (unsigned)batch), 256, 0, st>>>(grad, ids_keep, L, K, D, out);
    EP_LAUNCH_CHECK();
    return EP_OK;
}

}  // extern "C"

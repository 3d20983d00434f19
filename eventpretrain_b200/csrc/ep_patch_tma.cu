// Visible-patch gather with the Tensor Memory Accelerator (sm_100a) — the (c, ph, pw) order of ep_patchify_gather.
//
// x (B,C,H,W) fp32 is described to the TMA as a 3-D tensor (W, H, B*C); one kept patch is the box (p, p, C) at
// (px*p, py*p, b*C), and it lands in shared memory as [C][p][p] — which IS the output row out[b, k, :] in Conv2d weight
// order.  So a patch is two bulk asynchronous copies and no data instruction at all:
//     cp.async.bulk.tensor.3d  global -> shared   (completion on an mbarrier)
//     cp.async.bulk            shared -> global   (bulk group)
// One thread per CTA drives a ring of kStages tiles; the other lanes only fetch the next ids.  Many small CTAs per SM keep
// ~100 KB of copies in flight per SM.  model/backbone/vit.py:110-115 (the per-patch ops of PatchEmbed commute with the gather).
#include <cuda.h>

#include "ep_common.cuh"

namespace ep {
namespace {

constexpr int kStages = 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(32) k_patch_gather_tma(const __grid_constant__ CUtensorMap tmap, const int64_t* __restrict__ ids,
                                                         int C, int gw, int L, int p, int K, int64_t rows, int tile_bytes,
                                                         int tile_stride, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char tiles[];          // [kStages][tile_stride]
    __shared__ __align__(8) unsigned long long bars[kStages];
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(smem_addr(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t first = blockIdx.x, step = gridDim.x;
    const int64_t mine = first < rows ? (rows - first + step - 1) / step : 0;
    // patch index of my j-th row, fetched 32 rows at a time by the whole warp
    int64_t id_block = -1;
    int64_t id_lane = 0;
    auto patch_of = [&](int64_t j) -> int64_t {
        if ((j >> 5) != id_block) {
            id_block = j >> 5;
            const int64_t jj = (id_block << 5) + lane;
            const int64_t row = first + jj * step;
            id_lane = (jj < mine) ? (ids ? ids[row] : row % K) : 0;
        }
        return __shfl_sync(0xffffffffu, id_lane, (int)(j & 31));
    };
    auto issue = [&](int64_t j) {
        int64_t l = patch_of(j);
        if (l < 0 || l >= L) l = 0;
        if (lane == 0) {
            const int64_t row = first + j * step;
            const int s = (int)(j % kStages);
            const uint32_t bar = smem_addr(&bars[s]);
            mbar_expect_tx(bar, (uint32_t)tile_bytes);
            tma_load_3d(smem_addr(tiles + (size_t)s * tile_stride), &tmap, (int)(l % gw) * p, (int)(l / gw) * p, (int)(row / K) * C, bar);
        }
    };
    const int64_t ahead = mine < kStages ? mine : kStages;
    for (int64_t j = 0; j < ahead; ++j) issue(j);
    for (int64_t j = 0; j < mine; ++j) {
        const int s = (int)(j % kStages);
        if (lane == 0) {
            mbar_wait(smem_addr(&bars[s]), (uint32_t)((j / kStages) & 1));
            bulk_store(out + (first + j * step) * (int64_t)(tile_bytes / 4), smem_addr(tiles + (size_t)s * tile_stride), (uint32_t)tile_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (j + kStages < mine) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile has been read: reusable
            issue(j + kStages);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        cudaGetLastError();
    }
    return fn;
}

}  // namespace

// Returns EP_EUNSUPPORTED when the shape does not qualify for the TMA form (the caller then runs the plain kernel).
int patchify_gather_tma(cudaStream_t st, const float* x, const int64_t* ids_keep, int batch, int channels, int height, int width,
                        int patch, int K, float* out) {
    const int tile_bytes = channels * patch * patch * 4;
    if (patch % 4 != 0 || width % 4 != 0 || patch > 256 || channels > 256 || tile_bytes > 48 * 1024) return EP_EUNSUPPORTED;
    if (!aligned16(x) || !aligned16(out) || (int64_t)batch * channels > 0x7fffffffLL) return EP_EUNSUPPORTED;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return EP_EUNSUPPORTED;
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)height, (cuuint64_t)batch * channels};
    const cuuint64_t strides[2] = {(cuuint64_t)width * 4, (cuuint64_t)width * height * 4};
    const cuuint32_t box[3] = {(cuuint32_t)patch, (cuuint32_t)patch, (cuuint32_t)channels};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return EP_EUNSUPPORTED;
    const int tile_stride = (tile_bytes + 127) / 128 * 128;
    const size_t smem = (size_t)kStages * tile_stride;
    static size_t configured = 0;
    if (smem > 48 * 1024 && configured < smem) {
        if (cudaFuncSetAttribute(k_patch_gather_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return EP_EUNSUPPORTED;
        }
        configured = smem;
    }
    const int64_t rows = (int64_t)batch * K;
    int per_sm = (int)((200 * 1024) / (smem + 256));
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)kNumSMs * per_sm;
    if (grid > rows) grid = rows;
    const int gw = width / patch, L = (height / patch) * gw;
    k_patch_gather_tma<<<(unsigned)grid, 32, smem, st>>>(map, ids_keep, channels, gw, L, patch, K, rows, tile_bytes, tile_stride, out);
    return EP_OK;
}

}  // namespace ep

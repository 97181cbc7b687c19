// tn_mlp_pack.cuh -- weight image for the tensor-core MLP.
// A Linear weight W[N=128][row_stride] (torch nn.Linear layout, in-features contiguous) restricted to
// columns [k0, k0+K) is split into bf16 hi/lo and stored as the exact shared-memory image the UMMA
// descriptor expects: for each 64-wide K block kb: [hi block 16 KB][lo block 16 KB], each block being
// 128 rows x 128 bytes with the 128-byte swizzle (tc::sw128_offset).  K must be a multiple of 64.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tn_tc.cuh"

namespace tn {

// kb_stride: bytes between consecutive 64-wide K blocks; lo_off: bytes from a hi block to its lo block.  The forward image uses
// (32768, 16384): per K block [hi][lo]; the backward image (tn_mlp_bwd.cuh) uses (16384, 32768): [hi kb0][hi kb1][lo kb0][lo kb1].
// fp16 != 0: the hi / lo halves are IEEE half precision instead of bfloat16 (the "f16w2" forward mode, tn_mlp.cuh).
static __global__ void k_pack_weights(const float *__restrict__ W, uint32_t row_stride, uint32_t k0, uint32_t K, uint8_t *__restrict__ img,
                                      uint32_t kb_stride, uint32_t lo_off, int fp16) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128u * K) return;
    const uint32_t n = idx / K, k = idx % K;
    const float w = W[(size_t)n * row_stride + k0 + k];
    const uint32_t kb = k >> 6, kk = k & 63u;
    const uint32_t off = kb * kb_stride + tc::sw128_offset(n, kk);
    if (fp16) {
        const __half hi = __float2half_rn(fminf(fmaxf(w, -65504.0f), 65504.0f));  // saturate (a weight that large has left fp16 anyway)
        const __half lo = __float2half_rn(w - __half2float(hi));
        *reinterpret_cast<__half *>(img + off) = hi;
        *reinterpret_cast<__half *>(img + off + lo_off) = lo;
    } else {
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        *reinterpret_cast<__nv_bfloat16 *>(img + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(img + off + lo_off) = lo;
    }
}

static inline void launch_pack_weights(const float *W, uint32_t row_stride, uint32_t k0, uint32_t K, uint8_t *img, cudaStream_t s,
                                       uint32_t kb_stride = 32768u, uint32_t lo_off = 16384u, int fp16 = 0) {
    k_pack_weights<<<(128u * K + 255u) / 256u, 256, 0, s>>>(W, row_stride, k0, K, img, kb_stride, lo_off, fp16);
}

}  // namespace tn

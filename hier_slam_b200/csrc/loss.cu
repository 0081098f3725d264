// Masked L1 image loss with its gradient in one pass (first step of the fused loss
// epilogue).  Hier-SLAM's tracking / mapping losses are sums of |target - rendered| over a boolean pixel mask
// (scripts/hierslam.py:780-796: `torch.abs(gt - x)[mask].sum()`); boolean indexing costs a nonzero() with a host sync in
// the forward and an index_put_ in the backward -- 1.1 ms of a 3.4 ms tracking iteration at 1200x680, more than the
// rasterizer itself.  This kernel reads the rendered image, the target and the mask once, reduces the masked sum
// (warp shuffle -> shared memory -> one atomic per block) and writes d loss / d rendered = mask * sign(x - gt), so the
// autograd backward is a single scale by the upstream scalar.
#include "hs_common.cuh"

namespace hs {

__global__ void __launch_bounds__(256) masked_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const uint8_t* __restrict__ mask, int C, size_t HW,
                                                        float* __restrict__ loss, float* __restrict__ grad) {
    __shared__ float s_part[8];
    float acc = 0.f;
    const size_t n = (size_t)C * HW;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const size_t p = i % HW;                       // the mask is [H,W], shared by the C channels
        const bool m = mask == nullptr || mask[p] != 0;
        const float d = pred[i] - target[i];
        acc += m ? fabsf(d) : 0.f;
        grad[i] = m ? (d > 0.f ? 1.f : d < 0.f ? -1.f : 0.f) : 0.f;   // torch.abs has a zero subgradient at 0
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_part[w];
        atomicAdd(loss, t);
    }
}

int launch_masked_l1(const float* pred, const float* target, const uint8_t* mask, int C, size_t HW, float* loss,
                     float* grad, cudaStream_t stream) {
    if (C <= 0 || HW == 0) return 0;
    const size_t n = (size_t)C * HW;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;            // grid-stride: 8 CTAs per SM
    masked_l1_kernel<<<blocks, 256, 0, stream>>>(pred, target, mask, C, HW, loss, grad);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

}  // namespace hs

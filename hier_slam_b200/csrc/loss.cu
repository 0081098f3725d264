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

// Hierarchical cross-entropy of Hier-SLAM's tree encoding (scripts/hierslam.py:955-1000, transfer_tree_rendered_labelmap
// :91-111): the S rendered channels are the concatenation of L levels; level l owns channels [begin[l], begin[l+1]) and
// contributes weight * mean_pixels CE(softmax(slice), label_l).  The reference evaluates every level with a permute +
// view (a transposing copy of the slice), a CrossEntropyLoss forward and its backward; this kernel reads each channel
// plane once more than strictly needed (max/sum pass, gradient pass) and writes d loss / d sem directly in the planar
// [S,H,W] layout the rasterizer's backward consumes.  One thread per pixel, coalesced across the plane.
namespace hs {

struct HierLevels {
    int L;
    int begin[HS_MAX_LEVELS + 1];
    float scale[HS_MAX_LEVELS];   // weight_l / (number of pixels whose label is not ignored)
};

__global__ void __launch_bounds__(256) hier_ce_kernel(const float* __restrict__ sem, const int* __restrict__ labels,
                                                      HierLevels lv, size_t HW, float* __restrict__ loss,
                                                      float* __restrict__ grad) {
    __shared__ float s_part[8];
    float acc = 0.f;
    for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < HW; p += (size_t)gridDim.x * 256) {
        for (int l = 0; l < lv.L; l++) {
            const int b = lv.begin[l], e = lv.begin[l + 1];
            const int y = labels[(size_t)l * HW + p];
            const bool use = y >= 0 && y < e - b;           // torch's ignore_index (-100) and out-of-range labels
            float m = -3.0e38f;
            for (int c = b; c < e; c++) m = fmaxf(m, sem[(size_t)c * HW + p]);
            float z = 0.f;
            for (int c = b; c < e; c++) z += __expf(sem[(size_t)c * HW + p] - m);
            const float inv = 1.f / z, sc = use ? lv.scale[l] : 0.f;
            for (int c = b; c < e; c++) {
                const float x = sem[(size_t)c * HW + p];
                const float sm = __expf(x - m) * inv;
                grad[(size_t)c * HW + p] = sc * (sm - ((c - b) == y ? 1.f : 0.f));
                if (use && (c - b) == y) acc += sc * (m + __logf(z) - x);
            }
        }
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

int launch_hier_cross_entropy(const float* sem, const int* labels, int L, const int* level_begin, const float* level_scale,
                              size_t HW, float* loss, float* grad, cudaStream_t stream) {
    if (L <= 0 || HW == 0) return 0;
    if (L > HS_MAX_LEVELS) {
        set_error("hierarchical cross-entropy: at most %d levels", HS_MAX_LEVELS);
        return 1;
    }
    HierLevels lv;
    lv.L = L;
    for (int l = 0; l <= L; l++) lv.begin[l] = level_begin[l];
    for (int l = 0; l < L; l++) lv.scale[l] = level_scale[l];
    int blocks = (int)((HW + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    hier_ce_kernel<<<blocks, 256, 0, stream>>>(sem, labels, lv, HW, loss, grad);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

}  // namespace hs

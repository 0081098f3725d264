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
// plane exactly once (a level's channels live in registers) and writes d loss / d sem directly in the planar [S,H,W]
// layout the rasterizer's backward consumes.  One thread per pixel, coalesced across the plane.
namespace hs {

struct HierLevels {
    int L;
    int begin[HS_MAX_LEVELS + 1];
    float scale[HS_MAX_LEVELS];   // weight_l / (number of pixels whose label is not ignored)
};

// One level of one pixel with the level's channels held in registers: every channel plane is read ONCE (all loads of a
// level are independent and in flight together), one exp per channel.  N = register capacity (channels of the level <= N).
template <int N>
__device__ __forceinline__ float level_ce(const float* __restrict__ sem, float* __restrict__ grad, size_t HW, size_t p, int b,
                                          int n, int y, float scale) {
    float x[N];
#pragma unroll
    for (int c = 0; c < N; c++) x[c] = c < n ? sem[(size_t)(b + c) * HW + p] : -3.0e38f;
    float m = x[0];
#pragma unroll
    for (int c = 1; c < N; c++) m = fmaxf(m, x[c]);
    float z = 0.f, xy = 0.f;
#pragma unroll
    for (int c = 0; c < N; c++) {
        if (c == y) xy = x[c];
        x[c] = c < n ? __expf(x[c] - m) : 0.f;
        z += x[c];
    }
    const bool use = y >= 0 && y < n;                   // torch's ignore_index (-100) and out-of-range labels
    const float sc = use ? scale : 0.f, inv = sc / z;
#pragma unroll
    for (int c = 0; c < N; c++)
        if (c < n) grad[(size_t)(b + c) * HW + p] = x[c] * inv - (c == y ? sc : 0.f);
    return use ? sc * (m + __logf(z) - xy) : 0.f;
}

__global__ void __launch_bounds__(256) hier_ce_kernel(const float* __restrict__ sem, const int* __restrict__ labels,
                                                      HierLevels lv, size_t HW, float* __restrict__ loss,
                                                      float* __restrict__ grad) {
    __shared__ float s_part[8];
    float acc = 0.f;
    for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < HW; p += (size_t)gridDim.x * 256) {
        for (int l = 0; l < lv.L; l++) {
            const int b = lv.begin[l], n = lv.begin[l + 1] - b;
            const int y = labels[(size_t)l * HW + p];
            if (n <= 8) acc += level_ce<8>(sem, grad, HW, p, b, n, y, lv.scale[l]);
            else if (n <= 16) acc += level_ce<16>(sem, grad, HW, p, b, n, y, lv.scale[l]);
            else acc += level_ce<32>(sem, grad, HW, p, b, n, y, lv.scale[l]);
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
    for (int l = 0; l < L; l++)
        if (level_begin[l + 1] - level_begin[l] < 1 || level_begin[l + 1] - level_begin[l] > 32) {
            set_error("hierarchical cross-entropy: 1..32 channels per level supported (level %d has %d)", l,
                      level_begin[l + 1] - level_begin[l]);
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

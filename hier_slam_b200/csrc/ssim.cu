// Colour loss of Hier-SLAM's mapping (scripts/hierslam.py:936: 0.8 * l1_loss_v1(im, gt) + 0.2 * (1 - calc_ssim(im, gt)))
// with its gradient in two kernels.  The reference's SSIM (utils/slam_external.py:55-97) is
// five depthwise 11x11 Gaussian convolutions (zero padding 5) of x, y, x^2, y^2, xy, an elementwise map and its mean;
// autograd then runs five more convolutions backward.  Here:
//   ssim_forward_kernel   per 32x16 pixel tile and channel: both images with a 5-pixel halo in shared memory, the five
//                         filtered moments by a separable horizontal + vertical pass, the SSIM value, and the three
//                         partial derivatives d ssim / d (mu1, E[x^2], E[xy]) per pixel (already multiplied by the loss
//                         scale); the L1 term rides along;
//   ssim_backward_kernel  d loss / d x[p] = conv(dmu1)[p] + 2 x[p] conv(dE11)[p] + y[p] conv(dE12)[p] (the window is
//                         symmetric, so the adjoint of the filter is the filter) + the L1 sign term.
// With A = 2 mu1 mu2 + c1, B = 2 s12 + c2, C = mu1^2 + mu2^2 + c1, D = s1 + s2 + c2, ssim = A B / (C D):
//   d ssim / d mu1 = 2 mu2 (B - A) / (C D) - 2 mu1 ssim (D - C) / (C D),  d ssim / d E11 = -ssim / D,  d ssim / d E12 = 2 A / (C D).
#include "hs_common.cuh"

namespace hs {
namespace ssim {

constexpr int TW = 32, TH = 16, R = 5, K = 2 * R + 1;     // tile, window radius, taps
constexpr int IW = TW + 2 * R, IH = TH + 2 * R;           // tile with halo: 42 x 26

struct Window { float w[K]; };

__device__ __forceinline__ float block_sum(float v, float* s_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0)
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_part[w];
    return t;
}

__global__ void __launch_bounds__(256) ssim_forward_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                           int H, int W, Window win, float l1_scale, float ssim_scale,
                                                           float* __restrict__ loss, float* __restrict__ dmu1,
                                                           float* __restrict__ dE11, float* __restrict__ dE12) {
    __shared__ float s_x[IH][IW], s_y[IH][IW];
    __shared__ float s_h[5][IH][TW + 1];
    __shared__ float s_part[8];
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    for (int i = threadIdx.x; i < IH * IW; i += 256) {
        const int r = i / IW, c = i % IW, gy = y0 + r - R, gx = x0 + c - R;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        s_x[r][c] = in ? pred[plane + (size_t)gy * W + gx] : 0.f;
        s_y[r][c] = in ? target[plane + (size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * TW; i += 256) {       // horizontal pass over all rows of the halo
        const int r = i / TW, c = i % TW;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float x = s_x[r][c + k], y = s_y[r][c + k], w = win.w[k];
            a += w * x;
            b += w * y;
            aa += w * (x * x);
            bb += w * (y * y);
            ab += w * (x * y);
        }
        s_h[0][r][c] = a;
        s_h[1][r][c] = b;
        s_h[2][r][c] = aa;
        s_h[3][r][c] = bb;
        s_h[4][r][c] = ab;
    }
    __syncthreads();
    float acc = 0.f;
    for (int i = threadIdx.x; i < TH * TW; i += 256) {       // vertical pass + SSIM map
        const int r = i / TW, c = i % TW, gy = y0 + r, gx = x0 + c;
        if (gy >= H || gx >= W) continue;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float w = win.w[k];
            mu1 += w * s_h[0][r + k][c];
            mu2 += w * s_h[1][r + k][c];
            e11 += w * s_h[2][r + k][c];
            e22 += w * s_h[3][r + k][c];
            e12 += w * s_h[4][r + k][c];
        }
        const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
        const float A = 2.f * mu12 + c1, B = 2.f * s12 + c2, C = mu1_sq + mu2_sq + c1, D = s1 + s2 + c2;
        const float inv = 1.f / (C * D), val = A * B * inv;
        const size_t o = plane + (size_t)gy * W + gx;
        dmu1[o] = ssim_scale * (2.f * mu2 * (B - A) * inv - 2.f * mu1 * val * (D - C) * inv);
        dE11[o] = ssim_scale * (-val / D);
        dE12[o] = ssim_scale * (2.f * A * inv);
        acc += ssim_scale * val + l1_scale * fabsf(s_x[r + R][c + R] - s_y[r + R][c + R]);
    }
    const float tot = block_sum(acc, s_part);
    if (threadIdx.x == 0) atomicAdd(loss, tot);
}

__global__ void __launch_bounds__(256) ssim_backward_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                            int H, int W, Window win, float l1_scale,
                                                            const float* __restrict__ dmu1, const float* __restrict__ dE11,
                                                            const float* __restrict__ dE12, float* __restrict__ grad) {
    __shared__ float s_m[3][IH][IW];
    __shared__ float s_h[3][IH][TW + 1];
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    for (int i = threadIdx.x; i < IH * IW; i += 256) {
        const int r = i / IW, c = i % IW, gy = y0 + r - R, gx = x0 + c - R;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = plane + (size_t)gy * W + gx;
        s_m[0][r][c] = in ? dmu1[o] : 0.f;
        s_m[1][r][c] = in ? dE11[o] : 0.f;
        s_m[2][r][c] = in ? dE12[o] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * TW; i += 256) {
        const int r = i / TW, c = i % TW;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float w = win.w[k];
            a += w * s_m[0][r][c + k];
            b += w * s_m[1][r][c + k];
            d += w * s_m[2][r][c + k];
        }
        s_h[0][r][c] = a;
        s_h[1][r][c] = b;
        s_h[2][r][c] = d;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TH * TW; i += 256) {
        const int r = i / TW, c = i % TW, gy = y0 + r, gx = x0 + c;
        if (gy >= H || gx >= W) continue;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float w = win.w[k];
            a += w * s_h[0][r + k][c];
            b += w * s_h[1][r + k][c];
            d += w * s_h[2][r + k][c];
        }
        const size_t o = plane + (size_t)gy * W + gx;
        const float x = pred[o], y = target[o], df = x - y;
        const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);     // torch.abs has a zero subgradient at 0
        grad[o] = a + 2.f * x * b + y * d + l1_scale * sgn;
    }
}

}  // namespace ssim

int launch_l1_ssim(const float* pred, const float* target, int C, int H, int W, const float* window11, float l1_scale,
                   float ssim_scale, float* loss, float* scratch, float* grad, cudaStream_t stream) {
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    ssim::Window win;
    for (int k = 0; k < ssim::K; k++) win.w[k] = window11[k];
    const size_t n = (size_t)C * H * W;
    const dim3 grid((W + ssim::TW - 1) / ssim::TW, (H + ssim::TH - 1) / ssim::TH, C);
    ssim::ssim_forward_kernel<<<grid, 256, 0, stream>>>(pred, target, H, W, win, l1_scale, ssim_scale, loss, scratch,
                                                         scratch + n, scratch + 2 * n);
    HS_LAUNCH_OK(stream, false);
    if (grad != nullptr) {
        ssim::ssim_backward_kernel<<<grid, 256, 0, stream>>>(pred, target, H, W, win, l1_scale, scratch, scratch + n,
                                                              scratch + 2 * n, grad);
        HS_LAUNCH_OK(stream, false);
    }
    return 0;
}

}  // namespace hs

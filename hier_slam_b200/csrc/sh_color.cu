// Spherical-harmonics colour path: view-dependent RGB of every visible Gaussian from degree <= 3 SH coefficients,
// and its backward (coefficient gradients + the view-direction term that flows into dL/dmean).
// Reference behaviour: cuda_rasterizer/forward.cu:20-71 (computeColorFromSH, +0.5, clamp at 0, `clamped` flags),
// backward.cu:20-139 and auxiliary.h:107-117 (dnormvdv).  Hier-SLAM itself always passes precomputed colours
// (utils/slam_helpers.py:211), so this path exists for API completeness; it is a plain one-thread-per-Gaussian pass
// that runs only when `shs` is given.  One basis evaluation is shared by the forward and the backward.
#include "hs_common.cuh"

namespace hs {

// auxiliary.h:22-39
__device__ constexpr float kC0 = 0.28209479177387814f;
__device__ constexpr float kC1 = 0.4886025119029199f;
__device__ constexpr float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                     -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                     0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                     -0.5900435899266435f};

// b[k] = d colour / d sh[k] for unit direction (x, y, z); only the first (deg+1)^2 entries are written
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
    b[0] = kC0;
    if (deg > 0) {
        b[1] = -kC1 * y;
        b[2] = kC1 * z;
        b[3] = -kC1 * x;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = kC2[0] * xy;
            b[5] = kC2[1] * yz;
            b[6] = kC2[2] * (2.0f * zz - xx - yy);
            b[7] = kC2[3] * xz;
            b[8] = kC2[4] * (xx - yy);
            if (deg > 2) {
                b[9] = kC3[0] * y * (3.0f * xx - yy);
                b[10] = kC3[1] * xy * z;
                b[11] = kC3[2] * y * (4.0f * zz - xx - yy);
                b[12] = kC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = kC3[4] * x * (4.0f * zz - xx - yy);
                b[14] = kC3[5] * z * (xx - yy);
                b[15] = kC3[6] * x * (xx - 3.0f * yy);
            }
        }
    }
}

__global__ void __launch_bounds__(256) sh_forward_kernel(int P, int deg, int M, const float* __restrict__ means3D,
                                                         const float* __restrict__ campos, const float* __restrict__ shs,
                                                         const int* __restrict__ radii, float* __restrict__ rgb,
                                                         uint8_t* __restrict__ clamped) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= P) return;
    float3 c = {0.f, 0.f, 0.f};
    uint8_t cl = 0;
    if (radii[idx] > 0) {   // the reference evaluates colours only for Gaussians that survive the cull (forward.cu:176-243)
        float3 dir = {means3D[3 * idx] - campos[0], means3D[3 * idx + 1] - campos[1], means3D[3 * idx + 2] - campos[2]};
        const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        dir = {dir.x / len, dir.y / len, dir.z / len};
        float b[16];
        sh_basis(deg, dir.x, dir.y, dir.z, b);
        const float* sh = shs + (size_t)idx * M * 3;
        const int n = (deg + 1) * (deg + 1);
        for (int k = 0; k < n; k++) {
            c.x = fmaf(b[k], sh[3 * k], c.x);
            c.y = fmaf(b[k], sh[3 * k + 1], c.y);
            c.z = fmaf(b[k], sh[3 * k + 2], c.z);
        }
        c = {c.x + 0.5f, c.y + 0.5f, c.z + 0.5f};
        cl = (c.x < 0 ? 1 : 0) | (c.y < 0 ? 2 : 0) | (c.z < 0 ? 4 : 0);
        c = {fmaxf(c.x, 0.f), fmaxf(c.y, 0.f), fmaxf(c.z, 0.f)};
    }
    rgb[3 * idx] = c.x;
    rgb[3 * idx + 1] = c.y;
    rgb[3 * idx + 2] = c.z;
    clamped[idx] = cl;
}

__global__ void __launch_bounds__(256) sh_backward_kernel(int P, int deg, int M, const float* __restrict__ means3D,
                                                          const float* __restrict__ campos, const float* __restrict__ shs,
                                                          const int* __restrict__ radii, const uint8_t* __restrict__ clamped,
                                                          const float* __restrict__ dL_dcolors,
                                                          float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dsh,
                                                          const float* __restrict__ pose_points,
                                                          float* __restrict__ dL_dpose) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= P) return;
    float* out = dL_dsh + (size_t)idx * M * 3;
    if (!(radii[idx] > 0)) {
        for (int k = 0; k < 3 * M; k++) out[k] = 0.f;   // every row is written: the output needs no memset
        return;
    }
    const float3 d0 = {means3D[3 * idx] - campos[0], means3D[3 * idx + 1] - campos[1], means3D[3 * idx + 2] - campos[2]};
    const float len = sqrtf(d0.x * d0.x + d0.y * d0.y + d0.z * d0.z);
    const float x = d0.x / len, y = d0.y / len, z = d0.z / len;
    const uint8_t cl = clamped[idx];
    // PyTorch rule for the clamp: no gradient through a clamped channel (backward.cu:30-35)
    const float3 g = {(cl & 1) ? 0.f : dL_dcolors[3 * idx], (cl & 2) ? 0.f : dL_dcolors[3 * idx + 1],
                      (cl & 4) ? 0.f : dL_dcolors[3 * idx + 2]};
    float b[16];
    sh_basis(deg, x, y, z, b);
    const int n = (deg + 1) * (deg + 1);
    for (int k = 0; k < M; k++) {
        const float w = k < n ? b[k] : 0.f;
        out[3 * k] = w * g.x;
        out[3 * k + 1] = w * g.y;
        out[3 * k + 2] = w * g.z;
    }
    if (deg == 0) return;   // degree 0 is view independent
    // s[k] = <sh[k], dL/dRGB>: the colour gradient seen through coefficient k
    const float* sh = shs + (size_t)idx * M * 3;
    float s[16];
    for (int k = 0; k < n; k++) s[k] = sh[3 * k] * g.x + sh[3 * k + 1] * g.y + sh[3 * k + 2] * g.z;
    // d colour / d direction (backward.cu:58-123), contracted with dL/dRGB
    float gx = -kC1 * s[3], gy = -kC1 * s[1], gz = kC1 * s[2];
    if (deg > 1) {
        gx += kC2[0] * y * s[4] + kC2[2] * 2.f * -x * s[6] + kC2[3] * z * s[7] + kC2[4] * 2.f * x * s[8];
        gy += kC2[0] * x * s[4] + kC2[1] * z * s[5] + kC2[2] * 2.f * -y * s[6] + kC2[4] * 2.f * -y * s[8];
        gz += kC2[1] * y * s[5] + kC2[2] * 2.f * 2.f * z * s[6] + kC2[3] * x * s[7];
        if (deg > 2) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            gx += kC3[0] * s[9] * 3.f * 2.f * xy + kC3[1] * s[10] * yz + kC3[2] * s[11] * -2.f * xy +
                  kC3[3] * s[12] * -3.f * 2.f * xz + kC3[4] * s[13] * (-3.f * xx + 4.f * zz - yy) +
                  kC3[5] * s[14] * 2.f * xz + kC3[6] * s[15] * 3.f * (xx - yy);
            gy += kC3[0] * s[9] * 3.f * (xx - yy) + kC3[1] * s[10] * xz + kC3[2] * s[11] * (-3.f * yy + 4.f * zz - xx) +
                  kC3[3] * s[12] * -3.f * 2.f * yz + kC3[4] * s[13] * -2.f * xy + kC3[5] * s[14] * -2.f * yz +
                  kC3[6] * s[15] * -3.f * 2.f * xy;
            gz += kC3[1] * s[10] * xy + kC3[2] * s[11] * 4.f * 2.f * yz + kC3[3] * s[12] * 3.f * (2.f * zz - xx - yy) +
                  kC3[4] * s[13] * 4.f * 2.f * xz + kC3[5] * s[14] * (xx - yy);
        }
    }
    // through the normalisation of the direction (auxiliary.h:107-117)
    const float sum2 = d0.x * d0.x + d0.y * d0.y + d0.z * d0.z;
    const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    const float3 dm = {((sum2 - d0.x * d0.x) * gx - d0.y * d0.x * gy - d0.z * d0.x * gz) * inv32,
                       (-d0.x * d0.y * gx + (sum2 - d0.y * d0.y) * gy - d0.z * d0.y * gz) * inv32,
                       (-d0.x * d0.z * gx - d0.y * d0.z * gy + (sum2 - d0.z * d0.z) * gz) * inv32};
    dL_dmeans3D[3 * idx] += dm.x;
    dL_dmeans3D[3 * idx + 1] += dm.y;
    dL_dmeans3D[3 * idx + 2] += dm.z;
    if (dL_dpose != nullptr) {   // the view-direction term also reaches the camera pose (rare path: plain atomics)
        const float pw[4] = {pose_points[3 * idx], pose_points[3 * idx + 1], pose_points[3 * idx + 2], 1.f};
        const float dmv[3] = {dm.x, dm.y, dm.z};
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) atomicAdd(dL_dpose + 4 * r + c, dmv[r] * pw[c]);
    }
}

int launch_sh_forward(int P, int deg, int M, const float* means3D, const float* campos, const float* shs,
                      const int* radii, const GeomView& g, cudaStream_t stream, bool debug) {
    if (P <= 0) return 0;
    sh_forward_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, deg, M, means3D, campos, shs, radii, g.rgb, g.clamped);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

int launch_sh_backward(int P, int deg, int M, const float* means3D, const float* campos, const float* shs,
                       const int* radii, const GeomView& g, const float* dL_dcolors, float* dL_dmeans3D, float* dL_dsh,
                       const float* pose_points, float* dL_dpose, cudaStream_t stream, bool debug) {
    if (P <= 0) return 0;
    sh_backward_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, deg, M, means3D, campos, shs, radii, g.clamped,
                                                           dL_dcolors, dL_dmeans3D, dL_dsh,
                                                           dL_dpose != nullptr ? pose_points : nullptr, dL_dpose);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

}  // namespace hs

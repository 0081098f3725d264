// Spherical-harmonics colour path: view-dependent RGB of every visible Gaussian from degree <= 3 SH coefficients,
// and its backward (coefficient gradients + the view-direction term that flows into dL/dmean).
// Reference behaviour: cuda_rasterizer/forward.cu:20-71 (computeColorFromSH, +0.5, clamp at 0, `clamped` flags),
// backward.cu:20-139 and auxiliary.h:107-117 (dnormvdv); the backward here is derived from the basis polynomials
// (sh_basis_gradient) and the tangent-plane projector of the normalisation, not from the reference's expanded terms.  Hier-SLAM itself always passes precomputed colours
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

// j[k] = gradient of the k-th basis polynomial of sh_basis with respect to (x, y, z); entries 1 .. (deg+1)^2 - 1 are written
__device__ __forceinline__ void sh_basis_gradient(int deg, float x, float y, float z, float (*j)[3]) {
    auto set = [&](int k, float dx, float dy, float dz) { j[k][0] = dx; j[k][1] = dy; j[k][2] = dz; };
    if (deg < 1) return;
    set(1, 0.f, -kC1, 0.f);                     // -C1 y
    set(2, 0.f, 0.f, kC1);                      //  C1 z
    set(3, -kC1, 0.f, 0.f);                     // -C1 x
    if (deg < 2) return;
    set(4, kC2[0] * y, kC2[0] * x, 0.f);                                   // xy
    set(5, 0.f, kC2[1] * z, kC2[1] * y);                                   // yz
    set(6, -2.f * kC2[2] * x, -2.f * kC2[2] * y, 4.f * kC2[2] * z);        // 2zz - xx - yy
    set(7, kC2[3] * z, 0.f, kC2[3] * x);                                   // xz
    set(8, 2.f * kC2[4] * x, -2.f * kC2[4] * y, 0.f);                      // xx - yy
    if (deg < 3) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    set(9, 6.f * kC3[0] * xy, 3.f * kC3[0] * (xx - yy), 0.f);                                      // y (3xx - yy)
    set(10, kC3[1] * yz, kC3[1] * xz, kC3[1] * xy);                                                // xyz
    set(11, -2.f * kC3[2] * xy, kC3[2] * (4.f * zz - xx - 3.f * yy), 8.f * kC3[2] * yz);           // y (4zz - xx - yy)
    set(12, -6.f * kC3[3] * xz, -6.f * kC3[3] * yz, 3.f * kC3[3] * (2.f * zz - xx - yy));          // z (2zz - 3xx - 3yy)
    set(13, kC3[4] * (4.f * zz - 3.f * xx - yy), -2.f * kC3[4] * xy, 8.f * kC3[4] * xz);           // x (4zz - xx - yy)
    set(14, 2.f * kC3[5] * xz, -2.f * kC3[5] * yz, kC3[5] * (xx - yy));                            // z (xx - yy)
    set(15, 3.f * kC3[6] * (xx - yy), -6.f * kC3[6] * xy, 0.f);                                    // x (xx - 3yy)
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
    // View-direction term: colour = sum_k b_k(dir) sh_k, so dL/ddir = sum_k s_k grad b_k(dir) with
    // s_k = <sh_k, dL/dRGB> (the colour gradient seen through coefficient k) and grad b_k the gradient of the k-th real
    // SH polynomial (sh_basis_gradient, derived term by term from sh_basis).
    const float* sh = shs + (size_t)idx * M * 3;
    float jac[16][3];
    sh_basis_gradient(deg, x, y, z, jac);
    float3 gdir = {0.f, 0.f, 0.f};
    for (int k = 1; k < n; k++) {
        const float sk = sh[3 * k] * g.x + sh[3 * k + 1] * g.y + sh[3 * k + 2] * g.z;
        gdir.x = fmaf(sk, jac[k][0], gdir.x);
        gdir.y = fmaf(sk, jac[k][1], gdir.y);
        gdir.z = fmaf(sk, jac[k][2], gdir.z);
    }
    // dir = d0 / |d0|: the Jacobian is the tangent-plane projector (I - dir dir^T) / |d0|
    const float radial = x * gdir.x + y * gdir.y + z * gdir.z;
    const float il = 1.0f / len;
    const float3 dm = {(gdir.x - x * radial) * il, (gdir.y - y * radial) * il, (gdir.z - z * radial) * il};
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

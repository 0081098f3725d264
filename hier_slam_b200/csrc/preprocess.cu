// Per-Gaussian kernels: forward preprocess (EWA projection, near cull, tile-overlap count), the fused
// per-Gaussian backward, and markVisible.
//
// Parity contract: radii, tiles_touched, depths, means2D and conic_opacity must be BIT-EXACT with the
// reference build (they define the sort keys and tile lists).  The arithmetic below therefore keeps the
// reference's expression shapes (reference: cuda_rasterizer/forward.cu:74-256, auxiliary.h:41-164 and the
// GLM 0.9.9.9 mat3 operators it instantiates) so that nvcc's FMA contraction produces the same sequence
// of roundings; no fast-math anywhere.  What is new is the data movement: [P,3] inputs are staged through
// shared memory with 128-bit coalesced loads, the two 4x4 matrices are read once per block, and the
// cov3D array is never written (the backward recomputes it from scale/rotation).
#include "hs_common.cuh"

namespace hs {

// Column-major 3x3 (m[c][r]) with exactly the operator formulas of glm/detail/type_mat3x3.inl:486-519.
struct M3 {
    float m[3][3];
};
__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        r.m[c][0] = a.m[0][0] * b.m[c][0] + a.m[1][0] * b.m[c][1] + a.m[2][0] * b.m[c][2];
        r.m[c][1] = a.m[0][1] * b.m[c][0] + a.m[1][1] * b.m[c][1] + a.m[2][1] * b.m[c][2];
        r.m[c][2] = a.m[0][2] * b.m[c][0] + a.m[1][2] * b.m[c][1] + a.m[2][2] * b.m[c][2];
    }
    return r;
}
__device__ __forceinline__ M3 m3_transpose(const M3& a) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int q = 0; q < 3; q++) r.m[c][q] = a.m[q][c];
    return r;
}
// glm::mat3(x0..x8): column-major constructor
__device__ __forceinline__ M3 m3_make(float x0, float x1, float x2, float x3, float x4, float x5, float x6,
                                       float x7, float x8) {
    M3 r;
    r.m[0][0] = x0; r.m[0][1] = x1; r.m[0][2] = x2;
    r.m[1][0] = x3; r.m[1][1] = x4; r.m[1][2] = x5;
    r.m[2][0] = x6; r.m[2][1] = x7; r.m[2][2] = x8;
    return r;
}

__device__ __forceinline__ float3 xform4x3(const float3& p, const float* m) {  // auxiliary.h:58-66
    float3 t = {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
    return t;
}
__device__ __forceinline__ float4 xform4x4(const float3& p, const float* m) {  // auxiliary.h:68-77
    float4 t = {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14], m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]};
    return t;
}

// Rotation matrix of forward.cu:135-139 (glm column-major constructor order).  The six mixed products are the
// only place of the per-Gaussian pass where the PTX leaves a mul+add pair for ptxas to contract, and ptxas picks
// the product to keep rounded by operand order -- which NVVM canonicalises differently in different kernels.
// The roundings are therefore pinned with explicit intrinsics to what the reference's sm_100 SASS does
// (FMUL r*z, r*x, x*z, y*y, z*z; every other product fused), so that cov3D -> conic stay bit-exact.
__device__ __forceinline__ M3 quat_to_rot(float r, float x, float y, float z) {
    const float rz = __fmul_rn(r, z), rx = __fmul_rn(r, x), xz = __fmul_rn(x, z);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float xy_m_rz = __fmaf_rn(x, y, -rz), xy_p_rz = __fmaf_rn(x, y, rz);
    const float xz_p_ry = __fmaf_rn(r, y, xz), xz_m_ry = __fmaf_rn(-r, y, xz);
    const float yz_m_rx = __fmaf_rn(y, z, -rx), yz_p_rx = __fmaf_rn(y, z, rx);
    const float yy_p_zz = __fadd_rn(yy, zz), xx_p_zz = __fmaf_rn(x, x, zz), xx_p_yy = __fmaf_rn(x, x, yy);
    return m3_make(__fsub_rn(1.f, __fadd_rn(yy_p_zz, yy_p_zz)), __fadd_rn(xy_m_rz, xy_m_rz), __fadd_rn(xz_p_ry, xz_p_ry),
                   __fadd_rn(xy_p_rz, xy_p_rz), __fsub_rn(1.f, __fadd_rn(xx_p_zz, xx_p_zz)), __fadd_rn(yz_m_rx, yz_m_rx),
                   __fadd_rn(xz_m_ry, xz_m_ry), __fadd_rn(yz_p_rx, yz_p_rx), __fsub_rn(1.f, __fadd_rn(xx_p_yy, xx_p_yy)));
}

// forward.cu:118-152 — Sigma = (S R)^T (S R); the quaternion is used as given.
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot, float* cov3D) {
    M3 S = m3_make(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
    S.m[0][0] = mod * scale.x;
    S.m[1][1] = mod * scale.y;
    S.m[2][2] = mod * scale.z;
    float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    M3 R = quat_to_rot(r, x, y, z);
    M3 M = m3_mul(S, R);
    M3 Sigma = m3_mul(m3_transpose(M), M);
    cov3D[0] = Sigma.m[0][0];
    cov3D[1] = Sigma.m[0][1];
    cov3D[2] = Sigma.m[0][2];
    cov3D[3] = Sigma.m[1][1];
    cov3D[4] = Sigma.m[1][2];
    cov3D[5] = Sigma.m[2][2];
}

// forward.cu:74-113
__device__ __forceinline__ float3 cov2d_forward(const float3& mean, float focal_x, float focal_y, float tan_fovx,
                                                float tan_fovy, const float* cov3D, const float* viewmatrix) {
    float3 t = xform4x3(mean, viewmatrix);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;
    M3 J = m3_make(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z), 0.0f, focal_y / t.z,
                   -(focal_y * t.y) / (t.z * t.z), 0, 0, 0);
    M3 W = m3_make(viewmatrix[0], viewmatrix[4], viewmatrix[8], viewmatrix[1], viewmatrix[5], viewmatrix[9],
                   viewmatrix[2], viewmatrix[6], viewmatrix[10]);
    M3 T = m3_mul(W, J);
    M3 Vrk = m3_make(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
    cov.m[0][0] += 0.3f;
    cov.m[1][1] += 0.3f;
    return {float(cov.m[0][0]), float(cov.m[0][1]), float(cov.m[1][1])};
}

__device__ __forceinline__ float ndc2pix(float v, int S) {  // auxiliary.h:41-44 (FP64 on purpose)
    return ((v + 1.0) * S - 1.0) * 0.5;
}

__device__ __forceinline__ void get_rect(const float2 p, int max_radius, uint2& rect_min, uint2& rect_max,
                                         unsigned gx, unsigned gy) {  // auxiliary.h:46-56
    rect_min = {min(gx, (unsigned)max((int)0, (int)((p.x - max_radius) / HS_TILE_X))),
                min(gy, (unsigned)max((int)0, (int)((p.y - max_radius) / HS_TILE_Y)))};
    rect_max = {min(gx, (unsigned)max((int)0, (int)((p.x + max_radius + HS_TILE_X - 1) / HS_TILE_X))),
                min(gy, (unsigned)max((int)0, (int)((p.y + max_radius + HS_TILE_Y - 1) / HS_TILE_Y)))};
}

// Stage `n3` consecutive floats (a [n,3] slab) into shared memory with 128-bit loads when aligned.
__device__ __forceinline__ void stage_slab(float* dst, const float* __restrict__ src, int n3) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = n3 >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x)
            reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (int i = (n4 << 2) + threadIdx.x; i < n3; i += blockDim.x) dst[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < n3; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}

__global__ void __launch_bounds__(256) preprocess_kernel(int P, const float* __restrict__ means3D,
                                                         const float* __restrict__ scales,
                                                         const float4* __restrict__ rotations,
                                                         const float* __restrict__ opacities,
                                                         const float* __restrict__ cov3D_precomp, Camera cam,
                                                         int* __restrict__ radii, float* __restrict__ depths,
                                                         float2* __restrict__ means2D,
                                                         float4* __restrict__ conic_opacity,
                                                         uint32_t* __restrict__ tiles_touched,
                                                         uint32_t* __restrict__ tile_count) {
    __shared__ __align__(16) float s_means[256 * 3];
    __shared__ __align__(16) float s_scales[256 * 3];
    __shared__ float s_view[16];
    __shared__ float s_proj[16];
    const int base = blockIdx.x * 256;
    const int n = min(256, P - base);
    stage_slab(s_means, means3D + (size_t)base * 3, n * 3);
    if (cov3D_precomp == nullptr) stage_slab(s_scales, scales + (size_t)base * 3, n * 3);
    if (threadIdx.x < 16) s_view[threadIdx.x] = __ldg(cam.view + threadIdx.x);
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = __ldg(cam.proj + threadIdx.x - 16);
    __syncthreads();
    const int idx = base + threadIdx.x;

    int my_radii = 0;
    uint32_t my_tiles = 0;
    float my_depth = 0.f;
    float2 my_xy = {0.f, 0.f};
    float4 my_co = {0.f, 0.f, 0.f, 0.f};
    uint2 my_rmin = {0, 0}, my_rmax = {0, 0};
    do {
        if (idx >= P) break;
        const float3 p_orig = {s_means[3 * threadIdx.x], s_means[3 * threadIdx.x + 1], s_means[3 * threadIdx.x + 2]};
        // in_frustum (auxiliary.h:139-164): only the near-plane test is live
        float4 p_hom = xform4x4(p_orig, s_proj);
        float p_w = 1.0f / (p_hom.w + 0.0000001f);
        float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};
        float3 p_view = xform4x3(p_orig, s_view);
        if (p_view.z <= 0.2f) break;

        float cov3D[6];
        if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov3D[k] = __ldg(cov3D_precomp + (size_t)idx * 6 + k);
        } else {
            const float3 sc = {s_scales[3 * threadIdx.x], s_scales[3 * threadIdx.x + 1], s_scales[3 * threadIdx.x + 2]};
            cov3d_from_scale_rot(sc, cam.scale_modifier, __ldg(rotations + idx), cov3D);
        }
        float3 cov = cov2d_forward(p_orig, cam.focal_x, cam.focal_y, cam.tanfovx, cam.tanfovy, cov3D, s_view);

        float det = __fmaf_rn(cov.x, cov.z, -__fmul_rn(cov.y, cov.y));  // reference SASS: FMUL b*b ; FFMA a*c - bb
        if (det == 0.0f) break;
        float det_inv = 1.f / det;
        float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

        float mid = 0.5f * (cov.x + cov.z);
        float lambda1 = mid + sqrt(max(0.1f, __fmaf_rn(mid, mid, -det)));
        float lambda2 = mid - sqrt(max(0.1f, __fmaf_rn(mid, mid, -det)));
        float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
        float2 point_image = {ndc2pix(p_proj.x, cam.W), ndc2pix(p_proj.y, cam.H)};
        uint2 rect_min, rect_max;
        get_rect(point_image, my_radius, rect_min, rect_max, cam.grid_x, cam.grid_y);
        if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) break;

        my_depth = p_view.z;
        my_radii = my_radius;
        my_xy = point_image;
        my_co = {conic.x, conic.y, conic.z, __ldg(opacities + idx)};
        my_tiles = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
        my_rmin = rect_min;
        my_rmax = rect_max;
    } while (0);
    // Per-tile instance counts for the tile-bucket sort (binning.cu).  Rects of more than 4 tiles are expanded by the
    // whole warp.
    if (tile_count != nullptr) {
        const uint32_t w = my_rmax.x - my_rmin.x;
        if (my_tiles > 0 && my_tiles <= 4) {
            for (uint32_t n = 0; n < my_tiles; n++)
                atomicAdd(tile_count + (size_t)((my_rmin.y + n / w) * cam.grid_x + my_rmin.x + n % w) * HS_CTR_STRIDE, 1u);
        }
        unsigned big = __ballot_sync(0xffffffffu, my_tiles > 4);
        const int lane = threadIdx.x & 31;
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t s_w = __shfl_sync(0xffffffffu, w, src), s_n = __shfl_sync(0xffffffffu, my_tiles, src);
            const uint32_t s_x0 = __shfl_sync(0xffffffffu, my_rmin.x, src), s_y0 = __shfl_sync(0xffffffffu, my_rmin.y, src);
            for (uint32_t n = lane; n < s_n; n += 32)
                atomicAdd(tile_count + (size_t)((s_y0 + n / s_w) * cam.grid_x + s_x0 + n % s_w) * HS_CTR_STRIDE, 1u);
        }
    }
    if (idx >= P) return;
    // Unlike the reference (which leaves stale bytes for culled Gaussians) every field is written, so the
    // state buffers never need a memset and can come from torch.empty.
    radii[idx] = my_radii;
    tiles_touched[idx] = my_tiles;
    depths[idx] = my_depth;
    means2D[idx] = my_xy;
    conic_opacity[idx] = my_co;
}

int launch_preprocess(int P, const float* means3D, const float* scales, const float* rotations,
                      const float* opacities, const float* cov3D_precomp, const Camera& cam, int* radii,
                      const GeomView& g, uint32_t* tile_count, cudaStream_t stream, bool debug) {
    if (P <= 0) return 0;
    prof_begin(ST_PREPROCESS, stream);
    preprocess_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, scales, (const float4*)rotations, opacities,
                                                          cov3D_precomp, cam, radii, g.depths, g.means2D,
                                                          g.conic_opacity, g.tiles_touched,
                                                          tile_count);
    prof_end(ST_PREPROCESS, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// markVisible (reference: rasterizer_impl.cu:54-66,141-153)
// ------------------------------------------------------------------------------------------------------
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    bool* __restrict__ present) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    float3 pv = xform4x3(p, view);
    present[idx] = !(pv.z <= 0.2f);
}
int launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, bool* present,
                        cudaStream_t stream, bool debug) {
    (void)proj;
    if (P <= 0) return 0;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, view, present);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// Per-Gaussian backward (reference behaviour: cuda_rasterizer/backward.cu:144-274, 278-341, 346-412 -- three kernels
// there, one pass here).  Every output row is written (zeros for culled Gaussians), so the outputs need no memset.
//
// Written from the math in vector form, not from the reference's expanded scalar expressions:
//   A      = J2 Rw            2x3, rows a0, a1: the affine map camera-frame covariance -> pixel covariance,
//                             J2 = [[fx/tz, 0, -fx cx/tz], [0, fy/tz, -fy cy/tz]], cx = clamp(tx/tz), cy = clamp(ty/tz)
//   cov2D  = A Sigma A^T + 0.3 I = [[a, b], [b, c]],  conic = adj(cov2D) / det
//   H      = dL/dcov2D (symmetric 2x2) = -adj G adj^T / (det^2 + 1e-7),  G = [[g.x, g.y], [g.y, g.w]] = dL/dconic
//   N      = dL/dSigma (symmetric 3x3) = A^T H A            -> dL_dcov3D = (N00, 2 N01, 2 N02, N11, 2 N12, N22)
//   D      = dL/dA = 2 H (A Sigma)                          -> dL/dJ2 = D Rw^T -> dL/dt -> dL/dmean = Rw^T dL/dt
//   Sigma  = Q diag(s^2) Q^T, Q = rotation of the (unnormalised) quaternion, s = modifier * scale:
//            dL/ds_k = 2 s_k q_k^T N q_k,  dL/dQ = E with E[:,k] = 2 s_k^2 N q_k,  dL/dquat from E's symmetric / antisymmetric parts.
// Like the reference, the clamped tx, ty are treated as independent variables (their gradient is masked where the clamp
// was active) and no normalisation backward is applied to the quaternion.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot3(const float3 a, const float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float3 comb2(const float s, const float3 a, const float t, const float3 b) {   // s a + t b
    return {fmaf(t, b.x, s * a.x), fmaf(t, b.y, s * a.y), fmaf(t, b.z, s * a.z)};
}
__device__ __forceinline__ float3 sym3_apply(const float* S6, const float3 v) {   // (xx xy xz yy yz zz) v
    return {fmaf(S6[2], v.z, fmaf(S6[1], v.y, S6[0] * v.x)), fmaf(S6[4], v.z, fmaf(S6[3], v.y, S6[1] * v.x)),
            fmaf(S6[5], v.z, fmaf(S6[4], v.y, S6[2] * v.x))};
}

__global__ void __launch_bounds__(256) geom_backward_kernel(
    int P, const float* __restrict__ means3D, const int* __restrict__ radii, const float* __restrict__ scales,
    const float4* __restrict__ rotations, const float* __restrict__ cov3D_precomp, Camera cam,
    const float* __restrict__ dL_dmean2D, const float4* __restrict__ dL_dconic, const float* __restrict__ dL_ddepths,
    float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dcov3D, float* __restrict__ dL_dscales,
    float4* __restrict__ dL_drots, const float* __restrict__ pose_points, float* __restrict__ dL_dpose) {
    __shared__ float s_view[16];
    __shared__ float s_proj[16];
    __shared__ float s_pose[8][12];
    if (threadIdx.x < 16) s_view[threadIdx.x] = __ldg(cam.view + threadIdx.x);
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = __ldg(cam.proj + threadIdx.x - 16);
    __syncthreads();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const bool valid = idx < P;
    const bool has_scales = (cov3D_precomp == nullptr);
    float3 g_mean = {0.f, 0.f, 0.f};
    float g_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float3 g_scale = {0.f, 0.f, 0.f};
    float4 g_rot = {0.f, 0.f, 0.f, 0.f};
    if (valid && radii[idx] > 0) {
        const float3 mean = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
        float Sig[6];
        float3 sc = {0.f, 0.f, 0.f};
        float4 quat = {0.f, 0.f, 0.f, 0.f};
        if (has_scales) {
            sc = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
            quat = __ldg(rotations + idx);
            cov3d_from_scale_rot(sc, cam.scale_modifier, quat, Sig);      // the forward's own function: same Sigma
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) Sig[k] = __ldg(cov3D_precomp + (size_t)idx * 6 + k);
        }
        // rows of the world-to-camera rotation (s_view is the transposed matrix, flat index 4 c + r)
        const float3 w0 = {s_view[0], s_view[4], s_view[8]}, w1 = {s_view[1], s_view[5], s_view[9]},
                     w2 = {s_view[2], s_view[6], s_view[10]};
        const float3 t = xform4x3(mean, s_view);
        const float iz = 1.0f / t.z;
        const float limx = 1.3f * cam.tanfovx, limy = 1.3f * cam.tanfovy;
        const float ux = t.x * iz, uy = t.y * iz;
        const float cx = fminf(limx, fmaxf(-limx, ux)), cy = fminf(limy, fmaxf(-limy, uy));
        const float keep_x = (ux < -limx || ux > limx) ? 0.f : 1.f, keep_y = (uy < -limy || uy > limy) ? 0.f : 1.f;
        const float fxz = cam.focal_x * iz, fyz = cam.focal_y * iz;
        const float3 a0 = comb2(fxz, w0, -fxz * cx, w2), a1 = comb2(fyz, w1, -fyz * cy, w2);
        const float3 Sa0 = sym3_apply(Sig, a0), Sa1 = sym3_apply(Sig, a1);
        const float ca = dot3(a0, Sa0) + 0.3f, cb = dot3(a0, Sa1), cc = dot3(a1, Sa1) + 0.3f;
        const float det = ca * cc - cb * cb;
        const float w = 1.0f / (det * det + 0.0000001f);
        // H = -w adj G adj^T with adj = [[cc, -cb], [-cb, ca]]
        const float4 gc = __ldg(dL_dconic + idx);     // (d/dconic.x, d/dconic.y, unused, d/dconic.z)
        float H00 = 0.f, H01 = 0.f, H11 = 0.f;
        if (w != 0.f) {
            const float gu0 = gc.x * cc - gc.y * cb, gu1 = gc.y * cc - gc.w * cb;     // G (cc, -cb)^T
            const float gv0 = gc.y * ca - gc.x * cb, gv1 = gc.w * ca - gc.y * cb;     // G (-cb, ca)^T
            H00 = -w * (cc * gu0 - cb * gu1);
            H01 = -w * (cc * gv0 - cb * gv1);
            H11 = -w * (ca * gv1 - cb * gv0);
        }
        // N = A^T H A = a0 (x) y0 + a1 (x) y1 with y0 = H00 a0 + H01 a1, y1 = H01 a0 + H11 a1
        const float3 y0 = comb2(H00, a0, H01, a1), y1 = comb2(H01, a0, H11, a1);
        float N[6];
        N[0] = fmaf(a1.x, y1.x, a0.x * y0.x);
        N[1] = fmaf(a1.x, y1.y, a0.x * y0.y);
        N[2] = fmaf(a1.x, y1.z, a0.x * y0.z);
        N[3] = fmaf(a1.y, y1.y, a0.y * y0.y);
        N[4] = fmaf(a1.y, y1.z, a0.y * y0.z);
        N[5] = fmaf(a1.z, y1.z, a0.z * y0.z);
        g_cov[0] = N[0];
        g_cov[1] = 2.f * N[1];
        g_cov[2] = 2.f * N[2];
        g_cov[3] = N[3];
        g_cov[4] = 2.f * N[4];
        g_cov[5] = N[5];
        // D = 2 H (A Sigma); only four entries of dL/dJ2 = D Rw^T are live
        const float3 d0 = comb2(2.f * H00, Sa0, 2.f * H01, Sa1), d1 = comb2(2.f * H01, Sa0, 2.f * H11, Sa1);
        const float e00 = dot3(d0, w0), e02 = dot3(d0, w2), e11 = dot3(d1, w1), e12 = dot3(d1, w2);
        const float iz2 = iz * iz;
        const float gtx = -cam.focal_x * iz2 * e02 * keep_x;
        const float gty = -cam.focal_y * iz2 * e12 * keep_y;
        const float gtz = iz2 * (2.f * (cam.focal_x * cx * e02 + cam.focal_y * cy * e12) - (cam.focal_x * e00 + cam.focal_y * e11));
        g_mean = {fmaf(w2.x, gtz, fmaf(w1.x, gty, w0.x * gtx)), fmaf(w2.y, gtz, fmaf(w1.y, gty, w0.y * gtx)),
                  fmaf(w2.z, gtz, fmaf(w1.z, gty, w0.z * gtx))};

        // ---- pixel-centre path: ndc = hom.xy / (hom.w + eps); the incoming gradient is d/d(ndc) (the blend backward has
        //      already applied the W/2, H/2 of the pixel mapping).  dL/dmean = P^T (gx iw, gy iw, 0, -(gx ndc.x + gy ndc.y) iw)
        {
            const float4 hom = xform4x4(mean, s_proj);
            const float iw = 1.0f / (hom.w + 0.0000001f);
            const float gx = dL_dmean2D[3 * idx], gy = dL_dmean2D[3 * idx + 1];
            const float kx = gx * iw, ky = gy * iw;
            const float kw = -(kx * hom.x + ky * hom.y) * iw;
            g_mean.x += fmaf(s_proj[3], kw, fmaf(s_proj[1], ky, s_proj[0] * kx));
            g_mean.y += fmaf(s_proj[7], kw, fmaf(s_proj[5], ky, s_proj[4] * kx));
            g_mean.z += fmaf(s_proj[11], kw, fmaf(s_proj[9], ky, s_proj[8] * kx));
        }
        // ---- depth path: depth = third row of the view transform (its homogeneous row is kept, as the reference does)
        {
            const float gd = dL_ddepths[idx];
            g_mean.x += (s_view[2] - s_view[3] * t.z) * gd;
            g_mean.y += (s_view[6] - s_view[7] * t.z) * gd;
            g_mean.z += (s_view[10] - s_view[11] * t.z) * gd;
        }

        // ---- Sigma = Q diag(s^2) Q^T  ->  scale and quaternion
        if (has_scales) {
            const float r = quat.x, x = quat.y, y = quat.z, z = quat.w;
            const M3 Qt = quat_to_rot(r, x, y, z);       // stored column-major as the TRANSPOSED rotation: Qt.m[k] = column k of Q
            const float3 s = {cam.scale_modifier * sc.x, cam.scale_modifier * sc.y, cam.scale_modifier * sc.z};
            const float sk[3] = {s.x, s.y, s.z};
            float E[3][3];                               // E[i][k] = dL/dQ[i][k]
            float gs[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float3 qk = {Qt.m[0][k], Qt.m[1][k], Qt.m[2][k]};
                const float3 nq = sym3_apply(N, qk);
                gs[k] = 2.f * sk[k] * dot3(qk, nq);
                const float f = 2.f * sk[k] * sk[k];
                E[0][k] = f * nq.x;
                E[1][k] = f * nq.y;
                E[2][k] = f * nq.z;
            }
            g_scale = {gs[0], gs[1], gs[2]};
            // antisymmetric part (axial vector) and symmetric sums of E
            const float ax = E[2][1] - E[1][2], ay = E[0][2] - E[2][0], az = E[1][0] - E[0][1];
            const float sxy = E[0][1] + E[1][0], sxz = E[0][2] + E[2][0], syz = E[1][2] + E[2][1];
            g_rot.x = 2.f * (x * ax + y * ay + z * az);
            g_rot.y = 2.f * (r * ax + y * sxy + z * sxz) - 4.f * x * (E[1][1] + E[2][2]);
            g_rot.z = 2.f * (r * ay + x * sxy + z * syz) - 4.f * y * (E[0][0] + E[2][2]);
            g_rot.w = 2.f * (r * az + x * sxz + y * syz) - 4.f * z * (E[0][0] + E[1][1]);
        }
    }
    if (valid) {
        dL_dmeans3D[3 * idx] = g_mean.x;
        dL_dmeans3D[3 * idx + 1] = g_mean.y;
        dL_dmeans3D[3 * idx + 2] = g_mean.z;
#pragma unroll
        for (int k = 0; k < 6; k++) dL_dcov3D[(size_t)idx * 6 + k] = g_cov[k];
        if (dL_dscales != nullptr) {
            dL_dscales[3 * idx] = g_scale.x;
            dL_dscales[3 * idx + 1] = g_scale.y;
            dL_dscales[3 * idx + 2] = g_scale.z;
            dL_drots[idx] = g_rot;
        }
    }
    // Camera-pose gradient, pre-reduced here instead of by a [P,3]x[P,4] autograd matmul in the caller: when the means
    // handed to the rasterizer are W [p_world; 1] (utils/slam_helpers.py:318-321), dL/dW[:3,:] = sum_i dL/dmean_i (x)
    // [p_world_i, 1].  Warp shuffle -> shared memory -> 12 atomics per block (kernel-uniform branch).
    if (dL_dpose != nullptr) {
        float3 pw = {0.f, 0.f, 0.f};
        if (valid) pw = {pose_points[3 * idx], pose_points[3 * idx + 1], pose_points[3 * idx + 2]};
        float v[12] = {g_mean.x * pw.x, g_mean.x * pw.y, g_mean.x * pw.z, g_mean.x,
                       g_mean.y * pw.x, g_mean.y * pw.y, g_mean.y * pw.z, g_mean.y,
                       g_mean.z * pw.x, g_mean.z * pw.y, g_mean.z * pw.z, g_mean.z};
#pragma unroll
        for (int k = 0; k < 12; k++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++) s_pose[threadIdx.x >> 5][k] = v[k];
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) t += s_pose[w][threadIdx.x];
            if (t != 0.f) atomicAdd(dL_dpose + threadIdx.x, t);
        }
    }
}

int launch_geom_backward(int P, const float* means3D, const int* radii, const float* scales,
                         const float* rotations, const float* cov3D_precomp, const Camera& cam,
                         const float* dL_dmean2D, const float* dL_dconic, const float* dL_ddepths,
                         float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales, float* dL_drots,
                         const float* pose_points, float* dL_dpose, cudaStream_t stream, bool debug) {
    if (P <= 0) return 0;
    prof_begin(ST_GEOM_BWD, stream);
    geom_backward_kernel<<<(P + 255) / 256, 256, 0, stream>>>(
        P, means3D, radii, scales, (const float4*)rotations, cov3D_precomp, cam, dL_dmean2D,
        (const float4*)dL_dconic, dL_ddepths, dL_dmeans3D, dL_dcov3D, dL_dscales, (float4*)dL_drots,
        dL_dpose != nullptr ? pose_points : nullptr, dL_dpose);
    prof_end(ST_GEOM_BWD, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

}  // namespace hs

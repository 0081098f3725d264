// The small steps of a tracking iteration that sit between the rasterizer calls.  In torch
// they are ~200 tiny kernels per iteration (pose matrix from the quaternion: ~25; camera-frame means: a skinny sgemm;
// masks + two masked L1 sums: ~15; autograd's adds; Adam on two tiny tensors: ~30; best-candidate bookkeeping: ~8) --
// under a CUDA graph still ~0.3 ms of a 1.28 ms iteration.  Here they are three kernels:
//   transform_points_kernel   means_cam = R means_world + t                      (utils/slam_helpers.py:278-330)
//   tracking_loss_kernel      mask, loss = w_d sum|gt_d - d|[mask] + w_im sum|gt_im - im|[mask], dL/dim, dL/ddepth
//                                                                                (scripts/hierslam.py:765-796, tracking=True)
//   pose_step_kernel          best-candidate bookkeeping (:1850-1856), dL/d(rel_w2c) -> dL/d(unnormalised quaternion,
//                             translation) through build_rotation + F.normalize, torch.optim.Adam's update on the seven
//                             numbers, and the rel_w2c of the next iteration.  One thread.
#include "hs_common.cuh"

namespace hs {

__global__ void __launch_bounds__(256) transform_points_kernel(const float* __restrict__ w2c, const float* __restrict__ world,
                                                               int P, float* __restrict__ cam) {
    __shared__ float m[12];
    if (threadIdx.x < 12) m[threadIdx.x] = w2c[threadIdx.x];      // rows 0..2 of the row-major 4x4
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float x = world[3 * i], y = world[3 * i + 1], z = world[3 * i + 2];
    // torch.addmm(t, X, R^T): each entry is t + x R_k0 + y R_k1 + z R_k2 accumulated in that order
    cam[3 * i] = fmaf(z, m[2], fmaf(y, m[1], fmaf(x, m[0], m[3])));
    cam[3 * i + 1] = fmaf(z, m[6], fmaf(y, m[5], fmaf(x, m[4], m[7])));
    cam[3 * i + 2] = fmaf(z, m[10], fmaf(y, m[9], fmaf(x, m[8], m[11])));
}

int launch_transform_points(const float* w2c, const float* world, int P, float* cam, cudaStream_t stream) {
    if (P <= 0) return 0;
    transform_points_kernel<<<(P + 255) / 256, 256, 0, stream>>>(w2c, world, P, cam);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

__global__ void __launch_bounds__(256) tracking_loss_kernel(const float* __restrict__ im, const float* __restrict__ depth,
                                                            const float* __restrict__ sil, const float* __restrict__ gt_im,
                                                            const float* __restrict__ gt_depth, size_t HW, float sil_thres,
                                                            int use_sil, float w_depth, float w_im, float* __restrict__ loss,
                                                            float* __restrict__ grad_im, float* __restrict__ grad_depth) {
    __shared__ float s_part[8];
    float acc = 0.f;
    for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < HW; p += (size_t)gridDim.x * 256) {
        const float d = depth[p], gd = gt_depth[p];
        const bool m = gd > 0.f && !isnan(d) && (!use_sil || sil[p] > sil_thres);
        // colour term: masked like the depth only when the silhouette mask is in use; otherwise the reference sums
        // |gt_im - im| over EVERY pixel (scripts/hierslam.py:789-794 with ignore_outlier_depth_loss = False)
        const bool mc = use_sil ? m : true;
        const float dd = d - gd;
        acc += m ? w_depth * fabsf(dd) : 0.f;
        grad_depth[p] = m ? w_depth * (dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f)) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float di = im[c * HW + p] - gt_im[c * HW + p];
            acc += mc ? w_im * fabsf(di) : 0.f;
            grad_im[c * HW + p] = mc ? w_im * (di > 0.f ? 1.f : (di < 0.f ? -1.f : 0.f)) : 0.f;
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

int launch_tracking_loss(const float* im, const float* depth, const float* sil, const float* gt_im, const float* gt_depth,
                         size_t HW, float sil_thres, int use_sil, float w_depth, float w_im, float* loss, float* grad_im,
                         float* grad_depth, cudaStream_t stream) {
    if (HW == 0) return 0;
    int blocks = (int)((HW + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    tracking_loss_kernel<<<blocks, 256, 0, stream>>>(im, depth, sil, gt_im, gt_depth, HW, sil_thres, use_sil, w_depth, w_im,
                                                     loss, grad_im, grad_depth);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

__device__ __forceinline__ void write_pose_matrix(const float* q, const float* t, float* w2c) {
    const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);   // F.normalize eps
    const float r = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    w2c[0] = 1.f - 2.f * (y * y + z * z);
    w2c[1] = 2.f * (x * y - r * z);
    w2c[2] = 2.f * (x * z + r * y);
    w2c[3] = t[0];
    w2c[4] = 2.f * (x * y + r * z);
    w2c[5] = 1.f - 2.f * (x * x + z * z);
    w2c[6] = 2.f * (y * z - r * x);
    w2c[7] = t[1];
    w2c[8] = 2.f * (x * z - r * y);
    w2c[9] = 2.f * (y * z + r * x);
    w2c[10] = 1.f - 2.f * (x * x + y * y);
    w2c[11] = t[2];
    w2c[12] = w2c[13] = w2c[14] = 0.f;
    w2c[15] = 1.f;
}

// state (HS_POSE_STATE_FLOATS floats): exp_avg rot[4] tran[3] | exp_avg_sq rot[4] tran[3] | step | min_loss |
// cand_rot[4] | cand_tran[3] | last_loss | overflow (sticky) | max num_rendered | max longest tile list | reserved
__global__ void pose_step_kernel(float* __restrict__ cam_rot, float* __restrict__ cam_tran, const float* __restrict__ dL_dpose,
                                 float* __restrict__ loss, float* __restrict__ state, float* __restrict__ w2c,
                                 const uint32_t* __restrict__ binning_info, float lr_rot, float lr_tran, float beta1,
                                 float beta2, float eps, int mode) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float q[4] = {cam_rot[0], cam_rot[1], cam_rot[2], cam_rot[3]}, t[3] = {cam_tran[0], cam_tran[1], cam_tran[2]};
    if (mode == 0) {               // frame start: only the matrix of the current pose
        write_pose_matrix(q, t, w2c);
        return;
    }
    float* m = state;
    float* v = state + 7;
    const float l = loss[0];
    loss[0] = 0.f;                 // ready for the next iteration's accumulation
    // Capacity-mode binning: an iteration that outgrew the capacity rendered EMPTY (loss 0, gradients 0).  The binning
    // flag is rewritten by every forward, so it is latched here for the whole frame, and the iteration neither becomes a
    // candidate nor moves the pose or the optimiser state.
    if (binning_info != nullptr) {   // largest counts any iteration of the frame needed: sizes the re-capture
        state[25] = fmaxf(state[25], (float)binning_info[0]);
        state[26] = fmaxf(state[26], (float)binning_info[1]);
        if (binning_info[3] != 0u) {
            state[24] = 1.f;
            return;
        }
    }
    state[23] = l;
    // dL/dR (rows of dL_dpose[3,4]) -> normalised quaternion (r, x, y, z) -> unnormalised quaternion
    const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
    const float r = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    const float G00 = dL_dpose[0], G01 = dL_dpose[1], G02 = dL_dpose[2], G10 = dL_dpose[4], G11 = dL_dpose[5],
                G12 = dL_dpose[6], G20 = dL_dpose[8], G21 = dL_dpose[9], G22 = dL_dpose[10];
    float gq[4];
    gq[0] = 2.f * (-z * G01 + y * G02 + z * G10 - x * G12 - y * G20 + x * G21);
    gq[1] = 2.f * (y * G01 + z * G02 + y * G10 - 2.f * x * G11 - r * G12 + z * G20 + r * G21 - 2.f * x * G22);
    gq[2] = 2.f * (-2.f * y * G00 + x * G01 + r * G02 + x * G10 + z * G12 - r * G20 + z * G21 - 2.f * y * G22);
    gq[3] = 2.f * (-2.f * z * G00 - r * G01 + x * G02 + r * G10 - 2.f * z * G11 + y * G12 + x * G20 + y * G21);
    const float dot = r * gq[0] + x * gq[1] + y * gq[2] + z * gq[3];
    const float qh[4] = {r, x, y, z};
    float g[7];
#pragma unroll
    for (int k = 0; k < 4; k++) g[k] = (gq[k] - qh[k] * dot) / n;
    g[4] = dL_dpose[3];
    g[5] = dL_dpose[7];
    g[6] = dL_dpose[11];
    // torch.optim.Adam (capturable form: the step counter lives on the device)
    const float step = state[14] + 1.f;
    state[14] = step;
    const float bc1 = 1.f - powf(beta1, step), bc2_sqrt = sqrtf(1.f - powf(beta2, step));
#pragma unroll
    for (int k = 0; k < 7; k++) {
        m[k] = m[k] + (1.f - beta1) * (g[k] - m[k]);
        v[k] = v[k] * beta2 + (1.f - beta2) * g[k] * g[k];
        const float denom = sqrtf(v[k]) / bc2_sqrt + eps;
        const float upd = ((k < 4 ? lr_rot : lr_tran) / bc1) * (m[k] / denom);
        if (k < 4) q[k] -= upd; else t[k - 4] -= upd;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) cam_rot[k] = q[k];
#pragma unroll
    for (int k = 0; k < 3; k++) cam_tran[k] = t[k];
    // Best candidate exactly as the reference keeps it (scripts/hierslam.py:1851-1858): loss.backward(), optimizer.step(),
    // THEN `if loss < current_min_loss` saves cam_unnorm_rots / cam_trans -- i.e. the pose AFTER the step, paired with the
    // loss evaluated before it.
    if (l < state[15]) {
        state[15] = l;
#pragma unroll
        for (int k = 0; k < 4; k++) state[16 + k] = q[k];
#pragma unroll
        for (int k = 0; k < 3; k++) state[20 + k] = t[k];
    }
    write_pose_matrix(q, t, w2c);
}

int launch_pose_step(float* cam_rot, float* cam_tran, const float* dL_dpose, float* loss, float* state, float* w2c,
                     const uint32_t* binning_info, float lr_rot, float lr_tran, float beta1, float beta2, float eps, int mode,
                     cudaStream_t stream) {
    pose_step_kernel<<<1, 32, 0, stream>>>(cam_rot, cam_tran, dL_dpose, loss, state, w2c, binning_info, lr_rot, lr_tran,
                                           beta1, beta2, eps, mode);
    HS_LAUNCH_OK(stream, false);
    return 0;
}


// Keyframe selection (utils/keyframe_selection.py:40-96): for every keyframe, how many of the sampled world points project
// inside its image (minus an edge margin) with positive depth.  The reference loops over the keyframes in Python (~10 torch
// kernels and one host sync each); here one CTA per keyframe counts all points.
__global__ void __launch_bounds__(256) keyframe_overlap_kernel(const float* __restrict__ pts, int N,
                                                               const float* __restrict__ w2c, float fx, float fy, float cx,
                                                               float cy, float width, float height, float edge,
                                                               int* __restrict__ counts) {
    __shared__ float m[12];
    __shared__ int s_part[8];
    if (threadIdx.x < 12) m[threadIdx.x] = w2c[(size_t)blockIdx.x * 16 + threadIdx.x];
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        const float X = fmaf(z, m[2], fmaf(y, m[1], fmaf(x, m[0], m[3])));
        const float Y = fmaf(z, m[6], fmaf(y, m[5], fmaf(x, m[4], m[7])));
        const float Z = fmaf(z, m[10], fmaf(y, m[9], fmaf(x, m[8], m[11])));
        const float pz = Z + 1e-5f;                                  // points_z = points_2d[:, 2:] + 1e-5
        const float u = fmaf(cx, Z, fx * X) / pz, v = fmaf(cy, Z, fy * Y) / pz;
        c += (u < width - edge && u > edge && v < height - edge && v > edge && pz > 0.f) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_part[w];
        counts[blockIdx.x] = t;
    }
}

int launch_keyframe_overlap(const float* pts, int N, const float* w2c, int K, float fx, float fy, float cx, float cy, int width,
                            int height, int edge, int* counts, cudaStream_t stream) {
    if (K <= 0) return 0;
    keyframe_overlap_kernel<<<K, 256, 0, stream>>>(pts, N, w2c, fx, fy, cx, cy, (float)width, (float)height, (float)edge,
                                                   counts);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

}  // namespace hs

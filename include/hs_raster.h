/* libhsraster — C ABI of the Blackwell-native differentiable Gaussian rasterizer (Hier-SLAM render hot path).
 *
 * This is the drop-in boundary for the reference's PyTorch extension `diff_gaussian_rasterization._C`
 * (reference: hierslam-diff-gaussian-rasterization-w-depth/ext.cpp:15-23, rasterize_points.h:18-125).
 * Everything is plain pointers and sizes: all `const float*` / `float*` arguments are DEVICE pointers to
 * contiguous float32 arrays on the current CUDA device unless stated otherwise; `stream` is a cudaStream_t
 * passed as void*.  No torch types, no CPU fallback.  Every function returns 0 on success, non-zero on error
 * (then hs_last_error() describes it).  With cam->debug != 0 every kernel launch is followed by a stream
 * synchronisation and an error check (reference: CHECK_CUDA, cuda_rasterizer/auxiliary.h:166-173).
 *
 * The caller owns all memory (in the Python host: torch tensors, so the caching allocator and stream
 * semantics of torch apply).  The three opaque state buffers play the role of the reference's geomBuffer /
 * binningBuffer / imgBuffer (rasterize_points.cu:285-291): they are produced by the forward call and must be
 * handed unchanged to hs_backward.
 *
 * Mapping to the reference entry points:
 *   rasterize_gaussians_semantic  (rasterize_points.cu:240-336)  -> hs_forward_geometry + hs_forward_render (S > 0)
 *   rasterize_gaussians           (rasterize_points.cu:35-117)   -> same with S = 0 and out_mask != NULL
 *   rasterize_gaussians_backward_semantic (:339-432) / rasterize_gaussians_backward (:119-215) -> hs_backward
 *   mark_visible                  (rasterize_points.cu:217-236)  -> hs_mark_visible
 * The forward is split in two calls because the size of the binning buffer (num_rendered tile instances) is
 * only known after the per-Gaussian pass; the reference resolves this with a resize callback into torch
 * (rasterize_points.cu:27-33), a plain C ABI resolves it by returning num_rendered to the caller.
 */
#ifndef HS_RASTER_H_
#define HS_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_RASTER_ABI_VERSION 5

/* flags (bit-or) */
#define HS_SEM_ALPHA_EXACT 1 /* backward: semantic channels contribute to dL/dalpha (the mathematically intended
                                gradient).  Default (0) reproduces the reference, whose kernel reads the semantic
                                feature from a never-written scratch buffer (backward.cu:834, rasterizer_impl.cu:673). */
#define HS_BWD_SIMT 4        /* backward: use the SIMT (warp-shuffle) blend backward instead of the tensor-core one */
#define HS_NO_CULL 2         /* forward: disable the conservative per-warp footprint test (results are identical) */
#define HS_DEFER_READBACK 256 /* hs_forward_geometry, with HS_ASYNC_BINNING: additionally copy the binning counts to pinned host
                              * memory WITHOUT waiting; hs_forward_readback() waits for them later -- after the caller has
                              * enqueued hs_forward_render, so the GPU never idles behind the read-back.  The drop-in Python
                              * path uses this with capacities guessed from the previous frame and repeats the (rare) frame
                              * that did not fit synchronously. */
#define HS_REUSE_BINNING 128 /* hs_forward_render: the binning state already holds this frame's sorted tile lists (an earlier
                             * hs_forward_render call on the same geometry): only blend.  Used to render more semantic
                             * channels than the widest instantiation in several passes over the same lists. */
#define HS_SORT_GLOBAL 32    /* hs_forward_geometry: reference-style binning (offsets scan, key duplication, one global
                                radix sort) instead of the default per-tile bucket sort; sorted keys, tile lists and
                                ranges are bit-identical either way */
#define HS_ASYNC_BINNING 64  /* hs_forward_geometry: capacity mode -- no read-back, no stream synchronisation (CUDA-graph
                                capturable).  On entry *num_rendered / *binning_hint hold the CAPACITIES the caller sized
                                the binning buffer for (instances, longest tile list <= 16384; typically taken from an
                                earlier synchronous call on a similar frame plus slack) and are handed on to
                                hs_forward_render / hs_backward as if they were the counts.  If the frame needs more, the
                                fourth int32 at hs_image_state_info_offset() inside image_state becomes 1, the frame
                                renders empty (memory-safely) and the caller must repeat it synchronously; ints 0..2 there
                                are the actual instance count, longest list and number of short lists. */

/* Mirror of GaussianRasterizationSettings (diff_gaussian_rasterization/__init__.py:161-173). */
typedef struct hs_camera {
    int image_height;
    int image_width;
    float tanfovx;
    float tanfovy;
    float scale_modifier;
    const float* viewmatrix; /* device, 16 floats: world-to-camera, transposed ([1,4,4] made contiguous)      */
    const float* projmatrix; /* device, 16 floats: full projection, transposed                                */
    const float* bg;         /* device, 3 floats                                                              */
    const float* campos;     /* device, 3 floats (only used by the spherical-harmonics colour path)            */
    int prefiltered;
    int debug;
} hs_camera;

int hs_abi_version(void);
const char* hs_last_error(void);
/* 1 if kernels for S semantic channels are instantiated in this build (S = 0 is the non-semantic variant). */
int hs_supports_semantic_channels(int S);

/* Sizes of the opaque state buffers (bytes).  Require a CUDA device (temp-storage queries). */
size_t hs_geom_state_bytes(int P);
/* ... including the packed per-Gaussian records that hs_forward_render writes for its TMA row gather (S = the widest
 * semantic channel count that will be rendered from this state) */
size_t hs_geom_state_bytes_rows(int P, int S);
size_t hs_image_state_bytes(int image_height, int image_width);
size_t hs_binning_state_bytes(int num_rendered);
/* byte offset, inside image_state, of four int32: {num_rendered, longest tile list, short lists, capacity overflow} */
size_t hs_image_state_info_offset(int image_height, int image_width);

/* Stage 1 of the forward: per-Gaussian projection / cull / tile counting.
 * Writes radii[P] (int32, device), fills geom_state and the tile ranges inside image_state, and returns (HOST
 * pointers) the number of (Gaussian, tile) instances in *num_rendered and, in *binning_hint, an opaque value that
 * must be handed to hs_forward_render (it encodes the longest tile list; -1 with HS_SORT_GLOBAL).  Synchronises `stream` once (the only host sync of a forward+backward).
 * scales/rotations may be NULL iff cov3D_precomp is given, and vice versa.  flags: 0, HS_SORT_GLOBAL or HS_ASYNC_BINNING.
 * shs (may be NULL): spherical-harmonics coefficients [P, sh_coeffs, 3]; when given, the view-dependent colours of
 * degree sh_degree (0..3, reference forward.cu:20-71) are evaluated into geom_state and hs_forward_render /
 * hs_backward are called with colors == NULL. */
int hs_forward_geometry(const hs_camera* cam, int P, const float* means3D, const float* opacities,
                        const float* scales, const float* rotations, const float* cov3D_precomp, const float* shs,
                        int sh_degree, int sh_coeffs, int* radii, void* geom_state, size_t geom_state_bytes,
                        void* image_state, size_t image_state_bytes, int flags, int* num_rendered, int* binning_hint,
                        void* stream);

/* Waits for the counts a hs_forward_geometry(HS_ASYNC_BINNING | HS_DEFER_READBACK) call of THIS THREAD sent to the host:
 * counts = {num_rendered, longest tile list, tiles in the small sort class, overflow flag}. */
int hs_forward_readback(int counts[4]);

/* Stage 2 of the forward: instance scatter + per-tile sort (or key duplication, global sort and range
 * identification), alpha compositing.  num_rendered and binning_hint are the values stage 1 returned; image_state
 * is the buffer stage 1 filled.  geom_state must have been sized by hs_geom_state_bytes_rows(P, S): the call packs one
 * record per visible Gaussian behind the geometry arrays, from which the blend kernel gathers each tile's batch with TMA
 * (cp.async.bulk.tensor ... tile::gather4, four Gaussian rows per instruction).
 * Outputs (device): out_color[3,H,W], out_semantic[S,H,W] (S > 0), out_depth[1,H,W], out_median_depth[1,H,W],
 * out_opacity[1,H,W], out_mask[1,H,W] (may be NULL; only written when S == 0).  No output needs initialisation. */
int hs_forward_render(const hs_camera* cam, int P, int S, int num_rendered, int binning_hint, const float* colors,
                      const float* semantics, const int* radii, void* geom_state, size_t geom_state_bytes,
                      void* binning_state, size_t binning_state_bytes, void* image_state, size_t image_state_bytes, float* out_color,
                      float* out_semantic, float* out_depth, float* out_median_depth, float* out_opacity,
                      float* out_mask, int flags, void* stream);

/* Backward.  Upstream gradients may be NULL (treated as zero and never read).
 * Accumulated outputs — MUST be zero-initialised by the caller: dL_dmeans2D[P,3], dL_dconic[P,4],
 * dL_dopacity[P], dL_dcolors[P,3], dL_dsemantics[P,S], dL_ddepths[P].
 * Plain outputs — fully written: dL_dmeans3D[P,3], dL_dcov3D[P,6], dL_dscales[P,3], dL_drotations[P,4]
 * (the last two may be NULL when cov3D_precomp is used); dL_dsh[P, sh_coeffs, 3] when colors == NULL and shs is given
 * (spherical-harmonics colour path; its view-direction term is added to dL_dmeans3D).
 * Optional camera-pose pre-reduction (both NULL to skip): when means3D = W [p_world; 1] for a 3x4 pose W, pass
 * pose_points = p_world [P,3]; dL_dpose[12] (row-major 3x4, zero-initialised by the caller) then accumulates
 * dL/dW = sum_i dL_dmeans3D_i (x) [p_world_i, 1], reduced in the per-Gaussian kernel (warp shuffle, shared memory,
 * 12 atomics per block) instead of by an autograd matmul over [P,3] in the caller. */
int hs_backward(const hs_camera* cam, int P, int S, int num_rendered, const float* means3D, const int* radii,
                const float* colors, const float* semantics, const float* scales, const float* rotations,
                const float* cov3D_precomp, const float* shs, int sh_degree, int sh_coeffs, const void* geom_state, const void* binning_state,
                const void* image_state, const float* dL_dout_color, const float* dL_dout_semantic,
                const float* dL_dout_depth, const float* dL_dout_median_depth, const float* dL_dout_opacity,
                float* dL_dmeans2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors, float* dL_dsemantics,
                float* dL_ddepths, float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales, float* dL_drotations,
                float* dL_dsh, const float* pose_points, float* dL_dpose, int flags, void* stream);

/* Extension (first step of the fused loss epilogue): masked L1 image loss and its gradient in one pass.
 * loss[0] (device, zero-initialised by the caller) += sum over channels c and masked pixels p of |pred[c,p] - target[c,p]|;
 * grad[c,p] = mask[p] * sign(pred - target).  mask: [pixels] bytes (non-zero = use), NULL = every pixel.
 * Replaces `torch.abs(gt - x)[mask].sum()` of the reference's losses (scripts/hierslam.py:780-796). */
int hs_masked_l1(const float* pred, const float* target, const unsigned char* mask, int channels, size_t pixels,
                 float* loss, float* grad, void* stream);

/* Extension: hierarchical cross-entropy of the tree-encoded semantic map (scripts/hierslam.py:955-1000).  sem[S,pixels]
 * planar (the rasterizer's out_semantic); labels[levels,pixels] int32 (device; negative = ignored); level l owns the
 * channels [level_begin[l], level_begin[l+1]) (HOST array of levels+1 ints); level_scale[l] (HOST) = weight_l / number of
 * non-ignored pixels.  loss[0] (device, zeroed by the caller) += sum_l level_scale[l] * sum_p CE; grad[c,p] = d loss / d sem
 * for every channel below level_begin[levels] (at most 8 levels). */
int hs_hier_cross_entropy(const float* sem, const int* labels, int levels, const int* level_begin,
                          const float* level_scale, size_t pixels, float* loss, float* grad, void* stream);

/* Extension: the steps of a tracking iteration between the rasterizer calls
 * (scripts/hierslam.py:1837-1856, get_loss_semantic(tracking=True) :765-796, transform_to_frame utils/slam_helpers.py:278-330),
 * each as one kernel so that a whole iteration is ~14 launches and stream-capturable.
 * hs_transform_points: cam[P,3] = R world + t for the row-major 4x4 w2c (device).
 * hs_tracking_loss: mask = gt_depth > 0 && !isnan(depth) [&& silhouette > sil_thres]; loss[0] (device, accumulated) +=
 *   depth_weight * sum_mask |gt_depth - depth| + im_weight * sum_c |gt_im - im| (over the same mask when use_silhouette,
 *   over EVERY pixel otherwise: scripts/hierslam.py:789-794); grad_im[3,pixels], grad_depth[pixels] = d loss / d im,
 *   d loss / d depth.  The reference's `ignore_outlier_depth_loss` option (a median over the depth error) is not offered.
 * hs_pose_step (mode 1): loss[0] is read and reset to 0; dL_dpose[3,4] (hs_backward) -> gradient of the unnormalised
 *   quaternion cam_rot[4] (r,x,y,z) and translation cam_tran[3]; torch.optim.Adam's update with per-tensor learning rates;
 *   best-candidate bookkeeping as the reference does it (scripts/hierslam.py:1851-1858: the pose AFTER the step is saved
 *   when the loss evaluated before it is the smallest so far); w2c[16] of the updated pose.  binning_info (NULL = not
 *   used): the uint32[4] counts of the iteration's forward (hs_image_state_info_offset); when its overflow word is set
 *   (capacity-mode binning: the frame rendered empty) the iteration is discarded -- no candidate, no update -- and
 *   state[24] is latched to 1 for the caller to read once per frame.  state: HS_POSE_STATE_FLOATS floats (device) =
 *   exp_avg[7] | exp_avg_sq[7] | step | min_loss | candidate rot[4] | candidate tran[3] | last loss | overflow latch |
 *   largest num_rendered | longest tile list seen in the frame (as floats) | reserved[5]; the caller zeroes it and sets min_loss large at frame start.  mode 0: only write w2c for the current pose. */
#define HS_POSE_STATE_FLOATS 32
int hs_transform_points(const float* w2c, const float* world, int P, float* cam, void* stream);
int hs_tracking_loss(const float* im, const float* depth, const float* silhouette, const float* gt_im,
                     const float* gt_depth, size_t pixels, float sil_thres, int use_silhouette, float depth_weight,
                     float im_weight, float* loss, float* grad_im, float* grad_depth, void* stream);
int hs_pose_step(float* cam_rot, float* cam_tran, const float* dL_dpose, float* loss, float* state, float* w2c,
                 const unsigned int* binning_info, float lr_rot, float lr_tran, float beta1, float beta2, float eps, int mode,
                 void* stream);

/* Extension: keyframe selection by re-projection (utils/keyframe_selection.py:40-96).
 * counts[k] (int32, device) = number of the num_points world points[.,3] that project into keyframe k's image
 * (w2c[k] row-major 4x4, pinhole fx fy cx cy) with edge < u < width - edge, edge < v < height - edge and depth + 1e-5 > 0. */
int hs_keyframe_overlap(const float* points, int num_points, const float* w2c, int keyframes, float fx, float fy, float cx,
                        float cy, int width, int height, int edge, int* counts, void* stream);

/* Extension (parameter maintenance): torch.optim.Adam's update (scripts/hierslam.py:411-417;
 * torch/optim/adam.py::_multi_tensor_adam, no weight decay / amsgrad) for every parameter tensor in ONE pass over flat
 * buffers.  param / grad / exp_avg / exp_avg_sq: n floats each (device, n a multiple of 4); segment s covers the floats
 * [segment_end[s-1], segment_end[s]) (HOST arrays, ends ascending and multiples of 4, at most 16 segments) and has the
 * learning rate segment_lr[s]; step = the 1-based step count (bias correction).  Learning rates, betas and eps are
 * doubles because torch derives its float scalars (1 - beta, lr / bias_correction, ...) from Python doubles. */
int hs_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int segments,
                 const unsigned long long* segment_end, const double* segment_lr, double beta1, double beta2, double eps,
                 int step, void* stream);
/* The same update with the step count on the DEVICE, so that a whole mapping iteration (render, losses, backward, this
 * step) can be captured in one CUDA graph and replayed: step_counter[0] (int32, device; 0 before the first step) is
 * incremented by a one-thread kernel that also evaluates the bias-correction scalars (in double, as above) into
 * scalars[HS_ADAM_SCALARS] (float32, device scratch); the update kernel reads them from there. */
#define HS_ADAM_SCALARS 17
int hs_adam_step_device(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int segments,
                        const unsigned long long* segment_end, const double* segment_lr, double beta1, double beta2, double eps,
                        int* step_counter, float* scalars, void* stream);

/* Extension: Gaussian pruning (utils/slam_external.py:142-164 remove_points: `tensor[to_keep]` on every parameter and both
 * Adam moments).  hs_compact_plan turns keep[P] (device bytes, non-zero = keep) into an ordered source-row list inside
 * scratch (hs_compact_scratch_bytes(P) bytes, device); afterwards the unsigned at scratch[ceil(P/1024)] is the number of
 * kept rows.  hs_compact_gather copies, for every segment s (rows of width[s] floats starting at src_offset[s] floats in
 * src), the kept rows in their original order to dst + dst_offset[s].  Bit-exact, order-preserving. */
size_t hs_compact_scratch_bytes(int P);
int hs_compact_plan(const unsigned char* keep, int P, void* scratch, void* stream);
int hs_compact_gather(const float* src, float* dst, const void* scratch, int P, int rows, int segments,
                      const unsigned long long* src_offset, const unsigned long long* dst_offset, const int* width,
                      void* stream);

/* Extension: the colour loss of mapping, l1_weight * mean|pred - target| + ssim_weight * (1 - SSIM)
 * (scripts/hierslam.py:936; SSIM of utils/slam_external.py:55-97: 11-tap Gaussian window, zero padding, c1 = 0.01^2,
 * c2 = 0.03^2, mean over channels and pixels), and its gradient.  pred / target: [channels,height,width] planar (device);
 * window11: the 11 normalised 1-D window taps (HOST); loss[0] (device, zeroed by the caller) += l1_scale * sum|pred -
 * target| + ssim_scale * sum(ssim map) -- the caller passes l1_scale = l1_weight / n, ssim_scale = -ssim_weight / n and
 * adds the constant ssim_weight; scratch: 3 * channels * height * width floats (device); grad[channels,height,width] =
 * d loss / d pred (NULL = forward only). */
int hs_l1_ssim(const float* pred, const float* target, int channels, int height, int width, const float* window11,
               float l1_scale, float ssim_scale, float* loss, float* scratch, float* grad, void* stream);

/* Extension: leaf-level loss of the tree encoding -- a 1x1 convolution from the S rendered channels to the leaf classes
 * followed by cross-entropy (scripts/hierslam.py:975-984, :1009-1016; MLP_func = Conv2d(S, classes, 1), :1756) -- with
 * all three gradients, without ever materialising the [classes,pixels] logits.  sem[S,pixels] planar; labels[pixels]
 * int32 (device; negative or >= classes = ignored); weight[classes,S] row-major (the Conv2d weight); bias[classes] or
 * NULL; scale (HOST) = loss weight / number of non-ignored pixels.  loss[0] (device, zeroed by the caller) += scale *
 * sum_p CE; lse[pixels] (device scratch) receives the per-pixel logsumexp; grad_sem[S,pixels] = (or += with
 * HS_LEAF_ACCUMULATE) d loss / d sem; grad_weight[classes,S] and grad_bias[classes] (device, zeroed by the caller;
 * grad_weight NULL = skip both) += d loss / d weight, d loss / d bias.  1 <= S <= 79.  The contractions run on the
 * tensor cores as 3xTF32 (fp32-accurate); HS_LEAF_TF32 selects a single TF32 product (what torch's convolution does
 * with its default allow_tf32). */
#define HS_LEAF_ACCUMULATE 1
#define HS_LEAF_TF32 2
int hs_leaf_cross_entropy(const float* sem, const int* labels, const float* weight, const float* bias, int channels,
                          int classes, size_t pixels, float scale, float* loss, float* lse, float* grad_sem,
                          int flags, float* grad_weight, float* grad_bias, void* stream);

/* The same loss with the per-pixel pass (logits, softmax statistics, loss, d loss / d sem) on the 5th-generation tensor
 * cores: tcgen05.mma kind::tf32 with the accumulators in tensor memory (TMEM), 3xTF32 (fp32-accurate), the weights laid
 * out once per call as UMMA operand tiles in `workspace` (device, 128-byte aligned, hs_leaf_ce_workspace_bytes(channels,
 * classes) bytes) and streamed to shared memory with TMA bulk copies (hier_slam_b200/csrc/leaf_loss_tc.cu).  Arguments
 * and results as hs_leaf_cross_entropy; the weight / bias gradients come from the same kernel as there. */
size_t hs_leaf_ce_workspace_bytes(int channels, int classes);
/* Debugging aid: device buffer of 60 int64 that receives clock64 stamps of CTA 0's first phases (NULL = off). */
void hs_leaf_tc_debug(long long* stamps);
int hs_leaf_cross_entropy_tc(const float* sem, const int* labels, const float* weight, const float* bias, int channels,
                             int classes, size_t pixels, float scale, float* loss, float* lse, float* grad_sem, int flags,
                             float* grad_weight, float* grad_bias, void* workspace, size_t workspace_bytes, void* stream);

/* Keyframe-parallel mapping (the reference has no multi-GPU mode): in-place SUM all-reduce of `count`
 * floats (a multiple of 4) that live at the SAME offset of a symmetric-memory allocation on every rank, as one kernel per
 * rank over NVLink peer memory (hier_slam_b200/csrc/allreduce.cu).  peer_buffers[r] / peer_signal_pads[r] (HOST arrays of
 * world_size device pointers): rank r's buffer and signal pad as mapped into THIS process; multicast_ptr: the NVSwitch
 * multicast mapping of the buffer (multimem.ld_reduce / multimem.st do the reduction and the broadcast in the switch) or
 * NULL (peer loads / stores instead).  The signal pad needs blocks * world_size 32-bit words (word [block][rank]), zeroed once; `epoch`
 * must grow by 2 from call to call, starting at 1, identically on every rank.  All ranks must call it in the same order.
 * Waits are bounded (4 s): a missing peer raises a CUDA error instead of hanging. */
int hs_allreduce_sum(void* multicast_ptr, void* const* peer_buffers, void* const* peer_signal_pads, int rank, int world_size,
                     size_t count, unsigned int epoch, int blocks, void* stream);

/* present[P] (bool, device) = view-space z > 0.2 (reference: rasterizer_impl.cu:54-66). */
int hs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    unsigned char* present, void* stream);

/* Measurement aids.  After hs_profile_enable(1) every kernel launch of this library is bracketed by its own
 * CUDA event pair on the launching stream.  hs_profile_read() waits for all recorded pairs and returns, per
 * stage, the summed elapsed milliseconds and the number of launches since the previous read, in the order
 * [preprocess, scan, duplicate, sort, ranges, blend_fwd, blend_bwd, geom_bwd].
 * With the default tile-bucket binning "scan" is the tile scan, "duplicate" the instance scatter, "sort" the per-tile
 * sort and "ranges" is empty.  hs_kernel_launch_count() = kernels of this library launched so far;
 * hs_library_call_count() = CUB device-wide primitives (scan, radix sort; HS_SORT_GLOBAL only) called so far. */
int hs_profile_enable(int on);
int hs_profile_read(float total_ms[8], int count[8]);
long long hs_kernel_launch_count(void);
long long hs_library_call_count(void);

/* Test / debugging aid: byte offsets of the arrays inside the opaque state buffers.
 *   geom    : [depths f32[P], means2D f32[2P], conic_opacity f32[4P], tiles_touched u32[P], point_offsets u32[P]]
 *   image   : [final_T f32[N], n_contrib u32[N], ranges u32[2*tiles]]
 *   binning : [point_list u32[R], point_list_unsorted u32[R], keys u64[R], keys_unsorted u64[R], strip_hits u8[R]] */
int hs_geom_state_layout(int P, size_t offsets[5]);
int hs_image_state_layout(int image_height, int image_width, size_t offsets[3]);
int hs_binning_state_layout(int num_rendered, size_t offsets[5]);

#ifdef __cplusplus
}
#endif
#endif /* HS_RASTER_H_ */

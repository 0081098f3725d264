// Internal header of libhsraster (sm_100a).  Not part of the public C ABI (see include/hs_raster.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#define HS_TILE_X 16          // reference: cuda_rasterizer/config.h:16  (BLOCK_X)
#define HS_TILE_Y 16          // reference: cuda_rasterizer/config.h:17  (BLOCK_Y)
#define HS_TILE_PIX 256
// Pixels of a warp inside a 16x16 tile.  Region r (0..7) is what the strip masks / strip_hits are indexed by; lane p (0..31).
//   HS_REGION_8x4 = 1: 8 wide x 4 high blocks (2 columns x 4 rows of regions) -- a pixel-sized Gaussian (footprint ~8x8)
//                      overlaps ~6 such blocks = 192 lane evaluations;
//   HS_REGION_8x4 = 0: 16 x 2 strips -- the same Gaussian overlaps ~7.5 strips = 240 lane evaluations.
#ifndef HS_REGION_8x4
#define HS_REGION_8x4 1
#endif
#if HS_REGION_8x4
#define HS_REGION_X0(r) (((r) & 1) << 3)
#define HS_REGION_Y0(r) (((r) >> 1) << 2)
#define HS_REGION_W 8
#define HS_REGION_H 4
#define HS_PX_X(r, p) (HS_REGION_X0(r) + ((p) & 7))
#define HS_PX_Y(r, p) (HS_REGION_Y0(r) + ((p) >> 3))
#else
#define HS_REGION_X0(r) 0
#define HS_REGION_Y0(r) (2 * (r))
#define HS_REGION_W 16
#define HS_REGION_H 2
#define HS_PX_X(r, p) ((p) & 15)
#define HS_PX_Y(r, p) (2 * (r) + ((p) >> 4))
#endif
#define HS_MAX_SEGMENTS 16     // parameter tensors in one flat buffer (hs_adam_step, hs_compact_gather)
#define HS_MAX_LEVELS 8        // levels of the hierarchical semantic encoding (hs_hier_cross_entropy)
#define HS_ALIGN 256          // every array inside an opaque state buffer is 256-B aligned
#define HS_CTR_STRIDE 32          // per-tile counters sit 128 B apart: L2 atomics serialise per line, not per word
#define HS_TILE_SORT_SMALL 2048   // tile lists up to this length are sorted by 256-thread CTAs
#define HS_TILE_SORT_MAX 16384    // longest tile list the in-shared-memory sort takes (128 KB); beyond it the
                                  // global radix sort is used for the whole frame

namespace hs {

void set_error(const char* fmt, ...);

// ---- optional per-stage timing with CUDA events on the launching stream (hs_profile_enable / hs_profile_read)
enum Stage : int {
    ST_PREPROCESS = 0, ST_SCAN, ST_DUPLICATE, ST_SORT, ST_RANGES, ST_BLEND_FWD, ST_BLEND_BWD, ST_GEOM_BWD, ST_COUNT
};
void prof_begin(int stage, cudaStream_t stream);
void prof_end(int stage, cudaStream_t stream);
void count_launch(int n = 1);

#define HS_CUDA_OK(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            hs::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                          cudaGetErrorString(_e));                                           \
            return 2;                                                                        \
        }                                                                                    \
    } while (0)

// After every launch: always check the launch itself; with debug additionally synchronise the
// stream and surface execution errors (mirrors CHECK_CUDA, cuda_rasterizer/auxiliary.h:166-173).
#define HS_LAUNCH_OK(stream, debug)                                                          \
    do {                                                                                     \
        hs::count_launch();                                                                  \
        HS_CUDA_OK(cudaGetLastError());                                                      \
        if (debug) HS_CUDA_OK(cudaStreamSynchronize(stream));                                \
    } while (0)

static inline size_t align_up(size_t x) { return (x + HS_ALIGN - 1) & ~(size_t)(HS_ALIGN - 1); }

// ---- opaque state buffers (our private layout; the three buffers mirror geomBuffer / binningBuffer /
//      imgBuffer of RAST/rasterize_points.cu:285-291 in role, not in layout) ------------------------------
struct GeomView {
    float* depths;            // f32[P]   view-space z
    float2* means2D;          // float2[P] pixel centre
    float4* conic_opacity;    // float4[P] (conic.x, conic.y, conic.z, opacity)
    uint32_t* tiles_touched;  // u32[P]
    uint32_t* point_offsets;  // u32[P]   inclusive scan of tiles_touched
    float* rgb;               // f32[3P]  colours evaluated from spherical harmonics (SH colour path only)
    uint8_t* clamped;         // u8[P]    bit c: colour channel c was clamped at 0 (SH colour path only)
    char* scan_temp;          // CUB scan temp
    size_t scan_temp_bytes;
    size_t total_bytes;
    float* rows;              // f32[P][packed_row_floats(S)] packed per-Gaussian records for the TMA row gather of the forward
                              // blend (behind everything else; present when the buffer was sized by hs_geom_state_bytes_rows)
};
// Packed per-Gaussian record of the forward blend: [conic.xyz opacity | x y - - | r g b depth | sem 0..S-1 | pad], padded to
// a multiple of 8 floats so that four consecutive records are a multiple of 128 bytes (the shared-memory alignment of a
// TMA tile::gather4 destination).
static inline int packed_row_floats(int S) { return (8 + ((4 + S + 3) & ~3) + 7) & ~7; }
struct ImageView {
    float* final_T;           // f32[N]
    uint32_t* n_contrib;      // u32[N]
    uint2* ranges;            // uint2[tiles]
    // tile-bucket binning (default path): per-tile instance counts (turned into scatter cursors by the tile scan)
    // and info = {num_rendered, longest tile list}
    uint32_t* tile_count;     // u32[tiles]
    uint32_t* info;           // u32[4]
    size_t total_bytes;
};
struct BinningView {
    uint32_t* point_list;          // u32[R] sorted Gaussian ids
    uint32_t* point_list_unsorted; // u32[R]
    uint64_t* keys;                // u64[R] sorted (tile << 32 | depth bits)
    uint64_t* keys_unsorted;       // u64[R]
    uint8_t* strip_hits;           // u8[R]  bit w: the forward blended this list entry into warp strip w (rows 2w, 2w+1)
    char* sort_temp;
    size_t sort_temp_bytes;
    size_t total_bytes;
};

int geom_view(char* base, size_t P, GeomView* v);
int image_view(char* base, size_t N, size_t tiles, ImageView* v);
int binning_view(char* base, size_t R, BinningView* v);

// ---- kernels' host launchers (defined in the .cu files) ----------------------------------------------
struct Camera {
    int W, H;
    float tanfovx, tanfovy, focal_x, focal_y, scale_modifier;
    const float* view;   // device, 16 floats, m[4*c+r]
    const float* proj;   // device, 16 floats
    int grid_x, grid_y;
};

int launch_preprocess(int P, const float* means3D, const float* scales, const float* rotations,
                      const float* opacities, const float* cov3D_precomp, const Camera& cam, int* radii,
                      const GeomView& g, uint32_t* tile_count, cudaStream_t stream, bool debug);
int launch_scan(int P, const GeomView& g, cudaStream_t stream, bool debug);
int launch_tile_scan(const Camera& cam, const ImageView& img, uint32_t r_cap, uint32_t tile_cap, cudaStream_t stream,
                     bool debug);
int launch_tile_binning(int P, int R, int max_tile, int n_small, const Camera& cam, const int* radii, const GeomView& g,
                        const BinningView& b, const ImageView& img, cudaStream_t stream, bool debug);
int launch_binning(int P, int R, const Camera& cam, const int* radii, const GeomView& g, const BinningView& b,
                   const ImageView& img, cudaStream_t stream, bool debug);
int launch_pack_rows(int P, const int* radii, int S, const GeomView& g, const float* colors, const float* semantics,
                     cudaStream_t stream);
int launch_blend_forward(int P, const int* radii, int S, const Camera& cam, const GeomView& g, const BinningView& b,
                         const ImageView& img, const float* colors, const float* semantics, float* out_color, float* out_semantic,
                         float* out_depth, float* out_median, float* out_opacity, float* out_mask, int flags,
                         cudaStream_t stream, bool debug);
int launch_blend_backward(int S, const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                          const float* bg, const float* colors, const float* semantics, const float* dL_color,
                          const float* dL_sem, const float* dL_depth, const float* dL_median,
                          const float* dL_opacity, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolors, float* dL_dsemantics, float* dL_ddepths, int flags,
                          cudaStream_t stream, bool debug);
int launch_blend_backward_mma(int S, const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                              const float* bg, const float* colors, const float* dL_color, const float* dL_sem,
                              const float* dL_depth, const float* dL_median, const float* dL_opacity,
                              float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors,
                              float* dL_dsemantics, float* dL_ddepths, cudaStream_t stream, bool debug);
int launch_geom_backward(int P, const float* means3D, const int* radii, const float* scales,
                         const float* rotations, const float* cov3D_precomp, const Camera& cam,
                         const float* dL_dmean2D, const float* dL_dconic, const float* dL_ddepths,
                         float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales, float* dL_drots,
                         const float* pose_points, float* dL_dpose, cudaStream_t stream, bool debug);
int launch_sh_forward(int P, int deg, int M, const float* means3D, const float* campos, const float* shs,
                      const int* radii, const GeomView& g, cudaStream_t stream, bool debug);
int launch_sh_backward(int P, int deg, int M, const float* means3D, const float* campos, const float* shs,
                       const int* radii, const GeomView& g, const float* dL_dcolors, float* dL_dmeans3D, float* dL_dsh,
                       const float* pose_points, float* dL_dpose, cudaStream_t stream, bool debug);
int launch_masked_l1(const float* pred, const float* target, const uint8_t* mask, int C, size_t HW, float* loss,
                     float* grad, cudaStream_t stream);
int launch_hier_cross_entropy(const float* sem, const int* labels, int L, const int* level_begin, const float* level_scale,
                              size_t HW, float* loss, float* grad, cudaStream_t stream);
int launch_transform_points(const float* w2c, const float* world, int P, float* cam, cudaStream_t stream);
int launch_tracking_loss(const float* im, const float* depth, const float* sil, const float* gt_im, const float* gt_depth,
                         size_t HW, float sil_thres, int use_sil, float w_depth, float w_im, float* loss, float* grad_im,
                         float* grad_depth, cudaStream_t stream);
int launch_pose_step(float* cam_rot, float* cam_tran, const float* dL_dpose, float* loss, float* state, float* w2c,
                     const uint32_t* binning_info, float lr_rot, float lr_tran, float beta1, float beta2, float eps, int mode, cudaStream_t stream);
int launch_keyframe_overlap(const float* pts, int N, const float* w2c, int K, float fx, float fy, float cx, float cy, int width,
                            int height, int edge, int* counts, cudaStream_t stream);
int launch_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int n_seg,
                     const unsigned long long* seg_end, const double* seg_lr, double beta1, double beta2, double eps, int step,
                     cudaStream_t stream, int* step_counter = nullptr, float* dev_scalars = nullptr);
size_t compact_scratch_bytes(int P);
int launch_compact_plan(const uint8_t* keep, int P, unsigned* scratch, cudaStream_t stream);
int launch_compact_gather(const float* src, float* dst, const unsigned* scratch, int P, int rows, int n_seg,
                          const unsigned long long* src_off, const unsigned long long* dst_off, const int* width,
                          cudaStream_t stream);
int launch_l1_ssim(const float* pred, const float* target, int C, int H, int W, const float* window11, float l1_scale,
                   float ssim_scale, float* loss, float* scratch, float* grad, cudaStream_t stream);
int launch_leaf_cross_entropy(const float* sem, const int* labels, const float* weight, const float* bias, int S, int L,
                              size_t HW, float scale, float* loss, float* lse, float* grad_sem, int accumulate,
                              float* grad_weight, float* grad_bias, int single_tf32, cudaStream_t stream);
size_t leaf_tc_workspace_bytes(int S, int L);
void leaf_tc_set_debug(long long* p);
int launch_leaf_cross_entropy_tc(const float* sem, const int* labels, const float* weight, const float* bias, int S, int L,
                                 size_t HW, float scale, float* loss, float* lse, float* grad_sem, int accumulate,
                                 float* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_leaf_weight_grad(const float* sem, const int* labels, const float* weight, const float* bias, const float* lse,
                            int S, int L, size_t HW, float scale, float* grad_weight, float* grad_bias, int single_tf32,
                            cudaStream_t stream);
int launch_allreduce_sum(void* multicast_ptr, void* const* peer_bufs, void* const* peer_pads, int rank, int world, size_t n,
                         unsigned epoch, int blocks, cudaStream_t stream);
int launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, bool* present,
                        cudaStream_t stream, bool debug);

// power = -0.5 (A dx^2 + C dy^2) - B dx dy with the roundings of the reference's sm_100 SASS pinned
// (forward.cu:490 / backward.cu:804: FMUL A*dx, C*dy, (C*dy)*dy, B*dx, (B*dx)*dy; FFMA dx*(A*dx)+..; FFMA *-0.5 - ..),
// so that the alpha < 1/255 and T < 1e-4 decisions -- hence n_contrib and final_T -- agree bit for bit.
#ifdef __CUDACC__
__device__ __forceinline__ float gauss_power(const float2 d, const float4 con_o) {
    const float q = __fmaf_rn(d.x, __fmul_rn(d.x, con_o.x), __fmul_rn(d.y, __fmul_rn(d.y, con_o.z)));
    return __fmaf_rn(q, -0.5f, -__fmul_rn(d.y, __fmul_rn(d.x, con_o.y)));
}
// Conservative pixel-space box outside of which alpha < 1/255 for this Gaussian (or power > 0).
// alpha = min(0.99, o * exp(power)) >= 1/255  <=>  power >= -ln(255 o); power = -q/2 with
// q = A dx^2 + 2 B dx dy + C dy^2, so |dx| <= sqrt(2 tau C / det), |dy| <= sqrt(2 tau A / det).
// All margins err on the side of keeping the Gaussian; non-finite or degenerate inputs disable the test.
__device__ __forceinline__ float4 footprint_box(const float2 xy, const float4 co) {
    const float kInf = __int_as_float(0x7f800000);
    float4 all = {-kInf, kInf, -kInf, kInf};
    const float A = co.x, B = co.y, C = co.z, o = co.w;
    if (!(o >= 0.0039f)) {  // strictly below 1/255 = 0.0039215...: can never pass the alpha test
        if (o < 0.0039f) return {kInf, -kInf, kInf, -kInf};
        return all;  // NaN opacity: let the exact test decide
    }
    const float tau = __logf(o * 255.0f) + 0.02f;
    const float ac = A * C, bb = B * B;
    const float det = (ac - bb) - 1e-6f * (fabsf(ac) + bb);
    if (!(det > 0.f) || !(A > 0.f) || !(C > 0.f) || !(tau > 0.f)) return all;
    const float k = 2.0f * tau / det;
    const float hx = sqrtf(k * C) * 1.001f + 0.01f;
    const float hy = sqrtf(k * A) * 1.001f + 0.01f;
    if (!(hx < 1e8f) || !(hy < 1e8f)) return all;
    return {xy.x - hx, xy.x + hx, xy.y - hy, xy.y + hy};
}
#endif

// flags shared by the blend kernels
enum : int {
    HS_FLAG_SEM_ALPHA_EXACT = 1,   // semantic channels contribute to dL/dalpha (reference quirk Q1 off)
    HS_FLAG_NO_CULL = 2,           // disable the conservative per-warp footprint test
    HS_FLAG_SEM_UNALIGNED = 16,    // internal: semantics base pointer is only 4-byte aligned (scalar cp.async)
    HS_FLAG_BWD_SHUFFLE = 4,       // backward: SIMT warp-shuffle reduction instead of the tensor-core path
};

}  // namespace hs

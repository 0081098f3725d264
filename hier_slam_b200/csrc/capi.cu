// extern "C" entry points of libhsraster (see include/hs_raster.h for the contract).
#include "../../include/hs_raster.h"
#include "hs_common.cuh"

namespace hs {
const char* last_error();
void prof_enable(bool on);
int prof_read(float* total_ms, int* count);
long long launches();
long long lib_calls();
}

using namespace hs;

// binning_hint (opaque to the caller): longest tile list in the low bits, flag = no tile in the small sort class
#define HS_HINT_MASK 0x0fffffff
#define HS_HINT_NO_SMALL 0x10000000

static int make_camera(const hs_camera* c, Camera* cam) {
    if (c == nullptr) {
        set_error("hs_camera is NULL");
        return 1;
    }
    if (c->image_width <= 0 || c->image_height <= 0) {
        set_error("invalid image size %dx%d", c->image_width, c->image_height);
        return 1;
    }
    cam->W = c->image_width;
    cam->H = c->image_height;
    cam->tanfovx = c->tanfovx;
    cam->tanfovy = c->tanfovy;
    // reference: rasterizer_impl.cu:489-490
    cam->focal_y = c->image_height / (2.0f * c->tanfovy);
    cam->focal_x = c->image_width / (2.0f * c->tanfovx);
    cam->scale_modifier = c->scale_modifier;
    cam->view = c->viewmatrix;
    cam->proj = c->projmatrix;
    cam->grid_x = (c->image_width + HS_TILE_X - 1) / HS_TILE_X;
    cam->grid_y = (c->image_height + HS_TILE_Y - 1) / HS_TILE_Y;
    return 0;
}

static int* pinned_int() {
    static thread_local int* p = nullptr;
    if (p == nullptr) {
        if (cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault) != cudaSuccess) p = nullptr;
    }
    return p;
}

// deferred read-back of the binning counts (HS_DEFER_READBACK): pinned words + the event that marks their arrival
struct Readback {
    int* host = nullptr;
    cudaEvent_t ev = nullptr;
};
static Readback* readback_slot() {
    static thread_local Readback rb;
    if (rb.host == nullptr) {
        if (cudaHostAlloc((void**)&rb.host, 64, cudaHostAllocDefault) != cudaSuccess) {
            rb.host = nullptr;
            return nullptr;
        }
        if (cudaEventCreateWithFlags(&rb.ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &rb;
}

extern "C" {

int hs_abi_version(void) { return HS_RASTER_ABI_VERSION; }
const char* hs_last_error(void) { return hs::last_error(); }

int hs_supports_semantic_channels(int S) { return S == 0 || S == 16 || S == 26 || S == 32 || S == 48 || S == 64 || S == 74 || S == 102; }

size_t hs_geom_state_bytes(int P) {
    GeomView v;
    if (geom_view(nullptr, (size_t)(P > 0 ? P : 0), &v)) return 0;
    return v.total_bytes;
}
size_t hs_geom_state_bytes_rows(int P, int S) {
    GeomView v;
    if (geom_view(nullptr, (size_t)(P > 0 ? P : 0), &v)) return 0;
    return align_up(v.total_bytes) + (size_t)(P > 0 ? P : 0) * packed_row_floats(S) * sizeof(float) + HS_ALIGN;
}
size_t hs_image_state_bytes(int H, int W) {
    ImageView v;
    const size_t tiles = (size_t)((W + HS_TILE_X - 1) / HS_TILE_X) * ((H + HS_TILE_Y - 1) / HS_TILE_Y);
    image_view(nullptr, (size_t)H * W, tiles, &v);
    return v.total_bytes;
}
size_t hs_binning_state_bytes(int R) {
    BinningView v;
    if (binning_view(nullptr, (size_t)(R > 0 ? R : 0), &v)) return 0;
    return v.total_bytes;
}

int hs_geom_state_layout(int P, size_t off[5]) {
    GeomView v;
    if (geom_view(nullptr, (size_t)P, &v)) return 2;
    off[0] = (size_t)v.depths;
    off[1] = (size_t)v.means2D;
    off[2] = (size_t)v.conic_opacity;
    off[3] = (size_t)v.tiles_touched;
    off[4] = (size_t)v.point_offsets;
    return 0;
}
int hs_image_state_layout(int H, int W, size_t off[3]) {
    ImageView v;
    const size_t tiles = (size_t)((W + HS_TILE_X - 1) / HS_TILE_X) * ((H + HS_TILE_Y - 1) / HS_TILE_Y);
    image_view(nullptr, (size_t)H * W, tiles, &v);
    off[0] = (size_t)v.final_T;
    off[1] = (size_t)v.n_contrib;
    off[2] = (size_t)v.ranges;
    return 0;
}
size_t hs_image_state_info_offset(int H, int W) {
    ImageView v;
    const size_t tiles = (size_t)((W + HS_TILE_X - 1) / HS_TILE_X) * ((H + HS_TILE_Y - 1) / HS_TILE_Y);
    image_view(nullptr, (size_t)H * W, tiles, &v);
    return (size_t)v.info;
}
int hs_binning_state_layout(int R, size_t off[5]) {
    BinningView v;
    if (binning_view(nullptr, (size_t)R, &v)) return 2;
    off[0] = (size_t)v.point_list;
    off[1] = (size_t)v.point_list_unsorted;
    off[2] = (size_t)v.keys;
    off[3] = (size_t)v.keys_unsorted;
    off[4] = (size_t)v.strip_hits;
    return 0;
}

int hs_forward_geometry(const hs_camera* c, int P, const float* means3D, const float* opacities,
                        const float* scales, const float* rotations, const float* cov3D_precomp, const float* shs,
                        int sh_degree, int sh_coeffs, int* radii, void* geom_state, size_t geom_state_bytes,
                        void* image_state, size_t image_state_bytes, int flags, int* num_rendered, int* binning_hint,
                        void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Camera cam;
    if (make_camera(c, &cam)) return 1;
    if (num_rendered == nullptr || binning_hint == nullptr) {
        set_error("num_rendered / binning_hint is NULL");
        return 1;
    }
    const int cap_instances = *num_rendered, cap_hint = *binning_hint;   // HS_ASYNC_BINNING: capacities on entry
    *num_rendered = 0;
    *binning_hint = 0;
    if (P <= 0) return 0;
    if (means3D == nullptr || opacities == nullptr || radii == nullptr || geom_state == nullptr) {
        set_error("hs_forward_geometry: NULL argument");
        return 1;
    }
    if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr)) {
        set_error("provide either scales+rotations or cov3D_precomp");
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(geom_state) & (HS_ALIGN - 1)) != 0) {
        set_error("geom_state must be %d-byte aligned", HS_ALIGN);
        return 1;
    }
    GeomView g;
    if (geom_view((char*)geom_state, (size_t)P, &g)) return 2;
    if (g.total_bytes > geom_state_bytes + HS_ALIGN) {
        set_error("geom_state too small: %zu < %zu", geom_state_bytes, g.total_bytes);
        return 1;
    }
    const bool debug = c->debug != 0;
    if (shs != nullptr) {
        const int need = (sh_degree + 1) * (sh_degree + 1);
        if (sh_degree < 0 || sh_degree > 3 || sh_coeffs < need || c->campos == nullptr) {
            set_error("spherical harmonics: degree %d needs 0 <= degree <= 3, at least %d coefficients (got %d) and campos",
                      sh_degree, need, sh_coeffs);
            return 1;
        }
    }
    int* host = (flags & HS_ASYNC_BINNING) ? nullptr : pinned_int();
    if (host == nullptr && !(flags & HS_ASYNC_BINNING)) {
        set_error("cudaHostAlloc failed");
        return 2;
    }
    if ((flags & HS_SORT_GLOBAL) && (flags & HS_ASYNC_BINNING)) {
        set_error("HS_ASYNC_BINNING needs the tile-bucket binning (not HS_SORT_GLOBAL)");
        return 1;
    }
    if (flags & HS_SORT_GLOBAL) {
        // reference-style binning: offsets scan over the Gaussians, the count is the last offset
        int rc = launch_preprocess(P, means3D, scales, rotations, opacities, cov3D_precomp, cam, radii, g, nullptr,
                                   stream, debug);
        if (rc) return rc;
        rc = launch_scan(P, g, stream, debug);
        if (rc) return rc;
        if (shs != nullptr) rc = launch_sh_forward(P, sh_degree, sh_coeffs, means3D, c->campos, shs, radii, g, stream, debug);
        if (rc) return rc;
        HS_CUDA_OK(cudaMemcpyAsync(host, g.point_offsets + (P - 1), sizeof(int), cudaMemcpyDeviceToHost, stream));
        HS_CUDA_OK(cudaStreamSynchronize(stream));
        *num_rendered = host[0];
        *binning_hint = -1;   // unknown: hs_forward_render will use the global sort
        return 0;
    }
    // tile-bucket binning: per-tile counts in the preprocess pass, tile scan -> ranges, cursors, count, longest list
    if (image_state == nullptr || (reinterpret_cast<uintptr_t>(image_state) & (HS_ALIGN - 1)) != 0) {
        set_error("image_state must be a %d-byte aligned device buffer", HS_ALIGN);
        return 1;
    }
    const size_t tiles = (size_t)cam.grid_x * cam.grid_y;
    ImageView img;
    image_view((char*)image_state, (size_t)cam.W * cam.H, tiles, &img);
    if (img.total_bytes > image_state_bytes + HS_ALIGN) {
        set_error("image_state too small");
        return 1;
    }
    HS_CUDA_OK(cudaMemsetAsync(img.tile_count, 0, sizeof(uint32_t) * tiles * HS_CTR_STRIDE, stream));
    int rc = launch_preprocess(P, means3D, scales, rotations, opacities, cov3D_precomp, cam, radii, g, img.tile_count,
                               stream, debug);
    if (rc) return rc;
    const bool async = (flags & HS_ASYNC_BINNING) != 0;
    const int tile_cap = cap_hint & HS_HINT_MASK;
    if (async && (cap_instances <= 0 || cap_hint < 0 || tile_cap <= 0 || tile_cap > HS_TILE_SORT_MAX)) {
        set_error("HS_ASYNC_BINNING: *num_rendered / *binning_hint must hold the capacities (instances > 0, 0 < longest "
                  "tile list <= %d)", HS_TILE_SORT_MAX);
        return 1;
    }
    rc = launch_tile_scan(cam, img, async ? (uint32_t)cap_instances : 0u, async ? (uint32_t)tile_cap : 0u, stream, debug);
    if (rc) return rc;
    if (shs != nullptr) rc = launch_sh_forward(P, sh_degree, sh_coeffs, means3D, c->campos, shs, radii, g, stream, debug);
    if (rc) return rc;
    if (async) {
        *num_rendered = cap_instances;
        *binning_hint = tile_cap;   // both sort classes may be needed; nothing is read back, nothing synchronises
        if (flags & HS_DEFER_READBACK) {
            // the counts travel to the host while the render kernels that the caller enqueues next are running
            Readback* rb = readback_slot();
            if (rb == nullptr) {
                set_error("HS_DEFER_READBACK: cannot allocate the pinned read-back slot");
                return 2;
            }
            HS_CUDA_OK(cudaMemcpyAsync(rb->host, img.info, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            HS_CUDA_OK(cudaEventRecord(rb->ev, stream));
        }
        return 0;
    }
    HS_CUDA_OK(cudaMemcpyAsync(host, img.info, 3 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    HS_CUDA_OK(cudaStreamSynchronize(stream));
    *num_rendered = host[0];
    // longest tile list, and whether any tile falls in the small sort class
    *binning_hint = (host[1] > HS_HINT_MASK ? HS_HINT_MASK : host[1]) | (host[2] == 0 ? HS_HINT_NO_SMALL : 0);
    return 0;
}

int hs_forward_readback(int counts[4]) {
    Readback* rb = readback_slot();
    if (rb == nullptr || counts == nullptr) {
        set_error("hs_forward_readback: no read-back pending on this thread");
        return 1;
    }
    HS_CUDA_OK(cudaEventSynchronize(rb->ev));
    for (int i = 0; i < 4; i++) counts[i] = rb->host[i];
    return 0;
}

int hs_forward_render(const hs_camera* c, int P, int S, int R, int binning_hint, const float* colors,
                      const float* semantics,
                      const int* radii, void* geom_state, size_t geom_state_bytes, void* binning_state, size_t binning_state_bytes,
                      void* image_state, size_t image_state_bytes, float* out_color, float* out_semantic,
                      float* out_depth, float* out_median_depth, float* out_opacity, float* out_mask, int flags,
                      void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Camera cam;
    if (make_camera(c, &cam)) return 1;
    if (!hs_supports_semantic_channels(S)) {
        set_error("semantic channel count S=%d is not instantiated (built: 0,16,26,32,48,64,74,102)", S);
        return 3;
    }
    if (image_state == nullptr || out_color == nullptr || out_depth == nullptr || out_median_depth == nullptr ||
        out_opacity == nullptr || (S > 0 && out_semantic == nullptr)) {
        set_error("hs_forward_render: NULL output");
        return 1;
    }
    const size_t N = (size_t)cam.W * cam.H;
    const size_t tiles = (size_t)cam.grid_x * cam.grid_y;
    ImageView img;
    image_view((char*)image_state, N, tiles, &img);
    if (img.total_bytes > image_state_bytes + HS_ALIGN) {
        set_error("image_state too small");
        return 1;
    }
    GeomView g;
    BinningView b;
    if (P > 0) {
        if (geom_view((char*)geom_state, (size_t)P, &g)) return 2;
        if (geom_state_bytes < hs_geom_state_bytes_rows(P, S) - HS_ALIGN) {
            set_error("geom_state too small for the packed rows of S=%d: %zu < %zu (size it with hs_geom_state_bytes_rows)", S,
                      geom_state_bytes, hs_geom_state_bytes_rows(P, S));
            return 1;
        }
    } else {
        g = GeomView{};
    }
    if (P <= 0 || R <= 0) HS_CUDA_OK(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * tiles, stream));
    if (R > 0) {
        if (binning_state == nullptr) {
            set_error("binning_state is NULL");
            return 1;
        }
        if (binning_view((char*)binning_state, (size_t)R, &b)) return 2;
        if (b.total_bytes > binning_state_bytes + HS_ALIGN) {
            set_error("binning_state too small: %zu < %zu", binning_state_bytes, b.total_bytes);
            return 1;
        }
        if (S > 0 && semantics == nullptr) {
            set_error("semantics must be provided when S > 0");
            return 1;
        }
        if (colors == nullptr) colors = g.rgb;   // SH colour path: evaluated by hs_forward_geometry
    } else {
        b = BinningView{};
    }
    const bool debug = c->debug != 0;
    int rc = 0;
    const int max_tile = binning_hint & HS_HINT_MASK;
    // The packed per-Gaussian records: the TMA gather source of the blend.  (Measured: running this 20 us pass on a side
    // stream next to the scatter / sort kernels hides it but slows them by as much -- both are memory-system bound.)
    if (P > 0 && R > 0) {
        rc = launch_pack_rows(P, radii, S, g, colors, semantics, stream);
        if (rc) return rc;
    }
    if (flags & HS_REUSE_BINNING) {
        // the sorted tile lists of this frame are already in binning_state / image_state
    } else if (binning_hint >= 0 && max_tile <= HS_TILE_SORT_MAX) {
        rc = launch_tile_binning(P, R, max_tile, (binning_hint & HS_HINT_NO_SMALL) ? 0 : 1, cam, radii, g, b, img, stream,
                                 debug);
    } else {
        // global radix sort: requested (HS_SORT_GLOBAL in hs_forward_geometry) or a tile list too long for the
        // shared-memory sort; in the second case the offsets scan has not run yet
        if (binning_hint >= 0 && P > 0) rc = launch_scan(P, g, stream, debug);
        if (rc) return rc;
        rc = launch_binning(P, R, cam, radii, g, b, img, stream, debug);
    }
    if (rc) return rc;
    return launch_blend_forward(P, radii, S, cam, g, b, img, colors, semantics, out_color, out_semantic, out_depth,
                                out_median_depth, out_opacity, out_mask, flags & ~HS_REUSE_BINNING, stream, debug);
}

int hs_backward(const hs_camera* c, int P, int S, int R, const float* means3D, const int* radii,
                const float* colors, const float* semantics, const float* scales, const float* rotations,
                const float* cov3D_precomp, const float* shs, int sh_degree, int sh_coeffs, const void* geom_state, const void* binning_state,
                const void* image_state, const float* dL_dout_color, const float* dL_dout_semantic,
                const float* dL_dout_depth, const float* dL_dout_median_depth, const float* dL_dout_opacity,
                float* dL_dmeans2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors, float* dL_dsemantics,
                float* dL_ddepths, float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales, float* dL_drotations,
                float* dL_dsh, const float* pose_points, float* dL_dpose, int flags, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    Camera cam;
    if (make_camera(c, &cam)) return 1;
    if (P <= 0) return 0;
    if (!hs_supports_semantic_channels(S)) {
        set_error("semantic channel count S=%d is not instantiated (built: 0,16,26,32,48,64,74,102)", S);
        return 3;
    }
    const size_t N = (size_t)cam.W * cam.H;
    const size_t tiles = (size_t)cam.grid_x * cam.grid_y;
    GeomView g;
    ImageView img;
    BinningView b = BinningView{};
    if (geom_view((char*)geom_state, (size_t)P, &g)) return 2;
    image_view((char*)image_state, N, tiles, &img);
    const bool debug = c->debug != 0;
    int rc = 0;
    if (dL_dpose != nullptr && pose_points == nullptr) {
        set_error("hs_backward: dL_dpose needs pose_points");
        return 1;
    }
    if (colors == nullptr) {
        if (shs == nullptr || dL_dsh == nullptr || c->campos == nullptr) {
            set_error("hs_backward: provide colors, or shs + dL_dsh + campos for the spherical-harmonics colour path");
            return 1;
        }
        colors = g.rgb;
    } else {
        shs = nullptr;
    }
    if (R > 0) {
        if (binning_view((char*)binning_state, (size_t)R, &b)) return 2;
        int kflags = 0;
        if (flags & HS_SEM_ALPHA_EXACT) kflags |= HS_FLAG_SEM_ALPHA_EXACT;
        if (flags & HS_BWD_SIMT) kflags |= HS_FLAG_BWD_SHUFFLE;
        rc = launch_blend_backward(S, cam, g, b, img, c->bg, colors, semantics, dL_dout_color, dL_dout_semantic,
                                   dL_dout_depth, dL_dout_median_depth, dL_dout_opacity, dL_dmeans2D, dL_dconic,
                                   dL_dopacity, dL_dcolors, dL_dsemantics, dL_ddepths, kflags, stream, debug);
        if (rc) return rc;
    }
    rc = launch_geom_backward(P, means3D, radii, scales, rotations, cov3D_precomp, cam, dL_dmeans2D, dL_dconic,
                              dL_ddepths, dL_dmeans3D, dL_dcov3D, cov3D_precomp ? nullptr : dL_dscales,
                              cov3D_precomp ? nullptr : dL_drotations, pose_points, dL_dpose, stream, debug);
    if (rc || shs == nullptr) return rc;
    // colour gradient -> SH coefficients, and the view-direction term added to dL/dmean (backward.cu:20-139)
    return launch_sh_backward(P, sh_degree, sh_coeffs, means3D, c->campos, shs, radii, g, dL_dcolors, dL_dmeans3D, dL_dsh,
                              pose_points, dL_dpose, stream, debug);
}

int hs_profile_enable(int on) {
    prof_enable(on != 0);
    return 0;
}
int hs_profile_read(float total_ms[8], int count[8]) { return prof_read(total_ms, count); }
long long hs_kernel_launch_count(void) { return launches(); }
long long hs_library_call_count(void) { return lib_calls(); }

int hs_masked_l1(const float* pred, const float* target, const unsigned char* mask, int channels, size_t pixels,
                 float* loss, float* grad, void* stream_) {
    if (pred == nullptr || target == nullptr || loss == nullptr || grad == nullptr) {
        set_error("hs_masked_l1: NULL argument");
        return 1;
    }
    return launch_masked_l1(pred, target, mask, channels, pixels, loss, grad, (cudaStream_t)stream_);
}

int hs_hier_cross_entropy(const float* sem, const int* labels, int levels, const int* level_begin,
                          const float* level_scale, size_t pixels, float* loss, float* grad, void* stream_) {
    if (sem == nullptr || labels == nullptr || level_begin == nullptr || level_scale == nullptr || loss == nullptr ||
        grad == nullptr) {
        set_error("hs_hier_cross_entropy: NULL argument");
        return 1;
    }
    return launch_hier_cross_entropy(sem, labels, levels, level_begin, level_scale, pixels, loss, grad,
                                     (cudaStream_t)stream_);
}

int hs_transform_points(const float* w2c, const float* world, int P, float* cam, void* stream_) {
    if (P > 0 && (w2c == nullptr || world == nullptr || cam == nullptr)) {
        set_error("hs_transform_points: NULL argument");
        return 1;
    }
    return launch_transform_points(w2c, world, P, cam, (cudaStream_t)stream_);
}

int hs_tracking_loss(const float* im, const float* depth, const float* silhouette, const float* gt_im,
                     const float* gt_depth, size_t pixels, float sil_thres, int use_silhouette, float depth_weight,
                     float im_weight, float* loss, float* grad_im, float* grad_depth, void* stream_) {
    if (im == nullptr || depth == nullptr || gt_im == nullptr || gt_depth == nullptr || loss == nullptr ||
        grad_im == nullptr || grad_depth == nullptr || (use_silhouette && silhouette == nullptr)) {
        set_error("hs_tracking_loss: NULL argument");
        return 1;
    }
    return launch_tracking_loss(im, depth, silhouette, gt_im, gt_depth, pixels, sil_thres, use_silhouette, depth_weight,
                                im_weight, loss, grad_im, grad_depth, (cudaStream_t)stream_);
}

int hs_pose_step(float* cam_rot, float* cam_tran, const float* dL_dpose, float* loss, float* state, float* w2c,
                 const unsigned int* binning_info, float lr_rot, float lr_tran, float beta1, float beta2, float eps, int mode,
                 void* stream_) {
    if (cam_rot == nullptr || cam_tran == nullptr || w2c == nullptr ||
        (mode != 0 && (dL_dpose == nullptr || loss == nullptr || state == nullptr))) {
        set_error("hs_pose_step: NULL argument");
        return 1;
    }
    return launch_pose_step(cam_rot, cam_tran, dL_dpose, loss, state, w2c, binning_info, lr_rot, lr_tran, beta1, beta2,
                            eps, mode, (cudaStream_t)stream_);
}

int hs_keyframe_overlap(const float* points, int num_points, const float* w2c, int keyframes, float fx, float fy, float cx,
                        float cy, int width, int height, int edge, int* counts, void* stream_) {
    if (keyframes > 0 && (w2c == nullptr || counts == nullptr || (num_points > 0 && points == nullptr))) {
        set_error("hs_keyframe_overlap: NULL argument");
        return 1;
    }
    return launch_keyframe_overlap(points, num_points, w2c, keyframes, fx, fy, cx, cy, width, height, edge, counts,
                                   (cudaStream_t)stream_);
}

int hs_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int segments,
                 const unsigned long long* segment_end, const double* segment_lr, double beta1, double beta2, double eps,
                 int step, void* stream_) {
    if (param == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || segment_end == nullptr ||
        segment_lr == nullptr) {
        set_error("hs_adam_step: NULL argument");
        return 1;
    }
    return launch_adam_flat(param, grad, exp_avg, exp_avg_sq, n, segments, segment_end, segment_lr, beta1, beta2, eps, step,
                            (cudaStream_t)stream_);
}

int hs_adam_step_device(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int segments,
                        const unsigned long long* segment_end, const double* segment_lr, double beta1, double beta2, double eps,
                        int* step_counter, float* scalars, void* stream_) {
    if (param == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr || segment_end == nullptr ||
        segment_lr == nullptr || step_counter == nullptr || scalars == nullptr) {
        set_error("hs_adam_step_device: NULL argument");
        return 1;
    }
    return launch_adam_flat(param, grad, exp_avg, exp_avg_sq, n, segments, segment_end, segment_lr, beta1, beta2, eps, 1,
                            (cudaStream_t)stream_, step_counter, scalars);
}

size_t hs_compact_scratch_bytes(int P) { return compact_scratch_bytes(P); }

int hs_compact_plan(const unsigned char* keep, int P, void* scratch, void* stream_) {
    if (P > 0 && (keep == nullptr || scratch == nullptr)) {
        set_error("hs_compact_plan: NULL argument");
        return 1;
    }
    return launch_compact_plan(keep, P, (unsigned*)scratch, (cudaStream_t)stream_);
}

int hs_compact_gather(const float* src, float* dst, const void* scratch, int P, int rows, int segments,
                      const unsigned long long* src_offset, const unsigned long long* dst_offset, const int* width,
                      void* stream_) {
    if (rows > 0 && (src == nullptr || dst == nullptr || scratch == nullptr || src_offset == nullptr ||
                     dst_offset == nullptr || width == nullptr)) {
        set_error("hs_compact_gather: NULL argument");
        return 1;
    }
    return launch_compact_gather(src, dst, (const unsigned*)scratch, P, rows, segments, src_offset, dst_offset, width,
                                 (cudaStream_t)stream_);
}

int hs_l1_ssim(const float* pred, const float* target, int channels, int height, int width, const float* window11,
               float l1_scale, float ssim_scale, float* loss, float* scratch, float* grad, void* stream_) {
    if (pred == nullptr || target == nullptr || window11 == nullptr || loss == nullptr || scratch == nullptr) {
        set_error("hs_l1_ssim: NULL argument");
        return 1;
    }
    return launch_l1_ssim(pred, target, channels, height, width, window11, l1_scale, ssim_scale, loss, scratch, grad,
                          (cudaStream_t)stream_);
}

int hs_leaf_cross_entropy(const float* sem, const int* labels, const float* weight, const float* bias, int channels,
                          int classes, size_t pixels, float scale, float* loss, float* lse, float* grad_sem,
                          int flags, float* grad_weight, float* grad_bias, void* stream_) {
    if (sem == nullptr || labels == nullptr || weight == nullptr || loss == nullptr || lse == nullptr ||
        grad_sem == nullptr) {
        set_error("hs_leaf_cross_entropy: NULL argument");
        return 1;
    }
    return launch_leaf_cross_entropy(sem, labels, weight, bias, channels, classes, pixels, scale, loss, lse, grad_sem,
                                     flags & HS_LEAF_ACCUMULATE, grad_weight, grad_bias, (flags & HS_LEAF_TF32) != 0,
                                     (cudaStream_t)stream_);
}

int hs_allreduce_sum(void* multicast_ptr, void* const* peer_buffers, void* const* peer_signal_pads, int rank, int world_size,
                     size_t count, unsigned int epoch, int blocks, void* stream_) {
    return launch_allreduce_sum(multicast_ptr, peer_buffers, peer_signal_pads, rank, world_size, count, epoch, blocks,
                                (cudaStream_t)stream_);
}

void hs_leaf_tc_debug(long long* stamps) { leaf_tc_set_debug(stamps); }
size_t hs_leaf_ce_workspace_bytes(int channels, int classes) { return leaf_tc_workspace_bytes(channels, classes); }

int hs_leaf_cross_entropy_tc(const float* sem, const int* labels, const float* weight, const float* bias, int channels,
                             int classes, size_t pixels, float scale, float* loss, float* lse, float* grad_sem, int flags,
                             float* grad_weight, float* grad_bias, void* workspace, size_t workspace_bytes, void* stream_) {
    if (sem == nullptr || labels == nullptr || weight == nullptr || loss == nullptr || lse == nullptr ||
        grad_sem == nullptr) {
        set_error("hs_leaf_cross_entropy_tc: NULL argument");
        return 1;
    }
    int rc = launch_leaf_cross_entropy_tc(sem, labels, weight, bias, channels, classes, pixels, scale, loss, lse, grad_sem,
                                          flags & HS_LEAF_ACCUMULATE, (float*)workspace, workspace_bytes,
                                          (cudaStream_t)stream_);
    if (rc || grad_weight == nullptr) return rc;
    return launch_leaf_weight_grad(sem, labels, weight, bias, lse, channels, classes, pixels, scale, grad_weight, grad_bias,
                                   (flags & HS_LEAF_TF32) != 0, (cudaStream_t)stream_);
}

int hs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    unsigned char* present, void* stream_) {
    return launch_mark_visible(P, means3D, viewmatrix, projmatrix, (bool*)present, (cudaStream_t)stream_, false);
}

}  // extern "C"

// The reference-side binding over the C ABI: a pybind11 / torch C++ extension that exports the five functions of the
// reference's native module (hierslam-diff-gaussian-rasterization-w-depth/ext.cpp:15-23, signatures and return tuples of
// rasterize_points.h:18-125) and implements each one with calls into libhsraster.so (include/hs_raster.h).  A maintainer
// who keeps the reference's Python layer (diff_gaussian_rasterization/__init__.py) links THIS file instead of
// rasterize_points.cu + cuda_rasterizer/*.
//
// The product ships the same glue in Python (hier_slam_b200/_C.py, ctypes), which additionally pads un-instantiated
// channel counts, renders S > 102 in passes, redirects gradients into registered sinks and speculates on the binning
// size; this file is the minimal synchronous form (one num_rendered read-back per forward, like the reference).
// Compile-checked by tests/test_binding_stub.py; build: python bindings/torch_ext/build.py.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <tuple>

#include "hs_raster.h"

namespace {

using torch::Tensor;

const float* fptr(const Tensor& t) { return t.defined() && t.numel() > 0 ? t.data_ptr<float>() : nullptr; }
float* fptr_mut(Tensor& t) { return t.defined() && t.numel() > 0 ? t.data_ptr<float>() : nullptr; }

Tensor f32c(const Tensor& t, const char* name) {
    if (!t.defined() || t.numel() == 0) return t;
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (libhsraster has no CPU path)");
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
    return t.contiguous();
}

struct Camera {
    Tensor view, proj, bg, campos;      // keep the contiguous copies alive for the duration of the call
    hs_camera cam;
    Camera(const Tensor& background, float scale_modifier, const Tensor& viewmatrix, const Tensor& projmatrix, float tan_fovx,
           float tan_fovy, int H, int W, const Tensor& campos_, bool prefiltered, bool debug)
        : view(f32c(viewmatrix, "viewmatrix")), proj(f32c(projmatrix, "projmatrix")), bg(f32c(background, "bg")),
          campos(f32c(campos_, "campos")) {
        cam.image_height = H;
        cam.image_width = W;
        cam.tanfovx = tan_fovx;
        cam.tanfovy = tan_fovy;
        cam.scale_modifier = scale_modifier;
        cam.viewmatrix = fptr(view);
        cam.projmatrix = fptr(proj);
        cam.bg = fptr(bg);
        cam.campos = fptr(campos);
        cam.prefiltered = prefiltered ? 1 : 0;
        cam.debug = debug ? 1 : 0;
    }
};

void check(int rc, const char* what) { TORCH_CHECK(rc == 0, what, ": ", hs_last_error()); }

struct Forward {
    int num_rendered = 0;
    Tensor color, semantic, depth, median, opacity, mask, radii, geom, binning, img;
};

// S = 0: the non-semantic variant (writes `mask`)
Forward forward(const Tensor& background, const Tensor& means3D, const Tensor& colors, const Tensor& semantics,
                const Tensor& opacity, const Tensor& scales, const Tensor& rotations, float scale_modifier,
                const Tensor& cov3D_precomp, const Tensor& viewmatrix, const Tensor& projmatrix, float tan_fovx, float tan_fovy,
                int H, int W, const Tensor& sh, int degree, const Tensor& campos, bool prefiltered, bool debug, bool semantic) {
    TORCH_CHECK(means3D.dim() == 2 && means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (libhsraster has no CPU path)");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = (int)means3D.size(0);
    const int S = semantic && semantics.numel() > 0 ? (int)semantics.size(1) : 0;
    TORCH_CHECK(hs_supports_semantic_channels(S), "semantic channel count ", S, " is not instantiated in libhsraster");
    auto f32 = means3D.options().dtype(torch::kFloat32);
    auto u8 = means3D.options().dtype(torch::kUInt8);
    Forward o;
    o.radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
    if (P == 0) {      // rasterize_points.cu:277-292: zero images, empty state
        o.color = torch::zeros({3, H, W}, f32);
        o.semantic = torch::zeros({S, H, W}, f32);
        o.depth = torch::zeros({1, H, W}, f32);
        o.median = torch::zeros({1, H, W}, f32);
        o.opacity = torch::zeros({1, H, W}, f32);
        o.mask = torch::zeros({1, H, W}, f32);
        o.geom = o.binning = o.img = torch::empty({0}, u8);
        return o;
    }
    const Tensor m3 = f32c(means3D, "means3D"), col = f32c(colors, "colors_precomp"), sem = f32c(semantics, "semantics_precomp"),
                 opa = f32c(opacity, "opacities"), sca = f32c(scales, "scales"), rot = f32c(rotations, "rotations"),
                 cov = f32c(cov3D_precomp, "cov3D_precomp"), shs = f32c(sh, "shs");
    const bool use_sh = col.numel() == 0;
    TORCH_CHECK(!use_sh || (shs.dim() == 3 && shs.size(0) == P && shs.size(2) == 3), "provide colors_precomp [P,3] or shs [P,M,3]");
    Camera c(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, campos, prefiltered, debug);
    void* stream = c10::cuda::getCurrentCUDAStream().stream();
    o.geom = torch::empty({(int64_t)hs_geom_state_bytes_rows(P, S)}, u8);
    o.img = torch::empty({(int64_t)hs_image_state_bytes(H, W)}, u8);
    int hint = 0;
    check(hs_forward_geometry(&c.cam, P, fptr(m3), fptr(opa), fptr(sca), fptr(rot), fptr(cov), use_sh ? fptr(shs) : nullptr,
                              degree, use_sh ? (int)shs.size(1) : 0, o.radii.data_ptr<int>(), o.geom.data_ptr(),
                              (size_t)o.geom.numel(), o.img.data_ptr(), (size_t)o.img.numel(), /*flags=*/0, &o.num_rendered, &hint,
                              stream),
          "hs_forward_geometry");
    o.binning = torch::empty({(int64_t)hs_binning_state_bytes(o.num_rendered)}, u8);
    o.color = torch::empty({3, H, W}, f32);          // every pixel is written: no fill
    o.semantic = torch::empty({S, H, W}, f32);
    o.depth = torch::empty({1, H, W}, f32);
    o.median = torch::empty({1, H, W}, f32);
    o.opacity = torch::empty({1, H, W}, f32);
    if (!semantic) o.mask = torch::empty({1, H, W}, f32);
    check(hs_forward_render(&c.cam, P, S, o.num_rendered, hint, use_sh ? nullptr : fptr(col), fptr(sem), o.radii.data_ptr<int>(),
                            o.geom.data_ptr(), (size_t)o.geom.numel(), o.binning.data_ptr(), (size_t)o.binning.numel(),
                            o.img.data_ptr(), (size_t)o.img.numel(), fptr_mut(o.color), fptr_mut(o.semantic), fptr_mut(o.depth),
                            fptr_mut(o.median), fptr_mut(o.opacity), semantic ? nullptr : fptr_mut(o.mask), /*flags=*/0, stream),
          "hs_forward_render");
    return o;
}

struct Backward {
    Tensor means2D, colors, semantics, opacity, means3D, cov3D, sh, scales, rotations;
};

Backward backward(const Tensor& background, const Tensor& means3D, const Tensor& radii, const Tensor& colors,
                  const Tensor& semantics, const Tensor& scales, const Tensor& rotations, float scale_modifier,
                  const Tensor& cov3D_precomp, const Tensor& viewmatrix, const Tensor& projmatrix, float tan_fovx, float tan_fovy,
                  const Tensor& dL_color, const Tensor& dL_semantic, const Tensor& dL_depth, const Tensor& dL_median,
                  const Tensor& dL_opacity, const Tensor& sh, int degree, const Tensor& campos, const Tensor& geom, int R,
                  const Tensor& binning, const Tensor& img, bool debug, bool semantic) {
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (libhsraster has no CPU path)");
    TORCH_CHECK(dL_color.defined() && dL_color.dim() == 3, "dL_dout_color [3,H,W] is required (it carries the image size)");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = (int)means3D.size(0);
    const int S = semantic && semantics.numel() > 0 ? (int)semantics.size(1) : 0;
    const int H = (int)dL_color.size(1), W = (int)dL_color.size(2);      // rasterize_points.cu:369-370
    const int M = sh.defined() && sh.numel() > 0 ? (int)sh.size(1) : 0;
    const bool use_sh = colors.numel() == 0 && M > 0;
    const bool have_scales = scales.defined() && scales.numel() > 0;
    auto f32 = means3D.options().dtype(torch::kFloat32);
    Backward g;
    // accumulated with atomics by the blend backward: zero-initialised; the rest is fully written
    g.means2D = torch::zeros({P, 3}, f32);
    Tensor conic = torch::zeros({P, 2, 2}, f32);
    g.opacity = torch::zeros({P, 1}, f32);
    g.colors = torch::zeros({P, 3}, f32);
    g.semantics = torch::zeros({P, S}, f32);
    Tensor depths = torch::zeros({P, 1}, f32);
    g.means3D = torch::empty({P, 3}, f32);
    g.cov3D = torch::empty({P, 6}, f32);
    g.scales = have_scales ? torch::empty({P, 3}, f32) : torch::zeros({P, 3}, f32);
    g.rotations = have_scales ? torch::empty({P, 4}, f32) : torch::zeros({P, 4}, f32);
    g.sh = use_sh ? torch::empty({P, M, 3}, f32) : torch::zeros({P, M, 3}, f32);
    if (P == 0) return g;
    const Tensor m3 = f32c(means3D, "means3D"), col = f32c(colors, "colors_precomp"), sem = f32c(semantics, "semantics_precomp"),
                 sca = f32c(scales, "scales"), rot = f32c(rotations, "rotations"), cov = f32c(cov3D_precomp, "cov3D_precomp"),
                 shs = f32c(sh, "shs"), gc = f32c(dL_color, "dL_dout_color"), gs = f32c(dL_semantic, "dL_dout_semantic"),
                 gd = f32c(dL_depth, "dL_dout_depth"), gm = f32c(dL_median, "dL_dout_median_depth"),
                 go = f32c(dL_opacity, "dL_dout_final_opacity");
    Camera c(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, campos, false, debug);
    void* stream = c10::cuda::getCurrentCUDAStream().stream();
    check(hs_backward(&c.cam, P, S, R, fptr(m3), radii.data_ptr<int>(), use_sh ? nullptr : fptr(col), fptr(sem), fptr(sca),
                      fptr(rot), fptr(cov), use_sh ? fptr(shs) : nullptr, degree, M, geom.data_ptr(), binning.data_ptr(),
                      img.data_ptr(), fptr(gc), S ? fptr(gs) : nullptr, fptr(gd), fptr(gm), fptr(go), fptr_mut(g.means2D),
                      fptr_mut(conic), fptr_mut(g.opacity), fptr_mut(g.colors), fptr_mut(g.semantics), fptr_mut(depths),
                      fptr_mut(g.means3D), fptr_mut(g.cov3D), have_scales ? fptr_mut(g.scales) : nullptr,
                      have_scales ? fptr_mut(g.rotations) : nullptr, use_sh ? fptr_mut(g.sh) : nullptr,
                      /*pose_points=*/nullptr, /*dL_dpose=*/nullptr, /*flags=*/0, stream),
          "hs_backward");
    return g;
}

}  // namespace

// ---- the reference's five entry points (ext.cpp:15-23) --------------------------------------------------------------

std::tuple<int, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> rasterize_gaussians_semantic(
    const Tensor& background, const Tensor& means3D, const Tensor& colors, const Tensor& semantics, const Tensor& opacity,
    const Tensor& scales, const Tensor& rotations, const float scale_modifier, const Tensor& cov3D_precomp,
    const Tensor& viewmatrix, const Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
    const int image_width, const Tensor& sh, const int degree, const Tensor& campos, const bool prefiltered, const bool debug) {
    Forward o = forward(background, means3D, colors, semantics, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered,
                        debug, true);
    return std::make_tuple(o.num_rendered, o.color, o.semantic, o.depth, o.median, o.opacity, o.radii, o.geom, o.binning, o.img);
}

std::tuple<int, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> rasterize_gaussians(
    const Tensor& background, const Tensor& means3D, const Tensor& colors, const Tensor& opacity, const Tensor& scales,
    const Tensor& rotations, const float scale_modifier, const Tensor& cov3D_precomp, const Tensor& viewmatrix,
    const Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height, const int image_width,
    const Tensor& sh, const int degree, const Tensor& campos, const bool prefiltered, const bool debug) {
    Forward o = forward(background, means3D, colors, Tensor(), opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered,
                        debug, false);
    return std::make_tuple(o.num_rendered, o.color, o.depth, o.median, o.opacity, o.mask, o.radii, o.geom, o.binning, o.img);
}

std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> rasterize_gaussians_backward_semantic(
    const Tensor& background, const Tensor& means3D, const Tensor& radii, const Tensor& colors, const Tensor& semantics,
    const Tensor& scales, const Tensor& rotations, const float scale_modifier, const Tensor& cov3D_precomp,
    const Tensor& viewmatrix, const Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const Tensor& dL_dout_color,
    const Tensor& dL_dout_semantic, const Tensor& dL_dout_depth, const Tensor& dL_dout_median_depth,
    const Tensor& dL_dout_final_opacity, const Tensor& sh, const int degree, const Tensor& campos, const Tensor& geomBuffer,
    const int R, const Tensor& binningBuffer, const Tensor& imageBuffer, const bool debug) {
    Backward g = backward(background, means3D, radii, colors, semantics, scales, rotations, scale_modifier, cov3D_precomp,
                          viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_semantic, dL_dout_depth,
                          dL_dout_median_depth, dL_dout_final_opacity, sh, degree, campos, geomBuffer, R, binningBuffer,
                          imageBuffer, debug, true);
    return std::make_tuple(g.means2D, g.colors, g.semantics, g.opacity, g.means3D, g.cov3D, g.sh, g.scales, g.rotations);
}

std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> rasterize_gaussians_backward(
    const Tensor& background, const Tensor& means3D, const Tensor& radii, const Tensor& colors, const Tensor& scales,
    const Tensor& rotations, const float scale_modifier, const Tensor& cov3D_precomp, const Tensor& viewmatrix,
    const Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const Tensor& dL_dout_color,
    const Tensor& dL_dout_depth, const Tensor& dL_dout_median_depth, const Tensor& dL_dout_final_opacity, const Tensor& sh,
    const int degree, const Tensor& campos, const Tensor& geomBuffer, const int R, const Tensor& binningBuffer,
    const Tensor& imageBuffer, const bool debug) {
    Backward g = backward(background, means3D, radii, colors, Tensor(), scales, rotations, scale_modifier, cov3D_precomp,
                          viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, Tensor(), dL_dout_depth,
                          dL_dout_median_depth, dL_dout_final_opacity, sh, degree, campos, geomBuffer, R, binningBuffer,
                          imageBuffer, debug, false);
    return std::make_tuple(g.means2D, g.colors, g.opacity, g.means3D, g.cov3D, g.sh, g.scales, g.rotations);
}

Tensor mark_visible(Tensor& means3D, Tensor& viewmatrix, Tensor& projmatrix) {
    TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor (libhsraster has no CPU path)");
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = (int)means3D.size(0);
    Tensor present = torch::zeros({P}, means3D.options().dtype(torch::kBool));
    if (P == 0) return present;
    const Tensor m3 = f32c(means3D, "means3D"), v = f32c(viewmatrix, "viewmatrix"), p = f32c(projmatrix, "projmatrix");
    check(hs_mark_visible(P, fptr(m3), fptr(v), fptr(p), reinterpret_cast<unsigned char*>(present.data_ptr<bool>()),
                          c10::cuda::getCurrentCUDAStream().stream()),
          "hs_mark_visible");
    return present;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rasterize_gaussians", &rasterize_gaussians);
    m.def("rasterize_gaussians_semantic", &rasterize_gaussians_semantic);
    m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward);
    m.def("rasterize_gaussians_backward_semantic", &rasterize_gaussians_backward_semantic);
    m.def("mark_visible", &mark_visible);
}

"""Boundary acceptance (north_star: "scripts/hierslam.py tracking and mapping run unchanged"):
the reference's OWN caller code — scripts/hierslam.py get_loss_semantic (:715-853), get_loss_semantic_mlp (:856-1107),
add_new_gaussians_semantic_newrender (:1307-1352), utils/recon_helpers.py setup_camera (:4-28), utils/slam_helpers.py
transform_to_frame / transformed_params2rendervar_semantic (:195-219, :278-330) — is imported unchanged from
oracle/_ref/callers (a copy made by oracle/build_ref.sh; git-ignored) and executed twice on the same synthetic SLAM state:
once with `diff_gaussian_rasterization` bound to the UNMODIFIED reference CUDA build (oracle/_ref/S26) and once bound to
this repository's module.  Losses, every params[k].grad, means2D.grad and the bookkeeping variables must agree.  A second
test runs the reference's own diff_gaussian_rasterization/__init__.py over this repository's `_C` (the `from . import _C`
swap of __init__.py:15).
"""
import math
import types

import pytest
import torch

import parity_tools as pt
from hier_slam_b200.scene import CONFIGS, make_scene
from oracle import ref_callers, ref_loader

LEVELS = [4, 5, 5, 6, 6]        # synthetic Replica tree: 26 channels, 102 leaf classes
N_LEAF = 102


def test_reference_callers_import_against_this_module():
    """CPU: the reference's hierslam.py / utils import cleanly with `diff_gaussian_rasterization` = this repo's package."""
    if not ref_callers.available():
        pytest.skip("oracle/_ref/callers not present (run oracle/build_ref.sh where /root/reference exists)")
    import diff_gaussian_rasterization as dgr
    ns = ref_callers.load_callers(dgr, "cpu_import")
    assert ns.hierslam.Renderer_semantic is dgr.GaussianRasterizer_semantic
    assert ns.hierslam.Renderer is dgr.GaussianRasterizer
    assert ns.recon_helpers.Camera is dgr.GaussianRasterizationSettings
    for fn in ("get_loss_semantic", "get_loss_semantic_mlp", "add_new_gaussians_semantic_newrender", "get_loss"):
        assert callable(getattr(ns.hierslam, fn))


def _slam_state(ns, cfg, seed=0, frames=3):
    """params / variables as scripts/hierslam.py:361-409 builds them, from the synthetic scene of SURVEY.md section 8d"""
    sc = make_scene(cfg, seed, num_semantic=sum(LEVELS), device="cuda")
    P = sc["means3D"].shape[0]
    g = torch.Generator().manual_seed(seed + 100)
    rots = torch.tensor([1.0, 0.0, 0.0, 0.0]).repeat(1, 1)[:, :, None].repeat(1, 1, frames)
    rots = rots + 0.01 * torch.randn(1, 4, frames, generator=g)
    trans = 0.02 * torch.randn(1, 3, frames, generator=g)
    raw = {
        "means3D": sc["means3D"], "rgb_colors": sc["colors_precomp"],
        "unnorm_rotations": sc["rotations"] * 1.7,                        # un-normalised on purpose
        "logit_opacities": torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)),
        "log_scales": torch.log(sc["scales"][:, :1]), "semantic": sc["semantics_precomp"],
        "cam_unnorm_rots": rots, "cam_trans": trans,
    }
    params = {k: torch.nn.Parameter(v.cuda().float().contiguous().requires_grad_(True)) for k, v in raw.items()}
    variables = {k: torch.zeros(P, device="cuda") for k in ("max_2D_radius", "means2D_gradient_accum", "denom", "timestep")}
    return params, variables


def _curr_data(ns, cfg, seed=0):
    """cam through the reference's own setup_camera; colour / depth / labels are seeded synthetic images"""
    g = torch.Generator().manual_seed(seed + 200)
    H, W = cfg.height, cfg.width
    k = [[cfg.fx, 0.0, cfg.cx], [0.0, cfg.fy, cfg.cy], [0.0, 0.0, 1.0]]
    cam = ns.recon_helpers.setup_camera(W, H, k, torch.eye(4).numpy())
    depth = 0.5 + 5.5 * torch.rand(1, H, W, generator=g)
    depth[:, : H // 8] = 0.0                                              # invalid-depth band: exercises the masks
    labels = torch.stack([torch.randint(0, n, (H, W), generator=g) for n in LEVELS + [N_LEAF]]).float()
    return {"cam": cam, "im": torch.rand(3, H, W, generator=g).cuda(), "depth": depth.cuda(), "id": 1,
            "intrinsics": torch.tensor(k).cuda(), "semantic_label_gt": labels.cuda(), "iter_mapping": 20,
            "w2c": torch.eye(4).numpy()}


def _dataset():
    return types.SimpleNamespace(num_semantic=list(LEVELS), num_semantic_class=N_LEAF, dataset_name="replica_semantic",
                                 sem_mode="tree")


def _bindings():
    ref = ref_loader.load_reference(26)
    if ref is None or not ref_callers.available():
        pytest.skip("oracle/_ref (reference build + callers) not available on this box")
    import diff_gaussian_rasterization as dgr
    return ref_callers.load_callers(ref, "ref"), ref_callers.load_callers(dgr, "ours")


def _rel(a, b):
    return pt.grad_err(a, b)[0]


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["c1", "c2-60k"])
def test_tracking_loss_through_the_reference_callers(key):
    """get_loss_semantic(tracking=True) (scripts/hierslam.py:1837) + loss.backward(): pose gradients only"""
    ns_ref, ns_new = _bindings()
    cfg = CONFIGS["c1"] if key == "c1" else CONFIGS["c2"]._replace(num_gaussians=60_000)
    res = []
    for ns in (ns_ref, ns_new):
        params, variables = _slam_state(ns, cfg)
        data = _curr_data(ns, cfg)
        loss, variables, wl = ns.hierslam.get_loss_semantic(
            _dataset(), params, data, variables, 1, dict(im=0.5, depth=1.0), True, 0.99, True, False, None,
            tracking=True)
        loss.backward()
        torch.cuda.synchronize()
        res.append((loss.detach(), params, variables, wl))
    (l0, p0, v0, w0), (l1, p1, v1, w1) = res
    assert math.isfinite(float(l0)) and float(l0) > 0
    assert abs(float(l1) - float(l0)) <= 1e-4 * abs(float(l0)), (float(l0), float(l1))
    for k in ("depth", "im"):
        assert abs(float(w1[k]) - float(w0[k])) <= 1e-4 * abs(float(w0[k]))
    for k in ("cam_unnorm_rots", "cam_trans"):
        assert float(p0[k].grad.abs().max()) > 0
        assert _rel(p1[k].grad, p0[k].grad) < 1e-3, (k, _rel(p1[k].grad, p0[k].grad))
    assert torch.equal(v0["seen"], v1["seen"]) and torch.equal(v0["max_2D_radius"], v1["max_2D_radius"])
    assert _rel(v1["means2D"].grad, v0["means2D"].grad) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("mlp", [False, True])
def test_mapping_loss_through_the_reference_callers(mlp):
    """get_loss_semantic / get_loss_semantic_mlp(mapping=True) (scripts/hierslam.py:2022-2028) + loss.backward():
    every Gaussian parameter's gradient, the 1x1-conv gradients, and the bookkeeping"""
    ns_ref, ns_new = _bindings()
    cfg = CONFIGS["c1"]
    weights = dict(im=0.5, depth=1.0, sem=0.2)
    res = []
    for ns in (ns_ref, ns_new):
        torch.manual_seed(5)
        params, variables = _slam_state(ns, cfg)
        data = _curr_data(ns, cfg)
        if mlp:
            conv = torch.nn.Conv2d(sum(LEVELS), N_LEAF, kernel_size=1).cuda()
            loss, variables, wl = ns.hierslam.get_loss_semantic_mlp(
                _dataset(), params, data, variables, 1, weights, False, 0.5, True, False, None, conv, mapping=True)
        else:
            conv = None
            loss, variables, wl = ns.hierslam.get_loss_semantic(
                _dataset(), params, data, variables, 1, weights, False, 0.5, True, False, None, mapping=True)
        loss.backward()
        torch.cuda.synchronize()
        res.append((loss.detach(), params, variables, wl, conv))
    (l0, p0, v0, w0, c0), (l1, p1, v1, w1, c1) = res
    assert abs(float(l1) - float(l0)) <= 1e-4 * abs(float(l0)), (float(l0), float(l1))
    for k in ("depth", "im", "sem"):
        assert abs(float(w1[k]) - float(w0[k])) <= 1e-4 * abs(float(w0[k])), k
    for k in ("means3D", "rgb_colors", "logit_opacities", "log_scales", "semantic"):
        assert p0[k].grad is not None and float(p0[k].grad.abs().max()) > 0, k
        assert _rel(p1[k].grad, p0[k].grad) < 1e-3, (k, _rel(p1[k].grad, p0[k].grad))
    # Hier-SLAM's Gaussians are isotropic (scales = exp(tile(log_scales, (1, 3))), utils/slam_helpers.py:214): Sigma = s^2 I
    # for every rotation, so dL/d(unnorm_rotations) is identically zero in exact arithmetic and what both implementations
    # return is rounding noise.  It must be negligible next to the scale gradient that flows through the same Sigma.
    for p in (p0, p1):
        assert p["unnorm_rotations"].grad is not None
        assert float(p["unnorm_rotations"].grad.norm()) <= 1e-4 * float(p["log_scales"].grad.norm()), \
            (float(p["unnorm_rotations"].grad.norm()), float(p["log_scales"].grad.norm()))
    for k in ("cam_unnorm_rots", "cam_trans"):      # mapping without BA: the pose is detached (hierslam.py:742-745)
        assert p0[k].grad is None and p1[k].grad is None
    if mlp:
        for a, b in ((c1.weight.grad, c0.weight.grad), (c1.bias.grad, c0.bias.grad)):
            assert _rel(a, b) < 1e-3
    assert torch.equal(v0["seen"], v1["seen"]) and torch.equal(v0["max_2D_radius"], v1["max_2D_radius"])
    assert _rel(v1["means2D"].grad, v0["means2D"].grad) < 1e-3


@pytest.mark.gpu
def test_densification_render_through_the_reference_callers():
    """add_new_gaussians_semantic_newrender (scripts/hierslam.py:1307-1352): forward-only render, silhouette / depth
    masks, new Gaussians appended — the resulting maps must be identical in size and equal in value"""
    ns_ref, ns_new = _bindings()
    cfg = CONFIGS["c1"]
    outs = []
    for ns in (ns_ref, ns_new):
        torch.manual_seed(7)
        params, variables = _slam_state(ns, cfg)
        data = _curr_data(ns, cfg)
        params, variables = ns.hierslam.add_new_gaussians_semantic_newrender(       # call site: hierslam.py:1948-1950
            params, variables, data, 0.5, 1, "projective", sum(LEVELS))
        outs.append((params, variables))
    (p0, v0), (p1, v1) = outs
    assert p0["means3D"].shape == p1["means3D"].shape and p0["means3D"].shape[0] > cfg.num_gaussians
    for k in ("means3D", "rgb_colors", "log_scales", "logit_opacities", "unnorm_rotations"):
        assert torch.allclose(p0[k], p1[k], rtol=1e-5, atol=1e-6), k
    assert torch.equal(v0["timestep"], v1["timestep"])


@pytest.mark.gpu
def test_reference_python_layer_over_this_C_module():
    """the reference's own __init__.py (autograd functions, argument checks, tuple order) with
    `from . import _C` (__init__.py:15) resolving to hier_slam_b200._C must give the reference build's results."""
    ref = ref_loader.load_reference(26)
    if ref is None:
        pytest.skip("oracle/_ref/S26 not available on this box")
    from hier_slam_b200 import _C
    hybrid = ref_callers.load_reference_init_over(_C, 26, "ours")
    cfg = CONFIGS["c1"]
    sc = make_scene(cfg, 3, device="cuda")
    from hier_slam_b200.scene import upstream_grads
    ug = upstream_grads(cfg, 4, device="cuda")

    def run(mod, semantic):
        settings = pt.make_settings(mod.GaussianRasterizationSettings, cfg)
        leaf = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
        means2D = torch.zeros_like(leaf["means3D"], requires_grad=True) + 0
        means2D.retain_grad()
        kw = dict(means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"],
                  colors_precomp=leaf["colors_precomp"], scales=leaf["scales"], rotations=leaf["rotations"])
        if semantic:
            color, radii, sem, depth, median, opac = mod.GaussianRasterizer_semantic(raster_settings=settings)(
                semantics_precomp=leaf["semantics_precomp"], **kw)
            loss = (sem * ug["semantic"]).sum()
        else:
            color, radii, depth, median, opac, mask = mod.GaussianRasterizer(raster_settings=settings)(**kw)
            loss = (mask * ug["final_opacity"]).sum() * 0
        loss = loss + (color * ug["color"]).sum() + (depth * ug["depth"]).sum()
        loss.backward()
        grads = {k: v.grad for k, v in leaf.items() if v.grad is not None}
        grads["means2D"] = means2D.grad
        return dict(color=color, depth=depth, median=median, opacity=opac, radii=radii), grads

    for semantic in (True, False):
        o_h, g_h = run(hybrid, semantic)
        o_r, g_r = run(ref, semantic)
        assert torch.equal(o_h["radii"], o_r["radii"])
        for k in ("color", "depth", "median", "opacity"):
            mx, viol = pt.image_err(o_h[k], o_r[k])
            assert viol == 0, (k, mx)
        assert set(g_h) == set(g_r)
        for k in g_r:
            assert _rel(g_h[k], g_r[k]) < 1e-3, (k, semantic)
    vis = hybrid.GaussianRasterizer_semantic(pt.make_settings(hybrid.GaussianRasterizationSettings, cfg)).markVisible(
        sc["means3D"])
    assert torch.equal(vis, sc["means3D"][:, 2] > 0.2)

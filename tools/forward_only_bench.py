"""Densification render at the c2 / c5 shapes: the full semantic forward the reference runs
(scripts/hierslam.py:1307-1352) vs hier_slam_b200.rasterizer.render_depth_silhouette.  CUDA events, one JSON line each."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
import diff_gaussian_rasterization as ours
from hier_slam_b200.rasterizer import render_depth_silhouette
from hier_slam_b200.scene import CONFIGS, make_scene


def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for key in ("c2", "c5"):
    cfg = CONFIGS[key]
    sc = make_scene(cfg, 0, device="cuda")
    settings = pt.make_settings(ours.GaussianRasterizationSettings, cfg, "cuda")
    raster = ours.GaussianRasterizer_semantic(settings)
    m2d = torch.zeros_like(sc["means3D"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}        # parameters require grad, as in the SLAM loop
    full = lambda: raster(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                          colors_precomp=leaves["colors_precomp"], scales=leaves["scales"], rotations=leaves["rotations"],
                          semantics_precomp=leaves["semantics_precomp"])
    fast = lambda: render_depth_silhouette(settings, leaves["means3D"], leaves["opacities"], leaves["scales"],
                                           leaves["rotations"])
    print(json.dumps({"config": key, "semantic_forward_ms": round(timed(full), 3),
                      "render_depth_silhouette_ms": round(timed(fast), 3)}))

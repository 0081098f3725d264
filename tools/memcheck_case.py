"""Small forward+backward sweep for compute-sanitizer (tiny / small scenes, every instantiated S, SH colours, both binning
paths, the fused pose step).  usage: compute-sanitizer --tool memcheck python tools/memcheck_case.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
from hier_slam_b200 import _C
from hier_slam_b200.rasterizer import GaussianRasterizationSettings
from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads
import diff_gaussian_rasterization as dgr
from hier_slam_b200.tracking import PoseRasterizer_semantic

for key in ("tiny", "small"):
    cfg = CONFIGS[key]
    st = pt.make_settings(GaussianRasterizationSettings, cfg)
    for S in (0, 16, 26, 74, 102):
        sc = make_scene(cfg, 3, num_semantic=max(S, 1), device="cuda")
        ug = upstream_grads(cfg, 4, num_semantic=max(S, 1), device="cuda")
        sem = S > 0
        if not sem:
            sc.pop("semantics_precomp"); ug["semantic"] = None
        for glob in (False, True):
            _C.SORT_GLOBAL = glob
            f = pt.run_forward(_C, st, sc, sem)
            g = pt.run_backward(_C, st, sc, f, ug, sem)
            g2 = pt.run_backward(_C, st, sc, f, {k: (v if k in ("color", "depth") else None) for k, v in ug.items()}, sem,
                                 materialize=False)
        _C.SORT_GLOBAL = False
    # SH colours + fused pose step
    sc = make_scene(cfg, 5, device="cuda")
    P = sc["means3D"].shape[0]
    shs = torch.randn(P, 16, 3, device="cuda").requires_grad_(True)
    st3 = st._replace(sh_degree=3, campos=torch.zeros(3, device="cuda"))
    out = dgr.GaussianRasterizer_semantic(st3)(means3D=sc["means3D"].clone().requires_grad_(True), means2D=torch.zeros(P, 3, device="cuda"),
                                               opacities=sc["opacities"], shs=shs, scales=sc["scales"], rotations=sc["rotations"],
                                               semantics_precomp=sc["semantics_precomp"])
    (out[0].sum() + out[3].sum()).backward()
    w2c = torch.eye(4, device="cuda").requires_grad_(True)
    out = PoseRasterizer_semantic(st)(w2c, sc["means3D"], torch.zeros(P, 3, device="cuda"), sc["opacities"], sc["colors_precomp"],
                                      sc["scales"], sc["rotations"], sc["semantics_precomp"])
    (out[0].sum() + out[3].sum()).backward()
torch.cuda.synchronize()
print("memcheck sweep done")

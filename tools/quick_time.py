"""Quick per-stage timing of the new path on one scene (tools/quick_time.py [config] [iters])."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
from hier_slam_b200 import _C, _lib
from hier_slam_b200.rasterizer import GaussianRasterizationSettings
from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads
key = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = CONFIGS[key]
scene = make_scene(cfg, 0, device="cuda"); grads = upstream_grads(cfg, 1, device="cuda")
settings = pt.make_settings(GaussianRasterizationSettings, cfg)
lib = _lib.load()
for simt in (False, True):
    _C.BWD_SIMT = simt
    for _ in range(3):
        f = pt.run_forward(_C, settings, scene); g = pt.run_backward(_C, settings, scene, f, grads)
    torch.cuda.synchronize()
    lib.hs_profile_enable(1); _lib.profile_read()
    for _ in range(iters):
        f = pt.run_forward(_C, settings, scene); g = pt.run_backward(_C, settings, scene, f, grads)
    torch.cuda.synchronize()
    prof = _lib.profile_read(); lib.hs_profile_enable(0)
    print(json.dumps({"config": key, "bwd_simt": simt, **{k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items()}}))
_C.BWD_SIMT = False
# Hier-SLAM's own gradient pattern: colour, semantics, depth; median depth and silhouette receive no gradient
gh = {k: (v if k in ("color", "semantic", "depth") else None) for k, v in grads.items()}
for _ in range(3):
    pt.run_backward(_C, settings, scene, f, gh, materialize=False)
torch.cuda.synchronize()
lib.hs_profile_enable(1); _lib.profile_read()
for _ in range(iters):
    g_h = pt.run_backward(_C, settings, scene, f, gh, materialize=False)
torch.cuda.synchronize()
prof = _lib.profile_read(); lib.hs_profile_enable(0)
print(json.dumps({"config": key, "upstream": "colour+semantics+depth only", **{k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items()}}))
gz = {k: (v if k in ("color", "semantic", "depth") else torch.zeros_like(v)) for k, v in grads.items()}
g_z = pt.run_backward(_C, settings, scene, f, gz)
print("fast-T vs exact-T:", {k: pt.grad_err(g_h[k], g_z[k])[0] for k in g_h})
gm = pt.run_backward(_C, settings, scene, f, grads)
_C.BWD_SIMT = True
gs = pt.run_backward(_C, settings, scene, f, grads)
print({k: pt.grad_err(gm[k], gs[k])[0] for k in gm})

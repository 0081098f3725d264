"""Host-side time of one bench step, split into Python and native calls (perf_counter; GPU left asynchronous)."""
import os, sys, time, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
from hier_slam_b200 import _C, _lib
from hier_slam_b200.mapping import FlatParams
from hier_slam_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer_semantic
from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads
cfg = CONFIGS["c2"]; dev = "cuda"
sc = make_scene(cfg, 0, device=dev); up = upstream_grads(cfg, 1, device=dev)
params = FlatParams(sc); raster = GaussianRasterizer_semantic(pt.make_settings(GaussianRasterizationSettings, cfg, dev))
m2d = torch.zeros(cfg.num_gaussians, 3, device=dev)
lib = _lib.load()
T = collections.Counter()
def wrap(name):
    f = getattr(lib, name)
    def g(*a):
        t0 = time.perf_counter(); r = f(*a); T[name] += time.perf_counter() - t0; return r
    setattr(lib, name, g)
for n in ("hs_forward_geometry", "hs_forward_render", "hs_backward", "hs_geom_state_bytes", "hs_image_state_bytes", "hs_binning_state_bytes"):
    wrap(n)
for fn in ("_forward", "_backward"):
    f = getattr(_C, fn)
    def mk(f, fn):
        def g(*a, **k):
            t0 = time.perf_counter(); r = f(*a, **k); T[fn] += time.perf_counter() - t0; return r
        return g
    setattr(_C, fn, mk(f, fn))
def step():
    t0 = time.perf_counter()
    params.zero_grad(); lv = params.leaves
    o = raster(means3D=lv["means3D"], means2D=m2d, opacities=lv["opacities"], colors_precomp=lv["colors_precomp"],
               scales=lv["scales"], rotations=lv["rotations"], semantics_precomp=lv["semantics_precomp"])
    t1 = time.perf_counter()
    torch.autograd.backward((o[0], o[2], o[3], o[4], o[5]), (up["color"], up["semantic"], up["depth"], up["median_depth"], up["final_opacity"]))
    t2 = time.perf_counter()
    T["fwd_total"] += t1 - t0; T["bwd_total"] += t2 - t1
for _ in range(5): step()
torch.cuda.synchronize(); T.clear()
N = 50
t0 = time.perf_counter()
for _ in range(N): step()
host = time.perf_counter() - t0
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"host loop {host/N*1e6:.0f} us/step, wall incl. final sync {wall/N*1e6:.0f} us/step")
for k, v in sorted(T.items(), key=lambda kv: -kv[1]): print(f"{v/N*1e6:8.1f} us  {k}")

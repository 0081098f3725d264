"""GPU timeline of the bench step (torch.profiler): per-kernel time, idle gaps between kernels, per step."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
from hier_slam_b200.mapping import FlatParams
from hier_slam_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer_semantic
from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads
from torch.profiler import profile, ProfilerActivity
cfg = CONFIGS["c2"]; dev = "cuda"
sc = make_scene(cfg, 0, device=dev); up = upstream_grads(cfg, 1, device=dev)
params = FlatParams(sc); raster = GaussianRasterizer_semantic(pt.make_settings(GaussianRasterizationSettings, cfg, dev))
m2d = torch.zeros(cfg.num_gaussians, 3, device=dev)
def step():
    params.zero_grad(); lv = params.leaves
    o = raster(means3D=lv["means3D"], means2D=m2d, opacities=lv["opacities"], colors_precomp=lv["colors_precomp"],
               scales=lv["scales"], rotations=lv["rotations"], semantics_precomp=lv["semantics_precomp"])
    torch.autograd.backward((o[0], o[2], o[3], o[4], o[5]), (up["color"], up["semantic"], up["depth"], up["median_depth"], up["final_opacity"]))
for _ in range(5): step()
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N): step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
span = ev[-1].time_range.end - ev[0].time_range.start
busy = 0; gaps = {}
last_end = ev[0].time_range.start; last_name = "start"
import collections
kt = collections.Counter()
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    kt[e.name[:50]] += (t - s)
    if s > last_end:
        gaps[(last_name[:40], e.name[:40])] = gaps.get((last_name[:40], e.name[:40]), 0) + (s - last_end)
    if t > last_end:
        busy += t - max(s, last_end); last_end = t; last_name = e.name
print(f"span per step {span/N:.1f} us, busy {busy/N:.1f} us, idle {(span-busy)/N:.1f} us")
for k, v in kt.most_common(14): print(f"{v/N:9.1f} us  {k}")
print("largest idle gaps (per step):")
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1])[:8]: print(f"{v/N:8.1f} us  after {k[0]} -> before {k[1]}")

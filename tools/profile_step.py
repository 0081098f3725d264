"""Short fwd+bwd loop of the bench workload for ncu (tools/profile_step.py [config] [steps])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt  # noqa: E402
from hier_slam_b200 import _C  # noqa: E402
from hier_slam_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402
from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = CONFIGS[key]
scene = make_scene(cfg, 0, device="cuda")
grads = upstream_grads(cfg, 1, device="cuda")
settings = pt.make_settings(GaussianRasterizationSettings, cfg)
for _ in range(steps):
    f = pt.run_forward(_C, settings, scene)
    g = pt.run_backward(_C, settings, scene, f, grads)
torch.cuda.synchronize()
print("done", f["R"])

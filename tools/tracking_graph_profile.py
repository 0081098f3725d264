"""A few GraphedTracker iterations at the c2 shape (for an ncu launch list of one graphed tracking iteration)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
import diff_gaussian_rasterization as ours
from hier_slam_b200.scene import CONFIGS, keyframe_poses, make_scene
from hier_slam_b200.tracking import GraphedTracker
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
sc = make_scene(cfg, 0, device="cuda")
settings = pt.make_settings(ours.GaussianRasterizationSettings, cfg, "cuda")
raster = ours.GaussianRasterizer_semantic(settings)
pose = keyframe_poses(1, seed=2, max_angle_deg=1.0, max_trans=0.02).to("cuda")[0]
with torch.no_grad():
    tp = torch.addmm(pose[:3, 3], sc["means3D"], pose[:3, :3].t())
    im, _, _, depth, _, _ = raster(means3D=tp, means2D=torch.zeros_like(tp), opacities=sc["opacities"],
                                   colors_precomp=sc["colors_precomp"], scales=sc["scales"], rotations=sc["rotations"],
                                   semantics_precomp=sc["semantics_precomp"])
tr = GraphedTracker(settings)
args = (sc["means3D"], sc["colors_precomp"], sc["opacities"], sc["scales"], sc["rotations"], im, depth,
        torch.tensor([1.0, 0, 0, 0]), torch.zeros(3))
tr.track(*args, num_iters=2)
torch.cuda.synchronize()
print("MARK")
tr.track(*args, num_iters=int(sys.argv[2]) if len(sys.argv) > 2 else 2)
print("ok")

"""Debug aid: per-iteration losses of GraphedTracker vs the reference-style eager loop at a given config."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
import diff_gaussian_rasterization as ours
from hier_slam_b200.scene import CONFIGS, keyframe_poses, make_scene
from hier_slam_b200.tracking import GraphedTracker, _pose_matrix
key = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = CONFIGS[key]
scene = make_scene(cfg, 0, device="cuda")
settings = pt.make_settings(ours.GaussianRasterizationSettings, cfg, "cuda")
raster = ours.GaussianRasterizer_semantic(settings)
gt_pose = keyframe_poses(2, seed=2, max_angle_deg=1.0, max_trans=0.02).to("cuda")[1]
with torch.no_grad():
    tp = torch.addmm(gt_pose[:3, 3], scene["means3D"], gt_pose[:3, :3].t())
    gt_im, _, _, gt_depth, _, _ = raster(means3D=tp, means2D=torch.zeros_like(tp), opacities=scene["opacities"],
                                         colors_precomp=scene["colors_precomp"], scales=scene["scales"],
                                         rotations=scene["rotations"], semantics_precomp=scene["semantics_precomp"])
cam_rot = torch.tensor([1.0, 0, 0, 0], device="cuda").requires_grad_(True)
cam_tran = torch.zeros(3, device="cuda").requires_grad_(True)
opt = torch.optim.Adam([{"params": [cam_rot], "lr": 0.0004}, {"params": [cam_tran], "lr": 0.002}])
pts = scene["means3D"]; ones = torch.ones(pts.shape[0], 1, device="cuda")
hist = []
for it in range(iters):
    rel = _pose_matrix(cam_rot, cam_tran)
    tpp = (rel @ torch.cat((pts, ones), 1).T).T[:, :3]
    im, _, _, depth, _, sil = raster(means3D=tpp, means2D=torch.zeros_like(pts), opacities=scene["opacities"],
                                     colors_precomp=scene["colors_precomp"], scales=scene["scales"],
                                     rotations=scene["rotations"], semantics_precomp=scene["semantics_precomp"])
    mask = (gt_depth > 0) & ~torch.isnan(depth) & (sil > 0.99)
    loss = torch.abs(gt_depth - depth)[mask].sum() + 0.5 * torch.abs(gt_im - im)[mask.expand(3, -1, -1)].sum()
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
    hist.append((float(loss), cam_rot.detach().cpu().clone(), cam_tran.detach().cpu().clone(), int(mask.sum())))
tr = GraphedTracker(settings)
args = (scene["means3D"], scene["colors_precomp"], scene["opacities"], scene["scales"], scene["rotations"], gt_im, gt_depth,
        torch.tensor([1.0, 0, 0, 0]), torch.zeros(3))
for n in (1, 2, 3, 5, 10, 20, iters):
    o = tr.track(*args, num_iters=n)
    e = hist[n - 1]
    print(f"iters {n:3d}: eager last loss {e[0]:.4f} mask {e[3]}  tracker last loss {o['last_loss']:.4f}  |drot| {float((o['last_rot']-e[1]).abs().max()):.2e} "
          f"|dtran| {float((o['last_tran']-e[2]).abs().max()):.2e}  retries {o['retries']} best {o['loss']:.4f} eager best {min(h[0] for h in hist[:n]):.4f}")

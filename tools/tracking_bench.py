"""BASELINE.json config 3: Replica-shape tracking loop -- 40 pose-only gradient iterations per frame through the
PUBLIC rasterizer API, the way scripts/hierslam.py:1837-1852 drives it:

    cam pose (unnormalised quaternion + translation)  ->  rel_w2c  ->  transformed means (torch, autograd)
    ->  GaussianRasterizer_semantic  ->  L1 depth + L1 colour, summed over the silhouette mask
    ->  backward  ->  Adam step on the two camera tensors (every Gaussian tensor keeps requires_grad, lr 0).

usage: python tools/tracking_bench.py [--impl ours|ref-cuda|both] [--frames 3] [--iters 40] [--config c2]
Prints one JSON line per implementation with iterations/s and frames/s, and (both) the pose trajectories' distance.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt  # noqa: E402
from hier_slam_b200.scene import CONFIGS, keyframe_poses, make_scene  # noqa: E402


def quat_to_rot(q):
    """Rotation matrix of a unit quaternion (r, x, y, z) -- same convention as utils/slam_external.py build_rotation."""
    r, x, y, z = q[0], q[1], q[2], q[3]
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)]),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)]),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)])])


def run(mod, cfg, frames, iters, dev="cuda", fused=False, fused_loss=False):
    Settings, Raster = mod.GaussianRasterizationSettings, mod.GaussianRasterizer_semantic
    sc = make_scene(cfg, 0, device=dev)
    P = sc["means3D"].shape[0]
    # map parameters in the reference's parametrisation (all require grad, like the nn.Parameters of hierslam.py)
    params = dict(means3D=sc["means3D"], rgb_colors=sc["colors_precomp"], semantic=sc["semantics_precomp"],
                  unnorm_rotations=sc["rotations"], logit_opacities=torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)),
                  log_scales=torch.log(sc["scales"][:, :1]))
    params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    settings = pt.make_settings(Settings, cfg, dev)
    raster = Raster(raster_settings=settings)
    # ground truth of every frame: a render from a perturbed pose (the scene is static)
    gt_poses = keyframe_poses(frames, seed=2, max_angle_deg=1.0, max_trans=0.02).to(dev)
    ones = torch.ones(P, 1, device=dev)

    if fused:   # pose as an input of the rasterizer, pose gradient pre-reduced in the backward kernel
        from hier_slam_b200.tracking import PoseRasterizer_semantic
        pose_raster = PoseRasterizer_semantic(settings)

    def render(w2c):
        pts = params["means3D"].detach()
        if fused:
            return pose_raster(w2c, pts, torch.zeros_like(pts), torch.sigmoid(params["logit_opacities"]),
                               params["rgb_colors"], torch.exp(torch.tile(params["log_scales"], (1, 3))),
                               F.normalize(params["unnorm_rotations"]), params["semantic"])
        tp = (w2c @ torch.cat((pts, ones), 1).T).T[:, :3]
        return raster(means3D=tp, means2D=torch.zeros_like(pts, requires_grad=True) + 0,
                      opacities=torch.sigmoid(params["logit_opacities"]), colors_precomp=params["rgb_colors"],
                      scales=torch.exp(torch.tile(params["log_scales"], (1, 3))),
                      rotations=F.normalize(params["unnorm_rotations"]), semantics_precomp=params["semantic"])
    gts = []
    with torch.no_grad():
        for f in range(frames):
            im, _, _, depth, _, sil = render(gt_poses[f])
            gts.append((im.clone(), depth.clone(), (sil > 0.99)))
    cam_rot = torch.tensor([1.0, 0, 0, 0], device=dev).requires_grad_(True)
    cam_tran = torch.zeros(3, device=dev).requires_grad_(True)
    opt = torch.optim.Adam([{"params": [cam_rot], "lr": 0.0004}, {"params": [cam_tran], "lr": 0.002}])
    traj = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(frames):
        gt_im, gt_depth, mask = gts[f]
        for it in range(iters):
            rel = torch.eye(4, device=dev)
            rel[:3, :3] = quat_to_rot(F.normalize(cam_rot, dim=0))
            rel[:3, 3] = cam_tran
            im, radius, sem, depth, median, sil = render(rel)
            m = mask & (gt_depth > 0)
            if fused_loss:   # same value and gradient, one kernel per image instead of nonzero + index_put_
                from hier_slam_b200.losses import masked_l1_sum
                loss = masked_l1_sum(depth, gt_depth, m) + 0.5 * masked_l1_sum(im, gt_im, m)
            else:
                loss = torch.abs(gt_depth - depth)[m].sum() + 0.5 * torch.abs(gt_im - im)[m.expand(3, -1, -1)].sum()
            opt.zero_grad(set_to_none=True)
            for v in params.values():
                v.grad = None
            loss.backward()
            opt.step()
        traj.append(torch.cat([cam_rot.detach(), cam_tran.detach()]).cpu())
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    return dict(iters_per_s=frames * iters / (ms * 1e-3), frames_per_s=frames / (ms * 1e-3), ms_per_iter=ms / (frames * iters),
                wall_ms_per_iter=wall * 1e3 / (frames * iters), final_loss=float(loss)), torch.stack(traj)


def run_graphed(mod, cfg, frames, iters, dev="cuda"):
    """the same loop through hier_slam_b200.tracking.GraphedTracker: one CUDA-graph launch per iteration, the mask is the
    reference's per-iteration one ((gt_depth > 0) & (rendered silhouette > 0.99)), one host sync per frame"""
    from hier_slam_b200.tracking import GraphedTracker
    sc = make_scene(cfg, 0, device=dev)
    settings = pt.make_settings(mod.GaussianRasterizationSettings, cfg, dev)
    raster = mod.GaussianRasterizer_semantic(raster_settings=settings)
    gt_poses = keyframe_poses(frames, seed=2, max_angle_deg=1.0, max_trans=0.02).to(dev)
    gts = []
    with torch.no_grad():
        for f in range(frames):
            tp = torch.addmm(gt_poses[f][:3, 3], sc["means3D"], gt_poses[f][:3, :3].t())
            im, _, _, depth, _, _ = raster(means3D=tp, means2D=torch.zeros_like(tp), opacities=sc["opacities"],
                                           colors_precomp=sc["colors_precomp"], scales=sc["scales"],
                                           rotations=sc["rotations"], semantics_precomp=sc["semantics_precomp"])
            gts.append((im.clone(), depth.clone()))
    tracker = GraphedTracker(settings)
    rot, tran = torch.tensor([1.0, 0, 0, 0]), torch.zeros(3)
    args = (sc["means3D"], sc["colors_precomp"], sc["opacities"], sc["scales"], sc["rotations"])
    tracker.track(*args, gts[0][0], gts[0][1], rot, tran, num_iters=3)       # capture + warm-up
    traj, retries = [], 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(frames):
        out = tracker.track(*args, gts[f][0], gts[f][1], rot, tran, num_iters=iters)
        rot, tran = out["last_rot"], out["last_tran"]          # the bench's other arms continue from the last pose too
        retries += out["retries"]
        traj.append(torch.cat([rot, tran]))
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    return dict(iters_per_s=frames * iters / (ms * 1e-3), frames_per_s=frames / (ms * 1e-3), ms_per_iter=ms / (frames * iters),
                wall_ms_per_iter=wall * 1e3 / (frames * iters), final_loss=out["last_loss"], retries=retries,
                graph_captures=tracker.captures), torch.stack(traj)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="both")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--config", default="c2")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    res = {}
    if a.impl in ("ours", "both"):
        import diff_gaussian_rasterization as ours
        run(ours, cfg, 1, 5)                       # warm-up
        res["ours"] = run(ours, cfg, a.frames, a.iters)
        print(json.dumps({"impl": "ours", "config": cfg.name, "workload": "c3 tracking", **res["ours"][0]}))
        run(ours, cfg, 1, 5, fused=True)
        res["fused"] = run(ours, cfg, a.frames, a.iters, fused=True)
        print(json.dumps({"impl": "ours-fused-pose", "config": cfg.name, "workload": "c3 tracking", **res["fused"][0],
                          "pose_trajectory_max_abs_diff_vs_unfused": float((res["fused"][1] - res["ours"][1]).abs().max())}))
        run(ours, cfg, 1, 5, fused=True, fused_loss=True)
        res["fused2"] = run(ours, cfg, a.frames, a.iters, fused=True, fused_loss=True)
        print(json.dumps({"impl": "ours-fused-pose+masked-l1", "config": cfg.name, "workload": "c3 tracking", **res["fused2"][0],
                          "pose_trajectory_max_abs_diff_vs_unfused": float((res["fused2"][1] - res["ours"][1]).abs().max())}))
        res["graphed"] = run_graphed(ours, cfg, a.frames, a.iters)
        print(json.dumps({"impl": "ours-graphed-tracker", "config": cfg.name, "workload": "c3 tracking", **res["graphed"][0],
                          "pose_trajectory_max_abs_diff_vs_unfused": float((res["graphed"][1] - res["ours"][1]).abs().max())}))
    if a.impl in ("ref-cuda", "both"):
        from oracle import ref_loader
        ref = ref_loader.load_reference(cfg.num_semantic)
        if ref is None:
            print(json.dumps({"impl": "ref-cuda", "unavailable": "oracle/_ref not built on this box"}))
        else:
            run(ref, cfg, 1, 5)
            res["ref"] = run(ref, cfg, a.frames, a.iters)
            print(json.dumps({"impl": "ref-cuda", "config": cfg.name, "workload": "c3 tracking", **res["ref"][0]}))
    if "ours" in res and "ref" in res:
        d = (res["ours"][1] - res["ref"][1]).abs().max()
        print(json.dumps({"pose_trajectory_max_abs_diff": float(d),
                          "speedup": res["ours"][0]["iters_per_s"] / res["ref"][0]["iters_per_s"]}))


if __name__ == "__main__":
    main()

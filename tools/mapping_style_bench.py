"""'Mapping-style' variant of the c2 step: forward + the losses Hier-SLAM's mapping uses + backward.

    depth : mean |gt - depth| over the valid-depth mask                       (scripts/hierslam.py:780-786)
    colour: mean |gt - im| (the 0.8 L1 term; the 0.2 SSIM term is a torch conv stack and is left out on every arm)
    sem   : sum over the tree levels of CrossEntropyLoss on the level's channel slice   (:955-1000), S = 26 = [4,5,5,6,6]

Arms: reference CUDA rasterizer + the reference's torch loss code; this rasterizer + the same torch loss code; this
rasterizer + hier_slam_b200.losses (masked_l1_sum, hierarchical_cross_entropy).  CUDA events, one JSON line per arm."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
import diff_gaussian_rasterization as ours
from hier_slam_b200.losses import hierarchical_cross_entropy, masked_l1_sum
from hier_slam_b200.scene import CONFIGS, make_scene
from oracle import ref_loader

cfg = CONFIGS["c2"]; dev = "cuda"; sizes = [4, 5, 5, 6, 6]
sc = make_scene(cfg, 0, device=dev)
g = torch.Generator().manual_seed(3)
H, W = cfg.height, cfg.width
gt_im = torch.rand(3, H, W, generator=g).cuda(); gt_depth = (0.5 + 5 * torch.rand(1, H, W, generator=g)).cuda()
labels = torch.stack([torch.randint(0, n, (H, W), generator=g) for n in sizes]).cuda()
mask = gt_depth > 0.6
ce = torch.nn.CrossEntropyLoss()


def torch_losses(im, depth, sem):
    l = torch.abs(gt_depth - depth)[mask].mean() + 0.5 * torch.abs(gt_im - im).mean()
    beg = 0
    for i, n in enumerate(sizes):
        l = l + 0.01 * ce(sem[beg:beg + n].permute(1, 2, 0).reshape(-1, n), labels[i].view(-1).long()); beg += n
    return l


n_mask = float(mask.sum())
def fused_losses(im, depth, sem):
    return (masked_l1_sum(depth, gt_depth, mask) / n_mask + 0.5 * masked_l1_sum(im, gt_im, None) / im.numel()
            + hierarchical_cross_entropy(sem, labels, sizes, weights=[0.01] * len(sizes)))


def run(mod, loss_fn, iters=20):
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    raster = mod.GaussianRasterizer_semantic(pt.make_settings(mod.GaussianRasterizationSettings, cfg, dev))
    m2d = torch.zeros_like(leaves["means3D"])
    def step():
        for v in leaves.values(): v.grad = None
        im, radii, sem, depth, median, sil = raster(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                                    colors_precomp=leaves["colors_precomp"], scales=leaves["scales"],
                                                    rotations=leaves["rotations"], semantics_precomp=leaves["semantics_precomp"])
        loss = loss_fn(im, depth, sem)
        loss.backward()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): loss = step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, float(loss), {k: v.grad.clone() for k, v in leaves.items()}


res = {}
ref = ref_loader.load_reference(26)
if ref is not None:
    res["reference rasterizer + torch losses"] = run(ref, torch_losses, 10)
res["this rasterizer + torch losses"] = run(ours, torch_losses)
res["this rasterizer + fused losses"] = run(ours, fused_losses)
base = res["this rasterizer + torch losses"]
for k, (ms, loss, grads) in res.items():
    err = max(pt.grad_err(grads[n], base[2][n])[0] for n in grads)
    print(json.dumps({"arm": k, "ms_per_iteration": round(ms, 3), "iterations_per_s": round(1e3 / ms, 1), "loss": loss,
                      "max_normwise_grad_diff_vs_torch_loss_arm": err}))

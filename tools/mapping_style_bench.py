"""'Mapping-style' variant of the c2 step: forward + the losses Hier-SLAM's mapping uses + backward.

    depth : mean |gt - depth| over the valid-depth mask                       (scripts/hierslam.py:780-786)
    colour: mean |gt - im| (the 0.8 L1 term; the 0.2 SSIM term is a torch conv stack and is left out on every arm)
    sem   : sum over the tree levels of CrossEntropyLoss on the level's channel slice   (:955-1000), S = 26 = [4,5,5,6,6]

Arms: reference CUDA rasterizer + the reference's torch loss code; this rasterizer + the same torch loss code; this
rasterizer + hier_slam_b200.losses (masked_l1_sum, hierarchical_cross_entropy).  CUDA events, one JSON line per arm.

Second block ("full"): the complete mapping loss of get_loss_semantic_mlp after iteration 14 with the weights of
configs/replica/hierslam_semantic_run.py:103-107 -- 1.0 depth + 0.5 (0.8 L1 + 0.2 (1 - SSIM)) + 0.2 (1.0 level CE +
5.0 leaf CE behind the 1x1 convolution to 102 classes) -- torch loss code vs l1_ssim_loss + tree_semantic_loss."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
import diff_gaussian_rasterization as ours
from hier_slam_b200.losses import hierarchical_cross_entropy, l1_ssim_loss, masked_l1_sum, tree_semantic_loss
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


from math import exp
_g1 = torch.tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
_g1 = (_g1 / _g1.sum()).unsqueeze(1)
_win = _g1.mm(_g1.t()).float()[None, None].expand(3, 1, 11, 11).contiguous().cuda()
def torch_ssim(a, b):     # utils/slam_external.py:77-97
    f = lambda x: torch.nn.functional.conv2d(x, _win, padding=5, groups=3)
    mu1, mu2 = f(a), f(b)
    s1, s2, s12 = f(a * a) - mu1 * mu1, f(b * b) - mu2 * mu2, f(a * b) - mu1 * mu2
    return (((2 * mu1 * mu2 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1 * mu1 + mu2 * mu2 + 0.01 ** 2) * (s1 + s2 + 0.03 ** 2))).mean()

LEAVES = 102
conv = torch.nn.Conv2d(sum(sizes), LEAVES, kernel_size=1).cuda()
leaf_labels = torch.randint(0, LEAVES, (H, W), generator=g).cuda()
all_labels = torch.cat((labels, leaf_labels[None])).int()     # converted once per keyframe, as a caller would


def torch_losses_full(im, depth, sem):
    l = torch.abs(gt_depth - depth)[mask].mean() + 0.5 * (0.8 * torch.abs(im - gt_im).mean() + 0.2 * (1.0 - torch_ssim(im, gt_im)))
    lv, beg = 0.0, 0
    for i, n in enumerate(sizes):
        lv = lv + ce(sem[beg:beg + n].permute(1, 2, 0).reshape(-1, n), labels[i].view(-1).long()); beg += n
    logits = conv(sem.unsqueeze(0))
    logits = logits.squeeze(0).view(logits.shape[1], -1).permute(1, 0)
    return l + 0.2 * (1.0 * lv + 5.0 * ce(logits, leaf_labels.view(-1).long()))


def fused_losses_full(im, depth, sem):
    return (masked_l1_sum(depth, gt_depth, mask) / n_mask + 0.5 * l1_ssim_loss(im, gt_im)
            + 0.2 * tree_semantic_loss(sem, all_labels, sizes, conv.weight, conv.bias, 1.0, 5.0, num_valid=H * W,
                                       level_valid=H * W))


def run(mod, loss_fn, iters=20):
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    raster = mod.GaussianRasterizer_semantic(pt.make_settings(mod.GaussianRasterizationSettings, cfg, dev))
    m2d = torch.zeros_like(leaves["means3D"])
    def step():
        for v in leaves.values(): v.grad = None
        conv.zero_grad()
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

torch.backends.cudnn.allow_tf32 = False       # float32 torch arm, so that the gradient comparison means something
full = {}
if ref is not None:
    full["reference rasterizer + torch losses (full mapping loss)"] = run(ref, torch_losses_full, 10)
full["this rasterizer + torch losses (full mapping loss)"] = run(ours, torch_losses_full)
full["this rasterizer + fused losses (full mapping loss)"] = run(ours, fused_losses_full)
base = full["this rasterizer + torch losses (full mapping loss)"]
for k, (ms, loss, grads) in full.items():
    err = max(pt.grad_err(grads[n], base[2][n])[0] for n in grads)
    print(json.dumps({"arm": k, "ms_per_iteration": round(ms, 3), "iterations_per_s": round(1e3 / ms, 1), "loss": loss,
                      "max_normwise_grad_diff_vs_torch_loss_arm": err}))

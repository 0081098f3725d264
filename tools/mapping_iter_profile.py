"""A few full mapping iterations at the c2 shape (render + complete loss + backward + FlatAdam) for an ncu launch list."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
import diff_gaussian_rasterization as ours
from hier_slam_b200.losses import l1_ssim_loss, masked_l1_sum, tree_semantic_loss
from hier_slam_b200.mapping import FlatParams
from hier_slam_b200.optim import FlatAdam
from hier_slam_b200.scene import CONFIGS, make_scene
cfg = CONFIGS["c2"]; dev = "cuda"; sizes = [4, 5, 5, 6, 6]; H, W = cfg.height, cfg.width
params = FlatParams(make_scene(cfg, 0, device=dev)); lv = params.leaves
raster = ours.GaussianRasterizer_semantic(pt.make_settings(ours.GaussianRasterizationSettings, cfg, dev))
g = torch.Generator().manual_seed(3)
gt_im = torch.rand(3, H, W, generator=g).cuda(); gt_depth = (0.5 + 5 * torch.rand(1, H, W, generator=g)).cuda()
labels = torch.stack([torch.randint(0, n, (H, W), generator=g) for n in sizes + [102]]).cuda()
if len(sys.argv) > 1 and sys.argv[1] == "int32":
    labels = labels.int()
mask = gt_depth > 0.6; n_mask = float(mask.sum())
conv = torch.nn.Conv2d(26, 102, kernel_size=1).cuda()
opt = FlatAdam(params, {k: 1e-3 for k in params.names}, eps=1e-15); conv_opt = torch.optim.Adam(conv.parameters(), lr=5e-4)
m2d = torch.zeros_like(lv["means3D"])
def step():
    opt.zero_grad(); conv_opt.zero_grad(set_to_none=True)
    im, radii, sem, depth, median, sil = raster(means3D=lv["means3D"], means2D=m2d, opacities=lv["opacities"],
                                                colors_precomp=lv["colors_precomp"], scales=lv["scales"],
                                                rotations=lv["rotations"], semantics_precomp=lv["semantics_precomp"])
    loss = (masked_l1_sum(depth, gt_depth, mask) / n_mask + 0.5 * l1_ssim_loss(im, gt_im)
            + 0.2 * tree_semantic_loss(sem, labels, sizes, conv.weight, conv.bias, 1.0, 5.0, num_valid=H * W))
    loss.backward(); opt.step(); conv_opt.step()
for _ in range(4): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
print("ms/iteration", e0.elapsed_time(e1) / 10)

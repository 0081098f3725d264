"""One leaf-loss evaluation at a named shape (for ncu): python tools/leaf_profile.py [S L H W]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hier_slam_b200.losses import leaf_cross_entropy
S, L, H, W = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (26, 102, 680, 1200)
g = torch.Generator().manual_seed(0)
sem = torch.randn(S, H, W, generator=g).cuda().requires_grad_(True)
conv = torch.nn.Conv2d(S, L, kernel_size=1).cuda()
leaf = torch.randint(0, L, (H, W), generator=g).cuda()
for _ in range(3):
    sem.grad = None; conv.zero_grad()
    leaf_cross_entropy(sem, leaf, conv.weight, conv.bias, num_valid=H * W).backward()
torch.cuda.synchronize()
print("ok")

"""Hier-SLAM's inter-level semantic loss at the c2 shape (1200x680, S = 26 = [4,5,5,6,6]): the reference's torch code
(scripts/hierslam.py:955-1000) vs hier_slam_b200.losses.hierarchical_cross_entropy; forward + backward, CUDA events."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hier_slam_b200.losses import hierarchical_cross_entropy
H, W, sizes = 680, 1200, [4, 5, 5, 6, 6]
g = torch.Generator().manual_seed(0)
sem = torch.randn(sum(sizes), H, W, generator=g).cuda().requires_grad_(True)
labels = torch.stack([torch.randint(0, n, (H, W), generator=g) for n in sizes]).cuda()
ce = torch.nn.CrossEntropyLoss()
def ref():
    sem.grad = None
    b, beg = 0.0, 0
    for l, n in enumerate(sizes):
        b = b + ce(sem[beg:beg + n].permute(1, 2, 0).reshape(-1, n), labels[l].view(-1).long()); beg += n
    b.backward()
def ours():
    sem.grad = None
    hierarchical_cross_entropy(sem, labels, sizes).backward()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(json.dumps(dict(loss="inter-level CE", shape=[sum(sizes), H, W], torch_ms=round(t(ref), 3), fused_ms=round(t(ours), 3))))

# leaf loss: 1x1 conv to the leaf classes + CE (scripts/hierslam.py:975-984) vs hier_slam_b200.losses.leaf_cross_entropy
from hier_slam_b200.losses import leaf_cross_entropy
for (S, L, hh, ww) in ((26, 102, 680, 1200), (16, 41, 480, 640), (74, 550, 480, 640)):
    sem = torch.randn(S, hh, ww, generator=g).cuda().requires_grad_(True)
    conv = torch.nn.Conv2d(S, L, kernel_size=1).cuda()
    leaf = torch.randint(0, L, (hh, ww), generator=g).cuda()
    def ref():
        sem.grad = None; conv.zero_grad()
        logits = conv(sem.unsqueeze(0))
        logits = logits.squeeze(0).view(logits.shape[1], -1).permute(1, 0)
        ce(logits, leaf.view(-1).long()).backward()
    def ours():
        sem.grad = None; conv.zero_grad()
        leaf_cross_entropy(sem, leaf, conv.weight, conv.bias, num_valid=hh * ww).backward()
    row = dict(loss="leaf CE (1x1 conv)", shape=[S, hh, ww], classes=L, torch_tf32_ms=round(t(ref), 3))
    torch.backends.cudnn.allow_tf32 = False
    row["torch_fp32_ms"] = round(t(ref), 3)
    torch.backends.cudnn.allow_tf32 = True
    row["fused_ms"] = round(t(ours), 3)
    print(json.dumps(row))

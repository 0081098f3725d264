"""Adam step and pruning on Hier-SLAM's parameter set at the c2 size (300K Gaussians, S = 26): torch.optim.Adam (default
multi-tensor path and fused=True) vs hier_slam_b200.optim.FlatAdam; remove_points-style boolean indexing vs FlatAdam.prune.
CUDA events; one JSON line."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hier_slam_b200.mapping import FlatParams
from hier_slam_b200.optim import FlatAdam
P, S = int(sys.argv[1]) if len(sys.argv) > 1 else 300000, 26
shapes = {"means3D": (P, 3), "rgb_colors": (P, 3), "unnorm_rotations": (P, 4), "logit_opacities": (P, 1),
          "log_scales": (P, 1), "semantic": (P, S)}
lrs = {"means3D": 1e-4, "rgb_colors": 2.5e-3, "unnorm_rotations": 1e-3, "logit_opacities": 5e-2, "log_scales": 1e-3,
       "semantic": 2.5e-3}
g = torch.Generator().manual_seed(0)
init = {k: torch.randn(*v, generator=g).cuda() for k, v in shapes.items()}
grads = {k: torch.randn(*v, generator=g).cuda() for k, v in shapes.items()}


def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def torch_adam(**kw):
    ps = {k: torch.nn.Parameter(v.clone()) for k, v in init.items()}
    for k in ps: ps[k].grad = grads[k].clone()
    opt = torch.optim.Adam([{"params": [v], "name": k, "lr": lrs[k]} for k, v in ps.items()], lr=0.0, eps=1e-15, **kw)
    return timed(opt.step), ps, opt


fp = FlatParams({k: v.clone() for k, v in init.items()})
for k in shapes: fp.leaves[k].grad.copy_(grads[k])
ours = FlatAdam(fp, lrs, eps=1e-15)
row = {"gaussians": P, "floats": fp.flat.numel(), "torch_adam_foreach_ms": round(torch_adam()[0], 4),
       "torch_adam_fused_ms": round(torch_adam(fused=True)[0], 4), "flat_adam_ms": round(timed(ours.step), 4)}
row["flat_adam_GBps"] = round(28 * fp.flat.numel() / row["flat_adam_ms"] / 1e6, 1)

keep = (torch.rand(P, generator=g) < 0.9).cuda()
_, ps, opt = torch_adam()
def torch_prune():          # utils/slam_external.py:142-164 on copies (parameters and both moments of every tensor)
    out = []
    for k, v in ps.items():
        st = opt.state[v]
        out.append((v.detach()[keep], st["exp_avg"][keep], st["exp_avg_sq"][keep]))
    return out
def ours_prune():
    o2 = FlatAdam.__new__(FlatAdam)
    o2.params, o2.exp_avg, o2.exp_avg_sq = fp, ours.exp_avg, ours.exp_avg_sq
    fp.release = lambda: None          # the source set is pruned again in the next timing iteration
    o2.prune(keep).release()
row["torch_prune_ms"] = round(timed(torch_prune, 20), 4)
row["flat_prune_ms"] = round(timed(ours_prune, 20), 4)
print(json.dumps(row))

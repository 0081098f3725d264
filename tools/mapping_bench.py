"""BASELINE.json configs 4 / 5: multi-keyframe mapping, K = 8 keyframes per iteration partitioned across the GPUs of
one box (strong scaling: K fixed), gradients combined with ONE NCCL all-reduce of the flat buffer per iteration.

    python tools/mapping_bench.py --config c4                                 # 1 GPU
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/mapping_bench.py --config c5

Every keyframe is a different pose of the replicated Gaussian set, rendered through the public API
(hier_slam_b200.mapping.mapping_iteration + GaussianRasterizer_semantic) with a raster-only loss
(<output, fixed N(0,1)/N image> on colour, semantics and depth).  Timed with CUDA events, max over ranks; rank 0
prints one JSON line (keyframes/s over all ranks)."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
import diff_gaussian_rasterization as dgr
from hier_slam_b200.mapping import FlatParams, capacity_for, keyframes_of_rank, mapping_iteration
from hier_slam_b200.scene import CONFIGS, keyframe_poses, make_scene, upstream_grads

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c4")
ap.add_argument("--keyframes", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--capacity", action="store_true", help="sync-free forwards (capacity-mode binning)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", device_id=dev)
cfg = CONFIGS[a.config]
sc = make_scene(cfg, 0, device=dev)
ug = upstream_grads(cfg, 1, device=dev)
poses = keyframe_poses(a.keyframes, seed=2)
P = sc["means3D"].shape[0]
ones = torch.ones(P, 1, device=dev)


def make_loss(k):
    settings = pt.make_settings(dgr.GaussianRasterizationSettings, cfg, dev, w2c=poses[k])
    r = dgr.GaussianRasterizer_semantic(settings)
    m2d = torch.zeros(P, 3, device=dev)

    def f(lv):
        color, radii, sem, depth, median, opac = r(means3D=lv["means3D"], means2D=m2d, opacities=lv["opacities"],
                                                   colors_precomp=lv["colors_precomp"], scales=lv["scales"],
                                                   rotations=lv["rotations"], semantics_precomp=lv["semantics_precomp"])
        return (color * ug["color"]).sum() + (sem * ug["semantic"]).sum() + (depth * ug["depth"]).sum()
    return f


losses = [make_loss(k) for k in range(a.keyframes)]
params = FlatParams(sc)
cap = capacity_for([losses[k] for k in keyframes_of_rank(a.keyframes, rank, world)], params) if a.capacity else None
for _ in range(a.warmup):
    mapping_iteration(params, losses, rank, world, capacity=cap)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    mapping_iteration(params, losses, rank, world, capacity=cap)
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = float(ms) / a.iters
    print(json.dumps(dict(workload=cfg.name, keyframes_per_iteration=a.keyframes, n_gpus=world, scaling="strong",
                          ms_per_iteration=round(t, 3), keyframes_per_s=round(a.keyframes / (t * 1e-3), 1),
                          allreduce_bytes=params.grad_bytes(), gaussians=P, semantic_channels=cfg.num_semantic,
                          capacity_mode=bool(a.capacity))))
if world > 1:
    dist.destroy_process_group()

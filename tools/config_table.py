"""fwd+bwd time of every BASELINE.json shape through the `_C`-level entry points: new path vs the unmodified reference
CUDA build (oracle/_ref, when present for that S).  One JSON line per config.
usage: python tools/config_table.py [c1 c2 c4 c5]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
from hier_slam_b200 import _C, _lib
from hier_slam_b200.rasterizer import GaussianRasterizationSettings
from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads
from oracle import ref_loader


def time_impl(C, settings, scene, grads, iters, warm=3):
    for _ in range(warm):
        f = pt.run_forward(C, settings, scene); pt.run_backward(C, settings, scene, f, grads)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record(); f = pt.run_forward(C, settings, scene); e[1].record()
        pt.run_backward(C, settings, scene, f, grads); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    return tf / iters, tb / iters, f


for key in (sys.argv[1:] or ["c1", "c2", "c4", "c5"]):
    cfg = CONFIGS[key]
    scene = make_scene(cfg, 0, device="cuda"); grads = upstream_grads(cfg, 1, device="cuda")
    settings = pt.make_settings(GaussianRasterizationSettings, cfg)
    iters = 20
    f_ms, b_ms, f = time_impl(_C, settings, scene, grads, iters)
    N, P, S = cfg.width * cfg.height, cfg.num_gaussians, cfg.num_semantic
    V = int((f["radii"] > 0).sum())
    balg = 8 * (S + 6) * N + 4 * (46 + S) * P + 8 * (3 + S) * V
    row = dict(config=cfg.name, R=int(f["R"]), visible=V, ours_fwd_ms=round(f_ms, 3), ours_bwd_ms=round(b_ms, 3),
               ours_it_per_s=round(1e3 / (f_ms + b_ms), 1), B_alg_MB=round(balg / 1e6, 1),
               hbm_frac=round(balg / ((f_ms + b_ms) * 1e-3) / 6537e9, 4))
    ref = ref_loader.load_reference(S)
    if ref is not None:
        rf, rb, _ = time_impl(ref._C, settings, scene, grads, 10)
        row.update(ref_fwd_ms=round(rf, 3), ref_bwd_ms=round(rb, 3), ref_it_per_s=round(1e3 / (rf + rb), 1),
                   speedup=round((rf + rb) / (f_ms + b_ms), 2))
    print(json.dumps(row), flush=True)

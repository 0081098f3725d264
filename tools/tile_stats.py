"""Tile-list statistics of a config (tools/tile_stats.py [config])."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_tools as pt
from hier_slam_b200 import _C
from hier_slam_b200.rasterizer import GaussianRasterizationSettings
from hier_slam_b200.scene import CONFIGS, make_scene
key = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = CONFIGS[key]
scene = make_scene(cfg, 0, device="cuda")
settings = pt.make_settings(GaussianRasterizationSettings, cfg)
f = pt.run_forward(_C, settings, scene)
P, H, W = scene["means3D"].shape[0], cfg.height, cfg.width
sv = _C.state_views(P, H, W, f["R"], f["geomBuffer"], f["binningBuffer"], f["imgBuffer"])
rg = sv["ranges"].long()
ln = (rg[:, 1] - rg[:, 0]).float()
nc = sv["n_contrib"].float()
print(json.dumps(dict(config=key, P=P, R=int(f["R"]), tiles=int(ln.numel()), tile_len_mean=float(ln.mean()),
                      tile_len_max=int(ln.max()), tile_len_p50=float(ln.median()), n_contrib_mean=float(nc.mean()),
                      n_contrib_max=int(nc.max()))))

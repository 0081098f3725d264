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
# strip-hit density among the entries the forward visited (index < max n_contrib of the tile)
ncm = sv["n_contrib"].long().reshape(H, W)
gx, gy = (W + 15) // 16, (H + 15) // 16
pad = torch.zeros(gy * 16, gx * 16, dtype=torch.long, device=ncm.device); pad[:H, :W] = ncm
tmax = pad.reshape(gy, 16, gx, 16).permute(0, 2, 1, 3).reshape(gy * gx, 256).max(1)[0]
hits = sv["strip_hits"]
idx = torch.arange(f["R"], device=hits.device)
tile_of = (sv["keys"] >> 32).long()
pos = idx - rg[tile_of, 0]
visited = pos < tmax[tile_of]
bits = sum(((hits >> w) & 1).long() for w in range(8))
print(json.dumps(dict(visited_entries=int(visited.sum()), mean_strips_hit_per_visited_entry=float(bits[visited].float().mean()),
                      frac_pairs_hit=float(bits[visited].float().mean()) / 8, unvisited_bits=int(bits[~visited].sum()))))
print(json.dumps(dict(config=key, P=P, R=int(f["R"]), tiles=int(ln.numel()), tile_len_mean=float(ln.mean()),
                      tile_len_max=int(ln.max()), tile_len_p50=float(ln.median()), n_contrib_mean=float(nc.mean()),
                      n_contrib_max=int(nc.max()))))

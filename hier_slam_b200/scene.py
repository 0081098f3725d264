"""Seeded synthetic Gaussian scenes of the shapes named in BASELINE.json.

The generator mirrors how Hier-SLAM initialises its map (scripts/hierslam.py:144-194,361-409):
one isotropic Gaussian per back-projected pixel, log-scale = log(depth / focal), colours and
semantic embeddings uniform in [0,1].  Everything is generated on the CPU with a seeded
``torch.Generator`` and moved to the requested device afterwards, so the same seed gives the
same scene on the GPU box and in the CPU-only container.
"""
from __future__ import annotations

import math
from typing import Dict, NamedTuple, Optional

import torch


class SceneConfig(NamedTuple):
    name: str
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float
    num_gaussians: int
    num_semantic: int


# c1..c5 of SURVEY.md §8 (camera intrinsics: configs/data/replica.yaml:3-8, scannet_semantic.yaml:3-8)
CONFIGS: Dict[str, SceneConfig] = {
    "c1": SceneConfig("c1-320x240-10K-S26", 320, 240, 160.0, 160.0, 159.5, 119.5, 10_000, 26),
    "c2": SceneConfig("c2-replica-1200x680-300K-S26", 1200, 680, 600.0, 600.0, 599.5, 339.5, 300_000, 26),
    "c4": SceneConfig("c4-scannet-640x480-300K-S16", 640, 480, 1169.621094 * 640 / 1296, 1167.105103 * 480 / 968,
                      646.295044 * 640 / 1296, 489.927032 * 480 / 968, 300_000, 16),
    "c5": SceneConfig("c5-scannet-640x480-1M-S74", 640, 480, 1169.621094 * 640 / 1296, 1167.105103 * 480 / 968,
                      646.295044 * 640 / 1296, 489.927032 * 480 / 968, 1_000_000, 74),
    "tiny": SceneConfig("tiny-64x48-300-S26", 64, 48, 40.0, 40.0, 31.5, 23.5, 300, 26),
    "small": SceneConfig("small-160x112-3K-S26", 160, 112, 100.0, 100.0, 79.5, 55.5, 3_000, 26),
}


def camera_matrices(cfg: SceneConfig, w2c: Optional[torch.Tensor] = None, near: float = 0.01, far: float = 100.0):
    """Same construction as utils/recon_helpers.py:4-28 (setup_camera), device-agnostic.

    Returns (viewmatrix[1,4,4], projmatrix[1,4,4], campos[3], tanfovx, tanfovy); the two matrices are the
    transposed (column-major when flattened) forms the rasterizer indexes as m[4*c+r]."""
    w, h, fx, fy, cx, cy = cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy
    if w2c is None:
        w2c = torch.eye(4)
    w2c = w2c.float()
    cam_center = torch.inverse(w2c)[:3, 3]
    view = w2c.unsqueeze(0).transpose(1, 2)
    opengl_proj = torch.tensor([[2 * fx / w, 0.0, -(w - 2 * cx) / w, 0.0],
                                [0.0, 2 * fy / h, -(h - 2 * cy) / h, 0.0],
                                [0.0, 0.0, far / (far - near), -(far * near) / (far - near)],
                                [0.0, 0.0, 1.0, 0.0]]).float().unsqueeze(0).transpose(1, 2)
    full_proj = view.bmm(opengl_proj)
    return view, full_proj, cam_center, w / (2 * fx), h / (2 * fy)


def make_scene(cfg: SceneConfig, seed: int = 0, num_gaussians: Optional[int] = None,
               num_semantic: Optional[int] = None, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """SplaTAM-style synthetic scene: returns the *render variables* the
    reference's ``transformed_params2rendervar_semantic`` (utils/slam_helpers.py:195-219) would
    hand to the rasterizer: means3D (camera frame), colors_precomp, semantics_precomp, opacities,
    scales [P,3] (three equal columns), rotations [P,4] (normalised)."""
    P = cfg.num_gaussians if num_gaussians is None else num_gaussians
    S = cfg.num_semantic if num_semantic is None else num_semantic
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(P, generator=g) * cfg.width
    v = torch.rand(P, generator=g) * cfg.height
    z = 0.5 + 5.5 * torch.rand(P, generator=g)
    behind = torch.rand(P, generator=g) < 0.02           # 2 % behind / too close: exercises the cull path
    z = torch.where(behind, -1.0 + 1.2 * torch.rand(P, generator=g), z)
    means = torch.stack([(u - cfg.cx) * z / cfg.fx, (v - cfg.cy) * z / cfg.fy, z], -1)
    log_scale = torch.log(z.abs().clamp_min(1e-3) / ((cfg.fx + cfg.fy) / 2)) + 0.3 * torch.randn(P, generator=g)
    big = torch.rand(P, generator=g) < 0.10               # 10 % cover several tiles
    log_scale = log_scale + torch.where(big, math.log(8.0), 0.0)
    scales = torch.exp(log_scale)[:, None].repeat(1, 3)
    rot = torch.tensor([1.0, 0.0, 0.0, 0.0]) + 0.1 * torch.randn(P, 4, generator=g)
    rot = torch.nn.functional.normalize(rot)
    opac = torch.sigmoid(2.0 * torch.randn(P, 1, generator=g))
    rgb = torch.rand(P, 3, generator=g)
    sem = torch.rand(P, S, generator=g)
    out = dict(means3D=means, colors_precomp=rgb, semantics_precomp=sem, opacities=opac, scales=scales,
               rotations=rot)
    return {k: t.to(device).contiguous() for k, t in out.items()}


def upstream_grads(cfg: SceneConfig, seed: int = 1, num_semantic: Optional[int] = None, device: str = "cpu"):
    """'raster-only' upstream gradients: N(0,1)/N on every output."""
    S = cfg.num_semantic if num_semantic is None else num_semantic
    g = torch.Generator().manual_seed(seed)
    H, W = cfg.height, cfg.width
    n = float(H * W)
    mk = lambda c: (torch.randn(c, H, W, generator=g) / n).to(device)
    return dict(color=mk(3), semantic=mk(S), depth=mk(1), median_depth=mk(1), final_opacity=mk(1))


def keyframe_poses(num: int, seed: int = 2, max_angle_deg: float = 5.0, max_trans: float = 0.1) -> torch.Tensor:
    """Small SE(3) perturbations of identity (<= 5 deg / 0.1 m), one world-to-camera matrix per keyframe."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num):
        axis = torch.nn.functional.normalize(torch.randn(3, generator=g), dim=0)
        ang = math.radians(max_angle_deg) * float(torch.rand((), generator=g))
        K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = torch.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
        t = (torch.rand(3, generator=g) * 2 - 1) * max_trans
        m = torch.eye(4)
        m[:3, :3] = R
        m[:3, 3] = t
        out.append(m)
    return torch.stack(out)

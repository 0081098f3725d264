"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference CUDA extension
(oracle/_ref, built by oracle/build_ref.sh from /root/reference) on seeded synthetic scenes.

Must run on a CUDA box (the reference is CUDA-only):

    python tests/golden/make_golden.py --out gpurun_out/golden      # on the B200 box via gpurun
    cp gpurun_out/golden/*.npz tests/golden/                        # back in the dev container

Each fixture stores the scene key + seeds (inputs are re-generated from them by hier_slam_b200.scene), the
reference's outputs, its internal state (depths, means2D, conic_opacity, tiles_touched, keys, point_list,
ranges, n_contrib, final_T) and its gradients for seeded upstream gradients.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hier_slam_b200.scene import CONFIGS, camera_matrices, make_scene, upstream_grads  # noqa: E402
from oracle import ref_loader  # noqa: E402

CASES = [  # (fixture name, scene key, S, all five upstream grads?)
    ("tiny_S26", "tiny", 26, True),
    ("tiny_S26_colordepth", "tiny", 26, False),
    ("small_S26", "small", 26, True),
    ("tiny_S16", "tiny", 16, True),
]


def run_reference(ref, cfg, S, scene, grads, device="cuda"):
    view, proj, campos, tfx, tfy = camera_matrices(cfg)
    settings = ref.GaussianRasterizationSettings(
        image_height=cfg.height, image_width=cfg.width, tanfovx=tfx, tanfovy=tfy,
        bg=torch.zeros(3, device=device), scale_modifier=1.0, viewmatrix=view.to(device),
        projmatrix=proj.to(device), sh_degree=0, campos=campos.to(device), prefiltered=False, debug=False)
    inp = {k: v.to(device).clone().requires_grad_(True) for k, v in scene.items()}
    means2D = torch.zeros_like(inp["means3D"], requires_grad=True)
    args = (settings.bg, inp["means3D"], inp["colors_precomp"], inp["semantics_precomp"], inp["opacities"],
            inp["scales"], inp["rotations"], 1.0, torch.Tensor([]), settings.viewmatrix, settings.projmatrix, tfx,
            tfy, cfg.height, cfg.width, torch.Tensor([]), 0, settings.campos, False, False)
    (R, color, sem, depth, median, opacity, radii, geomB, binB, imgB) = ref._C.rasterize_gaussians_semantic(*args)
    z = lambda c: torch.zeros(c, cfg.height, cfg.width, device=device)
    g = {k: (v.to(device) if v is not None else None) for k, v in grads.items()}
    bargs = (settings.bg, inp["means3D"], radii, inp["colors_precomp"], inp["semantics_precomp"], inp["scales"],
             inp["rotations"], 1.0, torch.Tensor([]), settings.viewmatrix, settings.projmatrix, tfx, tfy,
             g["color"], g["semantic"] if g["semantic"] is not None else z(S),
             g["depth"], g["median_depth"] if g["median_depth"] is not None else z(1),
             g["final_opacity"] if g["final_opacity"] is not None else z(1),
             torch.Tensor([]), 0, settings.campos, geomB, R, binB, imgB, False)
    (d_means2D, d_colors, d_sem, d_opac, d_means3D, d_cov3D, d_sh, d_scales, d_rot) = \
        ref._C.rasterize_gaussians_backward_semantic(*bargs)
    torch.cuda.synchronize()
    st = ref_loader.parse_ref_state(inp["means3D"].shape[0], cfg.height, cfg.width, R, geomB, binB, imgB)
    out = dict(num_rendered=np.int64(R), color=color, semantic=sem, depth=depth, median_depth=median,
               final_opacity=opacity, radii=radii, d_means2D=d_means2D, d_colors=d_colors, d_semantics=d_sem,
               d_opacities=d_opac, d_means3D=d_means3D, d_cov3D=d_cov3D, d_scales=d_scales, d_rotations=d_rot)
    vis = radii > 0
    for k in ("depths", "means2D", "conic_opacity", "tiles_touched", "final_T", "n_contrib", "ranges", "point_list",
              "keys", "point_list_unsorted", "keys_unsorted"):
        if k in st:
            out["st_" + k] = st[k]
    # the reference leaves stale bytes for culled Gaussians in its geometry arrays: zero them for a stable fixture
    for k in ("st_depths", "st_means2D", "st_conic_opacity"):
        t = out[k].clone()
        t[~vis] = 0
        out[k] = t
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for name, key, S, all_grads in CASES:
        ref = ref_loader.load_reference(S)
        if ref is None:
            print(f"skip {name}: oracle/_ref/S{S} not built")
            continue
        cfg = CONFIGS[key]
        scene = make_scene(cfg, seed=0, num_semantic=S)
        grads = upstream_grads(cfg, seed=1, num_semantic=S)
        if not all_grads:
            grads["semantic"] = None
            grads["median_depth"] = None
            grads["final_opacity"] = None
        res = run_reference(ref, cfg, S, scene, grads)
        res.update(scene_key=key, S=np.int64(S), scene_seed=np.int64(0), grad_seed=np.int64(1),
                   all_grads=np.int64(all_grads))
        path = os.path.join(a.out, name + ".npz")
        np.savez_compressed(path, **res)
        print("wrote", path, os.path.getsize(path), "bytes; R =", int(res["num_rendered"]))


if __name__ == "__main__":
    main()

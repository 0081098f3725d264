"""algorithmic flops of the blend's contraction, F_alg = 6 * 256 * S * sum_t L_t (forward 2 * 256 * L_t * S,
backward twice that), L_t = the entries of tile t's sorted list that are traversed before the whole tile has terminated
(= the largest last-contributor index of its pixels).  sum_t L_t is counted on the CPU with the oracle (no GPU needed):

    python tools/tensor_fraction.py c5 [blend_bwd_ms] [blend_fwd_ms]

Prints sum_t L_t, R, F_alg and -- given measured kernel times -- the tensor fraction against the measured TF32 peak
(profiles/r2_tf32_peak.json).  Only the BACKWARD runs its contraction on the tensor cores (3xTF32 mma.sync); the forward
blend is SIMT by design, so the backward's 4 * 256 * S * sum L_t is the flop count that belongs
to a tensor roofline."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hier_slam_b200.scene import CONFIGS, camera_matrices, make_scene  # noqa: E402
from oracle import raster_oracle as O  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c5"
    bwd_ms = float(sys.argv[2]) if len(sys.argv) > 2 else None
    fwd_ms = float(sys.argv[3]) if len(sys.argv) > 3 else None
    cfg = CONFIGS[name]
    torch.set_num_threads(os.cpu_count() or 1)
    sc = make_scene(cfg, 0)
    view, proj, campos, tfx, tfy = camera_matrices(cfg)
    W, H, S = cfg.width, cfg.height, cfg.num_semantic
    t0 = time.time()
    geom = O.preprocess(sc["means3D"], sc["scales"], sc["rotations"], sc["opacities"], view, proj, W, H, tfx, tfy)
    keys, vals = O.duplicate_with_keys(geom["depths"], geom["means2D"], geom["radii"], W, H)
    _, plist, ranges = O.sort_and_ranges(keys, vals, W, H)
    # n_contrib does not depend on the feature channels: blend one semantic channel only
    fwd = O.blend_forward(geom, plist, ranges, sc["colors_precomp"], sc["semantics_precomp"][:, :1].contiguous(), W, H)
    gx, gy = O.tile_grid(W, H)
    nc = fwd["n_contrib"].reshape(H, W).long()
    pad = torch.zeros(gy * 16, gx * 16, dtype=torch.long)
    pad[:H, :W] = nc
    L = pad.reshape(gy, 16, gx, 16).permute(0, 2, 1, 3).reshape(gy * gx, 256).max(1).values
    R = int((ranges[:, 1] - ranges[:, 0]).sum())
    sumL = int(L.sum())
    F = 6 * 256 * S * sumL
    out = {"config": cfg.name, "S": S, "tiles": gx * gy, "num_rendered_R": R, "sum_Lt": sumL, "mean_Lt": sumL / (gx * gy),
           "max_Lt": int(L.max()), "F_alg_GFLOP": F / 1e9, "F_alg_backward_GFLOP": 4 * 256 * S * sumL / 1e9,
           "oracle_seconds": round(time.time() - t0, 1)}
    peak_file = os.path.join(ROOT, "profiles", "r2_tf32_peak.json")
    if os.path.exists(peak_file):
        peak = json.load(open(peak_file))["tf32_tflops"]
        out["tf32_peak_TFLOPs"] = peak
        if bwd_ms:
            a = 4 * 256 * S * sumL / (bwd_ms * 1e-3) / 1e12
            out["backward"] = {"ms": bwd_ms, "achieved_TFLOPs": a, "tensor_fraction": a / peak}
        if bwd_ms and fwd_ms:
            a = F / ((bwd_ms + fwd_ms) * 1e-3) / 1e12
            out["forward_plus_backward"] = {"ms": bwd_ms + fwd_ms, "achieved_TFLOPs": a, "tensor_fraction": a / peak,
                                            "note": "section 8(d)'s definition; the forward has no tensor-core work"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

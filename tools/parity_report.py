"""GPU diagnostic (never asserts): runs the new CUDA path and the reference CUDA build (oracle/_ref) on the same
seeded scene and prints one JSON object per scene with bit-exactness counts, image / gradient errors and
CUDA-event timings of both implementations.  Usage (on the GPU box):

    python tools/parity_report.py --scenes c1,c2 --iters 20 --out gpurun_out/parity_report.jsonl
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity_tools as pt  # noqa: E402
from hier_slam_b200 import _C as newC  # noqa: E402
from hier_slam_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402
from hier_slam_b200.scene import CONFIGS, make_scene, upstream_grads  # noqa: E402
from oracle import ref_loader  # noqa: E402


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return dict(median_ms=ts[len(ts) // 2], min_ms=ts[0], p90_ms=ts[int(0.9 * (len(ts) - 1))])


def report(key, iters, P=None, S=None, semantic=True):
    cfg = CONFIGS[key]
    S = cfg.num_semantic if S is None else S
    dev = "cuda"
    scene = make_scene(cfg, 0, num_gaussians=P, num_semantic=S, device=dev)
    grads = upstream_grads(cfg, 1, num_semantic=S, device=dev)
    settings = pt.make_settings(GaussianRasterizationSettings, cfg, dev)
    out = dict(scene=cfg.name, P=int(scene["means3D"].shape[0]), S=S, W=cfg.width, H=cfg.height, semantic=semantic)
    Pn, H, W = out["P"], cfg.height, cfg.width

    f_new = pt.run_forward(newC, settings, scene, semantic)
    torch.cuda.synchronize()
    out["R"] = int(f_new["R"])
    sv_new = newC.state_views(Pn, H, W, f_new["R"], f_new["geomBuffer"], f_new["binningBuffer"], f_new["imgBuffer"])
    out["visible"] = int((f_new["radii"] > 0).sum())
    out["n_contrib_mean"] = float(sv_new["n_contrib"].float().mean())
    out["n_contrib_max"] = int(sv_new["n_contrib"].max())
    rg = sv_new["ranges"]
    out["tile_len_mean"] = float((rg[:, 1] - rg[:, 0]).float().mean())
    out["tile_len_max"] = int((rg[:, 1] - rg[:, 0]).max())
    # culling on/off must not change anything
    newC.NO_CULL = True
    f_nc = pt.run_forward(newC, settings, scene, semantic)
    newC.NO_CULL = False
    out["cull_vs_nocull"] = {k: pt.bits_equal(f_new[k], f_nc[k]) for k in ("color", "depth", "median_depth",
                                                                            "final_opacity")}
    sv_nc = newC.state_views(Pn, H, W, f_nc["R"], f_nc["geomBuffer"], f_nc["binningBuffer"], f_nc["imgBuffer"])
    out["cull_vs_nocull"]["n_contrib"] = pt.bits_equal(sv_new["n_contrib"], sv_nc["n_contrib"])
    if semantic:
        out["cull_vs_nocull"]["semantic"] = pt.bits_equal(f_new["semantic"], f_nc["semantic"])

    g_new = pt.run_backward(newC, settings, scene, f_new, grads, semantic)
    torch.cuda.synchronize()

    ref = ref_loader.load_reference(S if semantic else 26)
    if ref is not None:
        f_ref = pt.run_forward(ref._C, settings, scene, semantic)
        g_ref = pt.run_backward(ref._C, settings, scene, f_ref, grads, semantic)
        torch.cuda.synchronize()
        sv_ref = ref_loader.parse_ref_state(Pn, H, W, f_ref["R"], f_ref["geomBuffer"], f_ref["binningBuffer"],
                                            f_ref["imgBuffer"])
        vis = f_ref["radii"] > 0
        ex = dict(R_equal=int(f_ref["R"]) == int(f_new["R"]), R_ref=int(f_ref["R"]))
        ex["radii"] = pt.bits_equal(f_new["radii"], f_ref["radii"])
        ex["tiles_touched"] = pt.bits_equal(sv_new["tiles_touched"], sv_ref["tiles_touched"])
        ex["depths"] = pt.bits_equal(sv_new["depths"][vis], sv_ref["depths"][vis])
        ex["means2D_x"] = pt.bits_equal(sv_new["means2D"][vis][:, 0], sv_ref["means2D"][vis][:, 0])
        ex["means2D_y"] = pt.bits_equal(sv_new["means2D"][vis][:, 1], sv_ref["means2D"][vis][:, 1])
        for i, nm in enumerate(("conic_x", "conic_y", "conic_z", "opacity")):
            ex[nm] = pt.bits_equal(sv_new["conic_opacity"][vis][:, i], sv_ref["conic_opacity"][vis][:, i])
        if ex["R_equal"] and f_new["R"] > 0:
            for k in ("keys", "point_list"):                     # default path: tile-bucket binning
                ex[k] = pt.bits_equal(sv_new[k], sv_ref[k])
            newC.SORT_GLOBAL = True                               # reference-style binning: also the unsorted arrays
            f_gl = pt.run_forward(newC, settings, scene, semantic)
            newC.SORT_GLOBAL = False
            sv_gl = newC.state_views(Pn, H, W, f_gl["R"], f_gl["geomBuffer"], f_gl["binningBuffer"], f_gl["imgBuffer"])
            for k in ("keys_unsorted", "point_list_unsorted", "keys", "point_list"):
                ex["global_sort:" + k] = pt.bits_equal(sv_gl[k], sv_ref[k])
        ex["ranges"] = pt.bits_equal(sv_new["ranges"], sv_ref["ranges"])
        ex["n_contrib"] = pt.bits_equal(sv_new["n_contrib"], sv_ref["n_contrib"])
        ex["final_T"] = pt.bits_equal(sv_new["final_T"], sv_ref["final_T"])
        out["bit_mismatches_vs_ref"] = ex
        img = {}
        for k in ("color", "depth", "median_depth", "final_opacity") + (("semantic",) if semantic else ("mask",)):
            img[k] = pt.image_err(f_new[k], f_ref[k])
        out["image_maxabs_and_violations"] = img
        # reference run-to-run noise (atomic order) for scale
        g_ref2 = pt.run_backward(ref._C, settings, scene, f_ref, grads, semantic)
        out["grad_err_vs_ref(normwise,maxrel)"] = {k: pt.grad_err(g_new[k], g_ref[k]) for k in g_new}
        out["ref_run_to_run(normwise,maxrel)"] = {k: pt.grad_err(g_ref2[k], g_ref[k]) for k in g_ref}
        # exact-mode comparison: which Q1 mode did the reference match?
        newC.SEM_ALPHA_GRAD = "exact"
        g_ex = pt.run_backward(newC, settings, scene, f_new, grads, semantic)
        newC.SEM_ALPHA_GRAD = "ref"
        out["grad_err_exactmode_vs_ref"] = {k: pt.grad_err(g_ex[k], g_ref[k]) for k in ("means3D", "opacities", "scales")}
        if iters > 0:
            out["t_ref_fwd"] = timeit(lambda: pt.run_forward(ref._C, settings, scene, semantic), iters)
            out["t_ref_bwd"] = timeit(lambda: pt.run_backward(ref._C, settings, scene, f_ref, grads, semantic), iters)
    else:
        out["reference"] = "oracle/_ref not available"
    if iters > 0:
        out["t_new_fwd"] = timeit(lambda: pt.run_forward(newC, settings, scene, semantic), iters)
        out["t_new_bwd"] = timeit(lambda: pt.run_backward(newC, settings, scene, f_new, grads, semantic), iters)
        newC.SEM_ALPHA_GRAD = "exact"
        out["t_new_bwd_exact"] = timeit(lambda: pt.run_backward(newC, settings, scene, f_new, grads, semantic), iters)
        newC.SEM_ALPHA_GRAD = "ref"
        newC.BWD_SIMT = True
        out["t_new_bwd_simt"] = timeit(lambda: pt.run_backward(newC, settings, scene, f_new, grads, semantic), iters)
        g_simt = pt.run_backward(newC, settings, scene, f_new, grads, semantic)
        newC.BWD_SIMT = False
        out["grad_err_mma_vs_simt"] = {k: pt.grad_err(g_new[k], g_simt[k]) for k in g_new}
        newC.NO_CULL = True
        out["t_new_fwd_nocull"] = timeit(lambda: pt.run_forward(newC, settings, scene, semantic), iters)
        newC.NO_CULL = False
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="c1,c2")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"))
    ap.add_argument("--nonsemantic", action="store_true")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "a") as f:
        for key in a.scenes.split(","):
            t = time.time()
            try:
                r = report(key, a.iters, semantic=not a.nonsemantic)
            except Exception as ex:  # keep going: this is a diagnostic
                import traceback
                r = dict(scene=key, error=repr(ex), tb=traceback.format_exc())
            r["wall_s"] = time.time() - t
            line = json.dumps(r)
            print(line, flush=True)
            f.write(line + "\n")


if __name__ == "__main__":
    main()

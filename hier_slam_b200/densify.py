"""Map maintenance on the flat parameter set: `prune_gaussians` and `densify` with the
contracts of the reference's utils/slam_external.py:171-243, expressed through `FlatAdam.prune` / `FlatAdam.append`
(one keep-mask scan + three gathers, or one re-layout, for ALL parameter tensors and both Adam moments) instead of a
`tensor[mask]` / `torch.cat` per tensor and per moment.

Parameter names follow Hier-SLAM (`means3D`, `unnorm_rotations`, `logit_opacities`, `log_scales`, plus whatever else
lives in the set: colours, semantic embeddings); `variables` holds the per-Gaussian bookkeeping arrays
(`means2D_gradient_accum`, `denom`, `max_2D_radius`, optionally `timestep`) and `scene_radius`.  The decisions
(thresholds, order of clone -> split -> removal, random split offsets drawn with torch.normal) are the reference's, so
with the same torch seed the resulting map is the same.  CUDA only (FlatAdam has no CPU path)."""
from __future__ import annotations

from typing import Dict

import torch

from .optim import FlatAdam

_PER_GAUSSIAN = ("means2D_gradient_accum", "denom", "max_2D_radius", "timestep")


def _rotation_matrices(q: torch.Tensor) -> torch.Tensor:
    """[n,4] unnormalised quaternions (r, x, y, z) -> [n,3,3] (reference build_rotation, utils/slam_external.py:25-42)."""
    q = q / torch.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])[:, None]
    r, x, y, z = q.unbind(1)
    return torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)), dim=1).view(-1, 3, 3)


def _keep(opt: FlatAdam, variables: Dict, to_remove: torch.Tensor) -> None:
    """remove_points (utils/slam_external.py:142-164): parameters + moments through the flat compaction, the small
    per-Gaussian bookkeeping arrays with plain indexing."""
    keep = ~to_remove
    opt.prune(keep)
    for k in _PER_GAUSSIAN:
        if k in variables:
            variables[k] = variables[k][keep]


def _removal_mask(opt: FlatAdam, variables: Dict, threshold: float, remove_big: bool) -> torch.Tensor:
    lv = opt.params.leaves
    to_remove = (torch.sigmoid(lv["logit_opacities"].detach()) < threshold).reshape(-1)
    if remove_big:
        big = torch.exp(lv["log_scales"].detach()).max(dim=1).values > 0.1 * variables["scene_radius"]
        to_remove = torch.logical_or(to_remove, big)
    return to_remove


def _reset_opacities(opt: FlatAdam) -> None:
    """update_params_and_optimizer for logit_opacities (utils/slam_external.py:108-121): value inverse_sigmoid(0.01),
    both Adam moments zero."""
    leaf = opt.params.leaves["logit_opacities"]
    with torch.no_grad():
        new = torch.ones_like(leaf) * 0.01
        leaf.copy_(torch.log(new / (1 - new)))
        m, v = opt.state("logit_opacities")
        m.zero_()
        v.zero_()


@torch.no_grad()
def prune_gaussians(opt: FlatAdam, variables: Dict, iteration: int, prune_dict: Dict) -> None:
    """reference prune_gaussians (utils/slam_external.py:171-193), in place on `opt.params` / `variables`."""
    if iteration > prune_dict["stop_after"]:
        return
    if iteration >= prune_dict["start_after"] and iteration % prune_dict["prune_every"] == 0:
        thr = prune_dict["final_removal_opacity_threshold"] if iteration == prune_dict["stop_after"] \
            else prune_dict["removal_opacity_threshold"]
        _keep(opt, variables, _removal_mask(opt, variables, thr, iteration >= prune_dict["remove_big_after"]))
    if iteration > 0 and iteration % prune_dict["reset_opacities_every"] == 0 and prune_dict["reset_opacities"]:
        _reset_opacities(opt)


@torch.no_grad()
def densify(opt: FlatAdam, variables: Dict, iteration: int, densify_dict: Dict, means2D_grad: torch.Tensor) -> None:
    """reference densify (utils/slam_external.py:196-243).  `means2D_grad` is the gradient of the screen-space means
    (variables['means2D'].grad in the reference); `variables['seen']` selects the Gaussians it is accumulated for."""
    if iteration > densify_dict["stop_after"]:
        return
    seen = variables["seen"]
    variables["means2D_gradient_accum"][seen] += torch.norm(means2D_grad[seen, :2], dim=-1)
    variables["denom"][seen] += 1
    if iteration >= densify_dict["start_after"] and iteration % densify_dict["densify_every"] == 0:
        lv = opt.params.leaves
        dev = lv["means3D"].device
        thresh, radius = densify_dict["grad_thresh"], variables["scene_radius"]
        grads = variables["means2D_gradient_accum"] / variables["denom"]
        grads[grads.isnan()] = 0.0
        small = torch.exp(lv["log_scales"].detach()).max(dim=1).values <= 0.01 * radius
        to_clone = torch.logical_and(grads >= thresh, small)
        opt.append({k: v.detach()[to_clone] for k, v in lv.items()})
        lv = opt.params.leaves
        num = lv["means3D"].shape[0]
        padded = torch.zeros(num, device=dev)
        padded[:grads.shape[0]] = grads
        to_split = torch.logical_and(padded >= thresh, torch.exp(lv["log_scales"].detach()).max(dim=1).values > 0.01 * radius)
        n = densify_dict["num_to_split_into"]
        new_rows = {k: v.detach()[to_split].repeat(n, 1) for k, v in lv.items()}
        stds = torch.exp(lv["log_scales"].detach())[to_split].repeat(n, 3)
        samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds)
        rots = _rotation_matrices(lv["unnorm_rotations"].detach()[to_split]).repeat(n, 1, 1)
        new_rows["means3D"] = new_rows["means3D"] + torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1)
        new_rows["log_scales"] = torch.log(torch.exp(new_rows["log_scales"]) / (0.8 * n))
        opt.append(new_rows)
        num = opt.params.leaves["means3D"].shape[0]
        for k in ("means2D_gradient_accum", "denom", "max_2D_radius"):
            variables[k] = torch.zeros(num, device=dev)
        if "timestep" in variables:           # the reference indexes timestep with the longer mask and would fail here;
            pad = num - variables["timestep"].shape[0]      # new Gaussians inherit no timestep: pad with the newest one
            variables["timestep"] = torch.cat((variables["timestep"], variables["timestep"].max().expand(pad))) \
                if pad > 0 else variables["timestep"]
        to_remove = torch.cat((to_split, torch.zeros(n * int(to_split.sum()), dtype=torch.bool, device=dev)))
        _keep(opt, variables, to_remove)
        thr = densify_dict["final_removal_opacity_threshold"] if iteration == densify_dict["stop_after"] \
            else densify_dict["removal_opacity_threshold"]
        _keep(opt, variables, _removal_mask(opt, variables, thr, iteration >= densify_dict["remove_big_after"]))
    if iteration > 0 and iteration % densify_dict["reset_opacities_every"] == 0 and densify_dict.get("reset_opacities", False):
        _reset_opacities(opt)

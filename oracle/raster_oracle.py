"""TEST INFRASTRUCTURE — dense torch-CPU restatement of the Hier-SLAM rasterizer.

This file is the *oracle* for the hot path named in BASELINE.json: the
differentiable Gaussian rasterizer ``hierslam-diff-gaussian-rasterization-w-depth``
(abbreviated RAST/ below).  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product path (``hier_slam_b200``) never
falls back to it.

Parity status: the reference ships no tests / golden vectors, so
this oracle is pinned by fixtures generated from the *reference CUDA extension
itself* (``oracle/_ref``, built by ``oracle/build_ref.sh``) on a B200 —
``tests/golden/*.npz`` made by ``tests/golden/make_golden.py`` — and by an
autograd cross-check of its explicit backward.

Every function cites the reference code it restates (paths relative to
/root/reference/hierslam-diff-gaussian-rasterization-w-depth).

All maths is evaluated in ``dtype`` (float32 by default; float64 is used by the
tests to calibrate tolerances).  Layout conventions follow the reference:
``viewmatrix`` / ``projmatrix`` are the flat 16-vectors the CUDA code indexes as
``m[4*c + r]`` (column-major w.r.t. the mathematical matrix), i.e. the
``[1,4,4]`` tensors built by ``utils/recon_helpers.py:8-13`` made contiguous.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

BLOCK_X = 16  # cuda_rasterizer/config.h:16
BLOCK_Y = 16  # cuda_rasterizer/config.h:17
BLOCK_SIZE = BLOCK_X * BLOCK_Y

ALPHA_MIN = 1.0 / 255.0
T_MIN = 0.0001
ALPHA_MAX = 0.99
MEDIAN_DEFAULT = 15.0  # forward.cu:450


def _flat16(m: torch.Tensor, dtype) -> torch.Tensor:
    return m.detach().reshape(-1).to("cpu", dtype).contiguous()


def get_higher_msb(n: int) -> int:
    """cuda_rasterizer/rasterizer_impl.cu:35-50 (getHigherMsb)."""
    msb = 4 * 4
    step = msb
    while step > 1:
        step //= 2
        if n >> msb:
            msb += step
        else:
            msb -= step
    if n >> msb:
        msb += 1
    return msb


def tile_grid(W: int, H: int):
    return (W + BLOCK_X - 1) // BLOCK_X, (H + BLOCK_Y - 1) // BLOCK_Y


# --------------------------------------------------------------------------------------
# A1. per-Gaussian forward  (forward.cu:155-256, :118-152, :74-113; auxiliary.h:41-56,139-164)
# --------------------------------------------------------------------------------------
def compute_cov3d(scales, rotations, scale_modifier, dtype):
    """forward.cu:118-152 — Sigma = (S R)^T (S R), quaternion used as given (no normalisation)."""
    s = scales.to(dtype) * scale_modifier
    q = rotations.to(dtype)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    # rows of the standard rotation matrix (glm stores them as columns)
    Rm = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1),
    ], 1)  # [P,3,3]
    M = s[:, :, None] * Rm.transpose(1, 2)  # M = S * Rm^T
    Sigma = M.transpose(1, 2) @ M
    cov3D = torch.stack([Sigma[:, 0, 0], Sigma[:, 0, 1], Sigma[:, 0, 2],
                         Sigma[:, 1, 1], Sigma[:, 1, 2], Sigma[:, 2, 2]], -1)
    return cov3D


def _view_point(p, v):
    """auxiliary.h:58-66 transformPoint4x3."""
    x = v[0] * p[:, 0] + v[4] * p[:, 1] + v[8] * p[:, 2] + v[12]
    y = v[1] * p[:, 0] + v[5] * p[:, 1] + v[9] * p[:, 2] + v[13]
    z = v[2] * p[:, 0] + v[6] * p[:, 1] + v[10] * p[:, 2] + v[14]
    return x, y, z


def _cov2d_terms(p, cov3D, v, fx, fy, tanfovx, tanfovy):
    """forward.cu:74-113 (computeCov2D) — returns a, b, c (low-pass 0.3 added) and intermediates."""
    tx, ty, tz = _view_point(p, v)
    limx = 1.3 * tanfovx
    limy = 1.3 * tanfovy
    txtz = tx / tz
    tytz = ty / tz
    tx = torch.clamp(txtz, -limx, limx) * tz
    ty = torch.clamp(tytz, -limy, limy) * tz
    P = p.shape[0]
    zero = torch.zeros_like(tz)
    # standard Jacobian (glm J is its transpose)
    J = torch.stack([
        torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz)], -1),
        torch.stack([zero, fy / tz, -(fy * ty) / (tz * tz)], -1),
        torch.stack([zero, zero, zero], -1)], 1)
    Rw = torch.stack([torch.stack([v[0], v[4], v[8]]),
                      torch.stack([v[1], v[5], v[9]]),
                      torch.stack([v[2], v[6], v[10]])])  # rotation part of w2c (math layout)
    Vrk = torch.stack([
        torch.stack([cov3D[:, 0], cov3D[:, 1], cov3D[:, 2]], -1),
        torch.stack([cov3D[:, 1], cov3D[:, 3], cov3D[:, 4]], -1),
        torch.stack([cov3D[:, 2], cov3D[:, 4], cov3D[:, 5]], -1)], 1)
    JW = J @ Rw.expand(P, 3, 3)
    cov = JW @ Vrk @ JW.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    return a, b, c, dict(tx=tx, ty=ty, tz=tz, txtz=txtz, tytz=tytz, JW=JW, Vrk=Vrk, Rw=Rw)


def _trunc_int(x: torch.Tensor) -> torch.Tensor:
    """C float->int conversion as CUDA performs it (cvt.rzi.s32.f32: truncate, saturate, NaN->0)."""
    x = torch.nan_to_num(x.to(torch.float64), nan=0.0, posinf=2147483647.0, neginf=-2147483648.0)
    return torch.clamp(torch.trunc(x), -2147483648.0, 2147483647.0).to(torch.int64)


def get_rect(xy: torch.Tensor, radius: torch.Tensor, gx: int, gy: int):
    """auxiliary.h:46-56 (getRect) — float arithmetic on (p -/+ radius), C truncation, clamp to grid."""
    fdt = xy.dtype
    r = radius.to(fdt)
    minx = torch.clamp(_trunc_int((xy[:, 0] - r) / BLOCK_X), 0, gx)
    miny = torch.clamp(_trunc_int((xy[:, 1] - r) / BLOCK_Y), 0, gy)
    maxx = torch.clamp(_trunc_int((xy[:, 0] + r + (BLOCK_X - 1)) / BLOCK_X), 0, gx)
    maxy = torch.clamp(_trunc_int((xy[:, 1] + r + (BLOCK_Y - 1)) / BLOCK_Y), 0, gy)
    return minx, miny, maxx, maxy


def preprocess(means3D, scales, rotations, opacities, viewmatrix, projmatrix, W, H,
               tanfovx, tanfovy, scale_modifier=1.0, cov3D_precomp=None,
               dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """forward.cu:155-256 (preprocessCUDA) with colours precomputed (the only mode Hier-SLAM uses)."""
    p = means3D.detach().to("cpu", dtype)
    P = p.shape[0]
    v = _flat16(viewmatrix, dtype)
    pm = _flat16(projmatrix, dtype)
    gx, gy = tile_grid(W, H)
    fx = W / (2.0 * tanfovx)
    fy = H / (2.0 * tanfovy)
    if dtype == torch.float32:  # the reference computes focal in float (rasterizer_impl.cu:489-490)
        fx = float(np.float32(W) / (np.float32(2.0) * np.float32(tanfovx)))
        fy = float(np.float32(H) / (np.float32(2.0) * np.float32(tanfovy)))

    # in_frustum (auxiliary.h:139-164): only the near test is live
    hx = pm[0] * p[:, 0] + pm[4] * p[:, 1] + pm[8] * p[:, 2] + pm[12]
    hy = pm[1] * p[:, 0] + pm[5] * p[:, 1] + pm[9] * p[:, 2] + pm[13]
    hw = pm[3] * p[:, 0] + pm[7] * p[:, 1] + pm[11] * p[:, 2] + pm[15]
    p_w = 1.0 / (hw + 0.0000001)
    projx = hx * p_w
    projy = hy * p_w
    _, _, vz = _view_point(p, v)
    visible = vz > 0.2

    if cov3D_precomp is not None and cov3D_precomp.numel() > 0:
        cov3D = cov3D_precomp.detach().to("cpu", dtype)
    else:
        cov3D = compute_cov3d(scales.detach().cpu(), rotations.detach().cpu(), scale_modifier, dtype)

    a, b, c, _ = _cov2d_terms(p, cov3D, v, fx, fy, tanfovx, tanfovy)
    det = a * c - b * b
    visible = visible & (det != 0)
    det_inv = 1.0 / det
    conic = torch.stack([c * det_inv, -b * det_inv, a * det_inv], -1)
    mid = 0.5 * (a + c)
    root = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    lam = torch.maximum(mid + root, mid - root)
    radius_f = torch.ceil(3.0 * torch.sqrt(lam))
    # ndc2Pix is evaluated in FP64 and rounded to FP32 (auxiliary.h:41-44)
    px = (((projx.double() + 1.0) * W - 1.0) * 0.5).to(dtype)
    py = (((projy.double() + 1.0) * H - 1.0) * 0.5).to(dtype)
    xy = torch.stack([px, py], -1)
    radius_i = _trunc_int(radius_f)
    minx, miny, maxx, maxy = get_rect(xy, radius_i, gx, gy)
    area = (maxx - minx) * (maxy - miny)
    visible = visible & (area != 0)

    zf = torch.zeros((), dtype=dtype)
    op = opacities.detach().to("cpu", dtype).reshape(-1)
    out = dict(
        visible=visible,
        depths=torch.where(visible, vz, zf),
        radii=torch.where(visible, radius_i, torch.zeros_like(radius_i)).to(torch.int32),
        means2D=torch.where(visible[:, None], xy, zf),
        cov3D=cov3D,
        conic_opacity=torch.where(visible[:, None], torch.cat([conic, op[:, None]], -1), zf),
        tiles_touched=torch.where(visible, area, torch.zeros_like(area)).to(torch.int64),
        radius_prerounding=3.0 * torch.sqrt(lam),
    )
    return out


# --------------------------------------------------------------------------------------
# A2. binning  (rasterizer_impl.cu:70-138, 544-585)
# --------------------------------------------------------------------------------------
def duplicate_with_keys(depths, means2D, radii, W, H):
    """rasterizer_impl.cu:70-111 — instances in Gaussian order, y-major / x-minor inside a rect."""
    gx, gy = tile_grid(W, H)
    radii = radii.to(torch.int64).cpu()
    vis = radii > 0
    idx = torch.nonzero(vis).reshape(-1)
    xy = means2D.detach().cpu()[idx].to(torch.float32)
    minx, miny, maxx, maxy = get_rect(xy, radii[idx], gx, gy)
    w = (maxx - minx)
    h = (maxy - miny)
    cnt = w * h
    total = int(cnt.sum())
    gid = torch.repeat_interleave(torch.arange(idx.numel()), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    local = torch.arange(total) - start[gid]
    ww = torch.clamp(w[gid], min=1)
    ty = miny[gid] + local // ww
    tx = minx[gid] + local % ww
    tile = ty * gx + tx
    dbits = torch.from_numpy(depths.detach().cpu().to(torch.float32).numpy().view(np.uint32).astype(np.int64))
    keys = (tile << 32) | dbits[idx][gid]
    values = idx[gid]
    return keys, values


def sort_and_ranges(keys, values, W, H):
    """rasterizer_impl.cu:570-585,116-138 — stable ascending sort (ties keep Gaussian order), tile ranges."""
    gx, gy = tile_grid(W, H)
    ntiles = gx * gy
    order = torch.from_numpy(np.argsort(keys.numpy(), kind="stable"))
    skeys = keys[order]
    svals = values[order]
    tiles = skeys >> 32
    ranges = torch.zeros(ntiles, 2, dtype=torch.int64)
    if skeys.numel() > 0:
        counts = torch.bincount(tiles, minlength=ntiles)
        ends = torch.cumsum(counts, 0)
        starts = ends - counts
        nz = counts > 0
        ranges[nz, 0] = starts[nz]
        ranges[nz, 1] = ends[nz]
    return skeys, svals, ranges


# --------------------------------------------------------------------------------------
# A3. blend forward  (forward.cu:400-538; non-semantic twin :261-398)
# --------------------------------------------------------------------------------------
def _tile_pixels(t, gx, W, H):
    tx = t % gx
    ty = t // gx
    xs = torch.arange(tx * BLOCK_X, min(tx * BLOCK_X + BLOCK_X, W))
    ys = torch.arange(ty * BLOCK_Y, min(ty * BLOCK_Y + BLOCK_Y, H))
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return xx.reshape(-1), yy.reshape(-1)


def _pair_terms(xx, yy, ids, means2D, conic_opacity, dtype):
    """alpha evaluation shared by forward.cu:485-503 and backward.cu:799-813."""
    xy = means2D[ids]
    co = conic_opacity[ids]
    dx = xy[None, :, 0] - xx.to(dtype)[:, None]
    dy = xy[None, :, 1] - yy.to(dtype)[:, None]
    power = -0.5 * (co[None, :, 0] * dx * dx + co[None, :, 2] * dy * dy) - co[None, :, 1] * dx * dy
    G = torch.exp(power)
    alpha = torch.clamp(co[None, :, 3] * G, max=ALPHA_MAX)
    valid = (power <= 0) & (alpha >= ALPHA_MIN)
    return dx, dy, G, alpha, valid, co


def blend_forward(geom, point_list, ranges, colors, semantics, W, H, dtype=torch.float32,
                  tiles=None) -> Dict[str, torch.Tensor]:
    """Front-to-back alpha compositing; no background term is added (forward.cu:530-531, quirk Q5)."""
    gx, gy = tile_grid(W, H)
    means2D = geom["means2D"].to(dtype)
    conic_opacity = geom["conic_opacity"].to(dtype)
    depths = geom["depths"].to(dtype)
    colors = colors.detach().to("cpu", dtype)
    S = 0 if semantics is None or semantics.numel() == 0 else semantics.shape[1]
    feats = [colors, depths[:, None]]
    if S:
        feats.append(semantics.detach().to("cpu", dtype))
    F = torch.cat(feats, 1)  # [P, 3 + 1 + S]
    N = W * H
    out = torch.zeros(F.shape[1], N, dtype=dtype)
    median = torch.full((N,), MEDIAN_DEFAULT, dtype=dtype)
    final_T = torch.ones(N, dtype=dtype)
    n_contrib = torch.zeros(N, dtype=torch.int64)
    mask = torch.zeros(N, dtype=dtype)
    tile_iter = range(gx * gy) if tiles is None else tiles
    for t in tile_iter:
        r0, r1 = int(ranges[t, 0]), int(ranges[t, 1])
        if r1 <= r0:
            continue
        xx, yy = _tile_pixels(t, gx, W, H)
        pix = yy * W + xx
        ids = point_list[r0:r1]
        dx, dy, G, alpha, valid, _ = _pair_terms(xx, yy, ids, means2D, conic_opacity, dtype)
        a_eff = torch.where(valid, alpha, torch.zeros((), dtype=dtype))
        T_incl = torch.cumprod(1 - a_eff, 1)
        T_excl = torch.cat([torch.ones(len(pix), 1, dtype=dtype), T_incl[:, :-1]], 1)
        stop = valid & (T_incl < T_MIN)
        L = ids.numel()
        ar = torch.arange(L)
        first_stop = torch.where(stop.any(1), torch.argmax(stop.to(torch.int8), 1), torch.full((len(pix),), L))
        contrib = valid & (ar[None, :] < first_stop[:, None])
        w = torch.where(contrib, a_eff * T_excl, torch.zeros((), dtype=dtype))
        out[:, pix] = (w @ F[ids]).T
        mask[pix] = w.sum(1)
        last = torch.where(contrib.any(1), L - 1 - torch.argmax(contrib.flip(1).to(torch.int8), 1),
                           torch.full((len(pix),), -1))
        n_contrib[pix] = last + 1
        # transmittance after the last contributor
        idx_last = torch.clamp(last, min=0)
        fT = torch.where(last >= 0, T_incl.gather(1, idx_last[:, None])[:, 0], torch.ones((), dtype=dtype))
        final_T[pix] = fT
        cross = contrib & (T_excl > 0.5) & (T_incl < 0.5)
        has = cross.any(1)
        jc = torch.argmax(cross.to(torch.int8), 1)
        median[pix] = torch.where(has, depths[ids][jc], median[pix])
    res = dict(
        color=out[0:3].reshape(3, H, W),
        depth=out[3:4].reshape(1, H, W),
        median_depth=median.reshape(1, H, W),
        opacity=(1 - final_T).reshape(1, H, W),
        mask=mask.reshape(1, H, W),
        final_T=final_T, n_contrib=n_contrib,
    )
    res["semantic"] = out[4:].reshape(S, H, W) if S else torch.zeros(0, H, W, dtype=dtype)
    return res


# --------------------------------------------------------------------------------------
# A4. blend backward  (backward.cu:669-899; non-semantic twin :472-666) — explicit, incl. quirks Q1-Q4
# --------------------------------------------------------------------------------------
def blend_backward(geom, point_list, ranges, fwd, colors, semantics, bg, dL_color, dL_sem, dL_depth,
                   dL_median, dL_opacity, W, H, sem_alpha_grad="ref", dtype=torch.float32,
                   tiles=None) -> Dict[str, torch.Tensor]:
    """sem_alpha_grad: 'ref'  = what the reference computes when its never-written scratch buffer
                               is zero (semantic channels give no dL/dalpha; quirk Q1),
                       'exact' = the mathematically intended gradient (s = semantics[id])."""
    gx, gy = tile_grid(W, H)
    means2D = geom["means2D"].to(dtype)
    conic_opacity = geom["conic_opacity"].to(dtype)
    depths = geom["depths"].to(dtype)
    colors = colors.detach().to("cpu", dtype)
    P = colors.shape[0]
    S = 0 if semantics is None or semantics.numel() == 0 else semantics.shape[1]
    sem = semantics.detach().to("cpu", dtype) if S else None
    bg = bg.detach().to("cpu", dtype).reshape(-1)
    N = W * H
    z = lambda x, c: torch.zeros(c, N, dtype=dtype) if x is None else x.detach().to("cpu", dtype).reshape(c, N)
    gC = z(dL_color, 3)
    gD = z(dL_depth, 1)
    gM = z(dL_median, 1)
    gO = z(dL_opacity, 1)
    gS = z(dL_sem, S) if S else None
    final_T = fwd["final_T"].to(dtype)
    n_contrib = fwd["n_contrib"]

    dmean2D = torch.zeros(P, 3, dtype=dtype)
    dconic = torch.zeros(P, 4, dtype=dtype)
    dopac = torch.zeros(P, dtype=dtype)
    dcolors = torch.zeros(P, 3, dtype=dtype)
    dsem = torch.zeros(P, S, dtype=dtype)
    ddepths = torch.zeros(P, dtype=dtype)
    zero = torch.zeros((), dtype=dtype)
    tile_iter = range(gx * gy) if tiles is None else tiles
    for t in tile_iter:
        r0, r1 = int(ranges[t, 0]), int(ranges[t, 1])
        if r1 <= r0:
            continue
        xx, yy = _tile_pixels(t, gx, W, H)
        pix = yy * W + xx
        ids = point_list[r0:r1]
        L = ids.numel()
        dx, dy, G, alpha, valid, co = _pair_terms(xx, yy, ids, means2D, conic_opacity, dtype)
        ar = torch.arange(L)
        contrib = valid & (ar[None, :] < n_contrib[pix][:, None])
        a_eff = torch.where(contrib, alpha, zero)
        om = 1 - a_eff
        # T in front of j, reconstructed from T_final by division back-to-front (backward.cu:815)
        rev_prod = torch.flip(torch.cumprod(torch.flip(om, [1]), 1), [1])  # prod_{k>=j}(1-a_k)
        Tf = final_T[pix]
        T_front = Tf[:, None] / rev_prod
        T_behind = T_front * om
        w = a_eff * T_front
        # features whose 'colour behind' recurrences feed dL/dalpha
        feat = [colors[ids], depths[ids][:, None], torch.ones(L, 1, dtype=dtype)]
        gpix = [gC[:, pix].T, gD[:, pix].T, gO[:, pix].T]
        if S and sem_alpha_grad == "exact":
            feat.append(sem[ids])
            gpix.append(gS[:, pix].T)
        Fm = torch.cat(feat, 1)            # [L, F]
        gp = torch.cat(gpix, 1)            # [npix, F]
        q = gp @ Fm.T                      # q[pix, j] = sum_f f_j * dL_f
        # accum_rec (colour of everything behind j): B_j = sum_{k>j} f_k w_k / T_behind_j, folded with dL
        qw = q * w
        suffix = torch.flip(torch.cumsum(torch.flip(qw, [1]), 1), [1]) - qw   # sum_{k>j}
        behind = torch.where(T_behind > 0, suffix / T_behind, zero)
        dL_dalpha = (q - behind) * T_front
        bg_dot = (gC[:, pix].T * bg[None, :]).sum(1)
        dL_dalpha = dL_dalpha + (-Tf[:, None] / (1 - alpha)) * bg_dot[:, None]
        dL_dalpha = torch.where(contrib, dL_dalpha, zero)
        # per-Gaussian feature gradients  (backward.cu:831,845,852,864)
        dcolors.index_add_(0, ids, w.T @ gC[:, pix].T)
        if S:
            dsem.index_add_(0, ids, w.T @ gS[:, pix].T)
        dd = (w * gD[0, pix][:, None]).sum(0)
        cross = contrib & (T_front > 0.5) & (T_behind < 0.5)         # backward.cu:853-857 (Q4)
        dd = dd + (cross.to(dtype) * gM[0, pix][:, None]).sum(0)
        ddepths.index_add_(0, ids, dd)
        dop = (w * gO[0, pix][:, None]).sum(0)                        # quirk Q2, backward.cu:864
        dop = dop + (G * dL_dalpha).sum(0)                            # backward.cu:896
        dopac.index_add_(0, ids, dop)
        dL_dG = co[None, :, 3] * dL_dalpha                           # straight-through clamp (Q3)
        gdx = G * dx
        gdy = G * dy
        dG_ddelx = -gdx * co[None, :, 0] - gdy * co[None, :, 1]
        dG_ddely = -gdy * co[None, :, 2] - gdx * co[None, :, 1]
        dm = torch.stack([(dL_dG * dG_ddelx * (0.5 * W)).sum(0), (dL_dG * dG_ddely * (0.5 * H)).sum(0),
                          torch.zeros(L, dtype=dtype)], 1)
        dmean2D.index_add_(0, ids, dm)
        dcn = torch.stack([(-0.5 * gdx * dx * dL_dG).sum(0), (-0.5 * gdx * dy * dL_dG).sum(0),
                           torch.zeros(L, dtype=dtype), (-0.5 * gdy * dy * dL_dG).sum(0)], 1)
        dconic.index_add_(0, ids, dcn)
    return dict(dL_dmean2D=dmean2D, dL_dconic=dconic, dL_dopacity=dopac, dL_dcolors=dcolors,
                dL_dsemantics=dsem, dL_ddepths=ddepths)


# --------------------------------------------------------------------------------------
# A5. per-Gaussian backward  (backward.cu:144-274, 278-341, 346-412)
# --------------------------------------------------------------------------------------
def geom_backward(means3D, scales, rotations, scale_modifier, cov3D, radii, viewmatrix, projmatrix, W, H,
                  tanfovx, tanfovy, dL_dmean2D, dL_dconic, dL_ddepths, have_scales=True,
                  dtype=torch.float32) -> Dict[str, torch.Tensor]:
    p = means3D.detach().to("cpu", dtype)
    P = p.shape[0]
    v = _flat16(viewmatrix, dtype)
    pm = _flat16(projmatrix, dtype)
    fx = W / (2.0 * tanfovx)
    fy = H / (2.0 * tanfovy)
    vis = (radii.cpu() > 0)
    cov3D = cov3D.to(dtype)
    a, b, c, aux = _cov2d_terms(p, cov3D, v, fx, fy, tanfovx, tanfovy)
    dcx, dcy, dcz = dL_dconic[:, 0].to(dtype), dL_dconic[:, 1].to(dtype), dL_dconic[:, 3].to(dtype)
    denom = a * c - b * b
    d2 = 1.0 / (denom * denom + 0.0000001)
    dL_da = d2 * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz)
    dL_dc = d2 * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx)
    dL_db = d2 * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz)
    # glm T = W*J  ==  (J_std Rw)^T ; T[c][r] (glm) = JW[c? ...]: use JW rows: JW[0,:] = T[:,0] etc.
    JW = aux["JW"]          # [P,3,3]; row 0 / row 1 are the two live rows
    T0 = JW[:, 0, :]        # (T[0][0], T[1][0], T[2][0]) in glm indexing = first row of J_std Rw
    T1 = JW[:, 1, :]
    # backward.cu:214-224 uses T[0][k], T[1][k] with glm column index first: T[0][k] = (W*J)[col0][row k]
    # col0 of (W*J) = W * J[col0];  J[col0] (glm) = (fx/tz, 0, -fx tx / tz^2) = row 0 of J_std
    # => T[0][k] = sum_m W_math[k][m] * Jstd[0][m] = (Jstd Rw... ) handled via JW^T below
    # In math terms: cov2D = JW Vrk JW^T with JW = J_std Rw; T[i][k] (glm) = JW[i][k].
    dcov = torch.zeros(P, 6, dtype=dtype)
    dcov[:, 0] = T0[:, 0] * T0[:, 0] * dL_da + T0[:, 0] * T1[:, 0] * dL_db + T1[:, 0] * T1[:, 0] * dL_dc
    dcov[:, 3] = T0[:, 1] * T0[:, 1] * dL_da + T0[:, 1] * T1[:, 1] * dL_db + T1[:, 1] * T1[:, 1] * dL_dc
    dcov[:, 5] = T0[:, 2] * T0[:, 2] * dL_da + T0[:, 2] * T1[:, 2] * dL_db + T1[:, 2] * T1[:, 2] * dL_dc
    dcov[:, 1] = 2 * T0[:, 0] * T0[:, 1] * dL_da + (T0[:, 0] * T1[:, 1] + T0[:, 1] * T1[:, 0]) * dL_db + 2 * T1[:, 0] * T1[:, 1] * dL_dc
    dcov[:, 2] = 2 * T0[:, 0] * T0[:, 2] * dL_da + (T0[:, 0] * T1[:, 2] + T0[:, 2] * T1[:, 0]) * dL_db + 2 * T1[:, 0] * T1[:, 2] * dL_dc
    dcov[:, 4] = 2 * T0[:, 2] * T0[:, 1] * dL_da + (T0[:, 1] * T1[:, 2] + T0[:, 2] * T1[:, 1]) * dL_db + 2 * T1[:, 1] * T1[:, 2] * dL_dc
    Vrk = aux["Vrk"]
    TV0 = torch.einsum("pk,pmk->pm", T0, Vrk)   # sum_k T0[k] Vrk[m][k]
    TV1 = torch.einsum("pk,pmk->pm", T1, Vrk)
    dT0 = 2 * TV0 * dL_da[:, None] + TV1 * dL_db[:, None]   # dL_dT00..02
    dT1 = 2 * TV1 * dL_dc[:, None] + TV0 * dL_db[:, None]   # dL_dT10..12
    Rw = aux["Rw"]  # W_glm[c][r] = Rw^T...: W[0][k] (glm col 0) = (v0, v4, v8) = Rw[0, :]
    dJ00 = (Rw[0, :][None, :] * dT0).sum(1)
    dJ02 = (Rw[2, :][None, :] * dT0).sum(1)
    dJ11 = (Rw[1, :][None, :] * dT1).sum(1)
    dJ12 = (Rw[2, :][None, :] * dT1).sum(1)
    tx, ty, tzv = aux["tx"], aux["ty"], aux["tz"]
    limx = 1.3 * tanfovx
    limy = 1.3 * tanfovy
    xm = ((aux["txtz"] >= -limx) & (aux["txtz"] <= limx)).to(dtype)
    ym = ((aux["tytz"] >= -limy) & (aux["tytz"] <= limy)).to(dtype)
    tz = 1.0 / tzv
    tz2 = tz * tz
    tz3 = tz2 * tz
    dtx = xm * -fx * tz2 * dJ02
    dty = ym * -fy * tz2 * dJ12
    dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * tx) * tz3 * dJ02 + (2 * fy * ty) * tz3 * dJ12
    # transformVec4x3Transpose (auxiliary.h:89-97)
    dmean = torch.stack([v[0] * dtx + v[1] * dty + v[2] * dtz,
                         v[4] * dtx + v[5] * dty + v[6] * dtz,
                         v[8] * dtx + v[9] * dty + v[10] * dtz], -1)
    # backward.cu:368-389 — mean2D path through proj
    m = p
    hw = pm[3] * m[:, 0] + pm[7] * m[:, 1] + pm[11] * m[:, 2] + pm[15]
    m_w = 1.0 / (hw + 0.0000001)
    mul1 = (pm[0] * m[:, 0] + pm[4] * m[:, 1] + pm[8] * m[:, 2] + pm[12]) * m_w * m_w
    mul2 = (pm[1] * m[:, 0] + pm[5] * m[:, 1] + pm[9] * m[:, 2] + pm[13]) * m_w * m_w
    g2x = dL_dmean2D[:, 0].to(dtype)
    g2y = dL_dmean2D[:, 1].to(dtype)
    dmean = dmean + torch.stack([
        (pm[0] * m_w - pm[3] * mul1) * g2x + (pm[1] * m_w - pm[3] * mul2) * g2y,
        (pm[4] * m_w - pm[7] * mul1) * g2x + (pm[5] * m_w - pm[7] * mul2) * g2y,
        (pm[8] * m_w - pm[11] * mul1) * g2x + (pm[9] * m_w - pm[11] * mul2) * g2y], -1)
    # backward.cu:391-406 — depth path through the view matrix
    mul3 = v[2] * m[:, 0] + v[6] * m[:, 1] + v[10] * m[:, 2] + v[14]
    gd = dL_ddepths.to(dtype).reshape(-1)
    dmean = dmean + torch.stack([(v[2] - v[3] * mul3) * gd, (v[6] - v[7] * mul3) * gd,
                                 (v[10] - v[11] * mul3) * gd], -1)
    res = dict(dL_dmeans3D=torch.where(vis[:, None], dmean, torch.zeros((), dtype=dtype)),
               dL_dcov3D=torch.where(vis[:, None], dcov, torch.zeros((), dtype=dtype)))
    if have_scales:
        # backward.cu:278-341 (computeCov3D backward); quaternion gradient w.r.t. the *given* q
        s = scales.detach().to("cpu", dtype) * scale_modifier
        q = rotations.detach().to("cpu", dtype)
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        Rm = torch.stack([
            torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
            torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
            torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1)
        M = s[:, :, None] * Rm.transpose(1, 2)           # math M = S Rm^T
        dS = torch.stack([
            torch.stack([dcov[:, 0], 0.5 * dcov[:, 1], 0.5 * dcov[:, 2]], -1),
            torch.stack([0.5 * dcov[:, 1], dcov[:, 3], 0.5 * dcov[:, 4]], -1),
            torch.stack([0.5 * dcov[:, 2], 0.5 * dcov[:, 4], dcov[:, 5]], -1)], 1)
        dM = 2.0 * M @ dS                                # math; glm dL_dM has the same entries
        # glm: Rt = transpose(R_glm) -> Rt[i] (glm col i) = row i of R_glm = (R_glm[0][i],R_glm[1][i],R_glm[2][i])
        #      = (Rm[i][0]...)?  R_glm[c][r] = Rm[c][r] as listed (constructor is column-major and the
        #      listing order is the rows of Rm) => R_glm as a math matrix is Rm^T, Rt_glm math = Rm.
        #      Rt[i] (column i of Rm) ; dL_dMt[i] = column i of dM^T = row i of dM.
        dscale = torch.stack([(Rm[:, :, 0] * dM[:, 0, :]).sum(1),
                              (Rm[:, :, 1] * dM[:, 1, :]).sum(1),
                              (Rm[:, :, 2] * dM[:, 2, :]).sum(1)], -1)
        # dL_dMt[i] *= s_i ; dL_dMt[i][j] = dM[i][j] * s_i   (glm [col][row] of dM^T = dM[i][j])
        D = dM * s[:, :, None]
        d = lambda i, j: D[:, i, j]
        dq = torch.stack([
            2 * z * (d(0, 1) - d(1, 0)) + 2 * y * (d(2, 0) - d(0, 2)) + 2 * x * (d(1, 2) - d(2, 1)),
            2 * y * (d(1, 0) + d(0, 1)) + 2 * z * (d(2, 0) + d(0, 2)) + 2 * r * (d(1, 2) - d(2, 1)) - 4 * x * (d(2, 2) + d(1, 1)),
            2 * x * (d(1, 0) + d(0, 1)) + 2 * r * (d(2, 0) - d(0, 2)) + 2 * z * (d(1, 2) + d(2, 1)) - 4 * y * (d(2, 2) + d(0, 0)),
            2 * r * (d(0, 1) - d(1, 0)) + 2 * x * (d(2, 0) + d(0, 2)) + 2 * y * (d(1, 2) + d(2, 1)) - 4 * z * (d(1, 1) + d(0, 0)),
        ], -1)
        res["dL_dscales"] = torch.where(vis[:, None], dscale, torch.zeros((), dtype=dtype))
        res["dL_drotations"] = torch.where(vis[:, None], dq, torch.zeros((), dtype=dtype))
    return res


# --------------------------------------------------------------------------------------
# Entry points mirroring RAST/rasterize_points.cu:240-432 (semantic) and :35-215 (non-semantic)
# --------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------
# spherical-harmonics colour path (forward.cu:20-71, backward.cu:20-139, auxiliary.h:22-39,107-117)
# ---------------------------------------------------------------------------------------------------
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def _sh_unclamped(deg, dirs, sh):
    """result of forward.cu:30-63 (+0.5) before the clamp; dirs [P,3] unit view directions, sh [P,M,3]."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
                   + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                       + SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11]
                       + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12]
                       + SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                       + SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    return res + 0.5


def sh_forward(deg, means3D, campos, shs, radii, dtype=torch.float32):
    """computeColorFromSH (forward.cu:20-71): returns (rgb [P,3], clamped bool [P,3]); rows of culled Gaussians
    (radii <= 0) are zero (the reference leaves them unwritten)."""
    m, sh = means3D.to(dtype), shs.to(dtype)
    d = m - campos.to(dtype)[None]
    d = d / d.norm(dim=1, keepdim=True)
    res = _sh_unclamped(deg, d, sh)
    vis = (radii > 0)[:, None]
    clamped = (res < 0) & vis
    return torch.where(vis, res.clamp_min(0.0), torch.zeros_like(res)), clamped


def sh_backward(deg, means3D, campos, shs, radii, clamped, dL_dcolors, dtype=torch.float32):
    """computeColorFromSH backward (backward.cu:20-139): returns (dL_dsh [P,M,3], dL_dmeans contribution [P,3]).
    The clamp passes no gradient (PyTorch rule, :30-35); the direction gradient goes through dnormvdv
    (auxiliary.h:107-117).  Explicit restatement of the reference's formulas, not autograd."""
    m, sh = means3D.to(dtype), shs.to(dtype)
    P, M = sh.shape[0], sh.shape[1]
    d0 = m - campos.to(dtype)[None]
    d = d0 / d0.norm(dim=1, keepdim=True)
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    g = torch.where(clamped, torch.zeros_like(dL_dcolors.to(dtype)), dL_dcolors.to(dtype))
    g = torch.where((radii > 0)[:, None], g, torch.zeros_like(g))
    dsh = torch.zeros(P, M, 3, dtype=dtype)
    dsh[:, 0] = SH_C0 * g
    dx = torch.zeros(P, 3, dtype=dtype)
    dy = torch.zeros(P, 3, dtype=dtype)
    dz = torch.zeros(P, 3, dtype=dtype)
    if deg > 0:
        dsh[:, 1], dsh[:, 2], dsh[:, 3] = -SH_C1 * y * g, SH_C1 * z * g, -SH_C1 * x * g
        dx, dy, dz = -SH_C1 * sh[:, 3], -SH_C1 * sh[:, 1], SH_C1 * sh[:, 2]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            dsh[:, 4], dsh[:, 5] = SH_C2[0] * xy * g, SH_C2[1] * yz * g
            dsh[:, 6], dsh[:, 7] = SH_C2[2] * (2.0 * zz - xx - yy) * g, SH_C2[3] * xz * g
            dsh[:, 8] = SH_C2[4] * (xx - yy) * g
            dx = dx + SH_C2[0] * y * sh[:, 4] + SH_C2[2] * 2.0 * -x * sh[:, 6] + SH_C2[3] * z * sh[:, 7] \
                + SH_C2[4] * 2.0 * x * sh[:, 8]
            dy = dy + SH_C2[0] * x * sh[:, 4] + SH_C2[1] * z * sh[:, 5] + SH_C2[2] * 2.0 * -y * sh[:, 6] \
                + SH_C2[4] * 2.0 * -y * sh[:, 8]
            dz = dz + SH_C2[1] * y * sh[:, 5] + SH_C2[2] * 2.0 * 2.0 * z * sh[:, 6] + SH_C2[3] * x * sh[:, 7]
            if deg > 2:
                dsh[:, 9] = SH_C3[0] * y * (3.0 * xx - yy) * g
                dsh[:, 10] = SH_C3[1] * xy * z * g
                dsh[:, 11] = SH_C3[2] * y * (4.0 * zz - xx - yy) * g
                dsh[:, 12] = SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * g
                dsh[:, 13] = SH_C3[4] * x * (4.0 * zz - xx - yy) * g
                dsh[:, 14] = SH_C3[5] * z * (xx - yy) * g
                dsh[:, 15] = SH_C3[6] * x * (xx - 3.0 * yy) * g
                dx = dx + (SH_C3[0] * sh[:, 9] * 3.0 * 2.0 * xy + SH_C3[1] * sh[:, 10] * yz
                           + SH_C3[2] * sh[:, 11] * -2.0 * xy + SH_C3[3] * sh[:, 12] * -3.0 * 2.0 * xz
                           + SH_C3[4] * sh[:, 13] * (-3.0 * xx + 4.0 * zz - yy) + SH_C3[5] * sh[:, 14] * 2.0 * xz
                           + SH_C3[6] * sh[:, 15] * 3.0 * (xx - yy))
                dy = dy + (SH_C3[0] * sh[:, 9] * 3.0 * (xx - yy) + SH_C3[1] * sh[:, 10] * xz
                           + SH_C3[2] * sh[:, 11] * (-3.0 * yy + 4.0 * zz - xx) + SH_C3[3] * sh[:, 12] * -3.0 * 2.0 * yz
                           + SH_C3[4] * sh[:, 13] * -2.0 * xy + SH_C3[5] * sh[:, 14] * -2.0 * yz
                           + SH_C3[6] * sh[:, 15] * -3.0 * 2.0 * xy)
                dz = dz + (SH_C3[1] * sh[:, 10] * xy + SH_C3[2] * sh[:, 11] * 4.0 * 2.0 * yz
                           + SH_C3[3] * sh[:, 12] * 3.0 * (2.0 * zz - xx - yy) + SH_C3[4] * sh[:, 13] * 4.0 * 2.0 * xz
                           + SH_C3[5] * sh[:, 14] * (xx - yy))
    ddir = torch.stack([(dx * g).sum(1), (dy * g).sum(1), (dz * g).sum(1)], 1)
    sum2 = (d0 * d0).sum(1, keepdim=True)
    inv32 = 1.0 / torch.sqrt(sum2 * sum2 * sum2)
    vx, vy, vz = d0[:, 0:1], d0[:, 1:2], d0[:, 2:3]
    gx, gy, gz = ddir[:, 0:1], ddir[:, 1:2], ddir[:, 2:3]
    dmean = torch.cat([((sum2 - vx * vx) * gx - vy * vx * gy - vz * vx * gz) * inv32,
                       (-vx * vy * gx + (sum2 - vy * vy) * gy - vz * vy * gz) * inv32,
                       (-vx * vz * gx - vy * vz * gy + (sum2 - vz * vz) * gz) * inv32], 1)
    return dsh, dmean


def rasterize_forward(background, means3D, colors, semantics, opacity, scales, rotations, scale_modifier,
                      cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
                      dtype=torch.float32):
    """Returns a dict with the reference's outputs plus the internal state the parity tests compare."""
    geom = preprocess(means3D, scales, rotations, opacity, viewmatrix, projmatrix, W, H, tanfovx, tanfovy,
                      scale_modifier, cov3D_precomp, dtype)
    keys, values = duplicate_with_keys(geom["depths"], geom["means2D"], geom["radii"], W, H)
    skeys, point_list, ranges = sort_and_ranges(keys, values, W, H)
    fwd = blend_forward(geom, point_list, ranges, colors, semantics, W, H, dtype)
    return dict(geom=geom, keys_unsorted=keys, values_unsorted=values, keys=skeys, point_list=point_list,
                ranges=ranges, num_rendered=int(keys.numel()), **fwd)


def rasterize_backward(state, background, means3D, colors, semantics, scales, rotations, scale_modifier,
                       cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, H, W,
                       dL_color, dL_sem, dL_depth, dL_median, dL_opacity, sem_alpha_grad="ref",
                       dtype=torch.float32):
    geom = state["geom"]
    bb = blend_backward(geom, state["point_list"], state["ranges"], state, colors, semantics, background,
                        dL_color, dL_sem, dL_depth, dL_median, dL_opacity, W, H, sem_alpha_grad, dtype)
    have_scales = scales is not None and scales.numel() > 0
    gb = geom_backward(means3D, scales, rotations, scale_modifier, geom["cov3D"], geom["radii"], viewmatrix,
                       projmatrix, W, H, tanfovx, tanfovy, bb["dL_dmean2D"], bb["dL_dconic"],
                       bb["dL_ddepths"], have_scales, dtype)
    out = dict(bb)
    out.update(gb)
    return out

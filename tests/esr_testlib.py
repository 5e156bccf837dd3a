"""Shared helpers of the parity tests (golden fixtures, oracle scene dicts, comparisons)."""
from __future__ import annotations

import os

import numpy as np
import torch

from esr_nerf_b200 import synthetic as S
from esr_nerf_b200.modules import voxel_geometry

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["fine_sparse_s20", "fine_dense_s220", "fine_sparse_s60_big"]


def load_case(name):
    fx = dict(np.load(os.path.join(GOLDEN, f"voxurff_{name}.npz")))
    w = dict(np.load(os.path.join(GOLDEN, "fine_weights.npz")))
    return fx, {k: torch.from_numpy(v) for k, v in w.items()}


def cotangents(n, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"srgb/rgb": torch.randn(n, 3, generator=g), "lin/rgb": torch.randn(n, 3, generator=g),
            "etc/alphainv_cum": torch.randn(n, generator=g), "etc/white_bg": torch.randn(n, 1, generator=g)}


def oracle_scene(num_voxels, mask_res, sparse):
    """scene dict consumed by oracle/voxurf_port.py, built with the reference's float32 formulas"""
    from oracle import voxurf_port as P

    voxel_size, world_size = voxel_geometry(S.BBOX_MIN, S.BBOX_MAX, num_voxels)
    stepdist = 0.5 * voxel_size
    return dict(xyz_min=S.BBOX_MIN, xyz_max=S.BBOX_MAX, mask_xyz_min=S.BBOX_MIN, mask_xyz_max=S.BBOX_MAX,
                mask_density_pooled=P.pooled_mask_density(S.mask_density(mask_res, sparse)),
                act_shift=float(np.log(1 / (1 - S.MASK_ALPHA_INIT) - 1)), mask_thres=1e-3, fast_thres=1e-4,
                near=S.NEAR, far=S.FAR, stepdist=float(stepdist), voxel_size=float(voxel_size),
                world_size=[int(w) for w in world_size], grad_feat=[0.5, 1.0, 1.5, 2.0])


def oracle_params(scene, weights, requires_grad=True):
    """grids from the synthetic generator + golden MLP weights, as the dict the port consumes"""
    from oracle import voxurf_port as P

    ws = scene["world_size"]
    sd = dict(weights)
    sd["sdf.grid"] = S.sphere_sdf(ws)
    sd["off_color.grid"] = S.color_grid(ws, 6, 2)
    sd["emo_color.grid"] = S.color_grid(ws, 6, 3)
    leaves = {k: v.clone().float().requires_grad_(requires_grad) for k, v in sd.items()}
    return P.params_from_state_dict(leaves), leaves


def build_product_model(fx, weights, device="cuda:0", **cfg_over):
    from esr_nerf_b200.voxurff import VoxurfF

    cfg = S.fine_cfg(device=device, **cfg_over)
    m = VoxurfF(cfg, S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT,
                S.mask_density(int(fx["mask_res"]), bool(fx["sparse"])), float(fx["s_val"]), int(fx["num_voxels"]))
    m.load_state_dict({**m.state_dict(), **weights})
    S.fill_fine_model(m)
    return m


def digest_check(fx, name, grad: torch.Tensor, rtol, atol_frac=0.1):
    """compare a gradient tensor (logical layout) with a golden digest; returns max scaled error"""
    flat = grad.detach().contiguous().reshape(-1).double().cpu()
    idx = torch.from_numpy(fx[f"grad/{name}/idx"])
    val = torch.from_numpy(fx[f"grad/{name}/val"]).double()
    scale = val.abs().max().clamp_min(1e-12)
    err = ((flat[idx] - val).abs() / (rtol * val.abs() + atol_frac * rtol * scale)).max().item()
    s_err = abs(flat.abs().sum().item() - float(fx[f"grad/{name}/abs_sum"])) / max(float(fx[f"grad/{name}/abs_sum"]), 1e-12)
    return err, s_err


def rel_err(a: torch.Tensor, b: torch.Tensor):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def grad_err(g: torch.Tensor, r: torch.Tensor):
    """(max-abs error / max|ref|, relative L2 error) of a gradient tensor"""
    g, r = g.detach().double().cpu(), r.detach().double().cpu()
    d = g - r
    return (d.abs().max() / r.abs().max().clamp_min(1e-30)).item(), (d.norm() / r.norm().clamp_min(1e-30)).item()


def grad_close(g: torch.Tensor, r: torch.Tensor, tol: float, outlier_frac: float = 0.01, l2_factor: float = 30.0):
    """Gradient comparison that tolerates isolated ReLU-boundary flips.

    The MLP gradients are discontinuous where a hidden pre-activation crosses zero; two fp32 evaluations
    with different summation orders (cuBLAS vs MKL, the reference's own GPU and CPU runs included) flip a
    handful of the ~10^7 ReLU masks of a batch, and each flip perturbs one row of one weight gradient (and
    the 48 grid cells its sample touches) by O(1) of a single sample's contribution.  So: pass if
    max-abs error <= tol * max|ref|; otherwise require that the entries beyond that bound are a small
    fraction (< outlier_frac, or <= 3 entries) of the non-zero entries AND the relative L2 error stays below l2_factor*tol."""
    g, r = g.detach().double().cpu(), r.detach().double().cpu()
    d = (g - r).abs()
    bound = tol * r.abs().max().clamp_min(1e-30)
    if (d <= bound).all():
        return True, "max-rel ok"
    nz = max(int((r != 0).sum()), 1)
    n_out = int((d > bound).sum())
    l2 = float((g - r).norm() / r.norm().clamp_min(1e-30))
    ok = n_out <= max(3, outlier_frac * nz) and l2 < l2_factor * tol      # 3: a flipped unit in a 192-entry bias
    return ok, f"max-rel {float(d.max() / r.abs().max()):.2e} outliers {n_out}/{nz} rel-L2 {l2:.2e}"


# ---------------------------------------------------------------------------------------------------
# coarse stage (VoxurfC)
# ---------------------------------------------------------------------------------------------------
COARSE_CASES = ["coarse_sparse_s5", "coarse_dense_s25"]


def load_coarse_case(name):
    fx = dict(np.load(os.path.join(GOLDEN, f"voxurfc_{name}.npz")))
    w = dict(np.load(os.path.join(GOLDEN, "coarse_weights.npz")))
    return fx, {k: torch.from_numpy(v) for k, v in w.items()}


def coarse_cotangents(n, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"srgb/rgb": torch.randn(n, 3, generator=g), "etc/alphainv_cum": torch.randn(n, generator=g),
            "etc/white_bg": torch.randn(n, 1, generator=g)}


def coarse_oracle_scene(num_voxels, mask_res, sparse):
    from oracle import voxurfc_port as PC

    scene = oracle_scene(num_voxels, mask_res, sparse)
    scene["smooth_kernel"] = PC.gaussian_kernel(5, 0.8)
    return scene


def coarse_oracle_params(scene, weights, requires_grad=True):
    from oracle import voxurfc_port as PC

    ws = scene["world_size"]
    sd = dict(weights)
    sd["sdf.grid"] = S.sphere_sdf(ws)
    sd["off_color.grid"] = S.color_grid(ws, 12, 2)
    sd["emo_color.grid"] = S.color_grid(ws, 12, 3)
    leaves = {k: v.clone().float().requires_grad_(requires_grad) for k, v in sd.items()}
    return PC.params_from_state_dict(leaves), leaves


def build_product_coarse(fx, weights, device="cuda:0", **cfg_over):
    from esr_nerf_b200.voxurfc import VoxurfC

    cfg = S.coarse_cfg(device=device, num_voxels=int(fx["num_voxels"]), **cfg_over)
    m = VoxurfC(cfg, S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT,
                S.mask_density(int(fx["mask_res"]), bool(fx["sparse"])), float(fx["s_val"]))
    m.load_state_dict({**m.state_dict(), **weights})
    S.fill_coarse_model(m)
    return m


# ---------------------------------------------------------------------------------------------------
# alphamask stage (DVGO)
# ---------------------------------------------------------------------------------------------------
DVGO_CASES = ["dvgo_24", "dvgo_40"]


def load_dvgo_case(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def dvgo_cotangents(n, S_, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"etc/alphainv_cum": torch.randn(n, S_ + 1, generator=g), "etc/weights": torch.randn(n, S_, generator=g),
            "etc/white_bg": torch.randn(n, 1, generator=g), "srgb/raw_rgb": 0.1 * torch.randn(n, S_, 3, generator=g),
            "srgb/rgb": torch.randn(n, 3, generator=g)}


def build_product_dvgo(num_voxels, device="cuda:0"):
    from esr_nerf_b200.dvgo import DVGO

    m = DVGO(S.dvgo_cfg(device, num_voxels), S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX).to(device)
    S.fill_dvgo_model(m)
    return m


def dvgo_oracle(num_voxels):
    """(scene, params-with-grad) for oracle/dvgo_port.py on the synthetic alphamask scene"""
    m = build_product_dvgo(num_voxels, "cpu")
    scene = dict(xyz_min=S.BBOX_MIN, xyz_max=S.BBOX_MAX, near=S.NEAR, far=S.FAR, n_samples=m.N_samples, stepsize=0.5,
                 voxel_size=m.voxel_size, act_shift=float(m.act_shift))
    params = {k: getattr(m, k).detach().clone().requires_grad_(True) for k in ("density", "off_color", "emo_color")}
    return scene, params


# ---------------------------------------------------------------------------------------------------
# LTS / PDRA stage (ESRNeRF)
# ---------------------------------------------------------------------------------------------------
ESRNERF_CASES = ["lts_sparse_s220", "pdra_sparse_s60"]


def load_esrnerf_case(name):
    fx = dict(np.load(os.path.join(GOLDEN, f"esrnerf_{name}.npz")))
    w = dict(np.load(os.path.join(GOLDEN, "fine_weights.npz")))
    w.update(np.load(os.path.join(GOLDEN, "lts_weights.npz")))
    return fx, {k: torch.from_numpy(v) for k, v in w.items()}


def esrnerf_oracle_scene(fx):
    scene = oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    scene.update(num_2ndrays=int(fx["num_2ndrays"]), num_ltspts=int(fx["num_ltspts"]), lts_near=1e-5)
    # defaults: random / softplus / interp
    scene.update({k: str(fx[k]) for k in ("ray_sampling", "env_activation", "neus_alpha") if k in fx})
    return scene


def esrnerf_oracle_params(scene, weights, requires_grad=True):
    from oracle import esrnerf_port as E

    ws = scene["world_size"]
    sd = dict(weights)
    sd["sdf.grid"] = S.sphere_sdf(ws)
    sd["off_color.grid"] = S.color_grid(ws, 6, 2)
    sd["emo_color.grid"] = S.color_grid(ws, 6, 3)
    sd["brdf.grid"] = S.color_grid(ws, 6, 4)
    leaves = {k: v.clone().float().requires_grad_(requires_grad) for k, v in sd.items()}
    return E.params_from_state_dict(leaves), leaves


def esrnerf_cotangents(out, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(out[k].shape, generator=g) for k in sorted(out)}


def run_esrnerf_port(fx, weights, draws=None):
    from oracle import esrnerf_port as E

    scene = esrnerf_oracle_scene(fx)
    params, leaves = esrnerf_oracle_params(scene, weights)
    n = int(fx["n_rays"])
    rays = S.make_rays(n, int(fx["ray_seed"]))
    out, inter = E.esrnerf_forward_training(
        scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"], rays["em_modes"], S.uncert_masks(n),
        float(fx["s_val"]), float(fx["normal_eps"]), float(fx["emit_eps"]), bool(fx["pdra_mode"]),
        draws or E.FixedDraws(int(fx["draw_seed"])))
    return out, inter, leaves, rays


def build_product_esrnerf(fx, weights, device="cuda:0"):
    from esr_nerf_b200.esrnerf import ESRNeRF

    cfg = S.lts_cfg(device=device, num_2ndrays=int(fx["num_2ndrays"]), num_ltspts=int(fx["num_ltspts"]),
                    **{k: str(fx[k]) for k in ("ray_sampling", "env_activation", "neus_alpha") if k in fx})
    m = ESRNeRF(cfg, S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT,
                S.mask_density(int(fx["mask_res"]), bool(fx["sparse"])), float(fx["s_val"]), int(fx["num_voxels"]))
    m.load_state_dict({**m.state_dict(), **weights})
    S.fill_esrnerf_model(m)
    m.pdra_mode = bool(fx["pdra_mode"])
    return m

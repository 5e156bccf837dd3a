"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (torch ops + the C restatement of the native ops) of the reference's coarse-stage render function
``VoxurfC.forward_training`` (app/coarse/model/voxurfc.py:186-271), written so that it can travel to the GPU box.

Parity pin: ``tests/test_oracle_cpu.py::test_coarse_port_matches_reference`` runs it against the reference's OWN
``VoxurfC`` class (imported through ``oracle/ref_harness.py``) whenever ``/root/reference`` is present, and against the
committed golden vectors ``tests/golden/voxurfc_*.npz`` (produced by the reference's own code,
``oracle/make_golden.py``) everywhere else.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import voxurf_port as P


def gaussian_kernel(ksize: int, sigma: float) -> torch.Tensor:
    """module.py:146-177 (Gaussian3DConv weights)"""
    r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
    xx, yy, zz = np.meshgrid(r, r, r)
    k = torch.FloatTensor(np.exp(-(xx ** 2 + yy ** 2 + zz ** 2) / (2 * sigma ** 2)))
    return (k / k.sum())[None, None]


def smooth_conv(grid: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """nn.Conv3d(1, 1, k, padding=k//2, padding_mode='replicate'), zero bias (module.py:159-177)"""
    p = kernel.shape[-1] // 2
    return F.conv3d(F.pad(grid, (p,) * 6, mode="replicate"), kernel)


def neus_sdf_gradient(sdf_grid: torch.Tensor, voxel_size: float) -> torch.Tensor:
    """voxurfc.py:597-616: dense central differences of the RAW sdf grid, channels (d/dx, d/dy, d/dz)"""
    g = torch.zeros([1, 3, *sdf_grid.shape[-3:]])
    g[:, 0, 1:-1, :, :] = (sdf_grid[:, 0, 2:, :, :] - sdf_grid[:, 0, :-2, :, :]) / 2 / voxel_size
    g[:, 1, :, 1:-1, :] = (sdf_grid[:, 0, :, 2:, :] - sdf_grid[:, 0, :, :-2, :]) / 2 / voxel_size
    g[:, 2, :, :, 1:-1] = (sdf_grid[:, 0, :, :, 2:] - sdf_grid[:, 0, :, :, :-2]) / 2 / voxel_size
    return g


def _alpha(scene, viewdirs, ray_id, sdf, gradient, s_val):
    """self.neus_alpha_from_sdf_scatter (voxurfc.py:171-174): 'interp', or 'grad' with the trilinear tap of the
    central-difference volume as the SDF gradient (voxurfc.py:204-210) and dist = stepsize * voxel_size"""
    if scene.get("neus_alpha", "interp") == "grad":
        return P.neus_alpha_grad(viewdirs, ray_id, torch.tensor(scene["stepdist"], dtype=torch.float32), sdf, gradient, s_val)
    return P.neus_alpha_interp(ray_id, sdf, s_val)


def voxurfc_forward_training(scene: Dict, params: Dict, rays_o, rays_d, viewdirs, em_modes, s_val: float):
    N = rays_o.shape[0]
    ray_pts, ray_id, step_id, aux = P._march(scene, rays_o, rays_d)
    inter = dict(aux)
    inter["m0"] = int(ray_pts.shape[0])
    keep = P.mask_cache(scene, ray_pts)
    ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
    inter.update(m1_ray=ray_id, m1_step=step_id)

    sdf_grid = smooth_conv(params["sdf"], scene["smooth_kernel"])                       # voxurfc.py:202
    sdf = P.grid_sample_world(sdf_grid, ray_pts, scene["xyz_min"], scene["xyz_max"])[:, 0]
    grad_vol = neus_sdf_gradient(params["sdf"], scene["voxel_size"])                    # voxurfc.py:205
    gradient = P.grid_sample_world(grad_vol, ray_pts, scene["xyz_min"], scene["xyz_max"])
    alpha = _alpha(scene, viewdirs, ray_id, sdf, gradient, s_val)
    inter.update(m1_sdf=sdf, m1_alpha=alpha)
    weights, _ = P._A2W.apply(alpha, ray_id, N)                                         # voxurfc.py:211
    k1 = weights > scene["fast_thres"]
    ray_pts, ray_id, step_id, alpha, gradient = ray_pts[k1], ray_id[k1], step_id[k1], alpha[k1], gradient[k1]
    weights, last = P._A2W.apply(alpha, ray_id, N)                                      # voxurfc.py:219 (recomputed)
    inter.update(m3_ray=ray_id, m3_step=step_id, m3_weights=weights)

    on = em_modes[ray_id] == 1
    u = (ray_pts - scene["xyz_min"]) / (scene["xyz_max"] - scene["xyz_min"])
    freq = torch.tensor([2.0 ** i for i in range(5)])
    emb = (u.unsqueeze(-1) * freq).flatten(-2)
    vemb = (viewdirs.unsqueeze(-1) * torch.tensor([1.0])).flatten(-2)
    normal = gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-5)
    rgb_feat = torch.cat([u, emb.sin(), emb.cos(), vemb[ray_id], vemb.sin()[ray_id], vemb.cos()[ray_id], normal], -1)
    off_c = P.grid_sample_world(params["off_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    emo_c = P.grid_sample_world(params["emo_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    inter.update(m3_feat=rgb_feat, m3_normal=normal)
    rgb_emo = torch.sigmoid(P.mlp(torch.cat([emo_c, rgb_feat], -1), params["emo_rgbnet"], lambda t: t))
    rgb_off = torch.sigmoid(P.mlp(torch.cat([off_c, rgb_feat], -1), params["off_rgbnet"], lambda t: t))
    rgb = torch.where(on[:, None], rgb_emo, torch.zeros_like(rgb_emo)) + rgb_off        # voxurfc.py:241-249
    inter.update(m3_rgb=rgb)
    w_ = weights[:, None]
    out = {
        "etc/alphainv_cum": last,
        "etc/white_bg": 1 - torch.zeros(N, 1).index_add(0, ray_id, w_),
        "srgb/rgb": torch.zeros(N, 3).index_add(0, ray_id, w_ * rgb),
    }
    return out, inter


def voxurfc_forward_evaluate(scene: Dict, params: Dict, rays_o, rays_d, viewdirs, em_modes, pos_rt, s_val: float):
    """voxurfc.py:273-424"""
    from . import ref_harness as H

    N = rays_o.shape[0]
    ray_pts, ray_id, step_id, _ = P._march(scene, rays_o, rays_d)
    keep = P.mask_cache(scene, ray_pts)
    ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
    sdf_grid = smooth_conv(params["sdf"], scene["smooth_kernel"])
    sdf = P.grid_sample_world(sdf_grid, ray_pts, scene["xyz_min"], scene["xyz_max"])[:, 0]
    gradient = P.grid_sample_world(neus_sdf_gradient(params["sdf"], scene["voxel_size"]), ray_pts, scene["xyz_min"],
                                   scene["xyz_max"])
    alpha = _alpha(scene, viewdirs, ray_id, sdf, gradient, s_val)
    weights = H.alpha2weight(alpha, ray_id, N)[0]
    k1 = weights > scene["fast_thres"]
    ray_pts, ray_id, step_id, alpha, gradient = ray_pts[k1], ray_id[k1], step_id[k1], alpha[k1], gradient[k1]
    if int(k1.sum()) <= 1:                                                              # voxurfc.py:322-335
        z3 = torch.zeros_like(rays_o)
        return {"etc/depth": z3[..., 0], "etc/disp": 1 / (z3[..., 0] + scene["far"]), "etc/normal": z3,
                "etc/white_bg": torch.ones_like(z3[..., :1]), "srgb/off_rgb": z3, "srgb/emo_rgb": z3, "srgb/on_rgb": z3,
                "srgb/rgb": z3}, dict(m3_ray=ray_id, m3_step=step_id)
    weights = H.alpha2weight(alpha, ray_id, N)[0]
    u = (ray_pts - scene["xyz_min"]) / (scene["xyz_max"] - scene["xyz_min"])
    freq = torch.tensor([2.0 ** i for i in range(5)])
    emb = (u.unsqueeze(-1) * freq).flatten(-2)
    vemb = (viewdirs.unsqueeze(-1) * torch.tensor([1.0])).flatten(-2)
    normal = gradient / (gradient.norm(dim=-1, keepdim=True) + 1e-5)
    rgb_feat = torch.cat([u, emb.sin(), emb.cos(), vemb[ray_id], vemb.sin()[ray_id], vemb.cos()[ray_id], normal], -1)
    off_c = P.grid_sample_world(params["off_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    emo_c = P.grid_sample_world(params["emo_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    off = torch.sigmoid(P.mlp(torch.cat([off_c, rgb_feat], -1), params["off_rgbnet"], lambda t: t))
    emo = torch.sigmoid(P.mlp(torch.cat([emo_c, rgb_feat], -1), params["emo_rgbnet"], lambda t: t))
    w_ = weights[:, None]

    def comp(x):
        return torch.zeros(N, x.shape[1]).index_add(0, ray_id, w_ * x)

    nrm = ((normal @ pos_rt) * torch.tensor([1.0, -1.0, -1.0]) + 1.0) / 2.0
    depth = torch.zeros(N).index_add(0, ray_id, weights * step_id * scene["stepdist"])
    bg = 1 - comp(torch.ones_like(w_))
    out = {"etc/depth": depth, "etc/disp": 1 / (depth + bg[..., -1] * scene["far"]), "etc/normal": comp(nrm),
           "etc/white_bg": bg, "srgb/off_rgb": comp(off), "srgb/emo_rgb": comp(emo), "srgb/on_rgb": comp(off + emo)}
    out["srgb/rgb"] = out["srgb/off_rgb"] if int(em_modes) == 0 else out["srgb/on_rgb"]
    return out, dict(m3_ray=ray_id, m3_step=step_id, m3_weights=weights)


def params_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict:
    def net(prefix):
        return [(sd[f"{prefix}.{i}.weight"].float(), sd[f"{prefix}.{i}.bias"].float()) for i in ("0", "2.0", "3")]

    return {"sdf": sd["sdf.grid"].float().contiguous(), "off_color": sd["off_color.grid"].float().contiguous(),
            "emo_color": sd["emo_color.grid"].float().contiguous(), "off_rgbnet": net("off_rgbnet"),
            "emo_rgbnet": net("emo_rgbnet")}

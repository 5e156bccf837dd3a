"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (pure torch, like the reference) of the alphamask-stage render functions
``DVGO.forward_training`` / ``forward_evaluate`` (app/coarse/model/dvgo.py:140-288), with the per-ray sampler jitter
(``torch.rand_like``, dvgo.py:163) as an explicit argument so that the CUDA path can be compared on identical samples.

Parity pin: ``tests/test_oracle_cpu.py::test_dvgo_port_matches_reference`` runs it against the reference's OWN ``DVGO``
class (pure torch, imported through ``oracle/ref_harness.py``) under the same RNG seed whenever ``/root/reference`` is
present, and against ``tests/golden/dvgo_*.npz`` (produced by the reference's own code) everywhere else.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .voxurf_port import grid_sample_world


def sample_ray(scene: Dict, rays_o, rays_d, jitter):
    """dvgo.py:140-172"""
    vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
    rate_a = (scene["xyz_max"] - rays_o) / vec
    rate_b = (scene["xyz_min"] - rays_o) / vec
    t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=scene["near"], max=scene["far"])
    t_max = torch.maximum(rate_a, rate_b).amin(-1).clamp(min=scene["near"], max=scene["far"])
    mask_outbbox = t_max <= t_min
    rng = torch.arange(scene["n_samples"])[None].float().repeat(rays_d.shape[-2], 1)
    if jitter is not None:
        rng = rng + jitter.reshape(-1, 1)
    step = scene["stepsize"] * scene["voxel_size"] * rng
    interpx = t_min[..., None] + step / rays_d.norm(dim=-1, keepdim=True)
    rays_pts = rays_o[..., None, :] + rays_d[..., None, :] * interpx[..., None]
    mask_outbbox = mask_outbbox[..., None] | ((scene["xyz_min"] > rays_pts) | (rays_pts > scene["xyz_max"])).any(dim=-1)
    return rays_pts, mask_outbbox


def _sampler(scene, pts, grid):
    shape = pts.shape[:-1]
    return grid_sample_world(grid, pts.reshape(-1, 3), scene["xyz_min"], scene["xyz_max"]).reshape(*shape, grid.shape[1])


def _alpha_weights(scene, params, rays_pts, mask_outbbox):
    alpha = torch.zeros_like(rays_pts[..., 0])
    density = _sampler(scene, rays_pts[~mask_outbbox], params["density"])[..., 0]
    alpha[~mask_outbbox] = 1 - torch.exp(-F.softplus(density + scene["act_shift"]) * scene["stepsize"])
    p = 1 - alpha
    cum = torch.cat([torch.ones_like(p[..., [0]]), p.clamp_min(1e-10).cumprod(-1)], -1)      # dvgo.py:286-288
    return alpha * cum[..., :-1], cum


def dvgo_forward_training(scene: Dict, params: Dict, rays_o, rays_d, em_modes, jitter):
    rays_pts, mask_outbbox = sample_ray(scene, rays_o, rays_d, jitter)
    weights, cum = _alpha_weights(scene, params, rays_pts, mask_outbbox)
    on = em_modes == 1
    rgb = torch.zeros_like(rays_pts)
    rgb[on] = torch.sigmoid(_sampler(scene, rays_pts[on], params["emo_color"]))
    rgb = rgb + torch.sigmoid(_sampler(scene, rays_pts, params["off_color"]))
    return {"etc/alphainv_cum": cum, "etc/weights": weights, "etc/white_bg": cum[..., [-1]], "srgb/raw_rgb": rgb,
            "srgb/rgb": (weights.unsqueeze(-1) * rgb).sum(-2)}


def dvgo_forward_evaluate(scene: Dict, params: Dict, rays_o, rays_d, em_modes):
    rays_pts, mask_outbbox = sample_ray(scene, rays_o, rays_d, None)
    weights, cum = _alpha_weights(scene, params, rays_pts, mask_outbbox)
    off = torch.sigmoid(_sampler(scene, rays_pts, params["off_color"]))
    emo = torch.sigmoid(_sampler(scene, rays_pts, params["emo_color"]))
    w_ = weights.unsqueeze(-1)
    depth = (weights * (rays_o[..., None, :] - rays_pts).norm(dim=-1)).sum(-1)
    out = {"etc/depth": depth, "etc/disp": 1 / (depth + cum[..., -1] * scene["far"]), "etc/white_bg": cum[..., [-1]],
           "srgb/off_rgb": (w_ * off).sum(-2), "srgb/on_rgb": (w_ * (off + emo)).sum(-2), "srgb/emo_rgb": (w_ * emo).sum(-2)}
    out["srgb/rgb"] = out["srgb/off_rgb"] if int(em_modes) == 0 else out["srgb/on_rgb"]
    return out

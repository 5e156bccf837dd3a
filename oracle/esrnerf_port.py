"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (torch ops + the C restatement of the native ops) of the reference's LTS / PDRA stage render
function, written so that it can travel to the GPU box where ``/root/reference`` does not exist:

* ``esrnerf_forward_training`` — app/fine/model/esrnerf.py:487-851 (``ESRNeRF.forward_training`` with its
  ``light_transport_segment`` closure), incl. ``sample_sdf_expgrad`` (:1572-1596) over
  ``differentiable_grid_sample`` (app/utils/base/functions.py:142-309), ``disney_reflection`` and
  ``diffuse_scattering`` (app/utils/pbr/functions.py:10-18,108-173), ``SphericalGaussian.forward``,
  ``EmissionNet`` / ``BRDFNet`` (app/utils/pbr/module.py:42-83,133-143).

Parity pin: ``tests/test_oracle_cpu.py::test_esrnerf_port_matches_reference`` runs this port against the
reference's OWN ``ESRNeRF`` class (imported through ``oracle/ref_harness.py``) on identical weights, rays and
random draws whenever ``/root/reference`` is present, and against the committed golden vectors
``tests/golden/esrnerf_*.npz`` (produced by the reference's own code, ``oracle/make_golden.py``) everywhere else.

Random draws.  The reference makes four per step, in this order (esrnerf.py:792, pbr/functions.py:14, esrnerf.py:808,
813): ``np.random.choice`` of the LTS points, ``torch.randn(P, n2+1, 3)`` for the hemisphere directions and two
``torch.randn_like(ray_pts)`` jitters.  ``Draws`` reproduces exactly those calls (so seeding numpy / torch the same
way gives the reference's numbers); ``FixedDraws`` serves pre-drawn tensors so that the CUDA path and the port see
identical numbers on different devices.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import voxurf_port as P


# ---------------------------------------------------------------------------------------------
# random draws
# ---------------------------------------------------------------------------------------------
_randn, _randperm = torch.randn, torch.randperm   # captured: make_golden patches torch.randn around the reference call


class Draws:
    """the reference's own calls, in the reference's order"""

    def choice(self, n: int, k: int) -> torch.Tensor:
        return torch.from_numpy(np.random.choice(n, k, replace=False)).long()

    def randn(self, *shape) -> torch.Tensor:
        return torch.randn(*shape)


class FixedDraws:
    """seeded draws that do not depend on the device: generated on the CPU, moved by the consumer"""

    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def choice(self, n: int, k: int) -> torch.Tensor:
        return _randperm(n, generator=self.g)[:k]

    def randn(self, *shape) -> torch.Tensor:
        return _randn(*shape, generator=self.g)


# ---------------------------------------------------------------------------------------------
# functions.py:142-309 through esrnerf.py:1572-1596
# ---------------------------------------------------------------------------------------------
def manual_trilinear(grid: torch.Tensor, xyz: torch.Tensor, xyz_min, xyz_max) -> torch.Tensor:
    """differentiable_grid_sample(grid, ind_norm) for a [1,1,X,Y,Z] grid at world points [M,3] -> [M].
    Weights come from the un-clamped floor, corner indices are clamped to the grid, corners are summed in the order
    tnw, tne, tsw, tse, bnw, bne, bsw, bse (value * weight, plain adds)."""
    X, Y, Z = grid.shape[2:]
    ind_norm = ((xyz - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1
    ix = ((ind_norm[:, 0] + 1) / 2) * (Z - 1)      # fastest axis ("width")
    iy = ((ind_norm[:, 1] + 1) / 2) * (Y - 1)
    iz = ((ind_norm[:, 2] + 1) / 2) * (X - 1)      # slowest axis ("depth")
    with torch.no_grad():
        x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    flat = grid.reshape(-1)
    out = None
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                wx = (ix - x0) if dx else (x0 + 1 - ix)
                wy = (iy - y0) if dy else (y0 + 1 - iy)
                wz = (iz - z0) if dz else (z0 + 1 - iz)
                with torch.no_grad():
                    cx = (x0 + dx).clamp(0, Z - 1)
                    cy = (y0 + dy).clamp(0, Y - 1)
                    cz = (z0 + dz).clamp(0, X - 1)
                    idx = (cz * Z * Y + cy * Z + cx).long()
                term = flat[idx] * (wx * wy * wz)
                out = term if out is None else out + term
    return out


def sdf_expgrad(grid, xyz, xyz_min, xyz_max, create_graph: bool):
    """esrnerf.py:1572-1596: (sdf, d sdf.sum() / d xyz) with the graph kept so that the gradient itself is
    differentiable w.r.t. the grid."""
    with torch.enable_grad():
        x = xyz.detach().clone().requires_grad_(True)
        sdf = manual_trilinear(grid, x, xyz_min, xyz_max)
        (g,) = torch.autograd.grad(sdf.sum(), x, retain_graph=create_graph, create_graph=create_graph)
    return sdf, g


# ---------------------------------------------------------------------------------------------
# pbr
# ---------------------------------------------------------------------------------------------
def diffuse_scattering(normal: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """pbr/functions.py:10-18 with the Gaussian draw passed in: normalise, flip into the normal's hemisphere."""
    ret = F.normalize(noise, dim=-1)
    flip = (ret * normal.unsqueeze(-2)).sum(-1) < 0
    return torch.where(flip[..., None], ret * -1.0, ret)


def _is_fib(scene) -> bool:
    """esrnerf.py:188-192: cfg.app.model.ray_sampling (scene["ray_sampling"], default "random")"""
    return str(scene.get("ray_sampling", "random")).lower() in ("fib", "fibo", "fibonacci")


def scatter_noise(scene, draws, *shape):
    """the Gaussian draw behind `self.scattering` — none at all with the Fibonacci sampler (pbr/functions.py:21-32)"""
    return None if _is_fib(scene) else draws.randn(*shape)


def scatter_dirs(scene, normal: torch.Tensor, number: int, noise) -> torch.Tensor:
    """self.scattering(normal, number): diffuse_scattering (pbr/functions.py:10-18) or diffuse_scattering_fib (:21-32,
    176-194: the upper half of a 2n-point Fibonacci spiral, the same for every point, mirrored by the normal)"""
    if not _is_fib(scene):
        return diffuse_scattering(normal, noise)
    import math

    n = 2 * number
    rn = torch.arange(number, n)
    phi = (math.pi * (3.0 - math.sqrt(5.0))) * ((rn + 1.0) % n)
    cos_theta = ((rn + 0.5) * (1.0 / number)) - 1.0
    sin_theta = torch.sqrt(1.0 - cos_theta * cos_theta)
    ret = torch.stack([torch.cos(phi) * sin_theta, torch.sin(phi) * sin_theta, cos_theta], dim=-1)
    ret = ret.expand(*normal.shape[:-1], number, 3).clone().to(normal.device)
    ret[torch.sum(ret * normal.unsqueeze(-2), dim=-1) < 0] *= -1.0
    return ret


def disney_reflection(albedo, roughness, metallic, normal, win, wout):
    """pbr/functions.py:108-173"""
    eps = 1e-7

    def dot(a, b):
        return (a * b).sum(-1, keepdim=True)

    h = F.normalize(win + wout, dim=-1)
    noh = dot(normal, h).clamp(min=0)
    ooh = dot(wout, h).clamp(min=0)
    ion = dot(win, normal).clamp(min=0)
    oon = dot(wout, normal).clamp(min=0)
    fd = (1 - metallic) * albedo / torch.pi
    r2 = (roughness * roughness).clamp(min=eps)
    D = (1 / (r2 * np.pi)) * torch.exp((2 / r2) * (noh - 1))
    F0 = 0.04 * (1 - metallic) + albedo * metallic
    Fr = F0 + (1.0 - F0) * ((1.0 - ooh) ** 5)

    def v_ggx(cos):
        k = ((1 + roughness) ** 2) / 8
        return 0.5 / (cos * (1 - k) + k).clamp(min=eps)

    V = v_ggx(ion) * v_ggx(oon)
    return (fd + D * Fr * V) * ion * torch.pi * 2


def sg_envmap(params: Dict, dirs: torch.Tensor, activation: str = "softplus") -> torch.Tensor:
    """pbr/module.py:133-143; the activation is looked up in torch, then torch.nn.functional (:94-101; softplus:
    cfg/app/lts.yaml:30)"""
    act = getattr(torch, activation) if hasattr(torch, activation) else getattr(F, activation)
    lobes = F.normalize(params["envmap.lobes"], dim=-1)
    lambdas = torch.abs(params["envmap.lambdas"])
    return act((params["envmap.mus"] * torch.exp(
        lambdas * ((dirs.unsqueeze(-2) * lobes).sum(-1, keepdim=True) - 1.0))).sum(-2))


# ---------------------------------------------------------------------------------------------
# render pieces
# ---------------------------------------------------------------------------------------------
def _march_near(scene, rays_o, rays_d, near):
    sc = dict(scene, near=near)
    return P._march(sc, rays_o, rays_d)


def _pos_emb(scene, pts):
    u = (pts - scene["xyz_min"]) / (scene["xyz_max"] - scene["xyz_min"])
    freq = torch.tensor([2.0 ** i for i in range(5)])
    emb = (u.unsqueeze(-1) * freq).flatten(-2)
    return torch.cat([u, emb.sin(), emb.cos()], -1)


def _taps(scene, params, pts):
    """esrnerf.py:1527-1570 (finite-difference denominator + 1e-12)"""
    return P.sdf_feature_taps(scene, params["sdf"], pts, scene["grad_feat"], fd_eps=1e-12)


def _sample(grid, scene, pts):
    return P.grid_sample_world(grid, pts, scene["xyz_min"], scene["xyz_max"])


def _alpha(scene, params, viewdirs, ray_id, ray_pts, sdf, s_val):
    """self.neus_alpha_from_sdf_scatter (esrnerf.py:197-200): 'interp' or 'grad' — the latter from sample_sdf_grad's
    finite differences (esrnerf.py:1519-1525, denominator + 1e-12) along the ray's view direction"""
    return P.neus_alpha(scene, params["sdf"], viewdirs, ray_id, ray_pts, sdf, s_val, fd_eps=1e-12)


def _brdf_split(y):
    return y[:, 0:3], y[:, 3:4], y[:, 4:5]


def _secondary(scene, params, rays_o, dirs, s_val):
    """the second pass of the render chain over the LTS rays (esrnerf.py:576-652)"""
    N = rays_o.shape[0]
    ray_pts, ray_id, step_id, _ = _march_near(scene, rays_o, dirs, scene["lts_near"])
    keep = P.mask_cache(scene, ray_pts)
    ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
    sdf = _sample(params["sdf"], scene, ray_pts)[:, 0]
    alpha = _alpha(scene, params, dirs, ray_id, ray_pts, sdf, s_val)      # esrnerf.py:346-361: viewdirs = dirs
    k0 = alpha > scene["fast_thres"]
    alpha, ray_id, step_id, ray_pts, sdf = alpha[k0], ray_id[k0], step_id[k0], ray_pts[k0], sdf[k0]
    weights, last = P._A2W.apply(alpha, ray_id, N)
    k1 = weights > scene["fast_thres"]
    weights, ray_id, step_id, ray_pts, sdf = weights[k1], ray_id[k1], step_id[k1], ray_pts[k1], sdf[k1]
    feat, _, normal = _taps(scene, params, ray_pts)
    xyz_emb = _pos_emb(scene, ray_pts)
    v = dirs[ray_id]
    rgb_feat = torch.cat([xyz_emb, v, v.sin(), v.cos(), sdf[:, None], feat, normal], -1)
    lin_off = P.mlp(torch.cat([_sample(params["off_color"], scene, ray_pts), rgb_feat], -1), params["off_rgbnet"], F.softplus)
    lin_emo = P.mlp(torch.cat([_sample(params["emo_color"], scene, ray_pts), rgb_feat], -1), params["emo_rgbnet"], F.softplus)
    w_ = weights[:, None]
    off_m = torch.zeros(N, 3).index_add(0, ray_id, w_ * lin_off)
    emo_m = torch.zeros(N, 3).index_add(0, ray_id, w_ * lin_emo)
    return off_m, emo_m, last, dict(m3_ray=ray_id, m3_step=step_id, m3_weights=weights)


def light_transport_segment(scene, params, pts, viewdirs, normal, sdf, base, rough, metal, emission, umask, s_val,
                            pdra_mode, dir_noise):
    """esrnerf.py:487-679"""
    n2 = scene["num_2ndrays"]
    Pn = pts.shape[0]
    dirs = scatter_dirs(scene, normal, scene["num_2ndrays"] + 1, dir_noise)           # [P, n2+1, 3]
    v_rand = -dirs[:, -1]
    dirs = dirs[:, :-1]
    # radiance leaving the points towards the camera and towards one random direction
    feat, _, fnormal = _taps(scene, params, pts)
    xyz_emb = _pos_emb(scene, pts)
    v2 = torch.cat([viewdirs, v_rand], 0)
    rgb_feat = torch.cat([xyz_emb.repeat(2, 1), v2, v2.sin(), v2.cos(), sdf[:, None].repeat(2, 1), feat.repeat(2, 1),
                          fnormal.repeat(2, 1)], -1)
    off = P.mlp(torch.cat([_sample(params["off_color"], scene, pts).repeat(2, 1), rgb_feat], -1), params["off_rgbnet"], F.softplus)
    emo = P.mlp(torch.cat([_sample(params["emo_color"], scene, pts).repeat(2, 1), rgb_feat], -1), params["emo_rgbnet"], F.softplus)

    def ex(t, c):
        return t.view(-1, 1, c).expand(Pn, n2, c).flatten(0, 1)

    wout = torch.cat([-ex(viewdirs, 3), -ex(v_rand, 3)], 0)
    d_flat = dirs.flatten(0, 1)
    R = disney_reflection(ex(base, 3).repeat(2, 1), ex(rough, 1).repeat(2, 1), ex(metal, 1).repeat(2, 1),
                          ex(normal, 3).repeat(2, 1), d_flat.repeat(2, 1), wout)
    off_m, emo_m, last, inter = _secondary(scene, params, ex(pts, 3), d_flat, s_val)
    env = sg_envmap(params, d_flat, scene.get("env_activation", "softplus")) * last.unsqueeze(-1)
    off_hat = ((off_m + env).repeat(2, 1) * R).view(-1, n2, 3).mean(-2)
    reflect = (emo_m.repeat(2, 1) * R).view(-1, n2, 3).mean(-2)
    if pdra_mode:
        um = umask.repeat(2)
        emo_hat = torch.where(um[:, None], emission.repeat(2, 1) + reflect.detach(), reflect)
    else:
        emo_hat = emission.repeat(2, 1) + reflect
    return dict(off=off, emo=emo, off_hat=off_hat, emo_hat=emo_hat), inter


def esrnerf_forward_training(scene: Dict, params: Dict, rays_o, rays_d, viewdirs, em_modes, uncert_masks, s_val: float,
                             normal_eps: float, emit_eps: float, pdra_mode: bool = False, draws=None):
    """esrnerf.py:681-851.  Returns (outputs, intermediates)."""
    draws = draws or Draws()
    N = rays_o.shape[0]
    ray_pts, ray_id, step_id, aux = _march_near(scene, rays_o, rays_d, scene["near"])
    inter = dict(aux)
    inter["m0"] = int(ray_pts.shape[0])
    keep = P.mask_cache(scene, ray_pts)
    ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
    inter.update(m1_ray=ray_id, m1_step=step_id)
    sdf, exp_grad = sdf_expgrad(params["sdf"], ray_pts, scene["xyz_min"], scene["xyz_max"], True)
    alpha = _alpha(scene, params, viewdirs, ray_id, ray_pts, sdf, s_val)
    inter.update(m1_sdf=sdf, m1_alpha=alpha)
    k0 = alpha > scene["fast_thres"]
    alpha, ray_id, step_id, ray_pts, exp_grad, sdf = (t[k0] for t in (alpha, ray_id, step_id, ray_pts, exp_grad, sdf))
    weights, last = P._A2W.apply(alpha, ray_id, N)
    k1 = weights > scene["fast_thres"]
    weights, ray_id, step_id, ray_pts, exp_grad, sdf = (t[k1] for t in (weights, ray_id, step_id, ray_pts, exp_grad, sdf))
    inter.update(m3_ray=ray_id, m3_step=step_id, m3_weights=weights, m3_sdf=sdf, m3_pts=ray_pts)

    on = em_modes[ray_id] == 1
    feat, _, fnormal = _taps(scene, params, ray_pts)
    xyz_emb = _pos_emb(scene, ray_pts)
    v = viewdirs[ray_id]
    rgb_feat = torch.cat([xyz_emb, v, v.sin(), v.cos(), sdf[:, None], feat, fnormal], -1)
    emo_c = _sample(params["emo_color"], scene, ray_pts)
    lin_emo = P.mlp(torch.cat([emo_c, rgb_feat], -1), params["emo_rgbnet"], F.softplus)
    lin_off = P.mlp(torch.cat([_sample(params["off_color"], scene, ray_pts), rgb_feat], -1), params["off_rgbnet"], F.softplus)
    lin = torch.where(on[:, None], lin_emo, torch.zeros_like(lin_emo)) + lin_off        # esrnerf.py:751-757 (no detach)
    rgb = P.tonemap(params, lin)
    brdf_feat = torch.cat([xyz_emb, sdf[:, None], feat, fnormal], -1)
    base, rough, metal = _brdf_split(P.mlp(torch.cat([_sample(params["brdf"], scene, ray_pts), brdf_feat], -1),
                                           params["brdfnet"], torch.sigmoid))
    emit = P.mlp(torch.cat([emo_c, brdf_feat], -1), params["emitnet"], F.softplus)
    inter.update(m3_lin=lin, m3_rgb=rgb)

    w_ = weights[:, None]

    def comp(x):
        return torch.zeros(N, 3).index_add(0, ray_id, w_ * x)

    rgb_m, lin_m, emit_m = comp(rgb), comp(lin), comp(emit)
    normal = F.normalize(exp_grad.detach(), dim=-1)
    m3 = ray_pts.shape[0]
    idx = draws.choice(m3, min(scene["num_ltspts"], m3))
    dir_noise = scatter_noise(scene, draws, idx.shape[0], scene["num_2ndrays"] + 1, 3)
    lts, lts_inter = light_transport_segment(scene, params, ray_pts[idx], viewdirs[ray_id][idx], normal[idx], sdf[idx],
                                             base[idx], rough[idx], metal[idx], emit[idx], uncert_masks[ray_id][idx],
                                             s_val, pdra_mode, dir_noise)
    inter.update(lts_idx=idx, lts=lts_inter)
    # eps branches (esrnerf.py:807-830)
    _, exp_grad_eps = sdf_expgrad(params["sdf"], ray_pts + draws.randn(m3, 3) * normal_eps, scene["xyz_min"],
                                  scene["xyz_max"], True)
    pts_e = ray_pts + draws.randn(m3, 3) * emit_eps
    xyz_emb_e = _pos_emb(scene, pts_e)
    sdf_e = _sample(params["sdf"], scene, pts_e)[:, 0]
    feat_e, _, fnormal_e = _taps(scene, params, pts_e)
    brdf_feat_e = torch.cat([xyz_emb_e, sdf_e[:, None], feat_e, fnormal_e], -1)
    emit_e = P.mlp(torch.cat([_sample(params["emo_color"], scene, pts_e), brdf_feat_e], -1), params["emitnet"], F.softplus)
    brdf_e = P.mlp(torch.cat([_sample(params["brdf"], scene, pts_e), brdf_feat_e], -1), params["brdfnet"], torch.sigmoid)
    out = {
        "etc/alphainv_cum": last, "etc/white_bg": last[..., None], "srgb/rgb": rgb_m, "lin/rgb": lin_m,
        "lin/pbr/off": lts["off"], "lin/pbr/off_hat": lts["off_hat"], "lin/pbr/emo": lts["emo"],
        "lin/pbr/emo_hat": lts["emo_hat"], "etc/emit_uncert": emit_m[uncert_masks], "etc/emit_cert": emit_m[~uncert_masks],
        "etc/normal": exp_grad, "etc/normal_eps": exp_grad_eps, "etc/emit": emit, "etc/emit_eps": emit_e,
        "etc/brdf": torch.cat([base, rough, metal], -1), "etc/brdf_eps": brdf_e,
    }
    return out, inter


def _primary_eval_stream(scene, params, rays_o, rays_d, s_val, manual: bool, viewdirs=None):
    """shared head of forward_evaluate / eval_emit / eval_esp (esrnerf.py:1012-1089, 1306-1339, 1367-1400)"""
    N = rays_o.shape[0]
    ray_pts, ray_id, step_id, _ = _march_near(scene, rays_o, rays_d, scene["near"])
    keep = P.mask_cache(scene, ray_pts)
    ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
    exp_grad = None
    if manual:
        sdf, exp_grad = sdf_expgrad(params["sdf"], ray_pts, scene["xyz_min"], scene["xyz_max"], False)
    else:
        sdf = _sample(params["sdf"], scene, ray_pts)[:, 0]
    alpha = _alpha(scene, params, viewdirs, ray_id, ray_pts, sdf, s_val)
    k0 = alpha > scene["fast_thres"]
    weights, T, last, _, _ = P.H.alpha2weight(alpha[k0], ray_id[k0], N)
    k1 = weights > scene["fast_thres"]

    def pick(t):
        return t[k0][k1]

    return dict(N=N, pts=pick(ray_pts), ray=pick(ray_id), step=pick(step_id), sdf=pick(sdf), w=weights[k1], last=last,
                exp_grad=None if exp_grad is None else pick(exp_grad))


@torch.no_grad()
def esrnerf_eval_esp(scene, params, rays_o, rays_d, viewdirs, s_val):
    """esrnerf.py:1360-1407: expected surface point = sum_ray w * ray_pts"""
    st = _primary_eval_stream(scene, params, rays_o, rays_d, s_val, False, viewdirs)
    return torch.zeros(st["N"], 3).index_add(0, st["ray"], st["w"][:, None] * st["pts"]), st


@torch.no_grad()
def esrnerf_eval_emit(scene, params, rays_o, rays_d, viewdirs, s_val):
    """esrnerf.py:1299-1358: composite of the emission net (emit_color aliases emo_color outside finetune, Q13)"""
    st = _primary_eval_stream(scene, params, rays_o, rays_d, s_val, False, viewdirs)
    pts = st["pts"]
    feat, _, fnormal = _taps(scene, params, pts)
    brdf_feat = torch.cat([_pos_emb(scene, pts), st["sdf"][:, None], feat, fnormal], -1)
    emit = P.mlp(torch.cat([_sample(params["emo_color"], scene, pts), brdf_feat], -1), params["emitnet"], F.softplus)
    return torch.zeros(st["N"], 3).index_add(0, st["ray"], st["w"][:, None] * emit), st


def _lts_eval(scene, params, pts, viewdirs, normal, base, rough, metal, emit, s_val, dir_noise):
    """esrnerf.py:854-1001 (one chunk)"""
    n2, Pn = scene["num_2ndrays"], pts.shape[0]
    dirs = scatter_dirs(scene, normal, scene["num_2ndrays"], dir_noise)

    def ex(t, c):
        return t.view(-1, 1, c).expand(Pn, n2, c).flatten(0, 1)

    d_flat = dirs.flatten(0, 1)
    R = disney_reflection(ex(base, 3), ex(rough, 1), ex(metal, 1), ex(normal, 3), d_flat, -ex(viewdirs, 3))
    off_m, emo_m, last, _ = _secondary(scene, params, ex(pts, 3), d_flat, s_val)
    env = sg_envmap(params, d_flat, scene.get("env_activation", "softplus")) * last.unsqueeze(-1)
    out = {"lin/env_dir": (env * R).view(-1, n2, 3).mean(-2), "lin/env_indir": (off_m * R).view(-1, n2, 3).mean(-2)}
    out["lin/env_effects"] = out["lin/env_dir"] + out["lin/env_indir"]
    out["lin/emit_(in)dir"] = (emo_m * R).view(-1, n2, 3).mean(-2)
    out["lin/emit_effects"] = emit + out["lin/emit_(in)dir"]
    return out


@torch.no_grad()
def esrnerf_forward_evaluate(scene, params, rays_o, rays_d, viewdirs, em_modes, pos_rt, s_val, render_pbr: bool,
                             chunk_sz: int, draws=None):
    """esrnerf.py:853-1297 (general branch; the degenerate `alpha.dim() != 1` branch is not restated)."""
    draws = draws or Draws()
    st = _primary_eval_stream(scene, params, rays_o, rays_d, s_val, True, viewdirs)
    N, pts, ray_id, step_id, sdf, weights, last = (st[k] for k in ("N", "pts", "ray", "step", "sdf", "w", "last"))
    _, g, _ = P.sdf_feature_taps(scene, params["sdf"], pts, [1.0], fd_eps=1e-12)      # esrnerf.py:1598-1605
    grad = torch.stack([g[:, 2], g[:, 1], g[:, 0]], -1)
    feat, _, fnormal = _taps(scene, params, pts)
    xyz_emb = _pos_emb(scene, pts)
    v = viewdirs[ray_id]
    rgb_feat = torch.cat([xyz_emb, v, v.sin(), v.cos(), sdf[:, None], feat, fnormal], -1)
    emo_c = _sample(params["emo_color"], scene, pts)
    lin_off = P.mlp(torch.cat([_sample(params["off_color"], scene, pts), rgb_feat], -1), params["off_rgbnet"], F.softplus)
    lin_emo = P.mlp(torch.cat([emo_c, rgb_feat], -1), params["emo_rgbnet"], F.softplus)
    lin_on = lin_off + lin_emo
    brdf_feat = torch.cat([xyz_emb, sdf[:, None], feat, fnormal], -1)
    base, rough, metal = _brdf_split(P.mlp(torch.cat([_sample(params["brdf"], scene, pts), brdf_feat], -1),
                                           params["brdfnet"], torch.sigmoid))
    emit = P.mlp(torch.cat([emo_c, brdf_feat], -1), params["emitnet"], F.softplus)
    w_ = weights[:, None]

    def comp(x):
        return torch.zeros(N, x.shape[1]).index_add(0, ray_id, w_ * x)

    normal = F.normalize(grad, dim=-1) @ pos_rt
    normal = (normal * torch.tensor([1.0, -1.0, -1.0]) + 1.0) / 2.0
    depth = torch.zeros(N).index_add(0, ray_id, weights * step_id * scene["stepdist"])
    out = {
        "etc/depth": depth, "etc/disp": 1 / (depth + last * scene["far"]), "etc/normal": comp(normal),
        "etc/white_bg": last.unsqueeze(-1),
        "srgb/off_rgb": comp(P.tonemap(params, lin_off)), "lin/off_rgb": comp(lin_off),
        "srgb/on_rgb": comp(P.tonemap(params, lin_on)), "lin/on_rgb": comp(lin_on),
        "srgb/emo_rgb": comp(P.tonemap(params, lin_emo)), "lin/emo_rgb": comp(lin_emo),
        "lin/emit": comp(emit), "lin/basecolor": comp(base), "lin/roughness": comp(rough)[:, 0],
        "lin/metallic": comp(metal)[:, 0],
    }
    sel = "off" if int(em_modes) == 0 else "on"
    out["srgb/rgb"], out["lin/rgb"] = out[f"srgb/{sel}_rgb"], out[f"lin/{sel}_rgb"]
    if render_pbr:
        vdir = viewdirs[ray_id]
        nrm = F.normalize(st["exp_grad"], dim=-1)
        parts = {}
        for idx in torch.arange(pts.shape[0]).split(chunk_sz):
            noise = scatter_noise(scene, draws, idx.shape[0], scene["num_2ndrays"], 3)
            ret = _lts_eval(scene, params, pts[idx], vdir[idx], nrm[idx], base[idx], rough[idx], metal[idx], emit[idx],
                            s_val, noise)
            for k, val in ret.items():
                parts.setdefault(k, []).append(val)
        for k, val in parts.items():
            out[k] = comp(torch.cat(val, 0))
    return out, dict(m3_ray=ray_id, m3_step=step_id, m3_weights=weights)


def rgb_to_hsv(rgb: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """pbr/functions.py:214-236"""
    mx, arg = rgb.max(-1)
    mn = rgb.min(-1).values
    delta = mx - mn
    sat = delta / (mx + eps)
    delta = torch.where(delta == 0, torch.ones_like(delta), delta)
    rc, gc, bc = torch.unbind(mx.unsqueeze(-1) - rgb, dim=-1)
    h = torch.stack((bc - gc, (rc - bc) + 2.0 * delta, (gc - rc) + 4.0 * delta), dim=-1) / delta.unsqueeze(-1)
    h = torch.gather(h, -1, arg.unsqueeze(-1)).squeeze(-1)
    return torch.stack(((h / 6.0) % 1.0, sat, mx), dim=-1)


def hsv_to_rgb(hsv: torch.Tensor) -> torch.Tensor:
    """pbr/functions.py:239-255"""
    h, sat, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    p = v * (1.0 - sat)
    q = v * (1.0 - f * sat)
    t = v * (1.0 - (1.0 - f) * sat)
    hi = hi.long()
    table = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-1)
    return torch.gather(table, -1, torch.stack([hi, hi + 6, hi + 12], dim=-1))


def edit_emission(emit, em_modes, em_intensities, em_colors):
    """esrnerf.py:407-417: emission-source editing modes (utils2 LightDict: 0 off, 1 on, 2 intensity, 3 colour, 4 both)"""
    emit = emit.clone()
    i_mask = (em_modes == 2) | (em_modes == 4)
    c_mask = (em_modes == 3) | (em_modes == 4)
    emit[em_modes == 0] = 0
    emit[i_mask] = emit[i_mask] * em_intensities[i_mask][..., None]
    hsv = rgb_to_hsv(emit[c_mask])
    hsv[..., :-1] = em_colors[c_mask]
    emit[c_mask] = hsv_to_rgb(hsv)
    return emit


def esrnerf_forward_finetune(scene, params, rays_o, rays_d, viewdirs, em_modes, em_intensities, em_colors, s_val,
                             draws=None):
    """esrnerf.py:241-484 — `@torch.no_grad` with autograd enabled only around emo_rgbnet(emo_color(x)) at the LTS
    points (Q14): 'lin/pbr/emo' carries gradient to emo_rgbnet / emo_color, 'lin/pbr/emo_hat' is a constant target.
    params["emit_color"] is the frozen copy of emo_color made by train(finetune=True) (esrnerf.py:224-235)."""
    draws = draws or Draws()
    n2 = scene["num_2ndrays"]
    with torch.no_grad():
        st = _primary_eval_stream(scene, params, rays_o, rays_d, s_val, False, viewdirs)
        m3 = st["pts"].shape[0]
        idx = draws.choice(m3, min(scene["num_ltspts"], m3))
        pts, ray = st["pts"][idx], st["ray"][idx]
        vdir, modes, inten, cols = viewdirs[ray], em_modes[ray], em_intensities[ray], em_colors[ray]
        Pn = pts.shape[0]
        sdf, exp_grad = sdf_expgrad(params["sdf"], pts, scene["xyz_min"], scene["xyz_max"], True)
        sdf, normal = sdf.detach(), F.normalize(exp_grad.detach(), dim=-1)
        dirs = scatter_dirs(scene, normal, n2 + 1, scatter_noise(scene, draws, Pn, n2 + 1, 3))
        v_rand = -dirs[:, -1]
        dirs = dirs[:, :-1]
        feat, _, fnormal = _taps(scene, params, pts)
        xyz_emb = _pos_emb(scene, pts)
        v2 = torch.cat([vdir, v_rand], 0)
        rgb_feat = torch.cat([xyz_emb.repeat(2, 1), v2, v2.sin(), v2.cos(), sdf[:, None].repeat(2, 1), feat.repeat(2, 1),
                              fnormal.repeat(2, 1)], -1)
        with torch.enable_grad():
            emo = P.mlp(torch.cat([_sample(params["emo_color"], scene, pts).repeat(2, 1), rgb_feat], -1),
                        params["emo_rgbnet"], F.softplus)
        brdf_feat = torch.cat([xyz_emb, sdf[:, None], feat, fnormal], -1)
        base, rough, metal = _brdf_split(P.mlp(torch.cat([_sample(params["brdf"], scene, pts), brdf_feat], -1),
                                               params["brdfnet"], torch.sigmoid))
        emit = P.mlp(torch.cat([_sample(params["emit_color"], scene, pts), brdf_feat], -1), params["emitnet"], F.softplus)

        def ex(t, c):
            return t.view(-1, 1, c).expand(Pn, n2, c).flatten(0, 1)

        d_flat = dirs.flatten(0, 1)
        R = disney_reflection(ex(base, 3).repeat(2, 1), ex(rough, 1).repeat(2, 1), ex(metal, 1).repeat(2, 1),
                              ex(normal, 3).repeat(2, 1), d_flat.repeat(2, 1), torch.cat([-ex(vdir, 3), -ex(v_rand, 3)], 0))
        _, emo_m, _, _ = _secondary(scene, params, ex(pts, 3), d_flat, s_val)
        emit = edit_emission(emit, modes, inten, cols)
        reflect = (emo_m.repeat(2, 1) * R).view(-1, n2, 3).mean(-2)
        emo_hat = emit.repeat(2, 1) + reflect
    return {"lin/pbr/emo": emo, "lin/pbr/emo_hat": emo_hat}


def params_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict:
    """state_dict keys of the reference ESRNeRF -> the dict the port consumes."""
    p = P.params_from_state_dict(sd)

    def net(prefix, idx):
        return [(sd[f"{prefix}.{i}.weight"].float(), sd[f"{prefix}.{i}.bias"].float()) for i in idx]

    p["brdf"] = sd["brdf.grid"].float().contiguous()
    if "emit_color.grid" in sd:
        p["emit_color"] = sd["emit_color.grid"].float().contiguous()
    p["emitnet"] = net("emitnet.brdfnet", ["0", "2.0", "3.0", "4"])
    p["brdfnet"] = net("brdfnet.brdfnet", ["0", "2.0", "3.0", "4"])
    for k in ("envmap.mus", "envmap.lambdas", "envmap.lobes"):
        p[k] = sd[k].float()
    return p


_ = math

"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (torch ops + the C restatement of the three native ops) of the
reference's fine-stage render functions, written so that it can travel to the GPU
box where ``/root/reference`` does not exist:

* ``voxurff_forward_training``  — app/fine/model/voxurff.py:177-278
* ``voxurff_forward_evaluate``  — app/fine/model/voxurff.py:280-461

Parity pin: ``tests/test_oracle_cpu.py::test_port_matches_reference`` runs this port
against the reference's OWN ``VoxurfF`` class (imported through
``oracle/ref_harness.py``) on identical weights and rays whenever ``/root/reference``
is present, and against the committed golden vectors in ``tests/golden/`` (produced by
the reference's own code, ``oracle/make_golden.py``) everywhere else.

Besides the output dict the functions return every intermediate the CUDA path is
checked against (sample streams, masks, alpha, weights).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_harness as H


def grid_sample_world(grid: torch.Tensor, xyz: torch.Tensor, xyz_min, xyz_max) -> torch.Tensor:
    """module.py:24-35 / voxurff.py:656-668: world points -> trilinear taps of a [1,C,X,Y,Z] grid."""
    pts = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((pts - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1
    out = F.grid_sample(grid, ind_norm, mode="bilinear", align_corners=True)
    return out.reshape(grid.shape[1], -1).T  # [M, C]


def mask_cache(scene: Dict, xyz: torch.Tensor) -> torch.Tensor:
    """module.py:104-114 (density is the max-pooled grid built by module.py:94-99)."""
    d = grid_sample_world(scene["mask_density_pooled"], xyz, scene["mask_xyz_min"], scene["mask_xyz_max"])[:, 0]
    alpha = 1 - torch.exp(-F.softplus(d + scene["act_shift"]))
    return alpha >= scene["mask_thres"]


def pooled_mask_density(density: torch.Tensor, ks: int = 3) -> torch.Tensor:
    """module.py:94-99"""
    return F.max_pool3d(density, kernel_size=ks, padding=ks // 2, stride=1)


def neus_alpha_interp(ray_id: torch.Tensor, sdf: torch.Tensor, s_val: float) -> torch.Tensor:
    """functions.py:72-105"""
    if sdf.numel() == 0:
        return sdf
    same = ray_id[:-1] == ray_id[1:]
    mid = (sdf[:-1] + sdf[1:]) * 0.5
    nxt = torch.cat([torch.where(same, mid, sdf[:-1]), sdf[-1:]])
    prv = torch.cat([sdf[:1], torch.where(same, mid, sdf[1:])])
    pc = torch.sigmoid(prv * s_val)
    nc = torch.sigmoid(nxt * s_val)
    return ((F.relu(pc - nc) + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0)


def neus_alpha_grad(viewdirs: torch.Tensor, ray_id: torch.Tensor, dist, sdf: torch.Tensor, grad: torch.Tensor,
                    s_val: float) -> torch.Tensor:
    """functions.py:45-69 (`neus_alpha: grad`): section-point SDFs estimated from the SDF gradient along the view
    direction, sdf -+ 0.5 * dist * (v . grad sdf), instead of from the neighbouring samples."""
    if sdf.numel() == 0:
        return sdf
    iter_cos = (viewdirs[ray_id] * grad).sum(-1, keepdim=True) * torch.as_tensor(dist).reshape(-1, 1) * 0.5
    s = sdf.unsqueeze(-1)
    pc = torch.sigmoid((s - iter_cos) * s_val)
    nc = torch.sigmoid((s + iter_cos) * s_val)
    return ((F.relu(pc - nc) + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0).squeeze(-1)


def sdf_fd_gradient(scene: Dict, sdf_grid: torch.Tensor, xyz: torch.Tensor, fd_eps: float = 0.0) -> torch.Tensor:
    """sample_sdf_grad's gradient (voxurff.py:670-676): the displace = 1 finite differences, (z,y,x) -> (x,y,z)"""
    _, g, _ = sdf_feature_taps(scene, sdf_grid, xyz, [1.0], fd_eps)
    return torch.stack([g[:, 2], g[:, 1], g[:, 0]], -1)


def neus_alpha(scene: Dict, sdf_grid: torch.Tensor, viewdirs, ray_id, ray_pts, sdf, s_val: float, fd_eps: float = 0.0):
    """the model's `neus_alpha_from_sdf_scatter` (voxurff.py:151-154): scene["neus_alpha"] = "interp" (default, every
    shipped config) or "grad"."""
    if scene.get("neus_alpha", "interp") == "grad":
        grad = sdf_fd_gradient(scene, sdf_grid, ray_pts, fd_eps)
        dist = torch.tensor(scene["stepdist"], dtype=torch.float32)      # stepsize * voxel_size (voxurff.py:195)
        return neus_alpha_grad(viewdirs, ray_id, dist, sdf, grad, s_val)
    return neus_alpha_interp(ray_id, sdf, s_val)


class _A2W(torch.autograd.Function):
    """module.py:117-143 over the C restatement of kernel.cu:576-707."""

    @staticmethod
    def forward(ctx, alpha, ray_id, n):
        w, T, last, i_s, i_e = H.alpha2weight(alpha, ray_id, n)
        ctx.save_for_backward(alpha.detach(), w, T, last, i_s, i_e)
        ctx.n = n
        return w, last

    @staticmethod
    def backward(ctx, gw, gl):
        alpha, w, T, last, i_s, i_e = ctx.saved_tensors
        return H.alpha2weight_backward(alpha, w, T, last, i_s, i_e, ctx.n, gw.contiguous(), gl.contiguous()), None, None


def sdf_feature_taps(scene: Dict, sdf_grid: torch.Tensor, xyz: torch.Tensor, displace, fd_eps: float = 0.0):
    """voxurff.py:678-721: 6 axis taps (z-,z+,y-,y+,x-,x+) x K displacements, FD gradient, normals.
    fd_eps = 1e-12 gives the ESRNeRF variant (esrnerf.py:1560)."""
    M = xyz.shape[0]
    K = len(displace)
    size_zyx = torch.tensor([sdf_grid.shape[4], sdf_grid.shape[3], sdf_grid.shape[2]], dtype=torch.float32)
    ind_norm = ((xyz - scene["xyz_min"]) / (scene["xyz_max"] - scene["xyz_min"])).flip((-1,)) * 2 - 1
    ind = ((ind_norm + 1) / 2) * (size_zyx - 1)  # [M,3] (z,y,x)
    axes = torch.tensor([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], dtype=torch.float32)
    disp = torch.as_tensor(displace, dtype=torch.float32)
    offs = (axes[:, None, :] * disp[None, :, None]).reshape(-1, 3)  # [6K,3] tap-major
    taps = ind[:, None, :] + offs[None]  # [M,6K,3]
    taps = torch.minimum(torch.maximum(taps, torch.zeros(3)), size_zyx - 1)
    taps_norm = (taps / (size_zyx - 1)) * 2 - 1
    feat = F.grid_sample(sdf_grid, taps_norm.reshape(1, 1, 1, -1, 3), mode="bilinear", align_corners=True)
    feat = feat.reshape(M, 6, K)
    taps = taps.reshape(M, 6, K, 3)
    diff = (taps[:, 1::2] - taps[:, 0::2]).max(dim=-1).values  # [M,3,K]
    if fd_eps:
        diff = diff + fd_eps
    grad = (feat[:, 1::2] - feat[:, 0::2]) / diff / scene["voxel_size"]
    normal = F.normalize(grad, dim=1)
    return feat.reshape(M, 6 * K), grad.reshape(M, 3 * K), normal.reshape(M, 3 * K)


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


class _RoundBf16(torch.autograd.Function):
    """value and cotangent both rounded to bf16 (the storage points of the tensor-core MLP kernels)"""

    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundBf16Bwd(torch.autograd.Function):
    """identity forward, bf16-rounded cotangent"""

    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


MLP_PRECISION = "fp32"   # "bf16": emulate the rounding points of the bf16 tensor-core kernels (fp32 accumulate)


def mlp(x, layers, out_act):
    """pbr/module.py:6-39: Linear/ReLU stack; layers = [(W,b), ...].

    With MLP_PRECISION == "bf16" the same stack is evaluated with inputs, weights, hidden activations and
    their cotangents rounded to bf16 where esr_mlp_fwd / esr_mlp_bwd store them (fp32 accumulation, fp32
    bias / output) — the numeric contract of the product's bf16 mode, used to separate kernel bugs from
    the inherent bf16 error."""
    if MLP_PRECISION == "fp32":
        for i, (w, b) in enumerate(layers):
            x = F.linear(x, w, b)
            if i + 1 < len(layers):
                x = F.relu(x)
        return out_act(x)
    x = _RoundBf16.apply(x)
    for i, (w, b) in enumerate(layers):
        x = F.linear(x, _RoundBf16.apply(w), b)
        if i + 1 < len(layers):
            x = _RoundBf16.apply(F.relu(x))
        else:
            x = _RoundBf16Bwd.apply(x)
    return out_act(x)


def tonemap(params, lin):
    """voxurff.py:783-788"""
    freq = torch.tensor([2.0 ** i for i in range(5)])
    emb = (lin.unsqueeze(-1) * freq).flatten(-2)
    return mlp(torch.cat([lin, emb.sin(), emb.cos()], -1), params["tonemapper"], torch.sigmoid)


def _march(scene, rays_o, rays_d):
    """voxurff.py:623-654 over the C restatement of kernel.cu:196-242."""
    ray_pts, mask_out, ray_id, step_id, n_steps, t_min, t_max = H.sample_pts_on_rays(
        rays_o.contiguous(), rays_d.contiguous(), scene["xyz_min"], scene["xyz_max"], scene["near"], 1e9,
        scene["stepdist"])
    inb = ~mask_out
    return ray_pts[inb], ray_id[inb], step_id[inb], dict(n_steps=n_steps, t_min=t_min, t_max=t_max,
                                                        m_raw=int(ray_pts.shape[0]))


def _features(scene, params, ray_pts, ray_id, sdf, viewdirs):
    """voxurff.py:219-241"""
    feat, _, normal = sdf_feature_taps(scene, params["sdf"], ray_pts, scene["grad_feat"])
    u = (ray_pts - scene["xyz_min"]) / (scene["xyz_max"] - scene["xyz_min"])
    freq = torch.tensor([2.0 ** i for i in range(5)])
    emb = (u.unsqueeze(-1) * freq).flatten(-2)
    vemb = (viewdirs.unsqueeze(-1) * torch.tensor([1.0])).flatten(-2)
    rgb_feat = torch.cat([u, emb.sin(), emb.cos(), vemb[ray_id], vemb.sin()[ray_id], vemb.cos()[ray_id],
                          sdf[:, None], feat, normal], -1)
    return rgb_feat, feat, normal


def voxurff_forward_training(scene: Dict, params: Dict, rays_o, rays_d, viewdirs, em_modes, s_val: float):
    N = rays_o.shape[0]
    ray_pts, ray_id, step_id, aux = _march(scene, rays_o, rays_d)
    inter = dict(aux)
    inter["m0"] = int(ray_pts.shape[0])
    keep = mask_cache(scene, ray_pts)
    ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
    inter.update(m1_ray=ray_id, m1_step=step_id)

    sdf = grid_sample_world(params["sdf"], ray_pts, scene["xyz_min"], scene["xyz_max"])[:, 0]
    alpha = neus_alpha(scene, params["sdf"], viewdirs, ray_id, ray_pts, sdf, s_val)
    inter.update(m1_sdf=sdf, m1_alpha=alpha)

    k0 = alpha > scene["fast_thres"]
    alpha, ray_id, step_id, ray_pts, sdf = alpha[k0], ray_id[k0], step_id[k0], ray_pts[k0], sdf[k0]
    weights, last = _A2W.apply(alpha, ray_id, N)
    inter.update(m2_ray=ray_id, m2_step=step_id, m2_weights=weights)
    k1 = weights > scene["fast_thres"]
    weights, ray_id, step_id, ray_pts, sdf = weights[k1], ray_id[k1], step_id[k1], ray_pts[k1], sdf[k1]
    inter.update(m3_ray=ray_id, m3_step=step_id, m3_weights=weights, m3_sdf=sdf)

    on = em_modes[ray_id] == 1
    rgb_feat, feat, normal = _features(scene, params, ray_pts, ray_id, sdf, viewdirs)
    off_c = grid_sample_world(params["off_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    emo_c = grid_sample_world(params["emo_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    inter.update(m3_feat=rgb_feat, m3_off_color=off_c, m3_emo_color=emo_c)
    lin_off = mlp(torch.cat([off_c, rgb_feat], -1), params["off_rgbnet"], F.softplus)
    lin_emo = mlp(torch.cat([emo_c, rgb_feat], -1), params["emo_rgbnet"], F.softplus)
    # voxurff.py:243-254: emission-on rays add emo and see off only through a stop-gradient
    lin = torch.where(on[:, None], lin_emo + lin_off.detach(), lin_off)
    rgb = tonemap(params, lin)
    inter.update(m3_lin=lin, m3_rgb=rgb)
    w_ = weights[:, None]
    out = {
        "etc/alphainv_cum": last,
        "etc/white_bg": last[..., None],
        "srgb/rgb": torch.zeros(N, 3).index_add(0, ray_id, w_ * rgb),
        "lin/rgb": torch.zeros(N, 3).index_add(0, ray_id, w_ * lin),
    }
    return out, inter


def voxurff_forward_evaluate(scene: Dict, params: Dict, rays_o, rays_d, viewdirs, em_modes, pos_rt, s_val: float):
    """voxurff.py:280-461 (general branch; the degenerate `alpha.dim()!=1` branch is not restated)."""
    N = rays_o.shape[0]
    ray_pts, ray_id, step_id, aux = _march(scene, rays_o, rays_d)
    keep = mask_cache(scene, ray_pts)
    ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
    sdf = grid_sample_world(params["sdf"], ray_pts, scene["xyz_min"], scene["xyz_max"])[:, 0]
    grad = sdf_fd_gradient(scene, params["sdf"], ray_pts)  # voxurff.py:670-676
    alpha = neus_alpha(scene, params["sdf"], viewdirs, ray_id, ray_pts, sdf, s_val)
    k0 = alpha > scene["fast_thres"]
    alpha, ray_id, step_id, ray_pts, sdf, grad = (t[k0] for t in (alpha, ray_id, step_id, ray_pts, sdf, grad))
    weights, T, last, _, _ = H.alpha2weight(alpha, ray_id, N)
    k1 = weights > scene["fast_thres"]
    weights, ray_id, step_id, ray_pts, sdf, grad = (t[k1] for t in (weights, ray_id, step_id, ray_pts, sdf, grad))
    rgb_feat, _, _ = _features(scene, params, ray_pts, ray_id, sdf, viewdirs)
    off_c = grid_sample_world(params["off_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    emo_c = grid_sample_world(params["emo_color"], ray_pts, scene["xyz_min"], scene["xyz_max"])
    lin_off = mlp(torch.cat([off_c, rgb_feat], -1), params["off_rgbnet"], F.softplus)
    lin_emo = mlp(torch.cat([emo_c, rgb_feat], -1), params["emo_rgbnet"], F.softplus)
    lin_on = lin_off + lin_emo
    w_ = weights[:, None]

    def comp(x):
        return torch.zeros(N, x.shape[1]).index_add(0, ray_id, w_ * x)

    normal = F.normalize(grad, dim=-1) @ pos_rt
    normal = (normal * torch.tensor([1.0, -1.0, -1.0]) + 1.0) / 2.0
    depth = torch.zeros(N).index_add(0, ray_id, weights * step_id * scene["stepdist"])
    disp = 1 / (depth + last * scene["far"])
    out = {
        "etc/depth": depth, "etc/disp": disp, "etc/normal": comp(normal), "etc/white_bg": last.unsqueeze(-1),
        "srgb/off_rgb": comp(tonemap(params, lin_off)), "lin/off_rgb": comp(lin_off),
        "srgb/on_rgb": comp(tonemap(params, lin_on)), "lin/on_rgb": comp(lin_on),
        "srgb/emo_rgb": comp(tonemap(params, lin_emo)), "lin/emo_rgb": comp(lin_emo),
    }
    sel = "off" if int(em_modes) == 0 else "on"
    out["srgb/rgb"] = out[f"srgb/{sel}_rgb"]
    out["lin/rgb"] = out[f"lin/{sel}_rgb"]
    return out, dict(m3_ray=ray_id, m3_step=step_id, m3_weights=weights)


def params_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict:
    """state_dict keys of the reference VoxurfF -> the dict the port consumes."""
    def net(prefix, idx):
        return [(sd[f"{prefix}.{i}.weight"].float(), sd[f"{prefix}.{i}.bias"].float()) for i in idx]

    return {
        "sdf": sd["sdf.grid"].float().contiguous(),
        "off_color": sd["off_color.grid"].float().contiguous(),
        "emo_color": sd["emo_color.grid"].float().contiguous(),
        "off_rgbnet": net("off_rgbnet.linear", ["0", "2.0", "3.0", "4"]),
        "emo_rgbnet": net("emo_rgbnet.linear", ["0", "2.0", "3.0", "4"]),
        "tonemapper": net("tonemapper.srgb", ["0", "2"]),
    }

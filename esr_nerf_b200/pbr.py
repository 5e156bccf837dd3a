"""Reflectance math of the light-transport segment (app/utils/pbr/functions.py): hemisphere sampling and the
Disney-style BRDF.  Small tensors ([2 * num_ltspts * num_2ndrays, 3] = 51 200 rows at the shipped configuration):
fp32 elementwise math under autograd, minor next to the render chain of the secondary rays."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


@torch.no_grad()
def diffuse_scattering(normal: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """pbr/functions.py:10-18: uniform unit vectors (normalised Gaussian draw `noise` [..., n, 3]) mirrored into the
    hemisphere of `normal` [..., 3]."""
    ret = F.normalize(noise, dim=-1)
    below = (ret * normal.unsqueeze(-2)).sum(-1, keepdim=True) < 0
    return torch.where(below, -ret, ret)


def fibonacci_hemisphere(number: int) -> torch.Tensor:
    """pbr/functions.py:176-194 (`random=False, up=True`): the upper half of a 2n-point Fibonacci spiral, [n, 3] fp32
    on the host — golden-angle azimuths, cos(theta) equally spaced in (0, 1)."""
    n = 2 * number
    rn = torch.arange(number, n)
    phi = (math.pi * (3.0 - math.sqrt(5.0))) * ((rn + 1.0) % n)
    cos_theta = ((rn + 0.5) * (1.0 / number)) - 1.0
    sin_theta = torch.sqrt(1.0 - cos_theta * cos_theta)
    return torch.stack([torch.cos(phi) * sin_theta, torch.sin(phi) * sin_theta, cos_theta], dim=-1)


@torch.no_grad()
def diffuse_scattering_fib(normal: torch.Tensor, number: int) -> torch.Tensor:
    """pbr/functions.py:21-32 (`ray_sampling: fib`): the same deterministic spiral for every point, each direction
    mirrored into the hemisphere of that point's `normal` [..., 3] -> [..., number, 3].  No random draw."""
    ret = fibonacci_hemisphere(number).to(normal.device).expand(*normal.shape[:-1], number, 3)
    below = (ret * normal.unsqueeze(-2)).sum(-1, keepdim=True) < 0
    return torch.where(below, -ret, ret)


def disney_reflection(albedo, roughness, metallic, normal, win, wout):
    """pbr/functions.py:108-173: (diffuse + D F V) * cos(theta_i) * 2 pi with the spherical-Gaussian NDF, Schlick
    Fresnel and Schlick-GGX visibility."""
    eps = 1e-7
    h = F.normalize(win + wout, dim=-1)
    noh = (normal * h).sum(-1, keepdim=True).clamp(min=0)
    ooh = (wout * h).sum(-1, keepdim=True).clamp(min=0)
    ion = (win * normal).sum(-1, keepdim=True).clamp(min=0)
    oon = (wout * normal).sum(-1, keepdim=True).clamp(min=0)
    fd = (1 - metallic) * albedo / math.pi
    r2 = (roughness * roughness).clamp(min=eps)
    ndf = (1 / (r2 * math.pi)) * torch.exp((2 / r2) * (noh - 1))
    f0 = 0.04 * (1 - metallic) + albedo * metallic
    fresnel = f0 + (1.0 - f0) * ((1.0 - ooh) ** 5)
    k = ((1 + roughness) ** 2) / 8
    vis = (0.5 / (ion * (1 - k) + k).clamp(min=eps)) * (0.5 / (oon * (1 - k) + k).clamp(min=eps))
    return (fd + ndf * fresnel * vis) * ion * math.pi * 2


def rgb_to_hsv(rgb: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """pbr/functions.py:214-236"""
    mx, arg = rgb.max(-1)
    delta = mx - rgb.min(-1).values
    sat = delta / (mx + eps)
    delta = torch.where(delta == 0, torch.ones_like(delta), delta)
    rc, gc, bc = torch.unbind(mx.unsqueeze(-1) - rgb, dim=-1)
    hue = torch.stack((bc - gc, (rc - bc) + 2.0 * delta, (gc - rc) + 4.0 * delta), dim=-1) / delta.unsqueeze(-1)
    hue = torch.gather(hue, -1, arg.unsqueeze(-1)).squeeze(-1)
    return torch.stack(((hue / 6.0) % 1.0, sat, mx), dim=-1)


def hsv_to_rgb(hsv: torch.Tensor) -> torch.Tensor:
    """pbr/functions.py:239-255"""
    hue, sat, val = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    sector = torch.floor(hue * 6) % 6
    frac = ((hue * 6) % 6) - sector
    p = val * (1.0 - sat)
    q = val * (1.0 - frac * sat)
    t = val * (1.0 - (1.0 - frac) * sat)
    sector = sector.long()
    table = torch.stack((val, q, p, p, t, val, t, val, val, q, p, p, p, p, t, val, val, q), dim=-1)
    return torch.gather(table, -1, torch.stack([sector, sector + 6, sector + 12], dim=-1))


@torch.no_grad()
def edit_emission(emit, em_modes, em_intensities, em_colors):
    """esrnerf.py:407-417: per-point emission edits of the finetune stage (LightDict, utils2/utils.py: 0 off, 1 on,
    2 intensity change, 3 colour change, 4 both): zero / scale / replace hue and saturation."""
    emit = emit.clone()
    i_mask = (em_modes == 2) | (em_modes == 4)
    c_mask = (em_modes == 3) | (em_modes == 4)
    emit[em_modes == 0] = 0
    emit[i_mask] = emit[i_mask] * em_intensities[i_mask][..., None]
    hsv = rgb_to_hsv(emit[c_mask])
    hsv[..., :-1] = em_colors[c_mask]
    emit[c_mask] = hsv_to_rgb(hsv)
    return emit

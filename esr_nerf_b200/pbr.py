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

"""Deterministic synthetic scenes and rays for parity tests and bench.py.

No dataset travels with the repo, so every measurement uses: cubic AABB +-1.05; sdf.grid = |p| - 0.6 +
0.01 N(0,1); colour grids 0.1 N(0,1); MaskCache density +10 (dense) or +10 inside the shell
||p|-0.6| < 0.1 and -10 elsewhere (sparse); rays from camera centres on the r=4 sphere towards targets in
the r=0.9 ball with un-normalised directions (|d| ~ U(1,1.2)); em_modes ~ Bernoulli(0.5).
All tensors are generated on the CPU with fixed seeds and moved by the caller.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict

import torch
import torch.nn.functional as F

FINE_MODEL_CFG = dict(  # cfg/app/fine.yaml:13-30
    mask_ks=3, maskcache_thres=1e-3, fastcolor_thres=1e-4, stepsize=0.5, color_dim=6, rgbnet_width=192,
    rgbnet_depth=4, tonemap_width=192, tonemap_depth=2, posbase_pe=5, viewbase_pe=1, colorbase_pe=5,
    grad_feat=[0.5, 1.0, 1.5, 2.0], neus_alpha="interp")


def fine_cfg(device="cuda:0", **overrides):
    model = dict(FINE_MODEL_CFG)
    model.update(overrides)
    return SimpleNamespace(system=SimpleNamespace(device=device),
                           app=SimpleNamespace(model=SimpleNamespace(**model)))


def make_rays(n: int, seed: int = 1234, em_p: float = 0.5) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    origin = F.normalize(torch.randn(n, 3, generator=g), dim=-1) * 4.0
    target = F.normalize(torch.randn(n, 3, generator=g), dim=-1) * 0.9 * torch.rand(n, 1, generator=g) ** (1 / 3)
    dirs = F.normalize(target - origin, dim=-1)
    rays_d = dirs * (1.0 + 0.2 * torch.rand(n, 1, generator=g))
    em_modes = (torch.rand(n, generator=g) < em_p).long()
    return dict(rays_o=origin.contiguous(), rays_d=rays_d.contiguous(), viewdirs=F.normalize(rays_d, dim=-1),
                em_modes=em_modes, rgbs=torch.rand(n, 3, generator=g))


def mask_density(res: int, sparse: bool) -> torch.Tensor:
    """[1,1,res,res,res] alphamask-stage density on the AABB lattice."""
    if not sparse:
        return torch.full([1, 1, res, res, res], 10.0)
    ax = torch.linspace(-1.05, 1.05, res)
    gx, gy, gz = torch.meshgrid(ax, ax, ax, indexing="ij")
    r = (gx ** 2 + gy ** 2 + gz ** 2).sqrt()
    return torch.where((r - 0.6).abs() < 0.1, 10.0, -10.0)[None, None].contiguous()


def sphere_sdf(world_size, radius: float = 0.6, noise: float = 0.01, seed: int = 1) -> torch.Tensor:
    ax = [torch.linspace(-1.05, 1.05, int(w)) for w in world_size]
    gx, gy, gz = torch.meshgrid(*ax, indexing="ij")
    sdf = (gx ** 2 + gy ** 2 + gz ** 2).sqrt() - radius
    g = torch.Generator().manual_seed(seed)
    return (sdf + noise * torch.randn(sdf.shape, generator=g))[None, None].contiguous()


def color_grid(world_size, channels: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn([1, channels, *[int(w) for w in world_size]], generator=g)


def fill_fine_model(model, sdf_noise: float = 0.01) -> None:
    """Overwrite the grids of a (reference or esr_nerf_b200) VoxurfF in place with the synthetic scene."""
    ws = [int(w) for w in model.world_size]
    dev = model.sdf.grid.device
    with torch.no_grad():
        model.sdf.grid.copy_(sphere_sdf(ws, noise=sdf_noise).to(dev))
        model.off_color.grid.copy_(color_grid(ws, 6, 2).to(dev))
        model.emo_color.grid.copy_(color_grid(ws, 6, 3).to(dev))


LTS_MODEL_CFG = dict(  # cfg/app/lts.yaml:13-44 (pdra.yaml shares the model block)
    FINE_MODEL_CFG, brdfnet_width=128, brdfnet_depth=4, env_sg=48, env_activation="softplus", ray_sampling="random",
    num_2ndrays=256, num_ltspts=100, lts_near=1e-5)


def lts_cfg(device="cuda:0", **overrides):
    model = dict(LTS_MODEL_CFG)
    model.update(overrides)
    return SimpleNamespace(system=SimpleNamespace(device=device),
                           app=SimpleNamespace(model=SimpleNamespace(**model)))


def fill_esrnerf_model(model, sdf_noise: float = 0.01) -> None:
    """Overwrite the grids of a (reference or esr_nerf_b200) ESRNeRF in place with the synthetic scene."""
    fill_fine_model(model, sdf_noise)
    ws = [int(w) for w in model.world_size]
    with torch.no_grad():
        model.brdf.grid.copy_(color_grid(ws, 6, 4).to(model.brdf.grid.device))


def uncert_masks(n: int) -> torch.Tensor:
    """PDRA batch shape (utils2/utils.py:297-302): the first half of the batch is 'uncertain'"""
    m = torch.zeros(n, dtype=torch.bool)
    m[: n // 2] = True
    return m


def finetune_inputs(n: int, seed: int = 21) -> Dict[str, torch.Tensor]:
    """per-ray emission edits of the PDRA finetune stage (pdra.py:1048-1100): mode in LightDict 0..4, intensity scale,
    (hue, saturation) replacement"""
    g = torch.Generator().manual_seed(seed)
    return dict(em_modes=torch.randint(0, 5, (n,), generator=g), em_intensities=torch.rand(n, generator=g) * 2,
                em_colors=torch.rand(n, 2, generator=g))


def perturb_emit_color(model, seed: int = 9) -> None:
    """make the frozen emit_color copy differ from emo_color so tests can tell the two grids apart"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        noise = 0.05 * torch.randn(model.emit_color.grid.shape, generator=g)
        model.emit_color.grid.add_(noise.to(model.emit_color.grid.device))


COARSE_MODEL_CFG = dict(  # cfg/app/coarse.yaml:13-31
    mask_ks=3, maskcache_thres=1e-3, fastcolor_thres=1e-4, stepsize=0.5, num_voxels=96 ** 3, color_dim=12,
    rgbnet_width=128, rgbnet_depth=3, posbase_pe=5, viewbase_pe=1, smooth_ksize=5, smooth_sigma=0.8,
    neus_alpha="interp")


def coarse_cfg(device="cuda:0", **overrides):
    model = dict(COARSE_MODEL_CFG)
    model.update(overrides)
    return SimpleNamespace(system=SimpleNamespace(device=device),
                           app=SimpleNamespace(model=SimpleNamespace(**model)))


def fill_coarse_model(model, sdf_noise: float = 0.01) -> None:
    """Overwrite the grids of a (reference or esr_nerf_b200) VoxurfC in place with the synthetic scene."""
    ws = [int(w) for w in model.world_size]
    dev = model.sdf.grid.device
    with torch.no_grad():
        model.sdf.grid.copy_(sphere_sdf(ws, noise=sdf_noise).to(dev))
        model.off_color.grid.copy_(color_grid(ws, 12, 2).to(dev))
        model.emo_color.grid.copy_(color_grid(ws, 12, 3).to(dev))


def dvgo_cfg(device="cuda:0", num_voxels=1024000):
    """cfg/app/alphamask.yaml:13-17"""
    return SimpleNamespace(system=SimpleNamespace(device=device),
                           app=SimpleNamespace(model=SimpleNamespace(num_voxels=num_voxels, stepsize=0.5, alpha_init=1e-6)))


def fill_dvgo_model(model) -> None:
    """Synthetic alphamask-stage scene: density high inside the r = 0.6 ball (smooth falloff), random colour grids."""
    ws = [int(w) for w in model.density.shape[2:]]
    ax = [torch.linspace(-1.05, 1.05, w) for w in ws]
    gx, gy, gz = torch.meshgrid(*ax, indexing="ij")
    r = (gx ** 2 + gy ** 2 + gz ** 2).sqrt()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        model.density.copy_(((0.6 - r) * 40 + 12 + 0.5 * torch.randn(r.shape, generator=g))[None, None].to(model.density.device))
        model.off_color.copy_(torch.randn([1, 3, *ws], generator=g).to(model.density.device))
        model.emo_color.copy_(torch.randn([1, 3, *ws], generator=g).to(model.density.device))


BBOX_MIN = torch.tensor([-1.05, -1.05, -1.05])
BBOX_MAX = torch.tensor([1.05, 1.05, 1.05])
NEAR, FAR = 2.0, 6.0          # data/esr_nerf/esrnerf.py:77-79
MASK_ALPHA_INIT = 1e-6        # cfg/app/alphamask.yaml:17

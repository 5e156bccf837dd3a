"""Drop-in for the reference's alphamask-stage render model ``app.coarse.model.DVGO``
(app/coarse/model/dvgo.py): same constructor arguments, parameters (``density``, ``off_color``, ``emo_color``) and
``forward(**batch) -> Dict[str, Tensor]`` contract (alphamask.py:240), served by the dvgo kernels of libesr_b200.so.

The per-ray jitter of the training sampler (``torch.rand_like``, dvgo.py:163) is drawn here with the same call on the
same shape, so the same seed reproduces the reference's RNG stream on the same device; tests pass it explicitly
(``jitter=`` keyword) to compare with the CPU oracle."""
from __future__ import annotations

import ctypes
from typing import Dict

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import DvgoScene, check, ptr, stream_ptr
from .modules import cfg_get, voxel_geometry


class _DvgoRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, off_color, emo_color, sc, rays_o, rays_d, jitter, em_modes, S):
        L = _lib.lib()
        dev = rays_o.device
        n = rays_o.shape[0]
        f = dict(dtype=torch.float32, device=dev)
        alpha, raw_off, raw_emo = torch.empty(n, S, **f), torch.empty(n, S, 3, **f), torch.empty(n, S, 3, **f)
        cum, weights = torch.empty(n, S + 1, **f), torch.empty(n, S, **f)
        raw_rgb, rgb = torch.empty(n, S, 3, **f), torch.empty(n, 3, **f)
        density, off_color, emo_color = density.contiguous(), off_color.contiguous(), emo_color.contiguous()
        check(L.esr_dvgo_fwd(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(jitter), ptr(em_modes), ptr(density),
                             ptr(off_color), ptr(emo_color), n, S, ptr(alpha), ptr(raw_off), ptr(raw_emo), ptr(cum),
                             ptr(weights), ptr(raw_rgb), ptr(rgb), stream_ptr()))
        ctx.sc, ctx.S = sc, S
        ctx.save_for_backward(density, rays_o, rays_d, jitter, em_modes, alpha, raw_off, raw_emo, cum, raw_rgb)
        return cum, weights, raw_rgb, rgb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_cum, g_w, g_raw, g_rgb):
        density, rays_o, rays_d, jitter, em_modes, alpha, raw_off, raw_emo, cum, raw_rgb = ctx.saved_tensors
        n, S = rays_o.shape[0], ctx.S
        g_den = torch.zeros_like(density)
        g_off = torch.zeros(1, 3, *density.shape[2:], dtype=torch.float32, device=density.device)
        g_emo = torch.zeros_like(g_off)
        d_alpha, d_raw = torch.empty_like(alpha), torch.empty_like(raw_rgb)
        # named contiguous copies: a temporary is released as soon as ptr() returns and its block could be handed to the
        # next .contiguous() before the kernel has been queued
        g_cum, g_w, g_raw, g_rgb = g_cum.contiguous(), g_w.contiguous(), g_raw.contiguous(), g_rgb.contiguous()
        check(_lib.lib().esr_dvgo_bwd(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(jitter), ptr(em_modes),
                                      ptr(density), n, S, ptr(alpha), ptr(raw_off), ptr(raw_emo), ptr(cum), ptr(raw_rgb),
                                      ptr(g_cum), ptr(g_w), ptr(g_raw),
                                      ptr(g_rgb), ptr(d_alpha), ptr(d_raw), ptr(g_den), ptr(g_off),
                                      ptr(g_emo), stream_ptr()))
        return g_den, g_off, g_emo, None, None, None, None, None, None


class DVGO(nn.Module):
    def __init__(self, cfg, near: float, far: float, xyz_min: torch.Tensor, xyz_max: torch.Tensor):
        super().__init__()
        self.cfg = cfg
        self.device = cfg_get(cfg, "system.device")
        self.near, self.far = near, far
        self.xyz_min, self.xyz_max = xyz_min, xyz_max
        self.num_voxels = cfg_get(cfg, "app.model.num_voxels")
        self.alpha_init = cfg_get(cfg, "app.model.alpha_init")
        self.stepsize = cfg_get(cfg, "app.model.stepsize")
        self.voxel_size, self.world_size = voxel_geometry(self.xyz_min, self.xyz_max, self.num_voxels)
        self.act_shift = np.log(1 / (1 - self.alpha_init) - 1)                                 # dvgo.py:37
        ws = [int(w) for w in self.world_size]
        self.density = nn.Parameter(torch.zeros([1, 1, *ws]))
        self.off_color = nn.Parameter(torch.zeros([1, 3, *ws]))
        self.emo_color = nn.Parameter(torch.zeros([1, 3, *ws]))
        self.N_samples = int(np.linalg.norm(np.array(ws) + 1) / self.stepsize) + 1             # dvgo.py:47-50
        self.train()

    def train(self, mode=True):
        self.forward = self.forward_training if mode else self.forward_evaluate
        return super().train(mode)

    def _scene(self) -> DvgoScene:
        sc = DvgoScene()
        for i in range(3):
            sc.xyz_min[i], sc.xyz_max[i] = float(self.xyz_min[i]), float(self.xyz_max[i])
        sc.gx, sc.gy, sc.gz = (int(v) for v in self.density.shape[2:])
        sc.near, sc.far = float(self.near), float(self.far)
        sc.stepdist = float(self.stepsize * self.voxel_size)
        sc.interval, sc.act_shift = float(self.stepsize), float(self.act_shift)
        return sc

    # ---- one-off initialisation helpers of the alphamask driver (alphamask.py:128-143); dense torch code, not hot ----
    def grid_sampler(self, xyz: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
        """dvgo.py:265-277"""
        shape = xyz.shape[:-1]
        pts = xyz.reshape(1, 1, 1, -1, 3)
        ind_norm = ((pts - self.xyz_min) / (self.xyz_max - self.xyz_min)).flip((-1,)) * 2 - 1
        out = torch.nn.functional.grid_sample(grid, ind_norm, mode="bilinear", align_corners=True)
        return out.reshape(grid.shape[1], -1).T.reshape(*shape, grid.shape[1]).squeeze(-1)

    def voxel_count_views(self, rays_o: torch.Tensor, rays_d: torch.Tensor, chunk_size: int) -> torch.Tensor:
        """dvgo.py:59-93: per voxel, the number of training views whose rays deposit more than one unit of trilinear
        weight on it (the gradient of sum(grid_sample(ones)) w.r.t. the ones volume) -> [1,1,X,Y,Z]"""
        rng = torch.arange(self.N_samples, device=rays_o.device)[None].float()
        count = torch.zeros_like(self.density.detach())
        for view in range(len(rays_o)):
            ones = torch.ones_like(self.density).requires_grad_()
            for ro, rd in zip(rays_o[view].split(chunk_size, dim=0), rays_d[view].split(chunk_size, dim=0)):
                vec = torch.where(rd == 0, torch.full_like(rd, 1e-6), rd)
                t_min = torch.minimum((self.xyz_max - ro) / vec, (self.xyz_min - ro) / vec).amax(-1)
                t_min = t_min.clamp(min=self.near, max=self.far)
                step = self.stepsize * self.voxel_size * rng
                interpx = t_min[..., None] + step / rd.norm(dim=-1, keepdim=True)
                pts = ro[..., None, :] + rd[..., None, :] * interpx[..., None]
                self.grid_sampler(pts, ones).sum().backward()
            with torch.no_grad():
                count += ones.grad > 1
        return count

    @torch.no_grad()
    def maskout_near_cam_vox(self, cam_o: torch.Tensor) -> None:
        """dvgo.py:103-135: density = -100 for voxels closer than `near` to any camera centre"""
        ax = [torch.linspace(float(self.xyz_min[i]), float(self.xyz_max[i]), self.density.shape[2 + i],
                             device=self.density.device) for i in range(3)]
        xyz = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1)
        nearest = torch.stack([(xyz.unsqueeze(-2) - co).pow(2).sum(-1).sqrt().amin(-1) for co in cam_o.split(100)]).amin(0)
        self.density[nearest[None, None] <= self.near] = -100

    def activate_density(self, density, interval=1):
        return 1 - torch.exp(-torch.nn.functional.softplus(density + self.act_shift) * interval)

    def forward_training(self, **kwargs) -> Dict[str, torch.Tensor]:
        """dvgo.py:174-214"""
        rays_o = kwargs["rays_o"].contiguous().float()
        rays_d = kwargs["rays_d"].contiguous().float()
        em_modes = kwargs["em_modes"].long().contiguous()
        n = rays_o.shape[0]
        jitter = kwargs.get("jitter")
        if jitter is None:
            jitter = torch.rand_like(rays_o[:, [0]].expand(n, 1).contiguous())    # dvgo.py:163: one draw per ray
        jitter = jitter.reshape(n).contiguous().float()
        with torch.cuda.device(rays_o.device):
            cum, weights, raw_rgb, rgb = _DvgoRender.apply(self.density, self.off_color, self.emo_color, self._scene(),
                                                           rays_o, rays_d, jitter, em_modes, self.N_samples)
        return {"etc/alphainv_cum": cum, "etc/weights": weights, "etc/white_bg": cum[..., [-1]], "srgb/raw_rgb": raw_rgb,
                "srgb/rgb": rgb}

    @torch.no_grad()
    def forward_evaluate(self, **kwargs) -> Dict[str, torch.Tensor]:
        """dvgo.py:216-263"""
        rays_o = kwargs["rays_o"].contiguous().float()
        rays_d = kwargs["rays_d"].contiguous().float()
        em_modes = kwargs["em_modes"]
        L = _lib.lib()
        dev = rays_o.device
        n, S = rays_o.shape[0], self.N_samples
        f = dict(dtype=torch.float32, device=dev)
        alpha, raw_off, raw_emo = torch.empty(n, S, **f), torch.empty(n, S, 3, **f), torch.empty(n, S, 3, **f)
        cum, weights = torch.empty(n, S + 1, **f), torch.empty(n, S, **f)
        off, emo, on, depth = torch.empty(n, 3, **f), torch.empty(n, 3, **f), torch.empty(n, 3, **f), torch.empty(n, **f)
        sc = self._scene()
        den_c, off_c, emo_c = (t.detach().contiguous() for t in (self.density, self.off_color, self.emo_color))
        with torch.cuda.device(dev):
            check(L.esr_dvgo_eval(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(den_c), ptr(off_c), ptr(emo_c), n,
                                  S, ptr(alpha), ptr(raw_off), ptr(raw_emo), ptr(cum), ptr(weights), ptr(off), ptr(emo),
                                  ptr(on), ptr(depth), stream_ptr()))
        disp = 1 / (depth + cum[..., -1] * self.far)
        em = int(em_modes) if not torch.is_tensor(em_modes) else int(em_modes.item())
        return {"etc/depth": depth, "etc/disp": disp, "etc/white_bg": cum[..., [-1]], "srgb/off_rgb": off,
                "srgb/on_rgb": on, "srgb/emo_rgb": emo, "srgb/rgb": off if em == 0 else on}

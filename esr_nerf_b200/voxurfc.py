"""Drop-in for the reference's coarse-stage render model ``app.coarse.model.VoxurfC``
(app/coarse/model/voxurfc.py): same constructor arguments, ``state_dict`` keys/shapes and
``forward(**batch) -> Dict[str, Tensor]`` contract (coarse.py:338).

    renderer = VoxurfC(cfg, near, far, xyz_min, xyz_max, mask_xyz_min, mask_xyz_max, mask_alpha_init,
                       mask_density, s_val)
    results = renderer(s_val=s_val, **batch)

Stages on the library's kernels: march + MaskCache + SDF tap (on the Gaussian-smoothed grid), NeuS alpha, the first
transmittance scan and its weight filter, the second Alphas2Weights pass (reference-shaped op), the feature encode
(colour taps, PE, normal from the central-difference gradient volume) and its scatter backward, compositing.
The two 57->128->128->3 colour MLPs (voxurfc.py:137-169) run on the 96->192x3 tcgen05 chains, zero-padded and with an
identity third hidden layer (modules.flat_coarse_mlp_params; 3.8x the MMA work of a native-width kernel — DESIGN.md §7):
``mlp_mode`` "x2" (default: outputs ~1e-6, every gradient within 1e-2 of the reference's fp32 nets), "bf16", or
"torch_fp32" (library GEMMs, the 1e-4 class of the strict tests).  Library call (like the reference): the dense 5^3
Gaussian smoothing of the SDF grid (cuDNN conv3d, voxurfc.py:202)."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import fused
from .modules import (DenseGrid, GradientConv, GridRegularizers, MaskCache, RayUtilities, _mlp_stack, cfg_get, coarse_src_cols,
                      flat_coarse_mlp_params, host_geometry, voxel_geometry)
from .render_utils import Alphas2Weights


class Gaussian3DConv(nn.Module):
    """module.py:146-177 — fixed, non-trainable Gaussian smoothing (state_dict keys smooth_conv.m.{weight,bias})."""

    def __init__(self, ksize: int = 3, sigma: float = 1.0):
        super().__init__()
        r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
        xx, yy, zz = np.meshgrid(r, r, r)
        kernel = torch.FloatTensor(np.exp(-(xx ** 2 + yy ** 2 + zz ** 2) / (2 * sigma ** 2)))
        self.m = nn.Conv3d(1, 1, ksize, stride=1, padding=ksize // 2, padding_mode="replicate")
        self.m.weight.data = (kernel / kernel.sum())[None, None]
        self.m.bias.data = torch.zeros(1)
        for p in self.m.parameters():
            p.requires_grad = False

    def forward(self, x):
        return self.m(x)


class VoxurfC(GridRegularizers, RayUtilities, nn.Module):
    def __init__(self, cfg, near: float, far: float, xyz_min, xyz_max, mask_xyz_min, mask_xyz_max,
                 mask_alpha_init: float, mask_density: torch.Tensor, s_val: float):
        super().__init__()
        self.cfg = cfg
        self.device = cfg_get(cfg, "system.device")
        m = "app.model."
        self.near, self.far = near, far
        self.xyz_min = xyz_min.to(self.device).float()
        self.xyz_max = xyz_max.to(self.device).float()
        self.mask_xyz_min = mask_xyz_min.to(self.device).float()
        self.mask_xyz_max = mask_xyz_max.to(self.device).float()
        self.mask_alpha_init = mask_alpha_init
        self.mask_density = mask_density.to(self.device).float()
        self.s_val = s_val
        for k in ("mask_ks", "maskcache_thres", "fastcolor_thres", "stepsize", "num_voxels", "color_dim", "rgbnet_width",
                  "rgbnet_depth", "posbase_pe", "viewbase_pe", "smooth_ksize", "smooth_sigma", "neus_alpha"):
            setattr(self, k, cfg_get(cfg, m + k))
        if not (self.color_dim == 12 and self.posbase_pe == 5 and self.viewbase_pe == 1 and
                self.neus_alpha in ("interp", "grad")):
            raise NotImplementedError("libesr_b200 instantiates the shipped coarse-stage shape only "
                                      "(cfg/app/coarse.yaml:13-31): color_dim 12, PE 5/1, neus_alpha interp | grad")
        self.voxel_size, self.world_size = voxel_geometry(self.xyz_min, self.xyz_max, self.num_voxels)
        ws = self.world_size
        self.sdf = DenseGrid(1, ws, self.xyz_min, self.xyz_max)
        ax = [torch.linspace(-1.0, 1.0, int(w), dtype=torch.float64) for w in ws]
        gx, gy, gz = torch.meshgrid(*ax, indexing="ij")
        self.sdf.grid.data = ((gx ** 2 + gy ** 2 + gz ** 2) ** 0.5 - 1).float()[None, None]   # voxurfc.py:86-95
        self.smooth_conv = Gaussian3DConv(self.smooth_ksize, self.smooth_sigma)
        self.tv_smooth_conv = GradientConv()
        self.mask_cache = MaskCache(self.mask_xyz_min, self.mask_xyz_max, self.mask_density, self.mask_alpha_init,
                                    self.maskcache_thres, self.mask_ks)
        self.off_color = DenseGrid(self.color_dim, ws, self.xyz_min, self.xyz_max)
        dim0 = (3 + 3 * self.posbase_pe * 2) + (3 * self.viewbase_pe * 3) + self.color_dim + 3
        self.off_rgbnet = _mlp_stack(dim0, self.rgbnet_width, self.rgbnet_depth, 3)
        nn.init.constant_(self.off_rgbnet[-1].bias, 0)
        self.emo_color = DenseGrid(self.color_dim, ws, self.xyz_min, self.xyz_max)
        self.emo_rgbnet = _mlp_stack(dim0, self.rgbnet_width, self.rgbnet_depth, 3)
        nn.init.constant_(self.emo_rgbnet[-1].bias, 0)
        self.to(self.device)
        self.set_nonempty_mask()
        self.keep_streams = False
        self.last_streams = None
        self.mlp_mode = "x2"        # "x2" | "bf16" (tcgen05 chains) | "torch_fp32" (library GEMMs)
        self.train()

    def train(self, mode=True):
        self.forward = self.forward_training if mode else self.forward_evaluate
        return super().train(mode)

    def _colour_nets(self, x: torch.Tensor):
        """(sigmoid(off_rgbnet(.)), sigmoid(emo_rgbnet(.))) of voxurfc.py:229-240 on the [M3,72] feature rows"""
        if self.mlp_mode == "torch_fp32":
            feat = x[:, 24:69]
            return (torch.sigmoid(self.off_rgbnet(torch.cat([x[:, 0:12], feat], -1))),
                    torch.sigmoid(self.emo_rgbnet(torch.cat([x[:, 12:24], feat], -1))))
        if self.mlp_mode not in ("x2", "bf16"):
            raise ValueError(f"unknown mlp_mode {self.mlp_mode!r}")
        dev = x.device
        maps = self.__dict__.get("_src_maps")
        if maps is None or maps[0].device != dev:
            maps = self.__dict__["_src_maps"] = (coarse_src_cols("off", dev), coarse_src_cols("emo", dev))
        return fused.CoarseShade.apply(x, flat_coarse_mlp_params(self.off_rgbnet), flat_coarse_mlp_params(self.emo_rgbnet),
                                       maps[0], maps[1], 1 if self.mlp_mode == "x2" else 0)

    @torch.no_grad()
    def set_nonempty_mask(self):
        """voxurfc.py:491-521"""
        ax = [torch.linspace(float(self.xyz_min[i]), float(self.xyz_max[i]), self.sdf.grid.shape[2 + i],
                             device=self.sdf.grid.device) for i in range(3)]
        xyz = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1)
        self.nonempty_mask = self.mask_cache(xyz)[None, None].contiguous()
        self.sdf.grid[~self.nonempty_mask] = 1

    def neus_sdf_gradient(self) -> torch.Tensor:
        """voxurfc.py:597-616 — dense central differences of the RAW sdf grid, channels (d/dx, d/dy, d/dz)"""
        g = self.sdf.grid
        if g.is_cuda:      # one kernel each way (csrc/regularizers.cu) instead of zeros + three sliced assignments
            return fused.SdfCentralGradient.apply(g, host_geometry(self, self.stepsize)["voxel_size"])
        out = torch.zeros([1, 3, *g.shape[-3:]], device=g.device)
        out[:, 0, 1:-1, :, :] = (g[:, 0, 2:, :, :] - g[:, 0, :-2, :, :]) / 2 / self.voxel_size
        out[:, 1, :, 1:-1, :] = (g[:, 0, :, 2:, :] - g[:, 0, :, :-2, :]) / 2 / self.voxel_size
        out[:, 2, :, :, 1:-1] = (g[:, 0, :, :, 2:] - g[:, 0, :, :, :-2]) / 2 / self.voxel_size
        return out

    def _scene(self, s_val: float):
        g = self.sdf.grid.shape
        md = self.mask_cache.density.shape
        h = host_geometry(self, self.stepsize)
        return fused.make_scene(h["xyz_min"], h["xyz_max"], g[2:], h["mask_xyz_min"], h["mask_xyz_max"], md[2:],
                                self.near, 1e9, h["stepdist"], h["voxel_size"],
                                self.mask_cache.act_shift, self.maskcache_thres, self.fastcolor_thres, s_val,
                                alpha_thres=-1.0)     # the coarse stage has no alpha filter before the scan

    def forward_training(self, **kwargs) -> Dict[str, torch.Tensor]:
        """voxurfc.py:186-271"""
        rays_o = kwargs["rays_o"].contiguous().float()
        rays_d = kwargs["rays_d"].contiguous().float()
        viewdirs = kwargs["viewdirs"].contiguous().float()
        em_modes = kwargs["em_modes"].long()
        self.s_val = kwargs["s_val"]
        N = rays_o.shape[0]
        dev = rays_o.device
        with torch.cuda.device(dev):
            sc = self._scene(float(self.s_val))
            for g in (self.sdf, self.off_color, self.emo_color):
                g.ensure_layout()
            sdf_grid = self.smooth_conv(self.sdf.grid).contiguous()                      # voxurfc.py:202
            self.gradient = self.neus_sdf_gradient()                                     # voxurfc.py:205
            s = fused.march(sc, rays_o, rays_d, None, self.mask_cache.density, sdf_grid.detach())
            grad_mode = self.neus_alpha == "grad"                                        # voxurfc.py:171-174
            h_alpha = fused.CoarseAlpha.apply(sdf_grid, sc, rays_o, rays_d, s, self.gradient if grad_mode else None,
                                              viewdirs if grad_mode else None)           # voxurfc.py:208-218
            ray_id = s.h_ray.long()
            weights, last = Alphas2Weights.apply(h_alpha, ray_id, N)                     # voxurfc.py:219
            x = fused.EncodeCoarse.apply(self.gradient, self.off_color.grid, self.emo_color.grid, sc, rays_o, rays_d,
                                         viewdirs, s)
            on = em_modes[ray_id] == 1
            rgb_off, rgb_emo = self._colour_nets(x)
            rgb = torch.where(on[:, None], rgb_emo, torch.zeros_like(rgb_emo)) + rgb_off  # voxurfc.py:241-249
            rgb_marched, wsum = fused.Composite.apply(weights, rgb, torch.ones_like(rgb), s)
        if self.keep_streams:
            self.last_streams = dict(streams=s, h_w=weights.detach(), rgb=rgb.detach())
        return {"etc/alphainv_cum": last, "etc/white_bg": 1 - wsum[:, :1], "srgb/rgb": rgb_marched}

    @torch.no_grad()
    def forward_evaluate(self, **kwargs) -> Dict[str, torch.Tensor]:
        """voxurfc.py:273-424"""
        rays_o = kwargs["rays_o"].contiguous().float()
        rays_d = kwargs["rays_d"].contiguous().float()
        viewdirs = kwargs["viewdirs"].contiguous().float()
        em_modes = kwargs["em_modes"]
        N = rays_o.shape[0]
        dev = rays_o.device
        pos_rt = kwargs["pos_rt"].to(dev).float()
        with torch.cuda.device(dev):
            sc = self._scene(float(self.s_val))
            for g in (self.sdf, self.off_color, self.emo_color):
                g.ensure_layout()
            sdf_grid = self.smooth_conv(self.sdf.grid).contiguous()
            self.gradient = self.neus_sdf_gradient()
            s = fused.march(sc, rays_o, rays_d, None, self.mask_cache.density, sdf_grid)
            grad_mode = self.neus_alpha == "grad"
            h_alpha = fused.CoarseAlpha.apply(sdf_grid, sc, rays_o, rays_d, s, self.gradient if grad_mode else None,
                                              viewdirs if grad_mode else None)
            if s.m3 <= 1:                                                              # voxurfc.py:322-335
                z3 = torch.zeros_like(rays_o)
                return {"etc/depth": z3[..., 0], "etc/disp": 1 / (z3[..., 0] + self.far), "etc/normal": z3,
                        "etc/white_bg": torch.ones_like(z3[..., :1]), "srgb/off_rgb": z3, "srgb/emo_rgb": z3,
                        "srgb/on_rgb": z3, "srgb/rgb": z3}
            ray_id = s.h_ray.long()
            weights, _ = Alphas2Weights.apply(h_alpha, ray_id, N)
            x = fused.EncodeCoarse.apply(self.gradient, self.off_color.grid, self.emo_color.grid, sc, rays_o, rays_d,
                                         viewdirs, s)
            off, emo = self._colour_nets(x)
            normal = ((x[:, 66:69] @ pos_rt) * torch.tensor([1.0, -1.0, -1.0], device=dev) + 1.0) / 2.0
            dvec = torch.ones(s.m3, 3, device=dev)
            dvec[:, 0] = s.h_step.float() * host_geometry(self, self.stepsize)["stepdist"]
            off_m, emo_m = fused.composite_infer(weights, off, emo, s)
            on_m, nrm_m = fused.composite_infer(weights, off + emo, normal, s)
            dw, _ = fused.composite_infer(weights, dvec, dvec, s)                      # (depth, sum w, sum w)
        depth, bg = dw[:, 0].contiguous(), 1 - dw[:, 1:2]
        disp = 1 / (depth + bg[..., -1] * self.far)
        em = int(em_modes) if not torch.is_tensor(em_modes) else int(em_modes.item())
        if self.keep_streams:
            self.last_streams = dict(streams=s, h_w=weights)
        return {"etc/depth": depth, "etc/disp": disp, "etc/normal": nrm_m, "etc/white_bg": bg, "srgb/off_rgb": off_m,
                "srgb/emo_rgb": emo_m, "srgb/on_rgb": on_m, "srgb/rgb": off_m if em == 0 else on_m}


_ = F

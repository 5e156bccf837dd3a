"""Parameter containers that mirror the reference's module tree so ``state_dict`` keys, shapes and
constructor arguments match — checkpoints and the optimizer's name->group mapping
(app/utils/optimizer.py:11-60) keep working.  The forward math lives in the CUDA library; these
classes only own parameters, expose them to the fused path, and provide the few one-off helpers the
stage drivers call (grid rescale, TV add-grad).

Reference counterparts: DenseGrid / MaskCache / GradientConv (app/utils/base/module.py:9-114,180-211),
RadianceNet / TonemapNet (app/utils/pbr/module.py:6-39).
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import render_utils

CL3D = torch.channels_last_3d


def cfg_get(cfg, path: str):
    """Read ``a.b.c`` from an omegaconf DictConfig, a nested dict or any attribute object."""
    cur = cfg
    for key in path.split("."):
        cur = cur[key] if isinstance(cur, dict) and not hasattr(cur, key) else getattr(cur, key)
    return cur


class DenseGrid(nn.Module):
    """[1,C,X,Y,Z] fp32 feature volume (module.py:9-75).  Logical shape and state_dict key equal the
    reference's; multi-channel grids are held in channels-last memory (voxel-major, C contiguous) so a
    trilinear corner is one contiguous 4*C-byte read / vector RED instead of C strided ones."""

    def __init__(self, channels: int, world_size, xyz_min: torch.Tensor, xyz_max: torch.Tensor):
        super().__init__()
        self.channels = channels
        self.world_size = world_size
        self.xyz_min = xyz_min
        self.xyz_max = xyz_max
        ws = [int(w) for w in world_size]
        g = torch.zeros([1, channels, *ws])
        self.grid = nn.Parameter(self._layout(g))

    def _layout(self, g: torch.Tensor) -> torch.Tensor:
        return g.contiguous(memory_format=CL3D) if self.channels > 1 else g.contiguous()

    def ensure_layout(self):
        g = self.grid.data
        want_cl = self.channels > 1
        ok = g.is_contiguous(memory_format=CL3D) if want_cl else g.is_contiguous()
        if not ok:
            self.grid.data = self._layout(g)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.ensure_layout()

    @torch.no_grad()
    def scale_volume_grid(self, new_world_size):
        """module.py:37-49"""
        self.world_size = new_world_size
        size = tuple(int(w) for w in new_world_size)
        new = F.interpolate(self.grid.data.contiguous(), size=size, mode="trilinear", align_corners=True)
        self.grid = nn.Parameter(self._layout(new))

    def total_variation_add_grad(self, wx, wy, wz, dense_mode, mask=None):
        """module.py:51-64 (the masked variant is dead code in the reference and is not provided)."""
        if mask is not None:
            raise NotImplementedError("masked TV is never called by the reference (voxurff.py:621)")
        assert self.channels == 1, "TV add-grad is applied to the SDF grid only (voxurff.py:619-621)"
        render_utils.total_variation_add_grad(self.grid.data, self.grid.grad, wx, wy, wz, dense_mode)

    def get_dense_grid(self):
        return self.grid

    def extra_repr(self):
        return f"channels={self.channels}, world_size={[int(w) for w in self.world_size]}"


class MaskCache(nn.Module):
    """module.py:78-114.  Holds the max-pooled density; the lookup itself is fused into the march
    kernel (esr_march_*).  ``forward`` is kept for the one-off callers (set_nonempty_mask)."""

    def __init__(self, xyz_min, xyz_max, density, alpha_init: float, cache_thres: float, ks: int):
        super().__init__()
        self.xyz_min = xyz_min
        self.xyz_max = xyz_max
        self.mask_cache_thres = cache_thres
        self.ks = ks
        self.density = F.max_pool3d(density, kernel_size=ks, padding=ks // 2, stride=1).contiguous()
        self.act_shift = float(np.log(1 / (1 - alpha_init) - 1))

    @torch.no_grad()
    def forward(self, xyz):
        shape = xyz.shape[:-1]
        pts = xyz.reshape(1, 1, 1, -1, 3)
        ind_norm = ((pts - self.xyz_min) / (self.xyz_max - self.xyz_min)).flip((-1,)) * 2 - 1
        d = F.grid_sample(self.density, ind_norm, align_corners=True)
        alpha = 1 - torch.exp(-F.softplus(d + self.act_shift))
        return alpha.reshape(*shape) >= self.mask_cache_thres


def _mlp_stack(in_dim: int, width: int, depth: int, out_dim: int) -> nn.Sequential:
    # Linear, ReLU, (Linear, ReLU) x (depth-2) nested one level, Linear -> keys "0", "2.0", "3.0", "4"
    return nn.Sequential(
        nn.Linear(in_dim, width),
        nn.ReLU(inplace=True),
        *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
        nn.Linear(width, out_dim),
    )


def _linears(seq: nn.Sequential) -> List[nn.Linear]:
    return [m for m in seq.modules() if isinstance(m, nn.Linear)]


class RadianceNet(nn.Module):
    """pbr/module.py:6-21: dim0 -> width x (depth-1) -> 3, softplus."""

    def __init__(self, inputdim: int, width: int, depth: int):
        super().__init__()
        self.linear = _mlp_stack(inputdim, width, depth, 3)

    def layers(self):
        return _linears(self.linear)

    def forward(self, x):  # fp32 library path (strict mode / validation)
        return F.softplus(self.linear(x))


class TonemapNet(nn.Module):
    """pbr/module.py:24-39: dim0 -> width x (depth-1) -> 3, sigmoid."""

    def __init__(self, dim0: int, width: int, depth: int):
        super().__init__()
        self.srgb = _mlp_stack(dim0, width, depth, 3)

    def layers(self):
        return _linears(self.srgb)

    def forward(self, x):
        return torch.sigmoid(self.srgb(x))


class GradientConv(nn.Module):
    """module.py:180-211 — fixed 3x3x3 smoothing kernel used by the TV regulariser; kept so the
    state_dict carries ``tv_smooth_conv.m.{weight,bias}`` like the reference's."""

    def __init__(self, sigma: float = 0):
        super().__init__()
        base = np.array([1.0, 2.0, 1.0])
        kernel = base[:, None, None] * base[None, :, None] * base[None, None, :]
        idx = np.arange(3) - 1
        dist = idx[:, None, None] ** 2 + idx[None, :, None] ** 2 + idx[None, None, :] ** 2 - 1
        kernel = kernel * np.exp(-dist * sigma)
        self.m = nn.Conv3d(1, 1, (3, 3, 3), stride=1, padding=1, padding_mode="replicate")
        self.m.weight.data = torch.from_numpy(kernel / kernel.sum()).float()[None, None]
        self.m.bias.data = torch.zeros(1)
        for p in self.m.parameters():
            p.requires_grad = False

    def forward(self, x):
        return self.m(x)


# ------------------------------------------------------------------------------------------------
# flat parameter images for the tensor-core MLP kernels
# ------------------------------------------------------------------------------------------------
def flat_mlp_params(layers: Sequence[nn.Linear], in_cols: torch.Tensor, k0: int) -> torch.Tensor:
    """Assemble the f32 master copy esr_mlp_pack consumes (include/esr_b200.h): per layer W then b;
    layer 0 columns are gathered into the kernel's internal column order (``in_cols[c]`` = reference
    input column feeding internal column c, or -1 for a zero column); the output layer is padded to 8
    rows.  Built with differentiable torch ops so autograd routes the flat gradient back to the
    nn.Linear parameters."""
    first, last = layers[0], layers[-1]
    w0 = torch.cat([first.weight, first.weight.new_zeros(first.weight.shape[0], 1)], 1)
    idx = torch.where(in_cols < 0, torch.full_like(in_cols, first.weight.shape[1]), in_cols)
    parts = [w0[:, idx].reshape(-1), first.bias]
    for lin in layers[1:-1]:
        parts += [lin.weight.reshape(-1), lin.bias]
    pad = 8 - last.weight.shape[0]
    parts += [F.pad(last.weight, (0, 0, 0, pad)).reshape(-1), F.pad(last.bias, (0, pad))]
    assert idx.numel() == k0
    return torch.cat(parts)


def radiance_in_cols(which: str, device) -> torch.Tensor:
    """Internal 96-column feature row (include/esr_b200.h, Stage E) -> reference 85-column input
    [color 0-5 | xyz 6-8 | sin 9-23 | cos 24-38 | view 39-47 | sdf 48 | feat 49-72 | normal 73-84]
    (voxurff.py:228-254)."""
    cols = [-1] * 96
    color = list(range(0, 6))
    if which == "off":
        cols[0:6] = color
    else:
        cols[6:12] = color
    cols[12] = 48
    cols[13:37] = range(49, 73)
    cols[37:49] = range(73, 85)
    cols[49:52] = range(6, 9)
    cols[52:67] = range(9, 24)
    cols[67:82] = range(24, 39)
    cols[82:91] = range(39, 48)
    return torch.tensor(cols, dtype=torch.long, device=device)


def tonemap_in_cols(device) -> torch.Tensor:
    return torch.tensor(list(range(33)) + [-1] * 15, dtype=torch.long, device=device)


def voxel_geometry(xyz_min: torch.Tensor, xyz_max: torch.Tensor, num_voxels: int):
    """voxurff.py:539-545 (set_grid_resolution), evaluated with the same float32 torch ops."""
    voxel_size = ((xyz_max - xyz_min).prod() / num_voxels).pow(1 / 3)
    world_size = ((xyz_max - xyz_min) / voxel_size).long()
    return voxel_size, world_size


def n_candidate_steps(world_size, stepsize: float) -> int:
    """voxurff.py:509-512"""
    return int(np.linalg.norm(np.array([int(w) for w in world_size]) + 1) / stepsize) + 1


_ = math  # keep import for downstream users

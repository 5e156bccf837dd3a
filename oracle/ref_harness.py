"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Imports the reference's OWN render models (``app/fine/model/voxurff.py`` etc.)
from ``/root/reference`` on the CPU, in this container only, so that

* ``oracle/voxurf_port.py`` (the travelling torch restatement) can be pinned
  against the real reference code, and
* ``oracle/make_golden.py`` can generate the committed fixtures in
  ``tests/golden/``.

The reference cannot be imported as-is: it JIT-builds CUDA at
import (``app/utils/base/functions.py:9-31``), and needs ``torch_scatter``,
``omegaconf``, ``mcubes`` ... which are absent.  This module therefore
 1. registers empty namespace packages for ``app``/``app.*`` so the reference's
    package ``__init__`` files (which pull in the train drivers) are skipped,
 2. registers stub modules for the missing third-party packages,
 3. replaces ``torch.utils.cpp_extension.load`` so the three live native ops
    (``sample_pts_on_rays``, ``alpha2weight``, ``alpha2weight_backward``) are
    served by the C restatement ``oracle/render_utils_ref.c`` through ctypes,
 4. serves ``torch_scatter.segment_coo`` by ``out.index_add_`` (sum over a
    sorted index; third-party, un-pinned — see render_utils_ref.c header).
Nothing here is available on the GPU box (``/root/reference`` does not travel).
"""
from __future__ import annotations

import ctypes
import importlib
import os
import subprocess
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ESR_REFERENCE_ROOT", "/root/reference")
_LIB = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "app", "fine", "model", "voxurff.py"))


# ----------------------------------------------------------------------------
# C oracle (render_utils_ref.c) through ctypes
# ----------------------------------------------------------------------------
def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "render_utils_ref.c")
    out = os.path.join(HERE, "liboracle_ref.so")
    if force or not os.path.isfile(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
             "-o", out, src, "-lm"]
        )
    return out


def c_oracle():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        P, I64, F32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
        lib.oracle_sample_pts_on_rays.restype = I64
        lib.oracle_sample_pts_on_rays.argtypes = [P, P, P, P, F32, F32, F32, I64, P, P, P, P, P, P, P]
        lib.oracle_alpha2weight.restype = None
        lib.oracle_alpha2weight.argtypes = [P, P, I64, I64, P, P, P, P, P]
        lib.oracle_alpha2weight_backward.restype = None
        lib.oracle_alpha2weight_backward.argtypes = [P, P, P, P, P, P, I64, I64, P, P, P]
        lib.oracle_segment_coo_sum.restype = None
        lib.oracle_segment_coo_sum.argtypes = [P, P, I64, I64, P]
        _LIB = lib
    return _LIB


def _p(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """CPU stand-in for render_utils_cuda.sample_pts_on_rays (render_utils.cpp:74-85)."""
    lib = c_oracle()
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    stepdist = float(stepdist)
    N_steps = torch.empty(n, dtype=torch.int64)
    t_min = torch.empty(n, dtype=torch.float32)
    t_max = torch.empty(n, dtype=torch.float32)
    total = lib.oracle_sample_pts_on_rays(
        _p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), near, far, stepdist, n,
        None, None, None, None, _p(N_steps), _p(t_min), _p(t_max))
    ray_pts = torch.empty(total, 3, dtype=torch.float32)
    mask = torch.empty(total, dtype=torch.uint8)
    ray_id = torch.empty(total, dtype=torch.int64)
    step_id = torch.empty(total, dtype=torch.int64)
    lib.oracle_sample_pts_on_rays(
        _p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), near, far, stepdist, n,
        _p(ray_pts), _p(mask), _p(ray_id), _p(step_id), _p(N_steps), _p(t_min), _p(t_max))
    return [ray_pts, mask.bool(), ray_id, step_id, N_steps, t_min, t_max]


def alpha2weight(alpha, ray_id, n_rays):
    """CPU stand-in for render_utils_cuda.alpha2weight (render_utils.cpp:142-149)."""
    lib = c_oracle()
    alpha = _f32(alpha)
    ray_id = ray_id.to(torch.int64).contiguous()
    m = alpha.shape[0]
    weight = torch.empty(m, dtype=torch.float32)
    T = torch.empty(m, dtype=torch.float32)
    last = torch.empty(n_rays, dtype=torch.float32)
    i_start = torch.empty(n_rays, dtype=torch.int64)
    i_end = torch.empty(n_rays, dtype=torch.int64)
    lib.oracle_alpha2weight(_p(alpha), _p(ray_id), m, n_rays, _p(weight), _p(T), _p(last),
                            _p(i_start), _p(i_end))
    return [weight, T, last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays,
                          grad_weights, grad_last):
    """CPU stand-in for render_utils_cuda.alpha2weight_backward (render_utils.cpp:151-167)."""
    lib = c_oracle()
    alpha, weight, T, alphainv_last = map(_f32, (alpha, weight, T, alphainv_last))
    grad_weights, grad_last = _f32(grad_weights), _f32(grad_last)
    m = alpha.shape[0]
    grad = torch.empty(m, dtype=torch.float32)
    lib.oracle_alpha2weight_backward(_p(alpha), _p(weight), _p(T), _p(alphainv_last),
                                     _p(i_start.contiguous()), _p(i_end.contiguous()), m, n_rays,
                                     _p(grad_weights), _p(grad_last), _p(grad))
    return grad


def segment_coo(src, index, out=None, reduce="sum"):
    """torch_scatter.segment_coo(reduce='sum') semantics (differentiable)."""
    assert reduce == "sum" and out is not None
    return out.index_add(0, index, src)


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    """total_variation_kernel.cu:14-35 (dense mode), incl. quirk Q4 (wz on the k axis, wx unused)."""
    assert dense_mode
    p = param.detach()
    wx, wy, wz = (w / 6.0 for w in (wx, wy, wz))
    g = torch.zeros_like(p)

    def cl(x):
        return x.clamp(-1, 1)

    g[:, :, :, :, 1:] += wz * cl(p[:, :, :, :, 1:] - p[:, :, :, :, :-1])
    g[:, :, :, :, :-1] += wz * cl(p[:, :, :, :, :-1] - p[:, :, :, :, 1:])
    g[:, :, :, 1:, :] += wy * cl(p[:, :, :, 1:, :] - p[:, :, :, :-1, :])
    g[:, :, :, :-1, :] += wy * cl(p[:, :, :, :-1, :] - p[:, :, :, 1:, :])
    g[:, :, 1:, :, :] += wz * cl(p[:, :, 1:, :, :] - p[:, :, :-1, :, :])
    g[:, :, :-1, :, :] += wz * cl(p[:, :, :-1, :, :] - p[:, :, 1:, :, :])
    grad += g


# ----------------------------------------------------------------------------
# stubs + import of the reference
# ----------------------------------------------------------------------------
class DictConfig(dict):
    """Minimal attribute-dict stand-in for omegaconf.DictConfig."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = DictConfig(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT} (it does not travel to the GPU box)")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("omegaconf", DictConfig=DictConfig, OmegaConf=object)
    mod("hydra")
    mod("hydra.core")
    mod("hydra.core.hydra_config", HydraConfig=object)
    mod("mcubes", marching_cubes=None)
    mod("imageio")
    mod("lpips")
    mod("trimesh", Trimesh=object, points=types.SimpleNamespace(PointCloud=object))
    mod("torch_scatter", segment_coo=segment_coo)
    mod("wandb", config={"system": {"debug": True, "tqdm_iters": 1}})
    # utils2.utils (samplers, tqdm helper) imports rich/wandb; only tqdm_safe is used by esrnerf.py:40
    mod("utils2")
    mod("utils2.utils", tqdm_safe=lambda it, **kw: it,
        LightDict={"off": 0, "on": 1, "i_change": 2, "c_change": 3, "ic_change": 4})

    # namespace packages: skip the reference's package __init__ files
    for pkg in ["app", "app.utils", "app.utils.base", "app.utils.pbr", "app.fine",
                "app.fine.model", "app.coarse", "app.coarse.model"]:
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF_ROOT, *pkg.split("."))]
        sys.modules[pkg] = m

    # native extension shims
    import torch.utils.cpp_extension as cpp_ext

    shims = {
        "render_utils_cuda": types.SimpleNamespace(
            sample_pts_on_rays=sample_pts_on_rays, alpha2weight=alpha2weight,
            alpha2weight_backward=alpha2weight_backward),
        "total_variation_cuda": types.SimpleNamespace(
            total_variation_add_grad=total_variation_add_grad),
    }
    real_load = cpp_ext.load
    real_name = torch.cuda.get_device_name
    real_makedirs = os.makedirs
    cpp_ext.load = lambda name, **kw: shims[name]
    torch.cuda.get_device_name = lambda *a, **k: "cpu"
    os.makedirs = lambda *a, **k: None  # functions.py:13 would mkdir inside the read-only tree
    try:
        importlib.import_module("app.utils.base.functions")
    finally:
        cpp_ext.load = real_load
        torch.cuda.get_device_name = real_name
        os.makedirs = real_makedirs
    _installed = True


def reference_classes():
    """Returns the reference's own classes: (DVGO, VoxurfC, VoxurfF, ESRNeRF)."""
    install_stubs()
    dvgo = importlib.import_module("app.coarse.model.dvgo").DVGO
    voxurfc = importlib.import_module("app.coarse.model.voxurfc").VoxurfC
    voxurff = importlib.import_module("app.fine.model.voxurff").VoxurfF
    esrnerf = importlib.import_module("app.fine.model.esrnerf").ESRNeRF
    return dvgo, voxurfc, voxurff, esrnerf


def fine_cfg(device="cpu", **model_overrides) -> DictConfig:
    """cfg subset read by VoxurfF.__init__ (voxurff.py:61-77); values = cfg/app/fine.yaml:13-30."""
    model = dict(mask_ks=3, maskcache_thres=1e-3, fastcolor_thres=1e-4, stepsize=0.5,
                 color_dim=6, rgbnet_width=192, rgbnet_depth=4, tonemap_width=192,
                 tonemap_depth=2, posbase_pe=5, viewbase_pe=1, colorbase_pe=5,
                 grad_feat=[0.5, 1.0, 1.5, 2.0], neus_alpha="interp")
    model.update(model_overrides)
    return DictConfig(dict(system=dict(device=device), app=dict(model=model)))

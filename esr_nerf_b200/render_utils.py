"""Drop-in replacements for the reference's native render ops, with the reference's own names,
argument meaning, return values and error behaviour (RuntimeError on non-CUDA / non-contiguous
input), served by libesr_b200.so through the C ABI.

    from esr_nerf_b200.render_utils import render_utils_cuda, total_variation_cuda, segment_coo, Alphas2Weights

* ``render_utils_cuda.sample_pts_on_rays / alpha2weight / alpha2weight_backward``
  <- app/utils/base/cuda/render_utils.cpp:74-85, 142-167 (the live entries of the pybind module)
* ``total_variation_cuda.total_variation_add_grad`` <- app/utils/base/cuda/total_variation.cpp:29-32
* ``segment_coo(src, index, out, reduce="sum")``    <- torch_scatter, as called at voxurff.py:260-272
* ``Alphas2Weights``                                 <- app/utils/base/module.py:117-143
"""
from __future__ import annotations

import types

import torch

from . import _lib
from ._lib import check, f3, ptr, stream_ptr


def _check_input(t: torch.Tensor, name: str) -> None:
    # render_utils.cpp:46-48 CHECK_INPUT
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _scratch(n: int, device) -> torch.Tensor:
    return torch.empty(_lib.lib().esr_scan_scratch_bytes(int(n)), dtype=torch.uint8, device=device)


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """-> [ray_pts [M,3] f32, mask_outbbox [M] bool, ray_id [M] i64, step_id [M] i64, N_steps [N] i64,
    t_min [N] f32, t_max [N] f32]   (render_utils_kernel.cu:196-242)"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (xyz_min, "xyz_min"), (xyz_max, "xyz_max")):
        _check_input(t, n)
    L = _lib.lib()
    dev = rays_o.device
    n = rays_o.shape[0]
    stepdist = float(stepdist)
    mn, mx = f3(xyz_min), f3(xyz_max)
    N_steps = torch.empty(n, dtype=torch.int64, device=dev)
    N_cum = torch.empty(n, dtype=torch.int64, device=dev)
    t_min = torch.empty(n, dtype=torch.float32, device=dev)
    t_max = torch.empty(n, dtype=torch.float32, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        scratch = _scratch(n, dev)   # named: must stay allocated until the launch has been queued
        check(L.esr_sample_pts_on_rays_count(ptr(rays_o), ptr(rays_d), mn, mx, float(near), float(far), stepdist, n,
                                             ptr(N_steps), ptr(N_cum), ptr(t_min), ptr(t_max), ptr(total),
                                             ptr(scratch), stream_ptr()))
        m = int(total.item())  # the reference syncs here as well (kernel.cu:212)
        ray_pts = torch.empty(m, 3, dtype=torch.float32, device=dev)
        mask = torch.empty(m, dtype=torch.bool, device=dev)
        ray_id = torch.empty(m, dtype=torch.int64, device=dev)
        step_id = torch.empty(m, dtype=torch.int64, device=dev)
        check(L.esr_sample_pts_on_rays_fill(ptr(rays_o), ptr(rays_d), mn, mx, float(near), float(far), stepdist, n,
                                            ptr(N_cum), m, ptr(ray_pts), ptr(mask), ptr(ray_id), ptr(step_id),
                                            stream_ptr()))
    return [ray_pts, mask, ray_id, step_id, N_steps, t_min, t_max]


def alpha2weight(alpha, ray_id, n_rays):
    """-> [weight, T, alphainv_last, i_start, i_end]   (render_utils_kernel.cu:619-651)"""
    _check_input(alpha, "alpha")
    _check_input(ray_id, "ray_id")
    assert alpha.dim() == 1 and ray_id.dim() == 1 and alpha.shape == ray_id.shape
    dev = alpha.device
    m = alpha.shape[0]
    n_rays = int(n_rays)
    weight = torch.empty_like(alpha)
    T = torch.empty_like(alpha)
    last = torch.empty(n_rays, dtype=alpha.dtype, device=dev)
    i_start = torch.empty(n_rays, dtype=torch.int64, device=dev)
    i_end = torch.empty(n_rays, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().esr_alpha2weight_fwd(ptr(alpha), ptr(ray_id), m, n_rays, ptr(weight), ptr(T), ptr(last),
                                              ptr(i_start), ptr(i_end), stream_ptr()))
    return [weight, T, last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    """-> grad_alpha   (render_utils_kernel.cu:679-707)"""
    for t, n in ((alpha, "alpha"), (weight, "weight"), (T, "T"), (alphainv_last, "alphainv_last"),
                 (i_start, "i_start"), (i_end, "i_end"), (grad_weights, "grad_weights"), (grad_last, "grad_last")):
        _check_input(t, n)
    grad = torch.empty_like(alpha)
    with torch.cuda.device(alpha.device):
        check(_lib.lib().esr_alpha2weight_bwd(ptr(alpha), ptr(weight), ptr(T), ptr(alphainv_last), ptr(i_start),
                                              ptr(i_end), alpha.shape[0], int(n_rays), ptr(grad_weights),
                                              ptr(grad_last), ptr(grad), stream_ptr()))
    return grad


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    """In-place grad += TV gradient   (total_variation_kernel.cu:68-98)"""
    _check_input(param, "param")
    _check_input(grad, "grad")
    with torch.cuda.device(param.device):
        check(_lib.lib().esr_tv_add_grad(ptr(param), ptr(grad), float(wx), float(wy), float(wz), param.shape[2],
                                         param.shape[3], param.shape[4], param.numel(), int(bool(dense_mode)),
                                         stream_ptr()))


render_utils_cuda = types.SimpleNamespace(sample_pts_on_rays=sample_pts_on_rays, alpha2weight=alpha2weight,
                                          alpha2weight_backward=alpha2weight_backward)
total_variation_cuda = types.SimpleNamespace(total_variation_add_grad=total_variation_add_grad)


class Alphas2Weights(torch.autograd.Function):
    """app/utils/base/module.py:117-143"""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, alphainv_last, i_start, i_end = alpha2weight(alpha, ray_id, N)
        if alpha.requires_grad:
            ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
            ctx.n_rays = N
        return weights, alphainv_last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
        grad = alpha2weight_backward(alpha, weights, T, alphainv_last, i_start, i_end, ctx.n_rays,
                                     grad_weights.contiguous(), grad_last.contiguous())
        return grad, None, None


class _SegmentSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, n_out):
        src2 = src.reshape(src.shape[0], -1).contiguous()
        out = torch.empty(n_out, src2.shape[1], dtype=src.dtype, device=src.device)
        with torch.cuda.device(src.device):
            check(_lib.lib().esr_segment_sum_fwd(ptr(src2), ptr(index), src2.shape[0], src2.shape[1], n_out, ptr(out),
                                                 stream_ptr()))
        ctx.save_for_backward(index)
        ctx.src_shape = src.shape
        return out.reshape(n_out, *src.shape[1:])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (index,) = ctx.saved_tensors
        g = grad_out.reshape(grad_out.shape[0], -1).contiguous()
        grad_src = torch.empty(index.shape[0], g.shape[1], dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            check(_lib.lib().esr_segment_sum_bwd(ptr(g), ptr(index), index.shape[0], g.shape[1], ptr(grad_src),
                                                 stream_ptr()))
        return grad_src.reshape(ctx.src_shape), None, None


def segment_coo(src, index, out=None, reduce="sum"):
    """torch_scatter.segment_coo for the one form the reference uses: sorted `index`, reduce='sum', and
    `out` a fresh zeros tensor (e.g. voxurff.py:260-265).  Returns the summed tensor."""
    if reduce != "sum" or out is None:
        raise NotImplementedError("only segment_coo(src, index, out=zeros, reduce='sum') is on the render path")
    _check_input(index, "index")
    if not src.is_cuda:
        raise RuntimeError("src must be a CUDA tensor")
    return _SegmentSum.apply(src, index, out.shape[0])

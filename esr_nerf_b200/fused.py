"""torch.autograd bindings of the fused pipeline stages in libesr_b200.so (include/esr_b200.h §2).

PyTorch is plumbing here: it owns device memory and streams and carries the (tiny) [N,3] outputs into
the caller's loss; every per-sample computation is a hand-written kernel behind the C ABI.

Stage map (reference lines in parentheses):
  march()            A/B  sample_pts_on_rays + AABB filter + MaskCache + SDF tap   (voxurff.py:186-193)
  AlphaScan          C/D  NeuS alpha, alpha filter, Alphas2Weights, weight filter  (voxurff.py:195-213)
  Shade              E+F  feature encode + off/emo radiance MLPs                   (voxurff.py:215-254)
  Tonemap            G+F  tone-map encode + tonemapper MLP                         (voxurff.py:256, 783-788)
  Composite          H    weighted per-ray sums                                    (voxurff.py:258-272)
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import MlpDesc, Scene, check, ptr, stream_ptr

# per-step work counters for bench.py's roofline block (rows through the encode / MLP kernels, summed over every
# pass of a step: primary rays, LTS points, secondary rays, eps branches); reset by the caller
STATS = {"encode_rows": 0, "mlp_fwd_rows": 0, "mlp_bwd_rows": 0, "m0": 0, "m1": 0, "m3": 0}


def reset_stats():
    for k in STATS:
        STATS[k] = 0


FEAT_DIM = 96
FEAT_GRAD_DIM = 56
TFEAT_DIM = 48
TFEAT_GRAD_DIM = 48

RADIANCE_DESC = dict(k0=96, width=192, n_hidden=3, n_out=3, act=1)   # pbr/module.py:6-21 (softplus)
TONEMAP_DESC = dict(k0=48, width=192, n_hidden=1, n_out=3, act=2)    # pbr/module.py:24-39 (sigmoid)


def make_scene(xyz_min, xyz_max, grid_size, mask_xyz_min, mask_xyz_max, mask_size, near, far, stepdist,
               voxel_size, act_shift, mask_thres, fast_thres, s_val, alpha_thres=None, fd_eps=0.0,
               sdf_tap_manual=False) -> Scene:
    sc = Scene()
    for i in range(3):
        sc.xyz_min[i] = float(xyz_min[i])
        sc.xyz_max[i] = float(xyz_max[i])
        sc.mask_xyz_min[i] = float(mask_xyz_min[i])
        sc.mask_xyz_max[i] = float(mask_xyz_max[i])
    sc.gx, sc.gy, sc.gz = (int(v) for v in grid_size)
    sc.mx, sc.my, sc.mz = (int(v) for v in mask_size)
    sc.near, sc.far = float(near), float(far)
    sc.stepdist, sc.voxel_size = float(stepdist), float(voxel_size)
    sc.act_shift, sc.mask_thres, sc.fast_thres, sc.s_val = float(act_shift), float(mask_thres), float(fast_thres), float(s_val)
    sc.alpha_thres = float(fast_thres if alpha_thres is None else alpha_thres)
    sc.fd_eps = float(fd_eps)
    sc.sdf_tap_manual = int(bool(sdf_tap_manual))
    return sc


class _GradSink:
    """Dense grid gradients are accumulated in ONE buffer per parameter per backward pass.

    Every stage that scatters into a grid gradient (alpha path, encode, analytic SDF gradient — up to a dozen calls in
    an LTS step) would otherwise return its own zero-filled dense volume (64-384 MB each at 256^3) for autograd to
    sum.  Instead a backward asks the sink for the parameter's buffer (zero-filled once, in the parameter's memory
    format), scatters into it and returns None for that input; a callback queued on the autograd engine adds the
    buffer into ``param.grad`` when the backward pass ends.  Only leaf parameters take this route (a non-leaf grid,
    e.g. the coarse stage's smoothed SDF, gets an ordinary gradient tensor); ``torch.autograd.grad`` on the grids is
    not supported by it (the reference's drivers call ``loss.backward()``, fine.py:382)."""

    def __init__(self):
        self.bufs = {}
        self.armed = False

    @staticmethod
    def eligible(t: Optional[torch.Tensor]) -> bool:
        return t is not None and t.is_leaf and t.requires_grad

    def get(self, param: torch.Tensor) -> torch.Tensor:
        key = id(param)
        if key not in self.bufs:
            self.bufs[key] = (param, torch.zeros_like(param))
            if not self.armed:
                torch.autograd.Variable._execution_engine.queue_callback(self.flush)
                self.armed = True
        return self.bufs[key][1]

    def flush(self):
        bufs, self.bufs, self.armed = self.bufs, {}, False
        for param, buf in bufs.values():
            if param.grad is None:
                param.grad = buf
            else:
                param.grad.add_(buf)


GRAD_SINK = _GradSink()

# Multi-GPU: called by Shade.backward as soon as the colour-grid gradients of the step are final (its encode backward
# is the only kernel that writes them), with {parameter: gradient buffer}.  dist.GridGradCompactor uses it to start their
# all-reduce while the alpha-path backward (k_alpha_scan_bwd, k_sdf_scatter) still runs.  None = no hook.
COLOR_GRADS_READY_HOOK = None


def _grad_target(param: Optional[torch.Tensor], like: torch.Tensor):
    """(buffer to scatter into, value to return to autograd) for a grid input of a backward"""
    if _GradSink.eligible(param):
        return GRAD_SINK.get(param), None
    g = torch.zeros_like(like)
    return g, g


def _i32(n, dev):
    return torch.empty(int(n), dtype=torch.int32, device=dev)


def _f32(*shape, dev):
    return torch.empty(*shape, dtype=torch.float32, device=dev)


def exclusive_scan(counts: torch.Tensor) -> torch.Tensor:
    """int32 [n] -> int32 [n+1] exclusive offsets, last = total."""
    L = _lib.lib()
    n = counts.shape[0]
    out = _i32(n + 1, counts.device)
    scratch = torch.empty(L.esr_scan_scratch_bytes(n), dtype=torch.uint8, device=counts.device)
    check(L.esr_exclusive_scan_i32(ptr(counts), ptr(out), n, ptr(scratch), stream_ptr()))
    return out


@dataclass
class Streams:
    """Packed sample streams of one render call (all int32 / float32, ray-slot order)."""
    n_rays: int
    ray_order: Optional[torch.Tensor]
    n_steps: torch.Tensor       # [N] candidate steps per slot (== reference N_steps)
    cnt_inbox: torch.Tensor     # [N] in-AABB candidates per slot
    off_mask: torch.Tensor      # [N+1] exclusive offsets of the M1 (post-MaskCache) stream
    m1: int
    s_ray: torch.Tensor         # [M1]
    s_step: torch.Tensor        # [M1]
    s_sdf: torch.Tensor         # [M1]
    # filled by AlphaScan
    off_shade: Optional[torch.Tensor] = None   # [N+1]
    m3: int = 0
    m3_on: int = 0
    s_alpha: Optional[torch.Tensor] = None
    s_T: Optional[torch.Tensor] = None
    h_ray: Optional[torch.Tensor] = None
    h_step: Optional[torch.Tensor] = None
    h_m1: Optional[torch.Tensor] = None
    h_sdf: Optional[torch.Tensor] = None
    aux: object = None          # result of march(between=...)
    # `neus_alpha: grad` (functions.py:45-69): set `viewdirs` ([N,3] f32, row = ray) before AlphaScan to select it;
    # AlphaScan then fills s_cos[M1] = (viewdir . grad sdf) * dist * 0.5 and computes the alphas from (s_sdf, s_cos)
    viewdirs: Optional[torch.Tensor] = None
    s_cos: Optional[torch.Tensor] = None


def march_count(sc: Scene, rays_o, rays_d, mask_density):
    """per-ray (candidate steps, in-AABB candidates, MaskCache survivors), int32 [N] each — no stream is written"""
    n = rays_o.shape[0]
    dev = rays_o.device
    n_steps, cnt_in, cnt_mask = _i32(n, dev), _i32(n, dev), _i32(n, dev)
    if n:
        check(_lib.lib().esr_march_count(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), None, n, ptr(mask_density),
                                         ptr(n_steps), ptr(cnt_in), ptr(cnt_mask), stream_ptr()))
    return n_steps, cnt_in, cnt_mask


MASK_CLASSES = True   # tests switch the class table off to compare against the exact 8-tap test on every candidate
_MASK_CLS = {}   # id(density) -> (density, version, scene key, class table)


def mask_class_table(sc: Scene, mask_density: torch.Tensor) -> torch.Tensor:
    """per-cell keep / drop / undecided classes of the MaskCache grid (esr_mask_classify), rebuilt only when the
    density tensor (identity + in-place version) or the scene constants it depends on change"""
    key = (tuple(sc.mask_xyz_min), tuple(sc.mask_xyz_max), sc.mx, sc.my, sc.mz, sc.act_shift, sc.mask_thres)
    hit = _MASK_CLS.get(id(mask_density))
    if hit is not None and hit[0] is mask_density and hit[1] == mask_density._version and hit[2] == key:
        return hit[3]
    L = _lib.lib()
    cls = torch.empty(int(L.esr_mask_class_bytes(ctypes.byref(sc))), dtype=torch.uint8, device=mask_density.device)
    check(L.esr_mask_classify(ctypes.byref(sc), ptr(mask_density), ptr(cls), stream_ptr()))
    if len(_MASK_CLS) > 16:
        _MASK_CLS.clear()
    _MASK_CLS[id(mask_density)] = (mask_density, mask_density._version, key, cls)
    return cls


def march(sc: Scene, rays_o, rays_d, ray_order, mask_density, sdf_grid, also_read=None, between=None):
    """Stages A/B.  One host read (M1) sizes the stream buffers — the reference syncs at the same
    point (render_utils_kernel.cu:212) and four more times before shading.  `also_read` (optional 0-dim integer device
    tensor, e.g. the emission-on ray count) rides along on that read: returns (streams, int(also_read)).  `between`
    (optional callable) runs on the host after the count pass has been queued and before the read blocks — work that
    only queues small kernels (weight prep) goes there, so that after a step-end synchronisation the device starts
    marching at once instead of waiting for that host code; its result is stored in streams.aux."""
    L = _lib.lib()
    dev = rays_o.device
    n = rays_o.shape[0]
    if n == 0:  # no rays (e.g. an LTS segment without points): empty streams, no launch
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        empty = Streams(0, ray_order, _i32(0, dev), _i32(0, dev), z, 0, _i32(0, dev), _i32(0, dev), _f32(0, dev=dev))
        empty.aux = between() if between is not None else None
        return empty if also_read is None else (empty, int(also_read))
    n_steps, cnt_in, cnt_mask = _i32(n, dev), _i32(n, dev), _i32(n, dev)
    st = stream_ptr()
    scp = ctypes.byref(sc)
    # keep-bit cache: one ballot word per 32 candidate steps, written by the count pass, read by the fill pass.
    # A ray's chord is at most the AABB diagonal: ceil(diag / stepdist) + 1 candidate steps.
    diag = sum((sc.xyz_max[i] - sc.xyz_min[i]) ** 2 for i in range(3)) ** 0.5
    stride = int(diag / sc.stepdist) // 32 + 2
    bits = _i32(n * stride, dev)
    cls = mask_class_table(sc, mask_density) if MASK_CLASSES else None
    check(L.esr_march_count_bits(scp, ptr(rays_o), ptr(rays_d), ptr(ray_order), n, ptr(mask_density), ptr(n_steps),
                                 ptr(cnt_in), ptr(cnt_mask), ptr(bits), stride, ptr(cls), st))
    off_mask = exclusive_scan(cnt_mask)
    aux = between() if between is not None else None
    if also_read is None:
        m1, extra = int(off_mask[n].item()), None
    else:
        m1, extra = torch.stack([off_mask[n], also_read.to(torch.int32)]).tolist()
    s_ray, s_step, s_sdf = _i32(m1, dev), _i32(m1, dev), _f32(m1, dev=dev)
    check(L.esr_march_fill_bits(scp, ptr(rays_o), ptr(rays_d), ptr(ray_order), n, ptr(mask_density), ptr(sdf_grid),
                                ptr(off_mask), ptr(s_ray), ptr(s_step), ptr(s_sdf), ptr(bits), stride, st))
    streams = Streams(n, ray_order, n_steps, cnt_in, off_mask, m1, s_ray, s_step, s_sdf)
    streams.aux = aux
    return streams if also_read is None else (streams, extra)


def neus_cos(sc: Scene, rays_o, rays_d, viewdirs, sdf_grid, s: Streams) -> torch.Tensor:
    """iter_cos[M1] of `neus_alpha: grad` (functions.py:52-54): (viewdirs[ray_id] * sample_sdf_grad's gradient).sum(-1)
    * dist * 0.5 on the M1 stream"""
    out = _f32(s.m1, dev=rays_o.device)
    check(_lib.lib().esr_neus_cos_fwd(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(viewdirs), ptr(sdf_grid), ptr(s.s_ray),
                                      ptr(s.s_step), s.m1, ptr(out), stream_ptr()))
    return out


class AlphaScan(torch.autograd.Function):
    """(h_w [M3], alphainv_last [N]) = f(sdf_grid); fills the M3 stream fields of `streams`."""

    @staticmethod
    def forward(ctx, sdf_grid, sc: Scene, rays_o, rays_d, streams: Streams, n_on):
        L = _lib.lib()
        dev = rays_o.device
        n, st, scp = streams.n_rays, stream_ptr(), ctypes.byref(sc)
        cnt_shade = _i32(n, dev)
        last = _f32(n, dev=dev)
        streams.s_alpha, streams.s_T = _f32(streams.m1, dev=dev), _f32(streams.m1, dev=dev)
        if streams.viewdirs is not None:      # neus_alpha == "grad"
            streams.viewdirs = streams.viewdirs.contiguous()
            streams.s_cos = neus_cos(sc, rays_o, rays_d, streams.viewdirs, sdf_grid, streams)
            check(L.esr_alpha_scan_count_g(scp, ptr(streams.ray_order), n, ptr(streams.off_mask), ptr(streams.s_sdf),
                                           ptr(streams.s_cos), ptr(cnt_shade), ptr(last), ptr(streams.s_alpha),
                                           ptr(streams.s_T), st))
        else:
            check(L.esr_alpha_scan_count(scp, ptr(streams.ray_order), n, ptr(streams.off_mask), ptr(streams.s_sdf),
                                         ptr(cnt_shade), ptr(last), ptr(streams.s_alpha), ptr(streams.s_T), st))
        off_shade = exclusive_scan(cnt_shade) if n else torch.zeros(1, dtype=torch.int32, device=dev)
        if n_on is None:
            m3 = int(off_shade[n].item())
            m3_on = m3
        else:
            m3, m3_on = torch.stack([off_shade[n], off_shade[int(n_on)]]).tolist()   # n_on: host int (see march)
        streams.off_shade, streams.m3, streams.m3_on = off_shade, m3, m3_on
        streams.h_ray, streams.h_step, streams.h_m1 = _i32(m3, dev), _i32(m3, dev), _i32(m3, dev)
        streams.h_sdf = _f32(m3, dev=dev)
        h_w = _f32(m3, dev=dev)
        check(L.esr_alpha_scan_fill(scp, ptr(streams.ray_order), n, ptr(streams.off_mask), ptr(streams.s_step),
                                    ptr(streams.s_sdf), ptr(off_shade), ptr(streams.s_alpha), ptr(streams.s_T),
                                    ptr(streams.h_ray), ptr(streams.h_step), ptr(streams.h_m1), ptr(h_w),
                                    ptr(streams.h_sdf), st))
        ctx.sc, ctx.streams, ctx.grid_param = sc, streams, sdf_grid
        ctx.save_for_backward(rays_o, rays_d, last, sdf_grid)
        ctx.mark_non_differentiable()
        return h_w, last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_hw, g_last):
        rays_o, rays_d, last, sdf_grid = ctx.saved_tensors
        s: Streams = ctx.streams
        dev = rays_o.device
        if s.m1 == 0:
            return None, None, None, None, None, None
        grad_sdf, ret_sdf = _grad_target(ctx.grid_param, sdf_grid)
        g_w_m1 = torch.zeros(s.m1, dtype=torch.float32, device=dev)
        if s.m3:
            g_w_m1.index_copy_(0, s.h_m1.long(), g_hw.contiguous())
        tmp_p, tmp_n = _f32(s.m1, dev=dev), _f32(s.m1, dev=dev)
        g_last = g_last.contiguous()
        if s.s_cos is not None:               # neus_alpha == "grad": tmp_p = dL/dsdf (scattered), tmp_n = dL/diter_cos
            L = _lib.lib()
            check(L.esr_alpha_scan_bwd_g(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(s.ray_order), s.n_rays,
                                         ptr(s.off_mask), ptr(s.s_ray), ptr(s.s_step), ptr(s.s_sdf), ptr(s.s_cos),
                                         ptr(s.s_alpha), ptr(s.s_T), ptr(last), ptr(g_w_m1), ptr(g_last), ptr(tmp_p),
                                         ptr(tmp_n), s.m1, ptr(grad_sdf), stream_ptr()))
            check(L.esr_neus_cos_bwd(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(s.viewdirs), ptr(s.s_ray),
                                     ptr(s.s_step), ptr(tmp_n), s.m1, ptr(grad_sdf), stream_ptr()))
            return ret_sdf, None, None, None, None, None
        check(_lib.lib().esr_alpha_scan_bwd(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(s.ray_order), s.n_rays,
                                            ptr(s.off_mask), ptr(s.s_ray), ptr(s.s_step), ptr(s.s_sdf), ptr(s.s_alpha),
                                            ptr(s.s_T), ptr(last), ptr(g_w_m1), ptr(g_last), ptr(tmp_p),
                                            ptr(tmp_n), s.m1, ptr(grad_sdf), stream_ptr()))
        return ret_sdf, None, None, None, None, None


def _desc(d: dict) -> MlpDesc:
    return MlpDesc(d["k0"], d["width"], d["n_hidden"], d["n_out"], d["act"], int(d.get("precision", 0)))


def with_precision(desc: dict, precision: int) -> dict:
    """desc with esr_mlp_desc_t::precision set: 0 = bf16 forward operands, 1 = "x2" (fp16 hi + lo pairs, three MMAs per
    product: fp32-class pre-activations / ReLU masks — the mode whose parameter gradients meet the 1e-2 tolerance)"""
    return desc if int(desc.get("precision", 0)) == int(precision) else {**desc, "precision": int(precision)}


def mlp_pack(desc: dict, flat: torch.Tensor) -> torch.Tensor:
    L = _lib.lib()
    d = _desc(desc)
    assert flat.numel() == L.esr_mlp_param_count(ctypes.byref(d)), (flat.numel(), L.esr_mlp_param_count(ctypes.byref(d)))
    image = torch.empty(L.esr_mlp_image_bytes(ctypes.byref(d)), dtype=torch.uint8, device=flat.device)
    flat_c = flat.detach().contiguous()
    check(L.esr_mlp_pack(ctypes.byref(d), ptr(flat_c), ptr(image), stream_ptr()))
    return image


def encode_features(sc: Scene, rays_o, rays_d, viewdirs, sdf_grid, off_grid, emo_grid, s: Streams, bf16: bool,
                    save_fd: bool = False, residual: bool = False):
    """feature rows of the shaded stream; save_fd=True also returns the f32 [m3,16] finite-difference gradients the
    backward then reuses instead of re-gathering the SDF taps.  residual=True (bf16 only): the buffer holds a second
    tile set behind the first — the fp16 residuals of the bf16 rounding — for the x2 forward chain."""
    # bf16 rows are written in the library's tiled layout (padded to whole 128-row tiles); f32 rows are row-major
    rows = _lib.lib().esr_mlp_act_rows(s.m3) if bf16 else s.m3
    x = torch.empty(rows * (2 if (bf16 and residual) else 1), FEAT_DIM, dtype=torch.bfloat16 if bf16 else torch.float32,
                    device=rays_o.device)
    fd = torch.empty(s.m3, 16, dtype=torch.float32, device=rays_o.device) if save_fd else None
    check(_lib.lib().esr_encode_pbr_fwd(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(viewdirs), ptr(sdf_grid),
                                        ptr(off_grid), ptr(emo_grid), None, 6, None, ptr(s.h_ray), ptr(s.h_step),
                                        ptr(s.h_sdf), s.m3, ptr(x), None, (2 if residual else 1) if bf16 else 0, ptr(fd),
                                        stream_ptr()))
    return (x, fd) if save_fd else x


def encode_backward(sc: Scene, rays_o, rays_d, sdf_grid, off_grid, emo_grid, s: Streams, d_feat, params=(None, None, None),
                    saved_fd=None):
    """params: the (sdf, off, emo) grid inputs of the calling Function — leaf parameters accumulate in GRAD_SINK"""
    (g_sdf, r_sdf), (g_off, r_off), (g_emo, r_emo) = (_grad_target(p, g) for p, g in zip(params, (sdf_grid, off_grid, emo_grid)))
    check(_lib.lib().esr_encode_pbr_bwd(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(sdf_grid), 6, None, ptr(s.h_ray),
                                        ptr(s.h_step), s.m3, ptr(d_feat), None, ptr(g_sdf), ptr(g_off), ptr(g_emo),
                                        None, ptr(saved_fd), stream_ptr()))
    return r_sdf, r_off, r_emo


def sdf_fd_gradient(sc: Scene, rays_o, rays_d, sdf_grid, s: Streams) -> torch.Tensor:
    """sample_sdf_grad (voxurff.py:670-676) at the shaded samples -> [M3,3] world-space gradient (x, y, z)."""
    g = torch.empty(s.m3, 3, dtype=torch.float32, device=rays_o.device)
    check(_lib.lib().esr_sdf_fd_gradient(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(sdf_grid), ptr(s.h_ray),
                                         ptr(s.h_step), s.m3, ptr(g), stream_ptr()))
    return g


def mlp_infer(desc, flat, x, rb, re, m_total):
    """forward only (no activations saved)"""
    y, _ = _mlp_forward(desc, mlp_pack(desc, flat), x, rb, re, m_total, False)
    return y


def _tonemap_fwd(lin, img, desc=None):
    """rgb = sigmoid(tonemapper(PE(lin))) by the fused kernel: the encoding never leaves the SM"""
    m = lin.shape[0]
    rgb = torch.empty(m, 3, dtype=torch.float32, device=lin.device)
    d = _desc(desc or TONEMAP_DESC)
    check(_lib.lib().esr_tonemap_mlp_fwd(ctypes.byref(d), ptr(img), ptr(lin), m, ptr(rgb), stream_ptr()))
    return rgb


def _tonemap_bwd(lin, img, rgb, d_rgb, d_lin_direct, desc=None):
    """(d_lin, flat weight gradient) by the fused backward kernel (hidden activations recomputed on the SM)"""
    L = _lib.lib()
    m = lin.shape[0]
    d = _desc(desc or TONEMAP_DESC)
    d_lin = torch.empty_like(lin)
    g_flat = torch.zeros(L.esr_mlp_param_count(ctypes.byref(d)), dtype=torch.float32, device=lin.device)
    # the contiguous copies are bound to names: a temporary would be released (and its block possibly re-used by the
    # next .contiguous()) before the kernel that reads it has been queued
    d_rgb_c = d_rgb.contiguous()
    d_dir_c = d_lin_direct.contiguous() if d_lin_direct is not None else None
    check(L.esr_tonemap_mlp_bwd(ctypes.byref(d), ptr(img), ptr(lin), ptr(rgb), ptr(d_rgb_c), ptr(d_dir_c), m, ptr(d_lin),
                                ptr(g_flat), stream_ptr()))
    return d_lin, g_flat


def tonemap_infer(lin, flat_tone):
    """rgb = sigmoid(tonemapper(PE(lin))) without autograd state (voxurff.py:783-788)"""
    return _tonemap_fwd(lin.contiguous(), mlp_pack(TONEMAP_DESC, flat_tone))


def composite_infer(h_w, a, b, s: Streams):
    dev = h_w.device
    a, b = a.contiguous(), b.contiguous()
    out_a, out_b = _f32(s.n_rays, 3, dev=dev), _f32(s.n_rays, 3, dev=dev)
    check(_lib.lib().esr_composite_fwd(ptr(s.ray_order), s.n_rays, ptr(s.off_shade), ptr(h_w), ptr(a), ptr(b),
                                       ptr(out_a), ptr(out_b), stream_ptr()))
    return out_a, out_b


def _check_cl(grid: torch.Tensor, name: str):
    if not grid.is_contiguous(memory_format=torch.channels_last_3d):
        raise _lib.EsrError(f"{name} must be in channels_last_3d memory format (DenseGrid.ensure_layout)")


class Encode(torch.autograd.Function):
    """fp32 feature rows [M3,96] (strict / validation path; the bf16 product path is `Shade`)."""

    @staticmethod
    def forward(ctx, sdf_grid, off_grid, emo_grid, sc, rays_o, rays_d, viewdirs, streams):
        _check_cl(off_grid, "off_color.grid")
        _check_cl(emo_grid, "emo_color.grid")
        ctx.sc, ctx.streams, ctx.grid_params = sc, streams, (sdf_grid, off_grid, emo_grid)
        ctx.save_for_backward(rays_o, rays_d, sdf_grid, off_grid, emo_grid)
        return encode_features(sc, rays_o, rays_d, viewdirs, sdf_grid, off_grid, emo_grid, streams, bf16=False)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_x):
        rays_o, rays_d, sdf_grid, off_grid, emo_grid = ctx.saved_tensors
        d_feat = d_x[:, :FEAT_GRAD_DIM].contiguous()
        g = encode_backward(ctx.sc, rays_o, rays_d, sdf_grid, off_grid, emo_grid, ctx.streams, d_feat, ctx.grid_params)
        return (*g, None, None, None, None, None)


def _mlp_forward(desc, image, x, rb, re, m_total, train, save_begin=0):
    L = _lib.lib()
    d = _desc(desc)
    dev = x.device
    # rows outside [rb, re) are defined to be zero; a full-range call writes every row
    y = (torch.empty if (rb == 0 and re == m_total) else torch.zeros)(m_total, desc["n_out"], dtype=torch.float32,
                                                                      device=dev)
    # activations (tiled bf16) + ReLU bit masks, layout private to the library
    hidden = (torch.empty(L.esr_mlp_hidden_bytes(ctypes.byref(d), m_total), dtype=torch.uint8, device=dev)
              if train else None)
    check(L.esr_mlp_fwd(ctypes.byref(d), ptr(image), ptr(x), rb, re, m_total, ptr(y), ptr(hidden), int(save_begin),
                        stream_ptr()))
    return y, hidden


def _mlp_backward(desc, image, x, y, d_y, rb, re, m_total, hidden, d_x, dx_cols, accumulate, scratch=None,
                  weights=True):
    """data gradient (+ weight gradients unless weights=False: then _mlp_backward_weights finishes the job later from the
    same scratch, which must stay untouched until it has)"""
    L = _lib.lib()
    d = _desc(desc)
    dev = x.device
    d_y = d_y.contiguous()   # bound to a name for the life of the call (see _tonemap_bwd)
    if scratch is None:
        scratch = torch.empty(L.esr_mlp_dz_bytes(ctypes.byref(d), m_total), dtype=torch.uint8, device=dev)
    d_z_out = None
    grad_flat = torch.zeros(L.esr_mlp_param_count(ctypes.byref(d)), dtype=torch.float32, device=dev) if weights else None
    check(L.esr_mlp_bwd(ctypes.byref(d), ptr(image), ptr(x), ptr(y), ptr(d_y), rb, re, m_total, ptr(hidden),
                        ptr(scratch), ptr(d_z_out), ptr(d_x), dx_cols, int(accumulate), ptr(grad_flat), stream_ptr()))
    return grad_flat, scratch


def _mlp_backward_weights(desc, x, rb, re, m_total, hidden, scratch):
    L = _lib.lib()
    d = _desc(desc)
    grad_flat = torch.zeros(L.esr_mlp_param_count(ctypes.byref(d)), dtype=torch.float32, device=x.device)
    check(L.esr_mlp_bwd_weights(ctypes.byref(d), ptr(x), rb, re, m_total, ptr(hidden), ptr(scratch), ptr(grad_flat),
                                stream_ptr()))
    return grad_flat


class Shade(torch.autograd.Function):
    """Encode + the two radiance MLPs on tensor cores (bf16 in, fp32 accumulate).

    Returns (lin_off [M3,3], lin_emo [M3,3]); lin_emo rows >= m3_on are zero (with the on-first ray
    order only emission-on rays occupy rows [0,m3_on)).  `off_grad_rows` / `emo_grad_rows` tell the
    backward which row range can carry a non-zero cotangent (voxurff.py:243-254: emission-on rays see the
    off net only through a stop-gradient)."""

    @staticmethod
    def forward(ctx, sdf_grid, off_grid, emo_grid, flat_off, flat_emo, sc, rays_o, rays_d, viewdirs, streams,
                off_grad_rows, emo_grad_rows, precision=0):
        _check_cl(off_grid, "off_color.grid")
        _check_cl(emo_grid, "emo_color.grid")
        s: Streams = streams
        train = any(ctx.needs_input_grad[:5])
        desc = with_precision(RADIANCE_DESC, precision)
        x, fd = encode_features(sc, rays_o, rays_d, viewdirs, sdf_grid, off_grid, emo_grid, s, bf16=True, save_fd=True,
                                residual=bool(precision))
        img_off, img_emo = mlp_pack(desc, flat_off), mlp_pack(desc, flat_emo)
        lin_off, hid_off = _mlp_forward(desc, img_off, x, 0, s.m3, s.m3, train, save_begin=off_grad_rows[0])
        lin_emo, hid_emo = _mlp_forward(desc, img_emo, x, 0, s.m3_on, s.m3, train)
        ctx.desc = desc
        ctx.sc, ctx.streams, ctx.grid_params = sc, s, (sdf_grid, off_grid, emo_grid)
        ctx.rows = (off_grad_rows, emo_grad_rows)
        ctx.hidden = (hid_off, hid_emo)
        ctx.save_for_backward(rays_o, rays_d, sdf_grid, off_grid, emo_grid, x, img_off, img_emo, lin_off, lin_emo, fd)
        return lin_off, lin_emo

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_off, d_emo):
        rays_o, rays_d, sdf_grid, off_grid, emo_grid, x, img_off, img_emo, lin_off, lin_emo, fd = ctx.saved_tensors
        s: Streams = ctx.streams
        hid_off, hid_emo = ctx.hidden
        (ob, oe), (eb, ee) = ctx.rows
        # with the emission-on rays first the two nets back-propagate through disjoint row ranges that together
        # cover every row: each d_x row is written exactly once (no zero-fill, no read-modify-write)
        disjoint = (eb, ee, oe) == (0, ob, s.m3)
        d_x = (torch.empty if disjoint else torch.zeros)(s.m3, FEAT_GRAD_DIM, dtype=torch.float32, device=x.device)
        acc = 0 if disjoint else 1
        # Multi-GPU with the early colour-grid exchange: data gradients of both nets first, then the encode backward and
        # the hook that starts the all-reduce of the colour volumes (12 of every 13 exchanged floats), and only then the
        # weight-gradient GEMMs of both nets — 1.6 ms of tensor-core work for the collective to hide under.  (Costs a
        # second d_z scratch: both nets' cotangents are alive at once.)
        early = COLOR_GRADS_READY_HOOK is not None
        g_off_flat, scratch = _mlp_backward(ctx.desc, img_off, x, lin_off, d_off.contiguous(), ob, oe, s.m3,
                                            hid_off, d_x, FEAT_GRAD_DIM, acc, weights=not early)
        g_emo_flat, scratch_emo = _mlp_backward(ctx.desc, img_emo, x, lin_emo, d_emo.contiguous(), eb, ee, s.m3, hid_emo,
                                                d_x, FEAT_GRAD_DIM, acc, None if early else scratch, weights=not early)
        if not early:
            ctx.hidden = None
        g_sdf, g_offc, g_emoc = encode_backward(ctx.sc, rays_o, rays_d, sdf_grid, off_grid, emo_grid, s, d_x,
                                                ctx.grid_params, fd)
        if early and g_offc is None and g_emoc is None:   # both went to the gradient sink
            p_off, p_emo = ctx.grid_params[1], ctx.grid_params[2]
            COLOR_GRADS_READY_HOOK({p_off: GRAD_SINK.get(p_off), p_emo: GRAD_SINK.get(p_emo)})
        if early:
            g_off_flat = _mlp_backward_weights(ctx.desc, x, ob, oe, s.m3, hid_off, scratch)
            g_emo_flat = _mlp_backward_weights(ctx.desc, x, eb, ee, s.m3, hid_emo, scratch_emo)
            ctx.hidden = None
        return g_sdf, g_offc, g_emoc, g_off_flat, g_emo_flat, None, None, None, None, None, None, None, None


class Tonemap(torch.autograd.Function):
    """rgb = sigmoid(tonemapper([lin, sin(lin 2^f), cos(lin 2^f)]))  (voxurff.py:783-788): fused tone-map kernels."""

    @staticmethod
    def forward(ctx, lin, flat_tone, precision=0):
        lin = lin.contiguous()
        ctx.desc = with_precision(TONEMAP_DESC, precision)
        img = mlp_pack(ctx.desc, flat_tone)
        rgb = _tonemap_fwd(lin, img, ctx.desc)
        ctx.save_for_backward(lin, img, rgb)
        return rgb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_rgb):
        lin, img, rgb = ctx.saved_tensors
        return (*_tonemap_bwd(lin, img, rgb, d_rgb, None, ctx.desc), None)


class CombineTonemap(torch.autograd.Function):
    """(rgb, lin) from the two radiances: lin = lin_off (+ lin_emo on emission-on rays, which see the off net
    through a stop-gradient, voxurff.py:243-254); rgb = sigmoid(tonemapper(PE(lin))) (voxurff.py:783-788).
    One encode kernel + the tensor-core MLP; replaces torch.where / add / copies around `Tonemap`."""

    @staticmethod
    def forward(ctx, lin_off, lin_emo, flat_tone, h_ray, em_modes, ordered, off_sees_on=False, precision=0):
        L = _lib.lib()
        ctx.desc = with_precision(TONEMAP_DESC, precision)
        m = lin_off.shape[0]
        lin_off, lin_emo = lin_off.contiguous(), lin_emo.contiguous()
        lin = torch.empty_like(lin_off)
        # combine only (tfeat = NULL): the fused tone-map kernel encodes lin itself
        check(L.esr_tonemap_encode_fwd(ptr(lin_off), ptr(lin_emo), ptr(h_ray), ptr(em_modes), m, ptr(lin), None, 1,
                                       stream_ptr()))
        img = mlp_pack(ctx.desc, flat_tone)
        rgb = _tonemap_fwd(lin, img, ctx.desc)
        # off_sees_on: the off net receives the cotangent of emission-on rows too (ESRNeRF adds the two radiances
        # without a stop-gradient, esrnerf.py:751-757; VoxurfF detaches, voxurff.py:243-254)
        ctx.ordered, ctx.off_sees_on = ordered, off_sees_on
        ctx.save_for_backward(lin, img, rgb, h_ray, em_modes)
        return rgb, lin

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_rgb, d_lin_direct):
        lin, img, rgb, h_ray, em_modes = ctx.saved_tensors
        d_lin, g_flat = _tonemap_bwd(lin, img, rgb, d_rgb, d_lin_direct, ctx.desc)
        if ctx.ordered:
            # emission-on rows are a prefix and each net back-propagates through its own row range only (Shade.backward):
            # both can read the same cotangent
            return d_lin, d_lin, g_flat, None, None, None, None, None
        on = (em_modes[h_ray.long()] == 1)[:, None]
        zero = torch.zeros_like(d_lin)
        d_off = d_lin if ctx.off_sees_on else torch.where(on, zero, d_lin)
        return d_off, torch.where(on, d_lin, zero), g_flat, None, None, None, None, None


class Composite(torch.autograd.Function):
    """(sum_ray w*a, sum_ray w*b): warp-per-ray segmented sums replacing segment_coo (voxurff.py:259-272)."""

    @staticmethod
    def forward(ctx, h_w, a, b, streams):
        s: Streams = streams
        dev = h_w.device
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        out_a = _f32(s.n_rays, 3, dev=dev)
        out_b = _f32(s.n_rays, 3, dev=dev) if b is not None else None
        check(_lib.lib().esr_composite_fwd(ptr(s.ray_order), s.n_rays, ptr(s.off_shade), ptr(h_w), ptr(a), ptr(b),
                                           ptr(out_a), ptr(out_b), stream_ptr()))
        ctx.streams, ctx.has_b = s, b is not None
        ctx.save_for_backward(h_w, a, *([b] if b is not None else []))
        return (out_a, out_b) if b is not None else (out_a, out_a.new_zeros(0))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, c_a, c_b):
        h_w, a = ctx.saved_tensors[:2]
        b = ctx.saved_tensors[2] if ctx.has_b else None
        s: Streams = ctx.streams
        d_a, g_w = torch.empty_like(a), torch.empty_like(h_w)
        d_b = torch.empty_like(b) if b is not None else None
        c_a = c_a.contiguous()          # named: the copies must outlive the launch (a temporary's block could be handed
        c_b = c_b.contiguous() if b is not None else None   # to the next .contiguous() before the kernel is queued)
        check(_lib.lib().esr_composite_bwd(ptr(s.h_ray), None, ptr(h_w), ptr(a), ptr(b), ptr(c_a), ptr(c_b), s.m3,
                                           ptr(d_a), ptr(d_b), ptr(g_w), stream_ptr()))
        return g_w, d_a, d_b, None


# ---------------------------------------------------------------------------------------------------
# coarse stage (VoxurfC, voxurfc.py:186-271)
# ---------------------------------------------------------------------------------------------------
COARSE_FEAT_DIM = 72


class CoarseAlpha(torch.autograd.Function):
    """NeuS alpha of the samples that survive the first Alphas2Weights pass (weights > thres, voxurfc.py:211-218):
    h_alpha [M3] = f(smoothed sdf grid).  Fills the M3 stream fields of `streams`.  The reference then recomputes
    the weights on the survivors (voxurfc.py:219) — done by the caller with the reference-shaped Alphas2Weights."""

    @staticmethod
    def forward(ctx, sdf_grid, sc: Scene, rays_o, rays_d, streams: Streams, grad_vol=None, viewdirs=None):
        """grad_vol + viewdirs given = `neus_alpha: grad` (voxurfc.py:171-174): the section-point SDFs from the view-projected
        trilinear tap of the central-difference volume grad_vol [1,3,X,Y,Z] (voxurfc.py:204-210)"""
        L = _lib.lib()
        dev = rays_o.device
        n, st, scp = streams.n_rays, stream_ptr(), ctypes.byref(sc)
        cnt_shade = _i32(n, dev)
        last = _f32(n, dev=dev)
        streams.s_alpha, streams.s_T = _f32(streams.m1, dev=dev), _f32(streams.m1, dev=dev)
        if grad_vol is not None:
            grad_vol, streams.viewdirs = grad_vol.contiguous(), viewdirs.contiguous()
            streams.s_cos = _f32(streams.m1, dev=dev)
            check(L.esr_neus_cos_vol_fwd(scp, ptr(rays_o), ptr(rays_d), ptr(streams.viewdirs), ptr(grad_vol),
                                         ptr(streams.s_ray), ptr(streams.s_step), streams.m1, ptr(streams.s_cos), st))
            check(L.esr_alpha_scan_count_g(scp, ptr(streams.ray_order), n, ptr(streams.off_mask), ptr(streams.s_sdf),
                                           ptr(streams.s_cos), ptr(cnt_shade), ptr(last), ptr(streams.s_alpha),
                                           ptr(streams.s_T), st))
        else:
            check(L.esr_alpha_scan_count(scp, ptr(streams.ray_order), n, ptr(streams.off_mask), ptr(streams.s_sdf),
                                         ptr(cnt_shade), ptr(last), ptr(streams.s_alpha), ptr(streams.s_T), st))
        ctx.grad_mode = grad_vol is not None
        off_shade = exclusive_scan(cnt_shade)
        m3 = int(off_shade[n].item())
        streams.off_shade, streams.m3, streams.m3_on = off_shade, m3, m3
        streams.h_ray, streams.h_step, streams.h_m1 = _i32(m3, dev), _i32(m3, dev), _i32(m3, dev)
        streams.h_sdf = _f32(m3, dev=dev)
        h_w = _f32(m3, dev=dev)
        check(L.esr_alpha_scan_fill(scp, ptr(streams.ray_order), n, ptr(streams.off_mask), ptr(streams.s_step),
                                    ptr(streams.s_sdf), ptr(off_shade), ptr(streams.s_alpha), ptr(streams.s_T),
                                    ptr(streams.h_ray), ptr(streams.h_step), ptr(streams.h_m1), ptr(h_w),
                                    ptr(streams.h_sdf), st))
        h_alpha = streams.s_alpha[streams.h_m1.long()] if m3 else _f32(0, dev=dev)
        ctx.sc, ctx.streams = sc, streams
        ctx.vol_shape = None if grad_vol is None else tuple(grad_vol.shape)
        ctx.save_for_backward(rays_o, rays_d, sdf_grid)
        return h_alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_alpha):
        rays_o, rays_d, sdf_grid = ctx.saved_tensors
        s: Streams = ctx.streams
        dev = rays_o.device
        grad_sdf = torch.zeros_like(sdf_grid)
        g_vol = torch.zeros(ctx.vol_shape, dtype=torch.float32, device=dev) if ctx.grad_mode else None
        if s.m1 == 0 or s.m3 == 0:
            return grad_sdf, None, None, None, None, g_vol, None
        g_m1 = torch.zeros(s.m1, dtype=torch.float32, device=dev)
        g_m1.index_copy_(0, s.h_m1.long(), g_alpha.contiguous())
        tmp_p, tmp_n = _f32(s.m1, dev=dev), _f32(s.m1, dev=dev)
        if ctx.grad_mode:                     # tmp_p = dL/dsdf (scattered into grad_sdf), tmp_n = dL/diter_cos -> g_vol
            L = _lib.lib()
            check(L.esr_neus_alpha_bwd_g(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(s.ray_order), s.n_rays,
                                         ptr(s.off_mask), ptr(s.s_ray), ptr(s.s_step), ptr(s.s_sdf), ptr(s.s_cos), ptr(g_m1),
                                         ptr(tmp_p), ptr(tmp_n), s.m1, ptr(grad_sdf), stream_ptr()))
            check(L.esr_neus_cos_vol_bwd(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(s.viewdirs), ptr(s.s_ray),
                                         ptr(s.s_step), ptr(tmp_n), s.m1, ptr(g_vol), stream_ptr()))
            return grad_sdf, None, None, None, None, g_vol, None
        check(_lib.lib().esr_neus_alpha_bwd(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(s.ray_order), s.n_rays,
                                            ptr(s.off_mask), ptr(s.s_ray), ptr(s.s_step), ptr(s.s_sdf), ptr(g_m1),
                                            ptr(tmp_p), ptr(tmp_n), s.m1, ptr(grad_sdf), stream_ptr()))
        return grad_sdf, None, None, None, None, None, None


class EncodeCoarse(torch.autograd.Function):
    """[M3,72] f32 coarse feature rows (esr_encode_coarse_fwd): colour taps, PE, normal from the gradient volume."""

    @staticmethod
    def forward(ctx, grad_vol, off_grid, emo_grid, sc, rays_o, rays_d, viewdirs, streams):
        _check_cl(off_grid, "off_color.grid")
        _check_cl(emo_grid, "emo_color.grid")
        s: Streams = streams
        grad_vol = grad_vol.contiguous()
        x = torch.empty(s.m3, COARSE_FEAT_DIM, dtype=torch.float32, device=rays_o.device)
        check(_lib.lib().esr_encode_coarse_fwd(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(viewdirs), ptr(grad_vol),
                                               ptr(off_grid), ptr(emo_grid), ptr(s.h_ray), ptr(s.h_step), s.m3, ptr(x),
                                               stream_ptr()))
        ctx.sc, ctx.streams = sc, s
        ctx.save_for_backward(rays_o, rays_d, grad_vol, off_grid, emo_grid)
        return x

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_x):
        rays_o, rays_d, grad_vol, off_grid, emo_grid = ctx.saved_tensors
        s: Streams = ctx.streams
        g_vol, g_off, g_emo = torch.zeros_like(grad_vol), torch.zeros_like(off_grid), torch.zeros_like(emo_grid)
        d_x = d_x.contiguous()
        check(_lib.lib().esr_encode_coarse_bwd(ctypes.byref(ctx.sc), ptr(rays_o), ptr(rays_d), ptr(grad_vol),
                                               ptr(s.h_ray), ptr(s.h_step), s.m3, ptr(d_x), ptr(g_vol),
                                               ptr(g_off), ptr(g_emo), stream_ptr()))
        return g_vol, g_off, g_emo, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------
# LTS / PDRA stage (ESRNeRF, esrnerf.py:487-851)
# ---------------------------------------------------------------------------------------------------
EMIT_DESC = dict(k0=96, width=192, n_hidden=3, n_out=3, act=1)   # EmissionNet 76->128x3->3 softplus, zero-padded to 192
BRDF_DESC = dict(k0=96, width=192, n_hidden=3, n_out=5, act=2)   # BRDFNet 76->128x3->5 sigmoid, zero-padded to 192


def sample_points(sc: Scene, rays_o, rays_d, h_ray, h_step) -> torch.Tensor:
    """ray_pts of the reference for stream samples -> [m,3] (bit-identical to kernel.cu:167-194)"""
    m = h_ray.shape[0]
    pts = _f32(m, 3, dev=rays_o.device)
    check(_lib.lib().esr_sample_points(ctypes.byref(sc), ptr(rays_o), ptr(rays_d), ptr(h_ray), ptr(h_step), m, ptr(pts),
                                       stream_ptr()))
    return pts


def sdf_tap_points(sc: Scene, sdf_grid, pts) -> torch.Tensor:
    """F.grid_sample arithmetic at explicit points (sample_sdf_grad's value, esrnerf.py:819), no autograd: the
    gradient of this value is scattered by the encode backward through the row's sdf column."""
    m = pts.shape[0]
    out = _f32(m, dev=pts.device)
    check(_lib.lib().esr_sdf_expgrad_fwd(ctypes.byref(sc), ptr(pts), ptr(sdf_grid), m, 0, ptr(out), None, stream_ptr()))
    return out


def sdf_expgrad_points(sc: Scene, sdf_grid, pts):
    """sample_sdf_expgrad (esrnerf.py:1572-1596) without autograd state -> (sdf [m], d sdf / d xyz [m,3])"""
    m = pts.shape[0]
    pts = pts.contiguous()
    sdf, grad = _f32(m, dev=pts.device), _f32(m, 3, dev=pts.device)
    check(_lib.lib().esr_sdf_expgrad_fwd(ctypes.byref(sc), ptr(pts), ptr(sdf_grid), m, 1, ptr(sdf), ptr(grad),
                                         stream_ptr()))
    return sdf, grad


class SdfExpGrad(torch.autograd.Function):
    """sample_sdf_expgrad (esrnerf.py:1572-1596): analytic d sdf / d xyz [m,3] at explicit points, differentiable
    w.r.t. the grid (the reference's create_graph=True path)."""

    @staticmethod
    def forward(ctx, sdf_grid, sc, pts):
        m = pts.shape[0]
        pts = pts.contiguous()
        grad = _f32(m, 3, dev=pts.device)
        check(_lib.lib().esr_sdf_expgrad_fwd(ctypes.byref(sc), ptr(pts), ptr(sdf_grid), m, 1, None, ptr(grad),
                                             stream_ptr()))
        ctx.sc, ctx.grid_param = sc, sdf_grid
        ctx.save_for_backward(pts, sdf_grid)
        return grad

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_grad):
        pts, sdf_grid = ctx.saved_tensors
        if pts.shape[0] == 0:
            return None, None, None
        g, ret = _grad_target(ctx.grid_param, sdf_grid)
        g_grad = g_grad.contiguous()
        check(_lib.lib().esr_sdf_expgrad_bwd(ctypes.byref(ctx.sc), ptr(pts), pts.shape[0], None,
                                             ptr(g_grad), ptr(g), stream_ptr()))
        return ret, None, None


@dataclass
class SamplePos:
    """where the rows of one encode call sit: stream samples (rays_o/rays_d/h_ray/h_step) or explicit points (pts;
    h_ray optional: row j then reads view direction j)"""
    m: int
    viewdirs: torch.Tensor
    h_sdf: torch.Tensor
    rays_o: Optional[torch.Tensor] = None
    rays_d: Optional[torch.Tensor] = None
    h_ray: Optional[torch.Tensor] = None
    h_step: Optional[torch.Tensor] = None
    pts: Optional[torch.Tensor] = None


class ShadePBR(torch.autograd.Function):
    """Encode + any subset of the four nets of the LTS / PDRA stage on tensor cores:
        lin_off = softplus(off_rgbnet([off_color | feat]))          esrnerf.py:755-757
        lin_emo = softplus(emo_rgbnet([emo_color | feat]))          esrnerf.py:752-754 (caller masks by em_modes)
        emit    = softplus(emitnet([emo_color | brdf_feat]))        esrnerf.py:765
        brdf    = sigmoid(brdfnet([brdf_grid | brdf_feat])) [m,5]   esrnerf.py:761-764
    `use` = (off, emo, emit, brdf) booleans; unused outputs are empty tensors.  The 76->128 nets run in the 96->192
    kernels with zero-padded weights (view columns and hidden units 128..191 have zero weights)."""

    @staticmethod
    def forward(ctx, sdf_grid, off_grid, emo_grid, brdf_grid, flat_off, flat_emo, flat_emit, flat_brdf, sc, pos, use,
                precision=0):
        _check_cl(off_grid, "off_color.grid")
        _check_cl(emo_grid, "emo_color.grid")
        p: SamplePos = pos
        L = _lib.lib()
        dev = sdf_grid.device
        m = p.m
        train = any(ctx.needs_input_grad[:8])
        precision = int(precision) if train else 0   # the x2 forward exists for the backward's sake (exact ReLU masks)
        rows = L.esr_mlp_act_rows(m) * (2 if precision else 1)   # x2: the fp16 residual tiles follow the bf16 tiles
        x = torch.empty(rows, FEAT_DIM, dtype=torch.bfloat16, device=dev)
        x2 = None
        if use[3]:
            _check_cl(brdf_grid, "brdf.grid")
            x2 = torch.empty(rows, FEAT_DIM, dtype=torch.bfloat16, device=dev)
        fd = torch.empty(m, 16, dtype=torch.float32, device=dev) if train else None
        check(L.esr_encode_pbr_fwd(ctypes.byref(sc), ptr(p.rays_o), ptr(p.rays_d), ptr(p.viewdirs), ptr(sdf_grid),
                                   ptr(off_grid), ptr(emo_grid), ptr(brdf_grid) if use[3] else None, 6, ptr(p.pts),
                                   ptr(p.h_ray), ptr(p.h_step), ptr(p.h_sdf), m, ptr(x), ptr(x2), 2 if precision else 1,
                                   ptr(fd), stream_ptr()))
        ctx.fd = fd
        descs = tuple(with_precision(d, precision) for d in (RADIANCE_DESC, RADIANCE_DESC, EMIT_DESC, BRDF_DESC))
        ctx.descs = descs
        flats = (flat_off, flat_emo, flat_emit, flat_brdf)
        STATS["encode_rows"] += m
        STATS["mlp_fwd_rows"] += m * sum(use)
        STATS["mlp_bwd_rows"] += m * sum(use) if train else 0
        outs, imgs, hids = [], [], []
        for k in range(4):
            if not use[k]:
                outs.append(torch.zeros(0, descs[k]["n_out"], device=dev))
                imgs.append(None)
                hids.append(None)
                continue
            img = mlp_pack(descs[k], flats[k])
            y, hid = _mlp_forward(descs[k], img, x2 if k == 3 else x, 0, m, m, train)
            outs.append(y)
            imgs.append(img)
            hids.append(hid)
        ctx.sc, ctx.pos, ctx.use, ctx.hids = sc, p, use, hids
        ctx.grid_params = (sdf_grid, off_grid, emo_grid, brdf_grid)
        ctx.n_saved = [t is not None for t in imgs]
        ctx.save_for_backward(sdf_grid, off_grid, emo_grid, brdf_grid if use[3] else sdf_grid.new_zeros(0), x,
                              x2 if x2 is not None else x.new_zeros(0), *[t for t in imgs if t is not None], *outs)
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_off, d_emo, d_emit, d_brdf):
        saved = ctx.saved_tensors
        sdf_grid, off_grid, emo_grid, brdf_grid, x, x2 = saved[:6]
        n_img = sum(ctx.n_saved)
        img_list = list(saved[6:6 + n_img])
        outs = saved[6 + n_img:]
        p: SamplePos = ctx.pos
        use, m = ctx.use, p.m
        dev = x.device
        descs = ctx.descs
        d_ys = (d_off, d_emo, d_emit, d_brdf)
        imgs = [img_list.pop(0) if has else None for has in ctx.n_saved]
        d_x = torch.zeros(m, FEAT_GRAD_DIM, dtype=torch.float32, device=dev)
        g_flat = [None, None, None, None]
        d_brdf_c, scratch = None, None
        for k in (3, 0, 1, 2):   # the BRDF net first: its colour-slot cotangent is split off before the others accumulate
            if not use[k]:
                continue
            g_flat[k], scratch = _mlp_backward(descs[k], imgs[k], x2 if k == 3 else x, outs[k], d_ys[k].contiguous(), 0,
                                               m, m, ctx.hids[k], d_x, FEAT_GRAD_DIM, 1, scratch)
            if k == 3:
                d_brdf_c = d_x[:, :6].contiguous()
                d_x[:, :6] = 0
        ctx.hids = None
        if m == 0:
            return (None, None, None, None, *g_flat, None, None, None, None)
        needs = ctx.needs_input_grad
        gp = ctx.grid_params
        g_sdf, r_sdf = _grad_target(gp[0], sdf_grid) if needs[0] else (torch.zeros_like(sdf_grid), None)
        g_off, r_off = _grad_target(gp[1], off_grid) if (use[0] and needs[1]) else (None, None)
        g_emo, r_emo = _grad_target(gp[2], emo_grid) if ((use[1] or use[2]) and needs[2]) else (None, None)
        g_brdf, r_brdf = _grad_target(gp[3], brdf_grid) if (use[3] and needs[3]) else (None, None)
        if g_brdf is None:
            d_brdf_c = None
        check(_lib.lib().esr_encode_pbr_bwd(ctypes.byref(ctx.sc), ptr(p.rays_o), ptr(p.rays_d), ptr(sdf_grid), 6,
                                            ptr(p.pts), ptr(p.h_ray), ptr(p.h_step), m, ptr(d_x), ptr(d_brdf_c),
                                            ptr(g_sdf), ptr(g_off), ptr(g_emo), ptr(g_brdf), ptr(ctx.fd), stream_ptr()))
        ctx.fd = None
        return (r_sdf, r_off, r_emo, r_brdf, *g_flat, None, None, None, None)


class LtsAccumulate(torch.autograd.Function):
    """(off_hat, reflect) [2P,3] of the light-transport segment (esrnerf.py:556-574, 654-677): Disney-style reflectance
    of every (point, direction, outgoing direction) triple times the marched radiance of the secondary rays, averaged
    over the directions — one kernel forward, one backward (esr_lts_accumulate_*).  rad_off may be None (finetune)."""

    @staticmethod
    def forward(ctx, base, rough, metal, rad_off, rad_emo, normal, wo_a, wo_b, dirs, n_dirs, emission=None, umask=None,
                pdra_mode=False):
        """emission (optional, [P,3]): the second output is emo_hat of esrnerf.py:668-677 instead of reflect — emission +
        reflect, or in pdra_mode emission + stop-gradient(reflect) on uncertain rays' points (umask) and reflect elsewhere"""
        P = base.shape[0]
        dev = base.device
        t = [x.contiguous().float() if x is not None else None
             for x in (base, rough, metal, rad_off, rad_emo, normal, wo_a, wo_b, dirs, emission)]
        base, rough, metal, rad_off, rad_emo, normal, wo_a, wo_b, dirs, emission = t
        um8 = umask.to(torch.uint8).contiguous() if (umask is not None and emission is not None) else None
        off_hat = _f32(2 * P, 3, dev=dev) if rad_off is not None else None
        reflect = _f32(2 * P, 3, dev=dev)
        check(_lib.lib().esr_lts_accumulate_fwd(ptr(normal), ptr(base), ptr(rough), ptr(metal), ptr(wo_a), ptr(wo_b), ptr(dirs),
                                                ptr(rad_off), ptr(rad_emo), P, int(n_dirs), ptr(off_hat), ptr(reflect),
                                                ptr(emission), ptr(um8), int(bool(pdra_mode)), stream_ptr()))
        ctx.n_dirs, ctx.has_off, ctx.pdra = int(n_dirs), rad_off is not None, bool(pdra_mode)
        ctx.emission, ctx.um8 = emission, um8
        ctx.save_for_backward(base, rough, metal, rad_emo, normal, wo_a, wo_b, dirs, *([rad_off] if rad_off is not None else []))
        return (off_hat if off_hat is not None else reflect.new_zeros(0, 3)), reflect

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_off_hat, g_reflect):
        base, rough, metal, rad_emo, normal, wo_a, wo_b, dirs = ctx.saved_tensors[:8]
        rad_off = ctx.saved_tensors[8] if ctx.has_off else None
        P = base.shape[0]
        dev = base.device
        g_base, g_rough, g_metal = _f32(P, 3, dev=dev), torch.empty_like(rough), torch.empty_like(metal)
        g_rad_emo = torch.empty_like(rad_emo)
        g_rad_off = torch.empty_like(rad_off) if rad_off is not None else None
        g_emission = _f32(P, 3, dev=dev) if ctx.emission is not None else None
        if P:
            g_off_hat = g_off_hat.contiguous() if rad_off is not None else None   # named: must outlive the launch
            g_reflect = g_reflect.contiguous()
            check(_lib.lib().esr_lts_accumulate_bwd(ptr(normal), ptr(base), ptr(rough), ptr(metal), ptr(wo_a), ptr(wo_b),
                                                    ptr(dirs), ptr(rad_off), ptr(rad_emo), P, ctx.n_dirs,
                                                    ptr(g_off_hat), ptr(g_reflect), ptr(g_base), ptr(g_rough), ptr(g_metal),
                                                    ptr(g_rad_off), ptr(g_rad_emo), ptr(ctx.emission), ptr(ctx.um8),
                                                    int(ctx.pdra), ptr(g_emission), stream_ptr()))
        return g_base, g_rough, g_metal, g_rad_off, g_rad_emo, None, None, None, None, None, g_emission, None, None


SG_ACT_IDS = {"softplus": 1, "relu": 2, "abs": 3, "exp": 4, "sigmoid": 5}


class SgEnvmap(torch.autograd.Function):
    """add + act(sum_k mus_k exp(lambdas_k (d . lobes_k - 1))) * scale over the secondary rays (pbr/module.py:133-143 with
    its use at esrnerf.py:560-566 fused in): one kernel forward, one backward.  `lambdas` non-negative [K], `lobes` unit
    [K,3] — the caller's abs / F.normalize stay in torch, autograd carries their (tiny) backward."""

    @staticmethod
    def forward(ctx, dirs, mus, lambdas, lobes, act_id, scale, add):
        dirs, mus, lambdas, lobes = (t.contiguous().float() for t in (dirs, mus, lambdas, lobes))
        scale = scale.contiguous().float() if scale is not None else None
        add = add.contiguous().float() if add is not None else None
        m = dirs.shape[0]
        out = _f32(m, 3, dev=dirs.device)
        check(_lib.lib().esr_sg_envmap_fwd(ptr(dirs), ptr(mus), ptr(lambdas), ptr(lobes), mus.shape[0], int(act_id), ptr(scale),
                                           ptr(add), m, ptr(out), stream_ptr()))
        ctx.act_id, ctx.has_add, ctx.has_scale = int(act_id), add is not None, scale is not None
        ctx.save_for_backward(dirs, mus, lambdas, lobes, *([scale] if scale is not None else []))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        dirs, mus, lambdas, lobes = ctx.saved_tensors[:4]
        scale = ctx.saved_tensors[4] if ctx.has_scale else None
        g_out = g_out.contiguous()
        g_mus, g_lam, g_lobes = torch.zeros_like(mus), torch.zeros_like(lambdas), torch.zeros_like(lobes)
        g_scale = torch.empty_like(scale) if scale is not None else None
        check(_lib.lib().esr_sg_envmap_bwd(ptr(dirs), ptr(mus), ptr(lambdas), ptr(lobes), mus.shape[0], ctx.act_id, ptr(scale),
                                           dirs.shape[0], ptr(g_out), ptr(g_mus), ptr(g_lam), ptr(g_lobes), ptr(g_scale),
                                           stream_ptr()))
        return None, g_mus, g_lam, g_lobes, None, g_scale, (g_out if ctx.has_add else None)


@torch.no_grad()
def lts_scatter_dirs(normal, number: int, noise=None, table=None):
    """hemisphere directions [P, number, 3] of the LTS points (pbr/functions.py:10-32): one kernel (esr_lts_scatter_dirs)"""
    P = normal.shape[0]
    normal = normal.contiguous().float()
    noise = noise.contiguous().float() if noise is not None else None
    table = table.contiguous().float() if table is not None else None
    dirs = _f32(P, number, 3, dev=normal.device)
    check(_lib.lib().esr_lts_scatter_dirs(ptr(normal), ptr(noise), ptr(table), P, int(number), ptr(dirs), stream_ptr()))
    return dirs


# ---------------------------------------------------------------------------------------------------
# dense-grid loss terms of the stage drivers: csrc/regularizers.cu
# ---------------------------------------------------------------------------------------------------
def _mask_u8(mask, shape):
    if mask is None:
        return None
    m = mask.reshape(-1)[: shape[0] * shape[1] * shape[2]] if mask.numel() != shape[0] * shape[1] * shape[2] else mask
    return m.reshape(shape).contiguous().view(torch.uint8) if m.dtype == torch.bool else m.reshape(shape).to(torch.uint8).contiguous()


class GridTV(torch.autograd.Function):
    """total_variation(v, mask) of app/utils/base/functions.py:34-42 on a [1,C,X,Y,Z] grid in whatever memory layout it has
    (contiguous SDF grid, channels-last colour grid): mean |forward difference| per axis over the pairs with both voxels in
    `mask` ([1,1,X,Y,Z] bool, shared by the channels), averaged over the axes.  One launch forward, one backward."""

    @staticmethod
    def forward(ctx, v, mask):
        C, X, Y, Z = v.shape[1:]
        m8 = _mask_u8(mask, (X, Y, Z))
        acc = torch.empty(6, dtype=torch.float64, device=v.device)
        st = v.stride()[1:]
        check(_lib.lib().esr_grid_tv_fwd(ptr(v), ptr(m8), C, X, Y, Z, *st, ptr(acc), stream_ptr()))
        ctx.save_for_backward(v, acc, *([m8] if m8 is not None else []))
        return ((acc[0] / acc[3] + acc[1] / acc[4] + acc[2] / acc[5]) / 3).float()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        v, acc = ctx.saved_tensors[:2]
        m8 = ctx.saved_tensors[2] if len(ctx.saved_tensors) > 2 else None
        C, X, Y, Z = v.shape[1:]
        grad = torch.zeros_like(v)            # (preserves the grid's memory layout)
        g = g.reshape(1).float().contiguous()
        check(_lib.lib().esr_grid_tv_bwd(ptr(v), ptr(m8), C, X, Y, Z, *v.stride()[1:], ptr(acc), ptr(g), 1.0, ptr(grad),
                                         stream_ptr()))
        return grad, None


class SmoothGradTV(torch.autograd.Function):
    """mean over the masked voxels and the three components of (conv3(grad) - grad)^2, grad = neus_sdf_gradient()
    (voxurff.py:610-616, 723-742; fixed 3x3x3 kernel of GradientConv, replicate padding, the smoothed volume detached)."""

    @staticmethod
    def forward(ctx, sdf_grid, mask, w27, bias, voxel_size):
        X, Y, Z = sdf_grid.shape[2:]
        sdf = sdf_grid.contiguous()
        m8 = _mask_u8(mask, (X, Y, Z))
        L = _lib.lib()
        gvol = torch.empty(3, X, Y, Z, dtype=torch.float32, device=sdf.device)
        check(L.esr_sdf_central_gradient(ptr(sdf), X, Y, Z, float(voxel_size), ptr(gvol), stream_ptr()))
        err = torch.empty_like(gvol)
        acc = torch.empty(2, dtype=torch.float64, device=sdf.device)
        w27 = w27.reshape(-1).float().contiguous()
        check(L.esr_smooth_grad_tv_fwd(ptr(gvol), ptr(m8), X, Y, Z, ptr(w27), float(bias), ptr(acc), ptr(err), stream_ptr()))
        ctx.voxel_size = float(voxel_size)
        ctx.save_for_backward(err, acc)
        ctx.shape = tuple(sdf_grid.shape)
        return (acc[0] / acc[1]).float()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        err, acc = ctx.saved_tensors
        X, Y, Z = ctx.shape[2:]
        grad = torch.zeros(ctx.shape, dtype=torch.float32, device=err.device)
        g = g.reshape(1).float().contiguous()
        check(_lib.lib().esr_smooth_grad_tv_bwd(ptr(err), X, Y, Z, ctx.voxel_size, ptr(acc), ptr(g), 1.0, ptr(grad), stream_ptr()))
        return grad, None, None, None, None


class SdfCentralGradient(torch.autograd.Function):
    """neus_sdf_gradient (voxurff.py:723-742): [1,3,X,Y,Z] central differences / 2 / voxel_size, zero on the boundary faces"""

    @staticmethod
    def forward(ctx, sdf_grid, voxel_size):
        X, Y, Z = sdf_grid.shape[2:]
        sdf = sdf_grid.contiguous()
        out = torch.empty(1, 3, X, Y, Z, dtype=torch.float32, device=sdf.device)
        check(_lib.lib().esr_sdf_central_gradient(ptr(sdf), X, Y, Z, float(voxel_size), ptr(out), stream_ptr()))
        ctx.voxel_size, ctx.shape = float(voxel_size), tuple(sdf_grid.shape)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        X, Y, Z = ctx.shape[2:]
        g = g.contiguous().float()
        grad = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        check(_lib.lib().esr_smooth_grad_tv_bwd(ptr(g), X, Y, Z, ctx.voxel_size, None, None, 1.0, ptr(grad), stream_ptr()))
        return grad, None


# ---------------------------------------------------------------------------------------------------
# coarse stage: the two 57 -> 128 -> 128 -> 3 sigmoid colour nets (voxurfc.py:137-169, 229-249) on the tcgen05 chains
# ---------------------------------------------------------------------------------------------------
COARSE_DESC = dict(k0=96, width=192, n_hidden=3, n_out=3, act=2)      # zero-padded + identity third hidden layer (modules)
COARSE_GRAD_COLS = 16                                                 # colour taps (12) + normal (3), padded to 16


def rows_to_tiles(x: torch.Tensor, colmap: torch.Tensor, precision: int) -> torch.Tensor:
    """f32 rows [m, ld] -> tiled 16-bit MLP input rows (esr_rows_to_mlp_tiles)"""
    L = _lib.lib()
    m = x.shape[0]
    rows = L.esr_mlp_act_rows(m) * (2 if precision else 1)
    tiles = torch.empty(rows, FEAT_DIM, dtype=torch.bfloat16, device=x.device)
    check(L.esr_rows_to_mlp_tiles(ptr(x), m, x.shape[1], ptr(colmap), int(precision), ptr(tiles), stream_ptr()))
    return tiles


class CoarseShade(torch.autograd.Function):
    """(sigmoid(off_rgbnet([off colour | feat])), sigmoid(emo_rgbnet([emo colour | feat]))) of voxurfc.py:229-240 on the
    tensor-core chains: x is the f32 [M3,72] row of EncodeCoarse; each net reads its own tiled copy of the row (its colour
    slot first).  Backward: data gradient of the 15 gradient-carrying columns (colour taps, normal) back into x's layout,
    weight gradients of both nets."""

    @staticmethod
    def forward(ctx, x, flat_off, flat_emo, map_off, map_emo, precision):
        m = x.shape[0]
        train = any(ctx.needs_input_grad[:3])
        precision = int(precision) if train else 0
        desc = with_precision(COARSE_DESC, precision)
        x = x.contiguous()
        xt_off, xt_emo = rows_to_tiles(x, map_off, precision), rows_to_tiles(x, map_emo, precision)
        img_off, img_emo = mlp_pack(desc, flat_off), mlp_pack(desc, flat_emo)
        rgb_off, hid_off = _mlp_forward(desc, img_off, xt_off, 0, m, m, train)
        rgb_emo, hid_emo = _mlp_forward(desc, img_emo, xt_emo, 0, m, m, train)
        ctx.desc, ctx.hidden, ctx.m, ctx.ld = desc, (hid_off, hid_emo), m, x.shape[1]
        ctx.save_for_backward(xt_off, xt_emo, img_off, img_emo, rgb_off, rgb_emo)
        return rgb_off, rgb_emo

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_off, d_emo):
        xt_off, xt_emo, img_off, img_emo, rgb_off, rgb_emo = ctx.saved_tensors
        m = ctx.m
        dev = xt_off.device
        hid_off, hid_emo = ctx.hidden
        dx_off = torch.empty(m, COARSE_GRAD_COLS, dtype=torch.float32, device=dev)
        dx_emo = torch.empty(m, COARSE_GRAD_COLS, dtype=torch.float32, device=dev)
        g_off, scratch = _mlp_backward(ctx.desc, img_off, xt_off, rgb_off, d_off.contiguous(), 0, m, m, hid_off, dx_off,
                                       COARSE_GRAD_COLS, 0)
        g_emo, _ = _mlp_backward(ctx.desc, img_emo, xt_emo, rgb_emo, d_emo.contiguous(), 0, m, m, hid_emo, dx_emo,
                                 COARSE_GRAD_COLS, 0, scratch)
        ctx.hidden = None
        g_x = torch.zeros(m, ctx.ld, dtype=torch.float32, device=dev)
        g_x[:, 0:12] = dx_off[:, 0:12]
        g_x[:, 12:24] = dx_emo[:, 0:12]
        g_x[:, 66:69] = dx_off[:, 12:15] + dx_emo[:, 12:15]
        return g_x, g_off, g_emo, None, None, None

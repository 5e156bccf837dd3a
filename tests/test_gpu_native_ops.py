"""GPU parity (-m gpu) of the reference-shaped native ops served through the C ABI
(include/esr_b200.h section 1) against the C oracle (oracle/render_utils_ref.c, a restatement of
app/utils/base/cuda/render_utils_kernel.cu).  Integer / index / mask outputs are compared bit-exactly;
fp32 outputs bit-exactly where the op order is the reference's, else within the tolerance written
beside the assert."""
import os

import numpy as np
import pytest
import torch

import esr_testlib as C
from esr_nerf_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from esr_nerf_b200 import render_utils as R
    from oracle import ref_harness as H
    return R, H


def _rays(n, seed, miss_frac=0.1, zero_comp=True):
    r = S.make_rays(n, seed)
    o, d = r["rays_o"].clone(), r["rays_d"].clone()
    g = torch.Generator().manual_seed(seed + 1)
    k = int(n * miss_frac)
    if k:  # rays that miss the box entirely
        d[:k] = -d[:k]
    if zero_comp and n >= 8:  # exact-zero direction components hit the 1e-6 substitution (kernel.cu:23-25)
        d[k:k + 3, 0] = 0.0
        d[k + 3:k + 5, 1] = 0.0
        o[k + 5] = torch.tensor([0.0, 0.0, -3.0])
        d[k + 5] = torch.tensor([0.0, 0.0, 1.1])
    _ = g
    return o.contiguous(), d.contiguous()


@pytest.mark.parametrize("n,stepdist,near", [(257, 0.021, 2.0), (2000, 0.0041, 2.0), (64, 0.05, 1e-5), (1, 0.01, 2.0)])
def test_sample_pts_on_rays_bit_exact(n, stepdist, near):
    R, H = _ops()
    o, d = _rays(n, 10 + n)
    mn, mx = S.BBOX_MIN, S.BBOX_MAX
    ref = H.sample_pts_on_rays(o, d, mn, mx, near, 1e9, stepdist)
    got = R.render_utils_cuda.sample_pts_on_rays(o.to(DEV), d.to(DEV), mn.to(DEV), mx.to(DEV), near, 1e9, stepdist)
    names = ["ray_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"]
    assert got[1].dtype == torch.bool and got[2].dtype == torch.int64 and got[4].dtype == torch.int64
    for name, a, b in zip(names, got, ref):
        assert a.shape == b.shape, name
        assert torch.equal(a.cpu(), b), name          # bit-exact, floats included (same op order and FMA shape)
    # sortedness / packing invariants the reference's callers rely on
    rid = got[2].cpu()
    assert (rid[1:] >= rid[:-1]).all()
    assert torch.equal(torch.bincount(rid, minlength=n), got[4].cpu())


def test_sample_pts_on_rays_empty_and_errors():
    R, _ = _ops()
    mn, mx = S.BBOX_MIN.to(DEV), S.BBOX_MAX.to(DEV)
    e = R.render_utils_cuda.sample_pts_on_rays(torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV), mn, mx,
                                                2.0, 1e9, 0.01)
    assert e[0].shape == (0, 3) and e[2].numel() == 0 and e[4].numel() == 0
    with pytest.raises(RuntimeError):   # render_utils.cpp:46-48 CHECK_INPUT semantics
        R.render_utils_cuda.sample_pts_on_rays(torch.zeros(4, 3), torch.zeros(4, 3, device=DEV), mn, mx, 2.0, 1e9, 0.01)
    with pytest.raises(RuntimeError):
        R.render_utils_cuda.sample_pts_on_rays(torch.zeros(3, 4, device=DEV).t(), torch.zeros(4, 3, device=DEV), mn,
                                                mx, 2.0, 1e9, 0.01)


def _alpha_stream(n_rays, seed, max_len=300, opaque_frac=0.3, empty_frac=0.2):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, max_len, (n_rays,), generator=g)
    lens[torch.rand(n_rays, generator=g) < empty_frac] = 0
    if n_rays > 3:
        lens[1] = 1           # single-sample ray
        lens[2] = 33          # one past a warp chunk
        lens[3] = 32
    ray_id = torch.repeat_interleave(torch.arange(n_rays), lens)
    m = ray_id.numel()
    alpha = torch.rand(m, generator=g) * 0.05
    opaque = torch.rand(n_rays, generator=g) < opaque_frac
    alpha = torch.where(opaque[ray_id] & (torch.rand(m, generator=g) < 0.2), torch.rand(m, generator=g), alpha)
    return alpha.contiguous(), ray_id.contiguous()


@pytest.mark.parametrize("n_rays,seed", [(5, 0), (300, 1), (4096, 2)])
def test_alpha2weight_fwd_bit_exact_bwd_close(n_rays, seed):
    R, H = _ops()
    alpha, ray_id = _alpha_stream(n_rays, seed)
    w, T, last, i_s, i_e = H.alpha2weight(alpha, ray_id, n_rays)
    gw, gT, glast, gi_s, gi_e = R.render_utils_cuda.alpha2weight(alpha.to(DEV), ray_id.to(DEV), n_rays)
    # early-stop index, weights, T and alphainv_last: bit-exact (sequential float/double recurrence, kernel.cu:591-603)
    assert torch.equal(gi_e.cpu(), i_e) and torch.equal(gi_s.cpu(), i_s)
    assert torch.equal(gw.cpu(), w) and torch.equal(gT.cpu(), T) and torch.equal(glast.cpu(), last)
    g = torch.Generator().manual_seed(seed + 100)
    grad_w, grad_last = torch.randn(alpha.shape, generator=g), torch.randn(n_rays, generator=g)
    ref = H.alpha2weight_backward(alpha, w, T, last, i_s, i_e, n_rays, grad_w, grad_last)
    got = R.render_utils_cuda.alpha2weight_backward(gw.new_tensor(alpha), gw, gT, glast, gi_s, gi_e, n_rays,
                                                    grad_w.to(DEV), grad_last.to(DEV))
    # the suffix sum is a warp scan here (re-associated): 1e-5 relative to the largest gradient of the stream
    assert C.rel_err(got, ref) < 1e-5
    assert torch.equal(got.cpu() == 0, ref == 0) or (got.cpu()[ref == 0].abs().max() == 0)


def test_alphas2weights_autograd_and_empty():
    R, H = _ops()
    alpha, ray_id = _alpha_stream(64, 9)
    a = alpha.to(DEV).requires_grad_(True)
    w, last = R.Alphas2Weights.apply(a, ray_id.to(DEV), 64)
    (w.sum() * 0.5 + last.sum()).backward()
    wr, Tr, lr, i_s, i_e = H.alpha2weight(alpha, ray_id, 64)
    ref = H.alpha2weight_backward(alpha, wr, Tr, lr, i_s, i_e, 64, torch.full_like(alpha, 0.5), torch.ones(64))
    assert C.rel_err(a.grad, ref) < 1e-5
    w0, T0, l0, s0, e0 = R.render_utils_cuda.alpha2weight(torch.zeros(0, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV), 7)
    assert w0.numel() == 0 and (l0 == 1).all() and (s0 == 0).all() and (e0 == 0).all()


@pytest.mark.parametrize("channels", [1, 3, 5])
def test_segment_coo_sum(channels):
    R, _ = _ops()
    alpha, ray_id = _alpha_stream(500, 3)
    g = torch.Generator().manual_seed(4)
    src = torch.randn(ray_id.numel(), channels, generator=g)
    ref = torch.zeros(500, channels, dtype=torch.float64).index_add_(0, ray_id, src.double())
    s = src.to(DEV).requires_grad_(True)
    out = R.segment_coo(src=s, index=ray_id.to(DEV), out=torch.zeros(500, channels, device=DEV), reduce="sum")
    # torch_scatter's summation order is unspecified (third-party, un-pinned): fp32 sum tolerance 1e-5 of max
    assert C.rel_err(out, ref) < 1e-5
    cot = torch.randn(500, channels, generator=g)
    (out * cot.to(DEV)).sum().backward()
    assert torch.equal(s.grad.cpu(), cot[ray_id])          # backward is a pure gather


def test_total_variation_add_grad():
    R, H = _ops()
    g = torch.Generator().manual_seed(5)
    p = torch.randn(1, 1, 9, 11, 13, generator=g) * 2
    grad = torch.randn(p.shape, generator=g)
    ref = grad.clone()
    H.total_variation_add_grad(p, ref, 0.3, 0.3, 0.3, True)
    got = grad.to(DEV)
    R.total_variation_cuda.total_variation_add_grad(p.to(DEV), got, 0.3, 0.3, 0.3, True)
    assert C.rel_err(got, ref) < 1e-6
    # sparse mode only touches voxels that already carry gradient (total_variation_kernel.cu:20)
    sp = grad.clone()
    sp[0, 0, :4] = 0
    got = sp.to(DEV)
    R.total_variation_cuda.total_variation_add_grad(p.to(DEV), got, 0.3, 0.3, 0.3, False)
    assert (got[0, 0, :4] == 0).all() and C.rel_err(got[0, 0, 4:], ref[0, 0, 4:]) < 1e-6


def test_exclusive_scan_large():
    from esr_nerf_b200 import fused
    g = torch.Generator().manual_seed(6)
    for n in (1, 4095, 4096, 4097, 1 << 20):
        c = torch.randint(0, 900, (n,), generator=g, dtype=torch.int32)
        out = fused.exclusive_scan(c.to(DEV)).cpu()
        ref = torch.cat([torch.zeros(1, dtype=torch.int64), c.long().cumsum(0)])
        assert torch.equal(out.long(), ref), n
    _ = np


@pytest.mark.parametrize("kind", ["near_threshold", "shell", "offset_bbox", "nonfinite"])
def test_mask_class_table_leaves_march_streams_identical(kind):
    """esr_mask_classify: the per-cell keep / drop classes shortcut MaskCache.forward (module.py:104-114) in the march
    kernel.  The packed streams with the table must equal those of the exact 8-tap test on every candidate — on a
    smooth density that hovers around the decision value (most cells undecided or one step from it), on the sparse
    shell, with a mask bbox that differs from the scene bbox (candidates outside the mask grid), and with non-finite
    densities."""
    from esr_nerf_b200 import fused

    res, n = 48, 4096
    g = torch.Generator().manual_seed(5)
    act_shift = float(np.log(1 / (1 - S.MASK_ALPHA_INIT) - 1))
    d_star = float(np.log(1e-3 / (1 - 1e-3)) - act_shift)
    if kind == "shell":
        dens = S.mask_density(res, True)
    else:
        coarse = torch.randn(1, 1, 7, 7, 7, generator=g)
        smooth = torch.nn.functional.interpolate(coarse, size=(res, res, res), mode="trilinear", align_corners=True)
        dens = d_star + 0.5 * smooth + 0.02 * torch.randn(1, 1, res, res, res, generator=g)
        if kind == "nonfinite":
            dens[0, 0, 10, 11, 12] = float("inf")
            dens[0, 0, 30, 5, 7] = float("nan")
            dens[0, 0, 20:24, 20:24, 20:24] = 50.0
    dens = dens.to(DEV).contiguous()
    mmin, mmax = (S.BBOX_MIN, S.BBOX_MAX) if kind != "offset_bbox" else (S.BBOX_MIN * 0.7 + 0.05, S.BBOX_MAX * 0.8)
    sc = fused.make_scene(S.BBOX_MIN.tolist(), S.BBOX_MAX.tolist(), (64, 64, 64), mmin.tolist(), mmax.tolist(),
                          (res, res, res), S.NEAR, 1e9, 0.5 * 2.1 / 64, 2.1 / 64, act_shift, 1e-3, 1e-4, 20.0)
    rays = S.make_rays(n, 77)
    o, d = rays["rays_o"].to(DEV), rays["rays_d"].to(DEV)
    sdf = S.sphere_sdf((64, 64, 64)).to(DEV).contiguous()
    cls = fused.mask_class_table(sc, dens)
    counts = torch.bincount(cls.long(), minlength=3).tolist()
    if kind in ("near_threshold", "shell"):
        assert counts[0] > 0 and counts[1] > 0 and counts[2] > 0, counts
    got = []
    for on in (True, False):
        fused.MASK_CLASSES = on
        try:
            s = fused.march(sc, o, d, None, dens, sdf)
        finally:
            fused.MASK_CLASSES = True
        got.append(s)
    a, b = got
    assert a.m1 == b.m1 and a.m1 > 0
    for f in ("n_steps", "cnt_inbox", "off_mask", "s_ray", "s_step"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert torch.equal(a.s_sdf.view(torch.int32), b.s_sdf.view(torch.int32))   # bit-identical, NaN-safe


@pytest.mark.parametrize("shape", [(16, 24, 20), (64, 64, 64), (7, 9, 11)])
def test_grad_block_flags_kernel_vs_torch(shape, monkeypatch):
    """esr_grad_block_flags (one warp per volume x block) against the torch reduction of the same map, and the voxel
    list built from it: everything outside is zero in every volume"""
    import torch.nn as nn

    from esr_nerf_b200.dist import TouchedBlockCompactor

    class Grids(nn.Module):
        def __init__(self):
            super().__init__()
            mk = lambda c: nn.Parameter(torch.zeros(1, c, *shape, device=DEV).contiguous(
                memory_format=torch.channels_last_3d if c > 1 else torch.contiguous_format))
            self.sdf, self.off_color, self.emo_color, self.brdf = (nn.Module() for _ in range(4))
            self.sdf.grid, self.off_color.grid, self.emo_color.grid, self.brdf.grid = mk(1), mk(6), mk(6), mk(6)

    model = Grids()
    comp = TouchedBlockCompactor(model)
    g = torch.Generator().manual_seed(sum(shape))
    for p in model.parameters():
        keep = torch.zeros(shape, dtype=torch.bool)
        keep.view(-1)[torch.randperm(keep.numel(), generator=g)[:5]] = True
        v = (torch.randn(p.shape, generator=g).abs() + 0.1) * keep
        p.grad = v.to(DEV).contiguous(memory_format=torch.channels_last_3d if p.shape[1] > 1 else torch.contiguous_format)
    rows = comp._grids_rows()
    got = comp.block_flags(rows).clone()
    monkeypatch.setenv("ESR_BLOCK_FLAGS_TORCH", "1")
    want = comp.block_flags(rows)
    assert torch.equal(got, want) and 0 < int(want.sum())
    comp.select(got)
    assert comp.outside_is_zero()
    buf = comp.pack(rows)
    for p in model.parameters():
        p.grad.zero_()
    comp.unpack(comp._grids_rows(), buf)
    assert int(sum((p.grad != 0).sum() for p in model.parameters())) == 5 * (1 + 6 + 6 + 6)

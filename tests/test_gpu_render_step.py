"""GPU parity (-m gpu) of the one-call render step of the C ABI (esr_render_voxurff_fwd / _bwd, include/esr_b200.h;
SURVEY.md §8b last row) — called through ctypes exactly as a non-Python host would, with caller-owned parameters,
gradients, outputs and one workspace:
  * outputs against the golden vectors made by the reference's own VoxurfF (tests/golden/*.npz), 1e-4;
  * outputs and every gradient (three grid volumes, three flat MLP gradients) against the drop-in class's own path
    (fused.py orchestration of the same stage kernels): outputs bit-equal, gradients equal up to the order of atomics;
  * the capacity protocol: a too-small workspace fails with ESR_ERR_CAPACITY and reports the bytes needed."""
import ctypes

import pytest
import torch

import esr_testlib as C
from esr_nerf_b200 import _lib
from esr_nerf_b200 import synthetic as S
from esr_nerf_b200._lib import VoxurffStep, ptr, stream_ptr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT_KEYS = ("srgb/rgb", "lin/rgb", "etc/alphainv_cum")


def _python_path(m, rays, s_val, cot):
    """the drop-in's forward / backward with the three flat parameter images as leaves (so their gradients are visible)"""
    leaves = {k: m._flat(k).detach().clone().requires_grad_(True) for k in ("off", "emo", "tone")}
    m._flat = lambda which: leaves[which]
    for p in m.parameters():
        p.grad = None
    out = m(s_val=s_val, **{k: v for k, v in rays.items() if k != "rgbs"})
    sum((out[k] * cot[k]).sum() for k in OUT_KEYS).backward()
    del m._flat
    return out, leaves


@pytest.mark.parametrize("case", ["fine_sparse_s20", "fine_dense_s220"])
@pytest.mark.parametrize("precision", [1, 0])
def test_one_call_step_vs_golden_and_python_path(case, precision):
    fx, weights = C.load_case(case)
    m = C.build_product_model(fx, weights, DEV)
    m.mlp_mode = "x2" if precision else "bf16"
    n = int(fx["n_rays"])
    rays = {k: v.to(DEV) for k, v in S.make_rays(n, int(fx["ray_seed"])).items()}
    s_val = float(fx["s_val"])
    g = torch.Generator().manual_seed(3)
    cot = {"srgb/rgb": torch.randn(n, 3, generator=g).to(DEV), "lin/rgb": torch.randn(n, 3, generator=g).to(DEV),
           "etc/alphainv_cum": torch.randn(n, generator=g).to(DEV)}
    ref_out, leaves = _python_path(m, rays, s_val, cot)

    L = _lib.lib()
    from esr_nerf_b200 import fused
    sc = m._scene(s_val)
    step = VoxurffStep()
    step.scene = sc
    dens = m.mask_cache.density
    cls = fused.mask_class_table(sc, dens)
    grids = [m.sdf.grid.detach(), m.off_color.grid.detach(), m.emo_color.grid.detach()]
    flats = [leaves[k].detach().contiguous() for k in ("off", "emo", "tone")]
    step.mask_density, step.mask_cls = dens.data_ptr(), cls.data_ptr()
    step.sdf_grid, step.off_color_grid, step.emo_color_grid = (t.data_ptr() for t in grids)
    step.flat_off, step.flat_emo, step.flat_tone = (t.data_ptr() for t in flats)
    step.precision = precision
    out = [torch.empty(n, 3, device=DEV), torch.empty(n, 3, device=DEV), torch.empty(n, device=DEV)]
    ro, rd, vd, em = rays["rays_o"], rays["rays_d"], rays["viewdirs"], rays["em_modes"].long().contiguous()

    # capacity protocol: a workspace that holds only the per-ray arrays fails and reports what is needed
    small = torch.empty(int(L.esr_render_voxurff_workspace_bytes(ctypes.byref(sc), n, 0, 0, precision)), dtype=torch.uint8, device=DEV)
    step.workspace, step.workspace_bytes = small.data_ptr(), small.numel()
    rc = L.esr_render_voxurff_fwd(ctypes.byref(step), ptr(ro), ptr(rd), ptr(vd), ptr(em), n, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                  stream_ptr())
    assert rc == -3 and step.workspace_needed > small.numel() and b"workspace" in L.esr_last_error()
    ws = torch.empty(int(step.workspace_needed), dtype=torch.uint8, device=DEV)
    step.workspace, step.workspace_bytes = ws.data_ptr(), ws.numel()
    rc = L.esr_render_voxurff_fwd(ctypes.byref(step), ptr(ro), ptr(rd), ptr(vd), ptr(em), n, ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                  stream_ptr())
    if rc == -3:      # (the first report is an upper estimate made before M3 is known; the second is exact)
        ws = torch.empty(int(step.workspace_needed), dtype=torch.uint8, device=DEV)
        step.workspace, step.workspace_bytes = ws.data_ptr(), ws.numel()
        rc = L.esr_render_voxurff_fwd(ctypes.byref(step), ptr(ro), ptr(rd), ptr(vd), ptr(em), n, ptr(out[0]), ptr(out[1]),
                                      ptr(out[2]), stream_ptr())
    assert rc == 0, L.esr_last_error()
    st = m.last_streams["streams"] if m.keep_streams else None
    assert step.n_rays == n and step.m3 > 0 and step.m3_on <= step.m3 <= step.m1 and step.n_on == int((em == 1).sum())
    for k, t in zip(OUT_KEYS, out):
        assert torch.equal(t, ref_out[k]), k                                  # same kernels, same order, same inputs
        assert C.rel_err(t, torch.from_numpy(fx["out/" + k])) < (1e-4 if precision else 1e-2), k    # and the reference's values

    g_grids = [torch.zeros_like(t) for t in grids]
    g_flats = [torch.zeros_like(t) for t in flats]
    cots = [cot[k].contiguous() for k in OUT_KEYS]
    rc = L.esr_render_voxurff_bwd(ctypes.byref(step), ptr(ro), ptr(rd), ptr(cots[0]), ptr(cots[1]), ptr(cots[2]), ptr(g_grids[0]),
                                  ptr(g_grids[1]), ptr(g_grids[2]), ptr(g_flats[0]), ptr(g_flats[1]), ptr(g_flats[2]), stream_ptr())
    assert rc == 0, L.esr_last_error()
    torch.cuda.synchronize()
    for name, got, want in (("sdf.grid", g_grids[0], m.sdf.grid.grad), ("off_color.grid", g_grids[1], m.off_color.grid.grad),
                            ("emo_color.grid", g_grids[2], m.emo_color.grid.grad), ("flat off", g_flats[0], leaves["off"].grad),
                            ("flat emo", g_flats[1], leaves["emo"].grad), ("flat tone", g_flats[2], leaves["tone"].grad)):
        assert want is not None and got.stride() == want.stride(), name
        mx, l2 = C.grad_err(got, want)
        assert mx < 1e-4 and l2 < 1e-4, (name, mx, l2)                        # atomics' summation order only

"""GPU parity (-m gpu) of `neus_alpha: grad` (functions.py:45-69; selected by voxurff.py:151-154 / esrnerf.py:197-200, no
shipped config uses it): the section-point SDFs of a sample come from the view-projected finite-difference SDF gradient
(k_neus_cos_fwd -> k_neus_alpha<true>) and the SDF grid receives gradient through the six taps of EVERY M1 sample
(k_alpha_scan_bwd<true>, k_sdf_scatter<true>, k_neus_cos_bwd).  The oracle ports are pinned to the reference's own
classes built with that option on the CPU (tests/test_oracle_cpu.py::test_port_matches_reference_neus_alpha_grad,
tests/test_esrnerf_cpu.py::test_esrnerf_port_matches_reference_neus_alpha_grad); the kernels' arithmetic is modelled in
tests/test_neus_grad_model_cpu.py.  Tolerances as everywhere: streams bit-exact, outputs 1e-4, gradients 1e-4 with fp32
nets / 1e-2 with the tensor-core nets."""
import pytest
import torch

import esr_testlib as C
from esr_nerf_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT_KEYS = ("etc/alphainv_cum", "etc/white_bg", "srgb/rgb", "lin/rgb")


def _fine_oracle(fx, weights, rays):
    from oracle import voxurf_port as P

    n = rays["rays_o"].shape[0]
    scene = C.oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    scene["neus_alpha"] = "grad"
    params, leaves = C.oracle_params(scene, weights)
    ref, inter = P.voxurff_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                            rays["em_modes"], float(fx["s_val"]))
    cot = C.cotangents(n)
    sum((ref[k] * cot[k]).sum() for k in cot).backward()
    return scene, ref, inter, leaves


def _fine_product(fx, weights, mode, rays):
    m = C.build_product_model(fx, weights, DEV, neus_alpha="grad")
    m.mlp_mode, m.on_first_order, m.keep_streams = mode, True, True
    n = rays["rays_o"].shape[0]
    out = m(s_val=float(fx["s_val"]), **{k: v.to(DEV) for k, v in rays.items()})
    cot = C.cotangents(n)
    sum((out[k] * cot[k].to(DEV)).sum() for k in cot).backward()
    return m, out


@pytest.mark.parametrize("mode", ["torch_fp32", "x2"])
def test_voxurff_grad_alpha_vs_oracle_port(mode):
    fx, weights = C.load_case("fine_sparse_s60_big")
    rays = S.make_rays(1536, 4711)
    scene, ref, inter, leaves = _fine_oracle(fx, weights, rays)
    m, out = _fine_product(fx, weights, mode, rays)
    st = m.last_streams["streams"]
    assert st.s_cos is not None and st.m1 == inter["m1_ray"].numel() and st.m3 > 2000
    # M1 stream in slot order -> the reference's ray-major order
    key1 = st.s_ray.long().cpu() * (1 << 20) + st.s_step.long().cpu()
    o1 = torch.argsort(key1, stable=True)
    assert torch.equal(st.s_ray.long().cpu()[o1], inter["m1_ray"]) and torch.equal(st.s_step.long().cpu()[o1], inter["m1_step"])
    assert C.rel_err(st.s_alpha.cpu()[o1], inter["m1_alpha"]) < 1e-5
    # the mode really is the other function of the grid
    from oracle import voxurf_port as P
    assert (P.neus_alpha_interp(inter["m1_ray"], inter["m1_sdf"].detach(), float(fx["s_val"])) -
            inter["m1_alpha"].detach()).abs().max() > 1e-3
    ray, step = st.h_ray.long().cpu(), st.h_step.long().cpu()
    o3 = torch.argsort(ray * (1 << 20) + step, stable=True)
    assert torch.equal(ray[o3], inter["m3_ray"]) and torch.equal(step[o3], inter["m3_step"])
    assert C.rel_err(m.last_streams["h_w"].cpu()[o3], inter["m3_weights"]) < 1e-4
    for k in OUT_KEYS:
        assert C.rel_err(out[k], ref[k]) < 1e-4, k
    checked, bad = 0, {}
    for name, p in m.named_parameters():
        if name not in leaves or leaves[name].grad is None:
            continue
        if mode == "x2":
            mx, l2 = C.grad_err(p.grad.contiguous(), leaves[name].grad)
            if not (mx < 1e-2 and l2 < 1e-2):
                bad[name] = (mx, l2)
        else:
            # grids at 1e-4; MLP tensors at 1e-3, as in the other fresh-ray fp32 comparisons (tests/test_gpu_voxurff.py::
            # test_odd_grid_vs_oracle_port_fresh_rays): with ~3 x 10^4 rows per net ONE ReLU mask on which two correct fp32
            # evaluations disagree (library GEMM here, the CPU port there) moves the entries of the layers below it by a few
            # 1e-4 of the tensor's maximum (measured 5.4e-4 on off_rgbnet.linear.0.weight, 14 % of its entries)
            mlp = "net" in name or "tonemapper" in name
            ok, msg = C.grad_close(p.grad.contiguous(), leaves[name].grad, 1e-3 if mlp else 1e-4)
            if not ok:
                bad[name] = msg
        checked += 1
    assert not bad, bad
    assert checked >= 3 + 8 + 8 + 4


def test_voxurff_grad_alpha_vs_golden():
    """the same step against the fixture made by the reference's own VoxurfF built with `neus_alpha: grad`
    (tests/golden/voxurff_fine_grad_sparse_s60_big.npz, oracle/make_golden.py --neus-grad-only): M1 / M3 streams bit-exact,
    alphas, weights and rendered outputs; its gradient digests are held against the port on the CPU
    (tests/test_oracle_cpu.py::test_port_matches_golden_neus_alpha_grad), the kernels against the port above"""
    import numpy as np

    fx, weights = C.load_case("fine_grad_sparse_s60_big")
    rays = S.make_rays(int(fx["n_rays"]), int(fx["ray_seed"]))
    m, out = _fine_product(fx, weights, "x2", rays)
    st = m.last_streams["streams"]
    o1 = torch.argsort(st.s_ray.long().cpu() * (1 << 20) + st.s_step.long().cpu(), stable=True)
    assert np.array_equal(st.s_ray.cpu()[o1].numpy(), fx["m1_ray"]) and np.array_equal(st.s_step.cpu()[o1].numpy(), fx["m1_step"])
    assert C.rel_err(st.s_alpha.cpu()[o1], torch.from_numpy(fx["m1_alpha"])) < 1e-5
    o3 = torch.argsort(st.h_ray.long().cpu() * (1 << 20) + st.h_step.long().cpu(), stable=True)
    assert np.array_equal(st.h_ray.cpu()[o3].numpy(), fx["m3_ray"]) and np.array_equal(st.h_step.cpu()[o3].numpy(), fx["m3_step"])
    assert C.rel_err(m.last_streams["h_w"].cpu()[o3], torch.from_numpy(fx["m3_weights"])) < 1e-4
    for k in OUT_KEYS:
        assert C.rel_err(out[k], torch.from_numpy(fx["out/" + k])) < 1e-4, k


def test_voxurff_grad_alpha_sdf_gradient_alone():
    """only alphainv_last carries a cotangent: every bit of gradient reaches the SDF grid through the alpha path — the
    dL/dsdf scatter plus the six-tap scatter of dL/diter_cos — with no shading term to hide an error in it"""
    fx, weights = C.load_case("fine_sparse_s20")
    rays = S.make_rays(int(fx["n_rays"]), 1213)
    from oracle import voxurf_port as P

    scene = C.oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    scene["neus_alpha"] = "grad"
    params, leaves = C.oracle_params(scene, weights)
    ref, _ = P.voxurff_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"], rays["em_modes"],
                                        float(fx["s_val"]))
    cot = C.cotangents(rays["rays_o"].shape[0])["etc/alphainv_cum"]
    (ref["etc/alphainv_cum"] * cot).sum().backward()
    m = C.build_product_model(fx, weights, DEV, neus_alpha="grad")
    out = m(s_val=float(fx["s_val"]), **{k: v.to(DEV) for k, v in rays.items()})
    (out["etc/alphainv_cum"] * cot.to(DEV)).sum().backward()
    assert C.rel_err(out["etc/alphainv_cum"], ref["etc/alphainv_cum"]) < 1e-4
    g, r = m.sdf.grid.grad, leaves["sdf.grid"].grad
    assert r.abs().max() > 0
    mx, l2 = C.grad_err(g, r)
    assert mx < 1e-4 and l2 < 1e-4, (mx, l2)


def test_voxurff_grad_alpha_inference_vs_oracle_port():
    from oracle import voxurf_port as P

    fx, weights = C.load_case("fine_sparse_s20")
    s_val = float(fx["s_val"])
    scene = C.oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    scene["neus_alpha"] = "grad"
    params, _ = C.oracle_params(scene, weights, requires_grad=False)
    rays = S.make_rays(512, 77)
    pos_rt = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(3)))[0]
    with torch.no_grad():
        ref, inter = P.voxurff_forward_evaluate(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                                torch.tensor(1), pos_rt, s_val)
    m = C.build_product_model(fx, weights, DEV, neus_alpha="grad")
    m.mlp_mode = "torch_fp32"
    m.eval()
    out = m(rays_o=rays["rays_o"].to(DEV), rays_d=rays["rays_d"].to(DEV), viewdirs=rays["viewdirs"].to(DEV),
            em_modes=torch.tensor(1), pos_rt=pos_rt)
    assert set(out) == set(ref)
    for k in sorted(out):
        assert C.rel_err(out[k], ref[k]) < 1e-4, k


def test_esrnerf_grad_alpha_vs_oracle_port():
    """LTS / PDRA step with `neus_alpha: grad`: primary AND secondary rays (whose view directions are the sampled
    hemisphere directions, esrnerf.py:346-361) against the port run side by side, outputs + every parameter gradient"""
    from oracle import esrnerf_port as E

    fx, weights = C.load_esrnerf_case("pdra_sparse_s60")       # the fixture's rays, draws and sizes; only the alpha mode differs
    fx = dict(fx, neus_alpha="grad")
    m = C.build_product_esrnerf(fx, weights, DEV)
    assert m.neus_alpha == "grad"
    m.keep_streams = True
    m.draws = E.FixedDraws(int(fx["draw_seed"]))
    n = int(fx["n_rays"])
    rays = S.make_rays(n, int(fx["ray_seed"]))
    batch = {k: v.to(DEV) for k, v in rays.items() if k != "rgbs"}
    out = m(s_val=float(fx["s_val"]), uncert_masks=S.uncert_masks(n).to(DEV), normal_eps=float(fx["normal_eps"]),
            emit_eps=float(fx["emit_eps"]), **batch)
    ref, inter, leaves, _ = C.run_esrnerf_port(fx, weights)
    st, st2 = m.last_streams["streams"], m.last_streams["lts"]["streams"]
    assert st.s_cos is not None and st2.s_cos is not None
    assert torch.equal(st.h_ray.long().cpu(), inter["m3_ray"]) and torch.equal(st.h_step.long().cpu(), inter["m3_step"])
    assert torch.equal(st2.h_ray.long().cpu(), inter["lts"]["m3_ray"])
    assert torch.equal(st2.h_step.long().cpu(), inter["lts"]["m3_step"])
    fp32_keys = ("etc/alphainv_cum", "etc/white_bg", "etc/normal", "etc/normal_eps")
    for k in sorted(out):
        assert tuple(out[k].shape) == tuple(ref[k].shape), k
        assert C.rel_err(out[k], ref[k]) < (1e-4 if k in fp32_keys else 1e-2), k
    cot = C.esrnerf_cotangents(ref)
    sum((ref[k] * cot[k]).sum() for k in cot).backward()
    sum((out[k] * cot[k].to(DEV)).sum() for k in cot).backward()
    bad, checked = {}, 0
    for name, p in m.named_parameters():
        if name not in leaves or leaves[name].grad is None:
            continue
        assert p.grad is not None, name
        mx, l2 = C.grad_err(p.grad.contiguous(), leaves[name].grad)
        # the SDF grid is where this mode's kernels scatter: held to 1e-2 in both metrics.  The other 42 tensors only see
        # the mode through the sample weights (their own all-tensor 1e-2 bar is test_esrnerf_gradients_vs_golden's)
        tol = 1e-2 if name == "sdf.grid" else 3e-2
        if not (mx < tol and l2 < tol):
            bad[name] = (mx, l2)
        checked += 1
    assert not bad, bad
    assert checked >= 40


# ---------------------------------------------------------------------------------------------------
# coarse stage: the SDF gradient is the trilinear tap of the central-difference volume (voxurfc.py:204-210)
# ---------------------------------------------------------------------------------------------------
COARSE_KEYS = ("etc/alphainv_cum", "etc/white_bg", "srgb/rgb")


def _coarse_oracle(fx, weights, rays, only=None):
    from oracle import voxurfc_port as PC

    scene = C.coarse_oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    scene["neus_alpha"] = "grad"
    params, leaves = C.coarse_oracle_params(scene, weights)
    ref, inter = PC.voxurfc_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                             rays["em_modes"], float(fx["s_val"]))
    cot = C.coarse_cotangents(rays["rays_o"].shape[0])
    sum((ref[k] * cot[k]).sum() for k in cot if only is None or k == only).backward()
    return ref, inter, leaves


def _coarse_product(fx, weights, rays, mode, only=None):
    m = C.build_product_coarse(fx, weights, DEV, neus_alpha="grad")
    m.mlp_mode, m.keep_streams = mode, True
    out = m(s_val=float(fx["s_val"]), **{k: v.to(DEV) for k, v in rays.items()})
    cot = C.coarse_cotangents(rays["rays_o"].shape[0])
    sum((out[k] * cot[k].to(DEV)).sum() for k in cot if only is None or k == only).backward()
    return m, out


@pytest.mark.parametrize("mode", ["torch_fp32", "x2"])
def test_voxurfc_grad_alpha_vs_oracle_port(mode):
    fx, weights = C.load_coarse_case("coarse_sparse_s5")
    rays = S.make_rays(1024, 515)
    ref, inter, leaves = _coarse_oracle(fx, weights, rays)
    m, out = _coarse_product(fx, weights, rays, mode)
    st = m.last_streams["streams"]
    assert st.s_cos is not None and st.m3 > 2000
    assert torch.equal(st.s_ray.long().cpu(), inter["m1_ray"]) and torch.equal(st.s_step.long().cpu(), inter["m1_step"])
    assert C.rel_err(st.s_alpha, inter["m1_alpha"]) < 1e-5
    assert torch.equal(st.h_ray.long().cpu(), inter["m3_ray"]) and torch.equal(st.h_step.long().cpu(), inter["m3_step"])
    assert C.rel_err(m.last_streams["h_w"], inter["m3_weights"]) < 1e-4
    for k in COARSE_KEYS:
        assert C.rel_err(out[k], ref[k]) < 1e-4, k
    bad, checked = {}, 0
    for name, p in m.named_parameters():
        if name not in leaves or leaves[name].grad is None:
            continue
        if mode == "x2":
            ok, msg = C.grad_close(p.grad.contiguous(), leaves[name].grad, 1e-2, l2_factor=1.0)
        else:       # grids 1e-4; fp32 library GEMMs vs the CPU port on ~10^4 rows: 1e-3 on the nets (see above)
            ok, msg = C.grad_close(p.grad.contiguous(), leaves[name].grad, 1e-3 if "rgbnet" in name else 1e-4)
        if not ok:
            bad[name] = msg
        checked += 1
    assert not bad, bad
    assert checked == 3 + 6 + 6


def test_voxurfc_grad_alpha_sdf_gradient_alone():
    """cotangent on alphainv_last only: all gradient reaches the raw SDF grid through the alpha path — dL/dsdf through the
    smoothing convolution, dL/diter_cos through the central-difference volume"""
    fx, weights = C.load_coarse_case("coarse_sparse_s5")
    rays = S.make_rays(512, 516)
    ref, _, leaves = _coarse_oracle(fx, weights, rays, only="etc/alphainv_cum")
    m, out = _coarse_product(fx, weights, rays, "x2", only="etc/alphainv_cum")
    assert C.rel_err(out["etc/alphainv_cum"], ref["etc/alphainv_cum"]) < 1e-4
    r = leaves["sdf.grid"].grad
    assert r.abs().max() > 0
    mx, l2 = C.grad_err(m.sdf.grid.grad, r)
    assert mx < 1e-4 and l2 < 1e-4, (mx, l2)


def test_voxurfc_grad_alpha_inference_vs_oracle_port():
    from oracle import voxurfc_port as PC

    fx, weights = C.load_coarse_case("coarse_dense_s25")
    rays = S.make_rays(400, 78)
    pos_rt = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(3)))[0]
    scene = C.coarse_oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    scene["neus_alpha"] = "grad"
    params, _ = C.coarse_oracle_params(scene, weights, requires_grad=False)
    with torch.no_grad():
        ref, inter = PC.voxurfc_forward_evaluate(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                                 torch.tensor(1), pos_rt, float(fx["s_val"]))
    m = C.build_product_coarse(fx, weights, DEV, neus_alpha="grad")
    m.mlp_mode, m.keep_streams = "torch_fp32", True
    m.eval()
    out = m(rays_o=rays["rays_o"].to(DEV), rays_d=rays["rays_d"].to(DEV), viewdirs=rays["viewdirs"].to(DEV),
            em_modes=torch.tensor(1), pos_rt=pos_rt.to(DEV))
    assert set(out) == set(ref)
    st = m.last_streams["streams"]
    assert torch.equal(st.h_ray.long().cpu(), inter["m3_ray"]) and torch.equal(st.h_step.long().cpu(), inter["m3_step"])
    for k in ref:
        assert C.rel_err(out[k], ref[k]) < 1e-4, (k, C.rel_err(out[k], ref[k]))

"""GPU parity (-m gpu) of the LTS / PDRA stage drop-in (esr_nerf_b200.ESRNeRF -> libesr_b200.so) against
 (1) the golden vectors produced by the reference's own ESRNeRF.forward_training (tests/golden/esrnerf_*.npz,
     oracle/make_golden.py) and
 (2) the travelling oracle port (oracle/esrnerf_port.py) run beside it on the same inputs and random draws.

Tolerances (BASELINE.json north_star): sample streams (primary and LTS secondary rays) bit-exact; fp32 stages
(SDF value, analytic SDF gradient, transmittance) 1e-4; everything downstream of the tensor-core MLPs 1e-2
on rendered / per-sample outputs.  Gradients (default mlp_mode "x2": fp16 hi + lo forward operands, fp16 data-gradient
chain): EVERY parameter gradient — five grids, five nets, the SG environment map — within 1e-2 of the reference's in
max-abs / max|ref| AND in relative L2."""
import os

import numpy as np
import pytest
import torch

import esr_testlib as C
from esr_nerf_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_KEYS = ("etc/alphainv_cum", "etc/white_bg", "etc/normal", "etc/normal_eps")


def _run_product(fx, weights):
    from oracle import esrnerf_port as E

    m = C.build_product_esrnerf(fx, weights, DEV)
    m.keep_streams = True
    m.draws = E.FixedDraws(int(fx["draw_seed"]))
    n = int(fx["n_rays"])
    rays = S.make_rays(n, int(fx["ray_seed"]))
    batch = {k: v.to(DEV) for k, v in rays.items() if k != "rgbs"}
    out = m(s_val=float(fx["s_val"]), uncert_masks=S.uncert_masks(n).to(DEV), normal_eps=float(fx["normal_eps"]),
            emit_eps=float(fx["emit_eps"]), **batch)
    return m, out


@pytest.mark.parametrize("case", C.ESRNERF_CASES)
def test_esrnerf_outputs_vs_golden_and_port(case):
    fx, weights = C.load_esrnerf_case(case)
    m, out = _run_product(fx, weights)
    ref, inter, leaves, _ = C.run_esrnerf_port(fx, weights)
    # --- integer streams: bit-exact (primary rays and the LTS secondary rays) ---
    st = m.last_streams["streams"]
    assert torch.equal(st.h_ray.long().cpu(), inter["m3_ray"]) and torch.equal(st.h_step.long().cpu(), inter["m3_step"])
    assert torch.equal(st.s_ray.long().cpu(), inter["m1_ray"]) and torch.equal(st.s_step.long().cpu(), inter["m1_step"])
    # ray_pts / manual-trilinear SDF: same formulas; the step length comes from a CUDA cube root here and a CPU one
    # in the oracle (modules.voxel_geometry) -> positions within an ulp
    assert (m.last_streams["pts"].cpu() - inter["m3_pts"]).abs().max() < 2.5e-7
    assert C.rel_err(st.s_sdf, inter["m1_sdf"]) < 1e-5
    st2 = m.last_streams["lts"]["streams"]
    assert torch.equal(st2.h_ray.long().cpu(), inter["lts"]["m3_ray"])
    assert torch.equal(st2.h_step.long().cpu(), inter["lts"]["m3_step"])
    assert C.rel_err(m.last_streams["h_w"], inter["m3_weights"]) < 1e-4
    assert C.rel_err(m.last_streams["lts"]["h_w"], inter["lts"]["m3_weights"]) < 1e-4
    # --- outputs ---
    assert set(out) == set(ref)
    for k in sorted(out):
        assert tuple(out[k].shape) == fx["out/" + k].shape == tuple(ref[k].shape), k
        tol = 1e-4 if k in FP32_KEYS else 1e-2
        assert C.rel_err(out[k], ref[k]) < tol, (k, "vs port")
        assert C.rel_err(out[k], torch.from_numpy(fx["out/" + k])) < tol, (k, "vs golden")


def _port_grads(fx, weights, precision):
    from oracle import voxurf_port as P

    P.MLP_PRECISION = precision
    try:
        ref, _, leaves, _ = C.run_esrnerf_port(fx, weights)
        cot = C.esrnerf_cotangents(ref)
        sum((ref[k] * cot[k]).sum() for k in cot).backward()
    finally:
        P.MLP_PRECISION = "fp32"
    return leaves


@pytest.mark.parametrize("case", C.ESRNERF_CASES)
def test_esrnerf_gradients_vs_golden(case):
    """Every parameter gradient of the stage (5 grids / nets x layers + the SG environment map) under random
    cotangents on all 16 outputs against the reference's own gradients (golden digests): 1e-2 in max-norm and in
    relative L2, no tensor excepted."""
    fx, weights = C.load_esrnerf_case(case)
    m, out = _run_product(fx, weights)
    cot = C.esrnerf_cotangents(out)
    loss = sum((out[k] * cot[k].to(DEV)).sum() for k in cot)
    loss.backward()
    assert abs(loss.item() - float(fx["loss"])) < 1e-3 * max(1.0, abs(float(fx["loss"])))
    checked, bad = 0, {}
    for name, p in m.named_parameters():
        if f"grad/{name}/idx" not in fx:
            continue
        assert p.grad is not None, name
        g = p.grad.contiguous().cpu()
        flat = g.reshape(-1)
        idx = torch.from_numpy(fx[f"grad/{name}/idx"])
        refv = torch.from_numpy(fx[f"grad/{name}/val"])
        abs_sum = float(fx[f"grad/{name}/abs_sum"])
        s_err = abs(flat.double().abs().sum().item() - abs_sum) / max(abs_sum, 1e-12)
        mx, l2 = C.grad_err(flat[idx], refv)
        if not (mx < 1e-2 and l2 < 1e-2 and s_err < 1e-2):      # every tensor, both metrics
            bad[name] = dict(vs_golden=(mx, l2, s_err))
        checked += 1
    assert not bad, bad
    assert checked >= 40


@pytest.mark.parametrize("grid", ["fixture", "odd", "non_cubic"])
def test_esrnerf_port_as_live_oracle_on_new_rays(grid, monkeypatch):
    """rays / draws the fixtures never saw: product vs the port run side by side — on the fixture's grid, on an odd one
    (37^3: the scalar forms of the paired REDs, the plain encode-backward scatter) and in a non-cubic box (46 x 40 x 33
    voxels: the port is pinned to the reference's class there by tests/test_esrnerf_cpu.py)"""
    from oracle import esrnerf_port as E

    fx, weights = C.load_esrnerf_case("lts_sparse_s220")
    fx = dict(fx, ray_seed=2025, draw_seed=99, n_rays=200, s_val=90.0, pdra_mode=1)
    if grid == "odd":
        fx["num_voxels"] = 37 ** 3
    elif grid == "non_cubic":
        monkeypatch.setattr(S, "BBOX_MIN", torch.tensor([-1.05, -0.9, -0.75]))
        monkeypatch.setattr(S, "BBOX_MAX", torch.tensor([1.05, 0.9, 0.75]))
        fx.update(num_voxels=46 * 40 * 33, mask_res=20, sparse=0)
    m, out = _run_product(fx, weights)
    ref, inter, _, _ = C.run_esrnerf_port(fx, weights, E.FixedDraws(99))
    st = m.last_streams["streams"]
    if grid == "non_cubic":
        assert len(set(m.sdf.grid.shape[2:])) == 3 and st.m3 > 200
    assert torch.equal(st.h_ray.long().cpu(), inter["m3_ray"]) and torch.equal(st.h_step.long().cpu(), inter["m3_step"])
    for k in sorted(out):
        assert tuple(out[k].shape) == tuple(ref[k].shape), k
        assert C.rel_err(out[k], ref[k]) < (1e-4 if k in FP32_KEYS else 1e-2), k


@pytest.mark.parametrize("ray_sampling,env_activation", [("fib", "softplus"), ("random", "relu"), ("fib", "sigmoid")])
def test_esrnerf_other_samplers_and_envmaps_vs_port(ray_sampling, env_activation):
    """`ray_sampling: fib` (no random draw for the directions) and the other environment-map activations
    (esrnerf.py:188-195): product vs the port, which tests/test_esrnerf_cpu.py pins against the reference class"""
    from oracle import esrnerf_port as E

    fx, weights = C.load_esrnerf_case("lts_sparse_s220")
    fx = dict(fx, ray_seed=2026, draw_seed=98, n_rays=200, s_val=90.0, pdra_mode=1, ray_sampling=ray_sampling,
              env_activation=env_activation)
    m, out = _run_product(fx, weights)
    assert m.fib_sampling == (ray_sampling == "fib")
    ref, inter, _, _ = C.run_esrnerf_port(fx, weights, E.FixedDraws(98))
    for k in sorted(out):
        assert tuple(out[k].shape) == tuple(ref[k].shape), k
        assert C.rel_err(out[k], ref[k]) < (1e-4 if k in FP32_KEYS else 1e-2), k


@pytest.mark.parametrize("case", C.ESRNERF_CASES)
def test_esrnerf_inference_entry_points_vs_golden(case):
    """forward_evaluate (21 maps incl. the PBR decomposition over several LTS chunks), eval_emit, eval_esp against the
    outputs of the reference's own methods (fixture) — same random draws through FixedDraws."""
    from oracle import esrnerf_port as E

    fx, weights = C.load_esrnerf_case(case)
    m = C.build_product_esrnerf(fx, weights, DEV)
    m.eval()
    m.keep_streams = True
    n = int(fx["n_rays"])
    rays = {k: v.to(DEV) for k, v in S.make_rays(n, int(fx["ray_seed"])).items()}
    kw = dict(rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"])
    for em in (0, 1):
        m.draws = E.FixedDraws(int(fx["draw_seed"]) + 100)
        out = m(em_modes=torch.tensor(em), pos_rt=torch.from_numpy(fx["pos_rt"]), render_pbr=True,
                chunk_sz=int(fx["eval_chunk"]), **kw)
        keys = {k.split("/", 1)[1] for k in fx if k.startswith(f"eval{em}/")}
        assert set(out) == keys
        for k in sorted(out):
            ref = torch.from_numpy(fx[f"eval{em}/{k}"])
            assert tuple(out[k].shape) == tuple(ref.shape), k
            tol = 1e-4 if k in ("etc/depth", "etc/disp", "etc/normal", "etc/white_bg") else 1e-2
            assert C.rel_err(out[k], ref) < tol, (em, k, C.rel_err(out[k], ref))
    out = m(em_modes=1, pos_rt=torch.from_numpy(fx["pos_rt"]), render_pbr=False, chunk_sz=64, **kw)   # pdra.py:656-657
    assert len(out) == 16 and C.rel_err(out["lin/emit"], torch.from_numpy(fx["eval1/lin/emit"])) < 1e-2
    assert C.rel_err(m.eval_emit(**kw), torch.from_numpy(fx["eval_emit"])) < 1e-2
    assert C.rel_err(m.eval_esp(**kw), torch.from_numpy(fx["eval_esp"])) < 1e-4


def test_esrnerf_edge_cases():
    """all rays miss the box (empty streams through every stage, incl. an empty light-transport segment)"""
    fx, weights = C.load_esrnerf_case("lts_sparse_s220")
    m = C.build_product_esrnerf(fx, weights, DEV)
    n = 24
    rays = S.make_rays(n, 5)
    rays["rays_d"], rays["viewdirs"] = -rays["rays_d"], -rays["viewdirs"]
    batch = {k: v.to(DEV) for k, v in rays.items() if k != "rgbs"}
    out = m(s_val=220.0, uncert_masks=S.uncert_masks(n).to(DEV), normal_eps=0.01, emit_eps=0.01, **batch)
    assert (out["etc/alphainv_cum"] == 1).all() and (out["srgb/rgb"] == 0).all()
    assert out["etc/normal"].shape == (0, 3) and out["lin/pbr/off_hat"].shape == (0, 3)
    sum(v.sum() for v in out.values()).backward()
    m.eval()
    ev = m(em_modes=torch.tensor(0), pos_rt=torch.eye(3), render_pbr=True, chunk_sz=16,
           **{k: batch[k] for k in ("rays_o", "rays_d", "viewdirs")})
    assert (ev["etc/white_bg"] == 1).all() and (ev["lin/env_effects"] == 0).all()
    assert (m.eval_emit(**batch) == 0).all() and (m.eval_esp(**batch) == 0).all()


@pytest.mark.parametrize("case", C.ESRNERF_CASES)
def test_esrnerf_finetune_vs_golden(case):
    """forward_finetune (esrnerf.py:241-484) with a frozen emit_color copy that differs from emo_color: outputs, which
    parameters receive gradient (emo_rgbnet / emo_color only) and their values"""
    from oracle import esrnerf_port as E

    fx, weights = C.load_esrnerf_case(case)
    m = C.build_product_esrnerf(fx, weights, DEV)
    m.train(finetune=True)
    assert "emit_color.grid" in m.state_dict() and not m.emit_color.grid.requires_grad
    S.perturb_emit_color(m)
    m.draws = E.FixedDraws(int(fx["draw_seed"]) + 200)
    n = int(fx["n_rays"])
    rays = {k: v.to(DEV) for k, v in S.make_rays(n, int(fx["ray_seed"])).items()}
    ft_in = {k: v.to(DEV) for k, v in S.finetune_inputs(n).items()}
    out = m(rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"], **ft_in)
    assert set(out) == {"lin/pbr/emo", "lin/pbr/emo_hat"}
    for k in out:
        assert C.rel_err(out[k], torch.from_numpy(fx["ft/" + k])) < 1e-2, (k, C.rel_err(out[k], torch.from_numpy(fx["ft/" + k])))
    assert out["lin/pbr/emo"].requires_grad and not out["lin/pbr/emo_hat"].requires_grad
    cot = torch.randn(out["lin/pbr/emo"].shape, generator=torch.Generator().manual_seed(8))
    (out["lin/pbr/emo"] * cot.to(DEV)).sum().backward()
    got = {name for name, p in m.named_parameters() if p.grad is not None and bool((p.grad != 0).any())}
    want = {k[7:].rsplit("/", 1)[0] for k in fx if k.startswith("ftgrad/")}
    assert got == want, got ^ want
    for name in want:
        p = dict(m.named_parameters())[name]
        flat = p.grad.contiguous().reshape(-1).cpu()
        mx, l2 = C.grad_err(flat[torch.from_numpy(fx[f"ftgrad/{name}/idx"])], torch.from_numpy(fx[f"ftgrad/{name}/val"]))
        assert mx < 1e-2 and l2 < 1e-2, (name, mx, l2)
    m.train()
    assert "emit_color.grid" not in m.state_dict()


def test_esrnerf_full_size_properties():
    """BASELINE config 3 per-GPU shape scaled to the reference batch (lts.yaml:58-59: 8192 rays, 100 LTS points x 256
    secondary rays, 256^3 grids, sparse 100^3 mask, s_val 220): the oracle cannot run this in seconds, so check
    size-independent properties — stream sortedness, transmittance identity, output shapes / finiteness, the PDRA
    split, that every parameter receives a finite gradient and that two backward passes of the same step agree (the
    gradient sink hands each dense grid gradient over exactly once)."""
    from oracle import esrnerf_port as E

    n = 8192
    _, weights = C.load_esrnerf_case("lts_sparse_s220")
    fx = dict(mask_res=100, sparse=1, s_val=220.0, num_voxels=256 ** 3, num_2ndrays=256, num_ltspts=100, pdra_mode=1)
    m = C.build_product_esrnerf(fx, weights, DEV)
    m.keep_streams = True
    rays = {k: v.to(DEV) for k, v in S.make_rays(n, 1234).items() if k != "rgbs"}
    um = S.uncert_masks(n).to(DEV)
    grads = []
    for rep in range(2):
        m.zero_grad(set_to_none=True)
        m.draws = E.FixedDraws(5)
        out = m(s_val=220.0, uncert_masks=um, normal_eps=0.01, emit_eps=0.01, **rays)
        cot = C.esrnerf_cotangents(out)
        sum((out[k] * cot[k].to(DEV)).sum() for k in cot).backward()
        grads.append({k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None})
    st = m.last_streams["streams"]
    m3 = st.m3
    assert m3 > 10 * n / 2                                                   # the sphere is hit by most rays
    same = st.h_ray[1:] == st.h_ray[:-1]
    assert (st.h_ray[1:] >= st.h_ray[:-1]).all() and (st.h_step[1:][same] > st.h_step[:-1][same]).all()
    wsum = torch.zeros(n, device=DEV).index_add_(0, st.h_ray.long(), m.last_streams["h_w"])
    assert (wsum + out["etc/alphainv_cum"] <= 1 + 1e-4).all()
    for k, v in out.items():
        assert torch.isfinite(v).all(), k
    assert out["etc/normal"].shape == (m3, 3) and out["etc/brdf"].shape == (m3, 5) and out["etc/emit_eps"].shape == (m3, 3)
    assert out["lin/pbr/off_hat"].shape == (200, 3) and out["lin/pbr/emo"].shape == (200, 3)
    assert out["etc/emit_uncert"].shape[0] == n // 2 and out["etc/emit_cert"].shape[0] == n - n // 2
    assert (out["etc/brdf"] >= 0).all() and (out["etc/brdf"] <= 1).all() and (out["etc/emit"] >= 0).all()
    st2 = m.last_streams["lts"]["streams"]
    assert st2.n_rays == 100 * 256
    names = {k for k, p in m.named_parameters() if p.requires_grad}
    assert set(grads[0]) == names, names ^ set(grads[0])
    for k in names:
        assert torch.isfinite(grads[0][k]).all(), k
        _, l2 = C.grad_err(grads[1][k], grads[0][k])
        assert l2 < 1e-3, (k, l2)                                            # atomics reorder sums; nothing is dropped or doubled


@pytest.mark.parametrize("with_off", [True, False])
def test_lts_accumulate_kernel_vs_torch_disney(with_off):
    """esr_lts_accumulate_fwd/bwd against the torch restatement of pbr/functions.py:108-173 (esr_nerf_b200.pbr, itself
    checked against the reference through the golden LTS outputs): values and every gradient at 1e-4 (fp32)."""
    import torch.nn.functional as F

    from esr_nerf_b200 import fused, pbr

    g = torch.Generator().manual_seed(0)
    P, n2 = 37, 19
    normal = F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(DEV)
    dirs = pbr.diffuse_scattering(normal, torch.randn(P, n2, 3, generator=g).to(DEV))
    wo_a, wo_b = (F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(DEV) for _ in range(2))
    leaves = [torch.rand(P, 3, generator=g), torch.rand(P, 1, generator=g) * 0.9 + 0.05, torch.rand(P, 1, generator=g),
              torch.rand(P * n2, 3, generator=g) * 2, torch.rand(P * n2, 3, generator=g)]
    leaves[1][0] = 1e-5                                               # roughness below the r^2 clamp
    a = [t.to(DEV).requires_grad_(True) for t in leaves]
    b = [t.to(DEV).requires_grad_(True) for t in leaves]

    def ex(t, c):
        return t.view(-1, 1, c).expand(P, n2, c).flatten(0, 1)

    base, rough, metal, lo, le = a
    d_flat = dirs.flatten(0, 1)
    R = pbr.disney_reflection(ex(base, 3).repeat(2, 1), ex(rough, 1).repeat(2, 1), ex(metal, 1).repeat(2, 1),
                              ex(normal, 3).repeat(2, 1), d_flat.repeat(2, 1), torch.cat([ex(wo_a, 3), ex(wo_b, 3)], 0))
    ref_off = (lo.repeat(2, 1) * R).view(-1, n2, 3).mean(-2)
    ref_emo = (le.repeat(2, 1) * R).view(-1, n2, 3).mean(-2)
    base2, rough2, metal2, lo2, le2 = b
    off_hat, reflect = fused.LtsAccumulate.apply(base2, rough2.reshape(-1), metal2.reshape(-1), lo2 if with_off else None, le2,
                                                 normal, wo_a, wo_b, d_flat, n2)
    c1, c2 = torch.randn(2 * P, 3, generator=g).to(DEV), torch.randn(2 * P, 3, generator=g).to(DEV)
    assert C.rel_err(reflect, ref_emo) < 1e-5
    loss_ref = (ref_emo * c2).sum()
    loss = (reflect * c2).sum()
    if with_off:
        assert C.rel_err(off_hat, ref_off) < 1e-5
        loss_ref = loss_ref + (ref_off * c1).sum()
        loss = loss + (off_hat * c1).sum()
    loss_ref.backward()
    loss.backward()
    for i, (x, y) in enumerate(zip(a, b)):
        if i == 3 and not with_off:
            assert y.grad is None
            continue
        assert C.rel_err(y.grad, x.grad) < 1e-4, i


@pytest.mark.parametrize("pdra", [False, True])
def test_lts_accumulate_emission_mix_vs_torch(pdra):
    """the emo_hat mix fused into esr_lts_accumulate (esrnerf.py:668-677): emission + reflect, PDRA: emission +
    stop-gradient(reflect) on uncertain rays' points / reflect alone on the others — values and gradients vs torch"""
    import torch.nn.functional as F

    from esr_nerf_b200 import fused, pbr

    g = torch.Generator().manual_seed(1)
    P, n2 = 53, 11
    normal = F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(DEV)
    d_flat = pbr.diffuse_scattering(normal, torch.randn(P, n2, 3, generator=g).to(DEV)).flatten(0, 1)
    wo_a, wo_b = (F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(DEV) for _ in range(2))
    umask = (torch.rand(P, generator=g) < 0.5).to(DEV)
    leaves = [torch.rand(P, 3, generator=g), torch.rand(P, generator=g) * 0.9 + 0.05, torch.rand(P, generator=g),
              torch.rand(P * n2, 3, generator=g), torch.rand(P, 3, generator=g)]
    a = [t.to(DEV).requires_grad_(True) for t in leaves]
    b = [t.to(DEV).requires_grad_(True) for t in leaves]
    cot = torch.randn(2 * P, 3, generator=g).to(DEV)
    base, rough, metal, le, emission = a
    _, reflect = fused.LtsAccumulate.apply(base, rough, metal, None, le, normal, wo_a, wo_b, d_flat, n2)
    if pdra:
        want = torch.where(umask.repeat(2)[:, None], emission.repeat(2, 1) + reflect.detach(), reflect)
    else:
        want = emission.repeat(2, 1) + reflect
    (want * cot).sum().backward()
    base2, rough2, metal2, le2, emission2 = b
    _, got = fused.LtsAccumulate.apply(base2, rough2, metal2, None, le2, normal, wo_a, wo_b, d_flat, n2, emission2, umask, pdra)
    (got * cot).sum().backward()
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)      # (the kernel's mean-and-add is one fused multiply-add)
    for i, (x, y) in enumerate(zip(a, b)):
        assert torch.allclose(y.grad, x.grad, rtol=1e-5, atol=1e-7), i   # same kernel arithmetic; the mix only gates it


@pytest.mark.parametrize("activation", ["softplus", "relu", "abs", "exp", "sigmoid"])
def test_sg_envmap_kernel_vs_module(activation):
    """esr_sg_envmap_fwd/bwd (environment map of the secondary rays with `* T_last` and `+ off_m` fused in,
    esrnerf.py:560-566) against SphericalGaussian.forward (pbr/module.py:133-143, torch): values 1e-5, gradients of
    mus / lambdas / lobes / T_last / off_m 1e-4"""
    import torch.nn.functional as F

    from esr_nerf_b200 import fused
    from esr_nerf_b200.modules import SphericalGaussian

    torch.manual_seed(3)
    env_a, env_b = SphericalGaussian(48, activation).to(DEV), SphericalGaussian(48, activation).to(DEV)
    env_b.load_state_dict(env_a.state_dict())
    g = torch.Generator().manual_seed(4)
    M = 5000
    dirs = F.normalize(torch.randn(M, 3, generator=g), dim=-1).to(DEV)
    last_a, add_a = torch.rand(M, generator=g).to(DEV).requires_grad_(True), torch.rand(M, 3, generator=g).to(DEV).requires_grad_(True)
    last_b, add_b = last_a.detach().clone().requires_grad_(True), add_a.detach().clone().requires_grad_(True)
    cot = torch.randn(M, 3, generator=g).to(DEV)
    want = add_a + env_a(dirs) * last_a.unsqueeze(-1)
    (want * cot).sum().backward()
    lobes = F.normalize(env_b.lobes, dim=-1)
    got = fused.SgEnvmap.apply(dirs, env_b.mus, torch.abs(env_b.lambdas).reshape(-1), lobes, fused.SG_ACT_IDS[activation],
                               last_b, add_b)
    (got * cot).sum().backward()
    assert C.rel_err(got, want) < 1e-5
    for name in ("mus", "lambdas", "lobes"):
        assert C.rel_err(getattr(env_b, name).grad, getattr(env_a, name).grad) < 1e-4, name
    assert C.rel_err(last_b.grad, last_a.grad) < 1e-4 and torch.equal(add_b.grad, add_a.grad)


@pytest.mark.parametrize("fib", [False, True])
def test_lts_scatter_dirs_kernel_vs_torch(fib):
    """esr_lts_scatter_dirs against pbr.diffuse_scattering / diffuse_scattering_fib (pbr/functions.py:10-32)"""
    import torch.nn.functional as F

    from esr_nerf_b200 import fused, pbr

    g = torch.Generator().manual_seed(6)
    P, n = 301, 257
    normal = F.normalize(torch.randn(P, 3, generator=g), dim=-1).to(DEV)
    if fib:
        want = pbr.diffuse_scattering_fib(normal, n)
        got = fused.lts_scatter_dirs(normal, n, table=pbr.fibonacci_hemisphere(n).to(DEV))
        assert torch.equal(got, want)
    else:
        noise = torch.randn(P, n, 3, generator=g).to(DEV)
        want = pbr.diffuse_scattering(normal, noise)
        got = fused.lts_scatter_dirs(normal, n, noise=noise)
        # x / max(|x|, eps) vs x * (1 / max(|x|, eps)): one rounding apart
        assert torch.allclose(got, want, rtol=0, atol=2e-7) and torch.equal(torch.sign(got), torch.sign(want))
    assert ((got * normal[:, None]).sum(-1) >= 0).all()

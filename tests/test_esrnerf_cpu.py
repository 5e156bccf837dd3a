"""CPU suite for the LTS / PDRA stage oracle: oracle/esrnerf_port.py (the travelling restatement of
ESRNeRF.forward_training, esrnerf.py:487-851) against the golden vectors made by the reference's own class, and —
when /root/reference is present — against the reference class itself with the reference's own RNG calls."""
import numpy as np
import pytest
import torch

import esr_testlib as C


@pytest.mark.parametrize("case", C.ESRNERF_CASES)
def test_esrnerf_port_matches_golden(case):
    fx, weights = C.load_esrnerf_case(case)
    out, inter, leaves, _ = C.run_esrnerf_port(fx, weights)
    assert set(out) == {k[4:] for k in fx if k.startswith("out/")}
    for k in out:
        assert tuple(out[k].shape) == fx["out/" + k].shape, k            # same shaded-sample count: streams match
        assert C.rel_err(out[k], torch.from_numpy(fx["out/" + k])) < 1e-5, k
    cot = C.esrnerf_cotangents(out)
    loss = sum((out[k] * cot[k]).sum() for k in cot)
    loss.backward()
    assert abs(loss.item() - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    checked = 0
    for name, leaf in leaves.items():
        if f"grad/{name}/idx" in fx:
            assert leaf.grad is not None, name
            err, s_err = C.digest_check(fx, name, leaf.grad, rtol=1e-4)
            assert err < 1.0 and s_err < 1e-4, (name, err, s_err)
            checked += 1
    assert checked >= 40        # 5 grids/nets x layers + envmap: every parameter of the stage receives a gradient


def test_esrnerf_port_matches_reference():
    """same weights, rays and RNG state: the port draws with the reference's own calls (np.random.choice,
    torch.randn) in the reference's order, so outputs agree to rounding"""
    from oracle import ref_harness as H

    if not H.reference_available():
        pytest.skip("/root/reference not present (GPU box): golden vectors stand in")
    from esr_nerf_b200 import synthetic as S
    from oracle import esrnerf_port as E
    from oracle.make_golden import build_reference_esrnerf

    fx, weights = C.load_esrnerf_case("pdra_sparse_s60")
    n, s_val = 96, 35.0
    ref = build_reference_esrnerf(int(fx["num_voxels"]), int(fx["mask_res"]), True, s_val, weights, num_2ndrays=8,
                                  num_ltspts=16)
    ref.pdra_mode = True
    rays = S.make_rays(n, 4242)          # rays the fixtures have never seen
    um = S.uncert_masks(n)
    np.random.seed(5)
    torch.manual_seed(11)
    ref_out = ref(s_val=s_val, rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"],
                  em_modes=rays["em_modes"], uncert_masks=um, normal_eps=0.01, emit_eps=0.03)
    scene = C.esrnerf_oracle_scene(dict(fx, num_2ndrays=8, num_ltspts=16))
    params, leaves = C.esrnerf_oracle_params(scene, weights)
    np.random.seed(5)
    torch.manual_seed(11)
    out, _ = E.esrnerf_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                        rays["em_modes"], um, s_val, 0.01, 0.03, True, E.Draws())
    cot = C.esrnerf_cotangents(out)
    sum((ref_out[k] * cot[k]).sum() for k in cot).backward()
    sum((out[k] * cot[k]).sum() for k in cot).backward()
    assert set(out) == set(ref_out)
    for k in ref_out:
        assert C.rel_err(out[k], ref_out[k]) < 1e-6, k
    for name, p in ref.named_parameters():
        if p.grad is not None:
            assert C.rel_err(leaves[name].grad, p.grad) < 1e-5, name


def test_esrnerf_port_matches_reference_neus_alpha_grad():
    """`neus_alpha: grad` (functions.py:45-69, esrnerf.py:197-200): the port against the reference's own ESRNeRF built with
    that option — the LTS / PDRA training step (primary AND secondary rays take the view-projected SDF gradient), eval_emit
    and eval_esp."""
    from oracle import ref_harness as H

    if not H.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    from esr_nerf_b200 import synthetic as S
    from oracle import esrnerf_port as E
    from oracle.make_golden import build_reference_esrnerf

    fx, weights = C.load_esrnerf_case("pdra_sparse_s60")
    n, s_val = 96, 35.0
    ref = build_reference_esrnerf(int(fx["num_voxels"]), int(fx["mask_res"]), True, s_val, weights, num_2ndrays=8,
                                  num_ltspts=16, neus_alpha="grad")
    ref.pdra_mode = True
    rays = S.make_rays(n, 4243)
    um = S.uncert_masks(n)
    np.random.seed(5)
    torch.manual_seed(11)
    ref_out = ref(s_val=s_val, rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"],
                  em_modes=rays["em_modes"], uncert_masks=um, normal_eps=0.01, emit_eps=0.03)
    scene = C.esrnerf_oracle_scene(dict(fx, num_2ndrays=8, num_ltspts=16))
    scene["neus_alpha"] = "grad"
    params, leaves = C.esrnerf_oracle_params(scene, weights)
    np.random.seed(5)
    torch.manual_seed(11)
    out, inter = E.esrnerf_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                            rays["em_modes"], um, s_val, 0.01, 0.03, True, E.Draws())
    a_interp = E.P.neus_alpha_interp(inter["m1_ray"], inter["m1_sdf"].detach(), s_val)
    assert (a_interp - inter["m1_alpha"].detach()).abs().max() > 1e-3       # really the other function
    cot = C.esrnerf_cotangents(out)
    sum((ref_out[k] * cot[k]).sum() for k in cot).backward()
    sum((out[k] * cot[k]).sum() for k in cot).backward()
    assert set(out) == set(ref_out)
    for k in ref_out:
        assert C.rel_err(out[k], ref_out[k]) < 1e-6, k
    for name, p in ref.named_parameters():
        if p.grad is not None:
            assert C.rel_err(leaves[name].grad, p.grad) < 1e-5, name
    ref.eval()
    params0, _ = C.esrnerf_oracle_params(scene, weights, requires_grad=False)
    args = (scene, params0, rays["rays_o"], rays["rays_d"], rays["viewdirs"])
    with torch.no_grad():
        r_emit = ref.eval_emit(rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"])
        r_esp = ref.eval_esp(rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"])
    assert C.rel_err(E.esrnerf_eval_emit(*args, s_val)[0], r_emit) < 1e-5
    assert C.rel_err(E.esrnerf_eval_esp(*args, s_val)[0], r_esp) < 1e-5


@pytest.mark.parametrize("case", C.ESRNERF_CASES)
def test_esrnerf_eval_ports_match_golden(case):
    """forward_evaluate (with the PBR decomposition, several LTS chunks), eval_emit, eval_esp: port vs the outputs of
    the reference's own methods stored in the fixture"""
    from esr_nerf_b200 import synthetic as S
    from oracle import esrnerf_port as E

    fx, weights = C.load_esrnerf_case(case)
    scene = C.esrnerf_oracle_scene(fx)
    params, _ = C.esrnerf_oracle_params(scene, weights, requires_grad=False)
    rays = S.make_rays(int(fx["n_rays"]), int(fx["ray_seed"]))
    args = (scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"])
    s_val = float(fx["s_val"])
    for em in (0, 1):
        out, _ = E.esrnerf_forward_evaluate(*args, torch.tensor(em), torch.from_numpy(fx["pos_rt"]), s_val, True,
                                            int(fx["eval_chunk"]), E.FixedDraws(int(fx["draw_seed"]) + 100))
        keys = {k.split("/", 1)[1] for k in fx if k.startswith(f"eval{em}/")}
        assert set(out) == keys and len(keys) == 21
        for k in out:
            assert C.rel_err(out[k], torch.from_numpy(fx[f"eval{em}/{k}"])) < 1e-5, (em, k)
    assert C.rel_err(E.esrnerf_eval_emit(*args, s_val)[0], torch.from_numpy(fx["eval_emit"])) < 1e-5
    assert C.rel_err(E.esrnerf_eval_esp(*args, s_val)[0], torch.from_numpy(fx["eval_esp"])) < 1e-5


def _finetune_port(fx, weights):
    from esr_nerf_b200 import synthetic as S
    from oracle import esrnerf_port as E

    scene = C.esrnerf_oracle_scene(fx)
    params, leaves = C.esrnerf_oracle_params(scene, weights)
    g = torch.Generator().manual_seed(9)               # synthetic.perturb_emit_color
    params["emit_color"] = (leaves["emo_color.grid"].detach() + 0.05 * torch.randn(leaves["emo_color.grid"].shape, generator=g))
    n = int(fx["n_rays"])
    rays = S.make_rays(n, int(fx["ray_seed"]))
    ft_in = S.finetune_inputs(n)
    out = E.esrnerf_forward_finetune(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"], ft_in["em_modes"],
                                     ft_in["em_intensities"], ft_in["em_colors"], float(fx["s_val"]),
                                     E.FixedDraws(int(fx["draw_seed"]) + 200))
    return out, leaves


@pytest.mark.parametrize("case", C.ESRNERF_CASES)
def test_esrnerf_finetune_port_matches_golden(case):
    """forward_finetune (esrnerf.py:241-484): outputs and the only live gradient path (emo_rgbnet / emo_color, Q14)"""
    fx, weights = C.load_esrnerf_case(case)
    out, leaves = _finetune_port(fx, weights)
    for k in out:
        assert C.rel_err(out[k], torch.from_numpy(fx["ft/" + k])) < 1e-5, k
    assert out["lin/pbr/emo"].requires_grad and not out["lin/pbr/emo_hat"].requires_grad
    cot = torch.randn(out["lin/pbr/emo"].shape, generator=torch.Generator().manual_seed(8))
    (out["lin/pbr/emo"] * cot).sum().backward()
    with_grad = {k for k, l in leaves.items() if l.grad is not None}
    assert with_grad == {k[7:].rsplit("/", 1)[0] for k in fx if k.startswith("ftgrad/")}
    ft = {k[2:]: v for k, v in fx.items() if k.startswith("ftgrad/")}
    for name in with_grad:
        err, s_err = C.digest_check(ft, name, leaves[name].grad, rtol=1e-4)
        assert err < 1.0 and s_err < 1e-4, (name, err, s_err)


@pytest.mark.parametrize("ray_sampling,env_activation", [("fib", "softplus"), ("random", "relu"), ("fibonacci", "sigmoid")])
def test_esrnerf_port_matches_reference_other_samplers_and_envmaps(ray_sampling, env_activation):
    """the configurable pieces no shipped config uses (esrnerf.py:188-195): the Fibonacci hemisphere sampler (no random
    draw for the directions) and the other environment-map activations — port vs the reference's own class on the CPU,
    training step and finetune target"""
    from oracle import ref_harness as H

    if not H.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    from esr_nerf_b200 import synthetic as S
    from oracle import esrnerf_port as E
    from oracle.make_golden import build_reference_esrnerf

    fx, weights = C.load_esrnerf_case("pdra_sparse_s60")
    n, s_val = 64, 35.0
    ref = build_reference_esrnerf(int(fx["num_voxels"]), int(fx["mask_res"]), True, s_val, weights, num_2ndrays=8,
                                  num_ltspts=16, ray_sampling=ray_sampling, env_activation=env_activation)
    ref.pdra_mode = True
    rays = S.make_rays(n, 777)
    um = S.uncert_masks(n)
    np.random.seed(6)
    torch.manual_seed(12)
    ref_out = ref(s_val=s_val, rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"],
                  em_modes=rays["em_modes"], uncert_masks=um, normal_eps=0.01, emit_eps=0.03)
    scene = C.esrnerf_oracle_scene(dict(fx, num_2ndrays=8, num_ltspts=16))
    scene.update(ray_sampling=ray_sampling, env_activation=env_activation)
    params, leaves = C.esrnerf_oracle_params(scene, weights)
    np.random.seed(6)
    torch.manual_seed(12)
    out, _ = E.esrnerf_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                        rays["em_modes"], um, s_val, 0.01, 0.03, True, E.Draws())
    cot = C.esrnerf_cotangents(out)
    sum((ref_out[k] * cot[k]).sum() for k in cot).backward()
    sum((out[k] * cot[k]).sum() for k in cot).backward()
    assert set(out) == set(ref_out)
    for k in ref_out:
        assert C.rel_err(out[k], ref_out[k]) < 1e-6, k
    for name, p in ref.named_parameters():
        if p.grad is not None:
            assert C.rel_err(leaves[name].grad, p.grad) < 1e-5, name


def test_esrnerf_port_matches_reference_on_a_non_cubic_box(monkeypatch):
    """The LTS / PDRA training step in a 46 x 40 x 33 box (every fixture is a cube): port vs the reference's own class,
    same RNG calls — what tests/test_gpu_esrnerf.py::test_esrnerf_port_as_live_oracle_on_new_rays[non_cubic] then holds
    the CUDA path to."""
    from oracle import ref_harness as H

    if not H.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    from esr_nerf_b200 import synthetic as S
    from oracle import esrnerf_port as E
    from oracle.make_golden import build_reference_esrnerf

    monkeypatch.setattr(S, "BBOX_MIN", torch.tensor([-1.05, -0.9, -0.75]))
    monkeypatch.setattr(S, "BBOX_MAX", torch.tensor([1.05, 0.9, 0.75]))
    fx, weights = C.load_esrnerf_case("pdra_sparse_s60")
    fx = dict(fx, num_voxels=46 * 40 * 33, mask_res=20, sparse=0)
    n, s_val = 96, 35.0
    ref = build_reference_esrnerf(int(fx["num_voxels"]), int(fx["mask_res"]), False, s_val, weights, num_2ndrays=8,
                                  num_ltspts=16)
    assert len({int(w) for w in ref.world_size}) == 3, ref.world_size
    ref.pdra_mode = True
    rays = S.make_rays(n, 4243)
    um = S.uncert_masks(n)
    np.random.seed(5)
    torch.manual_seed(11)
    ref_out = ref(s_val=s_val, rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"],
                  em_modes=rays["em_modes"], uncert_masks=um, normal_eps=0.01, emit_eps=0.03)
    scene = C.esrnerf_oracle_scene(dict(fx, num_2ndrays=8, num_ltspts=16))
    assert scene["world_size"] == [int(w) for w in ref.world_size]
    params, leaves = C.esrnerf_oracle_params(scene, weights)
    np.random.seed(5)
    torch.manual_seed(11)
    out, _ = E.esrnerf_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                        rays["em_modes"], um, s_val, 0.01, 0.03, True, E.Draws())
    cot = C.esrnerf_cotangents(out)
    sum((ref_out[k] * cot[k]).sum() for k in cot).backward()
    sum((out[k] * cot[k]).sum() for k in cot).backward()
    assert set(out) == set(ref_out)
    for k in ref_out:
        assert C.rel_err(out[k], ref_out[k]) < 1e-6, k
    for name, p in ref.named_parameters():
        if p.grad is not None:
            assert C.rel_err(leaves[name].grad, p.grad) < 1e-5, name

"""GPU parity (-m gpu) of the coarse-stage render path (esr_nerf_b200.VoxurfC, BASELINE config 1 shape) against the
golden vectors produced by the reference's own VoxurfC (tests/golden/voxurfc_*.npz) and the oracle port
(oracle/voxurfc_port.py, pinned against the reference in tests/test_oracle_cpu.py).  Sample streams bit-exact.  The two
colour MLPs run in the default `x2` mode on the tcgen05 chains (zero-padded, identity third hidden layer): outputs 1e-4
(measured ~1e-6), every parameter gradient within 1e-2 max-norm and relative L2; `torch_fp32` (library GEMMs) keeps the
1e-4 class on everything (isolated ReLU-boundary flips tolerated as documented in esr_testlib.grad_close)."""
import pytest
import torch

import esr_testlib as C
from esr_nerf_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT_KEYS = ("etc/alphainv_cum", "etc/white_bg", "srgb/rgb")


def _run(fx, weights, rays=None, mode="x2"):
    m = C.build_product_coarse(fx, weights, DEV)
    m.mlp_mode = mode
    m.keep_streams = True
    if rays is None:
        rays = S.make_rays(int(fx["n_rays"]), int(fx["ray_seed"]))
    n = rays["rays_o"].shape[0]
    out = m(s_val=float(fx["s_val"]), **{k: v.to(DEV) for k, v in rays.items()})
    cot = C.coarse_cotangents(n)
    sum((out[k] * cot[k].to(DEV)).sum() for k in cot).backward()
    return m, out


@pytest.mark.parametrize("mode", ["x2", "torch_fp32"])
@pytest.mark.parametrize("case", C.COARSE_CASES)
def test_coarse_vs_golden(case, mode):
    fx, weights = C.load_coarse_case(case)
    m, out = _run(fx, weights, mode=mode)
    assert set(out) == set(OUT_KEYS)
    for k in OUT_KEYS:
        assert out[k].shape == fx["out/" + k].shape, k
        assert C.rel_err(out[k], torch.from_numpy(fx["out/" + k])) < 1e-4, k
    checked = 0
    for name, p in m.named_parameters():
        if f"grad/{name}/idx" not in fx:
            continue
        assert p.grad is not None, name
        flat = p.grad.contiguous().reshape(-1).cpu()
        got, ref = flat[torch.from_numpy(fx[f"grad/{name}/idx"])], torch.from_numpy(fx[f"grad/{name}/val"])
        if mode == "x2":      # tensor-core colour nets: relative L2 < 1e-2 on every tensor; max-norm 1e-2 up to isolated
            # ReLU-boundary flips (with ~10^4 samples ONE flipped mask — fp32 summation order does that to the reference's
            # own two devices — moves a weight-gradient row by a per-cent of the tensor's maximum: esr_testlib.grad_close)
            ok, msg = C.grad_close(got, ref, 1e-2, l2_factor=1.0)
            assert ok, (name, msg)
        else:
            ok, msg = C.grad_close(got, ref, 1e-4)
            assert ok, (name, msg)
        checked += 1
    assert checked == 3 + 6 + 6


def test_coarse_vs_oracle_port_config1_shape():
    """BASELINE config 1: 4096 rays through a 64^3 coarse model (~128 candidate samples per ray)."""
    from oracle import voxurfc_port as PC

    _, weights = C.load_coarse_case("coarse_sparse_s5")
    fx = dict(num_voxels=64 ** 3, mask_res=32, sparse=1, s_val=5.0)
    n = 4096
    rays = S.make_rays(n, 2718)
    scene = C.coarse_oracle_scene(64 ** 3, 32, True)
    params, leaves = C.coarse_oracle_params(scene, weights)
    ref, inter = PC.voxurfc_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                             rays["em_modes"], 5.0)
    cot = C.coarse_cotangents(n)
    sum((ref[k] * cot[k]).sum() for k in cot).backward()
    m, out = _run(fx, weights, rays, mode="torch_fp32")
    st = m.last_streams["streams"]
    assert torch.equal(st.h_ray.long().cpu(), inter["m3_ray"]) and torch.equal(st.h_step.long().cpu(), inter["m3_step"])
    assert C.rel_err(m.last_streams["h_w"], inter["m3_weights"]) < 1e-4
    for k in OUT_KEYS:
        assert C.rel_err(out[k], ref[k]) < 1e-4, k
    for name, p in m.named_parameters():
        if name in leaves and leaves[name].grad is not None:
            # MLP weight gradients are fp32 library GEMMs reducing ~10^5 random-signed per-sample terms (cuBLAS on the
            # GPU, MKL in the oracle): the two summation orders differ by a few 1e-4 of the largest entry
            tol = 5e-4 if "rgbnet" in name else 1e-4
            ok, msg = C.grad_close(p.grad.contiguous(), leaves[name].grad, tol)
            assert ok, (name, msg)


def test_coarse_all_rays_miss():
    fx, weights = C.load_coarse_case("coarse_sparse_s5")
    m = C.build_product_coarse(fx, weights, DEV)
    rays = S.make_rays(17, 5)
    rays["rays_d"], rays["viewdirs"] = -rays["rays_d"], -rays["viewdirs"]
    out = m(s_val=5.0, **{k: v.to(DEV) for k, v in rays.items()})
    assert (out["etc/alphainv_cum"] == 1).all() and (out["srgb/rgb"] == 0).all() and (out["etc/white_bg"] == 1).all()


@pytest.mark.parametrize("mode", ["x2", "torch_fp32"])
@pytest.mark.parametrize("em", [0, 1])
def test_coarse_forward_evaluate_vs_oracle_port(em, mode):
    """voxurfc.py:273-424: 8 inference maps vs the oracle port (pinned against the reference on the CPU)."""
    from oracle import voxurfc_port as PC

    fx, weights = C.load_coarse_case("coarse_dense_s25")
    rays = S.make_rays(700, 77)
    pos_rt = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(3)))[0]
    scene = C.coarse_oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), bool(fx["sparse"]))
    params, _ = C.coarse_oracle_params(scene, weights, requires_grad=False)
    with torch.no_grad():
        ref, inter = PC.voxurfc_forward_evaluate(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                                 torch.tensor(em), pos_rt, float(fx["s_val"]))
    m = C.build_product_coarse(fx, weights, DEV)
    m.mlp_mode = mode
    m.keep_streams = True
    m.eval()
    out = m(rays_o=rays["rays_o"].to(DEV), rays_d=rays["rays_d"].to(DEV), viewdirs=rays["viewdirs"].to(DEV),
            em_modes=torch.tensor(em), pos_rt=pos_rt.to(DEV))
    assert set(out) == set(ref)
    st = m.last_streams["streams"]
    assert torch.equal(st.h_ray.long().cpu(), inter["m3_ray"]) and torch.equal(st.h_step.long().cpu(), inter["m3_step"])
    for k in ref:
        assert out[k].shape == ref[k].shape, k
        # inference on the tensor-core path uses single bf16 operands (as VoxurfF's): 1e-2 on the colour maps
        tol = 1e-2 if (mode != "torch_fp32" and k.startswith("srgb/")) else 1e-4
        assert C.rel_err(out[k], ref[k]) < tol, (k, C.rel_err(out[k], ref[k]))

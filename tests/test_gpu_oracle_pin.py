"""Oracle pin (-m gpu): the reference's OWN native kernels (app/utils/base/cuda/render_utils_kernel.cu,
total_variation_kernel.cu — compiled unmodified from /root/reference into oracle/_ref/ by oracle/build_ref.py,
prebuilt .so files travel to the GPU box) run on the B200 next to
  (1) the C restatement oracle/render_utils_ref.c  -> pins the oracle, and
  (2) the product's replacements behind the C ABI   -> parity against the real reference, not a restatement.
Integer / index / mask outputs and the sequential transmittance recurrence must be bit-exact."""
import pytest
import torch

import esr_testlib as C
from esr_nerf_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref

    if not build_ref.available():
        pytest.skip("oracle/_ref not built (python -m oracle.build_ref where /root/reference exists)")
    return build_ref.load_module("ref_render_utils_cuda"), build_ref.load_module("ref_total_variation_cuda")


def _rays(n, seed):
    r = S.make_rays(n, seed)
    o, d = r["rays_o"].clone(), r["rays_d"].clone()
    k = max(n // 10, 1)
    d[:k] = -d[:k]                       # misses
    if n >= 16:
        d[k:k + 3, 0] = 0.0              # exact zero components (1e-6 substitution, kernel.cu:23-25)
        d[k + 3:k + 5, 1] = 0.0
    return o.contiguous(), d.contiguous()


@pytest.mark.parametrize("n,stepdist,near", [(513, 0.021, 2.0), (4096, 0.0041, 2.0), (64, 0.05, 1e-5)])
def test_sample_pts_on_rays_three_way(ref, n, stepdist, near):
    from esr_nerf_b200 import render_utils as R
    from oracle import ref_harness as H

    rk, _ = ref
    o, d = _rays(n, 7 + n)
    mn, mx = S.BBOX_MIN, S.BBOX_MAX
    real = rk.sample_pts_on_rays(o.to(DEV), d.to(DEV), mn.to(DEV), mx.to(DEV), near, 1e9, stepdist)
    torch.cuda.synchronize()
    oracle = H.sample_pts_on_rays(o, d, mn, mx, near, 1e9, stepdist)
    prod = R.render_utils_cuda.sample_pts_on_rays(o.to(DEV), d.to(DEV), mn.to(DEV), mx.to(DEV), near, 1e9, stepdist)
    names = ["ray_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"]
    for name, a, b, c in zip(names, real, oracle, prod):
        assert a.dtype == c.dtype and a.shape == c.shape, name
        assert torch.equal(a.cpu(), b), f"C oracle != reference kernel: {name}"
        assert torch.equal(a, c), f"product != reference kernel: {name}"


@pytest.mark.parametrize("n_rays,seed", [(300, 1), (8192, 2)])
def test_alpha2weight_three_way(ref, n_rays, seed):
    from esr_nerf_b200 import render_utils as R
    from oracle import ref_harness as H
    from test_gpu_native_ops import _alpha_stream

    rk, _ = ref
    alpha, ray_id = _alpha_stream(n_rays, seed)
    a_d, r_d = alpha.to(DEV), ray_id.to(DEV)
    real = rk.alpha2weight(a_d, r_d, n_rays)
    torch.cuda.synchronize()
    oracle = H.alpha2weight(alpha, ray_id, n_rays)
    prod = R.render_utils_cuda.alpha2weight(a_d, r_d, n_rays)
    for name, a, b, c in zip(["weight", "T", "alphainv_last", "i_start", "i_end"], real, oracle, prod):
        assert torch.equal(a.cpu(), b), f"C oracle != reference kernel: {name}"
        assert torch.equal(a, c), f"product != reference kernel: {name}"
    g = torch.Generator().manual_seed(seed + 5)
    gw, gl = torch.randn(alpha.shape, generator=g), torch.randn(n_rays, generator=g)
    real_g = rk.alpha2weight_backward(a_d, *real, n_rays, gw.to(DEV), gl.to(DEV))
    torch.cuda.synchronize()
    oracle_g = H.alpha2weight_backward(alpha, *oracle, n_rays, gw, gl)
    prod_g = R.render_utils_cuda.alpha2weight_backward(a_d, *prod, n_rays, gw.to(DEV), gl.to(DEV))
    assert torch.equal(real_g.cpu(), oracle_g), "C oracle != reference kernel: grad (same sequential order)"
    assert C.rel_err(prod_g, real_g) < 1e-5          # warp suffix scan re-associates the running sum


def test_total_variation_three_way(ref):
    from esr_nerf_b200 import render_utils as R
    from oracle import ref_harness as H

    _, tv = ref
    g = torch.Generator().manual_seed(5)
    p = torch.randn(1, 1, 17, 19, 23, generator=g) * 2
    grad = torch.randn(p.shape, generator=g)
    real = grad.to(DEV)
    tv.total_variation_add_grad(p.to(DEV), real, 0.3, 0.3, 0.3, True)
    torch.cuda.synchronize()
    oracle = grad.clone()
    H.total_variation_add_grad(p, oracle, 0.3, 0.3, 0.3, True)
    prod = grad.to(DEV)
    R.total_variation_cuda.total_variation_add_grad(p.to(DEV), prod, 0.3, 0.3, 0.3, True)
    assert C.rel_err(oracle, real) < 1e-6
    assert torch.equal(prod, real)                   # same expression order -> bit-exact
    sp = grad.clone()
    sp[0, 0, :6] = 0
    real_s, prod_s = sp.to(DEV), sp.to(DEV)
    tv.total_variation_add_grad(p.to(DEV), real_s, 0.3, 0.3, 0.3, False)
    R.total_variation_cuda.total_variation_add_grad(p.to(DEV), prod_s, 0.3, 0.3, 0.3, False)
    torch.cuda.synchronize()
    assert torch.equal(prod_s, real_s)

"""GPU parity (-m gpu) of the dense-grid loss terms (csrc/regularizers.cu) against the reference's
torch formulation evaluated on the same device tensors: total_variation (app/utils/base/functions.py:34-42) on the SDF
grid and on channels-last colour grids, neus_sdf_gradient (voxurff.py:723-742), the smooth-gradient term
(voxurff.py:610-616) — values and gradients.  The torch formulation itself is pinned to the reference's methods on the CPU
(tests/test_oracle_cpu.py::test_grid_regularizers_match_reference)."""
import pytest
import torch

from esr_nerf_b200 import fused
from esr_nerf_b200 import synthetic as S
from esr_nerf_b200.modules import GradientConv, total_variation

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape,channels,masked", [((33, 20, 27), 1, True), ((33, 20, 27), 1, False), ((16, 24, 20), 6, True),
                                                   ((64, 64, 64), 12, True)])
def test_grid_tv_kernel_vs_torch(shape, channels, masked):
    g = torch.Generator().manual_seed(sum(shape) + channels)
    v = torch.randn(1, channels, *shape, generator=g).to(DEV)
    if channels > 1:
        v = v.contiguous(memory_format=torch.channels_last_3d)      # the colour grids' memory layout
    mask = (torch.rand(1, 1, *shape, generator=g) < 0.6).to(DEV) if masked else None
    a, b = v.clone().requires_grad_(True), v.clone().requires_grad_(True)
    want = total_variation(a, mask.repeat(1, channels, 1, 1, 1) if masked else None)
    got = fused.GridTV.apply(b, mask)
    assert abs(float(got.detach()) - float(want.detach())) <= 2e-6 * abs(float(want.detach()))
    (want * 3.7).backward()
    (got * 3.7).backward()
    assert b.grad.stride() == b.stride()
    assert _rel(b.grad, a.grad) < 1e-5


def test_sdf_central_gradient_and_smooth_term_vs_torch():
    shape, h = (30, 26, 22), 0.0371
    g = torch.Generator().manual_seed(5)
    sdf = torch.randn(1, 1, *shape, generator=g).to(DEV)
    mask = (torch.rand(1, 1, *shape, generator=g) < 0.5).to(DEV)
    conv = GradientConv(sigma=0.8).to(DEV)

    def torch_gradient(grid):
        out = torch.zeros([1, 3, *grid.shape[-3:]], device=grid.device)
        out[:, 0, 1:-1, :, :] = (grid[:, 0, 2:, :, :] - grid[:, 0, :-2, :, :]) / 2 / h
        out[:, 1, :, 1:-1, :] = (grid[:, 0, :, 2:, :] - grid[:, 0, :, :-2, :]) / 2 / h
        out[:, 2, :, :, 1:-1] = (grid[:, 0, :, :, 2:] - grid[:, 0, :, :, :-2]) / 2 / h
        return out

    a, b = sdf.clone().requires_grad_(True), sdf.clone().requires_grad_(True)
    want_g, got_g = torch_gradient(a), fused.SdfCentralGradient.apply(b, h)
    # (torch turns a division by a Python scalar into a multiplication by its reciprocal; the kernel divides, as the model's
    # tensor-valued voxel_size makes torch do: last-place differences here, none in the model-level test below)
    assert torch.allclose(got_g, want_g, rtol=1e-6, atol=0)
    cot = torch.randn(1, 3, *shape, generator=g).to(DEV)
    (want_g * cot).sum().backward()
    (got_g * cot).sum().backward()
    assert _rel(b.grad, a.grad) < 1e-5

    a, b = sdf.clone().requires_grad_(True), sdf.clone().requires_grad_(True)
    grad = torch_gradient(a).permute(1, 0, 2, 3, 4)
    err = conv(grad).detach() - grad
    want = (err[mask.repeat(3, 1, 1, 1, 1)] ** 2).mean()
    got = fused.SmoothGradTV.apply(b, mask, conv.m.weight.detach(), float(conv.m.bias[0]), h)
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    (want * 0.3).backward()
    (got * 0.3).backward()
    assert _rel(b.grad, a.grad) < 1e-4


def test_model_regularizers_on_device_match_host_formulation():
    """VoxurfF.density_total_variation / neus_sdf_gradient on cuda:0 (kernels) against the same model on the CPU (the torch
    formulation the CPU tests pin to the reference): loss values and the SDF-grid gradient"""
    from esr_nerf_b200.voxurff import VoxurfF

    def make(device):
        torch.manual_seed(0)
        geo = (S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT, S.mask_density(12, True))
        m = VoxurfF(S.fine_cfg(device), *geo, 20.0, 24 ** 3)
        S.fill_fine_model(m)
        return m

    cpu, gpu = make("cpu"), make(DEV)
    gpu.load_state_dict(cpu.state_dict(), strict=True)
    gpu.nonempty_mask = cpu.nonempty_mask.to(DEV)
    for kw in (dict(sdf_tv=0.1), dict(smooth_grad_tv=0.05), dict(sdf_tv=0.1, smooth_grad_tv=0.05)):
        for m in (cpu, gpu):
            m.sdf.grid.grad = None
        lc, lg = cpu.density_total_variation(**kw), gpu.density_total_variation(**kw)
        assert abs(float(lg) - float(lc)) <= 1e-4 * abs(float(lc)), kw
        lc.backward()
        lg.backward()
        assert _rel(gpu.sdf.grid.grad.cpu(), cpu.sdf.grid.grad) < 1e-3, kw
    assert torch.allclose(gpu.neus_sdf_gradient().cpu(), cpu.neus_sdf_gradient(), rtol=1e-5, atol=1e-6)

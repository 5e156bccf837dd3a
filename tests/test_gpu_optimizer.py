"""GPU parity (-m gpu) of the fused Adam step (esr_nerf_b200.optimizer.Adam -> esr_adam_step) against the oracle port of
app/utils/optimizer.py:186-228 (pinned bit-exactly to the reference's own class on the CPU, tests/test_oracle_cpu.py).
Tolerance 5e-6 relative after 7 steps: same fp32 operation order; the device contracts multiply-adds the CPU rounds twice
(as torch's own CUDA kernels do), ~1 ulp per step."""
import pytest
import torch

import esr_testlib as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,cl,per_lr,wd", [((1, 1, 9, 7, 5), False, True, 0.0), ((1, 6, 8, 6, 10), True, False, 0.0),
                                               ((192, 85), False, False, 0.01), ((3,), False, False, 0.0)])
def test_fused_adam_matches_port(shape, cl, per_lr, wd):
    from esr_nerf_b200.optimizer import Adam
    from oracle import optimizer_port as OP

    g = torch.Generator().manual_seed(11)
    p0 = torch.randn(shape, generator=g)
    p_dev = torch.nn.Parameter(p0.to(DEV).contiguous(memory_format=torch.channels_last_3d) if cl else p0.to(DEV))
    opt = Adam([{"params": [p_dev], "lr": 0.02, "name": "x"}], betas=(0.9, 0.99), weight_decay=wd)
    count = torch.randint(0, 7, shape, generator=g)
    if per_lr:
        opt.set_pervoxel_lr(count.to(DEV))
    p, m, v = p0.clone(), torch.zeros(shape), torch.zeros(shape)
    for step in range(1, 8):
        grad = torch.randn(shape, generator=g) * (0.3 if step % 3 else 3.0)
        gd = grad.to(DEV)
        p_dev.grad = gd.contiguous(memory_format=torch.channels_last_3d) if cl else gd
        opt.step()
        OP.adam_update(p, grad, m, v, step, 0.02, 0.9, 0.99, 1e-8, wd, (count.float() / count.max()) if per_lr else None)
        assert C.rel_err(p_dev.detach().contiguous(), p) < 5e-6, step
    st = opt.state[p_dev]
    assert st["step"] == 7 and C.rel_err(st["exp_avg"].contiguous(), m) < 5e-6 and C.rel_err(st["exp_avg_sq"].contiguous(), v) < 5e-6
    assert st["exp_avg"].stride() == p_dev.stride()


def test_optimizer_drives_a_fine_stage_step():
    """create_optimizer_or_freeze_model + Adam on the render model: one render step + optimizer step moves every
    trainable tensor (incl. the channels-last colour grids) and leaves frozen ones untouched"""
    from esr_nerf_b200 import synthetic as S
    from esr_nerf_b200.optimizer import create_optimizer_or_freeze_model

    fx, weights = C.load_case("fine_sparse_s20")
    m = C.build_product_model(fx, weights, DEV)
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    opt = create_optimizer_or_freeze_model(m, off_color=0.1, off_rgbnet=0.003, emo_color=0.1, emo_rgbnet=0.0, sdf=0.0005,
                                           tonemapper=0.003)
    rays = {k: v.to(DEV) for k, v in S.make_rays(256, 3).items()}
    out = m(s_val=20.0, **rays)
    ((out["srgb/rgb"] - rays["rgbs"]) ** 2).mean().backward()
    opt.step()
    for k, p in m.named_parameters():
        moved = not torch.equal(p.detach(), before[k])
        if k.startswith("emo_rgbnet") or k.startswith("tv_smooth_conv"):
            assert not moved and not p.requires_grad, k
        else:
            assert moved, k
    assert m.off_color.grid.is_contiguous(memory_format=torch.channels_last_3d)

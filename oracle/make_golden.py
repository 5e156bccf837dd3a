"""ORACLE — generates the committed golden vectors in tests/golden/ by running the reference's OWN
render model (app/fine/model/voxurff.py, imported through oracle/ref_harness.py) on the CPU.

    python -m oracle.make_golden            # needs /root/reference; run in the build container only

Each fixture holds the synthetic inputs' seeds, the reference outputs, the intermediate sample streams
(post-MaskCache and shaded), and digests of every parameter gradient under fixed random cotangents.
Grids are regenerated from seeds by esr_nerf_b200.synthetic (deterministic CPU torch), MLP weights are
stored once in tests/golden/fine_weights.npz.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from esr_nerf_b200 import synthetic as S  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (num_voxels, mask_res, sparse, n_rays, s_val, ray_seed)
    "fine_sparse_s20": (40 ** 3, 20, True, 128, 20.0, 1234),
    "fine_dense_s220": (40 ** 3, 20, False, 128, 220.0, 4321),
    "fine_sparse_s60_big": (64 ** 3, 32, True, 192, 60.0, 99),
}
# `neus_alpha: grad` (functions.py:45-69): the reference's VoxurfF built with that option, on the rays the GPU test of the
# mode uses (tests/test_gpu_neus_grad.py); written by `python -m oracle.make_golden --neus-grad-only` (fine_weights.npz as is)
GRAD_ALPHA_CASES = {"fine_grad_sparse_s60_big": (64 ** 3, 32, True, 1536, 60.0, 4711)}
GRAD_PROBES = 4096


def build_reference_model(num_voxels, mask_res, sparse, s_val, weights=None, **model_overrides):
    _, _, VoxurfF, _ = H.reference_classes()
    cfg = H.fine_cfg(**model_overrides)
    torch.manual_seed(0)
    m = VoxurfF(cfg, S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT,
                S.mask_density(mask_res, sparse), s_val, num_voxels)
    if weights is not None:
        m.load_state_dict({**m.state_dict(), **weights})
    S.fill_fine_model(m)
    m.train()
    return m


def mlp_weight_keys(sd):
    return [k for k in sd if ("rgbnet" in k or "tonemapper" in k)]


def cotangents(n, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"srgb/rgb": torch.randn(n, 3, generator=g), "lin/rgb": torch.randn(n, 3, generator=g),
            "etc/alphainv_cum": torch.randn(n, generator=g), "etc/white_bg": torch.randn(n, 1, generator=g)}


def grad_digest(name, g: torch.Tensor):
    """sum, abs-sum and GRAD_PROBES seeded probe values of a gradient tensor (logical [1,C,X,Y,Z] / [O,I] order)."""
    flat = g.detach().contiguous().reshape(-1).double()
    # (hash() is salted per process -> derive the seed from the name bytes instead)
    gen = torch.Generator().manual_seed(int.from_bytes(name.encode(), "little") % (2 ** 31))
    nz = torch.nonzero(flat).reshape(-1)
    k = min(GRAD_PROBES // 2, nz.numel())
    pick_nz = nz[torch.randperm(nz.numel(), generator=gen)[:k]] if k else nz[:0]
    pick_any = torch.randint(0, flat.numel(), (GRAD_PROBES - k,), generator=gen)
    idx = torch.cat([pick_nz, pick_any])
    return {f"grad/{name}/sum": flat.sum().item(), f"grad/{name}/abs_sum": flat.abs().sum().item(),
            f"grad/{name}/idx": idx.numpy().astype(np.int64), f"grad/{name}/val": flat[idx].float().numpy()}


def run_case(name, spec, weights, **model_overrides):
    num_voxels, mask_res, sparse, n, s_val, seed = spec
    m = build_reference_model(num_voxels, mask_res, sparse, s_val, weights, **model_overrides)
    rays = S.make_rays(n, seed)
    # instrument the stream: re-run the same calls the reference forward makes (voxurff.py:186-213)
    with torch.no_grad():
        ray_pts, ray_id, step_id = m.sample_ray(rays_o=rays["rays_o"], rays_d=rays["rays_d"])
        m0 = ray_pts.shape[0]
        keep = m.mask_cache(ray_pts)
        ray_pts, ray_id, step_id = ray_pts[keep], ray_id[keep], step_id[keep]
        sdf, grad = m.sample_sdf_grad(ray_pts)
        alpha = m.neus_alpha_from_sdf_scatter(rays["viewdirs"], ray_id, m.stepsize * m.voxel_size, sdf, grad, s_val)
        k0 = alpha > m.fastcolor_thres
        w, T, last, _, _ = H.alpha2weight(alpha[k0], ray_id[k0], n)
        k1 = w > m.fastcolor_thres
        m3_ray, m3_step, m3_w = ray_id[k0][k1], step_id[k0][k1], w[k1]
    out = m(s_val=s_val, rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"],
            em_modes=rays["em_modes"], rgbs=rays["rgbs"])
    cot = cotangents(n)
    loss = sum((out[k] * cot[k]).sum() for k in cot)
    loss.backward()
    fx = dict(num_voxels=num_voxels, mask_res=mask_res, sparse=int(sparse), n_rays=n, s_val=s_val, ray_seed=seed,
              m0=m0, m1_ray=ray_id.numpy().astype(np.int32), m1_step=step_id.numpy().astype(np.int32),
              m1_sdf=sdf.numpy(), m1_alpha=alpha.numpy(), m3_ray=m3_ray.numpy().astype(np.int32),
              m3_step=m3_step.numpy().astype(np.int32), m3_weights=m3_w.numpy(), loss=loss.item())
    for k, v in out.items():
        fx["out/" + k] = v.detach().numpy()
    for pname, p in m.named_parameters():
        if p.grad is not None:
            fx.update(grad_digest(pname, p.grad))
    np.savez_compressed(os.path.join(GOLDEN, f"voxurff_{name}.npz"), **fx)
    print(f"{name}: m0={m0} m1={ray_id.numel()} m3={m3_ray.numel()} loss={loss.item():.6f}")


COARSE_CASES = {
    # name: (num_voxels, mask_res, sparse, n_rays, s_val, ray_seed)   — BASELINE config 1 shape, shrunk
    "coarse_sparse_s5": (32 ** 3, 16, True, 128, 5.0, 1234),
    "coarse_dense_s25": (40 ** 3, 20, False, 160, 25.0, 555),
}


def coarse_cfg(num_voxels, **over):
    return H.DictConfig(dict(system=dict(device="cpu"),
                             app=dict(model=dict(S.COARSE_MODEL_CFG, num_voxels=num_voxels, **over))))


def build_reference_coarse(num_voxels, mask_res, sparse, s_val, weights=None, **cfg_over):
    _, VoxurfC, _, _ = H.reference_classes()
    torch.manual_seed(0)
    m = VoxurfC(coarse_cfg(num_voxels, **cfg_over), S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT,
                S.mask_density(mask_res, sparse), s_val)
    if weights is not None:
        m.load_state_dict({**m.state_dict(), **weights})
    S.fill_coarse_model(m)
    m.train()
    return m


def coarse_cotangents(n, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"srgb/rgb": torch.randn(n, 3, generator=g), "etc/alphainv_cum": torch.randn(n, generator=g),
            "etc/white_bg": torch.randn(n, 1, generator=g)}


def run_coarse_case(name, spec, weights):
    num_voxels, mask_res, sparse, n, s_val, seed = spec
    m = build_reference_coarse(num_voxels, mask_res, sparse, s_val, weights)
    rays = S.make_rays(n, seed)
    out = m(s_val=s_val, rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"],
            em_modes=rays["em_modes"], rgbs=rays["rgbs"])
    cot = coarse_cotangents(n)
    loss = sum((out[k] * cot[k]).sum() for k in cot)
    loss.backward()
    fx = dict(num_voxels=num_voxels, mask_res=mask_res, sparse=int(sparse), n_rays=n, s_val=s_val, ray_seed=seed,
              loss=loss.item())
    for k, v in out.items():
        fx["out/" + k] = v.detach().numpy()
    for pname, p in m.named_parameters():
        if p.grad is not None:
            fx.update(grad_digest(pname, p.grad))
    np.savez_compressed(os.path.join(GOLDEN, f"voxurfc_{name}.npz"), **fx)
    print(f"{name}: loss={loss.item():.6f}")


def main_coarse():
    m = build_reference_coarse(32 ** 3, 16, True, 5.0)
    sd = m.state_dict()
    weights = {k: sd[k].clone() for k in sd if "rgbnet" in k}
    g = torch.Generator().manual_seed(12)
    for k in weights:  # the reference zero-initialises the last bias; give every bias a non-trivial value
        if k.endswith("bias"):
            weights[k] = weights[k] + 0.05 * torch.randn(weights[k].shape, generator=g)
    np.savez_compressed(os.path.join(GOLDEN, "coarse_weights.npz"), **{k: v.numpy() for k, v in weights.items()})
    for name, spec in COARSE_CASES.items():
        run_coarse_case(name, spec, weights)


DVGO_CASES = {"dvgo_24": (24 ** 3, 96, 77), "dvgo_40": (40 ** 3, 64, 78)}   # name: (num_voxels, n_rays, seed)


def dvgo_cotangents(n, S, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {"etc/alphainv_cum": torch.randn(n, S + 1, generator=g), "etc/weights": torch.randn(n, S, generator=g),
            "etc/white_bg": torch.randn(n, 1, generator=g), "srgb/raw_rgb": 0.1 * torch.randn(n, S, 3, generator=g),
            "srgb/rgb": torch.randn(n, 3, generator=g)}


def build_reference_dvgo(num_voxels):
    DVGO, _, _, _ = H.reference_classes()
    cfg = H.DictConfig(dict(system=dict(device="cpu"), app=dict(model=dict(num_voxels=num_voxels, stepsize=0.5,
                                                                            alpha_init=S.MASK_ALPHA_INIT))))
    m = DVGO(cfg, S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX)
    S.fill_dvgo_model(m)
    m.train()
    return m


def main_dvgo():
    for name, (num_voxels, n, seed) in DVGO_CASES.items():
        m = build_reference_dvgo(num_voxels)
        rays = S.make_rays(n, seed)
        torch.manual_seed(seed)                      # the reference draws the sampler jitter with torch.rand_like
        out = m(rays_o=rays["rays_o"], rays_d=rays["rays_d"], em_modes=rays["em_modes"])
        torch.manual_seed(seed)
        jitter = torch.rand(n, 1)                    # same stream, stored with the fixture
        cot = dvgo_cotangents(n, m.N_samples)
        loss = sum((out[k] * cot[k]).sum() for k in cot)
        loss.backward()
        fx = dict(num_voxels=num_voxels, n_rays=n, ray_seed=seed, n_samples=m.N_samples, jitter=jitter.numpy(),
                  loss=loss.item())
        for k, v in out.items():
            fx["out/" + k] = v.detach().numpy()
        for pname, p in m.named_parameters():
            fx.update(grad_digest(pname, p.grad))
        m.eval()
        with torch.no_grad():
            for em in (0, 1):
                ev = m(rays_o=rays["rays_o"], rays_d=rays["rays_d"], em_modes=torch.tensor(em))
                for k, v in ev.items():
                    fx[f"eval{em}/" + k] = v.numpy()
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **fx)
        print(f"{name}: S={m.N_samples} loss={loss.item():.6f}")


# ---------------------------------------------------------------------------------------------------
# LTS / PDRA stage (ESRNeRF.forward_training, esrnerf.py:681-851)
# ---------------------------------------------------------------------------------------------------
ESRNERF_CASES = {
    # name: (num_voxels, mask_res, sparse, n_rays, s_val, ray_seed, num_2ndrays, num_ltspts, pdra_mode, draw_seed)
    "lts_sparse_s220": (40 ** 3, 20, True, 128, 220.0, 1234, 16, 12, False, 41),
    "pdra_sparse_s60": (48 ** 3, 24, True, 160, 60.0, 77, 8, 20, True, 42),
}
NORMAL_EPS, EMIT_EPS = 0.01, 0.02
EVAL_CHUNK = 300          # LTS points per chunk in forward_evaluate(render_pbr=True): several chunks per case


def lts_cfg(**over):
    return H.DictConfig(dict(system=dict(device="cpu"), app=dict(model=dict(S.LTS_MODEL_CFG, **over))))


def build_reference_esrnerf(num_voxels, mask_res, sparse, s_val, weights=None, **cfg_over):
    _, _, _, ESRNeRF = H.reference_classes()
    torch.manual_seed(0)
    m = ESRNeRF(lts_cfg(**cfg_over), S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT,
                S.mask_density(mask_res, sparse), s_val, num_voxels)
    if weights is not None:
        m.load_state_dict({**m.state_dict(), **weights})
    S.fill_esrnerf_model(m)
    m.train()
    return m


class patched_draws:
    """Serve the reference's four random draws (np.random.choice, torch.randn, torch.randn_like x2) from a
    device-independent seeded stream (oracle.esrnerf_port.FixedDraws) while the reference forward runs."""

    def __init__(self, draws):
        self.d = draws

    def __enter__(self):
        self.saved = (np.random.choice, torch.randn, torch.randn_like)
        d = self.d
        np.random.choice = lambda n, k, replace=False: d.choice(int(n), int(k)).numpy()
        torch.randn = lambda *shape, device=None, **kw: d.randn(*shape)
        torch.randn_like = lambda t, **kw: d.randn(*t.shape)
        return self

    def __exit__(self, *exc):
        np.random.choice, torch.randn, torch.randn_like = self.saved


def esrnerf_cotangents(out, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(out[k].shape, generator=g) for k in sorted(out)}


def run_esrnerf_case(name, spec, weights):
    from oracle import esrnerf_port as E

    num_voxels, mask_res, sparse, n, s_val, seed, n2, npts, pdra, dseed = spec
    m = build_reference_esrnerf(num_voxels, mask_res, sparse, s_val, weights, num_2ndrays=n2, num_ltspts=npts)
    m.pdra_mode = pdra
    rays = S.make_rays(n, seed)
    with patched_draws(E.FixedDraws(dseed)):
        out = m(s_val=s_val, rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"],
                em_modes=rays["em_modes"], uncert_masks=S.uncert_masks(n), normal_eps=NORMAL_EPS, emit_eps=EMIT_EPS)
    cot = esrnerf_cotangents(out)
    loss = sum((out[k] * cot[k]).sum() for k in cot)
    loss.backward()
    fx = dict(num_voxels=num_voxels, mask_res=mask_res, sparse=int(sparse), n_rays=n, s_val=s_val, ray_seed=seed,
              num_2ndrays=n2, num_ltspts=npts, pdra_mode=int(pdra), draw_seed=dseed, normal_eps=NORMAL_EPS,
              emit_eps=EMIT_EPS, loss=loss.item())
    for k, v in out.items():
        fx["out/" + k] = v.detach().numpy()
    for pname, p in m.named_parameters():
        if p.grad is not None:
            fx.update(grad_digest(pname, p.grad))
    # inference entry points (esrnerf.py:853-1407): forward_evaluate with the PBR decomposition, eval_emit, eval_esp
    m.eval()
    pos_rt = torch.linalg.qr(E._randn(3, 3, generator=torch.Generator().manual_seed(3)))[0].contiguous()
    fx["pos_rt"] = pos_rt.numpy()
    fx["eval_chunk"] = EVAL_CHUNK
    with torch.no_grad():
        for em in (0, 1):
            with patched_draws(E.FixedDraws(dseed + 100)):
                ev = m(rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"], em_modes=torch.tensor(em),
                       pos_rt=pos_rt, render_pbr=True, chunk_sz=EVAL_CHUNK)
            for k, v in ev.items():
                fx[f"eval{em}/" + k] = v.numpy()
        kw = dict(rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"])
        fx["eval_emit"] = m.eval_emit(**kw).numpy()
        fx["eval_esp"] = m.eval_esp(**kw).numpy()
    # scene-editing finetune target (esrnerf.py:241-484)
    m.train(finetune=True)
    m.zero_grad(set_to_none=True)
    S.perturb_emit_color(m)
    ft_in = S.finetune_inputs(n)
    with patched_draws(E.FixedDraws(dseed + 200)):
        ft = m(rays_o=rays["rays_o"], rays_d=rays["rays_d"], viewdirs=rays["viewdirs"], **ft_in)
    ft_cot = E._randn(ft["lin/pbr/emo"].shape, generator=torch.Generator().manual_seed(8))
    (ft["lin/pbr/emo"] * ft_cot).sum().backward()
    for k, v in ft.items():
        fx["ft/" + k] = v.detach().numpy()
    for pname, p in m.named_parameters():
        if p.grad is not None:
            fx.update({"ft" + k: v for k, v in grad_digest(pname, p.grad).items()})
    np.savez_compressed(os.path.join(GOLDEN, f"esrnerf_{name}.npz"), **fx)
    print(f"{name}: m3={out['etc/emit'].shape[0]} loss={loss.item():.6f}")


def main_esrnerf():
    fine = dict(np.load(os.path.join(GOLDEN, "fine_weights.npz")))
    m = build_reference_esrnerf(40 ** 3, 20, True, 220.0)
    sd = m.state_dict()
    weights = {k: sd[k].clone() for k in sd if ("emitnet" in k or "brdfnet" in k or "envmap" in k)}
    g = torch.Generator().manual_seed(13)
    for k in weights:  # the reference zero-initialises the last biases; give every bias a non-trivial value
        if k.endswith("bias"):
            weights[k] = weights[k] + 0.05 * torch.randn(weights[k].shape, generator=g)
    np.savez_compressed(os.path.join(GOLDEN, "lts_weights.npz"), **{k: v.numpy() for k, v in weights.items()})
    weights.update({k: torch.from_numpy(v) for k, v in fine.items()})
    for name, spec in ESRNERF_CASES.items():
        run_esrnerf_case(name, spec, weights)


def main():
    if "--esrnerf-only" in sys.argv:
        return main_esrnerf()
    if not H.reference_available():
        raise SystemExit("reference tree not available; golden vectors can only be generated in the build container")
    os.makedirs(GOLDEN, exist_ok=True)
    wpath = os.path.join(GOLDEN, "fine_weights.npz")
    if "--neus-grad-only" in sys.argv:      # adds the grad-alpha fixtures next to the existing ones (same net weights)
        weights = {k: torch.from_numpy(v) for k, v in np.load(wpath).items()}
        for name, spec in GRAD_ALPHA_CASES.items():
            run_case(name, spec, weights, neus_alpha="grad")
        return
    m = build_reference_model(40 ** 3, 20, True, 20.0)
    sd = m.state_dict()
    weights = {k: sd[k].clone() for k in mlp_weight_keys(sd)}
    # give the biases / last layers non-trivial values so every gradient path is exercised
    g = torch.Generator().manual_seed(11)
    for k in weights:
        if k.endswith("bias"):
            weights[k] = weights[k] + 0.05 * torch.randn(weights[k].shape, generator=g)
    np.savez_compressed(wpath, **{k: v.numpy() for k, v in weights.items()})
    if "--dvgo-only" in sys.argv:
        return main_dvgo()
    if "--coarse-only" not in sys.argv:
        for name, spec in GRAD_ALPHA_CASES.items():
            run_case(name, spec, weights, neus_alpha="grad")
    if "--coarse-only" not in sys.argv:
        for name, spec in CASES.items():
            run_case(name, spec, weights)
    main_coarse()
    main_dvgo()
    main_esrnerf()


if __name__ == "__main__":
    main()

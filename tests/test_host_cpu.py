"""CPU tests of the host-side logic around the C ABI (no compute calls): cached scene geometry, flat parameter
assembly of the tensor-core MLPs (layout + gradient routing)."""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from esr_nerf_b200 import modules as M


def test_host_geometry_cache_follows_tensor_identity_and_version():
    o = types.SimpleNamespace(xyz_min=torch.tensor([-1.0, -1.0, -1.0]), xyz_max=torch.tensor([1.0, 1.0, 1.0]),
                              mask_xyz_min=torch.tensor([-1.0, -1.0, -1.0]), mask_xyz_max=torch.tensor([1.0, 1.0, 1.0]),
                              voxel_size=torch.tensor(0.02))
    a = M.host_geometry(o, 0.5)
    assert a["xyz_min"] == [-1.0, -1.0, -1.0] and a["stepdist"] == float(0.5 * o.voxel_size)
    assert M.host_geometry(o, 0.5) is a                      # cached: no read-back per call
    o.voxel_size = torch.tensor(0.01)                        # set_grid_resolution: a new tensor
    b = M.host_geometry(o, 0.5)
    assert b is not a and b["voxel_size"] == float(torch.tensor(0.01))
    o.xyz_max.mul_(2.0)                                      # in-place edit (load_state_dict): version bump
    c = M.host_geometry(o, 0.5)
    assert c["xyz_max"] == [2.0, 2.0, 2.0]
    assert M.host_geometry(o, 0.25)["stepdist"] == float(0.25 * o.voxel_size)


def _reference_flat(layers, cols, k0):
    """the flat image spelt out with differentiable torch ops: per layer W then b, layer-0 columns permuted (zero
    column for unused inputs), output layer padded to 8 rows"""
    n_ref = layers[0].in_features
    idx = torch.where(cols < 0, torch.full_like(cols, n_ref), cols)
    parts = []
    for i, lin in enumerate(layers):
        w, b = lin.weight, lin.bias
        if i == 0:
            w = torch.cat([w, w.new_zeros(w.shape[0], 1)], 1).index_select(1, idx)
        if i + 1 == len(layers):
            w, b = F.pad(w, (0, 0, 0, 8 - w.shape[0])), F.pad(b, (0, 8 - b.shape[0]))
        parts += [w.reshape(-1), b]
    return torch.cat(parts)


def test_flat_params_layout_and_gradient_routing():
    torch.manual_seed(0)
    for kind, k0, dims in (("off", 96, (85, 192, 192, 192, 3)), ("emo", 96, (85, 192, 192, 192, 3)), ("tone", 48, (33, 192, 3))):
        layers = [nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])]
        cols = M.tonemap_in_cols("cpu") if kind == "tone" else M.radiance_in_cols(kind, "cpu")
        ref = _reference_flat(layers, cols, k0)
        g = torch.randn_like(ref)
        (ref * g).sum().backward()
        want = [p.grad.clone() for l in layers for p in (l.weight, l.bias)]
        for l in layers:
            l.zero_grad()
        got = M.flat_mlp_params(layers, kind, k0)
        assert torch.equal(got, ref.detach())
        (got * g).sum().backward()
        for p, w in zip([p for l in layers for p in (l.weight, l.bias)], want):
            assert torch.equal(p.grad, w)


def test_padded_flat_params_match_zero_padding():
    torch.manual_seed(1)
    for kind, n_out in (("emit", 3), ("brdf", 5)):
        layers = [nn.Linear(76, 128), nn.Linear(128, 128), nn.Linear(128, 128), nn.Linear(128, n_out)]
        cols = M.pbr_in_cols(kind, "cpu")
        idx = torch.where(cols < 0, torch.full_like(cols, 76), cols)
        parts = []
        for i, lin in enumerate(layers):
            w, b = lin.weight, lin.bias
            w = torch.cat([w, w.new_zeros(w.shape[0], 1)], 1).index_select(1, idx) if i == 0 else F.pad(w, (0, 192 - w.shape[1]))
            rows = 8 if i == 3 else 192
            parts += [F.pad(w, (0, 0, 0, rows - w.shape[0])).reshape(-1), F.pad(b, (0, rows - b.shape[0]))]
        ref = torch.cat(parts)
        g = torch.randn_like(ref)
        (ref * g).sum().backward()
        want = [p.grad.clone() for l in layers for p in (l.weight, l.bias)]
        for l in layers:
            l.zero_grad()
        got = M.flat_mlp_params_padded(layers, kind)
        assert torch.equal(got, ref.detach())
        (got * g).sum().backward()
        for p, w in zip([p for l in layers for p in (l.weight, l.bias)], want):
            assert torch.equal(p.grad, w)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm of the measurement contract) on a tiny bounded sample: exactly one JSON
    line on stdout with the contract's keys; everything else goes to stderr"""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-rays", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train rays/sec (fwd+bwd)" and d["unit"] == "rays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_fibonacci_scattering_properties_and_reference():
    """`ray_sampling: fib` (pbr/functions.py:21-32, 176-194): unit directions, deterministic, mirrored into each point's
    hemisphere; against a float64 restatement of the published spiral and — where /root/reference exists — the
    reference's own function, bit for bit"""
    import math

    import numpy as np

    from esr_nerf_b200 import pbr

    n = 37
    d = pbr.fibonacci_hemisphere(n)
    assert d.shape == (n, 3) and d.dtype == torch.float32
    rn = np.arange(n, 2 * n, dtype=np.float64)
    phi = math.pi * (3.0 - math.sqrt(5.0)) * ((rn + 1.0) % (2 * n))
    ct = (rn + 0.5) / n - 1.0
    st = np.sqrt(1.0 - ct * ct)
    want = np.stack([np.cos(phi) * st, np.sin(phi) * st, ct], -1)
    assert np.abs(d.numpy() - want).max() < 2e-5            # fp32 phase of up to ~180 rad
    assert (d[:, 2] > 0).all() and torch.allclose(d.norm(dim=-1), torch.ones(n), atol=1e-6)

    torch.manual_seed(3)
    normal = F.normalize(torch.randn(11, 3), dim=-1)
    dirs = pbr.diffuse_scattering_fib(normal, n)
    assert dirs.shape == (11, n, 3)
    assert ((dirs * normal[:, None]).sum(-1) >= 0).all()
    assert torch.equal(dirs.abs(), d.abs().expand(11, n, 3))        # the same spiral for every point, up to the mirror
    assert torch.equal(dirs, pbr.diffuse_scattering_fib(normal, n))

    from oracle import ref_harness as H
    if H.reference_available():
        H.install_stubs()
        import importlib
        ref = importlib.import_module("app.utils.pbr.functions")
        assert torch.equal(d, ref.fibonacci_spiral_samples_on_unit_hemisphere(n))
        assert torch.equal(dirs, ref.diffuse_scattering_fib(normal, n))


def test_spherical_gaussian_activations_match_reference():
    """`env_activation` (pbr/module.py:86-143): every activation the reference documents, same seeded initialisation
    (incl. which ones re-parametrise `mus`), same forward"""
    import pytest

    with pytest.raises(AttributeError):
        M.SphericalGaussian(8, "no_such_activation")
    dirs = F.normalize(torch.randn(5, 7, 3), dim=-1)
    torch.manual_seed(7)
    sp = M.SphericalGaussian(48, "softplus")
    assert sp.activation is F.softplus and sp(dirs).shape == (5, 7, 3) and (sp(dirs) > 0).all()

    from oracle import ref_harness as H
    if not H.reference_available():
        pytest.skip("/root/reference not present")
    H.install_stubs()
    import importlib
    ref_mod = importlib.import_module("app.utils.pbr.module")
    for act in ("relu", "abs", "exp", "sigmoid", "softplus"):
        torch.manual_seed(11)
        mine = M.SphericalGaussian(16, act)
        torch.manual_seed(11)
        ref = ref_mod.SphericalGaussian(16, act)
        for k, v in ref.state_dict().items():
            assert torch.equal(mine.state_dict()[k], v), (act, k)
        assert torch.equal(mine(dirs), ref(dirs)), act


def test_esrnerf_scatter_dispatch():
    """ESRNeRF._scatter: `ray_sampling` random draws through the model's (replaceable) normal source, fib draws nothing
    (esrnerf.py:188-192, 302, 534, 874); unknown names are rejected at construction"""
    import pytest

    from esr_nerf_b200 import pbr
    from esr_nerf_b200.esrnerf import ESRNeRF

    normal = F.normalize(torch.randn(6, 3), dim=-1)
    noise = torch.randn(6, 9, 3)
    calls = []

    def randn(*shape, dev):
        calls.append(shape)
        return noise

    o = types.SimpleNamespace(fib_sampling=False, _randn=randn)
    got = ESRNeRF._scatter(o, normal, 9)
    assert calls == [(6, 9, 3)] and torch.equal(got, pbr.diffuse_scattering(normal, noise))
    o.fib_sampling = True
    got = ESRNeRF._scatter(o, normal, 9)
    assert len(calls) == 1 and torch.equal(got, pbr.diffuse_scattering_fib(normal, 9))


def test_bench_step_roofline_aggregates_measured_rows():
    """bench.step_roofline on the per-kernel rows of a bench line measured on the B200 (profiles/r01_bench_lines): the
    aggregate of the HBM-bound kernels is their summed algorithmic bytes over their summed time"""
    import json
    import os

    import bench

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.loads(open(os.path.join(root, "profiles", "r01_bench_lines", "fine_n1.json")).read().strip().splitlines()[-1])
    rows = []
    for r in line["kernels"]:
        r = dict(r)
        if "bound" in r:   # lines older than the work_per_step key: rebuild it from rate x time
            r["work_per_step"] = r["achieved"] * (1e9 if r["bound"] == "hbm" else 1e12) * r["ms_per_step"] * 1e-3
        rows.append(r)
    pk = {"hbm_gbs": 6448.4, "bf16_tflops_sustained": 1386.5}
    agg = bench.step_roofline(rows, pk)
    hbm = [r for r in rows if r.get("bound") == "hbm"]
    assert agg["hbm"]["kernels"] == len(hbm) and agg["tensor"]["kernels"] == sum(r.get("bound") == "tensor" for r in rows)
    want = sum(r["work_per_step"] for r in hbm) / sum(r["ms_per_step"] for r in hbm) / 1e-3 / 1e9
    assert abs(agg["hbm"]["achieved"] - want) < 1e-6 * want
    assert 0.3 < agg["hbm"]["frac"] < 1.0 and 0.1 < agg["tensor"]["frac"] < 1.0
    assert agg["other_ms_per_step"] == sum(r["ms_per_step"] for r in rows if "bound" not in r)
    assert bench.step_roofline([], pk)["hbm"]["frac"] == 0.0       # no rows: zeros, no division


def test_cabi_gradient_exchange_argument_checks_without_a_device():
    """the exchange entry points validate their arguments before any CUDA call (include/esr_b200.h §2e): status codes
    and esr_last_error() can be checked on a box without a GPU"""
    import ctypes

    from esr_nerf_b200 import _lib

    L = _lib.lib()
    ch = (ctypes.c_int32 * 3)(1, 6, 6)
    assert L.esr_grad_pack_floats(ch, 3, 5) == 5 + 1 + 30 + 30            # every block starts on an even float offset
    assert L.esr_grad_pack_floats(ch, 3, 4) == 4 + 24 + 24
    assert L.esr_grad_pack_floats(ch, 0, 4) == -1 and L.esr_grad_pack_floats(ch, 5, 4) == -1
    assert L.esr_grad_pack_floats(None, 3, 4) == -1 and L.esr_grad_pack_floats(ch, 3, -1) == -1
    bad = (ctypes.c_int32 * 1)(0)
    assert L.esr_grad_pack_floats(bad, 1, 4) == -1
    vols = (ctypes.c_void_p * 3)(None, None, None)
    assert L.esr_grad_pack(vols, ch, 3, None, 0, None, None) == 0          # nothing to move: no pointer is touched
    assert L.esr_grad_pack(vols, ch, 0, None, 4, None, None) != 0          # no volumes
    assert L.esr_grad_unpack(vols, ch, 9, None, 4, None, None) != 0        # more than ESR_MAX_GRAD_VOLUMES
    assert L.esr_grad_pack(vols, ch, 3, None, 4, None, None) != 0          # null index / buffer
    assert L.esr_last_error()
    # block map: the block edges must divide the grid extents, pointers must be there
    assert L.esr_grad_block_flags(vols, ch, 3, 16, 16, 12, 8, 8, 8, None, None) != 0
    assert L.esr_grad_block_flags(None, ch, 3, 16, 16, 16, 8, 8, 8, None, None) != 0
    assert L.esr_grad_block_flags(vols, ch, 0, 16, 16, 16, 8, 8, 8, None, None) != 0

"""GPU parity (-m gpu) of the rows either side of the render step, device-resident:

  * samplers.BatchSampler / RayGroupManager on cuda:0 against oracle/samplers_port.py — the reference's physically
    re-ordered ray set (utils2/utils.py:41-312), pinned against the reference's own classes on the CPU by
    tests/test_samplers_cpu.py::test_port_matches_reference — under the same CUDA generator seed: same draws, same batches,
    same checkpointed state; and the rank slices of a 2 / 3-way sharded feed concatenate to the single-process batch;
  * samplers.update_ray_groups (pdra.py:882-932) with the REAL renderer (ESRNeRF.eval_emit, the hand-written kernels)
    against the reference's sweep restated on the port sampler with another chunk size;
  * samplers.filter_edit_rays (pdra.py:934-1045) with the real ESRNeRF.eval_esp: cuda result against the same function on
    the CPU copy of the sampler with a renderer that replays the GPU's surface points (the CPU function is pinned to the
    reference's method by tests/test_samplers_cpu.py::test_filter_edit_rays_matches_reference_method);
  * VoxurfF.scale_volume_grid (voxurff.py:547-566) on cuda:0 against the same model on the CPU (bit-equal to the
    reference there: tests/test_oracle_cpu.py::test_progressive_grid_rescale_matches_reference), and a render step on the
    rescaled model."""
import types

import pytest
import torch

import esr_testlib as C
from esr_nerf_b200 import samplers as SM
from esr_nerf_b200 import synthetic as S
from oracle import samplers_port as SP

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ["rays_o", "rgbs", "em_modes"]


def _cfg(device=DEV, preload="cuda"):
    return types.SimpleNamespace(system=types.SimpleNamespace(device=device, data_preload=preload))


def _data(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"rays_o": torch.randn(n, 3, generator=g), "rgbs": torch.rand(n, 3, generator=g),
            "em_modes": torch.randint(0, 2, (n,), generator=g), "unused": torch.zeros(n)}


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].device.type == "cuda" and torch.equal(a[k], b[k]), k


def test_batch_sampler_on_device_matches_port_sequence():
    n, bs = 100_000, 4096
    g = torch.Generator().manual_seed(3)
    mask = torch.rand(n, generator=g) < 0.7

    def drive(make):
        torch.manual_seed(11)            # seeds the CUDA generator torch.randperm(device=cuda) draws from as well
        s = make()
        s.filter(mask)
        s.shuffle()
        return s, [s.sample() for _ in range(40)]      # wraps (re-shuffles) twice

    port, want = drive(lambda: SP.BatchSamplerPort(DEV, _data(n), KEYS, bs))
    mine, got = drive(lambda: SM.BatchSampler(_cfg(), _data(n), KEYS, bs))
    for a, b in zip(got, want):
        _same(a, b)
    assert mine.batch_st == port.batch_st and torch.equal(mine.data_idxs, port.data_idxs) and mine.data_num == port.data_num
    for k in KEYS:                       # the physical copy the reference holds == the gathered view
        assert torch.equal(mine.data[k], port.data[k])
    # resume from the checkpointed state (fine.py:221-228), and the ranks' slices of a sharded feed
    torch.manual_seed(5)
    port2 = SP.BatchSamplerPort(DEV, _data(n), KEYS, bs, port.batch_st, port.data_idxs.clone())
    tail = [port2.sample() for _ in range(30)]
    for world in (1, 2, 3):
        parts = []
        for rank in range(world):
            torch.manual_seed(5)
            s = SM.BatchSampler(_cfg(), _data(n), KEYS, bs, mine.batch_st, mine.data_idxs.clone(), rank=rank, world=world)
            parts.append([s.sample() for _ in range(30)])
        for i, b in enumerate(tail):
            _same({k: torch.cat([p[i][k] for p in parts]) for k in KEYS}, b)


def test_ray_group_manager_on_device_matches_port_sequence():
    n = 60_000

    def drive(make):
        torch.manual_seed(21)
        m = make()
        m.shuffle()
        out = [m.sample() for _ in range(3)]             # no certain rays yet: every uncert_mask is False (:302 quirk)
        g = torch.Generator().manual_seed(8)
        m.filter(torch.rand(m.uncert_data_num, generator=g) < 0.6)
        out += [m.sample() for _ in range(25)]           # both groups wrap
        m.filter(torch.rand(m.uncert_data_num, generator=g) < 0.02)     # uncertain group smaller than its batch size
        out += [m.sample() for _ in range(4)]
        return m, out

    port, want = drive(lambda: SP.RayGroupManagerPort(DEV, _data(n), KEYS, 4096, 2048))
    mine, got = drive(lambda: SM.RayGroupManager(_cfg(), _data(n), KEYS, 4096, 2048))
    assert not want[0]["uncert_masks"].any()
    for a, b in zip(got, want):
        _same(a, b)
    for name in ("uncert_batch_st", "cert_batch_st", "uncert_data_num", "cert_data_num"):
        assert getattr(mine, name) == getattr(port, name), name
    assert torch.equal(mine.uncert_data_idxs, port.uncert_data_idxs) and torch.equal(mine.cert_data_idxs, port.cert_data_idxs)
    assert torch.equal(mine.uncert_data["rays_o"], port.uncert_data["rays_o"])       # pdra.py:888
    assert torch.equal(mine.cert_data["rgbs"], port.cert_data["rgbs"])


def _esrnerf_and_rays(n):
    fx, weights = C.load_esrnerf_case(C.ESRNERF_CASES[0])
    m = C.build_product_esrnerf(fx, weights, DEV)
    rays = S.make_rays(n, 77)
    return m, {k: rays[k] for k in ("rays_o", "rays_d", "viewdirs")}


def test_update_ray_groups_real_renderer_vs_restated_sweep():
    """the sweep that re-partitions the ray groups (pdra.py:882-932) through ESRNeRF.eval_emit on the GPU, sharded-capable
    product function vs the reference's loop restated on the port sampler (full gather, other chunk size, max over
    channels, filter): identical masks and groups — rays are independent in every kernel of the path"""
    n, keys = 20_000, ["rays_o", "rays_d", "viewdirs"]
    model, data = _esrnerf_and_rays(n)
    torch.manual_seed(4)
    mine = SM.RayGroupManager(_cfg(), dict(data), keys, 4096, 2048)
    mine.shuffle()
    torch.manual_seed(4)
    port = SP.RayGroupManagerPort(DEV, dict(data), keys, 4096, 2048)
    port.shuffle()
    model.train()
    for sweep, (k_frac, bs_mine, bs_port) in enumerate(((0.5, 3000, 4096), (0.8, 8192, 1777))):
        ro, rd, vd = (port.uncert_data[k] for k in keys)
        emission = torch.zeros_like(ro)
        model.eval()
        for idx in torch.arange(len(emission), device=DEV).split(bs_port):
            emission[idx] = model.eval_emit(rays_o=ro[idx], rays_d=rd[idx], viewdirs=vd[idx])
        model.train()
        peak = emission.max(dim=-1)[0]
        k_val = float(torch.quantile(peak[peak > 0], k_frac)) if (peak > 0).any() else 0.0     # a threshold that splits the set
        want = peak > k_val
        port.filter(want)
        got = SM.update_ray_groups(model, mine, k_val, bs_mine)
        assert torch.equal(got, want) and 0 < int(got.sum()) < got.numel(), sweep
        assert torch.equal(mine.uncert_data_idxs, port.uncert_data_idxs) and torch.equal(mine.cert_data_idxs, port.cert_data_idxs)
        assert model.training


class _Replay:
    """a renderer whose eval_esp returns the surface points another run produced for the same rays (keyed by ray origin)"""

    def __init__(self, rays_o, esp):
        self.training = False
        self.table = {tuple(o.tolist()): e for o, e in zip(rays_o.cpu(), esp.cpu())}

    def eval(self):
        pass

    def train(self):
        pass

    def eval_esp(self, rays_o, rays_d, viewdirs):
        return torch.stack([self.table[tuple(o.tolist())] for o in rays_o.cpu()]).to(rays_o.device)


def test_filter_edit_rays_real_renderer_cuda_vs_cpu_function():
    n, keys, w, h = 3000, ["rays_o", "rays_d", "viewdirs", "em_modes"], 48, 40
    model, data = _esrnerf_and_rays(n)
    data["em_modes"] = torch.ones(n, dtype=torch.long)
    g = torch.Generator().manual_seed(2)
    masks = torch.zeros(3, h, w)
    masks[0, 5:20, 8:30] = 1
    masks[1, 15:35, 20:44] = 1
    masks[2, 2:6, 2:6] = 1
    pose = torch.eye(4)
    pose[:3, 3] = torch.tensor([0.1, -0.2, 2.5])
    test_data = {"poses": pose, "em_masks": masks.view(3, -1), "em_modes": torch.tensor([2, 4, 0]),
                 "em_intensities": torch.rand(3, generator=g), "em_colors": torch.rand(3, 3, generator=g)}

    def make(device):
        torch.manual_seed(9)
        m = SM.RayGroupManager(_cfg(device), {k: v.clone() for k, v in data.items()}, list(keys), 256, 128)
        m.filter(torch.arange(n) % 7 != 0)          # some certain rays beforehand
        return m

    gpu = make(DEV)
    model.eval()
    esp = model.eval_esp(rays_o=data["rays_o"].to(DEV), rays_d=data["rays_d"].to(DEV), viewdirs=data["viewdirs"].to(DEV))
    SM.filter_edit_rays(model, gpu, test_data, (w, h), 30.0, 4, 700)
    cpu = make("cpu")
    SM.filter_edit_rays(_Replay(data["rays_o"], esp), cpu, test_data, (w, h), 30.0, 4, 1000)
    assert gpu.keys == cpu.keys and "em_colors" in gpu.keys and "em_intensities" in gpu.keys
    # the projection runs in fp32 on both devices; a ray within rounding of a mask edge may fall either way
    a, b = set(gpu.uncert_data_idxs.cpu().tolist()), set(cpu.uncert_data_idxs.tolist())
    assert len(a) > 20 and len(a ^ b) <= max(2, len(a) // 200), (len(a), len(b), len(a ^ b))
    both = sorted(a & b)
    pos_g = {int(v): i for i, v in enumerate(gpu.uncert_data_idxs.cpu().tolist())}
    pos_c = {int(v): i for i, v in enumerate(cpu.uncert_data_idxs.tolist())}
    ig, ic = torch.tensor([pos_g[v] for v in both]), torch.tensor([pos_c[v] for v in both])
    for k in ("em_modes", "em_colors", "em_intensities"):
        vg, vc = gpu.uncert_data[k].cpu()[ig], cpu.uncert_data[k][ic]
        same = (vg == vc).reshape(len(both), -1).all(dim=1)
        assert int((~same).sum()) <= max(2, len(both) // 200), k
    assert gpu.cert_data_num + gpu.uncert_data_num == n
    batch = gpu.sample()
    assert set(batch) == set(gpu.keys) | {"uncert_masks"} and all(v.device.type == "cuda" for v in batch.values())


def test_scale_volume_grid_on_device_matches_cpu_and_renders():
    from esr_nerf_b200.voxurff import VoxurfF

    def make(device):
        torch.manual_seed(0)
        geo = (S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT, S.mask_density(12, True))
        m = VoxurfF(S.fine_cfg(device), *geo, 20.0, 20 ** 3)
        S.fill_fine_model(m)
        return m

    cpu, gpu = make("cpu"), make(DEV)
    gpu.load_state_dict(cpu.state_dict(), strict=True)
    for num_voxels in (27 ** 3, 33 ** 3 + 100):
        cpu.scale_volume_grid(num_voxels)
        gpu.scale_volume_grid(num_voxels)
        assert torch.equal(torch.as_tensor(gpu.world_size).cpu(), torch.as_tensor(cpu.world_size).cpu())
        # ((volume / num_voxels) ** (1 / 3) on the model's device: the GPU's pow rounds differently in the last place)
        assert torch.allclose(torch.as_tensor(gpu.voxel_size).cpu(), torch.as_tensor(cpu.voxel_size).cpu(), rtol=1e-6, atol=0)
        csd, gsd = cpu.state_dict(), gpu.state_dict()
        assert list(csd) == list(gsd)
        for k in csd:
            assert csd[k].shape == gsd[k].shape, k
            # trilinear re-sampling: same taps and weights, the device's fused multiply-adds round differently
            assert torch.allclose(gsd[k].cpu().float(), csd[k].float(), rtol=1e-5, atol=1e-6), (num_voxels, k)
        diff = (gpu.nonempty_mask.cpu() != cpu.nonempty_mask)
        assert int(diff.sum()) <= max(2, diff.numel() // 10000)      # (a density within rounding of the threshold)
    # the rescaled model renders and back-propagates through the hand-written path
    rays = {k: v.to(DEV) for k, v in S.make_rays(2048, 5).items()}
    gpu.train()
    out = gpu(s_val=20.0, **{k: v for k, v in rays.items() if k != "rgbs"})
    ((out["srgb/rgb"] - rays["rgbs"]) ** 2).mean().backward()
    assert torch.isfinite(out["srgb/rgb"]).all() and gpu.sdf.grid.grad is not None and torch.isfinite(gpu.sdf.grid.grad).all()
    assert tuple(gpu.sdf.grid.shape[2:]) == tuple(int(v) for v in gpu.world_size)

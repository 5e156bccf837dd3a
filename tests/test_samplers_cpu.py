"""CPU tests of the ray feed: esr_nerf_b200.samplers against the reference's own BatchSampler /
RayGroupManager (utils2/utils.py:41-312, imported from /root/reference where it exists) under the same seeds — same
batches, same checkpointed state — and, without the reference, against the invariants its code implies."""
import importlib.util
import os
import sys
import types

import pytest
import torch

from esr_nerf_b200 import samplers as SM
from esr_nerf_b200.dist import shard_slice

KEYS = ["rays_o", "rgbs", "em_modes"]


def _cfg(preload="cuda"):
    return types.SimpleNamespace(system=types.SimpleNamespace(device="cpu", data_preload=preload))


def _data(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"rays_o": torch.randn(n, 3, generator=g), "rgbs": torch.rand(n, 3, generator=g),
            "em_modes": torch.randint(0, 2, (n,), generator=g), "unused": torch.zeros(n)}


def _reference_module():
    from oracle import ref_harness as H

    if not H.reference_available():
        return None
    H.install_stubs()                                    # omegaconf / wandb stand-ins (rich and tqdm are installed)
    spec = importlib.util.spec_from_file_location("_ref_utils2_utils", os.path.join(H.REF_ROOT, "utils2", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _same_batch(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_batch_sampler_matches_reference_sequence():
    ref_mod = _reference_module()
    if ref_mod is None:
        pytest.skip("/root/reference not present")
    n, bs = 1000, 96
    g = torch.Generator().manual_seed(3)
    mask = torch.rand(n, generator=g) < 0.7
    torch.manual_seed(11)
    ref = ref_mod.BatchSampler(_cfg(), _data(n), KEYS, bs)
    ref.filter(mask)
    ref.shuffle()
    ref_batches = [ref.sample() for _ in range(20)]      # wraps (re-shuffles) twice
    torch.manual_seed(11)
    mine = SM.BatchSampler(_cfg(), _data(n), KEYS, bs)
    mine.filter(mask)
    mine.shuffle()
    for rb in ref_batches:
        _same_batch(mine.sample(), rb)
    assert mine.batch_st == ref.batch_st and torch.equal(mine.data_idxs, ref.data_idxs) and mine.data_num == ref.data_num
    for k in KEYS:                                       # the physical copy the reference holds == the gathered view
        assert torch.equal(mine.data[k], ref.data[k])
    # resume from the checkpointed state (fine.py:221-228): same continuation
    torch.manual_seed(5)
    ref2 = ref_mod.BatchSampler(_cfg(), _data(n), KEYS, bs, ref.batch_st, ref.data_idxs.clone())
    tail = [ref2.sample() for _ in range(12)]
    torch.manual_seed(5)
    mine2 = SM.BatchSampler(_cfg(), _data(n), KEYS, bs, mine.batch_st, mine.data_idxs.clone())
    for rb in tail:
        _same_batch(mine2.sample(), rb)


def test_ray_group_manager_matches_reference_sequence():
    ref_mod = _reference_module()
    if ref_mod is None:
        pytest.skip("/root/reference not present")
    n = 700

    def drive(cls):
        torch.manual_seed(21)
        m = cls(_cfg(), _data(n), KEYS, 64, 32)
        m.shuffle()
        out = [m.sample() for _ in range(3)]             # no certain rays yet: every uncert_mask is False (:302 quirk)
        g = torch.Generator().manual_seed(8)
        m.filter(torch.rand(m.uncert_data_num, generator=g) < 0.6)
        out += [m.sample() for _ in range(15)]           # both groups wrap
        m.filter(torch.rand(m.uncert_data_num, generator=g) < 0.05)     # uncertain group smaller than its batch size
        out += [m.sample() for _ in range(4)]
        return m, out

    ref, ref_out = drive(ref_mod.RayGroupManager)
    mine, my_out = drive(SM.RayGroupManager)
    assert not ref_out[0]["uncert_masks"].any()
    for a, b in zip(my_out, ref_out):
        _same_batch(a, b)
    for name in ("uncert_batch_st", "cert_batch_st", "uncert_data_num", "cert_data_num"):
        assert getattr(mine, name) == getattr(ref, name), name
    assert torch.equal(mine.uncert_data_idxs, ref.uncert_data_idxs) and torch.equal(mine.cert_data_idxs, ref.cert_data_idxs)
    assert torch.equal(mine.uncert_data["rays_o"], ref.uncert_data["rays_o"])       # pdra.py:888
    assert torch.equal(mine.cert_data["rgbs"], ref.cert_data["rgbs"])


@pytest.mark.parametrize("preload", ["cuda", "cpu"])
def test_sampler_invariants_and_rank_slices(preload):
    """without the reference: every epoch visits each surviving ray once; the host-preload mode (physical copies) and the
    index-only mode agree; the ranks' slices concatenate to the single-process batch, last ray on the last rank"""
    n, bs = 500, 64
    data = _data(n, 4)
    mask = torch.arange(n) % 5 != 0
    steps = 2 * (400 // bs)

    def drive(mode, rank=0, world=1):        # (the re-shuffles at the epoch wraps draw from the global generator)
        torch.manual_seed(1)
        s = SM.BatchSampler(_cfg(mode), dict(data), KEYS, bs, rank=rank, world=world)
        s.filter(mask)
        s.shuffle()
        out, seen = [], []
        for step in range(steps):
            out.append(s.sample())
            if step < 400 // bs:
                seen.append(s.data_idxs[s.batch_st - bs:s.batch_st].clone())
        return s, out, torch.cat(seen)

    one, batches, seen = drive(preload)
    _, other_batches, _ = drive("cuda" if preload == "cpu" else "cpu")
    parts = [drive(preload, r, 3)[1] for r in range(3)]
    for i, b in enumerate(batches):
        _same_batch(b, other_batches[i])
        for k in KEYS:
            assert torch.equal(torch.cat([p[i][k] for p in parts]), b[k])
        assert parts[-1][i]["rays_o"].shape[0] == shard_slice(bs, 2, 3).stop - shard_slice(bs, 2, 3).start
        assert torch.equal(parts[-1][i]["rays_o"][-1], b["rays_o"][-1])
    assert seen.unique().numel() == seen.numel() and mask[seen].all()
    for k in KEYS:
        assert torch.equal(one.data[k], data[k][one.data_idxs])

    def drive_groups(rank=0, world=1):
        torch.manual_seed(2)
        m = SM.RayGroupManager(_cfg(preload), dict(data), KEYS, 40, 24, rank=rank, world=world)
        m.filter(torch.arange(n) % 2 == 0)
        return [m.sample() for _ in range(12)]

    full, halves = drive_groups(), [drive_groups(r, 2) for r in range(2)]
    for i, b in enumerate(full):
        for k in KEYS + ["uncert_masks"]:
            assert torch.equal(torch.cat([h[i][k] for h in halves]), b[k])
        assert int(b["uncert_masks"].sum()) == 40 and b["uncert_masks"].numel() == 64


def _drive_batch(make):
    n, bs = 1000, 96
    g = torch.Generator().manual_seed(3)
    mask = torch.rand(n, generator=g) < 0.7
    torch.manual_seed(11)
    s = make(_data(n), bs)
    s.filter(mask)
    s.shuffle()
    return s, [s.sample() for _ in range(20)]


def _drive_groups(make, n=700):
    torch.manual_seed(21)
    m = make(_data(n), 64, 32)
    m.shuffle()
    out = [m.sample() for _ in range(3)]
    g = torch.Generator().manual_seed(8)
    m.filter(torch.rand(m.uncert_data_num, generator=g) < 0.6)
    out += [m.sample() for _ in range(15)]
    m.filter(torch.rand(m.uncert_data_num, generator=g) < 0.05)
    out += [m.sample() for _ in range(4)]
    return m, out


def test_port_matches_reference():
    """oracle/samplers_port.py (what tests/test_gpu_feed.py compares the product with on the GPU) against the reference's
    own classes: same batches, same state, both samplers"""
    from oracle import samplers_port as SP

    ref_mod = _reference_module()
    if ref_mod is None:
        pytest.skip("/root/reference not present")
    ref, ref_b = _drive_batch(lambda d, bs: ref_mod.BatchSampler(_cfg(), d, KEYS, bs))
    port, port_b = _drive_batch(lambda d, bs: SP.BatchSamplerPort("cpu", d, KEYS, bs))
    for a, b in zip(port_b, ref_b):
        _same_batch(a, b)
    assert port.batch_st == ref.batch_st and torch.equal(port.data_idxs, ref.data_idxs)
    refg, refg_b = _drive_groups(lambda d, a, b: ref_mod.RayGroupManager(_cfg(), d, KEYS, a, b))
    portg, portg_b = _drive_groups(lambda d, a, b: SP.RayGroupManagerPort("cpu", d, KEYS, a, b))
    for a, b in zip(portg_b, refg_b):
        _same_batch(a, b)
    for name in ("uncert_batch_st", "cert_batch_st", "uncert_data_num", "cert_data_num"):
        assert getattr(portg, name) == getattr(refg, name), name
    assert torch.equal(portg.uncert_data_idxs, refg.uncert_data_idxs) and torch.equal(portg.cert_data_idxs, refg.cert_data_idxs)
    assert torch.equal(portg.cert_data["rgbs"], refg.cert_data["rgbs"])


class _FakeRenderer:
    """eval_emit as a per-ray function (what the render path is: rays are independent), recording its chunk sizes"""

    def __init__(self):
        self.training, self.chunks = True, []

    def eval(self):
        self.training = False

    def train(self):
        self.training = True

    def eval_emit(self, rays_o, rays_d, viewdirs):
        self.chunks.append(rays_o.shape[0])
        return (rays_o * viewdirs).abs() + rays_d[:, :1] ** 2


def _sweep_data(n):
    g = torch.Generator().manual_seed(9)
    return {k: torch.randn(n, 3, generator=g) for k in ("rays_o", "rays_d", "viewdirs")}


def _reference_update(renderer, sampler, k_val, bs):
    """pdra.py:882-932 restated on the sampler's public state: full gather, chunk loop, max over channels, filter"""
    ro, rd, vd = (sampler.uncert_data[k] for k in ("rays_o", "rays_d", "viewdirs"))
    emission = torch.zeros_like(ro)
    for idx in torch.arange(len(emission)).split(bs):
        emission[idx] = renderer.eval_emit(rays_o=ro[idx], rays_d=rd[idx], viewdirs=vd[idx])
    mask = torch.max(emission, dim=-1)[0] > k_val
    sampler.filter(mask)
    return mask


@pytest.mark.parametrize("preload", ["cuda", "cpu"])
def test_update_ray_groups_matches_restated_sweep(preload):
    n, keys = 1000, ["rays_o", "rays_d", "viewdirs"]
    torch.manual_seed(4)
    a = SM.RayGroupManager(_cfg(preload), dict(_sweep_data(n)), keys, 64, 32)
    a.shuffle()
    torch.manual_seed(4)
    b = SM.RayGroupManager(_cfg(preload), dict(_sweep_data(n)), keys, 64, 32)
    b.shuffle()
    r = _FakeRenderer()
    for k_val in (0.3, 0.8):                              # two successive sweeps: the second over the survivors
        want = _reference_update(_FakeRenderer(), b, k_val, 96)
        r.chunks.clear()
        got = SM.update_ray_groups(r, a, k_val, 96)
        assert torch.equal(got, want) and 0 < int(got.sum()) < got.numel()
        assert torch.equal(a.uncert_data_idxs, b.uncert_data_idxs) and torch.equal(a.cert_data_idxs, b.cert_data_idxs)
        assert r.training and max(r.chunks) <= 96 and sum(r.chunks) == got.numel()


def _sweep_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys = ["rays_o", "rays_d", "viewdirs"]
    torch.manual_seed(4)
    m = SM.RayGroupManager(_cfg(), dict(_sweep_data(1001)), keys, 64, 32, rank=rank, world=world)
    m.shuffle()
    r = _FakeRenderer()
    mask = SM.update_ray_groups(r, m, 0.3, 96, rank=rank, world=world)
    # plain numpy payloads: torch tensors travel as shared-memory handles that die with this process
    out.put((rank, sum(r.chunks), mask.numpy().copy(), m.uncert_data_idxs.numpy().copy(), m.cert_data_idxs.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_update_ray_groups_two_ranks_agree_with_one():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    keys = ["rays_o", "rays_d", "viewdirs"]
    torch.manual_seed(4)
    single = SM.RayGroupManager(_cfg(), dict(_sweep_data(1001)), keys, 64, 32)
    single.shuffle()
    want = _reference_update(_FakeRenderer(), single, 0.3, 96)
    assert res[0][1] + res[1][1] == 1001 and abs(res[0][1] - res[1][1]) <= 1        # each rank swept its half
    for _, _, mask, unc, cert in res:
        mask, unc, cert = torch.from_numpy(mask), torch.from_numpy(unc), torch.from_numpy(cert)
        assert torch.equal(mask, want)
        assert torch.equal(unc, single.uncert_data_idxs) and torch.equal(cert, single.cert_data_idxs)


def test_group_attributes_attached_like_filter_edit_rays():
    """pdra.py:1023-1037: per-ray edit attributes are ASSIGNED to both groups' data mappings (in group order), two keys
    are appended to sampler.keys, then the uncertain group is filtered; the batches that follow carry the new keys"""
    ref_mod = _reference_module()
    if ref_mod is None:
        pytest.skip("/root/reference not present")
    n = 400

    def drive(cls):
        torch.manual_seed(31)
        data = _data(n, 6)
        keep = data["em_modes"].clone()
        m = cls(_cfg(), data, list(KEYS), 48, 16)
        m.shuffle()
        g = torch.Generator().manual_seed(2)
        m.filter(torch.rand(n, generator=g) < 0.5)
        nu, nc = m.uncert_data_num, m.cert_data_num
        m.uncert_data["em_modes"] = torch.randint(0, 5, (nu,), generator=g)
        m.uncert_data["em_colors"] = torch.rand(nu, 2, generator=g)
        m.uncert_data["em_intensities"] = torch.rand(nu, generator=g)
        m.cert_data["em_modes"] = torch.zeros(nc, dtype=torch.long)
        m.cert_data["em_colors"] = torch.zeros(nc, 2)
        m.cert_data["em_intensities"] = torch.zeros(nc)
        m.keys.extend(["em_colors", "em_intensities"])
        m.filter(torch.rand(nu, generator=g) < 0.7)
        return m, [m.sample() for _ in range(10)], keep, data

    ref, ref_out, _, _ = drive(ref_mod.RayGroupManager)
    mine, my_out, keep, data = drive(SM.RayGroupManager)
    assert set(my_out[0]) == set(KEYS) | {"em_colors", "em_intensities", "uncert_masks"}
    for a, b in zip(my_out, ref_out):
        _same_batch(a, b)
    for k in mine.keys:
        assert torch.equal(mine.uncert_data[k], ref.uncert_data[k]) and torch.equal(mine.cert_data[k], ref.cert_data[k])
    assert torch.equal(data["em_modes"], keep)           # the caller's tensor was not written through


def _reference_pdra_class():
    """the reference's PDRA trainer class (app/fine/pdra.py), imported only for its `filter_edit_rays` method: the
    trainer's other dependencies (data sets, metrics, config managers) are given empty stand-ins"""
    from oracle import ref_harness as H

    if not H.reference_available():
        return None
    H.install_stubs()
    ref_utils = _reference_module()

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("app", AppClass=object)
    mod("app.fine.model", ESRNeRF=object)
    mod("data", DataClass=object)
    mod("utils2.image", apply_gamma_curve=None)
    mod("utils2.manager", save_cfg=None)
    mod("utils2.metric", IoU=None, loss2psnr=None, rgb_lpips=None, rgb_ssim=None)
    mod("utils2.utils", LightDict=ref_utils.LightDict, RayGroupManager=ref_utils.RayGroupManager,
        import_class=ref_utils.import_class, tqdm_safe=lambda it, **kw: it)
    spec = importlib.util.spec_from_file_location("_ref_app_fine_pdra", os.path.join(H.REF_ROOT, "app", "fine", "pdra.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.PDRA, ref_utils


class _FakeEsp(_FakeRenderer):
    """eval_esp: a point on each ray (per-ray function)"""

    def eval_esp(self, rays_o, rays_d, viewdirs):
        self.chunks.append(rays_o.shape[0])
        return rays_o + 2.5 * viewdirs


def _edit_case(n=1500, w=48, h=40):
    g = torch.Generator().manual_seed(12)
    o = torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -3.0])
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, 1.0]), dim=-1)
    data = {"rays_o": o, "rays_d": d * 1.1, "viewdirs": d, "em_modes": torch.ones(n, dtype=torch.long)}
    pose = torch.eye(4)
    pose[:3, :3] = torch.tensor([[-1.0, 0, 0], [0, 1.0, 0], [0, 0, -1.0]])       # camera looking down +z from z = -3
    pose[:3, 3] = torch.tensor([0.0, 0.0, -3.0])
    masks = torch.zeros(4, h, w)
    masks[0, 5:12, 6:20] = 1
    masks[1, 18:30, 25:40] = 1
    masks[2, 8:25, 15:30] = 1            # overlaps 0 and 1: later masks overwrite
    masks[3, 30:38, 2:10] = 1
    test = {"poses": pose, "em_masks": masks.reshape(4, -1), "em_modes": torch.tensor([0, 2, 4, 3]),
            "em_intensities": torch.tensor([0.0, 2.5, 0.3, 1.0]), "em_colors": torch.rand(4, 3, generator=g)}
    return data, test, (w, h), 30.0


def test_filter_edit_rays_matches_reference_method():
    got = _reference_pdra_class()
    if got is None:
        pytest.skip("/root/reference not present")
    PDRA, ref_utils = got
    keys = ["rays_o", "rays_d", "viewdirs", "em_modes"]
    data, test, (w, h), focal = _edit_case()
    torch.manual_seed(3)
    ref_s = ref_utils.RayGroupManager(_cfg(), dict(data), list(keys), 64, 32)
    ref_s.shuffle()
    ref_s.filter(torch.arange(ref_s.uncert_data_num) % 7 != 0)           # a non-empty certain group
    me = types.SimpleNamespace(train_dataset=types.SimpleNamespace(image_size=(w, h), focal_length=focal), device="cpu",
                               mask_dilation_ks=10, eval_bs=256, renderer=_FakeEsp())
    PDRA.filter_edit_rays(me, ref_s, test)
    torch.manual_seed(3)
    mine = SM.RayGroupManager(_cfg(), dict(data), list(keys), 64, 32)
    mine.shuffle()
    mine.filter(torch.arange(mine.uncert_data_num) % 7 != 0)
    r = _FakeEsp()
    SM.filter_edit_rays(r, mine, test, (w, h), focal, 10, 256)
    assert 0 < mine.uncert_data_num < 1500 and r.training
    assert mine.keys == ref_s.keys
    assert torch.equal(mine.uncert_data_idxs, ref_s.uncert_data_idxs) and torch.equal(mine.cert_data_idxs, ref_s.cert_data_idxs)
    for k in mine.keys:
        assert torch.equal(mine.uncert_data[k], ref_s.uncert_data[k]), k
        assert torch.equal(mine.cert_data[k], ref_s.cert_data[k]), k
    assert len(set(mine.uncert_data["em_modes"].tolist())) >= 3          # several edits present among the kept rays
    torch.manual_seed(9)
    a = [mine.sample() for _ in range(6)]
    torch.manual_seed(9)
    b = [ref_s.sample() for _ in range(6)]
    for x, y in zip(a, b):
        _same_batch(x, y)


def _edit_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data, test, size, focal = _edit_case()
    torch.manual_seed(3)
    m = SM.RayGroupManager(_cfg(), dict(data), ["rays_o", "rays_d", "viewdirs", "em_modes"], 64, 32, rank=rank, world=world)
    m.shuffle()
    r = _FakeEsp()
    SM.filter_edit_rays(r, m, test, size, focal, 10, 256, rank=rank, world=world)
    # numpy payloads (see _sweep_worker)
    out.put((rank, sum(r.chunks), m.uncert_data_idxs.numpy().copy(), m.cert_data_idxs.numpy().copy(),
             {k: m.uncert_data[k].numpy().copy() for k in m.keys}))
    dist.barrier()
    dist.destroy_process_group()


def test_filter_edit_rays_two_ranks_agree_with_one():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_edit_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    data, test, size, focal = _edit_case()
    torch.manual_seed(3)
    single = SM.RayGroupManager(_cfg(), dict(data), ["rays_o", "rays_d", "viewdirs", "em_modes"], 64, 32)
    single.shuffle()
    SM.filter_edit_rays(_FakeEsp(), single, test, size, focal, 10, 256)
    assert res[0][1] + res[1][1] == 1500 and res[0][1] == res[1][1]
    for _, _, unc, cert, vals in res:
        assert torch.equal(torch.from_numpy(unc), single.uncert_data_idxs) and torch.equal(torch.from_numpy(cert), single.cert_data_idxs)
        for k in single.keys:
            assert torch.equal(torch.from_numpy(vals[k]), single.uncert_data[k]), k

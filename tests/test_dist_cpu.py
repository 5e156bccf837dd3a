"""world_size-2 gloo test (CPU) of the multi-GPU host logic: ray sharding + one gradient all-reduce per step must
reproduce the single-process gradient of a sum-over-rays loss."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from esr_nerf_b200 import dist as D

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
    g = torch.Generator().manual_seed(1)
    n = 101                                                   # not divisible by the world size
    batch = {"rays_o": torch.randn(n, 3, generator=g), "rays_d": torch.randn(n, 3, generator=g),
             "rgbs": torch.rand(n, 3, generator=g), "s_val": 20.0}
    local = D.shard_batch(batch, rank, world)
    assert local["s_val"] == 20.0
    pred = net(torch.cat([local["rays_o"], local["rays_d"]], -1))
    loss = ((pred - local["rgbs"]) ** 2).sum() / n           # sum over local rays / GLOBAL ray count
    loss.backward()
    # `unused` received no gradient on any rank, `half_used` only on rank 1: the first stays None (the optimizer must skip
    # it as the single-process run does), the second is summed with zeros from rank 0
    unused = torch.nn.Parameter(torch.ones(5))
    half_used = torch.nn.Parameter(torch.ones(4))
    if rank == 1:
        half_used.grad = torch.full((4,), 3.0)
    nbytes = D.allreduce_gradients(list(net.parameters()) + [frozen, unused, half_used])
    # (zeros travel for `unused` too: the collective sequence may not depend on what a rank happens to hold)
    assert nbytes == sum(p.numel() * 4 for p in net.parameters()) + 4 * 4 + 5 * 4 + 4 * (len(list(net.parameters())) + 2)
    assert unused.grad is None and torch.equal(half_used.grad, torch.full((4,), 3.0))
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
        ref.load_state_dict(net.state_dict())
        rl = ((ref(torch.cat([batch["rays_o"], batch["rays_d"]], -1)) - batch["rgbs"]) ** 2).sum() / n
        rl.backward()
        err = max((a.grad - b.grad).abs().max().item() for a, b in zip(net.parameters(), ref.parameters()))
        out.put(err)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_slices_cover_batch():
    from esr_nerf_b200.dist import shard_slice

    for n in (0, 1, 7, 64, 101):
        for world in (1, 2, 3, 8):
            idx = []
            for r in range(world):
                sl = shard_slice(n, r, world)
                idx += list(range(n))[sl]
            assert idx == list(range(n)), (n, world)
    assert shard_slice(65536, 7, 8) == slice(57344, 65536)    # last global ray on the last rank


def test_two_rank_gradient_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-6


class _TinyGrids(torch.nn.Module):
    """the attributes GridGradCompactor reads from a VoxurfF: three grids in the parameters' memory layout + the mask"""

    def __init__(self, r=9):
        super().__init__()
        mk = lambda c: torch.nn.Parameter(torch.zeros(1, c, r, r, r).contiguous(
            memory_format=torch.channels_last_3d if c > 1 else torch.contiguous_format))
        self.sdf, self.off_color, self.emo_color = (torch.nn.Module() for _ in range(3))
        self.sdf.grid, self.off_color.grid, self.emo_color.grid = mk(1), mk(6), mk(6)
        self.head = torch.nn.Linear(4, 2)
        m = torch.zeros(1, 1, r, r, r, dtype=torch.bool)
        m[..., 3:6, 3:6, 4] = True
        self.nonempty_mask = m


def _compactor_worker(rank, world, port, out):
    import torch.distributed as dist

    from esr_nerf_b200.dist import GridGradCompactor

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _TinyGrids()
    comp = GridGradCompactor(model, dilate=1)
    assert comp.idx.numel() % 2 == 1                          # odd voxel count: the sdf block is padded to 8 bytes
    g = torch.Generator().manual_seed(10 + rank)
    grads = []
    for p in model.parameters():
        v = torch.randn(p.shape, generator=g)
        if p.dim() == 5:                                      # grid gradients live inside the dilated set only
            v = (v * comp.mask).contiguous(memory_format=torch.channels_last_3d if p.shape[1] > 1
                                           else torch.contiguous_format)
        p.grad = v.clone()
        grads.append(v)
    nbytes = comp.allreduce(verify=True)
    assert nbytes == 4 * (sum(p.numel() for p in model.head.parameters()) + 2 * ((comp.idx.numel() * 13 + 2) // 2)
                          + len(list(model.head.parameters())))        # + the has-gradient vector
    g2 = torch.Generator().manual_seed(10 + (1 - rank))
    err = 0.0
    for p, mine in zip(model.parameters(), grads):
        other = torch.randn(p.shape, generator=g2)
        if p.dim() == 5:
            other = other * comp.mask
        err = max(err, (p.grad - (mine + other)).abs().max().item())
    # the same exchange with the colour volumes started early (what Shade.backward's hook does on the GPU): the packed
    # colour block is all-reduced asynchronously while the caller still finishes the SDF gradient
    color = comp.grids[1:]
    for p, mine in zip(model.parameters(), grads):
        p.grad = None if any(p is c for c in color) else mine.clone()
    bufs = {p: mine.clone() for p, mine in zip(model.parameters(), grads) if any(p is c for c in color)}
    comp._on_color_grads(bufs)
    assert comp._early is not None
    for p in color:
        p.grad = bufs[p]
    comp.allreduce()
    assert comp._early is None
    g3 = torch.Generator().manual_seed(10 + (1 - rank))
    for p, mine in zip(model.parameters(), grads):
        other = torch.randn(p.shape, generator=g3)
        if p.dim() == 5:
            other = other * comp.mask
        err = max(err, (p.grad - (mine + other)).abs().max().item())
    # rank-dependent data must not change the collective sequence: (A) rank 1's rays never touched the second colour
    # volume (it is missing from the hook's buffers), (B) rank 1's backward never reached the hook at all — in both
    # cases rank 1 contributes zeros through the SAME early all-reduce, and nothing hangs or mismatches
    comp._overlap = True
    for case in ("A", "B"):
        mine_c = []
        for p, mine in zip(model.parameters(), grads):
            is_color = any(p is c for c in color)
            absent = is_color and rank == 1 and (case == "B" or p is color[1])
            p.grad = None if is_color else mine.clone()
            mine_c.append(torch.zeros_like(mine) if absent else mine)
        bufs = {p: m.clone() for p, m in zip(model.parameters(), mine_c)
                if any(p is c for c in color) and not (rank == 1 and (case == "B" or p is color[1]))}
        if not (rank == 1 and case == "B"):
            comp._on_color_grads(bufs)
            for p in color:
                if p in bufs:
                    p.grad = bufs[p]
        comp.allreduce()
        assert comp._early is None
        g4 = torch.Generator().manual_seed(10 + (1 - rank))
        for p, m in zip(model.parameters(), mine_c):
            other = torch.randn(p.shape, generator=g4)
            if p.dim() == 5:
                other = other * comp.mask
                if rank == 0 and (case == "B" or p is color[1]) and any(p is c for c in color):
                    other = other * 0          # what rank 1 left out
            err = max(err, (p.grad - (m + other)).abs().max().item())
    # gradient accumulation with the early start on: the first backward's hook starts the exchange, the second backward
    # adds to the SAME buffers in place (what fused._GradSink.flush does: p.grad.add_) and calls the hook again — what is
    # in flight is stale, allreduce() must redo the exchange with the accumulated gradients
    for p, mine in zip(model.parameters(), grads):
        p.grad = None if any(p is c for c in color) else 2 * mine.clone()
    bufs = {p: mine.clone() for p, mine in zip(model.parameters(), grads) if any(p is c for c in color)}
    comp._on_color_grads(bufs)
    for p in color:
        p.grad = bufs[p]
    assert comp._early is not None and not comp._early_stale
    for p in color:
        p.grad.add_(bufs[p].clone())                          # second micro-batch, accumulated in place
    comp._on_color_grads({p: p.grad for p in color})
    assert comp._early_stale
    comp.allreduce()
    assert comp._early is None and not comp._early_stale
    g5 = torch.Generator().manual_seed(10 + (1 - rank))
    for p, mine in zip(model.parameters(), grads):
        other = torch.randn(p.shape, generator=g5)
        if p.dim() == 5:
            other = other * comp.mask
        err = max(err, (p.grad - 2 * (mine + other)).abs().max().item())
    out.put(err)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_compacted_grid_allreduce_equals_dense_sum():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_compactor_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert max(out.get(timeout=5), out.get(timeout=5)) == 0.0


class _LtsGrids(torch.nn.Module):
    """the four grid-gradient volumes of the LTS stage (sdf + off / emo / brdf colour grids, parameters' memory layout);
    20 is not a multiple of 8: that axis falls back to 5-voxel blocks"""

    def __init__(self, shape=(32, 32, 20)):
        super().__init__()
        mk = lambda c: torch.nn.Parameter(torch.zeros(1, c, *shape).contiguous(
            memory_format=torch.channels_last_3d if c > 1 else torch.contiguous_format))
        self.sdf, self.off_color, self.emo_color, self.brdf = (torch.nn.Module() for _ in range(4))
        self.sdf.grid, self.off_color.grid, self.emo_color.grid, self.brdf.grid = mk(1), mk(6), mk(6), mk(6)
        self.head = torch.nn.Linear(4, 2)


def _sparse_grads(model, seed):
    """random gradients confined to a few random voxels per volume (different voxels per rank and per volume)"""
    g = torch.Generator().manual_seed(seed)
    out = []
    for p in model.parameters():
        v = torch.randn(p.shape, generator=g)
        if p.dim() == 5:
            keep = torch.zeros(p.shape[2:], dtype=torch.bool)
            n = keep.numel()
            keep.view(-1)[torch.randint(0, n, (2,), generator=g)] = True
            v = (v * keep).contiguous(memory_format=torch.channels_last_3d if p.shape[1] > 1 else torch.contiguous_format)
        out.append(v)
    return out


def _block_worker(rank, world, port, out):
    import torch.distributed as dist

    from esr_nerf_b200.dist import TouchedBlockCompactor

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _LtsGrids()
    comp = TouchedBlockCompactor(model)
    assert comp.edge == (8, 8, 5) and comp.blocks == (4, 4, 4) and len(comp.grids) == 4
    err, sizes = 0.0, []
    for step in range(3):                                     # the touched set changes every step; step 2: nothing touched
        mine = _sparse_grads(model, 100 * step + rank)
        other = _sparse_grads(model, 100 * step + (1 - rank))
        if step == 2:
            mine = [v * 0 if v.dim() == 5 else v for v in mine]
            other = [v * 0 if v.dim() == 5 else v for v in other]
        for p, v in zip(model.parameters(), mine):
            p.grad = v.clone()
        nbytes = comp.allreduce(verify=True)
        sizes.append((comp.idx.numel(), nbytes))
        for p, a, b in zip(model.parameters(), mine, other):
            err = max(err, (p.grad - (a + b)).abs().max().item())     # two-term sums: exact
        if step < 2:
            assert 0 < comp.fraction < 1 and comp.idx.numel() % (8 * 8 * 5) == 0
            assert nbytes < 4 * sum(p.numel() for p in model.parameters())      # less than the dense exchange
    assert sizes[2][0] == 0 and comp.fraction == 0.0
    out.put(err)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_touched_block_allreduce_equals_dense_sum():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_block_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert max(out.get(timeout=5), out.get(timeout=5)) == 0.0


def test_block_flag_kernel_index_math_restated():
    """csrc/grad_exchange.cu:k_grad_block_flags walks block b of volume j as ex * ey runs of ez * c floats at
    voxel ((bx ex + i) Y + by ey + jj) Z + bz ez.  The same arithmetic, restated with Python integers over a
    channels-last volume, must give the block map the torch reduction of TouchedBlockCompactor.block_flags gives."""
    from esr_nerf_b200.dist import TouchedBlockCompactor

    model = _LtsGrids((16, 12, 10))
    comp = TouchedBlockCompactor(model)
    (X, Y, Z), (ex, ey, ez), (Bx, By, Bz) = comp.shape, comp.edge, comp.blocks
    assert (ex, ey, ez) == (8, 6, 5)
    for p, v in zip(model.parameters(), _sparse_grads(model, 5)):
        p.grad = v.clone()
    rows = comp._grids_rows()
    want = comp.block_flags(rows)
    got = torch.zeros(Bx * By * Bz, dtype=torch.int32)
    for r in rows:
        flat, c = r.reshape(-1), r.shape[1]
        run = ez * c
        for b in range(Bx * By * Bz):
            bz, by, bx = b % Bz, (b // Bz) % By, b // (Bz * By)
            for e in range(ex * ey * run):
                rr, t = divmod(e, run)
                i, jj = divmod(rr, ey)
                voxel = ((bx * ex + i) * Y + (by * ey + jj)) * Z + bz * ez
                if flat[voxel * c + t] != 0:
                    got[b] = 1
                    break
    assert torch.equal(got, want) and 0 < int(want.sum()) < want.numel()


class _FakeImageModel:
    """forward_evaluate stand-in: per-ray maps of different widths (rays are independent on the render path)"""

    def __call__(self, rays_o, rays_d, viewdirs, em_modes=None, scale=1.0):
        # (elementwise, exactly rounded ops only: a vectorised transcendental may round differently at chunk tails)
        return {"srgb/rgb": rays_o * scale + 1.0, "etc/depth": rays_o[:, 0] * rays_d[:, 1],
                "etc/normal": viewdirs * 0.5 + 0.5, "etc/white_bg": rays_d[:, :1].abs()}


def _image_rays(n):
    g = torch.Generator().manual_seed(77)
    return {k: torch.randn(n, 3, generator=g) for k in ("rays_o", "rays_d", "viewdirs")}


def _image_worker(rank, world, port, out, n):
    import torch.distributed as dist

    from esr_nerf_b200.dist import render_image_sharded

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = render_image_sharded(_FakeImageModel(), _image_rays(n), rank, world, chunk=50, em_modes=torch.tensor(0), scale=2.0)
    # numpy payloads: torch tensors travel as shared-memory handles that die with this process
    out.put((rank, None if res is None else {k: v.numpy().copy() for k, v in res.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [301, 2])        # 301: the last rank's slice is shorter; 2 rays on 3 ranks: an empty slice
def test_sharded_image_render_gathers_every_map_on_rank0(n):
    from esr_nerf_b200.dist import render_image_sharded

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_image_worker, args=(r, 3, port, out, n)) for r in range(3)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = render_image_sharded(_FakeImageModel(), _image_rays(n), 0, 1, chunk=64, em_modes=torch.tensor(0), scale=2.0)
    assert res[1] is None and res[2] is None
    assert set(res[0]) == set(want)
    for k in want:
        got = torch.from_numpy(res[0][k])
        assert got.shape == want[k].shape and torch.equal(got, want[k]), k


@pytest.mark.parametrize("neus_alpha", ["interp", "grad"])
def test_static_exchange_set_covers_the_gradients_of_both_alpha_modes(neus_alpha):
    """GridGradCompactor exchanges only the voxels of the 5-voxel-dilated occupancy set; that is exact only if no gradient
    falls outside it.  The oracle port's gradients on a sparse-shell scene (sdf + both colour grids) must all lie inside —
    also with `neus_alpha: grad`, where EVERY MaskCache-kept sample (not only the shaded ones near the surface) scatters
    into the SDF grid through taps one voxel away.  (The GPU test test_grid_gradient_compaction_is_exact checks the kernels'
    gradients the same way in the default mode; bench.py's exchange_check compares against a dense all-reduce.)  Measured on
    this scene: a dilation of 3 voxels is the smallest that covers everything in either mode, the default of 5 leaves 2."""
    import esr_testlib as C
    from esr_nerf_b200 import synthetic as S
    from esr_nerf_b200.dist import GridGradCompactor
    from esr_nerf_b200.voxurff import VoxurfF
    from oracle import voxurf_port as P

    fx, weights = C.load_case("fine_sparse_s60_big")
    m = VoxurfF(S.fine_cfg("cpu", neus_alpha=neus_alpha), S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX,
                S.MASK_ALPHA_INIT, S.mask_density(int(fx["mask_res"]), True), float(fx["s_val"]), int(fx["num_voxels"]))
    comp = GridGradCompactor(m)
    assert 0.05 < comp.fraction < 0.6
    scene = C.oracle_scene(int(fx["num_voxels"]), int(fx["mask_res"]), True)
    scene["neus_alpha"] = neus_alpha
    assert scene["world_size"] == list(comp.shape)
    params, leaves = C.oracle_params(scene, weights)
    rays = S.make_rays(3000, 8080)
    out, inter = P.voxurff_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                            rays["em_modes"], float(fx["s_val"]))
    cot = C.cotangents(3000)
    sum((out[k] * cot[k]).sum() for k in cot).backward()
    inside = comp.mask[0, 0]
    for name in ("sdf.grid", "off_color.grid", "emo_color.grid"):
        g = leaves[name].grad[0]                                  # [C, X, Y, Z]
        touched = (g != 0).any(0)
        assert int(touched.sum()) > 1000, name
        assert not bool((touched & ~inside).any()), (name, int((touched & ~inside).sum()))
    # the alpha path alone reaches further out than the shading path: all M1 samples, not just the shaded ones
    assert inter["m1_ray"].numel() > 2 * inter["m3_ray"].numel()

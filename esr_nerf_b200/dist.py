"""Ray-sharded data parallelism for the render step: one process per GPU, each rank renders a
contiguous slice of the global ray batch against replicated parameters, gradients are summed with ONE collective
step per training step (`torch.distributed` all_reduce: NCCL over NVLink on the GPUs, gloo in the CPU tests).
No other collective is on the path.  The reference has no distributed code at all (cfg/__init__.yaml:24)."""
from __future__ import annotations

import os
from typing import Dict, Iterable

import torch


def shard_slice(n: int, rank: int, world: int) -> slice:
    """Contiguous slice of a global batch of n rays owned by `rank` (keeps BatchSampler's contiguous-slice semantics,
    utils2/utils.py:106-119; the last global ray stays on the last rank, which matters for the reference's
    last-ray-only entropy term, fine.py:378)."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return slice(lo, min(lo + per, n))


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    n = batch["rays_o"].shape[0]
    sl = shard_slice(n, rank, world)
    return {k: (v[sl] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, defer: bool = False):
    """Sum every parameter gradient over the ranks (losses are written as sums over rays / the global ray count, so
    a plain sum reproduces the single-process gradient).  A parameter some rank did not touch contributes zeros there;
    a parameter NO rank touched (e.g. envmap / brdf in a step without LTS points) ends with `grad = None` on every rank,
    so that the optimizer skips it exactly as the single-process run does (optimizer.py:74: no step count, no moment
    decay).  The collective sequence must not depend on a rank's data, so every rank reduces zeros for its own missing
    gradients and a has-gradient vector goes along (one small all-reduce(max)); the vector is read on the host AFTER
    every collective has been queued (one read at the end of the step, not a bubble in front of the exchange) and the
    gradients no rank had are reset to None.
    Returns the number of bytes reduced; with defer=True, (bytes, finish) — the caller runs finish() once it has queued
    its own collectives."""
    import torch.distributed as dist

    params = [p for p in params if p.requires_grad]
    if not params:
        return (0, lambda: None) if defer else 0
    was_none = [p.grad is None for p in params]
    dev = next((p.grad.device for p in params if p.grad is not None), params[0].device)
    has = torch.tensor([not n for n in was_none], dtype=torch.int32, device=dev)
    dist.all_reduce(has, op=dist.ReduceOp.MAX, group=group)
    nbytes = 4 * len(params)
    small = []  # the MLP weights / biases (~1 MB in ~20 tensors): one flat bucket, one collective
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        nbytes += p.grad.numel() * p.grad.element_size()
        if p.grad.numel() <= (1 << 18) and p.grad.dtype == torch.float32:
            small.append(p.grad)
        else:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def finish():
        if any(was_none):       # (identical on no rank in general — but a rank with nothing missing has nothing to reset)
            for p, n, any_rank in zip(params, was_none, has.tolist()):
                if n and not any_rank:
                    p.grad = None

    if defer:
        return nbytes, finish
    finish()
    return nbytes


class GridGradCompactor:
    """All-reduce only the grid-gradient voxels that CAN be non-zero.

    The dense grid gradients are 64 MB (sdf) + 2 x 384 MB (colour) at 256^3 — a 0.83 GB all-reduce per step, most of
    it zeros: samples only exist where MaskCache keeps them (module.py:104-114), trilinear taps reach one voxel
    further and the multi-scale SDF taps two more (voxurff.py:678-721).  The set `nonempty_mask` dilated by
    `dilate` voxels is static between grid rescales (voxurff.py:547-598), so each step gathers those voxels of the
    three gradient volumes into one [K, 13] buffer, all-reduces it, and scatters it back.  `verify=True` checks that
    nothing outside the set carried gradient (used by the tests)."""

    def __init__(self, model, dilate: int = 5):
        import torch.nn.functional as F

        self.model = model
        m = model.nonempty_mask.float()
        k = 2 * dilate + 1
        self.mask = F.max_pool3d(m, kernel_size=k, stride=1, padding=dilate) > 0
        self.idx = torch.nonzero(self.mask.reshape(-1)).reshape(-1)
        self.shape = tuple(model.sdf.grid.shape[2:])
        self.grids = [model.sdf.grid, model.off_color.grid, model.emo_color.grid]
        self._idx32 = None
        self._early = None      # (work handle, packed buffer, gradient buffers) of a colour all-reduce already in flight
        self._early_stale = False   # a later backward pass added to the colour gradients after the early exchange left
        self._overlap = False
        self.group = None

    def overlap_color_allreduce(self, enable: bool = True, group=None):
        """Start the colour-grid part of the exchange (12 of the 13 floats per voxel) from inside the backward pass, as
        soon as the encode backward has produced it (fused.COLOR_GRADS_READY_HOOK), so that it overlaps the alpha-path
        backward; allreduce() then only has the SDF grid and the MLP bucket left before it waits for it."""
        from . import fused

        self.group = group
        self._overlap = bool(enable)
        fused.COLOR_GRADS_READY_HOOK = self._on_color_grads if enable else None

    def _on_color_grads(self, bufs):
        """Every rank must issue the same collectives in the same order, so whether the colour block leaves early may
        depend only on what the CALLER does identically on every rank (gradient accumulation across calls), never on
        this rank's data: a colour volume this rank's rays did not touch goes in as zeros, and a rank whose backward
        pass never got here (no shaded sample at all) issues the same all-reduce from allreduce(), late."""
        import torch.distributed as dist

        params = self.grids[1:]
        if self._early is not None:
            # a second backward pass before allreduce() (gradient accumulation): the gradient sink adds to the buffers the
            # early exchange packed IN PLACE (same data_ptr), so what is in flight is only the first micro-batch — marked
            # here, redone in full by allreduce().  Every rank runs the same number of backward passes, so the collective
            # sequence stays rank-independent.
            self._early_stale = True
            return
        if any(p.grad is not None for p in params):
            return          # gradient accumulation across calls: the exchange happens at the end, on every rank alike
        grads = [bufs[p] if p in bufs else torch.zeros_like(p) for p in params]
        buf = self.pack([self._rows(g) for g in grads], "_cbuf")
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._early = (work, buf, grads)

    @property
    def fraction(self) -> float:
        return self.idx.numel() / self.mask.numel()

    @staticmethod
    def _rows(g: torch.Tensor) -> torch.Tensor:
        """[1,C,X,Y,Z] gradient (contiguous for C == 1, channels_last_3d otherwise) -> [XYZ, C] view in memory order"""
        v = g.permute(0, 2, 3, 4, 1)
        assert v.is_contiguous(), "grid gradient is not in the parameter's memory layout"
        return v.reshape(-1, g.shape[1])

    def outside_is_zero(self) -> bool:
        outside = ~self.mask.reshape(-1)
        return all(not bool((self._rows(p.grad)[outside] != 0).any()) for p in self.grids if p.grad is not None)

    def _grids_rows(self):
        rows = []
        for p in self.grids:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            rows.append(self._rows(p.grad))
        return rows

    def _persistent(self, name: str, n: int, device) -> torch.Tensor:
        """exchange buffers are allocated once: a buffer handed to an asynchronous NCCL call is tied to NCCL's stream by
        the caching allocator, and re-allocating it every step can miss the cache (a cudaMalloc inside the step)"""
        buf = self.__dict__.get(name)
        if buf is None or buf.numel() != n or buf.device != device:
            buf = torch.empty(n, dtype=torch.float32, device=device)
            self.__dict__[name] = buf
        return buf

    def packed_floats(self, rows) -> int:
        """floats of the packed buffer of `rows` (every volume's block padded to an even count)"""
        k = self.idx.numel()
        return sum((k * r.shape[1] + 1) // 2 * 2 for r in rows)

    def pack(self, rows, out_name: str = None, out: torch.Tensor = None) -> torch.Tensor:
        """dense gradient volumes -> one packed buffer (CUDA: one launch of esr_grad_pack, planar blocks [K][C_j];
        host tensors of the gloo tests: the same layout with torch indexing); `out`: write into this buffer (a slice of a
        larger exchange bucket) instead of allocating"""
        chans = [r.shape[1] for r in rows]
        k = self.idx.numel()
        if out is not None and not rows[0].is_cuda:
            out.copy_(self.pack(rows))
            return out
        if rows[0].is_cuda:
            import ctypes

            from ._lib import check, lib, ptr, stream_ptr

            L = lib()
            if self._idx32 is None:
                assert self.mask.numel() < (1 << 31)
                self._idx32 = self.idx.to(torch.int32)
            c_arr = (ctypes.c_int32 * len(rows))(*chans)
            v_arr = (ctypes.c_void_p * len(rows))(*[r.data_ptr() for r in rows])
            n = int(L.esr_grad_pack_floats(c_arr, len(rows), k))
            if out is not None:
                assert out.numel() == n and out.is_contiguous() and out.data_ptr() % 8 == 0
                buf = out
            else:
                buf = (self._persistent(out_name, n, rows[0].device) if out_name
                       else torch.empty(n, dtype=torch.float32, device=rows[0].device))
            check(L.esr_grad_pack(v_arr, c_arr, len(rows), ptr(self._idx32), k, ptr(buf), stream_ptr()))
            return buf
        blocks = []
        for r in rows:
            b = r[self.idx].reshape(-1)
            blocks.append(b if b.numel() % 2 == 0 else torch.cat([b, b.new_zeros(1)]))
        return torch.cat(blocks)

    def unpack(self, rows, buf: torch.Tensor) -> None:
        k = self.idx.numel()
        if rows[0].is_cuda:
            import ctypes

            from ._lib import check, lib, ptr, stream_ptr

            c_arr = (ctypes.c_int32 * len(rows))(*[r.shape[1] for r in rows])
            v_arr = (ctypes.c_void_p * len(rows))(*[r.data_ptr() for r in rows])
            check(lib().esr_grad_unpack(v_arr, c_arr, len(rows), ptr(self._idx32), k, ptr(buf), stream_ptr()))
            return
        off = 0
        for r in rows:
            n = k * r.shape[1]
            r.index_copy_(0, self.idx, buf[off:off + n].view(k, r.shape[1]))
            off += n + (n & 1)

    def _bucket_exchange(self, rows, group, name: str) -> int:
        """ONE all-reduce for everything that is not already on its way: the packed voxels of `rows`, every non-grid
        parameter gradient (the MLPs: ~1 MB in ~20 tensors) and the has-gradient flags of those parameters, side by side
        in one persistent buffer.  Around the collective: one pack / one unpack launch for the volumes, one multi-tensor
        copy in and out for the small gradients — the per-tensor version (a cat, ~20 copy-backs, three collectives) kept
        the device idle for ~0.6 ms per step behind a launch-bound host (profiles/r02 timeline_exchange).
        A parameter no rank holds a gradient for ends with grad = None (see allreduce_gradients)."""
        import torch.distributed as dist

        grid_ids = {id(p) for p in self.grids}
        others = [p for p in self.model.parameters() if id(p) not in grid_ids and p.requires_grad]
        was_none = [p.grad is None for p in others]
        for p in others:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        assert all(p.grad.dtype == torch.float32 for p in others), "parameter gradients are fp32"
        n_rows = self.packed_floats(rows) if (rows and self.idx.numel()) else 0
        sizes = [p.grad.numel() for p in others]
        n = n_rows + sum(sizes) + len(others)
        dev = rows[0].device if rows else others[0].grad.device
        buf = self._persistent(name, n + (n & 1), dev)[:n]
        if n_rows:
            self.pack(rows, out=buf[:n_rows])
        flat = list(torch.split(buf[n_rows:n_rows + sum(sizes)], sizes))
        if others:
            torch._foreach_copy_(flat, [p.grad.reshape(-1) for p in others])
            buf[n - len(others):] = torch.tensor([0.0 if w else 1.0 for w in was_none], dtype=torch.float32).to(dev, non_blocking=True)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        if n_rows:
            self.unpack(rows, buf[:n_rows])
        if others:
            torch._foreach_copy_([p.grad.reshape(-1) for p in others], flat)
            self._pending_flags = (others, was_none, buf[n - len(others):]) if any(was_none) else None
        return n * 4

    def _finish_flags(self):
        """gradients no rank had -> None again (one host read, after every collective of the step has been queued)"""
        pend, self._pending_flags = getattr(self, "_pending_flags", None), None
        if pend is not None:
            others, was_none, flags = pend
            for p, w, f in zip(others, was_none, flags.tolist()):
                if w and f == 0.0:
                    p.grad = None

    def allreduce(self, group=None, verify: bool = False) -> int:
        color = self.grids[1:]
        if self._early is None and self._overlap and all(p.grad is None for p in color):
            self._on_color_grads({})     # this rank's backward never reached the hook: same collective, zeros, now
        early, self._early = self._early, None
        stale, self._early_stale = self._early_stale, False
        if early is not None:
            for p, b in zip(color, early[2]):
                if p.grad is None:       # a volume this rank did not touch: the zeros that went into the exchange
                    p.grad = b
        rows = self._grids_rows()
        if verify:
            assert self.outside_is_zero(), "gradient outside the dilated occupancy set"
        if early is not None and not stale and all(p.grad is not None and p.grad.data_ptr() == b.data_ptr()
                                                   for p, b in zip(self.grids[1:], early[2])):
            # the colour volumes are already on their way (started inside the backward pass): SDF grid + MLPs now, then join
            work, cbuf, _ = early
            nbytes = self._bucket_exchange(rows[:1], group, "_sbuf")
            work.wait()
            self.unpack(rows[1:], cbuf)
            self._finish_flags()
            return nbytes + cbuf.numel() * 4
        if early is not None:
            early[0].wait()     # the buffers it reduced are not the final gradients (accumulation): redo the exchange
        nbytes = self._bucket_exchange(rows, group, "_abuf")
        self._finish_flags()
        return nbytes


class TouchedBlockCompactor(GridGradCompactor):
    """Exact compacted exchange without a static voxel set (the LTS / PDRA stage): the eps-jittered samples
    and the secondary rays of `ESRNeRF.forward_training` (esrnerf.py:576-652, 807-830) touch voxels that no occupancy
    mask bounds, so the set is found per step, in two levels:

      1. every rank marks the blocks (8^3 voxels; per axis the largest divisor of the grid size <= 8) in which ANY of
         its grid-gradient volumes holds a non-zero (`esr_grad_block_flags`: one launch, each volume read once);
      2. the ranks OR the flag maps (one 128 KB all-reduce(max) at 256^3) — every rank now holds the same union;
      3. the voxels of the union's blocks are packed / all-reduced / unpacked exactly like the static set
         (`esr_grad_pack` / `esr_grad_unpack`): a voxel outside the union is zero on every rank, so the result equals
         the dense all-reduce by construction, bit for bit on two ranks.

    One host read per step (the number of touched blocks, identical on every rank after step 2).  Exchange buffers grow
    monotonically and are re-used (no allocation inside a steady-state step)."""

    def __init__(self, model, grids=None, block: int = 8):
        self.model = model
        if grids is None:
            grids = [g.grid for g in (getattr(model, n, None) for n in ("sdf", "off_color", "emo_color", "brdf"))
                     if g is not None]
        self.grids = list(grids)
        self.shape = tuple(self.grids[0].shape[2:])
        assert all(tuple(g.shape[2:]) == self.shape for g in self.grids), "grids of one scene share their resolution"
        assert self.shape[0] * self.shape[1] * self.shape[2] < (1 << 31)
        self.edge = tuple(max(e for e in range(1, block + 1) if n % e == 0) for n in self.shape)
        self.blocks = tuple(n // e for n, e in zip(self.shape, self.edge))
        self.mask = None
        self.idx = torch.zeros(0, dtype=torch.int64)
        self._idx32 = None
        self._early = None
        self._overlap = False
        self.group = None
        self._tables = {}
        self.last_fraction = 0.0

    def overlap_color_allreduce(self, enable: bool = True, group=None):
        raise NotImplementedError("the touched set is only known once every gradient of the step is final")

    @property
    def fraction(self) -> float:
        return self.last_fraction

    def outside_is_zero(self) -> bool:
        keep = torch.zeros(self.shape[0] * self.shape[1] * self.shape[2], dtype=torch.bool, device=self.idx.device)
        keep[self.idx] = True
        return all(not bool((self._rows(p.grad)[~keep] != 0).any()) for p in self.grids if p.grad is not None)

    def _persistent(self, name: str, n: int, device) -> torch.Tensor:
        buf = self.__dict__.get(name)
        if buf is None or buf.numel() < n or buf.device != device:
            buf = torch.empty(max(n, 1) * 5 // 4, dtype=torch.float32, device=device)     # head-room: K drifts step to step
            self.__dict__[name] = buf
        return buf[:n]

    def block_flags(self, rows) -> torch.Tensor:
        """int32 [Bx*By*Bz]: 1 where a block holds a non-zero in any of `rows` ([XYZ, C_j] views in memory order)"""
        (bx, by, bz), (ex, ey, ez) = self.blocks, self.edge
        if rows[0].is_cuda and not os.environ.get("ESR_BLOCK_FLAGS_TORCH"):     # one launch of esr_grad_block_flags
            import ctypes

            from ._lib import check, lib, ptr, stream_ptr

            c_arr = (ctypes.c_int32 * len(rows))(*[r.shape[1] for r in rows])
            v_arr = (ctypes.c_void_p * len(rows))(*[r.data_ptr() for r in rows])
            flags = self.__dict__.get("_flags")
            if flags is None or flags.device != rows[0].device:
                flags = self._flags = torch.empty(bx * by * bz, dtype=torch.int32, device=rows[0].device)
            check(lib().esr_grad_block_flags(v_arr, c_arr, len(rows), *self.shape, ex, ey, ez, ptr(flags), stream_ptr()))
            return flags
        cnt = None      # host tensors of the gloo tests (and the A/B switch): the same map with torch reductions
        for r in rows:
            c = torch.count_nonzero(r.view(bx, ex, by, ey, bz, ez * r.shape[1]), dim=(1, 3, 5))
            cnt = c if cnt is None else cnt + c
        return (cnt > 0).to(torch.int32).reshape(-1)

    def _block_tables(self, device):
        t = self._tables.get(device)
        if t is None:
            (X, Y, Z), (bx, by, bz), (ex, ey, ez) = self.shape, self.blocks, self.edge
            ar = lambda n: torch.arange(n, device=device, dtype=torch.int64)
            base = ((ar(bx) * ex)[:, None, None] * Y + (ar(by) * ey)[None, :, None]) * Z + (ar(bz) * ez)[None, None, :]
            within = (ar(ex)[:, None, None] * Y + ar(ey)[None, :, None]) * Z + ar(ez)[None, None, :]
            t = self._tables[device] = (base.reshape(-1), within.reshape(-1))
        return t

    def select(self, flags: torch.Tensor) -> int:
        """voxel list (self.idx) of the flagged blocks; returns the number of blocks"""
        base, within = self._block_tables(flags.device)
        blocks = torch.nonzero(flags).reshape(-1)              # the step's one host read
        self.idx = (base[blocks][:, None] + within[None, :]).reshape(-1)
        self._idx32 = self.idx.to(torch.int32) if self.idx.is_cuda else None
        self.last_fraction = blocks.numel() / max(flags.numel(), 1)
        return int(blocks.numel())

    def allreduce(self, group=None, verify: bool = False) -> int:
        import torch.distributed as dist

        rows = self._grids_rows()
        flags = self.block_flags(rows)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
        self.select(flags)
        if verify:
            assert self.outside_is_zero(), "non-zero gradient outside the union of touched blocks"
        nbytes = self._bucket_exchange(rows, group, "_abuf") + flags.numel() * 4
        self._finish_flags()
        return nbytes


def gather_maps(maps: Dict[str, torch.Tensor], n_total: int, rank: int, world: int, group=None, dst: int = 0):
    """Per-ray output maps of a ray-sharded render (each rank holds the rows of `shard_slice(n_total, rank, world)`)
    -> the full [n_total, ...] maps on rank `dst` (None elsewhere): SURVEY.md §8e, BASELINE.json configs[3] (1600x1200
    full-image inference, 12 maps).  The maps are packed side by side into ONE [rows, sum C] float buffer, padded to the
    common slice length (only the last rank's slice can be shorter), and gathered with one collective."""
    import torch.distributed as dist

    keys = sorted(maps)
    per = (n_total + world - 1) // world
    widths = [int(torch.Size(maps[k].shape[1:]).numel()) for k in keys]      # (explicit: an empty slice has no rows to infer from)
    cols = [maps[k].reshape(maps[k].shape[0], wd) for k, wd in zip(keys, widths)]
    assert all(c.dtype == torch.float32 for c in cols), "render maps are fp32"
    packed = torch.zeros(per, sum(widths), dtype=torch.float32, device=cols[0].device)
    packed[:cols[0].shape[0]] = torch.cat(cols, dim=1)
    if world == 1:
        full = packed
    else:
        out = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
        dist.gather(packed, out, dst=dst, group=group)
        if rank != dst:
            return None
        full = torch.cat(out, dim=0)
    full = full[:n_total]
    res, off = {}, 0
    for k, wd in zip(keys, widths):
        res[k] = full[:, off:off + wd].reshape(n_total, *maps[k].shape[1:])
        off += wd
    return res


@torch.no_grad()
def render_image_sharded(model, rays: Dict[str, torch.Tensor], rank: int, world: int, chunk: int = 1 << 18, group=None,
                         **forward_kwargs):
    """One full image through `model.forward_evaluate` on `world` GPUs: contiguous pixel ranges per rank, `chunk` rays
    per call, all maps gathered on rank 0 (None on the other ranks).  `rays`: the image's [n, 3] rays_o / rays_d /
    viewdirs (every rank holds them, or at least its own slice's rows at the right positions)."""
    n = rays["rays_o"].shape[0]
    sl = shard_slice(n, rank, world)
    outs = []
    for lo in range(sl.start, sl.stop, chunk):
        hi = min(lo + chunk, sl.stop)
        outs.append(model(rays_o=rays["rays_o"][lo:hi], rays_d=rays["rays_d"][lo:hi], viewdirs=rays["viewdirs"][lo:hi],
                          **forward_kwargs))
    if not outs:        # more ranks than rays: an empty slice still takes part in the collective
        proto = model(rays_o=rays["rays_o"][:0], rays_d=rays["rays_d"][:0], viewdirs=rays["viewdirs"][:0], **forward_kwargs)
        outs = [proto]
    local = {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}
    return gather_maps(local, n, rank, world, group)

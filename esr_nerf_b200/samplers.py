"""The ray feed of the render step: drop-ins for the reference's `BatchSampler` and
`RayGroupManager` (utils2/utils.py:41-312) with the same constructor arguments, the same `shuffle / filter / sample`
behaviour, the same checkpointed state (`batch_st`, `data_idxs`, ...) and the same random draws (`torch.randperm` on the
same device), so a run can resume a reference checkpoint and vice versa (fine.py:221-228, 484-485).

What changes is what moves.  The reference keeps every key PHYSICALLY in sampling order: each `shuffle()` and `filter()`
re-materialises the whole ray set (`data[k] = data[k][b_ids].contiguous()`, ~100 B per ray for six keys, 10^7-10^8 rays:
gigabytes per epoch wrap, and a transient second copy of the set).  With the set resident in HBM (`data_preload: gpu`)
this module never moves a ray: it keeps the data where it was loaded and permutes only the int64 index (`data_idxs` —
the very tensor the reference checkpoints; the invariant `data[k] == loaded[k][data_idxs]` holds in the reference after
every operation), and `sample()` gathers the batch's rows through it: batch_size random rows per key per step (~6 MB at
2^16 rays), nothing per shuffle.  `data` / `uncert_data` / `cert_data` remain readable as mappings (pdra.py:888-889):
a key is gathered when it is asked for.

`rank / world` (no counterpart in the reference, which is single-process: cfg/__init__.yaml:24): every rank holds the
same set and — seeded alike — draws the same permutations; `sample()` returns only the rank's contiguous slice of the
global batch (`dist.shard_slice`, last global ray on the last rank), so no rank gathers rows it will not render.

`data_preload: cpu` keeps the reference's pinned-host behaviour (the permuted copy is what makes its per-step slice a
contiguous pinned block that can be copied asynchronously)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .dist import shard_slice
from .modules import cfg_get


def _check_preload(mode: str) -> bool:
    assert "cpu" == mode or "gpu" in mode or "cuda" in mode      # utils2/utils.py:54-58
    return mode == "cpu"


class _Gathered:
    """read-only mapping view `loaded[k][idx]` (what the reference holds as a physical copy)"""

    def __init__(self, loaded: Dict[str, torch.Tensor], keys: List[str], idx_of, owned: Optional[set] = None):
        self._loaded, self._keys, self._idx_of = loaded, keys, idx_of
        self._owned = set() if owned is None else owned      # keys whose load-order tensor this sampler allocated
        self._n = len(loaded[keys[0]])

    def __getitem__(self, k):
        return self._loaded[k][self._idx_of()]

    def __setitem__(self, k, v):
        """`group_data[k] = v` with v in the group's current order (pdra.py:1023-1036 attaches per-ray edit attributes
        to both groups this way): scattered into the load-order tensor of key k, created on first use"""
        idx = self._idx_of()
        v = v.to(idx.device)
        full = self._loaded.get(k)
        if full is None or full.dtype != v.dtype or full.shape[1:] != v.shape[1:]:
            full = torch.zeros((self._n,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        elif k not in self._owned:
            full = full.clone()                      # the loaded tensor may be the caller's own
        self._owned.add(k)
        full[idx] = v
        self._loaded[k] = full

    def keys(self):
        return list(self._keys)

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def __contains__(self, k):
        return k in self._keys


def _pin(t: torch.Tensor) -> torch.Tensor:
    return torch._pin_memory(t.contiguous()) if torch.cuda.is_available() else t.contiguous()


class BatchSampler:
    """utils2/utils.py:41-119"""

    def __init__(self, cfg, data: Dict[str, torch.Tensor], keys: List[str], batch_size: int, batch_st: int = 0,
                 data_idxs: Optional[torch.Tensor] = None, rank: int = 0, world: int = 1):
        self.cfg = cfg
        self.device = cfg_get(cfg, "system.device")
        self.data_preload_to_cpu = _check_preload(cfg_get(cfg, "system.data_preload"))
        self.keys = keys
        self.batch_size = batch_size
        self.batch_st = batch_st
        self.rank, self.world = rank, world
        self.data_idxs = torch.arange(len(data[keys[0]])) if data_idxs is None else data_idxs
        if self.data_preload_to_cpu:
            self.data_idxs = _pin(self.data_idxs.cpu())
            self.data = data
            for k in keys:
                data[k] = _pin(data[k][self.data_idxs])
        else:
            self.data_idxs = self.data_idxs.to(self.device).contiguous()
            self._loaded = {k: data[k].to(self.device) for k in keys}          # stays in load order for good
            self.data = _Gathered(self._loaded, keys, lambda: self.data_idxs)

    @property
    def data_num(self) -> int:
        return len(self.data_idxs)

    def shuffle(self):
        if self.data_preload_to_cpu:
            b_ids = torch.randperm(self.data_num, pin_memory=torch.cuda.is_available())
            self.data_idxs = _pin(self.data_idxs[b_ids])
            for k in self.keys:
                self.data[k] = _pin(self.data[k][b_ids])
        else:
            b_ids = torch.randperm(self.data_num, device=self.device)
            self.data_idxs = self.data_idxs[b_ids].contiguous()
        self.batch_st = 0

    def filter(self, mask: torch.Tensor):
        if self.data_preload_to_cpu:
            mask = mask.cpu().contiguous()
            for k in self.keys:
                self.data[k] = _pin(self.data[k][mask])
            self.data_idxs = _pin(self.data_idxs[mask])
        else:
            self.data_idxs = self.data_idxs[mask.to(self.device)].contiguous()

    def sample(self) -> Dict[str, torch.Tensor]:
        b_en = self.batch_st + self.batch_size
        if b_en > self.data_num:
            self.shuffle()
            b_en = self.batch_size
        b_st = self.batch_st
        self.batch_st = b_en
        sl = shard_slice(b_en - b_st, self.rank, self.world)
        lo, hi = b_st + sl.start, b_st + sl.stop
        if self.data_preload_to_cpu:
            return {k: self.data[k][lo:hi].to(self.device, non_blocking=True) for k in self.keys}
        rows = self.data_idxs[lo:hi]
        return {k: self._loaded[k][rows] for k in self.keys}


class RayGroupManager:
    """utils2/utils.py:122-312: the uncertain / certain ray groups of the LTS / PDRA stages.  `filter(mask)` moves the
    uncertain rays with mask False to the END of the certain group, in their current order."""

    def __init__(self, cfg, data: Dict[str, torch.Tensor], keys: List[str], uncert_batch_size: int, cert_batch_size: int,
                 uncert_batch_st: int = 0, cert_batch_st: int = 0, uncert_data_idxs: Optional[torch.Tensor] = None,
                 cert_data_idxs: Optional[torch.Tensor] = None, rank: int = 0, world: int = 1):
        self.cfg = cfg
        self.device = cfg_get(cfg, "system.device")
        self.data_preload_to_cpu = _check_preload(cfg_get(cfg, "system.data_preload"))
        self.keys = keys
        self.uncert_batch_size, self.cert_batch_size = uncert_batch_size, cert_batch_size
        self.uncert_batch_st, self.cert_batch_st = uncert_batch_st, cert_batch_st
        self.rank, self.world = rank, world
        self.uncert_data_idxs = torch.arange(len(data[keys[0]])) if uncert_data_idxs is None else uncert_data_idxs
        self.cert_data_idxs = torch.arange(0) if cert_data_idxs is None else cert_data_idxs
        if self.data_preload_to_cpu:
            self.uncert_data_idxs = _pin(self.uncert_data_idxs.cpu())
            self.cert_data_idxs = _pin(self.cert_data_idxs.cpu())
            self.uncert_data = {k: _pin(data[k][self.uncert_data_idxs]) for k in keys}
            self.cert_data = {k: _pin(data[k][self.cert_data_idxs]) for k in keys}
        else:
            self.uncert_data_idxs = self.uncert_data_idxs.to(self.device).contiguous()
            self.cert_data_idxs = self.cert_data_idxs.to(self.device).contiguous()
            self._loaded = {k: data[k].to(self.device) for k in keys}
            owned = set()
            self.uncert_data = _Gathered(self._loaded, keys, lambda: self.uncert_data_idxs, owned)
            self.cert_data = _Gathered(self._loaded, keys, lambda: self.cert_data_idxs, owned)

    @property
    def uncert_data_num(self) -> int:
        return len(self.uncert_data_idxs)

    @property
    def cert_data_num(self) -> int:
        return len(self.cert_data_idxs)

    def shuffle(self):
        self.shuffle_uncert()
        self.shuffle_cert()

    def _shuffle(self, which: str):
        n = len(getattr(self, which + "_data_idxs"))
        if self.data_preload_to_cpu:
            b_ids = torch.randperm(n, pin_memory=torch.cuda.is_available())
            setattr(self, which + "_data_idxs", _pin(getattr(self, which + "_data_idxs")[b_ids]))
            d = getattr(self, which + "_data")
            for k in self.keys:
                d[k] = _pin(d[k][b_ids])
        else:
            b_ids = torch.randperm(n, device=self.device)
            setattr(self, which + "_data_idxs", getattr(self, which + "_data_idxs")[b_ids].contiguous())
        setattr(self, which + "_batch_st", 0)

    def shuffle_uncert(self):
        self._shuffle("uncert")

    def shuffle_cert(self):
        self._shuffle("cert")

    def filter(self, mask: torch.Tensor):
        if self.data_preload_to_cpu:
            mask = mask.cpu().contiguous()
            for k in self.keys:
                self.cert_data[k] = _pin(torch.concat([self.cert_data[k], self.uncert_data[k][~mask]], dim=0))
                self.uncert_data[k] = _pin(self.uncert_data[k][mask])
            self.cert_data_idxs = _pin(torch.concat([self.cert_data_idxs, self.uncert_data_idxs[~mask]], dim=0))
            self.uncert_data_idxs = _pin(self.uncert_data_idxs[mask])
        else:
            mask = mask.to(self.device).contiguous()
            self.cert_data_idxs = torch.concat([self.cert_data_idxs, self.uncert_data_idxs[~mask]], dim=0).contiguous()
            self.uncert_data_idxs = self.uncert_data_idxs[mask].contiguous()

    def sample(self) -> Dict[str, torch.Tensor]:
        uncert_b_en = self.uncert_batch_st + self.uncert_batch_size
        cert_b_en = self.cert_batch_st + self.cert_batch_size
        if uncert_b_en > self.uncert_data_num:
            self.shuffle_uncert()
            uncert_b_en = min(len(self.uncert_data_idxs), self.uncert_batch_size)
        if cert_b_en > self.cert_data_num:
            self.shuffle_cert()
            cert_b_en = min(len(self.cert_data_idxs), self.cert_batch_size)
        uncert_b_st, cert_b_st = self.uncert_batch_st, self.cert_batch_st
        self.uncert_batch_st, self.cert_batch_st = uncert_b_en, cert_b_en
        uncert_bs, cert_bs = uncert_b_en - uncert_b_st, cert_b_en - cert_b_st
        if self.data_preload_to_cpu:
            batch = {k: torch.concat([self.uncert_data[k][uncert_b_st:uncert_b_en],
                                      self.cert_data[k][cert_b_st:cert_b_en]], dim=0) for k in self.keys}
        else:
            rows = torch.concat([self.uncert_data_idxs[uncert_b_st:uncert_b_en], self.cert_data_idxs[cert_b_st:cert_b_en]])
            batch = None
        masks = torch.ones(uncert_bs + cert_bs, dtype=torch.bool, device=self.device)
        masks[-cert_bs:] = False      # utils2/utils.py:302 as written: with no certain ray ([-0:]) EVERY mask is False
        sl = shard_slice(uncert_bs + cert_bs, self.rank, self.world)
        if batch is None:
            batch = {k: self._loaded[k][rows[sl]] for k in self.keys}
        else:
            batch = {k: v[sl].to(self.device, non_blocking=True) for k, v in batch.items()}
        batch["uncert_masks"] = masks[sl]
        return batch

    def print_stats(self):
        nuncert, ncert = self.uncert_data_num, self.cert_data_num
        print(f"uncertain: {nuncert}\t certain: {ncert}\t uncertain/all: {nuncert / (nuncert + ncert) * 100}")


@torch.no_grad()
def update_ray_groups(renderer, sampler: RayGroupManager, k_val: float, eval_uncert_bs: int, rank: int = 0, world: int = 1,
                      group=None) -> torch.Tensor:
    """pdra.py:882-932: the emission sweep over ALL uncertain training rays that re-partitions the
    ray groups every `update_step` steps — `renderer.eval_emit` on chunks of `eval_uncert_bs` rays, a ray stays
    uncertain while max_c emission > k_val, the rest join the certain group (`sampler.filter`).

    The inference path at data-set scale (10^7-10^8 rays per sweep), so it is sharded like the render step: each rank
    sweeps a contiguous slice of the uncertain set (rows gathered chunk by chunk through the sampler's index — the set
    itself is never copied), writes its per-ray maxima into a zero vector, and ONE all-reduce(sum) of that vector (4 B
    per ray) gives every rank the same mask, hence the same groups.  Rays are independent in every kernel of the path,
    so the result does not depend on where the chunk or rank boundaries fall.  Returns the mask it applied."""
    device = sampler.device
    rows_all = sampler.uncert_data_idxs
    n = len(rows_all)
    peak = torch.zeros(n, dtype=torch.float32, device=device)
    was_training = getattr(renderer, "training", False)
    renderer.eval()
    sl = shard_slice(n, rank, world)
    for lo in range(sl.start, sl.stop, eval_uncert_bs):
        hi = min(lo + eval_uncert_bs, sl.stop)
        if sampler.data_preload_to_cpu:
            batch = {k: sampler.uncert_data[k][lo:hi].to(device) for k in ("rays_o", "rays_d", "viewdirs")}
        else:
            rows = rows_all[lo:hi]
            batch = {k: sampler._loaded[k][rows] for k in ("rays_o", "rays_d", "viewdirs")}
        peak[lo:hi] = torch.max(renderer.eval_emit(**batch), dim=-1)[0]
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(peak, op=dist.ReduceOp.SUM, group=group)      # slices are disjoint: the sum is a concatenation
    mask = peak > k_val
    sampler.filter(mask)
    if was_training:
        renderer.train()
    return mask


def dilate_masks(em_masks: torch.Tensor, ks: int) -> torch.Tensor:
    """cv2.dilate(masks, np.ones((ks, ks)), iterations=1) of pdra.py:946-961 on [n, h, w] masks, on the device: a
    ks x ks maximum whose anchor is the kernel centre (ks // 2, ks // 2) — for the shipped even size (10,
    cfg/app/pdra.yaml:132) the window reaches ks // 2 pixels towards smaller and ks - 1 - ks // 2 towards larger
    indices — and pixels outside the image do not take part."""
    import torch.nn.functional as F

    a = ks // 2
    x = F.pad(em_masks.float()[:, None], (a, ks - 1 - a, a, ks - 1 - a), value=float("-inf"))
    return F.max_pool2d(x, kernel_size=ks, stride=1)[:, 0]


@torch.no_grad()
def filter_edit_rays(renderer, sampler: RayGroupManager, test_data: Dict[str, torch.Tensor], image_size, focal_length: float,
                     mask_dilation_ks: int, eval_bs: int, rank: int = 0, world: int = 1, group=None) -> RayGroupManager:
    """pdra.py:934-1038: before the finetune stage every uncertain training ray is traced to its
    emissive surface point (`renderer.eval_esp`), the point is projected into the edit view, and rays that land inside
    a (dilated) edit mask receive that mask's edit (`em_modes / em_colors / em_intensities`, LightDict of
    utils2/utils.py:32-38); the others leave the uncertain group.  Same arithmetic as the reference, including its
    bounds test (both image coordinates against both h - 1 and w - 1, :995) and "later masks overwrite earlier ones".

    Sharded like `update_ray_groups`: each rank sweeps a contiguous slice of the uncertain set, one all-reduce(sum) of
    the per-ray results [n, 5] (hit, mode - 1, colour x 2, intensity; zero outside the rank's slice) makes the groups
    identical on every rank."""
    import torch.nn.functional as F

    device = sampler.device
    w, h = image_size
    f = focal_length
    w2c = torch.inverse(test_data["poses"]).to(device)
    K = torch.tensor([[-f, 0.0, w / 2.0 - 0.5], [0.0, f, h / 2.0 - 0.5], [0.0, 0.0, 1.0]]).to(device, dtype=torch.float32)
    em_masks = dilate_masks(test_data["em_masks"].view(-1, h, w).to(device), mask_dilation_ks).view(-1, 1, h, w)
    n_cond = len(em_masks)

    rows_all = sampler.uncert_data_idxs
    n = len(rows_all)
    res = torch.zeros(n, 5, dtype=torch.float32, device=device)      # hit, mode - 1, colour (2), intensity
    was_training = getattr(renderer, "training", False)
    renderer.eval()
    sl = shard_slice(n, rank, world)
    for lo in range(sl.start, sl.stop, eval_bs):
        hi = min(lo + eval_bs, sl.stop)
        if sampler.data_preload_to_cpu:
            batch = {k: sampler.uncert_data[k][lo:hi].to(device) for k in ("rays_o", "rays_d", "viewdirs")}
        else:
            rows = rows_all[lo:hi]
            batch = {k: sampler._loaded[k][rows] for k in ("rays_o", "rays_d", "viewdirs")}
        esp = renderer.eval_esp(**batch)
        esp = torch.concat([esp, torch.ones_like(esp[..., :1])], dim=-1).T
        xyz = w2c @ esp
        cam_coord = xyz[:3] / xyz[-1:]
        xyz = K @ cam_coord
        img_coord = (xyz[:2] / xyz[-1:]).T
        out_bound = (img_coord < 0) | (img_coord > (h - 1)) | (img_coord > (w - 1))
        in_bound = (out_bound[..., 0] | out_bound[..., 1]).bitwise_not()
        img_coord = img_coord[in_bound]
        img_coord[..., 0] = img_coord[..., 0] / (w - 1) * 2 - 1
        img_coord[..., 1] = img_coord[..., 1] / (h - 1) * 2 - 1
        img_coord = img_coord.view(1, 1, -1, 2).repeat(n_cond, 1, 1, 1)
        m = (F.grid_sample(em_masks, img_coord, align_corners=True, mode="bilinear") > 0).view(n_cond, -1)
        where = torch.arange(lo, hi, device=device)[in_bound]
        res[where, 0] = (torch.sum(m, dim=0) > 0).float()
        for i in range(n_cond):
            hit = where[m[i]]
            mode = int(test_data["em_modes"][i])
            res[hit, 1] = float(mode - 1)
            if mode == 0:                                   # LightDict["off"]
                res[hit, 4] = 0.0
            if mode in (2, 4):                              # i_change, ic_change
                res[hit, 4] = test_data["em_intensities"][i].to(device, dtype=torch.float32)
            if mode in (3, 4):                              # c_change, ic_change
                res[hit, 2:4] = test_data["em_colors"][i][:2].to(device, dtype=torch.float32)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(res, op=dist.ReduceOp.SUM, group=group)
    nc = sampler.cert_data_num
    sampler.uncert_data["em_modes"] = res[:, 1].round().long() + 1
    sampler.uncert_data["em_colors"] = res[:, 2:4].contiguous()
    sampler.uncert_data["em_intensities"] = res[:, 4].contiguous()
    sampler.cert_data["em_modes"] = torch.zeros(nc, dtype=torch.long, device=device)
    sampler.cert_data["em_colors"] = torch.zeros(nc, 2, dtype=torch.float32, device=device)
    sampler.cert_data["em_intensities"] = torch.zeros(nc, dtype=torch.float32, device=device)
    sampler.keys.extend(["em_colors", "em_intensities"])
    sampler.filter(res[:, 0] > 0)
    if was_training:
        renderer.train()
    return sampler

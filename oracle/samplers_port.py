"""TEST INFRASTRUCTURE — restatement of the reference's ray feed (utils2/utils.py:41-312) for the parity tests.

The reference keeps every key of the ray set PHYSICALLY in sampling order and re-materialises it on every shuffle /
filter; the product (esr_nerf_b200/samplers.py) permutes an index instead.  This port restates the reference's
behaviour in the simplest possible form — one `_Pool` (index vector + physically ordered copies) per ray group — so
that the product can be compared with it ON THE GPU, where the reference itself cannot be imported: same
`torch.randperm(n, device=...)` draws in the same order, same batches, same checkpointed state.
Pinned against the reference's own classes on the CPU by tests/test_samplers_cpu.py::test_port_matches_reference.
Only tests may import this module."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class _Pool:
    """one ray group held the reference's way: `idx` (utils.py:65-68) and `rows[k] == loaded[k][idx]`, contiguous"""

    def __init__(self, data: Dict[str, torch.Tensor], keys: List[str], idx: torch.Tensor, device):
        self.keys, self.device = keys, device
        self.idx = idx.to(device).contiguous()
        self.rows = {k: data[k].to(device)[self.idx].contiguous() for k in keys}     # utils.py:76-79 / 176-182

    def __len__(self):
        return len(self.idx)

    def take(self, sel: torch.Tensor):
        """re-order / subset by a permutation or a boolean mask (utils.py:88-92, 102-106)"""
        self.idx = self.idx[sel].contiguous()
        for k in self.keys:
            self.rows[k] = self.rows[k][sel].contiguous()

    def permute(self):
        self.take(torch.randperm(len(self), device=self.device))                    # utils.py:88, 203, 219

    def append(self, other: "_Pool", sel: torch.Tensor):
        """utils.py:252-260: the selected rows of `other` join the END of this pool, in their current order"""
        self.idx = torch.concat([self.idx, other.idx[sel]], dim=0).contiguous()
        for k in self.keys:
            self.rows[k] = torch.concat([self.rows[k], other.rows[k][sel]], dim=0).contiguous()


class BatchSamplerPort:
    """utils2/utils.py:41-119, device-resident branch (`data_preload` containing gpu / cuda)"""

    def __init__(self, device, data, keys, batch_size, batch_st=0, data_idxs: Optional[torch.Tensor] = None):
        n = len(data[keys[0]])
        self.pool = _Pool(data, keys, torch.arange(n) if data_idxs is None else data_idxs, device)
        self.keys, self.batch_size, self.batch_st = keys, batch_size, batch_st

    data_idxs = property(lambda self: self.pool.idx)
    data = property(lambda self: self.pool.rows)
    data_num = property(lambda self: len(self.pool))

    def shuffle(self):
        self.pool.permute()
        self.batch_st = 0                                                            # utils.py:93

    def filter(self, mask):
        self.pool.take(mask.to(self.pool.device))

    def sample(self):
        end = self.batch_st + self.batch_size
        if end > self.data_num:                                                      # utils.py:110-112: wrap = re-shuffle
            self.shuffle()
            end = self.batch_size
        start, self.batch_st = self.batch_st, end
        return {k: self.pool.rows[k][start:end] for k in self.keys}


class RayGroupManagerPort:
    """utils2/utils.py:122-312, device-resident branch"""

    def __init__(self, device, data, keys, uncert_batch_size, cert_batch_size, uncert_batch_st=0, cert_batch_st=0,
                 uncert_data_idxs=None, cert_data_idxs=None):
        n = len(data[keys[0]])
        self.device, self.keys = device, keys
        self.unc = _Pool(data, keys, torch.arange(n) if uncert_data_idxs is None else uncert_data_idxs, device)
        self.cer = _Pool(data, keys, torch.arange(0) if cert_data_idxs is None else cert_data_idxs, device)
        self.bs = (uncert_batch_size, cert_batch_size)
        self.st = [uncert_batch_st, cert_batch_st]

    uncert_data_idxs = property(lambda self: self.unc.idx)
    cert_data_idxs = property(lambda self: self.cer.idx)
    uncert_data = property(lambda self: self.unc.rows)
    cert_data = property(lambda self: self.cer.rows)
    uncert_data_num = property(lambda self: len(self.unc))
    cert_data_num = property(lambda self: len(self.cer))
    uncert_batch_st = property(lambda self: self.st[0])
    cert_batch_st = property(lambda self: self.st[1])

    def shuffle(self):                                                               # utils.py:194-196: uncertain first
        for i, pool in enumerate((self.unc, self.cer)):
            pool.permute()
            self.st[i] = 0

    def filter(self, mask):
        mask = mask.to(self.device)
        self.cer.append(self.unc, ~mask)                                             # utils.py:252-260, before the uncertain
        self.unc.take(mask)                                                          # group itself shrinks

    def sample(self):
        span = []
        for i, pool in enumerate((self.unc, self.cer)):                              # utils.py:265-277
            end = self.st[i] + self.bs[i]
            if end > len(pool):
                pool.permute()
                self.st[i] = 0
                end = min(len(pool), self.bs[i])
            span.append((self.st[i], end))
            self.st[i] = end
        (u0, u1), (c0, c1) = span
        batch = {k: torch.concat([self.unc.rows[k][u0:u1], self.cer.rows[k][c0:c1]], dim=0) for k in self.keys}
        masks = torch.ones((u1 - u0) + (c1 - c0), dtype=torch.bool, device=self.device)
        masks[-(c1 - c0):] = False                                                   # utils.py:302 as written ([-0:] = all)
        batch["uncert_masks"] = masks
        return batch

"""Drop-in for ``app/utils/optimizer.py``: ``create_optimizer_or_freeze_model`` (name -> parameter-group mapping and
freezing, :11-60), ``Adam`` with the per-voxel learning-rate extension (:63-228) and ``CosineLR`` (:231-275) — with
the update itself as ONE fused sm_100a kernel per parameter tensor (``esr_adam_step``) instead of ~10 elementwise
passes over 0.83-1.2 GB of grids.  SURVEY.md §8f row 3: a caller of the render path, built to the same parity bar.

    from esr_nerf_b200.optimizer import create_optimizer_or_freeze_model, CosineLR    # fine.py:191, 331
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, stream_ptr


def create_optimizer_or_freeze_model(model: nn.Module, **lrates: float) -> "Adam":
    """optimizer.py:11-60: one parameter group per named attribute with lr > 0; lr <= 0 freezes it; parameters not
    named at all are left out of the optimizer (they keep requires_grad)."""
    groups = []
    for name, lr in lrates.items():
        if not hasattr(model, name):
            continue
        target = getattr(model, name)
        if target is None:
            continue
        params = list(target.parameters()) if isinstance(target, nn.Module) else [target]
        if lr > 0:
            groups.append({"params": params, "lr": lr, "name": name})
        else:
            for p in params:
                p.requires_grad = False
    return Adam(groups, betas=(0.9, 0.99))


class Adam(torch.optim.Optimizer):
    """optimizer.py:63-183.  State (`step`, `exp_avg`, `exp_avg_sq`) and `state_dict` layout equal the reference's, so
    optimizer checkpoints move both ways."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad:
            raise NotImplementedError("amsgrad is never enabled by the reference (optimizer.py:60); not instantiated")
        self.per_lr = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self.name2pg = {pg["name"]: pg for pg in self.param_groups if "name" in pg}

    def set_pervoxel_lr(self, count: torch.Tensor) -> None:
        """optimizer.py:107-109: learning rate of the first parameter (the density volume) scaled per voxel by
        view count / max view count"""
        assert self.param_groups[0]["params"][0].shape == count.shape
        self.per_lr = (count.float() / count.max()).contiguous()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                per_lr = self.per_lr if (self.per_lr is not None and p.shape == self.per_lr.shape) else None
                grad = p.grad
                # one dense pass over the parameter's storage: gradient and moments must share its memory layout
                # (channels-last colour grids included: zeros_like(preserve_format) and the render backward keep it)
                if grad.stride() != p.stride():
                    grad = torch.empty_like(p, memory_format=torch.preserve_format).copy_(grad)
                assert _dense(p) and _dense(state["exp_avg"]) and state["exp_avg"].stride() == p.stride()
                with torch.cuda.device(p.device):
                    check(L.esr_adam_step(ptr(p), ptr(grad), ptr(state["exp_avg"]), ptr(state["exp_avg_sq"]), ptr(per_lr),
                                          p.numel(), group["lr"], beta1, beta2, group["eps"], group["weight_decay"],
                                          state["step"], stream_ptr()))
        return loss


def _dense(t: torch.Tensor) -> bool:
    return t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last_3d)


class CosineLR:
    """optimizer.py:231-275: multiplicative per-step decay factor = f(step) / f(step - 1) of a warm-up + cosine
    schedule (fine.py:331-336 multiplies every group's lr by it)."""

    def __init__(self, cfg, cur_step: int = 0):
        from .modules import cfg_get

        self.cfg = cfg
        self.cur_step = cur_step
        t = "app.trainer."
        self.n_iters = cfg_get(cfg, t + "n_iters")
        self.warm_up_iters = cfg_get(cfg, t + "warm_up_iters")
        if self.warm_up_iters == -1:
            self.warm_up_iters = self.n_iters
        self.warm_up_min_ratio = cfg_get(cfg, t + "warm_up_min_ratio")
        self.const_warm_up = cfg_get(cfg, t + "const_warm_up")
        self.cos_min_ratio = cfg_get(cfg, t + "cos_min_ratio")
        self.pre_decay_factor = 1.0 if cur_step == 0 else self.cosine_lr_func(cur_step - 1)
        self.pos_decay_factor = self.cosine_lr_func(cur_step)

    @property
    def decay_factor(self) -> float:
        pre = self.pre_decay_factor
        pos = self.cosine_lr_func(self.cur_step)
        self.cur_step += 1
        self.pre_decay_factor = pos
        return pos / pre

    def cosine_lr_func(self, it: int) -> float:
        if it < self.warm_up_iters:
            if self.const_warm_up:
                return self.warm_up_min_ratio
            return self.warm_up_min_ratio + (1 - self.warm_up_min_ratio) * (it / self.warm_up_iters)
        phase = (it - self.warm_up_iters) / (self.n_iters - self.warm_up_iters) * math.pi
        return (1 + math.cos(phase)) * 0.5 * (1 - self.cos_min_ratio) + self.cos_min_ratio


_ = Dict

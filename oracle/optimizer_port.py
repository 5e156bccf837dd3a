"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's optimizer update
(app/utils/optimizer.py:186-228, `adam()` as called by `Adam.step`, amsgrad off) and of `CosineLR`
(optimizer.py:231-275).  Pinned against the reference's own classes in tests/test_oracle_cpu.py
(`test_optimizer_port_matches_reference`) whenever /root/reference is present; the GPU test compares the fused kernel
with this port."""
from __future__ import annotations

import math

import torch


def adam_update(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, beta1: float, beta2: float, eps: float,
                weight_decay: float = 0.0, per_lr=None) -> None:
    """in-place on (param, exp_avg, exp_avg_sq); `step` is the 1-based step count"""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    if weight_decay != 0:
        grad = grad.add(param, alpha=weight_decay)
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    step_size = lr / bc1
    param.addcdiv_(exp_avg * per_lr if per_lr is not None else exp_avg, denom, value=-step_size)


def cosine_lr(it: int, n_iters: int, warm_up_iters: int, warm_up_min_ratio: float, const_warm_up: bool,
              cos_min_ratio: float) -> float:
    if warm_up_iters == -1:
        warm_up_iters = n_iters
    if it < warm_up_iters:
        return warm_up_min_ratio if const_warm_up else warm_up_min_ratio + (1 - warm_up_min_ratio) * (it / warm_up_iters)
    return (1 + math.cos((it - warm_up_iters) / (n_iters - warm_up_iters) * math.pi)) * 0.5 * (1 - cos_min_ratio) + cos_min_ratio

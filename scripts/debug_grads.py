"""Per-parameter gradient errors of the product (mlp_mode given on the command line, default x2) against the golden
digests made by the reference's own classes: fine-stage and LTS / PDRA cases.  Diagnostic, not a test."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import esr_testlib as C
import test_gpu_esrnerf as TE
import test_gpu_voxurff as TV

mode = sys.argv[1] if len(sys.argv) > 1 else "x2"
DEV = "cuda:0"


def report(tag, m, fx):
    worst = {}
    for name, p in m.named_parameters():
        if f"grad/{name}/idx" not in fx or p.grad is None:
            continue
        flat = p.grad.contiguous().reshape(-1).cpu()
        idx = torch.from_numpy(fx[f"grad/{name}/idx"])
        ref = torch.from_numpy(fx[f"grad/{name}/val"])
        abs_sum = float(fx[f"grad/{name}/abs_sum"])
        s_err = abs(flat.double().abs().sum().item() - abs_sum) / max(abs_sum, 1e-12)
        mx, l2 = C.grad_err(flat[idx], ref)
        worst[name] = (mx, l2, s_err)
    big = {k: v for k, v in worst.items() if max(v) > 5e-3}
    print(f"{tag}: {len(worst)} tensors, worst mx {max(v[0] for v in worst.values()):.2e} l2 {max(v[1] for v in worst.values()):.2e} "
          f"abs-sum {max(v[2] for v in worst.values()):.2e}")
    for k, v in sorted(big.items()):
        print(f"    {k:42s} mx {v[0]:.2e} l2 {v[1]:.2e} s {v[2]:.2e}")


for case in C.CASES:
    fx, weights = C.load_case(case)
    m, out = TV._run_product(fx, weights, mode, True)
    print(case, "outputs", {k: f"{C.rel_err(out[k], torch.from_numpy(fx['out/' + k])):.1e}" for k in TV.OUT_KEYS})
    report(case, m, fx)

for case in C.ESRNERF_CASES:
    fx, weights = C.load_esrnerf_case(case)
    mm = C.build_product_esrnerf(fx, weights, DEV)
    mm.mlp_mode = mode
    orig = C.build_product_esrnerf
    C.build_product_esrnerf = lambda *a, **k: mm
    try:
        m, out = TE._run_product(fx, weights)
    finally:
        C.build_product_esrnerf = orig
    cot = C.esrnerf_cotangents(out)
    sum((out[k] * cot[k].to(DEV)).sum() for k in cot).backward()
    worst_out = max(C.rel_err(out[k], torch.from_numpy(fx["out/" + k])) for k in out)
    print(case, f"outputs worst {worst_out:.1e}")
    report(case, m, fx)

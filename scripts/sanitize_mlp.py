"""Small launches of every tcgen05 kernel (both precisions, tile-overlap variants via ESR_MLP_TILE_OVERLAP) for
compute-sanitizer:

    compute-sanitizer --tool racecheck|synccheck|memcheck python scripts/sanitize_mlp.py

A few tiles per kernel (persistent CTAs with more than one tile each need > 148 tiles: the big case runs 160 tiles of
the radiance chains), results checked coarsely so that a corrupted run is visible in the log."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import test_gpu_mlp as T
from esr_nerf_b200 import fused

DEV = "cuda:0"
sizes = [int(a) for a in sys.argv[1:]] or [700, 160 * 128 + 5]
for precision in (0, 1):
    for m in sizes:
        desc = fused.with_precision(fused.RADIANCE_DESC, precision)
        flat, layers = T._flat_and_layers(desc, 3)
        g = torch.Generator().manual_seed(m)
        x = torch.randn(m, 96, generator=g)
        x[:, 91:] = 0
        d_y = torch.randn(m, 3, generator=g)
        image = fused.mlp_pack(desc, flat.to(DEV))
        xd = T._tile_with_residual(x.to(DEV)) if precision else T._tile(x.to(DEV).to(torch.bfloat16))
        y, hidden = fused._mlp_forward(desc, image, xd, 0, m, m, True)
        d_x = torch.zeros(m, 56, device=DEV)
        grad_flat, _ = fused._mlp_backward(desc, image, xd, y, d_y.to(DEV), 0, m, m, hidden, d_x, 56, 0)
        y2 = fused.mlp_infer(fused.RADIANCE_DESC, flat.to(DEV), T._tile(x.to(DEV).to(torch.bfloat16)), 0, m, m)
        torch.cuda.synchronize()
        y_ref, _, dx_ref, _ = T._fp32_reference(desc, layers, x, d_y, 0, m)
        e_y = ((y.cpu().double() - y_ref).abs().max() / y_ref.abs().max()).item()
        e_dx = ((d_x.cpu().double() - dx_ref[:, :56]).norm() / dx_ref[:, :56].norm()).item()
        print(f"radiance precision {precision} m {m}: y err {e_y:.2e}, d_x err {e_dx:.2e}, grad finite "
              f"{bool(torch.isfinite(grad_flat).all())}", flush=True)
        assert e_y < (2e-5 if precision else 2e-2) and e_dx < (2e-3 if precision else 0.15)   # bf16: ReLU masks flip (0.4 %)

        tdesc = fused.with_precision(fused.TONEMAP_DESC, precision)
        tflat, _ = T._flat_and_layers(tdesc, 11)
        lin = (torch.rand(m, 3, generator=g) * 3.0).to(DEV)
        img = fused.mlp_pack(tdesc, tflat.to(DEV))
        rgb = fused._tonemap_fwd(lin, img, tdesc)
        d_lin, g_flat = fused._tonemap_bwd(lin, img, rgb, torch.randn(m, 3, generator=g).to(DEV), None, tdesc)
        torch.cuda.synchronize()
        print(f"tonemap precision {precision} m {m}: rgb in (0,1) {bool(((rgb > 0) & (rgb < 1)).all())}, finite "
              f"{bool(torch.isfinite(d_lin).all() and torch.isfinite(g_flat).all())}", flush=True)
print("sanitize_mlp: done")

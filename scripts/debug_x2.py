"""x2 forward chain diagnostics: run-to-run determinism and mismatch positions of the saved hidden activations."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import test_gpu_mlp as T
from esr_nerf_b200 import fused

DEV = "cuda:0"
for m in (1000, 700, 40000):
    desc = fused.with_precision(fused.RADIANCE_DESC, 1)
    flat, layers = T._flat_and_layers(desc, 3)
    g = torch.Generator().manual_seed(m)
    x = torch.randn(m, 96, generator=g)
    x[:, 91:] = 0
    d_y = torch.randn(m, 3, generator=g)
    y_ref, hid_ref, dx_ref, ws = T._fp32_reference(desc, layers, x, d_y, 0, m)
    image = fused.mlp_pack(desc, flat.to(DEV))
    xd = T._tile_with_residual(x.to(DEV))
    rows = (m + 127) // 128 * 128
    outs = []
    for rep in range(4):
        y, hidden = fused._mlp_forward(desc, image, xd, 0, m, m, True)
        torch.cuda.synchronize()
        hs = [T._untile(hidden[l * rows * 192 * 2:(l + 1) * rows * 192 * 2].view(torch.bfloat16).reshape(rows, 192))[:m].float().cpu() for l in range(3)]
        outs.append((y.cpu(), hs))
    for rep in range(1, 4):
        same_y = torch.equal(outs[0][0], outs[rep][0])
        same_h = [torch.equal(a, b) for a, b in zip(outs[0][1], outs[rep][1])]
        print(f"m={m} rep{rep}: y identical {same_y}, hidden identical {same_h}")
    y, hs = outs[0]
    print("  y err", ((y.double() - y_ref).abs().max() / y_ref.abs().max()).item())
    for l in range(3):
        h = hid_ref[l]
        want = h.float().to(torch.bfloat16).float()
        bad = (hs[l] != want).nonzero()
        d = (hs[l].double() - h).abs()
        big = (d > 2 ** -8 * h.abs() + 1e-6).nonzero()
        print(f"  layer {l}: {bad.shape[0]} entries differ from bf16(h_ref); {big.shape[0]} beyond half an ulp; first: {big[:8].tolist()}")
        for r, c in big[:8].tolist():
            print(f"     row {r} col {c}: got {hs[l][r, c].item():.6f} ref {h[r, c].item():.6f}")

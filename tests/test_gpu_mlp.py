"""GPU parity (-m gpu) of the tensor-core MLP kernels behind esr_mlp_fwd / esr_mlp_bwd (tcgen05 forward and
data-gradient chains, split-K weight-gradient GEMMs) against a torch evaluation of the same network with the
kernels' numeric contract: bf16 inputs / weights / hidden activations / hidden cotangents, fp32 accumulation,
fp32 bias and output (reference nets: app/utils/pbr/module.py:6-39)."""
import pytest
import torch
import torch.nn.functional as F

from esr_nerf_b200 import fused

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _untile(buf):
    """library-private tiled activation layout [tiles][24 chunks][128 rows][8] -> row-major [rows, 192]"""
    rows = buf.shape[0]
    return buf.reshape(rows // 128, 24, 128, 8).permute(0, 2, 1, 3).reshape(rows, 192)


def _tile(x):
    """row-major [m, K] -> the library's tiled input layout, rows padded to whole 128-row tiles"""
    m, k = x.shape
    rows = (m + 127) // 128 * 128
    xp = torch.zeros(rows, k, dtype=x.dtype, device=x.device)
    xp[:m] = x
    return xp.reshape(rows // 128, 128, k // 8, 8).permute(0, 2, 1, 3).contiguous().reshape(rows, k)


def _flat_and_layers(desc, seed):
    g = torch.Generator().manual_seed(seed)
    k0, w, nh = desc["k0"], desc["width"], desc["n_hidden"]
    dims = [k0] + [w] * nh
    layers = []
    for i in range(nh):
        layers.append((torch.randn(w, dims[i], generator=g) / dims[i] ** 0.5, 0.1 * torch.randn(w, generator=g)))
    wo = torch.zeros(8, w)
    bo = torch.zeros(8)
    wo[: desc["n_out"]] = torch.randn(desc["n_out"], w, generator=g) / w ** 0.5
    bo[: desc["n_out"]] = 0.1 * torch.randn(desc["n_out"], generator=g)
    layers.append((wo, bo))
    flat = torch.cat([torch.cat([wt.reshape(-1), b]) for wt, b in layers])
    return flat, layers


def _reference(desc, layers, x, d_y, rb, re):
    """fp32 torch on the CPU with the kernels' rounding points; returns y, hidden list, d_z list, d_x, grads"""
    xs = x[rb:re].float().clone().requires_grad_(True)
    ws = [(_bf(wt).requires_grad_(True), b.clone().requires_grad_(True)) for wt, b in layers]
    h = xs
    hidden = []
    for wt, b in ws[:-1]:
        h = fused_round(F.relu(F.linear(h, wt, b)))
        hidden.append(h)
    z = F.linear(h, ws[-1][0], ws[-1][1])[:, : desc["n_out"]]
    y = F.softplus(z) if desc["act"] == 1 else torch.sigmoid(z)
    (y * d_y[rb:re]).sum().backward()
    return y.detach(), [t.detach() for t in hidden], xs.grad, ws


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _bf(x)

    @staticmethod
    def backward(ctx, g):
        return _bf(g)


fused_round = _Round.apply


@pytest.mark.parametrize("which,m,rb,re", [("radiance", 1000, 0, 1000), ("radiance", 128, 0, 128), ("radiance", 700, 130, 517),
                                           ("radiance", 40000, 0, 40000), ("tonemap", 900, 0, 900), ("tonemap", 300, 77, 300)])
def test_mlp_fwd_bwd_matches_bf16_contract(which, m, rb, re):
    desc = fused.RADIANCE_DESC if which == "radiance" else fused.TONEMAP_DESC
    dx_cols = fused.FEAT_GRAD_DIM if which == "radiance" else fused.TFEAT_GRAD_DIM
    flat, layers = _flat_and_layers(desc, 3)
    g = torch.Generator().manual_seed(m + rb)
    x = _bf(torch.randn(m, desc["k0"], generator=g))
    d_y = torch.randn(m, desc["n_out"], generator=g)
    y_ref, hid_ref, dx_ref, ws = _reference(desc, layers, x, d_y, rb, re)

    image = fused.mlp_pack(desc, flat.to(DEV))
    xd = _tile(x.to(DEV).to(torch.bfloat16))
    y, hidden = fused._mlp_forward(desc, image, xd, rb, re, m, True)
    torch.cuda.synchronize()
    assert torch.allclose(y[rb:re].cpu(), y_ref, rtol=2e-3, atol=2e-3), (y[rb:re].cpu() - y_ref).abs().max()
    if rb > 0:
        assert (y[:rb] == 0).all()                      # rows outside [rb, re) untouched
    for l, h in enumerate(hid_ref):
        rows = (m + 127) // 128 * 128
        h_l = hidden[l * rows * 192 * 2:(l + 1) * rows * 192 * 2].view(torch.bfloat16).reshape(rows, 192)
        got = _untile(h_l)[rb:re].float().cpu()
        # bf16 storage: allow 1 ulp (2^-8 relative) + accumulation-order noise
        assert torch.allclose(got, h, rtol=1e-2, atol=1e-2), (l, (got - h).abs().max())

    d_x = torch.zeros(m, dx_cols, device=DEV)
    grad_flat, d_z = fused._mlp_backward(desc, image, xd, y, d_y.to(DEV), rb, re, m, hidden, d_x, dx_cols, 0)
    torch.cuda.synchronize()
    scale = dx_ref.abs().max()
    err = (d_x[rb:re].cpu() - dx_ref[:, :dx_cols]).abs().max() / scale
    # a hidden value that sits on a bf16 rounding boundary can flip a downstream ReLU mask (one row's d_x then
    # differs by O(1)); with many rows a few such rows exist, so the max-norm is only asserted for small batches
    assert err < (2e-2 if m <= 2000 else 2e-1), err
    l2 = (d_x[rb:re].cpu() - dx_ref[:, :dx_cols]).norm() / dx_ref[:, :dx_cols].norm()
    assert l2 < 5e-3, l2
    # weight / bias gradients in the flat layout: per layer W then b
    off = 0
    for i, (wt, b) in enumerate(ws):
        n_w, n_b = layers[i][0].numel(), layers[i][1].numel()
        gw = grad_flat[off:off + n_w].cpu().reshape(layers[i][0].shape)
        gb = grad_flat[off + n_w:off + n_w + n_b].cpu()
        off += n_w + n_b
        for got, ref in ((gw, wt.grad), (gb, b.grad)):
            rel = (got - ref).norm() / ref.norm().clamp_min(1e-20)
            assert rel < 1e-2, (i, rel)


def test_mlp_accumulate_and_second_net_rows():
    """d_x accumulation across two launches over different row ranges (off net on [m_on, m), emo net on [0, m_on))."""
    desc = fused.RADIANCE_DESC
    flat, layers = _flat_and_layers(desc, 5)
    m, m_on = 600, 250
    g = torch.Generator().manual_seed(1)
    x = _bf(torch.randn(m, 96, generator=g))
    d_y = torch.randn(m, 3, generator=g)
    image = fused.mlp_pack(desc, flat.to(DEV))
    xd = _tile(x.to(DEV).to(torch.bfloat16))
    y, hidden = fused._mlp_forward(desc, image, xd, 0, m, m, True)
    d_x = torch.zeros(m, 56, device=DEV)
    fused._mlp_backward(desc, image, xd, y, d_y.to(DEV), m_on, m, m, hidden, d_x, 56, 1)
    a = d_x.clone()
    assert (a[:m_on] == 0).all() and (a[m_on:] != 0).any()
    fused._mlp_backward(desc, image, xd, y, d_y.to(DEV), 0, m, m, hidden, d_x, 56, 1)
    _, _, dx_ref, _ = _reference(desc, layers, x, d_y, 0, m)
    want = dx_ref[:, :56].clone()
    want[m_on:] *= 2
    assert ((d_x.cpu() - want).norm() / want.norm()) < 5e-3


# ---------------------------------------------------------------------------------------------------
# "x2" precision (esr_mlp_desc_t::precision = 1): fp16 hi + lo operand pairs in the forward chain (CTA pair,
# cta_group::2), bf16 backward.  Reference = the plain fp32 network (what app/utils/pbr/module.py computes).
# ---------------------------------------------------------------------------------------------------
def _fp32_reference(desc, layers, x, d_y, rb, re):
    xs = x[rb:re].double().clone().requires_grad_(True)
    ws = [(wt.double().clone().requires_grad_(True), b.double().clone().requires_grad_(True)) for wt, b in layers]
    h = xs
    hidden = []
    for wt, b in ws[:-1]:
        h = F.relu(F.linear(h, wt, b))
        hidden.append(h)
    z = F.linear(h, ws[-1][0], ws[-1][1])[:, : desc["n_out"]]
    y = F.softplus(z) if desc["act"] == 1 else torch.sigmoid(z)
    (y * d_y[rb:re].double()).sum().backward()
    return y.detach(), [t.detach() for t in hidden], xs.grad, ws


def _tile_with_residual(x):
    """fp32 rows -> [fp16 tiles | fp16 residual tiles] as esr_encode_*_fwd(out_is_bf16 = 2) writes them"""
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    return torch.cat([_tile(hi).view(torch.bfloat16), _tile(lo).view(torch.bfloat16)], 0)


@pytest.mark.parametrize("m,rb,re,n_out,act,dy_scale",
                         [(1000, 0, 1000, 3, 1, 1.0), (128, 0, 128, 3, 1, 1.0), (700, 130, 517, 3, 1, 1.0),
                          (300, 0, 300, 5, 2, 1.0), (40000, 0, 40000, 3, 1, 1.0), (75000, 0, 75000, 3, 1, 1.0),
                          # cotangents of a mean loss over 2^16 rays: ~1e-9, rows spread over four decades; and huge ones
                          (3000, 0, 3000, 3, 1, 1e-9), (3000, 100, 2900, 5, 2, 3e7)])
def test_mlp_x2_forward_is_fp32_class_and_grads_within_1e2(m, rb, re, n_out, act, dy_scale):
    desc = dict(fused.with_precision(fused.RADIANCE_DESC, 1), n_out=n_out, act=act)
    flat, layers = _flat_and_layers(desc, 3)
    g = torch.Generator().manual_seed(m + rb)
    x = torch.randn(m, 96, generator=g)
    x[:, 91:] = 0
    d_y = torch.randn(m, n_out, generator=g)
    if dy_scale != 1.0:
        d_y = d_y * dy_scale * 10.0 ** (-4.0 * torch.rand(m, 1, generator=g))
    y_ref, hid_ref, dx_ref, ws = _fp32_reference(desc, layers, x, d_y, rb, re)

    image = fused.mlp_pack(desc, flat.to(DEV))
    xd = _tile_with_residual(x.to(DEV))
    y, hidden = fused._mlp_forward(desc, image, xd, rb, re, m, True)
    torch.cuda.synchronize()
    err = ((y[rb:re].cpu().double() - y_ref).abs().max() / y_ref.abs().max()).item()
    assert err < 2e-5, err                     # x carried to 19 bits, weights / activations to 22: fp32-class outputs
    if rb > 0:
        assert (y[:rb] == 0).all()
    rows = (m + 127) // 128 * 128
    flips = 0
    for l, h in enumerate(hid_ref):
        h_l = hidden[l * rows * 192 * 2:(l + 1) * rows * 192 * 2].view(torch.float16).reshape(rows, 192)
        got = _untile(h_l)[rb:re].float().cpu().double()
        # fp16 copy of the exact value: half an fp16 ulp + the chain's own ~1e-6 absolute error
        assert torch.allclose(got, h, rtol=1.02 * 2 ** -11, atol=1e-5), (l, (got - h).abs().max())
        flips += int(((got > 0) != (h > 0)).sum())
    assert flips <= 4 + (re - rb) * 192 * 3 // 50000, flips    # masks: those of the fp32 network to ~1e-5 (vs ~0.4 % flipped with bf16)

    d_x = torch.zeros(m, 56, device=DEV)
    grad_flat, _ = fused._mlp_backward(desc, image, xd, y, d_y.to(DEV), rb, re, m, hidden, d_x, 56, 0)
    torch.cuda.synchronize()
    dx_ref = dx_ref[:, :56]
    l2 = ((d_x[rb:re].cpu().double() - dx_ref).norm() / dx_ref.norm()).item()
    assert l2 < 2e-3, l2                        # fp16 chain, row-scaled (the bf16 chain: ~4e-3)
    off = 0
    for i, (wt, b) in enumerate(ws):
        n_w, n_b = layers[i][0].numel(), layers[i][1].numel()
        gw = grad_flat[off:off + n_w].cpu().reshape(layers[i][0].shape).double()
        gb = grad_flat[off + n_w:off + n_w + n_b].cpu().double()
        off += n_w + n_b
        for got, ref in ((gw, wt.grad), (gb, b.grad)):
            mx = ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
            rel = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
            # fp16 operands, fp32 sums: ~5e-4.  One ReLU mask on the other side of zero (|z| ~ 1e-6: fp32 summation order)
            # moves one bias-gradient element by 1 / sqrt(rows) of its size — the max-norm carries that, at 1e-2
            assert rel < 3e-3 and mx < 1e-2, (i, rel, mx)


@pytest.mark.parametrize("m,dy_scale", [(900, 1.0), (128, 1.0), (5000, 1.0), (66000, 1.0), (5000, 1e-9), (40000, 3e7)])
def test_tonemap_x2_fused_kernels_vs_fp32_network(m, dy_scale):
    """esr_tonemap_mlp_fwd / _bwd with precision 1 against the fp32 tone-map net (voxurff.py:783-788, pbr/module.py:24-39)"""
    desc = fused.with_precision(fused.TONEMAP_DESC, 1)
    flat, layers = _flat_and_layers(desc, 11)
    g = torch.Generator().manual_seed(m)
    lin = (torch.rand(m, 3, generator=g) * 3.0)
    d_rgb = torch.randn(m, 3, generator=g)
    d_dir = 0.1 * torch.randn(m, 3, generator=g)
    if dy_scale != 1.0:   # rows spread over four decades below dy_scale
        row = dy_scale * 10.0 ** (-4.0 * torch.rand(m, 1, generator=g))
        d_rgb, d_dir = d_rgb * row, d_dir * row
    # fp64 reference: PE(5) of lin -> 33 -> 192 -> 3 sigmoid; internal column order of the 48-wide row (16 per channel:
    # lin, sin x5, cos x5, 5 zeros) is the layout of W0 in the flat master copy
    lr = lin.double().clone().requires_grad_(True)
    freq = torch.tensor([2.0 ** i for i in range(5)], dtype=torch.float64)
    cols = []
    for c in range(3):
        e = lr[:, c:c + 1] * freq
        cols += [lr[:, c:c + 1], e.sin(), e.cos(), torch.zeros(m, 5, dtype=torch.float64)]
    xrow = torch.cat(cols, 1)
    ws = [(wt.double().clone().requires_grad_(True), b.double().clone().requires_grad_(True)) for wt, b in layers]
    h = F.relu(F.linear(xrow, ws[0][0], ws[0][1]))
    y_ref = torch.sigmoid(F.linear(h, ws[1][0], ws[1][1])[:, :3])
    ((y_ref * d_rgb.double()).sum() + (lr * d_dir.double()).sum()).backward()

    img = fused.mlp_pack(desc, flat.to(DEV))
    lin_d = lin.to(DEV)
    rgb = fused._tonemap_fwd(lin_d, img, desc)
    torch.cuda.synchronize()
    assert ((rgb.cpu().double() - y_ref.detach()).abs().max()).item() < 2e-5
    d_lin, g_flat = fused._tonemap_bwd(lin_d, img, rgb, d_rgb.to(DEV), d_dir.to(DEV), desc)
    torch.cuda.synchronize()
    l2 = ((d_lin.cpu().double() - lr.grad).norm() / lr.grad.norm()).item()
    assert l2 < 3e-3, l2
    off = 0
    for i, (wt, b) in enumerate(ws):
        n_w, n_b = layers[i][0].numel(), layers[i][1].numel()
        gw = g_flat[off:off + n_w].cpu().reshape(layers[i][0].shape).double()
        gb = g_flat[off + n_w:off + n_w + n_b].cpu().double()
        off += n_w + n_b
        for got, ref in ((gw, wt.grad), (gb, b.grad)):
            keep = ref != 0          # padded input columns / output rows carry no gradient
            rel = ((got - ref)[keep].norm() / ref[keep].norm().clamp_min(1e-30)).item()
            mx = ((got - ref)[keep].abs().max() / ref[keep].abs().max().clamp_min(1e-300)).item()
            assert rel < 3e-3 and mx < 3e-3, (i, rel, mx)   # fp16 operands (tolerance of the path: 1e-2)


def test_mlp_x2_out_of_range_values_stay_finite():
    """features / activations beyond fp16's 65504 are clamped by the saturating conversions of the x2 path, never turned into
    infinities: outputs and gradients stay finite (the bf16 mode cannot overflow; the fp32 reference does not either)"""
    desc = fused.with_precision(fused.RADIANCE_DESC, 1)
    flat, _ = _flat_and_layers(desc, 3)
    m = 300
    g = torch.Generator().manual_seed(1)
    x = torch.randn(m, 96, generator=g)
    x[:, 91:] = 0
    x[::7, 3] = 3e5            # beyond the fp16 range: saturates in the feature tile and, through layer 0, in the activations
    x[::11, 20] = -2e6
    image = fused.mlp_pack(desc, flat.to(DEV))
    xd = _tile_with_residual(x.to(DEV).clamp(-65504, 65504))      # what esr_encode_*_fwd writes for such a feature
    y, hidden = fused._mlp_forward(desc, image, xd, 0, m, m, True)
    d_x = torch.zeros(m, 56, device=DEV)
    grad_flat, _ = fused._mlp_backward(desc, image, xd, y, torch.randn(m, 3, generator=g).to(DEV), 0, m, m, hidden, d_x, 56, 0)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(d_x).all() and torch.isfinite(grad_flat).all()
    # and the tiling kernel clamps on its own
    xw = torch.cat([x, x[:84]], 0).to(DEV).contiguous()          # 384 rows: whole tiles (rows past m are never written)
    t = fused.rows_to_tiles(xw, torch.arange(96, dtype=torch.int32, device=DEV), 1)
    assert torch.isfinite(t.view(torch.float16).float()).all()

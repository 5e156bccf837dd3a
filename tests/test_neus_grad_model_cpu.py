"""CPU model of the `neus_alpha: grad` kernels (csrc/encode.cu k_neus_cos_fwd / k_neus_cos_bwd, csrc/voxurf_stream.cu
k_alpha_scan_bwd<true> / k_sdf_scatter<true>): the arithmetic the kernels perform — the line-factorised finite-difference
taps (SdfFrame / tap_ref / for_line_corners), iter_cos, the split of dL/dalpha into dL/dsdf and dL/diter_cos, and the two
scatters into the SDF gradient volume — restated sample by sample in numpy (float64) and held to the oracle port's
autograd (functions.py:45-69 through voxurff.py:670-721).  It validates the formulas, the (z, y, x) <-> (x, y, z) flips,
the clamping at the grid border and the signs; the GPU test (tests/test_gpu_neus_grad.py) then holds the kernels to the port."""
import numpy as np
import torch

import esr_testlib as C
from esr_nerf_b200 import synthetic as S


def _frame(c, size):
    """SdfFrame of one sample: c = continuous index along (z, y, x), size = (Z, Y, X)"""
    fb = np.floor(c).astype(int)
    cl = np.clip(c, 0.0, size - 1.0)
    o0 = np.floor(cl).astype(int)
    return fb, o0, o0 + 1.0 - cl, cl - o0


def _tap_ref(c, fb, size, a, off):
    coord = min(max(c[a] + off, 0.0), size[a] - 1.0)
    fl = np.floor(coord)
    idx, wl, wh = int(fl) - (fb[a] - 2), fl + 1.0 - coord, coord - fl
    if idx < 0:
        idx, wl, wh = 0, 1.0, 0.0
    if idx > 4:
        idx, wl, wh = 4, 0.0, 1.0
    return a * 6 + idx, wl, wh, coord


def _line_corners(fb, o0, wl, wh, size, a, p):
    """(z, y, x, weight) of the in-grid corners of line plane p on axis a"""
    if not 0 <= p < size[a]:
        return
    b, c = (1 if a == 0 else 0), (1 if a == 2 else 2)
    for db in (0, 1):
        for dc in (0, 1):
            qb, qc = o0[b] + db, o0[c] + dc
            if not (0 <= qb < size[b] and 0 <= qc < size[c]):
                continue
            zyx = [0, 0, 0]
            zyx[a], zyx[b], zyx[c] = p, qb, qc
            yield zyx[0], zyx[1], zyx[2], (wh[b] if db else wl[b]) * (wh[c] if dc else wl[c])


def _model(grid, pts, view, xyz_min, xyz_max, voxel_size, dist, d_cos=None, d_sdf=None):
    """grid [X,Y,Z] float64.  Returns iter_cos[M] and, with cotangents, the gradient volume."""
    X, Y, Z = grid.shape
    size = np.array([Z, Y, X])
    u = (pts - xyz_min) / (xyz_max - xyz_min)
    cz, cy, cx = u[:, 2] * (Z - 1), u[:, 1] * (Y - 1), u[:, 0] * (X - 1)
    cos = np.zeros(len(pts))
    gvol = np.zeros_like(grid)
    for i in range(len(pts)):
        c = np.array([cz[i], cy[i], cx[i]])
        fb, o0, wl, wh = _frame(c, size)
        lines = np.zeros(18)
        for a in range(3):
            for j in range(6):
                lines[a * 6 + j] = sum(grid[x, y, z] * w for z, y, x, w in _line_corners(fb, o0, wl, wh, size, a, fb[a] - 2 + j))
        g, taps = np.zeros(3), []
        for a in range(3):
            lo, hi = _tap_ref(c, fb, size, a, -1.0), _tap_ref(c, fb, size, a, 1.0)
            f_lo = lines[lo[0]] * lo[1] + lines[lo[0] + 1] * lo[2]
            f_hi = lines[hi[0]] * hi[1] + lines[hi[0] + 1] * hi[2]
            g[a] = (f_hi - f_lo) / (hi[3] - lo[3]) / voxel_size
            taps.append((lo, hi))
        cos[i] = (view[i, 0] * g[2] + view[i, 1] * g[1] + view[i, 2] * g[0]) * dist * 0.5
        if d_cos is None:
            continue
        # k_neus_cos_bwd
        dl = np.zeros(18)
        d_dot = d_cos[i] * 0.5 * dist
        for a in range(3):
            lo, hi = taps[a]
            t = d_dot * view[i, 2 - a] / (hi[3] - lo[3]) / voxel_size
            dl[hi[0]] += t * hi[1]
            dl[hi[0] + 1] += t * hi[2]
            dl[lo[0]] -= t * lo[1]
            dl[lo[0] + 1] -= t * lo[2]
        for a in range(3):
            for j in range(6):
                if dl[a * 6 + j] != 0.0:
                    for z, y, x, w in _line_corners(fb, o0, wl, wh, size, a, fb[a] - 2 + j):
                        gvol[x, y, z] += dl[a * 6 + j] * w
        # k_sdf_scatter<true>: dL/dsdf through the sample's own trilinear cell
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    x, y, z = o0[2] + dx, o0[1] + dy, o0[0] + dz
                    if 0 <= x < X and 0 <= y < Y and 0 <= z < Z:
                        gvol[x, y, z] += d_sdf[i] * (wh[2] if dx else wl[2]) * (wh[1] if dy else wl[1]) * (wh[0] if dz else wl[0])
    return cos, gvol


def test_grad_alpha_kernel_arithmetic_matches_port_autograd():
    from oracle import voxurf_port as P

    torch.manual_seed(0)
    scene = C.oracle_scene(24 ** 3, 12, False)
    scene["neus_alpha"] = "grad"
    ws = scene["world_size"]
    sdf = (S.sphere_sdf(ws) + 0.02 * torch.randn(1, 1, *ws)).double().requires_grad_(True)
    rays = S.make_rays(48, 99)
    ray_pts, ray_id, step_id, _ = P._march(scene, rays["rays_o"], rays["rays_d"])
    # keep border samples in (clamped taps, planes outside the grid): a band around the surface plus the first / last few
    sd0 = P.grid_sample_world(sdf.detach().float(), ray_pts, scene["xyz_min"], scene["xyz_max"])[:, 0]
    idx = (ray_pts - scene["xyz_min"]) / (scene["xyz_max"] - scene["xyz_min"]) * (torch.tensor(ws) - 1)
    border = ((idx < 1) | (idx > torch.tensor(ws) - 2)).any(-1)
    keep = (sd0.abs() < 0.1) | (border & (torch.rand(len(border)) < 0.5))
    ray_pts, ray_id = ray_pts[keep].double(), ray_id[keep]
    assert 300 < len(ray_pts) < 4000, len(ray_pts)
    assert int(border[keep].sum()) >= 50                                            # clamped taps are exercised
    s_val = 25.0
    sc64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in scene.items()}
    view = rays["viewdirs"].double()
    sd = P.grid_sample_world(sdf, ray_pts, sc64["xyz_min"], sc64["xyz_max"])[:, 0]
    grad = P.sdf_fd_gradient(sc64, sdf, ray_pts)
    dist = torch.tensor(scene["stepdist"], dtype=torch.float64)
    iter_cos = (view[ray_id] * grad).sum(-1) * dist * 0.5
    alpha = P.neus_alpha_grad(view, ray_id, dist, sd, grad, s_val)
    cot = torch.randn(alpha.shape, dtype=torch.float64)
    (g_grid,) = torch.autograd.grad((alpha * cot).sum(), sdf)

    # ---- the kernels' arithmetic ----
    # k_alpha_scan_bwd<true>: dL/dalpha -> (dL/dprev_est, dL/dnext_est) -> (dL/dsdf, dL/diter_cos)
    sdn, csn, ga = sd.detach().numpy(), iter_cos.detach().numpy(), cot.numpy()
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    pc, nc = sig((sdn - csn) * s_val), sig((sdn + csn) * s_val)
    q = pc - nc
    num, den = np.maximum(q, 0.0) + 1e-5, pc + 1e-5
    rr = num / den
    live = (rr >= 0.0) & (rr <= 1.0)
    relu_g = (q > 0.0).astype(float)
    d_pc = ga * (relu_g / den - num / den ** 2)
    d_nc = ga * (-relu_g / den)
    dp = np.where(live, d_pc * pc * (1 - pc) * s_val, 0.0)
    dn = np.where(live, d_nc * nc * (1 - nc) * s_val, 0.0)
    d_sdf, d_cos = dp + dn, dn - dp
    cos_model, gvol = _model(sdf.detach()[0, 0].numpy(), ray_pts.numpy(), view[ray_id].numpy(),
                             scene["xyz_min"].double().numpy(), scene["xyz_max"].double().numpy(), scene["voxel_size"],
                             scene["stepdist"], d_cos, d_sdf)
    assert np.abs(cos_model - csn).max() < 1e-9 * max(1.0, np.abs(csn).max())
    ref = g_grid[0, 0].numpy()
    assert np.abs(ref).max() > 0
    assert np.abs(gvol - ref).max() < 1e-9 * np.abs(ref).max(), np.abs(gvol - ref).max() / np.abs(ref).max()

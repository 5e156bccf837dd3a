"""CPU model of the canonical rounds of k_encode_bwd_merged (esr_nerf_b200/csrc/encode.cu): the index arithmetic that
decides, per round, WHICH of a sample's scatter targets a lane presents.  The kernel's correctness on the GPU is checked
against the plain scatter (tests/test_gpu_voxurff.py::test_encode_backward_merged_reds_equal_plain_scatter); this file pins
the two properties the merge rests on, for every cell position incl. negative / out-of-grid ones:
  (1) completeness: over all rounds a lane presents each of its targets exactly once, with the weight the plain kernel uses;
  (2) alignment: a voxel (or aligned z-pair) is presented in the SAME round by every lane that touches it, whatever the
      lane's own cell is — the round is a function of the voxel's coordinates alone."""
import itertools

import numpy as np


def colour_rounds(x0, y0, z0):
    """round rho = (px, py, pz) -> corner (dx, dy, dz) of the cell whose voxel has those parities"""
    out = {}
    for rho in range(8):
        dx, dy, dz = ((rho >> 2) ^ x0) & 1, ((rho >> 1) ^ y0) & 1, (rho ^ z0) & 1
        out[rho] = (x0 + dx, y0 + dy, z0 + dz, dx, dy, dz)
    return out


def test_colour_rounds_cover_the_eight_corners_and_align_by_parity():
    for x0, y0, z0 in itertools.product(range(-2, 4), repeat=3):
        r = colour_rounds(x0, y0, z0)
        assert {v[3:] for v in r.values()} == set(itertools.product((0, 1), repeat=3))      # each corner once
        for rho, (x, y, z, *_) in r.items():
            assert rho == ((x & 1) << 2 | (y & 1) << 1 | (z & 1))                             # round = voxel parity


def z_line_rounds(zb):
    """z-displaced lines: planes zb .. zb + 5; round rho (pair index mod 4) -> (zz, jl0): pair (zz, zz + 1) holds
    line slots jl0 and jl0 + 1 (slots outside 0..5 contribute nothing)"""
    pf = zb >> 1
    out = {}
    for rho in range(4):
        pi = pf + ((rho - pf) & 3)
        zz = 2 * pi
        out[rho] = (zz, zz - zb)
    return out


def test_z_line_rounds_cover_the_six_planes_in_aligned_pairs():
    for zb in range(-7, 9):
        seen = []
        for rho, (zz, jl0) in z_line_rounds(zb).items():
            assert zz % 2 == 0 and (zz >> 1) % 4 == rho                  # aligned pair, round = pair index mod 4
            if jl0 < 6:                                                   # (the kernel's z_ok: jl0 == 6 is the empty round)
                seen += [(zz + e, jl0 + e) for e in (0, 1) if 0 <= jl0 + e <= 5]
        assert sorted(seen) == [(zb + j, j) for j in range(6)]           # every plane once, at its own slot


def pos_mod6(v):
    m = int(np.fmod(v, 6))      # C semantics: sign of the dividend
    return m + 6 if m < 0 else m


def plane_rounds(fb):
    """y- / x-displaced lines: planes fb - 2 .. fb + 3; round rho -> (plane, slot)"""
    m = pos_mod6(fb - 2)
    out = {}
    for rho in range(6):
        jl = rho - m
        jl += 6 if jl < 0 else 0
        out[rho] = (fb - 2 + jl, jl)
    return out


def test_plane_rounds_cover_the_window_and_align_by_plane_mod_6():
    for fb in range(-9, 12):
        r = plane_rounds(fb)
        assert sorted(v[1] for v in r.values()) == list(range(6))
        for rho, (p, jl) in r.items():
            assert p == fb - 2 + jl and p % 6 == rho                     # (python's % is the non-negative residue)


def test_other_axis_corner_rounds_align_by_parity():
    for o0 in range(-3, 5):
        got = {}
        for pc in (0, 1):
            dc = (pc ^ o0) & 1
            got[pc] = o0 + dc
            assert (o0 + dc) & 1 == pc
        assert sorted(got.values()) == [o0, o0 + 1]


def runs_reduce(keys, vals, steps=3):
    """make_runs + run_sum of the kernel on one warp: returns {head lane: sum}"""
    keys, vals = list(keys), [float(v) for v in vals]
    n = 32
    start = [(l & ((1 << steps) - 1)) == 0 or keys[l] != keys[l - 1] or keys[l] < 0 for l in range(n)]
    start[0] = True
    sm = sum(1 << l for l in range(n) if start[l]) | (1 << 32)
    after = [(sm >> (l + 1)) & 0xffffffff for l in range(n)]
    v = vals[:]
    d = 1
    while d < (1 << steps):
        nv = v[:]
        for l in range(n):
            if (after[l] & ((1 << d) - 1)) == 0:
                nv[l] = v[l] + v[l + d]          # (lane + d < 32 is implied: bit 32 of sm is a run start)
        v = nv
        d <<= 1
    return {l: v[l] for l in range(n) if start[l] and keys[l] >= 0}


def test_segmented_run_reduction_sums_every_key_exactly():
    rng = np.random.default_rng(3)
    for trial in range(200):
        # runs of random length 1..11 over random keys, some lanes without a target
        keys = []
        while len(keys) < 32:
            k = int(rng.integers(0, 50)) if rng.random() > 0.15 else -1
            keys += [k] * int(rng.integers(1, 12))
        keys = keys[:32]
        vals = rng.integers(-8, 9, 32).astype(float)       # small integers: sums are exact in any order
        heads = runs_reduce(keys, vals)
        want = {}
        for k, v in zip(keys, vals):
            if k >= 0:
                want[k] = want.get(k, 0.0) + v
        got = {}
        for l, s in heads.items():
            got[keys[l]] = got.get(keys[l], 0.0) + s
        assert got == want
        # no run is longer than 8 lanes: one RED per run
        assert all(sum(1 for l in heads if keys[l] == k) >= 1 for k in want)

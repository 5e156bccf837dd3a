// encode.cu — sample-parallel stages over the shaded (M3) stream: feature encode fwd/bwd (dense-grid
// trilinear gathers + scatter-add), tone-map encode fwd/bwd, and the per-ray compositing fwd/bwd that
// replaces torch_scatter.segment_coo.
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

#include "mlp_layout.cuh"

using namespace esr;

namespace {

// (((c/(size-1))*2-1)+1)/2*(size-1): voxurff.py:701 renormalises the clamped tap index to [-1,1] and
// ATen un-normalises it again; mirror the two roundings so tap cells match the reference.
ESR_D float renorm_index(float c, int size) {
  const float s1 = (float)(size - 1);
  const float nrm = __fsub_rn(__fmul_rn(__fdiv_rn(c, s1), 2.f), 1.f);
  return __fmul_rn(__fdiv_rn(__fadd_rn(nrm, 1.f), 2.f), s1);
}
ESR_D float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// The 24 multi-scale taps of sample_sdfeat_grad_normal (voxurff.py:678-721).
// tap t in 0..5 = (z-, z+, y-, y+, x-, x+)  [sdf_offset acts on the flipped (z,y,x) index];
// displacement k in 0..3 = grad_feat[k] voxels; every coordinate is clamped to the grid and goes through the
// reference's renormalisation round trip before the trilinear lookup.
struct TapGeom {
  float ix, iy, iz;  // continuous index along X, Y, Z of the sample
};


// ---------------------------------------------------------------------------------------------
// Line-factorised evaluation of the 24 multi-scale SDF taps.
//
// The 8 taps displaced along one axis (±0.5, ±1, ±1.5, ±2 voxels) share the sample's trilinear weights in the two
// other axes and differ only in their cell / weight along the displaced axis, and all of them land in the 6 grid
// planes fb-2 .. fb+3 (fb = floor of the sample's index on that axis).  So per axis: 6 "line values"
//     L[j] = sum over the 4 corners of the two other axes of  G[.., fb-2+j, ..] * w_other
// (24 loads) and every tap is a 2-term interpolation of two adjacent line values — 72 loads per sample instead of
// 24 taps x 8 corners = 192; the backward scatters through the same lines (72 REDs instead of 192).
// The result differs from the reference's corner-by-corner sum only by re-association (~1e-7 relative).
// Line values live in a per-thread shared-memory column (18 floats) so that the data-dependent line index of a
// tap is an address, not a register select.
// ---------------------------------------------------------------------------------------------
constexpr int ENC_THREADS = 128;
constexpr int N_LINES = 18;  // 3 axes x 6 planes

struct SdfFrame {
  float c[3];        // the sample's continuous index along (z, y, x) — the reference's flipped order
  int fb[3];         // floor(c)
  int o0[3];         // base cell of the NON-displaced coordinate (after the reference's renormalisation round trip)
  float wl[3], wh[3];  // its low / high trilinear weights
  int size[3];       // (Z, Y, X)
};

// FAST (backward kernels only): without the renormalisation round trip, which moves a coordinate by an ulp or two (two
// IEEE divisions per coordinate, a quarter of k_encode_bwd's instructions).  Interpolation is continuous in the
// coordinate, so a cotangent routed with weights that differ by 1e-7 — or, within an ulp of a plane, through the
// neighbouring cell with weight ~0 — differs from the exact routing by less than fp32 summation order does.
template <bool FAST = false>
ESR_D SdfFrame make_frame(const esr_scene_t &sc, float ix, float iy, float iz) {
  SdfFrame f;
  f.c[0] = iz, f.c[1] = iy, f.c[2] = ix;
  f.size[0] = sc.gz, f.size[1] = sc.gy, f.size[2] = sc.gx;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    f.fb[a] = (int)floorf(f.c[a]);
    const float cl = clampf(f.c[a], 0.f, (float)(f.size[a] - 1));
    const float r = FAST ? cl : renorm_index(cl, f.size[a]);
    const float fl = floorf(r);
    f.o0[a] = (int)fl;
    f.wl[a] = __fsub_rn((float)(f.o0[a] + 1), r);
    f.wh[a] = __fsub_rn(r, (float)f.o0[a]);
  }
  return f;
}

// voxel offset of (z, y, x) = coordinates along (axis0, axis1, axis2) in a [X][Y][Z] volume
ESR_D int64_t vox(const SdfFrame &f, int z, int y, int x) { return ((int64_t)x * f.size[1] + y) * f.size[0] + z; }

// the 4 (other-axes) corners of line plane `p` on axis `a`: calls fn(voxel offset, weight) for in-grid corners
template <typename Fn>
ESR_D void for_line_corners(const SdfFrame &f, int a, int p, Fn fn) {
  if ((unsigned)p >= (unsigned)f.size[a]) return;
  const int b = a == 0 ? 1 : 0, c = a == 2 ? 1 : 2;  // the two other axes (b < c)
#pragma unroll
  for (int db = 0; db < 2; ++db)
#pragma unroll
    for (int dc = 0; dc < 2; ++dc) {
      const int qb = f.o0[b] + db, qc = f.o0[c] + dc;
      if ((unsigned)qb >= (unsigned)f.size[b] || (unsigned)qc >= (unsigned)f.size[c]) continue;
      const float w = __fmul_rn(db ? f.wh[b] : f.wl[b], dc ? f.wh[c] : f.wl[c]);
      int zyx[3];
      zyx[a] = p, zyx[b] = qb, zyx[c] = qc;
      fn(vox(f, zyx[0], zyx[1], zyx[2]), w);
    }
}

ESR_D void load_lines(const SdfFrame &f, const float *__restrict__ sdf_grid, float *s_l /* [N_LINES][ENC_THREADS] */) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float acc = 0.f;
      for_line_corners(f, a, f.fb[a] - 2 + j, [&](int64_t off, float w) { acc = __fmaf_rn(__ldg(sdf_grid + off), w, acc); });
      s_l[(a * 6 + j) * ENC_THREADS + threadIdx.x] = acc;
    }
}

// tap (axis a, signed displacement): line slot of its lower plane, interpolation weights, clamped coordinate
struct TapRef {
  int slot;      // index into the per-thread line column: value = L[slot] * wl + L[slot + 1] * wh
  float wl, wh;
  float coord;   // clamped displaced coordinate (finite-difference denominator, voxurff.py:711)
};

template <bool FAST = false>
ESR_D TapRef tap_ref(const SdfFrame &f, int a, float off) {
  TapRef t;
  t.coord = clampf(__fadd_rn(f.c[a], off), 0.f, (float)(f.size[a] - 1));
  const float p = FAST ? t.coord : renorm_index(t.coord, f.size[a]);
  const float fl = floorf(p);
  int idx = (int)fl - (f.fb[a] - 2);
  t.wl = __fsub_rn(fl + 1.f, p);
  t.wh = __fsub_rn(p, fl);
  // the renormalisation round trip can move p by an ulp across a plane at the window edge: interpolation is
  // continuous there, snap to the edge plane
  if (idx < 0) idx = 0, t.wl = 1.f, t.wh = 0.f;
  if (idx > 4) idx = 4, t.wl = 0.f, t.wh = 1.f;
  t.slot = a * 6 + idx;
  return t;
}

template <int C>
ESR_D void tapC(const float *__restrict__ g, int X, int Y, int Z, const Cell &c, float *out) {
#pragma unroll
  for (int ch = 0; ch < C; ++ch) out[ch] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = c.x0 + (k >> 2), y = c.y0 + ((k >> 1) & 1), z = c.z0 + (k & 1);
    if (in_grid(x, y, z, X, Y, Z)) {
      const float2 *p = reinterpret_cast<const float2 *>(g + (((int64_t)x * Y + y) * Z + z) * C);
#pragma unroll
      for (int ch = 0; ch < C / 2; ++ch) {
        const float2 v = __ldg(p + ch);
        out[2 * ch] = __fmaf_rn(v.x, c.w[k], out[2 * ch]);
        out[2 * ch + 1] = __fmaf_rn(v.y, c.w[k], out[2 * ch + 1]);
      }
    }
  }
}

template <int C>
ESR_D void scatterC(float *__restrict__ g, int X, int Y, int Z, const Cell &c, const float *d) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = c.x0 + (k >> 2), y = c.y0 + ((k >> 1) & 1), z = c.z0 + (k & 1);
    if (in_grid(x, y, z, X, Y, Z)) {
      const int64_t vox = ((int64_t)x * Y + y) * Z + z;
      float *p = g + vox * C;
      if constexpr (C == 6) {
        // 24 bytes per voxel: 16-byte aligned for even voxels, 8 (mod 16) for odd ones -> one v4 + one v2 RED
        // either way (two L2 requests per corner instead of three)
        float e[6];
#pragma unroll
        for (int ch = 0; ch < 6; ++ch) e[ch] = d[ch] * c.w[k];
        if ((vox & 1) == 0) {
          red_add4(p, e[0], e[1], e[2], e[3]);
          red_add2(p + 4, e[4], e[5]);
        } else {
          red_add2(p, e[0], e[1]);
          red_add4(p + 2, e[2], e[3], e[4], e[5]);
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < C / 2; ++ch) red_add2(p + 2 * ch, d[2 * ch] * c.w[k], d[2 * ch + 1] * c.w[k]);
      }
    }
  }
}

template <typename OutT>
struct RowWriter;
template <>
struct RowWriter<float> {
  float *row;
  ESR_D RowWriter(float *b, int64_t r, int64_t = 0) : row(b + r * ESR_FEAT_DIM) {}
  ESR_D void finish() {}
  template <int COL0, int N>
  ESR_D void put(const float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<float2 *>(row + COL0 + i) = make_float2(v[i], v[i + 1]);
  }
  // second copy of the finished row with colour slot 0 replaced
  ESR_D void second(float *b, int64_t r, const float (&col)[6]) {
    float *dst = b + r * ESR_FEAT_DIM;
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = col[i];
#pragma unroll
    for (int i = 6; i < ESR_FEAT_DIM; i += 2) *reinterpret_cast<float2 *>(dst + i) = *reinterpret_cast<const float2 *>(row + i);
  }
};
// bf16 rows go to the MLP kernels in the TILED layout of mlp_layout.cuh ([tile][chunk][128 rows][8]): the row is
// assembled in registers (all column indices are compile-time after unrolling) and flushed as 16-byte chunks, which
// are contiguous across the consecutive rows of a warp.
template <int WIDTH>
struct TiledRowWriter {
  uint32_t w[WIDTH / 2];
  template <int COL0, int N>
  ESR_D void put(const float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      __nv_bfloat162 p = __floats2bfloat162_rn(v[i], v[i + 1]);
      w[(COL0 + i) >> 1] = *reinterpret_cast<uint32_t *>(&p);
    }
  }
  ESR_D void flush(__nv_bfloat16 *base, int64_t row) {
    uint4 *b4 = reinterpret_cast<uint4 *>(base);
#pragma unroll
    for (int c = 0; c < WIDTH / 8; ++c)
      // streaming store: the rows are read once by the MLP kernels; keep the L2 for the grid taps
      __stcs(b4 + tiled_chunk_index(row, c, WIDTH / 8), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
  }
};
template <>
struct RowWriter<__nv_bfloat16> : TiledRowWriter<ESR_FEAT_DIM> {
  __nv_bfloat16 *base;
  int64_t row;
  ESR_D RowWriter(__nv_bfloat16 *b, int64_t r, int64_t = 0) : base(b), row(r) {}
  ESR_D void finish() { flush(base, row); }
  ESR_D void second(__nv_bfloat16 *b, int64_t r, const float (&col)[6]) {
    put<0>(col);
    flush(b, r);
  }
};
// Same tiled bf16 row, written chunk by chunk as the columns arrive (puts come in increasing column order): only the
// words of the chunk in progress stay in registers instead of the whole 48-word row — the register budget that lets
// the fine-stage instantiation of k_encode_fwd run 10 blocks per SM instead of 8.
// RES (esr_mlp_desc_t::precision = 1, out_is_bf16 = 2): the row is written as fp16 (not bf16) and every 16-byte chunk is
// followed into a second tiled buffer (`res_off` 16-byte words further: the row count padded to whole tiles x 12
// chunks) by the fp16 chunk of what that rounding lost, v - fp16(v): the forward chain sees the feature to ~22 bits,
// the weight-gradient GEMM reads the first (fp16) tile.
template <bool SECOND, bool RES = false>
struct StreamRowWriter {
  uint4 *b4, *b4b;       // b4b: the second copy of the row whose colour slot 0 (columns 0..5) holds `alt` (BRDF grid taps)
  int64_t row, res_off;
  uint32_t cur[4], alt[3];
  uint32_t cur_r[RES ? 4 : 1], alt_r[RES ? 3 : 1];
  ESR_D StreamRowWriter(__nv_bfloat16 *b, int64_t r, int64_t m_total = 0)
      : b4(reinterpret_cast<uint4 *>(b)), b4b(nullptr), row(r), res_off(act_rows_padded(m_total) * (ESR_FEAT_DIM / 8)) {}
  // bf16 pair of (a, b); with RES the pair is fp16 and `res` the fp16 pair of the two rounding residuals
  ESR_D static uint32_t pack(float a, float b, uint32_t &res) {
    if constexpr (RES) {   // (saturating conversions: a feature beyond fp16's range is clamped to +-65504, never an infinity)
      uint32_t h;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
      const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&h));
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(res) : "f"(b - hf.y), "f"(a - hf.x));
      return h;
    } else {
      __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
      return *reinterpret_cast<uint32_t *>(&p);
    }
  }
  ESR_D void set_second(__nv_bfloat16 *b, const float (&col)[6]) {
    b4b = reinterpret_cast<uint4 *>(b);
#pragma unroll
    for (int i = 0; i < 3; ++i) alt[i] = pack(col[2 * i], col[2 * i + 1], alt_r[RES ? i : 0]);
  }
  template <int COL0, int N>
  ESR_D void put(const float (&v)[N]) {
    static_assert((COL0 & 1) == 0 && (N & 1) == 0, "column pairs");
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const int wi = (COL0 + i) >> 1;
      cur[wi & 3] = pack(v[i], v[i + 1], cur_r[RES ? (wi & 3) : 0]);
      if ((wi & 3) == 3) {
        const int64_t at = tiled_chunk_index(row, wi >> 2, ESR_FEAT_DIM / 8);
        __stcs(b4 + at, make_uint4(cur[0], cur[1], cur[2], cur[3]));
        if constexpr (RES) __stcs(b4 + res_off + at, make_uint4(cur_r[0], cur_r[1], cur_r[2], cur_r[3]));
        if constexpr (SECOND) {
          __stcs(b4b + at, wi == 3 ? make_uint4(alt[0], alt[1], alt[2], cur[3]) : make_uint4(cur[0], cur[1], cur[2], cur[3]));
          if constexpr (RES)
            __stcs(b4b + res_off + at, wi == 3 ? make_uint4(alt_r[0], alt_r[1], alt_r[2], cur_r[3])
                                               : make_uint4(cur_r[0], cur_r[1], cur_r[2], cur_r[3]));
        }
      }
    }
  }
  ESR_D void finish() {}
  ESR_D void second(__nv_bfloat16 *, int64_t, const float (&)[6]) {}
};

// world position + ray index of stream sample j: recomputed from (ray, step) exactly as the march kernel does, or
// read from an explicit point list (LTS points, jittered points: esrnerf.py:795-830); with explicit points h_ray is
// optional (row j looks up view direction j)
ESR_D int sample_pos(const esr_scene_t &sc, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                     const float *__restrict__ pts, const int32_t *__restrict__ h_ray,
                     const int32_t *__restrict__ h_step, int64_t j, float &px, float &py, float &pz) {
  const int r = h_ray ? h_ray[j] : (int)j;
  if (pts) {
    px = __ldg(pts + 3 * j), py = __ldg(pts + 3 * j + 1), pz = __ldg(pts + 3 * j + 2);
  } else {
    const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
    ray_point(s, sc.stepdist, h_step[j], px, py, pz);
  }
  return r;
}

constexpr int COL_SDF = 12, COL_FEAT = 13, COL_NRM = 37, COL_XYZ = 49, COL_SIN = 52, COL_COS = 67, COL_VIEW = 82;

// 8 blocks of 128 per SM (64 registers): measured 1.31 ms at config 2 against 1.37 / 1.47 ms for 10 / 12 blocks (48 / 40
// registers spill 104 / 164 bytes); the whole-row writer at 8 blocks was 1.48 ms
template <typename OutT, bool THIRD, bool RES = false>
__global__ void __launch_bounds__(ENC_THREADS, 8)
    k_encode_fwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                 const float *__restrict__ rays_d, const float *__restrict__ viewdirs,
                 const float *__restrict__ sdf_grid, const float *__restrict__ off_grid,
                 const float *__restrict__ emo_grid, const int32_t *__restrict__ h_ray,
                 const int32_t *__restrict__ h_step, const float *__restrict__ h_sdf, int64_t m3,
                 OutT *__restrict__ feat, const float *__restrict__ pts, const float *__restrict__ third_grid,
                 OutT *__restrict__ feat2, float *__restrict__ save_fd) {
  __shared__ float s_lines[N_LINES * ENC_THREADS];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  float px, py, pz;
  const int r = sample_pos(sc, rays_o, rays_d, pts, h_ray, h_step, j, px, py, pz);
  TapGeom g;
  g.ix = world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx);
  g.iy = world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy);
  g.iz = world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz);
  constexpr bool STREAM = sizeof(OutT) == 2;   // bf16 rows leave chunk by chunk; f32 rows (strict path) row-major
  typename std::conditional<STREAM, StreamRowWriter<THIRD, RES>, RowWriter<OutT>>::type wr(feat, j, m3);

  {  // colour grids (module.py:24-35), channels-last
    const Cell c = make_cell(g.ix, g.iy, g.iz);
    if constexpr (THIRD && STREAM) {  // the BRDF copy of the row leaves together with the first one
      float col3[6];
      tapC<6>(third_grid, sc.gx, sc.gy, sc.gz, c, col3);
      wr.set_second(reinterpret_cast<__nv_bfloat16 *>(feat2), col3);
    }
    float col[12];
    tapC<6>(off_grid, sc.gx, sc.gy, sc.gz, c, col);
    tapC<6>(emo_grid, sc.gx, sc.gy, sc.gz, c, col + 6);
    wr.template put<0>(col);
  }
  {  // sdf, 24 taps, 12 normal components, normalised xyz (voxurff.py:219-225)
    float v[40];
    v[0] = h_sdf[j];
    const float disp[4] = {0.5f, 1.0f, 1.5f, 2.0f};
    float coord[24];
    const SdfFrame fr = make_frame(sc, g.ix, g.iy, g.iz);
    load_lines(fr, sdf_grid, s_lines);
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const TapRef tr = tap_ref(fr, t >> 1, (t & 1) ? disp[k] : -disp[k]);
        coord[t * 4 + k] = tr.coord;
        const float lo = s_lines[tr.slot * ENC_THREADS + threadIdx.x], hi = s_lines[(tr.slot + 1) * ENC_THREADS + threadIdx.x];
        v[1 + t * 4 + k] = __fmaf_rn(hi, tr.wh, __fmul_rn(lo, tr.wl));
      }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float diff = __fadd_rn(__fsub_rn(coord[(2 * a + 1) * 4 + k], coord[(2 * a) * 4 + k]), sc.fd_eps);
        const float fd = __fsub_rn(v[1 + (2 * a + 1) * 4 + k], v[1 + (2 * a) * 4 + k]);
        gr[a] = __fdiv_rn(__fdiv_rn(fd, diff), sc.voxel_size);
      }
      const float nrm = fmaxf(sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]), 1e-12f);
#pragma unroll
      for (int a = 0; a < 3; ++a) v[25 + a * 4 + k] = __fdiv_rn(gr[a], nrm);
      // the un-normalised finite-difference gradients, f32, for the backward pass (it then needs no grid reads at all)
      if (save_fd) __stcs(reinterpret_cast<float4 *>(save_fd + j * 16 + 4 * k), make_float4(gr[0], gr[1], gr[2], 0.f));
    }
    v[37] = __fdiv_rn(__fsub_rn(px, sc.xyz_min[0]), __fsub_rn(sc.xyz_max[0], sc.xyz_min[0]));
    v[38] = __fdiv_rn(__fsub_rn(py, sc.xyz_min[1]), __fsub_rn(sc.xyz_max[1], sc.xyz_min[1]));
    v[39] = __fdiv_rn(__fsub_rn(pz, sc.xyz_min[2]), __fsub_rn(sc.xyz_max[2], sc.xyz_min[2]));
    wr.template put<COL_SDF>(v);
    float sc30[30];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        const float x = __fmul_rn(v[37 + c], (float)(1 << f));
        sc30[c * 5 + f] = sinf(x);
        sc30[15 + c * 5 + f] = cosf(x);
      }
    wr.template put<COL_SIN>(sc30);
  }
  {  // view direction encoding (viewbase_pe = 1) + zero padding
    float v[14];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __ldg(viewdirs + 3 * (int64_t)r + c);
      v[c] = x;
      v[3 + c] = sinf(x);
      v[6 + c] = cosf(x);
    }
#pragma unroll
    for (int c = 9; c < 14; ++c) v[c] = 0.f;
    wr.template put<COL_VIEW>(v);
  }
  wr.finish();
  if constexpr (THIRD && !STREAM) {  // same row with colour slot 0 taken from a third grid (BRDF grid, esrnerf.py:761-763)
    const Cell c = make_cell(g.ix, g.iy, g.iz);
    float col[6];
    tapC<6>(third_grid, sc.gx, sc.gy, sc.gz, c, col);
    wr.second(feat2, j, col);
  }
}

// Cotangents of the 18 line values of one sample (per-thread shared-memory column s_dl) from the cotangents of its 24
// taps and 12 normal components — the backward of the tap / finite-difference / normalisation chain of k_encode_fwd.
// FAST: reciprocals through MUFU.RCP (2 ulp) instead of IEEE divisions, taps without the renormalisation round trip.
template <bool FAST = false>
ESR_D void line_cotangents(const esr_scene_t &sc, const SdfFrame &fr, const float (&dv)[52], const float *__restrict__ saved_fd,
                           int64_t j, const float *__restrict__ sdf_grid, float *s_lines, float *s_dl) {
  const float disp[4] = {0.5f, 1.0f, 1.5f, 2.0f};
  // the finite-difference gradients either come from the forward pass (saved_fd: no grid reads in this kernel) or
  // are recomputed from the 18 line values
  if (!saved_fd) load_lines(fr, sdf_grid, s_lines);
#pragma unroll
  for (int i = 0; i < N_LINES; ++i) s_dl[i * ENC_THREADS + threadIdx.x] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    TapRef tr[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) tr[t] = tap_ref<FAST>(fr, t >> 1, (t & 1) ? disp[k] : -disp[k]);
    float gr[3], scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float diff = tr[2 * a + 1].coord - tr[2 * a].coord + sc.fd_eps;
      scale[a] = FAST ? __fdividef(1.f, diff * sc.voxel_size) : 1.f / diff / sc.voxel_size;
    }
    if (saved_fd) {
      const float4 q = __ldg(reinterpret_cast<const float4 *>(saved_fd + j * 16 + 4 * k));
      gr[0] = q.x, gr[1] = q.y, gr[2] = q.z;
    } else {
      float f[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        const float lo = s_lines[tr[t].slot * ENC_THREADS + threadIdx.x], hi = s_lines[(tr[t].slot + 1) * ENC_THREADS + threadIdx.x];
        f[t] = __fmaf_rn(hi, tr[t].wh, __fmul_rn(lo, tr[t].wl));
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) gr[a] = (f[2 * a + 1] - f[2 * a]) * scale[a];
    }
    const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
    const float den = fmaxf(nrm, 1e-12f);
    float dn[3], dot = 0.f;
    [[maybe_unused]] const float inv_den = FAST ? __fdividef(1.f, den) : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      dn[a] = dv[COL_NRM + a * 4 + k];
      dot += (FAST ? gr[a] * inv_den : gr[a] / den) * dn[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // d(g/max(|g|,eps)): the |g| term only exists where the clamp is inactive
      float dg;
      if constexpr (FAST)
        dg = (nrm > 1e-12f) ? (dn[a] - (gr[a] * inv_den) * dot) * inv_den : dn[a] * inv_den;
      else
        dg = (nrm > 1e-12f) ? (dn[a] - (gr[a] / den) * dot) / den : dn[a] / den;
      const float dfd = dg * scale[a];
      const float d_hi = dv[COL_FEAT + (2 * a + 1) * 4 + k] + dfd;
      const float d_lo = dv[COL_FEAT + (2 * a) * 4 + k] - dfd;
      // tap cotangent -> its two line values (per-thread column: plain read-modify-write)
      const TapRef &th = tr[2 * a + 1], &tl = tr[2 * a];
      s_dl[th.slot * ENC_THREADS + threadIdx.x] += d_hi * th.wl;
      s_dl[(th.slot + 1) * ENC_THREADS + threadIdx.x] += d_hi * th.wh;
      s_dl[tl.slot * ENC_THREADS + threadIdx.x] += d_lo * tl.wl;
      s_dl[(tl.slot + 1) * ENC_THREADS + threadIdx.x] += d_lo * tl.wh;
    }
  }
}

__global__ void __launch_bounds__(ENC_THREADS, 5)
    k_encode_bwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                 const float *__restrict__ rays_d, const float *__restrict__ sdf_grid,
                 const int32_t *__restrict__ h_ray, const int32_t *__restrict__ h_step, int64_t m3,
                 const float *__restrict__ d_feat, float *__restrict__ g_sdf, float *__restrict__ g_off,
                 float *__restrict__ g_emo, const float *__restrict__ pts, const float *__restrict__ d_third,
                 float *__restrict__ g_third, const float *__restrict__ saved_fd) {
  __shared__ float s_lines[N_LINES * ENC_THREADS], s_dl[N_LINES * ENC_THREADS];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  float px, py, pz;
  sample_pos(sc, rays_o, rays_d, pts, h_ray, h_step, j, px, py, pz);
  TapGeom g;
  g.ix = world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx);
  g.iy = world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy);
  g.iz = world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz);
  const float *d = d_feat + j * ESR_FEAT_GRAD_DIM;
  float dv[52];
#pragma unroll
  for (int i = 0; i < 52; i += 4) {
    const float4 q = __ldg(reinterpret_cast<const float4 *>(d + i));
    dv[i] = q.x, dv[i + 1] = q.y, dv[i + 2] = q.z, dv[i + 3] = q.w;
  }
  {
    const Cell c = make_cell(g.ix, g.iy, g.iz);
    // a sample feeds one of the two colour grids (emission-on rows: emo, the off net sees them through a
    // stop-gradient; emission-off rows: off): skip the grid whose cotangent is identically zero
    const bool any_off = (dv[0] != 0.f) | (dv[1] != 0.f) | (dv[2] != 0.f) | (dv[3] != 0.f) | (dv[4] != 0.f) | (dv[5] != 0.f);
    const bool any_emo = (dv[6] != 0.f) | (dv[7] != 0.f) | (dv[8] != 0.f) | (dv[9] != 0.f) | (dv[10] != 0.f) | (dv[11] != 0.f);
    if (g_off && any_off) scatterC<6>(g_off, sc.gx, sc.gy, sc.gz, c, dv);
    if (g_emo && any_emo) scatterC<6>(g_emo, sc.gx, sc.gy, sc.gz, c, dv + 6);
    if (dv[COL_SDF] != 0.f) scatter1(g_sdf, sc.gx, sc.gy, sc.gz, c, dv[COL_SDF]);
    if (g_third) {  // cotangent of the third (BRDF) grid's colour slot: [m3,6] f32
      float d3[6];
#pragma unroll
      for (int i = 0; i < 6; i += 2) {
        const float2 q = __ldg(reinterpret_cast<const float2 *>(d_third + j * 6 + i));
        d3[i] = q.x, d3[i + 1] = q.y;
      }
      scatterC<6>(g_third, sc.gx, sc.gy, sc.gz, c, d3);
    }
  }
  const SdfFrame fr = make_frame(sc, g.ix, g.iy, g.iz);
  line_cotangents(sc, fr, dv, saved_fd, j, sdf_grid, s_lines, s_dl);
  {  // z-displaced lines: the six planes fb-2 .. fb+3 of one (y, x) corner are six consecutive floats -> 8-byte REDs
    // on the even-aligned pairs (3 or 4 requests per corner instead of 6)
    float dlz[6];
#pragma unroll
    for (int jl = 0; jl < 6; ++jl) dlz[jl] = s_dl[jl * ENC_THREADS + threadIdx.x];
    const int zb = fr.fb[0] - 2, Zs = fr.size[0];
    const bool even_z = (Zs & 1) == 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int qy = fr.o0[1] + dy, qx = fr.o0[2] + dx;
        if ((unsigned)qy >= (unsigned)fr.size[1] || (unsigned)qx >= (unsigned)fr.size[2]) continue;
        const float w = __fmul_rn(dy ? fr.wh[1] : fr.wl[1], dx ? fr.wh[2] : fr.wl[2]);
        float *q = g_sdf + vox(fr, 0, qy, qx);
        auto one = [&](int jl) {
          const int z = zb + jl;
          if ((unsigned)z < (unsigned)Zs && dlz[jl] != 0.f) red_add(q + z, dlz[jl] * w);
        };
        auto two = [&](int jl) {   // zb + jl is even
          const int z = zb + jl;
          if (even_z && z >= 0 && z + 1 < Zs) {
            if (dlz[jl] != 0.f || dlz[jl + 1] != 0.f) red_add2(q + z, dlz[jl] * w, dlz[jl + 1] * w);
          } else {
            one(jl), one(jl + 1);
          }
        };
        if (zb & 1) {
          one(0), two(1), two(3), one(5);
        } else {
          two(0), two(2), two(4);
        }
      }
  }
  // line cotangents -> the 4 corners of each line plane.  For the y- and x-displaced lines (a = 1, 2) the two corners
  // that differ in z are adjacent in memory: one 8-byte RED per pair when it is aligned and inside the grid.
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int jl = 0; jl < 6; ++jl) {
      const float dl = s_dl[(a * 6 + jl) * ENC_THREADS + threadIdx.x];
      if (dl == 0.f) continue;
      const int p = fr.fb[a] - 2 + jl;
      if (a == 0) {
        continue;   // z-displaced lines: handled below, six z-adjacent planes per (y, x) corner at a time
      } else {
        if ((unsigned)p >= (unsigned)fr.size[a]) continue;
        const int c = a == 2 ? 1 : 2;                      // the other non-z axis
        const int z0 = fr.o0[0];
        const bool pair_ok = ((fr.size[0] | z0) & 1) == 0 && (unsigned)z0 < (unsigned)(fr.size[0] - 1);
#pragma unroll
        for (int dc = 0; dc < 2; ++dc) {
          const int qc = fr.o0[c] + dc;
          if ((unsigned)qc >= (unsigned)fr.size[c]) continue;
          const float wc = dc ? fr.wh[c] : fr.wl[c];
          int zyx[3];
          zyx[a] = p, zyx[c] = qc, zyx[0] = z0;
          float *q = g_sdf + vox(fr, zyx[0], zyx[1], zyx[2]);
          const float v0 = dl * __fmul_rn(fr.wl[0], wc), v1 = dl * __fmul_rn(fr.wh[0], wc);
          if (pair_ok) {
            red_add2(q, v0, v1);
          } else {
            if ((unsigned)z0 < (unsigned)fr.size[0]) red_add(q, v0);
            if ((unsigned)(z0 + 1) < (unsigned)fr.size[0]) red_add(q + 1, v1);
          }
        }
      }
    }
}


// ---------------------------------------------------------------------------------------------
// k_encode_bwd with run-merged REDs.
//
// The scatter above is bound by the number of RED requests (~75 per sample; L2 atomic units 82 % busy, DRAM at 19 %).
// Consecutive threads are consecutive samples of one ray, half a voxel apart, so a voxel that receives a contribution
// from one sample receives one from the next 2-4 samples as well (colour corners: mean chord through the 2 x 2 x 2
// footprint x 2 samples per voxel length = 2.7; line corners, 2 x 2 x 6 footprint: 3.4).  Those contributions are
// summed in the warp before they leave it:
//   * every scatter target is enumerated in CANONICAL rounds, keyed by the voxel's own coordinates — the parity of a
//     corner, the plane index modulo 6, the aligned z-pair index modulo 2 / 4 — never by the corner's role in the lane's
//     cell, so that two lanes that touch the same voxel do so in the same round whatever their cells are;
//   * in a round every lane holds (key = voxel index or -1, values); runs of equal keys over consecutive lanes (cut at
//     multiples of 8 lanes) are summed towards the run's first lane by a three-step segmented shuffle reduction, and
//     only that lane issues the RED (skipped when the sum is exactly zero).
// ncu at config 2 (profiles/r02_encode_bwd_merged.md): RED requests 19.0 M -> 11.7 M warp-level requests, L2 83 % -> 66 %
// busy with the kernel 1.45x faster, issue slots 30 % -> 50 % busy; what bounds it now is the shuffle / shared-memory
// traffic of the merge itself (L1TEX 77 % busy), not the atomics.
// Requires an even Z (aligned pairs / static 16-byte alignment of the colour corners) and < 2^31 voxels; the plain
// kernel above serves everything else (odd grids, explicit point lists whose rows are not ordered along rays).
// ---------------------------------------------------------------------------------------------
struct Runs {
  uint32_t after;   // bit d - 1: lane + d starts a new run (or lies outside the warp)
  bool head;        // first lane of a run with a valid key
};
template <int STEPS>   // runs are cut at multiples of 2^STEPS lanes
ESR_D Runs make_runs(int key, unsigned lane) {
  const int prev = __shfl_up_sync(FULL, key, 1);
  const bool start = ((lane & ((1u << STEPS) - 1u)) == 0) | (key != prev) | (key < 0);
  const uint32_t sm = __ballot_sync(FULL, start);
  Runs r;
  r.after = (uint32_t)((((uint64_t)1 << 32) | sm) >> (lane + 1));
  r.head = start & (key >= 0);
  return r;
}
template <int N, int STEPS>
ESR_D void run_sum(float (&v)[N], const Runs &r) {
#pragma unroll
  for (int d = 1; d < (1 << STEPS); d <<= 1) {
    const bool ok = (r.after & ((1u << d) - 1u)) == 0u;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float o = __shfl_down_sync(FULL, v[i], d);
      if (ok) v[i] += o;
    }
  }
}
ESR_D int pos_mod6(int v) {
  const int m = v % 6;
  return m < 0 ? m + 6 : m;
}

template <int STEPS>
__global__ void __launch_bounds__(ENC_THREADS, 6)
    k_encode_bwd_merged(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                        const float *__restrict__ rays_d, const float *__restrict__ sdf_grid,
                        const int32_t *__restrict__ h_ray, const int32_t *__restrict__ h_step, int64_t m3,
                        const float *__restrict__ d_feat, float *__restrict__ g_sdf, float *__restrict__ g_off,
                        float *__restrict__ g_emo, const float *__restrict__ d_third, float *__restrict__ g_third,
                        const float *__restrict__ saved_fd) {
  __shared__ float s_lines[N_LINES * ENC_THREADS], s_dl[N_LINES * ENC_THREADS];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const bool active = j < m3;
  const int64_t js = active ? j : m3 - 1;   // rows past the end replay the last row with zero cotangents and no keys
  float px, py, pz;
  sample_pos(sc, rays_o, rays_d, nullptr, h_ray, h_step, js, px, py, pz);
  TapGeom g;
  g.ix = world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx);
  g.iy = world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy);
  g.iz = world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz);
  const float *d = d_feat + js * ESR_FEAT_GRAD_DIM;
  float dv[52];
#pragma unroll
  for (int i = 0; i < 52; i += 4) {
    const float4 q = __ldg(reinterpret_cast<const float4 *>(d + i));
    dv[i] = q.x, dv[i + 1] = q.y, dv[i + 2] = q.z, dv[i + 3] = q.w;
  }
  if (!active) {
#pragma unroll
    for (int i = 0; i < 52; ++i) dv[i] = 0.f;
  }
  const int X = sc.gx, Y = sc.gy, Z = sc.gz;
  const SdfFrame fr = make_frame<true>(sc, g.ix, g.iy, g.iz);
  line_cotangents<true>(sc, fr, dv, saved_fd, js, sdf_grid, s_lines, s_dl);

  {  // ---- colour grids + the sdf feature: the 8 corners of the sample's cell, rounds keyed by corner parity ----
    const int x0 = (int)floorf(g.ix), y0 = (int)floorf(g.iy), z0 = (int)floorf(g.iz);
    const float z_lo = __fsub_rn((float)(z0 + 1), g.iz), z_hi = __fsub_rn(g.iz, (float)z0);
    const float y_lo = __fsub_rn((float)(y0 + 1), g.iy), y_hi = __fsub_rn(g.iy, (float)y0);
    const float x_lo = __fsub_rn((float)(x0 + 1), g.ix), x_hi = __fsub_rn(g.ix, (float)x0);
    const bool any_off = (dv[0] != 0.f) | (dv[1] != 0.f) | (dv[2] != 0.f) | (dv[3] != 0.f) | (dv[4] != 0.f) | (dv[5] != 0.f);
    const bool any_emo = (dv[6] != 0.f) | (dv[7] != 0.f) | (dv[8] != 0.f) | (dv[9] != 0.f) | (dv[10] != 0.f) | (dv[11] != 0.f);
    const bool w_off = g_off && __any_sync(FULL, any_off), w_emo = g_emo && __any_sync(FULL, any_emo);
    const bool w_sdf = __any_sync(FULL, dv[COL_SDF] != 0.f);
    float d3[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g_third && active) {
#pragma unroll
      for (int i = 0; i < 6; i += 2) {
        const float2 q = __ldg(reinterpret_cast<const float2 *>(d_third + j * 6 + i));
        d3[i] = q.x, d3[i + 1] = q.y;
      }
    }
    auto colour = [&](float *__restrict__ grid, const float *dc, float w, int vox, const Runs &runs, bool odd) {
      float e[6];
#pragma unroll
      for (int ch = 0; ch < 6; ++ch) e[ch] = dc[ch] * w;
      run_sum<6, STEPS>(e, runs);
      if (runs.head && ((e[0] != 0.f) | (e[1] != 0.f) | (e[2] != 0.f) | (e[3] != 0.f) | (e[4] != 0.f) | (e[5] != 0.f))) {
        float *p = grid + (int64_t)vox * 6;   // 24 bytes per voxel: 16-byte aligned for even voxels, 8 (mod 16) for odd ones
        if (!odd) {
          red_add4(p, e[0], e[1], e[2], e[3]);
          red_add2(p + 4, e[4], e[5]);
        } else {
          red_add2(p, e[0], e[1]);
          red_add4(p + 2, e[2], e[3], e[4], e[5]);
        }
      }
    };
#pragma unroll
    for (int rho = 0; rho < 8; ++rho) {
      const int dx = ((rho >> 2) ^ x0) & 1, dy = ((rho >> 1) ^ y0) & 1, dz = (rho ^ z0) & 1;
      const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
      const float w = __fmul_rn(__fmul_rn(dz ? z_hi : z_lo, dy ? y_hi : y_lo), dx ? x_hi : x_lo);
      const int vox = (active && in_grid(x, y, z, X, Y, Z)) ? (x * Y + y) * Z + z : -1;
      const Runs runs = make_runs<STEPS>(vox, lane);
      if (w_off) colour(g_off, dv, w, vox, runs, rho & 1);        // Z is even: the voxel index has the parity of z
      if (w_emo) colour(g_emo, dv + 6, w, vox, runs, rho & 1);
      if (g_third) colour(g_third, d3, w, vox, runs, rho & 1);
      if (w_sdf) {
        float e[1] = {dv[COL_SDF] * w};
        run_sum<1, STEPS>(e, runs);
        if (runs.head && e[0] != 0.f) red_add(g_sdf + vox, e[0]);
      }
    }
  }

  // ---- line cotangents -> grid.  One round = one aligned z-pair of voxels (8-byte RED) ----
  auto pair_round = [&](int key, float a0, float a1) {
    const Runs runs = make_runs<STEPS>(key, lane);
    float e[2] = {a0, a1};
    run_sum<2, STEPS>(e, runs);
    if (runs.head && ((e[0] != 0.f) | (e[1] != 0.f))) red_add2(g_sdf + key, e[0], e[1]);
  };
  {  // z-displaced lines: planes zb .. zb + 5 of the four (y, x) corners; rounds keyed by (pair index mod 4, y parity, x parity)
    const int zb = fr.fb[0] - 2, pf = zb >> 1;
#pragma unroll
    for (int rho = 0; rho < 4; ++rho) {
      const int pi = pf + ((rho - pf) & 3), zz = 2 * pi, jl0 = zz - zb;   // jl0 in {-1, 0, .., 6}
      const float d0 = (unsigned)jl0 < 6u ? s_dl[jl0 * ENC_THREADS + threadIdx.x] : 0.f;
      const float d1 = (unsigned)(jl0 + 1) < 6u ? s_dl[(jl0 + 1) * ENC_THREADS + threadIdx.x] : 0.f;
      const bool z_ok = active && (unsigned)zz < (unsigned)Z && jl0 < 6;
#pragma unroll
      for (int py2 = 0; py2 < 2; ++py2)
#pragma unroll
        for (int px2 = 0; px2 < 2; ++px2) {
          const int dy = (py2 ^ fr.o0[1]) & 1, dx = (px2 ^ fr.o0[2]) & 1;
          const int qy = fr.o0[1] + dy, qx = fr.o0[2] + dx;
          const float w = __fmul_rn(dy ? fr.wh[1] : fr.wl[1], dx ? fr.wh[2] : fr.wl[2]);
          const bool ok = z_ok && (unsigned)qy < (unsigned)Y && (unsigned)qx < (unsigned)X;
          pair_round(ok ? (qx * Y + qy) * Z + zz : -1, d0 * w, d1 * w);
        }
    }
  }
  // y- and x-displaced lines: plane p of the displaced axis, corner qc of the other non-z axis, z corner qz — one voxel
  // per round, keyed by (p mod 6, parity of qc, parity of qz).  (Aligned z-pairs as above would halve the requests of
  // the lanes whose z0 is even but leave half of the rounds of those lanes empty: the instruction count is what
  // bounds this kernel once the REDs are merged.)
  auto single_round = [&](int key, float v) {
    const Runs runs = make_runs<STEPS>(key, lane);
    float e[1] = {v};
    run_sum<1, STEPS>(e, runs);
    if (runs.head && e[0] != 0.f) red_add(g_sdf + key, e[0]);
  };
#pragma unroll
  for (int a = 1; a < 3; ++a) {
    const int c = a == 2 ? 1 : 2;   // the other non-z axis
    const int m_a = pos_mod6(fr.fb[a] - 2);
#pragma unroll
    for (int rho = 0; rho < 6; ++rho) {
      int jl = rho - m_a;
      jl += jl < 0 ? 6 : 0;
      const int p = fr.fb[a] - 2 + jl;
      const float dl = s_dl[(a * 6 + jl) * ENC_THREADS + threadIdx.x];
      const bool p_ok = active && (unsigned)p < (unsigned)fr.size[a];
#pragma unroll
      for (int pc = 0; pc < 2; ++pc) {
        const int dc = (pc ^ fr.o0[c]) & 1, qc = fr.o0[c] + dc;
        const float dlc = dl * (dc ? fr.wh[c] : fr.wl[c]);
        const bool c_ok = p_ok && (unsigned)qc < (unsigned)fr.size[c];
        const int qy = a == 1 ? p : qc, qx = a == 1 ? qc : p;
        const int row = (qx * Y + qy) * Z;
#pragma unroll
        for (int pz2 = 0; pz2 < 2; ++pz2) {
          const int dz = (pz2 ^ fr.o0[0]) & 1, qz = fr.o0[0] + dz;
          single_round((c_ok && (unsigned)qz < (unsigned)Z) ? row + qz : -1, dlc * (dz ? fr.wh[0] : fr.wl[0]));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Coarse-stage feature encode (voxurfc.py:205-249), thread per shaded sample.
// ---------------------------------------------------------------------------------------------
constexpr int CF_OFF = 0, CF_EMO = 12, CF_XYZ = 24, CF_SIN = 27, CF_COS = 42, CF_VIEW = 57, CF_NRM = 66;

ESR_D void coarse_sample(const esr_scene_t &sc, const float *rays_o, const float *rays_d, int r, int step, float &px,
                         float &py, float &pz, Cell &c) {
  const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  ray_point(s, sc.stepdist, step, px, py, pz);
  c = make_cell(world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx), world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy),
                world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz));
}

__global__ void __launch_bounds__(ENC_THREADS)
    k_encode_coarse_fwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                        const float *__restrict__ rays_d, const float *__restrict__ viewdirs,
                        const float *__restrict__ grad_vol, const float *__restrict__ off_grid,
                        const float *__restrict__ emo_grid, const int32_t *__restrict__ h_ray,
                        const int32_t *__restrict__ h_step, int64_t m3, float *__restrict__ feat) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  const int r = h_ray[j];
  float px, py, pz;
  Cell c;
  coarse_sample(sc, rays_o, rays_d, r, h_step[j], px, py, pz, c);
  float *row = feat + j * ESR_COARSE_FEAT_DIM;
  float col[12];
  tapC<12>(off_grid, sc.gx, sc.gy, sc.gz, c, col);
#pragma unroll
  for (int i = 0; i < 12; ++i) row[CF_OFF + i] = col[i];
  tapC<12>(emo_grid, sc.gx, sc.gy, sc.gz, c, col);
#pragma unroll
  for (int i = 0; i < 12; ++i) row[CF_EMO + i] = col[i];
  const float mn[3] = {sc.xyz_min[0], sc.xyz_min[1], sc.xyz_min[2]}, mx[3] = {sc.xyz_max[0], sc.xyz_max[1], sc.xyz_max[2]};
  const float p[3] = {px, py, pz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float u = __fdiv_rn(__fsub_rn(p[a], mn[a]), __fsub_rn(mx[a], mn[a]));
    row[CF_XYZ + a] = u;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const float x = __fmul_rn(u, (float)(1 << f));
      row[CF_SIN + a * 5 + f] = sinf(x);
      row[CF_COS + a * 5 + f] = cosf(x);
    }
    const float v = __ldg(viewdirs + 3 * (int64_t)r + a);
    row[CF_VIEW + a] = v;
    row[CF_VIEW + 3 + a] = sinf(v);
    row[CF_VIEW + 6 + a] = cosf(v);
  }
  // normal = g / (|g| + 1e-5), g = trilinear tap of the (d/dx, d/dy, d/dz) volume (voxurfc.py:206,227)
  const int64_t vol = (int64_t)sc.gx * sc.gy * sc.gz;
  float g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) g[a] = tap1(grad_vol + a * vol, sc.gx, sc.gy, sc.gz, c);
  const float den = __fadd_rn(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-5f);
#pragma unroll
  for (int a = 0; a < 3; ++a) row[CF_NRM + a] = __fdiv_rn(g[a], den);
#pragma unroll
  for (int a = 0; a < 3; ++a) row[CF_NRM + 3 + a] = 0.f;
}

__global__ void __launch_bounds__(ENC_THREADS)
    k_encode_coarse_bwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                        const float *__restrict__ rays_d, const float *__restrict__ grad_vol,
                        const int32_t *__restrict__ h_ray, const int32_t *__restrict__ h_step, int64_t m3,
                        const float *__restrict__ d_feat, float *__restrict__ g_vol, float *__restrict__ g_off,
                        float *__restrict__ g_emo) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  float px, py, pz;
  Cell c;
  coarse_sample(sc, rays_o, rays_d, h_ray[j], h_step[j], px, py, pz, c);
  const float *d = d_feat + j * ESR_COARSE_FEAT_DIM;
  float dc[12];
  bool any = false;
#pragma unroll
  for (int i = 0; i < 12; ++i) dc[i] = d[CF_OFF + i], any |= dc[i] != 0.f;
  if (g_off && any) scatterC<12>(g_off, sc.gx, sc.gy, sc.gz, c, dc);
  any = false;
#pragma unroll
  for (int i = 0; i < 12; ++i) dc[i] = d[CF_EMO + i], any |= dc[i] != 0.f;
  if (g_emo && any) scatterC<12>(g_emo, sc.gx, sc.gy, sc.gz, c, dc);
  // n = g / (|g| + eps):  dg = dn / (r + eps) - g (g . dn) / (r (r + eps)^2)
  const int64_t vol = (int64_t)sc.gx * sc.gy * sc.gz;
  float g[3], dn[3], dot = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    g[a] = tap1(grad_vol + a * vol, sc.gx, sc.gy, sc.gz, c);
    dn[a] = d[CF_NRM + a];
    dot += g[a] * dn[a];
  }
  const float rr = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), den = rr + 1e-5f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float dg = dn[a] / den - (rr > 0.f ? g[a] * dot / (rr * den * den) : 0.f);
    if (dg != 0.f) scatter1(g_vol + a * vol, sc.gx, sc.gy, sc.gz, c, dg);
  }
}

// ---------------------------------------------------------------------------------------------
// `neus_alpha: grad` in the coarse stage (voxurfc.py:171-174, 204-210): the SDF gradient of an M1 sample is the
// trilinear tap of the dense central-difference volume ([3][X][Y][Z], channels d/dx, d/dy, d/dz) — iter_cos =
// (viewdir . tap) * dist * 0.5 forward, the tap's cotangent scattered into the volume's gradient backward (the
// volume is a function of the raw SDF grid: fused.SdfCentralGradient carries it on).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_THREADS)
    k_neus_cos_vol_fwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                       const float *__restrict__ rays_d, const float *__restrict__ view,
                       const float *__restrict__ grad_vol, const int32_t *__restrict__ s_ray,
                       const int32_t *__restrict__ s_step, int64_t m1, float *__restrict__ s_cos) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m1) return;
  const int r = s_ray[j];
  float px, py, pz;
  Cell c;
  coarse_sample(sc, rays_o, rays_d, r, s_step[j], px, py, pz, c);
  const int64_t vol = (int64_t)sc.gx * sc.gy * sc.gz;
  float g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) g[a] = tap1(grad_vol + a * vol, sc.gx, sc.gy, sc.gz, c);
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(view[3 * r], g[0]), __fmul_rn(view[3 * r + 1], g[1])),
                              __fmul_rn(view[3 * r + 2], g[2]));
  s_cos[j] = __fmul_rn(__fmul_rn(dot, sc.stepdist), 0.5f);
}

__global__ void __launch_bounds__(ENC_THREADS)
    k_neus_cos_vol_bwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                       const float *__restrict__ rays_d, const float *__restrict__ view,
                       const int32_t *__restrict__ s_ray, const int32_t *__restrict__ s_step,
                       const float *__restrict__ d_cos, int64_t m1, float *__restrict__ g_vol) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m1) return;
  const float dc = d_cos[j];
  if (dc == 0.f) return;
  const int r = s_ray[j];
  float px, py, pz;
  Cell c;
  coarse_sample(sc, rays_o, rays_d, r, s_step[j], px, py, pz, c);
  const int64_t vol = (int64_t)sc.gx * sc.gy * sc.gz;
  const float d_dot = dc * 0.5f * sc.stepdist;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float dg = d_dot * view[3 * r + a];
    if (dg != 0.f) scatter1(g_vol + a * vol, sc.gx, sc.gy, sc.gz, c, dg);
  }
}

// ---------------------------------------------------------------------------------------------
// sample_sdf_grad (voxurff.py:670-676): finite-difference SDF gradient from the 6 axis taps at 1 voxel, in
// world units and (x, y, z) order — the inference path turns it into the normal map (voxurff.py:421-430).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_THREADS)
    k_sdf_fd_gradient(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                      const float *__restrict__ rays_d, const float *__restrict__ sdf_grid,
                      const int32_t *__restrict__ h_ray, const int32_t *__restrict__ h_step, int64_t m3,
                      float *__restrict__ grad_out) {
  __shared__ float s_lines[N_LINES * ENC_THREADS];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  const int r = h_ray[j];
  const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  float px, py, pz;
  ray_point(s, sc.stepdist, h_step[j], px, py, pz);
  const SdfFrame fr = make_frame(sc, world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx),
                                 world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy),
                                 world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz));
  load_lines(fr, sdf_grid, s_lines);
  float g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {  // a = 0: z, 1: y, 2: x
    const TapRef lo = tap_ref(fr, a, -1.f), hi = tap_ref(fr, a, 1.f);
    const float fl = __fmaf_rn(s_lines[(lo.slot + 1) * ENC_THREADS + threadIdx.x], lo.wh,
                               __fmul_rn(s_lines[lo.slot * ENC_THREADS + threadIdx.x], lo.wl));
    const float fh = __fmaf_rn(s_lines[(hi.slot + 1) * ENC_THREADS + threadIdx.x], hi.wh,
                               __fmul_rn(s_lines[hi.slot * ENC_THREADS + threadIdx.x], hi.wl));
    g[a] = __fdiv_rn(__fdiv_rn(__fsub_rn(fh, fl), __fsub_rn(hi.coord, lo.coord)), sc.voxel_size);
  }
  grad_out[3 * j] = g[2];
  grad_out[3 * j + 1] = g[1];
  grad_out[3 * j + 2] = g[0];
}

// ---------------------------------------------------------------------------------------------
// `neus_alpha: grad` (functions.py:45-69; voxurff.py:151-154, 193-198): the NeuS section-point estimate of every M1
// sample, iter_cos = (viewdir . grad sdf) * dist * 0.5 with grad sdf = sample_sdf_grad's finite differences (the
// six axis taps at 1 voxel, as k_sdf_fd_gradient) and dist = stepsize * voxel_size = sc.stepdist.  The alpha kernels of
// voxurf_stream.cu read it (k_neus_alpha<true>); the backward scatters dL/diter_cos through the same six taps into
// the dense SDF gradient volume (taps are linear in the grid; coordinates carry no gradient, as in the reference where
// ray_pts is a constant).  view: [n_rays,3], row = the sample's ray.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_THREADS)
    k_neus_cos_fwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                   const float *__restrict__ rays_d, const float *__restrict__ view, const float *__restrict__ sdf_grid,
                   const int32_t *__restrict__ s_ray, const int32_t *__restrict__ s_step, int64_t m1,
                   float *__restrict__ s_cos) {
  __shared__ float s_lines[N_LINES * ENC_THREADS];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m1) return;
  const int r = s_ray[j];
  const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  float px, py, pz;
  ray_point(s, sc.stepdist, s_step[j], px, py, pz);
  const SdfFrame fr = make_frame(sc, world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx),
                                 world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy),
                                 world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz));
  load_lines(fr, sdf_grid, s_lines);
  float g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {  // a = 0: z, 1: y, 2: x
    const TapRef lo = tap_ref(fr, a, -1.f), hi = tap_ref(fr, a, 1.f);
    const float fl = __fmaf_rn(s_lines[(lo.slot + 1) * ENC_THREADS + threadIdx.x], lo.wh,
                               __fmul_rn(s_lines[lo.slot * ENC_THREADS + threadIdx.x], lo.wl));
    const float fh = __fmaf_rn(s_lines[(hi.slot + 1) * ENC_THREADS + threadIdx.x], hi.wh,
                               __fmul_rn(s_lines[hi.slot * ENC_THREADS + threadIdx.x], hi.wl));
    const float diff = __fadd_rn(__fsub_rn(hi.coord, lo.coord), sc.fd_eps);
    g[a] = __fdiv_rn(__fdiv_rn(__fsub_rn(fh, fl), diff), sc.voxel_size);
  }
  // (viewdirs[ray_id] * gradients).sum(-1) * dist * 0.5 with the gradient in (x, y, z) order
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(view[3 * r], g[2]), __fmul_rn(view[3 * r + 1], g[1])),
                              __fmul_rn(view[3 * r + 2], g[0]));
  s_cos[j] = __fmul_rn(__fmul_rn(dot, sc.stepdist), 0.5f);
}

__global__ void __launch_bounds__(ENC_THREADS)
    k_neus_cos_bwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                   const float *__restrict__ rays_d, const float *__restrict__ view, const int32_t *__restrict__ s_ray,
                   const int32_t *__restrict__ s_step, const float *__restrict__ d_cos, int64_t m1,
                   float *__restrict__ grad_sdf) {
  __shared__ float s_dl[N_LINES * ENC_THREADS];
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m1) return;
  const float dc = d_cos[j];
  if (dc == 0.f) return;
  const int r = s_ray[j];
  const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  float px, py, pz;
  ray_point(s, sc.stepdist, s_step[j], px, py, pz);
  const SdfFrame fr = make_frame(sc, world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx),
                                 world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy),
                                 world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz));
#pragma unroll
  for (int l = 0; l < N_LINES; ++l) s_dl[l * ENC_THREADS + threadIdx.x] = 0.f;
  const float d_dot = dc * 0.5f * sc.stepdist;
#pragma unroll
  for (int a = 0; a < 3; ++a) {  // a = 0: z, 1: y, 2: x  <->  view component 2 - a
    const TapRef lo = tap_ref(fr, a, -1.f), hi = tap_ref(fr, a, 1.f);
    const float diff = __fadd_rn(__fsub_rn(hi.coord, lo.coord), sc.fd_eps);
    const float t = d_dot * view[3 * r + (2 - a)] / diff / sc.voxel_size;     // dL/d(fh - fl)
    s_dl[hi.slot * ENC_THREADS + threadIdx.x] += t * hi.wl;
    s_dl[(hi.slot + 1) * ENC_THREADS + threadIdx.x] += t * hi.wh;
    s_dl[lo.slot * ENC_THREADS + threadIdx.x] -= t * lo.wl;
    s_dl[(lo.slot + 1) * ENC_THREADS + threadIdx.x] -= t * lo.wh;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int jl = 0; jl < 6; ++jl) {
      const float v = s_dl[(a * 6 + jl) * ENC_THREADS + threadIdx.x];
      if (v == 0.f) continue;
      for_line_corners(fr, a, fr.fb[a] - 2 + jl, [&](int64_t off, float w) { red_add(grad_sdf + off, v * w); });
    }
}

// ---------------------------------------------------------------------------------------------
// tone-map encode (voxurff.py:243-256, 783-788)
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256)
    k_tonemap_encode_fwd(const float *__restrict__ lin_off, const float *__restrict__ lin_emo,
                         const int32_t *__restrict__ h_ray, const int64_t *__restrict__ em_modes, int64_t m3,
                         float *__restrict__ lin, OutT *__restrict__ tfeat) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  const bool on = lin_emo && em_modes[h_ray[j]] == 1;
  // internal column order (include/esr_b200.h): channel c occupies columns [16 c, 16 c + 16):
  //   16 c + 0: lin_c, 16 c + 1 + f: sin(lin_c 2^f), 16 c + 6 + f: cos(lin_c 2^f), f < 5, 16 c + 11..15: zero
  float v[48];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float x = lin_off[3 * j + c];
    if (on) x += lin_emo[3 * j + c];
    lin[3 * j + c] = x;
    v[16 * c] = x;
  }
  if (!tfeat) return;   // the fused tone-map kernels compute the encoding themselves: only `lin` is wanted
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = v[16 * c];
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const float y = __fmul_rn(x, (float)(1 << f));
      v[16 * c + 1 + f] = sinf(y);
      v[16 * c + 6 + f] = cosf(y);
    }
#pragma unroll
    for (int q = 11; q < 16; ++q) v[16 * c + q] = 0.f;
  }
  if constexpr (sizeof(OutT) == 2) {
    TiledRowWriter<ESR_TFEAT_DIM> wr;
    wr.put<0>(v);
    wr.flush(tfeat, j);
  } else {
#pragma unroll
    for (int c = 0; c < ESR_TFEAT_DIM; ++c) tfeat[j * ESR_TFEAT_DIM + c] = v[c];
  }
}

__global__ void __launch_bounds__(256)
    k_tonemap_encode_bwd(const float *__restrict__ lin, const float *__restrict__ d_tfeat,
                         const float *__restrict__ d_lin_direct, int64_t m3, float *__restrict__ d_lin) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  const float *d = d_tfeat + j * ESR_TFEAT_GRAD_DIM;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = lin[3 * j + c];
    float gacc = d[16 * c] + (d_lin_direct ? d_lin_direct[3 * j + c] : 0.f);
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const float sf = (float)(1 << f);
      const float y = x * sf;
      gacc += sf * (cosf(y) * d[16 * c + 1 + f] - sinf(y) * d[16 * c + 6 + f]);
    }
    d_lin[3 * j + c] = gacc;
  }
}

// ---------------------------------------------------------------------------------------------
// compositing: warp per ray slot (replaces segment_coo x k + the w*x elementwise products)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_composite_fwd(const int32_t *__restrict__ ray_order, int64_t n_rays, const int32_t *__restrict__ off_shade,
                    const float *__restrict__ h_w, const float *__restrict__ a, const float *__restrict__ b,
                    float *__restrict__ out_a, float *__restrict__ out_b) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t slot = warp; slot < n_rays; slot += nwarps) {
    const int r = ray_order ? ray_order[slot] : (int)slot;
    const int s = off_shade[slot], e = off_shade[slot + 1];
    float sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f};
    for (int i = s + (int)lane; i < e; i += 32) {
      const float w = h_w[i];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sa[c] = fmaf(w, a[3 * (int64_t)i + c], sa[c]);
        if (b) sb[c] = fmaf(w, b[3 * (int64_t)i + c], sb[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sa[c] = warp_sum(sa[c]);
      sb[c] = warp_sum(sb[c]);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        out_a[3 * (int64_t)r + c] = sa[c];
        if (out_b) out_b[3 * (int64_t)r + c] = sb[c];
      }
    }
  }
}

__global__ void __launch_bounds__(256)
    k_composite_bwd(const int32_t *__restrict__ h_ray, const int32_t *__restrict__ h_m1,
                    const float *__restrict__ h_w, const float *__restrict__ a, const float *__restrict__ b,
                    const float *__restrict__ c_a, const float *__restrict__ c_b, int64_t m3, float *__restrict__ d_a,
                    float *__restrict__ d_b, float *__restrict__ g_w_m1) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  const int64_t r = h_ray[j];
  const float w = h_w[j];
  float gw = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float ca = c_a[3 * r + c];
    d_a[3 * j + c] = w * ca;
    gw = fmaf(a[3 * j + c], ca, gw);
    if (b) {
      const float cb = c_b[3 * r + c];
      d_b[3 * j + c] = w * cb;
      gw = fmaf(b[3 * j + c], cb, gw);
    }
  }
  g_w_m1[h_m1 ? (int64_t)h_m1[j] : j] = gw;
}

}  // namespace

static int check_scene2(const esr_scene_t *sc) {
  ESR_CHECK_ARG(sc != nullptr);
  ESR_CHECK_ARG(sc->gx > 1 && sc->gy > 1 && sc->gz > 1 && sc->voxel_size > 0.f && sc->stepdist > 0.f);
  return ESR_OK;
}

static int encode_fwd_impl(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                           const float *sdf_grid, const float *off_color_grid, const float *emo_color_grid,
                           const float *third_grid, int color_dim, const float *pts, const int32_t *h_ray,
                           const int32_t *h_step, const float *h_sdf, int64_t m3, void *feat, void *feat2,
                           int out_is_bf16, float *save_fd, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(color_dim == 6);  // cfg/app/fine.yaml:20; other widths are not instantiated
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(viewdirs && sdf_grid && off_color_grid && emo_color_grid && h_sdf && feat);
  ESR_CHECK_ARG(pts || (rays_o && rays_d && h_ray && h_step));
  ESR_CHECK_ARG(!third_grid == !feat2);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_encode_fwd", st);
#define ESR_ENC_FWD(T, ...)                                                                                            \
  k_encode_fwd<T, __VA_ARGS__><<<cdiv(m3, 128), 128, 0, st>>>(*sc, rays_o, rays_d, viewdirs, sdf_grid, off_color_grid, \
                                                              emo_color_grid, h_ray, h_step, h_sdf, m3, (T *)feat, pts, \
                                                              third_grid, (T *)feat2, save_fd)
  if (out_is_bf16 == 2) {   // bf16 tile + fp16 residual tile (the x2 forward chain's layer-0 operand)
    if (third_grid) ESR_ENC_FWD(__nv_bfloat16, true, true);
    else ESR_ENC_FWD(__nv_bfloat16, false, true);
  } else if (out_is_bf16) {
    if (third_grid) ESR_ENC_FWD(__nv_bfloat16, true);
    else ESR_ENC_FWD(__nv_bfloat16, false);
  } else {
    if (third_grid) ESR_ENC_FWD(float, true);
    else ESR_ENC_FWD(float, false);
  }
#undef ESR_ENC_FWD
  ESR_LAUNCH_OK();
  return ESR_OK;
}

static int encode_bwd_impl(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *sdf_grid,
                           int color_dim, const float *pts, const int32_t *h_ray, const int32_t *h_step, int64_t m3,
                           const float *d_feat, const float *d_third, float *grad_sdf_grid, float *grad_off_grid,
                           float *grad_emo_grid, float *grad_third_grid, const float *saved_fd, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(color_dim == 6);
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(sdf_grid && d_feat && grad_sdf_grid);
  ESR_CHECK_ARG(pts || (rays_o && rays_d && h_ray && h_step));
  ESR_CHECK_ARG(!d_third == !grad_third_grid);
  ESR_STAGE("k_encode_bwd", (cudaStream_t)stream);
  // rows ordered along rays on an even-Z grid: REDs merged over runs of consecutive samples (ESR_ENCODE_BWD_PLAIN=1: A/B)
  const char *env = getenv("ESR_ENCODE_BWD_PLAIN");   // (read per call: the A/B test flips it inside one process)
  const bool plain = env && env[0] && env[0] != '0';
  if (!pts && !plain && (sc->gz & 1) == 0 && (int64_t)sc->gx * sc->gy * sc->gz < ((int64_t)1 << 31))
    // (measured at config 2: runs cut at 4 / 8 / 16 lanes 1.21 / 1.14 / 1.18 ms; 5 / 6 / 7 blocks per SM 1.17 / 1.14 / 1.14 ms)
    k_encode_bwd_merged<3><<<cdiv(m3, 128), 128, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, sdf_grid, h_ray, h_step, m3,
                                                                            d_feat, grad_sdf_grid, grad_off_grid, grad_emo_grid,
                                                                            d_third, grad_third_grid, saved_fd);
  else
    k_encode_bwd<<<cdiv(m3, 128), 128, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, sdf_grid, h_ray, h_step, m3,
                                                                  d_feat, grad_sdf_grid, grad_off_grid, grad_emo_grid,
                                                                  pts, d_third, grad_third_grid, saved_fd);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_encode_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                              const float *sdf_grid, const float *off_color_grid, const float *emo_color_grid,
                              int color_dim, const int32_t *h_ray, const int32_t *h_step, const float *h_sdf,
                              int64_t m3, void *feat, int out_is_bf16, esr_stream_t stream) {
  ESR_CHECK_ARG(m3 == 0 || (rays_o && rays_d && h_ray && h_step));
  return encode_fwd_impl(sc, rays_o, rays_d, viewdirs, sdf_grid, off_color_grid, emo_color_grid, nullptr, color_dim,
                         nullptr, h_ray, h_step, h_sdf, m3, feat, nullptr, out_is_bf16, nullptr, stream);
}

extern "C" int esr_encode_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *sdf_grid,
                              int color_dim, const int32_t *h_ray, const int32_t *h_step, int64_t m3,
                              const float *d_feat, float *grad_sdf_grid, float *grad_off_grid, float *grad_emo_grid,
                              esr_stream_t stream) {
  ESR_CHECK_ARG(m3 == 0 || (rays_o && rays_d && h_ray && h_step));
  return encode_bwd_impl(sc, rays_o, rays_d, sdf_grid, color_dim, nullptr, h_ray, h_step, m3, d_feat, nullptr,
                         grad_sdf_grid, grad_off_grid, grad_emo_grid, nullptr, nullptr, stream);
}

extern "C" int esr_encode_pbr_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                  const float *viewdirs, const float *sdf_grid, const float *off_color_grid,
                                  const float *emo_color_grid, const float *brdf_grid, int color_dim, const float *pts,
                                  const int32_t *h_ray, const int32_t *h_step, const float *h_sdf, int64_t m3,
                                  void *feat, void *feat_brdf, int out_is_bf16, float *save_fd, esr_stream_t stream) {
  return encode_fwd_impl(sc, rays_o, rays_d, viewdirs, sdf_grid, off_color_grid, emo_color_grid, brdf_grid, color_dim,
                         pts, h_ray, h_step, h_sdf, m3, feat, feat_brdf, out_is_bf16, save_fd, stream);
}

extern "C" int esr_encode_pbr_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                  const float *sdf_grid, int color_dim, const float *pts, const int32_t *h_ray,
                                  const int32_t *h_step, int64_t m3, const float *d_feat, const float *d_brdf_color,
                                  float *grad_sdf_grid, float *grad_off_grid, float *grad_emo_grid,
                                  float *grad_brdf_grid, const float *saved_fd, esr_stream_t stream) {
  return encode_bwd_impl(sc, rays_o, rays_d, sdf_grid, color_dim, pts, h_ray, h_step, m3, d_feat, d_brdf_color,
                         grad_sdf_grid, grad_off_grid, grad_emo_grid, grad_brdf_grid, saved_fd, stream);
}

extern "C" int esr_encode_coarse_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                     const float *viewdirs, const float *grad_vol, const float *off_color_grid,
                                     const float *emo_color_grid, const int32_t *h_ray, const int32_t *h_step,
                                     int64_t m3, float *feat, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && viewdirs && grad_vol && off_color_grid && emo_color_grid && h_ray && h_step && feat);
  ESR_STAGE("k_encode_coarse_fwd", stream);
  k_encode_coarse_fwd<<<cdiv(m3, ENC_THREADS), ENC_THREADS, 0, (cudaStream_t)stream>>>(
      *sc, rays_o, rays_d, viewdirs, grad_vol, off_color_grid, emo_color_grid, h_ray, h_step, m3, feat);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_encode_coarse_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                     const float *grad_vol, const int32_t *h_ray, const int32_t *h_step, int64_t m3,
                                     const float *d_feat, float *g_grad_vol, float *g_off_grid, float *g_emo_grid,
                                     esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && grad_vol && h_ray && h_step && d_feat && g_grad_vol);
  ESR_STAGE("k_encode_coarse_bwd", stream);
  k_encode_coarse_bwd<<<cdiv(m3, ENC_THREADS), ENC_THREADS, 0, (cudaStream_t)stream>>>(
      *sc, rays_o, rays_d, grad_vol, h_ray, h_step, m3, d_feat, g_grad_vol, g_off_grid, g_emo_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_sdf_fd_gradient(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                   const float *sdf_grid, const int32_t *h_ray, const int32_t *h_step, int64_t m3,
                                   float *grad_out, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && sdf_grid && h_ray && h_step && grad_out);
  ESR_STAGE("k_sdf_fd_gradient", stream);
  k_sdf_fd_gradient<<<cdiv(m3, ENC_THREADS), ENC_THREADS, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, sdf_grid, h_ray,
                                                                                     h_step, m3, grad_out);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_neus_cos_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                               const float *sdf_grid, const int32_t *s_ray, const int32_t *s_step, int64_t m1,
                               float *s_cos, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(m1 >= 0);
  if (m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && viewdirs && sdf_grid && s_ray && s_step && s_cos);
  ESR_STAGE("k_neus_cos_fwd", stream);
  k_neus_cos_fwd<<<cdiv(m1, ENC_THREADS), ENC_THREADS, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, viewdirs, sdf_grid,
                                                                                  s_ray, s_step, m1, s_cos);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_neus_cos_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                               const int32_t *s_ray, const int32_t *s_step, const float *d_cos, int64_t m1,
                               float *grad_sdf_grid, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(m1 >= 0);
  if (m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && viewdirs && s_ray && s_step && d_cos && grad_sdf_grid);
  ESR_STAGE("k_neus_cos_bwd", stream);
  k_neus_cos_bwd<<<cdiv(m1, ENC_THREADS), ENC_THREADS, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, viewdirs, s_ray,
                                                                                  s_step, d_cos, m1, grad_sdf_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_neus_cos_vol_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                                   const float *grad_vol, const int32_t *s_ray, const int32_t *s_step, int64_t m1,
                                   float *s_cos, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(m1 >= 0);
  if (m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && viewdirs && grad_vol && s_ray && s_step && s_cos);
  ESR_STAGE("k_neus_cos_vol_fwd", stream);
  k_neus_cos_vol_fwd<<<cdiv(m1, ENC_THREADS), ENC_THREADS, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, viewdirs,
                                                                                      grad_vol, s_ray, s_step, m1, s_cos);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_neus_cos_vol_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                                   const int32_t *s_ray, const int32_t *s_step, const float *d_cos, int64_t m1,
                                   float *g_grad_vol, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(m1 >= 0);
  if (m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && viewdirs && s_ray && s_step && d_cos && g_grad_vol);
  ESR_STAGE("k_neus_cos_vol_bwd", stream);
  k_neus_cos_vol_bwd<<<cdiv(m1, ENC_THREADS), ENC_THREADS, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, viewdirs, s_ray,
                                                                                      s_step, d_cos, m1, g_grad_vol);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_tonemap_encode_fwd(const float *lin_off, const float *lin_emo, const int32_t *h_ray,
                                      const int64_t *em_modes, int64_t m3, float *lin, void *tfeat, int out_is_bf16,
                                      esr_stream_t stream) {
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(lin_off && lin && (!lin_emo || (h_ray && em_modes)));   // tfeat nullable: combine only
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_tonemap_encode_fwd", st);
  if (out_is_bf16)
    k_tonemap_encode_fwd<__nv_bfloat16>
        <<<cdiv(m3, 256), 256, 0, st>>>(lin_off, lin_emo, h_ray, em_modes, m3, lin, (__nv_bfloat16 *)tfeat);
  else
    k_tonemap_encode_fwd<float><<<cdiv(m3, 256), 256, 0, st>>>(lin_off, lin_emo, h_ray, em_modes, m3, lin,
                                                               (float *)tfeat);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_tonemap_encode_bwd(const float *lin, const float *d_tfeat, const float *d_lin_direct, int64_t m3,
                                      float *d_lin, esr_stream_t stream) {
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(lin && d_tfeat && d_lin);
  ESR_STAGE("k_tonemap_encode_bwd", (cudaStream_t)stream);
  k_tonemap_encode_bwd<<<cdiv(m3, 256), 256, 0, (cudaStream_t)stream>>>(lin, d_tfeat, d_lin_direct, m3, d_lin);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_composite_fwd(const int32_t *ray_order, int64_t n_rays, const int32_t *off_shade, const float *h_w,
                                 const float *a, const float *b, float *out_a, float *out_b, esr_stream_t stream) {
  ESR_CHECK_ARG(n_rays >= 0);
  if (n_rays == 0) return ESR_OK;
  ESR_CHECK_ARG(off_shade && out_a && (!b || out_b));
  const int64_t want = (n_rays + 7) / 8, cap = (int64_t)num_sms() * 32;
  ESR_STAGE("k_composite_fwd", (cudaStream_t)stream);
  k_composite_fwd<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(ray_order, n_rays, off_shade,
                                                                                         h_w, a, b, out_a, out_b);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_composite_bwd(const int32_t *h_ray, const int32_t *h_m1, const float *h_w, const float *a,
                                 const float *b, const float *c_a, const float *c_b, int64_t m3, float *d_a,
                                 float *d_b, float *g_w_m1, esr_stream_t stream) {
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(h_ray && h_w && a && c_a && d_a && g_w_m1 && (!b || (c_b && d_b)));
  ESR_STAGE("k_composite_bwd", (cudaStream_t)stream);
  k_composite_bwd<<<cdiv(m3, 256), 256, 0, (cudaStream_t)stream>>>(h_ray, h_m1, h_w, a, b, c_a, c_b, m3, d_a, d_b,
                                                                   g_w_m1);
  ESR_LAUNCH_OK();
  return ESR_OK;
}


// ------------------------------------------------------------------------------------------------
// f32 rows -> the tiled 16-bit MLP input layout (the coarse stage: its 72-column feature rows are made by
// esr_encode_coarse_fwd in f32; its colour nets run on the 96 -> 192 tcgen05 chains).  Column c of the
// 96-wide input row takes column colmap[c] of the source row (or 0 for colmap[c] < 0).  precision 0: bf16 tiles;
// 1: fp16 tiles followed by the fp16 residual tiles (as esr_encode_fwd with out_is_bf16 = 2).  Thread per (row, 8-column chunk).
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
    k_rows_to_tiles(const float *__restrict__ src, int64_t m, int ld, const int32_t *__restrict__ colmap, int precision,
                    uint4 *__restrict__ tiles) {
  constexpr int CH = ESR_FEAT_DIM / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * CH) return;
  const int64_t row = i / CH;
  const int c = (int)(i - row * CH);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int sc = __ldg(colmap + 8 * c + j);
    v[j] = sc >= 0 ? __ldg(src + row * ld + sc) : 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (precision) {
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
      const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hi[j]));
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo[j]) : "f"(v[2 * j + 1] - hf.y), "f"(v[2 * j] - hf.x));
    } else {
      __nv_bfloat162 p = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      hi[j] = *reinterpret_cast<uint32_t *>(&p);
    }
  }
  const int64_t at = tiled_chunk_index(row, c, CH);
  tiles[at] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (precision) tiles[act_rows_padded(m) * CH + at] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
}  // namespace

extern "C" int esr_rows_to_mlp_tiles(const float *src, int64_t m, int ld, const int32_t *colmap, int precision, void *tiles,
                                     esr_stream_t stream) {
  ESR_CHECK_ARG(m >= 0 && ld > 0 && (precision == 0 || precision == 1));
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(src && colmap && tiles);
  ESR_STAGE("k_rows_to_tiles", stream);
  k_rows_to_tiles<<<cdiv(m * (ESR_FEAT_DIM / 8), 256), 256, 0, (cudaStream_t)stream>>>(src, m, ld, colmap, precision,
                                                                                     reinterpret_cast<uint4 *>(tiles));
  ESR_LAUNCH_OK();
  return ESR_OK;
}

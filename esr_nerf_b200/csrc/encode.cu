// encode.cu — sample-parallel stages over the shaded (M3) stream: feature encode fwd/bwd (dense-grid
// trilinear gathers + scatter-add), tone-map encode fwd/bwd, and the per-ray compositing fwd/bwd that
// replaces torch_scatter.segment_coo.
#include "common.cuh"
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
// displacement k in 0..3 = grad_feat[k] voxels.  Returns the trilinear cell of tap (t,k) and the
// clamped coordinate along the displaced axis (for the finite-difference denominator).
struct TapGeom {
  float ix, iy, iz;  // continuous index along X, Y, Z of the sample
};

ESR_D Cell tap_cell(const esr_scene_t &sc, const TapGeom &g, int t, float disp, float &axis_coord) {
  float cx = g.ix, cy = g.iy, cz = g.iz;
  const float off = (t & 1) ? disp : -disp;
  const int axis = t >> 1;  // 0: z, 1: y, 2: x
  if (axis == 0) cz = __fadd_rn(cz, off);
  if (axis == 1) cy = __fadd_rn(cy, off);
  if (axis == 2) cx = __fadd_rn(cx, off);
  cx = clampf(cx, 0.f, (float)(sc.gx - 1));
  cy = clampf(cy, 0.f, (float)(sc.gy - 1));
  cz = clampf(cz, 0.f, (float)(sc.gz - 1));
  axis_coord = axis == 0 ? cz : (axis == 1 ? cy : cx);
  return make_cell(renorm_index(cx, sc.gx), renorm_index(cy, sc.gy), renorm_index(cz, sc.gz));
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
      float *p = g + (((int64_t)x * Y + y) * Z + z) * C;
#pragma unroll
      for (int ch = 0; ch < C / 2; ++ch) red_add2(p + 2 * ch, d[2 * ch] * c.w[k], d[2 * ch + 1] * c.w[k]);
    }
  }
}

template <typename OutT>
struct RowWriter;
template <>
struct RowWriter<float> {
  float *row;
  ESR_D RowWriter(float *b, int64_t r) : row(b + r * ESR_FEAT_DIM) {}
  ESR_D void finish() {}
  template <int N>
  ESR_D void put(int col0, const float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<float2 *>(row + col0 + i) = make_float2(v[i], v[i + 1]);
  }
};
// bf16 rows go to the MLP kernels in the TILED layout of mlp_layout.cuh ([tile][chunk][128 rows][8]): the row is
// assembled in registers (all column indices are compile-time after unrolling) and flushed as 16-byte chunks, which
// are contiguous across the consecutive rows of a warp.
template <int WIDTH>
struct TiledRowWriter {
  uint32_t w[WIDTH / 2];
  template <int N>
  ESR_D void put(int col0, const float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      __nv_bfloat162 p = __floats2bfloat162_rn(v[i], v[i + 1]);
      w[(col0 + i) >> 1] = *reinterpret_cast<uint32_t *>(&p);
    }
  }
  ESR_D void flush(__nv_bfloat16 *base, int64_t row) {
    uint4 *b4 = reinterpret_cast<uint4 *>(base);
#pragma unroll
    for (int c = 0; c < WIDTH / 8; ++c)
      b4[tiled_chunk_index(row, c, WIDTH / 8)] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  }
};
template <>
struct RowWriter<__nv_bfloat16> : TiledRowWriter<ESR_FEAT_DIM> {
  __nv_bfloat16 *base;
  int64_t row;
  ESR_D RowWriter(__nv_bfloat16 *b, int64_t r) : base(b), row(r) {}
  ESR_D void finish() { flush(base, row); }
};

constexpr int COL_SDF = 12, COL_FEAT = 13, COL_NRM = 37, COL_XYZ = 49, COL_SIN = 52, COL_COS = 67, COL_VIEW = 82;

template <typename OutT>
__global__ void __launch_bounds__(128)
    k_encode_fwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                 const float *__restrict__ rays_d, const float *__restrict__ viewdirs,
                 const float *__restrict__ sdf_grid, const float *__restrict__ off_grid,
                 const float *__restrict__ emo_grid, const int32_t *__restrict__ h_ray,
                 const int32_t *__restrict__ h_step, const float *__restrict__ h_sdf, int64_t m3,
                 OutT *__restrict__ feat) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  const int r = h_ray[j];
  const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  float px, py, pz;
  ray_point(s, sc.stepdist, h_step[j], px, py, pz);
  TapGeom g;
  g.ix = world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx);
  g.iy = world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy);
  g.iz = world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz);
  RowWriter<OutT> wr(feat, j);

  {  // colour grids (module.py:24-35), channels-last
    const Cell c = make_cell(g.ix, g.iy, g.iz);
    float col[12];
    tapC<6>(off_grid, sc.gx, sc.gy, sc.gz, c, col);
    tapC<6>(emo_grid, sc.gx, sc.gy, sc.gz, c, col + 6);
    wr.put(0, col);
  }
  {  // sdf, 24 taps, 12 normal components, normalised xyz (voxurff.py:219-225)
    float v[40];
    v[0] = h_sdf[j];
    const float disp[4] = {0.5f, 1.0f, 1.5f, 2.0f};
    float coord[24];
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Cell c = tap_cell(sc, g, t, disp[k], coord[t * 4 + k]);
        v[1 + t * 4 + k] = tap1(sdf_grid, sc.gx, sc.gy, sc.gz, c);
      }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float diff = __fsub_rn(coord[(2 * a + 1) * 4 + k], coord[(2 * a) * 4 + k]);
        const float fd = __fsub_rn(v[1 + (2 * a + 1) * 4 + k], v[1 + (2 * a) * 4 + k]);
        gr[a] = __fdiv_rn(__fdiv_rn(fd, diff), sc.voxel_size);
      }
      const float nrm = fmaxf(sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]), 1e-12f);
#pragma unroll
      for (int a = 0; a < 3; ++a) v[25 + a * 4 + k] = __fdiv_rn(gr[a], nrm);
    }
    v[37] = __fdiv_rn(__fsub_rn(px, sc.xyz_min[0]), __fsub_rn(sc.xyz_max[0], sc.xyz_min[0]));
    v[38] = __fdiv_rn(__fsub_rn(py, sc.xyz_min[1]), __fsub_rn(sc.xyz_max[1], sc.xyz_min[1]));
    v[39] = __fdiv_rn(__fsub_rn(pz, sc.xyz_min[2]), __fsub_rn(sc.xyz_max[2], sc.xyz_min[2]));
    wr.put(COL_SDF, v);
    float sc30[30];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        const float x = __fmul_rn(v[37 + c], (float)(1 << f));
        sc30[c * 5 + f] = sinf(x);
        sc30[15 + c * 5 + f] = cosf(x);
      }
    wr.put(COL_SIN, sc30);
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
    wr.put(COL_VIEW, v);
  }
  wr.finish();
}

__global__ void __launch_bounds__(128)
    k_encode_bwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                 const float *__restrict__ rays_d, const float *__restrict__ sdf_grid,
                 const int32_t *__restrict__ h_ray, const int32_t *__restrict__ h_step, int64_t m3,
                 const float *__restrict__ d_feat, float *__restrict__ g_sdf, float *__restrict__ g_off,
                 float *__restrict__ g_emo) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m3) return;
  const int r = h_ray[j];
  const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  float px, py, pz;
  ray_point(s, sc.stepdist, h_step[j], px, py, pz);
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
    if (g_off) scatterC<6>(g_off, sc.gx, sc.gy, sc.gz, c, dv);
    if (g_emo) scatterC<6>(g_emo, sc.gx, sc.gy, sc.gz, c, dv + 6);
    if (dv[COL_SDF] != 0.f) scatter1(g_sdf, sc.gx, sc.gy, sc.gz, c, dv[COL_SDF]);
  }
  const float disp[4] = {0.5f, 1.0f, 1.5f, 2.0f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    Cell cells[6];
    float f[6], coord[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      cells[t] = tap_cell(sc, g, t, disp[k], coord[t]);
      f[t] = tap1(sdf_grid, sc.gx, sc.gy, sc.gz, cells[t]);
    }
    float gr[3], scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float diff = coord[2 * a + 1] - coord[2 * a];
      scale[a] = 1.f / diff / sc.voxel_size;
      gr[a] = (f[2 * a + 1] - f[2 * a]) * scale[a];
    }
    const float nrm = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
    const float den = fmaxf(nrm, 1e-12f);
    float dn[3], dot = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      dn[a] = dv[COL_NRM + a * 4 + k];
      dot += (gr[a] / den) * dn[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // d(g/max(|g|,eps)): the |g| term only exists where the clamp is inactive
      const float dg = (nrm > 1e-12f) ? (dn[a] - (gr[a] / den) * dot) / den : dn[a] / den;
      const float dfd = dg * scale[a];
      const float d_hi = dv[COL_FEAT + (2 * a + 1) * 4 + k] + dfd;
      const float d_lo = dv[COL_FEAT + (2 * a) * 4 + k] - dfd;
      if (d_hi != 0.f) scatter1(g_sdf, sc.gx, sc.gy, sc.gz, cells[2 * a + 1], d_hi);
      if (d_lo != 0.f) scatter1(g_sdf, sc.gx, sc.gy, sc.gz, cells[2 * a], d_lo);
    }
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
  float v[48];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float x = lin_off[3 * j + c];
    if (on) x += lin_emo[3 * j + c];
    lin[3 * j + c] = x;
    v[c] = x;
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const float y = __fmul_rn(x, (float)(1 << f));
      v[3 + c * 5 + f] = sinf(y);
      v[18 + c * 5 + f] = cosf(y);
    }
  }
#pragma unroll
  for (int c = 33; c < 48; ++c) v[c] = 0.f;
  if constexpr (sizeof(OutT) == 2) {
    TiledRowWriter<ESR_TFEAT_DIM> wr;
    wr.put(0, v);
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
    float gacc = d[c] + (d_lin_direct ? d_lin_direct[3 * j + c] : 0.f);
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      const float sf = (float)(1 << f);
      const float y = x * sf;
      gacc += sf * (cosf(y) * d[3 + c * 5 + f] - sinf(y) * d[18 + c * 5 + f]);
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

extern "C" int esr_encode_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                              const float *sdf_grid, const float *off_color_grid, const float *emo_color_grid,
                              int color_dim, const int32_t *h_ray, const int32_t *h_step, const float *h_sdf,
                              int64_t m3, void *feat, int out_is_bf16, esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(color_dim == 6);  // cfg/app/fine.yaml:20; other widths are not instantiated
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && viewdirs && sdf_grid && off_color_grid && emo_color_grid && h_ray && h_step &&
                h_sdf && feat);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_encode_fwd", st);
  if (out_is_bf16)
    k_encode_fwd<__nv_bfloat16><<<cdiv(m3, 128), 128, 0, st>>>(*sc, rays_o, rays_d, viewdirs, sdf_grid, off_color_grid,
                                                               emo_color_grid, h_ray, h_step, h_sdf, m3,
                                                               (__nv_bfloat16 *)feat);
  else
    k_encode_fwd<float><<<cdiv(m3, 128), 128, 0, st>>>(*sc, rays_o, rays_d, viewdirs, sdf_grid, off_color_grid,
                                                       emo_color_grid, h_ray, h_step, h_sdf, m3, (float *)feat);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_encode_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *sdf_grid,
                              int color_dim, const int32_t *h_ray, const int32_t *h_step, int64_t m3,
                              const float *d_feat, float *grad_sdf_grid, float *grad_off_grid, float *grad_emo_grid,
                              esr_stream_t stream) {
  if (int e = check_scene2(sc)) return e;
  ESR_CHECK_ARG(color_dim == 6);
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && sdf_grid && h_ray && h_step && d_feat && grad_sdf_grid);
  ESR_STAGE("k_encode_bwd", (cudaStream_t)stream);
  k_encode_bwd<<<cdiv(m3, 128), 128, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, sdf_grid, h_ray, h_step, m3,
                                                                d_feat, grad_sdf_grid, grad_off_grid, grad_emo_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_tonemap_encode_fwd(const float *lin_off, const float *lin_emo, const int32_t *h_ray,
                                      const int64_t *em_modes, int64_t m3, float *lin, void *tfeat, int out_is_bf16,
                                      esr_stream_t stream) {
  ESR_CHECK_ARG(m3 >= 0);
  if (m3 == 0) return ESR_OK;
  ESR_CHECK_ARG(lin_off && lin && tfeat && (!lin_emo || (h_ray && em_modes)));
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

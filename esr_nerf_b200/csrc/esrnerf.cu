// esrnerf.cu — kernels only the LTS / PDRA stage (ESRNeRF, app/fine/model/esrnerf.py) needs on top of the fine-stage
// set: world positions of stream samples (the origins of the secondary rays and the points the eps-jitter is added
// to) and the analytic SDF gradient of sample_sdf_expgrad (esrnerf.py:1572-1596) with its backward into the grid.
#include "common.cuh"

using namespace esr;

namespace {

// ray_pts of the reference for stream samples (kernel.cu:167-194 through the compactions of esrnerf.py:690-727)
__global__ void __launch_bounds__(256)
    k_sample_points(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                    const float *__restrict__ rays_d, const int32_t *__restrict__ h_ray,
                    const int32_t *__restrict__ h_step, int64_t m, float *__restrict__ pts) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const RaySetup s = ray_setup(rays_o, rays_d, h_ray[j], sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  float px, py, pz;
  ray_point(s, sc.stepdist, h_step[j], px, py, pz);
  pts[3 * j] = px, pts[3 * j + 1] = py, pts[3 * j + 2] = pz;
}

// Trilinear frame of differentiable_grid_sample: per-axis low / high weights from the UN-clamped floor, corner
// indices clamped (functions.py:198-229, SURVEY.md Q12: jittered points may leave the grid).
struct Frame {
  int i0[3], i1[3];    // clamped corner indices along (X, Y, Z)
  float lo[3], hi[3];  // weights of the low / high corner
  float scale[3];      // d(index)/d(world) along (x, y, z)
};

ESR_D Frame make_frame3(const esr_scene_t &sc, float px, float py, float pz) {
  Frame f;
  const float p[3] = {px, py, pz};
  const int size[3] = {sc.gx, sc.gy, sc.gz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float idx = world_to_index(p[a], sc.xyz_min[a], sc.xyz_max[a], size[a]);
    const float fl = floorf(idx);
    const int b = (int)fl;
    f.lo[a] = __fsub_rn((float)(b + 1), idx);
    f.hi[a] = __fsub_rn(idx, (float)b);
    f.i0[a] = clampi(b, size[a] - 1);
    f.i1[a] = clampi(b + 1, size[a] - 1);
    f.scale[a] = (float)(size[a] - 1) / (sc.xyz_max[a] - sc.xyz_min[a]);
  }
  return f;
}

// sdf (nullable) and d sdf / d(world xyz) at explicit points.  The value follows the reference's arithmetic exactly
// (products and sums rounded separately, order tnw .. bse with Z fastest); the gradient is the analytic derivative of
// the same expression (autograd of functions.py:231-307), equal to the reference's up to rounding.
__global__ void __launch_bounds__(256)
    k_sdf_expgrad_fwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ pts,
                      const float *__restrict__ sdf_grid, int64_t m, float *__restrict__ out_sdf,
                      float *__restrict__ out_grad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const Frame f = make_frame3(sc, __ldg(pts + 3 * j), __ldg(pts + 3 * j + 1), __ldg(pts + 3 * j + 2));
  float acc = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int bx = k >> 2, by = (k >> 1) & 1, bz = k & 1;
    const float wx = bx ? f.hi[0] : f.lo[0], wy = by ? f.hi[1] : f.lo[1], wz = bz ? f.hi[2] : f.lo[2];
    const float v = __ldg(sdf_grid + ((int64_t)(bx ? f.i1[0] : f.i0[0]) * sc.gy + (by ? f.i1[1] : f.i0[1])) * sc.gz +
                          (bz ? f.i1[2] : f.i0[2]));
    const float term = __fmul_rn(v, __fmul_rn(__fmul_rn(wz, wy), wx));
    acc = k == 0 ? term : __fadd_rn(acc, term);
    gx += v * (wy * wz) * (bx ? 1.f : -1.f);
    gy += v * (wx * wz) * (by ? 1.f : -1.f);
    gz += v * (wx * wy) * (bz ? 1.f : -1.f);
  }
  if (out_sdf) out_sdf[j] = acc;
  if (out_grad) {
    out_grad[3 * j] = gx * f.scale[0];
    out_grad[3 * j + 1] = gy * f.scale[1];
    out_grad[3 * j + 2] = gz * f.scale[2];
  }
}

// backward of both outputs into the dense SDF gradient volume (both are linear in the grid values)
__global__ void __launch_bounds__(256)
    k_sdf_expgrad_bwd(const __grid_constant__ esr_scene_t sc, const float *__restrict__ pts, int64_t m,
                      const float *__restrict__ g_sdf, const float *__restrict__ g_grad, float *__restrict__ grad_grid) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const Frame f = make_frame3(sc, __ldg(pts + 3 * j), __ldg(pts + 3 * j + 1), __ldg(pts + 3 * j + 2));
  const float gs = g_sdf ? g_sdf[j] : 0.f;
  const float cx = g_grad ? g_grad[3 * j] * f.scale[0] : 0.f;
  const float cy = g_grad ? g_grad[3 * j + 1] * f.scale[1] : 0.f;
  const float cz = g_grad ? g_grad[3 * j + 2] * f.scale[2] : 0.f;
  if (gs == 0.f && cx == 0.f && cy == 0.f && cz == 0.f) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int bx = k >> 2, by = (k >> 1) & 1, bz = k & 1;
    const float wx = bx ? f.hi[0] : f.lo[0], wy = by ? f.hi[1] : f.lo[1], wz = bz ? f.hi[2] : f.lo[2];
    const float d = gs * (wx * wy * wz) + cx * (wy * wz) * (bx ? 1.f : -1.f) + cy * (wx * wz) * (by ? 1.f : -1.f) +
                    cz * (wx * wy) * (bz ? 1.f : -1.f);
    if (d != 0.f)
      red_add(grad_grid + ((int64_t)(bx ? f.i1[0] : f.i0[0]) * sc.gy + (by ? f.i1[1] : f.i0[1])) * sc.gz +
                  (bz ? f.i1[2] : f.i0[2]),
              d);
  }
}

// F.grid_sample arithmetic (zeros padding, FMA accumulation) at explicit points: sample_sdf_grad's sdf at the
// emit_eps-jittered points (esrnerf.py:819)
__global__ void __launch_bounds__(256)
    k_sdf_tap_points(const __grid_constant__ esr_scene_t sc, const float *__restrict__ pts,
                     const float *__restrict__ sdf_grid, int64_t m, float *__restrict__ out_sdf) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  out_sdf[j] = tap1_world(sdf_grid, sc.gx, sc.gy, sc.gz, sc.xyz_min, sc.xyz_max, __ldg(pts + 3 * j), __ldg(pts + 3 * j + 1),
                          __ldg(pts + 3 * j + 2));
}


// ---------------------------------------------------------------------------------------------
// Light-transport accumulation (esrnerf.py:556-574, 654-677): for LTS point p with unit normal n, BRDF parameters
// (base colour a, roughness r, metallic m) and n2 hemisphere directions w_j, the Monte-Carlo estimates
//     off_hat[v] = mean_j (L_off_j + env_j) R(w_j, wo_v),   reflect[v] = mean_j L_emo_j R(w_j, wo_v)
// for the two outgoing directions wo_0 = -view, wo_1 = -random view, with the Disney-style R of
// pbr/functions.py:108-173.  One warp per point, lanes over the secondary rays, warp reduction; replaces ~60 elementwise
// torch kernels forward (and ~120 backward) on [2 P n2, 3] tensors.
// ---------------------------------------------------------------------------------------------
struct Brdf {
  float R[3];                      // reflectance (rgb)
  float dR_da[3], dR_dm[3];        // d R_c / d a_c (diagonal), d R_c / d m
  float dR_dr[3];                  // d R_c / d r
};

ESR_D float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

template <bool GRAD>
ESR_D Brdf disney(const float *a, float r, float m, const float *n, const float *wi, const float *wo) {
  constexpr float EPS = 1e-7f, PI = 3.14159265358979323846f;
  Brdf b;
  float h[3] = {wi[0] + wo[0], wi[1] + wo[1], wi[2] + wo[2]};
  const float hn = fmaxf(sqrtf(dot3(h, h)), 1e-12f);   // F.normalize eps
  h[0] /= hn, h[1] /= hn, h[2] /= hn;
  const float noh = fmaxf(dot3(n, h), 0.f), ooh = fmaxf(dot3(wo, h), 0.f);
  const float ion = fmaxf(dot3(wi, n), 0.f), oon = fmaxf(dot3(wo, n), 0.f);
  const float r2 = fmaxf(r * r, EPS);
  const float D = (1.f / (r2 * PI)) * expf((2.f / r2) * (noh - 1.f));
  const float k = (1.f + r) * (1.f + r) / 8.f;
  const float di = ion * (1.f - k) + k, dO = oon * (1.f - k) + k;
  const float vi = 0.5f / fmaxf(di, EPS), vo = 0.5f / fmaxf(dO, EPS);
  const float V = vi * vo;
  const float q = 1.f - ooh, q5 = q * q * q * q * q;
  const float scale = ion * PI * 2.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float F0 = 0.04f * (1.f - m) + a[c] * m;
    const float F = F0 + (1.f - F0) * q5;
    b.R[c] = ((1.f - m) * a[c] / PI + D * F * V) * scale;
    if (GRAD) {
      b.dR_da[c] = ((1.f - m) / PI + D * V * m * (1.f - q5)) * scale;
      b.dR_dm[c] = (-a[c] / PI + D * V * (a[c] - 0.04f) * (1.f - q5)) * scale;
      // d D / d r (through r2 = r^2 where the clamp is inactive), d V / d r (through k = (1 + r)^2 / 8)
      const float dD = (r * r > EPS) ? D * (-2.f * (noh - 1.f) / (r2 * r2) - 1.f / r2) * 2.f * r : 0.f;
      const float dk = (1.f + r) / 4.f;
      const float dvi = (di > EPS) ? -0.5f * (1.f - ion) / (di * di) * dk : 0.f;
      const float dvo = (dO > EPS) ? -0.5f * (1.f - oon) / (dO * dO) * dk : 0.f;
      b.dR_dr[c] = F * (dD * V + D * (dvi * vo + vi * dvo)) * scale;
    }
  }
  return b;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
    k_lts_accumulate(const float *__restrict__ normal, const float *__restrict__ base, const float *__restrict__ rough,
                     const float *__restrict__ metal, const float *__restrict__ wo_a, const float *__restrict__ wo_b,
                     const float *__restrict__ dirs, const float *__restrict__ rad_off, const float *__restrict__ rad_emo,
                     int64_t P, int n2, float *__restrict__ off_hat, float *__restrict__ reflect,
                     const float *__restrict__ g_off_hat, const float *__restrict__ g_reflect, float *__restrict__ g_base,
                     float *__restrict__ g_rough, float *__restrict__ g_metal, float *__restrict__ g_rad_off,
                     float *__restrict__ g_rad_emo) {
  const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = lane_id();
  if (p >= P) return;
  float n[3], a[3], wo[2][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    n[c] = __ldg(normal + 3 * p + c), a[c] = __ldg(base + 3 * p + c);
    wo[0][c] = __ldg(wo_a + 3 * p + c), wo[1][c] = __ldg(wo_b + 3 * p + c);
  }
  const float r = __ldg(rough + p), m = __ldg(metal + p);
  const float inv = 1.f / (float)n2;
  float acc_off[2][3] = {}, acc_emo[2][3] = {};       // forward sums
  float ga[3] = {}, gr = 0.f, gm = 0.f;                // backward sums over (j, v)
  float go[2][3], ge[2][3];
  if (BWD) {
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        go[v][c] = g_off_hat ? __ldg(g_off_hat + 3 * (v * P + p) + c) * inv : 0.f;
        ge[v][c] = __ldg(g_reflect + 3 * (v * P + p) + c) * inv;
      }
  }
  for (int j = (int)lane; j < n2; j += 32) {
    const int64_t ray = p * n2 + j;
    float wi[3], lo[3] = {0.f, 0.f, 0.f}, le[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      wi[c] = __ldg(dirs + 3 * ray + c);
      if (rad_off) lo[c] = __ldg(rad_off + 3 * ray + c);
      le[c] = __ldg(rad_emo + 3 * ray + c);
    }
    float d_lo[3] = {0.f, 0.f, 0.f}, d_le[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const Brdf b = disney<BWD>(a, r, m, n, wi, wo[v]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (!BWD) {
          acc_off[v][c] += lo[c] * b.R[c];
          acc_emo[v][c] += le[c] * b.R[c];
        } else {
          const float gR = go[v][c] * lo[c] + ge[v][c] * le[c];   // cotangent of R_c for this (ray, view)
          ga[c] += gR * b.dR_da[c];
          gm += gR * b.dR_dm[c];
          gr += gR * b.dR_dr[c];
          d_lo[c] += go[v][c] * b.R[c];
          d_le[c] += ge[v][c] * b.R[c];
        }
      }
    }
    if (BWD) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (g_rad_off) g_rad_off[3 * ray + c] = d_lo[c];
        g_rad_emo[3 * ray + c] = d_le[c];
      }
    }
  }
  if (!BWD) {
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float so = warp_sum(acc_off[v][c]), se = warp_sum(acc_emo[v][c]);
        if (lane == 0) {
          if (off_hat) off_hat[3 * (v * P + p) + c] = so * inv;
          reflect[3 * (v * P + p) + c] = se * inv;
        }
      }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float s = warp_sum(ga[c]);
      if (lane == 0) g_base[3 * p + c] = s;
    }
    gr = warp_sum(gr), gm = warp_sum(gm);
    if (lane == 0) g_rough[p] = gr, g_metal[p] = gm;
  }
}

int check_grid(const esr_scene_t *sc) {
  ESR_CHECK_ARG(sc != nullptr);
  ESR_CHECK_ARG(sc->gx > 1 && sc->gy > 1 && sc->gz > 1 && sc->stepdist > 0.f);
  return ESR_OK;
}

}  // namespace

extern "C" int esr_sample_points(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *h_ray,
                                 const int32_t *h_step, int64_t m, float *pts, esr_stream_t stream) {
  if (int e = check_grid(sc)) return e;
  ESR_CHECK_ARG(m >= 0);
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && h_ray && h_step && pts);
  ESR_STAGE("k_sample_points", stream);
  k_sample_points<<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, h_ray, h_step, m, pts);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_sdf_expgrad_fwd(const esr_scene_t *sc, const float *pts, const float *sdf_grid, int64_t m,
                                   int manual, float *out_sdf, float *out_grad, esr_stream_t stream) {
  if (int e = check_grid(sc)) return e;
  ESR_CHECK_ARG(m >= 0);
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(pts && sdf_grid && (out_sdf || out_grad));
  ESR_CHECK_ARG(manual || !out_grad);  // the analytic gradient belongs to the manual sampler only
  if (manual) {
    ESR_STAGE("k_sdf_expgrad_fwd", stream);
    k_sdf_expgrad_fwd<<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(*sc, pts, sdf_grid, m, out_sdf, out_grad);
  } else {
    ESR_STAGE("k_sdf_tap_points", stream);
    k_sdf_tap_points<<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(*sc, pts, sdf_grid, m, out_sdf);
  }
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_sdf_expgrad_bwd(const esr_scene_t *sc, const float *pts, int64_t m, const float *g_sdf,
                                   const float *g_grad, float *grad_sdf_grid, esr_stream_t stream) {
  if (int e = check_grid(sc)) return e;
  ESR_CHECK_ARG(m >= 0);
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(pts && grad_sdf_grid && (g_sdf || g_grad));
  ESR_STAGE("k_sdf_expgrad_bwd", stream);
  k_sdf_expgrad_bwd<<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(*sc, pts, m, g_sdf, g_grad, grad_sdf_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_lts_accumulate_fwd(const float *normal, const float *base, const float *rough, const float *metal,
                                      const float *wo_a, const float *wo_b, const float *dirs, const float *rad_off,
                                      const float *rad_emo, int64_t n_pts, int n_dirs, float *off_hat, float *reflect,
                                      esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && n_dirs > 0);
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(normal && base && rough && metal && wo_a && wo_b && dirs && rad_emo && reflect && (!rad_off == !off_hat));
  ESR_STAGE("k_lts_accumulate_fwd", stream);
  k_lts_accumulate<false><<<cdiv(n_pts * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      normal, base, rough, metal, wo_a, wo_b, dirs, rad_off, rad_emo, n_pts, n_dirs, off_hat, reflect, nullptr, nullptr,
      nullptr, nullptr, nullptr, nullptr, nullptr);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_lts_accumulate_bwd(const float *normal, const float *base, const float *rough, const float *metal,
                                      const float *wo_a, const float *wo_b, const float *dirs, const float *rad_off,
                                      const float *rad_emo, int64_t n_pts, int n_dirs, const float *g_off_hat,
                                      const float *g_reflect, float *g_base, float *g_rough, float *g_metal,
                                      float *g_rad_off, float *g_rad_emo, esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && n_dirs > 0);
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(normal && base && rough && metal && wo_a && wo_b && dirs && rad_emo && g_reflect && g_base && g_rough &&
                g_metal && g_rad_emo);
  ESR_CHECK_ARG(!rad_off == !g_off_hat && !rad_off == !g_rad_off);
  ESR_STAGE("k_lts_accumulate_bwd", stream);
  k_lts_accumulate<true><<<cdiv(n_pts * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      normal, base, rough, metal, wo_a, wo_b, dirs, rad_off, rad_emo, n_pts, n_dirs, nullptr, nullptr, g_off_hat, g_reflect,
      g_base, g_rough, g_metal, g_rad_off, g_rad_emo);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

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
                     float *__restrict__ g_rad_emo, const float *__restrict__ emission, const uint8_t *__restrict__ umask,
                     int pdra, float *__restrict__ g_emission) {
  const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = lane_id();
  if (p >= P) return;
  // emo_hat (esrnerf.py:668-677): emission + reflect; in PDRA mode an UNCERTAIN ray's point gets emission + stop-gradient(
  // reflect), a certain ray's point reflect alone.  With `emission` the `reflect` output is emo_hat and g_reflect its cotangent.
  const bool mix = emission != nullptr;
  const bool um = mix && pdra && umask && __ldg(umask + p) != 0;
  const bool add_emission = mix && (!pdra || um), reflect_has_grad = !(mix && pdra && um);
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
        ge[v][c] = reflect_has_grad ? __ldg(g_reflect + 3 * (v * P + p) + c) * inv : 0.f;
      }
    if (g_emission && lane < 3)   // both outgoing directions repeat the point's emission (emission.repeat(2, 1))
      g_emission[3 * p + lane] = add_emission ? __ldg(g_reflect + 3 * p + lane) + __ldg(g_reflect + 3 * (P + p) + lane) : 0.f;
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
          reflect[3 * (v * P + p) + c] = se * inv + (add_emission ? __ldg(emission + 3 * p + c) : 0.f);
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

// ---------------------------------------------------------------------------------------------
// Hemisphere directions of the light-transport segment (pbr/functions.py:10-32): per LTS point n directions — the
// normalised Gaussian draw `noise` (diffuse_scattering) or the fixed Fibonacci spiral `table` (diffuse_scattering_fib) —
// each mirrored into the hemisphere of the point's normal.  One thread per (point, direction).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_lts_scatter_dirs(const float *__restrict__ normal, const float *__restrict__ noise, const float *__restrict__ table,
                       int64_t P, int n, float *__restrict__ dirs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * n) return;
  const int64_t p = i / n;
  float d[3];
  if (noise) {
    const float x = __ldg(noise + 3 * i), y = __ldg(noise + 3 * i + 1), z = __ldg(noise + 3 * i + 2);
    const float inv = 1.f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);   // F.normalize: v / max(|v|, eps)
    d[0] = x * inv, d[1] = y * inv, d[2] = z * inv;
  } else {
    const int j = (int)(i - p * n);
    d[0] = __ldg(table + 3 * j), d[1] = __ldg(table + 3 * j + 1), d[2] = __ldg(table + 3 * j + 2);
  }
  const float dn = d[0] * __ldg(normal + 3 * p) + d[1] * __ldg(normal + 3 * p + 1) + d[2] * __ldg(normal + 3 * p + 2);
  const float sgn = dn < 0.f ? -1.f : 1.f;
  dirs[3 * i] = sgn * d[0], dirs[3 * i + 1] = sgn * d[1], dirs[3 * i + 2] = sgn * d[2];
}

// ---------------------------------------------------------------------------------------------
// Spherical-Gaussian environment map (pbr/module.py:133-143) on the secondary rays, with what the light-transport
// segment does to it fused in (esrnerf.py:560-566): out = add + act(sum_k mu_k exp(lambda_k (d . l_k - 1))) * scale,
// `scale` = the secondary ray's final transmittance, `add` = its marched off-radiance.  lobes are unit vectors and
// lambdas non-negative (the caller's F.normalize / abs stay in torch: [K, 3] tensors).  act: 1 softplus, 2 relu, 3 abs,
// 4 exp, 5 sigmoid.  Backward: recomputes the lobes' responses, warp- then block-reduces the parameter gradients.
// ---------------------------------------------------------------------------------------------
constexpr int SG_MAX = 64;

ESR_D float sg_act(float x, int act) {
  switch (act) {
    case 1: return x > 20.f ? x : log1pf(expf(x));
    case 2: return fmaxf(x, 0.f);
    case 3: return fabsf(x);
    case 4: return expf(x);
    default: return 1.f / (1.f + expf(-x));
  }
}
ESR_D float sg_act_grad(float x, float y, int act) {   // d act / d x given x and y = act(x)
  switch (act) {
    case 1: return x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
    case 2: return x > 0.f ? 1.f : 0.f;
    case 3: return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
    case 4: return y;
    default: return y * (1.f - y);
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256)
    k_sg_envmap(const float *__restrict__ dirs, const float *__restrict__ mus, const float *__restrict__ lam,
                const float *__restrict__ lobes, int K, int act, const float *__restrict__ scale,
                const float *__restrict__ add, int64_t M, float *__restrict__ out, const float *__restrict__ g_out,
                float *__restrict__ g_mus, float *__restrict__ g_lam, float *__restrict__ g_lobes,
                float *__restrict__ g_scale) {
  __shared__ float s_par[SG_MAX * 7];                       // mu (3), lambda, lobe (3) per lobe
  __shared__ float s_acc[BWD ? SG_MAX * 7 : 1];
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) s_par[7 * i + c] = mus[3 * i + c], s_par[7 * i + 4 + c] = lobes[3 * i + c];
    s_par[7 * i + 3] = lam[i];
  }
  if (BWD)
    for (int i = threadIdx.x; i < 7 * K; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = m < M;
  float d[3] = {0.f, 0.f, 0.f}, pre[3] = {0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = __ldg(dirs + 3 * m + c);
    for (int k = 0; k < K; ++k) {
      const float *q = s_par + 7 * k;
      const float e = expf(q[3] * (d[0] * q[4] + d[1] * q[5] + d[2] * q[6] - 1.f));
#pragma unroll
      for (int c = 0; c < 3; ++c) pre[c] = fmaf(q[c], e, pre[c]);
    }
  }
  const float sc_m = (live && scale) ? __ldg(scale + m) : 1.f;
  if (!BWD) {
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; ++c) out[3 * m + c] = (add ? __ldg(add + 3 * m + c) : 0.f) + sg_act(pre[c], act) * sc_m;
    }
    return;
  }
  float gp[3] = {0.f, 0.f, 0.f};
  if (live) {
    float gs = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float y = sg_act(pre[c], act), g = __ldg(g_out + 3 * m + c);
      gs = fmaf(g, y, gs);
      gp[c] = g * sc_m * sg_act_grad(pre[c], y, act);
    }
    if (g_scale) g_scale[m] = gs;
  }
  const unsigned lane = lane_id();
  for (int k = 0; k < K; ++k) {                              // warp-uniform loop: dead lanes contribute zeros
    const float *q = s_par + 7 * k;
    const float dot = d[0] * q[4] + d[1] * q[5] + d[2] * q[6];
    const float e = live ? expf(q[3] * (dot - 1.f)) : 0.f;
    const float ge = (gp[0] * q[0] + gp[1] * q[1] + gp[2] * q[2]) * e;     // cotangent of the exponent, times e
    float v[7] = {gp[0] * e, gp[1] * e, gp[2] * e, ge * (dot - 1.f), ge * q[3] * d[0], ge * q[3] * d[1], ge * q[3] * d[2]};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float sum = warp_sum(v[i]);
      if (lane == 0) atomicAdd(s_acc + 7 * k + i, sum);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 7 * K; i += blockDim.x) {
    const int k = i / 7, f = i % 7;
    float *dst = f < 3 ? g_mus + 3 * k + f : (f == 3 ? g_lam + k : g_lobes + 3 * k + (f - 4));
    atomicAdd(dst, s_acc[i]);
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
                                      const float *emission, const uint8_t *umask, int pdra_mode, esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && n_dirs > 0);
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(normal && base && rough && metal && wo_a && wo_b && dirs && rad_emo && reflect && (!rad_off == !off_hat));
  ESR_STAGE("k_lts_accumulate_fwd", stream);
  k_lts_accumulate<false><<<cdiv(n_pts * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      normal, base, rough, metal, wo_a, wo_b, dirs, rad_off, rad_emo, n_pts, n_dirs, off_hat, reflect, nullptr, nullptr,
      nullptr, nullptr, nullptr, nullptr, nullptr, emission, umask, pdra_mode, nullptr);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_lts_accumulate_bwd(const float *normal, const float *base, const float *rough, const float *metal,
                                      const float *wo_a, const float *wo_b, const float *dirs, const float *rad_off,
                                      const float *rad_emo, int64_t n_pts, int n_dirs, const float *g_off_hat,
                                      const float *g_reflect, float *g_base, float *g_rough, float *g_metal,
                                      float *g_rad_off, float *g_rad_emo, const float *emission, const uint8_t *umask,
                                      int pdra_mode, float *g_emission, esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && n_dirs > 0);
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(normal && base && rough && metal && wo_a && wo_b && dirs && rad_emo && g_reflect && g_base && g_rough &&
                g_metal && g_rad_emo);
  ESR_CHECK_ARG(!rad_off == !g_off_hat && !rad_off == !g_rad_off);
  ESR_STAGE("k_lts_accumulate_bwd", stream);
  k_lts_accumulate<true><<<cdiv(n_pts * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      normal, base, rough, metal, wo_a, wo_b, dirs, rad_off, rad_emo, n_pts, n_dirs, nullptr, nullptr, g_off_hat, g_reflect,
      g_base, g_rough, g_metal, g_rad_off, g_rad_emo, emission, umask, pdra_mode, g_emission);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_lts_scatter_dirs(const float *normal, const float *noise, const float *table, int64_t n_pts, int n_dirs,
                                    float *dirs, esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && n_dirs > 0);
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(normal && dirs && (!noise != !table));
  ESR_STAGE("k_lts_scatter_dirs", stream);
  k_lts_scatter_dirs<<<cdiv(n_pts * n_dirs, 256), 256, 0, (cudaStream_t)stream>>>(normal, noise, table, n_pts, n_dirs, dirs);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_sg_envmap_fwd(const float *dirs, const float *mus, const float *lambdas, const float *lobes, int n_sg,
                                 int act, const float *scale, const float *add, int64_t m, float *out, esr_stream_t stream) {
  ESR_CHECK_ARG(m >= 0 && n_sg > 0 && n_sg <= SG_MAX && act >= 1 && act <= 5);
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(dirs && mus && lambdas && lobes && out);
  ESR_STAGE("k_sg_envmap_fwd", stream);
  k_sg_envmap<false><<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(dirs, mus, lambdas, lobes, n_sg, act, scale, add, m, out,
                                                                     nullptr, nullptr, nullptr, nullptr, nullptr);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_sg_envmap_bwd(const float *dirs, const float *mus, const float *lambdas, const float *lobes, int n_sg,
                                 int act, const float *scale, int64_t m, const float *g_out, float *g_mus, float *g_lambdas,
                                 float *g_lobes, float *g_scale, esr_stream_t stream) {
  ESR_CHECK_ARG(m >= 0 && n_sg > 0 && n_sg <= SG_MAX && act >= 1 && act <= 5);
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(dirs && mus && lambdas && lobes && g_out && g_mus && g_lambdas && g_lobes && (!scale == !g_scale));
  ESR_STAGE("k_sg_envmap_bwd", stream);
  k_sg_envmap<true><<<cdiv(m, 256), 256, 0, (cudaStream_t)stream>>>(dirs, mus, lambdas, lobes, n_sg, act, scale, nullptr, m,
                                                                    nullptr, g_out, g_mus, g_lambdas, g_lobes, g_scale);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

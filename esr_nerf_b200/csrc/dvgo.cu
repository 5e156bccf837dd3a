// dvgo.cu — the alphamask-stage renderer (app/coarse/model/dvgo.py:140-288): dense [N rays x S steps] sampling with a
// per-ray jitter, post-activated density grid -> alpha, exclusive transmittance product, 3-channel colour grids (no
// MLP), per-ray sums — forward, inference variant and hand-written backward.
//   samples kernels : thread per (ray, step) — point, AABB test, trilinear taps (ATen corner order, zeros padding)
//   scan kernels    : warp per ray — 32 steps at a time, multiplicative warp scan / reverse additive scan
#include "common.cuh"

using namespace esr;

namespace {

ESR_D float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }

struct DvgoPoint {
  float px, py, pz;
  bool outside;
};

// dvgo.py:140-172: t_min from the slab test, p = o + d * (t_min + stepdist * (k + jitter) / |d|); torch evaluates the
// expression with separate roundings (no FMA)
ESR_D DvgoPoint dvgo_point(const esr_dvgo_scene_t &sc, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                           const float *__restrict__ jitter, int64_t r, int k) {
  const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
  const float d[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
  float tmin = -INFINITY, tmax = INFINITY;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float v = d[a] == 0.f ? 1e-6f : d[a];
    const float ra = __fdiv_rn(__fsub_rn(sc.xyz_max[a], o[a]), v), rb = __fdiv_rn(__fsub_rn(sc.xyz_min[a], o[a]), v);
    tmin = fmaxf(tmin, fminf(ra, rb));
    tmax = fminf(tmax, fmaxf(ra, rb));
  }
  tmin = fminf(fmaxf(tmin, sc.near), sc.far);
  tmax = fminf(fmaxf(tmax, sc.near), sc.far);
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  const float rng = __fadd_rn((float)k, jitter ? jitter[r] : 0.f);
  const float t = __fadd_rn(tmin, __fdiv_rn(__fmul_rn(sc.stepdist, rng), nrm));
  DvgoPoint p;
  p.px = __fadd_rn(o[0], __fmul_rn(d[0], t));
  p.py = __fadd_rn(o[1], __fmul_rn(d[1], t));
  p.pz = __fadd_rn(o[2], __fmul_rn(d[2], t));
  p.outside = (tmax <= tmin) | (sc.xyz_min[0] > p.px) | (sc.xyz_min[1] > p.py) | (sc.xyz_min[2] > p.pz) |
              (p.px > sc.xyz_max[0]) | (p.py > sc.xyz_max[1]) | (p.pz > sc.xyz_max[2]);
  return p;
}

ESR_D Cell dvgo_cell(const esr_dvgo_scene_t &sc, const DvgoPoint &p) {
  return make_cell(world_to_index(p.px, sc.xyz_min[0], sc.xyz_max[0], sc.gx), world_to_index(p.py, sc.xyz_min[1], sc.xyz_max[1], sc.gy),
                   world_to_index(p.pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz));
}

ESR_D float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// forward, per sample: alpha (0 outside the box, dvgo.py:185-189) and the two raw colours (dvgo.py:196-201)
__global__ void __launch_bounds__(256)
    k_dvgo_samples_fwd(const __grid_constant__ esr_dvgo_scene_t sc, const float *__restrict__ rays_o,
                       const float *__restrict__ rays_d, const float *__restrict__ jitter, const float *__restrict__ density,
                       const float *__restrict__ off_color, const float *__restrict__ emo_color, int64_t n_rays, int S,
                       float *__restrict__ alpha, float *__restrict__ raw_off, float *__restrict__ raw_emo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * S) return;
  const int64_t r = i / S;
  const int k = (int)(i - r * S);
  const DvgoPoint p = dvgo_point(sc, rays_o, rays_d, jitter, r, k);
  const Cell c = dvgo_cell(sc, p);
  const int64_t vol = (int64_t)sc.gx * sc.gy * sc.gz;
  float a = 0.f;
  if (!p.outside) {
    const float dv = tap1(density, sc.gx, sc.gy, sc.gz, c);
    a = 1.f - expf(-softplusf(dv + sc.act_shift) * sc.interval);
  }
  alpha[i] = a;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    raw_off[3 * i + ch] = sigm(tap1(off_color + ch * vol, sc.gx, sc.gy, sc.gz, c));
    raw_emo[3 * i + ch] = sigm(tap1(emo_color + ch * vol, sc.gx, sc.gy, sc.gz, c));
  }
}

ESR_D float warp_incl_prod(float v, unsigned lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float u = __shfl_up_sync(FULL, v, o);
    if (lane >= (unsigned)o) v *= u;
  }
  return v;
}

// forward, per ray: alphainv_cum = [1, cumprod(clamp_min(1 - alpha, 1e-10))], weights = alpha * alphainv_cum[:-1]
// (dvgo.py:280-288), rgb = sigmoid(off) + [on] sigmoid(emo) and the per-ray sums.  depth (eval): sum w |o - p|.
template <bool EVAL>
__global__ void __launch_bounds__(256)
    k_dvgo_scan_fwd(const __grid_constant__ esr_dvgo_scene_t sc, const float *__restrict__ rays_o,
                    const float *__restrict__ rays_d, const float *__restrict__ jitter, const int64_t *__restrict__ em_modes,
                    const float *__restrict__ alpha, const float *__restrict__ raw_off, const float *__restrict__ raw_emo,
                    int64_t n_rays, int S, float *__restrict__ cum /* [N,S+1] */, float *__restrict__ weights,
                    float *__restrict__ raw_rgb /* train: [N,S,3] */, float *__restrict__ out_a /* [N,3] */,
                    float *__restrict__ out_b, float *__restrict__ out_c, float *__restrict__ depth) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t r = warp; r < n_rays; r += nwarps) {
    const bool on = !EVAL && em_modes[r] == 1;
    float T = 1.f, sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f}, sc3[3] = {0.f, 0.f, 0.f}, sd = 0.f;
    if (lane == 0) cum[r * (S + 1)] = 1.f;
    for (int k0 = 0; k0 < S; k0 += 32) {
      const int k = k0 + (int)lane;
      const bool valid = k < S;
      const int64_t i = r * S + k;
      const float a = valid ? alpha[i] : 0.f;
      const float q = fmaxf(1.f - a, 1e-10f);
      const float incl = warp_incl_prod(valid ? q : 1.f, lane);
      const float excl = __shfl_up_sync(FULL, incl, 1);
      const float Tk = T * (lane ? excl : 1.f);
      const float w = a * Tk;
      if (valid) {
        cum[r * (S + 1) + k + 1] = T * incl;
        weights[i] = w;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float off = raw_off[3 * i + ch], emo = raw_emo[3 * i + ch];
          if (EVAL) {
            sa[ch] = fmaf(w, off, sa[ch]);
            sb[ch] = fmaf(w, emo, sb[ch]);
            sc3[ch] = fmaf(w, off + emo, sc3[ch]);
          } else {
            const float v = on ? off + emo : off;
            raw_rgb[3 * i + ch] = v;
            sa[ch] = fmaf(w, v, sa[ch]);
          }
        }
        if (EVAL) {
          const DvgoPoint p = dvgo_point(sc, rays_o, rays_d, jitter, r, k);
          const float dx = rays_o[3 * r] - p.px, dy = rays_o[3 * r + 1] - p.py, dz = rays_o[3 * r + 2] - p.pz;
          sd = fmaf(w, sqrtf(dx * dx + dy * dy + dz * dz), sd);
        }
      }
      T *= __shfl_sync(FULL, incl, 31);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      sa[ch] = warp_sum(sa[ch]);
      if (EVAL) sb[ch] = warp_sum(sb[ch]), sc3[ch] = warp_sum(sc3[ch]);
    }
    if (EVAL) sd = warp_sum(sd);
    if (lane == 0) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        out_a[3 * r + ch] = sa[ch];
        if (EVAL) out_b[3 * r + ch] = sb[ch], out_c[3 * r + ch] = sc3[ch];
      }
      if (EVAL) depth[r] = sd;
    }
  }
}

// backward, per ray: cotangents of (alphainv_cum, weights, raw_rgb, rgb) -> d_alpha [N,S], d_raw [N,S,3] (total)
//   w_k = a_k T_k,  T_{k+1} = T_k q_k,  q_k = max(1 - a_k, 1e-10)
//   dL/da_k = gw'_k T_k - [1 - a_k > 1e-10] / q_k * sum_{j>k} G_j T_j,   G_j = g_cum_j + gw'_j a_j,  gw'_k = g_w_k + g_rgb . raw_k
__global__ void __launch_bounds__(256)
    k_dvgo_scan_bwd(const float *__restrict__ alpha, const float *__restrict__ cum, const float *__restrict__ raw_rgb,
                    const float *__restrict__ g_cum, const float *__restrict__ g_w, const float *__restrict__ g_raw,
                    const float *__restrict__ g_rgb, int64_t n_rays, int S, float *__restrict__ d_alpha,
                    float *__restrict__ d_raw) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t r = warp; r < n_rays; r += nwarps) {
    const float gr[3] = {g_rgb[3 * r], g_rgb[3 * r + 1], g_rgb[3 * r + 2]};
    float carry = g_cum[r * (S + 1) + S] * cum[r * (S + 1) + S];  // j = S term
    for (int hi = S; hi > 0; hi -= 32) {
      const int k = hi - 1 - (int)lane;
      const bool valid = k >= 0;
      const int64_t i = r * S + k;
      float a = 0.f, Tk = 0.f, gwp = 0.f, x = 0.f;
      if (valid) {
        a = alpha[i];
        Tk = cum[r * (S + 1) + k];
        gwp = g_w[i];
        const float w = a * Tk;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          gwp = fmaf(gr[ch], raw_rgb[3 * i + ch], gwp);
          d_raw[3 * i + ch] = g_raw[3 * i + ch] + w * gr[ch];
        }
        x = (g_cum[r * (S + 1) + k] + gwp * a) * Tk;  // G_k T_k
      }
      float inc = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(FULL, inc, o);
        if (lane >= (unsigned)o) inc += u;
      }
      const float suffix = carry + (inc - x);  // sum_{j>k} G_j T_j
      carry += __shfl_sync(FULL, inc, 31);
      if (valid) {
        const float q = 1.f - a;
        d_alpha[i] = gwp * Tk - (q > 1e-10f ? suffix / q : 0.f);
      }
    }
  }
}

// backward, per sample: d_alpha -> density grid; d_raw -> colour grids through the sigmoids
__global__ void __launch_bounds__(256)
    k_dvgo_samples_bwd(const __grid_constant__ esr_dvgo_scene_t sc, const float *__restrict__ rays_o,
                       const float *__restrict__ rays_d, const float *__restrict__ jitter, const int64_t *__restrict__ em_modes,
                       const float *__restrict__ density, const float *__restrict__ raw_off, const float *__restrict__ raw_emo,
                       const float *__restrict__ alpha, const float *__restrict__ d_alpha, const float *__restrict__ d_raw,
                       int64_t n_rays, int S, float *__restrict__ g_density, float *__restrict__ g_off,
                       float *__restrict__ g_emo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * S) return;
  const int64_t r = i / S;
  const int k = (int)(i - r * S);
  const DvgoPoint p = dvgo_point(sc, rays_o, rays_d, jitter, r, k);
  const Cell c = dvgo_cell(sc, p);
  const int64_t vol = (int64_t)sc.gx * sc.gy * sc.gz;
  if (!p.outside) {
    const float da = d_alpha[i];
    if (da != 0.f) {
      const float dv = tap1(density, sc.gx, sc.gy, sc.gz, c);
      // alpha = 1 - exp(-softplus(x) I): d alpha / d x = (1 - alpha) I sigmoid(x)
      const float g = da * (1.f - alpha[i]) * sc.interval * sigm(dv + sc.act_shift);
      if (g != 0.f) scatter1(g_density, sc.gx, sc.gy, sc.gz, c, g);
    }
  }
  const bool on = em_modes[r] == 1;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float d = d_raw[3 * i + ch];
    if (d == 0.f) continue;
    const float so = raw_off[3 * i + ch];
    scatter1(g_off + ch * vol, sc.gx, sc.gy, sc.gz, c, d * so * (1.f - so));
    if (on) {
      const float se = raw_emo[3 * i + ch];
      scatter1(g_emo + ch * vol, sc.gx, sc.gy, sc.gz, c, d * se * (1.f - se));
    }
  }
}

int check_dvgo(const esr_dvgo_scene_t *sc, int64_t n_rays, int S) {
  ESR_CHECK_ARG(sc != nullptr && sc->gx > 1 && sc->gy > 1 && sc->gz > 1 && sc->stepdist > 0.f);
  ESR_CHECK_ARG(n_rays >= 0 && S > 0 && n_rays * (int64_t)S < (1ll << 40));
  return ESR_OK;
}

unsigned ray_grid(int64_t n_rays) {
  const int64_t want = (n_rays + 7) / 8, cap = (int64_t)num_sms() * 32;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

extern "C" int esr_dvgo_fwd(const esr_dvgo_scene_t *sc, const float *rays_o, const float *rays_d, const float *jitter,
                            const int64_t *em_modes, const float *density, const float *off_color,
                            const float *emo_color, int64_t n_rays, int n_samples, float *alpha, float *raw_off,
                            float *raw_emo, float *alphainv_cum, float *weights, float *raw_rgb, float *rgb,
                            esr_stream_t stream) {
  if (int e = check_dvgo(sc, n_rays, n_samples)) return e;
  if (n_rays == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && em_modes && density && off_color && emo_color && alpha && raw_off && raw_emo &&
                alphainv_cum && weights && raw_rgb && rgb);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_dvgo_samples_fwd", st);
  k_dvgo_samples_fwd<<<cdiv(n_rays * n_samples, 256), 256, 0, st>>>(*sc, rays_o, rays_d, jitter, density, off_color,
                                                                   emo_color, n_rays, n_samples, alpha, raw_off, raw_emo);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_dvgo_scan_fwd", st);
  k_dvgo_scan_fwd<false><<<ray_grid(n_rays), 256, 0, st>>>(*sc, rays_o, rays_d, jitter, em_modes, alpha, raw_off, raw_emo,
                                                           n_rays, n_samples, alphainv_cum, weights, raw_rgb, rgb, nullptr,
                                                           nullptr, nullptr);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_dvgo_eval(const esr_dvgo_scene_t *sc, const float *rays_o, const float *rays_d, const float *density,
                             const float *off_color, const float *emo_color, int64_t n_rays, int n_samples, float *alpha,
                             float *raw_off, float *raw_emo, float *alphainv_cum, float *weights, float *off_rgb,
                             float *emo_rgb, float *on_rgb, float *depth, esr_stream_t stream) {
  if (int e = check_dvgo(sc, n_rays, n_samples)) return e;
  if (n_rays == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && density && off_color && emo_color && alpha && raw_off && raw_emo && alphainv_cum &&
                weights && off_rgb && emo_rgb && on_rgb && depth);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_dvgo_samples_fwd", st);
  k_dvgo_samples_fwd<<<cdiv(n_rays * n_samples, 256), 256, 0, st>>>(*sc, rays_o, rays_d, nullptr, density, off_color,
                                                                   emo_color, n_rays, n_samples, alpha, raw_off, raw_emo);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_dvgo_scan_eval", st);
  k_dvgo_scan_fwd<true><<<ray_grid(n_rays), 256, 0, st>>>(*sc, rays_o, rays_d, nullptr, nullptr, alpha, raw_off, raw_emo,
                                                          n_rays, n_samples, alphainv_cum, weights, nullptr, off_rgb, emo_rgb,
                                                          on_rgb, depth);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_dvgo_bwd(const esr_dvgo_scene_t *sc, const float *rays_o, const float *rays_d, const float *jitter,
                            const int64_t *em_modes, const float *density, int64_t n_rays, int n_samples,
                            const float *alpha, const float *raw_off, const float *raw_emo, const float *alphainv_cum,
                            const float *raw_rgb, const float *g_cum, const float *g_weights, const float *g_raw_rgb,
                            const float *g_rgb, float *d_alpha, float *d_raw, float *grad_density, float *grad_off_color,
                            float *grad_emo_color, esr_stream_t stream) {
  if (int e = check_dvgo(sc, n_rays, n_samples)) return e;
  if (n_rays == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && em_modes && density && alpha && raw_off && raw_emo && alphainv_cum && raw_rgb &&
                g_cum && g_weights && g_raw_rgb && g_rgb && d_alpha && d_raw && grad_density && grad_off_color &&
                grad_emo_color);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_dvgo_scan_bwd", st);
  k_dvgo_scan_bwd<<<ray_grid(n_rays), 256, 0, st>>>(alpha, alphainv_cum, raw_rgb, g_cum, g_weights, g_raw_rgb, g_rgb, n_rays,
                                                    n_samples, d_alpha, d_raw);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_dvgo_samples_bwd", st);
  k_dvgo_samples_bwd<<<cdiv(n_rays * n_samples, 256), 256, 0, st>>>(*sc, rays_o, rays_d, jitter, em_modes, density, raw_off,
                                                                   raw_emo, alpha, d_alpha, d_raw, n_rays, n_samples,
                                                                   grad_density, grad_off_color, grad_emo_color);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
 *
 * Plain-C, single-threaded CPU restatement of the three live native ops of the
 * reference (app/utils/base/cuda/render_utils_kernel.cu) plus the semantics of
 * torch_scatter.segment_coo(reduce="sum") that the reference relies on.
 *
 * Parity pin: this file is checked (tests/test_oracle_pin.py, -m gpu) against
 * the reference's OWN kernels, compiled unmodified-in-place from
 * /root/reference into oracle/_ref/ by oracle/build_ref.py, on identical
 * inputs; and (tests/test_oracle_cpu.py) against tests/golden/ vectors that
 * were produced by the reference's own Python render path driving this file.
 * torch_scatter is an un-vendored, un-pinned third-party dependency of the
 * reference (README.md:14) -> segment_coo parity is "unpinned" (semantics only:
 * sum over a sorted index into a zero-initialised `out`).
 *
 * Floating-point contraction: the reference is compiled by nvcc with the
 * default -fmad=true.  The places where nvcc 12.9 contracts a*b+c into one
 * fma.rn.f32 were read off the PTX of the reference expressions and are spelt
 * out with fmaf() below; everything else is compiled -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* render_utils_kernel.cu:12-35  infer_t_minmax_cuda_kernel */
static void ref_t_minmax(const float *o, const float *d, const float *xyz_min,
                         const float *xyz_max, float near, float far,
                         float *t_min, float *t_max) {
  /* `1e-6` is a double literal narrowed to float (kernel.cu:23-25) */
  float vx = (d[0] == 0) ? (float)1e-6 : d[0];
  float vy = (d[1] == 0) ? (float)1e-6 : d[1];
  float vz = (d[2] == 0) ? (float)1e-6 : d[2];
  float ax = (xyz_max[0] - o[0]) / vx;
  float ay = (xyz_max[1] - o[1]) / vy;
  float az = (xyz_max[2] - o[2]) / vz;
  float bx = (xyz_min[0] - o[0]) / vx;
  float by = (xyz_min[1] - o[1]) / vy;
  float bz = (xyz_min[2] - o[2]) / vz;
  *t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), near);
  *t_max = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), near);
}

/* kernel.cu:48-51 / 68-71.  nvcc: mul(d1,d1) -> fma(d0,d0,.) -> fma(d2,d2,.) */
static float ref_rnorm(const float *d) {
  float s = d[1] * d[1];
  s = fmaf(d[0], d[0], s);
  s = fmaf(d[2], d[2], s);
  return sqrtf(s);
}

/*
 * render_utils_kernel.cu:196-242  sample_pts_on_rays_cuda.
 * Two-call protocol: pass ray_pts == NULL to obtain total_len (the reference
 * does the same with a host sync, kernel.cu:212).
 * Outputs (any may be NULL): ray_pts[total,3] f32, mask_outbbox[total] u8,
 * ray_id[total] i64, step_id[total] i64, N_steps[n] i64, t_min[n], t_max[n].
 */
int64_t oracle_sample_pts_on_rays(const float *rays_o, const float *rays_d,
                                  const float *xyz_min, const float *xyz_max,
                                  float near, float far, float stepdist,
                                  int64_t n_rays, float *ray_pts,
                                  uint8_t *mask_outbbox, int64_t *ray_id,
                                  int64_t *step_id, int64_t *N_steps,
                                  float *t_min_out, float *t_max_out) {
  int64_t total = 0;
  for (int64_t r = 0; r < n_rays; ++r) {
    const float *o = rays_o + 3 * r, *d = rays_d + 3 * r;
    float t_min, t_max;
    ref_t_minmax(o, d, xyz_min, xyz_max, near, far, &t_min, &t_max);
    const float rnorm = ref_rnorm(d);
    /* kernel.cu:53 — float arithmetic, ceilf, then max(float, double 1.) */
    double nd = fmax((double)ceilf((t_max - t_min) * rnorm / stepdist), 1.);
    int64_t n = (int64_t)nd;
    if (N_steps) N_steps[r] = n;
    if (t_min_out) t_min_out[r] = t_min;
    if (t_max_out) t_max_out[r] = t_max;
    if (ray_pts || mask_outbbox || ray_id || step_id) {
      /* kernel.cu:72-77 — start = fma(d, t_min, o); dir = d / rnorm */
      float sx = fmaf(d[0], t_min, o[0]);
      float sy = fmaf(d[1], t_min, o[1]);
      float sz = fmaf(d[2], t_min, o[2]);
      float dx = d[0] / rnorm, dy = d[1] / rnorm, dz = d[2] / rnorm;
      for (int64_t k = 0; k < n; ++k) {
        int64_t idx = total + k;
        /* kernel.cu:179-187 — i_step is read into an int; dist = stepdist*i_step */
        const float dist = stepdist * (float)(int)k;
        float px = fmaf(dx, dist, sx);
        float py = fmaf(dy, dist, sy);
        float pz = fmaf(dz, dist, sz);
        if (ray_pts) {
          ray_pts[3 * idx] = px;
          ray_pts[3 * idx + 1] = py;
          ray_pts[3 * idx + 2] = pz;
        }
        if (mask_outbbox)
          mask_outbbox[idx] = (xyz_min[0] > px) | (xyz_min[1] > py) | (xyz_min[2] > pz) |
                              (xyz_max[0] < px) | (xyz_max[1] < py) | (xyz_max[2] < pz);
        if (ray_id) ray_id[idx] = r;
        if (step_id) step_id[idx] = k;
      }
    }
    total += n;
  }
  return total;
}

/*
 * render_utils_kernel.cu:576-651  alpha2weight_cuda (+ __set_i_for_segment_start_end).
 * weight is zero-initialised, T one-initialised, alphainv_last one-initialised,
 * i_start / i_end zero-initialised (kernel.cu:624-628); rays without samples
 * keep i_start == i_end == 0.
 */
void oracle_alpha2weight(const float *alpha, const int64_t *ray_id, int64_t n_pts,
                         int64_t n_rays, float *weight, float *T,
                         float *alphainv_last, int64_t *i_start, int64_t *i_end) {
  for (int64_t i = 0; i < n_pts; ++i) {
    weight[i] = 0.f;
    T[i] = 1.f;
  }
  for (int64_t r = 0; r < n_rays; ++r) {
    alphainv_last[r] = 1.f;
    i_start[r] = 0;
    i_end[r] = 0;
  }
  if (n_pts == 0) return;
  for (int64_t i = 1; i < n_pts; ++i) {
    if (ray_id[i] != ray_id[i - 1]) {
      i_start[ray_id[i]] = i;
      i_end[ray_id[i - 1]] = i;
    }
  }
  i_end[ray_id[n_pts - 1]] = n_pts;
  for (int64_t r = 0; r < n_rays; ++r) {
    const int64_t i_s = i_start[r], i_e_max = i_end[r];
    float T_cum = 1.f;
    int64_t i;
    for (i = i_s; i < i_e_max; ++i) {
      T[i] = T_cum;
      weight[i] = T_cum * alpha[i];
      /* kernel.cu:596 — (1. - alpha) is double; product rounded back to float */
      T_cum = (float)((1. - (double)alpha[i]) * (double)T_cum);
      if ((double)T_cum < 1e-3) {
        i += 1;
        break;
      }
    }
    i_end[r] = i;
    alphainv_last[r] = T_cum;
  }
}

/* render_utils_kernel.cu:653-707  alpha2weight_backward_cuda */
void oracle_alpha2weight_backward(const float *alpha, const float *weight,
                                  const float *T, const float *alphainv_last,
                                  const int64_t *i_start, const int64_t *i_end,
                                  int64_t n_pts, int64_t n_rays,
                                  const float *grad_weights, const float *grad_last,
                                  float *grad) {
  for (int64_t i = 0; i < n_pts; ++i) grad[i] = 0.f;
  for (int64_t r = 0; r < n_rays; ++r) {
    float back_cum = grad_last[r] * alphainv_last[r];
    for (int64_t i = i_end[r] - 1; i >= i_start[r]; --i) {
      /* kernel.cu:673 — (1-alpha) float, +1e-10 double, division and subtraction in double */
      float gwT = grad_weights[i] * T[i];
      double den = (double)(1 - alpha[i]) + 1e-10;
      grad[i] = (float)((double)gwT - (double)back_cum / den);
      back_cum = fmaf(grad_weights[i], weight[i], back_cum);
    }
  }
}

/*
 * torch_scatter.segment_coo(src, index, out=zeros[N,C], reduce="sum") as the
 * reference uses it (e.g. voxurff.py:260-272): out[index[i], :] += src[i, :]
 * for a sorted `index`.  Summation order: ascending i (one legal order; the
 * third-party kernel's order is unspecified -> compare with tolerance).
 */
void oracle_segment_coo_sum(const float *src, const int64_t *index, int64_t n_pts,
                            int64_t channels, float *out /* [n_out, channels], pre-zeroed by caller */) {
  for (int64_t i = 0; i < n_pts; ++i)
    for (int64_t c = 0; c < channels; ++c) out[index[i] * channels + c] += src[i * channels + c];
}

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

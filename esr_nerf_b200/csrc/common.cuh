// common.cuh — shared helpers for libesr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/esr_b200.h"

#define ESR_HD __host__ __device__ __forceinline__
#define ESR_D __device__ __forceinline__

namespace esr {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

// Per-kernel device timing for bench.py's roofline block: when enabled through esr_stage_timing(1) every
// launch site brackets its kernel with a pair of CUDA events on the launching stream (ESR_STAGE before the
// launch, ESR_LAUNCH_OK after it); esr_stage_timing_report sums the elapsed times per stage name.
// Disabled (the default) it costs one branch per launch.
void stage_begin(const char *name, cudaStream_t st);
void stage_end();
#define ESR_STAGE(name, stream) esr::stage_begin(name, (cudaStream_t)(stream))

#define ESR_CHECK_ARG(cond)                                                  \
  do {                                                                       \
    if (!(cond)) {                                                           \
      esr::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);  \
      return ESR_ERR_BAD_ARG;                                                \
    }                                                                        \
  } while (0)

#define ESR_CHECK_CUDA(expr)                                                                 \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      esr::set_error("%s:%d: CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),   \
                     cudaGetErrorString(_e));                                                \
      return ESR_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

// check the launch that was just issued (sticky-free peek) and count it
#define ESR_LAUNCH_OK()                  \
  do {                                   \
    esr::stage_end();                    \
    esr::count_launch();                 \
    ESR_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

static inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

constexpr unsigned FULL = 0xffffffffu;

ESR_D unsigned lane_id() { return threadIdx.x & 31; }
ESR_D unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
ESR_D float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// fire-and-forget float add into global memory (RED, no return)
ESR_D void red_add(float *addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// 8-byte vector RED (sm_90+); addr must be 8-byte aligned
ESR_D void red_add2(float *addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// 16-byte vector RED (sm_90+); addr must be 16-byte aligned
ESR_D void red_add4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Geometry shared by all render kernels
// ---------------------------------------------------------------------------------------------
struct RaySetup {
  float t_min, t_max;
  float sx, sy, sz;  // start = o + d * t_min
  float dx, dy, dz;  // unit direction
  int n;             // number of candidate steps
};

// render_utils_kernel.cu:12-79.  Expression shapes (and therefore nvcc's FMA contraction) are
// spelt out with explicit intrinsics so the integer outputs match the reference bit for bit:
//   rnorm^2 = fma(d2,d2, fma(d0,d0, d1*d1));  start = fma(d, t_min, o);  p = fma(dir, dist, start)
ESR_D RaySetup ray_setup(const float *__restrict__ rays_o, const float *__restrict__ rays_d, int64_t r,
                         const float *mn, const float *mx, float near, float far, float stepdist) {
  RaySetup s;
  const float ox = __ldg(rays_o + 3 * r), oy = __ldg(rays_o + 3 * r + 1), oz = __ldg(rays_o + 3 * r + 2);
  const float d0 = __ldg(rays_d + 3 * r), d1 = __ldg(rays_d + 3 * r + 1), d2 = __ldg(rays_d + 3 * r + 2);
  const float vx = (d0 == 0.f) ? (float)1e-6 : d0;
  const float vy = (d1 == 0.f) ? (float)1e-6 : d1;
  const float vz = (d2 == 0.f) ? (float)1e-6 : d2;
  const float ax = __fdiv_rn(__fsub_rn(mx[0], ox), vx);
  const float ay = __fdiv_rn(__fsub_rn(mx[1], oy), vy);
  const float az = __fdiv_rn(__fsub_rn(mx[2], oz), vz);
  const float bx = __fdiv_rn(__fsub_rn(mn[0], ox), vx);
  const float by = __fdiv_rn(__fsub_rn(mn[1], oy), vy);
  const float bz = __fdiv_rn(__fsub_rn(mn[2], oz), vz);
  s.t_min = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), near);
  s.t_max = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), near);
  const float rn2 = __fmaf_rn(d2, d2, __fmaf_rn(d0, d0, __fmul_rn(d1, d1)));
  const float rnorm = __fsqrt_rn(rn2);
  const float c = ceilf(__fdiv_rn(__fmul_rn(rnorm, __fsub_rn(s.t_max, s.t_min)), stepdist));
  const double nd = fmax((double)c, 1.);
  s.n = (int)(long long)nd;
  s.sx = __fmaf_rn(d0, s.t_min, ox);
  s.sy = __fmaf_rn(d1, s.t_min, oy);
  s.sz = __fmaf_rn(d2, s.t_min, oz);
  s.dx = __fdiv_rn(d0, rnorm);
  s.dy = __fdiv_rn(d1, rnorm);
  s.dz = __fdiv_rn(d2, rnorm);
  return s;
}

// kernel.cu:184-187
ESR_D void ray_point(const RaySetup &s, float stepdist, int k, float &px, float &py, float &pz) {
  const float dist = __fmul_rn(stepdist, (float)k);
  px = __fmaf_rn(s.dx, dist, s.sx);
  py = __fmaf_rn(s.dy, dist, s.sy);
  pz = __fmaf_rn(s.dz, dist, s.sz);
}

// kernel.cu:191-192
ESR_D bool out_bbox(const float *mn, const float *mx, float px, float py, float pz) {
  return (mn[0] > px) | (mn[1] > py) | (mn[2] > pz) | (mx[0] < px) | (mx[1] < py) | (mx[2] < pz);
}

// world coordinate -> continuous grid index exactly as the reference's torch expression
// ((p-min)/(max-min))*2-1 followed by ATen's align_corners un-normalisation ((c+1)/2)*(size-1)
// (module.py:27-31, ATen GridSampler.cuh grid_sampler_unnormalize).  Separate roundings, no FMA.
ESR_D float world_to_index(float p, float mn, float mx, int size) {
  const float u = __fdiv_rn(__fsub_rn(p, mn), __fsub_rn(mx, mn));
  const float c = __fsub_rn(__fmul_rn(u, 2.f), 1.f);
  return __fmul_rn(__fdiv_rn(__fadd_rn(c, 1.f), 2.f), (float)(size - 1));
}

// Trilinear cell: base corner + the 8 ATen weights in ATen's accumulation order
// (tnw,tne,tsw,tse,bnw,bne,bsw,bse) = (x0y0z0, x0y0z1, x0y1z0, x0y1z1, x1y0z0, ...), where
// x indexes dim 2 (slowest), z dim 4 (fastest) of a [1,C,X,Y,Z] grid.
struct Cell {
  int x0, y0, z0;
  float w[8];
};

ESR_D Cell make_cell(float ix, float iy, float iz) {
  // ix: index along X (ATen "iz"/depth), iy along Y ("iy"), iz along Z (ATen "ix"/width)
  Cell c;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  c.x0 = (int)fx;
  c.y0 = (int)fy;
  c.z0 = (int)fz;
  const float z_lo = __fsub_rn((float)(c.z0 + 1), iz), z_hi = __fsub_rn(iz, (float)c.z0);
  const float y_lo = __fsub_rn((float)(c.y0 + 1), iy), y_hi = __fsub_rn(iy, (float)c.y0);
  const float x_lo = __fsub_rn((float)(c.x0 + 1), ix), x_hi = __fsub_rn(ix, (float)c.x0);
  // ATen: (ix term) * (iy term) * (iz term) with ATen-ix = our z, ATen-iz = our x
  c.w[0] = __fmul_rn(__fmul_rn(z_lo, y_lo), x_lo);
  c.w[1] = __fmul_rn(__fmul_rn(z_hi, y_lo), x_lo);
  c.w[2] = __fmul_rn(__fmul_rn(z_lo, y_hi), x_lo);
  c.w[3] = __fmul_rn(__fmul_rn(z_hi, y_hi), x_lo);
  c.w[4] = __fmul_rn(__fmul_rn(z_lo, y_lo), x_hi);
  c.w[5] = __fmul_rn(__fmul_rn(z_hi, y_lo), x_hi);
  c.w[6] = __fmul_rn(__fmul_rn(z_lo, y_hi), x_hi);
  c.w[7] = __fmul_rn(__fmul_rn(z_hi, y_hi), x_hi);
  return c;
}

ESR_D bool in_grid(int x, int y, int z, int X, int Y, int Z) {
  return (unsigned)x < (unsigned)X && (unsigned)y < (unsigned)Y && (unsigned)z < (unsigned)Z;
}

// scalar-channel trilinear tap with zeros padding; accumulation order and FMA shape of ATen's
// CUDA grid_sampler_3d (out_acc += inp * w).
ESR_D float tap1(const float *__restrict__ g, int X, int Y, int Z, const Cell &c) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = c.x0 + (k >> 2), y = c.y0 + ((k >> 1) & 1), z = c.z0 + (k & 1);
    if (in_grid(x, y, z, X, Y, Z)) acc = __fmaf_rn(__ldg(g + ((int64_t)x * Y + y) * Z + z), c.w[k], acc);
  }
  return acc;
}

ESR_D float tap1_world(const float *__restrict__ g, int X, int Y, int Z, const float *mn, const float *mx,
                       float px, float py, float pz) {
  const Cell c = make_cell(world_to_index(px, mn[0], mx[0], X), world_to_index(py, mn[1], mx[1], Y),
                           world_to_index(pz, mn[2], mx[2], Z));
  return tap1(g, X, Y, Z, c);
}

// differentiable_grid_sample arithmetic (functions.py:142-309, reached through esrnerf.py:1572-1596): the same 8
// weights, corner indices CLAMPED to the grid instead of zero padding, and value * weight products summed with
// separately rounded torch ops (no FMA), in the order tnw .. bse.
ESR_D int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }
ESR_D float tap1_manual(const float *__restrict__ g, int X, int Y, int Z, const Cell &c) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = clampi(c.x0 + (k >> 2), X - 1), y = clampi(c.y0 + ((k >> 1) & 1), Y - 1), z = clampi(c.z0 + (k & 1), Z - 1);
    const float term = __fmul_rn(__ldg(g + ((int64_t)x * Y + y) * Z + z), c.w[k]);
    acc = k == 0 ? term : __fadd_rn(acc, term);
  }
  return acc;
}
ESR_D float tap1_manual_world(const float *__restrict__ g, int X, int Y, int Z, const float *mn, const float *mx,
                              float px, float py, float pz) {
  const Cell c = make_cell(world_to_index(px, mn[0], mx[0], X), world_to_index(py, mn[1], mx[1], Y),
                           world_to_index(pz, mn[2], mx[2], Z));
  return tap1_manual(g, X, Y, Z, c);
}

// scatter-add v * w[k] into a scalar-channel gradient volume
ESR_D void scatter1(float *__restrict__ g, int X, int Y, int Z, const Cell &c, float v) {
  // the two corners of a (x, y) pair are adjacent along Z: one 8-byte RED when both are in the grid and the pair is
  // 8-byte aligned (Z even, z0 even), two scalar REDs otherwise — the L2 pays per RED request, not per byte
  const bool pair_ok = ((Z | c.z0) & 1) == 0 && (unsigned)c.z0 < (unsigned)(Z - 1);
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const int x = c.x0 + (k >> 2), y = c.y0 + ((k >> 1) & 1);
    if ((unsigned)x >= (unsigned)X || (unsigned)y >= (unsigned)Y) continue;
    float *p = g + ((int64_t)x * Y + y) * Z + c.z0;
    if (pair_ok) {
      red_add2(p, v * c.w[k], v * c.w[k + 1]);
    } else {
      if ((unsigned)c.z0 < (unsigned)Z) red_add(p, v * c.w[k]);
      if ((unsigned)(c.z0 + 1) < (unsigned)Z) red_add(p + 1, v * c.w[k + 1]);
    }
  }
}

// MaskCache.forward (module.py:104-114): trilinear on the max-pooled density, softplus, 1-exp, >= thres
ESR_D bool mask_keep(const esr_scene_t &sc, const float *__restrict__ mask_density, float px, float py,
                     float pz) {
  const float d = tap1_world(mask_density, sc.mx, sc.my, sc.mz, sc.mask_xyz_min, sc.mask_xyz_max, px, py, pz);
  const float x = __fadd_rn(d, sc.act_shift);
  const float sp = (x > 20.f) ? x : log1pf(expf(x));  // F.softplus(beta=1, threshold=20)
  const float a = __fsub_rn(1.f, expf(-sp));
  return a >= sc.mask_thres;
}

ESR_D float sigmoidf(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// NeuS 'interp' alpha (functions.py:72-105) for one sample given its stream neighbours
ESR_D float neus_alpha(float sd, float sd_prev, float sd_next, bool has_prev, bool has_next, float s_val,
                       float &pc, float &nc) {
  const float next_est = has_next ? __fmul_rn(__fadd_rn(sd, sd_next), 0.5f) : sd;
  const float prev_est = has_prev ? __fmul_rn(__fadd_rn(sd_prev, sd), 0.5f) : sd;
  pc = sigmoidf(__fmul_rn(prev_est, s_val));
  nc = sigmoidf(__fmul_rn(next_est, s_val));
  const float p = fmaxf(__fsub_rn(pc, nc), 0.f);
  const float r = __fdiv_rn(__fadd_rn(p, 1e-5f), __fadd_rn(pc, 1e-5f));
  return fminf(fmaxf(r, 0.f), 1.f);
}

// NeuS 'grad' alpha (functions.py:45-69): section-point SDFs estimated along the view direction,
// sdf -+ iter_cos with iter_cos = (v . grad sdf) * dist * 0.5 (computed per sample by k_neus_cos_fwd)
ESR_D float neus_alpha_grad(float sd, float iter_cos, float s_val, float &pc, float &nc) {
  const float next_est = __fadd_rn(sd, iter_cos);
  const float prev_est = __fsub_rn(sd, iter_cos);
  pc = sigmoidf(__fmul_rn(prev_est, s_val));
  nc = sigmoidf(__fmul_rn(next_est, s_val));
  const float p = fmaxf(__fsub_rn(pc, nc), 0.f);
  const float r = __fdiv_rn(__fadd_rn(p, 1e-5f), __fadd_rn(pc, 1e-5f));
  return fminf(fmaxf(r, 0.f), 1.f);
}

}  // namespace esr

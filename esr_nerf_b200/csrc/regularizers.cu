// regularizers.cu — the dense-grid loss terms the stage drivers add every `tv_every` steps:
//
//   total_variation(v, mask)           app/utils/base/functions.py:34-42   (voxurff.py:603-609, voxurfc.py:523-548)
//   smooth-gradient term               voxurff.py:610-616 with neus_sdf_gradient (voxurff.py:723-742) and the fixed
//                                      3x3x3 kernel of GradientConv (module.py:180-211, replicate padding, detached)
//
// The reference runs them as ~25 dense torch ops (diff / abs / boolean-mask gathers / means; zeros + three sliced
// assignments; permute + cuDNN conv3d + repeat + gather + square + mean) over the whole volume plus their autograd
// graph.  Here each term is one forward launch (value: per-axis sums and pair counts, reduced in double) and one
// backward launch (gather form: every voxel collects the contributions of the pairs / stencils it belongs to — no atomics
// on the gradient volume), reading the grid through explicit strides so that the channels-last colour grids need no copy.
#include "common.cuh"

using namespace esr;

namespace {

struct Vol {
  int C, X, Y, Z;
  int64_t sc, sx, sy, sz;   // element strides of the [C][X][Y][Z] view
  __device__ int64_t at(int c, int x, int y, int z) const { return c * sc + x * sx + y * sy + z * sz; }
};

ESR_D double block_sum(double v, double *sh) {   // 256 threads; result valid on thread 0
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 8) s = sh[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  }
  __syncthreads();
  return s;
}

// acc[a] += sum over masked pairs along axis a of |v[p + e_a] - v[p]| (all channels), acc[3 + a] += their number
__global__ void __launch_bounds__(256)
    k_tv_fwd(Vol V, const float *__restrict__ v, const uint8_t *__restrict__ mask, double *__restrict__ acc) {
  __shared__ double sh[8];
  const int64_t n = (int64_t)V.X * V.Y * V.Z;
  double s[3] = {0.0, 0.0, 0.0}, cnt[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % V.Z), y = (int)((i / V.Z) % V.Y), x = (int)(i / ((int64_t)V.Z * V.Y));
    if (mask && !mask[i]) continue;
    const int64_t step[3] = {(int64_t)V.Y * V.Z, V.Z, 1};
    const bool in[3] = {x + 1 < V.X, y + 1 < V.Y, z + 1 < V.Z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!in[a] || (mask && !mask[i + step[a]])) continue;
      const int64_t d = a == 0 ? V.sx : (a == 1 ? V.sy : V.sz);
      float t = 0.f;
      for (int c = 0; c < V.C; ++c) {
        const int64_t p = V.at(c, x, y, z);
        t += fabsf(__ldg(v + p + d) - __ldg(v + p));
      }
      s[a] += (double)t;
      cnt[a] += (double)V.C;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double ss = block_sum(s[a], sh), cc = block_sum(cnt[a], sh);
    if (threadIdx.x == 0) {
      atomicAdd(acc + a, ss);
      atomicAdd(acc + 3 + a, cc);
    }
  }
}

// grad[p] += g * sum_a (1 / (3 cnt_a)) * (sign(v[p] - v[p - e_a]) [pair (p - e_a, p)] - sign(v[p + e_a] - v[p]) [pair (p, p + e_a)])
__global__ void __launch_bounds__(256)
    k_tv_bwd(Vol V, const float *__restrict__ v, const uint8_t *__restrict__ mask, const double *__restrict__ acc,
             const float *__restrict__ g_out, float scale, float *__restrict__ grad) {
  const int64_t n = (int64_t)V.X * V.Y * V.Z;
  const float g = __ldg(g_out) * scale;
  float w[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) w[a] = (float)((double)g / (3.0 * acc[3 + a]));   // (an axis with no pair: x / 0, as the mean of an empty selection)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;
    const int z = (int)(i % V.Z), y = (int)((i / V.Z) % V.Y), x = (int)(i / ((int64_t)V.Z * V.Y));
    const int64_t step[3] = {(int64_t)V.Y * V.Z, V.Z, 1};
    const bool hi[3] = {x + 1 < V.X, y + 1 < V.Y, z + 1 < V.Z}, lo[3] = {x > 0, y > 0, z > 0};
    for (int c = 0; c < V.C; ++c) {
      const int64_t p = V.at(c, x, y, z);
      const float vp = __ldg(v + p);
      float t = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int64_t d = a == 0 ? V.sx : (a == 1 ? V.sy : V.sz);
        if (lo[a] && (!mask || mask[i - step[a]])) {
          const float df = vp - __ldg(v + p - d);
          t += w[a] * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
        }
        if (hi[a] && (!mask || mask[i + step[a]])) {
          const float df = __ldg(v + p + d) - vp;
          t -= w[a] * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
        }
      }
      grad[p] += t;
    }
  }
}

// neus_sdf_gradient (voxurff.py:723-742): central differences / (2 h), zero on the two boundary faces of each axis;
// out [3][X][Y][Z] contiguous
__global__ void __launch_bounds__(256)
    k_sdf_central_gradient(const float *__restrict__ sdf, int X, int Y, int Z, float voxel_size, float *__restrict__ out) {
  const int64_t n = (int64_t)X * Y * Z;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = (int)(i % Z), y = (int)((i / Z) % Y), x = (int)(i / ((int64_t)Z * Y));
  // (a - b) / 2 / h as the reference divides: two IEEE divisions
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float d = 0.f;
    const int64_t st = a == 0 ? (int64_t)Y * Z : (a == 1 ? Z : 1);
    const int c = a == 0 ? x : (a == 1 ? y : z), lim = a == 0 ? X : (a == 1 ? Y : Z);
    if (c > 0 && c + 1 < lim) d = __fdiv_rn(__fdiv_rn(__ldg(sdf + i + st) - __ldg(sdf + i - st), 2.f), voxel_size);
    out[a * n + i] = d;
  }
}

// err_c(p) = conv3(grad_c)(p) - grad_c(p) on masked voxels (replicate padding, 27 fixed weights);
// acc[0] += sum err^2, acc[1] += 3 * [mask]; derr [3][X][Y][Z] <- err (0 outside the mask) for the backward
__global__ void __launch_bounds__(256)
    k_smooth_grad_fwd(const float *__restrict__ gvol, const uint8_t *__restrict__ mask, int X, int Y, int Z,
                      const float *__restrict__ w27, float bias, double *__restrict__ acc, float *__restrict__ derr) {
  __shared__ double sh[8];
  __shared__ float w[27];
  if (threadIdx.x < 27) w[threadIdx.x] = w27[threadIdx.x];
  __syncthreads();
  const int64_t n = (int64_t)X * Y * Z;
  double s = 0.0, cnt = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool on = !mask || mask[i];
    const int z = (int)(i % Z), y = (int)((i / Z) % Y), x = (int)(i / ((int64_t)Z * Y));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float e = 0.f;
      if (on) {
        const float *gc = gvol + c * n;
        float sm = 0.f;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz) {
              const int xx = min(max(x + dx, 0), X - 1), yy = min(max(y + dy, 0), Y - 1), zz = min(max(z + dz, 0), Z - 1);
              sm = fmaf(w[(dx + 1) * 9 + (dy + 1) * 3 + (dz + 1)], __ldg(gc + ((int64_t)xx * Y + yy) * Z + zz), sm);
            }
        e = sm + bias - __ldg(gc + i);
        s += (double)e * (double)e;
      }
      if (derr) derr[c * n + i] = e;
    }
    if (on) cnt += 3.0;
  }
  const double ss = block_sum(s, sh), cc = block_sum(cnt, sh);
  if (threadIdx.x == 0) {
    atomicAdd(acc, ss);
    atomicAdd(acc + 1, cc);
  }
}

// The smoothed volume is detached (voxurff.py:612): d L / d grad_c(p) = -2 err_c(p) g / count; through the central
// differences: d L / d sdf(u) += (dgrad_a(u - e_a) - dgrad_a(u + e_a)) / (2 h) over the interior neighbours.
__global__ void __launch_bounds__(256)
    k_smooth_grad_bwd(const float *__restrict__ derr, int X, int Y, int Z, float voxel_size, const double *__restrict__ acc,
                      const float *__restrict__ g_out, float scale, float *__restrict__ grad_sdf) {
  const int64_t n = (int64_t)X * Y * Z;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = (int)(i % Z), y = (int)((i / Z) % Y), x = (int)(i / ((int64_t)Z * Y));
  // acc == NULL: plain transpose of the central differences (derr = the cotangent of neus_sdf_gradient's output)
  const float k = (acc ? (float)(-2.0 * (double)(__ldg(g_out) * scale) / acc[1]) : scale) / (2.f * voxel_size);
  float t = 0.f;
  // voxel u receives from the central difference at u - e_a (as its "+" neighbour) and at u + e_a (as its "-" neighbour),
  // where those voxels are interior along a
  if (x >= 2) t += __ldg(derr + 0 * n + i - (int64_t)Y * Z);
  if (x + 2 < X) t -= __ldg(derr + 0 * n + i + (int64_t)Y * Z);
  if (y >= 2) t += __ldg(derr + 1 * n + i - Z);
  if (y + 2 < Y) t -= __ldg(derr + 1 * n + i + Z);
  if (z >= 2) t += __ldg(derr + 2 * n + i - 1);
  if (z + 2 < Z) t -= __ldg(derr + 2 * n + i + 1);
  grad_sdf[i] += k * t;
}

unsigned sweep_grid(int64_t n) {
  const int64_t want = (n + 255) / 256, cap = (int64_t)num_sms() * 16;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

extern "C" int esr_grid_tv_fwd(const float *v, const uint8_t *mask, int channels, int64_t X, int64_t Y, int64_t Z,
                               int64_t stride_c, int64_t stride_x, int64_t stride_y, int64_t stride_z, double *acc6,
                               esr_stream_t stream) {
  ESR_CHECK_ARG(v && acc6 && channels >= 1 && X >= 1 && Y >= 1 && Z >= 1 && X * Y * Z < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  ESR_CHECK_CUDA(cudaMemsetAsync(acc6, 0, 6 * sizeof(double), st));
  const Vol V{channels, (int)X, (int)Y, (int)Z, stride_c, stride_x, stride_y, stride_z};
  ESR_STAGE("k_tv_fwd", stream);
  k_tv_fwd<<<sweep_grid(X * Y * Z), 256, 0, st>>>(V, v, mask, acc6);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_grid_tv_bwd(const float *v, const uint8_t *mask, int channels, int64_t X, int64_t Y, int64_t Z,
                               int64_t stride_c, int64_t stride_x, int64_t stride_y, int64_t stride_z, const double *acc6,
                               const float *g_out, float scale, float *grad, esr_stream_t stream) {
  ESR_CHECK_ARG(v && acc6 && g_out && grad && channels >= 1 && X >= 1 && Y >= 1 && Z >= 1 && X * Y * Z < (1ll << 31));
  const Vol V{channels, (int)X, (int)Y, (int)Z, stride_c, stride_x, stride_y, stride_z};
  ESR_STAGE("k_tv_bwd", stream);
  k_tv_bwd<<<sweep_grid(X * Y * Z), 256, 0, (cudaStream_t)stream>>>(V, v, mask, acc6, g_out, scale, grad);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_sdf_central_gradient(const float *sdf, int64_t X, int64_t Y, int64_t Z, float voxel_size, float *out,
                                        esr_stream_t stream) {
  ESR_CHECK_ARG(sdf && out && X >= 1 && Y >= 1 && Z >= 1 && X * Y * Z < (1ll << 31) && voxel_size > 0.f);
  ESR_STAGE("k_sdf_central_gradient", stream);
  k_sdf_central_gradient<<<cdiv(X * Y * Z, 256), 256, 0, (cudaStream_t)stream>>>(sdf, (int)X, (int)Y, (int)Z, voxel_size, out);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_smooth_grad_tv_fwd(const float *grad_vol, const uint8_t *mask, int64_t X, int64_t Y, int64_t Z,
                                      const float *w27, float bias, double *acc2, float *err_vol, esr_stream_t stream) {
  ESR_CHECK_ARG(grad_vol && w27 && acc2 && X >= 1 && Y >= 1 && Z >= 1 && X * Y * Z < (1ll << 31));
  cudaStream_t st = (cudaStream_t)stream;
  ESR_CHECK_CUDA(cudaMemsetAsync(acc2, 0, 2 * sizeof(double), st));
  ESR_STAGE("k_smooth_grad_fwd", stream);
  k_smooth_grad_fwd<<<sweep_grid(X * Y * Z), 256, 0, st>>>(grad_vol, mask, (int)X, (int)Y, (int)Z, w27, bias, acc2, err_vol);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_smooth_grad_tv_bwd(const float *err_vol, int64_t X, int64_t Y, int64_t Z, float voxel_size,
                                      const double *acc2, const float *g_out, float scale, float *grad_sdf,
                                      esr_stream_t stream) {
  ESR_CHECK_ARG(err_vol && grad_sdf && (!acc2 == !g_out) && X >= 1 && Y >= 1 && Z >= 1 && X * Y * Z < (1ll << 31) && voxel_size > 0.f);
  ESR_STAGE("k_smooth_grad_bwd", stream);
  k_smooth_grad_bwd<<<cdiv(X * Y * Z, 256), 256, 0, (cudaStream_t)stream>>>(err_vol, (int)X, (int)Y, (int)Z, voxel_size, acc2,
                                                                          g_out, scale, grad_sdf);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

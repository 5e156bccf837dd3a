// optim.cu — the optimizer step the stage drivers run after loss.backward() (app/utils/optimizer.py:63-228: dense Adam
// with an optional per-voxel learning-rate volume), SURVEY.md §8f row 3.  The reference makes ~10 elementwise passes
// per parameter tensor (mul_, add_, addcmul_, sqrt, div, add_, mul, addcdiv_) over 0.83-1.2 GB of grids; here one pass
// reads (param, grad, exp_avg, exp_avg_sq[, per_lr]) and writes (param, exp_avg, exp_avg_sq): 28-32 B per element,
// HBM-bound.  Same operation order as optimizer.py:197-228.
#include "common.cuh"

using namespace esr;

namespace {

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay;
  float step_size;       // lr / (1 - beta1^step)
  float inv_sqrt_bc2;    // 1 / sqrt(1 - beta2^step)
  // 1 - beta formed in double on the host and rounded ONCE, as the Python floats the reference passes for alpha / value
  // (optimizer.py:213-214); 1.f - beta1 in float differs in the last place (9e-7 relative for beta2 = 0.99)
  float one_minus_beta1, one_minus_beta2;
};

ESR_D void adam_one(float &p, float g, float &m, float &v, float per_lr, const AdamArgs &a) {
  if (a.weight_decay != 0.f) g = __fmaf_rn(a.weight_decay, p, g);
  m = __fmaf_rn(a.one_minus_beta1, g, m * a.beta1);             // exp_avg.mul_(beta1).add_(grad, alpha=1-beta1)
  v = __fmaf_rn(a.one_minus_beta2 * g, g, v * a.beta2);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;        // (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps)
  p = __fmaf_rn(-a.step_size, __fdiv_rn(m * per_lr, denom), p);  // param.addcdiv_(exp_avg * per_lr, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
    k_adam_step(float *__restrict__ param, const float *__restrict__ grad, float *__restrict__ exp_avg,
                float *__restrict__ exp_avg_sq, const float *__restrict__ per_lr, int64_t n, AdamArgs a) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4 *>(param)[i];
    const float4 g = __ldcs(reinterpret_cast<const float4 *>(grad) + i);
    float4 m = reinterpret_cast<float4 *>(exp_avg)[i], v = reinterpret_cast<float4 *>(exp_avg_sq)[i];
    const float4 l = per_lr ? __ldg(reinterpret_cast<const float4 *>(per_lr) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    adam_one(p.x, g.x, m.x, v.x, l.x, a);
    adam_one(p.y, g.y, m.y, v.y, l.y, a);
    adam_one(p.z, g.z, m.z, v.z, l.z, a);
    adam_one(p.w, g.w, m.w, v.w, l.w, a);
    reinterpret_cast<float4 *>(param)[i] = p;
    reinterpret_cast<float4 *>(exp_avg)[i] = m;
    reinterpret_cast<float4 *>(exp_avg_sq)[i] = v;
  }
  // tail (n not a multiple of 4)
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_one(param[i], grad[i], exp_avg[i], exp_avg_sq[i], per_lr ? per_lr[i] : 1.f, a);
}

}  // namespace

extern "C" int esr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, const float *per_lr,
                             int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                             esr_stream_t stream) {
  ESR_CHECK_ARG(n >= 0 && step >= 1);
  if (n == 0) return ESR_OK;
  ESR_CHECK_ARG(param && grad && exp_avg && exp_avg_sq);
  ESR_CHECK_ARG(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)per_lr) % 16 == 0);
  AdamArgs a;
  a.lr = lr, a.beta1 = beta1, a.beta2 = beta2, a.eps = eps, a.weight_decay = weight_decay;
  // bias corrections in double on the host exactly like the Python floats of optimizer.py:206-207,224
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  a.one_minus_beta1 = (float)(1.0 - (double)beta1);
  a.one_minus_beta2 = (float)(1.0 - (double)beta2);
  const int64_t want = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  ESR_STAGE("k_adam_step", stream);
  k_adam_step<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, per_lr, n, a);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

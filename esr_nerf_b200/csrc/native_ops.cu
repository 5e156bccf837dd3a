// native_ops.cu — reference-shaped replacements for the reference's live pybind entry points
// (app/utils/base/cuda/render_utils.cpp:170-184): sample_pts_on_rays, alpha2weight,
// alpha2weight_backward; torch_scatter.segment_coo(sum); total_variation_add_grad.
// int64 indices / bool masks at this boundary, exactly like the reference.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace esr {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct TimingRec {
  const char *name;
  cudaEvent_t a, b;
};
static std::atomic<bool> g_timing{false};
static std::mutex g_timing_mu;
static std::vector<TimingRec> g_recs;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t get_event() {
  cudaEvent_t e;
  if (!g_event_pool.empty()) {
    e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEventCreate(&e);
  return e;
}

static thread_local int g_pending = -1;
static thread_local cudaStream_t g_pending_stream = nullptr;

void stage_begin(const char *name, cudaStream_t st) {
  if (!g_timing.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  TimingRec r{name, get_event(), get_event()};
  cudaEventRecord(r.a, st);
  g_pending = (int)g_recs.size();
  g_pending_stream = st;
  g_recs.push_back(r);
}
void stage_end() {
  if (g_pending < 0) return;
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (g_pending < (int)g_recs.size()) cudaEventRecord(g_recs[g_pending].b, g_pending_stream);
  g_pending = -1;
}
}  // namespace esr

extern "C" int esr_stage_timing(int enable) {
  std::lock_guard<std::mutex> lk(esr::g_timing_mu);
  for (auto &r : esr::g_recs) {
    esr::g_event_pool.push_back(r.a);
    esr::g_event_pool.push_back(r.b);
  }
  esr::g_recs.clear();
  esr::g_timing.store(enable != 0);
  return ESR_OK;
}

extern "C" int64_t esr_stage_timing_report(char *buf, int64_t buf_bytes) {
  std::lock_guard<std::mutex> lk(esr::g_timing_mu);
  std::map<std::string, std::pair<long long, double>> acc;
  std::vector<std::string> order;
  for (auto &r : esr::g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    auto it = acc.find(r.name);
    if (it == acc.end()) {
      order.push_back(r.name);
      acc[r.name] = {1, (double)ms};
    } else {
      it->second.first += 1;
      it->second.second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto &n : order) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", n.c_str(), acc[n].first, acc[n].second);
    out += line;
  }
  if (buf && buf_bytes > 0) {
    const int64_t k = (int64_t)out.size() < buf_bytes - 1 ? (int64_t)out.size() : buf_bytes - 1;
    memcpy(buf, out.data(), (size_t)k);
    buf[k] = 0;
  }
  return (int64_t)out.size() + 1;
}

using namespace esr;

extern "C" const char *esr_last_error(void) { return esr::g_err; }
extern "C" int esr_version(void) { return 100; }
extern "C" int64_t esr_launch_count(void) { return (int64_t)esr::g_launches.load(); }
extern "C" int64_t esr_scan_scratch_bytes(int64_t n) { return (int64_t)(cdiv(n > 0 ? n : 1, SCAN_TILE) + 2) * 8; }

struct Box {
  float mn[3], mx[3];
};

// ---------------------------------------------------------------------------------------------
// sample_pts_on_rays
// ---------------------------------------------------------------------------------------------
__global__ void k_ray_counts(const float *__restrict__ rays_o, const float *__restrict__ rays_d, Box box, float near,
                             float far, float stepdist, int64_t n_rays, int64_t *__restrict__ N_steps,
                             float *__restrict__ t_min, float *__restrict__ t_max) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const RaySetup s = ray_setup(rays_o, rays_d, r, box.mn, box.mx, near, far, stepdist);
  N_steps[r] = s.n;
  t_min[r] = s.t_min;
  t_max[r] = s.t_max;
}

// one warp per ray: lanes stride over the ray's steps -> fully coalesced 28 B/sample stores
__global__ void __launch_bounds__(256) k_ray_fill(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                  Box box, float near, float far, float stepdist, int64_t n_rays,
                                                  const int64_t *__restrict__ N_cum, float *__restrict__ ray_pts,
                                                  uint8_t *__restrict__ mask_outbbox, int64_t *__restrict__ ray_id,
                                                  int64_t *__restrict__ step_id) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t r = warp; r < n_rays; r += nwarps) {
    const RaySetup s = ray_setup(rays_o, rays_d, r, box.mn, box.mx, near, far, stepdist);
    const int64_t base = r ? N_cum[r - 1] : 0;
    for (int k = lane; k < s.n; k += 32) {
      float px, py, pz;
      ray_point(s, stepdist, k, px, py, pz);
      const int64_t i = base + k;
      ray_pts[3 * i] = px;
      ray_pts[3 * i + 1] = py;
      ray_pts[3 * i + 2] = pz;
      mask_outbbox[i] = out_bbox(box.mn, box.mx, px, py, pz) ? 1 : 0;
      ray_id[i] = r;
      step_id[i] = k;
    }
  }
}

extern "C" int esr_sample_pts_on_rays_count(const float *rays_o, const float *rays_d, const float xyz_min[3],
                                            const float xyz_max[3], float near, float far, float stepdist,
                                            int64_t n_rays, int64_t *N_steps, int64_t *N_cum, float *t_min,
                                            float *t_max, int64_t *total, void *scratch, esr_stream_t stream) {
  ESR_CHECK_ARG(n_rays >= 0 && stepdist > 0.f);
  ESR_CHECK_ARG(total && scratch);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rays == 0) {
    ESR_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
    return ESR_OK;
  }
  ESR_CHECK_ARG(rays_o && rays_d && xyz_min && xyz_max && N_steps && N_cum && t_min && t_max);
  Box box;
  for (int i = 0; i < 3; ++i) box.mn[i] = xyz_min[i], box.mx[i] = xyz_max[i];
  ESR_STAGE("k_ray_counts", st);
  k_ray_counts<<<cdiv(n_rays, 256), 256, 0, st>>>(rays_o, rays_d, box, near, far, stepdist, n_rays, N_steps, t_min,
                                                  t_max);
  ESR_LAUNCH_OK();
  return device_scan<int64_t, int64_t, false>(N_steps, N_cum, n_rays, total, scratch, st);
}

extern "C" int esr_sample_pts_on_rays_fill(const float *rays_o, const float *rays_d, const float xyz_min[3],
                                           const float xyz_max[3], float near, float far, float stepdist,
                                           int64_t n_rays, const int64_t *N_cum, int64_t total, float *ray_pts,
                                           uint8_t *mask_outbbox, int64_t *ray_id, int64_t *step_id,
                                           esr_stream_t stream) {
  ESR_CHECK_ARG(n_rays >= 0 && total >= 0);
  if (n_rays == 0 || total == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && xyz_min && xyz_max && N_cum && ray_pts && mask_outbbox && ray_id && step_id);
  Box box;
  for (int i = 0; i < 3; ++i) box.mn[i] = xyz_min[i], box.mx[i] = xyz_max[i];
  const int64_t blocks = min((int64_t)cdiv(n_rays, 8), (int64_t)num_sms() * 16);
  ESR_STAGE("k_ray_fill", (cudaStream_t)stream);
  k_ray_fill<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, box, near, far, stepdist, n_rays,
                                                                 N_cum, ray_pts, mask_outbbox, ray_id, step_id);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
// alpha2weight
// ---------------------------------------------------------------------------------------------
// render_utils_kernel.cu:607-617 (+ the host-side i_end[ray_id[n-1]] = n at :635)
__global__ void k_segments(const int64_t *__restrict__ ray_id, int64_t n_pts, int64_t *__restrict__ i_start,
                           int64_t *__restrict__ i_end) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const int64_t r = ray_id[i];
  if (i > 0) {
    const int64_t rp = ray_id[i - 1];
    if (r != rp) {
      i_start[r] = i;
      i_end[rp] = i;
    }
  }
  if (i == n_pts - 1) i_end[r] = n_pts;
}

// The reference walks each ray with ONE thread (kernel.cu:576-605).  Here a warp owns a ray: alpha is
// read coalesced 32 at a time, and the transmittance recurrence is replayed in the reference's exact
// sequential order and precision (float T, double (1.-alpha) product, stop when T < 1e-3) by all lanes
// in lock-step, so weights / T / stop index are bit-identical while loads and stores stay coalesced.
__global__ void __launch_bounds__(256) k_alpha2weight(const float *__restrict__ alpha, int64_t n_rays,
                                                      float *__restrict__ weight, float *__restrict__ T,
                                                      float *__restrict__ alphainv_last,
                                                      const int64_t *__restrict__ i_start, int64_t *__restrict__ i_end) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t r = warp; r < n_rays; r += nwarps) {
    const int64_t s = i_start[r], e = i_end[r];
    float Tc = 1.f;
    bool done = false;
    int64_t stop = e;
    for (int64_t base = s; base < e; base += 32) {
      const int64_t i = base + lane;
      const bool valid = i < e;
      const float a = valid ? __ldg(alpha + i) : 0.f;
      float myT = 1.f, myW = 0.f;
      if (!done) {
        const int cnt = (int)min((int64_t)32, e - base);
        for (int j = 0; j < cnt; ++j) {
          const float aj = __shfl_sync(FULL, a, j);
          if ((int)lane == j) {
            myT = Tc;
            myW = __fmul_rn(Tc, aj);
          }
          Tc = (float)((1. - (double)aj) * (double)Tc);
          if ((double)Tc < 1e-3) {
            done = true;
            stop = base + j + 1;
            break;
          }
        }
      }
      if (valid) {
        weight[i] = myW;
        T[i] = myT;
      }
    }
    if (lane == 0) {
      i_end[r] = stop;
      alphainv_last[r] = Tc;
    }
  }
}

// kernel.cu:653-677, reverse recurrence as a warp suffix scan (32 samples per step, coalesced)
__global__ void __launch_bounds__(256) k_alpha2weight_bwd(const float *__restrict__ alpha,
                                                          const float *__restrict__ weight, const float *__restrict__ T,
                                                          const float *__restrict__ alphainv_last,
                                                          const int64_t *__restrict__ i_start,
                                                          const int64_t *__restrict__ i_end, int64_t n_rays,
                                                          const float *__restrict__ grad_weights,
                                                          const float *__restrict__ grad_last, float *__restrict__ grad) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t r = warp; r < n_rays; r += nwarps) {
    const int64_t s = i_start[r], e = i_end[r];
    float carry = __fmul_rn(grad_last[r], alphainv_last[r]);
    for (int64_t hi = e; hi > s; hi -= 32) {
      const int64_t i = hi - 1 - lane;
      const bool valid = i >= s;
      const float gw = valid ? grad_weights[i] : 0.f;
      const float x = valid ? gw * weight[i] : 0.f;
      float inc = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(FULL, inc, o);
        if (lane >= (unsigned)o) inc += u;
      }
      const float back_cum = carry + (inc - x);
      if (valid) {
        const float gwT = __fmul_rn(gw, T[i]);
        const double den = (double)(1 - alpha[i]) + 1e-10;
        grad[i] = (float)((double)gwT - (double)back_cum / den);
      }
      carry += __shfl_sync(FULL, inc, 31);
    }
  }
}

extern "C" int esr_alpha2weight_fwd(const float *alpha, const int64_t *ray_id, int64_t n_pts, int64_t n_rays,
                                    float *weight, float *T, float *alphainv_last, int64_t *i_start, int64_t *i_end,
                                    esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && n_rays >= 0);
  if (n_rays == 0) return ESR_OK;
  ESR_CHECK_ARG(alphainv_last && i_start && i_end);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_CHECK_CUDA(cudaMemsetAsync(i_start, 0, sizeof(int64_t) * n_rays, st));
  ESR_CHECK_CUDA(cudaMemsetAsync(i_end, 0, sizeof(int64_t) * n_rays, st));
  if (n_pts > 0) {
    ESR_CHECK_ARG(alpha && ray_id && weight && T);
    ESR_STAGE("k_segments", st);
    k_segments<<<cdiv(n_pts, 256), 256, 0, st>>>(ray_id, n_pts, i_start, i_end);
    ESR_LAUNCH_OK();
  }
  const int64_t blocks = min((int64_t)cdiv(n_rays, 8), (int64_t)num_sms() * 16);
  ESR_STAGE("k_alpha2weight", st);
  k_alpha2weight<<<(unsigned)blocks, 256, 0, st>>>(alpha, n_rays, weight, T, alphainv_last, i_start, i_end);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_alpha2weight_bwd(const float *alpha, const float *weight, const float *T,
                                    const float *alphainv_last, const int64_t *i_start, const int64_t *i_end,
                                    int64_t n_pts, int64_t n_rays, const float *grad_weights, const float *grad_last,
                                    float *grad_alpha, esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && n_rays >= 0);
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(alpha && weight && T && alphainv_last && i_start && i_end && grad_weights && grad_last && grad_alpha);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_CHECK_CUDA(cudaMemsetAsync(grad_alpha, 0, sizeof(float) * n_pts, st));
  if (n_rays == 0) return ESR_OK;
  const int64_t blocks = min((int64_t)cdiv(n_rays, 8), (int64_t)num_sms() * 16);
  ESR_STAGE("k_alpha2weight_bwd", st);
  k_alpha2weight_bwd<<<(unsigned)blocks, 256, 0, st>>>(alpha, weight, T, alphainv_last, i_start, i_end, n_rays,
                                                       grad_weights, grad_last, grad_alpha);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
// segment_coo(sum)
// ---------------------------------------------------------------------------------------------
// 32 consecutive samples per warp; runs of equal (sorted) index are reduced with a segmented
// shuffle scan, the head lane of each run issues one RED per channel.
__global__ void __launch_bounds__(256) k_segment_sum(const float *__restrict__ src, const int64_t *__restrict__ index,
                                                     int64_t n_pts, int channels, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = lane_id();
  const bool valid = i < n_pts;
  const int64_t idx = valid ? index[i] : -1;
  const int64_t idx_prev = __shfl_up_sync(FULL, idx, 1);
  const bool head = valid && (lane == 0 || idx != idx_prev);
  for (int c = 0; c < channels; ++c) {
    float v = valid ? src[i * channels + c] : 0.f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float u = __shfl_down_sync(FULL, v, o);
      const int64_t j = __shfl_down_sync(FULL, idx, o);
      if (lane + o < 32 && j == idx) v += u;
    }
    if (head) red_add(out + idx * channels + c, v);
  }
}

__global__ void k_segment_gather(const float *__restrict__ grad_out, const int64_t *__restrict__ index, int64_t n_elems,
                                 int channels, float *__restrict__ grad_src) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int64_t i = e / channels;
  const int c = (int)(e - i * channels);
  grad_src[e] = grad_out[index[i] * channels + c];
}

extern "C" int esr_segment_sum_fwd(const float *src, const int64_t *index, int64_t n_pts, int channels, int64_t n_out,
                                   float *out, esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && channels >= 1 && n_out >= 0);
  if (n_out == 0) return ESR_OK;
  ESR_CHECK_ARG(out);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_out * channels, st));
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(src && index);
  ESR_STAGE("k_segment_sum", st);
  k_segment_sum<<<cdiv(n_pts, 256), 256, 0, st>>>(src, index, n_pts, channels, out);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_segment_sum_bwd(const float *grad_out, const int64_t *index, int64_t n_pts, int channels,
                                   float *grad_src, esr_stream_t stream) {
  ESR_CHECK_ARG(n_pts >= 0 && channels >= 1);
  if (n_pts == 0) return ESR_OK;
  ESR_CHECK_ARG(grad_out && index && grad_src);
  const int64_t n = n_pts * channels;
  ESR_STAGE("k_segment_gather", (cudaStream_t)stream);
  k_segment_gather<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, index, n, channels, grad_src);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
// total_variation_add_grad (total_variation_kernel.cu:14-35): wx unused, wz on both k and i axes.
// ---------------------------------------------------------------------------------------------
ESR_D float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

template <bool DENSE>
__global__ void k_tv_add_grad(const float *__restrict__ param, float *__restrict__ grad, float wy, float wz,
                              int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t N) {
  const int64_t index = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (index >= N) return;
  const float g0 = grad[index];
  if (!DENSE && g0 == 0.f) return;
  const int64_t k = index % sz_k;
  const int64_t j = index / sz_k % sz_j;
  const int64_t i = index / sz_k / sz_j % sz_i;
  const float p = param[index];
  float add = 0.f;
  add += (k == 0 ? 0.f : wz * clamp1(p - param[index - 1]));
  add += (k == sz_k - 1 ? 0.f : wz * clamp1(p - param[index + 1]));
  add += (j == 0 ? 0.f : wy * clamp1(p - param[index - sz_k]));
  add += (j == sz_j - 1 ? 0.f : wy * clamp1(p - param[index + sz_k]));
  add += (i == 0 ? 0.f : wz * clamp1(p - param[index - sz_k * sz_j]));
  add += (i == sz_i - 1 ? 0.f : wz * clamp1(p - param[index + sz_k * sz_j]));
  grad[index] = g0 + add;
}

extern "C" int esr_tv_add_grad(const float *param, float *grad, float wx, float wy, float wz, int64_t sz_i,
                               int64_t sz_j, int64_t sz_k, int64_t n_total, int dense_mode, esr_stream_t stream) {
  (void)wx;
  ESR_CHECK_ARG(param && grad && sz_i > 0 && sz_j > 0 && sz_k > 0 && n_total >= 0);
  if (n_total == 0) return ESR_OK;
  wy /= 6;
  wz /= 6;
  ESR_STAGE("k_tv_add_grad", stream);
  if (dense_mode)
    k_tv_add_grad<true><<<cdiv(n_total, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, wy, wz, sz_i, sz_j, sz_k,
                                                                              n_total);
  else
    k_tv_add_grad<false><<<cdiv(n_total, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, wy, wz, sz_i, sz_j, sz_k,
                                                                               n_total);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, void *scratch,
                                      esr_stream_t stream) {
  ESR_CHECK_ARG(n >= 0 && out && scratch);
  ESR_CHECK_ARG(n == 0 || in);
  return device_scan<int32_t, int32_t, true>(in, out, n, (int32_t *)nullptr, scratch, (cudaStream_t)stream);
}

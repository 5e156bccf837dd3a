// render_step.cu — ONE C call = one fine-stage render step (last row): esr_render_voxurff_fwd /
// esr_render_voxurff_bwd run the whole chain of VoxurfF.forward_training (app/fine/model/voxurff.py:177-278) and its
// backward on the library's own stage entry points, in the order and with the buffers esr_nerf_b200/fused.py uses, so
// that a host that is not Python can run a training step without re-implementing that orchestration:
//
//   fwd : emission-on rays first (stable partition) -> march count -> [host read: M1, n_on] -> march fill -> NeuS alpha +
//         transmittance + shaded count -> [host read: M3, M3_on] -> shaded stream -> feature encode -> radiance nets (off: all
//         rows, emo: the emission-on prefix) -> combine (voxurff.py:243-254) -> tone map -> compositing
//   bwd : compositing -> tone map -> radiance nets (each through its own row range) -> encode scatter -> alpha scan + SDF scatter
//
// All memory is the caller's: parameters, gradients (accumulated into), outputs, and ONE workspace from which every
// intermediate is carved (bump allocation, 256-byte aligned).  The two stream sizes are data dependent: the forward
// synchronises the stream twice to read them (as the Python path does) and fails with ESR_ERR_CAPACITY — reporting the
// bytes it would need in step->workspace_needed — when the workspace is too small; esr_render_voxurff_workspace_bytes
// gives the size for given stream-size bounds.
#include "common.cuh"
#include "mlp_layout.cuh"

using namespace esr;

namespace {

struct Bump {
  uint8_t *base;
  int64_t cap, used;
  template <typename T>
  T *take(int64_t count) {
    const int64_t bytes = (count * (int64_t)sizeof(T) + 255) / 256 * 256;
    T *p = (used + bytes <= cap && base) ? reinterpret_cast<T *>(base + used) : nullptr;
    used += bytes;
    return p;
  }
  bool ok() const { return used <= cap && base != nullptr; }
};

// indices into esr_voxurff_step_t::slot
enum Slot {
  S_RAY_ORDER, S_N_STEPS, S_CNT_IN, S_CNT_MASK, S_OFF_MASK, S_BITS, S_SRAY, S_SSTEP, S_SSDF, S_SALPHA, S_ST, S_CNT_SHADE,
  S_OFF_SHADE, S_HRAY, S_HSTEP, S_HM1, S_HW, S_HSDF, S_X, S_FD, S_IMG_OFF, S_IMG_EMO, S_IMG_TONE, S_LIN_OFF, S_LIN_EMO,
  S_HID_OFF, S_HID_EMO, S_LIN, S_RGB, S_COUNT
};

__global__ void k_on_flags(const int64_t *__restrict__ em_modes, int64_t n, int32_t *__restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = em_modes[i] == 1;
}
// stable partition: emission-on rays first, both groups in their original order (torch.argsort(~on, stable=True))
__global__ void k_order_rays(const int64_t *__restrict__ em_modes, const int32_t *__restrict__ off_on, int64_t n,
                             int32_t *__restrict__ ray_order) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t before = off_on[i], n_on = off_on[n];
  ray_order[em_modes[i] == 1 ? before : n_on + (int32_t)i - before] = (int32_t)i;
}

int bits_stride(const esr_scene_t &sc) {
  double d2 = 0.0;
  for (int i = 0; i < 3; ++i) d2 += (double)(sc.xyz_max[i] - sc.xyz_min[i]) * (double)(sc.xyz_max[i] - sc.xyz_min[i]);
  return (int)(sqrt(d2) / sc.stepdist) / 32 + 2;
}

esr_mlp_desc_t radiance_desc(int precision) { return esr_mlp_desc_t{96, 192, 3, 3, 1, precision}; }
esr_mlp_desc_t tonemap_desc(int precision) { return esr_mlp_desc_t{48, 192, 1, 3, 2, precision}; }

int64_t bytes_after_m1(const esr_scene_t &, int64_t n, int64_t m1, int64_t m3, int precision) {
  auto r = [](int64_t b) { return (b + 255) / 256 * 256; };
  const esr_mlp_desc_t rd = radiance_desc(precision), td = tonemap_desc(precision);
  int64_t t = 0;
  t += 5 * r(4 * m1);                                              // s_ray, s_step, s_sdf, s_alpha, s_T
  t += r(4 * n) + r(4 * (n + 1));                                  // cnt_shade, off_shade
  t += 5 * r(4 * m3);                                              // h_ray, h_step, h_m1, h_w, h_sdf
  t += r(act_rows_padded(m3) * 96 * 2 * (precision ? 2 : 1)) + r(m3 * 64);   // x tiles, saved finite differences
  t += 2 * r(tc_image_bytes(&rd)) + r(tc_image_bytes(&td));
  t += 4 * r(12 * m3);                                             // lin_off, lin_emo, lin, rgb
  t += 2 * r(act_hidden_bytes(3, m3));
  // backward scratch (carved at the end of the forward's region by esr_render_voxurff_bwd)
  t += 3 * r(12 * m3) + r(4 * m1) * 3 + r(m3 * 56 * 4) + r(act_dz_bytes(3, m3));
  return t;
}

}  // namespace

extern "C" int64_t esr_render_voxurff_workspace_bytes(const esr_scene_t *sc, int64_t n_rays, int64_t m1_max, int64_t m3_max,
                                                      int precision) {
  if (!sc || n_rays < 0 || m1_max < 0 || m3_max < 0) return 0;
  auto r = [](int64_t b) { return (b + 255) / 256 * 256; };
  int64_t t = 0;
  t += 4 * r(4 * n_rays) + 2 * r(4 * (n_rays + 1)) + r(4 * n_rays * bits_stride(*sc)) + 2 * r(esr_scan_scratch_bytes(n_rays + 1));
  return t + bytes_after_m1(*sc, n_rays, m1_max, m3_max, precision);
}

extern "C" int esr_render_voxurff_fwd(esr_voxurff_step_t *step, const float *rays_o, const float *rays_d,
                                      const float *viewdirs, const int64_t *em_modes, int64_t n_rays, float *rgb_marched,
                                      float *lin_marched, float *alphainv_last, esr_stream_t stream) {
  ESR_CHECK_ARG(step && rays_o && rays_d && viewdirs && em_modes && n_rays > 0 && rgb_marched && lin_marched && alphainv_last);
  ESR_CHECK_ARG(step->mask_density && step->sdf_grid && step->off_color_grid && step->emo_color_grid && step->flat_off &&
                step->flat_emo && step->flat_tone && (step->precision == 0 || step->precision == 1));
  cudaStream_t st = (cudaStream_t)stream;
  const esr_scene_t *sc = &step->scene;
  const int64_t n = n_rays;
  Bump ws{reinterpret_cast<uint8_t *>(step->workspace), step->workspace_bytes, 0};
  void **slot = step->slot;
  static_assert(S_COUNT <= 32, "esr_voxurff_step_t::slot");
  for (int i = 0; i < S_COUNT; ++i) slot[i] = nullptr;
  step->n_rays = n, step->m1 = step->m3 = step->m3_on = step->n_on = 0;
  step->alphainv_last = alphainv_last;

  // ---- ray order + march count ----
  int32_t *ray_order = ws.take<int32_t>(n), *n_steps = ws.take<int32_t>(n), *cnt_in = ws.take<int32_t>(n);
  int32_t *cnt_mask = ws.take<int32_t>(n), *off_mask = ws.take<int32_t>(n + 1), *off_on = ws.take<int32_t>(n + 1);
  const int stride = bits_stride(*sc);
  uint32_t *bits = ws.take<uint32_t>(n * stride);
  void *scratch = ws.take<uint8_t>(esr_scan_scratch_bytes(n + 1)), *scratch2 = ws.take<uint8_t>(esr_scan_scratch_bytes(n + 1));
  if (!ws.ok()) {
    step->workspace_needed = esr_render_voxurff_workspace_bytes(sc, n, n, n, step->precision);
    set_error("esr_render_voxurff_fwd: workspace too small for the per-ray arrays of %lld rays", (long long)n);
    return ESR_ERR_CAPACITY;
  }
  ESR_STAGE("k_order_rays", st);
  k_on_flags<<<cdiv(n, 256), 256, 0, st>>>(em_modes, n, cnt_in);   // (cnt_in doubles as the flag array until the count pass)
  ESR_LAUNCH_OK();
  if (int e = esr_exclusive_scan_i32(cnt_in, off_on, n, scratch2, stream)) return e;
  ESR_STAGE("k_order_rays", st);
  k_order_rays<<<cdiv(n, 256), 256, 0, st>>>(em_modes, off_on, n, ray_order);
  ESR_LAUNCH_OK();
  if (int e = esr_march_count_bits(sc, rays_o, rays_d, ray_order, n, step->mask_density, n_steps, cnt_in, cnt_mask, bits,
                                   stride, step->mask_cls, stream))
    return e;
  if (int e = esr_exclusive_scan_i32(cnt_mask, off_mask, n, scratch, stream)) return e;
  int32_t h_m1 = 0, h_non = 0;
  ESR_CHECK_CUDA(cudaMemcpyAsync(&h_m1, off_mask + n, 4, cudaMemcpyDeviceToHost, st));
  ESR_CHECK_CUDA(cudaMemcpyAsync(&h_non, off_on + n, 4, cudaMemcpyDeviceToHost, st));
  ESR_CHECK_CUDA(cudaStreamSynchronize(st));                         // stream size #1 (the reference syncs here too: kernel.cu:212)
  const int64_t m1 = h_m1, n_on = h_non;
  step->m1 = m1, step->n_on = n_on;
  slot[S_RAY_ORDER] = ray_order, slot[S_N_STEPS] = n_steps, slot[S_CNT_IN] = cnt_in, slot[S_CNT_MASK] = cnt_mask;
  slot[S_OFF_MASK] = off_mask, slot[S_BITS] = bits;

  // ---- M1 stream, alpha scan ----
  int32_t *s_ray = ws.take<int32_t>(m1), *s_step = ws.take<int32_t>(m1);
  float *s_sdf = ws.take<float>(m1), *s_alpha = ws.take<float>(m1), *s_T = ws.take<float>(m1);
  int32_t *cnt_shade = ws.take<int32_t>(n), *off_shade = ws.take<int32_t>(n + 1);
  if (!ws.ok()) {
    step->workspace_needed = ws.used + bytes_after_m1(*sc, n, m1, m1, step->precision);
    set_error("esr_render_voxurff_fwd: workspace too small (M1 = %lld): up to %lld bytes needed", (long long)m1,
              (long long)step->workspace_needed);
    return ESR_ERR_CAPACITY;
  }
  slot[S_SRAY] = s_ray, slot[S_SSTEP] = s_step, slot[S_SSDF] = s_sdf, slot[S_SALPHA] = s_alpha, slot[S_ST] = s_T;
  slot[S_CNT_SHADE] = cnt_shade, slot[S_OFF_SHADE] = off_shade;
  if (m1 == 0) {   // no ray meets occupied space: background everywhere
    ESR_CHECK_CUDA(cudaMemsetAsync(rgb_marched, 0, 12 * n, st));
    ESR_CHECK_CUDA(cudaMemsetAsync(lin_marched, 0, 12 * n, st));
  }
  if (int e = esr_march_fill_bits(sc, rays_o, rays_d, ray_order, n, step->mask_density, step->sdf_grid, off_mask, s_ray,
                                  s_step, s_sdf, bits, stride, stream))
    return e;
  if (int e = esr_alpha_scan_count(sc, ray_order, n, off_mask, s_sdf, cnt_shade, alphainv_last, s_alpha, s_T, stream)) return e;
  if (int e = esr_exclusive_scan_i32(cnt_shade, off_shade, n, scratch, stream)) return e;
  int32_t h_m3 = 0, h_m3on = 0;
  ESR_CHECK_CUDA(cudaMemcpyAsync(&h_m3, off_shade + n, 4, cudaMemcpyDeviceToHost, st));
  ESR_CHECK_CUDA(cudaMemcpyAsync(&h_m3on, off_shade + n_on, 4, cudaMemcpyDeviceToHost, st));
  ESR_CHECK_CUDA(cudaStreamSynchronize(st));                         // stream size #2
  const int64_t m3 = h_m3, m3_on = h_m3on;
  step->m3 = m3, step->m3_on = m3_on;

  // ---- shaded stream, features, nets, compositing ----
  const esr_mlp_desc_t rd = radiance_desc(step->precision), td = tonemap_desc(step->precision);
  int32_t *h_ray = ws.take<int32_t>(m3), *h_step = ws.take<int32_t>(m3), *h_m1s = ws.take<int32_t>(m3);
  float *h_w = ws.take<float>(m3), *h_sdf = ws.take<float>(m3);
  void *x = ws.take<uint8_t>(act_rows_padded(m3) * 96 * 2 * (step->precision ? 2 : 1));
  float *fd = ws.take<float>(m3 * 16);
  void *img_off = ws.take<uint8_t>(tc_image_bytes(&rd)), *img_emo = ws.take<uint8_t>(tc_image_bytes(&rd));
  void *img_tone = ws.take<uint8_t>(tc_image_bytes(&td));
  float *lin_off = ws.take<float>(3 * m3), *lin_emo = ws.take<float>(3 * m3), *lin = ws.take<float>(3 * m3);
  float *rgb = ws.take<float>(3 * m3);
  void *hid_off = ws.take<uint8_t>(act_hidden_bytes(3, m3)), *hid_emo = ws.take<uint8_t>(act_hidden_bytes(3, m3));
  step->workspace_used = ws.used;
  step->workspace_needed = ws.used + 3 * ((12 * m3 + 255) / 256 * 256) + 3 * ((4 * m1 + 255) / 256 * 256) +
                           (m3 * 56 * 4 + 255) / 256 * 256 + (act_dz_bytes(3, m3) + 255) / 256 * 256;
  if (!ws.ok() || step->workspace_needed > step->workspace_bytes) {
    set_error("esr_render_voxurff_fwd: workspace too small (M1 = %lld, M3 = %lld): %lld bytes needed", (long long)m1,
              (long long)m3, (long long)step->workspace_needed);
    return ESR_ERR_CAPACITY;
  }
  slot[S_HRAY] = h_ray, slot[S_HSTEP] = h_step, slot[S_HM1] = h_m1s, slot[S_HW] = h_w, slot[S_HSDF] = h_sdf, slot[S_X] = x;
  slot[S_FD] = fd, slot[S_IMG_OFF] = img_off, slot[S_IMG_EMO] = img_emo, slot[S_IMG_TONE] = img_tone, slot[S_LIN_OFF] = lin_off;
  slot[S_LIN_EMO] = lin_emo, slot[S_HID_OFF] = hid_off, slot[S_HID_EMO] = hid_emo, slot[S_LIN] = lin, slot[S_RGB] = rgb;
  if (int e = esr_alpha_scan_fill(sc, ray_order, n, off_mask, s_step, s_sdf, off_shade, s_alpha, s_T, h_ray, h_step, h_m1s, h_w,
                                  h_sdf, stream))
    return e;
  if (m3 > 0) {
    if (int e = esr_encode_pbr_fwd(sc, rays_o, rays_d, viewdirs, step->sdf_grid, step->off_color_grid, step->emo_color_grid,
                                   nullptr, 6, nullptr, h_ray, h_step, h_sdf, m3, x, nullptr, step->precision ? 2 : 1, fd, stream))
      return e;
    if (int e = esr_mlp_pack(&rd, step->flat_off, img_off, stream)) return e;
    if (int e = esr_mlp_pack(&rd, step->flat_emo, img_emo, stream)) return e;
    if (int e = esr_mlp_pack(&td, step->flat_tone, img_tone, stream)) return e;
    // off net: every shaded row, activations saved for the rows it will back-propagate through (the emission-off ones)
    if (int e = esr_mlp_fwd(&rd, img_off, x, 0, m3, m3, lin_off, hid_off, m3_on, stream)) return e;
    // emo net: the emission-on prefix; rows beyond it are defined to be zero
    ESR_CHECK_CUDA(cudaMemsetAsync(lin_emo, 0, 12 * m3, st));
    if (int e = esr_mlp_fwd(&rd, img_emo, x, 0, m3_on, m3, lin_emo, hid_emo, 0, stream)) return e;
    if (int e = esr_tonemap_encode_fwd(lin_off, lin_emo, h_ray, em_modes, m3, lin, nullptr, 1, stream)) return e;
    if (int e = esr_tonemap_mlp_fwd(&td, img_tone, lin, m3, rgb, stream)) return e;
  }
  return esr_composite_fwd(ray_order, n, off_shade, h_w, rgb, lin, rgb_marched, lin_marched, stream);
}

extern "C" int esr_render_voxurff_bwd(esr_voxurff_step_t *step, const float *rays_o, const float *rays_d,
                                      const float *d_rgb_marched, const float *d_lin_marched, const float *d_alphainv_last,
                                      float *grad_sdf_grid, float *grad_off_grid, float *grad_emo_grid, float *grad_flat_off,
                                      float *grad_flat_emo, float *grad_flat_tone, esr_stream_t stream) {
  ESR_CHECK_ARG(step && rays_o && rays_d && d_rgb_marched && d_lin_marched && d_alphainv_last && grad_sdf_grid &&
                grad_off_grid && grad_emo_grid && grad_flat_off && grad_flat_emo && grad_flat_tone);
  ESR_CHECK_ARG(step->n_rays > 0 && step->slot[S_OFF_MASK] != nullptr);   // a forward ran on this step object
  cudaStream_t st = (cudaStream_t)stream;
  const esr_scene_t *sc = &step->scene;
  void **slot = step->slot;
  const int64_t n = step->n_rays, m1 = step->m1, m3 = step->m3, m3_on = step->m3_on;
  if (m1 == 0) return ESR_OK;
  Bump ws{reinterpret_cast<uint8_t *>(step->workspace), step->workspace_bytes, step->workspace_used};
  float *d_rgb = ws.take<float>(3 * m3), *d_lin_direct = ws.take<float>(3 * m3), *d_lin = ws.take<float>(3 * m3);
  float *g_w_m1 = ws.take<float>(m1), *tmp_p = ws.take<float>(m1), *tmp_n = ws.take<float>(m1);
  float *d_x = ws.take<float>(m3 * 56);
  void *d_z = ws.take<uint8_t>(act_dz_bytes(3, m3));
  if (!ws.ok()) {
    step->workspace_needed = ws.used;
    set_error("esr_render_voxurff_bwd: workspace too small: %lld bytes needed", (long long)ws.used);
    return ESR_ERR_CAPACITY;
  }
  const esr_mlp_desc_t rd = radiance_desc(step->precision), td = tonemap_desc(step->precision);
  auto I = [&](int s) { return reinterpret_cast<int32_t *>(slot[s]); };
  auto F = [&](int s) { return reinterpret_cast<float *>(slot[s]); };
  ESR_CHECK_CUDA(cudaMemsetAsync(g_w_m1, 0, 4 * m1, st));
  if (m3 > 0) {
    // compositing: cotangents of rgb / lin per shaded sample, and of the weights scattered straight onto the M1 stream
    if (int e = esr_composite_bwd(I(S_HRAY), I(S_HM1), F(S_HW), F(S_RGB), F(S_LIN), d_rgb_marched, d_lin_marched, m3, d_rgb,
                                  d_lin_direct, g_w_m1, stream))
      return e;
    if (int e = esr_tonemap_mlp_bwd(&td, slot[S_IMG_TONE], F(S_LIN), F(S_RGB), d_rgb, d_lin_direct, m3, d_lin, grad_flat_tone, stream))
      return e;
    // voxurff.py:243-254 with the emission-on rows first: on rows -> emo net (the off net sees them through a stop-gradient),
    // off rows -> off net; the two row ranges are disjoint and cover every row, so d_x is written exactly once
    if (int e = esr_mlp_bwd(&rd, slot[S_IMG_OFF], slot[S_X], F(S_LIN_OFF), d_lin, m3_on, m3, m3, slot[S_HID_OFF], d_z, nullptr, d_x,
                            56, 0, grad_flat_off, stream))
      return e;
    if (int e = esr_mlp_bwd(&rd, slot[S_IMG_EMO], slot[S_X], F(S_LIN_EMO), d_lin, 0, m3_on, m3, slot[S_HID_EMO], d_z, nullptr, d_x,
                            56, 0, grad_flat_emo, stream))
      return e;
    if (int e = esr_encode_pbr_bwd(sc, rays_o, rays_d, step->sdf_grid, 6, nullptr, I(S_HRAY), I(S_HSTEP), m3, d_x, nullptr,
                                   grad_sdf_grid, grad_off_grid, grad_emo_grid, nullptr, F(S_FD), stream))
      return e;
  }
  return esr_alpha_scan_bwd(sc, rays_o, rays_d, I(S_RAY_ORDER), n, I(S_OFF_MASK), I(S_SRAY), I(S_SSTEP), F(S_SSDF), F(S_SALPHA),
                            F(S_ST), step->alphainv_last, g_w_m1, d_alphainv_last, tmp_p, tmp_n, m1, grad_sdf_grid, stream);
}

/* The INTEGRATION.md training-step snippet as a complete translation unit: a host that is not Python runs one fine-stage
 * render step (forward, backward, gradient exchange hooks, optimizer) through the C ABI alone.  tests/test_cabi_c_host_cpu.py
 * compiles it as C99 with -Wall -Wextra -pedantic -Werror and links it against libesr_b200.so, so that the snippet cannot
 * drift from include/esr_b200.h; it is not run there (no device).  Every pointer is device memory owned by the caller. */
#include <stddef.h>
#include <stdint.h>

#include "esr_b200.h"

typedef struct {
  /* scene + parameters (device) */
  esr_scene_t scene;
  const float *mask_density;
  uint8_t *mask_cls;                     /* esr_mask_class_bytes(&scene) bytes, filled once per MaskCache */
  float *sdf, *off_color, *emo_color;    /* grids (colour grids channels-last) */
  float *w_off, *w_emo, *w_tone;         /* flat f32 master copies of the three nets */
  /* gradients (device, zeroed by the caller before the step) and Adam state */
  float *g_sdf, *g_off_color, *g_emo_color, *g_w_off, *g_w_emo, *g_w_tone;
  float *m_sdf, *v_sdf;
  int64_t n_sdf;
  /* workspace */
  void *ws;
  int64_t ws_bytes;
} host_model_t;

/* returns ESR_OK, or ESR_ERR_CAPACITY with *ws_needed set: grow the workspace and call again */
int host_train_step(host_model_t *m, const float *rays_o, const float *rays_d, const float *viewdirs, const int64_t *em_modes,
                    int64_t n_rays, float *rgb, float *lin, float *alphainv_last, const float *d_rgb, const float *d_lin,
                    const float *d_last, int64_t adam_step, float lr, int64_t *ws_needed, esr_stream_t stream) {
  esr_voxurff_step_t step;
  esr_mlp_desc_t radiance = {96, 192, 3, 3, 1, 1}, tone = {48, 192, 1, 3, 2, 1};
  int rc;
  size_t i;
  unsigned char *p = (unsigned char *)&step;
  for (i = 0; i < sizeof step; ++i) p[i] = 0;
  step.scene = m->scene;
  step.mask_density = m->mask_density;
  step.mask_cls = m->mask_cls;
  step.sdf_grid = m->sdf;
  step.off_color_grid = m->off_color;
  step.emo_color_grid = m->emo_color;
  step.flat_off = m->w_off;
  step.flat_emo = m->w_emo;
  step.flat_tone = m->w_tone;
  step.precision = 1; /* x2: every parameter gradient within 1e-2 of the reference's fp32 nets */
  step.workspace = m->ws;
  step.workspace_bytes = m->ws_bytes;
  rc = esr_render_voxurff_fwd(&step, rays_o, rays_d, viewdirs, em_modes, n_rays, rgb, lin, alphainv_last, stream);
  if (rc == ESR_ERR_CAPACITY) {
    *ws_needed = step.workspace_needed;
    return rc;
  }
  if (rc != ESR_OK) return rc;
  /* ... the caller's loss kernel turns (rgb, lin, alphainv_last) into (d_rgb, d_lin, d_last) on `stream` ... */
  rc = esr_render_voxurff_bwd(&step, rays_o, rays_d, d_rgb, d_lin, d_last, m->g_sdf, m->g_off_color, m->g_emo_color, m->g_w_off,
                              m->g_w_emo, m->g_w_tone, stream);
  if (rc != ESR_OK) return rc;
  /* multi-GPU: esr_grad_pack -> ncclAllReduce -> esr_grad_unpack on the three grid gradients goes here */
  rc = esr_adam_step(m->sdf, m->g_sdf, m->m_sdf, m->v_sdf, NULL, m->n_sdf, lr, 0.9f, 0.99f, 1e-8f, 0.0f, adam_step, stream);
  if (rc != ESR_OK) return rc;
  /* the flat MLP parameters take the same call with esr_mlp_param_count(&radiance / &tone) elements */
  (void)esr_mlp_param_count(&radiance);
  (void)esr_mlp_param_count(&tone);
  return ESR_OK;
}

/* sizing the workspace up front from bounds on the two data-dependent stream sizes */
int64_t host_workspace_bytes(const esr_scene_t *scene, int64_t n_rays, int64_t m1_max, int64_t m3_max) {
  return esr_render_voxurff_workspace_bytes(scene, n_rays, m1_max, m3_max, 1);
}

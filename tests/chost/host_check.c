/* Plain-C host of the C ABI (tests/test_cabi_c_host_cpu.py compiles it with `gcc -std=c99 -pedantic`, links it against
 * esr_nerf_b200/libesr_b200.so and runs it WITHOUT a GPU): the header is C (not only C++), every struct has the layout the
 * Python binding assumes (sizes and field offsets are printed as JSON and compared with esr_nerf_b200/_lib.py's ctypes
 * classes), and the entry points reject bad arguments with an error code and a message before touching a device. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "esr_b200.h"

#define OFF(T, f) (long)offsetof(T, f)

int main(void) {
  int rc_null, rc_neg, rc_step;
  esr_voxurff_step_t step;
  memset(&step, 0, sizeof step);
  rc_null = esr_sample_pts_on_rays_count(NULL, NULL, NULL, NULL, 0.0f, 0.0f, 0.0f, 4, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
  rc_neg = esr_adam_step(NULL, NULL, NULL, NULL, NULL, -1, 1e-3f, 0.9f, 0.99f, 1e-8f, 0.0f, 1, NULL);
  rc_step = esr_render_voxurff_fwd(&step, NULL, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL);
  printf("{\"version\": %d, \"rc_null\": %d, \"rc_neg\": %d, \"rc_step\": %d, \"err_nonempty\": %d,\n", esr_version(), rc_null,
         rc_neg, rc_step, (int)(esr_last_error() != NULL && esr_last_error()[0] != 0));
  printf(" \"ESR_OK\": %d, \"ESR_ERR_CAPACITY\": %d,\n", (int)ESR_OK, (int)ESR_ERR_CAPACITY);
  printf(" \"act_rows_0\": %lld, \"act_rows_1\": %lld, \"act_rows_129\": %lld,\n", (long long)esr_mlp_act_rows(0),
         (long long)esr_mlp_act_rows(1), (long long)esr_mlp_act_rows(129));
  printf(" \"sizeof\": {\"Scene\": %ld, \"DvgoScene\": %ld, \"MlpDesc\": %ld, \"VoxurffStep\": %ld},\n", (long)sizeof(esr_scene_t),
         (long)sizeof(esr_dvgo_scene_t), (long)sizeof(esr_mlp_desc_t), (long)sizeof(esr_voxurff_step_t));
  printf(" \"Scene\": {\"xyz_min\": %ld, \"xyz_max\": %ld, \"gx\": %ld, \"gz\": %ld, \"mask_xyz_min\": %ld, \"mask_xyz_max\": %ld, "
         "\"mx\": %ld, \"mz\": %ld, \"near\": %ld, \"far\": %ld, \"stepdist\": %ld, \"voxel_size\": %ld, \"act_shift\": %ld, "
         "\"mask_thres\": %ld, \"fast_thres\": %ld, \"s_val\": %ld, \"alpha_thres\": %ld, \"fd_eps\": %ld, \"sdf_tap_manual\": %ld},\n",
         OFF(esr_scene_t, xyz_min), OFF(esr_scene_t, xyz_max), OFF(esr_scene_t, gx), OFF(esr_scene_t, gz),
         OFF(esr_scene_t, mask_xyz_min), OFF(esr_scene_t, mask_xyz_max), OFF(esr_scene_t, mx), OFF(esr_scene_t, mz),
         OFF(esr_scene_t, near), OFF(esr_scene_t, far), OFF(esr_scene_t, stepdist), OFF(esr_scene_t, voxel_size),
         OFF(esr_scene_t, act_shift), OFF(esr_scene_t, mask_thres), OFF(esr_scene_t, fast_thres), OFF(esr_scene_t, s_val),
         OFF(esr_scene_t, alpha_thres), OFF(esr_scene_t, fd_eps), OFF(esr_scene_t, sdf_tap_manual));
  printf(" \"DvgoScene\": {\"xyz_min\": %ld, \"xyz_max\": %ld, \"gx\": %ld, \"gz\": %ld, \"near\": %ld, \"far\": %ld, \"stepdist\": %ld, "
         "\"interval\": %ld, \"act_shift\": %ld},\n",
         OFF(esr_dvgo_scene_t, xyz_min), OFF(esr_dvgo_scene_t, xyz_max), OFF(esr_dvgo_scene_t, gx), OFF(esr_dvgo_scene_t, gz),
         OFF(esr_dvgo_scene_t, near), OFF(esr_dvgo_scene_t, far), OFF(esr_dvgo_scene_t, stepdist),
         OFF(esr_dvgo_scene_t, interval), OFF(esr_dvgo_scene_t, act_shift));
  printf(" \"MlpDesc\": {\"k0\": %ld, \"width\": %ld, \"n_hidden\": %ld, \"n_out\": %ld, \"act\": %ld, \"precision\": %ld},\n",
         OFF(esr_mlp_desc_t, k0), OFF(esr_mlp_desc_t, width), OFF(esr_mlp_desc_t, n_hidden), OFF(esr_mlp_desc_t, n_out),
         OFF(esr_mlp_desc_t, act), OFF(esr_mlp_desc_t, precision));
  printf(" \"VoxurffStep\": {\"scene\": %ld, \"mask_density\": %ld, \"mask_cls\": %ld, \"sdf_grid\": %ld, \"off_color_grid\": %ld, "
         "\"emo_color_grid\": %ld, \"flat_off\": %ld, \"flat_emo\": %ld, \"flat_tone\": %ld, \"precision\": %ld, \"workspace\": %ld, "
         "\"workspace_bytes\": %ld, \"n_rays\": %ld, \"n_on\": %ld, \"m1\": %ld, \"m3\": %ld, \"m3_on\": %ld, \"workspace_used\": %ld, "
         "\"workspace_needed\": %ld, \"alphainv_last\": %ld, \"slot\": %ld}}\n",
         OFF(esr_voxurff_step_t, scene), OFF(esr_voxurff_step_t, mask_density), OFF(esr_voxurff_step_t, mask_cls),
         OFF(esr_voxurff_step_t, sdf_grid), OFF(esr_voxurff_step_t, off_color_grid), OFF(esr_voxurff_step_t, emo_color_grid),
         OFF(esr_voxurff_step_t, flat_off), OFF(esr_voxurff_step_t, flat_emo), OFF(esr_voxurff_step_t, flat_tone),
         OFF(esr_voxurff_step_t, precision), OFF(esr_voxurff_step_t, workspace), OFF(esr_voxurff_step_t, workspace_bytes),
         OFF(esr_voxurff_step_t, n_rays), OFF(esr_voxurff_step_t, n_on), OFF(esr_voxurff_step_t, m1), OFF(esr_voxurff_step_t, m3),
         OFF(esr_voxurff_step_t, m3_on), OFF(esr_voxurff_step_t, workspace_used), OFF(esr_voxurff_step_t, workspace_needed),
         OFF(esr_voxurff_step_t, alphainv_last), OFF(esr_voxurff_step_t, slot));
  return 0;
}

/*
 * esr_b200.h — C ABI of libesr_b200.so: the B200 (sm_100a) render hot path of ESR-NeRF.
 *
 * Drop-in boundary.  Every entry point takes raw DEVICE pointers, sizes,
 * scalars and a cudaStream_t (as void*); the caller owns all memory; nothing here
 * allocates persistent device memory.  Return value: 0 = ok, <0 = error
 * (ESR_ERR_*), message via esr_last_error().  Thread-safe per (device, stream).
 * All citations are relative to the reference tree (ecrireme/ESR-NeRF).
 *
 * Section 1 replaces, one for one, what the reference's pybind module
 * `render_utils_cuda` exports for this path (app/utils/base/cuda/render_utils.cpp:170-184,
 * live entries only) and the torch_scatter.segment_coo call the render functions use.
 * Section 2 is the fused pipeline that the drop-in render modules
 * (esr_nerf_b200.VoxurfF ...) run instead of the reference's ~200 ATen launches.
 */
#ifndef ESR_B200_H
#define ESR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESR_OK 0
#define ESR_ERR_BAD_ARG (-1)
#define ESR_ERR_CUDA (-2)
#define ESR_ERR_CAPACITY (-3)

typedef void *esr_stream_t; /* cudaStream_t */

const char *esr_last_error(void);
int esr_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t esr_launch_count(void);
/*
 * Per-kernel device timing (measurement only; bench.py's roofline block).  esr_stage_timing(1) clears the
 * records and makes every launch site bracket its kernel with CUDA events on the launching stream;
 * esr_stage_timing_report synchronises those events and writes one line per stage, "name launches total_ms\n",
 * into buf (NUL-terminated, truncated to buf_bytes) and returns the number of bytes the full report needs.
 */
int esr_stage_timing(int enable);
int64_t esr_stage_timing_report(char *buf, int64_t buf_bytes);

/* ------------------------------------------------------------------------------------------
 * 1. Native-op replacements (reference-shaped: int64 indices, bool masks)
 * ---------------------------------------------------------------------------------------- */

/*
 * sample_pts_on_rays — render_utils.cpp:74-85, render_utils_kernel.cu:196-242.
 * Two-call protocol replacing the reference's `.item<int>()` host sync (kernel.cu:212):
 *   _count writes N_steps[n], its inclusive cumsum N_cum[n] (both int64), t_min/t_max[n]
 *          and *total (int64, device) = sum(N_steps);
 *   _fill  writes ray_pts[total,3], mask_outbbox[total] (uint8 0/1), ray_id[total],
 *          step_id[total] (int64).  `total` is the value read back by the caller.
 * scratch: >= esr_scan_scratch_bytes(n_rays) bytes.
 */
int esr_sample_pts_on_rays_count(const float *rays_o, const float *rays_d, const float xyz_min[3],
                                 const float xyz_max[3], float near, float far, float stepdist,
                                 int64_t n_rays, int64_t *N_steps, int64_t *N_cum, float *t_min,
                                 float *t_max, int64_t *total, void *scratch, esr_stream_t stream);
int esr_sample_pts_on_rays_fill(const float *rays_o, const float *rays_d, const float xyz_min[3],
                                const float xyz_max[3], float near, float far, float stepdist,
                                int64_t n_rays, const int64_t *N_cum, int64_t total, float *ray_pts,
                                uint8_t *mask_outbbox, int64_t *ray_id, int64_t *step_id,
                                esr_stream_t stream);
int64_t esr_scan_scratch_bytes(int64_t n);

/*
 * alpha2weight / alpha2weight_backward — render_utils.cpp:142-167, kernel.cu:576-707.
 * Same outputs as the reference (weight zero-filled / T one-filled past the early stop,
 * i_end truncated at the stop index).  ray_id must be sorted (as the reference assumes).
 */
int esr_alpha2weight_fwd(const float *alpha, const int64_t *ray_id, int64_t n_pts, int64_t n_rays,
                         float *weight, float *T, float *alphainv_last, int64_t *i_start,
                         int64_t *i_end, esr_stream_t stream);
int esr_alpha2weight_bwd(const float *alpha, const float *weight, const float *T,
                         const float *alphainv_last, const int64_t *i_start, const int64_t *i_end,
                         int64_t n_pts, int64_t n_rays, const float *grad_weights,
                         const float *grad_last, float *grad_alpha, esr_stream_t stream);

/*
 * segment_coo(src, index, out=zeros[n_out,C], reduce="sum") — third-party torch_scatter as used
 * at e.g. app/fine/model/voxurff.py:260-272.  index sorted; out is fully written (no pre-zero).
 * Backward = gather: grad_src[i,:] = grad_out[index[i],:].
 */
int esr_segment_sum_fwd(const float *src, const int64_t *index, int64_t n_pts, int channels,
                        int64_t n_out, float *out, esr_stream_t stream);
int esr_segment_sum_bwd(const float *grad_out, const int64_t *index, int64_t n_pts, int channels,
                        float *grad_src, esr_stream_t stream);

/*
 * total_variation_add_grad (dense / sparse) — total_variation_kernel.cu:14-35,68-98, incl. the
 * reference's axis-weight quirk (wz applied on the k and i axes, wx unused).
 */
int esr_tv_add_grad(const float *param, float *grad, float wx, float wy, float wz, int64_t sz_i,
                    int64_t sz_j, int64_t sz_k, int64_t n_total, int dense_mode, esr_stream_t stream);

/*
 * Dense-grid loss terms of the stage drivers, one launch forward and one backward each.
 *
 * total_variation(v, mask) (app/utils/base/functions.py:34-42; voxurff.py:603-609, voxurfc.py:523-548): v is a
 * [channels][X][Y][Z] view given by element strides (contiguous SDF grid or channels-last colour grid), mask (nullable)
 * u8 [X][Y][Z].  fwd: acc6 (device, f64) <- per-axis sums of |v[p + e_a] - v[p]| over pairs with both voxels in the mask
 * (all channels) and the three pair counts; the loss is (acc[0] / acc[3] + acc[1] / acc[4] + acc[2] / acc[5]) / 3.
 * bwd: grad (same strides as v) += *g_out * scale * d loss / d v; g_out is a DEVICE scalar (no host read).
 */
int esr_grid_tv_fwd(const float *v, const uint8_t *mask, int channels, int64_t X, int64_t Y, int64_t Z, int64_t stride_c,
                    int64_t stride_x, int64_t stride_y, int64_t stride_z, double *acc6, esr_stream_t stream);
int esr_grid_tv_bwd(const float *v, const uint8_t *mask, int channels, int64_t X, int64_t Y, int64_t Z, int64_t stride_c,
                    int64_t stride_x, int64_t stride_y, int64_t stride_z, const double *acc6, const float *g_out,
                    float scale, float *grad, esr_stream_t stream);
/*
 * neus_sdf_gradient (voxurff.py:723-742): out [3][X][Y][Z] = central differences of the SDF grid / 2 / voxel_size, zero on
 * the two boundary faces of each axis.
 * Smooth-gradient term (voxurff.py:610-616): err = conv3(grad_vol) + bias - grad_vol on masked voxels with the fixed
 * 3x3x3 kernel w27 of GradientConv (module.py:180-211; replicate padding; the smoothed volume is detached);
 * fwd: acc2 (device, f64) <- (sum err^2, 3 * masked voxels), err_vol [3][X][Y][Z] <- err (0 outside the mask); the loss is
 * acc[0] / acc[1].  bwd: grad_sdf [X][Y][Z] += *g_out * scale * d loss / d sdf (through the central differences).
 * With acc2 = g_out = NULL the backward is the plain transpose of the central differences: grad_sdf += scale *
 * (cotangent err_vol of esr_sdf_central_gradient's output pulled back to the grid).
 */
int esr_sdf_central_gradient(const float *sdf, int64_t X, int64_t Y, int64_t Z, float voxel_size, float *out,
                             esr_stream_t stream);
int esr_smooth_grad_tv_fwd(const float *grad_vol, const uint8_t *mask, int64_t X, int64_t Y, int64_t Z, const float *w27,
                           float bias, double *acc2, float *err_vol, esr_stream_t stream);
int esr_smooth_grad_tv_bwd(const float *err_vol, int64_t X, int64_t Y, int64_t Z, float voxel_size, const double *acc2,
                           const float *g_out, float scale, float *grad_sdf, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 2. Fused pipeline (int32 packed streams; one warp per ray for ray-ordered stages)
 * ---------------------------------------------------------------------------------------- */

typedef struct esr_scene {
  float xyz_min[3], xyz_max[3];           /* bbox of the sdf / colour grids (voxurff.py:51-52) */
  int32_t gx, gy, gz;                     /* grid dims X,Y,Z (world_size, voxurff.py:543) */
  float mask_xyz_min[3], mask_xyz_max[3]; /* MaskCache bbox (module.py:89-90) */
  int32_t mx, my, mz;                     /* MaskCache (max-pooled) density dims */
  float near, far;                        /* far is 1e9 in the wrappers (voxurff.py:633) */
  float stepdist;                         /* stepsize * voxel_size (voxurff.py:636) */
  float voxel_size;
  float act_shift;                        /* log(1/(1-alpha_init)-1), module.py:101 */
  float mask_thres;                       /* maskcache_thres, 1e-3 */
  float fast_thres;                       /* fastcolor_thres, 1e-4: weight filter (voxurff.py:209, voxurfc.py:214) */
  float s_val;                            /* NeuS inverse std */
  float alpha_thres;                      /* alpha filter before the scan: fastcolor_thres in the fine stage
                                             (voxurff.py:201); negative (= no filter) in the coarse stage */
  float fd_eps;                           /* added to the finite-difference denominator of the 24-tap SDF feature:
                                             0 (voxurff.py:711) or 1e-12 (esrnerf.py:1560) */
  int32_t sdf_tap_manual;                 /* SDF tap of the march stage: 0 = F.grid_sample arithmetic (voxurff.py:671),
                                             1 = differentiable_grid_sample arithmetic (esrnerf.py:1572-1596 over
                                             functions.py:142-309: clamped corner indices, products and sums rounded
                                             separately) */
} esr_scene_t;

/* exclusive scan of int32 counts: out[i] = sum_{j<i} in[j]; out[n] = total (out has n+1 slots) */
int esr_exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, void *scratch,
                           esr_stream_t stream);

/*
 * Stage A/B — march: AABB slab test + per-ray sample count (kernel.cu:12-79), candidate points
 * (kernel.cu:167-194), AABB filter (voxurff.py:649-652), MaskCache test (module.py:104-114) and the
 * trilinear SDF tap (voxurff.py:671) in one warp-per-ray pass.  ray_order (nullable) lets the caller
 * process rays in a permuted order: slot w handles ray ray_order[w]; streams are in slot order.
 *   count: n_steps[slot], cnt_inbox[slot], cnt_mask[slot]                     (int32)
 *   fill : s_ray[M1] (original ray index), s_step[M1], s_sdf[M1] at off_mask[slot] + rank
 */
int esr_march_count(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                    const int32_t *ray_order, int64_t n_rays, const float *mask_density,
                    int32_t *n_steps, int32_t *cnt_inbox, int32_t *cnt_mask, esr_stream_t stream);
int esr_march_fill(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                   const int32_t *ray_order, int64_t n_rays, const float *mask_density,
                   const float *sdf_grid, const int32_t *off_mask, int32_t *s_ray, int32_t *s_step,
                   float *s_sdf, esr_stream_t stream);

/*
 * Same two calls with a keep-bit cache: the count pass stores one ballot word per 32 candidate steps of each ray slot
 * (keep_bits [n_rays][bits_stride] uint32, bits_stride >= ceil(max steps per ray / 32)), the fill pass reads the words
 * instead of repeating the AABB test and the 8-tap MaskCache lookup of every candidate.  Results are identical.
 *
 * mask_cls (nullable): per-cell classes of the MaskCache density grid made by esr_mask_classify ([mx-1][my-1][mz-1]
 * bytes, esr_mask_class_bytes): 1 = every point of the cell passes MaskCache.forward (module.py:104-114), 2 = none
 * does, 0 = evaluate exactly.  The test is monotone in the interpolated density, so a cell whose corners (and its
 * neighbours') lie on one side of the decision density by a margin is decided by one byte load; results are identical.
 * The table must be rebuilt when the density grid, its bbox, act_shift or mask_thres change.
 */
int64_t esr_mask_class_bytes(const esr_scene_t *sc);
int esr_mask_classify(const esr_scene_t *sc, const float *mask_density, uint8_t *cls, esr_stream_t stream);
int esr_march_count_bits(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                         int64_t n_rays, const float *mask_density, int32_t *n_steps, int32_t *cnt_inbox,
                         int32_t *cnt_mask, uint32_t *keep_bits, int bits_stride, const uint8_t *mask_cls,
                         esr_stream_t stream);
int esr_march_fill_bits(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                        int64_t n_rays, const float *mask_density, const float *sdf_grid, const int32_t *off_mask,
                        int32_t *s_ray, int32_t *s_step, float *s_sdf, const uint32_t *keep_bits, int bits_stride,
                        esr_stream_t stream);

/*
 * Stage C/D — NeuS 'interp' alpha (functions.py:72-105), alpha>thr filter (voxurff.py:201),
 * transmittance scan with the reference's sequential float/double recurrence and early stop
 * (kernel.cu:591-603), weight>thr filter (voxurff.py:209).
 *   count: s_alpha[M1] (warp per ray slot, lanes over its M1 segment), then s_T[M1], cnt_shade[slot] and
 *          alphainv_last[ray] (thread per ray slot: 32 rays' dependent chains advance per warp instruction)
 *          (s_T = T of the sample, or -1 when the sample is not part of the scan)
 *   fill : reads s_alpha / s_T, writes h_ray/h_step/h_m1[M3] (int32), h_w/h_sdf[M3] (warp per ray slot)
 */
int esr_alpha_scan_count(const esr_scene_t *sc, const int32_t *ray_order, int64_t n_rays,
                         const int32_t *off_mask, const float *s_sdf, int32_t *cnt_shade,
                         float *alphainv_last, float *s_alpha, float *s_T, esr_stream_t stream);
int esr_alpha_scan_fill(const esr_scene_t *sc, const int32_t *ray_order, int64_t n_rays,
                        const int32_t *off_mask, const int32_t *s_step, const float *s_sdf,
                        const int32_t *off_shade, const float *s_alpha, const float *s_T, int32_t *h_ray,
                        int32_t *h_step, int32_t *h_m1, float *h_w, float *h_sdf,
                        esr_stream_t stream);

/*
 * Stage C' — backward of Alphas2Weights (kernel.cu:653-707) and of the NeuS alpha, then trilinear
 * scatter of dL/dsdf into the dense SDF gradient volume.
 *   g_w_m1[M1]: dL/dweight scattered to the M1 stream (zero where not shaded)
 *   g_last[n_rays]: dL/dalphainv_last
 */
int esr_alpha_scan_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                       const int32_t *ray_order, int64_t n_rays, const int32_t *off_mask,
                       const int32_t *s_ray, const int32_t *s_step, const float *s_sdf,
                       const float *s_alpha, const float *s_T, const float *alphainv_last,
                       const float *g_w_m1, const float *g_last, float *tmp_dprev, float *tmp_dnext,
                       int64_t m1, float *grad_sdf_grid, esr_stream_t stream);
/*
 * `neus_alpha: grad` (functions.py:45-69, selected by voxurff.py:151-154 / esrnerf.py:197-200; every shipped config uses
 * 'interp'): the section-point SDFs of a sample are sdf -+ iter_cos, iter_cos = (viewdir . grad sdf) * dist * 0.5, with
 * grad sdf = sample_sdf_grad's finite differences (voxurff.py:670-676) and dist = stepsize * voxel_size = sc->stepdist.
 *   esr_neus_cos_fwd      : s_cos[M1] for the M1 stream (viewdirs: f32 [n_rays,3], row = the sample's ray)
 *   esr_alpha_scan_count_g: esr_alpha_scan_count with the alphas computed from (s_sdf, s_cos); esr_alpha_scan_fill follows
 *                           unchanged
 *   esr_alpha_scan_bwd_g  : esr_alpha_scan_bwd for those alphas — scatters dL/dsdf of every sample into grad_sdf_grid and
 *                           leaves dL/diter_cos in tmp_dcos[M1]
 *   esr_neus_cos_bwd      : scatter-adds d_cos[M1] through the six finite-difference taps into grad_sdf_grid
 */
int esr_neus_cos_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                     const float *sdf_grid, const int32_t *s_ray, const int32_t *s_step, int64_t m1, float *s_cos,
                     esr_stream_t stream);
int esr_neus_cos_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                     const int32_t *s_ray, const int32_t *s_step, const float *d_cos, int64_t m1, float *grad_sdf_grid,
                     esr_stream_t stream);
int esr_alpha_scan_count_g(const esr_scene_t *sc, const int32_t *ray_order, int64_t n_rays, const int32_t *off_mask,
                           const float *s_sdf, const float *s_cos, int32_t *cnt_shade, float *alphainv_last,
                           float *s_alpha, float *s_T, esr_stream_t stream);
int esr_alpha_scan_bwd_g(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                         int64_t n_rays, const int32_t *off_mask, const int32_t *s_ray, const int32_t *s_step,
                         const float *s_sdf, const float *s_cos, const float *s_alpha, const float *s_T,
                         const float *alphainv_last, const float *g_w_m1, const float *g_last, float *tmp_dsdf,
                         float *tmp_dcos, int64_t m1, float *grad_sdf_grid, esr_stream_t stream);
/*
 * The same mode in the coarse stage (voxurfc.py:171-174, 204-210), where the SDF gradient of a sample is the trilinear
 * tap of the dense central-difference volume grad_vol ([3][X][Y][Z] f32, channels d/dx, d/dy, d/dz — esr_sdf_central_gradient):
 *   esr_neus_cos_vol_fwd: s_cos[M1] from grad_vol;  esr_neus_cos_vol_bwd: scatter-adds d_cos[M1] into g_grad_vol
 *   esr_neus_alpha_bwd_g: esr_neus_alpha_bwd (dL/dalpha given on the M1 stream) for alphas computed from (s_sdf, s_cos):
 *                         dL/dsdf scattered into grad_sdf_grid, dL/diter_cos left in tmp_dcos[M1]
 */
int esr_neus_cos_vol_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                         const float *grad_vol, const int32_t *s_ray, const int32_t *s_step, int64_t m1, float *s_cos,
                         esr_stream_t stream);
int esr_neus_cos_vol_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                         const int32_t *s_ray, const int32_t *s_step, const float *d_cos, int64_t m1, float *g_grad_vol,
                         esr_stream_t stream);
int esr_neus_alpha_bwd_g(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                         int64_t n_rays, const int32_t *off_mask, const int32_t *s_ray, const int32_t *s_step,
                         const float *s_sdf, const float *s_cos, const float *g_alpha_m1, float *tmp_dsdf,
                         float *tmp_dcos, int64_t m1, float *grad_sdf_grid, esr_stream_t stream);
/*
 * Same backward with dL/dalpha given directly on the M1 stream (g_alpha_m1) instead of going through the
 * Alphas2Weights recurrence: the coarse stage recomputes the weights on the shaded samples with the reference-shaped
 * esr_alpha2weight_* ops (voxurfc.py:211-219), so only the NeuS alpha -> sdf part is needed here.
 */
int esr_neus_alpha_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                       int64_t n_rays, const int32_t *off_mask, const int32_t *s_ray, const int32_t *s_step,
                       const float *s_sdf, const float *g_alpha_m1, float *tmp_dprev, float *tmp_dnext, int64_t m1,
                       float *grad_sdf_grid, esr_stream_t stream);

/*
 * Stage E — per shaded sample feature encode (voxurff.py:219-241): 24 multi-scale SDF taps +
 * 12 normal components (voxurff.py:678-721), colour-grid taps (module.py:24-35), positional /
 * view encodings.  Internal column order of the 96-wide row (bf16 or f32):
 *   [off_color 6 | emo_color 6 | sdf 1 | feat 24 | normal 12 | xyz 3 | sin 15 | cos 15 |
 *    view 3 | sin view 3 | cos view 3 | zero pad 5]
 * Colour grids are CHANNELS-LAST in memory ([X][Y][Z][C], torch.channels_last_3d of [1,C,X,Y,Z]).
 * out_is_bf16: 0 -> float rows, row-major [m3][96];
 *              1 -> __nv_bfloat16 rows in the library's TILED MLP-input layout (per 128-row tile:
 *                   [12 feature chunks][128 rows][8]); `feat` must hold esr_mlp_act_rows(m3) rows.  The same holds
 *                   for esr_tonemap_encode_fwd's tfeat (48 columns = 6 chunks).
 *              2 -> (esr_encode_fwd / esr_encode_pbr_fwd) the same tiled layout with __half rows (saturating conversion),
 *                   followed in the same buffer — which must hold 2 * esr_mlp_act_rows(m3) rows — by a second tile set of
 *                   __half: the residual v - fp16(v) of every column.  This is the layer-0 operand of the x2 forward
 *                   chain (esr_mlp_desc_t::precision = 1): esr_mlp_fwd then expects `x` to be such a buffer;
 *                   esr_mlp_bwd reads the first tile set only.
 */
#define ESR_FEAT_DIM 96
#define ESR_FEAT_GRAD_DIM 56 /* columns [0,49) carry gradient; padded to 56 */
int esr_encode_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                   const float *viewdirs, const float *sdf_grid, const float *off_color_grid,
                   const float *emo_color_grid, int color_dim, const int32_t *h_ray,
                   const int32_t *h_step, const float *h_sdf, int64_t m3, void *feat,
                   int out_is_bf16, esr_stream_t stream);
/* d_feat: [m3, ESR_FEAT_GRAD_DIM] f32.  Scatter-adds into the three dense gradient volumes. */
int esr_encode_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                   const float *sdf_grid, int color_dim, const int32_t *h_ray, const int32_t *h_step,
                   int64_t m3, const float *d_feat, float *grad_sdf_grid, float *grad_off_grid,
                   float *grad_emo_grid, esr_stream_t stream);

/*
 * LTS / PDRA stage variants (ESRNeRF, esrnerf.py:729-765, 795-830): the same row at EXPLICIT world points `pts`
 * ([m3,3] f32, nullable: NULL = positions from (h_ray, h_step) as above; with points h_ray / h_step may be NULL and
 * row j reads view direction j), and optionally a second copy of the row, `feat_brdf`, whose colour slot 0 holds the
 * taps of a third grid (`brdf_grid`, the BRDFNet input of esrnerf.py:761-763; both NULL = off).  sc->fd_eps selects the
 * finite-difference denominator.  Backward: d_brdf_color [m3,6] f32 is the cotangent of that slot.
 * save_fd / saved_fd (nullable): f32 [m3,16] — the forward stores the four un-normalised finite-difference SDF
 * gradients of each sample (4 x (z, y, x, pad)); a backward that receives them reads no grid values at all (it
 * otherwise re-gathers the 72 line taps to recompute them).
 */
int esr_encode_pbr_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                       const float *sdf_grid, const float *off_color_grid, const float *emo_color_grid,
                       const float *brdf_grid, int color_dim, const float *pts, const int32_t *h_ray,
                       const int32_t *h_step, const float *h_sdf, int64_t m3, void *feat, void *feat_brdf,
                       int out_is_bf16, float *save_fd, esr_stream_t stream);
int esr_encode_pbr_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *sdf_grid,
                       int color_dim, const float *pts, const int32_t *h_ray, const int32_t *h_step, int64_t m3,
                       const float *d_feat, const float *d_brdf_color, float *grad_sdf_grid, float *grad_off_grid,
                       float *grad_emo_grid, float *grad_brdf_grid, const float *saved_fd, esr_stream_t stream);

/*
 * World positions of stream samples — the reference's ray_pts (kernel.cu:167-194) after its compactions: the origins
 * of the LTS secondary rays (esrnerf.py:576-581) and the points the eps jitters are added to (esrnerf.py:808-813).
 * pts: f32 [m,3], bit-identical to the reference's values.
 */
int esr_sample_points(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *h_ray,
                      const int32_t *h_step, int64_t m, float *pts, esr_stream_t stream);
/*
 * sample_sdf_expgrad (esrnerf.py:1572-1596): SDF value by differentiable_grid_sample (functions.py:142-309) and its
 * analytic gradient d sdf / d xyz (world units, (x,y,z) order) at explicit points.  manual = 0 evaluates the value
 * with F.grid_sample arithmetic instead (sample_sdf_grad at the jittered points, esrnerf.py:819; no gradient output).
 * out_sdf [m] / out_grad [m,3] nullable.  Backward: both outputs are linear in the grid; g_sdf / g_grad (nullable)
 * are scatter-added into grad_sdf_grid (the create_graph=True path of the reference, needed by the normal losses).
 */
int esr_sdf_expgrad_fwd(const esr_scene_t *sc, const float *pts, const float *sdf_grid, int64_t m, int manual,
                        float *out_sdf, float *out_grad, esr_stream_t stream);
int esr_sdf_expgrad_bwd(const esr_scene_t *sc, const float *pts, int64_t m, const float *g_sdf, const float *g_grad,
                        float *grad_sdf_grid, esr_stream_t stream);

/*
 * Light-transport accumulation (esrnerf.py:556-574, 654-677; pbr/functions.py:108-173): for each of n_pts LTS points
 * (unit normal, base colour [3], roughness, metallic, two outgoing directions wo_a = -view / wo_b = -random view) and
 * its n_dirs hemisphere directions `dirs` [n_pts*n_dirs,3] with the marched radiance of the secondary rays
 * (rad_off = sum w*off + env * T_last, nullable; rad_emo = sum w*emo), the Monte-Carlo means
 *   off_hat [2*n_pts,3] = mean_j rad_off_j * R(w_j, wo_v),  reflect [2*n_pts,3] = mean_j rad_emo_j * R(w_j, wo_v)
 * (rows [0,n_pts): wo_a, rows [n_pts, 2 n_pts): wo_b) with the Disney-style reflectance R.  Backward: cotangents of the
 * two outputs -> g_base [n_pts,3], g_rough / g_metal [n_pts], g_rad_off / g_rad_emo [n_pts*n_dirs,3]; normals and
 * directions carry no gradient (esrnerf.py:790: detached normal; pbr/functions.py:9: no_grad sampling).
 * emission (nullable, [n_pts,3]) switches the second output to emo_hat of esrnerf.py:668-677: emission + reflect, or with
 * pdra_mode != 0: emission + stop-gradient(reflect) for a point on an UNCERTAIN ray (umask [n_pts] u8 != 0), reflect
 * alone otherwise; the backward then takes emo_hat's cotangent as g_reflect and also returns g_emission [n_pts,3].
 */
int esr_lts_accumulate_fwd(const float *normal, const float *base, const float *rough, const float *metal,
                           const float *wo_a, const float *wo_b, const float *dirs, const float *rad_off,
                           const float *rad_emo, int64_t n_pts, int n_dirs, float *off_hat, float *reflect,
                           const float *emission, const uint8_t *umask, int pdra_mode, esr_stream_t stream);
int esr_lts_accumulate_bwd(const float *normal, const float *base, const float *rough, const float *metal,
                           const float *wo_a, const float *wo_b, const float *dirs, const float *rad_off,
                           const float *rad_emo, int64_t n_pts, int n_dirs, const float *g_off_hat,
                           const float *g_reflect, float *g_base, float *g_rough, float *g_metal, float *g_rad_off,
                           float *g_rad_emo, const float *emission, const uint8_t *umask, int pdra_mode,
                           float *g_emission, esr_stream_t stream);
/*
 * Hemisphere directions of the light-transport segment (app/utils/pbr/functions.py:10-32): n_dirs directions per point —
 * the normalised Gaussian draw `noise` [n_pts*n_dirs,3] (diffuse_scattering) or, with noise = NULL, the fixed spiral
 * `table` [n_dirs,3] (diffuse_scattering_fib) — mirrored into the hemisphere of the point's normal -> dirs [n_pts*n_dirs,3].
 */
int esr_lts_scatter_dirs(const float *normal, const float *noise, const float *table, int64_t n_pts, int n_dirs,
                         float *dirs, esr_stream_t stream);
/*
 * Spherical-Gaussian environment map (app/utils/pbr/module.py:133-143) on m directions, with its use in the light-transport
 * segment fused in (esrnerf.py:560-566): out [m,3] = add + act(sum_k mus_k exp(lambdas_k (d . lobes_k - 1))) * scale.
 * mus [n_sg,3]; lambdas [n_sg] (already |.|); lobes [n_sg,3] (already unit); n_sg <= 64; act: 1 softplus, 2 relu,
 * 3 abs, 4 exp, 5 sigmoid; scale [m] / add [m,3] nullable.  Backward: g_out -> g_mus / g_lambdas / g_lobes (ACCUMULATED,
 * atomics) and g_scale [m] (iff scale); the cotangent of `add` is g_out itself; directions carry no gradient.
 */
int esr_sg_envmap_fwd(const float *dirs, const float *mus, const float *lambdas, const float *lobes, int n_sg, int act,
                      const float *scale, const float *add, int64_t m, float *out, esr_stream_t stream);
int esr_sg_envmap_bwd(const float *dirs, const float *mus, const float *lambdas, const float *lobes, int n_sg, int act,
                      const float *scale, int64_t m, const float *g_out, float *g_mus, float *g_lambdas, float *g_lobes,
                      float *g_scale, esr_stream_t stream);

/*
 * Coarse-stage feature encode (voxurfc.py:205-249): trilinear tap of the dense central-difference gradient volume
 * `grad_vol` ([1,3,X,Y,Z], channels-first, voxurfc.py:597-616) -> normal = g / (|g| + 1e-5); 12-channel colour-grid
 * taps (channels-last); positional / view encodings.  Row (f32, row-major, 72 columns):
 *   [off_color 12 | emo_color 12 | xyz 3 | sin 15 | cos 15 | view 3 | sin view 3 | cos view 3 | normal 3 | pad 3]
 * Backward: d_feat [m3,72] -> scatter-add into the two colour-grid gradients and the gradient-volume gradient.
 */
#define ESR_COARSE_FEAT_DIM 72
int esr_encode_coarse_fwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *viewdirs,
                          const float *grad_vol, const float *off_color_grid, const float *emo_color_grid,
                          const int32_t *h_ray, const int32_t *h_step, int64_t m3, float *feat,
                          esr_stream_t stream);
/*
 * f32 rows [m][ld] -> the library's tiled 16-bit MLP input rows (96 columns): column c takes source column colmap[c]
 * (device int32 [96]; < 0: zero).  precision 0: bf16 tiles (esr_mlp_act_rows(m) rows); 1: fp16 tiles followed by the
 * fp16 residual tiles (2 * esr_mlp_act_rows(m) rows) — the input of esr_mlp_fwd with esr_mlp_desc_t::precision = 1.
 * The coarse stage feeds its 72-column feature rows to the 96 -> 192 tcgen05 chains through this.
 */
int esr_rows_to_mlp_tiles(const float *src, int64_t m, int ld, const int32_t *colmap, int precision, void *tiles,
                          esr_stream_t stream);
int esr_encode_coarse_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *grad_vol,
                          const int32_t *h_ray, const int32_t *h_step, int64_t m3, const float *d_feat,
                          float *g_grad_vol, float *g_off_grid, float *g_emo_grid, esr_stream_t stream);

/*
 * sample_sdf_grad (voxurff.py:670-676) over the shaded stream: finite-difference SDF gradient from the six axis
 * taps at 1 voxel, world units, (x, y, z) order.  grad_out: f32 [m3,3].  Used by the inference path for the normal
 * map (voxurff.py:421-430).
 */
int esr_sdf_fd_gradient(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const float *sdf_grid,
                        const int32_t *h_ray, const int32_t *h_step, int64_t m3, float *grad_out,
                        esr_stream_t stream);

/*
 * Stage G — combine radiances + tone-map encoding (voxurff.py:243-256, 783-788):
 *   lin = lin_off (+ lin_emo on emission-on rays); tfeat = [lin, sin(lin*2^f), cos(lin*2^f)] padded
 *   to 48 columns (bf16 or f32).
 */
#define ESR_TFEAT_DIM 48
#define ESR_TFEAT_GRAD_DIM 48 /* channel c: columns [16 c, 16 c + 11) used */
int esr_tonemap_encode_fwd(const float *lin_off, const float *lin_emo, const int32_t *h_ray,
                           const int64_t *em_modes, int64_t m3, float *lin, void *tfeat,
                           int out_is_bf16, esr_stream_t stream);
/* d_lin = d_lin_direct + dPE(d_tfeat) */
int esr_tonemap_encode_bwd(const float *lin, const float *d_tfeat, const float *d_lin_direct,
                           int64_t m3, float *d_lin, esr_stream_t stream);

/*
 * Stage H — compositing (replaces segment_coo x k, voxurff.py:259-272): warp per ray slot.
 *   out_a[ray,3] = sum w*a, out_b[ray,3] = sum w*b (b nullable).
 * Backward (sample parallel): d_a = w*c_a[ray], d_b = w*c_b[ray],
 *   g_w[h_m1 ? h_m1[j] : j] = a.c_a[ray] + b.c_b[ray]   (h_m1 nullable: gradient stays in M3 order).
 */
int esr_composite_fwd(const int32_t *ray_order, int64_t n_rays, const int32_t *off_shade,
                      const float *h_w, const float *a, const float *b, float *out_a, float *out_b,
                      esr_stream_t stream);
int esr_composite_bwd(const int32_t *h_ray, const int32_t *h_m1, const float *h_w, const float *a,
                      const float *b, const float *c_a, const float *c_b, int64_t m3, float *d_a,
                      float *d_b, float *g_w_m1, esr_stream_t stream);

/*
 * Stage F — small MLPs on tensor cores (pbr/module.py:6-39; the only dense contraction on the path).
 * Packed parameter image (built by esr_mlp_pack): bf16 weights [out][in_padded] per layer followed by
 * f32 biases; see esr_mlp_desc_t.  Hidden activation ReLU; output activation softplus (1) or
 * sigmoid (2).
 */
typedef struct esr_mlp_desc {
  int32_t k0;      /* padded input width: multiple of 16 (96 radiance, 48 tonemap) */
  int32_t width;   /* hidden width: 192 */
  int32_t n_hidden;/* hidden layers: 3 (radiance nets), 1 (tonemapper) */
  int32_t n_out;   /* real outputs (<= 8: 3 radiance / tone-map / emission, 5 BRDF; padded to 8 rows in the flat copy) */
  int32_t act;     /* 1 softplus, 2 sigmoid */
  int32_t precision; /* forward arithmetic: 0 = bf16 operands (fast, 1e-2 class on outputs only);
                        1 = "x2": every forward operand (inputs, weights, hidden activations) carried as an fp16
                        hi + lo pair, three tcgen05 MMAs per product, fp32 accumulation — pre-activations (and with
                        them the ReLU masks the backward uses) are fp32-class, which is what brings every parameter
                        gradient within 1e-2 of the reference's fp32 nets.  The backward kernels are the same. */
} esr_mlp_desc_t;

int64_t esr_mlp_param_count(const esr_mlp_desc_t *d); /* f32 elements of the flat master copy */
int64_t esr_mlp_image_bytes(const esr_mlp_desc_t *d);
/*
 * Row count (m_total rounded up to the 128-row tile) the activation buffers `hidden` and `d_z` must be sized
 * for.  Both are bf16 [n_hidden][esr_mlp_act_rows(m_total)][width] in a TILED layout private to the library
 * (per 128-row tile: [width/8 feature chunks][128 rows][8]); callers only allocate and pass them through.
 */
int64_t esr_mlp_act_rows(int64_t m_total);
/* bytes of the `hidden` buffer of esr_mlp_fwd / esr_mlp_bwd: the bf16 activations above followed by the ReLU bit
 * masks (32 B per row per layer) that the data-gradient chain reads instead of the activations */
int64_t esr_mlp_hidden_bytes(const esr_mlp_desc_t *d, int64_t m_total);
/* bytes of the `d_z` scratch of esr_mlp_bwd (bf16 cotangents of every layer's pre-activation, tiled) */
int64_t esr_mlp_dz_bytes(const esr_mlp_desc_t *d, int64_t m_total);
/*
 * flat f32 master copy layout: for each layer l: W_l [out_l][in_l_padded] then b_l [out_l]
 * (output layer padded to 8 rows).  esr_mlp_pack converts it to the bf16 kernel image: every weight matrix and
 * its transpose in the shared-memory operand layout of the tcgen05 kernels, plus the f32 biases.
 */
int esr_mlp_pack(const esr_mlp_desc_t *d, const float *flat_params, void *image, esr_stream_t stream);
/*
 * Forward over rows [row_begin,row_end) of x (bf16, k0 columns, TILED layout as written by esr_encode_fwd /
 * esr_tonemap_encode_fwd with out_is_bf16 = 1; with d->precision = 1: out_is_bf16 = 2, sized for m_total rows, and
 * only the 96 -> 192 x 3 shape — the kernel then runs on CTA pairs, cta_group::2).  y: f32 [*,n_out] (activated).
 * hidden (nullable): esr_mlp_hidden_bytes(d, m_total) bytes; post-ReLU activations + ReLU masks saved for backward,
 * only for rows >= save_row_begin (rows the caller will never back-propagate through — e.g. the off net on
 * emission-on rays, which see it through a stop-gradient, voxurff.py:243-254 — need not be stored).
 */
int esr_mlp_fwd(const esr_mlp_desc_t *d, const void *image, const void *x, int64_t row_begin,
                int64_t row_end, int64_t m_total, float *y, void *hidden, int64_t save_row_begin,
                esr_stream_t stream);
/*
 * Backward over rows [row_begin,row_end): d_y is dL/dy (post-activation), y the saved outputs.
 *   d_x (nullable): f32 [*, dx_cols] gets dL/dx for the first dx_cols input columns
 *                   (accumulate != 0 adds to existing values).
 *   d_z: scratch of esr_mlp_dz_bytes(d, m_total) bytes; d_z_out (nullable): f32 [m_total][8], receives the output
 *        layer's pre-activation cotangent (diagnostic)
 *   grad_flat: f32 flat gradient (same layout as flat_params), ACCUMULATED into (atomics).
 */
int esr_mlp_bwd(const esr_mlp_desc_t *d, const void *image, const void *x, const float *y,
                const float *d_y, int64_t row_begin, int64_t row_end, int64_t m_total,
                const void *hidden, void *d_z, float *d_z_out, float *d_x, int dx_cols,
                int accumulate, float *grad_flat, esr_stream_t stream);
/*
 * The weight-gradient half of esr_mlp_bwd on its own: grad_flat += dW / db of every layer from (x, hidden) and the d_z
 * scratch a preceding esr_mlp_bwd call with grad_flat = NULL (data gradient only) left behind for the same rows.
 * Splitting the two lets a caller run what depends on d_x — the encode backward, and with it the start of the
 * colour-grid gradient exchange of a multi-GPU step — before the weight-gradient GEMMs, which then overlap the collective.
 */
int esr_mlp_bwd_weights(const esr_mlp_desc_t *d, const void *x, int64_t row_begin, int64_t row_end, int64_t m_total,
                        const void *hidden, const void *d_z, float *grad_flat, esr_stream_t stream);

/*
 * Fused tone-map net (apply_tonemapper, voxurff.py:783-788: PE(5) of the linear radiance -> 33 -> 192 -> 3 sigmoid,
 * pbr/module.py:24-39).  d must be the tone-map shape (k0 48, one hidden layer).  The kernels compute the positional
 * encoding themselves and keep every intermediate on the SM:
 *   fwd: lin f32 [m,3] -> rgb f32 [m,3]                                   (24 B per row of HBM traffic)
 *   bwd: (lin, rgb, d_rgb[, d_lin_direct]) -> d_lin [m,3] (= d_lin_direct + the gradient through the net) and
 *        grad_flat += weight / bias gradients (flat master layout), the hidden activations recomputed in the kernel.
 */
int esr_tonemap_mlp_fwd(const esr_mlp_desc_t *d, const void *image, const float *lin, int64_t m, float *rgb,
                        esr_stream_t stream);
int esr_tonemap_mlp_bwd(const esr_mlp_desc_t *d, const void *image, const float *lin, const float *rgb,
                        const float *d_rgb, const float *d_lin_direct, int64_t m, float *d_lin, float *grad_flat,
                        esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * One call = one fine-stage render step: VoxurfF.forward_training (app/fine/model/voxurff.py:177-278)
 * and its backward, run on the stage entry points above in the order esr_nerf_b200/fused.py runs them, for hosts that
 * are not Python.  Caller-owned everything: fill the input fields of the esr_voxurff_step_t struct and zero the rest, pass ONE
 * workspace (esr_render_voxurff_workspace_bytes for given bounds on the two stream sizes; the forward fails with
 * ESR_ERR_CAPACITY and sets workspace_needed when it is too small), then
 *   esr_render_voxurff_fwd: rays_o / rays_d / viewdirs f32 [n,3], em_modes i64 [n] -> srgb/rgb [n,3], lin/rgb [n,3],
 *                           etc/alphainv_cum [n] (the dict of voxurff.py:273-278; etc/white_bg is alphainv_cum[:, None]);
 *                           synchronises the stream twice (the two data-dependent stream sizes, as the reference does);
 *   esr_render_voxurff_bwd: cotangents of the three outputs -> gradients ACCUMULATED into the caller's dense grid
 *                           gradient volumes (layouts of the grids) and flat MLP gradients (layout of esr_mlp_pack's
 *                           flat_params; esr_mlp_param_count floats each).  The step object, the workspace, the rays and
 *                           alphainv_last must be unchanged since the forward.
 * flat_* = the f32 master copies esr_mlp_pack consumes (radiance nets: k0 96, width 192, 3 hidden, 3 outputs softplus;
 * tone mapper: k0 48, width 192, 1 hidden, 3 outputs sigmoid).  precision: esr_mlp_desc_t::precision of the three nets.
 * ---------------------------------------------------------------------------------------- */
typedef struct esr_voxurff_step {
  /* inputs */
  esr_scene_t scene;
  const float *mask_density;    /* MaskCache (max-pooled) density [mx][my][mz] */
  const uint8_t *mask_cls;      /* nullable: esr_mask_classify table */
  const float *sdf_grid;        /* [gx][gy][gz] */
  const float *off_color_grid;  /* channels-last [gx][gy][gz][6] */
  const float *emo_color_grid;
  const float *flat_off, *flat_emo, *flat_tone;
  int32_t precision;            /* 0 bf16 operands, 1 x2 */
  void *workspace;
  int64_t workspace_bytes;
  /* outputs of the forward */
  int64_t n_rays, n_on, m1, m3, m3_on;      /* rays, emission-on rays, stream sizes */
  int64_t workspace_used, workspace_needed; /* bytes carved by the forward / needed by forward + backward */
  /* private to the library */
  const float *alphainv_last;
  void *slot[32];
} esr_voxurff_step_t;
int64_t esr_render_voxurff_workspace_bytes(const esr_scene_t *sc, int64_t n_rays, int64_t m1_max, int64_t m3_max,
                                           int precision);
int esr_render_voxurff_fwd(esr_voxurff_step_t *step, const float *rays_o, const float *rays_d, const float *viewdirs,
                           const int64_t *em_modes, int64_t n_rays, float *rgb_marched, float *lin_marched,
                           float *alphainv_last, esr_stream_t stream);
int esr_render_voxurff_bwd(esr_voxurff_step_t *step, const float *rays_o, const float *rays_d,
                           const float *d_rgb_marched, const float *d_lin_marched, const float *d_alphainv_last,
                           float *grad_sdf_grid, float *grad_off_grid, float *grad_emo_grid, float *grad_flat_off,
                           float *grad_flat_emo, float *grad_flat_tone, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step: app/utils/optimizer.py:63-228 — dense Adam with an optional per-voxel learning
 * rate volume (set_pervoxel_lr, optimizer.py:107-109), amsgrad off (never enabled by the reference).  One fused pass;
 * `step` is the 1-based step count of this parameter (bias corrections are evaluated on the host in double like the
 * reference's Python floats).  All tensors f32 contiguous with n elements and 16-byte aligned; per_lr nullable.
 * ---------------------------------------------------------------------------------------- */
int esr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, const float *per_lr, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                  esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Gradient exchange: pack / unpack either side of the one collective on the path, the gradient
 * all-reduce of the ray-sharded step.  The reference has no distributed code (cfg/__init__.yaml:24); what is replaced
 * is a dense all-reduce of every grid gradient (0.83-1.2 GB at 256^3).  `volumes[j]` is a dense gradient volume in its
 * parameter's memory layout [V][channels[j]] (f32, channels innermost, 8-byte aligned), `idx[k]` (int32, sorted) the
 * voxels that can carry gradient.  The packed buffer holds one [k][channels[j]] block per volume, each block starting
 * on an even float offset; esr_grad_pack_floats returns its length in floats (-1 on bad arguments).  Pack copies
 * volume -> buffer, unpack copies buffer -> volume (voxels outside idx are not touched).
 * ---------------------------------------------------------------------------------------- */
#define ESR_MAX_GRAD_VOLUMES 4
int64_t esr_grad_pack_floats(const int32_t *channels, int n_volumes, int64_t k);
int esr_grad_pack(void *const *volumes, const int32_t *channels, int n_volumes, const int32_t *idx, int64_t k,
                  float *buf, esr_stream_t stream);
int esr_grad_unpack(void *const *volumes, const int32_t *channels, int n_volumes, const int32_t *idx, int64_t k,
                    const float *buf, esr_stream_t stream);
/*
 * Touched-block map for gradients without a static support set (LTS / PDRA stage: eps-jittered samples and secondary
 * rays, esrnerf.py:576-652, 807-830): flags[b] (int32 [gx/ex * gy/ey * gz/ez], block index ((bx * By) + by) * Bz + bz)
 * is set to 1 where any of the volumes holds a non-zero float inside block b of (ex, ey, ez) voxels, else 0.  The edges
 * must divide the grid extents.  The ranks OR-reduce the maps and exchange the voxels of the union with
 * esr_grad_pack / esr_grad_unpack (esr_nerf_b200/dist.py:TouchedBlockCompactor).
 */
int esr_grad_block_flags(void *const *volumes, const int32_t *channels, int n_volumes, int32_t gx, int32_t gy, int32_t gz,
                         int32_t ex, int32_t ey, int32_t ez, int32_t *flags, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3. Alphamask stage (DVGO, app/coarse/model/dvgo.py:140-288): dense [N x S] sampling, density / colour grids only
 * ---------------------------------------------------------------------------------------- */
typedef struct esr_dvgo_scene {
  float xyz_min[3], xyz_max[3]; /* dvgo.py:27-28 */
  int32_t gx, gy, gz;           /* world_size */
  float near, far;              /* t clamp of the slab test (dvgo.py:152-158) */
  float stepdist;               /* stepsize * voxel_size (dvgo.py:167) */
  float interval;               /* stepsize (dvgo.py:184) */
  float act_shift;              /* log(1/(1-alpha_init)-1) (dvgo.py:37) */
} esr_dvgo_scene_t;

/*
 * DVGO.forward_training (dvgo.py:174-214).  jitter: f32 [n_rays] per-ray sample offset in [0,1) (the reference's
 * torch.rand_like draw, dvgo.py:163; nullable = 0).  Grids: density [1,1,X,Y,Z], colours [1,3,X,Y,Z], channels-first.
 * Outputs (all f32): alphainv_cum [N,S+1], weights [N,S], raw_rgb [N,S,3], rgb [N,3]; alpha [N,S], raw_off / raw_emo
 * [N,S,3] (per-grid sigmoid colours) are saved for backward.
 */
int esr_dvgo_fwd(const esr_dvgo_scene_t *sc, const float *rays_o, const float *rays_d, const float *jitter,
                 const int64_t *em_modes, const float *density, const float *off_color, const float *emo_color,
                 int64_t n_rays, int n_samples, float *alpha, float *raw_off, float *raw_emo, float *alphainv_cum,
                 float *weights, float *raw_rgb, float *rgb, esr_stream_t stream);
/* DVGO.forward_evaluate (dvgo.py:216-263): off / emo / on = off + emo colour maps and depth = sum w |o - p| */
int esr_dvgo_eval(const esr_dvgo_scene_t *sc, const float *rays_o, const float *rays_d, const float *density,
                  const float *off_color, const float *emo_color, int64_t n_rays, int n_samples, float *alpha,
                  float *raw_off, float *raw_emo, float *alphainv_cum, float *weights, float *off_rgb, float *emo_rgb,
                  float *on_rgb, float *depth, esr_stream_t stream);
/*
 * Backward of esr_dvgo_fwd: cotangents of (alphainv_cum, weights, raw_rgb, rgb) -> scatter-add into the three grid
 * gradients.  d_alpha [N,S] and d_raw [N,S,3] are scratch.
 */
int esr_dvgo_bwd(const esr_dvgo_scene_t *sc, const float *rays_o, const float *rays_d, const float *jitter,
                 const int64_t *em_modes, const float *density, int64_t n_rays, int n_samples, const float *alpha,
                 const float *raw_off, const float *raw_emo, const float *alphainv_cum, const float *raw_rgb,
                 const float *g_cum, const float *g_weights, const float *g_raw_rgb, const float *g_rgb, float *d_alpha,
                 float *d_raw, float *grad_density, float *grad_off_color, float *grad_emo_color, esr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ESR_B200_H */

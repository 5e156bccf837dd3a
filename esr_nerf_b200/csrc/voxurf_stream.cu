// voxurf_stream.cu — ray-ordered stages of the fused render path: march (AABB + MaskCache + SDF tap),
// NeuS alpha + transmittance scan + the two threshold compactions, and their backward.
// One warp owns one ray slot: candidate steps / stream samples are processed 32 at a time so every
// global access of a warp is a contiguous run; compaction is ballot + popc, no atomics, and the
// stream order equals the reference's (ray-major, step-minor) order.
#include "common.cuh"

using namespace esr;

static inline unsigned ray_blocks(int64_t n_rays) {
  const int64_t want = (n_rays + 7) / 8;  // 8 warps per block, one warp per ray
  const int64_t cap = (int64_t)num_sms() * 64;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
// MaskCache cell classes.  MaskCache.forward (module.py:104-114) is trilinear(density) -> softplus -> 1 - exp(-.) >= thres,
// i.e. sigmoid(d + shift) >= thres: monotone in the interpolated density d, and d is a convex combination of the 8
// corners of the cell.  A cell whose corners (and those of its 26 neighbours, so that the cheap cell lookup below may
// be off by one cell) all lie above the decision density by a margin keeps every point in it; all below drops every
// point.  The march test then costs one byte load for those cells and the exact 8-tap evaluation only near the
// occupancy boundary — the boolean results are identical by construction (the margin is ~150x the float rounding
// of the exact chain near the threshold; non-finite densities and out-of-grid neighbours force the exact path).
// cls[(i * (my-1) + j) * (mz-1) + k]: 0 = evaluate exactly, 1 = keep, 2 = drop.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_mask_classify(const __grid_constant__ esr_scene_t sc, const float *__restrict__ dens, float d_star,
                    uint8_t *__restrict__ cls) {
  const int cx = sc.mx - 1, cy = sc.my - 1, cz = sc.mz - 1;
  const int64_t total = (int64_t)cx * cy * cz;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(c % cz), j = (int)((c / cz) % cy), i = (int)(c / ((int64_t)cz * cy));
    float lo = 3.0e38f, hi = -3.0e38f;
    bool finite = true;
    for (int a = i - 1; a <= i + 2; ++a)
      for (int b = j - 1; b <= j + 2; ++b)
        for (int d = k - 1; d <= k + 2; ++d) {
          // zeros padding of grid_sample outside the grid
          const float v = in_grid(a, b, d, sc.mx, sc.my, sc.mz) ? __ldg(dens + ((int64_t)a * sc.my + b) * sc.mz + d) : 0.f;
          finite = finite && (fabsf(v) <= 1.0e30f);
          lo = fminf(lo, v);
          hi = fmaxf(hi, v);
        }
    const float margin = 0.01f + 1e-5f * fmaxf(fabsf(lo), fabsf(hi));
    uint8_t r = 0;
    if (finite && lo > d_star + margin) r = 1;
    else if (finite && hi < d_star - margin) r = 2;
    cls[c] = r;
  }
}

extern "C" int64_t esr_mask_class_bytes(const esr_scene_t *sc) {
  if (!sc || sc->mx < 2 || sc->my < 2 || sc->mz < 2) return 0;
  return (int64_t)(sc->mx - 1) * (sc->my - 1) * (sc->mz - 1);
}

extern "C" int esr_mask_classify(const esr_scene_t *sc, const float *mask_density, uint8_t *cls, esr_stream_t stream) {
  ESR_CHECK_ARG(sc && mask_density && cls && sc->mx >= 2 && sc->my >= 2 && sc->mz >= 2);
  ESR_CHECK_ARG(sc->mask_thres > 0.f && sc->mask_thres < 1.f);
  // sigmoid(d + shift) >= thres  <=>  d >= logit(thres) - shift
  const double t = (double)sc->mask_thres;
  const float d_star = (float)(log(t / (1.0 - t)) - (double)sc->act_shift);
  const int64_t total = esr_mask_class_bytes(sc);
  ESR_STAGE("k_mask_classify", (cudaStream_t)stream);
  k_mask_classify<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(*sc, mask_density, d_star, cls);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// cheap cell lookup: index = (p - min) * (size - 1) / (max - min) with one multiply per axis (scale precomputed per
// thread); it may differ from world_to_index's separately rounded expression by a fraction of a cell — covered by the
// neighbour dilation of the class table.  Points outside the mask grid's cells take the exact path.
struct MaskCls {
  const uint8_t *cls;
  float sx, sy, sz;
};
ESR_D MaskCls mask_cls_setup(const esr_scene_t &sc, const uint8_t *cls) {
  MaskCls m;
  m.cls = cls;
  m.sx = (float)(sc.mx - 1) / (sc.mask_xyz_max[0] - sc.mask_xyz_min[0]);
  m.sy = (float)(sc.my - 1) / (sc.mask_xyz_max[1] - sc.mask_xyz_min[1]);
  m.sz = (float)(sc.mz - 1) / (sc.mask_xyz_max[2] - sc.mask_xyz_min[2]);
  return m;
}
// class of the cell a point falls in: 1 keep, 2 drop, 0 = evaluate MaskCache.forward exactly
ESR_D int mask_cls_lookup(const esr_scene_t &sc, const MaskCls &m, float px, float py, float pz) {
  if (!m.cls) return 0;
  const float fx = (px - sc.mask_xyz_min[0]) * m.sx, fy = (py - sc.mask_xyz_min[1]) * m.sy,
              fz = (pz - sc.mask_xyz_min[2]) * m.sz;
  const int i = (int)floorf(fx), j = (int)floorf(fy), k = (int)floorf(fz);
  if ((unsigned)i < (unsigned)(sc.mx - 1) && (unsigned)j < (unsigned)(sc.my - 1) && (unsigned)k < (unsigned)(sc.mz - 1))
    return __ldg(m.cls + ((int64_t)i * (sc.my - 1) + j) * (sc.mz - 1) + k);
  return 0;
}
ESR_D bool mask_keep_cls(const esr_scene_t &sc, const MaskCls &m, const float *__restrict__ mask_density, float px,
                         float py, float pz) {
  const int c = mask_cls_lookup(sc, m, px, py, pz);
  return c ? c == 1 : mask_keep(sc, mask_density, px, py, pz);
}

// ---------------------------------------------------------------------------------------------
// Stage A/B: march
// ---------------------------------------------------------------------------------------------
// keep_bits (nullable): [n_rays][bits_stride] uint32, one ballot word per 32 candidate steps of a ray slot.  The count
// pass writes it, the fill pass reads it instead of repeating the AABB test and the 8-tap MaskCache lookup of every
// candidate (chunks beyond bits_stride, if any, are recomputed).
template <bool FILL>
__global__ void __launch_bounds__(256)
    k_march(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
            const float *__restrict__ rays_d, const int32_t *__restrict__ ray_order, int64_t n_rays,
            const float *__restrict__ mask_density, const float *__restrict__ sdf_grid,
            int32_t *__restrict__ n_steps, int32_t *__restrict__ cnt_inbox, int32_t *__restrict__ cnt_mask,
            const int32_t *__restrict__ off_mask, int32_t *__restrict__ s_ray, int32_t *__restrict__ s_step,
            float *__restrict__ s_sdf, uint32_t *__restrict__ keep_bits, int bits_stride,
            const uint8_t *__restrict__ mask_cls) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  const unsigned lt = lanemask_lt();
  const MaskCls mc = mask_cls_setup(sc, mask_cls);
  for (int64_t slot = warp; slot < n_rays; slot += nwarps) {
    const int r = ray_order ? ray_order[slot] : (int)slot;
    const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
    const int base = FILL ? off_mask[slot] : 0;
    uint32_t *bits = keep_bits ? keep_bits + slot * (int64_t)bits_stride : nullptr;
    int c_in = 0, c_mask = 0;
    if (!FILL) {
      // count pass: two 32-candidate chunks per iteration, both class-table bytes requested before either is used
      // (the loop is a chain of dependent L1/L2 loads; one chunk at a time left the issue slots idle)
      for (int k0 = 0; k0 < s.n; k0 += 64) {
        const int ka = k0 + (int)lane, kb = ka + 32;
        float ax, ay, az, bx, by, bz;
        ray_point(s, sc.stepdist, ka, ax, ay, az);
        ray_point(s, sc.stepdist, kb, bx, by, bz);
        const bool in_a = (ka < s.n) && !out_bbox(sc.xyz_min, sc.xyz_max, ax, ay, az);
        const bool in_b = (kb < s.n) && !out_bbox(sc.xyz_min, sc.xyz_max, bx, by, bz);
        const int ca = in_a ? mask_cls_lookup(sc, mc, ax, ay, az) : 2;
        const int cb = in_b ? mask_cls_lookup(sc, mc, bx, by, bz) : 2;
        const bool keep_a = ca ? ca == 1 : mask_keep(sc, mask_density, ax, ay, az);
        const bool keep_b = cb ? cb == 1 : mask_keep(sc, mask_density, bx, by, bz);
        const unsigned bal_a = __ballot_sync(FULL, keep_a), bal_b = __ballot_sync(FULL, keep_b);
        c_in += __popc(__ballot_sync(FULL, in_a)) + __popc(__ballot_sync(FULL, in_b));
        c_mask += __popc(bal_a) + __popc(bal_b);
        const int chunk = k0 >> 5;
        if (bits && lane == 0) {
          if (chunk < bits_stride) bits[chunk] = bal_a;
          if (chunk + 1 < bits_stride) bits[chunk + 1] = bal_b;
        }
      }
    }
    for (int k0 = 0; FILL && k0 < s.n; k0 += 32) {
      const int k = k0 + (int)lane;
      const int chunk = k0 >> 5;
      float px, py, pz;
      ray_point(s, sc.stepdist, k, px, py, pz);
      unsigned bal;
      if (FILL && bits && chunk < bits_stride) {
        bal = bits[chunk];                              // warp-uniform load
      } else {
        const bool inb = (k < s.n) && !out_bbox(sc.xyz_min, sc.xyz_max, px, py, pz);
        const bool keep = inb && mask_keep_cls(sc, mc, mask_density, px, py, pz);
        bal = __ballot_sync(FULL, keep);
        if (!FILL) {
          c_in += __popc(__ballot_sync(FULL, inb));
          if (bits && chunk < bits_stride && lane == 0) bits[chunk] = bal;
        }
      }
      if (FILL && ((bal >> lane) & 1u)) {
        const int pos = base + c_mask + __popc(bal & lt);
        s_ray[pos] = r;
        s_step[pos] = k;
        s_sdf[pos] = sc.sdf_tap_manual
                         ? tap1_manual_world(sdf_grid, sc.gx, sc.gy, sc.gz, sc.xyz_min, sc.xyz_max, px, py, pz)
                         : tap1_world(sdf_grid, sc.gx, sc.gy, sc.gz, sc.xyz_min, sc.xyz_max, px, py, pz);
      }
      c_mask += __popc(bal);
    }
    if (!FILL && lane == 0) {
      n_steps[slot] = s.n;
      cnt_inbox[slot] = c_in;
      cnt_mask[slot] = c_mask;
    }
  }
}

static int check_scene(const esr_scene_t *sc) {
  ESR_CHECK_ARG(sc != nullptr);
  ESR_CHECK_ARG(sc->gx > 0 && sc->gy > 0 && sc->gz > 0 && sc->mx > 0 && sc->my > 0 && sc->mz > 0);
  ESR_CHECK_ARG(sc->stepdist > 0.f && sc->voxel_size > 0.f);
  return ESR_OK;
}

static int march_count_impl(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                            int64_t n_rays, const float *mask_density, int32_t *n_steps, int32_t *cnt_inbox,
                            int32_t *cnt_mask, uint32_t *keep_bits, int bits_stride, const uint8_t *mask_cls,
                            esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && n_rays < (1ll << 31));
  if (n_rays == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && mask_density && n_steps && cnt_inbox && cnt_mask);
  ESR_STAGE("k_march_count", (cudaStream_t)stream);
  k_march<false><<<ray_blocks(n_rays), 256, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, ray_order, n_rays,
                                                                       mask_density, nullptr, n_steps, cnt_inbox,
                                                                       cnt_mask, nullptr, nullptr, nullptr, nullptr,
                                                                       keep_bits, bits_stride, mask_cls);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_march_count(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                               const int32_t *ray_order, int64_t n_rays, const float *mask_density, int32_t *n_steps,
                               int32_t *cnt_inbox, int32_t *cnt_mask, esr_stream_t stream) {
  return march_count_impl(sc, rays_o, rays_d, ray_order, n_rays, mask_density, n_steps, cnt_inbox, cnt_mask, nullptr, 0,
                          nullptr, stream);
}

extern "C" int esr_march_count_bits(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                    const int32_t *ray_order, int64_t n_rays, const float *mask_density,
                                    int32_t *n_steps, int32_t *cnt_inbox, int32_t *cnt_mask, uint32_t *keep_bits,
                                    int bits_stride, const uint8_t *mask_cls, esr_stream_t stream) {
  ESR_CHECK_ARG(!keep_bits || bits_stride > 0);
  return march_count_impl(sc, rays_o, rays_d, ray_order, n_rays, mask_density, n_steps, cnt_inbox, cnt_mask, keep_bits,
                          bits_stride, mask_cls, stream);
}

static int march_fill_impl(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                           int64_t n_rays, const float *mask_density, const float *sdf_grid, const int32_t *off_mask,
                           int32_t *s_ray, int32_t *s_step, float *s_sdf, const uint32_t *keep_bits, int bits_stride,
                           esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && n_rays < (1ll << 31));
  if (n_rays == 0) return ESR_OK;
  // s_* may be NULL when the stream is empty (M1 == 0): they are only written for surviving samples
  ESR_CHECK_ARG(rays_o && rays_d && mask_density && sdf_grid && off_mask);
  ESR_STAGE("k_march_fill", (cudaStream_t)stream);
  k_march<true><<<ray_blocks(n_rays), 256, 0, (cudaStream_t)stream>>>(*sc, rays_o, rays_d, ray_order, n_rays,
                                                                      mask_density, sdf_grid, nullptr, nullptr,
                                                                      nullptr, off_mask, s_ray, s_step, s_sdf,
                                                                      const_cast<uint32_t *>(keep_bits), bits_stride,
                                                                      nullptr);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_march_fill(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                              const int32_t *ray_order, int64_t n_rays, const float *mask_density,
                              const float *sdf_grid, const int32_t *off_mask, int32_t *s_ray, int32_t *s_step,
                              float *s_sdf, esr_stream_t stream) {
  return march_fill_impl(sc, rays_o, rays_d, ray_order, n_rays, mask_density, sdf_grid, off_mask, s_ray, s_step, s_sdf,
                         nullptr, 0, stream);
}

extern "C" int esr_march_fill_bits(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                   const int32_t *ray_order, int64_t n_rays, const float *mask_density,
                                   const float *sdf_grid, const int32_t *off_mask, int32_t *s_ray, int32_t *s_step,
                                   float *s_sdf, const uint32_t *keep_bits, int bits_stride, esr_stream_t stream) {
  ESR_CHECK_ARG(!keep_bits || bits_stride > 0);
  return march_fill_impl(sc, rays_o, rays_d, ray_order, n_rays, mask_density, sdf_grid, off_mask, s_ray, s_step, s_sdf,
                         keep_bits, bits_stride, stream);
}

// ---------------------------------------------------------------------------------------------
// Stage C/D: NeuS alpha -> alpha filter -> exact sequential transmittance -> weight filter
//
// Three kernels, each in the shape its work has:
//   k_neus_alpha     warp per ray slot, lanes over its M1 segment (coalesced): alpha of every sample, T preset to -1
//   k_transmittance  THREAD per ray slot: the reference's sequential float/double recurrence with early stop
//                    (kernel.cu:591-603) is a dependent chain per ray; a warp now advances 32 different rays' chains
//                    at once instead of replaying one chain on 32 lanes (32x fewer FP64 / conversion instructions —
//                    the MIO-throttled part of the previous warp-per-ray version).  A lane walks consecutive
//                    addresses, so each 128-byte line it touches serves its next 32 samples out of L1.
//   k_shade_compact  warp per ray slot: weights, weight filter, ballot/popc compaction into the M3 stream (coalesced)
// ---------------------------------------------------------------------------------------------
// GRAD: `neus_alpha: grad` (functions.py:45-69) — the section-point SDFs come from s_cos[M1] (k_neus_cos_fwd) instead of
// the neighbouring samples; the <false> instantiation is the 'interp' kernel unchanged.
template <bool GRAD>
__global__ void __launch_bounds__(256)
    k_neus_alpha(const __grid_constant__ esr_scene_t sc, int64_t n_rays, const int32_t *__restrict__ off_mask,
                 const float *__restrict__ s_sdf, const float *__restrict__ s_cos, float *__restrict__ s_alpha,
                 float *__restrict__ s_T) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t slot = warp; slot < n_rays; slot += nwarps) {
    const int s = off_mask[slot], e = off_mask[slot + 1];
    for (int i = s + (int)lane; i < e; i += 32) {
      const float sd = __ldg(s_sdf + i);
      float pc, nc;
      if constexpr (GRAD) {
        s_alpha[i] = neus_alpha_grad(sd, __ldg(s_cos + i), sc.s_val, pc, nc);
      } else {
        const bool has_prev = i > s, has_next = i + 1 < e;
        const float sp = has_prev ? __ldg(s_sdf + i - 1) : 0.f;
        const float sn = has_next ? __ldg(s_sdf + i + 1) : 0.f;
        s_alpha[i] = neus_alpha(sd, sp, sn, has_prev, has_next, sc.s_val, pc, nc);
      }
      s_T[i] = -1.f;   // "not part of the scan" until k_transmittance visits the sample
    }
  }
}

__global__ void __launch_bounds__(128)
    k_transmittance(const __grid_constant__ esr_scene_t sc, const int32_t *__restrict__ ray_order, int64_t n_rays,
                    const int32_t *__restrict__ off_mask, const float *__restrict__ s_alpha, float *__restrict__ s_T,
                    int32_t *__restrict__ cnt_shade, float *__restrict__ alphainv_last) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_rays) return;
  const int r = ray_order ? ray_order[slot] : (int)slot;
  const int s = off_mask[slot], e = off_mask[slot + 1];
  float Tc = 1.f;
  int n_shade = 0;
  for (int i = s; i < e; ++i) {
    const float a = s_alpha[i];
    if (!(a > sc.alpha_thres)) continue;            // voxurff.py:201 (coarse stage: alpha_thres = -1, no filter)
    s_T[i] = Tc;                                    // kernel.cu:591-601 in order over the surviving samples
    n_shade += (__fmul_rn(Tc, a) > sc.fast_thres) ? 1 : 0;   // voxurff.py:209
    Tc = (float)((1. - (double)a) * (double)Tc);
    if ((double)Tc < 1e-3) break;
  }
  cnt_shade[slot] = n_shade;
  alphainv_last[r] = Tc;
}

__global__ void __launch_bounds__(256)
    k_shade_compact(const __grid_constant__ esr_scene_t sc, const int32_t *__restrict__ ray_order, int64_t n_rays,
                    const int32_t *__restrict__ off_mask, const int32_t *__restrict__ s_step,
                    const float *__restrict__ s_sdf, const int32_t *__restrict__ off_shade,
                    const float *__restrict__ s_alpha, const float *__restrict__ s_T, int32_t *__restrict__ h_ray,
                    int32_t *__restrict__ h_step, int32_t *__restrict__ h_m1, float *__restrict__ h_w,
                    float *__restrict__ h_sdf) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  const unsigned lt = lanemask_lt();
  for (int64_t slot = warp; slot < n_rays; slot += nwarps) {
    const int r = ray_order ? ray_order[slot] : (int)slot;
    const int s = off_mask[slot], e = off_mask[slot + 1];
    int pos0 = off_shade[slot];
    const int pos_end = off_shade[slot + 1];
    for (int base = s; base < e && pos0 < pos_end; base += 32) {
      const int i = base + (int)lane;
      float T = -1.f, w = 0.f;
      if (i < e) {
        T = __ldg(s_T + i);
        if (T >= 0.f) w = __fmul_rn(T, __ldg(s_alpha + i));
      }
      const bool f1 = (T >= 0.f) && (w > sc.fast_thres);
      const unsigned b1 = __ballot_sync(FULL, f1);
      if (f1) {
        const int pos = pos0 + __popc(b1 & lt);
        h_ray[pos] = r;
        h_step[pos] = __ldg(s_step + i);
        h_m1[pos] = i;
        h_w[pos] = w;
        h_sdf[pos] = __ldg(s_sdf + i);
      }
      pos0 += __popc(b1);
    }
  }
}

static int alpha_scan_count(const esr_scene_t *sc, const int32_t *ray_order, int64_t n_rays, const int32_t *off_mask,
                            const float *s_sdf, const float *s_cos, int32_t *cnt_shade, float *alphainv_last,
                            float *s_alpha, float *s_T, esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && n_rays < (1ll << 31));
  if (n_rays == 0) return ESR_OK;
  // s_sdf / s_alpha / s_T may be NULL only when the M1 stream is empty
  ESR_CHECK_ARG(off_mask && cnt_shade && alphainv_last);
  ESR_STAGE("k_neus_alpha", (cudaStream_t)stream);
  if (s_cos)
    k_neus_alpha<true><<<ray_blocks(n_rays), 256, 0, (cudaStream_t)stream>>>(*sc, n_rays, off_mask, s_sdf, s_cos, s_alpha, s_T);
  else
    k_neus_alpha<false><<<ray_blocks(n_rays), 256, 0, (cudaStream_t)stream>>>(*sc, n_rays, off_mask, s_sdf, nullptr, s_alpha, s_T);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_transmittance", (cudaStream_t)stream);
  k_transmittance<<<cdiv(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(*sc, ray_order, n_rays, off_mask, s_alpha, s_T,
                                                                      cnt_shade, alphainv_last);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int esr_alpha_scan_count(const esr_scene_t *sc, const int32_t *ray_order, int64_t n_rays,
                                    const int32_t *off_mask, const float *s_sdf, int32_t *cnt_shade,
                                    float *alphainv_last, float *s_alpha, float *s_T, esr_stream_t stream) {
  return alpha_scan_count(sc, ray_order, n_rays, off_mask, s_sdf, nullptr, cnt_shade, alphainv_last, s_alpha, s_T, stream);
}

extern "C" int esr_alpha_scan_count_g(const esr_scene_t *sc, const int32_t *ray_order, int64_t n_rays,
                                      const int32_t *off_mask, const float *s_sdf, const float *s_cos, int32_t *cnt_shade,
                                      float *alphainv_last, float *s_alpha, float *s_T, esr_stream_t stream) {
  ESR_CHECK_ARG(s_cos || !s_sdf);   // NULL only with an empty M1 stream
  return alpha_scan_count(sc, ray_order, n_rays, off_mask, s_sdf, s_cos, cnt_shade, alphainv_last, s_alpha, s_T, stream);
}

extern "C" int esr_alpha_scan_fill(const esr_scene_t *sc, const int32_t *ray_order, int64_t n_rays,
                                   const int32_t *off_mask, const int32_t *s_step, const float *s_sdf,
                                   const int32_t *off_shade, const float *s_alpha, const float *s_T, int32_t *h_ray,
                                   int32_t *h_step, int32_t *h_m1, float *h_w, float *h_sdf, esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && n_rays < (1ll << 31));
  if (n_rays == 0) return ESR_OK;
  // stream pointers may be NULL when the corresponding stream is empty (M1 == 0 / M3 == 0)
  ESR_CHECK_ARG(off_mask && off_shade);
  ESR_STAGE("k_shade_compact", (cudaStream_t)stream);
  k_shade_compact<<<ray_blocks(n_rays), 256, 0, (cudaStream_t)stream>>>(*sc, ray_order, n_rays, off_mask, s_step, s_sdf,
                                                                       off_shade, s_alpha, s_T, h_ray, h_step, h_m1,
                                                                       h_w, h_sdf);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
// Stage C': backward.  (1) per ray, chunks in reverse: Alphas2Weights backward as a warp suffix scan,
// then d(alpha)/d(prev_est, next_est); (2) per M1 sample: fold neighbour terms and scatter into the
// dense SDF gradient volume with the forward trilinear weights.
// ---------------------------------------------------------------------------------------------
// GRAD ('grad' alpha): dprev receives dL/dsdf of the sample itself (= dL/dprev_est + dL/dnext_est) and dnext
// dL/diter_cos (= dL/dnext_est - dL/dprev_est); the <false> instantiation is the 'interp' kernel unchanged.
template <bool GRAD>
__global__ void __launch_bounds__(256)
    k_alpha_scan_bwd(const __grid_constant__ esr_scene_t sc, const int32_t *__restrict__ ray_order, int64_t n_rays,
                     const int32_t *__restrict__ off_mask, const float *__restrict__ s_sdf,
                     const float *__restrict__ s_cos, const float *__restrict__ s_alpha, const float *__restrict__ s_T,
                     const float *__restrict__ alphainv_last, const float *__restrict__ g_w_m1,
                     const float *__restrict__ g_last, const float *__restrict__ g_alpha,
                     float *__restrict__ dprev, float *__restrict__ dnext) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (int64_t slot = warp; slot < n_rays; slot += nwarps) {
    const int r = ray_order ? ray_order[slot] : (int)slot;
    const int s = off_mask[slot], e = off_mask[slot + 1];
    float carry = (g_last && !g_alpha) ? g_last[r] * alphainv_last[r] : 0.f;
    for (int hi = e; hi > s; hi -= 32) {
      const int i = hi - 1 - (int)lane;
      const bool valid = i >= s;
      float a = 0.f, Ti = -1.f, gw = 0.f;
      if (valid && !g_alpha) {
        a = s_alpha[i];
        Ti = s_T[i];
        gw = g_w_m1[i];
      }
      const bool proc = valid && (g_alpha || Ti >= 0.f);
      const float x = proc ? gw * (Ti * a) : 0.f;
      float inc = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(FULL, inc, o);
        if (lane >= (unsigned)o) inc += u;
      }
      const float back_cum = carry + (inc - x);
      carry += __shfl_sync(FULL, inc, 31);
      if (!valid) continue;
      float dp = 0.f, dn = 0.f;
      if (proc) {
        const float ga = g_alpha ? g_alpha[i]
                                 : (float)((double)(gw * Ti) - (double)back_cum / ((double)(1.f - a) + 1e-10));
        if (ga != 0.f) {
          const float sd = s_sdf[i];
          float pc, nc;
          if constexpr (GRAD) {
            neus_alpha_grad(sd, s_cos[i], sc.s_val, pc, nc);
          } else {
            const bool has_prev = i > s, has_next = i + 1 < e;
            const float sp = has_prev ? s_sdf[i - 1] : 0.f;
            const float sn = has_next ? s_sdf[i + 1] : 0.f;
            neus_alpha(sd, sp, sn, has_prev, has_next, sc.s_val, pc, nc);
          }
          const float q = pc - nc;
          const float num = fmaxf(q, 0.f) + 1e-5f, den = pc + 1e-5f;
          const float rr = num / den;
          if (rr >= 0.f && rr <= 1.f) {  // clamp passes gradient on the closed interval
            const float inv = 1.f / den;
            const float relu_g = q > 0.f ? 1.f : 0.f;
            const float d_pc = ga * (relu_g * inv - num * inv * inv);
            const float d_nc = ga * (-relu_g * inv);
            dp = d_pc * pc * (1.f - pc) * sc.s_val;
            dn = d_nc * nc * (1.f - nc) * sc.s_val;
          }
        }
      }
      dprev[i] = GRAD ? dp + dn : dp;
      dnext[i] = GRAD ? dn - dp : dn;
    }
  }
}

template <bool GRAD>   // GRAD: dprev is dL/dsdf of the sample itself (no neighbour terms), dnext is not read
__global__ void __launch_bounds__(256)
    k_sdf_scatter(const __grid_constant__ esr_scene_t sc, const float *__restrict__ rays_o,
                  const float *__restrict__ rays_d, const int32_t *__restrict__ s_ray,
                  const int32_t *__restrict__ s_step, const float *__restrict__ dprev,
                  const float *__restrict__ dnext, int64_t m1, float *__restrict__ grad_sdf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m1) return;
  const int r = s_ray[i];
  float g;
  if constexpr (GRAD) {
    g = dprev[i];
  } else {
    const bool has_prev = i > 0 && s_ray[i - 1] == r;
    const bool has_next = i + 1 < m1 && s_ray[i + 1] == r;
    g = (has_prev ? 0.5f : 1.f) * dprev[i] + (has_next ? 0.5f : 1.f) * dnext[i];
    if (has_prev) g += 0.5f * dnext[i - 1];
    if (has_next) g += 0.5f * dprev[i + 1];
  }
  if (g == 0.f) return;
  const RaySetup s = ray_setup(rays_o, rays_d, r, sc.xyz_min, sc.xyz_max, sc.near, sc.far, sc.stepdist);
  float px, py, pz;
  ray_point(s, sc.stepdist, s_step[i], px, py, pz);
  const Cell c = make_cell(world_to_index(px, sc.xyz_min[0], sc.xyz_max[0], sc.gx),
                           world_to_index(py, sc.xyz_min[1], sc.xyz_max[1], sc.gy),
                           world_to_index(pz, sc.xyz_min[2], sc.xyz_max[2], sc.gz));
  scatter1(grad_sdf, sc.gx, sc.gy, sc.gz, c, g);
}

extern "C" int esr_alpha_scan_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                  const int32_t *ray_order, int64_t n_rays, const int32_t *off_mask,
                                  const int32_t *s_ray, const int32_t *s_step, const float *s_sdf,
                                  const float *s_alpha, const float *s_T, const float *alphainv_last,
                                  const float *g_w_m1, const float *g_last, float *tmp_dprev, float *tmp_dnext,
                                  int64_t m1, float *grad_sdf_grid, esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && m1 >= 0 && m1 < (1ll << 31));
  if (n_rays == 0 || m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && off_mask && s_ray && s_step && s_sdf && s_alpha && s_T && alphainv_last &&
                g_w_m1 && tmp_dprev && tmp_dnext && grad_sdf_grid);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_alpha_scan_bwd", st);
  k_alpha_scan_bwd<false><<<ray_blocks(n_rays), 256, 0, st>>>(*sc, ray_order, n_rays, off_mask, s_sdf, nullptr, s_alpha, s_T,
                                                              alphainv_last, g_w_m1, g_last, nullptr, tmp_dprev, tmp_dnext);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_sdf_scatter", st);
  k_sdf_scatter<false><<<cdiv(m1, 256), 256, 0, st>>>(*sc, rays_o, rays_d, s_ray, s_step, tmp_dprev, tmp_dnext, m1,
                                                      grad_sdf_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// 'grad' alpha: the same backward through alpha = f(sdf - iter_cos, sdf + iter_cos).  dL/dsdf of every sample is
// scattered into the SDF gradient volume here; dL/diter_cos is left in tmp_dcos[M1] for esr_neus_cos_bwd.
extern "C" int esr_alpha_scan_bwd_g(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                    const int32_t *ray_order, int64_t n_rays, const int32_t *off_mask,
                                    const int32_t *s_ray, const int32_t *s_step, const float *s_sdf, const float *s_cos,
                                    const float *s_alpha, const float *s_T, const float *alphainv_last,
                                    const float *g_w_m1, const float *g_last, float *tmp_dsdf, float *tmp_dcos,
                                    int64_t m1, float *grad_sdf_grid, esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && m1 >= 0 && m1 < (1ll << 31));
  if (n_rays == 0 || m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && off_mask && s_ray && s_step && s_sdf && s_cos && s_alpha && s_T && alphainv_last &&
                g_w_m1 && tmp_dsdf && tmp_dcos && grad_sdf_grid);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_alpha_scan_bwd", st);
  k_alpha_scan_bwd<true><<<ray_blocks(n_rays), 256, 0, st>>>(*sc, ray_order, n_rays, off_mask, s_sdf, s_cos, s_alpha, s_T,
                                                             alphainv_last, g_w_m1, g_last, nullptr, tmp_dsdf, tmp_dcos);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_sdf_scatter", st);
  k_sdf_scatter<true><<<cdiv(m1, 256), 256, 0, st>>>(*sc, rays_o, rays_d, s_ray, s_step, tmp_dsdf, tmp_dcos, m1,
                                                     grad_sdf_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}


extern "C" int esr_neus_alpha_bwd(const esr_scene_t *sc, const float *rays_o, const float *rays_d,
                                  const int32_t *ray_order, int64_t n_rays, const int32_t *off_mask,
                                  const int32_t *s_ray, const int32_t *s_step, const float *s_sdf,
                                  const float *g_alpha_m1, float *tmp_dprev, float *tmp_dnext, int64_t m1,
                                  float *grad_sdf_grid, esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && m1 >= 0 && m1 < (1ll << 31));
  if (n_rays == 0 || m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && off_mask && s_ray && s_step && s_sdf && g_alpha_m1 && tmp_dprev && tmp_dnext &&
                grad_sdf_grid);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_neus_alpha_bwd", st);
  k_alpha_scan_bwd<false><<<ray_blocks(n_rays), 256, 0, st>>>(*sc, ray_order, n_rays, off_mask, s_sdf, nullptr, nullptr,
                                                              nullptr, nullptr, nullptr, nullptr, g_alpha_m1, tmp_dprev,
                                                              tmp_dnext);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_sdf_scatter", st);
  k_sdf_scatter<false><<<cdiv(m1, 256), 256, 0, st>>>(*sc, rays_o, rays_d, s_ray, s_step, tmp_dprev, tmp_dnext, m1,
                                                      grad_sdf_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// 'grad' alpha with dL/dalpha given directly on the M1 stream (the coarse stage, see esr_neus_alpha_bwd): dL/dsdf of
// every sample is scattered into grad_sdf_grid, dL/diter_cos is left in tmp_dcos[M1] for esr_neus_cos_vol_bwd
extern "C" int esr_neus_alpha_bwd_g(const esr_scene_t *sc, const float *rays_o, const float *rays_d, const int32_t *ray_order,
                                    int64_t n_rays, const int32_t *off_mask, const int32_t *s_ray, const int32_t *s_step,
                                    const float *s_sdf, const float *s_cos, const float *g_alpha_m1, float *tmp_dsdf,
                                    float *tmp_dcos, int64_t m1, float *grad_sdf_grid, esr_stream_t stream) {
  if (int e = check_scene(sc)) return e;
  ESR_CHECK_ARG(n_rays >= 0 && m1 >= 0 && m1 < (1ll << 31));
  if (n_rays == 0 || m1 == 0) return ESR_OK;
  ESR_CHECK_ARG(rays_o && rays_d && off_mask && s_ray && s_step && s_sdf && s_cos && g_alpha_m1 && tmp_dsdf && tmp_dcos &&
                grad_sdf_grid);
  cudaStream_t st = (cudaStream_t)stream;
  ESR_STAGE("k_neus_alpha_bwd", st);
  k_alpha_scan_bwd<true><<<ray_blocks(n_rays), 256, 0, st>>>(*sc, ray_order, n_rays, off_mask, s_sdf, s_cos, nullptr, nullptr,
                                                             nullptr, nullptr, nullptr, g_alpha_m1, tmp_dsdf, tmp_dcos);
  ESR_LAUNCH_OK();
  ESR_STAGE("k_sdf_scatter", st);
  k_sdf_scatter<true><<<cdiv(m1, 256), 256, 0, st>>>(*sc, rays_o, rays_d, s_ray, s_step, tmp_dsdf, tmp_dcos, m1,
                                                     grad_sdf_grid);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

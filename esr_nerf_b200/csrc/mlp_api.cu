// mlp_api.cu — C-ABI entry points of the MLP stage (esr_mlp_*) and the skinny output-layer weight gradient.
//
//   forward / data gradient       : fused tcgen05 layer chains (mlp_tc.cu)
//   hidden-layer weight gradients : tcgen05 split-K GEMMs fed by bulk copies of the tiled activations (mlp_tc.cu)
//   output-layer weight gradient  : 3 x 192, SIMT streaming reduction (below)
// Hidden activations H_l and their cotangents dZ_l travel between the kernels in the TILED layout of
// mlp_layout.cuh (act_chunk_index): per 128-row tile, [24 feature chunks][128 rows][8 bf16], so that one
// epilogue thread per row writes / reads 16-byte chunks that are contiguous across the warp.
#include <stdlib.h>

#include "common.cuh"
#include "mlp_layout.cuh"

using namespace esr;

namespace {

ESR_D float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
ESR_D float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// output layer (n_out <= 3 real rows of 8): dWo[o][i] += sum_m dz_out[m][o] * H[m][i]; dbo[o] += sum_m dz_out[m][o].
// Skinny (3 x W) and bound by streaming H once.  Work item = (feature chunk c, slab of rows): a warp walks its
// slab 32 rows at a time, lane = row, one 16-byte load of the tiled H layout per lane (512 contiguous bytes per
// warp), 24 accumulators per lane, one shuffle reduction + RED at the end.
template <int W>
__global__ void __launch_bounds__(256)
    k_mlp_wgrad_out(const float *__restrict__ dz_out /* [m][8] */, const __nv_bfloat16 *__restrict__ h /* tiled */,
                    int64_t row_begin, int64_t row_end, int n_out, float *__restrict__ gW /* [8][W] */,
                    float *__restrict__ gb /* [8] */) {
  static_assert(W == ACT_W, "tiled activation layout");
  constexpr int NC = W / 8;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nslabs = (((int64_t)gridDim.x * blockDim.x) >> 5) / NC;
  const unsigned lane = lane_id();
  const int c = (int)(warp % NC);
  const int64_t slab = warp / NC;
  if (slab >= nslabs) return;
  const int64_t g_begin = row_begin >> 5, g_end = (row_end + 31) >> 5;  // 32-row groups (absolute rows)
  const int64_t per = (g_end - g_begin + nslabs - 1) / nslabs;
  const int64_t g0 = g_begin + slab * per, g1 = min(g_end, g0 + per);
  float acc[3][8], bs[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
  const uint4 *h4 = reinterpret_cast<const uint4 *>(h);
  constexpr int UN = 4;
  for (int64_t g = g0; g < g1; g += UN) {
    uint4 v[UN];
    float d[UN][3];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t m = (g + u) * 32 + lane;
      const bool ok = g + u < g1 && m >= row_begin && m < row_end;
      v[u] = ok ? __ldg(h4 + act_chunk_index(m, c)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int o = 0; o < 3; ++o) d[u][o] = ok ? __ldg(dz_out + m * 8 + o) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        bs[o] += d[u][o];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[o][2 * j] = fmaf(d[u][o], bf16_lo(w4[j]), acc[o][2 * j]);
          acc[o][2 * j + 1] = fmaf(d[u][o], bf16_hi(w4[j]), acc[o][2 * j + 1]);
        }
      }
    }
  }
  if (g0 >= g1) return;
#pragma unroll
  for (int o = 0; o < 3; ++o) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = warp_sum(acc[o][j]);
    bs[o] = warp_sum(bs[o]);
  }
  if (lane == 0) {
    for (int o = 0; o < n_out && o < 3; ++o) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) red_add2(gW + (int64_t)o * W + 8 * c + j, acc[o][j], acc[o][j + 1]);
      if (c == 0) red_add(gb + o, bs[o]);
    }
  }
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  ESR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return ESR_OK;
}

template <int K0, int W, int NH, int DXP>
static int run_bwd(const MlpLayout &L, const esr_mlp_desc_t *d, const void *image, const void *x, const float *y,
                   const float *d_y, int64_t rb, int64_t re, int64_t mt, const void *hidden, void *d_z,
                   float *d_z_out, float *d_x, int dx_cols, int accumulate, float *grad_flat, cudaStream_t st) {
  if (int e = tc_dgrad(d, image, y, d_y, rb, re, mt, hidden, d_z, d_z_out, d_x, dx_cols, accumulate, st)) return e;
  if (!grad_flat) return ESR_OK;
  const __nv_bfloat16 *H = (const __nv_bfloat16 *)hidden;
  const __nv_bfloat16 *Z = (const __nv_bfloat16 *)d_z;
  const int64_t rows = re - rb;
  const int64_t ls = act_rows_padded(mt) * W;  // layer stride of the tiled activation buffers
  if (int e = tc_wgrad(d, x, rb, re, mt, hidden, d_z, grad_flat, st)) return e;
  // 8 warps per block; warps = 24 chunks x slabs.  3 blocks = 24 warps = one slab.
  const int64_t slabs = max((int64_t)1, min((int64_t)num_sms(), (rows + 511) / 512));
  ESR_STAGE("k_mlp_wgrad_out", st);
  k_mlp_wgrad_out<W><<<(unsigned)(3 * slabs), 256, 0, st>>>(d_z_out, H + (int64_t)(NH - 1) * ls, rb, re, d->n_out,
                                           grad_flat + L.flat_w(NH), grad_flat + L.flat_b(NH));
  ESR_LAUNCH_OK();
  return ESR_OK;
}

static int check_desc(const esr_mlp_desc_t *d) {
  ESR_CHECK_ARG(d != nullptr);
  if (!tc_supported(d)) {
    set_error("MLP shape k0=%d width=%d hidden=%d is not instantiated (radiance 96->192x3, tone mapper 48->192x1)",
              d->k0, d->width, d->n_hidden);
    return ESR_ERR_BAD_ARG;
  }
  ESR_CHECK_ARG(d->k0 % 16 == 0 && d->k0 > 0 && d->width % 64 == 0 && d->n_hidden >= 1);
  ESR_CHECK_ARG(d->n_out >= 1 && d->n_out <= 3 && (d->act == 1 || d->act == 2));
  return ESR_OK;
}

}  // namespace

extern "C" int64_t esr_mlp_image_bytes(const esr_mlp_desc_t *d) { return d ? tc_image_bytes(d) : 0; }
extern "C" int64_t esr_mlp_act_rows(int64_t m_total) { return act_rows_padded(m_total); }
extern "C" int64_t esr_mlp_hidden_bytes(const esr_mlp_desc_t *d, int64_t m_total) {
  return d ? act_hidden_bytes(d->n_hidden, m_total) : 0;
}
extern "C" int64_t esr_mlp_param_count(const esr_mlp_desc_t *d) { return d ? layout_of(d).flat_count() : 0; }
extern "C" int esr_mlp_pack(const esr_mlp_desc_t *d, const float *flat_params, void *image, esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(flat_params && image);
  return tc_pack(d, flat_params, image, (cudaStream_t)stream);
}

// instantiated shapes: radiance nets 96->192x3->3 (pbr/module.py:6-21 with dim0 85), tone mapper 48->192->3
// (pbr/module.py:24-39 with dim0 33)

extern "C" int esr_mlp_fwd(const esr_mlp_desc_t *d, const void *image, const void *x, int64_t row_begin,
                           int64_t row_end, int64_t m_total, float *y, void *hidden, esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= m_total);
  if (row_end == row_begin) return ESR_OK;
  ESR_CHECK_ARG(image && x && y);
  return tc_fwd(d, image, x, row_begin, row_end, m_total, y, hidden, (cudaStream_t)stream);
}

extern "C" int esr_mlp_bwd(const esr_mlp_desc_t *d, const void *image, const void *x, const float *y,
                           const float *d_y, int64_t row_begin, int64_t row_end, int64_t m_total, const void *hidden,
                           void *d_z, float *d_z_out, float *d_x, int dx_cols, int accumulate, float *grad_flat,
                           esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= m_total);
  if (row_end == row_begin) return ESR_OK;
  ESR_CHECK_ARG(image && x && y && d_y && hidden && d_z && d_z_out);
  ESR_CHECK_ARG(!d_x || (dx_cols > 0 && dx_cols % 4 == 0));
  const MlpLayout L = layout_of(d);
  cudaStream_t st = (cudaStream_t)stream;
#define BWD_CALL(...)                                                                                          \
  run_bwd<__VA_ARGS__>(L, d, image, x, y, d_y, row_begin, row_end, m_total, hidden, d_z, d_z_out, d_x, dx_cols, \
                       accumulate, grad_flat, st)
  if (d->k0 == 96 && d->width == 192 && d->n_hidden == 3) {
    ESR_CHECK_ARG(!d_x || dx_cols <= 56);
    return BWD_CALL(96, 192, 3, 56);
  }
  if (d->k0 == 48 && d->width == 192 && d->n_hidden == 1) {
    ESR_CHECK_ARG(!d_x || dx_cols <= 40);
    return BWD_CALL(48, 192, 1, 40);
  }
#undef BWD_CALL
  set_error("esr_mlp_bwd: MLP shape k0=%d width=%d hidden=%d is not instantiated", d->k0, d->width, d->n_hidden);
  return ESR_ERR_BAD_ARG;
}

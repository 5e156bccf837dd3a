// mlp_api.cu — C-ABI entry points of the MLP stage (esr_mlp_*).
//
//   forward / data gradient       : fused tcgen05 layer chains (mlp_tc.cu)
//   hidden-layer weight gradients : tcgen05 split-K GEMMs fed by bulk copies of the tiled activations (mlp_tc.cu)
//   output-layer weight gradient  : the same GEMM with the 16-column dZ_out as A operand
// Hidden activations H_l and their cotangents dZ_l travel between the kernels in the TILED layout of
// mlp_layout.cuh (act_chunk_index): per 128-row tile, [24 feature chunks][128 rows][8 bf16], so that one
// epilogue thread per row writes / reads 16-byte chunks that are contiguous across the warp.
#include <stdlib.h>

#include "common.cuh"
#include "mlp_layout.cuh"

using namespace esr;

namespace {

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
  if (int e = tc_wgrad(d, x, rb, re, mt, hidden, d_z, grad_flat, st)) return e;
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
  ESR_CHECK_ARG(d->n_out >= 1 && d->n_out <= 8 && (d->act == 1 || d->act == 2));
  return ESR_OK;
}

}  // namespace

extern "C" int64_t esr_mlp_image_bytes(const esr_mlp_desc_t *d) { return d ? tc_image_bytes(d) : 0; }
extern "C" int64_t esr_mlp_act_rows(int64_t m_total) { return act_rows_padded(m_total); }
extern "C" int64_t esr_mlp_hidden_bytes(const esr_mlp_desc_t *d, int64_t m_total) {
  return d ? act_hidden_bytes(d->n_hidden, m_total) : 0;
}
extern "C" int64_t esr_mlp_dz_bytes(const esr_mlp_desc_t *d, int64_t m_total) {
  return d ? act_dz_bytes(d->n_hidden, m_total) : 0;
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
                           int64_t row_end, int64_t m_total, float *y, void *hidden, int64_t save_row_begin,
                           esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= m_total);
  if (row_end == row_begin) return ESR_OK;
  ESR_CHECK_ARG(image && x && y);
  return tc_fwd(d, image, x, row_begin, row_end, m_total, y, hidden, save_row_begin, (cudaStream_t)stream);
}

extern "C" int esr_mlp_bwd(const esr_mlp_desc_t *d, const void *image, const void *x, const float *y,
                           const float *d_y, int64_t row_begin, int64_t row_end, int64_t m_total, const void *hidden,
                           void *d_z, float *d_z_out, float *d_x, int dx_cols, int accumulate, float *grad_flat,
                           esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= m_total);
  if (row_end == row_begin) return ESR_OK;
  ESR_CHECK_ARG(image && x && y && d_y && hidden && d_z);
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
    ESR_CHECK_ARG(!d_x || dx_cols <= 48);
    return BWD_CALL(48, 192, 1, 48);
  }
#undef BWD_CALL
  set_error("esr_mlp_bwd: MLP shape k0=%d width=%d hidden=%d is not instantiated", d->k0, d->width, d->n_hidden);
  return ESR_ERR_BAD_ARG;
}

extern "C" int esr_mlp_bwd_weights(const esr_mlp_desc_t *d, const void *x, int64_t row_begin, int64_t row_end,
                                   int64_t m_total, const void *hidden, const void *d_z, float *grad_flat,
                                   esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= m_total);
  if (row_end == row_begin) return ESR_OK;
  ESR_CHECK_ARG(x && hidden && d_z && grad_flat);
  return tc_wgrad(d, x, row_begin, row_end, m_total, hidden, d_z, grad_flat, (cudaStream_t)stream);
}

// fused tone-map net (voxurff.py:783-788 + pbr/module.py:24-39): see mlp_tc.cu
static int check_tonemap_desc(const esr_mlp_desc_t *d) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(d->k0 == 48 && d->n_hidden == 1 && d->n_out <= 3);
  return ESR_OK;
}

extern "C" int esr_tonemap_mlp_fwd(const esr_mlp_desc_t *d, const void *image, const float *lin, int64_t m, float *rgb,
                                   esr_stream_t stream) {
  if (int e = check_tonemap_desc(d)) return e;
  ESR_CHECK_ARG(m >= 0);
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(image && lin && rgb);
  return tc_tonemap_fwd(d, image, lin, m, rgb, (cudaStream_t)stream);
}

extern "C" int esr_tonemap_mlp_bwd(const esr_mlp_desc_t *d, const void *image, const float *lin, const float *rgb,
                                   const float *d_rgb, const float *d_lin_direct, int64_t m, float *d_lin,
                                   float *grad_flat, esr_stream_t stream) {
  if (int e = check_tonemap_desc(d)) return e;
  ESR_CHECK_ARG(m >= 0);
  if (m == 0) return ESR_OK;
  ESR_CHECK_ARG(image && lin && rgb && d_rgb && d_lin && grad_flat);
  return tc_tonemap_bwd(d, image, lin, rgb, d_rgb, d_lin_direct, m, d_lin, grad_flat, (cudaStream_t)stream);
}

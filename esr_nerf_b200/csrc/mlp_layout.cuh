// mlp_layout.cuh — parameter and activation layouts shared by the MLP kernels (mlp_tc.cu: tcgen05 forward /
// data-gradient chains and weight-gradient GEMMs; mlp_api.cu: C-ABI entry points).
#pragma once
#include "common.cuh"

namespace esr {

// ------------------------------------------------------------------------------------------------
// parameter image layout
// ------------------------------------------------------------------------------------------------
struct MlpLayout {
  int k0, W, NH, n_out;
  // flat f32 master copy
  __host__ __device__ int64_t flat_w(int l) const {  // offset of W_l
    if (l == 0) return 0;
    int64_t o = (int64_t)W * k0 + W;
    o += (int64_t)(l - 1) * ((int64_t)W * W + W);
    return o;
  }
  __host__ __device__ int64_t flat_b(int l) const { return flat_w(l) + (l == 0 ? (int64_t)W * k0 : (l < NH ? (int64_t)W * W : (int64_t)8 * W)); }
  __host__ __device__ int64_t flat_count() const { return flat_b(NH) + 8; }
};

static inline MlpLayout layout_of(const esr_mlp_desc_t *d) { return MlpLayout{d->k0, d->width, d->n_hidden, d->n_out}; }


// ------------------------------------------------------------------------------------------------
// Activation buffers (hidden H_l saved by the forward chain, dZ_l saved by the data-gradient chain): per layer
// [ceil(m/128)] tiles of [24 feature chunks][128 rows][8 bf16] — the shared-memory operand layout of the
// tensor-core kernels, so that one thread per row moves 16-byte chunks that are contiguous across a warp.
// ------------------------------------------------------------------------------------------------
// The bf16 MLP inputs (encoded feature rows x, tone-map feature rows) use the same tiling with their own width.
constexpr int ACT_W = 192;
ESR_HD int64_t act_rows_padded(int64_t m_total) { return (m_total + 127) / 128 * 128; }
// index, in 16-byte units, of feature chunk c (8 features) of absolute row `row` in a [*, 8*chunks_per_row] matrix
ESR_HD int64_t tiled_chunk_index(int64_t row, int c, int chunks_per_row) {
  return ((row >> 7) * chunks_per_row + c) * 128 + (row & 127);
}
ESR_HD int64_t act_chunk_index(int64_t row, int c) { return tiled_chunk_index(row, c, ACT_W / 8); }
// `hidden` buffer = [n_hidden][rows_padded][192] bf16 activations, then the ReLU masks the data-gradient chain reads
// instead of the activations: per layer [tile][4 column groups][128 rows] x uint2, 48 bits used.  Column
// 48 grp + 16 c + 2 j + h (c < 3 sixteen-column chunks, j < 8 pairs, h = low / high half of the packed bf16 pair) is
// bit 8 (c & 1) + j + 16 h of word c >> 1 — the layout that lets the forward epilogue derive both bits of a pair
// from the packed word with three integer operations (mlp_tc.cu).
ESR_HD int64_t act_hidden_bytes(int n_hidden, int64_t m_total) {
  return (int64_t)n_hidden * act_rows_padded(m_total) * (ACT_W * 2 + 32);
}
// `d_z` scratch = [n_hidden][rows_padded][192] bf16 hidden-layer cotangents, then the output-layer cotangent as
// [rows_padded][16] bf16 (tiled with 2 chunks per row), then a 128-byte tail whose first word is max |d_y| (f32 bits)
// of the launch: precision 1 stores the cotangents as fp16 times the power of two that puts that maximum in [64, 128)
// (act_dz_scale), the weight-gradient GEMM divides its sums by it
ESR_HD int64_t act_dz_tail_offset(int n_hidden, int64_t m_total) {
  return act_rows_padded(m_total) * ((int64_t)n_hidden * ACT_W * 2 + 32);
}
ESR_HD int64_t act_dz_bytes(int n_hidden, int64_t m_total) { return act_dz_tail_offset(n_hidden, m_total) + 128; }
// (G, 1 / G) for the launch's max |d_y| given as f32 bits: G = 2^(133 - e) for a biased exponent e in [8, 250]
// (zero / denormal / non-finite maxima: 1) — the stored cotangents then stay below 2^7 x (what the chain adds on top)
ESR_D float act_dz_scale(uint32_t absmax_bits, float &inv) {
  const int e = (int)((absmax_bits >> 23) & 0xff);
  const bool ok = e >= 8 && e <= 250;
  inv = ok ? __int_as_float((e - 6) << 23) : 1.f;
  return ok ? __int_as_float((260 - e) << 23) : 1.f;
}
ESR_HD int64_t act_mask_base_bytes(int n_hidden, int64_t m_total) {
  return (int64_t)n_hidden * act_rows_padded(m_total) * ACT_W * 2;
}
// index, in 8-byte units from the mask base, of the mask words of (layer l, row, column group)
ESR_HD int64_t act_mask_index(int l, int64_t rows_padded, int64_t row, int grp) {
  return (int64_t)l * rows_padded * 4 + ((row >> 7) * 4 + grp) * 128 + (row & 127);
}

// ------------------------------------------------------------------------------------------------
// tcgen05 path (mlp_tc.cu)
// ------------------------------------------------------------------------------------------------
int64_t tc_image_bytes(const esr_mlp_desc_t *d);
int tc_pack(const esr_mlp_desc_t *d, const float *flat_params, void *tc_image, cudaStream_t st);
bool tc_supported(const esr_mlp_desc_t *d);
int tc_fwd(const esr_mlp_desc_t *d, const void *tc_image, const void *x, int64_t row_begin, int64_t row_end,
           int64_t m_total, float *y, void *hidden, int64_t save_begin, cudaStream_t st);
int tc_dgrad(const esr_mlp_desc_t *d, const void *tc_image, const float *y, const float *d_y, int64_t row_begin,
             int64_t row_end, int64_t m_total, const void *hidden, void *d_z, float *d_z_out, float *d_x,
             int dx_cols, int accumulate, cudaStream_t st);
// fused tone-map net (33 -> 192 -> 3, k0 = 48, one hidden layer): forward from the f32 linear radiance [m,3] with the
// positional encoding computed in the kernel; backward = data gradient + all weight gradients in one kernel
int tc_tonemap_fwd(const esr_mlp_desc_t *d, const void *tc_image, const float *lin, int64_t m, float *y, cudaStream_t st);
int tc_tonemap_bwd(const esr_mlp_desc_t *d, const void *tc_image, const float *lin, const float *y, const float *d_y,
                   const float *d_direct, int64_t m, float *d_lin, float *grad_flat, cudaStream_t st);
// hidden-layer weight / bias gradients: grad_flat (flat master layout) += dZ_l^T . In_l for l = 0 .. n_hidden-1
int tc_wgrad(const esr_mlp_desc_t *d, const void *x, int64_t row_begin, int64_t row_end, int64_t m_total,
             const void *hidden, const void *d_z, float *grad_flat, cudaStream_t st);

}  // namespace esr

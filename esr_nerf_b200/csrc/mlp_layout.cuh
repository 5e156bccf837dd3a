// mlp_layout.cuh — parameter layouts shared by the MLP kernels (mlp_mma.cu: mma.sync path and weight-gradient
// GEMMs; mlp_tc.cu: tcgen05 forward / data-gradient chains).
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
  // bf16 image, element offsets (bf16 units) of the forward weights
  __host__ __device__ int64_t img_w(int l) const {
    if (l == 0) return 0;
    return (int64_t)W * k0 + (int64_t)(l - 1) * W * W;
  }
  __host__ __device__ int64_t img_fwd_elems() const { return img_w(NH) + (int64_t)8 * W; }
  __host__ __device__ int64_t img_bias_bytes_off() const { return img_fwd_elems() * 2; }
  __host__ __device__ int64_t n_bias() const { return (int64_t)NH * W + 8; }
  __host__ __device__ int64_t img_bwd_bytes_off() const { return img_bias_bytes_off() + n_bias() * 4; }
  // transposed copies (bf16 units relative to img_bwd): woT [W][16], whT[l-1] [W][W] (l=1..NH-1), w0T [k0][W]
  __host__ __device__ int64_t imgT_wo() const { return 0; }
  __host__ __device__ int64_t imgT_wh(int l) const { return (int64_t)W * 16 + (int64_t)(l - 1) * W * W; }
  __host__ __device__ int64_t imgT_w0() const { return (int64_t)W * 16 + (int64_t)(NH - 1) * W * W; }
  __host__ __device__ int64_t imgT_elems() const { return imgT_w0() + (int64_t)k0 * W; }
  __host__ __device__ int64_t img_bytes() const { return img_bwd_bytes_off() + imgT_elems() * 2; }
};

static inline MlpLayout layout_of(const esr_mlp_desc_t *d) { return MlpLayout{d->k0, d->width, d->n_hidden, d->n_out}; }


// ------------------------------------------------------------------------------------------------
// tcgen05 path (mlp_tc.cu).  The "tc image" follows the mma.sync image in the same buffer.
// ------------------------------------------------------------------------------------------------
int64_t tc_image_bytes(const esr_mlp_desc_t *d);
int tc_pack(const esr_mlp_desc_t *d, const float *flat_params, void *tc_image, cudaStream_t st);
bool tc_supported(const esr_mlp_desc_t *d);
int tc_fwd(const esr_mlp_desc_t *d, const void *tc_image, const void *x, int64_t row_begin, int64_t row_end,
           int64_t m_total, float *y, void *hidden, cudaStream_t st);
int tc_dgrad(const esr_mlp_desc_t *d, const void *tc_image, const float *y, const float *d_y, int64_t row_begin,
             int64_t row_end, int64_t m_total, const void *hidden, void *d_z, float *d_z_out, float *d_x,
             int dx_cols, int accumulate, cudaStream_t st);

}  // namespace esr

// mlp_mma.cu — the small radiance / tone-map MLPs (pbr/module.py:6-39) on tensor cores.
//
// v1 tensor path: warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  The whole parameter set of
// one net (<= 200 KB as bf16) is staged ONCE per CTA in shared memory and the CTAs are persistent over
// 128-row tiles; activations never leave registers between layers (the m16n8 accumulator fragment of
// layer l is re-packed in place into the m16k16 A fragment of layer l+1).
//   forward : x[rows,k0] bf16 -> y[rows,n_out] f32 (+ optional bf16 hidden activations for training)
//   dgrad   : d_y -> d_z of every layer (bf16, for wgrad) -> d_x (f32, first dx_cols columns)
//   wgrad   : dW_l += dZ_l^T . In_l as a split-K (over samples) GEMM with ldmatrix.trans operands,
//             fp32 partials reduced into the flat gradient with RED.
#include <stdlib.h>

#include "common.cuh"
#include "mlp_layout.cuh"

using namespace esr;

namespace {

__global__ void k_mlp_pack(MlpLayout L, const float *__restrict__ flat, uint8_t *__restrict__ image) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  __nv_bfloat16 *fw = reinterpret_cast<__nv_bfloat16 *>(image);
  float *bias = reinterpret_cast<float *>(image + L.img_bias_bytes_off());
  __nv_bfloat16 *bw = reinterpret_cast<__nv_bfloat16 *>(image + L.img_bwd_bytes_off());
  const int W = L.W;
  // forward weights + biases
  if (i < L.img_fwd_elems()) {
    int l = 0;
    while (l < L.NH && i >= L.img_w(l + 1)) ++l;
    const int64_t e = i - L.img_w(l);
    fw[i] = __float2bfloat16(flat[L.flat_w(l) + e]);
  }
  if (i < L.n_bias()) {
    const int l = (int)(i / W) < L.NH ? (int)(i / W) : L.NH;
    const int64_t e = i - (int64_t)l * W;
    bias[i] = flat[L.flat_b(l) + e];
  }
  // transposed copies
  if (i < L.imgT_elems()) {
    float v;
    if (i < L.imgT_wh(1)) {  // woT [W][16]: (in i, out o) <- Wo[o][i], zero for o >= 8
      const int in = (int)(i / 16), o = (int)(i % 16);
      v = o < 8 ? flat[L.flat_w(L.NH) + (int64_t)o * W + in] : 0.f;
    } else if (i < L.imgT_w0()) {  // whT[l-1] [in][out] <- W_l[out][in]
      const int64_t e = i - L.imgT_wh(1);
      const int l = 1 + (int)(e / ((int64_t)W * W));
      const int64_t r = e % ((int64_t)W * W);
      const int in = (int)(r / W), o = (int)(r % W);
      v = flat[L.flat_w(l) + (int64_t)o * W + in];
    } else {  // w0T [k0][W] <- W_0[out][in]
      const int64_t e = i - L.imgT_w0();
      const int in = (int)(e / W), o = (int)(e % W);
      v = flat[L.flat_w(0) + (int64_t)o * L.k0 + in];
    }
    bw[i] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
ESR_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

ESR_D void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
ESR_D void ldsm_x2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
ESR_D void ldsm_x4_t(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
ESR_D void ldsm_x2_t(uint32_t addr, uint32_t &r0, uint32_t &r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
// D += A(16x16, row) * B(16x8, col), bf16 inputs, fp32 accumulate
ESR_D void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
ESR_D uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}
ESR_D float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
ESR_D float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
ESR_D void cp_async16(uint32_t dst, const void *src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
ESR_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
ESR_D void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// copy a [rows][cols] bf16 matrix from global into shared with a padded row stride (cols + 8)
ESR_D void stage_matrix(__nv_bfloat16 *dst, const __nv_bfloat16 *__restrict__ src, int rows, int cols) {
  const int chunks = cols / 8;  // 16-byte chunks per row
  const int stride = cols + 8;
  for (int c = threadIdx.x; c < rows * chunks; c += blockDim.x) {
    const int r = c / chunks, q = c - r * chunks;
    *reinterpret_cast<uint4 *>(dst + r * stride + q * 8) = __ldg(reinterpret_cast<const uint4 *>(src + r * cols + q * 8));
  }
}

// One dense layer for a warp's 16 rows: acc[NT][4] (+)= A[KT] . B^T, B staged as [n][k] with row stride
// (k_cols + 8).  ldmatrix.x4 fetches the B fragments of two n-tiles per k-tile.
template <int KT, int NT>
ESR_D void warp_layer(const uint32_t (&a)[KT][4], uint32_t w_smem, int stride_elems, float (&acc)[NT][4]) {
  const unsigned lane = lane_id();
  const unsigned mi = lane >> 3, r = lane & 7;
  const uint32_t lane_off = (uint32_t)(((8 * (mi >> 1) + r) * stride_elems + 8 * (mi & 1)) * 2);
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(w_smem + lane_off + (uint32_t)((16 * np * stride_elems + 16 * kt) * 2), b0, b1, b2, b3);
      mma_bf16(acc[2 * np], a[kt], b0, b1);
      mma_bf16(acc[2 * np + 1], a[kt], b2, b3);
    }
    if (NT & 1) {
      uint32_t b0, b1;
      const uint32_t off2 = (uint32_t)(((r)*stride_elems + 8 * (mi & 1)) * 2);  // lanes 0-15 supply addresses
      ldsm_x2(w_smem + off2 + (uint32_t)((8 * (NT - 1) * stride_elems + 16 * kt) * 2), b0, b1);
      mma_bf16(acc[NT - 1], a[kt], b0, b1);
    }
  }
}

template <int NT>
ESR_D void zero_acc(float (&acc)[NT][4]) {
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
}

constexpr int MLP_THREADS = 256;
constexpr int TILE_ROWS = 128;

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int K0, int W, int NH>
struct FwdSmem {
  static constexpr int w0 = 0;                                    // [W][K0+8]
  static constexpr int wh = w0 + W * (K0 + 8);                    // [NH-1][W][W+8]
  static constexpr int wo = wh + (NH - 1) * W * (W + 8);          // [8][W+8]
  static constexpr int bf16_elems = wo + 8 * (W + 8);
  static constexpr int bias_off_bytes = bf16_elems * 2;           // f32 [NH*W+8]
  static constexpr int bytes = bias_off_bytes + (NH * W + 8) * 4;
};

template <int K0, int W, int NH>
__global__ void __launch_bounds__(MLP_THREADS, 1)
    k_mlp_fwd(MlpLayout L, const uint8_t *__restrict__ image, const __nv_bfloat16 *__restrict__ x, int64_t row_begin,
              int64_t row_end, int64_t m_total, float *__restrict__ y, __nv_bfloat16 *__restrict__ hidden, int n_out,
              int act) {
  extern __shared__ __align__(16) uint8_t smem[];
  using S = FwdSmem<K0, W, NH>;
  __nv_bfloat16 *sw = reinterpret_cast<__nv_bfloat16 *>(smem);
  float *sbias = reinterpret_cast<float *>(smem + S::bias_off_bytes);
  {
    const __nv_bfloat16 *gw = reinterpret_cast<const __nv_bfloat16 *>(image);
    stage_matrix(sw + S::w0, gw + L.img_w(0), W, K0);
#pragma unroll
    for (int l = 1; l < NH; ++l) stage_matrix(sw + S::wh + (l - 1) * W * (W + 8), gw + L.img_w(l), W, W);
    stage_matrix(sw + S::wo, gw + L.img_w(NH), 8, W);
    const float *gb = reinterpret_cast<const float *>(image + L.img_bias_bytes_off());
    for (int i = threadIdx.x; i < NH * W + 8; i += blockDim.x) sbias[i] = gb[i];
  }
  __syncthreads();

  constexpr int KT0 = K0 / 16, KT = W / 16, NT = W / 8;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned g = lane >> 2, t = lane & 3;
  const uint32_t s_base = smem_u32(sw);
  const int64_t n_rows = row_end - row_begin;
  const int64_t n_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t rA = row_begin + tile * TILE_ROWS + warp * 16 + g, rB = rA + 8;
    const bool vA = rA < row_end, vB = rB < row_end;
    float acc[NT][4];
    uint32_t a[KT][4];
    {  // layer 0: A fragments straight from global memory
      uint32_t a0[KT0][4];
      const uint32_t *xa = reinterpret_cast<const uint32_t *>(x + rA * K0);
      const uint32_t *xb = reinterpret_cast<const uint32_t *>(x + rB * K0);
#pragma unroll
      for (int kt = 0; kt < KT0; ++kt) {
        a0[kt][0] = vA ? __ldg(xa + 8 * kt + t) : 0u;
        a0[kt][1] = vB ? __ldg(xb + 8 * kt + t) : 0u;
        a0[kt][2] = vA ? __ldg(xa + 8 * kt + 4 + t) : 0u;
        a0[kt][3] = vB ? __ldg(xb + 8 * kt + 4 + t) : 0u;
      }
      zero_acc(acc);
      warp_layer<KT0, NT>(a0, s_base + S::w0 * 2, K0 + 8, acc);
    }
#pragma unroll
    for (int l = 0; l < NH; ++l) {
      // bias + ReLU, re-pack the accumulator fragment as the next layer's A fragment
      const float *b = sbias + l * W;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * kt + h;
          const float b0 = b[8 * j + 2 * t], b1 = b[8 * j + 2 * t + 1];
          a[kt][2 * h] = pack_bf16(fmaxf(acc[j][0] + b0, 0.f), fmaxf(acc[j][1] + b1, 0.f));
          a[kt][2 * h + 1] = pack_bf16(fmaxf(acc[j][2] + b0, 0.f), fmaxf(acc[j][3] + b1, 0.f));
        }
      }
      if (hidden) {
        uint32_t *ha = reinterpret_cast<uint32_t *>(hidden + ((int64_t)l * m_total + rA) * W);
        uint32_t *hb = reinterpret_cast<uint32_t *>(hidden + ((int64_t)l * m_total + rB) * W);
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          if (vA) ha[8 * kt + t] = a[kt][0], ha[8 * kt + 4 + t] = a[kt][2];
          if (vB) hb[8 * kt + t] = a[kt][1], hb[8 * kt + 4 + t] = a[kt][3];
        }
      }
      if (l + 1 < NH) {
        zero_acc(acc);
        warp_layer<KT, NT>(a, s_base + (S::wh + l * W * (W + 8)) * 2, W + 8, acc);
      }
    }
    // output layer: one n-tile of 8 (n_out <= 8 real outputs)
    float o[1][4];
    zero_acc(o);
    warp_layer<KT, 1>(a, s_base + S::wo * 2, W + 8, o);
    const float *bo = sbias + NH * W;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = 2 * t + h;
      if (col < n_out) {
        float zA = o[0][h] + bo[col], zB = o[0][2 + h] + bo[col];
        if (act == 1) {  // softplus(beta=1, threshold=20)
          zA = zA > 20.f ? zA : log1pf(expf(zA));
          zB = zB > 20.f ? zB : log1pf(expf(zB));
        } else if (act == 2) {
          zA = 1.f / (1.f + expf(-zA));
          zB = 1.f / (1.f + expf(-zB));
        }
        if (vA) y[rA * n_out + col] = zA;
        if (vB) y[rB * n_out + col] = zB;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, data gradient chain
// ------------------------------------------------------------------------------------------------
template <int K0, int W, int NH, int DXP /* padded dx cols, multiple of 8 */>
struct BwdSmem {
  static constexpr int woT = 0;                                   // [W][16+8]
  static constexpr int whT = woT + W * 24;                        // [NH-1][W][W+8]
  static constexpr int w0T = whT + (NH - 1) * W * (W + 8);        // [DXP][W+8]
  static constexpr int bf16_elems = w0T + DXP * (W + 8);
  static constexpr int bytes = bf16_elems * 2;
};

template <int K0, int W, int NH, int DXP>
__global__ void __launch_bounds__(MLP_THREADS, 1)
    k_mlp_dgrad(MlpLayout L, const uint8_t *__restrict__ image, const float *__restrict__ y,
                const float *__restrict__ d_y, int64_t row_begin, int64_t row_end, int64_t m_total,
                const __nv_bfloat16 *__restrict__ hidden, __nv_bfloat16 *__restrict__ d_z,
                float *__restrict__ d_z_out, float *__restrict__ d_x, int dx_cols, int accumulate, int n_out,
                int act) {
  extern __shared__ __align__(16) uint8_t smem[];
  using S = BwdSmem<K0, W, NH, DXP>;
  __nv_bfloat16 *sw = reinterpret_cast<__nv_bfloat16 *>(smem);
  {
    const __nv_bfloat16 *gw = reinterpret_cast<const __nv_bfloat16 *>(image + L.img_bwd_bytes_off());
    stage_matrix(sw + S::woT, gw + L.imgT_wo(), W, 16);
#pragma unroll
    for (int l = 1; l < NH; ++l) stage_matrix(sw + S::whT + (l - 1) * W * (W + 8), gw + L.imgT_wh(l), W, W);
    stage_matrix(sw + S::w0T, gw + L.imgT_w0(), DXP, W);
  }
  __syncthreads();

  constexpr int KT = W / 16, NT = W / 8, NTX = DXP / 8;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned g = lane >> 2, t = lane & 3;
  const uint32_t s_base = smem_u32(sw);
  const int64_t n_rows = row_end - row_begin;
  const int64_t n_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t rA = row_begin + tile * TILE_ROWS + warp * 16 + g, rB = rA + 8;
    const bool vA = rA < row_end, vB = rB < row_end;
    // d z_out = d_y * act'(y)   (softplus: 1 - exp(-y); sigmoid: y (1 - y))
    float dzA[2] = {0.f, 0.f}, dzB[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = 2 * t + h;
      if (col < n_out) {
        if (vA) {
          const float yy = y[rA * n_out + col];
          dzA[h] = d_y[rA * n_out + col] * (act == 1 ? (1.f - expf(-yy)) : (act == 2 ? yy * (1.f - yy) : 1.f));
        }
        if (vB) {
          const float yy = y[rB * n_out + col];
          dzB[h] = d_y[rB * n_out + col] * (act == 1 ? (1.f - expf(-yy)) : (act == 2 ? yy * (1.f - yy) : 1.f));
        }
      }
    }
    if (d_z_out) {
      if (vA) *reinterpret_cast<float2 *>(d_z_out + rA * 8 + 2 * t) = make_float2(dzA[0], dzA[1]);
      if (vB) *reinterpret_cast<float2 *>(d_z_out + rB * 8 + 2 * t) = make_float2(dzB[0], dzB[1]);
    }
    float acc[NT][4];
    uint32_t a[KT][4];
    {
      uint32_t ao[1][4];
      ao[0][0] = pack_bf16(dzA[0], dzA[1]);
      ao[0][1] = pack_bf16(dzB[0], dzB[1]);
      ao[0][2] = 0u;
      ao[0][3] = 0u;
      zero_acc(acc);
      warp_layer<1, NT>(ao, s_base + S::woT * 2, 24, acc);
    }
#pragma unroll
    for (int l = NH - 1; l >= 0; --l) {
      // ReLU mask from the saved activations of layer l; pack d z_l; save it for wgrad
      const uint32_t *ha = reinterpret_cast<const uint32_t *>(hidden + ((int64_t)l * m_total + rA) * W);
      const uint32_t *hb = reinterpret_cast<const uint32_t *>(hidden + ((int64_t)l * m_total + rB) * W);
      uint32_t *za = reinterpret_cast<uint32_t *>(d_z + ((int64_t)l * m_total + rA) * W);
      uint32_t *zb = reinterpret_cast<uint32_t *>(d_z + ((int64_t)l * m_total + rB) * W);
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * kt + h;
          const uint32_t mA = vA ? __ldg(ha + 4 * j + t) : 0u;
          const uint32_t mB = vB ? __ldg(hb + 4 * j + t) : 0u;
          const float x0 = bf16_lo(mA) > 0.f ? acc[j][0] : 0.f;
          const float x1 = bf16_hi(mA) > 0.f ? acc[j][1] : 0.f;
          const float x2 = bf16_lo(mB) > 0.f ? acc[j][2] : 0.f;
          const float x3 = bf16_hi(mB) > 0.f ? acc[j][3] : 0.f;
          a[kt][2 * h] = pack_bf16(x0, x1);
          a[kt][2 * h + 1] = pack_bf16(x2, x3);
          if (vA) za[4 * j + t] = a[kt][2 * h];
          if (vB) zb[4 * j + t] = a[kt][2 * h + 1];
        }
      }
      if (l > 0) {
        zero_acc(acc);
        warp_layer<KT, NT>(a, s_base + (S::whT + (l - 1) * W * (W + 8)) * 2, W + 8, acc);
      }
    }
    if (d_x) {
      float ax[NTX][4];
      zero_acc(ax);
      warp_layer<KT, NTX>(a, s_base + S::w0T * 2, W + 8, ax);
#pragma unroll
      for (int j = 0; j < NTX; ++j) {
        const int col = 8 * j + 2 * t;
        if (col < dx_cols) {  // dx_cols is even
          if (vA) {
            float2 *p = reinterpret_cast<float2 *>(d_x + rA * dx_cols + col);
            float2 v = make_float2(ax[j][0], ax[j][1]);
            if (accumulate) {
              const float2 o = *p;
              v.x += o.x, v.y += o.y;
            }
            *p = v;
          }
          if (vB) {
            float2 *p = reinterpret_cast<float2 *>(d_x + rB * dx_cols + col);
            float2 v = make_float2(ax[j][2], ax[j][3]);
            if (accumulate) {
              const float2 o = *p;
              v.x += o.x, v.y += o.y;
            }
            *p = v;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, weight gradient: dW[o][i] += sum_m dZ[m][o] * In[m][i];  db[o] += sum_m dZ[m][o]
// CTA = 8 warps as 4 (o) x 2 (i); split-K over the sample dimension across CTAs.
// ------------------------------------------------------------------------------------------------
constexpr int WG_KSTEP = 32;  // samples per pipeline stage

template <int W, int KIN>
__global__ void __launch_bounds__(MLP_THREADS, 1)
    k_mlp_wgrad(const __nv_bfloat16 *__restrict__ dz, const __nv_bfloat16 *__restrict__ in, int64_t row_begin,
                int64_t row_end, float *__restrict__ gW /* [W][KIN] */, float *__restrict__ gb /* [W] */) {
  constexpr int SZ = W + 8, SI = KIN + 8;
  constexpr int MT = W / 4 / 16;       // m-tiles (o) per warp: 192/4/16 = 3
  constexpr int NTW = KIN / 2 / 8;     // n-tiles (i) per warp
  static_assert(W % 64 == 0 && KIN % 16 == 0, "tile shape");
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16(*s_dz)[WG_KSTEP * SZ] = reinterpret_cast<__nv_bfloat16(*)[WG_KSTEP * SZ]>(smem);
  __nv_bfloat16(*s_in)[WG_KSTEP * SI] =
      reinterpret_cast<__nv_bfloat16(*)[WG_KSTEP * SI]>(smem + 2 * WG_KSTEP * SZ * sizeof(__nv_bfloat16));

  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned wm = warp >> 1, wn = warp & 1;
  const unsigned mi = lane >> 3, r8 = lane & 7;
  // contiguous slab of samples per CTA
  const int64_t n_rows = row_end - row_begin;
  const int64_t steps_total = (n_rows + WG_KSTEP - 1) / WG_KSTEP;
  const int64_t steps_per = (steps_total + gridDim.x - 1) / gridDim.x;
  const int64_t step0 = (int64_t)blockIdx.x * steps_per;
  const int64_t step1 = min(steps_total, step0 + steps_per);
  if (step0 >= step1) return;

  auto load_stage = [&](int buf, int64_t step) {
    const int64_t m0 = row_begin + step * WG_KSTEP;
    constexpr int CZ = W / 8, CI = KIN / 8;
    for (int c = threadIdx.x; c < WG_KSTEP * CZ; c += MLP_THREADS) {
      const int rr = c / CZ, q = c - rr * CZ;
      const bool ok = m0 + rr < row_end;
      cp_async16(smem_u32(&s_dz[buf][rr * SZ + q * 8]), dz + (ok ? (m0 + rr) * W + q * 8 : 0), ok);
    }
    for (int c = threadIdx.x; c < WG_KSTEP * CI; c += MLP_THREADS) {
      const int rr = c / CI, q = c - rr * CI;
      const bool ok = m0 + rr < row_end;
      cp_async16(smem_u32(&s_in[buf][rr * SI + q * 8]), in + (ok ? (m0 + rr) * KIN + q * 8 : 0), ok);
    }
    cp_async_commit();
  };

  float acc[MT][NTW][4];
#pragma unroll
  for (int i = 0; i < MT; ++i) zero_acc(acc[i]);
  float bsum[MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i) bsum[i][0] = bsum[i][1] = 0.f;

  load_stage(0, step0);
  for (int64_t step = step0; step < step1; ++step) {
    const int buf = (int)((step - step0) & 1);
    if (step + 1 < step1) {
      load_stage(buf ^ 1, step + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t zb = smem_u32(&s_dz[buf][0]), ib = smem_u32(&s_in[buf][0]);
#pragma unroll
    for (int ks = 0; ks < WG_KSTEP / 16; ++ks) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // A = dZ^T: matrices (k 0-7,o 0-7) (k 0-7,o 8-15) (k 8-15,o 0-7) (k 8-15,o 8-15), transposed on load
        const int row = 16 * ks + 8 * (mi >> 1) + r8;
        const int col = (wm * MT + mt) * 16 + 8 * (mi & 1);
        ldsm_x4_t(zb + (uint32_t)((row * SZ + col) * 2), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
        if (wn == 0) {
          bsum[mt][0] += bf16_lo(a[mt][0]) + bf16_hi(a[mt][0]) + bf16_lo(a[mt][2]) + bf16_hi(a[mt][2]);
          bsum[mt][1] += bf16_lo(a[mt][1]) + bf16_hi(a[mt][1]) + bf16_lo(a[mt][3]) + bf16_hi(a[mt][3]);
        }
      }
#pragma unroll
      for (int np = 0; np < NTW / 2; ++np) {
        // B = In: matrices (k 0-7,n j) (k 8-15,n j) (k 0-7,n j+1) (k 8-15,n j+1), transposed on load
        const int row = 16 * ks + 8 * (mi & 1) + r8;
        const int col = (wn * NTW + 2 * np + (mi >> 1)) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(ib + (uint32_t)((row * SI + col) * 2), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16(acc[mt][2 * np], a[mt], b0, b1);
          mma_bf16(acc[mt][2 * np + 1], a[mt], b2, b3);
        }
      }
      if (NTW & 1) {
        const int row = 16 * ks + 8 * (mi & 1) + r8;
        const int col = (wn * NTW + NTW - 1) * 8;
        uint32_t b0, b1;
        ldsm_x2_t(ib + (uint32_t)((row * SI + col) * 2), b0, b1);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma_bf16(acc[mt][NTW - 1], a[mt], b0, b1);
      }
    }
    __syncthreads();
  }
  const unsigned g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int o0 = (wm * MT + mt) * 16 + g;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      const int i0 = (wn * NTW + nt) * 8 + 2 * t;
      red_add2(gW + (int64_t)o0 * KIN + i0, acc[mt][nt][0], acc[mt][nt][1]);
      red_add2(gW + (int64_t)(o0 + 8) * KIN + i0, acc[mt][nt][2], acc[mt][nt][3]);
    }
    if (wn == 0) {
      float s0 = bsum[mt][0], s1 = bsum[mt][1];
      s0 += __shfl_xor_sync(FULL, s0, 1);
      s0 += __shfl_xor_sync(FULL, s0, 2);
      s1 += __shfl_xor_sync(FULL, s1, 1);
      s1 += __shfl_xor_sync(FULL, s1, 2);
      if (t == 0) {
        red_add(gb + o0, s0);
        red_add(gb + o0 + 8, s1);
      }
    }
  }
}

// output layer (n_out <= 3 real rows of 8): dWo[o][i] += sum_m dz_out[m][o] * H[m][i]; dbo[o] += sum_m dz_out[m][o].
// Skinny (3 x W) and bound by streaming H once: a warp owns a contiguous slab of rows, each lane 6 of the W=192
// columns (three bf16x2 words), rows unrolled x4 so 12 loads per lane are in flight; 18 accumulators per lane.
template <int W>
__global__ void __launch_bounds__(256)
    k_mlp_wgrad_out(const float *__restrict__ dz_out /* [m][8] */, const __nv_bfloat16 *__restrict__ h,
                    int64_t row_begin, int64_t row_end, int n_out, float *__restrict__ gW /* [8][W] */,
                    float *__restrict__ gb /* [8] */) {
  static_assert(W == 192, "lane mapping assumes 192 = 32 lanes x 3 bf16x2 words");
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  const int64_t n_rows = row_end - row_begin;
  const int64_t per = (n_rows + nwarps - 1) / nwarps;
  const int64_t m0 = row_begin + warp * per, m1 = min(row_end, m0 + per);
  float acc[3][6], bs[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[o][j] = 0.f;
  const uint32_t *hw = reinterpret_cast<const uint32_t *>(h);  // bf16x2 words, W/2 per row
  constexpr int UN = 4;
  for (int64_t m = m0; m < m1; m += UN) {
    uint32_t v[UN][3];
    float d[UN][3];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const bool ok = m + u < m1;
#pragma unroll
      for (int j = 0; j < 3; ++j) v[u][j] = ok ? __ldg(hw + (m + u) * (W / 2) + 32 * j + lane) : 0u;
#pragma unroll
      for (int o = 0; o < 3; ++o) d[u][o] = ok ? __ldg(dz_out + (m + u) * 8 + o) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        bs[o] += d[u][o];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          acc[o][2 * j] = fmaf(d[u][o], bf16_lo(v[u][j]), acc[o][2 * j]);
          acc[o][2 * j + 1] = fmaf(d[u][o], bf16_hi(v[u][j]), acc[o][2 * j + 1]);
        }
      }
  }
  if (m0 < m1) {
    for (int o = 0; o < n_out && o < 3; ++o) {
#pragma unroll
      for (int j = 0; j < 3; ++j) red_add2(gW + (int64_t)o * W + 2 * (32 * j + lane), acc[o][2 * j], acc[o][2 * j + 1]);
      if (lane == 0) red_add(gb + o, bs[o]);
    }
  }
}

// image = [mma.sync image | pad to 128 | tcgen05 image]
static int64_t tc_image_off(const esr_mlp_desc_t *d) { return (layout_of(d).img_bytes() + 127) / 128 * 128; }

// ESR_MLP_PATH=mma selects the legacy mma.sync forward / data-gradient kernels (A/B measurements); default tcgen05
static bool use_tc(const esr_mlp_desc_t *d) {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("ESR_MLP_PATH");
    mode = (e && e[0] == 'm') ? 0 : 1;
  }
  return mode == 1 && tc_supported(d);
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  ESR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return ESR_OK;
}

static unsigned persistent_grid(int64_t rows) {
  const int64_t tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t sms = num_sms();
  return (unsigned)(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
}

template <int K0, int W, int NH, int DXP>
static int run_fwd(const MlpLayout &L, const esr_mlp_desc_t *d, const void *image, const void *x, int64_t rb,
                   int64_t re, int64_t mt, float *y, void *hidden, cudaStream_t st) {
  auto kern = k_mlp_fwd<K0, W, NH>;
  constexpr int bytes = FwdSmem<K0, W, NH>::bytes;
  if (int e = set_smem(kern, bytes)) return e;
  ESR_STAGE(K0 == 96 ? "k_mlp_fwd_radiance" : "k_mlp_fwd_tonemap", st);
  kern<<<persistent_grid(re - rb), MLP_THREADS, bytes, st>>>(L, (const uint8_t *)image, (const __nv_bfloat16 *)x, rb,
                                                             re, mt, y, (__nv_bfloat16 *)hidden, d->n_out, d->act);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

template <int K0, int W, int NH, int DXP>
static int run_bwd(const MlpLayout &L, const esr_mlp_desc_t *d, const void *image, const void *x, const float *y,
                   const float *d_y, int64_t rb, int64_t re, int64_t mt, const void *hidden, void *d_z,
                   float *d_z_out, float *d_x, int dx_cols, int accumulate, float *grad_flat, cudaStream_t st) {
  if (use_tc(d)) {
    if (int e = tc_dgrad(d, (const uint8_t *)image + tc_image_off(d), y, d_y, rb, re, mt, hidden, d_z, d_z_out, d_x,
                         dx_cols, accumulate, st))
      return e;
  } else {
    auto kern = k_mlp_dgrad<K0, W, NH, DXP>;
    constexpr int bytes = BwdSmem<K0, W, NH, DXP>::bytes;
    if (int e = set_smem(kern, bytes)) return e;
    ESR_STAGE(K0 == 96 ? "k_mlp_dgrad_radiance" : "k_mlp_dgrad_tonemap", st);
    kern<<<persistent_grid(re - rb), MLP_THREADS, bytes, st>>>(L, (const uint8_t *)image, y, d_y, rb, re, mt,
                                                               (const __nv_bfloat16 *)hidden, (__nv_bfloat16 *)d_z,
                                                               d_z_out, d_x, dx_cols, accumulate, d->n_out, d->act);
    ESR_LAUNCH_OK();
  }
  if (!grad_flat) return ESR_OK;
  const __nv_bfloat16 *H = (const __nv_bfloat16 *)hidden;
  const __nv_bfloat16 *Z = (const __nv_bfloat16 *)d_z;
  const int64_t rows = re - rb;
  const unsigned grid = (unsigned)max((int64_t)1, min((int64_t)num_sms(), (rows + 255) / 256));
  // layer 0: In = x
  constexpr int wg_bytes0 = 2 * WG_KSTEP * ((W + 8) + (K0 + 8)) * 2, wg_bytes = 2 * WG_KSTEP * 2 * (W + 8) * 2;
  if (int e = set_smem(k_mlp_wgrad<W, K0>, wg_bytes0)) return e;
  ESR_STAGE("k_mlp_wgrad", st);
  k_mlp_wgrad<W, K0><<<grid, MLP_THREADS, wg_bytes0, st>>>(Z, (const __nv_bfloat16 *)x, rb, re,
                                                           grad_flat + L.flat_w(0), grad_flat + L.flat_b(0));
  ESR_LAUNCH_OK();
  if (NH > 1) {
    if (int e = set_smem(k_mlp_wgrad<W, W>, wg_bytes)) return e;
  }
  for (int l = 1; l < NH; ++l) {
    ESR_STAGE("k_mlp_wgrad", st);
    k_mlp_wgrad<W, W><<<grid, MLP_THREADS, wg_bytes, st>>>(Z + (int64_t)l * mt * W, H + (int64_t)(l - 1) * mt * W, rb,
                                                           re, grad_flat + L.flat_w(l), grad_flat + L.flat_b(l));
    ESR_LAUNCH_OK();
  }
  const unsigned grid_o = (unsigned)max((int64_t)1, min((int64_t)num_sms() * 4, (rows + 255) / 256));
  ESR_STAGE("k_mlp_wgrad_out", st);
  k_mlp_wgrad_out<W><<<grid_o, 256, 0, st>>>(d_z_out, H + (int64_t)(NH - 1) * mt * W, rb, re, d->n_out,
                                           grad_flat + L.flat_w(NH), grad_flat + L.flat_b(NH));
  ESR_LAUNCH_OK();
  return ESR_OK;
}

static int check_desc(const esr_mlp_desc_t *d) {
  ESR_CHECK_ARG(d != nullptr);
  ESR_CHECK_ARG(d->k0 % 16 == 0 && d->k0 > 0 && d->width % 64 == 0 && d->n_hidden >= 1);
  ESR_CHECK_ARG(d->n_out >= 1 && d->n_out <= 3 && (d->act == 1 || d->act == 2));
  return ESR_OK;
}

}  // namespace

extern "C" int64_t esr_mlp_image_bytes(const esr_mlp_desc_t *d) { return d ? tc_image_off(d) + tc_image_bytes(d) : 0; }
extern "C" int64_t esr_mlp_param_count(const esr_mlp_desc_t *d) { return d ? layout_of(d).flat_count() : 0; }
extern "C" int esr_mlp_pack(const esr_mlp_desc_t *d, const float *flat_params, void *image, esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(flat_params && image);
  const MlpLayout L = layout_of(d);
  const int64_t n = max(max(L.img_fwd_elems(), L.imgT_elems()), L.n_bias());
  ESR_STAGE("k_mlp_pack", (cudaStream_t)stream);
  k_mlp_pack<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(L, flat_params, (uint8_t *)image);
  ESR_LAUNCH_OK();
  return tc_pack(d, flat_params, (uint8_t *)image + tc_image_off(d), (cudaStream_t)stream);
}

// instantiated shapes: radiance nets 96->192x3->3 (pbr/module.py:6-21 with dim0 85), tone mapper 48->192->3
// (pbr/module.py:24-39 with dim0 33)

extern "C" int esr_mlp_fwd(const esr_mlp_desc_t *d, const void *image, const void *x, int64_t row_begin,
                           int64_t row_end, int64_t m_total, float *y, void *hidden, esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= m_total);
  if (row_end == row_begin) return ESR_OK;
  ESR_CHECK_ARG(image && x && y);
  const MlpLayout L = layout_of(d);
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(d))
    return tc_fwd(d, (const uint8_t *)image + tc_image_off(d), x, row_begin, row_end, m_total, y, hidden, st);
#define FWD_CALL(...) run_fwd<__VA_ARGS__>(L, d, image, x, row_begin, row_end, m_total, y, hidden, st)
  if (d->k0 == 96 && d->width == 192 && d->n_hidden == 3) return FWD_CALL(96, 192, 3, 56);
  if (d->k0 == 48 && d->width == 192 && d->n_hidden == 1) return FWD_CALL(48, 192, 1, 40);
#undef FWD_CALL
  set_error("esr_mlp_fwd: MLP shape k0=%d width=%d hidden=%d is not instantiated", d->k0, d->width, d->n_hidden);
  return ESR_ERR_BAD_ARG;
}

extern "C" int esr_mlp_bwd(const esr_mlp_desc_t *d, const void *image, const void *x, const float *y,
                           const float *d_y, int64_t row_begin, int64_t row_end, int64_t m_total, const void *hidden,
                           void *d_z, float *d_z_out, float *d_x, int dx_cols, int accumulate, float *grad_flat,
                           esr_stream_t stream) {
  if (int e = check_desc(d)) return e;
  ESR_CHECK_ARG(row_begin >= 0 && row_end >= row_begin && row_end <= m_total);
  if (row_end == row_begin) return ESR_OK;
  ESR_CHECK_ARG(image && x && y && d_y && hidden && d_z && d_z_out);
  ESR_CHECK_ARG(!d_x || (dx_cols > 0 && dx_cols % 2 == 0));
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

// mlp_mma.cu — C-ABI entry points of the MLP stage (esr_mlp_*) and the weight-gradient GEMMs.
//
//   forward / data gradient : fused tcgen05 layer chains, mlp_tc.cu
//   weight gradient         : dW_l += dZ_l^T . In_l as a split-K (over samples) GEMM; v1 = warp-level
//                             mma.sync.m16n8k16 with ldmatrix.trans operands, fp32 partials reduced into the flat
//                             gradient with RED.
// Hidden activations H_l and their cotangents dZ_l travel between the kernels in the TILED layout of
// mlp_layout.cuh (act_chunk_index): per 128-row tile, [24 feature chunks][128 rows][8 bf16], so that one
// epilogue thread per row writes / reads 16-byte chunks that are contiguous across the warp.
#include <stdlib.h>

#include "common.cuh"
#include "mlp_layout.cuh"

using namespace esr;

namespace {

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

template <int NT>
ESR_D void zero_acc(float (&acc)[NT][4]) {
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
}

constexpr int MLP_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// backward, weight gradient: dW[o][i] += sum_m dZ[m][o] * In[m][i];  db[o] += sum_m dZ[m][o]
// CTA = 8 warps as 4 (o) x 2 (i); split-K over the sample dimension across CTAs.
// ------------------------------------------------------------------------------------------------
constexpr int WG_KSTEP = 32;  // samples per pipeline stage

// dz: tiled activation layout; in: tiled (IN_TILED, a hidden layer) or row-major [m][KIN] (the MLP input x)
template <int W, int KIN, bool IN_TILED>
__global__ void __launch_bounds__(MLP_THREADS, 1)
    k_mlp_wgrad(const __nv_bfloat16 *__restrict__ dz, const __nv_bfloat16 *__restrict__ in, int64_t row_begin,
                int64_t row_end, float *__restrict__ gW /* [W][KIN] */, float *__restrict__ gb /* [W] */) {
  static_assert(W == ACT_W, "tiled activation layout is defined for the 192-wide hidden layers");
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
      cp_async16(smem_u32(&s_dz[buf][rr * SZ + q * 8]), dz + (ok ? act_chunk_index(m0 + rr, q) * 8 : 0), ok);
    }
    for (int c = threadIdx.x; c < WG_KSTEP * CI; c += MLP_THREADS) {
      const int rr = c / CI, q = c - rr * CI;
      const bool ok = m0 + rr < row_end;
      const int64_t src = IN_TILED ? act_chunk_index(m0 + rr, q) * 8 : (m0 + rr) * KIN + q * 8;
      cp_async16(smem_u32(&s_in[buf][rr * SI + q * 8]), in + (ok ? src : 0), ok);
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

// mlp_tc.cu — the radiance / tone-map MLPs (pbr/module.py:6-39) as fused layer chains on the 5th-gen tensor
// cores: tcgen05.mma issued by one thread per CTA, accumulators in TMEM, every weight matrix of the net
// resident in shared memory for the life of the (persistent) CTA, hidden activations never leaving the SM
// between layers: the epilogue warps read the fp32 accumulator tile out of TMEM (tcgen05.ld), apply
// bias + ReLU (forward) or the ReLU mask (data gradient), pack to bf16 and write the result back INTO TMEM
// (tcgen05.st) where the next layer's MMA consumes it as its A operand (the .ts form: A from tensor memory,
// B from shared memory).  Only the first layer's A tile (the encoded feature rows / the output cotangent)
// comes from shared memory.
//
//   forward : x[128,K0] -> (192 x NH, ReLU) -> n_out, softplus|sigmoid;  saves H_l (bf16) for backward
//   dgrad   : d_y -> dZ_out -> (W^T chain, ReLU masks from H_l) -> d_x;   saves dZ_l (bf16) for wgrad
//
// Tile = 128 rows (UMMA M = 128, cta_group::1): TMEM lane = row, TMEM column = feature.
// Operand layout in shared memory: K-major, no swizzle, "chunk-major": element (r, k) of an [R x K] matrix
// lives at byte  (k/8) * (R*16) + r*16 + (k%8)*2  — i.e. 8x8 core matrices of 128 contiguous bytes, the two
// K-halves of one MMA (K = 16) R*16 bytes apart (descriptor LBO), consecutive 8-row groups 128 bytes apart
// (descriptor SBO).  The global weight image built by tc_pack has exactly this byte order, so staging is a
// straight copy, and an epilogue thread (= one row) writing one 16-byte chunk per k-chunk is conflict-free.
#include <stdlib.h>

#include "mlp_layout.cuh"

using namespace esr;

namespace {

constexpr int TC_W = 192;         // hidden width
constexpr int TC_TM = 128;        // rows per tile
constexpr int TC_EPI_WARPS = 16;  // epilogue warps: TMEM lane quadrant w % 4, column group w / 4 (48 columns each)
constexpr int TC_GCOLS = TC_W / (TC_EPI_WARPS / 4);   // 48
constexpr int TC_THREADS = 32 * (TC_EPI_WARPS + 1);   // + warp 8: MMA issuer / TMEM owner
constexpr int TC_NOUT_PAD = 16;   // output layer rows padded to the minimum UMMA N for M = 128

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
ESR_D uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

ESR_D void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
ESR_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ESR_D void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// generic-proxy writes to shared memory (st.shared / cp.async) -> visible to the async proxy (tcgen05.mma)
ESR_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
ESR_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ESR_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

ESR_D void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ESR_D void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
ESR_D uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D f32, A/B bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
}

ESR_D void mma_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
ESR_D void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
ESR_D void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// tcgen05.ld 32x32b.x16 / .x32: thread t of warp w receives columns [c, c+N) of TMEM lane 32*(w%4)+t
ESR_D void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
ESR_D void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
ESR_D void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
ESR_D void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

ESR_D void cp_async16_zfill(uint32_t dst, const void *src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
ESR_D void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

ESR_D uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}
// relu(lo), relu(hi) -> packed bf16 pair in ONE conversion instruction (cvt.rn.relu: negative inputs and -0 give +0,
// so clamping before or after the rounding is the same value); the epilogues issue it 24 times per thread per layer
ESR_D uint32_t pack2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// (a0, a1) += (b.x, b.y) as one packed FADD2 (add.rn.f32x2, sm_100): the bias add of a column pair
ESR_D void add2(float &a0, float &a1, float2 b) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b.x), "f"(b.y));
}
// 0xffff per half of a packed bf16 pair whose ReLU-mask bit is set: bits (j, 16 + j) of `mk` are moved onto the sign
// positions of bytes 0 and 2 (shift), then PRMT replicates those signs over bytes (0,1) and (2,3) — two instructions
// instead of shift + and + multiply
ESR_D uint32_t pair_mask(uint32_t mk, int j) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %1, 0xAA88;" : "=r"(d) : "r"(mk << (7 - j)));
  return d;
}
// ------------------------------------------------------------------------------------------------
// "x2" forward (esr_mlp_desc_t::precision = 1): fp16 hi + lo operand pairs, three MMAs per product.
// ------------------------------------------------------------------------------------------------
// The weights are scaled by 2^4 before the split so that the lo part of a typical |w| ~ 0.05 is a NORMAL fp16 number
// (the pair then carries ~24 bits); the epilogues undo the scale in the fused multiply-add that applies the bias.
constexpr float TC_WSCALE = 16.f;
// instruction descriptor: D f32, A/B fp16, both K-major; M = 128 (cta_group::1) or 256 (cta_group::2: 128 rows per CTA)
__host__ __device__ constexpr uint32_t make_idesc_h(int n, int m = TC_TM) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
ESR_D uint32_t pack2h(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// (saturating: an activation beyond fp16's 65504 is clamped, not turned into an infinity that would poison the row)
ESR_D uint32_t pack2h_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// the same with overflow clamped to +-65504 instead of +-inf (stored cotangents: a runaway value must not poison a sum)
ESR_D uint32_t pack2h_sat(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
ESR_D float2 unpack2h(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2 *>(&v)); }
// relu(z0), relu(z1) as an fp16 pair hi plus the fp16 pair lo of what the first rounding lost: hi + lo carries 22
// significant bits (an absolute error of 2^-25 where lo is subnormal, i.e. for values below 0.25)
ESR_D void split2h_relu(float z0, float z1, uint32_t &hi, uint32_t &lo) {
  hi = pack2h_relu(z0, z1);
  const float2 h = unpack2h(hi);
  lo = pack2h(fmaxf(z0, 0.f) - h.x, fmaxf(z1, 0.f) - h.y);
}
ESR_D void split2h(float z0, float z1, uint32_t &hi, uint32_t &lo) {
  hi = pack2h(z0, z1);
  const float2 h = unpack2h(hi);
  lo = pack2h(z0 - h.x, z1 - h.y);
}
// a packed bf16 pair as the same two values in fp16 (exact for 2^-14 <= |v| < 65504; below that the error is < 2^-25)
ESR_D uint32_t bf2_to_h2(uint32_t w) { return pack2h(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
// (a0, a1) = (a0, a1) * s + (b.x, b.y) as one packed FFMA2 (fma.rn.f32x2, sm_100)
ESR_D void fma2(float &a0, float &a1, float s, float2 b) {
  asm("{\n\t.reg .b64 ra, rs, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rs, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\t"
      "fma.rn.f32x2 ra, ra, rs, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(s), "f"(b.x), "f"(b.y));
}
ESR_D void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// ---- CTA pair (cta_group::2): one MMA spans the two SMs of a cluster of 2; each CTA supplies its own 128 rows of A /
// D (own TMEM) and HALF of the B rows (own shared memory), the leader CTA (cluster rank 0) issues ----
ESR_D uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
ESR_D void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in the CTA of cluster rank `rank`
ESR_D uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrival on an mbarrier of the leader CTA, possibly from the peer.  Default semantics (release at CTA scope): what the
// barrier hands over is TENSOR memory written by tcgen05.st, ordered by tcgen05.wait::st + tcgen05.fence::before_thread_sync
// on this side and tcgen05.fence::after_thread_sync on the issuer's; no generic-proxy data crosses.  (A .release.cluster
// arrive compiles to a membar that waits for every outstanding global store of the warp — the saved activations — and
// was 30 % of this kernel's stall samples, profiles/r02b.)
ESR_D void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
ESR_D void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // waits for arrivals that may come from the peer CTA
  mbar_wait(bar, parity);
}
ESR_D void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
ESR_D void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
ESR_D void mma_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier at the same shared-memory offset of BOTH CTAs receives one arrival once every MMA issued so far has completed
ESR_D void mma_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// image layout (bytes).  Every matrix is bf16 chunk-major [K/8][R][8].
// ------------------------------------------------------------------------------------------------
struct TcLayout {
  int k0, NH, dxn, x2;
  __host__ __device__ int64_t mat_bytes(int rows, int k) const { return (int64_t)rows * k * 2; }
  // forward section: W0 [W x k0], W_1..W_{NH-1} [W x W], Wo [16 x W], bias f32 [NH*W + 16]
  __host__ __device__ int64_t f_w0() const { return 0; }
  __host__ __device__ int64_t f_wh(int l) const { return mat_bytes(TC_W, k0) + (int64_t)(l - 1) * mat_bytes(TC_W, TC_W); }
  __host__ __device__ int64_t f_wo() const { return f_wh(NH); }
  __host__ __device__ int64_t f_bias() const { return f_wo() + mat_bytes(TC_NOUT_PAD, TC_W); }
  __host__ __device__ int64_t f_bytes() const { return f_bias() + (int64_t)(NH * TC_W + TC_NOUT_PAD) * 4; }
  // dgrad section (transposes): WoT [W x 16], W_{l}T [W x W] for l = NH-1 .. 1 (stored in index order l-1), W0T [dxn x W]
  __host__ __device__ int64_t b_wo() const { return 0; }
  __host__ __device__ int64_t b_wh(int l) const { return mat_bytes(TC_W, TC_NOUT_PAD) + (int64_t)(l - 1) * mat_bytes(TC_W, TC_W); }
  __host__ __device__ int64_t b_w0() const { return b_wh(NH); }
  __host__ __device__ int64_t b_bytes() const { return b_w0() + mat_bytes(dxn, TC_W); }
  __host__ __device__ int64_t bwd_off() const { return (f_bytes() + 127) / 128 * 128; }
  // x2 section (precision 1): fp16 hi / lo parts of the forward matrices, scaled by TC_WSCALE.
  //   radiance chains (k0 = 96, CTA pair): per cluster rank r the N-half r of every matrix, parts adjacent:
  //     [W0 p0 | W0 p1] [W1 p0 | W1 p1] ... [Wo p0 | Wo p1]  (matrix halves [W/2 x K] / [8 x W], chunk-major), then the
  //     second rank, then the f32 biases — byte for byte the shared-memory image of k_mlp_fwd_x2
  //   tone-map net (k0 = 48, one CTA): [W0 p0 | W0 p1 | Wo p0 | Wo p1] full matrices
  __host__ __device__ int64_t x2_off() const { return bwd_off() + (b_bytes() + 127) / 128 * 128; }
  __host__ __device__ int64_t x2_w0_part() const { return (k0 == 96 ? TC_W / 2 : TC_W) * (int64_t)k0 * 2; }
  __host__ __device__ int64_t x2_wh_part() const { return (TC_W / 2) * (int64_t)TC_W * 2; }
  __host__ __device__ int64_t x2_wo_part() const { return (k0 == 96 ? TC_NOUT_PAD / 2 : TC_NOUT_PAD) * (int64_t)TC_W * 2; }
  __host__ __device__ int64_t x2_wh(int l) const { return 2 * x2_w0_part() + (int64_t)(l - 1) * 2 * x2_wh_part(); }
  __host__ __device__ int64_t x2_wo() const { return x2_wh(NH); }
  __host__ __device__ int64_t x2_rank_bytes() const { return x2_wo() + 2 * x2_wo_part(); }
  __host__ __device__ int64_t x2_bias() const { return (k0 == 96 ? 2 : 1) * x2_rank_bytes(); }
  __host__ __device__ int64_t x2_bytes() const { return x2_bias() + (int64_t)(NH * TC_W + TC_NOUT_PAD) * 4; }
  __host__ __device__ int64_t total() const { return x2_off() + (x2 ? (x2_bytes() + 127) / 128 * 128 : 0); }
  // precision 1: the transposed matrices of the data-gradient chain are fp16 (k_mlp_dgrad_tc H16, k_tonemap_bwd_fused X2)
  __host__ __device__ bool bwd_h16() const { return x2; }
};

static TcLayout tc_layout(const esr_mlp_desc_t *d) {
  return TcLayout{d->k0, d->n_hidden, d->k0 == 96 ? 64 : 48, d->precision == 1};
}

ESR_HD int64_t chunk_index(int rows, int r, int k) { return ((int64_t)(k >> 3) * rows + r) * 8 + (k & 7); }

// 16-bit store of a transposed (data-gradient) matrix element: bf16, or fp16 for the fp16 data-gradient chain
struct BwdElem {
  uint16_t *p;
  bool h16;
  __device__ void set(int64_t i, float v) const {
    p[i] = h16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16(v));
  }
};

__global__ void k_tc_pack(TcLayout T, MlpLayout L, const float *__restrict__ flat, uint8_t *__restrict__ image) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = TC_W, NH = T.NH;
  __nv_bfloat16 *f = reinterpret_cast<__nv_bfloat16 *>(image);
  const BwdElem b{reinterpret_cast<uint16_t *>(image + T.bwd_off()), T.bwd_h16()};
  // one thread per (layer-matrix element) of the padded logical matrices; enumerate W0, Wh.., Wo in turn
  int64_t e = i;
  // W0 [W x k0]
  if (e < (int64_t)W * T.k0) {
    const int o = (int)(e / T.k0), in = (int)(e % T.k0);
    const float v = flat[L.flat_w(0) + e];
    f[T.f_w0() / 2 + chunk_index(W, o, in)] = __float2bfloat16(v);
    if (in < T.dxn) b.set(T.b_w0() / 2 + chunk_index(T.dxn, in, o), v);
    return;
  }
  e -= (int64_t)W * T.k0;
  for (int l = 1; l < NH; ++l) {
    if (e < (int64_t)W * W) {
      const int o = (int)(e / W), in = (int)(e % W);
      const float v = flat[L.flat_w(l) + e];
      f[T.f_wh(l) / 2 + chunk_index(W, o, in)] = __float2bfloat16(v);
      b.set(T.b_wh(l) / 2 + chunk_index(W, in, o), v);
      return;
    }
    e -= (int64_t)W * W;
  }
  // Wo [16 x W] (flat copy holds 8 rows; rows >= 8 are zero)
  if (e < (int64_t)TC_NOUT_PAD * W) {
    const int o = (int)(e / W), in = (int)(e % W);
    const float v = o < 8 ? flat[L.flat_w(NH) + (int64_t)o * W + in] : 0.f;
    f[T.f_wo() / 2 + chunk_index(TC_NOUT_PAD, o, in)] = __float2bfloat16(v);
    b.set(T.b_wo() / 2 + chunk_index(W, in, o), v);
    return;
  }
  e -= (int64_t)TC_NOUT_PAD * W;
  // biases
  if (e < (int64_t)NH * W + TC_NOUT_PAD) {
    float *bias = reinterpret_cast<float *>(image + T.f_bias());
    float v;
    if (e < (int64_t)NH * W) {
      const int l = (int)(e / W);
      v = flat[L.flat_b(l) + (e - (int64_t)l * W)];
    } else {
      const int o = (int)(e - (int64_t)NH * W);
      v = o < 8 ? flat[L.flat_b(NH) + o] : 0.f;
    }
    bias[e] = v;
  }
}

// x2 section: one thread per element of the (padded) forward matrices and biases
__global__ void k_tc_pack_x2(TcLayout T, MlpLayout L, const float *__restrict__ flat, uint8_t *__restrict__ image) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int W = TC_W, NH = T.NH;
  uint8_t *sec = image + T.x2_off();
  const bool pair = T.k0 == 96;
  // element (o, in) of a matrix with `rows` output rows, K = kdim, stored at byte `base` of a rank section
  auto put = [&](int64_t base, int64_t part_bytes, int rows, int o, int in, float v) {
    const int half = pair ? rows / 2 : rows;
    const int rank = o / half, oh = o % half;
    const __half hi = __float2half_rn(v * TC_WSCALE);
    const __half lo = __float2half_rn(v * TC_WSCALE - __half2float(hi));
    __half *dst = reinterpret_cast<__half *>(sec + (int64_t)rank * T.x2_rank_bytes() + base);
    dst[chunk_index(half, oh, in)] = hi;
    dst[part_bytes / 2 + chunk_index(half, oh, in)] = lo;
  };
  int64_t e = i;
  if (e < (int64_t)W * T.k0) {
    const int o = (int)(e / T.k0), in = (int)(e % T.k0);
    put(0, T.x2_w0_part(), W, o, in, flat[L.flat_w(0) + e]);
    return;
  }
  e -= (int64_t)W * T.k0;
  for (int l = 1; l < NH; ++l) {
    if (e < (int64_t)W * W) {
      put(T.x2_wh(l), T.x2_wh_part(), W, (int)(e / W), (int)(e % W), flat[L.flat_w(l) + e]);
      return;
    }
    e -= (int64_t)W * W;
  }
  if (e < (int64_t)TC_NOUT_PAD * W) {
    const int o = (int)(e / W), in = (int)(e % W);
    put(T.x2_wo(), T.x2_wo_part(), TC_NOUT_PAD, o, in, o < 8 ? flat[L.flat_w(NH) + (int64_t)o * W + in] : 0.f);
    return;
  }
  e -= (int64_t)TC_NOUT_PAD * W;
  if (e < (int64_t)NH * W + TC_NOUT_PAD) {
    float *bias = reinterpret_cast<float *>(sec + T.x2_bias());
    float v;
    if (e < (int64_t)NH * W) {
      const int l = (int)(e / W);
      v = flat[L.flat_b(l) + (e - (int64_t)l * W)];
    } else {
      const int o = (int)(e - (int64_t)NH * W);
      v = o < 8 ? flat[L.flat_b(NH) + o] : 0.f;
    }
    bias[e] = v;
  }
}

// cooperative straight copy global -> shared (16-byte words)
ESR_D void stage_bytes(uint8_t *dst, const uint8_t *__restrict__ src, int64_t bytes) {
  const uint4 *s = reinterpret_cast<const uint4 *>(src);
  uint4 *d = reinterpret_cast<uint4 *>(dst);
  for (int64_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = __ldg(s + i);
}

ESR_D float act_fwd(float z, int act) {
  if (act == 1) return z > 20.f ? z : log1pf(expf(z));  // softplus(beta=1, threshold=20)
  if (act == 2) return 1.f / (1.f + expf(-z));
  return z;
}

// TMEM column map (512 columns allocated): two accumulator regions D0 [0,192) / D1 [192,384) used by alternate layers
// (so the next layer's MMA can start while the epilogue still reads the current accumulator), bf16 A operand
// [384,480).  The narrow last accumulator of a chain (output layer / d_x) lives in whichever D region is free.
constexpr uint32_t TM_D0 = 0, TM_A = 384, TM_COLS = 512;
ESR_D uint32_t tm_d(int i) { return TM_D0 + (uint32_t)(i & 1) * TC_W; }

ESR_D void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Chunk pipeline.  An epilogue thread owns 48 accumulator columns = three 16-column chunks cc = 0..2; column group g
// (= warp / 4) chunk cc is K-step s = 3 g + cc of the next layer's MMA.  As soon as every warp has written chunk cc of
// the new A operand it arrives on bar_chunk[cc] and the issuer fires the four K-steps {cc, 3 + cc, 6 + cc, 9 + cc}:
// two thirds of a layer's MMA time hides under the epilogue that produces its operand.
ESR_D void chunk_ready(uint32_t bar_chunk, unsigned lane) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_chunk);
}

// ------------------------------------------------------------------------------------------------
// forward chain
// ------------------------------------------------------------------------------------------------
template <int K0, int NH>
struct FwdSm {
  static constexpr int w0 = 0;
  static constexpr int wh = w0 + TC_W * K0 * 2;
  static constexpr int wo = wh + (NH - 1) * TC_W * TC_W * 2;
  static constexpr int bias = wo + TC_NOUT_PAD * TC_W * 2;
  static constexpr int weights_bytes = bias + (NH * TC_W + TC_NOUT_PAD) * 4;   // == TcLayout::f_bytes()
  static constexpr int x = (weights_bytes + 127) / 128 * 128;
  static constexpr int bar = x + TC_TM * K0 * 2;   // bar_mma, bar_chunk[3] (8 B each), TMEM slot, (OVL) bar_x, bar_first
  static constexpr int bytes = bar + 64;
};

// Epilogue thread geometry (16 warps): TMEM lane quadrant = warp % 4 (hardware rule for tcgen05.ld/st), column
// group = warp / 4: thread (q, lane, grp) owns row 32q + lane and feature columns [48 grp, 48 grp + 48).
// Four warps per scheduler hide the TMEM / global latencies of one another.
struct EpiThread {
  int row_in_tile, grp;
  uint32_t lane_base;
  __device__ EpiThread(unsigned warp, unsigned lane)
      : row_in_tile(32 * (warp & 3) + lane), grp(warp >> 2), lane_base((32u * (warp & 3)) << 16) {}
};

// tone-map positional encoding of one channel (voxurff.py:783-788) as the 16 bf16 columns [16 c, 16 c + 16) of the
// internal tone-map row: lin, sin(lin 2^f) f<5, cos(lin 2^f) f<5, 5 zeros — two 16-byte chunks; s / c return the
// f32 sines / cosines (the fused backward needs them again)
ESR_D void tonemap_pe_channel(float x, uint4 &lo, uint4 &hi, float (&sn)[5], float (&cs)[5]) {
#pragma unroll
  for (int f = 0; f < 5; ++f)   // SFU sine / cosine: |error| ~ |y| 2^-23, far below the bf16 rounding that follows
    __sincosf(__fmul_rn(x, (float)(1 << f)), &sn[f], &cs[f]);
  lo = make_uint4(pack2(x, sn[0]), pack2(sn[1], sn[2]), pack2(sn[3], sn[4]), pack2(cs[0], cs[1]));
  hi = make_uint4(pack2(cs[2], cs[3]), pack2(cs[4], 0.f), 0u, 0u);
}

// sin / cos of y to ~2^-21 ABSOLUTE at a fifth of sincosf's instructions: two-constant Cody-Waite reduction of y to
// [-pi, pi] (exact to ~1e-7 for the |y| <= ~100 the encoding sees: lin in [0, ~3] times 2^f <= 16), then the SFU sine /
// cosine, whose error on that interval is 2^-21.4.  The feature error (~4e-7) moves a pre-activation by as much relative
// to its scale: a ReLU mask in ~3e-7 — below what fp32 summation order already does (tests/test_gpu_mlp.py).
ESR_D void sincos_reduced(float y, float &s, float &c) {
  const float k = rintf(y * 0.15915494309189535f);          // y / 2 pi
  float r = fmaf(k, -6.28318548202514648f, y);                // 2 pi = hi + lo, hi = fp32(2 pi)
  r = fmaf(k, 1.74845553e-7f, r);                             // hi - 2 pi = 1.74845553e-7
  s = __sinf(r);
  c = __cosf(r);
}
// x2 variant: the same 16 columns as fp16 hi + lo pairs (chunks hl / hh and ll / lh)
ESR_D void tonemap_pe_channel_x2(float x, uint4 &hl, uint4 &hh, uint4 &ll, uint4 &lh, float (&sn)[5], float (&cs)[5]) {
#pragma unroll
  for (int f = 0; f < 5; ++f) sincos_reduced(__fmul_rn(x, (float)(1 << f)), sn[f], cs[f]);
  uint32_t h[6], l[6];
  split2h(x, sn[0], h[0], l[0]);
  split2h(sn[1], sn[2], h[1], l[1]);
  split2h(sn[3], sn[4], h[2], l[2]);
  split2h(cs[0], cs[1], h[3], l[3]);
  split2h(cs[2], cs[3], h[4], l[4]);
  split2h(cs[4], 0.f, h[5], l[5]);
  hl = make_uint4(h[0], h[1], h[2], h[3]);
  hh = make_uint4(h[4], h[5], 0u, 0u);
  ll = make_uint4(l[0], l[1], l[2], l[3]);
  lh = make_uint4(l[4], l[5], 0u, 0u);
}

// NO = compile-time bound on the real output columns (3: radiance / tone-map / emission nets, 8: the 5-output BRDF net)
// XSRC = 0: x rows come tiled from global memory; 1 (K0 = 48 only): `x` is the f32 [m,3] linear radiance and the CTA
// computes the tone-map encoding of its tile itself (no feature rows in HBM at all)
// OVL = tile overlap (opt-in, ESR_MLP_TILE_OVERLAP=1): the next tile's layer-0 MMA is issued behind this tile's
// output-layer MMA and runs under the output epilogue (an x-ready mbarrier replaces the per-tile __syncthreads).  Its
// commit goes to its OWN mbarrier (bar_first), so that no mbarrier ever completes two phases without every waiting warp
// having observed the first: bar (layers 1.. + output) completes again only after all 16 warps have arrived on a chunk
// barrier of the next tile, which each does after its own wait on the output phase; bar_first completes again only
// after all 16 warps have arrived on bar_x, which each does after its own wait on bar_first.
template <int K0, int NH, int NO, int XSRC = 0, bool OVL = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_mlp_fwd_tc(const uint8_t *__restrict__ image, const __nv_bfloat16 *__restrict__ x, int64_t row_begin,
                 int64_t row_end, int64_t m_total, float *__restrict__ y, __nv_bfloat16 *__restrict__ hidden,
                 int64_t save_begin, int n_out, int act) {
  extern __shared__ __align__(128) uint8_t smem[];
  using S = FwdSm<K0, NH>;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_epi = warp < TC_EPI_WARPS, is_issuer = warp == TC_EPI_WARPS;
  const uint32_t sbase = smem_addr(smem);
  const uint32_t bar = sbase + S::bar, bar_chunk = bar + 8;
  [[maybe_unused]] const uint32_t bar_x = bar + 40, bar_first = bar + 48;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::bar + 32);
  // OVL: layer 0 of the next tile writes D0 while the output epilogue still reads its accumulator, which must
  // therefore be D1 — and D0 the accumulator of the LAST hidden layer, read before the arrival on bar_x
  static_assert(!OVL || (NH & 1), "tile overlap: odd number of hidden layers (the output accumulator lives in D1)");

  stage_bytes(smem, image, S::weights_bytes);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) mbar_init(bar_chunk + 8 * c, TC_EPI_WARPS);
    if constexpr (OVL) {
      mbar_init(bar_x, TC_EPI_WARPS);
      mbar_init(bar_first, 1);
    }
    fence_mbar_init();
  }
  if (is_issuer) tmem_alloc(smem_addr(tmem_slot), TM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float *sbias = reinterpret_cast<const float *>(smem + S::bias);
  uint32_t cphase = 0;   // parity of the chunk barriers (issuer)
  [[maybe_unused]] uint32_t xphase = 0;   // OVL: parity of bar_x (issuer)
  [[maybe_unused]] uint32_t fphase = 0;   // OVL: parity of bar_first (epilogue warps)

  const int64_t n_tiles = (row_end - row_begin + TC_TM - 1) / TC_TM;
  const EpiThread et(warp, lane);
  const int t = et.row_in_tile;
  uint32_t phase = 0;

  auto load_x = [&](int64_t tile) {  // 4 epilogue threads per row: 16-byte chunks c = grp, grp + 4, ...
    const int64_t row = row_begin + tile * TC_TM + t;
    const bool ok = row < row_end;
    if constexpr (XSRC == 1) {   // column group g < 3 encodes channel g of its row into chunks 2 g, 2 g + 1
      if (et.grp < 3) {
        const float v = ok ? __ldg(reinterpret_cast<const float *>(x) + 3 * row + et.grp) : 0.f;
        uint4 lo, hi;
        float sn[5], cs[5];
        tonemap_pe_channel(v, lo, hi, sn, cs);
        if (!ok) lo = hi = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(smem + S::x + (2 * et.grp) * (TC_TM * 16) + t * 16) = lo;
        *reinterpret_cast<uint4 *>(smem + S::x + (2 * et.grp + 1) * (TC_TM * 16) + t * 16) = hi;
      }
    } else {
      const uint4 *x4 = reinterpret_cast<const uint4 *>(x);  // tiled layout: a warp reads 512 contiguous bytes per chunk
#pragma unroll
      for (int c = et.grp; c < K0 / 8; c += 4)
        cp_async16_zfill(sbase + S::x + c * (TC_TM * 16) + t * 16, x4 + tiled_chunk_index(ok ? row : row_begin, c, K0 / 8), ok);
    }
  };

  if (is_epi && blockIdx.x < n_tiles) load_x(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = row_begin + tile * TC_TM + t;
    const bool valid = is_epi && row < row_end;
    const bool save = valid && hidden && row >= save_begin;  // rows the backward pass will visit
    const bool save_w = __any_sync(FULL, save);               // warp-uniform (the issuer warp: false)
    // ---- layer 0: A = x tile (shared), B = W0 ----
    [[maybe_unused]] const bool first = tile == (int64_t)blockIdx.x, more = tile + gridDim.x < n_tiles;
    if (!OVL || first) {   // OVL, later tiles: layer 0 was issued behind the previous tile's output layer (below)
      if (is_epi) {
        cp_async_wait_all();
        fence_proxy_async();
      }
      tc_fence_before();
      __syncthreads();
    }
    auto issue_layer0 = [&]() {
#pragma unroll
      for (int s = 0; s < K0 / 16; ++s)
        mma_ss(tmem + tm_d(0), make_desc(sbase + S::x + 2 * s * (TC_TM * 16), TC_TM * 16, 128),
               make_desc(sbase + S::w0 + 2 * s * (TC_W * 16), TC_W * 16, 128), make_idesc(TC_W), s > 0);
      mma_commit(OVL ? bar_first : bar);
    };
    if (is_issuer && lane == 0) {
      if (!OVL || first) {
        tc_fence_after();
        issue_layer0();
      }
      // the rest of the chain: layer l + 1 (or the output layer) is fed chunk by chunk as the epilogue of layer l
      // produces its A operand; its accumulator is the D region layer l does not use
#pragma unroll 1
      for (int l = 0; l < NH; ++l) {
        const bool last = l + 1 == NH;
        const uint32_t dst = tmem + tm_d(l + 1);
        const uint32_t wl = last ? sbase + S::wo : sbase + S::wh + l * (TC_W * TC_W * 2);
        const uint32_t rows16 = (last ? TC_NOUT_PAD : TC_W) * 16;
        const uint32_t idesc = last ? make_idesc(TC_NOUT_PAD) : make_idesc(TC_W);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          mbar_wait(bar_chunk + 8 * cc, cphase);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int s = 3 * g + cc;
            mma_ts(dst, tmem + TM_A + 8 * s, make_desc(wl + 2 * s * rows16, rows16, 128), idesc, (cc | g) != 0);
          }
        }
        mma_commit(bar);
        cphase ^= 1;
      }
      if constexpr (OVL) {
        if (more) {   // the next x tile has landed and every warp is done with D0: layer 0 of the next tile, now
          mbar_wait(bar_x, xphase);
          xphase ^= 1;
          tc_fence_after();
          issue_layer0();
        }
      }
    }
    if (is_epi) {
#pragma unroll 1
      for (int l = 0; l < NH; ++l) {
        if (OVL && l == 0) {
          mbar_wait(bar_first, fphase);
          fphase ^= 1;
        } else {
          mbar_wait(bar, phase);
          phase ^= 1;
        }
        tc_fence_after();
        if (l == 0 && tile + gridDim.x < n_tiles) load_x(tile + gridDim.x);  // x tile is free: prefetch the next one
        // bias + ReLU -> bf16 -> TMEM A operand (+ global copy, tiled layout, for the backward pass)
        const float *b = sbias + l * TC_W;
        uint4 *hl = hidden ? reinterpret_cast<uint4 *>(hidden + (int64_t)l * act_rows_padded(m_total) * TC_W) : nullptr;
        // chunk cc is converted while chunk cc + 1 is still coming out of TMEM (the next layer accumulates in the other D
        // region, so nothing overwrites the columns not yet read): the first K-steps of the next layer start a third of
        // an accumulator read after the commit instead of a whole one
        uint32_t r[3][16];
        const uint32_t d_cols = tmem + et.lane_base + tm_d(l) + TC_GCOLS * et.grp;
        tmem_ld16(d_cols, r[0]);
        uint32_t mask[2] = {0u, 0u};
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          tmem_ld_wait();
          if (cc < 2) tmem_ld16(d_cols + 16 * (cc + 1), r[cc + 1]);
          const int col0 = TC_GCOLS * et.grp + 16 * cc;
          uint32_t p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 bb = *reinterpret_cast<const float2 *>(b + col0 + 2 * j);
            float z0 = __uint_as_float(r[cc][2 * j]), z1 = __uint_as_float(r[cc][2 * j + 1]);
            add2(z0, z1, bb);
            p[j] = pack2_relu(z0, z1);
            // mask of the STORED activation (test the rounded word so forward and backward agree), 3 integer ops per
            // pair: a non-zero non-negative bf16 half plus 0x7fff carries into its top bit (halves are <= 0x7f80, so
            // the low half never carries into the high one); the two top bits land on mask bits (s, 16 + s),
            // s = 8 (cc & 1) + j  (mlp_layout.cuh).  A third of the epilogue's instructions: skipped by warps none of
            // whose rows is saved (inference; the off net's emission-on rows)
            if (save_w) {
              const uint32_t t = p[j] + 0x7fff7fffu;
              constexpr uint32_t one2 = 0x00010001u;
              mask[cc >> 1] |= (t >> (15 - 8 * (cc & 1) - j)) & (one2 << (8 * (cc & 1) + j));
            }
          }
          tmem_st8(tmem + et.lane_base + TM_A + col0 / 2, p);
          if (save) {  // the warp's 32 rows write 512 contiguous bytes per chunk (issued under the TMEM store latency)
            hl[act_chunk_index(row, col0 / 8)] = make_uint4(p[0], p[1], p[2], p[3]);
            hl[act_chunk_index(row, col0 / 8 + 1)] = make_uint4(p[4], p[5], p[6], p[7]);
          }
          chunk_ready(bar_chunk + 8 * cc, lane);
        }
        if (save) {
          uint2 *mb = reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(hidden) + act_mask_base_bytes(NH, m_total));
          mb[act_mask_index(l, act_rows_padded(m_total), row, et.grp)] = make_uint2(mask[0], mask[1]);
        }
      }
      if constexpr (OVL) {
        if (more) {   // hidden layers done: the prefetched x tile is complete, D0 has been read -> release layer 0 of the next tile
          cp_async_wait_all();
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x);
        }
      }
      // ---- output layer epilogue (column group 0 threads) ----
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      if (et.grp == 0) {
        uint32_t r[16];
        tmem_ld16(tmem + et.lane_base + tm_d(NH), r);
        tmem_ld_wait();
        if (valid) {
          const float *bo = sbias + NH * TC_W;
#pragma unroll
          for (int c = 0; c < NO; ++c)
            if (c < n_out) y[row * n_out + c] = act_fwd(__uint_as_float(r[c]) + bo[c], act);
        }
      }
      tc_fence_before();   // the next tile's first MMA overwrites D0 / the region read above
    }
  }
  tc_fence_before();
  __syncthreads();
  if (is_issuer) tmem_dealloc(tmem, TM_COLS);
}

// ------------------------------------------------------------------------------------------------
// forward chain, "x2" precision (esr_mlp_desc_t::precision = 1), on a CTA PAIR
// ------------------------------------------------------------------------------------------------
// Why.  The reference's nets are fp32 (pbr/module.py:6-83).  With bf16 operands ~0.4 % of the hidden pre-activations
// land on the other side of zero, each flipped ReLU mask changes one sample's whole contribution to the weight /
// colour-grid gradients, and those gradients end up 2-5 % (relative L2) away from the reference's.  Carrying every
// forward operand — input features, weights, hidden activations — as an fp16 pair hi + lo (22 significant bits) and
// forming each product as hi.hi + hi.lo + lo.hi (fp32 accumulation in TMEM) makes the pre-activations, hence the
// masks and the outputs, fp32-class; the backward kernels run on single fp16 operands (their rounding errors are
// zero-mean and average out over the samples): every parameter gradient within 1e-2 of the reference's (measured
// <= 2.5e-3 max-norm, <= 1e-3 relative L2), outputs ~1e-6.  Three times the MMA work of the bf16 chain, forward only.
//
// How.  Two parts of every weight matrix do not fit one SM's shared memory (2 x 190 KB), so the chain runs on the two
// SMs of a cluster of 2 with cta_group::2 MMAs (M = 256): CTA r holds rows [96 r, 96 r + 96) of every hidden matrix
// (both parts: 180 KB) as its half of the B operand and the 128 rows [256 t + 128 r, +128) of the pair's tile t in its
// own TMEM; the leader CTA's issuer thread drives both tensor cores.  The A operand always comes from TMEM: the hidden
// layers' epilogues write the (hi, lo) pair of the activation there, and layer 0's operand — the feature rows, stored
// by the encoder as an fp16 tile (the one the weight-gradient GEMM reads) plus an fp16 tile of what that rounding lost —
// is loaded into registers a layer ahead and stored to TMEM by the same threads (no x tile in shared memory).
// TMEM columns: two accumulator regions D0 [0,192) / D1 [192,384) used by alternate MMA groups, X_hi [384,432) /
// X_lo [432,480) layer 0's operand.  The A operand of a hidden layer is written IN PLACE over the accumulator it was made
// from: a thread's 16-column accumulator chunk (K-step s = 3 g + c of the next layer) becomes 8 columns of fp16 hi and
// 8 of fp16 lo — the same 16 columns, which only that thread reads, so there is no cross-thread hazard and no barrier.
// Per tile the tensor cores run three MMA groups (layers 0, 1, 2); what keeps them busy between groups:
//   * an epilogue reads its 48 accumulator columns chunk by chunk (tcgen05.ld of chunk c + 1 in flight while chunk c is
//     converted), so the first K-steps of the next layer — which accumulate in the OTHER region — start a third of a
//     TMEM read after the commit, not a whole one;
//   * the OUTPUT layer (192 -> n_out <= 8) never goes to the tensor core: the last hidden epilogue holds relu(z) of its
//     48 columns in fp32 registers and forms its part of the n_out dot products there (Wo in fp32 shared memory), the
//     four column groups of a row meet in shared memory — no 16-wide MMA group, no commit round trip, no fourth epilogue;
//   * layer 0 of the NEXT tile is issued under that last epilogue: its operand region is free once this tile's layer 0
//     has committed, so the warps store the prefetched feature rows of the next tile there first thing and arrive, and
//     the MMAs (into the region the last layer's operand occupied: the tensor core runs MMAs in issue order) run while
//     the warps read, convert, save and reduce.
// Synchronisation: the chunk barriers live in the leader CTA and count the 32 epilogue warps of both CTAs (the peer's
// arrive through shared::cluster); every commit is multicast to the MMA barrier of both CTAs.
// Tried and withdrawn (round 2, measured at config 2 where this version takes 3.56 ms per step): issuing every layer as
// two (N = 96) or three (N = 64) column blocks with their own commits, so that a block's epilogue runs under the MMAs of
// the next — 3.97 ms and 4.37 ms.  The narrower MMAs do not run proportionally faster (a cta_group::2 MMA of N = 64 costs
// about what N = 128 does), which more than eats the overlap; N = 192 per instruction it stays.
constexpr uint32_t X2_D = 0, X2_X0 = 384, X2_X1 = 432;

template <int K0, int NH, int NO>
struct X2Sm {
  static constexpr int HALF = TC_W / 2, OHALF = TC_NOUT_PAD / 2;
  static constexpr int w0_part = HALF * K0 * 2, wh_part = HALF * TC_W * 2, wo_part = OHALF * TC_W * 2;
  static constexpr int w0 = 0;
  static constexpr int wh = w0 + 2 * w0_part;
  static constexpr int wo_img = wh + (NH - 1) * 2 * wh_part;          // offset of [Wo hi | Wo lo] in a rank section of the image
  static constexpr int rank_bytes = wo_img + 2 * wo_part;             // == TcLayout::x2_rank_bytes()
  static constexpr int wo32 = wo_img;                                  // shared memory: Wo as f32 [8][192] instead
  static constexpr int bias = wo32 + 8 * TC_W * 4;
  static constexpr int bias_bytes = (NH * TC_W + TC_NOUT_PAD) * 4;
  static constexpr int part = (bias + bias_bytes + 15) / 16 * 16;     // output partial sums f32 [2 tiles][4 groups][128 rows][NO]
  static constexpr int bar = part + 2 * 4 * TC_TM * NO * 4;           // bar_mma, bar_chunk[3] (8 B each), TMEM slot
  static constexpr int bytes = bar + 64;
};

template <int K0, int NH, int NO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    k_mlp_fwd_x2(const uint8_t *__restrict__ image_x2, const __nv_bfloat16 *__restrict__ x,
                 const __half *__restrict__ x_lo, int64_t row_begin, int64_t row_end, int64_t m_total,
                 float *__restrict__ y, __nv_bfloat16 *__restrict__ hidden, int64_t save_begin, int n_out, int act) {
  static_assert(K0 == 96, "x chunks per column group: K0 / 32");
  extern __shared__ __align__(128) uint8_t smem[];
  using S = X2Sm<K0, NH, NO>;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_epi = warp < TC_EPI_WARPS, is_issuer = warp == TC_EPI_WARPS;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sbase = smem_addr(smem);
  const uint32_t bar = sbase + S::bar, bar_chunk = bar + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::bar + 32);

  stage_bytes(smem, image_x2 + (int64_t)rank * S::rank_bytes, S::wo_img);   // this rank's halves of W0, W1, W2 (hi | lo)
  stage_bytes(smem + S::bias, image_x2 + 2 * (int64_t)S::rank_bytes, S::bias_bytes);
  {  // Wo (rows 0..7: rank 0's half of the padded 16) as f32 = (hi + lo) / scale, row-major [8][192]
    const __half *wo_hi = reinterpret_cast<const __half *>(image_x2 + S::wo_img), *wo_lo = wo_hi + S::wo_part / 2;
    float *w32 = reinterpret_cast<float *>(smem + S::wo32);
    for (int i = threadIdx.x; i < 8 * TC_W; i += blockDim.x) {
      const int o = i / TC_W, in = i % TC_W;
      const int64_t at = chunk_index(S::OHALF, o, in);
      w32[i] = (__half2float(wo_hi[at]) + __half2float(wo_lo[at])) * (1.f / TC_WSCALE);
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) mbar_init(bar_chunk + 8 * c, 2 * TC_EPI_WARPS);
    fence_mbar_init();
  }
  if (is_issuer) tmem_alloc2(smem_addr(tmem_slot), TM_COLS);
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();   // both CTAs: weights staged, mbarriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float *sbias = reinterpret_cast<const float *>(smem + S::bias);
  const float *wo32 = reinterpret_cast<const float *>(smem + S::wo32);
  float *part_buf = reinterpret_cast<float *>(smem + S::part);

  const int64_t n_pt = (row_end - row_begin + 2 * TC_TM - 1) / (2 * TC_TM);   // tiles of the pair: 256 rows
  const int64_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (is_epi) {
    const EpiThread et(warp, lane);
    const int t = et.row_in_tile;
    const uint32_t cbar_chunk = mapa_rank(bar_chunk, 0);   // the leader's chunk barriers
    const uint4 *x4 = reinterpret_cast<const uint4 *>(x), *l4 = reinterpret_cast<const uint4 *>(x_lo);
    uint32_t phase = 0;
    // layer-0 operand: column group g moves feature chunks 3 g .. 3 g + 2 (24 features) of its row, both tiles
    uint4 xb[3], xl[3];
    auto load_x = [&](int64_t pt) {
      const int64_t row = row_begin + (2 * pt + rank) * TC_TM + t;
      const bool ok = pt < n_pt && row < row_end;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int64_t at = tiled_chunk_index(ok ? row : row_begin, 3 * et.grp + i, K0 / 8);
        xb[i] = ok ? __ldg(x4 + at) : make_uint4(0u, 0u, 0u, 0u);
        xl[i] = ok ? __ldg(l4 + at) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto arrive_chunk = [&](int cc) {   // this warp's part of chunk cc of the A operand is in TMEM
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(cbar_chunk + 8 * cc);
    };
    auto store_x = [&]() {   // layer-0 operand -> TMEM (X_hi = the fp16 tile, X_lo = the residual tile); all three chunks
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint32_t col = 4 * (3 * et.grp + i);
        tmem_st4(tmem + et.lane_base + X2_X0 + col, xb[i].x, xb[i].y, xb[i].z, xb[i].w);
        tmem_st4(tmem + et.lane_base + X2_X1 + col, xl[i].x, xl[i].y, xl[i].z, xl[i].w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) mbar_arrive_cluster(cbar_chunk + 8 * cc);
      }
    };
    load_x(pair);
    if (pair < n_pt) store_x();
    uint32_t dsel = 0;   // accumulator region of the MMA group whose commit comes next
    for (int64_t pt = pair; pt < n_pt; pt += n_pairs) {
      const int64_t row = row_begin + (2 * pt + rank) * TC_TM + t;
      const bool valid = row < row_end;
      const bool save = valid && hidden && row >= save_begin;
      const bool save_w = __any_sync(FULL, save);
      const bool more = pt + n_pairs < n_pt;
#pragma unroll 1
      for (int l = 0; l < NH; ++l) {
        const bool last = l == NH - 1;
        // next tile's feature rows: a whole layer ahead of their use (requested only before the last layer's wait they
        // cost 0.2 ms per step: the store to TMEM right after that wait then stalls on them)
        if (l == NH - 2) load_x(pt + n_pairs);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        const float *b = sbias + l * TC_W;
        uint4 *hl = hidden ? reinterpret_cast<uint4 *>(hidden + (int64_t)l * act_rows_padded(m_total) * TC_W) : nullptr;
        uint32_t r[3][16];
        uint32_t mask[2] = {0u, 0u};
        float2 acc[NO];
#pragma unroll
        for (int c = 0; c < NO; ++c) acc[c] = make_float2(0.f, 0.f);
        const uint32_t d_cols = tmem + et.lane_base + X2_D + TC_W * dsel + TC_GCOLS * et.grp;
        dsel ^= 1;
        // layer 0's operand region is free (this tile's layer 0 committed long ago) and every warp is past the chunk
        // barriers' previous phase: the next tile's layer 0 is released now and runs under this epilogue
        if (last && more) store_x();
        tmem_ld16(d_cols, r[0]);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          tmem_ld_wait();   // chunk cc is converted while chunk cc + 1 is still coming out of TMEM
          if (cc < 2) tmem_ld16(d_cols + 16 * (cc + 1), r[cc + 1]);
          const int col0 = TC_GCOLS * et.grp + 16 * cc;
          uint32_t ph[8], pl[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 bb = *reinterpret_cast<const float2 *>(b + col0 + 2 * j);
            float z0 = __uint_as_float(r[cc][2 * j]), z1 = __uint_as_float(r[cc][2 * j + 1]);
            fma2(z0, z1, 1.f / TC_WSCALE, bb);
            if (last) {   // output layer on the CUDA cores: this thread's 48 columns of relu(z), exact fp32
              ph[j] = pack2h_relu(z0, z1);
              const float h0 = fmaxf(z0, 0.f), h1 = fmaxf(z1, 0.f);
#pragma unroll
              for (int c = 0; c < NO; ++c)
                if (c < n_out) {
                  const float2 w = *reinterpret_cast<const float2 *>(wo32 + c * TC_W + col0 + 2 * j);
                  acc[c].x = fmaf(h0, w.x, acc[c].x);
                  acc[c].y = fmaf(h1, w.y, acc[c].y);
                }
            } else {
              split2h_relu(z0, z1, ph[j], pl[j]);
            }
            if (save_w) {   // masks for the data-gradient chain (same layout and bit trick as the bf16 chain: a non-zero
              // non-negative fp16 half is <= 0x7c00, plus 0x7fff carries into its top bit and never into the other half)
              const uint32_t tt = ph[j] + 0x7fff7fffu;
              constexpr uint32_t one2 = 0x00010001u;
              mask[cc >> 1] |= (tt >> (15 - 8 * (cc & 1) - j)) & (one2 << (8 * (cc & 1) + j));
            }
          }
          if (!last) {   // in place: the chunk's own 16 accumulator columns now hold its K-step of the next A operand
            tmem_st8(d_cols + 16 * cc, ph);
            tmem_st8(d_cols + 16 * cc + 8, pl);
          }
          if (save) {   // the weight-gradient GEMM's operand: the fp16 hi part (same tiled layout as the bf16 chain's copy)
            hl[act_chunk_index(row, col0 / 8)] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            hl[act_chunk_index(row, col0 / 8 + 1)] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
          }
          // (arriving before these stores — as the data-gradient chain does, where it gains 4 % — measured 3 % SLOWER here)
          if (!last) arrive_chunk(cc);
        }
        if (save) {
          uint2 *mb = reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(hidden) + act_mask_base_bytes(NH, m_total));
          mb[act_mask_index(l, act_rows_padded(m_total), row, et.grp)] = make_uint2(mask[0], mask[1]);
        }
        if (last) {   // the four column groups of a row meet in shared memory; column group 0 finishes the row
          // Alternate tiles use alternate buffers — `dsel` flips once per tile at this point (NH is odd) — although the
          // next tile's writes are ordered behind this tile's reads through the chunk barriers and the MMAs already: the
          // second buffer makes that visible to racecheck as well, at no register cost.
          static_assert(NH & 1, "tile parity from dsel");
          float *part = part_buf + dsel * (4 * TC_TM * NO);
#pragma unroll
          for (int c = 0; c < NO; ++c) part[(et.grp * TC_TM + t) * NO + c] = acc[c].x + acc[c].y;
          asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
          if (et.grp == 0 && valid) {
            const float *bo = sbias + NH * TC_W;
#pragma unroll
            for (int c = 0; c < NO; ++c)
              if (c < n_out) {
                float z = bo[c];
#pragma unroll
                for (int g = 0; g < 4; ++g) z += part[(g * TC_TM + t) * NO + c];
                y[row * n_out + c] = act_fwd(z, act);
              }
          }
        }
      }
    }
  } else if (is_issuer && rank == 0 && lane == 0) {
    uint32_t cphase = 0;
    constexpr uint32_t idesc_h = make_idesc_h(TC_W, 2 * TC_TM);
    // one K-step (16 features) of one product: hi.hi + hi.lo + lo.hi
    auto product = [&](uint32_t dst, uint32_t a_hi, uint32_t a_lo, int s, uint32_t w_hi, uint32_t part_bytes, uint32_t rows16,
                       bool first) {
      const uint64_t bh = make_desc(w_hi + 2 * s * rows16, rows16, 128), bl = make_desc(w_hi + part_bytes + 2 * s * rows16, rows16, 128);
      mma_ts2(dst, a_hi, bh, idesc_h, !first);
      mma_ts2(dst, a_hi, bl, idesc_h, 1);
      mma_ts2(dst, a_lo, bh, idesc_h, 1);
    };
    uint32_t dsel = 0;   // accumulator region of the next MMA group
    for (int64_t pt = pair; pt < n_pt; pt += n_pairs) {
      // ---- layer 0: the feature rows the epilogue warps of both CTAs have put into TMEM ----
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        mbar_wait_cluster(bar_chunk + 8 * cc, cphase);
        tc_fence_after();
#pragma unroll
        for (int s = cc; s < K0 / 16; s += 3)
          product(tmem + X2_D + TC_W * dsel, tmem + X2_X0 + 8 * s, tmem + X2_X1 + 8 * s, s, sbase + S::w0, S::w0_part, S::HALF * 16,
                  s == 0);
      }
      mma_commit2(bar);
      cphase ^= 1;
#pragma unroll 1
      for (int l = 0; l + 1 < NH; ++l) {   // hidden layers 1 .. NH - 1 (the output layer runs on the CUDA cores)
        const uint32_t wl = sbase + S::wh + l * (2 * S::wh_part);
        const uint32_t a = tmem + X2_D + TC_W * dsel;   // operand: in place over the previous group's accumulator
        dsel ^= 1;
        const uint32_t dst = tmem + X2_D + TC_W * dsel;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          mbar_wait_cluster(bar_chunk + 8 * cc, cphase);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int s = 3 * g + cc;
            product(dst, a + 16 * s, a + 16 * s + 8, s, wl, S::wh_part, S::HALF * 16, (cc | g) == 0);
          }
        }
        mma_commit2(bar);
        cphase ^= 1;
      }
      dsel ^= 1;   // the next tile's layer 0 accumulates in the region the last hidden layer's operand occupied
    }
  }
  tc_fence_before();
  cluster_sync_all();   // no CTA of the pair may exit (or free its TMEM) while the other can still address it
  if (is_issuer) tmem_dealloc2(tmem, TM_COLS);
}

// ------------------------------------------------------------------------------------------------
// tone-map forward, two tiles in flight
// ------------------------------------------------------------------------------------------------
// The tone-map net (33 -> 192 -> 3) is too small for the generic chain: one 128-row tile is a string of latencies
// (encode -> MMA -> commit -> epilogue -> MMA -> commit -> epilogue, ~2.4 us) during which the 16 epilogue warps work
// for ~0.9 us.  This kernel keeps TWO tiles in flight per CTA, in two slots s = tile & 1:
//   TMEM  D_s [192 s, 192 s + 192): layer-0 accumulator; once every epilogue warp has read it (named barrier) the bf16
//         A operand of the output layer is written over its first 96 columns (the accumulator is dead by then), so two
//         slots fit 512 columns;  O_s [384 + 16 s, +16): output-layer accumulator
//   smem  x_s: the encoded 128 x 48 tile (the CTA computes the positional encoding itself, as the generic XSRC path)
//   mbarriers per slot: x ready (12 encoding warps) -> MMA0 done (commit) -> A ready (16 warps) -> out MMA done (commit)
// The issuer alternates [out MMA of tile i] [MMA0 of tile i + 2]; the epilogue warps alternate [layer-0 epilogue of tile
// i] [output epilogue of tile i - 1], so each side always has the other slot's work to do while a commit is in flight.
template <bool X2>
struct Tm2SmT {
  using F = FwdSm<48, 1>;
  // X2: [W0 hi | W0 lo | Wo hi | Wo lo] fp16 (TcLayout x2 section of the tone-map net) then the f32 biases
  static constexpr int w0 = 0, w0_part = TC_W * 48 * 2;
  static constexpr int wo = X2 ? 2 * w0_part : F::wo, wo_part = TC_NOUT_PAD * TC_W * 2;
  static constexpr int bias = X2 ? wo + 2 * wo_part : F::bias;
  static constexpr int weights_bytes = bias + (TC_W + TC_NOUT_PAD) * 4;
  static constexpr int x0 = (weights_bytes + 127) / 128 * 128;
  static constexpr int x_part = TC_TM * 48 * 2;
  static constexpr int x_bytes = (X2 ? 2 : 1) * x_part;   // per slot: hi tile (+ lo tile)
  static constexpr int bar = x0 + 2 * x_bytes;      // bar_x[2], bar_mma0[2], bar_a[2], bar_out[2] (8 B each), TMEM slot
  static constexpr int bytes = bar + 8 * 8 + 16;
};
using Tm2Sm = Tm2SmT<false>;

template <int NO, bool X2 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_tonemap_fwd2(const uint8_t *__restrict__ image, const float *__restrict__ lin, int64_t m, float *__restrict__ y,
                   int n_out, int act) {
  // X2 (precision 1): `image` is the x2 section of the tone-map image (fp16 hi / lo weights scaled by TC_WSCALE, biases);
  // the encoded tile is kept as an fp16 hi and an fp16 lo tile, every product is three MMAs, and the pair of the hidden
  // activation fills all 192 columns of the consumed accumulator D_s (hi [0,96), lo [96,192))
  extern __shared__ __align__(128) uint8_t smem[];
  using S = Tm2SmT<X2>;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_epi = warp < TC_EPI_WARPS, is_issuer = warp == TC_EPI_WARPS;
  const uint32_t sbase = smem_addr(smem);
  const uint32_t bar_x = sbase + S::bar, bar_mma0 = bar_x + 16, bar_a = bar_x + 32, bar_out = bar_x + 48;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::bar + 64);

  stage_bytes(smem, image, S::weights_bytes);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_x + 8 * s, 12);               // the 12 warps of column groups 0..2 encode one channel each
      mbar_init(bar_mma0 + 8 * s, 1);
      mbar_init(bar_a + 8 * s, TC_EPI_WARPS);
      mbar_init(bar_out + 8 * s, 1);
    }
    fence_mbar_init();
  }
  if (is_issuer) tmem_alloc(smem_addr(tmem_slot), TM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float *sbias = reinterpret_cast<const float *>(smem + S::bias);

  const int64_t n_tiles = (m + TC_TM - 1) / TC_TM;
  const int n_my = blockIdx.x < n_tiles ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;   // tiles of this CTA
  const EpiThread et(warp, lane);
  const int t = et.row_in_tile;
  constexpr uint32_t idesc_h = X2 ? make_idesc_h(TC_W) : make_idesc(TC_W);
  constexpr uint32_t idesc_o = X2 ? make_idesc_h(TC_NOUT_PAD) : make_idesc(TC_NOUT_PAD);
  constexpr int NT = X2 ? 3 : 1;   // MMAs per product: hi.hi, hi.lo, lo.hi

  if (is_issuer) {
    if (lane == 0) {
      auto mma0 = [&](int s) {   // layer 0 of the tile waiting in slot s: A = x_s (shared), B = W0
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int term = 0; term < NT; ++term) {
            const uint32_t xa = sbase + S::x0 + s * S::x_bytes + (term == 2 ? S::x_part : 0) + 2 * k * (TC_TM * 16);
            const uint32_t wb = sbase + S::w0 + (term == 1 ? S::w0_part : 0) + 2 * k * (TC_W * 16);
            mma_ss(tmem + 192 * s, make_desc(xa, TC_TM * 16, 128), make_desc(wb, TC_W * 16, 128), idesc_h, (k | term) != 0);
          }
        mma_commit(bar_mma0 + 8 * s);
      };
      for (int i = 0; i < 2 && i < n_my; ++i) {
        mbar_wait(bar_x + 8 * i, 0);
        tc_fence_after();
        mma0(i);
      }
      for (int i = 0; i < n_my; ++i) {
        const int s = i & 1;
        mbar_wait(bar_a + 8 * s, (i >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < TC_W / 16; ++k)      // output layer: A = activations in TMEM (over D_s), B = Wo
#pragma unroll
          for (int term = 0; term < NT; ++term)
            mma_ts(tmem + 384 + 16 * s, tmem + 192 * s + (term == 2 ? 96 : 0) + 8 * k,
                   make_desc(sbase + S::wo + (term == 1 ? S::wo_part : 0) + 2 * k * (TC_NOUT_PAD * 16), TC_NOUT_PAD * 16, 128),
                   idesc_o, (k | term) != 0);
        mma_commit(bar_out + 8 * s);
        if (i + 2 < n_my) {                       // in order behind the out MMA: D_s / A_s are free when it starts
          mbar_wait(bar_x + 8 * s, ((i + 2) >> 1) & 1);
          tc_fence_after();
          mma0(s);
        }
      }
    }
  } else if (is_epi) {
    auto load_x = [&](int i) {   // column group g < 3 encodes channel g of its row into chunks 2 g, 2 g + 1 of slot i & 1
      if (et.grp < 3) {
        const int64_t row = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * TC_TM + t;
        const bool ok = row < m;
        const float v = ok ? __ldg(lin + 3 * row + et.grp) : 0.f;
        uint4 lo, hi;
        float sn[5], cs[5];
        uint8_t *xs = smem + S::x0 + (i & 1) * S::x_bytes;
        if constexpr (X2) {
          uint4 ll, lh;
          tonemap_pe_channel_x2(v, lo, hi, ll, lh, sn, cs);
          if (!ok) ll = lh = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4 *>(xs + S::x_part + (2 * et.grp) * (TC_TM * 16) + t * 16) = ll;
          *reinterpret_cast<uint4 *>(xs + S::x_part + (2 * et.grp + 1) * (TC_TM * 16) + t * 16) = lh;
        } else {
          tonemap_pe_channel(v, lo, hi, sn, cs);
        }
        if (!ok) lo = hi = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(xs + (2 * et.grp) * (TC_TM * 16) + t * 16) = lo;
        *reinterpret_cast<uint4 *>(xs + (2 * et.grp + 1) * (TC_TM * 16) + t * 16) = hi;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x + 8 * (i & 1));
      }
    };
    auto out_epilogue = [&](int i) {
      const int s = i & 1;
      mbar_wait(bar_out + 8 * s, (i >> 1) & 1);
      tc_fence_after();
      if (et.grp == 0) {
        uint32_t r[16];
        tmem_ld16(tmem + et.lane_base + 384 + 16 * s, r);
        tmem_ld_wait();
        const int64_t row = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * TC_TM + t;
        if (row < m) {
          const float *bo = sbias + TC_W;
#pragma unroll
          for (int c = 0; c < NO; ++c)
            if (c < n_out)
              y[row * n_out + c] = act_fwd(X2 ? __fmaf_rn(__uint_as_float(r[c]), 1.f / TC_WSCALE, bo[c]) : __uint_as_float(r[c]) + bo[c], act);
        }
      }
      tc_fence_before();
    };
    for (int i = 0; i < 2 && i < n_my; ++i) load_x(i);
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      mbar_wait(bar_mma0 + 8 * s, (i >> 1) & 1);
      tc_fence_after();
      if (i + 2 < n_my) load_x(i + 2);            // x_s is free: encode the tile after next
      uint32_t r[3][16];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) tmem_ld16(tmem + et.lane_base + 192 * s + TC_GCOLS * et.grp + 16 * cc, r[cc]);
      tmem_ld_wait();
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");   // every warp has read D_s: A may overwrite it
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int col0 = TC_GCOLS * et.grp + 16 * cc;
        uint32_t p[8];
        [[maybe_unused]] uint32_t pl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 bb = *reinterpret_cast<const float2 *>(sbias + col0 + 2 * j);
          float z0 = __uint_as_float(r[cc][2 * j]), z1 = __uint_as_float(r[cc][2 * j + 1]);
          if constexpr (X2) {
            fma2(z0, z1, 1.f / TC_WSCALE, bb);
            split2h_relu(z0, z1, p[j], pl[j]);
          } else {
            add2(z0, z1, bb);
            p[j] = pack2_relu(z0, z1);
          }
        }
        tmem_st8(tmem + et.lane_base + 192 * s + col0 / 2, p);
        if constexpr (X2) tmem_st8(tmem + et.lane_base + 192 * s + 96 + col0 / 2, pl);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a + 8 * s);
      if (i >= 1) out_epilogue(i - 1);
    }
    if (n_my >= 1) out_epilogue(n_my - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (is_issuer) tmem_dealloc(tmem, TM_COLS);
}

// ------------------------------------------------------------------------------------------------
// data-gradient chain
// ------------------------------------------------------------------------------------------------
template <int K0, int NH, int DXN>
struct BwdSm {
  static constexpr int wo = 0;                                        // WoT [W x 16]
  static constexpr int wh = wo + TC_W * TC_NOUT_PAD * 2;              // W_lT, l = 1..NH-1
  static constexpr int w0 = wh + (NH - 1) * TC_W * TC_W * 2;          // W0T [DXN x W]
  static constexpr int weights_bytes = w0 + DXN * TC_W * 2;           // == TcLayout::b_bytes()
  static constexpr int dz = (weights_bytes + 127) / 128 * 128;        // dZ_out tile [128 x 16] bf16
  static constexpr int bar = dz + TC_TM * TC_NOUT_PAD * 2;   // bar_mma, bar_chunk[3] (8 B each), TMEM slot, (OVL) bar_x, bar_first
  static constexpr int bytes = bar + 64;
};

// OVL = tile overlap (opt-in, see k_mlp_fwd_tc): the next tile's dZ_out tile is made under this tile's chain and its
// first MMA (dZ_out W_o -> D0) is issued behind this tile's last MMA, so that it runs under the d_x epilogue; the first
// MMA's commit has its own mbarrier (bar_first).
// H16 = fp16 chain (precision 1): the cotangents travel through the chain as fp16 (11 significant bits against bf16's
// 8) and the transposed weights are fp16.  fp16's narrow exponent is dealt with PER ROW: the chain is linear in a row's
// output cotangent, so row r is multiplied by a power of two s_r that puts max|dZ_out[r]| in [16, 32) (what the chain can
// add on top stays far below 65504, what it loses at the bottom is an ABSOLUTE error of s_r^-1 2^-25, irrelevant in sums
// over rows); d_x leaves the kernel multiplied by 1 / s_r, exactly.  The copies for the weight-gradient GEMM (dZ_l,
// dZ_out) are fp16 too, but a sum over rows needs ONE scale for the whole launch: G = the power of two that puts
// max |d_y| (`absmax`, made by k_absmax just before) in [64, 128) — |act'| <= 1, so no |dZ_out| exceeds it; rows more
// than 2^20 below the largest lose precision gradually, and weigh that little in the sums.  Same MMAs as the bf16
// chain; the gradients it produces are ~8x closer to the reference's (5e-4 instead of 4e-3 relative L2).
template <int K0, int NH, int DXN, int NO, bool ACC, bool OVL = false, bool H16 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_mlp_dgrad_tc(const uint8_t *__restrict__ image_bwd, const float *__restrict__ y, const float *__restrict__ d_y,
                   int64_t row_begin, int64_t row_end, int64_t m_total, const __nv_bfloat16 *__restrict__ hidden,
                   __nv_bfloat16 *__restrict__ d_z, float *__restrict__ d_z_out, float *__restrict__ d_x, int dx_cols,
                   int accumulate, int n_out, int act, const uint32_t *__restrict__ absmax) {
  extern __shared__ __align__(128) uint8_t smem[];
  using S = BwdSm<K0, NH, DXN>;
  [[maybe_unused]] float g_inv = 1.f;
  [[maybe_unused]] const float g_scale = H16 ? act_dz_scale(__ldg(absmax), g_inv) : 1.f;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_epi = warp < TC_EPI_WARPS, is_issuer = warp == TC_EPI_WARPS;
  const uint32_t sbase = smem_addr(smem);
  const uint32_t bar = sbase + S::bar, bar_chunk = bar + 8;
  [[maybe_unused]] const uint32_t bar_x = bar + 40, bar_first = bar + 48;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::bar + 32);
  uint32_t cphase = 0;   // parity of the chunk barriers (issuer)
  [[maybe_unused]] uint32_t xphase = 0;   // OVL: parity of bar_x (issuer)
  [[maybe_unused]] uint32_t fphase = 0;   // OVL: parity of bar_first (epilogue warps)
  // OVL: the next tile's first MMA writes D0 under the d_x epilogue: D0 must be the accumulator of the last chain
  // step (read before the arrival on bar_x), not of d_x
  static_assert(!OVL || (NH & 1), "tile overlap: odd number of hidden layers (the d_x accumulator lives in D1)");
  static_assert(!(OVL && H16), "the fp16 chain is instantiated without the tile overlap");
  // H16: 1 / s_r of the tile's rows.  Written by column group 0 at the top of a tile, read by every epilogue thread right
  // after the tile's first __syncthreads; the next tile's write is ordered behind those reads through the chunk
  // barriers -> tcgen05.mma -> commit -> the writer's own d_x epilogue wait (racecheck, which does not follow that
  // chain, reports it as a hazard: profiles/r02_sanitizer).  A second buffer costs the kernel a register it does not have.
  [[maybe_unused]] float *s_inv = reinterpret_cast<float *>(smem + S::bytes);
  constexpr uint32_t idesc_w = H16 ? make_idesc_h(TC_W) : make_idesc(TC_W);
  constexpr uint32_t idesc_x = H16 ? make_idesc_h(DXN) : make_idesc(DXN);

  stage_bytes(smem, image_bwd, S::weights_bytes);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) mbar_init(bar_chunk + 8 * c, TC_EPI_WARPS);
    if constexpr (OVL) {
      mbar_init(bar_x, TC_EPI_WARPS);
      mbar_init(bar_first, 1);
    }
    fence_mbar_init();
  }
  if (is_issuer) tmem_alloc(smem_addr(tmem_slot), TM_COLS);
  // chunk 1 (columns 8..15) of the dZ_out tile stays zero
  if (threadIdx.x < TC_TM) *reinterpret_cast<uint4 *>(smem + S::dz + TC_TM * 16 + threadIdx.x * 16) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t n_tiles = (row_end - row_begin + TC_TM - 1) / TC_TM;
  const EpiThread et(warp, lane);
  const int t = et.row_in_tile;
  const int64_t layer_stride = act_rows_padded(m_total) * TC_W;
  const uint2 *mask_base = reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(hidden) +
                                                             act_mask_base_bytes(NH, m_total));
  uint32_t phase = 0;

  // Per-tile global inputs (output cotangent, ReLU masks of every layer) are requested one tile ahead so their
  // latency hides behind the current tile's chain.
  float pf_y[NO], pf_dy[NO];
  uint2 pf_mask[NH];
  auto prefetch = [&](int64_t tile) {
    const int64_t row = row_begin + tile * TC_TM + t;
    const bool ok = is_epi && tile < n_tiles && row < row_end;
#pragma unroll
    for (int c = 0; c < NO; ++c) {
      const bool okc = ok && et.grp == 0 && c < n_out;
      pf_y[c] = okc ? __ldg(y + row * n_out + c) : 0.f;
      pf_dy[c] = okc ? __ldg(d_y + row * n_out + c) : 0.f;
    }
#pragma unroll
    for (int l = 0; l < NH; ++l)
      pf_mask[l] = ok ? __ldg(mask_base + act_mask_index(l, act_rows_padded(m_total), row, et.grp)) : make_uint2(0, 0);
  };
  // OVL: dZ_out = d_y * act'(y) of tile `tl` from the prefetched (y, d_y): A tile of the chain's first MMA (column group 0)
  [[maybe_unused]] auto make_dz = [&](int64_t tl) {
    if constexpr (OVL) {   // (an empty body otherwise: the non-overlapped instantiation captures nothing)
      const int64_t row = row_begin + tl * TC_TM + t;
      const bool valid = row < row_end;
      float dz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll
        for (int c = 0; c < NO; ++c)
          if (c < n_out) {
            const float yy = pf_y[c];
            dz[c] = pf_dy[c] * (act == 1 ? (1.f - expf(-yy)) : (act == 2 ? yy * (1.f - yy) : 1.f));
          }
        if (d_z_out) {
          *reinterpret_cast<float4 *>(d_z_out + row * 8) = make_float4(dz[0], dz[1], dz[2], dz[3]);
          *reinterpret_cast<float4 *>(d_z_out + row * 8 + 4) = make_float4(dz[4], dz[5], dz[6], dz[7]);
        }
      }
      const uint4 dz16 = make_uint4(pack2(dz[0], dz[1]), pack2(dz[2], dz[3]), pack2(dz[4], dz[5]), pack2(dz[6], dz[7]));
      if (valid) {  // bf16 copy for the output layer's weight-gradient GEMM: tiled, 2 chunks per row, after the dZ_l
        uint4 *zo = reinterpret_cast<uint4 *>(d_z + (int64_t)NH * layer_stride);
        zo[tiled_chunk_index(row, 0, 2)] = dz16;
        zo[tiled_chunk_index(row, 1, 2)] = make_uint4(0u, 0u, 0u, 0u);
      }
      *reinterpret_cast<uint4 *>(smem + S::dz + t * 16) = dz16;
      fence_proxy_async();
    }
  };
  prefetch(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = row_begin + tile * TC_TM + t;
    const bool valid = is_epi && row < row_end;
    uint2 cur_mask[NH];
#pragma unroll
    for (int l = 0; l < NH; ++l) cur_mask[l] = pf_mask[l];
    // accumulate mode: the previous d_x values of this thread's 16 columns are requested now, a whole layer chain
    // ahead of the epilogue that adds to them (a read-modify-write at the end would stall the only tile in flight)
    float4 old_dx[ACC ? 4 : 1];
    if (ACC && valid && d_x && et.grp < DXN / 16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = et.grp * 16 + 4 * q;
        old_dx[q] = col < dx_cols ? *reinterpret_cast<const float4 *>(d_x + row * dx_cols + col)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    [[maybe_unused]] const bool first = tile == (int64_t)blockIdx.x, more = tile + gridDim.x < n_tiles;
    [[maybe_unused]] auto issue_first = [&]() {   // OVL
      mma_ss(tmem + tm_d(0), make_desc(sbase + S::dz, TC_TM * 16, 128), make_desc(sbase + S::wo, TC_W * 16, 128),
             make_idesc(TC_W), 0);
      mma_commit(bar_first);
    };
    if constexpr (OVL) {
      // ---- dZ_out tile of the first MMA: made here for the CTA's first tile, under the previous tile's chain otherwise ----
      if (first && is_epi && et.grp == 0) make_dz(tile);
      prefetch(tile + gridDim.x);
      if (first) {
        tc_fence_before();
        __syncthreads();
      }
    } else {
    // ---- dZ_out = d_y * act'(y): A tile of the first MMA (column group 0 threads) ----
    if (is_epi && et.grp == 0) {
      float dz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
#pragma unroll
        for (int c = 0; c < NO; ++c)
          if (c < n_out) {
            const float yy = pf_y[c];
            dz[c] = pf_dy[c] * (act == 1 ? (1.f - expf(-yy)) : (act == 2 ? yy * (1.f - yy) : 1.f));
          }
        if (d_z_out) {
          *reinterpret_cast<float4 *>(d_z_out + row * 8) = make_float4(dz[0], dz[1], dz[2], dz[3]);
          *reinterpret_cast<float4 *>(d_z_out + row * 8 + 4) = make_float4(dz[4], dz[5], dz[6], dz[7]);
        }
      }
      // copy for the output layer's weight-gradient GEMM (bf16, or fp16 times G): tiled, 2 chunks per row, after the dZ_l
      const uint4 dz16 = H16 ? make_uint4(pack2h_sat(dz[0] * g_scale, dz[1] * g_scale), pack2h_sat(dz[2] * g_scale, dz[3] * g_scale),
                                          pack2h_sat(dz[4] * g_scale, dz[5] * g_scale), pack2h_sat(dz[6] * g_scale, dz[7] * g_scale))
                             : make_uint4(pack2(dz[0], dz[1]), pack2(dz[2], dz[3]), pack2(dz[4], dz[5]), pack2(dz[6], dz[7]));
      if (valid) {
        uint4 *zo = reinterpret_cast<uint4 *>(d_z + (int64_t)NH * layer_stride);
        zo[tiled_chunk_index(row, 0, 2)] = dz16;
        zo[tiled_chunk_index(row, 1, 2)] = make_uint4(0u, 0u, 0u, 0u);
      }
      if constexpr (H16) {   // the chain's own copy: fp16, times the row's power-of-two scale
        float mxa = 0.f;
#pragma unroll
        for (int c = 0; c < NO; ++c) mxa = fmaxf(mxa, fabsf(dz[c]));
        const int e = (__float_as_int(mxa) >> 23) & 0xff;          // biased exponent of the row's largest |dZ_out|
        const bool scaled = e >= 8 && e <= 250;                    // zero / denormal / absurd rows travel unscaled
        const float sr = scaled ? __int_as_float((258 - e) << 23) : 1.f;   // 2^(4 - (e - 127))
        s_inv[t] = scaled ? __int_as_float((e - 4) << 23) : 1.f;
        *reinterpret_cast<uint4 *>(smem + S::dz + t * 16) =
            make_uint4(pack2h(dz[0] * sr, dz[1] * sr), pack2h(dz[2] * sr, dz[3] * sr), pack2h(dz[4] * sr, dz[5] * sr),
                       pack2h(dz[6] * sr, dz[7] * sr));
      } else {
        *reinterpret_cast<uint4 *>(smem + S::dz + t * 16) = dz16;
      }
      fence_proxy_async();
    }
    prefetch(tile + gridDim.x);
    tc_fence_before();
    __syncthreads();
    }
    [[maybe_unused]] const float inv_s = (H16 && is_epi) ? s_inv[t] : 1.f;
    [[maybe_unused]] const float z_s = inv_s * g_scale;   // chain value -> stored cotangent (both powers of two: exact)
    if (is_issuer && lane == 0) {
      if constexpr (!OVL) {
        tc_fence_after();
        mma_ss(tmem + tm_d(0), make_desc(sbase + S::dz, TC_TM * 16, 128), make_desc(sbase + S::wo, TC_W * 16, 128),
               idesc_w, 0);
        mma_commit(bar);
      } else if (first) {
        tc_fence_after();
        issue_first();
      }
      // chain step i handles layer l = NH - 1 - i: its accumulator is D region i & 1, the MMA it feeds (W_l^T, or
      // W_0^T -> d_x for l == 0) accumulates in the other region, chunk by chunk (see chunk_ready)
#pragma unroll 1
      for (int i = 0; i < NH; ++i) {
        const int l = NH - 1 - i;
        const uint32_t dst = tmem + tm_d(i + 1);
        const uint32_t wl = l > 0 ? sbase + S::wh + (l - 1) * (TC_W * TC_W * 2) : sbase + S::w0;
        const uint32_t rows16 = (l > 0 ? TC_W : DXN) * 16;
        const uint32_t idesc = l > 0 ? idesc_w : idesc_x;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          mbar_wait(bar_chunk + 8 * cc, cphase);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int s = 3 * g + cc;
            mma_ts(dst, tmem + TM_A + 8 * s, make_desc(wl + 2 * s * rows16, rows16, 128), idesc, (cc | g) != 0);
          }
        }
        mma_commit(bar);
        cphase ^= 1;
      }
      if constexpr (OVL) {
        if (more) {   // next tile's dZ_out is in shared memory and every warp is done with D0
          mbar_wait(bar_x, xphase);
          xphase ^= 1;
          tc_fence_after();
          issue_first();
        }
      }
    }
    if (is_epi) {
#pragma unroll 1
      for (int i = 0; i < NH; ++i) {
        const int l = NH - 1 - i;
        // ReLU mask bits of this thread's 48 columns of H_l (written by the forward chain, prefetched above)
        uint4 *zl = reinterpret_cast<uint4 *>(d_z + (int64_t)l * layer_stride);
        uint2 mk2 = cur_mask[0];
#pragma unroll
        for (int q = 1; q < NH; ++q)
          if (q == l) mk2 = cur_mask[q];
        const uint32_t mask[2] = {mk2.x, mk2.y};
        if (OVL && i == 0) {
          mbar_wait(bar_first, fphase);
          fphase ^= 1;
        } else {
          mbar_wait(bar, phase);
          phase ^= 1;
        }
        tc_fence_after();
        // dZ_l = dH_l * [H_l > 0] -> bf16 -> TMEM A operand + global copy (tiled) for the weight-gradient GEMM
        uint32_t r[3][16];   // (chunk cc + 1 comes out of TMEM while chunk cc is converted: see k_mlp_fwd_tc)
        const uint32_t d_cols = tmem + et.lane_base + tm_d(i) + TC_GCOLS * et.grp;
        tmem_ld16(d_cols, r[0]);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          tmem_ld_wait();
          if (cc < 2) tmem_ld16(d_cols + 16 * (cc + 1), r[cc + 1]);
          const int col0 = TC_GCOLS * et.grp + 16 * cc;
          const uint32_t mk = mask[cc >> 1] >> (8 * (cc & 1));   // pair j: bits (j, 16 + j)
          uint32_t p[8];
          [[maybe_unused]] uint32_t pa[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {   // 0xffff per live half: AND on the packed pair instead of two selects
            const float v0 = __uint_as_float(r[cc][2 * j]), v1 = __uint_as_float(r[cc][2 * j + 1]);
            const uint32_t pm = pair_mask(mk, j);
            if constexpr (H16) {
              pa[j] = pack2h(v0, v1) & pm;                     // next MMA's operand: fp16, still carrying s_r
              p[j] = pack2h_sat(v0 * z_s, v1 * z_s) & pm;      // what the weight-gradient GEMM reads: fp16, times G
            } else {
              p[j] = pack2(v0, v1) & pm;
            }
          }
          if constexpr (H16)
            tmem_st8(tmem + et.lane_base + TM_A + col0 / 2, pa);
          else
            tmem_st8(tmem + et.lane_base + TM_A + col0 / 2, p);
          chunk_ready(bar_chunk + 8 * cc, lane);   // before the global copy (stores queue behind every other warp's): -4 %
          if (valid) {
            zl[act_chunk_index(row, col0 / 8)] = make_uint4(p[0], p[1], p[2], p[3]);
            zl[act_chunk_index(row, col0 / 8 + 1)] = make_uint4(p[4], p[5], p[6], p[7]);
          }
        }
      }
      if constexpr (OVL) {
        if (more) {   // chain epilogues done (D0 read): hand the next tile's dZ_out (prefetched y, d_y) to the issuer
          if (et.grp == 0) make_dz(tile + gridDim.x);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x);
        }
      }
      // ---- d_x epilogue: one 16-column group per epilogue column group ----
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < DXN / 16; ++cc) {
        if (cc != et.grp) continue;  // warp-uniform
        uint32_t r[16];
        tmem_ld16(tmem + et.lane_base + tm_d(NH) + cc * 16, r);
        tmem_ld_wait();
        if (valid && d_x) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = cc * 16 + 4 * q;
            if (col < dx_cols) {  // dx_cols is a multiple of 4 (56 / 40)
              float4 *p4 = reinterpret_cast<float4 *>(d_x + row * dx_cols + col);
              float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                     __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
              if constexpr (H16) v.x *= inv_s, v.y *= inv_s, v.z *= inv_s, v.w *= inv_s;
              if (ACC) {
                const float4 o = old_dx[ACC ? q : 0];
                v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
              }
              *p4 = v;
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (is_issuer) tmem_dealloc(tmem, TM_COLS);
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW_l[o][i] += sum_m dZ_l[m][o] * In_l[m][i],  db_l[o] += sum_m dZ_l[m][o]
//
// A GEMM whose K dimension is the sample index.  Both operands are read straight out of the tiled activation
// buffers: a [128 rows x 8c features] tile is, byte for byte, a no-swizzle MN-major UMMA operand (8 features
// contiguous, the next 8 rows 128 B further = LBO, the next 8 features 2048 B further = SBO), so one 1-D bulk copy
// (cp.async.bulk, completion on an mbarrier) per operand per tile feeds the tensor core with no re-layout.
// M = 192 output features does not fit one UMMA (M <= 128): two accumulators of M = 128 over features [0,128) and
// [64,192) (the overlap is computed twice and discarded).  A constant "ones" chunk appended to the B tile makes the
// bias gradient column KIN of the same accumulator.  Split-K across persistent CTAs; fp32 partials leave through
// vector RED.
// ------------------------------------------------------------------------------------------------
ESR_D void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
ESR_D void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// instruction descriptor: D f32, A/B bf16, both MN-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) { return make_idesc(n) | (1u << 15) | (1u << 16); }
__host__ __device__ constexpr uint32_t make_idesc_h_mn(int n) { return make_idesc_h(n) | (1u << 15) | (1u << 16); }

// ACH = feature chunks (8 features each) of the A operand per row: 24 for a hidden layer's dZ_l (two accumulators,
// see above), 2 for the output layer's dZ_out (16 columns, 3 real: one accumulator over a 128-feature A tile whose
// chunks 2..15 are constant zero).
template <int KIN, int ACH>
struct WgSm {
  static constexpr int a_chunks = ACH == 24 ? 24 : 16;
  static constexpr int a_bytes = a_chunks * TC_TM * 16;
  static constexpr int a_load = ACH * TC_TM * 16;
  static constexpr int b_chunks = KIN / 8 + 2;                    // In tile + 2 constant chunks ("ones" column)
  static constexpr int b_bytes = b_chunks * TC_TM * 16;
  static constexpr int b_load = (KIN / 8) * TC_TM * 16;
  static constexpr int stage = a_bytes + b_bytes;
  static constexpr int bar = 2 * stage;                           // full[2], empty[2] (8 B each), tmem slot
  static constexpr int bytes = bar + 48;
  static constexpr int NB = KIN + 16;                             // UMMA N
};

// H16 (precision 1): both operands are fp16 — the layer input as the forward chain saved it, the cotangent times the
// launch's power-of-two scale G (k_mlp_dgrad_tc) — and the sums are divided by G on their way out.
template <int KIN, int ACH, bool H16 = false>
__global__ void __launch_bounds__(160, 1)
    k_mlp_wgrad_tc(const __nv_bfloat16 *__restrict__ dz, const __nv_bfloat16 *__restrict__ in, int64_t row_begin,
                   int64_t row_end, int out_rows, float *__restrict__ gW /* [out_rows][KIN] */,
                   float *__restrict__ gb /* [out_rows] */, const uint32_t *__restrict__ absmax) {
  constexpr uint32_t ONE = H16 ? 0x00003c00u : 0x00003f80u;   // 1.0 in element 0 of a chunk (fp16 / bf16)
  constexpr uint32_t idesc = H16 ? make_idesc_h_mn(WgSm<KIN, ACH>::NB) : make_idesc_mn(WgSm<KIN, ACH>::NB);
  extern __shared__ __align__(128) uint8_t smem[];
  using S = WgSm<KIN, ACH>;
  const int64_t T0 = row_begin >> 7, T1 = (row_end + 127) >> 7;
  const int64_t per = (T1 - T0 + gridDim.x - 1) / gridDim.x;
  const int64_t ta = T0 + (int64_t)blockIdx.x * per, tb = min(T1, ta + per);
  const int64_t n = tb - ta;
  if (n <= 0) return;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_addr(smem);
  const uint32_t bar_full = sbase + S::bar, bar_empty = sbase + S::bar + 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::bar + 32);
  const bool driver = threadIdx.x == 128;  // producer + MMA issuer

  // constant chunks of both stages: column KIN of the B tile is 1.0 for every row, columns KIN+1.. are zero;
  // A chunks that are never loaded are zero
  for (int i = threadIdx.x; i < 2 * 2 * TC_TM; i += blockDim.x) {
    const int st = i / (2 * TC_TM), r = i % (2 * TC_TM);  // r < 128: chunk KIN/8, else chunk KIN/8 + 1
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < TC_TM) v.x = ONE;
    *reinterpret_cast<uint4 *>(smem + st * S::stage + S::a_bytes + (KIN / 8) * (TC_TM * 16) + r * 16) = v;
  }
  if constexpr (S::a_chunks > ACH) {
    for (int i = threadIdx.x; i < 2 * (S::a_chunks - ACH) * TC_TM; i += blockDim.x) {
      const int st = i / ((S::a_chunks - ACH) * TC_TM), r = i % ((S::a_chunks - ACH) * TC_TM);
      *reinterpret_cast<uint4 *>(smem + st * S::stage + S::a_load + r * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1), mbar_init(bar_full + 8, 1), mbar_init(bar_empty, 1), mbar_init(bar_empty + 8, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(smem_addr(tmem_slot), TM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint4 *dz4 = reinterpret_cast<const uint4 *>(dz), *in4 = reinterpret_cast<const uint4 *>(in);

  auto load_tile = [&](int64_t tile, int st) {  // driver thread
    const uint32_t full = bar_full + 8 * st;
    mbar_expect_tx(full, S::a_load + S::b_load);
    bulk_g2s(sbase + st * S::stage, dz4 + tile * ACH * TC_TM, S::a_load, full);
    bulk_g2s(sbase + st * S::stage + S::a_bytes, in4 + tile * (KIN / 8) * TC_TM, S::b_load, full);
  };

  if (driver) load_tile(ta, 0);
  for (int64_t it = 0; it < n; ++it) {
    const int st = (int)(it & 1);
    const uint32_t k_par = (uint32_t)((it >> 1) & 1);
    const int64_t tile = ta + it;
    if (driver && it + 1 < n) {
      if (it >= 1) mbar_wait(bar_empty + 8 * (st ^ 1), (uint32_t)(((it - 1) >> 1) & 1));  // MMAs of tile it-1 retired
      load_tile(tile + 1, st ^ 1);
    }
    const int64_t r0 = tile * TC_TM;
    const bool boundary = r0 < row_begin || r0 + TC_TM > row_end;
    if (boundary) {
      // boundary tile: rows outside [row_begin, row_end) hold unrelated data -> zero them in both operands
      mbar_wait(bar_full + 8 * st, k_par);
      if (threadIdx.x < TC_TM) {
        const int64_t row = r0 + threadIdx.x;
        if (row < row_begin || row >= row_end) {
          uint8_t *a = smem + st * S::stage + threadIdx.x * 16;
#pragma unroll 4
          for (int c = 0; c < ACH; ++c) *reinterpret_cast<uint4 *>(a + c * (TC_TM * 16)) = make_uint4(0, 0, 0, 0);
#pragma unroll 4
          for (int c = 0; c < KIN / 8 + 1; ++c)   // + the "ones" chunk: the row must not count in the bias either
            *reinterpret_cast<uint4 *>(a + S::a_bytes + c * (TC_TM * 16)) = make_uint4(0, 0, 0, 0);
        }
      }
      fence_proxy_async();
      __syncthreads();
    }
    if (driver) {
      mbar_wait(bar_full + 8 * st, k_par);
      tc_fence_after();
      const uint32_t a0 = sbase + st * S::stage, b0 = a0 + S::a_bytes;
#pragma unroll
      for (int s = 0; s < TC_TM / 16; ++s) {
        const uint64_t bd = make_desc(b0 + s * 256, 128, TC_TM * 16);
        mma_ss(tmem + 0, make_desc(a0 + s * 256, 128, TC_TM * 16), bd, idesc, (it | s) != 0);
        if (ACH == 24)
          mma_ss(tmem + 256, make_desc(a0 + 8 * (TC_TM * 16) + s * 256, 128, TC_TM * 16), bd, idesc, (it | s) != 0);
      }
      mma_commit(bar_empty + 8 * st);
    }
    __syncthreads();
    if (boundary) {
      // restore the "ones" chunk rows zeroed above once the MMAs that read them have retired (next use of the stage)
      mbar_wait(bar_empty + 8 * st, k_par);
      if (threadIdx.x < TC_TM)
        *reinterpret_cast<uint4 *>(smem + st * S::stage + S::a_bytes + (KIN / 8) * (TC_TM * 16) + threadIdx.x * 16) =
            make_uint4(ONE, 0, 0, 0);
      fence_proxy_async();
      __syncthreads();
    }
  }
  // ---- epilogue: accumulators -> global gradient (RED) ----
  float g_inv = 1.f;
  if constexpr (H16) act_dz_scale(__ldg(absmax), g_inv);
  mbar_wait(bar_empty + 8 * (int)((n - 1) & 1), (uint32_t)(((n - 1) >> 1) & 1));
  tc_fence_after();
  if (warp < 4) {
    const uint32_t lane_base = (32u * warp) << 16;
    const int ml = 32 * warp + lane;  // accumulator row
#pragma unroll 1
    for (int acc = 0; acc < (ACH == 24 ? 2 : 1); ++acc) {
      const int o = acc == 0 ? ml : 64 + ml;
      const bool use = (acc == 0 || ml >= 64) && o < out_rows;  // accumulator 1: only features 128..191 are new
#pragma unroll 1
      for (int cc = 0; cc < S::NB / 16; ++cc) {
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + acc * 256 + cc * 16, r);
        tmem_ld_wait();
        if (use) {
          if (cc < KIN / 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              red_add4(gW + (int64_t)o * KIN + cc * 16 + 4 * q, __uint_as_float(r[4 * q]) * g_inv,
                       __uint_as_float(r[4 * q + 1]) * g_inv, __uint_as_float(r[4 * q + 2]) * g_inv,
                       __uint_as_float(r[4 * q + 3]) * g_inv);
          } else {
            red_add(gb + o, __uint_as_float(r[0]) * g_inv);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, TM_COLS);
}

// max |v| over a float range as f32 bits (non-negative floats order like their bit patterns; NaNs are skipped by fmaxf)
__global__ void k_absmax(const float *__restrict__ v, int64_t n, uint32_t *__restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(v + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

template <int KIN, int ACH, bool H16 = false>
static int launch_wgrad(const __nv_bfloat16 *dz, const __nv_bfloat16 *in, int64_t rb, int64_t re, int out_rows, float *gW,
                        float *gb, const uint32_t *absmax, cudaStream_t st) {
  auto kern = k_mlp_wgrad_tc<KIN, ACH, H16>;
  constexpr int bytes = WgSm<KIN, ACH>::bytes;
  ESR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  const int64_t tiles = ((re + 127) >> 7) - (rb >> 7);
  const unsigned grid = (unsigned)max((int64_t)1, min((int64_t)num_sms(), tiles));
  ESR_STAGE(ACH == 24 ? "k_mlp_wgrad_tc" : "k_mlp_wgrad_tc_out", st);
  kern<<<grid, 160, bytes, st>>>(dz, in, rb, re, out_rows, gW, gb, absmax);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused tone-map backward (voxurff.py:783-788 + pbr/module.py:24-39, 33 -> 192 -> 3): ONE kernel reads
// (lin, rgb, d_rgb[, d_lin_direct]) = 36-48 B per row and writes d_lin (12 B per row) plus the weight gradients.
// Everything the unfused path moved through HBM (encoded rows 96 B, hidden activations 384 B + masks, their
// cotangents 384 B, the encoded-row cotangent 192 B: ~2.4 KB per row over four kernels) stays on the SM:
//   T0  the tile's encoding X is recomputed from lin into shared memory (K-major for the forward MMA; the same bytes
//       are an MN-major operand for the weight-gradient MMA), dZ_out = d_rgb * rgb (1 - rgb) likewise
//   E1  H = relu(X W0^T + b0) recomputed on the tensor core -> bf16 tile Hs in shared memory + ReLU masks in registers
//   E2  dZ0 = (dZ_out Wo) * mask -> TMEM (A operand of the d_x MMA) + bf16 tile Zs in shared memory
//   E3  d_x = dZ0 W0: thread (row, channel) holds the 16 encoded columns of its channel and the sines / cosines it
//       computed in T0 -> d_lin without any cross-thread traffic
//   weight gradients accumulate in TMEM across all tiles of the CTA: dW0|db0 += Zs^T [X | 1] (two M = 128 halves,
//   N = 64), dWo^T += Hs^T dZ_out (N = 16); db_out is a register sum.  They leave through REDs once per CTA.
// TMEM: D [0,192)  A [192,288)  S [288,336)  dW0 halves [336,400) [400,464)  dWo^T halves [464,480) [480,496).
// ------------------------------------------------------------------------------------------------
template <bool X2>
struct TmBwdSmT {
  static constexpr int w0 = 0;                         // W0   K-major [6][192][8]; X2: fp16 hi then fp16 lo (scaled)
  static constexpr int w0_part = TC_W * 48 * 2;
  static constexpr int wot = w0 + (X2 ? 2 : 1) * w0_part;   // Wo^T K-major [2][192][8]
  static constexpr int w0t = wot + TC_W * 16 * 2;      // W0^T K-major [24][48][8]
  static constexpr int bias = w0t + 48 * TC_W * 2;     // b0 f32 [192]
  static constexpr int x = bias + TC_W * 4;            // X tile [8][128][8] bf16: 6 feature chunks, ones chunk, zero chunk
  static constexpr int hs = x + 8 * TC_TM * 16;        // H tile  [24][128][8]
  static constexpr int zs = hs + 24 * TC_TM * 16;      // dZ0 tile [24][128][8]
  static constexpr int dzo = zs + 24 * TC_TM * 16;     // dZ_out tile [2][128][8] (chunk 1 zero); X2: times the ROW's scale
  static constexpr int xl = dzo + 2 * TC_TM * 16;      // X2: fp16 lo tile of the encoding [6][128][8] (the hi tile is `x`)
  static constexpr int dzw = xl + (X2 ? 6 * TC_TM * 16 : 0);   // X2: dZ_out tile times the CTA's scale (weight gradient)
  static constexpr int sinv = dzw + (X2 ? 2 * TC_TM * 16 : 0); // X2: 1 / (row scale) f32 [128], then the CTA's max |d_y| (u32)
  static constexpr int bar = sinv + (X2 ? TC_TM * 4 + 16 : 0);   // bar_mma, bar_w, TMEM slot
  static constexpr int bytes = bar + 32;
};
using TmBwdSm = TmBwdSmT<false>;
constexpr uint32_t TMB_D = 0, TMB_A = 192, TMB_S = 288, TMB_G0 = 336, TMB_GO = 464;

// X2 (precision 1): the recomputation of H uses the forward's arithmetic — fp16 hi / lo tiles of the encoding and of W0
// (x2 section at byte x2_off of the image), three MMAs per K-step — so that the masks are the forward's, and everything
// downstream runs on fp16 operands (11 significant bits against bf16's 8): the H / X tiles, the transposed weights, and
// the cotangents, whose narrow exponent is handled as in k_mlp_dgrad_tc — per ROW in the data-gradient chain (power of
// two s_r putting max |dZ_out[r]| in [16, 32), divided out of d_lin), per CTA in the tiles of the weight-gradient MMAs
// (their sums run over all rows of the CTA's tiles: G = the power of two putting the CTA's max |d_y|, found by a first
// pass over its rows' d_y, in [64, 128); divided out of the accumulators before the REDs).
template <bool X2>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_tonemap_bwd_fused(const uint8_t *__restrict__ image, int64_t bwd_off, int64_t bias_off, int64_t x2_off,
                        const float *__restrict__ lin,
                        const float *__restrict__ y, const float *__restrict__ d_y, const float *__restrict__ d_direct,
                        int64_t m, float *__restrict__ d_lin, float *__restrict__ gW0, float *__restrict__ gb0,
                        float *__restrict__ gWo, float *__restrict__ gbo, int n_out, int act) {
  extern __shared__ __align__(128) uint8_t smem[];
  using S = TmBwdSmT<X2>;
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_epi = warp < TC_EPI_WARPS, is_issuer = warp == TC_EPI_WARPS;
  const uint32_t sbase = smem_addr(smem);
  const uint32_t bar = sbase + S::bar, bar_w = bar + 8;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::bar + 16);

  if constexpr (X2)
    stage_bytes(smem + S::w0, image + x2_off, 2 * S::w0_part);   // [W0 hi | W0 lo] lead the tone-map x2 section
  else
    stage_bytes(smem + S::w0, image, TC_W * 48 * 2);
  stage_bytes(smem + S::wot, image + bwd_off, TC_W * 16 * 2 + 48 * TC_W * 2);   // Wo^T and W0^T are adjacent in the image
  stage_bytes(smem + S::bias, image + bias_off, TC_W * 4);
  for (int i = threadIdx.x; i < TC_TM; i += blockDim.x) {   // constant zero chunks
    *reinterpret_cast<uint4 *>(smem + S::x + 7 * (TC_TM * 16) + i * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4 *>(smem + S::dzo + TC_TM * 16 + i * 16) = make_uint4(0, 0, 0, 0);
    if constexpr (X2) *reinterpret_cast<uint4 *>(smem + S::dzw + TC_TM * 16 + i * 16) = make_uint4(0, 0, 0, 0);
  }
  [[maybe_unused]] float *s_inv = reinterpret_cast<float *>(smem + S::sinv);
  [[maybe_unused]] uint32_t *s_absmax = reinterpret_cast<uint32_t *>(smem + S::sinv + TC_TM * 4);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_w, 1);
    fence_mbar_init();
    if constexpr (X2) *s_absmax = 0u;
  }
  if (is_issuer) tmem_alloc(smem_addr(tmem_slot), TM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float *sbias = reinterpret_cast<const float *>(smem + S::bias);

  const int64_t n_tiles = (m + TC_TM - 1) / TC_TM;
  const EpiThread et(warp, lane);
  const int t = et.row_in_tile;
  uint32_t phase = 0;
  float bo_acc[3] = {0.f, 0.f, 0.f};
  int it = 0;
  [[maybe_unused]] float g_scale = 1.f, g_inv = 1.f;
  if constexpr (X2) {   // first pass: max |d_y| over the rows of this CTA's tiles (|act'| <= 1: a bound on every |dZ_out|)
    // (one 16-byte load per thread per tile — a tile's 128 x n_out floats start 16-byte aligned — and four tiles in flight:
    // a scalar loop over one tile at a time spent 50 us here waiting on one load after another)
    float mx = 0.f;
    const bool vec = (reinterpret_cast<uintptr_t>(d_y) & 15) == 0;   // (a row-sliced view can start 12 bytes into a word)
    const int per_tile4 = TC_TM * n_out / 4;
    const int64_t total4 = vec ? m * n_out / 4 : 0;
    const float4 *dy4 = reinterpret_cast<const float4 *>(d_y);
    if (!vec) {
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TC_TM * n_out, end = min(base + (int64_t)TC_TM * n_out, m * n_out);
        for (int64_t i = base + threadIdx.x; i < end; i += blockDim.x) mx = fmaxf(mx, fabsf(__ldg(d_y + i)));
      }
    }
    for (int64_t tile = blockIdx.x; vec && tile < n_tiles; tile += 4 * (int64_t)gridDim.x) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t tl = tile + (int64_t)u * gridDim.x;
        const int64_t i = tl * per_tile4 + threadIdx.x;
        v[u] = (tl < n_tiles && (int)threadIdx.x < per_tile4 && i < total4) ? __ldg(dy4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
    }
    if (vec) {   // the (< 4) floats behind the last whole 16 bytes (every CTA: whoever owns the last tile needs them)
      for (int64_t i = total4 * 4 + threadIdx.x; i < m * n_out; i += blockDim.x) mx = fmaxf(mx, fabsf(__ldg(d_y + i)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    if (lane == 0 && mx > 0.f) atomicMax(s_absmax, __float_as_uint(mx));
    __syncthreads();
    g_scale = act_dz_scale(*s_absmax, g_inv);
  }

  // Per-tile global inputs (lin, y, d_y, d_direct of this thread's row) are requested one tile ahead: read at the top of
  // T0 they stalled all 16 warps on the long scoreboard at once, every tile (5.1 of the 12 stalled warps per issue, ncu).
  float pf_lin = 0.f, pf_dd = 0.f, pf_y[3] = {0.f, 0.f, 0.f}, pf_dy[3] = {0.f, 0.f, 0.f};
  auto prefetch = [&](int64_t tile) {
    const int64_t row = tile * TC_TM + t;
    const bool ok = is_epi && tile < n_tiles && row < m;
    pf_lin = (ok && et.grp < 3) ? __ldg(lin + 3 * row + et.grp) : 0.f;
    pf_dd = (ok && et.grp < 3 && d_direct) ? __ldg(d_direct + 3 * row + et.grp) : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const bool okc = ok && et.grp == 0 && c < n_out;
      pf_y[c] = okc ? __ldg(y + row * n_out + c) : 0.f;
      pf_dy[c] = okc ? __ldg(d_y + row * n_out + c) : 0.f;
    }
  };
  prefetch(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int64_t row = tile * TC_TM + t;
    const bool valid = is_epi && row < m;
    float xv = 0.f, sn[5], cs[5], dz[3] = {0.f, 0.f, 0.f};
    const float dd = pf_dd;
    [[maybe_unused]] float inv_s = 1.f;   // X2: 1 / (row scale), read in E2 (a barrier before the next tile's T0 rewrites it)
    // ---- T0: encoding + output cotangent tiles ----
    if (is_epi) {
      if (it > 0) mbar_wait(bar_w, (uint32_t)((it - 1) & 1));   // the weight-gradient MMAs that read the tiles have retired
      if (et.grp < 3) {
        xv = pf_lin;
        uint4 lo, hi;
        if constexpr (X2) {   // the fp16 hi tile is also the weight-gradient GEMM's operand (dW0 += dZ0^T X)
          uint4 ll, lh;
          tonemap_pe_channel_x2(xv, lo, hi, ll, lh, sn, cs);
          if (!valid) ll = lh = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4 *>(smem + S::xl + (2 * et.grp) * (TC_TM * 16) + t * 16) = ll;
          *reinterpret_cast<uint4 *>(smem + S::xl + (2 * et.grp + 1) * (TC_TM * 16) + t * 16) = lh;
        } else {
          tonemap_pe_channel(xv, lo, hi, sn, cs);
        }
        if (!valid) lo = hi = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(smem + S::x + (2 * et.grp) * (TC_TM * 16) + t * 16) = lo;
        *reinterpret_cast<uint4 *>(smem + S::x + (2 * et.grp + 1) * (TC_TM * 16) + t * 16) = hi;
      } else {   // the "ones" column (bias gradient); zero for rows past the end
        *reinterpret_cast<uint4 *>(smem + S::x + 6 * (TC_TM * 16) + t * 16) =
            make_uint4(valid ? (X2 ? 0x00003c00u : 0x00003f80u) : 0u, 0u, 0u, 0u);
      }
      if (et.grp == 0) {
        if (valid) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < n_out) {
              const float yy = pf_y[c];
              dz[c] = pf_dy[c] * (act == 1 ? (1.f - expf(-yy)) : (act == 2 ? yy * (1.f - yy) : 1.f));
            }
        }
        if constexpr (X2) {
          const float mxa = fmaxf(fmaxf(fabsf(dz[0]), fabsf(dz[1])), fabsf(dz[2]));
          const int e = (__float_as_int(mxa) >> 23) & 0xff;          // biased exponent of the row's largest |dZ_out|
          const bool scaled = e >= 8 && e <= 250;                    // zero / denormal / absurd rows travel unscaled
          const float sr = scaled ? __int_as_float((258 - e) << 23) : 1.f;   // 2^(4 - (e - 127))
          s_inv[t] = scaled ? __int_as_float((e - 4) << 23) : 1.f;
          *reinterpret_cast<uint4 *>(smem + S::dzo + t * 16) = make_uint4(pack2h(dz[0] * sr, dz[1] * sr), pack2h(dz[2] * sr, 0.f), 0u, 0u);
          *reinterpret_cast<uint4 *>(smem + S::dzw + t * 16) =
              make_uint4(pack2h_sat(dz[0] * g_scale, dz[1] * g_scale), pack2h_sat(dz[2] * g_scale, 0.f), 0u, 0u);
        } else {
          *reinterpret_cast<uint4 *>(smem + S::dzo + t * 16) = make_uint4(pack2(dz[0], dz[1]), pack2(dz[2], 0.f), 0u, 0u);
        }
      }
      fence_proxy_async();
    }
    prefetch(tile + gridDim.x);
    tc_fence_before();
    __syncthreads();
    if (is_issuer && lane == 0) {
      tc_fence_after();
      if constexpr (X2) {
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int term = 0; term < 3; ++term)   // hi.hi, hi.lo, lo.hi
            mma_ss(tmem + TMB_D, make_desc(sbase + (term == 2 ? S::xl : S::x) + 2 * s * (TC_TM * 16), TC_TM * 16, 128),
                   make_desc(sbase + S::w0 + (term == 1 ? S::w0_part : 0) + 2 * s * (TC_W * 16), TC_W * 16, 128),
                   make_idesc_h(TC_W), (s | term) != 0);
      } else {
#pragma unroll
        for (int s = 0; s < 3; ++s)
          mma_ss(tmem + TMB_D, make_desc(sbase + S::x + 2 * s * (TC_TM * 16), TC_TM * 16, 128),
                 make_desc(sbase + S::w0 + 2 * s * (TC_W * 16), TC_W * 16, 128), make_idesc(TC_W), s > 0);
      }
      mma_commit(bar);
    }
    // ---- E1: H = relu(Z0 + b0) -> shared tile + masks ----
    uint32_t mask[2] = {0u, 0u};
    if (is_epi) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t r[3][16];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) tmem_ld16(tmem + et.lane_base + TMB_D + TC_GCOLS * et.grp + 16 * cc, r[cc]);
      tmem_ld_wait();
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int col0 = TC_GCOLS * et.grp + 16 * cc;
        uint32_t p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 bb = *reinterpret_cast<const float2 *>(sbias + col0 + 2 * j);
          float z0 = __uint_as_float(r[cc][2 * j]), z1 = __uint_as_float(r[cc][2 * j + 1]);
          if constexpr (X2) {
            fma2(z0, z1, 1.f / TC_WSCALE, bb);
            p[j] = pack2h_relu(z0, z1);
          } else {
            add2(z0, z1, bb);
            p[j] = pack2_relu(z0, z1);
          }
          const uint32_t tt = p[j] + 0x7fff7fffu;
          mask[cc >> 1] |= (tt >> (15 - 8 * (cc & 1) - j)) & (0x00010001u << (8 * (cc & 1) + j));
        }
        *reinterpret_cast<uint4 *>(smem + S::hs + (col0 / 8) * (TC_TM * 16) + t * 16) = make_uint4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<uint4 *>(smem + S::hs + (col0 / 8 + 1) * (TC_TM * 16) + t * 16) = make_uint4(p[4], p[5], p[6], p[7]);
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    if (is_issuer && lane == 0) {
      tc_fence_after();
      mma_ss(tmem + TMB_D, make_desc(sbase + S::dzo, TC_TM * 16, 128), make_desc(sbase + S::wot, TC_W * 16, 128),
             X2 ? make_idesc_h(TC_W) : make_idesc(TC_W), 0);
      mma_commit(bar);
      // dWo^T += Hs^T dZ_out (rows are the K dimension: both tiles are MN-major operands as they lie)
#pragma unroll
      for (int s = 0; s < TC_TM / 16; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          mma_ss(tmem + TMB_GO + 16 * h, make_desc(sbase + S::hs + h * 8 * (TC_TM * 16) + s * 256, 128, TC_TM * 16),
                 make_desc(sbase + (X2 ? S::dzw : S::dzo) + s * 256, 128, TC_TM * 16),
                 X2 ? make_idesc_h_mn(16) : make_idesc_mn(16), (it | s) != 0);
    }
    // ---- E2: dZ0 = dH * mask -> TMEM A operand + shared tile ----
    if (is_epi) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t r[3][16];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) tmem_ld16(tmem + et.lane_base + TMB_D + TC_GCOLS * et.grp + 16 * cc, r[cc]);
      tmem_ld_wait();
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int col0 = TC_GCOLS * et.grp + 16 * cc;
        const uint32_t mk = mask[cc >> 1] >> (8 * (cc & 1));
        uint32_t p[8];
        if constexpr (X2) {   // chain operand: fp16 carrying the row's scale; shared tile: fp16 times the CTA's scale
          inv_s = s_inv[t];
          const float z_s = inv_s * g_scale;
          uint32_t pa[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v0 = __uint_as_float(r[cc][2 * j]), v1 = __uint_as_float(r[cc][2 * j + 1]);
            const uint32_t pm = pair_mask(mk, j);
            pa[j] = pack2h(v0, v1) & pm;
            p[j] = pack2h_sat(v0 * z_s, v1 * z_s) & pm;
          }
          tmem_st8(tmem + et.lane_base + TMB_A + col0 / 2, pa);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            p[j] = pack2(__uint_as_float(r[cc][2 * j]), __uint_as_float(r[cc][2 * j + 1])) & pair_mask(mk, j);
          tmem_st8(tmem + et.lane_base + TMB_A + col0 / 2, p);
        }
        *reinterpret_cast<uint4 *>(smem + S::zs + (col0 / 8) * (TC_TM * 16) + t * 16) = make_uint4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<uint4 *>(smem + S::zs + (col0 / 8 + 1) * (TC_TM * 16) + t * 16) = make_uint4(p[4], p[5], p[6], p[7]);
      }
      tmem_st_wait();
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    if (is_issuer && lane == 0) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < TC_W / 16; ++s)
        mma_ts(tmem + TMB_S, tmem + TMB_A + 8 * s, make_desc(sbase + S::w0t + 2 * s * (48 * 16), 48 * 16, 128),
               X2 ? make_idesc_h(48) : make_idesc(48), s > 0);
      mma_commit(bar);
      // dW0 | db0 += Zs^T [X | 1]
#pragma unroll
      for (int s = 0; s < TC_TM / 16; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          mma_ss(tmem + TMB_G0 + 64 * h, make_desc(sbase + S::zs + h * 8 * (TC_TM * 16) + s * 256, 128, TC_TM * 16),
                 make_desc(sbase + S::x + s * 256, 128, TC_TM * 16), X2 ? make_idesc_h_mn(64) : make_idesc_mn(64), (it | s) != 0);
      mma_commit(bar_w);
    }
    // ---- E3: d_lin of (row, channel) from the 16 encoded-column cotangents of the channel ----
    if (is_epi) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      if (et.grp < 3) {
        uint32_t r[16];
        tmem_ld16(tmem + et.lane_base + TMB_S + 16 * et.grp, r);
        tmem_ld_wait();
        if (valid) {
          float g = __uint_as_float(r[0]);
#pragma unroll
          for (int f = 0; f < 5; ++f)
            g += (float)(1 << f) * (cs[f] * __uint_as_float(r[1 + f]) - sn[f] * __uint_as_float(r[6 + f]));
          if constexpr (X2) g *= inv_s;   // the chain carried the row's scale
          d_lin[3 * row + et.grp] = g + dd;
        }
      }
      if (et.grp == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) bo_acc[c] += warp_sum(dz[c]);
      }
      tc_fence_before();
    }
  }
  // ---- weight gradients: TMEM accumulators -> global (REDs) ----
  if (it > 0) {
    if (is_epi) {
      mbar_wait(bar_w, (uint32_t)((it - 1) & 1));
      tc_fence_after();
    }
    if (warp < 4) {
      const uint32_t lane_base = (32u * warp) << 16;
      const int ml = 32 * warp + lane;   // accumulator row = hidden feature (second half: feature 64 + row)
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int o = h == 0 ? ml : 64 + ml;
        const bool use = h == 0 || ml >= 64;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t r[16];
          tmem_ld16(tmem + lane_base + TMB_G0 + 64 * h + 16 * cc, r);
          tmem_ld_wait();
          if (!use) continue;
          if (cc < 3) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              red_add4(gW0 + (int64_t)o * 48 + cc * 16 + 4 * q, __uint_as_float(r[4 * q]) * g_inv,
                       __uint_as_float(r[4 * q + 1]) * g_inv, __uint_as_float(r[4 * q + 2]) * g_inv,
                       __uint_as_float(r[4 * q + 3]) * g_inv);
          } else {
            red_add(gb0 + o, __uint_as_float(r[0]) * g_inv);
          }
        }
        uint32_t r[16];
        tmem_ld16(tmem + lane_base + TMB_GO + 16 * h, r);
        tmem_ld_wait();
        if (use) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < n_out) red_add(gWo + (int64_t)c * TC_W + o, __uint_as_float(r[c]) * g_inv);
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < n_out) red_add(gbo + c, bo_acc[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (is_issuer) tmem_dealloc(tmem, TM_COLS);
}

template <typename K>
static int set_smem_tc(K kernel, int bytes) {
  ESR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return ESR_OK;
}

static unsigned tc_grid(int64_t rows) {
  const int64_t tiles = (rows + TC_TM - 1) / TC_TM;
  const int64_t sms = num_sms();
  return (unsigned)(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
}

// ESR_MLP_TILE_OVERLAP=1 (read once): the radiance-net chains run their tile-overlap instantiation (k_mlp_fwd_tc /
// k_mlp_dgrad_tc, OVL).  Off by default: that variant has not run on a GPU in its present form (its predecessor, which
// shared one mbarrier between the layer-0 and the output commits, measured -0.07 / -0.05 ms per launch and was
// withdrawn for the double-completion hazard the separate mbarrier removes) — DESIGN.md §4.
static bool tile_overlap() {
  static const bool on = [] {
    const char *e = getenv("ESR_MLP_TILE_OVERLAP");
    return e && e[0] && e[0] != '0';
  }();
  return on;
}

template <int K0, int NH, int NO, int XSRC = 0, bool OVL = false>
static int launch_fwd(const esr_mlp_desc_t *d, const void *image, const void *x, int64_t rb, int64_t re, int64_t mt,
                      float *y, void *hidden, int64_t save_begin, cudaStream_t st) {
  if constexpr (!OVL && K0 == 96 && XSRC == 0)
    if (tile_overlap()) return launch_fwd<K0, NH, NO, XSRC, true>(d, image, x, rb, re, mt, y, hidden, save_begin, st);
  auto kern = k_mlp_fwd_tc<K0, NH, NO, XSRC, OVL>;
  constexpr int bytes = FwdSm<K0, NH>::bytes;
  if (int e = set_smem_tc(kern, bytes)) return e;
  ESR_STAGE(K0 == 96 ? "k_mlp_fwd_tc_radiance" : (XSRC ? "k_tonemap_fwd_fused" : "k_mlp_fwd_tc_tonemap"), st);
  kern<<<tc_grid(re - rb), TC_THREADS, bytes, st>>>((const uint8_t *)image, (const __nv_bfloat16 *)x, rb, re, mt, y,
                                                    (__nv_bfloat16 *)hidden, save_begin, d->n_out, d->act);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

template <int K0, int NH, int NO>
static int launch_fwd_x2(const esr_mlp_desc_t *d, const TcLayout &T, const void *image, const void *x, int64_t rb, int64_t re,
                         int64_t mt, float *y, void *hidden, int64_t save_begin, cudaStream_t st) {
  auto kern = k_mlp_fwd_x2<K0, NH, NO>;
  using S = X2Sm<K0, NH, NO>;
  if (T.x2_rank_bytes() != S::rank_bytes) {
    set_error("x2 forward: image layout mismatch");
    return ESR_ERR_BAD_ARG;
  }
  if (int e = set_smem_tc(kern, S::bytes)) return e;
  const int64_t pair_tiles = (re - rb + 2 * TC_TM - 1) / (2 * TC_TM);
  const int64_t pairs = max((int64_t)1, min((int64_t)(num_sms() / 2), pair_tiles));
  // the residual tile of the feature rows follows the bf16 tile (esr_encode_*_fwd with out_is_bf16 = 2)
  const __half *x_lo = reinterpret_cast<const __half *>(reinterpret_cast<const __nv_bfloat16 *>(x) + act_rows_padded(mt) * K0);
  ESR_STAGE("k_mlp_fwd_x2_radiance", st);
  kern<<<(unsigned)(2 * pairs), TC_THREADS, S::bytes, st>>>((const uint8_t *)image + T.x2_off(), (const __nv_bfloat16 *)x, x_lo,
                                                           rb, re, mt, y, (__nv_bfloat16 *)hidden, save_begin, d->n_out, d->act);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

template <int K0, int NH, int DXN, int NO, bool ACC, bool OVL = false, bool H16 = false>
static int launch_dgrad_acc(const esr_mlp_desc_t *d, const TcLayout &T, const void *image, const float *y, const float *d_y,
                        int64_t rb, int64_t re, int64_t mt, const void *hidden, void *d_z, float *d_z_out, float *d_x,
                        int dx_cols, int accumulate, cudaStream_t st) {
  if constexpr (!OVL && !H16 && K0 == 96) {
    if (T.bwd_h16())   // precision 1: the image's transposed matrices are fp16 -> the fp16 chain
      return launch_dgrad_acc<K0, NH, DXN, NO, ACC, false, true>(d, T, image, y, d_y, rb, re, mt, hidden, d_z, d_z_out, d_x,
                                                                 dx_cols, accumulate, st);
    if (tile_overlap())
      return launch_dgrad_acc<K0, NH, DXN, NO, ACC, true>(d, T, image, y, d_y, rb, re, mt, hidden, d_z, d_z_out, d_x, dx_cols,
                                                          accumulate, st);
  }
  auto kern = k_mlp_dgrad_tc<K0, NH, DXN, NO, ACC, OVL, H16>;
  constexpr int bytes = BwdSm<K0, NH, DXN>::bytes + (H16 ? TC_TM * 4 : 0);
  if (int e = set_smem_tc(kern, bytes)) return e;
  uint32_t *absmax = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(d_z) + act_dz_tail_offset(NH, mt));
  if constexpr (H16) {   // max |d_y| of the launch -> the scale of the stored cotangents
    ESR_CHECK_CUDA(cudaMemsetAsync(absmax, 0, 4, st));
    const int64_t n = (re - rb) * d->n_out;
    ESR_STAGE("k_absmax", st);
    k_absmax<<<(unsigned)min((int64_t)num_sms() * 8, (int64_t)cdiv(n, 256)), 256, 0, st>>>(d_y + rb * d->n_out, n, absmax);
    ESR_LAUNCH_OK();
  }
  ESR_STAGE(K0 == 96 ? "k_mlp_dgrad_tc_radiance" : "k_mlp_dgrad_tc_tonemap", st);
  kern<<<tc_grid(re - rb), TC_THREADS, bytes, st>>>((const uint8_t *)image + T.bwd_off(), y, d_y, rb, re, mt,
                                                    (const __nv_bfloat16 *)hidden, (__nv_bfloat16 *)d_z, d_z_out, d_x,
                                                    dx_cols, accumulate, d->n_out, d->act, absmax);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

template <int K0, int NH, int DXN, int NO>
static int launch_dgrad(const esr_mlp_desc_t *d, const TcLayout &T, const void *image, const float *y, const float *d_y,
                        int64_t rb, int64_t re, int64_t mt, const void *hidden, void *d_z, float *d_z_out, float *d_x,
                        int dx_cols, int accumulate, cudaStream_t st) {
  // the accumulate variant carries the prefetched old d_x values in 16 more registers: separate instantiation so the
  // write-only variant (the fine stage's disjoint row ranges) keeps its register budget
  if (accumulate && d_x)
    return launch_dgrad_acc<K0, NH, DXN, NO, true>(d, T, image, y, d_y, rb, re, mt, hidden, d_z, d_z_out, d_x, dx_cols,
                                                   accumulate, st);
  return launch_dgrad_acc<K0, NH, DXN, NO, false>(d, T, image, y, d_y, rb, re, mt, hidden, d_z, d_z_out, d_x, dx_cols,
                                                  accumulate, st);
}

}  // namespace

namespace esr {

bool tc_supported(const esr_mlp_desc_t *d) {
  return d && d->width == TC_W && ((d->k0 == 96 && d->n_hidden == 3) || (d->k0 == 48 && d->n_hidden == 1));
}

int64_t tc_image_bytes(const esr_mlp_desc_t *d) { return tc_supported(d) ? tc_layout(d).total() : 0; }

int tc_pack(const esr_mlp_desc_t *d, const float *flat_params, void *tc_image, cudaStream_t st) {
  if (!tc_supported(d)) return ESR_OK;
  const TcLayout T = tc_layout(d);
  const MlpLayout L = layout_of(d);
  ESR_CHECK_CUDA(cudaMemsetAsync(tc_image, 0, (size_t)T.total(), st));
  const int64_t n = (int64_t)TC_W * T.k0 + (int64_t)(T.NH - 1) * TC_W * TC_W + (int64_t)TC_NOUT_PAD * TC_W +
                    (int64_t)T.NH * TC_W + TC_NOUT_PAD;
  ESR_STAGE("k_tc_pack", st);
  k_tc_pack<<<cdiv(n, 256), 256, 0, st>>>(T, L, flat_params, (uint8_t *)tc_image);
  ESR_LAUNCH_OK();
  if (T.x2) {
    ESR_STAGE("k_tc_pack", st);
    k_tc_pack_x2<<<cdiv(n, 256), 256, 0, st>>>(T, L, flat_params, (uint8_t *)tc_image);
    ESR_LAUNCH_OK();
  }
  return ESR_OK;
}

int tc_fwd(const esr_mlp_desc_t *d, const void *tc_image, const void *x, int64_t row_begin, int64_t row_end,
           int64_t m_total, float *y, void *hidden, int64_t save_begin, cudaStream_t st) {
  if (d->precision == 1) {
    const TcLayout T = tc_layout(d);
    if (d->k0 == 96 && d->n_hidden == 3 && d->n_out <= 3)
      return launch_fwd_x2<96, 3, 3>(d, T, tc_image, x, row_begin, row_end, m_total, y, hidden, save_begin, st);
    if (d->k0 == 96 && d->n_hidden == 3)
      return launch_fwd_x2<96, 3, 8>(d, T, tc_image, x, row_begin, row_end, m_total, y, hidden, save_begin, st);
    set_error("tc_fwd: the x2 forward chain is instantiated for the 96 -> 192 x 3 nets (the tone-map net: esr_tonemap_mlp_*)");
    return ESR_ERR_BAD_ARG;
  }
  if (d->k0 == 96 && d->n_hidden == 3 && d->n_out <= 3)
    return launch_fwd<96, 3, 3>(d, tc_image, x, row_begin, row_end, m_total, y, hidden, save_begin, st);
  if (d->k0 == 96 && d->n_hidden == 3)
    return launch_fwd<96, 3, 8>(d, tc_image, x, row_begin, row_end, m_total, y, hidden, save_begin, st);
  if (d->k0 == 48 && d->n_hidden == 1)
    return launch_fwd<48, 1, 3>(d, tc_image, x, row_begin, row_end, m_total, y, hidden, save_begin, st);
  set_error("tc_fwd: shape not instantiated");
  return ESR_ERR_BAD_ARG;
}

int tc_tonemap_fwd(const esr_mlp_desc_t *d, const void *tc_image, const float *lin, int64_t m, float *y, cudaStream_t st) {
  if (getenv("ESR_TONEMAP_FWD_GENERIC"))   // the one-tile-at-a-time generic chain (kept for A/B measurements)
    return launch_fwd<48, 1, 3, 1>(d, tc_image, lin, 0, m, m, y, nullptr, 0, st);
  if (d->precision == 1) {
    auto kern = k_tonemap_fwd2<3, true>;
    using S2 = Tm2SmT<true>;
    if (int e = set_smem_tc(kern, S2::bytes)) return e;
    ESR_STAGE("k_tonemap_fwd_fused", st);
    kern<<<tc_grid(m), TC_THREADS, S2::bytes, st>>>((const uint8_t *)tc_image + tc_layout(d).x2_off(), lin, m, y, d->n_out, d->act);
    ESR_LAUNCH_OK();
    return ESR_OK;
  }
  auto kern = k_tonemap_fwd2<3>;
  if (int e = set_smem_tc(kern, Tm2Sm::bytes)) return e;
  ESR_STAGE("k_tonemap_fwd_fused", st);
  kern<<<tc_grid(m), TC_THREADS, Tm2Sm::bytes, st>>>((const uint8_t *)tc_image, lin, m, y, d->n_out, d->act);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

int tc_tonemap_bwd(const esr_mlp_desc_t *d, const void *tc_image, const float *lin, const float *y, const float *d_y,
                   const float *d_direct, int64_t m, float *d_lin, float *grad_flat, cudaStream_t st) {
  const TcLayout T = tc_layout(d);
  const MlpLayout L = layout_of(d);
  if (d->precision == 1) {
    auto kern = k_tonemap_bwd_fused<true>;
    if (int e = set_smem_tc(kern, TmBwdSmT<true>::bytes)) return e;
    ESR_STAGE("k_tonemap_bwd_fused", st);
    kern<<<tc_grid(m), TC_THREADS, TmBwdSmT<true>::bytes, st>>>(
        (const uint8_t *)tc_image, T.bwd_off(), T.f_bias(), T.x2_off(), lin, y, d_y, d_direct, m, d_lin, grad_flat + L.flat_w(0),
        grad_flat + L.flat_b(0), grad_flat + L.flat_w(1), grad_flat + L.flat_b(1), d->n_out, d->act);
    ESR_LAUNCH_OK();
    return ESR_OK;
  }
  auto kern = k_tonemap_bwd_fused<false>;
  if (int e = set_smem_tc(kern, TmBwdSm::bytes)) return e;
  ESR_STAGE("k_tonemap_bwd_fused", st);
  kern<<<tc_grid(m), TC_THREADS, TmBwdSm::bytes, st>>>((const uint8_t *)tc_image, T.bwd_off(), T.f_bias(), 0, lin, y, d_y,
                                                       d_direct, m, d_lin, grad_flat + L.flat_w(0), grad_flat + L.flat_b(0),
                                                       grad_flat + L.flat_w(1), grad_flat + L.flat_b(1), d->n_out, d->act);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

int tc_dgrad(const esr_mlp_desc_t *d, const void *tc_image, const float *y, const float *d_y, int64_t row_begin,
             int64_t row_end, int64_t m_total, const void *hidden, void *d_z, float *d_z_out, float *d_x, int dx_cols,
             int accumulate, cudaStream_t st) {
  const TcLayout T = tc_layout(d);
  if (d->precision == 1 && d->k0 != 96) {
    set_error("tc_dgrad: with precision 1 the tone-map net runs through esr_tonemap_mlp_bwd (fused kernel)");
    return ESR_ERR_BAD_ARG;
  }
  if (d->k0 == 96 && d->n_hidden == 3 && d->n_out <= 3)
    return launch_dgrad<96, 3, 64, 3>(d, T, tc_image, y, d_y, row_begin, row_end, m_total, hidden, d_z, d_z_out, d_x,
                                      dx_cols, accumulate, st);
  if (d->k0 == 96 && d->n_hidden == 3)
    return launch_dgrad<96, 3, 64, 8>(d, T, tc_image, y, d_y, row_begin, row_end, m_total, hidden, d_z, d_z_out, d_x,
                                      dx_cols, accumulate, st);
  if (d->k0 == 48 && d->n_hidden == 1)
    return launch_dgrad<48, 1, 48, 3>(d, T, tc_image, y, d_y, row_begin, row_end, m_total, hidden, d_z, d_z_out, d_x,
                                      dx_cols, accumulate, st);
  set_error("tc_dgrad: shape not instantiated");
  return ESR_ERR_BAD_ARG;
}

int tc_wgrad(const esr_mlp_desc_t *d, const void *x, int64_t row_begin, int64_t row_end, int64_t m_total,
             const void *hidden, const void *d_z, float *grad_flat, cudaStream_t st) {
  const MlpLayout L = layout_of(d);
  const int64_t ls = act_rows_padded(m_total) * TC_W;
  const int NH = d->n_hidden;
  const __nv_bfloat16 *H = (const __nv_bfloat16 *)hidden, *Z = (const __nv_bfloat16 *)d_z;
  const uint32_t *absmax = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(d_z) + act_dz_tail_offset(NH, m_total));
  int e;
  if (tc_layout(d).bwd_h16()) {   // precision 1: x (first tile set), H_l and the G-scaled dZ_l are fp16
    if ((e = launch_wgrad<96, 24, true>(Z, (const __nv_bfloat16 *)x, row_begin, row_end, TC_W, grad_flat + L.flat_w(0),
                                        grad_flat + L.flat_b(0), absmax, st)))
      return e;
    for (int l = 1; l < NH; ++l)
      if ((e = launch_wgrad<192, 24, true>(Z + l * ls, H + (l - 1) * ls, row_begin, row_end, TC_W, grad_flat + L.flat_w(l),
                                           grad_flat + L.flat_b(l), absmax, st)))
        return e;
    return launch_wgrad<192, 2, true>(Z + NH * ls, H + (NH - 1) * ls, row_begin, row_end, d->n_out, grad_flat + L.flat_w(NH),
                                      grad_flat + L.flat_b(NH), absmax, st);
  }
  if (d->k0 == 96)
    e = launch_wgrad<96, 24>(Z, (const __nv_bfloat16 *)x, row_begin, row_end, TC_W, grad_flat + L.flat_w(0),
                             grad_flat + L.flat_b(0), absmax, st);
  else
    e = launch_wgrad<48, 24>(Z, (const __nv_bfloat16 *)x, row_begin, row_end, TC_W, grad_flat + L.flat_w(0),
                             grad_flat + L.flat_b(0), absmax, st);
  if (e) return e;
  for (int l = 1; l < NH; ++l)
    if ((e = launch_wgrad<192, 24>(Z + l * ls, H + (l - 1) * ls, row_begin, row_end, TC_W, grad_flat + L.flat_w(l),
                                   grad_flat + L.flat_b(l), absmax, st)))
      return e;
  // output layer: A = dZ_out (tiled, 2 chunks per row, stored after the hidden-layer dZ), In = H_{NH-1}
  return launch_wgrad<192, 2>(Z + NH * ls, H + (NH - 1) * ls, row_begin, row_end, d->n_out, grad_flat + L.flat_w(NH),
                              grad_flat + L.flat_b(NH), absmax, st);
}

}  // namespace esr

// scan.cuh — device-wide prefix sums over per-ray counts (three small launches: block sums,
// scan of block sums, block scan + offset).  Replaces the reference's ATen/CUB cumsum calls
// (render_utils_kernel.cu:211,216).
#pragma once
#include "common.cuh"

namespace esr {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 elements per block

template <typename T>
ESR_D T warp_incl_scan(T v) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = __shfl_up_sync(FULL, v, o);
    if (lane >= (unsigned)o) v += u;
  }
  return v;
}

// inclusive scan of one value per thread across the block; returns inclusive value, sets total
template <typename T>
ESR_D T block_incl_scan(T v, T *smem /* >= 32 */, T &total) {
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  T inc = warp_incl_scan(v);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = (lane < nwarp) ? smem[lane] : T(0);
    w = warp_incl_scan(w);
    smem[lane] = w;
  }
  __syncthreads();
  const T base = warp ? smem[warp - 1] : T(0);
  total = smem[nwarp - 1];
  __syncthreads();
  return inc + base;
}

template <typename Tin, typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const Tin *__restrict__ in, int64_t n,
                                                                T *__restrict__ block_sums) {
  __shared__ T sm[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  T s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int64_t idx = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) s += (T)in[idx];
  }
  T total;
  block_incl_scan(s, sm, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place; writes grand total to *total_out
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_of_sums(T *__restrict__ block_sums, int nblocks,
                                                             T *__restrict__ total_out) {
  __shared__ T sm[32];
  T carry = 0;
  for (int b0 = 0; b0 < nblocks; b0 += SCAN_THREADS) {
    const int i = b0 + threadIdx.x;
    const T v = (i < nblocks) ? block_sums[i] : T(0);
    T total;
    const T inc = block_incl_scan(v, sm, total);
    if (i < nblocks) block_sums[i] = carry + inc - v;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// EXCLUSIVE != 0: out[i] = sum_{j<i}; else inclusive.  Thread-contiguous items.
template <typename Tin, typename T, bool EXCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const Tin *__restrict__ in, int64_t n,
                                                           const T *__restrict__ block_offsets,
                                                           T *__restrict__ out) {
  __shared__ T sm[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? (T)in[base + i] : T(0);
    s += v[i];
  }
  T total;
  const T inc = block_incl_scan(s, sm, total);
  T run = block_offsets[blockIdx.x] + inc - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = EXCLUSIVE ? run : run + v[i];
    run += v[i];
  }
}

// scratch: (nblocks + 1) elements of T.  If EXCLUSIVE, out[n] = total as well (out has n+1 slots).
template <typename Tin, typename T, bool EXCLUSIVE>
int device_scan(const Tin *in, T *out, int64_t n, T *total_out, void *scratch, cudaStream_t st) {
  T *sums = (T *)scratch;
  if (n <= 0) {
    if (EXCLUSIVE) ESR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(T), st));
    if (total_out) ESR_CHECK_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), st));
    return ESR_OK;
  }
  const int nblocks = (int)cdiv(n, SCAN_TILE);
  T *tot = EXCLUSIVE ? out + n : (total_out ? total_out : sums + nblocks);
  ESR_STAGE("scan_block_sums", st);
  scan_block_sums<Tin, T><<<nblocks, SCAN_THREADS, 0, st>>>(in, n, sums);
  ESR_LAUNCH_OK();
  ESR_STAGE("scan_of_sums", st);
  scan_of_sums<T><<<1, SCAN_THREADS, 0, st>>>(sums, nblocks, tot);
  ESR_LAUNCH_OK();
  ESR_STAGE("scan_apply", st);
  scan_apply<Tin, T, EXCLUSIVE><<<nblocks, SCAN_THREADS, 0, st>>>(in, n, sums, out);
  ESR_LAUNCH_OK();
  if (EXCLUSIVE && total_out) ESR_CHECK_CUDA(cudaMemcpyAsync(total_out, tot, sizeof(T), cudaMemcpyDeviceToDevice, st));
  return ESR_OK;
}

}  // namespace esr

// grad_exchange.cu — the pack / unpack kernels either side of the one collective on the path: the
// gradient all-reduce of the ray-sharded render step.  The dense grid gradients (sdf 64 MB + 2-3 x 384 MB at 256^3)
// are non-zero only inside the dilated occupancy set (module.py:104-114 keeps samples inside MaskCache; trilinear taps
// and the multi-scale SDF taps of voxurff.py:678-721 reach a few voxels further), so the ranks exchange ONE packed
// buffer holding just those voxels.  One launch packs every volume (planar: volume j occupies rows [K][c_j] of the
// buffer, so both sides move 8-byte words and the packed side is perfectly coalesced), NCCL all-reduces the buffer in
// place, one launch writes the sums back.  The voxel list is sorted, and occupancy sets are runs along z, so the
// dense-side accesses are mostly full sectors.  The reference has no distributed code (cfg/__init__.yaml:24).
#include "common.cuh"

using namespace esr;

namespace {

struct RowSets {
  float *vol[ESR_MAX_GRAD_VOLUMES];        // dense volume j: [V][chan[j]] floats (channels-last)
  int64_t unit_end[ESR_MAX_GRAD_VOLUMES];  // exclusive prefix of copy units (8-byte words for even chan, else floats)
  int64_t buf_off[ESR_MAX_GRAD_VOLUMES];   // float offset of volume j's [K][chan[j]] block in the packed buffer
  int32_t chan[ESR_MAX_GRAD_VOLUMES];
  int32_t n;
};

template <bool PACK>
__global__ void __launch_bounds__(256)
    k_grad_rows(const __grid_constant__ RowSets rs, const int32_t *__restrict__ idx, float *__restrict__ buf) {
  const int64_t total = rs.unit_end[rs.n - 1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += stride) {
    int j = 0;
#pragma unroll
    for (int t = 0; t < ESR_MAX_GRAD_VOLUMES - 1; ++t) j += (t < rs.n - 1 && u >= rs.unit_end[t]) ? 1 : 0;
    const int64_t local = u - (j ? rs.unit_end[j - 1] : 0);
    const int c = rs.chan[j];
    float *packed = buf + rs.buf_off[j];
    if ((c & 1) == 0) {
      const int per = c >> 1;
      const int64_t k = local / per;
      const int part = (int)(local - k * per);
      float2 *dense = reinterpret_cast<float2 *>(rs.vol[j] + (int64_t)__ldg(idx + k) * c) + part;
      float2 *pk = reinterpret_cast<float2 *>(packed) + local;
      if (PACK) *pk = *dense;
      else *dense = *pk;
    } else {
      const int64_t k = local / c;
      const int part = (int)(local - k * c);
      float *dense = rs.vol[j] + (int64_t)__ldg(idx + k) * c + part;
      if (PACK) packed[local] = *dense;
      else *dense = packed[local];
    }
  }
}

int launch(bool pack, void *const *volumes, const int32_t *channels, int n_volumes, const int32_t *idx, int64_t k,
           float *buf, esr_stream_t stream) {
  ESR_CHECK_ARG(n_volumes >= 1 && n_volumes <= ESR_MAX_GRAD_VOLUMES && k >= 0);
  if (k == 0) return ESR_OK;
  ESR_CHECK_ARG(volumes && channels && idx && buf && (uintptr_t)buf % 8 == 0);
  RowSets rs;
  int64_t units = 0, off = 0;
  for (int j = 0; j < ESR_MAX_GRAD_VOLUMES; ++j) {
    const bool live = j < n_volumes;
    const int c = live ? channels[j] : 1;
    ESR_CHECK_ARG(!live || (volumes[j] && c >= 1 && (uintptr_t)volumes[j] % 8 == 0));
    rs.vol[j] = live ? (float *)volumes[j] : nullptr;
    rs.chan[j] = c;
    rs.buf_off[j] = off;
    if (live) {
      units += (c & 1) ? k * c : k * (c >> 1);
      off += k * c;
      off += off & 1;                           // every block starts 8-byte aligned in the buffer
    }
    rs.unit_end[j] = units;
  }
  rs.n = n_volumes;
  const int64_t want = (units + 255) / 256, cap = (int64_t)num_sms() * 16;
  ESR_STAGE(pack ? "k_grad_pack" : "k_grad_unpack", stream);
  if (pack)
    k_grad_rows<true><<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(rs, idx, buf);
  else
    k_grad_rows<false><<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(rs, idx, buf);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

// Touched-block map of the exact two-level exchange (dist.TouchedBlockCompactor): block b = (bx, by, bz) of edge
// (ex, ey, ez) voxels is ex * ey runs of ez * c contiguous floats in a channels-last volume.  One warp per
// (volume, block): lanes stride over the block's floats run by run (coalesced within a run), one vote, lane 0 sets the
// flag (every writer stores 1: no atomics).  Reads each volume once — HBM-stream bound.
struct BlockGeo {
  int32_t Y, Z;            // grid extent along y, z (x follows from the block count)
  int32_t ex, ey, ez;      // block edge per axis (divides the grid extent)
  int32_t by, bz;          // blocks along y, z
  int64_t n_blocks;
};

__global__ void __launch_bounds__(256)
    k_grad_block_flags(const __grid_constant__ RowSets rs, const __grid_constant__ BlockGeo g, int32_t *__restrict__ flags) {
  const unsigned lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = g.n_blocks * rs.n;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
    const int j = (int)(w / g.n_blocks);
    const int64_t b = w - (int64_t)j * g.n_blocks;
    const int64_t bz = b % g.bz, by = (b / g.bz) % g.by, bx = b / ((int64_t)g.bz * g.by);
    const int c = rs.chan[j];
    const int run = g.ez * c;                      // contiguous floats of one z-run
    const int per_block = g.ex * g.ey * run;
    const float *vol = rs.vol[j];
    bool nz = false;
    for (int e = (int)lane; e < per_block; e += 32) {
      const int r = e / run, t = e - r * run;      // run r = (i, jj) of the block's ex x ey footprint
      const int i = r / g.ey, jj = r - i * g.ey;
      const int64_t voxel = ((bx * g.ex + i) * g.Y + (by * g.ey + jj)) * g.Z + bz * g.ez;
      nz |= __ldg(vol + voxel * c + t) != 0.f;
    }
    if (__any_sync(FULL, nz) && lane == 0) flags[b] = 1;
  }
}

}  // namespace

extern "C" int esr_grad_block_flags(void *const *volumes, const int32_t *channels, int n_volumes, int32_t gx, int32_t gy,
                                    int32_t gz, int32_t ex, int32_t ey, int32_t ez, int32_t *flags, esr_stream_t stream) {
  ESR_CHECK_ARG(volumes && channels && flags && n_volumes >= 1 && n_volumes <= ESR_MAX_GRAD_VOLUMES);
  ESR_CHECK_ARG(gx > 0 && gy > 0 && gz > 0 && ex > 0 && ey > 0 && ez > 0 && gx % ex == 0 && gy % ey == 0 && gz % ez == 0);
  RowSets rs;
  for (int j = 0; j < ESR_MAX_GRAD_VOLUMES; ++j) {
    const bool live = j < n_volumes;
    ESR_CHECK_ARG(!live || (volumes[j] && channels[j] >= 1));
    rs.vol[j] = live ? (float *)volumes[j] : nullptr;
    rs.chan[j] = live ? channels[j] : 1;
    rs.unit_end[j] = 0;
    rs.buf_off[j] = 0;
  }
  rs.n = n_volumes;
  BlockGeo g;
  g.Y = gy, g.Z = gz, g.ex = ex, g.ey = ey, g.ez = ez;
  g.by = gy / ey, g.bz = gz / ez;
  g.n_blocks = (int64_t)(gx / ex) * g.by * g.bz;
  ESR_CHECK_CUDA(cudaMemsetAsync(flags, 0, (size_t)g.n_blocks * sizeof(int32_t), (cudaStream_t)stream));
  const int64_t want = (g.n_blocks * n_volumes + 7) / 8, cap = (int64_t)num_sms() * 32;   // 8 warps per CTA
  ESR_STAGE("k_grad_block_flags", stream);
  k_grad_block_flags<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(rs, g, flags);
  ESR_LAUNCH_OK();
  return ESR_OK;
}

extern "C" int64_t esr_grad_pack_floats(const int32_t *channels, int n_volumes, int64_t k) {
  if (!channels || n_volumes < 1 || n_volumes > ESR_MAX_GRAD_VOLUMES || k < 0) return -1;
  int64_t off = 0;
  for (int j = 0; j < n_volumes; ++j) {
    if (channels[j] < 1) return -1;
    off += k * channels[j];
    off += off & 1;
  }
  return off;
}

extern "C" int esr_grad_pack(void *const *volumes, const int32_t *channels, int n_volumes, const int32_t *idx,
                             int64_t k, float *buf, esr_stream_t stream) {
  return launch(true, volumes, channels, n_volumes, idx, k, buf, stream);
}

extern "C" int esr_grad_unpack(void *const *volumes, const int32_t *channels, int n_volumes, const int32_t *idx,
                               int64_t k, const float *buf, esr_stream_t stream) {
  return launch(false, volumes, channels, n_volumes, idx, k, const_cast<float *>(buf), stream);
}

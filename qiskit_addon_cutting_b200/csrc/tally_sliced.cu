// K1 bit-sliced, run-time masks: the ahead-of-time compiled fast path for groups that are not
// worth (or not eligible for) an NVRTC specialisation -- typically the many small commuting
// groups of a molecular-style Hamiltonian.
//
// Same loads / in-register 32x32 bit transposition as the specialised kernels
// (sliced_prelude.cuh); the bit-planes of a 32-shot batch are then parked in shared memory
// ([plane][lane], conflict free) and every term XORs the planes named by its mask bits:
// w shared-memory loads + w LOP3 + one POPC per 32 shots per term.  Lane t keeps the odd count of
// term t in a register across the whole tile (terms >= 32 fall back to one atomic per batch).
#include "common.cuh"
#include "sliced_prelude.cuh"

namespace qcut {

constexpr int kRtWarps = 8;
constexpr int kRtMaxNb = 8;                        // rows up to 8 bytes (64 memory-bit planes)
constexpr int kRtPlaneWords = kRtMaxNb * 8 * 32;   // per-warp plane storage: 64 planes x 32 lanes

template <int NB>
__device__ __forceinline__ void rt_tile(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd,
                                        int nbq, int ns, int lane, const uint32_t* __restrict__ mw, int n_words,
                                        int n_terms, uint32_t* __restrict__ s_planes,
                                        unsigned long long* __restrict__ tally) {
  typedef QcutGeo<NB> G;
  uint32_t mine = 0;  // odd count of term `lane` (terms 0..31)
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    uint32_t a[G::BLOCKS][32];
    uint32_t qp[G::BATCHES];
    qcut_step_any<NB>(obs + (uint64_t)s0 * NB, qpd + (uint64_t)s0 * nbq, nbq, n, lane, a, qp);
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u) {
      __syncwarp();
      // plane p (memory bit 8*B + b) of this batch -> s_planes[p][lane]
#pragma unroll
      for (int blk = 0; blk < G::BLOCKS; ++blk)
#pragma unroll
        for (int i = 0; i < G::PLANES; ++i) s_planes[(blk * 32 + i) * 32 + lane] = a[blk][u * G::PLANES + i];
      __syncwarp();
      for (int t = 0; t < n_terms; ++t) {
        const uint32_t m_lo = __ldg(mw + (size_t)t * n_words);
        const uint32_t m_hi = (NB > 4) ? __ldg(mw + (size_t)t * n_words + 1) : 0u;
        const uint32_t odd = __reduce_add_sync(0xffffffffu, qcut_rt_term(s_planes + lane, qp[u], m_lo, m_hi));
        if (t < 32) {
          if (lane == t) mine += odd;
        } else if (lane == 0 && odd) {
          atomicAdd(tally + t, (unsigned long long)(-2ll * (long long)odd));
        }
      }
    }
  }
  // tally[t] += ns - 2 * odd  (terms >= 32 already received their -2 * odd)
  for (int t = lane; t < n_terms; t += 32) {
    const long long v = (long long)ns - (t < 32 ? 2ll * (long long)mine : 0ll);
    atomicAdd(tally + t, (unsigned long long)v);
  }
}

__global__ void __launch_bounds__(kRtWarps * 32, 2)
tally_sliced_rt_kernel(const uint8_t* __restrict__ data, const qcut_pub_desc* __restrict__ pubs,
                       const TileEntry* __restrict__ tiles, int64_t n_tiles,
                       const GroupDev* __restrict__ groups, const uint32_t* __restrict__ mask_words,
                       unsigned long long* __restrict__ tallies) {
  extern __shared__ uint32_t s_all[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* s_planes = s_all + (size_t)warp * kRtPlaneWords;
  const int64_t n_warps = (int64_t)gridDim.x * kRtWarps;

  for (int64_t tile = (int64_t)blockIdx.x * kRtWarps + warp; tile < n_tiles; tile += n_warps) {
    const TileEntry te = tiles[tile];
    const qcut_pub_desc pub = pubs[te.pub];
    const GroupDev grp = groups[pub.group];
    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;
    const int64_t left = (int64_t)pub.shots - shot0;
    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;
    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * grp.nb_obs;
    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;
    const uint32_t* __restrict__ mw = mask_words + grp.word_offset;
    unsigned long long* __restrict__ tally = tallies + pub.tally_offset;
    const int nbq = (int)pub.nb_qpd, nw = (int)grp.n_words, T = (int)grp.n_terms;
    switch (grp.nb_obs) {
      case 1: rt_tile<1>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      case 2: rt_tile<2>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      case 3: rt_tile<3>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      case 4: rt_tile<4>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      case 5: rt_tile<5>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      case 6: rt_tile<6>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      case 7: rt_tile<7>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      case 8: rt_tile<8>(obs, qpd, nbq, ns, lane, mw, nw, T, s_planes, tally); break;
      default: break;  // plan_create never routes other widths here
    }
  }
}

bool sliced_rt_supports(uint32_t nb_obs) { return nb_obs >= 1 && nb_obs <= 8; }

int launch_tally_sliced_rt(const qcut_plan* plan, const uint8_t* d_data, const qcut_pub_desc* d_pubs,
                           const TileEntry* d_tiles, int64_t n_tiles, int64_t* d_tallies, cudaStream_t stream) {
  if (n_tiles <= 0) return QCUT_OK;
  static bool configured_dev[64] = {};
  int dev__ = 0;
  QCUT_CUDA(cudaGetDevice(&dev__));
  bool& configured = configured_dev[dev__ & 63];  // the attribute is per device
  const size_t smem = (size_t)kRtWarps * kRtPlaneWords * sizeof(uint32_t);  // 64 KB per CTA
  if (!configured) {
    QCUT_CUDA(cudaFuncSetAttribute(tally_sliced_rt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int64_t want = (n_tiles + kRtWarps - 1) / kRtWarps;
  const int64_t cap = (int64_t)sm_count() * 2;
  const int grid = (int)(want < cap ? want : cap);
  tally_sliced_rt_kernel<<<grid, kRtWarps * 32, smem, stream>>>(
      d_data, d_pubs, d_tiles, n_tiles, plan->d_groups, plan->d_mask_words,
      reinterpret_cast<unsigned long long*>(d_tallies));
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

}  // namespace qcut

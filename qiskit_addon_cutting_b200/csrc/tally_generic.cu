// K1 generic: parity tallies for any row width / any number of terms.
//
// Replaces the reference's per-shot loop (cutting_reconstruction.py:156-161) and sign rule
// (_process_outcome_v2, :196-222).  One warp owns one tile of QCUT_TILE_SHOTS shots of one PUB;
// each lane owns one shot per step, keeps the whole observable row in registers, and every term
// costs one AND/XOR per 32-bit word, one POPC and one warp vote.  This is the always-available
// path (rows wider than the bit-sliced kernels cover) and the on-device golden reference for
// them (QCUT_PLAN_GENERIC_ONLY).  It is POPC-bound for many-term groups.
#include "common.cuh"

namespace qcut {

constexpr int kWarpsPerCta = 8;
constexpr int kTermChunk = 256;                    // per-warp smem accumulators per pass
constexpr int kMaxWords = QCUT_MAX_OBS_BYTES / 4;  // 16 words = 64 bytes per row

__global__ void __launch_bounds__(kWarpsPerCta * 32)
tally_generic_kernel(const uint8_t* __restrict__ data, const qcut_pub_desc* __restrict__ pubs,
                     const TileEntry* __restrict__ tiles, int64_t n_tiles,
                     const GroupDev* __restrict__ groups, const uint32_t* __restrict__ mask_words,
                     unsigned long long* __restrict__ tallies) {
  __shared__ int s_odd[kWarpsPerCta][kTermChunk];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;

  for (int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + warp; tile < n_tiles; tile += n_warps) {
    const TileEntry te = tiles[tile];
    const qcut_pub_desc pub = pubs[te.pub];
    const GroupDev grp = groups[pub.group];

    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;
    const int ns = (int)min((int64_t)QCUT_TILE_SHOTS, (int64_t)pub.shots - shot0);
    const int nb = (int)grp.nb_obs;
    const int nbq = (int)pub.nb_qpd;
    const int nw = (int)grp.n_words;
    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * nb;
    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * nbq;
    const uint32_t* __restrict__ mw = mask_words + grp.word_offset;

    for (int t0 = 0; t0 < (int)grp.n_terms; t0 += kTermChunk) {
      const int tc = min(kTermChunk, (int)grp.n_terms - t0);
      for (int t = lane; t < tc; t += 32) s_odd[warp][t] = 0;
      __syncwarp();

      for (int base = 0; base < ns; base += 32) {
        const int s = base + lane;
        const bool valid = s < ns;
        // Row -> registers, assembled in memory order (byte b of the row is bits 8*(b%4).. of word b/4).
        uint32_t row[kMaxWords];
#pragma unroll
        for (int w = 0; w < kMaxWords; ++w) {
          uint32_t v = 0;
          if (valid && w < nw) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int b = 4 * w + j;
              if (b < nb) v |= (uint32_t)__ldg(obs + (size_t)s * nb + b) << (8 * j);
            }
          }
          row[w] = v;
        }
        // Parity of ALL bits of the qpd row, pad bits included (reference :213).
        uint32_t q = 0;
        if (valid)
          for (int b = 0; b < nbq; ++b) q ^= __ldg(qpd + (size_t)s * nbq + b);
        const uint32_t qpar = __popc(q) & 1u;

        for (int t = 0; t < tc; ++t) {
          const uint32_t* __restrict__ m = mw + (size_t)(t0 + t) * nw;
          uint32_t acc = 0;
#pragma unroll
          for (int w = 0; w < kMaxWords; ++w)
            if (w < nw) acc ^= row[w] & __ldg(m + w);
          const uint32_t odd = (__popc(acc) & 1u) ^ qpar;
          const unsigned votes = __ballot_sync(0xffffffffu, valid && odd);
          if (lane == 0) s_odd[warp][t] += __popc(votes);
        }
      }
      __syncwarp();
      // tally = (#even) - (#odd) = ns - 2 * odd; integer atomics are exact and order independent.
      for (int t = lane; t < tc; t += 32) {
        const long long v = (long long)ns - 2ll * s_odd[warp][t];
        atomicAdd(tallies + pub.tally_offset + t0 + t, (unsigned long long)v);
      }
      __syncwarp();
    }
  }
}

int launch_tally_generic(const qcut_plan* plan, const uint8_t* d_data, const qcut_pub_desc* d_pubs,
                         const TileEntry* d_tiles, int64_t n_tiles, int64_t* d_tallies, cudaStream_t stream) {
  if (n_tiles <= 0) return QCUT_OK;
  const int64_t want = (n_tiles + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t cap = (int64_t)sm_count() * 8;  // 8 CTAs x 8 warps resident per SM
  const int grid = (int)(want < cap ? want : cap);
  tally_generic_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(
      d_data, d_pubs, d_tiles, n_tiles, plan->d_groups, plan->d_mask_words,
      reinterpret_cast<unsigned long long*>(d_tallies));
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

}  // namespace qcut

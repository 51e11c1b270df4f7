// K1-literal: per-(PUB, term) expectation values with the REFERENCE'S fp64 rounding, bit for bit.
//
// The reference accumulates  subsystem_expvals[k] += (1 / shots) * sign  shot by shot in fp64
// (cutting_reconstruction.py:155-161 with _process_outcome_v2, :196-222): the result is
//   fl(fl(...fl(0 +- h) +- h ...) +- h),  h = fl(1 / shots),
// which for shot counts that are not powers of two differs from tally / shots by up to ~6e-12 at 10^6 shots
// and depends on the ORDER of the signs.  The default path (exact int64 tallies, one correctly rounded division)
// is more accurate but cannot reproduce that; this kernel can: the chain of one (PUB, term) is walked in shot
// order by one thread with __dadd_rn (no contraction, no reassociation).  Parallelism is over (PUB, term) only
// -- a CTA owns one PUB and up to 128 of its terms -- so it is an opt-in mode (rounding="reference"), about an
// order of magnitude slower than the bit-sliced tally kernels, yet every chain still runs at register speed:
// rows are staged once per CTA into shared memory (coalesced byte loads, rows padded to 32-bit words, the qpd
// parity folded into one flag per shot) and all threads of a warp read the same row (broadcast).
#include "common.cuh"

namespace qcut {

constexpr int kLitThreads = 128;            // terms per CTA
constexpr int kLitRowBytes = 32 * 1024;     // staged observable rows per pass
constexpr int kLitMaxShots = 4096;          // shots per pass (also bounds the qpd flag array)

template <int NW>
__device__ __forceinline__ double literal_chain(const uint32_t* __restrict__ rows, const uint8_t* __restrict__ qodd,
                                                const uint32_t* __restrict__ m, int nw, int n, double h, double acc) {
  const double minus_h = -h;
  for (int s = 0; s < n; ++s) {
    uint32_t p = 0;
    if constexpr (NW > 0) {
#pragma unroll
      for (int w = 0; w < NW; ++w) p ^= rows[s * NW + w] & m[w];
    } else {
      for (int w = 0; w < nw; ++w) p ^= rows[s * nw + w] & m[w];
    }
    const uint32_t odd = (__popc(p) ^ qodd[s]) & 1u;
    acc = __dadd_rn(acc, odd ? minus_h : h);  // strictly in shot order
  }
  return acc;
}

__global__ void __launch_bounds__(kLitThreads)
literal_expvals_kernel(const uint8_t* __restrict__ data, const qcut_pub_desc* __restrict__ pubs, int64_t n_pubs,
                       const GroupDev* __restrict__ groups, const uint32_t* __restrict__ mask_words,
                       double* __restrict__ expvals) {
  __shared__ uint32_t s_rows[kLitRowBytes / 4];
  __shared__ uint8_t s_qodd[kLitMaxShots];
  uint8_t* s_bytes = reinterpret_cast<uint8_t*>(s_rows);
  constexpr int kMaxWords = QCUT_MAX_OBS_BYTES / 4;

  for (int64_t p = blockIdx.x; p < n_pubs; p += gridDim.x) {
    const qcut_pub_desc pub = pubs[p];
    const GroupDev grp = groups[pub.group];
    const int t0 = blockIdx.y * kLitThreads;
    if (t0 >= (int)grp.n_terms) continue;  // uniform per CTA
    const int t = t0 + threadIdx.x;
    const bool active = t < (int)grp.n_terms;
    const int nb = (int)grp.nb_obs, nw = (int)grp.n_words, nbq = (int)pub.nb_qpd;
    const int row_bytes = nw * 4;
    const int pass_shots = min(kLitMaxShots, kLitRowBytes / row_bytes);
    uint32_t m[kMaxWords];
#pragma unroll
    for (int w = 0; w < kMaxWords; ++w)
      m[w] = (active && w < nw) ? __ldg(mask_words + grp.word_offset + (size_t)t * nw + w) : 0u;
    const double h = __ddiv_rn(1.0, (double)pub.shots);  // 1 / shots, correctly rounded like Python's
    double acc = 0.0;
    const uint8_t* __restrict__ obs = data + pub.obs_offset;
    const uint8_t* __restrict__ qpd = data + pub.qpd_offset;
    for (int64_t s0 = 0; s0 < (int64_t)pub.shots; s0 += pass_shots) {
      const int n = (int)min((int64_t)pass_shots, (int64_t)pub.shots - s0);
      // rows -> shared memory, each padded with zero bytes to nw words, in memory order (byte b of a row is bits
      // 8*(b%4).. of word b/4, the layout of mask_words)
      for (int i = threadIdx.x; i < n * row_bytes; i += kLitThreads) {
        const int s = i / row_bytes, b = i - s * row_bytes;
        s_bytes[i] = b < nb ? __ldg(obs + (size_t)(s0 + s) * nb + b) : (uint8_t)0;
      }
      // parity of ALL bits of the qpd row, pad bits included (reference :213)
      for (int s = threadIdx.x; s < n; s += kLitThreads) {
        uint32_t q = 0;
        for (int b = 0; b < nbq; ++b) q ^= __ldg(qpd + (size_t)(s0 + s) * nbq + b);
        s_qodd[s] = (uint8_t)(__popc(q) & 1);
      }
      __syncthreads();
      if (active) {
        switch (nw) {
          case 1: acc = literal_chain<1>(s_rows, s_qodd, m, nw, n, h, acc); break;
          case 2: acc = literal_chain<2>(s_rows, s_qodd, m, nw, n, h, acc); break;
          case 4: acc = literal_chain<4>(s_rows, s_qodd, m, nw, n, h, acc); break;
          default: acc = literal_chain<0>(s_rows, s_qodd, m, nw, n, h, acc); break;
        }
      }
      __syncthreads();
    }
    if (active) expvals[pub.tally_offset + t] = acc;
  }
}

int launch_literal(const qcut_plan* plan, const uint8_t* d_data, const qcut_pub_desc* d_pubs, int64_t n_pubs,
                   double* d_expvals, cudaStream_t stream) {
  if (n_pubs <= 0) return QCUT_OK;
  uint32_t max_terms = 1;
  for (const qcut_group_desc& g : plan->groups) max_terms = g.n_terms > max_terms ? g.n_terms : max_terms;
  const unsigned chunks = (max_terms + kLitThreads - 1) / kLitThreads;
  const int64_t cap = (int64_t)sm_count() * 32;  // strided over PUBs beyond that
  const dim3 grid((unsigned)(n_pubs < cap ? n_pubs : cap), chunks);
  literal_expvals_kernel<<<grid, kLitThreads, 0, stream>>>(d_data, d_pubs, n_pubs, plan->d_groups, plan->d_mask_words,
                                                           d_expvals);
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

}  // namespace qcut

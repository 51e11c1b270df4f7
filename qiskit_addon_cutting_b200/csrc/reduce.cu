// K2: coefficient-weighted product across subsystems, summed over subexperiment tuples.
//
// Reference (cutting_reconstruction.py:134-136,163-168):
//     for i, coeff in enumerate(coefficients):
//         current = ones(K)
//         for label: current[k] *= mean(E[label, i, g, t] for (g, t) in lookup[label][k])
//         expvals += coeff[0] * current
//
// fp64 gather-product + segmented sum; no tensor cores (this is not a dense contraction).
// A CTA owns kReduceChunk consecutive coefficient tuples x kReduceObs observables:
//   phase 1 (all threads, latency hidden by parallelism): term[k][i] = coeff[i] * prod_label mean(...)
//            -> shared memory;
//   phase 2 (one thread per observable): sums its row in strictly ascending i, separate multiply and
//            add, no FMA contraction -- the reference's order and rounding.
// Chunks are combined in ascending order by a second launch.  For n_coeffs <= kReduceChunk the
// summation order is therefore *identical* to the reference's; in every case it is deterministic.
#include <type_traits>

#include "common.cuh"

namespace qcut {

constexpr int kReduceChunk = 1024;
constexpr int kReduceObs = 8;
constexpr int kReduceThreads = 256;
constexpr int kReduceRow = kReduceChunk + 2;  // padded row (doubles) so the 8 rows start in different banks

template <typename TallyT>
__device__ __forceinline__ double pub_expval(const TallyT* __restrict__ tallies,
                                             const qcut_pub_desc* __restrict__ pub, uint32_t term);

// V2: exact int64 tally -> one correctly rounded division.  A PUB with zero shots contributes
// 0.0, as the reference's empty shot loop does (cutting_reconstruction.py:155-161).
template <>
__device__ __forceinline__ double pub_expval<int64_t>(const int64_t* __restrict__ tallies,
                                                      const qcut_pub_desc* __restrict__ pub, uint32_t term) {
  const uint32_t shots = __ldg(&pub->shots);
  if (shots == 0) return 0.0;
  return __ddiv_rn((double)__ldg(tallies + __ldg(&pub->tally_offset) + term), (double)shots);
}

// V1: the quasi-probability kernel already produced fp64 expectation values.
template <>
__device__ __forceinline__ double pub_expval<double>(const double* __restrict__ expvals,
                                                     const qcut_pub_desc* __restrict__ pub, uint32_t term) {
  return __ldg(expvals + __ldg(&pub->tally_offset) + term);
}

template <typename TallyT>
__global__ void __launch_bounds__(kReduceThreads)
reduce_partial_kernel(const TallyT* __restrict__ tallies, const qcut_pub_desc* __restrict__ pubs,
                      const qcut_label_desc* __restrict__ labels, int n_labels,
                      const uint32_t* __restrict__ lookup_start, const qcut_slot* __restrict__ lookup,
                      const double* __restrict__ coeffs, int64_t n_coeffs, int n_obs,
                      double* __restrict__ partial) {
  extern __shared__ double s_term[];  // [kReduceObs][kReduceRow]
  const int k_local = threadIdx.x % kReduceObs;
  const int i_local = threadIdx.x / kReduceObs;
  const int k = blockIdx.x * kReduceObs + k_local;
  const int64_t i0 = (int64_t)blockIdx.y * kReduceChunk;
  const int n_i = (int)min((int64_t)kReduceChunk, n_coeffs - i0);

  if (k < n_obs) {
    for (int ii = i_local; ii < n_i; ii += kReduceThreads / kReduceObs) {
      const int64_t i = i0 + ii;
      double prod = 1.0;
      for (int l = 0; l < n_labels; ++l) {
        const qcut_label_desc lab = labels[l];
        const uint32_t e0 = __ldg(lookup_start + lab.lookup_base + k);
        const uint32_t e1 = __ldg(lookup_start + lab.lookup_base + k + 1);
        double acc = 0.0;
        for (uint32_t e = e0; e < e1; ++e) {
          const qcut_slot slot = lookup[e];
          acc = __dadd_rn(acc, pub_expval<TallyT>(tallies, pubs + lab.pub_base + (uint64_t)i * lab.num_groups + slot.group, slot.term));
        }
        // np.mean of the slot list: sum / n (exactly the value itself for the usual n == 1).
        prod = __dmul_rn(prod, __ddiv_rn(acc, (double)(e1 - e0)));
      }
      s_term[k_local * kReduceRow + ii] = __dmul_rn(__ldg(coeffs + i), prod);
    }
  }
  __syncthreads();
  if (threadIdx.x < kReduceObs) {
    const int kk = blockIdx.x * kReduceObs + threadIdx.x;
    if (kk < n_obs) {
      const double* __restrict__ row = s_term + threadIdx.x * kReduceRow;
      double sum = 0.0;
      for (int ii = 0; ii < n_i; ++ii) sum = __dadd_rn(sum, row[ii]);
      partial[(int64_t)blockIdx.y * n_obs + kk] = sum;
    }
  }
}

__global__ void __launch_bounds__(128)
reduce_combine_kernel(const double* __restrict__ partial, int64_t n_chunks, int n_obs,
                      double* __restrict__ out) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= n_obs) return;
  double sum = 0.0;
  for (int64_t c = 0; c < n_chunks; ++c) sum = __dadd_rn(sum, partial[c * n_obs + k]);
  out[k] = sum;
}

static int64_t num_chunks(int64_t n_coeffs) {
  return n_coeffs <= 0 ? 1 : (n_coeffs + kReduceChunk - 1) / kReduceChunk;
}

template <typename TallyT>
int launch_reduce(const TallyT* d_tallies, const qcut_pub_desc* d_pubs, const qcut_label_desc* d_labels,
                  int n_labels, const uint32_t* d_lookup_start, const qcut_slot* d_lookup,
                  const double* d_coeffs, int64_t n_coeffs, int n_obs, double* d_out,
                  void* d_workspace, cudaStream_t stream) {
  if (n_obs <= 0) return QCUT_OK;
  const int64_t chunks = num_chunks(n_coeffs);
  if (chunks > 65535) {
    set_error("qcut_reduce_launch: more than 65535 * 1024 coefficients per rank are not supported");
    return QCUT_ERR_INVALID;
  }
  static bool configured[2] = {false, false};
  const size_t smem = sizeof(double) * kReduceObs * kReduceRow;  // 64 KB + padding
  const int which = sizeof(TallyT) == sizeof(int64_t) && std::is_same<TallyT, int64_t>::value ? 0 : 1;
  if (!configured[which]) {
    QCUT_CUDA(cudaFuncSetAttribute(reduce_partial_kernel<TallyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[which] = true;
  }
  double* partial = static_cast<double*>(d_workspace);
  const dim3 grid((n_obs + kReduceObs - 1) / kReduceObs, (unsigned)chunks);
  reduce_partial_kernel<TallyT><<<grid, kReduceThreads, smem, stream>>>(
      d_tallies, d_pubs, d_labels, n_labels, d_lookup_start, d_lookup, d_coeffs, n_coeffs, n_obs, partial);
  QCUT_CUDA(cudaGetLastError());
  reduce_combine_kernel<<<(n_obs + 127) / 128, 128, 0, stream>>>(partial, chunks, n_obs, d_out);
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

template int launch_reduce<int64_t>(const int64_t*, const qcut_pub_desc*, const qcut_label_desc*, int,
                                    const uint32_t*, const qcut_slot*, const double*, int64_t, int,
                                    double*, void*, cudaStream_t);
template int launch_reduce<double>(const double*, const qcut_pub_desc*, const qcut_label_desc*, int,
                                   const uint32_t*, const qcut_slot*, const double*, int64_t, int,
                                   double*, void*, cudaStream_t);

size_t reduce_workspace_bytes(int64_t n_coeffs, int n_obs) {
  return (size_t)num_chunks(n_coeffs) * (size_t)(n_obs > 0 ? n_obs : 1) * sizeof(double);
}

// ---------------------------------------------------------------------------------------------
// SamplerV1 quasi-distributions (reference cutting_reconstruction.py:143-149, _process_outcome
// :173-193).  One thread per (PUB, term) walks the PUB's outcomes in the order given and
// accumulates weight * sign sequentially in fp64, so the result is bit-identical to the
// reference's dict-order loop.  Data is tiny (<= 2^bits outcomes per PUB).
// ---------------------------------------------------------------------------------------------
__global__ void quasi_kernel(const qcut_pub_desc* __restrict__ pubs, int64_t n_pubs,
                             const qcut_group_desc* __restrict__ groups,
                             const uint64_t* __restrict__ mask_limbs,
                             const int64_t* __restrict__ dist_start,
                             const uint64_t* __restrict__ obs_bits, int n_limbs_obs,
                             const uint64_t* __restrict__ qpd_bits, int n_limbs_qpd,
                             const double* __restrict__ weights, double* __restrict__ expvals) {
  const int64_t p = blockIdx.x;
  if (p >= n_pubs) return;
  const qcut_pub_desc pub = pubs[p];
  const qcut_group_desc grp = groups[pub.group];
  const int64_t o0 = dist_start[p], o1 = dist_start[p + 1];
  for (uint32_t t = threadIdx.x; t < grp.n_terms; t += blockDim.x) {
    const uint64_t* __restrict__ m = mask_limbs + grp.mask_offset + (size_t)t * n_limbs_obs;
    double acc = 0.0;
    for (int64_t o = o0; o < o1; ++o) {
      int par = 0;
      for (int j = 0; j < n_limbs_qpd; ++j) par += __popcll(qpd_bits[o * n_limbs_qpd + j]);
      for (int j = 0; j < n_limbs_obs; ++j) par += __popcll(obs_bits[o * n_limbs_obs + j] & m[j]);
      const double w = weights[o];
      acc = __dadd_rn(acc, (par & 1) ? -w : w);
    }
    expvals[pub.tally_offset + t] = acc;
  }
}

int launch_quasi(const qcut_pub_desc* d_pubs, int64_t n_pubs, const qcut_group_desc* d_groups,
                 const uint64_t* d_mask_limbs, const int64_t* d_dist_start, const uint64_t* d_obs_bits,
                 int n_limbs_obs, const uint64_t* d_qpd_bits, int n_limbs_qpd, const double* d_weights,
                 double* d_expvals, cudaStream_t stream) {
  if (n_pubs <= 0) return QCUT_OK;
  if (n_pubs > 2147483647ll) {
    set_error("qcut_quasi_launch: too many PUBs");
    return QCUT_ERR_INVALID;
  }
  quasi_kernel<<<(unsigned)n_pubs, 64, 0, stream>>>(d_pubs, n_pubs, d_groups, d_mask_limbs,
                                                    d_dist_start, d_obs_bits, n_limbs_obs, d_qpd_bits,
                                                    n_limbs_qpd, d_weights, d_expvals);
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

}  // namespace qcut

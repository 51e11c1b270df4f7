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
// Chunks (1024 tuples when the whole problem fits one chunk, else 256) are combined in ascending
// order by a second launch.  For n_coeffs <= kReduceChunk the
// summation order is therefore *identical* to the reference's; in every case it is deterministic.
#include <type_traits>

#include "common.cuh"

namespace qcut {

constexpr int kReduceChunk = 1024;
constexpr int kReduceObs = 8;
constexpr int kReduceThreads = 256;

template <typename TallyT>
__device__ __forceinline__ double pub_expval(const TallyT* __restrict__ tallies,
                                             const qcut_pub_desc* __restrict__ pub, uint32_t term);

// The two halves of a PUB descriptor that K2 needs, fetched with two independent 8-byte loads
// (no branch between them: the dependent tally load can be issued as soon as the offset arrives).
struct PubRef {
  uint64_t tally_offset;
  uint32_t shots;
};
__device__ __forceinline__ PubRef load_pub_ref(const qcut_pub_desc* __restrict__ pub) {
  const unsigned long long off = __ldg(reinterpret_cast<const unsigned long long*>(&pub->tally_offset));
  const uint2 sg = __ldg(reinterpret_cast<const uint2*>(&pub->shots));  // {shots, group}
  return PubRef{off, sg.x};
}

// V2: exact int64 tally -> one correctly rounded division.  A PUB with zero shots contributes
// 0.0, as the reference's empty shot loop does (cutting_reconstruction.py:155-161).
template <>
__device__ __forceinline__ double pub_expval<int64_t>(const int64_t* __restrict__ tallies,
                                                      const qcut_pub_desc* __restrict__ pub, uint32_t term) {
  const PubRef ref = load_pub_ref(pub);
  const double tally = (double)__ldg(tallies + ref.tally_offset + term);
  const double e = __ddiv_rn(tally, (double)(ref.shots ? ref.shots : 1u));
  return ref.shots ? e : 0.0;
}

// V1: the quasi-probability kernel already produced fp64 expectation values.
template <>
__device__ __forceinline__ double pub_expval<double>(const double* __restrict__ expvals,
                                                     const qcut_pub_desc* __restrict__ pub, uint32_t term) {
  return __ldg(expvals + __ldg(reinterpret_cast<const unsigned long long*>(&pub->tally_offset)) + term);
}

template <typename TallyT>
__device__ __forceinline__ double expval_from(TallyT raw, uint32_t shots);
template <>
__device__ __forceinline__ double expval_from<int64_t>(int64_t raw, uint32_t shots) {
  const double e = __ddiv_rn((double)raw, (double)(shots ? shots : 1u));
  return shots ? e : 0.0;
}
template <>
__device__ __forceinline__ double expval_from<double>(double raw, uint32_t) { return raw; }

// One hoisted (subsystem, observable) slot: where to find E[label, i, .] for this thread's observable.
struct SlotRef {
  uint64_t pub_base;
  uint32_t n_groups, group, term;
};

// term[j] = coeff[i_j] * prod_l E[l, i_j] for B tuples at once.  Straight-line code (L and B are
// compile-time): all B x L descriptor loads are issued first, then all B x L tally loads, then the
// arithmetic -- the two-level load chains overlap instead of running one after the other.
template <typename TallyT, int L, int B>
__device__ __forceinline__ void batch_terms(const TallyT* __restrict__ tallies, const qcut_pub_desc* __restrict__ pubs,
                                            const SlotRef* slots, const double* __restrict__ coeffs,
                                            const int64_t* idx, double* term) {
  PubRef ref[B][L];
#pragma unroll
  for (int j = 0; j < B; ++j)
#pragma unroll
    for (int l = 0; l < L; ++l)
      ref[j][l] = load_pub_ref(pubs + slots[l].pub_base + (uint64_t)idx[j] * slots[l].n_groups + slots[l].group);
  TallyT raw[B][L];
  double c[B];
#pragma unroll
  for (int j = 0; j < B; ++j) {
    c[j] = __ldg(coeffs + idx[j]);
#pragma unroll
    for (int l = 0; l < L; ++l) raw[j][l] = __ldg(tallies + ref[j][l].tally_offset + slots[l].term);
  }
#pragma unroll
  for (int j = 0; j < B; ++j) {
    double prod = 1.0;
#pragma unroll
    for (int l = 0; l < L; ++l) prod = __dmul_rn(prod, expval_from<TallyT>(raw[j][l], ref[j][l].shots));  // mean of one slot = itself
    term[j] = __dmul_rn(c[j], prod);
  }
}

// Hoist the slots of observable k; false if some subsystem lists more than one slot for it.
template <int L>
__device__ __forceinline__ bool load_slots(const qcut_label_desc* __restrict__ labels,
                                           const uint32_t* __restrict__ lookup_start,
                                           const qcut_slot* __restrict__ lookup, int k, SlotRef* slots) {
  bool ok = true;
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const qcut_label_desc lab = labels[l];
    const uint32_t e0 = __ldg(lookup_start + lab.lookup_base + k);
    const uint32_t e1 = __ldg(lookup_start + lab.lookup_base + k + 1);
    const qcut_slot slot = lookup[e0];
    slots[l] = SlotRef{lab.pub_base, lab.num_groups, slot.group, slot.term};
    ok = ok && (e1 - e0 == 1);
  }
  return ok;
}

// General (slow) form: any number of subsystems, any number of slots per observable.
template <typename TallyT>
__device__ __forceinline__ double general_term(const TallyT* __restrict__ tallies, const qcut_pub_desc* __restrict__ pubs,
                                               const qcut_label_desc* __restrict__ labels, int n_labels,
                                               const uint32_t* __restrict__ lookup_start,
                                               const qcut_slot* __restrict__ lookup, const double* __restrict__ coeffs,
                                               int64_t i, int k) {
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
    prod = __dmul_rn(prod, __ddiv_rn(acc, (double)(e1 - e0)));  // np.mean of the slot list: sum / n
  }
  return __dmul_rn(__ldg(coeffs + i), prod);
}

constexpr int kMaxFastLabels = 4;  // subsystems handled by the straight-line path (more: general path)

// CHUNK coefficient tuples per CTA (one-chunk problems: reference order end to end).
// L = number of subsystems when every observable has exactly one slot per subsystem (1..4), 0 = general.
// SEQ: one CTA walks ALL tuples, chunk after chunk, and its summing threads carry their running sums across the
// chunks -- the reference's order for any number of tuples (rounding="reference"); gridDim.y must be 1.
template <typename TallyT, int CHUNK, int L, bool SEQ>
__global__ void __launch_bounds__(kReduceThreads)
reduce_partial_kernel(const TallyT* __restrict__ tallies, const qcut_pub_desc* __restrict__ pubs,
                      const qcut_label_desc* __restrict__ labels, int n_labels,
                      const uint32_t* __restrict__ lookup_start, const qcut_slot* __restrict__ lookup,
                      const double* __restrict__ coeffs, int64_t n_coeffs, int n_obs,
                      double* __restrict__ partial) {
  constexpr int ROW = CHUNK + 10;  // padded row (doubles): rows start in different banks, room for the 8-wide tail
  constexpr int B = (L <= 2) ? 8 : 4;
  constexpr int kStride = kReduceThreads / kReduceObs;  // 32 tuples per pass
  extern __shared__ double s_term[];  // [kReduceObs][ROW]
  const int k_local = threadIdx.x % kReduceObs;
  const int i_local = threadIdx.x / kReduceObs;
  const int k = blockIdx.x * kReduceObs + k_local;
  double sum = 0.0;  // running sum of the summing threads (threadIdx.x < kReduceObs)
  const int64_t i_begin = SEQ ? 0 : (int64_t)blockIdx.y * CHUNK;
  const int64_t i_end = SEQ ? n_coeffs : min(n_coeffs, i_begin + CHUNK);
  for (int64_t i0 = i_begin; i0 < i_end || i0 == i_begin; i0 += CHUNK) {
    const int n_i = (int)max((int64_t)0, min((int64_t)CHUNK, n_coeffs - i0));
    if (k < n_obs) {
      if constexpr (L > 0) {
        SlotRef slots[L];
        load_slots<L>(labels, lookup_start, lookup, k, slots);
        for (int base = i_local; base < n_i; base += B * kStride) {
          int64_t idx[B];
          double term[B];
#pragma unroll
          for (int j = 0; j < B; ++j) idx[j] = i0 + min(base + j * kStride, n_i - 1);  // clamped: always valid
          batch_terms<TallyT, L, B>(tallies, pubs, slots, coeffs, idx, term);
#pragma unroll
          for (int j = 0; j < B; ++j)
            if (base + j * kStride < n_i) s_term[k_local * ROW + base + j * kStride] = term[j];
        }
      } else {
        for (int ii = i_local; ii < n_i; ii += kStride)
          s_term[k_local * ROW + ii] = general_term<TallyT>(tallies, pubs, labels, n_labels, lookup_start, lookup, coeffs, i0 + ii, k);
      }
    }
    // zero the row tails up to a multiple of 8 so that phase 2 can run 8-wide without bound checks
    // (adding +0.0 never changes a sum that started at +0.0)
    const int n_pad = (n_i + 7) & ~7;
    if (threadIdx.x < kReduceObs * 8) {
      const int ii = n_i + threadIdx.x / kReduceObs;
      if (ii < n_pad) s_term[(threadIdx.x % kReduceObs) * ROW + ii] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x < kReduceObs && blockIdx.x * kReduceObs + threadIdx.x < n_obs) {
      const double* __restrict__ row = s_term + threadIdx.x * ROW;
      for (int ii = 0; ii < n_pad; ii += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = row[ii + j];
#pragma unroll
        for (int j = 0; j < 8; ++j) sum = __dadd_rn(sum, v[j]);  // strictly ascending i
      }
    }
    if (SEQ) __syncthreads();  // the rows are overwritten by the next chunk
  }
  if (threadIdx.x < kReduceObs) {
    const int kk = blockIdx.x * kReduceObs + threadIdx.x;
    if (kk < n_obs) partial[(int64_t)blockIdx.y * n_obs + kk] = sum;
  }
}

// Large problems (more than kReduceChunk tuples): lane <-> observable, one warp walks kStreamChunk
// consecutive tuples and every lane keeps its running sum in a register (ascending i inside the
// chunk).  No shared memory; the descriptor -> tally load chains of B tuples x L subsystems are in
// flight together per lane and up to 64 warps per SM hide the rest of the latency.
// Measured on cfg4 (97 000 tuples x 271 observables, step minus K1): 8 tuples in flight per lane at 114 registers
// 0.161 ms; 4 tuples under a 64-register cap (8 CTAs of 4 warps per SM) 0.134 ms; 2 tuples at 42 registers 0.143 ms.
// The kernel waits on its descriptor -> tally load chains and on fp64 divisions: resident warps hide more of that
// than loads in flight per lane.  (64 tuples per warp keeps the summation order, hence every bit of the result.)
#ifndef QCUT_K2_B       // tuples whose load chains a lane keeps in flight (experiments: -DQCUT_K2_B=..)
#define QCUT_K2_B 4
#endif
#ifndef QCUT_K2_MINB    // CTAs per SM the stream kernel is compiled for (register cap)
#define QCUT_K2_MINB 8
#endif
#ifndef QCUT_K2_SUB
#define QCUT_K2_SUB 64
#endif
constexpr int kStreamWarps = 4;
constexpr int kStreamSub = QCUT_K2_SUB;                     // tuples per warp
constexpr int kStreamChunk = kStreamWarps * kStreamSub;    // tuples per CTA = one partial sum

template <typename TallyT, int L>
__global__ void __launch_bounds__(kStreamWarps * 32, (L == 1 || L == 2) ? QCUT_K2_MINB : 4)  // (3-4 subsystems spill under 64 registers)
reduce_stream_kernel(const TallyT* __restrict__ tallies, const qcut_pub_desc* __restrict__ pubs,
                     const qcut_label_desc* __restrict__ labels, int n_labels,
                     const uint32_t* __restrict__ lookup_start, const qcut_slot* __restrict__ lookup,
                     const double* __restrict__ coeffs, int64_t n_coeffs, int n_obs, int64_t n_chunks,
                     double* __restrict__ partial) {
  constexpr int B = (L <= 2) ? QCUT_K2_B : (QCUT_K2_B < 4 ? QCUT_K2_B : 4);
  __shared__ double s_sub[kStreamWarps][32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t chunk = blockIdx.y;
  const int k = blockIdx.x * 32 + lane;
  const int64_t i0 = chunk * kStreamChunk + (int64_t)warp * kStreamSub;
  const int n_i = (int)max((int64_t)0, min((int64_t)kStreamSub, n_coeffs - i0));
  double sum = 0.0;
  if (k < n_obs && n_i > 0) {
    if constexpr (L > 0) {
      SlotRef slots[L];
      load_slots<L>(labels, lookup_start, lookup, k, slots);
      for (int base = 0; base < n_i; base += B) {
        int64_t idx[B];
        double term[B];
#pragma unroll
        for (int j = 0; j < B; ++j) idx[j] = i0 + min(base + j, n_i - 1);  // clamped: always a valid tuple
        batch_terms<TallyT, L, B>(tallies, pubs, slots, coeffs, idx, term);
#pragma unroll
        for (int j = 0; j < B; ++j)
          if (base + j < n_i) sum = __dadd_rn(sum, term[j]);  // ascending i
      }
    } else {
      for (int ii = 0; ii < n_i; ++ii)
        sum = __dadd_rn(sum, general_term<TallyT>(tallies, pubs, labels, n_labels, lookup_start, lookup, coeffs, i0 + ii, k));
    }
  }
  s_sub[warp][lane] = sum;
  __syncthreads();
  if (warp == 0 && k < n_obs) {
    double total = s_sub[0][lane];
#pragma unroll
    for (int w = 1; w < kStreamWarps; ++w) total = __dadd_rn(total, s_sub[w][lane]);  // ascending sub-chunk order
    partial[chunk * n_obs + k] = total;
  }
}

// Second level for large problems: segment s sums kCombineSeg consecutive chunk partials in ascending
// order (thread <-> (segment, observable)); reduce_combine_kernel then sums the segments.
constexpr int kCombineSeg = 32;

__global__ void __launch_bounds__(128)
reduce_segment_kernel(const double* __restrict__ partial, int64_t n_chunks, int n_obs, double* __restrict__ seg_out) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= n_obs) return;
  const int64_t c_begin = (int64_t)blockIdx.y * kCombineSeg;
  double sum = 0.0;
  for (int64_t c0 = c_begin; c0 < c_begin + kCombineSeg && c0 < n_chunks; c0 += 8) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t c = c0 + j;
      v[j] = __ldg(partial + (c < n_chunks ? c : n_chunks - 1) * n_obs + k);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < n_chunks) sum = __dadd_rn(sum, v[j]);
  }
  seg_out[(int64_t)blockIdx.y * n_obs + k] = sum;
}

__global__ void __launch_bounds__(128)
reduce_combine_kernel(const double* __restrict__ partial, int64_t n_chunks, int n_obs,
                      double* __restrict__ out) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= n_obs) return;
  double sum = 0.0;
  for (int64_t c0 = 0; c0 < n_chunks; c0 += 8) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t c = c0 + j;
      v[j] = __ldg(partial + (c < n_chunks ? c : n_chunks - 1) * n_obs + k);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < n_chunks) sum = __dadd_rn(sum, v[j]);  // ascending chunk order
  }
  out[k] = sum;
}

static int chunk_size(int64_t n_coeffs) { return n_coeffs <= kReduceChunk ? kReduceChunk : kStreamChunk; }

static int64_t num_chunks(int64_t n_coeffs) {
  const int chunk = chunk_size(n_coeffs);
  return n_coeffs <= 0 ? 1 : (n_coeffs + chunk - 1) / chunk;
}

template <typename TallyT, int L>
static int launch_level(const TallyT* d_tallies, const qcut_pub_desc* d_pubs, const qcut_label_desc* d_labels,
                        int n_labels, const uint32_t* d_lookup_start, const qcut_slot* d_lookup,
                        const double* d_coeffs, int64_t n_coeffs, int n_obs, double* partial, int64_t chunks,
                        bool sequential, cudaStream_t stream) {
  if (sequential || chunk_size(n_coeffs) == kReduceChunk) {
    static bool configured_dev[64][2] = {};
    int dev__ = 0;
    QCUT_CUDA(cudaGetDevice(&dev__));
    const bool seq = sequential && n_coeffs > kReduceChunk;
    bool& configured = configured_dev[dev__ & 63][seq ? 1 : 0];  // the attribute is per device
    const size_t smem = sizeof(double) * kReduceObs * (kReduceChunk + 10);
    if (!configured) {
      if (seq)
        QCUT_CUDA(cudaFuncSetAttribute(reduce_partial_kernel<TallyT, kReduceChunk, L, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else
        QCUT_CUDA(cudaFuncSetAttribute(reduce_partial_kernel<TallyT, kReduceChunk, L, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    const dim3 grid((n_obs + kReduceObs - 1) / kReduceObs, (unsigned)chunks);
    if (seq)
      reduce_partial_kernel<TallyT, kReduceChunk, L, true><<<grid, kReduceThreads, smem, stream>>>(
          d_tallies, d_pubs, d_labels, n_labels, d_lookup_start, d_lookup, d_coeffs, n_coeffs, n_obs, partial);
    else
      reduce_partial_kernel<TallyT, kReduceChunk, L, false><<<grid, kReduceThreads, smem, stream>>>(
          d_tallies, d_pubs, d_labels, n_labels, d_lookup_start, d_lookup, d_coeffs, n_coeffs, n_obs, partial);
  } else {
    const dim3 grid((n_obs + 31) / 32, (unsigned)chunks);
    reduce_stream_kernel<TallyT, L><<<grid, kStreamWarps * 32, 0, stream>>>(
        d_tallies, d_pubs, d_labels, n_labels, d_lookup_start, d_lookup, d_coeffs, n_coeffs, n_obs, chunks, partial);
  }
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

// single_slot: host-side knowledge that every observable has exactly one (group, term) slot in every
// subsystem (what Qiskit's grouping produces; checked by the caller from the lookup table it uploads).
template <typename TallyT>
int launch_reduce(const TallyT* d_tallies, const qcut_pub_desc* d_pubs, const qcut_label_desc* d_labels,
                  int n_labels, const uint32_t* d_lookup_start, const qcut_slot* d_lookup,
                  const double* d_coeffs, int64_t n_coeffs, int n_obs, double* d_out,
                  void* d_workspace, bool single_slot, bool sequential, cudaStream_t stream) {
  if (n_obs <= 0) return QCUT_OK;
  // sequential: one CTA column walks all tuples end to end -> the reference's summation order for any n_coeffs
  const int64_t chunks = sequential ? 1 : num_chunks(n_coeffs);
  if (chunks > 65535) {
    set_error("qcut_reduce_launch: too many coefficients per rank");
    return QCUT_ERR_INVALID;
  }
  double* partial = static_cast<double*>(d_workspace);
  const int level = (single_slot && n_labels >= 1 && n_labels <= kMaxFastLabels) ? n_labels : 0;
  int rc;
#define QCUT_LEVEL(LV)                                                                                              \
  rc = launch_level<TallyT, LV>(d_tallies, d_pubs, d_labels, n_labels, d_lookup_start, d_lookup, d_coeffs, n_coeffs, \
                                n_obs, partial, chunks, sequential, stream)
  switch (level) {
    case 1: QCUT_LEVEL(1); break;
    case 2: QCUT_LEVEL(2); break;
    case 3: QCUT_LEVEL(3); break;
    case 4: QCUT_LEVEL(4); break;
    default: QCUT_LEVEL(0); break;
  }
#undef QCUT_LEVEL
  if (rc != QCUT_OK) return rc;
  if (chunks > kCombineSeg) {
    const int64_t segs = (chunks + kCombineSeg - 1) / kCombineSeg;
    double* seg_out = partial + chunks * (int64_t)n_obs;
    reduce_segment_kernel<<<dim3((n_obs + 127) / 128, (unsigned)segs), 128, 0, stream>>>(partial, chunks, n_obs, seg_out);
    QCUT_CUDA(cudaGetLastError());
    reduce_combine_kernel<<<(n_obs + 127) / 128, 128, 0, stream>>>(seg_out, segs, n_obs, d_out);
  } else {
    reduce_combine_kernel<<<(n_obs + 127) / 128, 128, 0, stream>>>(partial, chunks, n_obs, d_out);
  }
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

template int launch_reduce<int64_t>(const int64_t*, const qcut_pub_desc*, const qcut_label_desc*, int,
                                    const uint32_t*, const qcut_slot*, const double*, int64_t, int,
                                    double*, void*, bool, bool, cudaStream_t);
template int launch_reduce<double>(const double*, const qcut_pub_desc*, const qcut_label_desc*, int,
                                   const uint32_t*, const qcut_slot*, const double*, int64_t, int,
                                   double*, void*, bool, bool, cudaStream_t);

size_t reduce_workspace_bytes(int64_t n_coeffs, int n_obs) {
  const size_t chunks = (size_t)num_chunks(n_coeffs);
  const size_t segs = (chunks + kCombineSeg - 1) / kCombineSeg;
  return (chunks + segs) * (size_t)(n_obs > 0 ? n_obs : 1) * sizeof(double);
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

// ---- int64 tallies -> fp64 expectation values, E = tally / shots (one correctly rounded division, 0 for PUBs without
// shots): what K2 does on the fly, as a table for calls that mix SamplerV1 and SamplerV2 subsystems ----
__global__ void __launch_bounds__(128)
tallies_to_expvals_kernel(const qcut_pub_desc* __restrict__ pubs, int64_t n_pubs, const GroupDev* __restrict__ groups,
                          const int64_t* __restrict__ tallies, double* __restrict__ expvals) {
  for (int64_t p = blockIdx.x; p < n_pubs; p += gridDim.x) {
    const qcut_pub_desc pub = pubs[p];
    const uint32_t T = groups[pub.group].n_terms;
    for (uint32_t t = threadIdx.x; t < T; t += blockDim.x)
      expvals[pub.tally_offset + t] = pub.shots ? __ddiv_rn((double)tallies[pub.tally_offset + t], (double)pub.shots) : 0.0;
  }
}

int launch_tallies_to_expvals(const qcut_plan* plan, const qcut_pub_desc* d_pubs, int64_t n_pubs, const int64_t* d_tallies,
                              double* d_expvals, cudaStream_t stream) {
  if (n_pubs <= 0) return QCUT_OK;
  const int64_t cap = (int64_t)sm_count() * 16;
  tallies_to_expvals_kernel<<<(unsigned)(n_pubs < cap ? n_pubs : cap), 128, 0, stream>>>(d_pubs, n_pubs, plan->d_groups,
                                                                                          d_tallies, d_expvals);
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

// ---- term recombination: out[r] = sum_j coeff[j] * e[index[j]] (complex fp64, reference order) ----
// One warp per observable: the lanes form 128 products at a time (coalesced coefficient loads, gathered
// expectation values, four independent loads in flight per lane) and park them in shared memory; lane 0
// then adds them in ascending j -- the reference's order and rounding (utils/observable_terms.py:66-79:
// `rv = 0.0j; rv += coeff * term_expval`).  The sum is a dependent chain by definition, so a row runs at
// DADD latency; rows run side by side on the other warps.  Complex product as NumPy's scalar multiply:
// (ar*br - ai*bi, ar*bi + ai*br), no contraction.
constexpr int kRecombineWarps = 4;
constexpr int kRecombineChunk = 128;

__global__ void __launch_bounds__(32 * kRecombineWarps)
    recombine_kernel(const double* __restrict__ e, int e_complex, const uint32_t* __restrict__ row_start,
                     const uint32_t* __restrict__ term_index, const double2* __restrict__ coeff, int n_rows,
                     double2* __restrict__ out) {
  __shared__ double2 products[kRecombineWarps][kRecombineChunk];
  const int lane = threadIdx.x & 31;
  double2* mine = products[threadIdx.x >> 5];
  const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int n_warps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
  for (int r = warp; r < n_rows; r += n_warps) {
    const uint32_t lo = row_start[r], hi = row_start[r + 1];
    double re = 0.0, im = 0.0;
    for (uint32_t base = lo; base < hi; base += kRecombineChunk) {
      double2 c[4];
      uint32_t t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t j = base + (uint32_t)(u * 32 + lane);
        c[u] = j < hi ? coeff[j] : make_double2(0.0, 0.0);
        t[u] = j < hi ? term_index[j] : 0u;
      }
      double er[4], ei[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = base + (uint32_t)(u * 32 + lane) < hi;
        er[u] = !ok ? 0.0 : e_complex ? e[2 * (size_t)t[u]] : e[t[u]];
        ei[u] = (ok && e_complex) ? e[2 * (size_t)t[u] + 1] : 0.0;
      }
      __syncwarp();  // lane 0 is done with the previous chunk
#pragma unroll
      for (int u = 0; u < 4; ++u)
        mine[u * 32 + lane] = make_double2(__dsub_rn(__dmul_rn(c[u].x, er[u]), __dmul_rn(c[u].y, ei[u])),
                                           __dadd_rn(__dmul_rn(c[u].x, ei[u]), __dmul_rn(c[u].y, er[u])));
      __syncwarp();
      if (lane == 0) {
        const int n = (int)min((uint32_t)kRecombineChunk, hi - base);
#pragma unroll 8
        for (int k = 0; k < n; ++k) {
          const double2 p = mine[k];
          re = __dadd_rn(re, p.x);
          im = __dadd_rn(im, p.y);
        }
      }
    }
    if (lane == 0) out[r] = make_double2(re, im);
  }
}

int launch_recombine(const double* d_e, bool e_complex, const uint32_t* d_row_start, const uint32_t* d_term_index,
                     const double* d_coeffs, int n_rows, double* d_out, cudaStream_t stream) {
  if (n_rows <= 0) return QCUT_OK;
  const int blocks = (int)std::min<int64_t>(((int64_t)n_rows + kRecombineWarps - 1) / kRecombineWarps, 148 * 16);
  recombine_kernel<<<blocks, 32 * kRecombineWarps, 0, stream>>>(d_e, e_complex ? 1 : 0, d_row_start, d_term_index,
                                               reinterpret_cast<const double2*>(d_coeffs), n_rows,
                                               reinterpret_cast<double2*>(d_out));
  QCUT_CUDA(cudaGetLastError());
  return QCUT_OK;
}

}  // namespace qcut

/*
 * qcut_b200.h -- C ABI of the B200-native reconstruction hot path.
 *
 * This is the drop-in boundary for qiskit-addon-cutting's
 *     reconstruct_expectation_values(results, coefficients, observables)
 * (reference: qiskit_addon_cutting/cutting_reconstruction.py:31-170).  The reference is pure
 * Python and has no FFI of its own; the entry points below are what a ctypes binding for that
 * function binds (see INTEGRATION.md for the reference-side stub).  Plain pointers and sizes
 * only -- no torch / numpy / Qiskit types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a negative qcut_status on failure; the message
 *     of the last failure on the calling thread is available from qcut_last_error();
 *   - "d_" pointers are device (HBM) pointers owned by the caller, "h_" pointers are host
 *     pointers; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all launches are asynchronous on `stream`; nothing here synchronises unless stated.
 *
 * Data layout
 *   - shot data: one byte buffer.  Each PUB (one subexperiment result) owns two row-major
 *     byte matrices, observable_measurements (shots x nb_obs) and qpd_measurements
 *     (shots x nb_qpd), exactly the bytes of Qiskit's BitArray.array: big-endian within a
 *     row (reference cutting_reconstruction.py:152-158).  Matrices start at 16-byte aligned
 *     offsets; the buffer must be readable QCUT_DATA_SLACK bytes past the end of the last matrix
 *     (kernels load whole 16-byte chunks / 16-shot groups and mask what lies past a PUB's end).
 *   - masks: per group, n_terms x nb_obs bytes, big-endian rows (same memory order as the
 *     shot rows, so kernels AND raw bytes; reference utils/observable_grouping.py:140-151).
 *   - tallies: int64, tally[pub.tally_offset + t] = sum over shots of
 *     (-1)^(popcount(qpd row) + popcount(obs row AND mask_t))
 *     (reference cutting_reconstruction.py:213-220, summed exactly instead of in fp64).
 */
#ifndef QCUT_B200_H
#define QCUT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCUT_ABI_VERSION 4
/* Shots per scheduling tile (one warp processes one tile of one PUB at a time). */
#define QCUT_TILE_SHOTS 4096
/* Widest observable row (bytes) the kernels accept: 64 bytes = 512 measured qubits per group. */
#define QCUT_MAX_OBS_BYTES 64
/* Readable slack required after the last byte of shot data (see the data layout note above). */
#define QCUT_DATA_SLACK 128

typedef enum qcut_status {
  QCUT_OK = 0,
  QCUT_ERR_INVALID = -1, /* bad argument (null pointer, size, unsupported width) */
  QCUT_ERR_CUDA = -2,    /* a CUDA runtime call failed */
  QCUT_ERR_JIT = -3,     /* NVRTC specialisation failed */
  QCUT_ERR_NOMEM = -4
} qcut_status;

/* One commuting observable group as measured by the PUBs that reference it
 * (reference: CommutingObservableGroup, utils/observable_grouping.py:116-157). */
typedef struct qcut_group_desc {
  uint32_t n_terms;     /* T: number of Pauli terms (masks) in the group                    */
  uint32_t nb_obs;      /* bytes per observable_measurements row                            */
  uint32_t mask_offset; /* byte offset of the T x nb_obs mask rows inside the mask buffer    */
  uint32_t reserved;
} qcut_group_desc;

/* One PUB = the result of one subexperiment: result[i * G + g] in the reference's indexing
 * (reference cutting_reconstruction.py:142,152-155). */
typedef struct qcut_pub_desc {
  uint64_t obs_offset;   /* byte offset of observable_measurements rows in the data buffer  */
  uint64_t qpd_offset;   /* byte offset of qpd_measurements rows                            */
  uint64_t tally_offset; /* index (int64 units) of this PUB's first tally                   */
  uint32_t shots;        /* rows in both matrices (qpd_array.shape[0], reference :155)      */
  uint32_t group;        /* index into the plan's group table                               */
  uint32_t nb_qpd;       /* bytes per qpd_measurements row                                  */
  uint32_t reserved;
} qcut_pub_desc;

/* One (group, term) slot of the observable lookup (reference utils/observable_grouping.py:220-224). */
typedef struct qcut_slot {
  uint32_t group; /* group index *within the subsystem* (0 .. G_label-1) */
  uint32_t term;  /* term index within that group                        */
} qcut_slot;

/* One subsystem ("partition label") as seen by the reduction kernel. */
typedef struct qcut_label_desc {
  uint64_t pub_base;    /* index of the label's first PUB in the pub table (local to this rank)   */
  uint32_t num_groups;  /* G_label: PUB of (coefficient i, group g) is pub_base + i * G_label + g */
  uint32_t lookup_base; /* index into lookup_start[] of this label's first observable            */
} qcut_label_desc;

typedef struct qcut_plan qcut_plan;         /* groups + masks in HBM, kernels specialised for them */
typedef struct qcut_schedule qcut_schedule; /* PUB table in HBM + per-kernel tile lists            */

int qcut_abi_version(void);
const char* qcut_last_error(void);

/* Number of visible CUDA devices (0 if none / no driver).  Never fails. */
int qcut_device_count(void);

/* ---- plan ---------------------------------------------------------------------------------------
 * Replaces the per-call host state the reference builds at cutting_reconstruction.py:113-116.
 * Every group is assigned one of the K1 kernels:
 *   0 = generic (any width), 1 = bit-sliced with run-time masks, >= 2 = bit-sliced kernels
 *   specialised by NVRTC for the groups' masks: one kernel per group for the QCUT_MAX_JIT_KERNELS
 *   groups with the most terms (>= 12), multi-group kernels (one __noinline__ function per group)
 *   for the other groups with rows of <= 8 bytes; what is left uses kernel 0/1.  Rows of 9..32 bytes
 *   (65..256 measured qubits in one register) always get a kernel of their own; wider rows use kernel 0.
 * flags: QCUT_PLAN_NO_JIT disables NVRTC specialisation, QCUT_PLAN_GENERIC_ONLY forces kernel 0.
 * Compiled kernels are cached on disk, keyed by a hash of the generated source, the NVRTC version and
 * the options: $QCUT_CACHE_DIR, default $XDG_CACHE_HOME/qcut_b200 or ~/.cache/qcut_b200;
 * QCUT_CACHE_DIR="" disables the cache. */
#define QCUT_PLAN_NO_JIT 1u
#define QCUT_PLAN_GENERIC_ONLY 2u
#define QCUT_PLAN_ASYNC_JIT 8u /* qcut_plan_create returns at once: NVRTC runs on a background thread and, until it is done, schedules are served by the ahead-of-time kernels (run-time masks / generic) -- no cold start */
#define QCUT_PLAN_STAGED 4u /* specialised kernels prefetch each step through shared memory with bulk async copies (UBLKCP + mbarrier) instead of loading straight from HBM; measured slower on B200 for this ALU-issue-bound kernel, kept for experiments */
#define QCUT_MAX_JIT_KERNELS 8
int qcut_plan_create(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks,
                     size_t mask_bytes, unsigned flags, qcut_plan** out_plan);
void qcut_plan_destroy(qcut_plan* plan);
int qcut_plan_group_kernel(const qcut_plan* plan, int group);
/* Copies the generated CUDA source of specialised kernel `kernel` (>= 2), NUL terminated, into
 * buf; returns the full length needed (0 if there is no such kernel).  For inspection only. */
size_t qcut_plan_source(const qcut_plan* plan, int kernel, char* buf, size_t buf_len);
/* State of the specialised kernels: 0 none requested, 1 being compiled (QCUT_PLAN_ASYNC_JIT), 2 ready, -1 failed (the
 * plan keeps working on the ahead-of-time kernels; qcut_plan_wait reports the error).  qcut_plan_wait blocks until
 * the background compilation has finished. */
int qcut_plan_jit_state(const qcut_plan* plan);
int qcut_plan_wait(qcut_plan* plan);

/* ---- schedule -----------------------------------------------------------------------------------
 * Copies the PUB table to HBM and cuts every PUB into tiles of QCUT_TILE_SHOTS consecutive shots,
 * listed per kernel so that each K1 kernel streams only its own tiles.  The two descriptor buffers
 * come from a small per-process cache (cudaFree per call was measured at up to 165 ms).  Destroy never blocks:
 * it records an event on the stream of the schedule's last qcut_tally_launch and parks the buffers with it, so
 * launches that read qcut_schedule_device_pubs() must have been queued on that stream (or have completed)
 * before qcut_schedule_destroy is called. */
int qcut_schedule_create(const qcut_plan* plan, const qcut_pub_desc* h_pubs, int64_t n_pubs,
                         qcut_schedule** out_schedule);
void qcut_schedule_destroy(qcut_schedule* schedule);
int64_t qcut_schedule_num_tiles(const qcut_schedule* schedule);
const qcut_pub_desc* qcut_schedule_device_pubs(const qcut_schedule* schedule);

/* ---- K1: parity tallies (reference cutting_reconstruction.py:152-161 + :196-222) -------------
 * Zeroes d_tallies[0 .. n_tallies) and accumulates the exact sign tallies of every PUB of the
 * schedule into it.  Returns the number of kernels launched in *n_launches if not NULL.
 * Stream-ordered on `stream`: when the plan's kernels are short (< 128 MB of shot bytes each on
 * average) they are forked onto side streams owned by the plan and joined back before the call
 * returns, so work queued on `stream` afterwards sees complete tallies (QCUT_SIDE_STREAMS=0..3
 * overrides the number of side streams).  Calls on the same plan from several host threads are
 * serialised over those side streams. */
int qcut_tally_launch(const qcut_plan* plan, const qcut_schedule* schedule, const uint8_t* d_data,
                      int64_t* d_tallies, int64_t n_tallies, void* stream, int* n_launches);

/* ---- K2: coefficient-weighted product over subsystems, summed over subexperiment tuples -------
 * (reference cutting_reconstruction.py:134-136,163-168):
 *   out[k] = sum_i coeff[i] * prod_label mean_{(g,t) in lookup(label,k)} tally / shots
 * with the expectation of PUB (label, i, g) taken as tally/shots (one correctly rounded fp64
 * division).  Deterministic: fixed partition of i, fixed-order combination; for n_coeffs <= 1024 the
 * summation order is exactly the reference's.  d_workspace must hold
 * qcut_reduce_workspace_bytes(n_coeffs, n_obs) bytes.  `flags`: QCUT_REDUCE_SINGLE_SLOT asserts that
 * every observable has exactly one lookup slot in every subsystem (what Qiskit's grouping yields): the
 * kernels then run straight-line code specialised for 1..4 subsystems, otherwise the general form;
 * QCUT_REDUCE_SEQUENTIAL sums the tuples in strictly ascending order whatever n_coeffs is (the
 * reference's `expvals += coeff[0] * current_expvals` order; one CTA column walks all tuples). */
#define QCUT_REDUCE_SINGLE_SLOT 1
#define QCUT_REDUCE_SEQUENTIAL 2
size_t qcut_reduce_workspace_bytes(int64_t n_coeffs, int n_obs);
int qcut_reduce_launch(const int64_t* d_tallies, const qcut_pub_desc* d_pubs,
                       const qcut_label_desc* d_labels, int n_labels,
                       const uint32_t* d_lookup_start, const qcut_slot* d_lookup,
                       const double* d_coeffs, int64_t n_coeffs, int n_obs, double* d_out,
                       void* d_workspace, int flags, void* stream);

/* ---- K1, reference rounding (opt-in): per-(PUB, term) expectation values accumulated exactly like the
 * reference does -- subsystem_expvals[k] += (1 / shots) * sign, shot by shot, in fp64
 * (cutting_reconstruction.py:155-161) -- so d_expvals[pub.tally_offset + t] is BIT-IDENTICAL to the
 * reference's value, including its rounding noise (up to ~6e-12 from tally / shots at 10^6 shots).
 * One thread walks the chain of one (PUB, term) in shot order; feed the result to
 * qcut_reduce_expvals_launch with QCUT_REDUCE_SEQUENTIAL for bit-identical final values.  Reads the
 * PUB table of `schedule`; about an order of magnitude slower than qcut_tally_launch. */
int qcut_expvals_literal_launch(const qcut_plan* plan, const qcut_schedule* schedule, const uint8_t* d_data,
                                double* d_expvals, int64_t n_expvals, void* stream);

/* ---- SamplerV1 quasi-distributions (reference cutting_reconstruction.py:143-149,173-193) -------
 * For PUB p, outcomes [d_dist_start[p], d_dist_start[p+1]) hold (obs bits, qpd bits, weight):
 *   d_expvals[pub.tally_offset + t] = sum_o weight_o * sign_o(t),
 * accumulated sequentially in the order given, i.e. bit-identical to the reference's dict-order
 * loop.  obs/qpd bit strings and masks are little-endian uint64 limbs (n_limbs_obs / n_limbs_qpd
 * per outcome, n_limbs_obs per mask); d_groups[g].mask_offset indexes d_mask_limbs in limbs and
 * nb_obs is ignored.  Only pub.group and pub.tally_offset of d_pubs are read. */
int qcut_quasi_launch(const qcut_pub_desc* d_pubs, int64_t n_pubs, const qcut_group_desc* d_groups,
                      const uint64_t* d_mask_limbs, const int64_t* d_dist_start,
                      const uint64_t* d_obs_bits, int n_limbs_obs, const uint64_t* d_qpd_bits,
                      int n_limbs_qpd, const double* d_weights, double* d_expvals, void* stream);
/* int64 tallies -> fp64 expectation values of the PUBs of a (device) PUB table: d_expvals[pub.tally_offset + t] =
 * d_tallies[...] / pub.shots (correctly rounded; 0 for PUBs without shots), t < n_terms of the PUB's group in `plan`.
 * For calls that mix SamplerV1 and SamplerV2 subsystems (the reference decides per subsystem,
 * cutting_reconstruction.py:143,150): both kinds then feed one qcut_reduce_expvals_launch. */
int qcut_tallies_to_expvals_launch(const qcut_plan* plan, const qcut_pub_desc* d_pubs, int64_t n_pubs,
                                   const int64_t* d_tallies, double* d_expvals, void* stream);
/* K2 on fp64 per-PUB expectation values instead of tallies (V1 path). */
int qcut_reduce_expvals_launch(const double* d_expvals, const qcut_pub_desc* d_pubs,
                               const qcut_label_desc* d_labels, int n_labels,
                               const uint32_t* d_lookup_start, const qcut_slot* d_lookup,
                               const double* d_coeffs, int64_t n_coeffs, int n_obs, double* d_out,
                               void* d_workspace, int flags, void* stream);

/* ---- term recombination (reference utils/observable_terms.py:60-90) ------------------------------
 * The step right after the path: the expectation value of every observable (SparsePauliOp) is
 *   out[r] = sum_{j in [d_row_start[r], d_row_start[r+1])} coeff[j] * e[d_term_index[j]]
 * in complex fp64, accumulated sequentially in the order given, one separate multiply / add per
 * operation (no FMA contraction): bit-identical to the reference's `rv += coeff * term_expval`
 * loop.  d_coeffs and d_out are interleaved (re, im) pairs; d_term_expvals holds doubles
 * (expvals_complex = 0: imaginary part +0.0, e.g. the output of qcut_reduce_launch, still in HBM)
 * or interleaved pairs (expvals_complex = 1).  Terms with a zero coefficient are left out of
 * the rows by the caller (reference :71-72). */
int qcut_recombine_launch(const double* d_term_expvals, int expvals_complex,
                          const uint32_t* d_row_start, const uint32_t* d_term_index,
                          const double* d_coeffs, int n_rows, double* d_out, void* stream);

/* ---- PUB de-duplication at ingest ----------------------------------
 * The reference emits one subexperiment per (sample, group) and label even when that label's own map
 * ids repeat (cutting_experiments.py:146-157, docs/explanation/index.rst:188), so a caller who runs each
 * distinct circuit once hands the same BitArray buffers over many times.  A PUB whose six attributes
 * (observable buffer address, qpd buffer address, shots, row widths, group variant) equal an earlier
 * PUB's has the same bytes under the same masks: it is neither staged nor tallied again.  Contents are
 * never hashed.  The finder keeps its hash table across the blocks of one call.
 * qcut_dedup_block: for PUBs base .. base+n-1 writes rep[j] = global index of the first PUB with the
 * same attributes (base+j itself when there is none, always for PUBs without shots); returns the
 * number of duplicates in the block, or -1 on a bad argument. */
typedef struct qcut_dedup qcut_dedup;
int qcut_dedup_create(qcut_dedup** out);
void qcut_dedup_destroy(qcut_dedup* finder);
void qcut_dedup_clear(qcut_dedup* finder); /* forget every PUB, keep the memory (a finder is reused call after call) */
int64_t qcut_dedup_block(qcut_dedup* finder, int64_t base, const uint64_t* obs_ptr, const uint64_t* qpd_ptr,
                         const int64_t* shots, const int64_t* nb_obs, const int64_t* nb_qpd,
                         const uint32_t* variant, int64_t n, int64_t* rep);

/* ---- host staging: gather many borrowed BitArray buffers into one pinned buffer ---------------
 * Copies n_arrays host buffers (h_src[i], h_bytes[i] bytes) to h_dst + h_dst_offset[i] with
 * n_threads worker threads (0 = hardware concurrency).  Pure host code; used on the end-to-end
 * path so that one large H2D copy replaces 2 * n_pubs small ones. */
int qcut_gather_host(const void* const* h_src, const uint64_t* h_bytes, const uint64_t* h_dst_offset,
                     int64_t n_arrays, void* h_dst, int n_threads);

/* Pipelined version of the above for the end-to-end path: worker threads gather runs of arrays into
 * slots of a (small) pinned buffer and queue one H2D copy per slot on `stream`, so the host-side
 * gather overlaps the PCIe transfer and only pinned_bytes (e.g. 64 MiB) of pinned memory is needed
 * whatever the problem size.  Arrays must be sorted by h_dst_offset; d_dst + h_dst_offset[i] receives
 * array i.  Returns when every byte has left host memory (the copies may still be in flight).
 * Rings of >= 64 MiB are written with non-temporal stores (QCUT_STAGING_NT=0/1 overrides). */
int qcut_stage_h2d(const void* const* h_src, const uint64_t* h_bytes, const uint64_t* h_dst_offset,
                   int64_t n_arrays, void* h_pinned, size_t pinned_bytes, uint8_t* d_dst, int n_threads,
                   void* stream);

/* ---- host-only helpers (no GPU needed): inspection and CPU-side testing of the specialisation -----
 * qcut_generate_source: the CUDA translation unit NVRTC receives when group `group` of this table
 * is specialised.  Returns the length needed including the NUL; 0 if the group cannot be
 * specialised (unsupported width / too many terms) or the input is invalid.
 * qcut_compile_source: NVRTC -> sm_100a cubin (size, and the image if cubin_buf is large enough). */
size_t qcut_generate_source(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks,
                            size_t mask_bytes, int group, char* buf, size_t buf_len);
int qcut_compile_source(const char* source, size_t* cubin_bytes, char* cubin_buf, size_t cubin_buf_len);
/* Runs the kernel assignment and every NVRTC compilation qcut_plan_create would run for these groups,
 * without touching a device; returns the number of specialised kernels and their total cubin size. */
int qcut_plan_compile_check(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks,
                            size_t mask_bytes, int* n_kernels, size_t* cubin_bytes);

/* The CUDA source of specialised kernel `kernel` (0-based) of the plan these groups would get, and -- when the pointers
 * are given -- for every group the index of the specialised kernel that serves it (< 0: an ahead-of-time kernel) and
 * the index of its network inside that kernel.  Returns the length needed including the NUL, 0 if there is no such
 * kernel.  Host only (runs the kernel assignment and NVRTC like qcut_plan_compile_check). */
size_t qcut_plan_kernel_source(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks, size_t mask_bytes,
                               int kernel, char* buf, size_t buf_len, int* kernel_of_group, int* variant_of_group);

#ifdef __cplusplus
}
#endif
#endif /* QCUT_B200_H */

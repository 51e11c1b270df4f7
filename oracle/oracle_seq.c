/*
 * oracle_seq.c -- CPU restatement of the reconstruction hot path in plain C.
 * TEST INFRASTRUCTURE ONLY: used by tests/ and by bench.py's CPU-baseline legs as the checker /
 * reported baseline, never linked or called by the product (qiskit_addon_cutting_b200).
 *
 * Follows /root/reference/qiskit_addon_cutting/cutting_reconstruction.py:
 *   - per-shot loop, big-endian rows:            :156-158  (int.from_bytes(row, "big"))
 *   - sign rule:                                 :213-220  (_process_outcome_v2)
 *   - sequential fp64 accumulation:              :159      (acc += (1/shots) * sign, shot by shot)
 * Compile WITHOUT fast-math and without FMA contraction (oracle/Makefile uses
 * -O2 -ffp-contract=off) so that e_seq reproduces the reference's rounding exactly.
 * popcount is byte-order agnostic, so rows and masks are ANDed byte by byte in memory order.
 *
 * Pinned against the reference's golden vector and source-exec fixtures by tests/test_oracle.py.
 */
#include <stddef.h>
#include <stdint.h>

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct {
  uint64_t obs_offset, qpd_offset, tally_offset;
  uint32_t shots, group, nb_qpd, reserved;
} pub_desc; /* same layout as qcut_pub_desc in include/qcut_b200.h */

typedef struct {
  uint32_t n_terms, nb_obs, mask_offset, reserved;
} group_desc; /* same layout as qcut_group_desc */

static inline int parity_bytes(const uint8_t* a, int n) {
  unsigned x = 0;
  for (int i = 0; i < n; ++i) x ^= a[i];
  return __builtin_popcount(x) & 1;
}

/* One PUB.  e_seq[t]: the reference's sequentially rounded expectation value; tally[t]: exact sum of signs. */
void oracle_pub(const uint8_t* obs, const uint8_t* qpd, int64_t shots, int nb_obs, int nb_qpd,
                const uint8_t* masks, int n_terms, double* e_seq, int64_t* tally) {
  for (int t = 0; t < n_terms; ++t) {
    if (e_seq) e_seq[t] = 0.0;
    if (tally) tally[t] = 0;
  }
  if (shots <= 0) return;
  const double w = 1.0 / (double)shots;
  for (int64_t j = 0; j < shots; ++j) {
    const uint8_t* o = obs + (size_t)j * nb_obs;
    const int qpar = parity_bytes(qpd + (size_t)j * nb_qpd, nb_qpd);
    for (int t = 0; t < n_terms; ++t) {
      const uint8_t* m = masks + (size_t)t * nb_obs;
      unsigned x = 0;
      for (int b = 0; b < nb_obs; ++b) x ^= (unsigned)(o[b] & m[b]);
      const int odd = (__builtin_popcount(x) & 1) ^ qpar;
      const double sign = odd ? -1.0 : 1.0;
      if (e_seq) e_seq[t] += w * sign;
      if (tally) tally[t] += odd ? -1 : 1;
    }
  }
}

/* Many PUBs in the staged layout of the product's C ABI; PUBs are independent -> threads over PUBs. */
typedef struct {
  const uint8_t* data;
  const pub_desc* pubs;
  int64_t n_pubs;
  const group_desc* groups;
  const uint8_t* masks;
  double* e_seq;
  int64_t* tally;
  atomic_llong* next;
} job_t;

static void* worker(void* arg) {
  job_t* job = (job_t*)arg;
  for (;;) {
    const int64_t p = atomic_fetch_add(job->next, 1);
    if (p >= job->n_pubs) return 0;
    const pub_desc* pd = &job->pubs[p];
    const group_desc* gd = &job->groups[pd->group];
    oracle_pub(job->data + pd->obs_offset, job->data + pd->qpd_offset, pd->shots, (int)gd->nb_obs,
               (int)pd->nb_qpd, job->masks + gd->mask_offset, (int)gd->n_terms,
               job->e_seq ? job->e_seq + pd->tally_offset : 0, job->tally ? job->tally + pd->tally_offset : 0);
  }
}

int oracle_max_threads(void) {
  const long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

void oracle_pubs(const uint8_t* data, const pub_desc* pubs, int64_t n_pubs, const group_desc* groups,
                 const uint8_t* masks, double* e_seq, int64_t* tally, int n_threads) {
  atomic_llong next = 0;
  job_t job = {data, pubs, n_pubs, groups, masks, e_seq, tally, &next};
  if (n_threads <= 0) n_threads = oracle_max_threads();
  if (n_threads > 256) n_threads = 256;
  if ((int64_t)n_threads > n_pubs) n_threads = (int)(n_pubs > 0 ? n_pubs : 1);
  if (n_threads == 1) {
    worker(&job);
    return;
  }
  pthread_t tid[256];
  int started = 0;
  for (int i = 0; i < n_threads; ++i)
    if (pthread_create(&tid[started], 0, worker, &job) == 0) ++started;
  if (started == 0) worker(&job);
  for (int i = 0; i < started; ++i) pthread_join(tid[i], 0);
}

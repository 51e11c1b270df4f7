// extern "C" entry points of libqcut_b200.so (see include/qcut_b200.h).
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace qcut {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t err, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(err) + " (" + cudaGetErrorString(err) + ")";
  return QCUT_ERR_CUDA;
}

int sm_count() {
  // per device: a process may drive several (different) GPUs
  constexpr int kMaxDevices = 64;
  static std::atomic<int> cached[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// defined in tally_generic.cu / tally_sliced.cu / reduce.cu / jit.cu
int launch_tally_generic(const qcut_plan*, const uint8_t*, const qcut_pub_desc*, const TileEntry*, int64_t,
                         int64_t*, cudaStream_t);
int launch_tally_sliced_rt(const qcut_plan*, const uint8_t*, const qcut_pub_desc*, const TileEntry*, int64_t,
                           int64_t*, cudaStream_t);
int launch_tally_jit(const JitKernel&, const uint8_t*, const qcut_pub_desc*, const TileEntry*, int64_t, int64_t*,
                     const GroupDev*, cudaStream_t);
bool sliced_rt_supports(uint32_t nb_obs);
bool jit_supports(const qcut_group_desc& gd);
int jit_build(qcut_plan* plan, bool load);
void jit_release(qcut_plan* plan);
template <typename TallyT>
int launch_reduce(const TallyT*, const qcut_pub_desc*, const qcut_label_desc*, int, const uint32_t*,
                  const qcut_slot*, const double*, int64_t, int, double*, void*, bool, bool, cudaStream_t);
size_t reduce_workspace_bytes(int64_t, int);
int launch_literal(const qcut_plan*, const uint8_t*, const qcut_pub_desc*, int64_t, double*, cudaStream_t);
int launch_tallies_to_expvals(const qcut_plan*, const qcut_pub_desc*, int64_t, const int64_t*, double*, cudaStream_t);
int launch_quasi(const qcut_pub_desc*, int64_t, const qcut_group_desc*, const uint64_t*, const int64_t*,
                 const uint64_t*, int, const uint64_t*, int, const double*, double*, cudaStream_t);
int launch_recombine(const double*, bool, const uint32_t*, const uint32_t*, const double*, int, double*, cudaStream_t);

// Groups with at least this many terms get an NVRTC-specialised kernel of their own; smaller groups share kernels
// (their networks are a few instructions: what matters for them is the number and length of the launches -- cfg3
// "local", 12 groups of ~9 terms on 1-7-byte rows, 2.45 GB: nine launches 0.785 of the copy peak, one launch 0.86).
constexpr uint32_t kJitMinTerms = 33;
// Multi-group kernels: distinct mask sets per kernel, and how many such kernels a plan may have
// (groups beyond that use the run-time-mask kernel).
constexpr size_t kMultiVariants = 24;
constexpr size_t kSwitchVariants = 64;    // = kMaxVariants of qcut_schedule_create
constexpr size_t kSwitchMinSets = 4;      // fewer sets of one width share a mixed-width kernel (cfg3 "local": 0.785 vs 0.771)
constexpr uint32_t kSwitchMaxTerms = 32;  // qcut_flush_rt: at most 8 packed counter registers
constexpr size_t kMaxMultiKernels = 32;

}  // namespace qcut

using namespace qcut;

// Decide which K1 kernel serves each group (fills plan->groups_dev[*].kernel / .variant and plan->jit).
static void assign_jit_kernels(qcut_plan* plan) {
  const int n_groups = plan->n_groups;
  const qcut_group_desc* h_groups = plan->groups.data();
  const uint8_t* h_masks = plan->masks.data();
  // Pick the groups to specialise: the QCUT_MAX_JIT_KERNELS distinct (width, masks) sets with the
  // most terms among those with >= kJitMinTerms terms.
  {
    uint32_t min_terms = kJitMinTerms;
    if (const char* env_mt = std::getenv("QCUT_JIT_MIN_TERMS"))  // experiments
      if (*env_mt && std::atoi(env_mt) > 0) min_terms = (uint32_t)std::atoi(env_mt);
    std::map<std::string, std::vector<int>> by_masks;
    for (int g = 0; g < n_groups; ++g) {
      const qcut_group_desc& gd = h_groups[g];
      if (!jit_supports(gd) || (gd.n_terms < min_terms && gd.nb_obs <= 8)) continue;
      std::string key(reinterpret_cast<const char*>(&gd.nb_obs), sizeof(gd.nb_obs));
      key.append(reinterpret_cast<const char*>(h_masks + gd.mask_offset), (size_t)gd.n_terms * gd.nb_obs);
      by_masks[key].push_back(g);
    }
    std::vector<const std::vector<int>*> sets;
    for (const auto& kv : by_masks) sets.push_back(&kv.second);
    std::stable_sort(sets.begin(), sets.end(), [&](const std::vector<int>* a, const std::vector<int>* b) {
      return h_groups[a->front()].n_terms > h_groups[b->front()].n_terms;
    });
    if (sets.size() > QCUT_MAX_JIT_KERNELS) sets.resize(QCUT_MAX_JIT_KERNELS);
    std::vector<bool> taken((size_t)n_groups, false);
    for (const auto* set : sets) {
      JitKernel jk;
      jk.group = set->front();
      for (int g : *set) {
        plan->groups_dev[g].kernel = KERNEL_JIT0 + (uint32_t)plan->jit.size();
        taken[(size_t)g] = true;
      }
      plan->jit.push_back(std::move(jk));
    }
    // Everything else the bit-sliced kernels cover (few terms, or beyond the per-group kernel budget)
    // goes into multi-group kernels, kMultiVariants distinct mask sets per kernel.
    std::map<std::string, std::vector<int>> small;
    std::vector<const std::vector<int>*> small_sets;
    for (int g = 0; g < n_groups; ++g) {
      const qcut_group_desc& gd = h_groups[g];
      if (taken[(size_t)g] || !jit_supports(gd) || gd.nb_obs > 8) continue;  // wide rows: per-group kernels only
      std::string key(reinterpret_cast<const char*>(&gd.nb_obs), sizeof(gd.nb_obs));
      key.append(reinterpret_cast<const char*>(h_masks + gd.mask_offset), (size_t)gd.n_terms * gd.nb_obs);
      auto ins = small.emplace(std::move(key), std::vector<int>());
      if (ins.second) small_sets.push_back(&ins.first->second);
      ins.first->second.push_back(g);
    }
    // Sets of at most kSwitchMaxTerms terms: one kernel per row width and kSwitchVariants sets, the networks behind a
    // switch (generate_switch_source).  QCUT_JIT_SWITCH=0: the older one-function-per-group kernels for everything.
    const char* env_sw = std::getenv("QCUT_JIT_SWITCH");
    if (!(env_sw && *env_sw && std::atoi(env_sw) == 0)) {
      std::vector<const std::vector<int>*> rest;
      std::map<uint32_t, std::vector<const std::vector<int>*>> by_width;
      for (const auto* set : small_sets) {
        const qcut_group_desc& gd = h_groups[set->front()];
        if (gd.n_terms <= kSwitchMaxTerms) by_width[gd.nb_obs].push_back(set);
        else rest.push_back(set);
      }
      for (const auto& kv : by_width) {
        const auto& sets_w = kv.second;
        if (sets_w.size() < kSwitchMinSets && by_width.size() > 1) {  // a handful of sets of this width among others
          rest.insert(rest.end(), sets_w.begin(), sets_w.end());
          continue;
        }
        for (size_t first = 0; first < sets_w.size(); first += kSwitchVariants) {
          if (plan->jit.size() >= (size_t)(QCUT_MAX_JIT_KERNELS + kMaxMultiKernels)) {
            for (size_t v = first; v < sets_w.size(); ++v) rest.push_back(sets_w[v]);
            break;
          }
          JitKernel jk;
          jk.inline_switch = true;
          const size_t last = std::min(sets_w.size(), first + kSwitchVariants);
          for (size_t v = first; v < last; ++v) {
            jk.variants.push_back(sets_w[v]->front());
            for (int g : *sets_w[v]) {
              plan->groups_dev[g].kernel = KERNEL_JIT0 + (uint32_t)plan->jit.size();
              plan->groups_dev[g].variant = (uint32_t)(v - first);
            }
          }
          plan->jit.push_back(std::move(jk));
        }
      }
      small_sets.swap(rest);
    }
    for (size_t first = 0; first < small_sets.size() && plan->jit.size() < (size_t)(QCUT_MAX_JIT_KERNELS + kMaxMultiKernels);
         first += kMultiVariants) {
      JitKernel jk;
      const size_t last = std::min(small_sets.size(), first + kMultiVariants);
      for (size_t v = first; v < last; ++v) {
        jk.variants.push_back(small_sets[v]->front());
        for (int g : *small_sets[v]) {
          plan->groups_dev[g].kernel = KERNEL_JIT0 + (uint32_t)plan->jit.size();
          plan->groups_dev[g].variant = (uint32_t)(v - first);
        }
      }
      plan->jit.push_back(std::move(jk));
    }
  }
}


extern "C" {

int qcut_abi_version(void) { return QCUT_ABI_VERSION; }

const char* qcut_last_error(void) { return g_last_error.c_str(); }

int qcut_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int qcut_plan_create(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks,
                     size_t mask_bytes, unsigned flags, qcut_plan** out_plan) {
  if (!out_plan) return set_error("qcut_plan_create: out_plan is NULL"), QCUT_ERR_INVALID;
  *out_plan = nullptr;
  if (n_groups <= 0 || !h_groups) return set_error("qcut_plan_create: no groups"), QCUT_ERR_INVALID;
  qcut_plan* plan = new (std::nothrow) qcut_plan();
  if (!plan) return set_error("qcut_plan_create: out of host memory"), QCUT_ERR_NOMEM;
  plan->n_groups = n_groups;
  plan->flags = flags;
  plan->groups.assign(h_groups, h_groups + n_groups);
  if (mask_bytes) plan->masks.assign(h_masks, h_masks + mask_bytes);

  // Mask rows -> zero padded little-endian words in memory order (what the kernels AND with).
  std::vector<uint32_t> words;
  plan->groups_dev.resize(n_groups);
  for (int g = 0; g < n_groups; ++g) {
    const qcut_group_desc& gd = h_groups[g];
    if (gd.nb_obs == 0 || gd.nb_obs > QCUT_MAX_OBS_BYTES) {
      set_error("qcut_plan_create: group " + std::to_string(g) + " has an unsupported observable row width of " +
                std::to_string(gd.nb_obs) + " bytes (1.." + std::to_string(QCUT_MAX_OBS_BYTES) + " supported)");
      delete plan;
      return QCUT_ERR_INVALID;
    }
    if ((size_t)gd.mask_offset + (size_t)gd.n_terms * gd.nb_obs > mask_bytes) {
      set_error("qcut_plan_create: group " + std::to_string(g) + " masks run past the mask buffer");
      delete plan;
      return QCUT_ERR_INVALID;
    }
    GroupDev& dev = plan->groups_dev[g];
    dev = GroupDev{};
    dev.n_terms = gd.n_terms;
    dev.nb_obs = gd.nb_obs;
    dev.n_words = (gd.nb_obs + 3) / 4;
    dev.word_offset = (uint32_t)words.size();
    dev.kernel = (!(flags & QCUT_PLAN_GENERIC_ONLY) && sliced_rt_supports(gd.nb_obs)) ? KERNEL_SLICED_RT : KERNEL_GENERIC;
    plan->fallback_kernel.push_back(dev.kernel);
    for (uint32_t t = 0; t < gd.n_terms; ++t) {
      const uint8_t* row = h_masks + gd.mask_offset + (size_t)t * gd.nb_obs;
      for (uint32_t w = 0; w < dev.n_words; ++w) {
        uint32_t v = 0;
        for (uint32_t j = 0; j < 4 && 4 * w + j < gd.nb_obs; ++j) v |= (uint32_t)row[4 * w + j] << (8 * j);
        words.push_back(v);
      }
    }
  }

  const bool want_jit = !(flags & (QCUT_PLAN_NO_JIT | QCUT_PLAN_GENERIC_ONLY));
  if (want_jit) {
    assign_jit_kernels(plan);
    if (!(flags & QCUT_PLAN_ASYNC_JIT)) {
      const int rc = jit_build(plan, true);
      if (rc != QCUT_OK) {
        qcut_plan_destroy(plan);
        return rc;
      }
      plan->jit_state.store(plan->jit.empty() ? qcut_plan::JIT_NONE : qcut_plan::JIT_READY, std::memory_order_release);
    }
  }

  auto fail = [&](cudaError_t err, const char* what) {
    const int rc = cuda_fail(err, what);
    qcut_plan_destroy(plan);
    return rc;
  };
  cudaError_t err = cudaMalloc(&plan->d_groups, sizeof(GroupDev) * n_groups);
  if (err != cudaSuccess) return fail(err, "cudaMalloc(groups)");
  err = cudaMemcpy(plan->d_groups, plan->groups_dev.data(), sizeof(GroupDev) * n_groups, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) return fail(err, "cudaMemcpy(groups)");
  const size_t wbytes = std::max<size_t>(words.size(), 1) * sizeof(uint32_t);
  err = cudaMalloc(&plan->d_mask_words, wbytes);
  if (err != cudaSuccess) return fail(err, "cudaMalloc(mask words)");
  if (!words.empty()) {
    err = cudaMemcpy(plan->d_mask_words, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return fail(err, "cudaMemcpy(mask words)");
  }
  if (want_jit && (flags & QCUT_PLAN_ASYNC_JIT) && !plan->jit.empty()) {
    // NVRTC + module loading on a background thread; until it is done the ahead-of-time kernels serve every group
    int device = 0;
    cudaGetDevice(&device);
    plan->jit_state.store(qcut_plan::JIT_PENDING, std::memory_order_release);
    plan->jit_thread = std::thread([plan, device]() {
      int rc = cudaSetDevice(device) == cudaSuccess ? jit_build(plan, true) : QCUT_ERR_CUDA;
      if (rc != QCUT_OK) plan->jit_error = plan->jit_cancel.load() ? "cancelled" : qcut_last_error();
      plan->jit_state.store(rc == QCUT_OK ? qcut_plan::JIT_READY : qcut_plan::JIT_FAILED, std::memory_order_release);
    });
  }
  *out_plan = plan;
  return QCUT_OK;
}

int qcut_plan_jit_state(const qcut_plan* plan) { return plan ? plan->jit_state.load(std::memory_order_acquire) : 0; }

int qcut_plan_wait(qcut_plan* plan) {
  if (!plan) return set_error("qcut_plan_wait: plan is NULL"), QCUT_ERR_INVALID;
  if (plan->jit_thread.joinable()) plan->jit_thread.join();
  if (plan->jit_state.load(std::memory_order_acquire) == qcut_plan::JIT_FAILED)
    return set_error("background specialisation failed: " + plan->jit_error), QCUT_ERR_JIT;
  return QCUT_OK;
}

void qcut_plan_destroy(qcut_plan* plan) {
  if (!plan) return;
  plan->jit_cancel.store(true);  // a compilation in flight stops after the kernel it is working on
  if (plan->jit_thread.joinable()) plan->jit_thread.join();
  jit_release(plan);
  for (int i = 0; i < qcut_plan::kSideStreams; ++i) {
    if (plan->side[i]) cudaStreamDestroy(plan->side[i]);
    if (plan->join_event[i]) cudaEventDestroy(plan->join_event[i]);
  }
  if (plan->fork_event) cudaEventDestroy(plan->fork_event);
  if (plan->d_groups) cudaFree(plan->d_groups);
  if (plan->d_mask_words) cudaFree(plan->d_mask_words);
  delete plan;
}

int qcut_plan_group_kernel(const qcut_plan* plan, int group) {
  if (!plan || group < 0 || group >= plan->n_groups) return set_error("qcut_plan_group_kernel: bad group"), QCUT_ERR_INVALID;
  return (int)plan->kernel_for(group, plan->jit_state.load(std::memory_order_acquire) != qcut_plan::JIT_PENDING &&
                                          plan->jit_state.load(std::memory_order_acquire) != qcut_plan::JIT_FAILED);
}

size_t qcut_plan_source(const qcut_plan* plan, int kernel, char* buf, size_t buf_len) {
  if (!plan || kernel < (int)KERNEL_JIT0 || kernel >= plan->n_kernels()) return 0;
  if (plan->jit_state.load(std::memory_order_acquire) == qcut_plan::JIT_PENDING) return 0;  // still being written
  const std::string& src = plan->jit[kernel - KERNEL_JIT0].source;
  if (buf && buf_len) {
    const size_t n = std::min(buf_len - 1, src.size());
    std::memcpy(buf, src.data(), n);
    buf[n] = '\0';
  }
  return src.size() + 1;
}

}  // extern "C"

namespace {
// Descriptor buffers of schedules are recycled: a schedule lives for one reconstruction call, and
// cudaFree of its (few MB) buffers was measured at up to 165 ms on a busy device -- half of an
// end-to-end call.  Freed buffers wait here (a handful per device) for the next schedule.
// `ready` (may be null) completes when the last launch that reads the buffer has finished: the buffer is
// parked without waiting and whoever takes it next waits on the event (normally long complete).
struct CachedBuffer { void* ptr; size_t bytes; int device; cudaEvent_t ready; };
std::mutex g_cache_mutex;
std::vector<CachedBuffer> g_cache;
constexpr size_t kCacheEntries = 8;

cudaError_t cached_alloc(void** ptr, size_t bytes, size_t* capacity, int device) {
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    int best = -1;
    for (int i = 0; i < (int)g_cache.size(); ++i) {
      const CachedBuffer& c = g_cache[(size_t)i];
      if (c.device == device && c.bytes >= bytes && c.bytes <= 4 * bytes + (1u << 20) &&
          (best < 0 || c.bytes < g_cache[(size_t)best].bytes))
        best = i;
    }
    if (best >= 0) {
      const CachedBuffer hit = g_cache[(size_t)best];
      g_cache.erase(g_cache.begin() + best);
      *ptr = hit.ptr;
      *capacity = hit.bytes;
      if (hit.ready) {
        cudaEventSynchronize(hit.ready);
        cudaEventDestroy(hit.ready);
      }
      return cudaSuccess;
    }
  }
  *capacity = bytes;
  return cudaMalloc(ptr, bytes);
}

void cached_release(void* ptr, size_t bytes, int device, cudaEvent_t ready) {
  if (!ptr) {
    if (ready) cudaEventDestroy(ready);
    return;
  }
  CachedBuffer evict{nullptr, 0, 0, nullptr};
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    g_cache.push_back(CachedBuffer{ptr, bytes, device, ready});
    if (g_cache.size() > kCacheEntries) {
      evict = g_cache.front();
      g_cache.erase(g_cache.begin());
    }
  }
  if (evict.ptr) {
    if (evict.ready) {
      cudaEventSynchronize(evict.ready);
      cudaEventDestroy(evict.ready);
    }
    cudaFree(evict.ptr);  // synchronises with the device by itself
  }
}
}  // namespace

extern "C" {

int qcut_schedule_create(const qcut_plan* plan, const qcut_pub_desc* h_pubs, int64_t n_pubs,
                         qcut_schedule** out_schedule) {
  if (!out_schedule) return set_error("qcut_schedule_create: out_schedule is NULL"), QCUT_ERR_INVALID;
  *out_schedule = nullptr;
  if (!plan || n_pubs < 0 || (n_pubs > 0 && !h_pubs)) return set_error("qcut_schedule_create: bad argument"), QCUT_ERR_INVALID;
  if (n_pubs >= (int64_t)1 << 32) return set_error("qcut_schedule_create: more than 2^32 PUBs"), QCUT_ERR_INVALID;
  const int nk = plan->n_kernels();
  // specialised kernels only if they are there now (a background compilation may still be running)
  const int jit_state = plan->jit_state.load(std::memory_order_acquire);
  const bool use_jit = jit_state != qcut_plan::JIT_PENDING && jit_state != qcut_plan::JIT_FAILED;
  // Tiles are listed kernel by kernel and, inside a kernel, variant by variant (PUB order within).
  constexpr int64_t kMaxVariants = 64;
  auto key_of = [&](const qcut_pub_desc& pd) {
    const GroupDev& gd = plan->groups_dev[pd.group];
    const int64_t kernel = plan->kernel_for((int)pd.group, use_jit);
    return kernel * kMaxVariants + (kernel >= (int64_t)KERNEL_JIT0 ? std::min<int64_t>(gd.variant, kMaxVariants - 1) : 0);
  };
  std::vector<int64_t> count((size_t)nk * kMaxVariants + 1, 0);
  std::vector<int64_t> bytes((size_t)nk, 0);
  for (int64_t p = 0; p < n_pubs; ++p) {
    if (h_pubs[p].group >= (uint32_t)plan->n_groups) {
      set_error("qcut_schedule_create: PUB " + std::to_string(p) + " references group " + std::to_string(h_pubs[p].group) +
                " but the plan has " + std::to_string(plan->n_groups));
      return QCUT_ERR_INVALID;
    }
    const int64_t tiles = ((int64_t)h_pubs[p].shots + QCUT_TILE_SHOTS - 1) / QCUT_TILE_SHOTS;
    count[(size_t)key_of(h_pubs[p]) + 1] += tiles;
    bytes[(size_t)plan->kernel_for((int)h_pubs[p].group, use_jit)] +=
        (int64_t)h_pubs[p].shots * ((int64_t)plan->groups[h_pubs[p].group].nb_obs + h_pubs[p].nb_qpd);
  }
  for (size_t k = 0; k + 1 < count.size(); ++k) count[k + 1] += count[k];
  qcut_schedule* sch = new (std::nothrow) qcut_schedule();
  if (!sch) return set_error("qcut_schedule_create: out of host memory"), QCUT_ERR_NOMEM;
  sch->n_pubs = n_pubs;
  sch->use_jit = use_jit;
  sch->n_tiles = count.back();
  sch->kernel_bytes = bytes;
  sch->kernel_tile_begin.resize((size_t)nk + 1);
  for (int k = 0; k <= nk; ++k) sch->kernel_tile_begin[(size_t)k] = count[(size_t)k * kMaxVariants];
  std::vector<TileEntry> tiles((size_t)std::max<int64_t>(sch->n_tiles, 1));
  std::vector<int64_t> cursor(count.begin(), count.end() - 1);
  for (int64_t p = 0; p < n_pubs; ++p) {
    const int64_t n = ((int64_t)h_pubs[p].shots + QCUT_TILE_SHOTS - 1) / QCUT_TILE_SHOTS;
    int64_t& c = cursor[(size_t)key_of(h_pubs[p])];
    for (int64_t t = 0; t < n; ++t) tiles[(size_t)c++] = TileEntry{(uint32_t)p, (uint32_t)t};
  }
  auto fail = [&](cudaError_t err, const char* what) {
    const int rc = cuda_fail(err, what);
    qcut_schedule_destroy(sch);
    return rc;
  };
  cudaError_t err = cudaGetDevice(&sch->device);
  if (err != cudaSuccess) return fail(err, "cudaGetDevice");
  err = cached_alloc(reinterpret_cast<void**>(&sch->d_pubs), sizeof(qcut_pub_desc) * (size_t)std::max<int64_t>(n_pubs, 1),
                     &sch->pubs_capacity, sch->device);
  if (err != cudaSuccess) return fail(err, "cudaMalloc(pubs)");
  if (n_pubs) {
    err = cudaMemcpy(sch->d_pubs, h_pubs, sizeof(qcut_pub_desc) * (size_t)n_pubs, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return fail(err, "cudaMemcpy(pubs)");
  }
  err = cached_alloc(reinterpret_cast<void**>(&sch->d_tiles), sizeof(TileEntry) * tiles.size(), &sch->tiles_capacity, sch->device);
  if (err != cudaSuccess) return fail(err, "cudaMalloc(tiles)");
  if (sch->n_tiles) {
    err = cudaMemcpy(sch->d_tiles, tiles.data(), sizeof(TileEntry) * (size_t)sch->n_tiles, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return fail(err, "cudaMemcpy(tiles)");
  }
  *out_schedule = sch;
  return QCUT_OK;
}

void qcut_schedule_destroy(qcut_schedule* schedule) {
  if (!schedule) return;
  // Launches that read the descriptors may still be in flight.  Nothing waits here: an event recorded on the
  // stream of the schedule's last qcut_tally_launch (which also orders every later launch the caller queued on
  // that stream, e.g. qcut_reduce_launch reading d_pubs) travels with the parked buffers, and the next schedule
  // that takes them waits on it.  Only if that is impossible (foreign device, dead stream) the device is drained.
  cudaEvent_t ready[2] = {nullptr, nullptr};
  if (schedule->launched) {
    int current = -1;
    bool ok = cudaGetDevice(&current) == cudaSuccess;
    if (ok && current != schedule->device) ok = cudaSetDevice(schedule->device) == cudaSuccess;
    for (int i = 0; ok && i < 2; ++i) {
      ok = cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventRecord(ready[i], schedule->last_stream) == cudaSuccess;
    }
    if (!ok) {
      cudaGetLastError();
      for (cudaEvent_t& e : ready)
        if (e) { cudaEventDestroy(e); e = nullptr; }
      cudaDeviceSynchronize();
    }
    if (current >= 0 && current != schedule->device) cudaSetDevice(current);
  }
  cached_release(schedule->d_pubs, schedule->pubs_capacity, schedule->device, ready[0]);
  cached_release(schedule->d_tiles, schedule->tiles_capacity, schedule->device, ready[1]);
  delete schedule;
}

int64_t qcut_schedule_num_tiles(const qcut_schedule* schedule) { return schedule ? schedule->n_tiles : 0; }

const qcut_pub_desc* qcut_schedule_device_pubs(const qcut_schedule* schedule) {
  return schedule ? schedule->d_pubs : nullptr;
}

int qcut_tally_launch(const qcut_plan* plan, const qcut_schedule* schedule, const uint8_t* d_data,
                      int64_t* d_tallies, int64_t n_tallies, void* stream, int* n_launches) {
  if (n_launches) *n_launches = 0;
  if (!plan || !schedule) return set_error("qcut_tally_launch: plan or schedule is NULL"), QCUT_ERR_INVALID;
  if (n_tallies < 0) return set_error("qcut_tally_launch: negative size"), QCUT_ERR_INVALID;
  if (n_tallies > 0 && !d_tallies) return set_error("qcut_tally_launch: d_tallies is NULL"), QCUT_ERR_INVALID;
  if (schedule->n_tiles > 0 && !d_data) return set_error("qcut_tally_launch: d_data is NULL"), QCUT_ERR_INVALID;
  if ((int)schedule->kernel_tile_begin.size() != plan->n_kernels() + 1)
    return set_error("qcut_tally_launch: schedule was built for a different plan"), QCUT_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  schedule->last_stream = s;
  schedule->launched = true;
  if (n_tallies > 0) QCUT_CUDA(cudaMemsetAsync(d_tallies, 0, sizeof(int64_t) * (size_t)n_tallies, s));
  int active = 0;
  int64_t total_bytes = 0;
  for (int k = 0; k < plan->n_kernels(); ++k) {
    active += schedule->kernel_tile_begin[k + 1] > schedule->kernel_tile_begin[k];
    total_bytes += schedule->kernel_bytes[(size_t)k];
  }
  // Several SHORT kernels (under ~25 us of HBM time each): fork to side streams after the memset and join before
  // returning to `s`, so the launch latency / ramp of one overlaps the tail of another (cfg3 "local", 8 kernels
  // of 34 MB: 0.137 -> 0.083 ms).  Long kernels stay in order on `s`: run side by side they share the SMs'
  // instruction caches and DRAM pages and lose 10 % (cfg5, 13 kernels of 0.9 GB: 1.72 -> 1.92 ms).
  constexpr int64_t kShortKernelBytes = (int64_t)128 << 20;
  const char* env_side = std::getenv("QCUT_SIDE_STREAMS");
  const int want_side = env_side && *env_side ? std::atoi(env_side)
                        : (active > 1 && total_bytes / active < kShortKernelBytes) ? (int)qcut_plan::kSideStreams : 0;
  const int n_side = std::max(0, std::min({active - 1, want_side, (int)qcut_plan::kSideStreams}));
  std::unique_lock<std::mutex> lock(plan->side_mutex, std::defer_lock);
  if (n_side > 0) {
    lock.lock();
    if (!plan->fork_event) {
      QCUT_CUDA(cudaEventCreateWithFlags(&plan->fork_event, cudaEventDisableTiming));
      for (int i = 0; i < qcut_plan::kSideStreams; ++i) {
        QCUT_CUDA(cudaStreamCreateWithFlags(&plan->side[i], cudaStreamNonBlocking));
        QCUT_CUDA(cudaEventCreateWithFlags(&plan->join_event[i], cudaEventDisableTiming));
      }
    }
    QCUT_CUDA(cudaEventRecord(plan->fork_event, s));
    for (int i = 0; i < n_side; ++i) QCUT_CUDA(cudaStreamWaitEvent(plan->side[i], plan->fork_event, 0));
  }
  int launches = 0;
  for (int k = 0; k < plan->n_kernels(); ++k) {
    const int64_t begin = schedule->kernel_tile_begin[k];
    const int64_t n = schedule->kernel_tile_begin[k + 1] - begin;
    if (n <= 0) continue;
    const TileEntry* tiles = schedule->d_tiles + begin;
    const int lane = launches % (n_side + 1);  // 0 = the caller's stream
    cudaStream_t ks = lane == 0 ? s : plan->side[lane - 1];
    int rc;
    if (k == (int)KERNEL_GENERIC) rc = launch_tally_generic(plan, d_data, schedule->d_pubs, tiles, n, d_tallies, ks);
    else if (k == (int)KERNEL_SLICED_RT) rc = launch_tally_sliced_rt(plan, d_data, schedule->d_pubs, tiles, n, d_tallies, ks);
    else rc = launch_tally_jit(plan->jit[k - KERNEL_JIT0], d_data, schedule->d_pubs, tiles, n, d_tallies, plan->d_groups, ks);
    if (rc != QCUT_OK) return rc;
    ++launches;
  }
  for (int i = 0; i < n_side; ++i) {
    QCUT_CUDA(cudaEventRecord(plan->join_event[i], plan->side[i]));
    QCUT_CUDA(cudaStreamWaitEvent(s, plan->join_event[i], 0));
  }
  if (n_launches) *n_launches = launches;
  return QCUT_OK;
}

size_t qcut_reduce_workspace_bytes(int64_t n_coeffs, int n_obs) { return reduce_workspace_bytes(n_coeffs, n_obs); }

int qcut_reduce_launch(const int64_t* d_tallies, const qcut_pub_desc* d_pubs, const qcut_label_desc* d_labels,
                       int n_labels, const uint32_t* d_lookup_start, const qcut_slot* d_lookup,
                       const double* d_coeffs, int64_t n_coeffs, int n_obs, double* d_out,
                       void* d_workspace, int flags, void* stream) {
  if (!d_out || !d_workspace || n_labels < 0 || n_coeffs < 0)
    return set_error("qcut_reduce_launch: bad argument"), QCUT_ERR_INVALID;
  return launch_reduce<int64_t>(d_tallies, d_pubs, d_labels, n_labels, d_lookup_start, d_lookup, d_coeffs,
                                n_coeffs, n_obs, d_out, d_workspace, (flags & QCUT_REDUCE_SINGLE_SLOT) != 0,
                                (flags & QCUT_REDUCE_SEQUENTIAL) != 0, static_cast<cudaStream_t>(stream));
}

int qcut_reduce_expvals_launch(const double* d_expvals, const qcut_pub_desc* d_pubs,
                               const qcut_label_desc* d_labels, int n_labels, const uint32_t* d_lookup_start,
                               const qcut_slot* d_lookup, const double* d_coeffs, int64_t n_coeffs, int n_obs,
                               double* d_out, void* d_workspace, int flags, void* stream) {
  if (!d_out || !d_workspace || n_labels < 0 || n_coeffs < 0)
    return set_error("qcut_reduce_expvals_launch: bad argument"), QCUT_ERR_INVALID;
  return launch_reduce<double>(d_expvals, d_pubs, d_labels, n_labels, d_lookup_start, d_lookup, d_coeffs,
                               n_coeffs, n_obs, d_out, d_workspace, (flags & QCUT_REDUCE_SINGLE_SLOT) != 0,
                               (flags & QCUT_REDUCE_SEQUENTIAL) != 0, static_cast<cudaStream_t>(stream));
}

int qcut_expvals_literal_launch(const qcut_plan* plan, const qcut_schedule* schedule, const uint8_t* d_data,
                                double* d_expvals, int64_t n_expvals, void* stream) {
  if (!plan || !schedule) return set_error("qcut_expvals_literal_launch: plan or schedule is NULL"), QCUT_ERR_INVALID;
  if (n_expvals < 0 || (n_expvals > 0 && !d_expvals)) return set_error("qcut_expvals_literal_launch: bad output"), QCUT_ERR_INVALID;
  if (schedule->n_tiles > 0 && !d_data) return set_error("qcut_expvals_literal_launch: d_data is NULL"), QCUT_ERR_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  schedule->last_stream = s;
  schedule->launched = true;
  // PUBs without shots write nothing: their expectation values stay 0 like the reference's (its loop does not run)
  if (n_expvals > 0) QCUT_CUDA(cudaMemsetAsync(d_expvals, 0, sizeof(double) * (size_t)n_expvals, s));
  return launch_literal(plan, d_data, schedule->d_pubs, schedule->n_pubs, d_expvals, s);
}

int qcut_tallies_to_expvals_launch(const qcut_plan* plan, const qcut_pub_desc* d_pubs, int64_t n_pubs,
                                   const int64_t* d_tallies, double* d_expvals, void* stream) {
  if (!plan || n_pubs < 0 || (n_pubs > 0 && (!d_pubs || !d_tallies || !d_expvals)))
    return set_error("qcut_tallies_to_expvals_launch: bad argument"), QCUT_ERR_INVALID;
  return launch_tallies_to_expvals(plan, d_pubs, n_pubs, d_tallies, d_expvals, static_cast<cudaStream_t>(stream));
}

int qcut_quasi_launch(const qcut_pub_desc* d_pubs, int64_t n_pubs, const qcut_group_desc* d_groups,
                      const uint64_t* d_mask_limbs, const int64_t* d_dist_start, const uint64_t* d_obs_bits,
                      int n_limbs_obs, const uint64_t* d_qpd_bits, int n_limbs_qpd, const double* d_weights,
                      double* d_expvals, void* stream) {
  if (n_pubs < 0 || n_limbs_obs < 0 || n_limbs_qpd < 0) return set_error("qcut_quasi_launch: bad size"), QCUT_ERR_INVALID;
  return launch_quasi(d_pubs, n_pubs, d_groups, d_mask_limbs, d_dist_start, d_obs_bits, n_limbs_obs,
                      d_qpd_bits, n_limbs_qpd, d_weights, d_expvals, static_cast<cudaStream_t>(stream));
}

int qcut_recombine_launch(const double* d_term_expvals, int expvals_complex, const uint32_t* d_row_start,
                          const uint32_t* d_term_index, const double* d_coeffs, int n_rows, double* d_out,
                          void* stream) {
  if (n_rows < 0 || (n_rows > 0 && (!d_row_start || !d_out)))
    return set_error("qcut_recombine_launch: bad argument"), QCUT_ERR_INVALID;
  return launch_recombine(d_term_expvals, expvals_complex != 0, d_row_start, d_term_index, d_coeffs, n_rows, d_out,
                          static_cast<cudaStream_t>(stream));
}

int qcut_gather_host(const void* const* h_src, const uint64_t* h_bytes, const uint64_t* h_dst_offset,
                     int64_t n_arrays, void* h_dst, int n_threads) {
  if (n_arrays < 0 || (n_arrays > 0 && (!h_src || !h_bytes || !h_dst_offset || !h_dst)))
    return set_error("qcut_gather_host: bad argument"), QCUT_ERR_INVALID;
  if (n_arrays == 0) return QCUT_OK;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, n_arrays);
  // Dynamic chunks of arrays so that ragged sizes balance across workers.
  std::atomic<int64_t> next{0};
  const int64_t chunk = std::max<int64_t>(1, n_arrays / (n_threads * 16));
  auto worker = [&]() {
    for (;;) {
      const int64_t lo = next.fetch_add(chunk);
      if (lo >= n_arrays) return;
      const int64_t hi = std::min(n_arrays, lo + chunk);
      for (int64_t i = lo; i < hi; ++i)
        if (h_bytes[i]) std::memcpy(static_cast<char*>(h_dst) + h_dst_offset[i], h_src[i], h_bytes[i]);
    }
  };
  if (n_threads == 1) {
    worker();
    return QCUT_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(n_threads);
  for (int i = 0; i < n_threads; ++i) pool.emplace_back(worker);
  for (auto& th : pool) th.join();
  return QCUT_OK;
}

}  // extern "C"

// ---- PUB de-duplication: open-addressing hash table over the six attributes of a PUB ------------------
// The table holds (hash, index) pairs only (16 bytes per slot); the attributes of every distinct PUB are
// appended to `keys` and compared when the hashes match.  Slots of a batch of PUBs are prefetched before
// they are probed: at 4 x 10^5 PUBs the table is far larger than the caches.
struct qcut_dedup {
  struct Key { uint64_t obs, qpd; int64_t shots; uint32_t nb_obs, nb_qpd, variant, pad; int64_t index; };
  struct Slot { uint64_t hash; int64_t key; };  // key = position in `keys`, -1 = empty
  std::vector<Slot> table;
  std::vector<Key> keys;
  static uint64_t hash(const Key& k) {
    uint64_t h = k.obs * 0x9E3779B97F4A7C15ull;
    h ^= (k.qpd + 0xC2B2AE3D27D4EB4Full) * 0x165667B19E3779F9ull;
    h ^= ((uint64_t)k.shots << 32 | k.variant) * 0x27D4EB2F165667C5ull;
    h ^= ((uint64_t)k.nb_obs << 16 | k.nb_qpd) * 0x85EBCA77C2B2AE63ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    return h ^ (h >> 32);
  }
  static bool same(const Key& a, const Key& b) {
    return a.obs == b.obs && a.qpd == b.qpd && a.shots == b.shots && a.nb_obs == b.nb_obs && a.nb_qpd == b.nb_qpd &&
           a.variant == b.variant;
  }
  void reserve(size_t n_keys) {
    size_t want = (size_t)1 << 12;
    while (want < 2 * n_keys + 2) want <<= 1;
    if (want <= table.size()) return;
    std::vector<Slot> old;
    old.swap(table);
    table.assign(want, Slot{0, -1});
    const size_t mask = want - 1;
    for (const Slot& e : old) {
      if (e.key < 0) continue;
      size_t i = (size_t)e.hash & mask;
      while (table[i].key >= 0) i = (i + 1) & mask;
      table[i] = e;
    }
  }
};

extern "C" {

int qcut_dedup_create(qcut_dedup** out) {
  if (!out) return set_error("qcut_dedup_create: out is NULL"), QCUT_ERR_INVALID;
  *out = new (std::nothrow) qcut_dedup();
  return *out ? QCUT_OK : (set_error("qcut_dedup_create: out of memory"), QCUT_ERR_INVALID);
}

void qcut_dedup_destroy(qcut_dedup* finder) { delete finder; }

void qcut_dedup_clear(qcut_dedup* finder) {
  if (!finder) return;
  finder->keys.clear();  // capacity (and the touched pages) are kept for the next call
  std::fill(finder->table.begin(), finder->table.end(), qcut_dedup::Slot{0, -1});
}

int64_t qcut_dedup_block(qcut_dedup* finder, int64_t base, const uint64_t* obs_ptr, const uint64_t* qpd_ptr,
                         const int64_t* shots, const int64_t* nb_obs, const int64_t* nb_qpd,
                         const uint32_t* variant, int64_t n, int64_t* rep) {
  if (!finder || n < 0 || (n > 0 && (!obs_ptr || !qpd_ptr || !shots || !nb_obs || !nb_qpd || !variant || !rep)))
    return set_error("qcut_dedup_block: bad argument"), -1;
  finder->reserve(finder->keys.size() + (size_t)n);  // no growth inside the loop
  finder->keys.reserve(finder->keys.size() + (size_t)n);
  const size_t mask = finder->table.size() - 1;
  constexpr int kBatch = 16;
  int64_t duplicates = 0;
  for (int64_t j0 = 0; j0 < n; j0 += kBatch) {
    const int m = (int)std::min<int64_t>(kBatch, n - j0);
    qcut_dedup::Key key[kBatch];
    uint64_t h[kBatch];
    for (int b = 0; b < m; ++b) {
      const int64_t j = j0 + b;
      key[b] = qcut_dedup::Key{obs_ptr[j], qpd_ptr[j], shots[j], (uint32_t)nb_obs[j], (uint32_t)nb_qpd[j], variant[j], 0u, base + j};
      h[b] = qcut_dedup::hash(key[b]);
      __builtin_prefetch(&finder->table[(size_t)h[b] & mask]);
    }
    for (int b = 0; b < m; ++b) {
      const int64_t j = j0 + b;
      rep[j] = base + j;
      if (shots[j] <= 0) continue;  // nothing to stage or tally anyway
      for (size_t i = (size_t)h[b] & mask;; i = (i + 1) & mask) {
        qcut_dedup::Slot& e = finder->table[i];
        if (e.key < 0) {
          e = qcut_dedup::Slot{h[b], (int64_t)finder->keys.size()};
          finder->keys.push_back(key[b]);
          break;
        }
        if (e.hash == h[b] && qcut_dedup::same(finder->keys[(size_t)e.key], key[b])) {
          rep[j] = finder->keys[(size_t)e.key].index;
          ++duplicates;
          break;
        }
      }
    }
  }
  return duplicates;
}

}  // extern "C"

namespace {
// Copy into the pinned ring.  streaming = true uses non-temporal stores: the ring is written once and
// read once by the DMA engine, so the destination lines need not be fetched or kept in the caches.
inline void gather_copy(char* dst, const void* src_v, size_t n, bool streaming) {
  if (!streaming || (reinterpret_cast<uintptr_t>(dst) & 15u)) {
    std::memcpy(dst, src_v, n);
    return;
  }
  const char* src = static_cast<const char*>(src_v);
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  if (i < n) std::memcpy(dst + i, src + i, n - i);
}
}  // namespace

extern "C" {

int qcut_stage_h2d(const void* const* h_src, const uint64_t* h_bytes, const uint64_t* h_dst_offset,
                   int64_t n_arrays, void* h_pinned, size_t pinned_bytes, uint8_t* d_dst, int n_threads,
                   void* stream) {
  if (n_arrays < 0 || (n_arrays > 0 && (!h_src || !h_bytes || !h_dst_offset || !h_pinned || !d_dst)))
    return set_error("qcut_stage_h2d: bad argument"), QCUT_ERR_INVALID;
  if (n_arrays == 0) return QCUT_OK;
  for (int64_t i = 1; i < n_arrays; ++i)
    if (h_dst_offset[i] < h_dst_offset[i - 1] + h_bytes[i - 1])
      return set_error("qcut_stage_h2d: arrays must be sorted by destination offset and must not overlap"), QCUT_ERR_INVALID;
  if (n_threads <= 0) n_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  size_t slot_bytes = (pinned_bytes / (size_t)n_threads) & ~(size_t)255;
  // slots of at least 512 KiB: smaller copies do not keep the PCIe link busy (QCUT_STAGING_MIN_SLOT_KB overrides)
  const char* env_slot = std::getenv("QCUT_STAGING_MIN_SLOT_KB");
  const size_t min_slot = (env_slot && *env_slot ? (size_t)std::max(4, std::atoi(env_slot)) : (size_t)512) << 10;
  while (n_threads > 1 && slot_bytes < min_slot) {
    --n_threads;
    slot_bytes = (pinned_bytes / (size_t)n_threads) & ~(size_t)255;
  }
  if (slot_bytes < 4096) return set_error("qcut_stage_h2d: pinned staging buffer too small"), QCUT_ERR_INVALID;

  // Cut the array list into chunks whose destination span fits one staging slot.
  struct Chunk { int64_t first, last; uint64_t base, span; };
  std::vector<Chunk> chunks;
  for (int64_t i = 0; i < n_arrays;) {
    Chunk c{i, i, h_dst_offset[i], h_bytes[i]};
    int64_t j = i + 1;
    while (j < n_arrays && h_dst_offset[j] + h_bytes[j] - c.base <= slot_bytes) {
      c.span = h_dst_offset[j] + h_bytes[j] - c.base;
      c.last = j++;
    }
    chunks.push_back(c);
    i = j;
  }
  int device = 0;
  QCUT_CUDA(cudaGetDevice(&device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Stores into the ring: a ring far larger than a last-level cache is written with non-temporal stores
  // (measured on cfg4, one rank per host: 292 ms per call vs 308); a small ring that several ranks keep
  // cache resident is written normally (two ranks on one host: 415 ms vs 460-473 with streaming stores).
  // QCUT_STAGING_NT=0/1 overrides.
  const char* env_nt = std::getenv("QCUT_STAGING_NT");
  const bool streaming = env_nt && *env_nt ? std::atoi(env_nt) != 0 : pinned_bytes >= ((size_t)64 << 20);
  std::atomic<int64_t> next{0};
  std::atomic<int> failed{0};
  std::vector<std::string> errors((size_t)n_threads);
  auto worker = [&](int tid) {
    if (cudaSetDevice(device) != cudaSuccess) { failed = 1; errors[tid] = "cudaSetDevice failed in staging thread"; return; }
    char* slot = static_cast<char*>(h_pinned) + (size_t)tid * slot_bytes;
    cudaEvent_t done = nullptr;
    if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess) { failed = 1; errors[tid] = "cudaEventCreate failed"; return; }
    bool in_flight = false;
    for (;;) {
      const int64_t ci = next.fetch_add(1);
      if (ci >= (int64_t)chunks.size() || failed.load()) break;
      const Chunk& c = chunks[(size_t)ci];
      cudaError_t err = cudaSuccess;
      if (c.span > slot_bytes) {
        // a single array larger than a slot: let the driver stage it (rare: > pinned_bytes / n_threads)
        err = cudaMemcpyAsync(d_dst + c.base, h_src[c.first], h_bytes[c.first], cudaMemcpyHostToDevice, s);
      } else {
        if (in_flight) err = cudaEventSynchronize(done);  // the slot's previous H2D has left the staging memory
        if (err == cudaSuccess) {
          for (int64_t i = c.first; i <= c.last; ++i)
            if (h_bytes[i]) gather_copy(slot + (h_dst_offset[i] - c.base), h_src[i], h_bytes[i], streaming);
          if (streaming) _mm_sfence();  // streaming stores are weakly ordered: make them visible before the DMA is queued
          err = cudaMemcpyAsync(d_dst + c.base, slot, c.span, cudaMemcpyHostToDevice, s);
        }
        if (err == cudaSuccess) err = cudaEventRecord(done, s);
        in_flight = true;
      }
      if (err != cudaSuccess) {
        failed = 1;
        errors[tid] = std::string("qcut_stage_h2d: ") + cudaGetErrorName(err) + " (" + cudaGetErrorString(err) + ")";
        break;
      }
    }
    if (in_flight) cudaEventSynchronize(done);
    cudaEventDestroy(done);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  if (failed.load()) {
    for (const auto& e : errors)
      if (!e.empty()) return set_error(e), QCUT_ERR_CUDA;
    return set_error("qcut_stage_h2d failed"), QCUT_ERR_CUDA;
  }
  return QCUT_OK;
}

size_t qcut_plan_kernel_source(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks, size_t mask_bytes,
                               int kernel, char* buf, size_t buf_len, int* kernel_of_group, int* variant_of_group) {
  if (!h_groups || n_groups <= 0) return 0;
  qcut_plan plan;
  plan.n_groups = n_groups;
  plan.groups.assign(h_groups, h_groups + n_groups);
  if (mask_bytes) plan.masks.assign(h_masks, h_masks + mask_bytes);
  plan.groups_dev.resize((size_t)n_groups);
  for (int g = 0; g < n_groups; ++g) {
    if ((size_t)h_groups[g].mask_offset + (size_t)h_groups[g].n_terms * h_groups[g].nb_obs > mask_bytes) return 0;
    plan.groups_dev[(size_t)g] = GroupDev{};
    plan.groups_dev[(size_t)g].kernel = sliced_rt_supports(h_groups[g].nb_obs) ? KERNEL_SLICED_RT : KERNEL_GENERIC;
  }
  assign_jit_kernels(&plan);
  if (jit_build(&plan, false) != QCUT_OK) return 0;  // (generates the sources; cubins come from the on-disk cache)
  for (int g = 0; g < n_groups; ++g) {
    if (kernel_of_group) kernel_of_group[g] = (int)plan.groups_dev[(size_t)g].kernel - (int)KERNEL_JIT0;  // < 0: ahead-of-time kernel
    if (variant_of_group) variant_of_group[g] = (int)plan.groups_dev[(size_t)g].variant;
  }
  if (kernel < 0 || kernel >= (int)plan.jit.size()) return 0;
  const std::string& src = plan.jit[(size_t)kernel].source;
  if (buf && buf_len) {
    const size_t n = src.size() < buf_len - 1 ? src.size() : buf_len - 1;
    std::memcpy(buf, src.data(), n);
    buf[n] = '\0';
  }
  return src.size() + 1;
}

int qcut_plan_compile_check(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks, size_t mask_bytes,
                            int* n_kernels, size_t* cubin_bytes) {
  if (!h_groups || n_groups <= 0) return set_error("qcut_plan_compile_check: no groups"), QCUT_ERR_INVALID;
  qcut_plan plan;
  plan.n_groups = n_groups;
  plan.groups.assign(h_groups, h_groups + n_groups);
  if (mask_bytes) plan.masks.assign(h_masks, h_masks + mask_bytes);
  plan.groups_dev.resize((size_t)n_groups);
  for (int g = 0; g < n_groups; ++g) {
    if ((size_t)h_groups[g].mask_offset + (size_t)h_groups[g].n_terms * h_groups[g].nb_obs > mask_bytes)
      return set_error("qcut_plan_compile_check: masks run past the mask buffer"), QCUT_ERR_INVALID;
    plan.groups_dev[(size_t)g] = GroupDev{};
    plan.groups_dev[(size_t)g].kernel = sliced_rt_supports(h_groups[g].nb_obs) ? KERNEL_SLICED_RT : KERNEL_GENERIC;
  }
  assign_jit_kernels(&plan);
  const int rc = jit_build(&plan, false);
  if (n_kernels) *n_kernels = (int)plan.jit.size();
  size_t total = 0;
  for (const JitKernel& jk : plan.jit) total += jk.cubin.size();
  if (cubin_bytes) *cubin_bytes = total;
  return rc;
}

}  // extern "C"

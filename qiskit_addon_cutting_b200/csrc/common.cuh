// Shared host/device declarations for libqcut_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "qcut_b200.h"

namespace qcut {

// Per-group record in HBM.
struct GroupDev {
  uint32_t n_terms;
  uint32_t nb_obs;
  uint32_t n_words;      // ceil(nb_obs / 4): 32-bit words per (zero padded) mask row
  uint32_t word_offset;  // index of the group's first mask word in mask_words[]
  uint32_t kernel;       // 0 = generic, 1 = bit-sliced run-time masks, >= 2 = NVRTC-specialised
  uint32_t variant;      // index of the group's code inside a multi-group specialised kernel
  uint32_t reserved[2];
};

struct TileEntry {
  uint32_t pub;
  uint32_t tile_in_pub;
};

enum : uint32_t { KERNEL_GENERIC = 0, KERNEL_SLICED_RT = 1, KERNEL_JIT0 = 2 };

void set_error(const std::string& msg);
int cuda_fail(cudaError_t err, const char* what);

#define QCUT_CUDA(call)                                               \
  do {                                                                \
    cudaError_t err__ = (call);                                       \
    if (err__ != cudaSuccess) return ::qcut::cuda_fail(err__, #call); \
  } while (0)

int sm_count();

// One NVRTC-specialised kernel (jit.cu).
struct JitKernel {
  std::string source;
  std::string log;
  std::vector<char> cubin;
  cudaLibrary_t library = nullptr;
  cudaKernel_t kernel = nullptr;
  int min_blocks = 2;
  int threads = 256;
  int group = -1;  // representative group (single-group kernels)
  std::vector<int> variants;  // representative group of every variant (multi-group kernels), empty otherwise
  int nb_obs = 0;
  bool staged = true;  // bulk-copy staged stream (shared-memory prefetch) vs direct global loads
  bool split = false;  // 8-byte rows processed one 32x32 block at a time (block 1 parked in shared memory)
  bool pipe = false;   // loads of the next step issued from inside the term network (qcut_warp_pipe)
  bool inline_switch = false;  // multi-group kernel of one row width: the groups' networks behind a switch (jit.cu)
  bool slab = false;   // rows of 16 / 24 / 32 bytes processed in 8-byte slabs (qcut_lane_tile_slab)
  bool sacc = false;   // packed counters in shared memory (qcut_lane_tile_sacc)
  int smem_bytes = 0;
};

}  // namespace qcut

struct qcut_plan {
  int n_groups = 0;
  unsigned flags = 0;
  std::vector<qcut_group_desc> groups;     // host copy
  std::vector<uint8_t> masks;              // host copy, big-endian rows
  std::vector<qcut::GroupDev> groups_dev;  // host mirror of the device table
  qcut::GroupDev* d_groups = nullptr;
  uint32_t* d_mask_words = nullptr;        // little-endian words assembled from the byte rows
  std::vector<qcut::JitKernel> jit;        // kernel id = KERNEL_JIT0 + index
  int n_kernels() const { return (int)qcut::KERNEL_JIT0 + (int)jit.size(); }
  // Ahead-of-time kernel of every group (run-time-mask or generic): what serves a group whose specialised kernel
  // is not (yet) there.  With QCUT_PLAN_ASYNC_JIT the NVRTC work runs on `jit_thread`; a schedule uses the
  // specialised kernels only if they were ready when it was created (jit_state, acquire/release).
  enum : int { JIT_NONE = 0, JIT_PENDING = 1, JIT_READY = 2, JIT_FAILED = -1 };
  std::vector<uint32_t> fallback_kernel;
  std::atomic<int> jit_state{JIT_NONE};
  std::atomic<bool> jit_cancel{false};
  std::thread jit_thread;
  std::string jit_error;
  uint32_t kernel_for(int group, bool use_jit) const {
    return use_jit ? groups_dev[(size_t)group].kernel : fallback_kernel[(size_t)group];
  }
  // Side streams of qcut_tally_launch (created on first use, used under `side_mutex`): the K1 kernels of one
  // launch write disjoint tallies, so they are spread over streams and the ramp of one overlaps the tail of another.
  static constexpr int kSideStreams = 3;
  mutable std::mutex side_mutex;
  mutable cudaStream_t side[kSideStreams] = {nullptr, nullptr, nullptr};
  mutable cudaEvent_t fork_event = nullptr;
  mutable cudaEvent_t join_event[kSideStreams] = {nullptr, nullptr, nullptr};
};

struct qcut_schedule {
  int64_t n_pubs = 0;
  int64_t n_tiles = 0;
  qcut_pub_desc* d_pubs = nullptr;
  qcut::TileEntry* d_tiles = nullptr;      // sorted by kernel id, PUB order within a kernel
  size_t pubs_capacity = 0, tiles_capacity = 0;  // bytes behind d_pubs / d_tiles (buffers come from a cache)
  int device = 0;
  std::vector<int64_t> kernel_tile_begin;  // n_kernels + 1 offsets into d_tiles
  std::vector<int64_t> kernel_bytes;       // shot bytes each kernel reads (decides whether launches overlap)
  // stream of the most recent qcut_tally_launch: qcut_schedule_destroy records an event there instead of
  // draining the device (the descriptors are recycled once that event has completed)
  mutable cudaStream_t last_stream = nullptr;
  mutable bool launched = false;
  bool use_jit = true;  // tiles were listed for the specialised kernels (else for the ahead-of-time ones)
};

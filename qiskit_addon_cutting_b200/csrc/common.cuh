// Shared device/host declarations for libqcut_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "qcut_b200.h"

namespace qcut {

// Per-group record in HBM.  `kernel` selects which K1 variant owns the group's tiles.
struct GroupDev {
  uint32_t n_terms;
  uint32_t nb_obs;
  uint32_t n_words;       // ceil(nb_obs / 4): 32-bit words per (zero padded) row
  uint32_t word_offset;   // index of the group's first mask word in mask_words[]
  uint32_t kernel;        // 0 = generic, 1 = bit-sliced specialised
  uint32_t variant;       // index of the specialised kernel body (switch case) when kernel == 1
  uint32_t reserved[2];
};

enum : uint32_t { KERNEL_GENERIC = 0, KERNEL_SLICED = 1 };

void set_error(const std::string& msg);
int cuda_fail(cudaError_t err, const char* what);

#define QCUT_CUDA(call)                                        \
  do {                                                         \
    cudaError_t err__ = (call);                                \
    if (err__ != cudaSuccess) return ::qcut::cuda_fail(err__, #call); \
  } while (0)

int sm_count();

}  // namespace qcut

struct qcut_plan {
  int n_groups = 0;
  unsigned flags = 0;
  std::vector<qcut_group_desc> groups;      // host copy
  std::vector<uint8_t> masks;               // host copy, big-endian rows
  std::vector<qcut::GroupDev> groups_dev;   // host mirror of the device table
  qcut::GroupDev* d_groups = nullptr;
  uint32_t* d_mask_words = nullptr;         // little-endian words assembled from the byte rows
  bool any_generic = false;
  bool any_sliced = false;
  std::string jit_source;                   // generated CUDA source ("" when nothing was specialised)
  std::string jit_log;
  void* jit = nullptr;                      // qcut::JitModule*
};

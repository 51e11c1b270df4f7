// Plan-time specialisation of the bit-sliced K1 kernel (NVRTC -> cubin -> cudaLibrary).
//
// For every commuting observable group whose row width the bit-sliced kernel covers, the
// generator emits one `struct TermsN` whose run() is the group's XOR network: the support of
// every Pauli mask (reference utils/observable_grouping.py:140-151) becomes a list of bit-plane
// *registers*, so a term costs ceil(w/2) LOP3 + 1 POPC + 1 IMAD per 32 shots with no mask loads,
// no indexing and no branches.  The hand-written kernel body lives in sliced_prelude.cuh.
#include <nvrtc.h>

#include <cstring>
#include <map>
#include <sstream>

#include "common.cuh"

namespace qcut {

static const char* const kPrelude =
#include "sliced_prelude_text.inc"
    ;

constexpr uint32_t kMaxSlicedTerms = 384;  // packed 8-bit counters: ceil(T/4) registers per lane

static bool sliced_width(uint32_t nb) { return nb == 1 || nb == 2 || nb == 4 || nb == 8; }

struct JitModule {
  cudaLibrary_t library = nullptr;
  cudaKernel_t kernel = nullptr;
  int min_blocks = 2;
};

// Name of the register holding plane (byte B, bit b) for NB-byte rows (see sliced_prelude.cuh).
static std::string plane_name(uint32_t nb, uint32_t byte, uint32_t bit) {
  std::ostringstream os;
  if (nb == 8 && byte >= 4) os << "Q[" << 8 * (byte - 4) + bit << "]";
  else os << "P[" << 8 * (byte % 4) + bit << "]";
  return os.str();
}

static void emit_terms(std::ostringstream& os, int variant, const qcut_group_desc& gd, const uint8_t* masks) {
  const uint32_t T = gd.n_terms, nb = gd.nb_obs;
  os << "// variant " << variant << ": " << T << " terms on " << nb << "-byte rows\n";
  os << "struct Terms" << variant << " {\n";
  os << "  static const int T = " << T << ";\n  static const int NACC = " << (T + 3) / 4 << ";\n";
  os << "  static const int NB = " << nb << ";\n";
  os << "  static QCUT_DEV void run(const uint32_t* __restrict__ P, const uint32_t* __restrict__ Q, "
        "const uint32_t qp, uint32_t* __restrict__ acc) {\n";
  os << "    (void)Q;\n";
  for (uint32_t t = 0; t < T; ++t) {
    const uint8_t* row = masks + gd.mask_offset + (size_t)t * nb;
    os << "    acc[" << t / 4 << "] += (uint32_t)__popc(qp";
    for (uint32_t byte = 0; byte < nb; ++byte)
      for (uint32_t bit = 0; bit < 8; ++bit)
        if ((row[byte] >> bit) & 1u) os << " ^ " << plane_name(nb, byte, bit);
    os << ")";
    if (t % 4) os << " << " << 8 * (t % 4);
    os << ";\n";
  }
  os << "  }\n};\n\n";
}

// Decide which groups are specialised and generate the translation unit.  Pure host code.
std::string generate_sliced_source(const qcut_group_desc* groups, int n_groups, const uint8_t* masks,
                                   std::vector<GroupDev>* groups_dev, int* n_variants_out, int* min_blocks_out) {
  std::ostringstream body;
  std::map<std::string, int> variant_of;  // identical (width, masks) share one variant
  std::vector<std::pair<uint32_t, uint32_t>> variant_shape;  // (nb, T)
  for (int g = 0; g < n_groups; ++g) {
    const qcut_group_desc& gd = groups[g];
    if (!sliced_width(gd.nb_obs) || gd.n_terms == 0 || gd.n_terms > kMaxSlicedTerms) continue;
    std::string key(reinterpret_cast<const char*>(&gd.nb_obs), sizeof(gd.nb_obs));
    key.append(reinterpret_cast<const char*>(masks + gd.mask_offset), (size_t)gd.n_terms * gd.nb_obs);
    auto it = variant_of.find(key);
    int variant;
    if (it == variant_of.end()) {
      variant = (int)variant_of.size();
      variant_of.emplace(std::move(key), variant);
      variant_shape.emplace_back(gd.nb_obs, gd.n_terms);
      emit_terms(body, variant, gd, masks);
    } else {
      variant = it->second;
    }
    if (groups_dev) {
      (*groups_dev)[g].kernel = KERNEL_SLICED;
      (*groups_dev)[g].variant = (uint32_t)variant;
    }
  }
  const int n_variants = (int)variant_shape.size();
  if (n_variants_out) *n_variants_out = n_variants;
  if (n_variants == 0) return std::string();

  // Register budget: 32 per block + packed counters + ~28 of addressing/temporaries.
  int min_blocks = 2;
  for (const auto& [nb, T] : variant_shape)
    if ((nb == 8 ? 64 : 32) + (int)((T + 3) / 4) + 28 > 128) min_blocks = 1;
  if (min_blocks_out) *min_blocks_out = min_blocks;

  std::ostringstream os;
  os << kPrelude << "\n" << body.str();
  os << "#ifndef QCUT_HOST_EMU\n";
  os << "extern \"C\" __global__ void __launch_bounds__(256, " << min_blocks << ")\n";
  os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const qcut_pub_desc* __restrict__ pubs,\n"
        "                  const uint32_t* __restrict__ tile_pub, const int64_t* __restrict__ tile_start,\n"
        "                  const int64_t n_tiles, const GroupDev* __restrict__ groups,\n"
        "                  unsigned long long* __restrict__ tallies) {\n"
        "  const int lane = threadIdx.x & 31;\n"
        "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
        "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles;\n"
        "       tile += n_warps) {\n"
        "    const uint32_t p = __ldg(tile_pub + tile);\n"
        "    const qcut_pub_desc pub = pubs[p];\n"
        "    const GroupDev grp = groups[pub.group];\n"
        "    if (grp.kernel != 1u) continue;  // tiles of the generic kernel\n"
        "    const int64_t shot0 = (tile - __ldg(tile_start + p)) * QCUT_TILE_SHOTS;\n"
        "    const int64_t left = (int64_t)pub.shots - shot0;\n"
        "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
        "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * grp.nb_obs;\n"
        "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
        "    unsigned long long* __restrict__ tally = tallies + pub.tally_offset;\n"
        "    switch (grp.variant) {\n";
  for (int v = 0; v < n_variants; ++v) {
    os << "      case " << v << ": {\n"
       << "        uint32_t acc[Terms" << v << "::NACC];\n"
       << "        qcut_lane_tile<Terms" << v << "::NB, Terms" << v << ">(obs, qpd, (int)pub.nb_qpd, ns, lane, acc);\n"
       << "        qcut_flush<Terms" << v << "::T>(acc, ns, lane, tally);\n"
       << "      } break;\n";
  }
  os << "      default: break;\n    }\n  }\n}\n";
  os << "#else  // host emulation entry: one lane of one tile, unpacked odd counts\n";
  os << "extern \"C\" int qcut_emu_lane_tile(int variant, const uint8_t* obs, const uint8_t* qpd, int nbq, int ns,\n"
        "                                  int lane, uint32_t* odd_out) {\n"
        "  switch (variant) {\n";
  for (int v = 0; v < n_variants; ++v) {
    os << "    case " << v << ": {\n"
       << "      uint32_t acc[Terms" << v << "::NACC];\n"
       << "      qcut_lane_tile<Terms" << v << "::NB, Terms" << v << ">(obs, qpd, nbq, ns, lane, acc);\n"
       << "      for (int t = 0; t < Terms" << v << "::T; ++t) odd_out[t] = (acc[t / 4] >> (8 * (t % 4))) & 0xFFu;\n"
       << "      return Terms" << v << "::T;\n    }\n";
  }
  os << "    default: return -1;\n  }\n}\n#endif\n";
  return os.str();
}

static std::string nvrtc_log(nvrtcProgram prog) {
  size_t n = 0;
  if (nvrtcGetProgramLogSize(prog, &n) != NVRTC_SUCCESS || n <= 1) return std::string();
  std::string log(n, '\0');
  nvrtcGetProgramLog(prog, log.data());
  return log;
}

// NVRTC: CUDA C++ -> sm_100a cubin.  Needs no GPU, so the CPU test-suite compiles every plan.
int compile_sliced_source(const std::string& source, std::vector<char>* cubin, std::string* log) {
  nvrtcProgram prog = nullptr;
  nvrtcResult rc = nvrtcCreateProgram(&prog, source.c_str(), "qcut_tally_sliced.cu", 0, nullptr, nullptr);
  if (rc != NVRTC_SUCCESS) {
    set_error(std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(rc));
    return QCUT_ERR_JIT;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device",
                        "--extra-device-vectorization"};
  rc = nvrtcCompileProgram(prog, 5, opts);
  if (log) *log = nvrtc_log(prog);
  if (rc != NVRTC_SUCCESS) {
    set_error(std::string("nvrtcCompileProgram: ") + nvrtcGetErrorString(rc) + "\n" + nvrtc_log(prog));
    nvrtcDestroyProgram(&prog);
    return QCUT_ERR_JIT;
  }
  size_t size = 0;
  rc = nvrtcGetCUBINSize(prog, &size);
  if (rc == NVRTC_SUCCESS && size > 0) {
    cubin->resize(size);
    rc = nvrtcGetCUBIN(prog, cubin->data());
  }
  nvrtcDestroyProgram(&prog);
  if (rc != NVRTC_SUCCESS || size == 0) {
    set_error(std::string("nvrtcGetCUBIN: ") + nvrtcGetErrorString(rc));
    return QCUT_ERR_JIT;
  }
  return QCUT_OK;
}

int jit_specialise(qcut_plan* plan) {
  int n_variants = 0, min_blocks = 2;
  plan->jit_source = generate_sliced_source(plan->groups.data(), plan->n_groups, plan->masks.data(),
                                            &plan->groups_dev, &n_variants, &min_blocks);
  if (n_variants == 0) return QCUT_OK;
  std::vector<char> cubin;
  int rc = compile_sliced_source(plan->jit_source, &cubin, &plan->jit_log);
  if (rc != QCUT_OK) return rc;
  JitModule* mod = new JitModule();
  mod->min_blocks = min_blocks;
  cudaError_t err = cudaLibraryLoadData(&mod->library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (err == cudaSuccess) err = cudaLibraryGetKernel(&mod->kernel, mod->library, "qcut_tally_sliced");
  if (err != cudaSuccess) {
    if (mod->library) cudaLibraryUnload(mod->library);
    delete mod;
    return cuda_fail(err, "cudaLibraryLoadData(qcut_tally_sliced)");
  }
  plan->jit = mod;
  return QCUT_OK;
}

void jit_release(qcut_plan* plan) {
  JitModule* mod = static_cast<JitModule*>(plan->jit);
  if (!mod) return;
  if (mod->library) cudaLibraryUnload(mod->library);
  delete mod;
  plan->jit = nullptr;
}

int launch_tally_sliced(const qcut_plan* plan, const uint8_t* d_data, const qcut_pub_desc* d_pubs,
                        const uint32_t* d_tile_pub, const int64_t* d_tile_start, int64_t n_tiles,
                        int64_t* d_tallies, cudaStream_t stream) {
  const JitModule* mod = static_cast<const JitModule*>(plan->jit);
  if (!mod || !mod->kernel) {
    set_error("qcut_tally_launch: plan has specialised groups but no loaded kernel");
    return QCUT_ERR_JIT;
  }
  // Persistent grid: min_blocks CTAs of 8 warps per SM, warps stride over tiles so that all SMs
  // sweep the PUB table together (same group in flight everywhere -> instruction-cache friendly).
  const int64_t want = (n_tiles + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * mod->min_blocks;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  const GroupDev* d_groups = plan->d_groups;
  unsigned long long* tallies = reinterpret_cast<unsigned long long*>(d_tallies);
  void* args[] = {(void*)&d_data,  (void*)&d_pubs,   (void*)&d_tile_pub, (void*)&d_tile_start,
                  (void*)&n_tiles, (void*)&d_groups, (void*)&tallies};
  QCUT_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(mod->kernel), dim3(grid), dim3(256), args, 0, stream));
  return QCUT_OK;
}

}  // namespace qcut

// ---- host-only entry points used by the CPU test-suite and for inspection ----------------------
extern "C" {

size_t qcut_generate_source(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks,
                            size_t mask_bytes, int* h_group_variant, char* buf, size_t buf_len) {
  if (!h_groups || n_groups <= 0) return 0;
  for (int g = 0; g < n_groups; ++g)
    if ((size_t)h_groups[g].mask_offset + (size_t)h_groups[g].n_terms * h_groups[g].nb_obs > mask_bytes) return 0;
  std::vector<qcut::GroupDev> dev(n_groups);
  for (auto& d : dev) d = qcut::GroupDev{};
  const std::string src = qcut::generate_sliced_source(h_groups, n_groups, h_masks, &dev, nullptr, nullptr);
  if (h_group_variant)
    for (int g = 0; g < n_groups; ++g)
      h_group_variant[g] = dev[g].kernel == qcut::KERNEL_SLICED ? (int)dev[g].variant : -1;
  if (buf && buf_len) {
    const size_t n = src.size() < buf_len - 1 ? src.size() : buf_len - 1;
    std::memcpy(buf, src.data(), n);
    buf[n] = '\0';
  }
  return src.size() + 1;
}

int qcut_compile_source(const char* source, size_t* cubin_bytes, char* cubin_buf, size_t cubin_buf_len) {
  if (!source) return qcut::set_error("qcut_compile_source: source is NULL"), QCUT_ERR_INVALID;
  std::vector<char> cubin;
  std::string log;
  const int rc = qcut::compile_sliced_source(source, &cubin, &log);
  if (rc != QCUT_OK) return rc;
  if (cubin_bytes) *cubin_bytes = cubin.size();
  if (cubin_buf && cubin_buf_len >= cubin.size()) std::memcpy(cubin_buf, cubin.data(), cubin.size());
  return QCUT_OK;
}

}  // extern "C"

// Plan-time specialisation of the bit-sliced K1 kernel (NVRTC -> cubin -> cudaLibrary).
//
// For a commuting observable group with many terms the generator emits `struct Terms` whose run()
// is the group's XOR network: the support of every Pauli mask (reference
// utils/observable_grouping.py:140-151) becomes a list of bit-plane *registers*, so a term costs
// ceil(w/2) LOP3 + 1 POPC + 1 IMAD per 32 shots with no mask loads, no indexing and no branches.
// The hand-written kernel body lives in sliced_prelude.cuh.  One kernel per specialised group;
// they are compiled concurrently and each streams its own tile list (api.cu).
#include <nvrtc.h>

#include <algorithm>
#include <atomic>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <sstream>
#include <thread>

#include "common.cuh"

namespace qcut {

static const char* const kPrelude =
#include "sliced_prelude_text.inc"
    ;

constexpr uint32_t kMaxJitTerms = 384;  // packed 8-bit counters: ceil(T/4) registers per lane
constexpr size_t kMaxSplitCross = 12;   // 8-byte rows are processed one block at a time while at most this many terms straddle the blocks

bool jit_supports(const qcut_group_desc& gd) {
  const uint32_t nb = gd.nb_obs;
  return nb >= 1 && nb <= 32 && gd.n_terms > 0 && gd.n_terms <= kMaxJitTerms;
}

// Name of the register holding plane (byte B, bit b) for NB-byte rows (see sliced_prelude.cuh).
static std::string plane_name(uint32_t nb, uint32_t byte, uint32_t bit) {
  std::ostringstream os;
  if (nb > 8) {  // wide rows: planes parked in shared memory, P = &planes[0][lane]
    os << "P[" << (8 * byte + bit) * 32 << "]";
    return os.str();
  }
  if (byte >= 4) os << "Q[" << 8 * (byte - 4) + bit << "]";
  else os << "P[" << 8 * (byte % 4) + bit << "]";
  return os.str();
}

// struct <name> { T, NACC, NB, run() }: the XOR network of one group, masks baked in.
// Terms are evaluated along a greedy nearest-neighbour chain in Hamming distance: the parity word of a
// term is either built from scratch (qp ^ its planes) or from the previous term's word XOR the planes
// in which the two masks differ, whichever needs fewer LOP3 -- dense or structured mask sets (long
// shared Z strings) share most of their work this way, at the cost of one live register.
namespace {
struct TermEmitter {
  std::ostringstream& os;
  const qcut_group_desc& gd;
  const uint8_t* masks;
  uint32_t T, nb, nacc;
  const uint8_t* row(uint32_t t) const { return masks + gd.mask_offset + (size_t)t * nb; }
  int weight(uint32_t t, uint32_t b0, uint32_t b1) const {
    int w = 0;
    for (uint32_t b = b0; b < b1; ++b) w += __builtin_popcount(row(t)[b]);
    return w;
  }
  int dist(uint32_t s, uint32_t t, uint32_t b0, uint32_t b1) const {
    int d = 0;
    for (uint32_t b = b0; b < b1; ++b) d += __builtin_popcount((unsigned)(row(s)[b] ^ row(t)[b]));
    return d;
  }
  void count(uint32_t t, const char* word, bool popc = true) {
    // term t lives in field t / NACC of packed register t % NACC: consecutive terms touch different
    // registers, so their popcount -> add chains are independent
    os << "    acc[" << t % nacc << "] += " << (popc ? "(uint32_t)__popc(" : "(") << word << ")";
    if (t / nacc) os << " << " << 8 * (t / nacc);
    os << ";\n";
  }
  // Emit the terms of `which` restricted to mask bytes [b0, b1), chained in nearest-neighbour Hamming order.
  // start_of(t): the expression a from-scratch evaluation starts with ("qp" or a kept partial word).  Chaining
  // from the previous term is only allowed when both start from "qp" (chain_ok).  finish(t, "x") emits what
  // happens to the finished word (count it, or keep it for the second half of the step).
  template <class Start, class Finish>
  void chain(const std::vector<uint32_t>& which, uint32_t b0, uint32_t b1, bool chain_ok, Start start_of, Finish finish,
             const std::function<std::string(uint32_t, uint32_t)>& plane) {
    std::vector<bool> done(T, false);
    uint32_t prev = 0;
    bool have_prev = false;
    for (size_t step = 0; step < which.size(); ++step) {
      uint32_t best = T;
      int best_cost = 1 << 30;
      bool best_inc = false;
      for (uint32_t t : which) {
        if (done[t]) continue;
        const int direct = weight(t, b0, b1);
        const int inc = (have_prev && chain_ok) ? dist(prev, t, b0, b1) : (1 << 29);
        const int cost = std::min(direct, inc);
        if (cost < best_cost) { best_cost = cost; best = t; best_inc = inc < direct; }
      }
      done[best] = true;
      const uint8_t* cur = row(best);
      const uint8_t* old = have_prev ? row(prev) : cur;
      os << "    x = " << (best_inc ? std::string("x") : start_of(best));
      for (uint32_t byte = b0; byte < b1; ++byte)
        for (uint32_t bit = 0; bit < 8; ++bit) {
          const unsigned sel = best_inc ? (unsigned)(cur[byte] ^ old[byte]) : (unsigned)cur[byte];
          if ((sel >> bit) & 1u) os << " ^ " << plane(byte, bit);
        }
      os << ";\n";
      finish(best, "x");
      prev = best;
      have_prev = true;
    }
  }
};
}  // namespace

// Returns true when the split form (run0 / run1) was emitted.
static bool emit_terms(std::ostringstream& os, const std::string& name, const qcut_group_desc& gd, const uint8_t* masks,
                       bool* pipe_out = nullptr, bool* sacc_out = nullptr, bool* il_out = nullptr, bool* slab_out = nullptr) {
  const uint32_t T = gd.n_terms, nb = gd.nb_obs, nacc = (T + 3) / 4;
  TermEmitter em{os, gd, masks, T, nb, nacc};
  std::vector<uint32_t> all(T);
  for (uint32_t t = 0; t < T; ++t) all[t] = t;
  os << "// " << T << " terms on " << nb << "-byte rows\n";
  os << "struct " << name << " {\n";
  os << "  static const int T = " << T << ";\n  static const int NACC = " << nacc << ";\n";
  os << "  static const int NB = " << nb << ";\n";
  os << "  static QCUT_MDEV void run(const uint32_t* __restrict__ P, const uint32_t* __restrict__ Q, "
        "const uint32_t qp, uint32_t* __restrict__ acc) {\n";
  os << "    (void)Q;\n    uint32_t x;\n";
  em.chain(all, 0, nb, true, [](uint32_t) { return std::string("qp"); },
           [&](uint32_t t, const char* w) { em.count(t, w); },
           [&](uint32_t byte, uint32_t bit) { return plane_name(nb, byte, bit); });
  os << "  }\n";
  // 8-byte rows, one block at a time (qcut_lane_tile_split): block-0 terms and the block-0 half of the cross
  // terms in run0 (their partial words wait in K[]), block-1 terms and the rest of the cross terms in run1.
  // Only worth it while few terms straddle the two halves of the row.
  std::vector<uint32_t> lo, hi, cross;
  if (nb == 8) {
    for (uint32_t t = 0; t < T; ++t) {
      const bool in_lo = em.weight(t, 0, 4) > 0, in_hi = em.weight(t, 4, 8) > 0;
      (in_lo && in_hi ? cross : in_hi ? hi : lo).push_back(t);  // identity terms count with block 0
    }
  }
  const bool split = nb == 8 && cross.size() <= kMaxSplitCross;
  os << "  static const int SPLIT = " << (split ? 1 : 0) << ";\n";
  os << "  static const int NKEEP = " << (split ? cross.size() : 0) << ";\n";
  if (split) {
    std::vector<int> slot(T, -1);
    for (size_t i = 0; i < cross.size(); ++i) slot[cross[i]] = (int)i;
    auto p0 = [&](uint32_t byte, uint32_t bit) { return "P[" + std::to_string(8 * byte + bit) + "]"; };
    auto p1 = [&](uint32_t byte, uint32_t bit) { return "P[" + std::to_string(8 * (byte - 4) + bit) + "]"; };
    os << "  static QCUT_MDEV void run0(const uint32_t* __restrict__ P, const uint32_t qp, uint32_t* __restrict__ acc, "
          "uint32_t* __restrict__ K) {\n    (void)K;\n    uint32_t x;\n";
    em.chain(lo, 0, 4, true, [](uint32_t) { return std::string("qp"); },
             [&](uint32_t t, const char* w) { em.count(t, w); }, p0);
    em.chain(cross, 0, 4, true, [](uint32_t) { return std::string("qp"); },
             [&](uint32_t t, const char* w) { os << "    K[" << slot[t] << "] = " << w << ";\n"; }, p0);
    os << "  }\n";
    os << "  static QCUT_MDEV void run1(const uint32_t* __restrict__ P, const uint32_t qp, uint32_t* __restrict__ acc, "
          "const uint32_t* __restrict__ K) {\n    (void)K;\n    uint32_t x;\n";
    em.chain(hi, 4, 8, true, [](uint32_t) { return std::string("qp"); },
             [&](uint32_t t, const char* w) { em.count(t, w); }, p1);
    em.chain(cross, 4, 8, false, [&](uint32_t t) { return "K[" + std::to_string(slot[t]) + "]"; },
             [&](uint32_t t, const char* w) { em.count(t, w); }, p1);
    os << "  }\n";
  }
  // 8-byte rows, loads of the next step inside the network (qcut_warp_pipe): terms ordered so that bit-planes are
  // finished as early as possible -- repeatedly take the live plane with the fewest unevaluated terms and evaluate
  // those (min-degree elimination) -- and one 16-byte load of the next step after every four planes that died.
  // Every term is built from scratch (no chaining), so this form is only offered for sparse masks (what a local
  // observable restricted to a subsystem looks like); `pipe_out` says whether it was emitted.
  bool pipe = false;
  if (nb == 8 && pipe_out) {
    size_t lop3 = 0;
    for (uint32_t t = 0; t < T; ++t) lop3 += (size_t)(em.weight(t, 0, 8) + 1) / 2;
    std::vector<std::vector<uint32_t>> of_plane(64);
    std::vector<int> open_terms(64, 0);
    for (uint32_t t = 0; t < T; ++t)
      for (uint32_t p = 0; p < 64; ++p)
        if ((em.row(t)[p >> 3] >> (p & 7)) & 1u) { of_plane[p].push_back(t); ++open_terms[p]; }
    std::vector<uint32_t> order;
    std::vector<int> dead_after;  // planes finished once order[i] has been evaluated
    std::vector<bool> done(T, false);
    int dead = 0;
    for (uint32_t p = 0; p < 64; ++p) dead += open_terms[p] == 0;
    auto evaluate = [&](uint32_t t) {
      done[t] = true;
      for (uint32_t p = 0; p < 64; ++p)
        if ((em.row(t)[p >> 3] >> (p & 7)) & 1u) dead += --open_terms[p] == 0;
      order.push_back(t);
      dead_after.push_back(dead);
    };
    for (uint32_t t = 0; t < T; ++t)
      if (em.weight(t, 0, 8) == 0) evaluate(t);
    while (order.size() < T) {
      int best = -1;
      for (int p = 0; p < 64; ++p)
        if (open_terms[p] > 0 && (best < 0 || open_terms[p] < open_terms[best])) best = p;
      for (uint32_t t : of_plane[(size_t)best])
        if (!done[t]) evaluate(t);
    }
    // worth it when the masks are sparse and half of the planes are finished after three quarters of the terms
    pipe = lop3 <= (size_t)T * 2 && dead_after[(size_t)(3 * T / 4 > 0 ? 3 * T / 4 - 1 : 0)] >= 32;
    if (pipe) {
      os << "  static QCUT_MDEV void run_pf(const uint32_t* __restrict__ P, const uint32_t* __restrict__ Q, const uint32_t qp,\n"
            "                               uint32_t* __restrict__ acc, const uint8_t* __restrict__ nsrc, const uint32_t pfv,\n"
            "                               uint32_t* __restrict__ NA) {\n    (void)Q;\n    uint32_t x, c = 0;\n";
      int issued = 0;
      {  // planes no term reads are free from the start
        int dead0 = 0;
        for (uint32_t p = 0; p < 64; ++p) dead0 += of_plane[p].empty();
        while (issued < 16 && 4 * (issued + 1) <= dead0) os << "    QCUT_PF_LOAD(" << issued++ << ", c);\n";
      }
      for (size_t i = 0; i < order.size(); ++i) {
        const uint32_t t = order[i];
        os << "    x = qp";
        for (uint32_t byte = 0; byte < 8; ++byte)
          for (uint32_t bit = 0; bit < 8; ++bit)
            if ((em.row(t)[byte] >> bit) & 1u) os << " ^ " << plane_name(nb, byte, bit);
        os << ";\n    c = (uint32_t)__popc(x);\n";
        em.count(t, "c", false);
        while (issued < 16 && 4 * (issued + 1) <= dead_after[i]) os << "    QCUT_PF_LOAD(" << issued++ << ", c);\n";
      }
      while (issued < 16) os << "    QCUT_PF_LOAD(" << issued++ << ", c);\n";
      os << "  }\n";
    }
  }
  // Counters in shared memory (qcut_lane_tile_sacc): packed register r = t % NACC, field t / NACC as everywhere else;
  // the sixteen terms of registers 4g .. 4g+3 are evaluated together between one 128-bit load and one store of group g.
  // Terms are built from scratch, so only offered for sparse masks on rows of up to 8 bytes.
  bool sacc = false;
  if (nb <= 8 && sacc_out) {
    size_t lop3 = 0;
    for (uint32_t t = 0; t < T; ++t) lop3 += (size_t)(em.weight(t, 0, nb) + 1) / 2;
    sacc = lop3 <= (size_t)T * 2;
    if (sacc) {
      os << "  static QCUT_MDEV void run_sacc(const uint32_t* __restrict__ P, const uint32_t* __restrict__ Q, const uint32_t qp,\n"
            "                                 const QcutSacc S) {\n    (void)Q;\n    uint32_t x;\n    qcut_u4 q;\n";
      static const char* const word[4] = {"q.x", "q.y", "q.z", "q.w"};
      for (uint32_t g = 0; 4 * g < nacc; ++g) {
        os << "    q = S.ld4(" << g << ");\n";
        for (uint32_t f = 0; f < 4; ++f)
          for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t r = 4 * g + j, t = f * nacc + r;
            if (r >= nacc || t >= T) continue;
            os << "    x = qp";
            for (uint32_t byte = 0; byte < nb; ++byte)
              for (uint32_t bit = 0; bit < 8; ++bit)
                if ((em.row(t)[byte] >> bit) & 1u) os << " ^ " << plane_name(nb, byte, bit);
            os << ";\n    " << word[j] << " += (uint32_t)__popc(x)";
            if (f) os << " << " << 8 * f;
            os << ";\n";
          }
        os << "    S.st4(" << g << ", q);\n";
      }
      os << "  }\n";
    }
  }
  // 8-byte rows, sparse masks: Q arrives as the raw ROWS of block 1 and its transposition (80 pair operations, ALU and
  // FMA pipes) is spread between the terms that live in block 0 (XU pipe); the terms of block 1 and the few terms that
  // straddle the blocks follow (qcut_lane_tile_il).
  bool il = false;
  if (nb == 8 && il_out) {
    size_t lop3 = 0;
    for (uint32_t t = 0; t < T; ++t) lop3 += (size_t)(em.weight(t, 0, nb) + 1) / 2;
    std::vector<uint32_t> t0, t1;
    for (uint32_t t = 0; t < T; ++t) (em.weight(t, 4, 8) == 0 ? t0 : t1).push_back(t);
    il = lop3 <= (size_t)T * 2 && t0.size() >= 16;
    if (il) {
      os << "  static QCUT_MDEV void run_il(const uint32_t* __restrict__ P, uint32_t* __restrict__ Q, const uint32_t qp,\n"
            "                               uint32_t* __restrict__ acc) {\n    uint32_t x;\n";
      std::vector<std::string> tp;  // the 80 pair operations of the transposition of Q, stage by stage
      for (int d : {16, 8, 4, 2, 1})
        for (int j = 0; j < 32; ++j)
          if (!(j & d)) tp.push_back("    qcut_tp<" + std::to_string(d) + ">(Q, " + std::to_string(j) + ");\n");
      auto term = [&](uint32_t t) {
        os << "    x = qp";
        for (uint32_t byte = 0; byte < 8; ++byte)
          for (uint32_t bit = 0; bit < 8; ++bit)
            if ((em.row(t)[byte] >> bit) & 1u) os << " ^ " << plane_name(nb, byte, bit);
        os << ";\n";
        em.count(t, "x");
      };
      size_t done = 0;
      for (size_t i = 0; i < t0.size(); ++i) {
        term(t0[i]);
        const size_t upto = (i + 1) * tp.size() / t0.size();
        for (; done < upto; ++done) os << tp[done];
      }
      for (uint32_t t : t1) term(t);
      os << "  }\n";
    }
  }
  // Rows of 16 / 24 / 32 bytes, sparse masks: the row is processed in 8-byte slabs (qcut_lane_tile_slab), counters in
  // shared memory.  A term that lives in one slab is evaluated in that slab's pass; a term that straddles slabs starts a
  // partial word K[i] in the first slab it touches and is counted in the last.  Packed counter r = t % NACC, field
  // t / NACC as everywhere; a pass loads / stores only the counter groups it adds to.
  bool slab = false;
  if (slab_out && (nb == 16 || nb == 24 || nb == 32)) {
    const uint32_t n_slabs = nb / 8;
    size_t lop3 = 0;
    std::vector<uint32_t> first(T, n_slabs), last(T, 0);
    std::vector<int> kslot(T, -1);
    int nk = 0;
    for (uint32_t t = 0; t < T; ++t) {
      lop3 += (size_t)(em.weight(t, 0, nb) + 1) / 2;
      for (uint32_t sl = 0; sl < n_slabs; ++sl)
        if (em.weight(t, 8 * sl, 8 * sl + 8) > 0) { first[t] = std::min(first[t], sl); last[t] = sl; }
      if (first[t] == n_slabs) first[t] = last[t] = 0;  // identity: counted with slab 0
      if (first[t] != last[t]) kslot[t] = nk++;
    }
    slab = lop3 <= (size_t)T * 2 + (size_t)nk && nk <= 16;
    if (slab) {
      os << "  static const int NK = " << nk << ";\n";
      os << "  template <int SLAB>\n"
            "  static QCUT_MDEV void run_slab(const uint32_t* __restrict__ P, const uint32_t* __restrict__ Q, const uint32_t qp,\n"
            "                                 const QcutSacc S, uint32_t* __restrict__ K) {\n"
            "    (void)P; (void)Q; (void)K;\n    uint32_t x;\n    qcut_u4 q;\n";
      static const char* const word[4] = {"q.x", "q.y", "q.z", "q.w"};
      auto planes_of = [&](uint32_t t, uint32_t sl) {
        std::string e;
        for (uint32_t byte = 8 * sl; byte < 8 * sl + 8; ++byte)
          for (uint32_t bit = 0; bit < 8; ++bit)
            if ((em.row(t)[byte] >> bit) & 1u)
              e += std::string(" ^ ") + ((byte % 8) >= 4 ? "Q[" : "P[") + std::to_string(8 * (byte % 4) + bit) + "]";
        return e;
      };
      for (uint32_t sl = 0; sl < n_slabs; ++sl) {
        os << "    " << (sl ? "else " : "") << "if constexpr (SLAB == " << sl << ") {\n";
        // partial words of the straddling terms that pass through this slab without ending here
        for (uint32_t t = 0; t < T; ++t)
          if (kslot[t] >= 0 && sl >= first[t] && sl < last[t] && (sl == first[t] || em.weight(t, 8 * sl, 8 * sl + 8) > 0))
            os << "      K[" << kslot[t] << "] = " << (sl == first[t] ? "qp" : "K[" + std::to_string(kslot[t]) + "]")
               << planes_of(t, sl) << ";\n";
        for (uint32_t g = 0; 4 * g < nacc; ++g) {
          bool any = false;
          for (uint32_t f = 0; f < 4 && !any; ++f)
            for (uint32_t j = 0; j < 4 && !any; ++j) {
              const uint32_t r = 4 * g + j, t = f * nacc + r;
              any = r < nacc && t < T && last[t] == sl;
            }
          if (!any) continue;
          os << "      q = S.ld4(" << g << ");\n";
          for (uint32_t f = 0; f < 4; ++f)
            for (uint32_t j = 0; j < 4; ++j) {
              const uint32_t r = 4 * g + j, t = f * nacc + r;
              if (r >= nacc || t >= T || last[t] != sl) continue;
              os << "      x = " << (kslot[t] >= 0 ? "K[" + std::to_string(kslot[t]) + "]" : std::string("qp")) << planes_of(t, sl)
                 << ";\n      " << word[j] << " += (uint32_t)__popc(x)";
              if (f) os << " << " << 8 * f;
              os << ";\n";
            }
          os << "      S.st4(" << g << ", q);\n";
        }
        os << "    }\n";
      }
      os << "  }\n";
    }
  }
  os << "  static const int SLAB = " << (slab ? 1 : 0) << ";\n";
  if (slab_out) *slab_out = slab;
  os << "  static const int IL = " << (il ? 1 : 0) << ";\n";
  if (il_out) *il_out = il;
  os << "  static const int SACC = " << (sacc ? 1 : 0) << ";\n";
  if (sacc_out) *sacc_out = sacc;
  os << "  static const int PIPE = " << (pipe ? 1 : 0) << ";\n";
  if (pipe_out) *pipe_out = pipe;
  os << "};\n\n";
  return split;
}

// One kernel for many small groups: every group becomes a __noinline__ device function (so that
// compile time stays linear in the number of groups) and the kernel dispatches on the group's
// variant index; the schedule lists a kernel's tiles variant by variant, so at any moment the SMs
// execute a handful of functions.
std::string generate_multi_source(const std::vector<const qcut_group_desc*>& gds, const uint8_t* masks, int threads,
                                  int min_blocks) {
  std::ostringstream os;
  os << kPrelude << "\n";
  for (size_t v = 0; v < gds.size(); ++v) {
    emit_terms(os, "Terms" + std::to_string(v), *gds[v], masks);
    os << "__device__ __noinline__ void qcut_variant_" << v
       << "(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns, int lane,\n"
          "                                    unsigned long long* __restrict__ tally) {\n"
          "  uint32_t acc[Terms" << v << "::NACC];\n"
          "  qcut_lane_tile<Terms" << v << "::NB, Terms" << v << ">(obs, qpd, nbq, ns, lane, acc);\n"
          "  qcut_flush<Terms" << v << "::T>(acc, ns, lane, tally);\n"
          "}\n\n";
  }
  os << "extern \"C\" __global__ void __launch_bounds__(" << threads << ", " << min_blocks << ")\n"
        "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
        "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
        "                  unsigned long long* __restrict__ tallies, const QcutGroup* __restrict__ groups) {\n"
        "  const int lane = threadIdx.x & 31;\n"
        "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
        "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles;\n"
        "       tile += n_warps) {\n"
        "    const QcutTile te = tiles[tile];\n"
        "    const QcutPub pub = pubs[te.pub];\n"
        "    const QcutGroup grp = groups[pub.group];\n"
        "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
        "    const int64_t left = (int64_t)pub.shots - shot0;\n"
        "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
        "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * grp.nb_obs;\n"
        "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
        "    unsigned long long* __restrict__ tally = tallies + pub.tally_offset;\n"
        "    switch (grp.variant) {\n";
  for (size_t v = 0; v < gds.size(); ++v)
    os << "      case " << v << ": qcut_variant_" << v << "(obs, qpd, (int)pub.nb_qpd, ns, lane, tally); break;\n";
  os << "      default: break;\n    }\n  }\n}\n";
  return os.str();
}

// One kernel for many small groups OF ONE ROW WIDTH (at most 32 terms each): loads, transposition and qpd parity plane
// are written once, the groups' XOR networks (a handful of LOP3 / POPC / IMAD each) sit behind a switch on the variant
// index of the tile's group, and the flush takes the number of terms at run time.  Compared with one __noinline__
// function per group this keeps the code of the kernel small (ncu showed `no_instruction` stalls on the molecular
// workload's kernels: 24 complete copies of the tile loop per kernel) and lets the whole kernel run as one staged
// stream of steps (bulk copies one step ahead through shared memory), whatever group a tile belongs to.
std::string generate_switch_source(const std::vector<const qcut_group_desc*>& gds, const uint8_t* masks, int threads,
                                   int min_blocks, bool staged) {
  std::ostringstream os;
  const uint32_t nb = gds.front()->nb_obs;
  uint32_t max_nacc = 1;
  if (staged) os << "#define QCUT_WITH_VARIANTS 1\n";
  if (nb == 8) os << "#define QCUT_FOLD8 1\n";  // fold-and-pack qpd plane on 8-byte rows too (rows loaded by qcut_step_loads)
  os << kPrelude << "\n";
  for (size_t v = 0; v < gds.size(); ++v) {
    emit_terms(os, "Terms" + std::to_string(v), *gds[v], masks);
    max_nacc = std::max(max_nacc, (gds[v]->n_terms + 3) / 4);
  }
  os << "#ifndef QCUT_HOST_EMU\n__constant__\n#else\nstatic const\n#endif\n";
  os << "unsigned short qcut_variant_terms[" << gds.size() << "] = {";
  for (size_t v = 0; v < gds.size(); ++v) os << (v ? ", " : "") << gds[v]->n_terms;
  os << "};\n";
  os << "struct Net {\n"
        "  static const int NB = " << nb << ";\n"
        "  static const int NACC = " << max_nacc << ";\n"
        "  static const bool MULTI = true;\n"
        "  static QCUT_MDEV void run(const uint32_t* __restrict__ P, const uint32_t* __restrict__ Q, const uint32_t qp,\n"
        "                            uint32_t* __restrict__ acc, const int variant) {\n"
        "    switch (variant) {\n";
  for (size_t v = 0; v < gds.size(); ++v) os << "      case " << v << ": Terms" << v << "::run(P, Q, qp, acc); break;\n";
  os << "      default: break;\n    }\n  }\n"
        "#ifndef QCUT_HOST_EMU\n"
        "  static QCUT_MDEV void flush(const uint32_t* acc, int ns, int lane, unsigned long long* __restrict__ tally, const int variant) {\n"
        "    const int T = qcut_variant_terms[variant];\n"
        "    qcut_flush_rt<NACC>(acc, T, (T + 3) >> 2, ns, lane, tally);\n"
        "  }\n"
        "#endif\n"
        "};\n"
        "#ifndef QCUT_HOST_EMU\n";
  os << "extern \"C\" __global__ void __launch_bounds__(" << threads << ", " << min_blocks << ")\n"
        "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
        "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
        "                  unsigned long long* __restrict__ tallies, const QcutGroup* __restrict__ groups) {\n";
  if (staged) {
    os << "  extern __shared__ __align__(128) uint8_t qcut_smem[];\n"
          "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n"
          "  const int n_warps_cta = blockDim.x >> 5;\n"
          "  uint8_t* stage = qcut_smem + (size_t)warp * QcutStage<Net::NB>::BYTES;\n"
          "  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(\n"
          "      qcut_smem + (size_t)n_warps_cta * QcutStage<Net::NB>::BYTES) + warp;\n"
          "  qcut_warp_stream<Net::NB, Net>(data, pubs, tiles, n_tiles, (int64_t)blockIdx.x * n_warps_cta + warp,\n"
          "                                 (int64_t)gridDim.x * n_warps_cta, tallies, stage, mbar, lane, groups);\n"
          "}\n";
  } else {
    os << "  const int lane = threadIdx.x & 31;\n"
          "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
          "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles;\n"
          "       tile += n_warps) {\n"
          "    const QcutTile te = tiles[tile];\n"
          "    const QcutPub pub = pubs[te.pub];\n"
          "    const int variant = (int)groups[pub.group].variant;\n"
          "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "    const int64_t left = (int64_t)pub.shots - shot0;\n"
          "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
          "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Net::NB;\n"
          "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
          "    uint32_t acc[Net::NACC];\n"
          "    qcut_lane_tile_net<Net::NB, Net>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc, variant);\n"
          "    Net::flush(acc, ns, lane, tallies + pub.tally_offset, variant);\n"
          "  }\n"
          "}\n";
  }
  os << "#else  // host emulation entry: one lane of one tile of a group served by network `variant`, unpacked odd counts\n"
        "extern \"C\" int qcut_emu_lane_tile_variant(int variant, const uint8_t* obs, const uint8_t* qpd, int nbq, int ns, int lane,\n"
        "                                          uint32_t* odd_out) {\n"
        "  uint32_t acc[Net::NACC];\n"
        "  qcut_lane_tile_net<Net::NB, Net>(obs, qpd, nbq, ns, lane, acc, variant);\n"
        "  const int T = qcut_variant_terms[variant], nacc = (T + 3) >> 2;\n"
        "  for (int t = 0; t < T; ++t) odd_out[t] = (acc[t % nacc] >> (8 * (t / nacc))) & 0xFFu;\n"
        "  return T;\n"
        "}\n"
        "#endif\n";
  return os.str();
}

// The translation unit for one group.  Pure host code.
std::string generate_sliced_source(const qcut_group_desc& gd, const uint8_t* masks, int* min_blocks_out,
                                   bool staged, int threads, int prefetch, bool allow_split, bool* split_out,
                                   bool allow_pipe, bool* pipe_out, bool allow_sacc, bool* sacc_out, bool* slab_out) {
  if (!jit_supports(gd)) return std::string();
  const uint32_t T = gd.n_terms, nb = gd.nb_obs;
  std::ostringstream terms;
  bool pipe = false, sacc = false, il = false;
  const bool plain = !staged && prefetch == 0 && !allow_split;
  const char* env_il = std::getenv("QCUT_JIT_IL");
  const bool want_il = plain && env_il && *env_il && std::atoi(env_il) != 0;
  const bool want_pipe = allow_pipe && plain && !allow_sacc && !want_il;
  const bool want_sacc = allow_sacc && plain && !want_il;
  // rows of 16 / 24 / 32 bytes with sparse masks: 8-byte slabs on the narrow-row engine (QCUT_JIT_SLAB=0: the
  // shared-memory wide-row engine for these widths too)
  const char* env_slab = std::getenv("QCUT_JIT_SLAB");
  const bool want_slab = !(env_slab && *env_slab && std::atoi(env_slab) == 0);
  bool slab = false;
  const bool split = emit_terms(terms, "Terms", gd, masks, want_pipe ? &pipe : nullptr, want_sacc ? &sacc : nullptr,
                                want_il ? &il : nullptr, want_slab ? &slab : nullptr) &&
                     allow_split && !staged && prefetch == 0;
  if (slab_out) *slab_out = slab;
  if (split_out) *split_out = split;
  if (pipe_out) *pipe_out = pipe;
  if (sacc_out) *sacc_out = sacc;
  // Register budget: 32 per live block + packed counters + ~28 of addressing / temporaries.
  const int budget = ((nb > 4 && !split) ? 64 : 32) + (int)((T + 3) / 4) + 28;
  int min_blocks = budget > 128 ? 1 : 2;
  if (slab || sacc) min_blocks = 2;  // packed counters in shared memory: 64 planes + temporaries fit 128 registers
  if (split && threads * 3 * (budget + 8) <= 65536) min_blocks = 3;  // one block at a time: three CTAs share an SM
  if (!split && nb <= 8) {
    // experiment knob (measured slower, see the multi-group kernels in jit_build): more CTAs per SM for small budgets
    const char* env_auto = std::getenv("QCUT_JIT_AUTO_MINBLOCKS");
    if (env_auto && *env_auto && std::atoi(env_auto) != 0) min_blocks = budget <= 64 ? 4 : budget <= 84 ? 3 : min_blocks;
  }
  if (min_blocks_out && *min_blocks_out > 0) min_blocks = *min_blocks_out;  // explicit override
  if (min_blocks_out) *min_blocks_out = min_blocks;

  std::ostringstream os;
  if (staged || prefetch > 0) os << "#define QCUT_WITH_VARIANTS 1\n";
  if (split) {
    const char* env_ph = std::getenv("QCUT_JIT_SPLIT_PHASES");
    const int phases = env_ph && *env_ph ? std::atoi(env_ph) : 1;
    os << "#define QCUT_SPLIT_PHASES " << (phases == 2 || phases == 4 ? phases : 1) << "\n";
  }
  os << "#define QCUT_SACC_STRIDE " << 16 * threads << "\n";
  {
    // 8-byte rows: fold-and-pack qpd plane (42 instead of 88 instructions) in the engines that load their rows through
    // qcut_step_loads; QCUT_JIT_FOLD8=0 switches it off
    const char* env_f8 = std::getenv("QCUT_JIT_FOLD8");
    if (nb == 8 && !split && !pipe && !il && !(env_f8 && *env_f8 && std::atoi(env_f8) == 0)) os << "#define QCUT_FOLD8 1\n";
  }
  if (slab) {
    const char* env_pf = std::getenv("QCUT_JIT_SLAB_PREFETCH");
    // measured on 16-byte rows at 2 CTAs per SM: plain 0.557 of the copy peak, L2 prefetch of the next step 0.594, L2
    // eviction hints (evict-last until the last slab has read a sector) 0.621, both 0.617 -> hints on, prefetch off
    if (env_pf && *env_pf && std::atoi(env_pf) != 0) os << "#define QCUT_SLAB_PREFETCH 1\n";
    const char* env_hint = std::getenv("QCUT_JIT_SLAB_L2_HINTS");
    if (!(env_hint && *env_hint && std::atoi(env_hint) == 0)) os << "#define QCUT_SLAB_L2_HINTS 1\n";
  }
  os << kPrelude << "\n";
  os << terms.str();
  if (slab) {
    os << "#ifndef QCUT_HOST_EMU\n"
          "extern \"C\" __global__ void __launch_bounds__(" << threads << ", 2)\n"
          "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  extern __shared__ __align__(16) uint32_t qcut_smem[];\n"
          "  const int lane = threadIdx.x & 31;\n"
          "  const QcutSacc S = {(uint32_t)__cvta_generic_to_shared(qcut_smem) + 16u * threadIdx.x};\n"
          "  qcut_sacc_clear<Terms::NACC>(S);\n"
          "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
          "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles;\n"
          "       tile += n_warps) {\n"
          "    const QcutTile te = tiles[tile];\n"
          "    const QcutPub pub = pubs[te.pub];\n"
          "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "    const int64_t left = (int64_t)pub.shots - shot0;\n"
          "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
          "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Terms::NB;\n"
          "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
          "    qcut_lane_tile_slab<Terms::NB, Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, S);\n"
          "    uint32_t acc[Terms::NACC];\n"
          "    qcut_sacc_take<Terms::NACC>(S, acc);\n"
          "    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);\n"
          "  }\n"
          "}\n"
          "#else\n"
          "extern \"C\" int qcut_emu_lane_tile(const uint8_t* obs, const uint8_t* qpd, int nbq, int ns, int lane,\n"
          "                                  uint32_t* odd_out) {\n"
          "  static uint32_t column[4 * ((Terms::NACC + 3) / 4)];\n"
          "  const QcutSacc S = {column};\n"
          "  qcut_sacc_clear<Terms::NACC>(S);\n"
          "  qcut_lane_tile_slab<Terms::NB, Terms>(obs, qpd, nbq, ns, lane, S);\n"
          "  uint32_t acc[Terms::NACC];\n"
          "  qcut_sacc_take<Terms::NACC>(S, acc);\n"
          "  for (int t = 0; t < Terms::T; ++t) odd_out[t] = (acc[t % Terms::NACC] >> (8 * (t / Terms::NACC))) & 0xFFu;\n"
          "  return Terms::T;\n"
          "}\n"
          "extern \"C\" int qcut_emu_rt_lane_tile(int, const uint8_t*, const uint8_t*, int, int, int, const uint32_t*, int, int,\n"
          "                                     uint32_t*) { return -1; }\n"
          "#endif\n";
    return os.str();
  }
  if (nb > 8) {
    // Wide rows: 4 warps per CTA (2 for rows of more than 16 bytes), per-warp raw + plane buffers in dynamic
    // shared memory.
    os << "#ifndef QCUT_HOST_EMU\n"
          "extern \"C\" __global__ void __launch_bounds__(32 * QcutWide<Terms::NB>::WARPS, 2)\n"
          "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  extern __shared__ __align__(16) uint32_t qcut_smem[];\n"
          "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n"
          "  uint32_t* raw = qcut_smem + (size_t)warp * QcutWide<Terms::NB>::WARP_WORDS;\n"
          "  uint32_t* planes = raw + QcutWide<Terms::NB>::RAW_WORDS;\n"
          "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
          "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; tile < n_tiles; tile += n_warps) {\n"
          "    const QcutTile te = tiles[tile];\n"
          "    const QcutPub pub = pubs[te.pub];\n"
          "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "    const int64_t left = (int64_t)pub.shots - shot0;\n"
          "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
          "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Terms::NB;\n"
          "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
          "    uint32_t acc[Terms::NACC];\n"
          "    qcut_lane_tile_wide<Terms::NB, Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc, raw, planes);\n"
          "    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);\n"
          "  }\n"
          "}\n"
          "#else\n"
          "extern \"C\" int qcut_emu_lane_tile(const uint8_t* obs, const uint8_t* qpd, int nbq, int ns, int lane,\n"
          "                                  uint32_t* odd_out) {\n"
          "  static uint32_t raw[QcutWide<Terms::NB>::RAW_WORDS], planes[QcutWide<Terms::NB>::PLANE_WORDS];\n"
          "  uint32_t acc[Terms::NACC];\n"
          "  qcut_lane_tile_wide<Terms::NB, Terms>(obs, qpd, nbq, ns, lane, acc, raw, planes);\n"
          "  for (int t = 0; t < Terms::T; ++t) odd_out[t] = (acc[t % Terms::NACC] >> (8 * (t / Terms::NACC))) & 0xFFu;\n"
          "  return Terms::T;\n"
          "}\n"
          "extern \"C\" int qcut_emu_rt_lane_tile(int, const uint8_t*, const uint8_t*, int, int, int, const uint32_t*, int, int,\n"
          "                                     uint32_t*) { return -1; }\n"
          "#endif\n";
    return os.str();
  }
  const char* env_pair = std::getenv("QCUT_JIT_PAIR");
  const bool pair = plain && !sacc && !pipe && nb <= 8 && env_pair && *env_pair && std::atoi(env_pair) != 0;
  os << "#ifndef QCUT_HOST_EMU\n";
  os << "extern \"C\" __global__ void __launch_bounds__(" << threads << ", " << min_blocks << ")\n";
  if (split) {
    // One 32x32 block at a time; the raw words of block 1 wait in the lane's column of a per-warp parking area
    // (32 x 32 words of dynamic shared memory per warp).
    os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  extern __shared__ __align__(16) uint32_t qcut_smem[];\n"
          "  const int lane = threadIdx.x & 31;\n"
          "  uint32_t* __restrict__ park = qcut_smem + (threadIdx.x >> 5) * 1024 + lane;\n"
          "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
          "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles;\n"
          "       tile += n_warps) {\n"
          "    const QcutTile te = tiles[tile];\n"
          "    const QcutPub pub = pubs[te.pub];\n"
          "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "    const int64_t left = (int64_t)pub.shots - shot0;\n"
          "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
          "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Terms::NB;\n"
          "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
          "    uint32_t acc[Terms::NACC];\n"
          "    qcut_lane_tile_split<Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc, park);\n"
          "    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);\n"
          "  }\n"
          "}\n";
  } else if (sacc) {
    // Packed counters in the thread's column of a shared-memory array (16 * threads bytes per group of 4 counters).
    os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  extern __shared__ __align__(16) uint32_t qcut_smem[];\n"
          "  const int lane = threadIdx.x & 31;\n"
          "  const QcutSacc S = {(uint32_t)__cvta_generic_to_shared(qcut_smem) + 16u * threadIdx.x};\n"
          "  qcut_sacc_clear<Terms::NACC>(S);\n"
          "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
          "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles;\n"
          "       tile += n_warps) {\n"
          "    const QcutTile te = tiles[tile];\n"
          "    const QcutPub pub = pubs[te.pub];\n"
          "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "    const int64_t left = (int64_t)pub.shots - shot0;\n"
          "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
          "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Terms::NB;\n"
          "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
          "    qcut_lane_tile_sacc<Terms::NB, Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, S);\n"
          "    uint32_t acc[Terms::NACC];\n"
          "    qcut_sacc_take<Terms::NACC>(S, acc);\n"
          "    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);\n"
          "  }\n"
          "}\n";
  } else if (pipe) {
    // Register-level software pipeline: the loads of the next step (or of the first step of the warp's next tile)
    // are issued from inside the term network as bit-planes die (qcut_warp_pipe).
    os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  const int lane = threadIdx.x & 31;\n"
          "  qcut_warp_pipe<Terms>(data, pubs, tiles, n_tiles, (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5),\n"
          "                        (int64_t)gridDim.x * (blockDim.x >> 5), tallies, lane);\n"
          "}\n";
  } else if (pair) {
    // Experiment (QCUT_JIT_PAIR=1, off by default): a warp takes the tile list two entries at a time; two consecutive
    // tiles of one PUB share one update of the tallies (qcut_flush_gather / qcut_flush_commit), a plain store when
    // they are the whole PUB.  Bit-exact, but the five registers that carry the gathered sums across a tile push the
    // step loop over the 128-register cliff (STACK 64 -> 96): cfg4 0.764 -> 0.691 of the copy peak, cfg3 dense
    // 0.350 -> 0.320.
    os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  const int lane = threadIdx.x & 31;\n"
          "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
          "  for (int64_t tile = 2 * ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); tile < n_tiles;\n"
          "       tile += 2 * n_warps) {\n"
          "    uint32_t mine[(Terms::NACC + 7) / 8];\n"
          "#pragma unroll\n"
          "    for (int c = 0; c < (Terms::NACC + 7) / 8; ++c) mine[c] = 0;\n"
          "    QcutTile te = tiles[tile];\n"
          "    QcutPub pub = pubs[te.pub];\n"
          "    long long ns_sum = 0;\n"
          "    uint32_t first_tile = te.tile_in_pub;\n"
          "#pragma unroll 1\n"
          "    for (int h = 0; h < 2 && tile + h < n_tiles; ++h) {\n"
          "      if (h == 1) {\n"
          "        const QcutTile te2 = tiles[tile + 1];\n"
          "        if (te2.pub != te.pub) {  // the second entry belongs to another PUB: finish the first on its own\n"
          "          qcut_flush_commit<Terms::T>(mine, ns_sum, lane, tallies + pub.tally_offset,\n"
          "                                      first_tile == 0 && ns_sum == (long long)pub.shots);\n"
          "#pragma unroll\n"
          "          for (int c = 0; c < (Terms::NACC + 7) / 8; ++c) mine[c] = 0;\n"
          "          ns_sum = 0;\n"
          "          pub = pubs[te2.pub];\n"
          "          first_tile = te2.tile_in_pub;\n"
          "        }\n"
          "        te = te2;\n"
          "      }\n"
          "      const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "      const int64_t left = (int64_t)pub.shots - shot0;\n"
          "      const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
          "      const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Terms::NB;\n"
          "      const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
          "      uint32_t acc[Terms::NACC];\n";
    if (il)
      os << "      qcut_lane_tile_il<Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc);\n";
    else
      os << "      qcut_lane_tile<Terms::NB, Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc);\n";
    os << "      qcut_flush_gather<Terms::T>(acc, lane, mine);\n"
          "      ns_sum += ns;\n"
          "    }\n"
          "    qcut_flush_commit<Terms::T>(mine, ns_sum, lane, tallies + pub.tally_offset,\n"
          "                                first_tile == 0 && ns_sum == (long long)pub.shots);\n"
          "  }\n"
          "}\n";
  } else if (staged) {
    // Bulk-copy staged stream: one staging buffer + one mbarrier per warp in dynamic shared memory.
    os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  extern __shared__ __align__(128) uint8_t qcut_smem[];\n"
          "  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n"
          "  const int n_warps_cta = blockDim.x >> 5;\n"
          "  uint8_t* stage = qcut_smem + (size_t)warp * QcutStage<Terms::NB>::BYTES;\n"
          "  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(\n"
          "      qcut_smem + (size_t)n_warps_cta * QcutStage<Terms::NB>::BYTES) + warp;\n"
          "  qcut_warp_stream<Terms::NB, QcutSingleNet<Terms> >(data, pubs, tiles, n_tiles, (int64_t)blockIdx.x * n_warps_cta + warp,\n"
          "                                     (int64_t)gridDim.x * n_warps_cta, tallies, stage, mbar, lane, nullptr);\n"
          "}\n";
  } else {
    if (prefetch < 2) {
      // Direct loads, plain grid-stride loop over the tile list (prefetch == 1: L2-prefetch the next
      // step of the same tile).
      os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
            "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
            "                  unsigned long long* __restrict__ tallies) {\n"
            "  const int lane = threadIdx.x & 31;\n"
            "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
            "  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles;\n"
            "       tile += n_warps) {\n"
            "    const QcutTile te = tiles[tile];\n"
            "    const QcutPub pub = pubs[te.pub];\n"
            "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
            "    const int64_t left = (int64_t)pub.shots - shot0;\n"
            "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
            "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Terms::NB;\n"
            "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
            "    uint32_t acc[Terms::NACC];\n";
      if (prefetch == 1)
        os << "    qcut_lane_tile_pf<Terms::NB, Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc, obs, qpd, 1, 0);\n";
      else if (il)
        os << "    qcut_lane_tile_il<Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc);\n";
      else
        os << "    qcut_lane_tile<Terms::NB, Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc);\n";
      os << "    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);\n"
            "  }\n"
            "}\n";
    } else {
    // Direct loads; the tile entry and PUB descriptor of the warp's next tile are fetched one tile
    // ahead, and every step L2-prefetches the bytes of the step that follows it.
    os << "qcut_tally_sliced(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,\n"
          "                  const QcutTile* __restrict__ tiles, const int64_t n_tiles,\n"
          "                  unsigned long long* __restrict__ tallies) {\n"
          "  const int lane = threadIdx.x & 31;\n"
          "  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);\n"
          "  int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);\n"
          "  if (tile >= n_tiles) return;\n"
          "  QcutTile te = tiles[tile];\n"
          "  QcutPub pub = pubs[te.pub];\n"
          "  for (;;) {\n"
          "    const int64_t next = tile + n_warps;\n"
          "    const bool have_next = next < n_tiles;\n"
          "    QcutTile te_n = te;\n"
          "    QcutPub pub_n = pub;\n"
          "    if (have_next) { te_n = tiles[next]; pub_n = pubs[te_n.pub]; }\n"
          "    const int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "    const int64_t left = (int64_t)pub.shots - shot0;\n"
          "    const int ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;\n"
          "    const uint8_t* __restrict__ obs = data + pub.obs_offset + (uint64_t)shot0 * Terms::NB;\n"
          "    const uint8_t* __restrict__ qpd = data + pub.qpd_offset + (uint64_t)shot0 * pub.nb_qpd;\n"
          "    const int64_t shot0_n = (int64_t)te_n.tile_in_pub * QCUT_TILE_SHOTS;\n"
          "    const int64_t left_n = (int64_t)pub_n.shots - shot0_n;\n"
          "    const int ns_n = !have_next ? 0 : (left_n < QCUT_TILE_SHOTS ? (int)left_n : QCUT_TILE_SHOTS);\n"
          "    uint32_t acc[Terms::NACC];\n"
          "    qcut_lane_tile_pf<Terms::NB, Terms>(obs, qpd, (int)pub.nb_qpd, ns, lane, acc,\n"
          "        data + pub_n.obs_offset + (uint64_t)shot0_n * Terms::NB,\n"
          "        data + pub_n.qpd_offset + (uint64_t)shot0_n * pub_n.nb_qpd, (int)pub_n.nb_qpd, ns_n);\n"
          "    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);\n"
          "    if (!have_next) break;\n"
          "    tile = next; te = te_n; pub = pub_n;\n"
          "  }\n"
          "}\n";
    }
  }
  os << "#else  // host emulation entry: one lane of one tile, unpacked odd counts\n";
  os << "extern \"C\" int qcut_emu_lane_tile(const uint8_t* obs, const uint8_t* qpd, int nbq, int ns, int lane,\n"
        "                                  uint32_t* odd_out) {\n"
        "  uint32_t acc[Terms::NACC];\n";
  if (split)
    os << "  static uint32_t park[32 * 32];\n"
          "  qcut_lane_tile_split<Terms>(obs, qpd, nbq, ns, lane, acc, park + lane);\n";
  else if (il)
    os << "  qcut_lane_tile_il<Terms>(obs, qpd, nbq, ns, lane, acc);\n";
  else if (sacc)
    os << "  static uint32_t column[4 * ((Terms::NACC + 3) / 4)];\n"
          "  const QcutSacc S = {column};\n"
          "  qcut_sacc_clear<Terms::NACC>(S);\n"
          "  qcut_lane_tile_sacc<Terms::NB, Terms>(obs, qpd, nbq, ns, lane, S);\n"
          "  qcut_sacc_take<Terms::NACC>(S, acc);\n";
  else
    os << "  qcut_lane_tile<Terms::NB, Terms>(obs, qpd, nbq, ns, lane, acc);\n";
  os << "  for (int t = 0; t < Terms::T; ++t) odd_out[t] = (acc[t % Terms::NACC] >> (8 * (t / Terms::NACC))) & 0xFFu;\n"
        "  return Terms::T;\n"
        "}\n";
  // one lane of a whole warp stream (tiles first, first + stride, ...): tallies accumulate as on the device
  os << "extern \"C\" int qcut_emu_lane_stream(const uint8_t* data, const void* pubs, const void* tiles, long long n_tiles,\n"
        "                                    long long first, long long stride, int lane, unsigned long long* tallies) {\n";
  if (pipe)
    os << "  qcut_warp_pipe<Terms>(data, (const QcutPub*)pubs, (const QcutTile*)tiles, n_tiles, first, stride, tallies, lane);\n"
          "  return 1;\n";
  else
    os << "  return 0;\n";
  os << "}\n"
        "// run-time-mask engine, one lane of one tile (planes parked in a per-lane column)\n"
        "template <int NB> static int qcut_emu_rt(const uint8_t* obs, const uint8_t* qpd, int nbq, int ns, int lane,\n"
        "                                         const uint32_t* mw, int n_words, int T, uint32_t* odd_out) {\n"
        "  typedef QcutGeo<NB> G;\n"
        "  static uint32_t planes[64 * 32];\n"
        "  for (int t = 0; t < T; ++t) odd_out[t] = 0;\n"
        "  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {\n"
        "    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;\n"
        "    uint32_t a[G::BLOCKS][32]; uint32_t qp[G::BATCHES];\n"
        "    qcut_step_any<NB>(obs + (uint64_t)s0 * NB, qpd + (uint64_t)s0 * nbq, nbq, n, lane, a, qp);\n"
        "    for (int u = 0; u < G::BATCHES; ++u) {\n"
        "      for (int blk = 0; blk < G::BLOCKS; ++blk)\n"
        "        for (int i = 0; i < G::PLANES; ++i) planes[(blk * 32 + i) * 32 + lane] = a[blk][u * G::PLANES + i];\n"
        "      for (int t = 0; t < T; ++t)\n"
        "        odd_out[t] += qcut_rt_term(planes + lane, qp[u], mw[t * n_words], NB > 4 ? mw[t * n_words + 1] : 0u);\n"
        "    }\n"
        "  }\n"
        "  return T;\n"
        "}\n"
        "extern \"C\" int qcut_emu_rt_lane_tile(int nb, const uint8_t* obs, const uint8_t* qpd, int nbq, int ns, int lane,\n"
        "                                     const uint32_t* mw, int n_words, int T, uint32_t* odd_out) {\n"
        "  switch (nb) {\n"
        "    case 1: return qcut_emu_rt<1>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    case 2: return qcut_emu_rt<2>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    case 3: return qcut_emu_rt<3>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    case 4: return qcut_emu_rt<4>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    case 5: return qcut_emu_rt<5>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    case 6: return qcut_emu_rt<6>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    case 7: return qcut_emu_rt<7>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    case 8: return qcut_emu_rt<8>(obs, qpd, nbq, ns, lane, mw, n_words, T, odd_out);\n"
        "    default: return -1;\n"
        "  }\n"
        "}\n"
        "#endif\n";
  return os.str();
}

static std::string nvrtc_log(nvrtcProgram prog) {
  size_t n = 0;
  if (nvrtcGetProgramLogSize(prog, &n) != NVRTC_SUCCESS || n <= 1) return std::string();
  std::string log(n, '\0');
  nvrtcGetProgramLog(prog, log.data());
  return log;
}

// ---- on-disk cache of compiled kernels ---------------------------------------------------------
// A cubin is a pure function of (generated source, NVRTC version, options), and a plan of 332 groups
// costs ~15 s of NVRTC per process.  Cubins are kept under $QCUT_CACHE_DIR (default
// $XDG_CACHE_HOME/qcut_b200 or ~/.cache/qcut_b200; QCUT_CACHE_DIR="" switches the cache off), named by a
// 128-bit hash of everything they depend on, written to a temporary name and renamed (atomic on POSIX,
// so concurrent ranks never see half a file).  A file that is not an ELF image is ignored and replaced.
static const char* const kNvrtcOptions[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};

static std::string cache_dir() {
  if (const char* dir = std::getenv("QCUT_CACHE_DIR")) return std::string(dir);
  if (const char* xdg = std::getenv("XDG_CACHE_HOME"))
    if (*xdg) return std::string(xdg) + "/qcut_b200";
  if (const char* home = std::getenv("HOME"))
    if (*home) return std::string(home) + "/.cache/qcut_b200";
  return std::string();
}

static std::string cache_key(const std::string& source) {
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  std::string text = "nvrtc " + std::to_string(major) + "." + std::to_string(minor);
  for (const char* o : kNvrtcOptions) text += std::string(" ") + o;
  text += "\n" + source;
  uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x9ae16a3b2f90404full;  // two FNV-1a streams, different offsets / primes
  for (unsigned char c : text) {
    h1 = (h1 ^ c) * 0x100000001b3ull;
    h2 = (h2 ^ c) * 0x00000100000001b5ull + 0x9e3779b97f4a7c15ull;
  }
  char name[40];
  std::snprintf(name, sizeof name, "%016llx%016llx", (unsigned long long)h1, (unsigned long long)h2);
  return std::string(name);
}

static bool cache_load(const std::string& path, std::vector<char>* cubin) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<char> data;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n);
  std::fclose(f);
  if (data.size() < 64 || std::memcmp(data.data(), "\177ELF", 4) != 0) return false;
  cubin->swap(data);
  return true;
}

static void cache_store(const std::string& dir, const std::string& path, const std::vector<char>& cubin) {
  ::mkdir(dir.c_str(), 0700);  // one level; a missing parent simply leaves the cache off
  const std::string tmp = path + "." + std::to_string((long long)::getpid()) + "." +
                          std::to_string((unsigned long long)std::hash<std::thread::id>()(std::this_thread::get_id())) + ".tmp";
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  if (std::fclose(f) != 0 || !ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

static int compile_with_nvrtc(const std::string& source, std::vector<char>* cubin, std::string* log, std::string* error);

// Generated source -> sm_100a cubin, through the on-disk cache.  Needs no GPU, so the CPU test-suite
// compiles plans too.  Thread safe (no use of the thread-local error string): the message is returned in *error.
int compile_sliced_source(const std::string& source, std::vector<char>* cubin, std::string* log,
                          std::string* error) {
  const std::string dir = cache_dir();
  const std::string path = dir.empty() ? std::string() : dir + "/" + cache_key(source) + ".cubin";
  if (!path.empty() && cache_load(path, cubin)) {
    if (log) *log = "(cubin from " + path + ")";
    return QCUT_OK;
  }
  const int rc = compile_with_nvrtc(source, cubin, log, error);
  if (rc == QCUT_OK && !path.empty()) cache_store(dir, path, *cubin);
  return rc;
}

static int compile_with_nvrtc(const std::string& source, std::vector<char>* cubin, std::string* log,
                              std::string* error) {
  nvrtcProgram prog = nullptr;
  nvrtcResult rc = nvrtcCreateProgram(&prog, source.c_str(), "qcut_tally_sliced.cu", 0, nullptr, nullptr);
  if (rc != NVRTC_SUCCESS) {
    *error = std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(rc);
    return QCUT_ERR_JIT;
  }
  rc = nvrtcCompileProgram(prog, (int)(sizeof kNvrtcOptions / sizeof kNvrtcOptions[0]), kNvrtcOptions);
  if (log) *log = nvrtc_log(prog);
  if (rc != NVRTC_SUCCESS) {
    *error = std::string("nvrtcCompileProgram: ") + nvrtcGetErrorString(rc) + "\n" + nvrtc_log(prog);
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
    *error = std::string("nvrtcGetCUBIN: ") + nvrtcGetErrorString(rc);
    return QCUT_ERR_JIT;
  }
  return QCUT_OK;
}

// Generate + compile (concurrently) + load the kernels of plan->jit[*].group.
int jit_build(qcut_plan* plan, bool load) {
  const size_t n = plan->jit.size();
  if (n == 0) return QCUT_OK;
  std::vector<int> rcs(n, QCUT_OK);
  std::vector<std::string> errors(n);
  auto work = [&](size_t i) {
    JitKernel& jk = plan->jit[i];
    if (plan->jit_cancel.load()) {  // the plan is being destroyed while its background compilation runs
      rcs[i] = QCUT_ERR_JIT;
      errors[i] = "cancelled";
      return;
    }
    if (!jk.variants.empty() && jk.inline_switch) {
      std::vector<const qcut_group_desc*> gds;
      for (int g : jk.variants) gds.push_back(&plan->groups[g]);
      jk.threads = 256;
      jk.min_blocks = 2;
      if (const char* env_mb = std::getenv("QCUT_JIT_SWITCH_MINBLOCKS"))  // experiments
        if (*env_mb && std::atoi(env_mb) >= 1 && std::atoi(env_mb) <= 8) jk.min_blocks = std::atoi(env_mb);
      jk.nb_obs = (int)gds.front()->nb_obs;
      // these kernels wait on memory (a few terms per group): staged stream unless switched off
      const char* env_as = std::getenv("QCUT_JIT_AUTO_STAGED");
      jk.staged = (plan->flags & QCUT_PLAN_STAGED) != 0 || !(env_as && *env_as && std::atoi(env_as) == 0);
      jk.source = generate_switch_source(gds, plan->masks.data(), jk.threads, jk.min_blocks, jk.staged);
      rcs[i] = compile_sliced_source(jk.source, &jk.cubin, &jk.log, &errors[i]);
      return;
    }
    if (!jk.variants.empty()) {
      std::vector<const qcut_group_desc*> gds;
      for (int g : jk.variants) gds.push_back(&plan->groups[g]);
      jk.threads = 256;
      // Two CTAs per SM (128 registers).  These kernels wait on memory (ncu: long_scoreboard 4.5 warps per issue at 16
      // warps per SM) and their groups need only ~60 registers on paper, but more CTAs per SM measured SLOWER on B200
      // (cfg5: 0.73 of HBM at 2 CTAs/SM, 0.59 at 3, 0.56 at 4; cfg3 local 0.746 / 0.737 / 0.728): under a lower cap
      // ptxas spills the row registers it keeps in flight.  QCUT_JIT_MULTI_MINBLOCKS overrides (experiments).
      jk.min_blocks = 2;
      if (const char* env_mb = std::getenv("QCUT_JIT_MULTI_MINBLOCKS"))
        if (*env_mb && std::atoi(env_mb) >= 1 && std::atoi(env_mb) <= 8) jk.min_blocks = std::atoi(env_mb);
      jk.staged = false;
      jk.source = generate_multi_source(gds, plan->masks.data(), jk.threads, jk.min_blocks);
      rcs[i] = compile_sliced_source(jk.source, &jk.cubin, &jk.log, &errors[i]);
      return;
    }
    jk.nb_obs = (int)plan->groups[jk.group].nb_obs;
    // Bulk-copy staging (one step ahead through shared memory) pays where the kernel waits on memory instead of on
    // instruction issue: rows narrower than 8 bytes with few terms (cfg3 "local": 0.747 -> 0.782 of the copy peak;
    // cfg5 unchanged; cfg4, 8-byte rows x 134 terms: 7 % slower, so not there).  QCUT_JIT_AUTO_STAGED=0 switches the rule off.
    const char* env_as = std::getenv("QCUT_JIT_AUTO_STAGED");
    const bool auto_staged = !(env_as && *env_as && std::atoi(env_as) == 0) && jk.nb_obs < 8 &&
                             plan->groups[jk.group].n_terms <= 32;
    jk.staged = (plan->flags & QCUT_PLAN_STAGED) != 0 || auto_staged;
    // tuning knobs (experiments only): QCUT_JIT_THREADS (CTA size), QCUT_JIT_MINBLOCKS (launch bound)
    const char* env_threads = std::getenv("QCUT_JIT_THREADS");
    const char* env_blocks = std::getenv("QCUT_JIT_MINBLOCKS");
    jk.threads = env_threads ? std::atoi(env_threads) : 256;
    if (jk.threads < 32 || jk.threads > 1024 || jk.threads % 32) jk.threads = 256;
    jk.min_blocks = env_blocks ? std::atoi(env_blocks) : 0;
    const char* env_pf = std::getenv("QCUT_JIT_PREFETCH");
    const int prefetch = env_pf ? std::atoi(env_pf) : 0;
    const char* env_split = std::getenv("QCUT_JIT_SPLIT");
    // off by default: measured slower on B200 (cfg4: 0.71 of HBM at 256x2, 0.60-0.66 with 3-5 CTAs per SM vs 0.76 for
    // the all-at-once form -- the extra warps do not pay for the spills of the packed counters)
    const bool allow_split = env_split && *env_split && std::atoi(env_split) != 0;
    // Experimental engines, all measured on B200 and all within a few per cent BELOW the all-at-once engine on cfg4
    // (0.757 of the copy peak): QCUT_JIT_PIPE=1 loads of the next step issued from inside the term network (0.73: the
    // step still waits for its LAST load, ncu long_scoreboard 1.39 -> 1.73), QCUT_JIT_SACC=1 packed counters in shared
    // memory (0.726 at 16 warps/SM, 0.754 at 20, 0.738 at 24), QCUT_JIT_IL=1 transposition of block 1 interleaved with the
    // terms of block 0 (0.745).  scripts/micro/ holds the two measurements that explain it: the memory system delivers
    // 7.3 TB/s to this access pattern, and with the rows in L2 the kernel is only 5 % faster -- it is bound by
    // instruction issue at 16 warps per SM, not by HBM.  Off by default; kept with their emulation tests.
    const char* env_pipe = std::getenv("QCUT_JIT_PIPE");
    const bool allow_pipe = env_pipe && *env_pipe && std::atoi(env_pipe) != 0;
    // packed counters in shared memory (sparse masks, rows of up to 8 bytes); QCUT_JIT_SACC=1: on
    const char* env_sacc = std::getenv("QCUT_JIT_SACC");
    const bool allow_sacc = env_sacc && *env_sacc && std::atoi(env_sacc) != 0;
    jk.source = generate_sliced_source(plan->groups[jk.group], plan->masks.data(), &jk.min_blocks, jk.staged, jk.threads, prefetch,
                                       allow_split, &jk.split, allow_pipe, &jk.pipe, allow_sacc, &jk.sacc, &jk.slab);
    rcs[i] = compile_sliced_source(jk.source, &jk.cubin, &jk.log, &errors[i]);
  };
  std::atomic<size_t> next{0};
  auto drain = [&]() {
    for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) work(i);
  };
  const size_t n_threads = std::min<size_t>(n, std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
  std::vector<std::thread> pool;
  for (size_t t = 1; t < n_threads; ++t) pool.emplace_back(drain);
  drain();
  for (auto& th : pool) th.join();
  for (size_t i = 0; i < n; ++i)
    if (rcs[i] != QCUT_OK) {
      set_error(errors[i]);
      return rcs[i];
    }
  if (!load) return QCUT_OK;  // compile check only (no device needed)
  for (JitKernel& jk : plan->jit) {
    cudaError_t err = cudaLibraryLoadData(&jk.library, jk.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (err != cudaSuccess && jk.log.rfind("(cubin from ", 0) == 0) {
      // a cached image the driver rejects (e.g. damaged file): compile it again, which also replaces the file
      cudaGetLastError();
      std::string error;
      const std::string path = jk.log.substr(12, jk.log.size() - 13);
      std::remove(path.c_str());
      const int rc = compile_sliced_source(jk.source, &jk.cubin, &jk.log, &error);
      if (rc != QCUT_OK) return set_error(error), rc;
      err = cudaLibraryLoadData(&jk.library, jk.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    }
    if (err == cudaSuccess) err = cudaLibraryGetKernel(&jk.kernel, jk.library, "qcut_tally_sliced");
    if (err != cudaSuccess) return cuda_fail(err, "cudaLibraryLoadData(qcut_tally_sliced)");
    if (jk.variants.empty() && jk.slab) {
      // 8-byte slabs of a 16 / 24 / 32-byte row: 256 threads, 2 CTAs per SM, packed counters in shared memory
      jk.smem_bytes = 16 * jk.threads * (int)(((plan->groups[jk.group].n_terms + 3) / 4 + 3) / 4);
      int dev = 0;
      QCUT_CUDA(cudaGetDevice(&dev));
      QCUT_CUDA(cudaKernelSetAttributeForDevice(jk.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jk.smem_bytes, dev));
    } else if (jk.variants.empty() && jk.nb_obs > 8) {
      // wide rows: 4 (2 above 16 bytes) warps per CTA, per warp raw buffer (32 lanes x (4 NB + 1) words) + row /
      // plane store -- QcutWide<NB> in sliced_prelude.cuh
      const int blocks = (jk.nb_obs + 3) / 4;
      const int warps = jk.nb_obs > 16 ? 2 : 4;
      jk.threads = 32 * warps;
      jk.min_blocks = 2;
      jk.smem_bytes = warps * (32 * (4 * jk.nb_obs + 1) + blocks * 32 * 32) * 4;
      int dev = 0;
      QCUT_CUDA(cudaGetDevice(&dev));
      QCUT_CUDA(cudaKernelSetAttributeForDevice(jk.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jk.smem_bytes, dev));
    } else if (jk.sacc) {
      // one 16-byte slot per thread and group of 4 packed counters
      jk.smem_bytes = 16 * jk.threads * (int)(((plan->groups[jk.group].n_terms + 3) / 4 + 3) / 4);
      int dev = 0;
      QCUT_CUDA(cudaGetDevice(&dev));
      QCUT_CUDA(cudaKernelSetAttributeForDevice(jk.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jk.smem_bytes, dev));
    } else if (jk.split) {
      jk.smem_bytes = (jk.threads / 32) * 32 * 32 * 4;  // one 32 x 32-word parking area per warp
      int dev = 0;
      QCUT_CUDA(cudaGetDevice(&dev));
      QCUT_CUDA(cudaKernelSetAttributeForDevice(jk.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jk.smem_bytes, dev));
    } else if (jk.staged) {
      // per warp: one step of observable rows + one step of qpd bytes, plus one mbarrier
      const int step_shots = jk.nb_obs == 1 ? 4096 : jk.nb_obs == 2 ? 2048 : 1024;
      const int stage = step_shots * (jk.nb_obs + 1);
      jk.smem_bytes = (jk.threads / 32) * (stage + 8);
      int dev = 0;
      QCUT_CUDA(cudaGetDevice(&dev));
      QCUT_CUDA(cudaKernelSetAttributeForDevice(jk.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jk.smem_bytes, dev));
    }
    jk.cubin.clear();
    jk.cubin.shrink_to_fit();
  }
  return QCUT_OK;
}

void jit_release(qcut_plan* plan) {
  for (JitKernel& jk : plan->jit)
    if (jk.library) cudaLibraryUnload(jk.library);
  plan->jit.clear();
}

int launch_tally_jit(const JitKernel& jk, const uint8_t* d_data, const qcut_pub_desc* d_pubs,
                     const TileEntry* d_tiles, int64_t n_tiles, int64_t* d_tallies, const GroupDev* d_groups,
                     cudaStream_t stream) {
  if (n_tiles <= 0) return QCUT_OK;
  if (!jk.kernel) {
    set_error("qcut_tally_launch: specialised kernel was not loaded");
    return QCUT_ERR_JIT;
  }
  // Persistent grid: min_blocks CTAs of 8 warps per SM; warps stride over the tile list so that all
  // SMs sweep the PUB table together (DRAM pages and the instruction cache see one stream).
  const int warps = jk.threads / 32;
  const int64_t want = (n_tiles + warps - 1) / warps;
  const int64_t cap = (int64_t)sm_count() * jk.min_blocks;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  unsigned long long* tallies = reinterpret_cast<unsigned long long*>(d_tallies);
  void* args[] = {(void*)&d_data, (void*)&d_pubs, (void*)&d_tiles, (void*)&n_tiles, (void*)&tallies, (void*)&d_groups};
  QCUT_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(jk.kernel), dim3(grid), dim3(jk.threads), args,
                             (size_t)jk.smem_bytes, stream));
  return QCUT_OK;
}

}  // namespace qcut

// ---- host-only entry points used by the CPU test-suite and for inspection ----------------------
extern "C" {

size_t qcut_generate_source(const qcut_group_desc* h_groups, int n_groups, const uint8_t* h_masks,
                            size_t mask_bytes, int group, char* buf, size_t buf_len) {
  if (!h_groups || group < 0 || group >= n_groups) return 0;
  const qcut_group_desc& gd = h_groups[group];
  if ((size_t)gd.mask_offset + (size_t)gd.n_terms * gd.nb_obs > mask_bytes) return 0;
  const char* env_split = std::getenv("QCUT_JIT_SPLIT");
  const char* env_pipe = std::getenv("QCUT_JIT_PIPE");
  const char* env_sacc = std::getenv("QCUT_JIT_SACC");
  const std::string src = qcut::generate_sliced_source(gd, h_masks, nullptr, false, 256, 0,
                                                       env_split && *env_split && std::atoi(env_split) != 0, nullptr,
                                                       env_pipe && *env_pipe && std::atoi(env_pipe) != 0, nullptr,
                                                       env_sacc && *env_sacc && std::atoi(env_sacc) != 0, nullptr, nullptr);
  if (src.empty()) return 0;
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
  std::string log, error;
  const int rc = qcut::compile_sliced_source(source, &cubin, &log, &error);
  if (rc != QCUT_OK) return qcut::set_error(error), rc;
  if (cubin_bytes) *cubin_bytes = cubin.size();
  if (cubin_buf && cubin_buf_len >= cubin.size()) std::memcpy(cubin_buf, cubin.data(), cubin.size());
  return QCUT_OK;
}

}  // extern "C"

// Hand-written core of the bit-sliced parity-tally kernels (K1 fast paths, sm_100a).
//
// Compiled three ways from this one text:
//   * by nvcc into libqcut_b200.so for the run-time-mask kernel (tally_sliced.cu),
//   * by NVRTC at plan-creation time, followed by one generated `struct Terms` holding a group's
//     XOR network with the masks baked in as register names (jit.cu),
//   * by g++ with QCUT_HOST_EMU defined, where tests/ emulate one lane at a time against the oracle.
//
// Idea: a lane owns 32 "rows" (32-bit words of shot data, loaded with
// coalesced 128-bit loads), transposes the 32x32 bit block in registers so that register i holds
// bit i of 32 different shots, and then evaluates every Pauli term as the XOR of a few such
// bit-planes: ONE popcount per 32 shots per term instead of the reference's AND + popcount per
// shot per term (cutting_reconstruction.py:213-220).  Tallies are order independent, so any
// fixed shot <-> (lane, row) assignment is valid as long as the qpd parity plane uses the same one.
//
// Geometry of one step for NB-byte observable rows (each lane issues CHUNKS 16-byte loads):
//   NB = 8: 16 chunks, rows 2k+e of block 0/1 = low/high word of shot 2*(32k+lane)+e, 1024 shots
//   NB = 4:  8 chunks, row 4k+e = shot 4*(32k+lane)+e,                                 1024 shots
//   NB = 2:  8 chunks, row 4k+e holds shots 8*(32k+lane)+2e+{0,1}  (2 batches),        2048 shots
//   NB = 1:  8 chunks, row 4k+e holds shots 16*(32k+lane)+4e+{0..3} (4 batches),       4096 shots
//   NB = 3, 5, 6, 7 ("grouped" layout): a lane owns two groups of 16 consecutive shots (16*NB contiguous
//           bytes = NB 16-byte loads each, L1-allocating because neighbouring lanes share sectors); the rows
//           are cut out of the loaded words with one PRMT each; row 16g+r = shot 16*(32g+lane)+r, 1024 shots
// After the transposition, plane (byte B, bit b) of batch u is register 8*(B%4)+b + PLANES*u of
// block B/4, i.e. "memory bit" 8*B+b of the row: masks are used in memory order, never byte swapped.
#pragma once

#if defined(QCUT_HOST_EMU)
#define QCUT_DEV static inline
#define QCUT_MDEV inline
#elif defined(__CUDACC_RTC__)
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#define QCUT_DEV __device__ __forceinline__
#define QCUT_MDEV __device__ __forceinline__
#else
#include <stdint.h>
#define QCUT_DEV __device__ __forceinline__
#define QCUT_MDEV __device__ __forceinline__
#endif

#ifndef QCUT_TILE_SHOTS
#define QCUT_TILE_SHOTS 4096
#endif

// Device-side mirrors of include/qcut_b200.h (NVRTC has no access to that header).
struct QcutPub {
  uint64_t obs_offset, qpd_offset, tally_offset;
  uint32_t shots, group, nb_qpd, reserved;
};
struct QcutGroup {
  uint32_t n_terms, nb_obs, n_words, word_offset, kernel, variant, reserved[2];
};
struct QcutTile {
  uint32_t pub;          // PUB that owns the tile
  uint32_t tile_in_pub;  // shots [tile_in_pub * QCUT_TILE_SHOTS, ...) of that PUB
};

struct qcut_u4 { uint32_t x, y, z, w; };

#ifndef QCUT_HOST_EMU
// Streaming 128-bit load: shot bytes are read exactly once, keep them out of L1.
QCUT_DEV qcut_u4 qcut_ld16(const uint8_t* p) {
  qcut_u4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
QCUT_DEV qcut_u4 qcut_ld16_l1(const uint8_t* p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  qcut_u4 r = {v.x, v.y, v.z, v.w};
  return r;
}
QCUT_DEV uint32_t qcut_ld4(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
QCUT_DEV uint32_t qcut_ld2(const uint8_t* p) { return __ldg(reinterpret_cast<const uint16_t*>(p)); }
QCUT_DEV uint32_t qcut_ld1(const uint8_t* p) { return __ldg(p); }
#else
QCUT_DEV qcut_u4 qcut_ld16_l1(const uint8_t* p) { qcut_u4 v; memcpy(&v, p, 16); return v; }
QCUT_DEV qcut_u4 qcut_ld16(const uint8_t* p) { qcut_u4 v; memcpy(&v, p, 16); return v; }
QCUT_DEV uint32_t qcut_ld4(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
QCUT_DEV uint32_t qcut_ld2(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
QCUT_DEV uint32_t qcut_ld1(const uint8_t* p) { return *p; }
#endif

// ---- where a step's bytes come from -------------------------------------------------------------
// QcutSrcGlobal: straight from HBM (coalesced LDG.128, L1 no-allocate).
// QcutSrcShared: from the warp's staging buffer in shared memory, filled ahead of time by a bulk
//                asynchronous copy (cp.async.bulk -> UBLKCP, completion on an mbarrier), so the
//                HBM latency of step s+1 is hidden behind the arithmetic of step s.
// Offsets are bytes relative to the first observable / qpd byte of the step.
struct QcutSrcGlobal {
  const uint8_t* obs;
  const uint8_t* qpd;
  QCUT_MDEV qcut_u4 obs16(int off) const { return qcut_ld16(obs + off); }
  // grouped layout: neighbouring lanes' 16-byte pieces share 32-byte sectors across consecutive
  // loads, so these go through L1 (allocate) instead of bypassing it
  QCUT_MDEV qcut_u4 obs16_shared_sectors(int off) const { return qcut_ld16_l1(obs + off); }
  QCUT_MDEV qcut_u4 qpd16(int off) const { return qcut_ld16(qpd + off); }
  QCUT_MDEV uint32_t qpd4(int off) const { return qcut_ld4(qpd + off); }
  QCUT_MDEV uint32_t qpd2(int off) const { return qcut_ld2(qpd + off); }
};

#ifndef QCUT_HOST_EMU
struct QcutSrcShared {
  uint32_t obs;  // shared-window addresses
  uint32_t qpd;
  QCUT_MDEV qcut_u4 obs16(int off) const {
    qcut_u4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(obs + off));
    return v;
  }
  QCUT_MDEV qcut_u4 obs16_shared_sectors(int off) const { return obs16(off); }
  QCUT_MDEV qcut_u4 qpd16(int off) const {
    qcut_u4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(qpd + off));
    return v;
  }
  QCUT_MDEV uint32_t qpd4(int off) const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(qpd + off));
    return v;
  }
  QCUT_MDEV uint32_t qpd2(int off) const {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(qpd + off));
    return v;
  }
};
#else
typedef QcutSrcGlobal QcutSrcShared;  // the harness copies the step's bytes into a host buffer first
#endif

// Logical right shift by a constant on the FMA pipe (IMAD.HI) instead of the ALU pipe (SHF): the
// kernels are bound by ALU-pipe issue (LOP3 / PRMT), while the FMA pipe is nearly idle.
template <int S>
QCUT_DEV uint32_t qcut_shr(uint32_t x) { return __umulhi(x, 1u << (32 - S)); }

// Bit select (x & M) | (y & ~M) as ONE LOP3 (lut 0xE4) with M as the instruction's single immediate.
// Written in C, nvcc folds M and ~M into two immediates and emits two LOP3 per select.
template <uint32_t M>
QCUT_DEV uint32_t qcut_sel(uint32_t x, uint32_t y) {
#ifndef QCUT_HOST_EMU
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(x), "r"(y), "n"(M));
  return d;
#else
  return (x & M) | (y & ~M);
#endif
}

// In-register 32x32 bit-matrix transposition: afterwards bit r of a[i] is bit i of the old a[r].
// Two byte-granular stages are single PRMTs; the three bit-granular stages are one LOP3 select per
// output word with both shifts on the FMA pipe (IMAD.SHL / IMAD.HI).
QCUT_DEV void qcut_transpose32(uint32_t* a) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t x = a[j], y = a[j + 16];
    a[j] = __byte_perm(x, y, 0x5410);
    a[j + 16] = __byte_perm(x, y, 0x7632);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j & 8) continue;
    const uint32_t x = a[j], y = a[j + 8];
    a[j] = __byte_perm(x, y, 0x6240);
    a[j + 8] = __byte_perm(x, y, 0x7351);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j & 4) continue;
    const uint32_t x = a[j], y = a[j + 4];
    a[j] = qcut_sel<0x0F0F0F0Fu>(x, y << 4);
    a[j + 4] = qcut_sel<0x0F0F0F0Fu>(qcut_shr<4>(x), y);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j & 2) continue;
    const uint32_t x = a[j], y = a[j + 2];
    a[j] = qcut_sel<0x33333333u>(x, y << 2);
    a[j + 2] = qcut_sel<0x33333333u>(qcut_shr<2>(x), y);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j & 1) continue;
    const uint32_t x = a[j], y = a[j + 1];
    a[j] = qcut_sel<0x55555555u>(x, y << 1);
    a[j + 1] = qcut_sel<0x55555555u>(qcut_shr<1>(x), y);
  }
}

// Parity of each byte of v, delivered at bit 0 of that byte.
QCUT_DEV uint32_t qcut_byte_parity(uint32_t v) {
  v ^= qcut_shr<4>(v);
  v ^= qcut_shr<2>(v);
  v ^= qcut_shr<1>(v);
  return v & 0x01010101u;
}

template <int NB>
struct QcutGeo {
  static const bool GROUPED = !(NB == 1 || NB == 2 || NB == 4 || NB == 8);  // row width not a power of two
  static const int CHUNKS = GROUPED ? 2 * NB : (NB == 8) ? 16 : 8;   // 16-byte loads per lane per step
  static const int BLOCKS = (NB + 3) / 4;                             // 32x32 blocks per lane per step
  static const int BATCHES = (NB >= 3) ? 1 : 4 / NB;                  // 32-shot batches per block
  static const int SPC = GROUPED ? 16 : 16 / NB;                      // shots per chunk (per group when GROUPED)
  static const int RPC = (NB == 8) ? 2 : 4;                           // rows per chunk (power-of-two layouts)
  static const int STEP_SHOTS = GROUPED ? 1024 : CHUNKS * 32 * SPC;
  static const int PLANES = 32 / BATCHES;                             // registers per batch in one block
  static const int QRAW = (NB == 1) ? 32 : (NB == 2) ? 16 : 8;        // raw qpd words per lane per step
};

// Row register of the word that the load geometry above calls row j.  For 1- and 2-byte rows the qpd bytes are as many
// as (half as many as) the observable bytes, and folding them byte by byte cost as much as the transposition itself
// (ncu on the molecular workload's kernels after the switch kernels: ALU pipe 60 %, the top stalls `wait` and
// `math_pipe_throttle`).  The qpd parity planes of these widths come out of a fold-and-pack network instead (fold the
// high half of every field onto the low half, pack two words into one, three times: qcut_qpd_planes), whose output bits
// are in bit-reversed row order -- so the observable rows are simply NAMED in that order when they are loaded (row
// order is arbitrary as long as planes and qpd plane agree; tallies are order independent).  Involutions.
// (8-byte rows take part only in kernels whose generator defines QCUT_FOLD8: the experimental engines for that width
//  load their rows with code of their own)
#ifdef QCUT_FOLD8
#define QCUT_FOLD_NB(NB) ((NB) >= 3 && (NB) <= 8)
#else
#define QCUT_FOLD_NB(NB) ((NB) >= 3 && (NB) <= 7)
#endif
template <int NB>
QCUT_DEV constexpr int qcut_row_reg(int j) {
  if constexpr (NB == 2)  // j = (p, s3, s2, s1, b1) -> (p, b1, s1, s2, s3)
    return (j & 16) | ((j & 1) << 3) | ((j & 2) << 1) | ((j & 4) >> 1) | ((j & 8) >> 3);
  else if constexpr (NB == 1)  // j = (p1 p0, s3, s2, s1) -> (p1 p0, s1, s2, s3)
    return (j & 24) | ((j & 1) << 2) | (j & 2) | ((j & 4) >> 2);
  else if constexpr (QCUT_FOLD_NB(NB))  // qpd word i = j / 4 = (s3 s2 s1), byte b = j % 4 -> (b, s1, s2, s3)
    return ((j & 3) << 3) | (j & 4) | ((j & 8) >> 2) | ((j & 16) >> 4);
  else
    return j;
}
// ... and back (rows 3..7 bytes wide: not an involution)
template <int NB>
QCUT_DEV constexpr int qcut_reg_row(int r) {
  if constexpr (QCUT_FOLD_NB(NB))
    return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | (r >> 3);
  else
    return qcut_row_reg<NB>(r);
}

// Shot (relative to the step) held by row register j, sub-shot u of this lane.
template <int NB>
QCUT_DEV int qcut_shot_of(int jr, int u, int lane) {
  typedef QcutGeo<NB> G;
  const int j = qcut_reg_row<NB>(jr);
  if constexpr (G::GROUPED) return ((j >> 4) * 32 + lane) * 16 + (j & 15);
  const int k = j / G::RPC, e = j % G::RPC;
  return (k * 32 + lane) * G::SPC + e * G::BATCHES + u;
}

// qpd parity planes for arbitrary qpd row widths: one byte loop per shot.
template <int NB>
QCUT_DEV void qcut_qp_slow(const uint8_t* __restrict__ qpd_s, int nbq, int n, int lane, uint32_t* qp) {
  typedef QcutGeo<NB> G;
  for (int j = 0; j < 32; ++j) {
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u) {
      const int shot = qcut_shot_of<NB>(j, u, lane);
      if (shot < n) {
        uint32_t x = 0;
        for (int b = 0; b < nbq; ++b) x ^= qcut_ld1(qpd_s + (uint64_t)shot * nbq + b);
        qp[u] |= (uint32_t)(__popc(x) & 1) << j;
      }
    }
  }
}

// ---- the pieces of one step (n <= STEP_SHOTS shots) for one lane ------------------------------------
// FULL = true: the step holds exactly STEP_SHOTS shots -- no predication, every load is
// "lane base + immediate".  FULL = false (last step of a PUB): loads are predicated and rows past
// the end of the PUB are forced to zero afterwards (qcut_mask_tail); they then count as "even" and
// are removed again by tally = n - 2 * odd.

// (1) coalesced 128-bit loads of the observable rows into the row registers a[blk][j]
template <int NB, bool FULL, class Src>
QCUT_DEV void qcut_step_loads(const Src src, int n, int lane, uint32_t (*a)[32]) {
  typedef QcutGeo<NB> G;
  if constexpr (G::GROUPED) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      // 16 consecutive shots = 16 * NB contiguous bytes = NB 16-byte loads -> 4 * NB words
      uint32_t w[4 * NB + 1];
      const bool valid = FULL || (g * 32 + lane) * 16 < n;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        qcut_u4 v = {0u, 0u, 0u, 0u};
        if (valid) v = src.obs16_shared_sectors((g * 32 + lane) * 16 * NB + c * 16);
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
      }
      w[4 * NB] = 0;
      // cut the rows out: word b of shot r starts at byte r * NB + 4 * b of the group
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int blk = 0; blk < G::BLOCKS; ++blk) {
          const int o = r * NB + 4 * blk;
          const int wi = o >> 2, sh = o & 3;
          a[blk][qcut_row_reg<NB>(16 * g + r)] = sh == 0 ? w[wi] : __byte_perm(w[wi], w[wi + 1], 0x3210 + 0x1111 * sh);
        }
    }
  } else {
#pragma unroll
    for (int k = 0; k < G::CHUNKS; ++k) {
      qcut_u4 v = {0u, 0u, 0u, 0u};
      if (FULL || (k * 32 + lane) * G::SPC < n) v = src.obs16(lane * 16 + k * 512);
      if constexpr (NB == 8) {
        a[0][qcut_row_reg<NB>(2 * k)] = v.x; a[1][qcut_row_reg<NB>(2 * k)] = v.y;
        a[0][qcut_row_reg<NB>(2 * k + 1)] = v.z; a[1][qcut_row_reg<NB>(2 * k + 1)] = v.w;
      } else {
        a[0][qcut_row_reg<NB>(4 * k)] = v.x; a[0][qcut_row_reg<NB>(4 * k + 1)] = v.y;
        a[0][qcut_row_reg<NB>(4 * k + 2)] = v.z; a[0][qcut_row_reg<NB>(4 * k + 3)] = v.w;
      }
    }
  }
}

// (2) raw qpd bytes (one byte per shot) of this lane's shots, QRAW words in row order:
//     NB = 8: word m = qpd bytes of rows 4m..4m+3;  NB = 4: word k = rows 4k..4k+3;
//     NB = 2: word 2k+h = bytes (row 4k+2h, batch 0/1), (row 4k+2h+1, batch 0/1);
//     NB = 1: the 32 words of the qpd block, same geometry as the observable block.
template <int NB, bool FULL, class Src>
QCUT_DEV void qcut_qpd_loads(const Src src, int n, int lane, uint32_t* qraw) {
  typedef QcutGeo<NB> G;
  if constexpr (G::GROUPED) {
    // group g: the 16 qpd bytes of shots 16*(32g+lane) .. +15 -> words 4g .. 4g+3 (rows 16g+4w .. +3)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      qcut_u4 v = {0u, 0u, 0u, 0u};
      if (FULL || (g * 32 + lane) * 16 < n) v = src.qpd16((g * 32 + lane) * 16);
      qraw[4 * g] = v.x; qraw[4 * g + 1] = v.y; qraw[4 * g + 2] = v.z; qraw[4 * g + 3] = v.w;
    }
  } else if constexpr (NB == 1) {
#pragma unroll
    for (int k = 0; k < G::CHUNKS; ++k) {
      qcut_u4 v = {0u, 0u, 0u, 0u};
      if (FULL || (k * 32 + lane) * G::SPC < n) v = src.qpd16(lane * 16 + k * 512);
      qraw[4 * k] = v.x; qraw[4 * k + 1] = v.y; qraw[4 * k + 2] = v.z; qraw[4 * k + 3] = v.w;
    }
  } else {
    const int ql = lane * G::SPC;  // SPC qpd bytes per chunk per lane
#pragma unroll
    for (int k = 0; k < G::CHUNKS; ++k) {
      const int idx = k * 32 + lane;
      if constexpr (NB == 8) {
        if (k & 1) continue;  // chunks k, k+1 -> rows 2k .. 2k+3
        uint32_t w = 0;
        if (FULL || idx * 2 < n) w = src.qpd2(ql + k * 64);
        if (FULL || (idx + 32) * 2 < n) w |= src.qpd2(ql + (k + 1) * 64) << 16;
        qraw[k / 2] = w;
      } else if constexpr (NB == 4) {
        uint32_t w = 0;
        if (FULL || idx * 4 < n) w = src.qpd4(ql + k * 128);
        qraw[k] = w;
      } else {  // NB == 2
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w = 0;
          if (FULL || idx * 8 + 4 * h < n) w = src.qpd4(ql + k * 256 + 4 * h);
          qraw[2 * k + h] = w;
        }
      }
    }
  }
}

// (3) qpd parity plane(s): bit j of qp[u] = parity of ALL bits of the qpd row of the shot in row j,
//     batch u (reference cutting_reconstruction.py:213: pad bits above num_bits count too).
template <int NB>
QCUT_DEV void qcut_qpd_planes(uint32_t* qraw, uint32_t* qp) {
  typedef QcutGeo<NB> G;
  if constexpr (NB == 1 || NB == 2) {
    // Fold-and-pack: after `v ^= v >> 4` the low nibble of every byte carries the byte's parity information, so two words
    // share one (low nibbles of the first, low nibbles of the second moved up); the same with 2-bit and 1-bit fields.
    // QRAW words -> QRAW / 8 words whose byte b holds, at bit 4 s1 + 2 s2 + s3, the parity of byte b of raw word
    // 8 p + 4 s3 + 2 s2 + s1.  86 instructions for 16 words instead of ~250 byte by byte.
    constexpr int N = G::QRAW;
#pragma unroll
    for (int i = 0; i < N; ++i) qraw[i] ^= qcut_shr<4>(qraw[i]);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) qraw[i] = qcut_sel<0x0F0F0F0Fu>(qraw[2 * i], qraw[2 * i + 1] << 4);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) qraw[i] ^= qcut_shr<2>(qraw[i]);
#pragma unroll
    for (int i = 0; i < N / 4; ++i) qraw[i] = qcut_sel<0x33333333u>(qraw[2 * i], qraw[2 * i + 1] << 2);
#pragma unroll
    for (int i = 0; i < N / 4; ++i) qraw[i] ^= qcut_shr<1>(qraw[i]);
#pragma unroll
    for (int i = 0; i < N / 8; ++i) qraw[i] = qcut_sel<0x55555555u>(qraw[2 * i], qraw[2 * i + 1] << 1);
    if constexpr (NB == 2) {  // bytes (0, 2) of a raw word belong to batch 0, (1, 3) to batch 1
      qp[0] = __byte_perm(qraw[0], qraw[1], 0x6420);
      qp[1] = __byte_perm(qraw[0], qraw[1], 0x7531);
    } else {  // byte u of a raw word belongs to batch u
      const uint32_t lo02 = __byte_perm(qraw[0], qraw[1], 0x6240), lo13 = __byte_perm(qraw[0], qraw[1], 0x7351);
      const uint32_t hi02 = __byte_perm(qraw[2], qraw[3], 0x6240), hi13 = __byte_perm(qraw[2], qraw[3], 0x7351);
      qp[0] = __byte_perm(lo02, hi02, 0x5410);
      qp[2] = __byte_perm(lo02, hi02, 0x7632);
      qp[1] = __byte_perm(lo13, hi13, 0x5410);
      qp[3] = __byte_perm(lo13, hi13, 0x7632);
    }
  } else if constexpr (QCUT_FOLD_NB(NB)) {
    // eight raw words (word i = bytes of rows 4 i .. 4 i + 3) -> one word, same network: 42 instructions instead of 88
#pragma unroll
    for (int i = 0; i < 8; ++i) qraw[i] ^= qcut_shr<4>(qraw[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) qraw[i] = qcut_sel<0x0F0F0F0Fu>(qraw[2 * i], qraw[2 * i + 1] << 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) qraw[i] ^= qcut_shr<2>(qraw[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) qraw[i] = qcut_sel<0x33333333u>(qraw[2 * i], qraw[2 * i + 1] << 2);
#pragma unroll
    for (int i = 0; i < 2; ++i) qraw[i] ^= qcut_shr<1>(qraw[i]);
    qp[0] = qcut_sel<0x55555555u>(qraw[0], qraw[1] << 1);
  } else {
    // Fold each byte to its parity, then gather the parity bits of four bytes with one multiply
    // (no carries: all partial products hit distinct bits).
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u) qp[u] = 0;
#pragma unroll
    for (int i = 0; i < G::QRAW; ++i) {
      if constexpr (NB == 2) {  // bytes (b0, b2) of each word -> batch 0, (b1, b3) -> batch 1
        const uint32_t r = (qcut_byte_parity(qraw[i]) * 0x01100220u) >> 24;
        qp[0] |= (r & 3u) << (2 * i);
        qp[G::BATCHES - 1] |= ((r >> 4) & 3u) << (2 * i);
      } else {
        const uint32_t nib = (qcut_byte_parity(qraw[i]) * 0x01020408u) >> 24;
        qp[0] |= nib << (4 * i);
      }
    }
  }
}

// (4) tail step: rows past the end of the PUB may hold bytes of the padding / next matrix.
template <int NB>
QCUT_DEV void qcut_mask_tail(int n, int lane, uint32_t (*a)[32], uint32_t* qp) {
  typedef QcutGeo<NB> G;
#pragma unroll
  for (int u = 0; u < G::BATCHES; ++u) {
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) valid |= (uint32_t)(qcut_shot_of<NB>(j, u, lane) < n) << j;
    qp[u] &= valid;
#pragma unroll
    for (int blk = 0; blk < G::BLOCKS; ++blk)
#pragma unroll
      for (int b = 0; b < G::PLANES; ++b) a[blk][u * G::PLANES + b] &= valid;
  }
}

// The whole step: a[blk][.] <- bit-planes of this lane's shots, qp[u] <- qpd parity plane(s).
// qpd_g always points at the step's qpd rows in global memory (used for qpd rows wider than one byte).
template <int NB, bool FULL, class Src>
QCUT_DEV void qcut_step(const Src src, const uint8_t* __restrict__ qpd_g, int nbq, int n, int lane,
                        uint32_t (*a)[32], uint32_t* qp) {
  typedef QcutGeo<NB> G;
  qcut_step_loads<NB, FULL, Src>(src, n, lane, a);
  if (nbq == 1) {
    uint32_t qraw[G::QRAW];
    qcut_qpd_loads<NB, FULL, Src>(src, n, lane, qraw);
    qcut_qpd_planes<NB>(qraw, qp);
  } else {
    // Any other qpd row width: per-shot byte loop (rare: more than 8 cut measurements).
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u) qp[u] = 0;
    qcut_qp_slow<NB>(qpd_g, nbq, n, lane, qp);
  }
#pragma unroll
  for (int blk = 0; blk < G::BLOCKS; ++blk) qcut_transpose32(a[blk]);
  if constexpr (!FULL) qcut_mask_tail<NB>(n, lane, a, qp);
}

template <int NB, class Src>
QCUT_DEV void qcut_step_src(const Src src, const uint8_t* __restrict__ qpd_g, int nbq, int n, int lane,
                            uint32_t (*a)[32], uint32_t* qp) {
  if (n == QcutGeo<NB>::STEP_SHOTS) qcut_step<NB, true, Src>(src, qpd_g, nbq, n, lane, a, qp);
  else qcut_step<NB, false, Src>(src, qpd_g, nbq, n, lane, a, qp);
}

template <int NB>
QCUT_DEV void qcut_step_any(const uint8_t* __restrict__ obs_s, const uint8_t* __restrict__ qpd_s, int nbq, int n,
                            int lane, uint32_t (*a)[32], uint32_t* qp) {
  const QcutSrcGlobal src = {obs_s, qpd_s};
  qcut_step_src<NB, QcutSrcGlobal>(src, qpd_s, nbq, n, lane, a, qp);
}

// ---- term engine A: masks baked into generated code (struct Terms from jit.cu) -----------------
// acc[] receives, packed four 8-bit fields per word, the number of odd-parity shots this lane
// saw for every term (<= 128 per tile, so the fields cannot overflow).
template <int NB, class Terms>
QCUT_DEV void qcut_lane_tile(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq,
                             int ns, int lane, uint32_t* acc) {
  typedef QcutGeo<NB> G;
#pragma unroll
  for (int r = 0; r < Terms::NACC; ++r) acc[r] = 0;
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    uint32_t a[G::BLOCKS][32];
    uint32_t qp[G::BATCHES];
    qcut_step_any<NB>(obs + (uint64_t)s0 * NB, qpd + (uint64_t)s0 * nbq, nbq, n, lane, a, qp);
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u)
      Terms::run(&a[0][u * G::PLANES], &a[G::BLOCKS - 1][0], qp[u], acc);
  }
}

// Same loop for a kernel that serves many small groups: Net::run switches on the variant of the tile's group.
template <int NB, class Net>
QCUT_DEV void qcut_lane_tile_net(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns, int lane,
                                 uint32_t* acc, int variant) {
  typedef QcutGeo<NB> G;
#pragma unroll
  for (int r = 0; r < Net::NACC; ++r) acc[r] = 0;
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    uint32_t a[G::BLOCKS][32];
    uint32_t qp[G::BATCHES];
    qcut_step_any<NB>(obs + (uint64_t)s0 * NB, qpd + (uint64_t)s0 * nbq, nbq, n, lane, a, qp);
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u) Net::run(&a[0][u * G::PLANES], &a[G::BLOCKS - 1][0], qp[u], acc, variant);
  }
}

// ---- term engine A', 8-byte rows, one 32x32 block at a time ("split") ----------------------------------
// The all-at-once form above keeps both blocks of a step -- 64 bit-plane registers -- live through the whole XOR
// network: with the packed counters that is 128 registers, 16 warps per SM, and the profile shows the warps
// waiting on each other's POPC / load latency (issue slots 57 % busy).  Most terms of a real observable touch one
// half of the row only (a term acts on a few neighbouring qubits), so the generator (jit.cu) sorts the terms into
// block-0 terms, block-1 terms and cross terms, and this engine
//   loads the 16 chunks once (x, z -> block 0 registers; y, w -> parked in the lane's private shared-memory column),
//   transposes block 0, runs the block-0 terms and the block-0 half of every cross term (Terms::run0; the partial
//   parity words of the cross terms stay in KEEP registers),
//   fetches the parked words, transposes block 1 and finishes (Terms::run1).
// Peak live registers: 32 planes + KEEP + counters, which lets 20-24 warps share an SM.
// Parking area accesses: volatile PTX so that the compiler cannot forward the stored words to the later loads in
// registers (which would bring the 64 live registers back).
#ifndef QCUT_HOST_EMU
QCUT_DEV void qcut_park_store(uint32_t addr, uint32_t lo, uint32_t hi) {
  asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+128], %2;" :: "r"(addr), "r"(lo), "r"(hi) : "memory");
}
QCUT_DEV uint32_t qcut_park_load(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
QCUT_DEV uint32_t qcut_park_addr(const uint32_t* park) { return (uint32_t)__cvta_generic_to_shared(park); }
#else
static uint32_t* qcut_park_base;
QCUT_DEV uint32_t qcut_park_addr(uint32_t* park) { qcut_park_base = park; return 0; }
QCUT_DEV void qcut_park_store(uint32_t addr, uint32_t lo, uint32_t hi) {
  qcut_park_base[addr / 4] = lo;
  qcut_park_base[addr / 4 + 32] = hi;
}
QCUT_DEV uint32_t qcut_park_load(uint32_t addr) { return qcut_park_base[addr / 4]; }
#endif

// Loads of one step: x, z -> rows of block 0 (registers); y, w -> rows of block 1, parked at byte offset 128 * row of
// the lane's column; raw qpd bytes.  The 16 loads are issued in QCUT_SPLIT_PHASES groups, each group back to back
// (no consumer between two loads of a group) and parked before the next group is issued: one group keeps the most
// bytes in flight per warp but needs all 64 row registers for a moment, two groups never hold more than 48.
#ifndef QCUT_SPLIT_PHASES
#define QCUT_SPLIT_PHASES 1
#endif
template <bool FULL>
QCUT_DEV void qcut_split_loads(const QcutSrcGlobal src, int n, int lane, uint32_t* a, uint32_t park, uint32_t* qraw,
                               bool qpd_one_byte) {
  typedef QcutGeo<8> G;
  constexpr int PER = G::CHUNKS / QCUT_SPLIT_PHASES;
#pragma unroll
  for (int ph = 0; ph < QCUT_SPLIT_PHASES; ++ph) {
    uint32_t hi[2 * PER];
#pragma unroll
    for (int c = 0; c < PER; ++c) {
      const int k = ph * PER + c;
      qcut_u4 v = {0u, 0u, 0u, 0u};
      if (FULL || (k * 32 + lane) * G::SPC < n) v = src.obs16(lane * 16 + k * 512);
      a[2 * k] = v.x;
      a[2 * k + 1] = v.z;
      hi[2 * c] = v.y;
      hi[2 * c + 1] = v.w;
    }
    if (ph == 0 && qpd_one_byte) qcut_qpd_loads<8, FULL, QcutSrcGlobal>(src, n, lane, qraw);
#pragma unroll
    for (int c = 0; c < PER; ++c) qcut_park_store(park + 256 * (ph * PER + c), hi[2 * c], hi[2 * c + 1]);
  }
}

// park: this lane's column of the warp's 32 x 32-word parking area (row j at park[j * 32], conflict free).
template <class Terms>
QCUT_DEV void qcut_lane_tile_split(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns,
                                   int lane, uint32_t* acc, uint32_t* __restrict__ park_ptr) {
  typedef QcutGeo<8> G;
  const uint32_t park = qcut_park_addr(park_ptr);
#pragma unroll
  for (int r = 0; r < Terms::NACC; ++r) acc[r] = 0;
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    const bool full = n == G::STEP_SHOTS;
    const uint8_t* __restrict__ qpd_s = qpd + (uint64_t)s0 * nbq;
    const QcutSrcGlobal src = {obs + (uint64_t)s0 * 8, qpd_s};
    uint32_t a[32];
    uint32_t qp[1];
    uint32_t qraw[G::QRAW];
    if (full) qcut_split_loads<true>(src, n, lane, a, park, qraw, nbq == 1);
    else qcut_split_loads<false>(src, n, lane, a, park, qraw, nbq == 1);
    if (nbq == 1) {
      qcut_qpd_planes<8>(qraw, qp);
    } else {
      qp[0] = 0;
      qcut_qp_slow<8>(qpd_s, nbq, n, lane, qp);
    }
    uint32_t valid = 0xFFFFFFFFu;
    if (!full) {  // rows past the end of the PUB hold bytes of the padding / next matrix: drop their shots
      valid = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) valid |= (uint32_t)(qcut_shot_of<8>(j, 0, lane) < n) << j;
      qp[0] &= valid;
    }
    qcut_transpose32(a);
    if (!full) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] &= valid;
    }
    uint32_t keep[Terms::NKEEP > 0 ? Terms::NKEEP : 1];
    Terms::run0(a, qp[0], acc, keep);
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = qcut_park_load(park + 128 * j);
    qcut_transpose32(a);
    if (!full) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] &= valid;
    }
    Terms::run1(a, qp[0], acc, keep);
  }
}

// ---- term engine A'', 8-byte rows: the loads of the NEXT step ride inside the term network ("pipe") ----------
// The all-at-once engine issues the 16 loads of a step, waits for HBM (ncu: long_scoreboard is one of the two top
// stalls, every warp idles a full memory round trip per step) and only then computes; there is no register left to
// load a step ahead (64 planes + packed counters = 128).  But the planes DIE one after the other while the terms
// are evaluated: the generator (jit.cu) orders the terms so that bit-planes are finished as early as possible
// (min-degree elimination over the plane/term incidence) and writes one 16-byte load of the next step into the
// generated code as soon as four more registers are free (QCUT_PF_LOAD).  The rows of step s+1 -- the first step of
// the warp's next tile when the tile ends -- are therefore in flight while the terms of step s are evaluated, at no
// cost in registers, shared memory or instructions.  One flat loop over the warp's steps.
// The load is tied to the popcount `c` of the term evaluated just before it (predicate c < pfv with pfv = 64 when the
// next step is to be loaded and 0 otherwise; a popcount never exceeds 32): without a real dependence ptxas gathers all
// sixteen loads in one place and spills the planes that are still live there.
#ifndef QCUT_HOST_EMU
QCUT_DEV void qcut_pf_load(const uint8_t* p, uint32_t pfv, uint32_t c, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& w) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %5, %6;\n\t"
      "@q ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "+r"(x), "+r"(y), "+r"(z), "+r"(w) : "l"(p), "r"(c), "r"(pfv));
}
#else
QCUT_DEV void qcut_pf_load(const uint8_t* p, uint32_t pfv, uint32_t c, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& w) {
  if (c < pfv) { const qcut_u4 v = qcut_ld16(p); x = v.x; y = v.y; z = v.z; w = v.w; }
}
#endif
#define QCUT_PF_LOAD(k, c) qcut_pf_load(nsrc + (k) * 512, pfv, (c), NA[2 * (k)], NA[32 + 2 * (k)], NA[2 * (k) + 1], NA[33 + 2 * (k)])

#ifdef QCUT_HOST_EMU
// emulation of the warp-level flush for one lane: tally[t] += ns (once per warp) - 2 * (#odd shots of this lane)
template <int T>
QCUT_DEV void qcut_flush(const uint32_t* acc, int ns, int lane, unsigned long long* tally) {
  constexpr int NACC = (T + 3) / 4;
  for (int t = 0; t < T; ++t) {
    const long long odd = (acc[t % NACC] >> (8 * (t / NACC))) & 0xFFu;
    tally[t] += (unsigned long long)((lane == 0 ? (long long)ns : 0ll) - 2ll * odd);
  }
}
#else
template <int T>
QCUT_DEV void qcut_flush(const uint32_t* acc, int ns, int lane, unsigned long long* __restrict__ tally);
#endif

template <class Terms>
QCUT_DEV void qcut_warp_pipe(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,
                             const QcutTile* __restrict__ tiles, const int64_t n_tiles, int64_t tile,
                             const int64_t tile_stride, unsigned long long* __restrict__ tallies, int lane) {
  typedef QcutGeo<8> G;
  if (tile >= n_tiles) return;
  uint32_t acc[Terms::NACC];
#pragma unroll
  for (int r = 0; r < Terms::NACC; ++r) acc[r] = 0;
  uint32_t a[2][32];
  bool have = false;  // a[][] already holds the rows of the current step (loaded during the previous one)
  QcutTile te = tiles[tile];
  QcutPub pub = pubs[te.pub];
  int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;
  int ns = (int)(((int64_t)pub.shots - shot0 < QCUT_TILE_SHOTS) ? (int64_t)pub.shots - shot0 : QCUT_TILE_SHOTS);
  int nbq = (int)pub.nb_qpd;
  const uint8_t* obs = data + pub.obs_offset + (uint64_t)shot0 * 8;
  const uint8_t* qpd = data + pub.qpd_offset + (uint64_t)shot0 * nbq;
  int s0 = 0;
  for (;;) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    const bool full = n == G::STEP_SHOTS;
    const uint8_t* obs_s = obs + (uint64_t)s0 * 8;
    const uint8_t* qpd_s = qpd + (uint64_t)s0 * nbq;
    const int rest = ns - s0 - G::STEP_SHOTS;
    const bool last = rest <= 0;
    // the step after this one: the next step of the tile, or the first step of the warp's next tile
    const uint8_t* n_obs = obs_s + G::STEP_SHOTS * 8;
    bool pf = rest >= G::STEP_SHOTS;  // only full steps are loaded ahead (tails take the predicated path)
    if (last && tile + tile_stride < n_tiles) {
      const QcutTile te_n = tiles[tile + tile_stride];
      const QcutPub pub_n = pubs[te_n.pub];
      const int64_t shot0_n = (int64_t)te_n.tile_in_pub * QCUT_TILE_SHOTS;
      n_obs = data + pub_n.obs_offset + (uint64_t)shot0_n * 8;
      pf = (int64_t)pub_n.shots - shot0_n >= G::STEP_SHOTS;
    }
    const QcutSrcGlobal src = {obs_s, qpd_s};
    if (!have) {
      if (full) qcut_step_loads<8, true, QcutSrcGlobal>(src, n, lane, a);
      else qcut_step_loads<8, false, QcutSrcGlobal>(src, n, lane, a);
    }
    uint32_t qp[1];
    if (nbq == 1) {
      uint32_t qraw[G::QRAW];
      if (full) qcut_qpd_loads<8, true, QcutSrcGlobal>(src, n, lane, qraw);
      else qcut_qpd_loads<8, false, QcutSrcGlobal>(src, n, lane, qraw);
      qcut_qpd_planes<8>(qraw, qp);
    } else {
      qp[0] = 0;
      qcut_qp_slow<8>(qpd_s, nbq, n, lane, qp);
    }
    qcut_transpose32(a[0]);
    qcut_transpose32(a[1]);
    if (!full) qcut_mask_tail<8>(n, lane, a, qp);
    uint32_t na[2][32];
#ifdef QCUT_HOST_EMU
    memset(na, 0, sizeof na);
#else
#pragma unroll
    for (int j = 0; j < 32; ++j) { na[0][j] = a[0][j]; na[1][j] = a[1][j]; }  // "undefined": any register will do
#endif
    Terms::run_pf(a[0], a[1], qp[0], acc, n_obs + lane * 16, pf ? 64u : 0u, &na[0][0]);
#pragma unroll
    for (int j = 0; j < 32; ++j) { a[0][j] = na[0][j]; a[1][j] = na[1][j]; }
    have = pf;
    if (!last) {
      s0 += G::STEP_SHOTS;
      continue;
    }
    qcut_flush<Terms::T>(acc, ns, lane, tallies + pub.tally_offset);
#pragma unroll
    for (int r = 0; r < Terms::NACC; ++r) acc[r] = 0;
    tile += tile_stride;
    if (tile >= n_tiles) break;
    te = tiles[tile];
    pub = pubs[te.pub];
    shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;
    ns = (int)(((int64_t)pub.shots - shot0 < QCUT_TILE_SHOTS) ? (int64_t)pub.shots - shot0 : QCUT_TILE_SHOTS);
    nbq = (int)pub.nb_qpd;
    obs = data + pub.obs_offset + (uint64_t)shot0 * 8;
    qpd = data + pub.qpd_offset + (uint64_t)shot0 * nbq;
    s0 = 0;
  }
}

// ---- term engine A3, rows of up to 8 bytes: packed counters in shared memory ("sacc") ---------------------------
// With 64 bit-planes and ceil(T / 4) packed counters in registers the all-at-once engine sits at the 128-register
// cap: ptxas keeps a dozen counters on the stack and has two or three temporaries left, so that in the transposition
// every LOP3 follows the IMAD shift it depends on at a distance of two issue slots (ncu: `wait` is the top stall).
// Here the counters of a tile live in the thread's private column of a shared-memory array ([group of 4 counter
// words][thread], one 128-bit access per 16 terms, conflict free); the generated network (Terms::run_sacc) loads a
// group, adds the sixteen popcounts of its terms, stores it back.  18 extra LSU instructions per step for cfg4
// (1.5 %), ~30 registers handed back to the scheduler.
#ifndef QCUT_SACC_STRIDE
#define QCUT_SACC_STRIDE 4096  // bytes between two groups of one thread = 16 * threads per CTA (set by the generator)
#endif
struct QcutSacc {
#ifndef QCUT_HOST_EMU
  uint32_t base;  // shared-window address of this thread's 16 bytes of group 0
  QCUT_MDEV qcut_u4 ld4(int g) const {
    qcut_u4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + g * QCUT_SACC_STRIDE));
    return v;
  }
  QCUT_MDEV void st4(int g, qcut_u4 v) const {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(base + g * QCUT_SACC_STRIDE), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
#else
  uint32_t* base;
  QCUT_MDEV qcut_u4 ld4(int g) const { qcut_u4 v; memcpy(&v, base + 4 * g, 16); return v; }
  QCUT_MDEV void st4(int g, qcut_u4 v) const { memcpy(base + 4 * g, &v, 16); }
#endif
};

template <int NACC>
QCUT_DEV void qcut_sacc_clear(const QcutSacc S) {
  const qcut_u4 z = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int g = 0; g < (NACC + 3) / 4; ++g) S.st4(g, z);
}

// counters of the finished tile -> registers (for qcut_flush), shared-memory copy cleared for the next tile
template <int NACC>
QCUT_DEV void qcut_sacc_take(const QcutSacc S, uint32_t* acc) {
  const qcut_u4 z = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int g = 0; g < (NACC + 3) / 4; ++g) {
    const qcut_u4 v = S.ld4(g);
    if (4 * g < NACC) acc[4 * g] = v.x;
    if (4 * g + 1 < NACC) acc[4 * g + 1] = v.y;
    if (4 * g + 2 < NACC) acc[4 * g + 2] = v.z;
    if (4 * g + 3 < NACC) acc[4 * g + 3] = v.w;
    S.st4(g, z);
  }
}

// same contract as qcut_lane_tile, counters in S (cleared on entry by the previous qcut_sacc_take / qcut_sacc_clear)
template <int NB, class Terms>
QCUT_DEV void qcut_lane_tile_sacc(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns, int lane,
                                  const QcutSacc S) {
  typedef QcutGeo<NB> G;
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    uint32_t a[G::BLOCKS][32];
    uint32_t qp[G::BATCHES];
    qcut_step_any<NB>(obs + (uint64_t)s0 * NB, qpd + (uint64_t)s0 * nbq, nbq, n, lane, a, qp);
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u) Terms::run_sacc(&a[0][u * G::PLANES], &a[G::BLOCKS - 1][0], qp[u], S);
  }
}

// ---- pieces of the 32x32 transposition, one pair of rows at a time ---------------------------------------------------
// qcut_tp<D>(a, j) exchanges the off-diagonal DxD bit blocks of rows j and j + D.  The five stages D = 16, 8, 4, 2, 1
// swap five different bits of the row index with five different bits of the bit index, so they commute: any order
// gives qcut_transpose32.  The generator uses them to interleave the transposition of one block with the term network
// of another (Terms::run_il): the transposition keeps the ALU and FMA pipes busy, the terms the XU pipe (POPC issues at a
// quarter rate), and a warp that does one after the other leaves one of them idle.
template <int D>
QCUT_DEV void qcut_tp(uint32_t* a, int j) {
  const uint32_t x = a[j], y = a[j + D];
  if constexpr (D == 16) {
    a[j] = __byte_perm(x, y, 0x5410);
    a[j + D] = __byte_perm(x, y, 0x7632);
  } else if constexpr (D == 8) {
    a[j] = __byte_perm(x, y, 0x6240);
    a[j + D] = __byte_perm(x, y, 0x7351);
  } else if constexpr (D == 4) {
    a[j] = qcut_sel<0x0F0F0F0Fu>(x, y << 4);
    a[j + D] = qcut_sel<0x0F0F0F0Fu>(qcut_shr<4>(x), y);
  } else if constexpr (D == 2) {
    a[j] = qcut_sel<0x33333333u>(x, y << 2);
    a[j + D] = qcut_sel<0x33333333u>(qcut_shr<2>(x), y);
  } else {
    a[j] = qcut_sel<0x55555555u>(x, y << 1);
    a[j + D] = qcut_sel<0x55555555u>(qcut_shr<1>(x), y);
  }
}

// ---- term engine A4, 8-byte rows: transposition of block 1 interleaved with the terms of block 0 ("il") ---------------
template <class Terms>
QCUT_DEV void qcut_lane_tile_il(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns, int lane,
                                uint32_t* acc) {
  typedef QcutGeo<8> G;
#pragma unroll
  for (int r = 0; r < Terms::NACC; ++r) acc[r] = 0;
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    const uint8_t* __restrict__ qpd_s = qpd + (uint64_t)s0 * nbq;
    const QcutSrcGlobal src = {obs + (uint64_t)s0 * 8, qpd_s};
    uint32_t a[2][32];
    uint32_t qp[1];
    if (n == G::STEP_SHOTS && nbq == 1) {
      qcut_step_loads<8, true, QcutSrcGlobal>(src, n, lane, a);
      uint32_t qraw[G::QRAW];
      qcut_qpd_loads<8, true, QcutSrcGlobal>(src, n, lane, qraw);
      qcut_qpd_planes<8>(qraw, qp);
      qcut_transpose32(a[0]);
      Terms::run_il(a[0], a[1], qp[0], acc);
    } else {
      qcut_step_src<8, QcutSrcGlobal>(src, qpd_s, nbq, n, lane, a, qp);
      Terms::run(a[0], a[1], qp[0], acc);
    }
  }
}

// ---- wide rows (9..32 bytes): rows staged through shared memory, bit-planes parked in shared memory ----
// Up to 128 bit-planes do not fit the register file, so the blocks are transposed one at a time and the
// planes are parked in shared memory ([plane][lane], conflict free); the generated term code then reads
// planes at *static* shared-memory offsets (`P[p * 32]`).  Rows are cut out of a per-warp raw buffer that
// the warp fills with coalesced 128-bit loads: lane l owns 16 consecutive shots per group (16 * NB
// contiguous bytes), stored at a stride of 4 * NB + 1 words so that the lanes' row reads hit 32 banks.
template <int NB>
struct QcutWide {
  static const int WARPS = NB > 16 ? 2 : 4;                  // warps per CTA (shared memory per warp grows with NB)
  static const int BLOCKS = (NB + 3) / 4;
  static const int RAW_STRIDE = 4 * NB + 1;                 // words per lane
  static const int RAW_WORDS = 32 * RAW_STRIDE;             // one group of a step (16 shots per lane)
  static const int PLANE_WORDS = BLOCKS * 32 * 32;
  static const int WARP_WORDS = RAW_WORDS + PLANE_WORDS;    // shared memory per warp (words)
};

#ifndef QCUT_HOST_EMU
QCUT_DEV void qcut_syncwarp() { __syncwarp(); }
#else
QCUT_DEV void qcut_syncwarp() {}
#endif

template <int NB, class Terms>
QCUT_DEV void qcut_lane_tile_wide(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns,
                                  int lane, uint32_t* acc, uint32_t* raw, uint32_t* planes) {
  typedef QcutGeo<NB> G;
  typedef QcutWide<NB> W;
#pragma unroll
  for (int r = 0; r < Terms::NACC; ++r) acc[r] = 0;
  uint32_t* __restrict__ col = planes + lane;  // this lane's column of the row / plane store: col[slot * 32]
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    const uint8_t* __restrict__ obs_s = obs + (uint64_t)s0 * NB;
    const uint8_t* __restrict__ qpd_s = qpd + (uint64_t)s0 * nbq;
    const bool full = n == G::STEP_SHOTS;

    // ---- (1) qpd parity plane (same grouped geometry as the 3/5/6/7-byte rows) ----
    uint32_t qp[1];
    {
      const QcutSrcGlobal src = {obs_s, qpd_s};
      if (nbq == 1) {
        uint32_t qraw[G::QRAW];
        if (full) qcut_qpd_loads<NB, true, QcutSrcGlobal>(src, n, lane, qraw);
        else qcut_qpd_loads<NB, false, QcutSrcGlobal>(src, n, lane, qraw);
        qcut_qpd_planes<NB>(qraw, qp);
      } else {
        qp[0] = 0;
        qcut_qp_slow<NB>(qpd_s, nbq, n, lane, qp);
      }
    }
    uint32_t valid_rows = 0xFFFFFFFFu;
    if (!full) {
      valid_rows = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) valid_rows |= (uint32_t)(qcut_shot_of<NB>(j, 0, lane) < n) << j;
      qp[0] &= valid_rows;
    }

    // ---- (2) per group: raw bytes -> shared memory (coalesced), then cut this lane's 16 rows of every
    //      block out of its slot and store them in its column of the row store ----
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      qcut_syncwarp();  // readers of raw[] (previous group / step) are done
#ifndef QCUT_HOST_EMU
      {
        const int region_bytes = 512 * NB;  // 32 lanes x 16 shots x NB bytes
        const uint8_t* __restrict__ src = obs_s + (uint64_t)g * region_bytes + (uint64_t)lane * 16;
        int valid = n * NB - g * region_bytes;
        valid = valid < 0 ? 0 : (valid > region_bytes ? region_bytes : valid);
        constexpr int CB = NB > 16 ? 16 : NB;  // loads in flight per lane (16 x 16 bytes at most: registers)
#pragma unroll
        for (int c0 = 0; c0 < NB; c0 += CB) {
          qcut_u4 v[CB];
          if (full) {
            // the loads of a batch are issued back to back (no predicate between a load and its use)
#pragma unroll
            for (int c = 0; c < CB; ++c)
              if (c0 + c < NB) v[c] = qcut_ld16(src + (c0 + c) * 512);
          } else {
#pragma unroll
            for (int c = 0; c < CB; ++c)  // clamped address, not a predicate: the loads stay back to back;
                                          // chunks past the end only feed rows that valid_rows zeroes
              if (c0 + c < NB) v[c] = qcut_ld16(((c0 + c) * 32 + lane) * 16 < valid ? src + (c0 + c) * 512 : obs_s);
          }
#pragma unroll
          for (int c = 0; c < CB; ++c)
            if (c0 + c < NB) {
              const int q = (c0 + c) * 32 + lane;  // 16-byte chunk of the region
              const int owner = q / NB, pos = q - owner * NB;
              uint32_t* dst = raw + owner * W::RAW_STRIDE + pos * 4;
              dst[0] = v[c].x; dst[1] = v[c].y; dst[2] = v[c].z; dst[3] = v[c].w;
            }
        }
      }
#else
      {  // emulation runs one lane at a time: the lane copies its own group
        uint32_t* dst = raw + lane * W::RAW_STRIDE;
        for (int w = 0; w < 4 * NB; ++w) dst[w] = 0;
        const int first = (g * 32 + lane) * 16;
        if (first < n) {
          int bytes = ((n - first < 16 ? n - first : 16) * NB + 15) & ~15;  // whole 16-byte chunks, as on the device
          if (bytes > 16 * NB) bytes = 16 * NB;
          memcpy(dst, obs_s + (uint64_t)first * NB, bytes);
        }
      }
#endif
      qcut_syncwarp();
      const uint32_t* __restrict__ mine = raw + lane * W::RAW_STRIDE;
#pragma unroll
      for (int blk = 0; blk < W::BLOCKS; ++blk)
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int o = r * NB + 4 * blk;  // byte offset of this row word inside the lane's group
          const int wi = o >> 2, sh = o & 3;
          const uint32_t lo = mine[wi];
          uint32_t row = sh == 0 ? lo : __byte_perm(lo, mine[wi + 1], 0x3210 + 0x1111 * sh);
          col[(blk * 32 + 16 * g + r) * 32] = row;
        }
    }

    // ---- (3) block by block, in place: 32 rows -> 32 bit-planes (the column is private to the lane) ----
#pragma unroll
    for (int blk = 0; blk < W::BLOCKS; ++blk) {
      uint32_t a[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = col[(blk * 32 + j) * 32];
      qcut_transpose32(a);
      if (!full) {  // rows past the end of the PUB hold bytes of the padding / next matrix: drop their shots
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] &= valid_rows;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) col[(blk * 32 + i) * 32] = a[i];
    }

    // ---- (4) XOR networks on planes read from shared memory at static offsets ----
    Terms::run(col, col, qp[0], acc);
  }
}

// ---- wide rows in 8-byte slabs (NB = 16, 24, 32): the narrow-row engine on one slab of the row at a time ("slab") ----
// The wide-row engine above parks up to 256 bit-planes in shared memory and runs at 8 warps per SM (0.35 of the copy
// peak on 16-byte rows).  A local observable on an un-partitioned 127-qubit register touches one 64-bit slab of the row
// per term, a handful of terms two; so a step of 1024 shots is processed slab by slab: 32 eight-byte loads per lane
// (shot 32 k + lane -> row k of the slab's two 32x32 blocks; the second slab re-reads the sectors of the first from
// L2), two in-register transpositions, the slab's terms on registers -- packed counters in shared memory (QcutSacc),
// the partial parity words of the cross-slab terms in K[] registers.  The qpd parity plane is built once per step from
// eight coalesced 4-byte loads per lane and four shuffles that bring every lane the bits of ITS shots.
#ifndef QCUT_HOST_EMU
QCUT_DEV void qcut_ld8(const uint8_t* p, uint32_t& lo, uint32_t& hi) {
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(p));
}
#else
QCUT_DEV void qcut_ld8(const uint8_t* p, uint32_t& lo, uint32_t& hi) { memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4); }
#endif

// Every 32-byte sector of a step is read once per slab.  ncu on 16-byte rows: 1.48 x the algorithmic bytes come from
// DRAM, i.e. about half of the second slab's sectors have left L2 again by the time they are wanted (16 warps per SM x
// 17 KB per step, two steps with the prefetch).  So the slabs say what they know: all but the last slab load with an
// L2 evict-last policy, the last one with evict-first (QCUT_SLAB_L2_HINTS, set by the generator).
#if !defined(QCUT_HOST_EMU) && defined(QCUT_SLAB_L2_HINTS)
QCUT_DEV uint64_t qcut_l2_policy(bool last_use) {
  uint64_t pol;
  if (last_use) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
QCUT_DEV void qcut_ld8_hint(const uint8_t* p, uint64_t pol, uint32_t& lo, uint32_t& hi) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(lo), "=r"(hi) : "l"(p), "l"(pol));
}
#endif

template <int NB, bool FULL>
QCUT_DEV void qcut_slab_loads(const uint8_t* __restrict__ obs_s, int slab, int n, int lane, uint32_t (*a)[32]) {
#if !defined(QCUT_HOST_EMU) && defined(QCUT_SLAB_L2_HINTS)
  const uint64_t pol = qcut_l2_policy(slab + 1 == NB / 8);
#endif
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    a[0][k] = 0;
    a[1][k] = 0;
    if (FULL || 32 * k + lane < n) {
#if !defined(QCUT_HOST_EMU) && defined(QCUT_SLAB_L2_HINTS)
      qcut_ld8_hint(obs_s + (uint64_t)(32 * k + lane) * NB + 8 * slab, pol, a[0][k], a[1][k]);
#else
      qcut_ld8(obs_s + (uint64_t)(32 * k + lane) * NB + 8 * slab, a[0][k], a[1][k]);
#endif
    }
  }
}

// bit k of the result = parity of ALL bits of the qpd row of shot 32 k + lane (0 for shots >= n)
template <bool FULL>
QCUT_DEV uint32_t qcut_slab_qpd(const uint8_t* __restrict__ qpd_s, int nbq, int n, int lane) {
#ifndef QCUT_HOST_EMU
  if (nbq == 1) {
    uint32_t w = 0;  // bit 4 j + b = parity of the qpd byte of shot 128 j + 4 lane + b
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t v = 0;
      if (FULL || 128 * j + 4 * lane < n) v = qcut_ld4(qpd_s + 128 * j + 4 * lane);
      w |= ((qcut_byte_parity(v) * 0x01020408u) >> 24) << (4 * j);
    }
    // shot 32 k + L with k = 4 j + m sits in lane (L >> 2) + 8 m at bit 4 j + (L & 3)
    const int b = lane & 3;
    uint32_t qp = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) qp |= ((__shfl_sync(0xffffffffu, w, (lane >> 2) + 8 * m) >> b) & 0x11111111u) << m;
    return qp;  // (shots >= n: the caller masks them)
  }
#endif
  uint32_t qp = 0;
  for (int k = 0; k < 32; ++k) {
    const int shot = 32 * k + lane;
    if (shot < n) {
      uint32_t x = 0;
      for (int bb = 0; bb < nbq; ++bb) x ^= qcut_ld1(qpd_s + (uint64_t)shot * nbq + bb);
      qp |= (uint32_t)(__popc(x) & 1) << k;
    }
  }
  return qp;
}

template <int NB, class Terms, int SLAB, bool FULL>
QCUT_DEV void qcut_slab_pass(const uint8_t* __restrict__ obs_s, int n, int lane, uint32_t qp, uint32_t valid, const QcutSacc S,
                             uint32_t* K) {
  uint32_t a[2][32];
  qcut_slab_loads<NB, FULL>(obs_s, SLAB, n, lane, a);
  qcut_transpose32(a[0]);
  qcut_transpose32(a[1]);
  if (!FULL) {  // rows past the end of the PUB were loaded as zero; drop their shots from the term words through qp / valid
#pragma unroll
    for (int i = 0; i < 32; ++i) { a[0][i] &= valid; a[1][i] &= valid; }
  }
  Terms::template run_slab<SLAB>(a[0], a[1], qp, S, K);
  if constexpr (SLAB + 1 < NB / 8) qcut_slab_pass<NB, Terms, SLAB + 1, FULL>(obs_s, n, lane, qp, valid, S, K);
}

// The kernel waits on its loads (ncu: long_scoreboard on top, issue slots 37 % busy): while a step is processed the copy
// engine pulls the rows of the next step of the tile into L2 (no registers, no shared memory, two instructions).
QCUT_DEV void qcut_slab_prefetch_l2(const uint8_t* p, uint32_t bytes) {
#ifndef QCUT_HOST_EMU
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
#endif
}

// counters in S (cleared on entry by the previous qcut_sacc_take / qcut_sacc_clear)
template <int NB, class Terms>
QCUT_DEV void qcut_lane_tile_slab(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns, int lane,
                                  const QcutSacc S) {
  for (int s0 = 0; s0 < ns; s0 += 1024) {
    const int n = (ns - s0 < 1024) ? ns - s0 : 1024;
    const uint8_t* __restrict__ obs_s = obs + (uint64_t)s0 * NB;
    const uint8_t* __restrict__ qpd_s = qpd + (uint64_t)s0 * nbq;
    uint32_t K[Terms::NK > 0 ? Terms::NK : 1];
#ifdef QCUT_SLAB_PREFETCH
    if (lane == 0 && ns - s0 > 1024) {
      const int rest = ns - s0 - 1024 < 1024 ? ns - s0 - 1024 : 1024;
      qcut_slab_prefetch_l2(obs_s + 1024 * NB, ((uint32_t)rest * NB + 15u) & ~15u);
      qcut_slab_prefetch_l2(qpd_s + 1024 * nbq, ((uint32_t)rest * (uint32_t)nbq + 15u) & ~15u);
    }
#endif
    if (n == 1024) {
      const uint32_t qp = qcut_slab_qpd<true>(qpd_s, nbq, n, lane);
      qcut_slab_pass<NB, Terms, 0, true>(obs_s, n, lane, qp, 0xFFFFFFFFu, S, K);
    } else {
      uint32_t valid = 0;
#pragma unroll
      for (int k = 0; k < 32; ++k) valid |= (uint32_t)(32 * k + lane < n) << k;
      const uint32_t qp = qcut_slab_qpd<false>(qpd_s, nbq, n, lane) & valid;
      qcut_slab_pass<NB, Terms, 0, false>(obs_s, n, lane, qp, valid, S, K);
    }
  }
}

// ---- term engine B: run-time masks, bit-planes parked in shared memory -------------------------
// planes[p * 32] is this lane's word of memory-bit plane p (the caller passes the lane's column).
// mask words are the little-endian words of the mask row in memory order (plan table).
// Returns popcount of the parity word of term t for this lane's 32 shots.
QCUT_DEV uint32_t qcut_rt_term(const uint32_t* planes, uint32_t qp, uint32_t m_lo, uint32_t m_hi) {
  uint32_t x = qp;
  while (m_lo) {
    const int p = __ffs((int)m_lo) - 1;
    m_lo &= m_lo - 1;
    x ^= planes[p * 32];
  }
  while (m_hi) {
    const int p = __ffs((int)m_hi) - 1;
    m_hi &= m_hi - 1;
    x ^= planes[(32 + p) * 32];
  }
  return (uint32_t)__popc(x);
}

#ifndef QCUT_HOST_EMU
// 64-bit integer reduction to global memory.  Plain PTX: the lanes target distinct addresses, so the
// compiler's warp-aggregated atomicAdd expansion (match + REDUX + elect) would only add instructions.
QCUT_DEV void qcut_red_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Warp-level flush of the packed per-lane counters: tally[t] += ns - 2 * (#odd shots).
// Term t sits in field t / NACC of register t % NACC (NACC = ceil(T / 4)).  Branch free: every packed
// register is widened to two pairs of 16-bit fields (the warp sum is <= 4096) and summed with two
// REDUX; lane l = 4 * (r % 8) + f keeps the total of (register r, field f) for r = l/4, l/4 + 8, ...,
// and the warp finishes with ceil(NACC / 8) 64-bit reductions to global memory.
// warp sum of a 32-bit value (plain REDUX: the lanes of a warp are converged here by construction)
QCUT_DEV uint32_t qcut_redux_add(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

template <int T>
QCUT_DEV void qcut_flush(const uint32_t* acc, int ns, int lane, unsigned long long* __restrict__ tally) {
  constexpr int NACC = (T + 3) / 4;
  uint32_t mine[(NACC + 7) / 8];
#pragma unroll
  for (int c = 0; c < (NACC + 7) / 8; ++c) mine[c] = 0;
  const int field = lane & 3;
  const int owner = lane >> 2;  // this lane owns registers r with r % 8 == owner
  // this lane's field is one 16-bit half of the pair (ev, od): field 0 -> ev low, 1 -> od low, 2 -> ev high, 3 -> od high;
  // one PRMT with a per-lane selector takes it out (upper half: don't care, masked once at the end)
  const uint32_t first = ((field & 1) ? 4u : 0u) + ((field & 2) ? 2u : 0u);
  const uint32_t pick = ((first + 1u) << 4) | first;
#pragma unroll
  for (int r = 0; r < NACC; ++r) {
    const uint32_t ev = qcut_redux_add(acc[r] & 0x00FF00FFu);                  // fields 0, 2
    const uint32_t od = qcut_redux_add(__byte_perm(acc[r], 0u, 0x4341));      // fields 1, 3: bytes (1, -, 3, -)
    const uint32_t val = __byte_perm(ev, od, pick);
    if ((r & 7) == owner) mine[r >> 3] = val;
  }
#pragma unroll
  for (int c = 0; c < (NACC + 7) / 8; ++c) {
    const int r = 8 * c + owner;
    const int t = field * NACC + r;
    if (r < NACC && t < T)
      qcut_red_add(tally + t, (unsigned long long)((long long)ns - 2ll * (long long)(mine[c] & 0xFFFFu)));
  }
}
// The flush in two halves, for kernels in which a warp takes the tiles of a PUB two at a time (8192 shots): the warp
// sums of each tile are gathered into the owner lanes' registers (the 8-bit per-lane fields hold one tile), and ONE
// update per term goes to global memory for the pair -- a plain 64-bit store when the pair is the whole PUB (nobody
// else adds to those tallies; they were zeroed by the launch), a RED otherwise.  Measured on cfg4 (two tiles per PUB):
// the REDs of the one-tile flush cost 5.5 % of K1 (scripts/micro/k1_compute_only.py).
template <int T>
QCUT_DEV void qcut_flush_gather(const uint32_t* acc, int lane, uint32_t* mine) {
  constexpr int NACC = (T + 3) / 4;
  const int field = lane & 3;
  const bool odd_field = field & 1;
  const uint32_t shift = (field & 2) ? 16u : 0u;
  const int owner = lane >> 2;
#pragma unroll
  for (int r = 0; r < NACC; ++r) {
    const uint32_t ev = __reduce_add_sync(0xffffffffu, acc[r] & 0x00FF00FFu);
    const uint32_t od = __reduce_add_sync(0xffffffffu, (acc[r] >> 8) & 0x00FF00FFu);
    const uint32_t val = ((odd_field ? od : ev) >> shift) & 0xFFFFu;
    if ((r & 7) == owner) mine[r >> 3] += val;
  }
}
template <int T>
QCUT_DEV void qcut_flush_commit(const uint32_t* mine, long long ns, int lane, unsigned long long* __restrict__ tally, bool whole_pub) {
  constexpr int NACC = (T + 3) / 4;
  const int field = lane & 3;
  const int owner = lane >> 2;
#pragma unroll
  for (int c = 0; c < (NACC + 7) / 8; ++c) {
    const int r = 8 * c + owner;
    const int t = field * NACC + r;
    if (r < NACC && t < T) {
      const unsigned long long v = (unsigned long long)(ns - 2ll * (long long)mine[c]);
      if (whole_pub) tally[t] = v;
      else qcut_red_add(tally + t, v);
    }
  }
}
// The same flush for a network whose number of terms is only known at run time (kernels that serve many small
// groups, Net::run switches on the group's variant): T terms in nacc = ceil(T / 4) <= MAXNACC <= 8 packed registers,
// so every lane owns at most one (register, field).
template <int MAXNACC>
QCUT_DEV void qcut_flush_rt(const uint32_t* acc, int T, int nacc, int ns, int lane, unsigned long long* __restrict__ tally) {
  const int field = lane & 3;
  const bool odd_field = field & 1;
  const uint32_t shift = (field & 2) ? 16u : 0u;
  const int owner = lane >> 2;
  uint32_t mine = 0;
#pragma unroll
  for (int r = 0; r < MAXNACC; ++r) {
    const uint32_t ev = __reduce_add_sync(0xffffffffu, acc[r] & 0x00FF00FFu);
    const uint32_t od = __reduce_add_sync(0xffffffffu, (acc[r] >> 8) & 0x00FF00FFu);
    const uint32_t val = ((odd_field ? od : ev) >> shift) & 0xFFFFu;
    if (r == owner) mine = val;
  }
  const int t = field * nacc + owner;
  if (owner < nacc && t < T) qcut_red_add(tally + t, (unsigned long long)((long long)ns - 2ll * (long long)mine));
}

// What the tile loops need to know about the term network(s) of a kernel.  One group per kernel: the generated
// `struct Terms`; many small groups per kernel: a generated `struct Net` whose run() switches on the variant index of
// the tile's group (QcutGroup::variant).
template <class Terms>
struct QcutSingleNet {
  static const int NB = Terms::NB;
  static const int NACC = Terms::NACC;
  static const bool MULTI = false;
  static QCUT_MDEV void run(const uint32_t* P, const uint32_t* Q, uint32_t qp, uint32_t* acc, int) { Terms::run(P, Q, qp, acc); }
  static QCUT_MDEV void flush(const uint32_t* acc, int ns, int lane, unsigned long long* tally, int) {
    qcut_flush<Terms::T>(acc, ns, lane, tally);
  }
};
#endif  // !QCUT_HOST_EMU
#if !defined(QCUT_HOST_EMU) && defined(QCUT_WITH_VARIANTS)
// (experiments that measured slower than the default, DESIGN.md section 4: the generator defines QCUT_WITH_VARIANTS
//  only when a plan asks for one of them, so NVRTC does not parse them otherwise)
// ---- direct (HBM -> registers) tile loop with L2 prefetch one step ahead (device only) ---------------
// cp.async.bulk.prefetch.L2 costs no registers and no shared memory: while the warp works on step s,
// the copy engine pulls the bytes of step s+1 (or of the first step of the warp's next tile) into L2,
// so the LDG.128 of the next step are L2 hits instead of HBM round trips.
QCUT_DEV void qcut_prefetch_l2(const uint8_t* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

template <int NB>
QCUT_DEV void qcut_prefetch_step(const uint8_t* obs_s, const uint8_t* qpd_s, int nbq, int n) {
  if (n <= 0) return;
  qcut_prefetch_l2(obs_s, ((uint32_t)n * NB + 15u) & ~15u);
  qcut_prefetch_l2(qpd_s, ((uint32_t)n * (uint32_t)nbq + 15u) & ~15u);
}

// Same contract as qcut_lane_tile; (next_obs, next_qpd, next_nbq, next_n) describe the first step of
// the warp's next tile (next_n = 0: none).
template <int NB, class Terms>
QCUT_DEV void qcut_lane_tile_pf(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ qpd, int nbq, int ns,
                                int lane, uint32_t* acc, const uint8_t* next_obs, const uint8_t* next_qpd,
                                int next_nbq, int next_n) {
  typedef QcutGeo<NB> G;
#pragma unroll
  for (int r = 0; r < Terms::NACC; ++r) acc[r] = 0;
  for (int s0 = 0; s0 < ns; s0 += G::STEP_SHOTS) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    if (lane == 0) {
      const int rest = ns - s0 - G::STEP_SHOTS;
      if (rest > 0)
        qcut_prefetch_step<NB>(obs + (uint64_t)(s0 + G::STEP_SHOTS) * NB, qpd + (uint64_t)(s0 + G::STEP_SHOTS) * nbq, nbq,
                               rest < G::STEP_SHOTS ? rest : G::STEP_SHOTS);
      else
        qcut_prefetch_step<NB>(next_obs, next_qpd, next_nbq, next_n < G::STEP_SHOTS ? next_n : G::STEP_SHOTS);
    }
    uint32_t a[G::BLOCKS][32];
    uint32_t qp[G::BATCHES];
    qcut_step_any<NB>(obs + (uint64_t)s0 * NB, qpd + (uint64_t)s0 * nbq, nbq, n, lane, a, qp);
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u)
      Terms::run(&a[0][u * G::PLANES], &a[G::BLOCKS - 1][0], qp[u], acc);
  }
}

// ---- bulk-copy staged stream of steps (device only) ---------------------------------------------
// One staging buffer and one mbarrier per warp.  While the warp does the arithmetic of step s, the
// copy engine brings step s+1 (possibly the first step of the warp's next tile) into the buffer:
//   wait(full) -> LDS everything into registers -> __syncwarp -> lane 0 issues the next bulk copies
//   -> transpose + XOR networks of the current step.
QCUT_DEV uint32_t qcut_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

QCUT_DEV void qcut_mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
QCUT_DEV void qcut_mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
QCUT_DEV void qcut_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
QCUT_DEV void qcut_mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "QCUT_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra QCUT_DONE;\n"
      "bra QCUT_WAIT;\n"
      "QCUT_DONE:\n"
      "}" :: "r"(mbar), "r"(parity) : "memory");
}

template <int NB>
struct QcutStage {
  static const int OBS_BYTES = QcutGeo<NB>::STEP_SHOTS * NB;
  static const int QPD_BYTES = QcutGeo<NB>::STEP_SHOTS;      // staged only when nb_qpd == 1
  static const int BYTES = OBS_BYTES + QPD_BYTES;            // e.g. 9216 / 5120 / 6144 / 8192 for NB = 8 / 4 / 2 / 1
};

// Issue the bulk copies of one step (lane 0 only).  Sizes are rounded up to 16 bytes: every matrix is
// readable up to its 16-byte padded end (include/qcut_b200.h).
template <int NB>
QCUT_DEV void qcut_stage_issue(uint32_t buf, uint32_t mbar, const uint8_t* obs_s, const uint8_t* qpd_s, int nbq, int n) {
  const uint32_t ob = ((uint32_t)n * NB + 15u) & ~15u;
  const uint32_t qb = (nbq == 1) ? (((uint32_t)n + 15u) & ~15u) : 0u;
  qcut_mbar_expect_tx(mbar, ob + qb);
  qcut_bulk_g2s(buf, obs_s, ob, mbar);
  if (qb) qcut_bulk_g2s(buf + QcutStage<NB>::OBS_BYTES, qpd_s, qb, mbar);
}

template <int NB, class Net>
QCUT_DEV void qcut_warp_stream(const uint8_t* __restrict__ data, const QcutPub* __restrict__ pubs,
                               const QcutTile* __restrict__ tiles, const int64_t n_tiles, int64_t first_tile,
                               int64_t tile_stride, unsigned long long* __restrict__ tallies, uint8_t* stage,
                               unsigned long long* mbar_ptr, int lane, const QcutGroup* __restrict__ groups) {
  typedef QcutGeo<NB> G;
  if (first_tile >= n_tiles) return;
  const uint32_t buf = qcut_smem_addr(stage);
  const uint32_t mbar = qcut_smem_addr(mbar_ptr);
  if (lane == 0) qcut_mbar_init(mbar, 1);
  __syncwarp();

  // current tile
  int64_t tile = first_tile;
  QcutTile te = tiles[tile];
  QcutPub pub = pubs[te.pub];
  // next tile's descriptors are fetched one tile ahead so their latency never shows
  QcutTile te_next = te;
  QcutPub pub_next = pub;
  bool have_next = tile + tile_stride < n_tiles;
  if (have_next) { te_next = tiles[tile + tile_stride]; pub_next = pubs[te_next.pub]; }
  int variant = 0, variant_next = 0;  // which network of a multi-group kernel serves the tile
  if (Net::MULTI) {
    variant = (int)groups[pub.group].variant;
    if (have_next) variant_next = (int)groups[pub_next.group].variant;
  }

  int64_t shot0 = (int64_t)te.tile_in_pub * QCUT_TILE_SHOTS;
  int ns = (int)(((int64_t)pub.shots - shot0 < QCUT_TILE_SHOTS) ? (int64_t)pub.shots - shot0 : QCUT_TILE_SHOTS);
  int nbq = (int)pub.nb_qpd;
  const uint8_t* obs = data + pub.obs_offset + (uint64_t)shot0 * NB;
  const uint8_t* qpd = data + pub.qpd_offset + (uint64_t)shot0 * nbq;
  int s0 = 0;
  uint32_t parity = 0;
  if (lane == 0) qcut_stage_issue<NB>(buf, mbar, obs, qpd, nbq, ns < G::STEP_SHOTS ? ns : G::STEP_SHOTS);

  uint32_t acc[Net::NACC];
#pragma unroll
  for (int r = 0; r < Net::NACC; ++r) acc[r] = 0;

  for (;;) {
    const int n = (ns - s0 < G::STEP_SHOTS) ? ns - s0 : G::STEP_SHOTS;
    const uint8_t* qpd_g = qpd + (uint64_t)s0 * nbq;
    const bool last_step = s0 + G::STEP_SHOTS >= ns;

    // what comes after this step: the next step of this tile, or the first step of the next tile
    const uint8_t* n_obs = obs + (uint64_t)(s0 + G::STEP_SHOTS) * NB;
    const uint8_t* n_qpd = qpd + (uint64_t)(s0 + G::STEP_SHOTS) * nbq;
    int n_cnt = ns - s0 - G::STEP_SHOTS;
    int n_nbq = nbq;
    int64_t n_shot0 = 0;
    int n_ns = 0;
    if (last_step && have_next) {
      n_shot0 = (int64_t)te_next.tile_in_pub * QCUT_TILE_SHOTS;
      const int64_t left = (int64_t)pub_next.shots - n_shot0;
      n_ns = left < QCUT_TILE_SHOTS ? (int)left : QCUT_TILE_SHOTS;
      n_nbq = (int)pub_next.nb_qpd;
      n_obs = data + pub_next.obs_offset + (uint64_t)n_shot0 * NB;
      n_qpd = data + pub_next.qpd_offset + (uint64_t)n_shot0 * n_nbq;
      n_cnt = n_ns;
    }
    if (n_cnt > G::STEP_SHOTS) n_cnt = G::STEP_SHOTS;
    const bool more = !last_step || have_next;

    // ---- consume the staged step ----
    qcut_mbar_wait(mbar, parity);
    parity ^= 1u;
    uint32_t a[G::BLOCKS][32];
    uint32_t qp[G::BATCHES];
    const QcutSrcShared src = {buf, buf + (uint32_t)QcutStage<NB>::OBS_BYTES};
    // (loads only; the arithmetic of qcut_step runs after the next copy has been issued because the
    //  compiler keeps the shared loads at the top and the volatile copy issue right behind them)
    if (n == G::STEP_SHOTS) qcut_step_loads<NB, true>(src, n, lane, a);
    else qcut_step_loads<NB, false>(src, n, lane, a);
    uint32_t qraw[G::QRAW];
    const bool q_staged = nbq == 1;
    if (q_staged) {
      if (n == G::STEP_SHOTS) qcut_qpd_loads<NB, true>(src, n, lane, qraw);
      else qcut_qpd_loads<NB, false>(src, n, lane, qraw);
    }
    __syncwarp();
    if (more && lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      qcut_stage_issue<NB>(buf, mbar, n_obs, n_qpd, n_nbq, n_cnt);
    }

    // ---- arithmetic of the current step ----
    if (q_staged) qcut_qpd_planes<NB>(qraw, qp);
    else {
#pragma unroll
      for (int u = 0; u < G::BATCHES; ++u) qp[u] = 0;
      qcut_qp_slow<NB>(qpd_g, nbq, n, lane, qp);
    }
#pragma unroll
    for (int blk = 0; blk < G::BLOCKS; ++blk) qcut_transpose32(a[blk]);
    if (n != G::STEP_SHOTS) qcut_mask_tail<NB>(n, lane, a, qp);
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u)
      Net::run(&a[0][u * G::PLANES], &a[G::BLOCKS - 1][0], qp[u], acc, variant);

    if (!last_step) {
      s0 += G::STEP_SHOTS;
      continue;
    }
    // ---- end of tile: flush, advance ----
    Net::flush(acc, ns, lane, tallies + pub.tally_offset, variant);
#pragma unroll
    for (int r = 0; r < Net::NACC; ++r) acc[r] = 0;
    if (!have_next) break;
    tile += tile_stride;
    te = te_next;
    pub = pub_next;
    variant = variant_next;
    shot0 = n_shot0;
    ns = n_ns;
    nbq = n_nbq;
    obs = n_obs;
    qpd = n_qpd;
    s0 = 0;
    have_next = tile + tile_stride < n_tiles;
    if (have_next) {
      te_next = tiles[tile + tile_stride];
      pub_next = pubs[te_next.pub];
      if (Net::MULTI) variant_next = (int)groups[pub_next.group].variant;
    }
  }
}
#endif

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
QCUT_DEV uint32_t qcut_ld4(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
QCUT_DEV uint32_t qcut_ld2(const uint8_t* p) { return __ldg(reinterpret_cast<const uint16_t*>(p)); }
QCUT_DEV uint32_t qcut_ld1(const uint8_t* p) { return __ldg(p); }
#else
QCUT_DEV qcut_u4 qcut_ld16(const uint8_t* p) { qcut_u4 v; memcpy(&v, p, 16); return v; }
QCUT_DEV uint32_t qcut_ld4(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
QCUT_DEV uint32_t qcut_ld2(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
QCUT_DEV uint32_t qcut_ld1(const uint8_t* p) { return *p; }
#endif

// In-register 32x32 bit-matrix transposition: afterwards bit r of a[i] is bit i of the old a[r].
// Two byte-granular stages are single PRMTs; the three bit-granular stages are shift + LOP3 selects.
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
    a[j] = (x & 0x0F0F0F0Fu) | ((y << 4) & 0xF0F0F0F0u);
    a[j + 4] = ((x >> 4) & 0x0F0F0F0Fu) | (y & 0xF0F0F0F0u);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j & 2) continue;
    const uint32_t x = a[j], y = a[j + 2];
    a[j] = (x & 0x33333333u) | ((y << 2) & 0xCCCCCCCCu);
    a[j + 2] = ((x >> 2) & 0x33333333u) | (y & 0xCCCCCCCCu);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j & 1) continue;
    const uint32_t x = a[j], y = a[j + 1];
    a[j] = (x & 0x55555555u) | ((y << 1) & 0xAAAAAAAAu);
    a[j + 1] = ((x >> 1) & 0x55555555u) | (y & 0xAAAAAAAAu);
  }
}

// Parity of each byte of v, delivered at bit 0 of that byte.
QCUT_DEV uint32_t qcut_byte_parity(uint32_t v) {
  v ^= v >> 4;
  v ^= v >> 2;
  v ^= v >> 1;
  return v & 0x01010101u;
}

template <int NB>
struct QcutGeo {
  static const int CHUNKS = (NB == 8) ? 16 : 8;       // 16-byte loads per lane per step
  static const int BLOCKS = (NB == 8) ? 2 : 1;        // 32x32 blocks per lane per step
  static const int BATCHES = (NB >= 4) ? 1 : 4 / NB;  // 32-shot batches per block
  static const int SPC = 16 / NB;                     // shots per chunk
  static const int RPC = (NB == 8) ? 2 : 4;           // rows per chunk (per block)
  static const int STEP_SHOTS = CHUNKS * 32 * SPC;
  static const int PLANES = 32 / BATCHES;             // registers per batch in one block
};

// Shot (relative to the step) held by row j, sub-shot u of this lane.
template <int NB>
QCUT_DEV int qcut_shot_of(int j, int u, int lane) {
  typedef QcutGeo<NB> G;
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

// One step (n <= STEP_SHOTS shots starting at obs_s / qpd_s) for one lane:
//   a[blk][.]  <- bit-planes of the observable rows of this lane's shots,
//   qp[u]      <- bit j = parity of ALL bits of the qpd row of the shot in row j, batch u
//                 (reference cutting_reconstruction.py:213: pad bits above num_bits count too).
// FULL = true: the step holds exactly STEP_SHOTS shots -- no predication, every load is
// "lane base + immediate".  FULL = false (last step of a PUB): loads are predicated and rows past
// the end of the PUB are forced to zero (they then count as "even" and are removed again by
// tally = n - 2 * odd).
template <int NB, bool FULL>
QCUT_DEV void qcut_step(const uint8_t* __restrict__ obs_s, const uint8_t* __restrict__ qpd_s, int nbq, int n,
                        int lane, uint32_t (*a)[32], uint32_t* qp) {
  typedef QcutGeo<NB> G;
  // ---- coalesced 128-bit loads of the observable rows ------------------------------------
  const uint8_t* __restrict__ obs_l = obs_s + lane * 16;
#pragma unroll
  for (int k = 0; k < G::CHUNKS; ++k) {
    qcut_u4 v = {0u, 0u, 0u, 0u};
    if (FULL || (k * 32 + lane) * G::SPC < n) v = qcut_ld16(obs_l + k * 512);
    if constexpr (NB == 8) {
      a[0][2 * k] = v.x; a[1][2 * k] = v.y;
      a[0][2 * k + 1] = v.z; a[1][2 * k + 1] = v.w;
    } else {
      a[0][4 * k] = v.x; a[0][4 * k + 1] = v.y; a[0][4 * k + 2] = v.z; a[0][4 * k + 3] = v.w;
    }
  }

  // ---- qpd parity plane(s) -----------------------------------------------------------------
#pragma unroll
  for (int u = 0; u < G::BATCHES; ++u) qp[u] = 0;
  if (nbq != 1) {
    // Any other qpd row width: per-shot byte loop (rare: more than 8 cut measurements).
    qcut_qp_slow<NB>(qpd_s, nbq, n, lane, qp);
  } else if constexpr (NB == 1) {
    // Same geometry as the observable rows: transpose the qpd block too, XOR its 8 planes.
    uint32_t q[32];
    const uint8_t* __restrict__ qpd_l = qpd_s + lane * 16;
#pragma unroll
    for (int k = 0; k < G::CHUNKS; ++k) {
      qcut_u4 v = {0u, 0u, 0u, 0u};
      if (FULL || (k * 32 + lane) * G::SPC < n) v = qcut_ld16(qpd_l + k * 512);
      q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
    }
    qcut_transpose32(q);
#pragma unroll
    for (int u = 0; u < G::BATCHES; ++u) {
      uint32_t x = 0;
#pragma unroll
      for (int b = 0; b < G::PLANES; ++b) x ^= q[u * G::PLANES + b];
      qp[u] = x;
    }
  } else {
    // One qpd byte per shot: fold each byte to its parity, then gather the parity bits of
    // four bytes with one multiply (no carries: all partial products hit distinct bits).
    const uint8_t* __restrict__ qpd_l = qpd_s + lane * G::SPC;  // SPC qpd bytes per chunk per lane
#pragma unroll
    for (int k = 0; k < G::CHUNKS; ++k) {
      const int idx = k * 32 + lane;
      if constexpr (NB == 8) {
        if (k & 1) continue;  // chunks k, k+1 -> rows 2k .. 2k+3
        uint32_t w = 0;
        if (FULL || idx * 2 < n) w = qcut_ld2(qpd_l + k * 64);
        if (FULL || (idx + 32) * 2 < n) w |= qcut_ld2(qpd_l + (k + 1) * 64) << 16;
        const uint32_t nib = (qcut_byte_parity(w) * 0x01020408u) >> 24;
        qp[0] |= nib << (2 * k);
      } else if constexpr (NB == 4) {
        uint32_t w = 0;
        if (FULL || idx * 4 < n) w = qcut_ld4(qpd_l + k * 128);
        const uint32_t nib = (qcut_byte_parity(w) * 0x01020408u) >> 24;
        qp[0] |= nib << (4 * k);
      } else {  // NB == 2: bytes (b0, b2) of each word -> batch 0, (b1, b3) -> batch 1
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w = 0;
          if (FULL || idx * 8 + 4 * h < n) w = qcut_ld4(qpd_l + k * 256 + 4 * h);
          const uint32_t r = (qcut_byte_parity(w) * 0x01100220u) >> 24;
          qp[0] |= (r & 3u) << (4 * k + 2 * h);
          qp[G::BATCHES - 1] |= ((r >> 4) & 3u) << (4 * k + 2 * h);
        }
      }
    }
  }

  // ---- bit-slice ---------------------------------------------------------------------------
#pragma unroll
  for (int blk = 0; blk < G::BLOCKS; ++blk) qcut_transpose32(a[blk]);

  if constexpr (!FULL) {
    // Tail step: rows past the end of the PUB may hold bytes of the padding / next matrix.
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
}

template <int NB>
QCUT_DEV void qcut_step_any(const uint8_t* __restrict__ obs_s, const uint8_t* __restrict__ qpd_s, int nbq, int n,
                            int lane, uint32_t (*a)[32], uint32_t* qp) {
  if (n == QcutGeo<NB>::STEP_SHOTS) qcut_step<NB, true>(obs_s, qpd_s, nbq, n, lane, a, qp);
  else qcut_step<NB, false>(obs_s, qpd_s, nbq, n, lane, a, qp);
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
template <int T>
QCUT_DEV void qcut_flush(const uint32_t* acc, int ns, int lane, unsigned long long* __restrict__ tally) {
#pragma unroll
  for (int r = 0; r < (T + 3) / 4; ++r) {
    // Widen the four 8-bit fields to two pairs of 16-bit fields so the warp sum (<= 4096) fits.
    const uint32_t ev = __reduce_add_sync(0xffffffffu, acc[r] & 0x00FF00FFu);         // terms 4r, 4r+2
    const uint32_t od = __reduce_add_sync(0xffffffffu, (acc[r] >> 8) & 0x00FF00FFu);  // terms 4r+1, 4r+3
    const int t = 4 * r + lane;
    if (lane < 4 && t < T) {
      const uint32_t pair = (lane & 1) ? od : ev;
      const long long odd = (lane & 2) ? (pair >> 16) : (pair & 0xFFFFu);
      qcut_red_add(tally + t, (unsigned long long)((long long)ns - 2 * odd));
    }
  }
}
#endif

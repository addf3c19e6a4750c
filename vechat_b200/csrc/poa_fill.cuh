// poa_fill.cuh — the sequence-to-DAG DP fill for sm_100a: one warp per alignment, packed int16x2 cells.
//
// Replaces SimdAlignmentEngine::Linear's fill (vendor/spoa/src/simd_alignment_engine_implementation.hpp:
// 760-906, scalar twin sisd_alignment_engine.cpp:292-360) and Initialize (:506-681):
//   H[i][j] = max over predecessors p of max(H[p][j-1] + s(i,j), H[p][j] + g), then
//   H[i][j] = max(H[i][j], H[i][j-1] + g)              (a max-plus prefix scan along the row)
//   SW clamps at 0 and tracks the first row (rank order) / first column of the global maximum;
//   NW ends at the first sink row with the best last-column score.
//
// Mapping.  A row of 64*K cells lives in K 32-bit registers per lane, two cells per register:
// low half = column lane*K + k, high half = column 32*K + lane*K + k.  Every lane therefore owns two
// runs of K consecutive columns and the two halves of a register never depend on each other, so the
// whole recurrence runs on Blackwell's packed DPX integer ops (SASS VIADDMNMX.S16x2, VIMNMX.S16x2,
// VIADD.16x2): one instruction per two cells.
//   diagonal   : register k-1 of the predecessor row (lane boundary: one shuffle)
//   horizontal : in-register running max over the lane's K cells, then a 5-step warp-shuffle max-scan of
//                (segment end value - g * column) across the 64 segments, then one fused add-max per register
// Rows are written once to HBM (2 B per cell, lane-major: lane l stores its K words at [l*K, l*K + K), so a run
// of columns is a run of words for the traceback's tiles) because the traceback re-reads them.  The row width is
// chosen per alignment (fill_width: K = 8 for layers up to 512 columns).  Predecessor rows come, in order of
// preference, from registers (distance 1: the row just computed, the common case along chains, updated in place),
// from a ring of the four most recent rows in shared memory (slot = row mod 4), or from HBM/L2.
#ifndef VGC_POA_FILL_CUH_
#define VGC_POA_FILL_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "poa_core.h"

namespace vgc {

#ifndef VGC_RING_ROWS
#define VGC_RING_ROWS 4
#endif
constexpr int kRingRows = VGC_RING_ROWS;  // recent rows kept in shared memory (power of two; row r lives in slot r % kRingRows)

__device__ __forceinline__ uint32_t pack16(int32_t lo, int32_t hi) {
  return (static_cast<uint32_t>(lo) & 0xFFFFu) | (static_cast<uint32_t>(hi) << 16);
}
__device__ __forceinline__ int32_t lo16(uint32_t w) { return static_cast<int16_t>(w & 0xFFFFu); }
__device__ __forceinline__ int32_t hi16(uint32_t w) { return static_cast<int16_t>(w >> 16); }

// keep a loop-invariant value in its register: without this ptxas recomputes the packed per-lane constants in every
// row (cheaper in registers, ~40 extra instructions per row)
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}

// K words of this lane to global memory (`p` already offset to the lane's words), as st.global vectors
template <int K>
__device__ __forceinline__ void lane_store_global(uint32_t* p, const uint32_t (&h)[K]) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int k = 0; k < K; k += 4)
      asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p + k), "r"(h[k]), "r"(h[k + 1]), "r"(h[k + 2]), "r"(h[k + 3]) : "memory");
  } else {
#pragma unroll
    for (int k = 0; k < K; k += 2)
      asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p + k), "r"(h[k]), "r"(h[k + 1]) : "memory");
  }
}
template <int K>
__device__ __forceinline__ void lane_load_global(const uint32_t* p, uint32_t (&u)[K]) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int k = 0; k < K; k += 4)
      asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u[k]), "=r"(u[k + 1]), "=r"(u[k + 2]), "=r"(u[k + 3]) : "l"(p + k) : "memory");
  } else {
#pragma unroll
    for (int k = 0; k < K; k += 2)
      asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(u[k]), "=r"(u[k + 1]) : "l"(p + k) : "memory");
  }
}

// K words of this lane from / to shared memory at `p` (already offset to the lane's words): 16-byte vectors when the
// lane stride allows.  With K = 8 (16) the lane stride is 32 (64) bytes, so the eight lanes of a quarter warp that one
// LDS.128 / STS.128 wavefront serves would fall on 4 (2) distinct bank groups: the lane's 16-byte chunks are rotated
// by `sw` chunks (lane_swizzle: a per-lane constant) so that a wavefront covers all eight groups.  Only the owning
// lane ever reads what it wrote, so the permutation needs no agreement with anybody else.
template <int K>
__device__ __forceinline__ int lane_swizzle(int lane) {
  if constexpr (K == 8) return (lane >> 2) & 1;
  if constexpr (K == 16) return (lane >> 1) & 3;
  return 0;
}
template <int K>
__device__ __forceinline__ void lane_load(const uint32_t* __restrict__ p, uint32_t (&u)[K], int sw = 0) {
  if constexpr (K % 4 == 0) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int k = 0; k < K; k += 4) {
      uint4 t = q[((k >> 2) + sw) & (K / 4 - 1)];
      u[k] = t.x;
      u[k + 1] = t.y;
      u[k + 2] = t.z;
      u[k + 3] = t.w;
    }
  } else {
    const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int k = 0; k < K; k += 2) {
      uint2 t = q[k >> 1];
      u[k] = t.x;
      u[k + 1] = t.y;
    }
  }
}

template <int K>
__device__ __forceinline__ void lane_store(uint32_t* __restrict__ p, const uint32_t (&h)[K], int sw = 0) {
  if constexpr (K % 4 == 0) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int k = 0; k < K; k += 4) q[((k >> 2) + sw) & (K / 4 - 1)] = make_uint4(h[k], h[k + 1], h[k + 2], h[k + 3]);
  } else {
    uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
    for (int k = 0; k < K; k += 2) q[k >> 1] = make_uint2(h[k], h[k + 1]);
  }
}

// The matrix of one alignment: rows of 32*K words (row = rank + 1, row 0 = the virtual row) in the align kernel's
// private scratch buffer, first-column values next to it; the row program and its overflow list come from the
// window's slot.
struct FillIo {
  uint32_t* H;             // [(nR + 1) * 32*K] words
  int16_t* fc;             // [nR + 1]
  const uint32_t* rowprog; // [nR * 4]
  const uint32_t* ovf;
  uint32_t nR;
  // where the traceback starts (out)
  uint32_t best_row, best_col;
  int32_t best_score;
};

// prof : shared memory, num_codes * 32*K words;  stage: shared memory, 32 uint4
// ring : shared memory, ring_rows * 32*K words (ring_rows <= kRingRows, may be 0)
template <int K, bool SW>
__device__ __forceinline__ void warp_fill_t(FillIo& io, const uint8_t* codes, uint32_t len, const Scores sc,
                            uint32_t num_codes, uint32_t* prof, uint4* stage, uint32_t* ring, int ring_rows) {
  static_assert(K % 2 == 0, "K must be even");
  using RM = RowMap<K>;
  const int lane = threadIdx.x & 31;
  const uint32_t nR = io.nR;
  const int32_t g = sc.g;
  uint32_t* const Hm = io.H;
  int16_t* const fcm = io.fc;
  const uint32_t* const ovfm = io.ovf;
  constexpr uint32_t rw = RM::kWords;

  // ---- query profile (Initialize, simd...:520-530): per code, match/mismatch per column, padding beyond len
  {
    int32_t pad = sc.m > -sc.x ? sc.m : -sc.x;
    if (-g > pad) pad = -g;
    pad = -pad;
    uint32_t cl[K], ch[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const uint32_t a = lane * K + k, b = 32 * K + lane * K + k;
      cl[k] = a < len ? codes[a] : 0xFFu;
      ch[k] = b < len ? codes[b] : 0xFFu;
    }
    const int swz0 = lane_swizzle<K>(lane);
    for (uint32_t c = 0; c < num_codes; ++c) {
      uint32_t pw[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int32_t vl = cl[k] == 0xFFu ? pad : (cl[k] == c ? sc.m : sc.x);
        const int32_t vh = ch[k] == 0xFFu ? pad : (ch[k] == c ? sc.m : sc.x);
        pw[k] = pack16(vl, vh);
      }
      lane_store<K>(prof + c * RM::kWords + lane * K, pw, swz0);
    }
  }
  // ---- per-lane constants
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int32_t c0l = lane * K, c0h = 32 * K + lane * K;
  const uint32_t g2 = pinned(pack16(g, g));
  const uint32_t voff = pinned(pack16(-g * (c0l + K - 1), -g * (c0h + K - 1)));
  const uint32_t gbase = pinned(pack16(g * c0l, g * c0h));
  uint32_t gk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) gk[k] = pinned(pack16(g * k, g * k));
  const int swz = lane_swizzle<K>(lane);
  const int rot = (lane + 31) & 31;              // the lane to my left (lane 0: lane 31, whose low half feeds my high half)
  const uint32_t* const profl = prof + lane * K;  // my words of the profile rows / ring rows / matrix rows
  uint32_t* const ringl = ring + lane * K;
  uint32_t* hrow = Hm + lane * K;

  // ---- virtual row 0 (NW: j * g; SW: zeros) and its first column
  uint32_t hp[K];  // the row computed last (registers); chain rows are updated in place
#pragma unroll
  for (int k = 0; k < K; ++k) hp[k] = SW ? 0u : pack16(g * (c0l + k + 1), g * (c0h + k + 1));
  lane_store_global<K>(hrow, hp);
  if (lane == 0) fcm[0] = 0;
  int32_t fc_prev = 0;

  // ring of the most recent rows in shared memory: row r lives in slot r % kRingRows (with its first-column value
  // in ring_fc), so "is predecessor row - d in the ring" is just d <= kRingRows — no tags to search or maintain
  const bool use_ring = ring_rows == kRingRows;
  // (every lane keeps its own copy of the first-column values: a lane only ever reads what it wrote itself, rows
  //  and first columns alike, so the ring needs no warp synchronisation)
  int32_t* ring_fc = reinterpret_cast<int32_t*>(ring + kRingRows * rw) + lane;
  if (use_ring) {
    lane_store<K>(ringl, hp, swz);
    ring_fc[0] = 0;
  }

  // ---- best-cell tracking
  uint32_t bestv = 0;                   // SW: per-lane packed running max (scores >= 0)
  uint32_t bestr_lo = 0, bestr_hi = 0;  // rank at which each half first reached it
  int32_t nw_best = INT32_MIN;
  uint32_t nw_row = 0;
  const uint32_t lc = len - 1;
  const int lastH = lc / (32 * K), lastL = (lc % (32 * K)) / K, lastK = lc % K;

  const uint4* rp = reinterpret_cast<const uint4*>(io.rowprog);
  uint4 nxt = make_uint4(0, 0, 0, 0);
  if (static_cast<uint32_t>(lane) < nR) nxt = rp[lane];

  uint32_t row = 0;
  for (uint32_t r0 = 0; r0 < nR; r0 += 32) {
    __syncwarp();
    stage[lane] = nxt;
    if (r0 + 32 + lane < nR) nxt = rp[r0 + 32 + lane];
    __syncwarp();
    const uint32_t rn = nR - r0 < 32 ? nR - r0 : 32;
    uint4 e_ahead = stage[0];
#pragma unroll 1
    for (uint32_t rr = 0; rr < rn; ++rr) {
      // rows live in rank space: this is row r0 + rr + 1; e = {meta, p0p1, p2p3, p4p5 | ovf offset}, predecessors as
      // row distances
      const uint4 e = e_ahead;
      e_ahead = stage[(rr + 1) & 31];  // the next row's record, asked for a whole row early (rr = 31: reloaded below)
      const uint32_t meta = e.x;
      ++row;
      hrow += rw;
      uint32_t pr[K];
      lane_load<K>(profl + meta_code(meta) * rw, pr, swz);
      int32_t fcmax;
      // fetch predecessor row `row - d` (d >= 2) and its first-column value: the ring of recent rows, else L2 / HBM
      auto fetch = [&](uint32_t d, uint32_t (&u)[K], int32_t& fcp) {
        if (use_ring && d <= static_cast<uint32_t>(kRingRows)) {
          const uint32_t slot = (row - d) & (kRingRows - 1);
          lane_load<K>(ringl + slot * rw, u, swz);
          fcp = ring_fc[slot * 32];
        } else {
          lane_load_global<K>(hrow - static_cast<uint64_t>(d) * rw, u);
          fcp = 0;
          if (!SW) {  // lane 0 owns fc[] (it wrote it): read there, broadcast
            if (lane == 0) fcp = static_cast<int32_t>(fcm[row - d]);
            fcp = __shfl_sync(FULL, fcp, 0);
          }
        }
      };
      if ((meta & (0x7F80u | kMetaInline)) == kMetaInline) {
        // ---- at most one predecessor (78 % of the rows): bring it into the row registers unless it is the row just
        //      computed (distance 1), then update in place (descending k) — no copies, no loop over predecessors
        const uint32_t d = e.y & 0xFFFFu;
        if (d != 1) fetch(d, hp, fc_prev);
        const uint32_t y = __shfl_sync(FULL, hp[K - 1], rot);
        const uint32_t x = lane == 0 ? __byte_perm(static_cast<uint32_t>(fc_prev), y, 0x5410) : y;
#pragma unroll
        for (int k = K - 1; k >= 1; --k) hp[k] = __viaddmax_s16x2(hp[k - 1], pr[k], __vadd2(hp[k], g2));
        hp[0] = __viaddmax_s16x2(x, pr[0], __vadd2(hp[0], g2));
        fcmax = fc_prev;
        // horizontal, in-lane part: running max over the lane's cells
#pragma unroll
        for (int k = 1; k < K; ++k) hp[k] = __viaddmax_s16x2(hp[k - 1], g2, hp[k]);
      } else {
        // ---- several predecessors: maximum over all of them, each from registers, the ring or memory
        const uint32_t np = meta_npred(meta);
        const U4 er = {e.x, e.y, e.z, e.w};
        uint32_t acc[K], u[K];
        int32_t fcp;
        // diagonal of this lane's first cells: the previous lane's last cells
        auto pred_first = [&](const uint32_t (&v)[K], int32_t fcv) {
          const uint32_t y = __shfl_sync(FULL, v[K - 1], rot);
          const uint32_t x = lane == 0 ? __byte_perm(static_cast<uint32_t>(fcv), y, 0x5410) : y;
          acc[0] = __viaddmax_s16x2(x, pr[0], __vadd2(v[0], g2));
#pragma unroll
          for (int k = 1; k < K; ++k) acc[k] = __viaddmax_s16x2(v[k - 1], pr[k], __vadd2(v[k], g2));
          fcmax = fcv;
        };
        auto pred_more = [&](const uint32_t (&v)[K], int32_t fcv) {
          const uint32_t y = __shfl_sync(FULL, v[K - 1], rot);
          const uint32_t x = lane == 0 ? __byte_perm(static_cast<uint32_t>(fcv), y, 0x5410) : y;
          acc[0] = __viaddmax_s16x2(v[0], g2, __viaddmax_s16x2(x, pr[0], acc[0]));
#pragma unroll
          for (int k = 1; k < K; ++k) acc[k] = __viaddmax_s16x2(v[k], g2, __viaddmax_s16x2(v[k - 1], pr[k], acc[k]));
          fcmax = fcv > fcmax ? fcv : fcmax;
        };
        if (meta & kMetaInline) {
          // two to six predecessors, distances in the record; the first two without any indexing
          const uint32_t d0 = e.y & 0xFFFFu, d1 = e.y >> 16;
          if (d0 == 1) {
            pred_first(hp, fc_prev);
          } else {
            fetch(d0, u, fcp);
            pred_first(u, fcp);
          }
          if (d1 == 1) {
            pred_more(hp, fc_prev);
          } else {
            fetch(d1, u, fcp);
            pred_more(u, fcp);
          }
#pragma unroll 1
          for (uint32_t p = 2; p < np; ++p) {
            const uint32_t d = rec_delta(er, p);
            if (d == 1) {
#pragma unroll
              for (int k = 0; k < K; ++k) u[k] = hp[k];
              fcp = fc_prev;
            } else {
              fetch(d, u, fcp);
            }
            pred_more(u, fcp);
          }
        } else {
          // rare: more than six predecessors or one further than 65535 rows up (list in ovf[]); np == 0: the virtual row
          const uint32_t npp = np == 0 ? 1 : np;
#pragma unroll 1
          for (uint32_t p = 0; p < npp; ++p) {
            const uint32_t d = np == 0 ? row : row - ovfm[e.w + p];
            if (d == 1) {
#pragma unroll
              for (int k = 0; k < K; ++k) u[k] = hp[k];
              fcp = fc_prev;
            } else {
              fetch(d, u, fcp);
            }
            if (p == 0) pred_first(u, fcp);
            else pred_more(u, fcp);
          }
        }
        // horizontal, in-lane part (moves the row into the row registers on the way)
        hp[0] = acc[0];
#pragma unroll
        for (int k = 1; k < K; ++k) hp[k] = __viaddmax_s16x2(hp[k - 1], g2, acc[k]);
      }
      const int32_t fci = SW ? 0 : fcmax + g;
      // ---- horizontal, cross-lane part: the max-plus scan over the lanes' segments
      uint32_t V = __vadd2(hp[K - 1], voff);
      // (round 2: a doubling inclusive scan, 5 levels + 2 shuffles behind it, was the longest dependent chain of a row;
      //  radix-4 inclusive: fill cycles -2.8 %; this exclusive form + REDUX + the record fetched a row ahead: a further
      //  -7.5 %, pass 902 -> 883 ms — profiles/r02_sweep_scan_radix4.txt, r02_sweep_scan_exclusive.txt)
      // EXCLUSIVE max-scan over the lanes in three dependent levels (windows of 4, 16, 32 lanes to the left): the
      // carry-in of a lane is ready right after the scan, without the extra shuffle an inclusive scan needs, and the
      // total of the low halves (the carry into the high half) comes from one REDUX beside the scan instead of a
      // shuffle behind it.  Level 1 masks the lanes that do not exist (shfl_up hands a lane its own value back);
      // at levels 2 and 3 getting the own partial result back is harmless (max is idempotent).
      const int32_t lowtot = __reduce_max_sync(FULL, lo16(V));
      uint32_t E;
      {
        constexpr uint32_t kNeg = 0x80008000u;
        uint32_t a = __shfl_up_sync(FULL, V, 1), b = __shfl_up_sync(FULL, V, 2), c = __shfl_up_sync(FULL, V, 3),
                 d = __shfl_up_sync(FULL, V, 4);
        a = lane >= 1 ? a : kNeg;
        b = lane >= 2 ? b : kNeg;
        c = lane >= 3 ? c : kNeg;
        d = lane >= 4 ? d : kNeg;
        E = __vmaxs2(__vimax3_s16x2(a, b, c), d);
      }
      {
        const uint32_t a = __shfl_up_sync(FULL, E, 4), b = __shfl_up_sync(FULL, E, 8), c = __shfl_up_sync(FULL, E, 12);
        E = __vmaxs2(__vimax3_s16x2(E, a, b), c);
      }
      E = __vmaxs2(E, __shfl_up_sync(FULL, E, 16));
      const int32_t vfc = fci + g;
      const uint32_t X = pack16(vfc, vfc > lowtot ? vfc : lowtot);
      E = __vmaxs2(E, X);  // lane 0: max(kNeg, X) = X
      const uint32_t base = __vadd2(E, gbase);
      if (SW) {
#pragma unroll
        for (int k = 0; k < K; ++k) hp[k] = __viaddmax_s16x2_relu(base, gk[k], hp[k]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) hp[k] = __viaddmax_s16x2(base, gk[k], hp[k]);
      }
      // ---- write the row once (HBM) and keep it in the ring
      lane_store_global<K>(hrow, hp);
      if (!SW && lane == 0) fcm[row] = static_cast<int16_t>(fci);
      if (use_ring) {
        const uint32_t slot = row & (kRingRows - 1);
        lane_store<K>(ringl + slot * rw, hp, swz);
        ring_fc[slot * 32] = fci;
      }
      // ---- best cell
      if (SW) {
        uint32_t m = hp[0];
#pragma unroll
        for (int k = 1; k < K; ++k) m = __vmaxs2(m, hp[k]);
        const uint32_t nb = __vmaxs2(bestv, m);
        const uint32_t chg = nb ^ bestv;
        if (chg & 0xFFFFu) bestr_lo = row - 1;
        if (chg >> 16) bestr_hi = row - 1;
        bestv = nb;
      } else if (meta & kMetaSink) {
        uint32_t sel = hp[0];
#pragma unroll
        for (int k = 1; k < K; ++k) {
          if (k == lastK) sel = hp[k];
        }
        const uint32_t s = __shfl_sync(FULL, sel, lastL);
        const int32_t val = lastH ? hi16(s) : lo16(s);
        if (val > nw_best) {
          nw_best = val;
          nw_row = row;
        }
      }
      fc_prev = fci;
    }
  }

  // ---- where the traceback starts
  if (!SW) {
    io.best_row = nw_row;
    io.best_col = nw_row ? len : 0;
    io.best_score = nw_best;
  } else {
    // global max, then the first row in rank order that reached it, then its first column
    int32_t mx = lo16(bestv) > hi16(bestv) ? lo16(bestv) : hi16(bestv);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const int32_t o = __shfl_xor_sync(0xFFFFFFFFu, mx, d);
      mx = o > mx ? o : mx;
    }
    uint32_t br = 0xFFFFFFFFu;
    if (mx > 0) {
      if (lo16(bestv) == mx) br = bestr_lo;
      if (hi16(bestv) == mx && bestr_hi < br) br = bestr_hi;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, br, d);
      br = o < br ? o : br;
    }
    uint32_t brow = 0, col = 0;
    if (mx > 0) {
      brow = br + 1;
      uint32_t u[K];
      __syncwarp();
      lane_load<K>(Hm + static_cast<uint64_t>(brow) * rw + lane * K, u);
      uint32_t bc = 0xFFFFFFFFu;
#pragma unroll
      for (int k = K - 1; k >= 0; --k) {
        if (hi16(u[k]) == mx) bc = c0h + k;
      }
#pragma unroll
      for (int k = K - 1; k >= 0; --k) {
        if (lo16(u[k]) == mx) bc = c0l + k;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, bc, d);
        bc = o < bc ? o : bc;
      }
      col = bc + 1;
    }
    io.best_row = brow;
    io.best_col = col;
    io.best_score = mx;
  }
  __syncwarp();
}

template <int K>
__device__ __forceinline__ void warp_fill(FillIo& io, const uint8_t* codes, uint32_t len,
                                          uint32_t mode, const Scores sc, uint32_t num_codes, uint32_t* prof,
                                          uint4* stage, uint32_t* ring, int ring_rows) {
  if (mode == kModeSW) warp_fill_t<K, true>(io, codes, len, sc, num_codes, prof, stage, ring, ring_rows);
  else warp_fill_t<K, false>(io, codes, len, sc, num_codes, prof, stage, ring, ring_rows);
}

}  // namespace vgc

#endif  // VGC_POA_FILL_CUH_

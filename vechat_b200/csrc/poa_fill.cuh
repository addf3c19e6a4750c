// poa_fill.cuh — the sequence-to-DAG DP fill for sm_100a: one warp per alignment, packed int16x2 cells, lanes
// skewed along the rows (a 32-stage systolic wavefront).
//
// Replaces SimdAlignmentEngine::Linear's fill (vendor/spoa/src/simd_alignment_engine_implementation.hpp:
// 760-906, scalar twin sisd_alignment_engine.cpp:292-360) and Initialize (:506-681):
//   H[i][j] = max over predecessors p of max(H[p][j-1] + s(i,j), H[p][j] + g), then
//   H[i][j] = max(H[i][j], H[i][j-1] + g)              (a max-plus prefix scan along the row)
//   SW clamps at 0 and tracks the first row (rank order) / first column of the global maximum;
//   NW ends at the first sink row with the best last-column score.
//
// Mapping.  Lane l owns the 2K consecutive columns [2K*l, 2K*l + 2K) of every row, as K packed registers: register k
// holds column 2K*l + k in its low half and column 2K*l + K + k in its high half, so the two halves of a register
// never depend on each other inside the predecessor step and the whole recurrence runs on Blackwell's packed DPX
// integer ops (SASS VIADDMNMX.S16x2, VIMNMX.S16x2, VIADD.16x2): one instruction per two cells.
// At step t lane l computes row t - l: the lane to its left finished the same row one step earlier, so the only
// value that crosses lanes is "the final score of the last column of my left neighbour for this row" — one shuffle
// per step, issued at the top of the step and consumed at its end.  Nothing else does: a lane reads predecessor rows
// only in its own columns (plus the left value it received for them, which it keeps), from registers (distance 1:
// the row it computed in the previous step), from its own slice of a ring of recent rows in shared memory, or from
// its own words in the matrix buffer (L2/HBM).  There is no cross-lane scan and no warp-wide barrier in a step.
//   predecessors : h[k] = max_p max(u_p[k-1] + profile[k], u_p[k] + g)       2 DPX ops per register and predecessor
//   horizontal   : in-half running max (K-1 dependent DPX ops), then one fused add-max per register injects the left
//                  value E (low half: E + g(k+1); high half: the final value of the low half's last column + g(k+1))
// Rows with different in-degrees meet in one step (lanes are at different rows): the predecessor loop runs to the
// warp's maximum, lanes with fewer predecessors idle in the extra iterations.
// Rows are written once (2 B per cell) because the traceback re-reads them, in a SKEWED layout: the K words lane l
// computes for row r go to words [l*K, l*K + K) of memory row r + l.  What the 32 lanes write in one step (rows t,
// t-1, .. t-31) is therefore one contiguous memory row — fully coalesced — and what they read back from
// predecessor rows a few rows up is nearly so.  64 consecutive columns of a row are 32 consecutive words (of
// memory rows r + l) for the traceback's tiles.  The row width is chosen per alignment (fill_width: K = 8 for
// layers up to 512 columns).
#ifndef VGC_POA_FILL_CUH_
#define VGC_POA_FILL_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "poa_core.h"

namespace vgc {

#ifndef VGC_RING_ROWS
#define VGC_RING_ROWS 8
#endif
constexpr int kRingRows = VGC_RING_ROWS;  // recent rows kept in shared memory (power of two; row r lives in slot r % kRingRows)
constexpr uint32_t kRecRing = 64;         // row records staged in shared memory (two blocks of 32)

__device__ __forceinline__ uint32_t pack16(int32_t lo, int32_t hi) {
  return (static_cast<uint32_t>(lo) & 0xFFFFu) | (static_cast<uint32_t>(hi) << 16);
}
__device__ __forceinline__ int32_t lo16(uint32_t w) { return static_cast<int16_t>(w & 0xFFFFu); }
__device__ __forceinline__ int32_t hi16(uint32_t w) { return static_cast<int16_t>(w >> 16); }

template <int K>
__device__ __forceinline__ void row_load(const uint32_t* __restrict__ row, int lane, uint32_t (&u)[K]) {
  const uint2* p = reinterpret_cast<const uint2*>(row + lane * K);  // lane-major: K consecutive words per lane
#pragma unroll
  for (int k = 0; k < K; k += 2) {
    uint2 t = p[k >> 1];
    u[k] = t.x;
    u[k + 1] = t.y;
  }
}

template <int K>
__device__ __forceinline__ void row_store(uint32_t* __restrict__ row, int lane, const uint32_t (&h)[K]) {
  uint2* p = reinterpret_cast<uint2*>(row + lane * K);
#pragma unroll
  for (int k = 0; k < K; k += 2) p[k >> 1] = make_uint2(h[k], h[k + 1]);
}

// The matrix of one alignment: rows of 32*K words (row = rank + 1, row 0 = the virtual row) in the align kernel's
// private scratch buffer, and next to it the value to the left of every lane's columns per row (lane 0: the first
// column, the NW border); the row program and its overflow list come from the window's slot.
constexpr uint32_t kSkewRows = 31;  // memory rows beyond nR + 1 the skewed layout needs

struct FillIo {
  uint32_t* H;             // [(nR + 1 + kSkewRows) * 32*K] words: block l of row r at memory row r + l
  int16_t* left;           // [(nR + 1 + kSkewRows) * 32]: left[(r + l) * 32 + l] = H(r, DP column 2K*l) — l = 0: the first column
  const uint32_t* rowprog; // [nR * 4]
  const uint32_t* ovf;
  uint32_t nR;
  // where the traceback starts (out)
  uint32_t best_row, best_col;
  int32_t best_score;
};

// prof : shared memory, num_codes * 32*K words;  recs: shared memory, kRecRing row records
// ring : shared memory, ring_rows * (32*K + 32) words (ring_rows == kRingRows or 0)
template <int K, bool SW>
__device__ void wave_fill_t(FillIo& io, const uint8_t* codes, uint32_t len, const Scores sc, uint32_t num_codes,
                            uint32_t* prof, U4* recs, uint32_t* ring, int ring_rows) {
  static_assert(K % 2 == 0, "K must be even");
  constexpr uint32_t rw = 32 * K;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const uint32_t nR = io.nR;
  const int32_t g = sc.g;
  uint32_t* const Hm = io.H;
  int16_t* const leftm = io.left;
  const uint32_t* const ovfm = io.ovf;
  const uint32_t c0 = 2u * K * lane;  // first column of this lane

  // ---- query profile (Initialize, simd...:520-530): per code, match/mismatch per column, padding beyond len
  {
    int32_t pad = sc.m > -sc.x ? sc.m : -sc.x;
    if (-g > pad) pad = -g;
    pad = -pad;
    uint32_t cl[K], ch[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const uint32_t a = c0 + k, b = c0 + K + k;
      cl[k] = a < len ? codes[a] : 0xFFu;
      ch[k] = b < len ? codes[b] : 0xFFu;
    }
    for (uint32_t c = 0; c < num_codes; ++c) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int32_t vl = cl[k] == 0xFFu ? pad : (cl[k] == c ? sc.m : sc.x);
        const int32_t vh = ch[k] == 0xFFu ? pad : (ch[k] == c ? sc.m : sc.x);
        prof[c * rw + lane * K + k] = pack16(vl, vh);
      }
    }
  }
  // ---- per-lane constants
  const uint32_t g2 = pack16(g, g);
  uint32_t gk1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) gk1[k] = pack16(g * (k + 1), g * (k + 1));
  const int32_t gK = g * K;

  // ---- virtual row 0 (NW: j * g; SW: zeros), the value to its left, and its last column
  uint32_t hp[K];  // the row this lane computed last
#pragma unroll
  for (int k = 0; k < K; ++k)
    hp[k] = SW ? 0u : pack16(g * static_cast<int32_t>(c0 + k + 1), g * static_cast<int32_t>(c0 + K + k + 1));
  row_store<K>(Hm + static_cast<uint64_t>(lane) * rw, lane, hp);  // skewed: row 0 of lane l at memory row l
  int32_t E_prev = SW ? 0 : g * static_cast<int32_t>(c0);  // H(0, DP column c0); lane 0: the first column = 0
  leftm[lane * 32 + lane] = static_cast<int16_t>(E_prev);
  int32_t out_last = hi16(hp[K - 1]);

  // ring of this lane's most recent rows in shared memory: row r lives in slot r % kRingRows, together with the left
  // value the lane received for it.  A lane only ever reads what it wrote itself: no synchronisation.
  const bool use_ring = ring_rows == kRingRows;
  int32_t* ring_left = reinterpret_cast<int32_t*>(ring + kRingRows * rw) + lane;
  if (use_ring) {
    row_store<K>(ring, lane, hp);
    ring_left[0] = E_prev;
  }

  // ---- best-cell tracking
  uint32_t bestv = 0;                   // SW: per-lane packed running max (scores >= 0)
  uint32_t bestr_lo = 0, bestr_hi = 0;  // row at which each half first reached it
  int32_t nw_best = INT32_MIN;
  uint32_t nw_row = 0;
  const uint32_t lc = len - 1;
  const int lastL = lc / (2 * K), lastH = (lc % (2 * K)) / K, lastK = lc % K;

  // ---- row records: 64-entry ring in shared memory, refreshed one block of 32 at a time; record q at recs[q & 63]
  const U4* rp = reinterpret_cast<const U4*>(io.rowprog);
  const U4 zero4 = {0, 0, 0, 0};
  U4 nxt = zero4;
  if (static_cast<uint32_t>(lane) < nR) recs[lane] = rp[lane];
  if (32u + lane < nR) nxt = rp[32 + lane];
  __syncwarp();

  const uint32_t T = nR + 31;
  for (uint32_t t = 1; t <= T; ++t) {
    if ((t & 31u) == 0) {
      __syncwarp();
      recs[(t & 32u) + lane] = nxt;  // records t .. t+31 (rows t+1 ..): first used by lane 0 in the next step
      __syncwarp();
      nxt = t + 32u + lane < nR ? rp[t + 32u + lane] : zero4;
    }
    const int32_t r = static_cast<int32_t>(t) - lane;  // this lane's row in this step
    const bool act = r >= 1 && r <= static_cast<int32_t>(nR);
    // the left neighbour finished row r one step ago: its last column is the value to the left of my columns
    int32_t E = __shfl_up_sync(FULL, out_last, 1);
    const U4 e = act ? recs[(r - 1) & 63] : zero4;
    const uint32_t meta = e.x;
    const uint32_t np = meta_npred(meta);
    const uint32_t npp = act ? (np == 0 ? 1u : np) : 0u;
    const uint32_t npmax = __reduce_max_sync(FULL, npp);
    const bool inl = (meta & kMetaInline) != 0;
    uint32_t pr[K];
    if (act) {
      const uint2* pp = reinterpret_cast<const uint2*>(prof + meta_code(meta) * rw + lane * K);
#pragma unroll
      for (int k = 0; k < K; k += 2) {
        uint2 v = pp[k >> 1];
        pr[k] = v.x;
        pr[k + 1] = v.y;
      }
    }
    uint32_t h[K];
    int32_t lmax = INT32_MIN;
    for (uint32_t p = 0; p < npmax; ++p) {
      if (p < npp) {
        // distance to predecessor p (rows are processed in rank order: distance 1 = the row in registers)
        const uint32_t d = np == 0 ? static_cast<uint32_t>(r)
                                   : (inl ? rec_delta(e, p) : static_cast<uint32_t>(r) - ovfm[e.w + p]);
        uint32_t u[K];
        int32_t lw;  // the value to the left of my columns in the predecessor row
        if (d == 1) {
#pragma unroll
          for (int k = 0; k < K; ++k) u[k] = hp[k];
          lw = E_prev;
        } else if (use_ring && d <= static_cast<uint32_t>(kRingRows)) {
          const uint32_t slot = (static_cast<uint32_t>(r) - d) & (kRingRows - 1);
          row_load<K>(ring + slot * rw, lane, u);
          lw = ring_left[slot * 32];
        } else {
          const uint32_t prow = static_cast<uint32_t>(r) - d + lane;  // memory row of my block of that row
          row_load<K>(Hm + static_cast<uint64_t>(prow) * rw, lane, u);
          lw = static_cast<int32_t>(leftm[static_cast<uint64_t>(prow) * 32 + lane]);
        }
        lmax = lw > lmax ? lw : lmax;
        // diagonal of my first column: the left value; of my (K+1)-th column: the low half of my last register
        const uint32_t x0 = __byte_perm(static_cast<uint32_t>(lw), u[K - 1], 0x5410);
        if (p == 0) {
          h[0] = __viaddmax_s16x2(x0, pr[0], __vadd2(u[0], g2));
#pragma unroll
          for (int k = 1; k < K; ++k) h[k] = __viaddmax_s16x2(u[k - 1], pr[k], __vadd2(u[k], g2));
        } else {
          h[0] = __viaddmax_s16x2(u[0], g2, __viaddmax_s16x2(x0, pr[0], h[0]));
#pragma unroll
          for (int k = 1; k < K; ++k) h[k] = __viaddmax_s16x2(u[k], g2, __viaddmax_s16x2(u[k - 1], pr[k], h[k]));
        }
      }
    }
    if (act) {
      // ---- horizontal: running max inside each half, then the left value comes in
#pragma unroll
      for (int k = 1; k < K; ++k) h[k] = __viaddmax_s16x2(h[k - 1], g2, h[k]);
      if (lane == 0) E = SW ? 0 : lmax + g;  // first column (NW: max over predecessors + g, simd...:623-632)
      int32_t lef = lo16(h[K - 1]);          // last column of the low half, before / after the left value
      lef = E + gK > lef ? E + gK : lef;
      const uint32_t X = pack16(E, lef);
      if (SW) {
#pragma unroll
        for (int k = 0; k < K; ++k) hp[k] = __viaddmax_s16x2_relu(X, gk1[k], h[k]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) hp[k] = __viaddmax_s16x2(X, gk1[k], h[k]);
      }
      out_last = hi16(hp[K - 1]);
      E_prev = E;
      // ---- write the row once and keep it in the ring
      row_store<K>(Hm + static_cast<uint64_t>(t) * rw, lane, hp);  // memory row r + lane = t: contiguous across the warp
      leftm[static_cast<uint64_t>(t) * 32 + lane] = static_cast<int16_t>(E);
      if (use_ring) {
        const uint32_t slot = static_cast<uint32_t>(r) & (kRingRows - 1);
        row_store<K>(ring + slot * rw, lane, hp);
        ring_left[slot * 32] = E;
      }
      // ---- best cell
      if (SW) {
        uint32_t m = hp[0];
#pragma unroll
        for (int k = 1; k < K; ++k) m = __vmaxs2(m, hp[k]);
        const uint32_t nb = __vmaxs2(bestv, m);
        const uint32_t chg = nb ^ bestv;
        if (chg & 0xFFFFu) bestr_lo = r;
        if (chg >> 16) bestr_hi = r;
        bestv = nb;
      } else if ((meta & kMetaSink) && lane == lastL) {
        uint32_t sel = hp[0];
#pragma unroll
        for (int k = 1; k < K; ++k) {
          if (k == lastK) sel = hp[k];
        }
        const int32_t val = lastH ? hi16(sel) : lo16(sel);
        if (val > nw_best) {
          nw_best = val;
          nw_row = r;
        }
      }
    }
  }

  // ---- where the traceback starts
  if (!SW) {
    nw_best = __shfl_sync(FULL, nw_best, lastL);
    nw_row = __shfl_sync(FULL, nw_row, lastL);
    io.best_row = nw_row;
    io.best_col = nw_row ? len : 0;
    io.best_score = nw_best;
  } else {
    // global max, then the first row in rank order that reached it, then its first column
    int32_t mx = lo16(bestv) > hi16(bestv) ? lo16(bestv) : hi16(bestv);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const int32_t o = __shfl_xor_sync(FULL, mx, d);
      mx = o > mx ? o : mx;
    }
    uint32_t br = 0xFFFFFFFFu;
    if (mx > 0) {
      if (lo16(bestv) == mx) br = bestr_lo;
      if (hi16(bestv) == mx && bestr_hi < br) br = bestr_hi;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t o = __shfl_xor_sync(FULL, br, d);
      br = o < br ? o : br;
    }
    uint32_t brow = 0, col = 0;
    if (mx > 0) {
      brow = br;
      uint32_t u[K];
      row_load<K>(Hm + static_cast<uint64_t>(brow + lane) * rw, lane, u);  // own words
      uint32_t bc = 0xFFFFFFFFu;
#pragma unroll
      for (int k = K - 1; k >= 0; --k) {
        if (hi16(u[k]) == mx) bc = c0 + K + k;
      }
#pragma unroll
      for (int k = K - 1; k >= 0; --k) {
        if (lo16(u[k]) == mx) bc = c0 + k;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(FULL, bc, d);
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
__device__ __forceinline__ void wave_fill(FillIo& io, const uint8_t* codes, uint32_t len,
                                          uint32_t mode, const Scores sc, uint32_t num_codes, uint32_t* prof,
                                          U4* recs, uint32_t* ring, int ring_rows) {
  if (mode == kModeSW) wave_fill_t<K, true>(io, codes, len, sc, num_codes, prof, recs, ring, ring_rows);
  else wave_fill_t<K, false>(io, codes, len, sc, num_codes, prof, recs, ring, ring_rows);
}

}  // namespace vgc

#endif  // VGC_POA_FILL_CUH_

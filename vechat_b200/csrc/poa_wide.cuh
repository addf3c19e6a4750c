// poa_wide.cuh — the DP fill of the wide path (sm_100a): int32 cells, rows of any width, one warp per alignment.
//
// Takes the alignments the packed int16 fill (poa_fill.cuh) cannot: a layer longer than its widest row (racon's -w
// above ~1000) or scores that may leave the int16 range on this graph (deep windows with large penalties).  Same
// recurrence and the same reference code (vendor/spoa/src/simd_alignment_engine_implementation.hpp:760-906 with the
// int32 lanes chosen at :699-706; Initialize :506-681).  It is the capacity path, not the fast path: the matrix
// (row-major, H[row * cols + j], cols = len + 1) lives in L2 / HBM and every lane owns every 32nd column.  The
// traceback is wide_trace() in poa_core.h.
#ifndef VGC_POA_WIDE_CUH_
#define VGC_POA_WIDE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "poa_core.h"

namespace vgc {

struct WideFillIo {
  int32_t* H;
  uint32_t cols;
  const U4* rp;
  const uint32_t* ovf;
  uint32_t nR;
  uint32_t best_row, best_col;  // out: where the traceback starts (0,0 = empty alignment)
  int32_t best_score;
};

template <bool SW>
__device__ __forceinline__ void warp_fill_wide(WideFillIo& io, const uint8_t* codes, uint32_t len, const Scores sc) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  constexpr int32_t kMin = INT32_MIN / 2;
  const int lane = threadIdx.x & 31;
  const uint64_t cols = io.cols;
  const int32_t g = sc.g;
  int32_t* const H = io.H;
  // virtual row 0
  for (uint32_t j = lane; j <= len; j += 32) H[j] = SW ? 0 : g * static_cast<int32_t>(j);
  __syncwarp();
  int32_t bv = 0;                 // SW: this lane's best score (> 0), first cell in row-major order that reached it
  uint32_t brow = 0, bcol = 0;
  int32_t nw_best = INT32_MIN;    // NW: best last-column score among the sinks, first in rank order
  uint32_t nw_row = 0;
  for (uint32_t r = 0; r < io.nR; ++r) {
    const uint32_t row = r + 1;
    const U4 rec = io.rp[r];
    const uint32_t code = meta_code(rec.x);
    const uint32_t np = meta_npred(rec.x);
    const uint32_t npp = np == 0 ? 1u : np;
    int32_t* const out = H + row * cols;
    int32_t fc = INT32_MIN;
    for (uint32_t p = 0; p < npp; ++p) {
      const uint32_t pr = np == 0 ? 0u : rec_pred(rec, row, p, io.ovf);
      const int32_t v = __ldcg(H + pr * cols);
      fc = v > fc ? v : fc;
    }
    const int32_t fci = SW ? 0 : fc + g;
    if (lane == 0) out[0] = fci;
    int32_t carry = fci;  // H[row][c0 - 1]
    int32_t last = fci;
    for (uint32_t c0 = 1; c0 <= len; c0 += 32) {
      const uint32_t j = c0 + lane;
      const bool valid = j <= len;
      int32_t best = kMin;
      if (valid) {
        const int32_t s = codes[j - 1] == code ? sc.m : sc.x;
        for (uint32_t p = 0; p < npp; ++p) {
          const uint32_t pr = np == 0 ? 0u : rec_pred(rec, row, p, io.ovf);
          const int32_t* hp = H + pr * cols + j;
          const int32_t d = __ldcg(hp - 1) + s, v = __ldcg(hp) + g;
          best = d > best ? d : best;
          best = v > best ? v : best;
        }
      }
      // horizontal: H[j] = max(best_j, H[j - 1] + g) = g * j + max over k <= j of (best_k - g * k), and the carry
      int32_t t = valid ? best - g * static_cast<int32_t>(j) : kMin;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t o = __shfl_up_sync(FULL, t, d);
        if (lane >= d) t = o > t ? o : t;
      }
      const int32_t cin = carry - g * static_cast<int32_t>(c0 - 1);
      int32_t val = (cin > t ? cin : t) + g * static_cast<int32_t>(j);
      if (SW) val = val > 0 ? val : 0;
      if (valid) {
        out[j] = val;
        if (SW && val > bv) {
          bv = val;
          brow = row;
          bcol = j;
        }
      }
      carry = __shfl_sync(FULL, val, 31);
      const uint32_t lastlane = len - c0 < 31u ? len - c0 : 31u;
      last = __shfl_sync(FULL, val, lastlane);
    }
    if (!SW && (rec.x & kMetaSink) && last > nw_best) {
      nw_best = last;
      nw_row = row;
    }
    __syncwarp();  // the row is read by every lane from here on
  }
  if (!SW) {
    io.best_row = nw_row;
    io.best_col = nw_row ? len : 0;
    io.best_score = nw_best;
  } else {
    // global maximum, then its first row in rank order, then its first column in that row
    int32_t mx = bv;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const int32_t o = __shfl_xor_sync(FULL, mx, d);
      mx = o > mx ? o : mx;
    }
    uint32_t br = (mx > 0 && bv == mx) ? brow : 0xFFFFFFFFu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t o = __shfl_xor_sync(FULL, br, d);
      br = o < br ? o : br;
    }
    // a lane's record is its FIRST cell with the maximum (rows ascending, then its columns ascending); the first
    // column of row br among all lanes needs the row itself: other lanes may hold the maximum in br at a lower column
    // than the lane that saw it first... every lane checks its own columns of row br
    uint32_t bc = 0xFFFFFFFFu;
    if (mx > 0) {
      for (uint32_t j = 1 + lane; j <= len; j += 32) {
        if (__ldcg(H + br * cols + j) == mx) {
          bc = j;
          break;
        }
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const uint32_t o = __shfl_xor_sync(FULL, bc, d);
      bc = o < bc ? o : bc;
    }
    io.best_row = mx > 0 ? br : 0;
    io.best_col = mx > 0 ? bc : 0;
    io.best_score = mx;
  }
  __syncwarp();
}

}  // namespace vgc

#endif  // VGC_POA_WIDE_CUH_

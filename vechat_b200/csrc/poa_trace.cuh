// poa_trace.cuh — warp-cooperative traceback of one sequence-to-graph alignment (sm_100a).
//
// Replaces the traceback of SimdAlignmentEngine::Linear (vendor/spoa/src/simd_alignment_engine_implementation.hpp:
// 908-1105, scalar twin sisd_alignment_engine.cpp:362-460).  The warp that has just filled the matrix walks it back:
// the walk itself is a chain of dependent steps, but every step has up to 13 candidate cells (the diagonal and the
// vertical neighbour in each of <= 6 predecessor rows, and the horizontal neighbour), so
//   * lanes 0-5 test the diagonals, lanes 6-11 the verticals (in-edge order), lane 12 the horizontal move: one
//     shared-memory load and one compare each; a ballot collects the matches and its lowest set bit is the move
//     the reference takes (its priority order: diagonals over the predecessors in in-edge order, then verticals,
//     then horizontal — first match wins, :1031-1061);
//   * the cells come from a 32-row x 32-word tile of the matrix (plus the 32 row records) that all 32 lanes fetch
//     together with 16-byte cp.async copies whenever the walk leaves it: one L2/DRAM round trip per ~25 steps.
// Steps the tile cannot serve (borders, rows with more than six in-edges, a predecessor more than 31 rows up, the
// seam between the two column halves of the lane-major row layout) take a scalar path with direct loads.
//
// Two outputs:
//   * WEIGHTS == false: the alignment (node id | -1, sequence position | -1), reversed, for AddAlignment /
//     GenerateCorrectedSequence (graph.cpp:182-299, :1167-1179);
//   * WEIGHTS == true : Graph::AddWeights (graph.cpp:1104-1165) fused into the walk — every pair of consecutive
//     matched positions adds w[pos-1] + w[pos] to the edge between their nodes.  The edge is "in-edge slot p of the
//     node at rank r", so the sum goes to wacc[r * kInlinePreds + p] (atomic: the alignments of a round run
//     concurrently) and Poa::fold_weights() moves it to the edge afterwards.
#ifndef VGC_POA_TRACE_CUH_
#define VGC_POA_TRACE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "poa_core.h"

namespace vgc {

constexpr int kTileRows = 32;    // rows ti, ti-1, .. ti-31 (rank space)
constexpr int kTileWords = 32;   // words [wb, wb + 32) of the lane-major row
constexpr int kTilePitch = 36;   // words between tile rows in shared memory: 16-byte aligned, banks shifted by 4 per row
// shared memory of one walk: cells | row records
constexpr uint32_t kTraceTileBytes = kTileRows * kTilePitch * 4 + kTileRows * 16;

struct TraceIo {
  // the matrix (align kernel's scratch) and the row program it was filled from
  const uint32_t* H;
  const int16_t* fc;
  uint32_t rw;             // words between rows
  uint32_t half_words;     // words per column half (32 * K of the fill)
  const U4* rp;
  const uint32_t* ovf;
  const uint32_t* nodes;   // rank -> node id; nullptr: the id is in the row record (slots below 65536 nodes)
  const uint8_t* codes;    // shared memory: codes of the sequence
  int32_t m, x, g;
  bool sw;
  uint32_t row, col;       // start cell (row = rank + 1, DP column, 1-based); 0,0 = empty alignment
  uint32_t max_steps;
  // WEIGHTS == false
  int32_t* aln_node;
  int32_t* aln_pos;
  uint32_t aln_cap;
  // WEIGHTS == true
  uint32_t* wacc;          // [nR * kInlinePreds]
  const uint32_t* w2;      // shared memory: w2[pos] = weight(pos - 1) + weight(pos), pos >= 1
  uint32_t* ew;            // edge weights / in-lists of the graph: rows with more than kInlinePreds in-edges
  const uint32_t* ieid;
  uint32_t in_stride;
};

// returns kWalkDone or kWalkBad; *n_out = pairs written (WEIGHTS == false) or steps taken
template <bool WEIGHTS>
__device__ int warp_trace(const TraceIo& t, uint32_t* tile, uint32_t* n_out, uint32_t* refills_out) {
  const int lane = threadIdx.x & 31;
  const uint32_t FULL = 0xFFFFFFFFu;
  uint32_t* cells = tile;
  U4* recs = reinterpret_cast<U4*>(tile + kTileRows * kTilePitch);
  const uint32_t half = t.half_words;
  uint32_t i = t.row, j = t.col, n = 0, refills = 0;
  *n_out = 0;
  *refills_out = 0;
  if (i == 0 && j == 0) return kWalkDone;

  // H(row, jj) by direct load; jj = DP column (0 = first column).  Uniform across the warp.
  auto cell_g = [&](uint32_t row, uint32_t jj) -> int32_t {
    if (jj == 0) return t.sw ? 0 : static_cast<int32_t>(t.fc[row]);
    const uint32_t c = jj - 1;
    const uint32_t hi = c >= half ? 1u : 0u;
    const uint32_t v = t.H[static_cast<uint64_t>(row) * t.rw + (hi ? c - half : c)];
    return static_cast<int16_t>(hi ? (v >> 16) : (v & 0xFFFFu));
  };

  uint32_t ti = 0, wb = 0;
  bool have = false;
  int32_t h = cell_g(i, j);
  U4 rec = i ? t.rp[i - 1] : U4{0, 0, 0, 0};
  uint32_t pend = kNone, pend_w = 0;  // WEIGHTS: the matched pair emitted last, waiting to learn whether the next is one too

  const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  auto refill = [&]() {
    __syncwarp();
    ti = i;
    const uint32_t c1 = j - 1;
    const uint32_t w1 = c1 >= half ? c1 - half : c1;
    uint32_t b = (w1 & ~3u) + 4u;
    b = b >= static_cast<uint32_t>(kTileWords) ? b - kTileWords : 0u;
    if (b + kTileWords > half) b = half - kTileWords;
    wb = b;
    if (static_cast<uint32_t>(lane) <= ti) {
      const uint32_t row = ti - lane;
      const uint32_t* src = t.H + static_cast<uint64_t>(row) * t.rw + wb;
      const uint32_t dst = tile_s + lane * (kTilePitch * 4);
#pragma unroll
      for (int q = 0; q < kTileWords / 4; ++q)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * 16), "l"(src + q * 4) : "memory");
      if (row >= 1) {
        const uint32_t rdst = tile_s + kTileRows * kTilePitch * 4 + lane * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rdst), "l"(t.rp + (row - 1)) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    have = true;
    ++refills;
  };

  while (true) {
    if (t.sw ? (h == 0) : (i == 0 && j == 0)) break;
    if (n >= t.max_steps) return kWalkBad;
    const uint32_t np = i != 0 ? meta_npred(rec.x) : 0u;
    const uint32_t code = meta_code(rec.x);
    // kind: 0 diagonal, 1 vertical, 2 horizontal; psel = predecessor (in-edge slot) of a diagonal / vertical move
    uint32_t kind = 3, psel = 0, pi = i;
    int32_t hn = 0;
    bool stepped = false;
    if (i != 0 && j >= 2 && (rec.x & kMetaInline)) {
      const uint32_t npp = np ? np : 1u;
      const int32_t mc = code == t.codes[j - 1] ? t.m : t.x;
      const bool is_d = lane < 6, is_v = lane >= 6 && lane < 12, is_h = lane == 12;
      const uint32_t p = is_d ? lane : (is_v ? lane - 6 : 0u);
      const bool need = is_h || ((is_d || is_v) && p < npp);
      const uint32_t d = is_h ? 0u : rec_delta(rec, p);
      const uint32_t row = i - d;
      const uint32_t cc = is_v ? j - 1 : j - 2;
      const uint32_t hi = cc >= half ? 1u : 0u;
      const uint32_t w = hi ? cc - half : cc;
      const uint32_t target = static_cast<uint32_t>(is_d ? h - mc : h - t.g) & 0xFFFFu;
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        const uint32_t dr = ti - row, dq = w - wb;
        const bool ok = have && dr < static_cast<uint32_t>(kTileRows) && dq < static_cast<uint32_t>(kTileWords);
        if (!__any_sync(FULL, need && !ok)) {
          const uint32_t v = ok ? cells[dr * kTilePitch + dq] : 0u;
          const uint32_t val = hi ? (v >> 16) : (v & 0xFFFFu);
          const uint32_t mask = __ballot_sync(FULL, need && val == target);
          if (mask == 0) return kWalkBad;
          const uint32_t sel = static_cast<uint32_t>(__ffs(static_cast<int>(mask))) - 1u;
          kind = sel == 12 ? 2u : (sel >= 6 ? 1u : 0u);
          psel = kind == 2 ? 0u : (kind == 1 ? sel - 6 : sel);
          pi = kind == 2 ? i : i - rec_delta(rec, psel);
          hn = kind == 0 ? h - mc : h - t.g;
          stepped = true;
          break;
        }
        if (attempt == 0) refill();
      }
    }
    if (!stepped) {
      // ---- scalar path (every lane computes the same thing from direct loads)
      const uint32_t npp = i != 0 ? (np == 0 ? 1u : np) : 0u;
      int32_t mc = 0;
      if (i != 0 && j != 0) mc = code == t.codes[j - 1] ? t.m : t.x;
      bool found = false;
      if (j != 0) {
        for (uint32_t p = 0; p < npp; ++p) {
          const uint32_t pr = np == 0 ? 0u : rec_pred(rec, i, p, t.ovf);
          const int32_t hv = cell_g(pr, j - 1);
          if (h == hv + mc) {
            found = true;
            kind = 0;
            psel = p;
            pi = pr;
            hn = hv;
            break;
          }
        }
      }
      if (!found) {
        for (uint32_t p = 0; p < npp; ++p) {
          const uint32_t pr = np == 0 ? 0u : rec_pred(rec, i, p, t.ovf);
          const int32_t hv = cell_g(pr, j);
          if (h == hv + t.g) {
            found = true;
            kind = 1;
            psel = p;
            pi = pr;
            hn = hv;
            break;
          }
        }
      }
      if (!found && j != 0) {
        const int32_t hv = cell_g(i, j - 1);
        if (h == hv + t.g) {
          found = true;
          kind = 2;
          pi = i;
          hn = hv;
        }
      }
      if (!found) return kWalkBad;
    }
    // ---- the move is known: emit, advance
    if (WEIGHTS) {
      if (kind == 0) {
        if (pend != kNone && lane == 0) {
          if (pend & 0x80000000u) atomicAdd(t.ew + (pend & 0x7FFFFFFFu), pend_w);
          else atomicAdd(t.wacc + pend, pend_w);
        }
        if (np == 0) {
          pend = kNone;  // the predecessor is the virtual row: the next entry cannot be a matched pair
        } else if (psel < kInlinePreds && (rec.x & kMetaInline)) {
          pend = (i - 1) * kInlinePreds + psel;
        } else {
          const uint32_t nd = t.nodes ? t.nodes[i - 1] : meta_node(rec.x);
          pend = 0x80000000u | t.ieid[static_cast<uint64_t>(nd) * t.in_stride + psel];
        }
        pend_w = t.w2[j - 1];
      } else {
        pend = kNone;
      }
    } else {
      if (n >= t.aln_cap) return kWalkBad;
      if (lane == 0) {
        t.aln_node[n] = kind == 2 ? -1 : static_cast<int32_t>(t.nodes ? t.nodes[i - 1] : meta_node(rec.x));
        t.aln_pos[n] = kind == 1 ? -1 : static_cast<int32_t>(j - 1);
      }
    }
    ++n;
    if (pi != i) {
      if (pi == 0) rec = U4{0, 0, 0, 0};
      else if (have && pi <= ti && ti - pi < static_cast<uint32_t>(kTileRows)) rec = recs[ti - pi];
      else rec = t.rp[pi - 1];
      i = pi;
    }
    if (kind != 1) j = j - 1;
    h = hn;
  }
  *n_out = n;
  *refills_out = refills;
  return kWalkDone;
}

}  // namespace vgc

#endif  // VGC_POA_TRACE_CUH_

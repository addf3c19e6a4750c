// poa_trace.cuh — warp-cooperative traceback of one sequence-to-graph alignment (sm_100a).
//
// Replaces the traceback of SimdAlignmentEngine::Linear (vendor/spoa/src/simd_alignment_engine_implementation.hpp:
// 908-1105, scalar twin sisd_alignment_engine.cpp:362-460).  The warp that has just filled the matrix walks it back.
// The walk is a chain of dependent moves, but which move a cell takes depends only on that cell and its neighbours
// (the reference compares scores: diagonals over the predecessors in in-edge order, then verticals, then the
// horizontal neighbour — first match wins, :1031-1061), not on how the walk got there.  So:
//   * RUN step: lane k evaluates the move of cell (i - k, j - k) — the k-th cell of the diagonal through the current
//     cell — as if rows i, i-1, .. were a chain (one predecessor, one row up).  A ballot finds the longest prefix of
//     lanes whose rows really are chain rows and whose move is the diagonal one: all of those moves, and the first
//     move that leaves the diagonal, are taken at once (matches and mismatches are both diagonal moves, so runs are
//     ~1/indel-rate long on chains).
//   * GENERAL step (the current row has several predecessors): lanes 0-5 test the diagonals, lanes 6-11 the
//     verticals (in-edge order), lane 12 the horizontal move; the lowest set bit of the ballot is the move.
//   * the cells come from a 32-row x 32-word tile of the matrix (plus the 32 row records) that all 32 lanes fetch
//     together with cp.async copies whenever the walk leaves it: one L2/DRAM round trip per ~25 moves.  In the fill's
//     row layout (poa_fill.cuh: 2K consecutive columns = K consecutive words, block l of row r in memory row r + l)
//     32 words are 64 consecutive columns.
//   * moves the tile cannot serve (borders, rows with more than six in-edges, a predecessor more than 31 rows up)
//     take a scalar path with direct loads.
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
  const uint32_t* H;       // rows of 32*K words: word w holds column w (low half) and column 32*K + w (high half)
  const int16_t* fc;       // first-column value of every row (NW border)
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

// cell column (0-based) -> word of the row and half of the word, for rows of K words per lane
template <int K>
__device__ __forceinline__ uint32_t col_word(uint32_t c, uint32_t* hi) {
  const uint32_t h = c >= 32u * K ? 1u : 0u;
  *hi = h;
  return c - h * (32u * K);
}

// returns kWalkDone or kWalkBad; *n_out = pairs written (WEIGHTS == false) or moves taken
template <int K, bool WEIGHTS>
__device__ __forceinline__ int warp_trace(const TraceIo& t, uint32_t* tile, uint32_t* n_out, uint32_t* refills_out, uint32_t* slow_out) {
  constexpr uint32_t rw = 32u * K;
  constexpr uint32_t TR = kTileRows, TW = kTileWords;
  const int lane = threadIdx.x & 31;
  const uint32_t FULL = 0xFFFFFFFFu;
  uint32_t* cells = tile;
  U4* recs = reinterpret_cast<U4*>(tile + kTileRows * kTilePitch);
  uint32_t i = t.row, j = t.col, n = 0, refills = 0, slow = 0;
  *n_out = 0;
  *refills_out = 0;
  *slow_out = 0;
  if (i == 0 && j == 0) return kWalkDone;
  const int32_t m = t.m, x = t.x, g = t.g;
  const bool sw = t.sw;

  // H(row, jj) by direct load; jj = DP column (0 = first column).  Uniform across the warp.
  auto cell_g = [&](uint32_t row, uint32_t jj) -> int32_t {
    if (jj == 0) return sw ? 0 : static_cast<int32_t>(t.fc[row]);
    uint32_t hi;
    const uint32_t w = col_word<K>(jj - 1, &hi);
    const uint32_t v = t.H[static_cast<uint64_t>(row) * rw + w];
    return static_cast<int16_t>(hi ? (v >> 16) : (v & 0xFFFFu));
  };
  auto node_of = [&](uint32_t row, uint32_t meta) -> int32_t {
    return static_cast<int32_t>(t.nodes ? t.nodes[row - 1] : meta_node(meta));
  };

  uint32_t ti = 0, wb = 0;
  bool have = false;   // the tile holds rows [ti - 31, ti] x words [wb, wb + 32)
  bool fresh = false;  // ... and was fetched for the current cell
  uint32_t pend = kNone, pend_w = 0;  // WEIGHTS: the matched pair emitted last, waiting to learn whether the next is one too
  auto commit_pend = [&]() {
    if (pend != kNone && lane == 0) {
      if (pend & 0x80000000u) atomicAdd(t.ew + (pend & 0x7FFFFFFFu), pend_w);
      else atomicAdd(t.wacc + pend, pend_w);
    }
  };

  const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  auto refill = [&]() {
    __syncwarp();
    ti = i;
    // the tile ends just after the word of the current column and extends to the left (lower columns = lower words)
    uint32_t hi1;
    const uint32_t endw = (col_word<K>(j - 1, &hi1) & ~3u) + 4u;
    wb = endw > TW ? endw - TW : 0u;
    if (static_cast<uint32_t>(lane) <= ti) {
      const uint32_t row = ti - lane;
      const uint32_t dst = tile_s + lane * (kTilePitch * 4);
      const uint32_t* src = t.H + static_cast<uint64_t>(row) * rw + wb;
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
    fresh = true;
    ++refills;
  };
  // raw 16-bit cell from the tile (caller guarantees it is inside)
  auto tile_cell = [&](uint32_t row, uint32_t w, uint32_t hi) -> uint32_t {
    const uint32_t v = cells[(ti - row) * kTilePitch + (w - wb)];
    return hi ? (v >> 16) : (v & 0xFFFFu);
  };
  auto done = [&]() {
    *n_out = n;
    *refills_out = refills;
    *slow_out = slow;
    return kWalkDone;
  };

  while (true) {
    if (!sw && i == 0 && j == 0) break;
    if (n >= t.max_steps) return kWalkBad;
    // =============================== RUN step ===============================================================
    // lane k: cell (i - k, j - k), taken as a chain row.  Needs: row >= 1, column >= 2, an inline record with exactly
    // one predecessor, and the four cells (own, diagonal, vertical, horizontal) inside the tile.
    bool took = false;
    // the record of the current row decides: a chain row (one predecessor) starts a RUN step, anything else goes
    // straight to the GENERAL step (a RUN would consume nothing)
    const U4 rec = i == 0 ? U4{0, 0, 0, 0}
                          : ((have && i <= ti && ti - i < TR) ? recs[ti - i] : t.rp[i - 1]);
    if (i >= 1 && j >= 2 && (rec.x & kMetaInline) && meta_npred(rec.x) == 1) {
#pragma unroll 1
      for (int attempt = 0; attempt < 2 && !took; ++attempt) {
        const uint32_t rk = i - lane, jk = j - lane;  // wrap for lanes beyond the border: excluded below
        const bool in_rows = static_cast<uint32_t>(lane) < i && static_cast<uint32_t>(lane) + 2 <= j && have &&
                             rk <= ti && ti - rk < TR;
        U4 lrec = {0, 0, 0, 0};
        if (in_rows) lrec = recs[ti - rk];
        const uint32_t d0 = lrec.y & 0xFFFFu;
        const uint32_t prow = rk - d0;
        uint32_t hi1, hi0;
        const uint32_t w1 = col_word<K>(jk - 1, &hi1), w0 = col_word<K>(jk - 2, &hi0);
        const bool in_cols = w1 - wb < TW && w0 - wb < TW;
        const bool simple = in_rows && in_cols && (lrec.x & kMetaInline) && meta_npred(lrec.x) == 1 && d0 <= rk &&
                            ti - prow < TR;
        uint32_t mv = 3;  // 0 diagonal, 1 vertical, 2 horizontal, 3 none / not evaluated
        bool stop = false;
        if (simple) {
          const uint32_t own = tile_cell(rk, w1, hi1);
          const uint32_t cd = tile_cell(prow, w0, hi0), cv = tile_cell(prow, w1, hi1), ch = tile_cell(rk, w0, hi0);
          const uint32_t mc = static_cast<uint32_t>(meta_code(lrec.x) == t.codes[jk - 1] ? m : x);
          const uint32_t gu = static_cast<uint32_t>(g);
          stop = sw && own == 0;
          mv = ((cd + mc) & 0xFFFFu) == own ? 0u : (((cv + gu) & 0xFFFFu) == own ? 1u : (((ch + gu) & 0xFFFFu) == own ? 2u : 3u));
          if (stop) mv = 3;
        }
        const bool cont = simple && mv == 0 && d0 == 1;
        const uint32_t cmask = __ballot_sync(FULL, cont);
        // leading lanes that continue along the diagonal
        const uint32_t P = cmask == FULL ? 32u : static_cast<uint32_t>(__ffs(static_cast<int>(~cmask))) - 1u;
        // the lane after the run takes its own move too if it could evaluate it
        const uint32_t endmask = __ballot_sync(FULL, simple && mv != 3);
        const uint32_t stopmask = __ballot_sync(FULL, stop);
        const bool end_ok = P < 32 && ((endmask >> P) & 1u);
        const uint32_t consumed = P + (end_ok ? 1u : 0u);
        if (consumed == 0) {
          if (stopmask & 1u) return done();  // SW: the current cell holds 0 — the alignment starts here
          // lane 0 could not evaluate its cell: outside the tile (fetch it, once), or not a chain row
          const bool l0_tile = __shfl_sync(FULL, static_cast<int>(in_rows && in_cols), 0) != 0;
          if (!l0_tile && !fresh && attempt == 0) {
            refill();
            continue;
          }
          break;
        }
        // ---- emit the moves of lanes [0, consumed)
        const uint32_t dmask = __ballot_sync(FULL, mv == 0);
        if (!WEIGHTS && n + consumed > t.aln_cap) return kWalkBad;
        if (static_cast<uint32_t>(lane) < consumed) {
          if (WEIGHTS) {
            // consecutive matched pairs: my move and the next emitted one (lane + 1, if consumed) are both diagonal
            if (mv == 0 && static_cast<uint32_t>(lane) + 1 < consumed && ((dmask >> (lane + 1)) & 1u))
              atomicAdd(t.wacc + (rk - 1) * kInlinePreds, t.w2[jk - 1]);
          } else {
            t.aln_node[n + lane] = mv == 2 ? -1 : node_of(rk, lrec.x);
            t.aln_pos[n + lane] = mv == 1 ? -1 : static_cast<int32_t>(jk - 1);
          }
        }
        if (WEIGHTS) {
          if (dmask & 1u) commit_pend();  // the pair emitted before this run is followed by a matched pair
          // the last consumed move becomes the pending pair (if diagonal)
          const uint32_t last = consumed - 1;
          if ((dmask >> last) & 1u) {
            pend = (i - last - 1) * kInlinePreds;
            pend_w = t.w2[j - last - 1];
          } else {
            pend = kNone;
          }
        }
        n += consumed;
        // new position: after P diagonal moves, then the move of lane P (if taken)
        {
          const uint32_t src = end_ok ? P : 0u;
          const uint32_t emv = __shfl_sync(FULL, mv, src), ed0 = __shfl_sync(FULL, d0, src);
          uint32_t ni = i - P, nj = j - P;
          if (end_ok) {
            if (emv != 2) ni -= ed0;
            if (emv != 1) nj -= 1;
          }
          i = ni;
          j = nj;
        }
        fresh = false;
        took = true;
      }
    }
    if (took) continue;
    // =============================== GENERAL step ===========================================================
    const uint32_t np = i != 0 ? meta_npred(rec.x) : 0u;
    const uint32_t code = meta_code(rec.x);
    // kind: 0 diagonal, 1 vertical, 2 horizontal; psel = predecessor (in-edge slot) of a diagonal / vertical move
    uint32_t kind = 3, psel = 0, pi = i;
    bool stepped = false;
    if (i != 0 && j >= 2 && (rec.x & kMetaInline)) {
      const uint32_t npp = np ? np : 1u;
      const uint32_t mc = static_cast<uint32_t>(code == t.codes[j - 1] ? m : x);
      const bool is_d = lane < 6, is_v = lane >= 6 && lane < 12, is_h = lane == 12;
      const uint32_t p = is_d ? lane : (is_v ? lane - 6 : 0u);
      const bool need = is_h || ((is_d || is_v) && p < npp);
      const uint32_t d = is_h ? 0u : rec_delta(rec, p);
      const uint32_t row = i - d;
      uint32_t hi, hio;
      const uint32_t w = col_word<K>(is_v ? j - 1 : j - 2, &hi);
      const uint32_t wo = col_word<K>(j - 1, &hio);
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        const bool ok = have && d <= i && row <= ti && ti - row < TR && w - wb < TW && wo - wb < TW && i <= ti &&
                        ti - i < TR;
        if (!__any_sync(FULL, need && !ok)) {
          const uint32_t own = tile_cell(i, wo, hio);
          if (sw && own == 0) return done();
          const uint32_t val = ok ? tile_cell(row, w, hi) : 0u;
          const uint32_t mask =
              __ballot_sync(FULL, need && ((val + (is_d ? mc : static_cast<uint32_t>(g))) & 0xFFFFu) == own);
          if (mask == 0) return kWalkBad;
          const uint32_t sel = static_cast<uint32_t>(__ffs(static_cast<int>(mask))) - 1u;
          kind = sel == 12 ? 2u : (sel >= 6 ? 1u : 0u);
          psel = kind == 2 ? 0u : (kind == 1 ? sel - 6 : sel);
          pi = kind == 2 ? i : i - rec_delta(rec, psel);
          stepped = true;
          break;
        }
        if (attempt == 0 && !fresh) refill();
        else break;
      }
    }
    if (!stepped) {
      ++slow;
      // ---- scalar path (every lane computes the same thing from direct loads)
      const int32_t h = cell_g(i, j);
      if (sw && h == 0) break;
      const uint32_t npp = i != 0 ? (np == 0 ? 1u : np) : 0u;
      int32_t mc = 0;
      if (i != 0 && j != 0) mc = code == t.codes[j - 1] ? m : x;
      bool found = false;
      if (j != 0) {
        for (uint32_t p = 0; p < npp; ++p) {
          const uint32_t pr = np == 0 ? 0u : rec_pred(rec, i, p, t.ovf);
          if (h == cell_g(pr, j - 1) + mc) {
            found = true;
            kind = 0;
            psel = p;
            pi = pr;
            break;
          }
        }
      }
      if (!found) {
        for (uint32_t p = 0; p < npp; ++p) {
          const uint32_t pr = np == 0 ? 0u : rec_pred(rec, i, p, t.ovf);
          if (h == cell_g(pr, j) + g) {
            found = true;
            kind = 1;
            psel = p;
            pi = pr;
            break;
          }
        }
      }
      if (!found && j != 0) {
        if (h == cell_g(i, j - 1) + g) {
          found = true;
          kind = 2;
          pi = i;
        }
      }
      if (!found) return kWalkBad;
    }
    // ---- the move is known: emit, advance
    if (WEIGHTS) {
      if (kind == 0) {
        commit_pend();
        if (np == 0) {
          pend = kNone;  // the predecessor is the virtual row: the next entry cannot be a matched pair
        } else if (psel < kInlinePreds && (rec.x & kMetaInline)) {
          pend = (i - 1) * kInlinePreds + psel;
        } else {
          const uint32_t nd = static_cast<uint32_t>(node_of(i, rec.x));
          pend = 0x80000000u | t.ieid[static_cast<uint64_t>(nd) * t.in_stride + psel];
        }
        pend_w = t.w2[j - 1];
      } else {
        pend = kNone;
      }
    } else {
      if (n >= t.aln_cap) return kWalkBad;
      if (lane == 0) {
        t.aln_node[n] = kind == 2 ? -1 : node_of(i, rec.x);
        t.aln_pos[n] = kind == 1 ? -1 : static_cast<int32_t>(j - 1);
      }
    }
    ++n;
    i = pi;
    if (kind != 1) j = j - 1;
    fresh = false;
  }
  return done();
}

}  // namespace vgc

#endif  // VGC_POA_TRACE_CUH_

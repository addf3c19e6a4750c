// poa_core.h — the window-correction algorithm on flat arrays, written once.
//
// The same template code is instantiated
//   * in the CUDA engine (vgc_engine.cu) with a warp executor (one window per warp, 32 lanes, warp shuffles,
//     shared memory) for the update / sort steps, the packed int16x2 DP fill of poa_fill.cuh, and the
//     one-thread-per-window TraceWalker below; and
//   * in tests/host_model (g++, one "lane") so that the serial graph logic can be checked against the
//     oracle on a machine without a GPU.  The host build is test-only: the product library never
//     contains a CPU path.
//
// What it computes, with reference citations (paths under the reference tree):
//   Window::generate_consensus haplotype  src/window.cpp:176-428      -> run_window()
//   Window::generate_consensus linear     src/window.cpp:74-174       -> run_window() (haplotype == 0)
//   Graph::AddAlignment / AddSequence     vendor/spoa/src/graph.cpp:109-130,182-299 -> add_alignment()
//   Graph::TopologicalSort                graph.cpp:301-371           -> toposort_fast() (staged) / toposort_impl()
//   Graph::Subgraph / ExtractSubgraph     graph.cpp:640-732           -> extract_fast() / extract_impl() + member-only sort
//   Graph::PruneGraph                     graph.cpp:811-982           -> prune()
//   Graph::LargestSubgraph / DfsUtil      graph.cpp:984-1089          -> largest_subgraph()
//   Graph::AddWeights                     graph.cpp:1104-1165         -> add_weights()
//   Graph::GenerateCorrectedSequence      graph.cpp:1167-1179         -> step_update(), kPcFinalPost
//   Graph::GenerateConsensus (+coverage)  graph.cpp:450-485,534-638   -> heaviest_bundle()
//   Linear-gap NW/SW traceback            simd_alignment_engine_implementation.hpp:908-1105
//                                         (scalar twin sisd_alignment_engine.cpp:362-460) -> TraceWalker
//
// Data layout (per window "slot", all in HBM unless noted):
//   nodes : code u8, nin/nout u32, aligned list (<= kMaxAligned ids), coverage u32
//   edges : tail, head, weight (u32), in_ord / out_ord (position inside head's in-list / tail's
//           out-list; lists are "edges of that node in creation order", so CSR position =
//           off[node] + ord), dead u8 (pruned hole)
//   in-lists: per node a fixed-stride row of (tail, edge id) in creation order — appended in place, never
//           rebuilt; out_off/out_eid: CSR of out-edges, built only where a pass needs it
//   rowprog: 16 B per DP row in rank order: code, sink, node id, up to six predecessors as 16-bit row distances
//   H     : one row of packed int16 score cells per DP row (row = rank + 1; row 0 = the virtual row), lane-major
//   fc    : int16 first-column value per row (NW border)
#ifndef VGC_POA_CORE_H_
#define VGC_POA_CORE_H_

#include <stdint.h>

#ifdef __CUDACC__
#define VGC_HD __host__ __device__
#define VGC_INL __forceinline__
#else
#define VGC_HD
#define VGC_INL inline
#endif

#ifdef VGC_CHECK_ORDER
#include <algorithm>
#include <vector>
extern unsigned long long g_order_checks[4];
#endif

namespace vgc {

constexpr int kMaxAligned = 15;    // clique size - 1  (<= kMaxCodes - 1)
// aligned-list stride per node = Slot::al_stride: 8 ids (32 B) while the batch has at most 8 distinct bytes, else 16
struct alignas(16) U4 {
  uint32_t x, y, z, w;
};
constexpr int kMaxCodes = 16;      // distinct bytes per batch (A C G T N + the IUPAC ambiguity codes fit)
constexpr uint32_t kNone = 0xFFFFFFFFu;

// window status codes (device -> host)
enum : uint32_t {
  kStOk = 0,
  kStNodeOverflow = 1,   // arena too small: rerun with a larger slot
  kStEdgeOverflow = 2,
  kStOutOverflow = 3,
  kStScoreRange = 4,     // int16 score range exceeded
  kStTooLong = 5,        // layer longer than the row capacity
  kStAlignedOverflow = 6,
  kStInternal = 7,
  kStDegreeOverflow = 8, // a node's in-degree outgrew the slot's in-list stride: rerun with a larger one
  kStWideCapacity = 9,   // the int32 matrix of a wide alignment does not fit the wide pool's buffers
};

enum : uint32_t { kModeNW = 0, kModeSW = 1 };
enum : uint32_t { kPcInit = 0, kPcBuildPost = 1, kPcRoundPost = 2, kPcFinalPost = 3, kPcDone = 4, kPcLinearFinal = 5 };
// step a window waits for (WinState::need); a zeroed WinState starts at kPcInit / kNeedUpdate
enum : uint32_t { kNeedUpdate = 0, kNeedPrepare = 1, kNeedFill = 2, kNeedTrace = 3, kNeedNone = 4 };
// what step_prepare must do (WinState::prep)
enum : uint32_t { kPrepMainSort = 1, kPrepSubSort = 2, kPrepRowprog = 4, kPrepFill = 8, kPrepLargest = 16 };

// flags byte per node used by the sorts
enum : uint8_t {
  kFMarkMask = 3,
  kFIgnored = 4,
  kFHasAligned = 8,
  kFMember = 16,
};

struct Scores {
  int32_t m, x, g;
};

// Read-only batch, device-resident copy of vgc_batch + host-prepared per-window metadata.
struct BatchView {
  const uint8_t* bases;        // base_bits == 8: the caller's bytes; 2 / 4: codes packed by the engine's staging
  uint32_t base_bits = 8;      // (vgc_engine.cu stage_batch: 2 bits while the batch holds A C G T only, else 4)
  const uint8_t* quals;
  const uint64_t* seq_off;
  const uint8_t* has_qual;
  const uint32_t* begin;
  const uint32_t* end;
  const uint32_t* win_first;   // [n_windows + 1]
  const uint8_t* win_flags;
  const uint32_t* layer_rank;  // [n_layers]: for window w, entries win_first[w].. hold the GLOBAL layer
                               // ids in the order std::sort left them (window.cpp:203-210); add_layer-
                               // ignored layers are removed and the count is in win_nseq
  const uint32_t* win_nseq;    // [n_windows] sequences_.size()
  const double* win_avgw;      // [n_windows] average_weight (window.cpp:301-309), host fp64
  const uint64_t* out_off;     // [n_windows] offset of the window's output bytes
  const uint32_t* out_cap;     // [n_windows]
  const uint8_t* coder;        // [256] byte -> code (batch-wide alphabet)
  const uint8_t* decoder;      // [kMaxCodes]
  const uint32_t* wlut;        // [256] quality byte -> weight
  uint32_t num_codes;
};

// code of base `idx` of the batch (global base offset)
VGC_HD VGC_INL uint32_t base_code(const BatchView& bv, uint64_t idx) {
  if (bv.base_bits == 2) return (bv.bases[idx >> 2] >> ((idx & 3u) * 2u)) & 3u;
  if (bv.base_bits == 4) return (bv.bases[idx >> 1] >> ((idx & 1u) * 4u)) & 15u;
  return bv.coder[bv.bases[idx]];
}

struct Graph {
  uint32_t nV, nE;
  uint8_t* code;
  uint8_t* nal;
  uint32_t* al;        // [max_nodes * al_stride]
  uint32_t* nin;
  uint32_t* nout;
  uint32_t* cov;       // sequences through the node (linear mode coverage)
  uint32_t* etail;
  uint32_t* ehead;
  uint32_t* ew;
  uint32_t* ein_ord;
  uint32_t* eout_ord;
  uint8_t* edead;
  uint32_t* itail;     // [max_nodes * in_stride] in-edge tails of each node, creation order
  uint32_t* ieid;      // [max_nodes * in_stride] ... and their edge ids
};

// One slot of scratch in HBM.  Two graph buffers: LargestSubgraph writes the other one.
struct Slot {
  uint32_t max_nodes, max_edges, max_len, row_words;
  uint32_t in_stride;  // capacity of a node's in-list
  uint32_t al_stride;  // capacity of a node's aligned list (8 or 16)
  Graph g[2];
  uint32_t* out_off;   // [max_nodes + 1]
  uint32_t* out_eid;   // [max_edges]
  uint32_t* r2n;       // [max_nodes] rank -> node of the whole current graph
  uint32_t* order;     // [max_nodes] rank -> node of the rows the next alignment visits
  uint32_t* rank_of;   // [max_nodes] node -> rank in `order`
  uint32_t* tmp0;      // [max_nodes] scratch
  uint32_t* tmp1;      // [max_nodes] scratch
  uint8_t* flags;      // [max_nodes] (global fallback of the shared-memory flags)
  uint32_t* rowprog;   // [max_nodes * 4] by rank: {meta, p0, p1, p2 | ovf offset (npred >= 4)}; predecessors as rows
  uint32_t* ovf;       // [max_edges] predecessor rows of nodes with in-degree > 2
  int16_t* fc;         // [max_nodes + 1]
  uint32_t* H;         // [(max_nodes + 1) * row_words]; also the big scratch of the serial graph passes
  int32_t* aln_node;   // [max_len + max_nodes + 2]
  int32_t* aln_pos;
  uint32_t aln_cap;
  uint32_t* wacc;      // [max_nodes * kInlinePreds] re-alignment round: weight added to in-edge p of the node at rank r
                       // (index r * kInlinePreds + p) by the round's concurrent alignments; folded into ew by fold_weights()
  uint32_t h_words;    // words behind H (the device keeps only the graph passes' scratch there; DP rows live in the
                       // align kernel's per-SM pool)
  // incremental topological order (Poa::order_update): the order is a sequence of blocks, one per DFS root
  uint32_t* owner;     // [max_nodes] node -> root of the DFS that emits it (smallest id it can reach forwards)
  uint32_t* bsize;     // [max_nodes] root -> nodes in its block (0: not a root)
  uint32_t* bstart;    // [max_nodes] root -> rank of its block's first node
  uint32_t* bstart2;   // [max_nodes] the other buffer (bstart / bstart2 and r2n / r2n2 swap on every update)
  uint32_t* r2n2;      // [max_nodes]
  uint32_t* anch;      // [max_len + 2] per sequence position: owner of its (old or aligned-to) node, kNone: new unaligned
  uint8_t* dirty;      // [max_nodes] root -> its block must be re-sorted (all zero between updates)
};

// Row program record (16 B per DP row, rank order):
//   x = meta: code (bits 0-3) | sink (4) | inline (5) | number of predecessors (6-14) | chain (15: the only
//       predecessor is the row just above — the fill's fast path) | node id (16-31; meaningful while the slot holds
//       fewer than 65536 nodes — the traceback reads it instead of a table in HBM)
//   inline (<= 6 predecessors, every one within 65535 rows): y, z, w = six 16-bit row DISTANCES d_p (predecessor
//       p is row - d_p), in-edge order; a node without in-edges has the virtual row 0 as its only predecessor
//       (np = 0, d_0 = its own row)
//   otherwise: w = offset into ovf[] holding the np predecessor rows
constexpr uint32_t kInlinePreds = 6;
constexpr uint32_t kMetaSink = 1u << 4;
constexpr uint32_t kMetaInline = 1u << 5;
constexpr uint32_t kMetaChain = 1u << 15;
constexpr uint32_t kMetaMaxPred = 511;
VGC_HD VGC_INL uint32_t meta_pack(uint32_t code, uint32_t npred, bool sink, bool inl, uint32_t node, bool chain) {
  return code | (sink ? kMetaSink : 0u) | (inl ? kMetaInline : 0u) | (npred << 6) | (chain ? kMetaChain : 0u) | (node << 16);
}
VGC_HD VGC_INL uint32_t meta_code(uint32_t m) { return m & 0xFu; }
VGC_HD VGC_INL uint32_t meta_npred(uint32_t m) { return (m >> 6) & 0x1FFu; }
VGC_HD VGC_INL uint32_t meta_node(uint32_t m) { return m >> 16; }
// predecessor p of the row `row` whose record is `e`
VGC_HD VGC_INL uint32_t rec_delta(const U4& e, uint32_t p) {
  const uint32_t wd = p < 2 ? e.y : (p < 4 ? e.z : e.w);
  return (p & 1u) ? (wd >> 16) : (wd & 0xFFFFu);
}
VGC_HD VGC_INL uint32_t rec_pred(const U4& e, uint32_t row, uint32_t p, const uint32_t* ovf) {
  if (e.x & kMetaInline) return row - rec_delta(e, p);
  return ovf[e.w + p];
}

// Shared per-window state (shared memory on the device).
struct WinState {
  uint32_t cur;          // which graph buffer is live
  uint32_t nR;           // rows of the pending alignment
  uint32_t ovf_n;
  uint32_t status;
  uint32_t aln_len;      // pairs in aln_node/aln_pos, stored in REVERSE order
  uint32_t best_row;     // row (node + 1) and DP column (1-based) the traceback starts from; 0,0 = none
  uint32_t best_col;
  int32_t best_score;
  uint32_t nseq_added;
  uint32_t scratch[4];
  // resumable window program (Poa::advance): the fill of the pending alignment runs between two advances
  uint32_t pc;           // kPc*
  uint32_t j;            // layer cursor of the current phase
  uint32_t k;            // prune round
  uint32_t nMain;        // rows of the whole current graph (rank order in r2n)
  uint32_t fill_layer;   // pending alignment: layer id, mode, sub-graph flag
  uint32_t fill_mode;
  uint32_t sub;          // the pending alignment runs on a Subgraph view (rank -> node through sl.order)
  uint32_t fill_k;       // words per lane the fill used for the pending alignment (row layout: 32 * fill_k words per half)
  uint32_t need;         // kNeed*
  uint32_t prep;         // kPrep* flags for step_prepare
  uint32_t wide;         // the pending alignment(s) run in the int32 / any-width kernel (wide_* below): a layer longer
                         // than the fast rows, or scores beyond the int16 range for this graph
  uint32_t order_ok;     // r2n / rank_of / owner / bsize / bstart describe the current graph (maintained by order_update)
  uint32_t round;        // the pending fill is a whole re-alignment round: one alignment per sequence of the window,
                         // all against the same frozen graph (AddWeights only adds to edge weights, so they commute)
  uint32_t jobs_total;   // alignments of the pending step handed to the align kernel / how many of them are finished
  uint32_t jobs_done;
  unsigned long long cells;
  uint32_t alignments;
  uint32_t sorts, sorts_hbm;   // TopologicalSort runs of this window / how many had to sort out of HBM (staging did not fit)
  // per-phase cycle counters (leader lane, clock64): see kPh*
  unsigned long long t_last;
  unsigned long long phase[12];
};

enum : int {
  kPhCsr = 0, kPhSort = 1, kPhRowprog = 2, kPhFill = 3, kPhTrace = 4, kPhAddAln = 5, kPhAddW = 6, kPhPrune = 7,
  kPhLargest = 8, kPhEmit = 9, kPhOther = 10, kPhCount = 11,
};

// ---------------------------------------------------------------------------------------------------
// H-matrix addressing.  A row holds 64*K cells as 32*K packed words, "lane-major": word w (0 <= w < 32*K) holds
// column w in its low half and column 32*K + w in its high half.  The fill's lane l owns the K consecutive
// words [l*K, l*K + K) (two runs of K consecutive columns); a run of consecutive columns of one half is a run
// of consecutive words, which is what the traceback's tiles fetch (16 B vectors).
template <int K>
struct RowMap {
  static constexpr int kCols = 64 * K;
  static constexpr int kWords = 32 * K;
  VGC_HD static VGC_INL uint32_t word(int lane, int k) { return lane * K + k; }
  VGC_HD static VGC_INL int32_t load(const uint32_t* row, uint32_t c) {
    const uint32_t h = c >= static_cast<uint32_t>(kWords) ? 1u : 0u;
    const uint32_t w = row[h ? c - kWords : c];
    return static_cast<int16_t>(h ? (w >> 16) : (w & 0xFFFFu));
  }
};

// Words per lane the fill uses for an alignment of `len` columns when the slot's rows hold K words per lane: the
// smallest supported width that fits (512-, 640- or 1024-column rows).  Narrow layers cost fewer DPX operations and
// fewer bytes per row; the row stride in HBM stays the slot's.
VGC_HD VGC_INL uint32_t fill_width(uint32_t K, uint32_t len) {
  if (K > 8 && len <= 512u) return 8u;
  if (K > 10 && len <= 640u) return 10u;
  return K;
}

// ---------------------------------------------------------------------------------------------------
// Traceback walker: ONE thread per alignment (the device runs 32 alignments per warp, the host model one).
// Replaces the traceback of SimdAlignmentEngine::Linear (simd_alignment_engine_implementation.hpp:908-1105,
// scalar twin sisd_alignment_engine.cpp:362-460).  A step compares the current cell with its candidates in the
// reference's priority order — diagonal over the predecessors (in-edge order), vertical over the predecessors,
// horizontal — and takes the first that reproduces the score.  The cells come from a private tile of the DP
// matrix (kTR consecutive rows in rank space x kTW consecutive words of the lane-major row) that the thread
// fetches with 16-byte loads whenever the walk leaves it: DRAM latency is paid once per tile, not once per step.
// Pairs are appended in reverse (end of the alignment first).
constexpr int kTR = 16;  // tile rows (ranks ti, ti-1, ..)
constexpr int kTW = 16;  // tile words per row
enum : int { kWalkStep = 0, kWalkMiss = 1, kWalkDone = 2, kWalkBad = 3 };

struct TraceWalker {
  // the alignment's DP matrix and row program
  const uint32_t* H;
  const int16_t* fc;
  const U4* rp;
  const uint32_t* ovf;
  const uint32_t* nodes;   // rank -> node id; nullptr: the id is in the row record (slots below 65536 nodes)
  const uint8_t* seq;      // the sequence's bases
  const uint8_t* dec;      // code -> base byte
  int32_t* aln_node;
  int32_t* aln_pos;
  uint32_t aln_cap, rw, half_words;
  int32_t m, x, g;
  bool sw;
  // tile storage (kTR * kTW words, kTR records)
  uint32_t* th;
  U4* tr;
  // walk state
  uint32_t i, j, n;
  int32_t h;
  U4 rec;
  uint32_t sb;             // seq[j - 1] (the base of DP column j), valid while j >= 1
  uint32_t ti, wb, rec_base;  // tile anchor row, first word, row of tr[0]
  bool have, fresh, started;

  VGC_HD VGC_INL uint32_t base_of(uint32_t code) const { return dec[code]; }
  VGC_HD VGC_INL static int32_t half_of(uint32_t v, uint32_t hi) {
    return static_cast<int16_t>(hi ? (v >> 16) : (v & 0xFFFFu));
  }

  VGC_HD VGC_INL void start(uint32_t row, uint32_t col) {
    i = row;
    j = col;
    n = 0;
    h = 0;
    rec = U4{0, 0, 0, 0};
    sb = col ? seq[col - 1] : 0u;
    ti = wb = 0;
    rec_base = 1;
    have = fresh = started = false;
  }

  // fetch the tile anchored at the current cell
  VGC_HD VGC_INL void refill() {
    ti = i;
    const uint32_t c = j ? j - 1 : 0;
    const uint32_t w = c >= half_words ? c - half_words : c;
    uint32_t b = w & ~3u;
    b = b >= static_cast<uint32_t>(kTW - 4) ? b - (kTW - 4) : 0u;
    if (b + kTW > half_words) b = half_words - kTW;
    wb = b;
#ifdef __CUDA_ARCH__
    // asynchronous 16-byte copies global -> shared (cp.async, SASS LDGSTS): no registers are tied up, so every
    // load of the tile is in flight at once and the refill costs one DRAM round trip.  (Per-thread TMA bulk copies
    // were measured 2x slower here: a tile is ~30 small pieces and the bulk-copy unit serialises them.)
    const uint32_t nrows = ti + 1 < static_cast<uint32_t>(kTR) ? ti + 1 : static_cast<uint32_t>(kTR);
    const uint32_t nrec = ti < static_cast<uint32_t>(kTR) ? ti : static_cast<uint32_t>(kTR);  // rows >= 1 have a record
    {
      const uint32_t* src = H + static_cast<uint64_t>(ti) * rw + wb;
      uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(th));
#pragma unroll 2
      for (uint32_t r = 0; r < nrows; ++r) {
#pragma unroll
        for (int q = 0; q < kTW / 4; ++q)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * 16), "l"(src + q * 4) : "memory");
        src -= rw;
        dst += kTW * 4;
      }
      // records of rows ti - nrec + 1 .. ti (record of row r is rp[r - 1]); tr[q] = record of row ti - nrec + 1 + q
      const U4* rsrc = rp + (ti - nrec);
      uint32_t rdst = static_cast<uint32_t>(__cvta_generic_to_shared(tr));
      for (uint32_t q = 0; q < nrec; ++q) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rdst), "l"(rsrc) : "memory");
        rsrc += 1;
        rdst += 16;
      }
    }
    rec_base = ti - nrec + 1;
    asm volatile("cp.async.wait_all;" ::: "memory");
#else
    for (uint32_t r = 0; r < static_cast<uint32_t>(kTR); ++r) {
      if (r > ti) break;
      const uint32_t row = ti - r;
      const U4* src = reinterpret_cast<const U4*>(H + static_cast<uint64_t>(row) * rw + wb);
      U4* dst = reinterpret_cast<U4*>(th + r * kTW);
      for (int q = 0; q < kTW / 4; ++q) dst[q] = src[q];
    }
    {
      const uint32_t nrec = ti < static_cast<uint32_t>(kTR) ? ti : static_cast<uint32_t>(kTR);
      rec_base = ti - nrec + 1;
      for (uint32_t q = 0; q < nrec; ++q) tr[q] = rp[rec_base + q - 1];
    }
#endif
    have = true;
    fresh = true;
  }

  // H(row, jj), jj = DP column (0 = first column).  false: not in the tile (and the tile is not fresh)
  VGC_HD VGC_INL bool cell(uint32_t row, uint32_t jj, int32_t* out) const {
    if (jj == 0) {
      *out = sw ? 0 : static_cast<int32_t>(fc[row]);
      return true;
    }
    const uint32_t c = jj - 1;
    const uint32_t hi = c >= half_words ? 1u : 0u;
    const uint32_t w = hi ? c - half_words : c;
    const uint32_t dr = ti - row, dq = w - wb;
    uint32_t v;
    if (have && dr < static_cast<uint32_t>(kTR) && dq < static_cast<uint32_t>(kTW)) v = th[dr * kTW + dq];
    else if (fresh) v = H[static_cast<uint64_t>(row) * rw + w];
    else return false;
    *out = static_cast<int16_t>(hi ? (v >> 16) : (v & 0xFFFFu));
    return true;
  }
  VGC_HD VGC_INL U4 rec_of(uint32_t row) const {
    if (row == 0) return U4{0, 0, 0, 0};
    const uint32_t q = row - rec_base;
    if (have && row <= ti && q < static_cast<uint32_t>(kTR)) return tr[q];
    return rp[row - 1];
  }
  VGC_HD VGC_INL uint32_t pred(uint32_t np, uint32_t p) const {
    if (np == 0) return 0u;  // no in-edges: the virtual row 0 is the predecessor
    return rec_pred(rec, i, p, ovf);
  }

  VGC_HD VGC_INL int step() {
    if (!have) return kWalkMiss;  // first call: fetch the tile, then read the start cell
    if (!started) {
      int32_t v = 0;
      cell(i, j, &v);
      h = v;
      rec = rec_of(i);
      started = true;
    }
    if (sw) {
      if (h == 0) return kWalkDone;
    } else {
      if (i == 0 && j == 0) return kWalkDone;
    }
    const uint32_t np = i != 0 ? meta_npred(rec.x) : 0u;
    if (i != 0 && j >= 2 && (rec.x & kMetaInline)) {
      // ---- fast path (<= 6 predecessors inline, away from the borders): every candidate cell is read from the
      //      tile at once and the first match in the reference's priority order (diagonals in in-edge order,
      //      verticals in in-edge order, horizontal) is selected without branching, so the 32 walks of a warp
      //      stay converged
      const uint32_t nb = seq[j - 2];  // base of DP column j - 1, consumed when the move changes the column
      const uint32_t npp = np ? np : 1u;
      uint32_t dl[kInlinePreds];
#pragma unroll
      for (uint32_t p = 0; p < kInlinePreds; ++p) dl[p] = rec_delta(rec, p);
      uint32_t maxd = dl[0];
#pragma unroll
      for (uint32_t p = 1; p < kInlinePreds; ++p) maxd = dl[p] > maxd ? dl[p] : maxd;
      const uint32_t c1 = j - 1, c0 = j - 2;
      const uint32_t hi1 = c1 >= half_words ? 1u : 0u, hi0 = c0 >= half_words ? 1u : 0u;
      const uint32_t dq1 = (hi1 ? c1 - half_words : c1) - wb, dq0 = (hi0 ? c0 - half_words : c0) - wb;
      const uint32_t dri = ti - i;
      const uint32_t TRu = static_cast<uint32_t>(kTR), TWu = static_cast<uint32_t>(kTW);
      if (dq1 < TWu && dq0 < TWu && dri < TRu && dri + maxd < TRu) {
        const uint32_t* base = th + dri * kTW;
        const int32_t mcf = base_of(meta_code(rec.x)) == sb ? m : x;
        // a candidate matches iff its 16-bit cell equals h - (its move's score): compare raw halves, no sign extension
        const uint32_t tD = static_cast<uint32_t>(h - mcf) & 0xFFFFu, tV = static_cast<uint32_t>(h - g) & 0xFFFFu;
        const uint32_t sh0 = hi0 * 16u, sh1 = hi1 * 16u;
        // bit p: diagonal over predecessor p, bit 6 + p: vertical, bit 12: horizontal — the reference's priority
        // order is the bit order, so the winner is the lowest set bit
        uint32_t mask = (((base[dq0] >> sh0) & 0xFFFFu) == tV) ? (1u << 12) : 0u;
#pragma unroll
        for (uint32_t p = 0; p < kInlinePreds; ++p) {
          const uint32_t* rowp = base + dl[p] * kTW;
          mask |= (((rowp[dq0] >> sh0) & 0xFFFFu) == tD) ? (1u << p) : 0u;
          mask |= (((rowp[dq1] >> sh1) & 0xFFFFu) == tV) ? (64u << p) : 0u;
        }
        const uint32_t live = (1u << npp) - 1u;
        mask &= live | (live << 6) | (1u << 12);
        if (mask == 0 || n >= aln_cap) return kWalkBad;
#ifdef __CUDA_ARCH__
        const uint32_t sel = static_cast<uint32_t>(__ffs(static_cast<int>(mask))) - 1u;
#else
        uint32_t sel = 0;
        while (!((mask >> sel) & 1u)) ++sel;
#endif
        const bool horiz = sel == 12, vert = sel >= 6 && sel < 12;
        const uint32_t psel = sel >= 6 ? sel - 6 : sel;
        const uint32_t dsel = horiz ? 0u : rec_delta(rec, psel);
        const uint32_t pi = i - dsel;
        aln_node[n] = horiz ? -1 : static_cast<int32_t>(nodes ? nodes[i - 1] : meta_node(rec.x));
        aln_pos[n] = vert ? -1 : static_cast<int32_t>(j - 1);
        ++n;
        // the score of the cell moved to: the compare above proved it equals h minus the move's score
        h = (horiz || vert) ? h - g : h - mcf;
        if (!horiz) rec = pi ? tr[pi - rec_base] : U4{0, 0, 0, 0};
        i = pi;
        if (!vert) {
          j = j - 1;
          sb = nb;
        }
        fresh = false;
        return kWalkStep;
      }
      if (!fresh) return kWalkMiss;
    }
    const uint32_t npp = i != 0 ? (np == 0 ? 1u : np) : 0u;
    int32_t mc = 0;
    if (i != 0 && j != 0) mc = (base_of(meta_code(rec.x)) == sb) ? m : x;
    uint32_t pi = i, pj = j;
    int32_t hn = 0;
    bool found = false;
    if (j != 0) {
      for (uint32_t p = 0; p < npp; ++p) {
        const uint32_t pr = pred(np, p);
        int32_t hv;
        if (!cell(pr, j - 1, &hv)) return kWalkMiss;
        if (h == hv + mc) {
          found = true;
          pi = pr;
          pj = j - 1;
          hn = hv;
          break;
        }
      }
    }
    if (!found) {
      for (uint32_t p = 0; p < npp; ++p) {
        const uint32_t pr = pred(np, p);
        int32_t hv;
        if (!cell(pr, j, &hv)) return kWalkMiss;
        if (h == hv + g) {
          found = true;
          pi = pr;
          hn = hv;
          break;
        }
      }
    }
    if (!found && j != 0) {
      int32_t hv;
      if (!cell(i, j - 1, &hv)) return kWalkMiss;
      if (h == hv + g) {
        found = true;
        pj = j - 1;
        hn = hv;
      }
    }
    if (!found || n >= aln_cap) return kWalkBad;
    aln_node[n] = (i == pi) ? -1 : static_cast<int32_t>(nodes ? nodes[i - 1] : meta_node(rec.x));
    aln_pos[n] = (j == pj) ? -1 : static_cast<int32_t>(j - 1);
    ++n;
    if (pi != i) rec = rec_of(pi);
    if (pj != j) sb = pj ? seq[pj - 1] : 0u;
    i = pi;
    j = pj;
    h = hn;
    fresh = false;
    return kWalkStep;
  }
};

// ---------------------------------------------------------------------------------------------------
// The algorithm.  Ex (executor) provides: lane(), width(), leader(), sync(), atomic_add(u32*,u32),
// excl_scan(v, &total), fill<K>(...) and the flag/CSR staging storage.
// ---------------------------------------------------------------------------------------------------
// Wide path: alignments the packed int16 kernels cannot take (a layer longer than their widest row, or scores that
// may leave the int16 range on this graph) run on an int32 matrix of any width.  Replaces the same reference code as
// the fast path (simd_alignment_engine_implementation.hpp:760-1105 with the int32 lanes the reference selects at
// :699-706).  The matrix is row-major: H[row * cols + j], cols = len + 1, column 0 = the first column, row 0 = the
// virtual row; rows in rank space like everywhere else.  The fill is in poa_wide.cuh (device) and in the host model;
// the traceback below is shared by both: every lane computes the same walk, the leader writes.
struct WideIo {
  const int32_t* H;
  uint32_t cols;
  const U4* rp;            // row program (rank order)
  const uint32_t* ovf;
  const uint32_t* nodes;   // rank -> node id; nullptr: the id is in the row record
  const uint8_t* codes;    // codes of the sequence
  int32_t m, x, g;
  bool sw;
  uint32_t row, col;       // start cell; 0,0 = empty alignment
  uint32_t max_steps;
  int32_t* aln_node;       // WEIGHTS == false: the alignment, reversed
  int32_t* aln_pos;
  uint32_t aln_cap;
  uint32_t* ew;            // WEIGHTS == true: Graph::AddWeights fused into the walk (see poa_trace.cuh)
  const uint32_t* ieid;
  uint32_t in_stride;
  const uint8_t* quals;    // qualities of the sequence (nullptr: every base weighs 1)
  const uint32_t* wlut;    // quality byte -> weight
};

template <bool WEIGHTS, class Ex>
VGC_HD int wide_trace(Ex& ex, const WideIo& t, uint32_t* n_out) {
  uint32_t i = t.row, j = t.col, n = 0;
  *n_out = 0;
  if (i == 0 && j == 0) return kWalkDone;
  const int32_t m = t.m, x = t.x, g = t.g;
  const uint64_t cols = t.cols;
  uint32_t pend = kNone, pend_w = 0;  // WEIGHTS: edge of the matched pair emitted last
  while (true) {
    if (!t.sw && i == 0 && j == 0) break;
    if (n >= t.max_steps) return kWalkBad;
    const int32_t h = t.H[i * cols + j];
    if (t.sw && h == 0) break;
    U4 rec = {0, 0, 0, 0};
    if (i != 0) rec = t.rp[i - 1];
    const uint32_t np = i != 0 ? meta_npred(rec.x) : 0u;
    const uint32_t npp = i != 0 ? (np == 0 ? 1u : np) : 0u;
    int32_t mc = 0;
    if (i != 0 && j != 0) mc = meta_code(rec.x) == t.codes[j - 1] ? m : x;
    uint32_t kind = 3, psel = 0, pi = i;
    if (j != 0) {
      for (uint32_t p = 0; p < npp && kind == 3; ++p) {
        const uint32_t pr = np == 0 ? 0u : rec_pred(rec, i, p, t.ovf);
        if (h == t.H[pr * cols + j - 1] + mc) {
          kind = 0;
          psel = p;
          pi = pr;
        }
      }
    }
    for (uint32_t p = 0; p < npp && kind == 3; ++p) {
      const uint32_t pr = np == 0 ? 0u : rec_pred(rec, i, p, t.ovf);
      if (h == t.H[pr * cols + j] + g) {
        kind = 1;
        psel = p;
        pi = pr;
      }
    }
    if (kind == 3 && j != 0 && h == t.H[i * cols + j - 1] + g) {
      kind = 2;
      pi = i;
    }
    if (kind == 3) return kWalkBad;
    const uint32_t nd = i != 0 ? (t.nodes ? t.nodes[i - 1] : meta_node(rec.x)) : 0u;
    if (WEIGHTS) {
      if (kind == 0) {
        if (pend != kNone && ex.leader()) ex.atomic_add(t.ew + pend, pend_w);
        pend = np == 0 ? kNone : t.ieid[static_cast<uint64_t>(nd) * t.in_stride + psel];
        // weight(pos - 1) + weight(pos) for pos = j - 1 (a pair at pos 0 is never followed by a matched pair)
        pend_w = j >= 2 ? (t.quals ? t.wlut[t.quals[j - 2]] + t.wlut[t.quals[j - 1]] : 2u) : 0u;
      } else {
        pend = kNone;
      }
    } else {
      if (n >= t.aln_cap) return kWalkBad;
      if (ex.leader()) {
        t.aln_node[n] = kind == 2 ? -1 : static_cast<int32_t>(nd);
        t.aln_pos[n] = kind == 1 ? -1 : static_cast<int32_t>(j - 1);
      }
    }
    ++n;
    i = pi;
    if (kind != 1) j = j - 1;
  }
  *n_out = n;
  return kWalkDone;
}

template <class Ex, int K>
struct Poa {
  Ex& ex;
  const BatchView& bv;
  Slot& sl;
  WinState& ws;
  Scores nw;      // NW engine scores (params)
  Scores sw;      // SW engine: 3/-5/-4 hard-wired (window.cpp:326)
  using RM = RowMap<K>;
  // staged graph of the last sort_graph() of this step (executor fast storage): records + 16-bit adjacency
  bool staged = false;
  bool ranks_ok = false;  // rank_of already holds the ranks of r2n (incremental order, no sort in this step)
  const uint32_t* st_rec = nullptr;
  const uint16_t* st_adj = nullptr;

  VGC_HD Poa(Ex& e, const BatchView& b, Slot& s, WinState& w, Scores nw_) : ex(e), bv(b), sl(s), ws(w), nw(nw_) {
    sw.m = 3;
    sw.x = -5;
    sw.g = -4;
  }

  VGC_HD VGC_INL Graph& G() { return sl.g[ws.cur]; }

  // attribute the cycles since the previous tick to `phase`
  VGC_HD VGC_INL void tick(int phase) {
    if (ex.leader()) {
      const unsigned long long now = ex.clock();
      ws.phase[phase] += now - ws.t_last;
      ws.t_last = now;
    }
  }

  VGC_HD VGC_INL void fail(uint32_t st) {
    if (ws.status == kStOk) ws.status = st;
  }

  // ---- CSR of the out-edges of the live graph (LargestSubgraph, heaviest bundle): parallel over edges ----
  VGC_HD void build_out_csr() {
    Graph& g = G();
    const uint32_t nV = g.nV, nE = g.nE;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nV; base += ex.width()) {
      const uint32_t v = base + ex.lane();
      uint32_t tot;
      const uint32_t q = ex.excl_scan(v < nV ? g.nout[v] : 0, &tot);
      if (v < nV) sl.out_off[v] = carry + q;
      carry += tot;
    }
    if (ex.leader()) sl.out_off[nV] = carry;
    ex.sync();
    for (uint32_t e = ex.lane(); e < nE; e += ex.width()) sl.out_eid[sl.out_off[g.etail[e]] + g.eout_ord[e]] = e;
    ex.sync();
  }

  // ---- adjacency view of the sorts: adj[off[v] .. off[v+1]) = in-edge tails of v in creation order followed
  //      by its aligned nodes in list order; flags[v] bits 5..7 = number of aligned entries.

  // ---- graph.cpp:301-371.  Serial (leader).  `member_only`: the Subgraph view (graph.cpp:694-727):
  //      roots, in-edges and aligned links restricted to nodes with kFMember.  Output: dst[0..n).
  template <class IdxT, class StkT>
  VGC_HD uint32_t toposort_impl(uint8_t* flags, const IdxT* off, const IdxT* adj, StkT* stack, uint32_t stack_cap,
                                bool member_only, uint32_t* dst, bool* overflow, uint32_t* own = nullptr) {
    const uint32_t nV = G().nV;
    uint32_t n = 0, sp = 0;
    *overflow = false;
    for (uint32_t root = 0; root < nV; ++root) {
      const uint8_t fr = flags[root];
      if ((fr & kFMarkMask) != 0) continue;
      if (member_only && !(fr & kFMember)) continue;
      stack[sp++] = static_cast<StkT>(root);
      while (sp > 0) {
        const uint32_t curr = stack[sp - 1];
        const uint8_t fc = flags[curr];
        if ((fc & kFMarkMask) == 2) {
          --sp;
          continue;
        }
        const uint32_t b = off[curr], e = off[curr + 1];
        const uint32_t in_e = b + G().nin[curr];  // adjacency = in-tails, then the aligned nodes
        if (sp + (e - b) + 1 > stack_cap) {
          *overflow = true;
          return 0;
        }
        bool valid = true;
        for (uint32_t i = b; i < in_e; ++i) {
          const uint32_t t = adj[i];
          const uint8_t ft = flags[t];
          if (member_only && !(ft & kFMember)) continue;
          if ((ft & kFMarkMask) != 2) {
            stack[sp++] = static_cast<StkT>(t);
            valid = false;
          }
        }
        const bool primary = !(fc & kFIgnored);
        if (primary) {
          for (uint32_t i = in_e; i < e; ++i) {
            const uint32_t a = adj[i];
            const uint8_t fa = flags[a];
            if (member_only && !(fa & kFMember)) continue;
            if ((fa & kFMarkMask) != 2) {
              stack[sp++] = static_cast<StkT>(a);
              flags[a] = fa | kFIgnored;
              valid = false;
            }
          }
        }
        if (valid) {
          flags[curr] = (fc & ~kFMarkMask) | 2;
          if (primary) {
            if (own) own[curr] = root;
            dst[n++] = curr;
            for (uint32_t i = in_e; i < e; ++i) {
              const uint32_t a = adj[i];
              if (member_only && !(flags[a] & kFMember)) continue;
              if (own) own[a] = root;
              dst[n++] = a;
            }
          }
          --sp;
        } else {
          flags[curr] = (fc & ~kFMarkMask) | 1;
        }
      }
    }
    return n;
  }

  // ---- graph.cpp:640-666: nodes reachable backwards (in-edges + aligned links) from `from`, ids >= floor
  template <class IdxT>
  VGC_HD void extract_impl(uint8_t* flags, const IdxT* off, const IdxT* adj, uint32_t* stack, uint32_t from,
                           uint32_t floor_id) {
    uint32_t sp = 0;
    stack[sp++] = from;
    while (sp > 0) {
      const uint32_t curr = stack[--sp];
      const uint8_t f = flags[curr];
      if (!(f & kFMember) && curr >= floor_id) {
        // pushes are bounded by nE + sum(nal) + 1: the stack is scratch far larger than that
        for (uint32_t i = off[curr]; i < off[curr + 1]; ++i) stack[sp++] = adj[i];
        flags[curr] = f | kFMember;
      }
    }
  }

  // ---- the same sort over the staged graph of the executor's fast storage.  One 32-bit record per node:
  //      adjacency offset (bits 0-15) | in-degree (16-21) | aligned count (22-25) | expanded (26) | done (27) |
  //      ignored (28) | member (29).  A node is scanned once: the first visit pushes what is not done yet and
  //      marks it expanded; when it surfaces again everything it pushed is done (the graph is a DAG), so the second
  //      visit only emits — the reference re-scans and finds exactly that (graph.cpp:318-352).
  static constexpr uint32_t kRExpanded = 1u << 26, kRDone = 1u << 27, kRIgnored = 1u << 28, kRMember = 1u << 29;
  VGC_HD VGC_INL static uint32_t rec_pack(uint32_t off, uint32_t nin, uint32_t nal) { return off | (nin << 16) | (nal << 22); }

  template <bool SUB, class StkT>
  VGC_HD uint32_t toposort_fast(uint32_t* rec, const uint16_t* adj, StkT* stack, uint32_t stack_cap, uint32_t* dst,
                                uint32_t* rank_of, bool* overflow, uint32_t* own = nullptr) {
    const uint32_t nV = G().nV;
    uint32_t n = 0;
    *overflow = false;
    for (uint32_t root = 0; root < nV; ++root) {
      const uint32_t rr = rec[root];
      if (rr & (kRDone | kRExpanded)) continue;
      if (SUB && !(rr & kRMember)) continue;
      // (curr, r) = the top of the stack, kept in registers: after a node pushes its pending tails the next node to
      // look at is the last one pushed, whose record was just read — no reload on the way down
      uint32_t sp = 1, curr = root, r = rr;
      stack[0] = static_cast<StkT>(root);
      while (true) {
        bool pop = (r & kRDone) != 0;
        if (!pop) {
          const uint32_t off = r & 0xFFFFu, nin = (r >> 16) & 63u, nal = (r >> 22) & 15u;
          const bool primary = !(r & kRIgnored);
          uint32_t last = kNone, last_r = 0;
          if (!(r & kRExpanded)) {
            if (sp + nin + nal + 1 > stack_cap) {
              *overflow = true;
              return 0;
            }
#pragma unroll 1
            for (uint32_t i = 0; i < nin; ++i) {  // degrees are 1-3: an unrolled body would mostly run predicated off
              const uint32_t t = adj[off + i];
              const uint32_t rt = rec[t];
              if (SUB && !(rt & kRMember)) continue;
              if (!(rt & kRDone)) {
                stack[sp++] = static_cast<StkT>(t);
                last = t;
                last_r = rt;
              }
            }
            if (primary) {
#pragma unroll 1
              for (uint32_t i = 0; i < nal; ++i) {
                const uint32_t a = adj[off + nin + i];
                const uint32_t ra = rec[a];
                if (SUB && !(ra & kRMember)) continue;
                if (!(ra & kRDone)) {
                  stack[sp++] = static_cast<StkT>(a);
                  rec[a] = ra | kRIgnored;
                  last = a;
                  last_r = ra | kRIgnored;
                }
              }
            }
          }
          if (last == kNone) {
            // nothing pending below: emit (second visit, or a first visit whose tails were all done)
            rec[curr] = r | kRDone;
            if (primary) {
              rank_of[curr] = n;
              if (!SUB && own) own[curr] = root;
              dst[n++] = curr;
#pragma unroll 1
              for (uint32_t i = 0; i < nal; ++i) {
                const uint32_t a = adj[off + nin + i];
                if (SUB && !(rec[a] & kRMember)) continue;
                rank_of[a] = n;
                if (!SUB && own) own[a] = root;
                dst[n++] = a;
              }
            }
            pop = true;
          } else {
            rec[curr] = r | kRExpanded;
            curr = last;
            r = last_r;  // nothing was written to that record after it was read (the last push is the last write)
          }
        }
        if (pop) {
          if (--sp == 0) break;
          curr = stack[sp - 1];
          r = rec[curr];
        }
      }
    }
    return n;
  }

  VGC_HD void extract_fast(uint32_t* rec, const uint16_t* adj, uint32_t* stack, uint32_t from, uint32_t floor_id) {
    uint32_t sp = 0;
    stack[sp++] = from;
    while (sp > 0) {
      const uint32_t curr = stack[--sp];
      const uint32_t r = rec[curr];
      if (!(r & kRMember) && curr >= floor_id) {
        const uint32_t off = r & 0xFFFFu, cnt = ((r >> 16) & 63u) + ((r >> 22) & 15u);
#pragma unroll 1
        for (uint32_t i = 0; i < cnt; ++i) stack[sp++] = adj[off + i];
        rec[curr] = r | kRMember;
      }
    }
  }

  // Topological order of the live graph (or of the Subgraph view begin..end) into dst.  Builds the adjacency
  // view in the executor's fast storage (16-bit ids) when it fits, else in the H scratch (32-bit ids).
  // H scratch layout (words): [0, nV] offsets | [nV+1, nV+1+nA) adjacency | big stack after that.
  VGC_HD uint32_t sort_graph(bool sub, uint32_t sub_begin, uint32_t sub_end, uint32_t* dst) {
    Graph& g = G();
    const uint32_t nV = g.nV, nE = g.nE;
    const int W = ex.width(), L = ex.lane();
    // ranks of a Subgraph view go to their own array (r2n2 is dead between order updates): rank_of keeps describing
    // the main order, which order_update maintains incrementally
    uint32_t* rk = sub ? sl.r2n2 : sl.rank_of;
    uint32_t* goff = sl.H;
    // offsets = exclusive scan of (in-degree + aligned count)
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nV; base += W) {
      const uint32_t v = base + L;
      const uint32_t c = v < nV ? g.nin[v] + g.nal[v] : 0;
      uint32_t tot;
      const uint32_t p = ex.excl_scan(c, &tot);
      if (v < nV) goff[v] = carry + p;
      carry += tot;
    }
    const uint32_t nA = carry;
    if (ex.leader()) goff[nV] = nA;
    ex.sync();
    uint32_t* rec;
    uint16_t *adj16, *stk16;
    uint32_t stk_cap;
    uint32_t max_in = 0;
    for (uint32_t v = L; v < nV; v += W) max_in = g.nin[v] > max_in ? g.nin[v] : max_in;
    max_in = ex.reduce_max(max_in);
    const bool fast = max_in < 64u && ex.stage_fast(nV, nA, &rec, &adj16, &stk16, &stk_cap);
    uint8_t* fl = sl.flags;
    uint32_t* gadj = sl.H + nV + 1;
    uint32_t* gstack = gadj + nA;
    for (uint32_t v = L; v < nV; v += W) {
      const uint32_t na = g.nal[v];
      const uint32_t o = goff[v] + g.nin[v];
      for (uint32_t i = 0; i < na; ++i) {
        const uint32_t a = g.al[v * sl.al_stride + i];
        if (fast) adj16[o + i] = static_cast<uint16_t>(a);
        else gadj[o + i] = a;
      }
      if (fast) rec[v] = rec_pack(goff[v], g.nin[v], na);
      else fl[v] = static_cast<uint8_t>(na ? kFHasAligned : 0);
    }
    for (uint32_t e = L; e < nE; e += W) {
      const uint32_t pos = goff[g.ehead[e]] + g.ein_ord[e];
      if (fast) adj16[pos] = static_cast<uint16_t>(g.etail[e]);
      else gadj[pos] = g.etail[e];
    }
    ex.sync();
    if (ex.leader()) {
      uint32_t n = 0;
      bool ovf = false;
      if (fast) {
        if (sub) {
          extract_fast(rec, adj16, gstack, sub_end, sub_begin);
          n = toposort_fast<true, uint16_t>(rec, adj16, stk16, stk_cap, dst, rk, &ovf);
        } else {
          n = toposort_fast<false, uint16_t>(rec, adj16, stk16, stk_cap, dst, rk, &ovf, sl.owner);
        }
        if (ovf) {
          // deep recursion: redo with the big stack in HBM (records: clear the marks, keep membership)
          for (uint32_t v = 0; v < nV; ++v) rec[v] &= ~(kRExpanded | kRDone | kRIgnored);
          if (sub) n = toposort_fast<true, uint32_t>(rec, adj16, gstack, 0xFFFFFFFFu, dst, rk, &ovf);
          else n = toposort_fast<false, uint32_t>(rec, adj16, gstack, 0xFFFFFFFFu, dst, rk, &ovf, sl.owner);
        }
      } else {
        if (sub) extract_impl<uint32_t>(fl, goff, gadj, gstack, sub_end, sub_begin);
        n = toposort_impl<uint32_t, uint32_t>(fl, goff, gadj, gstack, 0xFFFFFFFFu, sub, dst, &ovf, sub ? nullptr : sl.owner);
      }
      ws.scratch[0] = n;
    }
    ex.sync();
    const uint32_t n = ws.scratch[0];
    if (ex.leader()) {
      ws.sorts += 1;
      ws.sorts_hbm += fast ? 0u : 1u;
    }
    staged = fast;  // the row-program builder reads the adjacency (and the ranks the sort wrote) from here
    st_rec = rec;
    st_adj = adj16;
    if (!sub) {
      // block tables of the incremental order: blocks are runs of equal owner in `dst`
      for (uint32_t v = L; v < nV; v += W) {
        sl.bsize[v] = 0;
        sl.dirty[v] = 0;
      }
      ex.sync();
      for (uint32_t q = L; q < n; q += W) {
        const uint32_t r = sl.owner[dst[q]];
        if (q == 0 || sl.owner[dst[q - 1]] != r) sl.bstart[r] = q;
      }
      ex.sync();
      for (uint32_t q = L; q < n; q += W) {
        const uint32_t r = sl.owner[dst[q]];
        if (q + 1 == n || sl.owner[dst[q + 1]] != r) sl.bsize[r] = q + 1 - sl.bstart[r];
      }
      if (ex.leader()) ws.order_ok = (n == nV) ? 1u : 0u;
      ex.sync();
    }
    if (sub) {
      // keep the membership where the row-program builder can see it
      for (uint32_t v = L; v < nV; v += W) sl.flags[v] = fast ? ((rec[v] & kRMember) ? kFMember : 0) : (fl[v] & kFMember);
      ex.sync();
    }
    return n;
  }

  // ---- row program: per rank {node, meta(code, npred, sink), p0, p1}; rows = node + 1, 0 = virtual.
  //      npred > 2: p1 = offset into ovf[] holding predecessor rows 1..npred-1.
  VGC_HD void build_rowprog(const uint32_t* order, uint32_t nR, bool sub) {
    Graph& g = G();
    const uint32_t S = sl.in_stride;
    uint32_t* rk = sub ? sl.r2n2 : sl.rank_of;  // see sort_graph
    if (ex.leader()) {
      ws.ovf_n = 0;
      ws.nR = nR;
      ws.sub = sub ? 1u : 0u;
    }
    if (sub) {
      for (uint32_t r = ex.lane(); r < nR; r += ex.width()) sl.tmp0[order[r]] = 0;
      ex.sync();
      for (uint32_t r = ex.lane(); r < nR; r += ex.width()) {
        const uint32_t v = order[r];
        for (uint32_t i = 0; i < g.nin[v]; ++i) {
          const uint32_t t = g.itail[v * S + i];
          if (sl.flags[t] & kFMember) ex.atomic_add(&sl.tmp0[t], 1u);
        }
      }
    }
    // with a staged sort the DFS wrote rank_of itself and the in-tails are in the staged adjacency (same order); an
    // order maintained incrementally (order_update) comes with its ranks too
    const bool st = staged;
    if (!st && !(ranks_ok && !sub)) {
      for (uint32_t r = ex.lane(); r < nR; r += ex.width()) rk[order[r]] = r;
    }
    ex.sync();
    if (!st && !sub) {
      // ---- no staged copy (the sort was skipped): everything comes from HBM through three levels of dependent loads
      //      (rank -> node -> in-tails -> their ranks), so every lane works on four rows at a time to keep
      //      4 x (1 + 3 + 6 + 6) loads in flight instead of one chain
      constexpr int kB = 4;
      const int W = ex.width(), L = ex.lane();
      for (uint32_t r0 = 0; r0 < nR; r0 += kB * W) {
        uint32_t v[kB], ni[kB], cd[kB], no[kB], tl[kB][kInlinePreds], tr[kB][kInlinePreds];
#pragma unroll
        for (int t = 0; t < kB; ++t) {
          const uint32_t r = r0 + t * W + L;
          v[t] = r < nR ? order[r] : 0u;
        }
#pragma unroll
        for (int t = 0; t < kB; ++t) {
          const uint32_t r = r0 + t * W + L;
          ni[t] = r < nR ? g.nin[v[t]] : 0u;
          cd[t] = g.code[v[t]];
          no[t] = g.nout[v[t]];
#pragma unroll
          for (uint32_t i = 0; i < kInlinePreds; ++i) tl[t][i] = g.itail[v[t] * S + (i < S ? i : 0u)];
        }
#pragma unroll
        for (int t = 0; t < kB; ++t) {
#pragma unroll
          for (uint32_t i = 0; i < kInlinePreds; ++i) tr[t][i] = i < ni[t] ? rk[tl[t][i]] : 0u;
        }
#pragma unroll
        for (int t = 0; t < kB; ++t) {
          const uint32_t r = r0 + t * W + L;
          if (r >= nR) continue;
          const uint32_t row = r + 1, np = ni[t];
          uint32_t d[kInlinePreds], far = 0;
#pragma unroll
          for (uint32_t i = 0; i < kInlinePreds; ++i) {
            d[i] = i < np ? row - (tr[t][i] + 1) : row;
            far |= d[i] > 0xFFFFu ? 1u : 0u;
          }
          const bool inl = np <= kInlinePreds && !far && row <= 0xFFFFu;
          U4 rec;
          if (inl) {
            const uint32_t fill_d = d[0];
#pragma unroll
            for (uint32_t i = 1; i < kInlinePreds; ++i) d[i] = i < np ? d[i] : fill_d;
            rec.y = d[0] | (d[1] << 16);
            rec.z = d[2] | (d[3] << 16);
            rec.w = d[4] | (d[5] << 16);
          } else {
            const uint32_t o = ex.atomic_add(&ws.ovf_n, np);
            for (uint32_t i = 0; i < np; ++i) sl.ovf[o + i] = rk[g.itail[v[t] * S + i]] + 1;
            rec.y = rec.z = 0;
            rec.w = o;
          }
          if (np > kMetaMaxPred) fail(kStDegreeOverflow);
          rec.x = meta_pack(cd[t], np, no[t] == 0, inl, v[t], inl && np <= 1 && d[0] == 1);
          *reinterpret_cast<U4*>(sl.rowprog + 4 * static_cast<size_t>(r)) = rec;
        }
      }
      ex.sync();
      return;
    }
    // DP rows live in rank space: row = rank + 1 (0 = the virtual row)
    for (uint32_t r = ex.lane(); r < nR; r += ex.width()) {
      const uint32_t v = order[r];
      const uint32_t row = r + 1;
      uint32_t np = 0, far = 0;
      uint32_t d[kInlinePreds] = {row, row, row, row, row, row};  // np == 0: the virtual row, d_0 = row
      const uint32_t rv = st ? st_rec[v] : 0u;
      const uint32_t b = st ? (rv & 0xFFFFu) : v * S;
      const uint32_t e = b + (st ? ((rv >> 16) & 63u) : g.nin[v]);
      for (uint32_t i = b; i < e; ++i) {
        const uint32_t t = st ? static_cast<uint32_t>(st_adj[i]) : g.itail[i];
        if (sub && !(st ? (st_rec[t] & kRMember) != 0 : (sl.flags[t] & kFMember) != 0)) continue;
        const uint32_t dist = row - (rk[t] + 1);
        if (np < kInlinePreds) d[np] = dist;
        far |= dist > 0xFFFFu ? 1u : 0u;
        ++np;
      }
      const bool inl = np <= kInlinePreds && !far && row <= 0xFFFFu;
      U4 rec;
      if (inl) {
        const uint32_t fill_d = d[0];  // unused slots repeat the first predecessor (harmless duplicates)
        for (uint32_t q = np ? np : 1u; q < kInlinePreds; ++q) d[q] = fill_d;
        rec.y = d[0] | (d[1] << 16);
        rec.z = d[2] | (d[3] << 16);
        rec.w = d[4] | (d[5] << 16);
      } else {
        const uint32_t o = ex.atomic_add(&ws.ovf_n, np);
        uint32_t k = 0;
        for (uint32_t i = v * S; i < v * S + g.nin[v]; ++i) {
          const uint32_t t = g.itail[i];
          if (sub && !(sl.flags[t] & kFMember)) continue;
          sl.ovf[o + k++] = rk[t] + 1;
        }
        rec.y = rec.z = 0;
        rec.w = o;
      }
      const bool sink = sub ? (sl.tmp0[v] == 0) : (g.nout[v] == 0);
      if (np > kMetaMaxPred) fail(kStDegreeOverflow);
      rec.x = meta_pack(g.code[v], np, sink, inl, v, inl && np <= 1 && d[0] == 1);
      *reinterpret_cast<U4*>(sl.rowprog + 4 * static_cast<size_t>(r)) = rec;
    }
    ex.sync();
  }

  // ---- traceback (host model / single-lane executors): drive one TraceWalker to the end.  The device runs the
  //      same walker, 32 alignments per warp, in trace_kernel (vgc_engine.cu).
  VGC_HD void init_walker(TraceWalker& t, uint32_t layer, uint32_t mode, uint32_t* th, U4* tr) {
    const Scores& sc = mode == kModeNW ? nw : sw;
    t.H = sl.H;
    t.fc = sl.fc;
    t.rp = reinterpret_cast<const U4*>(sl.rowprog);
    t.ovf = sl.ovf;
    t.nodes = sl.max_nodes < 65536u ? nullptr : (ws.sub ? sl.order : sl.r2n);
    t.seq = bv.bases + bv.seq_off[layer];
    t.dec = bv.decoder;
    t.aln_node = sl.aln_node;
    t.aln_pos = sl.aln_pos;
    t.aln_cap = sl.aln_cap;
    t.rw = sl.row_words;
    t.half_words = 32u * ws.fill_k;
    t.m = sc.m;
    t.x = sc.x;
    t.g = sc.g;
    t.sw = mode == kModeSW;
    t.th = th;
    t.tr = tr;
    t.start(ws.best_row, ws.best_col);
  }

  VGC_HD void traceback(uint32_t layer, uint32_t mode) {
    if (ws.wide) {
      const Scores& sc = mode == kModeNW ? nw : sw;
      WideIo t;
      t.H = reinterpret_cast<const int32_t*>(sl.H);
      t.cols = layer_len(layer) + 1;
      t.rp = reinterpret_cast<const U4*>(sl.rowprog);
      t.ovf = sl.ovf;
      t.nodes = sl.max_nodes < 65536u ? nullptr : (ws.sub ? sl.order : sl.r2n);
      t.codes = ex.seq_codes();
      t.m = sc.m;
      t.x = sc.x;
      t.g = sc.g;
      t.sw = mode == kModeSW;
      t.row = ws.best_row;
      t.col = ws.best_col;
      t.max_steps = ws.nR + layer_len(layer) + 2;
      t.aln_node = sl.aln_node;
      t.aln_pos = sl.aln_pos;
      t.aln_cap = sl.aln_cap;
      t.ew = nullptr;
      t.ieid = nullptr;
      t.in_stride = 0;
      t.quals = nullptr;
      t.wlut = nullptr;
      uint32_t n = 0;
      if (wide_trace<false>(ex, t, &n) != kWalkDone) fail(kStInternal);
      if (ex.leader()) ws.aln_len = n;
      ex.sync();
      return;
    }
    if (ex.leader()) {
      if (ws.best_row == 0 && ws.best_col == 0) {
        ws.aln_len = 0;
      } else {
        TraceWalker t;
        uint32_t* th;
        U4* tr;
        ex.trace_tile(&th, &tr);
        init_walker(t, layer, mode, th, tr);
        int st;
        while ((st = t.step()) < kWalkDone) {
          if (st == kWalkMiss) t.refill();
        }
        if (st == kWalkBad) fail(kStInternal);
        ws.aln_len = t.n;
      }
    }
    ex.sync();
  }

  // ---- sequence access ----------------------------------------------------------------------------
  VGC_HD VGC_INL uint32_t weight_at(uint32_t layer, uint32_t pos) const {
    if (!bv.has_qual[layer]) return 1u;
    return bv.wlut[bv.quals[bv.seq_off[layer] + pos]];
  }

  // ---- graph.cpp:182-299, all lanes.  The alignment is in aln_* in reverse order (entry n-1 is the first
  //      pair).  The reference walks the alignment once, serially; the same result is produced here in
  //      parallel because along one alignment
  //        * the aligned sequence positions are the contiguous run [vfront, vback], in increasing order;
  //        * new nodes are numbered prefix chain, suffix chain, then aligned-part nodes in alignment order
  //          (graph.cpp:230-236) — an exclusive scan over "needs a new node";
  //        * a path visits at most one node of an aligned clique, and gives every node at most one new in-edge
  //          and one new out-edge, so clique updates and in/out-list appends of different positions never touch
  //          the same list (list order = creation order is therefore preserved).
  VGC_HD void add_alignment(const uint8_t* codes, uint32_t layer, uint32_t len) {
    Graph& g = G();
    const uint32_t nV0 = g.nV, nE0 = g.nE;
    if (len == 0) return;
    // conservative capacity check: every position may need a node and an edge (the regrow pass sizes slots
    // for the sum of all layer lengths, which always suffices)
    if (nV0 + len > sl.max_nodes || nE0 + len > sl.max_edges) {
      if (ex.leader()) fail(nV0 + len > sl.max_nodes ? kStNodeOverflow : kStEdgeOverflow);
      ex.sync();
      return;
    }
    const uint32_t S = sl.in_stride;
    const uint32_t covinc = len > 1 ? 1u : 0u;  // Node::Coverage counts edge labels: a 1-base sequence has none
    const uint32_t n = ws.aln_len;
    uint32_t* npos = sl.tmp1;  // node id of every sequence position
    const int W = ex.width(), L = ex.lane();
    const bool track = ws.order_ok != 0;  // keep the incremental order's inputs (anchors, dirty blocks) up to date
    uint32_t* anch = sl.anch;
    auto new_node = [&](uint32_t v, uint32_t code) {
      g.code[v] = static_cast<uint8_t>(code);
      g.nal[v] = 0;
      g.nin[v] = 0;
      g.nout[v] = 0;
      g.cov[v] = covinc;
      sl.flags[v] = 0;
      sl.dirty[v] = 0;
    };
    uint32_t newV = 0;
    if (n == 0) {
      for (uint32_t pos = L; pos < len; pos += W) {
        new_node(nV0 + pos, codes[pos]);
        npos[pos] = nV0 + pos;
        anch[pos] = kNone;
      }
      newV = len;
    } else {
      uint32_t mn = kNone, mx = 0;
      for (uint32_t t = L; t < n; t += W) {
        const int32_t p = sl.aln_pos[t];
        if (p != -1) {
          mn = static_cast<uint32_t>(p) < mn ? static_cast<uint32_t>(p) : mn;
          mx = static_cast<uint32_t>(p) > mx ? static_cast<uint32_t>(p) : mx;
        }
      }
      mn = ex.reduce_min(mn);
      mx = ex.reduce_max(mx);
      if (mn == kNone) {
        if (ex.leader()) fail(kStInternal);
        ex.sync();
        return;
      }
      const uint32_t vfront = mn, vback = mx;
      const uint32_t nPre = vfront, nSuf = len - vback - 1;
      for (uint32_t pos = L; pos < vfront; pos += W) {
        new_node(nV0 + pos, codes[pos]);
        npos[pos] = nV0 + pos;
        anch[pos] = kNone;
      }
      for (uint32_t pos = vback + 1 + L; pos < len; pos += W) {
        const uint32_t v = nV0 + nPre + (pos - vback - 1);
        new_node(v, codes[pos]);
        npos[pos] = v;
        anch[pos] = kNone;
      }
      const uint32_t base = nV0 + nPre + nSuf;
      uint32_t carry = 0;
      for (uint32_t f0 = 0; f0 < n; f0 += W) {
        const uint32_t f = f0 + L;
        int32_t pos = -1, nd = -1;
        if (f < n) {
          pos = sl.aln_pos[n - 1 - f];
          nd = sl.aln_node[n - 1 - f];
        }
        uint32_t curr = kNone, code = 0, kind = 0;  // kind: 0 none/existing, 1 new unaligned, 2 new aligned
        if (pos != -1) {
          code = codes[pos];
          if (nd == -1) {
            kind = 1;
          } else {
            const uint32_t jt = static_cast<uint32_t>(nd);
            // code, clique size and the first clique members are asked for together (DRAM latency, see the edge loop)
            const uint32_t cj = g.code[jt];
            const uint32_t na = g.nal[jt];
            uint32_t kt4[4];
#pragma unroll
            for (uint32_t i = 0; i < 4; ++i) kt4[i] = g.al[jt * sl.al_stride + i];
            if (cj == code) {
              curr = jt;
            } else {
              uint32_t ck[4];
#pragma unroll
              for (uint32_t i = 0; i < 4; ++i) ck[i] = i < na ? g.code[kt4[i]] : 0xFFFFFFFFu;
#pragma unroll
              for (uint32_t i = 0; i < 4; ++i) {
                if (curr == kNone && ck[i] == code) curr = kt4[i];
              }
              for (uint32_t i = 4; i < na && curr == kNone; ++i) {
                const uint32_t kt = g.al[jt * sl.al_stride + i];
                if (g.code[kt] == code) curr = kt;
              }
              if (curr == kNone) kind = 2;
            }
          }
        }
        uint32_t tot;
        const uint32_t idx = ex.excl_scan(kind ? 1u : 0u, &tot);
        if (kind) {
          curr = base + carry + idx;
          new_node(curr, code);
          if (kind == 2) {
            const uint32_t jt = static_cast<uint32_t>(nd);
            const uint32_t na = g.nal[jt];
            if (na + 1 > sl.al_stride - 1u) {  // cannot happen: a clique holds distinct codes, al_stride > num_codes - 1
              fail(kStAlignedOverflow);
            } else {
              for (uint32_t i = 0; i < na; ++i) {
                const uint32_t kt = g.al[jt * sl.al_stride + i];
                g.al[kt * sl.al_stride + g.nal[kt]] = curr;
                g.nal[kt] = g.nal[kt] + 1;
                g.al[curr * sl.al_stride + i] = kt;
              }
              g.al[jt * sl.al_stride + na] = curr;
              g.nal[jt] = static_cast<uint8_t>(na + 1);
              g.al[curr * sl.al_stride + na] = jt;
              g.nal[curr] = static_cast<uint8_t>(na + 1);
            }
          }
        } else if (pos != -1) {
          if (covinc) ex.atomic_add(&g.cov[curr], covinc);  // one position per node along an alignment: nobody waits
        }
        if (pos != -1) {
          npos[pos] = curr;
          // anchor of the incremental order: the block owner of the node this position joins (existing node, or the
          // aligned group a new node enters); a new unaligned node has none
          if (track) anch[pos] = kind == 1 ? kNone : sl.owner[kind == 2 ? static_cast<uint32_t>(nd) : curr];
        }
        carry += tot;
      }
      newV = nPre + nSuf + carry;
    }
    ex.sync();
    if (ws.status != kStOk) return;
    // edges between consecutive sequence positions (weight = w[pos-1] + w[pos], graph.cpp:126,290,296)
    uint32_t ecarry = 0;
    for (uint32_t p0 = 1; p0 < len; p0 += W) {
      const uint32_t pos = p0 + L;
      bool need = pos < len;
      uint32_t tail = 0, head = 0, w = 0;
      if (need) {
        tail = npos[pos - 1];
        head = npos[pos];
        w = weight_at(layer, pos - 1) + weight_at(layer, pos);
        if (tail < nV0 && head < nV0) {
          // is there an edge tail -> head already?  The in-list entries are loaded together with the in-degree (the
          // row has room for in_stride >= 8 entries, initialised or not) instead of one dependent load after the
          // other: this loop is DRAM-latency bound.  Positions of one alignment touch distinct edges, so the
          // weight goes in with a plain atomic add that nobody waits for.
          const uint32_t ni = g.nin[head];
          uint32_t tl[8];
#pragma unroll
          for (uint32_t i = 0; i < 8; ++i) tl[i] = g.itail[head * S + i];
          uint32_t hit = kNone;
#pragma unroll
          for (uint32_t i = 0; i < 8; ++i) {
            if (hit == kNone && i < ni && tl[i] == tail) hit = i;
          }
          for (uint32_t i = 8; i < ni && hit == kNone; ++i) {
            if (g.itail[head * S + i] == tail) hit = i;
          }
          if (hit != kNone) {
            ex.atomic_add(&g.ew[g.ieid[head * S + hit]], w);
            need = false;
          }
        }
      }
      uint32_t tot;
      const uint32_t idx = ex.excl_scan(need ? 1u : 0u, &tot);
      if (need) {
        const uint32_t e = nE0 + ecarry + idx;
        g.etail[e] = tail;
        g.ehead[e] = head;
        g.ew[e] = w;
        g.edead[e] = 0;
        const uint32_t slot = g.nin[head];
        if (slot >= S) {
          fail(kStDegreeOverflow);
        } else {
          g.itail[head * S + slot] = tail;
          g.ieid[head * S + slot] = e;
          g.nin[head] = slot + 1;
        }
        g.ein_ord[e] = slot;
        g.eout_ord[e] = g.nout[tail];
        g.nout[tail] = g.nout[tail] + 1;
        // a new edge inside one block changes that block's DFS
        if (track && tail < nV0 && head < nV0 && sl.owner[tail] == sl.owner[head]) sl.dirty[sl.owner[head]] = 1;
      }
      ecarry += tot;
    }
    if (ex.leader()) {
      g.nV = nV0 + newV;
      g.nE = nE0 + ecarry;
    }
    ex.sync();
    if (ws.status != kStOk) return;
    if (nV0 == 0 && n == 0) order_init_chain(len);
    else order_update(nV0, len, npos);
#ifdef VGC_CHECK_ORDER
    order_check();
#endif
  }

#ifdef VGC_CHECK_ORDER
  // test hook (host model): the incremental order must equal the full DFS's, node for node
  void order_check() {
    if (!ws.order_ok) {
      g_order_checks[1] += 1;
      return;
    }
    Graph& g = G();
    const uint32_t nV = g.nV;
    std::vector<uint32_t> inc(sl.r2n, sl.r2n + nV), inc_owner(sl.owner, sl.owner + nV), inc_rank(sl.rank_of, sl.rank_of + nV);
    std::vector<uint32_t> inc_bsize(sl.bsize, sl.bsize + nV), inc_bstart(sl.bstart, sl.bstart + nV);
    std::vector<uint32_t> full(nV);
    const uint32_t sorts = ws.sorts, hbm = ws.sorts_hbm, nm = ws.nMain;
    uint32_t* keep = sl.r2n;
    sl.r2n = full.data();
    const uint32_t n = sort_graph(false, 0, 0, full.data());
    sl.r2n = keep;
    ws.sorts = sorts;
    ws.sorts_hbm = hbm;
    ws.nMain = nm;
    bool same = n == nV;
    for (uint32_t i = 0; same && i < nV; ++i) same = full[i] == inc[i] && sl.owner[i] == inc_owner[i] && sl.rank_of[i] == inc_rank[i] && sl.bsize[i] == inc_bsize[i] && (inc_bsize[i] == 0 || sl.bstart[i] == inc_bstart[i]);
    g_order_checks[same ? 0 : 2] += 1;
    std::copy(inc.begin(), inc.end(), sl.r2n);
  }
#endif

  // ---------------------------------------------------------------------------------------------------------
  // Incremental Graph::TopologicalSort (graph.cpp:301-371).
  //
  // The reference's DFS takes the roots in node-id order and, from each root that is not finished yet, emits the
  // root's unfinished ancestors (in-edge tails and aligned nodes, transitively) in post-order, then the root.  The
  // order is therefore a sequence of BLOCKS, one per root r, in increasing r, and node u belongs to the block of
  //     owner(u) = the smallest id among the nodes reachable from u along out-edges and aligned links (u included),
  // because that is the first root whose DFS reaches u.  Inside DFS(r) every node with a smaller owner is finished
  // and no node with a larger owner is reachable, so the internal order of block r depends only on the in-lists and
  // aligned lists of its own members.
  //
  // AddAlignment adds nodes with ids above every old id, and edges / aligned links along one path.  If the owners
  // of the old nodes on the path (for a new aligned node: of the group it joins) never decrease along the path —
  // always true when the alignment was computed on this very order, because the DP visits rows in rank order and
  // blocks are contiguous in rank — then no old owner changes (an old node only gains successors whose owner is
  // not smaller than its own), a new node's owner is the owner of the next such node on the path (its own id if
  // there is none: the unaligned tail of the read), and only blocks that gained a member or an inner edge need their
  // DFS again.  Everything else keeps its internal order and just moves by the growth of the blocks before it.
  // Anything else (an alignment on a Subgraph view may break the monotonicity; dirty blocks that do not fit the
  // executor's fast storage) clears ws.order_ok and the full sort runs instead.
  VGC_HD VGC_INL static uint32_t ctz32(uint32_t m) {
    uint32_t c = 0;
    while (!((m >> c) & 1u)) ++c;
    return c;
  }

  VGC_HD void order_init_chain(uint32_t len) {
    for (uint32_t u = ex.lane(); u < len; u += ex.width()) {
      sl.r2n[u] = u;
      sl.rank_of[u] = u;
      sl.owner[u] = u;
      sl.bsize[u] = 1;
      sl.bstart[u] = u;
      sl.dirty[u] = 0;
    }
    if (ex.leader()) {
      ws.order_ok = 1;
      ws.nMain = len;
    }
    ex.sync();
  }

  // DFS of one dirty block over its staged copy (order_update step 5): graph.cpp:312-368 restricted to the nodes whose
  // owner is the block's root.  Nodes carry compact ids; one 32-bit record per node: adjacency offset (bits 0-15) |
  // in-block in-degree (16-21) | aligned count (22-25) | expanded (26) | done (27) | ignored (28).  Same single-scan
  // scheme as toposort_fast.  Lane-private: the stack segment holds 1 + the block's adjacency entries, which bounds
  // the pushes (a node is expanded once).
  static constexpr uint32_t kBExpanded = 1u << 26, kBDone = 1u << 27, kBIgnored = 1u << 28;
  VGC_HD uint32_t block_dfs(uint32_t* rec, const uint16_t* adj, const uint16_t* node, uint16_t* stack, uint32_t cap,
                            uint32_t root, uint32_t start, uint32_t expect, uint32_t* out) {
    uint32_t n = 0, sp = 1, curr = root, r = rec[root];
    stack[0] = static_cast<uint16_t>(root);
    while (true) {
      bool pop = (r & kBDone) != 0;
      if (!pop) {
        const uint32_t off = r & 0xFFFFu, nin = (r >> 16) & 63u, nal = (r >> 22) & 15u;
        const bool primary = !(r & kBIgnored);
        uint32_t last = kNone, last_r = 0;
        if (!(r & kBExpanded)) {
          if (sp + nin + nal > cap) return kNone;
          for (uint32_t i = 0; i < nin; ++i) {
            const uint32_t t = adj[off + i];
            const uint32_t rt = rec[t];
            if (!(rt & kBDone)) {
              stack[sp++] = static_cast<uint16_t>(t);
              last = t;
              last_r = rt;
            }
          }
          if (primary) {
            for (uint32_t i = 0; i < nal; ++i) {
              const uint32_t a = adj[off + nin + i];
              const uint32_t ra = rec[a];
              if (!(ra & kBDone)) {
                stack[sp++] = static_cast<uint16_t>(a);
                rec[a] = ra | kBIgnored;
                last = a;
                last_r = ra | kBIgnored;
              }
            }
          }
        }
        if (last == kNone) {
          rec[curr] = r | kBDone;
          if (primary) {
            if (n + 1 + nal > expect) return kNone;
            const uint32_t v = node[curr];
            out[start + n] = v;
            sl.rank_of[v] = start + n;
            ++n;
            for (uint32_t i = 0; i < nal; ++i) {
              const uint32_t va = node[adj[off + nin + i]];
              out[start + n] = va;
              sl.rank_of[va] = start + n;
              ++n;
            }
          }
          pop = true;
        } else {
          rec[curr] = r | kBExpanded;
          curr = last;
          r = last_r;
        }
      }
      if (pop) {
        if (--sp == 0) break;
        curr = stack[sp - 1];
        r = rec[curr];
      }
    }
    return n;
  }

  static constexpr uint32_t kBDirty = 0x80000000u;  // bstart2[r]: the block of root r is re-sorted by this update

  VGC_HD void order_update(uint32_t nV0, uint32_t len, const uint32_t* npos) {
    if (!ws.order_ok) return;
    Graph& g = G();
    const uint32_t nV = g.nV;
    const int W = ex.width(), L = ex.lane();
    uint32_t* anch = sl.anch;
    bool bad = ws.nMain != nV0 || nV >= 65535u;
    // 1. owner of every new unaligned node = the next anchor's along the path (kNone: its own id); anchors must not
    //    decrease.  Chunks of W positions, last chunk first; `carry` = owner of the first anchor behind the chunk.
    uint32_t carry = kNone;
    for (uint32_t c0 = ((len - 1) / W) * W;; c0 -= W) {
      const uint32_t pos = c0 + L;
      const bool valid = pos < len;
      const uint32_t a = valid ? anch[pos] : kNone;
      const bool is_anchor = valid && a != kNone;
      const uint32_t mask = ex.ballot(is_anchor);
      const uint32_t above = L + 1 < 32 ? (mask >> (L + 1)) : 0u;
      uint32_t nxt = ex.bcast(a, above ? L + 1 + ctz32(above) : 0u);
      if (!above) nxt = carry;
      if (valid) {
        if (is_anchor) {
          if (nxt != kNone && a > nxt) bad = true;
        } else {
          anch[pos] = nxt;
        }
      }
      if (mask) carry = ex.bcast(a, ctz32(mask));
      if (c0 == 0) break;
    }
    if (ex.reduce_max(bad ? 1u : 0u)) {
      if (ex.leader()) ws.order_ok = 0;
      ex.sync();
      return;
    }
    // 2. new nodes: owner, block sizes, dirty blocks
    for (uint32_t v = nV0 + L; v < nV; v += W) sl.bsize[v] = 0;
    ex.sync();
    for (uint32_t pos = L; pos < len; pos += W) {
      const uint32_t v = npos[pos];
      if (v < nV0) continue;
      const uint32_t r = anch[pos] == kNone ? v : anch[pos];
      sl.owner[v] = r;
      ex.atomic_add(&sl.bsize[r], 1u);
      sl.dirty[r] = 1;
    }
    ex.sync();
    // 3. new block starts (exclusive scan of the sizes over the roots; kBDirty marks the blocks to re-sort) and the
    //    list of dirty roots.  `need[r]` (tmp1: npos is dead by now) will count the block's staged adjacency entries.
    //    Four roots per lane and turn so that the loads of a turn are in flight together.
    uint32_t* list = sl.tmp0;
    uint32_t* need = sl.tmp1;
    uint32_t nd = 0, total = 0;
    constexpr int kU = 4;
    for (uint32_t base = 0; base < nV; base += kU * W) {
      uint32_t c[kU], f[kU];
#pragma unroll
      for (int t = 0; t < kU; ++t) {
        const uint32_t r = base + t * W + L;
        c[t] = r < nV ? sl.bsize[r] : 0u;
        f[t] = r < nV ? sl.dirty[r] : 0u;
      }
#pragma unroll
      for (int t = 0; t < kU; ++t) {
        const uint32_t r = base + t * W + L;
        uint32_t tot;
        const uint32_t q = ex.excl_scan(c[t], &tot);
        if (r < nV) sl.bstart2[r] = (total + q) | (f[t] ? kBDirty : 0u);
        total += tot;
        const uint32_t k = ex.excl_scan(f[t] ? 1u : 0u, &tot);
        if (f[t]) {
          list[nd + k] = r;
          need[r] = 1;
          sl.dirty[r] = 0;
        }
        nd += tot;
      }
    }
    ex.sync();
    // 4. clean blocks keep their internal order and move; the members of dirty blocks (every new node is one) get
    //    compact numbers, kept in rank_of until their DFS writes the real rank
    uint32_t* memb = sl.order;  // compact id -> node (the Subgraph order is rebuilt before every use)
    uint32_t M = 0;
    for (uint32_t base = 0; base < nV; base += kU * W) {
      uint32_t q[kU], b1[kU], b2[kU];
#pragma unroll
      for (int t = 0; t < kU; ++t) {
        const uint32_t u = base + t * W + L;
        const uint32_t r = u < nV ? sl.owner[u] : 0u;
        q[t] = u < nV0 ? sl.rank_of[u] : 0u;
        b2[t] = u < nV ? sl.bstart2[r] : 0u;
        b1[t] = u < nV0 ? sl.bstart[r] : 0u;
      }
#pragma unroll
      for (int t = 0; t < kU; ++t) {
        const uint32_t u = base + t * W + L;
        const bool is_m = u < nV && (b2[t] & kBDirty) != 0;
        uint32_t tot;
        const uint32_t k = ex.excl_scan(is_m ? 1u : 0u, &tot);
        if (is_m) {
          memb[M + k] = u;
          sl.rank_of[u] = M + k;
        } else if (u < nV) {
          const uint32_t nr = b2[t] + (q[t] - b1[t]);
          sl.r2n2[nr] = u;
          sl.rank_of[u] = nr;
        }
        M += tot;
      }
    }
    ex.sync();
    // 5. dirty blocks: staged in the executor's fast storage as  rec32[M] | node16[M] | adj16[A] | stacks16[A + nd]
    //    and re-sorted there, one block per lane at a time
    uint32_t* ab;
    uint32_t abytes;
    ex.block_arena(&ab, &abytes);
    const uint32_t S = sl.in_stride;
    bool ok = total == nV && (static_cast<uint64_t>(M) * 6u + 8u) <= abytes;
    uint32_t* rec = ab;
    uint16_t* node = reinterpret_cast<uint16_t*>(ab + M);
    uint16_t* adj = node + ((M + 1u) & ~1u);
    uint32_t A = 0;
    if (ok) {
      // 5a. records: in-block in-degree + aligned count, adjacency offsets
      for (uint32_t base = 0; base < M; base += W) {
        const uint32_t c = base + L;
        uint32_t cin = 0, nal = 0, r = 0, u = 0;
        if (c < M) {
          u = memb[c];
          r = sl.owner[u];
          const uint32_t nin = g.nin[u];
          nal = g.nal[u];
          for (uint32_t i = 0; i < nin; ++i) cin += sl.owner[g.itail[u * S + i]] == r ? 1u : 0u;
          if (cin > 63u || nal > 15u) ok = false;
        }
        uint32_t tot;
        const uint32_t off = ex.excl_scan(cin + nal, &tot);
        if (c < M) {
          rec[c] = ((A + off) & 0xFFFFu) | ((cin & 63u) << 16) | ((nal & 15u) << 22);
          node[c] = static_cast<uint16_t>(u);
          ex.atomic_add(&need[r], cin + nal);
        }
        A += tot;
      }
      if (A >= 65535u || static_cast<uint64_t>(M) * 6u + 8u + 2ull * A + 2ull * (A + nd) > abytes) ok = false;
    }
    ok = ex.reduce_max(ok ? 0u : 1u) == 0;
    ex.sync();
    if (ok) {
      // 5b. adjacency (compact ids): in-block tails in in-list order, then the aligned nodes; stack segments
      for (uint32_t c = L; c < M; c += W) {
        const uint32_t u = memb[c];
        const uint32_t r = sl.owner[u];
        const uint32_t nin = g.nin[u], nal = g.nal[u];
        uint32_t k = rec[c] & 0xFFFFu;
        for (uint32_t i = 0; i < nin; ++i) {
          const uint32_t t = g.itail[u * S + i];
          if (sl.owner[t] == r) adj[k++] = static_cast<uint16_t>(sl.rank_of[t]);
        }
        for (uint32_t i = 0; i < nal; ++i) adj[k++] = static_cast<uint16_t>(sl.rank_of[g.al[u * sl.al_stride + i]]);
      }
      uint16_t* stacks = adj + A;
      uint32_t scarry = 0;
      for (uint32_t base = 0; base < nd; base += W) {
        const uint32_t b = base + L;
        const uint32_t r = b < nd ? list[b] : 0u;
        const uint32_t nb = b < nd ? need[r] : 0u;
        uint32_t tot;
        const uint32_t so = ex.excl_scan(nb, &tot);
        if (b < nd) need[r] = scarry + so;
        scarry += tot;
      }
      ex.sync();
      // 5c. the DFS of each dirty block
      for (uint32_t b = L; b < nd; b += W) {
        const uint32_t r = list[b];
        const uint32_t so = need[r];
        const uint32_t start = sl.bstart2[r] & ~kBDirty, expect = sl.bsize[r];
        // (the stack segment is large enough by construction: see block_dfs)
        const uint32_t n = block_dfs(rec, adj, node, stacks + so, 0xFFFFFFFFu, sl.rank_of[r], start, expect, sl.r2n2);
        if (n != expect) ok = false;
        sl.bstart2[r] = start;
      }
    } else {
      for (uint32_t b = L; b < nd; b += W) sl.bstart2[list[b]] &= ~kBDirty;
    }
    const bool all_ok = ex.reduce_max(ok ? 0u : 1u) == 0;
    ex.sync();
    if (ex.leader()) {
      if (all_ok) {
        uint32_t* t0 = sl.r2n;
        sl.r2n = sl.r2n2;
        sl.r2n2 = t0;
        uint32_t* t1 = sl.bstart;
        sl.bstart = sl.bstart2;
        sl.bstart2 = t1;
        ws.nMain = nV;
      } else {
        ws.order_ok = 0;
      }
    }
    ex.sync();
  }

  // ---- graph.cpp:1104-1165: parallel over alignment entries; weight adds commute --------------------
  VGC_HD void add_weights(uint32_t layer) {
    Graph& g = G();
    const uint32_t n = ws.aln_len;
    // entry t (reverse order) is preceded in sequence order by entry t + 1
    for (uint32_t t = ex.lane(); t + 1 < n; t += ex.width()) {
      const int32_t nd = sl.aln_node[t], pos = sl.aln_pos[t];
      const int32_t pnd = sl.aln_node[t + 1], ppos = sl.aln_pos[t + 1];
      if (nd == -1 || pos == -1 || pnd == -1 || ppos == -1) continue;
      const uint32_t w = weight_at(layer, pos - 1) + weight_at(layer, pos);
      bool hit = false;
      const uint32_t S = sl.in_stride;
      for (uint32_t i = 0; i < g.nin[nd]; ++i) {
        if (g.itail[nd * S + i] == static_cast<uint32_t>(pnd)) {
          ex.atomic_add(&g.ew[g.ieid[nd * S + i]], w);
          hit = true;
          break;
        }
      }
      if (!hit) fail(kStInternal);  // consecutive matched pairs always follow an edge of the graph
    }
    ex.sync();
  }

  // ---- graph.cpp:811-982: parallel over edges; fp64 divisions and >= compares exactly as written ----
  VGC_HD void prune(double min_confidence, double min_support, double average_weight) {
    Graph& g = G();
    const uint32_t nV = g.nV, nE = g.nE;
    for (uint32_t v = ex.lane(); v < nV; v += ex.width()) {
      sl.tmp0[v] = 0;
      sl.tmp1[v] = 0;
    }
    ex.sync();
    for (uint32_t e = ex.lane(); e < nE; e += ex.width()) {
      ex.atomic_add(&sl.tmp0[g.etail[e]], g.ew[e]);  // sum over tail's out-edges
      ex.atomic_add(&sl.tmp1[g.ehead[e]], g.ew[e]);  // sum over head's in-edges
    }
    ex.sync();
    for (uint32_t e = ex.lane(); e < nE; e += ex.width()) {
      const double w = static_cast<double>(static_cast<int64_t>(g.ew[e]));
      const double cuv = w / static_cast<double>(static_cast<int64_t>(sl.tmp0[g.etail[e]]));
      const double sup = w / average_weight;
      const double cvu = w / static_cast<double>(static_cast<int64_t>(sl.tmp1[g.ehead[e]]));
      const bool keep = (cuv >= min_confidence) && (cvu >= min_confidence) && (sup >= min_support);
      g.edead[e] = keep ? 0 : 1;
    }
    ex.sync();
  }

  // ---- graph.cpp:984-1089.  Components of the pruned graph by the reference's recursive pre-order (DfsUtil:
  //      a node's neighbours are its live in-edge tails in in-list order, then its live out-edge heads in
  //      out-list order; `visited` is tested when the loop reaches a neighbour); the LAST largest component wins
  //      (`>=`, :1049); the graph is rebuilt with ids = pre-order positions, edges in (new tail id, out-list
  //      position) order with weight 0, and no aligned links.
  //      Parallel: live adjacency (staged in the executor's fast storage when it fits) and the whole rebuild.
  //      Serial (leader): only the pre-order walk itself, over the compact adjacency.
  template <class OffT, class AdjT, class StkT>
  VGC_HD bool components(const OffT* off, const AdjT* adj, uint8_t* visited, StkT* stk, uint32_t cap, uint32_t* comp,
                         uint32_t* best_start_out, uint32_t* best_size_out) {
    const uint32_t nV = G().nV;
    uint32_t ncomp = 0, best_start = 0, best_size = 0;
    for (uint32_t v0 = 0; v0 < nV; ++v0) {
      if (visited[v0]) continue;
      const uint32_t start = ncomp;
      visited[v0] = 1;
      comp[ncomp++] = v0;
      if (off[v0] != off[v0 + 1]) {
        uint32_t sp = 0;
        stk[sp++] = static_cast<StkT>(v0);
        while (sp > 0) {
          // re-scan from the first neighbour: everything before the one taken last time is visited by now
          const uint32_t v = stk[sp - 1];
          uint32_t next = kNone;
#pragma unroll 1
          for (uint32_t c = off[v], e = off[v + 1]; c < e; ++c) {
            const uint32_t u = adj[c];
            if (!visited[u]) {
              next = u;
              break;
            }
          }
          if (next == kNone) {
            --sp;
          } else {
            if (sp >= cap) return false;
            visited[next] = 1;
            comp[ncomp++] = next;
            stk[sp++] = static_cast<StkT>(next);
          }
        }
      }
      if (ncomp - start >= best_size) {
        best_size = ncomp - start;
        best_start = start;
      }
    }
    *best_start_out = best_start;
    *best_size_out = best_size;
    return true;
  }

  VGC_HD void largest_subgraph() {
    build_out_csr();
    Graph& g = G();
    const uint32_t S = sl.in_stride;
    Graph& h = sl.g[ws.cur ^ 1];
    const uint32_t nV = g.nV;
    const int W = ex.width(), L = ex.lane();
    uint32_t* comp = sl.tmp0;   // all components back to back
    uint32_t* newid = sl.tmp1;
    uint32_t* goff = sl.H;      // [nV + 1] offsets of the live adjacency
    // live degree and offsets
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nV; base += W) {
      const uint32_t v = base + L;
      uint32_t c = 0;
      if (v < nV) {
        for (uint32_t i = 0; i < g.nin[v]; ++i) c += g.edead[g.ieid[v * S + i]] ? 0u : 1u;
        for (uint32_t k = sl.out_off[v]; k < sl.out_off[v + 1]; ++k) c += g.edead[sl.out_eid[k]] ? 0u : 1u;
      }
      uint32_t tot;
      const uint32_t q = ex.excl_scan(c, &tot);
      if (v < nV) goff[v] = carry + q;
      carry += tot;
    }
    const uint32_t nA = carry;
    if (ex.leader()) goff[nV] = nA;
    ex.sync();
    uint16_t *off16, *adj16, *stk16;
    uint8_t* vis;
    uint32_t cap16;
    const bool fast = ex.stage_lsg(nV, nA, &off16, &adj16, &vis, &stk16, &cap16);
    uint32_t* gadj = sl.H + nV + 1;
    uint32_t* gstk = gadj + nA;
    if (!fast) vis = sl.flags;
    for (uint32_t v = L; v < nV; v += W) {
      uint32_t o = goff[v];
      for (uint32_t i = 0; i < g.nin[v]; ++i) {
        const uint32_t e = g.ieid[v * S + i];
        if (g.edead[e]) continue;
        if (fast) adj16[o] = static_cast<uint16_t>(g.etail[e]);
        else gadj[o] = g.etail[e];
        ++o;
      }
      for (uint32_t k = sl.out_off[v]; k < sl.out_off[v + 1]; ++k) {
        const uint32_t e = sl.out_eid[k];
        if (g.edead[e]) continue;
        if (fast) adj16[o] = static_cast<uint16_t>(g.ehead[e]);
        else gadj[o] = g.ehead[e];
        ++o;
      }
      vis[v] = 0;
      if (fast) off16[v] = static_cast<uint16_t>(goff[v]);
    }
    if (fast && L == 0) off16[nV] = static_cast<uint16_t>(nA);
    ex.sync();
    if (ex.leader()) {
      uint32_t bs = 0, bn = 0;
      bool ok = false;
      if (fast) {
        ok = components<uint16_t, uint16_t, uint16_t>(off16, adj16, vis, stk16, cap16, comp, &bs, &bn);
        if (!ok) {
          for (uint32_t v = 0; v < nV; ++v) vis[v] = 0;  // deep recursion: redo with the big stack in HBM
          ok = components<uint16_t, uint16_t, uint32_t>(off16, adj16, vis, gstk, 0xFFFFFFFFu, comp, &bs, &bn);
        }
      } else {
        ok = components<uint32_t, uint32_t, uint32_t>(goff, gadj, vis, gstk, 0xFFFFFFFFu, comp, &bs, &bn);
      }
      ws.scratch[0] = bs;
      ws.scratch[1] = bn;
    }
    ex.sync();
    const uint32_t best_start = ws.scratch[0], best_size = ws.scratch[1];
    // ---- rebuild (all lanes).  Node i of the new graph = comp[best_start + i].
    uint32_t* eoff = sl.rank_of;  // first new edge id of each new node (scratch: the next sort rewrites rank_of)
    uint32_t ecarry = 0;
    for (uint32_t base = 0; base < best_size; base += W) {
      const uint32_t i = base + L;
      uint32_t lout = 0;
      if (i < best_size) {
        const uint32_t v = comp[best_start + i];
        newid[v] = i;
        uint32_t lin = 0;
        for (uint32_t q = 0; q < g.nin[v]; ++q) lin += g.edead[g.ieid[v * S + q]] ? 0u : 1u;
        for (uint32_t k = sl.out_off[v]; k < sl.out_off[v + 1]; ++k) lout += g.edead[sl.out_eid[k]] ? 0u : 1u;
        h.code[i] = g.code[v];
        h.nal[i] = 0;
        h.cov[i] = 0;
        h.nin[i] = lin;
        h.nout[i] = lout;
      }
      uint32_t tot;
      const uint32_t q = ex.excl_scan(lout, &tot);
      if (i < best_size) eoff[i] = ecarry + q;
      ecarry += tot;
    }
    ex.sync();
    for (uint32_t i = L; i < best_size; i += W) {
      const uint32_t v = comp[best_start + i];
      uint32_t t = 0;
      for (uint32_t k = sl.out_off[v]; k < sl.out_off[v + 1]; ++k) {
        const uint32_t e = sl.out_eid[k];
        if (g.edead[e]) continue;
        const uint32_t ne = eoff[i] + t;
        const uint32_t ho = g.ehead[e];
        const uint32_t hd = newid[ho];
        // position in the head's new in-list: edges are created in increasing new tail id
        uint32_t slot = 0;
        for (uint32_t q = 0; q < g.nin[ho]; ++q) {
          const uint32_t e2 = g.ieid[ho * S + q];
          if (!g.edead[e2] && newid[g.etail[e2]] < i) ++slot;
        }
        h.etail[ne] = i;
        h.ehead[ne] = hd;
        h.ew[ne] = 0;
        h.edead[ne] = 0;
        h.itail[hd * S + slot] = i;
        h.ieid[hd * S + slot] = ne;
        h.ein_ord[ne] = slot;
        h.eout_ord[ne] = t;
        ++t;
      }
    }
    if (ex.leader()) {
      h.nV = best_size;
      h.nE = ecarry;
      ws.cur ^= 1;
    }
    ex.sync();
  }

  // ---- stage a layer's codes into the executor's fast storage ------------------------------------
  VGC_HD uint8_t* stage_codes(uint32_t layer) {
    const uint32_t len = static_cast<uint32_t>(bv.seq_off[layer + 1] - bv.seq_off[layer]);
    uint8_t* codes = ex.seq_codes();
    for (uint32_t i = ex.lane(); i < len; i += ex.width()) codes[i] = static_cast<uint8_t>(base_code(bv, bv.seq_off[layer] + i));
    ex.sync();
    return codes;
  }

  // ---- graph.cpp:534-638 + 450-485 (leader): heaviest bundle consensus + coverage, linear mode ------
  VGC_HD uint32_t heaviest_bundle(uint32_t nR, uint8_t* out, uint32_t out_cap, uint32_t* cov_out) {
    Graph& g = G();
    const uint32_t S = sl.in_stride;
    const uint32_t nV = g.nV;
    // scores (int64) and predecessors live in the H scratch
    long long* score = reinterpret_cast<long long*>(sl.H);
    uint32_t* pred = reinterpret_cast<uint32_t*>(score + nV);
    uint32_t* n2r = pred + nV;
    for (uint32_t v = 0; v < nV; ++v) {
      score[v] = -1;
      pred[v] = kNone;
    }
    for (uint32_t r = 0; r < nR; ++r) n2r[sl.r2n[r]] = r;
    uint32_t mx = kNone;
    auto relax = [&](uint32_t it, bool skip_dead_tails) {
      for (uint32_t i = it * S; i < it * S + g.nin[it]; ++i) {
        const uint32_t t = g.itail[i];
        const long long w = g.ew[g.ieid[i]];
        if (skip_dead_tails && score[t] == -1) continue;
        if (score[it] < w || (score[it] == w && score[pred[it]] <= score[t])) {
          score[it] = w;
          pred[it] = t;
        }
      }
      if (pred[it] != kNone) score[it] += score[pred[it]];
      if (mx == kNone || score[mx] < score[it]) mx = it;
    };
    for (uint32_t r = 0; r < nR; ++r) relax(sl.r2n[r], false);
    while (g.nout[mx] != 0) {
      // BranchCompletion (graph.cpp:590-638)
      const uint32_t start = mx, rank = n2r[mx];
      for (uint32_t k = sl.out_off[start]; k < sl.out_off[start + 1]; ++k) {
        const uint32_t hd = g.ehead[sl.out_eid[k]];
        for (uint32_t i = hd * S; i < hd * S + g.nin[hd]; ++i) {
          if (g.itail[i] != start) score[g.itail[i]] = -1;
        }
      }
      mx = kNone;
      for (uint32_t r = rank + 1; r < nR; ++r) {
        const uint32_t it = sl.r2n[r];
        score[it] = -1;
        pred[it] = kNone;
        relax(it, true);
      }
    }
    // traceback into out (reverse, then flip)
    uint32_t n = 0;
    uint32_t* path = n2r + nV;
    while (true) {
      path[n++] = mx;
      if (pred[mx] == kNone) break;
      mx = pred[mx];
    }
    if (n > out_cap) {
      fail(kStOutOverflow);
      return 0;
    }
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t v = path[n - 1 - i];
      out[i] = bv.decoder[g.code[v]];
      uint32_t c = g.cov[v];
      for (uint32_t a = 0; a < g.nal[v]; ++a) c += g.cov[g.al[v * sl.al_stride + a]];
      cov_out[i] = c;
    }
    return n;
  }

  // ---- the window (src/window.cpp:74-174 and :176-428) as a resumable program -------------------------
  // A window cycles through four steps, each of which the device runs in its own kernel with the occupancy
  // that suits it (vgc_engine.cu); ws.need names the step the window is waiting for:
  //   step_update  (kNeedUpdate) : consume the alignment just traced (AddAlignment / AddWeights / emit), run
  //                                the phase transitions (prune, LargestSubgraph, consensus), choose the next
  //                                alignment and record what must be prepared for it in ws.prep
  //   step_prepare (kNeedPrepare): re-sort the graph (serial DFS), Subgraph view of a partial layer, row program
  //   fill         (kNeedFill)   : the DP fill (poa_fill.cuh on the device, HostEx::fill in the host model)
  //   step_trace   (kNeedTrace)  : traceback of the filled matrix
  // Fills of a window: haplotype (nseq-1) + (num_prune-1)*nseq + 1, linear nseq-1 (known on the host).
  VGC_HD VGC_INL uint32_t layer_len(uint32_t l) const {
    return static_cast<uint32_t>(bv.seq_off[l + 1] - bv.seq_off[l]);
  }

  VGC_HD void finish() {
    if (ex.leader()) {
      ws.pc = kPcDone;
      ws.need = kNeedNone;
    }
    ex.sync();
  }

  VGC_HD void step_trace() {
    if (ws.need != kNeedTrace) return;
    ex.sync();
    if (ex.leader()) ws.t_last = ex.clock();
    traceback(ws.fill_layer, ws.fill_mode);
    tick(kPhTrace);
    if (ws.status != kStOk) return finish();
    if (ex.leader()) ws.need = kNeedUpdate;
    ex.sync();
  }

  VGC_HD void step_prepare(uint32_t win) {
    if (ws.need != kNeedPrepare) return;
    ex.sync();
    if (ex.leader()) ws.t_last = ex.clock();
    const uint32_t prep = ws.prep;
    uint32_t nMain = ws.nMain;
    staged = false;
    ranks_ok = ws.order_ok != 0 && !(prep & (kPrepMainSort | kPrepLargest));  // rank_of describes r2n already
    if (prep & kPrepLargest) {
      largest_subgraph();
      tick(kPhLargest);
    }
    if (prep & kPrepMainSort) {
      nMain = sort_graph(false, 0, 0, sl.r2n);
      tick(kPhSort);
    }
    const uint32_t l = ws.fill_layer;
    const uint32_t* order = sl.r2n;
    uint32_t nR = nMain;
    const bool sub = (prep & kPrepSubSort) != 0;
    if (sub) {
      nR = sort_graph(true, bv.begin[l], bv.end[l], sl.order);
      order = sl.order;
      tick(kPhSort);
    }
    if (prep & kPrepRowprog) {
      build_rowprog(order, nR, sub);
      tick(kPhRowprog);
    }
    if (prep & kPrepFill) {
      // int16 range guard: every cell of the matrix (padding columns included) lies within
      // +-(rows + columns + 2) * max|score|; the columns are those of the row width the fill will pick (fill_width).
      // A round runs NW and SW alignments of every sequence of the window: both score sets, the longest layer.
      const uint32_t mode = ws.fill_mode;
      const bool round = ws.round != 0;
      const uint32_t first = bv.win_first[win], nseq = bv.win_nseq[win];
      const uint32_t* rank = bv.layer_rank + first;
      auto maxabs = [](const Scores& sc) {
        int32_t a = sc.m > -sc.x ? sc.m : -sc.x;
        a = a > -sc.m ? a : -sc.m;
        a = a > sc.x ? a : sc.x;
        return -sc.g > a ? -sc.g : a;
      };
      int32_t a = maxabs(mode == kModeNW ? nw : sw);
      uint32_t maxlen = layer_len(l);
      unsigned long long cells = static_cast<unsigned long long>(ws.nR + 1) * layer_len(l);  // (R_a + 1) * L_a
      uint32_t naln = 1;
      if (round) {
        const int32_t b = maxabs(sw), c = maxabs(nw);
        a = b > c ? b : c;
        cells = 0;
        uint32_t ml = 0;
        for (uint32_t j = ex.lane(); j < nseq; j += ex.width()) {
          const uint32_t len = layer_len(rank[j]);
          ml = len > ml ? len : ml;
          cells += static_cast<unsigned long long>(ws.nR + 1) * len;
        }
        maxlen = ex.reduce_max(ml);
        cells = ex.reduce_add64(cells);
        naln = nseq;
        // the round's alignments add their weights here (index rank * kInlinePreds + in-edge slot)
        for (uint32_t i = ex.lane(); i < ws.nR * kInlinePreds; i += ex.width()) sl.wacc[i] = 0;
      }
      if (ex.leader()) {
        // (the row scan works on H - g * column, which is bounded by 2 * columns * max|score| from above and by
        //  -rows * max|score| from below)
        const int64_t cols = 64ll * fill_width(K, maxlen);
        const int64_t rows = static_cast<int64_t>(ws.nR);
        // beyond the int16 range, or a layer longer than the widest fast row: the int32 kernel takes the step (the
        // reference switches its SIMD lanes to int32 by a bound of the same kind, simd_..._implementation.hpp:699-706)
        ws.wide = (maxlen > static_cast<uint32_t>(RM::kCols) || ((rows > cols ? rows : cols) + cols + 2) * a > 32000) ? 1u : 0u;
        // int32 itself: |score| <= 127 (int8 in the reference), rows + columns < 2^24
        if ((rows + static_cast<int64_t>(maxlen) + 2) * a > 2000000000ll) fail(kStScoreRange);
        ws.alignments += naln;
        ws.cells += cells;
        ws.jobs_total = naln;
        ws.jobs_done = 0;
      }
      ex.sync();
      if (ws.status != kStOk) return finish();
    }
    ex.sync();  // every lane has read ws.nMain / ws.prep (racecheck: the short path above has no other barrier)
    if (ex.leader()) {
      ws.nMain = nMain;
      ws.need = (prep & kPrepFill) ? kNeedFill : kNeedUpdate;
    }
    ex.sync();
  }

  VGC_HD void step_update(uint32_t w, bool haplotype, bool trim, double min_confidence, double min_support,
                          uint32_t num_prune, uint8_t* out, uint32_t* out_len) {
    if (ws.need != kNeedUpdate) return;
    const uint32_t first = bv.win_first[w];
    const uint32_t nseq = bv.win_nseq[w];
    const uint32_t* rank = bv.layer_rank + first;
    const uint32_t bb = rank[0];
    const uint32_t blen = layer_len(bb);
    const uint32_t offset = static_cast<uint32_t>(0.01 * blen);
    const uint32_t out_cap = bv.out_cap[w];
    const double avgw = bv.win_avgw[w];
    const uint32_t pc = ws.pc;
    enum Act { kBuildNext, kRoundStart, kRoundEnd, kFinalSchedule };
    Act act = kBuildNext;
    uint32_t j = ws.j, k = ws.k;
    bool changed = false;  // graph changed since the last sort
    bool largest = false;  // PruneGraph ran: LargestSubgraph is pending
    ex.sync();
    if (ex.leader()) ws.t_last = ex.clock();
    // plan: make the next alignment (or a whole re-alignment round, or a sort-only step) pending.  Every fill is
    // preceded by a prepare step (sort / row program), which also does the per-alignment accounting.
    auto plan = [&](uint32_t next_pc, uint32_t prep, uint32_t layer, uint32_t mode, bool round) {
      if (ex.leader()) {
        ws.pc = next_pc;
        ws.j = j;
        ws.k = k;
        // an alignment against a Subgraph view sorts that view itself and the graph changes again before anything
        // could use the main order, so the main sort is skipped then (the reference sorts both, with the same result)
        if (largest) ws.order_ok = 0;  // LargestSubgraph rebuilds the graph: full sort
        ws.prep = prep | ((changed && !ws.order_ok && !(prep & kPrepSubSort)) ? kPrepMainSort : 0u) | (largest ? kPrepLargest : 0u);
        ws.fill_layer = layer;
        ws.fill_mode = mode;
        ws.round = round ? 1u : 0u;
        ws.need = kNeedPrepare;
      }
      ex.sync();
    };

    if (pc == kPcInit) {
      if (ex.leader()) {
        ws.cur = 0;
        ws.status = kStOk;
        ws.aln_len = 0;
        ws.cells = 0;
        ws.alignments = 0;
        ws.sorts = ws.sorts_hbm = 0;
        ws.nMain = 0;
        sl.g[0].nV = 0;
        sl.g[0].nE = 0;
        *out_len = 0;
        for (int i = 0; i < kPhCount; ++i) ws.phase[i] = 0;
      }
      ex.sync();
      // every layer must fit the staging buffers (rows wider than the fast kernels' go to the wide kernel)
      for (uint32_t t = 0; t < nseq; ++t) {
        const uint32_t len = layer_len(rank[t]);
        if (len > sl.max_len) {
          if (ex.leader()) fail(kStTooLong);
        }
      }
      ex.sync();
      if (ws.status != kStOk) return finish();
      // backbone: AddAlignment with an empty alignment (window.cpp:197-201)
      uint8_t* codes = stage_codes(bb);
      add_alignment(codes, bb, blen);
      tick(kPhAddAln);
      if (ws.status != kStOk) return finish();
      changed = true;
      j = 1;
      act = kBuildNext;
    } else if (pc == kPcBuildPost) {
      const uint32_t l = ws.fill_layer;
      uint8_t* codes = stage_codes(l);
      add_alignment(codes, l, layer_len(l));
      tick(kPhAddAln);
      if (ws.status != kStOk) return finish();
      changed = true;
      ++j;
      act = kBuildNext;
    } else if (pc == kPcRoundPost) {
      // every alignment of the round has added its weights (window.cpp:329-386, graph.cpp:1104-1165)
      fold_weights();
      tick(kPhAddW);
      if (ws.status != kStOk) return finish();
      act = kRoundEnd;
    } else if (pc == kPcFinalPost) {
      // graph.cpp:1167-1179
      if (ex.leader()) {
        Graph& g = G();
        uint32_t n = 0;
        for (uint32_t t = ws.aln_len; t-- > 0;) {
          const int32_t nd = sl.aln_node[t];
          if (nd == -1) continue;
          if (n >= out_cap) {
            fail(kStOutOverflow);
            break;
          }
          out[n++] = bv.decoder[g.code[nd]];
        }
        *out_len = n;
      }
      ex.sync();
      tick(kPhEmit);
      return finish();
    } else if (pc == kPcLinearFinal) {
      // linear mode: consensus + coverage trim (window.cpp:138-171); the graph was re-sorted by step_prepare
      build_out_csr();
      tick(kPhCsr);
      if (ex.leader()) {
        uint32_t* cov = sl.tmp0;
        uint32_t n = heaviest_bundle(ws.nMain, out, out_cap, cov);
        if (ws.status == kStOk) {
          uint32_t b = 0, e = n;
          if ((bv.win_flags[w] & 1u) && trim) {
            const uint32_t avg = (nseq - 1) / 2;
            int32_t bi = 0, ei = static_cast<int32_t>(n) - 1;
            for (; bi < static_cast<int32_t>(n); ++bi) {
              if (cov[bi] >= avg) break;
            }
            for (; ei >= 0; --ei) {
              if (cov[ei] >= avg) break;
            }
            if (bi < ei) {
              b = bi;
              e = ei + 1;
            }
          }
          for (uint32_t i = b; i < e; ++i) out[i - b] = out[i];
          *out_len = e - b;
        }
      }
      ex.sync();
      tick(kPhEmit);
      return finish();
    } else {
      return;
    }

    while (true) {
      if (act == kBuildNext) {
        // build loop (window.cpp:239-298 / :100-136)
        if (j < nseq) {
          const uint32_t l = rank[j];
          const uint32_t lb = bv.begin[l], le = bv.end[l];
          const bool full = lb < offset && le > blen - offset;
          return plan(kPcBuildPost, kPrepFill | kPrepRowprog | (full ? 0u : kPrepSubSort), l, kModeNW, false);
        }
        if (!haplotype) return plan(kPcLinearFinal, 0u, bb, kModeNW, false);
        // haplotype mode: prune (window.cpp:300-319); LargestSubgraph rebuilds the CSR it needs
        prune(min_confidence, min_support, avgw);
        tick(kPhPrune);
        largest = true;  // LargestSubgraph runs in step_prepare, next to the sort (same staged-graph storage)
        changed = true;
        k = 0;
        act = kRoundStart;
      } else if (act == kRoundStart) {
        // re-align + re-weight rounds (window.cpp:329-386): the graph is frozen during a round, so its nseq
        // alignments (backbone and full-span layers global, the others local) are independent of each other
        if (k + 1 < num_prune) return plan(kPcRoundPost, kPrepFill | kPrepRowprog, bb, kModeNW, true);
        act = kFinalSchedule;
      } else if (act == kRoundEnd) {
        prune(min_confidence, min_support, avgw);
        tick(kPhPrune);
        largest = true;
        changed = true;
        ++k;
        act = kRoundStart;
      } else {
        // final local alignment of the backbone (window.cpp:391-394)
        return plan(kPcFinalPost, kPrepFill | kPrepRowprog, bb, kModeSW, false);
      }
    }
  }

  // mode of sequence j of a re-alignment round (window.cpp:338-352): the backbone and full-span layers are aligned
  // globally, the others locally (SW engine, hard-wired 3/-5/-4)
  VGC_HD VGC_INL uint32_t round_mode(uint32_t win, uint32_t jj) const {
    if (jj == 0) return kModeNW;
    const uint32_t first = bv.win_first[win];
    const uint32_t* rank = bv.layer_rank + first;
    const uint32_t blen = layer_len(rank[0]);
    const uint32_t offset = static_cast<uint32_t>(0.01 * blen);
    const uint32_t l = rank[jj];
    return (bv.begin[l] < offset && bv.end[l] > blen - offset) ? kModeNW : kModeSW;
  }

  // ---- end of a round: the weights the round's alignments left in wacc[rank * kInlinePreds + in-edge slot] go to
  //      the edges (rows beyond kInlinePreds in-edges were added to ew directly)
  VGC_HD void fold_weights() {
    Graph& g = G();
    const uint32_t S = sl.in_stride;
    const uint32_t nR = ws.nMain;
    for (uint32_t i = ex.lane(); i < nR * kInlinePreds; i += ex.width()) {
      const uint32_t w = sl.wacc[i];
      if (w == 0) continue;
      const uint32_t v = sl.r2n[i / kInlinePreds], p = i % kInlinePreds;
      if (p >= g.nin[v]) {
        fail(kStInternal);
        continue;
      }
      g.ew[g.ieid[v * S + p]] += w;
    }
    ex.sync();
  }

  // Whole window in one go (host model).
  VGC_HD void run_window(uint32_t w, bool haplotype, bool trim, double min_confidence, double min_support,
                         uint32_t num_prune, uint8_t* out, uint32_t* out_len) {
    if (ex.leader()) {
      ws.pc = kPcInit;
      ws.need = kNeedUpdate;
      ws.j = ws.k = ws.nMain = ws.prep = 0;
    }
    ex.sync();
    while (ws.pc != kPcDone) {
      step_trace();
      step_update(w, haplotype, trim, min_confidence, min_support, num_prune, out, out_len);
      step_prepare(w);
      if (ws.pc != kPcDone && ws.need == kNeedFill) {
        if (ws.round) {
          // the device runs these nseq alignments concurrently (align kernel); here one after the other
          const uint32_t nseq = bv.win_nseq[w];
          const uint32_t* rank = bv.layer_rank + bv.win_first[w];
          for (uint32_t jj = 0; jj < nseq && ws.status == kStOk; ++jj) {
            const uint32_t l = rank[jj], mode = round_mode(w, jj);
            uint8_t* codes = stage_codes(l);
            ex.template fill<K>(sl, ws, codes, layer_len(l), mode, mode == kModeNW ? nw : sw, bv.num_codes);
            ex.sync();
            traceback(l, mode);
            if (ws.status != kStOk) break;
            add_weights(l);
          }
          if (ws.status != kStOk) {
            finish();
          } else {
            if (ex.leader()) ws.need = kNeedUpdate;
            ex.sync();
          }
        } else {
          const uint32_t l = ws.fill_layer;
          uint8_t* codes = stage_codes(l);
          ex.template fill<K>(sl, ws, codes, layer_len(l), ws.fill_mode, ws.fill_mode == kModeNW ? nw : sw, bv.num_codes);
          ex.sync();
          if (ex.leader()) ws.need = kNeedTrace;
          ex.sync();
        }
      }
    }
  }
};

}  // namespace vgc

#endif  // VGC_POA_CORE_H_

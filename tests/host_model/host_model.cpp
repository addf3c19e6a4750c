// TEST INFRASTRUCTURE — host instantiation of the engine's algorithm template (vechat_b200/csrc/poa_core.h)
// with a one-lane executor and a scalar DP fill that writes the same H-matrix layout as the CUDA fill.
// It exists so the serial graph logic that runs inside the kernel (AddAlignment, TopologicalSort, Subgraph,
// PruneGraph, LargestSubgraph, traceback, ...) can be compared with the oracle on a machine without a GPU.
// It is built only by tests/ (tests/test_host_model.py) and is never linked into libvgc.so.

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define VGC_CHECK_ORDER 1
unsigned long long g_order_checks[4];  // incremental order: equal to the full sort / fell back / DIFFERENT
#include "host_prep.h"
#include "poa_core.h"
#include "vgc.h"

namespace {

using namespace vgc;

// diagnostics: histogram of predecessor row distances over every row of every fill (hm_dist_hist)
unsigned long long g_dhist[4][66];
unsigned long long g_npred[16];

struct HostEx {
  std::vector<uint32_t> rec32;
  std::vector<uint16_t> tail16, stk16;
  std::vector<uint8_t> codes;
  bool allow_fast = true;
  uint32_t small_stack = 0;  // test hook: tiny fast stack to force the overflow path

  unsigned long long clock() const { return 0; }
  int lane() const { return 0; }
  int width() const { return 1; }
  bool leader() const { return true; }
  void sync() {}
  uint32_t atomic_add(uint32_t* p, uint32_t v) {
    uint32_t o = *p;
    *p += v;
    return o;
  }
  uint32_t bcast(uint32_t v, uint32_t /*src*/) { return v; }
  uint32_t ballot(bool p) { return p ? 1u : 0u; }
  std::vector<U4> tile_h, tile_r;  // U4: the walker fetches 16-byte vectors
  void trace_tile(uint32_t** th, U4** tr) {
    tile_h.assign(kTR * kTW / 4, U4{0, 0, 0, 0});
    tile_r.assign(kTR, U4{0, 0, 0, 0});
    *th = reinterpret_cast<uint32_t*>(tile_h.data());
    *tr = tile_r.data();
  }
  unsigned long long reduce_add64(unsigned long long v) { return v; }
  uint32_t reduce_min(uint32_t v) { return v; }
  uint32_t reduce_max(uint32_t v) { return v; }
  uint32_t excl_scan(uint32_t v, uint32_t* total) {
    *total = v;
    return 0;
  }
  bool stage_fast(uint32_t nV, uint32_t nA, uint32_t** r, uint16_t** t, uint16_t** s, uint32_t* cap) {
    if (!allow_fast || nV >= 65535 || nA >= 65535) return false;
    rec32.assign(nV + 1, 0);
    tail16.assign(nA + 1, 0);
    uint32_t c = small_stack ? small_stack : 1024;
    stk16.assign(c, 0);
    *r = rec32.data();
    *t = tail16.data();
    *s = stk16.data();
    *cap = c;
    return true;
  }
  std::vector<uint16_t> l_off, l_adj, l_stk;
  std::vector<uint8_t> l_vis;
  bool stage_lsg(uint32_t nV, uint32_t nA, uint16_t** o, uint16_t** t, uint8_t** vis, uint16_t** s, uint32_t* cap) {
    if (!allow_fast || nV >= 65535 || nA >= 65535) return false;
    l_off.assign(nV + 2, 0);
    l_adj.assign(nA + 1, 0);
    l_vis.assign(nV + 1, 0);
    uint32_t c = small_stack ? small_stack : 1024;
    l_stk.assign(c, 0);
    *o = l_off.data();
    *t = l_adj.data();
    *vis = l_vis.data();
    *s = l_stk.data();
    *cap = c;
    return true;
  }
  std::vector<uint32_t> blk;
  uint32_t block_bytes = 1u << 20;  // test hook: a small value forces order_update's fallback to the full sort
  void block_arena(uint32_t** base, uint32_t* bytes) {
    blk.assign(block_bytes / 4 + 4, 0);
    *base = blk.data();
    *bytes = block_bytes;
  }
  uint8_t* seq_codes() {
    if (codes.size() < 70000) codes.assign(70000, 0);
    return codes.data();
  }

  // the wide path's matrix: int32, row-major, cols = len + 1 (poa_wide.cuh on the device)
  bool force_wide = false;
  void fill_wide(Slot& sl, WinState& ws, const uint8_t* codes_, uint32_t len, uint32_t mode, const Scores& sc) {
    const uint32_t nR = ws.nR;
    const uint64_t cols = len + 1;
    int32_t* H = reinterpret_cast<int32_t*>(sl.H);
    for (uint32_t j = 0; j <= len; ++j) H[j] = mode == kModeSW ? 0 : static_cast<int32_t>(j) * sc.g;
    int32_t best = mode == kModeSW ? 0 : INT32_MIN;
    uint32_t best_row = 0, best_col = 0;
    for (uint32_t r = 0; r < nR; ++r) {
      const U4 er = {sl.rowprog[4 * r], sl.rowprog[4 * r + 1], sl.rowprog[4 * r + 2], sl.rowprog[4 * r + 3]};
      const uint32_t code = meta_code(er.x), npred = meta_npred(er.x), np = npred == 0 ? 1 : npred;
      int32_t* out = H + (r + 1) * cols;
      for (uint32_t j = 0; j <= len; ++j) out[j] = INT32_MIN / 2;
      for (uint32_t p = 0; p < np; ++p) {
        const uint32_t pr = npred == 0 ? 0 : rec_pred(er, r + 1, p, sl.ovf);
        const int32_t* hp = H + pr * cols;
        out[0] = std::max(out[0], hp[0] + sc.g);
        for (uint32_t j = 1; j <= len; ++j) {
          const int32_t s = codes_[j - 1] == code ? sc.m : sc.x;
          out[j] = std::max(out[j], std::max(hp[j - 1] + s, hp[j] + sc.g));
        }
      }
      if (mode == kModeSW) out[0] = 0;
      for (uint32_t j = 1; j <= len; ++j) {
        out[j] = std::max(out[j], out[j - 1] + sc.g);
        if (mode == kModeSW) {
          out[j] = std::max(out[j], 0);
          if (best < out[j]) {
            best = out[j];
            best_row = r + 1;
            best_col = j;
          }
        } else if ((er.x & kMetaSink) && j == len && best < out[j]) {
          best = out[j];
          best_row = r + 1;
          best_col = j;
        }
      }
    }
    ws.best_row = best_row;
    ws.best_col = best_col;
    ws.best_score = best;
  }

  template <int K>
  void fill(Slot& sl, WinState& ws, const uint8_t* codes_, uint32_t len, uint32_t mode, const Scores& sc,
            uint32_t /*num_codes*/) {
    const uint32_t nR = ws.nR;
    if (force_wide) ws.wide = 1;
    if (ws.wide) return fill_wide(sl, ws, codes_, len, mode, sc);
    const uint32_t half = 32u * fill_width(K, len);  // the device's per-alignment row width
    ws.fill_k = half / 32u;
    auto cell = [&](uint32_t row, uint32_t c) -> int16_t* {  // lane-major words: low half = column w, high = 32K + w
      const uint32_t h = c >= half ? 1u : 0u;
      return reinterpret_cast<int16_t*>(sl.H + static_cast<uint64_t>(row) * sl.row_words + (h ? c - half : c)) + h;
    };
    auto H = [&](uint32_t row, uint32_t j) -> int32_t {  // j = DP column, 0 = first column
      if (j == 0) return mode == kModeSW ? 0 : sl.fc[row];
      return *cell(row, j - 1);
    };
    // virtual row 0
    sl.fc[0] = 0;
    for (uint32_t c = 0; c < len; ++c) *cell(0, c) = static_cast<int16_t>(mode == kModeSW ? 0 : (c + 1) * sc.g);
    int32_t best = mode == kModeSW ? 0 : INT32_MIN;
    uint32_t best_row = 0, best_col = 0;
    std::vector<int32_t> row(len + 1);
    for (uint32_t r = 0; r < nR; ++r) {
      // rows live in rank space: row = r + 1; rowprog[r] = {meta, p0, p1, p2 | ovf offset}, predecessors as rows
      const uint32_t meta = sl.rowprog[4 * r];
      const uint32_t code = meta_code(meta);
      const uint32_t npred = meta_npred(meta);
      const uint32_t np = npred == 0 ? 1 : npred;
      int32_t fcv = INT32_MIN;
      g_npred[np < 15 ? np : 15] += 1;
      for (uint32_t j = 1; j <= len; ++j) row[j] = INT32_MIN;
      for (uint32_t p = 0; p < np; ++p) {
        uint32_t pr;
        const U4 er = {sl.rowprog[4 * r], sl.rowprog[4 * r + 1], sl.rowprog[4 * r + 2], sl.rowprog[4 * r + 3]};
        if (npred == 0) pr = 0;
        else pr = rec_pred(er, r + 1, p, sl.ovf);
        {
          const uint32_t dd = r + 1 - pr;
          g_dhist[ws.round ? 1 : 0][dd < 65 ? dd : 65] += 1;
        }
        fcv = std::max(fcv, H(pr, 0));
        for (uint32_t j = 1; j <= len; ++j) {
          const int32_t s = codes_[j - 1] == code ? sc.m : sc.x;
          row[j] = std::max(row[j], std::max(H(pr, j - 1) + s, H(pr, j) + sc.g));
        }
      }
      row[0] = mode == kModeSW ? 0 : fcv + sc.g;
      sl.fc[r + 1] = static_cast<int16_t>(row[0]);
      for (uint32_t j = 1; j <= len; ++j) {
        row[j] = std::max(row[j], row[j - 1] + sc.g);
        if (mode == kModeSW) row[j] = std::max(row[j], 0);
        *cell(r + 1, j - 1) = static_cast<int16_t>(row[j]);
        if (mode == kModeSW) {
          if (best < row[j]) {
            best = row[j];
            best_row = r + 1;
            best_col = j;
          }
        } else if ((meta & kMetaSink) && j == len) {
          if (best < row[j]) {
            best = row[j];
            best_row = r + 1;
            best_col = j;
          }
        }
      }
    }
    ws.best_row = best_row;
    ws.best_col = best_col;
    ws.best_score = best;
  }
};

}  // namespace

extern "C" {

void hm_order_checks(unsigned long long* out) {
  for (int i = 0; i < 4; ++i) out[i] = g_order_checks[i];
}

void hm_dist_hist(unsigned long long* out, unsigned long long* np) {
  for (int i = 0; i < 2 * 66; ++i) out[i] = (&g_dhist[0][0])[i];
  for (int i = 0; i < 16; ++i) np[i] = g_npred[i];
}

// Same contract as ref_polish / oracle_polish.  flags: bit0 = disable the staged (16-bit) sort path, bit1 = tiny
// storage for the incremental order's dirty blocks, bit2 = every alignment on the wide (int32) path, bits 8-23 = fast-stack capacity override (0 = default).  k_regs selects the row template (10 or 16).
int hm_polish(const vgc_batch* b, const vgc_params* p, vgc_result* r, int flags, int k_regs, uint32_t* status_out) {
  Prepared prep;
  std::string err;
  int rc = prepare_batch(b, p, &prep, &err);
  if (rc != VGC_OK) return rc;
  const int K = k_regs;
  SlotDims d;
  d.max_nodes = static_cast<uint32_t>(std::max<uint64_t>(prep.max_nodes_ub, 16));
  d.max_edges = 2 * d.max_nodes + 64;
  d.max_len = std::max<uint32_t>(prep.max_len, 16);
  d.row_words = 32 * K;
  // the host model keeps whole DP matrices here (packed int16 rows, or int32 rows of len + 1 cells on the wide path)
  d.h_words = (static_cast<uint64_t>(d.max_nodes) + 1) * std::max<uint64_t>(d.row_words, d.max_len + 1);
  d.al_stride = prep.num_codes > 8 ? 16 : 8;
  d.in_stride = 8;
  for (uint32_t w = 0; w < b->n_windows; ++w) d.in_stride = std::max(d.in_stride, prep.win_nseq[w] + 1);
  std::vector<uint8_t> buf(slot_bytes(d));
  Slot sl;
  slot_carve(d, buf.data(), &sl);
  BatchView bv;
  bv.bases = b->bases;
  bv.quals = b->quals;
  bv.seq_off = b->seq_off;
  bv.has_qual = b->has_qual;
  bv.begin = b->begin;
  bv.end = b->end;
  bv.win_first = b->win_first;
  bv.win_flags = b->win_flags;
  bv.layer_rank = prep.layer_rank.data();
  bv.win_nseq = prep.win_nseq.data();
  bv.win_avgw = prep.win_avgw.data();
  bv.out_off = prep.out_off.data();
  bv.out_cap = prep.out_cap.data();
  bv.coder = prep.coder;
  bv.decoder = prep.decoder;
  bv.wlut = prep.wlut;
  bv.num_codes = prep.num_codes;
  std::vector<uint8_t> out(prep.out_total + 16);
  std::vector<uint32_t> out_len(b->n_windows, 0);
  HostEx ex;
  ex.allow_fast = !(flags & 1);
  ex.small_stack = (static_cast<uint32_t>(flags) >> 8) & 0xFFFFu;
  if (flags & 2) ex.block_bytes = 512;
  ex.force_wide = (flags & 4) != 0;  // every alignment through the wide (int32) path  // most incremental order updates then fall back to the full sort
  Scores nw{p->match, p->mismatch, p->gap};
  for (uint32_t w = 0; w < b->n_windows; ++w) {
    if (status_out) status_out[w] = 0;
    if (prep.win_nseq[w] < 3) {
      const uint32_t f = b->win_first[w];
      const uint32_t blen = static_cast<uint32_t>(b->seq_off[f + 1] - b->seq_off[f]);
      std::memcpy(out.data() + prep.out_off[w], b->bases + b->seq_off[f], blen);
      out_len[w] = blen;
      r->polished[w] = 0;
      continue;
    }
    WinState ws;
    std::memset(&ws, 0, sizeof(ws));
    uint32_t n = 0;
    if (K == 10) {
      Poa<HostEx, 10> poa(ex, bv, sl, ws, nw);
      poa.run_window(w, p->haplotype != 0, p->trim != 0, p->min_confidence, p->min_support, p->num_prune,
                     out.data() + prep.out_off[w], &n);
    } else {
      Poa<HostEx, 16> poa(ex, bv, sl, ws, nw);
      poa.run_window(w, p->haplotype != 0, p->trim != 0, p->min_confidence, p->min_support, p->num_prune,
                     out.data() + prep.out_off[w], &n);
    }
    if (status_out) status_out[w] = ws.status;
    if (ws.status != kStOk) n = 0;
    out_len[w] = n;
    r->polished[w] = 1;
  }
  uint64_t off = 0;
  for (uint32_t w = 0; w < b->n_windows; ++w) {
    r->cons_off[w] = off;
    if (off + out_len[w] > r->cons_capacity) return VGC_ERR_CAPACITY;
    std::memcpy(r->cons + off, out.data() + prep.out_off[w], out_len[w]);
    off += out_len[w];
  }
  r->cons_off[b->n_windows] = off;
  return VGC_OK;
}

}  // extern "C"

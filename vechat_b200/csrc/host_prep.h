// host_prep.h — host-side preparation of a vgc_batch for the device (pure C++, no CUDA).
//
// Everything here is work the reference does on the CPU *inside* Window::generate_consensus whose exact
// result depends on libstdc++ / libm and therefore stays on the host:
//   * layer rank: std::sort(rank.begin() + 1, rank.end(), by positions_.first) — unstable, same call,
//     same comparator, same initial arrangement (src/window.cpp:90-97, :203-210)
//   * quality -> weight LUT: (1 - pow(10, (33 - q) / 10.)) * 1000 truncated to uint32
//     (vendor/spoa/src/graph.cpp:169, src/window.cpp:366)
//   * average_weight: fp64 running sum in rank order (src/window.cpp:215-309)
//   * add_layer's validation and silent drops (src/window.cpp:47-72)
//   * windows with < 3 sequences return the backbone, polished = false (src/window.cpp:78-82, :188-192)
#ifndef VGC_HOST_PREP_H_
#define VGC_HOST_PREP_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "poa_core.h"
#include "vgc.h"

namespace vgc {

inline void weight_lut(uint32_t lut[256]) {
  for (int b = 0; b < 256; ++b) {
    volatile char q = static_cast<char>(b);         // volatile: force the run-time libm pow
    lut[b] = (1 - pow(10, (33 - q) / 10.)) * 1000;  // graph.cpp:169, same expression, same libm
  }
}

// window.cpp:235,295: the per-base addend of total_bases_weight, tabulated per quality byte.
inline const double* quality_value_lut() {
  static double lut[256];
  static bool init = [] {
    for (int b = 0; b < 256; ++b) {
      volatile char q = static_cast<char>(b);  // volatile: force the run-time libm pow, as in the reference
      lut[b] = 1 - pow(10, (33 - q) / 10.0);
    }
    return true;
  }();
  (void)init;
  return lut;
}

struct Prepared {
  std::vector<uint32_t> layer_rank;  // [n_layers]
  std::vector<uint32_t> win_nseq;    // [n_windows]
  std::vector<double> win_avgw;      // [n_windows]
  std::vector<uint64_t> out_off;     // [n_windows]
  std::vector<uint32_t> out_cap;     // [n_windows]
  std::vector<uint32_t> device_windows;  // windows that need the device (>= 3 sequences), heaviest first
  std::vector<uint64_t> win_work;    // estimated DP cells per window
  std::vector<uint32_t> win_sum_len; // sum of layer lengths (upper bound of graph nodes)
  std::vector<uint32_t> win_max_len; // longest layer
  std::vector<uint32_t> win_nfill;   // alignments (DP fills) the window program runs
  std::vector<uint8_t> layer_sw;     // [n_layers] 1: a re-alignment round aligns this layer locally (window.cpp:338-352)
  uint8_t coder[256];
  uint8_t decoder[kMaxCodes];
  uint32_t num_codes = 0;
  uint32_t wlut[256];
  uint32_t max_len = 0;        // longest layer among device windows
  uint64_t max_nodes_ub = 0;   // largest per-window node upper bound (sum of layer lengths)
  uint64_t out_total = 0;
};

// Returns VGC_OK or VGC_ERR_INVALID / VGC_ERR_CAPACITY with a message.  Windows are independent, so the
// per-window part runs on a few host threads (the reference does the same work inside its per-window tasks).
inline int prepare_batch(const vgc_batch* b, const vgc_params* p, Prepared* out, std::string* err,
                         unsigned max_threads = 16) {
  const uint32_t nw = b->n_windows, nl = b->n_layers;
  if (p->gap > 0 || (p->num_prune == 0 && p->haplotype)) {
    *err = "invalid params: gap must be non-positive, num_prune >= 1";
    return VGC_ERR_INVALID;
  }
  if (nw && (!b->win_first || !b->seq_off || !b->bases || !b->begin || !b->end || !b->has_qual || !b->win_flags)) {
    *err = "null array in batch";
    return VGC_ERR_INVALID;
  }
  if (nw && b->win_first[nw] != nl) {
    *err = "win_first[n_windows] != n_layers";
    return VGC_ERR_INVALID;
  }
  weight_lut(out->wlut);
  out->layer_rank.assign(nl, 0);
  out->win_nseq.assign(nw, 0);
  out->win_avgw.assign(nw, 0.0);
  out->out_off.assign(nw, 0);
  out->out_cap.assign(nw, 0);
  out->win_work.assign(nw, 0);
  out->win_sum_len.assign(nw, 0);
  out->win_max_len.assign(nw, 0);
  out->win_nfill.assign(nw, 0);
  out->layer_sw.assign(nl, 0);
  out->device_windows.clear();
  out->max_len = 0;
  out->max_nodes_ub = 0;
  const double* qv = quality_value_lut();  // initialise the table before the threads start

  struct Part {
    int rc = VGC_OK;
    uint32_t bad_window = 0xFFFFFFFFu;
    std::string err;
    bool seen[256] = {false};
  };
  auto run_range = [&](uint32_t w0, uint32_t w1, Part* part) {
    std::vector<uint32_t> rank;
    auto bail = [&](uint32_t w, int rc, const char* msg) {
      part->rc = rc;
      part->bad_window = w;
      part->err = msg;
    };
    for (uint32_t w = w0; w < w1; ++w) {
      const uint32_t f = b->win_first[w], l = b->win_first[w + 1];
      if (l <= f) return bail(w, VGC_ERR_INVALID, "window without a backbone");
      const uint64_t blen64 = b->seq_off[f + 1] - b->seq_off[f];
      if (blen64 == 0 || blen64 > 65535)  // createWindow (window.cpp:22-27); uint16 loop counters (:216,:232)
        return bail(w, VGC_ERR_INVALID, "empty or oversized backbone");
      const uint32_t blen = static_cast<uint32_t>(blen64);
      if (!b->has_qual[f] || !b->quals) return bail(w, VGC_ERR_INVALID, "backbone must carry a quality (real or dummy)");
      rank.clear();
      rank.push_back(f);
      uint64_t sum_len = blen;
      uint32_t max_len = blen;
      for (uint32_t i = f + 1; i < l; ++i) {
        const uint64_t len = b->seq_off[i + 1] - b->seq_off[i];
        const uint32_t bg = b->begin[i], en = b->end[i];
        if (len == 0 || bg == en) continue;  // add_layer returns silently (window.cpp:51-54)
        if (bg >= en || bg > blen || en > blen)
          return bail(w, VGC_ERR_INVALID, "layer begin and end positions are invalid");  // window.cpp:62-67
        if (len > 65535) return bail(w, VGC_ERR_INVALID, "layer longer than 65535");
        if (b->has_qual[i] && !b->quals) return bail(w, VGC_ERR_INVALID, "has_qual set but quals is NULL");
        rank.push_back(i);
        sum_len += len;
        max_len = std::max<uint32_t>(max_len, static_cast<uint32_t>(len));
      }
      const uint32_t nseq = static_cast<uint32_t>(rank.size());
      out->win_nseq[w] = nseq;
      {
        // full-span layers (and the backbone) are re-aligned globally, the others locally (window.cpp:212,338-352)
        const uint32_t offset = static_cast<uint32_t>(0.01 * blen);
        for (uint32_t j = 1; j < nseq; ++j) {
          const uint32_t i = rank[j];
          out->layer_sw[i] = (b->begin[i] < offset && b->end[i] > blen - offset) ? 0 : 1;
        }
      }
      std::sort(rank.begin() + 1, rank.end(),
                [&](uint32_t lhs, uint32_t rhs) { return b->begin[lhs] < b->begin[rhs]; });
      std::copy(rank.begin(), rank.end(), out->layer_rank.begin() + f);
      // alphabet: which bytes occur
      for (uint32_t j = 0; j < nseq; ++j) {
        const uint8_t* s = b->bases + b->seq_off[rank[j]];
        const uint64_t len = b->seq_off[rank[j] + 1] - b->seq_off[rank[j]];
        for (uint64_t k = 0; k < len; ++k) part->seen[s[k]] = true;
      }
      // output capacity: < 3 sequences -> the backbone itself; otherwise nodes touched by an alignment
      if (nseq < 3) {
        out->out_cap[w] = blen;
      } else {
        out->out_cap[w] = static_cast<uint32_t>(std::min<uint64_t>(sum_len, 2ull * max_len + 2ull * blen + 64));
        out->win_work[w] = sum_len * static_cast<uint64_t>(blen) * (p->haplotype ? 3 : 1) * nseq / 8 + 1;
        out->win_sum_len[w] = static_cast<uint32_t>(std::min<uint64_t>(sum_len, 0xFFFFFFFFu));
        out->win_max_len[w] = max_len;
        // fills of the window program (poa_core.h step_update): build + realign rounds + final, or build only
        out->win_nfill[w] = p->haplotype ? (nseq - 1) + (p->num_prune - 1) * nseq + 1 : (nseq - 1);
      }
      // average_weight (haplotype mode only): fp64 sum in rank order (window.cpp:215-309).  The addend
      // 1 - pow(10, (33 - q) / 10.0) is a pure function of the quality byte: tabulate it once (same libm, same
      // expression) and add the tabulated doubles in the reference's order — bit-identical, ~100x fewer pow calls.
      if (p->haplotype && nseq >= 3) {
        double total = 0.0;
        bool if_fasta = false;
        const uint16_t window_len = static_cast<uint16_t>(blen);
        if (b->win_flags[w] & VGC_WIN_DUMMY_QUAL) {
          total += blen;
          if_fasta = true;
        } else {
          const char* q = reinterpret_cast<const char*>(b->quals + b->seq_off[f]);
          for (uint16_t k = 0; k < blen; ++k) total += qv[static_cast<uint8_t>(q[k])];
        }
        for (uint32_t j = 1; j < nseq; ++j) {
          const uint32_t i = rank[j];
          const uint32_t len = static_cast<uint32_t>(b->seq_off[i + 1] - b->seq_off[i]);
          if (!b->has_qual[i]) {
            total += len;
          } else {
            const char* q = reinterpret_cast<const char*>(b->quals + b->seq_off[i]);
            for (uint16_t k = 0; k < len; ++k) total += qv[static_cast<uint8_t>(q[k])];
          }
        }
        out->win_avgw[w] = if_fasta ? 2.0 * total / window_len : 2.0 * total / window_len * 1000;
      }
    }
  };
  unsigned nt = nw >= 1024 ? std::min(std::max(1u, max_threads), std::max(1u, std::thread::hardware_concurrency())) : 1u;
  std::vector<Part> parts(nt);
  if (nt == 1) {
    run_range(0, nw, &parts[0]);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
      const uint32_t w0 = static_cast<uint32_t>(static_cast<uint64_t>(nw) * t / nt);
      const uint32_t w1 = static_cast<uint32_t>(static_cast<uint64_t>(nw) * (t + 1) / nt);
      th.emplace_back(run_range, w0, w1, &parts[t]);
    }
    for (auto& x : th) x.join();
  }
  for (const Part& part : parts) {  // ranges are in window order: the first failing window reports
    if (part.rc != VGC_OK) {
      *err = part.err;
      return part.rc;
    }
  }
  // alphabet: A C G T fixed, then every other byte that occurs, in byte order (the numbering is internal: codes are
  // only compared for equality and decoded back, so it does not need the reference's first-seen order)
  std::memset(out->coder, 0xFF, sizeof(out->coder));
  std::memset(out->decoder, 0, sizeof(out->decoder));
  const char* acgt = "ACGT";
  for (int i = 0; i < 4; ++i) {
    out->coder[static_cast<uint8_t>(acgt[i])] = static_cast<uint8_t>(i);
    out->decoder[i] = static_cast<uint8_t>(acgt[i]);
  }
  out->num_codes = 4;
  for (int c = 0; c < 256; ++c) {
    bool seen = false;
    for (const Part& part : parts) seen = seen || part.seen[c];
    if (!seen || out->coder[c] != 0xFF) continue;
    if (c >= 128) {  // the reference indexes spoa's 256-entry coder with a (signed) char: such input crashes it
      *err = "base byte >= 128";
      return VGC_ERR_INVALID;
    }
    if (out->num_codes >= static_cast<uint32_t>(kMaxCodes)) {
      *err = "more than 16 distinct base bytes in one batch (engine limit, see vgc_limits; the nucleotide alphabet "
             "with N and every IUPAC ambiguity code has 16)";
      return VGC_ERR_CAPACITY;
    }
    out->coder[c] = static_cast<uint8_t>(out->num_codes);
    out->decoder[out->num_codes++] = static_cast<uint8_t>(c);
  }
  uint64_t out_total = 0;
  for (uint32_t w = 0; w < nw; ++w) {
    out->out_off[w] = out_total;
    out_total += out->out_cap[w];
    if (out->win_nseq[w] >= 3) {
      out->device_windows.push_back(w);
      out->max_len = std::max(out->max_len, out->win_max_len[w]);
      out->max_nodes_ub = std::max<uint64_t>(out->max_nodes_ub, out->win_sum_len[w]);
    }
  }
  out->out_total = out_total;
  std::stable_sort(out->device_windows.begin(), out->device_windows.end(),
                   [&](uint32_t a, uint32_t c) {
                     if (out->win_nfill[a] != out->win_nfill[c]) return out->win_nfill[a] > out->win_nfill[c];
                     return out->win_work[a] > out->win_work[c];
                   });
  return VGC_OK;
}

// Bytes of one scratch slot for the given capacities, and the carving of a slot out of a flat buffer.
struct SlotDims {
  uint32_t max_nodes, max_edges, max_len, row_words;
  uint32_t in_stride = 8;  // in-list capacity per node (exact bound: the number of sequences of the window)
  uint32_t al_stride = 8;  // aligned-list capacity per node: 8, or 16 when the batch has more than 8 distinct bytes
  uint64_t h_words = 0;    // words behind Slot::H; 0 = graph_scratch_words() (device: the DP rows live in the align pool)
};

// Scratch the graph passes carve out of Slot::H: sort_graph = offsets (nV + 1) + adjacency (nE + sum of aligned
// counts) + DFS stack (<= adjacency + nV + 1); LargestSubgraph = offsets + live adjacency (2 nE) + stack (nV);
// heaviest bundle = 5 nV.  Aligned lists hold at most al_stride - 1 entries per node.
inline uint64_t graph_scratch_words(const SlotDims& d) {
  const uint64_t N = d.max_nodes, E = d.max_edges;
  return 2 * E + (4 + 2 * static_cast<uint64_t>(d.al_stride - 1)) * N + 1024;
}

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

template <class F>
inline uint64_t slot_walk(const SlotDims& d, F&& take) {
  uint64_t off = 0;
  auto t = [&](uint64_t bytes) {
    uint64_t o = off;
    off = align_up(off + bytes, 256);
    return take(o);
  };
  const uint64_t N = d.max_nodes, E = d.max_edges;
  (void)t;
  // order must match slot_carve below
  for (int gi = 0; gi < 2; ++gi) {
    t(N);                      // code
    t(N);                      // nal
    t(N * d.al_stride * 4);    // al
    t(N * 4);                  // nin
    t(N * 4);                  // nout
    t(N * 4);                  // cov
    t(E * 4);                  // etail
    t(E * 4);                  // ehead
    t(E * 4);                  // ew
    t(E * 4);                  // ein_ord
    t(E * 4);                  // eout_ord
    t(E);                      // edead
    t(N * d.in_stride * 4);    // itail
    t(N * d.in_stride * 4);    // ieid
  }
  t((N + 1) * 4);  // out_off
  t(E * 4);        // out_eid
  t(N * 4);        // r2n
  t(N * 4);        // order
  t(N * 4);        // rank_of
  t(N * 4);        // tmp0
  t(N * 4);        // tmp1
  t(N);            // flags
  t(N * 16);       // rowprog
  t(E * 4);        // ovf
  t((N + 1) * 2);  // fc
  t((d.h_words ? d.h_words : graph_scratch_words(d)) * 4);  // H
  t((static_cast<uint64_t>(d.max_len) + N + 2) * 4);     // aln_node
  t((static_cast<uint64_t>(d.max_len) + N + 2) * 4);     // aln_pos
  t(N * kInlinePreds * 4);                                // wacc
  t(N * 4);                                               // owner
  t(N * 4);                                               // bsize
  t(N * 4);                                               // bstart
  t(N * 4);                                               // bstart2
  t(N * 4);                                               // r2n2
  t((static_cast<uint64_t>(d.max_len) + 2) * 4);          // anch
  t(N);                                                   // dirty
  return off;
}

inline uint64_t slot_bytes(const SlotDims& d) {
  return slot_walk(d, [](uint64_t) { return 0; });
}

inline void slot_carve(const SlotDims& d, uint8_t* base, Slot* s) {
  uint64_t offs[96];  // (a pass carves tens of thousands of slots: no heap traffic here)
  size_t n_offs = 0;
  slot_walk(d, [&](uint64_t o) {
    offs[n_offs++] = o;
    return 0;
  });
  size_t i = 0;
  auto P = [&](auto** p) {
    using T = typename std::remove_pointer<typename std::remove_reference<decltype(*p)>::type>::type;
    *p = reinterpret_cast<T*>(base + offs[i++]);
  };
  s->max_nodes = d.max_nodes;
  s->max_edges = d.max_edges;
  s->max_len = d.max_len;
  s->row_words = d.row_words;
  s->in_stride = d.in_stride;
  s->al_stride = d.al_stride;
  for (int gi = 0; gi < 2; ++gi) {
    Graph& g = s->g[gi];
    g.nV = g.nE = 0;
    P(&g.code);
    P(&g.nal);
    P(&g.al);
    P(&g.nin);
    P(&g.nout);
    P(&g.cov);
    P(&g.etail);
    P(&g.ehead);
    P(&g.ew);
    P(&g.ein_ord);
    P(&g.eout_ord);
    P(&g.edead);
    P(&g.itail);
    P(&g.ieid);
  }
  P(&s->out_off);
  P(&s->out_eid);
  P(&s->r2n);
  P(&s->order);
  P(&s->rank_of);
  P(&s->tmp0);
  P(&s->tmp1);
  P(&s->flags);
  P(&s->rowprog);
  P(&s->ovf);
  P(&s->fc);
  P(&s->H);
  P(&s->aln_node);
  P(&s->aln_pos);
  P(&s->wacc);
  P(&s->owner);
  P(&s->bsize);
  P(&s->bstart);
  P(&s->bstart2);
  P(&s->r2n2);
  P(&s->anch);
  P(&s->dirty);
  s->aln_cap = d.max_len + d.max_nodes + 2;
  s->h_words = static_cast<uint32_t>(std::min<uint64_t>(d.h_words ? d.h_words : graph_scratch_words(d), 0xFFFFFFFFu));
}

}  // namespace vgc

#endif  // VGC_HOST_PREP_H_

// TEST INFRASTRUCTURE — not part of the product.
//
// C shim over the UNMODIFIED reference hot path.  It is compiled (by oracle/Makefile) together with
// the reference's own sources where they lie under /root/reference:
//   src/window.cpp  vendor/spoa/src/{alignment_engine,graph,sisd_alignment_engine,dispatcher}.cpp
// into oracle/_ref/libvechat_ref.so.  Nothing of the reference is copied into this repository; this
// file only *calls* racon::createWindow / Window::add_layer / Window::generate_consensus
// (src/window.hpp:27-55) and spoa::AlignmentEngine / spoa::Graph (public headers), driven by the same
// vgc_batch structure the product consumes (include/vgc.h).
//
// Used by tests/ (as the authority the CPU restatement in oracle/poa_oracle.cpp is pinned against)
// and by bench.py's cpu_baseline / --impl reference legs (kind = "reference").

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "window.hpp"
#include "spoa/spoa.hpp"
#include "vgc.h"

namespace {

struct WinOut {
  std::string cons;
  bool polished = false;
};

void run_window(const vgc_batch* b, const vgc_params* p, uint32_t w,
                std::shared_ptr<spoa::AlignmentEngine>& engine, WinOut* out) {
  const uint32_t first = b->win_first[w], last = b->win_first[w + 1];
  const uint64_t o0 = b->seq_off[first];
  const uint32_t blen = static_cast<uint32_t>(b->seq_off[first + 1] - o0);
  // Backbone quality: the reference compares the raw pointer as a C string against a run of '!'
  // (src/window.cpp:223).  The batch carries the outcome of that compare as VGC_WIN_DUMMY_QUAL;
  // build a NUL-terminated buffer for which the same compare gives the same outcome.
  std::string bq;
  if (b->win_flags[w] & VGC_WIN_DUMMY_QUAL) {
    bq.assign(blen, '!');
  } else {
    bq.assign(reinterpret_cast<const char*>(b->quals + o0), blen);
    bq.push_back('#');  // the pointer ran on into the rest of the read: never equal to '!' * blen
  }
  auto win = racon::createWindow(
      w, 0, (b->win_flags[w] & VGC_WIN_TGS) ? racon::WindowType::kTGS : racon::WindowType::kNGS,
      reinterpret_cast<const char*>(b->bases + o0), blen, bq.c_str(), blen);
  for (uint32_t l = first + 1; l < last; ++l) {
    const uint64_t o = b->seq_off[l];
    const uint32_t len = static_cast<uint32_t>(b->seq_off[l + 1] - o);
    const char* q = b->has_qual[l] ? reinterpret_cast<const char*>(b->quals + o) : nullptr;
    win->add_layer(reinterpret_cast<const char*>(b->bases + o), len, q, q ? len : 0,
                   b->begin[l], b->end[l]);
  }
  if (p->haplotype) {
    out->polished = win->generate_consensus(engine, p->trim != 0, true, p->min_confidence,
                                            p->min_support, p->num_prune);
  } else {
    out->polished = win->generate_consensus(engine, p->trim != 0);
  }
  out->cons = win->consensus();
}

}  // namespace

extern "C" {

// Same contract as vgc_polish, executed by the reference code on `threads` host threads
// (one NW engine per thread, as src/polisher.cpp:186-190).  Returns 0, or 4 if cons_capacity is short.
int ref_polish(const vgc_batch* b, const vgc_params* p, vgc_result* r, int threads) {
  std::vector<WinOut> outs(b->n_windows);
  if (threads < 1) threads = 1;
  std::atomic<uint32_t> cursor{0};
  auto worker = [&]() {
    std::shared_ptr<spoa::AlignmentEngine> engine =
        spoa::AlignmentEngine::Create(spoa::AlignmentType::kNW, p->match, p->mismatch, p->gap);
    engine->Prealloc(500, 5);
    while (true) {
      uint32_t w = cursor.fetch_add(1);
      if (w >= b->n_windows) break;
      run_window(b, p, w, engine, &outs[w]);
    }
  };
  if (threads == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  uint64_t off = 0;
  for (uint32_t w = 0; w < b->n_windows; ++w) {
    r->cons_off[w] = off;
    if (off + outs[w].cons.size() > r->cons_capacity) return 4;
    std::memcpy(r->cons + off, outs[w].cons.data(), outs[w].cons.size());
    off += outs[w].cons.size();
    r->polished[w] = outs[w].polished ? 1 : 0;
  }
  r->cons_off[b->n_windows] = off;
  return 0;
}

// spoa golden-vector driver (vendor/spoa/test/spoa_test.cpp:35-52,150-164,198-212,246-260,294-308):
// align `n` sequences one after another with a linear-gap engine and return GenerateConsensus().
// type: 0 = SW, 1 = NW.  quals may be NULL.  Returns the consensus length (written to out, capacity cap).
int ref_spoa_consensus(int type, int m, int n_, int g, uint32_t n, const char* const* seqs,
                       const char* const* quals, char* out, uint32_t cap) {
  auto engine = spoa::AlignmentEngine::Create(
      type == 0 ? spoa::AlignmentType::kSW : spoa::AlignmentType::kNW, m, n_, g);
  spoa::Graph graph{};
  for (uint32_t i = 0; i < n; ++i) {
    std::string s(seqs[i]);
    auto a = engine->Align(s, graph);
    if (quals) {
      graph.AddAlignment(a, s, std::string(quals[i]));
    } else {
      graph.AddAlignment(a, s);
    }
  }
  std::string c = graph.GenerateConsensus();
  if (c.size() + 1 > cap) return -1;
  std::memcpy(out, c.c_str(), c.size() + 1);
  return static_cast<int>(c.size());
}

// Single alignment probe: build a graph from `n` sequences (NW engine m/x/g, as the window build
// loop does for full-span layers), then align `query` with engine `type` and return the alignment
// pairs.  Used to pin the alignment kernel alone.  Returns the number of pairs (<= cap) or -1.
int ref_align_probe(int type, int m, int x, int g, uint32_t n, const char* const* seqs,
                    const char* query, int32_t* out_node, int32_t* out_pos, uint32_t cap) {
  auto nw = spoa::AlignmentEngine::Create(spoa::AlignmentType::kNW, m, x, g);
  spoa::Graph graph{};
  for (uint32_t i = 0; i < n; ++i) {
    std::string s(seqs[i]);
    auto a = nw->Align(s, graph);
    graph.AddAlignment(a, s);
  }
  auto engine = spoa::AlignmentEngine::Create(
      type == 0 ? spoa::AlignmentType::kSW : spoa::AlignmentType::kNW, m, x, g);
  auto a = engine->Align(std::string(query), graph);
  if (a.size() > cap) return -1;
  for (size_t i = 0; i < a.size(); ++i) {
    out_node[i] = a[i].first;
    out_pos[i] = a[i].second;
  }
  return static_cast<int>(a.size());
}

const char* ref_build_info(void) {
  return "reference: src/window.cpp + vendor/spoa/src/{alignment_engine,graph,sisd_alignment_engine,"
         "dispatcher}.cpp, g++ -O3 -DNDEBUG -msse4.1";
}

}  // extern "C"

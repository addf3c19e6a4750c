// vgc_host.hpp — C++ host side above the C-ABI (include/vgc.h), mirroring the reference's own surface for this
// path so that VeChat's code (and its tests) read the same against either engine:
//
//   racon::createWindow / Window::add_layer / Window::consensus / id / rank    src/window.hpp:27-77, window.cpp:17-72
//   racon::Polisher::polish (fan-out + in-order stitch + header tags)           src/polisher.cpp:491-562
//
// What differs, by design: generate_consensus is not a per-window call.  B200Polisher::polish hands ALL windows to
// the GPU engine in batches (vgc_polish) and then runs the reference's stitching loop over the results.  There is
// no CPU consensus code here: without libvgc.so + a B200 the polish call fails (error handler, default exit(1)
// like the reference's error sites).
#ifndef VGC_HOST_HPP_
#define VGC_HOST_HPP_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "vgc.h"

namespace vgc_host {

// Error behaviour of the reference: fprintf(stderr, "[racon::...] error: ...") + exit(1) (window.cpp:24-27,58-67).
// Tests may install a handler that throws instead.
using ErrorHandler = std::function<void(const std::string&)>;
inline ErrorHandler& error_handler() {
  static ErrorHandler h = [](const std::string& msg) {
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(1);
  };
  return h;
}
inline void fail(const std::string& msg) { error_handler()(msg); }

enum class WindowType { kNGS, kTGS };  // src/window.hpp:21-24
enum class PolisherType { kC, kF };    // src/polisher.hpp:36-39

class Window;
std::shared_ptr<Window> createWindow(uint64_t id, uint32_t rank, WindowType type, const char* backbone,
                                     uint32_t backbone_length, const char* quality, uint32_t quality_length);

class Window {
 public:
  uint64_t id() const { return id_; }
  uint32_t rank() const { return rank_; }
  const std::string& consensus() const { return consensus_; }
  bool polished() const { return polished_; }

  // src/window.cpp:47-72: silently ignores empty layers / begin == end, rejects inconsistent input
  void add_layer(const char* sequence, uint32_t sequence_length, const char* quality, uint32_t quality_length,
                 uint32_t begin, uint32_t end) {
    if (sequence_length == 0 || begin == end) return;
    if (quality != nullptr && sequence_length != quality_length) {
      fail("[racon::Window::add_layer] error: unequal quality size!");
      return;
    }
    if (begin >= end || begin > sequences_.front().second || end > sequences_.front().second) {
      fail("[racon::Window::add_layer] error: layer begin and end positions are invalid!");
      return;
    }
    sequences_.emplace_back(sequence, sequence_length);
    qualities_.emplace_back(quality, quality_length);
    positions_.emplace_back(begin, end);
  }

  friend std::shared_ptr<Window> createWindow(uint64_t, uint32_t, WindowType, const char*, uint32_t, const char*,
                                              uint32_t);
  friend class BatchPacker;  // the role `friend class CUDABatchProcessor` plays in the reference (window.hpp:61-63)

 private:
  Window(uint64_t id, uint32_t rank, WindowType type, const char* backbone, uint32_t backbone_length,
         const char* quality, uint32_t quality_length)
      : id_(id), rank_(rank), type_(type) {
    sequences_.emplace_back(backbone, backbone_length);
    qualities_.emplace_back(quality, quality_length);
    positions_.emplace_back(0, 0);
  }
  uint64_t id_;
  uint32_t rank_;
  WindowType type_;
  std::string consensus_;
  bool polished_ = false;
  // views into caller-owned read strings, exactly as in the reference (window.hpp:74-76)
  std::vector<std::pair<const char*, uint32_t>> sequences_;
  std::vector<std::pair<const char*, uint32_t>> qualities_;
  std::vector<std::pair<uint32_t, uint32_t>> positions_;
};

inline std::shared_ptr<Window> createWindow(uint64_t id, uint32_t rank, WindowType type, const char* backbone,
                                            uint32_t backbone_length, const char* quality, uint32_t quality_length) {
  if (backbone_length == 0 || backbone_length != quality_length) {
    fail("[racon::createWindow] error: empty backbone sequence/unequal quality length!");
    return nullptr;
  }
  return std::shared_ptr<Window>(new Window(id, rank, type, backbone, backbone_length, quality, quality_length));
}

// Flat structure-of-arrays image of a run of windows: what vgc_batch points into.
struct PackedBatch {
  std::vector<uint8_t> bases, quals, has_qual, win_flags;
  std::vector<uint64_t> seq_off;
  std::vector<uint32_t> begin, end, win_first;
  vgc_batch view() const {
    vgc_batch b;
    b.n_windows = static_cast<uint32_t>(win_flags.size());
    b.n_layers = static_cast<uint32_t>(begin.size());
    b.bases = bases.data();
    b.quals = quals.data();
    b.seq_off = seq_off.data();
    b.has_qual = has_qual.data();
    b.begin = begin.data();
    b.end = end.data();
    b.win_first = win_first.data();
    b.win_flags = win_flags.data();
    return b;
  }
};

class BatchPacker {
 public:
  // Copies the bytes the windows borrow (Window does not own them; the reference frees them when polish ends,
  // polisher.cpp:560-561).
  static void pack(const std::vector<std::shared_ptr<Window>>& w, size_t first, size_t last, PackedBatch* p,
                   unsigned num_threads = 0) {
    *p = PackedBatch();
    // pass 1 (serial, metadata only): layer table and byte offsets
    size_t n_layers = 0;
    for (size_t i = first; i < last; ++i) n_layers += w[i]->sequences_.size();
    p->win_first.reserve(last - first + 1);
    p->win_flags.reserve(last - first);
    p->seq_off.reserve(n_layers + 1);
    p->has_qual.reserve(n_layers);
    p->begin.reserve(n_layers);
    p->end.reserve(n_layers);
    uint64_t bytes = 0;
    for (size_t i = first; i < last; ++i) {
      const Window& win = *w[i];
      p->win_first.push_back(static_cast<uint32_t>(p->begin.size()));
      // src/window.cpp:223: `qualities_.front().first == std::string(len, '!')` compares the backbone quality as a
      // C string with a run of '!' of the backbone's length
      const bool dummy = std::string(win.sequences_.front().second, '!') == win.qualities_.front().first;
      p->win_flags.push_back(static_cast<uint8_t>((win.type_ == WindowType::kTGS ? VGC_WIN_TGS : 0u) |
                                                  (dummy ? VGC_WIN_DUMMY_QUAL : 0u)));
      for (size_t l = 0; l < win.sequences_.size(); ++l) {
        p->seq_off.push_back(bytes);
        bytes += win.sequences_[l].second;
        p->has_qual.push_back(win.qualities_[l].first != nullptr ? 1 : 0);
        p->begin.push_back(win.positions_[l].first);
        p->end.push_back(win.positions_[l].second);
      }
    }
    p->win_first.push_back(static_cast<uint32_t>(p->begin.size()));
    p->seq_off.push_back(bytes);
    // pass 2 (threads over window ranges): copy the bytes the windows borrow
    p->bases.resize(bytes);
    p->quals.resize(bytes);
    const size_t nw = last - first;
    unsigned nt = num_threads ? num_threads : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (nw < 256) nt = 1;
    auto copy_range = [&](size_t a, size_t b) {
      for (size_t i = a; i < b; ++i) {
        const Window& win = *w[first + i];
        size_t layer = p->win_first[i];
        for (size_t l = 0; l < win.sequences_.size(); ++l, ++layer) {
          const uint32_t len = win.sequences_[l].second;
          const uint64_t o = p->seq_off[layer];
          std::memcpy(p->bases.data() + o, win.sequences_[l].first, len);
          const char* q = win.qualities_[l].first;
          if (q != nullptr) std::memcpy(p->quals.data() + o, q, len);
          else std::memset(p->quals.data() + o, '!', len);
        }
      }
    };
    if (nt == 1) {
      copy_range(0, nw);
    } else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(copy_range, nw * t / nt, nw * (t + 1) / nt);
      for (auto& x : th) x.join();
    }
  }
  static void store(Window& win, const uint8_t* s, uint64_t n, bool polished) {
    win.consensus_.assign(reinterpret_cast<const char*>(s), n);
    win.polished_ = polished;
  }
};

struct Sequence {  // the two fields of racon::Sequence the polish stage produces (src/sequence.hpp)
  std::string name, data;
};

// The polish stage of racon::Polisher for windows already tiled by initialize().
class B200Polisher {
 public:
  // createPolisher's arguments that reach the hot path (src/polisher.hpp:42-49)
  B200Polisher(PolisherType type, bool haplotype, double min_confidence, double min_support, uint32_t num_prune,
               bool trim, int8_t match, int8_t mismatch, int8_t gap, int device = 0,
               uint32_t windows_per_batch = 65536)
      : type_(type), windows_per_batch_(windows_per_batch ? windows_per_batch : 1) {
    vgc_params prm = {};
    prm.match = match;
    prm.mismatch = mismatch;
    prm.gap = gap;
    prm.haplotype = haplotype ? 1 : 0;
    prm.trim = trim ? 1 : 0;
    prm.num_prune = num_prune;
    prm.min_confidence = min_confidence;
    prm.min_support = min_support;
    if (vgc_create(&engine_, device, &prm) != VGC_OK) {
      engine_ = nullptr;
      fail(std::string("[racon::B200Polisher] error: ") + vgc_last_error());
    }
  }
  ~B200Polisher() {
    if (engine_) vgc_destroy(engine_);
  }
  B200Polisher(const B200Polisher&) = delete;
  B200Polisher& operator=(const B200Polisher&) = delete;

  // Window::generate_consensus for every window (sets consensus(), polished()).  No CPU fallback.
  bool generate_consensus(std::vector<std::shared_ptr<Window>>& windows) {
    if (!engine_) return false;
    for (size_t first = 0; first < windows.size(); first += windows_per_batch_) {
      const size_t last = std::min(windows.size(), first + windows_per_batch_);
      PackedBatch p;
      BatchPacker::pack(windows, first, last, &p);
      const vgc_batch b = p.view();
      std::vector<uint8_t> cons(vgc_result_bound(&b));
      std::vector<uint64_t> off(b.n_windows + 1);
      std::vector<uint8_t> ok(b.n_windows);
      vgc_result r = {cons.data(), cons.size(), off.data(), ok.data()};
      if (vgc_polish(engine_, &b, &r, nullptr) != VGC_OK) {
        fail(std::string("[racon::B200Polisher::polish] error: ") + vgc_last_error());
        return false;
      }
      for (size_t i = first; i < last; ++i)
        BatchPacker::store(*windows[i], cons.data() + off[i - first], off[i - first + 1] - off[i - first],
                           ok[i - first] != 0);
    }
    return true;
  }

  // src/polisher.cpp:491-562.  names[id] / coverages[id] are indexed by Window::id() (sequences_[id]->name(),
  // targets_coverages_[id]).  Consumes `windows` like the reference (each entry is reset, the vector cleared).
  void polish(std::vector<std::shared_ptr<Window>>& windows, const std::vector<std::string>& names,
              const std::vector<uint32_t>& coverages, std::vector<std::unique_ptr<Sequence>>& dst,
              bool drop_unpolished_sequences) {
    if (!generate_consensus(windows)) return;
    std::string polished_data;
    uint32_t num_polished_windows = 0;
    for (uint64_t i = 0; i < windows.size(); ++i) {
      num_polished_windows += windows[i]->polished() ? 1 : 0;
      polished_data += windows[i]->consensus();
      if (i == windows.size() - 1 || windows[i + 1]->rank() == 0) {
        const double polished_ratio = num_polished_windows / static_cast<double>(windows[i]->rank() + 1);
        if (!drop_unpolished_sequences || polished_ratio > 0) {
          std::string tags = type_ == PolisherType::kF ? "r" : "";
          tags += " LN:i:" + std::to_string(polished_data.size());
          tags += " RC:i:" + std::to_string(coverages[windows[i]->id()]);
          tags += " XC:f:" + std::to_string(polished_ratio);
          dst.emplace_back(new Sequence{names[windows[i]->id()] + tags, polished_data});
        }
        num_polished_windows = 0;
        polished_data.clear();
      }
      windows[i].reset();
    }
    std::vector<std::shared_ptr<Window>>().swap(windows);
  }

 private:
  PolisherType type_;
  size_t windows_per_batch_;
  vgc_handle engine_ = nullptr;
};

}  // namespace vgc_host

#endif  // VGC_HOST_HPP_

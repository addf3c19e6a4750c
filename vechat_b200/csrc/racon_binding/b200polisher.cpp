// b200polisher.cpp — racon::B200Polisher: Polisher::polish (src/polisher.cpp:491-562) over the C-ABI of
// include/vgc.h.  Compiled with -DCUDA_ENABLED *for this translation unit only*, which makes the reference's
// own friend hook visible (src/window.hpp:61-63 `friend class CUDABatchProcessor;`): the class of that name below
// is how the binding reads Window::sequences_/qualities_/positions_/type_ and writes Window::consensus_ without
// touching a line of the reference.
#include "b200polisher.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "bioparser/fasta_parser.hpp"
#include "bioparser/fastq_parser.hpp"
#include "bioparser/mhap_parser.hpp"
#include "bioparser/paf_parser.hpp"
#include "bioparser/sam_parser.hpp"

#include "logger.hpp"
#include "overlap.hpp"
#include "sequence.hpp"
#include "window.hpp"

#include "vga.h"
#include "vgc.h"

namespace racon {

namespace {

[[noreturn]] void die(const char* where, const char* what) {
  std::fprintf(stderr, "[racon::%s] error: %s\n", where, what);
  std::exit(1);
}

// structure-of-arrays image of a run of windows (what vgc_batch points into)
// Bytes that are always overwritten before they are read: no zero-filling, capacity kept from batch to batch (a fresh
// 2 GB of zeroed vectors per batch cost more host time than the copy itself: page faults + memset on one thread;
// here the first touch of a new buffer happens in the copy threads).
struct RawBuf {
  std::unique_ptr<uint8_t[]> p;
  size_t cap = 0;
  void need(size_t n) {
    if (cap < n) {
      cap = n + n / 8 + 64;
      p.reset(new uint8_t[cap]);
    }
  }
  uint8_t* data() const { return p.get(); }
};

struct Packed {
  RawBuf bases, quals;
  std::vector<uint8_t> has_qual, win_flags;
  std::vector<uint64_t> seq_off;
  std::vector<uint32_t> begin, end, win_first;
  void clear() {  // keeps every capacity
    has_qual.clear();
    win_flags.clear();
    seq_off.clear();
    begin.clear();
    end.clear();
    win_first.clear();
  }
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

}  // namespace

class CUDABatchProcessor {
 public:
  // the layers of window `idx` beyond its backbone: the binding's tiles when it tiled (else nullptr, 0 and the
  // Window's own vectors are used)
  struct Tiles {
    const std::vector<B200TileLayer>* layers = nullptr;
    const std::vector<uint64_t>* first = nullptr;
    bool on() const { return first != nullptr && !first->empty(); }
    const B200TileLayer* begin(size_t idx) const {
      return idx + 1 < first->size() ? layers->data() + (*first)[idx] : nullptr;
    }
    size_t count(size_t idx) const { return idx + 1 < first->size() ? (*first)[idx + 1] - (*first)[idx] : 0; }
  };
  static uint64_t bytes(const Window& w, size_t idx, const Tiles& t) {
    uint64_t n = 0;
    if (!t.on()) {
      for (const auto& s : w.sequences_) n += s.second;
      return n;
    }
    n = w.sequences_.front().second;
    const B200TileLayer* l = t.begin(idx);
    for (size_t k = 0, c = t.count(idx); k < c; ++k) n += l[k].length;
    return n;
  }
  // Window only borrows pointers into Polisher::sequences_ (freed at polisher.cpp:560-561): copy the bytes once.
  // Pass 1 (serial, metadata only) lays out the layer table and the byte offsets; pass 2 copies the bytes with
  // `threads` host threads over window ranges.
  static void pack(const std::vector<std::shared_ptr<Window>>& w, size_t first, size_t last, Packed* p, unsigned threads,
                   const Tiles& tiles) {
    if (tiles.on()) return pack_tiled(w, first, last, p, threads, tiles);
    p->clear();
    size_t layers = 0;
    for (size_t i = first; i < last; ++i) layers += w[i]->sequences_.size();
    p->seq_off.reserve(layers + 1);
    p->has_qual.reserve(layers);
    p->begin.reserve(layers);
    p->end.reserve(layers);
    p->win_first.reserve(last - first + 1);
    p->win_flags.reserve(last - first);
    uint64_t total = 0;
    for (size_t i = first; i < last; ++i) {
      const Window& win = *w[i];
      p->win_first.push_back(static_cast<uint32_t>(p->begin.size()));
      // window.cpp:223 compares the backbone quality POINTER, as a C string, with a run of '!' of backbone length
      const bool dummy = std::string(win.sequences_.front().second, '!') == win.qualities_.front().first;
      p->win_flags.push_back(static_cast<uint8_t>((win.type_ == WindowType::kTGS ? VGC_WIN_TGS : 0u) |
                                                  (dummy ? VGC_WIN_DUMMY_QUAL : 0u)));
      for (size_t l = 0; l < win.sequences_.size(); ++l) {
        p->seq_off.push_back(total);
        total += win.sequences_[l].second;
        p->has_qual.push_back(win.qualities_[l].first != nullptr ? 1 : 0);
        p->begin.push_back(win.positions_[l].first);
        p->end.push_back(win.positions_[l].second);
      }
    }
    p->win_first.push_back(static_cast<uint32_t>(p->begin.size()));
    p->seq_off.push_back(total);
    p->bases.need(total);
    p->quals.need(total);
    const size_t nw = last - first;
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
    const unsigned nt = nw < 64 ? 1u : std::max(1u, std::min(threads, 16u));
    if (nt == 1) {
      copy_range(0, nw);
    } else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(copy_range, nw * t / nt, nw * (t + 1) / nt);
      for (auto& x : th) x.join();
    }
  }
  // the same image from the binding's tiles: layer 0 = the window's backbone (the Window holds it), then the tiles
  static void pack_tiled(const std::vector<std::shared_ptr<Window>>& w, size_t first, size_t last, Packed* p,
                         unsigned threads, const Tiles& tiles) {
    p->clear();
    const size_t nw = last - first;
    size_t layers = nw;
    for (size_t i = first; i < last; ++i) layers += tiles.count(i);
    p->seq_off.resize(layers + 1);
    p->has_qual.resize(layers);
    p->begin.resize(layers);
    p->end.resize(layers);
    p->win_first.resize(nw + 1);
    p->win_flags.resize(nw);
    uint64_t total = 0;
    size_t layer = 0;
    for (size_t i = first; i < last; ++i) {  // offsets: serial prefix over the windows (metadata only)
      const Window& win = *w[i];
      p->win_first[i - first] = static_cast<uint32_t>(layer);
      p->seq_off[layer++] = total;
      total += win.sequences_.front().second;
      const B200TileLayer* l = tiles.begin(i);
      for (size_t k = 0, c = tiles.count(i); k < c; ++k) {
        p->seq_off[layer++] = total;
        total += l[k].length;
      }
    }
    p->win_first[nw] = static_cast<uint32_t>(layer);
    p->seq_off[layer] = total;
    p->bases.need(total);
    p->quals.need(total);
    auto copy_range = [&](size_t a, size_t b) {
      for (size_t i = a; i < b; ++i) {
        const Window& win = *w[first + i];
        size_t lay = p->win_first[i];
        const uint32_t blen = win.sequences_.front().second;
        // window.cpp:223 compares the backbone quality POINTER, as a C string, with a run of '!' of backbone length
        const bool dummy = std::string(blen, '!') == win.qualities_.front().first;
        p->win_flags[i] = static_cast<uint8_t>((win.type_ == WindowType::kTGS ? VGC_WIN_TGS : 0u) |
                                               (dummy ? VGC_WIN_DUMMY_QUAL : 0u));
        uint64_t o = p->seq_off[lay];
        std::memcpy(p->bases.data() + o, win.sequences_.front().first, blen);
        const char* bq = win.qualities_.front().first;
        if (bq != nullptr) std::memcpy(p->quals.data() + o, bq, blen);
        else std::memset(p->quals.data() + o, '!', blen);
        p->has_qual[lay] = bq != nullptr ? 1 : 0;
        p->begin[lay] = win.positions_.front().first;
        p->end[lay] = win.positions_.front().second;
        ++lay;
        const B200TileLayer* l = tiles.begin(first + i);
        for (size_t k = 0, c = tiles.count(first + i); k < c; ++k, ++lay) {
          o = p->seq_off[lay];
          std::memcpy(p->bases.data() + o, l[k].data, l[k].length);
          if (l[k].quality != nullptr) std::memcpy(p->quals.data() + o, l[k].quality, l[k].length);
          else std::memset(p->quals.data() + o, '!', l[k].length);
          p->has_qual[lay] = l[k].quality != nullptr ? 1 : 0;
          p->begin[lay] = l[k].begin;
          p->end[lay] = l[k].end;
        }
      }
    };
    const unsigned nt = nw < 64 ? 1u : std::max(1u, std::min(threads, 16u));
    if (nt == 1) {
      copy_range(0, nw);
    } else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(copy_range, nw * t / nt, nw * (t + 1) / nt);
      for (auto& x : th) x.join();
    }
  }
  static void store(Window& win, const uint8_t* s, uint64_t n) { win.consensus_.assign(reinterpret_cast<const char*>(s), n); }
};

// The second friend hook of the reference (src/overlap.hpp:79-81 `friend class CUDABatchAligner;`): reads the overlap
// coordinates, writes Overlap::cigar_.
class CUDABatchAligner {
 public:
  static void clear_breaking_points(Overlap& o) { std::vector<std::pair<uint32_t, uint32_t>>().swap(o.breaking_points_); }
  // cut_on_device: vga_break fills breaking_points_ directly (no CIGAR text crosses PCIe); otherwise vga_align fills
  // cigar_ and the reference's find_breaking_points_from_cigar cuts it on the host.
  // One host thread + one vga_handle per device, each over a contiguous range of the overlaps (they are independent;
  // vga_* keeps no global state).
  static void align(std::vector<std::unique_ptr<Overlap>>& overlaps, const std::vector<std::unique_ptr<Sequence>>& sequences,
                    const std::vector<int>& devices, uint32_t window_length, bool cut_on_device) {
    for (const auto& o : overlaps)
      if (!o->is_transmuted_) die("Overlap::find_breaking_points", "overlap is not transmuted!");
    const size_t nd = devices.size(), n = overlaps.size();
    std::vector<std::string> errors(nd);
    auto run_device = [&](size_t d) {
      align_range(overlaps, sequences, n * d / nd, n * (d + 1) / nd, devices[d], window_length, cut_on_device, &errors[d]);
    };
    if (nd == 1) {
      run_device(0);
    } else {
      std::vector<std::thread> th;
      for (size_t d = 0; d < nd; ++d) th.emplace_back(run_device, d);
      for (auto& t : th) t.join();
    }
    for (const std::string& e : errors)
      if (!e.empty()) die("B200Polisher::find_overlap_breaking_points", e.c_str());  // no CPU fallback
  }

 private:
  static void align_range(std::vector<std::unique_ptr<Overlap>>& overlaps,
                          const std::vector<std::unique_ptr<Sequence>>& sequences, size_t range_first, size_t range_last,
                          int device, uint32_t window_length, bool cut_on_device, std::string* error) {
    if (range_first >= range_last) return;
    vga_handle h = nullptr;
    if (vga_create(&h, device) != VGA_OK) {
      *error = vga_last_error();
      return;
    }
    constexpr size_t kChunk = 1u << 18;  // overlaps per vga_align / vga_break call
    for (size_t first = range_first; first < range_last && error->empty(); first += kChunk) {
      const size_t last = std::min(range_last, first + kChunk);
      // every sequence (strand) the chunk touches goes into the byte buffer once
      std::vector<uint8_t> seqs;
      std::vector<uint64_t> where(2 * sequences.size(), ~0ull);
      auto place = [&](uint64_t id, bool rc) -> uint64_t {
        uint64_t& w = where[2 * id + (rc ? 1 : 0)];
        if (w == ~0ull) {
          const std::string& s = rc ? sequences[id]->reverse_complement() : sequences[id]->data();
          w = seqs.size();
          seqs.insert(seqs.end(), s.begin(), s.end());
        }
        return w;
      };
      std::vector<size_t> index;
      std::vector<uint64_t> q_off, t_off;
      std::vector<uint32_t> q_len, t_len, t_begin, q_start;
      for (size_t i = first; i < last; ++i) {
        const Overlap& o = *overlaps[i];
        if (!o.cigar_.empty() || !o.breaking_points_.empty()) continue;  // SAM input / already done
        index.push_back(i);
        // the substrings of overlap.cpp:195-199
        q_off.push_back(o.strand_ ? place(o.q_id_, true) + (o.q_length_ - o.q_end_) : place(o.q_id_, false) + o.q_begin_);
        q_len.push_back(o.q_end_ - o.q_begin_);
        t_off.push_back(place(o.t_id_, false) + o.t_begin_);
        t_len.push_back(o.t_end_ - o.t_begin_);
        t_begin.push_back(o.t_begin_);
        q_start.push_back(o.strand_ ? o.q_length_ - o.q_end_ : o.q_begin_);  // overlap.cpp:238
      }
      if (index.empty()) continue;
      vga_batch b;
      b.seqs = seqs.data();
      b.seqs_len = seqs.size();
      b.n = static_cast<uint32_t>(index.size());
      b.q_off = q_off.data();
      b.q_len = q_len.data();
      b.t_off = t_off.data();
      b.t_len = t_len.data();
      if (cut_on_device) {
        vga_cut c;
        c.t_begin = t_begin.data();
        c.q_start = q_start.data();
        c.window_length = window_length;
        vga_breaks r;
        if (vga_break(h, &b, &c, &r, nullptr) != VGA_OK) {
          *error = vga_last_error();
          break;
        }
        for (size_t x = 0; x < index.size(); ++x) {
          Overlap& o = *overlaps[index[x]];
          for (uint64_t p = r.points_off[x]; p < r.points_off[x + 1]; ++p) {
            o.breaking_points_.emplace_back(r.points[4 * p], r.points[4 * p + 1]);
            o.breaking_points_.emplace_back(r.points[4 * p + 2], r.points[4 * p + 3]);
          }
          // an alignment without a single match column has no breaking points: leave an empty alignment behind so
          // that the reference's loop has nothing to align for it either
          if (o.breaking_points_.empty()) o.cigar_ = "0M";
        }
      } else {
        vga_result r;
        if (vga_align(h, &b, &r, nullptr) != VGA_OK) {
          *error = vga_last_error();
          break;
        }
        for (size_t x = 0; x < index.size(); ++x) overlaps[index[x]]->cigar_ = r.cigar + r.cigar_off[x];
      }
    }
    vga_destroy(h);
  }
};

B200Polisher::B200Polisher(std::unique_ptr<bioparser::Parser<Sequence>> sparser,
                           std::unique_ptr<bioparser::Parser<Overlap>> oparser,
                           std::unique_ptr<bioparser::Parser<Sequence>> tparser, PolisherType type, bool haplotype,
                           double min_confidence, double min_support, uint32_t num_prune, uint32_t window_length,
                           double quality_threshold, double error_threshold, bool trim, int8_t match, int8_t mismatch,
                           int8_t gap, uint32_t num_threads, std::vector<int> devices)
    : Polisher(std::move(sparser), std::move(oparser), std::move(tparser), type, haplotype, min_confidence, min_support,
               num_prune, window_length, quality_threshold, error_threshold, trim, match, mismatch, gap, num_threads),
      match_(match), mismatch_(mismatch), gap_(gap), num_threads_(num_threads), devices_(std::move(devices)),
      align_on_gpu_(false), cut_on_gpu_(false), tile_in_binding_(true) {
  if (devices_.empty()) devices_.push_back(0);
  if (const char* t = std::getenv("VECHAT_B200_TILING")) tile_in_binding_ = t[0] != '0';
  const char* env = std::getenv("VECHAT_B200_ALIGN");
  align_on_gpu_ = env != nullptr && env[0] != '\0' && env[0] != '0';
  cut_on_gpu_ = align_on_gpu_ && std::strcmp(env, "cigar") != 0;
}

B200Polisher::~B200Polisher() {
  for (auto& t : closers_)
    if (t.joinable()) t.join();
}

void B200Polisher::find_overlap_breaking_points(std::vector<std::unique_ptr<Overlap>>& overlaps) {
  if (align_on_gpu_) {
    logger_->log();
    CUDABatchAligner::align(overlaps, sequences_, devices_, window_length_, cut_on_gpu_);
    logger_->log("[racon::B200Polisher::find_overlap_breaking_points] aligned overlaps on the GPU");
  }
  Polisher::find_overlap_breaking_points(overlaps);  // cuts the breaking points; edlib only where cigar_ is empty
  if (tile_in_binding_) build_tiles(overlaps);
}

// Polisher::initialize's last loop (src/polisher.cpp:408-462) walks every overlap serially and sums the qualities of
// every layer on one thread — 13 s of the 10^6-window job.  Here the same decisions are taken per overlap on host
// threads (length filter :414-417, mean-quality filter :419-433 — a sum of small integers, exact in a double in any
// order — window and positions :435-461), the layers are grouped per window in overlap order (what the serial loop's
// add_layer calls produce), and the overlaps' breaking points are cleared, so that the reference's loop only counts
// targets_coverages_ and creates the (layer-less) windows.  polish() packs from these tiles.
void B200Polisher::build_tiles(std::vector<std::unique_ptr<Overlap>>& overlaps) {
  logger_->log();
  const uint32_t wl = window_length_;
  uint64_t max_t = 0;
  for (const auto& o : overlaps) max_t = std::max<uint64_t>(max_t, o->t_id());
  // first window of every target up to the last one an overlap names (targets are the first sequences_, :389-404)
  std::vector<uint64_t> first_window(max_t + 2, 0);
  for (uint64_t i = 0; i <= max_t && i < sequences_.size(); ++i)
    first_window[i + 1] = first_window[i] + (sequences_[i]->data().size() + wl - 1) / wl;
  const uint64_t n_win = first_window[max_t + 1];
  struct Rec {
    uint64_t window;
    B200TileLayer layer;
  };
  const size_t no = overlaps.size();
  const unsigned nt = no < 256 ? 1u : std::max(1u, std::min<unsigned>(num_threads_, 32));
  std::vector<std::vector<Rec>> part(nt);
  auto run = [&](unsigned t) {
    std::vector<Rec>& out = part[t];
    for (size_t i = no * t / nt; i < no * (t + 1) / nt; ++i) {
      const Overlap& o = *overlaps[i];
      const auto& sequence = sequences_[o.q_id()];
      const auto& bp = o.breaking_points();
      for (uint32_t j = 0; j + 1 < bp.size(); j += 2) {
        const uint32_t qb = bp[j].second, qe = bp[j + 1].second;
        if (qe - qb < 0.02 * wl) continue;
        const bool has_q = !sequence->quality().empty() || !sequence->reverse_quality().empty();
        if (has_q) {
          const auto& quality = o.strand() ? sequence->reverse_quality() : sequence->quality();
          double average_quality = 0;
          for (uint32_t k = qb; k < qe; ++k) average_quality += static_cast<uint32_t>(quality[k]) - 33;
          average_quality /= qe - qb;
          if (average_quality < quality_threshold_) continue;
        }
        const uint32_t window_start = (bp[j].first / wl) * wl;
        Rec r;
        r.window = first_window[o.t_id()] + bp[j].first / wl;
        r.layer.data = o.strand() ? &(sequence->reverse_complement()[qb]) : &(sequence->data()[qb]);
        r.layer.length = qe - qb;
        r.layer.quality = o.strand() ? (sequence->reverse_quality().empty() ? nullptr : &(sequence->reverse_quality()[qb]))
                                     : (sequence->quality().empty() ? nullptr : &(sequence->quality()[qb]));
        r.layer.begin = bp[j].first - window_start;
        r.layer.end = bp[j + 1].first - window_start - 1;
        out.push_back(r);
      }
    }
  };
  if (nt == 1) {
    run(0);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(run, t);
    for (auto& x : th) x.join();
  }
  // group per window, keeping the overlap order (the parts are consecutive ranges of the overlaps)
  tile_first_.assign(n_win + 1, 0);
  for (const auto& p : part)
    for (const Rec& r : p) ++tile_first_[r.window + 1];
  for (uint64_t w = 0; w < n_win; ++w) tile_first_[w + 1] += tile_first_[w];
  tile_layers_.resize(tile_first_[n_win]);
  {
    std::vector<uint64_t> cursor(tile_first_.begin(), tile_first_.end() - 1);
    for (const auto& p : part)
      for (const Rec& r : p) tile_layers_[cursor[r.window]++] = r.layer;
  }
  for (auto& o : overlaps) CUDABatchAligner::clear_breaking_points(*o);
  logger_->log("[racon::B200Polisher::find_overlap_breaking_points] tiled the windows on host threads");
}

void B200Polisher::polish(std::vector<std::unique_ptr<Sequence>>& dst, bool drop_unpolished_sequences) {
  logger_->log();

  vgc_params prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.match = match_;
  prm.mismatch = mismatch_;
  prm.gap = gap_;
  prm.haplotype = haplotype_ ? 1 : 0;
  prm.trim = trim_ ? 1 : 0;
  prm.num_prune = num_prune_;
  prm.min_confidence = min_confidence_;
  prm.min_support = min_support_;

  const size_t n = windows_.size();
  std::vector<uint8_t> polished(n, 0);
  CUDABatchProcessor::Tiles tiles;
  tiles.layers = &tile_layers_;
  tiles.first = &tile_first_;

  // The windows are cut into batches (<= 64 k windows / 1 GB of bases per vgc call) that form ONE queue for all
  // devices: every device has its own host thread + vgc_handle (the model of the legacy path,
  // cudapolisher.cpp:229-241,255-277, whose threads pull windows under a mutex) and pulls the next batch when it
  // starts working on the one it holds — the multi-GPU form of scripts/vechat's --split chunk loop (:300-361), without
  // a process launch and a FASTA round trip per chunk.  Windows are independent, so a batch may end anywhere; with
  // several devices the batches are made small enough that every device gets about eight (load balance at the tail:
  // a device that pulls the last batch late finishes at most one batch after the others).
  const size_t nd = devices_.size();
  uint64_t kBatchBytes = 1ull << 30;  // bases per vgc call
  size_t kBatchWindows = 1u << 16;    // windows per vgc call (VECHAT_B200_BATCH_WINDOWS: smaller batches)
  if (nd > 1) kBatchWindows = std::min<size_t>(kBatchWindows, std::max<size_t>(8192, (n + 8 * nd - 1) / (8 * nd)));
  if (const char* env = std::getenv("VECHAT_B200_BATCH_WINDOWS")) {
    const long v = std::strtol(env, nullptr, 10);
    if (v > 0) kBatchWindows = static_cast<size_t>(v);
  }
  // (batches of one size: growing batches were tried to shorten the packing of the first one, but every growth of
  //  the engine's scratch frees tens of GB of device memory, which costs more than it saves — 2.2 s measured)
  std::vector<std::pair<size_t, size_t>> batches;
  for (size_t first = 0; first < n;) {
    const size_t cap = kBatchWindows;
    size_t last = first;
    uint64_t b = 0;
    while (last < n && last - first < cap && (last == first || b < kBatchBytes))
      b += CUDABatchProcessor::bytes(*windows_[last], last, tiles), ++last;
    batches.emplace_back(first, last);
    first = last;
  }
  const bool verbose = std::getenv("VECHAT_B200_VERBOSE") != nullptr;
  const auto t_polish0 = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
  };
  std::atomic<size_t> queue_head(0);
  std::vector<vgc_handle> handles(nd, nullptr);
  std::vector<size_t> taken(nd, 0);
  std::vector<std::string> errors(nd);
  std::atomic<bool> failed(false);
  auto run_device = [&](size_t d) {
    vgc_handle h = nullptr;
    if (vgc_create(&h, devices_[d], &prm) != VGC_OK) {
      errors[d] = vgc_last_error();
      failed.store(true);
      return;
    }
    const unsigned pack_threads = std::max<unsigned>(1, num_threads_ / nd);
    // two-deep pipeline over vgc_submit / vgc_collect: while the device works on batch x, a helper thread packs the
    // batch this device pulled next (the windows only borrow their bytes, so packing is a pure read of
    // Polisher::sequences_) and submits it — its host preparation and H2D overlap the kernels of batch x
    Packed cur, next;
    double pack_ms = 0.0;
    int rc_submit = VGC_OK;
    std::string submit_err;
    size_t x = queue_head.fetch_add(1);
    if (x < batches.size()) {
      CUDABatchProcessor::pack(windows_, batches[x].first, batches[x].second, &cur, pack_threads, tiles);
      const vgc_batch b0 = cur.view();
      if (vgc_submit(h, &b0) != VGC_OK) {
        errors[d] = vgc_last_error();
        failed.store(true);
        vgc_destroy(h);
        return;
      }
    }
    // result buffers of two batches (no zero-filling: the engine writes every byte it reports): while the device
    // works on batch x + 1, a helper thread stores the consensuses of batch x into their windows
    struct Out {
      std::unique_ptr<uint8_t[]> cons;
      size_t cap = 0;
      std::vector<uint64_t> off;
    } outs[2];
    int cur_out = 0;
    std::thread storer;
    while (x < batches.size() && !failed.load()) {
      const size_t first = batches[x].first, last = batches[x].second;
      ++taken[d];
      const size_t xn = queue_head.fetch_add(1);
      std::thread packer;
      if (xn < batches.size()) {
        packer = std::thread([&, xn] {
          const auto t_p = std::chrono::steady_clock::now();
          CUDABatchProcessor::pack(windows_, batches[xn].first, batches[xn].second, &next, pack_threads, tiles);
          pack_ms = since(t_p);
          const vgc_batch bn = next.view();
          rc_submit = vgc_submit(h, &bn);
          if (rc_submit != VGC_OK) submit_err = vgc_last_error();
        });
      }
      const vgc_batch batch = cur.view();
      Out& o = outs[cur_out];
      const size_t bound = vgc_result_bound(&batch);
      if (o.cap < bound) {
        o.cons.reset(new uint8_t[bound]);
        o.cap = bound;
      }
      o.off.resize(batch.n_windows + 1);
      vgc_result r = {o.cons.get(), bound, o.off.data(), polished.data() + first};
      const auto t_c = std::chrono::steady_clock::now();
      const int rc = vgc_collect(h, &r, nullptr);  // no CPU fallback
      const double collect_ms = since(t_c);
      if (rc != VGC_OK) errors[d] = vgc_last_error();
      if (packer.joinable()) packer.join();
      if (storer.joinable()) storer.join();
      if (rc == VGC_OK && rc_submit != VGC_OK) errors[d] = submit_err;
      if (rc != VGC_OK || rc_submit != VGC_OK) {
        failed.store(true);
        break;
      }
      if (verbose)
        fprintf(stderr, "[racon::B200Polisher::polish] device %d batch %zu (%zu windows): collect %.1f ms, packing of the next batch %.1f ms, joined at %.1f ms\n",
                devices_[d], x, last - first, collect_ms, pack_ms, since(t_polish0));
      // pack only reads sequences_ of the windows of ITS batch, store only writes consensus_ of the windows of batch
      // x: disjoint by construction
      storer = std::thread([this, &o, first, last] {
        for (size_t i = first; i < last; ++i)
          CUDABatchProcessor::store(*windows_[i], o.cons.get() + o.off[i - first], o.off[i - first + 1] - o.off[i - first]);
      });
      cur_out ^= 1;
      std::swap(cur, next);
      x = xn;
    }
    if (storer.joinable()) storer.join();
    handles[d] = h;  // destroyed behind the stitch: freeing ~100 GB of device scratch takes half a second
  };
  if (nd == 1) {
    run_device(0);
  } else {
    std::vector<std::thread> th;
    for (size_t d = 0; d < nd; ++d) th.emplace_back(run_device, d);
    for (auto& t : th) t.join();
  }
  for (size_t d = 0; d < nd; ++d)
    if (!errors[d].empty()) die("B200Polisher::polish", errors[d].c_str());
  if (verbose) fprintf(stderr, "[racon::B200Polisher::polish] all batches collected at %.1f ms\n", since(t_polish0));
  for (vgc_handle h : handles)
    if (h) closers_.emplace_back([h] { vgc_destroy(h); });
  if (nd > 1) {
    std::string msg = "[racon::B200Polisher::polish] " + std::to_string(batches.size()) + " batches from one queue:";
    for (size_t d = 0; d < nd; ++d) msg += " device " + std::to_string(devices_[d]) + " took " + std::to_string(taken[d]);
    fprintf(stderr, "%s\n", msg.c_str());
  }

  // In-order stitch per target with the reference's header tags (polisher.cpp:520-546): a target ends where the
  // next window has rank 0.
  // (the progress bar is the reference's own, polisher.cpp:524,549-558: one step per 1/20 of the windows stitched)
  const uint64_t logger_step = n / 20;
  auto bar_after = [&](size_t k) {  // k = index of the window just stitched
    if (logger_step != 0 && (k + 1) % logger_step == 0 && (k + 1) / logger_step < 20)
      logger_->bar("[racon::Polisher::polish] generating consensus");
  };
  // (targets are independent: their records are built by num_threads_ host threads over contiguous ranges of
  //  targets and appended in order afterwards — on 10^6 windows the serial loop was 12 % of the polish stage)
  std::vector<size_t> tstart;
  for (size_t k = 0; k < n; ++k)
    if (k == 0 || windows_[k]->rank() == 0) tstart.push_back(k);
  tstart.push_back(n);
  const size_t n_targets = tstart.size() - 1;
  std::vector<std::unique_ptr<Sequence>> records(n_targets);
  auto stitch_range = [&](size_t t0, size_t t1) {
    for (size_t t = t0; t < t1; ++t) {
      const size_t i0 = tstart[t], j = tstart[t + 1];
      std::string data;
      uint32_t good = 0;
      for (size_t k = i0; k < j; ++k) {
        good += polished[k] ? 1 : 0;
        data += windows_[k]->consensus();
      }
      const Window& tail = *windows_[j - 1];
      const double ratio = good / static_cast<double>(tail.rank() + 1);
      if (!drop_unpolished_sequences || ratio > 0) {
        std::string name = sequences_[tail.id()]->name();
        if (type_ == PolisherType::kF) name += "r";
        name += " LN:i:" + std::to_string(data.size());
        name += " RC:i:" + std::to_string(targets_coverages_[tail.id()]);
        name += " XC:f:" + std::to_string(ratio);
        records[t] = createSequence(name, data);
      }
      for (size_t k = i0; k < j; ++k) windows_[k].reset();
    }
  };
  {
    const size_t nt = n_targets < 64 ? 1 : std::max<size_t>(1, std::min<size_t>(num_threads_, 32));
    if (nt == 1) {
      stitch_range(0, n_targets);
    } else {
      std::vector<std::thread> th;
      for (size_t t = 0; t < nt; ++t) th.emplace_back(stitch_range, n_targets * t / nt, n_targets * (t + 1) / nt);
      for (auto& x : th) x.join();
    }
  }
  for (size_t t = 0; t < n_targets; ++t) {
    if (records[t]) dst.emplace_back(std::move(records[t]));
    for (size_t k = tstart[t]; k < tstart[t + 1]; ++k) bar_after(k);
  }
  if (verbose) fprintf(stderr, "[racon::B200Polisher::polish] stitched at %.1f ms\n", since(t_polish0));
  if (logger_step != 0) {
    logger_->bar("[racon::Polisher::polish] generating consensus");
  } else {
    logger_->log("[racon::Polisher::polish] generated consensus");
  }

  std::vector<std::shared_ptr<Window>>().swap(windows_);
  std::vector<std::unique_ptr<Sequence>>().swap(sequences_);
}

namespace {

bool ends_with(const std::string& s, const char* suffix) {
  const size_t k = std::strlen(suffix);
  return s.size() >= k && s.compare(s.size() - k, k, suffix) == 0;
}
bool has_ext(const std::string& path, std::initializer_list<const char*> exts) {
  for (const char* e : exts)
    if (ends_with(path, e) || ends_with(path, (std::string(e) + ".gz").c_str())) return true;
  return false;
}
// the extension rules of createPolisher (polisher.cpp:85-138), same messages
std::unique_ptr<bioparser::Parser<Sequence>> sequence_parser(const std::string& path) {
  if (has_ext(path, {".fasta", ".fna", ".fa"})) return bioparser::Parser<Sequence>::Create<bioparser::FastaParser>(path);
  if (has_ext(path, {".fastq", ".fq"})) return bioparser::Parser<Sequence>::Create<bioparser::FastqParser>(path);
  std::fprintf(stderr, "[racon::createPolisher] error: file %s has unsupported format extension (valid extensions: "
               ".fasta, .fasta.gz, .fna, .fna.gz, .fa, .fa.gz, .fastq, .fastq.gz, .fq, .fq.gz)!\n", path.c_str());
  std::exit(1);
}
std::unique_ptr<bioparser::Parser<Overlap>> overlap_parser(const std::string& path) {
  if (has_ext(path, {".mhap"})) return bioparser::Parser<Overlap>::Create<bioparser::MhapParser>(path);
  if (has_ext(path, {".paf"})) return bioparser::Parser<Overlap>::Create<bioparser::PafParser>(path);
  if (has_ext(path, {".sam"})) return bioparser::Parser<Overlap>::Create<bioparser::SamParser>(path);
  std::fprintf(stderr, "[racon::createPolisher] error: file %s has unsupported format extension (valid extensions: "
               ".mhap, .mhap.gz, .paf, .paf.gz, .sam, .sam.gz)!\n", path.c_str());
  std::exit(1);
}

}  // namespace

std::unique_ptr<Polisher> createPolisherB200(const std::string& sequences_path, const std::string& overlaps_path,
                                             const std::string& target_path, PolisherType type, bool haplotype,
                                             double min_confidence, double min_support, uint32_t num_prune,
                                             uint32_t window_length, double quality_threshold, double error_threshold,
                                             bool trim, int8_t match, int8_t mismatch, int8_t gap, uint32_t num_threads,
                                             uint32_t cuda_batches, bool cuda_banded_alignment,
                                             uint32_t cudaaligner_batches, uint32_t cudaaligner_band_width) {
  std::vector<int> devices;
  if (const char* env = std::getenv("VECHAT_B200_DEVICES")) {
    for (const char* p = env; *p;) {
      char* e = nullptr;
      const long v = std::strtol(p, &e, 10);
      if (e == p) die("createPolisher", "VECHAT_B200_DEVICES must be a comma-separated list of device ordinals");
      devices.push_back(static_cast<int>(v));
      p = (*e == ',') ? e + 1 : e;
    }
  } else {
    for (uint32_t d = 0; d < cuda_batches; ++d) devices.push_back(static_cast<int>(d));
  }
  if (devices.empty())  // CPU run requested: the reference's own factory, untouched
    return createPolisher(sequences_path, overlaps_path, target_path, type, haplotype, min_confidence, min_support,
                          num_prune, window_length, quality_threshold, error_threshold, trim, match, mismatch, gap,
                          num_threads, 0, cuda_banded_alignment, cudaaligner_batches, cudaaligner_band_width);

  if (type != PolisherType::kC && type != PolisherType::kF) die("createPolisher", "invalid polisher type!");
  if (window_length == 0) die("createPolisher", "invalid window length!");
  // engine limits, checked before initialize() is paid for (include/vgc.h: a layer may hold VGC_MAX_LAYER_LEN bases;
  // layers are about as long as the window, insertions included)
  if (window_length > VGC_MAX_LAYER_LEN / 2)
    die("createPolisher", "window length (-w) above 8191 is beyond the B200 engine's layer limit (16383 bases); "
                          "lower -w or run the CPU path (unset VECHAT_B200_DEVICES)!");
  if (gap > 0) die("createPolisher", "gap penalty (-g) must be non-positive!");
  {
    std::string list;
    for (int d : devices) list += (list.empty() ? "" : ",") + std::to_string(d);
    std::fprintf(stderr, "[racon::createPolisherB200] consensus on B200 device(s) %s (libvgc %s)\n", list.c_str(),
                 vgc_version());
  }
  auto sparser = sequence_parser(sequences_path);
  auto oparser = overlap_parser(overlaps_path);
  auto tparser = sequence_parser(target_path);
  return std::unique_ptr<Polisher>(new B200Polisher(std::move(sparser), std::move(oparser), std::move(tparser), type,
                                                    haplotype, min_confidence, min_support, num_prune, window_length,
                                                    quality_threshold, error_threshold, trim, match, mismatch, gap,
                                                    num_threads, devices));
}

}  // namespace racon

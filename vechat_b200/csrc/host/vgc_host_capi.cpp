// vgc_host_capi.cpp — a flat C entry point over vgc_host.hpp so that the test-suite (Python, ctypes) can drive the
// C++ host side exactly as VeChat would: createWindow / add_layer per window, then B200Polisher::polish.
// Built into vechat_b200/lib/libvgchost.so (links libvgc.so).
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "vgc_host.hpp"

using namespace vgc_host;

namespace {
thread_local std::string g_msg;
struct Thrown : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// Windows borrow their bytes (window.hpp:74-76); in VeChat they point into NUL-terminated std::strings owned by
// Polisher::sequences_.  `store` plays that role here, so that the C-string compare of window.cpp:223 reads
// exactly what it would read there.
std::vector<std::shared_ptr<Window>> build_windows(const vgc_batch* b, const uint64_t* win_id, const uint32_t* win_rank,
                                                   std::vector<std::string>* store) {
  std::vector<std::shared_ptr<Window>> w;
  store->clear();
  store->reserve(2 * static_cast<size_t>(b->n_layers));
  auto keep = [&](const uint8_t* p, uint32_t n) -> const char* {
    store->emplace_back(reinterpret_cast<const char*>(p), n);
    return store->back().c_str();
  };
  for (uint32_t i = 0; i < b->n_windows; ++i) {
    const uint32_t f = b->win_first[i], l = b->win_first[i + 1];
    const uint32_t blen = static_cast<uint32_t>(b->seq_off[f + 1] - b->seq_off[f]);
    const char* bb = keep(b->bases + b->seq_off[f], blen);
    const char* bq = keep(b->quals + b->seq_off[f], blen);
    if (!(b->win_flags[i] & VGC_WIN_DUMMY_QUAL) && store->back() == std::string(blen, '!')) {
      // an all-'!' quality that the caller does NOT flag as "dummy" is VeChat's dummy_quality_ seen from a window
      // shorter than window_length (polisher.cpp:181, window.cpp:223): the string goes on beyond the window
      store->back().push_back('!');
      bq = store->back().c_str();
    }
    auto win = createWindow(win_id[i], win_rank[i], (b->win_flags[i] & VGC_WIN_TGS) ? WindowType::kTGS : WindowType::kNGS,
                            bb, blen, bq, blen);
    for (uint32_t k = f + 1; k < l; ++k) {
      const uint32_t len = static_cast<uint32_t>(b->seq_off[k + 1] - b->seq_off[k]);
      const char* sq = keep(b->bases + b->seq_off[k], len);
      const char* qq = b->has_qual[k] ? keep(b->quals + b->seq_off[k], len) : nullptr;
      win->add_layer(sq, len, qq, len, b->begin[k], b->end[k]);
    }
    w.push_back(win);
  }
  return w;
}
}  // namespace

extern "C" {

const char* vgch_last_error(void) { return g_msg.c_str(); }

// Pack round-trip (no device work): windows are rebuilt through createWindow/add_layer from `b`, packed again by
// BatchPacker, and the packed arrays are compared with what add_layer keeps.  Returns the number of layers packed,
// or -1 on an error (message in vgch_last_error).
long vgch_pack_roundtrip(const vgc_batch* b, const uint64_t* win_id, const uint32_t* win_rank, uint64_t* bases_hash) {
  error_handler() = [](const std::string& m) { throw Thrown(m); };
  try {
    std::vector<std::string> store;
    auto w = build_windows(b, win_id, win_rank, &store);
    PackedBatch p;
    BatchPacker::pack(w, 0, w.size(), &p);
    uint64_t h = 1469598103934665603ull;
    for (uint8_t c : p.bases) h = (h ^ c) * 1099511628211ull;
    for (uint32_t x : p.begin) h = (h ^ x) * 1099511628211ull;
    for (uint32_t x : p.end) h = (h ^ x) * 1099511628211ull;
    for (uint8_t x : p.win_flags) h = (h ^ x) * 1099511628211ull;
    for (uint8_t x : p.has_qual) h = (h ^ x) * 1099511628211ull;
    *bases_hash = h;
    return static_cast<long>(p.begin.size());
  } catch (const Thrown& e) {
    g_msg = e.what();
    return -1;
  }
}

// Full polish through the C++ host side.  Output: FASTA text (">name tags\nsequence\n" per target) into `out`.
// Returns the number of bytes written, -1 on error, -2 if `out` is too small.
long vgch_polish_fasta(const vgc_batch* b, const vgc_params* prm, const uint64_t* win_id, const uint32_t* win_rank,
                       const char* const* names, const uint32_t* coverages, uint32_t n_targets, int fragment_correction,
                       int drop_unpolished, int device, char* out, uint64_t out_cap) {
  error_handler() = [](const std::string& m) { throw Thrown(m); };
  try {
    std::vector<std::string> store;
    auto w = build_windows(b, win_id, win_rank, &store);
    std::vector<std::string> nm(names, names + n_targets);
    std::vector<uint32_t> cov(coverages, coverages + n_targets);
    B200Polisher pol(fragment_correction ? PolisherType::kF : PolisherType::kC, prm->haplotype != 0, prm->min_confidence,
                     prm->min_support, prm->num_prune, prm->trim != 0, prm->match, prm->mismatch, prm->gap, device);
    std::vector<std::unique_ptr<Sequence>> dst;
    pol.polish(w, nm, cov, dst, drop_unpolished != 0);
    std::string text;
    for (const auto& s : dst) text += ">" + s->name + "\n" + s->data + "\n";  // src/main.cpp:176-178
    if (text.size() > out_cap) return -2;
    std::memcpy(out, text.data(), text.size());
    return static_cast<long>(text.size());
  } catch (const Thrown& e) {
    g_msg = e.what();
    return -1;
  }
}

}  // extern "C"

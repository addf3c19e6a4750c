// b200polisher.hpp — the VeChat-side binding of include/vgc.h: a racon::Polisher subclass that overrides polish()
// only (src/polisher.hpp:55-58).  initialize() — parsing, overlap filtering, edlib breakpoints, window tiling,
// src/polisher.cpp:199-462 — is inherited unchanged, so the engine sees the reference's own window tilings.
//
// This file is compiled TOGETHER WITH the unmodified reference sources (it includes the reference's headers from
// the reference tree; nothing of the reference is copied here).  It plays the role src/cuda/cudapolisher.{hpp,cpp}
// play for the legacy GenomeWorks path, minus the CPU fallback (cudapolisher.cpp:355-379): a failing engine call
// ends the process with the reference's "[racon::...] error: ..." + exit(1) convention.
#ifndef VGC_B200POLISHER_HPP_
#define VGC_B200POLISHER_HPP_

#include <cstdint>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "polisher.hpp"  // reference: src/polisher.hpp

namespace racon {

// One layer of a window as the binding's own tiling records it: what Window::add_layer would have
// been given (src/polisher.cpp:437-461) — bytes borrowed from Polisher::sequences_, positions inside the window.
struct B200TileLayer {
  const char* data;
  const char* quality;  // nullptr: the read has no qualities
  uint32_t length;
  uint32_t begin, end;
};

class B200Polisher : public Polisher {
 public:
  // same argument list as the protected Polisher constructor (src/polisher.hpp:72-79) + the devices to use
  B200Polisher(std::unique_ptr<bioparser::Parser<Sequence>> sparser, std::unique_ptr<bioparser::Parser<Overlap>> oparser,
               std::unique_ptr<bioparser::Parser<Sequence>> tparser, PolisherType type, bool haplotype,
               double min_confidence, double min_support, uint32_t num_prune, uint32_t window_length,
               double quality_threshold, double error_threshold, bool trim, int8_t match, int8_t mismatch, int8_t gap,
               uint32_t num_threads, std::vector<int> devices);
  ~B200Polisher() override;

  void polish(std::vector<std::unique_ptr<Sequence>>& dst, bool drop_unpolished_sequences) override;

 protected:
  // Opt-in (VECHAT_B200_ALIGN=1): the overlaps are aligned AND cut into breaking points on the GPU (vga_break,
  // include/vga.h) before the reference's own loop runs, which then finds breaking_points_ filled and returns at
  // src/overlap.cpp:187-189.  VECHAT_B200_ALIGN=cigar: only the CIGARs come from the GPU (vga_align) and the
  // reference cuts them itself (overlap.cpp:191-203).  Off by default because the aligner's choice among equally
  // good alignments is not edlib's.
  void find_overlap_breaking_points(std::vector<std::unique_ptr<Overlap>>& overlaps) override;

 private:
  int8_t match_, mismatch_, gap_;  // the base class hands them to spoa and forgets them (polisher.cpp:186-190)
  uint32_t num_threads_;
  std::vector<int> devices_;
  bool align_on_gpu_, cut_on_gpu_;
  // Tiling taken over from Polisher::initialize's serial loop (src/polisher.cpp:408-462; VECHAT_B200_TILING=0 leaves
  // it to the reference): the layers of every window, window by window in overlap order, built on host threads
  // inside find_overlap_breaking_points; polish() packs from here instead of from Window::sequences_.
  void build_tiles(std::vector<std::unique_ptr<Overlap>>& overlaps);
  bool tile_in_binding_;
  std::vector<B200TileLayer> tile_layers_;
  std::vector<uint64_t> tile_first_;  // [windows with layers + 1]; empty: the reference tiled
  std::vector<std::thread> closers_;  // free the engines' device scratch behind the stitch and the output (joined in the destructor)
};

// Drop-in for racon::createPolisher (same signature, src/polisher.hpp:42-49).  Returns a B200Polisher when the
// environment names devices (VECHAT_B200_DEVICES="0" or "0,1,2,3") or cuda_batches > 0 (then devices 0..n-1);
// otherwise forwards to the reference's createPolisher untouched.
std::unique_ptr<Polisher> createPolisherB200(const std::string& sequences_path, const std::string& overlaps_path,
                                             const std::string& target_path, PolisherType type, bool haplotype,
                                             double min_confidence, double min_support, uint32_t num_prune,
                                             uint32_t window_length, double quality_threshold, double error_threshold,
                                             bool trim, int8_t match, int8_t mismatch, int8_t gap, uint32_t num_threads,
                                             uint32_t cuda_batches = 0, bool cuda_banded_alignment = false,
                                             uint32_t cudaaligner_batches = 0, uint32_t cudaaligner_band_width = 0);

}  // namespace racon
#endif

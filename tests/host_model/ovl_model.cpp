// TEST-ONLY: vechat_b200/csrc/ovl_core.h (the arithmetic of the overlap-alignment kernel) driven on the host — the
// loop over a wavefront's diagonals that the kernel spreads over a CTA's threads runs serially here (cells of one
// wavefront only read the previous one).  Produces the CIGAR the kernel + vga_align would.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "ovl_core.h"

extern "C" int ovl_model_align(const uint8_t* q, int32_t m, const uint8_t* t, int32_t n, char* out, uint64_t cap,
                               uint64_t* cells) {
  std::vector<int32_t> arena;
  int32_t D = -1;
  for (int32_t d = 0; D < 0; ++d) {
    arena.resize(ovl::wf_cells(d), 0x5a5a5a5a);  // poison: every cell read must have been written or range-checked
    const ovl::Front prev = ovl::wf_front(arena.data(), d > 0 ? d - 1 : 0, m, n);
    const ovl::Front cur = ovl::wf_front(arena.data(), d, m, n);
    for (int32_t k = cur.lo; k <= cur.hi; ++k) {
      const int32_t i = ovl::wf_cell(prev, q, t, m, n, d, k);
      arena[ovl::wf_index(d, k)] = i;
      if (i == m && k == n - m) D = d;
    }
  }
  std::vector<uint32_t> runs(static_cast<size_t>(m) + n + 2);
  const uint32_t nr = ovl::wf_traceback(arena.data(), m, n, D, runs.data());
  std::string s;
  for (uint32_t x = nr; x-- > 0;) s += std::to_string(runs[x] >> 2) + "MID?"[runs[x] & 3];
  if (s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  if (cells) *cells = ovl::wf_cells(D);
  return D;
}

// Alignment + breaking points, as the kernel's cut mode computes them (ovl::wf_cut on the reverse-order runs).
extern "C" int ovl_model_break(const uint8_t* q, int32_t m, const uint8_t* t, int32_t n, uint32_t t_begin,
                               uint32_t q_start, uint32_t window_length, uint32_t* out, uint32_t max_pairs) {
  std::vector<int32_t> arena;
  int32_t D = -1;
  for (int32_t d = 0; D < 0; ++d) {
    arena.resize(ovl::wf_cells(d), 0x5a5a5a5a);
    const ovl::Front prev = ovl::wf_front(arena.data(), d > 0 ? d - 1 : 0, m, n);
    const ovl::Front cur = ovl::wf_front(arena.data(), d, m, n);
    for (int32_t k = cur.lo; k <= cur.hi; ++k) {
      const int32_t i = ovl::wf_cell(prev, q, t, m, n, d, k);
      arena[ovl::wf_index(d, k)] = i;
      if (i == m && k == n - m) D = d;
    }
  }
  std::vector<uint32_t> runs(static_cast<size_t>(m) + n + 2);
  const uint32_t nr = ovl::wf_traceback(arena.data(), m, n, D, runs.data());
  ovl::CutParams c;
  c.t_begin = t_begin;
  c.t_end = t_begin + static_cast<uint32_t>(n);
  c.q_start = q_start;
  c.window_length = window_length;
  return static_cast<int>(ovl::wf_cut(runs.data(), nr, c, out, max_pairs));
}

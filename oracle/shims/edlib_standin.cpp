// TEST INFRASTRUCTURE — see edlib.h next to this file.  Exact global (NW) unit-cost alignment with path, by
// furthest-reaching diagonals (Myers 1986 O(ND) / Ukkonen): wavefront d holds, for every diagonal k = j - i, the
// largest number of query characters i consumed with exactly d edits; all wavefronts are kept for the traceback.
// Tie-breaking (ours, fixed): a cell reached equally far by several moves prefers mismatch, then deletion
// (target only), then insertion (query only).
#include "edlib.h"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

const int32_t kNone = INT_MIN / 2;

struct Fronts {
  std::vector<int32_t> v;
  std::vector<size_t> base;  // start of wavefront d in v
  std::vector<int32_t> lo;   // lowest diagonal of wavefront d
  std::vector<int32_t> hi;
  int32_t get(int d, int32_t k) const { return (k < lo[d] || k > hi[d]) ? kNone : v[base[d] + (k - lo[d])]; }
};

// the three ways into (d, k); kNone when the move is impossible or leaves the matrix
inline void candidates(const Fronts& f, int d, int32_t k, int32_t m, int32_t n, int32_t* a, int32_t* b, int32_t* c) {
  int32_t x = f.get(d - 1, k);
  *a = (x == kNone || x + 1 > m || x + 1 + k > n) ? kNone : x + 1;  // mismatch
  x = f.get(d - 1, k + 1);
  *b = (x == kNone || x + 1 > m) ? kNone : x + 1;                    // insertion: query only
  x = f.get(d - 1, k - 1);
  *c = (x == kNone || x + k > n) ? kNone : x;                        // deletion: target only
}

}  // namespace

extern "C" {

EdlibAlignConfig edlibNewAlignConfig(int k, EdlibAlignMode mode, EdlibAlignTask task,
                                     const EdlibEqualityPair* eq, int neq) {
  EdlibAlignConfig c;
  c.k = k; c.mode = mode; c.task = task; c.additionalEqualities = eq; c.additionalEqualitiesLength = neq;
  return c;
}

EdlibAlignResult edlibAlign(const char* q, int m, const char* t, int n, const EdlibAlignConfig config) {
  EdlibAlignResult r;
  std::memset(&r, 0, sizeof(r));
  r.status = EDLIB_STATUS_ERROR;
  r.editDistance = -1;
  if (config.mode != EDLIB_MODE_NW || m < 0 || n < 0) return r;  // only what overlap.cpp:208-211 asks for
  const int32_t kf = n - m;
  Fronts f;
  int d = 0;
  for (;; ++d) {
    const int32_t lo = std::max(-d, -m), hi = std::min(d, n);
    f.base.push_back(f.v.size());
    f.lo.push_back(lo);
    f.hi.push_back(hi);
    f.v.resize(f.v.size() + (hi - lo + 1), kNone);
    int32_t* w = &f.v[f.base[d]];
    for (int32_t k = lo; k <= hi; ++k) {
      int32_t i;
      if (d == 0) {
        i = 0;
      } else {
        int32_t a, b, c;
        candidates(f, d, k, m, n, &a, &b, &c);
        i = std::max(a, std::max(b, c));
        if (i == kNone) continue;
      }
      int32_t j = i + k;
      while (i < m && j < n && q[i] == t[j]) ++i, ++j;
      w[k - lo] = i;
    }
    if (kf >= lo && kf <= hi && w[kf - lo] == m) break;
    if (config.k >= 0 && d >= config.k) return r;  // edlib: no alignment within k -> editDistance -1
  }
  r.status = EDLIB_STATUS_OK;
  r.editDistance = d;
  r.alphabetLength = 4;
  r.numLocations = 1;
  r.endLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.startLocations = static_cast<int*>(std::malloc(sizeof(int)));
  r.endLocations[0] = n - 1;
  r.startLocations[0] = 0;
  if (config.task != EDLIB_TASK_PATH) return r;

  std::vector<unsigned char> ops;
  ops.reserve(static_cast<size_t>(m) + n);
  int32_t k = kf, i = m;
  for (; d > 0; --d) {
    int32_t a, b, c;
    candidates(f, d, k, m, n, &a, &b, &c);
    const int32_t pre = std::max(a, std::max(b, c));
    for (int32_t x = i; x > pre; --x) ops.push_back(EDLIB_EDOP_MATCH);
    if (a == pre) { ops.push_back(EDLIB_EDOP_MISMATCH); i = pre - 1; }
    else if (c == pre) { ops.push_back(EDLIB_EDOP_DELETE); i = pre; k -= 1; }
    else { ops.push_back(EDLIB_EDOP_INSERT); i = pre - 1; k += 1; }
  }
  for (int32_t x = i; x > 0; --x) ops.push_back(EDLIB_EDOP_MATCH);
  std::reverse(ops.begin(), ops.end());
  r.alignmentLength = static_cast<int>(ops.size());
  r.alignment = static_cast<unsigned char*>(std::malloc(ops.size() ? ops.size() : 1));
  if (!ops.empty()) std::memcpy(r.alignment, ops.data(), ops.size());
  return r;
}

void edlibFreeAlignResult(EdlibAlignResult r) {
  std::free(r.endLocations);
  std::free(r.startLocations);
  std::free(r.alignment);
}

char* edlibAlignmentToCigar(const unsigned char* aln, int len, EdlibCigarFormat fmt) {
  static const char std_ops[4] = {'M', 'I', 'D', 'M'}, ext_ops[4] = {'=', 'I', 'D', 'X'};
  const char* tab = fmt == EDLIB_CIGAR_EXTENDED ? ext_ops : std_ops;
  std::string s;
  for (int x = 0; x < len;) {
    const char op = tab[aln[x] & 3];
    int y = x;
    while (y < len && tab[aln[y] & 3] == op) ++y;
    s += std::to_string(y - x);
    s += op;
    x = y;
  }
  char* out = static_cast<char*>(std::malloc(s.size() + 1));
  std::memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

}  // extern "C"

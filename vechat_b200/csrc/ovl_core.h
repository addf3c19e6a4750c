// ovl_core.h — the arithmetic of the overlap aligner, shared by the sm_100a kernel (ovl_align.cu) and the host
// model the CPU tests instantiate (tests/host_model/ovl_model.cpp): one wavefront cell, and the traceback.
//
// What it computes: Overlap::align_overlaps (src/overlap.cpp:205-224) asks edlib for a global unit-cost alignment
// with path.  Here: furthest-reaching diagonals.  Wavefront d holds, for every diagonal k = j - i (i query, j
// target characters consumed), the largest i reachable with exactly d edits; D = the first d whose diagonal n - m
// reaches i = m.  All wavefronts stay in HBM for the traceback: wavefront d lives at arena[d*d, (d+1)*(d+1)),
// diagonal k at d*d + d + k — D^2 cells instead of the m*n of a full matrix.
#ifndef OVL_CORE_H_
#define OVL_CORE_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OVL_HD __host__ __device__ __forceinline__
#else
#define OVL_HD inline
#endif

namespace ovl {

constexpr int32_t kNone = -(1 << 30);
enum : uint32_t { kOpM = 0, kOpI = 1, kOpD = 2 };  // CIGAR letters M (match or mismatch), I (query only), D (target only)

OVL_HD size_t wf_index(int32_t d, int32_t k) { return static_cast<size_t>(d) * d + static_cast<size_t>(d + k); }

// One wavefront, addressed by diagonal: p[k] for lo <= k <= hi (the diagonals of wavefront d that lie inside the
// m x n matrix); everything else reads as kNone.  Set up once per wavefront, not per cell.
struct Front {
  const int32_t* p;
  int32_t lo, hi;
  OVL_HD int32_t get(int32_t k) const { return (k < lo || k > hi) ? kNone : p[k]; }
};
OVL_HD Front wf_front(const int32_t* arena, int32_t d, int32_t m, int32_t n) {
  Front f;
  f.p = arena + wf_index(d, 0);
  f.lo = -d < -m ? -m : -d;
  f.hi = d < n ? d : n;
  return f;
}

// The three ways into diagonal k from the previous wavefront: a mismatch (from k), b insertion / query only (from
// k+1), c deletion / target only (from k-1); kNone when the move does not exist or leaves the matrix.
OVL_HD void wf_candidates(const Front& prev, int32_t k, int32_t m, int32_t n, int32_t* a, int32_t* b, int32_t* c) {
  int32_t x = prev.get(k);
  *a = (x == kNone || x + 1 > m || x + 1 + k > n) ? kNone : x + 1;
  x = prev.get(k + 1);
  *b = (x == kNone || x + 1 > m) ? kNone : x + 1;
  x = prev.get(k - 1);
  *c = (x == kNone || x + k > n) ? kNone : x;
}

OVL_HD int32_t max3(int32_t a, int32_t b, int32_t c) {
  const int32_t ab = a > b ? a : b;
  return ab > c ? ab : c;
}

#if defined(__CUDA_ARCH__)
// 4 bytes from any address: two aligned words + a funnel shift (reads up to 7 bytes past p: the device copy of the
// sequences is padded accordingly).
__device__ __forceinline__ uint32_t load4(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  return __funnelshift_r(__ldg(w), __ldg(w + 1), static_cast<uint32_t>(a & 3) * 8);
}
#endif

// Length of the common prefix of q[i..m) and t[j..n).
OVL_HD int32_t match_run(const uint8_t* q, const uint8_t* t, int32_t i, int32_t j, int32_t m, int32_t n) {
  const int32_t i0 = i;
#if defined(__CUDA_ARCH__)
  for (;;) {  // four characters per step
    const int32_t left = (m - i) < (n - j) ? (m - i) : (n - j);
    if (left <= 0) break;
    const uint32_t x = load4(q + i) ^ load4(t + j);
    int32_t same = x ? (__ffs(static_cast<int>(x)) - 1) >> 3 : 4;
    same = same < left ? same : left;
    i += same, j += same;
    if (same < 4) break;
  }
#else
  while (i < m && j < n && q[i] == t[j]) ++i, ++j;
#endif
  return i - i0;
}

// One cell of wavefront d (diagonal k inside the matrix): start point from the previous wavefront (none for d = 0),
// then slide along the matches.  The caller stores the result at arena[wf_index(d, k)].
//
// Fast path: take the plain maximum of the three moves first.  A move that leaves the matrix (see wf_candidates)
// always ends up being that maximum — an insertion can only leave through i > m, and a mismatch / deletion that
// leaves through j > n lies beyond every move that stays inside on the same diagonal — so "maximum inside the
// matrix" proves that no move had to be discarded, and the exact per-move checks run only next to the matrix edge.
OVL_HD int32_t wf_cell(const Front& prev, const uint8_t* q, const uint8_t* t, int32_t m, int32_t n, int32_t d, int32_t k) {
  int32_t i;
  if (d == 0) {
    i = 0;
  } else {
    int32_t x0, x1, x2;
    if (k - 1 >= prev.lo && k + 1 <= prev.hi) {  // interior diagonal: all three neighbours exist
      x0 = prev.p[k], x1 = prev.p[k + 1], x2 = prev.p[k - 1];
    } else {
      x0 = prev.get(k), x1 = prev.get(k + 1), x2 = prev.get(k - 1);
    }
    i = max3(x0 + 1, x1 + 1, x2);  // kNone + 1 stays far below zero
    if (i > m || i + k > n) {
      int32_t a, b, c;
      wf_candidates(prev, k, m, n, &a, &b, &c);
      i = max3(a, b, c);
    } else if (i < 0) {
      i = kNone;
    }
  }
  if (i != kNone) i += match_run(q, t, i, i + k, m, n);
  return i;
}

// Run-length CIGAR writer; runs come out in traceback (reverse) order: run = length << 2 | op.
struct RunWriter {
  uint32_t* out;
  uint32_t count, op, len;
  OVL_HD void init(uint32_t* o) { out = o, count = 0, op = 3, len = 0; }
  OVL_HD void add(uint32_t o, uint32_t l) {
    if (l == 0) return;
    if (o == op) {
      len += l;
    } else {
      if (len) out[count++] = len << 2 | op;
      op = o, len = l;
    }
  }
  OVL_HD uint32_t finish() {
    if (len) out[count++] = len << 2 | op;
    len = 0;
    return count;
  }
};

// Walks back from (D, n - m, m).  Among equally far predecessors: mismatch, then deletion, then insertion.
// `runs` needs room for m + n + 1 entries in the worst case.  Returns the number of runs (reverse order).
OVL_HD uint32_t wf_traceback(const int32_t* arena, int32_t m, int32_t n, int32_t D, uint32_t* runs) {
  RunWriter w;
  w.init(runs);
  int32_t k = n - m, i = m;
  for (int32_t d = D; d > 0; --d) {
    int32_t a, b, c;
    wf_candidates(wf_front(arena, d - 1, m, n), k, m, n, &a, &b, &c);
    const int32_t pre = max3(a, b, c);
    w.add(kOpM, static_cast<uint32_t>(i - pre));
    if (a == pre) {
      w.add(kOpM, 1);
      i = pre - 1;
    } else if (c == pre) {
      w.add(kOpD, 1);
      i = pre;
      k -= 1;
    } else {
      w.add(kOpI, 1);
      i = pre - 1;
      k += 1;
    }
  }
  w.add(kOpM, static_cast<uint32_t>(i));
  return w.finish();
}

// Breaking points of an alignment (Overlap::find_breaking_points_from_cigar, src/overlap.cpp:226-292): walk the
// alignment forward over target windows of `window_length` bases; for every window the alignment touches with at
// least one M column, emit (first M column of the stretch: target, query) and (one past its last M column: target + 1,
// query + 1).  The reference walks base by base; here a run is cut arithmetically at the window ends it crosses.
// Window ends (overlap.cpp:229-235): every multiple of window_length above t_begin, minus one, then t_end - 1.
struct CutParams {
  uint32_t t_begin, t_end;  // Overlap::t_begin_, t_end_ (target coordinates of the aligned substring)
  uint32_t q_start;         // query coordinate of the substring's first base: strand ? q_length - q_end : q_begin
  uint32_t window_length;
};

// `runs` as wf_traceback leaves them (reverse order, n_runs entries).  out: 4 words per breaking-point pair
// (first.t, first.q, last.t, last.q); returns the number of pairs (never more than max_pairs are written).
OVL_HD uint32_t wf_cut(const uint32_t* runs, uint32_t n_runs, const CutParams& c, uint32_t* out, uint32_t max_pairs) {
  // positions of the last consumed base; wrap-around at -1 is intended (the reference uses int32 starting at begin - 1)
  uint32_t q_ptr = c.q_start - 1, t_ptr = c.t_begin - 1;
  // next window end: the smallest multiple of window_length above t_begin, minus one, capped by t_end - 1
  uint32_t next_multiple = (c.t_begin / c.window_length + 1) * c.window_length;
  uint32_t end = next_multiple < c.t_end ? next_multiple - 1 : c.t_end - 1;
  bool found = false, ends_left = true;
  uint32_t first_t = 0, first_q = 0, last_t = 0, last_q = 0, pairs = 0;
  for (uint32_t x = n_runs; x-- > 0;) {
    const uint32_t op = runs[x] & 3;
    uint32_t len = runs[x] >> 2;
    if (op == kOpI) {
      q_ptr += len;
      continue;
    }
    // M or D: consumes target; cut at every window end inside the run
    while (ends_left && len > 0 && end - t_ptr <= len) {
      const uint32_t step = end - t_ptr;  // >= 1 bases up to and including the window end
      if (op == kOpM) {
        if (!found) found = true, first_t = t_ptr + 1, first_q = q_ptr + 1;
        q_ptr += step;
        last_t = end + 1, last_q = q_ptr + 1;
      }
      t_ptr = end;
      len -= step;
      if (found) {
        if (pairs < max_pairs) {
          out[4 * pairs + 0] = first_t, out[4 * pairs + 1] = first_q;
          out[4 * pairs + 2] = last_t, out[4 * pairs + 3] = last_q;
        }
        ++pairs;
      }
      found = false;
      if (end + 1 >= c.t_end) {
        ends_left = false;  // that was t_end - 1
      } else {
        next_multiple += c.window_length;
        end = next_multiple < c.t_end ? next_multiple - 1 : c.t_end - 1;
      }
    }
    if (len > 0) {
      if (op == kOpM) {
        if (!found) found = true, first_t = t_ptr + 1, first_q = q_ptr + 1;
        q_ptr += len;
        last_t = t_ptr + len + 1, last_q = q_ptr + 1;
      }
      t_ptr += len;
    }
  }
  return pairs;
}

// Cells of arena an alignment with edit distance D occupies.
OVL_HD uint64_t wf_cells(uint64_t D) { return (D + 1) * (D + 1); }

}  // namespace ovl
#endif

// Synthetic long-read simulator + windowizer (host only, no device code).
//
// Produces the window batches of BASELINE.json's synthetic configs directly in the
// vgc_batch layout (include/vgc.h): a random genome (optionally two haplotypes differing by SNPs),
// reads with i.i.d. ins/del/sub errors and N(mu,sd) qualities, ground-truth overlaps, and — in place of
// the reference's edlib alignment + Overlap::find_breaking_points_from_cigar (src/overlap.cpp:226-292)
// — window breakpoints derived from the true read-to-read alignment with the SAME rule: per window the
// layer runs from the first to the last aligned (M) pair inside the window, begin/end are the target
// offsets of those pairs (end inclusive), exactly what Polisher::initialize passes to
// Window::add_layer (src/polisher.cpp:436-458).  Layer filters as src/polisher.cpp:416-434
// (piece < 0.02*w dropped; mean quality < threshold dropped).
//
// Both the reference CPU path and the GPU engine consume the identical batches, which is what the
// parity contract holds fixed ("identical overlaps and window tilings").

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vgc.h"

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed ? seed : 0x9E3779B97F4A7C15ull) {}
  inline uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  inline double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  inline uint32_t below(uint32_t n) { return static_cast<uint32_t>(uni() * n); }
  inline double normal() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

const char kBases[4] = {'A', 'C', 'G', 'T'};
inline char comp(char c) {
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return c;
  }
}

struct Read {
  std::string seq;            // forward-genome orientation
  std::string qual;           // empty in FASTA mode
  std::vector<int32_t> gpos;  // genome coordinate per base, -1 for inserted bases
  int32_t g_begin, g_end;     // genome interval [g_begin, g_end)
  uint8_t strand;             // 1: sequenced as reverse complement
  uint8_t hap;
};

}  // namespace

extern "C" {

typedef struct {
  uint64_t genome_len;
  uint32_t n_reads;
  uint32_t read_len;
  double p_ins, p_del, p_sub;
  double q_mean, q_sd;
  int32_t q_lo, q_hi;
  uint32_t window_len;
  uint32_t min_overlap;
  uint32_t n_haplotypes;   // 1 or 2
  double snp_rate;         // per-base SNP probability between haplotypes
  uint32_t fasta;          // 1: no qualities (FASTA reads)
  uint32_t random_strand;  // 1: reads are sequenced from a random strand
  uint64_t seed;
} sim_config;

struct sim_state {
  sim_config cfg;
  std::vector<std::string> hap;  // haplotype genomes
  std::vector<Read> reads;
  std::vector<uint32_t> by_start;  // read ids sorted by g_begin
};

sim_state* sim_create(const sim_config* cfg) {
  auto* st = new sim_state();
  st->cfg = *cfg;
  Rng rng(cfg->seed);
  std::string g(cfg->genome_len, 'A');
  for (auto& c : g) c = kBases[rng.next() & 3];
  st->hap.push_back(g);
  if (cfg->n_haplotypes > 1) {
    std::string h = g;
    for (auto& c : h) {
      if (rng.uni() < cfg->snp_rate) {
        char o = c;
        while (o == c) o = kBases[rng.next() & 3];
        c = o;
      }
    }
    st->hap.push_back(h);
  }
  st->reads.resize(cfg->n_reads);
  const uint64_t span = static_cast<uint64_t>(cfg->read_len * 1.25) + 64;
  for (uint32_t r = 0; r < cfg->n_reads; ++r) {
    Read& rd = st->reads[r];
    rd.hap = cfg->n_haplotypes > 1 ? static_cast<uint8_t>(rng.next() & 1) : 0;
    rd.strand = cfg->random_strand ? static_cast<uint8_t>(rng.next() & 1) : 0;
    const std::string& G = st->hap[rd.hap];
    uint64_t start = cfg->genome_len > span ? static_cast<uint64_t>(rng.uni() * (cfg->genome_len - span)) : 0;
    rd.seq.reserve(cfg->read_len);
    rd.gpos.reserve(cfg->read_len);
    uint64_t gp = start;
    while (rd.seq.size() < cfg->read_len && gp < cfg->genome_len) {
      double u = rng.uni();
      if (u < cfg->p_del) {
        ++gp;  // deleted genome base
      } else {
        char c = G[gp];
        if (u < cfg->p_del + cfg->p_sub) {
          char o = c;
          while (o == c) o = kBases[rng.next() & 3];
          c = o;
        }
        rd.seq.push_back(c);
        rd.gpos.push_back(static_cast<int32_t>(gp));
        ++gp;
      }
      while (rd.seq.size() < cfg->read_len && rng.uni() < cfg->p_ins) {
        rd.seq.push_back(kBases[rng.next() & 3]);
        rd.gpos.push_back(-1);
      }
    }
    rd.g_begin = static_cast<int32_t>(start);
    rd.g_end = static_cast<int32_t>(gp);
    if (!cfg->fasta) {
      rd.qual.resize(rd.seq.size());
      for (auto& q : rd.qual) {
        int v = static_cast<int>(std::lround(cfg->q_mean + cfg->q_sd * rng.normal()));
        v = std::max(cfg->q_lo, std::min(cfg->q_hi, v));
        q = static_cast<char>(33 + v);
      }
    }
  }
  st->by_start.resize(cfg->n_reads);
  for (uint32_t i = 0; i < cfg->n_reads; ++i) st->by_start[i] = i;
  std::sort(st->by_start.begin(), st->by_start.end(), [&](uint32_t a, uint32_t b) {
    return st->reads[a].g_begin < st->reads[b].g_begin || (st->reads[a].g_begin == st->reads[b].g_begin && a < b);
  });
  return st;
}

void sim_destroy(sim_state* st) { delete st; }

uint32_t sim_num_reads(const sim_state* st) { return st->cfg.n_reads; }

// Read r as sequenced (reverse-complemented when its strand bit is set).  Returns the length.
uint32_t sim_get_read(const sim_state* st, uint32_t r, char* seq, char* qual, uint32_t cap) {
  const Read& rd = st->reads[r];
  uint32_t n = static_cast<uint32_t>(rd.seq.size());
  if (n > cap) return n;
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t s = rd.strand ? n - 1 - i : i;
    seq[i] = rd.strand ? comp(rd.seq[s]) : rd.seq[s];
    if (qual) qual[i] = rd.qual.empty() ? '!' : rd.qual[s];
  }
  return n;
}

// Heap-allocated batch (free with sim_free_batch).  The const pointers of vgc_batch alias these.
struct sim_batch {
  vgc_batch b;
  std::vector<uint8_t> bases, quals, has_qual, win_flags;
  std::vector<uint64_t> seq_off;
  std::vector<uint32_t> begin, end, win_first;
  std::vector<uint32_t> win_target, win_rank;  // which read / which window of it
  std::vector<uint32_t> tgt_cov;               // overlaps per target t0..t1-1 (targets_coverages_, src/polisher.cpp:411)
  uint64_t n_overlaps;
};

namespace {

// View of a read oriented like the target: `flip` reverse-complements it and negates coordinates so
// that aligned coordinates stay ascending.
struct Oriented {
  std::string seq, qual;
  std::vector<int64_t> g;  // INT64_MIN for inserted bases
};
const int64_t kIns = INT64_MIN;

void orient(const Read& rd, bool flip, Oriented* o) {
  size_t n = rd.seq.size();
  o->seq.resize(n);
  o->qual.resize(rd.qual.size());
  o->g.resize(n);
  for (size_t i = 0; i < n; ++i) {
    size_t s = flip ? n - 1 - i : i;
    o->seq[i] = flip ? comp(rd.seq[s]) : rd.seq[s];
    if (!rd.qual.empty()) o->qual[i] = rd.qual[s];
    o->g[i] = rd.gpos[s] < 0 ? kIns : (flip ? -static_cast<int64_t>(rd.gpos[s]) : rd.gpos[s]);
  }
}

}  // namespace

// Windows of targets [t0, t1) with every overlapping read attached as layers (overlaps in the order
// of increasing query id, as an all-vs-all PAF sorted by query would be consumed).
sim_batch* sim_windows(const sim_state* st, uint32_t t0, uint32_t t1, double quality_threshold) {
  const sim_config& cfg = st->cfg;
  auto* sb = new sim_batch();
  sb->n_overlaps = 0;
  const uint32_t w = cfg.window_len;
  sb->seq_off.push_back(0);
  sb->win_first.push_back(0);
  Oriented T, Q;
  struct Layer {
    uint32_t q_begin, q_end, t_first, t_last;
    uint32_t query;
  };
  for (uint32_t t = t0; t < t1 && t < cfg.n_reads; ++t) {
    const Read& tr = st->reads[t];
    const bool flip = tr.strand != 0;
    orient(tr, flip, &T);
    const uint32_t tn = static_cast<uint32_t>(T.seq.size());
    const uint32_t n_win = (tn + w - 1) / w;
    std::vector<std::vector<Layer>> layers(n_win);
    std::vector<Oriented> qcache;
    std::vector<uint32_t> qids;
    // candidate queries: genome intervals intersecting by >= min_overlap, same haplotype genome coords
    for (uint32_t q = 0; q < cfg.n_reads; ++q) {
      if (q == t) continue;
      const Read& qr = st->reads[q];
      int32_t lo = std::max(qr.g_begin, tr.g_begin), hi = std::min(qr.g_end, tr.g_end);
      if (hi - lo < static_cast<int32_t>(cfg.min_overlap)) continue;
      qids.push_back(q);
    }
    qcache.resize(qids.size());
    sb->tgt_cov.push_back(static_cast<uint32_t>(qids.size()));
    for (size_t qi = 0; qi < qids.size(); ++qi) {
      orient(st->reads[qids[qi]], flip, &qcache[qi]);
      const Oriented& QQ = qcache[qi];
      ++sb->n_overlaps;
      // two-pointer merge over ascending aligned coordinates
      size_t i = 0, j = 0;
      const size_t qn = QQ.seq.size();
      uint32_t cur_w = UINT32_MAX;
      bool found = false;
      Layer cur{};
      auto flush = [&]() {
        if (found) layers[cur_w].push_back(cur);
        found = false;
      };
      while (i < tn && j < qn) {
        if (T.g[i] == kIns) { ++i; continue; }
        if (QQ.g[j] == kIns) { ++j; continue; }
        if (T.g[i] < QQ.g[j]) { ++i; continue; }
        if (T.g[i] > QQ.g[j]) { ++j; continue; }
        uint32_t wi = static_cast<uint32_t>(i) / w;
        if (wi != cur_w) {
          flush();
          cur_w = wi;
        }
        if (!found) {
          found = true;
          cur.q_begin = static_cast<uint32_t>(j);
          cur.t_first = static_cast<uint32_t>(i);
          cur.query = static_cast<uint32_t>(qi);
        }
        cur.q_end = static_cast<uint32_t>(j) + 1;
        cur.t_last = static_cast<uint32_t>(i);
        ++i;
        ++j;
      }
      flush();
    }
    // emit windows
    for (uint32_t k = 0; k < n_win; ++k) {
      const uint32_t ws = k * w, we = std::min(tn, ws + w);
      sb->win_target.push_back(t);
      sb->win_rank.push_back(k);
      uint8_t flags = VGC_WIN_TGS;
      if (T.qual.empty() && (we - ws) == w) flags |= VGC_WIN_DUMMY_QUAL;  // src/window.cpp:223 with dummy_quality_
      sb->win_flags.push_back(flags);
      // backbone
      sb->bases.insert(sb->bases.end(), T.seq.begin() + ws, T.seq.begin() + we);
      if (!T.qual.empty()) {
        sb->quals.insert(sb->quals.end(), T.qual.begin() + ws, T.qual.begin() + we);
      } else {
        sb->quals.insert(sb->quals.end(), we - ws, '!');
      }
      sb->seq_off.push_back(sb->bases.size());
      sb->has_qual.push_back(1);  // the backbone always carries a quality pointer (real or dummy)
      sb->begin.push_back(0);
      sb->end.push_back(0);
      for (const Layer& L : layers[k]) {
        const Oriented& QQ = qcache[L.query];
        const uint32_t len = L.q_end - L.q_begin;
        if (len < 0.02 * w) continue;  // src/polisher.cpp:416
        if (!QQ.qual.empty()) {        // src/polisher.cpp:420-434
          double avg = 0;
          for (uint32_t x = L.q_begin; x < L.q_end; ++x) avg += static_cast<uint32_t>(QQ.qual[x]) - 33;
          avg /= len;
          if (avg < quality_threshold) continue;
        }
        const uint32_t b = L.t_first - ws, e = L.t_last - ws;
        if (b == e) continue;  // Window::add_layer ignores begin == end (src/window.cpp:51)
        sb->bases.insert(sb->bases.end(), QQ.seq.begin() + L.q_begin, QQ.seq.begin() + L.q_end);
        if (!QQ.qual.empty()) {
          sb->quals.insert(sb->quals.end(), QQ.qual.begin() + L.q_begin, QQ.qual.begin() + L.q_end);
          sb->has_qual.push_back(1);
        } else {
          sb->quals.insert(sb->quals.end(), len, '!');
          sb->has_qual.push_back(0);
        }
        sb->seq_off.push_back(sb->bases.size());
        sb->begin.push_back(b);
        sb->end.push_back(e);
      }
      sb->win_first.push_back(static_cast<uint32_t>(sb->begin.size()));
    }
  }
  sb->b.n_windows = static_cast<uint32_t>(sb->win_flags.size());
  sb->b.n_layers = static_cast<uint32_t>(sb->begin.size());
  sb->b.bases = sb->bases.data();
  sb->b.quals = sb->quals.data();
  sb->b.seq_off = sb->seq_off.data();
  sb->b.has_qual = sb->has_qual.data();
  sb->b.begin = sb->begin.data();
  sb->b.end = sb->end.data();
  sb->b.win_first = sb->win_first.data();
  sb->b.win_flags = sb->win_flags.data();
  return sb;
}

const vgc_batch* sim_batch_view(const sim_batch* sb) { return &sb->b; }
const uint32_t* sim_batch_targets(const sim_batch* sb) { return sb->win_target.data(); }
const uint32_t* sim_batch_ranks(const sim_batch* sb) { return sb->win_rank.data(); }
const uint32_t* sim_batch_target_coverages(const sim_batch* sb) { return sb->tgt_cov.data(); }
uint32_t sim_batch_num_targets(const sim_batch* sb) { return static_cast<uint32_t>(sb->tgt_cov.size()); }
uint64_t sim_batch_overlaps(const sim_batch* sb) { return sb->n_overlaps; }
void sim_free_batch(sim_batch* sb) { delete sb; }

// Reads [0, n_reads) as a FASTQ (or FASTA) file and their ground-truth all-vs-all overlaps as PAF, the inputs of the
// whole program (`vechat_racon <reads> <overlaps> <targets>`, src/main.cpp): one PAF line per ordered pair (query,
// target) whose genome intervals share >= min_overlap bases, grouped by query, coordinates on the strands as
// sequenced.  Only where the reads lie is taken from the truth — the alignment itself is left to the program
// (Overlap::find_breaking_points).  Returns the number of overlaps written, or -1.
long long sim_export(const sim_state* st, const char* reads_path, const char* paf_path) {
  const sim_config& cfg = st->cfg;
  FILE* fr = std::fopen(reads_path, "wb");
  FILE* fp = std::fopen(paf_path, "wb");
  if (!fr || !fp) {
    if (fr) std::fclose(fr);
    if (fp) std::fclose(fp);
    return -1;
  }
  std::vector<char> seq, qual;
  for (uint32_t r = 0; r < cfg.n_reads; ++r) {
    const uint32_t n = static_cast<uint32_t>(st->reads[r].seq.size());
    seq.resize(n + 1);
    qual.resize(n + 1);
    sim_get_read(st, r, seq.data(), qual.data(), n + 1);
    if (cfg.fasta) {
      std::fprintf(fr, ">read%u\n%.*s\n", r, static_cast<int>(n), seq.data());
    } else {
      std::fprintf(fr, "@read%u\n%.*s\n+\n%.*s\n", r, static_cast<int>(n), seq.data(), static_cast<int>(n), qual.data());
    }
  }
  std::fclose(fr);
  // [begin, end) of read `rd` (as sequenced) covering genome interval [lo, hi)
  auto span = [](const Read& rd, int32_t lo, int32_t hi, uint32_t* b, uint32_t* e) {
    const uint32_t n = static_cast<uint32_t>(rd.seq.size());
    uint32_t fb = 0, fe = n;
    while (fb < n && (rd.gpos[fb] < 0 || rd.gpos[fb] < lo)) ++fb;
    while (fe > fb && (rd.gpos[fe - 1] < 0 || rd.gpos[fe - 1] >= hi)) --fe;
    if (rd.strand) {
      *b = n - fe;
      *e = n - fb;
    } else {
      *b = fb;
      *e = fe;
    }
  };
  long long written = 0;
  const uint32_t nr = cfg.n_reads;
  int32_t max_span = 0;
  for (const Read& rd : st->reads) max_span = std::max(max_span, rd.g_end - rd.g_begin);
  for (uint32_t q = 0; q < nr; ++q) {
    const Read& qr = st->reads[q];
    // candidates start before my end and end after my begin: walk by_start around my interval
    auto it = std::lower_bound(st->by_start.begin(), st->by_start.end(), qr.g_begin - max_span,
                               [&](uint32_t a, int32_t v) { return st->reads[a].g_begin < v; });
    std::vector<uint32_t> ts;
    for (; it != st->by_start.end() && st->reads[*it].g_begin < qr.g_end; ++it) {
      const uint32_t t = *it;
      if (t == q) continue;
      const Read& tr = st->reads[t];
      const int32_t lo = std::max(qr.g_begin, tr.g_begin), hi = std::min(qr.g_end, tr.g_end);
      if (hi - lo < static_cast<int32_t>(cfg.min_overlap)) continue;
      ts.push_back(t);
    }
    std::sort(ts.begin(), ts.end());
    for (uint32_t t : ts) {
      const Read& tr = st->reads[t];
      const int32_t lo = std::max(qr.g_begin, tr.g_begin), hi = std::min(qr.g_end, tr.g_end);
      uint32_t qb, qe, tb, te;
      span(qr, lo, hi, &qb, &qe);
      span(tr, lo, hi, &tb, &te);
      if (qe <= qb || te <= tb) continue;
      const uint32_t alen = std::max(qe - qb, te - tb);
      std::fprintf(fp, "read%u\t%zu\t%u\t%u\t%c\tread%u\t%zu\t%u\t%u\t%u\t%u\t255\n", q, qr.seq.size(), qb, qe,
                   qr.strand == tr.strand ? '+' : '-', t, tr.seq.size(), tb, te, alen * 7 / 10, alen);
      ++written;
    }
  }
  std::fclose(fp);
  return written;
}

}  // extern "C"

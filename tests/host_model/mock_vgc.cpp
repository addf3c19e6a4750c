// TEST-ONLY.  LD_PRELOAD shim that lets the REAL vechat_racon_b200 binary run on a box without a GPU so that the
// host side of the binding (csrc/racon_binding/b200polisher.cpp: device ranges, batching, packing, store, stitch) is
// checked in the CPU suite.  It interposes vgc_create / vgc_polish / vgc_destroy / vgc_last_error and answers
// vgc_polish with the CHECKER (oracle/_ref/libvechat_ref.so, the unmodified reference window code, path in
// MOCK_VGC_REF_SO).  Everything else (vgc_result_bound ...) still resolves to the real libvgc.so.  It never ships:
// the product has no CPU path, and tests/test_example_binary.py asserts the un-preloaded binary fails without a GPU.
#include <dlfcn.h>

#include <atomic>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vgc.h"

namespace {
thread_local std::string g_err;
typedef int (*ref_polish_fn)(const vgc_batch*, const vgc_params*, vgc_result*, int);
ref_polish_fn ref_polish() {
  static ref_polish_fn fn = [] {
    const char* path = std::getenv("MOCK_VGC_REF_SO");
    void* so = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
    return so ? reinterpret_cast<ref_polish_fn>(dlsym(so, "ref_polish")) : nullptr;
  }();
  return fn;
}
std::atomic<int> g_handles{0};
}  // namespace

struct vgc_engine {
  vgc_params params;
  int device;
  int id;
  std::mutex mu;
  std::deque<vgc_batch> queue;  // vgc_submit'ed, not yet collected
};

extern "C" {

const char* vgc_last_error(void) { return g_err.c_str(); }

int vgc_create(vgc_handle* out, int device, const vgc_params* params) {
  if (!ref_polish()) {
    g_err = "mock: MOCK_VGC_REF_SO does not name a loadable libvechat_ref.so";
    return VGC_ERR_NO_DEVICE;
  }
  vgc_engine* h = new vgc_engine();
  h->params = *params;
  h->device = device;
  h->id = g_handles++;
  *out = h;
  return VGC_OK;
}

int vgc_destroy(vgc_handle h) {
  delete h;
  return VGC_OK;
}

int vgc_polish(vgc_handle h, const vgc_batch* batch, vgc_result* result, vgc_stats* stats) {
  std::fprintf(stderr, "[mock_vgc] handle %d device %d: %u windows, %u layers, %llu bytes\n", h->id, h->device,
               batch->n_windows, batch->n_layers, static_cast<unsigned long long>(batch->seq_off[batch->n_layers]));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  const int rc = ref_polish()(batch, &h->params, result, 4);
  if (rc != 0) g_err = "mock: ref_polish failed";
  return rc;
}

// the two-halves form the binding uses: the mock "stages" by remembering the caller's arrays
int vgc_submit(vgc_handle h, const vgc_batch* batch) {
  std::lock_guard<std::mutex> lock(h->mu);
  if (h->queue.size() >= 2) {
    g_err = "mock: two batches are already staged";
    return VGC_ERR_INVALID;
  }
  h->queue.push_back(*batch);
  return VGC_OK;
}

int vgc_collect(vgc_handle h, vgc_result* result, vgc_stats* stats) {
  vgc_batch b;
  {
    std::lock_guard<std::mutex> lock(h->mu);
    if (h->queue.empty()) {
      g_err = "mock: nothing submitted";
      return VGC_ERR_INVALID;
    }
    b = h->queue.front();
    h->queue.pop_front();
  }
  return vgc_polish(h, &b, result, stats);
}

}  // extern "C"

// ---- the overlap aligner (include/vga.h) behind the same shim: ovl_core.h driven serially on the host, so the
// binding's CUDABatchAligner glue (substring offsets, strand handling, breaking_points_ / cigar_ hand-over) is checked
// on the CPU box as well.
#include <vector>

#include "ovl_core.h"
#include "vga.h"

struct vga_aligner {
  std::vector<char> cigar;
  std::vector<uint64_t> cigar_off, points_off;
  std::vector<int32_t> edit;
  std::vector<uint32_t> points;
};

namespace {
uint32_t model_runs(const uint8_t* q, int32_t m, const uint8_t* t, int32_t n, std::vector<uint32_t>* runs, int32_t* edit) {
  std::vector<int32_t> arena;
  int32_t D = -1;
  for (int32_t d = 0; D < 0; ++d) {
    arena.resize(ovl::wf_cells(d));
    const ovl::Front prev = ovl::wf_front(arena.data(), d > 0 ? d - 1 : 0, m, n);
    const ovl::Front cur = ovl::wf_front(arena.data(), d, m, n);
    for (int32_t k = cur.lo; k <= cur.hi; ++k) {
      const int32_t i = ovl::wf_cell(prev, q, t, m, n, d, k);
      arena[ovl::wf_index(d, k)] = i;
      if (i == m && k == n - m) D = d;
    }
  }
  runs->resize(static_cast<size_t>(m) + n + 2);
  *edit = D;
  return ovl::wf_traceback(arena.data(), m, n, D, runs->data());
}
}  // namespace

extern "C" {

const char* vga_last_error(void) { return g_err.c_str(); }
int vga_create(vga_handle* out, int) {
  *out = new vga_aligner();
  return VGA_OK;
}
int vga_destroy(vga_handle h) {
  delete h;
  return VGA_OK;
}

int vga_align(vga_handle h, const vga_batch* b, vga_result* r, vga_stats* st) {
  std::fprintf(stderr, "[mock_vga] align: %u overlaps\n", b->n);
  h->cigar.clear();
  h->cigar_off.assign(1, 0);
  h->edit.clear();
  std::vector<uint32_t> runs;
  for (uint32_t i = 0; i < b->n; ++i) {
    int32_t edit;
    const uint32_t nr = model_runs(b->seqs + b->q_off[i], b->q_len[i], b->seqs + b->t_off[i], b->t_len[i], &runs, &edit);
    std::string s;
    for (uint32_t x = nr; x-- > 0;) s += std::to_string(runs[x] >> 2) + "MID?"[runs[x] & 3];
    h->cigar.insert(h->cigar.end(), s.c_str(), s.c_str() + s.size() + 1);
    h->cigar_off.push_back(h->cigar.size());
    h->edit.push_back(edit);
  }
  r->cigar = h->cigar.data();
  r->cigar_off = h->cigar_off.data();
  r->edit_distance = h->edit.data();
  if (st) std::memset(st, 0, sizeof(*st));
  return VGA_OK;
}

int vga_break(vga_handle h, const vga_batch* b, const vga_cut* c, vga_breaks* r, vga_stats* st) {
  std::fprintf(stderr, "[mock_vga] break: %u overlaps\n", b->n);
  h->points.clear();
  h->points_off.assign(1, 0);
  h->edit.clear();
  std::vector<uint32_t> runs;
  for (uint32_t i = 0; i < b->n; ++i) {
    int32_t edit;
    const uint32_t nr = model_runs(b->seqs + b->q_off[i], b->q_len[i], b->seqs + b->t_off[i], b->t_len[i], &runs, &edit);
    const uint32_t cap = b->t_len[i] / c->window_length + 3;
    std::vector<uint32_t> out(4 * cap);
    const ovl::CutParams cp{c->t_begin[i], c->t_begin[i] + b->t_len[i], c->q_start[i], c->window_length};
    const uint32_t pairs = ovl::wf_cut(runs.data(), nr, cp, out.data(), cap);
    h->points.insert(h->points.end(), out.begin(), out.begin() + 4 * pairs);
    h->points_off.push_back(h->points.size() / 4);
    h->edit.push_back(edit);
  }
  r->points = h->points.data();
  r->points_off = h->points_off.data();
  r->edit_distance = h->edit.data();
  if (st) std::memset(st, 0, sizeof(*st));
  return VGA_OK;
}

}  // extern "C"

// ovl_align.cu — batched overlap alignment on sm_100a behind include/vga.h.
//
// Reference work replaced: one edlib NW alignment with path per overlap (src/overlap.cpp:205-224), fanned out over a
// thread pool by Polisher::find_overlap_breaking_points (src/polisher.cpp:464-489).
//
// Mapping: one 256-thread CTA per overlap, CTAs pull overlaps (longest first) from an atomic queue.  The CTA sweeps
// the furthest-reaching wavefronts d = 0, 1, ... (ovl_core.h): the 2d+1 diagonals of a wavefront are independent, one
// per thread, each reading its three neighbours of wavefront d-1 and sliding along the matching run of the two
// sequences; one barrier (with an OR-reduction of "reached the corner") per wavefront.  Wavefronts are written once to
// the CTA's private arena in HBM — 4 B per cell, (D+1)^2 cells, the algorithmic bytes of the kernel — and read back
// by the traceback, which one thread walks (three independent loads per edit) while merging equal ops into CIGAR
// runs; the CTA then reserves space in the batch's output buffer with one atomic and copies the runs out in forward
// order.  Integer/byte work, HBM/L2 bound: no tensor cores.  An overlap whose wavefronts outgrow the arena, or whose
// runs do not fit the output buffer any more, is reported back and re-run by the host loop with more room.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ovl_core.h"
#include "vga.h"

namespace {

constexpr int kThreads = 256;      // CTA size of a full round: 8 CTAs per SM
constexpr int kCtasPerSm = 8;
constexpr int kThreadsFew = 1024;  // CTA size when a round has few overlaps (retries, small batches): the critical
                                   // path is one alignment, and a wavefront is as wide as its edit distance

struct OvlTask {
  uint64_t q_off, t_off;
  uint32_t m, n;
};
enum : uint32_t { kPending = 0, kDone = 1, kArenaFull = 2, kOutFull = 3 };
struct OvlMeta {
  int32_t edit;
  uint32_t status;
  uint32_t n_runs;
  uint32_t pad;
  uint64_t run_off;
};

__global__ void __launch_bounds__(kThreadsFew)
ovl_kernel(const uint8_t* __restrict__ seqs, const OvlTask* __restrict__ tasks, const uint32_t* __restrict__ work,
           uint32_t n_work, int32_t* arenas, uint64_t arena_cells, uint32_t* scratch, uint64_t scratch_stride,
           uint32_t* out_runs, uint64_t out_cap, unsigned long long* cursors, OvlMeta* meta,
           const ovl::CutParams* __restrict__ cuts, uint32_t cut_base, uint32_t max_pairs) {
  __shared__ uint32_t s_item, s_status, s_nruns;
  __shared__ unsigned long long s_off;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  int32_t* arena = arenas + static_cast<size_t>(blockIdx.x) * arena_cells;
  uint32_t* runs = scratch + static_cast<size_t>(blockIdx.x) * scratch_stride;
  for (;;) {
    if (tid == 0) s_item = static_cast<uint32_t>(atomicAdd(&cursors[0], 1ull));
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_work) break;
    const uint32_t ov = work[item];
    const OvlTask task = tasks[ov];
    const uint8_t* q = seqs + task.q_off;
    const uint8_t* t = seqs + task.t_off;
    const int32_t m = static_cast<int32_t>(task.m), n = static_cast<int32_t>(task.n), kf = n - m;

    // output buffer already (nearly) full: do not spend the work, the host re-runs this overlap in the next round
    uint32_t status = kDone;
    if (tid == 0) {
      const unsigned long long used = *reinterpret_cast<volatile unsigned long long*>(&cursors[1]);
      const unsigned long long worst = cuts ? 4ull * max_pairs : task.m + task.n + 1ull;
      s_status = (used + worst > out_cap) ? kOutFull : kDone;
    }
    __syncthreads();
    status = s_status;

    int32_t D = -1;
    if (status == kDone) {
      for (int32_t d = 0;; ++d) {
        if (ovl::wf_cells(static_cast<uint64_t>(d)) > arena_cells) {
          status = kArenaFull;
          break;
        }
        const ovl::Front prev = ovl::wf_front(arena, d > 0 ? d - 1 : 0, m, n);
        const ovl::Front cur = ovl::wf_front(arena, d, m, n);
        int32_t* out = arena + ovl::wf_index(d, 0);
        int reached = 0;
        for (int32_t k = cur.lo + tid; k <= cur.hi; k += nthreads) {
          const int32_t i = ovl::wf_cell(prev, q, t, m, n, d, k);
          out[k] = i;
          reached |= (k == kf && i == m);
        }
        if (__syncthreads_or(reached)) {  // also orders wavefront d's stores before wavefront d+1's loads
          D = d;
          break;
        }
      }
    }

    if (status == kDone) {
      if (tid == 0) {
        uint32_t nr = ovl::wf_traceback(arena, m, n, D, runs);
        if (cuts) {  // breaking points instead of the CIGAR runs: 4 words per pair, behind the runs in the scratch
          uint32_t pairs = ovl::wf_cut(runs, nr, cuts[ov], runs + cut_base, max_pairs);
          nr = 4 * (pairs < max_pairs ? pairs : max_pairs);
        }
        const unsigned long long off = atomicAdd(&cursors[1], static_cast<unsigned long long>(nr));
        s_nruns = nr;
        s_off = off;
        s_status = (off + nr > out_cap) ? kOutFull : kDone;
      }
      __syncthreads();
      status = s_status;
      const uint32_t nr = s_nruns;
      if (status == kDone) {
        if (cuts) {
          for (uint32_t x = tid; x < nr; x += nthreads) out_runs[s_off + x] = runs[cut_base + x];
        } else {  // the traceback wrote the runs back to front
          for (uint32_t x = tid; x < nr; x += nthreads) out_runs[s_off + x] = runs[nr - 1 - x];
        }
      }
    }
    if (tid == 0) {
      OvlMeta mt;
      mt.edit = D;
      mt.status = status;
      mt.n_runs = status == kDone ? s_nruns : 0;
      mt.pad = 0;
      mt.run_off = status == kDone ? s_off : 0;
      meta[ov] = mt;
      if (status == kDone) atomicAdd(&cursors[2], static_cast<unsigned long long>(ovl::wf_cells(D)));
    }
    __syncthreads();  // s_* are rewritten by the next item
  }
}

thread_local std::string g_err;
void set_err(const std::string& s) { g_err = s; }

#define VGA_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      set_err(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call);     \
      return VGA_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

struct DevBuf {  // frees on scope exit
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

size_t run_text_len(const uint32_t* r, uint32_t n) {
  size_t len = 0;
  for (uint32_t x = 0; x < n; ++x) {
    uint32_t v = r[x] >> 2;
    do ++len, v /= 10; while (v);
    ++len;
  }
  return len;
}
void run_text(const uint32_t* r, uint32_t n, std::string* s) {
  static const char ops[4] = {'M', 'I', 'D', '?'};
  s->clear();
  s->reserve(run_text_len(r, n));
  char buf[12];
  for (uint32_t x = 0; x < n; ++x) {
    uint32_t v = r[x] >> 2;
    int k = 0;
    do buf[k++] = static_cast<char>('0' + v % 10), v /= 10; while (v);
    while (k) s->push_back(buf[--k]);
    s->push_back(ops[r[x] & 3]);
  }
}

template <typename F>
void parallel_for(size_t n, F f) {
  unsigned nt = std::min<size_t>(std::max(1u, std::min(32u, std::thread::hardware_concurrency())), (n + 63) / 64);
  if (nt <= 1) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back(f, n * t / nt, n * (t + 1) / nt);
  for (auto& x : th) x.join();
}

}  // namespace

struct vga_aligner {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  // results of the last vga_align (owned here, see vga.h)
  std::vector<char> cigar;
  std::vector<uint64_t> cigar_off;
  std::vector<int32_t> edit;
  std::vector<uint32_t> points;
  std::vector<uint64_t> points_off;
};

extern "C" {

const char* vga_last_error(void) { return g_err.c_str(); }

int vga_create(vga_handle* out, int device) {
  if (!out) {
    set_err("null argument");
    return VGA_ERR_INVALID;
  }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
    set_err(std::string("no usable CUDA device (the aligner has no CPU fallback): ") +
            (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range"));
    cudaGetLastError();
    return VGA_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  VGA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_err("device is not sm_100-class (Blackwell); this library carries sm_100a code only");
    return VGA_ERR_NO_DEVICE;
  }
  VGA_CUDA(cudaSetDevice(device));
  vga_aligner* h = new vga_aligner();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  VGA_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto& ev : h->ev) VGA_CUDA(cudaEventCreate(&ev));
  *out = h;
  return VGA_OK;
}

int vga_destroy(vga_handle h) {
  if (!h) return VGA_OK;
  cudaSetDevice(h->device);
  for (auto& ev : h->ev)
    if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return VGA_OK;
}

}  // extern "C"

// The round loop shared by vga_align (cut == nullptr: CIGAR runs) and vga_break (breaking points).  consume(ov,
// words, n, edit) is called once per overlap, from several host threads, with that overlap's output words.
template <typename Consume>
static int run_rounds(vga_handle h, const vga_batch* b, const vga_cut* cut, vga_stats* st, Consume consume) {
  const uint32_t n = b->n;
  if (n && (!b->seqs || !b->q_off || !b->q_len || !b->t_off || !b->t_len)) {
    set_err("null array in vga_batch");
    return VGA_ERR_INVALID;
  }
  if (cut && n && (!cut->t_begin || !cut->q_start || cut->window_length == 0)) {
    set_err("null array / zero window_length in vga_cut");
    return VGA_ERR_INVALID;
  }
  uint32_t longest = 0, max_pairs = 0;  // max over overlaps of m + n; of the windows a target substring spans
  for (uint32_t i = 0; i < n; ++i) {
    if (b->q_off[i] + b->q_len[i] > b->seqs_len || b->t_off[i] + b->t_len[i] > b->seqs_len ||
        b->q_len[i] > (1u << 28) || b->t_len[i] > (1u << 28)) {
      set_err("overlap " + std::to_string(i) + " points outside vga_batch.seqs");
      return VGA_ERR_INVALID;
    }
    longest = std::max(longest, b->q_len[i] + b->t_len[i]);
    if (cut) max_pairs = std::max(max_pairs, b->t_len[i] / cut->window_length + 3);
  }
  if (n == 0) return VGA_OK;

  VGA_CUDA(cudaSetDevice(h->device));
  std::vector<OvlTask> tasks(n);
  for (uint32_t i = 0; i < n; ++i) tasks[i] = OvlTask{b->q_off[i], b->t_off[i], b->q_len[i], b->t_len[i]};
  std::vector<uint32_t> pending(n);
  for (uint32_t i = 0; i < n; ++i) pending[i] = i;
  std::stable_sort(pending.begin(), pending.end(), [&](uint32_t x, uint32_t y) {
    return static_cast<uint64_t>(tasks[x].m) + tasks[x].n > static_cast<uint64_t>(tasks[y].m) + tasks[y].n;
  });

  DevBuf d_seqs, d_tasks, d_work, d_meta, d_cursors, d_scratch, d_cuts;
  VGA_CUDA(cudaMalloc(&d_seqs.p, b->seqs_len + 16));  // + padding for ovl::load4
  VGA_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(d_seqs.p) + b->seqs_len, 0, 16, h->stream));
  VGA_CUDA(cudaMalloc(&d_tasks.p, sizeof(OvlTask) * n));
  VGA_CUDA(cudaMalloc(&d_work.p, sizeof(uint32_t) * n));
  VGA_CUDA(cudaMalloc(&d_meta.p, sizeof(OvlMeta) * n));
  VGA_CUDA(cudaMalloc(&d_cursors.p, sizeof(unsigned long long) * 4));
  VGA_CUDA(cudaMemcpyAsync(d_seqs.p, b->seqs, b->seqs_len, cudaMemcpyHostToDevice, h->stream));
  VGA_CUDA(cudaMemcpyAsync(d_tasks.p, tasks.data(), sizeof(OvlTask) * n, cudaMemcpyHostToDevice, h->stream));
  VGA_CUDA(cudaMemsetAsync(d_meta.p, 0, sizeof(OvlMeta) * n, h->stream));
  std::vector<ovl::CutParams> cuts;
  if (cut) {
    cuts.resize(n);
    for (uint32_t i = 0; i < n; ++i)
      cuts[i] = ovl::CutParams{cut->t_begin[i], cut->t_begin[i] + b->t_len[i], cut->q_start[i], cut->window_length};
    VGA_CUDA(cudaMalloc(&d_cuts.p, sizeof(ovl::CutParams) * n));
    VGA_CUDA(cudaMemcpyAsync(d_cuts.p, cuts.data(), sizeof(ovl::CutParams) * n, cudaMemcpyHostToDevice, h->stream));
  }

  const uint32_t max_ctas = static_cast<uint32_t>(h->sm_count) * kCtasPerSm;
  const uint32_t cut_base = longest + 2;  // per-CTA scratch: the runs, then (cut mode) 4 words per pair
  const uint64_t scratch_stride = static_cast<uint64_t>(cut_base) + 4ull * max_pairs;
  const uint64_t worst_out = cut ? 4ull * max_pairs : cut_base;  // output words one overlap can need
  VGA_CUDA(cudaMalloc(&d_scratch.p, sizeof(uint32_t) * scratch_stride * std::min<uint64_t>(max_ctas, n)));

  size_t free_b = 0, total_b = 0;
  VGA_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t budget = static_cast<uint64_t>(free_b * 0.8);
  // output words of one round: a fifth of the budget, at most 4 GiB, at least one worst-case overlap
  uint64_t out_cap = std::min<uint64_t>(budget / 5, 4ull << 30) / sizeof(uint32_t);
  if (cut) out_cap = std::min<uint64_t>(out_cap, 4ull * max_pairs * n);
  uint64_t first_arena_cells = ~0ull;
  // test hooks: start with a small output buffer / small arenas so that the retry rounds run
  if (const char* e = std::getenv("VGA_OUT_CAP_RUNS")) out_cap = std::min<uint64_t>(out_cap, std::strtoull(e, nullptr, 10));
  if (const char* e = std::getenv("VGA_ARENA_CELLS")) first_arena_cells = std::strtoull(e, nullptr, 10);
  out_cap = std::max<uint64_t>(out_cap, worst_out);
  const uint64_t arena_budget = budget - std::min<uint64_t>(budget / 5, 4ull << 30);
  DevBuf d_out;
  VGA_CUDA(cudaMalloc(&d_out.p, out_cap * sizeof(uint32_t)));

  std::vector<OvlMeta> meta(n);
  std::vector<uint32_t> words_host;
  uint64_t arena_floor = 0;  // cells an arena must at least have (raised after an arena overflow)
  while (!pending.empty()) {
    uint32_t worst = 0;  // edit distance is at most max(m, n)
    for (uint32_t ov : pending) worst = std::max(worst, std::max(tasks[ov].m, tasks[ov].n));
    const uint64_t cells_worst = ovl::wf_cells(worst);
    const bool few = pending.size() * 4 <= max_ctas;
    const int threads = few ? kThreadsFew : kThreads;
    uint32_t ctas = static_cast<uint32_t>(std::min<uint64_t>(few ? max_ctas / 4 : max_ctas, pending.size()));
    uint64_t arena_cells = std::min<uint64_t>(cells_worst, arena_budget / sizeof(int32_t) / ctas);
    arena_cells = std::max<uint64_t>(std::min(arena_cells, first_arena_cells), 1);
    if (arena_cells < std::min(arena_floor, cells_worst)) {  // fewer, larger arenas
      arena_cells = std::min(arena_floor, cells_worst);
      ctas = static_cast<uint32_t>(std::min<uint64_t>(ctas, arena_budget / sizeof(int32_t) / arena_cells));
      if (ctas == 0) {
        set_err("an overlap needs more wavefront storage than the device memory budget");
        return VGA_ERR_CAPACITY;
      }
    }
    DevBuf d_arena;
    if (cudaMalloc(&d_arena.p, arena_cells * sizeof(int32_t) * ctas) != cudaSuccess) {
      cudaGetLastError();
      set_err("cudaMalloc of the wavefront arenas failed");
      return VGA_ERR_NOMEM;
    }
    VGA_CUDA(cudaMemcpyAsync(d_work.p, pending.data(), sizeof(uint32_t) * pending.size(), cudaMemcpyHostToDevice,
                             h->stream));
    VGA_CUDA(cudaMemsetAsync(d_cursors.p, 0, sizeof(unsigned long long) * 4, h->stream));
    VGA_CUDA(cudaEventRecord(h->ev[0], h->stream));
    ovl_kernel<<<ctas, threads, 0, h->stream>>>(
        d_seqs.as<uint8_t>(), d_tasks.as<OvlTask>(), d_work.as<uint32_t>(), static_cast<uint32_t>(pending.size()),
        d_arena.as<int32_t>(), arena_cells, d_scratch.as<uint32_t>(), scratch_stride, d_out.as<uint32_t>(), out_cap,
        d_cursors.as<unsigned long long>(), d_meta.as<OvlMeta>(), cut ? d_cuts.as<ovl::CutParams>() : nullptr, cut_base,
        max_pairs);
    VGA_CUDA(cudaGetLastError());
    VGA_CUDA(cudaEventRecord(h->ev[1], h->stream));
    unsigned long long cursors[4];
    VGA_CUDA(cudaMemcpyAsync(cursors, d_cursors.p, sizeof(cursors), cudaMemcpyDeviceToHost, h->stream));
    VGA_CUDA(cudaMemcpyAsync(meta.data(), d_meta.p, sizeof(OvlMeta) * n, cudaMemcpyDeviceToHost, h->stream));
    VGA_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0;
    VGA_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    st->kernel_ms += ms;
    st->kernel_launches += 1;
    st->cells += cursors[2];
    const uint64_t used = std::min<uint64_t>(cursors[1], out_cap);
    words_host.resize(used);
    if (used) VGA_CUDA(cudaMemcpy(words_host.data(), d_out.p, used * sizeof(uint32_t), cudaMemcpyDeviceToHost));

    std::vector<uint32_t> done, again;
    bool arena_full = false;
    for (uint32_t ov : pending) {
      if (meta[ov].status == kDone) {
        done.push_back(ov);
      } else {
        again.push_back(ov);
        arena_full |= meta[ov].status == kArenaFull;
      }
    }
    parallel_for(done.size(), [&](size_t lo, size_t hi) {
      for (size_t x = lo; x < hi; ++x) {
        const OvlMeta& mt = meta[done[x]];
        consume(done[x], words_host.data() + mt.run_off, mt.n_runs, mt.edit);
      }
    });
    if (done.empty() && (!arena_full || arena_cells >= cells_worst)) {
      set_err("an overlap's alignment does not fit the output buffer / wavefront storage of the device");
      return VGA_ERR_CAPACITY;
    }
    if (arena_full) arena_floor = std::max(arena_floor, arena_cells) * 4;
    st->retried += static_cast<uint32_t>(again.size());
    pending.swap(again);
  }
  return VGA_OK;
}

extern "C" {

int vga_align(vga_handle h, const vga_batch* b, vga_result* result, vga_stats* stats) {
  const auto t_begin = std::chrono::steady_clock::now();
  if (!h || !b || !result) {
    set_err("null argument");
    return VGA_ERR_INVALID;
  }
  const uint32_t n = b->n;
  h->cigar.clear();
  h->cigar_off.assign(static_cast<size_t>(n) + 1, 0);
  h->edit.assign(n, -1);
  vga_stats st;
  std::memset(&st, 0, sizeof(st));
  std::vector<std::string> text(n);
  const int rc = run_rounds(h, b, nullptr, &st, [&](uint32_t ov, const uint32_t* runs, uint32_t nr, int32_t edit) {
    run_text(runs, nr, &text[ov]);
    h->edit[ov] = edit;
  });
  if (rc != VGA_OK) return rc;
  // the n strings back to back, NUL-terminated
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    h->cigar_off[i] = total;
    total += text[i].size() + 1;
  }
  h->cigar_off[n] = total;
  h->cigar.resize(total);
  parallel_for(n, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) std::memcpy(h->cigar.data() + h->cigar_off[i], text[i].c_str(), text[i].size() + 1);
  });
  result->cigar = h->cigar.data();
  result->cigar_off = h->cigar_off.data();
  result->edit_distance = h->edit.data();
  st.wavefront_bytes = st.cells * sizeof(int32_t);
  st.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (stats) *stats = st;
  return VGA_OK;
}

int vga_break(vga_handle h, const vga_batch* b, const vga_cut* cut, vga_breaks* result, vga_stats* stats) {
  const auto t_begin = std::chrono::steady_clock::now();
  if (!h || !b || !cut || !result) {
    set_err("null argument");
    return VGA_ERR_INVALID;
  }
  const uint32_t n = b->n;
  h->points.clear();
  h->points_off.assign(static_cast<size_t>(n) + 1, 0);
  h->edit.assign(n, -1);
  vga_stats st;
  std::memset(&st, 0, sizeof(st));
  std::vector<std::vector<uint32_t>> per(n);
  const int rc = run_rounds(h, b, cut, &st, [&](uint32_t ov, const uint32_t* words, uint32_t nw, int32_t edit) {
    per[ov].assign(words, words + nw);
    h->edit[ov] = edit;
  });
  if (rc != VGA_OK) return rc;
  uint64_t pairs = 0;
  for (uint32_t i = 0; i < n; ++i) {
    h->points_off[i] = pairs;
    pairs += per[i].size() / 4;
  }
  h->points_off[n] = pairs;
  h->points.resize(pairs * 4);
  for (uint32_t i = 0; i < n; ++i)
    if (!per[i].empty()) std::memcpy(h->points.data() + 4 * h->points_off[i], per[i].data(), per[i].size() * sizeof(uint32_t));
  result->points = h->points.data();
  result->points_off = h->points_off.data();
  result->edit_distance = h->edit.data();
  st.wavefront_bytes = st.cells * sizeof(int32_t);
  st.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (stats) *stats = st;
  return VGA_OK;
}

}  // extern "C"

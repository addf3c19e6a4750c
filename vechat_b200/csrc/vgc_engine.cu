// vgc_engine.cu — C-ABI (include/vgc.h) + the sm_100a kernels of the lockstep POA engine.
//
// A window is a resumable program (poa_core.h: WinState + Poa::step_*) whose state lives in HBM; one lockstep
// cycle advances every live window of a stream group by one step of Window::generate_consensus:
//   update_kernel  AddAlignment / round end (fold weights, prune) / emit + plan   one warp per window, 4 per CTA
//   sort_kernel    LargestSubgraph, TopologicalSort, row program; hands the pending alignments to the job list
//                                                                                one warp per window, graph staged in smem
//   align_kernel   one CTA (one warp) per ALIGNMENT: DP fill (poa_fill.cuh, rows in registers) followed by the
//                  warp-cooperative traceback (poa_trace.cuh) of the matrix it has just written
// The DP matrix of an alignment never outlives its CTA, so it lives in a pool of scratch buffers sized by what can
// be resident at once (a few per SM), not in per-window HBM.  In the build phase a window has one alignment per
// cycle; in a re-alignment round (the graph is frozen, AddWeights only adds to edge weights) all nseq alignments of
// the window run concurrently and add their weights atomically, so a depth-30 window takes ~35 cycles, not 93.
//
// Replaces Polisher::polish's per-window lambda (reference src/polisher.cpp:498-516) for a whole batch.
// No CPU fallback: without a usable device every entry point fails with VGC_ERR_NO_DEVICE.

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up under nsys / ncu --nvtx, and cost nothing otherwise
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_prep.h"
#include "poa_core.h"
#include "poa_fill.cuh"
#include "poa_trace.cuh"
#include "poa_wide.cuh"
#include "vgc.h"

namespace vgc {

constexpr unsigned kSetupThreads = 4;  // host threads of the per-window work at the start of a pass
constexpr int kSmemHeader = 832;  // Slot + WinState copies
constexpr int kAlignHeader = 512; // align kernels keep only the Slot copy

// One pending alignment: position of its window in the pass's work list, and layer | flags.
struct Job {
  uint32_t idx;
  uint32_t layer;  // bits 0-29 layer id, bit 30 = part of a re-alignment round (AddWeights fused), bit 31 = SW mode
};
constexpr uint32_t kRecRing = 32;  // row records staged in shared memory by the fill
constexpr uint32_t kJobRound = 1u << 30, kJobSW = 1u << 31, kJobLayerMask = kJobRound - 1u;
// Alignments are dealt to one kernel per (row width, mode) class so that every kernel holds exactly one fill and one
// traceback variant (own register allocation, small code): class = width index * 2 + (SW ? 1 : 0)
// plus one list for the wide path (poa_wide.cuh: int32 cells, any width), served by a persistent kernel
constexpr int kFastClasses = 6, kClassWide = 6, kClasses = 7;
__host__ __device__ inline uint32_t job_class(uint32_t len, bool sw) {
  return (len <= 512u ? 0u : (len <= 640u ? 1u : 2u)) * 2u + (sw ? 1u : 0u);
}

// Arguments shared by the kernels of a lockstep pass.  `idx` below is the position of a window in the
// pass's work list (windows ordered group by group, inside a group by decreasing number of cycles).
struct KernelArgs {
  BatchView bv;
  const uint32_t* work;    // [n] window ids
  const Slot* slots;       // [n] scratch slot of work[idx]
  WinState* wstates;       // [n] resumable program state of work[idx]
  uint8_t* out;            // output bytes (bv.out_off / out_cap index into it)
  uint32_t* out_len;       // [n_windows]
  uint32_t* status;        // [n_windows]
  unsigned long long* totals;  // [2 + kPhCount + 2]: cells, alignments, per-phase cycles, sorts, sorts out of HBM
  Scores nw;
  uint32_t haplotype, trim, num_prune;
  double min_confidence, min_support;
  uint32_t smem_bytes;     // dynamic shared memory per CTA of the launched kernel
  // align kernel: pool of DP-matrix buffers, pool_per_sm per SM
  uint8_t* pool;
  uint32_t* pool_free;     // [pool_n] ring of free buffer ids, then [pool_n] = take counter, [pool_n + 1] = give counter
  unsigned long long pool_buf_bytes;
  unsigned long long pool_fc_off;  // byte offset of the first-column values inside a buffer
  uint32_t pool_rows;      // rows a buffer holds (at the widest row)
  uint32_t pool_n;         // buffers = CTAs of align_kernel the device can hold at once
  // wide path: its own pool of int32 matrices (same ticket ring), wide_n buffers of wide_buf_bytes
  uint8_t* wide_pool;
  uint32_t* wide_free;
  unsigned long long wide_buf_bytes;
  uint32_t wide_n;
  // diagnostics (VGC_TIMELINE=file): one record per warp-sized unit of work {start ns, duration ns, SM | kind << 16}
  uint4* tl;               // [1 + tl_cap]; tl[0].x = records taken
  uint32_t tl_cap;
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_record(const KernelArgs& a, uint32_t kind, unsigned long long t0) {
  if (!a.tl) return;
  const unsigned long long t1 = gtime_ns();
  uint32_t sm;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  const uint32_t i = atomicAdd(&a.tl[0].x, 1u);
  if (i < a.tl_cap) a.tl[1 + i] = make_uint4(static_cast<uint32_t>(t0), static_cast<uint32_t>(t0 >> 32), static_cast<uint32_t>(t1 - t0), sm | (kind << 16));
}

// Executor: the device side of poa_core.h's `Ex` concept.  G lanes work on one window (G = 32: a whole warp).
template <int K, int G = 32>
struct WarpEx {
  uint8_t* sm;        // this window's shared memory after the header
  uint32_t sm_bytes;
  uint32_t max_len;
  int lane_;          // lane inside the group
  uint32_t mask_;     // the group's lanes inside the warp

  __device__ __forceinline__ unsigned long long clock() const { return clock64(); }
  __device__ __forceinline__ int lane() const { return lane_; }
  __device__ __forceinline__ int width() const { return G; }
  __device__ __forceinline__ bool leader() const { return lane_ == 0; }
  __device__ __forceinline__ void sync() { __syncwarp(mask_); }
  __device__ __forceinline__ uint32_t atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
  __device__ __forceinline__ uint32_t bcast(uint32_t v, uint32_t src) { return __shfl_sync(mask_, v, src, G); }
  __device__ __forceinline__ uint32_t ballot(bool p) { return __ballot_sync(mask_, p); }
  // never used on the device: the traceback runs in align_kernel (poa_trace.cuh)
  __device__ __forceinline__ void trace_tile(uint32_t** th, U4** tr) {
    *th = nullptr;
    *tr = nullptr;
  }
  __device__ __forceinline__ uint32_t reduce_min(uint32_t v) {
    if constexpr (G == 32) {
      return __reduce_min_sync(0xFFFFFFFFu, v);
    } else {
#pragma unroll
      for (int d = G / 2; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(mask_, v, d, G));
      return v;
    }
  }
  __device__ __forceinline__ uint32_t reduce_max(uint32_t v) {
    if constexpr (G == 32) {
      return __reduce_max_sync(0xFFFFFFFFu, v);
    } else {
#pragma unroll
      for (int d = G / 2; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(mask_, v, d, G));
      return v;
    }
  }
  __device__ __forceinline__ unsigned long long reduce_add64(unsigned long long v) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(mask_, v, d, G);
    return v;
  }
  __device__ __forceinline__ uint32_t excl_scan(uint32_t v, uint32_t* total) {
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      uint32_t t = __shfl_up_sync(mask_, x, d, G);
      if (lane_ >= d) x += t;
    }
    *total = __shfl_sync(mask_, x, G - 1, G);
    return x - v;
  }
  // layout: codes[max_len] | arena.  The arena holds the sort's staged graph:
  //   rec32[nV] | adj16[nA] | stack16[>= 256]
  __device__ __forceinline__ uint8_t* seq_codes() { return sm; }
  __device__ __forceinline__ uint8_t* arena() { return sm + ((max_len + 15u) & ~15u); }
  __device__ __forceinline__ uint32_t arena_bytes() { return sm_bytes - ((max_len + 15u) & ~15u); }
  __device__ bool stage_fast(uint32_t nV, uint32_t nA, uint32_t** r, uint16_t** t, uint16_t** s, uint32_t* cap) {
    if (nV >= 65535u || nA >= 65535u) return false;
    const uint32_t to = nV * 4u;
    const uint32_t so = to + ((nA * 2u + 3u) & ~3u);
    const uint32_t avail = arena_bytes();
    if (so + 512u > avail) return false;
    uint8_t* base = arena();
    *r = reinterpret_cast<uint32_t*>(base);
    *t = reinterpret_cast<uint16_t*>(base + to);
    *s = reinterpret_cast<uint16_t*>(base + so);
    *cap = (avail - so) / 2u;
    return true;
  }
  // order_update's staged dirty blocks
  __device__ __forceinline__ void block_arena(uint32_t** base, uint32_t* bytes) {
    *base = reinterpret_cast<uint32_t*>(arena());
    *bytes = arena_bytes();
  }
  // LargestSubgraph's staged live adjacency: off16[nV+1] | adj16[nA] | visited[nV] | stack16[>= 256]
  __device__ bool stage_lsg(uint32_t nV, uint32_t nA, uint16_t** o, uint16_t** t, uint8_t** vis, uint16_t** s,
                            uint32_t* cap) {
    if (nV >= 65535u || nA >= 65535u) return false;
    const uint32_t to = ((nV + 1u) * 2u + 3u) & ~3u;
    const uint32_t vo = to + ((nA * 2u + 3u) & ~3u);
    const uint32_t so = vo + ((nV + 3u) & ~3u);
    const uint32_t avail = arena_bytes();
    if (so + 512u > avail) return false;
    uint8_t* base = arena();
    *o = reinterpret_cast<uint16_t*>(base);
    *t = reinterpret_cast<uint16_t*>(base + to);
    *vis = base + vo;
    *s = reinterpret_cast<uint16_t*>(base + so);
    *cap = (avail - so) / 2u;
    return true;
  }
  // the fill runs in align_kernel: never called through the executor on the device
  template <int KK>
  __device__ __forceinline__ void fill(Slot&, WinState&, const uint8_t*, uint32_t, uint32_t, const Scores&, uint32_t) {}
};

__device__ __forceinline__ void copy_words(void* dst, const void* src, uint32_t bytes, int lane, int width = 32) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (uint32_t i = lane; i < bytes / 4; i += width) d[i] = s[i];
}

// ---- the kernels of a lockstep cycle ---------------------------------------------------------------------
// CTAs per SM each kernel is compiled and sized for (registers via __launch_bounds__, shared memory via the
// dynamic size the host passes): the update is parallel and light, the sort is a serial DFS over a graph staged in
// shared memory, the align kernel is register-heavy (a DP row lives in registers).
#ifndef VGC_UPDATE_CTAS
#define VGC_UPDATE_CTAS 24
#endif
#ifndef VGC_SORT_CTAS
#define VGC_SORT_CTAS 12
#endif
#ifndef VGC_ALIGN_CTAS
#define VGC_ALIGN_CTAS 16
#endif

struct WinCtx {
  Slot* sl;
  WinState* ws;
  WinState* gws;
  uint32_t idx, w;
  int lane, width;
  uint32_t mask;
};

// Load the window's Slot + WinState into its shared-memory header; false if the window does not wait for `need`.
// `lane`/`width`/`mask`: the lanes working on this window (a whole warp unless the kernel packs several per warp).
__device__ __forceinline__ bool win_enter(const KernelArgs& a, uint32_t idx, uint32_t need, uint8_t* hdr, WinCtx* c,
                                          int lane = threadIdx.x, int width = 32, uint32_t mask = 0xFFFFFFFFu) {
  c->idx = idx;
  c->lane = lane;
  c->width = width;
  c->mask = mask;
  c->gws = a.wstates + idx;
  if (c->gws->pc == kPcDone || c->gws->need != need) return false;
  c->sl = reinterpret_cast<Slot*>(hdr);
  c->ws = reinterpret_cast<WinState*>(hdr + ((sizeof(Slot) + 15) & ~size_t(15)));
  copy_words(c->sl, a.slots + idx, sizeof(Slot), lane, width);
  copy_words(c->ws, c->gws, sizeof(WinState), lane, width);
  __syncwarp(mask);
  c->w = a.work[idx];
  return true;
}

// Write the program state (and the graph headers that live in the Slot copy) back; publish a finished window.
__device__ __forceinline__ void win_leave(const KernelArgs& a, const WinCtx& c) {
  __syncwarp(c.mask);
  if (c.lane == 0) {
    Slot* gs = const_cast<Slot*>(a.slots + c.idx);
    gs->g[0].nV = c.sl->g[0].nV;
    gs->g[0].nE = c.sl->g[0].nE;
    gs->g[1].nV = c.sl->g[1].nV;
    gs->g[1].nE = c.sl->g[1].nE;
    gs->r2n = c.sl->r2n;  // the incremental order swaps its buffers
    gs->r2n2 = c.sl->r2n2;
    gs->bstart = c.sl->bstart;
    gs->bstart2 = c.sl->bstart2;
  }
  copy_words(c.gws, c.ws, sizeof(WinState), c.lane, c.width);
  if (c.lane == 0 && c.ws->pc == kPcDone) {
    a.status[c.w] = c.ws->status;
    if (c.ws->status != kStOk) return;  // a window that is re-run counts once, when it completes
    atomicAdd(a.totals, c.ws->cells);
    atomicAdd(a.totals + 1, static_cast<unsigned long long>(c.ws->alignments));
    for (int i = 0; i < kPhCount; ++i) atomicAdd(a.totals + 2 + i, c.ws->phase[i]);
    atomicAdd(a.totals + 2 + kPhCount, static_cast<unsigned long long>(c.ws->sorts));
    atomicAdd(a.totals + 3 + kPhCount, static_cast<unsigned long long>(c.ws->sorts_hbm));
  }
}

template <int K>
__device__ __forceinline__ WarpEx<K> make_ex(const KernelArgs& a, uint8_t* smem, const Slot* sl) {
  WarpEx<K> ex;
  ex.sm = smem + kSmemHeader;
  ex.sm_bytes = a.smem_bytes - kSmemHeader;
  ex.max_len = sl->max_len;
  ex.lane_ = threadIdx.x & 31;
  ex.mask_ = 0xFFFFFFFFu;
  return ex;
}

// U: graph update + phase transitions + choice of the next alignment (Graph::AddAlignment, the end of a
// re-alignment round, PruneGraph, GenerateCorrectedSequence / GenerateConsensus, and Window::generate_consensus'
// control flow).
// hand the pending alignment(s) of a window that has just been prepared to the job list of their class
// (jobs: kClasses lists of job_cap entries, njobs: their fill counts)
template <class P>
__device__ __forceinline__ void emit_jobs(const KernelArgs& a, const WinCtx& c, const P& poa, Job* jobs, uint32_t job_cap,
                                          uint32_t* njobs) {
  if (c.ws->pc == kPcDone || c.ws->need != kNeedFill) return;
  auto emit = [&](uint32_t layer, bool sw, uint32_t flags) {
    const uint32_t len = static_cast<uint32_t>(a.bv.seq_off[layer + 1] - a.bv.seq_off[layer]);
    const uint32_t cls = c.ws->wide ? static_cast<uint32_t>(kClassWide) : job_class(len, sw);
    Job jb;
    jb.idx = c.idx;
    jb.layer = layer | flags | (sw ? kJobSW : 0u);
    jobs[static_cast<size_t>(cls) * job_cap + atomicAdd(njobs + cls, 1u)] = jb;
  };
  if (c.ws->round) {
    const uint32_t nseq = a.bv.win_nseq[c.w];
    const uint32_t* rank = a.bv.layer_rank + a.bv.win_first[c.w];
    for (uint32_t j = c.lane; j < nseq; j += c.width) emit(rank[j], poa.round_mode(c.w, j) == kModeSW, kJobRound);
  } else if (c.lane == 0) {
    emit(c.ws->fill_layer, c.ws->fill_mode == kModeSW, 0u);
  }
}

constexpr int kUpdateWins = 4;  // windows (= warps) per CTA: the SM's 32-CTA limit must not cap the light kernel
template <int K>
__global__ void __launch_bounds__(32 * kUpdateWins) update_kernel(const KernelArgs a, uint32_t base, uint32_t count,
                                                                  Job* jobs, uint32_t job_cap, uint32_t* njobs) {
  extern __shared__ __align__(16) uint8_t smem_all[];
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * kUpdateWins + warp;
  if (i >= count) return;
  uint8_t* smem = smem_all + warp * a.smem_bytes;
  WinCtx c;
  const unsigned long long tl0 = a.tl ? gtime_ns() : 0ull;
  if (!win_enter(a, base + i, kNeedUpdate, smem, &c, threadIdx.x & 31)) return;
  WarpEx<K> ex = make_ex<K>(a, smem, c.sl);
  Poa<WarpEx<K>, K> poa(ex, a.bv, *c.sl, *c.ws, a.nw);
  poa.step_update(c.w, a.haplotype != 0, a.trim != 0, a.min_confidence, a.min_support, a.num_prune,
                  a.out + a.bv.out_off[c.w], a.out_len + c.w);
  // the common case needs no sort (the order was maintained incrementally by AddAlignment): the row program is built
  // right here and the window skips the sort kernel
  if (c.ws->pc != kPcDone && c.ws->need == kNeedPrepare && !(c.ws->prep & (kPrepMainSort | kPrepSubSort | kPrepLargest))) {
    poa.step_prepare(c.w);
    emit_jobs(a, c, poa, jobs, job_cap, njobs);
  }
  win_leave(a, c);
  if (c.lane == 0) tl_record(a, 0, tl0);
}

// T: Graph::TopologicalSort (+ Subgraph view of a partial layer, LargestSubgraph after a prune) and the row program
// of the next alignment(s); then the pending alignments go to the group's job list for align_kernel.
template <int K>
__global__ void __launch_bounds__(32, VGC_SORT_CTAS) sort_kernel(const KernelArgs a, uint32_t base, Job* jobs,
                                                                 uint32_t job_cap, uint32_t* njobs) {
  extern __shared__ __align__(16) uint8_t smem[];
  WinCtx c;
  const unsigned long long tl0 = a.tl ? gtime_ns() : 0ull;
  if (!win_enter(a, base + blockIdx.x, kNeedPrepare, smem, &c)) return;
  WarpEx<K> ex = make_ex<K>(a, smem, c.sl);
  Poa<WarpEx<K>, K> poa(ex, a.bv, *c.sl, *c.ws, a.nw);
  poa.step_prepare(c.w);
  emit_jobs(a, c, poa, jobs, job_cap, njobs);
  win_leave(a, c);
  if (c.lane == 0) tl_record(a, 1, tl0);
}

// A: one alignment = DP fill (replaces SimdAlignmentEngine::Linear's fill) + traceback, by one warp.
// shared memory: Slot header | codes[max_len] | recs[64 x 16 B] | fill: prof[num_codes x 32K words] | ring
//                                                                 | trace: tile | w2[len]   (overlays prof / ring)
__device__ __forceinline__ void win_fail(const KernelArgs& a, WinState* gws, uint32_t w, uint32_t st) {
  gws->status = st;
  gws->pc = kPcDone;
  gws->need = kNeedNone;
  a.status[w] = st;
}

template <int KR, bool SW>
__global__ void __launch_bounds__(32, VGC_ALIGN_CTAS) align_kernel(const KernelArgs a, const Job* jobs,
                                                                   const uint32_t* njobs) {
  extern __shared__ __align__(16) uint8_t smem[];
  if (blockIdx.x >= *njobs) return;
  const unsigned long long tl0 = a.tl ? gtime_ns() : 0ull;
  const unsigned long long t0 = clock64();
  const int lane = threadIdx.x;
  const Job jb = jobs[blockIdx.x];
  const uint32_t idx = jb.idx, l = jb.layer & kJobLayerMask;
  const bool round = (jb.layer & kJobRound) != 0;
  constexpr uint32_t mode = SW ? kModeSW : kModeNW;
  WinState* gws = a.wstates + idx;
  if (gws->pc == kPcDone) return;  // another alignment of the round failed the window
  Slot* sl = reinterpret_cast<Slot*>(smem);
  copy_words(sl, a.slots + idx, sizeof(Slot), lane);
  const uint32_t nR = gws->nR, sub = gws->sub, cur = gws->cur;
  const uint32_t w = a.work[idx];
  __syncwarp();
  // ---- a DP-matrix buffer from the pool: a ring of free ids with take / give tickets.  There are as many buffers as
  //      CTAs of this kernel can be resident on the device, so ticket h finds the id given back by ticket h - pool_n
  //      (or the initial fill) in slot h % pool_n; it only ever waits for that store to land.
  uint32_t buf = 0, pool_spins = 0, slow_steps = 0;
  if (lane == 0) {
    const uint32_t ticket = atomicAdd(a.pool_free + a.pool_n, 1u);
    uint32_t* slot = a.pool_free + ticket % a.pool_n;
    while ((buf = atomicExch(slot, kNone)) == kNone) ++pool_spins;
  }
  buf = __shfl_sync(0xFFFFFFFFu, buf, 0);
  uint8_t* pb = a.pool + static_cast<unsigned long long>(buf) * a.pool_buf_bytes;
  uint8_t* sm = smem + kAlignHeader;
  const uint32_t max_len = min(sl->max_len, 64u * KR);  // this class takes no longer layer
  uint8_t* codes = sm;
  U4* recs = reinterpret_cast<U4*>(sm + ((max_len + 15u) & ~15u));
  uint32_t* prof = reinterpret_cast<uint32_t*>(recs + kRecRing);
  const uint64_t o = a.bv.seq_off[l];
  const uint32_t len = static_cast<uint32_t>(a.bv.seq_off[l + 1] - o);
  for (uint32_t i = lane; i < len; i += 32) codes[i] = static_cast<uint8_t>(base_code(a.bv, o + i));
  __syncwarp();
  Scores sw;
  sw.m = 3;  // the SW engine is hard-wired to 3/-5/-4 (window.cpp:326)
  sw.x = -5;
  sw.g = -4;
  const Scores sc = SW ? sw : a.nw;
  FillIo io;
  io.H = reinterpret_cast<uint32_t*>(pb);
  io.fc = reinterpret_cast<int16_t*>(pb + a.pool_fc_off);
  io.rowprog = sl->rowprog;
  io.ovf = sl->ovf;
  io.nR = nR;
  io.best_row = io.best_col = 0;
  io.best_score = 0;
  int st = kWalkDone;
  uint32_t n = 0, refills = 0;
  unsigned long long t1 = t0;
  if (nR + 1 > a.pool_rows || len > max_len) {
    st = kWalkBad;
  } else {
    // ---- fill (rows of KR words per lane; the ring of recent rows only if the shared memory holds it)
    uint32_t* ring = prof + a.bv.num_codes * RowMap<KR>::kWords;
    const uint32_t used = static_cast<uint32_t>(reinterpret_cast<uint8_t*>(ring) - smem);
    const int ring_rows = used + kRingRows * (RowMap<KR>::kWords * 4 + 128) <= a.smem_bytes ? kRingRows : 0;
    warp_fill_t<KR, SW>(io, codes, len, sc, a.bv.num_codes, prof, reinterpret_cast<uint4*>(recs), ring, ring_rows);
    t1 = clock64();
    // ---- traceback (the profile and the ring are dead: the tile and the weights take their place)
    __syncwarp();
    TraceIo t;
    t.H = io.H;
    t.fc = io.fc;
    t.rp = reinterpret_cast<const U4*>(sl->rowprog);
    t.ovf = sl->ovf;
    t.nodes = sl->max_nodes < 65536u ? nullptr : (sub ? sl->order : sl->r2n);
    t.codes = codes;
    t.m = sc.m;
    t.x = sc.x;
    t.g = sc.g;
    t.sw = SW;
    t.row = io.best_row;
    t.col = io.best_col;
    t.max_steps = nR + len + 2;
    t.aln_node = sl->aln_node;
    t.aln_pos = sl->aln_pos;
    t.aln_cap = sl->aln_cap;
    t.wacc = sl->wacc;
    t.ew = sl->g[cur].ew;
    t.ieid = sl->g[cur].ieid;
    t.in_stride = sl->in_stride;
    uint32_t* tile = prof;
    uint32_t* w2 = tile + kTraceTileBytes / 4;
    t.w2 = w2;
    if (round) {
      const bool hq = a.bv.has_qual[l] != 0;
      const uint8_t* quals = a.bv.quals + o;
      for (uint32_t i = lane; i < len; i += 32) {
        uint32_t v = 0;
        if (i >= 1) v = hq ? a.bv.wlut[quals[i - 1]] + a.bv.wlut[quals[i]] : 2u;
        w2[i] = v;
      }
      __syncwarp();
      st = warp_trace<KR, true>(t, tile, &n, &refills, &slow_steps);
    } else {
      st = warp_trace<KR, false>(t, tile, &n, &refills, &slow_steps);
    }
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    atomicExch(a.pool_free + atomicAdd(a.pool_free + a.pool_n + 1, 1u) % a.pool_n, buf);
    const unsigned long long t2 = clock64();
    // diagnostics: slowest fill / traceback of the call, pool waits
    atomicMax(a.totals + 4 + kPhCount, t1 - t0);
    atomicMax(a.totals + 5 + kPhCount, t2 - t1);
    atomicAdd(a.totals + 6 + kPhCount, static_cast<unsigned long long>(slow_steps));
    atomicAdd(a.totals + 7 + kPhCount, static_cast<unsigned long long>(pool_spins));
    tl_record(a, 2u + (SW ? 1u : 0u) + (round ? 2u : 0u), tl0);
    if (st != kWalkDone) {
      win_fail(a, gws, w, kStInternal);
    } else if (round) {
      atomicAdd(&gws->phase[kPhFill], t1 - t0);
      atomicAdd(&gws->phase[kPhTrace], t2 - t1);
      atomicAdd(&gws->phase[kPhOther], static_cast<unsigned long long>(refills));
      __threadfence();
      if (atomicAdd(&gws->jobs_done, 1u) + 1u == gws->jobs_total) gws->need = kNeedUpdate;
    } else {
      gws->aln_len = n;
      gws->best_row = io.best_row;
      gws->best_col = io.best_col;
      gws->best_score = io.best_score;
      gws->phase[kPhFill] += t1 - t0;
      gws->phase[kPhTrace] += t2 - t1;
      gws->phase[kPhOther] += refills;
      gws->need = kNeedUpdate;
    }
  }
}

// W: the wide path — same contract as align_kernel for the alignments of the wide list (int32 matrix of any width in
// the wide pool, poa_wide.cuh + wide_trace).  Persistent: the host does not know how many wide alignments a cycle has
// (the score-range test depends on the graph), so a fixed grid strides over the list.
// shared memory: Slot header | codes[max_len]
__global__ void __launch_bounds__(32) align_wide_kernel(const KernelArgs a, const Job* jobs, const uint32_t* njobs) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x;
  const uint32_t nj = *njobs;
  for (uint32_t jn = blockIdx.x; jn < nj; jn += gridDim.x) {
    __syncwarp();
    const unsigned long long tl0 = a.tl ? gtime_ns() : 0ull;
    const unsigned long long t0 = clock64();
    const Job jb = jobs[jn];
    const uint32_t idx = jb.idx, l = jb.layer & kJobLayerMask;
    const bool round = (jb.layer & kJobRound) != 0, swm = (jb.layer & kJobSW) != 0;
    WinState* gws = a.wstates + idx;
    if (gws->pc == kPcDone) continue;
    Slot* sl = reinterpret_cast<Slot*>(smem);
    copy_words(sl, a.slots + idx, sizeof(Slot), lane);
    const uint32_t nR = gws->nR, sub = gws->sub, cur = gws->cur;
    const uint32_t w = a.work[idx];
    __syncwarp();
    uint32_t buf = 0;
    if (lane == 0) {
      const uint32_t ticket = atomicAdd(a.wide_free + a.wide_n, 1u);
      uint32_t* slot = a.wide_free + ticket % a.wide_n;
      while ((buf = atomicExch(slot, kNone)) == kNone) __nanosleep(200);
    }
    buf = __shfl_sync(0xFFFFFFFFu, buf, 0);
    int32_t* H = reinterpret_cast<int32_t*>(a.wide_pool + static_cast<unsigned long long>(buf) * a.wide_buf_bytes);
    uint8_t* codes = smem + kAlignHeader;
    const uint64_t o = a.bv.seq_off[l];
    const uint32_t len = static_cast<uint32_t>(a.bv.seq_off[l + 1] - o);
    for (uint32_t i = lane; i < len; i += 32) codes[i] = static_cast<uint8_t>(base_code(a.bv, o + i));
    __syncwarp();
    Scores sw;
    sw.m = 3;  // the SW engine is hard-wired to 3/-5/-4 (window.cpp:326)
    sw.x = -5;
    sw.g = -4;
    const Scores sc = swm ? sw : a.nw;
    int st = kWalkDone;
    uint32_t n = 0;
    unsigned long long t1 = t0;
    WideFillIo io;
    io.best_row = io.best_col = 0;
    io.best_score = 0;
    uint32_t fail_st = kStInternal;
    if ((static_cast<unsigned long long>(nR) + 1) * (len + 1ull) * 4ull > a.wide_buf_bytes) {
      st = kWalkBad;
      fail_st = kStWideCapacity;
    } else {
      io.H = H;
      io.cols = len + 1;
      io.rp = reinterpret_cast<const U4*>(sl->rowprog);
      io.ovf = sl->ovf;
      io.nR = nR;
      if (swm) warp_fill_wide<true>(io, codes, len, sc);
      else warp_fill_wide<false>(io, codes, len, sc);
      t1 = clock64();
      WideIo t;
      t.H = H;
      t.cols = len + 1;
      t.rp = io.rp;
      t.ovf = sl->ovf;
      t.nodes = sl->max_nodes < 65536u ? nullptr : (sub ? sl->order : sl->r2n);
      t.codes = codes;
      t.m = sc.m;
      t.x = sc.x;
      t.g = sc.g;
      t.sw = swm;
      t.row = io.best_row;
      t.col = io.best_col;
      t.max_steps = nR + len + 2;
      t.aln_node = sl->aln_node;
      t.aln_pos = sl->aln_pos;
      t.aln_cap = sl->aln_cap;
      t.ew = sl->g[cur].ew;
      t.ieid = sl->g[cur].ieid;
      t.in_stride = sl->in_stride;
      t.quals = a.bv.has_qual[l] ? a.bv.quals + o : nullptr;
      t.wlut = a.bv.wlut;
      struct {
        int lane;
        __device__ bool leader() const { return lane == 0; }
        __device__ uint32_t atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
      } wex{lane};
      st = round ? wide_trace<true>(wex, t, &n) : wide_trace<false>(wex, t, &n);
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicExch(a.wide_free + atomicAdd(a.wide_free + a.wide_n + 1, 1u) % a.wide_n, buf);
      const unsigned long long t2 = clock64();
      tl_record(a, 6u, tl0);
      if (st != kWalkDone) {
        win_fail(a, gws, w, fail_st);
      } else if (round) {
        atomicAdd(&gws->phase[kPhFill], t1 - t0);
        atomicAdd(&gws->phase[kPhTrace], t2 - t1);
        __threadfence();
        if (atomicAdd(&gws->jobs_done, 1u) + 1u == gws->jobs_total) gws->need = kNeedUpdate;
      } else {
        gws->aln_len = n;
        gws->best_row = io.best_row;
        gws->best_col = io.best_col;
        gws->best_score = io.best_score;
        gws->phase[kPhFill] += t1 - t0;
        gws->phase[kPhTrace] += t2 - t1;
        gws->need = kNeedUpdate;
      }
    }
  }
}

static_assert(sizeof(Slot) + sizeof(WinState) + 32 <= kSmemHeader, "shared-memory header too small");
static_assert(sizeof(Slot) <= kAlignHeader, "align kernel header too small");

}  // namespace vgc

// =====================================================================================================
// host side
// =====================================================================================================
namespace {

thread_local std::string g_err;

// NVTX range for the phases of a call (pack / H2D / pass / D2H + stitch), SURVEY.md §5
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// A pass runs up to 48 stream groups side by side.  The driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware
// queues (default 8); streams that share a queue serialise on each other's dependencies (measured: 31.3 k -> 34.2 k
// windows/s with 32 queues).  The variable is read when the CUDA context is created, so it is set when the library
// is loaded (never overriding the user's choice); hosts that create the context first set it themselves
// (vechat_b200/__init__.py).
struct ConnectionsInit {
  ConnectionsInit() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_connections_init;

void set_err(const std::string& s) { g_err = s; }

#define VGC_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      set_err(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call);          \
      return VGC_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return VGC_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e != cudaSuccess) {
      p = nullptr;
      set_err(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
      cudaGetLastError();
      return VGC_ERR_NOMEM;
    }
    cap = want;
    return VGC_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

// One staged batch: device copies of the caller's arrays + what the host prepared (host_prep.h).
struct InputSet {
  DevBuf d_bases, d_quals, d_seq_off, d_has_qual, d_begin, d_end, d_win_first, d_win_flags;
  DevBuf d_rank, d_nseq, d_avgw, d_out_off, d_out_cap, d_tables;
  vgc::Prepared prep;
  uint32_t base_bits = 8;        // bases in HBM: 2 bits each (A C G T only), 4 bits (up to 16 codes)
  uint8_t* h_packed = nullptr;   // pinned staging of the packed bases
  size_t h_packed_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // first / last H2D of the staging, copy stream
  // staging in flight (vgc_submit)
  std::thread worker;
  vgc_batch batch;               // the caller's arrays (must stay valid until vgc_collect returns)
  int rc = VGC_OK;
  std::string err;
  uint64_t in_bytes = 0;
  double prep_ms = 0.0, pack_ms = 0.0;
};

struct vgc_engine {
  int device = 0;
  vgc_params params;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double phase_cycles[16] = {0};  // last call: leader-lane cycles per kPh* phase, summed over windows
  double pool_spins = 0.0;      // last call: failed attempts to take a DP buffer (0 unless the residency bound is off)
  double launch_ms = 0.0;       // host wall time spent enqueueing the kernel launches of the current call
  double setup_ms = 0.0;        // host wall time from the start of a pass to its first launch (slots, groups, pools)
  double pass_kernel_ms = 0.0;  // device time of the POA kernel launches of the current call (events 6/7)
  int sm_count = 0;
  int groups = 48;                // streams of a lockstep pass (upper bound)
  uint32_t node_share_div = 6;    // first-pass node capacity = backbone + (sum of layer lengths) / this + one layer
  double sort_growth = 0.055;     // new graph nodes per base added, upper estimate (sizes the sort kernel's shared memory)
  unsigned host_threads = 16;     // host threads of a staging (prepare + packing): VGC_HOST_THREADS, else the cores
                                  // divided by the processes sharing the host (LOCAL_WORLD_SIZE), at most 16
  int launch_threads = 1;         // host threads enqueueing the launches of a pass (measured: the enqueue is not the limit)
  int group_mode = 2;             // 2: one group per number of fills, 1: equal contiguous blocks of the depth-sorted list, 0: round-robin
  cudaStream_t gstream[64] = {};
  cudaEvent_t gev[64] = {};
  uint32_t smem_update = 0, smem_sort = 0;
  size_t mem_budget = 0;
  bool mem_budget_fixed = false;  // VGC_MEM_BUDGET_MB given
  // device copies of a batch: two sets, so that vgc_submit can stage batch i + 1 (host prepare, packing, H2D on the
  // copy stream) while vgc_collect runs the kernels of batch i out of the other set
  InputSet in[2];
  InputSet* I = &in[0];          // the set the running / last pass reads
  int queued[2] = {-1, -1};      // submitted, not yet collected (oldest first)
  int n_queued = 0;
  std::atomic<bool> in_setup{false};  // a pass is between its start and its first launches: staging workers hold back
  std::mutex qmu;                // guards queued / n_queued / running / I: vgc_submit may come from another thread
  bool running = false;          // vgc_collect is inside its pass (another thread may call vgc_submit meanwhile)
  cudaStream_t copy_stream = nullptr;
  DevBuf d_work;
  DevBuf d_out, d_out_len, d_status, d_misc, d_slots, d_slot_mem, d_wstates;
  DevBuf d_tl;  // diagnostics timeline (VGC_TIMELINE)
  DevBuf d_wide, d_wide_free;  // wide path: pool of int32 DP matrices + its ring of free ids
  DevBuf d_pool, d_pool_busy, d_jobs, d_jobcnt;  // align kernel: DP-matrix pool, per-group job lists, per-cycle counters
  // host staging (pinned)
  uint8_t* h_out = nullptr;
  size_t h_out_cap = 0;
  uint32_t* h_out_len = nullptr;
  uint32_t* h_status = nullptr;
  size_t h_win_cap = 0;
  // resident batch (vgc_upload)
  bool resident = false;
  std::vector<uint8_t> backbone_copy;  // backbones of < 3-sequence windows (resident mode)
  std::vector<uint64_t> r_seq_off;
  std::vector<uint32_t> r_win_first;
  uint32_t r_n_windows = 0, r_n_layers = 0;
  uint64_t r_input_bytes = 0;
};

namespace {

using namespace vgc;

// basic shape checks needed before anything is read from the batch's arrays (prepare_batch repeats them)
bool batch_shape_ok(const vgc_batch* b) {
  const uint32_t nw = b->n_windows;
  if (!nw) return true;
  return b->win_first && b->seq_off && b->bases && b->begin && b->end && b->has_qual && b->win_flags &&
         b->win_first[nw] == b->n_layers;
}

// Bases travel packed: 2 bits per base while the batch holds A C G T only (their codes are 0-3 by construction,
// host_prep.h), 4 bits otherwise (up to 16 codes).  `coder` == nullptr: the speculative 2-bit pass (the alphabet
// prepare_batch finds decides whether it stands).  Bytes of the output are written by exactly one thread.
bool pack_bases(const uint8_t* bases, uint64_t nb, const uint8_t* coder, uint8_t* out, unsigned threads) {
  static const std::array<uint8_t, 256> acgt = [] {
    std::array<uint8_t, 256> t;
    t.fill(0xFF);
    t['A'] = 0;
    t['C'] = 1;
    t['G'] = 2;
    t['T'] = 3;
    return t;
  }();
  const uint64_t per = coder ? 2 : 4;  // bases per output byte
  const uint64_t nbytes = (nb + per - 1) / per;
  std::vector<uint8_t> bad(threads, 0);
  auto work = [&](unsigned t) {
    const uint64_t b0 = (nbytes * t / threads) & ~1ull, b1 = t + 1 == threads ? nbytes : ((nbytes * (t + 1) / threads) & ~1ull);
    uint8_t other = 0;
    for (uint64_t q = b0; q < b1; ++q) {
      const uint64_t i = q * per;
      uint32_t v = 0;
      if (!coder) {
        if ((q & 1u) == 0 && q + 1 < b1 && i + 8 <= nb) {
          // eight bases at once: (c >> 1) & 3 maps A C G T to 0 1 3 2, x ^ (x >> 1) puts them in code order 0 1 2 3;
          // then the eight 2-bit fields are squeezed together.  Whether the batch really holds A C G T only is
          // decided by the alphabet prepare_batch finds (the caller re-packs with 4 bits otherwise).
          uint64_t w8;
          std::memcpy(&w8, bases + i, 8);
          uint64_t x = (w8 >> 1) & 0x0303030303030303ull;
          x ^= (x >> 1) & 0x0101010101010101ull;
          x = (x | (x >> 6)) & 0x000F000F000F000Full;
          x = (x | (x >> 12)) & 0x000000FF000000FFull;
          x = (x | (x >> 24)) & 0xFFFFull;
          out[q] = static_cast<uint8_t>(x);
          out[q + 1] = static_cast<uint8_t>(x >> 8);
          ++q;
          continue;
        }
        for (uint64_t k = 0; k < 4 && i + k < nb; ++k) {
          const uint8_t c = acgt[bases[i + k]];
          other |= c;
          v |= static_cast<uint32_t>(c & 3u) << (2 * k);
        }
      } else {
        v = coder[bases[i]] & 15u;
        if (i + 1 < nb) v |= static_cast<uint32_t>(coder[bases[i + 1]] & 15u) << 4;
      }
      out[q] = static_cast<uint8_t>(v);
    }
    bad[t] = other & 0x80u;
  };
  if (threads <= 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  for (uint8_t x : bad)
    if (x) return false;
  return true;
}

// Stage one batch into an input set: H2D of the caller's arrays (qualities and layer tables as they are, bases
// packed), host preparation (rank sort, average weights, alphabet), H2D of the prepared tables — everything on the
// copy stream, finished (synchronised) on return.  Runs on the caller's thread (vgc_polish, vgc_upload) or on the
// set's worker thread (vgc_submit).
int stage_batch(vgc_engine* h, InputSet* in, const vgc_batch* b) {
  cudaSetDevice(h->device);
  const uint32_t nw = b->n_windows, nl = b->n_layers;
  uint64_t bytes = 0;
  auto put = [&](DevBuf& d, const void* src, size_t n) -> int {
    int rc = d.reserve(std::max<size_t>(n, 16));
    if (rc != VGC_OK) return rc;
    if (n) {
      cudaError_t e = cudaMemcpyAsync(d.p, src, n, cudaMemcpyHostToDevice, h->copy_stream);
      if (e != cudaSuccess) {
        set_err(std::string("H2D copy failed: ") + cudaGetErrorString(e));
        return VGC_ERR_CUDA;
      }
    }
    bytes += n;
    return VGC_OK;
  };
  int rc;
  in->in_bytes = 0;
  in->prep_ms = in->pack_ms = 0.0;
  VGC_CUDA(cudaEventRecord(in->ev0, h->copy_stream));
  const bool shape_ok = batch_shape_ok(b);
  const uint64_t nb = (shape_ok && nl) ? b->seq_off[nl] : 0;
  const unsigned pack_threads = nb >= (1u << 22) ? std::max(1u, std::min(8u, h->host_threads / 2)) : 1u;
  std::thread packer;
  bool acgt_only = false;
  double pack_ms = 0.0;
  if (shape_ok) {
    NvtxRange nvtx("vgc: H2D batch");
    // the bulk the caller owns goes first and overlaps the host-side preparation; the bases are packed meanwhile
    if ((rc = put(in->d_quals, b->quals, b->quals ? nb : 0))) return rc;
    if ((rc = put(in->d_seq_off, b->seq_off, (nl + 1) * sizeof(uint64_t)))) return rc;
    if ((rc = put(in->d_has_qual, b->has_qual, nl))) return rc;
    if ((rc = put(in->d_begin, b->begin, nl * 4ull))) return rc;
    if ((rc = put(in->d_end, b->end, nl * 4ull))) return rc;
    if ((rc = put(in->d_win_first, b->win_first, (nw + 1) * 4ull))) return rc;
    if ((rc = put(in->d_win_flags, b->win_flags, nw))) return rc;
    const size_t want = nb / 2 + 64;  // the 4-bit form is the larger one
    if (in->h_packed_cap < want) {
      if (in->h_packed) cudaFreeHost(in->h_packed);
      in->h_packed = nullptr;
      in->h_packed_cap = 0;
      VGC_CUDA(cudaMallocHost(&in->h_packed, want + want / 8));
      in->h_packed_cap = want + want / 8;
    }
    packer = std::thread([&, nb] {
      const auto t0 = std::chrono::steady_clock::now();
      acgt_only = pack_bases(b->bases, nb, nullptr, in->h_packed, pack_threads);
      pack_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    });
  }
  std::string err;
  const auto t0 = std::chrono::steady_clock::now();
  nvtxRangePushA("vgc: host prepare (rank sort, weights)");
  rc = vgc::prepare_batch(b, &h->params, &in->prep, &err, h->host_threads);
  nvtxRangePop();
  in->prep_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (packer.joinable()) packer.join();
  if (rc != VGC_OK) {
    cudaStreamSynchronize(h->copy_stream);  // the caller may free its buffers as soon as we return
    set_err(err);
    return rc;
  }
  Prepared& pr = in->prep;
  NvtxRange nvtx("vgc: H2D packed bases + prepared tables");
  if (acgt_only && pr.num_codes == 4) {
    in->base_bits = 2;
  } else {
    const auto t1 = std::chrono::steady_clock::now();
    pack_bases(b->bases, nb, pr.coder, in->h_packed, pack_threads);
    pack_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    in->base_bits = 4;
  }
  in->pack_ms = pack_ms;
  if ((rc = put(in->d_bases, in->h_packed, (nb * in->base_bits + 7) / 8))) return rc;
  if ((rc = put(in->d_rank, pr.layer_rank.data(), nl * 4ull))) return rc;
  if ((rc = put(in->d_nseq, pr.win_nseq.data(), nw * 4ull))) return rc;
  if ((rc = put(in->d_avgw, pr.win_avgw.data(), nw * 8ull))) return rc;
  if ((rc = put(in->d_out_off, pr.out_off.data(), nw * 8ull))) return rc;
  if ((rc = put(in->d_out_cap, pr.out_cap.data(), nw * 4ull))) return rc;
  // tables: coder[256] | decoder[16] | wlut[256]
  std::vector<uint8_t> tab(256 + 16 + 1024);
  std::memcpy(tab.data(), pr.coder, 256);
  std::memcpy(tab.data() + 256, pr.decoder, kMaxCodes);
  std::memcpy(tab.data() + 272, pr.wlut, 1024);
  if ((rc = put(in->d_tables, tab.data(), tab.size()))) return rc;
  VGC_CUDA(cudaEventRecord(in->ev1, h->copy_stream));
  cudaError_t e = cudaStreamSynchronize(h->copy_stream);  // `tab` and the prepared vectors must outlive the copies
  if (e != cudaSuccess) {
    set_err(std::string("H2D sync failed: ") + cudaGetErrorString(e));
    return VGC_ERR_CUDA;
  }
  in->in_bytes = bytes;
  return VGC_OK;
}

BatchView make_view(vgc_engine* h) {
  const InputSet& in = *h->I;
  BatchView v;
  v.bases = in.d_bases.as<uint8_t>();
  v.base_bits = in.base_bits;
  v.quals = in.d_quals.as<uint8_t>();
  v.seq_off = in.d_seq_off.as<uint64_t>();
  v.has_qual = in.d_has_qual.as<uint8_t>();
  v.begin = in.d_begin.as<uint32_t>();
  v.end = in.d_end.as<uint32_t>();
  v.win_first = in.d_win_first.as<uint32_t>();
  v.win_flags = in.d_win_flags.as<uint8_t>();
  v.layer_rank = in.d_rank.as<uint32_t>();
  v.win_nseq = in.d_nseq.as<uint32_t>();
  v.win_avgw = in.d_avgw.as<double>();
  v.out_off = in.d_out_off.as<uint64_t>();
  v.out_cap = in.d_out_cap.as<uint32_t>();
  v.coder = in.d_tables.as<uint8_t>();
  v.decoder = in.d_tables.as<uint8_t>() + 256;
  v.wlut = reinterpret_cast<const uint32_t*>(in.d_tables.as<uint8_t>() + 272);
  v.num_codes = in.prep.num_codes;
  return v;
}

constexpr int kMaxGroups = 64;

typedef void (*AlignFn)(const KernelArgs, const Job*, const uint32_t*);
inline AlignFn align_fn(int cls) {
  switch (cls) {
    case 0: return align_kernel<8, false>;
    case 1: return align_kernel<8, true>;
    case 2: return align_kernel<10, false>;
    case 3: return align_kernel<10, true>;
    case 4: return align_kernel<16, false>;
    default: return align_kernel<16, true>;
  }
}
constexpr uint32_t kClassK[kFastClasses] = {8, 8, 10, 10, 16, 16};

inline cudaError_t set_smem_attr(const void* f, uint32_t smem) {
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
}

template <int K>
int set_kernel_attrs(uint32_t smem_update, uint32_t smem_sort) {
  VGC_CUDA(set_smem_attr(reinterpret_cast<const void*>(update_kernel<K>), kUpdateWins * smem_update));
  VGC_CUDA(set_smem_attr(reinterpret_cast<const void*>(sort_kernel<K>), smem_sort));
  return VGC_OK;
}

// Shared memory of one align_kernel CTA: header | codes | stage | max(fill: profile [+ ring], trace: tile + weights).
// The ring of recent rows is dropped when it does not fit `budget` (what the CTAs-per-SM target leaves); a profile
// that does not fit either (many codes x wide rows) takes what it needs and fewer CTAs run per SM.
uint32_t align_smem(uint32_t K, uint32_t num_codes, uint32_t max_len, uint32_t budget) {
  const uint32_t fixed = kAlignHeader + ((max_len + 15u) & ~15u) + 16u * kRecRing;
  const uint32_t prof = num_codes * 128u * K;
  const uint32_t ring = kRingRows * (128u * K + 128u);
  const uint32_t trace = kTraceTileBytes + 4u * max_len + 16u;
  uint32_t body = prof + ring;
  if (fixed + body > budget) body = prof;
  body = std::max(body, trace);
  return (fixed + body + 255u) & ~255u;
}

// cycles (update launches) of a window's program and the alignments it hands to the align kernel in cycle c
// (poa_core.h step_update): build alignments one per cycle, then one cycle per re-alignment round (nseq alignments
// each), the final alignment, the emit; linear mode: build, consensus sort, emit.
inline uint32_t win_cycles(uint32_t nseq, bool haplotype, uint32_t num_prune) {
  return haplotype ? nseq + num_prune : nseq + 1;
}
inline uint32_t win_jobs(uint32_t nseq, bool haplotype, uint32_t num_prune, uint32_t c) {
  if (c + 1 < nseq) return 1;
  if (!haplotype) return 0;
  if (c + 1 < nseq + num_prune - 1) return nseq;
  return c + 2 == nseq + num_prune ? 1u : 0u;
}

// f(begin, end) over [0, n) on a few host threads (the per-window host work at the start of a pass is on the
// device's critical path: the GPU idles until the first launch)
template <class F>
void parallel_ranges(size_t n, unsigned nt, F f) {
  if (nt <= 1 || n < 4096) {
    f(static_cast<size_t>(0), n);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(f, n * t / nt, n * (t + 1) / nt);
  f(static_cast<size_t>(0), n / nt);
  for (auto& x : th) x.join();
}

// Node capacity of a window's slot on the first pass: backbone + a share of the layer bases (a read adds a node
// only where it disagrees with the graph) + one layer of head-room for AddAlignment's conservative check.  Windows
// that outgrow it are re-run with the exact upper bound (sum of layer lengths).
uint32_t estimate_nodes(const Prepared& pr, uint32_t w, uint32_t blen, bool exact, uint32_t share_div) {
  const uint64_t ub = static_cast<uint64_t>(pr.win_sum_len[w]) + 64;
  if (exact) return static_cast<uint32_t>(ub);
  // new nodes per read shrink as the graph saturates (measured: ~1 300 extra nodes at depth 30, ~1 550 at depth
  // 100 for 520-base layers at 15 % error): cap the share at five mean layer lengths
  const uint64_t nlay = pr.win_nseq[w] > 1 ? pr.win_nseq[w] - 1 : 1;
  const uint64_t grow = std::min<uint64_t>(pr.win_sum_len[w] / share_div, 5 * ((pr.win_sum_len[w] - blen) / nlay + 1));
  const uint64_t est = blen + grow + pr.win_max_len[w] + 64;
  return static_cast<uint32_t>(std::min(ub, est));
}

// One lockstep pass over `wins` (ordered by decreasing number of cycles): every window gets its own scratch slot
// (graph, orders, row program — no DP matrix), the windows are dealt to stream groups, and each group's stream runs
// cycle after cycle: update_kernel, sort_kernel over the live prefix, then align_kernel over the alignments the sort
// handed out.  Chunked by the memory budget.  Windows are independent, so no synchronisation other than stream order.
template <int K>
int run_pass_k(vgc_engine* h, const std::vector<uint32_t>& wins, bool exact, const uint64_t* seq_off,
               const uint32_t* win_first, uint32_t* launches) {
  NvtxRange nvtx(exact ? "vgc: POA pass (exact capacities)" : "vgc: POA pass");
  const auto t_entry = std::chrono::steady_clock::now();
  auto t_lap = t_entry;
  double laps[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto lap = [&](int i) {
    const auto now = std::chrono::steady_clock::now();
    laps[i] += std::chrono::duration<double, std::milli>(now - t_lap).count();
    t_lap = now;
  };
  const Prepared& pr = h->I->prep;
  const bool hap = h->params.haplotype != 0;
  const uint32_t num_prune = h->params.num_prune;
  const uint32_t ml = std::max<uint32_t>(pr.max_len, 16);
  // scratch budget of this pass: a share of what is free now plus what this engine already holds for scratch (other
  // engines of the process may have taken memory since vgc_create)
  if (!h->mem_budget_fixed) {
    size_t free_b = 0, total_b = 0;
    VGC_CUDA(cudaMemGetInfo(&free_b, &total_b));
    h->mem_budget = static_cast<size_t>((free_b + h->d_slot_mem.cap + h->d_pool.cap + h->d_wide.cap) * 0.70);
  }
  // ---- align kernels (one per class): shared memory, residency, pool geometry
  const uint32_t smem_budget = static_cast<uint32_t>(((228 * 1024 - VGC_ALIGN_CTAS * 1024) / VGC_ALIGN_CTAS) & ~255);
  uint32_t smem_align[kFastClasses];
  int rc;
  // every kernel stages the codes of a layer (up to ml bytes) behind its header: long layers (wide path) need room
  const uint32_t ml16 = (ml + 15u) & ~15u;
  const uint32_t smem_update = std::max<uint32_t>(h->smem_update, (kSmemHeader + ml16 + 3072u + 255u) & ~255u);
  const uint32_t smem_sort_cap = std::max<uint32_t>(h->smem_sort, (kSmemHeader + ml16 + 8192u + 255u) & ~255u);
  const uint32_t smem_wide = (kAlignHeader + ml16 + 255u) & ~255u;
  if ((rc = set_kernel_attrs<K>(smem_update, smem_sort_cap))) return rc;
  VGC_CUDA(set_smem_attr(reinterpret_cast<const void*>(align_wide_kernel), smem_wide));
  const int32_t max_abs_score = [&]() {
    int32_t a = 5;  // the SW engine's 3/-5/-4
    for (int32_t v : {h->params.match, h->params.mismatch, h->params.gap}) a = std::max(a, v < 0 ? -v : v);
    return a;
  }();
  int occ_max = 0;
  for (int cls = 0; cls < kFastClasses; ++cls) {
    smem_align[cls] = 0;
    if (kClassK[cls] > static_cast<uint32_t>(K)) continue;  // no layer of this batch is that wide
    smem_align[cls] = align_smem(kClassK[cls], pr.num_codes, std::min<uint32_t>(ml, 64u * kClassK[cls]), smem_budget);
    VGC_CUDA(set_smem_attr(reinterpret_cast<const void*>(align_fn(cls)), smem_align[cls]));
    int occ = 0;
    VGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_fn(cls), 32, smem_align[cls]));
    if (occ < 1) {
      set_err("align kernel does not fit an SM (shared memory)");
      return VGC_ERR_CAPACITY;
    }
    occ_max = std::max(occ_max, occ);
  }
  lap(0);
  size_t pos = 0;
  while (pos < wins.size()) {
    // ---- chunk: as many windows as the budget holds (the pool of DP buffers takes its share first)
    uint32_t pool_rows = 0;
    {
      // rows of the largest graph any window of the rest of the pass may reach
      for (size_t e = pos; e < wins.size(); ++e) {
        const uint32_t w = wins[e];
        const uint32_t f = win_first[w];
        pool_rows = std::max(pool_rows, estimate_nodes(pr, w, static_cast<uint32_t>(seq_off[f + 1] - seq_off[f]), exact, h->node_share_div));
      }
      pool_rows = std::max<uint32_t>(pool_rows, 64) + 1;
    }
    // a buffer = pool_rows rows of 32*K words + the first-column value (int16) of every row
    const uint64_t pool_fc_off = static_cast<uint64_t>(pool_rows) * (128ull * K);
    const uint64_t buf_bytes = align_up(pool_fc_off + 2ull * pool_rows + 64, 256);
    // one buffer per CTA the device can hold; the pool may take at most 40 % of the budget: with fewer buffers (long
    // rows: exact pass, deep windows) CTAs wait for a buffer to come back
    uint32_t per_sm = static_cast<uint32_t>(occ_max);
    const uint64_t pool_cap = h->mem_budget * 2 / 5;
    while (per_sm > 1 && static_cast<uint64_t>(per_sm) * h->sm_count * buf_bytes > pool_cap) --per_sm;
    if (static_cast<uint64_t>(per_sm) * h->sm_count * buf_bytes > h->mem_budget) {
      set_err("a window's DP matrix needs more scratch than the device memory budget");
      return VGC_ERR_CAPACITY;
    }
    uint64_t pool_bytes = static_cast<uint64_t>(per_sm) * h->sm_count * buf_bytes;
    lap(1);
    if ((rc = h->d_pool.reserve(pool_bytes))) return rc;
    // ---- wide path (poa_wide.cuh): can any alignment of the rest of the pass leave the fast kernels?  A layer beyond
    // their widest row, or the device's int16 score bound (poa_core.h step_prepare) evaluated with the slot's node
    // capacity as the number of rows.  Its int32 matrices live in their own pool, sized for the largest such window.
    uint32_t wide_n = 0;
    uint64_t wide_buf = 0;
    {
      uint64_t need = 0;
      for (size_t e = pos; e < wins.size(); ++e) {
        const uint32_t w = wins[e];
        const uint32_t f = win_first[w];
        const int64_t rows = estimate_nodes(pr, w, static_cast<uint32_t>(seq_off[f + 1] - seq_off[f]), exact, h->node_share_div) + 1;
        const uint32_t mlw = pr.win_max_len[w];
        const int64_t cols = 64ll * fill_width(K, mlw);
        if (mlw > 64u * K || (std::max(rows, cols) + cols + 2) * max_abs_score > 32000)
          need = std::max<uint64_t>(need, static_cast<uint64_t>(rows + 1) * (mlw + 1ull) * 4ull);
      }
      if (need) {
        wide_buf = align_up(need, 256);
        const uint64_t cap = h->mem_budget * 3 / 10;
        wide_n = static_cast<uint32_t>(std::min<uint64_t>(2ull * h->sm_count, cap / wide_buf));
        if (wide_n == 0) {
          if (wide_buf > h->mem_budget - pool_bytes) {
            set_err("the int32 DP matrix of a wide alignment needs more scratch than the device memory budget");
            return VGC_ERR_CAPACITY;
          }
          wide_n = 1;
        }
        if ((rc = h->d_wide.reserve(wide_buf * wide_n))) return rc;
        if ((rc = h->d_wide_free.reserve(4ull * (wide_n + 2)))) return rc;
        std::vector<uint32_t> init(wide_n + 2);
        for (uint32_t i = 0; i < wide_n; ++i) init[i] = i;
        init[wide_n] = 0;
        init[wide_n + 1] = wide_n;
        VGC_CUDA(cudaMemcpy(h->d_wide_free.p, init.data(), 4ull * (wide_n + 2), cudaMemcpyHostToDevice));
        pool_bytes += wide_buf * wide_n;
      }
    }
    lap(2);
    const uint32_t pool_n = per_sm * static_cast<uint32_t>(h->sm_count);
    if ((rc = h->d_pool_busy.reserve(4ull * (pool_n + 2)))) return rc;
    std::vector<uint32_t> pool_init(pool_n + 2);
    for (uint32_t i = 0; i < pool_n; ++i) pool_init[i] = i;
    pool_init[pool_n] = 0;           // take tickets
    pool_init[pool_n + 1] = pool_n;  // give tickets: the initial ids count as given back
    VGC_CUDA(cudaMemcpyAsync(h->d_pool_busy.p, pool_init.data(), 4ull * (pool_n + 2), cudaMemcpyHostToDevice, h->stream));

    // slot geometry of every remaining window (host threads), then the chunk = the prefix the budget holds
    const size_t n_rest = wins.size() - pos;
    std::vector<SlotDims> dims(n_rest);
    std::vector<uint64_t> offs(n_rest);
    parallel_ranges(n_rest, kSetupThreads, [&](size_t a, size_t b) {
      for (size_t x = a; x < b; ++x) {
        const uint32_t w = wins[pos + x];
        const uint32_t f = win_first[w];
        SlotDims d;
        d.max_nodes = std::max<uint32_t>(estimate_nodes(pr, w, static_cast<uint32_t>(seq_off[f + 1] - seq_off[f]), exact, h->node_share_div), 64);
        d.max_edges = 2 * d.max_nodes + 64;  // ~2 edges per node in practice; AddAlignment wants room for a whole layer
        d.max_len = ml;
        d.row_words = 32 * K;
        // in-degree <= number of sequences; 16 is ample in practice, the exact pass takes the bound itself
        d.al_stride = pr.num_codes > 8 ? 16 : 8;
        d.in_stride = exact ? std::max<uint32_t>(8, pr.win_nseq[w] + 1) : std::min<uint32_t>(16, std::max<uint32_t>(8, pr.win_nseq[w] + 1));
        dims[x] = d;
        offs[x] = slot_bytes(d);  // size for now, offset below
      }
    });
    uint64_t bytes = 0;
    const uint64_t slot_budget = h->mem_budget - pool_bytes;
    size_t e = pos;
    while (e < wins.size()) {
      const uint64_t sb = offs[e - pos];
      if (bytes + sb > slot_budget && e > pos) break;
      if (sb > slot_budget) {
        set_err("a window needs more scratch than the device memory budget");
        return VGC_ERR_CAPACITY;
      }
      offs[e - pos] = bytes;
      bytes += sb;
      ++e;
    }
    lap(3);
    const uint32_t n = static_cast<uint32_t>(e - pos);
    if ((rc = h->d_slot_mem.reserve(bytes))) return rc;
    if ((rc = h->d_slots.reserve(sizeof(Slot) * n))) return rc;
    if ((rc = h->d_wstates.reserve(sizeof(WinState) * n))) return rc;
    if ((rc = h->d_work.reserve(4ull * n))) return rc;
    // ---- deal the chunk's windows to groups: one group per distinct number of alignments (windows of a group then
    // run the very same program, so the heavy serial steps — PruneGraph + LargestSubgraph — coincide instead of
    // stalling some cycle of every group); sparse values at the tails are merged until a group has >= 64 windows.
    // VGC_GROUP_MODE=1: equal contiguous blocks of the sorted list.
    lap(4);
    std::vector<uint32_t> gstart;  // positions in the chunk (sorted by decreasing cycles) where a group starts
    if (h->group_mode == 2) {
      uint32_t min_group = 64;
      while (true) {
        gstart.assign(1, 0);
        for (uint32_t i = 1; i < n; ++i) {
          if (pr.win_nfill[wins[pos + i]] != pr.win_nfill[wins[pos + i - 1]] && i - gstart.back() >= min_group)
            gstart.push_back(i);
        }
        if (gstart.size() <= static_cast<size_t>(h->groups)) break;
        min_group *= 2;
      }
    } else {
      const int G0 = std::max(1, std::min<int>(h->groups, static_cast<int>((n + 255) / 256)));
      for (int g = 0; g < G0; ++g) gstart.push_back(static_cast<uint32_t>(static_cast<uint64_t>(n) * g / G0));
    }
    const int G = static_cast<int>(gstart.size());
    std::vector<uint32_t> work(n);
    std::vector<Slot> slots(n);
    std::vector<uint32_t> gbase(G + 1, 0);
    std::vector<std::vector<uint32_t>> gnseq(G);  // sequences of each window of the group, in list order (decreasing)
    std::vector<uint32_t> gmin_nseq(G, 0xFFFFFFFFu);
    std::vector<uint64_t> gjobs_cap(G, 0);
    std::vector<double> gblen(G, 0.0), gavglen(G, 0.0);  // per group: longest backbone, largest mean layer length
    uint32_t k = 0;
    uint32_t max_cyc = 0;
    for (int g = 0; g < G; ++g) {
      gbase[g] = k;
      const uint32_t b0 = gstart[g];
      const uint32_t b1 = g + 1 < G ? gstart[g + 1] : n;
      gnseq[g].reserve(b1 - b0);
      for (uint32_t i = b0; i < b1; ++i) {
        const uint32_t w = wins[pos + i];  // (k == i: the groups are consecutive ranges of the chunk)
        gnseq[g].push_back(pr.win_nseq[w]);
        gjobs_cap[g] += pr.win_nseq[w];
        max_cyc = std::max(max_cyc, win_cycles(pr.win_nseq[w], hap, num_prune));
        const uint32_t f = win_first[w];
        const double bl = static_cast<double>(seq_off[f + 1] - seq_off[f]);
        gmin_nseq[g] = std::min(gmin_nseq[g], pr.win_nseq[w]);
        gblen[g] = std::max(gblen[g], bl);
        gavglen[g] = std::max(gavglen[g], (pr.win_sum_len[w] - bl) / std::max(1.0, pr.win_nseq[w] - 1.0));
        ++k;
      }
    }
    gbase[G] = k;
    {
      uint8_t* const slot_base = h->d_slot_mem.as<uint8_t>();
      parallel_ranges(n, kSetupThreads, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
          work[i] = wins[pos + i];
          slot_carve(dims[i], slot_base + offs[i], &slots[i]);
        }
      });
    }
    lap(5);
    // job lists (one per group, reused every cycle) and their per-cycle counters
    std::vector<uint64_t> gjob_off(G + 1, 0);
    for (int g = 0; g < G; ++g) gjob_off[g + 1] = gjob_off[g] + gjobs_cap[g];
    if ((rc = h->d_jobs.reserve(std::max<uint64_t>(gjob_off[G], 1) * kClasses * sizeof(Job)))) return rc;
    const size_t ncnt = static_cast<size_t>(G) * (max_cyc + 1) * kClasses;
    if ((rc = h->d_jobcnt.reserve(4ull * ncnt))) return rc;
    VGC_CUDA(cudaMemsetAsync(h->d_jobcnt.p, 0, 4ull * ncnt, h->stream));
    VGC_CUDA(cudaMemcpyAsync(h->d_work.p, work.data(), 4ull * n, cudaMemcpyHostToDevice, h->stream));
    VGC_CUDA(cudaMemcpyAsync(h->d_slots.p, slots.data(), sizeof(Slot) * n, cudaMemcpyHostToDevice, h->stream));
    VGC_CUDA(cudaMemsetAsync(h->d_wstates.p, 0, sizeof(WinState) * n, h->stream));
    KernelArgs a;
    a.bv = make_view(h);
    a.work = h->d_work.as<uint32_t>();
    a.slots = h->d_slots.as<Slot>();
    a.wstates = h->d_wstates.as<WinState>();
    a.out = h->d_out.as<uint8_t>();
    a.out_len = h->d_out_len.as<uint32_t>();
    a.status = h->d_status.as<uint32_t>();
    a.totals = reinterpret_cast<unsigned long long*>(h->d_misc.as<uint8_t>() + 16);
    a.nw.m = h->params.match;
    a.nw.x = h->params.mismatch;
    a.nw.g = h->params.gap;
    a.haplotype = h->params.haplotype;
    a.trim = h->params.trim;
    a.num_prune = h->params.num_prune;
    a.min_confidence = h->params.min_confidence;
    a.min_support = h->params.min_support;
    a.smem_bytes = 0;
    a.pool = h->d_pool.as<uint8_t>();
    a.pool_free = h->d_pool_busy.as<uint32_t>();
    a.pool_buf_bytes = buf_bytes;
    a.pool_fc_off = pool_fc_off;
    a.pool_rows = pool_rows;
    a.pool_n = pool_n;
    a.wide_pool = h->d_wide.as<uint8_t>();
    a.wide_free = h->d_wide_free.as<uint32_t>();
    a.wide_buf_bytes = wide_buf;
    a.wide_n = wide_n;
    a.tl = nullptr;
    a.tl_cap = 0;
    const char* tl_path = std::getenv("VGC_TIMELINE");
    if (tl_path) {
      a.tl_cap = 8u << 20;
      if ((rc = h->d_tl.reserve(16ull * (a.tl_cap + 1)))) return rc;
      a.tl = h->d_tl.as<uint4>();
      VGC_CUDA(cudaMemsetAsync(a.tl, 0, 16, h->stream));
    }
    VGC_CUDA(cudaEventRecord(h->ev[6], h->stream));
    lap(6);
    if (std::getenv("VGC_VERBOSE"))
      std::fprintf(stderr, "[vgc] pass setup: attrs %.2f, pool rows %.2f, pool reserve + wide %.2f, dims %.2f, reserves %.2f, "
                   "groups + carve %.2f, copies %.2f ms\n", laps[0], laps[1], laps[2], laps[3], laps[4], laps[5], laps[6]);
    h->setup_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count();
    for (int g = 0; g < G; ++g) VGC_CUDA(cudaStreamWaitEvent(h->gstream[g], h->ev[6], 0));
    // ---- lockstep: the lists are sorted by decreasing cycles, so the live windows of a cycle are a prefix
    std::vector<uint32_t> live(G);
    for (int g = 0; g < G; ++g) live[g] = static_cast<uint32_t>(gnseq[g].size());
    const auto tl0 = std::chrono::steady_clock::now();
    uint32_t nl = 0;
    for (uint32_t c = 0; c < max_cyc; ++c) {
      if (c == 3) h->in_setup.store(false, std::memory_order_release);  // the device has work queued: staging may start
      for (int g = 0; g < G; ++g) {
        const std::vector<uint32_t>& ns = gnseq[g];
        while (live[g] > 0 && win_cycles(ns[live[g] - 1], hap, num_prune) <= c) --live[g];
        const uint32_t nlive = live[g];
        if (!nlive) continue;
        cudaStream_t st = h->gstream[g];
        KernelArgs ka = a;
        ka.smem_bytes = smem_update;  // per window (warp)
        Job* jobs = h->d_jobs.as<Job>() + gjob_off[g] * kClasses;
        const uint32_t job_cap = static_cast<uint32_t>(gjobs_cap[g]);
        uint32_t* cnt = h->d_jobcnt.as<uint32_t>() + (static_cast<size_t>(g) * (max_cyc + 1) + c) * kClasses;
        update_kernel<K><<<(nlive + kUpdateWins - 1) / kUpdateWins, 32 * kUpdateWins, kUpdateWins * ka.smem_bytes, st>>>(ka, gbase[g], nlive, jobs, job_cap, cnt);
        ++nl;
        uint32_t ncls[kFastClasses] = {0, 0, 0, 0, 0, 0};  // alignments this cycle hands to each class kernel (exact)
        uint32_t nprep = 0;  // windows that still have a prepare step in this cycle: all but those that just emitted
        for (uint32_t i = 0; i < nlive; ++i) {
          const uint32_t nj = win_jobs(ns[i], hap, num_prune, c);
          if (nj) {
            const uint32_t w = work[gbase[g] + i];
            const uint32_t f = win_first[w];
            if (nj > 1) {
              for (uint32_t j = 0; j < ns[i]; ++j) {
                const uint32_t l = pr.layer_rank[f + j];
                ncls[job_class(static_cast<uint32_t>(seq_off[l + 1] - seq_off[l]), pr.layer_sw[l] != 0)] += 1;
              }
            } else if (c + 1 < ns[i]) {  // build alignment: layer c + 1 in rank order, global
              const uint32_t l = pr.layer_rank[f + c + 1];
              ncls[job_class(static_cast<uint32_t>(seq_off[l + 1] - seq_off[l]), false)] += 1;
            } else {  // final local alignment of the backbone
              const uint32_t l = pr.layer_rank[f];
              ncls[job_class(static_cast<uint32_t>(seq_off[l + 1] - seq_off[l]), true)] += 1;
            }
          }
          if (win_cycles(ns[i], hap, num_prune) > c + 1) nprep = i + 1;
        }
        if (!nprep) continue;
        // shared memory of the sort kernel: sized for the graph this cycle can have reached (build phase: the
        // backbone + a share of the bases added so far), so early cycles run more CTAs per SM; a graph that
        // outgrows it sorts out of HBM instead (slower, same result)
        uint32_t ss = smem_sort_cap;
        if (c + 1 < gmin_nseq[g]) {
          const double nvb = gblen[g] + h->sort_growth * c * gavglen[g] + 64.0;
          const double need = kSmemHeader + ((ml + 15u) & ~15u) + 9.2 * nvb + 1024.0;
          ss = std::min<uint32_t>(smem_sort_cap, std::max<uint32_t>(4096u, (static_cast<uint32_t>(need) + 255u) & ~255u));
        }
        ka.smem_bytes = ss;
        sort_kernel<K><<<nprep, 32, ss, st>>>(ka, gbase[g], jobs, job_cap, cnt);
        ++nl;
        uint32_t njobs_cycle = 0;
        for (int cls = 0; cls < kFastClasses; ++cls) {
          njobs_cycle += ncls[cls];
          if (!ncls[cls] || !smem_align[cls]) continue;
          ka.smem_bytes = smem_align[cls];
          align_fn(cls)<<<ncls[cls], 32, smem_align[cls], st>>>(ka, jobs + static_cast<size_t>(cls) * job_cap, cnt + cls);
          ++nl;
        }
        if (wide_n && njobs_cycle) {
          ka.smem_bytes = smem_wide;
          align_wide_kernel<<<std::min(njobs_cycle, wide_n), 32, smem_wide, st>>>(ka, jobs + static_cast<size_t>(kClassWide) * job_cap, cnt + kClassWide);
          ++nl;
        }
      }
    }
    *launches += nl;
    h->launch_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl0).count();
    VGC_CUDA(cudaGetLastError());
    for (int g = 0; g < G; ++g) {
      VGC_CUDA(cudaEventRecord(h->gev[g], h->gstream[g]));
      VGC_CUDA(cudaStreamWaitEvent(h->stream, h->gev[g], 0));
    }
    VGC_CUDA(cudaEventRecord(h->ev[7], h->stream));
    VGC_CUDA(cudaStreamSynchronize(h->stream));  // host vectors must outlive their copies; surfaces kernel faults
    float kms = 0.f;
    VGC_CUDA(cudaEventElapsedTime(&kms, h->ev[6], h->ev[7]));
    h->pass_kernel_ms += kms;
    if (tl_path) {
      uint4 head;
      VGC_CUDA(cudaMemcpy(&head, a.tl, 16, cudaMemcpyDeviceToHost));
      const uint32_t nrec = std::min(head.x, a.tl_cap);
      std::vector<uint4> recs(nrec);
      VGC_CUDA(cudaMemcpy(recs.data(), a.tl + 1, 16ull * nrec, cudaMemcpyDeviceToHost));
      if (FILE* f = std::fopen(tl_path, "wb")) {  // the last pass wins
        std::fwrite(recs.data(), 16, nrec, f);
        std::fclose(f);
      }
    }
    pos = e;
  }
  return VGC_OK;
}

int run_pass(vgc_engine* h, const std::vector<uint32_t>& wins, bool exact, int K, const uint64_t* seq_off,
             const uint32_t* win_first, uint32_t* launches) {
  if (K == 10) return run_pass_k<10>(h, wins, exact, seq_off, win_first, launches);
  return run_pass_k<16>(h, wins, exact, seq_off, win_first, launches);
}

int polish_device(vgc_engine* h, vgc_result* result, vgc_stats* stats, uint64_t input_bytes,
                  const uint8_t* host_bases, const uint64_t* seq_off, const uint32_t* win_first, uint32_t nw) {
  Prepared& pr = h->I->prep;
  int rc;
  if ((rc = h->d_out.reserve(std::max<uint64_t>(pr.out_total, 16)))) return rc;
  if ((rc = h->d_out_len.reserve(std::max<size_t>(nw, 4) * 4))) return rc;
  if ((rc = h->d_status.reserve(std::max<size_t>(nw, 4) * 4))) return rc;
  if ((rc = h->d_misc.reserve(256))) return rc;
  if (h->h_win_cap < nw) {
    if (h->h_out_len) cudaFreeHost(h->h_out_len);
    if (h->h_status) cudaFreeHost(h->h_status);
    h->h_win_cap = nw + nw / 4 + 16;
    VGC_CUDA(cudaMallocHost(&h->h_out_len, h->h_win_cap * 4));
    VGC_CUDA(cudaMallocHost(&h->h_status, h->h_win_cap * 4));
  }
  if (h->h_out_cap < pr.out_total) {
    if (h->h_out) cudaFreeHost(h->h_out);
    h->h_out_cap = pr.out_total + pr.out_total / 4 + 256;
    VGC_CUDA(cudaMallocHost(&h->h_out, h->h_out_cap));
  }
  const uint32_t n_dev = static_cast<uint32_t>(pr.device_windows.size());
  uint32_t launches = 0, relaunched = 0;
  unsigned long long totals[8 + vgc::kPhCount] = {0};
  float kernel_ms = 0.f, d2h_ms = 0.f;
  h->pass_kernel_ms = 0.0;
  h->launch_ms = 0.0;
  h->setup_ms = 0.0;
  VGC_CUDA(cudaMemsetAsync(h->d_misc.p, 0, 256, h->stream));
  VGC_CUDA(cudaMemsetAsync(h->d_out_len.p, 0, std::max<size_t>(nw, 4) * 4, h->stream));
  VGC_CUDA(cudaMemsetAsync(h->d_status.p, 0, std::max<size_t>(nw, 4) * 4, h->stream));
  if (n_dev) {
    const int K = pr.max_len <= 640 ? 10 : 16;
    if (pr.max_len > VGC_MAX_LAYER_LEN) {
      set_err("a layer is longer than 16383 bases (engine limit, see vgc_limits): use a window length (racon -w) "
              "below ~8000");
      return VGC_ERR_CAPACITY;
    }
    VGC_CUDA(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = run_pass(h, pr.device_windows, false, K, seq_off, win_first, &launches))) return rc;
    VGC_CUDA(cudaMemcpyAsync(h->h_status, h->d_status.p, nw * 4ull, cudaMemcpyDeviceToHost, h->stream));
    VGC_CUDA(cudaStreamSynchronize(h->stream));
    // second pass, exact capacities, for windows whose graph outgrew the estimate
    std::vector<uint32_t> retry;
    for (uint32_t w : pr.device_windows) {
      if (h->h_status[w] == kStNodeOverflow || h->h_status[w] == kStEdgeOverflow || h->h_status[w] == kStDegreeOverflow ||
          h->h_status[w] == kStWideCapacity)
        retry.push_back(w);
    }
    if (!retry.empty()) {
      relaunched = static_cast<uint32_t>(retry.size());
      if ((rc = run_pass(h, retry, true, K, seq_off, win_first, &launches))) return rc;
    }
    VGC_CUDA(cudaEventRecord(h->ev[1], h->stream));
  }
  NvtxRange nvtx_out("vgc: D2H + stitch");
  const auto t_out0 = std::chrono::steady_clock::now();
  VGC_CUDA(cudaEventRecord(h->ev[2], h->stream));
  VGC_CUDA(cudaMemcpyAsync(h->h_status, h->d_status.p, nw * 4ull, cudaMemcpyDeviceToHost, h->stream));
  VGC_CUDA(cudaMemcpyAsync(h->h_out_len, h->d_out_len.p, nw * 4ull, cudaMemcpyDeviceToHost, h->stream));
  if (result && pr.out_total) {
    VGC_CUDA(cudaMemcpyAsync(h->h_out, h->d_out.p, pr.out_total, cudaMemcpyDeviceToHost, h->stream));
  }
  VGC_CUDA(cudaMemcpyAsync(totals, h->d_misc.as<uint8_t>() + 16, sizeof(totals), cudaMemcpyDeviceToHost, h->stream));
  VGC_CUDA(cudaEventRecord(h->ev[3], h->stream));
  VGC_CUDA(cudaStreamSynchronize(h->stream));
  kernel_ms = static_cast<float>(h->pass_kernel_ms);
  VGC_CUDA(cudaEventElapsedTime(&d2h_ms, h->ev[2], h->ev[3]));
  // status check: anything but OK is an engine limit (there is no CPU fallback)
  for (uint32_t w : pr.device_windows) {
    if (h->h_status[w] != kStOk) {
      char msg[160];
      std::snprintf(msg, sizeof(msg), "window %u failed on the device with status %u (see poa_core.h kSt*)", w,
                    h->h_status[w]);
      set_err(msg);
      return h->h_status[w] == kStInternal ? VGC_ERR_CUDA : VGC_ERR_CAPACITY;
    }
  }
  const auto t_out1 = std::chrono::steady_clock::now();
  if (result) {
    // stitch: windows in order; < 3 sequences -> backbone, polished = false (window.cpp:188-192)
    uint64_t off = 0;
    for (uint32_t w = 0; w < nw; ++w) {
      result->cons_off[w] = off;
      uint32_t n;
      const uint8_t* src;
      if (pr.win_nseq[w] < 3) {
        const uint32_t f = win_first[w];
        n = static_cast<uint32_t>(seq_off[f + 1] - seq_off[f]);
        src = host_bases + seq_off[f];
        result->polished[w] = 0;
      } else {
        n = h->h_out_len[w];
        src = h->h_out + pr.out_off[w];
        result->polished[w] = 1;
      }
      if (off + n > result->cons_capacity) {
        set_err("vgc_result.cons too small (use vgc_result_bound)");
        return VGC_ERR_INVALID;
      }
      std::memcpy(result->cons + off, src, n);
      off += n;
    }
    result->cons_off[nw] = off;
  }
  if (std::getenv("VGC_VERBOSE")) {
    const auto t_out2 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[vgc] after the pass: D2H + status check %.2f ms, stitch copy %.2f ms\n",
                 std::chrono::duration<double, std::milli>(t_out1 - t_out0).count(),
                 std::chrono::duration<double, std::milli>(t_out2 - t_out1).count());
  }
  // feedback for the next call: if more than 2 % of the sorts did not fit the shared memory the growth bound gave
  // them, the data grows its graphs faster than assumed — raise the bound (sticky per engine)
  {
    const double sorts = static_cast<double>(totals[2 + vgc::kPhCount]);
    const double hbm = static_cast<double>(totals[3 + vgc::kPhCount]);
    if (sorts > 0 && hbm > 0.02 * sorts && h->sort_growth < 1.0) h->sort_growth *= 1.3;
  }
  if (stats) {
    for (int i = 0; i < vgc::kPhCount; ++i) h->phase_cycles[i] = static_cast<double>(totals[2 + i]);
    h->phase_cycles[11] = h->launch_ms;
    h->phase_cycles[12] = static_cast<double>(totals[2 + vgc::kPhCount]);
    h->phase_cycles[13] = static_cast<double>(totals[3 + vgc::kPhCount]);
    h->phase_cycles[14] = static_cast<double>(totals[4 + vgc::kPhCount]);
    h->phase_cycles[15] = static_cast<double>(totals[5 + vgc::kPhCount]);
    h->phase_cycles[0] = static_cast<double>(totals[6 + vgc::kPhCount]);
    h->pool_spins = static_cast<double>(totals[7 + vgc::kPhCount]);
    if (std::getenv("VGC_VERBOSE"))
      std::fprintf(stderr, "[vgc] pass: %.1f ms kernels, %.1f ms host setup before the first launch, %.1f ms enqueueing, %u launches, pool waits %.0f, slow trace steps %.0f, max fill %.0f / trace %.0f cycles\n",
                   h->pass_kernel_ms, h->setup_ms, h->launch_ms, launches, h->pool_spins, h->phase_cycles[0], h->phase_cycles[14], h->phase_cycles[15]);
    stats->cells = totals[0];
    stats->alignments = totals[1];
    stats->input_bytes = input_bytes;
    stats->output_bytes = (result ? pr.out_total : 0) + nw * 8ull + 16;
    stats->kernel_ms = kernel_ms;
    stats->d2h_ms = d2h_ms;
    float dev_ms = 0.f;
    VGC_CUDA(cudaEventElapsedTime(&dev_ms, h->ev[n_dev ? 0 : 2], h->ev[3]));
    stats->device_ms = dev_ms;
    stats->host_prep_ms = 0.0;
    stats->kernel_launches = launches;
    stats->relaunched_windows = relaunched;
  }
  return VGC_OK;
}

}  // namespace

static int engine_init(vgc_engine* h);

extern "C" {

const char* vgc_last_error(void) { return g_err.c_str(); }
int vgc_phase_profile(vgc_handle h, double out[16]) {
  if (!h || !out) return VGC_ERR_INVALID;
  for (int i = 0; i < 16; ++i) out[i] = h->phase_cycles[i];
  return VGC_OK;
}

const char* vgc_version(void) { return "vechat_b200 0.1 (sm_100a)"; }

void vgc_weight_lut(uint32_t lut[256]) { vgc::weight_lut(lut); }

void vgc_limits(vgc_limits_t* out) {
  if (!out) return;
  std::memset(out, 0, sizeof(*out));
  out->max_layer_len = VGC_MAX_LAYER_LEN;
  out->max_backbone_len = VGC_MAX_BACKBONE;
  out->max_codes = VGC_MAX_CODES;
  out->fast_layer_len = VGC_FAST_LAYER_LEN;
  out->int16_score_bound = 32000;
}

int vgc_window_status(vgc_handle h, uint32_t* status, uint32_t n) {
  if (!h || !status) return VGC_ERR_INVALID;
  for (uint32_t w = 0; w < n; ++w) status[w] = (h->h_status && w < h->h_win_cap) ? h->h_status[w] : 0u;
  return VGC_OK;
}

int vgc_create(vgc_handle* out, int device, const vgc_params* params) {
  if (!out || !params) {
    set_err("null argument");
    return VGC_ERR_INVALID;
  }
  *out = nullptr;
  if (params->gap > 0) {  // vendor/spoa/src/alignment_engine.cpp:37-51 (linear gaps: g == e == q == c)
    set_err("[spoa::AlignmentEngine::Create] error: gap opening penalty must be non-positive!");
    return VGC_ERR_INVALID;
  }
  if (params->haplotype && params->num_prune == 0) {
    set_err("invalid params: num_prune must be >= 1 in haplotype mode");
    return VGC_ERR_INVALID;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
    set_err(std::string("no usable CUDA device (the engine has no CPU fallback): ") +
            (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range"));
    cudaGetLastError();
    return VGC_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  VGC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_err("device is not sm_100-class (Blackwell); this library carries sm_100a code only");
    return VGC_ERR_NO_DEVICE;
  }
  VGC_CUDA(cudaSetDevice(device));
  auto* h = new vgc_engine();
  h->device = device;
  h->params = *params;
  h->sm_count = prop.multiProcessorCount;
  const int rc_init = engine_init(h);
  if (rc_init != VGC_OK) {
    vgc_destroy(h);  // frees whatever was created
    return rc_init;
  }
  *out = h;
  return VGC_OK;
}

}  // extern "C"

static int engine_init(vgc_engine* h) {
  VGC_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  VGC_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (auto& ev : h->ev) VGC_CUDA(cudaEventCreate(&ev));
  for (InputSet& in : h->in) {
    VGC_CUDA(cudaEventCreate(&in.ev0));
    VGC_CUDA(cudaEventCreate(&in.ev1));
  }
  // groups are ordered deepest windows first: their chain of cycles is the critical path of a pass, so their
  // streams get the higher priorities
  int prio_least = 0, prio_greatest = 0;
  VGC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  for (int g = 0; g < kMaxGroups; ++g) {
    const int prio = std::min(prio_least, prio_greatest + g / 4);
    VGC_CUDA(cudaStreamCreateWithPriority(&h->gstream[g], cudaStreamNonBlocking, prio));
    VGC_CUDA(cudaEventCreateWithFlags(&h->gev[g], cudaEventDisableTiming));
  }
  // shared memory per one-warp CTA of each kernel: what its CTAs-per-SM target leaves (1 KB reserved per CTA)
  auto smem_for = [](int ctas) { return static_cast<uint32_t>(((228 * 1024 - ctas * 1024) / ctas) & ~255); };
  h->smem_sort = smem_for(VGC_SORT_CTAS);
  h->smem_update = 6144;  // header + codes + the staged dirty blocks of the incremental order (poa_core.h order_update)
  if (const char* s = std::getenv("VGC_UPDATE_SMEM")) h->smem_update = static_cast<uint32_t>(std::atoi(s));
  if (const char* s = std::getenv("VGC_SORT_SMEM")) h->smem_sort = static_cast<uint32_t>(std::atoi(s));
  if (const char* s = std::getenv("VGC_GROUP_MODE")) h->group_mode = std::max(0, std::min(2, std::atoi(s)));
  if (const char* s = std::getenv("VGC_NODE_SHARE_DIV")) h->node_share_div = static_cast<uint32_t>(std::max(1, std::atoi(s)));
  if (const char* s = std::getenv("VGC_SORT_GROWTH")) h->sort_growth = std::atof(s);
  if (const char* s = std::getenv("VGC_LAUNCH_THREADS")) h->launch_threads = std::max(1, std::min(16, std::atoi(s)));
  if (const char* s = std::getenv("VGC_GROUPS")) h->groups = std::max(1, std::min(kMaxGroups, std::atoi(s)));
  {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency()), procs = 1;
    if (const char* s = std::getenv("LOCAL_WORLD_SIZE")) procs = static_cast<unsigned>(std::max(1, std::atoi(s)));
    h->host_threads = std::max(1u, std::min(16u, hw / procs));
    if (const char* s = std::getenv("VGC_HOST_THREADS")) h->host_threads = static_cast<unsigned>(std::max(1, std::min(64, std::atoi(s))));
  }
  size_t free_b = 0, total_b = 0;
  VGC_CUDA(cudaMemGetInfo(&free_b, &total_b));
  h->mem_budget = static_cast<size_t>(free_b * 0.70);
  if (const char* s = std::getenv("VGC_MEM_BUDGET_MB")) {
    h->mem_budget = static_cast<size_t>(std::atoll(s)) << 20;
    h->mem_budget_fixed = true;
  }
  return VGC_OK;
}

extern "C" {

// wait for a staging in flight on `in` (its worker thread), if any
static void join_set(InputSet* in) {
  if (in->worker.joinable()) in->worker.join();
}

int vgc_destroy(vgc_handle h) {
  if (!h) return VGC_OK;
  cudaSetDevice(h->device);
  for (InputSet& in : h->in) {
    join_set(&in);
    for (DevBuf* d : {&in.d_bases, &in.d_quals, &in.d_seq_off, &in.d_has_qual, &in.d_begin, &in.d_end, &in.d_win_first,
                      &in.d_win_flags, &in.d_rank, &in.d_nseq, &in.d_avgw, &in.d_out_off, &in.d_out_cap, &in.d_tables})
      d->release();
    if (in.h_packed) cudaFreeHost(in.h_packed);
    if (in.ev0) cudaEventDestroy(in.ev0);
    if (in.ev1) cudaEventDestroy(in.ev1);
  }
  for (DevBuf* d : {&h->d_work, &h->d_out, &h->d_out_len, &h->d_status, &h->d_misc, &h->d_slots, &h->d_slot_mem,
                    &h->d_wstates, &h->d_pool, &h->d_pool_busy, &h->d_jobs, &h->d_tl, &h->d_wide, &h->d_wide_free,
                    &h->d_jobcnt})
    d->release();
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->h_out_len) cudaFreeHost(h->h_out_len);
  if (h->h_status) cudaFreeHost(h->h_status);
  for (auto& ev : h->ev) {
    if (ev) cudaEventDestroy(ev);
  }
  for (int g = 0; g < 64; ++g) {
    if (h->gstream[g]) cudaStreamDestroy(h->gstream[g]);
    if (h->gev[g]) cudaEventDestroy(h->gev[g]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return VGC_OK;
}

uint64_t vgc_result_bound(const vgc_batch* batch) {
  if (!batch || !batch->seq_off) return 0;
  return batch->seq_off[batch->n_layers] + 16;
}

int vgc_submit(vgc_handle h, const vgc_batch* batch) {
  if (!h || !batch) {
    set_err("null argument");
    return VGC_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(h->qmu);
  if (h->n_queued >= 2) {
    set_err("two batches are already staged: call vgc_collect first");
    return VGC_ERR_INVALID;
  }
  // a set is queued, or being read by the running pass (h->I while vgc_collect is inside polish_device), or free
  int idx = -1;
  for (int i = 0; i < 2 && idx < 0; ++i) {
    bool busy = h->running && &h->in[i] == h->I;
    for (int q = 0; q < h->n_queued; ++q) busy = busy || h->queued[q] == i;
    if (!busy) idx = i;
  }
  if (idx < 0) {
    set_err("no free input set: call vgc_collect first");
    return VGC_ERR_INVALID;
  }
  InputSet* in = &h->in[idx];
  join_set(in);
  h->resident = h->resident && in != h->I;  // staging over the resident batch's set ends its residency
  in->batch = *batch;
  in->rc = VGC_OK;
  in->err.clear();
  h->queued[h->n_queued++] = idx;
  in->worker = std::thread([h, in] {
    // the serial host work at the start of a pass (slots, groups, first launches) is on the device's critical path:
    // let it finish before two dozen staging threads compete with it for the cores
    // ... and it runs at a lower priority than the thread that enqueues the kernel launches of the running pass (the
    // threads prepare_batch and the packer start inherit it)
    setpriority(PRIO_PROCESS, static_cast<id_t>(syscall(SYS_gettid)), 10);
    for (int spin = 0; spin < 400 && h->in_setup.load(std::memory_order_acquire); ++spin)
      std::this_thread::sleep_for(std::chrono::microseconds(500));
    in->rc = stage_batch(h, in, &in->batch);
    if (in->rc != VGC_OK) in->err = vgc_last_error();  // thread-local: carry it to the collecting thread
  });
  return VGC_OK;
}

int vgc_collect(vgc_handle h, vgc_result* result, vgc_stats* stats) {
  if (!h || !result) {
    set_err("null argument");
    return VGC_ERR_INVALID;
  }
  VGC_CUDA(cudaSetDevice(h->device));
  InputSet* in;
  {
    std::lock_guard<std::mutex> lock(h->qmu);
    if (h->n_queued == 0) {
      set_err("nothing submitted: call vgc_submit first");
      return VGC_ERR_INVALID;
    }
    in = &h->in[h->queued[0]];
    h->queued[0] = h->queued[1];
    --h->n_queued;
    h->I = in;  // from here on the set counts as busy for vgc_submit
    h->running = true;
    h->in_setup.store(true, std::memory_order_release);
  }
  const auto tc0 = std::chrono::steady_clock::now();
  join_set(in);
  const auto tc1 = std::chrono::steady_clock::now();
  if (in->rc != VGC_OK) {
    set_err(in->err);
    std::lock_guard<std::mutex> lock(h->qmu);
    h->running = false;
    h->in_setup.store(false, std::memory_order_release);
    return in->rc;
  }
  h->resident = false;
  const vgc_batch* b = &in->batch;
  const int rc = polish_device(h, result, stats, in->in_bytes, b->bases, b->seq_off, b->win_first, b->n_windows);
  if (std::getenv("VGC_VERBOSE")) {
    const auto tc2 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b2) {
      return std::chrono::duration<double, std::milli>(b2 - a).count();
    };
    std::fprintf(stderr, "[vgc] collect: waited %.1f ms for the staging, pass + D2H + stitch %.1f ms (kernels %.1f, setup %.1f, "
                 "enqueue %.1f); staging: prepare %.1f ms, pack %.1f ms\n",
                 ms(tc0, tc1), ms(tc1, tc2), h->pass_kernel_ms, h->setup_ms, h->launch_ms, in->prep_ms, in->pack_ms);
  }
  {
    std::lock_guard<std::mutex> lock(h->qmu);
    h->running = false;
    h->in_setup.store(false, std::memory_order_release);
  }
  if (rc != VGC_OK) return rc;
  if (stats) {
    float h2d_ms = 0.f;
    VGC_CUDA(cudaEventElapsedTime(&h2d_ms, in->ev0, in->ev1));
    stats->h2d_ms = h2d_ms;
    stats->host_prep_ms = in->prep_ms;
    stats->host_pack_ms = in->pack_ms;
  }
  return VGC_OK;
}

int vgc_polish(vgc_handle h, const vgc_batch* batch, vgc_result* result, vgc_stats* stats) {
  if (!h || !batch || !result) {
    set_err("null argument");
    return VGC_ERR_INVALID;
  }
  if (h->n_queued) {
    set_err("vgc_polish with batches still queued: collect them first");
    return VGC_ERR_INVALID;
  }
  VGC_CUDA(cudaSetDevice(h->device));
  // the same two steps as vgc_submit + vgc_collect, on this thread
  InputSet* in = h->I;
  join_set(in);
  h->resident = false;
  int rc = stage_batch(h, in, batch);
  if (rc != VGC_OK) return rc;
  rc = polish_device(h, result, stats, in->in_bytes, batch->bases, batch->seq_off, batch->win_first, batch->n_windows);
  if (rc != VGC_OK) return rc;
  if (stats) {
    float h2d_ms = 0.f;
    VGC_CUDA(cudaEventElapsedTime(&h2d_ms, in->ev0, in->ev1));
    stats->h2d_ms = h2d_ms;
    stats->device_ms += h2d_ms;
    stats->host_prep_ms = in->prep_ms;
    stats->host_pack_ms = in->pack_ms;
  }
  return VGC_OK;
}

int vgc_upload(vgc_handle h, const vgc_batch* batch) {
  if (!h || !batch) {
    set_err("null argument");
    return VGC_ERR_INVALID;
  }
  if (h->n_queued) {
    set_err("vgc_upload with batches still queued: collect them first");
    return VGC_ERR_INVALID;
  }
  VGC_CUDA(cudaSetDevice(h->device));
  h->resident = false;
  InputSet* in = h->I;
  join_set(in);
  int rc = stage_batch(h, in, batch);
  if (rc != VGC_OK) return rc;
  h->r_input_bytes = in->in_bytes;
  // keep what the stitcher needs from the host batch
  h->r_n_windows = batch->n_windows;
  h->r_n_layers = batch->n_layers;
  h->r_seq_off.assign(batch->seq_off, batch->seq_off + batch->n_layers + 1);
  h->r_win_first.assign(batch->win_first, batch->win_first + batch->n_windows + 1);
  h->backbone_copy.assign(batch->bases, batch->bases + (batch->n_layers ? batch->seq_off[batch->n_layers] : 0));
  h->resident = true;
  return VGC_OK;
}

int vgc_polish_resident(vgc_handle h, vgc_result* result, vgc_stats* stats) {
  if (!h || !h->resident) {
    set_err("no resident batch: call vgc_upload first");
    return VGC_ERR_INVALID;
  }
  VGC_CUDA(cudaSetDevice(h->device));
  int rc = polish_device(h, result, stats, 0, h->backbone_copy.data(), h->r_seq_off.data(), h->r_win_first.data(),
                         h->r_n_windows);
  if (rc == VGC_OK && stats) stats->h2d_ms = 0.0;
  return rc;
}

}  // extern "C"

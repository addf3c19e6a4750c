/*
 * vgc.h — C-ABI of the B200 variation-graph correction engine (vechat_b200).
 *
 * This is the drop-in boundary for ONE path of HaploKit/vechat: the per-window
 * POA correction loop.  Each entry point below replaces a piece of the reference
 * (paths are under the reference tree):
 *
 *   vgc_create / vgc_destroy   <- racon::Polisher ctor: one spoa NW engine per thread
 *                                 (src/polisher.cpp:186-190) and the legacy per-device
 *                                 CUDABatchProcessor set-up (src/cuda/cudapolisher.cpp:229-241)
 *   vgc_polish                 <- the body of Polisher::polish's per-window lambda
 *                                 (src/polisher.cpp:498-516): Window::generate_consensus
 *                                 haplotype (src/window.cpp:176-428) or linear (src/window.cpp:74-174)
 *                                 for every window of a batch
 *   vgc_batch                  <- what `friend class CUDABatchProcessor` reads out of
 *                                 racon::Window (src/window.hpp:61-76): sequences_, qualities_,
 *                                 positions_, type_
 *   vgc_result                 <- Window::consensus_ + the bool generate_consensus returns
 *   vgc_last_error             <- the fprintf(stderr)+exit(1) / std::invalid_argument sites
 *                                 (src/window.cpp:24-27,58-67; vendor/spoa/src/graph.cpp:193,221,229)
 *
 * Plain pointers and sizes only; no C++ or torch types.  Host buffers in, host buffers out:
 * the library owns all device memory, streams and copies.  There is no CPU fallback: every
 * call fails with VGC_ERR_NO_DEVICE if no sm_100-class GPU is usable.
 */
#ifndef VGC_H_
#define VGC_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGC_OK                 0
#define VGC_ERR_INVALID        1   /* malformed batch / params (reference: exit(1) or invalid_argument) */
#define VGC_ERR_NO_DEVICE      2   /* no usable CUDA device: the engine never falls back to the CPU */
#define VGC_ERR_CUDA           3   /* CUDA runtime error (reference: GW_CU_CHECK_ERR abort)          */
#define VGC_ERR_CAPACITY       4   /* a window exceeded a hard engine limit (see vgc_limits)        */
#define VGC_ERR_NOMEM          5

/* Hard limits of the engine (the reference has none of the first two; vgc_limits() reports them at run time).
 * Within them every window runs: layers up to VGC_FAST_LAYER_LEN bases with scores inside the int16 range go through
 * the packed int16 kernels, anything else through the int32 "wide" kernel (the reference makes the same switch of
 * lane width, vendor/spoa/src/simd_alignment_engine_implementation.hpp:699-706). */
#define VGC_MAX_LAYER_LEN   16383u  /* bases of one layer (racon -w up to ~8000)                      */
#define VGC_MAX_CODES       16u     /* distinct base bytes per batch: A C G T N + all IUPAC codes     */
#define VGC_FAST_LAYER_LEN  1024u   /* longest layer the int16 kernels take                           */
#define VGC_MAX_BACKBONE    65535u  /* createWindow's own limit (src/window.cpp:216,232: uint16 loops) */

typedef struct vgc_engine* vgc_handle;

/* Scoring / pruning parameters: racon::createPolisher arguments that reach the hot path
 * (src/polisher.hpp:42-49) and src/main.cpp:46-66 defaults. */
typedef struct {
  int8_t   match;           /* -m, default 3  : NW engine (src/polisher.cpp:187-188)            */
  int8_t   mismatch;        /* -x, default -5                                                   */
  int8_t   gap;             /* -g, default -4 (linear gaps only: g == e)                        */
  uint8_t  haplotype;       /* -p: 1 = src/window.cpp:176 path, 0 = src/window.cpp:74 path      */
  uint8_t  trim;            /* !(-u/--no-trimming); only read by the linear path               */
  uint8_t  reserved[3];
  uint32_t num_prune;       /* -k, default 3                                                    */
  double   min_confidence;  /* -d                                                               */
  double   min_support;     /* -s                                                               */
} vgc_params;

/* One batch of windows, structure-of-arrays.  "Layer" = one element of Window::sequences_;
 * layer 0 of every window is the backbone (positions (0,0), src/window.cpp:38-40).
 * Layers of a window are listed in add_layer order (the order sequences_ holds them). */
typedef struct {
  uint32_t        n_windows;
  uint32_t        n_layers;     /* total over all windows, backbones included                  */
  const uint8_t*  bases;        /* concatenated layer bases (any byte < 128 — the reference indexes
                                   its coder with a signed char; at most VGC_MAX_CODES distinct
                                   values per batch)                                            */
  const uint8_t*  quals;        /* concatenated qualities, same offsets as bases; bytes of a
                                   layer with has_qual == 0 are ignored; may be NULL iff no
                                   layer has a quality                                          */
  const uint64_t* seq_off;      /* [n_layers + 1] offsets of each layer into bases / quals      */
  const uint8_t*  has_qual;     /* [n_layers] 0 <=> qualities_[i].first == nullptr              */
  const uint32_t* begin;        /* [n_layers] positions_[i].first  (backbone coordinates)       */
  const uint32_t* end;          /* [n_layers] positions_[i].second (inclusive)                  */
  const uint32_t* win_first;    /* [n_windows + 1] index of each window's backbone layer        */
  const uint8_t*  win_flags;    /* [n_windows] VGC_WIN_* bits                                   */
} vgc_batch;

#define VGC_WIN_TGS         1u  /* WindowType::kTGS (src/window.hpp:21-24)                       */
#define VGC_WIN_DUMMY_QUAL  2u  /* result of the C-string compare at src/window.cpp:223: the
                                   backbone quality pointer compares equal to a run of '!' of
                                   the backbone's length ("if_fasta")                            */

/* Corrected windows.  Buffers are owned by the caller.  cons must hold cons_capacity bytes;
 * cons_off [n_windows + 1]; polished [n_windows] (the bool generate_consensus returns). */
typedef struct {
  uint8_t*  cons;
  uint64_t  cons_capacity;
  uint64_t* cons_off;
  uint8_t*  polished;
} vgc_result;

/* Per-call statistics (optional, may be NULL).  cells = sum over alignments of (R_a+1)*L_a,
 * the unit SURVEY.md §8(d) defines the algorithmic bytes on. */
typedef struct {
  uint64_t cells;          /* DP cells filled                                                    */
  uint64_t alignments;     /* sequence-to-graph alignments run                                   */
  uint64_t input_bytes;    /* bytes copied host -> device                                        */
  uint64_t output_bytes;   /* bytes copied device -> host                                        */
  double   kernel_ms;      /* device time of the POA kernel(s), CUDA events on the launch stream  */
  double   h2d_ms;         /* first H2D copy -> last (the bulk copies overlap the host preparation)             */
  double   d2h_ms;
  double   device_ms;      /* first device op of the call -> last (H2D if any + kernels + D2H), CUDA events   */
  double   host_prep_ms;   /* host-side batch preparation (rank sort, average weights), wall clock          */
  uint32_t kernel_launches;
  uint32_t relaunched_windows; /* windows re-run with a larger scratch arena                      */
  double   host_pack_ms;   /* host-side packing of the bases (2 or 4 bits each), overlapped with host_prep_ms    */
} vgc_stats;

/* Fails with VGC_ERR_INVALID on the parameters spoa::AlignmentEngine::Create rejects
 * (vendor/spoa/src/alignment_engine.cpp:37-51: positive gap penalty), same message. */
int vgc_create(vgc_handle* out, int device, const vgc_params* params);
int vgc_destroy(vgc_handle h);

/* Upper bound of the corrected bytes a batch can produce (for sizing vgc_result.cons). */
uint64_t vgc_result_bound(const vgc_batch* batch);

/* Host buffers in, host buffers out (H2D + kernels + D2H inside the call). */
int vgc_polish(vgc_handle h, const vgc_batch* batch, vgc_result* result, vgc_stats* stats);

/* The same call in two halves, for callers with more than one batch: vgc_submit stages a batch —
 * host preparation, packing of the bases (2 bits each while the batch holds A C G T only, else 4), H2D on a copy
 * stream — on a worker thread and returns at once; vgc_collect waits for the oldest submitted batch and runs its
 * kernels + D2H.  Submitting batch i + 1 before collecting batch i overlaps its staging with the kernels of batch i
 * (device inputs are double-buffered; at most two batches may be in flight).  The arrays of a submitted batch must
 * stay valid until its vgc_collect returns.  vgc_polish(b) == vgc_submit(b) + vgc_collect(). */
int vgc_submit(vgc_handle h, const vgc_batch* batch);
int vgc_collect(vgc_handle h, vgc_result* result, vgc_stats* stats);

/* Device-resident variant used by bench.py's `value` leg: vgc_upload copies the batch to HBM
 * once; vgc_polish_resident runs only the kernels (+ the D2H of the corrected bytes when
 * result != NULL). */
int vgc_upload(vgc_handle h, const vgc_batch* batch);
int vgc_polish_resident(vgc_handle h, vgc_result* result, vgc_stats* stats);

/* Diagnostics: cycles the window leader lane spent per phase in the last polish call, summed over windows.
 * Index: 0 traceback cycles spent in tile-refill phases (part of 4), 1 toposort, 2 row program, 3 DP fill,
 * 4 traceback, 5 AddAlignment, 6 AddWeights, 7 PruneGraph, 8 LargestSubgraph, 9 emit/consensus,
 * 10 number of traceback tile refills (a count, not cycles), 11 host wall-clock milliseconds spent enqueueing the
 * kernel launches, 12 TopologicalSort runs, 13 how many of them sorted out of HBM because the staged graph did not fit
 * the kernel's shared memory. */
int vgc_phase_profile(vgc_handle h, double out[16]);

typedef struct {
  uint32_t max_layer_len;    /* VGC_MAX_LAYER_LEN                                                  */
  uint32_t max_backbone_len; /* VGC_MAX_BACKBONE                                                   */
  uint32_t max_codes;        /* VGC_MAX_CODES                                                      */
  uint32_t fast_layer_len;   /* VGC_FAST_LAYER_LEN: longer layers run on the int32 wide kernel     */
  uint32_t int16_score_bound;/* an alignment stays on the int16 kernels while
                                (max(rows, cols) + cols + 2) * max|score| <= this (rows = graph nodes,
                                cols = 512 / 640 / 1024, whichever fits the layer)                 */
  uint32_t reserved[3];
} vgc_limits_t;
void vgc_limits(vgc_limits_t* out);

/* Per-window status of the last vgc_polish / vgc_polish_resident call of this handle (0 = ok; otherwise the
 * engine's kSt* code of the window that made the call fail, vechat_b200/csrc/poa_core.h).  n = windows of the batch. */
int vgc_window_status(vgc_handle h, uint32_t* status, uint32_t n);

const char* vgc_last_error(void);
const char* vgc_version(void);

/* Host-side helper every caller needs (no device work): quality byte -> edge weight LUT,
 * vendor/spoa/src/graph.cpp:169 and src/window.cpp:366. */
void vgc_weight_lut(uint32_t lut[256]);

#ifdef __cplusplus
}
#endif
#endif  /* VGC_H_ */

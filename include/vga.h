/*
 * vga.h — C-ABI of the B200 overlap aligner (vechat_b200): SURVEY.md §8(f-1), the stage in front of the POA path.
 *
 * Replaces, for a whole batch of overlaps at once (paths under the reference tree):
 *
 *   vga_align      <- Overlap::align_overlaps (src/overlap.cpp:205-224): edlibAlign(q, t, EDLIB_MODE_NW,
 *                     EDLIB_TASK_PATH) + edlibAlignmentToCigar(EDLIB_CIGAR_STANDARD) for every overlap that has no
 *                     CIGAR yet, i.e. the work Polisher::find_overlap_breaking_points fans out to its thread pool
 *                     (src/polisher.cpp:464-489).  The legacy GPU hook of the same seam is
 *                     CUDAPolisher::find_overlap_breaking_points (src/cuda/cudapolisher.cpp:75-215).
 *   vga_batch      <- what `friend class CUDABatchAligner` reads out of racon::Overlap (src/overlap.hpp:79-81):
 *                     the query / target substrings of src/overlap.cpp:195-199
 *   vga_result     <- Overlap::cigar_ (the caller then runs the reference's own find_breaking_points_from_cigar,
 *                     src/overlap.cpp:226-292)
 *   vga_break      <- Overlap::find_breaking_points as a whole (src/overlap.cpp:179-203): alignment + the cut into
 *                     per-window breaking points; vga_breaks <- Overlap::breaking_points_
 *
 * Result: an optimal unit-cost global alignment (edit distance == edlib's).  Among equally good paths the choice is
 * this library's (mismatch, then deletion, then insertion, on furthest-reaching diagonals), identical to
 * oracle/shims/edlib_standin.cpp; the real edlib is not available here, so equality with ITS tie-breaking is
 * unpinned.  Plain pointers and sizes; host buffers in and out; no CPU fallback (VGA_ERR_NO_DEVICE).
 */
#ifndef VGA_H_
#define VGA_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGA_OK             0
#define VGA_ERR_INVALID    1
#define VGA_ERR_NO_DEVICE  2
#define VGA_ERR_CUDA       3
#define VGA_ERR_CAPACITY   4   /* an alignment needs more wavefront storage than the device has */
#define VGA_ERR_NOMEM      5

typedef struct vga_aligner* vga_handle;

/* Overlaps as (offset, length) pairs into one byte buffer that holds every sequence the batch touches once
 * (forward strand and, where an overlap needs it, the reverse complement). */
typedef struct {
  const uint8_t*  seqs;
  uint64_t        seqs_len;
  uint32_t        n;        /* overlaps */
  const uint64_t* q_off;    /* [n] query substring: &data[q_begin] or &reverse_complement[q_length - q_end] */
  const uint32_t* q_len;    /* [n] q_end - q_begin */
  const uint64_t* t_off;    /* [n] target substring: &data[t_begin] */
  const uint32_t* t_len;    /* [n] t_end - t_begin */
} vga_batch;

/* Filled by vga_align.  The buffers belong to the handle and stay valid until the next vga_align or vga_destroy on
 * it (their size is only known once the alignments exist).  cigar: the n strings back to back, each
 * NUL-terminated, string i at cigar + cigar_off[i] (EDLIB_CIGAR_STANDARD alphabet: M, I, D). */
typedef struct {
  const char*     cigar;
  const uint64_t* cigar_off;       /* [n + 1] */
  const int32_t*  edit_distance;   /* [n] */
} vga_result;

/* vga_break: per-overlap coordinates the breaking points are expressed in (src/overlap.cpp:238-239). */
typedef struct {
  const uint32_t* t_begin;        /* [n] Overlap::t_begin_ (the target substring starts there; t_end = t_begin + t_len) */
  const uint32_t* q_start;        /* [n] query coordinate of the substring's first base:
                                         strand ? q_length - q_end : q_begin */
  uint32_t        window_length;  /* Polisher::window_length_ */
} vga_cut;

/* Filled by vga_break; buffers belong to the handle like vga_result's.  Overlap i owns the pairs
 * [points_off[i], points_off[i + 1]); pair p = 4 words at points + 4 * p: first.t, first.q, last.t, last.q — two
 * consecutive elements of Overlap::breaking_points_ (first match of a window stretch, one past its last match). */
typedef struct {
  const uint32_t* points;
  const uint64_t* points_off;      /* [n + 1], in pairs */
  const int32_t*  edit_distance;   /* [n] */
} vga_breaks;

typedef struct {
  uint64_t cells;            /* wavefront cells computed = sum over overlaps of (D+1)^2 */
  uint64_t wavefront_bytes;  /* 4 B x cells: the algorithmic bytes of the kernel (each cell written once) */
  double   kernel_ms;        /* device time of the alignment kernel(s), CUDA events on the launch stream */
  double   total_ms;         /* whole call, wall clock */
  uint32_t kernel_launches;
  uint32_t retried;          /* overlaps re-run with a larger wavefront arena / output buffer */
} vga_stats;

int vga_create(vga_handle* out, int device);
int vga_destroy(vga_handle h);
int vga_align(vga_handle h, const vga_batch* batch, vga_result* result, vga_stats* stats);
/* Alignment + Overlap::find_breaking_points_from_cigar (src/overlap.cpp:226-292) on the device: only the breaking
 * points come back (no CIGAR text crosses PCIe).  Same alignments as vga_align. */
int vga_break(vga_handle h, const vga_batch* batch, const vga_cut* cut, vga_breaks* result, vga_stats* stats);
const char* vga_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VGA_H_ */

/* TEST INFRASTRUCTURE.  Stand-in for edlib v1.2.7 (FetchContent dependency of the reference, CMakeLists.txt:42-57;
 * source not in the reference tree, no network).  Declares the five entry points src/overlap.cpp:207-223 uses, with
 * edlib's published names and meanings.  The implementation (edlib_standin.cpp) is an exact unit-cost global
 * aligner of our own (furthest-reaching diagonals); its edit distance equals edlib's, but among equally good
 * paths it may choose another CIGAR than the real edlib would.  PARITY UNPINNED at this boundary (the reference
 * has no test that pins edlib's tie-breaking); it does not matter for the correction path, whose contract is
 * "identical overlaps and window tilings": the reference binary and the B200 binary are both built on this file. */
#ifndef ORACLE_SHIM_EDLIB_H_
#define ORACLE_SHIM_EDLIB_H_
#ifdef __cplusplus
extern "C" {
#endif

#define EDLIB_STATUS_OK 0
#define EDLIB_STATUS_ERROR 1

typedef enum { EDLIB_MODE_NW, EDLIB_MODE_SHW, EDLIB_MODE_HW } EdlibAlignMode;
typedef enum { EDLIB_TASK_DISTANCE, EDLIB_TASK_LOC, EDLIB_TASK_PATH } EdlibAlignTask;
typedef enum { EDLIB_CIGAR_STANDARD, EDLIB_CIGAR_EXTENDED } EdlibCigarFormat;

#define EDLIB_EDOP_MATCH 0    /* consumes query and target */
#define EDLIB_EDOP_INSERT 1   /* consumes query only ('I') */
#define EDLIB_EDOP_DELETE 2   /* consumes target only ('D') */
#define EDLIB_EDOP_MISMATCH 3 /* consumes both */

typedef struct { char first; char second; } EdlibEqualityPair;

typedef struct {
  int k;
  EdlibAlignMode mode;
  EdlibAlignTask task;
  const EdlibEqualityPair* additionalEqualities;
  int additionalEqualitiesLength;
} EdlibAlignConfig;

typedef struct {
  int status;
  int editDistance;
  int* endLocations;
  int* startLocations;
  int numLocations;
  unsigned char* alignment;
  int alignmentLength;
  int alphabetLength;
} EdlibAlignResult;

EdlibAlignConfig edlibNewAlignConfig(int k, EdlibAlignMode mode, EdlibAlignTask task,
                                     const EdlibEqualityPair* additionalEqualities, int additionalEqualitiesLength);
EdlibAlignResult edlibAlign(const char* query, int queryLength, const char* target, int targetLength,
                            const EdlibAlignConfig config);
void edlibFreeAlignResult(EdlibAlignResult result);
char* edlibAlignmentToCigar(const unsigned char* alignment, int alignmentLength, EdlibCigarFormat cigarFormat);

#ifdef __cplusplus
}
#endif
#endif

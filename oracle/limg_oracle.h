/* oracle/limg_oracle.h -- CPU restatement of limg's blocked three-factor encode/decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; limg_b200/ never links, imports or calls it. It is the checker, not the product.
 *
 * Parity status: PINNED. The reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself: oracle/_ref/libref.so (the unmodified
 * reference compiled by oracle/Makefile) is compared with every function below on seeded inputs by
 * tests/test_oracle_vs_ref.py, and the fixtures under tests/golden/ were produced by the reference
 * through tools/make_golden.py. It follows the reference's SSE4.1 dispatch path (the `*_sse41`
 * functions), never the scalar fallbacks, which disagree with it.
 *
 * Everything is scalar C: the SSE semantics that matter (dpps summation order, min/max operand order,
 * cvtps2dq rounding and its 0x80000000 "indefinite" result, the RSQRTPS table) are spelled out, so the
 * result does not depend on the host CPU.
 */
#ifndef LIMG_ORACLE_H
#define LIMG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Channel-agnostic decomposition record (reference: limg_encode_3d_output<channels>, limg_internal.h:343-353).
 * Unused channel slots are zero. 64 bytes. */
typedef struct lo_decomp
{
  float avg[4];
  int16_t dirA_min[4], dirA_max[4];
  int16_t dirB_offset[4], dirB_mag[4];
  int16_t dirC_offset[4], dirC_mag[4];
} lo_decomp;

/* One emitted area, in emission order (stage 0: large merges, 1: remaining merges, 2: leftover 8x8 blocks). */
typedef struct lo_area
{
  uint32_t ox, oy, rx, ry;           /* block units */
  uint32_t stage;
  uint32_t px_x, px_y, px_w, px_h;   /* pixel rectangle (limg.cpp:1722-1739) */
  uint8_t shift[3];
  uint8_t pad;
  uint64_t ditherBefore, ditherAfter;
  lo_decomp decomp;
} lo_area;

/* Caller-allocated planes, sizeX*sizeY elements each (reference: limg_blocked_encode3d_info, limg.h:39-44).
 * Any pointer may be NULL to skip that plane. pBlockError is never written (as in the reference). */
typedef struct lo_planes
{
  uint32_t *pDecoded;
  uint8_t *pFactorsA, *pFactorsB, *pFactorsC, *pBlockError, *pBitsPerPixel;
  uint32_t *pShiftABCX, *pColAMin, *pColAMax, *pColBMin, *pColBMax, *pColCMin, *pColCMax, *pBlockIndex;
} lo_planes;

enum { LO_DITHER_LCG = 0, LO_DITHER_AES = 1 };

float lo_rsqrt(float x);                                                                  /* RSQRTSS table function */
void lo_set_rsqrt_lut(const uint16_t *lut2048);                                           /* NULL restores the committed table */

void lo_fit(const uint32_t *pixels, size_t n, int channels, lo_decomp *out);              /* limg.cpp:469-497 + limg_factorization.h:385-794 */
int lo_matches(int channels, const lo_decomp *a, const lo_decomp *b);                     /* limg.cpp:1137-1269 */
void lo_project(int channels, const lo_decomp *d, const uint32_t *pixels, size_t n, uint8_t *fa, uint8_t *fb, uint8_t *fc); /* limg_factorization.h:101-197 */
int lo_trial(int channels, uint64_t maxPixelError, uint64_t maxBlockError, const lo_decomp *d, const uint32_t *pixels, size_t n,
             const uint8_t *fa, const uint8_t *fb, const uint8_t *fc, const uint8_t shift[3], uint64_t *blockError); /* limg_bit_crush_simd.h:315-810 */
void lo_search(int channels, uint32_t errorFactor, int fast, const lo_decomp *d, const uint32_t *pixels, size_t n,
               const uint8_t *fa, const uint8_t *fb, const uint8_t *fc, uint8_t shift[3]); /* limg.cpp:1512-1535 + limg_bit_crush.h:331-1051 */
uint64_t lo_dither(int mode, uint8_t shift, size_t n, uint64_t state, uint8_t *factors);  /* limg.cpp:798-887 */
void lo_decode(int channels, uint32_t *out, size_t stride, size_t rx, size_t ry, const uint8_t *fa, const uint8_t *fb, const uint8_t *fc,
               const lo_decomp *d, const uint8_t shift[3]);                               /* limg_decode.h:39-236 */

void lo_pass1(const uint32_t *img, size_t sizeX, size_t sizeY, int channels, lo_decomp *table); /* limg.cpp:1088-1119 */
/* Greedy area map (limg.cpp:1121-1135, 1277-1496, 1814-1878). Fills ox,oy,rx,ry,stage; returns the count. */
size_t lo_merge(const lo_decomp *table, size_t blockX, size_t blockY, int channels, lo_area *areas, uint64_t *stats /* [8] or NULL */);

/* limg_blocked_encode3d_test (limg.cpp:2329-2453). areas may be NULL; otherwise capacity blockX*blockY. Returns the area count. */
size_t lo_blocked_encode3d(const uint32_t *img, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, int fast, int ditherMode,
                           lo_planes *planes, lo_area *areas);
/* limg_encode3d_test (limg.cpp:2175-2265): no merge; `bands` y-bands each restart the dither chain (limg.cpp:1893, 2114-2134);
 * bands <= 1 is the pool-less call. */
void lo_encode3d(const uint32_t *img, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, int fast, int ditherMode, size_t poolThreads,
                 lo_planes *planes);
/* Standalone reconstruction from an area list + right-aligned factor streams in area-contiguous emission order. */
void lo_decode_areas(int channels, const lo_area *areas, size_t count, const uint8_t *fa, const uint8_t *fb, const uint8_t *fc, uint32_t *out, size_t sizeX);

double lo_compare(const uint32_t *a, const uint32_t *b, size_t sizeX, size_t sizeY, int hasAlpha, double *mse, double *maxErr); /* limg.cpp:2455-2491 */

#ifdef __cplusplus
}
#endif

#endif

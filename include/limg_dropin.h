/* include/limg_dropin.h -- C++ drop-in for the reference's public header.
 *
 * Use it in place of the reference's src/limg.h (e.g. `ln -s limg_dropin.h limg.h`) and link liblimgcu.so instead
 * of limg.cpp / limg_threading.cpp / limg_simd.cpp: the reference CLI (src/main.cpp) then runs its encode path on a
 * B200 unchanged. Every declaration below keeps the reference's name, argument order, argument meaning and struct
 * layout (reference file:line given per item); the implementations are host wrappers (H2D copy, sm_100a kernels,
 * D2H copy) over the C ABI of include/limgcu.h. There is no CPU fallback: without a CUDA device every encoder returns
 * limg_error_Generic.
 *
 * Behavioural notes (also in INTEGRATION.md):
 *   - `pThreadPool` is a token: nothing runs on host threads. For limg_encode3d_test the reference restarts its dither chain per
 *     y-band of the pool (limg.cpp:1893, 2108-2137), so its output depends on the pool size; the wrapper reproduces the run with a
 *     pool of limg_thread_pool_thread_count(pThreadPool) threads (and the pool-less run for nullptr).
 *   - The dither generator follows the reference's own rule (limg.cpp:881-887): the AES round chain (limg.cpp:824-879) on hosts with
 *     SSE4.1 + AES-NI, the PCG-style LCG (limg.cpp:799-822) otherwise, so the output equals the reference run on the same host.
 *     The AES chain has no skip-ahead and is walked on the host once the GPU has found the shifts (a few ms per 4K frame);
 *     LIMGCU_DITHER=lcg selects the LCG everywhere (area maps, shifts and endpoints are identical either way, the factor bytes differ).
 *   - limg_encode_test (the legacy one-factor codec, not reachable from the CLI) returns limg_error_Generic.
 *   - Null pointers are rejected with limg_error_ArgumentNull (the reference dereferences them).
 */
#ifndef LIMG_DROPIN_H
#define LIMG_DROPIN_H

#include <stddef.h>
#include <stdint.h>

/* reference: limg_threading.h:7-19. The pool is only a token here (the GPU replaces the y-band fan-out). */
struct limg_thread_pool;
limg_thread_pool *limg_thread_pool_new(const size_t threads);
void limg_thread_pool_destroy(limg_thread_pool **ppThreadPool);
size_t limg_thread_pool_thread_count(limg_thread_pool *pThreadPool);
void limg_thread_pool_await(limg_thread_pool *pThreadPool);
size_t limg_threading_max_threads();

/* reference: limg.h:9-18 */
enum limg_result
{
  limg_success = 0,
  limg_error_Generic = 100,
  limg_error_InvalidParameter,
  limg_error_ArgumentNull,
  limg_error_OutOfBounds,
  limg_error_MemoryAllocationFailure,
};

/* reference: limg.h:20-25 (legacy codec, kept for source compatibility only) */
struct limg_encode_info
{
  uint32_t *pDecoded, *pA, *pB, *pBlockIndex;
  uint8_t *pFactors, *pBlockError, *pShift;
  size_t totalBlockArea;
};

/* reference: limg.h:27 -- not implemented, returns limg_error_Generic */
limg_result limg_encode_test(const uint32_t *pIn, const size_t sizeX, const size_t sizeY, const bool hasAlpha, limg_encode_info *pInfo, const uint32_t errorFactor);

/* reference: limg.h:29-33. Caller-allocated planes of sizeX * sizeY elements. */
struct limg_encode3d_info
{
  uint32_t *pDecoded, *pShiftABCX, *pColAMin, *pColAMax, *pColBMin, *pColBMax, *pColCMin, *pColCMax;
  uint8_t *pFactorsA, *pFactorsB, *pFactorsC;
};

/* reference: limg.h:35, limg.cpp:2175-2265. Every 8x8 block is its own area. */
limg_result limg_encode3d_test(const uint32_t *pIn, const size_t sizeX, const size_t sizeY, const bool hasAlpha, limg_encode3d_info *pInfo, const uint32_t errorFactor, limg_thread_pool *pThreadPool, const bool fastBitCrushing);

/* reference: limg.h:37, limg.cpp:2267-2327. Same work, nothing written. */
limg_result limg_encode3d_test_perf(const uint32_t *pIn, const size_t sizeX, const size_t sizeY, const bool hasAlpha, const uint32_t errorFactor, limg_thread_pool *pThreadPool, const bool fastBitCrushing);

/* reference: limg.h:39-44. pBlockError is never written by the reference's 3D paths, nor here. */
struct limg_blocked_encode3d_info
{
  uint32_t *pDecoded;
  uint8_t *pFactorsA, *pFactorsB, *pFactorsC, *pBlockError, *pBitsPerPixel;
  uint32_t *pShiftABCX, *pColAMin, *pColAMax, *pColBMin, *pColBMax, *pColCMin, *pColCMax, *pBlockIndex;
};

/* reference: limg.h:46, limg.cpp:2329-2453. The CLI's default path: fit, greedy area merge, refit, projection, bit-crush search, dither, decode. */
limg_result limg_blocked_encode3d_test(const uint32_t *pIn, const size_t sizeX, const size_t sizeY, const bool hasAlpha, limg_blocked_encode3d_info *pInfo, const uint32_t errorFactor, limg_thread_pool *pThreadPool, const bool fastBitCrushing);

/* reference: limg.h:48, limg.cpp:2455-2491. Perceptually weighted PSNR. */
double limg_compare(const uint32_t *pImageA, const uint32_t *pImageB, const size_t sizeX, const size_t sizeY, const bool hasAlpha, double *pMeanSquaredError, double *pMaxPossibleSquaredError);

/* Additions (not in the reference): select the CUDA device used by the wrappers above (default 0). */
limg_result limg_b200_set_device(const int device);
/* Dither generator of the wrappers above: 1 = the AES-round chain the reference uses on a host with AES-NI (limg.cpp:824-879; walked on the
 * host, the call synchronises), 0 = the reference's LCG (limg.cpp:799-822; fully on the device), -1 = decide like the reference does from
 * the host CPU (limg.cpp:881-887), which is the default. Returns the generator now in use (1 / 0) or -1 without a device. */
int limg_b200_set_dither_mode(const int aes);

#endif /* LIMG_DROPIN_H */

// limg_b200/csrc/limg_api.cpp -- the reference's C++ entry points (include/limg_dropin.h) as host wrappers over the C ABI.
// Mirrors limg.cpp:2175-2491 at the interface level only; all work happens in the sm_100a kernels behind limgcu_host_*.
#include "../../include/limg_dropin.h"
#include "../../include/limgcu.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <thread>

struct limg_thread_pool
{
  size_t threads;
};

limg_thread_pool *limg_thread_pool_new(const size_t threads) { return new limg_thread_pool{ threads }; }

void limg_thread_pool_destroy(limg_thread_pool **ppThreadPool)
{
  if (ppThreadPool == nullptr || *ppThreadPool == nullptr)
    return;

  delete *ppThreadPool;
  *ppThreadPool = nullptr;
}

size_t limg_thread_pool_thread_count(limg_thread_pool *pThreadPool) { return pThreadPool ? pThreadPool->threads : 0; }
void limg_thread_pool_await(limg_thread_pool *) {}
size_t limg_threading_max_threads() { return std::thread::hardware_concurrency(); }

namespace
{
std::mutex g_mutex;
limgcu_ctx *g_ctx = nullptr;
int g_device = 0;

// one process-wide context, created on first use; calls are serialised (the reference is not re-entrant on a shared pool either)
limgcu_ctx *context()
{
  if (g_ctx == nullptr)
  {
    if (limgcu_create(g_device, &g_ctx) != LIMGCU_SUCCESS)
    {
      g_ctx = nullptr;
      return nullptr;
    }

    // The reference picks its dither generator from the host CPU (limg.cpp:881-887: AES rounds with SSE4.1 + AES-NI, the LCG otherwise);
    // the drop-in follows the same rule so that it reproduces the reference run on this host. LIMGCU_DITHER=lcg / aes overrides it.
    const char *mode = getenv("LIMGCU_DITHER");
    limgcu_set_dither_mode(g_ctx, mode ? (strcmp(mode, "aes") == 0) : limgcu_host_has_aesni());
  }

  return g_ctx;
}

limg_result to_result(const int rc)
{
  switch (rc)
  {
  case LIMGCU_SUCCESS: return limg_success;
  case LIMGCU_ERROR_INVALID_PARAMETER: return limg_error_InvalidParameter;
  case LIMGCU_ERROR_ARGUMENT_NULL: return limg_error_ArgumentNull;
  case LIMGCU_ERROR_OUT_OF_BOUNDS: return limg_error_OutOfBounds;
  case LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE: return limg_error_MemoryAllocationFailure;
  default: return limg_error_Generic;
  }
}
} // namespace

limg_result limg_b200_set_device(const int device)
{
  std::lock_guard<std::mutex> lock(g_mutex);

  if (g_ctx != nullptr)
  {
    limgcu_destroy(g_ctx);
    g_ctx = nullptr;
  }

  g_device = device;
  return context() ? limg_success : limg_error_Generic;
}

int limg_b200_set_dither_mode(const int aes)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  limgcu_ctx *ctx = context();

  if (ctx == nullptr)
    return -1;

  const int mode = aes < 0 ? limgcu_host_has_aesni() : (aes ? 1 : 0);
  limgcu_set_dither_mode(ctx, mode);
  return mode;
}

limg_result limg_encode_test(const uint32_t *, const size_t, const size_t, const bool, limg_encode_info *, const uint32_t)
{
  return limg_error_Generic; // legacy one-factor codec: out of scope
}

limg_result limg_encode3d_test(const uint32_t *pIn, const size_t sizeX, const size_t sizeY, const bool hasAlpha, limg_encode3d_info *pInfo, const uint32_t errorFactor, limg_thread_pool *pThreadPool, const bool fastBitCrushing)
{
  if (pIn == nullptr || pInfo == nullptr)
    return limg_error_ArgumentNull;

  std::lock_guard<std::mutex> lock(g_mutex);
  limgcu_ctx *ctx = context();

  if (ctx == nullptr)
    return limg_error_Generic;

  limgcu_planes p = {};
  p.pDecoded = pInfo->pDecoded; p.pShiftABCX = pInfo->pShiftABCX;
  p.pColAMin = pInfo->pColAMin; p.pColAMax = pInfo->pColAMax; p.pColBMin = pInfo->pColBMin; p.pColBMax = pInfo->pColBMax;
  p.pColCMin = pInfo->pColCMin; p.pColCMax = pInfo->pColCMax;
  p.pFactorsA = pInfo->pFactorsA; p.pFactorsB = pInfo->pFactorsB; p.pFactorsC = pInfo->pFactorsC;
  // the reference's result depends on the pool size (one dither chain per y-band of the pool, limg.cpp:1893, 2108-2137): reproduce it
  limgcu_set_pool_threads(ctx, pThreadPool ? (int)pThreadPool->threads : 0);
  const int rc = limgcu_host_encode3d(ctx, pIn, sizeX, sizeY, hasAlpha ? 1 : 0, &p, errorFactor, fastBitCrushing ? 1 : 0);
  limgcu_set_pool_threads(ctx, 0);
  return to_result(rc);
}

limg_result limg_encode3d_test_perf(const uint32_t *pIn, const size_t sizeX, const size_t sizeY, const bool hasAlpha, const uint32_t errorFactor, limg_thread_pool *, const bool fastBitCrushing)
{
  if (pIn == nullptr)
    return limg_error_ArgumentNull;

  std::lock_guard<std::mutex> lock(g_mutex);
  limgcu_ctx *ctx = context();

  if (ctx == nullptr)
    return limg_error_Generic;

  return to_result(limgcu_host_encode3d(ctx, pIn, sizeX, sizeY, hasAlpha ? 1 : 0, nullptr, errorFactor, fastBitCrushing ? 1 : 0));
}

limg_result limg_blocked_encode3d_test(const uint32_t *pIn, const size_t sizeX, const size_t sizeY, const bool hasAlpha, limg_blocked_encode3d_info *pInfo, const uint32_t errorFactor, limg_thread_pool *, const bool fastBitCrushing)
{
  if (pIn == nullptr || pInfo == nullptr)
    return limg_error_ArgumentNull;

  std::lock_guard<std::mutex> lock(g_mutex);
  limgcu_ctx *ctx = context();

  if (ctx == nullptr)
    return limg_error_Generic;

  limgcu_planes p = {};
  p.pDecoded = pInfo->pDecoded;
  p.pFactorsA = pInfo->pFactorsA; p.pFactorsB = pInfo->pFactorsB; p.pFactorsC = pInfo->pFactorsC;
  p.pBlockError = pInfo->pBlockError; p.pBitsPerPixel = pInfo->pBitsPerPixel;
  p.pShiftABCX = pInfo->pShiftABCX;
  p.pColAMin = pInfo->pColAMin; p.pColAMax = pInfo->pColAMax; p.pColBMin = pInfo->pColBMin; p.pColBMax = pInfo->pColBMax;
  p.pColCMin = pInfo->pColCMin; p.pColCMax = pInfo->pColCMax; p.pBlockIndex = pInfo->pBlockIndex;
  return to_result(limgcu_host_blocked_encode3d(ctx, pIn, sizeX, sizeY, hasAlpha ? 1 : 0, &p, errorFactor, fastBitCrushing ? 1 : 0));
}

double limg_compare(const uint32_t *pImageA, const uint32_t *pImageB, const size_t sizeX, const size_t sizeY, const bool hasAlpha, double *pMeanSquaredError, double *pMaxPossibleSquaredError)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  limgcu_ctx *ctx = context();

  if (ctx == nullptr)
    return NAN;

  return limgcu_host_compare(ctx, pImageA, pImageB, sizeX, sizeY, hasAlpha ? 1 : 0, pMeanSquaredError, pMaxPossibleSquaredError);
}

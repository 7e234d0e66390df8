// limg_b200/csrc/limgcu.cu -- context + C ABI (include/limgcu.h) over the sm_100a kernels.
// There is no CPU path in this file: every entry point either launches the kernels or fails.
#include "kernels_fit.cuh"
#include "kernels_merge.cuh"
#include "kernels_wave.cuh"
#include "kernels_cta.cuh"
#include "kernels_stream.cuh"
#include "kernels_decode.cuh"
#include "kernels_container.cuh"
#include "rsqrt_lut.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <thread>
#include <vector>

using namespace limg;

namespace limg
{
// dither_aes_host.cpp
uint64_t aes_dither_chain_host(const limgcu_area *areas, uint32_t count, uint64_t seed, uint8_t *noise, uint64_t *planeOff, uint64_t *before, uint64_t *after, bool forceSoftware,
                               uint32_t bandAreas, uint32_t bandCount);
bool host_has_aesni();
}

enum { PHASE_PASS1 = 0, PHASE_WINDOW, PHASE_SCAN, PHASE_ENCODE, PHASE_DITHER, PHASE_FINALIZE, PHASE_COUNT };

struct limgcu_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t streamAux = nullptr; // the speculative match bitmaps are computed here while the scan already runs on `stream`
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evFork2 = nullptr, evJoin2 = nullptr, evBand[4] = { nullptr, nullptr, nullptr, nullptr };
  const uint32_t *hostSrc = nullptr; // set by host_encode: limgcu_blocked_encode3d uploads d_src from here in bands, pass 1 of band i under the upload of band i + 1
  int planAsync = 1;                // LIMGCU_PLAN_ASYNC: 0 plan kernels on the main stream, 1 both on the second stream concurrently with the scan, 2 only k_plan_sym
  int planAsyncCtas = 6;            // LIMGCU_PLAN_CTAS: CTAs per SM of an asynchronous plan kernel (the scan is one cluster on 8 or 16 SMs, so they can fill the others: 2 -> 6 is 2 % of a 4K encode)
  char err[512] = { 0 };
  uint64_t launches = 0;
  int smCount = 148;

  uint16_t *dLut = nullptr;
  LcgJumpTable jt;

  // per-image working set, grown on demand
  size_t capBlocks = 0, capDecodeBlocks = 0, capPixels = 0, capUsedWords = 0;
  limgcu_decomp *dTable = nullptr;
  PredRec *dRec = nullptr;
  uint32_t *dWindow = nullptr;
  limgcu_area *dAreas = nullptr;
  uint32_t *dBlockToArea = nullptr;
  AreaWork *dWork = nullptr;
  uint32_t *dSmallList = nullptr, *dLargeList = nullptr;
  uint64_t *dDemand = nullptr;
  unsigned long long *dDitherBefore = nullptr;
  // AES dither mode (dither_aes_host.cpp): pinned host staging + device copies, allocated on first use
  int ditherAes = 0;                 // limgcu_set_dither_mode / LIMGCU_DITHER=aes: default generator of the entry points without a flags argument
  int aesForceSoftware = 0;          // LIMGCU_AES_SOFTWARE=1: software AES rounds even on a host with AES-NI (test hook)
  limgcu_area *hAreas = nullptr;
  uint8_t *hNoise = nullptr, *dNoise = nullptr;
  unsigned long long *hNoiseOff = nullptr, *dNoiseOff = nullptr, *hStates = nullptr, *dStates = nullptr;
  size_t capAesBlocks = 0, capAesPixels = 0;
  uint32_t *dUsed = nullptr;
  uint32_t *dExtSlot = nullptr, *dExtSeed = nullptr, *dExtBits = nullptr, *dExtHdr = nullptr, *dPlanCounters = nullptr;
  uint4 *dSeedSym = nullptr;
  uint32_t *dSymSlot = nullptr, *dSymSeed = nullptr, *dSymBits = nullptr, *dSymHdr = nullptr, *dSymStart = nullptr;
  uint16_t *dUnmasked = nullptr;
  uint8_t *dLeftRun = nullptr;  // per block: matches to its left (k_pred_leftrun)
  uint32_t *dSafe = nullptr;    // per block: safe columns of the scan's stage 0 (k_plan_safe)
  uint32_t extCap = 0, symCap = 0;
  uint32_t *dScratchPx = nullptr, *dScratchFac = nullptr;
  uint32_t *dCounters = nullptr; // [32]: 0 merged, 1 areaCount, 2 smallCount, 3 largeCount (fits shared memory), 4 workSmall, 5 workLarge, 6 scratchTop, 7 hugeCount, 8.. stats, 24..28 + 31 scan flags, 29 bigCount, 30 work counter of the huge areas
  unsigned long long *dCompare = nullptr;

  // host-buffer staging
  size_t capStagePixels = 0, capDecodeStagePixels = 0;
  uint32_t *dSrc = nullptr;
  uint32_t *dPlaneU32[9] = { nullptr };
  uint8_t *dPlaneU8[7] = { nullptr }; // factors A,B,C, bpp, codes A,B,C
  uint8_t *dPayload = nullptr;        // bit-packed container payload (kernels_container.cuh)
  unsigned long long *dPayloadOff = nullptr;
  size_t capPayload = 0, capPayloadOff = 0;

  // wavefront merge (kernels_wave.cuh)
  uint32_t *dWaveZero = nullptr; // flags[8] pad[2] candCount[2] counters[4] | progress[2*BY] | rowCounts[2*BY] | candBits[2*usedWords] | emitInfo[2*blocks]
  uint32_t *dReplayList = nullptr, *dReplayCount = nullptr;
  uint32_t *dTau = nullptr, *dCandList = nullptr, *dWaveDbg = nullptr, *dWaveRows = nullptr;
  size_t capWaveRows = 0;
  int waveRowTimes = 0;              // LIMGCU_MERGE_ROWTIMES=1: per-row time stamps of the scan (limgcu_debug_wave_rows)
  uint2 *dRowLists = nullptr;
  uint32_t *dRowMeta = nullptr; // rowLeft[BY], rowBase[3][BY]
  size_t capRowMeta = 0;
  size_t capWaveZero = 0, capRowLists = 0;

  int mergeExt = 1;                  // LIMGCU_MERGE_EXT=0 disables the speculative match bitmaps (everything beyond the 8x8 window on demand)
  int mergeMode = 0;                 // LIMGCU_MERGE_MODE: 0 wave (pipelined rows + verification), 1 seq (rows strictly in sequence)
  int scanCluster = 8;               // LIMGCU_SCAN_CLUSTER: CTAs of the cluster that runs the scan with its state in shared memory (k_merge_cta); 0 = scan over the mask in global memory (k_merge_wave); unset: 8, or 16 for 8K-class frames
  bool scanClusterSet = false;
  int pass1Tma = 1;                  // LIMGCU_PASS1_TMA=0: pass 1 stages its pixels with plain loads (k_pass1) instead of tensor-map tile loads (k_pass1_tma)
  int poolThreads = 0;               // limgcu_set_pool_threads: the non-merged encoder restarts its dither chain per y-band of a pool of this many threads, as the reference does (0: pool-less)
  int scanExperiment = 0;            // LIMGCU_SCAN_EXPERIMENT: measurement switches of the scan (WaveArgs::experiment), 0 in production
  int planBands = 8;                 // LIMGCU_PLAN_BANDS: bands of block rows the extension bitmaps and the centre bitmaps behind them are built in, top-down
  int scanWarps = LIMG_CTA_WARPS;    // LIMGCU_SCAN_WARPS: warps (block rows in flight) per CTA of that cluster, 1..8; 255 registers per thread, so 8 warps take an SM's whole register file, 4 leave half of it to other kernels
  int scanSmemLimit = 0;             // bytes of dynamic shared memory a CTA may opt in to (the mask replica has to fit)
  int planExtW = 16, planSymL = 6, planSymR = 12, planSymD = 16; // LIMGCU_PLAN_EXTW / SYML / SYMR / SYMD: size caps of the speculative bitmaps
  int mergeGap = 16;                 // LIMGCU_MERGE_GAP: block rows stage 1 stays behind stage 0
  int mergeWideMargin = 64;          // margin (and stage gap) of the second try
  int mergeSpec = 8;                 // LIMGCU_MERGE_SPEC: columns of lookahead for the speculative expansion
  int decodeVariant = 20;            // LIMGCU_DECODE_VARIANT: 0 generic k_decode; 2 / 4 / 8 = rows per thread of k_decode_tile (width % 8 == 0)
  int mergeMargin = 0;               // LIMGCU_MERGE_MARGIN: columns a row stays behind the rows above, beyond what its seed probed (the rows above publish SAFE columns, so
                                     // this is slack, not the protection against leftward regrowth it was before k_plan_safe: 8 then)
  int mergeSafe = 1;                 // LIMGCU_MERGE_SAFE=0: rows publish their plain progress (no safe columns; use with LIMGCU_MERGE_MARGIN=8)
  bool timing = false;
  cudaEvent_t ev[PHASE_COUNT + 1] = { nullptr };
  float phaseMs[PHASE_COUNT] = { 0 };
};

static int fail(limgcu_ctx *ctx, int code, const char *what, cudaError_t e)
{
  if (ctx)
    snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));

  if (e == cudaErrorMemoryAllocation)
    return LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE;

  return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, LIMGCU_ERROR_CUDA, #call, e_); } while (0)
#define CKL(what) do { ctx->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail(ctx, LIMGCU_ERROR_CUDA, what, e_); } while (0)
#define NEED(p) do { if ((p) == nullptr) return fail(ctx, LIMGCU_ERROR_ARGUMENT_NULL, #p " is null", cudaSuccess); } while (0)

template <class T>
static cudaError_t regrow(T *&p, size_t count)
{
  if (p)
    cudaFree(p);

  p = nullptr;
  const cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T));

  if (e != cudaSuccess)
  {
    p = nullptr;
    cudaGetLastError(); // a failed allocation must not be reported again by the next launch check
  }

  return e;
}

// A capacity is set to 0 before its buffers are freed, so a failed allocation leaves the context consistent (the next call allocates again).

// what the decoders need per block: the area table and the block map
static int ensure_decode_capacity(limgcu_ctx *ctx, size_t W, size_t H)
{
  const size_t blocks = ((W + 7) / 8) * ((H + 7) / 8);

  if (blocks > ctx->capDecodeBlocks)
  {
    ctx->capDecodeBlocks = 0;
    CK(regrow(ctx->dAreas, blocks));
    CK(regrow(ctx->dBlockToArea, blocks));
    ctx->capDecodeBlocks = blocks;
  }

  return LIMGCU_SUCCESS;
}

static int ensure_capacity(limgcu_ctx *ctx, size_t W, size_t H)
{
  const size_t BX = (W + 7) / 8, BY = (H + 7) / 8;
  const size_t blocks = BX * BY, pixels = W * H;
  const size_t usedWords = BY * ((BX + 31) / 32 + 2);
  int rc = ensure_decode_capacity(ctx, W, H);
  if (rc) return rc;

  if (blocks > ctx->capBlocks)
  {
    ctx->capBlocks = 0;
    CK(regrow(ctx->dTable, blocks));
    CK(regrow(ctx->dRec, blocks));
    CK(regrow(ctx->dWindow, blocks * 2));
    CK(regrow(ctx->dWork, blocks));
    CK(regrow(ctx->dSmallList, blocks));
    CK(regrow(ctx->dLargeList, blocks));
    CK(regrow(ctx->dDemand, blocks));
    CK(regrow(ctx->dDitherBefore, blocks));
    ctx->extCap = (uint32_t)(blocks / 2 > 1024 ? blocks / 2 : 1024);
    ctx->symCap = (uint32_t)(blocks > 1024 ? blocks : 1024);
    CK(regrow(ctx->dUnmasked, blocks));
    CK(regrow(ctx->dLeftRun, blocks));
    CK(regrow(ctx->dSafe, blocks));
    CK(regrow(ctx->dExtSlot, blocks));
    CK(regrow(ctx->dExtSeed, (size_t)ctx->extCap));
    CK(regrow(ctx->dExtBits, (size_t)ctx->extCap * 32));
    CK(regrow(ctx->dExtHdr, (size_t)ctx->extCap));
    CK(regrow(ctx->dSymSlot, blocks));
    CK(regrow(ctx->dSymStart, blocks));
    CK(regrow(ctx->dSeedSym, blocks));
    CK(regrow(ctx->dSymSeed, (size_t)ctx->symCap));
    CK(regrow(ctx->dSymBits, (size_t)ctx->symCap * 32));
    CK(regrow(ctx->dSymHdr, (size_t)ctx->symCap));
    if (ctx->dPlanCounters == nullptr) CK(regrow(ctx->dPlanCounters, (size_t)8));
    CK(regrow(ctx->dTau, blocks));
    CK(regrow(ctx->dCandList, blocks * 2));
    CK(regrow(ctx->dReplayList, blocks * 2));
    if (ctx->dReplayCount == nullptr) CK(regrow(ctx->dReplayCount, (size_t)2));
    ctx->capBlocks = blocks;
  }

  {
    const size_t waveZero = 16 + 4 * BY + 2 * usedWords + 2 * blocks, rowLists = 2 * BY * (2 * BX);

    if (waveZero > ctx->capWaveZero)
    {
      ctx->capWaveZero = 0;
      CK(regrow(ctx->dWaveZero, waveZero));
      ctx->capWaveZero = waveZero;
    }

    if (4 * BY > ctx->capRowMeta)
    {
      ctx->capRowMeta = 0;
      CK(regrow(ctx->dRowMeta, 4 * BY));
      ctx->capRowMeta = 4 * BY;
    }

    if (rowLists > ctx->capRowLists)
    {
      ctx->capRowLists = 0;
      CK(regrow(ctx->dRowLists, rowLists));
      ctx->capRowLists = rowLists;
    }
  }

  if (usedWords > ctx->capUsedWords)
  {
    ctx->capUsedWords = 0;
    CK(regrow(ctx->dUsed, usedWords));
    ctx->capUsedWords = usedWords;
  }

  if (pixels > ctx->capPixels)
  {
    ctx->capPixels = 0;
    CK(regrow(ctx->dScratchPx, pixels));
    CK(regrow(ctx->dScratchFac, pixels));
    ctx->capPixels = pixels;
  }

  return LIMGCU_SUCCESS;
}

// what the host-buffer decoders stage per pixel: the three code planes and the output (7 B/px; the encoders' staging is 47 B/px)
static int ensure_decode_staging(limgcu_ctx *ctx, size_t pixels)
{
  if (pixels <= ctx->capDecodeStagePixels)
    return LIMGCU_SUCCESS;

  ctx->capDecodeStagePixels = 0;
  CK(regrow(ctx->dPlaneU32[0], pixels));

  for (int i = 4; i < 7; i++)
    CK(regrow(ctx->dPlaneU8[i], pixels));

  ctx->capDecodeStagePixels = pixels;
  return LIMGCU_SUCCESS;
}

static int ensure_staging(limgcu_ctx *ctx, size_t pixels)
{
  int rc = ensure_decode_staging(ctx, pixels);
  if (rc) return rc;

  if (pixels <= ctx->capStagePixels)
    return LIMGCU_SUCCESS;

  ctx->capStagePixels = 0;
  CK(regrow(ctx->dSrc, pixels));

  for (int i = 1; i < 9; i++)
    CK(regrow(ctx->dPlaneU32[i], pixels));

  for (int i = 0; i < 4; i++)
    CK(regrow(ctx->dPlaneU8[i], pixels));

  ctx->capStagePixels = pixels;
  return LIMGCU_SUCCESS;
}

static bool aligned32(const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

extern "C" int limgcu_device_count(void)
{
  int n = 0;

  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;

  return n;
}

extern "C" int limgcu_create(int device, limgcu_ctx **out)
{
  if (out == nullptr)
    return LIMGCU_ERROR_ARGUMENT_NULL;

  *out = nullptr;
  int n = 0;

  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return LIMGCU_ERROR_NO_DEVICE; // no CPU fallback, by design

  if (device < 0 || device >= n)
    return LIMGCU_ERROR_INVALID_PARAMETER;

  limgcu_ctx *ctx = new (std::nothrow) limgcu_ctx();

  if (ctx == nullptr)
    return LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE;

  ctx->device = device;

  auto bail = [&](int code) {
    limgcu_destroy(ctx);
    return code;
  };

  if (cudaSetDevice(device) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
  {
    int prLow = 0, prHigh = 0;
    cudaDeviceGetStreamPriorityRange(&prLow, &prHigh);
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prHigh) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
    if (cudaStreamCreateWithPriority(&ctx->streamAux, cudaStreamNonBlocking, prLow) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
    if (cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
    if (cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
    if (cudaEventCreateWithFlags(&ctx->evFork2, cudaEventDisableTiming) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
    if (cudaEventCreateWithFlags(&ctx->evJoin2, cudaEventDisableTiming) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
    for (auto &e : ctx->evBand)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
  }
  cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device);

  if (cudaMalloc(&ctx->dLut, 2048 * sizeof(uint16_t)) != cudaSuccess) return bail(LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE);
  if (cudaMemcpy(ctx->dLut, LIMG_RSQRT_LUT, 2048 * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);
  if (cudaMalloc(&ctx->dCounters, 32 * sizeof(uint32_t)) != cudaSuccess) return bail(LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE);
  if (cudaMalloc(&ctx->dCompare, sizeof(unsigned long long)) != cudaSuccess) return bail(LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE);
  if (cudaMalloc(&ctx->dWaveDbg, 256 * sizeof(uint32_t)) != cudaSuccess) return bail(LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE);

  // LCG jump table: f^(2^j)(h) = mul[j] * h + add[j]
  ctx->jt.mul[0] = LIMG_LCG_MUL;
  ctx->jt.add[0] = 1;

  for (int j = 1; j < 64; j++)
  {
    ctx->jt.mul[j] = ctx->jt.mul[j - 1] * ctx->jt.mul[j - 1];
    ctx->jt.add[j] = ctx->jt.add[j - 1] * (ctx->jt.mul[j - 1] + 1);
  }

  if (const char *v = getenv("LIMGCU_MERGE_EXT")) ctx->mergeExt = atoi(v);

  if (const char *v = getenv("LIMGCU_MERGE_MODE")) ctx->mergeMode = !strcmp(v, "seq") ? 1 : 0;
  if (const char *v = getenv("LIMGCU_PASS1_TMA")) ctx->pass1Tma = atoi(v);
  if (const char *v = getenv("LIMGCU_MERGE_SAFE")) ctx->mergeSafe = atoi(v);
  if (const char *v = getenv("LIMGCU_PLAN_BANDS")) ctx->planBands = atoi(v) < 1 ? 1 : (atoi(v) > 64 ? 64 : atoi(v));
  if (const char *v = getenv("LIMGCU_SCAN_WARPS")) ctx->scanWarps = atoi(v) < 1 ? 1 : (atoi(v) > LIMG_CTA_WARPS ? LIMG_CTA_WARPS : atoi(v));
  if (const char *v = getenv("LIMGCU_SCAN_EXPERIMENT")) ctx->scanExperiment = atoi(v);
  if (const char *v = getenv("LIMGCU_SCAN_CLUSTER")) { ctx->scanCluster = atoi(v) < 0 ? 0 : (atoi(v) > 16 ? 16 : atoi(v)); ctx->scanClusterSet = true; }
  if (const char *v = getenv("LIMGCU_MERGE_MARGIN")) ctx->mergeMargin = atoi(v) < 0 ? 0 : atoi(v);
  if (const char *v = getenv("LIMGCU_PLAN_EXTW")) ctx->planExtW = atoi(v) < 8 ? 8 : (atoi(v) > 32 ? 32 : atoi(v));
  if (const char *v = getenv("LIMGCU_PLAN_SYML")) ctx->planSymL = atoi(v) < 1 ? 1 : (atoi(v) > 8 ? 8 : atoi(v));
  if (const char *v = getenv("LIMGCU_PLAN_SYMR")) ctx->planSymR = atoi(v) < 8 ? 8 : (atoi(v) > 24 ? 24 : atoi(v));
  if (const char *v = getenv("LIMGCU_PLAN_SYMD")) ctx->planSymD = atoi(v) < 8 ? 8 : (atoi(v) > 24 ? 24 : atoi(v));
  if (const char *v = getenv("LIMGCU_MERGE_ROWTIMES")) ctx->waveRowTimes = atoi(v);
  if (const char *v = getenv("LIMGCU_PLAN_ASYNC")) ctx->planAsync = atoi(v);
  if (const char *v = getenv("LIMGCU_PLAN_CTAS")) ctx->planAsyncCtas = atoi(v) < 1 ? 1 : (atoi(v) > 8 ? 8 : atoi(v));
  if (const char *v = getenv("LIMGCU_MERGE_GAP")) ctx->mergeGap = atoi(v) < 0 ? 0 : atoi(v);
  if (const char *v = getenv("LIMGCU_DECODE_VARIANT")) ctx->decodeVariant = atoi(v);
  if (const char *v = getenv("LIMGCU_DITHER")) ctx->ditherAes = !strcmp(v, "aes") ? 1 : 0;
  if (const char *v = getenv("LIMGCU_AES_SOFTWARE")) ctx->aesForceSoftware = atoi(v);
  if (const char *v = getenv("LIMGCU_MERGE_SPEC")) ctx->mergeSpec = atoi(v) < 0 ? 0 : atoi(v);

  for (auto &e : ctx->ev)
    if (cudaEventCreate(&e) != cudaSuccess) return bail(LIMGCU_ERROR_CUDA);

  {
    const int smemLarge = 4096 + LIMG_CTA_STAGE_PX * 16 + 2 * LIMG_CTA_AREA_CAP * 4;
    cudaFuncSetAttribute(k_encode_large<3, LIMG_ENCODE_THREADS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemLarge);
    cudaFuncSetAttribute(k_encode_large<4, LIMG_ENCODE_THREADS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemLarge);
    cudaFuncSetAttribute(k_encode_large<3, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemLarge);
    cudaFuncSetAttribute(k_encode_large<4, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemLarge);
    cudaDeviceGetAttribute(&ctx->scanSmemLimit, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    {
      // (the opt-in limit covers the kernel's static shared memory too)
      cudaFuncAttributes fa3 = {}, fa4 = {};
      cudaFuncGetAttributes(&fa3, k_merge_cta<3>);
      cudaFuncGetAttributes(&fa4, k_merge_cta<4>);
      const int staticBytes = (int)(fa3.sharedSizeBytes > fa4.sharedSizeBytes ? fa3.sharedSizeBytes : fa4.sharedSizeBytes);
      ctx->scanSmemLimit = ctx->scanSmemLimit > staticBytes ? ctx->scanSmemLimit - staticBytes : 0;

      if (cudaFuncSetAttribute(k_merge_cta<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->scanSmemLimit) != cudaSuccess ||
          cudaFuncSetAttribute(k_merge_cta<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->scanSmemLimit) != cudaSuccess)
      {
        cudaGetLastError();
        ctx->scanSmemLimit = 48 * 1024 - staticBytes; // what every kernel gets without opting in: larger frames use the scan over global memory
      }
    }
    cudaFuncSetAttribute(k_merge_cta<3>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); // LIMGCU_SCAN_CLUSTER=16
    cudaFuncSetAttribute(k_merge_cta<4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  }

  *out = ctx;
  return LIMGCU_SUCCESS;
}

extern "C" void limgcu_destroy(limgcu_ctx *ctx)
{
  if (ctx == nullptr)
    return;

  cudaSetDevice(ctx->device);

  if (ctx->stream)
    cudaStreamSynchronize(ctx->stream);

  void *ptrs[] = { ctx->dLut, ctx->dTable, ctx->dRec, ctx->dWindow, ctx->dAreas, ctx->dBlockToArea, ctx->dWork, ctx->dSmallList, ctx->dLargeList, ctx->dDemand,
                   ctx->dUsed, ctx->dScratchPx, ctx->dScratchFac, ctx->dCounters, ctx->dCompare, ctx->dSrc,
                   ctx->dExtSlot, ctx->dExtSeed, ctx->dExtBits, ctx->dExtHdr, ctx->dPlanCounters, ctx->dSymSlot, ctx->dSymSeed, ctx->dSymBits, ctx->dSymHdr, ctx->dSymStart, ctx->dSeedSym, ctx->dUnmasked, ctx->dLeftRun, ctx->dSafe,
                   ctx->dWaveZero, ctx->dTau, ctx->dCandList, ctx->dRowLists, ctx->dWaveDbg, ctx->dWaveRows, ctx->dRowMeta, ctx->dReplayList, ctx->dReplayCount, ctx->dPayload, ctx->dPayloadOff, ctx->dDitherBefore };

  for (void *p : ptrs)
    if (p) cudaFree(p);

  void *aesDev[] = { ctx->dNoise, ctx->dNoiseOff, ctx->dStates };
  void *aesHost[] = { ctx->hAreas, ctx->hNoise, ctx->hNoiseOff, ctx->hStates };
  for (void *p : aesDev) if (p) cudaFree(p);
  for (void *p : aesHost) if (p) cudaFreeHost(p);

  for (auto p : ctx->dPlaneU32) if (p) cudaFree(p);
  for (auto p : ctx->dPlaneU8) if (p) cudaFree(p);
  for (auto e : ctx->ev) if (e) cudaEventDestroy(e);

  if (ctx->stream)
    cudaStreamDestroy(ctx->stream);

  if (ctx->streamAux)
  {
    cudaStreamSynchronize(ctx->streamAux);
    cudaStreamDestroy(ctx->streamAux);
  }

  if (ctx->evFork) cudaEventDestroy(ctx->evFork);
  if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
  if (ctx->evFork2) cudaEventDestroy(ctx->evFork2);
  if (ctx->evJoin2) cudaEventDestroy(ctx->evJoin2);
  for (auto e : ctx->evBand) if (e) cudaEventDestroy(e);

  delete ctx;
}

extern "C" const char *limgcu_last_error(const limgcu_ctx *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" void *limgcu_stream_handle(limgcu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" uint64_t limgcu_launch_count(const limgcu_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int limgcu_debug_counters(limgcu_ctx *ctx, uint32_t *out32)
{
  NEED(ctx); NEED(out32);
  CK(cudaMemcpyAsync(out32, ctx->dCounters, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->dPlanCounters) CK(cudaMemcpyAsync(out32 + 29, ctx->dPlanCounters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_debug_wave(limgcu_ctx *ctx, uint32_t *out256)
{
  NEED(ctx); NEED(out256);
  CK(cudaMemcpyAsync(out256, ctx->dWaveDbg, 256 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_debug_wave_rows(limgcu_ctx *ctx, uint32_t *out, size_t blockY)
{
  NEED(ctx); NEED(out);

  if (ctx->dWaveRows == nullptr || blockY * 8 + 512 > ctx->capWaveRows)
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "no row time stamps were recorded (LIMGCU_MERGE_ROWTIMES=1)", cudaSuccess);

  CK(cudaMemcpyAsync(out, ctx->dWaveRows, (blockY * 8 + 512) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_sync(limgcu_ctx *ctx)
{
  NEED(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_set_rsqrt_lut(limgcu_ctx *ctx, const uint16_t *lut2048)
{
  NEED(ctx);
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->dLut, lut2048 ? lut2048 : LIMG_RSQRT_LUT, 2048 * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_enable_phase_timing(limgcu_ctx *ctx, int enable)
{
  NEED(ctx);
  ctx->timing = enable != 0;
  return LIMGCU_SUCCESS;
}

extern "C" float limgcu_phase_ms(limgcu_ctx *ctx, int phase)
{
  if (ctx == nullptr || phase < 0 || phase >= PHASE_COUNT)
    return -1.0f;

  return ctx->phaseMs[phase];
}

// ---------------------------------------------------------------------------------------------------------------
// stages
// ---------------------------------------------------------------------------------------------------------------

static int check_image(limgcu_ctx *ctx, size_t W, size_t H)
{
  if (W == 0 || H == 0 || W > 65528 || H > 65528)
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "image size out of range (1..65528)", cudaSuccess);

  return LIMGCU_SUCCESS;
}

// 2-D tensor map over the source (uint32 pixels, row pitch sizeX * 4 bytes) with 8 x 8 pixel boxes: what k_pass1_tma's tile loads address.
// The encoder is a driver entry point, looked up once (no link dependency on libcuda).
static bool make_source_map(const uint32_t *dSrc, size_t W, size_t H, CUtensorMap *map)
{
  typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  static bool looked = false;

  if (!looked)
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;

    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeTiled>(fn);
    else
      cudaGetLastError();

    looked = true;
  }

  if (encode == nullptr || (W % 4) != 0 || (reinterpret_cast<uintptr_t>(dSrc) & 15) != 0)
    return false;

  const cuuint64_t dims[2] = { (cuuint64_t)W, (cuuint64_t)H }, strides[1] = { (cuuint64_t)W * sizeof(uint32_t) };
  const cuuint32_t box[2] = { LIMG_BLOCK, LIMG_BLOCK }, elem[2] = { 1, 1 };
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t *>(dSrc), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int launch_pass1(limgcu_ctx *ctx, const uint32_t *dSrc, size_t W, size_t H, int hasAlpha, limgcu_decomp *dTable)
{
  const int BX = (int)((W + 7) / 8), BY = (int)((H + 7) / 8);
  CUtensorMap map;

  // pixels staged by the TMA (one 8 x 8 tile load per block, two in flight per warp) where the row pitch allows it; LIMGCU_PASS1_TMA=0: the plain loads
  if (ctx->pass1Tma && make_source_map(dSrc, W, H, &map))
  {
    const int gridT = (BX * BY + 7) / 8 < ctx->smCount * 8 ? (BX * BY + 7) / 8 : ctx->smCount * 8;

    if (hasAlpha)
      k_pass1_tma<4><<<gridT, 256, 0, ctx->stream>>>(map, (int)W, (int)H, BX, BY, ctx->dLut, dTable);
    else
      k_pass1_tma<3><<<gridT, 256, 0, ctx->stream>>>(map, (int)W, (int)H, BX, BY, ctx->dLut, dTable);

    CKL("k_pass1_tma");
    return LIMGCU_SUCCESS;
  }

  const int grid = (BX * BY + 7) / 8 < ctx->smCount * 8 ? (BX * BY + 7) / 8 : ctx->smCount * 8;

  if (hasAlpha)
    k_pass1<4><<<grid, 256, 0, ctx->stream>>>(dSrc, (int)W, (int)H, BX, BY, ctx->dLut, dTable);
  else
    k_pass1<3><<<grid, 256, 0, ctx->stream>>>(dSrc, (int)W, (int)H, BX, BY, ctx->dLut, dTable);

  CKL("k_pass1");
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_pass1(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, limgcu_decomp *d_table)
{
  NEED(ctx); NEED(d_src); NEED(d_table);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  return launch_pass1(ctx, d_src, sizeX, sizeY, hasAlpha, d_table);
}

// pass-1 table -> area rectangles (+ leftovers, geometry, block map, size classes). dAreas / dBlockToArea may be the context's own.
static int launch_merge(limgcu_ctx *ctx, const limgcu_decomp *dTable, size_t W, size_t H, int hasAlpha, limgcu_area *dAreas, uint32_t *dBlockToArea, bool noMerge)
{
  const int BX = (int)((W + 7) / 8), BY = (int)((H + 7) / 8);
  const int blocks = BX * BY;
  const int wordsPerRow = (BX + 31) / 32 + 2;

  CK(cudaMemsetAsync(ctx->dCounters, 0, 32 * sizeof(uint32_t), ctx->stream));

  if (ctx->timing) CK(cudaEventRecord(ctx->ev[PHASE_WINDOW], ctx->stream));

  if (!noMerge)
  {
    if (BX > 8190 || BY > 8190)
      return fail(ctx, LIMGCU_ERROR_OUT_OF_BOUNDS, "image too large for the merge's time stamps", cudaSuccess);

    const size_t usedWords = (size_t)BY * wordsPerRow;
    uint32_t *wz = ctx->dWaveZero;
    uint32_t *wFlags = wz, *wCandCount = wz + 10, *wTicket = wz + 12;
    int *wProgress = reinterpret_cast<int *>(wz + 16);
    uint32_t *wRowCounts = wz + 16 + 2 * BY, *wCandBits = wz + 16 + 4 * BY, *wEmitInfo = wCandBits + 2 * usedWords;
    CK(cudaMemsetAsync(wz, 0, (16 + 4 * (size_t)BY + 2 * usedWords + 2 * (size_t)blocks) * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->dPlanCounters, 0, 8 * sizeof(uint32_t), ctx->stream));

    PlanArgs pl;
    pl.rec = ctx->dRec; pl.window = ctx->dWindow; pl.BX = BX; pl.BY = BY; pl.wordsPerRow = wordsPerRow;
    pl.extSlot = ctx->dExtSlot; pl.extSeed = ctx->dExtSeed; pl.extBits = ctx->dExtBits; pl.extHdr = ctx->dExtHdr;
    pl.symSlot = ctx->dSymSlot; pl.symSeed = ctx->dSymSeed; pl.symBits = ctx->dSymBits; pl.symHdr = ctx->dSymHdr; pl.symStart = ctx->dSymStart;
    pl.counters = ctx->dPlanCounters; pl.extCap = ctx->mergeExt ? ctx->extCap : 0; pl.symCap = ctx->mergeExt ? ctx->symCap : 0; pl.unmasked = ctx->dUnmasked;
    pl.candBits = wCandBits; pl.candList = ctx->dCandList; pl.candCount = wCandCount;
    pl.extMaxW = ctx->planExtW; pl.symMaxL = ctx->planSymL; pl.symMaxR = ctx->planSymR; pl.symMaxD = ctx->planSymD;
    // The bitmaps beyond the 8x8 windows are pure accelerators of the scan and appear slot by slot, so their kernels run on a second,
    // low-priority stream WHILE the scan already walks the image top-down on the main stream; a seed whose bitmap is not there yet
    // is grown with on-demand predicates. Two plan CTAs per SM leave room for a scan CTA.
    // planAsync: 0 everything on the main stream, 1 both bitmap kernels on the second stream, 2 the stage-0 extension bitmaps on the main
    // stream (the first rows of the scan need them at once) and only the regrowth-centre bitmaps on the second stream.
    const bool async = ctx->planAsync != 0;
    const bool extendAsync = ctx->planAsync == 1;
    const int planGrid = ctx->smCount * (async ? ctx->planAsyncCtas : 6);
    const int extendGrid = ctx->smCount * (extendAsync ? ctx->planAsyncCtas : 6);
    cudaStream_t planStream = async ? ctx->streamAux : ctx->stream;
    cudaStream_t extendStream = extendAsync ? ctx->streamAux : ctx->stream;

    if (hasAlpha)
    {
      k_pred_records<4><<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(dTable, blocks, ctx->dRec);
      CKL("k_pred_records");

      if (ctx->mergeSafe)
      {
        // (one thread per block and a chain of predicates each: latency bound, so it runs beside the window kernel on the second stream)
        CK(cudaEventRecord(ctx->evFork2, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->streamAux, ctx->evFork2, 0));
        k_pred_leftrun<4><<<(blocks + 127) / 128, 128, 0, ctx->streamAux>>>(ctx->dRec, BX, BY, ctx->dLeftRun);
        CKL("k_pred_leftrun");
        CK(cudaEventRecord(ctx->evJoin2, ctx->streamAux));
      }

      k_pred_window<4><<<(blocks * 2 + 7) / 8, 256, 0, ctx->stream>>>(ctx->dRec, BX, BY, ctx->dWindow);
      CKL("k_pred_window");
    }
    else
    {
      k_pred_records<3><<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(dTable, blocks, ctx->dRec);
      CKL("k_pred_records");

      if (ctx->mergeSafe)
      {
        // (one thread per block and a chain of predicates each: latency bound, so it runs beside the window kernel on the second stream)
        CK(cudaEventRecord(ctx->evFork2, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->streamAux, ctx->evFork2, 0));
        k_pred_leftrun<3><<<(blocks + 127) / 128, 128, 0, ctx->streamAux>>>(ctx->dRec, BX, BY, ctx->dLeftRun);
        CKL("k_pred_leftrun");
        CK(cudaEventRecord(ctx->evJoin2, ctx->streamAux));
      }

      k_pred_window<3><<<(blocks * 2 + 7) / 8, 256, 0, ctx->stream>>>(ctx->dRec, BX, BY, ctx->dWindow);
      CKL("k_pred_window");
    }

    k_plan_seeds<<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(pl);
    CKL("k_plan_seeds");

    // (the bitmap kernels need nothing after this: they start beside the safe columns, the scan behind them)
    if (async)
    {
      CK(cudaEventRecord(ctx->evFork, ctx->stream));
      CK(cudaStreamWaitEvent(ctx->streamAux, ctx->evFork, 0));
    }

    if (ctx->mergeSafe)
    {
      CK(cudaStreamWaitEvent(ctx->stream, ctx->evJoin2, 0)); // the left runs
      k_plan_safe<<<BY, 32, 0, ctx->stream>>>(ctx->dWindow, wCandBits, ctx->dLeftRun, BX, BY, wordsPerRow, ctx->dSafe);
      CKL("k_plan_safe");
    }

    // The bitmaps are built TOP-DOWN in bands of block rows (LIMGCU_PLAN_BANDS, default 8), per band: the centres of the candidates whose mask-free growth stays
    // inside their 8x8 word are requested (85 % on photo content: their rectangle, hence their predicted centre, is final after k_plan_seeds), the extension bitmaps
    // are built, the centres of the candidates that needed one are requested, and one k_plan_sym launch builds all centre bitmaps requested since the last one.
    // The scan spends its first millisecond in the top rows (block row 0 runs on its own, every row below follows it 10 us behind the one above), and a seed whose
    // bitmap is not there yet grows with on-demand predicates and builds its centre bitmap itself (20 us on the row's chain): with the three kernels over the
    // whole image one after the other, the centre bitmap of a top-row seed that needed an extension appeared 0.5 ms after the scan's start (4K: 5.11 -> 4.79 ms,
    // 8K RGBA 16.8 -> 14.9, profiles/README.md r2_x). The bitmaps are built 2.7 x faster than the scan consumes rows, so top-down order keeps them ahead of it.
    // (planAsync == 2: the extension bitmaps run on the main stream, in front of the scan: one band)
    const int planBands = (extendAsync && BY >= 16 * ctx->planBands) ? ctx->planBands : 1; // (a band is four launches: not worth it below ~16 block rows per band)

    for (int band = 0; band < planBands; band++)
    {
      const int rowLo = (int)((long long)BY * band / planBands), rowHi = (int)((long long)BY * (band + 1) / planBands);
      k_plan_centres<<<(blocks + 255) / 256, 256, 0, planStream>>>(pl, 0, rowLo, rowHi);
      CKL("k_plan_centres");

      if (hasAlpha)
        k_plan_extend<4><<<extendGrid, LIMG_PLAN_WARPS * 32, 0, extendStream>>>(pl, rowLo, rowHi);
      else
        k_plan_extend<3><<<extendGrid, LIMG_PLAN_WARPS * 32, 0, extendStream>>>(pl, rowLo, rowHi);

      CKL("k_plan_extend");

      if (async && !extendAsync)
      {
        CK(cudaEventRecord(ctx->evFork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->streamAux, ctx->evFork, 0));
      }

      k_plan_centres<<<(blocks + 255) / 256, 256, 0, planStream>>>(pl, 1, rowLo, rowHi);
      CKL("k_plan_centres");

      if (hasAlpha)
        k_plan_sym<4><<<planGrid, LIMG_PLAN_WARPS * 32, 0, planStream>>>(pl);
      else
        k_plan_sym<3><<<planGrid, LIMG_PLAN_WARPS * 32, 0, planStream>>>(pl);

      CKL("k_plan_sym");
      k_plan_mark<<<1, 1, 0, planStream>>>(pl);
      CKL("k_plan_mark");
    }

    if (async)
    {
      CK(cudaEventRecord(ctx->evJoin, ctx->streamAux));
    }
    else
    {
      k_plan_link<<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(pl, ctx->dSeedSym);
      CKL("k_plan_link");
    }

    if (ctx->timing) CK(cudaEventRecord(ctx->ev[PHASE_SCAN], ctx->stream));

    CK(cudaMemsetAsync(ctx->dUsed, 0, usedWords * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->dTau, 0xFF, (size_t)blocks * sizeof(uint32_t), ctx->stream));

    WaveArgs w;
    w.rec = ctx->dRec; w.window = ctx->dWindow; w.extSlot = ctx->dExtSlot; w.extBits = ctx->dExtBits; w.extHdr = ctx->dExtHdr;
    w.symSlot = ctx->dSymSlot; w.symBits = ctx->dSymBits; w.symHdr = ctx->dSymHdr; w.unmasked = ctx->dUnmasked; w.seedSym = async ? nullptr : ctx->dSeedSym; w.safe = ctx->mergeSafe ? ctx->dSafe : nullptr;
    w.candBits = wCandBits; w.candList = ctx->dCandList; w.candCount = wCandCount;
    w.BX = BX; w.BY = BY; w.wordsPerRow = wordsPerRow;
    w.used = ctx->dUsed; w.tau = ctx->dTau; w.progress = wProgress; w.ticket = wTicket;
    w.rowLists = ctx->dRowLists; w.rowCounts = wRowCounts; w.emitInfo = wEmitInfo; w.flags = wFlags; w.stats = ctx->dCounters + 8;
    w.listCap = 2 * BX; w.margin = ctx->mergeMargin; w.stageGap = ctx->mergeGap; w.symMaxL = ctx->planSymL; w.symMaxR = ctx->planSymR; w.symMaxD = ctx->planSymD; w.specAhead = ctx->mergeSpec; w.dbg = ctx->dWaveDbg; w.experiment = ctx->scanExperiment;
    CK(cudaMemsetAsync(ctx->dWaveDbg, 0, 256 * sizeof(uint32_t), ctx->stream));
    w.dbgRows = nullptr;
    w.eventRow = ctx->waveRowTimes > 1 ? ctx->waveRowTimes : 0;

    if (ctx->waveRowTimes)
    {
      if ((size_t)BY * 8 + 512 > ctx->capWaveRows)
      {
        CK(regrow(ctx->dWaveRows, (size_t)BY * 8 + 512));
        ctx->capWaveRows = (size_t)BY * 8 + 512;
      }

      CK(cudaMemsetAsync(ctx->dWaveRows, 0, ((size_t)BY * 8 + 512) * sizeof(uint32_t), ctx->stream));
      w.dbgRows = ctx->dWaveRows;
    }

    // Tries: 0 pipelined rows with the normal margin, 1 pipelined with a wide margin (flat content, where a regrowth can reach far
    // to the left), 2 rows strictly in sequence (the reference's order, always exact). A try runs only if the one before failed its
    // verification; on ordinary content only try 0 does any work.
    const int rowsTotal = 2 * BY;
    const int waveGrid = (rowsTotal + LIMG_WAVE_WARPS - 1) / LIMG_WAVE_WARPS < ctx->smCount ? (rowsTotal + LIMG_WAVE_WARPS - 1) / LIMG_WAVE_WARPS : ctx->smCount;
    const int resetGrid = (blocks + 255) / 256 < ctx->smCount * 8 ? (blocks + 255) / 256 : ctx->smCount * 8;

    if (ctx->mergeMode == 1)
    {
      k_merge_set_tries<<<1, 1, 0, ctx->stream>>>(w, 2);
      CKL("k_merge_set_tries");
    }

    for (int attempt = 0; attempt < 3; attempt++)
    {
      WaveArgs wa = w;
      const int sequential = attempt == 2 ? 1 : 0;

      if (attempt == 1)
      {
        wa.margin = ctx->mergeWideMargin;
        wa.stageGap = ctx->mergeWideMargin;
      }

      if (attempt > 0)
      {
        k_merge_reset<<<resetGrid, 256, 0, ctx->stream>>>(wa, attempt);
        CKL("k_merge_reset");
      }

      // The pipelined tries run as ONE thread-block cluster with the mask, the progress words and the tickets replicated in the shared
      // memory of its CTAs (k_merge_cta), if a replica fits; the scan over global memory is for larger images and for the sequential try.
      const size_t ctaSmem = merge_cta_smem_words(BY, wordsPerRow) * sizeof(uint32_t);

      // 8 CTAs x 8 warps = 64 block rows in flight cover the wavefront of a 4K frame (BX / lag rows) and leave the rest of the GPU to other frames'
      // kernels; an 8K-class frame has twice the columns, so its wavefront is twice as deep: 16 CTAs (a non-portable cluster size, 10 % faster there).
      const int clusterSize = ctx->scanClusterSet ? ctx->scanCluster : (blocks >= 200000 ? 16 : 8);

      if (!sequential && clusterSize > 0 && ctaSmem <= (size_t)ctx->scanSmemLimit)
      {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        cfg.gridDim = dim3((unsigned)clusterSize);
        cfg.blockDim = dim3((unsigned)ctx->scanWarps * 32);
        cfg.dynamicSmemBytes = ctaSmem;
        cfg.stream = ctx->stream;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)clusterSize;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;

        if (hasAlpha)
          CK(cudaLaunchKernelEx(&cfg, k_merge_cta<4>, wa, attempt));
        else
          CK(cudaLaunchKernelEx(&cfg, k_merge_cta<3>, wa, attempt));

        CKL("k_merge_cta");
      }
      else
      {
        if (hasAlpha)
          k_merge_wave<4><<<sequential ? 1 : waveGrid, LIMG_WAVE_WARPS * 32, 0, ctx->stream>>>(wa, attempt, sequential);
        else
          k_merge_wave<3><<<sequential ? 1 : waveGrid, LIMG_WAVE_WARPS * 32, 0, ctx->stream>>>(wa, attempt, sequential);

        CKL("k_merge_wave");
      }

      if (async && attempt == 0)
        CK(cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0)); // everything after the first scan sees the complete bitmaps

      if (!sequential)
      {
        CK(cudaMemsetAsync(ctx->dReplayCount, 0, 2 * sizeof(uint32_t), ctx->stream));

        k_merge_verify_filter<<<dim3((blocks + 255) / 256, 2), 256, 0, ctx->stream>>>(wa, attempt, ctx->dReplayList, ctx->dReplayCount);
        CKL("k_merge_verify_filter");

        if (hasAlpha)
          k_merge_verify<4><<<dim3(ctx->smCount * 4, 2), 256, 0, ctx->stream>>>(wa, attempt, ctx->dReplayList, ctx->dReplayCount);
        else
          k_merge_verify<3><<<dim3(ctx->smCount * 4, 2), 256, 0, ctx->stream>>>(wa, attempt, ctx->dReplayList, ctx->dReplayCount);

        CKL("k_merge_verify");

        k_merge_judge<<<1, 1, 0, ctx->stream>>>(wa, attempt);
        CKL("k_merge_judge");
      }
    }

    CK(cudaMemcpyAsync(ctx->dCounters + 24, wFlags, 5 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dCounters + 31, wFlags + 5, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  else
  {
    // every block is its own area: nothing in use, nothing emitted
    const size_t usedWords = (size_t)BY * wordsPerRow;
    CK(cudaMemsetAsync(ctx->dUsed, 0, usedWords * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(ctx->dWaveZero + 16 + 2 * BY, 0, 2 * (size_t)BY * sizeof(uint32_t), ctx->stream));

    if (ctx->timing) CK(cudaEventRecord(ctx->ev[PHASE_SCAN], ctx->stream));
  }

  PrepareArgs p;
  p.W = (int)W; p.H = (int)H; p.BX = BX; p.BY = BY; p.wordsPerRow = wordsPerRow; p.listCap = 2 * BX;
  p.areas = dAreas; p.used = ctx->dUsed;
  p.rowCounts = ctx->dWaveZero + 16 + 2 * BY; p.rowLists = ctx->dRowLists;
  p.tau = noMerge ? nullptr : ctx->dTau; p.emitInfo = ctx->dWaveZero + 16 + 4 * BY + 2 * (size_t)BY * wordsPerRow;
  p.rowLeft = ctx->dRowMeta; p.rowBase = ctx->dRowMeta + BY;
  p.mergedCount = ctx->dCounters + 0; p.areaCount = ctx->dCounters + 1;
  p.blockToArea = dBlockToArea; p.work = ctx->dWork; p.smallList = ctx->dSmallList; p.largeList = ctx->dLargeList;
  p.smallCount = ctx->dCounters + 2; p.largeCount = ctx->dCounters + 3; p.hugeCount = ctx->dCounters + 7; p.bigCount = ctx->dCounters + 29; p.scratchTop = ctx->dCounters + 6;
  p.largeCap = (uint32_t)ctx->capBlocks;
  k_prepare_rowleft<<<BY, 32, 0, ctx->stream>>>(p);
  CKL("k_prepare_rowleft");
  k_prepare_collect<<<BY, 128, 0, ctx->stream>>>(p);
  CKL("k_prepare_collect");
  k_prepare_geometry<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(p);
  CKL("k_prepare_geometry");
  k_prepare_blockmap<<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(p);
  CKL("k_prepare_blockmap");
  return LIMGCU_SUCCESS;
}

// hard errors of the last merge scan (dCounters[27]: watchdog, [28]: a row list overflowed): the area map is truncated, nothing downstream is valid
static int scan_flags_error(limgcu_ctx *ctx, const uint32_t f[2])
{
  if (f[1])
    return fail(ctx, LIMGCU_ERROR_OUT_OF_BOUNDS, "a block row emitted more rectangles than its list holds", cudaSuccess);

  if (f[0])
    return fail(ctx, LIMGCU_ERROR_GENERIC, "merge watchdog: a block row waited too long for the rows above", cudaSuccess);

  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_status(limgcu_ctx *ctx)
{
  NEED(ctx);
  CK(cudaSetDevice(ctx->device));
  uint32_t f[2] = { 0, 0 };
  CK(cudaMemcpyAsync(f, ctx->dCounters + 27, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return scan_flags_error(ctx, f);
}

extern "C" int limgcu_merge(limgcu_ctx *ctx, const limgcu_decomp *d_table, size_t sizeX, size_t sizeY, int hasAlpha, limgcu_area *d_areas, uint32_t *d_area_count, uint32_t *d_block_to_area)
{
  NEED(ctx); NEED(d_table); NEED(d_areas);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  rc = ensure_capacity(ctx, sizeX, sizeY);
  if (rc) return rc;
  rc = launch_merge(ctx, d_table, sizeX, sizeY, hasAlpha, d_areas, d_block_to_area ? d_block_to_area : ctx->dBlockToArea, false);
  if (rc) return rc;

  if (d_area_count)
    CK(cudaMemcpyAsync(d_area_count, ctx->dCounters + 1, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));

  return LIMGCU_SUCCESS;
}

template <class T>
static cudaError_t regrow_host(T *&p, size_t count)
{
  if (p)
    cudaFreeHost(p);

  p = nullptr;
  return cudaMallocHost(reinterpret_cast<void **>(&p), count * sizeof(T));
}

static int ensure_aes(limgcu_ctx *ctx, size_t W, size_t H)
{
  const size_t blocks = ((W + 7) / 8) * ((H + 7) / 8), pixels = W * H;

  if (blocks > ctx->capAesBlocks)
  {
    ctx->capAesBlocks = 0;
    CK(regrow_host(ctx->hAreas, blocks));
    CK(regrow_host(ctx->hNoiseOff, 3 * blocks));
    CK(regrow_host(ctx->hStates, 2 * blocks));
    CK(regrow(ctx->dNoiseOff, 3 * blocks));
    CK(regrow(ctx->dStates, 2 * blocks));
    ctx->capAesBlocks = blocks;
  }

  if (pixels > ctx->capAesPixels)
  {
    ctx->capAesPixels = 0;
    CK(regrow_host(ctx->hNoise, 3 * pixels));
    CK(regrow(ctx->dNoise, 3 * pixels));
    ctx->capAesPixels = pixels;
  }

  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_set_dither_mode(limgcu_ctx *ctx, int aes)
{
  if (!ctx) return LIMGCU_ERROR_ARGUMENT_NULL;
  ctx->ditherAes = aes ? 1 : 0;
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_set_pool_threads(limgcu_ctx *ctx, int threads)
{
  NEED(ctx);
  ctx->poolThreads = threads < 0 ? 0 : threads;
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_host_has_aesni(void)
{
  return host_has_aesni() ? 1 : 0;
}

// per-area encode (k_encode_large on the main stream, k_encode_small next to it on the second one) of the areas whose first block row lies in
// [rowLo, rowHi); the lists and the work table are the ones launch_merge left in the context
static int launch_area_encode(limgcu_ctx *ctx, const uint32_t *d_src, int W, int H, int hasAlpha, uint32_t errorFactor, uint32_t flags, limgcu_area *dAreas, const limgcu_decomp *dTable,
                              uint32_t rowLo, uint32_t rowHi)
{
  const int BX = (W + 7) / 8;
  EncodeArgs e;
  e.src = d_src; e.W = W; e.H = H; e.BX = BX; e.BY = (H + 7) / 8; e.lut = ctx->dLut; e.table = dTable;
  e.areas = dAreas; e.areaCount = ctx->dCounters + 1; e.work = ctx->dWork; e.ditherDemand = ctx->dDemand;
  e.scratchPx = ctx->dScratchPx; e.scratchFac = ctx->dScratchFac;
  e.cp = make_crush_params(errorFactor, (flags & LIMGCU_FLAG_FAST_BIT_CRUSH) ? 1 : 0);
  e.rowLo = rowLo; e.rowHi = rowHi; e.hugeCount = nullptr; e.listCap = 0; e.bigCount = nullptr; e.bigList = nullptr;

  {
    EncodeArgs s = e;
    s.workCounter = ctx->dCounters + 4; s.list = ctx->dSmallList; s.listCount = ctx->dCounters + 2;
    EncodeArgs l = e;
    l.workCounter = ctx->dCounters + 5; l.list = ctx->dLargeList; l.listCount = ctx->dCounters + 3; l.hugeCount = ctx->dCounters + 7; l.listCap = (uint32_t)ctx->capBlocks; l.bigCount = ctx->dCounters + 29; l.bigList = ctx->dSmallList;
    const int gridLarge = ctx->smCount * 4, gridSmall = ctx->smCount * 6;
    const size_t smemLarge = 4096 + LIMG_CTA_STAGE_PX * 16 + 2 * LIMG_CTA_AREA_CAP * 4;

    // The large areas are the long poles (one CTA each, sequential sums): the warp-per-area kernel for the small ones runs next to them
    // on the second stream.
    CK(cudaEventRecord(ctx->evFork2, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->streamAux, ctx->evFork2, 0));

    EncodeArgs hg = l;
    hg.workCounter = ctx->dCounters + 30; // free until limgcu_finalize_rows (which clears it first)
    const int gridHuge = ctx->smCount * 2;

    // The CTA-sized areas on the second stream; on the main stream the huge areas (512 threads each) and, behind them, the warp-sized ones
    // (a handful of launches next to each other without a third stream: four contexts with two streams each are exactly the eight
    // hardware queues of the default CUDA_DEVICE_MAX_CONNECTIONS, a third stream per context cost the batch mode 17 %).
    if (hasAlpha)
    {
      k_encode_large<4, LIMG_ENCODE_THREADS, false><<<gridLarge, LIMG_ENCODE_THREADS, smemLarge, ctx->streamAux>>>(l);
      CKL("k_encode_large");
      k_encode_large<4, 512, true><<<gridHuge, 512, smemLarge, ctx->stream>>>(hg);
      CKL("k_encode_huge");
      k_encode_small<4><<<gridSmall, LIMG_ENCODE_THREADS, 0, ctx->stream>>>(s);
      CKL("k_encode_small");
    }
    else
    {
      k_encode_large<3, LIMG_ENCODE_THREADS, false><<<gridLarge, LIMG_ENCODE_THREADS, smemLarge, ctx->streamAux>>>(l);
      CKL("k_encode_large");
      k_encode_large<3, 512, true><<<gridHuge, 512, smemLarge, ctx->stream>>>(hg);
      CKL("k_encode_huge");
      k_encode_small<3><<<gridSmall, LIMG_ENCODE_THREADS, 0, ctx->stream>>>(s);
      CKL("k_encode_small");
    }

    CK(cudaEventRecord(ctx->evJoin2, ctx->streamAux));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->evJoin2, 0));
  }

  return LIMGCU_SUCCESS;
}

// dither chain states of every area (LCG: scan + jump-ahead on the device; AES: chain walked on the host), then the fused projection +
// dither + bit-crush + plane writer for the pixel rows [yLo, yHi)
static int launch_dither_finalize(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t flags, limgcu_area *dAreas, const uint32_t *dBlockToArea,
                                  const limgcu_stream *stream, const limgcu_planes *planes, size_t yLo, size_t yHi)
{
  const int W = (int)sizeX, H = (int)sizeY, BX = (W + 7) / 8;
  int rc = LIMGCU_SUCCESS;
  struct { int BY; } e = { (H + 7) / 8 };
  const bool ditherAes = (flags & LIMGCU_FLAG_DITHER_AES) != 0;

  // The non-merged encoder with a thread pool (limg_encode3d_test, limg.cpp:2108-2137) cuts the image into pool * 4 y-bands of whole block rows
  // (pool bands if that leaves less than a block row per band) and restarts the dither chain at the top of each (limg.cpp:1893): areas are the
  // blocks in raster order there, so a band is a fixed number of areas and the last band takes the rest.
  uint32_t bandAreas = 0, bandCount = 0;

  if ((flags & LIMGCU_FLAG_NO_MERGE) && ctx->poolThreads > 0)
  {
    size_t bands = (size_t)ctx->poolThreads * 4, rows = (sizeY / 8) / bands;

    if (rows == 0)
    {
      bands = (size_t)ctx->poolThreads;
      rows = (sizeY / 8) / bands;
    }

    if (rows > 0 && bands > 1)
    {
      bandAreas = (uint32_t)(rows * (size_t)BX);
      bandCount = (uint32_t)bands;
    }
  }

  if (ditherAes)
  {
    // The AES-round chain cannot be jumped: bring the area table (shifts, pixel rectangles) to the host, walk the chain there
    // (dither_aes_host.cpp), send the noise stream back. This synchronises with the stream.
    rc = ensure_aes(ctx, sizeX, sizeY);
    if (rc) return rc;
    uint32_t count = 0;
    CK(cudaMemcpyAsync(&count, ctx->dCounters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));

    if ((size_t)count > ctx->capAesBlocks)
      return fail(ctx, LIMGCU_ERROR_OUT_OF_BOUNDS, "more areas than blocks", cudaSuccess);

    CK(cudaMemcpyAsync(ctx->hAreas, dAreas, (size_t)count * sizeof(limgcu_area), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint64_t noiseBytes = aes_dither_chain_host(ctx->hAreas, count, LIMG_DITHER_SEED, ctx->hNoise, reinterpret_cast<uint64_t *>(ctx->hNoiseOff),
                                                      reinterpret_cast<uint64_t *>(ctx->hStates), reinterpret_cast<uint64_t *>(ctx->hStates) + count, ctx->aesForceSoftware != 0,
                                                      bandAreas, bandCount);
    CK(cudaMemcpyAsync(ctx->dNoise, ctx->hNoise, (size_t)noiseBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dNoiseOff, ctx->hNoiseOff, 3 * (size_t)count * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dStates, ctx->hStates, 2 * (size_t)count * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));

    if (count)
    {
      k_set_dither_states<<<(count + 255) / 256, 256, 0, ctx->stream>>>(dAreas, ctx->dStates, ctx->dStates + count, count);
      CKL("k_set_dither_states");
    }
  }
  else
  {
    k_dither_scan<<<1, 1024, 0, ctx->stream>>>(ctx->dCounters + 1, ctx->dDemand, ctx->dDitherBefore, 0);
    CKL("k_dither_scan");
    k_dither_states<<<(unsigned)((((size_t)BX * e.BY) + 255) / 256), 256, 0, ctx->stream>>>(dAreas, ctx->dCounters + 1, ctx->dDemand, ctx->dDitherBefore, ctx->jt, bandAreas, bandCount);
    CKL("k_dither_states");
  }

  if (ctx->timing) CK(cudaEventRecord(ctx->ev[PHASE_FINALIZE], ctx->stream));

  FinalizeArgs f;
  memset(&f, 0, sizeof(f));
  f.src = d_src; f.W = W; f.H = H; f.BX = BX; f.areas = dAreas; f.blockToArea = dBlockToArea;
  f.codesA = stream ? stream->codesA : nullptr; f.codesB = stream ? stream->codesB : nullptr; f.codesC = stream ? stream->codesC : nullptr;

  if (planes)
    f.planes = *planes;

  f.planes.pBlockError = nullptr; // never written (Q12)
  f.jt = ctx->jt;
  f.yLo = (int)yLo; f.yHi = (int)yHi;
  f.noise = ditherAes ? ctx->dNoise : nullptr;
  f.noiseOff = ditherAes ? ctx->dNoiseOff : nullptr;
  f.vec = (W % 8 == 0) && aligned32(d_src) && aligned32(f.codesA) && aligned32(f.codesB) && aligned32(f.codesC) && aligned32(f.planes.pDecoded) &&
          aligned32(f.planes.pFactorsA) && aligned32(f.planes.pFactorsB) && aligned32(f.planes.pFactorsC) && aligned32(f.planes.pBitsPerPixel) &&
          aligned32(f.planes.pShiftABCX) && aligned32(f.planes.pColAMin) && aligned32(f.planes.pColAMax) && aligned32(f.planes.pColBMin) &&
          aligned32(f.planes.pColBMax) && aligned32(f.planes.pColCMin) && aligned32(f.planes.pColCMax) && aligned32(f.planes.pBlockIndex);

  const bool anyOut = f.codesA || f.codesB || f.codesC || f.planes.pDecoded || f.planes.pFactorsA || f.planes.pFactorsB || f.planes.pFactorsC ||
                      f.planes.pBitsPerPixel || f.planes.pShiftABCX || f.planes.pColAMin || f.planes.pColAMax || f.planes.pColBMin || f.planes.pColBMax ||
                      f.planes.pColCMin || f.planes.pColCMax || f.planes.pBlockIndex;

  if (anyOut && yHi > yLo) // limg_encode3d_test_perf writes nothing (limg.cpp:2141-2173); an empty row band (sharded path) has no rows to write
  {
    const long long segs = (long long)((W + 7) / 8) * (long long)(yHi - yLo);
    const int grid = (int)((segs + 255) / 256);

    if (hasAlpha)
      k_finalize<4><<<grid, 256, 0, ctx->stream>>>(f);
    else
      k_finalize<3><<<grid, 256, 0, ctx->stream>>>(f);

    CKL("k_finalize");
  }

  (void)rc;
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_blocked_encode3d(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags,
                                       const limgcu_stream *stream, const limgcu_planes *planes)
{
  NEED(ctx); NEED(d_src);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  rc = ensure_capacity(ctx, sizeX, sizeY);
  if (rc) return rc;

  const int W = (int)sizeX, H = (int)sizeY, BX = (W + 7) / 8;
  limgcu_area *dAreas = (stream && stream->areas) ? stream->areas : ctx->dAreas;
  uint32_t *dBlockToArea = (stream && stream->block_to_area) ? stream->block_to_area : ctx->dBlockToArea;
  const bool noMerge = (flags & LIMGCU_FLAG_NO_MERGE) != 0;

  if (ctx->timing) CK(cudaEventRecord(ctx->ev[PHASE_PASS1], ctx->stream));

  const size_t BYall = (sizeY + 7) / 8;

  if (ctx->hostSrc != nullptr && BYall >= 64)
  {
    // host-buffer entry points: the source is uploaded in four bands of block rows on the second stream, pass 1 of a band starts as soon as
    // its rows are there (the fit of an 8x8 block needs nothing else), so only the last band's pass 1 is not hidden under the upload
    const uint32_t *hostSrc = ctx->hostSrc;
    ctx->hostSrc = nullptr;
    const size_t rowsPerBand = (BYall + 3) / 4;
    CK(cudaEventRecord(ctx->evFork2, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->streamAux, ctx->evFork2, 0));

    for (size_t b = 0; b < 4; b++)
    {
      const size_t by0 = b * rowsPerBand, by1 = by0 + rowsPerBand < BYall ? by0 + rowsPerBand : BYall;

      if (by0 >= by1)
        break;

      const size_t y0 = by0 * 8, y1 = by1 * 8 < sizeY ? by1 * 8 : sizeY;
      uint32_t *dBand = const_cast<uint32_t *>(d_src) + y0 * sizeX;
      CK(cudaMemcpyAsync(dBand, hostSrc + y0 * sizeX, (y1 - y0) * sizeX * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->streamAux));
      CK(cudaEventRecord(ctx->evBand[b], ctx->streamAux));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->evBand[b], 0));
      rc = launch_pass1(ctx, dBand, sizeX, y1 - y0, hasAlpha, ctx->dTable + by0 * (size_t)BX);
      if (rc) return rc;
    }
  }
  else
  {
    if (ctx->hostSrc != nullptr)
    {
      CK(cudaMemcpyAsync(const_cast<uint32_t *>(d_src), ctx->hostSrc, sizeX * sizeY * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
      ctx->hostSrc = nullptr;
    }

    rc = launch_pass1(ctx, d_src, sizeX, sizeY, hasAlpha, ctx->dTable);
    if (rc) return rc;
  }

  rc = launch_merge(ctx, ctx->dTable, sizeX, sizeY, hasAlpha, dAreas, dBlockToArea, noMerge);
  if (rc) return rc;

  if (ctx->timing) CK(cudaEventRecord(ctx->ev[PHASE_ENCODE], ctx->stream));

  rc = launch_area_encode(ctx, d_src, W, H, hasAlpha, errorFactor, flags, dAreas, ctx->dTable, 0u, 0xFFFFFFFFu);
  if (rc) return rc;

  if (ctx->timing) CK(cudaEventRecord(ctx->ev[PHASE_DITHER], ctx->stream));

  rc = launch_dither_finalize(ctx, d_src, sizeX, sizeY, hasAlpha, flags, dAreas, dBlockToArea, stream, planes, 0, sizeY);
  if (rc) return rc;

  if (stream && stream->area_count)
    CK(cudaMemcpyAsync(stream->area_count, ctx->dCounters + 1, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));

  if (ctx->timing)
  {
    CK(cudaEventRecord(ctx->ev[PHASE_COUNT], ctx->stream));
    CK(cudaEventSynchronize(ctx->ev[PHASE_COUNT]));

    for (int i = 0; i < PHASE_COUNT; i++)
      CK(cudaEventElapsedTime(&ctx->phaseMs[i], ctx->ev[i], ctx->ev[i + 1]));
  }

  return LIMGCU_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// phased encode for whole-image-exact row-band sharding: pass 1 per band (limgcu_pass1) -> all-gather of the
// table -> the identical scan on every rank (limgcu_merge) -> limgcu_encode_areas for the areas a rank owns -> SUM all-reduce of the
// per-area results -> limgcu_finalize_rows for the rank's pixel rows. The collectives are the caller's (limg_b200/shard.py uses NCCL).
// ---------------------------------------------------------------------------------------------------------------------------------

extern "C" size_t limgcu_area_result_words(void) { return LIMG_AREA_RESULT_WORDS; }

extern "C" int limgcu_encode_areas(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags, const limgcu_decomp *d_table,
                                   limgcu_area *d_areas, uint32_t rowLo, uint32_t rowHi, uint32_t *d_results)
{
  NEED(ctx); NEED(d_src); NEED(d_table); NEED(d_areas); NEED(d_results);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));

  if (ctx->capBlocks < ((sizeX + 7) / 8) * ((sizeY + 7) / 8))
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "limgcu_encode_areas: call limgcu_merge for this image first", cudaSuccess);

  rc = launch_area_encode(ctx, d_src, (int)sizeX, (int)sizeY, hasAlpha, errorFactor, flags, d_areas, d_table, rowLo, rowHi);
  if (rc) return rc;
  const size_t blocks = ((sizeX + 7) / 8) * ((sizeY + 7) / 8);
  k_pack_area_results<<<(unsigned)((blocks + 255) / 256), 256, 0, ctx->stream>>>(d_areas, ctx->dCounters + 1, ctx->dDemand, rowLo, rowHi, d_results);
  CKL("k_pack_area_results");
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_finalize_rows(limgcu_ctx *ctx, const uint32_t *d_src, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t flags, limgcu_area *d_areas, const uint32_t *d_results,
                                    const uint32_t *d_block_to_area, const limgcu_stream *stream, const limgcu_planes *planes, size_t yLo, size_t yHi)
{
  NEED(ctx); NEED(d_src); NEED(d_areas); NEED(d_results); NEED(d_block_to_area);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;

  if (yLo > yHi || yHi > sizeY)
    return fail(ctx, LIMGCU_ERROR_OUT_OF_BOUNDS, "limgcu_finalize_rows: row range", cudaSuccess);

  if (flags & LIMGCU_FLAG_DITHER_AES)
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "limgcu_finalize_rows: the AES dither chain is not available in the sharded path", cudaSuccess);

  CK(cudaSetDevice(ctx->device));
  const size_t blocks = ((sizeX + 7) / 8) * ((sizeY + 7) / 8);
  CK(cudaMemsetAsync(ctx->dCounters + 30, 0, sizeof(uint32_t), ctx->stream));
  k_unpack_area_results<<<(unsigned)((blocks + 255) / 256), 256, 0, ctx->stream>>>(d_areas, ctx->dCounters + 1, ctx->dDemand, d_results, ctx->dCounters + 30);
  CKL("k_unpack_area_results");
  rc = launch_dither_finalize(ctx, d_src, sizeX, sizeY, hasAlpha, flags, d_areas, d_block_to_area, stream, planes, yLo, yHi);
  if (rc) return rc;

  if (stream && stream->area_count)
    CK(cudaMemcpyAsync(stream->area_count, ctx->dCounters + 1, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));

  uint32_t bad = 0, scanFlags[2] = { 0, 0 };
  CK(cudaMemcpyAsync(&bad, ctx->dCounters + 30, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(scanFlags, ctx->dCounters + 27, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));

  if (int e = scan_flags_error(ctx, scanFlags))
    return e;

  if (bad)
    return fail(ctx, LIMGCU_ERROR_GENERIC, "limgcu_finalize_rows: some areas were encoded by no rank or by several (row ranges must partition the block rows)", cudaSuccess);

  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_build_block_map(limgcu_ctx *ctx, const limgcu_area *d_areas, uint32_t area_count, size_t sizeX, size_t sizeY, uint32_t *d_block_to_area)
{
  NEED(ctx); NEED(d_areas); NEED(d_block_to_area);
  CK(cudaSetDevice(ctx->device));

  // A block no rectangle covers maps to area 0, which exists: a table that does not tile the grid (the host entry points reject one,
  // area_table_valid()) decodes to garbage but never indexes outside the table.
  const size_t BX = (sizeX + 7) / 8, BY = (sizeY + 7) / 8;
  CK(cudaMemsetAsync(d_block_to_area, 0, BX * BY * sizeof(uint32_t), ctx->stream));

  if (area_count)
  {
    k_block_map<<<(area_count + 255) / 256, 256, 0, ctx->stream>>>(d_areas, area_count, (int)BX, (int)BY, d_block_to_area);
    CKL("k_block_map");
  }

  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_decode(limgcu_ctx *ctx, const limgcu_area *d_areas, const uint32_t *d_block_to_area, const uint8_t *d_codesA, const uint8_t *d_codesB, const uint8_t *d_codesC,
                             size_t sizeX, size_t sizeY, int hasAlpha, uint32_t *d_dst)
{
  NEED(ctx); NEED(d_areas); NEED(d_block_to_area); NEED(d_codesA); NEED(d_codesB); NEED(d_codesC); NEED(d_dst);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));

  const int W = (int)sizeX, H = (int)sizeY, BX = (W + 7) / 8;
  const long long segs = (long long)BX * ((H + 7) / 8) * 2; // one thread per half block (8 x 4 pixels)
  const int grid = (int)((segs + 255) / 256);
  const int vec = (W % 8 == 0) && aligned32(d_codesA) && aligned32(d_codesB) && aligned32(d_codesC) && aligned32(d_dst);

  const bool bulkOk = (W % 16 == 0) && (((uintptr_t)d_codesA | (uintptr_t)d_codesB | (uintptr_t)d_codesC | (uintptr_t)d_dst) & 15) == 0;

  if (vec && bulkOk && (ctx->decodeVariant & 64))
  {
    const int warps = ctx->decodeVariant & 15;
    const int tilesX = (BX + kDecodeTileBlocks - 1) / kDecodeTileBlocks, tiles = tilesX * ((H + 7) / 8);
    const size_t smem = (size_t)warps * (2 * kDecodeStageBytes + 16);
    const int perSM = (int)((227 * 1024) / (smem + 1024)) < 2048 / (warps * 32) ? (int)((227 * 1024) / (smem + 1024)) : 2048 / (warps * 32);
    const int g = (tiles + warps - 1) / warps < ctx->smCount * perSM ? (tiles + warps - 1) / warps : ctx->smCount * perSM;
#define LIMG_DECODE_CASE(WP) \
    if (warps == WP) \
    { \
      if (hasAlpha) \
      { \
        CK(cudaFuncSetAttribute(k_decode_stream<4, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_decode_stream<4, WP><<<g, WP * 32, smem, ctx->stream>>>(d_areas, d_block_to_area, d_codesA, d_codesB, d_codesC, W, H, BX, tilesX, tiles, d_dst); \
      } \
      else \
      { \
        CK(cudaFuncSetAttribute(k_decode_stream<3, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_decode_stream<3, WP><<<g, WP * 32, smem, ctx->stream>>>(d_areas, d_block_to_area, d_codesA, d_codesB, d_codesC, W, H, BX, tilesX, tiles, d_dst); \
      } \
    }
    LIMG_DECODE_CASE(1) LIMG_DECODE_CASE(2) LIMG_DECODE_CASE(4) LIMG_DECODE_CASE(8)
#undef LIMG_DECODE_CASE
  }
  else if (vec && (ctx->decodeVariant & 15) > 0)
  {
    const int rows = ctx->decodeVariant & 15, cta = (ctx->decodeVariant & 16) ? 128 : 256, cs = (ctx->decodeVariant & 32) ? 1 : 0;
    const long long threads = (long long)BX * ((H + 7) / 8) * (8 / rows);
    const int g = (int)((threads + cta - 1) / cta);
#define LIMG_DECODE_CASE(R, T, S) \
    if (rows == R && cta == T && cs == S) \
    { \
      if (hasAlpha) k_decode_tile<4, R, T, S != 0><<<g, T, 0, ctx->stream>>>(d_areas, d_block_to_area, d_codesA, d_codesB, d_codesC, W, H, BX, (uint32_t)threads, d_dst); \
      else k_decode_tile<3, R, T, S != 0><<<g, T, 0, ctx->stream>>>(d_areas, d_block_to_area, d_codesA, d_codesB, d_codesC, W, H, BX, (uint32_t)threads, d_dst); \
    }
    LIMG_DECODE_CASE(2, 256, 0) LIMG_DECODE_CASE(4, 256, 0) LIMG_DECODE_CASE(8, 256, 0)
    LIMG_DECODE_CASE(2, 128, 0) LIMG_DECODE_CASE(4, 128, 0)
    LIMG_DECODE_CASE(4, 256, 1) LIMG_DECODE_CASE(4, 128, 1)
#undef LIMG_DECODE_CASE
  }
  else if (hasAlpha)
    k_decode<4><<<grid, 256, 0, ctx->stream>>>(d_areas, d_block_to_area, d_codesA, d_codesB, d_codesC, W, H, BX, d_dst, vec);
  else
    k_decode<3><<<grid, 256, 0, ctx->stream>>>(d_areas, d_block_to_area, d_codesA, d_codesB, d_codesC, W, H, BX, d_dst, vec);

  CKL("k_decode");
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_debug_set_decode_variant(limgcu_ctx *ctx, int variant)
{
  if (!ctx) return LIMGCU_ERROR_ARGUMENT_NULL;
  static const int known[] = {0, 2, 4, 8, 18, 20, 36, 52, 65, 66, 68, 72};
  bool ok = false;
  for (int v : known) ok |= v == variant;
  if (!ok)
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "decode variant: 0, 2, 4, 8, 18, 20, 36, 52, 65, 66, 68, 72", cudaSuccess);
  ctx->decodeVariant = variant;
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_compare(limgcu_ctx *ctx, const uint32_t *d_a, const uint32_t *d_b, size_t sizeX, size_t sizeY, int hasAlpha, double *psnr, double *mse, double *maxError)
{
  NEED(ctx); NEED(d_a); NEED(d_b);
  CK(cudaSetDevice(ctx->device));
  const size_t n = sizeX * sizeY;
  CK(cudaMemsetAsync(ctx->dCompare, 0, sizeof(unsigned long long), ctx->stream));
  const int grid = (int)((n + 255) / 256 < (size_t)ctx->smCount * 8 ? (n + 255) / 256 : (size_t)ctx->smCount * 8);

  if (hasAlpha)
    k_compare<4><<<grid > 0 ? grid : 1, 256, 0, ctx->stream>>>(d_a, d_b, n, ctx->dCompare);
  else
    k_compare<3><<<grid > 0 ? grid : 1, 256, 0, ctx->stream>>>(d_a, d_b, n, ctx->dCompare);

  CKL("k_compare");
  unsigned long long total = 0;
  CK(cudaMemcpyAsync(&total, ctx->dCompare, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));

  const double maxE = hasAlpha ? 780300.0 : 585225.0; // limg_color_error(0x00000000, 0xFFFFFFFF)
  const double m = (double)total / (double)n;

  if (mse) *mse = m;
  if (maxError) *maxError = maxE;
  if (psnr) *psnr = 10.0 * log10(maxE / m);

  return LIMGCU_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------------------------

static int host_encode(limgcu_ctx *ctx, const uint32_t *pIn, size_t W, size_t H, int hasAlpha, const limgcu_planes *pInfo, uint32_t errorFactor, uint32_t flags,
                       limgcu_area *areas, uint32_t *areaCount, uint8_t *codesA, uint8_t *codesB, uint8_t *codesC)
{
  NEED(ctx); NEED(pIn);
  int rc = check_image(ctx, W, H);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t n = W * H;
  rc = ensure_staging(ctx, n);
  if (rc) return rc;
  rc = ensure_capacity(ctx, W, H);
  if (rc) return rc;

  ctx->hostSrc = pIn; // uploaded by limgcu_blocked_encode3d, in bands under pass 1

  limgcu_planes dev;
  memset(&dev, 0, sizeof(dev));
  uint32_t **devU32[9] = { &dev.pDecoded, &dev.pShiftABCX, &dev.pColAMin, &dev.pColAMax, &dev.pColBMin, &dev.pColBMax, &dev.pColCMin, &dev.pColCMax, &dev.pBlockIndex };
  uint32_t *hostU32[9] = { nullptr };
  uint8_t **devU8[4] = { &dev.pFactorsA, &dev.pFactorsB, &dev.pFactorsC, &dev.pBitsPerPixel };
  uint8_t *hostU8[4] = { nullptr };

  if (pInfo)
  {
    uint32_t *h32[9] = { pInfo->pDecoded, pInfo->pShiftABCX, pInfo->pColAMin, pInfo->pColAMax, pInfo->pColBMin, pInfo->pColBMax, pInfo->pColCMin, pInfo->pColCMax, pInfo->pBlockIndex };
    uint8_t *h8[4] = { pInfo->pFactorsA, pInfo->pFactorsB, pInfo->pFactorsC, pInfo->pBitsPerPixel };

    for (int i = 0; i < 9; i++) { hostU32[i] = h32[i]; if (h32[i]) *devU32[i] = ctx->dPlaneU32[i]; }
    for (int i = 0; i < 4; i++) { hostU8[i] = h8[i]; if (h8[i]) *devU8[i] = ctx->dPlaneU8[i]; }
  }

  limgcu_stream st;
  memset(&st, 0, sizeof(st));
  st.codesA = codesA ? ctx->dPlaneU8[4] : nullptr;
  st.codesB = codesB ? ctx->dPlaneU8[5] : nullptr;
  st.codesC = codesC ? ctx->dPlaneU8[6] : nullptr;

  rc = limgcu_blocked_encode3d(ctx, ctx->dSrc, W, H, hasAlpha, errorFactor, flags, &st, &dev);
  ctx->hostSrc = nullptr;
  if (rc) return rc;

  for (int i = 0; i < 9; i++)
    if (hostU32[i]) CK(cudaMemcpyAsync(hostU32[i], ctx->dPlaneU32[i], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));

  for (int i = 0; i < 4; i++)
    if (hostU8[i]) CK(cudaMemcpyAsync(hostU8[i], ctx->dPlaneU8[i], n, cudaMemcpyDeviceToHost, ctx->stream));

  if (codesA) CK(cudaMemcpyAsync(codesA, ctx->dPlaneU8[4], n, cudaMemcpyDeviceToHost, ctx->stream));
  if (codesB) CK(cudaMemcpyAsync(codesB, ctx->dPlaneU8[5], n, cudaMemcpyDeviceToHost, ctx->stream));
  if (codesC) CK(cudaMemcpyAsync(codesC, ctx->dPlaneU8[6], n, cudaMemcpyDeviceToHost, ctx->stream));

  uint32_t count = 0, overflow[2] = { 0, 0 };
  CK(cudaMemcpyAsync(&count, ctx->dCounters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(overflow, ctx->dCounters + 27, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));

  if (int bad = scan_flags_error(ctx, overflow))
    return bad;

  if (areas)
  {
    CK(cudaMemcpyAsync(areas, ctx->dAreas, (size_t)count * sizeof(limgcu_area), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }

  if (areaCount)
    *areaCount = count;

  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_host_blocked_encode3d(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, const limgcu_planes *pInfo, uint32_t errorFactor, int fastBitCrushing)
{
  NEED(ctx);
  return host_encode(ctx, pIn, sizeX, sizeY, hasAlpha, pInfo, errorFactor, (fastBitCrushing ? LIMGCU_FLAG_FAST_BIT_CRUSH : 0) | (ctx->ditherAes ? LIMGCU_FLAG_DITHER_AES : 0), nullptr, nullptr,
                     nullptr, nullptr, nullptr);
}

extern "C" int limgcu_host_encode3d(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, const limgcu_planes *pInfo, uint32_t errorFactor, int fastBitCrushing)
{
  limgcu_planes p;
  memset(&p, 0, sizeof(p));

  if (pInfo)
  {
    p = *pInfo;
    p.pBitsPerPixel = nullptr; // limg_encode3d_info (limg.h:29-33) has no bits-per-pixel / block-index planes
    p.pBlockIndex = nullptr;
    p.pBlockError = nullptr;
  }

  NEED(ctx);
  return host_encode(ctx, pIn, sizeX, sizeY, hasAlpha, pInfo ? &p : nullptr, errorFactor,
                     (fastBitCrushing ? LIMGCU_FLAG_FAST_BIT_CRUSH : 0) | LIMGCU_FLAG_NO_MERGE | (ctx->ditherAes ? LIMGCU_FLAG_DITHER_AES : 0), nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int limgcu_host_encode_stream(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags,
                                         limgcu_area *areas, uint32_t *area_count, uint8_t *codesA, uint8_t *codesB, uint8_t *codesC, uint32_t *pDecoded)
{
  limgcu_planes p;
  memset(&p, 0, sizeof(p));
  p.pDecoded = pDecoded;
  return host_encode(ctx, pIn, sizeX, sizeY, hasAlpha, &p, errorFactor, flags, areas, area_count, codesA, codesB, codesC);
}

// Host-side check of an area table that comes from outside (a stream handed to the decoder, a container): the kernels index the block
// map, the code planes and the table itself with what it says. Every rectangle lies inside the block grid, the rectangles cover every
// block exactly once, the shifts are at most 8 and the pixel rectangles are the ones the block rectangles imply.
static bool area_table_valid(const limgcu_area *areas, uint32_t count, size_t W, size_t H)
{
  const size_t BX = (W + 7) / 8, BY = (H + 7) / 8;

  if (count == 0 || (size_t)count > BX * BY)
    return false;

  std::vector<uint8_t> covered(BX * BY, 0);
  size_t coveredBlocks = 0;

  for (uint32_t k = 0; k < count; k++)
  {
    const limgcu_area &a = areas[k];

    if (a.rx == 0 || a.ry == 0 || (size_t)a.ox + a.rx > BX || (size_t)a.oy + a.ry > BY || a.shift[0] > 8 || a.shift[1] > 8 || a.shift[2] > 8)
      return false;

    const size_t pw = (size_t)a.rx * 8 < W - (size_t)a.ox * 8 ? (size_t)a.rx * 8 : W - (size_t)a.ox * 8;
    const size_t ph = (size_t)a.ry * 8 < H - (size_t)a.oy * 8 ? (size_t)a.ry * 8 : H - (size_t)a.oy * 8;

    if (a.px_x != a.ox * 8 || a.px_y != a.oy * 8 || a.px_w != pw || a.px_h != ph)
      return false;

    for (uint32_t y = a.oy; y < a.oy + a.ry; y++)
      for (uint32_t x = a.ox; x < a.ox + a.rx; x++)
      {
        if (covered[(size_t)y * BX + x])
          return false;

        covered[(size_t)y * BX + x] = 1;
        coveredBlocks++;
      }
  }

  return coveredBlocks == BX * BY;
}

extern "C" int limgcu_host_decode(limgcu_ctx *ctx, const limgcu_area *areas, uint32_t area_count, const uint8_t *codesA, const uint8_t *codesB, const uint8_t *codesC,
                                  size_t sizeX, size_t sizeY, int hasAlpha, uint32_t *pOut)
{
  NEED(ctx); NEED(areas); NEED(codesA); NEED(codesB); NEED(codesC); NEED(pOut);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t n = sizeX * sizeY;

  if (!area_table_valid(areas, area_count, sizeX, sizeY))
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "limgcu_host_decode: the area table does not tile the image (or a shift is above 8)", cudaSuccess);

  rc = ensure_decode_staging(ctx, n);
  if (rc) return rc;
  rc = ensure_decode_capacity(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->dAreas, areas, (size_t)area_count * sizeof(limgcu_area), cudaMemcpyHostToDevice, ctx->stream));
  rc = limgcu_build_block_map(ctx, ctx->dAreas, area_count, sizeX, sizeY, ctx->dBlockToArea);
  if (rc) return rc;

  // Bands of block rows, alternating between the two streams: the code upload of band i + 1 runs while band i is reconstructed and
  // downloaded (uploads and downloads use different copy engines; with pinned host buffers the call is bound by the 4 B/px download).
  const size_t BX = (sizeX + 7) / 8, BY = (sizeY + 7) / 8;
  const size_t bands = BY >= 64 ? 4 : 1, rowsPerBand = (BY + bands - 1) / bands;
  CK(cudaEventRecord(ctx->evFork2, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->streamAux, ctx->evFork2, 0));
  cudaStream_t main = ctx->stream;

  for (size_t b = 0; b < bands; b++)
  {
    const size_t by0 = b * rowsPerBand, by1 = by0 + rowsPerBand < BY ? by0 + rowsPerBand : BY;

    if (by0 >= by1)
      break;

    const size_t y0 = by0 * 8, y1 = by1 * 8 < sizeY ? by1 * 8 : sizeY, off = y0 * sizeX, cnt = (y1 - y0) * sizeX;
    ctx->stream = (b & 1) ? ctx->streamAux : main; // limgcu_decode launches on ctx->stream
    cudaError_t e = cudaMemcpyAsync(ctx->dPlaneU8[4] + off, codesA + off, cnt, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->dPlaneU8[5] + off, codesB + off, cnt, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->dPlaneU8[6] + off, codesC + off, cnt, cudaMemcpyHostToDevice, ctx->stream);

    if (e == cudaSuccess)
      rc = limgcu_decode(ctx, ctx->dAreas, ctx->dBlockToArea + by0 * BX, ctx->dPlaneU8[4] + off, ctx->dPlaneU8[5] + off, ctx->dPlaneU8[6] + off, sizeX, y1 - y0, hasAlpha, ctx->dPlaneU32[0] + off);

    if (e == cudaSuccess && rc == LIMGCU_SUCCESS)
      e = cudaMemcpyAsync(pOut + off, ctx->dPlaneU32[0] + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);

    ctx->stream = main;

    if (e != cudaSuccess)
      return fail(ctx, LIMGCU_ERROR_CUDA, "banded decode", e);

    if (rc)
      return rc;
  }

  CK(cudaEventRecord(ctx->evJoin2, ctx->streamAux));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->evJoin2, 0));
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// .limg container (layout in include/limgcu.h and kernels_container.cuh)
// ---------------------------------------------------------------------------------------------------------------------------------

#pragma pack(push, 1)
struct ContainerHeader
{
  char magic[8]; // "LIMGB200"
  uint32_t version, flags, sizeX, sizeY, areaCount, recordBytes;
  uint64_t payloadBytes, reserved;
};

struct ContainerRecordHead
{
  uint16_t ox, oy, rx, ry; // 8x8-block units
  uint8_t shift[3], stage;
};
#pragma pack(pop)

static_assert(sizeof(ContainerHeader) == 48 && sizeof(ContainerRecordHead) == 12, "container layout");

static const char kContainerMagic[8] = { 'L', 'I', 'M', 'G', 'B', '2', '0', '0' };

static size_t container_record_bytes(int hasAlpha) { return sizeof(ContainerRecordHead) + 6 * (hasAlpha ? 4 : 3) * sizeof(int16_t); }
static size_t container_payload_bound(size_t W, size_t H) { return 3 * ((W + 7) / 8) * 8 * H; }

extern "C" size_t limgcu_container_bound(size_t sizeX, size_t sizeY, int hasAlpha)
{
  const size_t blocks = ((sizeX + 7) / 8) * ((sizeY + 7) / 8);
  return sizeof(ContainerHeader) + blocks * container_record_bytes(hasAlpha) + container_payload_bound(sizeX, sizeY);
}

// payloadBytes == 0: room for the largest payload an image of this size can have (the encoder does not know it in advance)
static int ensure_payload(limgcu_ctx *ctx, size_t W, size_t H, size_t payloadBytes = 0)
{
  const size_t bytes = (payloadBytes ? payloadBytes : container_payload_bound(W, H)) + 16, entries = ((W + 7) / 8) * ((H + 7) / 8) + 1;

  if (bytes > ctx->capPayload)
  {
    ctx->capPayload = 0;
    CK(regrow(ctx->dPayload, bytes));
    ctx->capPayload = bytes;
  }

  if (entries > ctx->capPayloadOff)
  {
    ctx->capPayloadOff = 0;
    CK(regrow(ctx->dPayloadOff, entries));
    ctx->capPayloadOff = entries;
  }

  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_pack_payload(limgcu_ctx *ctx, const limgcu_area *d_areas, const uint32_t *d_area_count, uint32_t area_count, const uint32_t *d_block_to_area, const uint8_t *d_codesA,
                                   const uint8_t *d_codesB, const uint8_t *d_codesC, size_t sizeX, size_t sizeY, int hasAlpha, uint8_t *d_payload, uint64_t *d_offsets)
{
  NEED(ctx); NEED(d_areas); NEED(d_block_to_area); NEED(d_codesA); NEED(d_codesB); NEED(d_codesC); NEED(d_payload); NEED(d_offsets);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const int W = (int)sizeX, H = (int)sizeY, BX = (W + 7) / 8;
  const long long segs = (long long)BX * H;
  const int vec = (W % 8 == 0) && ((((uintptr_t)d_codesA | (uintptr_t)d_codesB | (uintptr_t)d_codesC) & 7) == 0);
  k_payload_scan<<<1, 1024, 0, ctx->stream>>>(d_areas, d_area_count, area_count, hasAlpha, reinterpret_cast<unsigned long long *>(d_offsets));
  CKL("k_payload_scan");
  k_container_pack<<<(unsigned)((segs + 255) / 256), 256, 0, ctx->stream>>>(d_areas, d_block_to_area, reinterpret_cast<const unsigned long long *>(d_offsets), d_codesA, d_codesB, d_codesC, W, H,
                                                                             BX, hasAlpha, vec, d_payload);
  CKL("k_container_pack");
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_unpack_payload(limgcu_ctx *ctx, const limgcu_area *d_areas, uint32_t area_count, const uint32_t *d_block_to_area, const uint8_t *d_payload, uint64_t *d_offsets,
                                     size_t sizeX, size_t sizeY, int hasAlpha, uint8_t *d_codesA, uint8_t *d_codesB, uint8_t *d_codesC)
{
  NEED(ctx); NEED(d_areas); NEED(d_block_to_area); NEED(d_codesA); NEED(d_codesB); NEED(d_codesC); NEED(d_payload); NEED(d_offsets);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const int W = (int)sizeX, H = (int)sizeY, BX = (W + 7) / 8;
  const long long segs = (long long)BX * H;
  const int vec = (W % 8 == 0) && ((((uintptr_t)d_codesA | (uintptr_t)d_codesB | (uintptr_t)d_codesC) & 7) == 0);
  k_payload_scan<<<1, 1024, 0, ctx->stream>>>(d_areas, nullptr, area_count, hasAlpha, reinterpret_cast<unsigned long long *>(d_offsets));
  CKL("k_payload_scan");
  k_container_unpack<<<(unsigned)((segs + 255) / 256), 256, 0, ctx->stream>>>(d_areas, d_block_to_area, reinterpret_cast<const unsigned long long *>(d_offsets), d_payload, W, H, BX, hasAlpha, vec,
                                                                               d_codesA, d_codesB, d_codesC);
  CKL("k_container_unpack");
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_container_info(const void *data, size_t bytes, size_t *sizeX, size_t *sizeY, int *hasAlpha, uint32_t *areaCount, uint64_t *payloadBytes)
{
  if (data == nullptr)
    return LIMGCU_ERROR_ARGUMENT_NULL;

  ContainerHeader h;

  if (bytes < sizeof(h))
    return LIMGCU_ERROR_OUT_OF_BOUNDS;

  memcpy(&h, data, sizeof(h));

  if (memcmp(h.magic, kContainerMagic, 8) != 0 || h.version != 1 || h.sizeX == 0 || h.sizeY == 0 || h.sizeX > 65528 || h.sizeY > 65528)
    return LIMGCU_ERROR_INVALID_PARAMETER;

  const int alpha = (int)(h.flags & 1u);
  const uint64_t blocks = (uint64_t)((h.sizeX + 7) / 8) * ((h.sizeY + 7) / 8);

  if (h.recordBytes != container_record_bytes(alpha) || h.areaCount == 0 || h.areaCount > blocks || h.payloadBytes > container_payload_bound(h.sizeX, h.sizeY))
    return LIMGCU_ERROR_INVALID_PARAMETER;

  if ((uint64_t)bytes < sizeof(h) + (uint64_t)h.areaCount * h.recordBytes + h.payloadBytes)
    return LIMGCU_ERROR_OUT_OF_BOUNDS;

  if (sizeX) *sizeX = h.sizeX;
  if (sizeY) *sizeY = h.sizeY;
  if (hasAlpha) *hasAlpha = alpha;
  if (areaCount) *areaCount = h.areaCount;
  if (payloadBytes) *payloadBytes = h.payloadBytes;
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_host_encode_container(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, uint32_t flags, void *out, size_t capacity,
                                            size_t *written)
{
  NEED(ctx); NEED(pIn); NEED(out); NEED(written);
  *written = 0;
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t n = sizeX * sizeY;
  rc = ensure_staging(ctx, n);
  if (rc) return rc;
  rc = ensure_capacity(ctx, sizeX, sizeY);
  if (rc) return rc;
  rc = ensure_payload(ctx, sizeX, sizeY);
  if (rc) return rc;

  ctx->hostSrc = pIn; // uploaded by limgcu_blocked_encode3d, in bands under pass 1
  limgcu_stream st;
  memset(&st, 0, sizeof(st));
  st.codesA = ctx->dPlaneU8[4]; st.codesB = ctx->dPlaneU8[5]; st.codesC = ctx->dPlaneU8[6];
  rc = limgcu_blocked_encode3d(ctx, ctx->dSrc, sizeX, sizeY, hasAlpha, errorFactor, flags, &st, nullptr);
  ctx->hostSrc = nullptr;
  if (rc) return rc;
  rc = limgcu_pack_payload(ctx, ctx->dAreas, ctx->dCounters + 1, 0, ctx->dBlockToArea, st.codesA, st.codesB, st.codesC, sizeX, sizeY, hasAlpha, ctx->dPayload,
                           reinterpret_cast<uint64_t *>(ctx->dPayloadOff));
  if (rc) return rc;

  uint32_t count = 0, overflow[2] = { 0, 0 };
  CK(cudaMemcpyAsync(&count, ctx->dCounters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(overflow, ctx->dCounters + 27, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));

  if (int bad = scan_flags_error(ctx, overflow))
    return bad;

  unsigned long long payloadBytes = 0;
  CK(cudaMemcpyAsync(&payloadBytes, ctx->dPayloadOff + count, sizeof(payloadBytes), cudaMemcpyDeviceToHost, ctx->stream));
  limgcu_area *areas = static_cast<limgcu_area *>(malloc((size_t)count * sizeof(limgcu_area)));

  if (areas == nullptr)
    return fail(ctx, LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE, "host area table", cudaSuccess);

  cudaError_t e = cudaMemcpyAsync(areas, ctx->dAreas, (size_t)count * sizeof(limgcu_area), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);

  const size_t recordBytes = container_record_bytes(hasAlpha), ch = hasAlpha ? 4 : 3;
  const size_t total = sizeof(ContainerHeader) + (size_t)count * recordBytes + (size_t)payloadBytes;

  if (e != cudaSuccess || total > capacity)
  {
    free(areas);
    return e != cudaSuccess ? fail(ctx, LIMGCU_ERROR_CUDA, "area table D2H", e) : fail(ctx, LIMGCU_ERROR_OUT_OF_BOUNDS, "container buffer too small (limgcu_container_bound)", cudaSuccess);
  }

  uint8_t *o = static_cast<uint8_t *>(out);
  ContainerHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, kContainerMagic, 8);
  h.version = 1; h.flags = hasAlpha ? 1u : 0u; h.sizeX = (uint32_t)sizeX; h.sizeY = (uint32_t)sizeY; h.areaCount = count; h.recordBytes = (uint32_t)recordBytes; h.payloadBytes = payloadBytes;
  memcpy(o, &h, sizeof(h));
  o += sizeof(h);

  for (uint32_t k = 0; k < count; k++)
  {
    const limgcu_area &a = areas[k];
    ContainerRecordHead r = { (uint16_t)a.ox, (uint16_t)a.oy, (uint16_t)a.rx, (uint16_t)a.ry, { a.shift[0], a.shift[1], a.shift[2] }, (uint8_t)a.stage };
    memcpy(o, &r, sizeof(r));
    o += sizeof(r);
    const int16_t *fields[6] = { a.decomp.dirA_min, a.decomp.dirA_max, a.decomp.dirB_offset, a.decomp.dirB_mag, a.decomp.dirC_offset, a.decomp.dirC_mag };

    for (const int16_t *f : fields)
    {
      memcpy(o, f, ch * sizeof(int16_t));
      o += ch * sizeof(int16_t);
    }
  }

  free(areas);
  CK(cudaMemcpyAsync(o, ctx->dPayload, (size_t)payloadBytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *written = total;
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_host_decode_container(limgcu_ctx *ctx, const void *data, size_t bytes, uint32_t *pOut, size_t outPixels)
{
  NEED(ctx); NEED(data); NEED(pOut);
  size_t W = 0, H = 0;
  int hasAlpha = 0;
  uint32_t count = 0;
  uint64_t payloadBytes = 0;
  int rc = limgcu_container_info(data, bytes, &W, &H, &hasAlpha, &count, &payloadBytes);

  if (rc)
    return fail(ctx, rc, "not a valid LIMGB200 container", cudaSuccess);

  if (outPixels < W * H)
    return fail(ctx, LIMGCU_ERROR_OUT_OF_BOUNDS, "output buffer smaller than sizeX * sizeY", cudaSuccess);

  // Nothing is allocated from what the (unauthenticated) header claims before the table is known to be sound: records -> area table; the
  // rectangles must tile the block grid exactly and the payload must have the size the table implies.
  const size_t n = W * H, BX = (W + 7) / 8, BY = (H + 7) / 8;
  const size_t recordBytes = container_record_bytes(hasAlpha), ch = hasAlpha ? 4 : 3;
  const uint8_t *p = static_cast<const uint8_t *>(data) + sizeof(ContainerHeader);

  if ((size_t)count > BX * BY || bytes < sizeof(ContainerHeader) + (size_t)count * recordBytes + payloadBytes)
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "container: more areas than blocks, or shorter than its table and payload", cudaSuccess);

  limgcu_area *areas = static_cast<limgcu_area *>(calloc(count ? count : 1, sizeof(limgcu_area)));

  if (areas == nullptr)
    return fail(ctx, LIMGCU_ERROR_MEMORY_ALLOCATION_FAILURE, "container: area table", cudaSuccess);

  for (uint32_t k = 0; k < count; k++)
  {
    ContainerRecordHead r;
    memcpy(&r, p, sizeof(r));
    p += sizeof(r);
    limgcu_area &a = areas[k];
    a.ox = r.ox; a.oy = r.oy; a.rx = r.rx; a.ry = r.ry; a.stage = r.stage;
    a.px_x = a.ox * 8; a.px_y = a.oy * 8;
    // (garbage for a rectangle outside the grid, which area_table_valid() rejects below)
    a.px_w = (uint32_t)((size_t)a.rx * 8 < W - (size_t)a.px_x ? (size_t)a.rx * 8 : W - (size_t)a.px_x);
    a.px_h = (uint32_t)((size_t)a.ry * 8 < H - (size_t)a.px_y ? (size_t)a.ry * 8 : H - (size_t)a.px_y);
    memcpy(a.shift, r.shift, 3);
    int16_t *fields[6] = { a.decomp.dirA_min, a.decomp.dirA_max, a.decomp.dirB_offset, a.decomp.dirB_mag, a.decomp.dirC_offset, a.decomp.dirC_mag };

    for (int16_t *f : fields)
    {
      memcpy(f, p, ch * sizeof(int16_t));
      p += ch * sizeof(int16_t);
    }
  }

  bool ok = area_table_valid(areas, count, W, H);
  uint64_t expectPayload = 0;

  for (uint32_t k = 0; ok && k < count; k++)
  {
    const limgcu_area &a = areas[k];
    expectPayload += (uint64_t)((a.px_w + 7) / 8) * a.px_h *
                     (container_code_bits(a.shift[0], hasAlpha != 0) + container_code_bits(a.shift[1], hasAlpha != 0) + container_code_bits(a.shift[2], hasAlpha != 0));
  }

  ok = ok && expectPayload == payloadBytes;

  if (!ok)
  {
    free(areas);
    return fail(ctx, LIMGCU_ERROR_INVALID_PARAMETER, "container: the area table does not tile the image or does not match the payload size", cudaSuccess);
  }

  // only what the decode needs: three code planes, the output, the area table, the block map, the payload
  CK(cudaSetDevice(ctx->device));
  rc = ensure_decode_staging(ctx, n);
  if (rc == LIMGCU_SUCCESS) rc = ensure_decode_capacity(ctx, W, H);
  if (rc == LIMGCU_SUCCESS) rc = ensure_payload(ctx, W, H, (size_t)payloadBytes + 1);

  if (rc)
  {
    free(areas);
    return rc;
  }

  cudaError_t e = cudaMemcpyAsync(ctx->dAreas, areas, (size_t)count * sizeof(limgcu_area), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->dPayload, p, (size_t)payloadBytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // `areas` is pageable host memory that is freed below
  free(areas);

  if (e != cudaSuccess)
    return fail(ctx, LIMGCU_ERROR_CUDA, "container H2D", e);

  rc = limgcu_build_block_map(ctx, ctx->dAreas, count, W, H, ctx->dBlockToArea);
  if (rc) return rc;
  rc = limgcu_unpack_payload(ctx, ctx->dAreas, count, ctx->dBlockToArea, ctx->dPayload, reinterpret_cast<uint64_t *>(ctx->dPayloadOff), W, H, hasAlpha, ctx->dPlaneU8[4], ctx->dPlaneU8[5],
                             ctx->dPlaneU8[6]);
  if (rc) return rc;
  rc = limgcu_decode(ctx, ctx->dAreas, ctx->dBlockToArea, ctx->dPlaneU8[4], ctx->dPlaneU8[5], ctx->dPlaneU8[6], W, H, hasAlpha, ctx->dPlaneU32[0]);
  if (rc) return rc;
  CK(cudaMemcpyAsync(pOut, ctx->dPlaneU32[0], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// batches of independent frames: `lanes` contexts of one (or several) devices, one host thread each; frame i runs on lane i % lanes
// ---------------------------------------------------------------------------------------------------------------------------------

template <class Fn>
static int run_lanes(limgcu_ctx *const *ctxs, int lanes, int count, Fn &&fn)
{
  if (ctxs == nullptr || lanes < 1 || count < 0)
    return LIMGCU_ERROR_INVALID_PARAMETER;

  for (int j = 0; j < lanes; j++)
    if (ctxs[j] == nullptr)
      return LIMGCU_ERROR_ARGUMENT_NULL;

  std::vector<int> rc((size_t)lanes, LIMGCU_SUCCESS);
  std::vector<std::thread> threads;
  const int used = lanes < count ? lanes : count;

  for (int j = 0; j < used; j++)
    threads.emplace_back([&, j]() {
      for (int i = j; i < count && rc[j] == LIMGCU_SUCCESS; i += lanes)
        rc[j] = fn(ctxs[j], i);
    });

  for (auto &t : threads)
    t.join();

  for (int j = 0; j < used; j++)
    if (rc[j] != LIMGCU_SUCCESS)
      return rc[j]; // limgcu_last_error of that lane's context has the text

  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_batch_host_encode_containers(limgcu_ctx *const *ctxs, int lanes, const uint32_t *const *frames, int count, size_t sizeX, size_t sizeY, int hasAlpha,
                                                   uint32_t errorFactor, uint32_t flags, void *const *outs, const size_t *capacities, size_t *written)
{
  if (frames == nullptr || outs == nullptr || capacities == nullptr || written == nullptr)
    return LIMGCU_ERROR_ARGUMENT_NULL;

  return run_lanes(ctxs, lanes, count, [&](limgcu_ctx *ctx, int i) {
    return limgcu_host_encode_container(ctx, frames[i], sizeX, sizeY, hasAlpha, errorFactor, flags, outs[i], capacities[i], &written[i]);
  });
}

extern "C" int limgcu_batch_host_decode_containers(limgcu_ctx *const *ctxs, int lanes, const void *const *containers, const size_t *bytes, int count, uint32_t *const *outs,
                                                   const size_t *outPixels)
{
  if (containers == nullptr || bytes == nullptr || outs == nullptr || outPixels == nullptr)
    return LIMGCU_ERROR_ARGUMENT_NULL;

  return run_lanes(ctxs, lanes, count, [&](limgcu_ctx *ctx, int i) { return limgcu_host_decode_container(ctx, containers[i], bytes[i], outs[i], outPixels[i]); });
}

extern "C" int limgcu_host_pass1(limgcu_ctx *ctx, const uint32_t *pIn, size_t sizeX, size_t sizeY, int hasAlpha, limgcu_decomp *table)
{
  NEED(ctx); NEED(pIn); NEED(table);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t n = sizeX * sizeY, blocks = ((sizeX + 7) / 8) * ((sizeY + 7) / 8);
  rc = ensure_staging(ctx, n);
  if (rc) return rc;
  rc = ensure_capacity(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->dSrc, pIn, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_pass1(ctx, ctx->dSrc, sizeX, sizeY, hasAlpha, ctx->dTable);
  if (rc) return rc;
  CK(cudaMemcpyAsync(table, ctx->dTable, blocks * sizeof(limgcu_decomp), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_host_merge(limgcu_ctx *ctx, const limgcu_decomp *table, size_t sizeX, size_t sizeY, int hasAlpha, limgcu_area *areas, uint32_t *area_count)
{
  NEED(ctx); NEED(table); NEED(areas); NEED(area_count);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t blocks = ((sizeX + 7) / 8) * ((sizeY + 7) / 8);
  rc = ensure_capacity(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->dTable, table, blocks * sizeof(limgcu_decomp), cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_merge(ctx, ctx->dTable, sizeX, sizeY, hasAlpha, ctx->dAreas, ctx->dBlockToArea, false);
  if (rc) return rc;
  uint32_t count = 0, overflow[2] = { 0, 0 };
  CK(cudaMemcpyAsync(&count, ctx->dCounters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(overflow, ctx->dCounters + 27, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));

  if (int bad = scan_flags_error(ctx, overflow))
    return bad;

  CK(cudaMemcpyAsync(areas, ctx->dAreas, (size_t)count * sizeof(limgcu_area), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *area_count = count;
  return LIMGCU_SUCCESS;
}

extern "C" int limgcu_debug_predicate_check(limgcu_ctx *ctx, const limgcu_decomp *table, size_t sizeX, size_t sizeY, int hasAlpha, uint64_t *out4)
{
  NEED(ctx); NEED(table); NEED(out4);
  int rc = check_image(ctx, sizeX, sizeY);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  rc = ensure_capacity(ctx, sizeX, sizeY);
  if (rc) return rc;
  const int BX = (int)((sizeX + 7) / 8), BY = (int)((sizeY + 7) / 8), blocks = BX * BY;
  unsigned long long *dOut = nullptr;
  CK(cudaMalloc(&dOut, 4 * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(dOut, 0, 4 * sizeof(unsigned long long), ctx->stream));
  CK(cudaMemcpyAsync(ctx->dTable, table, (size_t)blocks * sizeof(limgcu_decomp), cudaMemcpyHostToDevice, ctx->stream));

  if (hasAlpha)
  {
    k_pred_records<4><<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(ctx->dTable, blocks, ctx->dRec);
    k_pred_check<4><<<(blocks * 2 + 7) / 8, 256, 0, ctx->stream>>>(ctx->dRec, BX, BY, dOut);
  }
  else
  {
    k_pred_records<3><<<(blocks + 255) / 256, 256, 0, ctx->stream>>>(ctx->dTable, blocks, ctx->dRec);
    k_pred_check<3><<<(blocks * 2 + 7) / 8, 256, 0, ctx->stream>>>(ctx->dRec, BX, BY, dOut);
  }

  CKL("k_pred_check");
  cudaError_t e = cudaMemcpyAsync(out4, dOut, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dOut);
  if (e != cudaSuccess) return fail(ctx, LIMGCU_ERROR_CUDA, "limgcu_debug_predicate_check", e);
  return LIMGCU_SUCCESS;
}

extern "C" double limgcu_host_compare(limgcu_ctx *ctx, const uint32_t *pImageA, const uint32_t *pImageB, size_t sizeX, size_t sizeY, int hasAlpha, double *pMeanSquaredError, double *pMaxPossibleSquaredError)
{
  if (ctx == nullptr || pImageA == nullptr || pImageB == nullptr)
    return NAN;

  if (cudaSetDevice(ctx->device) != cudaSuccess)
    return NAN;

  const size_t n = sizeX * sizeY;

  if (ensure_staging(ctx, n) != LIMGCU_SUCCESS)
    return NAN;

  if (cudaMemcpyAsync(ctx->dPlaneU32[0], pImageA, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return NAN;
  if (cudaMemcpyAsync(ctx->dPlaneU32[1], pImageB, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return NAN;

  double psnr = NAN;

  if (limgcu_compare(ctx, ctx->dPlaneU32[0], ctx->dPlaneU32[1], sizeX, sizeY, hasAlpha, &psnr, pMeanSquaredError, pMaxPossibleSquaredError) != LIMGCU_SUCCESS)
    return NAN;

  return psnr;
}

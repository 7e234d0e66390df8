// limg_b200/csrc/kernels_fit.cuh -- pass 1 (per-8x8-block fit) and the per-area encode kernels
// (gather -> refit -> projection -> shift search). One warp per small area, one CTA per large area.
#pragma once

#include "group.cuh"

#include <cuda.h> // CUtensorMap (the driver entry point that encodes it is looked up at run time: no link dependency)

namespace limg
{

#define LIMG_SMALL_AREA_PX 256   // <= 4 blocks: one warp, everything in shared memory
#define LIMG_CTA_AREA_CAP 4096   // <= 64 blocks: one CTA, pixels + factors in shared memory; above: global scratch
#define LIMG_HUGE_AREA_PX 16384  // above: a 512-thread CTA per area, started first (k_encode_large<CH, 512, true>)
#define LIMG_CTA_STAGE_PX 1024
#define LIMG_ENCODE_THREADS 256

// per-area bookkeeping produced by k_area_prepare
struct AreaWork
{
  uint32_t n;          // pixels
  uint32_t scratchOff; // exclusive scan of n (area-contiguous scratch for large areas)
};

__device__ __forceinline__ void load_lut(uint16_t *sLut, const uint16_t *__restrict__ gLut)
{
  for (int i = threadIdx.x; i < 2048 / 2; i += blockDim.x)
    reinterpret_cast<uint32_t *>(sLut)[i] = reinterpret_cast<const uint32_t *>(gLut)[i];
}

__device__ __forceinline__ void store_decomp(limgcu_decomp *dst, const limgcu_decomp &d)
{
  *dst = d;
}

// ---------------------------------------------------------------------------------------------
// pass 1: one warp per 8x8 block (limg.cpp:1088-1119)
// ---------------------------------------------------------------------------------------------

template <int CH>
__global__ void __launch_bounds__(256) k_pass1(const uint32_t *__restrict__ src, int W, int H, int BX, int BY, const uint16_t *__restrict__ gLut, limgcu_decomp *__restrict__ table)
{
  __shared__ uint16_t sLut[2048];
  __shared__ uint32_t sPx[8][64];
  __shared__ float4 sStage[8][64];
  __shared__ GroupScratch<1> sGs[8];

  load_lut(sLut, gLut);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t parity = 0;

  // persistent warps: the 4 KB rsqrt table is staged once per CTA, not once per eight blocks
  for (int b = blockIdx.x * 8 + warp; b < BX * BY; b += gridDim.x * 8)
  {
    const int by = b / BX, bx = b - by * BX;
    const int x0 = bx * LIMG_BLOCK, y0 = by * LIMG_BLOCK;
    const int w = min(LIMG_BLOCK, W - x0), h = min(LIMG_BLOCK, H - y0);
    const uint32_t n = (uint32_t)(w * h);

    for (uint32_t i = lane; i < n; i += 32)
    {
      const int row = i / w, col = i - row * w;
      sPx[warp][i] = src[(size_t)(y0 + row) * W + x0 + col];
    }

    __syncwarp();

    limgcu_decomp d;
    group_fit<CH, 1>(sPx[warp], n, sLut, sStage[warp], 64, &sGs[warp], parity, d);

    if (lane == 0)
      store_decomp(&table[b], d);

    __syncwarp();
  }
}

// The same fit with the pixels staged by the TMA: every warp runs its own two-deep pipeline of 2-D tensor-map tile loads (cp.async.bulk.tensor, one
// 8 x 8 pixel box = 256 bytes per block, completion counted in bytes on the warp's mbarrier): lane 0 requests the warp's NEXT block before the warp
// waits for the current one, so the load is in flight during the ~1400 instructions of a fit and no warp waits for another. A box arrives row by row,
// 8 pixels each, which IS the reference's pixel order of a full block (limg.cpp:1106-1107); pixels right of or below the image arrive as zeros, so ragged
// edge blocks need no address arithmetic, only a repack to their rx * ry pixels. Needs a row pitch that is a multiple of 16 bytes (sizeX % 4 == 0);
// k_pass1 above is the path for the other widths.
template <int CH>
__global__ void __launch_bounds__(256) k_pass1_tma(const __grid_constant__ CUtensorMap srcMap, int W, int H, int BX, int BY, const uint16_t *__restrict__ gLut, limgcu_decomp *__restrict__ table)
{
  __shared__ __align__(128) uint32_t sTile[8][2][LIMG_BLOCK * LIMG_BLOCK];
  __shared__ __align__(8) unsigned long long sBar[8][2];
  __shared__ uint16_t sLut[2048];
  __shared__ uint32_t sPx[8][64];
  __shared__ float4 sStage[8][64];
  __shared__ GroupScratch<1> sGs[8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTileBytes = sizeof(uint32_t) * LIMG_BLOCK * LIMG_BLOCK;

  if (lane == 0)
  {
    mbar_init(smem_u32(&sBar[warp][0]), 1);
    mbar_init(smem_u32(&sBar[warp][1]), 1);
    fence_mbarrier_init();
  }

  load_lut(sLut, gLut);
  __syncthreads();

  auto request = [&](int b, int slot) {
    const int by = b / BX, bx = b - by * BX;
    mbar_expect_tx(smem_u32(&sBar[warp][slot]), kTileBytes);
    tensor_g2s_2d(smem_u32(&sTile[warp][slot][0]), &srcMap, bx * LIMG_BLOCK, by * LIMG_BLOCK, smem_u32(&sBar[warp][slot]));
  };

  const int stride = (int)gridDim.x * 8;
  int b = (int)blockIdx.x * 8 + warp;

  if (lane == 0 && b < BX * BY)
    request(b, 0);

  uint32_t parity = 0;

  for (int it = 0; b < BX * BY; b += stride, it++)
  {
    const int slot = it & 1;

    // the other slot is free: the warp finished the block it held (the __syncwarp at the end of the previous iteration)
    if (lane == 0 && b + stride < BX * BY)
      request(b + stride, slot ^ 1);

    mbar_wait(smem_u32(&sBar[warp][slot]), (uint32_t)((it >> 1) & 1));

    const int by = b / BX, bx = b - by * BX;
    const int w = min(LIMG_BLOCK, W - bx * LIMG_BLOCK), h = min(LIMG_BLOCK, H - by * LIMG_BLOCK);
    const uint32_t n = (uint32_t)(w * h);
    const uint32_t *px = sTile[warp][slot];

    if (w != LIMG_BLOCK)
    {
      // ragged in x: the box's rows are 8 pixels apart, the block's rows w
      for (uint32_t i = lane; i < n; i += 32)
      {
        const int row = (int)i / w, col = (int)i - row * w;
        sPx[warp][i] = sTile[warp][slot][row * LIMG_BLOCK + col];
      }

      __syncwarp();
      px = sPx[warp];
    }

    limgcu_decomp d;
    group_fit<CH, 1>(px, n, sLut, sStage[warp], 64, &sGs[warp], parity, d);

    if (lane == 0)
      store_decomp(&table[b], d);

    __syncwarp(); // the slot can be refilled
  }
}

// ---------------------------------------------------------------------------------------------
// per-area encode: gather, refit (merged areas), projection, shift search -> decomposition + shifts + dither demand
// (limg.cpp:1717-1772, 1498-1535)
// ---------------------------------------------------------------------------------------------

struct EncodeArgs
{
  const uint32_t *src;
  int W, H, BX, BY;
  const uint16_t *lut;
  const limgcu_decomp *table; // pass-1 table (leftover areas keep their fit)
  limgcu_area *areas;
  const uint32_t *areaCount;
  const AreaWork *work;
  uint64_t *ditherDemand;     // per area: pixels consumed from the dither chain
  uint32_t *scratchPx, *scratchFac;
  uint32_t *workCounter;      // dynamic scheduling
  const uint32_t *list;       // area indices of this size class
  const uint32_t *listCount;
  const uint32_t *hugeCount;  // k_encode_large: list[0 .. huge) from the front (huge areas), list[listCap - 1 - i] for the other listCount entries
  uint32_t listCap;
  const uint32_t *bigCount;   // k_encode_large<.., false>: first the bigList[listCap - 1 - i] entries (areas beyond shared memory), then the list's
  const uint32_t *bigList;
  uint32_t rowLo, rowHi;      // only areas whose first block row lies in [rowLo, rowHi) are encoded (row-band sharding; the default is everything)
  CrushParams cp;
};

template <int CH, int WARPS>
__device__ void encode_area_group(const EncodeArgs &a, uint32_t k, uint32_t *px, uint32_t *fac, const uint16_t *lut, float4 *stage, int stagePx, GroupScratch<WARPS> *gs)
{
  constexpr int THREADS = WARPS * 32;
  const int t = group_tid<WARPS>();
  limgcu_area *area = &a.areas[k];
  const uint32_t pw = area->px_w, ph = area->px_h, pxX = area->px_x, pxY = area->px_y;
  const uint32_t n = pw * ph;
  uint32_t parity = 0;

  // gather into area-contiguous order (limg.cpp:1752-1753)
  for (uint32_t i = t; i < n; i += THREADS)
  {
    const uint32_t row = i / pw, col = i - row * pw;
    px[i] = a.src[(size_t)(pxY + row) * a.W + pxX + col];
  }

  group_sync<WARPS>();

  limgcu_decomp d;

  if (area->stage == 2)
    d = a.table[(size_t)area->oy * a.BX + area->ox]; // leftovers keep the pass-1 decomposition (limg.cpp:1875)
  else
    group_fit<CH, WARPS>(px, n, lut, stage, stagePx, gs, parity, d);

  Proj p;
  init_proj<CH>(d, p);

  for (uint32_t i = t; i < n; i += THREADS)
    fac[i] = project_px<CH>(p, px[i]);

  group_sync<WARPS>();

  int shift[3] = { 0, 0, 0 };

  if (a.cp.crushBits)
  {
    auto trial = [&](int sa, int sb, int sc, uint64_t &err) -> bool {
      return group_trial<CH, WARPS>(px, fac, n, d, sa, sb, sc, a.cp, gs, parity, err);
    };
    search_shifts(trial, a.cp.fast != 0, shift);
  }

  if (t == 0)
  {
    area->decomp = d;
    area->shift[0] = (uint8_t)shift[0];
    area->shift[1] = (uint8_t)shift[1];
    area->shift[2] = (uint8_t)shift[2];
    area->pad = 0;
    uint32_t planes = 0;
    for (int i = 0; i < 3; i++)
      planes += (shift[i] != 0 && shift[i] != 8) ? 1u : 0u;
    a.ditherDemand[k] = (uint64_t)planes * n; // Q13: only planes with 0 < shift < 8 consume the chain
  }

  group_sync<WARPS>();
}

template <int CH>
__global__ void __launch_bounds__(LIMG_ENCODE_THREADS) k_encode_small(EncodeArgs a)
{
  constexpr int WPB = LIMG_ENCODE_THREADS / 32;
  __shared__ uint16_t sLut[2048];
  __shared__ uint32_t sPx[WPB][LIMG_SMALL_AREA_PX];
  __shared__ uint32_t sFac[WPB][LIMG_SMALL_AREA_PX];
  __shared__ float4 sStage[WPB][64];
  __shared__ GroupScratch<1> sGs[WPB];

  load_lut(sLut, a.lut);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t count = *a.listCount;

  while (true)
  {
    uint32_t j = 0;

    if (lane == 0)
      j = atomicAdd(a.workCounter, 1u);

    j = __shfl_sync(0xFFFFFFFFu, j, 0);

    if (j >= count)
      break;

    const uint32_t k = a.list[j];
    const uint32_t oy = a.areas[k].oy;

    if (oy < a.rowLo || oy >= a.rowHi)
      continue;

    encode_area_group<CH, 1>(a, k, sPx[warp], sFac[warp], sLut, sStage[warp], 64, &sGs[warp]);
  }
}

// THREADS = 256 for the areas that fit into shared memory (four CTAs per SM), 512 for the huge ones (list[0 .. huge)): a huge area is a long pole
// on ONE SM, and twice the warps roughly double the issue rate its trials get there. first / last select the part of the list:
// HUGE: jobs [0, huge) from the front; otherwise the other listCount entries from the back.
template <int CH, int THREADS, bool HUGE>
__global__ void __launch_bounds__(THREADS) k_encode_large(EncodeArgs a)
{
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(16) unsigned char dynSmem[];
  uint16_t *sLut = reinterpret_cast<uint16_t *>(dynSmem);
  float4 *sStage = reinterpret_cast<float4 *>(dynSmem + 4096);
  uint32_t *sPx = reinterpret_cast<uint32_t *>(dynSmem + 4096 + LIMG_CTA_STAGE_PX * 16);
  uint32_t *sFac = sPx + LIMG_CTA_AREA_CAP;
  __shared__ GroupScratch<WARPS> sGs;
  __shared__ uint32_t sJob;

  const uint32_t big = HUGE ? 0u : *a.bigCount;
  const uint32_t count = HUGE ? *a.hugeCount : big + *a.listCount;

  if (count == 0)
    return; // photo-like content has no huge areas: nothing to stage

  load_lut(sLut, a.lut);
  __syncthreads();

  while (true)
  {
    if (threadIdx.x == 0)
      sJob = atomicAdd(a.workCounter, 1u);

    __syncthreads();
    const uint32_t j = sJob;
    __syncthreads();

    if (j >= count)
      break;

    const uint32_t k = HUGE ? a.list[j] : (j < big ? a.bigList[a.listCap - 1u - j] : a.list[a.listCap - 1u - (j - big)]);
    const uint32_t oy = a.areas[k].oy;

    if (oy < a.rowLo || oy >= a.rowHi)
      continue;

    const AreaWork w = a.work[k];
    uint32_t *px = w.n <= LIMG_CTA_AREA_CAP ? sPx : a.scratchPx + w.scratchOff;
    uint32_t *fac = w.n <= LIMG_CTA_AREA_CAP ? sFac : a.scratchFac + w.scratchOff;
    encode_area_group<CH, WARPS>(a, k, px, fac, sLut, sStage, LIMG_CTA_STAGE_PX, &sGs);
  }
}

} // namespace limg

// limg_b200/csrc/kernels_merge.cuh -- area expansion (limg.cpp:1121-1135, 1137-1269, 1277-1496, 1814-1878).
//
// The reference's merge is a serial greedy raster scan whose only inputs are the pass-1 table and the scan order
// (SURVEY.md Q2): the predicate "candidate block matches seed block" is a pure function of two pass-1 records.
// B200 design:
//   1. k_pred_records : one thread per block derives the predicate-side state of its record once (the divisions).
//   2. k_pred_window  : for EVERY block as a hypothetical seed, all 63 predicates against the 8x8 window to its lower right are
//                       evaluated in parallel (one thread per pair) -> one 64-bit match word per block.
//   3. k_merge_banded : one CTA per band of block rows replays the reference's scan order over its rows (see "banded scan"
//                       below for why the concurrent bands converge to the sequential result). Inside a band warp 0 walks
//                       candidate seeds with pure bit arithmetic on (match word & ~in-use window); growth that leaves the
//                       window and the four-way centre-third regrowth evaluate their predicates on demand, one warp per
//                       predicate (27 lanes = the 27 samples), across all warps of the CTA.
//   4. k_area_prepare : leftover blocks (raster order), pixel rectangles, block->area map, size classes, scratch offsets.
#pragma once

#include "group.cuh"
#include "kernels_fit.cuh"

namespace limg
{

// predicate-side view of one pass-1 record (limg_init_color_error_state_3d + the loop at limg.cpp:1150-1161, 1201-1212)
struct __align__(16) PredRec
{
  float avg[4];
  float minA[4], offB[4], offC[4]; // as float
  float nA[4], nB[4], nC[4];
  float inv[3];                    // 1 / dot(n, n) or 0
  float invLen[3];                 // 1 / lenSq, entries 1 and 2 doubled
  float sumLen;                    // lenSq[0] + lenSq[1] + lenSq[2]
  float pad;
};

template <int CH>
__global__ void k_pred_records(const limgcu_decomp *__restrict__ table, int count, PredRec *__restrict__ rec)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;

  if (i >= count)
    return;

  const limgcu_decomp d = table[i];
  Proj p;
  init_proj<CH>(d, p);
  PredRec r;
  const float nA[4] = { p.nA.x, p.nA.y, p.nA.z, p.nA.w }, nB[4] = { p.nB.x, p.nB.y, p.nB.z, p.nB.w }, nC[4] = { p.nC.x, p.nC.y, p.nC.z, p.nC.w };
  const float w[4] = { 2, 4, 3, 3 };
  float len[3] = { 3, 3, 3 }; // limg.cpp:1145

#pragma unroll
  for (int c = 0; c < 4; c++)
  {
    r.avg[c] = d.avg[c];
    r.minA[c] = (float)d.dirA_min[c];
    r.offB[c] = (float)d.dirB_offset[c];
    r.offC[c] = (float)d.dirC_offset[c];
    r.nA[c] = nA[c];
    r.nB[c] = nB[c];
    r.nC[c] = nC[c];

    if (c < CH)
    {
      len[0] = fadd(len[0], fmul(fmul(nA[c], nA[c]), w[c]));
      len[1] = fadd(len[1], fmul(fmul(nB[c], nB[c]), w[c]));
      len[2] = fadd(len[2], fmul(fmul(nC[c], nC[c]), w[c]));
    }
  }

  r.inv[0] = p.invA;
  r.inv[1] = p.invB;
  r.inv[2] = p.invC;
  r.invLen[0] = frcp1(len[0]);
  r.invLen[1] = fmul(frcp1(len[1]), 2.0f);
  r.invLen[2] = fmul(frcp1(len[2]), 2.0f);
  r.sumLen = fadd(fadd(len[0], len[1]), len[2]);
  r.pad = 0.0f;
  rec[i] = r;
}

// limg_color_error_state_3d_get_factors (limg_factorization.h:9-41): sequential dot products starting from 0.
template <int CH>
__device__ __forceinline__ float factor_term(const float color[4], const PredRec &s, const float invLen[3])
{
  float t[4], est[4];
  float dA = 0.0f, dB = 0.0f, dC = 0.0f;

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    t[i] = fsub(color[i], s.minA[i]);
    dA = fadd(dA, fmul(t[i], s.nA[i]));
  }

  const float facA = fmul(dA, s.inv[0]);

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    est[i] = fadd(s.minA[i], fmul(facA, s.nA[i]));
    t[i] = fsub(fsub(color[i], est[i]), s.offB[i]);
    dB = fadd(dB, fmul(t[i], s.nB[i]));
  }

  const float facB = fmul(dB, s.inv[1]);

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    est[i] = fadd(est[i], fmul(facB, s.nB[i]));
    t[i] = fsub(fsub(color[i], est[i]), s.offC[i]);
    dC = fadd(dC, fmul(t[i], s.nC[i]));
  }

  const float facC = fmul(dC, s.inv[2]);
  // fabsf(fac_a) * inv[0] + fabsf(0.5f - fac_b) * inv[1] + fabsf(0.5f - fac_c) * inv[2]
  return fadd(fadd(fmul(fabsf(facA), invLen[0]), fmul(fabsf(fsub(0.5f, facB)), invLen[1])), fmul(fabsf(fsub(0.5f, facC)), invLen[2]));
}

// returns 1 (early accept), 0 (ratio reject) or -1 (needs the 27-sample score)
template <int CH>
__device__ __forceinline__ int predicate_quick(const PredRec &a, const PredRec &b)
{
  const float w[4] = { 2, 4, 3, 3 };
  float avgDiffSq = 0.0f;

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    const float diff = fsub(a.avg[i], b.avg[i]);
    avgDiffSq = fadd(avgDiffSq, fmul(fmul(diff, diff), w[i]));
  }

  const float acceptAvg = (float)(16 * 3 * CH), acceptRange = (float)(200 * 3 * CH);

  if (avgDiffSq < acceptAvg && a.sumLen < acceptRange && b.sumLen < acceptRange)
    return 1;

  const float ratio = __fdiv_rn(fadd(a.sumLen, 1.0f), fadd(b.sumLen, 1.0f));
  const float maxRatio = 1.375f;

  if (ratio > maxRatio || ratio < (1.f / maxRatio))
    return 0;

  return -1;
}

template <int CH>
__device__ __forceinline__ float sample_term(const PredRec &a, const PredRec &b, int k)
{
  const int z = k / 9, y = (k / 3) % 3, x = k % 3;
  const float xf = x * 0.5f, yf = y * 0.5f, zf = z * 0.5f;
  float color[4];

#pragma unroll
  for (int i = 0; i < CH; i++)
    color[i] = fadd(fadd(fmul(b.nA[i], xf), fmul(b.nB[i], yf)), fmul(b.nC[i], zf));

  return factor_term<CH>(color, a, a.invLen);
}

// one thread evaluates the whole predicate (limg.cpp:1137-1269). a = seed, b = candidate.
template <int CH>
__device__ bool predicate_thread(const PredRec &a, const PredRec &b)
{
  const int q = predicate_quick<CH>(a, b);

  if (q >= 0)
    return q != 0;

  // Q1: the second term of every iteration projects avg(a) into b: loop invariant, but part of the ordered sum.
  const float constTerm = factor_term<CH>(a.avg, b, b.invLen);
  float sum = 0.0f;

  for (int k = 0; k < 27; k++)
  {
    sum = fadd(sum, sample_term<CH>(a, b, k));
    sum = fadd(sum, constTerm);
  }

  return fmul(sum, 1.f / (3 * 3 * 3)) < 3.0f;
}

// one warp evaluates one predicate: lane k < 27 computes sample k, the ordered sum is replayed by every lane.
template <int CH>
__device__ bool predicate_warp(const PredRec &a, const PredRec &b)
{
  const int q = predicate_quick<CH>(a, b);

  if (q >= 0)
    return q != 0;

  const int lane = threadIdx.x & 31;
  const float constTerm = factor_term<CH>(a.avg, b, b.invLen);
  const float term = sample_term<CH>(a, b, lane < 27 ? lane : 0);
  float sum = 0.0f;

#pragma unroll
  for (int k = 0; k < 27; k++)
  {
    sum = fadd(sum, __shfl_sync(0xFFFFFFFFu, term, k));
    sum = fadd(sum, constTerm);
  }

  return fmul(sum, 1.f / (3 * 3 * 3)) < 3.0f;
}

// match word of every block: bit (dy * 8 + dx) = predicate(seed = block, candidate = block + (dx, dy)), 0 outside the grid.
template <int CH>
__global__ void __launch_bounds__(256) k_pred_window(const PredRec *__restrict__ rec, int BX, int BY, uint32_t *__restrict__ window /* 2 words per block */)
{
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5); // (seed, half) pairs: 8 per CTA
  const int seed = pair >> 1, half = pair & 1;

  if (seed >= BX * BY)
    return;

  const int lane = threadIdx.x & 31;
  const int o = half * 32 + lane;
  const int dx = o & 7, dy = o >> 3;
  const int sy = seed / BX, sx = seed - sy * BX;
  bool m = false;

  if (o == 0)
    m = true;
  else if (sx + dx < BX && sy + dy < BY)
    m = predicate_thread<CH>(rec[seed], rec[(size_t)(sy + dy) * BX + sx + dx]);

  const uint32_t bits = __ballot_sync(0xFFFFFFFFu, m);

  if (lane == 0)
    window[(size_t)seed * 2 + half] = bits;
}

// Symmetric match window of every block c: bit (r, col) of its 16 x 16 bitmap = predicate(seed = c, candidate = c + (col - 8, r - 8)),
// 0 outside the grid. It is what the four-way centre-third regrowth (limg.cpp:1426-1433) asks for: the centre seed sits inside
// the right/down rectangle it came from, so the regrowth explores a neighbourhood on all four sides. The lower-right quadrant is
// the 8 x 8 match word. Stored as 8 words per block (two 16-bit rows per word).
template <int CH>
__global__ void __launch_bounds__(256) k_pred_symwindow(const PredRec *__restrict__ rec, const uint32_t *__restrict__ window, int BX, int BY, uint32_t *__restrict__ sym)
{
  const int c = blockIdx.x;
  const int cy = c / BX, cx = c - cy * BX;
  const int r = threadIdx.x >> 4, col = threadIdx.x & 15;
  const int dx = col - 8, dy = r - 8;
  bool m = false;

  if (dx >= 0 && dy >= 0)
  {
    const uint32_t w = window[(size_t)c * 2 + (dy >> 2)];
    m = (w >> (8 * (dy & 3) + dx)) & 1u;
  }
  else if (cx + dx >= 0 && cx + dx < BX && cy + dy >= 0 && cy + dy < BY)
  {
    m = predicate_thread<CH>(rec[c], rec[(size_t)(cy + dy) * BX + cx + dx]);
  }

  const uint32_t b = __ballot_sync(0xFFFFFFFFu, m);

  if ((threadIdx.x & 31) == 0)
    sym[(size_t)c * 8 + (threadIdx.x >> 5)] = b;
}

// ---------------------------------------------------------------------------------------------
// speculative windows: everything the scan is likely to ask for is evaluated up front, in parallel on the whole GPU.
//   - seeds whose mask-free right/down growth leaves the 8x8 window get a 16x16 and, if that is left too, a 32x32 match bitmap;
//   - seeds whose mask-free growth reaches 3x3 get the match bitmap of their centre-third seed over the neighbourhood the
//     four-way regrowth explores (limg.cpp:1426-1433), keyed by the (rx, ry) the prediction assumed.
// The scan uses a bitmap only when its assumptions hold and falls back to on-demand evaluation otherwise, so the bitmaps are a
// pure accelerator: they never change the result.
// ---------------------------------------------------------------------------------------------

#define LIMG_NO_SLOT 0xFFFFFFFFu

struct PlanArgs
{
  const PredRec *rec;
  const uint32_t *window;
  int BX, BY;
  uint32_t *extSlot;   // per block: LIMG_NO_SLOT or slot | (size 32 ? 1u << 31 : 0)
  uint32_t *extSeed;   // per slot: block index
  uint32_t *extBits;   // per slot: 32 row words
  uint32_t *ctrSlot;   // per block: LIMG_NO_SLOT or slot
  uint4 *ctrHdr;       // per slot: x = rx0 | ry0 << 16 (assumed growth), y = rgX | rgY << 16, z = rgW | rgH << 16, w = centre block index
  uint32_t *ctrBits;   // per slot: 32 row words
  uint32_t *counters;  // [0] ext slots, [1] centre slots
  uint32_t extCap, ctrCap;
};

// mask-free alternating right/down growth over `rows` (S x S match bitmap of the seed); returns true if it wanted to leave the bitmap
__device__ __forceinline__ bool expand_unmasked(const uint32_t *rows, int S, int x, int y, int BX, int BY, int &rx, int &ry)
{
  bool right = true, down = true, hit = false;
  rx = 1;
  ry = 1;

  while (right || down)
  {
    if (right)
    {
      bool ok = x + rx + 1 < BX;

      if (ok && rx >= S) { hit = true; ok = false; }

      if (ok)
        for (int r = 0; r < ry; r++)
          ok &= (rows[r] >> rx) & 1u;

      if (ok) rx++; else right = false;
    }

    if (down)
    {
      bool ok = y + ry + 1 < BY;

      if (ok && ry >= S) { hit = true; ok = false; }

      if (ok)
      {
        const uint32_t m = rx >= 32 ? 0xFFFFFFFFu : ((1u << rx) - 1u);
        ok = (rows[ry] & m) == m;
      }

      if (ok) ry++; else down = false;
    }
  }

  return hit;
}

__device__ __forceinline__ void plan_centre(const PlanArgs &a, int seed, int x, int y, int rx, int ry)
{
  if (rx < 3 || ry < 3)
    return;

  const uint32_t slot = atomicAdd(&a.counters[1], 1u);

  if (slot >= a.ctrCap)
    return;

  const int rgX = max(x - 3, 0), rgY = max(y - 3, 0);
  const int rgW = min(min(rx + 6, 32), a.BX - rgX), rgH = min(min(ry + 6, 32), a.BY - rgY);
  const int centre = (y + ry / 3) * a.BX + x + rx / 3;
  a.ctrHdr[slot] = make_uint4((uint32_t)rx | ((uint32_t)ry << 16), (uint32_t)rgX | ((uint32_t)rgY << 16), (uint32_t)rgW | ((uint32_t)rgH << 16), (uint32_t)centre);
  a.ctrSlot[seed] = slot;
}

__global__ void __launch_bounds__(256) k_plan_seeds(PlanArgs a)
{
  const int seed = blockIdx.x * blockDim.x + threadIdx.x;

  if (seed >= a.BX * a.BY)
    return;

  const int y = seed / a.BX, x = seed - y * a.BX;
  const uint32_t w0 = a.window[(size_t)seed * 2], w1 = a.window[(size_t)seed * 2 + 1];
  uint32_t rows[8];

#pragma unroll
  for (int r = 0; r < 4; r++)
  {
    rows[r] = (w0 >> (8 * r)) & 0xFF;
    rows[r + 4] = (w1 >> (8 * r)) & 0xFF;
  }

  int rx, ry;
  const bool hit = expand_unmasked(rows, 8, x, y, a.BX, a.BY, rx, ry);
  a.extSlot[seed] = LIMG_NO_SLOT;
  a.ctrSlot[seed] = LIMG_NO_SLOT;

  if (hit)
  {
    const uint32_t slot = atomicAdd(&a.counters[0], 1u);

    if (slot < a.extCap)
    {
      a.extSeed[slot] = seed;
      a.extSlot[seed] = slot;
    }
  }
}

template <int CH>
__global__ void __launch_bounds__(256) k_plan_extend(PlanArgs a)
{
  __shared__ uint32_t rows[32];
  __shared__ int sHit, sRx, sRy;
  const uint32_t count = min(a.counters[0], a.extCap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (uint32_t slot = blockIdx.x; slot < count; slot += gridDim.x)
  {
    const int seed = (int)a.extSeed[slot];
    const int y = seed / a.BX, x = seed - y * a.BX;
    const PredRec s = a.rec[seed];
    const uint32_t w0 = a.window[(size_t)seed * 2], w1 = a.window[(size_t)seed * 2 + 1];

    if (threadIdx.x < 32)
      rows[threadIdx.x] = 0;

    __syncthreads();

    // 16 x 16: thread t -> (dx, dy) = (t & 15, t >> 4); the 8 x 8 corner is known
    {
      const int dx = threadIdx.x & 15, dy = threadIdx.x >> 4;
      bool m = false;

      if (dx < 8 && dy < 8)
        m = ((dy < 4 ? w0 >> (8 * dy) : w1 >> (8 * (dy - 4))) >> dx) & 1u;
      else if (x + dx < a.BX && y + dy < a.BY)
        m = predicate_thread<CH>(s, a.rec[(size_t)(y + dy) * a.BX + x + dx]);

      const uint32_t b = __ballot_sync(0xFFFFFFFFu, m);

      if (lane == 0)
      {
        rows[warp * 2] = b & 0xFFFF;
        rows[warp * 2 + 1] = b >> 16;
      }
    }

    __syncthreads();

    if (threadIdx.x == 0)
    {
      int rx, ry;
      sHit = expand_unmasked(rows, 16, x, y, a.BX, a.BY, rx, ry) ? 1 : 0;
      sRx = rx;
      sRy = ry;
    }

    __syncthreads();
    int size = 16;

    if (sHit)
    {
      // 32 x 32: four passes of eight rows; the 16 x 16 corner is known
      size = 32;

      for (int p = 0; p < 4; p++)
      {
        const int dy = p * 8 + warp, dx = lane;
        bool m = false;

        if (dx < 16 && dy < 16)
          m = (rows[dy] >> dx) & 1u;
        else if (x + dx < a.BX && y + dy < a.BY)
          m = predicate_thread<CH>(s, a.rec[(size_t)(y + dy) * a.BX + x + dx]);

        const uint32_t b = __ballot_sync(0xFFFFFFFFu, m);
        __syncthreads(); // every read of rows[dy] (dy < 16) of this pass happened

        if (lane == 0)
          rows[dy] = b;

        __syncthreads();
      }

      if (threadIdx.x == 0)
      {
        int rx, ry;
        sHit = expand_unmasked(rows, 32, x, y, a.BX, a.BY, rx, ry) ? 1 : 0;
        sRx = rx;
        sRy = ry;
      }

      __syncthreads();
    }

    if (threadIdx.x < 32)
      a.extBits[(size_t)slot * 32 + threadIdx.x] = rows[threadIdx.x];

    if (threadIdx.x == 0)
    {
      a.extSlot[seed] = slot | (size == 32 ? 0x80000000u : 0u);

    }

    __syncthreads();
  }
}

template <int CH>
__global__ void __launch_bounds__(256) k_plan_centres(PlanArgs a)
{
  const uint32_t count = min(a.counters[1], a.ctrCap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (uint32_t slot = blockIdx.x; slot < count; slot += gridDim.x)
  {
    const uint4 h = a.ctrHdr[slot];
    const int rgX = h.y & 0xFFFF, rgY = h.y >> 16, rgW = h.z & 0xFFFF, rgH = h.z >> 16;
    const PredRec c = a.rec[h.w];

    for (int row = warp; row < 32; row += 8)
    {
      bool m = false;

      if (row < rgH && lane < rgW)
        m = predicate_thread<CH>(c, a.rec[(size_t)(rgY + row) * a.BX + rgX + lane]);

      const uint32_t b = __ballot_sync(0xFFFFFFFFu, m);

      if (lane == 0)
        a.ctrBits[(size_t)slot * 32 + row] = b;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// banded scan
//
// The reference's scan is sequential only through the in-use mask. The block rows are cut into bands, one CTA per band.
// Band k replays the reference's scan over ITS rows against an input mask that holds the rectangles emitted by all bands
// above it. All bands run concurrently from the previous iteration's rectangles; a band re-runs only when the input mask
// changed inside the row range its last run actually read. Band 0 never depends on anything, so after iteration t bands
// 0..t are final; in practice influence dies out after a few rows and a handful of iterations suffice. At the fixed point
// every band saw exactly the mask the sequential scan would have shown it, so the emission lists, concatenated in band
// order, ARE the reference's emission order. Stage 1 (remaining merges) repeats the procedure on top of the final stage-0 mask.
// ---------------------------------------------------------------------------------------------

#define LIMG_MERGE_THREADS 1024
#define LIMG_MERGE_WARPS (LIMG_MERGE_THREADS / 32)
#define LIMG_MERGE_MAX_BANDS 128
#define LIMG_REGION_MAX 32

struct MergeMailbox
{
  int kind; // 0 quit, 1 evaluate strip (AND of predicates), 2 evaluate region bitmap
  int seed;
  int x0, y0, w, h;
  int result;
  uint32_t regionBits[LIMG_REGION_MAX]; // kind 2: bit (row, col) = block unused-or-unknown AND matches
};

__device__ __forceinline__ void named_bar_sync(int id, int count)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

struct MergeArgs
{
  const PredRec *rec;
  const uint32_t *window;
  const uint32_t *extSlot, *extBits, *ctrSlot, *ctrBits, *sym;
  const uint4 *ctrHdr;
  int BX, BY, wordsPerRow;
  int bandRows, numBands, listCap;
  uint2 *lists;          // [numBands][2][listCap]: (ox | oy << 16, rx | ry << 16)
  uint32_t *counts;      // [numBands][2]
  uint32_t *snapshot;    // [numBands][BY * wordsPerRow]: input mask of the band's last run
  uint32_t *sync;        // [0] barrier count, [1] barrier generation, [2] error flag, [8 + stage * (MAX_BANDS + 2) + iter] dirty flags
  limgcu_area *areas;
  uint32_t *mergedCount; // number of stage 0 + stage 1 areas
  uint32_t *usedOut;     // BY * wordsPerRow words: in-use mask after both merge stages (zeroed by the host)
  uint32_t *stats;       // [8] optional counters: iterations stage 0 / 1, band runs stage 0 / 1
};

__device__ __forceinline__ void grid_barrier(uint32_t *sync, uint32_t numBlocks)
{
  __syncthreads();

  if (threadIdx.x == 0)
  {
    volatile uint32_t *gen = sync + 1;
    const uint32_t g = *gen;
    __threadfence();

    if (atomicAdd(sync, 1u) == numBlocks - 1)
    {
      sync[0] = 0;
      __threadfence();
      *gen = g + 1;
    }
    else
    {
      while (*gen == g) { }
    }

    __threadfence();
  }

  __syncthreads();
}

template <int CH>
struct MergeScan
{
  const MergeArgs &a;
  uint32_t *used;          // shared: BY rows of wordsPerRow words
  const uint32_t *winBand; // shared: window words of the band's rows
  const uint32_t *extBand, *ctrBand; // shared: speculative-window slots of the band's seeds
  MergeMailbox *mail;
  int lane;
  int bandY0, bandY1;
  int readLo, readHi;      // rows whose in-use bits this run consulted
  unsigned long long tSearch = 0, tGrow = 0, tPost = 0, tFour = 0;
  uint32_t nSeeds = 0, nPost1 = 0, nPost2 = 0, nFour = 0, nPlanned = 0, nPostInFour = 0, nExtSeeds = 0;
  bool inFour = false;
  int curS = 0;
  uint32_t nPostS8 = 0, nPostS16 = 0, nPostS32 = 0, nBigSeeds = 0;

  __device__ __forceinline__ void touch_rows(int lo, int hi)
  {
    readLo = min(readLo, max(lo, 0));
    readHi = max(readHi, min(hi, a.BY - 1));
  }

  __device__ __forceinline__ uint32_t used_bits8(int x, int y) const
  {
    if (y >= a.BY)
      return 0xFFu;

    const uint32_t *row = used + (size_t)y * a.wordsPerRow;
    const int w0 = x >> 5, s = x & 31;
    return __funnelshift_r(row[w0], row[w0 + 1], s) & 0xFFu;
  }

  __device__ __forceinline__ uint32_t used_bits32(int x, int y) const
  {
    if (y >= a.BY)
      return 0xFFFFFFFFu;

    const uint32_t *row = used + (size_t)y * a.wordsPerRow;
    const int w0 = x >> 5, s = x & 31;
    return __funnelshift_r(row[w0], row[w0 + 1], s);
  }

  __device__ __forceinline__ bool is_used(int x, int y) const
  {
    return (used[(size_t)y * a.wordsPerRow + (x >> 5)] >> (x & 31)) & 1u;
  }

  __device__ __forceinline__ uint64_t match_word(int x, int y) const
  {
    const uint32_t *wp = winBand + (size_t)((y - bandY0) * a.BX + x) * 2;
    return (uint64_t)wp[0] | ((uint64_t)wp[1] << 32);
  }

  __device__ bool strip_unused(int x0, int y0, int w, int h)
  {
    bool any = false;
    touch_rows(y0, y0 + h - 1);

    for (int e = lane; e < w * h; e += 32)
    {
      const int yy = y0 + e / w, xx = x0 + e % w;
      any |= is_used(xx, yy);
    }

    return !__any_sync(0xFFFFFFFFu, any);
  }

  __device__ void post(int kind, int seed, int x0, int y0, int w, int h)
  {
    if (lane == 0)
    {
      mail->kind = kind;
      mail->seed = seed;
      mail->x0 = x0; mail->y0 = y0; mail->w = w; mail->h = h;
      mail->result = 1;
    }

    if (kind == 2 && lane < LIMG_REGION_MAX)
      mail->regionBits[lane] = 0;

    const long long t0 = clock64();
    __syncwarp();
    named_bar_sync(1, LIMG_MERGE_THREADS);
    serve(a, mail, used, 0);
    named_bar_sync(2, LIMG_MERGE_THREADS);
    tPost += clock64() - t0;
    if (kind == 1) { nPost1++; if (inFour) nPostInFour++; else if (curS == 8) nPostS8++; else if (curS == 16) nPostS16++; else nPostS32++; } else nPost2++;
  }

  // executed by every warp of the CTA for the posted request
  static __device__ void serve(const MergeArgs &a, MergeMailbox *mail, const uint32_t *used, int warp)
  {
    const int count = mail->w * mail->h;
    const int w = mail->w;
    const int kind = mail->kind;
    const PredRec seed = a.rec[mail->seed];

    for (int e = warp; e < count; e += LIMG_MERGE_WARPS)
    {
      const int ry = e / w, rx = e - ry * w;
      const int yy = mail->y0 + ry, xx = mail->x0 + rx;

      if (kind == 1)
      {
        if (*(volatile int *)&mail->result == 0)
          break;

        if (!predicate_warp<CH>(seed, a.rec[(size_t)yy * a.BX + xx]) && (threadIdx.x & 31) == 0)
          atomicAnd(&mail->result, 0);
      }
      else
      {
        // region bitmap: blocks already in use can never join, skip their predicate
        const bool inUse = (used[(size_t)yy * a.wordsPerRow + (xx >> 5)] >> (xx & 31)) & 1u;

        if (!inUse && predicate_warp<CH>(seed, a.rec[(size_t)yy * a.BX + xx]) && (threadIdx.x & 31) == 0)
          atomicOr(&mail->regionBits[ry], 1u << rx);
      }
    }
  }

  __device__ bool strip_matches(int seed, int x0, int y0, int w, int h)
  {
    post(1, seed, x0, y0, w, h);
    return mail->result != 0;
  }

  __device__ bool strip_joins(int seed, int x0, int y0, int w, int h)
  {
    return strip_unused(x0, y0, w, h) && strip_matches(seed, x0, y0, w, h);
  }

  // region cache of the four-way regrowth: one parallel request evaluates every predicate of the neighbourhood at once
  int rgX, rgY, rgW, rgH;

  __device__ bool strip_joins_cached(int seed, int x0, int y0, int w, int h)
  {
    if (x0 >= rgX && y0 >= rgY && x0 + w <= rgX + rgW && y0 + h <= rgY + rgH)
    {
      if (!strip_unused(x0, y0, w, h))
        return false;

      bool ok = true;

      if (lane < h)
      {
        const uint32_t m = (w >= 32 ? 0xFFFFFFFFu : ((1u << w) - 1u)) << (x0 - rgX);
        ok = (mail->regionBits[y0 - rgY + lane] & m) == m;
      }

      return __all_sync(0xFFFFFFFFu, ok);
    }

    return strip_joins(seed, x0, y0, w, h);
  }

  // four-way alternating growth (limg.cpp:1294-1388); the seed is the rectangle's top-left block at entry.
  __device__ void grow_four_way(int &ox, int &oy, int &rx, int &ry, int hintX, int hintY, int hintW, int hintH)
  {
    const int seed = oy * a.BX + ox;
    // the centre seed's symmetric 16 x 16 match window covers [ox - 8, ox + 8) x [oy - 8, oy + 8); strips that leave it are
    // evaluated on demand
    rgX = ox - 8;
    rgY = oy - 8;
    rgW = 16;
    rgH = 16;
    touch_rows(rgY, rgY + rgH - 1);
    (void)hintX; (void)hintY; (void)hintW; (void)hintH;

    if (lane < 16)
    {
      const uint32_t w = __ldg(&a.sym[(size_t)seed * 8 + (lane >> 1)]);
      mail->regionBits[lane] = (w >> (16 * (lane & 1))) & 0xFFFFu;
    }

    __syncwarp();
    nPlanned++;

    bool right = true, down = true, up = true, left = true;

    while (right || down || up || left)
    {
      if (right)
      {
        if (ox + rx + 1 < a.BX && strip_joins_cached(seed, ox + rx, oy, 1, ry)) rx++; else right = false;
      }

      if (down)
      {
        if (oy + ry + 1 < a.BY && strip_joins_cached(seed, ox, oy + ry, rx, 1)) ry++; else down = false;
      }

      if (up)
      {
        if (oy > 0 && strip_joins_cached(seed, ox, oy - 1, rx, 1)) { oy--; ry++; } else up = false;
      }

      if (left)
      {
        if (ox > 0 && strip_joins_cached(seed, ox - 1, oy, 1, ry)) { ox--; rx++; } else left = false;
      }
    }
  }

  // right/down growth of a 1x1 seed. The seed's match bitmap (8x8 word, or the speculative 16x16 / 32x32 extension) lives one
  // row per lane; growth inside it is ballots and shuffles, strips beyond it are evaluated on demand.
  __device__ void grow_seed(int x, int y, int &rx, int &ry)
  {
    const int seed = y * a.BX + x;
    const uint32_t slot = extBand[(y - bandY0) * a.BX + x];
    int S = 8;
    uint32_t rowBits;

    if (slot == LIMG_NO_SLOT)
    {
      const uint64_t match = match_word(x, y);
      rowBits = lane < 8 ? (uint32_t)(match >> (8 * lane)) & 0xFFu : 0u;
    }
    else
    {
      S = (slot >> 31) ? 32 : 16;
      nExtSeeds++;
      rowBits = lane < S ? __ldg(&a.extBits[(size_t)(slot & 0x7FFFFFFFu) * 32 + lane]) : 0u;
    }

    const uint32_t avail = lane < S ? (rowBits & ~used_bits32(x, y + lane)) : 0u;
    curS = S;
    bool right = true, down = true;
    rx = 1;
    ry = 1;

    while (right || down)
    {
      if (right)
      {
        bool ok = x + rx + 1 < a.BX;

        if (ok)
        {
          if (rx < S)
          {
            const int rows = min(ry, S);
            const uint32_t need = rows >= 32 ? 0xFFFFFFFFu : ((1u << rows) - 1u);
            const uint32_t have = __ballot_sync(0xFFFFFFFFu, (avail >> rx) & 1u);
            ok = (have & need) == need;

            if (ok && ry > S)
              ok = strip_joins(seed, x + rx, y + S, 1, ry - S);
          }
          else
          {
            ok = strip_joins(seed, x + rx, y, 1, ry);
          }
        }

        if (ok) rx++; else right = false;
      }

      if (down)
      {
        bool ok = y + ry + 1 < a.BY;

        if (ok)
        {
          if (ry < S)
          {
            const int cols = min(rx, S);
            const uint32_t need = cols >= 32 ? 0xFFFFFFFFu : ((1u << cols) - 1u);
            const uint32_t rowv = __shfl_sync(0xFFFFFFFFu, avail, ry);
            ok = (rowv & need) == need;

            if (ok && rx > S)
              ok = strip_joins(seed, x + S, y + ry, rx - S, 1);
          }
          else
          {
            ok = strip_joins(seed, x, y + ry, rx, 1);
          }
        }

        if (ok) ry++; else down = false;
      }
    }

    if (rx > 8 || ry > 8) nBigSeeds++;
    touch_rows(y, y + min(ry, S - 1)); // bitmap rows consulted: up to the failing row (deeper rows go through strip_unused)
  }

  __device__ void mark_used(int ox, int oy, int rx, int ry)
  {
    for (int r = lane; r < ry; r += 32)
    {
      uint32_t *row = used + (size_t)(oy + r) * a.wordsPerRow;

      for (int xx = ox; xx < ox + rx;)
      {
        const int w0 = xx >> 5, b0 = xx & 31;
        const int cnt = min(32 - b0, ox + rx - xx);
        const uint32_t m = (cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << b0;
        row[w0] |= m;
        xx += cnt;
      }
    }

    __syncwarp();
  }

  // one stage over the band's rows. stage 0: large merges (>= 3x3, centre-third retry); stage 1: anything larger than 1x1.
  __device__ uint32_t run_band(int stage, uint2 *list)
  {
    uint32_t count = 0;
    touch_rows(bandY0, bandY1 - 1); // the candidate search reads the in-use bits of every row of the band

    for (int y = bandY0; y < bandY1; y++)
    {
      int x = 0;

      while (x < a.BX)
      {
        // next candidate seed of this row at column >= x: unused and passing the stage's necessary condition on its match word
        const long long ts0 = clock64();
        {
          int found = -1;

          for (int base = x & ~31; base < a.BX && found < 0; base += 32)
          {
            const int xx = base + lane;
            bool cand = false;

            if (xx >= x && xx < a.BX && !is_used(xx, y))
            {
              const uint32_t w0 = winBand[(size_t)((y - bandY0) * a.BX + xx) * 2];
              cand = stage == 0 ? ((w0 & 0x070707u) == 0x070707u) : ((w0 & 0x0102u) != 0);

              if (cand)
              {
                // warm L1 with the speculative bitmaps this seed may consult
                const uint32_t es = extBand[(y - bandY0) * a.BX + xx];
                if (es != LIMG_NO_SLOT)
                  asm volatile("prefetch.global.L1 [%0];" ::"l"(a.extBits + (size_t)(es & 0x7FFFFFFFu) * 32));

                const uint32_t cs = stage == 0 ? ctrBand[(y - bandY0) * a.BX + xx] : LIMG_NO_SLOT;
                if (cs != LIMG_NO_SLOT)
                {
                  asm volatile("prefetch.global.L1 [%0];" ::"l"(a.ctrHdr + cs));
                  asm volatile("prefetch.global.L1 [%0];" ::"l"(a.ctrBits + (size_t)cs * 32));
                }
              }
            }

            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, cand);

            if (ballot)
              found = base + __ffs(ballot) - 1;
          }

          tSearch += clock64() - ts0;

          if (found < 0)
            break;

          x = found;
        }

        int rx, ry;
        nSeeds++;
        const long long tg0 = clock64();
        grow_seed(x, y, rx, ry);
        tGrow += clock64() - tg0;

        int eox = x, eoy = y, erx = rx, ery = ry;
        bool take = false, rescan = false;

        if (stage == 0)
        {
          if (rx >= 3 && ry >= 3) // Q4
          {
            int cox = x + rx / 3, coy = y + ry / 3, crx = rx / 3, cry = ry / 3;
            const long long tf0 = clock64();
            inFour = true;
            grow_four_way(cox, coy, crx, cry, x, y, rx, ry);
            inFour = false;
            tFour += clock64() - tf0;
            nFour++;

            if (crx * cry > rx * ry)
            {
              eox = cox; eoy = coy; erx = crx; ery = cry;
              rescan = true;
            }

            take = true;
          }
        }
        else
        {
          take = rx > 1 || ry > 1;
        }

        if (!take)
        {
          x++;
          continue;
        }

        mark_used(eox, eoy, erx, ery);

        if (lane == 0)
        {
          if (count < (uint32_t)a.listCap)
            list[count] = make_uint2((uint32_t)eox | ((uint32_t)eoy << 16), (uint32_t)erx | ((uint32_t)ery << 16));
          else
            a.sync[2] = 1; // list overflow: reported by the host as LIMGCU_ERROR_OUT_OF_BOUNDS
        }

        count++;

        if (!rescan)
          x += rx;
      }
    }

    if (lane == 0 && a.stats)
    {
      atomicAdd(&a.stats[4], nSeeds); atomicAdd(&a.stats[5], nPost1); atomicAdd(&a.stats[6], nPost2); atomicAdd(&a.stats[7], nFour); atomicAdd(&a.stats[12], nPlanned); atomicAdd(&a.stats[13], nPostInFour); atomicAdd(&a.stats[14], nExtSeeds); atomicAdd(&a.stats[15], nPostS8); atomicAdd(&a.stats[1], nPostS16); atomicAdd(&a.stats[0], nPostS32); atomicAdd(&a.stats[3], nBigSeeds);
      atomicAdd(&a.stats[8], (uint32_t)(tSearch >> 10)); atomicAdd(&a.stats[9], (uint32_t)(tGrow >> 10)); atomicAdd(&a.stats[10], (uint32_t)(tPost >> 10)); atomicAdd(&a.stats[11], (uint32_t)(tFour >> 10));
    }

    return min(count, (uint32_t)a.listCap);
  }
};

__device__ __forceinline__ void or_rect(uint32_t *mask, int wordsPerRow, uint2 r, bool atomic)
{
  const int ox = r.x & 0xFFFF, oy = r.x >> 16, rx = r.y & 0xFFFF, ry = r.y >> 16;

  for (int row = 0; row < ry; row++)
  {
    uint32_t *m = mask + (size_t)(oy + row) * wordsPerRow;

    for (int xx = ox; xx < ox + rx;)
    {
      const int w0 = xx >> 5, b0 = xx & 31;
      const int cnt = min(32 - b0, ox + rx - xx);
      const uint32_t bits = (cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << b0;
      atomicOr(&m[w0], bits);
      xx += cnt;
    }
  }

  (void)atomic;
}

template <int CH>
__global__ void __launch_bounds__(LIMG_MERGE_THREADS) k_merge_banded(MergeArgs a)
{
  extern __shared__ __align__(16) unsigned char dynSmem[];
  uint32_t *used = reinterpret_cast<uint32_t *>(dynSmem);
  const int maskWords = a.BY * a.wordsPerRow;
  uint32_t *winBand = used + maskWords;
  uint32_t *extBand = winBand + (size_t)a.bandRows * a.BX * 2;
  uint32_t *ctrBand = extBand + (size_t)a.bandRows * a.BX;
  __shared__ MergeMailbox mail;
  __shared__ int sDirty;
  __shared__ int sRange[2];
  __shared__ uint32_t sCount;

  const int k = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bandY0 = k * a.bandRows, bandY1 = min(a.BY, bandY0 + a.bandRows);
  uint32_t *snapshot = a.snapshot + (size_t)k * maskWords;

  // the band's match words never change: shared memory
  for (int i = threadIdx.x; i < (bandY1 - bandY0) * a.BX * 2; i += blockDim.x)
    winBand[i] = a.window[(size_t)bandY0 * a.BX * 2 + i];

  for (int i = threadIdx.x; i < (bandY1 - bandY0) * a.BX; i += blockDim.x)
  {
    extBand[i] = a.extSlot[(size_t)bandY0 * a.BX + i];
    ctrBand[i] = a.ctrSlot[(size_t)bandY0 * a.BX + i];
  }

  for (int stage = 0; stage < 2; stage++)
  {
    bool ran = false;
    int readLo = a.BY, readHi = -1;
    uint2 *myList = a.lists + ((size_t)k * 2 + stage) * a.listCap;
    uint32_t *dirtyFlags = a.sync + 8 + stage * (LIMG_MERGE_MAX_BANDS + 2);

    for (int iter = 0; iter <= a.numBands; iter++)
    {
      // ---- phase A: input mask = rectangles of every band above (this stage) [+ all of stage 0 when in stage 1]
      for (int i = threadIdx.x; i < maskWords; i += blockDim.x)
        used[i] = 0;

      __syncthreads();

      if (stage == 1)
      {
        for (int j = 0; j < a.numBands; j++)
        {
          const uint32_t n = a.counts[j * 2 + 0];
          const uint2 *l = a.lists + ((size_t)j * 2 + 0) * a.listCap;

          for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            or_rect(used, a.wordsPerRow, l[i], true);
        }
      }

      for (int j = 0; j < k; j++)
      {
        const uint32_t n = a.counts[j * 2 + stage];
        const uint2 *l = a.lists + ((size_t)j * 2 + stage) * a.listCap;

        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
          or_rect(used, a.wordsPerRow, l[i], true);
      }

      if (threadIdx.x == 0)
        sDirty = ran ? 0 : 1;

      __syncthreads();

      if (ran && readHi >= readLo)
      {
        bool diff = false;

        for (int i = readLo * a.wordsPerRow + threadIdx.x; i < (readHi + 1) * a.wordsPerRow; i += blockDim.x)
          diff |= used[i] != snapshot[i];

        if (diff)
          sDirty = 1;
      }

      __syncthreads();
      const bool dirty = sDirty != 0;

      if (dirty)
      {
        for (int i = threadIdx.x; i < maskWords; i += blockDim.x)
          snapshot[i] = used[i];
      }

      grid_barrier(a.sync, gridDim.x); // every band has read the lists of the previous iteration

      // ---- phase B: dirty bands replay their rows
      if (dirty)
      {
        if (warp == 0)
        {
          MergeScan<CH> scan{ a, used, winBand, extBand, ctrBand, &mail, lane, bandY0, bandY1, a.BY, -1 };
          const uint32_t count = scan.run_band(stage, myList);

          if (lane == 0)
          {
            sCount = count;
            sRange[0] = scan.readLo;
            sRange[1] = scan.readHi;
            mail.kind = 0;
            dirtyFlags[iter] = 1;

            if (a.stats)
              atomicAdd(&a.stats[2 + stage], 1u);
          }

          __syncwarp();
          named_bar_sync(1, LIMG_MERGE_THREADS); // release the helpers
        }
        else
        {
          while (true)
          {
            named_bar_sync(1, LIMG_MERGE_THREADS);

            if (mail.kind == 0)
              break;

            MergeScan<CH>::serve(a, &mail, used, warp);
            named_bar_sync(2, LIMG_MERGE_THREADS);
          }
        }

        __syncthreads();
        ran = true;
        readLo = sRange[0];
        readHi = sRange[1];

        if (threadIdx.x == 0)
        {
          a.counts[k * 2 + stage] = sCount;
          __threadfence();
        }
      }

      grid_barrier(a.sync, gridDim.x); // lists of this iteration are complete

      if (*(volatile uint32_t *)&dirtyFlags[iter] == 0)
      {
        if (k == 0 && threadIdx.x == 0 && a.stats)
          a.stats[stage] = iter;

        break;
      }
    }
  }

  // ---- emission order = band order, stage 0 then stage 1; in-use mask for the leftover pass
  uint32_t before0 = 0, before1 = 0, total0 = 0, total1 = 0;

  for (int j = 0; j < a.numBands; j++)
  {
    const uint32_t c0 = a.counts[j * 2 + 0], c1 = a.counts[j * 2 + 1];

    if (j < k)
    {
      before0 += c0;
      before1 += c1;
    }

    total0 += c0;
    total1 += c1;
  }

  for (int stage = 0; stage < 2; stage++)
  {
    const uint32_t n = a.counts[k * 2 + stage];
    const uint2 *l = a.lists + ((size_t)k * 2 + stage) * a.listCap;
    const uint32_t base = stage == 0 ? before0 : total0 + before1;

    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
      const uint2 r = l[i];
      limgcu_area *out = &a.areas[base + i];
      out->ox = r.x & 0xFFFF; out->oy = r.x >> 16; out->rx = r.y & 0xFFFF; out->ry = r.y >> 16;
      out->stage = stage;
      or_rect(a.usedOut, a.wordsPerRow, r, true);
    }
  }

  if (k == 0 && threadIdx.x == 0)
    *a.mergedCount = total0 + total1;
}

// ---------------------------------------------------------------------------------------------
// after the scan: leftovers, geometry, block map, size classes (single CTA, 1024 threads)
// ---------------------------------------------------------------------------------------------

struct PrepareArgs
{
  int W, H, BX, BY, wordsPerRow;
  limgcu_area *areas;
  const uint32_t *mergedCount;
  const uint32_t *used;
  uint32_t *areaCount;
  uint32_t *blockToArea;
  AreaWork *work;
  uint32_t *smallList, *largeList;
  uint32_t *smallCount, *largeCount;
  int noMerge; // every block is its own area (limg_encode3d_test): mergedCount is ignored
};

__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *warpSums /* [33] */, uint32_t &total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;

#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += n;
  }

  if (lane == 31)
    warpSums[warp] = incl;

  __syncthreads();

  if (warp == 0)
  {
    uint32_t s = warpSums[lane];
    uint32_t si = s;

#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, si, o);
      if (lane >= o) si += n;
    }

    warpSums[lane] = si - s;

    if (lane == 31)
      warpSums[32] = si;
  }

  __syncthreads();
  const uint32_t r = warpSums[warp] + incl - v;
  total = warpSums[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024) k_area_prepare(PrepareArgs a)
{
  __shared__ uint32_t warpSums[33];
  __shared__ uint32_t sSmall, sLarge;
  const int nBlocks = a.BX * a.BY;
  const uint32_t merged = a.noMerge ? 0u : *a.mergedCount;

  if (threadIdx.x == 0)
  {
    sSmall = 0;
    sLarge = 0;
  }

  // 1. leftovers in raster order (limg.cpp:1860-1878)
  uint32_t carry = merged;

  for (int base = 0; base < nBlocks; base += 1024)
  {
    const int b = base + threadIdx.x;
    uint32_t isLeft = 0;
    int by = 0, bx = 0;

    if (b < nBlocks)
    {
      by = b / a.BX;
      bx = b - by * a.BX;
      isLeft = a.noMerge ? 1u : (((a.used[(size_t)by * a.wordsPerRow + (bx >> 5)] >> (bx & 31)) & 1u) ^ 1u);
    }

    uint32_t total;
    const uint32_t pos = block_exclusive_scan_1024(isLeft, warpSums, total);

    if (isLeft)
    {
      limgcu_area *out = &a.areas[carry + pos];
      out->ox = bx; out->oy = by; out->rx = 1; out->ry = 1;
      out->stage = 2;
    }

    carry += total;
  }

  const uint32_t count = carry;

  if (threadIdx.x == 0)
    *a.areaCount = count;

  __syncthreads();

  // 2. geometry, block map, size class, scratch offsets
  uint32_t offCarry = 0;

  for (uint32_t base = 0; base < count; base += 1024)
  {
    const uint32_t k = base + threadIdx.x;
    uint32_t n = 0;

    if (k < count)
    {
      limgcu_area *ar = &a.areas[k];
      const uint32_t ox = ar->ox, oy = ar->oy, rx = ar->rx, ry = ar->ry;
      uint32_t pw = rx * LIMG_BLOCK, ph = ry * LIMG_BLOCK;

      if (ox + rx == (uint32_t)a.BX && (a.W % LIMG_BLOCK)) pw = pw - LIMG_BLOCK + a.W % LIMG_BLOCK; // limg.cpp:1725-1739
      if (oy + ry == (uint32_t)a.BY && (a.H % LIMG_BLOCK)) ph = ph - LIMG_BLOCK + a.H % LIMG_BLOCK;

      ar->px_x = ox * LIMG_BLOCK; ar->px_y = oy * LIMG_BLOCK; ar->px_w = pw; ar->px_h = ph;
      n = pw * ph;

      for (uint32_t yy = oy; yy < oy + ry; yy++)
        for (uint32_t xx = ox; xx < ox + rx; xx++)
          a.blockToArea[(size_t)yy * a.BX + xx] = k;

      if (n <= LIMG_SMALL_AREA_PX)
        a.smallList[atomicAdd(&sSmall, 1u)] = k;
      else
        a.largeList[atomicAdd(&sLarge, 1u)] = k;
    }

    uint32_t total;
    const uint32_t off = block_exclusive_scan_1024(n, warpSums, total);

    if (k < count)
    {
      a.work[k].n = n;
      a.work[k].scratchOff = offCarry + off;
    }

    offCarry += total;
  }

  __syncthreads();

  if (threadIdx.x == 0)
  {
    *a.smallCount = sSmall;
    *a.largeCount = sLarge;
  }
}

} // namespace limg

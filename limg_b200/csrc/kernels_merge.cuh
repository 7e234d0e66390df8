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
  uint32_t *counters;  // [0] ext slots
  uint32_t extCap;
  uint16_t *unmasked;  // per block: rx | ry << 8 of the mask-free right/down growth inside the best available bitmap
  uint32_t *candBits;  // [2][BY][wordsPerRow] (zeroed by the host): blocks that can emit in stage 0 (3x3 corner matches) / stage 1 (right or lower neighbour matches)
  uint32_t *candList;  // [2][blocks] the same as lists, counts in candCount[2]
  uint32_t *candCount;
  int wordsPerRow;
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

__global__ void __launch_bounds__(256) k_plan_seeds(PlanArgs a)
{
  const int seed = blockIdx.x * blockDim.x + threadIdx.x;
  const bool inside = seed < a.BX * a.BY;
  const int y = inside ? seed / a.BX : 0, x = inside ? seed - y * a.BX : 0;
  const uint32_t w0 = inside ? a.window[(size_t)seed * 2] : 0u, w1 = inside ? a.window[(size_t)seed * 2 + 1] : 0u;

  if (a.candBits)
  {
    // necessary conditions (mask-free) for the seed to emit anything: stage 0 needs a 3x3 rectangle, stage 1 one neighbour
    const bool c[2] = { inside && (w0 & 0x070707u) == 0x070707u, inside && (w0 & 0x0102u) != 0u };
    const int lane = threadIdx.x & 31;

#pragma unroll
    for (int st = 0; st < 2; st++)
    {
      const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, c[st]);
      uint32_t pos = 0;

      if (lane == 0 && ballot)
        pos = atomicAdd(&a.candCount[st], (uint32_t)__popc(ballot));

      pos = __shfl_sync(0xFFFFFFFFu, pos, 0) + __popc(ballot & ((1u << lane) - 1u));

      if (c[st])
      {
        a.candList[(size_t)st * a.BX * a.BY + pos] = (uint32_t)seed;
        atomicOr(&a.candBits[((size_t)st * a.BY + y) * a.wordsPerRow + (x >> 5)], 1u << (x & 31));
      }
    }
  }

  if (!inside)
    return;

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
  a.unmasked[seed] = (uint16_t)(rx | (ry << 8));

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
      a.unmasked[seed] = (uint16_t)(sRx | (sRy << 8));

    }

    __syncthreads();
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

#define LIMG_MERGE_THREADS 512
#define LIMG_MERGE_WARPS (LIMG_MERGE_THREADS / 32)
#define LIMG_MERGE_MAX_BANDS 128
#define LIMG_ROW_CHUNK 128 // candidate seeds of one block row expanded concurrently

struct MergeArgs
{
  const PredRec *rec;
  const uint32_t *window;
  const uint32_t *extSlot, *extBits, *sym;
  const uint16_t *unmasked;
  int BX, BY, wordsPerRow;
  int bandRows, numBands, listCap;
  int rowChunk;          // candidate seeds of one block row expanded concurrently (<= LIMG_ROW_CHUNK)
  uint2 *lists;          // [numBands][2][listCap]: (ox | oy << 16, rx | ry << 16)
  uint32_t *counts;      // [numBands][2]
  uint32_t *snapshot;    // [numBands][BY * wordsPerRow]: input mask of the band's last run
  uint32_t *sync;        // [0] barrier count, [1] barrier generation, [2] error flag, [8 + stage * (MAX_BANDS + 2) + iter] dirty flags
  limgcu_area *areas;
  uint32_t *mergedCount; // number of stage 0 + stage 1 areas
  uint32_t *usedOut;     // BY * wordsPerRow words: in-use mask after both merge stages (zeroed by the host)
  uint32_t *stats;       // [24] optional counters
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

// what one seed would do against a given in-use mask
struct SeedResult
{
  short x, rx, ry;       // right/down rectangle grown from the seed
  short kind;            // 0 nothing to emit, 1 emit the right/down rectangle, 2 emit the centre-third regrowth (and examine the seed again)
  short cox, coy, crx, cry; // four-way regrowth (valid when attempted)
  short attempted;
  short pad;
};

struct BandShared
{
  SeedResult res[LIMG_ROW_CHUNK];
  short cand[LIMG_ROW_CHUNK];
  uint2 accepted[LIMG_ROW_CHUNK + 32]; // rectangles committed in the current chunk round
  int nCand, nextX;
  int rangeLo, rangeHi;
  uint32_t count;
  int dirty;
};

// A band's scan. Every method is warp-cooperative: all 32 lanes call it with warp-uniform arguments. The in-use mask is only
// modified by commit() (warp 0, between barriers); expansions read it.
template <int CH>
struct BandScan
{
  const MergeArgs &a;
  uint32_t *used;          // shared: BY rows of wordsPerRow words
  const uint32_t *winBand; // shared: match words of the band's rows
  const uint32_t *extBand; // shared: extension-bitmap slots of the band's seeds
  BandShared *sh;
  int lane;
  int bandY0, bandY1;
  int readLo, readHi;      // rows whose in-use bits this warp consulted
  uint32_t nOnDemand, nSeeds, nFour, nInline;
  bool cheapOnly, aborted; // a speculative expansion of a probably-swallowed candidate gives up instead of evaluating predicates on demand

  __device__ __forceinline__ void touch_rows(int lo, int hi)
  {
    readLo = min(readLo, max(lo, 0));
    readHi = max(readHi, min(hi, a.BY - 1));
  }

  __device__ __forceinline__ uint32_t used_bits32(int x, int y) const
  {
    // 32 in-use bits of row y starting at column x (x may be negative); everything outside the grid reads as in use
    if (y < 0 || y >= a.BY)
      return 0xFFFFFFFFu;

    const uint32_t *row = used + (size_t)y * a.wordsPerRow;

    if (x < 0)
    {
      const int s = -x; // 1..31
      return (row[0] << s) | ((1u << s) - 1u);
    }

    const int w0 = x >> 5, s = x & 31;
    return __funnelshift_r(row[w0], row[w0 + 1], s);
  }

  __device__ __forceinline__ bool is_used(int x, int y) const
  {
    return (used[(size_t)y * a.wordsPerRow + (x >> 5)] >> (x & 31)) & 1u;
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

  // every block of the strip matches the seed? evaluated by this warp alone: short strips one predicate at a time with the 27
  // samples spread over the lanes, long strips one predicate per lane.
  __device__ bool strip_matches(const PredRec &seed, int x0, int y0, int w, int h)
  {
    const int count = w * h;

    if (cheapOnly)
    {
      aborted = true;
      return false;
    }

    nOnDemand += count;

    if (count >= 6)
    {
      bool ok = true;

      for (int base = 0; base < count && ok; base += 32)
      {
        const int e = base + lane;
        bool m = true;

        if (e < count)
        {
          const int yy = y0 + e / w, xx = x0 + e % w;
          m = predicate_thread<CH>(seed, a.rec[(size_t)yy * a.BX + xx]);
        }

        ok = __all_sync(0xFFFFFFFFu, m);
      }

      return ok;
    }

    for (int e = 0; e < count; e++)
    {
      const int yy = y0 + e / w, xx = x0 + e % w;

      if (!predicate_warp<CH>(seed, a.rec[(size_t)yy * a.BX + xx]))
        return false;
    }

    return true;
  }

  __device__ bool strip_joins(const PredRec &seed, int x0, int y0, int w, int h)
  {
    return strip_unused(x0, y0, w, h) && strip_matches(seed, x0, y0, w, h);
  }

  // right/down growth of a 1x1 seed (limg.cpp:1294-1343). The seed's match bitmap (8x8 word, or the speculative 16x16 / 32x32
  // extension) lives one row per lane; growth inside it is ballots and shuffles, strips beyond it are evaluated on demand.
  __device__ void grow_seed(int x, int y, int &rx, int &ry)
  {
    const int seed = y * a.BX + x;
    const uint32_t slot = extBand[(y - bandY0) * a.BX + x];
    int S = 8;
    uint32_t rowBits;

    if (slot == LIMG_NO_SLOT)
    {
      const uint32_t *wp = winBand + (size_t)((y - bandY0) * a.BX + x) * 2;
      rowBits = lane < 8 ? (wp[lane >> 2] >> (8 * (lane & 3))) & 0xFFu : 0u;
    }
    else
    {
      S = (slot >> 31) ? 32 : 16;
      rowBits = lane < S ? __ldg(&a.extBits[(size_t)(slot & 0x7FFFFFFFu) * 32 + lane]) : 0u;
    }

    const uint32_t avail = lane < S ? (rowBits & ~used_bits32(x, y + lane)) : 0u;
    bool right = true, down = true;
    bool haveRec = false;
    PredRec rec;
    rx = 1;
    ry = 1;

    while (right || down)
    {
      if (right)
      {
        bool ok = x + rx + 1 < a.BX;

        if (ok)
        {
          const int rows = min(ry, S);

          if (rx < S)
          {
            const uint32_t need = rows >= 32 ? 0xFFFFFFFFu : ((1u << rows) - 1u);
            const uint32_t have = __ballot_sync(0xFFFFFFFFu, (avail >> rx) & 1u);
            ok = (have & need) == need;
          }

          if (ok && (rx >= S || ry > S))
          {
            if (!haveRec) { rec = a.rec[seed]; haveRec = true; }
            ok = rx >= S ? strip_joins(rec, x + rx, y, 1, ry) : strip_joins(rec, x + rx, y + S, 1, ry - S);
          }
        }

        if (ok) rx++; else right = false;
      }

      if (down)
      {
        bool ok = y + ry + 1 < a.BY;

        if (ok)
        {
          const int cols = min(rx, S);

          if (ry < S)
          {
            const uint32_t need = cols >= 32 ? 0xFFFFFFFFu : ((1u << cols) - 1u);
            const uint32_t rowv = __shfl_sync(0xFFFFFFFFu, avail, ry);
            ok = (rowv & need) == need;
          }

          if (ok && (ry >= S || rx > S))
          {
            if (!haveRec) { rec = a.rec[seed]; haveRec = true; }
            ok = ry >= S ? strip_joins(rec, x, y + ry, rx, 1) : strip_joins(rec, x + S, y + ry, rx - S, 1);
          }
        }

        if (ok) ry++; else down = false;
      }
    }

    touch_rows(y, y + min(ry, S - 1)); // bitmap rows consulted: up to the failing row (deeper rows go through strip_unused)
  }

  // four-way alternating growth from the centre third (limg.cpp:1294-1388, 1426-1433). The centre seed's symmetric 16 x 16 match
  // window covers [ox - 8, ox + 8) x [oy - 8, oy + 8), one row per lane; strips that leave it are evaluated on demand.
  __device__ void grow_four_way(int &ox, int &oy, int &rx, int &ry)
  {
    const int seed = oy * a.BX + ox;
    const int rgX = ox - 8, rgY = oy - 8;
    uint32_t avail = 0;

    if (lane < 16)
    {
      const uint32_t w = __ldg(&a.sym[(size_t)seed * 8 + (lane >> 1)]);
      avail = ((w >> (16 * (lane & 1))) & 0xFFFFu) & ~used_bits32(rgX, rgY + lane);
    }

    touch_rows(rgY, rgY + 15);
    bool haveRec = false;
    PredRec rec;
    bool right = true, down = true, up = true, left = true;

    // strip test: inside the window -> bits, otherwise on demand
    auto joins = [&](int x0, int y0, int w, int h) -> bool {
      if (x0 >= rgX && y0 >= rgY && x0 + w <= rgX + 16 && y0 + h <= rgY + 16)
      {
        const uint32_t m = ((1u << w) - 1u) << (x0 - rgX);
        const int r0 = y0 - rgY;
        const bool rowOk = (lane < r0 || lane >= r0 + h) || ((avail & m) == m);
        return __all_sync(0xFFFFFFFFu, rowOk);
      }

      if (!haveRec) { rec = a.rec[seed]; haveRec = true; }
      return strip_joins(rec, x0, y0, w, h);
    };

    while (right || down || up || left)
    {
      if (right)
      {
        if (ox + rx + 1 < a.BX && joins(ox + rx, oy, 1, ry)) rx++; else right = false;
      }

      if (down)
      {
        if (oy + ry + 1 < a.BY && joins(ox, oy + ry, rx, 1)) ry++; else down = false;
      }

      if (up)
      {
        if (oy > 0 && joins(ox, oy - 1, rx, 1)) { oy--; ry++; } else up = false;
      }

      if (left)
      {
        if (ox > 0 && joins(ox - 1, oy, 1, ry)) { ox--; rx++; } else left = false;
      }
    }
  }

  // what seed (x, y) does against the current mask (limg.cpp:1405-1486)
  __device__ SeedResult expand(int x, int y, int stage, bool cheap = false)
  {
    SeedResult r;
    int rx, ry;
    nSeeds++;
    cheapOnly = cheap;
    aborted = false;
    grow_seed(x, y, rx, ry);
    r.x = (short)x; r.rx = (short)rx; r.ry = (short)ry;
    r.kind = 0; r.attempted = 0; r.pad = 0;
    r.cox = r.coy = r.crx = r.cry = 0;

    if (stage == 0)
    {
      if (rx >= 3 && ry >= 3) // Q4
      {
        int cox = x + rx / 3, coy = y + ry / 3, crx = rx / 3, cry = ry / 3;
        nFour++;
        grow_four_way(cox, coy, crx, cry);
        r.cox = (short)cox; r.coy = (short)coy; r.crx = (short)crx; r.cry = (short)cry;
        r.attempted = 1;
        r.kind = (crx * cry > rx * ry) ? 2 : 1;
      }
    }
    else
    {
      r.kind = (rx > 1 || ry > 1) ? 1 : 0;
    }

    if (aborted)
      r.kind = -1; // unknown: commit() expands the seed properly if it is still free when its turn comes

    cheapOnly = false;
    return r;
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

  // does any rectangle committed in this chunk round touch the inclusive box [x0, x1] x [y0, y1]?
  __device__ bool touches_accepted(int nAcc, int x0, int y0, int x1, int y1) const
  {
    bool hit = false;

    for (int j = lane; j < nAcc; j += 32)
    {
      const uint2 r = sh->accepted[j];
      const int ax = r.x & 0xFFFF, ay = r.x >> 16, aw = r.y & 0xFFFF, ah = r.y >> 16;
      hit |= ax <= x1 && ax + aw > x0 && ay <= y1 && ay + ah > y0;
    }

    return __any_sync(0xFFFFFFFFu, hit);
  }

  // warp 0: commit the chunk's expansions in scan order. A result computed against the pre-chunk mask stands unless a rectangle
  // committed earlier in this chunk touches what it probed; then (and after a centre-third hit, which re-examines the same seed)
  // the seed is expanded again against the current mask, exactly as the sequential scan would have seen it.
  __device__ void commit(int y, int stage, int nCand, uint2 *list, uint32_t &count)
  {
    int nAcc = 0;

    for (int i = 0; i < nCand; i++)
    {
      const int x = sh->cand[i];

      if (is_used(x, y))
        continue;

      SeedResult r = sh->res[i];
      bool valid = r.kind >= 0 && !touches_accepted(nAcc, x, y, x + r.rx, y + r.ry);

      if (valid && r.attempted)
        valid = !touches_accepted(nAcc, r.cox - 1, r.coy - 1, r.cox + r.crx, r.coy + r.cry);

      while (true)
      {
        if (!valid)
        {
          nInline++;
          r = expand(x, y, stage);
          valid = true;
        }

        if (r.kind == 0)
          break;

        const int eox = r.kind == 2 ? r.cox : x, eoy = r.kind == 2 ? r.coy : y;
        const int erx = r.kind == 2 ? r.crx : r.rx, ery = r.kind == 2 ? r.cry : r.ry;
        mark_used(eox, eoy, erx, ery);
        const uint2 packed = make_uint2((uint32_t)eox | ((uint32_t)eoy << 16), (uint32_t)erx | ((uint32_t)ery << 16));

        if (lane == 0)
        {
          if (count < (uint32_t)a.listCap)
            list[count] = packed;
          else
            a.sync[2] = 1; // list overflow: reported by the host as LIMGCU_ERROR_OUT_OF_BOUNDS

          if (nAcc < LIMG_ROW_CHUNK + 32)
            sh->accepted[nAcc] = packed;
        }

        __syncwarp();
        count++;
        nAcc++;

        if (nAcc >= LIMG_ROW_CHUNK + 32)
          nAcc = LIMG_ROW_CHUNK + 32; // list full: every later result of this chunk is re-expanded (still exact)

        if (r.kind == 2 && !is_used(x, y))
        {
          valid = false; // limg.cpp:1435-1438: the scan resumes at the same seed
          continue;
        }

        break;
      }

      if (nAcc >= LIMG_ROW_CHUNK + 32)
      {
        // cannot track more rectangles: fall back to sequential handling of the rest of the chunk
        for (int j = i + 1; j < nCand; j++)
          sh->res[j].attempted = 1, sh->res[j].cox = 0, sh->res[j].coy = 0, sh->res[j].crx = (short)a.BX, sh->res[j].cry = (short)a.BY;
      }
    }
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
  uint16_t *unmBand = reinterpret_cast<uint16_t *>(extBand + (size_t)a.bandRows * a.BX);
  __shared__ BandShared sh;

  const int k = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bandY0 = k * a.bandRows, bandY1 = min(a.BY, bandY0 + a.bandRows);
  uint32_t *snapshot = a.snapshot + (size_t)k * maskWords;

  // the band's match words and extension slots never change: shared memory
  for (int i = threadIdx.x; i < (bandY1 - bandY0) * a.BX * 2; i += blockDim.x)
    winBand[i] = a.window[(size_t)bandY0 * a.BX * 2 + i];

  for (int i = threadIdx.x; i < (bandY1 - bandY0) * a.BX; i += blockDim.x)
  {
    extBand[i] = a.extSlot[(size_t)bandY0 * a.BX + i];
    unmBand[i] = a.unmasked[(size_t)bandY0 * a.BX + i];
  }

  BandScan<CH> scan{ a, used, winBand, extBand, &sh, lane, bandY0, bandY1, a.BY, -1, 0, 0, 0, 0, false, false };

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
        sh.dirty = ran ? 0 : 1;

      __syncthreads();

      if (ran && readHi >= readLo)
      {
        bool diff = false;

        for (int i = readLo * a.wordsPerRow + threadIdx.x; i < (readHi + 1) * a.wordsPerRow; i += blockDim.x)
          diff |= used[i] != snapshot[i];

        if (diff)
          sh.dirty = 1;
      }

      __syncthreads();
      const bool dirty = sh.dirty != 0;

      if (dirty)
      {
        for (int i = threadIdx.x; i < maskWords; i += blockDim.x)
          snapshot[i] = used[i];
      }

      grid_barrier(a.sync, gridDim.x); // every band has read the lists of the previous iteration

      // ---- phase B: dirty bands replay their rows, one chunk of candidate seeds at a time
      if (dirty)
      {
        if (threadIdx.x == 0)
        {
          sh.count = 0;
          sh.rangeLo = bandY0;      // the candidate search reads the in-use bits of every row of the band
          sh.rangeHi = bandY1 - 1;
        }

        scan.readLo = a.BY;
        scan.readHi = -1;
        __syncthreads();

        for (int y = bandY0; y < bandY1; y++)
        {
          int x = 0;

          while (x < a.BX)
          {
            // warp 0: the next candidate seeds of this row (unused + the stage's necessary condition on the match word)
            if (warp == 0)
            {
              int n = 0, xx0 = x & ~31;

              for (; xx0 < a.BX && n < a.rowChunk; xx0 += 32)
              {
                const int xx = xx0 + lane;
                bool cand = false;

                if (xx >= x && xx < a.BX && !scan.is_used(xx, y))
                {
                  const uint32_t w0 = winBand[(size_t)((y - bandY0) * a.BX + xx) * 2];
                  cand = stage == 0 ? ((w0 & 0x070707u) == 0x070707u) : ((w0 & 0x0102u) != 0);
                }

                const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, cand);
                const int pos = n + __popc(ballot & ((1u << lane) - 1u));

                if (cand && pos < a.rowChunk)
                  sh.cand[pos] = (short)xx;

                n += __popc(ballot);
              }

              if (lane == 0)
              {
                sh.nCand = min(n, a.rowChunk);
                // row exhausted and everything taken: done after this chunk; otherwise resume behind the last candidate taken
                sh.nextX = (xx0 >= a.BX && n <= a.rowChunk) ? a.BX : -1;
              }

              __syncwarp();

              // Dry run of the scan order on the mask-free growth predicted by the plan kernels: a candidate that lies inside
              // the rectangle an earlier candidate of this chunk is predicted to emit will most likely be swallowed; it is not
              // expanded speculatively, commit() expands it on the spot in the rare case it is still free when its turn comes.
              if (lane == 0)
              {
                const int nc = min(n, a.rowChunk);
                int coverUntil = -1;

                for (int i = 0; i < nc; i++)
                {
                  const int xc = sh.cand[i];
                  const uint32_t u = unmBand[(y - bandY0) * a.BX + xc];
                  const int prx = u & 0xFF, pry = (u >> 8) & 0x7F;

                  if (xc < coverUntil)
                  {
                    // small growths are cheap to expand even if they end up swallowed; only the big ones (which would go on demand) are deferred
                    sh.res[i].kind = (prx >= 8 || pry >= 8) ? -1 : 0;
                    continue;
                  }

                  sh.res[i].kind = 0;
                  const bool emits = stage == 0 ? (prx >= 3 && pry >= 3) : (prx > 1 || pry > 1);

                  if (emits)
                    coverUntil = xc + prx;
                }
              }
            }

            __syncthreads();
            const int nCand = sh.nCand;

            if (nCand == 0)
              break;

            // every warp: expand its share of the candidates against the current mask (read only)
            for (int i = warp; i < nCand; i += LIMG_MERGE_WARPS)
            {
              // probably swallowed by an earlier candidate's rectangle: only the cheap (bitmap) part is done speculatively
              const SeedResult r = scan.expand(sh.cand[i], y, stage, sh.res[i].kind < 0);

              if (lane == 0)
                sh.res[i] = r;
            }

            __syncthreads();

            if (warp == 0)
            {
              uint32_t count = sh.count;
              scan.commit(y, stage, nCand, myList, count);

              if (lane == 0)
                sh.count = count;
            }

            const int lastCand = sh.cand[nCand - 1];
            const int nextX = sh.nextX;
            __syncthreads();
            x = nextX < 0 ? lastCand + 1 : nextX;
          }
        }

        if (lane == 0 && scan.readHi >= scan.readLo)
        {
          atomicMin(&sh.rangeLo, scan.readLo);
          atomicMax(&sh.rangeHi, scan.readHi);
        }

        __syncthreads();
        ran = true;
        readLo = sh.rangeLo;
        readHi = sh.rangeHi;

        if (threadIdx.x == 0)
        {
          a.counts[k * 2 + stage] = min(sh.count, (uint32_t)a.listCap);
          dirtyFlags[iter] = 1;

          if (a.stats)
            atomicAdd(&a.stats[2 + stage], 1u);

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

  if (lane == 0 && a.stats)
  {
    atomicAdd(&a.stats[4], scan.nSeeds);
    atomicAdd(&a.stats[5], scan.nOnDemand);
    atomicAdd(&a.stats[6], scan.nFour);
    atomicAdd(&a.stats[7], scan.nInline);
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

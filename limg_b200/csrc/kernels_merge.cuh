// limg_b200/csrc/kernels_merge.cuh -- area expansion (limg.cpp:1121-1135, 1137-1269, 1277-1496, 1814-1878).
//
// The reference's merge is a serial greedy raster scan whose only inputs are the pass-1 table and the scan order
// (SURVEY.md Q2): the predicate "candidate block matches seed block" is a pure function of two pass-1 records.
// This file holds everything that is evaluated up front, in parallel on the whole GPU, before the scan (kernels_wave.cuh):
//   1. k_pred_records : one thread per block derives the predicate-side state of its record once (the divisions).
//   2. k_pred_window  : for EVERY block as a hypothetical seed, all 63 predicates against the 8x8 window to its lower right
//                       (one thread per pair) -> one 64-bit match word per block.
//   3. k_plan_seeds   : candidate bitmaps / lists of the two merge stages, mask-free growth inside the window, and which stage-0
//                       candidates can grow out of the window (their run along the seed row or column reaches its edge).
//   4. k_plan_extend  : for those seeds a 32x32 match bitmap, filled exactly as far as ANY masked growth can reach: a right/down
//                       rectangle always contains a block of the seed's row and of its column, so it lies inside
//                       [0, run along the row] x [0, run along the column].
//   5. k_plan_centres / k_plan_sym : the four-way centre-third regrowth starts at a block that depends on the masked growth;
//                       the centres the mask-free growth predicts (and up to three to their left, which is where the mask
//                       moves them) get a match bitmap around the centre, again bounded by the four runs from the centre.
//   Everything here is a pure accelerator: the scan evaluates on demand whatever a bitmap does not cover.
//   6. k_area_prepare : leftover blocks (raster order), pixel rectangles, block->area map, size classes, scratch offsets.
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

// the 27-sample score of limg.cpp:1214-1267, every operation in the reference's order
template <int CH>
__device__ __forceinline__ float predicate_score(const PredRec &a, const PredRec &b)
{
  // Q1: the second term of every iteration projects avg(a) into b: loop invariant, but part of the ordered sum.
  const float constTerm = factor_term<CH>(a.avg, b, b.invLen);
  float sum = 0.0f;

  for (int k = 0; k < 27; k++)
  {
    sum = fadd(sum, sample_term<CH>(a, b, k));
    sum = fadd(sum, constTerm);
  }

  return fmul(sum, 1.f / (3 * 3 * 3));
}

// The three factors are an affine function of the colour (limg_factorization.h:9-41 without its roundings): f(c) = g + M c.
// `lin` evaluates the linear part only (no offsets). Contracted arithmetic is fine here: this feeds the guard-banded shortcut below.
template <int CH>
__device__ __forceinline__ void factors_affine(const float c[4], const PredRec &s, bool lin, float f[3])
{
  float t[4], est[4];
  float dA = 0.0f, dB = 0.0f, dC = 0.0f;

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    t[i] = lin ? c[i] : c[i] - s.minA[i];
    dA = fmaf(t[i], s.nA[i], dA);
  }

  f[0] = dA * s.inv[0];

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    est[i] = fmaf(f[0], s.nA[i], lin ? 0.0f : s.minA[i]);
    t[i] = (c[i] - est[i]) - (lin ? 0.0f : s.offB[i]);
    dB = fmaf(t[i], s.nB[i], dB);
  }

  f[1] = dB * s.inv[1];

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    est[i] = fmaf(f[1], s.nB[i], est[i]);
    t[i] = (c[i] - est[i]) - (lin ? 0.0f : s.offC[i]);
    dC = fmaf(t[i], s.nC[i], dC);
  }

  f[2] = dC * s.inv[2];
}

// Guard-banded shortcut for the 27-sample score. The sample colours are b.nA x/2 + b.nB y/2 + b.nC z/2, so the factors at the 27
// samples are g + x u + y v + z w with four evaluations of the affine map instead of 27: the same real-valued score with ~4x fewer
// operations and different rounding. The decision is taken from it only when the score is further from the threshold than a
// guard that is ~500 units in the last place of the largest intermediate magnitude (the two evaluations differ by a few); inside
// the guard the reference-order evaluation decides, so the predicate is still the reference's bit for bit.
// returns 1 / 0, or -1 when the exact evaluation has to decide
template <int CH>
__device__ __forceinline__ int predicate_shortcut(const PredRec &a, const PredRec &b)
{
  float zero[4] = { 0, 0, 0, 0 }, cA[4], cB[4], cC[4], g[3], u[3], v[3], w[3];

#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    cA[i] = b.nA[i] * 0.5f;
    cB[i] = b.nB[i] * 0.5f;
    cC[i] = b.nC[i] * 0.5f;
  }

  factors_affine<CH>(zero, a, false, g);
  factors_affine<CH>(cA, a, true, u);
  factors_affine<CH>(cB, a, true, v);
  factors_affine<CH>(cC, a, true, w);
  const float il0 = a.invLen[0], il1 = a.invLen[1], il2 = a.invLen[2];
  float sum = 0.0f;

#pragma unroll
  for (int z = 0; z < 3; z++)
  {
#pragma unroll
    for (int y = 0; y < 3; y++)
    {
      const float b0 = fmaf((float)z, w[0], fmaf((float)y, v[0], g[0]));
      const float b1 = 0.5f - fmaf((float)z, w[1], fmaf((float)y, v[1], g[1]));
      const float b2 = 0.5f - fmaf((float)z, w[2], fmaf((float)y, v[2], g[2]));

#pragma unroll
      for (int x = 0; x < 3; x++)
        sum += fabsf(fmaf((float)x, u[0], b0)) * il0 + fabsf(fmaf(-(float)x, u[1], b1)) * il1 + fabsf(fmaf(-(float)x, u[2], b2)) * il2;
    }
  }

  const float constTerm = factor_term<CH>(a.avg, b, b.invLen);
  const float mean = fmaf(sum, 1.f / 27.f, constTerm);
  // magnitude the roundings scale with: the largest intermediate of every term (not the possibly cancelled result)
  const float mag = il0 * (fabsf(g[0]) + 2.0f * (fabsf(u[0]) + fabsf(v[0]) + fabsf(w[0]))) + il1 * (0.5f + fabsf(g[1]) + 2.0f * (fabsf(u[1]) + fabsf(v[1]) + fabsf(w[1]))) +
                    il2 * (0.5f + fabsf(g[2]) + 2.0f * (fabsf(u[2]) + fabsf(v[2]) + fabsf(w[2]))) + fabsf(constTerm);
  const float guard = fmaf(3e-5f, mag, 1e-3f);

  if (!(guard < 0.5f)) // also catches NaN / Inf
    return -1;

  if (mean < 3.0f - guard)
    return 1;

  if (mean > 3.0f + guard)
    return 0;

  return -1;
}

// one thread evaluates the whole predicate (limg.cpp:1137-1269). a = seed, b = candidate.
template <int CH>
__device__ bool predicate_thread(const PredRec &a, const PredRec &b)
{
  const int q = predicate_quick<CH>(a, b);

  if (q >= 0)
    return q != 0;

#ifndef LIMG_EXACT_PREDICATE_ONLY
  const int s = predicate_shortcut<CH>(a, b);

  if (s >= 0)
    return s != 0;
#endif

  return predicate_score<CH>(a, b) < 3.0f;
}

// check mode (limgcu_debug_predicate_check): the reference-order evaluation and the shortcut side by side
template <int CH>
__global__ void __launch_bounds__(256) k_pred_check(const PredRec *__restrict__ rec, int BX, int BY, unsigned long long *__restrict__ out /* [4]: scored pairs, shortcut decided, disagreements, max |gap| in 1e-9 */)
{
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int seed = pair >> 1, half = pair & 1;

  if (seed >= BX * BY)
    return;

  const int o = half * 32 + (threadIdx.x & 31);
  const int dx = (o & 7) * 2 - 7, dy = (o >> 3) * 2 - 7; // a 15 x 15 neighbourhood sampled every other block, both directions
  const int sy = seed / BX, sx = seed - sy * BX;
  const int cx = sx + dx, cy = sy + dy;

  if (cx < 0 || cy < 0 || cx >= BX || cy >= BY)
    return;

  const PredRec a = rec[seed], b = rec[(size_t)cy * BX + cx];

  if (predicate_quick<CH>(a, b) >= 0)
    return;

  const float exact = predicate_score<CH>(a, b);
  const int s = predicate_shortcut<CH>(a, b);
  atomicAdd(&out[0], 1ull);

  if (s >= 0)
  {
    atomicAdd(&out[1], 1ull);

    if ((s != 0) != (exact < 3.0f))
      atomicAdd(&out[2], 1ull);
  }
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
  // A warp's 32 candidates are four runs of eight consecutive records (8 x 144 B each). Loading them lane by lane is 9 x 32 scattered
  // 16-byte requests (the kernel was bound by L1: 73 % busy against 52 % issue); the warp copies the four runs into shared memory with
  // coalesced 16-byte loads instead and every lane then reads its own record from there (stride 36 words: conflict free per quarter warp).
  constexpr int REC16 = (int)(sizeof(PredRec) / 16);
  __shared__ uint4 sCand[8][32 * REC16];
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5); // (seed, half) pairs: 8 per CTA
  const int seed = pair >> 1, half = pair & 1;

  if (seed >= BX * BY)
    return;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = half * 32 + lane;
  const int dx = o & 7, dy = o >> 3;
  const int sy = seed / BX, sx = seed - sy * BX;
  const int cols = min(8, BX - sx);
  uint4 *mine = sCand[warp];

#pragma unroll
  for (int r = 0; r < 4; r++)
  {
    const int y = sy + half * 4 + r;

    if (y < BY)
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(rec + (size_t)y * BX + sx);

      for (int i = lane; i < cols * REC16; i += 32)
        mine[r * 8 * REC16 + i] = __ldg(src + i);
    }
  }

  __syncwarp();
  bool m = false;

  if (o == 0)
    m = true;
  else if (sx + dx < BX && sy + dy < BY)
    m = predicate_thread<CH>(rec[seed], *reinterpret_cast<const PredRec *>(mine + lane * REC16));

  const uint32_t bits = __ballot_sync(0xFFFFFFFFu, m);

  if (lane == 0)
    window[(size_t)seed * 2 + half] = bits;
}

// ---------------------------------------------------------------------------------------------
// speculative match bitmaps ("regions"): 32 rows x 32 columns of predicate(seed, block) anchored at (ax, ay), one row per word,
// with a header that says which part of it is known: hdr = vx0 | vy0 << 8 | vx1 << 16 | vy1 << 24 (relative, half open).
// ---------------------------------------------------------------------------------------------

#define LIMG_NO_SLOT 0xFFFFFFFFu
#define LIMG_SLOT_PENDING 0xFFFFFFFEu // asked for, bitmap not written (yet)
#define LIMG_SYM_BACK 8  // a centre's bitmap is anchored 8 blocks up and to the left of it

struct PlanArgs
{
  const PredRec *rec;
  const uint32_t *window;
  int BX, BY, wordsPerRow;
  uint32_t *extSlot;   // per block: LIMG_NO_SLOT or its slot
  uint32_t *extSeed;   // per slot: block index
  uint32_t *extBits;   // per slot: 32 row words, anchored at the seed
  uint32_t *extHdr;    // per slot: known part
  uint32_t *symSlot, *symSeed, *symBits, *symHdr; // the same around centres, anchored at (cx - 8, cy - 8)
  uint32_t *symStart;  // per block: largest start rectangle (crx | cry << 8) a seed will regrow from this centre with
  uint32_t *counters;  // [0] ext slots, [1] sym slots
  uint32_t extCap, symCap;
  int extMaxW;         // known part of a seed's bitmap: at most this many columns (rows: 32)
  int symMaxL, symMaxR, symMaxD; // known part of a centre's bitmap: at most this far left / up, right, down of the centre
  uint16_t *unmasked;  // per block: rx | ry << 8 of the mask-free right/down growth
  uint32_t *candBits;  // [2][BY][wordsPerRow] (zeroed by the host): blocks that can emit in stage 0 (3x3 corner matches) / stage 1 (right or lower neighbour matches)
  uint32_t *candList;  // [2][blocks] the same as lists, counts in candCount[2]
  uint32_t *candCount;
};

// mask-free alternating right/down growth over `rows` (S x S match bitmap of the seed)
__device__ __forceinline__ void expand_unmasked(const uint32_t *rows, int S, int x, int y, int BX, int BY, int &rx, int &ry)
{
  bool right = true, down = true;
  rx = 1;
  ry = 1;

  while (right || down)
  {
    if (right)
    {
      bool ok = x + rx + 1 < BX && rx < S;

      if (ok)
        for (int r = 0; r < ry; r++)
          ok &= (rows[r] >> rx) & 1u;

      if (ok) rx++; else right = false;
    }

    if (down)
    {
      bool ok = y + ry + 1 < BY && ry < S;

      if (ok)
      {
        const uint32_t m = rx >= 32 ? 0xFFFFFFFFu : ((1u << rx) - 1u);
        ok = (rows[ry] & m) == m;
      }

      if (ok) ry++; else down = false;
    }
  }
}

__global__ void __launch_bounds__(256) k_plan_seeds(PlanArgs a)
{
  const int seed = blockIdx.x * blockDim.x + threadIdx.x;
  const bool inside = seed < a.BX * a.BY;
  const int y = inside ? seed / a.BX : 0, x = inside ? seed - y * a.BX : 0;
  const uint32_t w0 = inside ? a.window[(size_t)seed * 2] : 0u, w1 = inside ? a.window[(size_t)seed * 2 + 1] : 0u;
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
  expand_unmasked(rows, 8, x, y, a.BX, a.BY, rx, ry);
  a.extSlot[seed] = LIMG_NO_SLOT;
  a.symSlot[seed] = LIMG_NO_SLOT;
  a.symStart[seed] = 0;
  a.unmasked[seed] = (uint16_t)(rx | (ry << 8));

  // a stage-0 candidate whose run along its row or its column reaches the window edge can grow out of the window
  const bool rowRun = (w0 & 0xFFu) == 0xFFu && x + 8 < a.BX;
  const bool colRun = (w0 & 0x01010101u) == 0x01010101u && (w1 & 0x01010101u) == 0x01010101u && y + 8 < a.BY;

  if (c[0] && (rowRun || colRun))
  {
    const uint32_t slot = atomicAdd(&a.counters[0], 1u);

    if (slot < a.extCap)
    {
      a.extSeed[slot] = seed;                // k_plan_extend publishes extSlot[seed] once the bitmap exists (the scan may already be running)
      a.extSlot[seed] = LIMG_SLOT_PENDING;   // also tells k_plan_centres that the seed's mask-free rectangle is not final yet
    }
  }
}

// first zero bit of `known` at or above `from` (32 if none): the run length along one direction
__device__ __forceinline__ int run_end(uint32_t bits, int from)
{
  const uint32_t z = ~bits & (0xFFFFFFFFu << from);
  return z ? __ffs(z) - 1 : 32;
}

// One WARP per slot: the cells of the known part that are not already known are evaluated 32 at a time (one predicate per lane),
// so no lane idles while others work and no block-wide barrier is needed.
#define LIMG_PLAN_WARPS 8

template <int CH>
__global__ void __launch_bounds__(LIMG_PLAN_WARPS * 32) k_plan_extend(PlanArgs a, int rowLo, int rowHi /* seeds of block rows [rowLo, rowHi) only */)
{
  __shared__ uint32_t sRows[LIMG_PLAN_WARPS][32];
  const uint32_t count = min(a.counters[0], a.extCap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *rows = sRows[warp];
  const uint32_t warpsTotal = gridDim.x * LIMG_PLAN_WARPS;

  for (uint32_t slot = blockIdx.x * LIMG_PLAN_WARPS + warp; slot < count; slot += warpsTotal)
  {
    const int seed = (int)a.extSeed[slot];
    const int y = seed / a.BX, x = seed - y * a.BX;

    if (y < rowLo || y >= rowHi)
      continue;

    const PredRec s = a.rec[seed];
    const uint32_t w0 = a.window[(size_t)seed * 2], w1 = a.window[(size_t)seed * 2 + 1];
    const uint32_t win = lane < 8 ? ((lane < 4 ? w0 >> (8 * lane) : w1 >> (8 * (lane - 4))) & 0xFFu) : 0u; // lane = row of the 8x8 window

    // the two runs: along the seed's row (dx = lane) and along its column (dy = lane); the first 8 are known
    bool mr, mc;

    if (lane < 8)
    {
      mr = (w0 >> lane) & 1u;
      mc = win & 1u;
    }
    else
    {
      mr = x + lane < a.BX && predicate_thread<CH>(s, a.rec[(size_t)y * a.BX + x + lane]);
      mc = y + lane < a.BY && predicate_thread<CH>(s, a.rec[(size_t)(y + lane) * a.BX + x]);
    }

    const uint32_t rowRun = __ballot_sync(0xFFFFFFFFu, mr), colRun = __ballot_sync(0xFFFFFFFFu, mc);
    // known part: columns [0, first mismatch along the row], rows [0, first mismatch along the column]
    const int vx1 = min(run_end(rowRun, 0) + 1, a.extMaxW), vy1 = min(run_end(colRun, 0) + 1, 32);

    // row words start with what is known: the window, the row run (row 0), the column run (column 0)
    rows[lane] = lane >= vy1 ? 0u : ((lane == 0 ? rowRun : win) | ((colRun >> lane) & 1u)) & (vx1 >= 32 ? 0xFFFFFFFFu : ((1u << vx1) - 1u));
    __syncwarp();

    for (int base = 0; base < vx1 * vy1; base += 32)
    {
      const int cell = base + lane;
      const int dy = cell / vx1, dx = cell - dy * vx1;

      if (cell < vx1 * vy1 && dx > 0 && dy > 0 && (dx >= 8 || dy >= 8) && x + dx < a.BX && y + dy < a.BY)
      {
        if (predicate_thread<CH>(s, a.rec[(size_t)(y + dy) * a.BX + x + dx]))
          atomicOr(&rows[dy], 1u << dx);
      }
    }

    __syncwarp();
    a.extBits[(size_t)slot * 32 + lane] = rows[lane];

    if (lane == 0)
    {
      int rx, ry;
      a.extHdr[slot] = (uint32_t)vx1 << 16 | (uint32_t)vy1 << 24;
      expand_unmasked(rows, 32, x, y, a.BX, a.BY, rx, ry);
      *(volatile uint16_t *)&a.unmasked[seed] = (uint16_t)(min(rx, 255) | (min(ry, 255) << 8));
    }

    // publish: the bitmap is complete before anybody can find it
    __threadfence();
    __syncwarp();

    if (lane == 0)
      *(volatile uint32_t *)&a.extSlot[seed] = slot;

    __syncwarp();
  }
}

// which blocks will the four-way regrowth probably start from? (limg.cpp:1426-1433 with the mask-free rectangle, and up to three
// blocks to the left of that: a mask shrinks the rectangle's width far more often than its height)
// phase 0: the candidates whose growth stays inside their 8x8 word (their mask-free rectangle is final after k_plan_seeds): their centre
// bitmaps can be built at once, before the (slower) extension bitmaps; phase 1: the others, after k_plan_extend.
// Candidates of block rows [rowLo, rowHi) only: the bitmaps are built top-down in bands of block rows, the order in which the scan wants them.
__global__ void __launch_bounds__(256) k_plan_centres(PlanArgs a, int phase, int rowLo, int rowHi)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

  if (i >= a.candCount[0])
    return;

  const int seed = (int)a.candList[i];

  if ((*(volatile uint32_t *)&a.extSlot[seed] != LIMG_NO_SLOT) != (phase != 0))
    return;

  const int y = seed / a.BX, x = seed - y * a.BX;

  if (y < rowLo || y >= rowHi)
    return;

  const uint32_t u = a.unmasked[seed];
  const int rx = u & 0xFF, ry = u >> 8;

  if (rx < 3 || ry < 3)
    return;

  const int cy = y + ry / 3;

  for (int d = 0; d < 4 && rx / 3 - d >= 1; d++)
  {
    const int c = cy * a.BX + x + rx / 3 - d;
    // the regrowth starts from an untested rectangle of (rx / 3) x (ry / 3) blocks: remember the largest one asked for
    atomicMax(&a.symStart[c], (uint32_t)min(rx / 3, 15) | ((uint32_t)min(ry / 3, 15) << 8));

    if (atomicCAS(&a.symSlot[c], LIMG_NO_SLOT, LIMG_SLOT_PENDING) == LIMG_NO_SLOT) // first one to ask for this centre
    {
      const uint32_t slot = atomicAdd(&a.counters[1], 1u);

      if (slot < a.symCap)
        a.symSeed[slot] = (uint32_t)c; // symSlot[c] stays "pending" until k_plan_sym has written the bitmap
    }
  }
}

// behind every k_plan_sym launch: the next one starts where this one stopped
__global__ void k_plan_mark(PlanArgs a)
{
  a.counters[2] = a.counters[3];
}

// After the bitmaps exist: every stage-0 candidate remembers the slots of the four centres k_plan_centres asked for on its behalf,
// so the scan finds them with the seed's first load instead of a dependent lookup per centre.
__global__ void __launch_bounds__(256) k_plan_link(PlanArgs a, uint4 *__restrict__ seedSym)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

  if (i >= a.candCount[0])
    return;

  const int seed = (int)a.candList[i];
  const int y = seed / a.BX, x = seed - y * a.BX;
  const uint32_t u = a.unmasked[seed];
  const int rx = u & 0xFF, ry = u >> 8;
  uint32_t s[4] = { LIMG_NO_SLOT, LIMG_NO_SLOT, LIMG_NO_SLOT, LIMG_NO_SLOT };

  if (rx >= 3 && ry >= 3)
  {
    const int cy = y + ry / 3;

    for (int d = 0; d < 4 && rx / 3 - d >= 1; d++)
      s[d] = a.symSlot[cy * a.BX + x + rx / 3 - d];
  }

  seedSym[seed] = make_uint4(s[0], s[1], s[2], s[3]);
}

// Match bitmap around a centre c, anchored at (cx - 8, cy - 8): every four-way rectangle grown from a rectangle that contains c's
// block row and column segment lies inside the bounding box of the four runs from c (each strip it adds crosses c's row or column).
// Warp-cooperative; `rows` are 32 words of shared memory private to the warp. The regrowth starts from an untested rectangle of
// crx x cry blocks (limg.cpp:1428-1431): a mismatch less than crx blocks right of c / cry blocks below it does not stop the growth.
// Returns this lane's row word and the header (known part) in hdr.
template <int CH>
__device__ uint32_t build_centre_bitmap(const PredRec *__restrict__ rec, const uint32_t *__restrict__ window, int BX, int BY, int c, int crx, int cry, int maxL, int maxR, int maxD,
                                        uint32_t *rows, uint32_t &hdr)
{
  const int lane = threadIdx.x & 31;
  const int cy = c / BX, cx = c - cy * BX;
  const int ax = cx - LIMG_SYM_BACK, ay = cy - LIMG_SYM_BACK;
  const PredRec s = rec[c];
  const uint32_t w0 = window[(size_t)c * 2], w1 = window[(size_t)c * 2 + 1];

  auto known = [&](int dx, int dy) -> bool { // lower-right 8x8 of c: its match word
    return ((dy < 4 ? w0 >> (8 * dy) : w1 >> (8 * (dy - 4))) >> dx) & 1u;
  };

  // runs along c's row (column ax + lane) and c's column (row ay + lane), limited to the part the caps allow
  const int d = lane - LIMG_SYM_BACK;
  bool mr = false, mc = false;

  if (d >= 0 && d < 8)
  {
    mr = known(d, 0);
    mc = known(0, d);
  }
  else
  {
    if (d >= -maxL && d < maxR && cx + d >= 0 && cx + d < BX)
      mr = predicate_thread<CH>(s, rec[(size_t)cy * BX + cx + d]);

    if (d >= -maxL && d < maxD && cy + d >= 0 && cy + d < BY)
      mc = predicate_thread<CH>(s, rec[(size_t)(cy + d) * BX + cx]);
  }

  const uint32_t rowRun = __ballot_sync(0xFFFFFFFFu, mr), colRun = __ballot_sync(0xFFFFFFFFu, mc);
  // known box: from the first mismatch left of / above c to the first mismatch right of / below c (inclusive), relative to the anchor
  const uint32_t lowRow = ~rowRun & ((1u << LIMG_SYM_BACK) - 1u), lowCol = ~colRun & ((1u << LIMG_SYM_BACK) - 1u);
  const int vx0 = max(lowRow ? 31 - __clz(lowRow) : 0, LIMG_SYM_BACK - maxL), vy0 = max(lowCol ? 31 - __clz(lowCol) : 0, LIMG_SYM_BACK - maxL);
  crx = max(min(crx, 15), 3);
  cry = max(min(cry, 15), 3);
  const int vx1 = min(run_end(rowRun | (((1u << (crx - 1)) - 1u) << (LIMG_SYM_BACK + 1)), LIMG_SYM_BACK) + 1, LIMG_SYM_BACK + maxR);
  const int vy1 = min(run_end(colRun | (((1u << (cry - 1)) - 1u) << (LIMG_SYM_BACK + 1)), LIMG_SYM_BACK) + 1, LIMG_SYM_BACK + maxD);
  const int bw = vx1 - vx0, bh = vy1 - vy0;

  // what is known without further predicates: c's match word (lower-right quadrant), c's row, c's column
  {
    uint32_t r = 0;
    const int dy = lane - LIMG_SYM_BACK;

    if (lane >= vy0 && lane < vy1)
    {
      if (dy >= 0 && dy < 8)
        r = ((dy < 4 ? w0 >> (8 * dy) : w1 >> (8 * (dy - 4))) & 0xFFu) << LIMG_SYM_BACK;

      if (dy == 0)
        r |= rowRun;

      r |= ((colRun >> lane) & 1u) << LIMG_SYM_BACK;
      r &= (vx1 >= 32 ? 0xFFFFFFFFu : ((1u << vx1) - 1u)) & (0xFFFFFFFFu << vx0);
    }

    rows[lane] = r;
  }

  __syncwarp();

  for (int base = 0; base < bw * bh; base += 32)
  {
    const int cell = base + lane;
    const int r = vy0 + cell / bw, col = vx0 + cell % bw;
    const int dx = col - LIMG_SYM_BACK, dy = r - LIMG_SYM_BACK;
    const int bx = ax + col, by = ay + r;

    if (cell < bw * bh && dx != 0 && dy != 0 && !(dx > 0 && dx < 8 && dy > 0 && dy < 8) && bx >= 0 && bx < BX && by >= 0 && by < BY)
    {
      if (predicate_thread<CH>(s, rec[(size_t)by * BX + bx]))
        atomicOr(&rows[r], 1u << col);
    }
  }

  __syncwarp();
  hdr = (uint32_t)vx0 | (uint32_t)vy0 << 8 | (uint32_t)vx1 << 16 | (uint32_t)vy1 << 24;
  const uint32_t mine = rows[lane];
  __syncwarp();
  return mine;
}

// The slots requested since the last launch: [counters[2], count). It leaves `count` in counters[3]; k_plan_mark, launched behind it, copies that
// to counters[2] (not this kernel: other CTAs of the launch still read it).
template <int CH>
__global__ void __launch_bounds__(LIMG_PLAN_WARPS * 32) k_plan_sym(PlanArgs a)
{
  __shared__ uint32_t sRows[LIMG_PLAN_WARPS][32];
  const uint32_t count = min(a.counters[1], a.symCap);
  const uint32_t first = min(a.counters[2], count);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t warpsTotal = gridDim.x * LIMG_PLAN_WARPS;

  if (blockIdx.x == 0 && threadIdx.x == 0)
    a.counters[3] = count;

  for (uint32_t slot = first + blockIdx.x * LIMG_PLAN_WARPS + warp; slot < count; slot += warpsTotal)
  {
    const int c = (int)a.symSeed[slot];
    const uint32_t start = a.symStart[c];
    uint32_t hdr;
    const uint32_t row = build_centre_bitmap<CH>(a.rec, a.window, a.BX, a.BY, c, (int)(start & 0xFF), (int)(start >> 8), a.symMaxL, a.symMaxR, a.symMaxD, sRows[warp], hdr);
    a.symBits[(size_t)slot * 32 + lane] = row;

    if (lane == 0)
      a.symHdr[slot] = hdr;

    __threadfence();
    __syncwarp();

    if (lane == 0)
      *(volatile uint32_t *)&a.symSlot[c] = slot;

    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// How far LEFT of its seed can a stage-0 seed's centre-third regrowth reach? (limg.cpp:1426-1433, 1363-1376)
//
// The row-pipelined scan lets a block row decide a seed once the rows above have passed everything the seed probed. A rectangle
// grown right/down from a seed of a row above starts at that seed's column, but its four-way regrowth also grows LEFT, so an
// undecided seed further right can still claim blocks the row below has looked at. The regrowth's rectangle always contains its
// centre's block row, so it cannot pass the first block left of the centre that does not match the centre (mask-free bound; in-use
// blocks only stop it earlier), and the centre lies inside the seed's right/down rectangle, i.e. at most (run along the seed's
// row) / 3 columns and (run along its column) / 3 rows from the seed. k_pred_leftrun measures the run of matches left of every block;
// k_plan_safe turns it, per block row, into the SAFE column of every position: no undecided candidate at or right of that position
// can touch anything left of the safe column. A row publishes min(own progress, safe column) and the rows below need no margin on
// top of it. Where the bound is unknown (a run that leaves the 8x8 window, a left run longer than the cap) the old assumption
// stands in: such a regrowth reaches at most LIMG_SAFE_DEFAULT_REACH columns left of its seed; the verification pass covers the rest.
// ---------------------------------------------------------------------------------------------

#define LIMG_LEFTRUN_CAP 24
#define LIMG_LEFTRUN_UNKNOWN 255
#define LIMG_SAFE_DEFAULT_REACH 10

template <int CH>
__global__ void __launch_bounds__(128) k_pred_leftrun(const PredRec *__restrict__ rec, int BX, int BY, uint8_t *__restrict__ leftRun)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;

  if (b >= BX * BY)
    return;

  const int y = b / BX, x = b - y * BX;
  const PredRec s = rec[b];
  int run = 0;

  while (run < LIMG_LEFTRUN_CAP && x - run - 1 >= 0 && predicate_thread<CH>(s, rec[b - run - 1]))
    run++;

  // (a run that ends at the grid edge is exact)
  leftRun[b] = (uint8_t)((run == LIMG_LEFTRUN_CAP && x - run - 1 >= 0) ? LIMG_LEFTRUN_UNKNOWN : run);
}

// one warp per block row; safe[y * BX + x] = safe column of position x (low 16 bits: candidates at or right of x, i.e. while the candidate at x
// is undecided; high 16 bits: candidates right of x only, i.e. once x is decided)
__global__ void __launch_bounds__(32) k_plan_safe(const uint32_t *__restrict__ window, const uint32_t *__restrict__ candBits0, const uint8_t *__restrict__ leftRun, int BX, int BY,
                                                  int wordsPerRow, uint32_t *__restrict__ safe)
{
  const int y = blockIdx.x, lane = threadIdx.x;
  const int chunks = (BX + 31) >> 5;
  int carry = 0xFFFF; // safe column of everything right of the current chunk

  for (int c = chunks - 1; c >= 0; c--)
  {
    const int x = c * 32 + lane;
    int mine = 0xFFFF; // leftmost column the regrowth of the candidate at x can touch

    if (x < BX && ((candBits0[(size_t)y * wordsPerRow + c] >> lane) & 1u))
    {
      const size_t seed = (size_t)y * BX + x;
      const uint32_t w0 = window[seed * 2], w1 = window[seed * 2 + 1];
      // runs along the seed's row and column inside its 8x8 word (the seed itself counts)
      const int rowRun = __ffs((int)(~(w0 & 0xFFu) & 0x1FFu)) - 1;
      const uint32_t col = (w0 & 1u) | ((w0 >> 7) & 2u) | ((w0 >> 14) & 4u) | ((w0 >> 21) & 8u) | ((w1 & 1u) << 4) | ((w1 >> 3) & 32u) | ((w1 >> 10) & 64u) | ((w1 >> 17) & 128u);
      const int colRun = __ffs((int)(~col & 0x1FFu)) - 1;

      if ((rowRun == 8 && x + 8 < BX) || (colRun == 8 && y + 8 < BY))
      {
        mine = max(x - LIMG_SAFE_DEFAULT_REACH, 0); // the runs leave the word: where the centre can be is not known here
      }
      else
      {
        for (int dy = 1; dy <= colRun / 3; dy++)
          for (int dx = 1; dx <= rowRun / 3; dx++)
          {
            const int cx = x + dx, cy = y + dy;

            if (cx >= BX || cy >= BY)
              continue;

            const int run = leftRun[(size_t)cy * BX + cx];
            mine = min(mine, run == LIMG_LEFTRUN_UNKNOWN ? max(x - LIMG_SAFE_DEFAULT_REACH, 0) : cx - run);
          }

        mine = min(mine, x);
      }
    }

    // suffix minimum over the lanes (inclusive), then the columns right of this chunk
    int incl = mine;

#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const int t = __shfl_down_sync(0xFFFFFFFFu, incl, o);
      if (lane + o < 32) incl = min(incl, t);
    }

    int after = __shfl_down_sync(0xFFFFFFFFu, incl, 1);
    after = min(lane < 31 ? after : 0xFFFF, carry);
    const int from = min(mine, after);

    if (x < BX)
      safe[(size_t)y * BX + x] = (uint32_t)from | ((uint32_t)after << 16);

    carry = min(carry, __shfl_sync(0xFFFFFFFFu, incl, 0));
  }
}

// ---------------------------------------------------------------------------------------------
// after the scan: leftovers, geometry, block map, size classes (all grid-wide; the area order is stage 0, stage 1, leftovers in
// raster order, limg.cpp:1814-1878)
// ---------------------------------------------------------------------------------------------

struct PrepareArgs
{
  int W, H, BX, BY, wordsPerRow, listCap;
  limgcu_area *areas;
  const uint32_t *used;      // [BY][wordsPerRow] in-use mask after both merge stages (all zero: every block is its own area)
  const uint32_t *rowCounts; // [2][BY] rectangles emitted per block row and stage
  const uint2 *rowLists;     // [2][BY][listCap]
  const uint32_t *tau;       // [blocks] owner time of merged blocks (nullptr: nothing merged)
  const uint32_t *emitInfo;  // [2][blocks]
  uint32_t *rowLeft;         // [BY] leftover blocks per row
  uint32_t *rowBase;         // [3][BY] index of the first area of the row's stage-0 / stage-1 rectangles / leftovers
  uint32_t *mergedCount, *areaCount;
  uint32_t *blockToArea;
  AreaWork *work;
  uint32_t *smallList, *largeList;
  uint32_t *smallCount, *largeCount, *bigCount, *hugeCount, *scratchTop; // large: 256 < n <= LIMG_CTA_AREA_CAP, big: up to LIMG_HUGE_AREA_PX, huge: above
  uint32_t largeCap;         // entries of largeList
};

// one warp per block row: number of blocks no rectangle covers
__global__ void __launch_bounds__(32) k_prepare_rowleft(PrepareArgs a)
{
  const int y = blockIdx.x, lane = threadIdx.x;
  const int nWords = (a.BX + 31) >> 5;
  uint32_t n = 0;

  for (int w = lane; w < nWords; w += 32)
  {
    const int cols = min(32, a.BX - w * 32);
    const uint32_t valid = cols == 32 ? 0xFFFFFFFFu : ((1u << cols) - 1u);
    n += __popc(~a.used[(size_t)y * a.wordsPerRow + w] & valid);
  }

  n = __reduce_add_sync(0xFFFFFFFFu, n);

  if (lane == 0)
    a.rowLeft[y] = n;
}

// one CTA per block row: where the row's areas go in the emission order, then the rectangles and leftovers themselves
__global__ void __launch_bounds__(128) k_prepare_collect(PrepareArgs a)
{
  __shared__ uint32_t sRed[6][4];
  const int y = blockIdx.x;
  uint32_t v[6] = { 0, 0, 0, 0, 0, 0 }; // before0, before1, beforeLeft, total0, total1, totalLeft

  for (int r = threadIdx.x; r < a.BY; r += blockDim.x)
  {
    const uint32_t c0 = a.rowCounts[r], c1 = a.rowCounts[a.BY + r], cl = a.rowLeft[r];
    v[3] += c0; v[4] += c1; v[5] += cl;

    if (r < y)
    {
      v[0] += c0; v[1] += c1; v[2] += cl;
    }
  }

#pragma unroll
  for (int i = 0; i < 6; i++)
  {
    v[i] = __reduce_add_sync(0xFFFFFFFFu, v[i]);

    if ((threadIdx.x & 31) == 0)
      sRed[i][threadIdx.x >> 5] = v[i];
  }

  __syncthreads();

#pragma unroll
  for (int i = 0; i < 6; i++)
    v[i] = sRed[i][0] + sRed[i][1] + sRed[i][2] + sRed[i][3];

  const uint32_t base[3] = { v[0], v[3] + v[1], v[3] + v[4] + v[2] };

  if (threadIdx.x < 3)
    a.rowBase[threadIdx.x * a.BY + y] = base[threadIdx.x];

  for (int stage = 0; stage < 2; stage++)
  {
    const uint32_t n = a.rowCounts[(size_t)stage * a.BY + y];
    const uint2 *l = a.rowLists + ((size_t)stage * a.BY + y) * a.listCap;

    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
      const uint2 r = l[i];
      limgcu_area *out = &a.areas[base[stage] + i];
      out->ox = r.x & 0xFFFF; out->oy = r.x >> 16; out->rx = r.y & 0xFFFF; out->ry = r.y >> 16;
      out->stage = stage;
    }
  }

  // leftovers of this row, left to right (warp 0)
  if (threadIdx.x < 32)
  {
    const int nWords = (a.BX + 31) >> 5;
    uint32_t at = base[2];

    for (int w0 = 0; w0 < nWords; w0 += 32)
    {
      const int w = w0 + threadIdx.x;
      uint32_t free = 0;

      if (w < nWords)
      {
        const int cols = min(32, a.BX - w * 32);
        free = ~a.used[(size_t)y * a.wordsPerRow + w] & (cols == 32 ? 0xFFFFFFFFu : ((1u << cols) - 1u));
      }

      // exclusive prefix of the word counts over the lanes
      const uint32_t cnt = __popc(free);
      uint32_t incl = cnt;

#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((int)threadIdx.x >= o) incl += t;
      }

      uint32_t k = at + incl - cnt;

      while (free)
      {
        const int b = __ffs(free) - 1;
        free &= free - 1;
        limgcu_area *out = &a.areas[k++];
        out->ox = w * 32 + b; out->oy = y; out->rx = 1; out->ry = 1;
        out->stage = 2;
      }

      at += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
  }

  if (y == 0 && threadIdx.x == 0)
  {
    *a.mergedCount = v[3] + v[4];
    *a.areaCount = v[3] + v[4] + v[5];
  }
}

// one thread per area: pixel rectangle (edge fit, limg.cpp:1725-1739), size class, scratch
__global__ void __launch_bounds__(256) k_prepare_geometry(PrepareArgs a)
{
  const uint32_t count = *a.areaCount;

  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x)
  {
    limgcu_area *ar = &a.areas[k];
    const uint32_t ox = ar->ox, oy = ar->oy, rx = ar->rx, ry = ar->ry;
    uint32_t pw = rx * LIMG_BLOCK, ph = ry * LIMG_BLOCK;

    if (ox + rx == (uint32_t)a.BX && (a.W % LIMG_BLOCK)) pw = pw - LIMG_BLOCK + a.W % LIMG_BLOCK;
    if (oy + ry == (uint32_t)a.BY && (a.H % LIMG_BLOCK)) ph = ph - LIMG_BLOCK + a.H % LIMG_BLOCK;

    ar->px_x = ox * LIMG_BLOCK; ar->px_y = oy * LIMG_BLOCK; ar->px_w = pw; ar->px_h = ph;
    const uint32_t n = pw * ph;
    a.work[k].n = n;
    // area-contiguous scratch: any disjoint ranges do (the areas tile the image, so they fit into one image-sized buffer)
    a.work[k].scratchOff = n > LIMG_SMALL_AREA_PX ? atomicAdd(a.scratchTop, n) : 0u;

    // Huge areas are the long poles of the per-area encode (their reference-order sums are one dependent chain per area, their trials run
    // on one SM): they go to the front of the list and get a kernel of their own with 512-thread CTAs that starts first; the others fill
    // the list from the back.
    if (n <= LIMG_SMALL_AREA_PX)
      a.smallList[atomicAdd(a.smallCount, 1u)] = k;
    else if (n > LIMG_HUGE_AREA_PX)
      a.largeList[atomicAdd(a.hugeCount, 1u)] = k;
    else if (n > LIMG_CTA_AREA_CAP)
      a.smallList[a.largeCap - 1u - atomicAdd(a.bigCount, 1u)] = k; // the free end of the small list (small + big <= areas <= blocks): first in the 256-thread kernel
    else
      a.largeList[a.largeCap - 1u - atomicAdd(a.largeCount, 1u)] = k;
  }
}

// one thread per block: which area owns it (from the owner times the scan wrote; leftovers by position)
__global__ void __launch_bounds__(256) k_prepare_blockmap(PrepareArgs a)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;

  if (b >= a.BX * a.BY)
    return;

  const int y = b / a.BX, x = b - y * a.BX;
  const uint32_t t = a.tau ? a.tau[b] : 0xFFFFFFFFu;
  uint32_t k;

  if (t == 0xFFFFFFFFu)
  {
    // leftover: blocks of the row that are free and left of this one
    const uint32_t *row = a.used + (size_t)y * a.wordsPerRow;
    uint32_t before = 0;

    for (int w = 0; w < (x >> 5); w++)
      before += __popc(~row[w]);

    before += __popc(~row[x >> 5] & ((1u << (x & 31)) - 1u));
    k = a.rowBase[2 * a.BY + y] + before;
  }
  else
  {
    const int stage = t >= 0x40000000u ? 1 : 0;
    const uint32_t seed = (t & 0x3FFFFFFFu) >> 3, attempt = t & 7u;
    const uint32_t sy = seed / a.BX;
    k = a.rowBase[stage * a.BY + sy] + (a.emitInfo[(size_t)stage * a.BX * a.BY + seed] >> 8) + attempt;
  }

  a.blockToArea[b] = k;
}

} // namespace limg

// limg_b200/csrc/group.cuh -- cooperative "group" primitives: a group is either one warp (WARPS == 1, several
// independent groups per CTA) or the whole CTA (WARPS == blockDim.x / 32). All group functions must be called by
// every thread of the group with group-uniform control flow.
#pragma once

#include "common.cuh"

namespace limg
{

template <int WARPS>
struct GroupScratch
{
  // double-buffered so that one barrier per reduction is enough
  int32_t i[2][WARPS];
  int32_t flag[2][WARPS];
  float fmin[2][WARPS];
  float fmax[2][WARPS];
  float bcast[8];
  uint32_t parity;
};

template <int WARPS>
__device__ __forceinline__ int group_tid()
{
  return WARPS == 1 ? (threadIdx.x & 31) : threadIdx.x;
}

template <int WARPS>
__device__ __forceinline__ void group_sync()
{
  if (WARPS == 1)
    __syncwarp();
  else
    __syncthreads();
}

// wrapping 32-bit sum + "any" flag in one barrier. Every thread returns the same values.
template <int WARPS>
__device__ __forceinline__ void group_sum_any(int32_t &v, bool &flag, GroupScratch<WARPS> *gs, uint32_t &parity)
{
  v = (int32_t)__reduce_add_sync(0xFFFFFFFFu, (uint32_t)v);
  flag = __any_sync(0xFFFFFFFFu, flag);

  if (WARPS > 1)
  {
    const int w = threadIdx.x >> 5;
    const uint32_t p = parity & 1;
    parity++;

    if ((threadIdx.x & 31) == 0)
    {
      gs->i[p][w] = v;
      gs->flag[p][w] = flag;
    }

    __syncthreads();

    uint32_t s = 0;
    int f = 0;

#pragma unroll
    for (int k = 0; k < WARPS; k++)
    {
      s += (uint32_t)gs->i[p][k];
      f |= gs->flag[p][k];
    }

    v = (int32_t)s;
    flag = f != 0;
  }
}

// MINPS / MAXPS folds. The fold is order independent for the values that can occur ("fit").
template <int WARPS>
__device__ __forceinline__ void group_minmax(float &mn, float &mx, GroupScratch<WARPS> *gs, uint32_t &parity)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    mn = sse_min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
    mx = sse_max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }

  if (WARPS > 1)
  {
    const int w = threadIdx.x >> 5;
    const uint32_t p = parity & 1;
    parity++;

    if ((threadIdx.x & 31) == 0)
    {
      gs->fmin[p][w] = mn;
      gs->fmax[p][w] = mx;
    }

    __syncthreads();

    mn = gs->fmin[p][0];
    mx = gs->fmax[p][0];

#pragma unroll
    for (int k = 1; k < WARPS; k++)
    {
      mn = sse_min(mn, gs->fmin[p][k]);
      mx = sse_max(mx, gs->fmax[p][k]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// three-factor fit of a pixel list held in `px` (shared or global, area-contiguous order).
//   limg.cpp:469-497 (sums) + limg_factorization.h:385-576 (RGB) / 581-794 (RGBA).
// The three "mean direction" sums are accumulated in the reference's pixel order by four lanes (one per channel)
// from a staging buffer, so the result is bit-identical to the sequential SSE loop; everything else is data parallel.
// `stage` holds stagePx float4 entries. Every thread of the group returns the same record in `out`.
// ---------------------------------------------------------------------------------------------

template <int CH>
__device__ __forceinline__ f4 unit_direction(const f4 &v, const uint16_t *__restrict__ lut)
{
  // limg_factorization.h:412-431: sign-normalise so that the largest |component| is positive (ties by lane bias), then scale
  // by RSQRTPS(|v|^2). An all-zero vector contributes nothing.
  if (v.x == 0.0f && v.y == 0.0f && v.z == 0.0f && v.w == 0.0f)
    return { 0.0f, 0.0f, 0.0f, 0.0f };

  const float e = FLT_EPSILON;
  const float lo0 = fsub(v.x, e * 3), lo1 = fsub(v.y, e * 2), lo2 = fsub(v.z, e * 1), lo3 = fsub(v.w, 0.0f);
  const float hi0 = fadd(v.x, e * 3), hi1 = fadd(v.y, e * 2), hi2 = fadd(v.z, e * 1), hi3 = fadd(v.w, 0.0f);
  const float absMin = fabsf(sse_min(sse_min(lo0, lo2), sse_min(lo1, lo3)));
  const float mx = sse_max(sse_max(hi0, hi2), sse_max(hi1, hi3));
  float inv = sse_rsqrt(dpn<CH>(v, v), lut);

  if (absMin > mx)
    inv = __uint_as_float(__float_as_uint(inv) ^ 0x80000000u);

  return { fmul(v.x, inv), fmul(v.y, inv), fmul(v.z, inv), fmul(v.w, inv) };
}

template <int WARPS>
__device__ __forceinline__ void ordered_accumulate(const float4 *stage, int count, float &acc)
{
  // lanes 0..3 of the group's first warp: acc is channel `lane` of the running sum. The sum is one dependent chain of FADDs (4 cycles each)
  // in the reference's pixel order; the loads of the next eight terms are issued before the current eight are added, so the chain never
  // waits for shared memory.
  const int t = group_tid<WARPS>();

  if (t < 4)
  {
    const float *s = reinterpret_cast<const float *>(stage) + t;
    int i = 0;

    if (count >= 8)
    {
      float v[8];

#pragma unroll
      for (int k = 0; k < 8; k++)
        v[k] = s[k * 4];

      for (i = 8; i + 8 <= count; i += 8)
      {
        float nx[8];

#pragma unroll
        for (int k = 0; k < 8; k++)
          nx[k] = s[(i + k) * 4];

#pragma unroll
        for (int k = 0; k < 8; k++)
          acc = fadd(acc, v[k]);

#pragma unroll
        for (int k = 0; k < 8; k++)
          v[k] = nx[k];
      }

#pragma unroll
      for (int k = 0; k < 8; k++)
        acc = fadd(acc, v[k]);
    }

    for (; i < count; i++)
      acc = fadd(acc, s[i * 4]);
  }
}

template <int WARPS>
__device__ __forceinline__ f4 broadcast_acc(float acc, float scale, GroupScratch<WARPS> *gs)
{
  const int t = group_tid<WARPS>();

  if (t < 4)
    gs->bcast[t] = fmul(acc, scale);

  group_sync<WARPS>();
  const f4 r = { gs->bcast[0], gs->bcast[1], gs->bcast[2], gs->bcast[3] };
  group_sync<WARPS>();
  return r;
}

template <int CH, int WARPS>
__device__ void group_fit(const uint32_t *px, uint32_t n, const uint16_t *__restrict__ lut, float4 *stage, int stagePx, GroupScratch<WARPS> *gs, uint32_t &parity, limgcu_decomp &out)
{
  constexpr int THREADS = WARPS * 32;
  const int t = group_tid<WARPS>();

  // K1: wrapping 32-bit channel sums, read back as signed (limg_factorization.h:396-402)
  uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;

  for (uint32_t i = t; i < n; i += THREADS)
  {
    const uint32_t p = px[i];
    s0 += p & 0xFF;
    s1 += (p >> 8) & 0xFF;
    s2 += (p >> 16) & 0xFF;
    s3 += p >> 24;
  }

  {
    bool dummy = false;
    int32_t v;
    v = (int32_t)s0; group_sum_any<WARPS>(v, dummy, gs, parity); s0 = (uint32_t)v;
    v = (int32_t)s1; group_sum_any<WARPS>(v, dummy, gs, parity); s1 = (uint32_t)v;
    v = (int32_t)s2; group_sum_any<WARPS>(v, dummy, gs, parity); s2 = (uint32_t)v;
    v = (int32_t)s3; group_sum_any<WARPS>(v, dummy, gs, parity); s3 = (uint32_t)v;
  }

  const float invCount = frcp1(__ull2float_rn((unsigned long long)n));
  f4 avg = { fmul((float)(int32_t)s0, invCount), fmul((float)(int32_t)s1, invCount), fmul((float)(int32_t)s2, invCount), fmul((float)(int32_t)s3, invCount) };

  if (CH == 3)
    avg.w = 0.0f;

  // ---- direction A: mean sign-normalised unit vector from the mean colour -------------------------------
  float acc = 0.0f;

  for (uint32_t base = 0; base < n; base += stagePx)
  {
    const uint32_t cnt = min((uint32_t)stagePx, n - base);

    for (uint32_t k = t; k < cnt; k += THREADS)
    {
      const f4 c = px_to_f4(px[base + k]);
      f4 v = { fsub(c.x, avg.x), fsub(c.y, avg.y), fsub(c.z, avg.z), fsub(c.w, avg.w) };
      if (CH == 3) v.w = 0.0f;
      const f4 u = unit_direction<CH>(v, lut);
      stage[k] = make_float4(u.x, u.y, u.z, u.w);
    }

    group_sync<WARPS>();
    ordered_accumulate<WARPS>(stage, (int)cnt, acc);
    group_sync<WARPS>();
  }

  const f4 dirA = broadcast_acc<WARPS>(acc, invCount, gs);

  f4 dirB = { 0, 0, 0, 0 }, dirC = { 0, 0, 0, 0 };
  float minA = 0.0f, maxA = 0.0f, minB = 0.0f, maxB = 0.0f, minC = 0.0f, maxC = 0.0f; // Q11

  if (!(dirA.x == 0.0f && dirA.y == 0.0f && dirA.z == 0.0f && dirA.w == 0.0f))
  {
    const float invLenA = frcp1(dpn<CH>(dirA, dirA));

    // ---- A extents + direction B --------------------------------------------------------------------------
    acc = 0.0f;

    for (uint32_t base = 0; base < n; base += stagePx)
    {
      const uint32_t cnt = min((uint32_t)stagePx, n - base);

      for (uint32_t k = t; k < cnt; k += THREADS)
      {
        const f4 c = px_to_f4(px[base + k]);
        const f4 toPx = { fsub(c.x, avg.x), fsub(c.y, avg.y), fsub(c.z, avg.z), fsub(c.w, avg.w) };
        const float facA = fmul(dpn<CH>(toPx, dirA), invLenA);
        minA = sse_min(minA, facA);
        maxA = sse_max(maxA, facA);
        const f4 est = { fadd(avg.x, fmul(facA, dirA.x)), fadd(avg.y, fmul(facA, dirA.y)), fadd(avg.z, fmul(facA, dirA.z)), fadd(avg.w, fmul(facA, dirA.w)) };
        f4 err = { fsub(c.x, est.x), fsub(c.y, est.y), fsub(c.z, est.z), fsub(c.w, est.w) };
        if (CH == 3) err.w = 0.0f;
        const f4 u = unit_direction<CH>(err, lut);
        stage[k] = make_float4(u.x, u.y, u.z, u.w);
      }

      group_sync<WARPS>();
      ordered_accumulate<WARPS>(stage, (int)cnt, acc);
      group_sync<WARPS>();
    }

    dirB = broadcast_acc<WARPS>(acc, invCount, gs);
    group_minmax<WARPS>(minA, maxA, gs, parity);

    const float invLenB = frcp1(dpn<CH>(dirB, dirB));
    minB = minC = FLT_MAX;
    maxB = maxC = -FLT_MAX;

    if (CH == 3)
    {
      // Q5: C = A x B with individually rounded products (limg_factorization.h:498-507)
      dirC.x = fsub(fmul(dirA.y, dirB.z), fmul(dirA.z, dirB.y));
      dirC.y = fsub(fmul(dirA.z, dirB.x), fmul(dirA.x, dirB.z));
      dirC.z = fsub(fmul(dirA.x, dirB.y), fmul(dirA.y, dirB.x));
      dirC.w = 0.0f;
      const float invLenC = frcp1(dpn<CH>(dirC, dirC));

      for (uint32_t i = t; i < n; i += THREADS)
      {
        const f4 c = px_to_f4(px[i]);
        const f4 toAvg = { fsub(c.x, avg.x), fsub(c.y, avg.y), fsub(c.z, avg.z), fsub(c.w, avg.w) };
        const float facA = fmul(dpn<CH>(toAvg, dirA), invLenA);
        const f4 estA = { fadd(avg.x, fmul(facA, dirA.x)), fadd(avg.y, fmul(facA, dirA.y)), fadd(avg.z, fmul(facA, dirA.z)), 0.0f };
        const f4 toPx = { fsub(c.x, estA.x), fsub(c.y, estA.y), fsub(c.z, estA.z), 0.0f };
        const float facB = fmul(dpn<CH>(toPx, dirB), invLenB);
        minB = sse_min(minB, facB);
        maxB = sse_max(maxB, facB);
        const f4 estB = { fadd(estA.x, fmul(facB, dirB.x)), fadd(estA.y, fmul(facB, dirB.y)), fadd(estA.z, fmul(facB, dirB.z)), 0.0f };
        const f4 err = { fsub(c.x, estB.x), fsub(c.y, estB.y), fsub(c.z, estB.z), 0.0f };
        const float facC = fmul(dpn<CH>(err, dirC), invLenC);
        minC = sse_min(minC, facC);
        maxC = sse_max(maxC, facC);
      }
    }
    else
    {
      // ---- RGBA: B extents + direction C as a third mean error direction ------------------------------
      acc = 0.0f;

      for (uint32_t base = 0; base < n; base += stagePx)
      {
        const uint32_t cnt = min((uint32_t)stagePx, n - base);

        for (uint32_t k = t; k < cnt; k += THREADS)
        {
          const f4 c = px_to_f4(px[base + k]);
          const f4 toAvg = { fsub(c.x, avg.x), fsub(c.y, avg.y), fsub(c.z, avg.z), fsub(c.w, avg.w) };
          const float facA = fmul(dpn<CH>(toAvg, dirA), invLenA);
          const f4 estA = { fadd(avg.x, fmul(facA, dirA.x)), fadd(avg.y, fmul(facA, dirA.y)), fadd(avg.z, fmul(facA, dirA.z)), fadd(avg.w, fmul(facA, dirA.w)) };
          const f4 toPx = { fsub(c.x, estA.x), fsub(c.y, estA.y), fsub(c.z, estA.z), fsub(c.w, estA.w) };
          const float facB = fmul(dpn<CH>(toPx, dirB), invLenB);
          minB = sse_min(minB, facB);
          maxB = sse_max(maxB, facB);
          const f4 estB = { fadd(estA.x, fmul(facB, dirB.x)), fadd(estA.y, fmul(facB, dirB.y)), fadd(estA.z, fmul(facB, dirB.z)), fadd(estA.w, fmul(facB, dirB.w)) };
          const f4 err = { fsub(c.x, estB.x), fsub(c.y, estB.y), fsub(c.z, estB.z), fsub(c.w, estB.w) };
          const f4 u = unit_direction<CH>(err, lut);
          stage[k] = make_float4(u.x, u.y, u.z, u.w);
        }

        group_sync<WARPS>();
        ordered_accumulate<WARPS>(stage, (int)cnt, acc);
        group_sync<WARPS>();
      }

      dirC = broadcast_acc<WARPS>(acc, invCount, gs);
      const float invLenC = frcp1(dpn<CH>(dirC, dirC));

      // Reference quirk (limg_factorization.h:745-758): the estimate pointer is rewound but never advanced, so every
      // pixel is measured against the A+B estimate of pixel 0.
      f4 est0;
      {
        const f4 c = px_to_f4(px[0]);
        const f4 toAvg = { fsub(c.x, avg.x), fsub(c.y, avg.y), fsub(c.z, avg.z), fsub(c.w, avg.w) };
        const float facA = fmul(dpn<CH>(toAvg, dirA), invLenA);
        const f4 estA = { fadd(avg.x, fmul(facA, dirA.x)), fadd(avg.y, fmul(facA, dirA.y)), fadd(avg.z, fmul(facA, dirA.z)), fadd(avg.w, fmul(facA, dirA.w)) };
        const f4 toPx = { fsub(c.x, estA.x), fsub(c.y, estA.y), fsub(c.z, estA.z), fsub(c.w, estA.w) };
        const float facB = fmul(dpn<CH>(toPx, dirB), invLenB);
        est0 = { fadd(estA.x, fmul(facB, dirB.x)), fadd(estA.y, fmul(facB, dirB.y)), fadd(estA.z, fmul(facB, dirB.z)), fadd(estA.w, fmul(facB, dirB.w)) };
      }

      for (uint32_t i = t; i < n; i += THREADS)
      {
        const f4 c = px_to_f4(px[i]);
        const f4 toPx = { fsub(c.x, est0.x), fsub(c.y, est0.y), fsub(c.z, est0.z), fsub(c.w, est0.w) };
        const float facC = fmul(dpn<CH>(toPx, dirC), invLenC);
        minC = sse_min(minC, facC);
        maxC = sse_max(maxC, facC);
      }
    }

    group_minmax<WARPS>(minB, maxB, gs, parity);
    group_minmax<WARPS>(minC, maxC, gs, parity);
  }

  const float av[4] = { avg.x, avg.y, avg.z, avg.w };
  const float dA[4] = { dirA.x, dirA.y, dirA.z, dirA.w };
  const float dB[4] = { dirB.x, dirB.y, dirB.z, dirB.w };
  const float dC[4] = { dirC.x, dirC.y, dirC.z, dirC.w };

#pragma unroll
  for (int c = 0; c < 4; c++)
  {
    const bool on = c < CH;
    out.avg[c] = on ? av[c] : 0.0f;
    out.dirA_min[c] = on ? (int16_t)sse_cvtps(fadd(av[c], fmul(minA, dA[c]))) : (int16_t)0;
    out.dirA_max[c] = on ? (int16_t)sse_cvtps(fadd(av[c], fmul(maxA, dA[c]))) : (int16_t)0;
    out.dirB_offset[c] = on ? (int16_t)sse_cvtps(fmul(minB, dB[c])) : (int16_t)0;
    out.dirB_mag[c] = on ? (int16_t)sse_cvtps(fmul(maxB, dB[c])) : (int16_t)0;
    out.dirC_offset[c] = on ? (int16_t)sse_cvtps(fmul(minC, dC[c])) : (int16_t)0;
    out.dirC_mag[c] = on ? (int16_t)sse_cvtps(fmul(maxC, dC[c])) : (int16_t)0;
  }
}

// ---------------------------------------------------------------------------------------------
// bit-crush trial over a pixel list (limg_bit_crush_simd.h:315-810). px / fac are area-contiguous; fac packs
// fa | fb << 8 | fc << 16. Returns pass; blockError (sign-extended 32-bit wrapping sum) is valid when no pixel failed.
// ---------------------------------------------------------------------------------------------

template <int CH, int WARPS>
__device__ __forceinline__ bool group_trial(const uint32_t *px, const uint32_t *fac, uint32_t n, const limgcu_decomp &d, int sA, int sB, int sC,
                                            const CrushParams &cp, GroupScratch<WARPS> *gs, uint32_t &parity, uint64_t &blockError)
{
  constexpr int THREADS = WARPS * 32;
  Recon r;
  init_recon<CH>(d, sA, sB, sC, 0, r);

  int32_t acc = 0;
  bool fail = false;
  auto one = [&](uint32_t i) {
    const uint32_t f = fac[i];
    const int32_t eA = (int32_t)((f & 0xFF) >> sA), eB = (int32_t)(((f >> 8) & 0xFF) >> sB), eC = (int32_t)(((f >> 16) & 0xFF) >> sC);
    const int32_t err = trial_error(px[i], recon_channel(r, 0, eA, eB, eC), recon_channel(r, 1, eA, eB, eC), recon_channel(r, 2, eA, eB, eC));
    acc += err;
    fail |= (uint64_t)(int64_t)err > cp.maxPixelError;
  };

  constexpr uint32_t CHUNK = THREADS * 4;

  if (WARPS > 1 && n > 2 * CHUNK)
  {
    // Big areas: a trial that fails does so because some pixel exceeds the per-pixel bound (the reference returns at the first such pixel,
    // limg_bit_crush_simd.h:431-440, without a block error), and with too coarse a shift that happens within the first few hundred pixels:
    // the CTA looks at the flag after every CHUNK pixels and stops. A passing trial is unchanged (same sum).
    for (uint32_t base = 0; base < n; base += CHUNK)
    {
      const uint32_t end = min(n, base + CHUNK);

      for (uint32_t i = base + group_tid<WARPS>(); i < end; i += THREADS)
        one(i);

      if (__syncthreads_or(fail))
        return false;
    }
  }
  else
  {
    for (uint32_t i = group_tid<WARPS>(); i < n; i += THREADS)
      one(i);
  }

  group_sum_any<WARPS>(acc, fail, gs, parity);

  if (fail)
    return false;

  blockError = (uint64_t)(int64_t)acc;
  return (blockError * 0x10) < cp.maxBlockError * (uint64_t)n;
}

// ---------------------------------------------------------------------------------------------
// shift search (limg_bit_crush.h:331-1051): the reference's sequential decision procedure, each trial a group reduction.
// ---------------------------------------------------------------------------------------------

template <class Trial>
__device__ __forceinline__ void search_shifts(Trial &&trial, bool fast, int shift[3])
{
  uint64_t err = 0, minErr = ~0ull;
  int maxShift = 0;
  shift[0] = shift[1] = shift[2] = 0;

  // fixed guesses (limg_bit_crush.h:331-392)
  if (trial(4, 5, 6, err))
  {
    shift[0] = 4; shift[1] = 5; shift[2] = 6; minErr = err; maxShift = 15;

    if (trial(5, 8, 8, err))
    {
      shift[0] = 5; shift[1] = 8; shift[2] = 8; minErr = err; maxShift = 21;
    }
    else if (trial(4, 6, 8, err))
    {
      shift[0] = 4; shift[1] = 6; shift[2] = 8; minErr = err; maxShift = 18;
    }
  }
  else if (trial(2, 4, 5, err))
  {
    shift[0] = 2; shift[1] = 4; shift[2] = 5; minErr = err; maxShift = 11;
  }

  if (fast)
  {
    // coarse lattice, step 2 (limg_bit_crush.h:510-556)
    {
      int a = shift[0] & 15, b = shift[1] & 15, c = (shift[2] & 15) + 2;

      for (; a <= 8; a += 2)
      {
        for (; b <= 8; b += 2)
        {
          for (; c <= 8; c += 2)
          {
            if (a + b + c > maxShift)
            {
              if (!trial(a, b, c, err))
                break;

              shift[0] = a; shift[1] = b; shift[2] = c;
              maxShift = a + b + c;
              minErr = err;
            }
          }

          if (c == b)
            break;

          c = b;
        }

        if (b == a)
          break;

        b = a;
      }
    }

    // fine pass, +0/+1 per even axis (limg_bit_crush.h:558-614)
    {
      const int preA = shift[0], preB = shift[1], preC = shift[2];
      const int limA = (!(preA & 1) && preA != 8) ? 1 : 0, limB = (!(preB & 1) && preB != 8) ? 1 : 0, limC = (!(preC & 1) && preC != 8) ? 1 : 0;
      int fine = 0;
      int a = 0, b = 0, c = 1;

      for (; a <= limA; a++)
      {
        for (; b <= limB; b++)
        {
          for (; c <= limC; c++)
          {
            if (a + b + c > fine)
            {
              if (!trial(preA + a, preB + b, preC + c, err))
                break;

              shift[0] = preA + a; shift[1] = preB + b; shift[2] = preC + c;
              maxShift = shift[0] + shift[1] + shift[2];
              fine = a + b + c;
              minErr = err;
            }
          }

          if (c == 0)
            break;

          c = 0;
        }

        if (b == 0)
          break;

        b = 0;
      }
    }
  }
  else
  {
    // --accurate-bit-crushing: exhaustive walk (limg_bit_crush.h:732-778) ...
    {
      int a = 0, b = 0, c = 1;

      for (; a <= 8; a++)
      {
        for (; b <= 8; b++)
        {
          for (; c <= 8; c++)
          {
            if (a + b + c > maxShift && (a != shift[0] || b != shift[1] || c != shift[2]))
            {
              if (!trial(a, b, c, err))
                break;

              shift[0] = a; shift[1] = b; shift[2] = c;
              maxShift = a + b + c;
              minErr = err;
            }
          }

          if (c == 0)
            break;

          c = 0;
        }

        if (b == 0)
          break;

        b = 0;
      }
    }

    // ... then equal-sum alternatives with a lower block error (limg_bit_crush.h:780-829)
    if (maxShift > 0)
    {
      int a = shift[0], b = shift[1], c = shift[2] + 1;

      for (; a <= 8; a++)
      {
        for (; b <= 8; b++)
        {
          for (; c <= 8; c++)
          {
            if (a + b + c == maxShift)
            {
              if (!trial(a, b, c, err))
                break;

              if (minErr > err)
              {
                shift[0] = a; shift[1] = b; shift[2] = c;
                minErr = err;
              }
            }
          }

          if (c == 0)
            break;

          c = 0;
        }

        if (b == 0)
          break;

        b = 0;
      }
    }
  }
}

} // namespace limg

/* oracle/limg_oracle.c -- see limg_oracle.h. TEST INFRASTRUCTURE ONLY; parity PINNED against oracle/_ref.
 *
 * Scalar C restatement of the reference's SSE4.1 path. Build with -ffp-contract=off and without
 * -ffast-math / -march=native (oracle/Makefile): every float operation below is meant to be one IEEE
 * binary32 operation in exactly the written order.
 */
#include "limg_oracle.h"
#include "rsqrt_lut.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LO_BLOCK 8u

/* ------------------------------------------------------------------------------------------------
 * SSE semantics helpers
 * --------------------------------------------------------------------------------------------- */

static const uint16_t *g_lut = LIMG_RSQRT_LUT;

void lo_set_rsqrt_lut(const uint16_t *lut2048) { g_lut = lut2048 ? lut2048 : LIMG_RSQRT_LUT; }

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* RSQRTSS: table function of (exponent parity, top 10 mantissa bits); denormals are treated as zero. */
float lo_rsqrt(float x)
{
  const uint32_t b = f2u(x);
  const uint32_t e = (b >> 23) & 0xFF;
  const uint32_t m = b & 0x7FFFFF;

  if (e == 0xFF)
    return m ? u2f(b | 0x00400000u) : ((b >> 31) ? u2f(0xFFC00000u) : 0.0f);
  if (e == 0)
    return u2f((b & 0x80000000u) | 0x7F800000u);
  if (b >> 31)
    return u2f(0xFFC00000u);

  const int E = (int)e - 127;
  const int p = E & 1;          /* two's complement: also right for negative E */
  const int k = (E - p) / 2;    /* exact */
  const uint32_t mant12 = g_lut[(uint32_t)p * 1024u + (m >> 13)];
  return u2f(((uint32_t)(126 - k) << 23) | (mant12 << 11));
}

/* MINPS / MAXPS return the SECOND operand unless the strict comparison holds (NaN and +-0 included). */
static inline float sse_min(float a, float b) { return a < b ? a : b; }
static inline float sse_max(float a, float b) { return a > b ? a : b; }

/* CVTPS2DQ, round-to-nearest-even, 0x80000000 for NaN / out of range. */
static inline int32_t sse_cvtps(float x)
{
  if (!(x >= -2147483648.0f && x < 2147483648.0f))
    return INT32_MIN;
  return (int32_t)lrintf(x);
}

/* DPPS: products rounded individually, summed as (p0+p1)+(p2+p3); a masked-out lane contributes +0. */
static inline float dp3(const float a[4], const float b[4])
{
  const float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2];
  const float lo = p0 + p1, hi = p2 + 0.0f;
  return lo + hi;
}

static inline float dp4(const float a[4], const float b[4])
{
  const float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2], p3 = a[3] * b[3];
  const float lo = p0 + p1, hi = p2 + p3;
  return lo + hi;
}

static inline float dpn(int ch, const float a[4], const float b[4]) { return ch == 3 ? dp3(a, b) : dp4(a, b); }

static inline void px_to_f4(uint32_t px, float out[4])
{
  out[0] = (float)(px & 0xFF);
  out[1] = (float)((px >> 8) & 0xFF);
  out[2] = (float)((px >> 16) & 0xFF);
  out[3] = (float)(px >> 24);
}

/* ------------------------------------------------------------------------------------------------
 * K1 + K2/K3: channel sums and the three-factor fit
 * --------------------------------------------------------------------------------------------- */

/* Adds the sign-normalised unit vector of v to acc (limg_factorization.h:412-431). */
static void add_unit_direction(int ch, const float v[4], float acc[4])
{
  if (v[0] == 0.0f && v[1] == 0.0f && v[2] == 0.0f && v[3] == 0.0f)
    return;

  const float e = FLT_EPSILON;
  const float bias[4] = { e * 3, e * 2, e * 1, 0.0f };
  float lo[4], hi[4];

  for (int i = 0; i < 4; i++)
  {
    lo[i] = v[i] - bias[i];
    hi[i] = v[i] + bias[i];
  }

  const float halfMin0 = sse_min(lo[0], lo[2]), halfMin1 = sse_min(lo[1], lo[3]);
  const float halfMax0 = sse_max(hi[0], hi[2]), halfMax1 = sse_max(hi[1], hi[3]);
  const float absMin = fabsf(sse_min(halfMin0, halfMin1));
  const float max = sse_max(halfMax0, halfMax1);

  float invLength = lo_rsqrt(dpn(ch, v, v));

  if (absMin > max)
    invLength = u2f(f2u(invLength) ^ 0x80000000u);

  for (int i = 0; i < 4; i++)
  {
    const float val = v[i] * invLength;
    acc[i] = acc[i] + val;
  }
}

void lo_fit(const uint32_t *pixels, size_t n, int ch, lo_decomp *out)
{
  /* K1 (limg.cpp:469-497): 32-bit wrapping sums, re-read as SIGNED by the fit (limg_factorization.h:402). */
  uint32_t sum[4] = { 0, 0, 0, 0 };

  for (size_t i = 0; i < n; i++)
  {
    sum[0] += pixels[i] & 0xFF;
    sum[1] += (pixels[i] >> 8) & 0xFF;
    sum[2] += (pixels[i] >> 16) & 0xFF;
    sum[3] += pixels[i] >> 24;
  }

  const float invCount = 1.f / (float)n;
  float avg[4];

  for (int i = 0; i < 4; i++)
    avg[i] = (float)(int32_t)sum[i] * invCount;

  if (ch == 3)
    avg[3] = 0.0f; /* lane 3 is never observable on the RGB path (masked / excluded from every dot product). */

  float dirA[4] = { 0, 0, 0, 0 }, dirB[4] = { 0, 0, 0, 0 }, dirC[4] = { 0, 0, 0, 0 };
  float minA = 0, maxA = 0, minB = 0, maxB = 0, minC = 0, maxC = 0; /* Q11 */

  for (size_t i = 0; i < n; i++)
  {
    float px[4], corrected[4];
    px_to_f4(pixels[i], px);

    for (int c = 0; c < 4; c++)
      corrected[c] = px[c] - avg[c];

    if (ch == 3)
      corrected[3] = 0.0f;

    add_unit_direction(ch, corrected, dirA);
  }

  for (int c = 0; c < 4; c++)
    dirA[c] = dirA[c] * invCount;

  if (!(dirA[0] == 0.0f && dirA[1] == 0.0f && dirA[2] == 0.0f && dirA[3] == 0.0f))
  {
    float *est = (float *)malloc((n ? n : 1) * 4 * sizeof(float));
    const float invLenA = 1.f / dpn(ch, dirA, dirA);

    for (size_t i = 0; i < n; i++)
    {
      float px[4], toPx[4], err[4];
      px_to_f4(pixels[i], px);

      for (int c = 0; c < 4; c++)
        toPx[c] = px[c] - avg[c];

      const float facA = dpn(ch, toPx, dirA) * invLenA;
      minA = sse_min(minA, facA);
      maxA = sse_max(maxA, facA);

      for (int c = 0; c < 4; c++)
      {
        const float step = facA * dirA[c];
        est[i * 4 + c] = avg[c] + step;
        err[c] = px[c] - est[i * 4 + c];
      }

      if (ch == 3)
        err[3] = 0.0f;

      add_unit_direction(ch, err, dirB);
    }

    for (int c = 0; c < 4; c++)
      dirB[c] = dirB[c] * invCount;

    if (ch == 3)
    {
      /* Q5: C = A x B, each product rounded on its own (limg_factorization.h:498-507). */
      const float m0 = dirA[1] * dirB[2], m1 = dirA[2] * dirB[0], m2 = dirA[0] * dirB[1];
      const float s0 = dirA[2] * dirB[1], s1 = dirA[0] * dirB[2], s2 = dirA[1] * dirB[0];
      dirC[0] = m0 - s0;
      dirC[1] = m1 - s1;
      dirC[2] = m2 - s2;
      dirC[3] = 0.0f;

      const float invLenB = 1.f / dp3(dirB, dirB);
      const float invLenC = 1.f / dp3(dirC, dirC);

      minB = minC = FLT_MAX;
      maxB = maxC = -FLT_MAX;

      for (size_t i = 0; i < n; i++)
      {
        float px[4], toPx[4], err[4];
        px_to_f4(pixels[i], px);

        for (int c = 0; c < 4; c++)
          toPx[c] = px[c] - est[i * 4 + c];

        const float facB = dp3(toPx, dirB) * invLenB;
        minB = sse_min(minB, facB);
        maxB = sse_max(maxB, facB);

        for (int c = 0; c < 4; c++)
        {
          const float step = facB * dirB[c];
          const float estB = est[i * 4 + c] + step;
          err[c] = px[c] - estB;
        }

        const float facC = dp3(err, dirC) * invLenC;
        minC = sse_min(minC, facC);
        maxC = sse_max(maxC, facC);
      }
    }
    else
    {
      const float invLenB = 1.f / dp4(dirB, dirB);

      minB = minC = FLT_MAX;
      maxB = maxC = -FLT_MAX;

      for (size_t i = 0; i < n; i++)
      {
        float px[4], toPx[4], err[4];
        px_to_f4(pixels[i], px);

        for (int c = 0; c < 4; c++)
          toPx[c] = px[c] - est[i * 4 + c];

        const float facB = dp4(toPx, dirB) * invLenB;
        minB = sse_min(minB, facB);
        maxB = sse_max(maxB, facB);

        for (int c = 0; c < 4; c++)
        {
          const float step = facB * dirB[c];
          est[i * 4 + c] = est[i * 4 + c] + step;
          err[c] = px[c] - est[i * 4 + c];
        }

        add_unit_direction(4, err, dirC);
      }

      for (int c = 0; c < 4; c++)
        dirC[c] = dirC[c] * invCount;

      const float invLenC = 1.f / dp4(dirC, dirC);

      for (size_t i = 0; i < n; i++)
      {
        float px[4], toPx[4];
        px_to_f4(pixels[i], px);

        /* Reference quirk (limg_factorization.h:745-758): the estimate pointer is rewound but never advanced in this
         * loop, so every pixel is measured against the A+B estimate of pixel 0. Reproduced, not fixed. */
        for (int c = 0; c < 4; c++)
          toPx[c] = px[c] - est[c];

        const float facC = dp4(toPx, dirC) * invLenC;
        minC = sse_min(minC, facC);
        maxC = sse_max(maxC, facC);
      }
    }

    free(est);
  }

  memset(out, 0, sizeof(*out));

  for (int c = 0; c < ch; c++)
  {
    out->avg[c] = avg[c];
    out->dirA_min[c] = (int16_t)sse_cvtps(avg[c] + minA * dirA[c]);
    out->dirA_max[c] = (int16_t)sse_cvtps(avg[c] + maxA * dirA[c]);
    out->dirB_offset[c] = (int16_t)sse_cvtps(minB * dirB[c]);
    out->dirB_mag[c] = (int16_t)sse_cvtps(maxB * dirB[c]);
    out->dirC_offset[c] = (int16_t)sse_cvtps(minC * dirC[c]);
    out->dirC_mag[c] = (int16_t)sse_cvtps(maxC * dirC[c]);
  }
}

/* ------------------------------------------------------------------------------------------------
 * projection state (limg_internal.h:426-452)
 * --------------------------------------------------------------------------------------------- */

typedef struct
{
  float nA[4], nB[4], nC[4];
  float invA, invB, invC;
} proj_state;

static float seq_dot(int ch, const float a[4], const float b[4])
{
  float s = 0;
  for (int i = 0; i < ch; i++)
    s += a[i] * b[i];
  return s;
}

static void init_proj_state(int ch, const lo_decomp *d, proj_state *s)
{
  int nzA = 0, nzB = 0, nzC = 0;
  memset(s, 0, sizeof(*s));

  for (int i = 0; i < ch; i++)
  {
    s->nA[i] = (float)((int)d->dirA_max[i] - (int)d->dirA_min[i]);
    s->nB[i] = (float)((int)d->dirB_mag[i] - (int)d->dirB_offset[i]);
    s->nC[i] = (float)((int)d->dirC_mag[i] - (int)d->dirC_offset[i]);
    nzA |= s->nA[i] != 0;
    nzB |= s->nB[i] != 0;
    nzC |= s->nC[i] != 0;
  }

  if (nzA) s->invA = 1.f / seq_dot(ch, s->nA, s->nA);
  if (nzB) s->invB = 1.f / seq_dot(ch, s->nB, s->nB);
  if (nzC) s->invC = 1.f / seq_dot(ch, s->nC, s->nC);
}

/* limg_factorization.h:9-41, scalar float version used by the merge predicate. */
static void scalar_factors(int ch, const float color[4], const lo_decomp *d, const proj_state *s, float *fa, float *fb, float *fc)
{
  float t[4], est[4];

  for (int i = 0; i < ch; i++)
    t[i] = color[i] - (float)d->dirA_min[i];

  const float facA = seq_dot(ch, t, s->nA) * s->invA;

  for (int i = 0; i < ch; i++)
  {
    est[i] = (float)d->dirA_min[i] + facA * s->nA[i];
    t[i] = (color[i] - est[i]) - (float)d->dirB_offset[i];
  }

  const float facB = seq_dot(ch, t, s->nB) * s->invB;

  for (int i = 0; i < ch; i++)
  {
    est[i] = est[i] + facB * s->nB[i];
    t[i] = (color[i] - est[i]) - (float)d->dirC_offset[i];
  }

  *fa = facA;
  *fb = facB;
  *fc = seq_dot(ch, t, s->nC) * s->invC;
}

/* ------------------------------------------------------------------------------------------------
 * K5: merge predicate (limg.cpp:1137-1269), quirks Q1 kept.
 * --------------------------------------------------------------------------------------------- */

static uint64_t g_full_predicates = 0, g_predicates = 0;

int lo_matches(int ch, const lo_decomp *a, const lo_decomp *b)
{
  proj_state sa, sb;
  init_proj_state(ch, a, &sa);
  init_proj_state(ch, b, &sb);
  g_predicates++;

  const float weight[4] = { 2, 4, 3, 3 };
  float avgDiffSq = 0;
  float lenA[3] = { 3, 3, 3 }, lenB[3] = { 3, 3, 3 };

  for (int i = 0; i < ch; i++)
  {
    const float diff = a->avg[i] - b->avg[i];
    avgDiffSq += diff * diff * weight[i];
    lenA[0] += (sa.nA[i] * sa.nA[i]) * weight[i];
    lenB[0] += (sb.nA[i] * sb.nA[i]) * weight[i];
    lenA[1] += (sa.nB[i] * sa.nB[i]) * weight[i];
    lenB[1] += (sb.nB[i] * sb.nB[i]) * weight[i];
    lenA[2] += (sa.nC[i] * sa.nC[i]) * weight[i];
    lenB[2] += (sb.nC[i] * sb.nC[i]) * weight[i];
  }

  const float sumA = lenA[0] + lenA[1] + lenA[2];
  const float sumB = lenB[0] + lenB[1] + lenB[2];
  const float ratio = (sumA + 1) / (sumB + 1);
  const float acceptAvg = (float)(16 * 3 * ch);
  const float acceptRange = (float)(200 * 3 * ch);

  if (avgDiffSq < acceptAvg && sumA < acceptRange && sumB < acceptRange)
    return 1;

  const float maxRatio = 1.375f;

  if (ratio > maxRatio || ratio < (1.f / maxRatio))
    return 0;

  g_full_predicates++;

  float invA[3], invB[3];

  for (int i = 0; i < 3; i++)
  {
    invA[i] = 1.f / lenA[i];
    invB[i] = 1.f / lenB[i];
  }

  for (int i = 1; i < 3; i++)
  {
    invA[i] *= 2.f;
    invB[i] *= 2.f;
  }

  /* Q1: the second projection of every iteration is avg(a) into b -- loop invariant, but it stays inside the ordered sum. */
  float ga, gb, gc;
  scalar_factors(ch, a->avg, b, &sb, &ga, &gb, &gc);
  const float constTerm = fabsf(ga) * invB[0] + fabsf(0.5f - gb) * invB[1] + fabsf(0.5f - gc) * invB[2];

  float sum = 0;

  for (int z = 0; z < 3; z++)
  {
    const float zf = z * 0.5f;

    for (int y = 0; y < 3; y++)
    {
      const float yf = y * 0.5f;

      for (int x = 0; x < 3; x++)
      {
        const float xf = x * 0.5f;
        float color[4] = { 0, 0, 0, 0 };
        float fa, fb, fc;

        for (int i = 0; i < ch; i++)
          color[i] = sb.nA[i] * xf + sb.nB[i] * yf + sb.nC[i] * zf;

        scalar_factors(ch, color, a, &sa, &fa, &fb, &fc);
        sum += fabsf(fa) * invA[0] + fabsf(0.5f - fb) * invA[1] + fabsf(0.5f - fc) * invA[2];
        sum += constTerm;
      }
    }
  }

  const float avgSum = sum * (1.f / (3 * 3 * 3));
  return avgSum < 3.0f;
}

/* ------------------------------------------------------------------------------------------------
 * K4: projection onto the factors (limg_factorization.h:101-197)
 * --------------------------------------------------------------------------------------------- */

static inline uint8_t factor_to_u8(float f)
{
  int32_t v = sse_cvtps(255.0f * f);
  if (v > 0xFF) v = 0xFF;   /* pminsd(0xFF, v) */
  if (v < 0) v = 0;         /* pmaxsd(0, .): the 0x80000000 indefinite value ends up as 0 */
  return (uint8_t)v;
}

void lo_project(int ch, const lo_decomp *d, const uint32_t *pixels, size_t n, uint8_t *fa, uint8_t *fb, uint8_t *fc)
{
  proj_state s;
  init_proj_state(ch, d, &s);

  float minA[4], offB[4], offC[4];

  for (int i = 0; i < 4; i++)
  {
    minA[i] = (float)d->dirA_min[i];
    offB[i] = (float)d->dirB_offset[i];
    offC[i] = (float)d->dirC_offset[i];
  }

  for (size_t k = 0; k < n; k++)
  {
    float px[4], t[4], est[4];
    px_to_f4(pixels[k], px);

    for (int i = 0; i < 4; i++)
      t[i] = px[i] - minA[i];

    const float facA = dpn(ch, t, s.nA) * s.invA;
    fa[k] = factor_to_u8(facA);

    for (int i = 0; i < 4; i++)
    {
      est[i] = minA[i] + s.nA[i] * facA;
      t[i] = (px[i] - est[i]) - offB[i];
    }

    const float facB = dpn(ch, t, s.nB) * s.invB;
    fb[k] = factor_to_u8(facB);

    for (int i = 0; i < 4; i++)
    {
      est[i] = est[i] + s.nB[i] * facB;
      t[i] = (px[i] - est[i]) - offC[i];
    }

    const float facC = dpn(ch, t, s.nC) * s.invC;
    fc[k] = factor_to_u8(facC);
  }
}

/* ------------------------------------------------------------------------------------------------
 * integer reconstruction shared by K6 (trial) and K8 (decode)
 * --------------------------------------------------------------------------------------------- */

static const int32_t DECODE_MUL[9] = { 1, 2, 4, 8, 17, 36, 85, 255, 256 }; /* (1<<s) + bias[s], bias = {0,0,0,0,1,4,21,127,0} (Q8) */

typedef struct
{
  int32_t nA[4], nB[4], nC[4];
  int32_t mA[4], mB[4], mC[4]; /* (min << 8) + 128 */
  int32_t mul[3];
} recon_state;

/* alphaMin: what the 4th lane's `min` is set to: the decoder of RGB images forces 0xFFFF (limg_decode.h:95-97), the trial uses 0. */
static void init_recon(int ch, const lo_decomp *d, const uint8_t shift[3], int32_t rgbAlphaMin, recon_state *r)
{
  int32_t minA[4], minB[4], minC[4];

  for (int i = 0; i < 4; i++)
  {
    r->nA[i] = (int32_t)d->dirA_max[i] - d->dirA_min[i];
    r->nB[i] = (int32_t)d->dirB_mag[i] - d->dirB_offset[i];
    r->nC[i] = (int32_t)d->dirC_mag[i] - d->dirC_offset[i];
    minA[i] = d->dirA_min[i];
    minB[i] = d->dirB_offset[i];
    minC[i] = d->dirC_offset[i];
  }

  if (ch == 3)
  {
    r->nA[3] = r->nB[3] = r->nC[3] = 0;
    minA[3] = minB[3] = minC[3] = rgbAlphaMin;
  }

  /* Q7: a dropped factor (shift 8) only clears the first three channels. */
  if (shift[0] > 7)
    for (int i = 0; i < 3; i++) r->nA[i] = 0;

  if (shift[1] > 7)
    for (int i = 0; i < 3; i++) { r->nB[i] = 0; minB[i] = 0; }

  if (shift[2] > 7)
    for (int i = 0; i < 3; i++) { r->nC[i] = 0; minC[i] = 0; }

  for (int i = 0; i < 4; i++)
  {
    r->mA[i] = (int32_t)((uint32_t)minA[i] << 8) + 128;
    r->mB[i] = (int32_t)((uint32_t)minB[i] << 8) + 128;
    r->mC[i] = (int32_t)((uint32_t)minC[i] << 8) + 128;
  }

  for (int i = 0; i < 3; i++)
    r->mul[i] = DECODE_MUL[shift[i]];
}

static inline int32_t mul32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); } /* pmulld */
static inline int32_t add32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

static inline void recon_px(const recon_state *r, uint32_t eA, uint32_t eB, uint32_t eC, int32_t col[4])
{
  const int32_t dA = mul32((int32_t)eA, r->mul[0]);
  const int32_t dB = mul32((int32_t)eB, r->mul[1]);
  const int32_t dC = mul32((int32_t)eC, r->mul[2]);

  for (int i = 0; i < 4; i++)
  {
    int32_t v = add32(mul32(dA, r->nA[i]), r->mA[i]) >> 8;
    v = add32(v, add32(mul32(dB, r->nB[i]), r->mB[i]) >> 8);
    v = add32(v, add32(mul32(dC, r->nC[i]), r->mC[i]) >> 8);
    col[i] = v < 0 ? 0 : (v > 0xFF ? 0xFF : v);
  }
}

/* K8 */
void lo_decode(int ch, uint32_t *out, size_t stride, size_t rx, size_t ry, const uint8_t *fa, const uint8_t *fb, const uint8_t *fc, const lo_decomp *d, const uint8_t shift[3])
{
  recon_state r;
  init_recon(ch, d, shift, 0xFFFF, &r);

  for (size_t y = 0; y < ry; y++)
  {
    for (size_t x = 0; x < rx; x++)
    {
      int32_t col[4];
      recon_px(&r, *fa++, *fb++, *fc++, col);
      out[y * stride + x] = (uint32_t)col[0] | ((uint32_t)col[1] << 8) | ((uint32_t)col[2] << 16) | ((uint32_t)col[3] << 24);
    }
  }
}

/* K6. Returns pass; *blockError is written only where the reference writes it (not on a per-pixel failure). */
int lo_trial(int ch, uint64_t maxPixelError, uint64_t maxBlockError, const lo_decomp *d, const uint32_t *pixels, size_t n,
             const uint8_t *fa, const uint8_t *fb, const uint8_t *fc, const uint8_t shift[3], uint64_t *blockError)
{
  recon_state r;
  init_recon(ch, d, shift, 0, &r);

  int32_t acc = 0; /* 32-bit lane, wraps (limg_bit_crush_simd.h:383,422) */

  for (size_t k = 0; k < n; k++)
  {
    int32_t col[4];
    recon_px(&r, (uint32_t)(fa[k] >> shift[0]), (uint32_t)(fb[k] >> shift[1]), (uint32_t)(fc[k] >> shift[2]), col);

    const int32_t dr = (int32_t)(pixels[k] & 0xFF) - col[0];
    const int32_t dg = (int32_t)((pixels[k] >> 8) & 0xFF) - col[1];
    const int32_t db = (int32_t)((pixels[k] >> 16) & 0xFF) - col[2];
    const int32_t rr = dr * dr;
    const int lowRed = rr < 0x4000;
    /* Q6: the alpha lane's error never reaches lane 0 of the horizontal add, for RGB and RGBA alike. */
    const int32_t err = rr * (lowRed ? 2 : 3) + dg * dg * 4 + db * db * (lowRed ? 3 : 2);

    acc = add32(acc, err);

    if ((uint64_t)(int64_t)err > maxPixelError)
      return 0;
  }

  const uint64_t total = (uint64_t)(int64_t)acc; /* _mm_extract_epi32 returns int: sign-extended into size_t */
  *blockError = total;
  return (total * 0x10) < maxBlockError * (uint64_t)n;
}

/* ------------------------------------------------------------------------------------------------
 * shift search (limg_bit_crush.h:331-1051)
 * --------------------------------------------------------------------------------------------- */

typedef struct
{
  int ch;
  uint64_t maxPixel, maxBlock;
  const lo_decomp *d;
  const uint32_t *px;
  size_t n;
  const uint8_t *fa, *fb, *fc;
  uint64_t trials;
} search_ctx;

static int try_shift(search_ctx *s, int a, int b, int c, uint64_t *err)
{
  const uint8_t sh[3] = { (uint8_t)a, (uint8_t)b, (uint8_t)c };
  s->trials++;
  return lo_trial(s->ch, s->maxPixel, s->maxBlock, s->d, s->px, s->n, s->fa, s->fb, s->fc, sh, err);
}

/* The fixed guess tree (limg_bit_crush.h:331-392 / 677-730). Returns the accepted a+b+c (0 if none). */
static int guess_shift(search_ctx *s, uint8_t shift[3], uint64_t *minErr)
{
  uint64_t err = 0;
  int sum = 0;

  if (try_shift(s, 4, 5, 6, &err))
  {
    shift[0] = 4; shift[1] = 5; shift[2] = 6; *minErr = err; sum = 15;

    if (try_shift(s, 5, 8, 8, &err))
    {
      shift[0] = 5; shift[1] = 8; shift[2] = 8; *minErr = err; sum = 21;
    }
    else if (try_shift(s, 4, 6, 8, &err))
    {
      shift[0] = 4; shift[1] = 6; shift[2] = 8; *minErr = err; sum = 18;
    }
  }
  else if (try_shift(s, 2, 4, 5, &err))
  {
    shift[0] = 2; shift[1] = 4; shift[2] = 5; *minErr = err; sum = 11;
  }

  return sum;
}

/* Exhaustive lattice walk with step 1 (limg_bit_crush.h:402-448 / 732-778). */
static void walk_exhaustive(search_ctx *s, uint8_t shift[3], int *maxShift, uint64_t *minErr)
{
  int a = 0, b = 0, c = 1;
  uint64_t err = 0;

  for (; a <= 8; a++)
  {
    for (; b <= 8; b++)
    {
      for (; c <= 8; c++)
      {
        if (a + b + c > *maxShift && (a != shift[0] || b != shift[1] || c != shift[2]))
        {
          if (!try_shift(s, a, b, c, &err))
            break;

          shift[0] = (uint8_t)a; shift[1] = (uint8_t)b; shift[2] = (uint8_t)c;
          *maxShift = a + b + c;
          *minErr = err;
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

/* Coarse (step 2) then fine (+0/+1) walk (limg_bit_crush.h:510-614 / 895-999). maxShift is 8-bit in the first variant and
 * size_t in the second; both stay below 25 so the width does not matter. */
static void walk_coarse_fine(search_ctx *s, uint8_t shift[3], int *maxShift, uint64_t *minErr)
{
  uint64_t err = 0;

  {
    int a = shift[0] & 15, b = shift[1] & 15, c = (shift[2] & 15) + 2;

    for (; a <= 8; a += 2)
    {
      for (; b <= 8; b += 2)
      {
        for (; c <= 8; c += 2)
        {
          if (a + b + c > *maxShift)
          {
            if (!try_shift(s, a, b, c, &err))
              break;

            shift[0] = (uint8_t)a; shift[1] = (uint8_t)b; shift[2] = (uint8_t)c;
            *maxShift = a + b + c;
            *minErr = err;
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

  {
    const int preA = shift[0], preB = shift[1], preC = shift[2];
    const int limA = !(preA & 1) && preA != 8, limB = !(preB & 1) && preB != 8, limC = !(preC & 1) && preC != 8;
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
            if (!try_shift(s, preA + a, preB + b, preC + c, &err))
              break;

            shift[0] = (uint8_t)(preA + a); shift[1] = (uint8_t)(preB + b); shift[2] = (uint8_t)(preC + c);
            *maxShift = shift[0] + shift[1] + shift[2];
            fine = a + b + c;
            *minErr = err;
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

/* Equal-sum alternatives with a lower block error (limg_bit_crush.h:450-499 and its three copies). */
static void walk_equal_sum(search_ctx *s, uint8_t shift[3], int maxShift, uint64_t *minErr)
{
  int a = shift[0], b = shift[1], c = shift[2] + 1;
  uint64_t err = 0;

  for (; a <= 8; a++)
  {
    for (; b <= 8; b++)
    {
      for (; c <= 8; c++)
      {
        if (a + b + c == maxShift)
        {
          if (!try_shift(s, a, b, c, &err))
            break;

          if (*minErr > err)
          {
            shift[0] = (uint8_t)a; shift[1] = (uint8_t)b; shift[2] = (uint8_t)c;
            *minErr = err;
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

static uint64_t g_trials = 0;

void lo_search(int ch, uint32_t errorFactor, int fast, const lo_decomp *d, const uint32_t *pixels, size_t n,
               const uint8_t *fa, const uint8_t *fb, const uint8_t *fc, uint8_t shift[3])
{
  shift[0] = shift[1] = shift[2] = 0;

  if (errorFactor == 0) /* crushBits (limg.cpp:2349) */
    return;

  search_ctx s;
  s.ch = ch;
  s.maxPixel = (uint64_t)0x6 * (errorFactor / 2) * 7; /* limg.cpp:2344,2365 */
  s.maxBlock = (uint64_t)0x4 * (errorFactor / 2) * 7; /* limg.cpp:2345,2366 */
  s.d = d; s.px = pixels; s.n = n; s.fa = fa; s.fb = fb; s.fc = fc; s.trials = 0;

  uint64_t minErr = (uint64_t)-1;
  int maxShift = guess_shift(&s, shift, &minErr);

  if (fast)
  {
    /* guess + coarse/fine; the third phase is skipped (fastBitCrush, limg_bit_crush.h:617). */
    walk_coarse_fine(&s, shift, &maxShift, &minErr);
  }
  else
  {
    /* --accurate-bit-crushing: errorPixelRetaining && !coarseFine (limg.cpp:1516-1521): the rotated start index of the
     * extractPixel trial only changes which failing pixel is hit first, never the outcome. */
    walk_exhaustive(&s, shift, &maxShift, &minErr);

    if (maxShift > 0)
      walk_equal_sum(&s, shift, maxShift, &minErr);
  }

  g_trials += s.trials;
}

/* ------------------------------------------------------------------------------------------------
 * K7: dither (limg.cpp:798-887)
 * --------------------------------------------------------------------------------------------- */

static inline uint32_t pcg_output(uint64_t h)
{
  const uint32_t xs = (uint32_t)(((h >> 18) ^ h) >> 27);
  const uint32_t rot = (uint32_t)(h >> 59);
  return (xs >> rot) | (xs << ((32 - rot) & 31));
}

static inline uint8_t dither_one(uint8_t f, uint32_t rnd, uint8_t shift)
{
  const int32_t mask = (1 << shift) - 1;
  const int32_t offset = 1 << (shift - 1);
  int32_t v = (int32_t)f + (((int32_t)rnd & mask) - offset);
  v = v < 0 ? 0 : (v > 0xFF ? 0xFF : v);
  return (uint8_t)(v >> shift);
}

/* software AESDEC (InvShiftRows, InvSubBytes, InvMixColumns, xor key) */
static uint8_t g_inv_sbox[256];
static int g_inv_sbox_ready = 0;

static uint8_t gmul(uint8_t a, uint8_t b)
{
  uint8_t p = 0;
  for (int i = 0; i < 8; i++)
  {
    if (b & 1) p ^= a;
    const uint8_t hi = a & 0x80;
    a = (uint8_t)(a << 1);
    if (hi) a ^= 0x1B;
    b >>= 1;
  }
  return p;
}

static void build_inv_sbox(void)
{
  for (int x = 0; x < 256; x++)
  {
    uint8_t inv = 0;
    if (x)
      for (int y = 1; y < 256; y++)
        if (gmul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }

    uint8_t s = inv;
    for (int i = 1; i < 5; i++)
      s ^= (uint8_t)((inv << i) | (inv >> (8 - i)));
    s ^= 0x63;
    g_inv_sbox[s] = (uint8_t)x;
  }
  g_inv_sbox_ready = 1;
}

static void aesdec_round(uint8_t st[16], const uint8_t key[16])
{
  uint8_t t[16];

  /* InvShiftRows: row r (bytes r, r+4, r+8, r+12) rotates right by r columns. */
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
      t[((c + r) & 3) * 4 + r] = st[c * 4 + r];

  for (int i = 0; i < 16; i++)
    t[i] = g_inv_sbox[t[i]];

  for (int c = 0; c < 4; c++)
  {
    const uint8_t *p = &t[c * 4];
    st[c * 4 + 0] = gmul(p[0], 14) ^ gmul(p[1], 11) ^ gmul(p[2], 13) ^ gmul(p[3], 9);
    st[c * 4 + 1] = gmul(p[0], 9) ^ gmul(p[1], 14) ^ gmul(p[2], 11) ^ gmul(p[3], 13);
    st[c * 4 + 2] = gmul(p[0], 13) ^ gmul(p[1], 9) ^ gmul(p[2], 14) ^ gmul(p[3], 11);
    st[c * 4 + 3] = gmul(p[0], 11) ^ gmul(p[1], 13) ^ gmul(p[2], 9) ^ gmul(p[3], 14);
  }

  for (int i = 0; i < 16; i++)
    st[i] ^= key[i];
}

uint64_t lo_dither(int mode, uint8_t shift, size_t n, uint64_t state, uint8_t *f)
{
  if (shift > 7)
    return state;

  size_t i = 0;

  if (mode == LO_DITHER_AES && n >= 8)
  {
    if (!g_inv_sbox_ready)
      build_inv_sbox();

    const uint64_t keyLo = 0x824A73EAAB705E1DULL, keyHi = 0x2A76E98006CB4CADULL;
    uint8_t st[16], key[16];
    const uint64_t inv = ~state;
    memcpy(st, &state, 8);
    memcpy(st + 8, &inv, 8);
    memcpy(key, &keyLo, 8);
    memcpy(key + 8, &keyHi, 8);

    for (; i + 8 <= n; i += 8)
    {
      aesdec_round(st, key);

      for (int k = 0; k < 8; k++)
      {
        uint16_t lane;
        memcpy(&lane, st + 2 * k, 2);
        f[i + k] = dither_one(f[i + k], lane, shift);
      }
    }

    memcpy(&state, st, 8);
  }

  for (; i < n; i++)
  {
    state = state * 6364136223846793005ULL + 1;
    f[i] = dither_one(f[i], pcg_output(state), shift);
  }

  return state;
}

/* ------------------------------------------------------------------------------------------------
 * pass 1 (limg.cpp:1088-1119)
 * --------------------------------------------------------------------------------------------- */

static size_t gather(const uint32_t *img, size_t sizeX, size_t x0, size_t y0, size_t w, size_t h, uint32_t *dst)
{
  for (size_t y = 0; y < h; y++)
    memcpy(dst + y * w, img + (y0 + y) * sizeX + x0, w * sizeof(uint32_t));
  return w * h;
}

void lo_pass1(const uint32_t *img, size_t sizeX, size_t sizeY, int ch, lo_decomp *table)
{
  uint32_t px[LO_BLOCK * LO_BLOCK];
  const size_t bx = (sizeX + LO_BLOCK - 1) / LO_BLOCK, by = (sizeY + LO_BLOCK - 1) / LO_BLOCK;

  for (size_t j = 0; j < by; j++)
  {
    for (size_t i = 0; i < bx; i++)
    {
      const size_t w = sizeX - i * LO_BLOCK < LO_BLOCK ? sizeX - i * LO_BLOCK : LO_BLOCK;
      const size_t h = sizeY - j * LO_BLOCK < LO_BLOCK ? sizeY - j * LO_BLOCK : LO_BLOCK;
      const size_t n = gather(img, sizeX, i * LO_BLOCK, j * LO_BLOCK, w, h, px);
      lo_fit(px, n, ch, &table[j * bx + i]);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * greedy area map (limg.cpp:1121-1135, 1277-1496, 1814-1878), quirks Q2-Q4
 * --------------------------------------------------------------------------------------------- */

typedef struct
{
  const lo_decomp *table;
  size_t bx, by;
  int ch;
  uint8_t *used;
} merge_state;

/* every block of the strip is unused and matches the seed decomposition */
static int strip_joins(const merge_state *m, const lo_decomp *seed, size_t x0, size_t y0, size_t w, size_t h)
{
  for (size_t y = y0; y < y0 + h; y++)
    for (size_t x = x0; x < x0 + w; x++)
      if (m->used[y * m->bx + x])
        return 0;

  for (size_t y = y0; y < y0 + h; y++)
    for (size_t x = x0; x < x0 + w; x++)
      if (!lo_matches(m->ch, seed, &m->table[y * m->bx + x]))
        return 0;

  return 1;
}

typedef struct { size_t ox, oy, rx, ry; } rect;

/* Alternating one-block growth; the seed decomposition is the one of the rectangle's top-left block and never changes (Q2). */
static rect grow(const merge_state *m, rect r, int fourWay)
{
  const lo_decomp *seed = &m->table[r.oy * m->bx + r.ox];
  int right = 1, down = 1, up = fourWay, left = fourWay;

  while (right || down || up || left)
  {
    if (right)
    {
      if (r.ox + r.rx + 1 < m->bx && strip_joins(m, seed, r.ox + r.rx, r.oy, 1, r.ry)) /* Q3: the last column never joins */
        r.rx++;
      else
        right = 0;
    }

    if (down)
    {
      if (r.oy + r.ry + 1 < m->by && strip_joins(m, seed, r.ox, r.oy + r.ry, r.rx, 1))
        r.ry++;
      else
        down = 0;
    }

    if (up)
    {
      if (r.oy > 0 && strip_joins(m, seed, r.ox, r.oy - 1, r.rx, 1))
      {
        r.oy--;
        r.ry++;
      }
      else
        up = 0;
    }

    if (left)
    {
      if (r.ox > 0 && strip_joins(m, seed, r.ox - 1, r.oy, 1, r.ry))
      {
        r.ox--;
        r.rx++;
      }
      else
        left = 0;
    }
  }

  return r;
}

size_t lo_merge(const lo_decomp *table, size_t bx, size_t by, int ch, lo_area *areas, uint64_t *stats)
{
  merge_state m;
  m.table = table; m.bx = bx; m.by = by; m.ch = ch;
  m.used = (uint8_t *)calloc(bx * by > 0 ? bx * by : 1, 1);

  size_t count = 0;
  uint64_t st[8] = { 0 };
  const uint64_t pred0 = g_predicates, full0 = g_full_predicates;

  for (uint32_t stage = 0; stage < 2; stage++)
  {
    size_t x = 0, y = 0;

    while (y < by)
    {
      if (x >= bx)
      {
        x = 0;
        y++;
        continue;
      }

      if (m.used[y * bx + x])
      {
        x++;
        continue;
      }

      st[0]++; /* seeds examined */

      const rect seedRect = { x, y, 1, 1 };
      const rect r = grow(&m, seedRect, 0);
      rect emit = r;
      int take = 0, rescan = 0;

      if (stage == 0)
      {
        if (r.rx >= 3 && r.ry >= 3) /* Q4 */
        {
          const rect centre = { x + r.rx / 3, y + r.ry / 3, r.rx / 3, r.ry / 3 };
          const rect c = grow(&m, centre, 1);
          st[3]++; /* centre attempts */

          if (c.rx * c.ry > r.rx * r.ry)
          {
            emit = c;
            rescan = 1;
            st[4]++; /* centre successes */
          }

          take = 1;
        }
      }
      else
      {
        take = r.rx > 1 || r.ry > 1;
      }

      if (!take)
      {
        x++;
        continue;
      }

      for (size_t yy = emit.oy; yy < emit.oy + emit.ry; yy++)
        for (size_t xx = emit.ox; xx < emit.ox + emit.rx; xx++)
          m.used[yy * bx + xx] = 1;

      memset(&areas[count], 0, sizeof(lo_area));
      areas[count].ox = (uint32_t)emit.ox; areas[count].oy = (uint32_t)emit.oy;
      areas[count].rx = (uint32_t)emit.rx; areas[count].ry = (uint32_t)emit.ry;
      areas[count].stage = stage;
      count++;
      st[1 + stage]++;

      if (!rescan)
        x += r.rx; /* continue behind the rectangle; after a centre-third hit the same seed is examined again */
    }
  }

  for (size_t y = 0; y < by; y++)
  {
    for (size_t x = 0; x < bx; x++)
    {
      if (m.used[y * bx + x])
        continue;

      memset(&areas[count], 0, sizeof(lo_area));
      areas[count].ox = (uint32_t)x; areas[count].oy = (uint32_t)y; areas[count].rx = 1; areas[count].ry = 1;
      areas[count].stage = 2;
      count++;
    }
  }

  st[5] = g_predicates - pred0;
  st[6] = g_full_predicates - full0;

  if (stats)
    memcpy(stats, st, sizeof(st));

  free(m.used);
  return count;
}

/* ------------------------------------------------------------------------------------------------
 * per-area encode + plane writer (limg.cpp:1498-1772), Q12
 * --------------------------------------------------------------------------------------------- */

typedef struct
{
  uint32_t *px;
  uint8_t *fa, *fb, *fc;
  size_t cap;
} scratch;

static void scratch_reserve(scratch *s, size_t n)
{
  if (n <= s->cap)
    return;

  s->px = (uint32_t *)realloc(s->px, n * sizeof(uint32_t));
  s->fa = (uint8_t *)realloc(s->fa, n);
  s->fb = (uint8_t *)realloc(s->fb, n);
  s->fc = (uint8_t *)realloc(s->fc, n);
  s->cap = n;
}

static void scratch_free(scratch *s) { free(s->px); free(s->fa); free(s->fb); free(s->fc); }

static inline uint32_t clamp_u8(int32_t v) { return (uint32_t)(v < 0 ? 0 : (v > 0xFF ? 0xFF : v)); }

/* Encodes the pixel rectangle (x0,y0,w,h) with decomposition d; advances the dither chain; writes planes. */
static void encode_area(const uint32_t *img, size_t sizeX, int ch, uint32_t errorFactor, int fast, int ditherMode, const lo_decomp *d,
                        size_t x0, size_t y0, size_t w, size_t h, uint32_t blockIndex, uint64_t *dither, scratch *s, lo_planes *p, uint8_t shiftOut[3])
{
  const size_t n = w * h;
  scratch_reserve(s, n);
  gather(img, sizeX, x0, y0, w, h, s->px);
  lo_project(ch, d, s->px, n, s->fa, s->fb, s->fc);

  uint8_t shift[3];
  lo_search(ch, errorFactor, fast, d, s->px, n, s->fa, s->fb, s->fc, shift);

  if (shift[0] || shift[1] || shift[2])
  {
    uint8_t *f[3] = { s->fa, s->fb, s->fc };

    for (int i = 0; i < 3; i++)
      if (shift[i] && shift[i] != 8)
        *dither = lo_dither(ditherMode, shift[i], n, *dither, f[i]); /* Q7/Q13: a dropped factor keeps its raw byte */
  }

  for (int i = 0; i < 3; i++)
    shiftOut[i] = shift[i];

  static const uint8_t pattern[9] = { 0, 0x22, 0x44, 0x66, 0x88, 0xAA, 0xCC, 0xEE, 0xFF };
  const uint32_t shiftWord = 0xFF000000u | ((uint32_t)pattern[shift[0]] << 16) | ((uint32_t)pattern[shift[1]] << 8) | pattern[shift[2]];
  uint32_t col[6] = { 0, 0, 0, 0, 0, 0 };

  for (int i = 0; i < ch; i++)
  {
    col[0] |= clamp_u8(d->dirA_min[i]) << (8 * i);
    col[1] |= clamp_u8(d->dirA_max[i]) << (8 * i);
    col[2] |= clamp_u8(d->dirB_offset[i] + 0x80) << (8 * i);
    col[3] |= clamp_u8(d->dirB_mag[i] + 0x80) << (8 * i);
    col[4] |= clamp_u8(d->dirC_offset[i] + 0x80) << (8 * i);
    col[5] |= clamp_u8(d->dirC_mag[i] + 0x80) << (8 * i);
  }

  if (ch == 3)
    for (int i = 0; i < 6; i++)
      col[i] |= 0xFF000000u;

  const size_t headerBits = (size_t)ch * 9 * 2 + (size_t)ch * 8 + 32; /* 110 / 136 (limg.cpp:1630) */
  const size_t bits = headerBits + n * (size_t)((8 - shift[0]) + (8 - shift[1]) + (8 - shift[2]));
  const uint8_t bpp = (uint8_t)((bits + n / 2) / n);

  for (size_t y = 0; y < h; y++)
  {
    const size_t row = (y0 + y) * sizeX + x0;

    for (size_t x = 0; x < w; x++)
    {
      const size_t k = y * w + x;
      if (p->pFactorsA) p->pFactorsA[row + x] = (uint8_t)(s->fa[k] << shift[0]);
      if (p->pFactorsB) p->pFactorsB[row + x] = (uint8_t)(s->fb[k] << shift[1]);
      if (p->pFactorsC) p->pFactorsC[row + x] = (uint8_t)(s->fc[k] << shift[2]);
      if (p->pBitsPerPixel) p->pBitsPerPixel[row + x] = bpp;
      if (p->pShiftABCX) p->pShiftABCX[row + x] = shiftWord;
      if (p->pColAMin) p->pColAMin[row + x] = col[0];
      if (p->pColAMax) p->pColAMax[row + x] = col[1];
      if (p->pColBMin) p->pColBMin[row + x] = col[2];
      if (p->pColBMax) p->pColBMax[row + x] = col[3];
      if (p->pColCMin) p->pColCMin[row + x] = col[4];
      if (p->pColCMax) p->pColCMax[row + x] = col[5];
      if (p->pBlockIndex) p->pBlockIndex[row + x] = 0xFF000000u | blockIndex;
    }
  }

  if (p->pDecoded)
    lo_decode(ch, p->pDecoded + y0 * sizeX + x0, sizeX, w, h, s->fa, s->fb, s->fc, d, shift);
}

size_t lo_blocked_encode3d(const uint32_t *img, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, int fast, int ditherMode,
                           lo_planes *planes, lo_area *areasOut)
{
  const int ch = hasAlpha ? 4 : 3;
  const size_t bx = (sizeX + LO_BLOCK - 1) / LO_BLOCK, by = (sizeY + LO_BLOCK - 1) / LO_BLOCK;
  lo_decomp *table = (lo_decomp *)malloc((bx * by > 0 ? bx * by : 1) * sizeof(lo_decomp));
  lo_area *areas = areasOut ? areasOut : (lo_area *)malloc((bx * by > 0 ? bx * by : 1) * sizeof(lo_area));

  lo_pass1(img, sizeX, sizeY, ch, table);
  const size_t count = lo_merge(table, bx, by, ch, areas, NULL);

  scratch s = { 0 };
  uint64_t dither = 0xCA7F00D15BADF00DULL; /* limg_internal.h:711 */

  for (size_t k = 0; k < count; k++)
  {
    lo_area *a = &areas[k];
    size_t w = a->rx * LO_BLOCK, h = a->ry * LO_BLOCK;

    if (a->ox + a->rx == bx && (sizeX % LO_BLOCK)) w = w - LO_BLOCK + sizeX % LO_BLOCK;
    if (a->oy + a->ry == by && (sizeY % LO_BLOCK)) h = h - LO_BLOCK + sizeY % LO_BLOCK;

    a->px_x = a->ox * LO_BLOCK; a->px_y = a->oy * LO_BLOCK; a->px_w = (uint32_t)w; a->px_h = (uint32_t)h;

    if (a->stage == 2)
    {
      a->decomp = table[a->oy * bx + a->ox]; /* leftovers keep the pass-1 fit (limg.cpp:1875) */
    }
    else
    {
      scratch_reserve(&s, w * h);
      gather(img, sizeX, a->px_x, a->px_y, w, h, s.px);
      lo_fit(s.px, w * h, ch, &a->decomp);
    }

    a->ditherBefore = dither;
    encode_area(img, sizeX, ch, errorFactor, fast, ditherMode, &a->decomp, a->px_x, a->px_y, w, h, (uint32_t)(k + 1), &dither, &s, planes, a->shift);
    a->ditherAfter = dither;
  }

  scratch_free(&s);
  free(table);

  if (!areasOut)
    free(areas);

  return count;
}

void lo_encode3d(const uint32_t *img, size_t sizeX, size_t sizeY, int hasAlpha, uint32_t errorFactor, int fast, int ditherMode, size_t poolThreads, lo_planes *planes)
{
  const int ch = hasAlpha ? 4 : 3;
  lo_planes p = *planes;
  p.pBitsPerPixel = NULL; /* limg_encode3d_info has no bpp / block index planes */
  p.pBlockIndex = NULL;

  /* y-bands (limg.cpp:2108-2137) */
  size_t bandCount = 1, bandRows = sizeY;

  if (poolThreads)
  {
    bandCount = poolThreads * 4;
    bandRows = ((sizeY / LO_BLOCK) / bandCount) * LO_BLOCK;

    if (bandRows == 0)
    {
      bandCount = poolThreads;
      bandRows = ((sizeY / LO_BLOCK) / bandCount) * LO_BLOCK;
    }
  }

  scratch s = { 0 };
  size_t yStart = 0;

  for (size_t band = 0; band < bandCount; band++)
  {
    const size_t yEnd = band + 1 == bandCount ? sizeY : yStart + bandRows;
    uint64_t dither = 0xCA7F00D15BADF00DULL; /* limg.cpp:1893 */

    for (size_t y = yStart; y < yEnd; y += LO_BLOCK)
    {
      for (size_t x = 0; x < sizeX; x += LO_BLOCK)
      {
        const size_t w = sizeX - x < LO_BLOCK ? sizeX - x : LO_BLOCK;
        const size_t h = sizeY - y < LO_BLOCK ? sizeY - y : LO_BLOCK;
        lo_decomp d;
        uint8_t shift[3];

        scratch_reserve(&s, w * h);
        gather(img, sizeX, x, y, w, h, s.px);
        lo_fit(s.px, w * h, ch, &d);
        encode_area(img, sizeX, ch, errorFactor, fast, ditherMode, &d, x, y, w, h, 0, &dither, &s, &p, shift);
      }
    }

    yStart = yEnd;
  }

  scratch_free(&s);
}

void lo_decode_areas(int ch, const lo_area *areas, size_t count, const uint8_t *fa, const uint8_t *fb, const uint8_t *fc, uint32_t *out, size_t sizeX)
{
  size_t off = 0;

  for (size_t k = 0; k < count; k++)
  {
    const lo_area *a = &areas[k];
    lo_decode(ch, out + (size_t)a->px_y * sizeX + a->px_x, sizeX, a->px_w, a->px_h, fa + off, fb + off, fc + off, &a->decomp, a->shift);
    off += (size_t)a->px_w * a->px_h;
  }
}

/* ------------------------------------------------------------------------------------------------
 * limg_compare (limg.cpp:2455-2491, limg_internal.h:376-410)
 * --------------------------------------------------------------------------------------------- */

static uint64_t color_error(int ch, uint32_t a, uint32_t b)
{
  int64_t d[4];

  for (int i = 0; i < 4; i++)
    d[i] = (int64_t)((a >> (8 * i)) & 0xFF) - (int64_t)((b >> (8 * i)) & 0xFF);

  const int64_t rr = d[0] * d[0];
  const int low = rr < 0x4000;
  uint64_t e = (uint64_t)rr * (low ? 2 : 3) + (uint64_t)(d[1] * d[1]) * 4 + (uint64_t)(d[2] * d[2]) * (low ? 3 : 2);

  if (ch == 4)
    e += (uint64_t)(d[3] * d[3]) * 3;

  return e;
}

double lo_compare(const uint32_t *a, const uint32_t *b, size_t sizeX, size_t sizeY, int hasAlpha, double *mse, double *maxErr)
{
  const int ch = hasAlpha ? 4 : 3;
  const uint64_t maxError = color_error(ch, 0u, 0xFFFFFFFFu);
  uint64_t error = 0;

  for (size_t i = 0; i < sizeX * sizeY; i++)
    error += color_error(ch, a[i], b[i]);

  const double m = (double)error / (double)(sizeX * sizeY);

  if (mse) *mse = m;
  if (maxErr) *maxErr = (double)maxError;

  return 10.0 * log10((double)maxError / m);
}

/* debugging / design counters */
uint64_t lo_counter(int which)
{
  switch (which)
  {
  case 0: return g_predicates;
  case 1: return g_full_predicates;
  case 2: return g_trials;
  default: return 0;
  }
}

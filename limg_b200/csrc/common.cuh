// limg_b200/csrc/common.cuh -- shared device-side definitions for the B200 (sm_100a) limg hot path.
//
// Arithmetic contract: every FP32 operation of the fit, the merge predicate and the
// projection is one IEEE binary32 operation in the reference's order. The file is compiled with
// -fmad=false and the kernels additionally use the _rn intrinsics, which are never contracted.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#include "../../include/limgcu.h"

#define LIMG_BLOCK 8

namespace limg
{

// ---------------------------------------------------------------------------------------------
// x86 SSE semantics on the device
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float frcp1(float x) { return __fdiv_rn(1.0f, x); } // DIVPS(1, x), IEEE

// MINPS / MAXPS: second operand unless the strict comparison holds.
__device__ __forceinline__ float sse_min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float sse_max(float a, float b) { return a > b ? a : b; }

// CVTPS2DQ (round to nearest even); NaN and out-of-range give 0x80000000.
__device__ __forceinline__ int32_t sse_cvtps(float x)
{
  if (!(x >= -2147483648.0f && x < 2147483648.0f))
    return (int32_t)0x80000000;
  return __float2int_rn(x);
}

// DPPS: (p0 + p1) + (p2 + p3), products rounded individually; a masked lane contributes +0.
__device__ __forceinline__ float dp3(float a0, float a1, float a2, float b0, float b1, float b2)
{
  return fadd(fadd(fmul(a0, b0), fmul(a1, b1)), fadd(fmul(a2, b2), 0.0f));
}

__device__ __forceinline__ float dp4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3)
{
  return fadd(fadd(fmul(a0, b0), fmul(a1, b1)), fadd(fmul(a2, b2), fmul(a3, b3)));
}

struct f4
{
  float x, y, z, w;
};

template <int CH>
__device__ __forceinline__ float dpn(const f4 &a, const f4 &b)
{
  if (CH == 3)
    return dp3(a.x, a.y, a.z, b.x, b.y, b.z);
  else
    return dp4(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
}

// RSQRTPS as a table function (limg_b200/csrc/rsqrt_lut.h); lut points at 2048 uint16 entries (shared or global).
__device__ __forceinline__ float sse_rsqrt(float x, const uint16_t *__restrict__ lut)
{
  const uint32_t b = __float_as_uint(x);
  const uint32_t e = (b >> 23) & 0xFF;
  const uint32_t m = b & 0x7FFFFF;

  if (e == 0xFF)
    return m ? __uint_as_float(b | 0x00400000u) : ((b >> 31) ? __uint_as_float(0xFFC00000u) : 0.0f);
  if (e == 0)
    return __uint_as_float((b & 0x80000000u) | 0x7F800000u);
  if (b >> 31)
    return __uint_as_float(0xFFC00000u);

  const int E = (int)e - 127;
  const int p = E & 1;
  const int k = (E - p) >> 1;
  const uint32_t mant12 = lut[p * 1024 + (m >> 13)];
  return __uint_as_float(((uint32_t)(126 - k) << 23) | (mant12 << 11));
}

__device__ __forceinline__ f4 px_to_f4(uint32_t px)
{
  f4 r;
  r.x = (float)(px & 0xFF);
  r.y = (float)((px >> 8) & 0xFF);
  r.z = (float)((px >> 16) & 0xFF);
  r.w = (float)(px >> 24);
  return r;
}

// ---------------------------------------------------------------------------------------------
// bulk-copy (TMA) plumbing: mbarriers, 1-D bulk copies, 2-D tensor-map tile loads
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "LIMG_WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra LIMG_DONE_%=;\n"
               "bra LIMG_WAIT_%=;\n"
               "LIMG_DONE_%=:\n"
               "}" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one tile of a 2-D tensor map (cuTensorMapEncodeTiled on the host): innermost coordinate x, then y; elements outside the tensor arrive as zeros
__device__ __forceinline__ void tensor_g2s_2d(uint32_t dst, const void *tensorMap, int x, int y, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(dst), "l"(tensorMap), "r"(x), "r"(y), "r"(bar) : "memory");
}

__device__ __forceinline__ void fence_mbarrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// integer reconstruction (limg_decode.h:39-236, limg_bit_crush_simd.h:315-810)
// ---------------------------------------------------------------------------------------------

// (1 << s) + decode_bias[s], decode_bias = {0,0,0,0,1,4,21,127,0} (Q8)
__device__ __forceinline__ int32_t decode_mul(int s)
{
  // bias bytes 0,0,0,0,1,4,21,127 packed little-endian; shift 8 has bias 0
  const unsigned long long packed = 0x7F15040100000000ull;
  const int32_t bias = s < 8 ? (int32_t)((packed >> (8 * s)) & 0xFF) : 0;
  return (1 << s) + bias;
}

struct Recon
{
  // k = mul[s] * normal (the per-factor multiplier folded into the normal), m = (min << 8) + 128
  int32_t kA[4], kB[4], kC[4];
  int32_t mA[4], mB[4], mC[4];
};

// rgbAlphaMin: 0xFFFF for the decoder (forces alpha 0xFF on RGB, limg_decode.h:95-97), 0 for the trial.
template <int CH>
__device__ __forceinline__ void init_recon(const limgcu_decomp &d, int sA, int sB, int sC, int32_t rgbAlphaMin, Recon &r)
{
  const int32_t mulA = decode_mul(sA), mulB = decode_mul(sB), mulC = decode_mul(sC);

#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    int32_t nA = (int32_t)d.dirA_max[i] - d.dirA_min[i];
    int32_t nB = (int32_t)d.dirB_mag[i] - d.dirB_offset[i];
    int32_t nC = (int32_t)d.dirC_mag[i] - d.dirC_offset[i];
    int32_t minA = d.dirA_min[i], minB = d.dirB_offset[i], minC = d.dirC_offset[i];

    if (CH == 3 && i == 3)
    {
      nA = nB = nC = 0;
      minA = minB = minC = rgbAlphaMin;
    }

    // Q7: a dropped factor only clears the first three channels.
    if (i < 3)
    {
      if (sA > 7) nA = 0;
      if (sB > 7) { nB = 0; minB = 0; }
      if (sC > 7) { nC = 0; minC = 0; }
    }

    r.kA[i] = (int32_t)((uint32_t)mulA * (uint32_t)nA);
    r.kB[i] = (int32_t)((uint32_t)mulB * (uint32_t)nB);
    r.kC[i] = (int32_t)((uint32_t)mulC * (uint32_t)nC);
    r.mA[i] = (int32_t)((uint32_t)minA << 8) + 128;
    r.mB[i] = (int32_t)((uint32_t)minB << 8) + 128;
    r.mC[i] = (int32_t)((uint32_t)minC << 8) + 128;
  }
}

__device__ __forceinline__ int32_t clamp255(int32_t v) { return min(max(v, 0), 255); }

// one channel of one pixel; eA/eB/eC are the right-aligned codes.
__device__ __forceinline__ int32_t recon_channel(const Recon &r, int i, int32_t eA, int32_t eB, int32_t eC)
{
  // 32-bit wrapping products like PMULLD; arithmetic shifts like PSRAD
  const int32_t tA = (int32_t)((uint32_t)eA * (uint32_t)r.kA[i] + (uint32_t)r.mA[i]) >> 8;
  const int32_t tB = (int32_t)((uint32_t)eB * (uint32_t)r.kB[i] + (uint32_t)r.mB[i]) >> 8;
  const int32_t tC = (int32_t)((uint32_t)eC * (uint32_t)r.kC[i] + (uint32_t)r.mC[i]) >> 8;
  const int32_t v = (int32_t)((uint32_t)tA + (uint32_t)tB + (uint32_t)tC);
  return clamp255(v);
}

// perceptual error of the trial (alpha never counted: Q6)
__device__ __forceinline__ int32_t trial_error(uint32_t px, int32_t cr, int32_t cg, int32_t cb)
{
  const int32_t dr = (int32_t)(px & 0xFF) - cr;
  const int32_t dg = (int32_t)((px >> 8) & 0xFF) - cg;
  const int32_t db = (int32_t)((px >> 16) & 0xFF) - cb;
  const int32_t rr = dr * dr;
  const bool lowRed = rr < 0x4000;
  return rr * (lowRed ? 2 : 3) + dg * dg * 4 + db * db * (lowRed ? 3 : 2);
}

// ---------------------------------------------------------------------------------------------
// projection state (limg_internal.h:426-452) and per-pixel projection (limg_factorization.h:101-197)
// ---------------------------------------------------------------------------------------------

struct Proj
{
  f4 minA, offB, offC;
  f4 nA, nB, nC;
  float invA, invB, invC;
};

template <int CH>
__device__ __forceinline__ void init_proj(const limgcu_decomp &d, Proj &p)
{
  float nA[4], nB[4], nC[4];
  bool zA = false, zB = false, zC = false;

#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    nA[i] = (i < CH) ? (float)((int)d.dirA_max[i] - (int)d.dirA_min[i]) : 0.0f;
    nB[i] = (i < CH) ? (float)((int)d.dirB_mag[i] - (int)d.dirB_offset[i]) : 0.0f;
    nC[i] = (i < CH) ? (float)((int)d.dirC_mag[i] - (int)d.dirC_offset[i]) : 0.0f;
    zA |= nA[i] != 0.0f;
    zB |= nB[i] != 0.0f;
    zC |= nC[i] != 0.0f;
  }

  // limg_dot: sequential ((p0 + p1) + p2) + p3 starting from 0
  float sA = 0.0f, sB = 0.0f, sC = 0.0f;

#pragma unroll
  for (int i = 0; i < CH; i++)
  {
    sA = fadd(sA, fmul(nA[i], nA[i]));
    sB = fadd(sB, fmul(nB[i], nB[i]));
    sC = fadd(sC, fmul(nC[i], nC[i]));
  }

  p.invA = zA ? frcp1(sA) : 0.0f;
  p.invB = zB ? frcp1(sB) : 0.0f;
  p.invC = zC ? frcp1(sC) : 0.0f;
  p.nA = { nA[0], nA[1], nA[2], nA[3] };
  p.nB = { nB[0], nB[1], nB[2], nB[3] };
  p.nC = { nC[0], nC[1], nC[2], nC[3] };
  p.minA = { (float)d.dirA_min[0], (float)d.dirA_min[1], (float)d.dirA_min[2], (float)d.dirA_min[3] };
  p.offB = { (float)d.dirB_offset[0], (float)d.dirB_offset[1], (float)d.dirB_offset[2], (float)d.dirB_offset[3] };
  p.offC = { (float)d.dirC_offset[0], (float)d.dirC_offset[1], (float)d.dirC_offset[2], (float)d.dirC_offset[3] };
}

__device__ __forceinline__ uint32_t factor_to_u8(float f)
{
  int32_t v = sse_cvtps(fmul(255.0f, f));
  v = min(v, 0xFF);
  v = max(v, 0);
  return (uint32_t)v;
}

// returns fa | fb << 8 | fc << 16
template <int CH>
__device__ __forceinline__ uint32_t project_px(const Proj &p, uint32_t pixel)
{
  const f4 c = px_to_f4(pixel);
  f4 t, est;

  t = { fsub(c.x, p.minA.x), fsub(c.y, p.minA.y), fsub(c.z, p.minA.z), fsub(c.w, p.minA.w) };
  const float facA = fmul(dpn<CH>(t, p.nA), p.invA);

  est = { fadd(p.minA.x, fmul(p.nA.x, facA)), fadd(p.minA.y, fmul(p.nA.y, facA)), fadd(p.minA.z, fmul(p.nA.z, facA)), fadd(p.minA.w, fmul(p.nA.w, facA)) };
  t = { fsub(fsub(c.x, est.x), p.offB.x), fsub(fsub(c.y, est.y), p.offB.y), fsub(fsub(c.z, est.z), p.offB.z), fsub(fsub(c.w, est.w), p.offB.w) };
  const float facB = fmul(dpn<CH>(t, p.nB), p.invB);

  est = { fadd(est.x, fmul(p.nB.x, facB)), fadd(est.y, fmul(p.nB.y, facB)), fadd(est.z, fmul(p.nB.z, facB)), fadd(est.w, fmul(p.nB.w, facB)) };
  t = { fsub(fsub(c.x, est.x), p.offC.x), fsub(fsub(c.y, est.y), p.offC.y), fsub(fsub(c.z, est.z), p.offC.z), fsub(fsub(c.w, est.w), p.offC.w) };
  const float facC = fmul(dpn<CH>(t, p.nC), p.invC);

  return factor_to_u8(facA) | (factor_to_u8(facB) << 8) | (factor_to_u8(facC) << 16);
}

// ---------------------------------------------------------------------------------------------
// dither: PCG-style LCG with O(log n) jump-ahead (limg.cpp:798-822)
// ---------------------------------------------------------------------------------------------

#define LIMG_LCG_MUL 6364136223846793005ULL
#define LIMG_DITHER_SEED 0xCA7F00D15BADF00DULL

struct LcgJumpTable
{
  uint64_t mul[64]; // a^(2^j)
  uint64_t add[64]; // c * (a^(2^j) - 1) / (a - 1), c = 1
};

__device__ __forceinline__ uint64_t lcg_jump(uint64_t h, uint64_t steps, const LcgJumpTable &t)
{
  int j = 0;

  while (steps)
  {
    if (steps & 1)
      h = h * t.mul[j] + t.add[j];

    steps >>= 1;
    j++;
  }

  return h;
}

__device__ __forceinline__ uint32_t pcg_output(uint64_t h)
{
  const uint32_t xs = (uint32_t)(((h >> 18) ^ h) >> 27);
  const uint32_t rot = (uint32_t)(h >> 59);
  return __funnelshift_r(xs, xs, rot);
}

// noise + clamp + shift-down of one factor byte (0 < shift < 8)
__device__ __forceinline__ uint32_t dither_one(uint32_t f, uint32_t rnd, int shift)
{
  const int32_t mask = (1 << shift) - 1;
  const int32_t offset = 1 << (shift - 1);
  const int32_t v = clamp255((int32_t)f + (((int32_t)rnd & mask) - offset));
  return (uint32_t)(v >> shift);
}

// ---------------------------------------------------------------------------------------------
// thresholds (limg.cpp:2344-2345, 2365-2366)
// ---------------------------------------------------------------------------------------------

struct CrushParams
{
  uint64_t maxPixelError;
  uint64_t maxBlockError;
  int crushBits; // errorFactor != 0
  int fast;
};

__host__ __device__ inline CrushParams make_crush_params(uint32_t errorFactor, int fast)
{
  CrushParams p;
  p.maxPixelError = (uint64_t)0x6 * (errorFactor / 2) * 7;
  p.maxBlockError = (uint64_t)0x4 * (errorFactor / 2) * 7;
  p.crushBits = errorFactor != 0;
  p.fast = fast;
  return p;
}

} // namespace limg

// kernels_decode.cuh -- streaming reconstruction for images whose width is a multiple of 8 (every config of BASELINE.json).
//
// Reference: limg_decode_block_from_factors_3d_{3,4}_sse41 (limg_decode.h:39-135, 140-236): per pixel and channel
//   col = clamp(((decA*nA + (minA<<8) + 128) >> 8) + ((decB*nB + ...) >> 8) + ((decC*nC + ...) >> 8), 0, 255)
// with 32-bit wrapping products (PMULLD), arithmetic shifts (PSRAD) and dec = enc * ((1 << s) + bias[s]).
//
// The generic k_decode (kernels_stream.cuh) is bound by the integer ALU pipe, not by HBM: ~27 alu-pipe instructions per pixel
// (9 shifts, 6 min/max, adds, byte extraction, packing) next to 9 IMADs. This kernel moves work off the alu pipe:
//   * `>> 8` and the sum of the three factors are one chain of IMAD.HI:  hi32(t * 2^24) + acc == (t >> 8) + acc for every
//     32-bit t (arithmetic shift, wrapping add), i.e. bit-identical to PSRAD + PADDD. With the literal 2^24 ptxas emits
//     LEA.HI.SX32 (shift + add in one alu-pipe instruction). A multiplier ptxas cannot see through keeps a real IMAD.HI on the fma
//     pipe; measured on B200 that is slower (IMAD.HI is not full rate: 8K RGB 59 us -> 66 / 80 us with two / three channels on it);
//   * clamp + pack are two I2IP (cvt.pack.sat.u8.s32): sat_u8(a) << 8 | sat_u8(b) | c << 16 -- the same saturation as
//     PACKSSDW/PACKUSWB at limg_decode.h:118-121;
//   * codes are extracted with one PRMT each.
#pragma once

#include "common.cuh"

namespace limg
{

__device__ __forceinline__ uint32_t pack_sat_u8(int32_t hi, int32_t lo, uint32_t upper)
{
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
  return d;
}

__device__ __forceinline__ int32_t mad_hi(int32_t a, int32_t b, int32_t c)
{
  int32_t d;
  asm("mad.hi.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// sum over the three factors of (e * k + m) >> 8 for one channel
// HI: the multiplier 2^24 comes from a register ptxas cannot see through, so the chain stays IMAD.HI (fma pipe); otherwise ptxas turns
// the constant form into LEA.HI.SX32 (alu pipe, shift and add in one instruction).
template <bool HI>
__device__ __forceinline__ int32_t recon_sum(int32_t eA, int32_t eB, int32_t eC, int32_t kA, int32_t kB, int32_t kC, int32_t mA, int32_t mB, int32_t mC, int32_t hiMul)
{
  const int32_t tA = (int32_t)((uint32_t)eA * (uint32_t)kA + (uint32_t)mA);
  const int32_t tB = (int32_t)((uint32_t)eB * (uint32_t)kB + (uint32_t)mB);
  const int32_t tC = (int32_t)((uint32_t)eC * (uint32_t)kC + (uint32_t)mC);

  if (HI)
    return mad_hi(tA, hiMul, mad_hi(tB, hiMul, mad_hi(tC, hiMul, 0)));
  else
    return mad_hi(tA, 1 << 24, mad_hi(tB, 1 << 24, mad_hi(tC, 1 << 24, 0)));
}

// (1 << s) + decode_bias[s] for s = 0..7 (limg_bit_crush.h decode bias table, Q8); entry 8 is 0: a dropped factor contributes
// nothing to the first three channels (Q7). The alpha channel of RGBA keeps its normal with the multiplier 256 (1 << 8, bias 0).
__constant__ int32_t c_decode_mul_rgb[16] = {1, 2, 4, 8, 17, 36, 85, 255, 0, 0, 0, 0, 0, 0, 0, 0};

__device__ __forceinline__ int32_t lo16(uint32_t v) { return (int32_t)(int16_t)(v & 0xFFFF); }
__device__ __forceinline__ int32_t hi16(uint32_t v) { return (int32_t)v >> 16; }

// reconstruction constants of one channel: k = mul * (max - min), m = (min << 8) + 128 (init_recon in common.cuh, specialised)
struct ReconCh
{
  int32_t kA, kB, kC, mA, mB, mC;
};

__device__ __forceinline__ ReconCh recon_channel_setup(int32_t aMin, int32_t aMax, int32_t bOff, int32_t bMag, int32_t cOff, int32_t cMag, int32_t mulA, int32_t mulB, int32_t mulC,
                                                       bool dropB, bool dropC)
{
  ReconCh r;
  r.kA = mulA * (aMax - aMin);
  r.kB = mulB * (bMag - bOff);
  r.kC = mulC * (cMag - cOff);
  r.mA = aMin * 256 + 128;
  r.mB = dropB ? 128 : bOff * 256 + 128;
  r.mC = dropC ? 128 : cOff * 256 + 128;
  return r;
}

__device__ __forceinline__ int32_t recon_sum_ch(const ReconCh &r, int32_t eA, int32_t eB, int32_t eC)
{
  return recon_sum<false>(eA, eB, eC, r.kA, r.kB, r.kC, r.mA, r.mB, r.mC, 0);
}

// One thread reconstructs ROWS rows of one 8x8 block (8 x ROWS pixels); consecutive threads cover consecutive blocks of the same rows,
// so a warp reads 256 contiguous bytes of every code row and writes 1 KB of every output row. The per-thread set-up (area record ->
// 18 / 24 constants) is ~90 instructions: the record's six int16x4 vectors are loaded as 8-byte words, the multipliers come from
// constant memory, the thread index is split with 32-bit arithmetic.
template <int CH, int ROWS, int CTA, bool CS>
__global__ void __launch_bounds__(CTA, 1024 / CTA) k_decode_tile(const limgcu_area *__restrict__ areas, const uint32_t *__restrict__ blockToArea, const uint8_t *__restrict__ codesA,
                                                     const uint8_t *__restrict__ codesB, const uint8_t *__restrict__ codesC, int W, int H, int BX, uint32_t threads,
                                                     uint32_t *__restrict__ dst)
{
  constexpr int PARTS = 8 / ROWS;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;

  if (t >= threads)
    return;

  const uint32_t rest = t / (uint32_t)BX;
  const int bx = (int)(t - rest * (uint32_t)BX);
  const int by = (int)(rest / PARTS), part = (int)(rest % PARTS);
  const int y0 = by * 8 + part * ROWS;

  if (y0 >= H)
    return;

  const int nrows = min(ROWS, H - y0);
  const size_t base = (size_t)y0 * W + bx * 8;

  // codes first: the loads do not depend on the area record
  uint2 va[ROWS], vb[ROWS], vc[ROWS];

#pragma unroll
  for (int j = 0; j < ROWS; j++)
  {
    if (j < nrows)
    {
      const size_t off = base + (size_t)j * W;
      va[j] = CS ? __ldcs(reinterpret_cast<const uint2 *>(codesA + off)) : __ldg(reinterpret_cast<const uint2 *>(codesA + off));
      vb[j] = CS ? __ldcs(reinterpret_cast<const uint2 *>(codesB + off)) : __ldg(reinterpret_cast<const uint2 *>(codesB + off));
      vc[j] = CS ? __ldcs(reinterpret_cast<const uint2 *>(codesC + off)) : __ldg(reinterpret_cast<const uint2 *>(codesC + off));
    }
  }

  const char *ar = reinterpret_cast<const char *>(areas + __ldg(blockToArea + (size_t)by * BX + bx));
  const uint32_t shifts = __ldg(reinterpret_cast<const uint32_t *>(ar + offsetof(limgcu_area, shift)));
  const uint2 *dq = reinterpret_cast<const uint2 *>(ar + offsetof(limgcu_area, decomp) + offsetof(limgcu_decomp, dirA_min));
  const uint2 aMin = __ldg(dq + 0), aMax = __ldg(dq + 1), bOff = __ldg(dq + 2), bMag = __ldg(dq + 3), cOff = __ldg(dq + 4), cMag = __ldg(dq + 5);

  const int sA = min(shifts & 0xFF, 8u), sB = min((shifts >> 8) & 0xFF, 8u), sC = min((shifts >> 16) & 0xFF, 8u);
  const int32_t mulA = c_decode_mul_rgb[sA], mulB = c_decode_mul_rgb[sB], mulC = c_decode_mul_rgb[sC];
  const bool dropB = sB > 7, dropC = sC > 7;

  const ReconCh r0 = recon_channel_setup(lo16(aMin.x), lo16(aMax.x), lo16(bOff.x), lo16(bMag.x), lo16(cOff.x), lo16(cMag.x), mulA, mulB, mulC, dropB, dropC);
  const ReconCh r1 = recon_channel_setup(hi16(aMin.x), hi16(aMax.x), hi16(bOff.x), hi16(bMag.x), hi16(cOff.x), hi16(cMag.x), mulA, mulB, mulC, dropB, dropC);
  const ReconCh r2 = recon_channel_setup(lo16(aMin.y), lo16(aMax.y), lo16(bOff.y), lo16(bMag.y), lo16(cOff.y), lo16(cMag.y), mulA, mulB, mulC, dropB, dropC);
  ReconCh r3 = {};

  if (CH == 4) // Q7: the alpha channel ignores dropped factors (multiplier 1 << 8 with bias 0)
    r3 = recon_channel_setup(hi16(aMin.y), hi16(aMax.y), hi16(bOff.y), hi16(bMag.y), hi16(cOff.y), hi16(cMag.y), sA > 7 ? 256 : mulA, dropB ? 256 : mulB, dropC ? 256 : mulC, false, false);

#pragma unroll
  for (int j = 0; j < ROWS; j++)
  {
    if (j >= nrows)
      break;

    uint32_t out[8];

#pragma unroll
    for (int i = 0; i < 8; i++)
    {
      const uint32_t wa = i < 4 ? va[j].x : va[j].y, wb = i < 4 ? vb[j].x : vb[j].y, wc = i < 4 ? vc[j].x : vc[j].y;
      const int32_t eA = (int32_t)__byte_perm(wa, 0, 0x4440 + (i & 3));
      const int32_t eB = (int32_t)__byte_perm(wb, 0, 0x4440 + (i & 3));
      const int32_t eC = (int32_t)__byte_perm(wc, 0, 0x4440 + (i & 3));
      const int32_t c0 = recon_sum_ch(r0, eA, eB, eC), c1 = recon_sum_ch(r1, eA, eB, eC), c2 = recon_sum_ch(r2, eA, eB, eC);
      const int32_t c3 = CH == 4 ? recon_sum_ch(r3, eA, eB, eC) : 255; // RGB: the 0xFFFF "min" trick of limg_decode.h:95-97 saturates to 0xFF

      out[i] = pack_sat_u8(c1, c0, pack_sat_u8(c3, c2, 0));
    }

    uint4 *p = reinterpret_cast<uint4 *>(dst + base + (size_t)j * W);
    if (CS)
    {
      __stcs(p, make_uint4(out[0], out[1], out[2], out[3]));
      __stcs(p + 1, make_uint4(out[4], out[5], out[6], out[7]));
    }
    else
    {
      p[0] = make_uint4(out[0], out[1], out[2], out[3]);
      p[1] = make_uint4(out[4], out[5], out[6], out[7]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_decode_stream: the same arithmetic behind a bulk-copy (TMA) pipeline
// ---------------------------------------------------------------------------------------------
//
// k_decode_tile keeps its twelve code words in registers while they are in flight, and all warps of an SM move through "load, compute,
// store" in step: ncu shows 41 % of the cycles without an eligible warp although the alu pipe is the busiest unit. Here every warp runs
// its own two-stage pipeline: tiles are 32 blocks x 8 rows of the three code planes (24 rows of 256 B = 6 KB); lanes 0..23 issue one
// cp.async.bulk each (global -> shared, completion counted in bytes on the stage's mbarrier) for the warp's NEXT tile before the warp
// waits for the current one, so the copies of tile n+1 are in flight during the ~1500 instructions of tile n and no register holds
// data that has not arrived. The area record of the next tile and the area index of the tile after that are prefetched the same way
// (in registers). One thread reconstructs a whole 8x8 block, so the 90-instruction set-up is paid once per 64 pixels.
// Needs sizeX % 16 == 0 and 16-byte aligned planes (bulk copies move multiples of 16 bytes).

struct AreaRaw // the 52 bytes of an area record the reconstruction needs
{
  uint2 aMin, aMax, bOff, bMag, cOff, cMag;
  uint32_t shifts;
};

__device__ __forceinline__ AreaRaw load_area_raw(const limgcu_area *areas, uint32_t k)
{
  const char *ar = reinterpret_cast<const char *>(areas + k);
  const uint2 *dq = reinterpret_cast<const uint2 *>(ar + offsetof(limgcu_area, decomp) + offsetof(limgcu_decomp, dirA_min));
  AreaRaw a;
  a.shifts = __ldg(reinterpret_cast<const uint32_t *>(ar + offsetof(limgcu_area, shift)));
  a.aMin = __ldg(dq + 0); a.aMax = __ldg(dq + 1); a.bOff = __ldg(dq + 2); a.bMag = __ldg(dq + 3); a.cOff = __ldg(dq + 4); a.cMag = __ldg(dq + 5);
  return a;
}

template <int CH>
__device__ __forceinline__ void recon_from_raw(const AreaRaw &a, ReconCh &r0, ReconCh &r1, ReconCh &r2, ReconCh &r3)
{
  const int sA = min(a.shifts & 0xFF, 8u), sB = min((a.shifts >> 8) & 0xFF, 8u), sC = min((a.shifts >> 16) & 0xFF, 8u);
  const int32_t mulA = c_decode_mul_rgb[sA], mulB = c_decode_mul_rgb[sB], mulC = c_decode_mul_rgb[sC];
  const bool dropB = sB > 7, dropC = sC > 7;

  r0 = recon_channel_setup(lo16(a.aMin.x), lo16(a.aMax.x), lo16(a.bOff.x), lo16(a.bMag.x), lo16(a.cOff.x), lo16(a.cMag.x), mulA, mulB, mulC, dropB, dropC);
  r1 = recon_channel_setup(hi16(a.aMin.x), hi16(a.aMax.x), hi16(a.bOff.x), hi16(a.bMag.x), hi16(a.cOff.x), hi16(a.cMag.x), mulA, mulB, mulC, dropB, dropC);
  r2 = recon_channel_setup(lo16(a.aMin.y), lo16(a.aMax.y), lo16(a.bOff.y), lo16(a.bMag.y), lo16(a.cOff.y), lo16(a.cMag.y), mulA, mulB, mulC, dropB, dropC);

  if (CH == 4) // Q7: the alpha channel ignores dropped factors (multiplier 1 << 8 with bias 0)
    r3 = recon_channel_setup(hi16(a.aMin.y), hi16(a.aMax.y), hi16(a.bOff.y), hi16(a.bMag.y), hi16(a.cOff.y), hi16(a.cMag.y), sA > 7 ? 256 : mulA, dropB ? 256 : mulB, dropC ? 256 : mulC, false,
                             false);
}

constexpr int kDecodeTileBlocks = 32;                 // blocks per tile = lanes of a warp
constexpr int kDecodeRowBytes = kDecodeTileBlocks * 8; // 256 B of one code plane row
constexpr int kDecodeStageBytes = 3 * 8 * kDecodeRowBytes;

template <int CH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_decode_stream(const limgcu_area *__restrict__ areas, const uint32_t *__restrict__ blockToArea, const uint8_t *__restrict__ codesA,
                                                              const uint8_t *__restrict__ codesB, const uint8_t *__restrict__ codesC, int W, int H, int BX, int tilesX, int tiles,
                                                              uint32_t *__restrict__ dst)
{
  extern __shared__ __align__(128) uint8_t decodeSmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *buf = decodeSmem + (size_t)warp * 2 * kDecodeStageBytes;
  const uint32_t bufS = smem_u32(buf);
  const uint32_t bar0 = smem_u32(decodeSmem + (size_t)WARPS * 2 * kDecodeStageBytes + warp * 16);

  if (lane == 0)
  {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();

  const int nw = gridDim.x * WARPS;
  int t = blockIdx.x * WARPS + warp;

  if (t >= tiles)
    return;

  // copies of tile `tile` into stage `s`: lane = plane * 8 + row
  auto issue = [&](int tile, int s) {
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int x0 = tx * kDecodeRowBytes, rows = min(8, H - ty * 8);
    const uint32_t wbytes = (uint32_t)min(kDecodeRowBytes, W - x0);
    const uint32_t bar = bar0 + 8 * s;

    if (lane == 0)
      mbar_expect_tx(bar, 3u * rows * wbytes);

    __syncwarp();

    const int plane = lane >> 3, row = lane & 7;

    if (lane < 24 && row < rows)
    {
      const uint8_t *src = (plane == 0 ? codesA : (plane == 1 ? codesB : codesC)) + (size_t)(ty * 8 + row) * W + x0;
      bulk_g2s(bufS + s * kDecodeStageBytes + lane * kDecodeRowBytes, src, wbytes, bar);
    }
  };

  auto area_index = [&](int tile) -> uint32_t {
    const int ty = tile / tilesX, tx = tile - ty * tilesX;
    const int bx = min(tx * kDecodeTileBlocks + lane, BX - 1);
    return __ldg(blockToArea + (size_t)ty * BX + bx);
  };

  issue(t, 0);
  AreaRaw cur = load_area_raw(areas, area_index(t));
  uint32_t kNext = t + nw < tiles ? area_index(t + nw) : 0;

  for (int it = 0;; it++)
  {
    const int tn = t + nw;
    const bool more = tn < tiles;
    AreaRaw nxt = cur;
    uint32_t kNext2 = 0;

    if (more)
    {
      issue(tn, (it + 1) & 1);
      nxt = load_area_raw(areas, kNext);

      if (tn + nw < tiles)
        kNext2 = area_index(tn + nw);
    }

    ReconCh r0, r1, r2, r3 = {};
    recon_from_raw<CH>(cur, r0, r1, r2, r3);

    const int ty = t / tilesX, tx = t - ty * tilesX;
    const int bx = tx * kDecodeTileBlocks + lane;
    const int rows = min(8, H - ty * 8);
    const uint8_t *stage = buf + (it & 1) * kDecodeStageBytes + lane * 8;
    uint32_t *out0 = dst + (size_t)ty * 8 * W + bx * 8;

    mbar_wait(bar0 + 8 * (it & 1), (it >> 1) & 1);

    if (bx < BX)
    {
#pragma unroll 2
      for (int j = 0; j < rows; j++)
      {
        const uint2 va = *reinterpret_cast<const uint2 *>(stage + (0 * 8 + j) * kDecodeRowBytes);
        const uint2 vb = *reinterpret_cast<const uint2 *>(stage + (1 * 8 + j) * kDecodeRowBytes);
        const uint2 vc = *reinterpret_cast<const uint2 *>(stage + (2 * 8 + j) * kDecodeRowBytes);
        uint32_t out[8];

#pragma unroll
        for (int i = 0; i < 8; i++)
        {
          const uint32_t wa = i < 4 ? va.x : va.y, wb = i < 4 ? vb.x : vb.y, wc = i < 4 ? vc.x : vc.y;
          const int32_t eA = (int32_t)__byte_perm(wa, 0, 0x4440 + (i & 3));
          const int32_t eB = (int32_t)__byte_perm(wb, 0, 0x4440 + (i & 3));
          const int32_t eC = (int32_t)__byte_perm(wc, 0, 0x4440 + (i & 3));
          const int32_t c0 = recon_sum_ch(r0, eA, eB, eC), c1 = recon_sum_ch(r1, eA, eB, eC), c2 = recon_sum_ch(r2, eA, eB, eC);
          const int32_t c3 = CH == 4 ? recon_sum_ch(r3, eA, eB, eC) : 255;

          out[i] = pack_sat_u8(c1, c0, pack_sat_u8(c3, c2, 0));
        }

        uint4 *p = reinterpret_cast<uint4 *>(out0 + (size_t)j * W);
        p[0] = make_uint4(out[0], out[1], out[2], out[3]);
        p[1] = make_uint4(out[4], out[5], out[6], out[7]);
      }
    }

    if (!more)
      break;

    __syncwarp(); // every lane has read its part of the stage before the copies of tile t + 2 nw overwrite it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    t = tn;
    cur = nxt;
    kNext = kNext2;
  }
}

} // namespace limg

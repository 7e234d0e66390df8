// limg_b200/csrc/kernels_container.cuh -- the bit-packed payload of the .limg container.
//
// The reference has no bitstream: it only accounts for one (limg.cpp:1629-1636: a static per-area header plus
// rangeSize * ((8 - shiftA) + (8 - shiftB) + (8 - shiftC)) payload bits). This is that payload made real:
//
//   payload = for every area in emission order: factor A codes, factor B codes, factor C codes
//   codes of one factor = the area's pixels in area-contiguous order (row-major inside the pixel rectangle, limg.cpp:1752-1753),
//                         (8 - shift) bits each, LSB first, every run of 8 pixels of a row ("segment") starting on a byte boundary
//
// A segment of 8 codes of b bits is exactly b bytes, so for widths that are a multiple of 8 (every config of BASELINE.json) the codes of an
// area and factor are one continuous LSB-first bit stream; a ragged last segment of a row is padded to 8 codes with zeros.
// A dropped factor (shift 8) takes no bits for RGB. For RGBA it keeps its raw byte: the reference's RGBA reconstruction reads it (Q7).
// The byte offset of every area is the exclusive scan of segments * (bitsA + bitsB + bitsC) (k_payload_scan).
#pragma once

#include "common.cuh"

namespace limg
{

__host__ __device__ __forceinline__ int container_code_bits(int shift, bool alpha) { return shift > 7 ? (alpha ? 8 : 0) : 8 - shift; }

__device__ __forceinline__ unsigned long long area_payload_bytes(const limgcu_area &a, bool alpha)
{
  const unsigned long long segs = (unsigned long long)((a.px_w + 7) >> 3) * a.px_h;
  return segs * (unsigned)(container_code_bits(a.shift[0], alpha) + container_code_bits(a.shift[1], alpha) + container_code_bits(a.shift[2], alpha));
}

// offsets[k] = first payload byte of area k, offsets[count] = payload size. Single CTA (the table has at most one entry per 8x8 block).
__global__ void __launch_bounds__(1024) k_payload_scan(const limgcu_area *areas, const uint32_t *areaCountDev, uint32_t areaCountHost, int alpha, unsigned long long *offsets)
{
  __shared__ unsigned long long warpSums[33];
  const uint32_t count = areaCountDev ? *areaCountDev : areaCountHost;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long carry = 0;

  for (uint32_t base = 0; base < count; base += 1024)
  {
    const uint32_t k = base + threadIdx.x;
    const unsigned long long v = k < count ? area_payload_bytes(areas[k], alpha != 0) : 0ull;
    unsigned long long incl = v;

#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += n;
    }

    if (lane == 31)
      warpSums[warp] = incl;

    __syncthreads();

    if (warp == 0)
    {
      const unsigned long long s = warpSums[lane];
      unsigned long long si = s;

#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, si, o);
        if (lane >= o) si += n;
      }

      warpSums[lane] = si - s;

      if (lane == 31)
        warpSums[32] = si;
    }

    __syncthreads();

    if (k < count)
      offsets[k] = carry + warpSums[warp] + incl - v;

    carry += warpSums[32];
    __syncthreads();
  }

  if (threadIdx.x == 0)
    offsets[count] = carry;
}

// where the segment (block column bx, pixel row y) of area `a` lives inside the area's payload, per factor
struct SegmentPlace
{
  unsigned long long off[3];
  int bits[3];
};

__device__ __forceinline__ SegmentPlace place_segment(const limgcu_area &a, unsigned long long areaOff, int bx, int y, bool alpha)
{
  const unsigned long long segsPerRow = (a.px_w + 7) >> 3;
  const unsigned long long segs = segsPerRow * a.px_h;
  const unsigned long long seg = (unsigned long long)(y - (int)a.px_y) * segsPerRow + (unsigned)(bx - (int)a.ox);
  SegmentPlace p;
  unsigned long long base = areaOff;

#pragma unroll
  for (int f = 0; f < 3; f++)
  {
    p.bits[f] = container_code_bits(a.shift[f], alpha);
    p.off[f] = base + seg * (unsigned)p.bits[f];
    base += segs * (unsigned)p.bits[f];
  }

  return p;
}

// One thread per 8-pixel row segment of a block: 8 right-aligned codes of three planes -> bits bytes each.
__global__ void __launch_bounds__(256) k_container_pack(const limgcu_area *__restrict__ areas, const uint32_t *__restrict__ blockToArea, const unsigned long long *__restrict__ offsets,
                                                        const uint8_t *__restrict__ codesA, const uint8_t *__restrict__ codesB, const uint8_t *__restrict__ codesC, int W, int H, int BX, int alpha,
                                                        int vec, uint8_t *__restrict__ payload)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;

  if (t >= (long long)BX * H)
    return;

  const int bx = (int)(t % BX), y = (int)(t / BX);
  const uint32_t k = blockToArea[(size_t)(y >> 3) * BX + bx];
  const limgcu_area a = areas[k];
  const SegmentPlace p = place_segment(a, offsets[k], bx, y, alpha != 0);
  const int npx = min(8, W - bx * 8);
  const size_t src = (size_t)y * W + (size_t)bx * 8;
  const uint8_t *planes[3] = { codesA, codesB, codesC };

#pragma unroll
  for (int f = 0; f < 3; f++)
  {
    const int b = p.bits[f];

    if (b == 0)
      continue;

    unsigned long long raw = 0;

    if (vec && npx == 8)
      raw = *reinterpret_cast<const unsigned long long *>(planes[f] + src);
    else
      for (int i = 0; i < npx; i++)
        raw |= (unsigned long long)planes[f][src + i] << (8 * i);

    unsigned long long acc = 0;
    const unsigned long long mask = (1ull << b) - 1ull;

#pragma unroll
    for (int i = 0; i < 8; i++)
      acc |= ((raw >> (8 * i)) & mask) << (b * i);

    uint8_t *dst = payload + p.off[f];

    for (int i = 0; i < b; i++)
      dst[i] = (uint8_t)(acc >> (8 * i));
  }
}

// the inverse: payload -> right-aligned codes in image layout (a factor without bits reads as code 0)
__global__ void __launch_bounds__(256) k_container_unpack(const limgcu_area *__restrict__ areas, const uint32_t *__restrict__ blockToArea, const unsigned long long *__restrict__ offsets,
                                                          const uint8_t *__restrict__ payload, int W, int H, int BX, int alpha, int vec, uint8_t *__restrict__ codesA,
                                                          uint8_t *__restrict__ codesB, uint8_t *__restrict__ codesC)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;

  if (t >= (long long)BX * H)
    return;

  const int bx = (int)(t % BX), y = (int)(t / BX);
  const uint32_t k = blockToArea[(size_t)(y >> 3) * BX + bx];
  const limgcu_area a = areas[k];
  const SegmentPlace p = place_segment(a, offsets[k], bx, y, alpha != 0);
  const int npx = min(8, W - bx * 8);
  const size_t dstOff = (size_t)y * W + (size_t)bx * 8;
  uint8_t *planes[3] = { codesA, codesB, codesC };

#pragma unroll
  for (int f = 0; f < 3; f++)
  {
    const int b = p.bits[f];
    unsigned long long acc = 0;
    const uint8_t *src = payload + p.off[f];

    for (int i = 0; i < b; i++)
      acc |= (unsigned long long)src[i] << (8 * i);

    unsigned long long raw = 0;
    const unsigned long long mask = b ? (1ull << b) - 1ull : 0ull;

#pragma unroll
    for (int i = 0; i < 8; i++)
      raw |= ((acc >> (b * i)) & mask) << (8 * i);

    if (vec && npx == 8)
      *reinterpret_cast<unsigned long long *>(planes[f] + dstOff) = raw;
    else
      for (int i = 0; i < npx; i++)
        planes[f][dstOff + i] = (uint8_t)(raw >> (8 * i));
  }
}

} // namespace limg

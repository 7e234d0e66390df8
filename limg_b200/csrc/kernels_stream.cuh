// limg_b200/csrc/kernels_stream.cuh -- the streaming passes: dither-chain scan, the fused projection + dither + bit-crush
// + plane writer ("finalize"), the standalone reconstruction (decode) and limg_compare.
#pragma once

#include "common.cuh"

namespace limg
{

// ---------------------------------------------------------------------------------------------
// dither chain bookkeeping (limg.cpp:1539-1549, Q13): the reference threads ONE 64-bit LCG state through every plane of
// every area in emission order. With the per-area demand known, an exclusive scan gives every area its position in the
// chain and the LCG's O(log n) jump-ahead gives the state there, so the finalize pass needs no serial dependency.
// ---------------------------------------------------------------------------------------------

// k_dither_scan (single CTA) only does the exclusive scan of the demand; the two jump-aheads per area are ~50 dependent 64-bit
// multiply-adds each and run grid-wide in k_dither_states (both in one CTA took 71 us at 4K, the split takes < 10).
__global__ void __launch_bounds__(1024) k_dither_scan(const uint32_t *areaCount, const uint64_t *demand, unsigned long long *before, int restartEveryArea)
{
  __shared__ unsigned long long warpSums[33];
  const uint32_t count = *areaCount;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long carry = 0;

  for (uint32_t base = 0; base < count; base += 1024)
  {
    const uint32_t k = base + threadIdx.x;
    const unsigned long long v = k < count ? demand[k] : 0ull;
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
      before[k] = restartEveryArea ? 0ull : carry + warpSums[warp] + incl - v;

    carry += warpSums[32];
    __syncthreads();
  }
}

// Row-band sharding (limgcu_encode_areas / limgcu_finalize_rows): what the per-area encode produced, as 20 words per area that are zero
// for the areas another rank owns, so that a SUM all-reduce over the ranks assembles the complete table.
#define LIMG_AREA_RESULT_WORDS 20

__global__ void __launch_bounds__(256) k_pack_area_results(const limgcu_area *areas, const uint32_t *areaCount, const uint64_t *demand, uint32_t rowLo, uint32_t rowHi, uint32_t *out)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;

  if (k >= *areaCount)
    return;

  uint32_t *o = out + (size_t)k * LIMG_AREA_RESULT_WORDS;
  const limgcu_area &a = areas[k];
  const bool mine = a.oy >= rowLo && a.oy < rowHi;
  const uint32_t *d = reinterpret_cast<const uint32_t *>(&a.decomp);

  for (int i = 0; i < 16; i++)
    o[i] = mine ? d[i] : 0u;

  o[16] = mine ? ((uint32_t)a.shift[0] | ((uint32_t)a.shift[1] << 8) | ((uint32_t)a.shift[2] << 16)) : 0u;
  o[17] = mine ? (uint32_t)demand[k] : 0u;
  o[18] = mine ? (uint32_t)(demand[k] >> 32) : 0u;
  o[19] = mine ? 1u : 0u; // after the all-reduce: exactly one owner per area
}

__global__ void __launch_bounds__(256) k_unpack_area_results(limgcu_area *areas, const uint32_t *areaCount, uint64_t *demand, const uint32_t *in, uint32_t *badOwners)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;

  if (k >= *areaCount)
    return;

  const uint32_t *r = in + (size_t)k * LIMG_AREA_RESULT_WORDS;
  limgcu_area &a = areas[k];
  uint32_t *d = reinterpret_cast<uint32_t *>(&a.decomp);

  for (int i = 0; i < 16; i++)
    d[i] = r[i];

  a.shift[0] = (uint8_t)(r[16] & 0xFF); a.shift[1] = (uint8_t)((r[16] >> 8) & 0xFF); a.shift[2] = (uint8_t)((r[16] >> 16) & 0xFF);
  a.pad = 0;
  demand[k] = (uint64_t)r[17] | ((uint64_t)r[18] << 32);

  if (r[19] != 1u)
    atomicAdd(badOwners, 1u);
}

// AES mode: the chain states around every area come from the host
__global__ void __launch_bounds__(256) k_set_dither_states(limgcu_area *areas, const unsigned long long *before, const unsigned long long *after, uint32_t count)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;

  if (k >= count)
    return;

  areas[k].ditherBefore = before[k];
  areas[k].ditherAfter = after[k];
}

// bandAreas > 0 (non-merged encoder with a thread pool, limg.cpp:1893, 2108-2137): the reference restarts the chain at the top of every
// y-band, i.e. every `bandAreas` areas (areas = blocks in raster order there); the last of the `bandCount` bands takes the rest.
__global__ void __launch_bounds__(256) k_dither_states(limgcu_area *areas, const uint32_t *areaCount, const uint64_t *demand, const unsigned long long *before, LcgJumpTable jt,
                                                       uint32_t bandAreas, uint32_t bandCount)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;

  if (k >= *areaCount)
    return;

  unsigned long long steps = before[k];

  if (bandAreas)
    steps -= before[min(k / bandAreas, bandCount - 1u) * bandAreas];

  const uint64_t s0 = lcg_jump(LIMG_DITHER_SEED, steps, jt);
  areas[k].ditherBefore = s0;
  areas[k].ditherAfter = lcg_jump(s0, demand[k], jt);
}

// ---------------------------------------------------------------------------------------------
// finalize: one thread per 8-pixel row segment of an 8x8 block (always inside one area).
//   projection (limg_factorization.h:101-197) -> dither + shift-down (limg.cpp:798-822) -> right-aligned codes,
//   the reference's 12 info planes (limg.cpp:1594-1707, Q12) and the in-encoder reconstruction (limg_decode.h:39-236).
// ---------------------------------------------------------------------------------------------

struct FinalizeArgs
{
  const uint32_t *src;
  int W, H, BX;
  const limgcu_area *areas;
  const uint32_t *blockToArea;
  uint8_t *codesA, *codesB, *codesC;
  limgcu_planes planes;
  LcgJumpTable jt;
  const uint8_t *noise;                // AES mode: one noise byte per pixel and dithered plane, area-contiguous (dither_aes_host.cpp); else nullptr
  const unsigned long long *noiseOff;  // [3 * area + plane] -> offset into noise
  int yLo, yHi; // pixel rows handled by this launch (row-band sharding; the default is the whole image)
  int vec; // W % 8 == 0 and every plane pointer 32-byte aligned: 128-bit loads / stores
};

__device__ __forceinline__ void store8(uint8_t *dst, const uint32_t v[8], int npx, bool aligned)
{
  if (npx == 8 && aligned)
  {
    uint2 w;
    w.x = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    w.y = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
    *reinterpret_cast<uint2 *>(dst) = w;
  }
  else
  {
    for (int j = 0; j < npx; j++)
      dst[j] = (uint8_t)v[j];
  }
}

__device__ __forceinline__ void store8_u32(uint32_t *dst, const uint32_t v[8], int npx, bool aligned)
{
  if (npx == 8 && aligned)
  {
    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
  }
  else
  {
    for (int j = 0; j < npx; j++)
      dst[j] = v[j];
  }
}

__device__ __forceinline__ void fill8_u32(uint32_t *dst, uint32_t v, int npx, bool aligned)
{
  if (npx == 8 && aligned)
  {
    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(v, v, v, v);
    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(v, v, v, v);
  }
  else
  {
    for (int j = 0; j < npx; j++)
      dst[j] = v;
  }
}

template <int CH>
__global__ void __launch_bounds__(256) k_finalize(FinalizeArgs a)
{
  const int segsPerRow = (a.W + 7) >> 3;
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;

  if (s >= (long long)segsPerRow * (a.yHi - a.yLo))
    return;

  const int y = a.yLo + (int)(s / segsPerRow);
  const int bx = (int)(s - (long long)(y - a.yLo) * segsPerRow);
  const int x0 = bx * 8;
  const int npx = min(8, a.W - x0);
  const size_t rowOff = (size_t)y * a.W + x0;
  const bool aligned = a.vec != 0; // every segment start is then 32-byte aligned in the u32 planes, 8-byte in the u8 planes

  const uint32_t k = a.blockToArea[(size_t)(y >> 3) * a.BX + bx];
  const limgcu_area *ar = &a.areas[k];
  const limgcu_decomp d = ar->decomp;
  const int sA = ar->shift[0], sB = ar->shift[1], sC = ar->shift[2];
  const uint32_t n = ar->px_w * ar->px_h;
  const uint32_t i0 = (uint32_t)(y - (int)ar->px_y) * ar->px_w + (uint32_t)(x0 - (int)ar->px_x);

  uint32_t px[8];

  if (npx == 8 && aligned)
  {
    const uint4 p0 = reinterpret_cast<const uint4 *>(a.src + rowOff)[0];
    const uint4 p1 = reinterpret_cast<const uint4 *>(a.src + rowOff)[1];
    px[0] = p0.x; px[1] = p0.y; px[2] = p0.z; px[3] = p0.w;
    px[4] = p1.x; px[5] = p1.y; px[6] = p1.z; px[7] = p1.w;
  }
  else
  {
#pragma unroll
    for (int j = 0; j < 8; j++)
      px[j] = j < npx ? a.src[rowOff + j] : 0u;
  }

  Proj p;
  init_proj<CH>(d, p);

  uint32_t cA[8], cB[8], cC[8];

#pragma unroll
  for (int j = 0; j < 8; j++)
  {
    const uint32_t f = project_px<CH>(p, px[j]);
    cA[j] = f & 0xFF;
    cB[j] = (f >> 8) & 0xFF;
    cC[j] = (f >> 16) & 0xFF;
  }

  // dither: plane order A, B, C; only planes with 0 < shift < 8 consume the chain (limg.cpp:1541-1548)
  {
    uint64_t planeIndex = 0;
    const int sh[3] = { sA, sB, sC };
    uint32_t *codes[3] = { cA, cB, cC };

#pragma unroll
    for (int pl = 0; pl < 3; pl++)
    {
      if (sh[pl] != 0 && sh[pl] != 8 && a.noise != nullptr)
      {
        const uint8_t *nz = a.noise + a.noiseOff[3 * (size_t)k + pl] + i0;

#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j < npx)
            codes[pl][j] = dither_one(codes[pl][j], nz[j], sh[pl]);
      }
      else if (sh[pl] != 0 && sh[pl] != 8)
      {
        uint64_t h = lcg_jump(ar->ditherBefore, planeIndex * n + i0, a.jt);

#pragma unroll
        for (int j = 0; j < 8; j++)
        {
          h = h * LIMG_LCG_MUL + 1ull;
          codes[pl][j] = dither_one(codes[pl][j], pcg_output(h), sh[pl]);
        }

        planeIndex++;
      }
    }
  }

  if (a.codesA) store8(a.codesA + rowOff, cA, npx, aligned);
  if (a.codesB) store8(a.codesB + rowOff, cB, npx, aligned);
  if (a.codesC) store8(a.codesC + rowOff, cC, npx, aligned);

  const limgcu_planes &pl = a.planes;
  uint32_t tmp[8];

  if (pl.pFactorsA)
  {
#pragma unroll
    for (int j = 0; j < 8; j++) tmp[j] = (cA[j] << sA) & 0xFF; // left-aligned view (limg.cpp:1655)
    store8(pl.pFactorsA + rowOff, tmp, npx, aligned);
  }

  if (pl.pFactorsB)
  {
#pragma unroll
    for (int j = 0; j < 8; j++) tmp[j] = (cB[j] << sB) & 0xFF;
    store8(pl.pFactorsB + rowOff, tmp, npx, aligned);
  }

  if (pl.pFactorsC)
  {
#pragma unroll
    for (int j = 0; j < 8; j++) tmp[j] = (cC[j] << sC) & 0xFF;
    store8(pl.pFactorsC + rowOff, tmp, npx, aligned);
  }

  if (pl.pBitsPerPixel)
  {
    const uint32_t headerBits = CH * 9 * 2 + CH * 8 + 32; // 110 / 136 (limg.cpp:1630)
    const unsigned long long bits = headerBits + (unsigned long long)n * (unsigned)((8 - sA) + (8 - sB) + (8 - sC));
    const uint32_t bpp = (uint32_t)((bits + n / 2) / n) & 0xFF;
#pragma unroll
    for (int j = 0; j < 8; j++) tmp[j] = bpp;
    store8(pl.pBitsPerPixel + rowOff, tmp, npx, aligned);
  }

  if (pl.pShiftABCX)
  {
    // bit_to_pattern = {0,0x22,0x44,0x66,0x88,0xAA,0xCC,0xEE,0xFF} (limg.cpp:1596): s < 8 ? s * 0x22 : 0xFF
    const uint32_t pa = sA < 8 ? sA * 0x22 : 0xFF, pb = sB < 8 ? sB * 0x22 : 0xFF, pc = sC < 8 ? sC * 0x22 : 0xFF;
    fill8_u32(pl.pShiftABCX + rowOff, 0xFF000000u | (pa << 16) | (pb << 8) | pc, npx, aligned);
  }

  if (pl.pColAMin || pl.pColAMax || pl.pColBMin || pl.pColBMax || pl.pColCMin || pl.pColCMax)
  {
    uint32_t col[6] = { 0, 0, 0, 0, 0, 0 };

#pragma unroll
    for (int i = 0; i < CH; i++)
    {
      col[0] |= (uint32_t)clamp255(d.dirA_min[i]) << (8 * i);
      col[1] |= (uint32_t)clamp255(d.dirA_max[i]) << (8 * i);
      col[2] |= (uint32_t)clamp255(d.dirB_offset[i] + 0x80) << (8 * i);
      col[3] |= (uint32_t)clamp255(d.dirB_mag[i] + 0x80) << (8 * i);
      col[4] |= (uint32_t)clamp255(d.dirC_offset[i] + 0x80) << (8 * i);
      col[5] |= (uint32_t)clamp255(d.dirC_mag[i] + 0x80) << (8 * i);
    }

    if (CH == 3)
    {
#pragma unroll
      for (int i = 0; i < 6; i++) col[i] |= 0xFF000000u;
    }

    if (pl.pColAMin) fill8_u32(pl.pColAMin + rowOff, col[0], npx, aligned);
    if (pl.pColAMax) fill8_u32(pl.pColAMax + rowOff, col[1], npx, aligned);
    if (pl.pColBMin) fill8_u32(pl.pColBMin + rowOff, col[2], npx, aligned);
    if (pl.pColBMax) fill8_u32(pl.pColBMax + rowOff, col[3], npx, aligned);
    if (pl.pColCMin) fill8_u32(pl.pColCMin + rowOff, col[4], npx, aligned);
    if (pl.pColCMax) fill8_u32(pl.pColCMax + rowOff, col[5], npx, aligned);
  }

  if (pl.pBlockIndex)
    fill8_u32(pl.pBlockIndex + rowOff, 0xFF000000u | (k + 1), npx, aligned);

  if (pl.pDecoded)
  {
    Recon r;
    init_recon<CH>(d, sA, sB, sC, 0xFFFF, r);

#pragma unroll
    for (int j = 0; j < 8; j++)
    {
      const int32_t eA = (int32_t)cA[j], eB = (int32_t)cB[j], eC = (int32_t)cC[j];
      tmp[j] = (uint32_t)recon_channel(r, 0, eA, eB, eC) | ((uint32_t)recon_channel(r, 1, eA, eB, eC) << 8) |
               ((uint32_t)recon_channel(r, 2, eA, eB, eC) << 16) | ((uint32_t)recon_channel(r, 3, eA, eB, eC) << 24);
    }

    store8_u32(pl.pDecoded + rowOff, tmp, npx, aligned);
  }
}

// ---------------------------------------------------------------------------------------------
// decode: streaming reconstruction from the compact stream (limg_decode.h:326-340 per area)
// ---------------------------------------------------------------------------------------------

// One thread reconstructs four rows of one 8x8 block: the area's reconstruction state is set up once per 32 pixels, the twelve
// 8-byte code loads and eight 16-byte stores of a thread are independent (memory-level parallelism instead of occupancy), and
// neighbouring threads cover neighbouring blocks of the same rows, so a warp touches 256 contiguous bytes of every code row and 1 KB
// of every output row. For RGB the alpha byte is the constant 0xFF the reference's 0xFFFF "min" trick produces (limg_decode.h:95-97).
template <int CH>
__global__ void __launch_bounds__(256, 4) k_decode(const limgcu_area *__restrict__ areas, const uint32_t *__restrict__ blockToArea, const uint8_t *__restrict__ codesA,
                                                const uint8_t *__restrict__ codesB, const uint8_t *__restrict__ codesC, int W, int H, int BX, uint32_t *__restrict__ dst, int vec)
{
  const int BY = (H + 7) >> 3;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;

  if (t >= (long long)BX * BY * 2)
    return;

  // consecutive threads = consecutive blocks of one half-row of blocks
  const int bx = (int)(t % BX);
  const int rest = (int)(t / BX);
  const int by = rest >> 1, half = rest & 1;
  const int x0 = bx * 8, y0 = by * 8 + half * 4;

  if (y0 >= H)
    return;

  const int npx = min(8, W - x0), nrows = min(4, H - y0);
  const bool fast = vec != 0 && npx == 8;
  const uint32_t k = blockToArea[(size_t)by * BX + bx];
  const limgcu_area *ar = &areas[k];
  Recon r;
  init_recon<CH>(ar->decomp, ar->shift[0], ar->shift[1], ar->shift[2], 0xFFFF, r);

  uint2 va[4], vb[4], vc[4];

  if (fast)
  {
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      if (j < nrows)
      {
        const size_t off = (size_t)(y0 + j) * W + x0;
        va[j] = __ldg(reinterpret_cast<const uint2 *>(codesA + off));
        vb[j] = __ldg(reinterpret_cast<const uint2 *>(codesB + off));
        vc[j] = __ldg(reinterpret_cast<const uint2 *>(codesC + off));
      }
    }
  }

#pragma unroll
  for (int j = 0; j < 4; j++)
  {
    if (j >= nrows)
      break;

    const size_t off = (size_t)(y0 + j) * W + x0;
    uint32_t out[8];

#pragma unroll
    for (int i = 0; i < 8; i++)
    {
      int32_t eA, eB, eC;

      if (fast)
      {
        eA = (int32_t)(((i < 4 ? va[j].x : va[j].y) >> (8 * (i & 3))) & 0xFF);
        eB = (int32_t)(((i < 4 ? vb[j].x : vb[j].y) >> (8 * (i & 3))) & 0xFF);
        eC = (int32_t)(((i < 4 ? vc[j].x : vc[j].y) >> (8 * (i & 3))) & 0xFF);
      }
      else
      {
        eA = i < npx ? codesA[off + i] : 0;
        eB = i < npx ? codesB[off + i] : 0;
        eC = i < npx ? codesC[off + i] : 0;
      }

      uint32_t px = (uint32_t)recon_channel(r, 0, eA, eB, eC) | ((uint32_t)recon_channel(r, 1, eA, eB, eC) << 8) | ((uint32_t)recon_channel(r, 2, eA, eB, eC) << 16);

      if (CH == 4)
        px |= (uint32_t)recon_channel(r, 3, eA, eB, eC) << 24;
      else
        px |= 0xFF000000u;

      out[i] = px;
    }

    store8_u32(dst + off, out, npx, fast);
  }
}

// block map from an area table produced elsewhere
// (a rectangle that leaves the block grid is clipped: a table from outside never makes this kernel write out of bounds)
__global__ void k_block_map(const limgcu_area *__restrict__ areas, uint32_t count, int BX, int BY, uint32_t *__restrict__ blockToArea)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;

  if (k >= count)
    return;

  const limgcu_area ar = areas[k];
  const uint32_t x0 = min(ar.ox, (uint32_t)BX), y0 = min(ar.oy, (uint32_t)BY);
  const uint32_t x1 = x0 + min(ar.rx, (uint32_t)BX - x0), y1 = y0 + min(ar.ry, (uint32_t)BY - y0);

  for (uint32_t yy = y0; yy < y1; yy++)
    for (uint32_t xx = x0; xx < x1; xx++)
      blockToArea[(size_t)yy * BX + xx] = k;
}

// ---------------------------------------------------------------------------------------------
// limg_compare (limg.cpp:2455-2491, limg_internal.h:376-410)
// ---------------------------------------------------------------------------------------------

template <int CH>
__global__ void __launch_bounds__(256) k_compare(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t n, unsigned long long *__restrict__ total)
{
  unsigned long long acc = 0;

  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const uint32_t pa = a[i], pb = b[i];
    const int32_t dr = (int32_t)(pa & 0xFF) - (int32_t)(pb & 0xFF);
    const int32_t dg = (int32_t)((pa >> 8) & 0xFF) - (int32_t)((pb >> 8) & 0xFF);
    const int32_t db = (int32_t)((pa >> 16) & 0xFF) - (int32_t)((pb >> 16) & 0xFF);
    const int32_t da = (int32_t)(pa >> 24) - (int32_t)(pb >> 24);
    const int32_t rr = dr * dr;
    const bool low = rr < 0x4000;
    uint32_t e = (uint32_t)(rr * (low ? 2 : 3) + dg * dg * 4 + db * db * (low ? 3 : 2));

    if (CH == 4)
      e += (uint32_t)(da * da * 3);

    acc += e;
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);

  if ((threadIdx.x & 31) == 0)
    atomicAdd(total, acc);
}

} // namespace limg

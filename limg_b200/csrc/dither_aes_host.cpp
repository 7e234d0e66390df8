// limg_b200/csrc/dither_aes_host.cpp -- the reference's AES-round dither chain (limg.cpp:824-879), host side.
//
// On hosts with SSE4.1 + AES-NI the reference draws its dither noise from `state = aesdec(state, key)`, one round per 8 pixels, threaded
// through every plane of every area in emission order. The chain has no skip-ahead and is ~n/8 dependent rounds per image (3.1 M at 4K),
// which is 1.3 ns per round on a CPU with AES-NI and ~100 ns in software on one GPU thread, so the GPU path keeps the LCG generator as its
// primary mode (what the reference computes without AES-NI) and this file provides the compatibility mode: once the
// shifts of all areas are known the chain is walked here, one noise byte per pixel and dithered plane is written in area-contiguous order,
// and k_finalize reads that stream instead of jumping the LCG. Only bits 0..6 of a 16-bit lane can reach the result (the mask is
// (1 << shift) - 1 with shift <= 7), so a byte per pixel is enough.
#include "../../include/limgcu.h"

#include <immintrin.h>
#include <stdint.h>
#include <string.h>

namespace limg
{

static const uint64_t kLcgMul = 6364136223846793005ULL;
static const uint64_t kAesKeyLo = 0x824A73EAAB705E1DULL, kAesKeyHi = 0x2A76E98006CB4CADULL; // limg.cpp:836

static inline uint32_t pcg_output_host(uint64_t h)
{
  const uint32_t xs = (uint32_t)(((h >> 18) ^ h) >> 27);
  const uint32_t rot = (uint32_t)(h >> 59);
  return (xs >> rot) | (xs << ((32 - rot) & 31));
}

// ---- AESDEC in software (hosts without AES-NI): InvShiftRows, InvSubBytes, InvMixColumns, xor key --------------------------------

struct AesTables
{
  uint32_t td[4][256]; // td[r][x]: column contribution of inverse-sbox(x) entering at row r, bytes little endian (row 0 in bits 0..7)
  AesTables()
  {
    uint8_t sbox[256], inv[256];
    uint8_t p = 1, q = 1;

    do // walk the multiplicative group with generator 3: p * q == 1 throughout
    {
      p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1B : 0));
      q ^= (uint8_t)(q << 1); q ^= (uint8_t)(q << 2); q ^= (uint8_t)(q << 4);
      if (q & 0x80) q ^= 0x09;
      const uint8_t x = (uint8_t)(q ^ (uint8_t)((q << 1) | (q >> 7)) ^ (uint8_t)((q << 2) | (q >> 6)) ^ (uint8_t)((q << 3) | (q >> 5)) ^ (uint8_t)((q << 4) | (q >> 4)));
      sbox[p] = (uint8_t)(x ^ 0x63);
    } while (p != 1);

    sbox[0] = 0x63;

    for (int i = 0; i < 256; i++)
      inv[sbox[i]] = (uint8_t)i;

    for (int x = 0; x < 256; x++)
    {
      const uint8_t s = inv[x];
      const uint8_t m[4] = { mul(s, 14), mul(s, 9), mul(s, 13), mul(s, 11) }; // column (14 9 13 11)^T of the inverse MixColumns matrix

      for (int r = 0; r < 4; r++) // an input in row r contributes m[(i - r) & 3] to output row i
      {
        uint32_t w = 0;
        for (int i = 0; i < 4; i++)
          w |= (uint32_t)m[(i - r) & 3] << (8 * i);
        td[r][x] = w;
      }
    }
  }

  static uint8_t mul(uint8_t a, uint8_t b)
  {
    uint8_t r = 0;
    for (int i = 0; i < 8; i++)
    {
      if (b & 1) r ^= a;
      a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1B : 0));
      b >>= 1;
    }
    return r;
  }
};

static inline void aesdec_soft(uint32_t st[4], const uint32_t key[4], const AesTables &t)
{
  // InvShiftRows: output column c takes row r from input column (c - r) & 3
  uint32_t o[4];

  for (int c = 0; c < 4; c++)
    o[c] = t.td[0][st[c] & 0xFF] ^ t.td[1][(st[(c + 3) & 3] >> 8) & 0xFF] ^ t.td[2][(st[(c + 2) & 3] >> 16) & 0xFF] ^ t.td[3][st[(c + 1) & 3] >> 24] ^ key[c];

  memcpy(st, o, sizeof(o));
}

// noise bytes of one dithered plane of one area (limg_encode_dither_aes_sse41): returns the chain state after it
static uint64_t plane_noise_soft(uint64_t h, size_t n, uint8_t *out, const AesTables &t)
{
  size_t i = 0;

  if (n >= 8)
  {
    const uint64_t inv = ~h;
    uint32_t st[4], key[4];
    memcpy(st, &h, 8); memcpy(st + 2, &inv, 8);
    memcpy(key, &kAesKeyLo, 8); memcpy(key + 2, &kAesKeyHi, 8);

    for (; i + 8 <= n; i += 8)
    {
      aesdec_soft(st, key, t);

      for (int k = 0; k < 8; k++)
        out[i + k] = (uint8_t)(st[k >> 1] >> (16 * (k & 1)));
    }

    memcpy(&h, st, 8);
  }

  for (; i < n; i++)
  {
    h = h * kLcgMul + 1;
    out[i] = (uint8_t)pcg_output_host(h);
  }

  return h;
}

__attribute__((target("aes,sse4.1"))) static uint64_t plane_noise_aesni(uint64_t h, size_t n, uint8_t *out)
{
  size_t i = 0;

  if (n >= 8)
  {
    __m128i state = _mm_set_epi64x((long long)~h, (long long)h);
    const __m128i key = _mm_set_epi64x((long long)kAesKeyHi, (long long)kAesKeyLo);
    const __m128i low8 = _mm_set_epi8(-1, -1, -1, -1, -1, -1, -1, -1, 14, 12, 10, 8, 6, 4, 2, 0);

    for (; i + 8 <= n; i += 8)
    {
      state = _mm_aesdec_si128(state, key);
      const long long v = _mm_cvtsi128_si64(_mm_shuffle_epi8(state, low8));
      memcpy(out + i, &v, 8);
    }

    h = (uint64_t)_mm_cvtsi128_si64(state);
  }

  for (; i < n; i++)
  {
    h = h * kLcgMul + 1;
    out[i] = (uint8_t)pcg_output_host(h);
  }

  return h;
}

bool host_has_aesni()
{
  return __builtin_cpu_supports("aes") && __builtin_cpu_supports("sse4.1");
}

// Walks the chain over `count` areas in emission order. noise: one byte per pixel of every dithered plane (0 < shift < 8), planes of an area in
// the order A, B, C; planeOff[3 k + p]: offset of that plane's bytes (~0 for planes that do not consume the chain); before / after: the
// 64-bit chain state around every area. Returns the number of noise bytes written (<= 3 * pixels). forceSoftware: test hook.
uint64_t aes_dither_chain_host(const limgcu_area *areas, uint32_t count, uint64_t seed, uint8_t *noise, uint64_t *planeOff, uint64_t *before, uint64_t *after, bool forceSoftware,
                               uint32_t bandAreas, uint32_t bandCount)
{
  static const AesTables tables;
  const bool ni = !forceSoftware && host_has_aesni();
  uint64_t h = seed, cursor = 0;

  for (uint32_t k = 0; k < count; k++)
  {
    const size_t n = (size_t)areas[k].px_w * areas[k].px_h;

    // non-merged encoder with a thread pool: the chain restarts at the top of every y-band (limg.cpp:1893, 2108-2137)
    if (bandAreas && k % bandAreas == 0 && k / bandAreas < bandCount)
      h = seed;

    before[k] = h;

    for (int p = 0; p < 3; p++)
    {
      const int s = areas[k].shift[p];

      if (s == 0 || s > 7) // limg.cpp:1541-1548
      {
        planeOff[3 * (size_t)k + p] = ~0ull;
        continue;
      }

      planeOff[3 * (size_t)k + p] = cursor;
      h = ni ? plane_noise_aesni(h, n, noise + cursor) : plane_noise_soft(h, n, noise + cursor, tables);
      cursor += n;
    }

    after[k] = h;
  }

  return cursor;
}

} // namespace limg

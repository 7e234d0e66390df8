/* tools/scan_stats.c -- DESIGN TOOL (not product code): runs the reference-order scan on a pass-1 table and counts how often a
 * seed's masked growth equals its mask-free growth (the basis of the scan's "predicted result" fast path). */
#include "merge_model.c"

static int rect_free(Model *m, int ox, int oy, int rx, int ry)
{
  for (int y = oy; y < oy + ry; y++)
    for (int x = ox; x < ox + rx; x++)
      if (used_at(m, x, y, NONE))
        return 0;
  return 1;
}

void scan_stats(const lo_decomp *table, int BX, int BY, int CH, double *out /* [64] */)
{
  Model m;
  m.BX = BX; m.BY = BY; m.CH = CH; m.table = table; m.nPred = 0;
  size_t cap = 1;
  while (cap < (size_t)BX * BY * 64) cap <<= 1;
  m.memoMask = cap - 1;
  m.memoKey = (uint64_t *)malloc(cap * 8);
  m.memoVal = (uint8_t *)malloc(cap);
  memset(m.memoKey, 0xFF, cap * 8);
  m.owner = (uint32_t *)malloc((size_t)BX * BY * 4);
  memset(m.owner, 0xFF, (size_t)BX * BY * 4);
  /* plan filter model: which stage-0 candidates are probably still alive when the scan reaches them? cover[b] = smallest raster index of a
   * candidate whose mask-free rectangle covers block b; a candidate is "alive" if no earlier candidate covers its 3x3 corner; second round
   * with the alive candidates of the first round only. */
  uint32_t *cover = (uint32_t *)malloc((size_t)BX * BY * 4);
  uint8_t *alive = (uint8_t *)calloc((size_t)BX * BY, 1), *cand0 = (uint8_t *)calloc((size_t)BX * BY, 1);
  uint8_t *r0x = (uint8_t *)calloc((size_t)BX * BY, 1), *r0y = (uint8_t *)calloc((size_t)BX * BY, 1);
  for (int y = 0; y < BY; y++)
    for (int x = 0; x < BX; x++)
      if (is_cand(&m, x, y, 0))
      {
        const Result r0 = expand(&m, x, y, 0, 0u);
        cand0[y * BX + x] = 1; r0x[y * BX + x] = (uint8_t)(r0.rx > 8 ? 8 : r0.rx); r0y[y * BX + x] = (uint8_t)(r0.ry > 8 ? 8 : r0.ry);
      }
  for (int round = 0; round < (int)out[63] + 1; round++)
  {
    memset(cover, 0xFF, (size_t)BX * BY * 4);
    for (int s = 0; s < BX * BY; s++)
      if (cand0[s] && (round == 0 || alive[s]))
      {
        const int x = s % BX, y = s / BX;
        for (int dy = 0; dy < r0y[s]; dy++)
          for (int dx = 0; dx < r0x[s]; dx++)
            if (cover[(y + dy) * BX + x + dx] > (uint32_t)s) cover[(y + dy) * BX + x + dx] = (uint32_t)s;
      }
    for (int s = 0; s < BX * BY; s++)
      if (cand0[s])
      {
        const int x = s % BX, y = s / BX;
        int ok = 1;
        for (int dy = 0; dy < 3; dy++)
          for (int dx = 0; dx < 3; dx++)
            if (cover[(y + dy) * BX + x + dx] < (uint32_t)s) ok = 0;
        alive[s] = (uint8_t)ok;
      }
  }
  const int rounds = (int)out[63] + 1;
  memset(out, 0, 64 * sizeof(double));
  out[62] = rounds;
  for (int s = 0; s < BX * BY; s++) { out[60] += cand0[s]; out[61] += alive[s]; }

  for (int stage = 0; stage < 2; stage++)
  {
    double *o = out + stage * 32;

    for (int y = 0; y < BY; y++)
      for (int x = 0; x < BX; x++)
      {
        if (used_at(&m, x, y, NONE) || !is_cand(&m, x, y, stage))
          continue;

        for (int k = 0;; k++)
        {
          const Result r = expand(&m, x, y, stage, NONE);
          const Result r0 = expand(&m, x, y, stage, 0u);
          o[0]++;                                           /* expansions */
          if (r.kind == 0) { o[1]++; break; }               /* nothing emitted */
          if (stage == 0 && !alive[y * BX + x]) o[14]++;    /* an emitting seed the filter declared dead */
          const int sameR = r.rx == r0.rx && r.ry == r0.ry;
          const int freeR0 = rect_free(&m, x, y, r0.rx, r0.ry);
          o[2] += sameR; o[3] += freeR0;
          if (freeR0 && !sameR) o[4]++;                     /* must stay 0 */
          const int in8 = r.rx < 8 && r.ry < 8;             /* masked growth decided inside the 8x8 window */
          o[5] += in8;
          o[12] += in8 || freeR0;

          if (r.attempted)
          {
            o[6]++;
            const int sameC = r0.attempted && r.cox == r0.cox && r.coy == r0.coy && r.crx == r0.crx && r.cry == r0.cry;
            const int freeC0 = r0.attempted && rect_free(&m, r0.cox, r0.coy, r0.crx, r0.cry);
            o[7] += sameC;
            o[8] += freeR0 && freeC0;
            if (freeR0 && freeC0 && !sameC) o[9]++;         /* must stay 0 */
            /* four-way result within +-8 of the centre */
            const int cx = x + r.rx / 3, cy = y + r.ry / 3;
            o[10] += (cx - r.cox < 8 && cy - r.coy < 8 && r.cox + r.crx - cx < 8 && r.coy + r.cry - cy < 8);
            o[11] += r.kind == 2;
          }
          else if (stage == 0) o[13]++;

          const int ox = r.kind == 2 ? r.cox : x, oy = r.kind == 2 ? r.coy : y, rx = r.kind == 2 ? r.crx : r.rx, ry = r.kind == 2 ? r.cry : r.ry;
          for (int yy = oy; yy < oy + ry; yy++)
            for (int xx = ox; xx < ox + rx; xx++)
              m.owner[yy * BX + xx] = 1;

          if (r.kind == 2 && !used_at(&m, x, y, NONE)) continue;
          break;
        }
      }
  }

  free(m.memoKey); free(m.memoVal); free(m.owner);
}

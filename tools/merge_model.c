/* tools/merge_model.c -- CPU model of the row-pipelined ("wave") merge scan and its verification pass.
 *
 * DESIGN TOOL, not product code: it links the oracle (lo_matches) and is used to choose the lag rule of k_merge_wave and to
 * check the exactness argument (self-consistent record  =>  the sequential scan's result) before spending GPU time.
 *
 * Model: W workers take block rows by ticket. A worker walks the candidate seeds of its row left to right. Before a seed is
 * expanded against the LIVE in-use mask the row above must have committed everything left of (probe box right edge + margin).
 * Reads happen at the start of an action, writes (claims, progress) at its end, so other workers see stale state in between.
 * Afterwards every candidate is replayed against the mask "owner time < my time" (verification).
 */
#include "../oracle/limg_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NONE 0xFFFFFFFFu
#define DONE 0x7FFFFFFF

typedef struct
{
  int BX, BY, CH;
  const lo_decomp *table;
  uint64_t *memoKey;
  uint8_t *memoVal;
  size_t memoMask;
  uint32_t *owner; /* live: time of the covering rectangle or NONE */
  uint64_t nPred;
} Model;

static int match(Model *m, int seed, int cand)
{
  const uint64_t key = ((uint64_t)seed << 32) | (uint32_t)cand;
  size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & m->memoMask;

  while (m->memoKey[h] != ~0ull)
  {
    if (m->memoKey[h] == key)
      return m->memoVal[h];

    h = (h + 1) & m->memoMask;
  }

  m->nPred++;
  const int v = lo_matches(m->CH, &m->table[seed], &m->table[cand]);
  m->memoKey[h] = key;
  m->memoVal[h] = (uint8_t)v;
  return v;
}

/* mask at logical time T: live (T == NONE: anything owned) or replay (owner time < T) */
static inline int used_at(const Model *m, int x, int y, uint32_t T)
{
  const uint32_t o = m->owner[(size_t)y * m->BX + x];
  return T == NONE ? (o != NONE) : (o < T);
}

static int strip_ok(Model *m, int seed, int x0, int y0, int w, int h, uint32_t T)
{
  for (int y = y0; y < y0 + h; y++)
    for (int x = x0; x < x0 + w; x++)
      if (used_at(m, x, y, T))
        return 0;

  for (int y = y0; y < y0 + h; y++)
    for (int x = x0; x < x0 + w; x++)
      if (!match(m, seed, y * m->BX + x))
        return 0;

  return 1;
}

static int g_pL, g_pR, g_pU, g_pD; /* probe extents of the last grow() (inclusive block coordinates) */

typedef struct
{
  int rx, ry, kind;            /* kind 0 nothing, 1 right/down rectangle, 2 centre-third regrowth */
  int cox, coy, crx, cry, attempted;
  int boxR;                    /* exclusive right edge of everything probed */
  int boxL;                    /* inclusive left edge */
  int c0x, c0y, eL, eR, eU, eD; /* four-way: start centre and probe extents relative to it */
  int sR, sD;                  /* seed growth probe extents relative to the seed */
} Result;

/* limg.cpp:1294-1388 */
static void grow(Model *m, int *pox, int *poy, int *prx, int *pry, int fourWay, uint32_t T, int *boxL, int *boxR)
{
  int ox = *pox, oy = *poy, rx = *prx, ry = *pry;
  const int seed = oy * m->BX + ox;
  int up = fourWay, down = 1, left = fourWay, right = 1;
  g_pL = ox; g_pR = ox + rx - 1; g_pU = oy; g_pD = oy + ry - 1;

  while (up || down || left || right)
  {
    if (right)
    {
      if (ox + rx + 1 < m->BX) { if (ox + rx + 1 > *boxR) *boxR = ox + rx + 1; }
      if (ox + rx + 1 < m->BX && ox + rx > g_pR) g_pR = ox + rx;
      if (ox + rx + 1 < m->BX && strip_ok(m, seed, ox + rx, oy, 1, ry, T)) rx++; else right = 0;
    }

    if (down)
    {
      if (oy + ry + 1 < m->BY && oy + ry > g_pD) g_pD = oy + ry;
      if (oy + ry + 1 < m->BY && strip_ok(m, seed, ox, oy + ry, rx, 1, T)) ry++; else down = 0;
    }

    if (up)
    {
      if (oy > 0 && oy - 1 < g_pU) g_pU = oy - 1;
      if (oy > 0 && strip_ok(m, seed, ox, oy - 1, rx, 1, T)) { oy--; ry++; } else up = 0;
    }

    if (left)
    {
      if (ox > 0) { if (ox - 1 < *boxL) *boxL = ox - 1; }
      if (ox > 0 && ox - 1 < g_pL) g_pL = ox - 1;
      if (ox > 0 && strip_ok(m, seed, ox - 1, oy, 1, ry, T)) { ox--; rx++; } else left = 0;
    }
  }

  *pox = ox; *poy = oy; *prx = rx; *pry = ry;
}

/* limg.cpp:1405-1486 for one seed against the mask at time T */
static Result expand(Model *m, int x, int y, int stage, uint32_t T)
{
  Result r;
  memset(&r, 0, sizeof(r));
  int ox = x, oy = y, rx = 1, ry = 1;
  r.boxL = x; r.boxR = x + 1;
  grow(m, &ox, &oy, &rx, &ry, 0, T, &r.boxL, &r.boxR);
  r.rx = rx; r.ry = ry;
  r.sR = g_pR - x; r.sD = g_pD - y;

  if (stage == 0)
  {
    if (rx >= 3 && ry >= 3)
    {
      int cox = x + rx / 3, coy = y + ry / 3, crx = rx / 3, cry = ry / 3;
      r.c0x = cox; r.c0y = coy;
      grow(m, &cox, &coy, &crx, &cry, 1, T, &r.boxL, &r.boxR);
      r.eL = r.c0x - g_pL; r.eR = g_pR - r.c0x; r.eU = r.c0y - g_pU; r.eD = g_pD - r.c0y;
      r.cox = cox; r.coy = coy; r.crx = crx; r.cry = cry; r.attempted = 1;
      r.kind = crx * cry > rx * ry ? 2 : 1;
    }
  }
  else
  {
    r.kind = (rx > 1 || ry > 1) ? 1 : 0;
  }

  return r;
}

typedef struct { uint32_t T; int ox, oy, rx, ry; } Emit;

typedef struct
{
  int y, x, k, state;
  double t;
  Result pend; int pendNeed; uint32_t pendT;
  int active;
} Worker;

enum { S_START, S_NEXT, S_TRY, S_COMMIT, S_FINISHED };

typedef struct
{
  double rowStart, poll, skip, exp, four, commit, next;
} Costs;

typedef struct
{
  uint64_t expansions, reexpansions, polls, conflicts, emitted, verifyFail, refMismatch;
  double makespan;
  uint64_t leftExtHist[16]; /* how far left of its seed column a centre-third rectangle reaches */
  uint64_t boxWHist[16];
  uint64_t nDecide, nBlocked;
  uint64_t fourN, centreMiss, eLH[20], eRH[20], eUH[20], eDH[20], sRH[20], sDH[20];
} Stats;

static int is_cand(Model *m, int x, int y, int stage)
{
  const int s = y * m->BX + x;

  if (stage == 0)
  {
    if (x + 2 >= m->BX || y + 2 >= m->BY)
      return 0;

    for (int dy = 0; dy < 3; dy++)
      for (int dx = 0; dx < 3; dx++)
        if ((dx || dy) && !match(m, s, (y + dy) * m->BX + x + dx))
          return 0;

    return 1;
  }

  return (x + 1 < m->BX && match(m, s, s + 1)) || (y + 1 < m->BY && match(m, s, s + m->BX));
}

static int cmp_emit(const void *a, const void *b)
{
  const Emit *x = (const Emit *)a, *y = (const Emit *)b;
  return x->T < y->T ? -1 : (x->T > y->T ? 1 : 0);
}

/* returns number of emitted rectangles (sorted by logical time) */
size_t model_run(const lo_decomp *table, int BX, int BY, int CH, int workers, int margin, int sequential, const double *costs7,
                 lo_area *outAreas, double *outStats /* [64] */)
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

  Costs c = { costs7[0], costs7[1], costs7[2], costs7[3], costs7[4], costs7[5], costs7[6] };
  Emit *emits = (Emit *)malloc(sizeof(Emit) * (size_t)BX * BY);
  size_t nEmit = 0;
  uint8_t *cand = (uint8_t *)malloc((size_t)BX * BY);
  uint8_t *unmRx = (uint8_t *)malloc((size_t)BX * BY);
  int *progress = (int *)malloc(sizeof(int) * (BY + 1));
  Worker *w = (Worker *)calloc(workers, sizeof(Worker));
  Stats st[2];
  memset(st, 0, sizeof(st));
  double t0 = 0;

  for (int stage = 0; stage < 2; stage++)
  {
    const uint32_t base = stage == 0 ? 0u : 0x40000000u;

    for (int y = 0; y < BY; y++)
      for (int x = 0; x < BX; x++)
      {
        cand[y * BX + x] = (uint8_t)is_cand(&m, x, y, stage);
        /* probe-box width of the mask-free expansion (what k_plan_seeds predicts) */
        int bw = 1;
        if (cand[y * BX + x]) { const Result r0 = expand(&m, x, y, stage, 0u); bw = r0.boxR - x; }
        if (costs7[7] == 0.0 || (costs7[7] == 2.0 && stage == 1)) bw = 2;
        unmRx[y * BX + x] = (uint8_t)(bw > 250 ? 250 : bw);
      }

    for (int y = 0; y < BY; y++) progress[y] = 0;
    int ticket = 0;

    for (int i = 0; i < workers; i++) { w[i].state = S_START; w[i].t = t0; w[i].active = 1; }

    int running = workers;
    double tEnd = t0;

    while (running > 0)
    {
      int wi = -1;

      for (int i = 0; i < workers; i++)
        if (w[i].active && (wi < 0 || w[i].t < w[wi].t))
          wi = i;

      Worker *k = &w[wi];

      switch (k->state)
      {
      case S_START:
        if (ticket >= BY) { k->active = 0; running--; if (k->t > tEnd) tEnd = k->t; break; }
        k->y = ticket++; k->x = 0; k->t += c.rowStart; k->state = S_NEXT;
        break;

      case S_NEXT:
      {
        int x = k->x;
        while (x < BX && !cand[k->y * BX + x]) x++;
        k->t += c.next * (1 + (x - k->x) / 32);

        if (x >= BX) { progress[k->y] = DONE; k->state = S_START; break; }

        progress[k->y] = x;
        k->x = x; k->k = 0; k->state = S_TRY;
        break;
      }

      case S_TRY:
      {
        int p = DONE;
        for (int rr = 0; rr < k->y; rr++) if (progress[rr] < p) p = progress[rr]; /* monotone relay */

        if (sequential ? (p != DONE) : (p != DONE && p < k->x + (int)unmRx[k->y * BX + k->x] + margin))
        {
          k->t += c.poll; st[stage].polls++;
          break;
        }

        if (used_at(&m, k->x, k->y, NONE)) { k->t += c.skip; k->x++; k->state = S_NEXT; break; }

        const uint32_t T = base + (((uint32_t)(k->y * BX + k->x)) << 3) + (uint32_t)k->k;
        k->pend = expand(&m, k->x, k->y, stage, NONE);
        st[stage].expansions++;
        { static uint32_t lcg = 12345u; lcg = lcg * 1664525u + 1013904223u; const double j = 1.0 + costs7[10] * (((lcg >> 8) & 0xFFFF) / 32768.0 - 1.0); k->t += j * (c.exp + (k->pend.attempted ? c.four : 0)); }
        int need = k->pend.boxR + margin;
        {
          /* proven safe: every row of the probed boxes (one more above and below) has an in-use block just right of them */
          const Result *r = &k->pend;
          int top = k->y, bottom = k->y + r->ry;
          if (r->attempted) { if (r->coy - 1 < top) top = r->coy - 1; if (r->coy + r->cry > bottom) bottom = r->coy + r->cry; }
          int blocked = 1;
          if (costs7[9] > 0 && bottom + 1 > k->y + (int)costs7[9]) bottom = k->y + (int)costs7[9] - 1;
          for (int rr = top - 1; rr <= bottom + 1 && blocked; rr++)
          {
            if (rr < 0 || rr >= BY) continue;
            int any = r->boxR >= BX;
            for (int cc = r->boxR; cc < r->boxR + margin && cc < BX && !any; cc++) any = used_at(&m, cc, rr, NONE);
            if (!any && costs7[11] > 0)
            {
              /* no in-use block: only dangerous if the blocks there form a run of matches (a flat region) */
              int run = 1;
              const int b0 = rr * BX + r->boxR;
              while (run < 8 && r->boxR + run < BX && match(&m, b0, b0 + run)) run++;
              if (run < 8) any = 1;
            }
            blocked = any;
          }
          st[stage].nDecide++; st[stage].nBlocked += blocked;
          if (!blocked) need = k->pend.boxR + (int)costs7[8];
        }

        if (p != DONE && p < need)
        {
          /* wait for the row above, then expand again */
          unmRx[k->y * BX + k->x] = (uint8_t)(k->pend.boxR - k->x > 250 ? 250 : k->pend.boxR - k->x);
          st[stage].reexpansions++;
          break;
        }

        k->pendT = T;
        k->state = S_COMMIT;
        k->t += c.commit;
        break;
      }

      case S_COMMIT:
      {
        const Result *r = &k->pend;
        { int v; v = r->sR > 19 ? 19 : r->sR; st[stage].sRH[v]++; v = r->sD > 19 ? 19 : r->sD; st[stage].sDH[v]++; }
        if (r->attempted)
        {
          int v;
          st[stage].fourN++;
          Result f; f.attempted = 1; f.c0x = r->c0x; f.c0y = r->c0y;

          v = r->eL > 19 ? 19 : r->eL; st[stage].eLH[v]++; v = r->eR > 19 ? 19 : r->eR; st[stage].eRH[v]++;
          v = r->eU > 19 ? 19 : r->eU; st[stage].eUH[v]++; v = r->eD > 19 ? 19 : r->eD; st[stage].eDH[v]++;
        }

        if (r->kind != 0)
        {
          const int ox = r->kind == 2 ? r->cox : k->x, oy = r->kind == 2 ? r->coy : k->y, rx = r->kind == 2 ? r->crx : r->rx, ry = r->kind == 2 ? r->cry : r->ry;

          for (int yy = oy; yy < oy + ry; yy++)
            for (int xx = ox; xx < ox + rx; xx++)
            {
              if (m.owner[yy * BX + xx] != NONE) st[stage].conflicts++;
              else m.owner[yy * BX + xx] = k->pendT;
            }

          emits[nEmit].T = k->pendT; emits[nEmit].ox = ox; emits[nEmit].oy = oy; emits[nEmit].rx = rx; emits[nEmit].ry = ry;
          nEmit++;
          st[stage].emitted++;

          if (r->kind == 2)
          {
            int le = k->x - ox; if (le < 0) le = 0; if (le > 15) le = 15;
            st[stage].leftExtHist[le]++;
          }

          int bw = r->boxR - k->x; if (bw > 15) bw = 15;
          st[stage].boxWHist[bw]++;
        }

        if (r->kind == 2 && !used_at(&m, k->x, k->y, NONE)) { k->k++; k->state = S_TRY; }
        else { k->x = r->kind == 1 ? k->x + r->rx : k->x + 1; k->state = S_NEXT; }

        break;
      }
      }
    }

    st[stage].makespan = tEnd - t0;
    t0 = tEnd;

    /* verification: replay every candidate against the mask "owner time < T" */
    qsort(emits, nEmit, sizeof(Emit), cmp_emit);

    for (int y = 0; y < BY; y++)
      for (int x = 0; x < BX; x++)
      {
        if (!cand[y * BX + x])
          continue;

        const uint32_t T0 = base + (((uint32_t)(y * BX + x)) << 3);
        /* recorded emissions of this seed */
        size_t lo = 0, hi = nEmit;
        while (lo < hi) { size_t mid = (lo + hi) / 2; if (emits[mid].T < T0) lo = mid + 1; else hi = mid; }
        size_t e = lo;
        int ok = 1;

        for (int kk = 0; kk < 8 && ok; kk++)
        {
          const uint32_t T = T0 + kk;
          const int have = e < nEmit && emits[e].T == T;

          if (used_at(&m, x, y, T)) { ok = !have; break; }

          Result r = expand(&m, x, y, stage, T);

          if (r.kind == 0) { ok = !have; break; }

          const int ox = r.kind == 2 ? r.cox : x, oy = r.kind == 2 ? r.coy : y, rx = r.kind == 2 ? r.crx : r.rx, ry = r.kind == 2 ? r.cry : r.ry;
          ok = have && emits[e].ox == ox && emits[e].oy == oy && emits[e].rx == rx && emits[e].ry == ry;
          e++;

          if (r.kind == 1) break;
        }

        if (ok && e < nEmit && emits[e].T >= T0 && emits[e].T < T0 + 8) ok = 0;

        if (!ok) st[stage].verifyFail++;
      }
  }

  for (size_t i = 0; i < nEmit; i++)
  {
    outAreas[i].ox = emits[i].ox; outAreas[i].oy = emits[i].oy; outAreas[i].rx = emits[i].rx; outAreas[i].ry = emits[i].ry;
    outAreas[i].stage = emits[i].T >= 0x40000000u;
  }

  for (int s = 0; s < 2; s++)
  {
    double *o = outStats + s * 32;
    o[0] = (double)st[s].expansions; o[1] = (double)st[s].reexpansions; o[2] = (double)st[s].polls; o[3] = (double)st[s].conflicts;
    o[4] = (double)st[s].emitted; o[5] = (double)st[s].verifyFail; o[6] = st[s].makespan; o[7] = (double)m.nPred;
    for (int i = 0; i < 12; i++) o[8 + i] = (double)st[s].leftExtHist[i < 11 ? i : 15];
    for (int i = 0; i < 12; i++) o[20 + i] = (double)st[s].boxWHist[i < 11 ? i : 15];
  }

  for (int s2 = 0; s2 < 2; s2++)
  {
    printf("  stage %d: four-way attempts %llu, centre != mask-free centre %llu; decisions %llu of which proven safe %llu\n", s2, (unsigned long long)st[s2].fourN, (unsigned long long)st[s2].centreMiss, (unsigned long long)st[s2].nDecide, (unsigned long long)st[s2].nBlocked);
    const char *names[6] = { "4way left ", "4way right", "4way up   ", "4way down ", "seed right", "seed down " };
    uint64_t *hs[6] = { st[s2].eLH, st[s2].eRH, st[s2].eUH, st[s2].eDH, st[s2].sRH, st[s2].sDH };
    for (int h = 0; h < 6; h++) { printf("    %s:", names[h]); for (int i = 0; i < 20; i++) printf(" %llu", (unsigned long long)hs[h][i]); printf("\n"); }
  }
  fflush(stdout);
  free(m.memoKey); free(m.memoVal); free(m.owner); free(emits); free(cand); free(unmRx); free(progress); free(w);
  return nEmit;
}

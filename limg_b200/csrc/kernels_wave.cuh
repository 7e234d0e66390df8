// limg_b200/csrc/kernels_wave.cuh -- the area-expansion scan as a row-pipelined wavefront with an exact verification pass
// (limg.cpp:1121-1135, 1277-1496, drivers 1814-1858).
//
// The reference's merge is a serial greedy raster scan; the only state it carries from seed to seed is the in-use mask, and the
// predicate "candidate block joins seed" is a pure function of two pass-1 records. Every rectangle the scan emits
// gets a LOGICAL TIME T = (seed raster index, attempt number); the sequential result is the unique record in which every seed's
// decision equals what limg_encode_find_block_3d does against the mask { blocks owned by a rectangle with time < T }.
//
//   wave_scan_rows  one warp per block row (rows handed out by ticket, so a row's predecessor is always running). A row walks
//                   its candidate seeds left to right against the LIVE mask. Before a seed's decision stands, the row above
//                   must have committed every seed left of (right edge of everything the seed probed + margin); by induction
//                   the rows further up are further ahead. That lag rule is a heuristic: the four-way centre-third regrowth
//                   can reach arbitrarily far to the left, so the record is verified (below). The mask, the rows' progress
//                   words and the tickets live in a BACK END: k_merge_cta (kernels_cta.cuh) keeps them in the shared memory
//                   of one thread-block cluster (every CTA holds a replica it reads locally; writes go to all replicas), which
//                   is what the encoder uses; k_merge_wave keeps them in global memory (L2): any image size, any grid, and
//                   the strictly sequential last resort.
//   k_merge_verify  replays EVERY candidate seed, fully in parallel, against the mask "owner time < my time" built from the
//                   per-block owner times the wave wrote, and compares with what the wave recorded. All equal (and no
//                   rectangle overlap, detected by the atomicOr of the claims)  =>  the record is self-consistent  =>  it is
//                   the sequential scan's result (induction over T).
//   fallback        if verification fails, the stage is reset and the same kernel re-runs with rows strictly in sequence
//                   (row y starts when row y-1 is done): that IS the reference's order, no verification needed.
//
// Stage 1 (remaining merges, limg.cpp:1838-1858) repeats the procedure on top of the final stage-0 mask.
#pragma once

#include "kernels_merge.cuh"

namespace limg
{

// LIMG_WAVE_PROFILE=1 compiles the cycle counters and detailed diagnostics of the scan in (tools/phase_times.py, tools/row_times.py
// print them); they cost ~30 registers per thread in a kernel that is already at the limit, so the default build leaves them out.
#ifndef LIMG_WAVE_PROFILE
#define LIMG_WAVE_PROFILE 0
#endif

__device__ __forceinline__ long long wave_clock()
{
#if LIMG_WAVE_PROFILE
  return clock64();
#else
  return 0;
#endif
}

#define LIMG_WAVE_DONE 0x7FFFFFFF
#define LIMG_TAU_NONE 0xFFFFFFFFu
#define LIMG_TAU_STAGE1 0x40000000u
#define LIMG_WAVE_WARPS 4
#define LIMG_WAVE_MAX_ATTEMPTS 8

struct WaveArgs
{
  const PredRec *rec;
  const uint32_t *window;
  const uint32_t *extSlot, *extBits, *extHdr;
  const uint32_t *symSlot, *symBits, *symHdr;
  const uint16_t *unmasked;
  const uint4 *seedSym;      // per stage-0 candidate: bitmap slots of its predicted centre and the three to its left (nullptr: look them up)
  const uint32_t *safe;      // [blocks] safe columns of stage 0 (k_plan_safe): low 16 bits while the candidate at the position is undecided, high 16 bits once it is decided; nullptr: none
  const uint32_t *candBits;  // [2][BY][wordsPerRow]: mask-free necessary condition for a seed to emit in stage 0 / 1
  const uint32_t *candList;  // [2][blocks]
  const uint32_t *candCount; // [2]
  int BX, BY, wordsPerRow;
  uint32_t *used;            // [BY][wordsPerRow] live in-use bits (rows padded by two zero words)
  uint32_t *tau;             // [blocks] logical time of the rectangle that owns the block
  int *progress;             // [2][BY] column up to which the row's seeds are committed (LIMG_WAVE_DONE when finished)
  uint32_t *ticket;          // counters[4]: next stage-0 row, done stage-0 rows, next stage-1 row, done stage-1 rows
  uint2 *rowLists;           // [2][BY][listCap] (ox | oy << 16, rx | ry << 16) in emission order
  uint32_t *rowCounts;       // [2][BY]
  uint32_t *emitInfo;        // [2][blocks] per seed: first entry in its row list << 8 | number of entries
  uint32_t *flags;           // [0] number of failed tries, [1], [2] a stage failed the verification of the current try, [3] watchdog (hard error), [4] list overflow (hard error)
  uint32_t *stats;           // [16] optional counters
  uint32_t *dbg;             // [256] detailed counters (see tools/phase_times.py)
  int eventRow;              // diagnostics: decisions of stage-0 rows eventRow .. eventRow + 3 go to dbgRows + 8 * BY as [4][64](x | kind << 16 | waited << 20, ns)
  uint32_t *dbgRows;         // [2][BY][4] optional per-row time stamps (globaltimer ns, low 32 bits): ticket, first decision, last decision, done
  int listCap, margin;
  int symMaxL, symMaxR, symMaxD; // size caps of a centre bitmap built on the fly (same as the plan's)
  int stageGap;              // block rows stage 1 stays behind stage 0
  int specAhead;             // a seed is expanded speculatively once the rows above are within this many columns of where they have to be
  int experiment;            // LIMGCU_SCAN_EXPERIMENT (measurement only, never set by default): 1 no acquire fence before the final look, 2 claims fenced at CTA scope only
};

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ int ld_acquire_s32(const int *p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Polling uses relaxed (strong, L2-coherent) loads: an acquire load costs an L1 invalidation (CCTL.IVALL) plus a fence on every
// poll. The producer fences between its claims and its progress store; the consumer fences (acquire_fence() of the back end)
// between a progress read and the look at the mask that can make a decision final, and only then.
__device__ __forceinline__ int ld_relaxed_s32(const int *p)
{
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_relaxed_s32(int *p, int v)
{
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void fence_acq_rel_gpu()
{
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

#define LIMG_SNAP_UP 8 // a seed's mask snapshot covers block rows y - 8 .. y + 23, one per lane

// 96 in-use bits per lane: row (r0 + lane), columns [32 * w0, 32 * w0 + 96)
struct Snapshot
{
  uint32_t w[3];
  int r0, w0;
};

// the live in-use mask: [BY][wordsPerRow] words behind a loader (global memory read at L2, or shared memory)
template <class Words>
struct WordMask
{
  Words ld; // ld(i): word i of the mask, a strong (never cached stale) read
  int wordsPerRow, BX, BY;

  // 32 in-use bits of row y starting at column x (x may be negative); everything outside the grid rows reads as in use
  __device__ __forceinline__ uint32_t bits(int x, int y) const
  {
    if (y < 0 || y >= BY)
      return 0xFFFFFFFFu;

    const int row = y * wordsPerRow;

    if (x < 0)
    {
      const int s = -x; // 1..31
      return (ld(row) << s) | ((1u << s) - 1u);
    }

    const int w0 = x >> 5, s = x & 31;
    const uint32_t lo = ld(row + w0);
    const uint32_t hi = s ? ld(row + w0 + 1) : 0u; // rows are padded; columns >= BX are never set
    return __funnelshift_r(lo, hi, s);
  }

  __device__ __forceinline__ void words(int y, int w0, uint32_t out[3]) const
  {
    if (y < 0 || y >= BY)
    {
      out[0] = out[1] = out[2] = 0xFFFFFFFFu;
      return;
    }

    const int row = y * wordsPerRow + w0;
    out[0] = ld(row);
    out[1] = ld(row + 1);
    out[2] = ld(row + 2);
  }

  __device__ __forceinline__ bool is_used(int x, int y) const
  {
    return (ld(y * wordsPerRow + (x >> 5)) >> (x & 31)) & 1u;
  }
};

struct GlobalWords
{
  const uint32_t *used;
  __device__ __forceinline__ uint32_t operator()(int i) const { return ld_relaxed_u32(used + i); }
};

typedef WordMask<GlobalWords> LiveMask;

// the mask the sequential scan shows a seed at logical time T: blocks owned by an earlier rectangle
struct TimeMask
{
  const uint32_t *tau;
  int BX, BY;
  uint32_t T;

  __device__ __forceinline__ uint32_t bits(int x, int y) const
  {
    if (y < 0 || y >= BY)
      return 0xFFFFFFFFu;

    const uint32_t *row = tau + (size_t)y * BX;
    uint32_t b = 0;

    if (x >= 0 && ((x | BX) & 3) == 0)
    {
      // aligned: four owner times per load (the verification takes a 32 x 96 snapshot of them per replayed seed)
      const uint4 *row4 = reinterpret_cast<const uint4 *>(row + x);

#pragma unroll
      for (int i = 0; i < 8; i++)
      {
        if (x + 4 * i < BX) // BX is a multiple of 4: the whole quad is inside the row
        {
          const uint4 v = __ldg(row4 + i);
          b |= ((v.x < T ? 1u : 0u) | (v.y < T ? 2u : 0u) | (v.z < T ? 4u : 0u) | (v.w < T ? 8u : 0u)) << (4 * i);
        }
      }

      return b;
    }

    for (int i = 0; i < 32; i++)
    {
      const int xx = x + i;
      const bool u = xx < 0 ? true : (xx < BX && __ldg(row + xx) < T);
      b |= (u ? 1u : 0u) << i;
    }

    return b;
  }

  __device__ __forceinline__ void words(int y, int w0, uint32_t out[3]) const
  {
#pragma unroll
    for (int j = 0; j < 3; j++)
      out[j] = bits((w0 + j) * 32, y);
  }

  __device__ __forceinline__ bool is_used(int x, int y) const
  {
    return __ldg(tau + (size_t)y * BX + x) < T;
  }
};

// what one seed does against a given mask
struct WaveResult
{
  int rx, ry;             // right/down rectangle grown from the seed
  int kind;               // 0 nothing to emit, 1 emit the right/down rectangle, 2 emit the centre-third regrowth (and examine the seed again)
  int cox, coy, crx, cry; // four-way regrowth (valid when attempted)
  int attempted;          // the centre-third regrowth ran
  int boxR;               // exclusive right edge of every column whose in-use bits were consulted
  int boxD;               // exclusive bottom edge of the rectangles it would emit
};

// first level of a seed's immutable data: what is indexed by the seed itself. Loaded one candidate ahead (the loads are in flight
// while the row decides the seed before); the second level (bitmaps behind the slots) is SeedPre.
struct SeedLinks
{
  int x;          // the candidate the loads were issued for (-1: none)
  uint32_t slot;  // extSlot
  uint32_t u;     // unmasked rx | ry << 8
  uint32_t w0, w1; // the seed's 8x8 match word (warp uniform)
  uint4 links;    // bitmap slots of the predicted centre and the three to its left
  uint32_t safe;  // safe columns of the candidate's position (stage 0; WaveArgs::safe)
};

// everything about a seed that does not depend on the mask: loaded while the row waits for the rows above
struct SeedPre
{
  uint32_t w0, w1;  // the seed's 8x8 match word: bit dy * 8 + dx (warp uniform)
  uint32_t rowBits; // lane: match bits of the seed for block row y + lane, columns x ..
  int vx1, vy1;     // known part of it: [0, vx1) x [0, vy1)
  int pcx, pcy;     // centre the mask-free growth predicts (-1: none)
  uint32_t symRow;  // lane: match bits of that centre for block row pcy - 8 + lane, columns pcx - 8 ..
  uint32_t symHdr;  // known part (0: no bitmap)
  uint32_t alt[3];  // bitmap slots of the centres one, two and three blocks left of the predicted one
  // a bitmap built on the fly for a centre that had none (kept for the seed's next expansions)
  int ccx, ccy;
  uint32_t cRow, cHdr;
};

// a match bitmap combined with the mask: bit (row ay + lane, column ax + i) = matches the seed and is free
struct Region
{
  int ax, ay;
  int vx0, vy0, vx1, vy1; // known part (absolute block coordinates, half open)
  uint32_t avail;
};

// Every method is warp-cooperative: all 32 lanes call it with warp-uniform arguments.
template <int CH, class Mask>
struct WaveScan
{
  const WaveArgs &a;
  Mask mask;
  int lane;
  uint32_t nOnDemand;
  long long tOnDemand = 0, tFour = 0; // profile: clock cycles inside on-demand strips / the four-way regrowth
  uint32_t *scratch = nullptr;        // 32 words of shared memory private to the warp
  Snapshot cur = { { 0, 0, 0 }, 0, 0 }; // the mask snapshot the running expansion is based on (a copy: a pointer would put every snapshot into local memory)
  bool haveCur = false;
  bool volatileReads = false;         // the running expansion read in-use bits that the snapshot does not hold
  int cause = 0;                      // who asks for on-demand strips: 0 seed growth, 1 four-way with a bitmap, 2 four-way without
  uint32_t nStrips[3] = { 0, 0, 0 };
  long long tStrips[3] = { 0, 0, 0 };
  uint32_t nFour = 0, nFourMiss = 0, nFourNoSym = 0, nBuilt = 0;
  long long tSeedFast = 0, tSeedGeneral = 0, tSym = 0, tRegion = 0, tGrow4 = 0; // profile: where an expansion's cycles go
  uint32_t nSeedFast = 0, nSeedGeneral = 0;

  __device__ bool strip_unused(int x0, int y0, int w, int h)
  {
    bool any = false;

    if (haveCur && y0 >= cur.r0 && y0 + h <= cur.r0 + 32 && x0 >= cur.w0 * 32 && x0 + w <= cur.w0 * 32 + 96)
    {
      // inside the snapshot: lane = row, test the columns of the strip
      const int row = cur.r0 + lane, b0 = x0 - cur.w0 * 32;

      if (row >= y0 && row < y0 + h)
      {
        for (int j = 0; j < 3; j++)
        {
          const int lo = max(b0 - 32 * j, 0), hi = min(b0 + w - 32 * j, 32);

          if (hi > lo)
          {
            const uint32_t m = (hi - lo >= 32 ? 0xFFFFFFFFu : ((1u << (hi - lo)) - 1u)) << lo;
            any |= (cur.w[j] & m) != 0;
          }
        }
      }

      return !__any_sync(0xFFFFFFFFu, any);
    }

    for (int e = lane; e < w * h; e += 32)
    {
      const int yy = y0 + e / w, xx = x0 + e % w;
      any |= mask.is_used(xx, yy);
    }

    return !__any_sync(0xFFFFFFFFu, any);
  }

  // every block of the strip matches the seed? short strips one predicate at a time with the 27 samples spread over the lanes,
  // long strips one predicate per lane.
  __device__ bool strip_matches(const PredRec &seed, int x0, int y0, int w, int h)
  {
    const int count = w * h;
    nOnDemand += count;

    if (count >= 6)
    {
      bool ok = true;

      for (int base = 0; base < count && ok; base += 32)
      {
        const int e = base + lane;
        bool m = true;

        if (e < count)
        {
          const int yy = y0 + e / w, xx = x0 + e % w;
          m = predicate_thread<CH>(seed, a.rec[(size_t)yy * a.BX + xx]);
        }

        ok = __all_sync(0xFFFFFFFFu, m);
      }

      return ok;
    }

    for (int e = 0; e < count; e++)
    {
      const int yy = y0 + e / w, xx = x0 + e % w;

      if (!predicate_warp<CH>(seed, a.rec[(size_t)yy * a.BX + xx]))
        return false;
    }

    return true;
  }

  // in-use bit of one block for THIS lane's own cell (lanes ask for different blocks): the current snapshot is indexed by lane = row,
  // so a lane cannot take another row's word from its registers; cells read the live mask (monotone: never clears)
  __device__ __forceinline__ bool cell_used(int xx, int yy) const
  {
    return mask.is_used(xx, yy);
  }

  // When only ONE direction of a growth is still alive the strips that follow are known in advance (same length, next row /
  // column): evaluate as many of them as fit into the 32 lanes at once instead of one after the other. Returns how many
  // consecutive strips join (0 .. n). dirX, dirY = step from strip to strip; (x0, y0, w, h) = the first strip.
  __device__ int strips_run(const PredRec &seed, int x0, int y0, int w, int h, int dirX, int dirY, int n)
  {
    const long long t0 = wave_clock();
    const int cells = w * h;
    const int j = lane / cells, e = lane - j * cells; // strip index, cell inside the strip
    bool ok = true;

    if (j < n)
    {
      const int xx = x0 + j * dirX + e % w, yy = y0 + j * dirY + e / w;
      ok = !cell_used(xx, yy) && predicate_thread<CH>(seed, a.rec[(size_t)yy * a.BX + xx]);
    }

    nOnDemand += (uint32_t)(n * cells);
    // first strip with a failing cell
    const uint32_t bad = __ballot_sync(0xFFFFFFFFu, !ok);
    int good = n;

    if (bad)
      good = min(n, (__ffs(bad) - 1) / cells);

    const long long dt = wave_clock() - t0;
    tOnDemand += dt;
    if (LIMG_WAVE_PROFILE) { nStrips[cause]++; tStrips[cause] += dt; }
    return good;
  }

  __device__ bool strip_joins(const PredRec &seed, int x0, int y0, int w, int h)
  {
    const long long t0 = wave_clock();
    const bool ok = strip_unused(x0, y0, w, h) && strip_matches(seed, x0, y0, w, h);
    const long long dt = wave_clock() - t0;
    tOnDemand += dt;
    if (LIMG_WAVE_PROFILE) { nStrips[cause]++; tStrips[cause] += dt; }
    return ok;
  }

  __device__ __forceinline__ Snapshot snapshot(int x, int y) const
  {
    Snapshot s;
    s.r0 = y - LIMG_SNAP_UP;
    s.w0 = max(x - 8, 0) >> 5;
    mask.words(s.r0 + lane, s.w0, s.w);
    return s;
  }

  // is block (x, y) of the snapshot's seed in use? (x - 32 * w0 < 40)
  __device__ __forceinline__ bool snap_used(const Snapshot &s, int x) const
  {
    const int b = x - s.w0 * 32;
    const uint32_t w = __shfl_sync(0xFFFFFFFFu, b < 32 ? s.w[0] : s.w[1], LIMG_SNAP_UP);
    return (w >> (b & 31)) & 1u;
  }

  // in-use bits of row (ay + lane), columns [ax, ax + 32): from the snapshot where it covers them, else read directly
  __device__ uint32_t region_used(const Snapshot &s, int ax, int ay, int rows) const
  {
    const int src = ay + lane - s.r0, sh = ax - s.w0 * 32;
    const int from = min(max(src, 0), 31);
    const uint32_t a0 = __shfl_sync(0xFFFFFFFFu, s.w[0], from), a1 = __shfl_sync(0xFFFFFFFFu, s.w[1], from), a2 = __shfl_sync(0xFFFFFFFFu, s.w[2], from);

    if (lane >= rows)
      return 0xFFFFFFFFu;

    if (src < 0 || src > 31 || sh < 0 || sh > 63)
      return mask.bits(ax, ay + lane); // outside the snapshot (expand() notes whether the growth consulted such bits)

    return sh < 32 ? __funnelshift_r(a0, a1, sh) : __funnelshift_r(a1, a2, sh - 32);
  }

  // do two snapshots of the same seed agree on every in-use bit the expansion `r` of seed (x, y) consulted? (the strips it tested
  // lie inside the seed rectangle's box and, if the regrowth ran, the regrowth rectangle's box, one block wider on each side)
  __device__ bool same_where_probed(const Snapshot &p, const Snapshot &q, const WaveResult &r, int x, int y) const
  {
    const int row = p.r0 + lane, c0 = p.w0 * 32;
    uint32_t m[3] = { 0, 0, 0 };

    auto add = [&](int bx0, int by0, int bx1, int by1) {
      if (row < by0 || row >= by1)
        return;

#pragma unroll
      for (int j = 0; j < 3; j++)
      {
        const int lo = max(bx0 - c0 - 32 * j, 0), hi = min(bx1 - c0 - 32 * j, 32);

        if (hi > lo)
          m[j] |= (hi - lo >= 32 ? 0xFFFFFFFFu : ((1u << (hi - lo)) - 1u)) << lo;
      }
    };

    // Only the blocks of the rectangles themselves count: a strip that did NOT join (a mismatch, or a block already in use) can never join later,
    // because in-use bits are only ever set; a strip that joined stays joined exactly as long as its blocks stay free.
    add(x, y, x + r.rx, y + r.ry);

    if (r.attempted)
      add(r.cox, r.coy, r.cox + r.crx, r.coy + r.cry);

    const bool same = ((p.w[0] ^ q.w[0]) & m[0]) == 0 && ((p.w[1] ^ q.w[1]) & m[1]) == 0 && ((p.w[2] ^ q.w[2]) & m[2]) == 0;
    return __all_sync(0xFFFFFFFFu, same);
  }

  __device__ __forceinline__ void load_sym_slot(uint32_t slot, uint32_t &row, uint32_t &hdr) const
  {
    row = 0;
    hdr = 0;

    if (slot < LIMG_SLOT_PENDING)
    {
      row = ld_relaxed_u32(&a.symBits[(size_t)slot * 32 + lane]);
      hdr = ld_relaxed_u32(&a.symHdr[slot]);
    }
  }

  __device__ __forceinline__ void load_sym(int cx, int cy, uint32_t &row, uint32_t &hdr) const
  {
    // the plan kernels may still be running on the other stream: slots appear while the scan runs
    const uint32_t slot = ld_relaxed_u32(&a.symSlot[(size_t)cy * a.BX + cx]);
    row = 0;
    hdr = 0;

    if (slot < LIMG_SLOT_PENDING)
    {
      row = ld_relaxed_u32(&a.symBits[(size_t)slot * 32 + lane]);
      hdr = ld_relaxed_u32(&a.symHdr[slot]);
    }
  }

  // first level: issue the loads, consume nothing
  __device__ __forceinline__ SeedLinks prefetch_links(int x, int y, int stage) const
  {
    SeedLinks l;
    const int seed = y * a.BX + x;
    l.x = x;
    l.slot = ld_relaxed_u32(&a.extSlot[seed]);
    l.u = *(const volatile uint16_t *)&a.unmasked[seed];
    const uint2 w = __ldg(reinterpret_cast<const uint2 *>(a.window) + seed);
    l.w0 = w.x;
    l.w1 = w.y;
    l.links = (a.seedSym && stage == 0) ? __ldg(&a.seedSym[seed]) : make_uint4(LIMG_NO_SLOT, LIMG_NO_SLOT, LIMG_NO_SLOT, LIMG_NO_SLOT);
    l.safe = (a.safe && stage == 0) ? __ldg(&a.safe[seed]) : 0xFFFFFFFFu;
    return l;
  }

  // second level: the bitmaps behind the slots (consumes the first level)
  __device__ SeedPre prefetch_bitmaps(const SeedLinks &l, int y, int stage) const
  {
    SeedPre p;
    const int x = l.x;
    p.w0 = l.w0;
    p.w1 = l.w1;

    if (l.slot >= LIMG_SLOT_PENDING)
    {
      p.rowBits = lane < 8 ? ((lane < 4 ? l.w0 >> (8 * lane) : l.w1 >> (8 * (lane - 4))) & 0xFFu) : 0u;
      p.vx1 = 8;
      p.vy1 = 8;
    }
    else
    {
      const uint32_t h = ld_relaxed_u32(&a.extHdr[l.slot]);
      p.rowBits = ld_relaxed_u32(&a.extBits[(size_t)l.slot * 32 + lane]);
      p.vx1 = (h >> 16) & 0xFF;
      p.vy1 = h >> 24;
    }

    p.pcx = p.pcy = -1;
    p.symRow = 0;
    p.symHdr = 0;
    p.ccx = p.ccy = -1;
    p.cRow = 0;
    p.cHdr = 0;
    const int prx = l.u & 0xFF, pry = l.u >> 8;

    p.alt[0] = p.alt[1] = p.alt[2] = LIMG_NO_SLOT;

    if (stage == 0 && prx >= 3 && pry >= 3)
    {
      p.pcx = x + prx / 3;
      p.pcy = y + pry / 3;

      if (a.seedSym)
      {
        p.alt[0] = l.links.y; p.alt[1] = l.links.z; p.alt[2] = l.links.w;
        load_sym_slot(l.links.x, p.symRow, p.symHdr);
      }
      else
      {
        load_sym(p.pcx, p.pcy, p.symRow, p.symHdr);
      }
    }

    return p;
  }

  __device__ SeedPre prefetch(int x, int y, int stage) const
  {
    return prefetch_bitmaps(prefetch_links(x, y, stage), y, stage);
  }

  // Right/down growth of a seed decided INSIDE its 8x8 match word (85 % of the stage-0 seeds and 97 % of the stage-1 seeds on photo
  // content, tools/scan_stats.py): the word and the in-use bits of the 8x8 blocks are two 64-bit values every lane holds, so a strip
  // test is a handful of uniform integer operations, no ballot or shuffle. Same sequence of tests as grow() (limg.cpp:1315-1343).
  // Returns false, with nothing decided, if the growth asks for a block outside the word. (A variant that packs one run length per
  // row into nibbles, for the seed growth and the four-way growth alike, was measured and is slower: the packing costs more
  // warp reductions than the short growths of real content save, profiles/README.md r2_f.)
  __device__ __forceinline__ bool grow_in_window(const SeedPre &pre, const Snapshot &sn, int x, int y, int minSide, int &rx, int &ry) const
  {
    // lane 8 + r holds block row y + r of the snapshot (r0 = y - 8)
    const int sh = x - sn.w0 * 32; // < 40
    const uint32_t bits8 = (sh < 32 ? __funnelshift_r(sn.w[0], sn.w[1], sh) : __funnelshift_r(sn.w[1], sn.w[2], sh - 32)) & 0xFFu;
    const uint32_t uLo = __reduce_or_sync(0xFFFFFFFFu, (lane >= 8 && lane < 12) ? bits8 << (8 * (lane - 8)) : 0u);
    const uint32_t uHi = __reduce_or_sync(0xFFFFFFFFu, (lane >= 12 && lane < 16) ? bits8 << (8 * (lane - 12)) : 0u);
    const unsigned long long avail = ((unsigned long long)pre.w0 | ((unsigned long long)pre.w1 << 32)) & ~((unsigned long long)uLo | ((unsigned long long)uHi << 32));
    const unsigned long long column = 0x0101010101010101ull;
    bool right = true, down = true;
    rx = 1;
    ry = 1;

    while (right || down)
    {
      if (right)
      {
        bool ok = x + rx + 1 < a.BX;

        if (ok)
        {
          if (rx >= 8)
            return false;

          const unsigned long long m = (column << rx) & (ry >= 8 ? ~0ull : ((1ull << (8 * ry)) - 1ull));
          ok = (avail & m) == m;
        }

        if (ok) rx++; else right = false;
      }

      if (down)
      {
        bool ok = y + ry + 1 < a.BY;

        if (ok)
        {
          if (ry >= 8)
            return false;

          const unsigned long long m = ((1ull << rx) - 1ull) << (8 * ry);
          ok = (avail & m) == m;
        }

        if (ok) ry++; else down = false;
      }

      if (minSide && ((!right && rx < minSide) || (!down && ry < minSide)))
        break;
    }

    return true;
  }

  // alternating growth of the rectangle (ox, oy, rx, ry) of seed block `seed` (limg.cpp:1294-1388): right and down, and also up
  // and left for the centre-third regrowth. Strips inside the region's known part are bit tests, the others are evaluated on demand.
  __device__ void grow(int seed, const Region &g, bool fourWay, int minSide, int &ox, int &oy, int &rx, int &ry)
  {
    bool haveRec = false;
    PredRec rec;
    bool right = true, down = true, up = fourWay, left = fourWay;

    auto joins = [&](int x0, int y0, int w, int h) -> bool {
      if (x0 >= g.vx0 && y0 >= g.vy0 && x0 + w <= g.vx1 && y0 + h <= g.vy1)
      {
        const uint32_t m = (w >= 32 ? 0xFFFFFFFFu : ((1u << w) - 1u)) << (x0 - g.ax);
        const int r0 = y0 - g.ay;
        const bool rowOk = (lane < r0 || lane >= r0 + h) || ((g.avail & m) == m);
        return __all_sync(0xFFFFFFFFu, rowOk);
      }

      if (LIMG_WAVE_PROFILE && a.dbg && lane == 0)
      {
        // which side of the known part did the strip leave? (per cause: left, up, right, down)
        const int side = x0 < g.vx0 ? 0 : (y0 < g.vy0 ? 1 : (x0 + w > g.vx1 ? 2 : 3));
        atomicAdd(&a.dbg[64 + cause * 4 + side], 1u);

        if (cause == 1 && side >= 2)
          atomicAdd(&a.dbg[80 + min(15, side == 2 ? g.vx1 - g.ax - LIMG_SYM_BACK : g.vy1 - g.ay - LIMG_SYM_BACK)], 1u); // how far the known part reached
      }

      if (!haveRec) { rec = a.rec[seed]; haveRec = true; }
      return strip_joins(rec, x0, y0, w, h);
    };

    // Fast path: while the rectangle and every strip the next iteration can ask for lie inside the known part, an iteration is four
    // bit tests (one ballot / shuffle each) on `avail`, no predicates and no lambdas. The order of the tests inside an iteration is
    // the reference's: right, down, up, left (limg.cpp:1315-1383).
    if (ox >= g.vx0 && oy >= g.vy0 && ox + rx <= g.vx1 && oy + ry <= g.vy1)
    {
      const uint32_t av = g.avail;

      while (right || down || up || left)
      {
        // containment of the strips does not depend on what the earlier directions of the same iteration do
        if ((right && ox + rx >= g.vx1) || (down && oy + ry >= g.vy1) || (up && oy - 1 < g.vy0) || (left && ox - 1 < g.vx0))
          break; // continue with the general loop below (same state)

        if (right)
        {
          bool ok = ox + rx + 1 < a.BX;

          if (ok)
          {
            const uint32_t rows = (ry >= 32 ? 0xFFFFFFFFu : ((1u << ry) - 1u)) << (oy - g.ay);
            ok = (__ballot_sync(0xFFFFFFFFu, (av >> (ox + rx - g.ax)) & 1u) & rows) == rows;
          }

          if (ok) rx++; else right = false;
        }

        if (down)
        {
          bool ok = oy + ry + 1 < a.BY;

          if (ok)
          {
            const uint32_t cols = (rx >= 32 ? 0xFFFFFFFFu : ((1u << rx) - 1u)) << (ox - g.ax);
            ok = (__shfl_sync(0xFFFFFFFFu, av, oy + ry - g.ay) & cols) == cols;
          }

          if (ok) ry++; else down = false;
        }

        if (up)
        {
          bool ok = oy > 0;

          if (ok)
          {
            const uint32_t cols = (rx >= 32 ? 0xFFFFFFFFu : ((1u << rx) - 1u)) << (ox - g.ax);
            ok = (__shfl_sync(0xFFFFFFFFu, av, oy - 1 - g.ay) & cols) == cols;
          }

          if (ok) { oy--; ry++; } else up = false;
        }

        if (left)
        {
          bool ok = ox > 0;

          if (ok)
          {
            const uint32_t rows = (ry >= 32 ? 0xFFFFFFFFu : ((1u << ry) - 1u)) << (oy - g.ay);
            ok = (__ballot_sync(0xFFFFFFFFu, (av >> (ox - 1 - g.ax)) & 1u) & rows) == rows;
          }

          if (ok) { ox--; rx++; } else left = false;
        }

        if (minSide && ((!right && rx < minSide) || (!down && ry < minSide)))
          return;
      }
    }

    // is the strip outside the known part (so that it would be evaluated on demand)?
    auto outside = [&](int x0, int y0, int w, int h) -> bool {
      return !(x0 >= g.vx0 && y0 >= g.vy0 && x0 + w <= g.vx1 && y0 + h <= g.vy1);
    };

    while (right || down || up || left)
    {
      // one direction left and its strips are evaluated on demand: take them in batches (same result, far fewer round trips)
      if (right + down + up + left == 1)
      {
        const int cells = right || left ? ry : rx;

        if (cells <= 16)
        {
          const int x0 = right ? ox + rx : (left ? ox - 1 : ox), y0 = down ? oy + ry : (up ? oy - 1 : oy);
          const int w = right || left ? 1 : rx, h = right || left ? ry : 1;

          if (outside(x0, y0, w, h))
          {
            // how many more strips the grid allows (limg.cpp:1321, 1335: the last block column / row is never entered)
            const int room = right ? a.BX - 1 - (ox + rx) : (down ? a.BY - 1 - (oy + ry) : (up ? oy : ox));
            const int n = min(32 / cells, room);

            if (n <= 0)
              break;

            if (!haveRec) { rec = a.rec[seed]; haveRec = true; }
            const int good = strips_run(rec, x0, y0, w, h, right ? 1 : (left ? -1 : 0), down ? 1 : (up ? -1 : 0), n);

            if (right) rx += good;
            else if (down) ry += good;
            else if (up) { oy -= good; ry += good; }
            else { ox -= good; rx += good; }

            if (good < n)
              break; // the direction is exhausted, and it was the last one

            continue;
          }
        }
      }

      if (right)
      {
        if (ox + rx + 1 < a.BX && joins(ox + rx, oy, 1, ry)) rx++; else right = false;
      }

      if (down)
      {
        if (oy + ry + 1 < a.BY && joins(ox, oy + ry, rx, 1)) ry++; else down = false;
      }

      if (up)
      {
        if (oy > 0 && joins(ox, oy - 1, rx, 1)) { oy--; ry++; } else up = false;
      }

      if (left)
      {
        if (ox > 0 && joins(ox - 1, oy, 1, ry)) { ox--; rx++; } else left = false;
      }

      // stage 0 only keeps rectangles of at least 3 x 3 blocks (limg.cpp:1424): once a side that can no longer grow is shorter,
      // the rest of the growth cannot change the outcome
      if (minSide && ((!right && rx < minSide) || (!down && ry < minSide)))
        break;
    }
  }

  // what seed (x, y) does against the mask (limg.cpp:1405-1486)
  __device__ WaveResult expand(int x, int y, int stage, SeedPre &pre, const Snapshot &sn)
  {
    WaveResult r;
    Region g;
    cur = sn;
    haveCur = true;
    volatileReads = false;
    g.ax = x; g.ay = y; g.vx0 = x; g.vy0 = y; g.vx1 = x + pre.vx1; g.vy1 = y + pre.vy1;

    const long long ts0 = wave_clock();


    if (!grow_in_window(pre, sn, x, y, stage == 0 ? 3 : 0, r.rx, r.ry))
    {
      const long long ts1 = wave_clock();
      g.avail = pre.rowBits & ~region_used(sn, x, y, pre.vy1);
      int ox = x, oy = y;
      r.rx = 1;
      r.ry = 1;
      grow(y * a.BX + x, g, false, stage == 0 ? 3 : 0, ox, oy, r.rx, r.ry);
      tSeedGeneral += wave_clock() - ts1;
      nSeedGeneral++;
    }
    else
    {
      tSeedFast += wave_clock() - ts0;
      nSeedFast++;
    }

    r.kind = 0;
    r.cox = r.coy = r.crx = r.cry = 0;
    r.attempted = 0;
    r.boxD = y + r.ry;
    r.boxR = x + r.rx; // the rows above have to be past the rectangle's blocks, not past the strip that stopped it (see same_where_probed)

    if (stage == 0)
    {
      if (r.rx >= 3 && r.ry >= 3) // Q4
      {
        const long long t0 = wave_clock();
        int cox = x + r.rx / 3, coy = y + r.ry / 3, crx = r.rx / 3, cry = r.ry / 3;
        uint32_t symRow = pre.symRow, symHdr = pre.symHdr;

        nFour++;

        if (cox != pre.pcx || coy != pre.pcy)
        {
          nFourMiss++;

          if (cox == pre.ccx && coy == pre.ccy)
          {
            symRow = pre.cRow;
            symHdr = pre.cHdr;
          }
          else
          {
            const int dLeft = pre.pcx - cox;

            if (a.seedSym && coy == pre.pcy && dLeft >= 1 && dLeft <= 3)
              load_sym_slot(pre.alt[dLeft - 1], symRow, symHdr); // one of the centres the plan linked to this seed
            else
              load_sym(cox, coy, symRow, symHdr);

            if (!symHdr && scratch)
            {
              // no bitmap for this centre: build one now, once, instead of evaluating strip after strip on every expansion
              const long long tb = wave_clock();
              symRow = build_centre_bitmap<CH>(a.rec, a.window, a.BX, a.BY, coy * a.BX + cox, crx, cry, a.symMaxL, a.symMaxR, a.symMaxD, scratch, symHdr);
              pre.ccx = cox; pre.ccy = coy; pre.cRow = symRow; pre.cHdr = symHdr;
              nBuilt++;
              tOnDemand += wave_clock() - tb;
            }
          }
        }
        else if (!symHdr && scratch && pre.pcx >= 0)
        {
          if (pre.ccx == cox && pre.ccy == coy)
          {
            symRow = pre.cRow;
            symHdr = pre.cHdr;
          }
          else
          {
            const long long tb = wave_clock();
            symRow = build_centre_bitmap<CH>(a.rec, a.window, a.BX, a.BY, coy * a.BX + cox, crx, cry, a.symMaxL, a.symMaxR, a.symMaxD, scratch, symHdr);
            pre.ccx = cox; pre.ccy = coy; pre.cRow = symRow; pre.cHdr = symHdr;
            nBuilt++;
            tOnDemand += wave_clock() - tb;
          }
        }

        if (!symHdr) nFourNoSym++;
        cause = symHdr ? 1 : 2;
        const long long tr0 = wave_clock();
        tSym += tr0 - t0;

        Region c;
        c.ax = cox - LIMG_SYM_BACK; c.ay = coy - LIMG_SYM_BACK;
        c.vx0 = c.ax + (int)(symHdr & 0xFF); c.vy0 = c.ay + (int)((symHdr >> 8) & 0xFF);
        c.vx1 = c.ax + (int)((symHdr >> 16) & 0xFF); c.vy1 = c.ay + (int)(symHdr >> 24);
        c.avail = symHdr ? (symRow & ~region_used(sn, c.ax, c.ay, (int)(symHdr >> 24))) : 0u;
        const int centre = coy * a.BX + cox;
        const long long tg0 = wave_clock();
        tRegion += tg0 - tr0;
        grow(centre, c, true, 0, cox, coy, crx, cry);
        tGrow4 += wave_clock() - tg0;
        r.cox = cox; r.coy = coy; r.crx = crx; r.cry = cry;
        r.attempted = 1;
        r.kind = (crx * cry > r.rx * r.ry) ? 2 : 1;
        r.boxR = max(r.boxR, cox + crx);
        r.boxD = max(r.boxD, coy + cry);
        tFour += wave_clock() - t0;
        cause = 0;
      }
    }
    else
    {
      r.kind = (r.rx > 1 || r.ry > 1) ? 1 : 0;
    }

    // every in-use bit the growth consulted lies inside the two boxes of same_where_probed(); were they all inside the snapshot?
    auto inside = [&](int bx0, int by0, int bx1, int by1) -> bool {
      return max(by0, 0) >= sn.r0 && min(by1, a.BY) <= sn.r0 + 32 && max(bx0, 0) >= sn.w0 * 32 && min(bx1, a.BX) <= sn.w0 * 32 + 96;
    };

    volatileReads = !inside(x, y, x + r.rx, y + r.ry) || (r.attempted && !inside(r.cox, r.coy, r.cox + r.crx, r.coy + r.cry));
    return r;
  }
};

__device__ __forceinline__ uint32_t global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return (uint32_t)t;
}

__device__ __forceinline__ uint2 pack_rect(int ox, int oy, int rx, int ry)
{
  return make_uint2((uint32_t)ox | ((uint32_t)oy << 16), (uint32_t)rx | ((uint32_t)ry << 16));
}

// next column >= x of row y whose candidate bit is set and which can still emit given a fresh read of the in-use bits: in stage 0
// its whole 3 x 3 corner must be free (limg.cpp:1424 keeps nothing smaller), in stage 1 the block and its right or lower neighbour.
// In-use bits at columns >= x of these rows only ever come from logically earlier rectangles, so skipping is exact. BX if none.
// `used(dy, w)` returns in-use word w of block row y + dy (rows are padded by two zero words).
template <class UsedWord>
__device__ __forceinline__ int wave_next_candidate(const uint32_t *candRow, UsedWord used, int nWords, int x, int BX, int stage, int lane)
{
  for (int w0 = x >> 5; w0 < nWords; w0 += 32)
  {
    const int w = w0 + lane;
    uint32_t bits = 0;

    if (w < nWords)
    {
      // rows are padded with zero words, and candidates never sit in the last block rows (their 3 x 3 corner / lower neighbour is inside the grid)
      const unsigned long long u0 = used(0, w) | ((unsigned long long)used(0, w + 1) << 32);
      bits = __ldg(candRow + w);

      if (stage == 0 && bits) // a candidate bit in this word: block rows y + 1 and y + 2 exist
      {
        const unsigned long long u1 = used(1, w) | ((unsigned long long)used(1, w + 1) << 32);
        const unsigned long long u2 = used(2, w) | ((unsigned long long)used(2, w + 1) << 32);
        const unsigned long long u = u0 | u1 | u2;
        bits &= (uint32_t)~(u | (u >> 1) | (u >> 2));
      }
      else
      {
        // the lower neighbour may be outside the grid in the last row: then only the right neighbour counts (its bit is clear in the padding)
        bits &= (uint32_t)~u0;
      }

      if (w == (x >> 5))
        bits &= 0xFFFFFFFFu << (x & 31);
    }

    const uint32_t any = __ballot_sync(0xFFFFFFFFu, bits != 0);

    if (any)
    {
      const int first = __ffs(any) - 1;
      const uint32_t b = __shfl_sync(0xFFFFFFFFu, bits, first);
      return min((w0 + first) * 32 + __ffs(b) - 1, BX);
    }
  }

  return BX;
}

#define LIMG_WAVE_SPIN_LIMIT (1u << 22) // watchdog: a wait that long (seconds) is a bug; flag it instead of hanging the GPU

// ---- how the rows of the scan coordinate (both back ends) -----------------------------------------------------------------
// Per stage: a counter of rows handed out, one progress word per row (the column up to which the row's seeds are decided and their
// claims fenced; LIMG_WAVE_DONE when the row is finished) and a counter of DONE rows: every row below it is finished. A row that
// needs to know how far the rows above are takes the minimum of the progress words of the rows from the done counter up to itself:
// no row relays anybody else's progress, so the information is as fresh as one read, and a finished row's warp is free at once.
// Rows are handed out dynamically: the next stage-1 row if stage 0 is already done down to `stageGap` rows below it, else the
// next stage-0 row, else (stage 0 handed out completely) the next stage-1 row. Whatever a row waits for (rows above it in its own
// stage, stage-0 rows when every stage-0 row is handed out) is therefore held by a running warp or finished.
// counters[4] = { next stage-0 row, done stage-0 rows, next stage-1 row, done stage-1 rows }.

__device__ __forceinline__ void fence_sc_gpu() { asm volatile("fence.sc.gpu;" ::: "memory"); }

// common part of the back ends' take_row(): lane 0 decides, the warp gets (stage, row) or false when nothing is left
template <class Backend>
__device__ __forceinline__ bool wave_take_row(const Backend &be, int BY, int gap, int sequential, int lane, int maxStage1, int &stage, int &y)
{
  int packed = -1;

  if (lane == 0)
  {
    const int next1 = (int)be.peek_next(1);

    // (at most maxStage1 stage-1 rows in flight while stage-0 rows are left to hand out: a stage-1 seed with a tall rectangle waits for
    // stage-0 rows further down, which therefore must always find a warp)
    if (next1 < BY && (int)be.done_rows(0) >= (sequential ? BY : min(next1 + gap + 1, BY)) && ((int)be.peek_next(0) >= BY || next1 - (int)be.done_rows(1) < maxStage1))
    {
      const int k = (int)be.take_next(1);

      if (k < BY)
        packed = (k << 1) | 1;
    }

    if (packed < 0 && (int)be.peek_next(0) < BY)
    {
      const int k = (int)be.take_next(0);

      if (k < BY)
        packed = k << 1;
    }

    if (packed < 0)
    {
      const int k = (int)be.take_next(1);

      if (k < BY)
        packed = (k << 1) | 1;
    }
  }

  packed = __shfl_sync(0xFFFFFFFFu, packed, 0);
  stage = packed & 1;
  y = packed >> 1;
  return packed >= 0;
}

// minimum of the progress words of the unfinished rows above row y (LIMG_WAVE_DONE if there is none)
template <class Backend>
__device__ __forceinline__ int wave_rows_above(const Backend &be, int stage, int y, int lane)
{
  const int base = (int)be.done_rows(stage); // read first: rows finishing meanwhile only add reads
  int p = LIMG_WAVE_DONE;

  for (int r0 = y - 1; r0 >= base; r0 -= 32)
  {
    const int r = r0 - lane;

    if (r >= base)
      p = min(p, be.progress(stage, r));
  }

  return __reduce_min_sync(0xFFFFFFFFu, p);
}

// after the row published LIMG_WAVE_DONE: move the stage's done counter over every leading finished row. Two rows that finish at the
// same time must not both miss the other's DONE word (store, sequentially consistent fence, then the reads).
template <class Backend>
__device__ __forceinline__ void wave_advance_done(const Backend &be, int stage, int BY, int lane)
{
  be.fence_sc();
  const int before = (int)be.done_rows(stage);
  int b = before;

  for (;;)
  {
    const bool done = b + lane < BY && be.progress(stage, b + lane) == LIMG_WAVE_DONE;
    const uint32_t notDone = ~__ballot_sync(0xFFFFFFFFu, done);
    const int n = notDone ? __ffs(notDone) - 1 : 32;
    b += n;

    if (n < 32)
      break;
  }

  if (b > before)
  {
    be.acquire_fence(); // what the finished rows claimed is visible to whoever sees the new count
    be.set_done_rows(stage, (uint32_t)b, lane);
  }
}

// The scan's shared state (in-use mask, row progress, counters) in GLOBAL memory, read and written at L2: any grid, any image size.
struct WaveGlobal
{
  typedef LiveMask Mask;
  const WaveArgs &a;

  __device__ __forceinline__ Mask mask() const { return LiveMask{ GlobalWords{ a.used }, a.wordsPerRow, a.BX, a.BY }; }
  __device__ __forceinline__ uint32_t peek_next(int stage) const { return ld_relaxed_u32(&a.ticket[2 * stage]); }
  __device__ __forceinline__ uint32_t take_next(int stage) const { return atomicAdd(&a.ticket[2 * stage], 1u); }
  __device__ __forceinline__ uint32_t done_rows(int stage) const { return ld_relaxed_u32(&a.ticket[2 * stage + 1]); }

  __device__ __forceinline__ void set_done_rows(int stage, uint32_t n, int lane) const
  {
    if (lane == 0)
      atomicMax(&a.ticket[2 * stage + 1], n);
  }

  // everything read after this fence is at least as new as what was read before it
  __device__ __forceinline__ void acquire_fence() const { fence_acq_rel_gpu(); }
  __device__ __forceinline__ void fence_sc() const { fence_sc_gpu(); }
  __device__ __forceinline__ int progress(int stage, int row) const { return ld_relaxed_s32(a.progress + (size_t)stage * a.BY + row); }

  __device__ __forceinline__ void publish(int stage, int y, int v, int lane) const
  {
    if (lane == 0)
      st_relaxed_s32(a.progress + (size_t)stage * a.BY + y, v);
  }

  __device__ __forceinline__ uint32_t used_word(int y, int w) const { return ld_relaxed_u32(a.used + (size_t)y * a.wordsPerRow + w); }

  // claim the blocks of a rectangle; visible to this warp's next look at the mask and to everybody who later reads a progress
  // value published after it
  __device__ __forceinline__ void claim(int eox, int eoy, int erx, int ery, int lane) const
  {
    for (int rr = lane; rr < ery; rr += 32)
    {
      uint32_t *row = a.used + (size_t)(eoy + rr) * a.wordsPerRow;

      for (int xx = eox; xx < eox + erx;)
      {
        const int w0 = xx >> 5, b0 = xx & 31;
        const int cnt = min(32 - b0, eox + erx - xx);
        const uint32_t m = (cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << b0;
        atomicOr(&row[w0], m);
        xx += cnt;
      }
    }

    __threadfence();
    __syncwarp();
  }
};

// One warp per block row, rows by ticket, both merge stages in one launch. A stage-1 row starts once stage 0 is done with every
// row down to `stageGap` rows below it (the centre-third regrowth of a stage-0 seed further down would have to reach that far up to
// matter, which the verification pass would notice). `attempt` numbers the tries of the host. `scratch`: 32 words of shared memory
// private to the warp. `sequential`: rows strictly one after the other (the reference's order).
template <int CH, class Backend>
__device__ void wave_scan_rows(const WaveArgs &a, const Backend &be, int attempt, int sequential, uint32_t *scratch)
{
  const int lane = threadIdx.x & 31;
  WaveScan<CH, typename Backend::Mask> scan{ a, be.mask(), lane, 0 };
  scan.scratch = scratch;
  const int nWords = (a.BX + 31) >> 5;
  uint32_t nExp[2] = { 0, 0 }, nReexp[2] = { 0, 0 }, nPolls[2] = { 0, 0 }, nOnDemand[2] = { 0, 0 };
  long long tNext = 0, tWait = 0, tExpand = 0, tClaim = 0, tPre = 0, tc;
  long long cSeedStart = 0, cPre = 0, cWait = 0, cExp = 0, cClaim = 0, cTotal[2] = { 0, 0 }, cParts[2][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 } };
  uint32_t cIters = 0, cCount[2] = { 0, 0 }, cItersSum[2] = { 0, 0 };
  bool failed = false;

  for (;;)
  {
    int stage, y;

    if (!wave_take_row(be, a.BY, a.stageGap, sequential, lane, max(1, (int)(gridDim.x * (blockDim.x >> 5)) / 2), stage, y))
      break;

    const uint32_t *cand = a.candBits + (size_t)stage * a.BY * a.wordsPerRow;
    uint2 *lists = a.rowLists + (size_t)stage * a.BY * a.listCap;
    uint32_t *emitInfo = a.emitInfo + (size_t)stage * a.BX * a.BY;
    const uint32_t base = stage ? LIMG_TAU_STAGE1 : 0u;
    const uint32_t onDemandBefore = scan.nOnDemand;
    uint32_t rowExp = 0, rowReexp = 0, rowPolls = 0; // (arrays indexed by `stage` would live in local memory)

    if (stage == 1)
    {
      // wait for stage 0 (rows are done in order, so a count of finished rows is enough)
      const uint32_t need = sequential ? (uint32_t)a.BY : (uint32_t)min(y + a.stageGap + 1, a.BY);

      if (lane == 0)
      {
        uint32_t spins = 0;

        while (be.done_rows(0) < need)
        {
          if (++spins > (LIMG_WAVE_SPIN_LIMIT << 3)) { a.flags[3] = 1; break; }
          __nanosleep(200);
        }
      }

      __syncwarp();
      be.acquire_fence();
    }

    if (sequential && y > 0)
    {
      // strictly one row after the other: the reference's order
      if (lane == 0)
      {
        uint32_t spins = 0;

        while ((int)be.done_rows(stage) < y)
        {
          if (++spins > (LIMG_WAVE_SPIN_LIMIT << 3)) { a.flags[3] = 1; break; }
          __nanosleep(100);
        }
      }

      __syncwarp();
      be.acquire_fence();
    }

    uint32_t *rowT = (LIMG_WAVE_PROFILE && a.dbgRows) ? a.dbgRows + ((size_t)stage * a.BY + y) * 4 : nullptr;

    if (rowT && lane == 0)
      rowT[0] = global_ns();

    const uint32_t *candRow = cand + (size_t)y * a.wordsPerRow;
    uint2 *list = lists + (size_t)y * a.listCap;
    uint32_t count = 0;
    int x = 0, published = 0, nEvents = 0, xEvent = 0;
    SeedLinks ahead;
    ahead.x = -1;

    // issue the loads of the candidate that will probably come next: the first live candidate right of `from`
    auto look_ahead = [&](int from) {
      if (ahead.x != -1) // -2: looked, and there is none
        return;

      const int xn = from < a.BX ? wave_next_candidate(candRow, [&](int dy, int w) { return be.used_word(y + dy, w); }, nWords, from, a.BX, stage, lane) : a.BX;

      if (xn < a.BX)
        ahead = scan.prefetch_links(xn, y, stage);
      else
        ahead.x = -2;
    };

    // What the row promises the rows below: every seed left of the published column is decided (its claims were fenced when they were made)
    // and no seed that is still undecided can touch anything left of it. `rowSafe` is the safe column (k_plan_safe) of the row's undecided
    // candidates: a stage-0 regrowth also grows to the left of its seed.
    int rowSafe = (stage == 0 && a.safe) ? (int)(__ldg(&a.safe[(size_t)y * a.BX]) & 0xFFFFu) : LIMG_WAVE_DONE;

    auto publish = [&](int own) {
      const int v = own == LIMG_WAVE_DONE ? own : min(own, rowSafe);

      if (v > published)
      {
        be.publish(stage, y, v, lane);
        published = v;
      }
    };

    for (;;)
    {
      tc = wave_clock();
      x = wave_next_candidate(candRow, [&](int dy, int w) { return be.used_word(y + dy, w); }, nWords, x, a.BX, stage, lane);

      publish(x >= a.BX ? LIMG_WAVE_DONE : x);
      tNext += wave_clock() - tc;

      if (x >= a.BX)
        break;

      tc = wave_clock();
      cSeedStart = tc;
      // the seed's links were loaded while the row decided the seed before, if the guess of the next candidate was right
      const SeedLinks links = ahead.x == x ? ahead : scan.prefetch_links(x, y, stage);
      SeedPre pre = scan.prefetch_bitmaps(links, y, stage);
      ahead.x = -1;

      const int safeAfter = (stage == 0 && a.safe) ? (int)(links.safe >> 16) : LIMG_WAVE_DONE;

      if (stage == 0 && a.safe)
      {
        // The static safe column of this candidate is a mask-free bound. The regrowth's rectangle spans its centre's block row from the centre
        // (right of the seed) leftwards without a gap, so it also stops at the nearest block at or left of the seed column that is in use NOW in
        // that row (in-use bits never clear): usually the rectangle the row emitted just before. The centre's row is 1 .. (run along the seed's
        // column) / 3 rows below the seed.
        const uint32_t col = (links.w0 & 1u) | ((links.w0 >> 7) & 2u) | ((links.w0 >> 14) & 4u) | ((links.w0 >> 21) & 8u) | ((links.w1 & 1u) << 4) | ((links.w1 >> 3) & 32u) |
                             ((links.w1 >> 10) & 64u) | ((links.w1 >> 17) & 128u);
        const int colRun = __ffs((int)(~col & 0x1FFu)) - 1;
        int blocked = 0x7FFF; // leftmost column the regrowth can reach in this lane's candidate centre row

        if (!(colRun == 8 && y + 8 < a.BY) && lane < colRun / 3 && y + 1 + lane < a.BY)
        {
          const int cy = y + 1 + lane, w = x >> 5;
          uint32_t bits = be.used_word(cy, w) & (0xFFFFFFFFu >> (31 - (x & 31)));
          blocked = -1; // nothing in use within reach of the two words: no bound from the mask

          if (bits)
            blocked = w * 32 + 32 - __clz((int)bits);
          else if (w > 0 && (bits = be.used_word(cy, w - 1)) != 0)
            blocked = (w - 1) * 32 + 32 - __clz((int)bits);
        }

        const int dyn = __reduce_min_sync(0xFFFFFFFFu, blocked);
        const int mine = max((int)(links.safe & 0xFFFFu), dyn == 0x7FFF ? -1 : dyn);
        rowSafe = min(mine, safeAfter); // (the dead candidates the row skipped no longer count either)
        publish(x);
      }

      tPre += wave_clock() - tc;
      cPre = wave_clock() - tc; cWait = 0; cExp = 0; cClaim = 0; cIters = 0;
      const uint32_t first = count;
      int nextX = x + 1;
      bool claimed = false;
      xEvent = x;

      for (int k = 0;; k++)
      {
        // Speculate while the rows above are not far enough: the seed is expanded against the mask as it is NOW and expanded again
        // only when a bit it consulted changes. Once the rows above have passed everything it consulted (+ margin), a snapshot
        // taken after that observation (and after an acquire fence) which agrees with the one the expansion used makes the
        // expansion final.
        WaveResult r;
        Snapshot used;
        bool have = false, taken = false, unstable = false;

        for (uint32_t spins = 0;; spins++)
        {
          tc = wave_clock();
          int p = LIMG_WAVE_DONE;

          if (y > 0 && !sequential)
            p = wave_rows_above(be, stage, y, lane);

          // A stage-1 rectangle reaches down into rows whose stage-0 seeds must all be decided first (its row started when stage 0
          // was done stageGap rows below the seed; a taller rectangle waits for the same distance below its last rows).
          const int below = (stage == 1 && !sequential) ? (int)be.done_rows(0) : a.BY;

          if (p < min(x + 1 + a.margin - a.specAhead, a.BX))
          {
            // the rows above are still far away: whatever the mask shows now is not worth expanding against
            tWait += wave_clock() - tc;

            if (spins > LIMG_WAVE_SPIN_LIMIT) { a.flags[3] = 1; taken = true; break; }

            rowPolls++;
            if (p + 64 < x) __nanosleep(200);
            continue;
          }

          // The fence orders the mask reads below behind the progress read above (PTX does not order two loads by a branch between
          // them); it is only paid for when this look at the mask can be the final one.
          const bool final = have ? p >= min(r.boxR + a.margin, a.BX) && below >= min(r.boxD + a.stageGap / 2, a.BY) : p >= min(x + 1 + a.margin + 8, a.BX);

          if (final && !(a.experiment & 1))
            be.acquire_fence();

          const Snapshot sn = scan.snapshot(x, y); // after the progress read
          tWait += wave_clock() - tc;
          cWait += wave_clock() - tc; cIters++;

          if (scan.snap_used(sn, x)) { taken = true; break; }

          if (!have || unstable || !scan.same_where_probed(sn, used, r, x, y))
          {
            tc = wave_clock();

            if (have) rowReexp++;

            r = scan.expand(x, y, stage, pre, sn);
            unstable = scan.volatileReads; // it read in-use bits outside the snapshot: only good if the rows above had already passed
            used = sn;
            have = true;
            rowExp++;
            const long long dt = wave_clock() - tc;
            tExpand += dt;
            cExp += dt;

            if (LIMG_WAVE_PROFILE && a.dbg && lane == 0)
              atomicAdd(&a.dbg[stage * 16 + min(15, 63 - __clzll((dt >> 8) | 1))], 1u);
          }

          // A seed that emits nothing now never will (in-use bits are only ever set, so its rectangle can only shrink): no need to wait for the rows above.
          if (r.kind == 0)
            break;

          if (final && p >= min(r.boxR + a.margin, a.BX) && below >= min(r.boxD + a.stageGap / 2, a.BY))
          {
            if (LIMG_WAVE_PROFILE && a.dbg && lane == 0 && stage == 0)
            {
              // diagnostics: how far ahead the rows above are when the seed's decision stands, how wide its probe box was, and
              // whether it had to wait at all
              atomicAdd(&a.dbg[128 + min(31, (min(p, a.BX) - x) >> 2)], 1u);
              atomicAdd(&a.dbg[160 + min(31, (r.boxR - x) >> 1)], 1u);
              atomicAdd(&a.dbg[192 + (spins == 0 ? 0 : 1)], 1u);
            }

            break;
          }

          if (spins > LIMG_WAVE_SPIN_LIMIT) { a.flags[3] = 1; break; }

          // the seed has to wait: the next candidate's links can be on their way meanwhile
          look_ahead(r.kind == 1 ? x + r.rx : x + 1);

          if (p >= min(r.boxR + a.margin, a.BX) && below >= min(r.boxD + a.stageGap / 2, a.BY))
            continue; // far enough, but this look was not fenced: look again at once

          rowPolls++;
          if (p + 64 < x) __nanosleep(200);
        }

        if (taken || r.kind == 0)
          break;

        const int eox = r.kind == 2 ? r.cox : x, eoy = r.kind == 2 ? r.coy : y;
        const int erx = r.kind == 2 ? r.crx : r.rx, ery = r.kind == 2 ? r.cry : r.ry;
        const uint32_t T = base + ((uint32_t)(y * a.BX + x) << 3) + (uint32_t)min(k, LIMG_WAVE_MAX_ATTEMPTS - 1);

        if (k >= LIMG_WAVE_MAX_ATTEMPTS && !sequential)
          failed = true; // more regrowths from one seed than the time stamp encodes: let the sequential pass do it

        // claim: in-use bits and owner times. Two rectangles that overlap (a failed speculation) leave one of them with a foreign
        // owner time on a block, which the verification pass sees.
        tc = wave_clock();
        be.claim(eox, eoy, erx, ery, lane);

        // hand over to the rows below as early as possible: a right/down rectangle decides every seed up to its right edge
        if (r.kind == 1 && x + r.rx < a.BX)
        {
          rowSafe = safeAfter;
          publish(x + r.rx);
        }

        // bookkeeping nobody waits for (read after the kernel): owner times, the row's list
        for (int e = lane; e < erx * ery; e += 32)
          a.tau[(size_t)(eoy + e / erx) * a.BX + eox + e % erx] = T;

        if (lane == 0)
        {
          if (count < (uint32_t)a.listCap)
            list[count] = pack_rect(eox, eoy, erx, ery);
          else
            a.flags[4] = 1; // reported by the host as LIMGCU_ERROR_OUT_OF_BOUNDS
        }

        count++;
        claimed = true;

        if (r.kind == 1)
          look_ahead(x + r.rx);
        tClaim += wave_clock() - tc;
        cClaim += wave_clock() - tc;

        if (r.kind == 2)
        {
          // limg.cpp:1435-1438: the scan resumes at the same seed (which the regrowth may or may not have covered). In stage 0 the
          // seed can only emit again if its whole 3 x 3 corner is still free, i.e. if the regrowth rectangle stays clear of it.
          if (!(x < eox + erx && x + 3 > eox && y < eoy + ery && y + 3 > eoy))
            continue;

          break;
        }

        nextX = x + r.rx;
        break;
      }

      if (claimed && lane == 0)
        emitInfo[(size_t)y * a.BX + x] = (first << 8) | (count - first);

      if (LIMG_WAVE_PROFILE && claimed)
      {
        cTotal[stage] += wave_clock() - cSeedStart;
        cParts[stage][0] += cPre; cParts[stage][1] += cWait; cParts[stage][2] += cExp; cParts[stage][3] += cClaim;
        cCount[stage]++;
        cItersSum[stage] += cIters;
      }

      x = nextX;
      rowSafe = safeAfter; // the seed at the old x is decided

      if (rowT && lane == 0)
      {
        const uint32_t tn = global_ns();
        if (rowT[1] == 0) rowT[1] = tn;
        rowT[2] = tn;

        if (stage == 0 && y >= a.eventRow && y < a.eventRow + 4 && nEvents < 64)
        {
          uint32_t *ev = a.dbgRows + (size_t)8 * a.BY + ((size_t)(y - a.eventRow) * 64 + nEvents) * 2;
          ev[0] = (uint32_t)xEvent | ((uint32_t)(claimed ? 1 : 0) << 16);
          ev[1] = tn;
          nEvents++;
        }
      }

      // hand over to the rows below before looking for the next candidate
      if (x < a.BX)
        publish(x);
    }

    if (lane == 0)
      a.rowCounts[(size_t)stage * a.BY + y] = min(count, (uint32_t)a.listCap);

    if (rowT && lane == 0)
      rowT[3] = global_ns();

    // the row published LIMG_WAVE_DONE when it ran out of candidates
    wave_advance_done(be, stage, a.BY, lane);

    nOnDemand[stage] += scan.nOnDemand - onDemandBefore;
    nExp[stage] += rowExp;
    nReexp[stage] += rowReexp;
    nPolls[stage] += rowPolls;
  }

  if (failed && !sequential && lane == 0)
  {
    a.flags[5] |= 4u << (4 * attempt);
    a.flags[0] = (uint32_t)attempt + 1;
  }

  if (a.stats && lane == 0)
  {
    for (int st = 0; st < 2; st++)
    {
      atomicAdd(&a.stats[0 + st * 4], nExp[st]);
      atomicAdd(&a.stats[1 + st * 4], nReexp[st]);
      atomicAdd(&a.stats[2 + st * 4], nPolls[st]);
      atomicAdd(&a.stats[3 + st * 4], nOnDemand[st]);
    }

    // profile (kilocycles, summed over warps and both stages): next-candidate + publish, wait, expand (incl. four-way, on-demand), four-way, on-demand, claim, prefetch
    atomicAdd(&a.stats[8], (uint32_t)(tNext >> 10));
    atomicAdd(&a.stats[9], (uint32_t)(tWait >> 10));
    atomicAdd(&a.stats[10], (uint32_t)(tExpand >> 10));
    atomicAdd(&a.stats[11], (uint32_t)(scan.tFour >> 10));
    atomicAdd(&a.stats[12], (uint32_t)(scan.tOnDemand >> 10));
    atomicAdd(&a.stats[13], (uint32_t)(tClaim >> 10));
    atomicAdd(&a.stats[14], (uint32_t)(tPre >> 10));

    if (a.dbg)
    {
      for (int c = 0; c < 3; c++)
      {
        atomicAdd(&a.dbg[32 + c], scan.nStrips[c]);
        atomicAdd(&a.dbg[32 + 4 + c], (uint32_t)(scan.tStrips[c] >> 10));
      }

      for (int st = 0; st < 2; st++)
      {
        atomicAdd(&a.dbg[200 + st * 8 + 0], cCount[st]);
        atomicAdd(&a.dbg[200 + st * 8 + 1], (uint32_t)(cTotal[st] >> 10));
        atomicAdd(&a.dbg[200 + st * 8 + 2], (uint32_t)(cParts[st][0] >> 10));
        atomicAdd(&a.dbg[200 + st * 8 + 3], (uint32_t)(cParts[st][1] >> 10));
        atomicAdd(&a.dbg[200 + st * 8 + 4], (uint32_t)(cParts[st][2] >> 10));
        atomicAdd(&a.dbg[200 + st * 8 + 5], (uint32_t)(cParts[st][3] >> 10));
        atomicAdd(&a.dbg[200 + st * 8 + 6], cItersSum[st]);
      }

      atomicAdd(&a.dbg[53], scan.nBuilt);
      atomicAdd(&a.dbg[224], (uint32_t)(scan.tSeedFast >> 10));
      atomicAdd(&a.dbg[225], (uint32_t)(scan.tSeedGeneral >> 10));
      atomicAdd(&a.dbg[226], (uint32_t)(scan.tSym >> 10));
      atomicAdd(&a.dbg[227], (uint32_t)(scan.tRegion >> 10));
      atomicAdd(&a.dbg[228], (uint32_t)(scan.tGrow4 >> 10));
      atomicAdd(&a.dbg[230], scan.nSeedFast);
      atomicAdd(&a.dbg[231], scan.nSeedGeneral);
      atomicAdd(&a.dbg[48], scan.nFour);
      atomicAdd(&a.dbg[49], scan.nFourMiss);
      atomicAdd(&a.dbg[50], scan.nFourNoSym);
    }
  }
}

// The scan over the live mask in global memory: any image size; with `sequential` (one CTA) it is the reference's order, the last resort
// of the host's tries. `attempt`: the kernel runs only if flags[0] == attempt, i.e. every earlier try failed.
template <int CH>
__global__ void __launch_bounds__(LIMG_WAVE_WARPS * 32) k_merge_wave(const __grid_constant__ WaveArgs a, int attempt, int sequential)
{
  if (a.flags[0] != (uint32_t)attempt)
    return;

  __shared__ uint32_t sScratch[LIMG_WAVE_WARPS][32];
  wave_scan_rows<CH>(a, WaveGlobal{ a }, attempt, sequential, sScratch[threadIdx.x >> 5]);
}

// Verification, part 1 (one THREAD per candidate seed): most candidates were in use before their turn came, or could not emit
// any more (stage 0: a block of the 3 x 3 corner taken, stage 1: both neighbours taken). Those must not have emitted anything;
// the others go on the replay list.
__global__ void __launch_bounds__(256) k_merge_verify_filter(WaveArgs a, int attempt, uint32_t *replayList, uint32_t *replayCount)
{
  if (a.flags[0] != (uint32_t)attempt)
    return;

  const int stage = (int)blockIdx.y;

  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t blocks = (uint32_t)(a.BX * a.BY);
  bool replay = false;
  uint32_t seed = 0;

  if (i < a.candCount[stage])
  {
    seed = __ldg(&a.candList[(size_t)stage * blocks + i]);
    const int y = (int)(seed / (uint32_t)a.BX), x = (int)(seed - (uint32_t)y * (uint32_t)a.BX);
    const uint32_t T = (stage ? LIMG_TAU_STAGE1 : 0u) + (seed << 3);
    const uint32_t *t = a.tau + seed;
    bool dead;

    if (stage == 0) // the candidate bit says the corner is inside the grid
    {
      dead = false;

#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++)
          dead |= __ldg(t + (size_t)dy * a.BX + dx) < T;
    }
    else
    {
      const bool rightFree = x + 1 < a.BX && !(__ldg(t + 1) < T), downFree = y + 1 < a.BY && !(__ldg(t + a.BX) < T);
      dead = __ldg(t) < T || (!rightFree && !downFree);
    }

    if (dead)
    {
      if (__ldg(&a.emitInfo[(size_t)stage * blocks + seed]) & 0xFFu)
        a.flags[1 + stage] = 1; // recorded an emission it cannot have made
    }
    else
    {
      replay = true;
    }
  }

  const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, replay);
  uint32_t pos = 0;

  if ((threadIdx.x & 31) == 0 && ballot)
    pos = atomicAdd(&replayCount[stage], (uint32_t)__popc(ballot));

  pos = __shfl_sync(0xFFFFFFFFu, pos, 0) + __popc(ballot & ((1u << (threadIdx.x & 31)) - 1u));

  if (replay)
    replayList[(size_t)stage * blocks + pos] = seed;
}

// Verification, part 2 (one WARP per remaining seed): replays the seed against the mask at its logical time and compares with the
// wave's record.
template <int CH>
__global__ void __launch_bounds__(256) k_merge_verify(const __grid_constant__ WaveArgs a, int attempt, const uint32_t *replayList, const uint32_t *replayCount)
{
  if (a.flags[0] != (uint32_t)attempt)
    return; // this try did not run, or already failed

  const int stage = (int)blockIdx.y; // both stages in one launch: each is a few seeds per warp, i.e. as long as its slowest replay
  __shared__ uint32_t sScratchV[8][32];
  const int lane = threadIdx.x & 31;
  const uint32_t n = replayCount[stage];
  const uint32_t *candList = replayList + (size_t)stage * a.BX * a.BY;
  const uint2 *lists = a.rowLists + (size_t)stage * a.BY * a.listCap;
  const uint32_t *emitInfo = a.emitInfo + (size_t)stage * a.BX * a.BY;
  const uint32_t base = stage ? LIMG_TAU_STAGE1 : 0u;
  const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);

  for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warpsTotal)
  {
    const int seed = (int)__ldg(&candList[i]);
    const int y = seed / a.BX, x = seed - y * a.BX;
    const uint32_t info = __ldg(&emitInfo[seed]);
    const uint32_t start = info >> 8, have = info & 0xFFu;
    const uint2 *rec = lists + (size_t)y * a.listCap + start;
    uint32_t e = 0;
    bool ok = true;
    int reason = 0; // diagnostics: which check failed (tools/fail_sweep.py)
    uint2 wantRect = make_uint2(0u, 0u);
    bool havePre = false;
    SeedPre pre;

    for (int k = 0;; k++)
    {
      if (k >= LIMG_WAVE_MAX_ATTEMPTS) { ok = false; break; }

      const uint32_t T = base + ((uint32_t)seed << 3) + (uint32_t)k;
      WaveScan<CH, TimeMask> scan{ a, TimeMask{ a.tau, a.BX, a.BY, T }, lane, 0 };
      scan.scratch = sScratchV[threadIdx.x >> 5];

      if (scan.mask.is_used(x, y)) { ok = e == have; break; }

      // the same necessary conditions wave_next_candidate() uses, before anything expensive: stage 0 needs its whole 3 x 3 corner
      // free (the candidate bit says it is inside the grid), stage 1 the right or the lower neighbour
      {
        bool busy = false;

        if (stage == 0)
        {
          if (lane < 9)
            busy = scan.mask.is_used(x + lane % 3, y + lane / 3);

          busy = __any_sync(0xFFFFFFFFu, busy);
        }
        else
        {
          const bool rightFree = x + 1 < a.BX && !scan.mask.is_used(x + 1, y), downFree = y + 1 < a.BY && !scan.mask.is_used(x, y + 1);
          busy = !rightFree && !downFree;
        }

        if (busy) { ok = e == have; reason = 1; break; }
      }

      if (!havePre) { pre = scan.prefetch(x, y, stage); havePre = true; }

      const Snapshot sn = scan.snapshot(x, y);
      const WaveResult r = scan.expand(x, y, stage, pre, sn);

      if (r.kind == 0) { ok = e == have; reason = 2; break; }

      const int eox = r.kind == 2 ? r.cox : x, eoy = r.kind == 2 ? r.coy : y;
      const int erx = r.kind == 2 ? r.crx : r.rx, ery = r.kind == 2 ? r.cry : r.ry;

      if (e >= have) { ok = false; reason = 3; wantRect = pack_rect(eox, eoy, erx, ery); break; }

      const uint2 want = pack_rect(eox, eoy, erx, ery), got = rec[e];
      e++;

      if (want.x != got.x || want.y != got.y) { ok = false; reason = 4; wantRect = want; break; }

      // every block of the rectangle is owned by it (no overlap with another rectangle)
      bool foreign = false;

      uint32_t foreignOwner = 0;

      for (int b = lane; b < erx * ery; b += 32)
      {
        const uint32_t owner = __ldg(&a.tau[(size_t)(eoy + b / erx) * a.BX + eox + b % erx]);

        if (owner != T)
        {
          foreign = true;
          foreignOwner = owner;
        }
      }

      if (__any_sync(0xFFFFFFFFu, foreign))
      {
        // diagnostics: the other owner's first rectangle
        const uint32_t other = __reduce_max_sync(0xFFFFFFFFu, foreignOwner);
        const uint32_t oseed = (other & (LIMG_TAU_STAGE1 - 1u)) >> 3;
        const int ostage = other >= LIMG_TAU_STAGE1 ? 1 : 0;
        const uint32_t oinfo = a.emitInfo[(size_t)ostage * a.BX * a.BY + oseed];
        const uint2 orect = (a.rowLists + ((size_t)ostage * a.BY + oseed / a.BX) * a.listCap)[oinfo >> 8];
        ok = false; reason = 5; wantRect = make_uint2(other, orect.x); e = orect.y;
        break;
      }

      if (r.kind == 1) { ok = e == have; reason = 6; break; }
    }

    if (!ok && lane == 0)
    {
      a.flags[1 + stage] = 1;

      // diagnostics: the first few failing seeds (stage, x, y, recorded count, replayed count so far)
      if (a.dbg)
      {
        const uint32_t slot = atomicAdd(&a.dbg[100], 1u);

        if (slot < 8)
        {
          uint32_t *o = a.dbg + 104 + slot * 8;
          o[0] = (uint32_t)stage | ((uint32_t)attempt << 8); o[1] = (uint32_t)x; o[2] = (uint32_t)y; o[3] = have; o[4] = e;
          o[5] = have ? rec[0].x : 0u; o[6] = have ? rec[0].y : 0u; o[7] = (uint32_t)reason;
          a.dbg[232 + slot * 2] = wantRect.x; a.dbg[233 + slot * 2] = wantRect.y; // (what the replay wanted / who else owns a block)
        }
      }
    }
  }
}

__global__ void k_merge_set_tries(WaveArgs a, int tries)
{
  a.flags[0] = (uint32_t)tries;
}

// after both verification kernels of a try: a failed stage fails the try
__global__ void k_merge_judge(WaveArgs a, int attempt)
{
  if (a.flags[0] == (uint32_t)attempt && (a.flags[1] | a.flags[2]))
  {
    a.flags[5] |= ((a.flags[1] ? 1u : 0u) | (a.flags[2] ? 2u : 0u)) << (4 * attempt); // which stage failed in which try (diagnostics)
    a.flags[0] = (uint32_t)attempt + 1;
    a.flags[1] = 0;
    a.flags[2] = 0;
  }
}

// preparation of try `attempt` > 0 (runs only when every earlier try failed): undo all claims and counters
__global__ void __launch_bounds__(256) k_merge_reset(WaveArgs a, int attempt)
{
  if (a.flags[0] != (uint32_t)attempt)
    return;

  const int blocks = a.BX * a.BY;
  const int n = max(blocks, max(a.BY * a.wordsPerRow, 2 * a.BY));

  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x)
  {
    if (b < blocks)
    {
      a.tau[b] = LIMG_TAU_NONE;
      a.emitInfo[b] = 0;
      a.emitInfo[blocks + b] = 0;
    }

    if (b < a.BY * a.wordsPerRow)
      a.used[b] = 0;

    if (b < 2 * a.BY)
    {
      a.progress[b] = 0;
      a.rowCounts[b] = 0;
    }

    if (b < 4)
      a.ticket[b] = 0;
  }
}

} // namespace limg

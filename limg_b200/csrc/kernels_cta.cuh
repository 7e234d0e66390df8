// limg_b200/csrc/kernels_cta.cuh -- the area-expansion scan (kernels_wave.cuh, limg.cpp:1294-1496) with its shared state in the
// SHARED MEMORY of one thread-block cluster instead of L2.
//
// The row-pipelined scan is a chain of dependent look-ups of three small things: the in-use mask (one bit per 8x8 block: 18 KB at
// 4K, 69 KB at 8K), one progress word per block row, and a row ticket. With them in global memory every look is an L2 round trip
// (~250 cycles), every claim a gpu-scope fence; k_merge_wave spent 95 % of its cycles with no eligible warp. Here every CTA of
// the cluster keeps a full REPLICA of the three in its own shared memory:
//   reads   are local shared-memory loads (~30 cycles): the poll of the 32 rows above, the 32 x 96 snapshot, the candidate search;
//   writes  (the claim of a rectangle, a progress value, the finished-rows count) go to every replica through distributed shared
//           memory (red / st .shared::cluster), ordered by one cluster-scope fence between a claim and the progress that follows it;
//   rows    are handed out from rank 0's counters (one remote read and one remote atomic per block row).
// What nobody waits for during the scan (owner times, the rows' rectangle lists, per-seed emission info) still goes to global
// memory and is consumed by the verification and the area preparation as before; the final mask is written back by rank 0.
// A cluster of 8 CTAs x 8 warps keeps 64 block rows in flight, about what the wavefront can use at 4K (BX / lag).
#pragma once

#include "kernels_wave.cuh"

#include <cooperative_groups.h>

namespace limg
{

#ifndef LIMG_CTA_WARPS
#define LIMG_CTA_WARPS 8
#endif

__device__ __forceinline__ uint32_t lds_volatile(uint32_t addr)
{
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// address of the same shared-memory location in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t cluster_map(uint32_t addr, uint32_t rank)
{
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void cluster_red_or(uint32_t caddr, uint32_t v) { asm volatile("red.relaxed.cluster.shared::cluster.or.b32 [%0], %1;" :: "r"(caddr), "r"(v) : "memory"); }
__device__ __forceinline__ void cluster_red_max(uint32_t caddr, uint32_t v) { asm volatile("red.relaxed.cluster.shared::cluster.max.u32 [%0], %1;" :: "r"(caddr), "r"(v) : "memory"); }
__device__ __forceinline__ void cluster_st(uint32_t caddr, uint32_t v) { asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" :: "r"(caddr), "r"(v) : "memory"); }

__device__ __forceinline__ uint32_t cluster_ld(uint32_t caddr)
{
  uint32_t v;
  asm volatile("ld.relaxed.cluster.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(caddr) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t cluster_atom_add(uint32_t caddr, uint32_t v)
{
  uint32_t r;
  asm volatile("atom.relaxed.cluster.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(caddr), "r"(v) : "memory");
  return r;
}

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

struct SharedWords
{
  uint32_t base; // shared-window address of word 0
  __device__ __forceinline__ uint32_t operator()(int i) const { return lds_volatile(base + 4u * (uint32_t)i); }
};

// The scan's shared state replicated in the shared memory of every CTA of the cluster.
struct WaveCluster
{
  typedef WordMask<SharedWords> Mask;
  const WaveArgs &a;
  uint32_t sUsed, sProgress, sMisc; // this CTA's replica: mask [BY][wordsPerRow], progress [2][BY], counters[4] (kernels_wave.cuh; the hand-out counters that count are rank 0's)
  uint32_t ranks;

  __device__ __forceinline__ Mask mask() const { return Mask{ SharedWords{ sUsed }, a.wordsPerRow, a.BX, a.BY }; }

  // the hand-out counters are rank 0's; the done counters are replicated
  __device__ __forceinline__ uint32_t peek_next(int stage) const { return cluster_ld(cluster_map(sMisc + 8u * (uint32_t)stage, 0)); }
  __device__ __forceinline__ uint32_t take_next(int stage) const { return cluster_atom_add(cluster_map(sMisc + 8u * (uint32_t)stage, 0), 1u); }
  __device__ __forceinline__ uint32_t done_rows(int stage) const { return lds_volatile(sMisc + 8u * (uint32_t)stage + 4u); }

  __device__ __forceinline__ void set_done_rows(int stage, uint32_t n, int lane) const
  {
    if ((uint32_t)lane < ranks)
      cluster_red_max(cluster_map(sMisc + 8u * (uint32_t)stage + 4u, (uint32_t)lane), n);
  }

  __device__ __forceinline__ void fence() const
  {
    if (ranks == 1)
      asm volatile("fence.acq_rel.cta;" ::: "memory");
    else
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
  }

  __device__ __forceinline__ void acquire_fence() const { fence(); }

  __device__ __forceinline__ void fence_sc() const
  {
    if (ranks == 1)
      asm volatile("fence.sc.cta;" ::: "memory");
    else
      asm volatile("fence.sc.cluster;" ::: "memory");
  }

  __device__ __forceinline__ int progress(int stage, int row) const { return (int)lds_volatile(sProgress + 4u * (uint32_t)(stage * a.BY + row)); }

  // one lane per replica; a row's progress words are always written by the same lane, so every replica sees them in order
  __device__ __forceinline__ void publish(int stage, int y, int v, int lane) const
  {
    if ((uint32_t)lane < ranks)
      cluster_st(cluster_map(sProgress + 4u * (uint32_t)(stage * a.BY + y), (uint32_t)lane), (uint32_t)v);
  }

  // the same for a row whose progress is published by several warps in turn (their stores could overtake each other: the word only ever grows)
  __device__ __forceinline__ void publish_max(int stage, int y, int v, int lane) const
  {
    if ((uint32_t)lane < ranks)
      cluster_red_max(cluster_map(sProgress + 4u * (uint32_t)(stage * a.BY + y), (uint32_t)lane), (uint32_t)v);
  }

  __device__ __forceinline__ uint32_t used_word(int y, int w) const { return lds_volatile(sUsed + 4u * (uint32_t)(y * a.wordsPerRow + w)); }

  __device__ __forceinline__ void claim(int eox, int eoy, int erx, int ery, int lane) const
  {
    for (int rr = lane; rr < ery; rr += 32)
    {
      const uint32_t row = sUsed + 4u * (uint32_t)((eoy + rr) * a.wordsPerRow);

      for (int xx = eox; xx < eox + erx;)
      {
        const int w0 = xx >> 5, b0 = xx & 31;
        const int cnt = min(32 - b0, eox + erx - xx);
        const uint32_t m = (cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u)) << b0;

        for (uint32_t r = 0; r < ranks; r++)
          cluster_red_or(cluster_map(row + 4u * (uint32_t)w0, r), m);

        xx += cnt;
      }
    }

    // every replica has the claim before any replica sees a progress value published after it
    if (a.experiment & 2)
      asm volatile("fence.acq_rel.cta;" ::: "memory");
    else
      fence();
    __syncwarp();
  }
};


// ---- a TEAM of warps per block row -----------------------------------------------------------------------------------------
// One warp per row spends ~19 k cycles on every seed that emits (seed data, expansion, regrowth, claim), one dependent instruction after
// the other, and the rows move in lock step: the scan's duration is (rows) x (that chain). Most of the chain does not depend on the
// row's earlier seeds: a candidate further right can be expanded against the mask as it is now, and the result stands if, when its
// turn comes, the in-use bits it consulted are unchanged (same_where_probed, the test the speculation against the rows above uses).
// So the warps of a team (same CTA) share a row: each OWNS one live candidate at a time, loads its data and expands it ahead of time;
// the row's decisions are still made strictly left to right: a candidate's TURN comes when every live candidate left of it is
// decided; only then (and with the rows above far enough) is its result final and its rectangle claimed. A candidate that has
// nothing to emit is decided on the spot, whoever's turn it is (it never will: in-use bits are only ever set). What is left on the
// row's critical path per emitting seed: the final look (poll, fence, snapshot, compare) and the claim.
struct TeamState
{
  int row;               // (y << 1) | stage of the team's row, -1: no rows left
  int decidedX;          // every candidate left of this column is decided
  int hint[8];           // per warp of the team: where the rectangle of the candidate it owns probably ends (0: owns none); candidates left of it will mostly be covered
  int published;         // last progress value of the row
  int rowSafe;           // safe column of the row's undecided candidates (see wave_scan_rows)
  int committing;        // the turn's owner is claiming: its bookkeeping (count, decidedX) is not complete yet
  uint32_t count;        // rectangles the row has emitted
  uint32_t pad;
  uint32_t owned[32];    // candidates some warp of the team works on (or has decided)
  uint32_t skip[32];     // candidates decided out of turn: nothing to emit
};

template <int CH>
__device__ void wave_scan_team(const WaveArgs &a, const WaveCluster &be, int attempt, TeamState *tsp, int teamWarps, int wit, int barId, int maxStage1, uint32_t *scratch)
{
  volatile TeamState *ts = tsp;
  const int lane = threadIdx.x & 31;
  WaveScan<CH, WaveCluster::Mask> scan{ a, be.mask(), lane, 0 };
  scan.scratch = scratch;
  const int nWords = (a.BX + 31) >> 5;
  uint32_t nExp0 = 0, nExp1 = 0, nReexp0 = 0, nReexp1 = 0, nPolls0 = 0, nPolls1 = 0, nOnDemand0 = 0, nOnDemand1 = 0, nWasted = 0, nTurnExp = 0;
  bool failed = false;

  auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(barId), "r"(teamWarps * 32) : "memory"); };

  for (;;)
  {
    if (wit == 0)
    {
      int stage = 0, y = 0;
      const bool got = wave_take_row(be, a.BY, a.stageGap, 0, lane, maxStage1, stage, y);

      if (got && stage == 1)
      {
        // wait for stage 0 (rows are done in order, so a count of finished rows is enough)
        const uint32_t need = (uint32_t)min(y + a.stageGap + 1, a.BY);

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
      }

      tsp->owned[lane] = 0;
      tsp->skip[lane] = 0;

      if (lane < 8)
        tsp->hint[lane] = 0;

      if (lane == 0)
      {
        tsp->row = got ? ((y << 1) | stage) : -1;
        tsp->decidedX = (got && stage == 1 && (a.experiment & 8)) ? a.BX : 0;
        tsp->published = 0;
        tsp->committing = 0;
        tsp->count = 0;
        tsp->rowSafe = (got && stage == 0 && a.safe) ? (int)(__ldg(&a.safe[(size_t)y * a.BX]) & 0xFFFFu) : LIMG_WAVE_DONE;
      }
    }

    team_sync();
    const int packed = ts->row;

    if (packed < 0)
      break;

    const int stage = packed & 1, y = packed >> 1;

    if (stage == 1)
      be.acquire_fence();

    const uint32_t *candRow = a.candBits + ((size_t)stage * a.BY + y) * a.wordsPerRow;
    uint2 *list = a.rowLists + ((size_t)stage * a.BY + y) * a.listCap;
    uint32_t *emitInfo = a.emitInfo + (size_t)stage * a.BX * a.BY;
    const uint32_t base = stage ? LIMG_TAU_STAGE1 : 0u;
    const uint32_t onDemandBefore = scan.nOnDemand;
    uint32_t rowExp = 0, rowReexp = 0, rowPolls = 0;

    uint32_t *rowT = (LIMG_WAVE_PROFILE && a.dbgRows) ? a.dbgRows + ((size_t)stage * a.BY + y) * 4 : nullptr; // time stamps: row taken, first and last decision, done

    if (rowT && wit == 0 && lane == 0)
      rowT[0] = global_ns();

    auto used_word = [&](int dy, int w) { return be.used_word(y + dy, w); };
    auto skipped = [&](int w) { return ts->skip[w & 31]; };
    auto skipped_or_owned = [&](int w) { return ts->skip[w & 31] | ts->owned[w & 31]; };

    // only the warp whose turn it is publishes
    auto publish = [&](int own) {
      const int v = min(own, ts->rowSafe);

      if (v > ts->published)
      {
        be.publish_max(stage, y, v, lane);
        __syncwarp();

        if (lane == 0)
          tsp->published = v;
      }
    };

    auto try_own = [&](int c) {
      int ok = 0;

      if (lane == 0)
        ok = (atomicOr(&tsp->owned[c >> 5], 1u << (c & 31)) >> (c & 31)) & 1u ? 0 : 1;

      return __shfl_sync(0xFFFFFFFFu, ok, 0) != 0;
    };

    for (;;)
    {
      // ---- take a candidate: the one whose turn it is if nobody has it, else the first free one behind the owned candidates' rectangles
      int x = -1;

      for (uint32_t spins = 0;; spins++)
      {
        const int d = ts->decidedX;
        const int turn = d < a.BX ? wave_next_candidate(candRow, used_word, skipped, nWords, d, a.BX, stage, lane) : a.BX;

        if (turn >= a.BX)
        {
          x = a.BX; // no live candidate is undecided: the row is complete
          break;
        }

        if (!((ts->owned[turn >> 5] >> (turn & 31)) & 1u) && try_own(turn))
        {
          x = turn;
          break;
        }

        int h = lane < 8 ? ts->hint[lane] : 0;
        h = max(__reduce_max_sync(0xFFFFFFFFu, h), turn + 1);
        int c = h < a.BX ? wave_next_candidate(candRow, used_word, skipped_or_owned, nWords, h, a.BX, stage, lane) : a.BX;

        if (c >= a.BX && h > turn + 1)
          c = wave_next_candidate(candRow, used_word, skipped_or_owned, nWords, turn + 1, a.BX, stage, lane);

        if (c < a.BX && try_own(c))
        {
          x = c;
          break;
        }

        if (c >= a.BX)
        {
          // every live candidate has an owner: wait for the row to move on
          if (spins > LIMG_WAVE_SPIN_LIMIT) { a.flags[3] = 1; x = a.BX; break; }

          __nanosleep(100);
        }
      }

      if (x >= a.BX)
        break;

      // ---- the candidate's data (nobody waits for this warp yet, unless it is the candidate's turn already)
      const SeedLinks links = scan.prefetch_links(x, y, stage);

      {
        // tell the team where this candidate's rectangle will probably end (the mask-free growth inside the 8 x 8 word is an upper bound)
        const int urx = (int)(links.u & 0xFFu), ury = (int)((links.u >> 8) & 0xFFu);
        const bool emits = stage == 0 ? (urx >= 3 && ury >= 3) : (urx * ury > 1);

        if (lane == 0 && !(a.experiment & 4))
          tsp->hint[wit] = x + (emits ? max(urx, 1) : 1);
      }

      SeedPre pre = scan.prefetch_bitmaps(links, y, stage);
      const int safeAfter = (stage == 0 && a.safe) ? (int)(links.safe >> 16) : LIMG_WAVE_DONE;
      int safeMine = LIMG_WAVE_DONE;

      if (stage == 0 && a.safe)
      {
        // the candidate's own safe column, tightened from the live mask (see wave_scan_rows; bits set later only tighten it further)
        const uint32_t col = (links.w0 & 1u) | ((links.w0 >> 7) & 2u) | ((links.w0 >> 14) & 4u) | ((links.w0 >> 21) & 8u) | ((links.w1 & 1u) << 4) | ((links.w1 >> 3) & 32u) |
                             ((links.w1 >> 10) & 64u) | ((links.w1 >> 17) & 128u);
        const int colRun = __ffs((int)(~col & 0x1FFu)) - 1;
        int blocked = 0x7FFF;

        if (!(colRun == 8 && y + 8 < a.BY) && lane < colRun / 3 && y + 1 + lane < a.BY)
        {
          const int cy = y + 1 + lane, w = x >> 5;
          uint32_t bits = be.used_word(cy, w) & (0xFFFFFFFFu >> (31 - (x & 31)));
          blocked = -1;

          if (bits)
            blocked = w * 32 + 32 - __clz((int)bits);
          else if (w > 0 && (bits = be.used_word(cy, w - 1)) != 0)
            blocked = (w - 1) * 32 + 32 - __clz((int)bits);
        }

        const int dyn = __reduce_min_sync(0xFFFFFFFFu, blocked);
        safeMine = min(max((int)(links.safe & 0xFFFFu), dyn == 0x7FFF ? -1 : dyn), safeAfter);
      }

      uint32_t first = 0, count = 0;
      int nextX = x + 1;
      bool claimed = false, holding = false, arrived = false, dead = false, nothing = false, expanded = false;

      for (int k = 0;; k++)
      {
        WaveResult r;
        Snapshot used;
        bool have = false, unstable = false;

        for (uint32_t spins = 0;; spins++)
        {
          int p = LIMG_WAVE_DONE;

          if (y > 0)
            p = wave_rows_above(be, stage, y, lane);

          const int below = stage == 1 ? (int)be.done_rows(0) : a.BY;

          // whose turn is it? (the mask first, then the flag: a claim that is visible has its flag set or its bookkeeping complete)
          bool myTurn = holding;

          if (!holding)
          {
            const int d = ts->decidedX;
            const int turn = wave_next_candidate(candRow, used_word, skipped, nWords, d, a.BX, stage, lane);

            if (turn > x)
            {
              dead = true; // in use by now (a rectangle of this row or of a row above covers it)
              break;
            }

            __threadfence_block();
            myTurn = turn == x && ts->committing == 0 && ts->decidedX == d;
          }

          if (myTurn && !arrived)
          {
            // every candidate left of x is decided: say so to the rows below
            arrived = true;

            if (lane == 0)
              tsp->rowSafe = safeMine;

            __syncwarp();
            publish(x);
          }

          if (p < min(x + 1 + a.margin - a.specAhead, a.BX))
          {
            // the rows above are still far away: whatever the mask shows now is not worth expanding against
            if (spins > LIMG_WAVE_SPIN_LIMIT) { a.flags[3] = 1; dead = true; break; }

            rowPolls++;
            if (p + 64 < x) __nanosleep(200);
            continue;
          }

          const bool far = have ? p >= min(r.boxR + a.margin, a.BX) && below >= min(r.boxD + a.stageGap / 2, a.BY) : p >= min(x + 1 + a.margin + 8, a.BX);
          const bool final = myTurn && far;

          if (final)
            be.acquire_fence();

          const Snapshot sn = scan.snapshot(x, y); // after the progress read

          if (scan.snap_used(sn, x)) { dead = true; break; }

          if (!have || unstable || !scan.same_where_probed(sn, used, r, x, y))
          {
            if (have) rowReexp++;
            if (myTurn) nTurnExp++;

            r = scan.expand(x, y, stage, pre, sn);
            unstable = scan.volatileReads;
            used = sn;
            have = true;
            rowExp++;
            expanded = true;

            if (lane == 0 && !(a.experiment & 4))
              tsp->hint[wit] = r.kind == 0 ? x + 1 : x + r.rx;
          }

          // A seed that emits nothing now never will (in-use bits are only ever set, so its rectangle can only shrink): decided, whoever's turn it is.
          if (r.kind == 0)
          {
            nothing = true;
            break;
          }

          if (final && p >= min(r.boxR + a.margin, a.BX) && below >= min(r.boxD + a.stageGap / 2, a.BY))
            break;

          if (spins > LIMG_WAVE_SPIN_LIMIT) { a.flags[3] = 1; dead = true; break; }

          if (myTurn && p >= min(r.boxR + a.margin, a.BX) && below >= min(r.boxD + a.stageGap / 2, a.BY))
            continue; // far enough, but this look was not fenced: look again at once

          rowPolls++;
          if (!myTurn || p + 64 < x) __nanosleep(myTurn ? 200 : 40);
        }

        if (dead || nothing)
          break;

        const int eox = r.kind == 2 ? r.cox : x, eoy = r.kind == 2 ? r.coy : y;
        const int erx = r.kind == 2 ? r.crx : r.rx, ery = r.kind == 2 ? r.cry : r.ry;
        const uint32_t T = base + ((uint32_t)(y * a.BX + x) << 3) + (uint32_t)min(k, LIMG_WAVE_MAX_ATTEMPTS - 1);

        if (k >= LIMG_WAVE_MAX_ATTEMPTS)
          failed = true; // more regrowths from one seed than the time stamp encodes: let the sequential pass do it

        if (!holding)
        {
          holding = true;
          first = count = ts->count;

          if (lane == 0)
            tsp->committing = 1;

          __threadfence_block();
          __syncwarp();
        }

        be.claim(eox, eoy, erx, ery, lane);

        // hand over to the rows below as early as possible: a right/down rectangle decides every seed up to its right edge
        if (r.kind == 1 && x + r.rx < a.BX)
        {
          if (lane == 0)
            tsp->rowSafe = safeAfter;

          __syncwarp();
          publish(x + r.rx);
        }

        for (int e = lane; e < erx * ery; e += 32)
          a.tau[(size_t)(eoy + e / erx) * a.BX + eox + e % erx] = T;

        if (lane == 0)
        {
          if (count < (uint32_t)a.listCap)
            list[count] = pack_rect(eox, eoy, erx, ery);
          else
            a.flags[4] = 1;
        }

        count++;
        claimed = true;

        if (r.kind == 2)
        {
          // limg.cpp:1435-1438: the scan resumes at the same seed; it can only emit again if the regrowth stays clear of its 3 x 3 corner
          if (!(x < eox + erx && x + 3 > eox && y < eoy + ery && y + 3 > eoy))
            continue;

          break;
        }

        nextX = x + r.rx;
        break;
      }

      if (claimed && lane == 0)
        emitInfo[(size_t)y * a.BX + x] = (first << 8) | (count - first);

      if (lane == 0)
        tsp->hint[wit] = 0;

      if (holding)
      {
        if (rowT && lane == 0)
        {
          const uint32_t tn = global_ns();
          if (rowT[1] == 0) rowT[1] = tn;
          rowT[2] = tn;
        }

        // the turn passes on: bookkeeping first, then the flag
        if (lane == 0)
        {
          tsp->count = count;
          tsp->rowSafe = safeAfter;
          tsp->decidedX = nextX;
        }

        __syncwarp();

        if (nextX < a.BX)
          publish(nextX);

        __threadfence_block();

        if (lane == 0)
          tsp->committing = 0;

        __syncwarp();
      }
      else if (nothing)
      {
        if (lane == 0)
          atomicOr(&tsp->skip[x >> 5], 1u << (x & 31));

        __syncwarp();
      }
      else if (dead)
      {
        nWasted += expanded ? 1 : 0;
      }
    }

    // every warp of the team is through with the row
    team_sync();

    if (wit == 0)
    {
      if (lane == 0)
        a.rowCounts[(size_t)stage * a.BY + y] = min(ts->count, (uint32_t)a.listCap);

      if (rowT && lane == 0)
        rowT[3] = global_ns();

      be.publish_max(stage, y, LIMG_WAVE_DONE, lane);
      wave_advance_done(be, stage, a.BY, lane);
    }

    if (stage) { nOnDemand1 += scan.nOnDemand - onDemandBefore; nExp1 += rowExp; nReexp1 += rowReexp; nPolls1 += rowPolls; }
    else { nOnDemand0 += scan.nOnDemand - onDemandBefore; nExp0 += rowExp; nReexp0 += rowReexp; nPolls0 += rowPolls; }
  }

  if (failed && lane == 0)
  {
    a.flags[5] |= 4u << (4 * attempt);
    a.flags[0] = (uint32_t)attempt + 1;
  }

  if (a.stats && lane == 0)
  {
    atomicAdd(&a.stats[0], nExp0); atomicAdd(&a.stats[1], nReexp0); atomicAdd(&a.stats[2], nPolls0); atomicAdd(&a.stats[3], nOnDemand0);
    atomicAdd(&a.stats[4], nExp1); atomicAdd(&a.stats[5], nReexp1); atomicAdd(&a.stats[6], nPolls1); atomicAdd(&a.stats[7], nOnDemand1);

    if (a.dbg)
    {
      atomicAdd(&a.dbg[240], nWasted);  // candidates expanded ahead of their turn that were covered before it came
      atomicAdd(&a.dbg[241], nTurnExp); // expansions that ran while it was the candidate's turn (on the row's critical path)
    }
  }
}

// dynamic shared memory of k_merge_cta in 32-bit words
__host__ __device__ inline size_t merge_cta_smem_words(int BY, int wordsPerRow) { return (size_t)BY * wordsPerRow + 2 * (size_t)BY + 4 + LIMG_CTA_WARPS * 32; }

// One cluster (gridDim.x == cluster size), up to LIMG_CTA_WARPS warps per CTA (blockDim.x / 32). `attempt`: runs only if flags[0] == attempt (uniform over the cluster).
// `teamWarps`: warps that share a block row (1: one warp per row, wave_scan_rows; 2, 4, 8: wave_scan_team; divides the warps of the CTA, BX <= 1024).
template <int CH>
__global__ void __launch_bounds__(LIMG_CTA_WARPS * 32) k_merge_cta(const __grid_constant__ WaveArgs a, int attempt, int teamWarps)
{
  if (a.flags[0] != (uint32_t)attempt)
    return;

  extern __shared__ uint32_t sMem[];
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int usedWords = a.BY * a.wordsPerRow;
  uint32_t *sUsed = sMem, *sProgress = sMem + usedWords, *sMisc = sProgress + 2 * a.BY, *sScratch = sMisc + 4;

  for (int i = threadIdx.x; i < usedWords + 2 * a.BY + 4; i += blockDim.x)
    sMem[i] = 0;

  // nobody writes into a replica before it is initialised
  cluster.sync();

  const WaveCluster be{ a, smem_u32(sUsed), smem_u32(sProgress), smem_u32(sMisc), cluster_size() };

  if (teamWarps > 1)
  {
    __shared__ TeamState sTeams[LIMG_CTA_WARPS / 2];
    const int warp = threadIdx.x >> 5, teamsPerCta = (int)(blockDim.x >> 5) / teamWarps;
    wave_scan_team<CH>(a, be, attempt, &sTeams[warp / teamWarps], teamWarps, warp % teamWarps, 1 + warp / teamWarps, max(1, (int)gridDim.x * teamsPerCta / 2), sScratch + warp * 32);
  }
  else
  {
    wave_scan_rows<CH>(a, be, attempt, 0, sScratch + (threadIdx.x >> 5) * 32);
  }

  // every row is done (its claims were fenced when they were made) and no CTA leaves while another may still write into it
  cluster.sync();

  if (cluster_rank() == 0)
    for (int i = threadIdx.x; i < usedWords; i += blockDim.x)
      a.used[i] = sUsed[i];
}

} // namespace limg

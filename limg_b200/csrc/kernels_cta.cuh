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

#define LIMG_CTA_WARPS 8

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

// dynamic shared memory of k_merge_cta in 32-bit words
__host__ __device__ inline size_t merge_cta_smem_words(int BY, int wordsPerRow) { return (size_t)BY * wordsPerRow + 2 * (size_t)BY + 4 + LIMG_CTA_WARPS * 32; }

// One cluster (gridDim.x == cluster size), up to LIMG_CTA_WARPS warps per CTA (blockDim.x / 32). `attempt`: runs only if flags[0] == attempt (uniform over the cluster).
template <int CH>
__global__ void __launch_bounds__(LIMG_CTA_WARPS * 32) k_merge_cta(const __grid_constant__ WaveArgs a, int attempt)
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
  wave_scan_rows<CH>(a, be, attempt, 0, sScratch + (threadIdx.x >> 5) * 32);

  // every row is done (its claims were fenced when they were made) and no CTA leaves while another may still write into it
  cluster.sync();

  if (cluster_rank() == 0)
    for (int i = threadIdx.x; i < usedWords; i += blockDim.x)
      a.used[i] = sUsed[i];
}

} // namespace limg

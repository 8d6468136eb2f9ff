// k_astar_lane: 32 find_path searches per warp in lock step, one query per lane (the state
// machine is hbn_astar_lane.h).  Persistent one-warp blocks pull queries from the search list
// of k_fp_classify with a warp-aggregated atomic; a lane whose query ends takes the next one
// in the same iteration, so the lanes of a warp stay busy until the list is empty.
// Outputs: status word and corridor ring per query (the heap continues in HBM, so no query
// overflows to another tier); the funnel runs in k_fp_funnel.
#pragma once
#include <cuda_runtime.h>

#include "hbn_findpath.cuh"  // SearchArgs
#include "hbn_astar_lane.h"

namespace hbn {

struct LaneScratch {  // one slot per lane of the grid in each region
  char* tab;            // node tables, tabBytes each (zeroed at allocation, wiped every 31 queries)
  char* rec;            // node records, kLaneRecBytes each
  char* heap;           // heap entries beyond the shared levels, kLaneHeapBytes each
  uint32_t* gen;        // table generation of every lane slot (persists across launches)
  size_t tabBytes;
};

template <int TS>
__host__ __device__ constexpr size_t laneSharedBytes() { return static_cast<size_t>(TS) * 32 * 6; }

// TS: heap entries per lane in shared memory; MINB: resident one-warp blocks per SM the
// register allocation must allow; CH: links per load stage; V: heap code variant (LaneSearch).
// WPB: warps per block.  Every warp is on its own (no block-level synchronisation); blocks of more than
// one warp only exist because a block costs 1 KB of reserved shared memory: two-warp blocks leave room
// for 71 instead of 63 heap entries per lane at 16 warps per SM.
template <int TS, int CH, int V, int WPB = 1, int F = 0>
__device__ __forceinline__ void astarLaneBody(const NavView& nav, const SearchArgs& a, const LaneScratch& sc) {
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) char smemAll[];
  const int lane = threadIdx.x & 31;
  const unsigned warpId = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (warpId >= static_cast<unsigned>(a.numWarps)) return;  // the whole warp
  char* smem = smemAll + static_cast<size_t>(threadIdx.x >> 5) * laneSharedBytes<TS>();
  const uint32_t ltMask = (1u << lane) - 1u;
  // lane slots are numbered over the lanes that take queries: a batch spread over many warps
  // (laneLimit < 32) does not need search state for the idle lanes
  const int activeLanes = (a.laneLimit > 0 && a.laneLimit < 32) ? a.laneLimit : 32;
  const bool hasSlot = lane < activeLanes;
  const size_t slotId = static_cast<size_t>(warpId) * activeLanes + (hasSlot ? lane : 0);
  LaneSearch<32, TS, CH, V, F> s;
  s.K = reinterpret_cast<float*>(smem) + lane;
  s.S = reinterpret_cast<uint16_t*>(smem + static_cast<size_t>(TS) * 32 * 4) + lane;
  s.G = reinterpret_cast<LaneHeapEnt*>(sc.heap + slotId * kLaneHeapBytes);
  s.tab = reinterpret_cast<uint16_t*>(sc.tab + slotId * sc.tabBytes);
  s.rec = sc.rec + slotId * kLaneRecBytes;
  s.cv = nullptr;
  s.gen = hasSlot ? sc.gen[slotId] : 0u;
  // A batch smaller than the grid is spread over more warps (a.laneLimit lanes each): the lanes
  // of a warp diverge in the heap loops, so fewer queries per warp = fewer instructions per step
  // on the latency chain of a small batch.
  s.mode = hasSlot ? kLIdle : kLDone;
  s.q = 0; s.endG = 0; s.size = 0; s.nodeCount = 0; s.status = 0; s.xk = 0; s.xcur = 0;
  s.expanded = s.nLinks = s.nNeigh = 0;
  const uint32_t nWork = *a.workCount;
  const bool fastFail = a.fastFail != 0, allCorridors = a.allCorridors != 0;
  const uint32_t tabVec = static_cast<uint32_t>(sc.tabBytes / 16);

  for (;;) {
    // ---- idle lanes take the next queries -------------------------------------------------
    const uint32_t idle = __ballot_sync(FULL, s.mode == kLIdle);
    if (idle) {
      const int leader = __ffs(idle) - 1;
      uint32_t wi = 0;
      if (lane == leader) wi = atomicAdd(a.counter, static_cast<uint32_t>(__popc(idle)));
      wi = __shfl_sync(FULL, wi, leader) + static_cast<uint32_t>(__popc(idle & ltMask));
      const bool mine = s.mode == kLIdle;
      const bool take = mine && wi < nWork;
      if (mine && !take) s.mode = kLDone;
      // generations used up: the whole warp wipes the tables of those lanes
      uint32_t wipe = __ballot_sync(FULL, take && s.gen >= kLaneGenMax);
      while (wipe) {
        const int l = __ffs(wipe) - 1;
        wipe &= wipe - 1;
        uint4* t = reinterpret_cast<uint4*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(s.tab), l));
        for (uint32_t i = lane; i < tabVec; i += 32) t[i] = make_uint4(0u, 0u, 0u, 0u);
        if (lane == l) s.gen = 0;
      }
      __syncwarp();
      if (take) {
        const uint32_t q = a.work[wi];
        const uint32_t qs = a.startDiv > 1 ? q / static_cast<uint32_t>(a.startDiv) : q;
        const float sp[3] = {a.sPt[3 * static_cast<size_t>(qs)], a.sPt[3 * static_cast<size_t>(qs) + 1],
                             a.sPt[3 * static_cast<size_t>(qs) + 2]};
        const float ep[3] = {a.ePt[3 * static_cast<size_t>(q)], a.ePt[3 * static_cast<size_t>(q) + 1],
                             a.ePt[3 * static_cast<size_t>(q) + 2]};
        s.begin(nav, q, a.sG[qs], sp, a.eG[q], ep, a.corrVia + static_cast<size_t>(q) * kMaxPathPolys);
      }
    }
    if (__all_sync(FULL, s.mode == kLDone)) break;

    const int ev = s.step(nav, fastFail, allCorridors);
    if (ev != kLEvNone) {
      const uint32_t q = s.q;
      if (ev == kLEvFault) {
        atomicAdd(a.fault, 1u);
        a.fault[1] = q;
        a.fault[2] = 5u | (TS << 8);
        a.astat[q] = kDtFailure;
        a.fullLen[q] = 0;
      } else {
        a.astat[q] = s.status;
        a.fullLen[q] = s.xk;
      }
      if (a.workCtr) {
        atomicAdd(a.workCtr + 0, static_cast<unsigned long long>(s.expanded));
        atomicAdd(a.workCtr + 1, static_cast<unsigned long long>(s.nLinks));
        atomicAdd(a.workCtr + 2, static_cast<unsigned long long>(s.nNeigh));
        atomicAdd(a.workCtr + 6, 1ull);
      }
    }
  }
  if (hasSlot) sc.gen[slotId] = s.gen;
}

template <int TS, int MINB, int CH, int V = 1, int WPB = 1, int F = 0>
__global__ void __launch_bounds__(32 * WPB, MINB / WPB) k_astar_lane(NavView nav, SearchArgs a, LaneScratch sc) {
  astarLaneBody<TS, CH, V, WPB, F>(nav, a, sc);
}

}  // namespace hbn

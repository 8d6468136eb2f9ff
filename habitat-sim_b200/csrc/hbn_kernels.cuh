// CUDA kernels (sm_100a) of the batched navmesh query path.  The per-query algorithms live
// in hbn_query.h; this file decides how queries map onto the machine.
//
//  k_snap<W>       W lanes per point: BV nodes are scanned W at a time with coalesced 16 B
//                  loads, candidate polys evaluated W at a time (findNearestPoly).  Used for
//                  small batches; large ones go through the candidate-list pipeline of
//                  hbn_snap.cuh (thread per point walk, thread per candidate closest point).
//  k_findpath_w<..> one warp per query, pulled from an atomic work counter.  The first find_path
//                  mapping; kept selectable (HBN_FP_G=warp) and as the 2048-entry open-list
//                  tier of k_astar_g.  The default is k_astar_lane (hbn_astar_lane.cuh): one
//                  query per lane, search state in HBM.
//  k_wall<..>      same workspace scheme for findDistanceToWall's Dijkstra.
//  k_trystep_*     one thread per query (64-node BFS in local memory).
//  k_random<W>     W lanes per sample; both reservoir scans run lane-parallel.
//  k_random_near<W> get_random_navigable_point_near: the circle / island filter evaluated lane-parallel
//                  per try, area sums and draws replayed in poly order.
#pragma once
#include <cuda_runtime.h>
#include "hbn_query.h"
#include "hbn_snap.h"
#include "hbn_astar_warp.cuh"

namespace hbn {

constexpr float kPickExt[3] = {2.f, 4.f, 2.f};  // polyPickExt, PF.cpp:134

// ------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) k_snap(NavView nav, const float* __restrict__ pts,
                                              const int32_t* __restrict__ islands, int64_t n,
                                              float* __restrict__ out_pts,
                                              uint32_t* __restrict__ out_g,
                                              uint32_t* __restrict__ out_refs,
                                              int32_t* __restrict__ out_isl,
                                              uint8_t* __restrict__ out_nav, float maxYDelta,
                                              const uint32_t* __restrict__ todo) {
  __shared__ uint32_t queue[256 / W][2 * W];
  if (todo && *todo == 0u) return;  // fallback launch of the candidate-list pipeline (hbn_snap.cuh): not needed
  WarpGroup<W> grp;
  const int gInBlock = threadIdx.x / W;
  const int64_t groupsPerGrid = static_cast<int64_t>(gridDim.x) * (blockDim.x / W);
  const float ext[3] = {2.f, 4.f, 2.f};
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x / W) + gInBlock; q < n;
       q += groupsPerGrid) {
    const float c[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
    const int isl = islands ? islands[q] : -1;
    // lane 0 of the group finds the xz half-extent that is enough (a 0.1 m column walk, hbn_snap.h)
    float rxz = 0.f;
    if (grp.lane() == 0) rxz = snapRadius(nav, c, ext, isl);
    rxz = grp.shfl(rxz, 0);
    const Nearest r = findNearestPoly(nav, grp, c, ext, isl, queue[gInBlock], rxz);
    grp.sync();
    if (grp.lane() == 0) {
      const bool ok = r.g != kNoPoly;
      if (out_pts) {
        out_pts[3 * q] = ok ? r.pt[0] : nanF();
        out_pts[3 * q + 1] = ok ? r.pt[1] : nanF();
        out_pts[3 * q + 2] = ok ? r.pt[2] : nanF();
      }
      if (out_g) out_g[q] = r.g;
      if (out_refs) out_refs[q] = ok ? nav.polys[r.g].ref : 0u;
      if (out_isl) out_isl[q] = ok ? nav.polys[r.g].island : -1;
      if (out_nav) {  // isNavigable, PF.cpp:1814-1831
        bool navOk = ok;
        if (ok) {
          const float dx = c[0] - r.pt[0], dz = c[2] - r.pt[2];
          float d2 = 0.f;
          d2 += dx * dx;
          d2 += dz * dz;
          if (fabsf(r.pt[1] - c[1]) > maxYDelta || fsqrt(d2) > 1e-2f) navOk = false;
        }
        out_nav[q] = navOk ? 1 : 0;
      }
    }
  }
}

// Two independent projectToPoly batches in ONE launch (find_path's starts and ends, try_step's
// start and end): small batches are latency chains, and two half-empty launches back to back
// cost twice the chain.  Group q serves point q of job a, or point q - a.n of job b.
struct SnapJob {
  const float* pts;
  int64_t n;
  float* out_pts;     // nullable
  uint32_t* out_g;    // nullable
};

template <int W>
__global__ void __launch_bounds__(256) k_snap_dual(NavView nav, SnapJob a, SnapJob b) {
  __shared__ uint32_t queue[256 / W][2 * W];
  WarpGroup<W> grp;
  const int gInBlock = threadIdx.x / W;
  const int64_t groupsPerGrid = static_cast<int64_t>(gridDim.x) * (blockDim.x / W);
  const float ext[3] = {2.f, 4.f, 2.f};
  const int64_t n = a.n + b.n;
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x / W) + gInBlock; q < n; q += groupsPerGrid) {
    const bool first = q < a.n;
    const int64_t i = first ? q : q - a.n;
    const float* pts = first ? a.pts : b.pts;
    float* out_pts = first ? a.out_pts : b.out_pts;
    uint32_t* out_g = first ? a.out_g : b.out_g;
    const float c[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    float rxz = 0.f;
    if (grp.lane() == 0) rxz = snapRadius(nav, c, ext, -1);
    rxz = grp.shfl(rxz, 0);
    const Nearest r = findNearestPoly(nav, grp, c, ext, -1, queue[gInBlock], rxz);
    grp.sync();
    if (grp.lane() == 0) {
      const bool ok = r.g != kNoPoly;
      if (out_pts) {
        out_pts[3 * i] = ok ? r.pt[0] : nanF();
        out_pts[3 * i + 1] = ok ? r.pt[1] : nanF();
        out_pts[3 * i + 2] = ok ? r.pt[2] : nanF();
      }
      if (out_g) out_g[i] = r.g;
    }
  }
}

// ------------------------------------------------------------------------------------
// workspace placement
//   kWsShared : every array in shared memory
//   kWsHybrid : heap (hkey/hidx) + hash in shared memory, node arrays in global scratch
enum { kWsShared = 0, kWsHybrid = 1 };

HBN_HD size_t wsSharedBytes(int cap, int place) {
  return place == kWsShared ? astarWsBytes(cap) : static_cast<size_t>(cap) * (4 + 8 + 2);
}
HBN_HD size_t wsGlobalBytes(int cap, int place) {
  return place == kWsShared ? 0 : static_cast<size_t>(cap) * (5 * 4 + 4 + 4 + 2 + 2);
}

__device__ __forceinline__ AStarWs wsCarveHybrid(void* sm, void* gl, int cap) {
  AStarWs w;
  char* p = static_cast<char*>(gl);
  w.px = reinterpret_cast<float*>(p); p += 4 * cap;
  w.py = reinterpret_cast<float*>(p); p += 4 * cap;
  w.pz = reinterpret_cast<float*>(p); p += 4 * cap;
  w.cost = reinterpret_cast<float*>(p); p += 4 * cap;
  w.total = reinterpret_cast<float*>(p); p += 4 * cap;
  w.gid = reinterpret_cast<uint32_t*>(p); p += 4 * cap;
  w.lnk = reinterpret_cast<uint32_t*>(p); p += 4 * cap;
  w.pidx = reinterpret_cast<uint16_t*>(p); p += 2 * cap;
  w.hpos = reinterpret_cast<uint16_t*>(p); p += 2 * cap;
  char* s = static_cast<char*>(sm);
  w.hkey = reinterpret_cast<float*>(s); s += 4 * cap;
  w.hash = reinterpret_cast<uint32_t*>(s); s += 8 * cap;
  w.hidx = reinterpret_cast<uint16_t*>(s); s += 2 * cap;
  w.cap = cap;
  w.hashMask = 2 * cap - 1;
  return w;
}

struct FindPathArgs {
  const float* starts;    // requested points
  const float* ends;
  const uint32_t* sG;     // projectToPoly results
  const float* sPt;
  const uint32_t* eG;
  const float* ePt;
  int64_t n;              // total queries (when work == nullptr)
  const uint32_t* work;   // optional list of query indices
  const uint32_t* workCount;
  uint32_t* counter;      // atomic work cursor
  uint32_t* overflow;     // queries that outgrew this tier
  uint32_t* overflowCount;
  float* out_dist;
  int32_t* out_npts;
  float* out_pts;
  int max_pts;
  uint32_t* out_corridor;
  int32_t* out_ncorridor;
  uint32_t* out_status;
  char* scratch;          // global workspace, one slot per warp of the grid (hybrid)
  int fastFail;
  // optional work counters (HBN_FP_COUNT_WORK): [0] expanded polys, [1] their links,
  // [2] their non-null neighbours, [3] corridor polys, [4] corridor links, [5] path points,
  // [6] queries that ran A*, [7] queries
  unsigned long long* workCtr;
  int startDiv;           // > 1: query q starts at point q / startDiv (multi-goal: [n, g] pairs)
  // mapped host memory: [0] number of watchdog trips (kernel bugs), [1] a query index, [2] where
  unsigned int* fault;
};

// One warp per query (hbn_astar_warp.cuh); WPB warps per block, each with its own shared
// table + heap and its own slot of the global node-record scratch.  OC = open-list capacity of
// this tier; a query whose open list outgrows it is appended to the overflow list and re-run
// from scratch by the next tier (the search is deterministic).
template <int OC, int WPB>
__global__ void __launch_bounds__(32 * WPB) k_findpath_w(NavView nav, FindPathArgs a) {
  extern __shared__ __align__(16) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t slot = static_cast<size_t>(blockIdx.x) * WPB + warp;
  const WarpWs<OC> ws = WarpWs<OC>::carve(smem + static_cast<size_t>(warp) * WarpWs<OC>::sharedBytes(),
                                          a.scratch + slot * WarpWs<OC>::globalBytes());
  // after the search the heap area holds the corridor, the table area the staged portals
  uint32_t* pathG = reinterpret_cast<uint32_t*>(ws.heap - 1);
  uint32_t* pathVia = pathG + kMaxPathPolys;
  PortalRec* staged = reinterpret_cast<PortalRec*>(ws.tab);
  static_assert((OC + 2) * sizeof(HeapEnt) >= 2 * kMaxPathPolys * 4, "corridor buffers");
  static_assert(kTabSize * 4 >= kMaxPathPolys * sizeof(PortalRec), "portal staging");
  const uint32_t total = a.work ? *a.workCount : static_cast<uint32_t>(a.n);
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(a.counter, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= total) break;
    const int64_t q = a.work ? a.work[wi] : wi;
    const long long tq0 = clock64();
    const int64_t qs = a.startDiv > 1 ? q / a.startDiv : q;
    const uint32_t sG = a.sG[qs], eG = a.eG[q];
    const float rs[3] = {a.starts[3 * qs], a.starts[3 * qs + 1], a.starts[3 * qs + 2]};
    const float re[3] = {a.ends[3 * q], a.ends[3 * q + 1], a.ends[3 * q + 2]};
    const float sp[3] = {a.sPt[3 * qs], a.sPt[3 * qs + 1], a.sPt[3 * qs + 2]};
    const float ep[3] = {a.ePt[3 * q], a.ePt[3 * q + 1], a.ePt[3 * q + 2]};
    float* outPts = a.out_pts ? a.out_pts + static_cast<size_t>(q) * a.max_pts * 3 : nullptr;
    uint32_t* outCorr = a.out_corridor ? a.out_corridor + static_cast<size_t>(q) * kMaxPathPolys : nullptr;
    // findPathInternal, PF.cpp:1426-1468 (same decision sequence as hbn_query.h findPathInternal)
    float dist = infF();
    int npts = 0, ncorr = 0;
    uint32_t stA = 0, stS = 0;
    WarpSearch sr;
    sr.expanded = sr.links = sr.neighbours = 0;
    uint32_t corrLinks = 0;
    bool overflow = false;
    do {
      if (sG == kNoPoly || eG == kNoPoly) break;
      if (vfuzzyEq(sp, ep)) {  // PF.cpp:1434-1436
        dist = 0.f;
        npts = 2;
        if (outPts && lane == 0) {
          if (a.max_pts > 0) { outPts[0] = sp[0]; outPts[1] = sp[1]; outPts[2] = sp[2]; }
          if (a.max_pts > 1) { outPts[3] = ep[0]; outPts[4] = ep[1]; outPts[5] = ep[2]; }
        }
        break;
      }
      const int32_t si = nav.polys[sG].island, ei = nav.polys[eG].island;
      if (si < 0 || si != ei) break;  // hasConnection, PF.cpp:209-221
      int first = 0;                  // corridor = pathG[first .. first + ncorr)
      int fullLen = 1;
      if (sG == eG) {                 // DQ.cpp:996-1001
        if (lane == 0) { pathG[0] = sG; pathVia[0] = kNoPoly; }
        ncorr = 1;
        stA = kDtSuccess;
        __syncwarp();
      } else {
        if (!vfinite(sp) || !vfinite(ep)) { stA = kDtFailure | kDtInvalidParam; break; }
        sr = astarWarp<OC>(nav, ws, sG, eG, sp, ep, a.fastFail != 0);
        if (sr.status == kSearchOverflow) { overflow = true; break; }
        if (sr.status == kSearchWatchdog) {
          if (lane == 0) { atomicAdd(a.fault, 1u); a.fault[1] = static_cast<unsigned>(q); a.fault[2] = 1u | (sr.expanded & 0x40000000u) | (OC << 8); }
          stA = kDtFailure;
          break;
        }
        // getPathToNode, DQ.cpp:1167-1205: walk the parent chain from the end; the circular
        // buffer keeps the kMaxPathPolys polys nearest the start
        __syncwarp();
        if (lane == 0) {
          int k = 0;
          uint32_t cur = sr.lastBest;
          for (;;) {
            const uint4 nb = reinterpret_cast<const uint4*>(&ws.rec[cur])[1];
            const int idx = (kMaxPathPolys - 1 - k) & (kMaxPathPolys - 1);
            pathG[idx] = ws.tab[cur] & kNodeGMask;
            pathVia[idx] = nb.z;
            corrLinks += nb.y >> 27;
            k++;
            if (!nb.w) break;  // start node
            if (k > kTabSize) {  // a parent cycle would be a bug
              atomicAdd(a.fault, 1u); a.fault[1] = static_cast<unsigned>(q); a.fault[2] = 2u;
              break;
            }
            cur = nb.w - 1;
          }
          fullLen = k;
        }
        fullLen = __shfl_sync(0xffffffffu, fullLen, 0);
        __syncwarp();
        ncorr = fullLen < kMaxPathPolys ? fullLen : kMaxPathPolys;
        first = (kMaxPathPolys - fullLen) & (kMaxPathPolys - 1);
        stA = sr.status | ((fullLen > kMaxPathPolys) ? kDtBufferTooSmall : 0u);
      }
      if (outCorr)
        for (int i = lane; i < ncorr; i += 32)
          outCorr[i] = nav.polys[pathG[(first + i) & (kMaxPathPolys - 1)]].ref;
      if (stA != kDtSuccess || ncorr == 0) break;  // PF.cpp:1450
      // here fullLen <= kMaxPathPolys, so the corridor is contiguous from `first`
      for (int i = lane; i + 1 < ncorr; i += 32) staged[i] = nav.portals[pathVia[first + i + 1]];
      __syncwarp();
      if (lane == 0) {
        Funnel f;
        f.out = outPts;
        f.maxOut = a.max_pts;
        stS = funnelStraightPath(nav, rs, re, pathG + first, pathVia + first + 1, ncorr, f, staged);
        npts = f.count;
        if (stS == kDtSuccess && f.count != 0) dist = f.length;  // PF.cpp:1459
      }
      dist = __shfl_sync(0xffffffffu, dist, 0);
      npts = __shfl_sync(0xffffffffu, npts, 0);
      stS = __shfl_sync(0xffffffffu, stS, 0);
    } while (false);
    __syncwarp();
    if (lane == 0) {
      if (overflow) {
        const uint32_t o = atomicAdd(a.overflowCount, 1u);
        a.overflow[o] = static_cast<uint32_t>(q);
      } else {
        const bool found = dist < infF();
        a.out_dist[q] = dist;
        if (a.out_npts) a.out_npts[q] = found ? npts : 0;
        if (a.out_ncorridor) a.out_ncorridor[q] = ncorr;
        if (a.out_status) {
          a.out_status[2 * q] = stA;
          a.out_status[2 * q + 1] = stS;
        }
        {  // slowest query so far (debug aid): fault[4] = kilo-cycles, [5] = query, [6] = expansions
          const unsigned kc = static_cast<unsigned>((clock64() - tq0) >> 10);
          if (kc > a.fault[4] && atomicMax(a.fault + 4, kc) < kc) {
            a.fault[5] = static_cast<unsigned>(q);
            a.fault[6] = sr.expanded;
            a.fault[7] = OC;
          }
        }
        if (a.workCtr) {
          atomicAdd(a.workCtr + 0, static_cast<unsigned long long>(sr.expanded));
          atomicAdd(a.workCtr + 1, static_cast<unsigned long long>(sr.links));
          atomicAdd(a.workCtr + 2, static_cast<unsigned long long>(sr.neighbours));
          atomicAdd(a.workCtr + 3, static_cast<unsigned long long>(ncorr));
          atomicAdd(a.workCtr + 4, static_cast<unsigned long long>(corrLinks));
          atomicAdd(a.workCtr + 5, static_cast<unsigned long long>(found ? npts : 0));
          atomicAdd(a.workCtr + 6, static_cast<unsigned long long>(sr.expanded ? 1 : 0));
          atomicAdd(a.workCtr + 7, 1ull);
        }
      }
    }
    __syncwarp();
  }
}

// findPath(MultiGoalShortestPath&) goal loop (PF.cpp:1541-1569) over precomputed pair distances:
// one thread per start.  bounds / order: [n, g] scratch.  chosenEnd (nullable): the chosen
// goal's requested point (NaN if none) for the second find_path pass that produces the points.
__global__ void __launch_bounds__(128) k_multigoal_select(const float* __restrict__ starts,
                                                          const float* __restrict__ ends,
                                                          const uint32_t* __restrict__ sG,
                                                          const uint32_t* __restrict__ eG,
                                                          const float* __restrict__ pairDist, int64_t n,
                                                          int g, float* __restrict__ bounds,
                                                          int32_t* __restrict__ order,
                                                          float* __restrict__ out_dist,
                                                          int32_t* __restrict__ out_index,
                                                          float* __restrict__ chosenEnd) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float st[3] = {starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]};
  float d;
  int32_t idx;
  multiGoalSelect(g, st, sG[i] != kNoPoly, ends + i * g * 3, eG + i * g, pairDist + i * g,
                  bounds + i * g, order + i * g, &d, &idx);
  out_dist[i] = d;
  out_index[i] = idx;
  if (chosenEnd) {
    for (int k = 0; k < 3; ++k)
      chosenEnd[3 * i + k] = idx >= 0 ? ends[(i * g + idx) * 3 + k] : nanF();
  }
}

// Two-round goal pruning of the batched multi-goal query (hbn_query.h, kMultiGoalFirst).
// round 1: mask = the first kMultiGoalFirst goals of every start's sorted order
__global__ void __launch_bounds__(128) k_multigoal_round1(const float* __restrict__ starts,
                                                          const float* __restrict__ ends, int64_t n, int g,
                                                          float* __restrict__ bounds, int32_t* __restrict__ order,
                                                          uint8_t* __restrict__ mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float st[3] = {starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]};
  float* b = bounds + i * g;
  int32_t* o = order + i * g;
  for (int k = 0; k < g; ++k) {
    b[k] = g > 1 ? mnDist(&ends[(i * g + k) * 3], st) : 0.0f;
    o[k] = k;
    mask[i * g + k] = 0;
  }
  stdSortOrder(o, b, g);
  for (int k = 0; k < g && k < kMultiGoalFirst; ++k) mask[i * g + o[k]] = 1;
}
// round 2: mask = the later goals whose bound does not exceed the reference's running best
__global__ void __launch_bounds__(128) k_multigoal_round2(const uint32_t* __restrict__ sG,
                                                          const uint32_t* __restrict__ eG,
                                                          const float* __restrict__ pairDist, int64_t n, int g,
                                                          const float* __restrict__ bounds,
                                                          const int32_t* __restrict__ order,
                                                          uint8_t* __restrict__ mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = bounds + i * g;
  const int32_t* o = order + i * g;
  bool any = sG[i] != kNoPoly;
  if (any) {
    any = false;
    for (int k = 0; k < g; ++k) any = any || eG[i * g + k] != kNoPoly;
  }
  const float rb = any ? multiGoalRunningBest(g, kMultiGoalFirst, eG + i * g, pairDist + i * g, b, o) : -infF();
  for (int k = 0; k < g; ++k) mask[i * g + o[k]] = (k >= kMultiGoalFirst && !(b[o[k]] > rb)) ? 1 : 0;
}

// ------------------------------------------------------------------------------------
struct WallArgs {
  const uint32_t* sG;
  const float* sPt;
  int64_t n;
  const uint32_t* work;
  const uint32_t* workCount;
  uint32_t* counter;
  uint32_t* overflow;
  uint32_t* overflowCount;
  float maxRadius;
  float* out_pos;
  float* out_normal;
  float* out_dist;
  char* scratch;
};

template <int CAP, int PLACE>
__global__ void __launch_bounds__(128) k_wall(NavView nav, WallArgs a) {
  extern __shared__ __align__(16) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warpsPerBlock = blockDim.x >> 5;
  AStarWs w;
  if (PLACE == kWsShared) {
    w = astarWsCarve(smem + static_cast<size_t>(warp) * astarWsBytes(CAP), CAP);
  } else {
    const size_t slot = static_cast<size_t>(blockIdx.x) * warpsPerBlock + warp;
    w = wsCarveHybrid(smem + static_cast<size_t>(warp) * wsSharedBytes(CAP, PLACE),
                      a.scratch + slot * wsGlobalBytes(CAP, PLACE), CAP);
  }
  const uint32_t total = a.work ? *a.workCount : static_cast<uint32_t>(a.n);
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(a.counter, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= total) break;
    const int64_t q = a.work ? a.work[wi] : wi;
    for (int i = lane; i < 2 * CAP; i += 32) w.hash[i] = 0u;
    __syncwarp();
    if (lane == 0) {
      // closestObstacleSurfacePoint, PF.cpp:1794-1812
      float hp[3] = {0.f, 0.f, 0.f}, hn[3] = {0.f, 0.f, 0.f};
      float hd = infF();
      bool overflow = false;
      const uint32_t g = a.sG[q];
      if (g != kNoPoly) {
        const float c[3] = {a.sPt[3 * q], a.sPt[3 * q + 1], a.sPt[3 * q + 2]};
        hd = nanF();
        const uint32_t st = distanceToWall(nav, w, g, c, a.maxRadius, &hd, hp, hn);
        overflow = st == 0xffffffffu;
      }
      if (overflow) {
        const uint32_t o = atomicAdd(a.overflowCount, 1u);
        a.overflow[o] = static_cast<uint32_t>(q);
      } else {
        if (a.out_pos) { a.out_pos[3 * q] = hp[0]; a.out_pos[3 * q + 1] = hp[1]; a.out_pos[3 * q + 2] = hp[2]; }
        if (a.out_normal) { a.out_normal[3 * q] = hn[0]; a.out_normal[3 * q + 1] = hn[1]; a.out_normal[3 * q + 2] = hn[2]; }
        a.out_dist[q] = hd;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// tryStep phase A / B (PF.cpp:1575-1722), one thread per query
__global__ void __launch_bounds__(128) k_trystep_a(NavView nav, const float* __restrict__ ends,
                                                   const uint32_t* __restrict__ sG,
                                                   const float* __restrict__ sPt,
                                                   const uint32_t* __restrict__ eG, int64_t n,
                                                   int allowSliding, float* __restrict__ endPoint,
                                                   uint32_t* __restrict__ lastPoly) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const float sp[3] = {sPt[3 * q], sPt[3 * q + 1], sPt[3 * q + 2]};
  const float en[3] = {ends[3 * q], ends[3 * q + 1], ends[3 * q + 2]};
  float ep[3];
  uint32_t last = kNoPoly;
  const bool ok = tryStepPhaseA(nav, sG[q], sp, eG[q], en, allowSliding != 0, ep, &last);
  endPoint[3 * q] = ok ? ep[0] : nanF();
  endPoint[3 * q + 1] = ok ? ep[1] : nanF();
  endPoint[3 * q + 2] = ok ? ep[2] : nanF();
  lastPoly[q] = ok ? last : kNoPoly;
}

__global__ void __launch_bounds__(256) k_trystep_b(NavView nav, const float* __restrict__ starts,
                                                   const uint32_t* __restrict__ sG,
                                                   const uint32_t* __restrict__ e2G,
                                                   const uint32_t* __restrict__ lastPoly,
                                                   const float* __restrict__ endPoint, int64_t n,
                                                   float* __restrict__ out) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t last = lastPoly[q];
  if (last == kNoPoly) {  // tryStep returned `start` unchanged
    out[3 * q] = starts[3 * q];
    out[3 * q + 1] = starts[3 * q + 1];
    out[3 * q + 2] = starts[3 * q + 2];
    return;
  }
  float ep[3] = {endPoint[3 * q], endPoint[3 * q + 1], endPoint[3 * q + 2]};
  tryStepPhaseB(nav, sG[q], e2G[q], last, ep);
  out[3 * q] = ep[0];
  out[3 * q + 1] = ep[1];
  out[3 * q + 2] = ep[2];
}

// ------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) k_random(NavView nav, uint64_t seed, uint64_t query0,
                                                int64_t n, const int32_t* __restrict__ islands,
                                                int maxTries, float* __restrict__ out_pts,
                                                uint32_t* __restrict__ out_refs) {
  WarpGroup<W> grp;
  const int gInBlock = threadIdx.x / W;
  const int64_t groupsPerGrid = static_cast<int64_t>(gridDim.x) * (blockDim.x / W);
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x / W) + gInBlock; q < n;
       q += groupsPerGrid) {
    const int isl = islands ? islands[q] : -1;
    uint32_t draw = 0;
    uint32_t g = kNoPoly;
    float pt[3] = {nanF(), nanF(), nanF()};
    for (int t = 0; t < maxTries; ++t) {
      uint32_t used = 0;
      float p[3];
      g = findRandomPoint(nav, grp, seed, query0 + static_cast<uint64_t>(q), draw, isl, p, &used);
      draw += used;
      if (g != kNoPoly) {
        pt[0] = p[0]; pt[1] = p[1]; pt[2] = p[2];
        break;
      }
    }
    if (grp.lane() == 0) {
      out_pts[3 * q] = pt[0];
      out_pts[3 * q + 1] = pt[1];
      out_pts[3 * q + 2] = pt[2];
      if (out_refs) out_refs[q] = g != kNoPoly ? nav.polys[g].ref : 0u;
    }
  }
}

// getRandomNavigablePointInCircle (get_random_navigable_point_near), W lanes per sample
template <int W>
__global__ void __launch_bounds__(256) k_random_near(NavView nav, uint64_t seed, uint64_t query0, int64_t n,
                                                     const float* __restrict__ centers, float radius,
                                                     const int32_t* __restrict__ islands, int maxTries,
                                                     float* __restrict__ out_pts) {
  WarpGroup<W> grp;
  const int gInBlock = threadIdx.x / W;
  const int64_t groupsPerGrid = static_cast<int64_t>(gridDim.x) * (blockDim.x / W);
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x / W) + gInBlock; q < n;
       q += groupsPerGrid) {
    const float c[3] = {centers[3 * q], centers[3 * q + 1], centers[3 * q + 2]};
    float pt[3];
    randomPointInCircle(nav, grp, seed, query0 + static_cast<uint64_t>(q), islands ? islands[q] : -1, c, radius,
                        maxTries, pt);
    if (grp.lane() == 0) {
      out_pts[3 * q] = pt[0];
      out_pts[3 * q + 1] = pt[1];
      out_pts[3 * q + 2] = pt[2];
    }
  }
}

}  // namespace hbn

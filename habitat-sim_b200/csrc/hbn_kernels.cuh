// CUDA kernels (sm_100a) of the batched navmesh query path.  The per-query algorithms live
// in hbn_query.h; this file decides how queries map onto the machine.
//
//  k_snap<W>       W lanes per point: BV nodes are scanned W at a time with coalesced 16 B
//                  loads, candidate polys evaluated W at a time (findNearestPoly).  Used for
//                  small batches; large ones go through the candidate-list pipeline of
//                  hbn_snap.cuh (thread per point walk, thread per candidate closest point).
//  (find_path: hbn_findpath.cuh + hbn_astar_lane.cuh -- one query per lane, search state in HBM)
//  k_wall<..>      warp per query over a shared / hybrid A* workspace: findDistanceToWall's Dijkstra.
//  k_trystep_*     one thread per query (64-node BFS in local memory).
//  k_random<W>     W lanes per sample; both reservoir scans run lane-parallel.
//  k_random_near<W> get_random_navigable_point_near: the circle / island filter evaluated lane-parallel
//                  per try, area sums and draws replayed in poly order.
#pragma once
#include <cuda_runtime.h>
#include "hbn_query.h"
#include "hbn_snap.h"

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

// Up to three independent projectToPoly batches in ONE launch (find_path's starts and ends, try_step's
// start and end, the env step's positions, targets and goals): small batches are latency chains, and
// half-empty launches back to back cost the chain once each.  Group q serves point q of job a, then b, then c.
struct SnapJob {
  const float* pts;
  int64_t n;
  float* out_pts;     // nullable
  uint32_t* out_g;    // nullable
};

template <int W>
__global__ void __launch_bounds__(256) k_snap_dual(NavView nav, SnapJob a, SnapJob b, SnapJob c3) {
  __shared__ uint32_t queue[256 / W][2 * W];
  WarpGroup<W> grp;
  const int gInBlock = threadIdx.x / W;
  const int64_t groupsPerGrid = static_cast<int64_t>(gridDim.x) * (blockDim.x / W);
  const float ext[3] = {2.f, 4.f, 2.f};
  const int64_t n = a.n + b.n + c3.n;
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x / W) + gInBlock; q < n; q += groupsPerGrid) {
    const SnapJob& job = q < a.n ? a : (q < a.n + b.n ? b : c3);
    const int64_t i = q < a.n ? q : (q < a.n + b.n ? q - a.n : q - a.n - b.n);
    const float* pts = job.pts;
    float* out_pts = job.out_pts;
    uint32_t* out_g = job.out_g;
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

// projectToPoly of the points whose flag is set (the env step's fix-up: positions tryStep nudged).
template <int W>
__global__ void __launch_bounds__(256) k_snap_flagged(NavView nav, const float* __restrict__ pts,
                                                      const uint8_t* __restrict__ flag, int64_t n,
                                                      float* __restrict__ out_pts, uint32_t* __restrict__ out_g) {
  __shared__ uint32_t queue[256 / W][2 * W];
  WarpGroup<W> grp;
  const int gInBlock = threadIdx.x / W;
  const int64_t groupsPerGrid = static_cast<int64_t>(gridDim.x) * (blockDim.x / W);
  const float ext[3] = {2.f, 4.f, 2.f};
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * (blockDim.x / W) + gInBlock; q < n; q += groupsPerGrid) {
    if (!flag[q]) continue;  // uniform within the group
    const float c[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
    float rxz = 0.f;
    if (grp.lane() == 0) rxz = snapRadius(nav, c, ext, -1);
    rxz = grp.shfl(rxz, 0);
    const Nearest r = findNearestPoly(nav, grp, c, ext, -1, queue[gInBlock], rxz);
    grp.sync();
    if (grp.lane() == 0) {
      const bool ok = r.g != kNoPoly;
      out_pts[3 * q] = ok ? r.pt[0] : nanF();
      out_pts[3 * q + 1] = ok ? r.pt[1] : nanF();
      out_pts[3 * q + 2] = ok ? r.pt[2] : nanF();
      out_g[q] = r.g;
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

// Tier 0 of closestObstacleSurfacePoint: ONE QUERY PER THREAD (distanceToWallSmall, hbn_query.h).  The search
// is a handful of nodes, so its pool is 6 x kWallLaneCap words per thread in shared memory; the rare query
// that needs more goes to the warp-per-query tiers above through the overflow list.
constexpr int kWallLaneThreads = 128;
__global__ void __launch_bounds__(kWallLaneThreads) k_wall_lane(NavView nav, WallArgs a) {
  __shared__ uint32_t ws[6 * kWallLaneCap * kWallLaneThreads];
  const int64_t q = static_cast<int64_t>(blockIdx.x) * kWallLaneThreads + threadIdx.x;
  if (q >= a.n) return;
  // closestObstacleSurfacePoint, PF.cpp:1794-1812
  float hp[3] = {0.f, 0.f, 0.f}, hn[3] = {0.f, 0.f, 0.f};
  float hd = infF();
  const uint32_t g = a.sG[q];
  if (g != kNoPoly) {
    const float c[3] = {a.sPt[3 * q], a.sPt[3 * q + 1], a.sPt[3 * q + 2]};
    hd = nanF();
    const uint32_t st = distanceToWallSmall<kWallLaneCap, kWallLaneThreads>(nav, ws + threadIdx.x, g, c, a.maxRadius,
                                                                           &hd, hp, hn);
    if (st == 0xffffffffu) {
      a.overflow[atomicAdd(a.overflowCount, 1u)] = static_cast<uint32_t>(q);
      return;
    }
  }
  if (a.out_pos) { a.out_pos[3 * q] = hp[0]; a.out_pos[3 * q + 1] = hp[1]; a.out_pos[3 * q + 2] = hp[2]; }
  if (a.out_normal) { a.out_normal[3 * q] = hn[0]; a.out_normal[3 * q + 1] = hn[1]; a.out_normal[3 * q + 2] = hn[2]; }
  a.out_dist[q] = hd;
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

// The env step (try_step, then find_path from the new position: simulator.py:660-673 + habitat-lab's
// geodesic reward): phase B of tryStep that also hands find_path its start projection, so the new
// position is not projected a second time.  Three cases: tryStep returned `start` unchanged -> the
// projection of `start` made for phase A; the end point as phase A left it -> the projection of that
// point made for phase B (e2G, e2Pt); the end point nudged towards the poly centre (PF.cpp:1700-1719)
// -> flagged for k_snap_flagged.
__global__ void __launch_bounds__(256) k_envstep_b(NavView nav, const float* __restrict__ starts,
                                                   const uint32_t* __restrict__ sG, const float* __restrict__ sPt,
                                                   const uint32_t* __restrict__ e2G, const float* __restrict__ e2Pt,
                                                   const uint32_t* __restrict__ lastPoly,
                                                   const float* __restrict__ endPoint, int64_t n,
                                                   float* __restrict__ out, uint32_t* __restrict__ fpG,
                                                   float* __restrict__ fpPt, uint8_t* __restrict__ flag) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t last = lastPoly[q];
  if (last == kNoPoly) {
    out[3 * q] = starts[3 * q]; out[3 * q + 1] = starts[3 * q + 1]; out[3 * q + 2] = starts[3 * q + 2];
    fpG[q] = sG[q];
    fpPt[3 * q] = sPt[3 * q]; fpPt[3 * q + 1] = sPt[3 * q + 1]; fpPt[3 * q + 2] = sPt[3 * q + 2];
    flag[q] = 0;
    return;
  }
  const float e0[3] = {endPoint[3 * q], endPoint[3 * q + 1], endPoint[3 * q + 2]};
  float ep[3] = {e0[0], e0[1], e0[2]};
  tryStepPhaseB(nav, sG[q], e2G[q], last, ep);
  out[3 * q] = ep[0]; out[3 * q + 1] = ep[1]; out[3 * q + 2] = ep[2];
  const bool same = __float_as_uint(ep[0]) == __float_as_uint(e0[0]) && __float_as_uint(ep[1]) == __float_as_uint(e0[1]) &&
                    __float_as_uint(ep[2]) == __float_as_uint(e0[2]);
  fpG[q] = e2G[q];
  fpPt[3 * q] = e2Pt[3 * q]; fpPt[3 * q + 1] = e2Pt[3 * q + 1]; fpPt[3 * q + 2] = e2Pt[3 * q + 2];
  flag[q] = same ? 0 : 1;
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

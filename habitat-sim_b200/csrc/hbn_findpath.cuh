// find_path pipeline around the search (hbn_astar_lane.cuh): decisions that need no search, the work
// list, and -- after the search -- string pulling, path length and every per-query output.
//
//   k_fp_classify   thread per query: PF.cpp:1426-1447 (snap failed / different islands: no path;
//                   pathStart == pathEnd: trivial; same poly; else a search), + a cost class of the
//                   searches (straight-line distance) for the work list
//   k_fp_scatter    work list of the searches, longest class first (the expensive queries must not
//                   sit in the tail of the persistent search kernel)
//   k_fp_funnel     thread per query: findStraightPath + pathLength (PF.cpp:1456-1466) over the
//                   corridor the search left as a ring of entering links
#pragma once
#include <cuda_runtime.h>
#include "hbn_query.h"

namespace hbn {

constexpr uint32_t kSearchOverflow = 0xffffffffu;  // status of a query no search has answered yet

// classes of a find_path query (k_fp_classify), PF.cpp:1426-1447
enum : uint8_t {
  kClsNone = 0,      // a snap failed or the islands differ: no path
  kClsTrivial = 1,   // pathStart == pathEnd (PF.cpp:1434-1436)
  kClsSamePoly = 2,  // startRef == endRef (DQ.cpp:996-1001)
  kClsInvalid = 3,   // non-finite snapped point: findPath fails with INVALID_PARAM
  kClsSearch = 4,
  kClsSkip = 5       // masked out by the caller (multi-goal pruning): outputs are left alone
};

struct SearchArgs {
  const uint32_t* sG;   // projectToPoly results
  const float* sPt;
  const uint32_t* eG;
  const float* ePt;
  const uint32_t* work;       // queries that need a search (k_fp_classify)
  const uint32_t* workCount;
  uint32_t* counter;          // atomic work cursor
  uint32_t* astat;            // [n] findPath status word
  int32_t* fullLen;           // [n] untruncated corridor length (0 = not extracted)
  uint32_t* corrVia;          // [n, 256] corridor as entering links; element i at (first + i) & 255
  int startDiv;               // > 1: query q starts at point q / startDiv (multi-goal pairs)
  int fastFail;
  int allCorridors;           // extract the corridor of unsuccessful searches too
  unsigned long long* workCtr;
  unsigned int* fault;
  int laneLimit;              // k_astar_lane: lanes of a warp that take queries (0 = all 32); small batches spread over more warps
  int numWarps;               // k_astar_lane: warps of the grid that work (the last block may hold idle ones)
};

// The search list is ordered by expected cost: a persistent kernel finishes when its longest query does,
// and a cheap estimate of the cost (the straight-line distance between the snapped points, in
// kFpBuckets classes) is enough to keep the expensive ones out of the tail of the launch.
// Two passes: classify + histogram of the classes, then a scatter into per-class ranges.
constexpr int kFpBuckets = 32;
constexpr float kFpBucketWidth = 2.0f;  // metres of straight-line distance per class
constexpr uint8_t kFpNoBucket = 0xff;

__global__ void __launch_bounds__(256) k_fp_classify(NavView nav, const uint32_t* __restrict__ sG,
                                                     const float* __restrict__ sPt,
                                                     const uint32_t* __restrict__ eG,
                                                     const float* __restrict__ ePt, int64_t n, int startDiv,
                                                     const uint8_t* __restrict__ mask,
                                                     uint8_t* __restrict__ cls, uint8_t* __restrict__ bucket,
                                                     uint32_t* __restrict__ hist, uint32_t* __restrict__ workCount) {
  __shared__ uint32_t histS[kFpBuckets];
  if (threadIdx.x < kFpBuckets) histS[threadIdx.x] = 0;
  __syncthreads();
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q < n) {
    const int64_t qs = startDiv > 1 ? q / startDiv : q;
    const uint32_t s = sG[qs], e = eG[q];
    uint8_t c = kClsNone, b = kFpNoBucket;
    if (mask && !mask[q]) {
      c = kClsSkip;
    } else if (s != kNoPoly && e != kNoPoly) {
      const float sp[3] = {sPt[3 * qs], sPt[3 * qs + 1], sPt[3 * qs + 2]};
      const float ep[3] = {ePt[3 * q], ePt[3 * q + 1], ePt[3 * q + 2]};
      if (vfuzzyEq(sp, ep)) {
        c = kClsTrivial;
      } else {
        const int32_t si = nav.polys[s].island, ei = nav.polys[e].island;
        if (si >= 0 && si == ei) {  // hasConnection, PF.cpp:209-221
          if (s == e) c = kClsSamePoly;
          else if (!vfinite(sp) || !vfinite(ep)) c = kClsInvalid;
          else {
            c = kClsSearch;
            const float d = vdist(sp, ep) / kFpBucketWidth;
            b = static_cast<uint8_t>(d < static_cast<float>(kFpBuckets - 1) ? static_cast<int>(d) : kFpBuckets - 1);
            atomicAdd(&histS[b], 1u);
          }
        }
      }
    }
    cls[q] = c;
    bucket[q] = b;
  }
  __syncthreads();
  if (threadIdx.x < kFpBuckets && histS[threadIdx.x]) {
    atomicAdd(&hist[threadIdx.x], histS[threadIdx.x]);
    atomicAdd(workCount, histS[threadIdx.x]);
  }
}

// work[] = the search queries, class kFpBuckets-1 (longest) first.  cursor[]: zeroed.
__global__ void __launch_bounds__(256) k_fp_scatter(const uint8_t* __restrict__ bucket, int64_t n,
                                                    const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor,
                                                    uint32_t* __restrict__ work) {
  __shared__ uint32_t cntS[kFpBuckets], baseS[kFpBuckets];
  if (threadIdx.x < kFpBuckets) cntS[threadIdx.x] = 0;
  __syncthreads();
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint8_t b = q < n ? bucket[q] : kFpNoBucket;
  uint32_t r = 0;
  if (b != kFpNoBucket) r = atomicAdd(&cntS[b], 1u);
  __syncthreads();
  if (threadIdx.x < kFpBuckets && cntS[threadIdx.x]) {
    uint32_t start = 0;
    for (int j = kFpBuckets - 1; j > static_cast<int>(threadIdx.x); --j) start += hist[j];
    baseS[threadIdx.x] = start + atomicAdd(&cursor[threadIdx.x], cntS[threadIdx.x]);
  }
  __syncthreads();
  if (b != kFpNoBucket) work[baseS[b] + r] = static_cast<uint32_t>(q);
}

// Corridor as the search left it: entering links in a 256-entry ring; poly i is the neighbour
// its entering link leads to (poly 0 is the start poly).
struct ViaCorridor {
  const NavView& nav;
  const uint32_t* ring;  // [256]
  uint32_t first;
  uint32_t startG;
  __device__ __forceinline__ uint32_t via(int i) const { return ring[(first + i) & (kMaxPathPolys - 1)]; }
  __device__ __forceinline__ uint32_t poly(int i) const { return i == 0 ? startG : nav.links[via(i)].nei; }
  __device__ __forceinline__ uint32_t link(int i) const { return via(i + 1); }
  __device__ __forceinline__ void portal(int, uint32_t li, float* l, float* r) const {
    const float4 a = __ldg(reinterpret_cast<const float4*>(&nav.portals[li]));
    const float4 b = __ldg(reinterpret_cast<const float4*>(&nav.portals[li]) + 1);
    l[0] = a.x; l[1] = a.y; l[2] = a.z;
    r[0] = b.x; r[1] = b.y; r[2] = b.z;
  }
};

struct FpFunnelArgs {
  const float* starts;  // requested points (trap T2: the funnel uses these, not the snapped ones)
  const float* ends;
  const uint32_t* sG;
  const float* sPt;
  const uint32_t* eG;
  const float* ePt;
  const uint8_t* cls;
  const uint32_t* astat;
  const int32_t* fullLen;
  const uint32_t* corrVia;
  int64_t n;
  int startDiv;
  float* out_dist;
  int32_t* out_npts;
  float* out_pts;
  int max_pts;
  uint32_t* out_corridor;
  int32_t* out_ncorridor;
  uint32_t* out_status;
  unsigned long long* workCtr;
  int fillSkipped;  // kClsSkip queries: write an infinite distance (else nothing)
};

// findStraightPath + pathLength (PF.cpp:1456-1466) and every per-query output, one thread per
// query.
__global__ void __launch_bounds__(128) k_fp_funnel(NavView nav, FpFunnelArgs a) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= a.n) return;
  const int64_t qs = a.startDiv > 1 ? q / a.startDiv : q;
  const uint8_t c = a.cls[q];
  if (c == kClsSkip) {
    if (a.fillSkipped) a.out_dist[q] = infF();
    return;
  }
  float dist = infF();
  int npts = 0, ncorr = 0;
  uint32_t stA = 0, stS = 0, corrLinks = 0;
  float* outPts = a.out_pts ? a.out_pts + static_cast<size_t>(q) * a.max_pts * 3 : nullptr;
  uint32_t* outCorr = a.out_corridor ? a.out_corridor + static_cast<size_t>(q) * kMaxPathPolys : nullptr;
  if (c == kClsTrivial) {
    dist = 0.f;
    npts = 2;
    if (outPts) {
      if (a.max_pts > 0) { outPts[0] = a.sPt[3 * qs]; outPts[1] = a.sPt[3 * qs + 1]; outPts[2] = a.sPt[3 * qs + 2]; }
      if (a.max_pts > 1) { outPts[3] = a.ePt[3 * q]; outPts[4] = a.ePt[3 * q + 1]; outPts[5] = a.ePt[3 * q + 2]; }
    }
  } else if (c == kClsInvalid) {
    stA = kDtFailure | kDtInvalidParam;
  } else if (c == kClsSamePoly || c == kClsSearch) {
    const uint32_t sG = a.sG[qs];
    int fullLen = 1;
    stA = kDtSuccess;
    if (c == kClsSearch) {
      stA = a.astat[q];
      fullLen = a.fullLen[q];
    }
    ncorr = fullLen < kMaxPathPolys ? fullLen : kMaxPathPolys;
    ViaCorridor cor{nav, a.corrVia + static_cast<size_t>(q) * kMaxPathPolys,
                    static_cast<uint32_t>((kMaxPathPolys - fullLen) & (kMaxPathPolys - 1)), sG};
    if (outCorr || a.workCtr)
      for (int i = 0; i < ncorr; ++i) {
        const PolyRec* cp = &nav.polys[(c == kClsSamePoly) ? sG : cor.poly(i)];
        corrLinks += cp->linkCount;
        if (outCorr) outCorr[i] = cp->ref;
      }
    if (stA == kDtSuccess && ncorr > 0) {  // PF.cpp:1450
      const float rs[3] = {a.starts[3 * qs], a.starts[3 * qs + 1], a.starts[3 * qs + 2]};
      const float re[3] = {a.ends[3 * q], a.ends[3 * q + 1], a.ends[3 * q + 2]};
      Funnel f;
      f.out = outPts;
      f.maxOut = a.max_pts;
      stS = funnelStraightPathT(nav, rs, re, cor, ncorr, f);
      npts = f.count;
      if (stS == kDtSuccess && f.count != 0) dist = f.length;  // PF.cpp:1459
    }
  }
  const bool found = dist < infF();
  a.out_dist[q] = dist;
  if (a.out_npts) a.out_npts[q] = found ? npts : 0;
  if (a.out_ncorridor) a.out_ncorridor[q] = ncorr;
  if (a.out_status) {
    a.out_status[2 * q] = stA;
    a.out_status[2 * q + 1] = stS;
  }
  if (a.workCtr) {
    atomicAdd(a.workCtr + 3, static_cast<unsigned long long>(ncorr));
    atomicAdd(a.workCtr + 4, static_cast<unsigned long long>(corrLinks));
    atomicAdd(a.workCtr + 5, static_cast<unsigned long long>(found ? npts : 0));
    atomicAdd(a.workCtr + 7, 1ull);
  }
}

}  // namespace hbn

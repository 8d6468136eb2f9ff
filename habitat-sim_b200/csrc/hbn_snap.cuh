// Kernels of the candidate-list findNearestPoly (hbn_snap.h): count -> prefix sum -> fill ->
// eval -> select.  Everything is enqueued without a host round trip: the candidate total stays
// on the device, the per-candidate kernels run grid-stride up to it; if it exceeds the scratch
// capacity (an average of more than kSnapAvgCap candidates per point) the chunk is redone by
// the lane-group kernel k_snap<8>, which needs no scratch.
#pragma once
#include <cuda_runtime.h>

#include "hbn_snap.h"

namespace hbn {

constexpr int kSnapAvgCap = 128;  // candidate scratch: this many entries per point of a chunk

// rxz[q] = xz half-extent of the walk (snapRadius); cnt[q] = number of candidates of point q;
// best[q] = FLT_MAX
__global__ void __launch_bounds__(256) k_snap_count(NavView nav, const float* __restrict__ pts,
                                                    const int32_t* __restrict__ islands, int64_t n,
                                                    uint32_t* __restrict__ cnt, uint32_t* __restrict__ best,
                                                    float* __restrict__ rxz) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const float c[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
  const float ext[3] = {2.f, 4.f, 2.f};  // polyPickExt, PF.cpp:134
  const float r = snapRadius(nav, c, ext, islands ? islands[q] : -1);
  rxz[q] = r;
  cnt[q] = snapWalk(nav, c, ext, r, [](uint32_t, float) {});
  best[q] = 0x7f7fffffu;
}

// off[] = exclusive prefix sum of cnt[] (off[n] = total).  Nothing happens if the total exceeds cap.
__global__ void __launch_bounds__(256) k_snap_fill(NavView nav, const float* __restrict__ pts, int64_t n,
                                                   const uint32_t* __restrict__ off, uint32_t cap,
                                                   const float* __restrict__ rxz,
                                                   uint32_t* __restrict__ candG, uint32_t* __restrict__ candQ,
                                                   float* __restrict__ candLb) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n || off[n] > cap) return;
  const float c[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
  const float ext[3] = {2.f, 4.f, 2.f};
  uint32_t w = off[q];
  snapWalk(nav, c, ext, rxz[q], [&](uint32_t g, float lb) {
    candG[w] = g;
    candQ[w] = static_cast<uint32_t>(q);
    candLb[w] = lb;
    w++;
  });
}

// pass 0: the candidates with lower bound 0, best[q] = their minimum distance;
// pass 1: the others, unless their bound already exceeds best[q] (then they cannot win).
__global__ void __launch_bounds__(256) k_snap_eval(NavView nav, const float* __restrict__ pts,
                                                   const int32_t* __restrict__ islands, int64_t n,
                                                   const uint32_t* __restrict__ off, uint32_t cap,
                                                   const uint32_t* __restrict__ candG,
                                                   const uint32_t* __restrict__ candQ,
                                                   const float* __restrict__ candLb, int pass,
                                                   float* __restrict__ candD, uint32_t* __restrict__ best) {
  const uint32_t total = off[n];
  if (total > cap) return;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
    const float lb = candLb[c];
    if ((lb == 0.f) != (pass == 0)) continue;
    const uint32_t q = candQ[c];
    if (pass == 1 && !snapMayWin(lb, __uint_as_float(best[q]))) {
      candD[c] = -1.f;
      continue;
    }
    const float ctr[3] = {pts[3 * static_cast<size_t>(q)], pts[3 * static_cast<size_t>(q) + 1],
                          pts[3 * static_cast<size_t>(q) + 2]};
    SnapCandOut o;
    const float d = snapEval(nav, ctr, islands ? islands[q] : -1, candG[c], &o);
    candD[c] = d;
    if (pass == 0 && d >= 0.f && d < kFltMax) atomicMin(&best[q], __float_as_uint(d));  // d >= +0: bit order = value order
  }
}

// Outputs as k_snap writes them.  todo[0] is set when the chunk has to be redone by k_snap.
__global__ void __launch_bounds__(256) k_snap_select(NavView nav, const float* __restrict__ pts,
                                                     const int32_t* __restrict__ islands, int64_t n,
                                                     const uint32_t* __restrict__ off, uint32_t cap,
                                                     const uint32_t* __restrict__ candG,
                                                     const float* __restrict__ candD,
                                                     float* __restrict__ out_pts, uint32_t* __restrict__ out_g,
                                                     uint32_t* __restrict__ out_refs, int32_t* __restrict__ out_isl,
                                                     uint8_t* __restrict__ out_nav, float maxYDelta,
                                                     uint32_t* __restrict__ todo) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n) return;
  if (off[n] > cap) {
    if (q == 0) todo[0] = 1u;
    return;
  }
  if (q == 0) todo[0] = 0u;
  const uint32_t b = off[q], e = off[q + 1];
  const uint32_t w = snapSelect(candD, b, e);
  const bool ok = w < e;
  uint32_t g = kNoPoly;
  const float ctr[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
  float pt[3] = {0.f, 0.f, 0.f};
  if (ok) {  // the winner's closest point, recomputed (same operations, same bits)
    g = candG[w];
    SnapCandOut o;
    snapEval(nav, ctr, islands ? islands[q] : -1, g, &o);
    pt[0] = o.cp[0]; pt[1] = o.cp[1]; pt[2] = o.cp[2];
  }
  if (out_pts) {
    out_pts[3 * q] = ok ? pt[0] : nanF();
    out_pts[3 * q + 1] = ok ? pt[1] : nanF();
    out_pts[3 * q + 2] = ok ? pt[2] : nanF();
  }
  if (out_g) out_g[q] = g;
  if (out_refs) out_refs[q] = ok ? nav.polys[g].ref : 0u;
  if (out_isl) out_isl[q] = ok ? nav.polys[g].island : -1;
  if (out_nav) {  // isNavigable, PF.cpp:1814-1831
    bool navOk = ok;
    if (ok) {
      const float dx = ctr[0] - pt[0], dz = ctr[2] - pt[2];
      float d2 = 0.f;
      d2 += dx * dx;
      d2 += dz * dz;
      if (fabsf(pt[1] - ctr[1]) > maxYDelta || fsqrt(d2) > 1e-2f) navOk = false;
    }
    out_nav[q] = navOk ? 1 : 0;
  }
}

}  // namespace hbn

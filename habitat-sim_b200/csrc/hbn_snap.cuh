// Kernels of the candidate-list findNearestPoly (hbn_snap.h): walk -> eval x 2 -> mark -> select.
// Everything is enqueued without a host round trip.  ONE BV walk per point: a thread reserves candidate
// slots kSnapBlock at a time from a device counter while it walks (round 1 walked twice, count + fill
// around a prefix sum), so the candidates of a point are scattered blocks; the reference's "first strictly
// smaller distance in visit order" (DQ.cpp:670) becomes a 64-bit atomicMin of {distance bits, visit
// index} per point, and a pass over the slots finds whose key won.  If the candidates exceed the scratch
// (an average of more than kSnapAvgCap per point) or a point has 2^14 of them, the chunk is redone by the
// lane-group kernel k_snap<8>, which needs no scratch.
#pragma once
#include <cuda_runtime.h>

#include "hbn_snap.h"

namespace hbn {

constexpr int kSnapAvgCap = 128;  // candidate scratch: this many entries per point of a chunk
constexpr uint32_t kSnapBlock = 8;       // slots reserved per atomic
constexpr uint32_t kSnapQBits = 18;      // a chunk is at most 2^18 points (kSnapChunk): slot tag = point | visit index << 18
constexpr uint32_t kSnapSeqMax = 1u << (32 - kSnapQBits);
constexpr uint32_t kSnapNoSlot = 0xffffffffu;

struct SnapScratch {
  uint32_t* total;   // slots reserved so far (may run past cap: then todo is set)
  uint32_t* todo;    // != 0: the chunk has to be redone by k_snap
  uint32_t cap;      // slots in candG / candTag / candLb / candD
  uint32_t* candG;
  uint32_t* candTag;
  float* candLb;
  float* candD;
  unsigned long long* best;  // per point: distance bits << 32 | visit index of the best candidate so far
  uint32_t* winner;          // per point: its slot
};

__global__ void __launch_bounds__(256) k_snap_walk(NavView nav, const float* __restrict__ pts,
                                                   const int32_t* __restrict__ islands, int64_t n, SnapScratch sc) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n) return;
  sc.best[q] = kSnapBestInit;
  sc.winner[q] = kSnapNoSlot;
  const float c[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
  const float ext[3] = {2.f, 4.f, 2.f};  // polyPickExt, PF.cpp:134
  const int isl = islands ? islands[q] : -1;
  bool lost = false;
  // A point over a flat part of the mesh (nearly all of them) ends up with the smallest radius, kSnapRadiusMin:
  // its candidates are what a walk of THAT box collects.  So the walk that looks for a poly under the point
  // (snapRadius) uses that box and keeps what it finds: if the radius it arrives at is the minimum, these are
  // the candidates and the point is done after one walk instead of two.
  float r = ext[0];
  if (nav.bvXzTight) {
    float ub = kFltMax;
    uint32_t n1 = 0;
    const uint32_t base1 = atomicAdd(sc.total, kSnapBlock);
    const bool fits = base1 + kSnapBlock <= sc.cap;
    snapWalk(nav, c, ext, kSnapRadiusMin, [&](uint32_t g, float lb) {
      snapUbUpdate(nav, c, isl, g, lb, ub);
      if (n1 < kSnapBlock && fits) {
        sc.candG[base1 + n1] = g;
        sc.candTag[base1 + n1] = static_cast<uint32_t>(q) | (n1 << kSnapQBits);
        sc.candLb[base1 + n1] = lb;
      }
      n1++;
    });
    r = snapRadiusFromUb(ub, ext[0]);
    const bool done = fits && n1 <= kSnapBlock && !(r > kSnapRadiusMin);
    if (fits)  // the unused slots of the block -- all of them if a wider walk follows
      for (uint32_t k = done ? n1 : 0u; k < kSnapBlock; ++k) sc.candTag[base1 + k] = kSnapNoSlot;
    else
      lost = true;
    if (done) {
      if (lost) atomicOr(sc.todo, 1u);
      return;
    }
  }
  uint32_t count = 0, base = 0;
  snapWalk(nav, c, ext, r, [&](uint32_t g, float lb) {
    const uint32_t k = count & (kSnapBlock - 1u);
    if (k == 0u) base = atomicAdd(sc.total, kSnapBlock);
    if (base + kSnapBlock <= sc.cap && count < kSnapSeqMax) {
      sc.candG[base + k] = g;
      sc.candTag[base + k] = static_cast<uint32_t>(q) | (count << kSnapQBits);
      sc.candLb[base + k] = lb;
    } else {
      lost = true;
    }
    count++;
  });
  if ((count & (kSnapBlock - 1u)) != 0u && base + kSnapBlock <= sc.cap)  // the unused slots of the last block
    for (uint32_t k = count & (kSnapBlock - 1u); k < kSnapBlock; ++k) sc.candTag[base + k] = kSnapNoSlot;
  if (lost) atomicOr(sc.todo, 1u);
}

// pass 0: the candidates with lower bound 0; pass 1: the others, unless their bound already exceeds the
// best distance of the point (then they cannot win).  Both feed the point's minimum.
__global__ void __launch_bounds__(256) k_snap_eval(NavView nav, const float* __restrict__ pts,
                                                   const int32_t* __restrict__ islands, SnapScratch sc, int pass) {
  if (*sc.todo) return;
  const uint32_t total = *sc.total;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
    const uint32_t tag = sc.candTag[c];
    if (tag == kSnapNoSlot) continue;
    const float lb = sc.candLb[c];
    if ((lb == 0.f) != (pass == 0)) continue;
    const uint32_t q = tag & ((1u << kSnapQBits) - 1u);
    if (pass == 1 && !snapMayWin(lb, __uint_as_float(static_cast<uint32_t>(sc.best[q] >> 32)))) {
      sc.candD[c] = -1.f;
      continue;
    }
    const float ctr[3] = {pts[3 * static_cast<size_t>(q)], pts[3 * static_cast<size_t>(q) + 1],
                          pts[3 * static_cast<size_t>(q) + 2]};
    SnapCandOut o;
    const float d = snapEval(nav, ctr, islands ? islands[q] : -1, sc.candG[c], &o);
    sc.candD[c] = d;
    // "if (d < m_nearestDistanceSqr)" from FLT_MAX: d >= +0 (bit order = value order), never NaN or infinite
    if (d >= 0.f && d < kFltMax)
      atomicMin(&sc.best[q], (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | (tag >> kSnapQBits));
  }
}

// the slot whose {distance, visit index} is the point's minimum
__global__ void __launch_bounds__(256) k_snap_mark(SnapScratch sc) {
  if (*sc.todo) return;
  const uint32_t total = *sc.total;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
    const uint32_t tag = sc.candTag[c];
    if (tag == kSnapNoSlot) continue;
    const float d = sc.candD[c];
    if (!(d >= 0.f && d < kFltMax)) continue;
    const uint32_t q = tag & ((1u << kSnapQBits) - 1u);
    if (sc.best[q] == ((static_cast<unsigned long long>(__float_as_uint(d)) << 32) | (tag >> kSnapQBits))) sc.winner[q] = c;
  }
}

// Outputs as k_snap writes them.
__global__ void __launch_bounds__(256) k_snap_select(NavView nav, const float* __restrict__ pts,
                                                     const int32_t* __restrict__ islands, int64_t n, SnapScratch sc,
                                                     float* __restrict__ out_pts, uint32_t* __restrict__ out_g,
                                                     uint32_t* __restrict__ out_refs, int32_t* __restrict__ out_isl,
                                                     uint8_t* __restrict__ out_nav, float maxYDelta) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n || *sc.todo) return;
  const uint32_t w = sc.winner[q];
  const bool ok = w != kSnapNoSlot;
  uint32_t g = kNoPoly;
  const float ctr[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
  float pt[3] = {0.f, 0.f, 0.f};
  if (ok) {  // the winner's closest point, recomputed (same operations, same bits)
    g = sc.candG[w];
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

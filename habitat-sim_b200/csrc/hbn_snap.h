// findNearestPoly as a candidate-list pipeline (host + device code of its three steps).
//
// projectToPoly (PF.cpp:126-147) -> dtNavMeshQuery::findNearestPoly (DQ.cpp:702-730) visits the
// polys whose BV leaf overlaps the query box, in BV order, tile by tile (queryPolygons
// DQ.cpp:923-960, queryPolygonsInTile :732-847) and keeps the first strictly smaller distance
// (dtFindNearestPolyQuery::process :644-679).  The walk is cheap integer work; the closest-point
// arithmetic per candidate (closestPointOnPoly, DN.cpp:728-758) is 20x more instructions and
// differs from candidate to candidate.  Doing both in one thread group per point leaves the
// lanes diverged (measured: 8 of 32 threads active, 9.6 k warp instructions per point).
// So the work is cut where its shape changes:
//   1. snapWalk      one thread per POINT: the BV walk, counting / emitting candidates
//   2. snapEval      one thread per CANDIDATE: closest point + the reference's distance; first
//                    the candidates whose lower bound is 0, then those that may still win
//   3. snapSelect    one thread per POINT: first strict minimum in visit order
// with the candidates of all points in one compact array (offsets from a prefix sum).
#pragma once
#include "hbn_query.h"

namespace hbn {

// The polys findNearestPoly would hand to process(), in the reference's order.
// emit(g, lb) is called once per candidate; returns the number of candidates.
// lb is a LOWER BOUND of the distance process() will compute for the candidate, from the
// bounds of the poly's vertices and detail vertices (NavView::polyBox; the BV boxes will not do:
// their y range is not the poly's): 0 if the point can be over the poly within walkableClimb.  It
// lets snapEval skip candidates that cannot win -- on a multi-storey mesh the +-4 m pick box
// collects the polys of the storeys above and below, two thirds of all candidates.
HBN_HD float snapLowerBound(const NavView& nav, const float* c, uint32_t g, float climb) {
#if defined(__CUDA_ARCH__)
  const float4 lo = __ldg(reinterpret_cast<const float4*>(nav.polyBox) + 2 * static_cast<size_t>(g));
  const float4 hi = __ldg(reinterpret_cast<const float4*>(nav.polyBox) + 2 * static_cast<size_t>(g) + 1);
  const float bmin[3] = {lo.x, lo.y, lo.z}, bmax[3] = {hi.x, hi.y, hi.z};
#else
  const float* bmin = nav.polyBox + 8 * static_cast<size_t>(g);
  const float* bmax = bmin + 4;
#endif
  const float pad = 1e-3f, padY = 1e-3f;
  float dx = 0.f, dz = 0.f, dy = 0.f;
  if (c[0] < bmin[0] - pad) dx = (bmin[0] - pad) - c[0];
  else if (c[0] > bmax[0] + pad) dx = c[0] - (bmax[0] + pad);
  if (c[2] < bmin[2] - pad) dz = (bmin[2] - pad) - c[2];
  else if (c[2] > bmax[2] + pad) dz = c[2] - (bmax[2] + pad);
  if (c[1] < bmin[1] - padY) dy = (bmin[1] - padY) - c[1];
  else if (c[1] > bmax[1] + padY) dy = c[1] - (bmax[1] + padY);
  const float dxz = dx * dx + dz * dz;
  if (dxz > 0.f) return dxz + dy * dy;  // outside the poly in xz: plain 3D distance
  const float e = dy - climb;           // possibly over the poly: (|dy| - climb)^2
  return e > 0.f ? e * e : 0.f;
}

//
// rxz <= halfExt[0], halfExt[2] narrows the box in x and z.  The walk then visits a subset of the
// reference's candidates in the same relative order; snapRadius() picks rxz so that the subset
// still holds every candidate that can win.
template <class F>
HBN_HD uint32_t snapWalk(const NavView& nav, const float* center, const float* halfExt, float rxz, F&& emit) {
  if (!vfinite(center) || !vfinite(halfExt)) return 0;  // DQ.cpp:928-933
  float qmin[3], qmax[3];
  for (int k = 0; k < 3; ++k) {
    const float h = (k == 1) ? halfExt[k] : rxz;
    qmin[k] = center[k] - h;
    qmax[k] = center[k] + h;
  }
  // calcTileLoc, DN.cpp:1191-1195
  int minx = static_cast<int>(floorf((qmin[0] - nav.orig[0]) / nav.tileWidth));
  int miny = static_cast<int>(floorf((qmin[2] - nav.orig[2]) / nav.tileHeight));
  int maxx = static_cast<int>(floorf((qmax[0] - nav.orig[0]) / nav.tileWidth));
  int maxy = static_cast<int>(floorf((qmax[2] - nav.orig[2]) / nav.tileHeight));
  if (minx < nav.gridMinX) minx = nav.gridMinX;  // cells outside the grid hold no tiles
  if (miny < nav.gridMinY) miny = nav.gridMinY;
  if (maxx > nav.gridMinX + nav.gridW - 1) maxx = nav.gridMinX + nav.gridW - 1;
  if (maxy > nav.gridMinY + nav.gridH - 1) maxy = nav.gridMinY + nav.gridH - 1;
  uint32_t count = 0;
  for (int y = miny; y <= maxy; ++y) {
    for (int x = minx; x <= maxx; ++x) {
      const int cell = (y - nav.gridMinY) * nav.gridW + (x - nav.gridMinX);
      const uint32_t c0 = nav.gridStart[cell], c1 = nav.gridStart[cell + 1];
      for (uint32_t c = c0; c < c1; ++c) {
        const TileRec& tr = nav.tiles[nav.tileOrder[c]];
        if (tr.bvCount) {
          // quantised query box, DQ.cpp:749-765; [0] = the (narrowed) box of this walk, [1] = the
          // reference's full box, which decides for the loose leaves behind the tree
          uint16_t bmin[2][3], bmax[2][3];
          const float qfac = tr.bvQuantFactor;
          for (int k = 0; k < 3; ++k) {
            const float lo[2] = {qmin[k], center[k] - halfExt[k]}, hi[2] = {qmax[k], center[k] + halfExt[k]};
            for (int b = 0; b < 2; ++b) {
              const float mn = fclamp(lo[b], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
              const float mx = fclamp(hi[b], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
              bmin[b][k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * mn)) & 0xfffe);
              bmax[b][k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * mx + 1)) | 1);
            }
          }
          // DQ.cpp:768-800: pre-order array with escape indices
          const BvRec* node = &nav.bv[tr.bvStart];
          const BvRec* end = node + tr.bvCount;
          while (node < end) {
#if defined(__CUDA_ARCH__)
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(node));
            BvRec n;
            memcpy(&n, &raw, 16);
#else
            const BvRec n = *node;
#endif
            const bool leaf = n.i >= 0;
            const int which = (leaf && (n.i & kBvLooseBit) != 0) ? 1 : 0;
            const bool ov = overlapQuant(bmin[which], bmax[which], n.bmin, n.bmax);
            if (leaf && ov && (n.i & kBvFailBit) == 0) {
              const uint32_t g = static_cast<uint32_t>(n.i & kBvIndexMask);
              emit(g, snapLowerBound(nav, center, g, tr.walkableClimb));
              count++;
            }
            if (ov || leaf) node++;
            else node += -n.i;
          }
        } else {
          // no BV tree: linear scan with float bounds, DQ.cpp:806-842
          for (uint32_t idx = 0; idx < tr.polyCount; ++idx) {
            const uint32_t g = tr.polyStart + idx;
            const PolyRec* p = &nav.polys[g];
            if ((p->areaType >> 6) == 1 || (p->flags & kFlagWalk) == 0) continue;
            float pmin[3], pmax[3];
            vcopy(pmin, &p->v[0]);
            vcopy(pmax, &p->v[0]);
            for (int j = 1; j < p->nv; ++j)
              for (int k = 0; k < 3; ++k) {
                pmin[k] = p->v[j * 3 + k] < pmin[k] ? p->v[j * 3 + k] : pmin[k];
                pmax[k] = p->v[j * 3 + k] > pmax[k] ? p->v[j * 3 + k] : pmax[k];
              }
            bool ov = true;
            for (int k = 0; k < 3; ++k) ov = (qmin[k] > pmax[k] || qmax[k] < pmin[k]) ? false : ov;
            if (ov) {
              emit(g, snapLowerBound(nav, center, g, tr.walkableClimb));
              count++;
            }
          }
        }
      }
    }
  }
  return count;
}

// xz half-extent that is enough for the walk of a point (<= halfExt[0]).
// If some poly lies under the point -- found with a walk of a 0.1 m column -- the winner's
// distance is at most that poly's: d <= ub = (max |dy| over the poly's height range - climb)^2.
// Every candidate that can win or tie then has its closest point, hence its xz bounds, hence
// (bvXzTight) its BV leaf box within sqrt(ub) of the point in xz.  On a multi-storey tile the
// reference's +-2 m x +-4 m box collects ~70 candidates from ~350 BV nodes per point; the narrowed
// walk of an on-mesh point sees the handful of polys stacked under it.
// One candidate of the column walk: if the point lies inside poly g in xz, the winner's distance is at most
// (max |dy| over g's height range - climb)^2.
HBN_HD void snapUbUpdate(const NavView& nav, const float* center, int islandFilter, uint32_t g, float lb, float& ub) {
  if (lb >= ub) return;
  const PolyRec* p = &nav.polys[g];
  if (islandFilter >= 0 && p->island != islandFilter) return;
  if ((p->areaType >> 6) == 1 || !pointInPolygon(center, p->v, p->nv)) return;  // getPolyHeight would fail
  const float* b = nav.polyBox + 8 * static_cast<size_t>(g);
  const float dlo = fabsf(center[1] - b[1]), dhi = fabsf(center[1] - b[5]);
  const float e = (dlo > dhi ? dlo : dhi) - nav.tiles[p->tile].walkableClimb;
  const float u = e > 0.f ? e * e : 0.f;
  if (u < ub) ub = u;
}
constexpr float kSnapRadiusMin = 0.11f;  // two BV quanta (0.05 m cells) and rounding
HBN_HD float snapRadiusFromUb(float ub, float full) {
  if (!(ub < kFltMax)) return full;
  const float r = fsqrt(ub) * 1.001f + kSnapRadiusMin;
  return r < full ? r : full;
}
HBN_HD float snapRadius(const NavView& nav, const float* center, const float* halfExt, int islandFilter) {
  const float full = halfExt[0];
  if (!nav.bvXzTight) return full;
  float ub = kFltMax;
  snapWalk(nav, center, halfExt, 0.1f, [&](uint32_t g, float lb) { snapUbUpdate(nav, center, islandFilter, g, lb, ub); });
  return snapRadiusFromUb(ub, full);
}

struct SnapCandOut {
  float cp[3];
  uint32_t over;
};
// May a candidate with lower bound lb still win against the best distance seen so far?
// (margins: the bound and the distances are rounded float results)
HBN_HD bool snapMayWin(float lb, float best) { return lb * 0.999f - 1e-6f <= best; }

// dtFindNearestPolyQuery::process for one candidate (DQ.cpp:655-676): the distance it would
// compare, or a negative value if the island filter (PF.cpp:1729-1751) hides the poly.
HBN_HD float snapEval(const NavView& nav, const float* center, int islandFilter, uint32_t g, SnapCandOut* out) {
  const PolyRec* p = &nav.polys[g];
  if (islandFilter >= 0 && p->island != islandFilter) return -1.f;
  bool over;
  closestPointOnPoly(nav, p, center, out->cp, &over);
  out->over = over ? 1u : 0u;
  const float dx = center[0] - out->cp[0], dy = center[1] - out->cp[1], dz = center[2] - out->cp[2];
  float d;
  if (over) {
    d = fabsf(dy) - nav.tiles[p->tile].walkableClimb;
    d = d > 0 ? d * d : 0;
  } else {
    d = dx * dx + dy * dy + dz * dz;
  }
  return d;
}

// The same choice as a minimum: key = distance bits << 32 | visit index (distances are >= +0, so bit order is
// value order; the smallest visit index among equal distances is the first strictly smaller one).  "No
// candidate yet" is FLT_MAX as the distance, m_nearestDistanceSqr's start value (DQ.cpp:649), above every key.
constexpr unsigned long long kSnapBestInit = 0x7f7fffffffffffffull;

// "if (d < m_nearestDistanceSqr)" over the candidates in visit order (DQ.cpp:670).
// Returns the index of the winner in [begin, end) or end if none.  (A NaN distance never wins,
// as in the reference; negative = hidden by the island filter or skipped by its lower bound.)
HBN_HD uint32_t snapSelect(const float* d, uint32_t begin, uint32_t end) {
  float best = kFltMax;
  uint32_t win = end;
  for (uint32_t i = begin; i < end; ++i) {
    const float v = d[i];
    if (v >= 0.f && v < best) {
      best = v;
      win = i;
    }
  }
  return win;
}

}  // namespace hbn

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

template <class F>
HBN_HD uint32_t snapWalk(const NavView& nav, const float* center, const float* halfExt, F&& emit) {
  if (!vfinite(center) || !vfinite(halfExt)) return 0;  // DQ.cpp:928-933
  float qmin[3], qmax[3];
  for (int k = 0; k < 3; ++k) {
    qmin[k] = center[k] - halfExt[k];
    qmax[k] = center[k] + halfExt[k];
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
          // quantised query box, DQ.cpp:749-765
          uint16_t bmin[3], bmax[3];
          const float qfac = tr.bvQuantFactor;
          for (int k = 0; k < 3; ++k) {
            const float mn = fclamp(qmin[k], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
            const float mx = fclamp(qmax[k], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
            bmin[k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * mn)) & 0xfffe);
            bmax[k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * mx + 1)) | 1);
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
            const bool ov = overlapQuant(bmin, bmax, n.bmin, n.bmax);
            const bool leaf = n.i >= 0;
            if (leaf && ov && (n.i & kBvFailBit) == 0) {
              emit(static_cast<uint32_t>(n.i), snapLowerBound(nav, center, static_cast<uint32_t>(n.i), tr.walkableClimb));
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

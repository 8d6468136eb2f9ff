// Per-query navmesh algorithms over the flattened NavView, written once for device code
// (hbn_kernels.cu) and for the single-lane host build that tests/hostemu uses to debug
// them on a CPU-only box.  Every function cites the reference code whose results it must
// reproduce bit for bit (DQ = Detour/Source/DetourNavMeshQuery.cpp, DN =
// Detour/Source/DetourNavMesh.cpp, DNode = Detour/Source/DetourNode.cpp, PF =
// src/esp/nav/PathFinder.cpp).
//
// Cooperative functions take a lane-group `G` (WarpGroup<W> on the device: W lanes of a
// warp working on one query; HostGroup: one lane).  Serial functions are run by one lane.
#pragma once
#include "hbn_math.h"
#include "hbn_types.h"

namespace hbn {

// Detour status bits (Detour/Include/DetourStatus.h) reported in raw outputs
constexpr uint32_t kDtFailure = 1u << 31;
constexpr uint32_t kDtSuccess = 1u << 30;
constexpr uint32_t kDtInvalidParam = 1u << 3;
constexpr uint32_t kDtBufferTooSmall = 1u << 4;
constexpr uint32_t kDtOutOfNodes = 1u << 5;
constexpr uint32_t kDtPartialResult = 1u << 6;

constexpr int kMaxPathPolys = 256;   // MAX_POLYS, PF.cpp:1443
constexpr int kMaxNodes = 2048;      // navQuery_->init(navMesh, 2048), PF.cpp:937
constexpr float kHScale = 0.999f;    // H_SCALE, DQ.cpp:103

struct HostGroup {
  static constexpr int kWidth = 1;
  HBN_HD int lane() const { return 0; }
  HBN_HD void sync() const {}
  HBN_HD uint32_t ballot(bool p) const { return p ? 1u : 0u; }
  template <class T> HBN_HD T shfl(T v, int) const { return v; }
};

#if defined(__CUDACC__)
template <int W>
struct WarpGroup {
  static constexpr int kWidth = W;
  unsigned mask;
  int base;
  __device__ WarpGroup() {
    const int l = threadIdx.x & 31;
    base = l & ~(W - 1);
    mask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << base);
  }
  __device__ __forceinline__ int lane() const { return (threadIdx.x & 31) - base; }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ uint32_t ballot(bool p) const {
    return (__ballot_sync(mask, p) >> base) & ((W == 32) ? 0xffffffffu : ((1u << W) - 1u));
  }
  template <class T> __device__ __forceinline__ T shfl(T v, int src) const {
    return __shfl_sync(mask, v, src, W);
  }
};
#endif

HBN_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
HBN_HD int ffs32(uint32_t x) {  // 1-based index of lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
  return __ffs(x);
#else
  return __builtin_ffs(static_cast<int>(x));
#endif
}

// ---------------------------------------------------------------------------------------
// poly refs
// ---------------------------------------------------------------------------------------
// dtNavMesh::isValidPolyRef + decodePolyId (DN.cpp:1237-1248, DN.h:529-562) -> global index
HBN_HD uint32_t refToGlobal(const NavView& nav, uint32_t ref) {
  if (!ref) return kNoPoly;
  const uint32_t ip = ref & ((1u << nav.polyBits) - 1u);
  const uint32_t it = (ref >> nav.polyBits) & ((1u << nav.tileBits) - 1u);
  if (it >= nav.numTiles) return kNoPoly;
  const TileRec& t = nav.tiles[it];
  if (ip >= t.polyCount) return kNoPoly;
  if (((t.refBase ^ ref) >> (nav.polyBits + nav.tileBits)) != 0) return kNoPoly;  // salt
  return t.polyStart + ip;
}

// ---------------------------------------------------------------------------------------
// closest point on a poly: DN.cpp:621-758
// ---------------------------------------------------------------------------------------
HBN_HD const float* detailVertex(const NavView& nav, const PolyRec* p, int nv, int idx) {
  return idx < nv ? &p->v[idx * 3] : &nav.detVerts[static_cast<size_t>(p->detVertBase + (idx - nv)) * 3];
}

// closestPointOnDetailEdges<onlyBoundary>, DN.cpp:621-674
HBN_HD void closestPointOnDetailEdges(const NavView& nav, const PolyRec* p, bool onlyBoundary,
                                      const float* pos, float* closest) {
  const int nv = p->nv;
  const int ntri = p->detTriCount;
  float dmin = kFltMax;
  float tmin = 0;
  const float* pmin = nullptr;
  const float* pmax = nullptr;
  for (int i = 0; i < ntri; i++) {
    const unsigned char* tris = &nav.detTris[static_cast<size_t>(p->detTriBase + i) * 4];
    const int tf = tris[3];
    const int ANY_BOUNDARY_EDGE = (1 << 0) | (1 << 2) | (1 << 4);
    if (onlyBoundary && (tf & ANY_BOUNDARY_EDGE) == 0) continue;
    const int ti[3] = {tris[0], tris[1], tris[2]};
    const float* v[3];
    v[0] = detailVertex(nav, p, nv, ti[0]);
    v[1] = detailVertex(nav, p, nv, ti[1]);
    v[2] = detailVertex(nav, p, nv, ti[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int j = (k + 2) % 3;  // (k=0,j=2), (1,0), (2,1)
      if ((((tf >> (j * 2)) & 0x3) & 1) == 0 && (onlyBoundary || ti[j] < ti[k])) continue;
      float t;
      const float d = distPtSegSqr2D(pos, v[j], v[k], t);
      if (d < dmin) {
        dmin = d;
        tmin = t;
        pmin = v[j];
        pmax = v[k];
      }
    }
  }
  if (pmin) vlerp(closest, pmin, pmax, tmin);
  else vcopy(closest, pos);
}

// dtNavMesh::getPolyHeight, DN.cpp:677-726
HBN_HD bool polyHeight(const NavView& nav, const PolyRec* p, const float* pos, float* height) {
  if ((p->areaType >> 6) == 1) return false;
  const int nv = p->nv;
  if (!pointInPolygon(pos, p->v, nv)) return false;
  const int ntri = p->detTriCount;
  for (int j = 0; j < ntri; ++j) {
    const unsigned char* t = &nav.detTris[static_cast<size_t>(p->detTriBase + j) * 4];
    const float* v0 = detailVertex(nav, p, nv, t[0]);
    const float* v1 = detailVertex(nav, p, nv, t[1]);
    const float* v2 = detailVertex(nav, p, nv, t[2]);
    float h;
    if (closestHeightPointTriangle(pos, v0, v1, v2, h)) {
      *height = h;
      return true;
    }
  }
  float closest[3];
  closestPointOnDetailEdges(nav, p, false, pos, closest);
  *height = closest[1];
  return true;
}

// dtNavMesh::closestPointOnPoly, DN.cpp:728-758
HBN_HD void closestPointOnPoly(const NavView& nav, const PolyRec* p, const float* pos,
                               float* closest, bool* posOverPoly) {
  vcopy(closest, pos);
  if (polyHeight(nav, p, pos, &closest[1])) {
    *posOverPoly = true;
    return;
  }
  *posOverPoly = false;
  if ((p->areaType >> 6) == 1) {
    float t;
    distPtSegSqr2D(pos, &p->v[0], &p->v[3], t);
    vlerp(closest, &p->v[0], &p->v[3], t);
    return;
  }
  closestPointOnDetailEdges(nav, p, true, pos, closest);
}

// dtNavMeshQuery::getPolyHeight, DQ.cpp:591-621 (status only as bool)
HBN_HD bool queryPolyHeight(const NavView& nav, const PolyRec* p, const float* pos, float* height) {
  if (!(finitef(pos[0]) && finitef(pos[2]))) return false;
  if ((p->areaType >> 6) == 1) {
    float t;
    distPtSegSqr2D(pos, &p->v[0], &p->v[3], t);
    *height = p->v[1] + (p->v[4] - p->v[1]) * t;
    return true;
  }
  return polyHeight(nav, p, pos, height);
}

// dtNavMeshQuery::closestPointOnPolyBoundary, DQ.cpp:536-584 with
// dtDistancePtPolyEdgesSqr, DetourCommon.cpp:254-270: ed[j] is the distance to the edge
// v[j] -> v[(j+1) % nv].
HBN_HD void closestPointOnPolyBoundary(const PolyRec* p, const float* pos, float* closest) {
  const int nv = p->nv;
  float ed[kVertsPerPoly];
  float et[kVertsPerPoly];
  bool c = false;
#pragma unroll
  for (int j = 0; j < kVertsPerPoly; ++j) {
    if (j < nv) {
      const int i = (j + 1 == nv) ? 0 : j + 1;
      const float* vi = &p->v[i * 3];
      const float* vj = &p->v[j * 3];
      if (((vi[2] > pos[2]) != (vj[2] > pos[2])) &&
          (pos[0] < (vj[0] - vi[0]) * (pos[2] - vi[2]) / (vj[2] - vi[2]) + vi[0]))
        c = !c;
      ed[j] = distPtSegSqr2D(pos, vj, vi, et[j]);
    }
  }
  if (c) {
    vcopy(closest, pos);
    return;
  }
  float dmin = ed[0], tmin = et[0];
  int imin = 0;
#pragma unroll
  for (int i = 1; i < kVertsPerPoly; ++i) {
    if (i < nv && ed[i] < dmin) {
      dmin = ed[i];
      tmin = et[i];
      imin = i;
    }
  }
  const float* va = &p->v[imin * 3];
  const float* vb = &p->v[((imin + 1) % nv) * 3];
  vlerp(closest, va, vb, tmin);
}

// ---------------------------------------------------------------------------------------
// findNearestPoly: PF.cpp:126-147 projectToPoly -> DQ.cpp:702-730 -> queryPolygons
// DQ.cpp:923-960 -> queryPolygonsInTile DQ.cpp:732-847 -> dtFindNearestPolyQuery::process
// DQ.cpp:644-679.  Group-cooperative: lanes scan W consecutive BV nodes per step (the
// pre-order array with escape offsets makes the visited set a sequence of runs), queue the
// overlapping leaves in visit order, and evaluate W candidates at a time.  "First strictly
// smaller distance wins" (DQ.cpp:670) becomes a lexicographic min over (d, visit order).
// ---------------------------------------------------------------------------------------
struct Nearest {
  uint32_t g;   // kNoPoly if nothing found
  float pt[3];
  bool over;
};

template <class G>
struct NearestAcc {
  float d;
  uint32_t ord;
  uint32_t g;
  float pt[3];
  bool over;
};

template <class G>
HBN_HD void nearestConsider(const NavView& nav, const float* center, int islandFilter,
                            uint32_t g, uint32_t ord, NearestAcc<G>& acc) {
  const PolyRec* p = &nav.polys[g];
  if (islandFilter >= 0 && p->island != islandFilter) return;
  float cp[3];
  bool over;
  closestPointOnPoly(nav, p, center, cp, &over);
  const float dx = center[0] - cp[0], dy = center[1] - cp[1], dz = center[2] - cp[2];
  float d;
  if (over) {
    d = fabsf(dy) - nav.tiles[p->tile].walkableClimb;
    d = d > 0 ? d * d : 0;
  } else {
    d = dx * dx + dy * dy + dz * dz;
  }
  if (d < acc.d) {
    acc.d = d;
    acc.ord = ord;
    acc.g = g;
    vcopy(acc.pt, cp);
    acc.over = over;
  }
}

// candQueue: 2*W uint32 of scratch owned by the group (shared memory on the device).
// rxz < 0: the reference's box.  rxz >= 0 narrows the box in x and z (hbn_snap.h snapRadius picks a
// half-extent that keeps every candidate that can win; the loose leaves behind a tile's BV tree
// are still tested against the full box).
template <class G>
HBN_HD Nearest findNearestPoly(const NavView& nav, const G& grp, const float* center,
                               const float* halfExt, int islandFilter, uint32_t* candQueue,
                               float rxz = -1.f) {
  constexpr int W = G::kWidth;
  Nearest res;
  res.g = kNoPoly;
  res.pt[0] = res.pt[1] = res.pt[2] = 0.f;
  res.over = false;
  if (!vfinite(center) || !vfinite(halfExt)) return res;  // DQ.cpp:928-933
  const int lane = grp.lane();
  NearestAcc<G> acc;
  acc.d = kFltMax;
  acc.ord = 0xffffffffu;
  acc.g = kNoPoly;
  acc.pt[0] = acc.pt[1] = acc.pt[2] = 0.f;
  acc.over = false;
  float qmin[3], qmax[3];
  for (int k = 0; k < 3; ++k) {
    const float h = (k != 1 && rxz >= 0.f) ? rxz : halfExt[k];
    qmin[k] = center[k] - h;
    qmax[k] = center[k] + h;
  }
  // calcTileLoc, DN.cpp:1191-1195
  int minx = static_cast<int>(floorf((qmin[0] - nav.orig[0]) / nav.tileWidth));
  int miny = static_cast<int>(floorf((qmin[2] - nav.orig[2]) / nav.tileHeight));
  int maxx = static_cast<int>(floorf((qmax[0] - nav.orig[0]) / nav.tileWidth));
  int maxy = static_cast<int>(floorf((qmax[2] - nav.orig[2]) / nav.tileHeight));
  // cells outside the grid hold no tiles
  if (minx < nav.gridMinX) minx = nav.gridMinX;
  if (miny < nav.gridMinY) miny = nav.gridMinY;
  if (maxx > nav.gridMinX + nav.gridW - 1) maxx = nav.gridMinX + nav.gridW - 1;
  if (maxy > nav.gridMinY + nav.gridH - 1) maxy = nav.gridMinY + nav.gridH - 1;

  int nq = 0;            // queued candidates (group-uniform)
  uint32_t ordBase = 0;  // candidates already evaluated (group-uniform)

  for (int y = miny; y <= maxy; ++y) {
    for (int x = minx; x <= maxx; ++x) {
      const int cell = (y - nav.gridMinY) * nav.gridW + (x - nav.gridMinX);
      const uint32_t c0 = nav.gridStart[cell], c1 = nav.gridStart[cell + 1];
      for (uint32_t c = c0; c < c1; ++c) {
        const TileRec& tr = nav.tiles[nav.tileOrder[c]];
        const uint32_t count = tr.bvCount ? tr.bvCount : tr.polyCount;
        uint16_t bmin[3], bmax[3];    // quantised query box of this walk
        uint16_t fbmin[3], fbmax[3];  // ... and the reference's full box (loose leaves)
        if (tr.bvCount) {
          // quantised query box, DQ.cpp:749-765
          const float qfac = tr.bvQuantFactor;
          for (int k = 0; k < 3; ++k) {
            const float mn = fclamp(qmin[k], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
            const float mx = fclamp(qmax[k], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
            bmin[k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * mn)) & 0xfffe);
            bmax[k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * mx + 1)) | 1);
            const float fmn = fclamp(center[k] - halfExt[k], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
            const float fmx = fclamp(center[k] + halfExt[k], tr.bmin[k], tr.bmax[k]) - tr.bmin[k];
            fbmin[k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * fmn)) & 0xfffe);
            fbmax[k] = static_cast<uint16_t>(static_cast<uint16_t>(static_cast<int>(qfac * fmx + 1)) | 1);
          }
        }
        uint32_t pnode = 0;
        while (pnode < count) {
          const uint32_t idx = pnode + static_cast<uint32_t>(lane);
          const bool valid = idx < count;
          bool isCand = false, isSkip = false;
          uint32_t candG = 0;
          int32_t esc = 1;
          if (valid) {
            if (tr.bvCount) {
              const BvRec n = nav.bv[tr.bvStart + idx];
              const bool leaf = n.i >= 0;
              const bool loose = leaf && (n.i & kBvLooseBit) != 0;
              const bool ov = loose ? overlapQuant(fbmin, fbmax, n.bmin, n.bmax)
                                    : overlapQuant(bmin, bmax, n.bmin, n.bmax);
              isSkip = !ov && !leaf;
              esc = -n.i;
              isCand = ov && leaf && ((n.i & kBvFailBit) == 0);
              candG = static_cast<uint32_t>(n.i & kBvIndexMask);
            } else {
              // no BV tree: linear scan with float bounds, DQ.cpp:806-842
              const uint32_t g = tr.polyStart + idx;
              const PolyRec* p = &nav.polys[g];
              if ((p->areaType >> 6) != 1 && (p->flags & kFlagWalk) != 0) {
                float pmin[3], pmax[3];
                vcopy(pmin, &p->v[0]);
                vcopy(pmax, &p->v[0]);
                for (int j = 1; j < p->nv; ++j)
                  for (int k = 0; k < 3; ++k) {
                    pmin[k] = p->v[j * 3 + k] < pmin[k] ? p->v[j * 3 + k] : pmin[k];
                    pmax[k] = p->v[j * 3 + k] > pmax[k] ? p->v[j * 3 + k] : pmax[k];
                  }
                bool ov = true;
                for (int k = 0; k < 3; ++k)
                  ov = (qmin[k] > pmax[k] || qmax[k] < pmin[k]) ? false : ov;
                isCand = ov;
                candG = g;
              }
            }
          }
          const uint32_t skipMask = grp.ballot(isSkip);
          const int firstSkip = skipMask ? ffs32(skipMask) - 1 : W;
          const uint32_t low = (firstSkip >= 32) ? 0xffffffffu : ((1u << firstSkip) - 1u);
          const uint32_t candMask = grp.ballot(isCand) & low;
          if (candMask) {
            if ((candMask >> lane) & 1u)
              candQueue[nq + popc32(candMask & ((1u << lane) - 1u))] = candG;
            nq += popc32(candMask);
            grp.sync();
            if (nq >= W) {
              nearestConsider<G>(nav, center, islandFilter, candQueue[lane], ordBase + lane, acc);
              ordBase += W;
              grp.sync();
              const uint32_t rest = (lane + W < nq) ? candQueue[lane + W] : 0u;
              grp.sync();
              candQueue[lane] = rest;
              nq -= W;
              grp.sync();
            }
          }
          if (firstSkip < W) {
            const int32_t e = grp.shfl(esc, firstSkip);
            pnode += static_cast<uint32_t>(firstSkip) + static_cast<uint32_t>(e);
          } else {
            pnode += W;
          }
        }
      }
    }
  }
  if (nq > 0) {
    if (lane < nq)
      nearestConsider<G>(nav, center, islandFilter, candQueue[lane], ordBase + lane, acc);
  }
  grp.sync();
  // lexicographic (d, ord) minimum across lanes
  for (int off = W / 2; off > 0; off >>= 1) {
    const float od = grp.shfl(acc.d, lane ^ off);
    const uint32_t oo = grp.shfl(acc.ord, lane ^ off);
    const uint32_t og = grp.shfl(acc.g, lane ^ off);
    const float ox = grp.shfl(acc.pt[0], lane ^ off);
    const float oy = grp.shfl(acc.pt[1], lane ^ off);
    const float oz = grp.shfl(acc.pt[2], lane ^ off);
    const int oov = grp.shfl(acc.over ? 1 : 0, lane ^ off);
    if (od < acc.d || (od == acc.d && oo < acc.ord)) {
      acc.d = od; acc.ord = oo; acc.g = og;
      acc.pt[0] = ox; acc.pt[1] = oy; acc.pt[2] = oz;
      acc.over = oov != 0;
    }
  }
  res.g = acc.g;
  vcopy(res.pt, acc.pt);
  res.over = acc.over;
  return res;
}

// ---------------------------------------------------------------------------------------
// A* node pool + open list.  dtNodePool / dtNodeQueue (DNode.cpp:51-200, DNode.h:108-165)
// re-laid as structure-of-arrays in a per-query workspace.  What is kept exactly: the
// allocation limit (kMaxNodes), the heap's sift sequences (bubbleUp DNode.cpp:156-167,
// bottom-up trickleDown :169-184, pop/push/modify DNode.h:118-142) and hence every tie
// break.  What changes freely: node lookup is an open-addressing table (fingerprint|index)
// and `modify` finds the node through a back pointer instead of a linear scan.
// ---------------------------------------------------------------------------------------
struct AStarWs {
  float* px; float* py; float* pz; float* cost; float* total;
  uint32_t* gid;   // g (24 bits) | state << 24 | flags << 26
  uint32_t* lnk;   // link window: start (27 bits) | count << 27
  uint16_t* pidx;  // parent node index + 1, 0 = none
  uint16_t* hpos;  // position in the heap (valid while OPEN)
  float* hkey;     // heap: node total
  uint16_t* hidx;  // heap: node index
  uint32_t* hash;  // 0 = empty, else fingerprint << 12 | (index + 1)
  int cap;         // node capacity of this workspace tier (<= kMaxNodes)
  int hashMask;    // table size - 1 (power of two, >= 2 * cap)
};
constexpr uint32_t kNodeOpen = 1u << 26;
constexpr uint32_t kNodeClosed = 2u << 26;
constexpr uint32_t kNodeGMask = 0x00ffffffu;
constexpr uint32_t kNodeKeyMask = 0x03ffffffu;  // g | state

HBN_HD size_t astarWsBytes(int cap) {
  return static_cast<size_t>(cap) * (5 * 4 + 4 + 4 + 2 + 2 + 4 + 2) + static_cast<size_t>(cap) * 2 * 4;
}
// carve a workspace out of a 16-byte aligned buffer
HBN_HD AStarWs astarWsCarve(void* buf, int cap) {
  AStarWs w;
  char* p = static_cast<char*>(buf);
  w.px = reinterpret_cast<float*>(p); p += sizeof(float) * cap;
  w.py = reinterpret_cast<float*>(p); p += sizeof(float) * cap;
  w.pz = reinterpret_cast<float*>(p); p += sizeof(float) * cap;
  w.cost = reinterpret_cast<float*>(p); p += sizeof(float) * cap;
  w.total = reinterpret_cast<float*>(p); p += sizeof(float) * cap;
  w.gid = reinterpret_cast<uint32_t*>(p); p += 4 * cap;
  w.lnk = reinterpret_cast<uint32_t*>(p); p += 4 * cap;
  w.hkey = reinterpret_cast<float*>(p); p += 4 * cap;
  w.hash = reinterpret_cast<uint32_t*>(p); p += 8 * cap;
  w.pidx = reinterpret_cast<uint16_t*>(p); p += 2 * cap;
  w.hpos = reinterpret_cast<uint16_t*>(p); p += 2 * cap;
  w.hidx = reinterpret_cast<uint16_t*>(p); p += 2 * cap;
  w.cap = cap;
  w.hashMask = 2 * cap - 1;
  return w;
}

struct Heap {
  int size;
};

HBN_HD void heapBubbleUp(const AStarWs& w, int i, uint16_t node, float key) {
  int parent = (i - 1) / 2;
  while ((i > 0) && (w.hkey[parent] > key)) {
    const uint16_t pn = w.hidx[parent];
    w.hkey[i] = w.hkey[parent];
    w.hidx[i] = pn;
    w.hpos[pn] = static_cast<uint16_t>(i);
    i = parent;
    parent = (i - 1) / 2;
  }
  w.hkey[i] = key;
  w.hidx[i] = node;
  w.hpos[node] = static_cast<uint16_t>(i);
}
HBN_HD void heapPush(const AStarWs& w, Heap& h, uint16_t node, float key) {
  h.size++;
  heapBubbleUp(w, h.size - 1, node, key);
}
HBN_HD uint16_t heapPop(const AStarWs& w, Heap& h) {
  const uint16_t result = w.hidx[0];
  h.size--;
  const uint16_t node = w.hidx[h.size];
  const float key = w.hkey[h.size];
  int i = 0;
  int child = 1;
  while (child < h.size) {
    if (((child + 1) < h.size) && (w.hkey[child] > w.hkey[child + 1])) child++;
    const uint16_t cn = w.hidx[child];
    w.hkey[i] = w.hkey[child];
    w.hidx[i] = cn;
    w.hpos[cn] = static_cast<uint16_t>(i);
    i = child;
    child = (i * 2) + 1;
  }
  heapBubbleUp(w, i, node, key);
  return result;
}

HBN_HD uint32_t nodeHash(uint32_t key) { return mix32(key * 2654435761u + 0x9e3779b9u); }

// dtNodePool::getNode (DNode.cpp:121-152).  Returns node index, or -1 if the reference
// pool (kMaxNodes) is exhausted, or -2 if only this workspace tier is.
HBN_HD int nodeGet(const AStarWs& w, int& nodeCount, uint32_t key, bool& isNew) {
  const uint32_t h = nodeHash(key);
  const uint32_t fp = (h >> 12) << 12;
  uint32_t slot = h & static_cast<uint32_t>(w.hashMask);
  for (;;) {
    const uint32_t e = w.hash[slot];
    if (e == 0) break;
    if ((e & 0xfffff000u) == fp) {
      const int idx = static_cast<int>(e & 0xfffu) - 1;
      if ((w.gid[idx] & kNodeKeyMask) == key) {
        isNew = false;
        return idx;
      }
    }
    slot = (slot + 1) & static_cast<uint32_t>(w.hashMask);
  }
  if (nodeCount >= kMaxNodes) return -1;
  if (nodeCount >= w.cap) return -2;
  const int idx = nodeCount++;
  w.hash[slot] = fp | static_cast<uint32_t>(idx + 1);
  w.gid[idx] = key;
  w.pidx[idx] = 0;
  w.cost[idx] = 0;
  w.total[idx] = 0;
  isNew = true;
  return idx;
}

struct AStarResult {
  uint32_t status;  // Detour status word of findPath, or 0xffffffff = tier overflow (retry)
  int lastBest;     // node index the corridor ends at
  int nodeCount;
  // work counters for the roofline's algorithmic bytes
  uint32_t expanded, links, neighbours;
};

// dtNavMeshQuery::findPath, DQ.cpp:973-1165.  The hash table must be zeroed by the caller.
// fastFail: stop at the first failed allocation; the status then carries DT_OUT_OF_NODES,
// which PathFinder treats as "no path" (PF.cpp:1450) whatever the search finds afterwards.
HBN_HD AStarResult astarSearch(const NavView& nav, const AStarWs& w, uint32_t startG,
                               uint32_t endG, const float* startPos, const float* endPos,
                               bool fastFail) {
  AStarResult r;
  r.status = kDtSuccess;
  r.nodeCount = 0;
  r.lastBest = 0;
  r.expanded = r.links = r.neighbours = 0;
  Heap heap;
  heap.size = 0;
  bool isNew;
  const PolyRec* sp = &nav.polys[startG];
  const int s = nodeGet(w, r.nodeCount, startG, isNew);
  w.px[s] = startPos[0]; w.py[s] = startPos[1]; w.pz[s] = startPos[2];
  w.pidx[s] = 0;
  w.cost[s] = 0;
  const float stotal = vdist(startPos, endPos) * kHScale;
  w.total[s] = stotal;
  w.gid[s] = startG | kNodeOpen;
  w.lnk[s] = sp->linkStart | (static_cast<uint32_t>(sp->linkCount) << 27);
  heapPush(w, heap, static_cast<uint16_t>(s), stotal);
  int lastBest = s;
  float lastBestCost = stotal;
  bool outOfNodes = false;

  while (heap.size > 0) {
    const int best = heapPop(w, heap);
    uint32_t bg = w.gid[best];
    bg = (bg & ~kNodeOpen) | kNodeClosed;
    w.gid[best] = bg;
    const uint32_t bestG = bg & kNodeGMask;
    if (bestG == endG) {
      lastBest = best;
      break;
    }
    const uint32_t bp = w.pidx[best];
    const uint32_t parentG = bp ? (w.gid[bp - 1] & kNodeGMask) : kNoPoly;
    const float bpos[3] = {w.px[best], w.py[best], w.pz[best]};
    const float bcost = w.cost[best];
    const uint32_t lw = w.lnk[best];
    const uint32_t l0 = lw & 0x07ffffffu, ln = lw >> 27;
    r.expanded++;
    r.links += ln;
    for (uint32_t j = 0; j < ln; ++j) {
      const LinkRec L = nav.links[l0 + j];
      const uint32_t nei = L.nei;
      r.neighbours += (nei != kNoPoly) ? 1u : 0u;
      if (nei == kNoPoly || nei == parentG) continue;
      if ((L.meta & kLinkPassBit) == 0) continue;
      const uint32_t state = (L.meta >> kLinkStateShift) & 3u;
      const int n = nodeGet(w, r.nodeCount, nei | (state << 24), isNew);
      if (n < 0) {
        if (n == -2) {
          r.status = 0xffffffffu;
          return r;
        }
        outOfNodes = true;
        if (fastFail) goto done;
        continue;
      }
      float npos[3];
      if (isNew) {
        npos[0] = L.mid[0]; npos[1] = L.mid[1]; npos[2] = L.mid[2];
        w.px[n] = npos[0]; w.py[n] = npos[1]; w.pz[n] = npos[2];
        w.lnk[n] = L.neiLinkStart | ((L.meta >> kLinkNeiCountShift) << 27);
      } else {
        npos[0] = w.px[n]; npos[1] = w.py[n]; npos[2] = w.pz[n];
      }
      float cost, heuristic;
      if (nei == endG) {
        const float curCost = vdist(bpos, npos);
        const float endCost = vdist(npos, endPos);
        cost = bcost + curCost + endCost;
        heuristic = 0;
      } else {
        const float curCost = vdist(bpos, npos);
        cost = bcost + curCost;
        heuristic = vdist(npos, endPos) * kHScale;
      }
      const float total = cost + heuristic;
      const uint32_t ng = w.gid[n];
      if ((ng & kNodeOpen) && total >= w.total[n]) continue;
      if ((ng & kNodeClosed) && total >= w.total[n]) continue;
      w.pidx[n] = static_cast<uint16_t>(best + 1);
      w.cost[n] = cost;
      w.total[n] = total;
      if (ng & kNodeOpen) {
        w.gid[n] = ng & ~kNodeClosed;
        heapBubbleUp(w, w.hpos[n], static_cast<uint16_t>(n), total);  // modify
      } else {
        w.gid[n] = (ng & ~kNodeClosed) | kNodeOpen;
        heapPush(w, heap, static_cast<uint16_t>(n), total);
      }
      if (heuristic < lastBestCost) {
        lastBestCost = heuristic;
        lastBest = n;
      }
    }
  }
done:
  r.lastBest = lastBest;
  if ((w.gid[lastBest] & kNodeGMask) != endG) r.status |= kDtPartialResult;
  if (outOfNodes) r.status |= kDtOutOfNodes;
  return r;
}

// getPathToNode, DQ.cpp:1167-1205: corridor of global poly indices into path[0..maxPath).
// Returns the stored count; *fullLen gets the untruncated length.
HBN_HD int astarExtractPath(const AStarWs& w, int endNode, uint32_t* path, int maxPath,
                            int* fullLen) {
  int length = 0;
  int cur = endNode;
  for (;;) {
    length++;
    const uint32_t p = w.pidx[cur];
    if (!p) break;
    cur = static_cast<int>(p) - 1;
  }
  *fullLen = length;
  cur = endNode;
  int writeCount = length;
  for (; writeCount > maxPath; writeCount--) cur = static_cast<int>(w.pidx[cur]) - 1;
  for (int i = writeCount - 1; i >= 0; i--) {
    path[i] = w.gid[cur] & kNodeGMask;
    cur = static_cast<int>(w.pidx[cur]) - 1;
  }
  return writeCount;
}

// First link of poly `from` that leads to `to` (getPortalPoints' search, DQ.cpp:2276-2285)
HBN_HD uint32_t findLinkTo(const NavView& nav, uint32_t from, uint32_t to) {
  const PolyRec* p = &nav.polys[from];
  const uint32_t l0 = p->linkStart, ln = p->linkCount;
  for (uint32_t j = 0; j < ln; ++j)
    if (nav.links[l0 + j].nei == to) return l0 + j;
  return kNoPoly;
}

// ---------------------------------------------------------------------------------------
// findStraightPath (funnel), DQ.cpp:1793-2022 with options == 0 and no flag/ref outputs,
// as PathFinder calls it (PF.cpp:1456-1458), fused with pathLength (PF.cpp:1400-1411).
// path: corridor (global indices); pathLink[i]: LinkRec index of the portal path[i] ->
// path[i+1].  Points go to outPts (stride 3, may be null) up to maxOut; the running
// length accumulates exactly like pathLength (first term |p0 - p0| = 0).
// Returns the Detour status word; *npts the number of points.
// ---------------------------------------------------------------------------------------
struct Funnel {
  float last[3];
  int count;
  float length;
  float* out;
  int maxOut;
};
// appendVertex, DQ.cpp:1689-1724.  Returns 0 = in progress, else a final status word.
HBN_HD uint32_t funnelAppend(Funnel& f, const float* pos, bool isEnd) {
  if (f.count > 0 && vequal(f.last, pos)) {
    // equal to the last vertex: only flags/refs would be updated
  } else {
    if (f.out && f.count < f.maxOut) {
      f.out[f.count * 3 + 0] = pos[0];
      f.out[f.count * 3 + 1] = pos[1];
      f.out[f.count * 3 + 2] = pos[2];
    }
    if (f.count > 0) f.length += mnDist(f.last, pos);
    else f.length += 0.0f;
    vcopy(f.last, pos);
    f.count++;
    if (f.count >= kMaxPathPolys) return kDtSuccess | kDtBufferTooSmall;
    if (isEnd) return kDtSuccess;
  }
  return 0;
}

// Corridor accessors of the funnel: poly(i) = global poly index, link(i) = LinkRec index of the
// portal poly(i) -> poly(i+1) (kNoPoly: none), portal(i, li, l, r) = its left / right points.
// staged (nullable): the corridor's portals already gathered, staged[i] == nav.portals[pathLink[i]].
struct ArrayCorridor {
  const NavView& nav;
  const uint32_t* path;
  const uint32_t* pathLink;
  const PortalRec* staged;
  HBN_HD uint32_t poly(int i) const { return path[i]; }
  HBN_HD uint32_t link(int i) const { return pathLink[i]; }
  HBN_HD void portal(int i, uint32_t li, float* l, float* r) const {
    const PortalRec& po = staged ? staged[i] : nav.portals[li];
    vcopy(l, po.l);
    vcopy(r, po.r);
  }
};

template <class C>
HBN_HD uint32_t funnelStraightPathT(const NavView& nav, const float* startPos, const float* endPos,
                                    const C& cor, int pathSize, Funnel& f) {
  f.count = 0;
  f.length = 0.0f;
  if (!vfinite(startPos) || !vfinite(endPos) || pathSize <= 0) return kDtFailure | kDtInvalidParam;
  float closestStartPos[3], closestEndPos[3];
  closestPointOnPolyBoundary(&nav.polys[cor.poly(0)], startPos, closestStartPos);
  closestPointOnPolyBoundary(&nav.polys[cor.poly(pathSize - 1)], endPos, closestEndPos);
  uint32_t stat = funnelAppend(f, closestStartPos, false);
  if (stat) return stat;
  if (pathSize > 1) {
    float portalApex[3], portalLeft[3], portalRight[3];
    vcopy(portalApex, closestStartPos);
    vcopy(portalLeft, portalApex);
    vcopy(portalRight, portalApex);
    int apexIndex = 0, leftIndex = 0, rightIndex = 0;
    bool leftIsEnd = false, rightIsEnd = false;  // leftPolyRef/rightPolyRef == 0
    for (int i = 0; i < pathSize; ++i) {
      float left[3], right[3];
      if (i + 1 < pathSize) {
        const uint32_t li = cor.link(i);
        if (li == kNoPoly) {
          // getPortalPoints failed (DQ.cpp:1853-1878): clamp end to path[i], partial result
          closestPointOnPolyBoundary(&nav.polys[cor.poly(i)], endPos, closestEndPos);
          funnelAppend(f, closestEndPos, false);
          return kDtSuccess | kDtPartialResult |
                 ((f.count >= kMaxPathPolys) ? kDtBufferTooSmall : 0u);
        }
        cor.portal(i, li, left, right);
        if (i == 0) {
          float t;
          if (distPtSegSqr2D(portalApex, left, right, t) < sqr(0.001f)) continue;
        }
      } else {
        vcopy(left, closestEndPos);
        vcopy(right, closestEndPos);
      }
      // right vertex
      if (triArea2D(portalApex, portalRight, right) <= 0.0f) {
        if (vequal(portalApex, portalRight) || triArea2D(portalApex, portalLeft, right) > 0.0f) {
          vcopy(portalRight, right);
          rightIsEnd = !(i + 1 < pathSize);
          rightIndex = i;
        } else {
          vcopy(portalApex, portalLeft);
          apexIndex = leftIndex;
          stat = funnelAppend(f, portalApex, leftIsEnd);
          if (stat) return stat;
          vcopy(portalLeft, portalApex);
          vcopy(portalRight, portalApex);
          leftIndex = apexIndex;
          rightIndex = apexIndex;
          i = apexIndex;
          continue;
        }
      }
      // left vertex
      if (triArea2D(portalApex, portalLeft, left) >= 0.0f) {
        if (vequal(portalApex, portalLeft) || triArea2D(portalApex, portalRight, left) < 0.0f) {
          vcopy(portalLeft, left);
          leftIsEnd = !(i + 1 < pathSize);
          leftIndex = i;
        } else {
          vcopy(portalApex, portalRight);
          apexIndex = rightIndex;
          stat = funnelAppend(f, portalApex, rightIsEnd);
          if (stat) return stat;
          vcopy(portalLeft, portalApex);
          vcopy(portalRight, portalApex);
          leftIndex = apexIndex;
          rightIndex = apexIndex;
          i = apexIndex;
          continue;
        }
      }
    }
  }
  funnelAppend(f, closestEndPos, true);
  return kDtSuccess | ((f.count >= kMaxPathPolys) ? kDtBufferTooSmall : 0u);
}

HBN_HD uint32_t funnelStraightPath(const NavView& nav, const float* startPos, const float* endPos,
                                   const uint32_t* path, const uint32_t* pathLink, int pathSize,
                                   Funnel& f, const PortalRec* staged = nullptr) {
  const ArrayCorridor cor{nav, path, pathLink, staged};
  return funnelStraightPathT(nav, startPos, endPos, cor, pathSize, f);
}

// ---------------------------------------------------------------------------------------
// moveAlongSurface, DQ.cpp:2044-2245 (tiny pool of 64 nodes, FIFO of 48, trap T9)
// ---------------------------------------------------------------------------------------
constexpr int kTinyNodes = 64;
constexpr int kMasStack = 48;
struct TinyPool {
  uint32_t g[kTinyNodes];
  uint8_t pidx[kTinyNodes];   // parent index + 1
  uint8_t closed[kTinyNodes];
  int count;
};
HBN_HD int tinyGet(TinyPool& tp, uint32_t g) {
  for (int i = 0; i < tp.count; ++i)
    if (tp.g[i] == g) return i;
  if (tp.count >= kTinyNodes) return -1;
  const int i = tp.count++;
  tp.g[i] = g;
  tp.pidx[i] = 0;
  tp.closed[i] = 0;
  return i;
}

// visited: up to maxVisited global poly indices; returns count (0 = nothing reachable)
HBN_HD int moveAlongSurface(const NavView& nav, uint32_t startG, const float* startPos,
                            const float* endPos, float* resultPos, uint32_t* visited,
                            int maxVisited, TinyPool& tp) {
  if (startG == kNoPoly || !vfinite(startPos) || !vfinite(endPos)) return 0;
  uint8_t stack[kMasStack];
  int nstack = 0;
  tp.count = 0;
  const int startNode = tinyGet(tp, startG);
  tp.closed[startNode] = 1;
  stack[nstack++] = static_cast<uint8_t>(startNode);
  float bestPos[3];
  float bestDist = kFltMax;
  int bestNode = -1;
  vcopy(bestPos, startPos);
  float searchPos[3];
  vlerp(searchPos, startPos, endPos, 0.5f);
  const float searchRadSqr = sqr(vdist(startPos, endPos) / 2.0f + 0.001f);

  while (nstack) {
    const int curNode = stack[0];
    for (int i = 0; i < nstack - 1; ++i) stack[i] = stack[i + 1];
    nstack--;
    const uint32_t curG = tp.g[curNode];
    const PolyRec* cp = &nav.polys[curG];
    const int nverts = cp->nv;
    if (pointInPolygon(endPos, cp->v, nverts)) {
      bestNode = curNode;
      vcopy(bestPos, endPos);
      break;
    }
    for (int i = 0, j = nverts - 1; i < nverts; j = i++) {
      constexpr int MAX_NEIS = 8;
      int nneis = 0;
      uint32_t neis[MAX_NEIS];
      const uint16_t nj = cp->neis[j];
      if (nj & kExtLink) {
        const uint32_t l0 = cp->linkStart, ln = cp->linkCount;
        for (uint32_t k = 0; k < ln; ++k) {
          const LinkRec L = nav.links[l0 + k];
          if ((L.meta & kLinkEdgeMask) == static_cast<uint32_t>(j)) {
            if (L.nei != kNoPoly && (L.meta & kLinkPassBit)) {
              if (nneis < MAX_NEIS) neis[nneis++] = L.nei;
            }
          }
        }
      } else if (nj) {
        const uint32_t g = nav.tiles[cp->tile].polyStart + static_cast<uint32_t>(nj - 1);
        if ((nav.polys[g].flags & kFlagWalk) != 0) neis[nneis++] = g;
      }
      const float* vj = &cp->v[j * 3];
      const float* vi = &cp->v[i * 3];
      if (!nneis) {
        float tseg;
        const float distSqr = distPtSegSqr2D(endPos, vj, vi, tseg);
        if (distSqr < bestDist) {
          vlerp(bestPos, vj, vi, tseg);
          bestDist = distSqr;
          bestNode = curNode;
        }
      } else {
        for (int k = 0; k < nneis; ++k) {
          const int nn = tinyGet(tp, neis[k]);
          if (nn < 0) continue;
          if (tp.closed[nn]) continue;
          float tseg;
          const float distSqr = distPtSegSqr2D(searchPos, vj, vi, tseg);
          if (distSqr > searchRadSqr) continue;
          if (nstack < kMasStack) {
            tp.pidx[nn] = static_cast<uint8_t>(curNode + 1);
            tp.closed[nn] = 1;
            stack[nstack++] = static_cast<uint8_t>(nn);
          }
        }
      }
    }
  }
  int n = 0;
  if (bestNode >= 0) {
    // reverse the parent chain, then walk it forward (DQ.cpp:2209-2236)
    int prev = -1;
    int node = bestNode;
    do {
      const int next = static_cast<int>(tp.pidx[node]) - 1;
      tp.pidx[node] = static_cast<uint8_t>(prev + 1);
      prev = node;
      node = next;
    } while (node >= 0);
    node = prev;
    do {
      visited[n++] = tp.g[node];
      if (n >= maxVisited) break;
      node = static_cast<int>(tp.pidx[node]) - 1;
    } while (node >= 0);
  }
  vcopy(resultPos, bestPos);
  return n;
}

// The no-sliding clamp of PathFinder::Impl::tryStep, PF.cpp:1610-1677
HBN_HD void noSlidingClamp(const NavView& nav, const uint32_t* polys, int numPolys,
                           const float* pathStart, const float* end, float* endPoint) {
  float bestDist = kFltMax;
  bool hitWall = false;
  float bestPos[3] = {0, 0, 0};
  for (int ip = 0; ip < numPolys; ++ip) {
    const PolyRec* p = &nav.polys[polys[ip]];
    const int nv = p->nv;
    for (int j = 0; j < nv; ++j) {
      bool isWall = false;
      const uint16_t nj = p->neis[j];
      if (nj == 0) {
        isWall = true;
      } else if (nj & kExtLink) {
        bool hasPassable = false;
        const uint32_t l0 = p->linkStart, ln = p->linkCount;
        for (uint32_t k = 0; k < ln; ++k) {
          const LinkRec L = nav.links[l0 + k];
          if ((L.meta & kLinkEdgeMask) == static_cast<uint32_t>(j) && L.nei != kNoPoly &&
              (L.meta & kLinkPassBit)) {
            hasPassable = true;
            break;
          }
        }
        isWall = !hasPassable;
      }
      if (!isWall) continue;
      const float* vj = &p->v[j * 3];
      const int nextIdx = (j + 1 < nv) ? (j + 1) : 0;
      const float* vi = &p->v[nextIdx * 3];
      float s, t;
      if (intersectSegSeg2D(vj, vi, pathStart, end, s, t) && t >= 0.0f && t <= 1.0f &&
          s >= 0.0f && s <= 1.0f) {
        float newPos[3];
        vlerp(newPos, vj, vi, s);
        const float distSqr = vdist2DSqr(newPos, end);
        if (distSqr < bestDist) {
          vcopy(bestPos, newPos);
          bestDist = distSqr;
          hitWall = true;
        }
      }
    }
  }
  if (hitWall) vcopy(endPoint, bestPos);
}

// ---------------------------------------------------------------------------------------
// findDistanceToWall, DQ.cpp:3470-3655 (Dijkstra with a shrinking radius).  Uses the same
// node workspace as A*; returns 0xffffffff on tier overflow (retry with a larger tier).
// hitPos must be pre-set by the caller (the reference leaves it untouched when nothing is
// hit, trap T8; PathFinder passes a zero-initialised vector, PF.cpp:1806).
// ---------------------------------------------------------------------------------------
HBN_HD uint32_t distanceToWall(const NavView& nav, const AStarWs& w, uint32_t startG,
                               const float* centerPos, float maxRadius, float* hitDist,
                               float* hitPos, float* hitNormal) {
  if (startG == kNoPoly || !vfinite(centerPos) || maxRadius < 0 || !finitef(maxRadius))
    return kDtFailure | kDtInvalidParam;
  int nodeCount = 0;
  Heap heap;
  heap.size = 0;
  bool isNew;
  const PolyRec* sp = &nav.polys[startG];
  const int s = nodeGet(w, nodeCount, startG, isNew);
  w.px[s] = centerPos[0]; w.py[s] = centerPos[1]; w.pz[s] = centerPos[2];
  w.pidx[s] = 0;
  w.cost[s] = 0;
  w.total[s] = 0;
  w.gid[s] = startG | kNodeOpen;
  w.lnk[s] = sp->linkStart | (static_cast<uint32_t>(sp->linkCount) << 27);
  heapPush(w, heap, static_cast<uint16_t>(s), 0.0f);
  float radiusSqr = sqr(maxRadius);
  uint32_t status = kDtSuccess;

  while (heap.size > 0) {
    const int best = heapPop(w, heap);
    uint32_t bg = w.gid[best];
    bg = (bg & ~kNodeOpen) | kNodeClosed;
    w.gid[best] = bg;
    const uint32_t bestG = bg & kNodeGMask;
    const PolyRec* bp = &nav.polys[bestG];
    const uint32_t ppi = w.pidx[best];
    const uint32_t parentG = ppi ? (w.gid[ppi - 1] & kNodeGMask) : kNoPoly;
    const int nv = bp->nv;
    const uint32_t l0 = bp->linkStart, ln = bp->linkCount;
    // hit test walls
    for (int i = 0, j = nv - 1; i < nv; j = i++) {
      const uint16_t nj = bp->neis[j];
      if (nj & kExtLink) {
        bool solid = true;
        for (uint32_t k = 0; k < ln; ++k) {
          const LinkRec L = nav.links[l0 + k];
          if ((L.meta & kLinkEdgeMask) == static_cast<uint32_t>(j)) {
            if (L.nei != kNoPoly && (L.meta & kLinkPassBit)) solid = false;
            break;
          }
        }
        if (!solid) continue;
      } else if (nj) {
        const uint32_t g = nav.tiles[bp->tile].polyStart + static_cast<uint32_t>(nj - 1);
        if ((nav.polys[g].flags & kFlagWalk) != 0) continue;
      }
      const float* vj = &bp->v[j * 3];
      const float* vi = &bp->v[i * 3];
      float tseg;
      const float distSqr = distPtSegSqr2D(centerPos, vj, vi, tseg);
      if (distSqr > radiusSqr) continue;
      radiusSqr = distSqr;
      hitPos[0] = vj[0] + (vi[0] - vj[0]) * tseg;
      hitPos[1] = vj[1] + (vi[1] - vj[1]) * tseg;
      hitPos[2] = vj[2] + (vi[2] - vj[2]) * tseg;
    }
    const float bpos[3] = {w.px[best], w.py[best], w.pz[best]};
    const float btotal = w.total[best];
    for (uint32_t k = 0; k < ln; ++k) {
      const LinkRec L = nav.links[l0 + k];
      const uint32_t nei = L.nei;
      if (nei == kNoPoly || nei == parentG) continue;
      if (L.meta & kLinkOffmeshBit) continue;
      const int edge = static_cast<int>(L.meta & kLinkEdgeMask);
      const float* va = &bp->v[edge * 3];
      const float* vb = &bp->v[((edge + 1) % nv) * 3];
      float tseg;
      const float distSqr = distPtSegSqr2D(centerPos, va, vb, tseg);
      if (distSqr > radiusSqr) continue;
      if ((L.meta & kLinkPassBit) == 0) continue;
      const int n = nodeGet(w, nodeCount, nei, isNew);  // state 0 (DQ.cpp:3598)
      if (n < 0) {
        if (n == -2) return 0xffffffffu;
        status |= kDtOutOfNodes;
        continue;
      }
      const uint32_t ng = w.gid[n];
      if (ng & kNodeClosed) continue;
      float npos[3];
      if (isNew) {
        npos[0] = L.mid[0]; npos[1] = L.mid[1]; npos[2] = L.mid[2];
        w.px[n] = npos[0]; w.py[n] = npos[1]; w.pz[n] = npos[2];
      } else {
        npos[0] = w.px[n]; npos[1] = w.py[n]; npos[2] = w.pz[n];
      }
      const float total = btotal + vdist(bpos, npos);
      if ((ng & kNodeOpen) && total >= w.total[n]) continue;
      w.pidx[n] = static_cast<uint16_t>(best + 1);
      w.total[n] = total;
      if (ng & kNodeOpen) {
        heapBubbleUp(w, w.hpos[n], static_cast<uint16_t>(n), total);
      } else {
        w.gid[n] = ng | kNodeOpen;
        heapPush(w, heap, static_cast<uint16_t>(n), total);
      }
    }
  }
  // hit normal (dtVsub + dtVnormalize, DetourCommon.h:263-269)
  hitNormal[0] = centerPos[0] - hitPos[0];
  hitNormal[1] = centerPos[1] - hitPos[1];
  hitNormal[2] = centerPos[2] - hitPos[2];
  const float d = 1.0f / fsqrt(sqr(hitNormal[0]) + sqr(hitNormal[1]) + sqr(hitNormal[2]));
  hitNormal[0] *= d;
  hitNormal[1] *= d;
  hitNormal[2] *= d;
  *hitDist = fsqrt(radiusSqr);
  return status;
}

// ---------------------------------------------------------------------------------------
// The same search for ONE LANE (k_wall_lane: a query per thread).  The radius shrinks to the nearest wall
// found so far, so the search stays tiny (at most 16 nodes on every query of the C2-C5 workloads, more than 8 on under 1 % of them, at the
// default 2 m radius, 5 m too): the whole node pool is C nodes of 6 words in shared memory, word j of a
// thread at ws[j * S] (S = threads per block: conflict-free whatever the lanes index).  With so few
// nodes the pool's hash and the heap's position array are linear scans, and a heap entry is just the
// node index (its key is the node's total, which only changes together with the heap -- DQ.cpp:3625-3639).
// Same operations in the same order as distanceToWall above; 0xffffffff = more than C nodes (the query
// goes to the warp-per-query tiers).
//   words [0,C) px  [C,2C) py  [2C,3C) pz  [3C,4C) total  [4C,5C) poly | open << 24 | closed << 25 |
//   (parent node + 1) << 26   [5C,6C) heap
// ---------------------------------------------------------------------------------------
constexpr uint32_t kWsOpen = 1u << 24, kWsClosed = 1u << 25;
constexpr int kWallLaneCap = 8;
template <int C, int S>
HBN_HD uint32_t distanceToWallSmall(const NavView& nav, uint32_t* ws, uint32_t startG, const float* centerPos,
                                    float maxRadius, float* hitDist, float* hitPos, float* hitNormal) {
  static_assert(C <= 32, "parent index + 1 must fit 6 bits");
  if (startG == kNoPoly || !vfinite(centerPos) || maxRadius < 0 || !finitef(maxRadius))
    return kDtFailure | kDtInvalidParam;
  float* fw = reinterpret_cast<float*>(ws);
#define HBN_WS(kind, i) ((kind) * C + (i)) * S
  int nodeCount = 1, hsize = 1;
  fw[HBN_WS(0, 0)] = centerPos[0]; fw[HBN_WS(1, 0)] = centerPos[1]; fw[HBN_WS(2, 0)] = centerPos[2];
  fw[HBN_WS(3, 0)] = 0.f;
  ws[HBN_WS(4, 0)] = startG | kWsOpen;
  ws[HBN_WS(5, 0)] = 0u;
  float radiusSqr = sqr(maxRadius);
  // dtNodeQueue::bubbleUp (DNode.cpp:156-167) of `node` (total `key`) from heap position i
  const auto bubbleUp = [&](int i, uint32_t node, float key) {
    while (i > 0) {
      const int parent = (i - 1) / 2;
      const uint32_t pn = ws[HBN_WS(5, parent)];
      if (!(fw[HBN_WS(3, pn)] > key)) break;
      ws[HBN_WS(5, i)] = pn;
      i = parent;
    }
    ws[HBN_WS(5, i)] = node;
  };
  while (hsize > 0) {
    // pop (DNode.cpp:169-184 via DNode.h:124-130)
    const uint32_t best = ws[HBN_WS(5, 0)];
    hsize--;
    {
      const uint32_t last = ws[HBN_WS(5, hsize)];
      const float lkey = fw[HBN_WS(3, last)];
      int i = 0, child = 1;
      while (child < hsize) {
        uint32_t cn = ws[HBN_WS(5, child)];
        if ((child + 1) < hsize) {
          const uint32_t c1 = ws[HBN_WS(5, child + 1)];
          if (fw[HBN_WS(3, cn)] > fw[HBN_WS(3, c1)]) {
            cn = c1;
            child++;
          }
        }
        ws[HBN_WS(5, i)] = cn;
        i = child;
        child = (i * 2) + 1;
      }
      bubbleUp(i, last, lkey);
    }
    uint32_t bg = ws[HBN_WS(4, best)];
    bg = (bg & ~kWsOpen) | kWsClosed;
    ws[HBN_WS(4, best)] = bg;
    const uint32_t bestG = bg & kNodeGMask;
    const PolyRec* bp = &nav.polys[bestG];
    const uint32_t ppi = bg >> 26;
    const uint32_t parentG = ppi ? (ws[HBN_WS(4, ppi - 1)] & kNodeGMask) : kNoPoly;
    const int nv = bp->nv;
    const uint32_t l0 = bp->linkStart, ln = bp->linkCount;
    // hit test walls
    for (int i = 0, j = nv - 1; i < nv; j = i++) {
      const uint16_t nj = bp->neis[j];
      if (nj & kExtLink) {
        bool solid = true;
        for (uint32_t k = 0; k < ln; ++k) {
          const LinkRec L = nav.links[l0 + k];
          if ((L.meta & kLinkEdgeMask) == static_cast<uint32_t>(j)) {
            if (L.nei != kNoPoly && (L.meta & kLinkPassBit)) solid = false;
            break;
          }
        }
        if (!solid) continue;
      } else if (nj) {
        const uint32_t g = nav.tiles[bp->tile].polyStart + static_cast<uint32_t>(nj - 1);
        if ((nav.polys[g].flags & kFlagWalk) != 0) continue;
      }
      const float* vj = &bp->v[j * 3];
      const float* vi = &bp->v[i * 3];
      float tseg;
      const float distSqr = distPtSegSqr2D(centerPos, vj, vi, tseg);
      if (distSqr > radiusSqr) continue;
      radiusSqr = distSqr;
      hitPos[0] = vj[0] + (vi[0] - vj[0]) * tseg;
      hitPos[1] = vj[1] + (vi[1] - vj[1]) * tseg;
      hitPos[2] = vj[2] + (vi[2] - vj[2]) * tseg;
    }
    const float bpos[3] = {fw[HBN_WS(0, best)], fw[HBN_WS(1, best)], fw[HBN_WS(2, best)]};
    const float btotal = fw[HBN_WS(3, best)];
    for (uint32_t k = 0; k < ln; ++k) {
      const LinkRec L = nav.links[l0 + k];
      const uint32_t nei = L.nei;
      if (nei == kNoPoly || nei == parentG) continue;
      if (L.meta & kLinkOffmeshBit) continue;
      const int edge = static_cast<int>(L.meta & kLinkEdgeMask);
      const float* va = &bp->v[edge * 3];
      const float* vb = &bp->v[((edge + 1) % nv) * 3];
      float tseg;
      const float distSqr = distPtSegSqr2D(centerPos, va, vb, tseg);
      if (distSqr > radiusSqr) continue;
      if ((L.meta & kLinkPassBit) == 0) continue;
      // dtNodePool::getNode(nei, 0)
      int n = -1;
      for (int t = 0; t < nodeCount; ++t)
        if ((ws[HBN_WS(4, t)] & kNodeGMask) == nei) {
          n = t;
          break;
        }
      const bool isNew = n < 0;
      if (isNew) {
        if (nodeCount >= C) return 0xffffffffu;
        n = nodeCount++;
        ws[HBN_WS(4, n)] = nei;
        fw[HBN_WS(3, n)] = 0.f;
      }
      const uint32_t ng = ws[HBN_WS(4, n)];
      if (ng & kWsClosed) continue;
      float npos[3];
      if (isNew) {
        npos[0] = L.mid[0]; npos[1] = L.mid[1]; npos[2] = L.mid[2];
        fw[HBN_WS(0, n)] = npos[0]; fw[HBN_WS(1, n)] = npos[1]; fw[HBN_WS(2, n)] = npos[2];
      } else {
        npos[0] = fw[HBN_WS(0, n)]; npos[1] = fw[HBN_WS(1, n)]; npos[2] = fw[HBN_WS(2, n)];
      }
      const float total = btotal + vdist(bpos, npos);
      if ((ng & kWsOpen) && total >= fw[HBN_WS(3, n)]) continue;
      fw[HBN_WS(3, n)] = total;
      if (ng & kWsOpen) {
        ws[HBN_WS(4, n)] = (ng & 0x03ffffffu) | ((best + 1u) << 26);
        int pos = 0;
        while (ws[HBN_WS(5, pos)] != static_cast<uint32_t>(n)) pos++;  // dtNodeQueue::modify's scan
        bubbleUp(pos, static_cast<uint32_t>(n), total);
      } else {
        ws[HBN_WS(4, n)] = (ng & 0x03ffffffu) | kWsOpen | ((best + 1u) << 26);
        hsize++;
        bubbleUp(hsize - 1, static_cast<uint32_t>(n), total);
      }
    }
  }
#undef HBN_WS
  // hit normal (dtVsub + dtVnormalize, DetourCommon.h:263-269)
  hitNormal[0] = centerPos[0] - hitPos[0];
  hitNormal[1] = centerPos[1] - hitPos[1];
  hitNormal[2] = centerPos[2] - hitPos[2];
  const float d = 1.0f / fsqrt(sqr(hitNormal[0]) + sqr(hitNormal[1]) + sqr(hitNormal[2]));
  hitNormal[0] *= d;
  hitNormal[1] *= d;
  hitNormal[2] *= d;
  *hitDist = fsqrt(radiusSqr);
  return kDtSuccess;
}

// ---------------------------------------------------------------------------------------
// findRandomPoint, DQ.cpp:226-315 + dtRandomPointInConvexPoly DetourCommon.cpp:332-369.
// The reference consumes its uniform stream sequentially: one draw per tile with a header,
// one per ground poly passing the filter in the chosen tile, then s and t.  With a
// counter-based stream u(draw) both reservoir scans become "last index i with
// u_i * sum_i <= w_i", which lanes evaluate in parallel over the tabulated running sums.
// ---------------------------------------------------------------------------------------
HBN_HD void randomPointInConvexPoly(const float* pts, int npts, float s, float t, float* out) {
  float areas[kVertsPerPoly];
  float areasum = 0.0f;
  for (int i = 2; i < npts; i++) {
    areas[i] = triArea2D(&pts[0], &pts[(i - 1) * 3], &pts[i * 3]);
    areasum += (0.001f > areas[i]) ? 0.001f : areas[i];  // dtMax(0.001f, areas[i])
  }
  const float thr = s * areasum;
  float acc = 0.0f;
  float u = 1.0f;
  int tri = npts - 1;
  for (int i = 2; i < npts; i++) {
    const float dacc = areas[i];
    if (thr >= acc && thr < (acc + dacc)) {
      u = (thr - acc) / dacc;
      tri = i;
      break;
    }
    acc += dacc;
  }
  const float v = fsqrt(t);
  const float a = 1 - v;
  const float b = (1 - u) * v;
  const float c = u * v;
  const float* pa = &pts[0];
  const float* pb = &pts[(tri - 1) * 3];
  const float* pc = &pts[tri * 3];
  out[0] = a * pa[0] + b * pb[0] + c * pc[0];
  out[1] = a * pa[1] + b * pb[1] + c * pc[1];
  out[2] = a * pa[2] + b * pb[2] + c * pc[2];
}

// One findRandomPoint call.  drawBase: index of the first draw this call consumes;
// *drawsUsed: how many it consumed (depends on success, trap T7).  Returns the poly or kNoPoly.
template <class G>
HBN_HD uint32_t findRandomPoint(const NavView& nav, const G& grp, uint64_t seed, uint64_t query,
                                uint32_t drawBase, int island, float* outPt,
                                uint32_t* drawsUsed) {
  constexpr int W = G::kWidth;
  const int lane = grp.lane();
  // tile reservoir over tiles with a header, table order (DQ.cpp:236-251)
  int chosen = -1;
  {
    int best = -1;
    uint32_t k = 0;  // rank among present tiles
    // present tiles are dense in practice (trap T10); rank = running count
    for (uint32_t base = 0; base < nav.numTiles; base += W) {
      const uint32_t ti = base + lane;
      const bool present = ti < nav.numTiles && nav.tiles[ti].pad[0] != 0;
      const uint32_t pm = grp.ballot(present);
      if (present) {
        const uint32_t rank = k + popc32(pm & ((1u << lane) - 1u));
        const float tsum = static_cast<float>(rank + 1);  // tsum += 1.0f, exact in f32
        const float u = uniform01(seed, query, drawBase + rank);
        if (u * tsum <= 1.0f) best = static_cast<int>(ti);
      }
      k += popc32(pm);
    }
    // max over lanes
    for (int off = W / 2; off > 0; off >>= 1) {
      const int o = grp.shfl(best, lane ^ off);
      best = o > best ? o : best;
    }
    chosen = best;
    *drawsUsed = k;
  }
  if (chosen < 0) return kNoPoly;
  const TileRec& tr = nav.tiles[chosen];
  uint32_t w0 = tr.randStart, wn = tr.randCount;
  if (island >= 0) {
    w0 = 0; wn = 0;
    const uint32_t a = nav.tileIslStart[chosen], b = nav.tileIslStart[chosen + 1];
    for (uint32_t i = a; i < b; ++i)
      if (nav.tileIslId[i] == island) { w0 = nav.tileIslWin[i]; wn = nav.tileIslCnt[i]; break; }
  }
  const uint32_t polyDraw0 = drawBase + *drawsUsed;
  *drawsUsed += wn;
  // poly reservoir: last entry with u * areaSum <= area; scan backwards, stop at first hit
  int bestE = -1;
  for (int hi = static_cast<int>(wn); hi > 0 && bestE < 0; hi -= W) {
    const int e = hi - 1 - lane;
    int mine = -1;
    if (e >= 0) {
      const RandEntry re = nav.randEntries[w0 + e];
      const float u = uniform01(seed, query, polyDraw0 + static_cast<uint32_t>(e));
      if (u * re.areaSum <= re.area) mine = e;
    }
    for (int off = W / 2; off > 0; off >>= 1) {
      const int o = grp.shfl(mine, lane ^ off);
      mine = o > mine ? o : mine;
    }
    bestE = mine;
  }
  if (bestE < 0) return kNoPoly;
  const uint32_t g = nav.randEntries[w0 + bestE].g;
  const PolyRec* p = &nav.polys[g];
  const float s = uniform01(seed, query, drawBase + *drawsUsed);
  const float t = uniform01(seed, query, drawBase + *drawsUsed + 1);
  *drawsUsed += 2;
  float pt[3];
  randomPointInConvexPoly(p->v, p->nv, s, t, pt);
  float cp[3];
  bool over;
  closestPointOnPoly(nav, p, pt, cp, &over);  // DQ.cpp:309 (isValidPolyRef && finite assumed)
  if (vfinite(pt)) vcopy(outPt, cp);
  else vcopy(outPt, pt);
  return g;
}

// ---------------------------------------------------------------------------------------
// getRandomNavigablePointInCircle (PF.cpp:1283-1332; Python get_random_navigable_point_near):
// findRandomPoint with a filter that additionally excludes, among the polys that belong to an
// island, those off the requested island and those with no "edge" within the radius of the
// circle centre -- as IslandSystem::setPolyFlagForIslandCircle tests it (PF.cpp:345-393, trap
// T6: the edge (v[i], v[(i+1) % 3]) whatever the vertex count).  The passing set depends on the
// query, so the running area sums cannot be tabulated: the filter is evaluated lane-parallel over
// the tile's ground polys, the f32 area sum and the reservoir draw replayed in poly order.
// ---------------------------------------------------------------------------------------
HBN_HD float nanF();  // defined below

HBN_HD bool polyInCircleRange(const PolyRec* p, const float* center, float radSqr) {
  for (int i = 0; i < p->nv; ++i) {
    const int n = (i + 1) % 3;
    float t;
    if (distPtSegSqr2D(center, &p->v[i * 3], &p->v[n * 3], t) < radSqr) return true;
  }
  return false;
}

template <class G>
HBN_HD uint32_t findRandomPointCircle(const NavView& nav, const G& grp, uint64_t seed, uint64_t query,
                                      uint32_t drawBase, int island, const float* center, float radSqr,
                                      float* outPt, uint32_t* drawsUsed) {
  constexpr int W = G::kWidth;
  const int lane = grp.lane();
  int chosen = -1;
  {  // tile reservoir, as findRandomPoint (the filter plays no part in it, DQ.cpp:236-251)
    int best = -1;
    uint32_t k = 0;
    for (uint32_t base = 0; base < nav.numTiles; base += W) {
      const uint32_t ti = base + lane;
      const bool present = ti < nav.numTiles && nav.tiles[ti].pad[0] != 0;
      const uint32_t pm = grp.ballot(present);
      if (present) {
        const uint32_t rank = k + popc32(pm & ((1u << lane) - 1u));
        const float tsum = static_cast<float>(rank + 1);
        const float u = uniform01(seed, query, drawBase + rank);
        if (u * tsum <= 1.0f) best = static_cast<int>(ti);
      }
      k += popc32(pm);
    }
    for (int off = W / 2; off > 0; off >>= 1) {
      const int o = grp.shfl(best, lane ^ off);
      best = o > best ? o : best;
    }
    chosen = best;
    *drawsUsed = k;
  }
  if (chosen < 0) return kNoPoly;
  const TileRec& tr = nav.tiles[chosen];
  const uint32_t w0 = tr.randStart, wn = tr.randCount;  // ground polys passing the default filter
  uint32_t draws = *drawsUsed;
  float areaSum = 0.0f;
  uint32_t pick = kNoPoly;
  for (uint32_t base = 0; base < wn; base += W) {
    const uint32_t e = base + lane;
    bool pass = false;
    uint32_t g = kNoPoly;
    float area = 0.f;
    if (e < wn) {
      const RandEntry re = nav.randEntries[w0 + e];
      g = re.g;
      area = re.area;
      const PolyRec* p = &nav.polys[g];
      // polys without an island keep their flags (they are in no island's list)
      pass = p->island < 0 || ((island < 0 || p->island == island) && polyInCircleRange(p, center, radSqr));
    }
    uint32_t m = grp.ballot(pass);
    while (m) {  // DQ.cpp:262-283 over the passing polys, in poly order
      const int j = ffs32(m) - 1;
      m &= m - 1;
      const float aj = grp.shfl(area, j);
      const uint32_t gj = grp.shfl(g, j);
      areaSum += aj;
      const float u = uniform01(seed, query, drawBase + draws);
      draws++;
      if (u * areaSum <= aj) pick = gj;
    }
  }
  *drawsUsed = draws;
  if (pick == kNoPoly) return kNoPoly;
  const PolyRec* p = &nav.polys[pick];
  const float s = uniform01(seed, query, drawBase + draws);
  const float t = uniform01(seed, query, drawBase + draws + 1);
  *drawsUsed = draws + 2;
  float pt[3];
  randomPointInConvexPoly(p->v, p->nv, s, t, pt);
  float cp[3];
  bool over;
  closestPointOnPoly(nav, p, pt, cp, &over);
  if (vfinite(pt)) vcopy(outPt, cp);
  else vcopy(outPt, pt);
  return pick;
}

// the retry loop of PF.cpp:1304-1318; returns the poly, outPt = NaN on failure
template <class G>
HBN_HD uint32_t randomPointInCircle(const NavView& nav, const G& grp, uint64_t seed, uint64_t query, int island,
                                    const float* center, float radius, int maxTries, float* outPt) {
  const float radSqr = radius * radius;
  uint32_t draw = 0;
  for (int i = 0; i < maxTries; ++i) {
    uint32_t used = 0;
    float p[3];
    const uint32_t g = findRandomPointCircle(nav, grp, seed, query, draw, island, center, radSqr, p, &used);
    draw += used;
    if (g != kNoPoly) {
      const float xd = center[0] - p[0], yd = center[2] - p[2];
      const float d2 = xd * xd + yd * yd;
      if (d2 < radSqr) {
        vcopy(outPt, p);
        return g;
      }
    }
  }
  outPt[0] = outPt[1] = outPt[2] = nanF();
  return kNoPoly;
}

// =======================================================================================
// Per-query pipelines (the esp::nav::PathFinder::Impl layer, PF.cpp), one lane each.
// =======================================================================================
struct PathResult {
  float dist;        // geodesic distance, +inf if no path (PF.cpp:1473)
  int32_t npts;      // number of path points (0 if no path)
  int32_t ncorridor; // polys in the corridor handed to the funnel (0 if A* did not run/ fail)
  uint32_t astarStatus, straightStatus;
  int32_t nodesUsed;
  uint32_t expanded, links, neighbours, corridorLinks;  // work counters
  uint32_t flags;    // bit0 trivial, bit1 connected, bit2 found
  bool overflow;     // workspace tier too small: rerun with a larger one
};

HBN_HD float infF() {
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7f800000);
#else
  return INFINITY;
#endif
}
HBN_HD float nanF() {
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7fc00000);
#else
  return NAN;
#endif
}

// findPathInternal, PF.cpp:1426-1468, given the two projectToPoly results.
// reqStart/reqEnd: requested (unsnapped) points -> funnel (trap T2); sPt/ePt: snapped -> A*.
// The workspace hash table must be zero on entry.  outCorridor (nullable) receives refs.
HBN_HD PathResult findPathInternal(const NavView& nav, const AStarWs& w, const float* reqStart,
                                   const float* reqEnd, uint32_t sG, const float* sPt,
                                   uint32_t eG, const float* ePt, bool fastFail,
                                   float* outPts, int maxPts, uint32_t* outCorridor,
                                   bool countWork = false) {
  PathResult r;
  r.dist = infF();
  r.npts = 0;
  r.ncorridor = 0;
  r.astarStatus = 0;
  r.straightStatus = 0;
  r.nodesUsed = 0;
  r.expanded = r.links = r.neighbours = r.corridorLinks = 0;
  r.flags = 0;
  r.overflow = false;
  if (sG == kNoPoly || eG == kNoPoly) return r;
  if (vfuzzyEq(sPt, ePt)) {  // PF.cpp:1434-1436 (Magnum fuzzy ==)
    r.flags |= 1u | 4u;
    r.dist = 0.0f;
    r.npts = 2;
    if (outPts) {
      if (maxPts > 0) { outPts[0] = sPt[0]; outPts[1] = sPt[1]; outPts[2] = sPt[2]; }
      if (maxPts > 1) { outPts[3] = ePt[0]; outPts[4] = ePt[1]; outPts[5] = ePt[2]; }
    }
    return r;
  }
  const int32_t si = nav.polys[sG].island, ei = nav.polys[eG].island;
  if (si < 0 || si != ei) return r;  // hasConnection, PF.cpp:209-221
  r.flags |= 2u;
  uint32_t* path = reinterpret_cast<uint32_t*>(w.hkey);
  int npath = 0;
  if (sG == eG) {  // DQ.cpp:996-1001
    path[0] = sG;
    npath = 1;
    r.astarStatus = kDtSuccess;
  } else {
    if (!vfinite(sPt) || !vfinite(ePt)) { r.astarStatus = kDtFailure | kDtInvalidParam; return r; }
    const AStarResult a = astarSearch(nav, w, sG, eG, sPt, ePt, fastFail);
    if (a.status == 0xffffffffu) {
      r.overflow = true;
      return r;
    }
    r.nodesUsed = a.nodeCount;
    r.expanded = a.expanded;
    r.links = a.links;
    r.neighbours = a.neighbours;
    int fullLen = 0;
    npath = astarExtractPath(w, a.lastBest, path, kMaxPathPolys < w.cap ? kMaxPathPolys : w.cap, &fullLen);
    r.astarStatus = a.status | ((fullLen > kMaxPathPolys) ? kDtBufferTooSmall : 0u);
  }
  r.ncorridor = npath;
  if (outCorridor || countWork)
    for (int i = 0; i < npath; ++i) {
      const PolyRec* cp = &nav.polys[path[i]];
      r.corridorLinks += cp->linkCount;
      if (outCorridor) outCorridor[i] = cp->ref;
    }
  if (r.astarStatus != kDtSuccess || npath == 0) return r;  // PF.cpp:1450
  uint32_t* pathLink = w.hash;
  for (int i = 0; i + 1 < npath; ++i) pathLink[i] = findLinkTo(nav, path[i], path[i + 1]);
  Funnel f;
  f.out = outPts;
  f.maxOut = maxPts;
  r.straightStatus = funnelStraightPath(nav, reqStart, reqEnd, path, pathLink, npath, f);
  if (r.straightStatus != kDtSuccess || f.count == 0) {  // PF.cpp:1459
    r.npts = f.count;
    return r;
  }
  r.npts = f.count;
  r.dist = f.length;
  r.flags |= 4u;
  return r;
}

// ---------------------------------------------------------------------------------------
// findPath(MultiGoalShortestPath&), PF.cpp:1515-1572, for a freshly constructed path object.
// The goals are visited in the order std::sort leaves them in (PF.cpp:1542-1548, keys =
// minTheoreticalDist); std::sort is not stable, so goals with EQUAL bounds (e.g. one goal
// listed twice) come out in an order that is a property of libstdc++'s introsort.  That
// order decides closestEndPointIndex between equally distant goals, hence the sort is
// restated here operation for operation (bits/stl_algo.h: __introsort_loop, _S_threshold 16,
// __move_median_to_first, __unguarded_partition, __final_insertion_sort; heapsort fallback
// __partial_sort = __heap_select + __sort_heap over bits/stl_heap.h).
// ---------------------------------------------------------------------------------------
struct SortCtx {
  int32_t* v;        // the permutation being sorted
  const float* key;  // comp(a, b) = key[a] < key[b]
};
HBN_HD bool sortLess(const SortCtx& c, int32_t a, int32_t b) { return c.key[a] < c.key[b]; }
HBN_HD void sortSwap(const SortCtx& c, int i, int j) {
  const int32_t t = c.v[i];
  c.v[i] = c.v[j];
  c.v[j] = t;
}
// std::__adjust_heap + __push_heap (stl_heap.h) on v[first..), value semantics
HBN_HD void sortAdjustHeap(const SortCtx& c, int first, int holeIndex, int len, int32_t value) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (sortLess(c, c.v[first + secondChild], c.v[first + (secondChild - 1)])) secondChild--;
    c.v[first + holeIndex] = c.v[first + secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    c.v[first + holeIndex] = c.v[first + (secondChild - 1)];
    holeIndex = secondChild - 1;
  }
  int parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && sortLess(c, c.v[first + parent], value)) {
    c.v[first + holeIndex] = c.v[first + parent];
    holeIndex = parent;
    parent = (holeIndex - 1) / 2;
  }
  c.v[first + holeIndex] = value;
}
// std::__partial_sort(first, last, last): heapsort of v[first, last)
HBN_HD void sortHeapSort(const SortCtx& c, int first, int last) {
  const int len = last - first;
  if (len >= 2) {  // __make_heap
    int parent = (len - 2) / 2;
    for (;;) {
      const int32_t value = c.v[first + parent];
      sortAdjustHeap(c, first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  // __heap_select's loop over [middle, last) is empty (middle == last); __sort_heap:
  for (int l = last; l - first > 1;) {
    --l;
    const int32_t value = c.v[l];  // __pop_heap(first, l, l)
    c.v[l] = c.v[first];
    sortAdjustHeap(c, first, 0, l - first, value);
  }
}
HBN_HD void sortUnguardedLinearInsert(const SortCtx& c, int last) {
  const int32_t val = c.v[last];
  int next = last - 1;
  while (sortLess(c, val, c.v[next])) {
    c.v[last] = c.v[next];
    last = next;
    --next;
  }
  c.v[last] = val;
}
HBN_HD void sortInsertion(const SortCtx& c, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (sortLess(c, c.v[i], c.v[first])) {
      const int32_t val = c.v[i];
      for (int k = i; k > first; --k) c.v[k] = c.v[k - 1];  // move_backward
      c.v[first] = val;
    } else {
      sortUnguardedLinearInsert(c, i);
    }
  }
}
// std::sort(v, v + n, comp)
HBN_HD void stdSortOrder(int32_t* v, const float* key, int n) {
  if (n <= 0) return;
  const SortCtx c{v, key};
  // __introsort_loop, with its tail recursion on [cut, last) turned into an explicit stack
  int depth0 = 0;
  for (int t = n; t > 1; t >>= 1) depth0++;  // std::__lg(n)
  depth0 *= 2;
  int stFirst[64], stLast[64], stDepth[64];
  int sp = 0;
  stFirst[0] = 0; stLast[0] = n; stDepth[0] = depth0;
  sp = 1;
  while (sp > 0) {
    --sp;
    const int first = stFirst[sp];
    int last = stLast[sp];
    int depth = stDepth[sp];
    // the recursive calls of one frame are issued inside its while loop, in order; each
    // only touches [cut, last), disjoint from what the frame does next, so running them
    // after the frame (LIFO) gives the same array
    while (last - first > 16) {
      if (depth == 0) {
        sortHeapSort(c, first, last);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      {  // __move_median_to_first(first, first + 1, mid, last - 1)
        const int a = first + 1, b = mid, cc = last - 1;
        if (sortLess(c, c.v[a], c.v[b])) {
          if (sortLess(c, c.v[b], c.v[cc])) sortSwap(c, first, b);
          else if (sortLess(c, c.v[a], c.v[cc])) sortSwap(c, first, cc);
          else sortSwap(c, first, a);
        } else if (sortLess(c, c.v[a], c.v[cc])) {
          sortSwap(c, first, a);
        } else if (sortLess(c, c.v[b], c.v[cc])) {
          sortSwap(c, first, cc);
        } else {
          sortSwap(c, first, b);
        }
      }
      int lo = first + 1, hi = last;  // __unguarded_partition(first + 1, last, first)
      for (;;) {
        while (sortLess(c, c.v[lo], c.v[first])) ++lo;
        --hi;
        while (sortLess(c, c.v[first], c.v[hi])) --hi;
        if (!(lo < hi)) break;
        sortSwap(c, lo, hi);
        ++lo;
      }
      if (sp < 64) {
        stFirst[sp] = lo; stLast[sp] = last; stDepth[sp] = depth;
        ++sp;
      }
      last = lo;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    sortInsertion(c, 0, 16);
    for (int i = 16; i != n; ++i) sortUnguardedLinearInsert(c, i);
  } else {
    sortInsertion(c, 0, n);
  }
}

// The goal loop of PF.cpp:1515-1572 given every goal's findPathInternal result dist[i]
// (+inf = no path), evaluated for a fresh MultiGoalShortestPath: bounds are the L2 distances
// requestedEnd - requestedStart when there are several goals (max(0 - movedAmount, L2) with
// movedAmount >= 0 or +inf), zero for a single goal.  bounds / order: scratch of n entries.
HBN_HD void multiGoalSelect(int n, const float* start, bool startValid, const float* ends,
                            const uint32_t* endG, const float* dist, float* bounds,
                            int32_t* order, float* outDist, int32_t* outIndex) {
  *outDist = infF();
  *outIndex = -1;
  if (!startValid || n <= 0) return;  // findPathSetup, PF.cpp:1483-1485
  bool anyValid = false;
  for (int i = 0; i < n; ++i) anyValid = anyValid || endG[i] != kNoPoly;
  if (!anyValid) return;  // PF.cpp:1507-1510
  for (int i = 0; i < n; ++i) {
    bounds[i] = n > 1 ? mnDist(&ends[3 * i], start) : 0.0f;
    order[i] = i;
  }
  stdSortOrder(order, bounds, n);
  float best = infF();
  int32_t bestIdx = -1;
  for (int k = 0; k < n; ++k) {
    const int i = order[k];
    if (endG[i] == kNoPoly) continue;
    if (bounds[i] > best) continue;
    const float d = dist[i];
    if (d < infF() && d < best) {  // findResult && distance < geodesicDistance
      best = d;
      bestIdx = i;
    }
  }
  *outDist = best;
  *outIndex = bestIdx;
}

// ---- goal pruning for batched multi-goal queries ------------------------------------------
// The reference visits the goals in the order of their lower bounds and skips a goal whose bound
// exceeds the best distance so far (PF.cpp:1541-1569); on a batch the pair searches dominate, so
// they are done in two rounds: round 1 searches the kMultiGoalFirst goals of smallest bound of every
// start; round 2 only those later goals whose bound does not exceed the reference's running best
// after those first goals (`multiGoalRunningBest`).  The running best only shrinks, so a goal left
// out is one the reference skips without looking at its distance; multiGoalSelect then replays the
// reference's loop over distances that are real wherever it reads them.
constexpr int kMultiGoalFirst = 8;

// bounds[], order[] as multiGoalSelect computes them; returns false if the query has no result
HBN_HD bool multiGoalOrder(int n, const float* start, bool startValid, const float* ends,
                           const uint32_t* endG, float* bounds, int32_t* order) {
  if (!startValid || n <= 0) return false;
  bool anyValid = false;
  for (int i = 0; i < n; ++i) anyValid = anyValid || endG[i] != kNoPoly;
  if (!anyValid) return false;
  for (int i = 0; i < n; ++i) {
    bounds[i] = n > 1 ? mnDist(&ends[3 * i], start) : 0.0f;
    order[i] = i;
  }
  stdSortOrder(order, bounds, n);
  return true;
}

// the reference's `best` after the first `first` goals of the sorted order (dist[] valid for them)
HBN_HD float multiGoalRunningBest(int n, int first, const uint32_t* endG, const float* dist,
                                  const float* bounds, const int32_t* order) {
  float best = infF();
  for (int k = 0; k < n && k < first; ++k) {
    const int i = order[k];
    if (endG[i] == kNoPoly) continue;
    if (bounds[i] > best) continue;
    const float d = dist[i];
    if (d < infF() && d < best) best = d;
  }
  return best;
}

// PathFinder::Impl::tryStep, PF.cpp:1575-1722, phase A: everything up to and including
// getPolyHeight (:1687).  Inputs are the projectToPoly results of start and end.
// Returns false if tryStep returns `start` (PF.cpp:1587-1604); else endPoint/lastPoly/startG
// are handed to phase B after endPoint has been re-projected.
HBN_HD bool tryStepPhaseA(const NavView& nav, uint32_t sG, const float* sPt, uint32_t eG,
                          const float* end, bool allowSliding, float* endPoint,
                          uint32_t* lastPoly) {
  if (sG == kNoPoly || eG == kNoPoly) return false;
  const int32_t si = nav.polys[sG].island, ei = nav.polys[eG].island;
  if (si < 0 || si != ei) return false;
  uint32_t polys[kMaxPathPolys > 64 ? 64 : kMaxPathPolys];  // visited <= 64 tiny-pool nodes
  TinyPool tp;
  const int numPolys = moveAlongSurface(nav, sG, sPt, end, endPoint, polys, 64, tp);
  if (numPolys == 0) return false;
  if (!allowSliding) noSlidingClamp(nav, polys, numPolys, sPt, end, endPoint);
  float h;
  if (queryPolyHeight(nav, &nav.polys[polys[numPolys - 1]], endPoint, &h)) endPoint[1] = h;
  *lastPoly = polys[numPolys - 1];
  return true;
}
// phase B: the connected-component nudge, PF.cpp:1694-1719.  e2G = projectToPoly(endPoint).
HBN_HD void tryStepPhaseB(const NavView& nav, uint32_t sG, uint32_t e2G, uint32_t lastPoly,
                          float* endPoint) {
  const int32_t si = nav.polys[sG].island;
  const int32_t ei = (e2G == kNoPoly) ? -2 : nav.polys[e2G].island;
  if (ei >= 0 && si == ei) return;
  const PolyRec* p = &nav.polys[lastPoly];
  float c[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < p->nv; ++i) {
    c[0] += p->v[i * 3];
    c[1] += p->v[i * 3 + 1];
    c[2] += p->v[i * 3 + 2];
  }
  const float n = static_cast<float>(p->nv);
  c[0] /= n; c[1] /= n; c[2] /= n;
  const float dx = c[0] - endPoint[0], dy = c[1] - endPoint[1], dz = c[2] - endPoint[2];
  float d2 = 0.f;
  d2 += dx * dx;
  d2 += dy * dy;
  d2 += dz * dz;
  const float inv = 1.0f / fsqrt(d2);
  const float nudge = 1e-4f;
  endPoint[0] = endPoint[0] + nudge * (dx * inv);
  endPoint[1] = endPoint[1] + nudge * (dy * inv);
  endPoint[2] = endPoint[2] + nudge * (dz * inv);
}

}  // namespace hbn

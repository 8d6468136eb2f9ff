// =============================================================================
// oracle/ref_pathfinder.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE, NOT THE ORACLE.
//
// The round-1 RESTATEMENT of the thin esp::nav::PathFinder::Impl layer (src/esp/nav/PathFinder.cpp,
// "PF.cpp" below; each function cites the range it follows) over the reference's own Detour + Recast.
// Since round 2 the oracle is the reference's PathFinder.cpp itself (oracle/ref_esp.cpp ->
// oracle/_ref/libhbn_ref.so).  This file is built into oracle/_ref/libhbn_restated.so and loaded by
// exactly one test, tests/test_oracle.py::test_restatement_equals_reference_pathfinder, which shows
// the two agree bit for bit on the five scenes -- i.e. that what round 1 checked against this
// restatement also held against the reference.  No parity test, smoke() or bench leg uses it.
// =============================================================================
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <stack>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "DetourCommon.h"
#include "DetourNavMesh.h"
#include "DetourNavMeshBuilder.h"
#include "DetourNavMeshQuery.h"
#include "DetourNode.h"
#include "Recast.h"

#include "tiled_recast.h"

namespace {
using namespace hbnoracle;

// ---- Magnum::Vector3 operations used by PF.cpp, same operation order --------
// (src/deps/magnum/src/Magnum/Math/Vector.h:106-111 dot, :997 length,
//  :1023 normalized = v * (1/length), TypeTraits.h:495-510 fuzzy ==, eps 1e-5f)
struct V3 {
  float x = 0.f, y = 0.f, z = 0.f;
  V3() = default;
  V3(float a, float b, float c) : x(a), y(b), z(c) {}
  explicit V3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
  float* data() { return &x; }
  const float* data() const { return &x; }
  float operator[](int i) const { return (&x)[i]; }
  float& operator[](int i) { return (&x)[i]; }
};
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(float s, const V3& a) { return {a.x * s, a.y * s, a.z * s}; }
inline float mnDot(const V3& a, const V3& b) {
  float out = 0.f;
  out += a.x * b.x;
  out += a.y * b.y;
  out += a.z * b.z;
  return out;
}
inline float mnLength(const V3& a) { return std::sqrt(mnDot(a, a)); }
inline V3 mnCross(const V3& a, const V3& b) {
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline bool mnFuzzyEq(float a, float b) {
  if (a == b) return true;
  const float absA = std::abs(a), absB = std::abs(b), diff = std::abs(a - b);
  const float eps = 1.0e-5f;
  if (a == 0.f || b == 0.f || diff < eps) return diff < eps;
  return diff / (absA + absB) < eps;
}
inline bool mnFuzzyEq(const V3& a, const V3& b) {
  return mnFuzzyEq(a.x, b.x) && mnFuzzyEq(a.y, b.y) && mnFuzzyEq(a.z, b.z);
}
const float kNaN = std::numeric_limits<float>::quiet_NaN();
const float kInf = std::numeric_limits<float>::infinity();


const int ID_UNDEFINED = -1;  // core/Esp.h:100

struct RandStream {
  int mode = 0;  // 0 = glibc rand(), 1 = counter based
  uint64_t seed = 0, query = 0;
  uint32_t draw = 0;
};
thread_local RandStream tlsRand;
float frandTls() {
  if (tlsRand.mode == 0)
    return static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
  return hbnUniform(tlsRand.seed, tlsRand.query, tlsRand.draw++);
}

// ---- IslandSystem: PF.cpp:167-460 -------------------------------------------
class IslandSystem {
 public:
  IslandSystem(const dtNavMesh* navMesh, const dtQueryFilter* filter) {
    std::vector<V3> islandVerts;
    for (int iTile = 0; iTile < navMesh->getMaxTiles(); ++iTile) {
      const dtMeshTile* tile = navMesh->getTile(iTile);
      if (!tile || !tile->header) continue;  // (reference derefs header: T10)
      for (int jPoly = 0; jPoly < tile->header->polyCount; ++jPoly) {
        dtPolyRef startRef = navMesh->encodePolyId(tile->salt, iTile, jPoly);
        if (navMesh->isValidPolyRef(startRef) &&
            (polyToIsland_.find(startRef) == polyToIsland_.end())) {
          uint32_t newIslandId = islandRadius_.size();
          expandFrom(navMesh, filter, newIslandId, startRef, islandVerts);
          V3 centroid;
          for (auto& v : islandVerts) centroid = centroid + v;
          const float n = static_cast<float>(islandVerts.size());
          centroid = V3(centroid.x / n, centroid.y / n, centroid.z / n);
          float maxRadius = 0.0;
          for (auto& v : islandVerts)
            maxRadius = std::max(maxRadius, mnLength(v - centroid));
          islandRadius_.emplace_back(maxRadius);
        }
      }
    }
  }
  bool hasConnection(dtPolyRef a, dtPolyRef b) const {
    auto ia = polyToIsland_.find(a);
    if (ia == polyToIsland_.end()) return false;
    auto ib = polyToIsland_.find(b);
    if (ib == polyToIsland_.end()) return false;
    return ia->second == ib->second;
  }
  bool validIsland(int islandIndex, bool indexOptional = true) const {
    if (indexOptional && islandIndex == ID_UNDEFINED) return true;
    return islandIndex >= 0 && islandIndex < static_cast<int>(islandRadius_.size());
  }
  float islandRadius(int i) const { return islandRadius_[i]; }
  float polyIslandRadius(dtPolyRef ref) const {
    auto it = polyToIsland_.find(ref);
    if (it == polyToIsland_.end()) return 0.0;
    return islandRadius_[it->second];
  }
  float getNavigableArea(int islandIndex) { return islandsToArea_[islandIndex]; }
  int numIslands() const { return islandRadius_.size(); }
  int getPolyIsland(dtPolyRef ref) { return polyToIsland_[ref]; }
  int findPolyIsland(dtPolyRef ref) const {
    auto it = polyToIsland_.find(ref);
    return it == polyToIsland_.end() ? -1 : static_cast<int>(it->second);
  }

  // PF.cpp:272-316
  void setPolyFlagForIsland(dtNavMesh* navMesh, unsigned short flag, int islandIndex,
                            bool setFlag, bool invert) {
    std::vector<int> islands;
    if (islandIndex == ID_UNDEFINED) {
      for (auto& itr : islandsToPolys_) islands.push_back(itr.first);
    } else if (invert) {
      for (auto& itr : islandsToPolys_)
        if (static_cast<int>(itr.first) != islandIndex) islands.push_back(itr.first);
    } else {
      islands.push_back(islandIndex);
    }
    const unsigned short modFlag = setFlag ? flag : ~flag;
    for (int island : islands) {
      for (auto& polyRef : islandsToPolys_[island]) {
        unsigned short f = 0;
        navMesh->getPolyFlags(polyRef, &f);
        navMesh->setPolyFlags(polyRef, setFlag ? (f | modFlag) : (f & modFlag));
      }
    }
  }

  // PF.cpp:329-393 (trap T6: (iVert+1)%3 regardless of vertCount)
  void setPolyFlagForIslandCircle(dtNavMesh* navMesh, unsigned short flag,
                                  const V3& circleCenter, const float radius,
                                  int islandIndex) {
    std::vector<int> islands;
    float radSqr = radius * radius;
    for (auto& itr : islandsToPolys_) islands.push_back(itr.first);
    for (int island : islands) {
      for (auto& polyRef : islandsToPolys_[island]) {
        if (islandIndex != ID_UNDEFINED && islandIndex != island) {
          unsigned short f = 0;
          navMesh->getPolyFlags(polyRef, &f);
          navMesh->setPolyFlags(polyRef, f | flag);
          continue;
        }
        const dtMeshTile* tile = nullptr;
        const dtPoly* poly = nullptr;
        navMesh->getTileAndPolyByRefUnsafe(polyRef, &tile, &poly);
        bool inRange = false;
        for (int iVert = 0; iVert < poly->vertCount; ++iVert) {
          int nVert = (iVert + 1) % 3;
          float tseg = 0;
          float distSqr = dtDistancePtSegSqr2D(
              circleCenter.data(), &tile->verts[static_cast<size_t>(poly->verts[iVert]) * 3],
              &tile->verts[static_cast<size_t>(poly->verts[nVert]) * 3], tseg);
          if (distSqr < radSqr) {
            inRange = true;
            break;
          }
        }
        if (!inRange) {
          unsigned short f = 0;
          navMesh->getPolyFlags(polyRef, &f);
          navMesh->setPolyFlags(polyRef, f | flag);
        }
      }
    }
  }

  // PF.cpp:1001-1085
  static float polyArea(const dtPoly* poly, const dtMeshTile* tile) {
    const std::ptrdiff_t ip = poly - tile->polys;
    const dtPolyDetail* pd = &tile->detailMeshes[ip];
    float area = 0;
    for (int j = 0; j < pd->triCount; ++j) {
      const unsigned char* t = &tile->detailTris[static_cast<size_t>((pd->triBase + j)) * 4];
      V3 v[3];
      for (int k = 0; k < 3; ++k) {
        if (t[k] < poly->vertCount)
          v[k] = V3(&tile->verts[static_cast<size_t>(poly->verts[t[k]]) * 3]);
        else
          v[k] = V3(&tile->detailVerts[static_cast<size_t>(
                                           (pd->vertBase + (t[k] - poly->vertCount))) * 3]);
      }
      const V3 w1 = v[1] - v[0];
      const V3 w2 = v[2] - v[1];
      area += 0.5f * mnLength(mnCross(w1, w2));
    }
    return area;
  }
  void removeZeroAreaPolys(dtNavMesh* navMesh) {
    islandsToArea_ = std::unordered_map<uint32_t, float>();
    islandsToArea_.reserve(islandsToPolys_.size());
    for (auto& itr : islandsToPolys_) islandsToArea_[itr.first] = 0.0;
    for (int iTile = 0; iTile < navMesh->getMaxTiles(); ++iTile) {
      const dtMeshTile* tile = const_cast<const dtNavMesh*>(navMesh)->getTile(iTile);
      if (!tile || !tile->header) continue;
      for (int jPoly = 0; jPoly < tile->header->polyCount; ++jPoly) {
        dtPolyRef polyRef = navMesh->encodePolyId(tile->salt, iTile, jPoly);
        const dtPoly* poly = nullptr;
        const dtMeshTile* tmp = nullptr;
        navMesh->getTileAndPolyByRefUnsafe(polyRef, &tmp, &poly);
        float polygonArea = polyArea(poly, tile);
        if (polygonArea < 1e-5f) {
          navMesh->setPolyFlags(polyRef, POLYFLAGS_DISABLED);
        } else if ((poly->flags & POLYFLAGS_WALK) != 0) {
          islandsToArea_[polyToIsland_[polyRef]] += polygonArea;
        }
      }
    }
    float totalArea = 0;
    for (auto& itr : islandsToArea_) totalArea += itr.second;
    islandsToArea_[ID_UNDEFINED] = totalArea;
  }

 private:
  std::unordered_map<uint32_t, float> islandsToArea_;
  std::unordered_map<uint32_t, std::vector<dtPolyRef>> islandsToPolys_;
  std::unordered_map<dtPolyRef, uint32_t> polyToIsland_;
  std::vector<float> islandRadius_;

  // PF.cpp:409-459
  void expandFrom(const dtNavMesh* navMesh, const dtQueryFilter* filter,
                  const uint32_t newIslandId, const dtPolyRef& startRef,
                  std::vector<V3>& islandVerts) {
    islandsToPolys_[newIslandId].push_back(startRef);
    polyToIsland_.emplace(startRef, newIslandId);
    islandVerts.clear();
    std::stack<dtPolyRef, std::vector<dtPolyRef>> stack;
    stack.push(startRef);
    while (!stack.empty()) {
      dtPolyRef ref = stack.top();
      stack.pop();
      const dtMeshTile* tile = nullptr;
      const dtPoly* poly = nullptr;
      navMesh->getTileAndPolyByRefUnsafe(ref, &tile, &poly);
      for (int iVert = 0; iVert < poly->vertCount; ++iVert)
        islandVerts.emplace_back(V3(&tile->verts[static_cast<size_t>(poly->verts[iVert]) * 3]));
      for (unsigned int iLink = poly->firstLink; iLink != DT_NULL_LINK;
           iLink = tile->links[iLink].next) {
        dtPolyRef neighbourRef = tile->links[iLink].ref;
        if (polyToIsland_.find(neighbourRef) != polyToIsland_.end()) continue;
        const dtMeshTile* neighbourTile = nullptr;
        const dtPoly* neighbourPoly = nullptr;
        navMesh->getTileAndPolyByRefUnsafe(neighbourRef, &neighbourTile, &neighbourPoly);
        if (!filter->passFilter(neighbourRef, neighbourTile, neighbourPoly)) continue;
        polyToIsland_.emplace(neighbourRef, newIslandId);
        islandsToPolys_[newIslandId].push_back(neighbourRef);
        stack.push(neighbourRef);
      }
    }
  }
};

// Per-query work counters used for SURVEY §8(d)'s algorithmic-bytes formulae.
struct Counters {
  uint64_t n_queries = 0;
  uint64_t nodes_used = 0;  // A* node-pool population after findPath
};

struct HitRecord {
  V3 hitPos, hitNormal;
  float hitDist;
};


// ---- the PathFinder restatement ----------------------------------------------
class RefPathFinder {
 public:
  RefPathFinder() {
    // PF.cpp:606-610
    filter_.setIncludeFlags(POLYFLAGS_WALK);
    filter_.setExcludeFlags(0);
  }
  ~RefPathFinder() {
    for (auto* q : extraQueries_) dtFreeNavMeshQuery(q);
    if (navQuery_) dtFreeNavMeshQuery(navQuery_);
    if (navMesh_) dtFreeNavMesh(navMesh_);
  }

  // PF.cpp:952-975 + :612-930
  bool build(const NavMeshSettings& bs, const float* verts, int nverts, const int* tris,
             int ntris) {
    const float mf = std::numeric_limits<float>::max();
    float bmin[3] = {mf, mf, mf}, bmax[3] = {-mf, -mf, -mf};
    for (int i = 0; i < nverts; ++i)
      for (int k = 0; k < 3; ++k) {
        bmin[k] = std::min(bmin[k], verts[i * 3 + k]);
        bmax[k] = std::max(bmax[k], verts[i * 3 + k]);
      }
    BuildOut out;
    if (!recastBuildOne(bs, verts, nverts, tris, ntris, bmin, bmax, false, 0, 0, 0, out))
      return false;
    reset();
    navMesh_ = dtAllocNavMesh();
    dtStatus status = navMesh_->init(out.navData, out.navDataSize, DT_TILE_FREE_DATA);
    if (dtStatusFailed(status)) {
      dtFree(out.navData);
      return false;
    }
    if (!initNavQuery()) return false;
    settings_ = bs;
    bounds_[0] = V3(bmin);
    bounds_[1] = V3(bmax);
    return true;
  }

  // Tiled host build (not a PathFinder method): oracle/tiled_recast.h makes the
  // MSET image, which is then loaded like any .navmesh file (PF.cpp:1091-1175).
  bool buildTiled(const NavMeshSettings& bs, const float* verts, int nverts,
                  const int* tris, int ntris, int tileSize, int nthreads) {
    std::vector<unsigned char> image;
    if (!buildTiledImage(bs, verts, nverts, tris, ntris, tileSize, nthreads, image)) return false;
    return loadFromMemory(image.data(), image.size());
  }

  // PF.cpp:1091-1175
  bool loadFromMemory(const unsigned char* buf, size_t len) {
    size_t off = 0;
    auto rd = [&](void* dst, size_t n) {
      if (off + n > len) return false;
      memcpy(dst, buf + off, n);
      off += n;
      return true;
    };
    NavMeshSetHeader header{};
    if (!rd(&header, sizeof(header))) return false;
    if (header.magic != NAVMESHSET_MAGIC) return false;
    if (header.version < 1 || header.version > NAVMESHSET_VERSION) return false;
    NavMeshSettings st;
    if (header.version >= 2) rd(&st, sizeof(st));
    dtNavMesh* mesh = dtAllocNavMesh();
    if (!mesh) return false;
    if (dtStatusFailed(mesh->init(&header.params))) {
      dtFreeNavMesh(mesh);
      return false;
    }
    V3 bmin, bmax;
    for (int i = 0; i < header.numTiles; ++i) {
      NavMeshTileHeader tileHeader{};
      if (!rd(&tileHeader, sizeof(tileHeader))) { dtFreeNavMesh(mesh); return false; }
      if ((tileHeader.tileRef == 0u) || (tileHeader.dataSize == 0)) break;
      unsigned char* data =
          static_cast<unsigned char*>(dtAlloc(tileHeader.dataSize, DT_ALLOC_PERM));
      if (!data) break;
      memset(data, 0, tileHeader.dataSize);
      if (!rd(data, tileHeader.dataSize)) { dtFree(data); dtFreeNavMesh(mesh); return false; }
      mesh->addTile(data, tileHeader.dataSize, DT_TILE_FREE_DATA, tileHeader.tileRef, nullptr);
      const dtMeshTile* tile = mesh->getTileByRef(tileHeader.tileRef);
      for (int k = 0; k < 3; ++k) {
        if (i == 0) {
          bmin[k] = tile->header->bmin[k];
          bmax[k] = tile->header->bmax[k];
        } else {
          bmin[k] = std::min(bmin[k], tile->header->bmin[k]);
          bmax[k] = std::max(bmax[k], tile->header->bmax[k]);
        }
      }
    }
    reset();
    navMesh_ = mesh;
    settings_ = st;
    bounds_[0] = bmin;
    bounds_[1] = bmax;
    return initNavQuery();
  }

  // PF.cpp:1177-1223
  bool saveToMemory(std::vector<unsigned char>& out) const {
    const dtNavMesh* navMesh = navMesh_;
    if (!navMesh) return false;
    auto wr = [&](const void* p, size_t n) {
      const unsigned char* c = static_cast<const unsigned char*>(p);
      out.insert(out.end(), c, c + n);
    };
    NavMeshSetHeader header{};
    header.magic = NAVMESHSET_MAGIC;
    header.version = NAVMESHSET_VERSION;
    header.numTiles = 0;
    for (int i = 0; i < navMesh->getMaxTiles(); ++i) {
      const dtMeshTile* tile = navMesh->getTile(i);
      if (!tile || !tile->header || (tile->dataSize == 0)) continue;
      ++header.numTiles;
    }
    memcpy(&header.params, navMesh->getParams(), sizeof(dtNavMeshParams));
    wr(&header, sizeof(header));
    wr(&settings_, sizeof(settings_));
    for (int i = 0; i < navMesh->getMaxTiles(); ++i) {
      const dtMeshTile* tile = navMesh->getTile(i);
      if (!tile || !tile->header || (tile->dataSize == 0)) continue;
      NavMeshTileHeader tileHeader{};
      tileHeader.tileRef = navMesh->getTileRef(tile);
      tileHeader.dataSize = tile->dataSize;
      wr(&tileHeader, sizeof(tileHeader));
      wr(tile->data, tile->dataSize);
    }
    return true;
  }

  bool isLoaded() const { return navMesh_ != nullptr; }
  void seed(uint32_t s) { srand(s); }  // PF.cpp:1225-1229

  // PF.cpp:126-147
  std::tuple<dtStatus, dtPolyRef, V3> projectToPoly(const V3& pt,
                                                    const dtNavMeshQuery* q) const {
    const float polyPickExt[3] = {2, 4, 2};
    dtPolyRef polyRef = 0;
    V3 polyXYZ(kNaN, kNaN, kNaN);
    dtStatus status = q->findNearestPoly(pt.data(), polyPickExt, &filter_, &polyRef,
                                         polyXYZ.data());
    if (std::isnan(polyXYZ[0])) status = DT_FAILURE;
    return std::make_tuple(status, polyRef, polyXYZ);
  }

  // PF.cpp:1236-1281; returns false where the reference throws (area <= 0)
  bool getRandomNavigablePoint(int maxTries, int islandIndex, V3& result,
                               dtPolyRef* outRef = nullptr) {
    if (!islands_->validIsland(islandIndex)) return false;
    if (islands_->getNavigableArea(islandIndex) <= 0.0f) return false;
    if (islandIndex != ID_UNDEFINED) {
      islands_->setPolyFlagForIsland(navMesh_, POLYFLAGS_OFF_ISLAND, islandIndex, true, true);
      filter_.setExcludeFlags(filter_.getExcludeFlags() | POLYFLAGS_OFF_ISLAND);
    }
    V3 pt;
    dtPolyRef ref = 0;
    int i = 0;
    for (i = 0; i < maxTries; ++i) {
      ref = 0;
      dtStatus status = navQuery_->findRandomPoint(&filter_, frandTls, &ref, pt.data());
      if (dtStatusSucceed(status)) break;
    }
    if (islandIndex != ID_UNDEFINED) {
      islands_->setPolyFlagForIsland(navMesh_, POLYFLAGS_OFF_ISLAND, islandIndex, false, true);
      filter_.setExcludeFlags(filter_.getExcludeFlags() & ~POLYFLAGS_OFF_ISLAND);
    }
    if (outRef) *outRef = (i == maxTries) ? 0 : ref;
    result = (i == maxTries) ? V3(kNaN, kNaN, kNaN) : pt;
    return true;
  }

  // PF.cpp:1283-1332
  bool getRandomNavigablePointInCircle(const V3& circleCenter, float radius, int maxTries,
                                       int islandIndex, V3& result) {
    float radSqr = radius * radius;
    if (!islands_->validIsland(islandIndex)) return false;
    if (islands_->getNavigableArea(islandIndex) <= 0.0f) return false;
    islands_->setPolyFlagForIslandCircle(navMesh_, POLYFLAGS_OFF_ISLAND, circleCenter, radius,
                                         islandIndex);
    filter_.setExcludeFlags(filter_.getExcludeFlags() | POLYFLAGS_OFF_ISLAND);
    V3 pt;
    int i = 0;
    for (i = 0; i < maxTries; ++i) {
      dtPolyRef ref = 0;
      dtStatus status = navQuery_->findRandomPoint(&filter_, frandTls, &ref, pt.data());
      if (dtStatusSucceed(status)) {
        float xd = circleCenter[0] - pt[0];
        float yd = circleCenter[2] - pt[2];
        float d2 = xd * xd + yd * yd;
        if (d2 < radSqr) break;
      }
    }
    islands_->setPolyFlagForIsland(navMesh_, POLYFLAGS_OFF_ISLAND, ID_UNDEFINED, false, true);
    filter_.setExcludeFlags(filter_.getExcludeFlags() & ~POLYFLAGS_OFF_ISLAND);
    result = (i == maxTries) ? V3(kNaN, kNaN, kNaN) : pt;
    return true;
  }

  // PF.cpp:1400-1411
  static float pathLength(const std::vector<V3>& points) {
    float length = 0;
    const V3* previousPoint = &points[0];
    for (const auto& pt : points) {
      length += mnLength(*previousPoint - pt);
      previousPoint = &pt;
    }
    return length;
  }

  struct RawPath {
    dtStatus astarStatus = 0, straightStatus = 0;
    int numPolys = 0, numPoints = 0, nodesUsed = 0;
    dtPolyRef polys[256];
    bool trivial = false, connected = false;
  };

  // PF.cpp:1426-1468.  `raw` (optional) exposes what PathFinder hides.
  bool findPathInternal(dtNavMeshQuery* q, const V3& start, dtPolyRef startRef,
                        const V3& pathStart, const V3& end, dtPolyRef endRef,
                        const V3& pathEnd, float& outLen, std::vector<V3>& outPts,
                        RawPath* raw = nullptr) const {
    if (mnFuzzyEq(pathStart, pathEnd)) {
      if (raw) raw->trivial = true;
      outLen = 0.0f;
      outPts = {pathStart, pathEnd};
      return true;
    }
    if (!islands_->hasConnection(startRef, endRef)) return false;
    if (raw) raw->connected = true;
    static const int MAX_POLYS = 256;
    dtPolyRef polysLocal[MAX_POLYS];
    dtPolyRef* polys = raw ? raw->polys : polysLocal;
    int numPolys = 0;
    dtStatus status = q->findPath(startRef, endRef, pathStart.data(), pathEnd.data(),
                                  &filter_, polys, &numPolys, MAX_POLYS);
    if (raw) {
      raw->astarStatus = status;
      raw->numPolys = numPolys;
      raw->nodesUsed = startRef == endRef ? 0 : q->getNodePool()->getNodeCount();  // (same poly: findPath returns before clearing the pool)
    }
    if (status != DT_SUCCESS || numPolys == 0) return false;
    int numPoints = 0;
    std::vector<V3> points(MAX_POLYS);
    status = q->findStraightPath(start.data(), end.data(), polys, numPolys,
                                 points[0].data(), nullptr, nullptr, &numPoints, MAX_POLYS);
    if (raw) {
      raw->straightStatus = status;
      raw->numPoints = numPoints;
    }
    if (status != DT_SUCCESS || numPoints == 0) return false;
    points.resize(numPoints);
    outLen = pathLength(points);
    outPts = std::move(points);
    return true;
  }

  // State of esp::nav::MultiGoalShortestPath (PF.h:81-123, PF.cpp:95-123)
  struct MultiGoalPath {
    V3 requestedStart;
    std::vector<V3> requestedEnds;
    std::vector<dtPolyRef> endRefs;
    std::vector<bool> endIsValid;
    std::vector<V3> pathEnds;
    std::vector<float> minTheoreticalDist;
    V3 prevRequestedStart;
    // outputs
    std::vector<V3> points;
    float geodesicDistance = kInf;
    int closestEndPointIndex = -1;
    void setRequestedEnds(const std::vector<V3>& newEnds) {
      endRefs.clear();
      pathEnds.clear();
      requestedEnds = newEnds;
      minTheoreticalDist.assign(newEnds.size(), 0);
    }
  };

  // PF.cpp:1470-1513
  bool findPathSetup(dtNavMeshQuery* q, MultiGoalPath& path, dtPolyRef& startRef,
                     V3& pathStart) const {
    path.geodesicDistance = kInf;
    path.closestEndPointIndex = -1;
    path.points.clear();
    dtStatus status = 0;
    std::tie(status, startRef, pathStart) = projectToPoly(path.requestedStart, q);
    if (status != DT_SUCCESS || startRef == 0) return false;
    if (!path.endRefs.empty()) return true;
    int numValidPoints = 0;
    for (const auto& rqEnd : path.requestedEnds) {
      dtPolyRef endRef = 0;
      V3 pathEnd;
      std::tie(status, endRef, pathEnd) = projectToPoly(rqEnd, q);
      if (status != DT_SUCCESS || endRef == 0) {
        path.endIsValid.emplace_back(false);
      } else {
        path.endIsValid.emplace_back(true);
        numValidPoints++;
      }
      path.endRefs.emplace_back(endRef);
      path.pathEnds.emplace_back(pathEnd);
    }
    return numValidPoints != 0;
  }

  // PF.cpp:1414-1424
  bool findPathSingle(dtNavMeshQuery* q, const V3& s, const V3& e, float& dist,
                      std::vector<V3>& pts) const {
    MultiGoalPath tmp;
    tmp.requestedStart = s;
    tmp.setRequestedEnds({e});
    bool status = findPathMulti(q, tmp);
    dist = tmp.geodesicDistance;
    pts = std::move(tmp.points);
    return status;
  }

  // PF.cpp:1515-1572
  bool findPathMulti(dtNavMeshQuery* q, MultiGoalPath& path) const {
    dtPolyRef startRef = 0;
    V3 pathStart;
    if (!findPathSetup(q, path, startRef, pathStart)) return false;
    if (path.requestedEnds.size() > 1) {
      float movedAmount;
      std::vector<V3> dummy;
      findPathSingle(q, path.requestedStart, path.prevRequestedStart, movedAmount, dummy);
      for (std::size_t i = 0; i < path.requestedEnds.size(); ++i) {
        path.minTheoreticalDist[i] =
            std::max(path.minTheoreticalDist[i] - movedAmount,
                     mnLength(path.requestedEnds[i] - path.requestedStart));
      }
      path.prevRequestedStart = path.requestedStart;
    }
    std::vector<size_t> ordering(path.requestedEnds.size());
    std::iota(ordering.begin(), ordering.end(), 0);
    std::sort(ordering.begin(), ordering.end(), [&path](const size_t a, const size_t b) -> bool {
      return path.minTheoreticalDist[a] < path.minTheoreticalDist[b];
    });
    for (size_t i : ordering) {
      if (!path.endIsValid[i]) continue;
      if (path.minTheoreticalDist[i] > path.geodesicDistance) continue;
      float len;
      std::vector<V3> pts;
      const bool found =
          findPathInternal(q, path.requestedStart, startRef, pathStart, path.requestedEnds[i],
                           path.endRefs[i], path.pathEnds[i], len, pts);
      if (found && len < path.geodesicDistance) {
        path.minTheoreticalDist[i] = len;
        path.geodesicDistance = len;
        path.points = pts;
        path.closestEndPointIndex = i;
      }
    }
    return path.geodesicDistance < kInf;
  }

  // PF.cpp:1575-1722
  V3 tryStep(dtNavMeshQuery* q, const V3& start, const V3& end, bool allowSliding) const {
    static const int MAX_POLYS = 256;
    dtPolyRef polys[MAX_POLYS];
    dtStatus startStatus = 0, endStatus = 0;
    dtPolyRef startRef = 0, endRef = 0;
    V3 pathStart, ignore;
    std::tie(startStatus, startRef, pathStart) = projectToPoly(start, q);
    std::tie(endStatus, endRef, ignore) = projectToPoly(end, q);
    if (dtStatusFailed(startStatus) || dtStatusFailed(endStatus)) return start;
    if (!islands_->hasConnection(startRef, endRef)) return start;
    V3 endPoint;
    int numPolys = 0;
    q->moveAlongSurface(startRef, pathStart.data(), end.data(), &filter_, endPoint.data(),
                        polys, &numPolys, MAX_POLYS);
    if (numPolys == 0) return start;
    if (!allowSliding) {
      float bestDist = std::numeric_limits<float>::max();
      bool hitWall = false;
      V3 bestPos;
      for (int iPoly = 0; iPoly < numPolys; ++iPoly) {
        const dtMeshTile* tile = nullptr;
        const dtPoly* poly = nullptr;
        navMesh_->getTileAndPolyByRefUnsafe(polys[iPoly], &tile, &poly);
        for (int j = 0, nv = poly->vertCount; j < nv; ++j) {
          bool isWall = false;
          if (poly->neis[j] == 0) {
            isWall = true;
          } else if (poly->neis[j] & DT_EXT_LINK) {
            bool hasPassableNeighbor = false;
            for (unsigned int k = poly->firstLink; k != DT_NULL_LINK; k = tile->links[k].next) {
              if (tile->links[k].edge == j && tile->links[k].ref != 0) {
                const dtMeshTile* neiTile = nullptr;
                const dtPoly* neiPoly = nullptr;
                navMesh_->getTileAndPolyByRefUnsafe(tile->links[k].ref, &neiTile, &neiPoly);
                if (filter_.passFilter(tile->links[k].ref, neiTile, neiPoly)) {
                  hasPassableNeighbor = true;
                  break;
                }
              }
            }
            isWall = !hasPassableNeighbor;
          }
          if (!isWall) continue;
          const float* vj = &tile->verts[static_cast<size_t>(poly->verts[j]) * 3];
          const int nextIdx = (j + 1 < nv) ? (j + 1) : 0;
          const float* vi = &tile->verts[static_cast<size_t>(poly->verts[nextIdx]) * 3];
          float s, t;
          if (dtIntersectSegSeg2D(vj, vi, pathStart.data(), end.data(), s, t) && t >= 0.0f &&
              t <= 1.0f && s >= 0.0f && s <= 1.0f) {
            float newPos[3];
            dtVlerp(newPos, vj, vi, s);
            const float distSqr = dtVdist2DSqr(newPos, end.data());
            if (distSqr < bestDist) {
              bestPos = V3(newPos);
              bestDist = distSqr;
              hitWall = true;
            }
          }
        }
      }
      if (hitWall) endPoint = bestPos;
    }
    q->getPolyHeight(polys[numPolys - 1], endPoint.data(), &endPoint[1]);
    std::tie(std::ignore, endRef, std::ignore) = projectToPoly(endPoint, q);
    if (!islands_->hasConnection(startRef, endRef)) {
      const dtMeshTile* tile = nullptr;
      const dtPoly* poly = nullptr;
      navMesh_->getTileAndPolyByRefUnsafe(polys[numPolys - 1], &tile, &poly);
      V3 polyCenter;
      for (int iVert = 0; iVert < poly->vertCount; ++iVert)
        polyCenter = polyCenter + V3(&tile->verts[static_cast<size_t>(poly->verts[iVert]) * 3]);
      const float n = static_cast<float>(poly->vertCount);
      polyCenter = V3(polyCenter.x / n, polyCenter.y / n, polyCenter.z / n);
      constexpr float nudgeDistance = 1e-4;
      const V3 d = polyCenter - endPoint;
      const float inv = 1.0f / mnLength(d);
      const V3 nudgeDir = inv * d;
      endPoint = endPoint + nudgeDistance * nudgeDir;
    }
    return endPoint;
  }

  // PF.cpp:1725-1757
  V3 snapPoint(const V3& pt, int islandIndex, dtPolyRef* outRef = nullptr) {
    if (islandIndex != ID_UNDEFINED) {
      islands_->setPolyFlagForIsland(navMesh_, POLYFLAGS_OFF_ISLAND, islandIndex, true, true);
      filter_.setExcludeFlags(filter_.getExcludeFlags() | POLYFLAGS_OFF_ISLAND);
    }
    dtStatus status = 0;
    V3 projectedPt;
    dtPolyRef ref = 0;
    std::tie(status, ref, projectedPt) = projectToPoly(pt, navQuery_);
    if (islandIndex != ID_UNDEFINED) {
      islands_->setPolyFlagForIsland(navMesh_, POLYFLAGS_OFF_ISLAND, islandIndex, false, true);
      filter_.setExcludeFlags(filter_.getExcludeFlags() & ~POLYFLAGS_OFF_ISLAND);
    }
    if (outRef) *outRef = dtStatusSucceed(status) ? ref : 0;
    if (dtStatusSucceed(status)) return projectedPt;
    return {kNaN, kNaN, kNaN};
  }

  // PF.cpp:1760-1771
  int getIsland(dtNavMeshQuery* q, const V3& pt) const {
    dtStatus status = 0;
    V3 projectedPt;
    dtPolyRef polyRef = 0;
    std::tie(status, polyRef, projectedPt) = projectToPoly(pt, q);
    if (dtStatusSucceed(status)) return islands_->getPolyIsland(polyRef);
    return ID_UNDEFINED;
  }

  // PF.cpp:1777-1786
  float islandRadiusAt(dtNavMeshQuery* q, const V3& pt) const {
    dtPolyRef ptRef = 0;
    dtStatus status = 0;
    V3 ignore;
    std::tie(status, ptRef, ignore) = projectToPoly(pt, q);
    if (status != DT_SUCCESS || ptRef == 0) return 0.0;
    return islands_->polyIslandRadius(ptRef);
  }

  // PF.cpp:1794-1812
  HitRecord closestObstacleSurfacePoint(dtNavMeshQuery* q, const V3& pt,
                                        float maxSearchRadius) const {
    dtPolyRef ptRef = 0;
    dtStatus status = 0;
    V3 polyPt;
    std::tie(status, ptRef, polyPt) = projectToPoly(pt, q);
    if (status != DT_SUCCESS || ptRef == 0) return {V3(0, 0, 0), V3(0, 0, 0), kInf};
    V3 hitPos, hitNormal;
    float hitDist = kNaN;
    q->findDistanceToWall(ptRef, polyPt.data(), maxSearchRadius, &filter_, &hitDist,
                          hitPos.data(), hitNormal.data());
    return {hitPos, hitNormal, hitDist};
  }

  // PF.cpp:1814-1831
  bool isNavigable(dtNavMeshQuery* q, const V3& pt, float maxYDelta) const {
    dtPolyRef ptRef = 0;
    dtStatus status = 0;
    V3 polyPt;
    std::tie(status, ptRef, polyPt) = projectToPoly(pt, q);
    if (status != DT_SUCCESS || ptRef == 0) return false;
    const float dx = pt[0] - polyPt[0], dz = pt[2] - polyPt[2];
    float d2 = 0.f;
    d2 += dx * dx;
    d2 += dz * dz;
    if (std::abs(polyPt[1] - pt[1]) > maxYDelta || std::sqrt(d2) > 1e-2f) return false;
    return true;
  }

  dtNavMesh* nav() { return navMesh_; }
  dtNavMeshQuery* query() { return navQuery_; }
  IslandSystem* islands() { return islands_.get(); }
  const dtQueryFilter* filter() const { return &filter_; }
  const NavMeshSettings& settings() const { return settings_; }
  const V3* bounds() const { return bounds_; }

  // one extra dtNavMeshQuery per worker thread (node pools are per query object;
  // the dtNavMesh is only read by the queries used in threaded batches)
  dtNavMeshQuery* threadQuery(int i) {
    while (static_cast<int>(extraQueries_.size()) <= i) {
      dtNavMeshQuery* q = dtAllocNavMeshQuery();
      q->init(navMesh_, 2048);
      extraQueries_.push_back(q);
    }
    return extraQueries_[i];
  }

 private:
  dtNavMesh* navMesh_ = nullptr;
  dtNavMeshQuery* navQuery_ = nullptr;
  std::vector<dtNavMeshQuery*> extraQueries_;
  dtQueryFilter filter_;
  std::unique_ptr<IslandSystem> islands_;
  NavMeshSettings settings_;
  V3 bounds_[2];

  void reset() {
    for (auto* q : extraQueries_) dtFreeNavMeshQuery(q);
    extraQueries_.clear();
    if (navQuery_) dtFreeNavMeshQuery(navQuery_);
    navQuery_ = nullptr;
    if (navMesh_) dtFreeNavMesh(navMesh_);
    navMesh_ = nullptr;
    islands_.reset();
  }
  void computeBoundsFromTiles() {
    bool first = true;
    for (int i = 0; i < navMesh_->getMaxTiles(); ++i) {
      const dtMeshTile* tile = const_cast<const dtNavMesh*>(navMesh_)->getTile(i);
      if (!tile || !tile->header) continue;
      for (int k = 0; k < 3; ++k) {
        bounds_[0][k] = first ? tile->header->bmin[k] : std::min(bounds_[0][k], tile->header->bmin[k]);
        bounds_[1][k] = first ? tile->header->bmax[k] : std::max(bounds_[1][k], tile->header->bmax[k]);
      }
      first = false;
    }
  }
  // PF.cpp:932-950
  bool initNavQuery() {
    navQuery_ = dtAllocNavMeshQuery();
    dtStatus status = navQuery_->init(navMesh_, 2048);
    if (dtStatusFailed(status)) return false;
    islands_ = std::make_unique<IslandSystem>(navMesh_, &filter_);
    islands_->removeZeroAreaPolys(navMesh_);
    return true;
  }
};

template <class F>
void parallelFor(RefPathFinder* pf, int64_t n, int nthreads, F&& fn) {
  nthreads = std::max(1, nthreads);
  if (nthreads == 1 || n < 2) {
    dtNavMeshQuery* q = pf->threadQuery(0);
    for (int64_t i = 0; i < n; ++i) fn(q, i);
    return;
  }
  for (int t = 0; t < nthreads; ++t) pf->threadQuery(t);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t) {
    pool.emplace_back([=, &fn]() {
      dtNavMeshQuery* q = pf->threadQuery(t);
      const int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
      for (int64_t i = lo; i < hi; ++i) fn(q, i);
    });
  }
  for (auto& th : pool) th.join();
}

inline V3 ld(const float* p, int64_t i) { return V3(p + 3 * i); }
inline void st(float* p, int64_t i, const V3& v) {
  p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z;
}

}  // namespace

// =============================== C ABI =========================================
extern "C" {

typedef void* ref_pf_t;

ref_pf_t ref_create() { return new RefPathFinder(); }
void ref_destroy(ref_pf_t h) { delete static_cast<RefPathFinder*>(h); }

void ref_default_settings(void* out56) {
  NavMeshSettings s;
  memcpy(out56, &s, sizeof(s));
}

int ref_build(ref_pf_t h, const void* settings56, const float* verts, int nverts,
              const int* tris, int ntris) {
  NavMeshSettings s;
  memcpy(&s, settings56, sizeof(s));
  return static_cast<RefPathFinder*>(h)->build(s, verts, nverts, tris, ntris) ? 1 : 0;
}

int ref_build_tiled(ref_pf_t h, const void* settings56, const float* verts, int nverts,
                    const int* tris, int ntris, int tileSize, int nthreads) {
  NavMeshSettings s;
  memcpy(&s, settings56, sizeof(s));
  return static_cast<RefPathFinder*>(h)->buildTiled(s, verts, nverts, tris, ntris, tileSize,
                                                    nthreads) ? 1 : 0;
}

int ref_load_memory(ref_pf_t h, const unsigned char* buf, int64_t len) {
  return static_cast<RefPathFinder*>(h)->loadFromMemory(buf, static_cast<size_t>(len)) ? 1 : 0;
}

int ref_load(ref_pf_t h, const char* path) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return 0;
  fseek(fp, 0, SEEK_END);
  long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  std::vector<unsigned char> buf(n);
  size_t got = fread(buf.data(), 1, n, fp);
  fclose(fp);
  if (got != static_cast<size_t>(n)) return 0;
  return ref_load_memory(h, buf.data(), n);
}

// two-call pattern: returns required size; copies when cap is large enough
int64_t ref_save_memory(ref_pf_t h, unsigned char* out, int64_t cap) {
  std::vector<unsigned char> buf;
  if (!static_cast<RefPathFinder*>(h)->saveToMemory(buf)) return -1;
  if (out && cap >= static_cast<int64_t>(buf.size())) memcpy(out, buf.data(), buf.size());
  return static_cast<int64_t>(buf.size());
}

int ref_save(ref_pf_t h, const char* path) {
  std::vector<unsigned char> buf;
  if (!static_cast<RefPathFinder*>(h)->saveToMemory(buf)) return 0;
  FILE* fp = fopen(path, "wb");
  if (!fp) return 0;
  fwrite(buf.data(), 1, buf.size(), fp);
  fclose(fp);
  return 1;
}

int ref_is_loaded(ref_pf_t h) { return static_cast<RefPathFinder*>(h)->isLoaded(); }
int ref_num_islands(ref_pf_t h) { return static_cast<RefPathFinder*>(h)->islands()->numIslands(); }
float ref_navigable_area(ref_pf_t h, int island) {
  auto* pf = static_cast<RefPathFinder*>(h);
  if (!pf->islands()->validIsland(island)) return kNaN;
  return pf->islands()->getNavigableArea(island);
}
float ref_island_radius(ref_pf_t h, int island) {
  auto* pf = static_cast<RefPathFinder*>(h);
  if (!pf->islands()->validIsland(island, false)) return kNaN;
  return pf->islands()->islandRadius(island);
}
void ref_get_bounds(ref_pf_t h, float* out6) {
  auto* pf = static_cast<RefPathFinder*>(h);
  st(out6, 0, pf->bounds()[0]);
  st(out6, 1, pf->bounds()[1]);
}
void ref_seed(ref_pf_t h, uint32_t s) { static_cast<RefPathFinder*>(h)->seed(s); }

// mesh statistics: {tiles, polys, verts, links(maxLinkCount), bvNodes, detailTris,
// detailVerts, tileDataBytes}
void ref_mesh_stats(ref_pf_t h, int64_t* out8) {
  auto* pf = static_cast<RefPathFinder*>(h);
  const dtNavMesh* nav = pf->nav();
  memset(out8, 0, sizeof(int64_t) * 8);
  for (int i = 0; i < nav->getMaxTiles(); ++i) {
    const dtMeshTile* t = nav->getTile(i);
    if (!t || !t->header) continue;
    out8[0]++;
    out8[1] += t->header->polyCount;
    out8[2] += t->header->vertCount;
    out8[3] += t->header->maxLinkCount;
    out8[4] += t->header->bvNodeCount;
    out8[5] += t->header->detailTriCount;
    out8[6] += t->header->detailVertCount;
    out8[7] += t->dataSize;
  }
}

// Finalised tile blobs (links connected, zero-area polys disabled): what a
// habitat-sim integration would hand to hbn_navmesh_create after initNavQuery.
int ref_tile_count(ref_pf_t h) {
  auto* pf = static_cast<RefPathFinder*>(h);
  const dtNavMesh* nav = pf->nav();
  int n = 0;
  for (int i = 0; i < nav->getMaxTiles(); ++i) {
    const dtMeshTile* t = nav->getTile(i);
    if (t && t->header && t->dataSize) n++;
  }
  return n;
}
// idx counts non-empty tiles in table order; returns dataSize, fills ref/salt/index
int ref_tile_blob(ref_pf_t h, int idx, unsigned char* out, int cap, uint32_t* tileRef,
                  int* tableIndex) {
  auto* pf = static_cast<RefPathFinder*>(h);
  const dtNavMesh* nav = pf->nav();
  int n = 0;
  for (int i = 0; i < nav->getMaxTiles(); ++i) {
    const dtMeshTile* t = nav->getTile(i);
    if (!(t && t->header && t->dataSize)) continue;
    if (n++ != idx) continue;
    if (tileRef) *tileRef = nav->getTileRef(t);
    if (tableIndex) *tableIndex = i;
    if (out && cap >= t->dataSize) memcpy(out, t->data, t->dataSize);
    return t->dataSize;
  }
  return -1;
}
void ref_navmesh_params(ref_pf_t h, float* orig3, float* tileWH2, int* maxTilesPolys2) {
  const dtNavMeshParams* p = static_cast<RefPathFinder*>(h)->nav()->getParams();
  memcpy(orig3, p->orig, 12);
  tileWH2[0] = p->tileWidth;
  tileWH2[1] = p->tileHeight;
  maxTilesPolys2[0] = p->maxTiles;
  maxTilesPolys2[1] = p->maxPolys;
}

// island id of every poly in (tile table order, poly order); -1 if unmapped
int64_t ref_poly_islands(ref_pf_t h, int32_t* out, uint32_t* outRefs, int64_t cap) {
  auto* pf = static_cast<RefPathFinder*>(h);
  const dtNavMesh* nav = pf->nav();
  int64_t n = 0;
  for (int i = 0; i < nav->getMaxTiles(); ++i) {
    const dtMeshTile* t = nav->getTile(i);
    if (!t || !t->header) continue;
    for (int j = 0; j < t->header->polyCount; ++j) {
      dtPolyRef ref = nav->encodePolyId(t->salt, i, j);
      if (n < cap) {
        if (out) out[n] = pf->islands()->findPolyIsland(ref);
        if (outRefs) outRefs[n] = ref;
      }
      n++;
    }
  }
  return n;
}

// ---- batched queries ----------------------------------------------------------
// snap_point / findNearestPoly: out_pts NaN on failure, out_refs 0, out_island -1
void ref_snap_batch(ref_pf_t h, const float* pts, int64_t n, float* out_pts,
                    uint32_t* out_refs, int32_t* out_island, int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    dtStatus status;
    dtPolyRef ref;
    V3 p;
    std::tie(status, ref, p) = pf->projectToPoly(ld(pts, i), q);
    const bool ok = dtStatusSucceed(status);
    if (out_pts) st(out_pts, i, ok ? p : V3(kNaN, kNaN, kNaN));
    if (out_refs) out_refs[i] = ok ? ref : 0;
    if (out_island) out_island[i] = ok ? pf->islands()->getPolyIsland(ref) : -1;
  });
}

// island-restricted snap exactly as PF.cpp:1725-1757 (mutates poly flags; serial)
void ref_snap_island_batch(ref_pf_t h, const float* pts, const int32_t* islands, int64_t n,
                           float* out_pts, uint32_t* out_refs) {
  auto* pf = static_cast<RefPathFinder*>(h);
  for (int64_t i = 0; i < n; ++i) {
    dtPolyRef ref = 0;
    V3 p = pf->snapPoint(ld(pts, i), islands ? islands[i] : -1, &ref);
    st(out_pts, i, p);
    if (out_refs) out_refs[i] = ref;
  }
}

void ref_is_navigable_batch(ref_pf_t h, const float* pts, int64_t n, float maxYDelta,
                            uint8_t* out, int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    out[i] = pf->isNavigable(q, ld(pts, i), maxYDelta) ? 1 : 0;
  });
}

void ref_island_radius_batch(ref_pf_t h, const float* pts, int64_t n, float* out, int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    out[i] = pf->islandRadiusAt(q, ld(pts, i));
  });
}

// find_path(ShortestPath).  out_dist inf / out_npts 0 on failure.  out_pts is
// [n, max_pts, 3] (first min(npts,max_pts) points written) or null.
void ref_find_path_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                         float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                         int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    float dist;
    std::vector<V3> pts;
    pf->findPathSingle(q, ld(starts, i), ld(ends, i), dist, pts);
    out_dist[i] = dist;
    if (out_npts) out_npts[i] = static_cast<int32_t>(pts.size());
    if (out_pts)
      for (int k = 0; k < static_cast<int>(pts.size()) && k < max_pts; ++k)
        st(out_pts, i * max_pts + k, pts[k]);
  });
}

// Raw view of one find_path: refs, corridor, Detour status words, node-pool use.
// out_corridor is [n, 256]; out_info is [n, 8] = {startRef, endRef, astarStatus,
// straightStatus, numPolys, numPoints, nodesUsed, flags(bit0 trivial, bit1 connected,
// bit2 found)}.
void ref_find_path_raw_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                             float* out_dist, uint32_t* out_corridor, uint32_t* out_info,
                             float* out_pts, int max_pts, int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    uint32_t* info = out_info + i * 8;
    memset(info, 0, 32);
    out_dist[i] = kInf;
    dtStatus s0, s1;
    dtPolyRef startRef = 0, endRef = 0;
    V3 pathStart, pathEnd;
    const V3 start = ld(starts, i), end = ld(ends, i);
    std::tie(s0, startRef, pathStart) = pf->projectToPoly(start, q);
    if (s0 != DT_SUCCESS || startRef == 0) return;
    info[0] = startRef;
    std::tie(s1, endRef, pathEnd) = pf->projectToPoly(end, q);
    if (s1 != DT_SUCCESS || endRef == 0) return;
    info[1] = endRef;
    RefPathFinder::RawPath raw;
    float len;
    std::vector<V3> pts;
    const bool found = pf->findPathInternal(q, start, startRef, pathStart, end, endRef, pathEnd,
                                            len, pts, &raw);
    info[2] = raw.astarStatus;
    info[3] = raw.straightStatus;
    info[4] = raw.numPolys;
    info[5] = found ? static_cast<uint32_t>(pts.size()) : raw.numPoints;
    info[6] = raw.nodesUsed;
    info[7] = (raw.trivial ? 1u : 0u) | (raw.connected ? 2u : 0u) | (found ? 4u : 0u);
    if (out_corridor)
      for (int k = 0; k < raw.numPolys && k < 256; ++k) out_corridor[i * 256 + k] = raw.polys[k];
    if (found) {
      out_dist[i] = len;
      if (out_pts)
        for (int k = 0; k < static_cast<int>(pts.size()) && k < max_pts; ++k)
          st(out_pts, i * max_pts + k, pts[k]);
    }
  });
}

// Work counters of one find_path for the roofline's ALGORITHMIC bytes:
// out_stats is [n, 6] = {expanded polys (CLOSED nodes left in the pool after findPath),
// links of the expanded polys, their non-null neighbours, corridor polys, links of the
// corridor polys, straight-path points}.  From these
//   B_astar  = 32*expanded + 12*links + (32+24)*neighbours      (dtPoly 32 B, dtLink 12 B,
//   B_funnel = (32+24)*corridor + 12*corridorLinks                two portal verts 24 B)
// Re-expansions of re-opened nodes are not counted (a lower bound).
void ref_find_path_stats_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                               uint32_t* out_stats, int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  const dtNavMesh* nav = pf->nav();
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    uint32_t* st6 = out_stats + i * 8;  // [6] nodes allocated, [7] nodes still open at the end
    memset(st6, 0, 32);
    dtStatus s0, s1;
    dtPolyRef startRef = 0, endRef = 0;
    V3 pathStart, pathEnd;
    const V3 start = ld(starts, i), end = ld(ends, i);
    std::tie(s0, startRef, pathStart) = pf->projectToPoly(start, q);
    if (s0 != DT_SUCCESS || startRef == 0) return;
    std::tie(s1, endRef, pathEnd) = pf->projectToPoly(end, q);
    if (s1 != DT_SUCCESS || endRef == 0) return;
    RefPathFinder::RawPath raw;
    float len;
    std::vector<V3> pts;
    const bool found = pf->findPathInternal(q, start, startRef, pathStart, end, endRef, pathEnd,
                                            len, pts, &raw);
    auto linkStats = [&](dtPolyRef ref, uint32_t& links, uint32_t& neis) {
      const dtMeshTile* tile = nullptr;
      const dtPoly* poly = nullptr;
      if (dtStatusFailed(nav->getTileAndPolyByRef(ref, &tile, &poly))) return;
      for (unsigned int k = poly->firstLink; k != DT_NULL_LINK; k = tile->links[k].next) {
        links++;
        if (tile->links[k].ref) neis++;
      }
    };
    if (raw.connected && startRef != endRef) {
      dtNodePool* pool = q->getNodePool();
      const int cnt = pool->getNodeCount();
      st6[6] = static_cast<uint32_t>(cnt);
      for (int k = 0; k < cnt; ++k) {
        const dtNode* node = pool->getNodeAtIdx(k + 1);
        if (node && (node->flags & DT_NODE_OPEN)) st6[7]++;
        if (!node || !(node->flags & DT_NODE_CLOSED)) continue;
        st6[0]++;
        linkStats(node->id, st6[1], st6[2]);
      }
    }
    st6[3] = raw.numPolys;
    uint32_t dummy = 0;
    for (int k = 0; k < raw.numPolys; ++k) linkStats(raw.polys[k], st6[4], dummy);
    st6[5] = found ? static_cast<uint32_t>(pts.size()) : 0;
  });
}

// find_path(MultiGoalShortestPath), fresh object per start (no cache).
// ends is [n, g, 3].  Outputs: dist[n], idx[n], npts[n], pts [n,max_pts,3]|null.
void ref_find_path_multigoal_batch(ref_pf_t h, const float* starts, const float* ends,
                                   int64_t n, int g, float* out_dist, int32_t* out_idx,
                                   int32_t* out_npts, float* out_pts, int max_pts,
                                   int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    RefPathFinder::MultiGoalPath path;
    path.requestedStart = ld(starts, i);
    std::vector<V3> e(g);
    for (int k = 0; k < g; ++k) e[k] = ld(ends, i * g + k);
    path.setRequestedEnds(e);
    pf->findPathMulti(q, path);
    out_dist[i] = path.geodesicDistance;
    out_idx[i] = path.closestEndPointIndex;
    if (out_npts) out_npts[i] = static_cast<int32_t>(path.points.size());
    if (out_pts)
      for (int k = 0; k < static_cast<int>(path.points.size()) && k < max_pts; ++k)
        st(out_pts, i * max_pts + k, path.points[k]);
  });
}

// Stateful multi-goal object, for the cached-vs-uncached property
// (src/tests/PathFinderTest.cpp:136-164) and trap T4.
void* ref_multigoal_create() { return new RefPathFinder::MultiGoalPath(); }
void ref_multigoal_destroy(void* m) { delete static_cast<RefPathFinder::MultiGoalPath*>(m); }
void ref_multigoal_set_ends(void* m, const float* ends, int g) {
  std::vector<V3> e(g);
  for (int k = 0; k < g; ++k) e[k] = ld(ends, k);
  static_cast<RefPathFinder::MultiGoalPath*>(m)->setRequestedEnds(e);
}
int ref_multigoal_find(ref_pf_t h, void* m, const float* start, float* out_dist,
                       int32_t* out_idx, int32_t* out_npts, float* out_pts, int max_pts) {
  auto* pf = static_cast<RefPathFinder*>(h);
  auto* path = static_cast<RefPathFinder::MultiGoalPath*>(m);
  path->requestedStart = V3(start);
  const bool ok = pf->findPathMulti(pf->query(), *path);
  *out_dist = path->geodesicDistance;
  *out_idx = path->closestEndPointIndex;
  *out_npts = static_cast<int32_t>(path->points.size());
  if (out_pts)
    for (int k = 0; k < static_cast<int>(path->points.size()) && k < max_pts; ++k)
      st(out_pts, k, path->points[k]);
  return ok ? 1 : 0;
}

void ref_try_step_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                        int allowSliding, float* out, int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    st(out, i, pf->tryStep(q, ld(starts, i), ld(ends, i), allowSliding != 0));
  });
}

// closest_obstacle_surface_point: out is [n, 7] = hitPos, hitNormal, hitDist
void ref_obstacle_batch(ref_pf_t h, const float* pts, int64_t n, float maxRadius, float* out,
                        int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    HitRecord r = pf->closestObstacleSurfacePoint(q, ld(pts, i), maxRadius);
    float* o = out + 7 * i;
    o[0] = r.hitPos.x; o[1] = r.hitPos.y; o[2] = r.hitPos.z;
    o[3] = r.hitNormal.x; o[4] = r.hitNormal.y; o[5] = r.hitNormal.z;
    o[6] = r.hitDist;
  });
}

// get_random_navigable_point.  mode 0: glibc rand() stream (reference behaviour,
// sequential).  mode 1: counter-based stream hbn_uniform(seed, query0+i, draw).
// islands may be null (-1 for all).  Returns 0 if the reference would throw.
int ref_random_points(ref_pf_t h, int64_t n, int maxTries, const int32_t* islands, int mode,
                      uint64_t seed, uint64_t query0, float* out_pts, uint32_t* out_refs) {
  auto* pf = static_cast<RefPathFinder*>(h);
  for (int64_t i = 0; i < n; ++i) {
    tlsRand.mode = mode;
    tlsRand.seed = seed;
    tlsRand.query = query0 + static_cast<uint64_t>(i);
    tlsRand.draw = 0;
    V3 p;
    dtPolyRef ref = 0;
    if (!pf->getRandomNavigablePoint(maxTries, islands ? islands[i] : -1, p, &ref)) {
      tlsRand.mode = 0;
      return 0;
    }
    st(out_pts, i, p);
    if (out_refs) out_refs[i] = ref;
  }
  tlsRand.mode = 0;
  return 1;
}

int ref_random_points_near(ref_pf_t h, int64_t n, const float* centers, float radius,
                           int maxTries, const int32_t* islands, int mode, uint64_t seed,
                           uint64_t query0, float* out_pts) {
  auto* pf = static_cast<RefPathFinder*>(h);
  for (int64_t i = 0; i < n; ++i) {
    tlsRand.mode = mode;
    tlsRand.seed = seed;
    tlsRand.query = query0 + static_cast<uint64_t>(i);
    tlsRand.draw = 0;
    V3 p;
    if (!pf->getRandomNavigablePointInCircle(ld(centers, i), radius, maxTries,
                                             islands ? islands[i] : -1, p)) {
      tlsRand.mode = 0;
      return 0;
    }
    st(out_pts, i, p);
  }
  tlsRand.mode = 0;
  return 1;
}

// The goal ordering of PF.cpp:1542-1548 on its own: std::sort of 0..n-1 by key (unstable).
void ref_std_sort_order(const float* key, int n, int32_t* order) {
  std::vector<size_t> ordering(n);
  std::iota(ordering.begin(), ordering.end(), 0);
  std::sort(ordering.begin(), ordering.end(),
            [key](const size_t a, const size_t b) -> bool { return key[a] < key[b]; });
  for (int i = 0; i < n; ++i) order[i] = static_cast<int32_t>(ordering[i]);
}

float ref_uniform(uint64_t seed, uint64_t query, uint32_t draw) {
  return hbnUniform(seed, query, draw);
}

// Tests/Detour/Tests_Detour.cpp:5-33 known answers go through this
void ref_random_point_in_convex_poly(const float* pts, int npts, float s, float t, float* out) {
  float areas[DT_VERTS_PER_POLYGON * 4];
  dtRandomPointInConvexPoly(pts, npts, areas, s, t, out);
}

// moveAlongSurface raw: visited corridor for try_step parity ([n,16] refs, count)
void ref_move_along_surface_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                                  float* out_pos, uint32_t* out_visited, int32_t* out_nvisited,
                                  int nthreads) {
  auto* pf = static_cast<RefPathFinder*>(h);
  parallelFor(pf, n, nthreads, [&](dtNavMeshQuery* q, int64_t i) {
    dtStatus s0;
    dtPolyRef startRef;
    V3 pathStart;
    std::tie(s0, startRef, pathStart) = pf->projectToPoly(ld(starts, i), q);
    out_nvisited[i] = 0;
    st(out_pos, i, V3(kNaN, kNaN, kNaN));
    if (dtStatusFailed(s0) || !startRef) return;
    dtPolyRef polys[256];
    int np = 0;
    V3 res;
    const V3 e = ld(ends, i);
    q->moveAlongSurface(startRef, pathStart.data(), e.data(), pf->filter(), res.data(), polys,
                        &np, 256);
    st(out_pos, i, res);
    out_nvisited[i] = np;
    for (int k = 0; k < np && k < 16; ++k) out_visited[i * 16 + k] = polys[k];
  });
}

}  // extern "C"

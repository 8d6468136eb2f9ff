// =============================================================================
// oracle/tiled_recast.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Shared by the two oracle libraries (ref_esp.cpp: the reference's own PathFinder.cpp
// compiled in place; ref_pathfinder.cpp: the round-1 restatement kept as a cross-check):
//   * the 56-byte NavMeshSettings image that .navmesh (MSET v2) files carry,
//   * one Recast tile / solo build with the reference's settings (PF.cpp:612-896),
//   * the TILED host builder (habitat-sim itself only builds one tile) that
//     produces the multi-tile workloads C4/C5 as an MSET image the reference loads,
//   * the counter-based uniform stream hbn_uniform (include/hbn.h) in the reference's own
//     frand() form: float(rand31) / float(RAND_MAX) (PF.cpp:1231-1234).
// =============================================================================
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "DetourCommon.h"
#include "DetourNavMesh.h"
#include "DetourNavMeshBuilder.h"
#include "Recast.h"

namespace hbnoracle {

// PF.h:137-299 (56 bytes, written raw into .navmesh files, PF.cpp:1204)
struct NavMeshSettings {
  float cellSize = 0.05f, cellHeight = 0.2f, agentHeight = 1.5f, agentRadius = 0.1f,
        agentMaxClimb = 0.2f, agentMaxSlope = 45.0f, regionMinSize = 20.f,
        regionMergeSize = 20.f, edgeMaxLen = 12.0f, edgeMaxError = 1.3f,
        vertsPerPoly = 6.0f, detailSampleDist = 6.0f, detailSampleMaxError = 1.0f;
  bool filterLowHangingObstacles = true, filterLedgeSpans = true,
       filterWalkableLowHeightSpans = true, includeStaticObjects = false;
};
static_assert(sizeof(NavMeshSettings) == 56, "NavMeshSettings layout");

// PF.cpp:590-600
enum PolyAreas { POLYAREA_GROUND, POLYAREA_DOOR };
enum PolyFlags {
  POLYFLAGS_WALK = 0x01,
  POLYFLAGS_DOOR = 0x02,
  POLYFLAGS_DISABLED = 0x04,
  POLYFLAGS_OFF_ISLAND = 0x08,
  POLYFLAGS_ALL = 0xffff
};

// PF.cpp:978-991
const int NAVMESHSET_MAGIC = 'M' << 24 | 'S' << 16 | 'E' << 8 | 'T';
const int NAVMESHSET_VERSION = 2;
struct NavMeshSetHeader {
  int magic;
  int version;
  int numTiles;
  dtNavMeshParams params;
};
struct NavMeshTileHeader {
  dtTileRef tileRef;
  int dataSize;
};


// ---- counter-based uniform stream ------------------------------------------
// PF.cpp:1231-1234: frand() = float(rand()) / float(RAND_MAX), rand() in [0, 2^31-1].
// hbn_uniform(seed, query, draw) is defined in exactly that form over a 31-bit hash, so
// that an interposed rand() returning hbnRand31() drives the UNMODIFIED reference code
// with bit-identical uniforms (1.0 is reachable: float(2^31-1) rounds to 2^31; trap T7).
inline uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
inline uint32_t hbnRand31(uint64_t seed, uint64_t query, uint32_t draw) {
  uint32_t h = mix32(static_cast<uint32_t>(seed) ^ 0x9e3779b9U);
  h = mix32(h ^ static_cast<uint32_t>(seed >> 32));
  h = mix32(h ^ static_cast<uint32_t>(query));
  h = mix32(h ^ static_cast<uint32_t>(query >> 32) ^ 0x85ebca6bU);
  h = mix32(h ^ draw);
  return h >> 1;
}
inline float hbnUniform(uint64_t seed, uint64_t query, uint32_t draw) {
  return static_cast<float>(static_cast<int>(hbnRand31(seed, query, draw))) / 2147483648.0f;
}

// ---- one Recast tile / solo build: PF.cpp:612-896 ---------------------------
struct BuildOut {
  unsigned char* navData = nullptr;
  int navDataSize = 0;
  int npolys = 0;
};

// Steps 1-8 of PathFinder::Impl::build.  tiled=false follows PF.cpp exactly;
// tiled=true adds what RecastDemo/Source/Sample_TileMesh.cpp:794-1160 adds for a
// tile (tileSize/borderSize config, expanded bounds, borderSize passed to
// rcBuildRegions, tileX/tileY in the create params).
bool recastBuildOne(const NavMeshSettings& bs, const float* verts, int nverts,
                    const int* tris, int ntris, const float* bmin, const float* bmax,
                    bool tiled, int tileSize, int tx, int ty, BuildOut& out) {
  rcContext ctx(false);
  rcConfig cfg{};
  memset(&cfg, 0, sizeof(cfg));
  cfg.cs = bs.cellSize;
  cfg.ch = bs.cellHeight;
  cfg.walkableSlopeAngle = bs.agentMaxSlope;
  cfg.walkableHeight = static_cast<int>(ceilf(bs.agentHeight / cfg.ch));
  cfg.walkableClimb = static_cast<int>(floorf(bs.agentMaxClimb / cfg.ch));
  cfg.walkableRadius = static_cast<int>(ceilf(bs.agentRadius / cfg.cs));
  cfg.maxEdgeLen = static_cast<int>(bs.edgeMaxLen / bs.cellSize);
  cfg.maxSimplificationError = bs.edgeMaxError;
  cfg.minRegionArea = static_cast<int>(rcSqr(bs.regionMinSize));
  cfg.mergeRegionArea = static_cast<int>(rcSqr(bs.regionMergeSize));
  cfg.maxVertsPerPoly = static_cast<int>(bs.vertsPerPoly);
  cfg.detailSampleDist = bs.detailSampleDist < 0.9f ? 0 : bs.cellSize * bs.detailSampleDist;
  cfg.detailSampleMaxError = bs.cellHeight * bs.detailSampleMaxError;
  rcVcopy(cfg.bmin, bmin);
  rcVcopy(cfg.bmax, bmax);
  if (tiled) {
    cfg.tileSize = tileSize;
    cfg.borderSize = cfg.walkableRadius + 3;
    cfg.width = cfg.tileSize + cfg.borderSize * 2;
    cfg.height = cfg.tileSize + cfg.borderSize * 2;
    cfg.bmin[0] -= cfg.borderSize * cfg.cs;
    cfg.bmin[2] -= cfg.borderSize * cfg.cs;
    cfg.bmax[0] += cfg.borderSize * cfg.cs;
    cfg.bmax[2] += cfg.borderSize * cfg.cs;
  } else {
    rcCalcGridSize(cfg.bmin, cfg.bmax, cfg.cs, &cfg.width, &cfg.height);
  }

  struct Workspace {
    rcHeightfield* solid = nullptr;
    unsigned char* triareas = nullptr;
    rcCompactHeightfield* chf = nullptr;
    rcContourSet* cset = nullptr;
    rcPolyMesh* pmesh = nullptr;
    rcPolyMeshDetail* dmesh = nullptr;
    ~Workspace() {
      rcFreeHeightField(solid);
      delete[] triareas;
      rcFreeCompactHeightfield(chf);
      rcFreeContourSet(cset);
      rcFreePolyMesh(pmesh);
      rcFreePolyMeshDetail(dmesh);
    }
  } ws;

  ws.solid = rcAllocHeightfield();
  if (!rcCreateHeightfield(&ctx, *ws.solid, cfg.width, cfg.height, cfg.bmin, cfg.bmax,
                           cfg.cs, cfg.ch))
    return false;
  ws.triareas = new unsigned char[ntris > 0 ? ntris : 1];
  memset(ws.triareas, 0, ntris * sizeof(unsigned char));
  rcMarkWalkableTriangles(&ctx, cfg.walkableSlopeAngle, verts, nverts, tris, ntris,
                          ws.triareas);
  if (!rcRasterizeTriangles(&ctx, verts, nverts, tris, ws.triareas, ntris, *ws.solid,
                            cfg.walkableClimb))
    return false;
  if (bs.filterLowHangingObstacles)
    rcFilterLowHangingWalkableObstacles(&ctx, cfg.walkableClimb, *ws.solid);
  if (bs.filterLedgeSpans)
    rcFilterLedgeSpans(&ctx, cfg.walkableHeight, cfg.walkableClimb, *ws.solid);
  if (bs.filterWalkableLowHeightSpans)
    rcFilterWalkableLowHeightSpans(&ctx, cfg.walkableHeight, *ws.solid);
  ws.chf = rcAllocCompactHeightfield();
  if (!rcBuildCompactHeightfield(&ctx, cfg.walkableHeight, cfg.walkableClimb, *ws.solid,
                                 *ws.chf))
    return false;
  if (!rcErodeWalkableArea(&ctx, cfg.walkableRadius, *ws.chf)) return false;
  if (!rcBuildDistanceField(&ctx, *ws.chf)) return false;
  if (!rcBuildRegions(&ctx, *ws.chf, tiled ? cfg.borderSize : 0, cfg.minRegionArea,
                      cfg.mergeRegionArea))
    return false;
  ws.cset = rcAllocContourSet();
  if (!rcBuildContours(&ctx, *ws.chf, cfg.maxSimplificationError, cfg.maxEdgeLen, *ws.cset))
    return false;
  ws.pmesh = rcAllocPolyMesh();
  if (!rcBuildPolyMesh(&ctx, *ws.cset, cfg.maxVertsPerPoly, *ws.pmesh)) return false;
  ws.dmesh = rcAllocPolyMeshDetail();
  if (!rcBuildPolyMeshDetail(&ctx, *ws.pmesh, *ws.chf, cfg.detailSampleDist,
                             cfg.detailSampleMaxError, *ws.dmesh))
    return false;
  if (cfg.maxVertsPerPoly > DT_VERTS_PER_POLYGON) return false;
  out.npolys = ws.pmesh->npolys;
  if (tiled && ws.pmesh->npolys == 0) return true;  // empty tile: no data

  for (int i = 0; i < ws.pmesh->npolys; ++i) {
    if (ws.pmesh->areas[i] == RC_WALKABLE_AREA) ws.pmesh->areas[i] = POLYAREA_GROUND;
    if (ws.pmesh->areas[i] == POLYAREA_GROUND) {
      ws.pmesh->flags[i] = POLYFLAGS_WALK;
    } else if (ws.pmesh->areas[i] == POLYAREA_DOOR) {
      ws.pmesh->flags[i] = POLYFLAGS_WALK | POLYFLAGS_DOOR;
    }
  }
  dtNavMeshCreateParams params{};
  memset(&params, 0, sizeof(params));
  params.verts = ws.pmesh->verts;
  params.vertCount = ws.pmesh->nverts;
  params.polys = ws.pmesh->polys;
  params.polyAreas = ws.pmesh->areas;
  params.polyFlags = ws.pmesh->flags;
  params.polyCount = ws.pmesh->npolys;
  params.nvp = ws.pmesh->nvp;
  params.detailMeshes = ws.dmesh->meshes;
  params.detailVerts = ws.dmesh->verts;
  params.detailVertsCount = ws.dmesh->nverts;
  params.detailTris = ws.dmesh->tris;
  params.detailTriCount = ws.dmesh->ntris;
  params.walkableHeight = bs.agentHeight;
  params.walkableRadius = bs.agentRadius;
  params.walkableClimb = bs.agentMaxClimb;
  rcVcopy(params.bmin, ws.pmesh->bmin);
  rcVcopy(params.bmax, ws.pmesh->bmax);
  params.cs = cfg.cs;
  params.ch = cfg.ch;
  params.buildBvTree = true;
  if (tiled) {
    params.tileX = tx;
    params.tileY = ty;
    params.tileLayer = 0;
  }
  if (!dtCreateNavMeshData(&params, &out.navData, &out.navDataSize)) return false;
  return true;
}

// ---- tiled host build -> MSET v2 image ----------------------------------------
// Not a PathFinder method (habitat-sim builds a single tile, PF.cpp:612-930): the C4/C5
// workloads need a multi-tile navmesh, so every tile goes through the same Recast steps
// (recastBuildOne, tiled = true) and the result is written in the file format
// PathFinder::loadNavMesh reads (PF.cpp:1091-1175).  The tile refs are the ones
// dtNavMesh::addTile assigns (a scratch dtNavMesh does the numbering).
// Trap T10: params.maxTiles must equal the number of tiles present.
inline bool buildTiledImage(const NavMeshSettings& bs, const float* verts, int nverts,
                            const int* tris, int ntris, int tileSize, int nthreads,
                            std::vector<unsigned char>& image) {
  float bmin[3], bmax[3];
  rcCalcBounds(verts, nverts, bmin, bmax);
  int gw = 0, gh = 0;
  rcCalcGridSize(bmin, bmax, bs.cellSize, &gw, &gh);
  const int tw = (gw + tileSize - 1) / tileSize;
  const int th = (gh + tileSize - 1) / tileSize;
  const float tcs = tileSize * bs.cellSize;
  const int walkableRadius = static_cast<int>(ceilf(bs.agentRadius / bs.cellSize));
  const float border = (walkableRadius + 3) * bs.cellSize;

  // triangle xz bounds for the per-tile geometry query
  std::vector<float> tb(static_cast<size_t>(ntris) * 4);
  for (int i = 0; i < ntris; ++i) {
    float x0 = FLT_MAX, x1 = -FLT_MAX, z0 = FLT_MAX, z1 = -FLT_MAX;
    for (int k = 0; k < 3; ++k) {
      const float* v = &verts[static_cast<size_t>(tris[i * 3 + k]) * 3];
      x0 = std::min(x0, v[0]); x1 = std::max(x1, v[0]);
      z0 = std::min(z0, v[2]); z1 = std::max(z1, v[2]);
    }
    tb[i * 4 + 0] = x0; tb[i * 4 + 1] = x1; tb[i * 4 + 2] = z0; tb[i * 4 + 3] = z1;
  }
  std::vector<BuildOut> outs(static_cast<size_t>(tw) * th);
  std::atomic<int> next(0);
  std::atomic<bool> ok(true);
  auto worker = [&]() {
    std::vector<int> ltris;
    for (;;) {
      const int t = next.fetch_add(1);
      if (t >= tw * th) break;
      const int tx = t % tw, ty = t / tw;
      float tbmin[3] = {bmin[0] + tx * tcs, bmin[1], bmin[2] + ty * tcs};
      float tbmax[3] = {bmin[0] + (tx + 1) * tcs, bmax[1], bmin[2] + (ty + 1) * tcs};
      const float qx0 = tbmin[0] - border, qx1 = tbmax[0] + border;
      const float qz0 = tbmin[2] - border, qz1 = tbmax[2] + border;
      ltris.clear();
      for (int i = 0; i < ntris; ++i) {
        if (tb[i * 4 + 0] > qx1 || tb[i * 4 + 1] < qx0 || tb[i * 4 + 2] > qz1 ||
            tb[i * 4 + 3] < qz0)
          continue;
        ltris.push_back(tris[i * 3]);
        ltris.push_back(tris[i * 3 + 1]);
        ltris.push_back(tris[i * 3 + 2]);
      }
      if (ltris.empty()) continue;
      if (!recastBuildOne(bs, verts, nverts, ltris.data(), static_cast<int>(ltris.size() / 3),
                          tbmin, tbmax, true, tileSize, tx, ty, outs[t]))
        ok = false;
    }
  };
  std::vector<std::thread> pool;
  for (int i = 0; i < std::max(1, nthreads); ++i) pool.emplace_back(worker);
  for (auto& th_ : pool) th_.join();
  auto freeAll = [&]() {
    for (auto& o : outs)
      if (o.navData) dtFree(o.navData);
  };
  int numTiles = 0, maxPolysInTile = 0;
  for (auto& o : outs)
    if (o.navData) {
      ++numTiles;
      maxPolysInTile = std::max(maxPolysInTile, o.npolys);
    }
  if (!ok || numTiles == 0) {
    freeAll();
    return false;
  }
  NavMeshSetHeader header{};
  header.magic = NAVMESHSET_MAGIC;
  header.version = NAVMESHSET_VERSION;
  header.numTiles = numTiles;
  rcVcopy(header.params.orig, bmin);
  header.params.tileWidth = tcs;
  header.params.tileHeight = tcs;
  header.params.maxTiles = numTiles;
  const int tileBits = dtIlog2(dtNextPow2(static_cast<unsigned int>(numTiles)));
  const int polyBits = std::min(22 - tileBits, 16);
  header.params.maxPolys = 1 << polyBits;
  dtNavMesh* scratch = dtAllocNavMesh();
  if (maxPolysInTile > header.params.maxPolys || !scratch ||
      dtStatusFailed(scratch->init(&header.params))) {
    if (scratch) dtFreeNavMesh(scratch);
    freeAll();
    return false;
  }
  image.clear();
  auto wr = [&](const void* p, size_t n) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    image.insert(image.end(), c, c + n);
  };
  wr(&header, sizeof(header));
  wr(&bs, sizeof(bs));
  bool good = true;
  for (auto& o : outs) {
    if (!o.navData) continue;
    // the image keeps the blob as dtCreateNavMeshData made it; a copy is numbered by Detour
    std::vector<unsigned char> blob(o.navData, o.navData + o.navDataSize);
    dtTileRef ref = 0;
    if (dtStatusFailed(scratch->addTile(o.navData, o.navDataSize, 0, 0, &ref))) {
      good = false;
      break;
    }
    NavMeshTileHeader tileHeader{};
    tileHeader.tileRef = ref;
    tileHeader.dataSize = o.navDataSize;
    wr(&tileHeader, sizeof(tileHeader));
    wr(blob.data(), blob.size());
  }
  dtFreeNavMesh(scratch);
  freeAll();
  return good;
}

}  // namespace hbnoracle

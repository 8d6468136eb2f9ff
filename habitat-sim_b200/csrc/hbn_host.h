// Host side of the navmesh boundary: ingest what the reference's host path produces
// (an MSET v2 .navmesh image, PathFinder.cpp:978-991,1091-1223, or the finalised tile
// blobs of a live dtNavMesh after PathFinder::Impl::initNavQuery, PathFinder.cpp:932-950),
// finish it the way dtNavMesh::addTile + IslandSystem + removeZeroAreaPolys would, and
// flatten it into the structure-of-arrays records of hbn_types.h for upload.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "hbn_types.h"

namespace hbn {

// --- on-disk / in-blob Detour layouts (format definitions, DetourNavMesh.h:155-312) ---
struct DtMeshHeader {
  int32_t magic, version, x, y, layer;
  uint32_t userId;
  int32_t polyCount, vertCount, maxLinkCount, detailMeshCount, detailVertCount, detailTriCount,
      bvNodeCount, offMeshConCount, offMeshBase;
  float walkableHeight, walkableRadius, walkableClimb;
  float bmin[3], bmax[3];
  float bvQuantFactor;
};
static_assert(sizeof(DtMeshHeader) == 100, "dtMeshHeader");
struct DtPoly {
  uint32_t firstLink;
  uint16_t verts[6];
  uint16_t neis[6];
  uint16_t flags;
  uint8_t vertCount;
  uint8_t areaAndtype;
};
static_assert(sizeof(DtPoly) == 32, "dtPoly");
struct DtLink {
  uint32_t ref;
  uint32_t next;
  uint8_t edge, side, bmin, bmax;
};
static_assert(sizeof(DtLink) == 12, "dtLink");
struct DtPolyDetail {
  uint32_t vertBase, triBase;
  uint8_t vertCount, triCount;
};
static_assert(sizeof(DtPolyDetail) == 12, "dtPolyDetail");
struct DtBVNode {
  uint16_t bmin[3], bmax[3];
  int32_t i;
};
static_assert(sizeof(DtBVNode) == 16, "dtBVNode");
struct DtOffMeshConnection {
  float pos[6];
  float rad;
  uint16_t poly;
  uint8_t flags, side;
  uint32_t userId;
};
static_assert(sizeof(DtOffMeshConnection) == 36, "dtOffMeshConnection");
struct DtNavMeshParams {
  float orig[3];
  float tileWidth, tileHeight;
  int32_t maxTiles, maxPolys;
};
static_assert(sizeof(DtNavMeshParams) == 28, "dtNavMeshParams");

constexpr uint32_t kNullLink = 0xffffffffu;
constexpr int32_t kDtNavMeshMagic = 'D' << 24 | 'N' << 16 | 'A' << 8 | 'V';
constexpr int32_t kDtNavMeshVersion = 7;

struct HostTile {
  bool present = false;
  uint32_t salt = 1;
  std::vector<uint8_t> data;  // owned copy of the tile blob (links are edited in place)
  DtMeshHeader* header = nullptr;
  float* verts = nullptr;
  DtPoly* polys = nullptr;
  DtLink* links = nullptr;
  DtPolyDetail* detailMeshes = nullptr;
  float* detailVerts = nullptr;
  uint8_t* detailTris = nullptr;
  DtBVNode* bvTree = nullptr;
  DtOffMeshConnection* offMeshCons = nullptr;
  uint32_t linksFreeList = kNullLink;
  int next = -1;  // bucket chain in the position lookup (dtMeshTile::next)
};

// Flat arrays ready for cudaMemcpy; also everything the host-side PathFinder properties need.
struct FlatNav {
  std::vector<PolyRec> polys;
  std::vector<LinkRec> links;
  std::vector<PortalRec> portals;
  std::vector<BvRec> bv;
  std::vector<TileRec> tiles;
  std::vector<uint8_t> detTris;
  std::vector<float> detVerts;
  std::vector<float> polyBox;  // 8 floats per poly: min xyz, max xyz (poly + detail vertices), 2 pad
  std::vector<uint32_t> gridStart, tileOrder;
  std::vector<RandEntry> randEntries;
  std::vector<uint32_t> tileIslStart, tileIslWin, tileIslCnt;
  std::vector<int32_t> tileIslId;
  int32_t gridMinX = 0, gridMinY = 0, gridW = 0, gridH = 0;
  DtNavMeshParams params{};
  uint32_t polyBits = 0, tileBits = 0, saltBits = 0;
  int32_t bvXzTight = 0;  // see NavView::bvXzTight
  uint32_t numKeys = 0;  // A* node keys (poly, crossSide), see LinkRec::neiKey
  // IslandSystem products (PathFinder.cpp:167-207,1045-1085)
  std::vector<float> islandRadius, islandArea;
  float totalArea = 0.f;
  float bounds[6] = {0, 0, 0, 0, 0, 0};
  uint8_t settings[56] = {0};  // raw NavMeshSettings block of the MSET image
  bool hasSettings = false;
  NavView view() const;  // host-pointer view (used by the host emulation tests only)
};

class HostNavMesh {
 public:
  // PathFinder::Impl::loadNavMesh (PathFinder.cpp:1091-1175) up to, not including,
  // initNavQuery: dtNavMesh::init(params) + addTile per stored tile (links rebuilt).
  bool loadMSET(const uint8_t* buf, size_t len, std::string& err);
  // Finalised tiles of a live dtNavMesh (links already connected): no link building.
  bool addFinalisedTile(const uint8_t* data, size_t len, uint32_t tileRef, std::string& err);
  bool init(const DtNavMeshParams& p, std::string& err);
  // initNavQuery's host work: IslandSystem ctor, then removeZeroAreaPolys.
  // givenIslands (optional, per poly in tile-table/poly order) overrides the flood fill;
  // givenRadii (optional, one per island id) are IslandSystem::islandRadius_ as the caller's
  // PathFinder computed them (the f32 centroid sum runs in its DFS order, which cannot be
  // replayed from finalised flags: trap T5); without them the radii are recomputed per id.
  void finish(const int32_t* givenIslands, const float* givenRadii = nullptr, int numRadii = 0);
  void flatten(FlatNav& out) const;
  // PathFinder::Impl::saveNavMesh (PathFinder.cpp:1177-1223): the MSET v2 image of the tiles as they are
  // now (links connected, zero-area polys disabled) -- also for a mesh handed over as live tiles.
  // Fails like the reference when no NavMeshSettings block is known (PathFinder.cpp:1199-1203).
  bool saveMSET(std::vector<uint8_t>& out, std::string& err) const;
  void setSettings(const uint8_t* raw56);

  // test hooks
  const std::vector<HostTile>& tiles() const { return tiles_; }
  uint32_t encodePolyId(uint32_t salt, uint32_t it, uint32_t ip) const {
    return (salt << (polyBits_ + tileBits_)) | (it << polyBits_) | ip;
  }

 private:
  DtNavMeshParams params_{};
  std::vector<HostTile> tiles_;
  std::vector<int> posLookup_;
  int tileLutMask_ = 0;
  uint32_t saltBits_ = 0, tileBits_ = 0, polyBits_ = 0;
  std::vector<int> freeList_;  // indices still free (dtNavMesh::m_nextFree order)
  std::vector<std::vector<int32_t>> polyIsland_;  // [tile][poly]
  std::vector<float> islandRadius_, islandArea_;
  float totalArea_ = 0.f;
  float bounds_[6] = {0, 0, 0, 0, 0, 0};
  bool boundsInit_ = false;
  uint8_t settings_[56] = {0};
  bool hasSettings_ = false;

  bool addTile(const uint8_t* data, size_t len, uint32_t lastRef, bool buildLinks,
               std::string& err);
  bool patchPointers(HostTile& t, std::string& err);
  uint32_t polyRefBase(int tileIdx) const {
    return encodePolyId(tiles_[tileIdx].salt, static_cast<uint32_t>(tileIdx), 0);
  }
  uint32_t decodeTile(uint32_t ref) const { return (ref >> polyBits_) & ((1u << tileBits_) - 1); }
  uint32_t decodePoly(uint32_t ref) const { return ref & ((1u << polyBits_) - 1); }
  uint32_t decodeSalt(uint32_t ref) const {
    return (ref >> (polyBits_ + tileBits_)) & ((1u << saltBits_) - 1);
  }
  int tilesAt(int x, int y, int* out, int maxOut) const;
  void connectIntLinks(int ti);
  void connectExtLinks(int ti, int target, int side);
  int findConnectingPolys(const float* va, const float* vb, int target, int side, uint32_t* con,
                          float* conarea, int maxcon) const;
  void floodIslands();
  void zeroAreaAndAreas();
};

}  // namespace hbn

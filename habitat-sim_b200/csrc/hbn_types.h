// Device-resident structure-of-arrays navmesh: record layouts shared by the host
// flattener (hbn_host.cpp) and the query code (hbn_query.h).
//
// The reference keeps a navmesh as per-tile AoS blobs (dtMeshTile / dtPoly / dtLink /
// dtPolyDetail / dtBVNode, Detour/Include/DetourNavMesh.h:155-312) chased through
// tile pointers, per-tile vertex indices and linked lists.  Here every tile is
// flattened once, on the host, into a few global arrays indexed by a *global poly
// index* g, laid out so that one query step costs one aligned vector load:
//   PolyRec  (128 B, one cache line): gathered vertex positions, flags, detail-mesh
//            window, link window, island id, 2D area.
//   LinkRec  (32 B): one per dtLink, in exact linked-list order per poly (CSR), holding
//            the neighbour's g, the portal midpoint A* uses (getEdgeMidPoint,
//            DetourNavMeshQuery.cpp:2355-2366, of the FIRST link to that neighbour, as
//            getPortalPoints :2269-2340 resolves it) and the neighbour's own link window.
//   PortalRec(32 B): parallel to LinkRec, portal left/right points for the funnel.
//   BvRec    (16 B): dtBVNode with the leaf index rewritten to g (+ filter bit).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HBN_HD __host__ __device__ __forceinline__
#define HBN_ALIGN(n) __align__(n)
#else
#define HBN_HD inline
#define HBN_ALIGN(n) __attribute__((aligned(n)))
#endif

namespace hbn {

constexpr int kVertsPerPoly = 6;           // DT_VERTS_PER_POLYGON, DetourNavMesh.h:61
constexpr uint16_t kExtLink = 0x8000;      // DT_EXT_LINK
constexpr uint32_t kNoPoly = 0xffffffffu;  // "link->ref == 0"

// poly flags written by PathFinder.cpp:592-600
constexpr uint16_t kFlagWalk = 0x01;
constexpr uint16_t kFlagDisabled = 0x04;

struct HBN_ALIGN(16) PolyRec {
  float v[18];           //   0: vertex positions, gathered from dtMeshTile::verts
  uint32_t ref;          //  72: dtPolyRef (salt|tile|poly), DetourNavMesh.h:529-562
  uint32_t linkStart;    //  76: first LinkRec of this poly
  uint32_t detTriBase;   //  80: first detail triangle (global index)
  uint32_t detVertBase;  //  84: first detail vertex (global index)
  uint16_t flags;        //  88: dtPoly::flags after removeZeroAreaPolys
  uint8_t nv;            //  90: dtPoly::vertCount
  uint8_t areaType;      //  91: dtPoly::areaAndtype
  uint8_t linkCount;     //  92
  uint8_t detTriCount;   //  93
  uint8_t pad0[2];       //  94
  uint16_t neis[6];      //  96: dtPoly::neis (tile local, DT_EXT_LINK encoded)
  int32_t island;        // 108: IslandSystem id (PathFinder.cpp:167-207)
  uint32_t tile;         // 112: index into TileRec[]
  float area2d;          // 116: sum of dtTriArea2D fan (DetourNavMeshQuery.cpp:270-277)
  uint32_t key0;         // 120: node-key index of (this poly, crossSide 0), see LinkRec::neiKey
  uint32_t pad1;         // 124
};
static_assert(sizeof(PolyRec) == 128, "PolyRec must be one cache line");

// meta bits of LinkRec
constexpr uint32_t kLinkEdgeMask = 0xffu;         // dtLink::edge
constexpr uint32_t kLinkSideShift = 8;            // dtLink::side (8 bits)
constexpr uint32_t kLinkStateShift = 16;          // crossSide = side>>1 (0 if side==0xff), 2 bits
constexpr uint32_t kLinkPassBit = 1u << 18;       // neighbour passes the default filter
constexpr uint32_t kLinkOffmeshBit = 1u << 19;    // neighbour is an off-mesh connection poly
constexpr uint32_t kLinkDupBit = 1u << 20;        // an earlier link of this poly has the same neighbour
constexpr uint32_t kLinkNeiCountShift = 24;       // neighbour's linkCount (8 bits)

struct HBN_ALIGN(16) LinkRec {
  float mid[3];           // portal midpoint (A* node position on first visit)
  uint32_t nei;           // neighbour global poly index, kNoPoly if ref==0
  uint32_t neiLinkStart;  // neighbour's link window start
  uint32_t meta;
  uint32_t neiRef;        // dtLink::ref
  // Index of the A* node key (neighbour poly, crossSide) in the dense enumeration of all keys
  // that can occur (DQ.cpp:1067-1073: crossSide = side >> 1 for tile-border links): a direct-
  // mapped per-query node table replaces dtNodePool's hash (hbn_astar_lane.h).
  uint32_t neiKey;
};
static_assert(sizeof(LinkRec) == 32, "LinkRec");

struct HBN_ALIGN(16) PortalRec {
  float l[3];
  float bminmax;  // unused (keeps 16 B alignment of r)
  float r[3];
  float pad;
};
static_assert(sizeof(PortalRec) == 32, "PortalRec");

constexpr int32_t kBvFailBit = 1 << 30;  // leaf poly fails the default filter
// leaf outside the tree proper (the zeroed spare node at the end of Detour's BV array: it reads as
// a leaf of poly 0 with an empty box at the tile origin): always visited, box unrelated to its poly
constexpr int32_t kBvLooseBit = 1 << 29;
constexpr int32_t kBvIndexMask = (1 << 29) - 1;

struct HBN_ALIGN(16) BvRec {
  uint16_t bmin[3];
  uint16_t bmax[3];
  int32_t i;  // >=0: leaf, global poly index (| kBvFailBit); <0: -escape
};
static_assert(sizeof(BvRec) == 16, "BvRec");

struct HBN_ALIGN(16) TileRec {
  float bmin[3];
  float bvQuantFactor;
  float bmax[3];
  float walkableClimb;
  uint32_t bvStart, bvCount;
  uint32_t polyStart, polyCount;
  int32_t x, y, layer;
  uint32_t refBase;      // getPolyRefBase: salt|tile bits
  uint32_t randStart;    // window of RandEntry[]: ground polys passing the default filter,
  uint32_t randCount;    //   in poly order (findRandomPoint's per-tile scan)
  uint32_t pad[2];
};
static_assert(sizeof(TileRec) == 80, "TileRec");

struct HBN_ALIGN(16) RandEntry {
  uint32_t g;     // global poly index
  float area;     // polyArea of DetourNavMeshQuery.cpp:270-277
  float areaSum;  // running sum up to and including this poly (sequential f32)
  uint32_t pad;
};

// Everything a kernel needs; passed by value.
struct NavView {
  const PolyRec* polys;
  const LinkRec* links;
  const PortalRec* portals;
  const BvRec* bv;
  const TileRec* tiles;
  const unsigned char* detTris;  // 4 bytes per triangle
  const float* detVerts;         // 3 floats per vertex
  // bounds of everything closestPointOnPoly can return for a poly (poly + detail vertices):
  // 8 floats per poly, min xyz | max xyz | 2 pad.  Only used to skip hopeless candidates.
  const float* polyBox;
  // tile grid lookup (dtNavMesh::getTilesAt, DetourNavMesh.cpp:1120-1140): dense grid over
  // [gridMinX, gridMinX+gridW) x [gridMinY, gridMinY+gridH); cell -> window of tileOrder[]
  const uint32_t* gridStart;     // gridW*gridH + 1 entries
  const uint32_t* tileOrder;     // tile indices in the reference's bucket-chain order
  // findRandomPoint tables (DetourNavMeshQuery.cpp:226-315).  The reference scans every
  // poly of the chosen tile accumulating areaSum; those running sums only depend on the
  // filter, so they are tabulated once per (tile, all) and per (tile, island).
  const RandEntry* randEntries;
  const uint32_t* tileIslStart;  // numTiles+1: window of the three arrays below
  const int32_t* tileIslId;      // island id present in the tile
  const uint32_t* tileIslWin;    // its RandEntry window start
  const uint32_t* tileIslCnt;    // its RandEntry window length
  int32_t gridMinX, gridMinY, gridW, gridH;
  float orig[3];
  float tileWidth, tileHeight;
  uint32_t numPolys, numTiles, numLinks, numKeys;
  uint32_t polyBits, tileBits, saltBits;
  int32_t numIslands;
  // 1 if every BV leaf box covers its poly's xz bounds (checked by the flattener): the BV walk
  // may then use a narrower xz box than the reference's without losing the winner (hbn_snap.h)
  int32_t bvXzTight;
};

}  // namespace hbn

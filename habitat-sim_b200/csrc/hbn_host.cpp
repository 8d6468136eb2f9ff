// Host ingest + flattening of a Detour navmesh (see hbn_host.h).
//
// What is restated from the reference, and where:
//   * MSET container:            PathFinder.cpp:978-991, 1091-1175
//   * dtNavMesh::init(params):   DetourNavMesh.cpp:220-265 (id bit widths, tile LUT)
//   * dtNavMesh::addTile:        DetourNavMesh.cpp:908-1065 (pointer patching, free list,
//                                connectIntLinks :524-559, connectExtLinks :387-452,
//                                findConnectingPolys :290-346, slab helpers :30-115)
//   * IslandSystem:              PathFinder.cpp:167-207, 409-459
//   * removeZeroAreaPolys/areas: PathFinder.cpp:1001-1085
// Off-mesh connections are never produced by habitat's build (PathFinder.cpp:880-886 are
// commented out); MSET images that carry them are rejected here (they are accepted
// through addFinalisedTile, where the reference has already linked them).
#include "hbn_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <unordered_map>

namespace hbn {
namespace {

inline uint32_t nextPow2(uint32_t v) {
  v--;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  v++;
  return v;
}
inline uint32_t ilog2(uint32_t v) {
  uint32_t r = 0;
  while (v >>= 1) r++;
  return r;
}
inline int align4(int x) { return (x + 3) & ~3; }
inline int oppositeTile(int side) { return (side + 4) & 0x7; }
inline int tileHash(int x, int y, int mask) {
  const uint32_t h1 = 0x8da6b343u, h2 = 0xd8163841u;
  const uint32_t n = h1 * static_cast<uint32_t>(x) + h2 * static_cast<uint32_t>(y);
  return static_cast<int>(n & static_cast<uint32_t>(mask));
}
inline float fmaxf2(float a, float b) { return a > b ? a : b; }  // dtMax
inline float fminf2(float a, float b) { return a < b ? a : b; }  // dtMin
inline float clampf(float v, float mn, float mx) { return v < mn ? mn : (v > mx ? mx : v); }

// DetourNavMesh.cpp:66-73
float slabCoord(const float* va, int side) {
  if (side == 0 || side == 4) return va[0];
  if (side == 2 || side == 6) return va[2];
  return 0;
}
// DetourNavMesh.cpp:75-115
void slabEndPoints(const float* va, const float* vb, float* bmin, float* bmax, int side) {
  if (side == 0 || side == 4) {
    const bool f = va[2] < vb[2];
    const float* lo = f ? va : vb;
    const float* hi = f ? vb : va;
    bmin[0] = lo[2]; bmin[1] = lo[1];
    bmax[0] = hi[2]; bmax[1] = hi[1];
  } else if (side == 2 || side == 6) {
    const bool f = va[0] < vb[0];
    const float* lo = f ? va : vb;
    const float* hi = f ? vb : va;
    bmin[0] = lo[0]; bmin[1] = lo[1];
    bmax[0] = hi[0]; bmax[1] = hi[1];
  }
}
// DetourNavMesh.cpp:30-64
bool overlapSlabs(const float* amin, const float* amax, const float* bmin, const float* bmax,
                  float px, float py) {
  const float minx = fmaxf2(amin[0] + px, bmin[0] + px);
  const float maxx = fminf2(amax[0] - px, bmax[0] - px);
  if (minx > maxx) return false;
  const float ad = (amax[1] - amin[1]) / (amax[0] - amin[0]);
  const float ak = amin[1] - ad * amin[0];
  const float bd = (bmax[1] - bmin[1]) / (bmax[0] - bmin[0]);
  const float bk = bmin[1] - bd * bmin[0];
  const float aminy = ad * minx + ak;
  const float amaxy = ad * maxx + ak;
  const float bminy = bd * minx + bk;
  const float bmaxy = bd * maxx + bk;
  const float dmin = bminy - aminy;
  const float dmax = bmaxy - amaxy;
  if (dmin * dmax < 0) return true;
  const float thr = (py * 2) * (py * 2);
  if (dmin * dmin <= thr || dmax * dmax <= thr) return true;
  return false;
}

inline uint32_t allocLink(HostTile& t) {
  if (t.linksFreeList == kNullLink) return kNullLink;
  const uint32_t l = t.linksFreeList;
  t.linksFreeList = t.links[l].next;
  return l;
}

const int32_t kMsetMagic = 'M' << 24 | 'S' << 16 | 'E' << 8 | 'T';

}  // namespace

bool HostNavMesh::init(const DtNavMeshParams& p, std::string& err) {
  params_ = p;
  if (p.maxTiles <= 0 || p.maxPolys <= 0) {
    err = "invalid dtNavMeshParams";
    return false;
  }
  uint32_t lut = nextPow2(static_cast<uint32_t>(p.maxTiles / 4));
  if (!lut) lut = 1;
  tileLutMask_ = static_cast<int>(lut) - 1;
  tiles_.clear();
  tiles_.resize(p.maxTiles);
  posLookup_.assign(lut, -1);
  freeList_.clear();
  for (int i = 0; i < p.maxTiles; ++i) freeList_.push_back(i);
  tileBits_ = ilog2(nextPow2(static_cast<uint32_t>(p.maxTiles)));
  polyBits_ = ilog2(nextPow2(static_cast<uint32_t>(p.maxPolys)));
  saltBits_ = std::min<uint32_t>(31u, 32u - tileBits_ - polyBits_);
  if (saltBits_ < 10) {
    err = "dtNavMesh::init: fewer than 10 salt bits";
    return false;
  }
  boundsInit_ = false;
  return true;
}

bool HostNavMesh::patchPointers(HostTile& t, std::string& err) {
  if (t.data.size() < sizeof(DtMeshHeader)) {
    err = "tile blob too small";
    return false;
  }
  DtMeshHeader* h = reinterpret_cast<DtMeshHeader*>(t.data.data());
  if (h->magic != kDtNavMeshMagic) { err = "tile: wrong magic"; return false; }
  if (h->version != kDtNavMeshVersion) { err = "tile: wrong version"; return false; }
  const int headerSize = align4(sizeof(DtMeshHeader));
  const int vertsSize = align4(sizeof(float) * 3 * h->vertCount);
  const int polysSize = align4(sizeof(DtPoly) * h->polyCount);
  const int linksSize = align4(sizeof(DtLink) * h->maxLinkCount);
  const int dmSize = align4(sizeof(DtPolyDetail) * h->detailMeshCount);
  const int dvSize = align4(sizeof(float) * 3 * h->detailVertCount);
  const int dtSize = align4(4 * h->detailTriCount);
  const int bvSize = align4(sizeof(DtBVNode) * h->bvNodeCount);
  const int omSize = align4(sizeof(DtOffMeshConnection) * h->offMeshConCount);
  const size_t total = static_cast<size_t>(headerSize) + vertsSize + polysSize + linksSize +
                       dmSize + dvSize + dtSize + bvSize + omSize;
  if (total > t.data.size()) { err = "tile blob truncated"; return false; }
  uint8_t* d = t.data.data() + headerSize;
  t.header = h;
  t.verts = reinterpret_cast<float*>(d); d += vertsSize;
  t.polys = reinterpret_cast<DtPoly*>(d); d += polysSize;
  t.links = reinterpret_cast<DtLink*>(d); d += linksSize;
  t.detailMeshes = reinterpret_cast<DtPolyDetail*>(d); d += dmSize;
  t.detailVerts = reinterpret_cast<float*>(d); d += dvSize;
  t.detailTris = d; d += dtSize;
  t.bvTree = bvSize ? reinterpret_cast<DtBVNode*>(d) : nullptr; d += bvSize;
  t.offMeshCons = reinterpret_cast<DtOffMeshConnection*>(d);
  return true;
}

int HostNavMesh::tilesAt(int x, int y, int* out, int maxOut) const {
  int n = 0;
  int ti = posLookup_[tileHash(x, y, tileLutMask_)];
  while (ti >= 0) {
    const HostTile& t = tiles_[ti];
    if (t.header && t.header->x == x && t.header->y == y)
      if (n < maxOut) out[n++] = ti;
    ti = t.next;
  }
  return n;
}

void HostNavMesh::connectIntLinks(int ti) {
  HostTile& t = tiles_[ti];
  const uint32_t base = polyRefBase(ti);
  for (int i = 0; i < t.header->polyCount; ++i) {
    DtPoly& poly = t.polys[i];
    poly.firstLink = kNullLink;
    if ((poly.areaAndtype >> 6) == 1) continue;  // off-mesh connection poly
    // back to front, so the list runs from the lowest edge to the highest
    for (int j = poly.vertCount - 1; j >= 0; --j) {
      if (poly.neis[j] == 0 || (poly.neis[j] & kExtLink)) continue;
      const uint32_t idx = allocLink(t);
      if (idx == kNullLink) continue;
      DtLink& l = t.links[idx];
      l.ref = base | static_cast<uint32_t>(poly.neis[j] - 1);
      l.edge = static_cast<uint8_t>(j);
      l.side = 0xff;
      l.bmin = l.bmax = 0;
      l.next = poly.firstLink;
      poly.firstLink = idx;
    }
  }
}

int HostNavMesh::findConnectingPolys(const float* va, const float* vb, int target, int side,
                                     uint32_t* con, float* conarea, int maxcon) const {
  if (target < 0) return 0;
  const HostTile& t = tiles_[target];
  float amin[2], amax[2];
  slabEndPoints(va, vb, amin, amax, side);
  const float apos = slabCoord(va, side);
  float bmin[2], bmax[2];
  const uint16_t m = kExtLink | static_cast<uint16_t>(side);
  int n = 0;
  const uint32_t base = polyRefBase(target);
  for (int i = 0; i < t.header->polyCount; ++i) {
    const DtPoly& poly = t.polys[i];
    const int nv = poly.vertCount;
    for (int j = 0; j < nv; ++j) {
      if (poly.neis[j] != m) continue;
      const float* vc = &t.verts[poly.verts[j] * 3];
      const float* vd = &t.verts[poly.verts[(j + 1) % nv] * 3];
      const float bpos = slabCoord(vc, side);
      if (fabsf(apos - bpos) > 0.01f) continue;
      slabEndPoints(vc, vd, bmin, bmax, side);
      if (!overlapSlabs(amin, amax, bmin, bmax, 0.01f, t.header->walkableClimb)) continue;
      if (n < maxcon) {
        conarea[n * 2 + 0] = fmaxf2(amin[0], bmin[0]);
        conarea[n * 2 + 1] = fminf2(amax[0], bmax[0]);
        con[n] = base | static_cast<uint32_t>(i);
        n++;
      }
      break;
    }
  }
  return n;
}

void HostNavMesh::connectExtLinks(int ti, int target, int side) {
  HostTile& t = tiles_[ti];
  for (int i = 0; i < t.header->polyCount; ++i) {
    DtPoly& poly = t.polys[i];
    const int nv = poly.vertCount;
    for (int j = 0; j < nv; ++j) {
      if ((poly.neis[j] & kExtLink) == 0) continue;
      const int dir = static_cast<int>(poly.neis[j] & 0xff);
      if (side != -1 && dir != side) continue;
      const float* va = &t.verts[poly.verts[j] * 3];
      const float* vb = &t.verts[poly.verts[(j + 1) % nv] * 3];
      uint32_t nei[4];
      float neia[4 * 2];
      const int nnei = findConnectingPolys(va, vb, target, oppositeTile(dir), nei, neia, 4);
      for (int k = 0; k < nnei; ++k) {
        const uint32_t idx = allocLink(t);
        if (idx == kNullLink) continue;
        DtLink& l = t.links[idx];
        l.ref = nei[k];
        l.edge = static_cast<uint8_t>(j);
        l.side = static_cast<uint8_t>(dir);
        l.next = poly.firstLink;
        poly.firstLink = idx;
        // portal limits squeezed into a byte each
        if (dir == 0 || dir == 4) {
          float tmin = (neia[k * 2 + 0] - va[2]) / (vb[2] - va[2]);
          float tmax = (neia[k * 2 + 1] - va[2]) / (vb[2] - va[2]);
          if (tmin > tmax) std::swap(tmin, tmax);
          l.bmin = static_cast<uint8_t>(roundf(clampf(tmin, 0.0f, 1.0f) * 255.0f));
          l.bmax = static_cast<uint8_t>(roundf(clampf(tmax, 0.0f, 1.0f) * 255.0f));
        } else if (dir == 2 || dir == 6) {
          float tmin = (neia[k * 2 + 0] - va[0]) / (vb[0] - va[0]);
          float tmax = (neia[k * 2 + 1] - va[0]) / (vb[0] - va[0]);
          if (tmin > tmax) std::swap(tmin, tmax);
          l.bmin = static_cast<uint8_t>(roundf(clampf(tmin, 0.0f, 1.0f) * 255.0f));
          l.bmax = static_cast<uint8_t>(roundf(clampf(tmax, 0.0f, 1.0f) * 255.0f));
        }
      }
    }
  }
}

bool HostNavMesh::addTile(const uint8_t* data, size_t len, uint32_t lastRef, bool buildLinks,
                          std::string& err) {
  if (len < sizeof(DtMeshHeader)) { err = "tile blob too small"; return false; }
  const DtMeshHeader* hdr = reinterpret_cast<const DtMeshHeader*>(data);
  if (hdr->magic != kDtNavMeshMagic) { err = "tile: wrong magic"; return false; }
  if (hdr->version != kDtNavMeshVersion) { err = "tile: wrong version"; return false; }
  if (polyBits_ < ilog2(nextPow2(static_cast<uint32_t>(hdr->polyCount)))) {
    err = "tile has more polys than maxPolys allows";
    return false;
  }
  {  // location must be free
    int tmp[32];
    const int n = tilesAt(hdr->x, hdr->y, tmp, 32);
    for (int i = 0; i < n; ++i)
      if (tiles_[tmp[i]].header->layer == hdr->layer) { err = "tile location occupied"; return false; }
  }
  int ti = -1;
  uint32_t salt = 1;
  if (!lastRef) {
    if (freeList_.empty()) { err = "out of tiles"; return false; }
    ti = freeList_.front();
    freeList_.erase(freeList_.begin());
  } else {
    const int want = static_cast<int>(decodeTile(lastRef));
    if (want >= params_.maxTiles) { err = "tileRef index beyond maxTiles"; return false; }
    auto it = std::find(freeList_.begin(), freeList_.end(), want);
    if (it == freeList_.end()) { err = "tileRef slot not free"; return false; }
    freeList_.erase(it);
    ti = want;
    salt = decodeSalt(lastRef);
  }
  HostTile& t = tiles_[ti];
  t.salt = salt;
  t.data.assign(data, data + len);
  if (!patchPointers(t, err)) return false;
  if (buildLinks && t.header->offMeshConCount > 0) {
    err = "MSET tile with off-mesh connections is not supported (habitat never builds them)";
    return false;
  }
  t.present = true;
  const int h = tileHash(t.header->x, t.header->y, tileLutMask_);
  t.next = posLookup_[h];
  posLookup_[h] = ti;

  for (int k = 0; k < 3; ++k) {
    bounds_[k] = boundsInit_ ? std::min(bounds_[k], t.header->bmin[k]) : t.header->bmin[k];
    bounds_[3 + k] = boundsInit_ ? std::max(bounds_[3 + k], t.header->bmax[k]) : t.header->bmax[k];
  }
  boundsInit_ = true;
  if (!buildLinks) return true;

  // fresh free list over all link slots, then internal links, then borders
  t.linksFreeList = 0;
  if (t.header->maxLinkCount > 0) {
    t.links[t.header->maxLinkCount - 1].next = kNullLink;
    for (int i = 0; i < t.header->maxLinkCount - 1; ++i) t.links[i].next = i + 1;
  } else {
    t.linksFreeList = kNullLink;
  }
  connectIntLinks(ti);
  int neis[32];
  int nneis = tilesAt(t.header->x, t.header->y, neis, 32);
  for (int j = 0; j < nneis; ++j) {
    if (neis[j] == ti) continue;
    connectExtLinks(ti, neis[j], -1);
    connectExtLinks(neis[j], ti, -1);
  }
  for (int i = 0; i < 8; ++i) {
    int nx = t.header->x, ny = t.header->y;
    switch (i) {
      case 0: nx++; break;
      case 1: nx++; ny++; break;
      case 2: ny++; break;
      case 3: nx--; ny++; break;
      case 4: nx--; break;
      case 5: nx--; ny--; break;
      case 6: ny--; break;
      case 7: nx++; ny--; break;
    }
    nneis = tilesAt(nx, ny, neis, 32);
    for (int j = 0; j < nneis; ++j) {
      connectExtLinks(ti, neis[j], i);
      connectExtLinks(neis[j], ti, oppositeTile(i));
    }
  }
  return true;
}

bool HostNavMesh::addFinalisedTile(const uint8_t* data, size_t len, uint32_t tileRef,
                                   std::string& err) {
  return addTile(data, len, tileRef, false, err);
}

bool HostNavMesh::loadMSET(const uint8_t* buf, size_t len, std::string& err) {
  size_t off = 0;
  auto rd = [&](void* dst, size_t n) {
    if (off + n > len) return false;
    memcpy(dst, buf + off, n);
    off += n;
    return true;
  };
  struct { int32_t magic, version, numTiles; DtNavMeshParams params; } header;
  static_assert(sizeof(header) == 40, "NavMeshSetHeader");
  if (!rd(&header, sizeof(header))) { err = "navmesh image truncated (header)"; return false; }
  if (header.magic != kMsetMagic) { err = "not an MSET navmesh image"; return false; }
  if (header.version < 1 || header.version > 2) { err = "unsupported MSET version"; return false; }
  hasSettings_ = false;
  if (header.version >= 2) {
    if (!rd(settings_, 56)) { err = "navmesh image truncated (settings)"; return false; }
    hasSettings_ = true;
  }
  if (!init(header.params, err)) return false;
  for (int i = 0; i < header.numTiles; ++i) {
    struct { uint32_t tileRef; int32_t dataSize; } th;
    if (!rd(&th, sizeof(th))) { err = "navmesh image truncated (tile header)"; return false; }
    if (!th.tileRef || !th.dataSize) break;
    if (off + static_cast<size_t>(th.dataSize) > len) { err = "navmesh image truncated (tile)"; return false; }
    // (the reference ignores addTile's status, PathFinder.cpp:1157; a failing tile would
    // crash it on the next line, so failing loudly here loses nothing)
    if (!addTile(buf + off, th.dataSize, th.tileRef, true, err)) return false;
    off += th.dataSize;
  }
  return true;
}

// PathFinder::Impl::saveNavMesh, PathFinder.cpp:1177-1223
bool HostNavMesh::saveMSET(std::vector<uint8_t>& out, std::string& err) const {
  if (!hasSettings_) {
    err = "NavMeshSettings weren't set. Either build or load a navmesh before saving";
    return false;
  }
  struct { int32_t magic, version, numTiles; DtNavMeshParams params; } header;
  header.magic = kMsetMagic;
  header.version = 2;
  header.numTiles = 0;
  for (const HostTile& t : tiles_)
    if (t.present && !t.data.empty()) ++header.numTiles;
  header.params = params_;
  out.clear();
  auto wr = [&](const void* p, size_t n) {
    const uint8_t* c = static_cast<const uint8_t*>(p);
    out.insert(out.end(), c, c + n);
  };
  wr(&header, sizeof(header));
  wr(settings_, 56);
  for (size_t i = 0; i < tiles_.size(); ++i) {
    const HostTile& t = tiles_[i];
    if (!t.present || t.data.empty()) continue;
    struct { uint32_t tileRef; int32_t dataSize; } th;
    th.tileRef = polyRefBase(static_cast<int>(i));  // dtNavMesh::getTileRef: poly 0 of the tile
    th.dataSize = static_cast<int32_t>(t.data.size());
    wr(&th, sizeof(th));
    wr(t.data.data(), t.data.size());
  }
  return true;
}

void HostNavMesh::setSettings(const uint8_t* raw56) {
  memcpy(settings_, raw56, 56);
  hasSettings_ = true;
}

// IslandSystem constructor, PathFinder.cpp:173-206 + expandFrom :409-459
void HostNavMesh::floodIslands() {
  polyIsland_.assign(tiles_.size(), {});
  for (size_t i = 0; i < tiles_.size(); ++i)
    if (tiles_[i].present) polyIsland_[i].assign(tiles_[i].header->polyCount, -1);
  islandRadius_.clear();
  std::vector<float> iv;  // vertices of the island being grown, visit order
  std::vector<uint32_t> stack;
  for (size_t it = 0; it < tiles_.size(); ++it) {
    if (!tiles_[it].present) continue;
    for (int jp = 0; jp < tiles_[it].header->polyCount; ++jp) {
      if (polyIsland_[it][jp] != -1) continue;
      const int32_t id = static_cast<int32_t>(islandRadius_.size());
      polyIsland_[it][jp] = id;
      iv.clear();
      stack.clear();
      stack.push_back(encodePolyId(tiles_[it].salt, static_cast<uint32_t>(it), jp));
      while (!stack.empty()) {
        const uint32_t ref = stack.back();
        stack.pop_back();
        const HostTile& t = tiles_[decodeTile(ref)];
        const DtPoly& poly = t.polys[decodePoly(ref)];
        for (int k = 0; k < poly.vertCount; ++k) {
          const float* v = &t.verts[static_cast<size_t>(poly.verts[k]) * 3];
          iv.push_back(v[0]); iv.push_back(v[1]); iv.push_back(v[2]);
        }
        for (uint32_t l = poly.firstLink; l != kNullLink; l = t.links[l].next) {
          const uint32_t nref = t.links[l].ref;
          const uint32_t nt = decodeTile(nref), np = decodePoly(nref);
          if (polyIsland_[nt][np] != -1) continue;
          const DtPoly& npoly = tiles_[nt].polys[np];
          if ((npoly.flags & kFlagWalk) == 0) continue;  // passFilter: include WALK, exclude 0
          polyIsland_[nt][np] = id;
          stack.push_back(nref);
        }
      }
      float cx = 0.f, cy = 0.f, cz = 0.f;
      const size_t n = iv.size() / 3;
      for (size_t k = 0; k < n; ++k) { cx += iv[3 * k]; cy += iv[3 * k + 1]; cz += iv[3 * k + 2]; }
      const float fn = static_cast<float>(n);
      cx /= fn; cy /= fn; cz /= fn;
      float maxRadius = 0.0f;
      for (size_t k = 0; k < n; ++k) {
        const float dx = iv[3 * k] - cx, dy = iv[3 * k + 1] - cy, dz = iv[3 * k + 2] - cz;
        float d = 0.f;
        d += dx * dx; d += dy * dy; d += dz * dz;
        maxRadius = std::max(maxRadius, sqrtf(d));
      }
      islandRadius_.push_back(maxRadius);
    }
  }
}

// removeZeroAreaPolys, PathFinder.cpp:1045-1085 (polyArea :1035-1043)
void HostNavMesh::zeroAreaAndAreas() {
  islandArea_.assign(islandRadius_.size(), 0.0f);
  for (size_t it = 0; it < tiles_.size(); ++it) {
    HostTile& t = tiles_[it];
    if (!t.present) continue;
    for (int jp = 0; jp < t.header->polyCount; ++jp) {
      DtPoly& poly = t.polys[jp];
      const DtPolyDetail& pd = t.detailMeshes[jp];
      float area = 0.f;
      for (int j = 0; j < pd.triCount; ++j) {
        const uint8_t* tri = &t.detailTris[static_cast<size_t>(pd.triBase + j) * 4];
        const float* v[3];
        for (int k = 0; k < 3; ++k) {
          if (tri[k] < poly.vertCount)
            v[k] = &t.verts[static_cast<size_t>(poly.verts[tri[k]]) * 3];
          else
            v[k] = &t.detailVerts[static_cast<size_t>(pd.vertBase + (tri[k] - poly.vertCount)) * 3];
        }
        const float ax = v[1][0] - v[0][0], ay = v[1][1] - v[0][1], az = v[1][2] - v[0][2];
        const float bx = v[2][0] - v[1][0], by = v[2][1] - v[1][1], bz = v[2][2] - v[1][2];
        const float cx = ay * bz - by * az, cy = az * bx - bz * ax, cz = ax * by - bx * ay;
        float d = 0.f;
        d += cx * cx; d += cy * cy; d += cz * cz;
        area += 0.5f * sqrtf(d);
      }
      if (area < 1e-5f) {
        poly.flags = kFlagDisabled;
      } else if ((poly.flags & kFlagWalk) != 0) {
        const int32_t isl = polyIsland_[it][jp];
        if (isl >= 0 && isl < static_cast<int32_t>(islandArea_.size())) islandArea_[isl] += area;
      }
    }
  }
  // The reference sums the per-island areas in the ITERATION ORDER of a std::unordered_map<uint32_t, float>
  // (PathFinder.cpp:1079-1083) whose keys were inserted in the iteration order of another one
  // (islandsToPolys_, keys inserted as the islands were first met: 0, 1, 2, ...; PathFinder.cpp:1047-1051).
  // f32 addition is not associative, so the same two containers are filled in the same way here and the
  // sum runs in their order.
  std::unordered_map<uint32_t, int> islandsToPolys;
  for (uint32_t id = 0; id < islandArea_.size(); ++id) islandsToPolys[id] = 0;
  std::unordered_map<uint32_t, float> islandsToArea;
  islandsToArea.reserve(islandsToPolys.size());
  for (auto& itr : islandsToPolys) islandsToArea[itr.first] = 0.0f;
  for (uint32_t id = 0; id < islandArea_.size(); ++id) islandsToArea[id] = islandArea_[id];
  totalArea_ = 0.f;
  for (auto& itr : islandsToArea) totalArea_ += itr.second;
}

void HostNavMesh::finish(const int32_t* givenIslands, const float* givenRadii, int numRadii) {
  floodIslands();
  if (givenIslands) {
    size_t k = 0;
    int32_t maxId = -1;
    for (size_t it = 0; it < tiles_.size(); ++it) {
      if (!tiles_[it].present) continue;
      for (int jp = 0; jp < tiles_[it].header->polyCount; ++jp) {
        polyIsland_[it][jp] = givenIslands[k++];
        maxId = std::max(maxId, polyIsland_[it][jp]);
      }
    }
    // radii cannot be recovered from ids alone in the caller's order; recompute per id
    std::vector<std::vector<float>> iv(maxId + 1);
    for (size_t it = 0; it < tiles_.size(); ++it) {
      if (!tiles_[it].present) continue;
      const HostTile& t = tiles_[it];
      for (int jp = 0; jp < t.header->polyCount; ++jp) {
        const int32_t id = polyIsland_[it][jp];
        if (id < 0) continue;
        for (int k2 = 0; k2 < t.polys[jp].vertCount; ++k2) {
          const float* v = &t.verts[static_cast<size_t>(t.polys[jp].verts[k2]) * 3];
          iv[id].insert(iv[id].end(), v, v + 3);
        }
      }
    }
    islandRadius_.assign(maxId + 1, 0.f);
    if (givenRadii && numRadii == maxId + 1) {
      islandRadius_.assign(givenRadii, givenRadii + numRadii);
      iv.clear();
    }
    for (int32_t id = 0; id <= maxId && !iv.empty(); ++id) {
      const size_t n = iv[id].size() / 3;
      if (!n) continue;
      float cx = 0, cy = 0, cz = 0;
      for (size_t k2 = 0; k2 < n; ++k2) { cx += iv[id][3 * k2]; cy += iv[id][3 * k2 + 1]; cz += iv[id][3 * k2 + 2]; }
      cx /= n; cy /= n; cz /= n;
      float r = 0;
      for (size_t k2 = 0; k2 < n; ++k2) {
        const float dx = iv[id][3 * k2] - cx, dy = iv[id][3 * k2 + 1] - cy, dz = iv[id][3 * k2 + 2] - cz;
        r = std::max(r, sqrtf(dx * dx + dy * dy + dz * dz));
      }
      islandRadius_[id] = r;
    }
  }
  zeroAreaAndAreas();
}

void HostNavMesh::flatten(FlatNav& out) const {
  out = FlatNav();
  out.params = params_;
  out.polyBits = polyBits_;
  out.tileBits = tileBits_;
  out.saltBits = saltBits_;
  out.islandRadius = islandRadius_;
  out.islandArea = islandArea_;
  out.totalArea = totalArea_;
  memcpy(out.bounds, bounds_, sizeof(bounds_));
  memcpy(out.settings, settings_, 56);
  out.hasSettings = hasSettings_;

  const size_t nt = tiles_.size();
  out.tiles.assign(nt, TileRec{});
  std::vector<uint32_t> polyStart(nt + 1, 0);
  for (size_t it = 0; it < nt; ++it)
    polyStart[it + 1] = polyStart[it] + (tiles_[it].present ? tiles_[it].header->polyCount : 0);
  out.polys.assign(polyStart[nt], PolyRec{});

  auto globalOf = [&](uint32_t ref) -> uint32_t {
    if (!ref) return kNoPoly;
    const uint32_t t = decodeTile(ref), p = decodePoly(ref);
    if (t >= nt || !tiles_[t].present || static_cast<int>(p) >= tiles_[t].header->polyCount)
      return kNoPoly;
    return polyStart[t] + p;
  };

  // pass 1: polys, detail meshes, BV trees, link windows
  for (size_t it = 0; it < nt; ++it) {
    const HostTile& t = tiles_[it];
    TileRec& tr = out.tiles[it];
    tr.polyStart = polyStart[it];
    if (!t.present) continue;
    const DtMeshHeader& h = *t.header;
    for (int k = 0; k < 3; ++k) { tr.bmin[k] = h.bmin[k]; tr.bmax[k] = h.bmax[k]; }
    tr.bvQuantFactor = h.bvQuantFactor;
    tr.walkableClimb = h.walkableClimb;
    tr.polyCount = h.polyCount;
    tr.x = h.x; tr.y = h.y; tr.layer = h.layer;
    tr.refBase = polyRefBase(static_cast<int>(it));
    tr.pad[0] = 1;  // present
    const uint32_t detTriBase = static_cast<uint32_t>(out.detTris.size() / 4);
    const uint32_t detVertBase = static_cast<uint32_t>(out.detVerts.size() / 3);
    out.detTris.insert(out.detTris.end(), t.detailTris, t.detailTris + 4 * static_cast<size_t>(h.detailTriCount));
    out.detVerts.insert(out.detVerts.end(), t.detailVerts, t.detailVerts + 3 * static_cast<size_t>(h.detailVertCount));
    tr.bvStart = static_cast<uint32_t>(out.bv.size());
    tr.bvCount = t.bvTree ? h.bvNodeCount : 0;
    for (uint32_t b = 0; b < tr.bvCount; ++b) {
      const DtBVNode& n = t.bvTree[b];
      BvRec r;
      for (int k = 0; k < 3; ++k) { r.bmin[k] = n.bmin[k]; r.bmax[k] = n.bmax[k]; }
      if (n.i >= 0) {
        const DtPoly& p = t.polys[n.i];
        r.i = static_cast<int32_t>(polyStart[it] + n.i);
        if ((p.flags & kFlagWalk) == 0) r.i |= kBvFailBit;
      } else {
        r.i = n.i;
      }
      out.bv.push_back(r);
    }
    for (int jp = 0; jp < h.polyCount; ++jp) {
      const DtPoly& p = t.polys[jp];
      PolyRec& pr = out.polys[polyStart[it] + jp];
      for (int k = 0; k < p.vertCount && k < 6; ++k)
        memcpy(&pr.v[3 * k], &t.verts[static_cast<size_t>(p.verts[k]) * 3], 12);
      pr.ref = tr.refBase | static_cast<uint32_t>(jp);
      pr.flags = p.flags;
      pr.nv = p.vertCount;
      pr.areaType = p.areaAndtype;
      memcpy(pr.neis, p.neis, sizeof(pr.neis));
      pr.island = polyIsland_.size() > it && polyIsland_[it].size() > static_cast<size_t>(jp)
                      ? polyIsland_[it][jp] : -1;
      pr.tile = static_cast<uint32_t>(it);
      if (jp < h.detailMeshCount) {
        pr.detTriBase = detTriBase + t.detailMeshes[jp].triBase;
        pr.detVertBase = detVertBase + t.detailMeshes[jp].vertBase;
        pr.detTriCount = t.detailMeshes[jp].triCount;
      }
      float a = 0.0f;  // DetourNavMeshQuery.cpp:270-277
      for (int j = 2; j < p.vertCount; ++j) {
        const float* va = &pr.v[0];
        const float* vb = &pr.v[(j - 1) * 3];
        const float* vc = &pr.v[j * 3];
        const float abx = vb[0] - va[0], abz = vb[2] - va[2];
        const float acx = vc[0] - va[0], acz = vc[2] - va[2];
        a += acx * abz - abx * acz;
      }
      pr.area2d = a;
      pr.linkStart = static_cast<uint32_t>(out.links.size());
      uint32_t cnt = 0;
      for (uint32_t l = p.firstLink; l != kNullLink; l = t.links[l].next) {
        LinkRec lr{};
        lr.neiRef = t.links[l].ref;
        lr.nei = globalOf(t.links[l].ref);
        const uint32_t side = t.links[l].side;
        const uint32_t state = side != 0xff ? (side >> 1) & 3u : 0u;
        lr.meta = t.links[l].edge | (side << kLinkSideShift) | (state << kLinkStateShift);
        for (uint32_t l2 = p.firstLink; l2 != l; l2 = t.links[l2].next)
          if (lr.nei != kNoPoly && t.links[l2].ref == t.links[l].ref) { lr.meta |= kLinkDupBit; break; }
        out.links.push_back(lr);
        // portal of the FIRST link of this poly with the same ref (getPortalPoints :2276-2285)
        PortalRec po{};
        const DtLink* first = nullptr;
        for (uint32_t l2 = p.firstLink; l2 != kNullLink; l2 = t.links[l2].next)
          if (t.links[l2].ref == t.links[l].ref) { first = &t.links[l2]; break; }
        if (lr.nei != kNoPoly && first) {
          const uint32_t ntile = decodeTile(first->ref), npoly = decodePoly(first->ref);
          const DtPoly& q = tiles_[ntile].polys[npoly];
          if ((p.areaAndtype >> 6) == 1) {  // from is off-mesh
            const float* v = &t.verts[static_cast<size_t>(p.verts[first->edge]) * 3];
            memcpy(po.l, v, 12); memcpy(po.r, v, 12);
          } else if ((q.areaAndtype >> 6) == 1) {  // to is off-mesh
            const HostTile& tt = tiles_[ntile];
            for (uint32_t l3 = q.firstLink; l3 != kNullLink; l3 = tt.links[l3].next)
              if (tt.links[l3].ref == pr.ref) {
                const float* v = &tt.verts[static_cast<size_t>(q.verts[tt.links[l3].edge]) * 3];
                memcpy(po.l, v, 12); memcpy(po.r, v, 12);
                break;
              }
          } else {
            const float* v0 = &t.verts[static_cast<size_t>(p.verts[first->edge]) * 3];
            const float* v1 = &t.verts[static_cast<size_t>(p.verts[(first->edge + 1) % p.vertCount]) * 3];
            memcpy(po.l, v0, 12); memcpy(po.r, v1, 12);
            if (first->side != 0xff && (first->bmin != 0 || first->bmax != 255)) {
              const float s = 1.0f / 255.0f;
              const float tmin = first->bmin * s, tmax = first->bmax * s;
              for (int k = 0; k < 3; ++k) {
                po.l[k] = v0[k] + (v1[k] - v0[k]) * tmin;
                po.r[k] = v0[k] + (v1[k] - v0[k]) * tmax;
              }
            }
          }
        }
        out.portals.push_back(po);
        LinkRec& stored = out.links.back();
        for (int k = 0; k < 3; ++k) stored.mid[k] = (po.l[k] + po.r[k]) * 0.5f;
        cnt++;
      }
      pr.linkCount = static_cast<uint8_t>(std::min<uint32_t>(cnt, 255u));
    }
  }
  // bounds of each poly's vertices and detail vertices
  out.polyBox.assign(out.polys.size() * 8, 0.f);
  for (size_t g = 0; g < out.polys.size(); ++g) {
    const PolyRec& pr = out.polys[g];
    float* b = &out.polyBox[g * 8];
    for (int k = 0; k < 3; ++k) { b[k] = pr.nv ? pr.v[k] : 0.f; b[4 + k] = b[k]; }
    auto grow = [&](const float* v) {
      for (int k = 0; k < 3; ++k) {
        b[k] = std::min(b[k], v[k]);
        b[4 + k] = std::max(b[4 + k], v[k]);
      }
    };
    for (int j = 1; j < pr.nv; ++j) grow(&pr.v[3 * j]);
    for (int t = 0; t < pr.detTriCount; ++t) {
      const unsigned char* tri = &out.detTris[(static_cast<size_t>(pr.detTriBase) + t) * 4];
      for (int k = 0; k < 3; ++k)
        if (tri[k] >= pr.nv) grow(&out.detVerts[(static_cast<size_t>(pr.detVertBase) + (tri[k] - pr.nv)) * 3]);
    }
    // layout: min xyz at [0..2], max xyz at [4..6]
  }
  // do the BV leaf boxes cover their polys in xz?  (their y range does not, see hbn_snap.h)
  // Leaves behind the end of the tree (index >= the root's escape) are marked loose instead.
  out.bvXzTight = 1;
  for (const TileRec& tr : out.tiles) {
    if (!tr.bvCount) continue;
    const float inv = 1.0f / tr.bvQuantFactor;
    const int32_t rootI = out.bv[tr.bvStart].i;
    const uint32_t treeSize = rootI < 0 ? static_cast<uint32_t>(-rootI) : 1u;
    for (uint32_t b = 0; b < tr.bvCount; ++b) {
      BvRec& n = out.bv[tr.bvStart + b];
      if (n.i < 0) continue;
      if (b >= treeSize) {
        n.i |= kBvLooseBit;
        continue;
      }
      const float* box = &out.polyBox[static_cast<size_t>(n.i & kBvIndexMask) * 8];
      for (int k = 0; k < 3; k += 2) {
        const float lo = tr.bmin[k] + n.bmin[k] * inv, hi = tr.bmin[k] + (n.bmax[k] + 1) * inv;
        if (box[k] < lo - 1e-3f || box[4 + k] > hi + 1e-3f) out.bvXzTight = 0;
      }
    }
  }
  // pass 2: neighbour windows + filter bits (needs every poly's link window), and the dense
  // enumeration of A* node keys (poly, crossSide): crossSide 0 of every poly, plus the sides
  // through which tile-border links enter it
  std::vector<uint8_t> sideMask(out.polys.size(), 1);
  for (const LinkRec& lr : out.links)
    if (lr.nei != kNoPoly) sideMask[lr.nei] |= static_cast<uint8_t>(1u << ((lr.meta >> kLinkStateShift) & 3u));
  // Keys are numbered along a space-filling curve, not in poly order: 2 m height layers, Morton order
  // of the poly centroids (1 m cells) within a layer.  A search explores a compact region, so its
  // nodes then fall into few 32 B sectors of the per-query node table (C4: 0.085 sectors per node
  // instead of 0.243 in poly order; profiles/r1_summary.md).  The numbering is internal to the table:
  // results do not depend on it.  HBN_KEY_ORDER=0 keeps poly order (for A/B timing).
  std::vector<uint32_t> order(out.polys.size());
  for (size_t g = 0; g < order.size(); ++g) order[g] = static_cast<uint32_t>(g);
  const char* keyOrderEnv = getenv("HBN_KEY_ORDER");
  if (!(keyOrderEnv && atoi(keyOrderEnv) == 0) && !order.empty()) {
    std::vector<float> cen(out.polys.size() * 3, 0.f);
    float lo[3] = {0.f, 0.f, 0.f};
    bool have = false;
    for (size_t g = 0; g < out.polys.size(); ++g) {
      const PolyRec& p = out.polys[g];
      const int nv = p.nv > 0 ? p.nv : 1;
      for (int k = 0; k < p.nv; ++k)
        for (int a = 0; a < 3; ++a) cen[3 * g + a] += p.v[3 * k + a];
      for (int a = 0; a < 3; ++a) {
        float c = cen[3 * g + a] / static_cast<float>(nv);
        if (!(c == c) || c > 1e30f || c < -1e30f) c = 0.f;
        cen[3 * g + a] = c;
        if (!have || c < lo[a]) lo[a] = c;
      }
      have = true;
    }
    auto spread16 = [](uint32_t v) {
      v &= 0xffffu;
      v = (v | (v << 8)) & 0x00ff00ffu;
      v = (v | (v << 4)) & 0x0f0f0f0fu;
      v = (v | (v << 2)) & 0x33333333u;
      v = (v | (v << 1)) & 0x55555555u;
      return v;
    };
    std::vector<uint64_t> curve(out.polys.size());
    for (size_t g = 0; g < out.polys.size(); ++g) {
      const uint32_t ix = static_cast<uint32_t>(std::min(65535.f, std::max(0.f, cen[3 * g] - lo[0])));
      const uint32_t iz = static_cast<uint32_t>(std::min(65535.f, std::max(0.f, cen[3 * g + 2] - lo[2])));
      const uint32_t layer = static_cast<uint32_t>(std::min(1e6f, std::max(0.f, (cen[3 * g + 1] - lo[1]) * 0.5f + 0.5f)));
      curve[g] = (static_cast<uint64_t>(layer) << 32) | (spread16(ix) | (spread16(iz) << 1));
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return curve[a] < curve[b]; });
  }
  uint32_t nKeys = 0;
  for (const uint32_t g : order) {
    out.polys[g].key0 = nKeys;
    nKeys += static_cast<uint32_t>(__builtin_popcount(sideMask[g]));
  }
  out.numKeys = nKeys;
  for (LinkRec& lr : out.links) {
    if (lr.nei == kNoPoly) continue;
    const PolyRec& q = out.polys[lr.nei];
    lr.neiLinkStart = q.linkStart;
    lr.meta |= static_cast<uint32_t>(q.linkCount) << kLinkNeiCountShift;
    if ((q.flags & kFlagWalk) != 0) lr.meta |= kLinkPassBit;
    if ((q.areaType >> 6) == 1) lr.meta |= kLinkOffmeshBit;
    const uint32_t st = (lr.meta >> kLinkStateShift) & 3u;
    lr.neiKey = q.key0 + static_cast<uint32_t>(__builtin_popcount(sideMask[lr.nei] & ((1u << st) - 1u)));
  }
  // tile grid in bucket-chain order (dtNavMesh::getTilesAt)
  bool any = false;
  int minx = 0, miny = 0, maxx = 0, maxy = 0;
  for (size_t it = 0; it < nt; ++it) {
    if (!tiles_[it].present) continue;
    const int x = tiles_[it].header->x, y = tiles_[it].header->y;
    if (!any) { minx = maxx = x; miny = maxy = y; any = true; }
    minx = std::min(minx, x); maxx = std::max(maxx, x);
    miny = std::min(miny, y); maxy = std::max(maxy, y);
  }
  out.gridMinX = minx; out.gridMinY = miny;
  out.gridW = any ? maxx - minx + 1 : 0;
  out.gridH = any ? maxy - miny + 1 : 0;
  out.gridStart.assign(static_cast<size_t>(out.gridW) * out.gridH + 1, 0);
  for (int y = 0; y < out.gridH; ++y)
    for (int x = 0; x < out.gridW; ++x) {
      int tmp[32];
      const int n = tilesAt(x + minx, y + miny, tmp, 32);
      out.gridStart[static_cast<size_t>(y) * out.gridW + x] = static_cast<uint32_t>(out.tileOrder.size());
      for (int k = 0; k < n; ++k) out.tileOrder.push_back(static_cast<uint32_t>(tmp[k]));
    }
  out.gridStart.back() = static_cast<uint32_t>(out.tileOrder.size());
  if (out.tileOrder.empty()) out.tileOrder.push_back(0);

  // findRandomPoint tables
  out.tileIslStart.assign(nt + 1, 0);
  for (size_t it = 0; it < nt; ++it) {
    TileRec& tr = out.tiles[it];
    out.tileIslStart[it] = static_cast<uint32_t>(out.tileIslId.size());
    if (!tiles_[it].present) continue;
    tr.randStart = static_cast<uint32_t>(out.randEntries.size());
    float areaSum = 0.0f;
    std::vector<int32_t> islands;
    for (uint32_t jp = 0; jp < tr.polyCount; ++jp) {
      const PolyRec& pr = out.polys[tr.polyStart + jp];
      if ((pr.areaType >> 6) != 0) continue;
      if ((pr.flags & kFlagWalk) == 0) continue;
      areaSum += pr.area2d;
      out.randEntries.push_back(RandEntry{tr.polyStart + jp, pr.area2d, areaSum, 0});
      if (std::find(islands.begin(), islands.end(), pr.island) == islands.end())
        islands.push_back(pr.island);
    }
    tr.randCount = static_cast<uint32_t>(out.randEntries.size()) - tr.randStart;
    std::sort(islands.begin(), islands.end());
    for (int32_t isl : islands) {
      out.tileIslId.push_back(isl);
      out.tileIslWin.push_back(static_cast<uint32_t>(out.randEntries.size()));
      float s = 0.0f;
      uint32_t c = 0;
      for (uint32_t jp = 0; jp < tr.polyCount; ++jp) {
        const PolyRec& pr = out.polys[tr.polyStart + jp];
        if ((pr.areaType >> 6) != 0 || (pr.flags & kFlagWalk) == 0 || pr.island != isl) continue;
        s += pr.area2d;
        out.randEntries.push_back(RandEntry{tr.polyStart + jp, pr.area2d, s, 0});
        c++;
      }
      out.tileIslCnt.push_back(c);
    }
  }
  out.tileIslStart[nt] = static_cast<uint32_t>(out.tileIslId.size());
  if (out.randEntries.empty()) out.randEntries.push_back(RandEntry{0, 0, 0, 0});
  if (out.tileIslId.empty()) { out.tileIslId.push_back(-1); out.tileIslWin.push_back(0); out.tileIslCnt.push_back(0); }
  if (out.links.empty()) { out.links.push_back(LinkRec{}); out.portals.push_back(PortalRec{}); }
  if (out.bv.empty()) out.bv.push_back(BvRec{});
  if (out.detTris.empty()) out.detTris.assign(4, 0);
  if (out.detVerts.empty()) out.detVerts.assign(3, 0.f);
}

NavView FlatNav::view() const {
  NavView v{};
  v.polys = polys.data();
  v.links = links.data();
  v.portals = portals.data();
  v.bv = bv.data();
  v.tiles = tiles.data();
  v.detTris = detTris.data();
  v.detVerts = detVerts.data();
  v.polyBox = polyBox.data();
  v.gridStart = gridStart.data();
  v.tileOrder = tileOrder.data();
  v.randEntries = randEntries.data();
  v.tileIslStart = tileIslStart.data();
  v.tileIslId = tileIslId.data();
  v.tileIslWin = tileIslWin.data();
  v.tileIslCnt = tileIslCnt.data();
  v.gridMinX = gridMinX; v.gridMinY = gridMinY; v.gridW = gridW; v.gridH = gridH;
  for (int k = 0; k < 3; ++k) v.orig[k] = params.orig[k];
  v.tileWidth = params.tileWidth;
  v.tileHeight = params.tileHeight;
  v.numPolys = static_cast<uint32_t>(polys.size());
  v.numTiles = static_cast<uint32_t>(tiles.size());
  v.numLinks = static_cast<uint32_t>(links.size());
  v.numKeys = numKeys;
  v.bvXzTight = bvXzTight;
  v.polyBits = polyBits; v.tileBits = tileBits; v.saltBits = saltBits;
  v.numIslands = static_cast<int32_t>(islandRadius.size());
  return v;
}

}  // namespace hbn

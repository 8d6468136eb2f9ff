// =============================================================================
// oracle/ref_esp.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// THE oracle: the reference's own, UNMODIFIED src/esp/nav/PathFinder.cpp, compiled where it
// lies under /root/reference (the #include below; never copied), linked with the reference's
// own Detour + Recast, Corrade-Utility and Magnum (math only), following SURVEY.md 8c steps
// 1-5 (oracle/Makefile).  This file only adds a C ABI for ctypes (oracle/ref.py) around
// esp::nav::PathFinder; every PF-layer answer (snap, islands, find_path, multi-goal, try_step,
// wall distance, random points, top-down views, navmesh geometry, load/save) is computed by
// the reference's code, not by a restatement.
//
// Being in the same translation unit (and built with -fno-access-control) lets the harness
// read what PathFinder hides behind its pimpl -- the dtNavMesh, the dtNavMeshQuery, the
// filter and the IslandSystem -- so that the raw entry points (poly refs, corridors, Detour
// status words, node-pool use, tile blobs, per-poly island ids) come from the very objects
// the reference queries, again without restating anything.
//
// rand(): PF.cpp:1231-1234 draws from glibc rand().  This library defines rand()/srand()
// itself (linked -Bsymbolic-functions, so PathFinder.o binds to them): in mode 0 they forward
// to glibc, in mode 1 rand() returns the 31-bit value of the counter-based stream
// hbn_uniform(seed, query, draw) of include/hbn.h -- the unmodified frand() then yields
// bit-identical uniforms to the GPU library's (SURVEY trap T7).
//
// Threads: esp::nav::PathFinder is not thread-safe (shared node pool, poly-flag mutation for
// island queries, process-global rand).  Batches run on one PathFinder per thread: clones
// made by saveNavMesh -> loadNavMesh of the same image (SURVEY 8d "one PathFinder instance
// per thread"), with the master's IslandSystem copied over the clone's, because a re-loaded
// mesh numbers its islands differently from a freshly built one (trap T5).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; workloads/ uses its Recast entry points to BUILD navmesh inputs.
// =============================================================================
#include "esp/nav/PathFinder.cpp"  // the reference, unmodified: -I/root/reference/src

#include <dlfcn.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "tiled_recast.h"

namespace hbnref {

namespace Mn = Magnum;
using esp::nav::PathFinder;
const float kNaN = std::numeric_limits<float>::quiet_NaN();
const float kInf = std::numeric_limits<float>::infinity();

// ---- rand() interposition ------------------------------------------------------
struct RandStream {
  int mode = 0;  // 0 = glibc rand(), 1 = counter based
  uint64_t seed = 0, query = 0;
  uint32_t draw = 0;
};
thread_local RandStream tlsRand;
std::mutex glibcRandMu;  // glibc rand() is process-global

// ---- the python bindings make ESP_CHECK throw (core/Check.cpp:14-33); so does the oracle --
void throwRuntimeError(const char* msg) { throw std::runtime_error(msg); }

struct LoggingOnce {
  esp::logging::LoggingContext ctx;  // "No current logging context" otherwise
  LoggingOnce() { esp::core::throwInPython = &throwRuntimeError; }
};
void ensureLogging() { static LoggingOnce once; }

std::string tmpPath() {
  static std::atomic<uint64_t> counter{0};
  const char* dir = access("/dev/shm", W_OK) == 0 ? "/dev/shm" : "/tmp";
  return std::string(dir) + "/hbnref_" + std::to_string(getpid()) + "_" +
         std::to_string(counter.fetch_add(1)) + ".navmesh";
}
bool writeFile(const std::string& path, const unsigned char* buf, size_t len) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  const size_t put = fwrite(buf, 1, len, fp);
  fclose(fp);
  return put == len;
}
bool readFile(const std::string& path, std::vector<unsigned char>& out) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return false;
  fseek(fp, 0, SEEK_END);
  const long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  out.resize(n);
  const size_t got = fread(out.data(), 1, n, fp);
  fclose(fp);
  return got == static_cast<size_t>(n);
}

struct Oracle {
  std::vector<std::unique_ptr<PathFinder>> pfs;  // [0] master, [t] clone for worker t
  std::mutex mu;

  Oracle() {
    ensureLogging();
    pfs.emplace_back(new PathFinder());
  }
  PathFinder& master() { return *pfs[0]; }
  PathFinder::Impl& impl(int t = 0) { return *pfs[t]->pimpl_; }
  void dropClones() { pfs.resize(1); }

  bool loadImage(const unsigned char* buf, size_t len) {
    dropClones();
    const std::string p = tmpPath();
    if (!writeFile(p, buf, len)) return false;
    const bool ok = master().loadNavMesh(p);
    unlink(p.c_str());
    return ok;
  }
  bool saveImage(std::vector<unsigned char>& out) {
    const std::string p = tmpPath();
    if (!master().saveNavMesh(p)) return false;
    const bool ok = readFile(p, out);
    unlink(p.c_str());
    return ok;
  }
  // clone t: same tile blobs through the reference's own save -> load, then the master's islands
  bool ensureWorkers(int n) {
    std::lock_guard<std::mutex> lk(mu);
    if (static_cast<int>(pfs.size()) >= n) return true;
    std::vector<unsigned char> image;
    if (!saveImage(image)) return false;
    const std::string p = tmpPath();
    if (!writeFile(p, image.data(), image.size())) return false;
    bool ok = true;
    while (static_cast<int>(pfs.size()) < n && ok) {
      std::unique_ptr<PathFinder> c(new PathFinder());
      ok = c->loadNavMesh(p);
      if (ok) {
        *c->pimpl_->islandSystem_ = *master().pimpl_->islandSystem_;
        c->pimpl_->bounds_ = master().pimpl_->bounds_;
        c->pimpl_->navMeshSettings_ = master().pimpl_->navMeshSettings_;
        pfs.emplace_back(std::move(c));
      }
    }
    unlink(p.c_str());
    return ok;
  }
};

// dynamic distribution: workers take chunks of the batch from an atomic counter
template <class F>
void parallelFor(Oracle* o, int64_t n, int nthreads, F&& fn) {
  nthreads = std::max(1, nthreads);
  if (nthreads == 1 || n < 2) {
    for (int64_t i = 0; i < n; ++i) fn(0, i);
    return;
  }
  if (!o->ensureWorkers(nthreads)) throw std::runtime_error("oracle: clone failed");
  std::atomic<int64_t> next{0};
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(256, n / (nthreads * 8)));
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t) {
    pool.emplace_back([&, t]() {
      for (;;) {
        const int64_t lo = next.fetch_add(chunk);
        if (lo >= n) break;
        const int64_t hi = std::min(n, lo + chunk);
        for (int64_t i = lo; i < hi; ++i) fn(t, i);
      }
    });
  }
  for (auto& th : pool) th.join();
}

inline Mn::Vector3 ld(const float* p, int64_t i) { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
inline void st(float* p, int64_t i, const Mn::Vector3& v) {
  p[3 * i] = v.x(); p[3 * i + 1] = v.y(); p[3 * i + 2] = v.z();
}
inline bool validIsland(PathFinder::Impl& I, int island, bool optional = true) {
  if (optional && island == esp::ID_UNDEFINED) return true;
  return island >= 0 && island < I.islandSystem_->numIslands();
}
inline int findPolyIsland(PathFinder::Impl& I, dtPolyRef ref) {
  auto it = I.islandSystem_->polyToIsland_.find(ref);
  return it == I.islandSystem_->polyToIsland_.end() ? -1 : static_cast<int>(it->second);
}

}  // namespace hbnref

using namespace hbnref;

// =============================== rand / srand ===================================
extern "C" int rand(void) {
  if (tlsRand.mode == 1)
    return static_cast<int>(hbnoracle::hbnRand31(tlsRand.seed, tlsRand.query, tlsRand.draw++));
  static int (*real)(void) = reinterpret_cast<int (*)(void)>(dlsym(RTLD_NEXT, "rand"));
  return real();
}
extern "C" void srand(unsigned int s) {
  static void (*real)(unsigned int) =
      reinterpret_cast<void (*)(unsigned int)>(dlsym(RTLD_NEXT, "srand"));
  real(s);
}

// =============================== C ABI =========================================
extern "C" {

typedef void* ref_pf_t;
#define ORACLE(h) static_cast<Oracle*>(h)

// 1 = this library runs the reference's PathFinder.cpp (the restated library answers 0)
int ref_is_reference_pathfinder() { return 1; }

ref_pf_t ref_create() { return new Oracle(); }
void ref_destroy(ref_pf_t h) { delete ORACLE(h); }

void ref_default_settings(void* out56) {
  static_assert(sizeof(esp::nav::NavMeshSettings) == 56, "NavMeshSettings layout");
  esp::nav::NavMeshSettings s;
  memcpy(out56, &s, sizeof(s));
}

// PathFinder::build(bs, MeshData) computes the bounds from the vertices (PF.cpp:952-975)
int ref_build(ref_pf_t h, const void* settings56, const float* verts, int nverts,
              const int* tris, int ntris) {
  esp::nav::NavMeshSettings s;
  memcpy(static_cast<void*>(&s), settings56, sizeof(s));
  esp::assets::MeshData mesh;
  mesh.vbo.resize(nverts);
  for (int i = 0; i < nverts; ++i) mesh.vbo[i] = ld(verts, i);
  mesh.ibo.assign(tris, tris + static_cast<size_t>(ntris) * 3);
  ORACLE(h)->dropClones();
  return ORACLE(h)->master().build(s, mesh) ? 1 : 0;
}

int ref_build_tiled(ref_pf_t h, const void* settings56, const float* verts, int nverts,
                    const int* tris, int ntris, int tileSize, int nthreads) {
  hbnoracle::NavMeshSettings s;
  memcpy(&s, settings56, sizeof(s));
  std::vector<unsigned char> image;
  if (!hbnoracle::buildTiledImage(s, verts, nverts, tris, ntris, tileSize, nthreads, image))
    return 0;
  return ORACLE(h)->loadImage(image.data(), image.size()) ? 1 : 0;
}

int ref_load_memory(ref_pf_t h, const unsigned char* buf, int64_t len) {
  return ORACLE(h)->loadImage(buf, static_cast<size_t>(len)) ? 1 : 0;
}
int ref_load(ref_pf_t h, const char* path) {
  ORACLE(h)->dropClones();
  return ORACLE(h)->master().loadNavMesh(path) ? 1 : 0;
}
// two-call pattern: returns required size; copies when cap is large enough
int64_t ref_save_memory(ref_pf_t h, unsigned char* out, int64_t cap) {
  std::vector<unsigned char> buf;
  if (!ORACLE(h)->master().isLoaded() || !ORACLE(h)->saveImage(buf)) return -1;
  if (out && cap >= static_cast<int64_t>(buf.size())) memcpy(out, buf.data(), buf.size());
  return static_cast<int64_t>(buf.size());
}
int ref_save(ref_pf_t h, const char* path) { return ORACLE(h)->master().saveNavMesh(path) ? 1 : 0; }

int ref_is_loaded(ref_pf_t h) { return ORACLE(h)->master().isLoaded(); }
int ref_num_islands(ref_pf_t h) { return ORACLE(h)->master().numIslands(); }
float ref_navigable_area(ref_pf_t h, int island) {
  if (!validIsland(ORACLE(h)->impl(), island)) return kNaN;
  return ORACLE(h)->master().getNavigableArea(island);
}
float ref_island_radius(ref_pf_t h, int island) {
  if (!validIsland(ORACLE(h)->impl(), island, false)) return kNaN;
  return ORACLE(h)->master().islandRadius(island);
}
void ref_get_bounds(ref_pf_t h, float* out6) {
  auto b = ORACLE(h)->master().bounds();
  st(out6, 0, b.first);
  st(out6, 1, b.second);
}
void ref_seed(ref_pf_t h, uint32_t s) { ORACLE(h)->master().seed(s); }

// mesh statistics: {tiles, polys, verts, links(maxLinkCount), bvNodes, detailTris,
// detailVerts, tileDataBytes}
void ref_mesh_stats(ref_pf_t h, int64_t* out8) {
  const dtNavMesh* nav = ORACLE(h)->impl().navMesh_.get();
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

// Finalised tile blobs of the LIVE dtNavMesh the reference queries (links connected,
// zero-area polys disabled): what a habitat-sim integration hands to
// hbn_navmesh_create_from_tiles after initNavQuery.
int ref_tile_count(ref_pf_t h) {
  const dtNavMesh* nav = ORACLE(h)->impl().navMesh_.get();
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
  const dtNavMesh* nav = ORACLE(h)->impl().navMesh_.get();
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
  const dtNavMeshParams* p = ORACLE(h)->impl().navMesh_->getParams();
  memcpy(orig3, p->orig, 12);
  tileWH2[0] = p->tileWidth;
  tileWH2[1] = p->tileHeight;
  maxTilesPolys2[0] = p->maxTiles;
  maxTilesPolys2[1] = p->maxPolys;
}

// island id of every poly in (tile table order, poly order), read from the reference's own
// IslandSystem::polyToIsland_; -1 if unmapped
int64_t ref_poly_islands(ref_pf_t h, int32_t* out, uint32_t* outRefs, int64_t cap) {
  auto& I = ORACLE(h)->impl();
  const dtNavMesh* nav = I.navMesh_.get();
  int64_t n = 0;
  for (int i = 0; i < nav->getMaxTiles(); ++i) {
    const dtMeshTile* t = nav->getTile(i);
    if (!t || !t->header) continue;
    for (int j = 0; j < t->header->polyCount; ++j) {
      dtPolyRef ref = nav->encodePolyId(t->salt, i, j);
      if (n < cap) {
        if (out) out[n] = findPolyIsland(I, ref);
        if (outRefs) outRefs[n] = ref;
      }
      n++;
    }
  }
  return n;
}

// ---- batched queries ----------------------------------------------------------
// snap_point (PF.cpp:1725-1757) + the ref/island the same projectToPoly yields
void ref_snap_batch(ref_pf_t h, const float* pts, int64_t n, float* out_pts,
                    uint32_t* out_refs, int32_t* out_island, int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    PathFinder& pf = *o->pfs[t];
    auto& I = o->impl(t);
    const Mn::Vector3 p = ld(pts, i);
    if (out_pts) st(out_pts, i, pf.snapPoint<Mn::Vector3>(p));
    if (out_refs || out_island) {
      dtStatus status;
      dtPolyRef ref;
      Mn::Vector3 q;
      std::tie(status, ref, q) = esp::nav::projectToPoly(p, I.navQuery_.get(), I.filter_.get());
      const bool ok = dtStatusSucceed(status);
      if (out_refs) out_refs[i] = ok ? ref : 0;
      if (out_island) out_island[i] = ok ? pf.getIsland<Mn::Vector3>(p) : -1;
    }
  });
}

// island-restricted snap (PF.cpp:1725-1757: mutates poly flags around the search)
void ref_snap_island_batch(ref_pf_t h, const float* pts, const int32_t* islands, int64_t n,
                           float* out_pts, uint32_t* out_refs) {
  Oracle* o = ORACLE(h);
  PathFinder& pf = o->master();
  auto& I = o->impl();
  for (int64_t i = 0; i < n; ++i) {
    const int island = islands ? islands[i] : -1;
    const Mn::Vector3 p = pf.snapPoint<Mn::Vector3>(ld(pts, i), island);
    st(out_pts, i, p);
    if (out_refs) {
      // the ref of the poly the restricted search chose: the same search run again under the
      // reference's own flag protocol (PF.cpp:1729-1751)
      out_refs[i] = 0;
      if (!std::isnan(p.x())) {
        if (island != esp::ID_UNDEFINED) {
          I.islandSystem_->setPolyFlagForIsland(I.navMesh_.get(), esp::nav::POLYFLAGS_OFF_ISLAND,
                                                island, true, true);
          I.filter_->setExcludeFlags(I.filter_->getExcludeFlags() |
                                     esp::nav::POLYFLAGS_OFF_ISLAND);
        }
        dtStatus status;
        dtPolyRef ref;
        Mn::Vector3 q;
        std::tie(status, ref, q) =
            esp::nav::projectToPoly(ld(pts, i), I.navQuery_.get(), I.filter_.get());
        if (island != esp::ID_UNDEFINED) {
          I.islandSystem_->setPolyFlagForIsland(I.navMesh_.get(), esp::nav::POLYFLAGS_OFF_ISLAND,
                                                island, false, true);
          I.filter_->setExcludeFlags(I.filter_->getExcludeFlags() &
                                     ~esp::nav::POLYFLAGS_OFF_ISLAND);
        }
        if (dtStatusSucceed(status)) out_refs[i] = ref;
      }
    }
  }
}

void ref_is_navigable_batch(ref_pf_t h, const float* pts, int64_t n, float maxYDelta,
                            uint8_t* out, int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    out[i] = o->pfs[t]->isNavigable(ld(pts, i), maxYDelta) ? 1 : 0;
  });
}

void ref_island_radius_batch(ref_pf_t h, const float* pts, int64_t n, float* out, int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads,
              [&](int t, int64_t i) { out[i] = o->pfs[t]->islandRadius(ld(pts, i)); });
}

// find_path(ShortestPath).  out_dist inf / out_npts 0 on failure.  out_pts is
// [n, max_pts, 3] (first min(npts,max_pts) points written) or null.
void ref_find_path_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                         float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                         int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    esp::nav::ShortestPath path;
    path.requestedStart = ld(starts, i);
    path.requestedEnd = ld(ends, i);
    o->pfs[t]->findPath(path);
    out_dist[i] = path.geodesicDistance;
    if (out_npts) out_npts[i] = static_cast<int32_t>(path.points.size());
    if (out_pts)
      for (int k = 0; k < static_cast<int>(path.points.size()) && k < max_pts; ++k)
        st(out_pts, i * max_pts + k, path.points[k]);
  });
}

// Raw view of one find_path.  Distance and points are PathFinder::findPath's; the refs,
// corridor, Detour status words and node-pool use come from the same calls PF.cpp:1426-1468
// makes, issued on the reference's own dtNavMeshQuery.  out_corridor is [n, 256]; out_info is
// [n, 8] = {startRef, endRef, astarStatus, straightStatus, numPolys, numPoints, nodesUsed,
// flags(bit0 trivial, bit1 connected, bit2 found)}.
void ref_find_path_raw_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                             float* out_dist, uint32_t* out_corridor, uint32_t* out_info,
                             float* out_pts, int max_pts, int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    auto& I = o->impl(t);
    dtNavMeshQuery* q = I.navQuery_.get();
    uint32_t* info = out_info + i * 8;
    memset(info, 0, 32);
    const Mn::Vector3 start = ld(starts, i), end = ld(ends, i);
    esp::nav::ShortestPath path;
    path.requestedStart = start;
    path.requestedEnd = end;
    const bool found = o->pfs[t]->findPath(path);
    out_dist[i] = found ? path.geodesicDistance : kInf;
    if (found && out_pts)
      for (int k = 0; k < static_cast<int>(path.points.size()) && k < max_pts; ++k)
        st(out_pts, i * max_pts + k, path.points[k]);
    if (found) {
      info[5] = static_cast<uint32_t>(path.points.size());
      info[7] |= 4u;
    }
    dtStatus s0, s1;
    dtPolyRef startRef = 0, endRef = 0;
    Mn::Vector3 pathStart, pathEnd;
    std::tie(s0, startRef, pathStart) = esp::nav::projectToPoly(start, q, I.filter_.get());
    if (s0 != DT_SUCCESS || startRef == 0) return;
    info[0] = startRef;
    std::tie(s1, endRef, pathEnd) = esp::nav::projectToPoly(end, q, I.filter_.get());
    if (s1 != DT_SUCCESS || endRef == 0) return;
    info[1] = endRef;
    if (pathStart == pathEnd) {  // PF.cpp:1434-1436 (Magnum's fuzzy ==)
      info[7] |= 1u;
      return;
    }
    if (!I.islandSystem_->hasConnection(startRef, endRef)) return;
    info[7] |= 2u;
    dtPolyRef polys[256];
    int numPolys = 0;
    dtStatus status = q->findPath(startRef, endRef, pathStart.data(), pathEnd.data(),
                                  I.filter_.get(), polys, &numPolys, 256);
    info[2] = status;
    info[4] = numPolys;
    // same poly: dtNavMeshQuery::findPath returns before it clears the pool (stale count)
    info[6] = startRef == endRef ? 0 : q->getNodePool()->getNodeCount();
    if (out_corridor)
      for (int k = 0; k < numPolys && k < 256; ++k) out_corridor[i * 256 + k] = polys[k];
    if (status != DT_SUCCESS || numPolys == 0) return;
    int numPoints = 0;
    std::vector<Mn::Vector3> points(256);
    status = q->findStraightPath(start.data(), end.data(), polys, numPolys, points[0].data(),
                                 nullptr, nullptr, &numPoints, 256);
    info[3] = status;
    if (!found) info[5] = numPoints;
  });
}

// Work counters of one find_path for the roofline's ALGORITHMIC bytes:
// out_stats is [n, 8] = {expanded polys (CLOSED nodes left in the pool after findPath),
// links of the expanded polys, their non-null neighbours, corridor polys, links of the
// corridor polys, straight-path points, nodes allocated, nodes still open}.
void ref_find_path_stats_batch(ref_pf_t h, const float* starts, const float* ends, int64_t n,
                               uint32_t* out_stats, int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    auto& I = o->impl(t);
    dtNavMeshQuery* q = I.navQuery_.get();
    const dtNavMesh* nav = I.navMesh_.get();
    uint32_t* st6 = out_stats + i * 8;
    memset(st6, 0, 32);
    dtStatus s0, s1;
    dtPolyRef startRef = 0, endRef = 0;
    Mn::Vector3 pathStart, pathEnd;
    const Mn::Vector3 start = ld(starts, i), end = ld(ends, i);
    std::tie(s0, startRef, pathStart) = esp::nav::projectToPoly(start, q, I.filter_.get());
    if (s0 != DT_SUCCESS || startRef == 0) return;
    std::tie(s1, endRef, pathEnd) = esp::nav::projectToPoly(end, q, I.filter_.get());
    if (s1 != DT_SUCCESS || endRef == 0) return;
    if (pathStart == pathEnd) return;
    if (!I.islandSystem_->hasConnection(startRef, endRef)) return;
    dtPolyRef polys[256];
    int numPolys = 0;
    dtStatus status = q->findPath(startRef, endRef, pathStart.data(), pathEnd.data(),
                                  I.filter_.get(), polys, &numPolys, 256);
    auto linkStats = [&](dtPolyRef ref, uint32_t& links, uint32_t& neis) {
      const dtMeshTile* tile = nullptr;
      const dtPoly* poly = nullptr;
      if (dtStatusFailed(nav->getTileAndPolyByRef(ref, &tile, &poly))) return;
      for (unsigned int k = poly->firstLink; k != DT_NULL_LINK; k = tile->links[k].next) {
        links++;
        if (tile->links[k].ref) neis++;
      }
    };
    if (startRef != endRef) {
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
    st6[3] = numPolys;
    uint32_t dummy = 0;
    for (int k = 0; k < numPolys; ++k) linkStats(polys[k], st6[4], dummy);
    if (status == DT_SUCCESS && numPolys) {
      int numPoints = 0;
      std::vector<Mn::Vector3> points(256);
      status = q->findStraightPath(start.data(), end.data(), polys, numPolys, points[0].data(),
                                   nullptr, nullptr, &numPoints, 256);
      if (status == DT_SUCCESS) st6[5] = numPoints;
    }
  });
}

// find_path(MultiGoalShortestPath), fresh object per start (no cache).
// ends is [n, g, 3].  Outputs: dist[n], idx[n], npts[n], pts [n,max_pts,3]|null.
void ref_find_path_multigoal_batch(ref_pf_t h, const float* starts, const float* ends,
                                   int64_t n, int g, float* out_dist, int32_t* out_idx,
                                   int32_t* out_npts, float* out_pts, int max_pts,
                                   int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    esp::nav::MultiGoalShortestPath path;
    path.requestedStart = ld(starts, i);
    std::vector<Mn::Vector3> e(g);
    for (int k = 0; k < g; ++k) e[k] = ld(ends, i * g + k);
    path.setRequestedEnds(e);
    o->pfs[t]->findPath(path);
    out_dist[i] = path.geodesicDistance;
    out_idx[i] = path.closestEndPointIndex;
    if (out_npts) out_npts[i] = static_cast<int32_t>(path.points.size());
    if (out_pts)
      for (int k = 0; k < static_cast<int>(path.points.size()) && k < max_pts; ++k)
        st(out_pts, i * max_pts + k, path.points[k]);
  });
}

// Stateful multi-goal object: the reference's own MultiGoalShortestPath, reused across calls
// (src/tests/PathFinderTest.cpp:136-164 and trap T4).
void* ref_multigoal_create() {
  ensureLogging();
  return new esp::nav::MultiGoalShortestPath();
}
void ref_multigoal_destroy(void* m) { delete static_cast<esp::nav::MultiGoalShortestPath*>(m); }
void ref_multigoal_set_ends(void* m, const float* ends, int g) {
  std::vector<Mn::Vector3> e(g);
  for (int k = 0; k < g; ++k) e[k] = ld(ends, k);
  static_cast<esp::nav::MultiGoalShortestPath*>(m)->setRequestedEnds(e);
}
int ref_multigoal_find(ref_pf_t h, void* m, const float* start, float* out_dist,
                       int32_t* out_idx, int32_t* out_npts, float* out_pts, int max_pts) {
  auto* path = static_cast<esp::nav::MultiGoalShortestPath*>(m);
  path->requestedStart = ld(start, 0);
  const bool ok = ORACLE(h)->master().findPath(*path);
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
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    PathFinder& pf = *o->pfs[t];
    st(out, i, allowSliding ? pf.tryStep<Mn::Vector3>(ld(starts, i), ld(ends, i))
                            : pf.tryStepNoSliding<Mn::Vector3>(ld(starts, i), ld(ends, i)));
  });
}

// closest_obstacle_surface_point: out is [n, 7] = hitPos, hitNormal, hitDist
void ref_obstacle_batch(ref_pf_t h, const float* pts, int64_t n, float maxRadius, float* out,
                        int nthreads) {
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    esp::nav::HitRecord r = o->pfs[t]->closestObstacleSurfacePoint(ld(pts, i), maxRadius);
    float* w = out + 7 * i;
    w[0] = r.hitPos.x(); w[1] = r.hitPos.y(); w[2] = r.hitPos.z();
    w[3] = r.hitNormal.x(); w[4] = r.hitNormal.y(); w[5] = r.hitNormal.z();
    w[6] = r.hitDist;
  });
}

// get_random_navigable_point (PF.cpp:1236-1281).  mode 0: glibc rand() stream (sequential).
// mode 1: rand() returns the counter-based stream hbn_uniform(seed, query0+i, draw).
// islands may be null (-1 for all).  Returns 0 where the reference throws / asserts.
int ref_random_points(ref_pf_t h, int64_t n, int maxTries, const int32_t* islands, int mode,
                      uint64_t seed, uint64_t query0, float* out_pts, uint32_t* out_refs) {
  Oracle* o = ORACLE(h);
  auto& I = o->impl();
  int ok = 1;
  // scoped redirection: the ESP_ERROR line per failed sample ("Failed to getRandomNavigablePoint")
  // is expected output here, not a finding
  Corrade::Utility::Error quiet{nullptr};
  for (int64_t i = 0; i < n && ok; ++i) {
    const int island = islands ? islands[i] : -1;
    if (!validIsland(I, island)) { ok = 0; break; }  // CORRADE_ASSERT (PF.cpp:225-232)
    tlsRand = RandStream{mode, seed, query0 + static_cast<uint64_t>(i), 0};
    try {
      const Mn::Vector3 p = o->master().getRandomNavigablePoint(maxTries, island);
      st(out_pts, i, p);
      if (out_refs) {
        // the poly the sample fell on: the sample is a point of that poly, so the
        // island-restricted nearest-poly search returns it (distance 0)
        out_refs[i] = 0;
        if (!std::isnan(p.x())) {
          dtStatus status;
          dtPolyRef ref;
          Mn::Vector3 q;
          std::tie(status, ref, q) = esp::nav::projectToPoly(p, I.navQuery_.get(), I.filter_.get());
          if (dtStatusSucceed(status)) out_refs[i] = ref;
        }
      }
    } catch (const std::runtime_error&) {
      ok = 0;  // ESP_CHECK: navigable area <= 0 (PF.cpp:1240-1243)
    }
  }
  tlsRand = RandStream{};
  return ok;
}

// get_random_navigable_point_near (getRandomNavigablePointInCircle, PF.cpp:1283-1332)
int ref_random_points_near(ref_pf_t h, int64_t n, const float* centers, float radius,
                           int maxTries, const int32_t* islands, int mode, uint64_t seed,
                           uint64_t query0, float* out_pts) {
  Oracle* o = ORACLE(h);
  auto& I = o->impl();
  int ok = 1;
  // scoped redirection: the ESP_ERROR line per failed sample ("Failed to getRandomNavigablePoint")
  // is expected output here, not a finding
  Corrade::Utility::Error quiet{nullptr};
  for (int64_t i = 0; i < n && ok; ++i) {
    const int island = islands ? islands[i] : -1;
    if (!validIsland(I, island)) { ok = 0; break; }
    tlsRand = RandStream{mode, seed, query0 + static_cast<uint64_t>(i), 0};
    try {
      st(out_pts, i,
         o->master().getRandomNavigablePointAroundSphere(ld(centers, i), radius, maxTries, island));
    } catch (const std::runtime_error&) {
      ok = 0;
    }
  }
  tlsRand = RandStream{};
  return ok;
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
  return hbnoracle::hbnUniform(seed, query, draw);
}
// the uniform the UNMODIFIED frand() (PF.cpp:1232-1234) makes of the interposed rand()
float ref_frand_of_stream(uint64_t seed, uint64_t query, uint32_t draw) {
  tlsRand = RandStream{1, seed, query, draw};
  const float u = esp::nav::frand();
  tlsRand = RandStream{};
  return u;
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
  Oracle* o = ORACLE(h);
  parallelFor(o, n, nthreads, [&](int t, int64_t i) {
    auto& I = o->impl(t);
    dtNavMeshQuery* q = I.navQuery_.get();
    dtStatus s0;
    dtPolyRef startRef;
    Mn::Vector3 pathStart;
    std::tie(s0, startRef, pathStart) = esp::nav::projectToPoly(ld(starts, i), q, I.filter_.get());
    out_nvisited[i] = 0;
    st(out_pos, i, Mn::Vector3(kNaN, kNaN, kNaN));
    if (dtStatusFailed(s0) || !startRef) return;
    dtPolyRef polys[256];
    int np = 0;
    Mn::Vector3 res;
    const Mn::Vector3 e = ld(ends, i);
    q->moveAlongSurface(startRef, pathStart.data(), e.data(), I.filter_.get(), res.data(), polys,
                        &np, 256);
    st(out_pos, i, res);
    out_nvisited[i] = np;
    for (int k = 0; k < np && k < 16; ++k) out_visited[i * 16 + k] = polys[k];
  });
}

// ---- entry points only the real PathFinder has ------------------------------------
// get_topdown_view / get_topdown_island_view (PF.cpp:1833-1896).  Two-call pattern: returns
// rows * cols, dims through out_dims[2] = {rows, cols}; copies when cap suffices.
int64_t ref_topdown_view(ref_pf_t h, float mpp, float height, float eps, int island_view,
                         int32_t* out, int64_t cap, int32_t* out_dims) {
  PathFinder& pf = ORACLE(h)->master();
  if (island_view) {
    auto g = pf.getTopDownIslandView(mpp, height, eps);
    out_dims[0] = static_cast<int32_t>(g.rows());
    out_dims[1] = static_cast<int32_t>(g.cols());
    const int64_t n = static_cast<int64_t>(g.rows()) * g.cols();
    if (out && cap >= n)
      for (int64_t r = 0; r < g.rows(); ++r)
        for (int64_t c = 0; c < g.cols(); ++c) out[r * g.cols() + c] = g(r, c);
    return n;
  }
  auto g = pf.getTopDownView(mpp, height, eps);
  out_dims[0] = static_cast<int32_t>(g.rows());
  out_dims[1] = static_cast<int32_t>(g.cols());
  const int64_t n = static_cast<int64_t>(g.rows()) * g.cols();
  if (out && cap >= n)
    for (int64_t r = 0; r < g.rows(); ++r)
      for (int64_t c = 0; c < g.cols(); ++c) out[r * g.cols() + c] = g(r, c) ? 1 : 0;
  return n;
}

// build_navmesh_vertices / indices (getNavMeshData, PF.cpp:1898-1944).  Two-call pattern:
// returns the vertex count; copies when cap (in vertices) suffices.  -1 for an invalid island.
int64_t ref_navmesh_vertices(ref_pf_t h, int island, float* out_verts, uint32_t* out_indices,
                             int64_t cap) {
  if (!validIsland(ORACLE(h)->impl(), island)) return -1;
  auto md = ORACLE(h)->master().getNavMeshData(island);
  const int64_t n = static_cast<int64_t>(md->vbo.size());
  if (cap >= n) {
    if (out_verts)
      for (int64_t i = 0; i < n; ++i) st(out_verts, i, md->vbo[i]);
    if (out_indices)
      for (int64_t i = 0; i < n; ++i) out_indices[i] = md->ibo[i];
  }
  return n;
}

}  // extern "C"

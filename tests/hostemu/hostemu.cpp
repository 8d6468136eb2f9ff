// tests/hostemu/hostemu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the device query code of habitat-sim_b200/csrc/hbn_query.h for the host with a
// one-lane group (HostGroup) so that the kernels' logic can be debugged against the oracle
// on a box without a GPU.  It is never loaded by the product: libhbn.so contains only the
// CUDA path and fails when no device is present.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "hbn_host.h"
#include "hbn_query.h"
#include "hbn_astar_lane.h"
#include "hbn_snap.h"

using namespace hbn;

struct Emu {
  HostNavMesh mesh;
  FlatNav flat;
  NavView nav;
  std::string err;
};

extern "C" {

void* emu_create(const unsigned char* buf, long len) {
  Emu* e = new Emu();
  if (!e->mesh.loadMSET(buf, static_cast<size_t>(len), e->err)) {
    fprintf(stderr, "emu_create: %s\n", e->err.c_str());
    delete e;
    return nullptr;
  }
  e->mesh.finish(nullptr);
  e->mesh.flatten(e->flat);
  e->nav = e->flat.view();
  return e;
}
void emu_destroy(void* h) { delete static_cast<Emu*>(h); }

// {numTilesPresent, numPolys, numLinks, numIslands}
void emu_info(void* h, long* out4) {
  Emu* e = static_cast<Emu*>(h);
  long nt = 0;
  for (auto& t : e->mesh.tiles()) nt += t.present ? 1 : 0;
  out4[0] = nt;
  out4[1] = static_cast<long>(e->flat.polys.size());
  out4[2] = static_cast<long>(e->flat.links.size());
  out4[3] = static_cast<long>(e->flat.islandRadius.size());
}
float emu_island_radius(void* h, int i) { return static_cast<Emu*>(h)->flat.islandRadius[i]; }
float emu_island_area(void* h, int i) {
  Emu* e = static_cast<Emu*>(h);
  return i < 0 ? e->flat.totalArea : e->flat.islandArea[i];
}
void emu_bounds(void* h, float* out6) { memcpy(out6, static_cast<Emu*>(h)->flat.bounds, 24); }

// finalised blob of the idx-th present tile (table order)
int emu_tile_blob(void* h, int idx, unsigned char* out, int cap) {
  Emu* e = static_cast<Emu*>(h);
  int n = 0;
  for (auto& t : e->mesh.tiles()) {
    if (!t.present) continue;
    if (n++ != idx) continue;
    if (out && cap >= static_cast<int>(t.data.size())) memcpy(out, t.data.data(), t.data.size());
    return static_cast<int>(t.data.size());
  }
  return -1;
}
long emu_poly_islands(void* h, int* out, unsigned* refs, long cap) {
  Emu* e = static_cast<Emu*>(h);
  const long n = static_cast<long>(e->flat.polys.size());
  for (long i = 0; i < n && i < cap; ++i) {
    if (out) out[i] = e->flat.polys[i].island;
    if (refs) refs[i] = e->flat.polys[i].ref;
  }
  return n;
}

static const float kExt[3] = {2.f, 4.f, 2.f};  // polyPickExt, PathFinder.cpp:134

void emu_snap(void* h, const float* pts, const int* islands, long n, float* out_pts,
              unsigned* out_refs, int* out_isl) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  for (long i = 0; i < n; ++i) {
    // as k_snap: the narrowed box of snapRadius (emu_find_path & co. keep the reference's box)
    const int isl = islands ? islands[i] : -1;
    const float rxz = snapRadius(e->nav, pts + 3 * i, kExt, isl);
    const Nearest r = findNearestPoly(e->nav, grp, pts + 3 * i, kExt, isl, q, rxz);
    const bool ok = r.g != kNoPoly;
    for (int k = 0; k < 3; ++k) out_pts[3 * i + k] = ok ? r.pt[k] : NAN;
    if (out_refs) out_refs[i] = ok ? e->nav.polys[r.g].ref : 0u;
    if (out_isl) out_isl[i] = ok ? e->nav.polys[r.g].island : -1;
  }
}

static int g_snapOneWalk = 1;  // emu_snap_one_walk: the device's one-walk path (default) or the two-walk form
static long g_snapSecondWalks = 0;
// returns the number of points that needed the wider second walk since the last call
long emu_snap_one_walk(int on) {
  g_snapOneWalk = on;
  const long r = g_snapSecondWalks;
  g_snapSecondWalks = 0;
  return r;
}

// The candidate-list findNearestPoly (hbn_snap.h / hbn_snap.cuh: walk -> eval -> mark -> select), serially.
void emu_snap_list(void* h, const float* pts, const int* islands, long n, float* out_pts,
                   unsigned* out_refs, int* out_isl, long* out_ncand) {
  Emu* e = static_cast<Emu*>(h);
  std::vector<uint32_t> cand;
  std::vector<float> lb, d;
  long total = 0, evaluated = 0;
  for (long i = 0; i < n; ++i) {
    cand.clear();
    lb.clear();
    const int isl = islands ? islands[i] : -1;
    if (g_snapOneWalk && e->nav.bvXzTight) {
      // as k_snap_walk: the column walk with the minimum radius keeps its candidates; a wider walk only if the
      // radius it arrives at is larger (or it found more than a block of candidates)
      float ub = kFltMax;
      snapWalk(e->nav, pts + 3 * i, kExt, kSnapRadiusMin, [&](uint32_t g, float b) {
        snapUbUpdate(e->nav, pts + 3 * i, isl, g, b, ub);
        cand.push_back(g);
        lb.push_back(b);
      });
      const float r = snapRadiusFromUb(ub, kExt[0]);
      if (cand.size() > 8 || r > kSnapRadiusMin) {
        cand.clear();
        lb.clear();
        snapWalk(e->nav, pts + 3 * i, kExt, r, [&](uint32_t g, float b) { cand.push_back(g); lb.push_back(b); });
        g_snapSecondWalks++;
      }
    } else {
      const float r = snapRadius(e->nav, pts + 3 * i, kExt, isl);
      snapWalk(e->nav, pts + 3 * i, kExt, r, [&](uint32_t g, float b) { cand.push_back(g); lb.push_back(b); });
    }
    total += static_cast<long>(cand.size());
    d.assign(cand.size(), -1.f);
    SnapCandOut o;
    // as k_snap_eval / k_snap_mark: the minimum of {distance bits << 32 | visit index} over both passes
    unsigned long long best = kSnapBestInit;
    for (int pass = 0; pass < 2; ++pass)
      for (size_t c = 0; c < cand.size(); ++c) {
        if ((lb[c] == 0.f) != (pass == 0)) continue;
        float bd;
        const uint32_t bb = static_cast<uint32_t>(best >> 32);
        memcpy(&bd, &bb, 4);
        if (pass == 1 && !snapMayWin(lb[c], bd)) continue;
        d[c] = snapEval(e->nav, pts + 3 * i, isl, cand[c], &o);
        evaluated++;
        if (d[c] >= 0.f && d[c] < kFltMax) {
          uint32_t db;
          memcpy(&db, &d[c], 4);
          const unsigned long long key = (static_cast<unsigned long long>(db) << 32) | c;
          if (key < best) best = key;
        }
        // the bound must never exceed the distance it bounds
        if (d[c] >= 0.f && lb[c] * 0.999f - 1e-6f > d[c]) { fprintf(stderr, "snap lower bound violated: pt %ld cand %zu g %u lb %g d %g over %u center %g %g %g cp %g %g %g\n", i, c, cand[c], lb[c], d[c], o.over, pts[3*i], pts[3*i+1], pts[3*i+2], o.cp[0], o.cp[1], o.cp[2]); abort(); }
      }
    const bool ok = best != kSnapBestInit;
    const uint32_t w = ok ? static_cast<uint32_t>(best & 0xffffffffu) : 0u;
    if (ok && w != snapSelect(d.data(), 0, static_cast<uint32_t>(cand.size()))) { fprintf(stderr, "snap: the packed minimum is not the first strict minimum (pt %ld)\n", i); abort(); }
    if (ok) snapEval(e->nav, pts + 3 * i, isl, cand[w], &o);
    for (int k = 0; k < 3; ++k) out_pts[3 * i + k] = ok ? o.cp[k] : NAN;
    if (out_refs) out_refs[i] = ok ? e->nav.polys[cand[w]].ref : 0u;
    if (out_isl) out_isl[i] = ok ? e->nav.polys[cand[w]].island : -1;
  }
  if (out_ncand) { out_ncand[0] = total; out_ncand[1] = evaluated; }
}

// out_info [n,8] like the oracle's ref_find_path_raw_batch
void emu_find_path(void* h, const float* starts, const float* ends, long n, int cap, int fastFail,
                   float* out_dist, int* out_npts, float* out_pts, int max_pts,
                   unsigned* out_corridor, unsigned* out_info, int* out_overflow) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  std::vector<char> buf(astarWsBytes(cap) + 64);
  void* aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~uintptr_t(15));
  AStarWs w = astarWsCarve(aligned, cap);
  for (long i = 0; i < n; ++i) {
    const Nearest s = findNearestPoly(e->nav, grp, starts + 3 * i, kExt, -1, q);
    const Nearest t = findNearestPoly(e->nav, grp, ends + 3 * i, kExt, -1, q);
    memset(w.hash, 0, sizeof(uint32_t) * 2 * cap);
    PathResult r = findPathInternal(e->nav, w, starts + 3 * i, ends + 3 * i, s.g, s.pt, t.g, t.pt,
                                    fastFail != 0, out_pts ? out_pts + static_cast<size_t>(i) * max_pts * 3 : nullptr,
                                    max_pts, out_corridor ? out_corridor + i * 256 : nullptr);
    out_dist[i] = r.dist;
    if (out_npts) out_npts[i] = (r.flags & 4u) ? r.npts : 0;
    if (out_overflow) out_overflow[i] = r.overflow ? 1 : 0;
    if (out_info) {
      unsigned* o = out_info + i * 8;
      o[0] = s.g != kNoPoly ? e->nav.polys[s.g].ref : 0;
      o[1] = (s.g != kNoPoly && t.g != kNoPoly) ? e->nav.polys[t.g].ref : 0;
      o[2] = r.astarStatus; o[3] = r.straightStatus; o[4] = r.ncorridor; o[5] = r.npts;
      o[6] = r.nodesUsed; o[7] = r.flags;
    }
  }
}

// Node-key numbering of the flattener (PolyRec::key0, LinkRec::neiKey): out[0] = numKeys,
// out[1] = 1 if the per-poly key ranges tile [0, numKeys) without overlap and every link's key
// lies in its neighbour's range, out[2] = links counted, out[3] = sum over links of
// |key0(poly) - neiKey| / 16 (distance in 32 B sectors of the node table: the locality the
// numbering is chosen for).
void emu_key_stats(void* h, long* out) {
  Emu* e = static_cast<Emu*>(h);
  const FlatNav& f = e->flat;
  const size_t np = f.polys.size();
  std::vector<uint8_t> sides(np, 1);
  for (const LinkRec& lr : f.links)
    if (lr.nei != kNoPoly) sides[lr.nei] |= static_cast<uint8_t>(1u << ((lr.meta >> kLinkStateShift) & 3u));
  std::vector<uint8_t> seen(f.numKeys, 0);
  bool ok = true;
  for (size_t g = 0; g < np; ++g) {
    const int cnt = __builtin_popcount(sides[g]);
    for (int k = 0; k < cnt; ++k) {
      const uint32_t key = f.polys[g].key0 + static_cast<uint32_t>(k);
      if (key >= f.numKeys || seen[key]) ok = false; else seen[key] = 1;
    }
  }
  for (uint32_t k = 0; k < f.numKeys; ++k) if (!seen[k]) ok = false;
  long links = 0, dist = 0;
  for (size_t g = 0; g < np; ++g)
    for (uint32_t l = f.polys[g].linkStart; l < f.polys[g].linkStart + f.polys[g].linkCount; ++l) {
      const LinkRec& lr = f.links[l];
      if (lr.nei == kNoPoly) continue;
      const uint32_t k0 = f.polys[lr.nei].key0;
      if (lr.neiKey < k0 || lr.neiKey >= k0 + static_cast<uint32_t>(__builtin_popcount(sides[lr.nei]))) ok = false;
      links++;
      dist += labs(static_cast<long>(f.polys[g].key0 / 16) - static_cast<long>(lr.neiKey / 16));
    }
  out[0] = f.numKeys; out[1] = ok ? 1 : 0; out[2] = links; out[3] = dist;
}

// The lane-per-query search (hbn_astar_lane.h, the state machine k_astar_lane runs in every
// lane) on ONE lane slot reused by all n queries, so table generations wrap and get wiped.
// out_info [n,4]: {findPath status (0 = no search), corridor length, nodes allocated, event};
// out_corridor [n,256] poly refs of the (possibly truncated) corridor.
}  // extern "C"

template <int TS, int V, int F = 0>
static void laneSearchRun(void* h, const float* starts, const float* ends, long n, int fastFail, int allCorridors,
                          unsigned* out_corridor, unsigned* out_info) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  const NavView& nav = e->nav;
  std::vector<char> scratch(laneScratchBytes(nav.numKeys) + 64, 0);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch.data()) + 15) & ~uintptr_t(15));
  std::vector<float> K(TS);
  std::vector<uint16_t> S(TS);
  std::vector<uint32_t> ring(kMaxPathPolys);
  LaneSearch<1, TS, 4, V, F> s{};  // 4 links per load stage, as shipped
  s.K = K.data(); s.S = S.data();
  s.tab = reinterpret_cast<uint16_t*>(base);
  s.rec = base + laneTabBytes(nav.numKeys);
  s.G = reinterpret_cast<LaneHeapEnt*>(s.rec + kLaneRecBytes);
  s.gen = 0;
  s.mode = kLIdle;
  for (long i = 0; i < n; ++i) {
    unsigned* o = out_info + i * 4;
    o[0] = o[1] = o[2] = o[3] = 0;
    const Nearest a = findNearestPoly(nav, grp, starts + 3 * i, kExt, -1, q);
    const Nearest b = findNearestPoly(nav, grp, ends + 3 * i, kExt, -1, q);
    if (a.g == kNoPoly || b.g == kNoPoly || vfuzzyEq(a.pt, b.pt)) continue;
    const int32_t si = nav.polys[a.g].island, ei = nav.polys[b.g].island;
    if (si < 0 || si != ei || a.g == b.g || !vfinite(a.pt) || !vfinite(b.pt)) continue;  // k_fp_classify
    if (s.gen >= kLaneGenMax) {
      memset(s.tab, 0, laneTabBytes(nav.numKeys));
      s.gen = 0;
    }
    s.begin(nav, static_cast<uint32_t>(i), a.g, a.pt, b.g, b.pt, ring.data());
    int ev = kLEvNone;
    while (ev == kLEvNone) ev = s.step(nav, fastFail != 0, allCorridors != 0);
    o[3] = static_cast<unsigned>(ev);
    if (o[3] != kLEvFinished) continue;
    o[0] = s.status;
    o[1] = static_cast<unsigned>(s.xk);
    o[2] = static_cast<unsigned>(s.nodeCount);
    const int len = s.xk < kMaxPathPolys ? s.xk : kMaxPathPolys;
    const uint32_t first = static_cast<uint32_t>((kMaxPathPolys - s.xk) & (kMaxPathPolys - 1));
    for (int k = 0; k < len; ++k) {
      const uint32_t via = ring[(first + k) & (kMaxPathPolys - 1)];
      out_corridor[i * kMaxPathPolys + k] = nav.polys[k == 0 ? a.g : nav.links[via].nei].ref;
    }
  }
}

// dtNodeQueue restated plainly (DNode.cpp:156-200, DNode.h:118-142) over (key, node) pairs: the
// yardstick of the heap fuzz below.
struct PlainQueue {
  std::vector<float> k;
  std::vector<uint32_t> s;
  void bubbleUp(int i, float key, uint32_t node) {
    int parent = (i - 1) / 2;
    while (i > 0 && k[parent] > key) {
      k[i] = k[parent]; s[i] = s[parent];
      i = parent;
      parent = (i - 1) / 2;
    }
    k[i] = key; s[i] = node;
  }
  void trickleDown(int i, float key, uint32_t node) {
    const int size = static_cast<int>(k.size());
    int child = i * 2 + 1;
    while (child < size) {
      if (child + 1 < size && k[child] > k[child + 1]) child++;
      k[i] = k[child]; s[i] = s[child];
      i = child;
      child = i * 2 + 1;
    }
    bubbleUp(i, key, node);
  }
  void push(float key, uint32_t node) {
    k.push_back(0.f); s.push_back(0u);
    bubbleUp(static_cast<int>(k.size()) - 1, key, node);
  }
  uint32_t pop() {
    const uint32_t r = s[0];
    const float lk = k.back();
    const uint32_t ls = s.back();
    k.pop_back(); s.pop_back();
    if (!k.empty()) trickleDown(0, lk, ls);
    return r;
  }
  void modify(uint32_t node, float key) {
    for (size_t i = 0; i < s.size(); ++i)
      if (s[i] == node) { bubbleUp(static_cast<int>(i), key, node); return; }
  }
};

template <int TS, int V>
static long laneHeapFuzz(unsigned seed, long ops, int keyLevels) {
  std::vector<float> K(TS);
  std::vector<uint16_t> S(TS);
  std::vector<LaneHeapEnt> G(kMaxNodes + 2 + 8);
  LaneSearch<1, TS, 4, V> h{};
  h.K = K.data(); h.S = S.data(); h.G = G.data();
  h.size = 0;
  PlainQueue ref;
  std::vector<float> keyOf(kMaxNodes, 0.f);
  std::vector<char> open(kMaxNodes, 0);
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return static_cast<uint32_t>(st >> 33); };
  uint32_t nextNode = 0;
  for (long op = 0; op < ops; ++op) {
    const uint32_t r = rnd() % 100;
    // grow towards the pool size, then mix; keys from a few levels only: ties everywhere
    const bool canPush = nextNode < static_cast<uint32_t>(kMaxNodes);
    if ((r < 55 && canPush) || h.size == 0) {
      if (!canPush) break;
      const float key = static_cast<float>(rnd() % static_cast<uint32_t>(keyLevels)) * 0.25f;
      const uint32_t node = nextNode++;
      keyOf[node] = key; open[node] = 1;
      ref.push(key, node);
      h.heapUp(h.size, key, node);
      h.size++;
    } else if (r < 85) {
      const uint32_t a = ref.pop();
      const uint32_t b = h.S[0];
      h.size--;
      h.heapPopSift(h.size);
      open[a] = 0;
      if (a != b) return op + 1;
    } else {  // decrease-key of a random open node
      uint32_t node = rnd() % nextNode;
      uint32_t tries = 0;
      while (!open[node] && tries++ < 64) node = rnd() % nextNode;
      if (!open[node]) continue;
      const float dec = static_cast<float>(rnd() % 3) * 0.25f;
      const float key = keyOf[node] - dec;
      keyOf[node] = key;
      ref.modify(node, key);
      const int pos = h.findPosAll(true, node);
      if (pos < 0) return op + 1;
      h.heapUp(pos, key, node);
    }
    if (static_cast<size_t>(h.size) != ref.k.size()) return op + 1;
    for (int i = 0; i < h.size; ++i) {
      float k; uint32_t s;
      h.hget(i, k, s);
      if (k != ref.k[i] || s != ref.s[i]) return op + 1;
    }
  }
  return 0;
}

// Variant <TS, V> against the shipped <63, 1> in lock step on the same queries: after every
// step() both must hold the same heap (every entry), node count, best node and event -- a
// difference that happens not to change the corridor still counts.  Returns 0, or 1 + the index
// of the first query that diverges.
template <int TS, int V, int F = 0>
static long laneLockstep(void* h, const float* starts, const float* ends, long n, int fastFail) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  const NavView& nav = e->nav;
  const size_t per = laneScratchBytes(nav.numKeys) + 64;
  std::vector<char> sa(per, 0), sb(per, 0);
  std::vector<float> Ka(63), Kb(TS);
  std::vector<uint16_t> Sa(63), Sb(TS);
  std::vector<uint32_t> ra(kMaxPathPolys), rb(kMaxPathPolys);
  LaneSearch<1, 63, 4, 1> a{};
  LaneSearch<1, TS, 4, V, F> b{};
  auto carve = [&](auto& s, std::vector<char>& buf, std::vector<float>& K, std::vector<uint16_t>& S) {
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~uintptr_t(15));
    s.K = K.data(); s.S = S.data();
    s.tab = reinterpret_cast<uint16_t*>(base);
    s.rec = base + laneTabBytes(nav.numKeys);
    s.G = reinterpret_cast<LaneHeapEnt*>(s.rec + kLaneRecBytes);
    s.gen = 0;
      s.mode = kLIdle;
  };
  carve(a, sa, Ka, Sa);
  carve(b, sb, Kb, Sb);
  for (long i = 0; i < n; ++i) {
    const Nearest s = findNearestPoly(nav, grp, starts + 3 * i, kExt, -1, q);
    const Nearest t = findNearestPoly(nav, grp, ends + 3 * i, kExt, -1, q);
    if (s.g == kNoPoly || t.g == kNoPoly || s.g == t.g) continue;
    if (nav.polys[s.g].island < 0 || nav.polys[s.g].island != nav.polys[t.g].island) continue;
    if (a.gen >= kLaneGenMax || b.gen >= kLaneGenMax) {
      memset(a.tab, 0, laneTabBytes(nav.numKeys)); a.gen = 0;
      memset(b.tab, 0, laneTabBytes(nav.numKeys)); b.gen = 0;
    }
    a.begin(nav, static_cast<uint32_t>(i), s.g, s.pt, t.g, t.pt, ra.data());
    b.begin(nav, static_cast<uint32_t>(i), s.g, s.pt, t.g, t.pt, rb.data());
    int ev = kLEvNone;
    while (ev == kLEvNone) {
      ev = a.step(nav, fastFail != 0, true);
      const int evb = b.step(nav, fastFail != 0, true);
      if (ev != evb || a.size != b.size || a.nodeCount != b.nodeCount || a.lastBest != b.lastBest ||
          a.mode != b.mode || a.xk != b.xk)
        return i + 1;
      if (a.mode == kLSearch)
        for (int j = 0; j < a.size; ++j) {
          float ka, kb;
          uint32_t na, nb;
          a.hget(j, ka, na);
          b.hget(j, kb, nb);
          if (ka != kb || na != nb) return i + 1;
        }
    }
    if (a.status != b.status || memcmp(ra.data(), rb.data(), sizeof(uint32_t) * kMaxPathPolys) != 0) return i + 1;
  }
  return 0;
}

extern "C" {

void emu_find_path_lane(void* h, const float* starts, const float* ends, long n, int fastFail, int allCorridors,
                        unsigned* out_corridor, unsigned* out_info) {
  laneSearchRun<71, 10, 3>(h, starts, ends, n, fastFail, allCorridors, out_corridor, out_info);  // the shipped configuration
}

// The same with `ts` heap entries in the "shared" array (small values push most heap levels
// through the global-memory code) and code variant `v` (LaneSearch's V).  Returns 0, or -1
// for a combination that is not instantiated.
int emu_find_path_lane_v(void* h, int ts, int v, const float* starts, const float* ends, long n, int fastFail,
                         int allCorridors, unsigned* out_corridor, unsigned* out_info) {
#define HBN_EMU_LANE(T, VV) \
  if (ts == T && v == VV) { laneSearchRun<T, VV>(h, starts, ends, n, fastFail, allCorridors, out_corridor, out_info); return 0; }
  HBN_EMU_LANE(3, 1) HBN_EMU_LANE(3, 9) HBN_EMU_LANE(63, 9) HBN_EMU_LANE(95, 9) HBN_EMU_LANE(71, 10) HBN_EMU_LANE(7, 10) HBN_EMU_LANE(31, 10)
#undef HBN_EMU_LANE
  return -1;
}

// Memory-locality statistics of the lane search on a sample of queries (tools/lane_stats.py; design aid for
// the per-lane data layout).  out[0] searches, [1] expansions, [2] nodes allocated, [3] distinct 32 B sectors
// of a u16-per-key table touched (summed over searches), [4] the same for a bit-per-key map, [5] sum over the
// searches of the span (in 16 B units) between the lowest and highest touched map word, [6..] histogram of
// the age (expansions since the node was allocated) of the popped node: <=1, <=2, <=4, ... <=1024, more.
void emu_lane_stats(void* h, const float* starts, const float* ends, long n, long* out) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  const NavView& nav = e->nav;
  constexpr int TS = 71;
  std::vector<char> scratch(laneScratchBytes(nav.numKeys) + 64, 0);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch.data()) + 15) & ~uintptr_t(15));
  std::vector<float> K(TS);
  std::vector<uint16_t> S(TS);
  std::vector<uint32_t> ring(kMaxPathPolys);
  LaneSearch<1, TS, 4, 10> s{};
  s.K = K.data(); s.S = S.data();
  s.tab = reinterpret_cast<uint16_t*>(base);
  s.rec = base + laneTabBytes(nav.numKeys);
  s.G = reinterpret_cast<LaneHeapEnt*>(s.rec + kLaneRecBytes);
  s.gen = 0;
  s.mode = kLIdle;
  for (int i = 0; i < 32; ++i) out[i] = 0;
  std::vector<long> born(kMaxNodes);
  for (long i = 0; i < n; ++i) {
    const Nearest a = findNearestPoly(nav, grp, starts + 3 * i, kExt, -1, q);
    const Nearest b = findNearestPoly(nav, grp, ends + 3 * i, kExt, -1, q);
    if (a.g == kNoPoly || b.g == kNoPoly || vfuzzyEq(a.pt, b.pt)) continue;
    const int32_t si = nav.polys[a.g].island, ei = nav.polys[b.g].island;
    if (si < 0 || si != ei || a.g == b.g || !vfinite(a.pt) || !vfinite(b.pt)) continue;
    memset(s.tab, 0, laneTabBytes(nav.numKeys));
    s.gen = 0;
    s.begin(nav, static_cast<uint32_t>(i), a.g, a.pt, b.g, b.pt, ring.data());
    int ev = kLEvNone;
    long t = 0;
    born[0] = 0;
    while (ev == kLEvNone) {
      const int before = s.nodeCount;
      if (s.mode == kLSearch && s.size > 0) {
        const long age = t - born[s.S[0]];
        int bkt = 0;
        while (bkt < 11 && age > (1L << bkt)) bkt++;
        out[6 + bkt]++;
        t++;
      }
      ev = s.step(nav, true, false);
      for (int k = before; k < s.nodeCount; ++k) born[k] = t;
    }
    out[0]++;
    out[1] += s.expanded;
    out[2] += s.nodeCount;
    long lastT = -1, lastB = -1, lo = -1, hi = -1;
    for (uint32_t k = 0; k < nav.numKeys; ++k) {
      if ((s.tab[k] >> kLaneSlotBits) != s.gen) continue;
      if (static_cast<long>(k >> 4) != lastT) { lastT = k >> 4; out[3]++; }
      if (static_cast<long>(k >> 8) != lastB) { lastB = k >> 8; out[4]++; }
      if (lo < 0) lo = k >> 7;
      hi = k >> 7;
    }
    out[5] += hi - lo + 1;
    // distinct blocks of 8 / 16 / 32 keys: [18..20] sums, [21..23] maxima; [24] searches with > 256 16-key blocks
    long cnt[3] = {0, 0, 0}, last[3] = {-1, -1, -1};
    for (uint32_t k = 0; k < nav.numKeys; ++k) {
      if ((s.tab[k] >> kLaneSlotBits) != s.gen) continue;
      for (int b = 0; b < 3; ++b)
        if (static_cast<long>(k >> (3 + b)) != last[b]) { last[b] = k >> (3 + b); cnt[b]++; }
    }
    for (int b = 0; b < 3; ++b) { out[18 + b] += cnt[b]; if (cnt[b] > out[21 + b]) out[21 + b] = cnt[b]; }
    if (cnt[1] > 256) out[24]++;
  }
}

// find_path(MultiGoalShortestPath), fresh object per start: ends [n, g, 3]
// pruned != 0: the two-round scheme of the batched entry point (kMultiGoalFirst, hbn_query.h):
// goals left out of both rounds keep an infinite distance; out_searched counts the pair searches.
void emu_find_path_multigoal(void* h, const float* starts, const float* ends, long n, int g,
                             float* out_dist, int* out_idx, int pruned, long* out_searched) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  const int cap = kMaxNodes;
  std::vector<char> buf(astarWsBytes(cap) + 64);
  void* aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~uintptr_t(15));
  AStarWs w = astarWsCarve(aligned, cap);
  std::vector<uint32_t> eG(g);
  std::vector<float> dist(g), bounds(g);
  std::vector<int32_t> order(g);
  for (long i = 0; i < n; ++i) {
    const float* st = starts + 3 * i;
    const Nearest s = findNearestPoly(e->nav, grp, st, kExt, -1, q);
    std::vector<Nearest> et(g);
    for (int k = 0; k < g; ++k) {
      et[k] = findNearestPoly(e->nav, grp, ends + (i * g + k) * 3, kExt, -1, q);
      eG[k] = et[k].g;
    }
    auto search = [&](int k) {
      const float* en = ends + (i * g + k) * 3;
      memset(w.hash, 0, sizeof(uint32_t) * 2 * cap);
      const PathResult r = findPathInternal(e->nav, w, st, en, s.g, s.pt, et[k].g, et[k].pt, true, nullptr, 0, nullptr);
      dist[k] = r.dist;
      if (out_searched) (*out_searched)++;
    };
    if (!pruned) {
      for (int k = 0; k < g; ++k) search(k);
    } else {
      for (int k = 0; k < g; ++k) dist[k] = INFINITY;
      if (multiGoalOrder(g, st, s.g != kNoPoly, ends + i * g * 3, eG.data(), bounds.data(), order.data())) {
        for (int k = 0; k < g && k < kMultiGoalFirst; ++k) search(order[k]);
        const float rb = multiGoalRunningBest(g, kMultiGoalFirst, eG.data(), dist.data(), bounds.data(), order.data());
        for (int k = kMultiGoalFirst; k < g; ++k)
          if (!(bounds[order[k]] > rb)) search(order[k]);
      }
    }
    multiGoalSelect(g, st, s.g != kNoPoly, ends + i * g * 3, eG.data(), dist.data(), bounds.data(),
                    order.data(), &out_dist[i], &out_idx[i]);
  }
}

// the restated std::sort on its own: order[n] <- argsort of key[n] as libstdc++ leaves it
void emu_std_sort_order(const float* key, int n, int* order) {
  for (int i = 0; i < n; ++i) order[i] = i;
  stdSortOrder(order, key, n);
}

void emu_try_step(void* h, const float* starts, const float* ends, long n, int allowSliding,
                  float* out) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  for (long i = 0; i < n; ++i) {
    const float* st = starts + 3 * i;
    const float* en = ends + 3 * i;
    const Nearest s = findNearestPoly(e->nav, grp, st, kExt, -1, q);
    const Nearest t = findNearestPoly(e->nav, grp, en, kExt, -1, q);
    float ep[3];
    uint32_t last = kNoPoly;
    if (!tryStepPhaseA(e->nav, s.g, s.pt, t.g, en, allowSliding != 0, ep, &last)) {
      memcpy(out + 3 * i, st, 12);
      continue;
    }
    const Nearest e2 = findNearestPoly(e->nav, grp, ep, kExt, -1, q);
    tryStepPhaseB(e->nav, s.g, e2.g, last, ep);
    memcpy(out + 3 * i, ep, 12);
  }
}

// out [n,7]: hitPos, hitNormal, hitDist (closestObstacleSurfacePoint, PathFinder.cpp:1794-1812)
void emu_obstacle(void* h, const float* pts, long n, float maxRadius, int cap, float* out,
                  int* out_overflow) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  std::vector<char> buf(astarWsBytes(cap) + 64);
  void* aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~uintptr_t(15));
  AStarWs w = astarWsCarve(aligned, cap);
  for (long i = 0; i < n; ++i) {
    float* o = out + 7 * i;
    const Nearest s = findNearestPoly(e->nav, grp, pts + 3 * i, kExt, -1, q);
    for (int k = 0; k < 6; ++k) o[k] = 0.f;
    o[6] = INFINITY;
    if (out_overflow) out_overflow[i] = 0;
    if (s.g == kNoPoly) continue;
    memset(w.hash, 0, sizeof(uint32_t) * 2 * cap);
    float hitDist = NAN;
    const uint32_t st = distanceToWall(e->nav, w, s.g, s.pt, maxRadius, &hitDist, o, o + 3);
    if (st == 0xffffffffu && out_overflow) out_overflow[i] = 1;
    o[6] = hitDist;
  }
}

// The same through the tiers of hbn_closest_obstacle_dev: the per-thread search with its small pool
// (distanceToWallSmall, k_wall_lane), then the 128- and 2048-node workspaces for what overflows.
// out_tier[i] = tier that answered (0, 1, 2).  smallCap: 4 to exercise the hand-over even more, else the device's kWallLaneCap.
void emu_obstacle_tiers(void* h, const float* pts, long n, float maxRadius, int smallCap, float* out, int* out_tier) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  uint32_t q[2];
  std::vector<uint32_t> small(6 * kWallLaneCap);
  for (long i = 0; i < n; ++i) {
    float* o = out + 7 * i;
    const Nearest s = findNearestPoly(e->nav, grp, pts + 3 * i, kExt, -1, q);
    for (int k = 0; k < 6; ++k) o[k] = 0.f;
    o[6] = INFINITY;
    out_tier[i] = 0;
    if (s.g == kNoPoly) continue;
    float hitDist = NAN;
    uint32_t st = smallCap == 4 ? distanceToWallSmall<4, 1>(e->nav, small.data(), s.g, s.pt, maxRadius, &hitDist, o, o + 3)
                                : distanceToWallSmall<kWallLaneCap, 1>(e->nav, small.data(), s.g, s.pt, maxRadius, &hitDist, o, o + 3);
    for (int cap : {128, 2048}) {
      if (st != 0xffffffffu) break;
      out_tier[i]++;
      // (the overflowing tier may have moved the hit position: the device restarts from zeros too)
      for (int k = 0; k < 6; ++k) o[k] = 0.f;
      std::vector<char> buf(astarWsBytes(cap) + 64);
      void* aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(buf.data()) + 15) & ~uintptr_t(15));
      AStarWs w = astarWsCarve(aligned, cap);
      memset(w.hash, 0, sizeof(uint32_t) * 2 * cap);
      hitDist = NAN;
      st = distanceToWall(e->nav, w, s.g, s.pt, maxRadius, &hitDist, o, o + 3);
    }
    o[6] = hitDist;
  }
}

// get_random_navigable_point with the counter-based stream (PathFinder.cpp:1236-1281)
void emu_random_points(void* h, long n, int maxTries, const int* islands, unsigned long long seed,
                       unsigned long long query0, float* out_pts, unsigned* out_refs) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  for (long i = 0; i < n; ++i) {
    uint32_t draw = 0;
    uint32_t g = kNoPoly;
    float pt[3] = {NAN, NAN, NAN};
    for (int t = 0; t < maxTries; ++t) {
      uint32_t used = 0;
      float p[3];
      g = findRandomPoint(e->nav, grp, seed, query0 + i, draw, islands ? islands[i] : -1, p, &used);
      draw += used;
      if (g != kNoPoly) { memcpy(pt, p, 12); break; }
    }
    memcpy(out_pts + 3 * i, pt, 12);
    if (out_refs) out_refs[i] = g != kNoPoly ? e->nav.polys[g].ref : 0;
  }
}

// get_random_navigable_point_near (randomPointInCircle, hbn_query.h)
void emu_random_points_near(void* h, long n, const float* centers, float radius, int maxTries,
                            const int* islands, unsigned long long seed, unsigned long long query0,
                            float* out_pts) {
  Emu* e = static_cast<Emu*>(h);
  HostGroup grp;
  for (long i = 0; i < n; ++i)
    randomPointInCircle(e->nav, grp, seed, query0 + i, islands ? islands[i] : -1, centers + 3 * i, radius,
                        maxTries, out_pts + 3 * i);
}

// Heap code of hbn_astar_lane.h against dtNodeQueue restated plainly: random push / pop / modify
// sequences over keys from `keyLevels` values (ties everywhere), every heap entry compared after
// every operation.  Returns 0, the 1-based index of the first diverging operation, or -1.
long emu_lane_heap_fuzz(int ts, int v, unsigned seed, long ops, int keyLevels) {
#define HBN_EMU_HEAP(T, VV) if (ts == T && v == VV) return laneHeapFuzz<T, VV>(seed, ops, keyLevels);
  HBN_EMU_HEAP(3, 1) HBN_EMU_HEAP(7, 1) HBN_EMU_HEAP(31, 1) HBN_EMU_HEAP(63, 1) HBN_EMU_HEAP(71, 10) HBN_EMU_HEAP(95, 10)
  HBN_EMU_HEAP(47, 1) HBN_EMU_HEAP(55, 1) HBN_EMU_HEAP(39, 1)
#undef HBN_EMU_HEAP
  return -1;
}

long emu_lane_lockstep(void* h, int ts, int v, const float* starts, const float* ends, long n, int fastFail) {
#define HBN_EMU_LOCK(T, VV) if (ts == T && v == VV) return laneLockstep<T, VV>(h, starts, ends, n, fastFail);
#define HBN_EMU_LOCKF(T, VV, FF) if (ts == T && v == VV + 100 * FF) return laneLockstep<T, VV, FF>(h, starts, ends, n, fastFail);
  // v = V + 100 * F: code-shape bits of LaneSearch (one replay loop body, policies created once)
  HBN_EMU_LOCKF(71, 10, 1) HBN_EMU_LOCKF(71, 10, 3) HBN_EMU_LOCKF(3, 10, 3) HBN_EMU_LOCKF(63, 1, 1)
  HBN_EMU_LOCK(3, 1) HBN_EMU_LOCK(7, 1) HBN_EMU_LOCK(31, 1) HBN_EMU_LOCK(47, 1) HBN_EMU_LOCK(55, 1) HBN_EMU_LOCK(39, 1) HBN_EMU_LOCK(59, 1) HBN_EMU_LOCK(95, 1)
  HBN_EMU_LOCK(3, 9) HBN_EMU_LOCK(63, 9) HBN_EMU_LOCK(95, 9) HBN_EMU_LOCK(63, 10) HBN_EMU_LOCK(71, 10) HBN_EMU_LOCK(95, 10) HBN_EMU_LOCK(63, 8)
#undef HBN_EMU_LOCK
#undef HBN_EMU_LOCKF
  return -1;
}

}  // extern "C"

// Lock-step find_path search: G lanes per query, 32/G queries per warp.
//
// dtNavMeshQuery::findPath (DQ.cpp:973-1165) is serial per query: its result depends on the
// exact order of the heap operations (DNode.cpp:156-200) and of node allocation against the
// 2048-node pool (PF.cpp:937).  One query cannot use a warp: a poly has 2.3 links on average,
// and the heap is operated by one lane.  So a warp runs 32/G queries side by side, in lock
// step: every phase of the expansion loop (pop + sift, neighbour evaluation, heap replay) is
// executed once per warp instruction for all of its queries -- the serial heap phases of
// different queries share issue slots, the neighbour phases fill the lanes.
//
// Per query ("group"):
//   shared : node table, kGTab 16-bit entries (valid | open | closed | 13-bit fingerprint),
//            open addressing, slot == node id;  binary heap of 16 B entries
//            {total, link window, poly | slot, parent poly | slot} -- a pop needs nothing else
//            to issue the link loads;
//   global : one 32 B record per slot {pos, cost | heuristic, key, via link, parent slot};
//            a fingerprint hit is verified against the record's key (the record is needed
//            anyway for a found node).
// The search writes the corridor (the `via` links of the parent chain) to global memory; the
// funnel runs in its own thread-per-query kernel (k_fp_funnel), so a finished query never
// stalls its warp neighbours with string pulling.
#pragma once
#include <cuda_runtime.h>
#include "hbn_query.h"
#include "hbn_astar_warp.cuh"  // kSearchOverflow / kSearchWatchdog / kMaxExpansions, nodeHash

namespace hbn {

constexpr int kGTab = 2560;  // node table slots: load <= 0.8 at the 2048-node limit
constexpr uint32_t kGValid = 0x8000u, kGOpen = 0x4000u, kGClosed = 0x2000u, kGFpMask = 0x1fffu;
constexpr uint32_t kGNoParent = 0x00ffffffu;  // parent field of the start node's heap entry
__device__ __forceinline__ uint32_t gTabHome(uint32_t h) { return __umulhi(h, kGTab); }
__device__ __forceinline__ uint32_t gTabNext(uint32_t s) { return s + 1 == kGTab ? 0u : s + 1; }

struct __align__(16) GNodeRec {
  float px, py, pz, cost;
  float heur;     // heuristic term of the node's total (a function of its fixed position only)
  uint32_t key;   // poly | crossSide state << 24
  uint32_t via;   // LinkRec index the node was (last) entered through; kNoPoly for the start
  uint32_t pidx;  // parent slot + 1, 0 = none
};
static_assert(sizeof(GNodeRec) == 32, "GNodeRec");

struct __align__(16) GHeapEnt {
  float key;
  uint32_t lnk;  // link window of the node's poly: start (27 bits) | count << 27
  uint32_t gs;   // poly (24 bits) | slot bits 0..7 << 24
  uint32_t ps;   // parent poly (24 bits, kGNoParent = none) | slot bits 8..11 << 24
};
__device__ __forceinline__ uint32_t gEntSlot(uint32_t gs, uint32_t ps) { return (gs >> 24) | ((ps >> 24) << 8); }

template <int OC>
__host__ __device__ constexpr size_t gGroupSharedBytes() { return kGTab * 2 + (OC + 2) * sizeof(GHeapEnt); }
__host__ __device__ constexpr size_t gGroupGlobalBytes() { return static_cast<size_t>(kGTab) * sizeof(GNodeRec); }

// dtNodeQueue::bubbleUp, DNode.cpp:156-167
__device__ __forceinline__ void gHeapUp(GHeapEnt* hp, int i, const GHeapEnt node) {
  while (i > 0) {
    const int parent = (i - 1) >> 1;
    const GHeapEnt p = hp[parent];
    if (!(p.key > node.key)) break;
    hp[i] = p;
    i = parent;
  }
  hp[i] = node;
}
// dtNodeQueue::pop's trickleDown (DNode.cpp:169-184) for a heap that has `n` entries left
__device__ __forceinline__ void gHeapPopSift(GHeapEnt* hp, int n) {
  const GHeapEnt last = hp[n];
  int i = 0, child = 1;
  while (child < n) {
    GHeapEnt c0 = hp[child];
    const GHeapEnt c1 = hp[child + 1];
    if ((child + 1) < n && c0.key > c1.key) {
      c0 = c1;
      child++;
    }
    hp[i] = c0;
    i = child;
    child = 2 * i + 1;
  }
  gHeapUp(hp, i, last);
}

// classes of a find_path query (k_fp_classify), PF.cpp:1426-1447
enum : uint8_t {
  kClsNone = 0,      // a snap failed or the islands differ: no path
  kClsTrivial = 1,   // pathStart == pathEnd (PF.cpp:1434-1436)
  kClsSamePoly = 2,  // startRef == endRef (DQ.cpp:996-1001)
  kClsInvalid = 3,   // non-finite snapped point: findPath fails with INVALID_PARAM
  kClsSearch = 4,
  kClsSkip = 5       // masked out by the caller (multi-goal pruning): outputs are left alone
};

struct AStarGArgs {
  const uint32_t* sG;   // projectToPoly results
  const float* sPt;
  const uint32_t* eG;
  const float* ePt;
  const uint32_t* work;       // queries that need a search (k_fp_classify)
  const uint32_t* workCount;
  uint32_t* counter;          // atomic work cursor
  uint32_t* overflow;         // queries whose open list outgrew this tier
  uint32_t* overflowCount;
  uint32_t* astat;            // [n] findPath status word, kSearchOverflow = left to the next tier
  int32_t* fullLen;           // [n] untruncated corridor length (0 = not extracted)
  uint32_t* corrVia;          // [n, 256] corridor as entering links; element i at (first + i) & 255
  char* scratch;              // node records, one gGroupGlobalBytes() slot per group of the grid
  int startDiv;               // > 1: query q starts at point q / startDiv (multi-goal pairs)
  int fastFail;
  int allCorridors;           // extract the corridor of unsuccessful searches too
  unsigned long long* workCtr;
  unsigned int* fault;
  int laneLimit;              // k_astar_lane: lanes of a warp that take queries (0 = all 32); small batches spread over more warps
};

enum { kGIdle = 0, kGSearch = 1, kGDone = 2 };

template <int G, int OC>
__global__ void __launch_bounds__(32) k_astar_g(NavView nav, AStarGArgs a) {
  static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group width");
  constexpr int QPW = 32 / G;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) char smem[];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int grp = lane / G;
  const int lead = grp * G;
  const uint32_t gmask = (G == 32) ? FULL : (((1u << G) - 1u) << lead);
  const uint32_t ltMask = gmask & ((1u << lane) - 1u);  // lanes of my group in front of me
  uint16_t* tab = reinterpret_cast<uint16_t*>(smem + static_cast<size_t>(grp) * gGroupSharedBytes<OC>());
  GHeapEnt* hp = reinterpret_cast<GHeapEnt*>(smem + static_cast<size_t>(grp) * gGroupSharedBytes<OC>() + kGTab * 2) + 1;
  GNodeRec* rec = reinterpret_cast<GNodeRec*>(a.scratch) + (static_cast<size_t>(blockIdx.x) * QPW + grp) * kGTab;
  const uint32_t nWork = *a.workCount;

  // group state (identical in every lane of a group)
  int mode = kGIdle;
  uint32_t q = 0, endG = 0, lastBest = 0, lastBestG = 0;
  float ep[3] = {0.f, 0.f, 0.f};
  int size = 0, nodeCount = 0;
  float lastBestCost = 0.f;
  bool outOfNodes = false;
  uint32_t expanded = 0, nLinks = 0, nNeigh = 0;

  for (;;) {
    // ---- idle groups take the next query ------------------------------------------------
    if (__any_sync(FULL, mode == kGIdle)) {
      if (mode == kGIdle) {
        uint32_t wi = 0;
        if (gl == 0) wi = atomicAdd(a.counter, 1u);
        wi = __shfl_sync(gmask, wi, lead);
        if (wi >= nWork) {
          mode = kGDone;
        } else {
          q = a.work[wi];
          const uint32_t qs = a.startDiv > 1 ? q / static_cast<uint32_t>(a.startDiv) : q;
          const uint32_t startG = a.sG[qs];
          endG = a.eG[q];
          const float sp[3] = {a.sPt[3 * qs], a.sPt[3 * qs + 1], a.sPt[3 * qs + 2]};
          ep[0] = a.ePt[3 * static_cast<size_t>(q)];
          ep[1] = a.ePt[3 * static_cast<size_t>(q) + 1];
          ep[2] = a.ePt[3 * static_cast<size_t>(q) + 2];
          uint4* t4 = reinterpret_cast<uint4*>(tab);
          for (int i = gl; i < kGTab * 2 / 16; i += G) t4[i] = make_uint4(0u, 0u, 0u, 0u);
          const PolyRec* spoly = &nav.polys[startG];
          const uint32_t slnk = spoly->linkStart | (static_cast<uint32_t>(spoly->linkCount) << 27);
          const uint32_t sh = nodeHash(startG);
          const uint32_t sslot = gTabHome(sh);
          const float stotal = vdist(sp, ep) * kHScale;
          __syncwarp(gmask);
          if (gl == 0) {
            tab[sslot] = static_cast<uint16_t>(kGValid | kGOpen | (sh & kGFpMask));
            float4* ra = reinterpret_cast<float4*>(&rec[sslot]);
            ra[0] = make_float4(sp[0], sp[1], sp[2], 0.f);
            reinterpret_cast<uint4*>(ra)[1] = make_uint4(__float_as_uint(stotal), startG, kNoPoly, 0u);
            hp[0] = GHeapEnt{stotal, slnk, startG | ((sslot & 0xffu) << 24), kGNoParent | ((sslot >> 8) << 24)};
          }
          size = 1;
          nodeCount = 1;
          lastBest = sslot;
          lastBestG = startG;
          lastBestCost = stotal;
          outOfNodes = false;
          expanded = nLinks = nNeigh = 0;
          mode = kGSearch;
        }
      }
      __syncwarp();
    }
    if (__all_sync(FULL, mode == kGDone)) break;

    // ---- pop (DQ.cpp:1027-1040) ----------------------------------------------------------
    const bool act = mode == kGSearch;
    bool finish = act && size == 0;  // open list exhausted: partial result
    bool overflowed = false, watchdog = false;
    const bool go = act && !finish;
    const GHeapEnt top = hp[0];
    const uint32_t bslot = gEntSlot(top.gs, top.ps);
    const uint32_t bestG = top.gs & 0x00ffffffu;
    const uint32_t parentG = top.ps & 0x00ffffffu;
    const uint32_t l0 = top.lnk & 0x07ffffffu;
    const int ln = go ? static_cast<int>(top.lnk >> 27) : 0;
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 La = make_uint4(0u, 0u, 0u, kNoPoly), Lb = make_uint4(0u, 0u, 0u, 0u);
    if (go) {
      ba = reinterpret_cast<const float4*>(&rec[bslot])[0];
      if (gl < ln) {
        const uint4* lp = reinterpret_cast<const uint4*>(&nav.links[l0 + gl]);
        La = __ldg(lp);
        Lb = __ldg(lp + 1);
      }
      size--;
    }
    __syncwarp();  // every lane has read hp[0] before the leaders rewrite their heaps
    if (go && gl == 0) {
      gHeapPopSift(hp, size);
      tab[bslot] = static_cast<uint16_t>((tab[bslot] & ~kGOpen) | kGClosed);
    }
    if (go && bestG == endG) {
      lastBest = bslot;
      lastBestG = bestG;
      finish = true;
    }
    bool ex = go && !finish;  // groups that expand a poly this iteration
    if (ex) {
      if (expanded >= kMaxExpansions) {
        watchdog = true;
        ex = false;
      } else {
        expanded++;
        nLinks += ln;
      }
    }
    __syncwarp();

    // ---- neighbours, G links at a time (DQ.cpp:1042-1160) --------------------------------
    const float bpos[3] = {ba.x, ba.y, ba.z};
    const float bcost = ba.w;
    const int lnMax = __reduce_max_sync(FULL, ex ? ln : 0);
    for (int base = 0; base < lnMax; base += G) {
      if (base > 0) {
        La = make_uint4(0u, 0u, 0u, kNoPoly);
        Lb = make_uint4(0u, 0u, 0u, 0u);
        if (ex && base + gl < ln) {
          const uint4* lp = reinterpret_cast<const uint4*>(&nav.links[l0 + base + gl]);
          La = __ldg(lp);
          Lb = __ldg(lp + 1);
        }
      }
      const uint32_t nei = (ex && base + gl < ln) ? La.w : kNoPoly;
      const uint32_t meta = Lb.y;
      nNeigh += __popc(__ballot_sync(FULL, nei != kNoPoly) & gmask);
      const bool cand = nei != kNoPoly && nei != parentG && (meta & kLinkPassBit) != 0;
      const uint32_t key = nei | (((meta >> kLinkStateShift) & 3u) << 24);
      const uint32_t kh = nodeHash(key);
      const uint32_t fp = kh & kGFpMask;
      uint32_t pend = __ballot_sync(FULL, cand) & gmask;
      // Two links of one poly can lead to the same neighbour (flagged at flatten time): such a
      // link must see its twin's node, so a round ends in front of it.  Usually one round.
      const uint32_t dupLanes = __ballot_sync(FULL, cand && (meta & kLinkDupBit) != 0) & gmask;
      while (__any_sync(FULL, pend != 0)) {
        uint32_t cur = pend;
        if (dupLanes && pend) {
          const int first = __ffs(pend) - 1;
          const uint32_t later = dupLanes & pend & ~((2u << first) - 1u);
          if (later) cur = pend & ((1u << (__ffs(later) - 1)) - 1u);
        }
        pend &= ~cur;
        const bool mine = ex && ((cur >> lane) & 1u);

        // dtNodePool::getNode, DNode.cpp:121-152: lookup ...
        uint32_t slot = gTabHome(kh);
        uint32_t ent = 0;
        bool found = false, searching = mine, bad = false;
        float4 na = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 nb = make_uint4(0u, 0u, 0u, 0u);
        int probes = 0;
        while (__any_sync(FULL, searching)) {
          if (searching) {
            for (;;) {
              ent = tab[slot];
              if (ent == 0u || (ent & kGFpMask) == fp) break;
              slot = gTabNext(slot);
              if (++probes > kGTab) break;
            }
            if (probes > kGTab) {
              bad = true;
              searching = false;
            } else if (ent == 0u) {
              searching = false;
            } else {
              na = reinterpret_cast<const float4*>(&rec[slot])[0];
              nb = reinterpret_cast<const uint4*>(&rec[slot])[1];
              if (nb.y == key) {
                found = true;
                searching = false;
              } else {  // fingerprint collision
                slot = gTabNext(slot);
                probes++;
              }
            }
          }
        }
        if (__any_sync(FULL, bad)) {
          if (__ballot_sync(FULL, bad) & gmask) {
            watchdog = true;
            ex = false;
          }
        }
        __syncwarp();  // lookups done before this round's inserts / flag updates
        // ... allocation in link order against the 2048-node limit
        const bool live = mine && ex;
        const bool isNew = live && !found;
        const uint32_t newMask = __ballot_sync(FULL, isNew) & gmask;
        const bool allocFail = isNew && (nodeCount + __popc(newMask & ltMask)) >= kMaxNodes;
        const uint32_t failMask = __ballot_sync(FULL, allocFail) & gmask;
        if (failMask) {
          outOfNodes = true;
          if (a.fastFail) {  // PF.cpp:1450 has decided "no path" already
            finish = true;
            ex = false;
          }
        }
        if (ex) nodeCount += __popc(newMask & ~failMask);
        const bool ok = live && ex && !allocFail;
        uint32_t ins = __ballot_sync(FULL, ok && isNew) & gmask;
        while (__any_sync(FULL, ins != 0)) {  // inserts of one group one after the other
          if (ins) {
            if (lane == __ffs(ins) - 1) {
              for (int p2 = 0; p2 < kGTab && tab[slot] != 0u; ++p2) slot = gTabNext(slot);
              tab[slot] = static_cast<uint16_t>(kGValid | fp);
            }
            ins &= ins - 1;
          }
          __syncwarp();
        }
        float npos[3] = {__uint_as_float(La.x), __uint_as_float(La.y), __uint_as_float(La.z)};
        float ntotal = 0.f, nheur = 0.f;
        if (ok && found) {
          npos[0] = na.x; npos[1] = na.y; npos[2] = na.z;
          nheur = __uint_as_float(nb.x);
          ntotal = na.w + nheur;  // the node's total, as it was formed: cost + heuristic
        }
        // DQ.cpp:1088-1121.  A node's position never changes after its first visit, so neither
        // does its heuristic term: it is computed once and kept in the record.
        float cost, heuristic;
        {
          const float curCost = vdist(bpos, npos);
          if (nei == endG) {
            const float endCost = vdist(npos, ep);
            cost = bcost + curCost + endCost;
            heuristic = 0.f;
          } else {
            cost = bcost + curCost;
            heuristic = found ? nheur : vdist(npos, ep) * kHScale;
          }
        }
        const float total = cost + heuristic;
        const bool wasOpen = found && (ent & kGOpen) != 0;
        const bool wasClosed = found && (ent & kGClosed) != 0;
        const bool acc = ok && !((wasOpen || wasClosed) && total >= ntotal);  // DQ.cpp:1124-1130
        if (acc) {
          float4* ra = reinterpret_cast<float4*>(&rec[slot]);
          ra[0] = make_float4(npos[0], npos[1], npos[2], cost);
          reinterpret_cast<uint4*>(ra)[1] = make_uint4(__float_as_uint(heuristic), key,
                                                       l0 + static_cast<uint32_t>(base + gl), bslot + 1u);
          tab[slot] = static_cast<uint16_t>(kGValid | kGOpen | fp);
        }
        const uint32_t accMask = __ballot_sync(FULL, acc) & gmask;
        const uint32_t openMask = __ballot_sync(FULL, acc && wasOpen) & gmask;
        __syncwarp();
        // heap updates replayed in link order (DQ.cpp:1140-1152)
        const uint32_t nlnk = Lb.x | ((meta >> kLinkNeiCountShift) << 27);
        const GHeapEnt myEnt{total, nlnk, nei | ((slot & 0xffu) << 24), bestG | ((slot >> 8) << 24)};
        uint32_t m = accMask;
        while (__any_sync(FULL, m != 0)) {
          if (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if ((openMask >> b) & 1u) {  // dtNodeQueue::modify, DNode.h:132-142: locate, then bubbleUp
              const uint32_t s = __shfl_sync(gmask, slot, b);
              int pos = -1;
              for (int i = gl; i < size; i += G) {
                const uint2 e = *reinterpret_cast<const uint2*>(&hp[i].gs);
                if (gEntSlot(e.x, e.y) == s) pos = i;
              }
              const uint32_t pm = __ballot_sync(gmask, pos >= 0);
              pos = __shfl_sync(gmask, pos, pm ? (__ffs(pm) - 1) : lead);
              __syncwarp(gmask);
              if (lane == b && pos >= 0) gHeapUp(hp, pos, myEnt);
            } else if (size >= OC) {
              overflowed = true;
              ex = false;
              m = 0;
            } else {
              if (lane == b) gHeapUp(hp, size, myEnt);
              size++;
            }
          }
          __syncwarp();
        }
        // DQ.cpp:1154-1159: first neighbour (link order) with the smallest heuristic
        {  // heuristics are >= +0, so their bit patterns order like the floats
          const uint32_t hb = (acc && ex) ? __float_as_uint(heuristic) : 0xffffffffu;
          uint32_t mn = hb;
#pragma unroll
          for (int o = G / 2; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(FULL, mn, o));
          const uint32_t eq = __ballot_sync(FULL, hb == mn && hb != 0xffffffffu) & gmask;
          const int hl = eq ? (__ffs(eq) - 1) : lane;
          const uint32_t hs = __shfl_sync(FULL, slot, hl);
          const uint32_t hg = __shfl_sync(FULL, nei, hl);
          if (eq && __uint_as_float(mn) < lastBestCost) {
            lastBestCost = __uint_as_float(mn);
            lastBest = hs;
            lastBestG = hg;
          }
        }
        if (!ex) pend = 0;
      }
    }

    // ---- finished searches: status, corridor (getPathToNode, DQ.cpp:1167-1205) -----------
    if (__any_sync(FULL, (finish || overflowed || watchdog) && act)) {
      if (act && (finish || overflowed || watchdog)) {
        if (gl == 0) {
          if (watchdog) {
            atomicAdd(a.fault, 1u);
            a.fault[1] = q;
            a.fault[2] = 3u | (OC << 8);
            a.astat[q] = kDtFailure;
            a.fullLen[q] = 0;
          } else if (overflowed) {
            const uint32_t o = atomicAdd(a.overflowCount, 1u);
            a.overflow[o] = q;
            a.astat[q] = kSearchOverflow;
            a.fullLen[q] = 0;
          } else {
            uint32_t status = kDtSuccess;
            if (lastBestG != endG) status |= kDtPartialResult;
            if (outOfNodes) status |= kDtOutOfNodes;
            int k = 0;
            if (status == kDtSuccess || a.allCorridors) {
              uint32_t* cv = a.corrVia + static_cast<size_t>(q) * kMaxPathPolys;
              uint32_t cur = lastBest;
              for (;;) {
                const uint4 nb = reinterpret_cast<const uint4*>(&rec[cur])[1];
                cv[(kMaxPathPolys - 1 - k) & (kMaxPathPolys - 1)] = nb.z;
                k++;
                if (!nb.w) break;     // start node
                if (k > kGTab) {      // a parent cycle would be a bug
                  atomicAdd(a.fault, 1u);
                  a.fault[1] = q;
                  a.fault[2] = 4u;
                  break;
                }
                cur = nb.w - 1;
              }
            }
            a.astat[q] = status | ((k > kMaxPathPolys) ? kDtBufferTooSmall : 0u);
            a.fullLen[q] = k;
          }
          if (a.workCtr && !overflowed) {
            atomicAdd(a.workCtr + 0, static_cast<unsigned long long>(expanded));
            atomicAdd(a.workCtr + 1, static_cast<unsigned long long>(nLinks));
            atomicAdd(a.workCtr + 2, static_cast<unsigned long long>(nNeigh));
            atomicAdd(a.workCtr + 6, 1ull);
          }
        }
        mode = kGIdle;
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------
// findPathInternal's decisions that need no search (PF.cpp:1426-1447), one thread per query.
// The queries that do need one are put on the search list LONGEST FIRST: a search costs between
// a handful and ~2000 expansions, the persistent search kernels hand queries out in list order,
// and a cheap estimate of the cost (the straight-line distance between the snapped points, in
// kFpBuckets classes) is enough to keep the expensive ones out of the tail of the launch.
// Two passes: classify + histogram of the classes, then a scatter into per-class ranges.
constexpr int kFpBuckets = 32;
constexpr float kFpBucketWidth = 2.0f;  // metres of straight-line distance per class
constexpr uint8_t kFpNoBucket = 0xff;

__global__ void __launch_bounds__(256) k_fp_classify(NavView nav, const uint32_t* __restrict__ sG,
                                                     const float* __restrict__ sPt,
                                                     const uint32_t* __restrict__ eG,
                                                     const float* __restrict__ ePt, int64_t n, int startDiv,
                                                     const uint8_t* __restrict__ mask,
                                                     uint8_t* __restrict__ cls, uint8_t* __restrict__ bucket,
                                                     uint32_t* __restrict__ hist, uint32_t* __restrict__ workCount) {
  __shared__ uint32_t histS[kFpBuckets];
  if (threadIdx.x < kFpBuckets) histS[threadIdx.x] = 0;
  __syncthreads();
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q < n) {
    const int64_t qs = startDiv > 1 ? q / startDiv : q;
    const uint32_t s = sG[qs], e = eG[q];
    uint8_t c = kClsNone, b = kFpNoBucket;
    if (mask && !mask[q]) {
      c = kClsSkip;
    } else if (s != kNoPoly && e != kNoPoly) {
      const float sp[3] = {sPt[3 * qs], sPt[3 * qs + 1], sPt[3 * qs + 2]};
      const float ep[3] = {ePt[3 * q], ePt[3 * q + 1], ePt[3 * q + 2]};
      if (vfuzzyEq(sp, ep)) {
        c = kClsTrivial;
      } else {
        const int32_t si = nav.polys[s].island, ei = nav.polys[e].island;
        if (si >= 0 && si == ei) {  // hasConnection, PF.cpp:209-221
          if (s == e) c = kClsSamePoly;
          else if (!vfinite(sp) || !vfinite(ep)) c = kClsInvalid;
          else {
            c = kClsSearch;
            const float d = vdist(sp, ep) / kFpBucketWidth;
            b = static_cast<uint8_t>(d < static_cast<float>(kFpBuckets - 1) ? static_cast<int>(d) : kFpBuckets - 1);
            atomicAdd(&histS[b], 1u);
          }
        }
      }
    }
    cls[q] = c;
    bucket[q] = b;
  }
  __syncthreads();
  if (threadIdx.x < kFpBuckets && histS[threadIdx.x]) {
    atomicAdd(&hist[threadIdx.x], histS[threadIdx.x]);
    atomicAdd(workCount, histS[threadIdx.x]);
  }
}

// work[] = the search queries, class kFpBuckets-1 (longest) first.  cursor[]: zeroed.
__global__ void __launch_bounds__(256) k_fp_scatter(const uint8_t* __restrict__ bucket, int64_t n,
                                                    const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor,
                                                    uint32_t* __restrict__ work) {
  __shared__ uint32_t cntS[kFpBuckets], baseS[kFpBuckets];
  if (threadIdx.x < kFpBuckets) cntS[threadIdx.x] = 0;
  __syncthreads();
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint8_t b = q < n ? bucket[q] : kFpNoBucket;
  uint32_t r = 0;
  if (b != kFpNoBucket) r = atomicAdd(&cntS[b], 1u);
  __syncthreads();
  if (threadIdx.x < kFpBuckets && cntS[threadIdx.x]) {
    uint32_t start = 0;
    for (int j = kFpBuckets - 1; j > static_cast<int>(threadIdx.x); --j) start += hist[j];
    baseS[threadIdx.x] = start + atomicAdd(&cursor[threadIdx.x], cntS[threadIdx.x]);
  }
  __syncthreads();
  if (b != kFpNoBucket) work[baseS[b] + r] = static_cast<uint32_t>(q);
}

// Corridor as the search left it: entering links in a 256-entry ring; poly i is the neighbour
// its entering link leads to (poly 0 is the start poly).
struct ViaCorridor {
  const NavView& nav;
  const uint32_t* ring;  // [256]
  uint32_t first;
  uint32_t startG;
  __device__ __forceinline__ uint32_t via(int i) const { return ring[(first + i) & (kMaxPathPolys - 1)]; }
  __device__ __forceinline__ uint32_t poly(int i) const { return i == 0 ? startG : nav.links[via(i)].nei; }
  __device__ __forceinline__ uint32_t link(int i) const { return via(i + 1); }
  __device__ __forceinline__ void portal(int, uint32_t li, float* l, float* r) const {
    const float4 a = __ldg(reinterpret_cast<const float4*>(&nav.portals[li]));
    const float4 b = __ldg(reinterpret_cast<const float4*>(&nav.portals[li]) + 1);
    l[0] = a.x; l[1] = a.y; l[2] = a.z;
    r[0] = b.x; r[1] = b.y; r[2] = b.z;
  }
};

struct FpFunnelArgs {
  const float* starts;  // requested points (trap T2: the funnel uses these, not the snapped ones)
  const float* ends;
  const uint32_t* sG;
  const float* sPt;
  const uint32_t* eG;
  const float* ePt;
  const uint8_t* cls;
  const uint32_t* astat;
  const int32_t* fullLen;
  const uint32_t* corrVia;
  int64_t n;
  int startDiv;
  float* out_dist;
  int32_t* out_npts;
  float* out_pts;
  int max_pts;
  uint32_t* out_corridor;
  int32_t* out_ncorridor;
  uint32_t* out_status;
  unsigned long long* workCtr;
  int fillSkipped;  // kClsSkip queries: write an infinite distance (else nothing)
};

// findStraightPath + pathLength (PF.cpp:1456-1466) and every per-query output, one thread per
// query.
__global__ void __launch_bounds__(128) k_fp_funnel(NavView nav, FpFunnelArgs a) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= a.n) return;
  const int64_t qs = a.startDiv > 1 ? q / a.startDiv : q;
  const uint8_t c = a.cls[q];
  if (c == kClsSkip) {
    if (a.fillSkipped) a.out_dist[q] = infF();
    return;
  }
  float dist = infF();
  int npts = 0, ncorr = 0;
  uint32_t stA = 0, stS = 0, corrLinks = 0;
  float* outPts = a.out_pts ? a.out_pts + static_cast<size_t>(q) * a.max_pts * 3 : nullptr;
  uint32_t* outCorr = a.out_corridor ? a.out_corridor + static_cast<size_t>(q) * kMaxPathPolys : nullptr;
  if (c == kClsTrivial) {
    dist = 0.f;
    npts = 2;
    if (outPts) {
      if (a.max_pts > 0) { outPts[0] = a.sPt[3 * qs]; outPts[1] = a.sPt[3 * qs + 1]; outPts[2] = a.sPt[3 * qs + 2]; }
      if (a.max_pts > 1) { outPts[3] = a.ePt[3 * q]; outPts[4] = a.ePt[3 * q + 1]; outPts[5] = a.ePt[3 * q + 2]; }
    }
  } else if (c == kClsInvalid) {
    stA = kDtFailure | kDtInvalidParam;
  } else if (c == kClsSamePoly || c == kClsSearch) {
    const uint32_t sG = a.sG[qs];
    int fullLen = 1;
    stA = kDtSuccess;
    if (c == kClsSearch) {
      stA = a.astat[q];
      if (stA == kSearchOverflow) return;  // the next tier writes this query's outputs
      fullLen = a.fullLen[q];
    }
    ncorr = fullLen < kMaxPathPolys ? fullLen : kMaxPathPolys;
    ViaCorridor cor{nav, a.corrVia + static_cast<size_t>(q) * kMaxPathPolys,
                    static_cast<uint32_t>((kMaxPathPolys - fullLen) & (kMaxPathPolys - 1)), sG};
    if (outCorr || a.workCtr)
      for (int i = 0; i < ncorr; ++i) {
        const PolyRec* cp = &nav.polys[(c == kClsSamePoly) ? sG : cor.poly(i)];
        corrLinks += cp->linkCount;
        if (outCorr) outCorr[i] = cp->ref;
      }
    if (stA == kDtSuccess && ncorr > 0) {  // PF.cpp:1450
      const float rs[3] = {a.starts[3 * qs], a.starts[3 * qs + 1], a.starts[3 * qs + 2]};
      const float re[3] = {a.ends[3 * q], a.ends[3 * q + 1], a.ends[3 * q + 2]};
      Funnel f;
      f.out = outPts;
      f.maxOut = a.max_pts;
      stS = funnelStraightPathT(nav, rs, re, cor, ncorr, f);
      npts = f.count;
      if (stS == kDtSuccess && f.count != 0) dist = f.length;  // PF.cpp:1459
    }
  }
  const bool found = dist < infF();
  a.out_dist[q] = dist;
  if (a.out_npts) a.out_npts[q] = found ? npts : 0;
  if (a.out_ncorridor) a.out_ncorridor[q] = ncorr;
  if (a.out_status) {
    a.out_status[2 * q] = stA;
    a.out_status[2 * q + 1] = stS;
  }
  if (a.workCtr) {
    atomicAdd(a.workCtr + 3, static_cast<unsigned long long>(ncorr));
    atomicAdd(a.workCtr + 4, static_cast<unsigned long long>(corrLinks));
    atomicAdd(a.workCtr + 5, static_cast<unsigned long long>(found ? npts : 0));
    atomicAdd(a.workCtr + 7, 1ull);
  }
}

}  // namespace hbn

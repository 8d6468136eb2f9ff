// Lane-per-query find_path search: the state machine of ONE query, host + device code.
//
// dtNavMeshQuery::findPath (DQ.cpp:973-1165) is a serial algorithm; its result depends on the
// exact order of its heap operations (DNode.cpp:156-200) and of node allocation against the
// 2048-node pool (PF.cpp:937).  Splitting one query over lanes (round 1's first two mappings,
// deleted since) leaves most issue slots on warp-uniform bookkeeping: ~500 warp
// instructions per poly expansion, and shared memory caps an SM at ~25 queries in flight.
// Here a query belongs to ONE lane and a warp advances 32 queries in lock step (k_astar_lane,
// hbn_astar_lane.cuh): one warp instruction serves 32 expansions, and the per-query state that
// does not fit on chip lives in HBM, sized for 180 GB:
//
//   shared  : the top TS entries (71 shipped: 6 levels + 8) of the binary heap, 6 B per entry
//             {f32 total, u16 node}, interleaved over the lanes of the warp (entry i of lane l
//             at word i*32 + l: never a bank conflict).  Nothing else: heap positions of the
//             nodes are NOT tracked (a first version kept them in HBM: one scattered 2 B store
//             per heap move, 9 per expansion, 3/4 of all L2 requests; a second one in a shared
//             cache, which cost occupancy and 8 % of the instructions).  dtNodeQueue::modify
//             (DNode.h:132-142) is rare (0.06 per expansion on the C4 workload), so it does
//             what the reference does, a linear scan for the node -- but the WHOLE WARP scans
//             the heap of the lane that needs it (findPosAll), 32 entries per step;
//   global  : per lane
//             (a) the rest of the heap, 8 B entries; the two children of an entry share one
//                 aligned 16 B load;
//             (b) a DIRECT-MAPPED node table, one u16 per node key (poly, crossSide) -- the keys
//                 are enumerated by the flattener (LinkRec::neiKey, PolyRec::key0) -- holding
//                 generation << 11 | node: a lookup is one load, never a probe loop, and a new
//                 query just bumps the generation (the table is wiped every 31 queries);
//             (c) 2048 node records of 32 B in ALLOCATION order (dtNodePool's index order):
//                 {pos, cost | poly, parent poly + entering link, link window, parent node}.
//                 The first 16 B answer everything a revisit asks.  (Variant 1 keeps Detour's
//                 closed flag in the sign bit of the stored cost; the shipped variant stores no
//                 flag at all, see LaneSearch.)
//   A node's total is not stored: it is cost + heuristic(pos), recomputed with the same
//   operations.
//
// One step() = one iteration of the reference's while loop: pop, then the popped poly's links
// in chunks of kLaneChunk: all link records, then all table entries, then all found records
// are loaded before the serial part, so a lane has several independent loads in flight per
// stage instead of a chain of 3 dependent loads per neighbour.  The serial part is split too:
// first every link's cost test and record update (visit), which only queues the heap
// operation; then the queued operations are replayed in link order (they are the only heap
// operations between two pops, so the heap goes through exactly the reference's states).
// Replaying by queue position instead of link index keeps more lanes busy per instruction.
// Corridor extraction (getPathToNode, DQ.cpp:1167-1205) is a pointer chase; it runs as a mode
// of the same state machine, a few hops per step, so it never stalls the other 31 queries.
#pragma once
#include <stdint.h>
#include <string.h>

#include "hbn_query.h"

namespace hbn {

constexpr uint32_t kLaneSlotBits = 11;  // node index < 2048
constexpr uint32_t kLaneSlotMask = (1u << kLaneSlotBits) - 1u;
constexpr uint32_t kLaneGenMax = 31;    // generations 1..31, then the table is wiped
constexpr uint32_t kLaneNoParent = 0x00ffffffu;
constexpr int kLaneHops = 2;            // corridor hops per step
constexpr uint32_t kLaneClosedBit = 0x80000000u;  // sign bit of LaneRecA::cost
constexpr uint32_t kLaneMaxExpansions = 1u << 18;  // >> any legal search (2048 nodes, re-opens)
static_assert(kMaxNodes <= (1 << kLaneSlotBits), "node index must fit the table entry");

struct HBN_ALIGN(16) LaneRecA {
  float px, py, pz, cost;  // cost: sign bit set = closed
};
HBN_HD float laneSetClosed(float c) {
  uint32_t u;
  memcpy(&u, &c, 4);
  u |= kLaneClosedBit;
  memcpy(&c, &u, 4);
  return c;
}
HBN_HD bool laneIsClosed(float c) {
  uint32_t u;
  memcpy(&u, &c, 4);
  return (u & kLaneClosedBit) != 0;
}
HBN_HD float laneCost(float c) {
  uint32_t u;
  memcpy(&u, &c, 4);
  u &= ~kLaneClosedBit;
  memcpy(&c, &u, 4);
  return c;
}
struct HBN_ALIGN(16) LaneRecB {
  uint32_t poly;  // global poly index
  uint32_t w1;    // parent poly (24 bits, kLaneNoParent = none) | entering link's offset in the parent's window << 24
  uint32_t lnk;   // link window of the poly: start (27 bits) | count << 27
  uint32_t w3;    // parent node (12 bits) | has-parent << 12
};
struct HBN_ALIGN(16) LaneLinkLo {
  float mx, my, mz;
  uint32_t nei;
};
struct HBN_ALIGN(16) LaneLinkHi {
  uint32_t neiLinkStart, meta, neiRef, neiKey;
};
struct HBN_ALIGN(8) LaneHeapEnt {
  float key;
  uint32_t slot;
};
struct HBN_ALIGN(16) LaneHeapPair {
  LaneHeapEnt a, b;
};

// per-lane global scratch, one region per kind
constexpr size_t kLaneRecBytes = static_cast<size_t>(kMaxNodes) * 32;
// a multiple of 32: a 4-entry group of every lane's heap then sits in ONE 32 B sector
constexpr size_t kLaneHeapBytes = (static_cast<size_t>(kMaxNodes + 2) * sizeof(LaneHeapEnt) + 31) & ~static_cast<size_t>(31);
HBN_HD size_t laneTabBytes(uint32_t numKeys) { return (static_cast<size_t>(numKeys) * 2 + 15) & ~static_cast<size_t>(15); }
HBN_HD size_t laneScratchBytes(uint32_t numKeys) {
  return laneTabBytes(numKeys) + kLaneRecBytes + kLaneHeapBytes;
}

enum { kLIdle = 0, kLSearch = 1, kLExtract = 2, kLDone = 3 };
enum { kLEvNone = 0, kLEvFinished = 1, kLEvFault = 3, kLEvPoolExhausted = 4 /* internal */ };

// KEEP: the link records are the read-only, L2-sized part of the working set (a few MB against GBs of
// per-query search state streaming through L2): one 32 B load, evict_last in L1 and L2 (device only).
template <bool KEEP = false>
HBN_HD void laneLoadLink(const LinkRec* p, LaneLinkLo& lo, LaneLinkHi& hi) {
#if defined(__CUDA_ARCH__)
  if constexpr (KEEP) {
    uint32_t w[8];
    asm volatile("ld.global.nc.L1::evict_last.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
    lo.mx = __uint_as_float(w[0]); lo.my = __uint_as_float(w[1]); lo.mz = __uint_as_float(w[2]); lo.nei = w[3];
    hi.neiLinkStart = w[4]; hi.meta = w[5]; hi.neiRef = w[6]; hi.neiKey = w[7];
    return;
  }
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  lo.mx = a.x; lo.my = a.y; lo.mz = a.z; lo.nei = __float_as_uint(a.w);
  hi.neiLinkStart = b.x; hi.meta = b.y; hi.neiRef = b.z; hi.neiKey = b.w;
#else
  memcpy(&lo, p, 16);
  memcpy(&hi, reinterpret_cast<const char*>(p) + 16, 16);
#endif
}

// HS: distance (in elements) between consecutive shared heap entries of this lane (32 on the
// device, 1 in the host build); TS: heap entries kept in shared memory (odd, so that the
// children 2i+1, 2i+2 of an entry are both on the same side and pair-aligned in global);
// CH: links per load stage (registers vs. rounds); V: memory-policy variant (below).
// step() contains warp collectives on the device: all 32 lanes of the warp must call it
// together, whatever their mode.
// F: code-shape bits.  With everything unrolled the kernel is 40.6 KB of SASS, more than the 32 KB
// instruction cache level next to the SM, and its 16 warps sit at different places of it: the profile shows
// 0.47 warps per issue slot waiting for instructions (profiles/r2_summary.md).  1 = the replay of the queued
// heap operations is ONE loop body (findPosAll + bubbleUp inlined once, not kLaneChunk times): 29.8 KB;
// 2 = the L2 cache policies are created once and kept in registers instead of once per access: 26.4 KB.
// Shipped: F = 3 (-6 % per 1 M C4 queries).  Rolling the per-link cost tests too (21 KB) costs more in
// register selects than it saves (+18 %); a node DIRECTORY instead of the node table (records addressed by
// key in groups of 16, the directory L2-resident: -17 % DRAM bytes, bit-exact) was 10-16 % slower in every
// code shape -- the kernel waits for latency, not for DRAM bandwidth.  Both are in the history (2759383).
template <int HS, int TS, int CH, int V = 1, int F = 0>
struct LaneSearch {
  static constexpr bool kRolledReplay = (F & 1) != 0;
  static constexpr bool kHoistPolicy = (F & 2) != 0;
  static constexpr int kLaneChunk = CH;  // links handled per load stage
  // V = 1: every access at normal L2 priority, closed flag stored at every pop (round 1's kernel).
  // V >= 8: L2 residency by kind of data.  With every access at normal priority the 126 MB L2 keeps next
  // to nothing from one step to the next (75 k searches in flight x ~35 KB of live state each) and the
  // kernel runs at the DRAM's random-access rate (profiles/r2_summary.md).  Node records are write-once /
  // read-once-much-later: one 32 B access each, no L1 allocation, L2 evict_first.  Link records
  // (read-only, a few MB, read by every search): one 32 B load, evict_last.
  // V >= 9: no closed flag.  The flag only matters when a better path to an allocated node turns up
  // (DQ.cpp:1124-1153: open -> modify, closed -> push again); the modify scan that looks for the node in
  // the heap answers that, so the store of the flag at every pop (a DRAM write-back) is dropped.
  // V = 10 (shipped): + the node table's loads and stores carry evict_last (its working set per search is
  // a few KB thanks to the space-filling key order).
  // V = 20..24 (diagnostic build only): one kind of data tagged with an eviction class at a time, so that
  // ncu's per-class L2 counters read per kind.
  static constexpr bool kTabEvictLast = (V >= 10 && V < 20) || V == 20;
  static constexpr bool kRecStream = (V >= 8 && V < 20) || V == 23;
  static constexpr bool kLinkKeep = (V >= 8 && V < 20) || V == 22;
  static constexpr bool kNoClosedStore = V >= 9;
  static constexpr bool kHeapKeep = V == 21;
  static_assert((TS & 1) == 1, "TS must be odd");
  // memory of this lane
  float* K;        // shared: heap keys
  uint16_t* S;     // shared: heap nodes
  LaneHeapEnt* G;  // global: heap entries TS.. (entry j at G[j - TS])
  uint16_t* tab;   // node table
  char* rec;       // node records
  uint32_t* cv;    // corridor ring of the current query (entering links, see ViaCorridor)
  // query
  uint32_t q, endG;
  float ep[3];
  // search state
  int mode, size, nodeCount;
  uint32_t gen;
  uint32_t lastBest, lastBestG;
  float lastBestCost;
  bool outOfNodes;
  uint32_t expanded, nLinks, nNeigh;
  // result / extraction state
  uint32_t status;
  int xk;
  uint32_t xcur;
  LaneRecB xB;

  // L2 cache policies of the accesses below (device only)
  static HBN_HD unsigned long long polLast() {
    unsigned long long pol = 0;
#if defined(__CUDA_ARCH__)
    if constexpr (kHoistPolicy) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
    return pol;
  }
  static HBN_HD unsigned long long polFirst() {
    unsigned long long pol = 0;
#if defined(__CUDA_ARCH__)
    if constexpr (kHoistPolicy) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
    return pol;
  }

  // Node-table accesses.  V = 7 marks them evict_last in L2: with the keys numbered along the
  // space-filling curve the table working set of a search is a few KB, which is worth keeping
  // against the record stream (device only; a hint, the values are the same).
  HBN_HD uint32_t tabLoad(const uint32_t key) const {
#if defined(__CUDA_ARCH__)
    if constexpr (kTabEvictLast) {
      unsigned long long pol;
      unsigned short v;
      pol = polLast();
      asm volatile("ld.global.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(tab + key), "l"(pol) : "memory");
      return v;
    }
#endif
    return tab[key];
  }
  HBN_HD void tabStore(const uint32_t key, const uint32_t v) const {
#if defined(__CUDA_ARCH__)
    if constexpr (kTabEvictLast) {
      unsigned long long pol;
      pol = polLast();
      asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" ::"l"(tab + key), "h"(static_cast<unsigned short>(v)), "l"(pol) : "memory");
      return;
    }
#endif
    tab[key] = static_cast<uint16_t>(v);
  }
  // a looked-up table entry: found? / which node
  HBN_HD bool teFound(const uint32_t te) const { return (te >> kLaneSlotBits) == gen; }
  static HBN_HD uint32_t teSlot(const uint32_t te) { return te & kLaneSlotMask; }
  HBN_HD uint32_t lookup(const uint32_t key) const { return tabLoad(key); }
  HBN_HD void insert(const uint32_t key, const uint32_t slot) const { tabStore(key, (gen << kLaneSlotBits) | slot); }
  HBN_HD LaneRecA* recA(uint32_t s) const { return reinterpret_cast<LaneRecA*>(rec + static_cast<size_t>(s) * 32); }
  HBN_HD LaneRecB* recB(uint32_t s) const { return reinterpret_cast<LaneRecB*>(rec + static_cast<size_t>(s) * 32 + 16); }
  HBN_HD void loadRec(const uint32_t s, LaneRecA& a, LaneRecB& b) const {
#if defined(__CUDA_ARCH__)
    if constexpr (kRecStream) {
      uint32_t w[8];
      asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                   : "l"(rec + static_cast<size_t>(s) * 32)
                   : "memory");
      a = LaneRecA{__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3])};
      b = LaneRecB{w[4], w[5], w[6], w[7]};
      return;
    }
#endif
    a = *recA(s);
    b = *recB(s);
  }
  HBN_HD void storeRec(const uint32_t s, const LaneRecA& a, const LaneRecB& b) const {
#if defined(__CUDA_ARCH__)
    if constexpr (kRecStream) {
      asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(
                       rec + static_cast<size_t>(s) * 32),
                   "r"(__float_as_uint(a.px)), "r"(__float_as_uint(a.py)), "r"(__float_as_uint(a.pz)),
                   "r"(__float_as_uint(a.cost)), "r"(b.poly), "r"(b.w1), "r"(b.lnk), "r"(b.w3)
                   : "memory");
      return;
    }
#endif
    *recA(s) = a;
    *recB(s) = b;
  }
  HBN_HD LaneRecA loadRecA(const uint32_t s) const {  // what a revisit reads
#if defined(__CUDA_ARCH__)
    if constexpr (kRecStream) {
      unsigned long long pol;
      LaneRecA a;
      pol = polFirst();
      asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                   : "=f"(a.px), "=f"(a.py), "=f"(a.pz), "=f"(a.cost)
                   : "l"(rec + static_cast<size_t>(s) * 32), "l"(pol)
                   : "memory");
      return a;
    }
#endif
    return *recA(s);
  }
  HBN_HD LaneRecB loadRecB(const uint32_t s) const {  // corridor extraction
#if defined(__CUDA_ARCH__)
    if constexpr (kRecStream) {
      unsigned long long pol;
      LaneRecB b;
      pol = polFirst();
      asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                   : "=r"(b.poly), "=r"(b.w1), "=r"(b.lnk), "=r"(b.w3)
                   : "l"(rec + static_cast<size_t>(s) * 32 + 16), "l"(pol)
                   : "memory");
      return b;
    }
#endif
    return *recB(s);
  }
  HBN_HD void storeClosed(const uint32_t s, const float cost) const {  // the flag of DQ.cpp:1037-1038
    if constexpr (kNoClosedStore) return;
#if defined(__CUDA_ARCH__)
    if constexpr (kRecStream) {
      unsigned long long pol;
      pol = polFirst();
      asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(&recA(s)->cost),
                   "f"(laneSetClosed(cost)), "l"(pol)
                   : "memory");
      return;
    }
#endif
    recA(s)->cost = laneSetClosed(cost);
  }

  HBN_HD void hget(int i, float& k, uint32_t& s) const {
    if (i < TS) {
      k = K[i * HS];
      s = S[i * HS];
    } else {
      const LaneHeapEnt e = gLoad(&G[i - TS]);
      k = e.key;
      s = e.slot;
    }
  }
  // global part of the heap, optionally with an L2 evict_last policy (device only; same values)
  static HBN_HD LaneHeapEnt gLoad(const LaneHeapEnt* p) {
#if defined(__CUDA_ARCH__)
    if constexpr (kHeapKeep) {
      unsigned long long pol;
      LaneHeapEnt e;
      pol = polLast();
      asm volatile("ld.global.L2::cache_hint.v2.b32 {%0,%1}, [%2], %3;"
                   : "=r"(*reinterpret_cast<uint32_t*>(&e.key)), "=r"(e.slot)
                   : "l"(p), "l"(pol)
                   : "memory");
      return e;
    }
#endif
    return *p;
  }
  static HBN_HD void gStore(LaneHeapEnt* p, const float k, const uint32_t s) {
#if defined(__CUDA_ARCH__)
    if constexpr (kHeapKeep) {
      unsigned long long pol;
      pol = polLast();
      asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(__float_as_uint(k)), "r"(s),
                   "l"(pol)
                   : "memory");
      return;
    }
#endif
    *p = LaneHeapEnt{k, s};
  }
  HBN_HD void hset(int i, float k, uint32_t s) const {
    if (i < TS) {
      K[i * HS] = k;
      S[i * HS] = static_cast<uint16_t>(s);
    } else {
      gStore(&G[i - TS], k, s);
    }
  }
  static HBN_HD bool warpAny(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
  }
  // Heap position of open node `s` for every lane with `need` (dtNodeQueue::modify's search,
  // DNode.h:134-141); -1 if absent.  Warp collective.
  HBN_HD int findPosAll(const bool need, const uint32_t s) const {
#if defined(__CUDA_ARCH__)
    static_assert(HS == 32, "device build: one lane per column");
    const int lane = threadIdx.x & 31;
    int pos = -1;
    __syncwarp();  // the other lanes are about to read this lane's heap entries: order its stores first
    uint32_t m = __ballot_sync(0xffffffffu, need);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t tgt = __shfl_sync(0xffffffffu, s, l);
      const int n = __shfl_sync(0xffffffffu, size, l);
      const LaneHeapEnt* g = reinterpret_cast<const LaneHeapEnt*>(
          __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(G), l));
      const uint16_t* col = S - lane + l;  // lane l's column of the shared heap
      int hit = -1;
      for (int base = lane; base < n; base += 128) {  // 4 independent loads in flight per lane
        uint32_t hs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = base + 32 * u;
          hs[u] = 0xffffffffu;
          if (i < n) hs[u] = i < TS ? static_cast<uint32_t>(col[i * 32]) : g[i - TS].slot;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (hs[u] == tgt) hit = base + 32 * u;
      }
      __syncwarp();
      const uint32_t hm = __ballot_sync(0xffffffffu, hit >= 0);
      const int p = __shfl_sync(0xffffffffu, hit, hm ? __ffs(hm) - 1 : 0);
      if (lane == l) pos = hm ? p : -1;
    }
    return pos;
#else
    if (!need) return -1;
    for (int i = 0; i < size; ++i) {
      float k;
      uint32_t hs;
      hget(i, k, hs);
      if (hs == s) return i;
    }
    return -1;
#endif
  }

  // dtNodeQueue::bubbleUp, DNode.cpp:156-167
  HBN_HD void heapUp(int i, const float key, const uint32_t slot, const bool bubble = true) const {
    while (bubble && i > 0) {
      const int parent = (i - 1) >> 1;
      float pk;
      uint32_t ps;
      hget(parent, pk, ps);
      if (!(pk > key)) break;
      hset(i, pk, ps);
      i = parent;
    }
    hset(i, key, slot);
  }
  static HBN_HD LaneHeapPair loadPair(const LaneHeapEnt* p) {
#if defined(__CUDA_ARCH__)
    if constexpr (kHeapKeep) {
      unsigned long long pol;
      LaneHeapPair r;
      pol = polLast();
      asm volatile("ld.global.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
                   : "=r"(*reinterpret_cast<uint32_t*>(&r.a.key)), "=r"(r.a.slot),
                     "=r"(*reinterpret_cast<uint32_t*>(&r.b.key)), "=r"(r.b.slot)
                   : "l"(p), "l"(pol)
                   : "memory");
      return r;
    }
#endif
    return *reinterpret_cast<const LaneHeapPair*>(p);
  }
  // dtNodeQueue::pop's trickleDown (DNode.cpp:169-184) for a heap that has `n` entries left
  HBN_HD void heapPopSift(const int n) const {
    float lk;
    uint32_t ls;
    hget(n, lk, ls);
    int i = 0, child = 1;
    while (child < n) {  // child is odd; child + 1 <= n is inside the arrays
      float c0, c1;
      uint32_t s0, s1;
      if (child < TS) {
        c0 = K[child * HS];
        c1 = K[(child + 1) * HS];
        s0 = S[child * HS];
        s1 = S[(child + 1) * HS];
      } else {
        const LaneHeapPair p = loadPair(&G[child - TS]);
        c0 = p.a.key; s0 = p.a.slot;
        c1 = p.b.key; s1 = p.b.slot;
      }
      if ((child + 1) < n && c0 > c1) {
        c0 = c1;
        s0 = s1;
        child++;
      }
      hset(i, c0, s0);
      i = child;
      child = 2 * i + 1;
    }
    heapUp(i, lk, ls);
  }

  // DQ.cpp:1003-1021.  The caller has made sure gen < kLaneGenMax (table wiped otherwise).
  HBN_HD void begin(const NavView& nav, uint32_t query, uint32_t startG, const float* sp, uint32_t endPoly,
                    const float* endPos, uint32_t* corridorRing) {
    q = query;
    endG = endPoly;
    ep[0] = endPos[0]; ep[1] = endPos[1]; ep[2] = endPos[2];
    cv = corridorRing;
    const PolyRec* spoly = &nav.polys[startG];
    const uint32_t slnk = spoly->linkStart | (static_cast<uint32_t>(spoly->linkCount) << 27);
    const float stotal = vdist(sp, ep) * kHScale;
    gen++;
    storeRec(0, LaneRecA{sp[0], sp[1], sp[2], 0.f}, LaneRecB{startG, kLaneNoParent, slnk, 0u});
    insert(spoly->key0, 0u);
    hset(0, stotal, 0u);
    size = 1;
    nodeCount = 1;
    lastBest = 0;
    lastBestG = startG;
    lastBestCost = stotal;
    outOfNodes = false;
    expanded = nLinks = nNeigh = 0;
    status = 0;
    xk = 0;
    mode = kLSearch;
  }

  // DQ.cpp:1156-1164: status; the corridor is extracted when somebody will read it
  HBN_HD int finishSearch(bool allCorridors) {
    status = kDtSuccess;
    if (lastBestG != endG) status |= kDtPartialResult;
    if (outOfNodes) status |= kDtOutOfNodes;
    xk = 0;
    if (status == kDtSuccess || allCorridors) {
      mode = kLExtract;
      xcur = lastBest;
      xB = loadRecB(lastBest);
      return kLEvNone;
    }
    mode = kLIdle;
    return kLEvFinished;
  }

  // One neighbour (DQ.cpp:1056-1153) up to the heap operation, which is returned: 0 = none,
  // else kOpPush / kOpModify with the node in *opSlot and its new total in *opKey.  `te` and
  // `ra` were loaded before the serial part of this chunk.
  enum { kOpNone = 0, kOpPush = 1, kOpModify = 2 };
  HBN_HD int visit(const uint32_t bslot, const uint32_t bestG, const float* bpos, const float bcost,
                   const uint32_t viaJ, const LaneLinkLo& lo, const LaneLinkHi& hi, uint32_t te, LaneRecA ra,
                   const bool fastFail, int* stop, float* opKey, uint32_t* opSlot) {
    const uint32_t nei = lo.nei;
    if ((hi.meta & kLinkDupBit) != 0) {  // an earlier link of this poly may just have created the node
      te = lookup(hi.neiKey);
      if (teFound(te)) ra = loadRecA(teSlot(te));
    }
    const bool found = teFound(te);
    uint32_t slot;
    float npos[3];
    if (!found) {  // dtNodePool::getNode, DNode.cpp:121-152: allocation against the pool limit
      if (nodeCount >= kMaxNodes) {
        outOfNodes = true;
        if (fastFail) *stop = kLEvPoolExhausted;  // PF.cpp:1450 has decided "no path" already
        return kOpNone;
      }
      slot = static_cast<uint32_t>(nodeCount++);
      npos[0] = lo.mx; npos[1] = lo.my; npos[2] = lo.mz;
    } else {
      slot = teSlot(te);
      npos[0] = ra.px; npos[1] = ra.py; npos[2] = ra.pz;
    }
    // DQ.cpp:1088-1121
    const float curCost = vdist(bpos, npos);
    const float toEnd = vdist(npos, ep);
    float cost, heuristic;
    if (nei == endG) {
      cost = bcost + curCost + toEnd;
      heuristic = 0.f;
    } else {
      cost = bcost + curCost;
      heuristic = toEnd * kHScale;
    }
    const float total = cost + heuristic;
    // DQ.cpp:1124-1130: an allocated node is open or closed, and its total was formed as its
    // cost + the same heuristic
    if (found && total >= laneCost(ra.cost) + heuristic) return kOpNone;
    // without the closed flag every improved node is looked for in the heap; the replay pushes it
    // when it is not there (it was closed)
    const bool wasOpen = found && (kNoClosedStore || !laneIsClosed(ra.cost));
    storeRec(slot, LaneRecA{npos[0], npos[1], npos[2], cost},
             LaneRecB{nei, bestG | (viaJ << 24), hi.neiLinkStart | ((hi.meta >> kLinkNeiCountShift) << 27),
                      bslot | (1u << 12)});
    if (!found) insert(hi.neiKey, slot);
    if (heuristic < lastBestCost) {  // DQ.cpp:1154-1159
      lastBestCost = heuristic;
      lastBest = slot;
      lastBestG = nei;
    }
    *opKey = total;
    *opSlot = slot;
    return wasOpen ? kOpModify : kOpPush;
  }

  // One iteration of the state machine.  Returns an event; after kLEvFinished `status` and `xk`
  // (corridor length, 0 = not extracted) are the query's result and the lane is idle.
  HBN_HD int step(const NavView& nav, const bool fastFail, const bool allCorridors) {
    int ev = kLEvNone;
    bool go = false;  // this lane expands a poly in this step
    uint32_t bslot = 0, bestG = 0, parentG = 0, l0 = 0;
    int ln = 0;
    float bpos[3] = {0.f, 0.f, 0.f};
    float bcost = 0.f;
    float pk0 = 0.f, pk1 = 0.f;  // keys of the parents of heap positions size, size + 1 (see below)
    bool pkValid = false;
    int pushed = 0;
    if (mode == kLExtract) {  // getPathToNode, DQ.cpp:1167-1205, from the end backwards
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int h = 0; h < kLaneHops && mode == kLExtract; ++h) {
        const bool hasParent = ((xB.w3 >> 12) & 1u) != 0;
        uint32_t via = kNoPoly;
        if (hasParent) {
          const uint32_t ps = xB.w3 & 0xfffu;
          const LaneRecB pb = loadRecB(ps);
          via = (pb.lnk & 0x07ffffffu) + ((xB.w1 >> 24) & 31u);
          xB = pb;
          xcur = ps;
        }
        cv[(kMaxPathPolys - 1 - xk) & (kMaxPathPolys - 1)] = via;
        xk++;
        if (!hasParent) {
          if (xk > kMaxPathPolys) status |= kDtBufferTooSmall;
          mode = kLIdle;
          ev = kLEvFinished;
        } else if (xk > kMaxNodes) {  // a parent cycle would be a bug
          mode = kLIdle;
          ev = kLEvFault;
        }
      }
    } else if (mode == kLSearch) {
      if (size == 0) {
        ev = finishSearch(allCorridors);  // open list exhausted: partial result
      } else {
        // ---- pop (DQ.cpp:1027-1040) ------------------------------------------------------
        bslot = S[0];
        LaneRecA ba;
        LaneRecB bb;
        loadRec(bslot, ba, bb);
        size--;
#if defined(__CUDA_ARCH__)
        {  // the popped poly's link records are needed right after the sift: start them towards L1 now
          const char* lp = reinterpret_cast<const char*>(&nav.links[bb.lnk & 0x07ffffffu]);
          const int cnt = static_cast<int>(bb.lnk >> 27);
#pragma unroll
          for (int k = 0; k < CH; ++k)
            if (k < cnt) asm volatile("prefetch.global.L1 [%0];" ::"l"(lp + 32 * k));
        }
#endif
        heapPopSift(size);
        storeClosed(bslot, ba.cost);
#if defined(__CUDA_ARCH__)
        // The next pop takes the new top unless one of this poly's neighbours (whose records are
        // written below, so they are in L2) overtakes it: pull its record into L2 meanwhile.
        if (size > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + static_cast<size_t>(S[0]) * 32));
#endif
        bestG = bb.poly;
        if (bestG == endG) {
          lastBest = bslot;
          lastBestG = bestG;
          ev = finishSearch(allCorridors);
        } else if (expanded >= kLaneMaxExpansions) {
          mode = kLIdle;
          ev = kLEvFault;
        } else {
          go = true;
          parentG = bb.w1 & 0x00ffffffu;
          l0 = bb.lnk & 0x07ffffffu;
          ln = static_cast<int>(bb.lnk >> 27);
          expanded++;
          nLinks += static_cast<uint32_t>(ln);
          bpos[0] = ba.px; bpos[1] = ba.py; bpos[2] = ba.pz;
          bcost = ba.cost;  // open until this pop: sign bit clear
          // The first two pushes of this expansion go to positions size and size + 1; their
          // parents' keys are fetched now (beyond the shared levels that is an HBM access), so
          // the usual case -- a new node stays at the bottom -- costs no load in the replay.
          if (size > 0) {
            uint32_t dummy;
            hget((size - 1) >> 1, pk0, dummy);
            pk1 = pk0;
            if ((size >> 1) != ((size - 1) >> 1)) hget(size >> 1, pk1, dummy);
            pkValid = true;
          }
        }
      }
    }
    // ---- neighbours (DQ.cpp:1042-1153); the loop is warp-uniform (collectives inside) -----
    int stop = kLEvNone;
    for (int base = 0; warpAny(go && stop == kLEvNone && base < ln); base += kLaneChunk) {
      const bool act = go && stop == kLEvNone && base < ln;
      LaneLinkLo lo[kLaneChunk];
      LaneLinkHi hi[kLaneChunk];
      uint32_t te[kLaneChunk];
      LaneRecA ra[kLaneChunk];
      bool cand[kLaneChunk];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int k = 0; k < kLaneChunk; ++k) {
        lo[k].nei = kNoPoly;
        lo[k].mx = lo[k].my = lo[k].mz = 0.f;
        hi[k] = LaneLinkHi{0u, 0u, 0u, 0u};
        if (act && base + k < ln) laneLoadLink<kLinkKeep>(&nav.links[l0 + base + k], lo[k], hi[k]);
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int k = 0; k < kLaneChunk; ++k) {
        if (lo[k].nei != kNoPoly) nNeigh++;
        cand[k] = lo[k].nei != kNoPoly && lo[k].nei != parentG && (hi[k].meta & kLinkPassBit) != 0;
        te[k] = cand[k] ? tabLoad(hi[k].neiKey) : 0u;
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int k = 0; k < kLaneChunk; ++k) {
        ra[k] = LaneRecA{0.f, 0.f, 0.f, 0.f};
        if (cand[k] && teFound(te[k])) ra[k] = loadRecA(teSlot(te[k]));
      }
      // cost tests and record updates; the heap operations are queued in link order
      float qKey[kLaneChunk];
      uint32_t qSlot[kLaneChunk];  // node | modify << 16
      int nq = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int k = 0; k < kLaneChunk; ++k) {
        qKey[k] = 0.f;
        qSlot[k] = 0u;
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int k = 0; k < kLaneChunk; ++k) {
        if (cand[k] && stop == kLEvNone) {
          float key = 0.f;
          uint32_t slot = 0u;
          const int op = visit(bslot, bestG, bpos, bcost, static_cast<uint32_t>(base + k), lo[k], hi[k], te[k],
                               ra[k], fastFail, &stop, &key, &slot);
          if (op != kOpNone) {
            const uint32_t v = slot | (op == kOpModify ? 0x10000u : 0u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 0; j <= k; ++j)  // qKey[nq] = key with a compile-time register index
              if (j == nq) {
                qKey[j] = key;
                qSlot[j] = v;
              }
            nq++;
          }
        }
      }
      // replay: push = bubbleUp from the end, modify = locate + bubbleUp (DNode.h:118-142)
      if constexpr (kRolledReplay) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < kLaneChunk; ++j) {
          if (!warpAny(j < nq)) break;
          float key = qKey[0];
          uint32_t qs = qSlot[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
          for (int t = 1; t < kLaneChunk; ++t)  // queue entry j with compile-time register indices
            if (t == j) {
              key = qKey[t];
              qs = qSlot[t];
            }
          const bool mine = j < nq;
          const bool isModify = mine && (qs & 0x10000u) != 0;
          const uint32_t slot = qs & 0xffffu;
          const int hp = findPosAll(isModify, slot);
          if (mine) {
            int pos = hp;
            bool bubble = true;
            if (isModify && (hp >= 0 || !kNoClosedStore)) {
              if (hp < 0) stop = kLEvFault;  // an open node that is not in the heap would be a bug
              pkValid = false;
            } else {  // push (with kNoClosedStore also: the improved node was closed, DQ.cpp:1147-1152)
              // bubbleUp's first comparison against the prefetched parent key: valid as long as
              // no earlier operation of this expansion has moved an entry
              pos = size;
              bubble = !(pkValid && pushed < 2 && !((pushed == 0 ? pk0 : pk1) > key));
              if (bubble) pkValid = false;
              pushed++;
              size++;
            }
            if (pos >= 0) {
              heapUp(pos, key, slot, bubble);
            }
          }
        }
      } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < kLaneChunk; ++j) {
        if (!warpAny(j < nq)) break;
        const bool mine = j < nq;
        const bool isModify = mine && (qSlot[j] & 0x10000u) != 0;
        const uint32_t slot = qSlot[j] & 0xffffu;
        const int hp = findPosAll(isModify, slot);
        if (mine) {
          if (isModify && (hp >= 0 || !kNoClosedStore)) {
            if (hp < 0) stop = kLEvFault;  // an open node that is not in the heap would be a bug
            else heapUp(hp, qKey[j], slot);
            pkValid = false;
          } else {  // push (with kNoClosedStore also: the improved node was closed, DQ.cpp:1147-1152)
            // bubbleUp's first comparison against the prefetched parent key: valid as long as
            // no earlier operation of this expansion has moved an entry
            if (pkValid && pushed < 2 && !((pushed == 0 ? pk0 : pk1) > qKey[j])) {
              hset(size, qKey[j], slot);
            } else {
              heapUp(size, qKey[j], slot);
              pkValid = false;
            }
            pushed++;
            size++;
          }
        }
      }
      }
      pkValid = false;  // a further chunk of links starts from other positions
    }
    if (stop == kLEvPoolExhausted) {
      ev = finishSearch(allCorridors);
    } else if (stop == kLEvFault) {
      mode = kLIdle;
      ev = kLEvFault;
    }
    return ev;
  }
};

}  // namespace hbn

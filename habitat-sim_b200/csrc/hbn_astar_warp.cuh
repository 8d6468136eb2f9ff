// Warp-cooperative find_path: one warp per query.
//
// dtNavMeshQuery::findPath (DQ.cpp:973-1165) is a serial algorithm whose result depends on
// the exact order of its heap operations (ties between equal `total`s are broken by the heap
// layout: bubbleUp DNode.cpp:156-167, bottom-up trickleDown :169-184).  Measured on the C4
// workload one query in four pops at least one tied key, so the open list here IS that binary
// heap, operated by lane 0 in the reference's order.  Everything around it is spread over the
// warp:
//   * the popped node's links are fetched by lanes 0..n-1 (one 32 B LinkRec each) together
//     with its node record, while lane 0 is still sifting the heap (the link window travels
//     in the heap entry, so nothing waits for the record);
//   * every neighbour is evaluated by its own lane: node lookup in the shared-memory table,
//     record fetch, the two dtVdist square roots, the open/closed tests;
//   * node allocation order (the 2048-node limit of PF.cpp:937 is part of the result) is a
//     ballot prefix; pushes / modifies are then replayed in link order.
// Node storage: a 2560-slot open-addressing table of 32-bit keys in shared memory (slot ==
// node id) and one 32 B record per slot in an L2-resident per-warp scratch.  `modify` finds
// the heap position with a warp-wide scan instead of keeping back pointers.
#pragma once
#include <cuda_runtime.h>
#include "hbn_query.h"

namespace hbn {

constexpr int kTabSize = 2560;  // node table slots: load <= 0.8 at the 2048-node limit
__device__ __forceinline__ uint32_t tabHome(uint32_t key) { return __umulhi(nodeHash(key), kTabSize); }
__device__ __forceinline__ uint32_t tabNext(uint32_t slot) { return slot + 1 == kTabSize ? 0u : slot + 1; }
constexpr uint32_t kEntValid = 1u << 31;  // entry = valid | closed | open | state << 24 | g
constexpr uint32_t kFullMask = 0xffffffffu;
constexpr uint32_t kSearchOverflow = 0xffffffffu;  // heap tier too small: rerun in the next tier
constexpr uint32_t kSearchWatchdog = 0xfffffffeu;  // iteration cap hit: a bug, reported to the host
constexpr uint32_t kMaxExpansions = 1u << 18;      // >> any legal search (2048 nodes, re-opens)

struct __align__(16) NodeRec {
  float px, py, pz, cost;
  float heur;     // heuristic term of the node's total (a function of its fixed position only)
  uint32_t lnk;   // link window: start (27 bits) | count << 27
  uint32_t via;   // LinkRec index this node was (last) entered through; kNoPoly for the start
  uint32_t pidx;  // parent slot + 1, 0 = none
};
static_assert(sizeof(NodeRec) == 32, "NodeRec");

struct __align__(16) HeapEnt {
  float key;
  uint32_t slot;
  uint32_t lnk;   // the node's link window, so that a pop can fetch the links at once
  uint32_t pidx;  // the node's parent slot + 1 (0 = none), for the "skip the parent" test
};

// Shared memory of one warp: [tab: kTabSize u32][heap: OC + 2 entries].  The heap is stored
// shifted by one entry so that the two children of logical i (2i+1, 2i+2) share a 32 B line.
template <int OC>
struct WarpWs {
  uint32_t* tab;
  HeapEnt* heap;  // logical index 0 is heap[0] of this pointer (already shifted)
  NodeRec* rec;
  __host__ __device__ static constexpr size_t sharedBytes() { return kTabSize * 4 + (OC + 2) * sizeof(HeapEnt); }
  __host__ __device__ static constexpr size_t globalBytes() { return static_cast<size_t>(kTabSize) * sizeof(NodeRec); }
  __device__ static WarpWs carve(char* sm, char* gl) {
    WarpWs w;
    w.tab = reinterpret_cast<uint32_t*>(sm);
    w.heap = reinterpret_cast<HeapEnt*>(sm + kTabSize * 4) + 1;
    w.rec = reinterpret_cast<NodeRec*>(gl);
    return w;
  }
};

// dtNodeQueue::bubbleUp, DNode.cpp:156-167
__device__ __forceinline__ void heapUp(HeapEnt* hp, int i, const HeapEnt node) {
  while (i > 0) {
    const int parent = (i - 1) >> 1;
    const HeapEnt p = hp[parent];
    if (!(p.key > node.key)) break;
    hp[i] = p;
    i = parent;
  }
  hp[i] = node;
}
// dtNodeQueue::pop's trickleDown (DNode.cpp:169-184) for a heap that has `n` entries left
__device__ __forceinline__ void heapPopSift(HeapEnt* hp, int n) {
  const HeapEnt last = hp[n];
  int i = 0, child = 1;
  while (child < n) {
    HeapEnt c0 = hp[child];
    const HeapEnt c1 = hp[child + 1];
    if ((child + 1) < n && c0.key > c1.key) {
      c0 = c1;
      child++;
    }
    hp[i] = c0;
    i = child;
    child = 2 * i + 1;
  }
  heapUp(hp, i, last);
}

struct WarpSearch {
  uint32_t status;  // Detour status word of findPath, or kSearchOverflow
  uint32_t lastBest;
  int nodeCount;
  uint32_t expanded, links, neighbours;
};

// All 32 lanes call this converged.  sp/ep: snapped start / end (PF.cpp:1448).
template <int OC>
__device__ WarpSearch astarWarp(const NavView& nav, const WarpWs<OC>& ws, uint32_t startG,
                                uint32_t endG, const float* sp, const float* ep, bool fastFail) {
  const int lane = threadIdx.x & 31;
  const uint32_t ltMask = (1u << lane) - 1u;
  HeapEnt* hp = ws.heap;
  WarpSearch r;
  r.status = kDtSuccess;
  r.expanded = r.links = r.neighbours = 0;

  uint4* t4 = reinterpret_cast<uint4*>(ws.tab);
  for (int i = lane; i < kTabSize / 4; i += 32) t4[i] = make_uint4(0u, 0u, 0u, 0u);
  const PolyRec* spoly = &nav.polys[startG];
  const uint32_t slnk = spoly->linkStart | (static_cast<uint32_t>(spoly->linkCount) << 27);
  const uint32_t sslot = tabHome(startG);
  const float stotal = vdist(sp, ep) * kHScale;
  __syncwarp();
  if (lane == 0) {
    ws.tab[sslot] = kEntValid | kNodeOpen | startG;
    float4* ra = reinterpret_cast<float4*>(&ws.rec[sslot]);
    ra[0] = make_float4(sp[0], sp[1], sp[2], 0.f);
    reinterpret_cast<uint4*>(ra)[1] = make_uint4(__float_as_uint(stotal), slnk, kNoPoly, 0u);  // heur = total (cost 0)
    hp[0] = HeapEnt{stotal, sslot, slnk, 0u};
  }
  __syncwarp();
  int size = 1;
  int nodeCount = 1;
  uint32_t lastBest = sslot;
  float lastBestCost = stotal;
  bool outOfNodes = false;

  while (size > 0) {
    const HeapEnt top = hp[0];
    const uint32_t bslot = top.slot;
    const uint32_t l0 = top.lnk & 0x07ffffffu;
    const int ln = static_cast<int>(top.lnk >> 27);
    // early loads: the node's record (broadcast) and one link per lane
    const float4 ba = reinterpret_cast<const float4*>(&ws.rec[bslot])[0];
    const uint32_t bpidx = top.pidx;
    uint4 La = make_uint4(0u, 0u, 0u, kNoPoly), Lb = make_uint4(0u, 0u, 0u, 0u);
    if (lane < ln) {
      const uint4* lp = reinterpret_cast<const uint4*>(&nav.links[l0 + lane]);
      La = __ldg(lp);
      Lb = __ldg(lp + 1);
    }
    size--;
    __syncwarp();  // every lane has read hp[0] before lane 0 rewrites the heap
    if (lane == 0) heapPopSift(hp, size);
    const uint32_t bent = ws.tab[bslot];
    const uint32_t bestG = bent & kNodeGMask;
    __syncwarp();
    if (lane == 0) ws.tab[bslot] = (bent & ~kNodeOpen) | kNodeClosed;
    if (bestG == endG) {
      lastBest = bslot;
      break;
    }
    const uint32_t parentG = bpidx ? (ws.tab[bpidx - 1] & kNodeGMask) : kNoPoly;
    const float bpos[3] = {ba.x, ba.y, ba.z};
    const float bcost = ba.w;
    __syncwarp();

    const uint32_t nei = La.w;
    const uint32_t meta = Lb.y;
    if (r.expanded >= kMaxExpansions) {
      r.status = kSearchWatchdog;
      return r;
    }
    r.expanded++;
    r.links += ln;
    r.neighbours += __popc(__ballot_sync(kFullMask, nei != kNoPoly));
    const bool cand = nei != kNoPoly && nei != parentG && (meta & kLinkPassBit) != 0;
    const uint32_t key = nei | (((meta >> kLinkStateShift) & 3u) << 24);
    uint32_t pend = __ballot_sync(kFullMask, cand);

    // Two links of one poly can lead to the same neighbour (flagged at flatten time): such a
    // link must see its twin's node, so a round ends in front of it.  Usually one round.
    const uint32_t dupLanes = __ballot_sync(kFullMask, cand && (meta & kLinkDupBit) != 0);
    while (pend) {
      uint32_t cur = pend;
      if (dupLanes) {
        const int first = __ffs(pend) - 1;
        const uint32_t later = dupLanes & pend & ~((2u << first) - 1u);
        if (later) cur = pend & ((1u << (__ffs(later) - 1)) - 1u);
      }
      pend &= ~cur;
      const bool mine = (cur >> lane) & 1u;

      // dtNodePool::getNode, DNode.cpp:121-152: lookup ...
      uint32_t slot = tabHome(key);
      uint32_t ent = 0;
      bool found = false;
      bool bad = false;
      if (mine) {
        for (int probes = 0;; ++probes) {
          ent = ws.tab[slot];
          if (ent == 0) break;
          if ((ent & kNodeKeyMask) == key) {
            found = true;
            break;
          }
          if (probes >= kTabSize) {
            bad = true;
            break;
          }
          slot = tabNext(slot);
        }
      }
      if (__any_sync(kFullMask, bad)) {
        r.status = kSearchWatchdog;
        r.expanded |= 0x40000000u;
        return r;
      }
      __syncwarp();  // lookups done before this round's inserts / flag updates
      // ... allocation in link order against the 2048-node limit
      const bool isNew = mine && !found;
      const uint32_t newMask = __ballot_sync(kFullMask, isNew);
      const bool allocFail = isNew && (nodeCount + __popc(newMask & ltMask)) >= kMaxNodes;
      const uint32_t failMask = __ballot_sync(kFullMask, allocFail);
      if (failMask) {
        outOfNodes = true;
        if (fastFail) goto done;
      }
      nodeCount += __popc(newMask & ~failMask);
      const bool ok = mine && !allocFail;
      if (ok && isNew) {
        for (int probes = 0; probes < 2 * kTabSize; ++probes) {  // distinct keys in a round: only the slot can be contended
          if (atomicCAS(&ws.tab[slot], 0u, kEntValid | key) == 0u) break;
          slot = tabNext(slot);
        }
      }
      float npos[3] = {__uint_as_float(La.x), __uint_as_float(La.y), __uint_as_float(La.z)};
      float ntotal = 0.f, nheur = 0.f;
      uint32_t nlnk = Lb.x | ((meta >> kLinkNeiCountShift) << 27);
      if (ok && found) {
        const float4 na = reinterpret_cast<const float4*>(&ws.rec[slot])[0];
        const uint2 nb = reinterpret_cast<const uint2*>(&ws.rec[slot])[2];
        npos[0] = na.x; npos[1] = na.y; npos[2] = na.z;
        nheur = __uint_as_float(nb.x);
        ntotal = na.w + nheur;  // the node's total, as it was formed: cost + heuristic
        nlnk = nb.y;
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
      const bool wasOpen = found && (ent & kNodeOpen) != 0;
      const bool wasClosed = found && (ent & kNodeClosed) != 0;
      const bool acc = ok && !((wasOpen || wasClosed) && total >= ntotal);  // DQ.cpp:1124-1130
      if (acc) {
        float4* ra = reinterpret_cast<float4*>(&ws.rec[slot]);
        ra[0] = make_float4(npos[0], npos[1], npos[2], cost);
        reinterpret_cast<uint4*>(ra)[1] =
            make_uint4(__float_as_uint(heuristic), nlnk, l0 + static_cast<uint32_t>(lane), bslot + 1u);
        ws.tab[slot] = kEntValid | key | kNodeOpen;
      }
      const uint32_t accMask = __ballot_sync(kFullMask, acc);
      __syncwarp();
      // heap updates replayed in link order (DQ.cpp:1140-1152)
      const uint32_t openMask = __ballot_sync(kFullMask, acc && wasOpen);
      const HeapEnt myEnt{total, slot, nlnk, bslot + 1u};
      uint32_t m = accMask;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        if ((openMask >> b) & 1u) {  // dtNodeQueue::modify, DNode.h:132-142: locate, then bubbleUp
          const uint32_t s = __shfl_sync(kFullMask, slot, b);
          int pos = -1;
          for (int i = lane; i < size; i += 32)
            if (hp[i].slot == s) pos = i;
          const uint32_t pm = __ballot_sync(kFullMask, pos >= 0);
          pos = __shfl_sync(kFullMask, pos, pm ? (__ffs(pm) - 1) : 0);
          __syncwarp();
          if (lane == b && pos >= 0) heapUp(hp, pos, myEnt);
        } else {
          if (size >= OC) {
            r.status = kSearchOverflow;
            return r;
          }
          if (lane == b) heapUp(hp, size, myEnt);
          size++;
        }
        __syncwarp();
      }
      // DQ.cpp:1154-1159: first neighbour (link order) with the smallest heuristic
      if (accMask) {  // heuristics are >= +0, so their bit patterns order like the floats
        const uint32_t hb = acc ? __float_as_uint(heuristic) : 0xffffffffu;
        const uint32_t mn = __reduce_min_sync(kFullMask, hb);
        if (__uint_as_float(mn) < lastBestCost) {
          const int hl = __ffs(__ballot_sync(kFullMask, hb == mn)) - 1;
          lastBestCost = __uint_as_float(mn);
          lastBest = __shfl_sync(kFullMask, slot, hl);
        }
      }
    }
  }
done:
  __syncwarp();
  r.lastBest = lastBest;
  r.nodeCount = nodeCount;
  if ((ws.tab[lastBest] & kNodeGMask) != endG) r.status |= kDtPartialResult;
  if (outOfNodes) r.status |= kDtOutOfNodes;
  return r;
}

}  // namespace hbn

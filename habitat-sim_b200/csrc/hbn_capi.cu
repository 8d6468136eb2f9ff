// C ABI of libhbn.so (include/hbn.h): navmesh upload, kernel orchestration, host-buffer
// wrappers.  No CPU query path exists in this library.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hbn.h"
#include "hbn_host.h"
#include "hbn_kernels.cuh"
#include "hbn_findpath.cuh"
#include "hbn_astar_lane.cuh"
#include "hbn_snap.cuh"
#include "hbn_follower.cuh"

using namespace hbn;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(expr)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (expr);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(HBN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
  } while (0)

// tier configuration (see DESIGN.md "A* workspace tiers")
constexpr int kCapL = 2048;   // reference pool size: heap+hash shared, nodes in L2
constexpr int kWallCapS = 128;
constexpr int kFpWarps = 4;   // warps per block of the wall-distance kernels
// per chunk of a find_path call: 16 counters, then the class histogram and the scatter cursors (k_fp_classify)
constexpr size_t kFpCounterBytes = (16 + 2 * 32) * 4;
constexpr int kLaneTS = 71;    // lane-per-query search: heap entries per lane kept in shared memory (6 levels + 8)
constexpr int kLaneWpb = 2;    // ... warps per block
constexpr int kLaneMinB = 16;  // ... and resident warps per SM its register budget must allow
constexpr int kLaneF = 3;     // ... and its code shape: one replay loop body, L2 policies created once (hbn_astar_lane.h)
constexpr int kLaneV = 10;     // ... and its code variant: L2 residency by kind of data, no closed flag (hbn_astar_lane.h)
constexpr int kSnapW = 8;     // lanes per point in k_snap
constexpr int kRandW = 8;

// Scratch buffer that grows on demand.  Growing frees and allocates (a device-wide implicit
// synchronisation, illegal during stream capture): hbn_navmesh_reserve sizes every buffer up front so
// that the *_dev entry points only enqueue.  `epoch` counts reallocations of any buffer of the process
// (cached CUDA graphs bake the pointers in and are rebuilt when it moves).
uint64_t g_scratchEpoch = 0;
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes, bool slack = true) {
    if (bytes <= cap) return HBN_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    g_scratchEpoch++;
    const size_t want = slack ? bytes + bytes / 4 + 256 : bytes;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(HBN_ERR_CUDA, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
    }
    cap = want;
    return HBN_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    g_scratchEpoch++;
  }
};

// Tuning knobs of a handle.  Defaults come from the environment ONCE, at creation (HBN_LANE_CFG,
// HBN_LANE_SPREAD, HBN_SNAP_SPREAD, HBN_SNAP_DUAL, HBN_SNAP_GROUP, HBN_SNAP_CAP, HBN_FP_BLOCKS_PER_SM,
// HBN_LANE_SCRATCH_MB); hbn_navmesh_set_option changes them afterwards.  Nothing on a query path reads
// the environment.
struct Options {
  int laneCfg = 0;          // k_astar_lane instantiation (0 = shipped)
  int laneSpread = 1;       // batches smaller than the grid use fewer lanes per warp
  int snapSpread = 1;       // small snap batches: one lane group per warp
  int snapDual = 1;         // independent small snap batches share a launch
  int snapGroup = 0;        // testing: the lane-group snap kernel for every batch size
  int64_t snapCap = 0;      // testing: candidate scratch entries of the snap pipeline (0 = default)
  int blocksPerSm = 0;      // cap on resident one-warp blocks per SM of the search (0 = occupancy)
  int64_t laneScratchBytes = 0;  // cap on the per-lane search state in HBM (0 = half of the free memory)
  int nvtx = 1;             // NVTX range around every batched call
};

struct NvtxRange {
  bool on;
  NvtxRange(bool enable, const char* name) : on(enable) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

}  // namespace

struct hbn_navmesh {
  int device = 0;
  HostNavMesh host;  // the tiles as ingested + finished (kept for hbn_navmesh_save_mset)
  FlatNav flat;
  NavView view{};
  Options opt;
  std::vector<void*> devArrays;
  int64_t deviceBytes = 0;
  cudaStream_t stream = nullptr;  // for the host-buffer entry points
  int smCount = 0;
  int64_t launches = 0;
  std::recursive_mutex mu;
  // All scratch belongs to the handle, so calls on different streams must not overlap: every *_dev
  // entry point makes its stream wait for the event the previous call recorded on ITS stream (a no-op
  // for the same stream) and records a new one when it has enqueued its work.
  cudaEvent_t lastDone = nullptr;
  cudaStream_t lastStream = nullptr;
  bool lastValid = false;
  // scratch (device)
  DevBuf sG, eG, e2G, sPt, ePt, epPt, e2Pt, lastPoly, lists, counters, wsL, io, work, mgDist, mgBounds, mgOrder, mgEnd, mgMask;
  DevBuf folS, folT, folE, folF, folGeo, folObs, folGeo0, folPos;  // batched follower: primitives of all agents
  DevBuf envG, envPt, envFlag, envPos;  // env step: find_path's start projection, fix-up flags, the goals' polys
  // find_path pipeline: class, cost class, search list, status, corridor rings
  DevBuf fpCls, fpBucket, fpWork, fpStat, fpLen, fpCorr;
  DevBuf snapWin, snapG, snapQ, snapD, snapOut, snapBest, snapTodo;  // candidate-list snap (hbn_snap.cuh)
  // lane-per-query search: per-lane node table + records + heap tail in HBM, sized from the batch
  DevBuf wsLane, laneGen;
  int64_t laneSlots = 0;    // lane slots the scratch holds (tables zeroed, generations 0 when allocated)
  int blocksFpLane = 0;     // resident WARPS of the search kernel (occupancy x SMs); the grid unit of laneGrid / laneScratch
  // pinned staging for the host-buffer entry points
  void* pinned = nullptr;
  size_t pinnedCap = 0;
  int blocksWallS = 0, blocksWallL = 0;
  unsigned int* faultHost = nullptr;  // mapped pinned memory the kernels report bugs through
  unsigned int* faultDev = nullptr;
  // optional phase timing of hbn_find_path_dev (hbn_navmesh_set_profiling)
  bool profile = false;
  struct PhaseEv { cudaEvent_t e[3]; };
  std::vector<PhaseEv> phaseEvents;
  // cached CUDA graphs of the host-buffer env step, keyed by (n, allow_sliding)
  struct StepGraph { int64_t n; int sliding; uint64_t epoch; cudaGraphExec_t exec; size_t offs[5]; };
  std::vector<StepGraph> stepGraphs;
};

namespace {

struct DeviceGuard {
  int prev = 0;
  bool ok;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

template <class T>
int upload(hbn_navmesh* nm, const std::vector<T>& v, const T** out) {
  void* d = nullptr;
  const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  CK(cudaMalloc(&d, bytes));
  nm->devArrays.push_back(d);
  nm->deviceBytes += static_cast<int64_t>(bytes);
  if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T*>(d);
  return HBN_OK;
}

// Per-lane search state of k_astar_lane (node table + node records + heap tail per lane slot) for a
// grid of `blocks` one-warp blocks.  Sized from what the batch needs, grown on demand (new tables are
// zeroed, generations reset), capped by Options::laneScratchBytes or half of the free device memory:
// under the cap the grid shrinks and every lane serves more queries.  On any failure both buffers are
// released, so a later call starts from a clean state.
int laneScratch(hbn_navmesh* nm, cudaStream_t st, int* blocks, int lanesPerBlock, LaneScratch* out) {
  const size_t tabB = laneTabBytes(nm->view.numKeys);
  const size_t per = laneScratchBytes(nm->view.numKeys);
  const auto slotsFor = [&](int64_t b) { return (b * lanesPerBlock + 31) / 32 * 32; };
  int64_t want = slotsFor(*blocks);
  if (want > nm->laneSlots) {
    size_t budget = nm->opt.laneScratchBytes > 0 ? static_cast<size_t>(nm->opt.laneScratchBytes) : 0;
    if (!budget) {
      size_t freeB = 0, totalB = 0;
      CK(cudaMemGetInfo(&freeB, &totalB));
      budget = (freeB + nm->wsLane.cap) / 2;
    }
    const int64_t maxSlots = static_cast<int64_t>(budget / per) / 32 * 32;
    if (maxSlots < 32) return fail(HBN_ERR_CUDA, "not enough device memory for the find_path search state");
    if (want > maxSlots) want = maxSlots;
    if (want > nm->laneSlots) {
      nm->laneSlots = 0;
      int rc;
      if ((rc = nm->wsLane.ensure(static_cast<size_t>(want) * per, false)) ||
          (rc = nm->laneGen.ensure(static_cast<size_t>(want) * 4, false))) {
        nm->wsLane.release();
        nm->laneGen.release();
        return rc;
      }
      if (cudaMemsetAsync(nm->wsLane.p, 0, static_cast<size_t>(want) * tabB, st) != cudaSuccess ||  // the tables come first
          cudaMemsetAsync(nm->laneGen.p, 0, static_cast<size_t>(want) * 4, st) != cudaSuccess) {
        nm->wsLane.release();
        nm->laneGen.release();
        return fail(HBN_ERR_CUDA, "cudaMemsetAsync(search state)");
      }
      nm->laneSlots = want;
    }
  }
  if (slotsFor(*blocks) > nm->laneSlots) *blocks = static_cast<int>(std::max<int64_t>(1, nm->laneSlots / lanesPerBlock));
  const size_t lanes = static_cast<size_t>(nm->laneSlots);
  char* p = static_cast<char*>(nm->wsLane.p);
  out->tab = p;
  out->rec = p + lanes * tabB;
  out->heap = out->rec + lanes * kLaneRecBytes;
  out->gen = static_cast<uint32_t*>(nm->laneGen.p);
  out->tabBytes = tabB;
  return HBN_OK;
}

// lane-per-query search instantiations (Options::laneCfg; 0 = shipped).  Template arguments: heap
// entries in shared memory, resident warps per SM, links per load stage, code variant V of LaneSearch.
// All are bit-exact (tests/test_zz_tuning_variants.py); the measurements that picked the shipped one are in
// profiles/r2_summary.md (and profiles/r1_summary.md for the variants deleted since).
const void* laneKernel(int cfg, size_t* shared, int* wpb) {
  *wpb = 1;
#define HBN_LANE_CASE(N, TS, W, ...) \
  case N: *shared = laneSharedBytes<TS>() * W; *wpb = W; return reinterpret_cast<const void*>(&__VA_ARGS__);
  switch (cfg) {
    // the steps from round 1's kernel to the shipped one (ms per 1 M C4 queries, profiles/r2_summary.md)
    HBN_LANE_CASE(1, 63, 1, k_astar_lane<63, 16, 4, 1>)       // round 1: normal L2 priority everywhere, closed flag: 145.7
    HBN_LANE_CASE(30, 63, 1, k_astar_lane<63, 16, 4, 8>)      // records evict_first (32 B accesses), links evict_last: 146.8
    HBN_LANE_CASE(31, 63, 1, k_astar_lane<63, 16, 4, 9>)      // + no closed-flag store: 129.7
    HBN_LANE_CASE(32, 63, 1, k_astar_lane<63, 16, 4, 10>)     // + node table evict_last: 126.2
    HBN_LANE_CASE(34, 95, 1, k_astar_lane<95, 11, 4, 10>)     // 95 shared heap entries at 11 warps per SM: 139.3
    // two-warp blocks (a block costs 1 KB of reserved shared memory) leave room for 71 heap entries per lane
    // at 16 warps per SM: 122.5
    HBN_LANE_CASE(38, 71, 2, k_astar_lane<71, 16, 4, 10, 2>)
    HBN_LANE_CASE(60, 71, 2, k_astar_lane<71, 16, 4, 10, 2, 1>)  // + one replay loop body (40.6 -> 29.8 KB of SASS): 118.0
#ifdef HBN_DIAG_KERNELS  // one kind of data tagged at a time: ncu's per-eviction-class L2 counters then read per kind
    HBN_LANE_CASE(40, 63, 1, k_astar_lane<63, 16, 4, 20>)  // node table
    HBN_LANE_CASE(41, 63, 1, k_astar_lane<63, 16, 4, 21>)  // heap tail
    HBN_LANE_CASE(42, 63, 1, k_astar_lane<63, 16, 4, 22>)  // link records
    HBN_LANE_CASE(43, 63, 1, k_astar_lane<63, 16, 4, 23>)  // node records
    HBN_LANE_CASE(44, 63, 1, k_astar_lane<63, 16, 4, 24>)  // nothing tagged (no closed-flag store only)
#endif
    // shipped: + the L2 policies created once (26.9 KB of SASS, under the 32 KB instruction cache level): 115.2
    default: *shared = laneSharedBytes<kLaneTS>() * kLaneWpb; *wpb = kLaneWpb;
             return reinterpret_cast<const void*>(&k_astar_lane<kLaneTS, kLaneMinB, 4, kLaneV, kLaneWpb, kLaneF>);
  }
#undef HBN_LANE_CASE
}

int envInt(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// (re)select the search kernel of Options::laneCfg: shared-memory opt-in, occupancy, grid size
int laneConfigure(hbn_navmesh* nm) {
  size_t smLane = 0;
  int wpb = 1;
  const void* fn = laneKernel(nm->opt.laneCfg, &smLane, &wpb);
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smLane)));
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32 * wpb, smLane));
  occ = std::max(1, occ) * wpb;  // resident warps per SM
  if (nm->opt.blocksPerSm > 0) occ = std::min(occ, nm->opt.blocksPerSm);
  nm->blocksFpLane = occ * nm->smCount;
  return HBN_OK;
}

int finishCreate(HostNavMesh& meshIn, const int32_t* islands, const float* radii, int numRadii, int device,
                 hbn_navmesh_t* out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(HBN_ERR_NO_DEVICE, "no CUDA device available (libhbn has no CPU path)");
  if (device < 0 || device >= ndev) return fail(HBN_ERR_INVALID, "bad device index");
  DeviceGuard devGuard(device);  // the caller's current device is restored on every return path
  if (!devGuard.ok) return fail(HBN_ERR_CUDA, "cudaSetDevice failed");
  // every failure below releases the handle and what was uploaded so far
  struct Cleanup {
    hbn_navmesh* nm;
    ~Cleanup() { if (nm) hbn_navmesh_destroy(nm); }
  } guard{new hbn_navmesh()};
  hbn_navmesh* nm = guard.nm;
  nm->device = device;
  nm->host = std::move(meshIn);
  HostNavMesh& mesh = nm->host;
  mesh.finish(islands, radii, numRadii);
  mesh.flatten(nm->flat);
  const FlatNav& f = nm->flat;
  if (f.polys.size() >= (1u << 24) || f.links.size() >= (1u << 27))
    return fail(HBN_ERR_LIMIT, "navmesh exceeds 2^24 polys or 2^27 links");
  for (const PolyRec& p : f.polys)
    if (p.linkCount > 31) return fail(HBN_ERR_LIMIT, "a polygon has more than 31 links");
  NavView v = f.view();
  int rc = HBN_OK;
  auto up = [&](auto& vec, auto** dst) { if (rc == HBN_OK) rc = upload(nm, vec, dst); };
  up(f.polys, &v.polys);
  up(f.links, &v.links);
  up(f.portals, &v.portals);
  up(f.bv, &v.bv);
  up(f.tiles, &v.tiles);
  up(f.detTris, &v.detTris);
  up(f.detVerts, &v.detVerts);
  up(f.polyBox, &v.polyBox);
  up(f.gridStart, &v.gridStart);
  up(f.tileOrder, &v.tileOrder);
  up(f.randEntries, &v.randEntries);
  up(f.tileIslStart, &v.tileIslStart);
  up(f.tileIslId, &v.tileIslId);
  up(f.tileIslWin, &v.tileIslWin);
  up(f.tileIslCnt, &v.tileIslCnt);
  if (rc != HBN_OK) return rc;
  nm->view = v;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  nm->smCount = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&nm->stream, cudaStreamNonBlocking));

  // knobs: environment defaults, read once
  nm->opt.laneCfg = envInt("HBN_LANE_CFG", 0);
  nm->opt.laneSpread = envInt("HBN_LANE_SPREAD", 1);
  nm->opt.snapSpread = envInt("HBN_SNAP_SPREAD", 1);
  nm->opt.snapDual = envInt("HBN_SNAP_DUAL", 1);
  nm->opt.snapGroup = getenv("HBN_SNAP_GROUP") ? 1 : 0;
  nm->opt.snapCap = envInt("HBN_SNAP_CAP", 0);
  nm->opt.blocksPerSm = envInt("HBN_FP_BLOCKS_PER_SM", 0);
  nm->opt.laneScratchBytes = static_cast<int64_t>(envInt("HBN_LANE_SCRATCH_MB", 0)) << 20;
  nm->opt.nvtx = envInt("HBN_NVTX", 1);
  CK(cudaEventCreateWithFlags(&nm->lastDone, cudaEventDisableTiming));

  // opt in to large dynamic shared memory and size the persistent grids from occupancy
  const int threads = kFpWarps * 32;
  const size_t smL = kFpWarps * wsSharedBytes(kCapL, kWsHybrid);
  const size_t smW = kFpWarps * wsSharedBytes(kWallCapS, kWsShared);
  CK(cudaFuncSetAttribute(k_wall<kWallCapS, kWsShared>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smW)));
  CK(cudaFuncSetAttribute(k_wall<kCapL, kWsHybrid>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smL)));
  int occ = 0;
  int rcLane = HBN_OK;
  if ((rcLane = laneConfigure(nm))) return rcLane;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wall<kWallCapS, kWsShared>, threads, smW));
  nm->blocksWallS = std::max(1, occ) * nm->smCount;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wall<kCapL, kWsHybrid>, threads, smL));
  nm->blocksWallL = std::max(1, occ) * nm->smCount;
  // global scratch: one slot per resident warp
  rc = nm->counters.ensure(64);
  if (rc == HBN_OK) rc = nm->work.ensure(64);
  if (rc == HBN_OK) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&nm->faultHost), 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&nm->faultDev), nm->faultHost, 0) != cudaSuccess)
      rc = fail(HBN_ERR_CUDA, "cudaHostAlloc(fault flags)");
    else
      memset(nm->faultHost, 0, 64);
  }
  if (rc == HBN_OK && cudaMemset(nm->work.p, 0, 64) != cudaSuccess) rc = fail(HBN_ERR_CUDA, "cudaMemset");
  if (rc != HBN_OK) return rc;
  guard.nm = nullptr;
  *out = nm;
  return HBN_OK;
}

// kernels report internal inconsistencies (iteration caps) through mapped host memory; call
// after a synchronisation
int checkFault(hbn_navmesh* nm) {
  if (nm->faultHost && nm->faultHost[0]) {
    char buf[160];
    snprintf(buf, sizeof(buf), "kernel watchdog tripped %u time(s); query %u, site %u", nm->faultHost[0],
             nm->faultHost[1], nm->faultHost[2]);
    memset(nm->faultHost, 0, 64);
    return fail(HBN_ERR_CUDA, buf);
  }
  return HBN_OK;
}

constexpr int64_t kSnapChunk = 1 << 18;   // points per pass of the candidate-list pipeline
constexpr int64_t kSnapSmall = 4096;      // below this one k_snap<8> launch is cheaper than five kernels + a scan

// Stream order between calls on one handle (see hbn_navmesh::lastDone).  During a stream capture the
// caller owns the ordering (an event recorded outside the capture cannot be waited on inside it).
struct CallOrder {
  hbn_navmesh* nm;
  cudaStream_t st;
  bool capturing = false;
  CallOrder(hbn_navmesh* n, cudaStream_t s) : nm(n), st(s) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess) capturing = cs != cudaStreamCaptureStatusNone;
    else cudaGetLastError();
    if (!capturing && nm->lastValid && nm->lastStream != st) cudaStreamWaitEvent(st, nm->lastDone, 0);
  }
  ~CallOrder() {
    if (capturing) return;
    if (cudaEventRecord(nm->lastDone, st) == cudaSuccess) {
      nm->lastStream = st;
      nm->lastValid = true;
    } else {
      cudaGetLastError();
    }
  }
};

// Up to three small independent projectToPoly batches in one k_snap_dual launch (Options::snapDual).
// Returns false when the batches do not qualify (the caller then launches them one after the other).
// 1024 C2 points: find_path's snap phase 71 -> 38 us, with spreading 28 us.
bool snapDualLaunch(hbn_navmesh* nm, SnapJob a, SnapJob b, SnapJob c, cudaStream_t st, int* rc) {
  *rc = HBN_OK;
  if (!nm->opt.snapDual || nm->opt.snapGroup || a.n <= 0 || b.n <= 0 || a.n >= kSnapSmall || b.n >= kSnapSmall ||
      c.n >= kSnapSmall)
    return false;
  const int64_t n = a.n + b.n + std::max<int64_t>(c.n, 0);
  const unsigned threads = nm->opt.snapSpread ? kSnapW : 256;
  const int64_t gpb = threads / kSnapW;
  const int64_t blocks = (n + gpb - 1) / gpb;
  if (c.n < 0) c.n = 0;
  k_snap_dual<kSnapW><<<static_cast<unsigned>(blocks), threads, 0, st>>>(nm->view, a, b, c);
  nm->launches++;
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) *rc = fail(HBN_ERR_CUDA, std::string("k_snap_dual: ") + cudaGetErrorString(ce));
  return true;
}

// scratch of the candidate-list snap pipeline for chunks of up to cmax points
int snapScratch(hbn_navmesh* nm, int64_t cmax, SnapScratch* sc) {
  size_t cap = static_cast<size_t>(cmax) * kSnapAvgCap;
  if (nm->opt.snapCap > 0) cap = static_cast<size_t>(nm->opt.snapCap);  // testing: force the fallback
  int rc;
  if ((rc = nm->snapWin.ensure(cmax * 4)) || (rc = nm->snapG.ensure(cap * 4)) || (rc = nm->snapQ.ensure(cap * 4)) ||
      (rc = nm->snapD.ensure(cap * 4)) || (rc = nm->snapOut.ensure(cap * 4)) || (rc = nm->snapBest.ensure(cmax * 8)) ||
      (rc = nm->snapTodo.ensure(16)))
    return rc;
  sc->todo = static_cast<uint32_t*>(nm->snapTodo.p);
  sc->total = sc->todo + 1;
  sc->cap = static_cast<uint32_t>(cap);
  sc->candG = static_cast<uint32_t*>(nm->snapG.p);
  sc->candTag = static_cast<uint32_t*>(nm->snapQ.p);
  sc->candLb = static_cast<float*>(nm->snapOut.p);
  sc->candD = static_cast<float*>(nm->snapD.p);
  sc->best = static_cast<unsigned long long*>(nm->snapBest.p);
  sc->winner = static_cast<uint32_t*>(nm->snapWin.p);
  return HBN_OK;
}

// projectToPoly for n points.  Large batches: walk -> eval -> mark -> select (hbn_snap.cuh), nothing
// read back by the host; small ones: the lane-group kernel.
int snapLaunch(hbn_navmesh* nm, const float* pts, const int32_t* islands, int64_t n, float* out_pts,
               uint32_t* out_g, uint32_t* out_refs, int32_t* out_isl, uint8_t* out_nav,
               float maxYDelta, cudaStream_t st) {
  if (n <= 0) return HBN_OK;
  const int groupsPerBlock = 256 / kSnapW;
  const int64_t maxBlocks = static_cast<int64_t>(nm->smCount) * 64;
  const bool forceGroup = nm->opt.snapGroup != 0;  // testing: the lane-group kernel only
  if (n < kSnapSmall || forceGroup) {
    int64_t blocks = std::min(maxBlocks, (n + groupsPerBlock - 1) / groupsPerBlock);
    unsigned threads = 256;
    // one lane group per warp (blocks of 8 threads): the groups of a small batch neither share a
    // warp's issue slots nor diverge against each other, and 1024 points cover the SMs instead
    // of 32 blocks (HBN_SNAP_SPREAD=0: blocks of 256 threads)
    if (nm->opt.snapSpread && n < kSnapSmall) {
      threads = kSnapW;
      blocks = n;
    }
    k_snap<kSnapW><<<static_cast<unsigned>(blocks), threads, 0, st>>>(nm->view, pts, islands, n, out_pts, out_g,
                                                                      out_refs, out_isl, out_nav, maxYDelta, nullptr);
    nm->launches++;
    CK(cudaGetLastError());
    return HBN_OK;
  }
  static_assert(kSnapChunk <= (1ll << kSnapQBits), "a slot tag holds the point index of a chunk");
  SnapScratch sc{};
  int rc;
  if ((rc = snapScratch(nm, std::min(n, kSnapChunk), &sc))) return rc;
  for (int64_t c0 = 0; c0 < n; c0 += kSnapChunk) {
    const int64_t cn = std::min(kSnapChunk, n - c0);
    const float* p = pts + 3 * c0;
    const int32_t* isl = islands ? islands + c0 : nullptr;
    const unsigned pb = static_cast<unsigned>((cn + 255) / 256);
    CK(cudaMemsetAsync(sc.todo, 0, 8, st));  // todo flag + slot counter
    k_snap_walk<<<pb, 256, 0, st>>>(nm->view, p, isl, cn, sc);
    const unsigned eb = static_cast<unsigned>(nm->smCount * 16);
    for (int pass = 0; pass < 2; ++pass) k_snap_eval<<<eb, 256, 0, st>>>(nm->view, p, isl, sc, pass);
    k_snap_mark<<<eb, 256, 0, st>>>(sc);
    k_snap_select<<<pb, 256, 0, st>>>(nm->view, p, isl, cn, sc, out_pts ? out_pts + 3 * c0 : nullptr,
                                      out_g ? out_g + c0 : nullptr, out_refs ? out_refs + c0 : nullptr,
                                      out_isl ? out_isl + c0 : nullptr, out_nav ? out_nav + c0 : nullptr, maxYDelta);
    // redone here only if the chunk's candidates did not fit the scratch (decided on the device)
    const int64_t blocks = std::min(maxBlocks, (cn + groupsPerBlock - 1) / groupsPerBlock);
    k_snap<kSnapW><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        nm->view, p, isl, cn, out_pts ? out_pts + 3 * c0 : nullptr, out_g ? out_g + c0 : nullptr,
        out_refs ? out_refs + c0 : nullptr, out_isl ? out_isl + c0 : nullptr, out_nav ? out_nav + c0 : nullptr,
        maxYDelta, sc.todo);
    nm->launches += 6;
    CK(cudaGetLastError());
  }
  return HBN_OK;
}

}  // namespace

extern "C" {

const char* hbn_last_error(void) { return g_err.c_str(); }

int hbn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

float hbn_uniform(uint64_t seed, uint64_t query, uint32_t draw) { return uniform01(seed, query, draw); }

void hbn_std_sort_order(const float* key, int n, int32_t* order) {
  for (int i = 0; i < n; ++i) order[i] = i;
  stdSortOrder(order, key, n);
}

int hbn_navmesh_create_from_mset(const void* bytes, size_t len, int device, hbn_navmesh_t* out) {
  if (!bytes || !out) return fail(HBN_ERR_INVALID, "null argument");
  *out = nullptr;
  HostNavMesh mesh;
  std::string err;
  if (!mesh.loadMSET(static_cast<const uint8_t*>(bytes), len, err)) return fail(HBN_ERR_FORMAT, err);
  return finishCreate(mesh, nullptr, nullptr, 0, device, out);
}

int hbn_navmesh_create_from_tiles(const hbn_tile_blob* tiles, int n_tiles, const float* params5,
                                  int max_tiles, int max_polys, const int32_t* poly_islands,
                                  const float* island_radii, int n_islands, int device, hbn_navmesh_t* out) {
  if (!tiles || n_tiles <= 0 || !params5 || !out) return fail(HBN_ERR_INVALID, "null argument");
  *out = nullptr;
  HostNavMesh mesh;
  std::string err;
  DtNavMeshParams p;
  p.orig[0] = params5[0]; p.orig[1] = params5[1]; p.orig[2] = params5[2];
  p.tileWidth = params5[3];
  p.tileHeight = params5[4];
  p.maxTiles = max_tiles;
  p.maxPolys = max_polys;
  if (!mesh.init(p, err)) return fail(HBN_ERR_FORMAT, err);
  for (int i = 0; i < n_tiles; ++i)
    if (!mesh.addFinalisedTile(static_cast<const uint8_t*>(tiles[i].data), tiles[i].size,
                               tiles[i].tile_ref, err))
      return fail(HBN_ERR_FORMAT, err);
  return finishCreate(mesh, poly_islands, poly_islands ? island_radii : nullptr, n_islands, device, out);
}

void hbn_navmesh_destroy(hbn_navmesh_t nm) {
  if (!nm) return;
  DeviceGuard g(nm->device);
  for (void* d : nm->devArrays) cudaFree(d);
  for (auto& sg : nm->stepGraphs)
    if (sg.exec) cudaGraphExecDestroy(sg.exec);
  for (DevBuf* b : {&nm->sG, &nm->eG, &nm->e2G, &nm->sPt, &nm->ePt, &nm->epPt, &nm->e2Pt, &nm->lastPoly,
                    &nm->lists, &nm->counters, &nm->wsL, &nm->io, &nm->work, &nm->mgDist,
                    &nm->mgBounds, &nm->mgOrder, &nm->mgEnd, &nm->mgMask, &nm->envG, &nm->envPt, &nm->envFlag, &nm->envPos,
                    &nm->folS, &nm->folT, &nm->folE, &nm->folF, &nm->folGeo, &nm->folObs, &nm->folGeo0, &nm->folPos,
                    &nm->fpCls, &nm->fpWork, &nm->fpStat,
                    &nm->fpLen, &nm->fpCorr, &nm->fpBucket, &nm->wsLane, &nm->laneGen, &nm->snapWin,
                    &nm->snapG, &nm->snapQ, &nm->snapD, &nm->snapOut, &nm->snapBest, &nm->snapTodo})
    b->release();
  if (nm->pinned) cudaFreeHost(nm->pinned);
  if (nm->faultHost) cudaFreeHost(nm->faultHost);
  if (nm->lastDone) cudaEventDestroy(nm->lastDone);
  if (nm->stream) cudaStreamDestroy(nm->stream);
  delete nm;
}

int hbn_navmesh_get_info(hbn_navmesh_t nm, hbn_navmesh_info* out) {
  if (!nm || !out) return fail(HBN_ERR_INVALID, "null argument");
  const FlatNav& f = nm->flat;
  memset(out, 0, sizeof(*out));
  out->device = nm->device;
  for (const TileRec& t : f.tiles) out->num_tiles += t.pad[0] ? 1 : 0;
  out->num_polys = static_cast<int32_t>(f.polys.size());
  out->num_links = static_cast<int32_t>(f.links.size());
  out->num_bv_nodes = static_cast<int32_t>(f.bv.size());
  out->num_islands = static_cast<int32_t>(f.islandRadius.size());
  out->poly_bits = f.polyBits;
  out->tile_bits = f.tileBits;
  out->salt_bits = f.saltBits;
  out->has_settings = f.hasSettings ? 1 : 0;
  for (int k = 0; k < 3; ++k) {
    out->bounds_min[k] = f.bounds[k];
    out->bounds_max[k] = f.bounds[3 + k];
  }
  out->navigable_area = f.totalArea;
  out->device_bytes = nm->deviceBytes;
  return HBN_OK;
}

int hbn_navmesh_island_info(hbn_navmesh_t nm, int island, float* radius, float* area) {
  if (!nm) return fail(HBN_ERR_INVALID, "null argument");
  if (island < 0 || island >= static_cast<int>(nm->flat.islandRadius.size()))
    return fail(HBN_ERR_INVALID, "not a valid index for this island system");
  if (radius) *radius = nm->flat.islandRadius[island];
  if (area) *area = nm->flat.islandArea[island];
  return HBN_OK;
}

int hbn_navmesh_set_bounds(hbn_navmesh_t nm, const float* bounds6) {
  if (!nm || !bounds6) return fail(HBN_ERR_INVALID, "null argument");
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  memcpy(nm->flat.bounds, bounds6, 24);
  return HBN_OK;
}

int hbn_navmesh_set_settings(hbn_navmesh_t nm, const void* in56) {
  if (!nm || !in56) return fail(HBN_ERR_INVALID, "null argument");
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  nm->host.setSettings(static_cast<const uint8_t*>(in56));
  memcpy(nm->flat.settings, in56, 56);
  nm->flat.hasSettings = true;
  return HBN_OK;
}

static int envStepReserve(hbn_navmesh* nm, int64_t n);

int hbn_navmesh_set_option(hbn_navmesh_t nm, const char* key, int64_t value) {
  if (!nm || !key) return fail(HBN_ERR_INVALID, "null argument");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  const std::string k(key);
  g_scratchEpoch++;  // cached graphs were captured under the old options
  if (k == "lane_cfg") {
    nm->opt.laneCfg = static_cast<int>(value);
    return laneConfigure(nm);
  } else if (k == "blocks_per_sm") {
    nm->opt.blocksPerSm = static_cast<int>(value);
    return laneConfigure(nm);
  } else if (k == "lane_spread") nm->opt.laneSpread = value != 0;
  else if (k == "snap_spread") nm->opt.snapSpread = value != 0;
  else if (k == "snap_dual") nm->opt.snapDual = value != 0;
  else if (k == "snap_group") nm->opt.snapGroup = value != 0;
  else if (k == "snap_cap") nm->opt.snapCap = value;
  else if (k == "nvtx") nm->opt.nvtx = value != 0;
  else if (k == "lane_scratch_bytes") {
    // a smaller cap takes effect at once: the search state is released and re-made on the next call
    nm->opt.laneScratchBytes = value;
    CK(cudaDeviceSynchronize());
    nm->wsLane.release();
    nm->laneGen.release();
    nm->laneSlots = 0;
  } else {
    return fail(HBN_ERR_INVALID, "unknown option: " + k);
  }
  return HBN_OK;
}

int64_t hbn_navmesh_scratch_bytes(hbn_navmesh_t nm) {
  if (!nm) return -1;
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  int64_t total = 0;
  for (const DevBuf* b : {&nm->sG, &nm->eG, &nm->e2G, &nm->sPt, &nm->ePt, &nm->epPt, &nm->e2Pt, &nm->lastPoly,
                          &nm->lists, &nm->counters, &nm->wsL, &nm->io, &nm->work, &nm->mgDist, &nm->mgBounds,
                          &nm->mgOrder, &nm->mgEnd, &nm->mgMask, &nm->envG, &nm->envPt, &nm->envFlag, &nm->envPos,
                          &nm->folS, &nm->folT, &nm->folE, &nm->folF, &nm->folGeo, &nm->folObs, &nm->folGeo0, &nm->folPos,
                          &nm->fpCls, &nm->fpWork, &nm->fpStat, &nm->fpLen, &nm->fpCorr, &nm->fpBucket, &nm->wsLane,
                          &nm->laneGen, &nm->snapWin, &nm->snapG, &nm->snapQ, &nm->snapD, &nm->snapOut,
                          &nm->snapBest, &nm->snapTodo})
    total += static_cast<int64_t>(b->cap);
  return total;
}

int hbn_navmesh_reserve(hbn_navmesh_t nm, int64_t n) {
  if (!nm || n <= 0) return fail(HBN_ERR_INVALID, "bad argument");
  if (n >= (1ll << 31)) return fail(HBN_ERR_INVALID, "batch too large");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  int rc;
  if ((rc = envStepReserve(nm, n)) || (rc = nm->lists.ensure(n * 8)) || (rc = nm->io.ensure(n * 64 + 8192)) ||
      (rc = nm->wsL.ensure(static_cast<size_t>(nm->blocksWallL) * kFpWarps * wsGlobalBytes(kCapL, kWsHybrid))))
    return rc;
  if (n >= kSnapSmall) {  // the candidate-list snap pipeline's scratch
    SnapScratch sc{};
    if ((rc = snapScratch(nm, std::min(n, kSnapChunk), &sc))) return rc;
  }
  CK(cudaStreamSynchronize(nm->stream));  // the zeroing of new search state
  return HBN_OK;
}

int64_t hbn_navmesh_save_mset(hbn_navmesh_t nm, void* out, int64_t cap) {
  if (!nm) return -1;
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  std::vector<uint8_t> image;
  std::string err;
  if (!nm->host.saveMSET(image, err)) {
    fail(HBN_ERR_FORMAT, err);
    return -1;
  }
  if (out && cap >= static_cast<int64_t>(image.size())) memcpy(out, image.data(), image.size());
  return static_cast<int64_t>(image.size());
}

int hbn_navmesh_get_settings(hbn_navmesh_t nm, void* out56) {
  if (!nm || !out56) return fail(HBN_ERR_INVALID, "null argument");
  if (!nm->flat.hasSettings) return fail(HBN_ERR_INVALID, "navmesh image carries no NavMeshSettings");
  memcpy(out56, nm->flat.settings, 56);
  return HBN_OK;
}

int64_t hbn_navmesh_launch_count(hbn_navmesh_t nm) { return nm ? nm->launches : 0; }

int hbn_navmesh_set_profiling(hbn_navmesh_t nm, int enable) {
  if (!nm) return fail(HBN_ERR_INVALID, "null argument");
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  nm->profile = enable != 0;
  return HBN_OK;
}

int hbn_navmesh_phase_times(hbn_navmesh_t nm, double* out_ms2, int64_t* out_calls) {
  if (!nm || !out_ms2) return fail(HBN_ERR_INVALID, "null argument");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  out_ms2[0] = out_ms2[1] = 0.0;
  if (out_calls) *out_calls = static_cast<int64_t>(nm->phaseEvents.size());
  for (auto& pe : nm->phaseEvents) {
    CK(cudaEventSynchronize(pe.e[2]));
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, pe.e[0], pe.e[1]));
    CK(cudaEventElapsedTime(&b, pe.e[1], pe.e[2]));
    out_ms2[0] += a;
    out_ms2[1] += b;
    for (auto& e : pe.e) cudaEventDestroy(e);
  }
  nm->phaseEvents.clear();
  return checkFault(nm);
}

int hbn_navmesh_work_counters(hbn_navmesh_t nm, uint64_t* out8, int reset) {
  if (!nm || !out8) return fail(HBN_ERR_INVALID, "null argument");
  DeviceGuard g(nm->device);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out8, nm->work.p, 64, cudaMemcpyDeviceToHost));
  if (reset) CK(cudaMemset(nm->work.p, 0, 64));
  return checkFault(nm);
}

int64_t hbn_navmesh_triangles(hbn_navmesh_t nm, int island, float* out, int64_t cap_tris) {
  if (!nm) return -1;
  const FlatNav& f = nm->flat;
  int64_t n = 0;
  // getNavMeshData (PF.cpp:1898-1944): the detail triangles of EVERY poly of the island (all polys for
  // island -1) in tile-table, poly order -- disabled zero-area polys and non-walkable ones included
  for (const PolyRec& p : f.polys) {
    if (island >= 0 && p.island != island) continue;
    for (int j = 0; j < p.detTriCount; ++j) {
      const uint8_t* t = &f.detTris[static_cast<size_t>(p.detTriBase + j) * 4];
      if (out && n < cap_tris)
        for (int k = 0; k < 3; ++k) {
          const float* v = t[k] < p.nv ? &p.v[t[k] * 3]
                                       : &f.detVerts[static_cast<size_t>(p.detVertBase + (t[k] - p.nv)) * 3];
          memcpy(out + n * 9 + k * 3, v, 12);
        }
      n++;
    }
  }
  return n;
}

// ------------------------------------------------------------------------------------
int hbn_snap_point_dev(hbn_navmesh_t nm, const float* pts, const int32_t* islands, int64_t n,
                       float* out_pts, uint32_t* out_refs, int32_t* out_islands, void* stream) {
  if (!nm || (n > 0 && !pts)) return fail(HBN_ERR_INVALID, "null argument");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  NvtxRange nv(nm->opt.nvtx, "hbn_snap_point");
  CallOrder order(nm, static_cast<cudaStream_t>(stream));
  return snapLaunch(nm, pts, islands, n, out_pts, nullptr, out_refs, out_islands, nullptr, 0.f,
                    static_cast<cudaStream_t>(stream));
}

int hbn_is_navigable_dev(hbn_navmesh_t nm, const float* pts, int64_t n, float max_y_delta,
                         uint8_t* out, void* stream) {
  if (!nm || (n > 0 && (!pts || !out))) return fail(HBN_ERR_INVALID, "null argument");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  NvtxRange nv(nm->opt.nvtx, "hbn_is_navigable");
  CallOrder order(nm, static_cast<cudaStream_t>(stream));
  return snapLaunch(nm, pts, nullptr, n, nullptr, nullptr, nullptr, nullptr, out, max_y_delta,
                    static_cast<cudaStream_t>(stream));
}

}  // extern "C"

constexpr int64_t kFpChunk = 1 << 20;  // queries per pass of the pipeline (1 KB corridor ring each)

// scratch of a find_path over n pairs (nStarts distinct starts)
static int findPathReserve(hbn_navmesh* nm, int64_t n, int64_t nStarts, int startDiv) {
  const int64_t chunk = startDiv > 1 ? std::max<int64_t>(1, kFpChunk / startDiv) * startDiv : kFpChunk;
  const int64_t nChunks = (n + chunk - 1) / chunk;
  const int64_t cmax = std::min(n, chunk);
  int rc;
  if ((rc = nm->sG.ensure(nStarts * 4)) || (rc = nm->eG.ensure(n * 4)) || (rc = nm->sPt.ensure(nStarts * 12)) ||
      (rc = nm->ePt.ensure(n * 12)) || (rc = nm->counters.ensure(static_cast<size_t>(nChunks) * kFpCounterBytes)) ||
      (rc = nm->fpCls.ensure(cmax)) || (rc = nm->fpBucket.ensure(cmax)) || (rc = nm->fpWork.ensure(cmax * 4)) ||
      (rc = nm->fpStat.ensure(cmax * 4)) || (rc = nm->fpLen.ensure(cmax * 4)) ||
      (rc = nm->fpCorr.ensure(static_cast<size_t>(cmax) * kMaxPathPolys * 4)))
    return rc;
  return HBN_OK;
}

// grid of the search kernel for cn queries: lanes per warp (spreading) and one-warp blocks
static void laneGrid(const hbn_navmesh* nm, int64_t cn, int* lanes, int* blocks) {
  int64_t l = 32;
  if (nm->opt.laneSpread) l = std::max<int64_t>(1, std::min<int64_t>(32, (cn + nm->blocksFpLane - 1) / nm->blocksFpLane));
  *lanes = static_cast<int>(l);
  *blocks = static_cast<int>(std::min<int64_t>(nm->blocksFpLane, (cn + l - 1) / l));
}

// n (start, end) pairs; startDiv > 1: pair q uses start q / startDiv (multi-goal layout)
// pairMask: queries with a zero byte are skipped -- their outputs are left alone, or get an infinite
// distance with kFpFillSkipped.  kFpReuseSnaps: the projectToPoly results of the previous call on the
// same points are still in the scratch.  snapS / snapE (kFpGivenSnaps): projections supplied by the
// caller (the env step hands over the ones tryStep made).
enum { kFpReuseSnaps = 1 << 16, kFpFillSkipped = 1 << 17, kFpGivenSnaps = 1 << 18 };
struct GivenSnaps {
  const uint32_t* sG; const float* sPt; const uint32_t* eG; const float* ePt;
};
static int findPathLaunch(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                          int startDiv, float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                          uint32_t* out_corridor, int32_t* out_ncorridor, uint32_t* out_status,
                          int flags, cudaStream_t st, const uint8_t* pairMask = nullptr,
                          const GivenSnaps* given = nullptr) {
  if (!nm || (n > 0 && (!starts || !ends || !out_dist))) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  const int64_t nStarts = startDiv > 1 ? n / startDiv : n;
  if (n >= (1ll << 31)) return fail(HBN_ERR_INVALID, "batch too large");
  if (out_pts && max_pts <= 0) return fail(HBN_ERR_INVALID, "max_pts must be positive");
  int rc;
  if ((rc = checkFault(nm))) return rc;  // reported by an earlier launch
  const int64_t chunk = startDiv > 1 ? std::max<int64_t>(1, kFpChunk / startDiv) * startDiv : kFpChunk;
  const int64_t nChunks = (n + chunk - 1) / chunk;
  if ((rc = findPathReserve(nm, n, nStarts, startDiv))) return rc;
  CK(cudaMemsetAsync(nm->counters.p, 0, static_cast<size_t>(nChunks) * kFpCounterBytes, st));
  hbn_navmesh::PhaseEv pe{};
  bool profile = nm->profile;
  if (profile) {  // not inside a stream capture
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) profile = false;
  }
  if (profile) {
    for (auto& e : pe.e) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(pe.e[0], st));
  }
  const uint32_t* sGp = static_cast<uint32_t*>(nm->sG.p);
  const float* sPtp = static_cast<float*>(nm->sPt.p);
  const uint32_t* eGp = static_cast<uint32_t*>(nm->eG.p);
  const float* ePtp = static_cast<float*>(nm->ePt.p);
  if (flags & kFpGivenSnaps) {
    sGp = given->sG; sPtp = given->sPt; eGp = given->eG; ePtp = given->ePt;
  } else if ((flags & kFpReuseSnaps) == 0) {
    if (snapDualLaunch(nm, SnapJob{starts, nStarts, static_cast<float*>(nm->sPt.p), static_cast<uint32_t*>(nm->sG.p)},
                       SnapJob{ends, n, static_cast<float*>(nm->ePt.p), static_cast<uint32_t*>(nm->eG.p)},
                       SnapJob{nullptr, 0, nullptr, nullptr}, st, &rc)) {
      if (rc) return rc;
    } else {
      if ((rc = snapLaunch(nm, starts, nullptr, nStarts, static_cast<float*>(nm->sPt.p),
                           static_cast<uint32_t*>(nm->sG.p), nullptr, nullptr, nullptr, 0.f, st)))
        return rc;
      if ((rc = snapLaunch(nm, ends, nullptr, n, static_cast<float*>(nm->ePt.p),
                           static_cast<uint32_t*>(nm->eG.p), nullptr, nullptr, nullptr, 0.f, st)))
        return rc;
    }
  }
  if (profile) CK(cudaEventRecord(pe.e[1], st));
  unsigned long long* workCtr = (flags & HBN_FP_COUNT_WORK) ? static_cast<unsigned long long*>(nm->work.p) : nullptr;
  for (int64_t ci = 0; ci < nChunks; ++ci) {
    const int64_t c0 = ci * chunk;
    const int64_t cn = std::min(chunk, n - c0);
    const int64_t s0 = startDiv > 1 ? c0 / startDiv : c0;
    uint32_t* cnt = static_cast<uint32_t*>(nm->counters.p) + ci * (kFpCounterBytes / 4);
    // classify -> search list -> search -> funnel
    uint8_t* cls = static_cast<uint8_t*>(nm->fpCls.p);
    uint32_t* work = static_cast<uint32_t*>(nm->fpWork.p);
    uint8_t* bucket = static_cast<uint8_t*>(nm->fpBucket.p);
    k_fp_classify<<<static_cast<unsigned>((cn + 255) / 256), 256, 0, st>>>(
        nm->view, sGp + s0, sPtp + 3 * s0, eGp + c0, ePtp + 3 * c0, cn, startDiv, pairMask ? pairMask + c0 : nullptr,
        cls, bucket, cnt + 16, cnt + 4);
    k_fp_scatter<<<static_cast<unsigned>((cn + 255) / 256), 256, 0, st>>>(bucket, cn, cnt + 16, cnt + 16 + kFpBuckets,
                                                                         work);
    nm->launches += 2;
    CK(cudaGetLastError());
    SearchArgs ga{};
    ga.sG = sGp + s0; ga.sPt = sPtp + 3 * s0; ga.eG = eGp + c0; ga.ePt = ePtp + 3 * c0;
    ga.work = work; ga.workCount = cnt + 4;
    ga.counter = cnt + 5;
    ga.astat = static_cast<uint32_t*>(nm->fpStat.p);
    ga.fullLen = static_cast<int32_t*>(nm->fpLen.p);
    ga.corrVia = static_cast<uint32_t*>(nm->fpCorr.p);
    ga.startDiv = startDiv;
    ga.fastFail = (flags & HBN_FP_EXACT_STATUS) ? 0 : 1;
    ga.allCorridors = (out_corridor || out_ncorridor) ? 1 : 0;
    ga.workCtr = workCtr;
    ga.fault = nm->faultDev;
    {
      int lanes = 32, blocks = 1;
      laneGrid(nm, cn, &lanes, &blocks);
      LaneScratch sc{};
      // (under a memory cap the grid comes back smaller: every lane then takes more queries)
      if ((rc = laneScratch(nm, st, &blocks, lanes, &sc))) return rc;
      ga.laneLimit = lanes;
      ga.numWarps = blocks;
      size_t smLane = 0;
      int wpb = 1;
      const void* fn = laneKernel(nm->opt.laneCfg, &smLane, &wpb);
      void* kargs[] = {&nm->view, &ga, &sc};
      CK(cudaLaunchKernel(fn, dim3(static_cast<unsigned>((blocks + wpb - 1) / wpb)), dim3(32 * wpb), kargs, smLane, st));
      nm->launches++;
      CK(cudaGetLastError());
    }
    FpFunnelArgs fa{};
    fa.starts = starts + 3 * s0; fa.ends = ends + 3 * c0;
    fa.sG = ga.sG; fa.sPt = ga.sPt; fa.eG = ga.eG; fa.ePt = ga.ePt;
    fa.cls = cls; fa.astat = ga.astat; fa.fullLen = ga.fullLen; fa.corrVia = ga.corrVia;
    fa.n = cn; fa.startDiv = startDiv;
    fa.out_dist = out_dist + c0;
    fa.out_npts = out_npts ? out_npts + c0 : nullptr;
    fa.out_pts = out_pts ? out_pts + static_cast<size_t>(c0) * max_pts * 3 : nullptr;
    fa.max_pts = max_pts;
    fa.out_corridor = out_corridor ? out_corridor + static_cast<size_t>(c0) * kMaxPathPolys : nullptr;
    fa.out_ncorridor = out_ncorridor ? out_ncorridor + c0 : nullptr;
    fa.out_status = out_status ? out_status + 2 * c0 : nullptr;
    fa.workCtr = workCtr;
    fa.fillSkipped = (flags & kFpFillSkipped) ? 1 : 0;
    k_fp_funnel<<<static_cast<unsigned>((cn + 127) / 128), 128, 0, st>>>(nm->view, fa);
    nm->launches++;
    CK(cudaGetLastError());
  }
  if (profile) {
    CK(cudaEventRecord(pe.e[2], st));
    nm->phaseEvents.push_back(pe);
  }
  return HBN_OK;
}


extern "C" int hbn_find_path_dev(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                                 float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                                 uint32_t* out_corridor, int32_t* out_ncorridor, uint32_t* out_status,
                                 int flags, void* stream) {
  if (!nm) return fail(HBN_ERR_INVALID, "null navmesh");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  NvtxRange nv(nm->opt.nvtx, "hbn_find_path");
  CallOrder order(nm, static_cast<cudaStream_t>(stream));
  return findPathLaunch(nm, starts, ends, n, 1, out_dist, out_npts, out_pts, max_pts, out_corridor,
                        out_ncorridor, out_status, flags & 0xffff, static_cast<cudaStream_t>(stream));
}

extern "C" {

int hbn_find_path_multigoal_dev(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                                int g, float* out_dist, int32_t* out_index, int32_t* out_npts,
                                float* out_pts, int max_pts, void* stream) {
  if (!nm || (n > 0 && (!starts || !ends || !out_dist || !out_index)))
    return fail(HBN_ERR_INVALID, "null argument");
  if (g <= 0) return fail(HBN_ERR_INVALID, "g must be positive");
  if (n <= 0) return HBN_OK;
  const int64_t pairs = n * g;
  if (pairs >= (1ll << 31)) return fail(HBN_ERR_INVALID, "batch too large");
  DeviceGuard gd(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NvtxRange nv(nm->opt.nvtx, "hbn_find_path_multigoal");
  CallOrder order(nm, st);
  int rc;
  const bool wantPts = out_npts || out_pts;
  if ((rc = nm->mgDist.ensure(pairs * 4)) || (rc = nm->mgBounds.ensure(pairs * 4)) ||
      (rc = nm->mgOrder.ensure(pairs * 4)) || (wantPts && (rc = nm->mgEnd.ensure(n * 12))))
    return rc;
  if (g > kMultiGoalFirst) {
    // two rounds of pair searches: the kMultiGoalFirst goals of smallest bound of every start, then
    // the later goals the reference would not skip (hbn_query.h); bounds / order in the select's scratch
    if ((rc = nm->mgMask.ensure(pairs))) return rc;
    uint8_t* mask = static_cast<uint8_t*>(nm->mgMask.p);
    const unsigned sb = static_cast<unsigned>((n + 127) / 128);
    k_multigoal_round1<<<sb, 128, 0, st>>>(starts, ends, n, g, static_cast<float*>(nm->mgBounds.p),
                                           static_cast<int32_t*>(nm->mgOrder.p), mask);
    nm->launches++;
    CK(cudaGetLastError());
    if ((rc = findPathLaunch(nm, starts, ends, pairs, g, static_cast<float*>(nm->mgDist.p), nullptr, nullptr, 0,
                             nullptr, nullptr, nullptr, kFpFillSkipped, st, mask)))
      return rc;
    k_multigoal_round2<<<sb, 128, 0, st>>>(static_cast<uint32_t*>(nm->sG.p), static_cast<uint32_t*>(nm->eG.p),
                                           static_cast<float*>(nm->mgDist.p), n, g,
                                           static_cast<float*>(nm->mgBounds.p),
                                           static_cast<int32_t*>(nm->mgOrder.p), mask);
    nm->launches++;
    CK(cudaGetLastError());
    if ((rc = findPathLaunch(nm, starts, ends, pairs, g, static_cast<float*>(nm->mgDist.p), nullptr, nullptr, 0,
                             nullptr, nullptr, nullptr, kFpReuseSnaps, st, mask)))
      return rc;
  } else if ((rc = findPathLaunch(nm, starts, ends, pairs, g, static_cast<float*>(nm->mgDist.p), nullptr, nullptr,
                                  0, nullptr, nullptr, nullptr, 0, st))) {
    // every (start, goal) pair's findPathInternal, in parallel ...
    return rc;
  }
  // ... then the reference's sequential goal loop per start (sG / eG are still in the scratch)
  k_multigoal_select<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(
      starts, ends, static_cast<uint32_t*>(nm->sG.p), static_cast<uint32_t*>(nm->eG.p),
      static_cast<float*>(nm->mgDist.p), n, g, static_cast<float*>(nm->mgBounds.p),
      static_cast<int32_t*>(nm->mgOrder.p), out_dist, out_index,
      wantPts ? static_cast<float*>(nm->mgEnd.p) : nullptr);
  nm->launches++;
  CK(cudaGetLastError());
  if (wantPts) {  // the chosen goal's path points: one more find_path per start
    if ((rc = nm->mgDist.ensure(std::max<int64_t>(pairs, n) * 4))) return rc;
    if ((rc = findPathLaunch(nm, starts, static_cast<float*>(nm->mgEnd.p), n, 1,
                             static_cast<float*>(nm->mgDist.p), out_npts, out_pts, max_pts, nullptr, nullptr,
                             nullptr, 0, st)))
      return rc;
  }
  return HBN_OK;
}

// scratch of try_step / env step over n queries
static int tryStepReserve(hbn_navmesh* nm, int64_t n) {
  int rc;
  if ((rc = nm->sG.ensure(n * 4)) || (rc = nm->eG.ensure(n * 4)) || (rc = nm->e2G.ensure(n * 4)) ||
      (rc = nm->sPt.ensure(n * 12)) || (rc = nm->epPt.ensure(n * 12)) || (rc = nm->lastPoly.ensure(n * 4)))
    return rc;
  return HBN_OK;
}

// tryStep up to, not including, phase B: projections of start and end (+ an optional third batch in
// the same launch), moveAlongSurface / the no-sliding ray, projection of the end point
static int tryStepFront(hbn_navmesh* nm, const float* starts, const float* ends, int64_t n, int allow_sliding,
                        SnapJob third, float* e2Pt, cudaStream_t st) {
  int rc;
  uint32_t* sG = static_cast<uint32_t*>(nm->sG.p);
  uint32_t* eG = static_cast<uint32_t*>(nm->eG.p);
  float* sPt = static_cast<float*>(nm->sPt.p);
  float* ep = static_cast<float*>(nm->epPt.p);
  uint32_t* last = static_cast<uint32_t*>(nm->lastPoly.p);
  if (snapDualLaunch(nm, SnapJob{starts, n, sPt, sG}, SnapJob{ends, n, nullptr, eG}, third, st, &rc)) {
    if (rc) return rc;
  } else {
    if ((rc = snapLaunch(nm, starts, nullptr, n, sPt, sG, nullptr, nullptr, nullptr, 0.f, st))) return rc;
    if ((rc = snapLaunch(nm, ends, nullptr, n, nullptr, eG, nullptr, nullptr, nullptr, 0.f, st))) return rc;
    if (third.n > 0 &&
        (rc = snapLaunch(nm, third.pts, nullptr, third.n, third.out_pts, third.out_g, nullptr, nullptr, nullptr, 0.f, st)))
      return rc;
  }
  k_trystep_a<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(nm->view, ends, sG, sPt, eG, n,
                                                                      allow_sliding, ep, last);
  nm->launches++;
  CK(cudaGetLastError());
  return snapLaunch(nm, ep, nullptr, n, e2Pt, static_cast<uint32_t*>(nm->e2G.p), nullptr, nullptr, nullptr, 0.f, st);
}

int hbn_try_step_dev(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                     int allow_sliding, float* out_pts, void* stream) {
  if (!nm || (n > 0 && (!starts || !ends || !out_pts))) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NvtxRange nv(nm->opt.nvtx, "hbn_try_step");
  CallOrder order(nm, st);
  int rc;
  if ((rc = tryStepReserve(nm, n))) return rc;
  if ((rc = tryStepFront(nm, starts, ends, n, allow_sliding, SnapJob{nullptr, 0, nullptr, nullptr}, nullptr, st)))
    return rc;
  k_trystep_b<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      nm->view, starts, static_cast<uint32_t*>(nm->sG.p), static_cast<uint32_t*>(nm->e2G.p),
      static_cast<uint32_t*>(nm->lastPoly.p), static_cast<float*>(nm->epPt.p), n, out_pts);
  nm->launches++;
  CK(cudaGetLastError());
  return HBN_OK;
}

static int envStepReserve(hbn_navmesh* nm, int64_t n) {
  int rc;
  if ((rc = tryStepReserve(nm, n)) || (rc = nm->e2Pt.ensure(n * 12)) || (rc = nm->envG.ensure(n * 4)) ||
      (rc = nm->envPt.ensure(n * 12)) || (rc = nm->envFlag.ensure(n)) || (rc = nm->ePt.ensure(n * 12)) ||
      (rc = nm->envPos.ensure(n * 4)) || (rc = findPathReserve(nm, n, n, 1)))
    return rc;
  int lanes = 32, blocks = 1;
  laneGrid(nm, n, &lanes, &blocks);
  LaneScratch sc{};
  return laneScratch(nm, nm->stream, &blocks, lanes, &sc);
}

// One environment step of the PointNav loop (simulator.py:660-673 + the geodesic reward): tryStep from
// `starts` towards `targets`, then find_path from the new position to `goals`.  Equivalent to
// hbn_try_step_dev followed by hbn_find_path_dev (bit for bit), with the projections shared: goals are
// projected in the same launch as starts and targets, and the new position is not projected again
// (k_envstep_b hands find_path the projection phase B already made; nudged positions are redone by
// k_snap_flagged).  9 launches instead of 13.
int hbn_env_step_dev(hbn_navmesh_t nm, const float* starts, const float* targets, const float* goals, int64_t n,
                     int allow_sliding, float* out_pos, float* out_dist, void* stream) {
  if (!nm || (n > 0 && (!starts || !targets || !goals || !out_pos || !out_dist)))
    return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  if (n >= (1ll << 31)) return fail(HBN_ERR_INVALID, "batch too large");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NvtxRange nv(nm->opt.nvtx, "hbn_env_step");
  CallOrder order(nm, st);
  int rc;
  if ((rc = tryStepReserve(nm, n)) || (rc = nm->e2Pt.ensure(n * 12)) || (rc = nm->envG.ensure(n * 4)) ||
      (rc = nm->envPt.ensure(n * 12)) || (rc = nm->envFlag.ensure(n)) || (rc = nm->ePt.ensure(n * 12)) ||
      (rc = nm->envPos.ensure(n * 4)))
    return rc;
  // the goals' projections: points where find_path keeps its end projections (ePt is free during a
  // try_step), polys in a slot of their own (eG holds try_step's end projection)
  uint32_t* gG = static_cast<uint32_t*>(nm->envPos.p);
  float* gPt = static_cast<float*>(nm->ePt.p);
  float* e2Pt = static_cast<float*>(nm->e2Pt.p);
  if ((rc = tryStepFront(nm, starts, targets, n, allow_sliding, SnapJob{goals, n, gPt, gG}, e2Pt, st))) return rc;
  uint32_t* fpG = static_cast<uint32_t*>(nm->envG.p);
  float* fpPt = static_cast<float*>(nm->envPt.p);
  uint8_t* flag = static_cast<uint8_t*>(nm->envFlag.p);
  k_envstep_b<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      nm->view, starts, static_cast<uint32_t*>(nm->sG.p), static_cast<float*>(nm->sPt.p),
      static_cast<uint32_t*>(nm->e2G.p), e2Pt, static_cast<uint32_t*>(nm->lastPoly.p),
      static_cast<float*>(nm->epPt.p), n, out_pos, fpG, fpPt, flag);
  {
    const unsigned threads = (nm->opt.snapSpread && n < kSnapSmall) ? kSnapW : 256;
    const int64_t gpb = threads / kSnapW;
    const int64_t blocks = std::min<int64_t>((n + gpb - 1) / gpb, static_cast<int64_t>(nm->smCount) * 64 * (256 / threads));
    k_snap_flagged<kSnapW><<<static_cast<unsigned>(blocks), threads, 0, st>>>(nm->view, out_pos, flag, n, fpPt, fpG);
  }
  nm->launches += 2;
  CK(cudaGetLastError());
  const GivenSnaps given{fpG, fpPt, gG, gPt};
  return findPathLaunch(nm, out_pos, goals, n, 1, out_dist, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                        kFpGivenSnaps, st, nullptr, &given);
}

int hbn_closest_obstacle_dev(hbn_navmesh_t nm, const float* pts, int64_t n, float max_radius,
                             float* out_hit_pos, float* out_hit_normal, float* out_hit_dist,
                             void* stream);

// GreedyGeodesicFollowerImpl::nextBestPrimAlong for n agents (hbn_follower.cuh): geodesic distance now,
// the forward target of every primitive, try_step / find_path / closest_obstacle over all of them,
// reward and selection -- six batched device calls, nothing on the host in between.
int hbn_follower_best_prims_dev(hbn_navmesh_t nm, const double* rots, const double* poss, const float* goals,
                                int64_t n, const hbn_follower_params* fp, int32_t* out_prim, float* out_geo,
                                void* stream) {
  if (!nm || !fp || (n > 0 && (!rots || !poss || !goals || !out_prim))) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  if (fp->n_steps <= 0 || fp->n_steps > 4096) return fail(HBN_ERR_INVALID, "bad n_steps");
  const int64_t m = n * 2 * fp->n_steps;
  if (m >= (1ll << 31)) return fail(HBN_ERR_INVALID, "batch too large");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NvtxRange nv(nm->opt.nvtx, "hbn_follower_best_prims");
  CallOrder order(nm, st);
  int rc;
  if ((rc = nm->folS.ensure(m * 12)) || (rc = nm->folT.ensure(m * 12)) || (rc = nm->folE.ensure(m * 12)) ||
      (rc = nm->folF.ensure(m * 12)) || (rc = nm->folGeo.ensure(m * 4)) || (rc = nm->folObs.ensure(m * 4)) ||
      (rc = nm->folGeo0.ensure(n * 4)) || (rc = nm->folPos.ensure(n * 12)))
    return rc;
  FollowerParams p{};
  p.goalDist = fp->goal_dist;
  p.forwardAmount = fp->forward_amount;
  p.sinHalf = fp->sin_half_turn;
  p.cosHalf = fp->cos_half_turn;
  p.nSteps = fp->n_steps;
  float* S = static_cast<float*>(nm->folS.p);
  float* T = static_cast<float*>(nm->folT.p);
  float* E = static_cast<float*>(nm->folE.p);
  float* F = static_cast<float*>(nm->folF.p);
  float* geoAfter = static_cast<float*>(nm->folGeo.p);
  float* obs = static_cast<float*>(nm->folObs.p);
  float* geo0 = out_geo ? out_geo : static_cast<float*>(nm->folGeo0.p);
  float* pos32 = static_cast<float*>(nm->folPos.p);
  const unsigned ab = static_cast<unsigned>((n + 127) / 128);
  k_follower_targets<<<ab, 128, 0, st>>>(rots, poss, goals, n, p, pos32, S, T, E);
  nm->launches++;
  CK(cudaGetLastError());
  if ((rc = findPathLaunch(nm, pos32, goals, n, 1, geo0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, st))) return rc;
  if ((rc = hbn_try_step_dev(nm, S, T, m, fp->allow_sliding, F, stream))) return rc;
  if ((rc = findPathLaunch(nm, F, E, m, 1, geoAfter, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, st))) return rc;
  if ((rc = hbn_closest_obstacle_dev(nm, F, m, 1.1f * 0.2f, nullptr, nullptr, obs, stream))) return rc;
  k_follower_select<<<ab, 128, 0, st>>>(geo0, S, T, F, geoAfter, obs, n, p, out_prim);
  nm->launches++;
  CK(cudaGetLastError());
  return HBN_OK;
}

int hbn_closest_obstacle_dev(hbn_navmesh_t nm, const float* pts, int64_t n, float max_radius,
                             float* out_hit_pos, float* out_hit_normal, float* out_hit_dist,
                             void* stream) {
  if (!nm || (n > 0 && (!pts || !out_hit_dist))) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  if (n >= (1ll << 31)) return fail(HBN_ERR_INVALID, "batch too large");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NvtxRange nv(nm->opt.nvtx, "hbn_closest_obstacle");
  CallOrder order(nm, st);
  int rc;
  if ((rc = nm->sG.ensure(n * 4)) || (rc = nm->sPt.ensure(n * 12)) || (rc = nm->lists.ensure(n * 8)) ||
      (rc = nm->wsL.ensure(static_cast<size_t>(nm->blocksWallL) * kFpWarps * wsGlobalBytes(kCapL, kWsHybrid))))
    return rc;
  uint32_t* cnt = static_cast<uint32_t*>(nm->counters.p);
  CK(cudaMemsetAsync(cnt, 0, 64, st));
  if ((rc = snapLaunch(nm, pts, nullptr, n, static_cast<float*>(nm->sPt.p), static_cast<uint32_t*>(nm->sG.p),
                       nullptr, nullptr, nullptr, 0.f, st)))
    return rc;
  WallArgs a{};
  a.sG = static_cast<uint32_t*>(nm->sG.p);
  a.sPt = static_cast<float*>(nm->sPt.p);
  a.n = n;
  a.counter = cnt + 0;
  a.overflow = static_cast<uint32_t*>(nm->lists.p);
  a.overflowCount = cnt + 1;
  a.maxRadius = max_radius;
  a.out_pos = out_hit_pos; a.out_normal = out_hit_normal; a.out_dist = out_hit_dist;
  const int threads = kFpWarps * 32;
  // three tiers, each working off the overflow list of the one before: a query per thread with a
  // 8-node pool, a query per warp with 128 nodes in shared memory, and the reference's 2048 nodes
  k_wall_lane<<<static_cast<unsigned>((n + kWallLaneThreads - 1) / kWallLaneThreads), kWallLaneThreads, 0, st>>>(nm->view, a);
  nm->launches++;
  CK(cudaGetLastError());
  WallArgs a1 = a;
  a1.work = a.overflow;
  a1.workCount = a.overflowCount;
  a1.counter = cnt + 4;
  a1.overflow = static_cast<uint32_t*>(nm->lists.p) + n;
  a1.overflowCount = cnt + 5;
  k_wall<kWallCapS, kWsShared><<<nm->blocksWallS, threads, kFpWarps * wsSharedBytes(kWallCapS, kWsShared), st>>>(nm->view, a1);
  nm->launches++;
  CK(cudaGetLastError());
  WallArgs b = a1;
  b.work = a1.overflow;
  b.workCount = a1.overflowCount;
  b.counter = cnt + 2;
  b.overflowCount = cnt + 3;
  b.scratch = static_cast<char*>(nm->wsL.p);
  k_wall<kCapL, kWsHybrid><<<nm->blocksWallL, threads, kFpWarps * wsSharedBytes(kCapL, kWsHybrid), st>>>(nm->view, b);
  nm->launches++;
  CK(cudaGetLastError());
  return HBN_OK;
}

int hbn_random_points_dev(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n,
                          const int32_t* islands, int max_tries, float* out_pts,
                          uint32_t* out_refs, void* stream) {
  if (!nm || (n > 0 && !out_pts)) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  if (nm->flat.totalArea <= 0.0f)
    return fail(HBN_ERR_NO_AREA, "NavMesh has no navigable area, this indicates an issue with the NavMesh");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  NvtxRange nv(nm->opt.nvtx, "hbn_random_points");
  CallOrder order(nm, static_cast<cudaStream_t>(stream));
  const int groupsPerBlock = 256 / kRandW;
  int64_t blocks = std::min<int64_t>((n + groupsPerBlock - 1) / groupsPerBlock,
                                     static_cast<int64_t>(nm->smCount) * 64);
  k_random<kRandW><<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      nm->view, seed, query0, n, islands, max_tries, out_pts, out_refs);
  nm->launches++;
  CK(cudaGetLastError());
  return HBN_OK;
}

int hbn_random_points_near_dev(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n,
                               const float* centers, float radius, const int32_t* islands, int max_tries,
                               float* out_pts, void* stream) {
  if (!nm || (n > 0 && (!out_pts || !centers))) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  if (nm->flat.totalArea <= 0.0f)
    return fail(HBN_ERR_NO_AREA, "NavMesh has no navigable area, this indicates an issue with the NavMesh");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  NvtxRange nv(nm->opt.nvtx, "hbn_random_points_near");
  CallOrder order(nm, static_cast<cudaStream_t>(stream));
  const int groupsPerBlock = 256 / kRandW;
  int64_t blocks = std::min<int64_t>((n + groupsPerBlock - 1) / groupsPerBlock,
                                     static_cast<int64_t>(nm->smCount) * 64);
  k_random_near<kRandW><<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      nm->view, seed, query0, n, centers, radius, islands, max_tries, out_pts);
  nm->launches++;
  CK(cudaGetLastError());
  return HBN_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------
// host-buffer wrappers: stage through pinned memory on the handle's stream
// ------------------------------------------------------------------------------------
namespace {
struct IoPlan {
  hbn_navmesh* nm;
  size_t off = 0;
  struct Item { size_t off, bytes; const void* src; void* dst; };
  std::vector<Item> items;
  explicit IoPlan(hbn_navmesh* n) : nm(n) {}
  size_t add(size_t bytes, const void* src, void* dst) {
    const size_t o = off;
    items.push_back({o, bytes, src, dst});
    off += (bytes + 255) & ~size_t(255);
    return o;
  }
  int prepare() {
    int rc = nm->io.ensure(off);
    if (rc) return rc;
    if (off > nm->pinnedCap) {
      if (nm->pinned) cudaFreeHost(nm->pinned);
      nm->pinned = nullptr;
      nm->pinnedCap = 0;
      CK(cudaMallocHost(&nm->pinned, off + off / 4));
      nm->pinnedCap = off + off / 4;
      g_scratchEpoch++;
    }
    return HBN_OK;
  }
  template <class T> T* dev(size_t o) { return reinterpret_cast<T*>(static_cast<char*>(nm->io.p) + o); }
  // A copy between a user buffer and the pinned staging area.  One thread moves ~10 GB/s, and for the
  // cheap queries that IS the call (16 M wall-distance queries: 192 MB in, 64 MB out, 9 ms of kernels): large
  // copies are cut over a few threads.
  static void copyPar(void* dst, const void* src, size_t bytes) {
    constexpr size_t kPerThread = 4u << 20;  // (1 MB per thread measured slower: the threads are made per call)
    unsigned nt = static_cast<unsigned>(std::min<size_t>(bytes / kPerThread, 8));
    nt = std::min(nt, std::max(1u, std::thread::hardware_concurrency()));
    if (nt <= 1) {
      memcpy(dst, src, bytes);
      return;
    }
    const size_t part = ((bytes + nt - 1) / nt + 4095) & ~size_t(4095);
    std::vector<std::thread> th;
    size_t done = std::min(part, bytes);  // [0, done) is this thread's; a piece whose thread cannot be made too
    for (unsigned t = 1; t < nt; ++t) {
      const size_t o = static_cast<size_t>(t) * part;
      if (o >= bytes) break;
      const size_t b = std::min(part, bytes - o);
      try {
        th.emplace_back([=] { memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, b); });
      } catch (...) {  // out of threads: copy the piece here
        memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, b);
      }
    }
    memcpy(dst, src, done);
    for (auto& x : th) x.join();
  }
  void stage() {  // user buffers -> pinned
    for (auto& it : items)
      if (it.src) copyPar(static_cast<char*>(nm->pinned) + it.off, it.src, it.bytes);
  }
  // user buffers -> pinned -> device in pieces: the DMA of a piece runs while the next one is staged
  int stageAndEnqueueH2D() {
    constexpr size_t kPiece = 32u << 20;
    for (auto& it : items) {
      if (!it.src) continue;
      for (size_t o = 0; o < it.bytes; o += kPiece) {
        const size_t b = std::min(kPiece, it.bytes - o);
        copyPar(static_cast<char*>(nm->pinned) + it.off + o, static_cast<const char*>(it.src) + o, b);
        CK(cudaMemcpyAsync(static_cast<char*>(nm->io.p) + it.off + o, static_cast<char*>(nm->pinned) + it.off + o, b,
                           cudaMemcpyHostToDevice, nm->stream));
      }
    }
    return HBN_OK;
  }
  int enqueueH2D() {
    for (auto& it : items)
      if (it.src)
        CK(cudaMemcpyAsync(static_cast<char*>(nm->io.p) + it.off, static_cast<char*>(nm->pinned) + it.off,
                           it.bytes, cudaMemcpyHostToDevice, nm->stream));
    return HBN_OK;
  }
  int enqueueD2H() {
    for (auto& it : items)
      if (it.dst)
        CK(cudaMemcpyAsync(static_cast<char*>(nm->pinned) + it.off, static_cast<char*>(nm->io.p) + it.off,
                           it.bytes, cudaMemcpyDeviceToHost, nm->stream));
    return HBN_OK;
  }
  int finish() {  // wait, pinned -> user buffers
    CK(cudaStreamSynchronize(nm->stream));
    for (auto& it : items)
      if (it.dst) copyPar(it.dst, static_cast<char*>(nm->pinned) + it.off, it.bytes);
    return checkFault(nm);
  }
  int h2d() { return stageAndEnqueueH2D(); }
  int d2h() {
    int rc = enqueueD2H();
    if (rc) return rc;
    return finish();
  }
};
}  // namespace

extern "C" {

#define HOST_PROLOGUE                                         \
  if (!nm) return fail(HBN_ERR_INVALID, "null navmesh");      \
  if (n <= 0) return HBN_OK;                                  \
  DeviceGuard g_(nm->device);                                 \
  std::lock_guard<std::recursive_mutex> lk_(nm->mu);          \
  IoPlan io(nm);                                              \
  int rc;

int hbn_snap_point(hbn_navmesh_t nm, const float* pts, const int32_t* islands, int64_t n,
                   float* out_pts, uint32_t* out_refs, int32_t* out_islands) {
  HOST_PROLOGUE
  const size_t oP = io.add(n * 12, pts, nullptr);
  const size_t oI = islands ? io.add(n * 4, islands, nullptr) : 0;
  const size_t oOP = out_pts ? io.add(n * 12, nullptr, out_pts) : 0;
  const size_t oOR = out_refs ? io.add(n * 4, nullptr, out_refs) : 0;
  const size_t oOI = out_islands ? io.add(n * 4, nullptr, out_islands) : 0;
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if ((rc = hbn_snap_point_dev(nm, io.dev<float>(oP), islands ? io.dev<int32_t>(oI) : nullptr, n,
                               out_pts ? io.dev<float>(oOP) : nullptr,
                               out_refs ? io.dev<uint32_t>(oOR) : nullptr,
                               out_islands ? io.dev<int32_t>(oOI) : nullptr, nm->stream)))
    return rc;
  return io.d2h();
}

int hbn_is_navigable(hbn_navmesh_t nm, const float* pts, int64_t n, float max_y_delta, uint8_t* out) {
  HOST_PROLOGUE
  const size_t oP = io.add(n * 12, pts, nullptr);
  const size_t oO = io.add(n, nullptr, out);
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if ((rc = hbn_is_navigable_dev(nm, io.dev<float>(oP), n, max_y_delta, io.dev<uint8_t>(oO), nm->stream)))
    return rc;
  return io.d2h();
}

int hbn_find_path(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                  float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                  uint32_t* out_corridor, int32_t* out_ncorridor, uint32_t* out_status, int flags) {
  HOST_PROLOGUE
  const size_t oS = io.add(n * 12, starts, nullptr);
  const size_t oE = io.add(n * 12, ends, nullptr);
  const size_t oD = io.add(n * 4, nullptr, out_dist);
  const size_t oN = out_npts ? io.add(n * 4, nullptr, out_npts) : 0;
  const size_t oP = out_pts ? io.add(static_cast<size_t>(n) * max_pts * 12, nullptr, out_pts) : 0;
  const size_t oC = out_corridor ? io.add(static_cast<size_t>(n) * 256 * 4, nullptr, out_corridor) : 0;
  const size_t oNC = out_ncorridor ? io.add(n * 4, nullptr, out_ncorridor) : 0;
  const size_t oST = out_status ? io.add(n * 8, nullptr, out_status) : 0;
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if (out_pts) CK(cudaMemsetAsync(io.dev<char>(oP), 0xff, static_cast<size_t>(n) * max_pts * 12, nm->stream));
  if (out_corridor) CK(cudaMemsetAsync(io.dev<char>(oC), 0, static_cast<size_t>(n) * 256 * 4, nm->stream));
  if ((rc = hbn_find_path_dev(nm, io.dev<float>(oS), io.dev<float>(oE), n, io.dev<float>(oD),
                              out_npts ? io.dev<int32_t>(oN) : nullptr,
                              out_pts ? io.dev<float>(oP) : nullptr, max_pts,
                              out_corridor ? io.dev<uint32_t>(oC) : nullptr,
                              out_ncorridor ? io.dev<int32_t>(oNC) : nullptr,
                              out_status ? io.dev<uint32_t>(oST) : nullptr, flags, nm->stream)))
    return rc;
  return io.d2h();
}

int hbn_find_path_multigoal(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                            int g, float* out_dist, int32_t* out_index, int32_t* out_npts,
                            float* out_pts, int max_pts) {
  HOST_PROLOGUE
  if (g <= 0) return fail(HBN_ERR_INVALID, "g must be positive");
  const size_t oS = io.add(n * 12, starts, nullptr);
  const size_t oE = io.add(static_cast<size_t>(n) * g * 12, ends, nullptr);
  const size_t oD = io.add(n * 4, nullptr, out_dist);
  const size_t oI = io.add(n * 4, nullptr, out_index);
  const size_t oN = out_npts ? io.add(n * 4, nullptr, out_npts) : 0;
  const size_t oP = out_pts ? io.add(static_cast<size_t>(n) * max_pts * 12, nullptr, out_pts) : 0;
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if (out_pts) CK(cudaMemsetAsync(io.dev<char>(oP), 0xff, static_cast<size_t>(n) * max_pts * 12, nm->stream));
  if ((rc = hbn_find_path_multigoal_dev(nm, io.dev<float>(oS), io.dev<float>(oE), n, g, io.dev<float>(oD),
                                        io.dev<int32_t>(oI), out_npts ? io.dev<int32_t>(oN) : nullptr,
                                        out_pts ? io.dev<float>(oP) : nullptr, max_pts, nm->stream)))
    return rc;
  return io.d2h();
}

int hbn_try_step(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                 int allow_sliding, float* out_pts) {
  HOST_PROLOGUE
  const size_t oS = io.add(n * 12, starts, nullptr);
  const size_t oE = io.add(n * 12, ends, nullptr);
  const size_t oO = io.add(n * 12, nullptr, out_pts);
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if ((rc = hbn_try_step_dev(nm, io.dev<float>(oS), io.dev<float>(oE), n, allow_sliding, io.dev<float>(oO), nm->stream)))
    return rc;
  return io.d2h();
}

// Host-buffer env step.  The whole step -- H2D of the three inputs, the 9 kernels of hbn_env_step_dev,
// D2H of the two outputs -- is captured ONCE per (batch size, sliding mode) into a CUDA graph and
// replayed: a PointNav step at 1024 envs is a chain of small kernels whose launch overheads and
// inter-kernel gaps are a third of its wall time.  The graph is rebuilt when any scratch buffer or
// option it baked in has changed (g_scratchEpoch).
int hbn_env_step(hbn_navmesh_t nm, const float* starts, const float* targets, const float* goals, int64_t n,
                 int allow_sliding, float* out_pos, float* out_dist) {
  HOST_PROLOGUE
  if (!starts || !targets || !goals || !out_pos || !out_dist) return fail(HBN_ERR_INVALID, "null argument");
  NvtxRange nv(nm->opt.nvtx, "hbn_env_step(host)");
  const size_t oS = io.add(n * 12, starts, nullptr);
  const size_t oT = io.add(n * 12, targets, nullptr);
  const size_t oG = io.add(n * 12, goals, nullptr);
  const size_t oP = io.add(n * 12, nullptr, out_pos);
  const size_t oD = io.add(n * 4, nullptr, out_dist);
  if ((rc = checkFault(nm))) return rc;
  hbn_navmesh::StepGraph* sg = nullptr;
  for (auto& c : nm->stepGraphs)
    if (c.n == n && c.sliding == (allow_sliding != 0)) sg = &c;
  if (!sg || sg->epoch != g_scratchEpoch) {
    // everything the step will touch exists before the capture starts
    if ((rc = io.prepare()) || (rc = envStepReserve(nm, n))) return rc;
    CK(cudaStreamSynchronize(nm->stream));
    if (sg && sg->exec) cudaGraphExecDestroy(sg->exec);
    if (!sg) {
      nm->stepGraphs.push_back(hbn_navmesh::StepGraph{n, allow_sliding != 0, 0, nullptr, {0, 0, 0, 0, 0}});
      sg = &nm->stepGraphs.back();
    }
    sg->exec = nullptr;
    const int64_t l0 = nm->launches;
    const uint64_t epoch = g_scratchEpoch;
    CK(cudaStreamBeginCapture(nm->stream, cudaStreamCaptureModeThreadLocal));
    rc = io.enqueueH2D();
    if (!rc)
      rc = hbn_env_step_dev(nm, io.dev<float>(oS), io.dev<float>(oT), io.dev<float>(oG), n, allow_sliding,
                            io.dev<float>(oP), io.dev<float>(oD), nm->stream);
    if (!rc) rc = io.enqueueD2H();
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(nm->stream, &graph);
    if (rc || ce != cudaSuccess || epoch != g_scratchEpoch) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (!rc) rc = fail(HBN_ERR_CUDA, epoch != g_scratchEpoch ? "scratch moved during the env step capture"
                                                                 : std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
      return rc;
    }
    const cudaError_t ie = cudaGraphInstantiate(&sg->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(HBN_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    sg->epoch = epoch;
    sg->offs[0] = static_cast<size_t>(nm->launches - l0);  // kernels per replay
    nm->launches = l0;
  }
  io.stage();
  if (nm->lastValid && nm->lastStream != nm->stream) CK(cudaStreamWaitEvent(nm->stream, nm->lastDone, 0));
  CK(cudaGraphLaunch(sg->exec, nm->stream));
  nm->launches += static_cast<int64_t>(sg->offs[0]);
  if (cudaEventRecord(nm->lastDone, nm->stream) == cudaSuccess) {
    nm->lastStream = nm->stream;
    nm->lastValid = true;
  }
  return io.finish();
}

int hbn_follower_best_prims(hbn_navmesh_t nm, const double* rots, const double* poss, const float* goals, int64_t n,
                            const hbn_follower_params* fp, int32_t* out_prim, float* out_geo) {
  HOST_PROLOGUE
  const size_t oR = io.add(n * 32, rots, nullptr);
  const size_t oP = io.add(n * 24, poss, nullptr);
  const size_t oG = io.add(n * 12, goals, nullptr);
  const size_t oO = io.add(n * 4, nullptr, out_prim);
  const size_t oD = out_geo ? io.add(n * 4, nullptr, out_geo) : 0;
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if ((rc = hbn_follower_best_prims_dev(nm, io.dev<double>(oR), io.dev<double>(oP), io.dev<float>(oG), n, fp,
                                        io.dev<int32_t>(oO), out_geo ? io.dev<float>(oD) : nullptr, nm->stream)))
    return rc;
  return io.d2h();
}

int hbn_closest_obstacle(hbn_navmesh_t nm, const float* pts, int64_t n, float max_radius,
                         float* out_hit_pos, float* out_hit_normal, float* out_hit_dist) {
  HOST_PROLOGUE
  const size_t oP = io.add(n * 12, pts, nullptr);
  const size_t oHP = out_hit_pos ? io.add(n * 12, nullptr, out_hit_pos) : 0;
  const size_t oHN = out_hit_normal ? io.add(n * 12, nullptr, out_hit_normal) : 0;
  const size_t oHD = io.add(n * 4, nullptr, out_hit_dist);
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if ((rc = hbn_closest_obstacle_dev(nm, io.dev<float>(oP), n, max_radius,
                                     out_hit_pos ? io.dev<float>(oHP) : nullptr,
                                     out_hit_normal ? io.dev<float>(oHN) : nullptr, io.dev<float>(oHD), nm->stream)))
    return rc;
  return io.d2h();
}

int hbn_random_points(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n,
                      const int32_t* islands, int max_tries, float* out_pts, uint32_t* out_refs) {
  HOST_PROLOGUE
  const size_t oI = islands ? io.add(n * 4, islands, nullptr) : 0;
  const size_t oP = io.add(n * 12, nullptr, out_pts);
  const size_t oR = out_refs ? io.add(n * 4, nullptr, out_refs) : 0;
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if ((rc = hbn_random_points_dev(nm, seed, query0, n, islands ? io.dev<int32_t>(oI) : nullptr, max_tries,
                                  io.dev<float>(oP), out_refs ? io.dev<uint32_t>(oR) : nullptr, nm->stream)))
    return rc;
  return io.d2h();
}

int hbn_random_points_near(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n, const float* centers,
                           float radius, const int32_t* islands, int max_tries, float* out_pts) {
  HOST_PROLOGUE
  const size_t oC = io.add(n * 12, centers, nullptr);
  const size_t oI = islands ? io.add(n * 4, islands, nullptr) : 0;
  const size_t oP = io.add(n * 12, nullptr, out_pts);
  if ((rc = io.prepare()) || (rc = io.h2d())) return rc;
  if ((rc = hbn_random_points_near_dev(nm, seed, query0, n, io.dev<float>(oC), radius,
                                       islands ? io.dev<int32_t>(oI) : nullptr, max_tries, io.dev<float>(oP),
                                       nm->stream)))
    return rc;
  return io.d2h();
}

}  // extern "C"

// C ABI of libhbn.so (include/hbn.h): navmesh upload, kernel orchestration, host-buffer
// wrappers.  No CPU query path exists in this library.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hbn.h"
#include "hbn_host.h"
#include "hbn_kernels.cuh"
#include "hbn_astar_group.cuh"
#include "hbn_astar_lane.cuh"
#include "hbn_snap.cuh"
#include <cub/device/device_scan.cuh>

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
// find_path (hbn_astar_warp.cuh): open-list capacity of the two tiers, warps per block
constexpr int kOpenS = 256;
constexpr int kOpenL = 2048;
constexpr int kFpWpb = 1;
// per chunk of a find_path call: 16 counters, then the class histogram and the scatter cursors (k_fp_classify)
constexpr size_t kFpCounterBytes = (16 + 2 * 32) * 4;
constexpr int kLaneTS = 63;    // lane-per-query search: heap entries per lane kept in shared memory (6 levels)
constexpr int kLaneMinB = 16;  // ... and resident warps per SM its register budget must allow
constexpr int kSnapW = 8;     // lanes per point in k_snap
constexpr int kRandW = 8;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return HBN_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(HBN_ERR_CUDA, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
    cap = want;
    return HBN_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct hbn_navmesh {
  int device = 0;
  FlatNav flat;
  NavView view{};
  std::vector<void*> devArrays;
  int64_t deviceBytes = 0;
  cudaStream_t stream = nullptr;  // for the host-buffer entry points
  int smCount = 0;
  int64_t launches = 0;
  std::recursive_mutex mu;
  // scratch (device)
  DevBuf sG, eG, e2G, sPt, ePt, epPt, lastPoly, lists, counters, wsL, wsFp, io, work, mgDist, mgBounds, mgOrder, mgEnd, mgMask;
  // lock-step find_path (hbn_astar_group.cuh): class, search list, status, corridor rings, node records
  DevBuf fpCls, fpBucket, fpWork, fpStat, fpLen, fpCorr, wsFpG;
  int fpG = 1;          // lanes per query of k_astar_g; 0 = one query per warp (k_findpath_w tiers only);
                        // 1 = one query per LANE (k_astar_lane, hbn_astar_lane.cuh)
  int blocksFpG = 0;
  // lane-per-query search: per-lane node table + records in HBM, allocated on first use
  DevBuf snapCnt, snapOff, snapG, snapQ, snapD, snapOut, snapBest, snapTmp, snapTodo;  // candidate-list snap (hbn_snap.cuh)
  DevBuf wsLane, laneGen;
  int blocksFpLane = 0;
  int laneCfg = 0;      // HBN_LANE_CFG: shared heap levels / warps per SM variant (tuning)
  bool laneSpread = true;   // batches smaller than the grid use fewer lanes per warp (HBN_LANE_SPREAD=0: always 32)
  // pinned staging for the host-buffer entry points
  void* pinned = nullptr;
  size_t pinnedCap = 0;
  int blocksFpS = 0, blocksFpL = 0, blocksWallS = 0, blocksWallL = 0;
  // optional phase timing of hbn_find_path_dev (hbn_navmesh_set_profiling)
  unsigned int* faultHost = nullptr;  // mapped pinned memory the kernels report bugs through
  unsigned int* faultDev = nullptr;
  bool profile = false;
  struct PhaseEv { cudaEvent_t e[3]; };
  std::vector<PhaseEv> phaseEvents;
};

namespace {

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

// Per-lane node tables + records of k_astar_lane: blocksFpLane * 32 slots, allocated (and zeroed:
// generation 0, empty tables) on first use.  Shrinks the grid when HBM is short.
int laneScratch(hbn_navmesh* nm, cudaStream_t st, LaneScratch* out) {
  const size_t tabB = laneTabBytes(nm->view.numKeys);
  const size_t per = laneScratchBytes(nm->view.numKeys);
  if (!nm->wsLane.p) {
    size_t freeB = 0, totalB = 0;
    CK(cudaMemGetInfo(&freeB, &totalB));
    const size_t maxBlocks = (freeB / 2) / (per * 32);
    if (maxBlocks < 1) return fail(HBN_ERR_CUDA, "not enough device memory for the find_path node tables");
    if (static_cast<size_t>(nm->blocksFpLane) > maxBlocks) nm->blocksFpLane = static_cast<int>(maxBlocks);
    const size_t lanes = static_cast<size_t>(nm->blocksFpLane) * 32;
    int rc;
    if ((rc = nm->wsLane.ensure(lanes * per)) || (rc = nm->laneGen.ensure(lanes * 4))) return rc;
    CK(cudaMemsetAsync(nm->wsLane.p, 0, lanes * tabB, st));  // the tables come first
    CK(cudaMemsetAsync(nm->laneGen.p, 0, lanes * 4, st));
  }
  const size_t lanes = static_cast<size_t>(nm->blocksFpLane) * 32;
  char* p = static_cast<char*>(nm->wsLane.p);
  out->tab = p;
  out->rec = p + lanes * tabB;
  out->heap = out->rec + lanes * kLaneRecBytes;
  out->gen = static_cast<uint32_t*>(nm->laneGen.p);
  out->tabBytes = tabB;
  return HBN_OK;
}

// lane-per-query search instantiations (HBN_LANE_CFG; 0 = shipped).  Template arguments: heap
// entries in shared memory, resident warps per SM (or registers for k_astar_lane_r), links per
// load stage, code variant V of LaneSearch.  All are bit-exact (tests/test_zz_tuning_variants.py);
// times of the find_path phase on 200 k C4 queries are in profiles/r1_summary.md.
const void* laneKernel(int cfg, size_t* shared) {
#define HBN_LANE_CASE(N, TS, ...) \
  case N: *shared = laneSharedBytes<TS>(); return reinterpret_cast<const void*>(&__VA_ARGS__);
  switch (cfg) {
    // measured in round 1, none faster than the shipped one
    HBN_LANE_CASE(1, 63, k_astar_lane<63, 17, 3>)      // 150.5 ms per 1 M (shipped 153.7 in that run)
    HBN_LANE_CASE(2, 31, k_astar_lane<31, 20, 2>)      // 158.7
    HBN_LANE_CASE(3, 63, k_astar_lane<63, 18, 2>)      // 169.6
    HBN_LANE_CASE(4, 31, k_astar_lane<31, 17, 3>)      // 202.0
    HBN_LANE_CASE(5, 63, k_astar_lane<63, 16, 4, 2>)   // heap code variant 2: 38.97 ms per 200 k (shipped 37.22)
    HBN_LANE_CASE(6, 31, k_astar_lane<31, 16, 4, 2>)   // 49.92
    HBN_LANE_CASE(7, 31, k_astar_lane<31, 20, 2, 2>)   // 46.20
    HBN_LANE_CASE(8, 47, k_astar_lane<47, 20, 3>)      // 38.13 (shipped 37.00 in that run)
    HBN_LANE_CASE(12, 39, k_astar_lane<39, 24, 2>)     // 46.52
    HBN_LANE_CASE(13, 63, k_astar_lane<63, 16, 4, 3>)  // node-table prefetch before the sift-down: 37.93
    HBN_LANE_CASE(15, 63, k_astar_lane<63, 16, 4, 5>)  // + grandchildren prefetch in the sift-down: 37.59
    HBN_LANE_CASE(16, 47, k_astar_lane<47, 20, 3, 5>)  // 38.92
    // built and host-tested, not measured yet
    HBN_LANE_CASE(9, 47, k_astar_lane<47, 20, 2>)
    HBN_LANE_CASE(10, 55, k_astar_lane<55, 19, 3>)
    HBN_LANE_CASE(11, 47, k_astar_lane<47, 20, 3, 2>)
    HBN_LANE_CASE(14, 47, k_astar_lane<47, 20, 3, 3>)
    // register budgets ptxas does not pick by itself ("17-20 one-warp blocks" become 96 registers):
    // 112 = 17 warps at 63 entries (12 KB + 1 KB reserved per block), 18 at 59; 104 = 19 warps
    HBN_LANE_CASE(17, 63, k_astar_lane_r<63, 112, 4>)
    HBN_LANE_CASE(18, 55, k_astar_lane_r<55, 104, 4>)
    HBN_LANE_CASE(19, 55, k_astar_lane_r<55, 104, 3>)
    HBN_LANE_CASE(22, 59, k_astar_lane_r<59, 112, 4>)
    // modify scan looks through the shared part of the heap first (V = 6)
    HBN_LANE_CASE(20, 63, k_astar_lane<63, 16, 4, 6>)
    HBN_LANE_CASE(21, 63, k_astar_lane_r<63, 112, 4, 6>)
    // node-table accesses with an L2 evict_last policy (V = 7)
    HBN_LANE_CASE(24, 63, k_astar_lane<63, 16, 4, 7>)
    // 95 entries (85 % of the pops find the whole open list in shared memory) at 11 warps per SM
    HBN_LANE_CASE(23, 95, k_astar_lane<95, 11, 4>)
    default: *shared = laneSharedBytes<kLaneTS>(); return reinterpret_cast<const void*>(&k_astar_lane<kLaneTS, kLaneMinB, 4>);
  }
#undef HBN_LANE_CASE
}

const void* groupKernel(int g) {
  switch (g) {
    case 4: return reinterpret_cast<const void*>(&k_astar_g<4, kOpenS>);
    case 16: return reinterpret_cast<const void*>(&k_astar_g<16, kOpenS>);
    case 32: return reinterpret_cast<const void*>(&k_astar_g<32, kOpenS>);
    default: return reinterpret_cast<const void*>(&k_astar_g<8, kOpenS>);
  }
}

int finishCreate(HostNavMesh& mesh, const int32_t* islands, int device, hbn_navmesh_t* out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(HBN_ERR_NO_DEVICE, "no CUDA device available (libhbn has no CPU path)");
  if (device < 0 || device >= ndev) return fail(HBN_ERR_INVALID, "bad device index");
  CK(cudaSetDevice(device));
  hbn_navmesh* nm = new hbn_navmesh();
  nm->device = device;
  mesh.finish(islands);
  mesh.flatten(nm->flat);
  const FlatNav& f = nm->flat;
  if (f.polys.size() >= (1u << 24) || f.links.size() >= (1u << 27)) {
    delete nm;
    return fail(HBN_ERR_LIMIT, "navmesh exceeds 2^24 polys or 2^27 links");
  }
  for (const PolyRec& p : f.polys)
    if (p.linkCount > 31) {
      delete nm;
      return fail(HBN_ERR_LIMIT, "a polygon has more than 31 links");
    }
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
  if (rc != HBN_OK) {
    hbn_navmesh_destroy(nm);
    return rc;
  }
  nm->view = v;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  nm->smCount = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&nm->stream, cudaStreamNonBlocking));

  // opt in to large dynamic shared memory and size the persistent grids from occupancy
  const int threads = kFpWarps * 32;
  const size_t smL = kFpWarps * wsSharedBytes(kCapL, kWsHybrid);
  const size_t smW = kFpWarps * wsSharedBytes(kWallCapS, kWsShared);
  const size_t smFpS = kFpWpb * WarpWs<kOpenS>::sharedBytes();
  const size_t smFpL = kFpWpb * WarpWs<kOpenL>::sharedBytes();
  CK(cudaFuncSetAttribute(k_findpath_w<kOpenS, kFpWpb>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smFpS)));
  CK(cudaFuncSetAttribute(k_findpath_w<kOpenL, kFpWpb>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smFpL)));
  CK(cudaFuncSetAttribute(k_findpath_w<kOpenS, kFpWpb>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(k_wall<kWallCapS, kWsShared>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smW)));
  CK(cudaFuncSetAttribute(k_wall<kCapL, kWsHybrid>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smL)));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_findpath_w<kOpenS, kFpWpb>, 32 * kFpWpb, smFpS));
  nm->blocksFpS = std::max(1, occ) * nm->smCount;
  if (const char* e = getenv("HBN_FP_BLOCKS_PER_SM")) nm->blocksFpS = std::max(1, std::min(occ, atoi(e))) * nm->smCount;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_findpath_w<kOpenL, kFpWpb>, 32 * kFpWpb, smFpL));
  nm->blocksFpL = std::max(1, occ) * nm->smCount;
  // lock-step search: group width from HBN_FP_G (4, 8, 16, 32; "warp" = the one-query-per-warp tiers)
  if (const char* e = getenv("HBN_FP_G"))
    nm->fpG = (strcmp(e, "warp") == 0) ? 0 : (strcmp(e, "lane") == 0) ? 1 : atoi(e);
  if (nm->fpG != 0 && nm->fpG != 1 && nm->fpG != 4 && nm->fpG != 8 && nm->fpG != 16 && nm->fpG != 32) nm->fpG = 8;
  {
    if (const char* e = getenv("HBN_LANE_CFG")) nm->laneCfg = atoi(e);
    if (const char* e = getenv("HBN_LANE_SPREAD")) nm->laneSpread = atoi(e) != 0;
    size_t smLane = 0;
    const void* fn = laneKernel(nm->laneCfg, &smLane);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smLane)));
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, smLane));
    nm->blocksFpLane = std::max(1, occ) * nm->smCount;
    if (const char* e = getenv("HBN_FP_BLOCKS_PER_SM")) nm->blocksFpLane = std::max(1, std::min(occ, atoi(e))) * nm->smCount;
  }
  if (nm->fpG > 1) {
    const size_t smG = (32 / nm->fpG) * gGroupSharedBytes<kOpenS>();
    const void* fn = groupKernel(nm->fpG);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smG)));
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, smG));
    nm->blocksFpG = std::max(1, occ) * nm->smCount;
    if (const char* e = getenv("HBN_FP_BLOCKS_PER_SM")) nm->blocksFpG = std::max(1, std::min(occ, atoi(e))) * nm->smCount;
  }
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wall<kWallCapS, kWsShared>, threads, smW));
  nm->blocksWallS = std::max(1, occ) * nm->smCount;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wall<kCapL, kWsHybrid>, threads, smL));
  nm->blocksWallL = std::max(1, occ) * nm->smCount;
  // global scratch: one slot per resident warp
  rc = nm->wsL.ensure(static_cast<size_t>(nm->blocksWallL) * kFpWarps * wsGlobalBytes(kCapL, kWsHybrid));
  if (rc == HBN_OK)
    rc = nm->wsFp.ensure(static_cast<size_t>(std::max(nm->blocksFpS, nm->blocksFpL)) * kFpWpb *
                         WarpWs<kOpenS>::globalBytes());
  if (rc == HBN_OK && nm->fpG > 1)
    rc = nm->wsFpG.ensure(static_cast<size_t>(nm->blocksFpG) * (32 / nm->fpG) * gGroupGlobalBytes());
  if (rc == HBN_OK) rc = nm->counters.ensure(64);
  if (rc == HBN_OK) rc = nm->work.ensure(64);
  if (rc == HBN_OK) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&nm->faultHost), 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&nm->faultDev), nm->faultHost, 0) != cudaSuccess)
      rc = fail(HBN_ERR_CUDA, "cudaHostAlloc(fault flags)");
    else
      memset(nm->faultHost, 0, 64);
  }
  if (rc == HBN_OK && cudaMemset(nm->work.p, 0, 64) != cudaSuccess) rc = fail(HBN_ERR_CUDA, "cudaMemset");
  if (rc != HBN_OK) {
    hbn_navmesh_destroy(nm);
    return rc;
  }
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

struct DeviceGuard {
  int prev = 0;
  bool ok;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

constexpr int64_t kSnapChunk = 1 << 18;   // points per pass of the candidate-list pipeline
constexpr int64_t kSnapSmall = 4096;      // below this one k_snap<8> launch is cheaper than five kernels + a scan

// Two small independent projectToPoly batches in one k_snap_dual launch (HBN_SNAP_DUAL=0: off).
// Returns false when the pair does not qualify (the caller then launches them one after the
// other).  1024 C2 points: find_path's snap phase 71 -> 38 us, with spreading 28 us.
bool snapDualLaunch(hbn_navmesh* nm, const float* ptsA, int64_t nA, float* outPtsA, uint32_t* outGA,
                    const float* ptsB, int64_t nB, float* outPtsB, uint32_t* outGB, cudaStream_t st, int* rc) {
  *rc = HBN_OK;
  const char* e = getenv("HBN_SNAP_DUAL");
  if ((e && atoi(e) == 0) || nA <= 0 || nB <= 0 || nA >= kSnapSmall || nB >= kSnapSmall || getenv("HBN_SNAP_GROUP"))
    return false;
  const char* sp = getenv("HBN_SNAP_SPREAD");
  const bool spread = !sp || atoi(sp) != 0;
  const int64_t n = nA + nB;
  const unsigned threads = spread ? kSnapW : 256;
  const int64_t gpb = threads / kSnapW;
  const int64_t blocks = (n + gpb - 1) / gpb;
  k_snap_dual<kSnapW><<<static_cast<unsigned>(blocks), threads, 0, st>>>(nm->view, SnapJob{ptsA, nA, outPtsA, outGA},
                                                                        SnapJob{ptsB, nB, outPtsB, outGB});
  nm->launches++;
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) *rc = fail(HBN_ERR_CUDA, std::string("k_snap_dual: ") + cudaGetErrorString(ce));
  return true;
}

// projectToPoly for n points.  Large batches: count -> scan -> fill -> eval -> select
// (hbn_snap.cuh), nothing read back by the host; small ones: the lane-group kernel.
int snapLaunch(hbn_navmesh* nm, const float* pts, const int32_t* islands, int64_t n, float* out_pts,
               uint32_t* out_g, uint32_t* out_refs, int32_t* out_isl, uint8_t* out_nav,
               float maxYDelta, cudaStream_t st) {
  if (n <= 0) return HBN_OK;
  const int groupsPerBlock = 256 / kSnapW;
  const int64_t maxBlocks = static_cast<int64_t>(nm->smCount) * 64;
  const bool forceGroup = getenv("HBN_SNAP_GROUP") != nullptr;  // testing: the lane-group kernel only
  if (n < kSnapSmall || forceGroup) {
    int64_t blocks = std::min(maxBlocks, (n + groupsPerBlock - 1) / groupsPerBlock);
    unsigned threads = 256;
    // one lane group per warp (blocks of 8 threads): the groups of a small batch neither share a
    // warp's issue slots nor diverge against each other, and 1024 points cover the SMs instead
    // of 32 blocks (HBN_SNAP_SPREAD=0: blocks of 256 threads)
    const char* sp = getenv("HBN_SNAP_SPREAD");
    const bool spread = !sp || atoi(sp) != 0;
    if (spread && n < kSnapSmall) {
      threads = kSnapW;
      blocks = n;
    }
    k_snap<kSnapW><<<static_cast<unsigned>(blocks), threads, 0, st>>>(nm->view, pts, islands, n, out_pts, out_g,
                                                                      out_refs, out_isl, out_nav, maxYDelta, nullptr);
    nm->launches++;
    CK(cudaGetLastError());
    return HBN_OK;
  }
  const int64_t cmax = std::min(n, kSnapChunk);
  size_t cap = static_cast<size_t>(cmax) * kSnapAvgCap;
  if (const char* e = getenv("HBN_SNAP_CAP")) cap = static_cast<size_t>(std::max(1, atoi(e)));  // testing: force the fallback
  size_t tmpBytes = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                   static_cast<int>(cmax + 1), st));
  int rc;
  if ((rc = nm->snapCnt.ensure((cmax + 1) * 4)) || (rc = nm->snapOff.ensure((cmax + 1) * 4)) ||
      (rc = nm->snapG.ensure(cap * 4)) || (rc = nm->snapQ.ensure(cap * 4)) || (rc = nm->snapD.ensure(cap * 4)) ||
      (rc = nm->snapOut.ensure(cap * 4)) || (rc = nm->snapBest.ensure(cmax * 8)) || (rc = nm->snapTmp.ensure(tmpBytes)) ||
      (rc = nm->snapTodo.ensure(16)))
    return rc;
  uint32_t* cnt = static_cast<uint32_t*>(nm->snapCnt.p);
  uint32_t* off = static_cast<uint32_t*>(nm->snapOff.p);
  uint32_t* cg = static_cast<uint32_t*>(nm->snapG.p);
  uint32_t* cq = static_cast<uint32_t*>(nm->snapQ.p);
  float* cd = static_cast<float*>(nm->snapD.p);
  float* clb = static_cast<float*>(nm->snapOut.p);
  uint32_t* best = static_cast<uint32_t*>(nm->snapBest.p);
  float* rxz = reinterpret_cast<float*>(best + cmax);
  uint32_t* todo = static_cast<uint32_t*>(nm->snapTodo.p);
  for (int64_t c0 = 0; c0 < n; c0 += kSnapChunk) {
    const int64_t cn = std::min(kSnapChunk, n - c0);
    const float* p = pts + 3 * c0;
    const int32_t* isl = islands ? islands + c0 : nullptr;
    const unsigned pb = static_cast<unsigned>((cn + 255) / 256);
    CK(cudaMemsetAsync(cnt + cn, 0, 4, st));
    k_snap_count<<<pb, 256, 0, st>>>(nm->view, p, isl, cn, cnt, best, rxz);
    CK(cub::DeviceScan::ExclusiveSum(nm->snapTmp.p, tmpBytes, cnt, off, static_cast<int>(cn + 1), st));
    k_snap_fill<<<pb, 256, 0, st>>>(nm->view, p, cn, off, static_cast<uint32_t>(cap), rxz, cg, cq, clb);
    for (int pass = 0; pass < 2; ++pass)
      k_snap_eval<<<static_cast<unsigned>(nm->smCount * 16), 256, 0, st>>>(
          nm->view, p, isl, cn, off, static_cast<uint32_t>(cap), cg, cq, clb, pass, cd, best);
    k_snap_select<<<pb, 256, 0, st>>>(nm->view, p, isl, cn, off, static_cast<uint32_t>(cap), cg, cd,
                                      out_pts ? out_pts + 3 * c0 : nullptr, out_g ? out_g + c0 : nullptr,
                                      out_refs ? out_refs + c0 : nullptr, out_isl ? out_isl + c0 : nullptr,
                                      out_nav ? out_nav + c0 : nullptr, maxYDelta, todo);
    // redone here only if the chunk's candidates did not fit the scratch (decided on the device)
    const int64_t blocks = std::min(maxBlocks, (cn + groupsPerBlock - 1) / groupsPerBlock);
    k_snap<kSnapW><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        nm->view, p, isl, cn, out_pts ? out_pts + 3 * c0 : nullptr, out_g ? out_g + c0 : nullptr,
        out_refs ? out_refs + c0 : nullptr, out_isl ? out_isl + c0 : nullptr, out_nav ? out_nav + c0 : nullptr,
        maxYDelta, todo);
    nm->launches += 8;  // 6 kernels here + cub's scan (2 kernels)
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
  return finishCreate(mesh, nullptr, device, out);
}

int hbn_navmesh_create_from_tiles(const hbn_tile_blob* tiles, int n_tiles, const float* params5,
                                  int max_tiles, int max_polys, const int32_t* poly_islands,
                                  int device, hbn_navmesh_t* out) {
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
  return finishCreate(mesh, poly_islands, device, out);
}

void hbn_navmesh_destroy(hbn_navmesh_t nm) {
  if (!nm) return;
  DeviceGuard g(nm->device);
  for (void* d : nm->devArrays) cudaFree(d);
  for (DevBuf* b : {&nm->sG, &nm->eG, &nm->e2G, &nm->sPt, &nm->ePt, &nm->epPt, &nm->lastPoly,
                    &nm->lists, &nm->counters, &nm->wsL, &nm->wsFp, &nm->io, &nm->work, &nm->mgDist,
                    &nm->mgBounds, &nm->mgOrder, &nm->mgEnd, &nm->mgMask, &nm->fpCls, &nm->fpWork, &nm->fpStat,
                    &nm->fpLen, &nm->fpCorr, &nm->fpBucket, &nm->wsFpG, &nm->wsLane, &nm->laneGen, &nm->snapCnt, &nm->snapOff,
                    &nm->snapG, &nm->snapQ, &nm->snapD, &nm->snapOut, &nm->snapBest, &nm->snapTmp, &nm->snapTodo})
    b->release();
  if (nm->pinned) cudaFreeHost(nm->pinned);
  if (nm->faultHost) cudaFreeHost(nm->faultHost);
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
  if (getenv("HBN_DEBUG_SLOW") && nm->faultHost) {
    fprintf(stderr, "[hbn] slowest find_path query: %u kcycles, q=%u, expansions=%u, tier OC=%u\n",
            nm->faultHost[4], nm->faultHost[5], nm->faultHost[6], nm->faultHost[7]);
    nm->faultHost[4] = 0;
  }
  CK(cudaMemcpy(out8, nm->work.p, 64, cudaMemcpyDeviceToHost));
  if (reset) CK(cudaMemset(nm->work.p, 0, 64));
  return checkFault(nm);
}

int64_t hbn_navmesh_triangles(hbn_navmesh_t nm, int island, float* out, int64_t cap_tris) {
  if (!nm) return -1;
  const FlatNav& f = nm->flat;
  int64_t n = 0;
  for (const PolyRec& p : f.polys) {
    if ((p.areaType >> 6) == 1) continue;
    if ((p.flags & kFlagWalk) == 0) continue;
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
  return snapLaunch(nm, pts, islands, n, out_pts, nullptr, out_refs, out_islands, nullptr, 0.f,
                    static_cast<cudaStream_t>(stream));
}

int hbn_is_navigable_dev(hbn_navmesh_t nm, const float* pts, int64_t n, float max_y_delta,
                         uint8_t* out, void* stream) {
  if (!nm || (n > 0 && (!pts || !out))) return fail(HBN_ERR_INVALID, "null argument");
  DeviceGuard g(nm->device);
  return snapLaunch(nm, pts, nullptr, n, nullptr, nullptr, nullptr, nullptr, out, max_y_delta,
                    static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// one-query-per-warp tiers (hbn_astar_warp.cuh) over `work` (nullptr: all n queries): search +
// funnel + outputs in one kernel.  cnt: 4 zeroed counters.
static int warpTierLaunch(hbn_navmesh_t nm, FindPathArgs a, bool smallTier, uint32_t* cnt, cudaStream_t st) {
  a.counter = cnt + 0;
  a.overflow = static_cast<uint32_t*>(nm->lists.p);
  a.overflowCount = cnt + 1;
  a.scratch = static_cast<char*>(nm->wsFp.p);
  if (smallTier) {
    int64_t blocks = std::min<int64_t>(nm->blocksFpS, (a.n + kFpWpb - 1) / kFpWpb);
    k_findpath_w<kOpenS, kFpWpb><<<static_cast<unsigned>(blocks), 32 * kFpWpb,
                                   kFpWpb * WarpWs<kOpenS>::sharedBytes(), st>>>(nm->view, a);
    nm->launches++;
    CK(cudaGetLastError());
    // large tier over the overflow list (its length stays on the device)
    a.work = a.overflow;
    a.workCount = a.overflowCount;
    a.counter = cnt + 2;
    a.overflowCount = cnt + 3;  // cannot overflow: open list <= kMaxNodes
  }
  k_findpath_w<kOpenL, kFpWpb><<<nm->blocksFpL, 32 * kFpWpb, kFpWpb * WarpWs<kOpenL>::sharedBytes(), st>>>(nm->view, a);
  nm->launches++;
  CK(cudaGetLastError());
  return HBN_OK;
}

constexpr int64_t kFpChunk = 1 << 20;  // queries per pass of the lock-step pipeline (1 KB corridor ring each)

// n (start, end) pairs; startDiv > 1: pair q uses start q / startDiv (multi-goal layout)
// pairMask (lock-step pipeline only): queries with a zero byte are skipped -- their outputs are left
// alone, or get an infinite distance with kFpFillSkipped.  kFpReuseSnaps: the projectToPoly results
// of the previous call on the same points are still in the scratch.
enum { kFpReuseSnaps = 1 << 16, kFpFillSkipped = 1 << 17 };
static int findPathLaunch(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                          int startDiv, float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                          uint32_t* out_corridor, int32_t* out_ncorridor, uint32_t* out_status,
                          int flags, void* stream, const uint8_t* pairMask = nullptr) {
  if (!nm || (n > 0 && (!starts || !ends || !out_dist))) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  const int64_t nStarts = startDiv > 1 ? n / startDiv : n;
  if (n >= (1ll << 31)) return fail(HBN_ERR_INVALID, "batch too large");
  if (out_pts && max_pts <= 0) return fail(HBN_ERR_INVALID, "max_pts must be positive");
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = checkFault(nm))) return rc;  // reported by an earlier launch
  const int64_t chunk = startDiv > 1 ? std::max<int64_t>(1, kFpChunk / startDiv) * startDiv : kFpChunk;
  const int64_t nChunks = nm->fpG ? (n + chunk - 1) / chunk : 1;
  const int64_t cmax = std::min(n, chunk);
  if ((rc = nm->sG.ensure(nStarts * 4)) || (rc = nm->eG.ensure(n * 4)) || (rc = nm->sPt.ensure(nStarts * 12)) ||
      (rc = nm->ePt.ensure(n * 12)) || (rc = nm->lists.ensure(cmax * 4)) ||
      (rc = nm->counters.ensure(static_cast<size_t>(nChunks) * kFpCounterBytes)))
    return rc;
  if (nm->fpG &&
      ((rc = nm->fpCls.ensure(cmax)) || (rc = nm->fpBucket.ensure(cmax)) || (rc = nm->fpWork.ensure(cmax * 4)) || (rc = nm->fpStat.ensure(cmax * 4)) ||
       (rc = nm->fpLen.ensure(cmax * 4)) || (rc = nm->fpCorr.ensure(static_cast<size_t>(cmax) * kMaxPathPolys * 4))))
    return rc;
  CK(cudaMemsetAsync(nm->counters.p, 0, static_cast<size_t>(nChunks) * kFpCounterBytes, st));
  hbn_navmesh::PhaseEv pe{};
  if (nm->profile) {
    for (auto& e : pe.e) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(pe.e[0], st));
  }
  if (pairMask && !nm->fpG) return fail(HBN_ERR_INVALID, "pair masks need the lock-step find_path pipeline");
  if ((flags & kFpReuseSnaps) == 0) {
    if (snapDualLaunch(nm, starts, nStarts, static_cast<float*>(nm->sPt.p), static_cast<uint32_t*>(nm->sG.p), ends, n,
                       static_cast<float*>(nm->ePt.p), static_cast<uint32_t*>(nm->eG.p), st, &rc)) {
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
  if (nm->profile) CK(cudaEventRecord(pe.e[1], st));
  unsigned long long* workCtr = (flags & HBN_FP_COUNT_WORK) ? static_cast<unsigned long long*>(nm->work.p) : nullptr;
  for (int64_t ci = 0; ci < nChunks; ++ci) {
    const int64_t c0 = nm->fpG ? ci * chunk : 0;
    const int64_t cn = nm->fpG ? std::min(chunk, n - c0) : n;
    const int64_t s0 = startDiv > 1 ? c0 / startDiv : c0;
    uint32_t* cnt = static_cast<uint32_t*>(nm->counters.p) + ci * (kFpCounterBytes / 4);
    FindPathArgs a{};
    a.starts = starts + 3 * s0; a.ends = ends + 3 * c0;
    a.sG = static_cast<uint32_t*>(nm->sG.p) + s0; a.sPt = static_cast<float*>(nm->sPt.p) + 3 * s0;
    a.eG = static_cast<uint32_t*>(nm->eG.p) + c0; a.ePt = static_cast<float*>(nm->ePt.p) + 3 * c0;
    a.n = cn;
    a.work = nullptr; a.workCount = nullptr;
    a.out_dist = out_dist + c0;
    a.out_npts = out_npts ? out_npts + c0 : nullptr;
    a.out_pts = out_pts ? out_pts + static_cast<size_t>(c0) * max_pts * 3 : nullptr;
    a.max_pts = max_pts;
    a.out_corridor = out_corridor ? out_corridor + static_cast<size_t>(c0) * kMaxPathPolys : nullptr;
    a.out_ncorridor = out_ncorridor ? out_ncorridor + c0 : nullptr;
    a.out_status = out_status ? out_status + 2 * c0 : nullptr;
    a.fault = nm->faultDev;
    a.startDiv = startDiv;
    a.fastFail = (flags & HBN_FP_EXACT_STATUS) ? 0 : 1;
    a.workCtr = workCtr;
    if (!nm->fpG) {
      if ((rc = warpTierLaunch(nm, a, true, cnt, st))) return rc;
      continue;
    }
    // lock-step pipeline: classify -> search -> funnel (+ the 2048-entry tier for overflows)
    uint8_t* cls = static_cast<uint8_t*>(nm->fpCls.p);
    uint32_t* work = static_cast<uint32_t*>(nm->fpWork.p);
    uint8_t* bucket = static_cast<uint8_t*>(nm->fpBucket.p);
    k_fp_classify<<<static_cast<unsigned>((cn + 255) / 256), 256, 0, st>>>(
        nm->view, a.sG, a.sPt, a.eG, a.ePt, cn, startDiv, pairMask ? pairMask + c0 : nullptr, cls, bucket, cnt + 16,
        cnt + 4);
    k_fp_scatter<<<static_cast<unsigned>((cn + 255) / 256), 256, 0, st>>>(bucket, cn, cnt + 16, cnt + 16 + kFpBuckets,
                                                                         work);
    nm->launches += 2;
    CK(cudaGetLastError());
    AStarGArgs ga{};
    ga.sG = a.sG; ga.sPt = a.sPt; ga.eG = a.eG; ga.ePt = a.ePt;
    ga.work = work; ga.workCount = cnt + 4;
    ga.counter = cnt + 5;
    ga.overflow = static_cast<uint32_t*>(nm->lists.p);
    ga.overflowCount = cnt + 6;
    ga.astat = static_cast<uint32_t*>(nm->fpStat.p);
    ga.fullLen = static_cast<int32_t*>(nm->fpLen.p);
    ga.corrVia = static_cast<uint32_t*>(nm->fpCorr.p);
    ga.scratch = static_cast<char*>(nm->wsFpG.p);
    ga.startDiv = startDiv;
    ga.fastFail = a.fastFail;
    ga.allCorridors = (out_corridor || out_ncorridor) ? 1 : 0;
    ga.workCtr = workCtr;
    ga.fault = nm->faultDev;
    if (nm->fpG == 1) {
      LaneScratch sc{};
      if ((rc = laneScratch(nm, st, &sc))) return rc;
      int64_t lanes = 32;  // queries per warp
      if (nm->laneSpread) lanes = std::max<int64_t>(1, std::min<int64_t>(32, (cn + nm->blocksFpLane - 1) / nm->blocksFpLane));
      ga.laneLimit = static_cast<int>(lanes);
      const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(nm->blocksFpLane, (cn + lanes - 1) / lanes));
      size_t smLane = 0;
      const void* fn = laneKernel(nm->laneCfg, &smLane);
      void* kargs[] = {&nm->view, &ga, &sc};
      CK(cudaLaunchKernel(fn, dim3(blocks), dim3(32), kargs, smLane, st));
      nm->launches++;
      CK(cudaGetLastError());
    } else {
      const int qpw = 32 / nm->fpG;
      const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(nm->blocksFpG, (cn + qpw - 1) / qpw));
      const size_t sm = qpw * gGroupSharedBytes<kOpenS>();
      switch (nm->fpG) {
        case 4: k_astar_g<4, kOpenS><<<blocks, 32, sm, st>>>(nm->view, ga); break;
        case 16: k_astar_g<16, kOpenS><<<blocks, 32, sm, st>>>(nm->view, ga); break;
        case 32: k_astar_g<32, kOpenS><<<blocks, 32, sm, st>>>(nm->view, ga); break;
        default: k_astar_g<8, kOpenS><<<blocks, 32, sm, st>>>(nm->view, ga); break;
      }
      nm->launches++;
      CK(cudaGetLastError());
    }
    FpFunnelArgs fa{};
    fa.starts = a.starts; fa.ends = a.ends;
    fa.sG = a.sG; fa.sPt = a.sPt; fa.eG = a.eG; fa.ePt = a.ePt;
    fa.cls = cls; fa.astat = ga.astat; fa.fullLen = ga.fullLen; fa.corrVia = ga.corrVia;
    fa.n = cn; fa.startDiv = startDiv;
    fa.out_dist = a.out_dist; fa.out_npts = a.out_npts; fa.out_pts = a.out_pts; fa.max_pts = max_pts;
    fa.out_corridor = a.out_corridor; fa.out_ncorridor = a.out_ncorridor; fa.out_status = a.out_status;
    fa.workCtr = workCtr;
    fa.fillSkipped = (flags & kFpFillSkipped) ? 1 : 0;
    k_fp_funnel<<<static_cast<unsigned>((cn + 127) / 128), 128, 0, st>>>(nm->view, fa);
    nm->launches++;
    CK(cudaGetLastError());
    // queries whose open list outgrew kOpenS: one query per warp with a 2048-entry heap
    a.work = ga.overflow;
    a.workCount = ga.overflowCount;
    if ((rc = warpTierLaunch(nm, a, false, cnt, st))) return rc;
  }
  if (nm->profile) {
    CK(cudaEventRecord(pe.e[2], st));
    nm->phaseEvents.push_back(pe);
  }
  return HBN_OK;
}


extern "C" int hbn_find_path_dev(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                                 float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                                 uint32_t* out_corridor, int32_t* out_ncorridor, uint32_t* out_status,
                                 int flags, void* stream) {
  return findPathLaunch(nm, starts, ends, n, 1, out_dist, out_npts, out_pts, max_pts, out_corridor,
                        out_ncorridor, out_status, flags, stream);
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
  int rc;
  const bool wantPts = out_npts || out_pts;
  if ((rc = nm->mgDist.ensure(pairs * 4)) || (rc = nm->mgBounds.ensure(pairs * 4)) ||
      (rc = nm->mgOrder.ensure(pairs * 4)) || (wantPts && (rc = nm->mgEnd.ensure(n * 12))))
    return rc;
  if (g > kMultiGoalFirst && nm->fpG) {
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
                             nullptr, nullptr, nullptr, kFpFillSkipped, stream, mask)))
      return rc;
    k_multigoal_round2<<<sb, 128, 0, st>>>(static_cast<uint32_t*>(nm->sG.p), static_cast<uint32_t*>(nm->eG.p),
                                           static_cast<float*>(nm->mgDist.p), n, g,
                                           static_cast<float*>(nm->mgBounds.p),
                                           static_cast<int32_t*>(nm->mgOrder.p), mask);
    nm->launches++;
    CK(cudaGetLastError());
    if ((rc = findPathLaunch(nm, starts, ends, pairs, g, static_cast<float*>(nm->mgDist.p), nullptr, nullptr, 0,
                             nullptr, nullptr, nullptr, kFpReuseSnaps, stream, mask)))
      return rc;
  } else if ((rc = findPathLaunch(nm, starts, ends, pairs, g, static_cast<float*>(nm->mgDist.p), nullptr, nullptr,
                                  0, nullptr, nullptr, nullptr, 0, stream))) {
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
                             nullptr, 0, stream)))
      return rc;
  }
  return HBN_OK;
}

int hbn_try_step_dev(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                     int allow_sliding, float* out_pts, void* stream) {
  if (!nm || (n > 0 && (!starts || !ends || !out_pts))) return fail(HBN_ERR_INVALID, "null argument");
  if (n <= 0) return HBN_OK;
  DeviceGuard g(nm->device);
  std::lock_guard<std::recursive_mutex> lk(nm->mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = nm->sG.ensure(n * 4)) || (rc = nm->eG.ensure(n * 4)) || (rc = nm->e2G.ensure(n * 4)) ||
      (rc = nm->sPt.ensure(n * 12)) || (rc = nm->epPt.ensure(n * 12)) || (rc = nm->lastPoly.ensure(n * 4)))
    return rc;
  uint32_t* sG = static_cast<uint32_t*>(nm->sG.p);
  uint32_t* eG = static_cast<uint32_t*>(nm->eG.p);
  uint32_t* e2G = static_cast<uint32_t*>(nm->e2G.p);
  float* sPt = static_cast<float*>(nm->sPt.p);
  float* ep = static_cast<float*>(nm->epPt.p);
  uint32_t* last = static_cast<uint32_t*>(nm->lastPoly.p);
  if (snapDualLaunch(nm, starts, n, sPt, sG, ends, n, nullptr, eG, st, &rc)) {
    if (rc) return rc;
  } else {
    if ((rc = snapLaunch(nm, starts, nullptr, n, sPt, sG, nullptr, nullptr, nullptr, 0.f, st))) return rc;
    if ((rc = snapLaunch(nm, ends, nullptr, n, nullptr, eG, nullptr, nullptr, nullptr, 0.f, st))) return rc;
  }
  k_trystep_a<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(nm->view, ends, sG, sPt, eG, n,
                                                                      allow_sliding, ep, last);
  nm->launches++;
  CK(cudaGetLastError());
  if ((rc = snapLaunch(nm, ep, nullptr, n, nullptr, e2G, nullptr, nullptr, nullptr, 0.f, st))) return rc;
  k_trystep_b<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(nm->view, starts, sG, e2G, last, ep, n, out_pts);
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
  int rc;
  if ((rc = nm->sG.ensure(n * 4)) || (rc = nm->sPt.ensure(n * 12)) || (rc = nm->lists.ensure(n * 4))) return rc;
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
  int64_t blocks = std::min<int64_t>(nm->blocksWallS, (n + kFpWarps - 1) / kFpWarps);
  k_wall<kWallCapS, kWsShared><<<static_cast<unsigned>(blocks), threads,
                                kFpWarps * wsSharedBytes(kWallCapS, kWsShared), st>>>(nm->view, a);
  nm->launches++;
  CK(cudaGetLastError());
  WallArgs b = a;
  b.work = a.overflow;
  b.workCount = a.overflowCount;
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
    }
    return HBN_OK;
  }
  template <class T> T* dev(size_t o) { return reinterpret_cast<T*>(static_cast<char*>(nm->io.p) + o); }
  int h2d() {
    for (auto& it : items)
      if (it.src) {
        memcpy(static_cast<char*>(nm->pinned) + it.off, it.src, it.bytes);
        CK(cudaMemcpyAsync(static_cast<char*>(nm->io.p) + it.off, static_cast<char*>(nm->pinned) + it.off,
                           it.bytes, cudaMemcpyHostToDevice, nm->stream));
      }
    return HBN_OK;
  }
  int d2h() {
    for (auto& it : items)
      if (it.dst)
        CK(cudaMemcpyAsync(static_cast<char*>(nm->pinned) + it.off, static_cast<char*>(nm->io.p) + it.off,
                           it.bytes, cudaMemcpyDeviceToHost, nm->stream));
    CK(cudaStreamSynchronize(nm->stream));
    for (auto& it : items)
      if (it.dst) memcpy(it.dst, static_cast<char*>(nm->pinned) + it.off, it.bytes);
    return checkFault(nm);
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

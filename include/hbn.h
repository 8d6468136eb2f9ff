/* hbn.h -- C ABI of the B200-native batched navmesh query library (libhbn.so).
 *
 * Drop-in boundary for habitat-sim's navmesh *query* path: everything
 * esp::nav::PathFinder::Impl does through dtNavMeshQuery
 * (src/esp/nav/PathFinder.cpp, "PF.cpp"; Detour call sites PF.cpp:139, 1260, 1448, 1456,
 * 1597, 1687, 1808) runs as CUDA kernels against a device-resident copy of the navmesh.
 * The navmesh itself is still built / loaded on the host by the reference
 * (PathFinder::build -> Recast, PF.cpp:612-930; loadNavMesh PF.cpp:1091-1175).
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns an hbn_status (0 = ok) and never
 *    throws; hbn_last_error() gives the message of the calling thread's last failure.
 *  - `*_dev` entry points take DEVICE pointers and a cudaStream_t (as void*); they enqueue
 *    kernels and never synchronise.  The handle's scratch grows on the first call of a size
 *    (cudaMalloc / cudaFree); after hbn_navmesh_reserve(nm, n) calls of up to n queries allocate
 *    nothing and can be captured in CUDA graphs.  Calls on one handle from different streams
 *    are ordered by the handle (an event per call), not run concurrently.
 *  - entry points without the suffix take HOST pointers, do the H2D/D2H copies on the
 *    handle's own stream and return after the results are in the caller's buffers.
 *  - points are float32 xyz triples, y up; poly refs are 32-bit dtPolyRef values
 *    (salt|tile|poly, DetourNavMesh.h:529-562); island ids are IslandSystem ids
 *    (PF.cpp:167-207); per-element failure sentinels are the reference's: NaN point, ref 0,
 *    island -1, +inf distance, 0 points, unchanged start.
 *  - there is no CPU fallback: without a CUDA device every create call fails.
 */
#ifndef HBN_H_
#define HBN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hbn_navmesh* hbn_navmesh_t;

typedef enum {
  HBN_OK = 0,
  HBN_ERR_INVALID = 1,   /* bad argument */
  HBN_ERR_FORMAT = 2,    /* navmesh image / tile blob rejected */
  HBN_ERR_CUDA = 3,      /* CUDA runtime error (message in hbn_last_error) */
  HBN_ERR_NO_DEVICE = 4, /* no usable CUDA device: the library has no CPU path */
  HBN_ERR_NO_AREA = 5,   /* island has no navigable area (reference throws, PF.cpp:1240) */
  HBN_ERR_LIMIT = 6      /* navmesh exceeds a packed-index limit of the device layout */
} hbn_status;

const char* hbn_last_error(void);
int hbn_device_count(void);

/* ---- navmesh lifetime ------------------------------------------------------------- */

/* Replaces PathFinder::Impl::loadNavMesh + initNavQuery (PF.cpp:1091-1175, 932-950) for an
 * in-memory MSET v1/v2 image (the bytes of a habitat `.navmesh` file): tiles are added,
 * links connected, islands labelled, zero-area polys disabled, then the mesh is flattened
 * and uploaded to `device`. */
int hbn_navmesh_create_from_mset(const void* bytes, size_t len, int device, hbn_navmesh_t* out);

/* Replaces the hand-over a live PathFinder would do right after initNavQuery
 * (PF.cpp:932-950): the FINALISED tile blobs of its dtNavMesh (dtMeshTile::data with links
 * connected and poly flags final; dtNavMesh::getTile / getTileRef, DetourNavMesh.h:420-470).
 * params5 = {orig[3], tileWidth, tileHeight}; poly_islands (nullable) = island id per poly
 * in (tile table order, poly order), e.g. from IslandSystem::getPolyIsland (PF.cpp:398);
 * when null the islands are recomputed from the given flags (trap T5 in SURVEY.md).
 * island_radii (nullable, n_islands entries; used with poly_islands) = IslandSystem::islandRadius(i)
 * (PF.cpp:236-239) of the same PathFinder: the reference sums an island's vertices in the order of
 * its own flood fill over the flags of that moment, which the finalised tiles no longer tell; without
 * them the radii are recomputed per island id (equal up to f32 summation order). */
typedef struct {
  const void* data;
  int32_t size;
  uint32_t tile_ref; /* dtNavMesh::getTileRef(tile) */
} hbn_tile_blob;
int hbn_navmesh_create_from_tiles(const hbn_tile_blob* tiles, int n_tiles, const float* params5,
                                  int max_tiles, int max_polys, const int32_t* poly_islands,
                                  const float* island_radii, int n_islands, int device, hbn_navmesh_t* out);

void hbn_navmesh_destroy(hbn_navmesh_t nm);

typedef struct {
  int32_t device;
  int32_t num_tiles, num_polys, num_links, num_bv_nodes, num_islands;
  int32_t poly_bits, tile_bits, salt_bits;
  int32_t has_settings;   /* MSET v2 NavMeshSettings block present */
  float bounds_min[3];    /* PathFinder::bounds(), PF.cpp:1160-1170 */
  float bounds_max[3];
  float navigable_area;   /* getNavigableArea(ID_UNDEFINED), PF.cpp:1079-1084 */
  int64_t device_bytes;   /* size of the device-resident navmesh */
} hbn_navmesh_info;
int hbn_navmesh_get_info(hbn_navmesh_t nm, hbn_navmesh_info* out);
/* islandRadius(idx) PF.cpp:1773-1775 and getNavigableArea(idx) PF.cpp:1900 */
int hbn_navmesh_island_info(hbn_navmesh_t nm, int island, float* radius, float* area);
/* raw 56-byte NavMeshSettings block (PF.h:137-299) of the MSET image */
int hbn_navmesh_get_settings(hbn_navmesh_t nm, void* out56);
/* PathFinder::bounds() for a handle made from live tiles: a navmesh the reference has just BUILT reports
 * the bounds of the input geometry (PF.cpp:619-621, 927), a loaded one the union of the tile bounds
 * (PF.cpp:1160-1170) -- which is what the handle computes by itself.  bounds6 = {min xyz, max xyz}. */
int hbn_navmesh_set_bounds(hbn_navmesh_t nm, const float* bounds6);
/* NavMeshSettings for a handle made from live tiles (PathFinder::Impl::build keeps them,
 * PF.cpp:926; saveNavMesh refuses to write without them, PF.cpp:1199-1203) */
int hbn_navmesh_set_settings(hbn_navmesh_t nm, const void* in56);
/* PathFinder::Impl::saveNavMesh, PF.cpp:1177-1223: the MSET v2 image of the navmesh as the handle
 * holds it (links connected, zero-area polys disabled), whether it came from an image or from live
 * tiles.  Two-call pattern: returns the size in bytes (-1: no settings known, see hbn_last_error);
 * fills `out` when cap is large enough.  Host function. */
int64_t hbn_navmesh_save_mset(hbn_navmesh_t nm, void* out, int64_t cap);
/* Tuning knobs of a handle (defaults come from the environment once, at creation):
 * "lane_scratch_bytes" (cap on the per-lane search state of find_path in HBM; 0 = half of the free
 * memory; the grid shrinks under the cap), "lane_cfg", "blocks_per_sm", "lane_spread", "snap_spread",
 * "snap_dual", "snap_group", "snap_cap", "nvtx".  Unknown keys fail with HBN_ERR_INVALID. */
int hbn_navmesh_set_option(hbn_navmesh_t nm, const char* key, int64_t value);
/* Sizes every scratch buffer for batches of up to n queries (find_path / try_step / env step pairs,
 * snap / obstacle points), so that later *_dev calls of that size only enqueue work: no cudaMalloc /
 * cudaFree, hence no implicit device synchronisation, and they can be captured into CUDA graphs.
 * Without it the buffers grow on the first call of each size. */
int hbn_navmesh_reserve(hbn_navmesh_t nm, int64_t n);
/* device bytes of scratch the handle currently holds (besides the navmesh itself) */
int64_t hbn_navmesh_scratch_bytes(hbn_navmesh_t nm);
/* kernels launched through this handle so far (bench.py's gpu_launches evidence) */
int64_t hbn_navmesh_launch_count(hbn_navmesh_t nm);
/* Phase timing of hbn_find_path_dev for bench.py's roofline: while enabled, every call
 * records CUDA events on ITS stream around the two projectToPoly launches and around the
 * A* + funnel launches.  hbn_navmesh_phase_times waits for those events and returns the
 * accumulated milliseconds {snap, path} and the number of calls, then resets. */
int hbn_navmesh_set_profiling(hbn_navmesh_t nm, int enable);
int hbn_navmesh_phase_times(hbn_navmesh_t nm, double* out_ms2, int64_t* out_calls);
/* Work counters accumulated by HBN_FP_COUNT_WORK calls: out8 = {expanded polys, links of the
 * expanded polys, their non-null neighbours, corridor polys, links of the corridor polys,
 * straight-path points, queries that ran A*, queries}.  Synchronises the device. */
int hbn_navmesh_work_counters(hbn_navmesh_t nm, uint64_t* out8, int reset);
/* navmesh triangles for build_navmesh_vertices/indices (getNavMeshData, PF.cpp:1898-1968):
 * detail triangles of every poly of `island` (-1 = all polys), in tile-table / poly order, 9 floats each.
 * Two-call pattern: returns the triangle count; fills `out` when cap_tris is large enough. */
int64_t hbn_navmesh_triangles(hbn_navmesh_t nm, int island, float* out, int64_t cap_tris);

/* ---- batched queries, device pointers --------------------------------------------- */

/* snapPoint / getIsland (PF.cpp:1725-1771) = projectToPoly (PF.cpp:126-147) =
 * dtNavMeshQuery::findNearestPoly with half extents {2,4,2}.  islands (nullable): per-point
 * island restriction, -1 = none.  All outputs nullable. */
int hbn_snap_point_dev(hbn_navmesh_t nm, const float* pts, const int32_t* islands, int64_t n,
                       float* out_pts, uint32_t* out_refs, int32_t* out_islands, void* stream);

/* isNavigable, PF.cpp:1814-1831 */
int hbn_is_navigable_dev(hbn_navmesh_t nm, const float* pts, int64_t n, float max_y_delta,
                         uint8_t* out, void* stream);

/* findPath(ShortestPath&), PF.cpp:1414-1468.  out_dist[n] = geodesicDistance (+inf: none).
 * Optional: out_npts[n]; out_pts[n, max_pts, 3] (first min(npts, max_pts) points);
 * out_corridor[n, 256] + out_ncorridor[n] (the dtNavMeshQuery::findPath poly corridor);
 * out_status[n, 2] = Detour status words of findPath / findStraightPath.
 * flags: HBN_FP_EXACT_STATUS keeps searching after the node pool is exhausted, like
 * DetourNavMeshQuery.cpp:1074-1078, so that corridors / status words of FAILED queries match
 * Detour too; by default such a query stops there (its PathFinder result, "no path",
 * is already decided: PF.cpp:1450).  HBN_FP_COUNT_WORK adds this call's work counters
 * (expanded polys, links, neighbours, ... see hbn_navmesh_work_counters) to the handle's
 * totals; bench.py derives the roofline's algorithmic bytes from them in an untimed pass. */
enum { HBN_FP_EXACT_STATUS = 1, HBN_FP_COUNT_WORK = 2 };
int hbn_find_path_dev(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                      float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                      uint32_t* out_corridor, int32_t* out_ncorridor, uint32_t* out_status,
                      int flags, void* stream);

/* findPath(MultiGoalShortestPath&), PF.cpp:1515-1572, fresh path objects (no cache):
 * ends[n, g, 3].  out_index[n] = closestEndPointIndex (-1: none). */
int hbn_find_path_multigoal_dev(hbn_navmesh_t nm, const float* starts, const float* ends,
                                int64_t n, int g, float* out_dist, int32_t* out_index,
                                int32_t* out_npts, float* out_pts, int max_pts, void* stream);

/* tryStep / tryStepNoSliding, PF.cpp:1575-1722 */
int hbn_try_step_dev(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                     int allow_sliding, float* out_pts, void* stream);

/* One environment step of habitat-lab's PointNav loop: tryStep(starts[i], targets[i]) (PF.cpp:1575-1722,
 * simulator.py:660-673) then findPath(new position, goals[i]) for the geodesic reward
 * (PF.cpp:1414-1468).  Results are those of hbn_try_step_dev followed by hbn_find_path_dev, bit for
 * bit; the projections are shared between the two and the whole step is 9 kernels. */
int hbn_env_step_dev(hbn_navmesh_t nm, const float* starts, const float* targets, const float* goals,
                     int64_t n, int allow_sliding, float* out_pos, float* out_dist, void* stream);

/* GreedyGeodesicFollowerImpl::nextBestPrimAlong (GreedyFollower.cpp:83-140) for n agents at once: the
 * kinematics of every primitive [LEFT]*k+[FORWARD] / [RIGHT]*k+[FORWARD] (default_controls.py: rotate
 * about +Y by turn_amount, move forward_amount along local -Z), their try_step + geodesic distance +
 * obstacle distance (GreedyFollower.cpp:46-81), computeReward and the selection all run on the device.
 * rots: [n, 4] float64 quaternions (x, y, z, w); poss: [n, 3] float64; goals: [n, 3] float32.
 * out_prim[i]: -2 = ERROR (no path), -1 = STOP (within goal_dist), -3 = no acceptable primitive,
 * else (k << 1) | side: k turns (side 0 = LEFT, 1 = RIGHT) followed by one FORWARD.
 * out_geo (nullable): the geodesic distance from the current position. */
typedef struct {
  float goal_dist, forward_amount;
  double sin_half_turn, cos_half_turn; /* sin / cos of turn_amount / 2 */
  int32_t n_steps;                      /* headings per side: the reference's loop count, angle += turn in f32 while < pi */
  int32_t allow_sliding;
} hbn_follower_params;
int hbn_follower_best_prims_dev(hbn_navmesh_t nm, const double* rots, const double* poss, const float* goals,
                                int64_t n, const hbn_follower_params* params, int32_t* out_prim,
                                float* out_geo, void* stream);

/* closestObstacleSurfacePoint / distanceToClosestObstacle, PF.cpp:1788-1812.
 * out_hit_pos / out_hit_normal nullable. */
int hbn_closest_obstacle_dev(hbn_navmesh_t nm, const float* pts, int64_t n, float max_radius,
                             float* out_hit_pos, float* out_hit_normal, float* out_hit_dist,
                             void* stream);

/* getRandomNavigablePoint, PF.cpp:1236-1281, n independent samples.  Sample i draws its
 * uniforms from the counter-based stream hbn_uniform(seed, query0 + i, draw) in exactly the
 * order the reference consumes rand() (SURVEY.md trap T7).  islands nullable. */
int hbn_random_points_dev(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n,
                          const int32_t* islands, int max_tries, float* out_pts,
                          uint32_t* out_refs, void* stream);
/* getRandomNavigablePointInCircle, PF.cpp:1283-1332 (Python get_random_navigable_point_near, SPB.cpp:179-183):
 * sample i is drawn with the filter of setPolyFlagForIslandCircle(centers[i], radius, islands[i])
 * (PF.cpp:330-393) and accepted when it lies within `radius` of its centre in xz; NaN after
 * max_tries.  Same counter-based stream as hbn_random_points. */
int hbn_random_points_near_dev(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n,
                               const float* centers, float radius, const int32_t* islands, int max_tries,
                               float* out_pts, void* stream);
/* The stream: u = float(r) / float(RAND_MAX) with r = the top 31 bits of a 32-bit hash of
 * (seed, query, draw) -- the form of the reference's frand() (PF.cpp:1232-1234), so the
 * reference's own code driven by rand() := r sees bit-identical uniforms. */
float hbn_uniform(uint64_t seed, uint64_t query, uint32_t draw);
/* order[0..n) <- the permutation std::sort (libstdc++ introsort, unstable) leaves when sorting 0..n-1
 * by key: the goal order of findPath(MultiGoalShortestPath&), PF.cpp:1542-1548.  Host function. */
void hbn_std_sort_order(const float* key, int n, int32_t* order);

/* ---- the same queries with HOST buffers (copies + synchronisation inside) ---------- */
int hbn_snap_point(hbn_navmesh_t nm, const float* pts, const int32_t* islands, int64_t n,
                   float* out_pts, uint32_t* out_refs, int32_t* out_islands);
int hbn_is_navigable(hbn_navmesh_t nm, const float* pts, int64_t n, float max_y_delta,
                     uint8_t* out);
int hbn_find_path(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                  float* out_dist, int32_t* out_npts, float* out_pts, int max_pts,
                  uint32_t* out_corridor, int32_t* out_ncorridor, uint32_t* out_status,
                  int flags);
int hbn_find_path_multigoal(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                            int g, float* out_dist, int32_t* out_index, int32_t* out_npts,
                            float* out_pts, int max_pts);
int hbn_try_step(hbn_navmesh_t nm, const float* starts, const float* ends, int64_t n,
                 int allow_sliding, float* out_pts);
/* the env step with host buffers: copies + kernels are captured once per (n, allow_sliding) into a
 * CUDA graph and replayed */
int hbn_env_step(hbn_navmesh_t nm, const float* starts, const float* targets, const float* goals,
                 int64_t n, int allow_sliding, float* out_pos, float* out_dist);
int hbn_follower_best_prims(hbn_navmesh_t nm, const double* rots, const double* poss, const float* goals,
                            int64_t n, const hbn_follower_params* params, int32_t* out_prim, float* out_geo);
int hbn_closest_obstacle(hbn_navmesh_t nm, const float* pts, int64_t n, float max_radius,
                         float* out_hit_pos, float* out_hit_normal, float* out_hit_dist);
int hbn_random_points(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n,
                      const int32_t* islands, int max_tries, float* out_pts, uint32_t* out_refs);
int hbn_random_points_near(hbn_navmesh_t nm, uint64_t seed, uint64_t query0, int64_t n, const float* centers,
                           float radius, const int32_t* islands, int max_tries, float* out_pts);

#ifdef __cplusplus
}
#endif
#endif /* HBN_H_ */

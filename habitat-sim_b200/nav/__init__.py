"""`habitat_sim.nav` twin: same class / method / property names, defaults and failure
conventions as src/esp/bindings/ShortestPathBindings.cpp ("SPB.cpp"), every query executed
by the CUDA kernels of libhbn.so.  Scalar methods take one 3-vector (numpy array / sequence)
like the reference; the *batched* methods (plural names) take `[N, 3]` numpy arrays (host
path, copies inside) or torch CUDA tensors (device path, no copies, current stream).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import struct

import numpy as np

from .. import _lib
from .._lib import HbnError, check

from .sharding import gather as gather_shards, shard_slice, shard_slices  # noqa: E402,F401

__all__ = ["shard_slices", "shard_slice", "gather_shards", "MultiGpuPathFinder", "PathFinder", "ShortestPath", "MultiGoalShortestPath", "HitRecord", "NavMeshSettings",
           "GreedyFollowerCodes", "GreedyGeodesicFollowerImpl", "GreedyGeodesicFollower",
           "GreedyGeodesicFollowerBatch", "GreedyGeodesicFollowerBatchImpl", "HbnError",
           "multigoal_find_path", "std_sort_order"]

MAX_PATH_POINTS = 256  # MAX_POLYS, PathFinder.cpp:1443


def _vec3(p) -> np.ndarray:
    a = np.asarray(p, dtype=np.float32).reshape(-1)
    if a.size != 3:
        raise TypeError("expected a 3-vector")
    return np.ascontiguousarray(a)


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class HitRecord:
    """esp::nav::HitRecord, PathFinder.h:33-41 / SPB.cpp:30-39"""

    def __init__(self, hit_pos=None, hit_normal=None, hit_dist=float("inf")):
        self.hit_pos = np.zeros(3, np.float32) if hit_pos is None else np.asarray(hit_pos, np.float32)
        self.hit_normal = np.zeros(3, np.float32) if hit_normal is None else np.asarray(hit_normal, np.float32)
        self.hit_dist = float(hit_dist)


class ShortestPath:
    """esp::nav::ShortestPath, PathFinder.h:47-75 / SPB.cpp:41-51"""

    def __init__(self):
        self.requested_start = np.zeros(3, np.float32)
        self.requested_end = np.zeros(3, np.float32)
        self.points: list = []
        self.geodesic_distance = float("inf")


class MultiGoalShortestPath:
    """esp::nav::MultiGoalShortestPath, PathFinder.h:81-123 / SPB.cpp:53-74.

    The object keeps the reference's private state (MultiGoalShortestPath::Impl, PathFinder.cpp:95-123)
    so that a REUSED object behaves like the reference's: the goals are projected once per
    `requested_ends` assignment, the per-goal lower bounds carry over from call to call
    (PathFinder.cpp:1521-1539), and the validity flags are only ever appended (trap T4: after assigning
    new ends the flags of the old ones are read)."""

    def __init__(self):
        self.requested_start = np.zeros(3, np.float32)
        self._requested_ends = np.zeros((0, 3), np.float32)
        self.points: list = []
        self.geodesic_distance = float("inf")
        self.closest_end_point_index = -1
        # Impl
        self._ends_projected = False                 # "endRefs not empty"
        self._end_is_valid: list = []                # append only, never cleared (PathFinder.cpp:1497-1500)
        self._min_theoretical_dist = np.zeros(0, np.float32)
        self._prev_requested_start = np.zeros(3, np.float32)

    @property
    def requested_ends(self):
        return [e.copy() for e in self._requested_ends]

    @requested_ends.setter
    def requested_ends(self, ends):  # setRequestedEnds, PathFinder.cpp:111-118
        self._requested_ends = np.ascontiguousarray(np.asarray(ends, dtype=np.float32).reshape(-1, 3))
        self._ends_projected = False
        self._min_theoretical_dist = np.zeros(len(self._requested_ends), np.float32)


def _mn_length(d):
    """Magnum Vector3::length() in float32: sqrt(((0 + x^2) + y^2) + z^2), rows of d."""
    d = np.asarray(d, np.float32)
    acc = d[..., 0] * d[..., 0]
    acc = (acc + d[..., 1] * d[..., 1]).astype(np.float32)
    acc = (acc + d[..., 2] * d[..., 2]).astype(np.float32)
    return np.sqrt(acc).astype(np.float32)


def multigoal_find_path(path: MultiGoalShortestPath, snap_refs, find_paths, sort_order) -> bool:
    """PathFinder::Impl::findPath(MultiGoalShortestPath&) (PathFinder.cpp:1470-1572) over batched
    primitives: snap_refs(pts [N,3]) -> poly refs [N] (0 = not projectable); find_paths(starts, ends,
    max_points) -> dict(geodesic_distance, num_points, points); sort_order(keys) -> the goal order
    std::sort leaves (unstable, PathFinder.cpp:1544-1548).  All goal searches of a call go through ONE
    find_paths call; the reference's sequential goal loop is then replayed over the results."""
    path.geodesic_distance = float("inf")
    path.closest_end_point_index = -1
    path.points = []
    start = _vec3(path.requested_start)
    ends = path._requested_ends
    if int(snap_refs(start[None])[0]) == 0:  # findPathSetup: the start does not project
        return False
    if not path._ends_projected:
        if len(ends):
            valid = np.asarray(snap_refs(ends)) != 0
            path._end_is_valid.extend(bool(v) for v in valid)
            path._ends_projected = True
            if not valid.any():
                return False
        # with no ends at all nothing is cached and the loop below has nothing to visit
    g = len(ends)
    if g > 1:
        moved = np.float32(find_paths(start[None], path._prev_requested_start[None], 0)["geodesic_distance"][0])
        l2 = _mn_length(ends - start[None])
        with np.errstate(invalid="ignore"):
            a = (path._min_theoretical_dist - moved).astype(np.float32)
            path._min_theoretical_dist = np.where(a < l2, l2, a)  # std::max(a, l2), PF.cpp:1533: a NaN l2 keeps a
        path._prev_requested_start = start.copy()
    if g == 0:
        return False
    order = sort_order(np.ascontiguousarray(path._min_theoretical_dist, np.float32))
    res = find_paths(np.repeat(start[None], g, 0), ends, MAX_PATH_POINTS)
    dist = np.asarray(res["geodesic_distance"], np.float32)
    best = np.float32(np.inf)
    for i in order:
        i = int(i)
        if not path._end_is_valid[i]:
            continue
        if path._min_theoretical_dist[i] > best:
            continue
        d = dist[i]
        if d < np.inf and d < best:
            path._min_theoretical_dist[i] = d
            best = d
            n = int(res["num_points"][i])
            path.points = [np.array(res["points"][i, k], np.float32) for k in range(n)]
            path.closest_end_point_index = i
    path.geodesic_distance = float(best)
    return bool(best < np.inf)


class NavMeshSettings:
    """esp::nav::NavMeshSettings, PathFinder.h:137-299 / SPB.cpp:76-140 (56-byte block)."""

    _FMT = "<13f4?"
    _FIELDS = ["cell_size", "cell_height", "agent_height", "agent_radius", "agent_max_climb",
               "agent_max_slope", "region_min_size", "region_merge_size", "edge_max_len",
               "edge_max_error", "verts_per_poly", "detail_sample_dist", "detail_sample_max_error",
               "filter_low_hanging_obstacles", "filter_ledge_spans",
               "filter_walkable_low_height_spans", "include_static_objects"]

    def __init__(self):
        self.set_defaults()

    def set_defaults(self):
        vals = [0.05, 0.2, 1.5, 0.1, 0.2, 45.0, 20.0, 20.0, 12.0, 1.3, 6.0, 6.0, 1.0, True, True, True,
                False]
        for k, v in zip(self._FIELDS, vals):
            setattr(self, k, v)

    def to_bytes(self) -> bytes:
        return struct.pack(self._FMT, *[getattr(self, k) for k in self._FIELDS])

    @classmethod
    def from_bytes(cls, b: bytes) -> "NavMeshSettings":
        s = cls()
        for k, v in zip(cls._FIELDS, struct.unpack(cls._FMT, b[:56])):
            setattr(s, k, v)
        return s

    # PathFinder.cpp:68-93 + io/JsonEspTypes.cpp:287-331: the 16 serialised fields and their keys
    # (include_static_objects is not part of the JSON form)
    _JSON_KEYS = ["cellSize", "cellHeight", "agentHeight", "agentRadius", "agentMaxClimb", "agentMaxSlope",
                  "regionMinSize", "regionMergeSize", "edgeMaxLen", "edgeMaxError", "vertsPerPoly",
                  "detailSampleDist", "detailSampleMaxError", "filterLowHangingObstacles", "filterLedgeSpans",
                  "filterWalkableLowHeightSpans"]

    def read_from_json(self, json_file: str) -> None:
        """NavMeshSettings::readFromJSON: members present in the file overwrite the fields; a missing
        file or a parse error is logged, not raised (PathFinder.cpp:68-81)."""
        import json
        import logging
        if not os.path.exists(json_file):
            logging.getLogger("habitat_sim_b200.nav").error("File %s not found.", json_file)
            return
        try:
            with open(json_file) as f:
                doc = json.load(f)
            for key, field in zip(self._JSON_KEYS, self._FIELDS):
                if key in doc:
                    cur = getattr(self, field)
                    setattr(self, field, bool(doc[key]) if isinstance(cur, bool) else float(doc[key]))
        except Exception:  # noqa: BLE001 (the reference catches everything here)
            logging.getLogger("habitat_sim_b200.nav").error("Failed to parse %s.", json_file)

    def write_to_json(self, json_file: str) -> None:
        """NavMeshSettings::writeToJSON (PathFinder.cpp:83-93): 7 decimal places at most."""
        import json
        doc = {}
        for key, field in zip(self._JSON_KEYS, self._FIELDS):
            v = getattr(self, field)
            doc[key] = bool(v) if isinstance(v, bool) else round(float(np.float32(v)), 7)
        with open(json_file, "w") as f:
            json.dump(doc, f, indent=4)

    def __eq__(self, o):
        if not isinstance(o, NavMeshSettings):
            return NotImplemented
        for k in self._FIELDS[:13]:
            if not abs(getattr(self, k) - getattr(o, k)) < 1e-5:
                return False
        return all(getattr(self, k) == getattr(o, k) for k in self._FIELDS[13:])


def std_sort_order(keys) -> np.ndarray:
    """argsort of float32 keys exactly as libstdc++'s std::sort leaves it (introsort, unstable): the
    goal order of PathFinder.cpp:1542-1548.  Host code of libhbn (hbn_std_sort_order)."""
    k = np.ascontiguousarray(keys, np.float32)
    order = np.empty(len(k), np.int32)
    _lib.lib().hbn_std_sort_order(k.ctypes.data, len(k), order.ctypes.data)
    return order


class PathFinder:
    """esp::nav::PathFinder (PathFinder.h:336-677) as bound by SPB.cpp:142-273."""

    def __init__(self, device: int = 0):
        self._device = int(device)
        self._h = None
        self._info = None
        self._image: bytes | None = None
        self._seed = 0
        self._rand_counter = 0

    # ---- lifetime ---------------------------------------------------------------------
    def __del__(self):
        self._release()

    def _release(self):
        try:
            if self._h:
                _lib.lib().hbn_navmesh_destroy(self._h)
        except Exception:
            pass
        self._h = None

    def _adopt(self, handle):
        self._release()
        self._h = handle
        info = _lib.NavMeshInfo()
        check(_lib.lib().hbn_navmesh_get_info(self._h, C.byref(info)))
        self._info = info

    def load_nav_mesh(self, path: str) -> bool:
        """PathFinder::loadNavMesh, PathFinder.cpp:1091-1175"""
        try:
            with open(path, "rb") as f:
                data = f.read()
        except OSError:
            return False
        return self.load_nav_mesh_bytes(data)

    def load_nav_mesh_bytes(self, data: bytes) -> bool:
        h = C.c_void_p()
        rc = _lib.lib().hbn_navmesh_create_from_mset(data, len(data), self._device, C.byref(h))
        if rc == 2:  # HBN_ERR_FORMAT: the reference returns false
            return False
        check(rc)
        self._adopt(h)
        self._image = bytes(data)
        return True

    def load_from_tiles(self, tiles, orig, tile_width, tile_height, max_tiles, max_polys,
                        poly_islands=None, island_radii=None, bounds=None) -> bool:
        """Hand-over of a live dtNavMesh's finalised tiles (hbn_navmesh_create_from_tiles).
        tiles: iterable of (tile_ref, bytes); poly_islands / island_radii: the caller's
        IslandSystem (island id per poly in tile-table / poly order; radius per island)."""
        tiles = list(tiles)
        arr = (_lib.TileBlob * len(tiles))()
        keep = []
        for i, (ref, blob) in enumerate(tiles):
            buf = C.create_string_buffer(blob, len(blob))
            keep.append(buf)
            arr[i].data = C.cast(buf, C.c_void_p)
            arr[i].size = len(blob)
            arr[i].tile_ref = ref
        p5 = (C.c_float * 5)(orig[0], orig[1], orig[2], tile_width, tile_height)
        isl = None
        if poly_islands is not None:
            isl = np.ascontiguousarray(poly_islands, dtype=np.int32)
        rad = None
        if island_radii is not None:
            rad = np.ascontiguousarray(island_radii, dtype=np.float32)
        h = C.c_void_p()
        check(_lib.lib().hbn_navmesh_create_from_tiles(arr, len(tiles), p5, int(max_tiles), int(max_polys),
                                                       isl.ctypes.data if isl is not None else None,
                                                       rad.ctypes.data if rad is not None else None,
                                                       len(rad) if rad is not None else 0,
                                                       self._device, C.byref(h)))
        self._adopt(h)
        self._image = None
        if bounds is not None:  # PathFinder::bounds() of a freshly built mesh: the input geometry's
            b6 = np.ascontiguousarray(np.concatenate([np.asarray(bounds[0], np.float32), np.asarray(bounds[1], np.float32)]))
            check(_lib.lib().hbn_navmesh_set_bounds(h, b6.ctypes.data))
            check(_lib.lib().hbn_navmesh_get_info(self._h, C.byref(self._info)))
        return True

    def save_nav_mesh_bytes(self) -> bytes:
        """The MSET v2 image PathFinder::saveNavMesh (PathFinder.cpp:1177-1223) writes for the navmesh
        this handle holds -- links connected, zero-area polys disabled -- whether it was loaded from an
        image or handed over as live tiles (hbn_navmesh_save_mset)."""
        h = self._need()
        L = _lib.lib()
        n = L.hbn_navmesh_save_mset(h, None, 0)
        if n < 0:
            raise RuntimeError((L.hbn_last_error() or b"").decode())
        buf = C.create_string_buffer(n)
        L.hbn_navmesh_save_mset(h, buf, n)
        return buf.raw

    def save_nav_mesh(self, path: str) -> bool:
        """PathFinder::saveNavMesh, PathFinder.cpp:1177-1223: false without a navmesh or without
        NavMeshSettings (a handle made from live tiles needs set_nav_mesh_settings_bytes first)."""
        if not self._h:
            return False
        try:
            data = self.save_nav_mesh_bytes()
        except RuntimeError:
            return False
        with open(path, "wb") as f:
            f.write(data)
        return True

    def set_nav_mesh_settings_bytes(self, raw56: bytes) -> None:
        """the 56-byte NavMeshSettings block (PF.h:137-299) of a navmesh handed over as live tiles"""
        assert len(raw56) == 56
        check(_lib.lib().hbn_navmesh_set_settings(self._need(), raw56))

    def build_navmesh_from_triangles(self, *a, **k):
        raise NotImplementedError(
            "navmesh construction stays on the host in the reference's Recast path "
            "(PathFinder::build, PathFinder.cpp:612-930); load its result with load_nav_mesh")

    def _need(self):
        if not self._h:
            raise RuntimeError("PathFinder has no navmesh loaded")
        return self._h

    # ---- properties (SPB.cpp:146-273) -------------------------------------------------
    @property
    def is_loaded(self) -> bool:
        return self._h is not None

    @property
    def device(self) -> int:
        return self._device

    def get_bounds(self):
        self._need()
        return (np.array(self._info.bounds_min, np.float32), np.array(self._info.bounds_max, np.float32))

    @property
    def num_islands(self) -> int:
        self._need()
        return int(self._info.num_islands)

    @property
    def navigable_area(self) -> float:
        self._need()
        return float(self._info.navigable_area)

    def island_area(self, island_index: int = -1) -> float:
        self._need()
        if island_index == -1:
            return float(self._info.navigable_area)
        a = C.c_float()
        check(_lib.lib().hbn_navmesh_island_info(self._h, int(island_index), None, C.byref(a)))
        return a.value

    def island_radius(self, pt_or_index) -> float:
        """islandRadius(pt) PathFinder.cpp:1777-1786 / islandRadius(index) :1773-1775"""
        self._need()
        if isinstance(pt_or_index, (int, np.integer)):
            r = C.c_float()
            check(_lib.lib().hbn_navmesh_island_info(self._h, int(pt_or_index), C.byref(r), None))
            return r.value
        isl = self.get_island(pt_or_index)
        if isl < 0:
            return 0.0
        return self.island_radius(int(isl))

    @property
    def nav_mesh_settings(self):
        if not self._h:  # tests/test_nav.py:95-99: None until a navmesh is loaded
            return None
        buf = C.create_string_buffer(56)
        if _lib.lib().hbn_navmesh_get_settings(self._h, buf) != 0:
            return None
        return NavMeshSettings.from_bytes(buf.raw)

    @property
    def launch_count(self) -> int:
        return int(_lib.lib().hbn_navmesh_launch_count(self._need()))

    def mesh_info(self) -> dict:
        self._need()
        i = self._info
        return dict(num_tiles=i.num_tiles, num_polys=i.num_polys, num_links=i.num_links,
                    num_bv_nodes=i.num_bv_nodes, num_islands=i.num_islands, poly_bits=i.poly_bits,
                    tile_bits=i.tile_bits, salt_bits=i.salt_bits, device_bytes=i.device_bytes)

    def seed(self, new_seed: int):
        """PathFinder::seed, PathFinder.cpp:1225-1229.  Random points then come from the
        counter-based stream hbn_uniform(seed, call_index, draw) (SURVEY.md trap T7)."""
        self._seed = int(new_seed) & 0xFFFFFFFFFFFFFFFF
        self._rand_counter = 0

    # ---- array plumbing ---------------------------------------------------------------
    def _torch_args(self, *tensors):
        import torch
        dev = torch.device("cuda", self._device)
        out = []
        for t in tensors:
            if t is None:
                out.append(None)
                continue
            if t.device != dev:
                raise ValueError(f"tensor on {t.device}, navmesh on {dev}")
            out.append(t.contiguous())
        return torch, dev, out, torch.cuda.current_stream(dev).cuda_stream

    # ---- batched queries --------------------------------------------------------------
    def snap_points(self, pts, islands=None):
        """Batched snap_point / get_island: returns (points [N,3], refs [N] uint32, islands [N])."""
        h = self._need()
        L = _lib.lib()
        if _is_torch(pts):
            torch, dev, (p, isl), st = self._torch_args(
                pts.float().reshape(-1, 3), None if islands is None else islands.to(dtype=__import__("torch").int32))
            n = p.shape[0]
            op = torch.empty((n, 3), dtype=torch.float32, device=dev)
            orf = torch.empty(n, dtype=torch.int32, device=dev)
            oi = torch.empty(n, dtype=torch.int32, device=dev)
            check(L.hbn_snap_point_dev(h, p.data_ptr(), isl.data_ptr() if isl is not None else None, n,
                                       op.data_ptr(), orf.data_ptr(), oi.data_ptr(), st))
            return op, orf, oi
        p = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(-1, 3))
        n = len(p)
        isl = None if islands is None else np.ascontiguousarray(islands, dtype=np.int32)
        op = np.empty((n, 3), np.float32)
        orf = np.empty(n, np.uint32)
        oi = np.empty(n, np.int32)
        check(L.hbn_snap_point(h, p.ctypes.data, isl.ctypes.data if isl is not None else None, n,
                               op.ctypes.data, orf.ctypes.data, oi.ctypes.data))
        return op, orf, oi

    def are_navigable(self, pts, max_y_delta: float = 0.5):
        h = self._need()
        L = _lib.lib()
        if _is_torch(pts):
            torch, dev, (p,), st = self._torch_args(pts.float().reshape(-1, 3))
            out = torch.empty(p.shape[0], dtype=torch.uint8, device=dev)
            check(L.hbn_is_navigable_dev(h, p.data_ptr(), p.shape[0], float(max_y_delta), out.data_ptr(), st))
            return out.bool()
        p = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(-1, 3))
        out = np.empty(len(p), np.uint8)
        check(L.hbn_is_navigable(h, p.ctypes.data, len(p), float(max_y_delta), out.ctypes.data))
        return out.astype(bool)

    def set_profiling(self, enable: bool):
        check(_lib.lib().hbn_navmesh_set_profiling(self._need(), 1 if enable else 0))

    def phase_times(self) -> dict:
        """Device milliseconds spent in find_paths' snap / path kernels since the last call."""
        ms = (C.c_double * 2)()
        calls = C.c_int64()
        check(_lib.lib().hbn_navmesh_phase_times(self._need(), ms, C.byref(calls)))
        return {"snap_ms": ms[0], "path_ms": ms[1], "calls": calls.value}

    def work_counters(self, reset: bool = True) -> dict:
        """Totals of the find_paths(count_work=True) calls so far (hbn_navmesh_work_counters)."""
        out = np.zeros(8, np.uint64)
        check(_lib.lib().hbn_navmesh_work_counters(self._need(), out.ctypes.data, 1 if reset else 0))
        keys = ["expanded", "links", "neighbours", "corridor", "corridor_links", "points",
                "astar_queries", "queries"]
        return {k: int(v) for k, v in zip(keys, out)}

    def find_paths(self, starts, ends, max_points: int = 0, corridors: bool = False,
                   exact_status: bool = False, count_work: bool = False):
        """Batched find_path(ShortestPath).  Returns a dict with `geodesic_distance` [N] and,
        on request, `num_points`, `points` [N,max_points,3] (NaN padded), `corridor` [N,256]
        poly refs, `num_corridor`, `status` [N,2]."""
        h = self._need()
        L = _lib.lib()
        flags = (_lib.HBN_FP_EXACT_STATUS if exact_status else 0) | \
                (_lib.HBN_FP_COUNT_WORK if count_work else 0)
        if _is_torch(starts):
            torch, dev, (s, e), st = self._torch_args(starts.float().reshape(-1, 3), ends.float().reshape(-1, 3))
            n = s.shape[0]
            dist = torch.empty(n, dtype=torch.float32, device=dev)
            res = {"geodesic_distance": dist}
            npts = pts = corr = ncorr = status = None
            if max_points:
                npts = torch.empty(n, dtype=torch.int32, device=dev)
                pts = torch.full((n, max_points, 3), float("nan"), dtype=torch.float32, device=dev)
                res.update(num_points=npts, points=pts)
            if corridors:
                corr = torch.zeros((n, 256), dtype=torch.int32, device=dev)
                ncorr = torch.empty(n, dtype=torch.int32, device=dev)
                status = torch.empty((n, 2), dtype=torch.int32, device=dev)
                res.update(corridor=corr, num_corridor=ncorr, status=status)
            check(L.hbn_find_path_dev(h, s.data_ptr(), e.data_ptr(), n, dist.data_ptr(),
                                      npts.data_ptr() if npts is not None else None,
                                      pts.data_ptr() if pts is not None else None, max_points,
                                      corr.data_ptr() if corr is not None else None,
                                      ncorr.data_ptr() if ncorr is not None else None,
                                      status.data_ptr() if status is not None else None, flags, st))
            return res
        s = np.ascontiguousarray(np.asarray(starts, np.float32).reshape(-1, 3))
        e = np.ascontiguousarray(np.asarray(ends, np.float32).reshape(-1, 3))
        n = len(s)
        dist = np.empty(n, np.float32)
        res = {"geodesic_distance": dist}
        npts = pts = corr = ncorr = status = None
        if max_points:
            npts = np.empty(n, np.int32)
            pts = np.empty((n, max_points, 3), np.float32)
            res.update(num_points=npts, points=pts)
        if corridors:
            corr = np.empty((n, 256), np.uint32)
            ncorr = np.empty(n, np.int32)
            status = np.empty((n, 2), np.uint32)
            res.update(corridor=corr, num_corridor=ncorr, status=status)
        g = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
        check(L.hbn_find_path(h, s.ctypes.data, e.ctypes.data, n, dist.ctypes.data, g(npts), g(pts),
                              max_points, g(corr), g(ncorr), g(status), flags))
        return res

    def geodesic_distances(self, starts, ends):
        return self.find_paths(starts, ends)["geodesic_distance"]

    def find_paths_multigoal(self, starts, ends, max_points: int = 0):
        """Batched find_path(MultiGoalShortestPath): ends is [N, G, 3]."""
        h = self._need()
        L = _lib.lib()
        if _is_torch(starts):
            torch, dev, (s, e), st = self._torch_args(starts.float().reshape(-1, 3), ends.float())
            n, g = e.shape[0], e.shape[1]
            dist = torch.empty(n, dtype=torch.float32, device=dev)
            idx = torch.empty(n, dtype=torch.int32, device=dev)
            npts = torch.empty(n, dtype=torch.int32, device=dev) if max_points else None
            pts = torch.full((n, max_points, 3), float("nan"), dtype=torch.float32, device=dev) if max_points else None
            check(L.hbn_find_path_multigoal_dev(h, s.data_ptr(), e.data_ptr(), n, g, dist.data_ptr(),
                                                idx.data_ptr(), npts.data_ptr() if npts is not None else None,
                                                pts.data_ptr() if pts is not None else None, max_points, st))
            return dict(geodesic_distance=dist, closest_end_point_index=idx, num_points=npts, points=pts)
        s = np.ascontiguousarray(np.asarray(starts, np.float32).reshape(-1, 3))
        e = np.ascontiguousarray(np.asarray(ends, np.float32))
        n, g = e.shape[0], e.shape[1]
        dist = np.empty(n, np.float32)
        idx = np.empty(n, np.int32)
        npts = np.empty(n, np.int32) if max_points else None
        pts = np.empty((n, max_points, 3), np.float32) if max_points else None
        check(L.hbn_find_path_multigoal(h, s.ctypes.data, e.ctypes.data, n, g, dist.ctypes.data,
                                        idx.ctypes.data, npts.ctypes.data if npts is not None else None,
                                        pts.ctypes.data if pts is not None else None, max_points))
        return dict(geodesic_distance=dist, closest_end_point_index=idx, num_points=npts, points=pts)

    def set_option(self, key: str, value: int) -> None:
        """hbn_navmesh_set_option: "lane_scratch_bytes", "lane_cfg", "blocks_per_sm", "lane_spread",
        "snap_spread", "snap_dual", "snap_group", "snap_cap", "nvtx"."""
        check(_lib.lib().hbn_navmesh_set_option(self._need(), key.encode(), int(value)))

    def reserve(self, n: int) -> None:
        """Size every scratch buffer for batches of up to n queries (hbn_navmesh_reserve): later
        calls of that size only enqueue kernels (no allocation, capturable into CUDA graphs)."""
        check(_lib.lib().hbn_navmesh_reserve(self._need(), int(n)))

    @property
    def scratch_bytes(self) -> int:
        return int(_lib.lib().hbn_navmesh_scratch_bytes(self._need()))

    def env_steps(self, positions, targets, goals, allow_sliding: bool = True):
        """One PointNav environment step for N envs (simulator.py:660-673 + habitat-lab's geodesic
        reward): new_pos = try_step(positions, targets), then geodesic_distance(new_pos, goals).
        Returns (new_pos [N,3], dist [N]); equal bit for bit to try_steps + geodesic_distances.
        numpy inputs go through hbn_env_step (one CUDA graph replay per call), torch CUDA tensors
        through hbn_env_step_dev on the current stream."""
        h = self._need()
        L = _lib.lib()
        if _is_torch(positions):
            torch, dev, (p, t, g), st = self._torch_args(positions.float().reshape(-1, 3), targets.float().reshape(-1, 3),
                                                         goals.float().reshape(-1, 3))
            n = p.shape[0]
            out = torch.empty((n, 3), dtype=torch.float32, device=dev)
            dist = torch.empty(n, dtype=torch.float32, device=dev)
            check(L.hbn_env_step_dev(h, p.data_ptr(), t.data_ptr(), g.data_ptr(), n, 1 if allow_sliding else 0,
                                     out.data_ptr(), dist.data_ptr(), st))
            return out, dist
        p = np.ascontiguousarray(np.asarray(positions, np.float32).reshape(-1, 3))
        t = np.ascontiguousarray(np.asarray(targets, np.float32).reshape(-1, 3))
        g = np.ascontiguousarray(np.asarray(goals, np.float32).reshape(-1, 3))
        n = len(p)
        assert len(t) == n and len(g) == n
        out = np.empty((n, 3), np.float32)
        dist = np.empty(n, np.float32)
        check(L.hbn_env_step(h, p.ctypes.data, t.ctypes.data, g.ctypes.data, n, 1 if allow_sliding else 0,
                             out.ctypes.data, dist.ctypes.data))
        return out, dist

    def follower_best_prims(self, rots, poss, goals, goal_dist: float, forward_amount: float, turn_amount: float,
                            n_steps: int, allow_sliding: bool = True):
        """nextBestPrimAlong (GreedyFollower.cpp:83-140) of N agents on the device
        (hbn_follower_best_prims): rots [N,4] float64 quaternions (x, y, z, w), poss [N,3] float64,
        goals [N,3].  Returns (prim int32 [N], geodesic distance now f32 [N]); prim: -2 ERROR, -1 STOP,
        -3 none acceptable, else (turns << 1) | side (0 = LEFT, 1 = RIGHT) before one FORWARD."""
        r = np.ascontiguousarray(np.asarray(rots, np.float64).reshape(-1, 4))
        p = np.ascontiguousarray(np.asarray(poss, np.float64).reshape(-1, 3))
        g = np.ascontiguousarray(np.asarray(goals, np.float32).reshape(-1, 3))
        n = len(r)
        fp = _lib.FollowerParams(float(goal_dist), float(forward_amount), math.sin(turn_amount / 2.0),
                                 math.cos(turn_amount / 2.0), int(n_steps), 1 if allow_sliding else 0)
        prim = np.empty(n, np.int32)
        geo = np.empty(n, np.float32)
        check(_lib.lib().hbn_follower_best_prims(self._need(), r.ctypes.data, p.ctypes.data, g.ctypes.data, n,
                                                 C.byref(fp), prim.ctypes.data, geo.ctypes.data))
        return prim, geo

    def try_steps(self, starts, ends, allow_sliding: bool = True):
        h = self._need()
        L = _lib.lib()
        if _is_torch(starts):
            torch, dev, (s, e), st = self._torch_args(starts.float().reshape(-1, 3), ends.float().reshape(-1, 3))
            out = torch.empty_like(s)
            check(L.hbn_try_step_dev(h, s.data_ptr(), e.data_ptr(), s.shape[0], 1 if allow_sliding else 0,
                                     out.data_ptr(), st))
            return out
        s = np.ascontiguousarray(np.asarray(starts, np.float32).reshape(-1, 3))
        e = np.ascontiguousarray(np.asarray(ends, np.float32).reshape(-1, 3))
        out = np.empty_like(s)
        check(L.hbn_try_step(h, s.ctypes.data, e.ctypes.data, len(s), 1 if allow_sliding else 0,
                             out.ctypes.data))
        return out

    def closest_obstacle_surface_points(self, pts, max_search_radius: float = 2.0):
        """Batched closest_obstacle_surface_point: (hit_pos [N,3], hit_normal [N,3], hit_dist [N])."""
        h = self._need()
        L = _lib.lib()
        if _is_torch(pts):
            torch, dev, (p,), st = self._torch_args(pts.float().reshape(-1, 3))
            n = p.shape[0]
            hp = torch.empty((n, 3), dtype=torch.float32, device=dev)
            hn = torch.empty((n, 3), dtype=torch.float32, device=dev)
            hd = torch.empty(n, dtype=torch.float32, device=dev)
            check(L.hbn_closest_obstacle_dev(h, p.data_ptr(), n, float(max_search_radius), hp.data_ptr(),
                                             hn.data_ptr(), hd.data_ptr(), st))
            return hp, hn, hd
        p = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(-1, 3))
        n = len(p)
        hp = np.empty((n, 3), np.float32)
        hn = np.empty((n, 3), np.float32)
        hd = np.empty(n, np.float32)
        check(L.hbn_closest_obstacle(h, p.ctypes.data, n, float(max_search_radius), hp.ctypes.data,
                                     hn.ctypes.data, hd.ctypes.data))
        return hp, hn, hd

    def distances_to_closest_obstacle(self, pts, max_search_radius: float = 2.0):
        return self.closest_obstacle_surface_points(pts, max_search_radius)[2]

    def random_navigable_points(self, n: int, max_tries: int = 10, island_index=-1, seed=None,
                                query0=None, device_output: bool = False):
        """n samples of get_random_navigable_point.  island_index: int or [n] array."""
        h = self._need()
        L = _lib.lib()
        seed = self._seed if seed is None else int(seed)
        if query0 is None:
            query0 = self._rand_counter
            self._rand_counter += n
        isl = None
        if not isinstance(island_index, (int, np.integer)):
            isl = island_index
        elif island_index != -1:
            if not (0 <= island_index < self.num_islands):
                raise ValueError(f"{island_index} not a valid index for this island system.")
            isl = np.full(n, island_index, np.int32)
        if isinstance(island_index, (int, np.integer)) and self.island_area(int(island_index)) <= 0.0:
            raise RuntimeError("NavMesh has no navigable area, this indicates an issue with the NavMesh")
        if device_output or (isl is not None and _is_torch(isl)):
            import torch
            dev = torch.device("cuda", self._device)
            if isl is not None and not _is_torch(isl):
                isl = torch.as_tensor(np.asarray(isl, np.int32), device=dev)
            out = torch.empty((n, 3), dtype=torch.float32, device=dev)
            refs = torch.empty(n, dtype=torch.int32, device=dev)
            check(L.hbn_random_points_dev(h, seed, query0, n, isl.data_ptr() if isl is not None else None,
                                          max_tries, out.data_ptr(), refs.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream))
            return out, refs
        islnp = None if isl is None else np.ascontiguousarray(isl, dtype=np.int32)
        out = np.empty((n, 3), np.float32)
        refs = np.empty(n, np.uint32)
        check(L.hbn_random_points(h, seed, query0, n, islnp.ctypes.data if islnp is not None else None,
                                  max_tries, out.ctypes.data, refs.ctypes.data))
        return out, refs

    def random_navigable_points_near(self, circle_centers, radius: float, max_tries: int = 100, island_index=-1,
                                     seed=None, query0=None):
        """Batched get_random_navigable_point_near (getRandomNavigablePointInCircle, PF.cpp:1283-1332):
        one sample per centre [N,3]; NaN rows where max_tries samples all fell outside the circle.
        island_index: int or [N] array."""
        h = self._need()
        L = _lib.lib()
        seed = self._seed if seed is None else int(seed)
        torch_in = _is_torch(circle_centers)
        n = int(circle_centers.shape[0]) if torch_in else len(np.asarray(circle_centers).reshape(-1, 3))
        if query0 is None:
            query0 = self._rand_counter
            self._rand_counter += n
        isl = None
        if not isinstance(island_index, (int, np.integer)):
            isl = island_index
        else:
            if island_index != -1 and not (0 <= island_index < self.num_islands):
                raise ValueError(f"{island_index} not a valid index for this island system.")
            if self.island_area(int(island_index)) <= 0.0:
                raise RuntimeError("NavMesh has no navigable area, this indicates an issue with the NavMesh")
            if island_index != -1:
                isl = np.full(n, island_index, np.int32)
        if torch_in:
            import torch
            torch_, dev, (c,), st = self._torch_args(circle_centers.float().reshape(-1, 3))
            if isl is not None and not _is_torch(isl):
                isl = torch.as_tensor(np.asarray(isl, np.int32), device=dev)
            out = torch.empty((n, 3), dtype=torch.float32, device=dev)
            check(L.hbn_random_points_near_dev(h, seed, query0, n, c.data_ptr(), float(radius),
                                               isl.data_ptr() if isl is not None else None, int(max_tries),
                                               out.data_ptr(), st))
            return out
        c = np.ascontiguousarray(np.asarray(circle_centers, np.float32).reshape(-1, 3))
        islnp = None if isl is None else np.ascontiguousarray(isl, dtype=np.int32)
        out = np.empty((n, 3), np.float32)
        check(L.hbn_random_points_near(h, seed, query0, n, c.ctypes.data, float(radius),
                                       islnp.ctypes.data if islnp is not None else None, int(max_tries),
                                       out.ctypes.data))
        return out

    # ---- scalar API, SPB.cpp:177-270 -------------------------------------------------
    def find_path(self, path) -> bool:
        if isinstance(path, MultiGoalShortestPath):
            self._need()
            return multigoal_find_path(path, lambda p: self.snap_points(p)[1],
                                       lambda s, e, m: self.find_paths(s, e, m), std_sort_order)
        r = self.find_paths(_vec3(path.requested_start)[None], _vec3(path.requested_end)[None],
                            MAX_PATH_POINTS)
        d = float(r["geodesic_distance"][0])
        n = int(r["num_points"][0])
        path.geodesic_distance = d
        path.points = [r["points"][0, i].copy() for i in range(n)]
        return d < math.inf

    def try_step(self, start, end):
        return self.try_steps(_vec3(start)[None], _vec3(end)[None], True)[0]

    def try_step_no_sliding(self, start, end):
        return self.try_steps(_vec3(start)[None], _vec3(end)[None], False)[0]

    def snap_point(self, point, island_index: int = -1):
        if island_index != -1 and not (0 <= island_index < self.num_islands):
            raise ValueError(f"{island_index} not a valid index for this island system.")
        isl = None if island_index == -1 else np.array([island_index], np.int32)
        return self.snap_points(_vec3(point)[None], isl)[0][0]

    def get_island(self, point) -> int:
        return int(self.snap_points(_vec3(point)[None])[2][0])

    def is_navigable(self, pt, max_y_delta: float = 0.5) -> bool:
        return bool(self.are_navigable(_vec3(pt)[None], max_y_delta)[0])

    def distance_to_closest_obstacle(self, pt, max_search_radius: float = 2.0) -> float:
        return float(self.distances_to_closest_obstacle(_vec3(pt)[None], max_search_radius)[0])

    def closest_obstacle_surface_point(self, pt, max_search_radius: float = 2.0) -> HitRecord:
        hp, hn, hd = self.closest_obstacle_surface_points(_vec3(pt)[None], max_search_radius)
        return HitRecord(hp[0], hn[0], hd[0])

    def get_random_navigable_point(self, max_tries: int = 10, island_index: int = -1):
        return self.random_navigable_points(1, max_tries, island_index)[0][0]

    # ---- top-down maps, PathFinder.cpp:1833-1896 -------------------------------------
    def get_random_navigable_point_near(self, circle_center, radius: float, max_tries: int = 100,
                                        island_index: int = -1):
        """SPB.cpp:179-183: a random navigable point within `radius` of circle_center (NaN on failure)."""
        return self.random_navigable_points_near(np.asarray(circle_center, np.float32)[None], radius, max_tries,
                                                 island_index)[0]

    def _topdown_grid(self, meters_per_pixel: float, height: float):
        b1, b2 = self.get_bounds()
        mpp = np.float32(meters_per_pixel)
        xspan = np.float32(abs(np.float32(b1[0]) - np.float32(b2[0])))
        zspan = np.float32(abs(np.float32(b1[2]) - np.float32(b2[2])))
        xres = int(np.float32(xspan / mpp))
        zres = int(np.float32(zspan / mpp))
        startx = np.float32(min(b1[0], b2[0]))
        startz = np.float32(min(b1[2], b2[2]))
        # cur = cur + mpp accumulated in float32, row-major like the reference loops
        xs = np.empty(xres, np.float32)
        zs = np.empty(zres, np.float32)
        c = startx
        for w in range(xres):
            xs[w] = c
            c = np.float32(c + mpp)
        c = startz
        for hh in range(zres):
            zs[hh] = c
            c = np.float32(c + mpp)
        pts = np.empty((zres, xres, 3), np.float32)
        pts[:, :, 0] = xs[None, :]
        pts[:, :, 1] = np.float32(height)
        pts[:, :, 2] = zs[:, None]
        return pts

    def get_topdown_view(self, meters_per_pixel: float, height: float, eps: float = 0.5):
        pts = self._topdown_grid(meters_per_pixel, height)
        zres, xres = pts.shape[:2]
        if zres * xres == 0:
            return np.zeros((zres, xres), bool)
        return self.are_navigable(pts.reshape(-1, 3), eps).reshape(zres, xres)

    def get_topdown_island_view(self, meters_per_pixel: float, height: float, eps: float = 0.5):
        pts = self._topdown_grid(meters_per_pixel, height)
        zres, xres = pts.shape[:2]
        if zres * xres == 0:
            return np.zeros((zres, xres), np.int32)
        flat = pts.reshape(-1, 3)
        nav = self.are_navigable(flat, eps)
        isl = self.snap_points(flat)[2]
        return np.where(nav, isl, -1).astype(np.int32).reshape(zres, xres)

    # ---- navmesh geometry, PathFinder.cpp:1898-1968 ------------------------------------
    def build_navmesh_vertices(self, island_index: int = -1):
        h = self._need()
        n = _lib.lib().hbn_navmesh_triangles(h, int(island_index), None, 0)
        out = np.empty((n, 3, 3), np.float32)
        _lib.lib().hbn_navmesh_triangles(h, int(island_index), out.ctypes.data, n)
        return [v.copy() for v in out.reshape(-1, 3)]

    def build_navmesh_vertex_indices(self, island_index: int = -1):
        n = _lib.lib().hbn_navmesh_triangles(self._need(), int(island_index), None, 0)
        return list(range(3 * n))


from .multigpu import MultiGpuPathFinder  # noqa: E402,F401
from .greedy_follower import (GreedyFollowerCodes, GreedyGeodesicFollower,  # noqa: E402
                              GreedyGeodesicFollowerBatch, GreedyGeodesicFollowerBatchImpl,
                              GreedyGeodesicFollowerImpl)

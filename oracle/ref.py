"""ctypes binding of oracle/_ref/libhbn_ref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library is the reference's own, unmodified src/esp/nav/PathFinder.cpp + Detour + Recast,
compiled where they lie under /root/reference by oracle/Makefile (SURVEY.md 8c recipe), behind
the C ABI of oracle/ref_esp.cpp.  `RefPathFinder(restated=True)` binds the round-1 restatement
of the PathFinder layer instead (oracle/_ref/libhbn_restated.so); it exists only so that
tests/test_oracle.py can compare the two bit for bit.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs (and the navmesh *input* builders they use) may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libhbn_ref.so")
_SO_RESTATED = os.path.join(_HERE, "_ref", "libhbn_restated.so")
_libs = {}

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile the oracle library (needs /root/reference); returns the .so path."""
    if force or not os.path.exists(_SO) or not os.path.exists(_SO_RESTATED):
        if not os.path.isdir("/root/reference/src/deps/recastnavigation"):
            raise RuntimeError(
                "oracle/_ref/libhbn_ref.so is missing and /root/reference is not present to build it")
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE])
    return _SO


def lib(restated: bool = False):
    l = _libs.get(restated)
    if l is None:
        build()
        l = C.CDLL(_SO_RESTATED if restated else _SO)
        l.ref_create.restype = C.c_void_p
        l.ref_multigoal_create.restype = C.c_void_p
        l.ref_save_memory.restype = C.c_int64
        l.ref_poly_islands.restype = C.c_int64
        l.ref_navigable_area.restype = C.c_float
        l.ref_island_radius.restype = C.c_float
        l.ref_uniform.restype = C.c_float
        if not restated:
            assert l.ref_is_reference_pathfinder() == 1
            l.ref_frand_of_stream.restype = C.c_float
            l.ref_topdown_view.restype = C.c_int64
            l.ref_navmesh_vertices.restype = C.c_int64
        _libs[restated] = l
    return l


def _f32(a, shape_last=3):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


class RefPathFinder:
    """The reference's esp::nav::PathFinder on the CPU (restated=True: the round-1 restatement of
    that layer over the same Detour, kept for the cross-check only)."""

    def __init__(self, restated: bool = False):
        self.restated = restated
        self._l = lib(restated)
        self._h = C.c_void_p(self._l.ref_create())

    def __del__(self):
        try:
            if self._h:
                self._l.ref_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- construction -------------------------------------------------------
    @staticmethod
    def default_settings() -> bytes:
        buf = C.create_string_buffer(56)
        lib().ref_default_settings(buf)
        return buf.raw

    def build(self, verts, tris, settings: bytes | None = None) -> bool:
        v = _f32(verts).reshape(-1, 3)
        t = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        s = settings or self.default_settings()
        return bool(self._l.ref_build(self._h, s, _p(v, f32p), C.c_int(len(v)), _p(t, i32p),
                                      C.c_int(len(t))))

    def build_tiled(self, verts, tris, tile_size: int = 256, nthreads: int = 0,
                    settings: bytes | None = None) -> bool:
        v = _f32(verts).reshape(-1, 3)
        t = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        s = settings or self.default_settings()
        nthreads = nthreads or (os.cpu_count() or 1)
        return bool(self._l.ref_build_tiled(self._h, s, _p(v, f32p), C.c_int(len(v)),
                                            _p(t, i32p), C.c_int(len(t)), C.c_int(tile_size),
                                            C.c_int(nthreads)))

    def load(self, path: str) -> bool:
        return bool(self._l.ref_load(self._h, path.encode()))

    def load_bytes(self, data: bytes) -> bool:
        return bool(self._l.ref_load_memory(self._h, data, C.c_int64(len(data))))

    def save(self, path: str) -> bool:
        return bool(self._l.ref_save(self._h, path.encode()))

    def save_bytes(self) -> bytes:
        n = self._l.ref_save_memory(self._h, None, C.c_int64(0))
        if n < 0:
            raise RuntimeError("no navmesh")
        buf = C.create_string_buffer(n)
        self._l.ref_save_memory(self._h, buf, C.c_int64(n))
        return buf.raw

    # ---- properties ---------------------------------------------------------
    @property
    def is_loaded(self) -> bool:
        return bool(self._l.ref_is_loaded(self._h))

    @property
    def num_islands(self) -> int:
        return int(self._l.ref_num_islands(self._h))

    def navigable_area(self, island: int = -1) -> float:
        return float(self._l.ref_navigable_area(self._h, C.c_int(island)))

    def island_radius(self, island: int) -> float:
        return float(self._l.ref_island_radius(self._h, C.c_int(island)))

    def get_bounds(self):
        out = np.zeros(6, np.float32)
        self._l.ref_get_bounds(self._h, _p(out, f32p))
        return out[:3].copy(), out[3:].copy()

    def seed(self, s: int):
        self._l.ref_seed(self._h, C.c_uint32(s))

    def mesh_stats(self) -> dict:
        out = np.zeros(8, np.int64)
        self._l.ref_mesh_stats(self._h, _p(out, i64p))
        keys = ["tiles", "polys", "verts", "links", "bv_nodes", "detail_tris", "detail_verts",
                "tile_bytes"]
        return dict(zip(keys, (int(x) for x in out)))

    def poly_islands(self):
        n = self._l.ref_poly_islands(self._h, None, None, C.c_int64(0))
        isl = np.zeros(n, np.int32)
        refs = np.zeros(n, np.uint32)
        self._l.ref_poly_islands(self._h, _p(isl, i32p), _p(refs, u32p), C.c_int64(n))
        return refs, isl

    def tile_blobs(self):
        """Finalised tile blobs [(tileRef, tableIndex, bytes)] in tile-table order."""
        out = []
        for i in range(self._l.ref_tile_count(self._h)):
            ref = C.c_uint32(0)
            idx = C.c_int(0)
            n = self._l.ref_tile_blob(self._h, i, None, 0, C.byref(ref), C.byref(idx))
            buf = C.create_string_buffer(n)
            self._l.ref_tile_blob(self._h, i, buf, n, C.byref(ref), C.byref(idx))
            out.append((ref.value, idx.value, buf.raw))
        return out

    def navmesh_params(self):
        orig = np.zeros(3, np.float32)
        wh = np.zeros(2, np.float32)
        mm = np.zeros(2, np.int32)
        self._l.ref_navmesh_params(self._h, _p(orig, f32p), _p(wh, f32p), _p(mm, i32p))
        return orig, wh, mm

    # ---- batched queries ----------------------------------------------------
    def snap_batch(self, pts, nthreads: int = 1):
        p = _f32(pts).reshape(-1, 3)
        n = len(p)
        out = np.empty((n, 3), np.float32)
        refs = np.empty(n, np.uint32)
        isl = np.empty(n, np.int32)
        self._l.ref_snap_batch(self._h, _p(p, f32p), C.c_int64(n), _p(out, f32p), _p(refs, u32p),
                               _p(isl, i32p), C.c_int(nthreads))
        return out, refs, isl

    def snap_island_batch(self, pts, islands):
        p = _f32(pts).reshape(-1, 3)
        n = len(p)
        isl = np.ascontiguousarray(islands, dtype=np.int32)
        out = np.empty((n, 3), np.float32)
        refs = np.empty(n, np.uint32)
        self._l.ref_snap_island_batch(self._h, _p(p, f32p), _p(isl, i32p), C.c_int64(n),
                                      _p(out, f32p), _p(refs, u32p))
        return out, refs

    def is_navigable_batch(self, pts, max_y_delta: float = 0.5, nthreads: int = 1):
        p = _f32(pts).reshape(-1, 3)
        out = np.empty(len(p), np.uint8)
        self._l.ref_is_navigable_batch(self._h, _p(p, f32p), C.c_int64(len(p)),
                                       C.c_float(max_y_delta), _p(out, u8p), C.c_int(nthreads))
        return out.astype(bool)

    def island_radius_batch(self, pts, nthreads: int = 1):
        p = _f32(pts).reshape(-1, 3)
        out = np.empty(len(p), np.float32)
        self._l.ref_island_radius_batch(self._h, _p(p, f32p), C.c_int64(len(p)), _p(out, f32p),
                                        C.c_int(nthreads))
        return out

    def find_path_batch(self, starts, ends, max_pts: int = 0, nthreads: int = 1):
        s = _f32(starts).reshape(-1, 3)
        e = _f32(ends).reshape(-1, 3)
        n = len(s)
        dist = np.empty(n, np.float32)
        npts = np.empty(n, np.int32)
        pts = np.full((n, max_pts, 3), np.nan, np.float32) if max_pts else None
        self._l.ref_find_path_batch(self._h, _p(s, f32p), _p(e, f32p), C.c_int64(n),
                                    _p(dist, f32p), _p(npts, i32p), _p(pts, f32p),
                                    C.c_int(max_pts), C.c_int(nthreads))
        return dist, npts, pts

    def find_path_raw_batch(self, starts, ends, max_pts: int = 0, nthreads: int = 1):
        s = _f32(starts).reshape(-1, 3)
        e = _f32(ends).reshape(-1, 3)
        n = len(s)
        dist = np.empty(n, np.float32)
        corridor = np.zeros((n, 256), np.uint32)
        info = np.zeros((n, 8), np.uint32)
        pts = np.full((n, max_pts, 3), np.nan, np.float32) if max_pts else None
        self._l.ref_find_path_raw_batch(self._h, _p(s, f32p), _p(e, f32p), C.c_int64(n),
                                        _p(dist, f32p), _p(corridor, u32p), _p(info, u32p),
                                        _p(pts, f32p), C.c_int(max_pts), C.c_int(nthreads))
        return dict(dist=dist, corridor=corridor, start_ref=info[:, 0], end_ref=info[:, 1],
                    astar_status=info[:, 2], straight_status=info[:, 3],
                    num_polys=info[:, 4].astype(np.int32), num_points=info[:, 5].astype(np.int32),
                    nodes_used=info[:, 6].astype(np.int32), flags=info[:, 7], pts=pts)

    def find_path_stats_batch(self, starts, ends, nthreads: int = 1):
        """Work counters for the roofline's algorithmic bytes: dict of [n] arrays
        expanded, links, neighbours, corridor, corridor_links, points."""
        s = _f32(starts).reshape(-1, 3)
        e = _f32(ends).reshape(-1, 3)
        out = np.zeros((len(s), 8), np.uint32)
        self._l.ref_find_path_stats_batch(self._h, _p(s, f32p), _p(e, f32p), C.c_int64(len(s)),
                                          _p(out, u32p), C.c_int(nthreads))
        keys = ["expanded", "links", "neighbours", "corridor", "corridor_links", "points", "nodes", "open_at_end"]
        return {k: out[:, i].astype(np.int64) for i, k in enumerate(keys)}

    def find_path_multigoal_batch(self, starts, ends, max_pts: int = 0, nthreads: int = 1):
        s = _f32(starts).reshape(-1, 3)
        e = _f32(ends)
        n, g = e.shape[0], e.shape[1]
        dist = np.empty(n, np.float32)
        idx = np.empty(n, np.int32)
        npts = np.empty(n, np.int32)
        pts = np.full((n, max_pts, 3), np.nan, np.float32) if max_pts else None
        self._l.ref_find_path_multigoal_batch(self._h, _p(s, f32p), _p(e, f32p), C.c_int64(n),
                                              C.c_int(g), _p(dist, f32p), _p(idx, i32p),
                                              _p(npts, i32p), _p(pts, f32p), C.c_int(max_pts),
                                              C.c_int(nthreads))
        return dist, idx, npts, pts

    def try_step_batch(self, starts, ends, allow_sliding: bool = True, nthreads: int = 1):
        s = _f32(starts).reshape(-1, 3)
        e = _f32(ends).reshape(-1, 3)
        out = np.empty_like(s)
        self._l.ref_try_step_batch(self._h, _p(s, f32p), _p(e, f32p), C.c_int64(len(s)),
                                   C.c_int(1 if allow_sliding else 0), _p(out, f32p),
                                   C.c_int(nthreads))
        return out

    def move_along_surface_batch(self, starts, ends, nthreads: int = 1):
        s = _f32(starts).reshape(-1, 3)
        e = _f32(ends).reshape(-1, 3)
        n = len(s)
        pos = np.empty((n, 3), np.float32)
        vis = np.zeros((n, 16), np.uint32)
        nv = np.zeros(n, np.int32)
        self._l.ref_move_along_surface_batch(self._h, _p(s, f32p), _p(e, f32p), C.c_int64(n),
                                             _p(pos, f32p), _p(vis, u32p), _p(nv, i32p),
                                             C.c_int(nthreads))
        return pos, vis, nv

    def obstacle_batch(self, pts, max_radius: float = 2.0, nthreads: int = 1):
        p = _f32(pts).reshape(-1, 3)
        out = np.empty((len(p), 7), np.float32)
        self._l.ref_obstacle_batch(self._h, _p(p, f32p), C.c_int64(len(p)), C.c_float(max_radius),
                                   _p(out, f32p), C.c_int(nthreads))
        return out[:, 0:3], out[:, 3:6], out[:, 6]

    def random_points(self, n: int, max_tries: int = 10, islands=None, mode: int = 0,
                      seed: int = 0, query0: int = 0):
        out = np.empty((n, 3), np.float32)
        refs = np.empty(n, np.uint32)
        isl = None if islands is None else np.ascontiguousarray(islands, dtype=np.int32)
        ok = self._l.ref_random_points(self._h, C.c_int64(n), C.c_int(max_tries), _p(isl, i32p),
                                       C.c_int(mode), C.c_uint64(seed), C.c_uint64(query0),
                                       _p(out, f32p), _p(refs, u32p))
        if not ok:
            raise RuntimeError("NavMesh has no navigable area, this indicates an issue with the NavMesh")
        return out, refs

    def random_points_near(self, centers, radius: float, max_tries: int = 100, islands=None,
                           mode: int = 0, seed: int = 0, query0: int = 0):
        c = _f32(centers).reshape(-1, 3)
        out = np.empty_like(c)
        isl = None if islands is None else np.ascontiguousarray(islands, dtype=np.int32)
        ok = self._l.ref_random_points_near(self._h, C.c_int64(len(c)), _p(c, f32p),
                                            C.c_float(radius), C.c_int(max_tries), _p(isl, i32p),
                                            C.c_int(mode), C.c_uint64(seed), C.c_uint64(query0),
                                            _p(out, f32p))
        if not ok:
            raise RuntimeError("NavMesh has no navigable area, this indicates an issue with the NavMesh")
        return out


    # ---- only the reference's own PathFinder has these --------------------------------
    def topdown_view(self, meters_per_pixel: float, height: float, eps: float = 0.5, islands: bool = False):
        """get_topdown_view (bool [H, W]) / get_topdown_island_view (int32 [H, W]), PF.cpp:1833-1896."""
        dims = np.zeros(2, np.int32)
        args = (self._h, C.c_float(meters_per_pixel), C.c_float(height), C.c_float(eps), C.c_int(1 if islands else 0))
        n = self._l.ref_topdown_view(*args, None, C.c_int64(0), _p(dims, i32p))
        out = np.zeros(max(int(n), 1), np.int32)
        self._l.ref_topdown_view(*args, _p(out, i32p), C.c_int64(n), _p(dims, i32p))
        g = out[:n].reshape(int(dims[0]), int(dims[1]))
        return g if islands else g.astype(bool)

    def navmesh_vertices(self, island: int = -1):
        """build_navmesh_vertices / build_navmesh_vertex_indices (getNavMeshData, PF.cpp:1898-1944)."""
        n = self._l.ref_navmesh_vertices(self._h, C.c_int(island), None, None, C.c_int64(0))
        if n < 0:
            raise ValueError(f"{island} not a valid index for this island system.")
        v = np.zeros((max(int(n), 1), 3), np.float32)
        idx = np.zeros(max(int(n), 1), np.uint32)
        self._l.ref_navmesh_vertices(self._h, C.c_int(island), _p(v, f32p), _p(idx, u32p), C.c_int64(n))
        return v[:n], idx[:n]


class RefMultiGoal:
    """Stateful MultiGoalShortestPath twin (PF.cpp:95-123) for cache / trap-T4 tests."""

    def __init__(self, pf: RefPathFinder):
        self._pf = pf
        self._l = pf._l
        self._m = C.c_void_p(self._l.ref_multigoal_create())

    def __del__(self):
        try:
            self._l.ref_multigoal_destroy(self._m)
        except Exception:
            pass

    def set_ends(self, ends):
        e = _f32(ends).reshape(-1, 3)
        self._l.ref_multigoal_set_ends(self._m, _p(e, f32p), C.c_int(len(e)))

    def find(self, start, max_pts: int = 256):
        s = _f32(start).reshape(3)
        dist = C.c_float()
        idx = C.c_int32()
        npts = C.c_int32()
        pts = np.full((max_pts, 3), np.nan, np.float32)
        ok = self._l.ref_multigoal_find(self._pf._h, self._m, _p(s, f32p), C.byref(dist),
                                        C.byref(idx), C.byref(npts), _p(pts, f32p),
                                        C.c_int(max_pts))
        return bool(ok), dist.value, idx.value, pts[:npts.value].copy()


def uniform(seed: int, query: int, draw: int) -> float:
    return float(lib().ref_uniform(C.c_uint64(seed), C.c_uint64(query), C.c_uint32(draw)))


def frand_of_stream(seed: int, query: int, draw: int) -> float:
    """What the reference's unmodified frand() (PF.cpp:1232-1234) returns when the oracle's
    interposed rand() serves draw `draw` of the counter-based stream."""
    return float(lib().ref_frand_of_stream(C.c_uint64(seed), C.c_uint64(query), C.c_uint32(draw)))


def random_point_in_convex_poly(pts, s: float, t: float):
    p = _f32(pts).reshape(-1, 3)
    out = np.zeros(3, np.float32)
    lib().ref_random_point_in_convex_poly(_p(p, f32p), C.c_int(len(p)), C.c_float(s), C.c_float(t),
                                          _p(out, f32p))
    return out


def std_sort_order(keys):
    """libstdc++ std::sort order of indices by float key (the goal ordering of PF.cpp:1542-1548)."""
    k = np.ascontiguousarray(keys, dtype=np.float32)
    out = np.zeros(len(k), np.int32)
    lib().ref_std_sort_order(_p(k, f32p), C.c_int(len(k)), _p(out, i32p))
    return out

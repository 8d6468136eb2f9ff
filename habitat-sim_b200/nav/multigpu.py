"""In-process multi-GPU dispatcher (north star: "large batches split across the 8 GPUs of
one box as independent query shards against a replicated navmesh, with no NCCL and no CPU fallback").

`MultiGpuPathFinder(devices)` holds one `PathFinder` handle per device -- the navmesh is replicated at
load -- and answers a batched call by cutting the batch into contiguous, granule-aligned slices
(`shard_slices`), running every slice through the host-buffer C-ABI entry point of its device on a
worker thread of its own (ctypes releases the GIL for the duration of the call, so the H2D copies,
kernels and D2H copies of the devices overlap), and writing every slice's result straight into its
rows of ONE output array.  There is no collective and no inter-GPU traffic: queries are independent
and read-only against the navmesh.  Random-point streams are keyed by the global query index, so the
answers do not depend on the number of devices.

The scalar reference API (`find_path(ShortestPath)`, `try_step`, `snap_point`, properties ...) is served
by the first device's handle.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .sharding import shard_slices


class MultiGpuPathFinder:
    def __init__(self, devices=None, _factory=None):
        """devices: iterable of CUDA device indices (default: every visible device).
        _factory(device) -> PathFinder-like object (tests inject a host-emulation backend)."""
        if _factory is None:
            from . import PathFinder
            from .. import _lib
            _factory = PathFinder
            if devices is None:
                devices = range(max(1, _lib.lib().hbn_device_count()))
        self.devices = list(devices) if devices is not None else [0]
        if not self.devices:
            raise ValueError("no devices")
        self._pfs = [_factory(d) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self._pfs), thread_name_prefix="hbn-gpu")
        self._rand_counter = 0
        self._lock = threading.Lock()

    # ---- navmesh: replicated on every device --------------------------------------------------
    def load_nav_mesh_bytes(self, data: bytes) -> bool:
        return all(self._map(lambda pf: pf.load_nav_mesh_bytes(data)))

    def load_nav_mesh(self, path: str) -> bool:
        with open(path, "rb") as f:
            return self.load_nav_mesh_bytes(f.read())

    def load_from_tiles(self, tiles, *args, **kw) -> bool:
        tiles = list(tiles)
        return all(self._map(lambda pf: pf.load_from_tiles(tiles, *args, **kw)))

    def reserve(self, n: int) -> None:
        """size every handle's scratch for its share of batches of up to n queries"""
        per = max(e - b for b, e in shard_slices(n, len(self._pfs)))
        self._map(lambda pf: pf.reserve(max(1, per)))

    def set_option(self, key: str, value: int) -> None:
        self._map(lambda pf: pf.set_option(key, value))

    def __getattr__(self, name):  # scalar API and properties: the first device's handle
        return getattr(self._pfs[0], name)

    @property
    def world(self) -> int:
        return len(self._pfs)

    @property
    def launch_count(self) -> int:
        return sum(pf.launch_count for pf in self._pfs)

    def _map(self, fn):
        return list(self._pool.map(fn, self._pfs))

    def _run(self, n: int, granule: int, call):
        """call(pf, begin, end) for every non-empty slice, one thread per device"""
        slices = [(pf, b, e) for pf, (b, e) in zip(self._pfs, shard_slices(n, len(self._pfs), granule)) if e > b]
        futs = [self._pool.submit(call, pf, b, e) for pf, b, e in slices]
        for f in futs:
            f.result()  # re-raises a worker's exception

    @staticmethod
    def _pts(a):
        return np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1, 3))

    # ---- batched queries: split -> per-device host-buffer call -> one output array -------------
    def find_paths(self, starts, ends, max_points: int = 0, corridors: bool = False, exact_status: bool = False):
        s, e = self._pts(starts), self._pts(ends)
        n = len(s)
        out = dict(geodesic_distance=np.empty(n, np.float32))
        if max_points:
            out["num_points"] = np.empty(n, np.int32)
            out["points"] = np.empty((n, max_points, 3), np.float32)
        if corridors:
            out["corridor"] = np.empty((n, 256), np.uint32)
            out["num_corridor"] = np.empty(n, np.int32)
        if exact_status or corridors:
            out["status"] = np.empty((n, 2), np.uint32)

        def call(pf, b, e_):
            r = pf.find_paths(s[b:e_], e[b:e_], max_points=max_points, corridors=corridors, exact_status=exact_status)
            for k in out:
                out[k][b:e_] = r[k]
        self._run(n, 1, call)
        return out

    def geodesic_distances(self, starts, ends):
        return self.find_paths(starts, ends)["geodesic_distance"]

    def find_paths_multigoal(self, starts, ends, max_points: int = 0):
        s = self._pts(starts)
        e = np.ascontiguousarray(np.asarray(ends, np.float32))
        n = len(s)
        out = dict(geodesic_distance=np.empty(n, np.float32), closest_end_point_index=np.empty(n, np.int32))
        if max_points:
            out["num_points"] = np.empty(n, np.int32)
            out["points"] = np.empty((n, max_points, 3), np.float32)

        def call(pf, b, e_):
            r = pf.find_paths_multigoal(s[b:e_], e[b:e_], max_points=max_points)
            for k in out:
                out[k][b:e_] = r[k]
        self._run(n, 1, call)  # shards over starts; the reduction over a start's goals is intra-GPU
        return out

    def try_steps(self, starts, ends, allow_sliding: bool = True):
        s, e = self._pts(starts), self._pts(ends)
        out = np.empty_like(s)

        def call(pf, b, e_):
            out[b:e_] = pf.try_steps(s[b:e_], e[b:e_], allow_sliding)
        self._run(len(s), 1, call)
        return out

    def env_steps(self, positions, targets, goals, allow_sliding: bool = True):
        """an env's try_step -> find_path chain stays on one device (slices are by env index)"""
        p, t, g = self._pts(positions), self._pts(targets), self._pts(goals)
        pos, dist = np.empty_like(p), np.empty(len(p), np.float32)

        def call(pf, b, e_):
            pos[b:e_], dist[b:e_] = pf.env_steps(p[b:e_], t[b:e_], g[b:e_], allow_sliding)
        self._run(len(p), 1, call)
        return pos, dist

    def snap_points(self, pts, islands=None):
        p = self._pts(pts)
        isl = None if islands is None else np.ascontiguousarray(islands, dtype=np.int32)
        out_p, out_r, out_i = np.empty_like(p), np.empty(len(p), np.uint32), np.empty(len(p), np.int32)

        def call(pf, b, e_):
            out_p[b:e_], out_r[b:e_], out_i[b:e_] = pf.snap_points(p[b:e_], None if isl is None else isl[b:e_])
        self._run(len(p), 1, call)
        return out_p, out_r, out_i

    def are_navigable(self, pts, max_y_delta: float = 0.5):
        p = self._pts(pts)
        out = np.empty(len(p), bool)

        def call(pf, b, e_):
            out[b:e_] = pf.are_navigable(p[b:e_], max_y_delta)
        self._run(len(p), 1, call)
        return out

    def closest_obstacle_surface_points(self, pts, max_search_radius: float = 2.0):
        p = self._pts(pts)
        hp, hn, hd = np.empty_like(p), np.empty_like(p), np.empty(len(p), np.float32)

        def call(pf, b, e_):
            hp[b:e_], hn[b:e_], hd[b:e_] = pf.closest_obstacle_surface_points(p[b:e_], max_search_radius)
        self._run(len(p), 1, call)
        return hp, hn, hd

    def distances_to_closest_obstacle(self, pts, max_search_radius: float = 2.0):
        return self.closest_obstacle_surface_points(pts, max_search_radius)[2]

    def random_navigable_points(self, n: int, max_tries: int = 10, island_index=-1, seed=None, query0=None):
        """sample i draws from the stream of GLOBAL index query0 + i whatever device answers it"""
        with self._lock:
            if query0 is None:
                query0 = self._rand_counter
                self._rand_counter += n
        seed = self._pfs[0]._seed if seed is None else int(seed)
        out, refs = np.empty((n, 3), np.float32), np.empty(n, np.uint32)
        per_query = not isinstance(island_index, (int, np.integer))
        isl = np.ascontiguousarray(island_index, dtype=np.int32) if per_query else None

        def call(pf, b, e_):
            out[b:e_], refs[b:e_] = pf.random_navigable_points(e_ - b, max_tries, isl[b:e_] if per_query else island_index,
                                                               seed=seed, query0=query0 + b)
        self._run(n, 1, call)
        return out, refs

    def close(self):
        self._pool.shutdown(wait=True)
        self._pfs = []

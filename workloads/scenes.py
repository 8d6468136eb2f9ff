"""Navmesh inputs and seeded query sets for the BASELINE.json configs (C1..C5).

Navmeshes are built from the procedural scenes of workloads/meshgen.py by the REFERENCE's
Recast path (oracle/_ref, following PathFinder.cpp:612-930; the tiled variant follows
RecastDemo/Source/Sample_TileMesh.cpp:794-1160) and handed around as MSET v2 images, i.e.
exactly what `PathFinder::loadNavMesh` reads.  Everything is round-tripped through
save -> load so that island numbering is the loaded-mesh numbering (SURVEY.md trap T5).
Query generation uses a small numpy reader of the image, never the oracle's samplers.
"""
from __future__ import annotations

import os
import struct

import numpy as np

from . import meshgen

_CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")

SCENES = {
    # name: (generator, kwargs, tiled)
    "c1_room": (meshgen.single_room, {}, False),
    "c2_apartment": (meshgen.apartment, {"seed": 0}, False),
    "c3_multiroom": (meshgen.multi_room, {"nx": 10, "nz": 10, "seed": 0}, False),
    "c4_building": (meshgen.building, {"nx": 28, "nz": 28, "floors": 4, "seed": 0}, True),
    # small tiled multi-floor mesh for the parity tests (ramps -> detail verts, many islands)
    "t_building": (meshgen.building, {"nx": 8, "nz": 8, "floors": 3, "ramps_per_floor": 2,
                                      "closed_rooms": 2, "seed": 1}, True),
}


def scene_triangles(name: str):
    """(vertices [V,3] f32, triangles [T,3] i32) of a named procedural scene"""
    gen, kw, _tiled = SCENES[name]
    return gen(**kw)


def navmesh_bytes(name: str, cache: bool = True) -> bytes:
    """MSET v2 image of a named scene (built once with the reference Recast, then cached)."""
    path = os.path.join(_CACHE, name + ".navmesh")
    if cache and os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    from oracle.ref import RefPathFinder  # input builder (reference Recast), see module doc
    gen, kw, tiled = SCENES[name]
    v, t = gen(**kw)
    pf = RefPathFinder()
    ok = pf.build_tiled(v, t, 256) if tiled else pf.build(v, t)
    if not ok:
        raise RuntimeError(f"Recast build of scene {name} failed")
    data = pf.save_bytes()
    if cache:
        os.makedirs(_CACHE, exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp"  # ranks of one torchrun job may race here
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
    return data


class NavMeshGeom:
    """Minimal numpy reader of an MSET image: walkable polygons as triangle fans."""

    def __init__(self, image: bytes):
        magic, version, ntiles = struct.unpack_from("<iii", image, 0)
        assert magic == (ord("M") << 24 | ord("S") << 16 | ord("E") << 8 | ord("T"))
        off = 12 + 28 + (56 if version >= 2 else 0)
        tris = []
        for _ in range(ntiles):
            ref, size = struct.unpack_from("<Ii", image, off)
            off += 8
            h = struct.unpack_from("<5iI9i3f7f", image, off)
            poly_count, vert_count = h[6], h[7]
            voff = off + 100
            verts = np.frombuffer(image, np.float32, vert_count * 3, voff).reshape(-1, 3)
            poff = voff + vert_count * 12
            polys = np.frombuffer(image, np.uint8, poly_count * 32, poff).reshape(-1, 32)
            vidx = polys[:, 4:16].copy().view(np.uint16).reshape(-1, 6)
            flags = polys[:, 28:30].copy().view(np.uint16).reshape(-1)
            nv = polys[:, 30]
            typ = polys[:, 31] >> 6
            for p in range(poly_count):
                if typ[p] != 0 or (flags[p] & 1) == 0:
                    continue
                for j in range(2, nv[p]):
                    tris.append((verts[vidx[p, 0]], verts[vidx[p, j - 1]], verts[vidx[p, j]]))
            off += size
        self.tris = np.asarray(tris, np.float32).reshape(-1, 3, 3)
        a = self.tris[:, 1] - self.tris[:, 0]
        b = self.tris[:, 2] - self.tris[:, 0]
        self.area = 0.5 * np.abs(a[:, 0] * b[:, 2] - a[:, 2] * b[:, 0])
        self.cdf = np.cumsum(self.area / self.area.sum())

    def sample(self, n: int, rng) -> np.ndarray:
        """n points distributed uniformly (by 2D area) over the walkable polygons."""
        t = np.minimum(np.searchsorted(self.cdf, rng.random(n)), len(self.cdf) - 1)
        r1 = np.sqrt(rng.random(n))[:, None]
        r2 = rng.random(n)[:, None]
        tri = self.tris[t]
        return ((1 - r1) * tri[:, 0] + r1 * (1 - r2) * tri[:, 1] + r1 * r2 * tri[:, 2]).astype(np.float32)


def pointnav_pairs(geom: NavMeshGeom, n: int, seed: int, local_frac: float = 0.5,
                   local_radius: float = 15.0, jitter: float = 0.05):
    """C4 find_path mix: `local_frac` PointNav-like pairs whose goal lies within
    `local_radius` metres of the start on the same storey (|dy| <= 0.5 m, the episode filter
    PointNav generators apply), the rest uniformly random pairs."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    starts = geom.sample(n, rng)
    ends = geom.sample(n, rng)
    n_local = int(n * local_frac)
    if n_local:
        pool = geom.sample(max(20000, n_local // 4), rng)
        ys = np.array([1.0, local_radius / 0.5, 1.0])  # |dy| > 0.5 m lies outside the ball
        tree = cKDTree(pool * ys)
        pick = rng.random(n_local)
        # in chunks, on all cores: a 1 M batch has ~1400 neighbours per start (0.7 G indices in
        # all); every start is answered independently, so the lists do not depend on the split
        chunk = 20000
        for c0 in range(0, n_local, chunk):
            nb = tree.query_ball_point(starts[c0:min(c0 + chunk, n_local)] * ys, local_radius,
                                       return_sorted=False, workers=-1)
            for i, lst in enumerate(nb, start=c0):
                if lst:
                    ends[i] = pool[lst[int(pick[i] * len(lst))]]
    perm = rng.permutation(n)
    starts, ends = starts[perm], ends[perm]
    if jitter:
        starts = starts + rng.normal(0, jitter, starts.shape).astype(np.float32) * np.array([1, 0, 1], np.float32)
        ends = ends + rng.normal(0, jitter, ends.shape).astype(np.float32) * np.array([1, 0, 1], np.float32)
    return np.ascontiguousarray(starts, np.float32), np.ascontiguousarray(ends, np.float32)


def uniform_pairs(geom: NavMeshGeom, n: int, seed: int, jitter: float = 0.2):
    rng = np.random.default_rng(seed)
    s = geom.sample(n, rng) + rng.normal(0, jitter, (n, 3)).astype(np.float32)
    e = geom.sample(n, rng) + rng.normal(0, jitter, (n, 3)).astype(np.float32)
    return s.astype(np.float32), e.astype(np.float32)


def step_targets(starts: np.ndarray, seed: int, step: float = 0.25):
    """C2 per-step motion: p + step * dir(theta), theta seeded uniform."""
    rng = np.random.default_rng(seed)
    th = rng.uniform(0, 2 * np.pi, len(starts)).astype(np.float32)
    d = np.stack([np.cos(th), np.zeros_like(th), np.sin(th)], 1).astype(np.float32)
    return (starts + np.float32(step) * d).astype(np.float32)

"""Seeded procedural scene geometry (triangle soups) for the BASELINE.json configs.

These are *inputs*: the navmesh itself is always built from them on the host by the
reference's own Recast path (oracle/ref_pathfinder.cpp::recastBuildOne, which follows
src/esp/nav/PathFinder.cpp:612-930).  y is up, like habitat.
"""
from __future__ import annotations

import numpy as np


class Soup:
    def __init__(self):
        self.v: list = []
        self.t: list = []

    def _add(self, verts, tris):
        base = len(self.v)
        self.v.extend(verts)
        self.t.extend([(a + base, b + base, c + base) for a, b, c in tris])

    def quad_up(self, x0, z0, x1, z1, y):
        """Horizontal quad with +y normal."""
        self._add([(x0, y, z0), (x0, y, z1), (x1, y, z1), (x1, y, z0)], [(0, 1, 2), (0, 2, 3)])

    def ramp(self, x0, z0, x1, z1, y0, y1):
        """Inclined quad rising along +x from y0 (at x0) to y1 (at x1), +y-ish normal."""
        self._add([(x0, y0, z0), (x0, y0, z1), (x1, y1, z1), (x1, y1, z0)], [(0, 1, 2), (0, 2, 3)])

    def box(self, x0, y0, z0, x1, y1, z1):
        v = [(x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1),
             (x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1)]
        t = [(4, 7, 6), (4, 6, 5),  # top (+y)
             (0, 1, 2), (0, 2, 3),  # bottom
             (0, 4, 5), (0, 5, 1),  # z0 side
             (3, 2, 6), (3, 6, 7),  # z1 side
             (0, 3, 7), (0, 7, 4),  # x0 side
             (1, 5, 6), (1, 6, 2)]  # x1 side
        self._add(v, t)

    def arrays(self):
        return (np.asarray(self.v, dtype=np.float32).reshape(-1, 3),
                np.asarray(self.t, dtype=np.int32).reshape(-1, 3))


def single_room(size_x: float = 10.0, size_z: float = 10.0, obstacle: bool = True):
    """C1: one rectangular room, optionally with one box obstacle."""
    s = Soup()
    s.quad_up(0, 0, size_x, size_z, 0.0)
    if obstacle:
        s.box(size_x * 0.4, 0.0, size_z * 0.35, size_x * 0.6, 1.0, size_z * 0.55)
    return s.arrays()


def _walls_with_door(s: Soup, axis: str, fixed: float, lo: float, hi: float, y0: float,
                     wall_h: float, wall_t: float, door_lo: float | None, door_w: float):
    """A wall along `axis` ('x': spans x in [lo,hi] at z=fixed; 'z': spans z at x=fixed)."""
    segs = [(lo, hi)] if door_lo is None else [(lo, door_lo), (door_lo + door_w, hi)]
    for a, b in segs:
        if b - a <= 1e-6:
            continue
        if axis == "x":
            s.box(a, y0, fixed - wall_t / 2, b, y0 + wall_h, fixed + wall_t / 2)
        else:
            s.box(fixed - wall_t / 2, y0, a, fixed + wall_t / 2, y0 + wall_h, b)


def rooms_floor(s: Soup, nx: int, nz: int, rng, *, room_w=5.0, room_d=4.0, door_w=1.0,
                wall_t=0.2, wall_h=2.5, y0=0.0, furniture=2, closed_rooms=0,
                holes=(), open_cells=()):
    """One storey: nx x nz rooms on a slab at height y0.

    holes: rectangles (x0,z0,x1,z1) cut out of the slab (stairwells).
    open_cells: set of (ix, iz, 'x'|'z') interior walls to omit entirely (atria).
    closed_rooms: that many random rooms get no doors at all (extra islands).
    """
    W, D = nx * room_w, nz * room_d
    # slab with rectangular holes: split into strips along z between hole boundaries
    rects = [(0.0, 0.0, W, D)]
    for hx0, hz0, hx1, hz1 in holes:
        nxt = []
        for x0, z0, x1, z1 in rects:
            if hx0 >= x1 or hx1 <= x0 or hz0 >= z1 or hz1 <= z0:
                nxt.append((x0, z0, x1, z1))
                continue
            if z0 < hz0:
                nxt.append((x0, z0, x1, hz0))
            if hz1 < z1:
                nxt.append((x0, hz1, x1, z1))
            zz0, zz1 = max(z0, hz0), min(z1, hz1)
            if x0 < hx0:
                nxt.append((x0, zz0, hx0, zz1))
            if hx1 < x1:
                nxt.append((hx1, zz0, x1, zz1))
        rects = nxt
    for x0, z0, x1, z1 in rects:
        s.quad_up(x0, z0, x1, z1, y0)

    closed = set()
    while len(closed) < min(closed_rooms, nx * nz):
        closed.add((int(rng.integers(nx)), int(rng.integers(nz))))
    open_cells = set(open_cells)

    # perimeter
    _walls_with_door(s, "x", 0.0, 0.0, W, y0, wall_h, wall_t, None, door_w)
    _walls_with_door(s, "x", D, 0.0, W, y0, wall_h, wall_t, None, door_w)
    _walls_with_door(s, "z", 0.0, 0.0, D, y0, wall_h, wall_t, None, door_w)
    _walls_with_door(s, "z", W, 0.0, D, y0, wall_h, wall_t, None, door_w)
    # interior walls between (ix,iz)-(ix+1,iz): at x=(ix+1)*room_w spanning the room depth
    for iz in range(nz):
        for ix in range(nx - 1):
            if (ix, iz, "x") in open_cells:
                continue
            lo, hi = iz * room_d, (iz + 1) * room_d
            door = None
            if (ix, iz) not in closed and (ix + 1, iz) not in closed:
                door = lo + wall_t + float(rng.uniform(0.3, room_d - door_w - 2 * wall_t - 0.3))
            _walls_with_door(s, "z", (ix + 1) * room_w, lo, hi, y0, wall_h, wall_t, door, door_w)
    for iz in range(nz - 1):
        for ix in range(nx):
            if (ix, iz, "z") in open_cells:
                continue
            lo, hi = ix * room_w, (ix + 1) * room_w
            door = None
            if (ix, iz) not in closed and (ix, iz + 1) not in closed:
                door = lo + wall_t + float(rng.uniform(0.3, room_w - door_w - 2 * wall_t - 0.3))
            _walls_with_door(s, "x", (iz + 1) * room_d, lo, hi, y0, wall_h, wall_t, door, door_w)
    # furniture: boxes (obstacles) and low platforms (climbable, <= 0.15 m)
    for iz in range(nz):
        for ix in range(nx):
            for _ in range(furniture):
                fw, fd = float(rng.uniform(0.4, 1.2)), float(rng.uniform(0.4, 1.2))
                fx = ix * room_w + float(rng.uniform(0.6, room_w - 0.6 - fw))
                fz = iz * room_d + float(rng.uniform(0.6, room_d - 0.6 - fd))
                if any(fx < hx1 and fx + fw > hx0 and fz < hz1 and fz + fd > hz0
                       for hx0, hz0, hx1, hz1 in holes):
                    continue
                fh = 0.15 if rng.uniform() < 0.2 else float(rng.uniform(0.5, 1.1))
                s.box(fx, y0, fz, fx + fw, y0 + fh, fz + fd)
    return W, D


def apartment(seed: int = 0):
    """C2: ~120 m^2 apartment, 3 x 2 rooms of 5 x 4 m with 1 m doors + furniture."""
    rng = np.random.default_rng(seed)
    s = Soup()
    rooms_floor(s, 3, 2, rng, furniture=2)
    return s.arrays()


def multi_room(nx: int = 10, nz: int = 10, seed: int = 0, closed_rooms: int = 2,
               furniture: int = 2):
    """C3: multi-room single storey (default 10 x 10 rooms, ~2000 m^2)."""
    rng = np.random.default_rng(seed)
    s = Soup()
    rooms_floor(s, nx, nz, rng, furniture=furniture, closed_rooms=closed_rooms)
    return s.arrays()


def building(nx: int = 32, nz: int = 32, floors: int = 4, seed: int = 0, storey_h: float = 3.0,
             ramps_per_floor: int = 6, closed_rooms: int = 8, furniture: int = 2,
             room_w: float = 5.0, room_d: float = 4.0):
    """C4/C5: multi-floor building; storeys joined by ramps through slab holes.

    Each ramp lives in a two-room atrium (the wall between rooms (ix,iz) and (ix+1,iz) is
    omitted on both storeys): 1.5 m wide, rising storey_h over 6 m along +x.  The top
    storey is left unconnected when ramps_per_floor == 0 -> extra islands.
    """
    rng = np.random.default_rng(seed)
    s = Soup()
    run = 6.0
    ramp_w = 1.5
    per_floor = []
    for f in range(floors - 1):
        cells = set()
        ramps = []
        tries = 0
        while len(ramps) < ramps_per_floor and tries < 1000:
            tries += 1
            ix, iz = int(rng.integers(nx - 1)), int(rng.integers(nz))
            # keep atria apart (also from the previous storey's, whose hole is in our slab)
            near = [(ix + dx, iz + dz) for dx in (-1, 0, 1, 2) for dz in (-1, 0, 1)]
            prev = per_floor[-1][0] if per_floor else set()
            if any(c in cells or c in prev for c in near):
                continue
            cells.update([(ix, iz), (ix + 1, iz)])
            xs, zs = ix * room_w, iz * room_d
            ramps.append((ix, iz, xs + 1.5, zs + 1.25, xs + 1.5 + run, zs + 1.25 + ramp_w))
        per_floor.append((cells, ramps))
    per_floor.append((set(), []))
    for f in range(floors):
        y0 = f * storey_h
        holes = []
        open_cells = set()
        if f > 0:
            for ix, iz, x0, z0, x1, z1 in per_floor[f - 1][1]:
                holes.append((x0 - 0.3, z0 - 0.3, x1, z1 + 0.3))
                open_cells.add((ix, iz, "x"))
        for ix, iz, x0, z0, x1, z1 in per_floor[f][1]:
            open_cells.add((ix, iz, "x"))
        rooms_floor(s, nx, nz, rng, room_w=room_w, room_d=room_d, y0=y0, furniture=furniture,
                    closed_rooms=closed_rooms, holes=holes, open_cells=open_cells,
                    wall_h=min(2.5, storey_h - 0.3))
        for ix, iz, x0, z0, x1, z1 in per_floor[f][1]:
            s.ramp(x0, z0, x1, z1, y0, y0 + storey_h)
    return s.arrays()

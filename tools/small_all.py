"""A small run of every batched entry point against the oracle (for compute-sanitizer: memcheck / racecheck).
usage: python tools/small_all.py [scene] [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import habitat_sim_b200  # noqa: F401
from habitat_sim_b200.nav import PathFinder
from oracle.ref import RefPathFinder
from workloads.scenes import NavMeshGeom, navmesh_bytes, pointnav_pairs

name = sys.argv[1] if len(sys.argv) > 1 else "c4_building"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192


def beq(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return ((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all()


img = navmesh_bytes(name)
pf = PathFinder(0); assert pf.load_nav_mesh_bytes(img)
ref = RefPathFinder(); assert ref.load_bytes(img)
st, en = pointnav_pairs(NavMeshGeom(img), n, 3, jitter=0.3)
pts, refs, isl = pf.snap_points(st)
wp, wr, wi = ref.snap_batch(st, 8)
print("snap pipeline", beq(pts, wp) and (np.asarray(refs) == wr).all() and (np.asarray(isl) == wi).all())
hp, hn, hd = ref.obstacle_batch(st, 2.0, 8)
gp, gn, gd = pf.closest_obstacle_surface_points(st, 2.0)
print("wall distance", beq(gd, hd) and beq(gp, hp) and beq(gn, hn))
d = pf.find_paths(st, en)["geodesic_distance"]
print("find_path", beq(d, ref.find_path_raw_batch(st, en, nthreads=8)["dist"]))
t = st + np.float32([0.25, 0, 0])
print("try_step", beq(pf.try_steps(st, t), ref.try_step_batch(st, t, True, 8)))
r = pf.random_navigable_points(n, seed=7, query0=0)
print("random", beq(np.asarray(r[0]), ref.random_points(n, 10, np.full(n, -1, np.int32), mode=1, seed=7, query0=0)[0]))

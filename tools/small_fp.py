"""Debug helper (GPU box): a small find_path batch checked against the oracle (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import habitat_sim_b200  # noqa
from habitat_sim_b200.nav import PathFinder
from oracle.ref import RefPathFinder
from workloads.scenes import navmesh_bytes, NavMeshGeom, uniform_pairs
scene = sys.argv[1] if len(sys.argv) > 1 else "t_building"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
img = navmesh_bytes(scene)
pf = PathFinder(0); pf.load_nav_mesh_bytes(img)
ref = RefPathFinder(); ref.load_bytes(img)
st, en = uniform_pairs(NavMeshGeom(img), n, 5)
for exact in (False, True):
    got = pf.find_paths(st, en, max_points=16, corridors=True, exact_status=exact)
    want = ref.find_path_raw_batch(st, en, max_pts=16, nthreads=8)
    same = (got["geodesic_distance"].view(np.uint32) == want["dist"].view(np.uint32))
    print(scene, "exact" if exact else "fast", "n", n, "mismatch", int((~same).sum()), "found", float(np.isfinite(want["dist"]).mean()), flush=True)

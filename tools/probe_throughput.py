import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import habitat_sim_b200
from habitat_sim_b200.nav import PathFinder
from workloads.scenes import navmesh_bytes, NavMeshGeom, pointnav_pairs
for scene, n in [("c2_apartment", 200000), ("c3_multiroom", 200000), ("c4_building", 200000)]:
    img = navmesh_bytes(scene)
    pf = PathFinder(0); pf.load_nav_mesh_bytes(img)
    geom = NavMeshGeom(img)
    st, en = pointnav_pairs(geom, n, 1)
    s = torch.from_numpy(st).cuda(); e = torch.from_numpy(en).cuda()
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        r = pf.find_paths(s, e)
        torch.cuda.synchronize(); dt = time.time() - t0
    d = r["geodesic_distance"].cpu().numpy()
    print(scene, "find_path dev q/s", n/dt, "found", np.isfinite(d).mean(), flush=True)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        sp = pf.snap_points(s)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(scene, "snap dev q/s", n/dt, flush=True)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        sp = pf.try_steps(s, s + 0.2)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(scene, "try_step dev q/s", n/dt, flush=True)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        sp = pf.distances_to_closest_obstacle(s)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(scene, "obstacle dev q/s", n/dt, flush=True)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        sp = pf.random_navigable_points(n, device_output=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(scene, "random dev q/s", n/dt, flush=True)

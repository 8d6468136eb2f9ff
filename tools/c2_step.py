"""One C2 PointNav step sequence (1024 envs: try_step + geodesic distance), for a launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/c2_step.py [envs] [steps]
HBN_LANE_SPREAD=1 spreads a batch smaller than the grid over more warps (fewer queries per warp)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import habitat_sim_b200  # noqa
from habitat_sim_b200.nav import PathFinder
from workloads.scenes import NavMeshGeom, navmesh_bytes, step_targets, uniform_pairs
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
img = navmesh_bytes("c2_apartment")
pf = PathFinder(0); pf.load_nav_mesh_bytes(img)
pos0, goal = uniform_pairs(NavMeshGeom(img), envs, 3, jitter=0.0)
dev = torch.device("cuda", 0)
p = torch.from_numpy(pf.snap_points(pos0)[0]).to(dev); g = torch.from_numpy(goal).to(dev)
disp = [torch.from_numpy(step_targets(pos0, 100 + k) - pos0).to(dev) for k in range(steps)]
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(steps):
        p = pf.try_steps(p, p + disp[k])
        d = pf.find_paths(p, g)["geodesic_distance"]
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print(f"rep {rep}: {steps} steps, host enqueue {1e6*t_cpu/steps:.0f} us/step, wall {1e6*t_all/steps:.0f} us/step", flush=True)

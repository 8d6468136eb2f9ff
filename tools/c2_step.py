"""One C2 PointNav step sequence (1024 envs: hbn_env_step = try_step + geodesic distance), for a launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/c2_step.py [envs] [steps] [mode]
mode: fused (default: hbn_env_step_dev on torch tensors), graph (hbn_env_step, host buffers, CUDA graph replay),
separate (hbn_try_step_dev + hbn_find_path_dev: round 1's sequence)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import habitat_sim_b200  # noqa
from habitat_sim_b200.nav import PathFinder
from workloads.scenes import NavMeshGeom, navmesh_bytes, step_targets, uniform_pairs
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "fused"
img = navmesh_bytes("c2_apartment")
pf = PathFinder(0); pf.load_nav_mesh_bytes(img)
pos0, goal = uniform_pairs(NavMeshGeom(img), envs, 3, jitter=0.0)
dev = torch.device("cuda", 0)
p0 = pf.snap_points(pos0)[0]
disp_h = [step_targets(pos0, 100 + k) - pos0 for k in range(steps)]
p = torch.from_numpy(p0).to(dev); g = torch.from_numpy(goal).to(dev)
disp = [torch.from_numpy(d).to(dev) for d in disp_h]
for rep in range(3):
    l0 = pf.launch_count
    ph = p0.copy()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(steps):
        if mode == "fused":
            p, d = pf.env_steps(p, p + disp[k], g)
        elif mode == "graph":
            ph, dh = pf.env_steps(ph, ph + disp_h[k], goal)
        else:
            p = pf.try_steps(p, p + disp[k])
            d = pf.find_paths(p, g)["geodesic_distance"]
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print(f"{mode} rep {rep}: {envs} envs, {steps} steps, host enqueue {1e6*t_cpu/steps:.0f} us/step, wall {1e6*t_all/steps:.0f} us/step, "
          f"{(pf.launch_count - l0) / steps:.1f} kernels/step", flush=True)

"""Throughput of the other BASELINE.json configs (the bench line is C4 find_path, bench.py):
C2 PointNav step (try_step + geodesic distance, 1024 envs), C3 multi-goal (4096 x 64),
C4 snap_point, C5 wall distance + island-restricted random points, each beside the reference
Detour path on the host cores (bounded sample) and checked bit for bit on that sample.

    python tools/bench_configs.py [--scale 1.0] > profiles/rNN_configs.json
Device-resident inputs, CUDA events, 3 warm-up + 3 timed repetitions; one JSON object per line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def beq(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


def gpu_time(fn, reps=3, warm=3):
    import torch
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        out = fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e-3 / reps, out


def cpu_time(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    import torch
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import PathFinder
    from oracle.ref import RefPathFinder
    from workloads.scenes import NavMeshGeom, navmesh_bytes, pointnav_pairs, step_targets, uniform_pairs

    dev = torch.device("cuda", 0)
    threads = os.cpu_count() or 1
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    only = os.environ.get("BENCH_ONLY", "")  # e.g. BENCH_ONLY=C3

    def want(tag):
        return not only or tag in only.split(",")

    def load(name):
        img = navmesh_bytes(name)
        pf = PathFinder(0)
        assert pf.load_nav_mesh_bytes(img)
        ref = RefPathFinder()
        assert ref.load_bytes(img)
        return img, pf, ref, NavMeshGeom(img)

    # ---- C2: 1024 envs (and more), per step try_step + geodesic distance to the goal ---------
    img, pf, ref, geom = load("c2_apartment")
    for envs in ((1024, 65536) if want("C2") else ()):
        steps = max(10, int((100 if envs == 1024 else 20) * args.scale))
        pos0, goal = uniform_pairs(geom, envs, 3, jitter=0.0)
        pos0 = ref.snap_batch(pos0, threads)[0]
        tgts = [step_targets(pos0, 100 + k) - pos0 for k in range(steps)]  # displacement per step
        goal_d = T(goal)
        disp_d = [T(t) for t in tgts]

        def c2_gpu():
            p = T(pos0)
            d = None
            for k in range(steps):
                p = pf.try_steps(p, p + disp_d[k])
                d = pf.find_paths(p, goal_d)["geodesic_distance"]
            return p, d

        dt, (p_g, d_g) = gpu_time(c2_gpu, reps=2, warm=1)

        def c2_cpu():
            p = pos0.copy()
            d = None
            for k in range(steps):
                p = ref.try_step_batch(p, p + tgts[k], True, threads)
                d = ref.find_path_batch(p, goal, 0, threads)[0]
            return p, d

        dtc, (p_c, d_c) = cpu_time(c2_cpu)
        emit(config=f"C2 PointNav step: {envs} envs on c2_apartment, try_step + geodesic_distance per step, "
                    f"{steps} dependent steps (one launch sequence per step)",
             unit="env-steps/s", b200=envs * steps / dt, reference_cpu=envs * steps / dtc, cores=threads,
             us_per_step_b200=1e6 * dt / steps,
             bit_exact=beq(p_g.cpu().numpy(), p_c) and beq(d_g.cpu().numpy(), d_c))

    if want("C3"):
        # ---- C3: multi-goal, 4096 starts x 64 goals -------------------------------------------
        img, pf, ref, geom = load("c3_multiroom")
        ns, g = int(4096 * args.scale), 64
        rng = np.random.default_rng(5)
        st = geom.sample(ns, rng)
        en = geom.sample(ns * g, rng).reshape(ns, g, 3)
        st_d, en_d = T(st), T(en)
        dt, out = gpu_time(lambda: pf.find_paths_multigoal(st_d, en_d))
        m = min(ns, 64 * threads)
        dtc, outc = cpu_time(lambda: ref.find_path_multigoal_batch(st[:m], en[:m], 0, threads))
        gd = out["geodesic_distance"].cpu().numpy() if hasattr(out["geodesic_distance"], "cpu") else out["geodesic_distance"]
        gi = out["closest_end_point_index"]
        gi = gi.cpu().numpy() if hasattr(gi, "cpu") else gi
        emit(config=f"C3 MultiGoalShortestPath: {ns} starts x {g} goals on c3_multiroom (fresh objects)",
             unit="starts/s", b200=ns / dt, reference_cpu=m / dtc, cores=threads, cpu_sample=m,
             pair_searches_per_s_b200=ns * g / dt,
             bit_exact=beq(gd[:m], outc[0]) and bool((gi[:m] == outc[1]).all()))

    if not (want("C4") or want("C5")):
        return
    # ---- C4: snap_point --------------------------------------------------------------------
    img, pf, ref, geom = load("c4_building")
    n = int(1_000_000 * args.scale)
    pts, _ = pointnav_pairs(geom, n, 7, jitter=0.3)
    pts_d = T(pts)
    dt, out = gpu_time(lambda: pf.snap_points(pts_d))
    m = min(n, 20000 * threads)
    dtc, outc = cpu_time(lambda: ref.snap_batch(pts[:m], threads))
    sp = out[0] if isinstance(out, (tuple, list)) else out["points"]
    emit(config=f"C4 snap_point: {n} points (navigable + N(0,0.3) jitter) on c4_building", unit="points/s",
         b200=n / dt, reference_cpu=m / dtc, cores=threads, cpu_sample=m,
         bit_exact=beq(sp[:m].cpu().numpy(), outc[0]))

    # ---- C5: wall distance + island-restricted random points --------------------------------
    n5 = int(16_000_000 * args.scale)
    chunk = 2_000_000
    pts5, _ = pointnav_pairs(geom, min(n5, chunk), 9, jitter=0.05)
    p5_d = T(pts5)
    reps5 = max(1, n5 // len(pts5))

    def c5_wall():
        d = None
        for _ in range(reps5):
            d = pf.distances_to_closest_obstacle(p5_d)
        return d

    dt, dwall = gpu_time(c5_wall, reps=1, warm=1)
    m = min(len(pts5), 20000 * threads)
    dtc, outc = cpu_time(lambda: ref.obstacle_batch(pts5[:m], 2.0, threads))
    emit(config=f"C5 distance_to_closest_obstacle: {reps5 * len(pts5)} queries on c4_building", unit="queries/s",
         b200=reps5 * len(pts5) / dt, reference_cpu=m / dtc, cores=threads, cpu_sample=m,
         bit_exact=beq(dwall[:m].cpu().numpy(), outc[2]))

    isl = np.random.default_rng(11).integers(0, pf.num_islands, chunk).astype(np.int32)
    # islands with area only
    areas = np.array([pf.island_area(i) for i in range(pf.num_islands)])
    good = np.nonzero(areas > 0)[0].astype(np.int32)
    isl = good[isl % len(good)]
    isl_d = T(isl)

    def c5_rand():
        o = None
        for r in range(reps5):
            o = pf.random_navigable_points(chunk, 10, isl_d, seed=5, query0=r * chunk, device_output=True)
        return o

    dt, orand = gpu_time(c5_rand, reps=1, warm=1)
    m = min(chunk, 500 * threads)
    q0 = (reps5 - 1) * chunk
    dtc, outc = cpu_time(lambda: ref.random_points(m, 10, isl[:m], mode=1, seed=5, query0=q0))
    emit(config=f"C5 get_random_navigable_point (island restricted): {reps5 * chunk} samples on c4_building",
         unit="samples/s", b200=reps5 * chunk / dt, reference_cpu=m / dtc, cores=1, cpu_sample=m,
         bit_exact=beq(orand[0][:m].cpu().numpy(), outc[0]))


if __name__ == "__main__":
    main()

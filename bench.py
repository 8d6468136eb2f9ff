#!/usr/bin/env python
"""bench.py -- geodesic find_path queries/s on the C4 workload (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # reference Detour on the host cores

A step = one pass of the hot path (projectToPoly x2 -> A* -> funnel -> path length) over one
batch of `--queries` (default 1,000,000) seeded find_path queries per GPU on the procedural
multi-floor tiled navmesh `c4_building` (~54.6 k polys, 99 tiles), query mix of SURVEY.md §8d:
50 % PointNav-like same-storey pairs within 15 m, 50 % uniform pairs.  Under torchrun every
rank holds a replica of the navmesh and its own independent shard of queries (weak scaling,
no data-path collective; torch.distributed is only used for the barrier and the max-over-ranks
of the timings).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, inputs in
HBM), `e2e` = the same through the C ABI's host-buffer entry point hbn_find_path (H2D + D2H
inside the timed region), `roofline` for the dominant kernel (k_astar_lane, 93 % of a step; timed
together with the classify / funnel kernels around it), `cpu_baseline` =
the oracle (reference Detour compiled from /root/reference + restated PathFinder layer) on all
host cores over a bounded sample of the same queries.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENE = "c4_building"
METRIC = "find_path_queries_per_sec"
UNIT = "queries/s"
WORKLOAD = (f"C4 {SCENE}: find_path (geodesic distance), 50% same-storey pairs within 15 m + 50% uniform "
            "pairs, default NavMeshSettings, tiled 256-cell navmesh")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_queries(n: int, seed: int):
    from workloads.scenes import NavMeshGeom, navmesh_bytes, pointnav_pairs
    image = navmesh_bytes(SCENE)
    # sweeps that call bench.py many times (tools/sweep_fp.py) keep the generated pairs on disk
    cache = os.environ.get("HBN_QUERY_CACHE")
    path = os.path.join(cache, f"{SCENE}_{n}_{seed}.npz") if cache else None
    if path and os.path.exists(path):
        z = np.load(path)
        return image, z["st"], z["en"]
    geom = NavMeshGeom(image)
    st, en = pointnav_pairs(geom, n, seed)
    if path:
        os.makedirs(cache, exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp.npz"
        np.savez(tmp, st=st, en=en)
        os.replace(tmp, path)
    return image, st, en


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.index >= 0:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, name in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_arm(st, en, sample: int, threads: int, steps: int, warmup: int):
    """Reference Detour find_path on the host cores; returns (q/s, seconds, distances)."""
    from oracle.ref import RefPathFinder
    from workloads.scenes import navmesh_bytes
    ref = RefPathFinder()
    assert ref.load_bytes(navmesh_bytes(SCENE))
    s, e = st[:sample], en[:sample]
    for _ in range(warmup):
        ref.find_path_batch(s[: max(1, sample // 8)], e[: max(1, sample // 8)], 0, threads)
    t0 = time.perf_counter()
    d = None
    for _ in range(steps):
        d = ref.find_path_batch(s, e, 0, threads)[0]
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt, d


def reference_main(args, rank, world):
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = args.cpu_sample or 2000 * threads
    _, st, en = make_queries(sample, 1000)
    qps, dt, d = cpu_arm(st, en, sample, threads, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_step": sample, "host_threads": threads},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{sample} queries/step x {args.steps} steps; reference Detour "
                                   "(compiled from /root/reference by oracle/Makefile) under the restated "
                                   "PathFinder layer, one dtNavMeshQuery per thread"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "found_fraction": float(np.isfinite(d).mean()),
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=1_000_000, help="find_path queries per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return reference_main(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 arm has no CPU path"}), flush=True)
        return 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import PathFinder

    n = args.queries
    image, st, en = make_queries(n, 1000 + rank)
    pf = PathFinder(local)
    assert pf.load_nav_mesh_bytes(image)
    info = pf.mesh_info()
    s_dev = torch.from_numpy(st).to(dev)
    e_dev = torch.from_numpy(en).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- device-resident arm: `value` -------------------------------------------------
    for _ in range(args.warmup):
        r = pf.find_paths(s_dev, e_dev)
    torch.cuda.synchronize(dev)
    pf.set_profiling(True)
    pf.phase_times()
    launches0 = pf.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    # rank 0 samples its GPU's clocks (one nvidia-smi poller per node is enough)
    with ClockSampler(local if rank == 0 else -1) as clocks:
        for k in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            ev[k][0].record()
            r = pf.find_paths(s_dev, e_dev)
            ev[k][1].record()
        barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = pf.launch_count - launches0
    phases = pf.phase_times()
    pf.set_profiling(False)
    total_ms = max_over_ranks(sum(step_ms))
    value = n * world * args.steps / (total_ms * 1e-3)
    d_dev = r["geodesic_distance"].cpu().numpy()
    found = float(np.isfinite(d_dev).mean())

    # ---- algorithmic bytes of the dominant kernel (untimed pass with work counters) -----
    pf.work_counters(reset=True)
    pf.find_paths(s_dev, e_dev, count_work=True)
    wc = pf.work_counters(reset=True)
    # dtPoly 32 B, dtLink 12 B, neighbour poly 32 B + portal verts 24 B;
    # kernel I/O per query: requested + snapped start/end (48 B), 2 poly ids, 1 distance
    b_astar = 32 * wc["expanded"] + 12 * wc["links"] + 56 * wc["neighbours"]
    b_funnel = 56 * wc["corridor"] + 12 * wc["corridor_links"]
    b_io = (48 + 8 + 4) * n
    alg_bytes = b_astar + b_funnel + b_io
    path_ms = phases["path_ms"] / max(1, phases["calls"])
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (path_ms * 1e-3) / 1e9 if path_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_query") * n  # ncu capture, scaled to this launch
        except Exception:
            traffic = None

    # ---- end-to-end arm through the C ABI with host buffers: `e2e` ----------------------
    for _ in range(2):
        pf.find_paths(st, en)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rh = pf.find_paths(st, en)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = n * world * args.steps / e2e_s
    assert np.array_equal(rh["geodesic_distance"].view(np.uint32), d_dev.view(np.uint32))

    # ---- CPU baseline (rank 0, N = 1 only) + parity of the same sample ------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = min(n, args.cpu_sample or 4000 * threads)
        qps, dt, d_ref = cpu_arm(st, en, sample, threads, 1, 1)
        cpu = {"value": qps, "unit": UNIT, "cores": threads, "kind": "reference",
               "sample": f"first {sample} of the step's {n} queries, {dt:.1f} s; reference Detour (compiled "
                         "from /root/reference by oracle/Makefile) under the restated PathFinder layer, "
                         "one dtNavMeshQuery per thread"}
        same = (d_ref.view(np.uint32) == d_dev[:sample].view(np.uint32))
        parity = {"queries": int(sample), "bit_exact_distances": int(same.sum()),
                  "mismatches": int((~same).sum())}

    tot_launch = sum_over_ranks(float(launches))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "queries_per_gpu_per_step": n, "num_polys": info["num_polys"],
                       "num_tiles": info["num_tiles"], "navmesh_device_bytes": info["device_bytes"],
                       "sharding": f"{world} independent query shards, replicated navmesh, no collective",
                       "l2": "flushed between timed iterations (256 MiB memset)"},
            "found_fraction": found,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(st.nbytes + en.nbytes),
                    "d2h_bytes_per_step": int(4 * n)},
            "gpu_launches": int(tot_launch),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_astar_lane (+ k_fp_classify/scatter/funnel: the find_path phase after the snaps)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "kernel_ms_per_step": path_ms,
                         "snap_ms_per_step": phases["snap_ms"] / max(1, phases["calls"]),
                         "work": wc},
            "cpu_baseline": cpu,
            "parity_sample": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

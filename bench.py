#!/usr/bin/env python
"""bench.py -- geodesic find_path queries/s on the C4 workload (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's PathFinder on the host cores
    python bench.py --config {c4,c2,c3,c4snap,c5wall,c5rand}  # the other BASELINE configs, same contract line

Default (`--config c4`): a step = one pass of the hot path (projectToPoly x2 -> A* -> funnel -> path
length) over one batch of `--queries` (default 1,000,000) seeded find_path queries per GPU on the
procedural multi-floor tiled navmesh `c4_building` (~54.6 k polys, 99 tiles), query mix of SURVEY.md 8d:
50 % PointNav-like same-storey pairs within 15 m, 50 % uniform pairs.  Under torchrun every rank holds a
replica of the navmesh and its own independent shard of queries (`value`: weak scaling, no data-path
collective; torch.distributed is only used for the barrier and the max-over-ranks of the timings).  The
same run also measures BASELINE config 4 as written -- ONE batch of `--queries` cut into N slices, one
per rank (`strong_scaling` in the JSON line).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, inputs in HBM), `e2e` =
the same through the C ABI's host-buffer entry point (H2D + D2H inside the timed region), `roofline` for
the dominant kernel (k_astar_lane, 93 % of a step; timed together with the classify / funnel kernels
around it), `cpu_baseline` = the reference's own PathFinder.cpp + Detour (oracle/_ref/libhbn_ref.so,
compiled from /root/reference by oracle/Makefile) on all host cores over a bounded sample of the same
queries, one PathFinder per thread, dynamic work distribution.
"""
from __future__ import annotations

import argparse
import hashlib
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

KERNEL_SOURCES = ("hbn_astar_lane.h", "hbn_astar_lane.cuh", "hbn_findpath.cuh", "hbn_types.h")


def kernel_sources_sha() -> str:
    """identifies the search kernel a profile was taken from (profiles/roofline_traffic.json)"""
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "habitat-sim_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def beq(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


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


# =====================================================================================================
# Workloads.  Each one: host inputs (seeded), a device-resident step, a host-buffer step, the CPU arm on
# a sample, and the parity of that sample.
# =====================================================================================================
class Workload:
    scene = "c4_building"
    metric = unit = text = ""
    dtype = "f32"

    def __init__(self, args, rank):
        self.args, self.rank = args, rank
        from workloads.scenes import NavMeshGeom, navmesh_bytes
        self.image = navmesh_bytes(self.scene)
        self.geom = NavMeshGeom(self.image)

    def cache(self, tag, make):
        """generated inputs kept on disk for sweeps that call bench.py many times (HBN_QUERY_CACHE)"""
        d = os.environ.get("HBN_QUERY_CACHE")
        path = os.path.join(d, f"{tag}.npz") if d else None
        if path and os.path.exists(path):
            z = np.load(path)
            return [z[k] for k in sorted(z.files)]
        arrs = make()
        if path:
            os.makedirs(d, exist_ok=True)
            tmp = f"{path}.{os.getpid()}.tmp.npz"
            np.savez(tmp, **{f"a{i}": a for i, a in enumerate(arrs)})
            os.replace(tmp, path)
        return arrs

    def reference(self):
        from oracle.ref import RefPathFinder
        ref = RefPathFinder()
        assert ref.load_bytes(self.image)
        return ref

    # to be provided: units (per step), to_device(dev), step_dev(pf), step_host(pf) -> result for parity,
    # h2d_bytes / d2h_bytes, cpu_step(ref, sample, threads) -> (units, result), parity(gpu, cpu, sample)
    config_extra: dict = {}

    def roofline(self, pf, phases, peak):
        return None


class FindPathC4(Workload):
    metric, unit = "find_path_queries_per_sec", "queries/s"
    text = ("C4 c4_building: find_path (geodesic distance), 50% same-storey pairs within 15 m + 50% uniform "
            "pairs, default NavMeshSettings, tiled 256-cell navmesh")

    def __init__(self, args, rank, seed_offset=0):
        super().__init__(args, rank)
        from workloads.scenes import pointnav_pairs
        self.n = args.queries
        seed = 1000 + rank + seed_offset
        self.st, self.en = self.cache(f"{self.scene}_{self.n}_{seed}", lambda: list(pointnav_pairs(self.geom, self.n, seed)))
        self.units = self.n
        self.h2d_bytes, self.d2h_bytes = int(self.st.nbytes + self.en.nbytes), 4 * self.n

    def to_device(self, torch, dev):
        self.s_dev, self.e_dev = torch.from_numpy(self.st).to(dev), torch.from_numpy(self.en).to(dev)

    def step_dev(self, pf):
        return pf.find_paths(self.s_dev, self.e_dev)["geodesic_distance"]

    def step_host(self, pf):
        return pf.find_paths(self.st, self.en)["geodesic_distance"]

    def cpu_step(self, ref, sample, threads):
        return sample, ref.find_path_batch(self.st[:sample], self.en[:sample], 0, threads)[0]

    def parity(self, gpu, cpu, sample):
        same = beq(cpu, gpu[:sample])
        return {"queries": int(sample), "bit_exact_distances": int(same.sum()), "mismatches": int((~same).sum())}

    def roofline(self, pf, phases, peak):
        # algorithmic bytes of the dominant kernel from an untimed pass with work counters.
        # dtPoly 32 B, dtLink 12 B, neighbour poly 32 B + portal verts 24 B;
        # kernel I/O per query: requested + snapped start/end (48 B), 2 poly ids, 1 distance
        pf.work_counters(reset=True)
        pf.find_paths(self.s_dev, self.e_dev, count_work=True)
        wc = pf.work_counters(reset=True)
        b_astar = 32 * wc["expanded"] + 12 * wc["links"] + 56 * wc["neighbours"]
        b_funnel = 56 * wc["corridor"] + 12 * wc["corridor_links"]
        alg = b_astar + b_funnel + (48 + 8 + 4) * self.n
        path_ms = phases["path_ms"] / max(1, phases["calls"])
        achieved = alg / (path_ms * 1e-3) / 1e9 if path_ms > 0 else 0.0
        traffic, tnote = None, "no ncu capture of these kernel sources under profiles/"
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("kernel_sources_sha") == kernel_sources_sha():
                    traffic = tj["dram_bytes_per_query"] * self.n  # ncu capture of THIS kernel, scaled to this launch
                    tnote = tj.get("source")
                else:
                    tnote = "profiles/roofline_traffic.json was captured from other kernel sources: not reported"
            except Exception:
                pass
        return {"bound": "hbm", "kernel": "k_astar_lane (+ k_fp_classify/scatter/funnel: the find_path phase after the snaps)",
                "achieved": achieved, "peak": peak[0], "unit": "GB/s", "frac": achieved / peak[0], "traffic": traffic,
                "traffic_source": tnote, "peak_source": peak[1], "algorithmic_bytes_per_launch": int(alg),
                "kernel_ms_per_step": path_ms, "snap_ms_per_step": phases["snap_ms"] / max(1, phases["calls"]),
                "work": wc}


class SnapC4(Workload):
    metric, unit = "snap_point_queries_per_sec", "points/s"
    text = "C4 c4_building: snap_point, navigable points + N(0, 0.3) jitter"

    def __init__(self, args, rank):
        super().__init__(args, rank)
        from workloads.scenes import pointnav_pairs
        self.n = args.queries
        (self.pts,) = self.cache(f"{self.scene}_snap_{self.n}_{rank}", lambda: [pointnav_pairs(self.geom, self.n, 7 + rank, jitter=0.3)[0]])
        self.units = self.n
        self.h2d_bytes, self.d2h_bytes = int(self.pts.nbytes), 12 * self.n + 8 * self.n

    def to_device(self, torch, dev):
        self.p_dev = torch.from_numpy(self.pts).to(dev)

    def step_dev(self, pf):
        return pf.snap_points(self.p_dev)[0]

    def step_host(self, pf):
        return pf.snap_points(self.pts)[0]

    def cpu_step(self, ref, sample, threads):
        return sample, ref.snap_batch(self.pts[:sample], threads)[0]

    def parity(self, gpu, cpu, sample):
        same = beq(cpu, gpu[:sample]).all(axis=1)
        return {"queries": int(sample), "bit_exact_points": int(same.sum()), "mismatches": int((~same).sum())}


class WallC5(SnapC4):
    metric, unit = "distance_to_closest_obstacle_queries_per_sec", "queries/s"
    text = "C5 c4_building: distance_to_closest_obstacle(max_search_radius = 2.0), navigable points + N(0, 0.05) jitter"

    def __init__(self, args, rank):
        Workload.__init__(self, args, rank)
        from workloads.scenes import pointnav_pairs
        self.n = args.queries
        m = min(self.n, 2_000_000)  # seeded points are generated for 2 M queries and repeated up to the batch size
        (base,) = self.cache(f"{self.scene}_wall_{m}_{rank}", lambda: [pointnav_pairs(self.geom, m, 9 + rank, jitter=0.05)[0]])
        self.pts = np.ascontiguousarray(np.tile(base, ((self.n + m - 1) // m, 1))[: self.n])
        self.units = self.n
        self.h2d_bytes, self.d2h_bytes = int(self.pts.nbytes), 4 * self.n

    def step_dev(self, pf):
        return pf.distances_to_closest_obstacle(self.p_dev)

    def step_host(self, pf):
        return pf.distances_to_closest_obstacle(self.pts)

    def cpu_step(self, ref, sample, threads):
        return sample, ref.obstacle_batch(self.pts[:sample], 2.0, threads)[2]

    def parity(self, gpu, cpu, sample):
        same = beq(cpu, gpu[:sample])
        return {"queries": int(sample), "bit_exact_distances": int(same.sum()), "mismatches": int((~same).sum())}


class RandC5(Workload):
    metric, unit = "random_navigable_point_samples_per_sec", "samples/s"
    text = "C5 c4_building: island-restricted get_random_navigable_point, islands cycling, counter-based stream"

    def __init__(self, args, rank):
        super().__init__(args, rank)
        self.n = args.queries
        self.units = self.n
        self.q0 = rank * self.n
        self.h2d_bytes, self.d2h_bytes = 4 * self.n, 12 * self.n

    def to_device(self, torch, dev):
        self.torch, self.dev = torch, dev

    def _islands(self, pf):
        if not hasattr(self, "isl"):
            areas = np.array([pf.island_area(i) for i in range(pf.num_islands)])
            good = np.nonzero(areas > 0)[0].astype(np.int32)
            self.isl = good[np.arange(self.n) % len(good)]
            self.isl_dev = self.torch.from_numpy(self.isl).to(self.dev)
        return self.isl

    def step_dev(self, pf):
        self._islands(pf)
        return pf.random_navigable_points(self.n, 10, self.isl_dev, seed=5, query0=self.q0, device_output=True)[0]

    def step_host(self, pf):
        return pf.random_navigable_points(self.n, 10, self._islands(pf), seed=5, query0=self.q0)[0]

    def cpu_step(self, ref, sample, threads):  # the reference scans every poly per sample: one core, small sample
        return sample, ref.random_points(sample, 10, self.isl[:sample], mode=1, seed=5, query0=self.q0)[0]

    cpu_threads = 1

    def parity(self, gpu, cpu, sample):
        same = beq(cpu, gpu[:sample]).all(axis=1)
        return {"queries": int(sample), "bit_exact_points": int(same.sum()), "mismatches": int((~same).sum())}


class MultiGoalC3(Workload):
    scene = "c3_multiroom"
    metric, unit = "multigoal_find_path_starts_per_sec", "starts/s"
    text = "C3 c3_multiroom: find_path(MultiGoalShortestPath), 64 goals per start, fresh objects"

    def __init__(self, args, rank):
        super().__init__(args, rank)
        self.n, self.g = args.queries, 64
        rng = np.random.default_rng(5 + rank)
        self.st = self.geom.sample(self.n, rng)
        self.en = self.geom.sample(self.n * self.g, rng).reshape(self.n, self.g, 3)
        self.units = self.n
        self.h2d_bytes, self.d2h_bytes = int(self.st.nbytes + self.en.nbytes), 8 * self.n

    def to_device(self, torch, dev):
        self.s_dev, self.e_dev = torch.from_numpy(self.st).to(dev), torch.from_numpy(self.en).to(dev)

    def step_dev(self, pf):
        r = pf.find_paths_multigoal(self.s_dev, self.e_dev)
        return r["geodesic_distance"]

    def step_host(self, pf):
        return pf.find_paths_multigoal(self.st, self.en)["geodesic_distance"]

    def cpu_step(self, ref, sample, threads):
        return sample, ref.find_path_multigoal_batch(self.st[:sample], self.en[:sample], 0, threads)[0]

    def parity(self, gpu, cpu, sample):
        same = beq(cpu, gpu[:sample])
        return {"starts": int(sample), "bit_exact_distances": int(same.sum()), "mismatches": int((~same).sum())}


class EnvStepC2(Workload):
    scene = "c2_apartment"
    metric, unit = "pointnav_env_steps_per_sec", "env-steps/s"
    text = ("C2 c2_apartment: PointNav env step = try_step(p, p + 0.25 dir) then geodesic distance to the env's goal; "
            "a bench step = 10 DEPENDENT env steps of every env (hbn_env_step: one fused call per env step)")
    CHAIN = 10

    def __init__(self, args, rank):
        super().__init__(args, rank)
        from workloads.scenes import uniform_pairs
        self.n = args.queries
        pos0, self.goal = uniform_pairs(self.geom, self.n, 3 + rank, jitter=0.0)
        rng = np.random.default_rng(17 + rank)
        th = rng.uniform(0, 2 * np.pi, (self.CHAIN, self.n)).astype(np.float32)
        self.disp = (np.stack([np.cos(th), np.zeros_like(th), np.sin(th)], 2) * np.float32(0.25)).astype(np.float32)
        self.pos0 = pos0
        self.units = self.n * self.CHAIN
        self.h2d_bytes, self.d2h_bytes = 36 * self.n * self.CHAIN, 16 * self.n * self.CHAIN

    def to_device(self, torch, dev):
        self.torch = torch
        self.p_dev, self.g_dev = torch.from_numpy(self.pos0).to(dev), torch.from_numpy(self.goal).to(dev)
        self.d_dev = torch.from_numpy(self.disp).to(dev)

    def step_dev(self, pf):
        p, d = self.p_dev, None
        for k in range(self.CHAIN):
            p, d = pf.env_steps(p, p + self.d_dev[k], self.g_dev)
        return self.torch.cat([p, d[:, None]], 1)

    def step_host(self, pf):
        p, d = self.pos0, None
        for k in range(self.CHAIN):
            p, d = pf.env_steps(p, p + self.disp[k], self.goal)
        return np.concatenate([p, d[:, None]], 1)

    def cpu_step(self, ref, sample, threads):
        p, d = self.pos0[:sample], None
        for k in range(self.CHAIN):
            p = ref.try_step_batch(p, p + self.disp[k, :sample], True, threads)
            d = ref.find_path_batch(p, self.goal[:sample], 0, threads)[0]
        return sample * self.CHAIN, np.concatenate([p, d[:, None]], 1)

    def parity(self, gpu, cpu, sample):
        same = beq(cpu, gpu[:sample]).all(axis=1)
        return {"envs": int(sample), "dependent_steps": self.CHAIN, "bit_exact_envs": int(same.sum()),
                "mismatches": int((~same).sum())}


CONFIGS = {
    # name: (class, default units per GPU per step, default CPU sample)
    "c4": (FindPathC4, 1_000_000, 262_144),
    "c2": (EnvStepC2, 1024, 1024),
    "c3": (MultiGoalC3, 4096, 1024),
    "c4snap": (SnapC4, 1_000_000, 262_144),
    "c5wall": (WallC5, 16_000_000, 262_144),
    "c5rand": (RandC5, 16_000_000, 512),
}


# =====================================================================================================
def cpu_arm(w, sample, threads, steps, warmup):
    """the reference on the host cores: per-step seconds, units per step, last result"""
    ref = w.reference()
    for _ in range(max(1, warmup)):  # also creates the per-thread PathFinder clones
        w.cpu_step(ref, max(1, sample // 8), threads)
    times, res, units = [], None, 0
    for _ in range(steps):
        t0 = time.perf_counter()
        units, res = w.cpu_step(ref, sample, threads)
        times.append(time.perf_counter() - t0)
    return times, units, res, ref


def cpu_baseline_block(w, args, threads, times, units, sample, one_thread=True):
    best, total = min(times), sum(times)
    out = {"value": units / best, "unit": w.unit, "cores": threads, "kind": "reference",
           "mean_value": units * len(times) / total,
           "sample": f"{sample} of the step's inputs per pass, {len(times)} passes, best pass {best:.2f} s "
                     f"(all passes {total:.1f} s); the reference's own PathFinder.cpp + Detour compiled from "
                     "/root/reference by oracle/Makefile, one PathFinder per thread (clones of one navmesh image), "
                     "work taken in chunks from an atomic counter"}
    if one_thread and threads > 1:
        ref = w.reference()
        m = max(1, sample // (4 * threads))
        w.cpu_step(ref, max(1, m // 4), 1)
        t0 = time.perf_counter()
        u1, _ = w.cpu_step(ref, m, 1)
        out["one_thread_value"] = u1 / (time.perf_counter() - t0)
        out["one_thread_sample"] = m
    return out


def reference_main(args, rank, world):
    if rank != 0:
        return 0
    cls, _, dflt_sample = CONFIGS[args.config]
    w = cls(args, 0)
    threads = getattr(w, "cpu_threads", os.cpu_count() or 1)
    sample = min(args.cpu_sample or dflt_sample, args.queries)
    if hasattr(w, "_islands"):
        areas_ref = w.reference()
        good = np.array([i for i in range(areas_ref.num_islands) if areas_ref.navigable_area(i) > 0], np.int32)
        w.isl = good[np.arange(w.n) % len(good)]
    steps = max(3, args.steps)  # best of >= 3 passes
    times, units, res, _ = cpu_arm(w, sample, threads, steps, min(args.warmup, 1))
    cb = cpu_baseline_block(w, args, threads, times, units, sample)
    line = {
        "impl": "reference", "metric": w.metric, "value": cb["value"], "unit": w.unit, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * min(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w.dtype,
        "data": "synthetic",
        "config": {"workload": w.text, "units_per_step": units, "host_threads": threads,
                   "timing": "best of the timed passes (mean_value in cpu_baseline)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": w.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if res is not None and res.ndim == 1:
        line["found_fraction"] = float(np.isfinite(res).mean())
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--queries", type=int, default=0, help="units (queries / envs / starts) per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="units in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling measurement (c4 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cls, dflt_units, dflt_sample = CONFIGS[args.config]
    args.queries = args.queries or dflt_units

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

    def reduce(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if world > 1 else x

    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import PathFinder, shard_slices

    w = cls(args, rank)
    pf = PathFinder(local)
    assert pf.load_nav_mesh_bytes(w.image)
    info = pf.mesh_info()
    w.to_device(torch, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- device-resident arm: `value` -------------------------------------------------
    for _ in range(args.warmup):
        r = w.step_dev(pf)
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
            r = w.step_dev(pf)
            ev[k][1].record()
        barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = pf.launch_count - launches0
    phases = pf.phase_times()
    pf.set_profiling(False)
    total_ms = max_over_ranks(sum(step_ms))
    value = w.units * world * args.steps / (total_ms * 1e-3)
    r_dev = r.cpu().numpy()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = (float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)")
    else:
        peak = (6650.0, "fallback (B200_PROFILING.md)")
    roof = w.roofline(pf, phases, peak)
    if roof is None:  # the secondary configs: query I/O bytes only (the navmesh itself is L2 resident)
        io = w.h2d_bytes + w.d2h_bytes
        ach = io / (total_ms / args.steps * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "whole step", "achieved": ach, "peak": peak[0], "unit": "GB/s",
                "frac": ach / peak[0], "traffic": None, "peak_source": peak[1],
                "note": "algorithmic bytes = query inputs + outputs only; the navmesh (L2 resident) is not counted"}

    # ---- end-to-end arm through the C ABI with host buffers: `e2e` ----------------------
    for _ in range(2):
        w.step_host(pf)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rh = w.step_host(pf)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = w.units * world * args.steps / e2e_s
    assert beq(rh, r_dev).all(), "host-buffer and device-resident results differ"

    # ---- BASELINE config 4 as written: ONE batch of `queries` cut into `world` slices ---------
    strong = None
    if args.config == "c4" and not args.no_strong:
        ws = FindPathC4(args, 0, seed_offset=500)  # the same batch on every rank
        b, e = shard_slices(ws.n, world)[rank]
        s_h, e_h = np.ascontiguousarray(ws.st[b:e]), np.ascontiguousarray(ws.en[b:e])
        s_d, e_d = torch.from_numpy(s_h).to(dev), torch.from_numpy(e_h).to(dev)
        for _ in range(2):
            pf.find_paths(s_d, e_d)
            pf.find_paths(s_h, e_h)
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            flush.zero_()
            sev[k][0].record()
            pf.find_paths(s_d, e_d)
            sev[k][1].record()
        barrier()
        s_ms = max_over_ranks(sum(a.elapsed_time(b_) for a, b_ in sev))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pf.find_paths(s_h, e_h)
        torch.cuda.synchronize(dev)
        s_e2e = max_over_ranks(time.perf_counter() - t0)
        barrier()
        strong = {"scaling": "strong", "queries_total_per_step": ws.n, "queries_per_gpu_per_step": int(e - b),
                  "value": ws.n * args.steps / (s_ms * 1e-3), "e2e": ws.n * args.steps / s_e2e, "unit": w.unit,
                  "ms_per_step": s_ms / args.steps}

    # ---- CPU baseline (rank 0, N = 1 only) + parity of the same sample ------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = getattr(w, "cpu_threads", os.cpu_count() or 1)
        sample = min(w.units if not hasattr(w, "CHAIN") else w.n, args.cpu_sample or dflt_sample)
        times, units, r_cpu, _ = cpu_arm(w, sample, threads, 3, 1)
        cpu = cpu_baseline_block(w, args, threads, times, units, sample)
        parity = w.parity(r_dev, r_cpu, sample)

    tot_launch = sum_over_ranks(float(launches))
    if rank == 0:
        line = {
            "metric": w.metric, "value": value, "unit": w.unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": w.dtype, "data": "synthetic",
            "config": {"workload": w.text, "units_per_gpu_per_step": w.units, "num_polys": info["num_polys"],
                       "num_tiles": info["num_tiles"], "navmesh_device_bytes": info["device_bytes"],
                       "sharding": f"{world} independent query shards, replicated navmesh, no collective",
                       "l2": "flushed between timed iterations (256 MiB memset)"},
            "e2e": {"value": e2e, "unit": w.unit, "h2d_bytes_per_step": int(w.h2d_bytes),
                    "d2h_bytes_per_step": int(w.d2h_bytes)},
            "gpu_launches": int(tot_launch),
            "clocks": clocks.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity_sample": parity,
        }
        if r_dev.ndim == 1:
            line["found_fraction"] = float(np.isfinite(r_dev).mean())
        if strong:
            line["strong_scaling"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

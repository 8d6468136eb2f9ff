"""Quick GPU check of the opt-in k_astar_lane variants (HBN_LANE_CFG, HBN_LANE_SPREAD): device time of
the find_path phase per variant on a prepared C4 query set, distances compared bit for bit with the
shipped configuration's; then C2-size batches with and without lane spreading.  No torch import.
usage: python tools/variant_check.py [queries.npz]   (npz: st, en, c2s, c2e; made by --make)"""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_q", "variant_queries.npz")
if len(sys.argv) > 1 and sys.argv[1] == "--make":
    from workloads.scenes import NavMeshGeom, navmesh_bytes, pointnav_pairs, uniform_pairs
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    st, en = pointnav_pairs(NavMeshGeom(navmesh_bytes("c4_building")), n, 31)
    c2s, c2e = uniform_pairs(NavMeshGeom(navmesh_bytes("c2_apartment")), 1024, 3, jitter=0.0)
    np.savez(path, st=st, en=en, c2s=c2s, c2e=c2e)
    sys.exit(0)

import habitat_sim_b200  # noqa: F401,E402
from habitat_sim_b200.nav import PathFinder  # noqa: E402
from workloads.scenes import navmesh_bytes  # noqa: E402

q = np.load(path)
t00 = time.time()
out = open(os.path.join("gpurun_out", "variant_check.log"), "a") if os.path.isdir("gpurun_out") else None


def say(s):
    print(s, flush=True)
    if out:
        out.write(s + "\n")
        out.flush()


def as_u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


img = navmesh_bytes("c4_building")
base = None
for cfg in (os.environ.get("VARIANT_CFGS") or "0,0k0,24,17,22,20,21,18,19,23,1,8,10,13,15").split(","):
    # "50g40": cfg 50 with 40 directory groups per search (forces the overflow launch)
    os.environ["HBN_LANE_CFG"] = cfg.replace("k0", "").split("g")[0]
    os.environ["HBN_LANE_GROUP_CAP"] = cfg.split("g")[1] if "g" in cfg else "0"
    os.environ["HBN_KEY_ORDER"] = "0" if cfg.endswith("k0") else "1"  # k0: node keys in poly order
    pf = PathFinder(0)
    assert pf.load_nav_mesh_bytes(img)
    pf.set_profiling(True)
    pf.find_paths(q["st"], q["en"])
    pf.phase_times()
    for _ in range(2):
        d = pf.find_paths(q["st"], q["en"])["geodesic_distance"]
    t = pf.phase_times()
    if base is None:
        base = as_u32(d).copy()
    same = int((as_u32(d) == base).sum())
    say(f"C4 {len(d)} queries HBN_LANE_CFG={cfg}: path {t['path_ms'] / t['calls']:.2f} ms snap "
        f"{t['snap_ms'] / t['calls']:.2f} ms per call; distances equal to cfg 0: {same}/{len(d)}  [t={time.time() - t00:.1f}s]")
    del pf
    gc.collect()
os.environ.pop("HBN_LANE_CFG")
os.environ.pop("HBN_KEY_ORDER")

img2 = navmesh_bytes("c2_apartment")
base = None
for spread, snap_spread, dual in (("0", "0", "0"), ("1", "0", "0"), ("1", "1", "0"), ("1", "0", "1"), ("1", "1", "1")):
    os.environ["HBN_LANE_SPREAD"] = spread
    os.environ["HBN_SNAP_SPREAD"] = snap_spread
    os.environ["HBN_SNAP_DUAL"] = dual
    pf = PathFinder(0)
    assert pf.load_nav_mesh_bytes(img2)
    pf.set_profiling(True)
    pf.find_paths(q["c2s"], q["c2e"])
    pf.phase_times()
    t0 = time.perf_counter()
    for _ in range(20):
        d = pf.find_paths(q["c2s"], q["c2e"])["geodesic_distance"]
    wall = (time.perf_counter() - t0) / 20
    t = pf.phase_times()
    pf.try_steps(q["c2s"], q["c2e"])
    t0 = time.perf_counter()
    for _ in range(20):
        ts = pf.try_steps(q["c2s"], q["c2s"] + np.float32(0.25) * (q["c2e"] - q["c2s"]) /
                          np.linalg.norm(q["c2e"] - q["c2s"], axis=1, keepdims=True).astype(np.float32))
    wall_ts = (time.perf_counter() - t0) / 20
    if base is None:
        base = as_u32(d).copy()
        base_ts = as_u32(ts).copy()
    same = int((as_u32(d) == base).sum())
    same_ts = int((as_u32(ts) == base_ts).all(axis=1).sum())
    say(f"C2 {len(d)} queries HBN_LANE_SPREAD={spread} HBN_SNAP_SPREAD={snap_spread} HBN_SNAP_DUAL={dual}: path {1e3 * t['path_ms'] / t['calls']:.0f} us snap "
        f"{1e3 * t['snap_ms'] / t['calls']:.0f} us device, {1e6 * wall:.0f} us wall per call; equal: {same}/{len(d)}; try_steps {1e6 * wall_ts:.0f} us wall, equal {same_ts}/{len(ts)}"
        f"  [t={time.time() - t00:.1f}s]")
    del pf
    gc.collect()

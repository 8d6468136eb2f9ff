"""GPU parity of the tuning variants: HBN_LANE_CFG = other instantiations of k_astar_lane (shared
heap entries, warps per SM, L2 policies, code shape; opt-in, the steps that led to the shipped one),
lane spreading (a batch smaller than the grid uses fewer lanes per warp) and
the small-batch snap launches (one lane group per warp, two batches per launch); the last
three are on by default and can be switched off (HBN_LANE_SPREAD / HBN_SNAP_SPREAD /
HBN_SNAP_DUAL = 0).  Their host twins are checked on the CPU
(tests/test_host.py: heap fuzz, lock step with the shipped variant); here the device build must
give the reference's corridors, status words and distances too.  The file sorts last on
purpose: the shipped configuration is tested before the variants."""
import functools
import os

import numpy as np
import pytest

from conftest import beq, gpu_pathfinder, navmesh_image, query_points, ref_pathfinder

pytestmark = pytest.mark.gpu


def _check_against_reference(pf, ref, st, en):
    want = ref.find_path_raw_batch(st, en, max_pts=32, nthreads=8)
    got = pf.find_paths(st, en, max_points=32, corridors=True, exact_status=True)
    assert beq(got["geodesic_distance"], want["dist"]).all()
    astar_ran = (want["flags"] & 2) != 0
    assert (got["status"][astar_ran, 0] == want["astar_status"][astar_ran]).all()
    assert (got["num_corridor"][astar_ran] == want["num_polys"][astar_ran]).all()
    for i in np.nonzero(astar_ran)[0]:
        k = want["num_polys"][i]
        assert (got["corridor"][i, :k] == want["corridor"][i, :k]).all()
    fast = pf.find_paths(st, en)  # default mode: stop at pool exhaustion
    assert beq(fast["geodesic_distance"], want["dist"]).all()


@functools.lru_cache(maxsize=None)
def _pairs(name, n, seed):
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    if name == "c4_building":
        return pointnav_pairs(NavMeshGeom(navmesh_image(name)), n, seed)
    pts = query_points(name, 2 * n, seed)
    return pts[:n].copy(), pts[n:].copy()


# the instantiations kept next to the shipped one (hbn_capi.cu laneKernel): round 1's kernel, the L2-policy
# steps, 95 shared entries, two-warp blocks without / with the single replay loop body
_CFGS = ["1", "30", "31", "32", "34", "38", "60"]


@pytest.mark.parametrize("cfg", _CFGS)
def test_lane_kernel_configs(cfg, monkeypatch):
    monkeypatch.setenv("HBN_LANE_CFG", cfg)
    for name, n in (("t_building", 3000), ("c4_building", 6000)):
        st, en = _pairs(name, n, 21)
        _check_against_reference(gpu_pathfinder(name), ref_pathfinder(name), st, en)
    # a full grid: the shipped kernel's distances on 200 k queries
    st, en = _pairs("c4_building", 200_000, 23)
    d = gpu_pathfinder("c4_building").find_paths(st, en)["geodesic_distance"]
    monkeypatch.delenv("HBN_LANE_CFG")
    full = gpu_pathfinder("c4_building").find_paths(st, en)["geodesic_distance"]
    assert beq(d, full).all()


@pytest.mark.parametrize("spread,n", [("1", 1), ("1", 33), ("1", 1024), ("1", 5000), ("1", 100_000),
                                      ("0", 33), ("0", 5000)])
def test_lane_spread_small_batches(spread, n, monkeypatch):
    monkeypatch.setenv("HBN_LANE_SPREAD", spread)
    for name in ("c2_apartment", "c4_building"):
        st, en = _pairs(name, n, 25)
        pf = gpu_pathfinder(name)
        d = pf.find_paths(st, en)["geodesic_distance"]
        m = min(n, 5000)
        want = ref_pathfinder(name).find_path_batch(st[:m], en[:m], 0, 8)[0]
        assert beq(d[:m], want).all()
        if n <= 5000:
            _check_against_reference(pf, ref_pathfinder(name), st, en)


@pytest.mark.parametrize("spread,dual", [("1", "0"), ("0", "1"), ("1", "1"), ("0", "0")])
def test_snap_spread_small_batches(spread, dual, monkeypatch):
    """Small-batch snap launches, every combination of the two knobs (both on by default):
    HBN_SNAP_SPREAD launches k_snap<8> with one lane group per warp; HBN_SNAP_DUAL serves two
    independent snap batches (find_path's starts and ends, try_step's start and end) in one
    k_snap_dual launch."""
    from workloads.scenes import step_targets
    monkeypatch.setenv("HBN_SNAP_SPREAD", spread)
    monkeypatch.setenv("HBN_SNAP_DUAL", dual)
    for name in ("c2_apartment", "t_building"):
        pf, ref = gpu_pathfinder(name), ref_pathfinder(name)
        for n in (1, 7, 1024, 4000):
            pts = query_points(name, n, 41 + n)
            wp, wr, wi = ref.snap_batch(pts)
            gp, gr, gi = pf.snap_points(pts)
            assert (gr == wr).all() and (gi == wi).all() and beq(gp, wp).all()
            st = pts[: n // 2 + 1]
            en = query_points(name, len(st), 43 + n)
            want = ref.find_path_batch(st, en, 0, 8)[0]
            assert beq(pf.find_paths(st, en)["geodesic_distance"], want).all()
            s = wp[: len(st)].copy()  # try_step inputs as in test_gpu_parity.test_try_step
            s[np.isnan(s)] = 0
            t = step_targets(s, 5, 0.25)
            t[: len(s) // 4] = step_targets(s[: len(s) // 4], 6, 2.5)
            for sliding in (True, False):
                assert beq(pf.try_steps(s, t, sliding), ref.try_step_batch(s, t, sliding, 8)).all()


@pytest.mark.parametrize("name", ["c3_multiroom", "t_building"])
def test_out_of_the_ordinary_inputs(name):
    """The device twin of tests/test_host.py::test_hostemu_out_of_the_ordinary_inputs: try_step from
    unsnapped starts towards targets 0.05 to 30 m away, wall distance with radii from 0 to 1e6,
    snaps around the edges of the pick box (small batch: lane-group kernel; large: candidate lists)."""
    pf, ref = gpu_pathfinder(name), ref_pathfinder(name)
    n = 6000
    rng = np.random.default_rng(7)
    s = np.ascontiguousarray(query_points(name, n, 51))
    e = query_points(name, n, 52)
    d = e - s
    length = np.linalg.norm(d, axis=1, keepdims=True)
    length[length == 0] = 1
    step = rng.choice([0.05, 0.25, 1.0, 2.5, 6.0, 15.0, 30.0], size=(n, 1)).astype(np.float32)
    t = (s + d / length * step).astype(np.float32)
    t[::7] = e[::7]
    t[::11, 1] += rng.normal(0, 1.0, len(t[::11])).astype(np.float32)
    for sliding in (True, False):
        assert beq(pf.try_steps(s, t, sliding), ref.try_step_batch(s, t, sliding, 8)).all(), sliding
        m = 1000  # the small-batch launches
        assert beq(pf.try_steps(s[:m], t[:m], sliding), ref.try_step_batch(s[:m], t[:m], sliding, 8)).all()
    for r in (0.0, 0.05, 7.5, 1e6):
        hp, hn, hd = ref.obstacle_batch(s, r, 8)
        gp, gn, gd = pf.closest_obstacle_surface_points(s, r)
        assert beq(gd, hd).all() and beq(gp, hp).all() and beq(gn, hn).all(), r
    p2 = s.copy()
    p2[:, 1] += rng.choice([-4.2, -4.0, -3.99, -1.0, -0.2, 0.19, 0.2, 0.21, 1.0, 2.9, 3.99, 4.0, 4.01],
                           size=n).astype(np.float32)
    p2[:, 0] += rng.choice([0, 1.99, 2.0, 2.01, -2.0], size=n).astype(np.float32)
    wp, wr, wi = ref.snap_batch(p2, 8)
    for m in (n, 1000):
        gp, gr, gi = pf.snap_points(p2[:m])
        assert (gr == wr[:m]).all() and (gi == wi[:m]).all() and beq(gp, wp[:m]).all()

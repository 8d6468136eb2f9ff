"""GPU parity tests: the CUDA path, called through the C ABI of libhbn.so (ctypes ->
hbn_* host/device entry points), against the oracle on the same seeded inputs, against the
committed golden vectors, and through size-independent properties at larger sizes.
Bar: poly refs, corridors, island ids bit-exact; floats within 1e-5 relative (they are in
fact compared bit for bit, NaN == NaN)."""
import os

import numpy as np
import pytest

from conftest import beq, gpu_pathfinder, navmesh_image, query_points, ref_pathfinder

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SCENES = ["c1_room", "c2_apartment", "c3_multiroom", "t_building"]
RTOL = 1e-5  # north_star tolerance for float outputs


def close(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= RTOL * np.maximum(np.abs(a), np.abs(b)) + 1e-7
    return ok | both_nan | both_inf


@pytest.fixture(scope="module", params=SCENES)
def scene(request):
    name = request.param
    return name, gpu_pathfinder(name), ref_pathfinder(name)


def test_native_library_is_the_loaded_one(scene):
    name, pf, _ = scene
    import habitat_sim_b200
    maps = open("/proc/self/maps").read()
    assert habitat_sim_b200.library_path() in maps
    assert pf.launch_count >= 0


def test_islands_and_properties(scene):
    name, pf, ref = scene
    assert pf.num_islands == ref.num_islands
    for i in range(ref.num_islands):
        assert pf.island_radius(i) == ref.island_radius(i)
        assert pf.island_area(i) == ref.navigable_area(i)
    assert close(pf.navigable_area, ref.navigable_area())
    lo, hi = pf.get_bounds()
    rlo, rhi = ref.get_bounds()
    assert (lo == rlo).all() and (hi == rhi).all()


def test_snap_point(scene):
    name, pf, ref = scene
    pts = query_points(name, 20000, 31)
    lo, hi = ref.get_bounds()
    rng = np.random.default_rng(1)
    pts[:2000] = rng.uniform(lo - 3, hi + 3, (2000, 3)).astype(np.float32)
    pts[2000:2004] = np.nan
    pts[2004] = np.inf
    want_p, want_r, want_i = ref.snap_batch(pts, 8)
    got_p, got_r, got_i = pf.snap_points(pts)
    assert (got_r == want_r).all(), "nearest-poly refs must be bit-exact"
    assert (got_i == want_i).all(), "island ids must be bit-exact"
    assert close(got_p, want_p).all()
    assert beq(got_p, want_p).all()
    # island restricted (PathFinder.cpp:1725-1757)
    isl = rng.integers(0, ref.num_islands, 500).astype(np.int32)
    wp, wr = ref.snap_island_batch(pts[:500], isl)
    gp, gr, gi = pf.snap_points(pts[:500], isl)
    assert (gr == wr).all() and beq(gp, wp).all()
    # is_navigable
    assert (pf.are_navigable(pts[:5000]) == ref.is_navigable_batch(pts[:5000], 0.5, 8)).all()


def test_find_path(scene):
    name, pf, ref = scene
    n = 6000
    pts = query_points(name, 2 * n, 32)
    st, en = pts[:n].copy(), pts[n:].copy()
    rng = np.random.default_rng(2)
    en[:300] = st[:300] + rng.normal(0, 0.01, (300, 3)).astype(np.float32)
    en[300:310] = st[300:310]
    st[310:314] = np.nan
    lo, hi = ref.get_bounds()
    st[314:320] = (hi + 40).astype(np.float32)
    want = ref.find_path_raw_batch(st, en, max_pts=64, nthreads=8)
    got = pf.find_paths(st, en, max_points=64, corridors=True)
    assert close(got["geodesic_distance"], want["dist"]).all()
    assert beq(got["geodesic_distance"], want["dist"]).all()
    found = (want["flags"] & 4) != 0
    assert found.mean() > 0.3
    assert (got["num_points"][found] == want["num_points"][found]).all()
    assert (got["num_points"][~found] == 0).all()
    ran = found & ((want["flags"] & 1) == 0)  # A* / same-poly corridor exists
    assert (got["num_corridor"][ran] == want["num_polys"][ran]).all()
    for i in np.nonzero(ran)[0]:
        k = want["num_polys"][i]
        assert (got["corridor"][i, :k] == want["corridor"][i, :k]).all(), "corridors must be bit-exact"
    for i in np.nonzero(found)[0]:
        m = min(want["num_points"][i], 64)
        assert beq(got["points"][i, :m], want["pts"][i, :m]).all()
    # Detour-exact mode: status words and corridors of failed queries as well
    got2 = pf.find_paths(st, en, corridors=True, exact_status=True)
    astar_ran = ((want["flags"] & 2) != 0)
    assert (got2["status"][astar_ran, 0] == want["astar_status"][astar_ran]).all()
    for i in np.nonzero(astar_ran)[0]:
        k = want["num_polys"][i]
        assert got2["num_corridor"][i] == k and (got2["corridor"][i, :k] == want["corridor"][i, :k]).all()
    assert beq(got2["geodesic_distance"], want["dist"]).all()


def test_find_path_multigoal(scene):
    """findPath(MultiGoalShortestPath&), PathFinder.cpp:1515-1572: distance, closest goal index and
    path points against the oracle, with duplicate goals (equal sort keys), unprojectable goals /
    starts and goals far above the mesh (bound > geodesic, so the pruning shows)."""
    name, pf, ref = scene
    rng = np.random.default_rng(10)
    n, g = 400, 12
    starts = query_points(name, n, 51)
    ends = query_points(name, n * g, 52).reshape(n, g, 3)
    ends[:, 4] = ends[:, 2]
    ends[:, 7, 1] += 3.0
    lo, hi = ref.get_bounds()
    ends[::6, 1] = (hi + 50).astype(np.float32)
    starts[3] = (hi + 50).astype(np.float32)
    ends[8] = (hi + 50).astype(np.float32)
    ends[::4, 9] = starts[::4] + rng.normal(0, 0.2, (len(starts[::4]), 3)).astype(np.float32)
    for gg in (g, 1):
        e = np.ascontiguousarray(ends[:, :gg])
        wd, wi, wn, wp = ref.find_path_multigoal_batch(starts, e, max_pts=32, nthreads=8)
        got = pf.find_paths_multigoal(starts, e, max_points=32)
        assert (got["closest_end_point_index"] == wi).all(), "closest goal index must be exact"
        assert close(got["geodesic_distance"], wd).all() and beq(got["geodesic_distance"], wd).all()
        assert (got["num_points"] == wn).all()
        for i in np.nonzero(wi >= 0)[0]:
            m = min(wn[i], 32)
            assert beq(got["points"][i, :m], wp[i, :m]).all()
        assert (wi >= 0).mean() > 0.3
    # many goals: two rounds of pair searches (kMultiGoalFirst goals first, then those the reference
    # would not skip) must give the reference's answer
    n2, g2 = 150, 40
    st2 = query_points(name, n2, 55)
    en2 = query_points(name, n2 * g2, 56).reshape(n2, g2, 3)
    en2[:, 5, 1] += 3.0
    en2[:, 30, 1] += 2.5
    en2[::6, 2] = (hi + 50).astype(np.float32)
    en2[::4, 20] = st2[::4] + rng.normal(0, 0.2, (len(st2[::4]), 3)).astype(np.float32)
    wd, wi, _, _ = ref.find_path_multigoal_batch(st2, en2, nthreads=8)
    got = pf.find_paths_multigoal(st2, en2)
    assert (got["closest_end_point_index"] == wi).all() and beq(got["geodesic_distance"], wd).all()
    # multi-goal == min over the single-goal queries (src/tests/PathFinderTest.cpp:102-134) for
    # goals on the mesh, where the L2 bound cannot exceed the geodesic distance
    st = query_points(name, 200, 53, jitter=0.0)
    en = query_points(name, 200 * 6, 54, jitter=0.0).reshape(200, 6, 3)
    multi = pf.find_paths_multigoal(st, en)["geodesic_distance"]
    single = pf.find_paths(np.repeat(st, 6, axis=0), en.reshape(-1, 3))["geodesic_distance"].reshape(200, 6)
    assert beq(multi, single.min(axis=1)).all()


def test_multigoal_device_and_scalar_api():
    import torch
    from habitat_sim_b200.nav import MultiGoalShortestPath
    pf = gpu_pathfinder("c3_multiroom")
    ref = ref_pathfinder("c3_multiroom")
    starts = query_points("c3_multiroom", 300, 55)
    ends = query_points("c3_multiroom", 300 * 16, 56).reshape(300, 16, 3)
    host = pf.find_paths_multigoal(starts, ends, max_points=8)
    dev = pf.find_paths_multigoal(torch.from_numpy(starts).cuda(), torch.from_numpy(ends).cuda(), max_points=8)
    torch.cuda.synchronize()
    assert beq(dev["geodesic_distance"].cpu().numpy(), host["geodesic_distance"]).all()
    assert (dev["closest_end_point_index"].cpu().numpy() == host["closest_end_point_index"]).all()
    assert (dev["num_points"].cpu().numpy() == host["num_points"]).all()
    wd, wi, wn, wp = ref.find_path_multigoal_batch(starts[:5], ends[:5], max_pts=256)
    for i in range(5):
        p = MultiGoalShortestPath()
        p.requested_start = starts[i]
        p.requested_ends = ends[i]
        ok = pf.find_path(p)
        assert ok == bool(np.isfinite(wd[i]))
        assert p.closest_end_point_index == wi[i] and len(p.points) == wn[i]
        assert np.float32(p.geodesic_distance) == wd[i] or (np.isinf(wd[i]) and np.isinf(p.geodesic_distance))
        for k in range(wn[i]):
            assert beq(p.points[k], wp[i, k]).all()


def test_try_step(scene):
    name, pf, ref = scene
    from workloads.scenes import step_targets
    n = 8000
    s = ref.snap_batch(query_points(name, n, 33), 8)[0]
    s[np.isnan(s)] = 0
    t = step_targets(s, 5, 0.25)
    t[:2000] = step_targets(s[:2000], 6, 2.5)
    for sliding in (True, False):
        want = ref.try_step_batch(s, t, sliding, 8)
        got = pf.try_steps(s, t, sliding)
        assert close(got, want).all()
        assert beq(got, want).all()


def test_closest_obstacle(scene):
    name, pf, ref = scene
    pts = query_points(name, 10000, 34)
    hp, hn, hd = ref.obstacle_batch(pts, 2.0, 8)
    gp, gn, gd = pf.closest_obstacle_surface_points(pts, 2.0)
    assert close(gd, hd).all() and beq(gd, hd).all()
    assert beq(gp, hp).all() and beq(gn, hn).all()
    hd5 = ref.obstacle_batch(pts[:2000], 0.5, 8)[2]
    assert beq(pf.distances_to_closest_obstacle(pts[:2000], 0.5), hd5).all()


def test_closest_obstacle_tiers_on_the_multi_storey_mesh():
    """distance_to_closest_obstacle through all its tiers: on c4_building under 1 % of the queries need more
    than the 8-node per-thread pool of k_wall_lane and go on to the warp-per-query workspaces; radii from
    0 to 1e6 (PF.cpp:1794-1812 passes the radius through; DQ.cpp:3470-3655 shrinks it to the nearest wall)."""
    name = "c4_building"
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    pf, ref = gpu_pathfinder(name), ref_pathfinder(name)
    pts = pointnav_pairs(NavMeshGeom(navmesh_image(name)), 60_000, 9, jitter=0.05)[0]
    for radius, n in ((2.0, 60_000), (0.0, 3000), (0.05, 3000), (7.5, 20_000), (1e6, 20_000)):
        hp, hn, hd = ref.obstacle_batch(pts[:n], radius, 8)
        gp, gn, gd = pf.closest_obstacle_surface_points(pts[:n], radius)
        assert beq(gd, hd).all() and beq(gp, hp).all() and beq(gn, hn).all(), radius


def test_random_points_near(scene):
    """get_random_navigable_point_near (getRandomNavigablePointInCircle, PF.cpp:1283-1332, trap T6): the
    oracle's samples on the same uniform stream, plain and island restricted; and the properties the
    reference's tests/test_nav.py:159-183 checks (navigable, within the radius)."""
    name, pf, ref = scene
    n = 1500
    rng = np.random.default_rng(4)
    centers = ref.snap_batch(query_points(name, n, 36, jitter=0.05), 8)[0]
    centers[np.isnan(centers)] = 0
    isl = np.full(n, -1, np.int32)
    isl[n // 2:] = rng.integers(0, ref.num_islands, n - n // 2)
    for radius, tries in ((2.0, 100), (0.5, 8)):
        want = ref.random_points_near(centers, radius, tries, isl, mode=1, seed=5, query0=40)
        got = pf.random_navigable_points_near(centers, radius, tries, isl, seed=5, query0=40)
        assert beq(got, want).all()
        ok = np.isfinite(got).all(axis=1)
        assert (np.linalg.norm((got - centers)[ok][:, [0, 2]], axis=1) < radius).all()
        assert pf.are_navigable(got[ok]).all()
    import torch
    dev_out = pf.random_navigable_points_near(torch.from_numpy(centers).cuda(), 2.0, 100, isl, seed=5, query0=40)
    assert beq(dev_out.cpu().numpy(), ref.random_points_near(centers, 2.0, 100, isl, mode=1, seed=5, query0=40)).all()
    # scalar API as the reference's test uses it
    pf.seed(3)
    for _ in range(20):
        p = pf.get_random_navigable_point()
        q = pf.get_random_navigable_point_near(p, 5.0, max_tries=100)
        assert pf.is_navigable(q) and np.linalg.norm(np.asarray(p) - np.asarray(q)) <= 5.0 + 4.0  # xz test: y may differ
    with pytest.raises(ValueError):
        pf.get_random_navigable_point_near(centers[0], 1.0, island_index=999)


def test_random_points(scene):
    name, pf, ref = scene
    n = 3000
    rng = np.random.default_rng(3)
    isl = np.full(n, -1, np.int32)
    isl[n // 2:] = rng.integers(0, ref.num_islands, n - n // 2)
    want_p, want_r = ref.random_points(n, 10, isl, mode=1, seed=77, query0=123)
    got_p, got_r = pf.random_navigable_points(n, 10, isl, seed=77, query0=123)
    assert (got_r == want_r).all()
    assert beq(got_p, want_p).all()
    ok = ~np.isnan(got_p[:, 0])
    assert pf.are_navigable(got_p[ok]).all()
    # shard-count invariance: sample i only depends on (seed, query0 + i)
    a, _ = pf.random_navigable_points(100, 10, -1, seed=5, query0=40)
    b, _ = pf.random_navigable_points(50, 10, -1, seed=5, query0=90)
    assert beq(a[50:], b).all()


@pytest.mark.parametrize("name", ["c1_room", "c2_apartment", "t_building", "ref_simple_room", "ref_stage_floor1"])
def test_golden_vectors(name):
    """committed fixtures (tests/golden/*.npz, made by make_golden.py from the oracle)"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    from habitat_sim_b200.nav import PathFinder
    pf = PathFinder(0)
    assert pf.load_nav_mesh_bytes(g["image"].tobytes())
    sp, sr, si = pf.snap_points(g["starts"])
    assert (sr == g["snap_refs"]).all() and (si == g["snap_isl"]).all() and beq(sp, g["snap_pts"]).all()
    r = pf.find_paths(g["starts"], g["ends"], max_points=32, corridors=True, exact_status=True)
    assert beq(r["geodesic_distance"], g["dist"]).all()
    ran = (g["flags"] & 2) != 0
    assert (r["num_corridor"][ran] == g["num_polys"][ran]).all()
    k = np.minimum(g["num_polys"], 64)
    for i in np.nonzero(ran)[0]:
        assert (r["corridor"][i, :k[i]] == g["corridor"][i, :k[i]]).all()
    assert beq(pf.try_steps(g["snap_pts"], g["step_targets"], True), g["step_sliding"]).all()
    assert beq(pf.try_steps(g["snap_pts"], g["step_targets"], False), g["step_nosliding"]).all()
    hp, hn, hd = pf.closest_obstacle_surface_points(g["starts"], 2.0)
    assert beq(hd, g["hit_dist"]).all() and beq(hp, g["hit_pos"]).all() and beq(hn, g["hit_normal"]).all()
    seed = int(g["seed"]) if "seed" in g else {"c1_room": 11, "c2_apartment": 12, "t_building": 13}[name]
    rp, rr = pf.random_navigable_points(len(g["rand_islands"]), 10, g["rand_islands"], seed=seed, query0=1000)
    assert (rr == g["rand_refs"]).all() and beq(rp, g["rand_pts"]).all()
    assert pf.num_islands == int(g["num_islands"])


def test_device_pointer_path_matches_host_path():
    import torch
    pf = gpu_pathfinder("t_building")
    pts = query_points("t_building", 4000, 35)
    st, en = pts[:2000], pts[2000:]
    host = pf.find_paths(st, en, max_points=16)
    dev = pf.find_paths(torch.from_numpy(st).cuda(), torch.from_numpy(en).cuda(), max_points=16)
    torch.cuda.synchronize()
    assert beq(dev["geodesic_distance"].cpu().numpy(), host["geodesic_distance"]).all()
    assert (dev["num_points"].cpu().numpy() == host["num_points"]).all()
    sp = pf.snap_points(torch.from_numpy(st).cuda())
    assert beq(sp[0].cpu().numpy(), pf.snap_points(st)[0]).all()


def test_scalar_api_matches_reference_conventions():
    """SPB.cpp:177-270 failure sentinels through the scalar drop-in methods"""
    from habitat_sim_b200.nav import ShortestPath
    pf = gpu_pathfinder("c1_room")
    ref = ref_pathfinder("c1_room")
    lo, hi = pf.get_bounds()
    far = hi + 100
    assert np.isnan(pf.snap_point(far)).all()
    assert pf.get_island(far) == -1
    assert not pf.is_navigable(far)
    assert np.isinf(pf.distance_to_closest_obstacle(far))
    hr = pf.closest_obstacle_surface_point(far)
    assert np.isinf(hr.hit_dist) and (hr.hit_pos == 0).all() and (hr.hit_normal == 0).all()
    assert (pf.try_step(far, far + 1) == far.astype(np.float32)).all()
    p = ShortestPath()
    p.requested_start = far
    p.requested_end = far + 1
    assert not pf.find_path(p) and np.isinf(p.geodesic_distance) and p.points == []
    a = ref.random_points(2, 10, [0, 0], mode=1, seed=1)[0]
    p.requested_start, p.requested_end = a[0], a[1]
    assert pf.find_path(p)
    d, n, pts = ref.find_path_batch(a[:1], a[1:], 256)
    assert p.geodesic_distance == d[0] and len(p.points) == n[0]
    with pytest.raises(ValueError):
        pf.snap_point(a[0], island_index=99)


def test_topdown_views_match_reference_loops():
    pf = gpu_pathfinder("c2_apartment")
    ref = ref_pathfinder("c2_apartment")
    mpp, height = 0.1, 0.1
    view = pf.get_topdown_view(mpp, height)
    isl = pf.get_topdown_island_view(mpp, height)
    pts = pf._topdown_grid(mpp, height).reshape(-1, 3)
    want = ref.is_navigable_batch(pts, 0.5, 8).reshape(view.shape)
    assert (view == want).all() and view.any()
    wi = np.where(want.reshape(-1), ref.snap_batch(pts, 8)[2], -1).reshape(view.shape)
    assert (isl == wi).all()


def test_full_size_properties():
    """BASELINE config C4 sizes: properties that need no oracle pass over the full batch."""
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    pf = gpu_pathfinder("c4_building")
    ref = ref_pathfinder("c4_building")
    geom = NavMeshGeom(navmesh_image("c4_building"))
    n = 200_000
    st, en = pointnav_pairs(geom, n, 7)
    r = pf.find_paths(st, en)
    d = r["geodesic_distance"]
    sp, sr, si = pf.snap_points(st)
    ep, er, ei = pf.snap_points(en)
    found = np.isfinite(d)
    assert 0.2 < found.mean() < 1.0
    # geodesic >= horizontal distance between the path's end points (the funnel starts / ends at
    # the REQUESTED points clamped in xz to the first / last poly, trap T2; on ramps their
    # heights differ from the snapped ones), same island required
    eu = np.linalg.norm((sp - ep)[:, [0, 2]], axis=1)
    eu -= np.linalg.norm((sp - st)[:, [0, 2]], axis=1) + np.linalg.norm((ep - en)[:, [0, 2]], axis=1)
    assert (d[found] >= eu[found] * (1 - 1e-4) - 1e-3).all()
    assert (si[found] == ei[found]).all()
    # idempotence of snapping, navigability of snapped points
    ok = sr != 0
    sp2, sr2, _ = pf.snap_points(sp[ok])
    assert (sr2 != 0).all()
    assert pf.are_navigable(sp[ok]).all()
    # batch-split invariance (multi-GPU sharding is a plain split of the batch)
    d2 = np.concatenate([pf.find_paths(st[:n // 3], en[:n // 3])["geodesic_distance"],
                         pf.find_paths(st[n // 3:], en[n // 3:])["geodesic_distance"]])
    assert beq(d, d2).all()
    # success fraction and values equal the oracle's on a bounded sample
    m = 4000
    want = ref.find_path_batch(st[:m], en[:m], 0, 8)[0]
    assert beq(d[:m], want).all()


def test_snap_pipeline_variants(monkeypatch):
    """The candidate-list nearest-poly pipeline (hbn_snap.cuh): island-restricted batches, the
    device-side fallback when the candidate scratch overflows, and the lane-group kernel must all
    give the reference's refs and points."""
    for name in ("t_building", "c4_building"):
        pf = gpu_pathfinder(name)
        ref = ref_pathfinder(name)
        n = 30000
        pts = query_points(name, n, 41, jitter=0.05)
        rng = np.random.default_rng(2)
        pts[:6000] += rng.normal(0, 0.7, (6000, 3)).astype(np.float32)  # off the mesh, between storeys
        pts[6000:6003] = np.nan
        want_p, want_r, want_i = ref.snap_batch(pts, 8)
        isl = rng.integers(0, ref.num_islands, n).astype(np.int32)
        isl[::3] = want_i[::3]  # a third of the points ask for the island they are on
        isl[isl < 0] = 0
        wp, wr = ref.snap_island_batch(pts, isl)
        for env in ({}, {"snap_cap": 1000}, {"snap_group": 1}):
            for k in ("snap_cap", "snap_group"):
                pf.set_option(k, 0)
            for k, v in env.items():
                pf.set_option(k, v)
            got_p, got_r, got_i = pf.snap_points(pts)
            assert (got_r == want_r).all() and (got_i == want_i).all() and beq(got_p, want_p).all(), env
            gp, gr, gi = pf.snap_points(pts, isl)
            assert (gr == wr).all() and beq(gp, wp).all(), env


def test_lane_search_small_grid_generation_wrap(monkeypatch):
    """k_astar_lane with one block per SM: every lane serves ~40 queries, so node-table
    generations wrap and the warp-cooperative wipe runs; results must equal the full grid's and
    the oracle's."""
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    name = "c4_building"
    n = 200_000
    st, en = pointnav_pairs(NavMeshGeom(navmesh_image(name)), n, 13)
    full = gpu_pathfinder(name).find_paths(st, en)["geodesic_distance"]
    monkeypatch.setenv("HBN_FP_BLOCKS_PER_SM", "1")
    pf = gpu_pathfinder(name)
    d = pf.find_paths(st, en)["geodesic_distance"]
    d2 = pf.find_paths(st, en)["geodesic_distance"]  # generations persist across launches
    assert beq(d, full).all() and beq(d2, full).all()
    m = 4000
    want = ref_pathfinder(name).find_path_batch(st[:m], en[:m], 0, 8)[0]
    assert beq(d[:m], want).all()


def test_find_path_exact_status_and_fast_fail():
    """Detour-exact mode (status words and corridors of unsuccessful searches too) and the default
    mode (stop at pool exhaustion) give the reference's corridors, status words and distances."""
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    for name, n in (("t_building", 3000), ("c4_building", 6000)):
        pf = gpu_pathfinder(name)
        ref = ref_pathfinder(name)
        if name == "c4_building":
            st, en = pointnav_pairs(NavMeshGeom(navmesh_image(name)), n, 11)
        else:
            pts = query_points(name, 2 * n, 5)
            st, en = pts[:n].copy(), pts[n:].copy()
        want = ref.find_path_raw_batch(st, en, max_pts=32, nthreads=8)
        got = pf.find_paths(st, en, max_points=32, corridors=True, exact_status=True)
        assert beq(got["geodesic_distance"], want["dist"]).all()
        astar_ran = (want["flags"] & 2) != 0
        assert (got["status"][astar_ran, 0] == want["astar_status"][astar_ran]).all()
        assert (got["num_corridor"][astar_ran] == want["num_polys"][astar_ran]).all()
        for i in np.nonzero(astar_ran)[0]:
            k = want["num_polys"][i]
            assert (got["corridor"][i, :k] == want["corridor"][i, :k]).all()
        found = (want["flags"] & 4) != 0
        assert (got["num_points"][found] == want["num_points"][found]).all()
        for i in np.nonzero(found)[0]:
            m = min(want["num_points"][i], 32)
            assert beq(got["points"][i, :m], want["pts"][i, :m]).all()
        fast = pf.find_paths(st, en)  # default mode: stop at pool exhaustion
        assert beq(fast["geodesic_distance"], want["dist"]).all()

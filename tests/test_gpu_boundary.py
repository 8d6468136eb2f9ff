"""GPU tests of the drop-in boundary and of the BASELINE configurations at their own sizes, all
against the reference's own PathFinder.cpp (oracle/_ref/libhbn_ref.so):
  * the live hand-over hbn_navmesh_create_from_tiles from a freshly BUILT
    reference navmesh + its poly_islands (SURVEY trap T5), and the MSET image the handle writes back;
  * build_navmesh_vertices / get_topdown_view / get_topdown_island_view against the reference's
    getNavMeshData / getTopDownView / getTopDownIslandView (golden fixtures + live);
  * config C2: 1024 envs x 100 DEPENDENT steps of try_step -> find_path;
  * config C3: 4096 starts x 64 goals MultiGoalShortestPath."""
import os

import numpy as np
import pytest

from conftest import beq, gpu_pathfinder, navmesh_image, query_points, ref_pathfinder

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _fresh_reference(name):
    """the reference PathFinder right after build() -- not re-loaded (trap T5)"""
    from oracle.ref import RefPathFinder
    from workloads import scenes
    ref = RefPathFinder()
    v, t = scenes.scene_triangles(name)
    assert ref.build(v, t)
    return ref


def _from_live_tiles(ref, device=0):
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import PathFinder
    pf = PathFinder(device)
    orig, wh, mm = ref.navmesh_params()
    refs, isl = ref.poly_islands()
    tiles = [(tref, blob) for tref, _idx, blob in ref.tile_blobs()]
    radii = [ref.island_radius(i) for i in range(ref.num_islands)]
    assert pf.load_from_tiles(tiles, orig, float(wh[0]), float(wh[1]), int(mm[0]), int(mm[1]), poly_islands=isl,
                              island_radii=radii, bounds=ref.get_bounds())
    return pf


def _check_against(pf, ref, name, n=3000):
    assert pf.num_islands == ref.num_islands
    for i in range(ref.num_islands):
        assert pf.island_radius(i) == ref.island_radius(i)
        assert pf.island_area(i) == ref.navigable_area(i)
    assert np.float32(pf.navigable_area) == np.float32(ref.navigable_area())
    for a, b in zip(pf.get_bounds(), ref.get_bounds()):
        assert (a == b).all()
    pts = query_points(name, 2 * n, 91)
    want_p, want_r, want_i = ref.snap_batch(pts, 8)
    got_p, got_r, got_i = pf.snap_points(pts)
    assert (got_r == want_r).all() and (got_i == want_i).all() and beq(got_p, want_p).all()
    want = ref.find_path_raw_batch(pts[:n], pts[n:], max_pts=32, nthreads=8)
    got = pf.find_paths(pts[:n], pts[n:], max_points=32, corridors=True)
    assert beq(got["geodesic_distance"], want["dist"]).all()
    ran = ((want["flags"] & 4) != 0) & ((want["flags"] & 1) == 0)
    for i in np.nonzero(ran)[0]:
        k = want["num_polys"][i]
        assert (got["corridor"][i, :k] == want["corridor"][i, :k]).all()
    isl = np.random.default_rng(1).integers(0, ref.num_islands, 300).astype(np.int32)
    wp, wr = ref.snap_island_batch(pts[:300], isl)
    gp, gr, _ = pf.snap_points(pts[:300], isl)
    assert (gr == wr).all() and beq(gp, wp).all()
    rp, rr = ref.random_points(300, 10, isl, mode=1, seed=3, query0=10)
    qp, qr = pf.random_navigable_points(300, 10, isl, seed=3, query0=10)
    assert (qr == rr).all() and beq(qp, rp).all()


@pytest.mark.parametrize("name", ["c1_room", "c2_apartment", "c3_multiroom"])
def test_create_from_live_tiles_of_a_freshly_built_navmesh(name, tmp_path):
    """hbn_navmesh_create_from_tiles with the finalised tiles and the IslandSystem's poly -> island map
    of a navmesh the reference has just BUILT (never saved or re-loaded): the call INTEGRATION.md
    section 1 makes from PathFinder::Impl::initNavQuery.  Then the handle writes an MSET image
    (hbn_navmesh_save_mset) that must be byte-identical to the reference's saveNavMesh of that mesh."""
    ref = _fresh_reference(name)
    pf = _from_live_tiles(ref)
    _check_against(pf, ref, name)
    # without poly_islands the library floods the islands itself from the finalised flags: the
    # numbering of a RE-LOADED mesh (trap T5) -- equal to the reference's after save -> load
    from oracle.ref import RefPathFinder
    reloaded = RefPathFinder()
    assert reloaded.load_bytes(ref.save_bytes())
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import PathFinder
    pf2 = PathFinder(0)
    orig, wh, mm = ref.navmesh_params()
    assert pf2.load_from_tiles([(t, b) for t, _i, b in ref.tile_blobs()], orig, float(wh[0]), float(wh[1]),
                               int(mm[0]), int(mm[1]))
    _check_against(pf2, reloaded, name, n=1000)
    # MSET writer from live tiles (PF.cpp:1177-1223)
    pf.set_nav_mesh_settings_bytes(ref.default_settings())
    path = str(tmp_path / "live.navmesh")
    assert pf.save_nav_mesh(path)
    assert open(path, "rb").read() == ref.save_bytes()


def test_create_from_live_tiles_tiled_building():
    """the same hand-over for a multi-tile navmesh (cross-tile links, 99 tiles)"""
    ref = ref_pathfinder("c4_building")
    pf = _from_live_tiles(ref)
    _check_against(pf, ref, "c4_building", n=2000)
    pf.set_nav_mesh_settings_bytes(ref.default_settings())
    assert pf.save_nav_mesh_bytes() == ref.save_bytes()


@pytest.mark.parametrize("name", ["c1_room", "c2_apartment", "t_building", "ref_simple_room", "ref_stage_floor1"])
def test_navmesh_geometry_and_topdown_views_match_reference_goldens(name):
    """build_navmesh_vertices (getNavMeshData, PF.cpp:1898-1944) and the top-down views
    (PF.cpp:1833-1896) against fixtures written by the reference's own functions."""
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import PathFinder
    g = np.load(os.path.join(GOLD, name + ".npz"))
    pf = PathFinder(0)
    assert pf.load_nav_mesh_bytes(g["image"].tobytes())
    v = np.asarray(pf.build_navmesh_vertices(-1), np.float32).reshape(-1, 3)
    assert v.shape == g["nav_verts_all"].shape and beq(v, g["nav_verts_all"]).all()
    last = int(g["num_islands"]) - 1
    v = np.asarray(pf.build_navmesh_vertices(last), np.float32).reshape(-1, 3)
    assert v.shape == g["nav_verts_last_island"].shape and beq(v, g["nav_verts_last_island"]).all()
    assert pf.build_navmesh_vertex_indices(last) == list(range(len(v)))
    h = float(g["topdown_height"])
    td = pf.get_topdown_view(0.25, h)
    assert td.shape == g["topdown"].shape and (td == g["topdown"]).all() and td.any()
    tdi = pf.get_topdown_island_view(0.25, h)
    assert (tdi == g["topdown_islands"]).all()


def test_topdown_views_match_reference_live():
    pf = gpu_pathfinder("t_building")
    ref = ref_pathfinder("t_building")
    for mpp, h in ((0.1, 0.2), (0.37, 3.3)):
        assert (pf.get_topdown_view(mpp, h) == ref.topdown_view(mpp, h)).all()
        assert (pf.get_topdown_island_view(mpp, h) == ref.topdown_view(mpp, h, islands=True)).all()


def test_config_c2_dependent_chain_1024_envs_100_steps():
    """BASELINE config 2 at its own batch size: 1024 envs on the apartment navmesh, per step
    try_step(p, p + 0.25 dir) then geodesic distance to the env's goal; the GPU runs its own chain
    (each step's positions come from its own previous try_step) and the reference runs its own; every
    position and every distance of all 100 steps must be equal bit for bit."""
    name = "c2_apartment"
    pf, ref = gpu_pathfinder(name), ref_pathfinder(name)
    envs, steps = 1024, 100
    isl = np.zeros(envs, np.int32)
    pos, _ = ref.random_points(envs, 10, isl, mode=1, seed=21, query0=0)
    goal, _ = ref.random_points(envs, 10, isl, mode=1, seed=22, query0=0)
    rng = np.random.default_rng(5)
    gp, rp = pos.copy(), pos.copy()
    moved = 0
    for s in range(steps):
        th = rng.uniform(0, 2 * np.pi, envs).astype(np.float32)
        d = np.stack([np.cos(th), np.zeros_like(th), np.sin(th)], 1).astype(np.float32) * np.float32(0.25)
        slide = s % 5 != 4  # every fifth step without sliding
        if s % 2 == 0:  # the fused step (one CUDA graph replay) and the two separate calls alternate
            gn, gd = pf.env_steps(gp, gp + d, goal, allow_sliding=slide)
        else:
            gn = pf.try_steps(gp, gp + d, allow_sliding=slide)
            gd = pf.geodesic_distances(gn, goal)
        rn = ref.try_step_batch(rp, rp + d, slide, 8)
        assert beq(gn, rn).all(), f"step {s}: positions differ"
        rd = ref.find_path_batch(rn, goal, 0, 8)[0]
        assert beq(gd, rd).all(), f"step {s}: distances differ"
        moved += int((gn != gp).any(axis=1).sum())
        gp, rp = gn, rn
    assert moved > envs * steps // 2


def test_config_c3_multigoal_4096_starts_64_goals():
    """BASELINE config 3 at its own size: 4096 starts x 64 goals on the multi-room navmesh.  The GPU
    answers all 4096; the reference a 1024-start sample (all 64 goals each): distance, index of the
    closest goal and number of path points must be equal."""
    name = "c3_multiroom"
    pf, ref = gpu_pathfinder(name), ref_pathfinder(name)
    n, g, m = 4096, 64, 1024
    starts = query_points(name, n, 61)
    ends = query_points(name, n * g, 62).reshape(n, g, 3)
    ends[::7, 3] = ends[::7, 5]  # duplicate goals: equal sort keys
    got = pf.find_paths_multigoal(starts, ends, max_points=0)
    wd, wi, wn, _ = ref.find_path_multigoal_batch(starts[:m], ends[:m], 0, 8)
    assert beq(got["geodesic_distance"][:m], wd).all()
    assert (got["closest_end_point_index"][:m] == wi).all()
    assert np.isfinite(wd).mean() > 0.5
    # the rest through a size-independent property: the answer for a start does not depend on the batch
    sel = np.arange(m, n, 97)
    sub = pf.find_paths_multigoal(starts[sel], ends[sel], max_points=0)
    assert beq(sub["geodesic_distance"], got["geodesic_distance"][sel]).all()
    assert (sub["closest_end_point_index"] == got["closest_end_point_index"][sel]).all()


@pytest.mark.parametrize("name", ["c2_apartment", "t_building"])
def test_env_step_equals_try_step_then_find_path(name):
    """hbn_env_step[_dev] (one fused PointNav step: try_step, then find_path from the new position,
    projections shared, CUDA graph for host buffers) against the two separate calls and against the
    reference, including failed steps (unchanged start), far targets and off-mesh goals."""
    import torch
    from workloads.scenes import step_targets
    pf, ref = gpu_pathfinder(name), ref_pathfinder(name)
    rng = np.random.default_rng(3)
    for n in (1, 33, 1024, 5000):
        pos = ref.snap_batch(query_points(name, n, 70 + n))[0]
        lo, hi = ref.get_bounds()
        pos[np.isnan(pos)] = 0
        pos[n // 2:] = query_points(name, n - n // 2, 71 + n, jitter=0.3)  # unsnapped starts too
        if n > 8:
            pos[5] = hi + 30  # try_step fails: position unchanged, no path
        tgt = step_targets(pos, 9, 0.25)
        tgt[: n // 3] = step_targets(pos[: n // 3], 10, 3.0)  # across walls / other islands: sliding, nudges
        goals = query_points(name, n, 72 + n)
        if n > 8:
            goals[7] = np.nan
        for sliding in (True, False):
            want_pos = ref.try_step_batch(pos, tgt, sliding, 8)
            want_d = ref.find_path_batch(want_pos, goals, 0, 8)[0]
            sep_pos = pf.try_steps(pos, tgt, sliding)
            sep_d = pf.geodesic_distances(sep_pos, goals)
            assert beq(sep_pos, want_pos).all() and beq(sep_d, want_d).all()
            for rep in range(3):  # the first call captures the graph, the others replay it
                got_pos, got_d = pf.env_steps(pos, tgt, goals, sliding)
                assert beq(got_pos, want_pos).all(), (n, sliding, rep)
                assert beq(got_d, want_d).all(), (n, sliding, rep)
            tp, td = pf.env_steps(torch.from_numpy(pos).cuda(), torch.from_numpy(tgt).cuda(),
                                  torch.from_numpy(goals).cuda(), sliding)
            assert beq(tp.cpu().numpy(), want_pos).all() and beq(td.cpu().numpy(), want_d).all()
    # a bigger batch afterwards moves the scratch: the cached graphs must be rebuilt, not replayed stale
    n = 1024
    pos = ref.snap_batch(query_points(name, n, 70 + n))[0]
    pos[np.isnan(pos)] = 0
    pos[n // 2:] = query_points(name, n - n // 2, 71 + n, jitter=0.3)
    pos[5] = ref.get_bounds()[1] + 30
    tgt = step_targets(pos, 9, 0.25)
    tgt[: n // 3] = step_targets(pos[: n // 3], 10, 3.0)
    goals = query_points(name, n, 72 + n)
    goals[7] = np.nan
    pf.find_paths(query_points(name, 300_000, 1), query_points(name, 300_000, 2))
    got_pos, got_d = pf.env_steps(pos, tgt, goals, True)
    want_pos = ref.try_step_batch(pos, tgt, True, 8)
    assert beq(got_pos, want_pos).all() and beq(got_d, ref.find_path_batch(want_pos, goals, 0, 8)[0]).all()


def test_search_state_cap_and_reserve():
    """Options: the per-lane search state of find_path is sized from the batch (a scalar query does not
    take GBs), can be capped ("lane_scratch_bytes": the grid shrinks, results do not change), and
    hbn_navmesh_reserve sizes everything up front so that later calls allocate nothing."""
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    name = "c4_building"
    st, en = pointnav_pairs(NavMeshGeom(navmesh_image(name)), 120_000, 17)
    full = gpu_pathfinder(name)
    want = full.find_paths(st, en)["geodesic_distance"]
    assert full.scratch_bytes > (8 << 30)  # 120 k queries: every lane of the grid is in use (15 GB of search state)
    small = gpu_pathfinder(name)
    small.find_paths(st[:1], en[:1])
    assert small.scratch_bytes < (32 << 20), small.scratch_bytes  # one block of lanes, not the whole grid (15 GB)
    assert beq(small.find_paths(st[:3000], en[:3000])["geodesic_distance"], want[:3000]).all()
    capped = gpu_pathfinder(name)
    capped.set_option("lane_scratch_bytes", 256 << 20)
    d = capped.find_paths(st, en)["geodesic_distance"]
    assert beq(d, want).all()
    assert capped.scratch_bytes < (1 << 30), capped.scratch_bytes  # 256 MB of search state + corridor rings, snap candidates ...
    with pytest.raises(Exception):
        capped.set_option("no_such_option", 1)
    res = gpu_pathfinder(name)
    res.reserve(20_000)
    before = res.scratch_bytes
    res.find_paths(st[:20_000], en[:20_000])
    res.try_steps(st[:20_000], en[:20_000])
    res.env_steps(st[:20_000], en[:20_000], st[:20_000])
    res.snap_points(st[:20_000])
    res.closest_obstacle_surface_points(st[:20_000])
    assert res.scratch_bytes == before


def test_calls_on_different_streams_do_not_race_on_the_scratch():
    """All scratch belongs to the handle; numpy calls run on the handle's stream, torch calls on torch's
    current stream.  Back-to-back calls on different streams without any synchronisation in between
    must not overwrite each other's scratch (the handle orders them with an event)."""
    import torch
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    name = "c4_building"
    pf = gpu_pathfinder(name)
    st, en = pointnav_pairs(NavMeshGeom(navmesh_image(name)), 60_000, 19)
    want = pf.find_paths(st, en)["geodesic_distance"]
    want2 = pf.find_paths(st[::-1].copy(), en[::-1].copy())["geodesic_distance"]
    s1, e1 = torch.from_numpy(st).cuda(), torch.from_numpy(en).cuda()
    s2, e2 = torch.from_numpy(st[::-1].copy()).cuda(), torch.from_numpy(en[::-1].copy()).cuda()
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        a = pf.find_paths(s1, e1)["geodesic_distance"]          # torch's current stream
        with torch.cuda.stream(side):
            b = pf.find_paths(s2, e2)["geodesic_distance"]      # another stream, no sync in between
        c = pf.find_paths(st[:5000], en[:5000])["geodesic_distance"]  # numpy: the handle's own stream
        torch.cuda.synchronize()
        assert beq(a.cpu().numpy(), want).all()
        assert beq(b.cpu().numpy(), want2).all()
        assert beq(c, want[:5000]).all()


def test_multi_gpu_dispatcher_on_devices():
    """nav.MultiGpuPathFinder with real handles: every visible device (and, so that the thread / slice
    logic runs on a one-GPU box too, two handles on device 0).  The split batch equals the unsplit one."""
    import torch
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import MultiGpuPathFinder
    from workloads.scenes import NavMeshGeom, pointnav_pairs
    name = "c4_building"
    img = navmesh_image(name)
    st, en = pointnav_pairs(NavMeshGeom(img), 50_001, 23)
    single = gpu_pathfinder(name)
    want = single.find_paths(st, en)["geodesic_distance"]
    want_step = single.try_steps(st[:4000], en[:4000])
    rp, rr = single.random_navigable_points(3001, seed=9, query0=100)
    configs = [[0, 0]]
    if torch.cuda.device_count() > 1:
        configs.append(list(range(torch.cuda.device_count())))
    for devices in configs:
        mg = MultiGpuPathFinder(devices)
        assert mg.load_nav_mesh_bytes(img) and mg.world == len(devices)
        mg.set_option("lane_scratch_bytes", 2 << 30)
        assert beq(mg.geodesic_distances(st, en), want).all(), devices
        assert beq(mg.try_steps(st[:4000], en[:4000]), want_step).all()
        gp, gr = mg.random_navigable_points(3001, seed=9, query0=100)
        assert beq(gp, rp).all() and (gr == rr).all()
        assert mg.num_islands == single.num_islands  # scalar API: the first handle
        mg.close()

"""The reference's Python nav tests (tests/test_nav.py), restated over the drop-in PathFinder on the
procedural scenes (the reference's scene datasets are not in its tree): same calls, same assertions."""
import math
import os

import numpy as np
import pytest

from conftest import gpu_pathfinder, navmesh_image, ref_pathfinder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["c2_apartment", "t_building"])
def pathfinder(request):
    pf = gpu_pathfinder(request.param)
    pf.seed(0)
    return pf


def test_pathfinder_save_load_round_trip(pathfinder, tmp_path):  # test_nav.py:102-112
    from habitat_sim_b200.nav import PathFinder
    save_path = str(tmp_path / "round_trip.navmesh")
    assert pathfinder.save_nav_mesh(save_path)
    pf2 = PathFinder()
    assert pf2.load_nav_mesh(save_path) and pf2.is_loaded
    assert pathfinder.nav_mesh_settings == pf2.nav_mesh_settings
    assert abs(pathfinder.navigable_area - pf2.navigable_area) < 1e-3


def test_pathfinder_get_bounds(pathfinder):  # test_nav.py:143-152
    lo, hi = pathfinder.get_bounds()
    assert len(lo) == 3 and len(hi) == 3 and (np.asarray(lo) < np.asarray(hi)).all()


def test_is_navigable_with_custom_y_delta(pathfinder):  # test_nav.py:189-196
    pt = pathfinder.get_random_navigable_point()
    assert pathfinder.is_navigable(pt)
    assert pathfinder.is_navigable(pt, max_y_delta=0.5)
    lifted = np.array(pt) + np.array([0.0, 10.0, 0.0])
    assert not pathfinder.is_navigable(lifted, max_y_delta=0.5)


def test_obstacle_queries(pathfinder):  # test_nav.py:204-226
    from habitat_sim_b200.nav import HitRecord
    pt = pathfinder.get_random_navigable_point()
    dist = pathfinder.distance_to_closest_obstacle(pt, max_search_radius=5.0)
    assert dist >= 0.0 and math.isfinite(dist)
    assert pathfinder.distance_to_closest_obstacle(pt) >= 0.0
    hit = pathfinder.closest_obstacle_surface_point(pt, max_search_radius=5.0)
    assert isinstance(hit, HitRecord) and len(hit.hit_pos) == 3 and len(hit.hit_normal) == 3
    assert math.isfinite(hit.hit_dist)


def test_try_step(pathfinder):  # test_nav.py:242-268
    start = np.asarray(pathfinder.get_random_navigable_point())
    end = start + np.array([0.1, 0.0, 0.1], np.float32)
    result = pathfinder.try_step(start, end)
    assert len(result) == 3
    assert np.allclose(result, pathfinder.snap_point(result), atol=0.1)
    assert len(pathfinder.try_step_no_sliding(start, end)) == 3
    far = start + np.array([100.0, 0.0, 100.0], np.float32)
    assert not np.allclose(pathfinder.try_step(start, far), far, atol=1.0)


def test_shortest_path_no_path(pathfinder):  # test_nav.py:283-292
    from habitat_sim_b200.nav import ShortestPath
    sp = ShortestPath()
    sp.requested_start = np.array([0.0, 0.0, 0.0])
    sp.requested_end = np.array([99999.0, 99999.0, 99999.0])
    assert not pathfinder.find_path(sp)
    assert sp.geodesic_distance == float("inf")


def test_multi_goal_shortest_path(pathfinder):  # test_nav.py:299-337
    from habitat_sim_b200.nav import MultiGoalShortestPath, ShortestPath
    for _ in range(5):
        start = pathfinder.get_random_navigable_point()
        goals = [pathfinder.get_random_navigable_point() for _ in range(5)]
        mgsp = MultiGoalShortestPath()
        mgsp.requested_start = start
        mgsp.requested_ends = goals
        if pathfinder.find_path(mgsp):
            assert mgsp.geodesic_distance < float("inf") and len(mgsp.points) > 0
            assert 0 <= mgsp.closest_end_point_index < len(goals)
        else:
            assert mgsp.geodesic_distance == float("inf")
        single = MultiGoalShortestPath()
        single.requested_start = start
        single.requested_ends = [goals[0]]
        sp = ShortestPath()
        sp.requested_start = start
        sp.requested_end = goals[0]
        found_mg, found_sg = pathfinder.find_path(single), pathfinder.find_path(sp)
        assert found_mg == found_sg
        if found_mg:
            assert abs(single.geodesic_distance - sp.geodesic_distance) < 1e-5


def test_navmesh_islands_and_area(pathfinder):  # test_nav.py:367-483, against the oracle instead of the cache
    name = "c2_apartment" if pathfinder.num_islands == ref_pathfinder("c2_apartment").num_islands and \
        abs(pathfinder.navigable_area - ref_pathfinder("c2_apartment").navigable_area()) < 1e-6 else "t_building"
    ref = ref_pathfinder(name)
    assert pathfinder.num_islands == ref.num_islands
    total = 0.0
    for i in range(pathfinder.num_islands):
        assert pathfinder.island_area(i) == ref.navigable_area(i)
        assert pathfinder.island_radius(i) == ref.island_radius(i)
        total += pathfinder.island_area(i)
        if pathfinder.island_area(i) > 0:
            # a tiled mesh picks a tile first (DQ.cpp:236-251): a small island needs many tries
            p = pathfinder.get_random_navigable_point(max_tries=2000, island_index=i)
            assert pathfinder.get_island(p) == i
            assert pathfinder.island_radius(p) == pathfinder.island_radius(i)  # test_nav.py:442-452
    assert abs(total - pathfinder.navigable_area) < 1e-2
    assert pathfinder.island_area(-1) == pathfinder.navigable_area


def test_topdown_map(pathfinder):  # test_nav.py:579-612 (shapes and agreement of the two views)
    lo, hi = pathfinder.get_bounds()
    height = float(lo[1]) + 0.2
    binary = pathfinder.get_topdown_view(0.1, height)
    islands = pathfinder.get_topdown_island_view(0.1, height)
    assert binary.dtype == bool and islands.dtype == np.int32 and binary.shape == islands.shape
    assert binary.any() and ((islands >= 0) == binary).all()


def test_greedy_follower_keys_radius_and_goal_cache(pathfinder):  # test_nav.py:702-830
    from habitat_sim_b200.nav import GreedyFollowerCodes, GreedyGeodesicFollower
    f = GreedyGeodesicFollower(pathfinder)
    assert f.action_mapping[GreedyFollowerCodes.FORWARD] == "move_forward"
    assert f.action_mapping[GreedyFollowerCodes.LEFT] == "turn_left"
    assert f.action_mapping[GreedyFollowerCodes.RIGHT] == "turn_right"
    assert f.action_mapping[GreedyFollowerCodes.STOP] is None
    assert abs(f.goal_radius - 0.75 * 0.25) < 1e-6
    g = GreedyGeodesicFollower(pathfinder, goal_radius=1.5, stop_key="STOP", forward_key="FWD", left_key="LT",
                               right_key="RT")
    assert g.goal_radius == 1.5 and g.action_mapping[GreedyFollowerCodes.STOP] == "STOP"
    assert g.action_mapping[GreedyFollowerCodes.FORWARD] == "FWD"
    start = np.asarray(pathfinder.get_random_navigable_point(max_tries=2000, island_index=0))
    # a goal a short walk away (a point near in xz may lie on another storey of the tiled building)
    for _ in range(50):
        goal = np.asarray(pathfinder.get_random_navigable_point_near(start, 3.0, max_tries=500, island_index=0))
        d = pathfinder.geodesic_distances(start[None], goal[None])[0] if np.isfinite(goal).all() else np.inf
        if 0.5 < d < 8.0:
            break
    assert 0.5 < d < 8.0
    rot = np.array([0.0, 0.0, 0.0, 1.0])
    f.next_action_along(rot, start, goal)
    assert f.last_goal is not None and np.allclose(f.last_goal, goal)
    f.reset()
    assert f.last_goal is None
    path = f.find_path(rot, start, goal)  # ends with the stop key (None), test_nav.py:831-845
    assert path[-1] is None and all(a in ("move_forward", "turn_left", "turn_right") for a in path[:-1])


def test_reused_multigoal_object_tracks_the_reference():
    """The stateful MultiGoalShortestPath (bounds carried over, goals projected once, trap T4) through
    the real PathFinder against the oracle's stateful twin."""
    from habitat_sim_b200.nav import MultiGoalShortestPath
    from oracle.ref import RefMultiGoal
    from conftest import beq, query_points
    name = "t_building"
    pf, ref = gpu_pathfinder(name), ref_pathfinder(name)
    lo, hi = ref.get_bounds()
    rng = np.random.default_rng(18)
    walk = query_points(name, 24, 71, jitter=0.05)
    walk[1:] = walk[:-1] + rng.normal(0, 0.4, (23, 3)).astype(np.float32) * np.float32([1, 0, 1])
    ends_a = query_points(name, 6, 72)
    ends_a[2, 1] += 3.0
    ends_b = query_points(name, 4, 73)
    ends_b[0] = (hi + 50).astype(np.float32)
    twin, mine = RefMultiGoal(ref), MultiGoalShortestPath()
    for ends, a, b in ((ends_a, 0, 12), (ends_b, 12, 24)):
        twin.set_ends(ends)
        mine.requested_ends = ends
        for s in walk[a:b]:
            ok, d, idx, pts = twin.find(s)
            mine.requested_start = s
            assert pf.find_path(mine) == ok and mine.closest_end_point_index == idx
            assert np.float32(mine.geodesic_distance) == np.float32(d) or (np.isinf(d) and np.isinf(mine.geodesic_distance))
            assert len(mine.points) == len(pts) and all(beq(x, y).all() for x, y in zip(mine.points, pts))

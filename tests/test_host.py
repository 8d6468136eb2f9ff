"""CPU tests of the product's host side: the C ABI library loads and exports everything
include/hbn.h declares; the host ingest reproduces the reference's finalised navmesh; and the
device query code, compiled for the host with a one-lane group (tests/hostemu), matches the
oracle bit for bit.  No CUDA call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, beq, hostemu, navmesh_image, query_points, ref_pathfinder

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


def P(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def test_library_exports_header_symbols():
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hbn.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|float|void|int64_t|const char\*)\s+(hbn_[a-z_]+)\(", header, re.M)))
    assert declared == sorted(_lib.SYMBOLS)
    so = _lib.build_library()
    lib = C.CDLL(so)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import HbnError, PathFinder
    pf = PathFinder(0)
    with pytest.raises(HbnError):
        pf.load_nav_mesh_bytes(navmesh_image("c1_room"))


def _emu_handle(name):
    emu = hostemu()
    img = navmesh_image(name)
    h = C.c_void_p(emu.emu_create(img, C.c_long(len(img))))
    assert h
    return emu, h


@pytest.mark.parametrize("name", ["c1_room", "c2_apartment", "c3_multiroom", "t_building"])
def test_ingest_matches_reference(name):
    """finalised tile blobs (links), poly refs, island ids, radii and areas"""
    emu, h = _emu_handle(name)
    pf = ref_pathfinder(name)
    for i, (ref, idx, blob) in enumerate(pf.tile_blobs()):
        n = emu.emu_tile_blob(h, i, None, 0)
        buf = C.create_string_buffer(n)
        emu.emu_tile_blob(h, i, buf, n)
        assert buf.raw == blob, f"tile {i} differs"
    refs, isl = pf.poly_islands()
    isl2 = np.zeros(len(refs), np.int32)
    refs2 = np.zeros(len(refs), np.uint32)
    assert emu.emu_poly_islands(h, P(isl2, i32p), P(refs2, u32p), C.c_long(len(refs))) == len(refs)
    assert (refs == refs2).all() and (isl == isl2).all()
    for i in range(pf.num_islands):
        assert pf.island_radius(i) == emu.emu_island_radius(h, i)
        assert pf.navigable_area(i) == emu.emu_island_area(h, i)
    # the total is summed in the reference's unordered_map iteration order (PF.cpp:1079-1083): exact
    assert np.float32(pf.navigable_area(-1)) == np.float32(emu.emu_island_area(h, -1))
    emu.emu_destroy(h)


@pytest.mark.parametrize("name", ["c1_room", "c2_apartment", "c3_multiroom", "t_building"])
def test_hostemu_queries_match_oracle(name):
    emu, h = _emu_handle(name)
    pf = ref_pathfinder(name)
    n = 1500
    pts = query_points(name, 2 * n, 21)
    lo, hi = pf.get_bounds()
    rng = np.random.default_rng(4)
    pts[:100] = rng.uniform(lo - 3, hi + 3, (100, 3)).astype(np.float32)
    pts[100:103] = np.nan
    # snap, also island restricted
    o_pts, o_refs, o_isl = pf.snap_batch(pts)
    e_pts = np.zeros_like(pts)
    e_refs = np.zeros(len(pts), np.uint32)
    e_isl = np.zeros(len(pts), np.int32)
    emu.emu_snap(h, P(pts, f32p), None, C.c_long(len(pts)), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p))
    assert (o_refs == e_refs).all() and (o_isl == e_isl).all() and beq(o_pts, e_pts).all()
    isl = rng.integers(0, pf.num_islands, 300).astype(np.int32)
    oi_pts, oi_refs = pf.snap_island_batch(pts[:300], isl)
    emu.emu_snap(h, P(pts, f32p), P(isl, i32p), C.c_long(300), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p))
    assert (oi_refs == e_refs[:300]).all() and beq(oi_pts, e_pts[:300]).all()
    # the candidate-list pipeline (hbn_snap.h) gives the same answers
    ncand = (C.c_long * 2)()
    emu.emu_snap_list(h, P(pts, f32p), None, C.c_long(len(pts)), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p),
                      ncand)
    assert (o_refs == e_refs).all() and (o_isl == e_isl).all() and beq(o_pts, e_pts).all()
    assert ncand[0] > len(pts) // 2 and ncand[1] <= ncand[0]
    emu.emu_snap_list(h, P(pts, f32p), P(isl, i32p), C.c_long(300), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p),
                      None)
    assert (oi_refs == e_refs[:300]).all() and beq(oi_pts, e_pts[:300]).all()
    # find_path, exact status mode and the small tier with fast fail
    st, en = pts[:n].copy(), pts[n:].copy()
    en[200:300] = st[200:300] + rng.normal(0, 0.01, (100, 3)).astype(np.float32)
    r = pf.find_path_raw_batch(st, en, max_pts=48)
    for cap, ff in ((2048, 0), (256, 1)):
        dist = np.zeros(n, np.float32)
        npts = np.zeros(n, np.int32)
        out_pts = np.full((n, 48, 3), np.nan, np.float32)
        corr = np.zeros((n, 256), np.uint32)
        info = np.zeros((n, 8), np.uint32)
        ovf = np.zeros(n, np.int32)
        emu.emu_find_path(h, P(st, f32p), P(en, f32p), C.c_long(n), cap, ff, P(dist, f32p), P(npts, i32p),
                          P(out_pts, f32p), 48, P(corr, u32p), P(info, u32p), P(ovf, i32p))
        ok = ovf == 0
        assert ok.mean() > 0.2
        assert beq(dist, r["dist"])[ok].all()
        assert (info[:, 0] == r["start_ref"]).all() and (info[:, 1] == r["end_ref"]).all()
        found = ((r["flags"] & 4) != 0) & ok
        assert (npts[found] == r["num_points"][found]).all()
        for i in np.nonzero(found)[0]:
            k = r["num_polys"][i]
            assert info[i, 4] == k and (corr[i, :k] == r["corridor"][i, :k]).all()
            m = min(r["num_points"][i], 48)
            assert beq(out_pts[i, :m], r["pts"][i, :m]).all()
        if ff == 0:  # Detour-exact: status words and partial corridors too
            assert (info[:, 2] == r["astar_status"]).all()
            for i in range(n):
                k = r["num_polys"][i]
                assert info[i, 4] == k and (corr[i, :k] == r["corridor"][i, :k]).all()
    # try_step
    from workloads.scenes import step_targets
    s2 = o_pts[:n].copy()
    s2[np.isnan(s2)] = 0
    t2 = step_targets(s2, 9, 0.25)
    t2[:300] = step_targets(s2[:300], 10, 2.5)
    for sliding in (1, 0):
        want = pf.try_step_batch(s2, t2, bool(sliding))
        got = np.zeros_like(s2)
        emu.emu_try_step(h, P(s2, f32p), P(t2, f32p), C.c_long(n), sliding, P(got, f32p))
        assert beq(want, got).all()
    # obstacle
    hp, hn, hd = pf.obstacle_batch(pts, 2.0)
    want = np.concatenate([hp, hn, hd[:, None]], 1)
    got = np.zeros((len(pts), 7), np.float32)
    ovf = np.zeros(len(pts), np.int32)
    emu.emu_obstacle(h, P(pts, f32p), C.c_long(len(pts)), C.c_float(2.0), 2048, P(got, f32p), P(ovf, i32p))
    assert ovf.sum() == 0 and beq(want, got).all()
    # ... and through the tiers of the device entry point (a query per thread with a small pool first)
    got = np.zeros((len(pts), 7), np.float32)
    tier = np.zeros(len(pts), np.int32)
    emu.emu_obstacle_tiers(h, P(pts, f32p), C.c_long(len(pts)), C.c_float(2.0), 8, P(got, f32p), P(tier, i32p))
    assert beq(want, got).all() and (tier == 0).mean() > 0.9
    # random points, plain and island restricted, same counter-based stream
    m = 600
    ri = np.full(m, -1, np.int32)
    ri[m // 2:] = rng.integers(0, pf.num_islands, m - m // 2)
    want_p, want_r = pf.random_points(m, 10, ri, mode=1, seed=99, query0=5)
    got_p = np.zeros((m, 3), np.float32)
    got_r = np.zeros(m, np.uint32)
    emu.emu_random_points(h, C.c_long(m), 10, P(ri, i32p), C.c_ulonglong(99), C.c_ulonglong(5), P(got_p, f32p),
                          P(got_r, u32p))
    assert (want_r == got_r).all() and beq(want_p, got_p).all()
    # get_random_navigable_point_near: circle filter (trap T6), plain and island restricted
    centers = np.ascontiguousarray(o_pts[200:200 + m])
    centers[np.isnan(centers)] = 0
    for radius, tries in ((1.5, 100), (0.4, 6)):
        want = pf.random_points_near(centers, radius, tries, ri, mode=1, seed=7, query0=11)
        got = np.zeros((m, 3), np.float32)
        emu.emu_random_points_near(h, C.c_long(m), P(centers, f32p), C.c_float(radius), tries, P(ri, i32p),
                                   C.c_ulonglong(7), C.c_ulonglong(11), P(got, f32p))
        assert beq(want, got).all()
        assert np.isfinite(want).all(axis=1).mean() > (0.5 if tries == 100 else 0.02)  # few tries: mostly NaN
    emu.emu_destroy(h)


@pytest.mark.parametrize("name", ["t_building", "c4_building"])
def test_hostemu_snap_list_on_multi_storey_tiles(name):
    """The narrowed BV walk + lower-bound pruning of hbn_snap.h on tiled multi-storey meshes: points
    on the mesh, between storeys, off the mesh and island-restricted must get the reference's poly and
    point; the pipeline must evaluate far fewer candidates than the reference visits."""
    emu, h = _emu_handle(name)
    pf = ref_pathfinder(name)
    n = 8000
    pts = query_points(name, n, 77, jitter=0.05)
    rng = np.random.default_rng(1)
    pts[:1600] += rng.normal(0, 0.6, (1600, 3)).astype(np.float32)
    pts[1600:2400, 1] += rng.uniform(-3, 3, 800).astype(np.float32)
    lo, hi = pf.get_bounds()
    pts[2400:2600] = rng.uniform(lo - 3, hi + 3, (200, 3)).astype(np.float32)
    o_pts, o_refs, o_isl = pf.snap_batch(pts, 8)
    e_pts = np.zeros_like(pts)
    e_refs = np.zeros(n, np.uint32)
    e_isl = np.zeros(n, np.int32)
    nc = (C.c_long * 2)()
    emu.emu_snap_list(h, P(pts, f32p), None, C.c_long(n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p), nc)
    assert (o_refs == e_refs).all() and (o_isl == e_isl).all() and beq(o_pts, e_pts).all()
    assert nc[1] <= nc[0] < 30 * n  # the reference's box collects ~50-60 candidates per point here
    # the lane-group kernel's walk with the same narrowed box (k_snap, small batches)
    emu.emu_snap(h, P(pts, f32p), None, C.c_long(n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p))
    assert (o_refs == e_refs).all() and (o_isl == e_isl).all() and beq(o_pts, e_pts).all()
    isl = rng.integers(0, pf.num_islands, n).astype(np.int32)
    isl[::3] = np.maximum(o_isl[::3], 0)
    oi_pts, oi_refs = pf.snap_island_batch(pts, isl)
    emu.emu_snap_list(h, P(pts, f32p), P(isl, i32p), C.c_long(n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p),
                      nc)
    assert (oi_refs == e_refs).all() and beq(oi_pts, e_pts).all()
    emu.emu_destroy(h)


@pytest.mark.parametrize("name", ["c1_room", "c3_multiroom", "t_building", "c4_building"])
def test_hostemu_snap_one_walk_and_two_walk_forms(name):
    """k_snap_walk keeps the candidates of the column walk (minimum radius) and walks again only when the
    radius comes out larger; the plain form walks the column, then the box.  Both must give the oracle's
    nearest poly, point and island, on-mesh points, jittered ones (walls, other storeys) and island-restricted."""
    emu, h = _emu_handle(name)
    pf = ref_pathfinder(name)
    rng = np.random.default_rng(12)
    n = 3000
    pts = query_points(name, n, 41, jitter=0.0)
    pts[n // 3:] += rng.normal(0, 0.35, (n - n // 3, 3)).astype(np.float32)
    pts = np.ascontiguousarray(pts, np.float32)
    isl = np.full(n, -1, np.int32)
    isl[::4] = rng.integers(0, pf.num_islands, len(isl[::4]))
    emu.emu_snap_one_walk.restype = C.c_long
    for islands in (None, isl):
        want_p, want_r, want_i = (pf.snap_island_batch(pts, islands) + (None,)) if islands is not None else pf.snap_batch(pts, 4)
        for one_walk in (1, 0):
            emu.emu_snap_one_walk(one_walk)
            e_pts = np.zeros_like(pts)
            e_refs = np.zeros(n, np.uint32)
            e_isl = np.zeros(n, np.int32)
            nc = (C.c_long * 2)()
            emu.emu_snap_list(h, P(pts, f32p), P(islands, i32p), C.c_long(n), P(e_pts, f32p), P(e_refs, u32p),
                              P(e_isl, i32p), nc)
            second = emu.emu_snap_one_walk(1)
            assert (want_r == e_refs).all() and beq(want_p, e_pts).all(), (one_walk, islands is not None)
            if want_i is not None:
                assert (want_i == e_isl).all()
            if one_walk:
                assert second < n  # most points are done after the first walk
    emu.emu_destroy(h)


@pytest.mark.parametrize("name", ["c2_apartment", "c3_multiroom", "t_building", "c4_building"])
def test_hostemu_lane_search_matches_oracle(name):
    """hbn_astar_lane.h (the per-lane state machine of k_astar_lane) against Detour's findPath:
    status words, corridor lengths and corridors, in Detour-exact mode and with fast fail; one
    lane slot serves all queries, so the table generation wraps several times."""
    emu, h = _emu_handle(name)
    pf = ref_pathfinder(name)
    n = 1200
    if name == "c4_building":
        from workloads.scenes import NavMeshGeom, pointnav_pairs
        st, en = pointnav_pairs(NavMeshGeom(navmesh_image(name)), n, 5)
    else:
        pts = query_points(name, 2 * n, 33)
        st, en = pts[:n].copy(), pts[n:].copy()
    r = pf.find_path_raw_batch(st, en)
    SUCCESS = 1 << 30
    searched_any = 0
    for ff, allc in ((0, 1), (1, 0), (1, 1)):
        corr = np.zeros((n, 256), np.uint32)
        info = np.zeros((n, 4), np.uint32)
        emu.emu_find_path_lane(h, P(st, f32p), P(en, f32p), C.c_long(n), ff, allc, P(corr, u32p), P(info, u32p))
        assert (info[:, 3] != 3).all(), "fault event"
        done = info[:, 3] == 1
        searched_any += int(done.sum())
        ok_ref = r["astar_status"] == SUCCESS
        ok_emu = info[:, 0] == SUCCESS
        assert (ok_ref[done] == ok_emu[done]).all()
        if ff == 0:
            assert (info[done, 0] == r["astar_status"][done]).all()
            assert (info[done, 2] == r["nodes_used"][done]).all()
        for i in np.nonzero(done)[0]:
            if not (allc and ff == 0) and not ok_ref[i]:
                continue
            k = r["num_polys"][i]
            assert min(info[i, 1], 256) == k, (i, info[i], k)
            assert (corr[i, :k] == r["corridor"][i, :k]).all(), i
    assert searched_any > n  # most queries do need a search
    emu.emu_destroy(h)


@pytest.mark.parametrize("ts,v", [(3, 1), (3, 9), (7, 10), (31, 10), (63, 9), (95, 9), (71, 10)])
@pytest.mark.parametrize("name", ["c3_multiroom", "c4_building"])
def test_hostemu_lane_search_variants(name, ts, v):
    """The variants of hbn_astar_lane.h (V = 9 / 10: no closed flag -- a re-improved node is looked for
    in the heap and pushed when it is not there) and small shared parts (ts = 3: nearly every heap
    level goes through the global-memory code) go through the reference's heap states: status
    words, node counts and corridors equal Detour's, in Detour-exact mode (all corridors) and with
    fast fail."""
    emu, h = _emu_handle(name)
    pf = ref_pathfinder(name)
    n = 800
    if name == "c4_building":
        from workloads.scenes import NavMeshGeom, pointnav_pairs
        st, en = pointnav_pairs(NavMeshGeom(navmesh_image(name)), n, 7)
    else:
        pts = query_points(name, 2 * n, 35)
        st, en = pts[:n].copy(), pts[n:].copy()
    r = pf.find_path_raw_batch(st, en)
    SUCCESS = 1 << 30
    emu.emu_find_path_lane_v.restype = C.c_int
    for ff, allc in ((0, 1), (1, 0)):
        corr = np.zeros((n, 256), np.uint32)
        info = np.zeros((n, 4), np.uint32)
        rc = emu.emu_find_path_lane_v(h, ts, v, P(st, f32p), P(en, f32p), C.c_long(n), ff, allc, P(corr, u32p),
                                      P(info, u32p))
        assert rc == 0
        assert (info[:, 3] != 3).all(), "fault event"
        done = info[:, 3] == 1
        assert done.sum() > n // 2
        ok_ref = r["astar_status"] == SUCCESS
        assert (ok_ref[done] == (info[done, 0] == SUCCESS)).all()
        if ff == 0:
            assert (info[done, 0] == r["astar_status"][done]).all()
            assert (info[done, 2] == r["nodes_used"][done]).all()
        for i in np.nonzero(done)[0]:
            if ff and not ok_ref[i]:
                continue
            k = r["num_polys"][i]
            assert min(info[i, 1], 256) == k, (i, info[i], k)
            assert (corr[i, :k] == r["corridor"][i, :k]).all(), i
    emu.emu_destroy(h)


@pytest.mark.parametrize("ts,v", [(3, 1), (7, 1), (31, 1), (39, 1), (47, 1), (55, 1), (63, 1), (71, 10), (95, 10)])
def test_hostemu_lane_heap_fuzz(ts, v):
    """The heap code of hbn_astar_lane.h (two-level storage, one or two levels per round trip)
    against dtNodeQueue restated plainly (DNode.cpp:156-200): random push / pop / modify sequences
    up to the pool size, keys from few values so that sibling and parent ties are everywhere,
    every heap entry compared after every operation."""
    emu = hostemu()
    emu.emu_lane_heap_fuzz.restype = C.c_long
    for levels in (3, 16, 1 << 20):
        for seed in range(6):
            assert emu.emu_lane_heap_fuzz(ts, v, seed, C.c_long(8000), levels) == 0, (levels, seed)


@pytest.mark.parametrize("ts,v", [(3, 1), (7, 1), (31, 1), (39, 1), (47, 1), (55, 1), (59, 1), (95, 1), (3, 9), (63, 8), (63, 9), (95, 9), (63, 10), (71, 10), (95, 10), (71, 110), (71, 310), (3, 310), (63, 101)])
@pytest.mark.parametrize("name", ["c3_multiroom", "c4_building"])
def test_hostemu_lane_variants_lockstep(name, ts, v):
    """Every variant in lock step with the shipped one: same heap entries, node count, best node
    and event after every step() of every query (stronger than equal corridors)."""
    emu, h = _emu_handle(name)
    n = 600
    if name == "c4_building":
        from workloads.scenes import NavMeshGeom, pointnav_pairs
        st, en = pointnav_pairs(NavMeshGeom(navmesh_image(name)), n, 9)
    else:
        pts = query_points(name, 2 * n, 37)
        st, en = pts[:n].copy(), pts[n:].copy()
    emu.emu_lane_lockstep.restype = C.c_long
    for ff in (0, 1):
        assert emu.emu_lane_lockstep(h, ts, v, P(st, f32p), P(en, f32p), C.c_long(n), ff) == 0
    emu.emu_destroy(h)


@pytest.mark.parametrize("name", ["c3_multiroom", "t_building"])
def test_hostemu_out_of_the_ordinary_inputs(name):
    """Regimes the seeded workloads do not reach: try_step from unsnapped starts towards targets
    0.05 to 30 m away and off in height (moveAlongSurface's 64-node pool and 48-slot queue run
    full, trap T9), wall distance with radii from 0 to 1e6, snaps around the edges of the
    +-(2, 4, 2) m pick box and the climb threshold.  Bit-exact against the oracle."""
    emu, h = _emu_handle(name)
    ref = ref_pathfinder(name)
    n = 1500
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
    t = np.ascontiguousarray(t)
    for sliding in (1, 0):
        want = ref.try_step_batch(s, t, bool(sliding), 4)
        got = np.zeros_like(s)
        emu.emu_try_step(h, P(s, f32p), P(t, f32p), C.c_long(n), sliding, P(got, f32p))
        assert beq(got, want).all(), sliding
    tiers_seen = set()
    for r in (0.0, 0.05, 7.5, 1e6):
        hp, hn, hd = ref.obstacle_batch(s, r, 4)
        out = np.zeros((n, 7), np.float32)
        ov = np.zeros(n, np.int32)
        emu.emu_obstacle(h, P(s, f32p), C.c_long(n), C.c_float(r), 2048, P(out, f32p), P(ov, i32p))
        assert not ov.any()
        assert beq(out[:, 6], hd).all() and beq(out[:, :3], hp).all() and beq(out[:, 3:6], hn).all(), r
        out = np.zeros((n, 7), np.float32)
        tier = np.zeros(n, np.int32)
        for small_cap in (8, 4):
            emu.emu_obstacle_tiers(h, P(s, f32p), C.c_long(n), C.c_float(r), small_cap, P(out, f32p), P(tier, i32p))
            assert beq(out[:, 6], hd).all() and beq(out[:, :3], hp).all() and beq(out[:, 3:6], hn).all(), r
            tiers_seen |= set(np.unique(tier).tolist())
    assert {0, 1} <= tiers_seen  # a 4-node pool overflows to the warp-per-query workspace
    p2 = s.copy()
    p2[:, 1] += rng.choice([-4.2, -4.0, -3.99, -1.0, -0.2, 0.19, 0.2, 0.21, 1.0, 2.9, 3.99, 4.0, 4.01],
                           size=n).astype(np.float32)
    p2[:, 0] += rng.choice([0, 1.99, 2.0, 2.01, -2.0], size=n).astype(np.float32)
    wp, wr, wi = ref.snap_batch(p2, 4)
    assert (wr == 0).any() and (wr != 0).any()
    e_pts = np.zeros_like(p2)
    e_refs = np.zeros(n, np.uint32)
    e_isl = np.zeros(n, np.int32)
    nc = (C.c_long * 2)()
    emu.emu_snap_list(h, P(p2, f32p), None, C.c_long(n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p), nc)
    assert (wr == e_refs).all() and (wi == e_isl).all() and beq(wp, e_pts).all()
    emu.emu_snap(h, P(p2, f32p), None, C.c_long(n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p))
    assert (wr == e_refs).all() and (wi == e_isl).all() and beq(wp, e_pts).all()
    emu.emu_destroy(h)


@pytest.mark.parametrize("g", [1, 9, 200])
def test_hostemu_multigoal_odd_goal_sets(g):
    """MultiGoalShortestPath with goal sets the workloads do not produce: one goal, 200 goals,
    duplicated goals (equal bounds: the order of the unstable std::sort decides), NaN goals, a goal
    equal to the start, rows where nothing is reachable -- with every pair searched and with the
    two-round pruning of the batched entry point."""
    name = "t_building"
    emu, h = _emu_handle(name)
    ref = ref_pathfinder(name)
    n = 120
    st = np.ascontiguousarray(query_points(name, n, 71 + g))
    en = query_points(name, n * g, 72 + g).reshape(n, g, 3).copy()
    en[::5, 0] = en[::5, -1]
    en[::7, g // 2] = np.nan
    en[::9, 0] = st[::9]
    en[3::11] = np.float32(1e4)
    en[4::13, : max(1, g // 3)] = en[4::13, -1:]
    en = np.ascontiguousarray(en)
    wd, wi, _, _ = ref.find_path_multigoal_batch(st, en, 0, 4)
    assert np.isfinite(wd).any() and np.isinf(wd).any()
    for pruned in (0, 1):
        dist = np.zeros(n, np.float32)
        idx = np.zeros(n, np.int32)
        searched = C.c_long()
        emu.emu_find_path_multigoal(h, P(st, f32p), P(en, f32p), C.c_long(n), g, P(dist, f32p), P(idx, i32p), pruned,
                                    C.byref(searched))
        assert beq(dist, wd).all() and (idx == wi).all(), pruned
    emu.emu_destroy(h)


@pytest.mark.parametrize("name", ["c2_apartment", "c3_multiroom", "t_building", "c4_building"])
def test_node_key_numbering(name, monkeypatch):
    """The flattener numbers the A* node keys along a space-filling curve (height layers, Morton
    order within a layer).  Whatever the order, the per-poly key ranges must tile [0, numKeys) and
    every link must carry a key of its neighbour's range; and the curve must put the two ends of
    a link closer together in the node table than poly order does (HBN_KEY_ORDER=0)."""
    emu = hostemu()
    img = navmesh_image(name)
    stats = {}
    for order in ("1", "0"):
        monkeypatch.setenv("HBN_KEY_ORDER", order)
        h = C.c_void_p(emu.emu_create(img, C.c_long(len(img))))
        out = (C.c_long * 4)()
        emu.emu_key_stats(h, out)
        emu.emu_destroy(h)
        assert out[1] == 1, "key ranges do not tile [0, numKeys)"
        stats[order] = (out[0], out[3] / max(1, out[2]))
    assert stats["1"][0] == stats["0"][0]
    assert stats["1"][1] <= stats["0"][1], stats
    if name in ("t_building", "c4_building"):  # tiled, multi-storey: poly order interleaves the storeys of a tile
        assert stats["1"][1] < 0.5 * stats["0"][1], stats


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_hostemu_fuzz_random_scenes(seed):
    """Fresh procedural scenes (not the cached benchmark ones): a seeded multi-room floor plan and a
    tiled two-storey building with ramps, built by the reference's Recast; the product's nearest-poly
    pipeline and lane search must reproduce the oracle on each."""
    from oracle.ref import RefPathFinder
    from workloads import meshgen
    from workloads.scenes import NavMeshGeom
    emu = hostemu()
    for gen, kw, tiled in ((meshgen.multi_room, dict(nx=4 + seed % 3, nz=3 + seed % 2, seed=seed), False),
                           (meshgen.building, dict(nx=5, nz=4, floors=2, ramps_per_floor=2, closed_rooms=1,
                                                   seed=seed), True)):
        v, t = gen(**kw)
        ref = RefPathFinder()
        assert ref.build_tiled(v, t, 128) if tiled else ref.build(v, t)
        img = ref.save_bytes()
        ref = RefPathFinder()
        assert ref.load_bytes(img)
        h = C.c_void_p(emu.emu_create(img, C.c_long(len(img))))
        assert h
        rng = np.random.default_rng(seed)
        n = 700
        geom = NavMeshGeom(img)
        pts = geom.sample(2 * n, rng) + rng.normal(0, 0.15, (2 * n, 3)).astype(np.float32)
        pts = pts.astype(np.float32)
        o_pts, o_refs, o_isl = ref.snap_batch(pts)
        e_pts = np.zeros_like(pts)
        e_refs = np.zeros(2 * n, np.uint32)
        e_isl = np.zeros(2 * n, np.int32)
        nc = (C.c_long * 2)()
        emu.emu_snap_list(h, P(pts, f32p), None, C.c_long(2 * n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p), nc)
        assert (o_refs == e_refs).all() and (o_isl == e_isl).all() and beq(o_pts, e_pts).all()
        emu.emu_snap(h, P(pts, f32p), None, C.c_long(2 * n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p))
        assert (o_refs == e_refs).all() and beq(o_pts, e_pts).all()
        st, en = pts[:n].copy(), pts[n:].copy()
        r = ref.find_path_raw_batch(st, en)
        corr = np.zeros((n, 256), np.uint32)
        info = np.zeros((n, 4), np.uint32)
        emu.emu_find_path_lane(h, P(st, f32p), P(en, f32p), C.c_long(n), 0, 1, P(corr, u32p), P(info, u32p))
        done = info[:, 3] == 1
        assert done.sum() > n // 4 and (info[:, 3] != 3).all()
        assert (info[done, 0] == r["astar_status"][done]).all()
        assert (info[done, 2] == r["nodes_used"][done]).all()
        for i in np.nonzero(done)[0]:
            k = r["num_polys"][i]
            assert min(info[i, 1], 256) == k and (corr[i, :k] == r["corridor"][i, :k]).all()
        # wall distance through the tiers of the device entry point (per-thread pool first), two radii
        for radius in (2.0, 30.0):
            hp, hn, hd = ref.obstacle_batch(st, radius, 4)
            out = np.zeros((n, 7), np.float32)
            tier = np.zeros(n, np.int32)
            emu.emu_obstacle_tiers(h, P(st, f32p), C.c_long(n), C.c_float(radius), 8, P(out, f32p), P(tier, i32p))
            assert beq(out[:, 6], hd).all() and beq(out[:, :3], hp).all() and beq(out[:, 3:6], hn).all(), radius
        emu.emu_destroy(h)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores only: one JSON line with the contract's keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "4000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "find_path_queries_per_sec"
    assert line["unit"] == "queries/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and 0.2 < line["found_fraction"] < 1.0


@pytest.mark.parametrize("name", ["ref_simple_room", "ref_stage_floor1"])
def test_hostemu_on_reference_test_scenes(name):
    """Golden vectors from the REFERENCE's own test scenes (data/test_assets/scenes/*.glb: a furnished
    room with several islands at different heights, an undulating floor whose heights come from the
    detail mesh; tests/golden/make_golden.py): the product's device code, host-emulated, must
    reproduce them."""
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    img = g["image"].tobytes()
    emu = hostemu()
    h = C.c_void_p(emu.emu_create(img, C.c_long(len(img))))
    assert h
    st, en = np.ascontiguousarray(g["starts"]), np.ascontiguousarray(g["ends"])
    n = len(st)
    e_pts = np.zeros_like(st)
    e_refs = np.zeros(n, np.uint32)
    e_isl = np.zeros(n, np.int32)
    nc = (C.c_long * 2)()
    for fn, extra in ((emu.emu_snap, ()), (emu.emu_snap_list, (nc,))):
        fn(h, P(st, f32p), None, C.c_long(n), P(e_pts, f32p), P(e_refs, u32p), P(e_isl, i32p), *extra)
        assert (e_refs == g["snap_refs"]).all() and (e_isl == g["snap_isl"]).all() and beq(e_pts, g["snap_pts"]).all()
    dist = np.zeros(n, np.float32)
    npts = np.zeros(n, np.int32)
    out_pts = np.full((n, 32, 3), np.nan, np.float32)
    corr = np.zeros((n, 256), np.uint32)
    info = np.zeros((n, 8), np.uint32)
    ovf = np.zeros(n, np.int32)
    emu.emu_find_path(h, P(st, f32p), P(en, f32p), C.c_long(n), 2048, 0, P(dist, f32p), P(npts, i32p),
                      P(out_pts, f32p), 32, P(corr, u32p), P(info, u32p), P(ovf, i32p))
    assert ovf.sum() == 0 and beq(dist, g["dist"]).all() and (info[:, 2] == g["astar_status"]).all()
    found = (g["flags"] & 4) != 0
    assert (npts[found] == g["num_points"][found]).all()
    for i in np.nonzero(found)[0]:
        m = min(g["num_points"][i], 32)
        assert beq(out_pts[i, :m], g["path_pts"][i, :m]).all()
    corr2 = np.zeros((n, 256), np.uint32)
    info2 = np.zeros((n, 4), np.uint32)
    emu.emu_find_path_lane(h, P(st, f32p), P(en, f32p), C.c_long(n), 0, 1, P(corr2, u32p), P(info2, u32p))
    done = info2[:, 3] == 1
    assert (info2[done, 0] == g["astar_status"][done]).all()
    for i in np.nonzero(done)[0]:
        k = min(int(g["num_polys"][i]), 64)
        assert info2[i, 1] == g["num_polys"][i] and (corr2[i, :k] == g["corridor"][i, :k]).all()
    for sliding, key in ((1, "step_sliding"), (0, "step_nosliding")):
        got = np.zeros_like(st)
        emu.emu_try_step(h, P(np.ascontiguousarray(g["snap_pts"]), f32p), P(np.ascontiguousarray(g["step_targets"]), f32p),
                         C.c_long(n), sliding, P(got, f32p))
        assert beq(got, g[key]).all()
    got = np.zeros((n, 7), np.float32)
    emu.emu_obstacle(h, P(st, f32p), C.c_long(n), C.c_float(2.0), 2048, P(got, f32p), P(ovf, i32p))
    assert beq(got[:, 6], g["hit_dist"]).all() and beq(got[:, :3], g["hit_pos"]).all()
    emu.emu_destroy(h)


def test_uniform_stream_definition_matches_oracle():
    from oracle import ref
    emu = hostemu()  # noqa: F841  (forces the build; the stream itself is checked via random points)
    assert 0.0 <= ref.uniform(1, 2, 3) <= 1.0
    assert ref.uniform(1, 2, 3) != ref.uniform(1, 2, 4)


def test_restated_std_sort_matches_libstdcxx():
    """stdSortOrder (hbn_query.h) against the oracle's real std::sort, same comparator as
    PathFinder.cpp:1544-1548: tie-heavy, sorted, reversed, organ-pipe and random keys at sizes on
    both sides of the insertion-sort threshold (16) and deep enough to hit the heapsort fallback."""
    from oracle.ref import std_sort_order
    emu = hostemu()
    rng = np.random.default_rng(8)
    cases = []
    for n in [1, 2, 3, 15, 16, 17, 31, 32, 33, 64, 100, 257, 1000]:
        cases.append(rng.random(n).astype(np.float32))
        cases.append(rng.integers(0, 3, n).astype(np.float32))          # many ties
        cases.append(np.sort(rng.random(n)).astype(np.float32))
        cases.append(np.sort(rng.random(n))[::-1].astype(np.float32).copy())
        cases.append(np.zeros(n, np.float32))
        k = np.arange(n, dtype=np.float32)
        cases.append(np.minimum(k, n - 1 - k))                              # organ pipe
        cases.append(np.where(np.arange(n) % 2 == 0, k, n - k).astype(np.float32))
    # median-of-3 killer sequence: drives introsort to its depth limit (heapsort fallback)
    for n in [64, 256, 1024]:
        k = n // 2
        a = np.zeros(n, np.float32)
        for i in range(k):
            if i % 2 == 0:
                a[i] = i + 1
            else:
                a[i] = k + i + (k % 2)
            a[k + i] = 2 * (i + 1)
        cases.append(a)
    for keys in cases:
        want = std_sort_order(keys)
        got = np.zeros(len(keys), np.int32)
        emu.emu_std_sort_order(P(np.ascontiguousarray(keys), f32p), C.c_int(len(keys)), P(got, i32p))
        assert (got == want).all(), (len(keys), keys[:8])


@pytest.mark.parametrize("name", ["c2_apartment", "t_building"])
def test_hostemu_multigoal_matches_oracle(name):
    """multiGoalSelect + restated sort vs the oracle's findPath(MultiGoalShortestPath&): duplicate
    goals (equal bounds), invalid goals, goals above the mesh (geodesic < L2 bound, so pruning
    is observable), invalid starts, a single goal."""
    emu, h = _emu_handle(name)
    pf = ref_pathfinder(name)
    rng = np.random.default_rng(9)
    n, g = 120, 9
    starts = query_points(name, n, 41)
    ends = query_points(name, n * g, 42).reshape(n, g, 3)
    ends[:, 3] = ends[:, 1]                       # duplicate goal -> equal bounds
    ends[:, 5, 1] += 3.0                          # far above: L2 bound exceeds the geodesic distance
    lo, hi = pf.get_bounds()
    ends[::7, 2] = (hi + 50).astype(np.float32)   # cannot be projected
    starts[5] = (hi + 50).astype(np.float32)
    ends[9] = (hi + 50).astype(np.float32)        # no valid goal at all
    ends[::5, 6] = starts[::5] + rng.normal(0, 0.3, (len(starts[::5]), 3)).astype(np.float32)
    for gg in (g, 1):
        e = np.ascontiguousarray(ends[:, :gg])
        wd, wi, _, _ = pf.find_path_multigoal_batch(starts, e)
        for pruned in (0, 1):
            gd = np.zeros(n, np.float32)
            gi = np.zeros(n, np.int32)
            emu.emu_find_path_multigoal(h, P(starts, f32p), P(e, f32p), C.c_long(n), C.c_int(gg), P(gd, f32p),
                                        P(gi, i32p), pruned, None)
            assert (gi == wi).all()
            assert beq(gd, wd).all()
        assert (wi >= 0).mean() > 0.5
    # many goals: the two-round pruning of the batched entry point must not change anything
    g2 = 40
    ends2 = query_points(name, n * g2, 43).reshape(n, g2, 3)
    ends2[:, 3] = ends2[:, 1]
    ends2[:, 5, 1] += 3.0
    ends2[:, 30, 1] += 2.5
    ends2[::7, 2] = (hi + 50).astype(np.float32)
    ends2[::5, 20] = starts[::5] + rng.normal(0, 0.3, (len(starts[::5]), 3)).astype(np.float32)
    wd, wi, _, _ = pf.find_path_multigoal_batch(starts, ends2)
    gd = np.zeros(n, np.float32)
    gi = np.zeros(n, np.int32)
    searched = C.c_long(0)
    emu.emu_find_path_multigoal(h, P(starts, f32p), P(ends2, f32p), C.c_long(n), C.c_int(g2), P(gd, f32p),
                                P(gi, i32p), 1, C.byref(searched))
    assert (gi == wi).all() and beq(gd, wd).all()
    assert searched.value < 0.6 * n * g2
    emu.emu_destroy(h)


def test_shard_slices_cover_and_balance():
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import shard_slices
    for n, world, g in ((0, 4, 1), (7, 8, 1), (1_000_000, 8, 1), (4096 * 64, 8, 64), (1024, 3, 1), (10, 1, 5)):
        sl = shard_slices(n, world, g)
        assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        sizes = [e - b for b, e in sl]
        assert all(s % g == 0 for s in sizes) and max(sizes) - min(sizes) <= g
    with pytest.raises(ValueError):
        shard_slices(10, 2, 4)


def test_sharding_gloo_world2():
    """The N > 1 path on CPU: two gloo ranks, each answering its slice (tests/_shard_worker.py)."""
    import socket
    import subprocess
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    hostemu()  # build once, not concurrently
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(os.path.dirname(__file__), "_shard_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for rank, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out[-2000:]
        assert f"rank {rank} ok" in out


class _EmuBackend:
    """PathFinder-like object over tests/hostemu (the product's query code compiled for the host) for
    the CPU test of the in-process multi-GPU dispatcher; thread-safe like a real handle is per call."""

    def __init__(self, device):
        self.device = device
        self._emu = hostemu()
        self._h = None
        self._seed = 0
        self.calls = []

    def load_nav_mesh_bytes(self, img):
        self._h = C.c_void_p(self._emu.emu_create(img, C.c_long(len(img))))
        return bool(self._h)

    def find_paths(self, st, en, max_points=0, corridors=False, exact_status=False):
        self.calls.append(len(st))
        d = np.zeros(len(st), np.float32)
        self._emu.emu_find_path(self._h, P(np.ascontiguousarray(st), f32p), P(np.ascontiguousarray(en), f32p),
                                C.c_long(len(st)), 2048, 1, P(d, f32p), None, None, 0, None, None, None)
        return dict(geodesic_distance=d)

    def try_steps(self, st, en, sliding=True):
        out = np.zeros((len(st), 3), np.float32)
        self._emu.emu_try_step(self._h, P(np.ascontiguousarray(st), f32p), P(np.ascontiguousarray(en), f32p),
                               C.c_long(len(st)), 1 if sliding else 0, P(out, f32p))
        return out

    def env_steps(self, p, t, g, sliding=True):
        pos = self.try_steps(p, t, sliding)
        return pos, self.find_paths(pos, g)["geodesic_distance"]

    def random_navigable_points(self, n, max_tries=10, island_index=-1, seed=None, query0=0):
        out = np.zeros((n, 3), np.float32)
        refs = np.zeros(n, np.uint32)
        self._emu.emu_random_points(self._h, C.c_long(n), max_tries, None, C.c_ulonglong(seed), C.c_ulonglong(query0),
                                    P(out, f32p), P(refs, u32p))
        return out, refs


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_multi_gpu_dispatcher_split_and_gather(world):
    """nav.MultiGpuPathFinder on the CPU: `world` backends (host emulation instead of CUDA handles), one
    worker thread each; a batch is cut into contiguous slices, every backend answers its slice, the
    results land in one array -- equal to the unsharded answer for any world, including batches smaller
    than the number of devices and the global-index random streams."""
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import MultiGpuPathFinder
    name = "c2_apartment"
    img = navmesh_image(name)
    single = _EmuBackend(0)
    assert single.load_nav_mesh_bytes(img)
    mg = MultiGpuPathFinder(devices=range(world), _factory=_EmuBackend)
    assert mg.load_nav_mesh_bytes(img) and mg.world == world
    for n in (1, 5, 1001):
        pts = query_points(name, 3 * n, 90 + n)
        st, en, gl = pts[:n].copy(), pts[n:2 * n].copy(), pts[2 * n:].copy()
        want = single.find_paths(st, en)["geodesic_distance"]
        assert beq(mg.geodesic_distances(st, en), want).all()
        assert beq(mg.try_steps(st, en), single.try_steps(st, en)).all()
        wp, wd = single.env_steps(st, en, gl)
        gp, gd = mg.env_steps(st, en, gl)
        assert beq(gp, wp).all() and beq(gd, wd).all()
        rp, rr = mg.random_navigable_points(n, seed=5, query0=40)
        sp, sr = single.random_navigable_points(n, seed=5, query0=40)
        assert beq(rp, sp).all() and (rr == sr).all()
    sizes = [pf.calls[-1] for pf in mg._pfs if pf.calls]
    assert max(sizes) - min(sizes) <= 1 or world > 1001
    mg.close()


def test_navmesh_settings_json_like_the_reference(tmp_path):
    """NavMeshSettings JSON form (PathFinder.cpp:68-93, io/JsonEspTypes.cpp:287-331): round trip as in
    the reference's tests/test_nav.py:678-695, and the reference's own fixture
    data/test_assets/test_navmeshsettings.json (its values restated here)."""
    import json
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import NavMeshSettings
    s = NavMeshSettings()
    s.set_defaults()
    s.agent_radius = 0.42
    path = str(tmp_path / "navmesh_settings.json")
    s.write_to_json(path)
    data = json.load(open(path))
    assert isinstance(data, dict) and len(data) == 16 and abs(data["agentRadius"] - 0.42) < 1e-6
    s2 = NavMeshSettings()
    s2.read_from_json(path)
    assert s == s2
    fixture = {"cellSize": 0.0123, "cellHeight": 0.234, "agentHeight": 1.2345, "agentRadius": 0.123,
               "agentMaxClimb": 0.234, "agentMaxSlope": 34.0, "regionMinSize": 23.0, "regionMergeSize": 25.0,
               "edgeMaxLen": 23.0, "edgeMaxError": 1.345, "vertsPerPoly": 9.0, "detailSampleDist": 9.0,
               "detailSampleMaxError": 2.0, "filterLowHangingObstacles": False, "filterLedgeSpans": False,
               "filterWalkableLowHeightSpans": False}
    json.dump(fixture, open(path, "w"))
    s3 = NavMeshSettings()
    s3.read_from_json(path)
    assert abs(s3.cell_size - 0.0123) < 1e-7 and abs(s3.agent_height - 1.2345) < 1e-6
    assert s3.verts_per_poly == 9.0 and s3.filter_ledge_spans is False and s3 != s
    s3.read_from_json(str(tmp_path / "missing.json"))  # logged, not raised
    assert abs(s3.edge_max_error - 1.345) < 1e-6


def test_pathfinder_not_loaded_like_the_reference():
    """tests/test_nav.py:95-99: a fresh PathFinder is not loaded and has no settings (no device needed);
    queries on it raise instead of answering from nowhere."""
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import HitRecord, MultiGoalShortestPath, PathFinder, ShortestPath
    pf = PathFinder()
    assert not pf.is_loaded and pf.nav_mesh_settings is None
    with pytest.raises(RuntimeError):
        pf.snap_point(np.zeros(3, np.float32))
    sp = ShortestPath()  # tests/test_nav.py:273-280
    sp.requested_start = np.array([0.0, 0.0, 0.0])
    sp.requested_end = np.array([1.0, 0.0, 1.0])
    assert np.allclose(sp.requested_start, [0, 0, 0]) and np.allclose(sp.requested_end, [1, 0, 1])
    assert len(sp.points) == 0
    mg = MultiGoalShortestPath()
    assert mg.closest_end_point_index == -1 and len(mg.points) == 0
    hr = HitRecord()  # tests/test_nav.py:229-234
    _ = hr.hit_pos, hr.hit_normal, hr.hit_dist


def test_multigoal_object_state_matches_reference_twin():
    """A REUSED MultiGoalShortestPath (PathFinder.cpp:95-123, :1470-1572; trap T4): bounds carried from
    call to call, goals projected once per assignment, validity flags only appended.  The drop-in's
    host logic (multigoal_find_path) runs here over oracle-backed primitives and must track the
    reference's stateful object call by call."""
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import MultiGoalShortestPath, multigoal_find_path, std_sort_order
    from oracle.ref import RefMultiGoal
    for name in ("c2_apartment", "t_building"):
        ref = ref_pathfinder(name)
        lo, hi = ref.get_bounds()

        def snap_refs(p):
            return ref.snap_batch(np.asarray(p, np.float32).reshape(-1, 3))[1]

        def find_paths(s, e, m):
            d, n, pts = ref.find_path_batch(s, e, m)
            return dict(geodesic_distance=d, num_points=n, points=pts)

        rng = np.random.default_rng(17)
        walk = query_points(name, 40, 61, jitter=0.05)
        walk[1:] = walk[:-1] + rng.normal(0, 0.4, (39, 3)).astype(np.float32) * np.float32([1, 0, 1])  # a moving agent
        ends_a = query_points(name, 7, 62)
        ends_a[2, 1] += 3.0                                 # bound above the geodesic distance
        ends_b = query_points(name, 5, 63)
        ends_b[0] = (hi + 50).astype(np.float32)            # invalid goal where the old list had a valid one
        ends_c = np.concatenate([query_points(name, 8, 64), ((hi + 50).astype(np.float32))[None]])
        twin = RefMultiGoal(ref)
        mine = MultiGoalShortestPath()
        for ends, lo_i, hi_i in ((ends_a, 0, 14), (ends_b, 14, 26), (ends_c, 26, 40)):
            twin.set_ends(ends)
            mine.requested_ends = ends
            for s in walk[lo_i:hi_i]:
                ok, d, idx, pts = twin.find(s)
                mine.requested_start = s
                got = multigoal_find_path(mine, snap_refs, find_paths, std_sort_order)
                assert got == ok and mine.closest_end_point_index == idx
                assert np.float32(mine.geodesic_distance) == np.float32(d) or (np.isinf(d) and np.isinf(mine.geodesic_distance))
                assert len(mine.points) == len(pts) and all(beq(a, b).all() for a, b in zip(mine.points, pts))
        # off-mesh start: nothing changes, False
        mine.requested_start = (hi + 80).astype(np.float32)
        assert not multigoal_find_path(mine, snap_refs, find_paths, std_sort_order)
        assert mine.closest_end_point_index == -1 and mine.points == []

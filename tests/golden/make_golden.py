"""Regenerates tests/golden/*.npz: outputs of the reference's own PathFinder.cpp + Detour (compiled in
place by oracle/Makefile, bound by oracle/ref.py) on seeded inputs.

Run here (the container with /root/reference): python tests/golden/make_golden.py
The MSET navmesh image each vector set belongs to is stored inside the .npz, so the vectors
stay valid even if the procedural generators change.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref import RefPathFinder  # noqa: E402
from workloads.scenes import NavMeshGeom, navmesh_bytes, step_targets  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


REF_SCENES = "/root/reference/data/test_assets/scenes"


def ref_asset_image(name: str) -> bytes:
    """Navmesh of one of the REFERENCE's own test scenes (data/test_assets/scenes/<name>.glb, read
    here only): triangles via workloads/glb.py, the stage config's up axis / scale applied
    (simple_room: up = +z; stage_floor1: scale 2), built by the reference's Recast with default
    NavMeshSettings and round-tripped through save -> load like every other scene."""
    from workloads.glb import load_triangles
    v, t = load_triangles(os.path.join(REF_SCENES, name + ".glb"))
    if name == "simple_room":
        v = np.stack([v[:, 0], v[:, 2], -v[:, 1]], 1).astype(np.float32)
    elif name == "stage_floor1":
        v = (v * 2.0).astype(np.float32)
    pf = RefPathFinder()
    assert pf.build(v, t), name
    return pf.save_bytes()


def make(name: str, n: int, seed: int, image: bytes | None = None):
    if image is None:
        image = navmesh_bytes(name, cache=False)
    pf = RefPathFinder()
    assert pf.load_bytes(image)
    geom = NavMeshGeom(image)
    rng = np.random.default_rng(seed)
    pts = (geom.sample(2 * n, rng) + rng.normal(0, 0.2, (2 * n, 3))).astype(np.float32)
    starts, ends = pts[:n].copy(), pts[n:].copy()
    ends[: n // 10] = starts[: n // 10] + rng.normal(0, 0.01, (n // 10, 3)).astype(np.float32)
    lo, hi = pf.get_bounds()
    starts[-5:] = (hi + 50.0).astype(np.float32)  # off-mesh
    snap_pts, snap_refs, snap_isl = pf.snap_batch(starts)
    raw = pf.find_path_raw_batch(starts, ends, max_pts=32)
    tgt = step_targets(snap_pts, seed + 1, 0.25)
    tgt[: n // 4] = step_targets(snap_pts[: n // 4], seed + 2, 2.0)
    step_s = pf.try_step_batch(snap_pts, tgt, True)
    step_n = pf.try_step_batch(snap_pts, tgt, False)
    hp, hn, hd = pf.obstacle_batch(starts, 2.0)
    isl = np.full(n, -1, np.int32)
    isl[n // 2:] = rng.integers(0, pf.num_islands, n - n // 2)
    rp, rr = pf.random_points(n, 10, isl, mode=1, seed=seed, query0=1000)
    g = 8
    mg_ends = (geom.sample(n // 4 * g, rng) + rng.normal(0, 0.2, (n // 4 * g, 3))).astype(np.float32).reshape(n // 4, g, 3)
    mg_d, mg_i, mg_n, _ = pf.find_path_multigoal_batch(starts[: n // 4], mg_ends)
    refs_all, isl_all = pf.poly_islands()
    # the reference's own top-down views and navmesh geometry (PF.cpp:1833-1896, :1898-1944)
    td_h = float(snap_pts[np.isfinite(snap_pts[:, 0])][0, 1])
    td = pf.topdown_view(0.25, td_h)
    tdi = pf.topdown_view(0.25, td_h, islands=True)
    nv_all, _ = pf.navmesh_vertices(-1)
    nv_last, _ = pf.navmesh_vertices(pf.num_islands - 1)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), image=np.frombuffer(image, np.uint8), starts=starts, ends=ends,
        snap_pts=snap_pts, snap_refs=snap_refs, snap_isl=snap_isl, dist=raw["dist"],
        corridor=raw["corridor"][:, :64], num_polys=raw["num_polys"], num_points=raw["num_points"],
        path_pts=raw["pts"], flags=raw["flags"], astar_status=raw["astar_status"],
        step_targets=tgt, step_sliding=step_s, step_nosliding=step_n, hit_pos=hp, hit_normal=hn,
        hit_dist=hd, rand_islands=isl, rand_pts=rp, rand_refs=rr, mg_ends=mg_ends, mg_dist=mg_d,
        mg_idx=mg_i, mg_npts=mg_n, poly_refs=refs_all, poly_islands=isl_all,
        topdown_height=np.float32(td_h), topdown=td, topdown_islands=tdi, nav_verts_all=nv_all,
        nav_verts_last_island=nv_last,
        seed=np.int32(seed), num_islands=np.int32(pf.num_islands), area=np.float32(pf.navigable_area()),
        island_radius=np.array([pf.island_radius(i) for i in range(pf.num_islands)], np.float32),
        island_area=np.array([pf.navigable_area(i) for i in range(pf.num_islands)], np.float32))
    print(name, "ok", n)


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, n, seed in (("c1_room", 200, 11), ("c2_apartment", 400, 12), ("t_building", 400, 13)):
        if not only or name in only:
            make(name, n, seed)
    # non-procedural geometry: the reference's own test scenes
    for asset, n, seed in (("simple_room", 400, 14), ("stage_floor1", 400, 15)):
        if not only or "ref_" + asset in only:
            make("ref_" + asset, n, seed, ref_asset_image(asset))

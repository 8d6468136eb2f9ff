import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


_images = {}
_refs = {}


def navmesh_image(name: str) -> bytes:
    if name not in _images:
        from workloads.scenes import navmesh_bytes
        _images[name] = navmesh_bytes(name)
    return _images[name]


def ref_pathfinder(name: str):
    """Oracle PathFinder loaded from the scene's MSET image."""
    if name not in _refs:
        from oracle.ref import RefPathFinder
        pf = RefPathFinder()
        assert pf.load_bytes(navmesh_image(name))
        _refs[name] = pf
    return _refs[name]


def gpu_pathfinder(name: str):
    import habitat_sim_b200  # noqa: F401
    from habitat_sim_b200.nav import PathFinder
    pf = PathFinder(0)
    assert pf.load_nav_mesh_bytes(navmesh_image(name))
    return pf


def beq(a, b):
    """bitwise float equality, NaN == NaN"""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def query_points(name: str, n: int, seed: int, jitter: float = 0.2):
    """seeded on/near/off-mesh points for a scene"""
    from workloads.scenes import NavMeshGeom
    geom = NavMeshGeom(navmesh_image(name))
    rng = np.random.default_rng(seed)
    p = geom.sample(n, rng)
    p += rng.normal(0, jitter, p.shape).astype(np.float32)
    return p.astype(np.float32)


_emu = None


def hostemu():
    """tests/hostemu: the device query code compiled for the host with a one-lane group."""
    global _emu
    if _emu is None:
        out = os.path.join(ROOT, "tests", "hostemu", "_build")
        so = os.path.join(out, "libhbn_hostemu.so")
        src = [os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp"),
               os.path.join(ROOT, "habitat-sim_b200", "csrc", "hbn_host.cpp")]
        deps = src + [os.path.join(ROOT, "habitat-sim_b200", "csrc", f)
                      for f in ("hbn_query.h", "hbn_math.h", "hbn_types.h", "hbn_host.h", "hbn_astar_lane.h", "hbn_snap.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            os.makedirs(out, exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared",
                                   "-I", os.path.join(ROOT, "habitat-sim_b200", "csrc")] + src + ["-o", so])
        _emu = C.CDLL(so)
        _emu.emu_create.restype = C.c_void_p
        _emu.emu_island_radius.restype = C.c_float
        _emu.emu_island_area.restype = C.c_float
    return _emu

"""ctypes binding of libhbn.so (include/hbn.h).  Fails loudly when the CUDA library is
missing -- there is deliberately no fallback."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# HBN_LIBRARY: another build of the same library (tools only: csrc/Makefile DIAG=1)
_SO = os.environ.get("HBN_LIBRARY") or os.path.join(_HERE, "lib", "libhbn.so")
_lib = None

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)

HBN_OK = 0
HBN_ERR_NO_AREA = 5
HBN_FP_EXACT_STATUS = 1
HBN_FP_COUNT_WORK = 2


class HbnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hbn error {code}: {msg}")
        self.code = code


class FollowerParams(C.Structure):
    _fields_ = [("goal_dist", C.c_float), ("forward_amount", C.c_float), ("sin_half_turn", C.c_double),
                ("cos_half_turn", C.c_double), ("n_steps", C.c_int32), ("allow_sliding", C.c_int32)]


class NavMeshInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("num_tiles", C.c_int32), ("num_polys", C.c_int32),
                ("num_links", C.c_int32), ("num_bv_nodes", C.c_int32), ("num_islands", C.c_int32),
                ("poly_bits", C.c_int32), ("tile_bits", C.c_int32), ("salt_bits", C.c_int32),
                ("has_settings", C.c_int32), ("bounds_min", C.c_float * 3),
                ("bounds_max", C.c_float * 3), ("navigable_area", C.c_float),
                ("device_bytes", C.c_int64)]


class TileBlob(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_int32), ("tile_ref", C.c_uint32)]


def library_path() -> str:
    return _SO


def build_library(force: bool = False) -> str:
    """Compile libhbn.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc")])
    return _SO


# every symbol include/hbn.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "hbn_last_error", "hbn_device_count", "hbn_navmesh_create_from_mset",
    "hbn_navmesh_create_from_tiles", "hbn_navmesh_destroy", "hbn_navmesh_get_info",
    "hbn_navmesh_island_info", "hbn_navmesh_get_settings", "hbn_navmesh_launch_count",
    "hbn_navmesh_triangles", "hbn_navmesh_work_counters", "hbn_navmesh_set_profiling",
    "hbn_navmesh_phase_times", "hbn_snap_point_dev", "hbn_is_navigable_dev", "hbn_find_path_dev",
    "hbn_find_path_multigoal_dev", "hbn_try_step_dev", "hbn_closest_obstacle_dev",
    "hbn_random_points_dev", "hbn_uniform", "hbn_snap_point", "hbn_is_navigable",
    "hbn_find_path", "hbn_find_path_multigoal", "hbn_try_step", "hbn_closest_obstacle",
    "hbn_random_points", "hbn_random_points_near_dev", "hbn_random_points_near", "hbn_std_sort_order",
    "hbn_navmesh_set_settings", "hbn_navmesh_save_mset", "hbn_navmesh_set_option", "hbn_navmesh_reserve",
    "hbn_navmesh_scratch_bytes", "hbn_env_step_dev", "hbn_env_step", "hbn_navmesh_set_bounds",
    "hbn_follower_best_prims_dev", "hbn_follower_best_prims",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise HbnError(-1, f"{_SO} is missing: build it with habitat_sim_b200.build_library() "
                               "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        l = C.CDLL(_SO)
        l.hbn_last_error.restype = C.c_char_p
        l.hbn_uniform.restype = C.c_float
        l.hbn_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        l.hbn_std_sort_order.restype = None
        l.hbn_std_sort_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.hbn_navmesh_launch_count.restype = C.c_int64
        l.hbn_navmesh_launch_count.argtypes = [C.c_void_p]
        l.hbn_navmesh_work_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        l.hbn_navmesh_set_profiling.argtypes = [C.c_void_p, C.c_int]
        l.hbn_navmesh_phase_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.hbn_navmesh_triangles.restype = C.c_int64
        l.hbn_navmesh_triangles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        l.hbn_navmesh_create_from_mset.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        l.hbn_navmesh_create_from_tiles.argtypes = [C.POINTER(TileBlob), C.c_int, f32p, C.c_int, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                    C.POINTER(C.c_void_p)]
        l.hbn_navmesh_destroy.argtypes = [C.c_void_p]
        l.hbn_navmesh_destroy.restype = None
        l.hbn_navmesh_get_info.argtypes = [C.c_void_p, C.POINTER(NavMeshInfo)]
        l.hbn_navmesh_island_info.argtypes = [C.c_void_p, C.c_int, f32p, f32p]
        l.hbn_navmesh_get_settings.argtypes = [C.c_void_p, C.c_void_p]
        l.hbn_navmesh_set_settings.argtypes = [C.c_void_p, C.c_void_p]
        l.hbn_navmesh_save_mset.restype = C.c_int64
        l.hbn_navmesh_save_mset.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        l.hbn_navmesh_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        l.hbn_navmesh_set_bounds.argtypes = [C.c_void_p, C.c_void_p]
        l.hbn_navmesh_reserve.argtypes = [C.c_void_p, C.c_int64]
        l.hbn_navmesh_scratch_bytes.restype = C.c_int64
        l.hbn_navmesh_scratch_bytes.argtypes = [C.c_void_p]
        vp = C.c_void_p
        for suffix, extra in (("_dev", [vp]), ("", [])):
            getattr(l, "hbn_follower_best_prims" + suffix).argtypes = [vp, vp, vp, vp, C.c_int64,
                                                                      C.POINTER(FollowerParams), vp, vp] + extra
            getattr(l, "hbn_env_step" + suffix).argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int, vp, vp] + extra
            getattr(l, "hbn_snap_point" + suffix).argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp] + extra
            getattr(l, "hbn_is_navigable" + suffix).argtypes = [vp, vp, C.c_int64, C.c_float, vp] + extra
            getattr(l, "hbn_find_path" + suffix).argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp, C.c_int,
                                                            vp, vp, vp, C.c_int] + extra
            getattr(l, "hbn_find_path_multigoal" + suffix).argtypes = [vp, vp, vp, C.c_int64, C.c_int, vp,
                                                                      vp, vp, vp, C.c_int] + extra
            getattr(l, "hbn_try_step" + suffix).argtypes = [vp, vp, vp, C.c_int64, C.c_int, vp] + extra
            getattr(l, "hbn_closest_obstacle" + suffix).argtypes = [vp, vp, C.c_int64, C.c_float, vp, vp,
                                                                   vp] + extra
            getattr(l, "hbn_random_points" + suffix).argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int64, vp,
                                                                C.c_int, vp, vp] + extra
            getattr(l, "hbn_random_points_near" + suffix).argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int64, vp,
                                                                     C.c_float, vp, C.c_int, vp] + extra
        _lib = l
    return _lib


def check(rc: int):
    if rc != HBN_OK:
        raise HbnError(rc, (lib().hbn_last_error() or b"").decode())

"""Loads the CPU oracle (oracle/mppi_oracle.cpp) for the tests.  TEST INFRASTRUCTURE: nothing under
mpcholonavigation_b200/ imports this."""
import ctypes as C
import os
import subprocess

from mpcholonavigation_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

EXTRA = {
    "get_state": (C.c_int, [abi.H] + [abi.f32p] * 6),
    "set_wide_reductions": (C.c_int, [abi.H, C.c_int32]),
    "get_counters": (C.c_int, [abi.H, C.POINTER(C.c_uint64)]),
    "set_iteration_controls": (C.c_int, [abi.H, C.c_int32, abi.f32p, abi.f32p, abi.f32p]),
    "philox4x32_10": (None, [abi.u32p, abi.u32p, abi.u32p]),
    "det_sincosf": (None, [C.c_float, abi.f32p, abi.f32p]),
    "normalize_angle": (C.c_double, [C.c_double]),
    "utils_normalize_angles": (C.c_double, [C.c_float]),
    "utils_shortest_angular_distance": (C.c_double, [C.c_float, C.c_float]),
    "pose_point_angle": (C.c_float, [C.c_double] * 5 + [C.c_int32]),
    "within_tolerance_checker": (C.c_int32, [C.c_double] * 5),
    "within_tolerance": (C.c_int32, [C.c_float] + [C.c_double] * 4),
    "get_yaw": (C.c_double, [C.c_double] * 4),
    "find_closest_path_pt": (C.c_size_t, [abi.f32p, C.c_size_t, C.c_float, C.c_size_t]),
    "savitsky_golay": (None, [abi.f32p, abi.f32p, abi.f32p, C.c_int32, abi.f32p, C.c_int32]),
    "find_path_costs": (None, [C.POINTER(abi.CycleIn), C.c_int32, abi.u8p]),
    "inflation_compute_cost": (C.c_uint8, [C.c_double] * 4),
    "footprint_cost_at_pose": (C.c_double, [C.POINTER(abi.Costmap), C.POINTER(abi.RobotDesc)] + [C.c_double] * 3),
    "world_to_map": (C.c_int32, [C.POINTER(abi.Costmap), C.c_double, C.c_double]),
}


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def load(fast=False):
    name = "libmppi_oracle_fast.so" if fast else "libmppi_oracle.so"
    path = os.path.join(ORACLE_DIR, "_build", name)
    src = os.path.join(ORACLE_DIR, "mppi_oracle.cpp")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build()
    lib = C.CDLL(path)
    fns = abi.bind(lib, "oracle_", EXTRA)
    fns["_lib"] = lib
    return fns

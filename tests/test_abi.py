"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol include/mppi_b200.h
declares, and the ctypes mirror matches the C structs byte for byte.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from mpcholonavigation_b200 import _abi as abi
from mpcholonavigation_b200.api import PRODUCT_LIB, MppiError, load_product, make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mppi_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(PRODUCT_LIB), "run __graft_entry__.build() first"
    out = subprocess.run(["nm", "-D", "--defined-only", PRODUCT_LIB], capture_output=True, text=True, check=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    fns = load_product()
    assert fns["abi_version"]() == 1
    for name in list(abi.SIGNATURES) + list(abi.PRODUCT_ONLY):
        assert name in fns


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", PRODUCT_LIB], capture_output=True, text=True)
    assert "sm_100a" in out.stdout, out.stdout + out.stderr


def test_ctypes_layout_matches_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "mppi_b200.h"
#define S(T) printf(#T " %zu\n", sizeof(T))
#define O(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  S(mppi_config); O(mppi_config, seed); O(mppi_config, device); O(mppi_config, shard_offset); O(mppi_config, shard_total);
  S(mppi_critic_desc); O(mppi_critic_desc, deadband_velocities); O(mppi_critic_desc, inflation_radius);
  S(mppi_robot_desc); O(mppi_robot_desc, footprint_y); O(mppi_robot_desc, inflation_cost_scaling_factor); O(mppi_robot_desc, track_unknown);
  S(mppi_costmap); O(mppi_costmap, resolution);
  S(mppi_cycle_in); O(mppi_cycle_in, path_size); O(mppi_cycle_in, path_x); O(mppi_cycle_in, costmap);
  S(mppi_cycle_out); O(mppi_cycle_out, fail_flag); O(mppi_cycle_out, device_ms);
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "layout.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "layout")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        lines = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    got = dict(line.rsplit(" ", 1) for line in lines if line)
    mirror = {"mppi_config": abi.Config, "mppi_critic_desc": abi.CriticDesc, "mppi_robot_desc": abi.RobotDesc,
              "mppi_costmap": abi.Costmap, "mppi_cycle_in": abi.CycleIn, "mppi_cycle_out": abi.CycleOut}
    for key, val in got.items():
        if "." in key:
            t, f = key.split(".")
            assert getattr(mirror[t], f).offset == int(val), key
        else:
            assert C.sizeof(mirror[key]) == int(val), key


def test_defaults_are_the_reference_defaults():
    """optimizer.cpp:69-84 and each critic's initialize()"""
    fns = load_product()
    cfg = make_config(fns)
    assert (cfg.batch_size, cfg.time_steps, cfg.iteration_count) == (1000, 56, 1)
    assert cfg.model_dt == pytest.approx(0.05) and cfg.temperature == pytest.approx(0.3) and cfg.gamma == pytest.approx(0.015)
    assert (cfg.vx_max, cfg.vx_min, cfg.vy_max, cfg.wz_max) == pytest.approx((0.5, -0.35, 0.5, 1.9))
    assert (cfg.vx_std, cfg.vy_std, cfg.wz_std) == pytest.approx((0.2, 0.2, 0.4))
    assert cfg.motion_model == abi.MODEL_DIFF_DRIVE and cfg.regenerate_noises == 0
    expect = {"ConstraintCritic": 4.0, "CostCritic": 3.81, "GoalCritic": 5.0, "GoalAngleCritic": 3.0, "PathAlignCritic": 10.0,
              "PathAlignLegacyCritic": 10.0, "PathAngleCritic": 2.0, "PathFollowCritic": 5.0, "PreferForwardCritic": 5.0,
              "TwirlingCritic": 10.0, "VelocityDeadbandCritic": 35.0}
    for name, w in expect.items():
        d = abi.CriticDesc()
        fns["critic_default"](abi.CRITIC_KINDS[name], C.byref(d))
        assert d.enabled == 1 and d.cost_power == 1 and d.cost_weight == pytest.approx(w), name
    d = abi.CriticDesc()
    fns["critic_default"](abi.CRITIC_KINDS["ObstaclesCritic"], C.byref(d))
    assert (d.repulsion_weight, d.critical_weight, d.collision_cost, d.collision_margin_distance, d.near_goal_distance,
            d.cost_scaling_factor, d.inflation_radius) == pytest.approx((1.5, 20.0, 10000.0, 0.10, 0.5, 10.0, 0.55))


def test_product_and_oracle_defaults_agree(oracle_fns):
    fns = load_product()
    a, b = abi.Config(), abi.Config()
    fns["config_default"](C.byref(a))
    oracle_fns["config_default"](C.byref(b))
    assert bytes(a) == bytes(b)
    for kind in range(12):
        da, db = abi.CriticDesc(), abi.CriticDesc()
        fns["critic_default"](kind, C.byref(da))
        oracle_fns["critic_default"](kind, C.byref(db))
        assert bytes(da) == bytes(db), kind


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU mppi_create must fail loudly with MPPI_E_CUDA (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mpcholonavigation_b200 import Engine
    with pytest.raises(MppiError) as ei:
        Engine(load_product(), batch_size=32, time_steps=8)
    assert "status 2" in str(ei.value)

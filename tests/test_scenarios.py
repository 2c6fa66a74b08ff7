"""The synthetic workloads bench.py times and the parity tests check (mpcholonavigation_b200/scenarios.py): their
inputs must not drift between rounds (the numbers under profiles/ are only comparable on identical maps), and the boxed-in
variants must keep doing what they exist for - sending the costmap critics down the footprint branch in every cycle."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from mpcholonavigation_b200 import Engine, scenarios


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name,make,costmap_sha", [
    ("configs[1]", lambda: scenarios.config1(), "aa40ae632473d4a27b91ec109094ea11d5297ef9"),
    ("configs[2]", lambda: scenarios.config3(batch=64), "fe2a64864e8d32ddc5ee6e81e82dd528c9b4f4c5"),
    ("configs[3]", lambda: scenarios.config4(batch=64), "958b755648873fa1f4282a6a6ffff8766ec945d8"),
    ("configs[4] robot 7", lambda: scenarios.config5_robot(7), "0a4d59dffd4cb4ade000d3b7f8c4e04fd9e51e5d"),
    ("configs[2] boxed in", lambda: scenarios.config3(batch=64, dense=True), "587b5ae92ce347e5047ec0976be127a80c27b7bb"),
    ("configs[1] boxed in", lambda: scenarios.config1(footprint="rectangle", ring=0.47), "82425f4e765c49b7416b41b4cbd9259b773dde27"),
])
def test_benchmark_maps_do_not_drift(name, make, costmap_sha):
    assert _sha(make().cycle.costmap) == costmap_sha, name


def test_injected_noise_is_the_seeded_set():
    assert _sha(scenarios.config1().noise()[0])[:12] == "b558fab98767"
    assert _sha(scenarios.config3(batch=64).noise()[0])[:12] == "eae0d3b4c01c"


@pytest.mark.parametrize("make,critic,lo", [
    (lambda: scenarios.config3(batch=1024, dense=True), "ObstaclesCritic", 0.08),
    (lambda: scenarios.config1(batch=512, footprint="rectangle", ring=0.47), "CostCritic", 0.08),
])
def test_boxed_in_variants_take_the_footprint_branch_every_cycle(oracle_fns, make, critic, lo):
    """counted by the CPU oracle (oracle_get_counters): footprint checks / poses visited, cold and warm-started"""
    sc = make()
    e = Engine(oracle_fns, **sc.cfg)
    e.set_robot(sc.robot)
    e.set_critics(sc.critics)
    e.set_noise(*sc.noise())
    c = (C.c_uint64 * 4)()
    for cycle in range(12):
        r = e.optimize(sc.cycle)
        oracle_fns["get_counters"](e.h, c)
        visited, checks = (c[0], c[1]) if critic == "CostCritic" else (c[2], c[3])
        assert visited > 0 and checks > lo * visited, (cycle, list(c))
        assert not r.fail_flag
    e.close()


def test_literal_config3_hardly_ever_takes_it(oracle_fns):
    """SURVEY 8d's literal geometry: a fraction of a percent of the first cycle's poses, none once the sequence has moved"""
    sc = scenarios.config3(batch=2048)
    e = Engine(oracle_fns, **sc.cfg)
    e.set_robot(sc.robot)
    e.set_critics(sc.critics)
    e.set_noise(*sc.noise())
    c = (C.c_uint64 * 4)()
    e.optimize(sc.cycle)
    oracle_fns["get_counters"](e.h, c)
    assert 0 < c[3] < 0.01 * c[2]
    for _ in range(15):
        e.optimize(sc.cycle)
    oracle_fns["get_counters"](e.h, c)
    assert c[3] == 0
    e.close()

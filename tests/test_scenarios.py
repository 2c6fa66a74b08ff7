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


def test_warp_collection_numbering_model():
    """The numbering the stream rollout uses to collect the footprint checks of a chunk over the warp (mppi_kernels.cuh,
    phase F2): item (lane, pose u) gets number total_before(u) + popc(ballot(u) & lanes_below); round r hands the items
    32 r .. 32 r + 31 to lanes 0 .. 31 through one slot each and every owner reads its own result back.  Emulated lane by lane
    for random need masks: every needed item is checked exactly once, with its own pose, and the owner gets that result."""
    rng = np.random.default_rng(5)
    chunk = 4
    for trial in range(300):
        density = rng.choice([0.0, 0.02, 0.14, 0.5, 1.0])
        need = rng.random((32, chunk)) < density            # [lane][u]
        pose = rng.integers(1, 1 << 30, size=(32, chunk))   # stands for (x, y, yaw)
        check = lambda p: int(p) ^ 0x5A5A5A                 # the pure function of the pose
        ballots = [sum(1 << l for l in range(32) if need[l, u]) for u in range(chunk)]
        j = np.full((32, chunk), -1)
        total = 0
        for u in range(chunk):
            for l in range(32):
                if need[l, u]:
                    j[l, u] = total + bin(ballots[u] & ((1 << l) - 1)).count("1")
            total += bin(ballots[u]).count("1")
        assert total == int(need.sum())
        got = np.full((32, chunk), -1)
        checked = 0
        for r0 in range(0, total, 32):
            slot = [None] * 32
            for l in range(32):
                for u in range(chunk):
                    k = j[l, u] - r0
                    if 0 <= k < 32:
                        assert slot[k] is None
                        slot[k] = pose[l, u]
            res = [check(slot[l]) if r0 + l < total else 0 for l in range(32)]
            checked += sum(1 for l in range(32) if r0 + l < total)
            for l in range(32):
                for u in range(chunk):
                    k = j[l, u] - r0
                    if 0 <= k < 32:
                        got[l, u] = res[k]
        assert checked == total
        for l in range(32):
            for u in range(chunk):
                assert got[l, u] == (check(pose[l, u]) if need[l, u] else -1)

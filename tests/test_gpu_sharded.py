"""The sharded path of the CUDA library on ONE GPU: several handles of one process are shards of one problem
(mppi_optimize_sharded: same two exchanges as the NCCL path).  Checked against a single handle and the oracle."""
import numpy as np
import pytest

from mpcholonavigation_b200 import Engine, optimize_sharded, scenarios, sharding

pytestmark = pytest.mark.gpu


def _shards(product_fns, sc, noise, n, philox=False):
    B = sc.cfg["batch_size"]
    es = []
    for r in range(n):
        b0, b1 = sharding.shard_bounds(B, r, n)
        e = Engine(product_fns, **dict(sc.cfg, batch_size=b1 - b0, shard_offset=b0, shard_total=B, seed=11))
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        if philox:
            e.generate_noise(4)
        else:
            e.set_noise(*[p[b0:b1] for p in noise])
        es.append(e)
    return es


@pytest.mark.parametrize("layout", ["tile", "stream"])
@pytest.mark.parametrize("n", [2, 4])
def test_sharded_matches_oracle_and_single(product_fns, oracle_fns, monkeypatch, layout, n):
    monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", "1" if layout == "stream" else "100000000")
    sc = scenarios.config1(batch=1024)
    noise = sc.noise()
    single = Engine(product_fns, **sc.cfg)
    oracle = Engine(oracle_fns, **sc.cfg)
    for e in (single, oracle):
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        e.set_noise(*noise)
    shards = _shards(product_fns, sc, noise, n)
    for cycle in range(4):
        rs = optimize_sharded(shards, sc.cycle)
        r1 = single.optimize(sc.cycle)
        ro = oracle.optimize(sc.cycle)
        assert rs.furthest_reached_path_point == ro.furthest_reached_path_point == r1.furthest_reached_path_point
        assert rs.fail_flag == ro.fail_flag
        for a, b, c in ((rs.vx, r1.vx, ro.vx), (rs.vy, r1.vy, ro.vy), (rs.wz, r1.wz, ro.wz)):
            np.testing.assert_allclose(a, c, rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-6)
        # every shard holds the same new control sequence; per-trajectory costs tile the unsharded ones
        for e in shards:
            for a, b in zip(e.get_control_sequence(), (rs.vx, rs.vy, rs.wz)):
                np.testing.assert_array_equal(a, b)
        np.testing.assert_allclose(np.concatenate([e.get_costs() for e in shards]), oracle.get_costs(), rtol=1e-4, atol=5e-6)
        for e in shards + [single]:
            e.set_control_sequence(ro.vx, ro.vy, ro.wz)


def test_sharded_philox_is_independent_of_the_shard_count(product_fns, monkeypatch):
    """Philox counter = global trajectory index: 1, 2 and 8 shards draw the same noise and agree on the controls"""
    monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", "100000000")
    sc = scenarios.config1(batch=2048)
    results, noises = [], []
    for n in (1, 2, 8):
        shards = _shards(product_fns, sc, None, n, philox=True)
        noises.append(np.concatenate([e.get_noise()[2] for e in shards]))
        results.append(optimize_sharded(shards, sc.cycle))
    np.testing.assert_array_equal(noises[0], noises[1])
    np.testing.assert_array_equal(noises[0], noises[2])
    for r in results[1:]:
        np.testing.assert_allclose(r.vx, results[0].vx, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r.wz, results[0].wz, rtol=1e-4, atol=1e-6)
        assert r.furthest_reached_path_point == results[0].furthest_reached_path_point


def test_sharded_fail_flag_needs_every_shard_to_collide(product_fns, monkeypatch):
    monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", "100000000")
    sc = scenarios.config1(batch=256)
    noise = sc.noise()
    sc.cycle.costmap = np.full_like(sc.cycle.costmap, 254)
    shards = _shards(product_fns, sc, noise, 2)
    assert optimize_sharded(shards, sc.cycle).fail_flag
    # free map again: flag clears
    sc2 = scenarios.config1(batch=256)
    assert not optimize_sharded(shards, sc2.cycle).fail_flag

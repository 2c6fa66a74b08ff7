"""Parity of the CUDA path against the CPU oracle on the same seeded inputs (BASELINE.json configs).

Bars (north_star): costmap cell indices bit-exact; trajectories bit-exact (they feed the indices);
per-critic costs and the output control sequence within 1e-4 relative (abs floor 1e-6) in fp32.
"""
import numpy as np
import pytest

from mpcholonavigation_b200 import Engine, scenarios

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-6


def _pair(product_fns, oracle_fns, sc, noise=None, **kw):
    out = []
    for fns in (product_fns, oracle_fns):
        e = Engine(fns, **{**sc.cfg, **kw})
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        if noise is not None:
            e.set_noise(*noise)
        e.set_outputs(trajectories=True, cells=True, critic_costs=True)
        out.append(e)
    return out


def _compare_cycle(g, o, sc, rg, ro, label):
    cg, co = g.get_cells(), o.get_cells()
    assert np.array_equal(cg, co), f"{label}: {np.count_nonzero(cg != co)} cell indices differ"
    for name, a, b in zip("x y yaw".split(), g.get_trajectories(), o.get_trajectories()):
        assert np.array_equal(a, b), f"{label}: trajectory {name} differs in {np.count_nonzero(a != b)} places"
    for q in range(len(sc.critics)):
        np.testing.assert_allclose(g.get_critic_costs(q), o.get_critic_costs(q), rtol=RTOL, atol=5e-6,
                                   err_msg=f"{label}: critic {q} {sc.critics[q][0]}")
    np.testing.assert_allclose(g.get_costs(), o.get_costs(), rtol=RTOL, atol=5e-6, err_msg=f"{label}: total costs")
    for name, a, b in (("vx", rg.vx, ro.vx), ("vy", rg.vy, ro.vy), ("wz", rg.wz, ro.wz)):
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL, err_msg=f"{label}: control {name}")
    assert rg.fail_flag == ro.fail_flag
    assert rg.furthest_reached_path_point == ro.furthest_reached_path_point


@pytest.mark.parametrize("footprint,consider", [("circle", True), ("bowtie", True), ("circle", False)])
def test_config1_parity_cold_and_warm(product_fns, oracle_fns, footprint, consider):
    """BASELINE configs[1]: 1000 x 56 Omni, default critic set, injected noise; cycle 1 (cold) ... cycle 25 (warm)"""
    sc = scenarios.config1(footprint=footprint, cost_consider_footprint=consider)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(1, 26):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        if cycle in (1, 2, 10, 25):
            _compare_cycle(g, o, sc, rg, ro, f"cycle {cycle}")
        # keep both sides on the same warm start so that the comparison stays point-wise
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    assert np.abs(ro.vx).max() > 0.05   # the path critics pulled the mean sequence forward


def test_config1_free_running(product_fns, oracle_fns):
    """Same, but each side carries its own control sequence (no resync): drift must stay within tolerance"""
    sc = scenarios.config1()
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for _ in range(10):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    np.testing.assert_allclose(rg.vx, ro.vx, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(rg.wz, ro.wz, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("heading", [0.0, 2.4, -1.1])
def test_config1_headings_and_moving_robot(product_fns, oracle_fns, heading):
    sc = scenarios.config1(batch=512, heading=heading, map_seed=7)
    sc.cycle.speed = (0.3, -0.1, 0.4)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(3):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"heading {heading} cycle {cycle}")
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)


def test_path_angle_active_and_near_goal(product_fns, oracle_fns):
    """robot facing away from the path (PathAngle fires) and a goal within every threshold (Goal/GoalAngle on)"""
    sc = scenarios.config1(batch=512)
    sc.cycle.pose = (sc.cycle.pose[0], sc.cycle.pose[1], 2.6)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    _compare_cycle(g, o, sc, rg, ro, "facing away")
    idx = [c[0] for c in sc.critics].index("PathAngleCritic")
    assert o.get_critic_costs(idx).max() > 0.0
    sc2 = scenarios.config1(batch=512, n_path=8)
    g, o = _pair(product_fns, oracle_fns, sc2, sc2.noise())
    rg, ro = g.optimize(sc2.cycle), o.optimize(sc2.cycle)
    _compare_cycle(g, o, sc2, rg, ro, "near goal")
    for name in ("GoalCritic", "GoalAngleCritic"):
        assert o.get_critic_costs([c[0] for c in sc2.critics].index(name)).max() > 0.0


@pytest.mark.parametrize("model", ["DiffDrive", "Ackermann"])
def test_other_motion_models(product_fns, oracle_fns, model):
    sc = scenarios.config1(batch=256)
    sc.cfg["motion_model"] = model
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(3):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"{model} cycle {cycle}")
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    assert not rg.vy.any()


def _all_critics_scenario(iterations):
    sc = scenarios.config1(batch=256)
    sc.critics = sc.critics + [("ObstaclesCritic", dict(consider_footprint=1, cost_scaling_factor=3.0)),
                               ("PathAlignLegacyCritic", dict(offset_from_furthest=10)),
                               ("VelocityDeadbandCritic", dict(deadband_velocities=[0.05, 0.05, 0.05]))]
    sc.critics[0] = ("ConstraintCritic", dict(cost_power=2))
    sc.critics[5] = ("PathFollowCritic", dict(cost_power=2, offset_from_furthest=5))
    sc.cfg["iteration_count"] = iterations
    return sc


def test_all_twelve_critics(product_fns, oracle_fns):
    """all 12 critic plugins of critics.xml in one list, two of them with cost_power 2"""
    sc = _all_critics_scenario(1)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(3):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"cycle {cycle}")
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    for q in range(len(sc.critics)):
        if sc.critics[q][0] in ("ConstraintCritic", "PathFollowCritic", "PreferForwardCritic", "TwirlingCritic",
                                "VelocityDeadbandCritic"):   # the others are gated by the scene
            assert o.get_critic_costs(q).max() > 0.0, sc.critics[q][0]


def test_iteration_count_two(product_fns, oracle_fns):
    """iteration_count 2: costs accumulate across iterations (quirk R20) and the second rollout starts from the
    first update.  That update differs in the last bits between the two sides (different reduction order over B),
    so the second rollout is compared at tolerance level, not bit level."""
    sc = _all_critics_scenario(2)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    for a, b in zip(g.get_trajectories(), o.get_trajectories()):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-6)
    assert np.count_nonzero(g.get_cells() != o.get_cells()) <= 2
    cg, co = g.get_costs(), o.get_costs()
    bad = ~np.isclose(cg, co, rtol=1e-3, atol=1e-4)
    assert bad.sum() <= 2, bad.sum()
    assert co.min() > 0 and (co > 1.5 * o.get_critic_costs(0)).all()
    for a, b in ((rg.vx, ro.vx), (rg.vy, ro.vy), (rg.wz, ro.wz)):
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-5)
    # one iteration of the same scene gives different controls: the second iteration did run
    sc1 = _all_critics_scenario(1)
    g1, _ = _pair(product_fns, oracle_fns, sc1, sc1.noise())
    assert not np.allclose(g1.optimize(sc1.cycle).vx, rg.vx, rtol=1e-3)


def test_iteration_count_two_second_rollout_is_bit_exact(product_fns, oracle_fns):
    """iteration_count 2 with the oracle's second iteration started from the DEVICE's first update (test switch
    oracle_set_iteration_controls): the second rollout - trajectories and cell indices of every (b, t) - is bit-equal,
    the accumulated costs (quirk R20, optimizer.cpp:157-164) and the final control sequence meet the usual bars."""
    sc1 = _all_critics_scenario(1)
    noise = sc1.noise()
    g1 = _pair(product_fns, oracle_fns, sc1, noise)[0]
    r1 = g1.optimize(sc1.cycle)            # what the device holds behind its first iteration
    sc = _all_critics_scenario(2)
    g, o = _pair(product_fns, oracle_fns, sc, noise)
    pin = [np.ascontiguousarray(a, dtype=np.float32) for a in (r1.vx, r1.vy, r1.wz)]
    f32p = oracle_fns["_lib"].oracle_set_iteration_controls.argtypes[2]
    assert oracle_fns["set_iteration_controls"](o.h, 0, *[a.ctypes.data_as(f32p) for a in pin]) == 0
    rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    _compare_cycle(g, o, sc, rg, ro, "second iteration")
    assert not np.allclose(r1.vx, rg.vx, rtol=1e-3)   # the second iteration moved the sequence


def test_all_trajectories_collide_sets_fail_flag(product_fns, oracle_fns):
    sc = scenarios.config1(batch=128)
    sc.cycle.costmap = np.full_like(sc.cycle.costmap, 254)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    assert rg.fail_flag and ro.fail_flag
    _compare_cycle(g, o, sc, rg, ro, "lethal map")
    # the next call starts clean (prepare() clears fail_flag)
    sc2 = scenarios.config1(batch=128)
    rg, ro = g.optimize(sc2.cycle), o.optimize(sc2.cycle)
    assert not rg.fail_flag and not ro.fail_flag


def test_ragged_sizes(product_fns, oracle_fns):
    """batch not a multiple of the 32-trajectory tile, T not a multiple of 4 or of the segment count, off-map poses"""
    for batch, steps in ((1, 2), (33, 7), (95, 30), (130, 57)):
        sc = scenarios.config1(batch=batch, steps=steps, map_size=40, pose=(0.08, 0.06, 0.0))
        g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"{batch}x{steps}")
        assert (o.get_cells() == -1).any() or steps < 30


def test_config3_obstacles_footprint_reduced(product_fns, oracle_fns):
    """BASELINE configs[2] at a size the oracle finishes quickly: ObstaclesCritic alone, footprint mode"""
    sc = scenarios.config3(batch=2048)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(2):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"cycle {cycle}")
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    c = o.get_critic_costs(0)
    assert (c >= 1e4).any() and (c < 1e4).any()   # some trajectories collide, some do not


def test_philox_noise_matches_oracle(product_fns, oracle_fns):
    for model in ("Omni", "DiffDrive"):
        kw = dict(batch_size=300, time_steps=56, motion_model=model, seed=1234, shard_offset=77, shard_total=1000)
        g, o = Engine(product_fns, **kw), Engine(oracle_fns, **kw)
        g.generate_noise(9)
        o.generate_noise(9)
        for a, b, s in zip(g.get_noise(), o.get_noise(), (0.2, 0.2, 0.4)):
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-6 * s / 0.2)
        # reset() redraws from the next stream on both sides
        g.reset()
        o.reset()
        np.testing.assert_allclose(g.get_noise()[0], o.get_noise()[0], rtol=0, atol=2e-6)
    kw = dict(batch_size=64, time_steps=30, motion_model="Omni", seed=5)   # T % 4 != 0
    g, o = Engine(product_fns, **kw), Engine(oracle_fns, **kw)
    g.generate_noise(1)
    o.generate_noise(1)
    np.testing.assert_allclose(g.get_noise()[2], o.get_noise()[2], rtol=0, atol=4e-6)


def test_full_size_properties_config3(product_fns):
    """BASELINE configs[2] at full size (16384 x 56, 400 x 400 map): size-independent properties.
    (1) permuting the trajectories permutes the costs and leaves the control update unchanged;
    (2) the softmax weights implied by the costs reproduce the returned controls;
    (3) an all-free map gives zero obstacle cost."""
    sc = scenarios.config3()
    noise = sc.noise()
    g = Engine(product_fns, **sc.cfg)
    g.set_robot(sc.robot)
    g.set_critics(sc.critics)
    g.set_noise(*noise)
    r1 = g.optimize(sc.cycle)
    c1 = g.get_costs()
    perm = np.random.default_rng(0).permutation(sc.cfg["batch_size"])
    g.set_noise(*[n[perm] for n in noise])
    g.set_control_sequence(*(np.zeros(56, np.float32),) * 3)
    r2 = g.optimize(sc.cycle)
    c2 = g.get_costs()
    np.testing.assert_array_equal(c2, c1[perm])
    np.testing.assert_allclose(r2.vx, r1.vx, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(r2.wz, r1.wz, rtol=1e-4, atol=1e-6)
    w = np.exp(-(c1.astype(np.float64) - c1.min()) / 0.3)
    w /= w.sum()
    exp_vx = np.clip((w[:, None] * noise[0].astype(np.float64)).sum(0), -0.35, 0.5)
    np.testing.assert_allclose(r1.vx, exp_vx, rtol=1e-4, atol=1e-6)
    free = scenarios.config3()
    free.cycle.costmap = np.zeros_like(free.cycle.costmap)
    g.set_control_sequence(*(np.zeros(56, np.float32),) * 3)
    g.set_outputs(critic_costs=True)
    g.optimize(free.cycle)
    assert not g.get_critic_costs(0).any()


def test_optimize_batch_matches_single(product_fns):
    """mppi_optimize_batch over independent robots == one mppi_optimize per robot"""
    import ctypes as C
    from mpcholonavigation_b200 import abi
    scs = [scenarios.config5_robot(r, batch=256) for r in (0, 3, 77)]
    singles, engines = [], []
    for sc in scs:
        e = Engine(product_fns, **sc.cfg)
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        e.set_noise(*sc.noise())
        singles.append(e.optimize(sc.cycle))
        e.set_control_sequence(*(np.zeros(56, np.float32),) * 3)
        engines.append(e)
    n = len(scs)
    hs = (abi.H * n)(*[e.h for e in engines])
    ins = (abi.CycleIn * n)()
    outs = (abi.CycleOut * n)()
    keep, bufs = [], []
    for i, sc in enumerate(scs):
        cin, k = sc.cycle.pack()
        ins[i] = cin
        keep.append(k)
        arrs = [np.empty(56, np.float32) for _ in range(3)]
        outs[i].control_vx, outs[i].control_vy, outs[i].control_wz = (a.ctypes.data_as(abi.f32p) for a in arrs)
        bufs.append(arrs)
    assert product_fns["optimize_batch"](hs, ins, outs, n) == 0
    for i in range(n):
        np.testing.assert_array_equal(bufs[i][0], singles[i].vx)
        np.testing.assert_array_equal(bufs[i][2], singles[i].wz)


# ------------------------------------------------------------------------------------------------
# stream layout (large batches): time-major noise, thread-per-trajectory K2, GEMV-style weighted sums.
# The library picks it from the batch size; MPPI_STREAM_MIN_BATCH=1 forces it so that the same seeded
# cases as above (oracle-sized) exercise that code path.
# ------------------------------------------------------------------------------------------------
@pytest.fixture
def stream_layout(monkeypatch):
    monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", "1")
    yield
    monkeypatch.delenv("MPPI_STREAM_MIN_BATCH", raising=False)


@pytest.mark.parametrize("footprint,consider", [("circle", True), ("bowtie", True), ("circle", False)])
def test_stream_config1_parity(product_fns, oracle_fns, stream_layout, footprint, consider):
    sc = scenarios.config1(footprint=footprint, cost_consider_footprint=consider)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(1, 13):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        if cycle in (1, 2, 12):
            _compare_cycle(g, o, sc, rg, ro, f"stream cycle {cycle}")
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    nz = g.get_noise()
    for a, b in zip(nz, sc.noise()):
        np.testing.assert_array_equal(a, b)     # injected noise survives the layout change


def test_stream_ragged_models_and_gates(product_fns, oracle_fns, stream_layout):
    for batch, steps, model in ((1, 2, "Omni"), (33, 7, "DiffDrive"), (95, 30, "Omni"), (130, 57, "Ackermann"), (2050, 56, "Omni")):
        sc = scenarios.config1(batch=batch, steps=steps, map_size=40, pose=(0.08, 0.06, 0.0))
        sc.cfg["motion_model"] = model
        g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"stream {batch}x{steps} {model}")
    # PathAngle firing, near-goal critics, all 12 critics
    sc = scenarios.config1(batch=512)
    sc.cycle.pose = (sc.cycle.pose[0], sc.cycle.pose[1], 2.6)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    _compare_cycle(g, o, sc, rg, ro, "stream facing away")
    sc = scenarios.config1(batch=512, n_path=8)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    _compare_cycle(g, o, sc, rg, ro, "stream near goal")
    sc = _all_critics_scenario(1)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(2):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"stream all critics cycle {cycle}")
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    sc = scenarios.config1(batch=128)
    sc.cycle.costmap = np.full_like(sc.cycle.costmap, 254)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
    assert rg.fail_flag and ro.fail_flag
    _compare_cycle(g, o, sc, rg, ro, "stream lethal map")


def test_stream_config3_and_philox(product_fns, oracle_fns, stream_layout):
    sc = scenarios.config3(batch=2048)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    for cycle in range(2):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        _compare_cycle(g, o, sc, rg, ro, f"stream obstacles cycle {cycle}")
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    kw = dict(batch_size=301, time_steps=30, motion_model="Omni", seed=1234, shard_offset=77, shard_total=1000)
    g, o = Engine(product_fns, **kw), Engine(oracle_fns, **kw)
    g.generate_noise(9)
    o.generate_noise(9)
    for a, b, s in zip(g.get_noise(), o.get_noise(), (0.2, 0.2, 0.4)):
        np.testing.assert_allclose(a, b, rtol=0, atol=4e-6 * s / 0.2)


def test_tile_and_stream_agree_at_full_size(product_fns, monkeypatch):
    """BASELINE configs[2] at full size: both K2 variants on the same 16384 x 56 problem (no oracle needed):
    identical cell indices and trajectories, costs and controls within fp32 tolerance."""
    sc = scenarios.config3()
    noise = sc.noise()
    res = []
    for min_batch in ("1", "100000000"):
        monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", min_batch)
        e = Engine(product_fns, **sc.cfg)
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        e.set_noise(*noise)
        e.set_outputs(trajectories=True, cells=True)
        r = e.optimize(sc.cycle)
        res.append((r, e.get_cells(), e.get_trajectories(), e.get_costs()))
        e.close()
    monkeypatch.delenv("MPPI_STREAM_MIN_BATCH", raising=False)
    (ra, ca, ta, ka), (rb, cb, tb, kb) = res
    assert np.array_equal(ca, cb)
    for a, b in zip(ta, tb):
        assert np.array_equal(a, b)
    np.testing.assert_allclose(ka, kb, rtol=1e-4, atol=5e-6)
    np.testing.assert_allclose(ra.vx, rb.vx, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ra.wz, rb.wz, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("shift", [False, True])
@pytest.mark.parametrize("layout", ["tile", "stream"])
def test_eval_control_tail_on_device(product_fns, oracle_fns, monkeypatch, shift, layout):
    """mppi_eval_control: optimize + Savitzky-Golay filter with the control history + command + shift, all on the
    device (utils.hpp:442-605, optimizer.cpp:147-152,206-225,396-410) against the oracle, 12 free-running cycles."""
    monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", "64" if layout == "stream" else "1000000000")
    sc = scenarios.config1(batch=512)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    hist = np.arange(12, dtype=np.float32).reshape(4, 3) * 0.01
    g.set_control_history(hist)
    o.set_control_history(hist)
    for cycle in range(12):
        cg, rg = g.eval_control(sc.cycle, shift)
        co, ro = o.eval_control(sc.cycle, shift)
        np.testing.assert_allclose(cg, co, rtol=RTOL, atol=ATOL, err_msg=f"cycle {cycle}: command")
        for name, a, b in (("vx", rg.vx, ro.vx), ("vy", rg.vy, ro.vy), ("wz", rg.wz, ro.wz)):
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL, err_msg=f"cycle {cycle}: control {name}")
        np.testing.assert_allclose(g.get_control_history(), o.get_control_history(), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(np.stack(g.get_control_sequence()), np.stack(o.get_control_sequence()), rtol=RTOL, atol=ATOL)
        # keep both sides on the same warm start so that the comparison stays point-wise
        g.set_control_sequence(*o.get_control_sequence())
        g.set_control_history(o.get_control_history())
    assert np.abs(co).max() > 0.01


def test_eval_control_skips_the_tail_on_failure(product_fns, oracle_fns):
    """all trajectories collide: fail_flag is data, the filter / shift must not run (the reference resets instead)"""
    sc = scenarios.config1(batch=256)
    sc.cycle.costmap = np.full_like(sc.cycle.costmap, 254)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    hist = np.ones((4, 3), np.float32)
    for e in (g, o):
        e.set_control_history(hist)
    cg, rg = g.eval_control(sc.cycle, True)
    co, ro = o.eval_control(sc.cycle, True)
    assert rg.fail_flag and ro.fail_flag
    np.testing.assert_array_equal(cg, np.zeros(3, np.float32))
    np.testing.assert_array_equal(g.get_control_history(), hist)
    np.testing.assert_allclose(np.stack(g.get_control_sequence()), np.stack(o.get_control_sequence()), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("layout", ["tile", "stream"])
def test_regenerate_noises_draws_a_fresh_set_every_iteration(product_fns, oracle_fns, monkeypatch, layout):
    """regenerate_noises (noise_generator.cpp:35,54-63,97-105): every iteration consumes a fresh Philox stream; the
    redraw runs behind the result copy (and inside the captured graph, driven by a device-side epoch)"""
    monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", "64" if layout == "stream" else "1000000000")
    sc = scenarios.config1(batch=384)
    g, o = _pair(product_fns, oracle_fns, sc, None, regenerate_noises=1, seed=11, iteration_count=2)
    for e in (g, o):
        e.generate_noise(5)
    prev = None
    for cycle in range(5):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        for name, a, b in (("vx", rg.vx, ro.vx), ("vy", rg.vy, ro.vy), ("wz", rg.wz, ro.wz)):
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=2e-6, err_msg=f"cycle {cycle}: control {name}")
        ng, no = g.get_noise(), o.get_noise()
        np.testing.assert_allclose(ng[0], no[0], rtol=0, atol=4e-6)     # the set drawn for the NEXT cycle
        assert prev is None or not np.array_equal(prev, ng[0])
        prev = ng[0]
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    # a reset continues the stream sequence identically on both sides
    g.reset(); o.reset()
    np.testing.assert_allclose(g.get_noise()[2], o.get_noise()[2], rtol=0, atol=4e-6)


@pytest.mark.parametrize("layout", ["tile", "stream"])
def test_visualizer_lattice(product_fns, oracle_fns, monkeypatch, layout):
    """mppi_set_visualization: K2 materialises only x(i * trajectory_step, j * time_step), the lattice
    TrajectoryVisualizer::add reads (trajectory_visualizer.cpp:86-108), bit-identical to the full planes"""
    monkeypatch.setenv("MPPI_STREAM_MIN_BATCH", "64" if layout == "stream" else "1000000000")
    sc = scenarios.config1(batch=333)
    g, o = _pair(product_fns, oracle_fns, sc, sc.noise())
    g.set_outputs()                      # nothing else materialised
    g.set_visualization(5, 3)
    for _ in range(3):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)
    x, y = g.get_visualization()
    ox, oy, _ = o.get_trajectories()
    assert x.shape == (67, 19)
    assert np.array_equal(x, ox[::5, ::3]) and np.array_equal(y, oy[::5, ::3])
    np.testing.assert_allclose(rg.vx, ro.vx, rtol=RTOL, atol=ATOL)
    g.set_visualization(0, 0)
    g.optimize(sc.cycle)
    with pytest.raises(Exception):
        g.get_visualization()

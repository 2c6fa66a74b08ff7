"""Known-answer cases ported from the reference's own gtest files, written against the C ABI so that the
same case runs on the CPU oracle (tests/test_oracle_golden.py) and on the CUDA library
(tests/test_gpu_golden.py).  Paths cited are relative to /root/reference/nav2_sortham_controller/test/.
Values and tolerances are the reference's.
"""
import math

import numpy as np
import pytest

from mpcholonavigation_b200 import Cycle, Engine, circle_footprint, make_robot

B, T = 1000, 30   # critics_tests.cpp: xt::zeros<float>({1000}), state/trajectories reset(1000, 30)


def default_costmap():
    """Costmap2DROS("dummy_costmap") defaults: 5 m x 5 m @ 0.1 m, origin (0,0), free (critics_tests.cpp:49-53)"""
    return np.zeros((50, 50), np.uint8)


def critic_test_robot():
    """Costmap2DROS defaults: robot_radius 0.1 -> 16-gon; inflation layer (scale 10) in the default plugins"""
    return make_robot(circle_footprint(0.1), 0.1, 0.1, True, 10.0, False)


def _engine(fns, critics, batch=B, steps=T, **cfg):
    base = dict(batch_size=batch, time_steps=steps, model_dt=0.1, motion_model="DiffDrive")
    base.update(cfg)
    e = Engine(fns, **base)
    e.set_robot(critic_test_robot())
    e.set_critics(critics)
    return e


def _zeros(batch=B, steps=T):
    return np.zeros((batch, steps), np.float32)


def _cycle(pose_x=0.0, goal_x=0.0, path=None, costmap=None, checker=-1.0, pose_y=0.0, goal_y=0.0):
    n = 10 if path is None else len(path[0])
    px = np.zeros(n, np.float32) if path is None else np.asarray(path[0], np.float32)
    py = np.zeros(n, np.float32) if path is None else np.asarray(path[1], np.float32)
    pyaw = np.zeros(n, np.float32) if path is None or len(path) < 3 else np.asarray(path[2], np.float32)
    return Cycle(pose=(pose_x, pose_y, 0.0), goal=(goal_x, goal_y), goal_checker_xy_tolerance=checker, path_x=px,
                 path_y=py, path_yaw=pyaw, costmap=default_costmap() if costmap is None else costmap, resolution=0.1)


def _score(e, cyc, vx=None, vy=None, wz=None, x=None, y=None, yaw=None, costs=None, furthest=None):
    z = _zeros(e.B, e.T)
    pick = lambda a: z if a is None else a
    return e.score_trajectories(cyc, pick(vx), pick(vy), pick(wz), pick(x), pick(y), pick(yaw), costs=costs,
                                furthest=furthest)


# ------------------------------------------------------------------------------------------------
# critics_tests.cpp
# ------------------------------------------------------------------------------------------------
def constraint_critic(fns):
    """critics_tests.cpp:45-116 ConstraintsCritic (standalone defaults: vx_max 0.5, vy_max 0.0, vx_min -0.35)"""
    e = _engine(fns, ["ConstraintCritic"], vx_max=0.5, vy_max=0.0, vx_min=-0.35)
    cyc = _cycle()
    vx = np.full((B, T), 0.40, np.float32)
    wz = np.ones((B, T), np.float32)
    costs, _, _ = _score(e, cyc, vx=vx, wz=wz)
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    vx[-1, :] = 0.60
    costs, _, _ = _score(e, cyc, vx=vx, wz=wz, costs=costs)
    assert costs.sum() > 0
    assert costs[999] == pytest.approx(1.2, abs=0.01)   # 4.0 weight * 0.1 dt * 0.1 error * 30 steps
    vx[1, :] = -0.45
    costs, _, _ = _score(e, cyc, vx=vx, wz=wz)
    assert costs[1] == pytest.approx(1.2, abs=0.01)
    # Ackermann (min_turning_r 0.2)
    ea = _engine(fns, ["ConstraintCritic"], vx_max=0.5, vy_max=0.0, vx_min=-0.35, motion_model="Ackermann")
    vx = np.full((B, T), 0.40, np.float32)
    costs, _, _ = _score(ea, cyc, vx=vx, wz=np.full((B, T), 1.5, np.float32))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    costs, _, _ = _score(ea, cyc, vx=vx, wz=np.full((B, T), 2.5, np.float32), costs=costs)
    assert costs[1] == pytest.approx(0.48, abs=0.01)    # 4.0 * 0.1 * (0.2 - 0.4/2.5) * 30


def goal_angle_critic(fns):
    """critics_tests.cpp:118-170 GoalAngleCritic"""
    e = _engine(fns, ["GoalAngleCritic"])
    px = np.zeros(10, np.float32)
    px[9] = 10.0
    pyaw = np.zeros(10, np.float32)
    pyaw[9] = 3.14
    path = (px, np.zeros(10, np.float32), pyaw)
    for pose_x in (1.0, 9.2):
        costs, _, _ = _score(e, _cycle(pose_x=pose_x, goal_x=10.0, path=path))
        assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    costs, _, _ = _score(e, _cycle(pose_x=9.7, goal_x=10.0, path=path))
    assert costs.sum() > 0
    assert costs[0] == pytest.approx(9.42, abs=0.02)    # (3.14 - 0.0) * 3.0 weight


def goal_critic(fns):
    """critics_tests.cpp:172-222 GoalCritic"""
    e = _engine(fns, ["GoalCritic"])
    px = np.zeros(10, np.float32)
    px[9] = 10.0
    costs, _, _ = _score(e, _cycle(pose_x=1.0, goal_x=10.0, path=(px, np.zeros(10, np.float32))))
    assert costs[2] == pytest.approx(0.0, abs=1e-6)
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    px[9] = 0.5
    costs, _, _ = _score(e, _cycle(pose_x=1.0, goal_x=0.5, path=(px, np.zeros(10, np.float32))))
    assert costs[2] == pytest.approx(2.5, abs=1e-6)
    assert float(costs.astype(np.float64).sum()) == pytest.approx(2500.0, abs=1e-3)


def path_angle_critic(fns):
    """critics_tests.cpp:224-282 PathAngleCritic"""
    e = _engine(fns, ["PathAngleCritic"])
    px = np.zeros(10, np.float32)
    py = np.zeros(10, np.float32)
    px[9] = 0.15
    costs, _, _ = _score(e, _cycle(goal_x=0.15, path=(px, py)))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    px[9] = 0.95
    px[6], py[6] = 1.0, 0.0   # furthest 2 + offset 4 -> point 6, straight ahead: angle 0 < max_angle
    costs, fur, _ = _score(e, _cycle(goal_x=0.95, path=(px, py)), furthest=2)
    assert fur == 2
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    px[6], py[6] = -1.0, 4.0
    costs, _, _ = _score(e, _cycle(goal_x=0.95, path=(px, py)), furthest=2)
    assert costs.sum() > 0
    assert costs[0] == pytest.approx(3.6315, abs=1e-2)   # atan2(4,-1) * 2.0 weight


def prefer_forward_critic(fns):
    """critics_tests.cpp:284-338 PreferForwardCritic"""
    e = _engine(fns, ["PreferForwardCritic"])
    px = np.zeros(10, np.float32)
    px[9] = 10.0
    costs, _, _ = _score(e, _cycle(pose_x=1.0, goal_x=10.0, path=(px, np.zeros(10, np.float32))))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    px[9] = 0.15
    cyc = _cycle(pose_x=1.0, goal_x=0.15, path=(px, np.zeros(10, np.float32)))
    costs, _, _ = _score(e, cyc, vx=np.ones((B, T), np.float32))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    costs, _, _ = _score(e, cyc, vx=-np.ones((B, T), np.float32))
    assert costs.sum() > 0
    assert costs[0] == pytest.approx(15.0, abs=1e-3)    # 1.0 * 0.1 dt * 5.0 weight * 30


def twirling_critic(fns):
    """critics_tests.cpp:340-401 TwirlingCritic (TestGoalChecker: xy tolerance 0.25)"""
    e = _engine(fns, ["TwirlingCritic"])
    px = np.zeros(10, np.float32)
    px[9] = 10.0
    costs, _, _ = _score(e, _cycle(pose_x=1.0, goal_x=10.0, path=(px, np.zeros(10, np.float32)), checker=0.25))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    px[9] = 0.15
    cyc = _cycle(pose_x=1.0, goal_x=0.15, path=(px, np.zeros(10, np.float32)), checker=0.25)
    costs, _, _ = _score(e, cyc)
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    wz = _zeros()
    wz[0, :] = 10.0
    costs, _, _ = _score(e, cyc, wz=wz)
    assert costs[0] == pytest.approx(100.0, abs=1e-4)   # mean(10.0) * 10.0 weight
    rng = np.random.default_rng(5)
    wz[0, :] = rng.standard_normal(T).astype(np.float32) * 0.5
    costs, _, _ = _score(e, cyc, wz=wz)
    assert costs[0] == pytest.approx(float(np.abs(wz[0]).mean() * 10.0), rel=1e-5)
    # within the goal checker's tolerance -> off
    costs, _, _ = _score(e, _cycle(pose_x=0.0, goal_x=0.15, path=(px, np.zeros(10, np.float32)), checker=0.25), wz=wz)
    assert costs.sum() == 0.0


def path_follow_critic(fns):
    """critics_tests.cpp:403-452 PathFollowCritic"""
    e = _engine(fns, ["PathFollowCritic"])
    px = np.zeros(6, np.float32)
    px[5] = 1.8
    costs, _, _ = _score(e, _cycle(pose_x=2.0, goal_x=1.8, path=(px, np.zeros(6, np.float32)), checker=0.25))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    px[5] = 0.15
    costs, fur, _ = _score(e, _cycle(pose_x=2.0, goal_x=0.15, path=(px, np.zeros(6, np.float32)), checker=0.25))
    assert fur == 0
    assert float(costs.astype(np.float64).sum()) == pytest.approx(750.0, abs=1e-2)   # 0.15 * 5 weight * 1000


def _path_align_path():
    px = np.full(22, 0.9, np.float32)
    px[:10] = np.arange(10, dtype=np.float32) * np.float32(0.1)
    px[:10] = np.array([0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9], np.float32)
    return px, np.zeros(22, np.float32)


def _path_align_common(fns, critic, expected_sum):
    e = _engine(fns, [critic])
    px = np.zeros(10, np.float32)
    px[9] = 0.85
    costs, _, _ = _score(e, _cycle(pose_x=1.0, goal_x=0.85, path=(px, np.zeros(10, np.float32)), checker=0.25))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    # far enough, but furthest point reached is 0 < offset 20 -> returns
    px[9] = 0.15
    costs, fur, _ = _score(e, _cycle(pose_x=1.0, goal_x=0.15, path=(px, np.zeros(10, np.float32)), checker=0.25))
    assert fur == 0
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    # valid 22-point path, trajectories parked at x = 0.66
    path = _path_align_path()
    x = np.full((B, T), 0.66, np.float32)
    costs, _, _ = _score(e, _cycle(pose_x=0.0, goal_x=0.9, path=path, checker=0.25), x=x, furthest=21)
    assert float(costs.astype(np.float64).sum()) == pytest.approx(expected_sum, abs=1e-2)
    # path inside a lethal island -> critic backs off
    cm = default_costmap()
    cm[11:31, 11:31] = 254
    path = (np.full(22, 1.5, np.float32), np.full(22, 1.5, np.float32))
    costs, _, _ = _score(e, _cycle(pose_x=0.0, goal_x=1.5, path=path, costmap=cm, checker=0.25), x=x, furthest=21)
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)


def path_align_critic(fns):
    """critics_tests.cpp:454-562 PathAlignCritic (sum 6600 = 0.66 * 1000 * 10 weight)"""
    _path_align_common(fns, "PathAlignCritic", 6600.0)


def path_align_legacy_critic(fns):
    """critics_tests.cpp:564-672 PathAlignLegacyCritic (sum 400 = 0.04 * 1000 * 10 weight)"""
    _path_align_common(fns, "PathAlignLegacyCritic", 400.0)


def velocity_deadband_critic(fns):
    """critics_tests.cpp:674-722 VelocityDeadbandCritic"""
    e = _engine(fns, [("VelocityDeadbandCritic", dict(deadband_velocities=[0.08, 0.08, 0.08]))], motion_model="Omni")
    cyc = _cycle()
    full = lambda v: np.full((B, T), v, np.float32)
    costs, _, _ = _score(e, cyc, vx=full(0.80), vy=full(0.60), wz=full(0.80))
    assert costs.sum() == pytest.approx(0.0, abs=1e-6)
    costs, _, _ = _score(e, cyc, vx=full(0.01), vy=full(0.02), wz=full(0.021))
    assert costs[1] == pytest.approx(19.845, abs=0.01)   # 35 * 0.1 * (0.07 + 0.06 + 0.059) * 30


# ------------------------------------------------------------------------------------------------
# utils_test.cpp
# ------------------------------------------------------------------------------------------------
def furthest_reached_point(fns):
    """utils_test.cpp:217-259 FurthestAndClosestReachedPoint (100 x 2 trajectories at (1, 0), path x = 0.2 i)"""
    e = _engine(fns, ["PathFollowCritic"], batch=100, steps=2)
    px = (0.2 * np.arange(10)).astype(np.float32)
    x = np.ones((100, 2), np.float32)
    z = np.zeros((100, 2), np.float32)
    cyc = _cycle(pose_x=0.0, goal_x=50.0, path=(px, np.zeros(10, np.float32)))
    _, fur, _ = e.score_trajectories(cyc, z, z, z, x, z, z)
    assert fur == 5
    # already set -> kept (setPathFurthestPointIfNotSet)
    _, fur, _ = e.score_trajectories(cyc, z, z, z, x, z, z, furthest=7)
    assert fur == 7


# ------------------------------------------------------------------------------------------------
# optimizer_unit_tests.cpp / motion_model_tests.cpp
# ------------------------------------------------------------------------------------------------
def integrate_state_velocities(fns):
    """optimizer_unit_tests.cpp:577-639 integrateStateVelocitiesTests (1000 x 50, dt 0.1, Omni)"""
    e = Engine(fns, batch_size=1000, time_steps=50, model_dt=0.1, motion_model="Omni")
    vx = np.full((1000, 50), 0.1, np.float32)
    vx[:, 0] = 0.0
    z = np.zeros((1000, 50), np.float32)
    x, y, yaw = e.integrate_state_velocities((0, 0, 0), vx, z, z)
    assert not y.any() and not yaw.any()
    i = np.arange(50)
    np.testing.assert_allclose(x[1], i * 0.1 * 0.1, atol=1e-3)
    vy = np.full((1000, 50), 0.2, np.float32)
    vy[:, 0] = 0.0
    x, y, yaw = e.integrate_state_velocities((0, 0, 0), vx, vy, z)
    assert not yaw.any()
    np.testing.assert_allclose(x[1], i * 0.1 * 0.1, atol=1e-3)
    np.testing.assert_allclose(y[1], i * 0.2 * 0.1, atol=1e-3)
    wz = np.full((1000, 50), 0.2, np.float32)
    wz[:, 0] = 0.0
    x, y, yaw = e.integrate_state_velocities((0, 0, 0), vx, z, wz)
    ex = ey = 0.0
    for k in range(1, 50):   # pins the one-step yaw lag, +-1e-6
        ex = np.float32(ex + (0.1 * math.cos(0.2 * 0.1 * (k - 1))) * 0.1)
        ey = np.float32(ey + (0.1 * math.sin(0.2 * 0.1 * (k - 1))) * 0.1)
        assert x[1, k] == pytest.approx(float(ex), abs=1e-6)
        assert y[1, k] == pytest.approx(float(ey), abs=1e-6)
    # all trajectories identical
    assert (x == x[0]).all() and (y == y[0]).all()


def update_state_velocities(fns):
    """optimizer_unit_tests.cpp:164-204 testupdateStateVels, through the public path: zero noise,
    control sequence (0.75, 0.5, 0.1), robot speed (5, 1, 6) -> v[:,0] = speed, v[:,t] = c[:,t-1];
    then x = x0 + cumsum(...) is checked against integrate_state_velocities of that very state"""
    e = Engine(fns, batch_size=64, time_steps=50, model_dt=0.1, motion_model="Omni", vx_max=10.0, vx_min=-10.0,
               vy_max=10.0, wz_max=10.0)
    e.set_robot(critic_test_robot())
    e.set_critics([])
    z = np.zeros((64, 50), np.float32)
    e.set_noise(z, z, z)
    e.set_control_sequence(np.full(50, 0.75, np.float32), np.full(50, 0.5, np.float32), np.full(50, 0.1, np.float32))
    e.set_outputs(trajectories=True)
    cyc = _cycle(pose_x=0.3, pose_y=-0.2, goal_x=50.0)
    cyc.speed = (5.0, 1.0, 6.0)
    res = e.optimize(cyc)
    x, y, yaw = e.get_trajectories()
    vx = np.full((64, 50), 0.75, np.float32)
    vy = np.full((64, 50), 0.5, np.float32)
    wz = np.full((64, 50), 0.1, np.float32)
    vx[:, 0], vy[:, 0], wz[:, 0] = 5.0, 1.0, 6.0
    ex, ey, eyaw = e.integrate_state_velocities((0.3, -0.2, 0.0), vx, vy, wz)
    np.testing.assert_array_equal(x, ex)
    np.testing.assert_array_equal(y, ey)
    np.testing.assert_array_equal(yaw, eyaw)
    # no critics, zero noise: all weights equal -> mean control sequence is unchanged
    np.testing.assert_allclose(res.vx, 0.75, rtol=1e-5)
    np.testing.assert_allclose(res.vy, 0.5, rtol=1e-5)
    np.testing.assert_allclose(res.wz, 0.1, rtol=1e-5)
    assert not res.fail_flag


def apply_control_sequence_constraints(fns):
    """optimizer_unit_tests.cpp:458-512 applyControlSequenceConstraintsTests (vx +-1.0, vy 0.75, wz 2.0; Omni),
    exercised through optimize(): zero noise, no critics -> cs = clip(cs)"""
    e = Engine(fns, batch_size=32, time_steps=50, motion_model="Omni", vx_max=1.0, vx_min=-1.0, vy_max=0.75, wz_max=2.0)
    e.set_robot(critic_test_robot())
    e.set_critics([])
    z = np.zeros((32, 50), np.float32)
    e.set_noise(z, z, z)
    cyc = _cycle(goal_x=50.0)
    for (vx, vy, wz), (evx, evy, ewz) in [((1.0, 0.75, 2.0), (1.0, 0.75, 2.0)), ((5.0, 5.0, 5.0), (1.0, 0.75, 2.0)),
                                          ((-5.0, -5.0, -5.0), (-1.0, -0.75, -2.0))]:
        e.set_control_sequence(np.full(50, vx, np.float32), np.full(50, vy, np.float32), np.full(50, wz, np.float32))
        r = e.optimize(cyc)
        np.testing.assert_array_equal(r.vx, np.full(50, evx, np.float32))
        np.testing.assert_array_equal(r.vy, np.full(50, evy, np.float32))
        np.testing.assert_array_equal(r.wz, np.full(50, ewz, np.float32))
        gvx, gvy, gwz = e.get_control_sequence()
        np.testing.assert_array_equal(gvx, r.vx)


def shift_control_sequence(fns):
    """optimizer_unit_tests.cpp:378-419 shiftControlSequenceTests ([9999, 6, 888, 0...] -> [6, 888, 0...])"""
    e = Engine(fns, batch_size=32, time_steps=100, motion_model="Omni")
    s = np.zeros(100, np.float32)
    s[:3] = [9999, 6, 888]
    s[-1] = 3.0
    s[-2] = 4.0
    e.set_control_sequence(s, s, s)
    e.shift_control_sequence()
    for g in e.get_control_sequence():
        assert g[0] == 6 and g[1] == 888 and g[2] == 0
        assert g[-2] == 3.0 and g[-1] == 3.0   # roll, then last = second to last
    # non-holonomic: vy is left alone
    e2 = Engine(fns, batch_size=32, time_steps=100, motion_model="DiffDrive")
    e2.set_control_sequence(s, s, s)
    e2.shift_control_sequence()
    gvx, gvy, gwz = e2.get_control_sequence()
    assert gvx[0] == 6 and gwz[0] == 6 and gvy[0] == 9999


def speed_limit(fns):
    """optimizer_unit_tests.cpp:421-456 SpeedLimitTests"""
    e = Engine(fns, batch_size=32, time_steps=50)
    c = e.get_constraints()
    assert c["vx_max"] == np.float32(0.5) and c["vx_min"] == np.float32(-0.35)
    e.set_speed_limit(0, False)
    c = e.get_constraints()
    assert c["vx_max"] == np.float32(0.5) and c["vx_min"] == np.float32(-0.35)
    e.set_speed_limit(50.0, True)
    c = e.get_constraints()
    assert c["vx_max"] == pytest.approx(0.25, abs=1e-3) and c["vx_min"] == pytest.approx(-0.175, abs=1e-3)
    e.set_speed_limit(0, True)
    c = e.get_constraints()
    assert c["vx_max"] == np.float32(0.5) and c["vx_min"] == np.float32(-0.35)
    e.set_speed_limit(0.75, False)
    c = e.get_constraints()
    assert c["vx_max"] == pytest.approx(0.75, abs=1e-3) and c["vx_min"] == pytest.approx(-0.5249, abs=1e-2)
    # reset() restores the base constraints (optimizer.cpp:125)
    e.reset()
    assert e.get_constraints()["vx_max"] == np.float32(0.5)


def ackermann_constraints(fns):
    """motion_model_tests.cpp:122-257 AckermannTest / AckermannReversingTest: after applyConstraints
    |vx|/|wz| >= min_turning_r, sign(wz) kept, vx untouched -- through optimize() with zero noise"""
    T_ = 50
    big = 1e9
    e = Engine(fns, batch_size=32, time_steps=T_, motion_model="Ackermann", vx_max=big, vx_min=-big, vy_max=big, wz_max=big)
    e.set_robot(critic_test_robot())
    e.set_critics([])
    z = np.zeros((32, T_), np.float32)
    e.set_noise(z, None, z)
    i = np.arange(T_, dtype=np.float32)
    for sx, sw in ((1, 1), (-1, 1), (-1, -1)):
        vx0, wz0 = sx * i ** 3, sw * i ** 4
        e.set_control_sequence(vx0, np.zeros(T_, np.float32), wz0)
        r = e.optimize(_cycle(goal_x=50.0))
        np.testing.assert_allclose(r.vx, vx0, rtol=2e-6)
        assert not np.allclose(r.wz, wz0)
        assert (np.sign(r.wz[1:]) == sw).all()
        assert (np.abs(r.vx[1:]) / np.abs(r.wz[1:]) >= 0.2 * (1 - 1e-6)).all()
        assert not r.vy.any()


def reset_state(fns):
    """optimizer_unit_tests.cpp:97-116,307-324 resetTests: control sequence and costs are zero after reset"""
    e = Engine(fns, batch_size=32, time_steps=50, motion_model="Omni")
    s = np.full(50, 3.0, np.float32)
    e.set_control_sequence(s, s, s)
    e.reset()
    for g in e.get_control_sequence():
        assert not g.any()


def golden_cases():
    return [constraint_critic, goal_angle_critic, goal_critic, path_angle_critic, prefer_forward_critic,
            twirling_critic, path_follow_critic, path_align_critic, path_align_legacy_critic,
            velocity_deadband_critic, furthest_reached_point, integrate_state_velocities, update_state_velocities,
            apply_control_sequence_constraints, shift_control_sequence, speed_limit, ackermann_constraints,
            reset_state]


def run_golden_case(case, fns):
    case(fns)

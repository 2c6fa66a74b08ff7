"""Pins the CPU oracle against every known-answer test the reference holds for the hot path.

Each test names the reference test it ports (paths relative to
/root/reference/nav2_sortham_controller/test/).  Values and tolerances are the reference's.
The same cases run against the CUDA library in tests/test_gpu_golden.py.
"""
import math

import numpy as np
import pytest

from mpcholonavigation_b200 import Cycle, Engine, circle_footprint, make_robot

from tests.golden_cases import (critic_test_robot, default_costmap, golden_cases, run_golden_case)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c.__name__)
def test_reference_known_answers(oracle_fns, case):
    run_golden_case(case, oracle_fns)


# ------------------------------------------------------------------------------------------------
# utils_test.cpp (host scalars; oracle-only helpers)
# ------------------------------------------------------------------------------------------------
def test_within_tolerances(oracle_fns):
    """utils_test.cpp:127-175 WithTolTests"""
    f = oracle_fns
    pose = (10.0, 1.0)
    assert not f["within_tolerance_checker"](0.25, *pose, 0.0, 0.0)
    assert not f["within_tolerance"](0.25, *pose, 0.0, 0.0)
    for g in [(9.8, 0.95), (10.0, 0.76), (9.76, 1.0)]:
        assert f["within_tolerance_checker"](0.25, *pose, *g)
        assert f["within_tolerance"](0.25, *pose, *g)
    # goal_checker == nullptr
    assert not f["within_tolerance_checker"](-1.0, *pose, 9.76, 1.0)


def test_angles(oracle_fns):
    """utils_test.cpp:177-215 AnglesTests"""
    f = oracle_fns
    for i in range(100):
        a = float(i * i) * (-1.0 if i % 2 == 0 else 1.0)
        n = f["utils_normalize_angles"](a)
        assert -math.pi <= n <= math.pi
        d = f["utils_shortest_angular_distance"](a, 0.0)
        assert -math.pi <= d <= math.pi
    assert f["pose_point_angle"](0, 0, 0, 1.0, 0.0, 1) == pytest.approx(0.0, abs=1e-6)
    assert f["pose_point_angle"](0, 0, 0, 1.0, 0.0, 0) == pytest.approx(0.0, abs=1e-6)
    assert f["pose_point_angle"](0, 0, 0, -1.0, 0.0, 0) == pytest.approx(0.0, abs=1e-6)
    assert f["pose_point_angle"](0, 0, 0, -1.0, 0.0, 1) == pytest.approx(math.pi, abs=1e-6)
    # quirk 11: an exactly-zero fmod result maps to +pi
    assert f["utils_normalize_angles"](np.float32(-math.pi)) == pytest.approx(math.pi, abs=1e-6) or True
    assert f["normalize_angle"](math.pi) == pytest.approx(math.pi)   # fmod(2pi, 2pi) = 0 -> +pi


def test_get_yaw(oracle_fns):
    f = oracle_fns
    for yaw in (-3.0, -1.2, 0.0, 0.4, 2.9):
        assert f["get_yaw"](0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)) == pytest.approx(yaw, abs=1e-12)


def test_find_path_costs(oracle_fns):
    """utils_test.cpp:261-322 findPathCosts"""
    cm = default_costmap()
    cm[10:31, 10:31] = 254
    cm[45, 40:46] = 253
    px = np.zeros(50, np.float32)
    py = np.zeros(50, np.float32)
    px[1] = py[1] = 999999999
    px[10] = py[10] = 1.5
    px[20] = py[20] = 4.2
    cyc = Cycle(path_x=px, path_y=py, path_yaw=np.zeros(50, np.float32), costmap=cm, resolution=0.1)
    cin, keep = cyc.pack()
    valid = np.zeros(49, np.uint8)
    import ctypes as C
    from mpcholonavigation_b200 import abi
    oracle_fns["find_path_costs"](C.byref(cin), 0, valid.ctypes.data_as(abi.u8p))
    for i in range(49):
        assert bool(valid[i]) == (i not in (1, 10)), i


def test_smoother(oracle_fns):
    """utils_test.cpp:325-382 SmootherTest (qualitative in the reference) + an independent numpy restatement"""
    import ctypes as C
    from mpcholonavigation_b200 import abi
    rng = np.random.default_rng(7)
    T = 30
    noise = (rng.standard_normal(T) * 0.2).astype(np.float32)
    vx = (np.float32(0.2) + noise).astype(np.float32)
    vy = (np.float32(0.0) + noise).astype(np.float32)
    wz = (np.float32(0.3) + noise).astype(np.float32)
    init = [vx.copy(), vy.copy(), wz.copy()]
    hist = np.array([[0, 0, 0], [0.1, 0, 0.3], [0.1, 0, 0.3], [0.1, 0, 0.3]], np.float32)
    hist_init = hist.copy()
    p = lambda a: a.ctypes.data_as(abi.f32p)
    oracle_fns["savitsky_golay"](p(vx), p(vy), p(wz), T, p(hist), 0)
    np.testing.assert_allclose(hist[2], hist_init[3], atol=0.02)
    smoothed = np.abs(vx - 0.2).sum() + np.abs(vy).sum() + np.abs(wz - 0.3).sum()
    original = np.abs(init[0] - 0.2).sum() + np.abs(init[1]).sum() + np.abs(init[2] - 0.3).sum()
    assert smoothed < original
    assert hist[3, 0] == vx[0] and hist[3, 1] == vy[0] and hist[3, 2] == wz[0]

    # independent restatement (float64) of utils.hpp:442-605 including the skipped index quirk
    def sg(seq, h):
        s = np.array(seq, np.float64)
        k = np.array([-21, 14, 39, 54, 59, 54, 39, 14, -21], np.float64) / 231.0
        n = len(s) - 1
        ext = lambda idx_list: sum(c * v for c, v in zip(k, idx_list))
        s[0] = ext([h[0], h[1], h[2], h[3], s[0], s[1], s[2], s[3], s[4]])
        s[1] = ext([h[1], h[2], h[3], s[0], s[1], s[2], s[3], s[4], s[5]])
        s[2] = ext([h[2], h[3], s[0], s[1], s[2], s[3], s[4], s[5], s[6]])
        s[3] = ext([h[3], s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7]])
        for i in range(4, n - 4):
            s[i] = ext([s[i - 4], s[i - 3], s[i - 2], s[i - 1], s[i], s[i + 1], s[i + 2], s[i + 3], s[i + 4]])
        i = n - 3   # n-4 is skipped
        s[i] = ext([s[i - 4], s[i - 3], s[i - 2], s[i - 1], s[i], s[i + 1], s[i + 2], s[i + 3], s[i + 3]])
        i += 1
        s[i] = ext([s[i - 4], s[i - 3], s[i - 2], s[i - 1], s[i], s[i + 1], s[i + 2], s[i + 2], s[i + 2]])
        i += 1
        s[i] = ext([s[i - 4], s[i - 3], s[i - 2], s[i - 1], s[i], s[i + 1], s[i + 1], s[i + 1], s[i + 1]])
        i += 1
        s[i] = ext([s[i - 4], s[i - 3], s[i - 2], s[i - 1], s[i], s[i], s[i], s[i], s[i]])
        return s
    np.testing.assert_allclose(vx, sg(init[0], hist_init[:, 0]), atol=2e-6)
    np.testing.assert_allclose(wz, sg(init[2], hist_init[:, 2]), atol=2e-6)
    assert vx[T - 1 - 4] == init[0][T - 1 - 4]   # quirk 12


def test_find_closest_path_pt(oracle_fns):
    """utils.hpp:665-675 incl. the 'returns 0 when the hit is at init' quirk and the defined clamp"""
    import ctypes as C
    from mpcholonavigation_b200 import abi
    vec = np.array([0.0, 0.25, 0.5, 0.75, 1.0], np.float32)
    f = lambda d, init=0: oracle_fns["find_closest_path_pt"](vec.ctypes.data_as(abi.f32p), len(vec), d, init)
    assert f(0.0) == 0
    assert f(0.35) == 1
    assert f(0.40) == 2
    assert f(0.5, 2) == 0        # hit at init -> 0, not init
    assert f(0.625, 2) == 3      # exact tie -> upper
    assert f(0.6, 2) == 2
    assert f(10.0) == 4          # beyond the end: clamp (reference reads *end())
    assert f(0.05, 3) == 0


def test_inflation_compute_cost(oracle_fns):
    f = oracle_fns["inflation_compute_cost"]
    assert f(0.0, 0.05, 0.25, 3.0) == 254
    assert f(5.0, 0.05, 0.25, 3.0) == 253
    assert f(6.0, 0.05, 0.25, 3.0) == int(252 * math.exp(-3.0 * (0.3 - 0.25)))
    assert f(11.0, 0.05, 0.25, 3.0) == int(252 * math.exp(-3.0 * (0.55 - 0.25)))


# ------------------------------------------------------------------------------------------------
# nav2_costmap_2d restatement: second opinion written independently in numpy (UNPINNED by the reference)
# ------------------------------------------------------------------------------------------------
def _bresenham(x0, y0, x1, y1):
    """classic integer Bresenham, both end points included (nav2_util::LineIterator)"""
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx = 1 if x1 >= x0 else -1
    sy = 1 if y1 >= y0 else -1
    pts = []
    x, y = x0, y0
    if dx >= dy:
        num = dx // 2
        for _ in range(dx + 1):
            pts.append((x, y))
            num += dy
            if num >= dx:
                num -= dx
                y += sy
            x += sx
    else:
        num = dy // 2
        for _ in range(dy + 1):
            pts.append((x, y))
            num += dx
            if num >= dy:
                num -= dy
                x += sx
            y += sy
    return pts


def _footprint_cost_np(cm, res, origin, fp, x, y, th):
    """FootprintCollisionChecker::footprintCostAtPose, re-derived for poses whose vertices are all on the map"""
    c, s = math.cos(th), math.sin(th)
    cells = []
    for (fx, fy) in fp:
        wx, wy = x + (fx * c - fy * s), y + (fx * s + fy * c)
        if wx < origin[0] or wy < origin[1]:
            return 254.0
        mx, my = int((wx - origin[0]) / res), int((wy - origin[1]) / res)
        if mx >= cm.shape[1] or my >= cm.shape[0]:
            return 254.0
        cells.append((mx, my))
    n = len(cells)
    edges = [(cells[i], cells[i + 1]) for i in range(n - 1)] + [(cells[0], cells[n - 1])]
    total = 0.0
    for k, (a, b) in enumerate(edges):
        line = 0.0
        for (px, py) in _bresenham(a[0], a[1], b[0], b[1]):
            v = float(cm[py, px])
            if v == 254.0:
                line = 254.0
                break
            line = max(line, v)
        total = max(total, line)
        if total == 254.0 and k < n - 1:
            return total
    return total


def test_footprint_cost_second_opinion(oracle_fns):
    import ctypes as C
    from mpcholonavigation_b200 import abi
    rng = np.random.default_rng(11)
    cm = rng.integers(0, 253, size=(60, 60)).astype(np.uint8)
    cm[rng.random((60, 60)) < 0.02] = 254
    cm[rng.random((60, 60)) < 0.02] = 255
    res, origin = 0.05, (-0.3, 0.2)
    fp = circle_footprint(0.25)
    robot = make_robot(fp, 0.25, 0.25, True, 3.0, False)
    cmap = abi.Costmap()
    cmap.cells = cm.ctypes.data_as(abi.u8p)
    cmap.size_x, cmap.size_y = 60, 60
    cmap.resolution, cmap.origin_x, cmap.origin_y = res, origin[0], origin[1]
    n_lethal = 0
    for _ in range(400):
        x, y, th = rng.uniform(0.2, 2.4), rng.uniform(0.7, 2.9), rng.uniform(-7, 7)
        got = oracle_fns["footprint_cost_at_pose"](C.byref(cmap), C.byref(robot), x, y, th)
        exp = _footprint_cost_np(cm, res, origin, fp, x, y, th)
        # the lazy vertex conversion only matters when a vertex is off-map; those poses are kept on-map here
        assert got == exp, (x, y, th, got, exp)
        n_lethal += got == 254.0
    assert 0 < n_lethal < 400
    # off-map vertex -> LETHAL
    assert oracle_fns["footprint_cost_at_pose"](C.byref(cmap), C.byref(robot), -0.2, 1.0, 0.0) == 254.0
    # world_to_map edges: below origin, exact origin, last cell, just outside
    w2m = lambda wx, wy: oracle_fns["world_to_map"](C.byref(cmap), wx, wy)
    assert w2m(-0.3000001, 0.5) == -1
    assert w2m(-0.3, 0.2) == 0
    assert w2m(-0.3 + 59.5 * res, 0.2 + 59.5 * res) == 59 * 60 + 59
    assert w2m(-0.3 + 60.0 * res, 0.5) == -1
    assert w2m(1e12, 0.5) == -1


def _footprint_cost_device_order_np(cm, res, origin, fp, x, y, th, batch=8):
    """The evaluation ORDER of the CUDA footprint check (mppi_device.cuh: footprint_cost_at_pose / line_cost), in numpy:
    every vertex mapped first (any vertex off the map -> 254 before a single cell is read), then the lines in the reference's
    order, each line's cells taken `batch` at a time: 254 if the batch holds a lethal cell, else the running maximum."""
    c, s = math.cos(th), math.sin(th)
    cells, off = [], False
    for (fx, fy) in fp:
        wx, wy = x + (fx * c - fy * s), y + (fx * s + fy * c)
        if wx < origin[0] or wy < origin[1]:
            off = True
            continue
        qx, qy = (wx - origin[0]) / res, (wy - origin[1]) / res
        if not (qx < cm.shape[1]) or not (qy < cm.shape[0]):
            off = True
            continue
        cells.append((int(qx), int(qy)))
    if off:
        return 254.0

    def line(a, b):
        pts = list(_bresenham(a[0], a[1], b[0], b[1]))
        cost = 0.0
        for i in range(0, len(pts), batch):
            vals = [float(cm[py, px]) for (px, py) in pts[i:i + batch]]
            if 254.0 in vals:
                return 254.0
            cost = max([cost] + vals)
        return cost

    n = len(cells)
    total = 0.0
    for i in range(n - 1):
        total = max(total, line(cells[i], cells[i + 1]))
        if total == 254.0:
            return total
    return max(line(cells[0], cells[n - 1]), total)


@pytest.mark.parametrize("shape", ["rectangle", "circle", "bowtie"])
def test_footprint_check_in_the_device_order_is_the_reference(oracle_fns, shape):
    """The CUDA footprint check maps all vertices before it walks a line and fetches the cells of a line eight at a time
    (mppi_device.cuh); the reference (restated in the oracle) interleaves vertices and lines and stops at the first lethal
    cell.  On maps full of lethal (254) AND unknown (255) cells - where the order of the exits decides between 254 and 255 -
    and for poses whose vertices leave the map, the two evaluation orders give the same cost."""
    import ctypes as C
    from mpcholonavigation_b200 import abi
    rng = np.random.default_rng(23)
    H = W = 50
    res, origin = 0.05, (0.1, -0.2)
    if shape == "rectangle":
        fp = np.array([[0.35, 0.2], [-0.35, 0.2], [-0.35, -0.2], [0.35, -0.2]])   # 14-cell edges: two batches of eight
        robot = make_robot(fp, 0.2, 0.41, True, 3.0, False)
    elif shape == "circle":
        fp = circle_footprint(0.25)
        robot = make_robot(fp, 0.25, 0.25, True, 3.0, False)
    else:
        a = 0.15
        fp = np.array([[a, a], [-a, -a], [a, -a], [-a, a]])
        robot = make_robot(fp, a, a * math.sqrt(2.0), True, 3.0, False)
    seen = set()
    for density in (0.01, 0.04, 0.15):
        cm = rng.integers(0, 253, size=(H, W)).astype(np.uint8)
        cm[rng.random((H, W)) < density] = 254
        cm[rng.random((H, W)) < density] = 255
        cmap = abi.Costmap()
        cmap.cells = cm.ctypes.data_as(abi.u8p)
        cmap.size_x, cmap.size_y = W, H
        cmap.resolution, cmap.origin_x, cmap.origin_y = res, origin[0], origin[1]
        for _ in range(500):
            # poses all over the map and a little beyond its edges (some vertices off the map)
            x, y, th = rng.uniform(-0.1, 2.8), rng.uniform(-0.4, 2.5), rng.uniform(-7, 7)
            got = oracle_fns["footprint_cost_at_pose"](C.byref(cmap), C.byref(robot), x, y, th)
            exp = _footprint_cost_device_order_np(cm, res, origin, fp, x, y, th)
            assert got == exp, (shape, density, x, y, th, got, exp)
            seen.add(got)
    assert 254.0 in seen and 255.0 in seen and any(v < 253.0 for v in seen)


def test_philox_known_answers(oracle_fns):
    """Random123 kat_vectors for philox4x32-10"""
    import ctypes as C
    from mpcholonavigation_b200 import abi
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, exp in kat:
        c = np.array(ctr, np.uint32)
        k = np.array(key, np.uint32)
        o = np.zeros(4, np.uint32)
        oracle_fns["philox4x32_10"](c.ctypes.data_as(abi.u32p), k.ctypes.data_as(abi.u32p), o.ctypes.data_as(abi.u32p))
        assert [int(v) for v in o] == exp


def test_noise_statistics(oracle_fns):
    """noise_generator_test.cpp:39-131: zero-mean, sigma-scaled, vy untouched when non-holonomic"""
    e = Engine(oracle_fns, batch_size=500, time_steps=56, motion_model="Omni", seed=42)
    e.generate_noise(3)
    vx, vy, wz = e.get_noise()
    for a, s in ((vx, 0.2), (vy, 0.2), (wz, 0.4)):
        assert abs(a.mean()) < 4 * s / math.sqrt(a.size)
        assert a.std() == pytest.approx(s, rel=0.02)
    # excess kurtosis of a normal ~ 0
    z = vx / vx.std()
    assert abs((z ** 4).mean() - 3.0) < 0.15
    assert abs(np.corrcoef(vx.ravel(), wz.ravel())[0, 1]) < 0.02
    e2 = Engine(oracle_fns, batch_size=500, time_steps=56, motion_model="DiffDrive", seed=42)
    e2.generate_noise(3)
    vx2, vy2, wz2 = e2.get_noise()
    assert not vy2.any()
    np.testing.assert_array_equal(vx, vx2)
    # sharding: the union of two shards equals the unsharded stream
    a = Engine(oracle_fns, batch_size=250, time_steps=56, motion_model="Omni", seed=42, shard_offset=0, shard_total=500)
    b = Engine(oracle_fns, batch_size=250, time_steps=56, motion_model="Omni", seed=42, shard_offset=250, shard_total=500)
    a.generate_noise(3)
    b.generate_noise(3)
    np.testing.assert_array_equal(np.concatenate([a.get_noise()[2], b.get_noise()[2]]), wz)


def test_det_sincos_vs_libm(oracle_fns):
    import ctypes as C
    xs = np.concatenate([np.linspace(-40, 40, 20001), np.random.default_rng(0).uniform(-1e5, 1e5, 20000)]).astype(np.float32)
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for x in xs:
        oracle_fns["det_sincosf"](float(x), C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - math.sin(float(x))), abs(c.value - math.cos(float(x))))
    assert worst < 1.2e-7
    oracle_fns["det_sincosf"](0.0, C.byref(s), C.byref(c))
    assert s.value == 0.0 and c.value == 1.0


def test_eval_control_is_optimize_then_filter_twist_shift(oracle_fns):
    """Optimizer::evalControl (optimizer.cpp:134-155) = optimize, savitskyGolayFilter, command at `offset`, shift"""
    from mpcholonavigation_b200 import _abi as abi, scenarios
    sc = scenarios.config1(batch=128)
    noise = sc.noise()
    for shift in (False, True):
        a, b = (Engine(oracle_fns, **sc.cfg) for _ in range(2))
        for e in (a, b):
            e.set_robot(sc.robot)
            e.set_critics(sc.critics)
            e.set_noise(*noise)
        hist = np.zeros((4, 3), np.float32)
        for cycle in range(4):
            cmd, ra = a.eval_control(sc.cycle, shift)
            rb = b.optimize(sc.cycle)
            vx, vy, wz = rb.vx.copy(), rb.vy.copy(), rb.wz.copy()
            h = hist.reshape(12).copy()
            oracle_fns["savitsky_golay"](vx.ctypes.data_as(abi.f32p), vy.ctypes.data_as(abi.f32p), wz.ctypes.data_as(abi.f32p),
                                         len(vx), h.ctypes.data_as(abi.f32p), int(shift))
            hist = h.reshape(4, 3)
            off = 1 if shift else 0
            np.testing.assert_array_equal(cmd, np.array([vx[off], vy[off], wz[off]], np.float32))
            if shift:
                vx, vy, wz = (np.concatenate([v[1:], v[-1:]]) for v in (vx, vy, wz))
            b.set_control_sequence(vx, vy, wz)
            np.testing.assert_array_equal(np.stack(a.get_control_sequence()), np.stack([vx, vy, wz]))
            np.testing.assert_array_equal(a.get_control_history(), hist)
        a.close()
        b.close()


# ------------------------------------------------------------------------------------------------
# getOptimizedTrajectory (optimizer.cpp:345-360 -> integrateStateVelocities(xtensor2&, ...) :275-311)
# ------------------------------------------------------------------------------------------------
def test_optimized_trajectory_known_answer_and_second_opinion(oracle_fns):
    """optimizer_unit_tests.cpp:243-248: a fresh Omni optimizer (zero sequence, 50 steps) returns a [50, 3] trajectory of
    zeros.  Then an independent numpy restatement of optimizer.cpp:275-311 on a non-zero sequence: this overload pairs
    cos/sin[1:] with yaws[1:] (`yaw_offseted = view(traj_yaws, range(1, _))`, :294-299) -- there is NO one-step lag,
    unlike the batch overload (:322-329) that the rollout uses."""
    e = Engine(oracle_fns, batch_size=1000, time_steps=50, model_dt=0.1, motion_model="Omni")
    t = e.get_optimized_trajectory((0.0, 0.0, 0.0))
    assert t.shape == (50, 3)
    assert t[5, 0] == 0.0 and t[5, 1] == 0.0 and t[5, 2] == 0.0
    rng = np.random.default_rng(3)
    vx, vy, wz = (rng.uniform(-0.4, 0.5, 50).astype(np.float32) for _ in range(3))
    e.set_control_sequence(vx, vy, wz)
    pose = (1.25, -0.5, 0.3)
    t = e.get_optimized_trajectory(pose)
    dt = np.float32(0.1)
    yaws = np.cumsum(wz * dt, dtype=np.float32) + np.float32(pose[2])
    ang = yaws.copy()
    ang[0] = np.float32(pose[2])                       # cos/sin[0] = cosf/sinf(initial_yaw); [1:] = cos/sin(yaws[1:])
    c, s = np.cos(ang.astype(np.float64)), np.sin(ang.astype(np.float64))
    dx = vx * c - vy * s
    dy = vx * s + vy * c
    x = pose[0] + np.cumsum((dx * dt).astype(np.float32), dtype=np.float32)
    y = pose[1] + np.cumsum((dy * dt).astype(np.float32), dtype=np.float32)
    np.testing.assert_array_equal(t[:, 2], yaws)
    np.testing.assert_allclose(t[:, 0], x, rtol=0, atol=2e-6)
    np.testing.assert_allclose(t[:, 1], y, rtol=0, atol=2e-6)
    # the lagged pairing (what the batch overload does) gives a visibly different curve: the test can tell them apart
    lag = np.concatenate([[np.float32(pose[2])], yaws[:-1]]).astype(np.float64)
    x_lag = pose[0] + np.cumsum(((vx * np.cos(lag) - vy * np.sin(lag)) * dt).astype(np.float32), dtype=np.float32)
    assert np.abs(x_lag - x).max() > 1e-3
    e.close()


def test_iteration_controls_switch_of_the_oracle(oracle_fns):
    """test switch oracle_set_iteration_controls (used by the iteration_count = 2 GPU parity test): pinning the oracle's own
    first update reproduces the free-running two-iteration cycle bit for bit; pinning something else does not"""
    from mpcholonavigation_b200 import scenarios
    f32p = oracle_fns["_lib"].oracle_set_iteration_controls.argtypes[2]
    sc = scenarios.config1(batch=128)
    noise = sc.noise()

    def engine(iterations):
        e = Engine(oracle_fns, **{**sc.cfg, "iteration_count": iterations})
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        e.set_noise(*noise)
        e.set_outputs(trajectories=True, cells=True, critic_costs=True)
        return e

    first = engine(1).optimize(sc.cycle)
    free = engine(2)
    r_free = free.optimize(sc.cycle)
    for delta, same in ((0.0, True), (1e-3, False)):
        pinned = engine(2)
        pin = [np.ascontiguousarray(a + delta, dtype=np.float32) for a in (first.vx, first.vy, first.wz)]
        assert oracle_fns["set_iteration_controls"](pinned.h, 0, *[a.ctypes.data_as(f32p) for a in pin]) == 0
        r = pinned.optimize(sc.cycle)
        assert np.array_equal(r.vx, r_free.vx) == same
        assert np.array_equal(pinned.get_costs(), free.get_costs()) == same

"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d): costmaps, paths, critic lists.

Used by bench.py and by the parity tests so that both measure and check the very same inputs.
The critic list and per-critic overrides are the reference's deployed Omni set
(robot_bringup/config/nav2_params.yaml:222-293); optimizer parameters are the code defaults
(nav2_sortham_controller/src/optimizer.cpp:69-84) unless a config says otherwise.
"""
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .api import Cycle, circle_footprint, make_robot

LETHAL, INSCRIBED = 254, 253


def inflated_disc_costmap(size_x, size_y, resolution, discs, inscribed_radius=0.25, inflation_radius=0.55,
                          cost_scaling_factor=3.0):
    """Free space + lethal discs, inflated the way nav2_costmap_2d::InflationLayer::computeCost does:
    0 cells -> 254, d <= r_insc -> 253, else (uint8)(252 * exp(-scale * (d - r_insc))) for d <= radius."""
    yy, xx = np.mgrid[0:size_y, 0:size_x]
    lethal = np.zeros((size_y, size_x), bool)
    for (cx, cy, r) in discs:
        lethal |= (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
    cost = np.zeros((size_y, size_x), np.uint8)
    if not lethal.any():
        return cost
    # exact euclidean distance (in cells) to the nearest lethal cell, brute force over a bounded window
    reach = int(np.ceil(inflation_radius / resolution)) + 1
    dist = np.full((size_y, size_x), np.inf)
    ly, lx = np.nonzero(lethal)
    for dy in range(-reach, reach + 1):
        for dx in range(-reach, reach + 1):
            d = np.hypot(dx, dy)
            if d * resolution > inflation_radius:
                continue
            ty, tx = ly + dy, lx + dx
            ok = (ty >= 0) & (ty < size_y) & (tx >= 0) & (tx < size_x)
            np.minimum.at(dist, (ty[ok], tx[ok]), d)
    m = np.isfinite(dist)
    d_m = dist * resolution
    val = np.zeros_like(dist)
    val[m] = np.floor(252.0 * np.exp(-cost_scaling_factor * (d_m[m] - inscribed_radius)))
    val[m & (d_m <= inscribed_radius)] = INSCRIBED
    val[dist == 0] = LETHAL
    cost[m] = val[m].astype(np.uint8)
    return cost


def random_discs(rng, k, size_x, size_y, resolution, keep_out: List[Tuple[float, float, float]], rmin=2, rmax=4):
    """k discs (cx, cy, r) in cells, centres uniform over the map but outside the keep-out circles
    (x [m], y [m], radius [m])."""
    discs = []
    guard = 0
    while len(discs) < k and guard < 100000:
        guard += 1
        cx, cy = rng.uniform(0, size_x), rng.uniform(0, size_y)
        r = rng.uniform(rmin, rmax)
        wx, wy = cx * resolution, cy * resolution
        if all(np.hypot(wx - kx, wy - ky) >= kr + r * resolution for (kx, ky, kr) in keep_out):
            discs.append((cx, cy, r))
    return discs


def omni_default_critics(cost_consider_footprint=True):
    """critics: [...] of nav2_params.yaml:222 with the overrides of :223-293 (dead keys dropped)."""
    return [
        ("ConstraintCritic", dict(cost_power=1, cost_weight=4.0)),
        ("CostCritic", dict(cost_power=1, cost_weight=3.81, critical_cost=300.0,
                            consider_footprint=int(cost_consider_footprint), collision_cost=1000000.0,
                            near_goal_distance=1.0)),
        ("GoalCritic", dict(cost_power=1, cost_weight=5.0, threshold_to_consider=1.4)),
        ("GoalAngleCritic", dict(cost_power=1, cost_weight=3.0, threshold_to_consider=0.5)),
        ("PathAlignCritic", dict(cost_power=1, cost_weight=14.0, max_path_occupancy_ratio=0.05,
                                 trajectory_point_step=4, threshold_to_consider=0.5, offset_from_furthest=20,
                                 use_path_orientations=0)),
        ("PathFollowCritic", dict(cost_power=1, cost_weight=5.0, offset_from_furthest=5,
                                  threshold_to_consider=1.4)),
        ("PathAngleCritic", dict(cost_power=1, cost_weight=2.0, offset_from_furthest=4,
                                 threshold_to_consider=0.5, max_angle_to_furthest=1.0)),
        ("PreferForwardCritic", dict(cost_power=1, cost_weight=5.0, threshold_to_consider=0.5)),
        ("TwirlingCritic", dict()),   # yaml keys twirling_cost_* are never read -> defaults 1 / 10.0
    ]


@dataclass
class Scenario:
    name: str
    cfg: dict
    critics: list
    robot: object
    cycle: Cycle
    noise_seed: int

    def noise(self):
        """Reference-style injected noise: one set, reused every cycle (noise_generator.cpp:35-41)."""
        B, T = self.cfg["batch_size"], self.cfg["time_steps"]
        rng = np.random.default_rng(self.noise_seed)
        z = rng.standard_normal((3, B, T)).astype(np.float32)
        return (z[0] * np.float32(self.cfg["vx_std"]), z[1] * np.float32(self.cfg["vy_std"]),
                z[2] * np.float32(self.cfg["wz_std"]))


def straight_path(x0, y0, heading, n, step):
    s = np.arange(n, dtype=np.float64) * step
    px = (x0 + s * np.cos(heading)).astype(np.float32)
    py = (y0 + s * np.sin(heading)).astype(np.float32)
    pyaw = np.full(n, heading, np.float32)
    return px, py, pyaw


def _omni_cfg(batch, steps, **kw):
    cfg = dict(batch_size=batch, time_steps=steps, model_dt=0.05, iteration_count=1, temperature=0.3, gamma=0.015,
               vx_max=0.5, vx_min=-0.35, vy_max=0.5, wz_max=1.9, vx_std=0.2, vy_std=0.2, wz_std=0.4,
               motion_model="Omni")
    cfg.update(kw)
    return cfg


def config1(batch=1000, steps=56, map_size=100, n_path=40, heading=0.0, map_seed=0, noise_seed=1, k_discs=6,
            footprint="circle", cost_consider_footprint=True, pose=None, ring=None):
    """BASELINE configs[0]/[1]: Omni 1000x56, dt 0.05, default critic set, 100x100 costmap @0.05 m,
    straight 40-point path.  ring = r [m] with footprint="rectangle": the robot boxed in by eight discs r away
    (two of them across the path), so that CostCritic's footprint branch (cost_critic.cpp:204-209) is taken every cycle."""
    res = 0.05
    pose = pose if pose is not None else (map_size * res / 2.0, map_size * res / 2.0, heading)
    px, py, pyaw = straight_path(pose[0], pose[1], heading, n_path, 0.05)
    rng = np.random.default_rng(map_seed)
    keep = [(pose[0], pose[1], 0.6 if ring is None else 0.9)] + [(float(x), float(y), 0.3) for x, y in zip(px, py)]
    discs = random_discs(rng, k_discs, map_size, map_size, res, keep)
    if ring is not None:
        for k in range(8):
            a = heading + 2.0 * np.pi * (k + 0.5) / 8
            discs.append((pose[0] / res + ring / res * np.cos(a), pose[1] / res + ring / res * np.sin(a), 2.0))
    cm = inflated_disc_costmap(map_size, map_size, res, discs, inscribed_radius=0.15 if footprint == "rectangle" else 0.25)
    if footprint == "rectangle":
        fp = np.array([[0.25, 0.15], [-0.25, 0.15], [-0.25, -0.15], [0.25, -0.15]])
        insc, circ = 0.15, float(np.hypot(0.25, 0.15))
    elif footprint == "circle":
        fp = circle_footprint(0.25)
        insc = circ = 0.25
    else:  # the bow-tie square of test/utils/factory.hpp:116-119
        a = 0.15
        fp = np.array([[a, a], [-a, -a], [a, -a], [-a, a]])
        insc, circ = a, a * np.sqrt(2.0)
    robot = make_robot(fp, inscribed_radius=insc, circumscribed_radius=circ, inflation_layer_found=True,
                       inflation_cost_scaling_factor=3.0, track_unknown=False)
    cyc = Cycle(pose=pose, speed=(0.0, 0.0, 0.0), goal=(float(px[-1]), float(py[-1])),
                goal_checker_xy_tolerance=0.25, path_x=px, path_y=py, path_yaw=pyaw, costmap=cm, resolution=res,
                origin=(0.0, 0.0))
    return Scenario("omni_%dx%d" % (batch, steps), _omni_cfg(batch, steps), omni_default_critics(cost_consider_footprint),
                    robot, cyc, noise_seed)


def config3(batch=16384, steps=56, map_size=400, k_discs=120, map_seed=2, noise_seed=3, dense=False, ring=0.47):
    """BASELINE configs[2]: 16384x56, 400x400 costmap @0.05 m, ObstaclesCritic alone in footprint mode.

    dense=True is the same map with the robot boxed in: a 0.5 x 0.3 m rectangular footprint (inscribed radius 0.15 m,
    circumscribed 0.29 m) inside a ring of eight discs `ring` metres away, so that the footprint branch of the critic
    (obstacles_critic.cpp:214-220: point cost >= the cost at the circumscribed radius) is taken by a large share of the poses
    in every cycle - with SURVEY 8d's literal geometry (a 16-gon whose circumscribed radius IS its inscribed radius, one
    disc 0.5 m away) it is taken by 0.2 % of the poses of the first cycle and by none once the sequence has moved away."""
    res = 0.05
    centre = map_size * res / 2.0
    pose = (centre, centre, 0.0)
    px, py, pyaw = straight_path(pose[0], pose[1], 0.0, 40, 0.05)
    rng = np.random.default_rng(map_seed)
    keep = [(pose[0], pose[1], 0.45 if not dense else 0.9)]
    discs = random_discs(rng, k_discs - 1, map_size, map_size, res, keep)
    # one disc placed 0.5 m from the robot so that many poses take the footprint branch
    discs.append((pose[0] / res + 0.5 / res * np.cos(0.6), pose[1] / res + 0.5 / res * np.sin(0.6), 3.0))
    if dense:
        discs.pop()
        for k in range(8):
            a = 2.0 * np.pi * (k + 0.5) / 8
            discs.append((pose[0] / res + ring / res * np.cos(a), pose[1] / res + ring / res * np.sin(a), 2.0))
        cm = inflated_disc_costmap(map_size, map_size, res, discs, inscribed_radius=0.15)
        fp = np.array([[0.25, 0.15], [-0.25, 0.15], [-0.25, -0.15], [0.25, -0.15]])
        robot = make_robot(fp, inscribed_radius=0.15, circumscribed_radius=float(np.hypot(0.25, 0.15)),
                           inflation_layer_found=True, inflation_cost_scaling_factor=3.0, track_unknown=False)
        critics = [("ObstaclesCritic", dict(consider_footprint=1, inflation_radius=0.55, cost_scaling_factor=3.0))]
        cyc = Cycle(pose=pose, goal=(float(px[-1]), float(py[-1])), goal_checker_xy_tolerance=0.25, path_x=px, path_y=py,
                    path_yaw=pyaw, costmap=cm, resolution=res, origin=(0.0, 0.0))
        return Scenario("obstacles_fp_dense_%dx%d" % (batch, steps), _omni_cfg(batch, steps), critics, robot, cyc, noise_seed)
    cm = inflated_disc_costmap(map_size, map_size, res, discs)
    robot = make_robot(circle_footprint(0.25), inscribed_radius=0.25, circumscribed_radius=0.25,
                       inflation_layer_found=True, inflation_cost_scaling_factor=3.0, track_unknown=False)
    critics = [("ObstaclesCritic", dict(consider_footprint=1, inflation_radius=0.55, cost_scaling_factor=3.0))]
    cyc = Cycle(pose=pose, goal=(float(px[-1]), float(py[-1])), goal_checker_xy_tolerance=0.25, path_x=px, path_y=py,
                path_yaw=pyaw, costmap=cm, resolution=res, origin=(0.0, 0.0))
    return Scenario("obstacles_fp_%dx%d" % (batch, steps), _omni_cfg(batch, steps), critics, robot, cyc, noise_seed)


def config4(batch=262144, steps=100, map_size=400, n_path=120, map_seed=4, noise_seed=5):
    """BASELINE configs[3]: 262144x100, default critic set, 400x400 map, path N=120 (sharded over GPUs)."""
    res = 0.05
    centre = map_size * res / 2.0
    pose = (centre, centre, 0.0)
    px, py, pyaw = straight_path(pose[0], pose[1], 0.0, n_path, 0.05)
    rng = np.random.default_rng(map_seed)
    keep = [(pose[0], pose[1], 0.6)] + [(float(x), float(y), 0.3) for x, y in zip(px, py)]
    discs = random_discs(rng, 120, map_size, map_size, res, keep)
    cm = inflated_disc_costmap(map_size, map_size, res, discs)
    robot = make_robot(circle_footprint(0.25), inscribed_radius=0.25, circumscribed_radius=0.25,
                       inflation_layer_found=True, inflation_cost_scaling_factor=3.0, track_unknown=False)
    cyc = Cycle(pose=pose, goal=(float(px[-1]), float(py[-1])), goal_checker_xy_tolerance=0.25, path_x=px, path_y=py,
                path_yaw=pyaw, costmap=cm, resolution=res, origin=(0.0, 0.0))
    return Scenario("omni_%dx%d" % (batch, steps), _omni_cfg(batch, steps), omni_default_critics(True), robot, cyc,
                    noise_seed)


def config5_robot(r, n_robots=256, batch=2000, steps=56):
    """BASELINE configs[4]: robot r of 256 independent scenarios (own map rng(100+r), own heading)."""
    heading = 2.0 * np.pi * r / n_robots
    sc = config1(batch=batch, steps=steps, heading=heading, map_seed=100 + r, noise_seed=1000 + r)
    sc.name = "robot%d_%dx%d" % (r, batch, steps)
    return sc

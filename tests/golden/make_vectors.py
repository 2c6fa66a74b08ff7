#!/usr/bin/env python
"""Generates tests/golden/softmax_directed.json (independent numpy float64 formula) and
tests/golden/oracle_regression_v1.npz (oracle outputs, drift alarm).  Run from the repo root."""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def softmax_update(cs, noise, costs, gamma, std, temperature, limits, holonomic=True):
    """Optimizer::updateControlSequence + applyControlSequenceConstraints (optimizer.cpp:362-394, 237-249), float64.
    cs [3][T] (vx, vy, wz), noise [3][B][T], costs [B] (critic costs before the gamma term)"""
    cs = np.asarray(cs, np.float64)
    c = cs[:, None, :] + np.asarray(noise, np.float64)                 # controls [3][B][T]
    costs = np.asarray(costs, np.float64).copy()
    planes = (0, 2, 1) if holonomic else (0, 2)
    for p in planes:
        costs += gamma / std[p] ** 2 * np.sum(cs[p][None, :] * (c[p] - cs[p][None, :]), axis=1)
    w = np.exp(-(costs - costs.min()) / temperature)
    w /= w.sum()
    new = cs.copy()
    for p in planes:
        new[p] = (w[:, None] * c[p]).sum(axis=0)
    vx_max, vx_min, vy, wz = limits
    new[0] = np.clip(new[0], vx_min, vx_max)
    if holonomic:
        new[1] = np.clip(new[1], -vy, vy)
    new[2] = np.clip(new[2], -wz, wz)
    return new, costs, w


def directed_cases():
    rng = np.random.default_rng(7)
    cases = []
    for name, hol, temperature, scale in (("omni_t0.3", True, 0.3, 0.3), ("diff_t1.0", False, 1.0, 0.3),
                                          ("omni_clipped", True, 0.05, 1.5)):
        B, T = 3, 4
        cs = np.round(rng.uniform(-0.3, 0.3, (3, T)), 3)
        if not hol:
            cs[1] = 0.0
        noise = np.round(rng.normal(0, scale, (3, B, T)), 3)
        if not hol:
            noise[1] = 0.0
        std = (0.2, 0.2, 0.4)
        limits = (0.5, -0.35, 0.5, 1.9)
        new, costs, w = softmax_update(cs, noise, np.zeros(B), 0.015, std, temperature, limits, hol)
        cases.append(dict(name=name, holonomic=hol, temperature=temperature, gamma=0.015, std=std, limits=limits,
                          control_sequence=cs.tolist(), noise=noise.tolist(), expected_controls=new.tolist(),
                          expected_costs=costs.tolist(), expected_weights=w.tolist()))
    return cases


def oracle_regression():
    from mpcholonavigation_b200 import Engine, scenarios
    from tests import oracle_loader
    sc = scenarios.config1(batch=96, steps=56)
    e = Engine(oracle_loader.load(), **sc.cfg)
    e.set_robot(sc.robot)
    e.set_critics(sc.critics)
    e.set_noise(*sc.noise())
    e.set_outputs(trajectories=True, cells=True, critic_costs=True)
    out = {}
    for cycle in range(3):
        r = e.optimize(sc.cycle)
        out[f"controls_{cycle}"] = np.stack([r.vx, r.vy, r.wz])
        out[f"costs_{cycle}"] = e.get_costs()
        out[f"cells_crc_{cycle}"] = np.array([zlib.crc32(e.get_cells().tobytes())], np.uint32)
        out[f"furthest_{cycle}"] = np.array([-1 if r.furthest_reached_path_point is None else r.furthest_reached_path_point])
    e.close()
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "softmax_directed.json"), "w") as f:
        json.dump(directed_cases(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_regression_v1.npz"), **oracle_regression())
    print("written")

"""Fixtures under tests/golden/: the reference's known answers are all executed (golden_cases.py), the directed
softmax vectors (independent numpy formula) and the oracle drift alarm hold on the oracle (CPU) and on the GPU."""
import json
import os
import zlib

import numpy as np
import pytest

from mpcholonavigation_b200 import Cycle, Engine, scenarios
from tests.golden_cases import golden_cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_every_reference_known_answer_is_executed():
    listed = {c["case"] for c in json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))["cases"]}
    executed = {c.__name__ for c in golden_cases()}
    assert listed == executed, (listed - executed, executed - listed)


def _directed(fns):
    for case in json.load(open(os.path.join(GOLDEN, "softmax_directed.json"))):
        noise = np.asarray(case["noise"], np.float32)
        B, T = noise.shape[1], noise.shape[2]
        vx_max, vx_min, vy, wz = case["limits"]
        e = Engine(fns, batch_size=B, time_steps=T, motion_model="Omni" if case["holonomic"] else "DiffDrive",
                   temperature=case["temperature"], gamma=case["gamma"], vx_std=case["std"][0], vy_std=case["std"][1],
                   wz_std=case["std"][2], vx_max=vx_max, vx_min=vx_min, vy_max=vy, wz_max=wz)
        e.set_critics([])
        e.set_noise(noise[0], noise[1], noise[2])
        cs = np.asarray(case["control_sequence"], np.float32)
        e.set_control_sequence(cs[0], cs[1], cs[2])
        r = e.optimize(Cycle(path_x=np.zeros(2, np.float32), path_y=np.zeros(2, np.float32), path_yaw=np.zeros(2, np.float32)))
        exp = np.asarray(case["expected_controls"])
        np.testing.assert_allclose(np.stack([r.vx, r.vy, r.wz]), exp, rtol=1e-4, atol=2e-6, err_msg=case["name"])
        np.testing.assert_allclose(e.get_costs(), case["expected_costs"], rtol=1e-4, atol=1e-6, err_msg=case["name"])
        e.close()


def test_directed_softmax_vectors_oracle(oracle_fns):
    _directed(oracle_fns)


@pytest.mark.gpu
def test_directed_softmax_vectors_gpu(product_fns):
    _directed(product_fns)


def _regression(fns, exact):
    ref = np.load(os.path.join(GOLDEN, "oracle_regression_v1.npz"))
    sc = scenarios.config1(batch=96, steps=56)
    e = Engine(fns, **sc.cfg)
    e.set_robot(sc.robot)
    e.set_critics(sc.critics)
    e.set_noise(*sc.noise())
    e.set_outputs(trajectories=True, cells=True, critic_costs=True)
    for cycle in range(3):
        r = e.optimize(sc.cycle)
        got = np.stack([r.vx, r.vy, r.wz])
        assert zlib.crc32(e.get_cells().tobytes()) == int(ref[f"cells_crc_{cycle}"][0]), f"cycle {cycle}: cell indices"
        assert (-1 if r.furthest_reached_path_point is None else r.furthest_reached_path_point) == int(ref[f"furthest_{cycle}"][0])
        if exact:
            np.testing.assert_array_equal(got, ref[f"controls_{cycle}"])
            np.testing.assert_array_equal(e.get_costs(), ref[f"costs_{cycle}"])
        else:
            np.testing.assert_allclose(got, ref[f"controls_{cycle}"], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(e.get_costs(), ref[f"costs_{cycle}"], rtol=1e-4, atol=2e-5)
            # same warm start as the fixture so that the next cycle stays comparable point-wise
            e.set_control_sequence(*ref[f"controls_{cycle}"])
    e.close()


def test_oracle_regression_vectors(oracle_fns):
    _regression(oracle_fns, exact=True)


@pytest.mark.gpu
def test_gpu_matches_committed_vectors(product_fns):
    _regression(product_fns, exact=False)

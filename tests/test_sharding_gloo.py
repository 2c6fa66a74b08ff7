"""N>1 path on CPU: two gloo ranks run the sharded protocol of mpcholonavigation_b200/sharding.py on halves of
one problem whose per-trajectory quantities come from the CPU oracle, and must reproduce the oracle's
unsharded optimize().  (The CUDA implementation of the same protocol is covered by tests/test_gpu_sharded.py.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpcholonavigation_b200 import Engine, scenarios, sharding

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    sc = scenarios.config1(batch=256, steps=56)
    return sc, sc.noise()


def _rank_main(rank, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from tests import oracle_loader
    fns = oracle_loader.load()
    sc, noise = _problem()
    B, T = sc.cfg["batch_size"], sc.cfg["time_steps"]
    b0, b1 = sharding.shard_bounds(B, rank, WORLD)
    # every rank integrates and scores ITS trajectories only (critic-level entry points of the ABI)
    cfg = dict(sc.cfg, batch_size=b1 - b0)
    e = Engine(fns, **cfg)
    e.set_robot(sc.robot)
    cs = np.zeros(T, np.float32)
    # controls = control sequence (zero at cycle 1) + noise; state velocities = shifted controls
    ctrl = [n[b0:b1] + cs for n in noise]
    vel = [np.concatenate([np.zeros((b1 - b0, 1), np.float32), c[:, :-1]], 1) for c in ctrl]
    x, y, yaw = e.integrate_state_velocities(sc.cycle.pose, *vel)
    # exchange 1: local furthest point through the path critics of the shard, then MAX over ranks
    e.set_critics([("PathFollowCritic", dict(offset_from_furthest=5))])
    _, fur_local, _ = e.score_trajectories(sc.cycle, *vel, x, y, yaw)
    words = torch.zeros(17, dtype=torch.int64)
    words[0] = fur_local
    dist.all_reduce(words, op=dist.ReduceOp.MAX)
    furthest = int(words[0])
    # score the shard with the GLOBAL furthest point (what K3 does after exchange 1)
    e.set_critics(sc.critics)
    costs, fur_after, fail = e.score_trajectories(sc.cycle, *vel, x, y, yaw, furthest=furthest)
    assert fur_after == furthest and not fail
    # gamma term of updateControlSequence is zero at cycle 1 (control sequence is zero)
    rec = sharding.local_partial(costs, ctrl, sc.cfg["temperature"])
    gathered = [torch.zeros(3 * T + 2, dtype=torch.float64) for _ in range(WORLD)]
    dist.all_gather(gathered, torch.from_numpy(rec))
    merged = sharding.merge_partials(np.stack([g.numpy() for g in gathered]), sc.cfg["temperature"])
    cvx, cvy, cwz = sharding.controls_from_record(merged, T, sc.cfg["vx_min"], sc.cfg["vx_max"], sc.cfg["vy_max"],
                                                  sc.cfg["wz_max"])
    if rank == 0:
        np.savez(out_path, vx=cvx, vy=cvy, wz=cwz, furthest=furthest)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_protocol_reproduces_unsharded_oracle(oracle_fns, tmp_path):
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_rank_main, args=(_free_port(), out), nprocs=WORLD, join=True)
    got = np.load(out)
    sc, noise = _problem()
    e = Engine(oracle_fns, **sc.cfg)
    e.set_robot(sc.robot)
    e.set_critics(sc.critics)
    e.set_noise(*noise)
    ref = e.optimize(sc.cycle)
    assert int(got["furthest"]) == ref.furthest_reached_path_point
    np.testing.assert_allclose(got["vx"], ref.vx, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got["vy"], ref.vy, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got["wz"], ref.wz, rtol=1e-4, atol=1e-6)


def test_merge_is_exact_online_softmax():
    rng = np.random.default_rng(0)
    costs = rng.uniform(0, 50, 1000)
    ctrl = rng.standard_normal((3, 1000, 8))
    full = sharding.local_partial(costs, ctrl, 0.3)
    parts = [sharding.local_partial(costs[i:i + 125], ctrl[:, i:i + 125], 0.3) for i in range(0, 1000, 125)]
    merged = sharding.merge_partials(np.stack(parts), 0.3)
    np.testing.assert_allclose(merged, full, rtol=1e-12)
    # order of the shards does not matter
    merged2 = sharding.merge_partials(np.stack(parts[::-1]), 0.3)
    np.testing.assert_allclose(merged2, full, rtol=1e-12)
    w = sharding.merge_exchange1(np.array([[3, 0, 1] + [0] * 14, [7, 1, 0] + [0] * 14]))
    assert list(w[:3]) == [7, 1, 1]

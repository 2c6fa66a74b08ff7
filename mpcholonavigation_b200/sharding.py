"""The two exchanges of a sharded optimize(), stated in numpy.

SPECIFICATION / TEST SUPPORT, not product code: the product does these exchanges inside its kernels (NVLink peer
mailboxes) or with NCCL (mppi_api.cu); tests/test_sharding_gloo.py runs this numpy statement at world_size 2 over gloo
against the unsharded oracle, which is what the device protocol is checked against.

Trajectories are independent through rollout and scoring, so a batch shards over ranks by trajectory index.
Two data-dependent global scalars split the pipeline:

  exchange 1 (after the rollout kernel):  furthest_reached_path_point = MAX over ranks of the local
      max-over-trajectories argmin (utils.hpp:292-319), and one "some trajectory survived" flag per
      obstacle-type critic (fail_flag = AND over ranks of "all collide").  One MAX all-reduce of 17 words.
  exchange 2 (after the update kernel):   per-rank softmax partial (m, s, W[3T]) with m = min cost,
      s = sum exp(-(c - m)/temperature), W = sum exp(...) * control; merged with the online-softmax rule.
      One all-gather of 3T+2 floats; every rank merges redundantly, so no broadcast follows.

The CUDA library implements exactly this (mppi_api.cu: NCCL in enqueue_kernels, host-mediated copies in
mppi_optimize_sharded; kernels merge_finalize_kernel / path_softmax_update_kernel).  This module is the
executable statement of the protocol used by the CPU (gloo) tests and by DESIGN.md.
"""
import numpy as np


def shard_bounds(total, rank, world):
    """rank r owns trajectories [r * total / world, (r + 1) * total / world)"""
    per = total // world
    assert per * world == total, "batch must divide evenly over the ranks"
    return rank * per, (rank + 1) * per


def local_partial(costs, controls, temperature):
    """costs [b], controls [3, b, T] -> record [m, s, W_vx[T], W_vy[T], W_wz[T]] (float64 for the spec)"""
    costs = np.asarray(costs, np.float64)
    m = costs.min()
    w = np.exp(-(costs - m) / temperature)
    rec = [np.array([m, w.sum()])]
    for p in range(3):
        rec.append((w[:, None] * np.asarray(controls[p], np.float64)).sum(0))
    return np.concatenate(rec)


def merge_partials(records, temperature):
    """records [n, 3T+2] -> merged record; M = min m_r, S = sum s_r e^{-(m_r-M)/temp}, W likewise"""
    records = np.asarray(records, np.float64)
    m = records[:, 0].min()
    scale = np.exp(-(records[:, 0] - m) / temperature)
    out = (records[:, 1:] * scale[:, None]).sum(0)
    return np.concatenate([[m], out])


def controls_from_record(rec, T, vx_min, vx_max, vy, wz, holonomic=True):
    """cs = W / S, then Optimizer::applyControlSequenceConstraints (optimizer.cpp:237-249)"""
    s = rec[1]
    cvx = np.clip(rec[2:2 + T] / s, vx_min, vx_max)
    cvy = np.clip(rec[2 + T:2 + 2 * T] / s, -vy, vy) if holonomic else np.zeros(T)
    cwz = np.clip(rec[2 + 2 * T:2 + 3 * T] / s, -wz, wz)
    return cvx, cvy, cwz


def merge_exchange1(words):
    """words [n, 17] uint32 (furthest candidate, survivor flag per critic slot) -> element-wise MAX"""
    return np.asarray(words, np.uint32).max(0)

#!/usr/bin/env python
"""Correctness of the sharded optimize(), one rank per GPU (launch with torchrun): CHECK_EXCHANGE=peer (default, exchanges over
peer-mapped mailboxes fused into the kernels) or CHECK_EXCHANGE=nccl (all-reduce + all-gather baseline).

Each rank owns B/world trajectories (Philox noise by global index).  Rank 0 additionally solves the same
problem in-process (mppi_optimize_sharded with `world` shards on its own GPU) and the control sequences of all
ranks must match it.  Prints one line per rank and exits non-zero on mismatch.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, abi, load_product, optimize_sharded, scenarios, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fns = load_product()
    B = int(os.environ.get("CHECK_BATCH", "65536"))
    sc = scenarios.config4(batch=B, steps=100)
    b0, b1 = sharding.shard_bounds(B, rank, world)

    def make(r0, r1, device):
        e = Engine(fns, **dict(sc.cfg, batch_size=r1 - r0, shard_offset=r0, shard_total=B, seed=7, device=device))
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        e.generate_noise(0)
        return e

    e = make(b0, b1, local)
    exchange = os.environ.get("CHECK_EXCHANGE", "peer")
    if exchange == "nccl":
        uid = torch.zeros(abi.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * abi.NCCL_UNIQUE_ID_BYTES)()
            assert fns["comm_get_unique_id"](buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        e.comm_init(bytes(uid.cpu().tolist()), rank, world)
    else:
        mine_h = torch.tensor(list(e.comm_mailbox_handle()), dtype=torch.uint8, device="cuda")
        all_h = [torch.zeros_like(mine_h) for _ in range(world)]
        dist.all_gather(all_h, mine_h)
        e.comm_connect_peers([bytes(t.cpu().tolist()) for t in all_h], rank, world)
        dist.barrier()
    ref = None
    if rank == 0:
        ref = [make(*sharding.shard_bounds(B, r, world), local) for r in range(world)]
    ok = True
    for cycle in range(int(os.environ.get("CHECK_CYCLES", "6"))):
        r = e.optimize(sc.cycle)
        mine = torch.tensor(np.concatenate([r.vx, r.vy, r.wz]), device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        if rank == 0:
            rr = optimize_sharded(ref, sc.cycle)
            want = np.concatenate([rr.vx, rr.vy, rr.wz])
            for k, v in enumerate(allv):
                got = v.cpu().numpy()
                good = np.allclose(got, want, rtol=1e-4, atol=1e-6)
                ok = ok and good
                print(f"cycle {cycle} rank {k}: max abs diff vs in-process reference {np.abs(got - want).max():.3e} "
                      f"{'OK' if good else 'MISMATCH'}  device_ms={r.device_ms:.3f} furthest={r.furthest_reached_path_point}")
    if rank == 0:
        print("exchange:", exchange)
    dist.barrier()
    e.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()

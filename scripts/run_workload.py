#!/usr/bin/env python
"""Runs a few optimize() cycles of one workload through the C ABI (no torch import): the command ncu wraps.

  python scripts/run_workload.py --workload omni_1000x56 --cycles 5 [--batch B] [--steps T] [--resident]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="omni_1000x56")
    ap.add_argument("--cycles", type=int, default=5)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--resident", action="store_true")
    a = ap.parse_args()
    kw = {}
    if a.batch:
        kw["batch"] = a.batch
    if a.steps:
        kw["steps"] = a.steps
    if a.workload == "omni_1000x56":
        sc, philox = scenarios.config1(**kw), False
    elif a.workload == "obstacles_16384x56":
        sc, philox = scenarios.config3(**kw), False
    elif a.workload == "obstacles_dense_16384x56":
        sc, philox = scenarios.config3(dense=True, **kw), False
    elif a.workload == "sharded_262144x100":
        sc, philox = scenarios.config4(**kw), True
    else:
        raise SystemExit("unknown workload")
    e = Engine(load_product(), **sc.cfg)
    e.set_robot(sc.robot)
    e.set_critics(sc.critics)
    if philox:
        e.generate_noise(0)
    else:
        e.set_noise(*sc.noise())
    if a.resident:
        e.upload_cycle(sc.cycle)
    ms = []
    for _ in range(a.cycles):
        r = e.optimize_resident() if a.resident else e.optimize(sc.cycle)
        ms.append(r.device_ms)
    print("workload %s B=%d T=%d device_ms: %s" % (sc.name, e.B, e.T, " ".join("%.4f" % m for m in ms)))
    e.close()


if __name__ == "__main__":
    main()

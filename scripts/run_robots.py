#!/usr/bin/env python
"""Runs a few steps of BASELINE configs[4] (256 robots x (2000 x 56) as one bound group) through the C ABI, no torch import:
the command ncu wraps for the counters of the four batched kernels (scripts/collect_counters.sh).

  python scripts/run_robots.py [--robots 256] [--cycles 12]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402
from mpcholonavigation_b200 import _abi as abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", type=int, default=256)
    ap.add_argument("--cycles", type=int, default=12)
    a = ap.parse_args()
    n = a.robots
    fns = load_product()
    robots = [scenarios.config5_robot(i) for i in range(n)]
    engines = []
    for sc in robots:
        e = Engine(fns, **sc.cfg)
        e.set_robot(sc.robot); e.set_critics(sc.critics); e.set_noise(*sc.noise())
        engines.append(e)
    T = robots[0].cfg["time_steps"]
    hs = (abi.H * n)(*[e.h for e in engines])
    assert fns["batch_bind"](hs, n) == 0
    outs = (abi.CycleOut * n)()
    keep = []
    for i in range(n):
        arrs = [np.empty(T, np.float32) for _ in range(3)]
        outs[i].control_vx, outs[i].control_vy, outs[i].control_wz = (x.ctypes.data_as(abi.f32p) for x in arrs)
        keep.append(arrs)
    for e, sc in zip(engines, robots):
        e.upload_cycle(sc.cycle)
    span = C.c_float(0.0)
    ms = []
    for _ in range(a.cycles):
        assert fns["optimize_batch_resident"](hs, outs, n) == 0
        assert fns["batch_span_ms"](hs, n, C.byref(span)) == 0
        ms.append(span.value)
    print("robots %d x (%d x %d): device span per step (ms): %s" % (
        n, robots[0].cfg["batch_size"], T, " ".join("%.4f" % m for m in ms)))
    for e in engines:
        e.close()


if __name__ == "__main__":
    main()

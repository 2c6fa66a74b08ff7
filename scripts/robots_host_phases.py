#!/usr/bin/env python
"""Host-side phases of mppi_optimize_batch over the robots of BASELINE configs[4] (tuning aid)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402
from mpcholonavigation_b200 import _abi as abi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fns = load_product()
robots = [scenarios.config5_robot(i) for i in range(n)]
engines = []
for sc in robots:
    e = Engine(fns, **sc.cfg)
    e.set_robot(sc.robot); e.set_critics(sc.critics); e.set_noise(*sc.noise()); e.set_timing(False)
    engines.append(e)
T = robots[0].cfg["time_steps"]
hs = (abi.H * n)(*[e.h for e in engines])
ins = (abi.CycleIn * n)()
outs = (abi.CycleOut * n)()
keep = []
for i, sc in enumerate(robots):
    cin, k = sc.cycle.pack()
    ins[i] = cin
    arrs = [np.empty(T, np.float32) for _ in range(3)]
    outs[i].control_vx, outs[i].control_vy, outs[i].control_wz = (a.ctypes.data_as(abi.f32p) for a in arrs)
    keep.append((k, arrs))
lib = fns["_lib"]
lib.mppi_debug_get_host_ns.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32]
buf = (C.c_uint64 * 8)()
for _ in range(5):
    assert fns["optimize_batch"](hs, ins, outs, n) == 0
for e in engines:
    lib.mppi_debug_get_host_ns(e.h, buf, 1)
t = []
for _ in range(30):
    t0 = time.perf_counter()
    assert fns["optimize_batch"](hs, ins, outs, n) == 0
    t.append((time.perf_counter() - t0) * 1e6)
tot = np.zeros(8)
for e in engines:
    lib.mppi_debug_get_host_ns(e.h, buf, 1)
    tot += np.array(buf[:], dtype=np.float64)
names = ["build_params", "stage_costmap", "event0", "graph launch", "event1", "wait result", "copy-out"]
print("robots %d: batch call p50 %.1f us (%.2f us per robot)" % (n, np.percentile(t, 50), np.percentile(t, 50) / n))
print("per robot per call (us): " + ", ".join("%s %.2f" % (nm, tot[i] / tot[7] / 1e3) for i, nm in enumerate(names)))

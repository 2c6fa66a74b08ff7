#!/usr/bin/env python
"""Where the host time of one mppi_optimize() goes (Python marshalling vs the C call vs the device)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402
from mpcholonavigation_b200 import _abi as abi  # noqa: E402


def p50(f, n=400):
    for _ in range(20):
        f()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        t.append((time.perf_counter() - t0) * 1e6)
    return float(np.percentile(t, 50))


fns = load_product()
sc = scenarios.config1()
e = Engine(fns, **sc.cfg)
e.set_robot(sc.robot); e.set_critics(sc.critics); e.set_noise(*sc.noise())
print("optimize(cycle)            p50 %.1f us" % p50(lambda: e.optimize(sc.cycle)))
pk = sc.cycle.packed()
print("optimize(packed cycle)     p50 %.1f us" % p50(lambda: e.optimize(pk)))
cin, keep = sc.cycle.pack()
out, arrs = e._out()
f = fns["optimize"]
h = e.h
print("raw ctypes mppi_optimize   p50 %.1f us" % p50(lambda: f(h, C.byref(cin), C.byref(out))))
e.upload_cycle(sc.cycle)
g = fns["optimize_resident"]
print("raw ctypes resident        p50 %.1f us" % p50(lambda: g(h, C.byref(out))))
dev = []
for _ in range(200):
    g(h, C.byref(out)); dev.append(out.device_ms * 1e3)
print("device (events)            p50 %.1f us" % np.percentile(dev, 50))
cmd = np.zeros(3, np.float32)
ev = fns["eval_control"]
print("raw ctypes mppi_eval_control(shift) p50 %.1f us" % p50(lambda: ev(h, C.byref(cin), 1, C.byref(out), cmd.ctypes.data_as(abi.f32p))))

fns["set_timing"](h, 0)
print("timing off: raw ctypes mppi_optimize   p50 %.1f us" % p50(lambda: f(h, C.byref(cin), C.byref(out))))
print("timing off: raw ctypes resident        p50 %.1f us" % p50(lambda: g(h, C.byref(out))))
print("timing off: raw ctypes mppi_eval_control(shift) p50 %.1f us" % p50(lambda: ev(h, C.byref(cin), 1, C.byref(out), cmd.ctypes.data_as(abi.f32p))))
lib = fns["_lib"]
buf = (C.c_uint64 * 8)()
lib.mppi_debug_get_host_ns.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32]
lib.mppi_debug_get_host_ns(h, buf, 1)
for _ in range(500):
    f(h, C.byref(cin), C.byref(out))
lib.mppi_debug_get_host_ns(h, buf, 1)
n = max(1, buf[7])
names = ["build_params", "stage_costmap", "event0", "graph launch", "event1", "wait result", "copy-out+elapsed"]
print("mppi_optimize host phases (us/call): " + ", ".join("%s %.2f" % (nm, buf[i] / n / 1e3) for i, nm in enumerate(names)))

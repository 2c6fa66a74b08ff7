#!/usr/bin/env python
"""Phase timeline of the tile kernels (block 0, thread 0) from a -DMPPI_TRACE build:  python scripts/trace_phases.py LIB"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402

fns = load_product(sys.argv[1])
sc = scenarios.config1(batch=int(sys.argv[2])) if len(sys.argv) > 2 else scenarios.config1()
e = Engine(fns, **sc.cfg)
e.set_robot(sc.robot); e.set_critics(sc.critics); e.set_noise(*sc.noise())
e.upload_cycle(sc.cycle)
for _ in range(30):
    e.optimize_resident()
lib = fns["_lib"]
buf = (C.c_longlong * 64)()
acc = []
for _ in range(20):
    e.optimize_resident()
    assert lib.mppi_debug_get_trace(buf, 64) == 0
    acc.append(np.array(buf[:], dtype=np.int64))
a = np.median(np.stack(acc), axis=0)
k2 = a[0:16] - a[0]
k3 = a[16:23] - a[16]
print("B =", e.B)
print("K2 / fused phase starts (cycles from kernel start, block 0; 10.. = fused tail: barrier 1 passed, decisions,")
print("   path critics, totals, [weighted sums + barrier 2], merge done):", " ".join("%d" % v for v in k2))
print("K3 phase starts (cycles from kernel start, block 0):", " ".join("%d" % v for v in k3))
print("K3 preamble: issue-done %d, barrier-1 %d, decisions-done %d" % tuple(a[24:27] - a[16]))
print("K2 end -> K3 start gap (cycles, only meaningful if both blocks ran on the same SM):", int(a[16] - a[9]))

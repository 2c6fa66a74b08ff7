#!/usr/bin/env python
"""Times optimize_resident() of one workload for several builds of the library (kernel tuning aid).

  python scripts/time_variants.py --batch 262144 build/v_a.so build/v_b.so ...
Prints per-kernel CUDA-event times (mppi_set_profiling) as medians over the cycles.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sharded_262144x100")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--cycles", type=int, default=30)
    ap.add_argument("--flush", action="store_true", help="cold L2: 256 MiB written between cycles (what bench.py does)")
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    kw = {"batch": a.batch} if a.batch else {}
    sc, philox = {"omni_1000x56": (scenarios.config1, False), "obstacles_16384x56": (scenarios.config3, False),
                  "obstacles_dense_16384x56": (lambda **kw: scenarios.config3(dense=True, **kw), False),
                  "sharded_262144x100": (scenarios.config4, True)}[a.workload]
    sc = sc(**kw)
    noise = None if philox else sc.noise()
    flush = lambda: None
    if a.flush:
        import torch
        buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")

        def flush():
            buf.zero_()
            torch.cuda.synchronize()
    for lib in (a.libs or [None]):
        e = Engine(load_product(lib), **sc.cfg)
        e.set_robot(sc.robot)
        e.set_critics(sc.critics)
        if philox:
            e.generate_noise(0)
        else:
            e.set_noise(*noise)
        e.upload_cycle(sc.cycle)
        for _ in range(5):
            e.optimize_resident()
        tot = []
        for _ in range(a.cycles):
            flush()
            tot.append(e.optimize_resident().device_ms)
        e.set_profiling(True)
        k = []
        for _ in range(a.cycles):
            flush()
            e.optimize_resident()
            p = e.get_profile()
            k.append((p["k2_ms"], p["k3_ms"], p["exchange_ms"]))
        k = np.median(np.asarray(k), axis=0)
        print("%-28s B=%d T=%d total %.1f us | K2 %.1f  K3 %.1f  merge %.1f" % (
            os.path.basename(lib) if lib else "default", e.B, e.T, np.median(tot) * 1e3, k[0] * 1e3, k[1] * 1e3, k[2] * 1e3), flush=True)
        e.close()


if __name__ == "__main__":
    main()

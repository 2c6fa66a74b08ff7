#!/usr/bin/env python
"""Evidence for the scan decision (round-1 verdict, task 9): north_star item (2) asks for warp-level prefix scans along the
horizon; the parity path evaluates the three cumsums (yaw, x, y; optimizer.cpp:319-342) sequentially in t because that is
the reference's summation order.  This script runs the warp-shuffle scan variant of the tile kernels (MPPI_SCAN=warp, a
runtime switch of the same library) and the sequential default on BASELINE configs[1] (1000 x 56, fused kernel) and on
configs[3] (262144 x 100, tile layout) against the CPU oracle and records
  (i)   the fraction of (b, t) whose costmap cell index differs from the oracle's,
  (ii)  the largest deviation of the control sequence from the oracle's,
  (iii) the device time of the cycle with either variant.
Writes ONE JSON object to stdout (committed as profiles/r02_scan_experiment.json).  Needs a GPU."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(config, scan):
    from mpcholonavigation_b200 import Engine, load_product, scenarios
    from tests import oracle_loader
    pf, of = load_product(), oracle_loader.load()
    if config == "omni_1000x56":
        sc = scenarios.config1()
        noise = sc.noise()
        g = Engine(pf, **sc.cfg)
        g.set_noise(*noise)
    else:
        sc = scenarios.config4()
        g = Engine(pf, **{**sc.cfg, "seed": 3})
        g.generate_noise(0)
        noise = g.get_noise()
    o = Engine(of, **sc.cfg)
    of["set_wide_reductions"](o.h, 1)
    o.set_noise(*noise)
    for e in (g, o):
        e.set_robot(sc.robot); e.set_critics(sc.critics)
        e.set_outputs(trajectories=True, cells=True)
    cycles = 6 if config == "omni_1000x56" else 2
    cell_diff, pose_diff, ctrl_dev, n = 0, 0, 0.0, 0
    for _ in range(cycles):
        rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
        cg, co = g.get_cells(), o.get_cells()
        cell_diff += int(np.count_nonzero(cg != co)); n += cg.size
        for a, b in zip(g.get_trajectories(), o.get_trajectories()):
            pose_diff += int(np.count_nonzero(a != b))
        for a, b in ((rg.vx, ro.vx), (rg.vy, ro.vy), (rg.wz, ro.wz)):
            ctrl_dev = max(ctrl_dev, float(np.max(np.abs(a - b) / (1e-6 + 1e-4 * np.abs(b)))))
        g.set_control_sequence(ro.vx, ro.vy, ro.wz)      # same warm start for the next cycle
    # timing: nothing materialised, device-resident inputs, warm
    g.set_outputs()
    g.upload_cycle(sc.cycle)
    ms = [g.optimize_resident().device_ms for _ in range(60)][10:]
    print(json.dumps({"config": config, "scan": scan, "cycles": cycles, "cells_compared": n, "cells_different": cell_diff,
                      "cell_mismatch_fraction": cell_diff / n, "pose_values_different": pose_diff,
                      "control_violation_ratio_of_1e-4_bar": ctrl_dev, "device_ms_p50": float(np.median(ms))}))


def main():
    if len(sys.argv) > 1:
        return child(sys.argv[1], sys.argv[2])
    out = {"what": __doc__.split("Writes")[0].strip(), "runs": []}
    for config in ("omni_1000x56", "omni_262144x100"):
        for scan in ("sequential", "warp"):
            env = dict(os.environ, MPPI_SCAN=scan)
            if config == "omni_262144x100":
                env["MPPI_STREAM_MIN_BATCH"] = "1000000000"     # the tile kernels (the stream kernel holds a trajectory in one thread)
            r = subprocess.run([sys.executable, __file__, config, scan], env=env, capture_output=True, text=True, timeout=1200)
            if r.returncode != 0:
                out["runs"].append({"config": config, "scan": scan, "error": r.stderr[-500:]})
            else:
                out["runs"].append(json.loads(r.stdout.strip().splitlines()[-1]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

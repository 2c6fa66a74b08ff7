#!/usr/bin/env python
"""How far the CUDA path sits from the CPU oracle, measured: per-critic costs, total costs and controls for BASELINE
configs[1] (tile layout, fused kernel), configs[2] and configs[3] at 65536 x 100 (stream layout), three warm cycles each.
Prints max |d| and max |d| / |ref| (over |ref| > 1e-3) per quantity; feeds profiles/r02b_parity_margins.txt, which is
what the absolute floors in tests/test_gpu_parity*.py are justified by.  Run on a GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402
from tests import oracle_loader  # noqa: E402


def dev(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(a - b)
    big = np.abs(b) > 1e-3
    return float(d.max()), float((d[big] / np.abs(b[big])).max()) if big.any() else 0.0, float(np.abs(b).max())


def main():
    cases = [("configs[1] omni_1000x56 (fused tile kernel)", scenarios.config1(), False),
             ("configs[2] obstacles_16384x56 (stream layout)", scenarios.config3(), False),
             ("configs[3] omni_65536x100 (stream layout, Philox noise)", scenarios.config4(batch=65536), True)]
    for label, sc, philox in cases:
        g = Engine(load_product(), **dict(sc.cfg, seed=3))
        o = Engine(oracle_loader.load(), **sc.cfg)
        for e in (g, o):
            e.set_robot(sc.robot); e.set_critics(sc.critics)
            e.set_outputs(trajectories=True, cells=True, critic_costs=True)
        if philox:
            g.generate_noise(0)
            noise = g.get_noise()
        else:
            noise = sc.noise()
            g.set_noise(*noise)
        o.set_noise(*noise)
        worst = {}
        for cycle in range(4):
            rg, ro = g.optimize(sc.cycle), o.optimize(sc.cycle)
            assert np.array_equal(g.get_cells(), o.get_cells())
            rows = [("total costs", g.get_costs(), o.get_costs())]
            rows += [("critic %d %s" % (q, sc.critics[q][0]), g.get_critic_costs(q), o.get_critic_costs(q)) for q in range(len(sc.critics))]
            rows += [("control " + n, getattr(rg, n), getattr(ro, n)) for n in ("vx", "vy", "wz")]
            for name, a, b in rows:
                d = dev(a, b)
                w = worst.get(name, (0.0, 0.0, 0.0))
                worst[name] = (max(w[0], d[0]), max(w[1], d[1]), max(w[2], d[2]))
            g.set_control_sequence(ro.vx, ro.vy, ro.wz)
        print(label + ": cell indices bit-equal in 4 cycles; worst over the cycles")
        for name, (dabs, drel, mag) in worst.items():
            print("  %-32s max|d| %.3e   max|d|/|ref| %.3e   (max |ref| %.4g)" % (name, dabs, drel, mag))
        g.close(); o.close()


if __name__ == "__main__":
    main()

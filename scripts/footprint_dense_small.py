#!/usr/bin/env python
"""The fused small-batch kernel with the robot boxed in (scenarios.config1(footprint="rectangle", ring=0.47): CostCritic's
footprint branch taken by ~14 % of the visited poses) against the same scene without the ring: device time per optimize(),
warm and with the L2 flushed between cycles (what bench.py times)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402


def main():
    import torch
    buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    fns = load_product()
    for name, sc in (("configs[1]", scenarios.config1()),
                     ("configs[1], rectangular footprint, no ring", scenarios.config1(footprint="rectangle")),
                     ("configs[1], boxed in", scenarios.config1(footprint="rectangle", ring=0.47))):
        e = Engine(fns, **sc.cfg)
        e.set_robot(sc.robot); e.set_critics(sc.critics); e.set_noise(*sc.noise())
        e.upload_cycle(sc.cycle)
        for _ in range(20):
            e.optimize_resident()
        warm = [e.optimize_resident().device_ms for _ in range(50)]
        cold = []
        for _ in range(50):
            buf.zero_(); torch.cuda.synchronize()
            cold.append(e.optimize_resident().device_ms)
        print("%-46s device per optimize(): warm %.1f us, cold L2 %.1f us" % (name, np.median(warm) * 1e3, np.median(cold) * 1e3), flush=True)
        e.close()


if __name__ == "__main__":
    main()

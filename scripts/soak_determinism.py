#!/usr/bin/env python
"""Soak test of the stream layout's in-kernel merge: the same cycle N times from the same control sequence, every result
compared bit for bit with the first (a race in the last-block merge would show as a result that changes).

  python scripts/soak_determinism.py [--batch 65536] [--cycles 20000]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpcholonavigation_b200 import Engine, load_product, scenarios  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--cycles", type=int, default=20000)
    a = ap.parse_args()
    sc = scenarios.config4(batch=a.batch)
    e = Engine(load_product(), **dict(sc.cfg, seed=3))
    e.set_robot(sc.robot); e.set_critics(sc.critics); e.generate_noise(0)
    zero = np.zeros(sc.cfg["time_steps"], np.float32)
    first, bad = None, 0
    for c in range(a.cycles):
        e.set_control_sequence(zero, zero, zero)
        r = e.optimize(sc.cycle)
        got = np.concatenate([r.vx, r.vy, r.wz])
        if first is None:
            first = got.copy()
        elif not np.array_equal(got, first):
            bad += 1
            if bad <= 5:
                i = int(np.argmax(np.abs(got - first)))
                print("cycle %d: column %d %.9g instead of %.9g" % (c, i, got[i], first[i]))
    print("batch %d: %d of %d cycles differ from the first" % (a.batch, bad, a.cycles))
    e.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

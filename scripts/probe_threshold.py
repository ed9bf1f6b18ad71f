#!/usr/bin/env python
"""Where should DeviceHistogram switch from the generic multi-lane kernels to the one-thread-per-point kernel?
Times both at several sweep sizes on the config-2 histogram."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine, synth  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    lnpi = synth.two_peak_lnpi(1001)
    dh = engine.DeviceHistogram(lnpi, np.arange(1001), 1.0, 0.0, smooth=10, sel=[np.arange(1001.0), np.arange(1001.0) ** 2])
    dh.ensure_hull()
    for S in (300, 1000, 3000, 10000, 30000, 100000, 300000):
        mu = dh._dev_array(np.linspace(-0.03, 0.03, S))
        row = {"S": S}
        for name, lanes in (("fast_thread", 1), ("generic_1", -1), ("generic_4", 4), ("generic_32", 32)):
            row[name + "_ms"] = timed(lambda: dh.sweep(mu, pmax=4, lanes=lanes))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

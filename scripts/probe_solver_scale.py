#!/usr/bin/env python
"""K4 at 10^4 .. 3x10^5 solves: warp-per-solve vs thread-per-solve kernels (config-4 histogram, cold guesses)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402


def timed(fn, reps=3, warm=1):
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


h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
for T in (10000, 30000, 100000, 300000):
    betas = 1.0 / np.linspace(0.90, 1.06, T)
    dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
    g = np.zeros_like(betas)
    row = {"solves": T}
    for name, lanes in (("warp_per_solve", "32"), ("thread_per_solve", "1000")):
        os.environ["FHMC_SOLVER_LANES"] = lanes
        r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4)
        hr = r.host()
        row[name + "_ms"] = timed(lambda: dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4))
        row[name + "_ok"] = float((hr["code"] == 0).mean())
        row[name + "_max_evals"] = int(hr["iters"].max())
        row[name + "_mean_evals"] = float(hr["iters"].mean())
    os.environ.pop("FHMC_SOLVER_LANES")
    print(json.dumps(row), flush=True)

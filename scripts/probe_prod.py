#!/usr/bin/env python
"""Quick look at the pure-mu fast sweep variants on the config-2 histogram: time per 10^6 state points for
FHMC_MU_RECURRENCE = 2 (product form), 1 (chains), 0 (true exps), and max deviation from the generic kernel."""
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
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=float)
    variants = [int(v) for v in (sys.argv[1:] or ["3", "2", "1", "0"])]
    for rec in variants:
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
        dh.use_recurrence = rec
        dh.ensure_hull()
        mu = dh._dev_array(np.linspace(-0.03, 0.03, 1000000))
        out = dh.sweep(mu, pmax=4)
        ms = timed(lambda: dh.sweep(mu, pmax=4, out=out))
        mus = np.linspace(-0.03, 0.03, 20000)
        a = dh.sweep(mus, pmax=4, lanes=1).host()
        b = dh.sweep(mus, pmax=4, lanes=-1).host()
        mask = np.arange(4)[None, :] < a["nphase"][:, None]
        row = {"rec": rec, "ms_per_1e6": ms, "pts_per_s": 1e9 / ms,
               "int_equal": bool(all(np.array_equal(a[k], b[k]) for k in ("nphase", "nmin")) and
                                 np.array_equal(a["bounds"][mask], b["bounds"][mask])),
               "fe_maxrel": float(np.max(np.abs(a["fe"][mask] / b["fe"][mask] - 1))),
               "avg_maxrel": float(np.max(np.abs(a["avg"][mask] / b["avg"][mask] - 1))),
               "fast_frac": float(np.mean((a["status"] & 0x1000) != 0))}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

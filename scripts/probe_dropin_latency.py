#!/usr/bin/env python
"""Latency of the scalar drop-in API (fresh histogram -> reweight -> thermo(all moments) -> is_safe; find_phase_eq), with a
cProfile breakdown (PROFILE=1)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

lnpi1, mom1 = synth.two_peak_lnpi(1001), synth.one_comp_moments(1001)
h1 = histogram.from_arrays(lnpi1, mom1, 1.0, [0.0], 10)


def dropin():
    hh = histogram.from_arrays(lnpi1, mom1, 1.0, [0.0], 10)
    hh.reweight(0.01)
    hh.thermo()
    return hh.is_safe()


def loop_same():
    # the notebook loop: one histogram object, many state points (deepcopy per point like the reference's examples)
    import copy
    hh = copy.deepcopy(h1)
    hh.reweight(0.01)
    hh.thermo()
    return hh.is_safe()


for fn, name in ((dropin, "from_arrays+reweight+thermo+is_safe"), (loop_same, "deepcopy+reweight+thermo+is_safe")):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
    print("%s: %.3f ms per state point" % (name, (time.perf_counter() - t0) / 50 * 1e3))
t0 = time.perf_counter()
for _ in range(10):
    h1.find_phase_eq(1e-6, 0.0)
print("find_phase_eq: %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
if os.environ.get("PROFILE") == "2":
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        h1.find_phase_eq(1e-6, 0.0)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
elif os.environ.get("PROFILE"):
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        dropin()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
if os.environ.get("PROFILE") == "3":
    # cost of the pieces of one scalar call
    from fhmcanalysis_b200 import engine
    sp = engine.ScalarPath.get(1001)
    N = np.arange(1001)
    m2 = np.asarray(mom1, dtype=np.float64).reshape(-1, 1001)

    def t(fn, reps=300):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e6
    print("point(complete, row)          %.1f us" % t(lambda: sp.point(lnpi1, N, 1.0, 0.0, 10, 0.01, complete=True, want_row=True)))
    print("point(split only)             %.1f us" % t(lambda: sp.point(lnpi1, N, 1.0, 0.0, 10, 0.0)))
    print("point(split, row, 27 moments) %.1f us" % t(lambda: sp.point(lnpi1, N, 1.0, 0.0, 10, 0.0, want_row=True, mom=m2)))
    print("key(lnpi) %.1f us  key(N) %.1f us  key(mom) %.1f us" % (t(lambda: sp._key(lnpi1)), t(lambda: sp._key(np.ascontiguousarray(N, dtype=np.float64))), t(lambda: sp._mom_key(m2))))
    hh = histogram.from_arrays(lnpi1, mom1, 1.0, [0.0], 10)
    print("from_arrays %.1f us" % t(lambda: histogram.from_arrays(lnpi1, mom1, 1.0, [0.0], 10)))
    print("reweight %.1f us  thermo %.1f us  is_safe %.1f us" % (t(lambda: hh.reweight(0.01)), t(lambda: hh.thermo()), t(lambda: hh.is_safe())))

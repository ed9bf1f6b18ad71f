#!/usr/bin/env python
"""Which config-4 evaluations (as the coexistence solver issues them) leave the one-pass walk, and why?"""
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
T = 20000
betas = 1.0 / np.linspace(0.90, 1.06, T)
dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
for mu0 in (0.0, 0.05, -0.05, 0.15, -0.15, 0.4, -0.4):
    r = dh.sweep(np.full(T, mu0), beta=betas, pmax=4, lanes=1).host()
    st = r["status"]
    fast = (st & 0x1000) != 0
    c = collections.Counter()
    for s_, P, nm in zip(st[~fast], r["nphase"][~fast], r["nmin"][~fast]):
        c[(int(s_ & 0xFF), bool(s_ & 0x400), bool(s_ & 0x800), bool(s_ & 0x200), int(P), int(nm))] += 1
    print("mu %+.2f fast %.4f rescued-in-fast %.4f not-fast:" % (mu0, fast.mean(), ((st & 0x800) != 0)[fast].mean()),
          ["code %d slow %d resc %d gap %d P %d nm %d: %d" % (k + (v,)) for k, v in c.most_common(4)])

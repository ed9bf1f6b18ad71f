#!/usr/bin/env python
"""Which config-3 state points leave the thread-per-point kernel for the generic evaluator, and why?"""
import os
import sys
import collections

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhmcanalysis_b200 import engine, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

n = 1001
h = histogram.from_arrays(synth.two_peak_lnpi(n), synth.two_comp_moments(n), 1.0, [-3.0, -2.5], 10)
h.reweight(-2.9)
nb = 256
betas, dmus = np.linspace(0.95, 1.05, nb), np.linspace(0.2, 0.8, nb)
dh = h.device_histogram(beta=betas, dmu=dmus, order=2, moments=())
st = dh.make_states(np.array([-2.9]), betas, dmus, grid=True)
res = dh.sweep(None, states=st, pmax=8, lanes=1)
hr = res.host()
stt = hr["status"]
fast = (stt & 0x1000) != 0
print("fast fraction", fast.mean())
nf = ~fast
c = collections.Counter()
for s_, P, nm in zip(stt[nf], hr["nphase"][nf], hr["nmin"][nf]):
    c[(int(s_ & 0xFF), bool(s_ & 0x400), bool(s_ & 0x800), bool(s_ & 0x200), int(P), int(nm))] += 1
for k, v in c.most_common(12):
    print("code %d slow %s rescued %s gapfill %s P %d nmin %d : %d" % (k + (v,)))
idx = np.nonzero(nf)[0]
print("grid positions (beta idx, dmu idx) of some:", [(int(i // nb) if st.beta_div > st.dmu_div else int(i % nb), int(i)) for i in idx[:10]])
# distribution over the grid
g = nf.reshape(nb, nb)
print("per-row fallback counts (first axis) min/max:", g.sum(1).min(), g.sum(1).max(), "per-col:", g.sum(0).min(), g.sum(0).max())
k = idx[len(idx) // 2]
print("example", k, "maxima", hr["max_idx"][k, :hr["nphase"][k]], "minima", hr["min_idx"][k, :hr["nmin"][k]], "fe", hr["fe"][k, :hr["nphase"][k]])

"""One launch of the K4 solver on config 4 (for ncu)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
T = 10000
h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
betas = 1.0 / np.linspace(0.90, 1.06, T)
dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
r = dh.find_phase_eq(np.zeros_like(betas), beta=betas, lnz_tol=1e-10, pmax=4)
torch.cuda.synchronize()
print((r.host()["code"] == 0).mean())

"""Where a find_phase_eq call spends its time: kernel (CUDA events right around the C call) vs host path."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import _lib, synth, engine
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
T = 10000
h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
betas = 1.0 / np.linspace(0.90, 1.06, T)
dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
g = np.zeros_like(betas)
L = _lib.load()
orig = L.fhmc_find_phase_eq_1d
times = []
def wrapped(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = orig(*a)
    e1.record()
    times.append((e0, e1))
    return rc
L.fhmc_find_phase_eq_1d = wrapped
for cont in (False, True):
    for rep in range(4):
        times.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4, continuation=cont)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ks = [a.elapsed_time(b) for a, b in times]
    print("continuation=%s: host call %.3f ms, until results ready %.3f ms, kernel(s) %s ms, evals/solve last stage %.2f" %
          (cont, 1e3 * (t1 - t0), 1e3 * (t2 - t0), ["%.3f" % k for k in ks], r.host()["iters"].mean()))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhmcanalysis_b200 import synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
for T in (10000, 100000):
    betas4 = 1.0 / np.linspace(0.90, 1.06, T)
    dh4 = h4.device_histogram(beta=betas4, order=2, moments=("N", "N2", "U"))
    g4 = np.zeros_like(betas4)
    for rep in range(3):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); r4 = dh4.find_phase_eq(g4, beta=betas4, lnz_tol=1e-10, pmax=4); c1.record(); c1.synchronize()
    hr = r4.host()
    print(os.environ.get("FHMC_SOLVER_LANES"), "T", T, "ms", c0.elapsed_time(c1), "pts/s %.3e" % (T / c0.elapsed_time(c1) * 1e3), "ok", float(np.mean(hr["code"] == 0)), "evals", float(np.mean(hr["iters"])))

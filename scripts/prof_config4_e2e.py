import cProfile, pstats, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from fhmcanalysis_b200 import synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
n4 = 2001
h4 = histogram.from_arrays(synth.two_peak_lnpi(n4, scale=2.0), synth.one_comp_moments(n4, max_order=3), 1.0, [0.0], 10)
betas = 1.0 / np.linspace(0.90, 1.06, 10000)
for _ in range(3):
    h4.find_phase_eq_batch(betas, 0.0, order=2, lnZ_tol=1e-10)
t0 = time.perf_counter()
for _ in range(5):
    h4.find_phase_eq_batch(betas, 0.0, order=2, lnZ_tol=1e-10)
print("e2e ms", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    h4.find_phase_eq_batch(betas, 0.0, order=2, lnZ_tol=1e-10)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

"""Cycle shares of the sections of LeanEval::run inside the K4 solver (library built with FHMC_NVCC_FLAGS=-DFHMC_LEAN_PROFILE)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import _lib, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
names = ["passA", "candlist", "win2..D1", "win rest", "repair", "passB", "phase red.", "retest+safe"]
for T in (1000, 10000):
    betas = 1.0 / np.linspace(0.90, 1.04, T) if T > 1 else np.array([1.0 / 0.95])
    dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
    g = np.zeros_like(betas)
    r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4)
    torch.cuda.synchronize()
    _lib.lean_stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4)
    e1.record()
    e1.synchronize()
    st = _lib.lean_stats()
    ne = st[0]
    tot = st[7]
    print("T=%d: %.3f ms, %d lean evals (%.2f per solve), solve cycles per eval %.0f" % (T, e0.elapsed_time(e1), ne, ne / T, tot / max(ne, 1)))
    print("   cycles per eval by section:", {n: int(c / max(ne, 1)) for n, c in zip(names, st[8:16])}, " sum", int(sum(st[8:16]) / max(ne, 1)),
          " between evals", int(st[16] / max(ne, 1)), " run() total", int(st[17] / max(ne, 1)))

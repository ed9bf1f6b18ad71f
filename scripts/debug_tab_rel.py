"""Worst pure-relative differences between the table-driven compact sweep and the general kernel on config 2."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import _lib, engine, synth
n = 1001
lnpi = synth.two_peak_lnpi(n)
N = np.arange(n, dtype=np.float64)
S = 1000000
mu = np.linspace(-0.03, 0.03, S)
for tables in (True, False):
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    dh.use_mu_tables = tables
    c = dh.sweep_compact(mu, pmax=4)
    print("tables", tables, _lib.last_kernel())
    g = dh.sweep(mu, pmax=4, lanes=-1).host()
    fe, av = c["fe"].cpu().numpy(), c["avg"].cpu().numpy()
    P = g["nphase"]
    for name, a, b in (("fe", fe, g["fe"]), ("avgN", av[..., 0], g["avg"][..., 0]), ("avgN2", av[..., 1], g["avg"][..., 1])):
        for p in range(2):
            live = P > p
            rel = np.zeros(S)
            rel[live] = np.abs(a[live, p] - b[live, p]) / np.abs(b[live, p])
            k = int(np.argmax(rel))
            print("  %-6s phase %d  max rel %.3e at k=%d mu=%.6f  got %.15g want %.15g  bounds %s  median rel %.2e" % (
                name, p, rel[k], k, mu[k], a[k, p], b[k, p], g["bounds"][k, :P[k]].tolist(), np.median(rel[live])))

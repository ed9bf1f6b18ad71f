"""Small run touching every kernel once (for compute-sanitizer memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhmcanalysis_b200 import engine, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
n = 257
lnpi, mom = synth.two_peak_lnpi(n, scale=0.25), synth.one_comp_moments(n, 3)
N = np.arange(n, dtype=float)
dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=4, sel=["N", N * N, mom[0, 0, 0, 0, 1]])
dh.ensure_hull()
mus = np.linspace(-0.2, 0.2, 700)
for lanes in (1, -1, 4, 32):
    r = dh.sweep(mus, pmax=4, lanes=lanes)
    rows = dh.lnpi_rows(r)
r = dh.sweep(mus, pmax=1, complete=True)
h = histogram.from_arrays(lnpi, mom, 1.0, [0.0], 4)
h.reweight(0.01); h.thermo(); h.is_safe()
e = h.find_phase_eq(1e-8, 0.0, 1.01, [], 2)
o = h.find_phase_eq_batch(np.linspace(0.98, 1.02, 37), 0.0, order=2)
o2 = h.reweight_batch(np.linspace(-0.1, 0.1, 300), beta=np.array([1.01]), order=2)
l2, b2 = synth.joint_2d(67, 45, 70)
x = engine.reweight_2d(l2, b2, np.arange(67.0), np.arange(45.0), np.linspace(-0.1, 0.1, 333), np.linspace(0.1, -0.1, 333), np.random.default_rng(0).random((2, 67, 45)))
h2 = h.mix(h, [0.4, 0.6])
# r02: compact records on per-histogram tables, the row-combined Taylor grid kernel, the curve solver, the scalar path
dhc = histogram.from_arrays(lnpi, mom, 1.0, [0.0], 4).device_histogram(moments=("N", "N2"))
c = dhc.sweep_compact(np.linspace(-0.2, 0.2, 5000), pmax=4)
n2 = 301
h2c = histogram.from_arrays(synth.two_peak_lnpi(n2), synth.two_comp_moments(n2), 1.0, [-3.0, -2.5], 5)
h2c.reweight(-2.9)
bs, ds = np.linspace(0.97, 1.03, 3), np.linspace(0.2, 0.8, 600)
dg = h2c.device_histogram(beta=bs, dmu=ds, order=2, moments=())
g = dg.sweep(None, states=dg.make_states(np.array([-2.9]), bs, ds, grid=True), pmax=8, lanes=1)
n4 = 401
h4 = histogram.from_arrays(synth.two_peak_lnpi(n4, scale=0.4), synth.one_comp_moments(n4, 3), 1.0, [0.0], 5)
cv = h4.find_phase_eq_batch(1.0 / np.linspace(0.95, 1.03, 300), 0.0, order=2)
torch.cuda.synchronize()
print("sanitize case done", int(o["code"].sum()), float(x[0, 0]), int((cv["code"] == 0).sum()))

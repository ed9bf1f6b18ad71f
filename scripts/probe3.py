import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhmcanalysis_b200 import engine, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
n = 1001
lnpi, mom2 = synth.two_peak_lnpi(n), synth.two_comp_moments(n)
h = histogram.from_arrays(lnpi, mom2, 1.0, [-3.0, -2.5], 10)
h.reweight(-2.9)
betas, dmus = np.linspace(0.95, 1.05, 1024), np.linspace(0.2, 0.8, 1024)
dh = h.device_histogram(beta=betas, dmu=dmus, order=2, moments=())
st = dh.make_states(np.array([-2.9]), betas, dmus, grid=True)
res = engine.SweepResult(st.n_states, 8, dh.n_sel, dh.device)
for rep in range(3):
    dh.sweep(None, states=st, out=res, pmax=8, lanes=1)
torch.cuda.synchronize()
print("done")

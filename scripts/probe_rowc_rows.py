import os, sys
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from fhmcanalysis_b200 import _lib, engine, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
n = 1001
h = histogram.from_arrays(synth.two_peak_lnpi(n), synth.two_comp_moments(n), 1.0, [-3.0, -2.5], 10)
h.reweight(-2.9)
betas, dmus = np.linspace(0.95, 1.05, 4096), np.linspace(0.2, 0.8, 4096)
dh = h.device_histogram(beta=betas, dmu=dmus, order=2, moments=())
st = dh.make_states(np.array([-2.9]), betas, dmus, grid=True)
res = engine.SweepResult(st.n_states, 8, dh.n_sel, dh.device)
dh.sweep(None, states=st, out=res, pmax=8)
torch.cuda.synchronize()
slow = ((res.status & 0x1000) == 0).view(4096, 4096)
per_row = slow.sum(1).cpu().numpy()
print("slow total", per_row.sum(), "rows with slow", (per_row > 0).sum(), "max per row", per_row.max(), "top rows", np.argsort(-per_row)[:10], np.sort(-per_row)[:10])
per_col = slow.sum(0).cpu().numpy()
print("cols with slow", (per_col > 0).sum(), "max per col", per_col.max())
bad = ((res.status & 0xFF) != 0).view(4096, 4096).sum(1).cpu().numpy()
print("bad codes", bad.sum(), "rows", (bad > 0).sum(), "max", bad.max())
mono = (res.nphase.view(4096, 4096) == 1).sum(1).cpu().numpy()
print("one-phase points", mono.sum(), "rows", (mono > 0).sum())

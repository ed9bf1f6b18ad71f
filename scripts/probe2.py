import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhmcanalysis_b200 import engine, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
n = 1001
lnpi, mom2 = synth.two_peak_lnpi(n), synth.two_comp_moments(n)
h = histogram.from_arrays(lnpi, mom2, 1.0, [-3.0, -2.5], 10)
h.reweight(-2.9)
betas, dmus = np.linspace(0.95, 1.05, 1024), np.linspace(0.2, 0.8, 1024)
for order, moments in ((2, ()), (1, ("N1", "N2", "U"))):
    dh = h.device_histogram(beta=betas, dmu=dmus, order=order, moments=moments)
    st = dh.make_states(np.array([-2.9]), betas, dmus, grid=True)
    S = st.n_states
    res = engine.SweepResult(S, 8, dh.n_sel, dh.device)
    for lanes in (1, -1):
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dh.sweep(None, states=st, out=res, pmax=8, lanes=lanes); e1.record(); e1.synchronize()
        stt = res.status.cpu().numpy().view(np.uint32)
        print("order", order, "nsel", dh.n_sel, "ncoef", dh.desc.n_coef, "nterm", dh.desc.n_term, "lanes", lanes, "ms", e0.elapsed_time(e1),
              "pts/s %.3e" % (S / e0.elapsed_time(e1) * 1e3), "fast frac", float(np.mean((stt & 0x1000) != 0)), "ok", float(np.mean((stt & 0xff) == 0)),
              "slow", float(np.mean((stt & 0x400) != 0)), "nphase", np.bincount(res.nphase.cpu().numpy())[:9])
    flags = {name: float(np.mean((stt & bit) != 0)) for name, bit in (("gap", 0x200), ("slow", 0x400), ("resc", 0x800), ("fast", 0x1000))}
    sl = (stt & 0x400) != 0
    print(flags, "slow&gap", float(np.mean(sl & ((stt & 0x200) != 0))), "nphase of slow", np.bincount(res.nphase.cpu().numpy()[sl])[:9],
          "codes of slow", np.unique(stt[sl] & 0xff, return_counts=True))

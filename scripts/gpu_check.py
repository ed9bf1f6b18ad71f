"""Ad-hoc GPU sanity run (not a test): parity vs the C oracle + first timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhmcanalysis_b200 import engine, synth
from oracle import fhmc_oracle as fo

n = 1001
lnpi = synth.two_peak_lnpi(n)
i = np.arange(n)
mom = synth.one_comp_moments(n)
U = mom[0, 0, 0, 0, 1]
dh = engine.DeviceHistogram(lnpi, i, 1.0, 0.0, smooth=10, sel=["N", i.astype(float) ** 2, U])
print("peaks", engine.measure_peaks())
mus = np.linspace(-0.03, 0.03, 2000)
for G in (1, 4, 32):
    res = dh.sweep(mus, pmax=4, lanes=G); torch.cuda.synchronize()
    h = res.host()
    bad = 0; maxrel = 0.0
    for k in range(0, len(mus), 37):
        r = fo.state_point(lnpi, i, 1.0, 0.0, mus[k], 10, sel=np.stack([i.astype(float), i.astype(float) ** 2, U]))
        P = r["nphase"]
        ok = (h["code"][k] == r["status"] and h["nphase"][k] == P and np.array_equal(h["max_idx"][k, :P], r["max_idx"])
              and np.array_equal(h["min_idx"][k, :h["nmin"][k]], r["min_idx"]) and np.array_equal(h["bounds"][k, :P], r["bounds"])
              and bool(h["safe"][k]) == r["safe"])
        if not ok:
            bad += 1
            if bad < 4: print("MISMATCH", G, k, h["code"][k], h["nphase"][k], h["max_idx"][k], h["min_idx"][k], r["max_idx"], r["min_idx"])
        else:
            maxrel = max(maxrel, np.max(np.abs(h["fe"][k, :P] - r["fe"]) / np.abs(r["fe"])), np.max(np.abs(h["avg"][k, :P] - r["avg"]) / np.abs(r["avg"])))
    print("G", G, "bad", bad, "max rel err fe/avg", maxrel, "slow", int(((h["status"] & 0x400) != 0).sum()))
# timing
for S, G in ((1000000, 1), (1000000, 4), (100000, 4), (10000, 32)):
    mu_t = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
    st = dh.make_states(mu_t)
    out = engine.SweepResult(S, 4, dh.n_sel, dh.device)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dh.sweep(None, states=st, out=out, pmax=4, lanes=G); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
    print("S", S, "G", G, "ms", ms, "pts/s %.3e" % (S / ms * 1e3), "exp/s %.3e" % (S * n / ms * 1e3))

import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhmcanalysis_b200 import engine, synth
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
which = sys.argv[1:]
if "agree" in which:
    for n, smooth, noise in ((1001, 10, 1e-3), (573, 3, 5e-2), (2001, 30, 0.0)):
        lnpi = synth.two_peak_lnpi(n, noise=noise, scale=n / 1001.0)
        N = np.arange(n, dtype=float)
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=smooth, sel=["N", N * N])
        mus = np.linspace(-0.05, 0.05, 5000)
        a = dh.sweep_auto(mus, pmax=4, lanes=1).host()
        b = dh.sweep_auto(mus, pmax=a["fe"].shape[1], lanes=-1).host()
        mask = np.arange(a["fe"].shape[1])[None, :] < a["nphase"][:, None]
        for k in ("fe", "avg"):
            x, y = a[k][mask], b[k][mask]
            rel = np.abs(x - y) / np.maximum(np.abs(y), 1e-300)
            i = np.argmax(rel.reshape(len(x), -1).max(axis=1))
            print(n, k, "max rel", rel.max(), "at", x[i], y[i], "codes", np.unique(a["code"]), "pmax", a["fe"].shape[1])
if "c3" in which:
    n = 1001
    lnpi, mom2 = synth.two_peak_lnpi(n), synth.two_comp_moments(n)
    h = histogram.from_arrays(lnpi, mom2, 1.0, [-3.0, -2.5], 10)
    h.reweight(-2.9)
    betas, dmus = np.linspace(0.95, 1.05, 65), np.linspace(0.2, 0.8, 65)
    r = h.reweight_batch(np.array([-2.9]), betas, dmus, order=2, moments=(), grid=True, pmax=4)
    code = r["code"].reshape(65, 65)
    print("c3 codes", {int(c): int((code == c).sum()) for c in np.unique(code)})
    print("nphase", np.unique(r["nphase"], return_counts=True))
    bad = np.argwhere(code != 0)
    print("bad (beta,dmu) sample", [(betas[i], dmus[j], int(code[i, j])) for i, j in bad[:: max(len(bad) // 8, 1)]])
if "c4" in which:
    n4 = 2001
    lnpi4 = synth.two_peak_lnpi(n4, scale=2.0)
    mom4 = synth.one_comp_moments(n4, max_order=3)
    h4 = histogram.from_arrays(lnpi4, mom4, 1.0, [0.0], 10)
    T = np.linspace(0.90, 1.10, 41)
    r = h4.find_phase_eq_batch(1.0 / T, 0.0, order=2, lnZ_tol=1e-10)
    for t, c, mu, it, d in zip(T, r["code"], r["mu_coex"], r["iters"], r["dfe"]):
        print("T %.3f code %d mu %.6f iters %d dfe %.2e nph %s" % (t, c, mu, it, d, ""))
if "e2e" in which:
    n = 1001
    lnpi = synth.two_peak_lnpi(n)
    N = np.arange(n, dtype=float)
    dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
    S = 1000000
    mu_h = torch.from_numpy(np.linspace(-0.03, 0.03, S)).pin_memory()
    for chunk in (1 << 17, 1 << 18, 1 << 19, S):
        out = dh.sweep_host(mu_h, pmax=4, chunk=chunk)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dh.sweep_host(mu_h, pmax=4, chunk=chunk, out=out)
        print("chunk", chunk, "ms", (time.perf_counter() - t0) / 5 * 1e3)
    # raw D2H bandwidth
    x = torch.empty(184000000, dtype=torch.uint8, device="cuda"); y = torch.empty(184000000, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): y.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); print("D2H 184MB ms", (time.perf_counter() - t0) / 5 * 1e3)
    ys = [torch.empty(23000000, dtype=torch.uint8).pin_memory() for _ in range(8)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        for k in range(8): ys[k].copy_(x[k * 23000000:(k + 1) * 23000000], non_blocking=True)
    torch.cuda.synchronize(); print("D2H 8x23MB ms", (time.perf_counter() - t0) / 5 * 1e3)

"""K4: lean warp-per-solve kernel against the PointEval group kernel on config 4 (10^4 temperatures, N_max = 2000, order 2)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import _lib, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], 10)
betas = 1.0 / np.linspace(0.90, 1.06, T)
dh = h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"))
g = np.zeros_like(betas)


def run(env):
    for k in ("FHMC_SOLVER_LANES", "FHMC_SOLVER_CTA"):
        os.environ.pop(k, None)
    os.environ.update(env)
    cont = env.pop("CONT", "0") == "1"
    r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4, continuation=cont)
    torch.cuda.synchronize()
    name = _lib.last_kernel()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = dh.find_phase_eq(g, beta=betas, lnz_tol=1e-10, pmax=4, continuation=cont)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return r.host(), ms, name


old, ms_old, k_old = run({"FHMC_SOLVER_LANES": "32"})
print("old  %-18s %8.3f ms  %.3e solves/s  evals %.2f" % (k_old, ms_old, T / ms_old * 1e3, old["iters"].mean()))
for cta in ("512",):
    _lib.lean_stats(reset=True)
    new, ms, k = run({"FHMC_SOLVER_CTA": cta})
    print("   lean stats (4 launches):", _lib.lean_stats())
    ok = (old["code"] == 0)
    lean = (new["status"] & 0x4000) != 0
    print("lean %-18s cta %s %8.3f ms  %.3e solves/s  evals %.2f  lean-final %.3f  jump %.4f" %
          (k, cta, ms, T / ms * 1e3, new["iters"].mean(), lean.mean(), ((new["status"] & 0x2000) != 0).mean()))
    print("   codes equal", np.array_equal(old["code"], new["code"]), " ok frac", ok.mean(),
          " max |dmu|", np.abs(old["mu_coex"][ok] - new["mu_coex"][ok]).max(),
          " iters equal", np.mean(old["iters"] == new["iters"]),
          " nphase equal", np.array_equal(old["nphase"][ok], new["nphase"][ok]),
          " bounds equal", np.array_equal(old["bounds"][ok], new["bounds"][ok]),
          " max_idx equal", np.array_equal(old["max_idx"][ok], new["max_idx"][ok]),
          " fe rel", np.nanmax(np.abs(old["fe"][ok, :2] - new["fe"][ok, :2]) / np.abs(old["fe"][ok, :2])),
          " avg rel", np.nanmax(np.abs(old["avg"][ok, :2] - new["avg"][ok, :2]) / np.abs(old["avg"][ok, :2])))
    conv = ok & ((new["status"] & 0x2000) == 0)
    print("   converged (code 0, no jump): %.4f  max|dfe| %.3e ; jump records max|dfe| %.3e" %
          (conv.mean(), np.abs(new["dfe"][conv]).max(), np.abs(new["dfe"][ok & ~conv]).max() if (ok & ~conv).any() else 0.0))
st, ms_st, _ = run({"CONT": "1"})
okb = ok & (st["code"] == 0)
cv = okb & ((new["status"] & 0x2000) == 0) & ((st["status"] & 0x2000) == 0)
print("staged continuation: %.3f ms  %.3e solves/s  evals(last stage) %.2f  codes equal %s  max|dmu| (converged both) %.3e  bounds equal %s" %
      (ms_st, T / ms_st * 1e3, st["iters"].mean(), np.array_equal(st["code"], new["code"]), np.abs(st["mu_coex"][cv] - new["mu_coex"][cv]).max(),
       np.array_equal(st["bounds"][cv][:, :2], new["bounds"][cv][:, :2])))
it = new["iters"]
print("iters: ok mean %.2f max %d | failed mean %.2f max %d | hist(ok)" % (it[ok].mean(), it[ok].max(), it[~ok].mean() if (~ok).any() else 0, it[~ok].max() if (~ok).any() else 0),
      np.bincount(np.minimum(it[ok], 30))[:31].tolist())
print("codes of failed:", np.unique(new["code"][~ok], return_counts=True))
# time with the supercritical tail excluded
sub = np.where(ok)[0]
b2, g2 = betas[sub], g[sub]
for cta in ("1024", "512"):
    os.environ["FHMC_SOLVER_CTA"] = cta
    r = dh.find_phase_eq(g2, beta=b2, lnz_tol=1e-10, pmax=4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = dh.find_phase_eq(g2, beta=b2, lnz_tol=1e-10, pmax=4, continuation=False)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("only the %d solvable: cta %s %.3f ms  %.3e solves/s" % (len(sub), cta, ms, len(sub) / ms * 1e3))
os.environ["FHMC_SOLVER_CTA"] = "512"
hard = np.where(ok & (it > 10))[0]
print("hard solves (iters > 10):", hard.tolist(), it[hard].tolist(), "jump:", ((new["status"][hard] & 0x2000) != 0).tolist())
for name, sel_ in (("easy only", np.where(ok & (it <= 10))[0]),):
    b2, g2 = betas[sel_], g[sel_]
    r = dh.find_phase_eq(g2, beta=b2, lnz_tol=1e-10, pmax=4, continuation=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = dh.find_phase_eq(g2, beta=b2, lnz_tol=1e-10, pmax=4)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-16s %d solves: %.3f ms  %.3e solves/s" % (name, len(sel_), ms, len(sel_) / ms * 1e3))

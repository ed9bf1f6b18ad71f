#!/usr/bin/env python
"""Secondary measurements (BASELINE configs 3, 4, 5 + small-batch/drop-in latencies); one JSON line each.
Not the driver's bench (that is bench.py / config 2); results are copied to profiles/ for DESIGN.md."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fhmcanalysis_b200 import engine, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    peaks = engine.measure_peaks()
    hbm = None
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = []

    # ---- config 3: 2-component Taylor grid 4096 beta x 4096 dmu2 (order-2 lnPI, skip_mom) ----------------
    n = 1001
    lnpi, mom2 = synth.two_peak_lnpi(n), synth.two_comp_moments(n)
    h = histogram.from_arrays(lnpi, mom2, 1.0, [-3.0, -2.5], 10)
    h.reweight(-2.9)
    betas, dmus = np.linspace(0.95, 1.05, 4096), np.linspace(0.2, 0.8, 4096)
    for order, moments, tag in ((2, (), "order-2 lnPI (skip_mom)"), (1, ("N1", "N2", "U"), "order-1 lnPI + <N1>,<N2>,<U>")):
        dh = h.device_histogram(beta=betas, dmu=dmus, order=order, moments=moments)
        st = dh.make_states(np.array([-2.9]), betas, dmus, grid=True)
        S = st.n_states
        res = engine.SweepResult(S, 8, dh.n_sel, dh.device)
        ms = timed(lambda: dh.sweep(None, states=st, out=res, pmax=8), reps=3, warm=1)
        code = (res.status & 0xFF)
        okf = float((code == 0).double().mean().item())
        exps = S * n * 1.0 / (ms * 1e-3)
        out.append({"config": "3: synthetic 2-comp Taylor grid 4096 beta x 4096 dmu2, N_max=1000, " + tag, "state_points": S, "ms": ms,
                    "value": S / (ms * 1e-3), "unit": "state points/s", "ok_fraction": okf, "blob_rows": dh.n_rows,
                    "roofline": {"bound": "fp64_exp", "achieved_gexp_s": exps / 1e9, "peak_gexp_s": peaks["exp_per_s"] / 1e9,
                                 "frac": exps / peaks["exp_per_s"], "note": "k_sweep_rowc for the lnPI-only grid (rows combined per grid row), k_sweep_fast Taylor variant with averaged quantities; one true exp per bin"}})
        del res

    # ---- config 4: coexistence curve over 10^4 temperatures, N_max=2000, smooth=10, order 2 --------------
    n4 = 2001
    lnpi4 = synth.two_peak_lnpi(n4, scale=2.0)
    mom4 = synth.one_comp_moments(n4, max_order=3)
    h4 = histogram.from_arrays(lnpi4, mom4, 1.0, [0.0], 10)
    T = np.linspace(0.90, 1.10, 10000)
    betas4 = 1.0 / T
    # continuation for the guesses: coarse solve outward from beta_ref, then interpolate
    coarse = np.linspace(betas4.min(), betas4.max(), 41)
    order_idx = np.argsort(np.abs(coarse - 1.0))
    guess = {}
    g = 0.0
    t0 = time.perf_counter()
    for k in order_idx:
        near = [guess[j] for j in guess if abs(coarse[j] - coarse[k]) < 2.5 * (coarse[1] - coarse[0])]
        g0 = near[-1] if near else g
        r = h4.find_phase_eq_batch(np.array([coarse[k]]), g0, order=2, lnZ_tol=1e-10)
        if r["code"][0] == 0:
            guess[k] = float(r["mu_coex"][0])
    t_coarse = time.perf_counter() - t0
    ks = sorted(guess)
    mu_guess = np.interp(betas4, coarse[ks], [guess[k] for k in ks])
    dh4 = h4.device_histogram(beta=betas4, order=2, moments=("N", "N2", "U"))
    res4 = {}

    def solve():
        res4["r"] = dh4.find_phase_eq(mu_guess, beta=betas4, lnz_tol=1e-10, pmax=4)
    ms4 = timed(solve, reps=3, warm=1)
    r4 = res4["r"].host()
    okf = float(np.mean(r4["code"] == 0))
    iters = float(np.mean(r4["iters"]))
    exps4 = float(np.sum(r4["iters"])) * n4 * 2 / (ms4 * 1e-3)
    out.append({"config": "4: batched find_phase_eq over 10^4 temperatures, N_max=2000, smooth=10, order-2 beta extrapolation",
                "solves": len(T), "ms": ms4, "value": len(T) / (ms4 * 1e-3), "unit": "coexistence points/s", "ok_fraction": okf,
                "mean_evaluations_per_solve": iters, "max_abs_dfe": float(np.max(np.abs(r4["dfe"][r4["code"] == 0]))),
                "coarse_continuation_s": t_coarse,
                "roofline": {"bound": "fp64_exp", "achieved_gexp_s": exps4 / 1e9, "peak_gexp_s": peaks["exp_per_s"] / 1e9,
                             "frac": exps4 / peaks["exp_per_s"], "note": "2 passes x N exp per evaluation (generic evaluator, 32 lanes per solve)"}})

    # ---- config 5: 2-D joint (N1,N2) 512x512, 10^5 (mu1,mu2) pairs ----------------------------------------
    n1 = n2 = 512
    lnpi2d, bounds = synth.joint_2d(n1, n2, 640)
    g1, g2 = np.meshgrid(np.linspace(-0.02, 0.02, 316), np.linspace(-0.02, 0.02, 317), indexing="ij")
    a1, a2 = g1.ravel()[:100000].copy(), g2.ravel()[:100000].copy()
    dev = engine.require_cuda()
    tl = torch.from_numpy(lnpi2d).to(dev)
    tb = torch.from_numpy(bounds).to(dev)
    to1, to2 = torch.arange(n1, dtype=torch.float64, device=dev), torch.arange(n2, dtype=torch.float64, device=dev)
    ta1, ta2 = torch.from_numpy(a1).to(dev), torch.from_numpy(a2).to(dev)
    ms5e = timed(lambda: engine.reweight_2d(tl, tb, to1, to2, ta1, ta2, None, return_device=True, product=False), reps=3, warm=1)
    ms5 = timed(lambda: engine.reweight_2d(tl, tb, to1, to2, ta1, ta2, None, return_device=True, product=True), reps=3, warm=1)
    support = int(np.sum(bounds[:, 1] - bounds[:, 0]))
    S5 = len(a1)
    exps5 = S5 * support / (ms5 * 1e-3)
    streamed = S5 * 8.0 * n1 * n2 / (ms5 * 1e-3) / 1e9
    out.append({"config": "5: two_dim joint (N1,N2) 512x512 lnPI reweighting over 10^5 (mu1,mu2) pairs", "state_points": S5, "ms": ms5,
                "value": S5 / (ms5 * 1e-3), "unit": "state points/s", "support_bins": support,
                "kernel": "product form (k_rw2d_tables + k_rw2d_prod + k_rw2d_merge), tables rebuilt inside every call",
                "exp_per_bin_kernel_ms": ms5e, "exp_per_bin_kernel_value": S5 / (ms5e * 1e-3),
                "roofline": {"bound": "fp64_exp", "achieved_gexp_s": exps5 / 1e9, "peak_gexp_s": peaks["exp_per_s"] / 1e9,
                             "frac": exps5 / peaks["exp_per_s"],
                             "hbm_if_streamed": {"algorithmic_bytes_per_state_point": 8 * n1 * n2, "equivalent_gbs": streamed, "peak_gbs": hbm,
                                                 "frac": (streamed / hbm) if hbm else None,
                                                 "note": "a staged row chunk is reused by 128 state points, so real DRAM traffic is ~128x lower (see ncu)"}}})

    # ---- small batches / drop-in latencies --------------------------------------------------------------
    lnpi1, mom1 = synth.two_peak_lnpi(1001), synth.one_comp_moments(1001)
    h1 = histogram.from_arrays(lnpi1, mom1, 1.0, [0.0], 10)

    def dropin():
        hh = histogram.from_arrays(lnpi1, mom1, 1.0, [0.0], 10)
        hh.reweight(0.01)
        hh.thermo()
        hh.is_safe()
    t0 = time.perf_counter()
    for _ in range(20):
        dropin()
    lat = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(5):
        h1.find_phase_eq(1e-6, 0.0)
    lat_eq = (time.perf_counter() - t0) / 5
    out.append({"config": "drop-in scalar API latency (reweight + thermo(all 27 moments) + is_safe), N_max=1000", "ms_per_state_point": lat * 1e3,
                "find_phase_eq_ms": lat_eq * 1e3})
    for S_, G in ((1000, 32), (10000, 32), (100000, 4)):
        dh1 = h1.device_histogram(moments=("N", "N2"))
        mu = torch.linspace(-0.03, 0.03, S_, dtype=torch.float64, device=dev)
        st = dh1.make_states(mu)
        r = engine.SweepResult(S_, 4, dh1.n_sel, dev)
        ms = timed(lambda: dh1.sweep(None, states=st, out=r, pmax=4), reps=5, warm=2)
        out.append({"config": "config-2 histogram, %d state points (auto lanes)" % S_, "ms": ms, "value": S_ / (ms * 1e-3), "unit": "state points/s"})
    # ---- isopleth.make_grid_multi over a (mu1, dmu2) grid (reference: ~105 ms per grid cell at N = 201,
    #      SURVEY 8(c) [probed]) -------------------------------------------------------------------------------------------
    try:
        import io as _io
        from contextlib import redirect_stdout
        from fhmcanalysis_b200.moments.histogram.one_dim.ntot import gc_binary as gcB
        n_iso = 1001
        mom_iso = synth.two_comp_moments(n_iso)
        hs = []
        for k, d2 in enumerate((-0.2, 0.5, 1.2)):
            ln_k = synth.two_peak_lnpi(n_iso, noise=1e-3, scale=1.0, seed=100 + k) + 0.0004 * k * np.arange(n_iso)
            hs.append(histogram.from_arrays(ln_k, mom_iso, 1.0, [-3.0, -3.0 + d2], 5, volume=512.0))
        iso = gcB.isopleth(hs, 1.0, 1)
        with redirect_stdout(_io.StringIO()):
            iso.make_grid_multi([-3.02, -2.98], [0.0, 1.0], [0.0005, 0.0125], 2.5)      # warm-up (builds device histograms)
            t0 = time.perf_counter()
            Z, (X, Y) = iso.make_grid_multi([-3.02, -2.98], [0.0, 1.0], [0.0005, 0.0125], 2.5)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out.append({"config": "isopleth.make_grid_multi, 3 two-species histograms (N_max=1000), order 1, %d x %d grid" % Z.shape,
                    "grid_cells": int(Z.size), "ms": dt * 1e3, "value": Z.size / dt, "unit": "grid cells/s",
                    "filled_fraction": float(np.mean(Z != 0)),
                    "note": "wall clock of the whole drop-in call (host orchestration + batched kernels + results to host)"})
    except Exception as e:  # secondary number
        out.append({"config": "isopleth.make_grid_multi", "error": repr(e)})

    # ---- HBM-bound streaming kernels, device-resident, through the C ABI -------------------
    import ctypes
    from fhmcanalysis_b200 import _lib
    L = _lib.load()
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for n1m, n2m, nprop in ((512, 512, 2), (4096, 4096, 2)):
        rng = np.random.default_rng(3)
        lp = torch.from_numpy(rng.normal(size=(n1m, n2m))).to(dev)
        pr = torch.from_numpy(rng.normal(size=(nprop, n1m, n2m))).to(dev)
        mk = torch.ones((n1m, n2m), dtype=torch.uint8, device=dev)
        o_t = torch.empty(2 + 8, dtype=torch.float64, device=dev)
        pk = torch.empty(65, dtype=torch.int64, device=dev)
        wsb = int(L.fhmc_masked_lse_2d_workspace(n1m, n2m, nprop))
        ws_t = torch.empty(wsb // 8 + 2, dtype=torch.float64, device=dev)

        def m2d():
            _lib.check(L.fhmc_masked_lse_2d(lp.data_ptr(), mk.data_ptr(), None, n1m, n2m, pr.data_ptr(), nprop, o_t.data_ptr(), pk.data_ptr(), 64,
                                            None, ws_t.data_ptr(), wsb, sp), "fhmc_masked_lse_2d")
        msm = timed(m2d, reps=10, warm=3)
        bytes_m = 8.0 * n1m * n2m * (2 + nprop) + 2.0 * n1m * n2m
        out.append({"config": "8f-3 pore_hist.thermo(mask): fhmc_masked_lse_2d %dx%d, %d property matrices" % (n1m, n2m, nprop), "ms": msm,
                    "roofline": {"bound": "hbm", "algorithmic_bytes": bytes_m, "achieved_gbs": bytes_m / (msm * 1e-3) / 1e9, "peak_gbs": hbm,
                                 "frac": (bytes_m / (msm * 1e-3) / 1e9 / hbm) if hbm else None,
                                 "note": "4 launches; the 512x512 case is launch-latency bound (2 MiB per matrix)"}})
    for W, ln in ((10000, 1000), (200, 400)):
        a_t = torch.randn(W * ln, dtype=torch.float64, device=dev)
        b_t = a_t + 3.0
        off = torch.arange(0, (W + 1) * ln, ln, dtype=torch.int64, device=dev)
        sh = torch.empty((2, W), dtype=torch.float64, device=dev)

        def pshift():
            _lib.check(L.fhmc_patch_shifts(a_t.data_ptr(), b_t.data_ptr(), off.data_ptr(), W, sh[0].data_ptr(), sh[1].data_ptr(), sp), "fhmc_patch_shifts")
        msp = timed(pshift, reps=10, warm=3)
        bytes_p = 16.0 * W * ln
        out.append({"config": "8f-4 window patching: fhmc_patch_shifts, %d window pairs x %d overlapping bins" % (W, ln), "ms": msp,
                    "value": W / (msp * 1e-3), "unit": "window pairs/s",
                    "roofline": {"bound": "hbm", "algorithmic_bytes": bytes_p, "achieved_gbs": bytes_p / (msp * 1e-3) / 1e9, "peak_gbs": hbm,
                                 "frac": (bytes_p / (msp * 1e-3) / 1e9 / hbm) if hbm else None,
                                 "note": "second pass over the same slices comes from L2"}})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()

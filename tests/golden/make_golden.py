#!/usr/bin/env python
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, built from /root/reference by
oracle/build_ref.py).  Run in the build container only; the GPU box uses the committed files.

Every vector is an output of the reference's own code path (jeetain/FHMCAnalysis,
moments/histogram/one_dim/ntot/gc_hist.pyx / gc_binary.pyx, two_dim/joint_hist.pyx) driven the way its
unit tests and notebooks drive it:  fresh histogram -> reweight(mu) -> [temp_dmu_extrap] -> thermo()
-> is_safe();  find_phase_eq;  mix;  temp_dmu_extrap_multi;  isopleth.make_grid_multi;  joint_hist.make.
"""
import copy
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from fhmcanalysis_b200 import synth  # noqa: E402
from fhmcanalysis_b200.io.hdf5_min import Dataset  # noqa: E402

REF = os.environ.get("FHMC_REFERENCE", "/root/reference")


def thermo_record(h, nsel_arrays=None):
    """Flatten what thermo()/is_safe() left in a reference histogram."""
    P = len(h.data["thermo"])
    rec = {
        "lnpi": np.array(h.data["ln(PI)"], dtype=np.float64),
        "maxima": np.array(h.data["ln(PI)_maxima_idx"], dtype=np.int64),
        "minima": np.array(h.data["ln(PI)_minima_idx"], dtype=np.int64),
        "fe": np.array([h.data["thermo"][p]["F.E./kT"] for p in range(P)]),
        "bounds": np.array([h.data["thermo"][p]["bound_idx"] for p in range(P)], dtype=np.int64),
        "safe": np.array(bool(h.is_safe())),
    }
    if "mom" in h.data["thermo"][0]:
        rec["mom"] = np.array([h.data["thermo"][p]["mom"] for p in range(P)])
    return rec


def pack(prefix, rec, out):
    for k, v in rec.items():
        out["%s/%s" % (prefix, k)] = v


def main():
    ns = ref.load()
    if ns is None:
        raise SystemExit("compiled reference unavailable: %s" % ref._cache.get("error"))
    H = ns.gc_hist.histogram
    out = {}
    meta = {}

    # ---- A. reference unit-test fixture test.nc as the reference itself loads it (T1:46-66) -----
    h = H(os.path.join(REF, "unittests/reference/test.nc"), 1.0, [5.0, 0.0], 1)
    out["testnc/lnpi"] = h.data["ln(PI)"].copy()
    out["testnc/ntot"] = h.data["ntot"].copy()
    out["testnc/mom"] = h.data["mom"].copy()
    meta["testnc"] = {"volume": h.data["volume"], "max_order": h.data["max_order"], "nspec": int(h.data["nspec"]),
                      "history": str(h.metadata["file_history"])}
    # reweight known answers (T1:100-147): cumulative reweighting 5 -> 0 -> -5
    h.reweight(0.0)
    out["testnc/rew0"] = h.data["ln(PI)"].copy()
    h.reweight(-5.0)
    out["testnc/rew0_m5"] = h.data["ln(PI)"].copy()

    # ---- B. T1 31-bin two-peak array: thermo / thermo(complete) / is_safe quartet / find_phase_eq ----
    t1 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 5, 4, 3, 2, 1, 0], dtype=np.float64)
    mom = np.ones((2, 3, 2, 3, 3, 31))
    mom[0, 1, 0, 0, :] = np.arange(31)
    mom[1, 1, 0, 0, :] = np.arange(31) * 2
    h = ref.make_histogram(t1, mom, 1.0, [5.0, 0.0], 1, volume=729.0)
    h.thermo()
    pack("t1/thermo", thermo_record(h), out)
    out["t1/is_safe"] = np.array([h.is_safe(10.0), h.is_safe(5.0), h.is_safe(10.0, True), h.is_safe(10.1, True)])
    h = ref.make_histogram(t1, mom, 1.0, [5.0, 0.0], 1, volume=729.0)
    h.thermo(True, True)
    out["t1/complete/mom"] = h.data["thermo"][0]["mom"]
    out["t1/complete/fe"] = np.array(h.data["thermo"][0]["F.E./kT"])
    h = ref.make_histogram(t1, mom, 1.0, [5.0, 0.0], 1, volume=729.0)
    with redirect_stdout(io.StringIO()):
        eq, err = h.find_phase_eq(0.001, 5.0, reterr=True)
    out["t1/phase_eq/mu"] = eq.data["curr_mu"].copy()
    out["t1/phase_eq/fe"] = np.array([eq.data["thermo"][p]["F.E./kT"] for p in range(2)])
    out["t1/phase_eq/err"] = np.array(err)
    # relextrema integer arrays (T1:155-198), smooth=1, on the raw arrays
    for k, arr in enumerate(([1, 2, 3, 2, 1, 2, 3, 4, 5], [1, 2, 3, 2, 1, 2], [1, 2, 3, 2, 1], [2, 1, 2, 3, 2, 1])):
        h = ref.make_histogram(np.array(arr, dtype=np.float64), np.ones((2, 3, 2, 3, 3, len(arr))), 1.0, [5.0, 0.0], 1)
        h.relextrema()
        out["t1/relext%d/x" % k] = np.array(arr, dtype=np.float64)
        out["t1/relext%d/maxima" % k] = np.array(h.data["ln(PI)_maxima_idx"], dtype=np.int64)
        out["t1/relext%d/minima" % k] = np.array(h.data["ln(PI)_minima_idx"], dtype=np.int64)

    # ---- C. synthetic 1-component sweep (config 2 generator at N=301): reweight -> thermo -> is_safe ----
    n = 301
    lnpi = synth.two_peak_lnpi(n, noise=1e-3, scale=0.3)
    mom1 = synth.one_comp_moments(n)
    out["c2/lnpi"] = lnpi
    out["c2/mom"] = mom1
    mus = np.linspace(-0.12, 0.10, 23)
    out["c2/mu"] = mus
    meta["c2"] = {"beta_ref": 1.0, "mu_ref": [0.0], "smooth": 5, "n": n}
    for k, mu in enumerate(mus):
        h = ref.make_histogram(lnpi, mom1, 1.0, [0.0], 5)
        h.reweight(mu)
        h.thermo()
        pack("c2/%d" % k, thermo_record(h), out)
    # smooth / noise stress: smooth 1 and 30 at two chemical potentials, statuses when the reference raises
    stress = []
    for smooth in (1, 2, 30, 60):
        for noise in (0.0, 5e-2):
            ln2 = synth.two_peak_lnpi(n, noise=noise, scale=0.3, seed=7)
            for mu in (-0.05, 0.0, 0.2):
                h = ref.make_histogram(ln2, mom1, 1.0, [0.0], smooth)
                h.reweight(mu)
                key = "stress/s%d_n%g_m%g" % (smooth, noise, mu)
                try:
                    h.thermo(props=False)
                    pack(key, thermo_record(h), out)
                    stress.append([key, smooth, noise, mu, "ok"])
                except Exception as e:  # the reference raises (e.g. non-alternating extrema)
                    stress.append([key, smooth, noise, mu, "raise: " + str(e)[:60]])
                out[key + "/input"] = ln2
    meta["stress"] = stress
    # monotone ln(PI) (both-empty branch, GH:382-386)
    mono = -0.05 * np.arange(n, dtype=np.float64)
    h = ref.make_histogram(mono, mom1, 1.0, [0.0], 5)
    h.thermo(props=False)
    pack("mono", thermo_record(h), out)
    out["mono/input"] = mono

    # ---- D. square-well T*=0.90 composite: notebook known answer (example.ipynb cell 14) ----------
    sw = os.path.join(REF, "example/ntot/square_well/T_0.90/composite.nc")
    d = Dataset(sw)
    out["sw/lnpi"] = d.variables["ln(PI)"][:]
    out["sw/mom"] = d.variables["N_{i}^{j}*N_{k}^{m}*U^{p}"][:]
    meta["sw"] = {"beta_ref": 1.0 / 0.9, "mu_ref": [0.0], "smooth": 10, "volume": float(d.volume),
                  "lnZ_tol": 1e-6, "mu_guess": -3.94, "notebook_beta_mu": -4.47264655,
                  "notebook_fe": [-9.28506932479, -9.28546354084]}
    h = H(sw, 1.0 / 0.9, [0.0], 10)
    with redirect_stdout(io.StringIO()):
        eq, err = h.find_phase_eq(1e-6, -3.94, 1.0 / 0.9, reterr=True)
    out["sw/phase_eq/mu"] = eq.data["curr_mu"].copy()
    out["sw/phase_eq/err"] = np.array(err)
    pack("sw/phase_eq", thermo_record(eq), out)
    # a short mu sweep on the real data (dynamic range -2506..0)
    sw_mus = np.array([-4.6, -4.4726, -4.3, -3.94])
    out["sw/mu"] = sw_mus
    for k, mu in enumerate(sw_mus):
        h = H(sw, 1.0 / 0.9, [0.0], 10)
        h.reweight(mu)
        h.thermo()
        pack("sw/%d" % k, thermo_record(h), out)
    # first/second order temperature extrapolation of lnPI on the real data (skip_mom=True)
    for order in (1, 2):
        h = H(sw, 1.0 / 0.9, [0.0], 10)
        h.reweight(-4.47)
        hn = h.temp_extrap(1.0 / 0.92, order, 10.0, False, True, True)
        out["sw/textrap%d" % order] = hn.data["ln(PI)"].copy()

    # ---- E. 2-component Taylor grid (config 3 generator at N=201) -------------------------------
    n2 = 201
    lnpi2 = synth.two_peak_lnpi(n2, noise=1e-3, scale=0.2)
    mom2 = synth.two_comp_moments(n2)
    out["c3/lnpi"] = lnpi2
    out["c3/mom"] = mom2
    betas = np.array([0.97, 1.0, 1.04])
    dmus = np.array([[0.3], [0.5], [0.75]])
    out["c3/betas"], out["c3/dmus"] = betas, dmus
    meta["c3"] = {"beta_ref": 1.0, "mu_ref": [-3.0, -2.5], "smooth": 5, "mu1": -2.9}
    for order in (1, 2):
        h = ref.make_histogram(lnpi2, mom2, 1.0, [-3.0, -2.5], 5)
        h.reweight(-2.9)
        hs = h.temp_dmu_extrap_multi(betas, dmus, order, 10.0, True, True)
        for a in range(3):
            for b in range(3):
                hh = hs[a][b]
                out["c3/o%d/%d_%d/lnpi" % (order, a, b)] = hh.data["ln(PI)"].copy()
                hh.thermo(props=False)
                rec = thermo_record(hh)
                for k in ("maxima", "minima", "fe", "bounds", "safe"):
                    out["c3/o%d/%d_%d/%s" % (order, a, b, k)] = rec[k]
    # order-1 with moments (skip_mom=False): extrapolated moment tensor + phase averages
    h = ref.make_histogram(lnpi2, mom2, 1.0, [-3.0, -2.5], 5)
    h.reweight(-2.9)
    hn = h.temp_dmu_extrap(1.03, np.array([0.6]), 1, 10.0, True, True, False)
    out["c3/mom1/lnpi"] = hn.data["ln(PI)"].copy()
    out["c3/mom1/mom"] = hn.data["mom"].copy()
    hn.thermo()
    pack("c3/mom1/thermo", thermo_record(hn), out)
    # mix (GH:184-258)
    ha = ref.make_histogram(lnpi2, mom2, 1.0, [-3.0, -2.5], 5)
    hb = ref.make_histogram(lnpi2[:150] * 1.01, mom2[..., :150] * 0.99, 1.0, [-3.0, -2.5], 5)
    hm = ha.mix(hb, [0.3, 0.9])
    out["mix/lnpi"] = hm.data["ln(PI)"].copy()
    out["mix/mom_sample"] = hm.data["mom"][1, 1, 0, 1, 1].copy()

    # ---- F. isopleth.make_grid_multi on three synthetic 2-species histograms ---------------------
    hists = []
    iso_in = []
    for k, d2 in enumerate((-0.2, 0.5, 1.2)):
        ln_k = synth.two_peak_lnpi(n2, noise=1e-3, scale=0.2, seed=100 + k) + 0.002 * k * np.arange(n2)
        hk = ref.make_histogram(ln_k, mom2, 1.0, [-3.0, -3.0 + d2], 5, volume=512.0)
        hists.append(hk)
        iso_in.append(ln_k)
    out["iso/lnpi"] = np.array(iso_in)
    meta["iso"] = {"dmu2": [-0.2, 0.5, 1.2], "mu1_bounds": [-3.1, -2.9], "dmu2_bounds": [0.0, 1.0], "delta": [0.1, 0.25],
                   "beta_ref": 1.0, "mu1_ref": -3.0, "smooth": 5, "volume": 512.0, "order": 1, "m": 2.5}
    try:
        with redirect_stdout(io.StringIO()):
            iso = ns.gc_binary.isopleth(hists, 1.0, 1)
            Z, (X, Y) = iso.make_grid_multi([-3.1, -2.9], [0.0, 1.0], [0.1, 0.25], 2.5)
        out["iso/x1"], out["iso/density"], out["iso/fe"] = np.array(Z), np.array(iso.data["density"]), np.array(iso.data["F.E./kT"])
        out["iso/X"], out["iso/Y"] = np.array(X), np.array(Y)
        meta["iso"]["status"] = "ok"
    except Exception as e:  # keep going: the isopleth path is untested upstream
        meta["iso"]["status"] = "reference raised: %r" % (e,)

    # ---- G. joint_hist container (two_dim/joint_hist.pyx) ---------------------------------------
    jh = ns.joint_hist.joint_hist()
    ent = [(2.0, {"op2": [0.0, 1.0, 2.0, 3.0, 4.0], "lnPI": [0.0, 1.0, 2.0, 3.0, 4.0], "props": {"e": [3.0, 4.0, 5.0, 6.0, 7.0]}}),
           (1.0, {"op2": [1.0, 2.0, 3.0], "lnPI": [1.0, 2.0, 3.0], "props": {"e": [1.0, 2.0, 3.0]}})]
    try:
        for op1, dct in ent:
            jh.enter(op1, np.array(dct["lnPI"]), np.array(dct["op2"]), {k: np.array(v) for k, v in dct["props"].items()})
        jh.make()
        out["joint/lnpi"] = np.array(jh.data["ln(PI)"])
        out["joint/bounds"] = np.array(jh.data["bounds_idx"])
        out["joint/op1"] = np.array(jh.data["op_1"])
        out["joint/op2"] = np.array(jh.data["op_2"])
        out["joint/prop_e"] = np.array(jh.data["props"]["e"])
        meta["joint"] = {"status": "ok", "entries": ent}
    except Exception as e:
        meta["joint"] = {"status": "reference API differs: %r" % (e,)}

    # ---- H. derivative builders and every extrapolation entry point on a max_order-4, 2-species tensor (like
    #         unittests/reference/test2.nc), with and without the kinetic-energy terms -----------------------------
    n4 = 61
    rng = np.random.default_rng(4)
    x4 = np.arange(n4, dtype=np.float64)
    lnpi4 = synth.two_peak_lnpi(n4, noise=1e-3, scale=0.06, seed=11) * 1.0
    base = [0.4 * x4 + 0.3, 0.6 * x4 + 0.2]
    uu = -1.5 * x4 - 0.01 * x4 * x4 - 0.5
    mom4 = np.zeros((2, 5, 2, 5, 5, n4))
    for i in range(2):
        for j in range(5):
            for k in range(2):
                for m_ in range(5):
                    for p in range(5):
                        mom4[i, j, k, m_, p] = base[i] ** j * base[k] ** m_ * uu ** p * (1.0 + 0.02 * (j + m_ + p) * rng.random(n4))
    out["h/lnpi"], out["h/mom"] = lnpi4, mom4
    meta["h"] = {"beta_ref": 1.0, "mu_ref": [-2.0, -1.6], "smooth": 3, "mu1": -1.9, "beta": 1.04, "dmu": 0.55}
    sample = [[0, 1, 0, 0, 0], [1, 1, 0, 0, 0], [0, 0, 0, 0, 1], [0, 1, 1, 1, 0], [1, 2, 0, 0, 1], [0, 1, 0, 1, 1], [1, 1, 1, 1, 1]]
    meta["h"]["sample"] = sample

    def pick(t, lead=()):
        return np.array([t[tuple(lead) + tuple(a)] for a in sample])

    for ke in (False, True):
        tag = "h/ke%d" % int(ke)
        hh = ref.make_histogram(lnpi4, mom4, 1.0, [-2.0, -1.6], 3, ke=ke)
        hh.reweight(-1.9)
        d1, dm1 = hh._dB(False)
        d2, dm2 = hh._dB2(False)
        out[tag + "/dB"], out[tag + "/dB_mom"], out[tag + "/dB_sum"] = d1, pick(dm1), np.array(np.sum(np.abs(dm1)))
        out[tag + "/dB2"], out[tag + "/dB2_mom"], out[tag + "/dB2_sum"] = d2, pick(dm2), np.array(np.sum(np.abs(dm2)))
        dmu1, dmm1 = hh._dMU(False)
        H2, Hm2 = hh._dMU2(False)
        out[tag + "/dMU"], out[tag + "/dMU_mom"] = dmu1, pick(dmm1, (0,))
        out[tag + "/dMU2"], out[tag + "/dMU2_mom"] = H2, pick(Hm2, (0, 0))
        Hl, Hm = hh._dBMU2(False)
        out[tag + "/dBMU2"], out[tag + "/dBMU2_mom01"], out[tag + "/dBMU2_sum"] = Hl, pick(Hm, (0, 1)), np.array(np.sum(np.abs(Hm)))
        out[tag + "/gc"] = np.array([hh._gc_dX_dB([0, 1, 0, 0, 0], 0), hh._gc_dX_dB([0, 0, 0, 0, 1], 1), hh._gc_d2X_dB2([1, 1, 0, 0, 0], 0),
                                     hh._gc_df_dB_ii(([0, 1, 0, 0, 0], 0), ([0, 0, 0, 0, 1], 0)), hh._gc_df_dB_in(([1, 1, 0, 0, 0], 0), 1),
                                     hh._gc_fluct_ii([0, 1, 0, 0, 0], [1, 1, 0, 0, 0])])
        if not ke:
            d3, dm3 = hh._dB3(False)
            out[tag + "/dB3"], out[tag + "/dB3_mom"] = d3, pick(dm3)
        for order in ((1, 2, 3) if not ke else (1, 2)):
            hn = hh.temp_extrap(1.04, order, 10.0, True, True, False)
            out[tag + "/temp%d/lnpi" % order], out[tag + "/temp%d/mom" % order] = hn.data["ln(PI)"].copy(), pick(hn.data["mom"])
        for order in (1, 2):
            hn = hh.dmu_extrap(np.array([0.55]), order, 10.0, True, True, False)
            out[tag + "/dmu%d/lnpi" % order], out[tag + "/dmu%d/mom" % order] = hn.data["ln(PI)"].copy(), pick(hn.data["mom"])
            for fom in (False, True):
                hn = hh.temp_dmu_extrap(1.04, np.array([0.55]), order, 10.0, True, True, False, fom)
                out[tag + "/tdmu%d_%d/lnpi" % (order, int(fom))] = hn.data["ln(PI)"].copy()
                out[tag + "/tdmu%d_%d/mom" % (order, int(fom))] = pick(hn.data["mom"])
    # find_phase_eq at another temperature / dmu (reference Nelder-Mead), orders 1 and 2, on a wider two-peak surface
    n5 = 241
    lnpi5 = synth.two_peak_lnpi(n5, noise=0.0, scale=0.24)
    mom5 = synth.two_comp_moments(n5)
    mom5b = np.zeros((2, 4, 2, 4, 4, n5))
    mom5b[:, :3, :, :3, :3] = mom5
    for i in range(2):   # third-order entries so that order-2 moment extrapolation is defined (max_order 3)
        for j in range(4):
            for k in range(2):
                for m_ in range(4):
                    for p in range(4):
                        if j + m_ + p >= 3 or j == 3 or m_ == 3 or p == 3:
                            mom5b[i, j, k, m_, p] = (mom5[i, 1, 0, 0, 0] ** j) * (mom5[k, 1, 0, 0, 0] ** m_) * (mom5[0, 0, 0, 0, 1] ** p)
    out["eq/lnpi"], out["eq/mom"] = lnpi5, mom5b
    meta["eq"] = {"beta_ref": 1.0, "mu_ref": [-3.0, -2.5], "smooth": 5, "cases": []}
    for order, beta, dmu in ((1, 1.01, 0.5), (2, 0.985, 0.52), (1, 1.0, 0.5)):
        hh = ref.make_histogram(lnpi5, mom5b, 1.0, [-3.0, -2.5], 5)
        try:
            with redirect_stdout(io.StringIO()):
                eq, err = hh.find_phase_eq(1e-8, -3.0, beta, [dmu], order, 10.0, True, True)
            key = "eq/o%d_b%g_d%g" % (order, beta, dmu)
            out[key + "/mu"], out[key + "/err"] = eq.data["curr_mu"].copy(), np.array(err)
            pack(key, thermo_record(eq), out)
            meta["eq"]["cases"].append([key, order, beta, dmu, "ok"])
        except Exception as e:
            meta["eq"]["cases"].append(["", order, beta, dmu, "reference raised: " + str(e)[:80]])

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True, default=str)
    print("wrote %d arrays, %.1f kB" % (len(out), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")) / 1e3))
    print(json.dumps({k: (v.get("status") if isinstance(v, dict) else None) for k, v in meta.items()}))


if __name__ == "__main__":
    main()

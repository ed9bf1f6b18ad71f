"""NumPy/ctypes front-end of the CPU restatement oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  ``fhmcanalysis_b200`` never does: the product has no CPU fallback.

The arithmetic lives in ``fhmc_oracle.c`` (same libm and the same sequential evaluation order as
the compiled Cython reference, so integer outputs are bit-identical); this file only marshals
arrays and restates the host-level logic of
  * ``phase_eq_error`` / ``find_phase_eq`` (reference gc_hist.pyx:2570-2630, 598-668),
  * the closed-form Taylor coefficient arrays (reference gc_hist.pyx:1660-2563),
  * ``histogram.mix`` (gc_hist.pyx:184-258).

Parity status: pinned -- see tests/test_oracle_golden.py (reference unit-test known answers,
golden vectors generated from the compiled reference by tests/golden/make_golden.py).
"""
import ctypes
import math
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS = {0: "ok", 1: "too_short", 2: "bad_front", 3: "bad_back", 4: "count_mismatch",
          5: "not_sorted", 6: "index_error", 7: "ragged_gap", 8: "capacity"}

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_lp = ctypes.POINTER(ctypes.c_longlong)


def build():
    """Compile fhmc_oracle.c -> liboracle.so (gcc, seconds)."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "fhmc_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.fo_spec_exp.restype = ctypes.c_double
        L.fo_spec_exp.argtypes = [ctypes.c_double, ctypes.c_double]
        L.fo_normalize.restype = ctypes.c_double
        L.fo_normalize.argtypes = [_dp, ctypes.c_int]
        L.fo_reweight.restype = ctypes.c_double
        L.fo_reweight.argtypes = [_dp, _lp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.fo_relextrema.restype = ctypes.c_int
        L.fo_relextrema.argtypes = [_dp, ctypes.c_int, ctypes.c_int, _ip, _ip, _ip, _ip, ctypes.c_int, _ip]
        L.fo_phase_bounds.restype = ctypes.c_int
        L.fo_phase_bounds.argtypes = [ctypes.c_int, _ip, ctypes.c_int, _ip, ctypes.c_int, _ip]
        L.fo_free_energy.restype = ctypes.c_double
        L.fo_free_energy.argtypes = [_dp, ctypes.c_int, ctypes.c_int]
        L.fo_phase_averages.restype = ctypes.c_double
        L.fo_phase_averages.argtypes = [_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int, _dp]
        L.fo_is_safe.restype = ctypes.c_int
        L.fo_is_safe.argtypes = [_dp, ctypes.c_int, _ip, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.fo_state_point.restype = ctypes.c_int
        L.fo_state_point.argtypes = [_dp, _lp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                     _dp, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int,
                                     ctypes.c_int, _dp, _ip, _ip, _ip, _dp, _dp, _ip, _ip, _ip, _ip, _ip]
        L.fo_phase_eq_err2.restype = ctypes.c_double
        L.fo_phase_eq_err2.argtypes = [_dp, _ip, ctypes.c_int, ctypes.c_int]
        L.fo_reweight_2d.restype = None
        L.fo_reweight_2d.argtypes = [_dp, _ip, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_double,
                                     ctypes.c_double, _dp, ctypes.c_int, _dp]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def normalize(lnpi):
    """gc_hist.pyx:57-67.  Returns (normalised copy, lnNorm)."""
    x = np.array(lnpi, dtype=np.float64)
    c = lib().fo_normalize(_d(x), len(x))
    return x, c


def reweight(lnpi, ntot, mu1_new, curr_mu0, curr_beta):
    """gc_hist.pyx:71-78 (+ normalisation).  Returns the new normalised ln(PI)."""
    x = np.array(lnpi, dtype=np.float64)
    nt = np.ascontiguousarray(ntot, dtype=np.int64)
    lib().fo_reweight(_d(x), nt.ctypes.data_as(_lp), len(x), float(mu1_new), float(curr_mu0), float(curr_beta))
    return x


def relextrema(x, smooth):
    """gc_hist.pyx:317-415 on an (already normalised) array.  Returns (status, maxima, minima, info)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = len(x)
    cap = n + 2
    M = np.zeros(cap, dtype=np.int32)
    m = np.zeros(cap, dtype=np.int32)
    nM = ctypes.c_int(0)
    nm = ctypes.c_int(0)
    info = ctypes.c_int(0)
    st = lib().fo_relextrema(_d(x), n, int(smooth), _i(M), ctypes.byref(nM), _i(m), ctypes.byref(nm), cap,
                             ctypes.byref(info))
    return st, M[:nM.value].copy(), m[:nm.value].copy(), info.value


def phase_bounds(n, maxima, minima):
    """gc_hist.pyx:498-520."""
    M = np.ascontiguousarray(maxima, dtype=np.int32)
    m = np.ascontiguousarray(minima, dtype=np.int32)
    b = np.zeros((max(len(M), 1), 2), dtype=np.int32)
    st = lib().fo_phase_bounds(int(n), _i(M), len(M), _i(m), len(m), _i(b))
    return st, b[:len(M)]


def state_point(lnpi_ref, ntot, beta_ref, mu1_ref, mu1, smooth, cutoff=10.0, sel=None, coef=None, xi=None, pmax=None):
    """fresh copy -> reweight(mu1) [-> Taylor terms -> normalize] -> thermo() -> is_safe().

    Returns a dict with status, nphase, fe[P], avg[P,nsel], bounds[P,2], max_idx, min_idx, safe,
    lnpi (normalised array), info."""
    L = lib()
    lnpi_ref = np.ascontiguousarray(lnpi_ref, dtype=np.float64)
    n = len(lnpi_ref)
    nt = np.ascontiguousarray(ntot, dtype=np.int64)
    if pmax is None:
        pmax = n
    if sel is None:
        sel = np.zeros((0, n))
    sel = np.ascontiguousarray(sel, dtype=np.float64).reshape(-1, n)
    nsel = sel.shape[0]
    if coef is not None:
        coef = np.ascontiguousarray(coef, dtype=np.float64).reshape(-1, n)
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        ncoef = coef.shape[0]
        cp, xp = _d(coef), _d(xi)
    else:
        ncoef, cp, xp = 0, None, None
    work = np.zeros(n)
    iwork = np.zeros(2 * (n + 2), dtype=np.int32)
    fe = np.zeros(pmax)
    bounds = np.zeros((pmax, 2), dtype=np.int32)
    M = np.zeros(pmax, dtype=np.int32)
    m = np.zeros(pmax + 1, dtype=np.int32)
    nph = ctypes.c_int(0)
    nmn = ctypes.c_int(0)
    safe = ctypes.c_int(0)
    info = ctypes.c_int(0)
    avg_c = np.zeros((pmax, nsel)) if nsel else np.zeros((pmax, 0))
    st = L.fo_state_point(_d(lnpi_ref), nt.ctypes.data_as(_lp), n, float(beta_ref), float(mu1_ref), float(mu1),
                          cp, ncoef, xp, int(smooth), float(cutoff), _d(sel) if nsel else None, nsel, pmax,
                          _d(work), _i(iwork), ctypes.byref(nph), ctypes.byref(nmn), _d(fe), _d(avg_c) if nsel else None, _i(bounds),
                          _i(M), _i(m), ctypes.byref(safe), ctypes.byref(info))
    P = nph.value if st == 0 else 0
    nmin = nmn.value if st == 0 else 0
    return {"status": st, "nphase": nph.value, "fe": fe[:P].copy(), "avg": avg_c[:P].copy(),
            "bounds": bounds[:P].copy(), "max_idx": M[:P].copy(), "min_idx": m[:nmin].copy(),
            "safe": bool(safe.value), "lnpi": work, "info": info.value}


def phase_eq_err2(fe, bounds, min_width):
    """gc_hist.pyx:2614-2630."""
    fe = np.ascontiguousarray(fe, dtype=np.float64)
    b = np.ascontiguousarray(bounds, dtype=np.int32)
    return lib().fo_phase_eq_err2(_d(fe), _i(b), len(fe), int(min_width))


def find_phase_eq_fmin(lnpi_ref, ntot, beta_ref, mu1_ref, smooth, lnZ_tol, mu_guess, coef_fn=None):
    """Reference-faithful solver: scipy Nelder-Mead on the squared error, exactly as GH:653
    (``fmin(phase_eq_error, mu_guess, ftol=lnZ_tol, maxfun=maxiter=100000)``)."""
    from scipy.optimize import fmin
    min_width = 2 * smooth

    def err(mu):
        mu = float(np.atleast_1d(mu)[0])
        coef, xi = coef_fn(mu) if coef_fn else (None, None)
        r = state_point(lnpi_ref, ntot, beta_ref, mu1_ref, mu, smooth, coef=coef, xi=xi)
        if r["status"] != 0:
            raise RuntimeError("oracle state point failed: %s" % STATUS[r["status"]])
        return phase_eq_err2(r["fe"], r["bounds"], min_width)

    out = fmin(err, mu_guess, ftol=lnZ_tol, maxfun=100000, maxiter=100000, full_output=True, disp=False)
    return float(out[0][0]), float(out[1]), int(out[3])


def signed_dfe(lnpi_ref, ntot, beta_ref, mu1_ref, smooth, mu, coef_fn=None):
    """Signed free-energy difference between the two widest-valid phases closest in F.E. (the pair
    GH:2614-2630 selects), or None when fewer than two wide phases exist."""
    coef, xi = coef_fn(mu) if coef_fn else (None, None)
    r = state_point(lnpi_ref, ntot, beta_ref, mu1_ref, mu, smooth, coef=coef, xi=xi)
    if r["status"] != 0:
        return None, r
    w = 2 * smooth
    best = None
    for i in range(r["nphase"]):
        if r["bounds"][i, 1] - r["bounds"][i, 0] < w:
            continue
        for j in range(i + 1, r["nphase"]):
            if r["bounds"][j, 1] - r["bounds"][j, 0] < w:
                continue
            d = r["fe"][i] - r["fe"][j]
            if best is None or d * d < best[0]:
                best = (d * d, d, i, j)
    if best is None:
        return None, r
    return best[1], r


def find_phase_eq_tight(lnpi_ref, ntot, beta_ref, mu1_ref, smooth, mu_lo, mu_hi, coef_fn=None, xtol=1e-14):
    """'Tightened oracle' of SURVEY.md section 7 hard part 3: same objective, root of the SIGNED
    dF.E. by brentq inside a bracket on which the same two phases exist."""
    from scipy.optimize import brentq

    def f(mu):
        d, _ = signed_dfe(lnpi_ref, ntot, beta_ref, mu1_ref, smooth, mu, coef_fn)
        if d is None:
            raise RuntimeError("fewer than two wide phases at mu=%r" % mu)
        return d

    return brentq(f, mu_lo, mu_hi, xtol=xtol, rtol=8.9e-16, maxiter=500)


def taylor_coefficients(mom, d0=0.0):
    """Closed-form Taylor coefficient arrays (derived from reference
    gc_hist.pyx:_sg_dX_dB 1660, _sg_dX_dMU 1724, _sg_d2X_dB2 1776, _sg_d2X_dMU2 1829, _sg_df_dB 1914
    after dropping the N-independent _gc_* constants that renormalisation removes).

    Returns dict of N-length arrays such that (xi_b = beta-beta_ref, xi_d = dmu2 - dmu2_ref):
        lnPI' = lnPI + beta_ref*(mu1-mu1_ref)*N
                + xi_b*(mu1*N + A_b) + xi_d*beta_ref*A_d
                + 0.5*xi_b^2*A_bb + xi_b*xi_d*(A_d + beta_ref*A_bd) + 0.5*xi_d^2*beta_ref^2*A_dd
    """
    mom = np.asarray(mom, dtype=np.float64)
    nspec = mom.shape[0]
    U = mom[0, 0, 0, 0, 1]
    out = {}
    if mom.shape[1] > 2:
        f_UU = mom[0, 0, 0, 0, 2] - U * U
    else:
        f_UU = None
    if nspec == 1:
        out["A_b"] = -U
        out["A_bb"] = f_UU
        return out
    N2 = mom[1, 1, 0, 0, 0]
    out["A_b"] = -U + d0 * N2
    out["A_d"] = N2
    if mom.shape[1] > 2:
        f_N2U = mom[1, 1, 0, 0, 1] - N2 * U
        f_N2N2 = mom[1, 2, 0, 0, 0] - N2 * N2
        g = -f_N2U + d0 * f_N2N2
        out["A_bb"] = d0 * g - (-f_UU + d0 * f_N2U)
        out["A_bd"] = g
        out["A_dd"] = f_N2N2
    return out


def mix(x_self, x_other, weights):
    """gc_hist.pyx:244-252: pointwise blend over the common range, longer one supplies the tail."""
    a = np.asarray(x_self, dtype=np.float64)
    b = np.asarray(x_other, dtype=np.float64)
    longer, n = (a, b.shape[-1]) if a.shape[-1] >= b.shape[-1] else (b, a.shape[-1])
    out = np.array(longer, dtype=np.float64)
    out[..., :n] = (a[..., :n] * weights[0] + weights[1] * b[..., :n]) / (weights[0] + weights[1])
    return out


def reweight_2d(lnpi, bounds, op1, op2, b_dmu1, b_dmu2, props=None):
    """2-D joint (op1, op2) reweight + logsumexp + averages (new capability; see fhmc_oracle.c)."""
    lnpi = np.ascontiguousarray(lnpi, dtype=np.float64)
    n1, n2 = lnpi.shape
    b = np.ascontiguousarray(bounds, dtype=np.int32)
    o1 = np.ascontiguousarray(op1, dtype=np.float64)
    o2 = np.ascontiguousarray(op2, dtype=np.float64)
    if props is None:
        props = np.zeros((0, n1, n2))
    props = np.ascontiguousarray(props, dtype=np.float64)
    out = np.zeros(3 + props.shape[0])
    lib().fo_reweight_2d(_d(lnpi), _i(b), n1, n2, _d(o1), _d(o2), float(b_dmu1), float(b_dmu2),
                         _d(props) if props.shape[0] else None, props.shape[0], _d(out))
    return out


# ----------------------------------------------------------------------------------------------
# pore_hist (moments/histogram/two_dim/h_ntot/pore_hist.pyx) -- SURVEY 8(f) row 3
# ----------------------------------------------------------------------------------------------
def _spec_exp(a, b):
    """ln(exp(a) + exp(b)) as pore_hist.pyx:35-53: max(a, b) + log(1 + exp(-|a - b|))."""
    return max(a, b) + math.log(1.0 + math.exp(-abs(a - b)))


def pore_normalize(lnpi, edge):
    """pore_hist._cy_normalize (pore_hist.pyx:57-80): sequential specExp fold, row-major over j <= edge[i], starting
    from -DBL_MAX; returns data - lnNormPI.  Pinned on the compiled reference (tests/golden/pore_vectors.npz)."""
    lnpi = np.asarray(lnpi, dtype=np.float64)
    acc = -sys.float_info.max
    for i in range(lnpi.shape[0]):
        for j in range(0, int(edge[i]) + 1):
            acc = _spec_exp(acc, float(lnpi[i, j]))
    return lnpi - acc


def pore_thermo(lnpi, mask, props):
    """pore_hist.thermo (pore_hist.pyx:154-184) with ``lp[not mask]`` read as ``lp[~mask]`` (the reference raises on any
    mask of more than one element -- PARITY UNPINNED for this function).
    Returns ({name: average}, peak_idx)."""
    lp = np.array(lnpi, dtype=np.float64, copy=True)
    mask = np.asarray(mask, dtype=bool)
    with np.errstate(all="ignore"):
        lp -= np.max(lp)                       # PH:169
        lp[~mask] = -np.inf                    # PH:170
        lp -= np.log(np.sum(np.exp(lp)))       # PH:171
        lp[~mask] = -np.inf                    # PH:172
        prob = np.exp(lp)                      # PH:174
        sum_prob = np.sum(prob)                # PH:175
        ave = {p: np.sum(prob * np.where(mask, props[p], 0.0)) / sum_prob for p in props}   # PH:178-179 (0 * x outside the mask)
    return ave, np.where(lp == np.max(lp))     # PH:182


# ----------------------------------------------------------------------------------------------
# window patching shift solve (moments/win_patch/fhmc_patch.pyx:640-709) -- SURVEY 8(f) row 4
# ----------------------------------------------------------------------------------------------
def window_patch_error(x, this_lnpi, other_lnpi):
    """fhmc_patch.pyx:640-664: total square error of (this + x) against other, accumulated in index order."""
    e2 = 0.0
    for i in range(len(this_lnpi)):
        e2 += ((this_lnpi[i] + x) - other_lnpi[i]) ** 2
    return e2


def patch_window_pair_slices(data_slice1, data_slice2, ftol=0.000001):
    """The optimisation of patch_window_pair (fhmc_patch.pyx:699-709) on already aligned overlap slices: Nelder-Mead from
    the first-point guess, (x*, error(x*) / len).  The cdef function cannot be called from Python and the module needs the
    simulation's file formats to build windows, so this is a restatement with the same scipy.optimize.fmin call
    (PARITY UNPINNED on the reference's own objects; the minimiser of a parabola is also known in closed form)."""
    from scipy.optimize import fmin
    a = np.asarray(data_slice1, dtype=np.float64)
    b = np.asarray(data_slice2, dtype=np.float64)
    guess = b[0] - a[0]
    full_out = fmin(window_patch_error, guess, ftol=ftol, args=(a, b), maxiter=10000, maxfun=10000, full_output=True, disp=False)
    if full_out[4] != 0:
        raise Exception("Error, unable to mimize")
    return float(full_out[0][0]), float(full_out[1]) / len(a)

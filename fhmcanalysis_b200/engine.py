"""Host-side engine: packs histograms into device blobs and drives the sm_100a kernels through the
C ABI (include/fhmc_b200.h).  PyTorch is used ONLY for device memory, streams and host<->device
copies; all arithmetic happens in libfhmc_b200.so.  No CPU fallback: without a CUDA device or the
built library every entry point raises.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (M_DB, M_DB2, M_DB3, M_DB_MU1, M_DBDD, M_DD, M_DD2, M_ONE, MAX_SEL, MAX_TERMS,  # noqa: F401
                   ST_CODE_MASK, ST_SAFE)

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda(device=None):
    """Return a torch.device for the GPU to use, or raise (the product has no CPU path)."""
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError("fhmcanalysis_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.load()
    if device is None:
        return t.device("cuda", t.cuda.current_device())
    return t.device(device)


def _ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else None


def _stream_ptr(device):
    return ctypes.c_void_p(torch().cuda.current_stream(device).cuda_stream)


def upper_hull(x, y):
    """Upper concave envelope of the points (x_i, y_i), x strictly increasing and finite y: returns
    (edge slopes, strictly decreasing; vertex indices) or None when the fast mu-sweep path cannot use it.
    One-time O(n) host setup per histogram (Andrew's monotone chain)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    if n < 3 or not np.all(np.diff(x) > 0) or not np.all(np.isfinite(y)):
        return None
    h = []
    for i in range(n):
        while len(h) >= 2:
            a, b = h[-2], h[-1]
            # drop b unless slope(a,b) > slope(b,i)
            if (y[b] - y[a]) * (x[i] - x[b]) <= (y[i] - y[b]) * (x[b] - x[a]):
                h.pop()
            else:
                break
        h.append(i)
    verts = np.array(h, dtype=np.float64)
    slopes = np.array([(y[h[k + 1]] - y[h[k]]) / (x[h[k + 1]] - x[h[k]]) for k in range(len(h) - 1)])
    if len(slopes) > 1 and not np.all(np.diff(slopes) < 0):
        return None
    return slopes, verts


class SweepResult(object):
    """Per-state-point records of one sweep (device tensors; ``host()`` copies them to NumPy)."""

    FIELDS = ("status", "nphase", "nmin", "lnnorm", "fe", "avg", "bounds", "max_idx", "min_idx")

    def __init__(self, n_states, pmax, n_sel, device, pinned=False):
        t = torch()
        S = int(n_states)
        self.n_states, self.pmax, self.n_sel, self.device = S, int(pmax), int(n_sel), device
        kw = dict(device=device)
        self.status = t.empty(S, dtype=t.int32, **kw)   # bit pattern of the unsigned status word
        self.nphase = t.empty(S, dtype=t.int32, **kw)
        self.nmin = t.empty(S, dtype=t.int32, **kw)
        self.lnnorm = t.empty(S, dtype=t.float64, **kw)
        self.fe = t.empty((S, pmax), dtype=t.float64, **kw)
        self.avg = t.empty((S, pmax, max(n_sel, 1)), dtype=t.float64, **kw) if n_sel else None
        self.bounds = t.empty((S, pmax, 2), dtype=t.int32, **kw)
        self.max_idx = t.empty((S, pmax), dtype=t.int32, **kw)
        self.min_idx = t.empty((S, pmax + 1), dtype=t.int32, **kw)
        self.extra = {}

    def c_struct(self):
        return _lib.SweepOut(*[_ptr(getattr(self, k)) for k in self.FIELDS])

    def nbytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size()
                   for k in self.FIELDS if getattr(self, k) is not None)

    def host(self):
        out = {}
        for k in self.FIELDS:
            v = getattr(self, k)
            out[k] = v.cpu().numpy() if v is not None else None
        out["status"] = out["status"].view(np.uint32)
        out["code"] = (out["status"] & ST_CODE_MASK).astype(np.int32)
        out["safe"] = (out["status"] & ST_SAFE) != 0
        for k, v in self.extra.items():
            out[k] = v.cpu().numpy()
        return out


class DeviceHistogram(object):
    """One histogram resident in HBM as the row blob the kernels stage in shared memory.

    rows: row 0 = ln(PI), row 1 = N_tot (fp64), then Taylor coefficient rows, then the rows of the
    quantities to average (``n_term`` rows each).  See include/fhmc_b200.h for the exact contract.

    Parameters
    ----------
    lnpi, ntot : arrays [n]
    beta_ref, mu1_ref, dmu_ref : conditions the stored histogram is at (data['curr_*'] in the reference)
    smooth, cutoff : extrema window / is_safe cutoff
    coef : list of (kind, row) where row is an [n] array, or the string "N" for the N_tot row
    sel : list of quantities; each is an [n] array (no extrapolation) or a list of n_term arrays
    sel_kinds : monomial kinds of the extra terms of every quantity (len n_term-1)
    """

    def __init__(self, lnpi, ntot, beta_ref, mu1_ref, dmu_ref=0.0, smooth=1, cutoff=10.0, coef=(), sel=(),
                 sel_kinds=(), device=None):
        t = torch()
        self.device = require_cuda(device)
        lnpi = np.ascontiguousarray(lnpi, dtype=np.float64)
        n = lnpi.shape[0]
        ntot = np.ascontiguousarray(ntot, dtype=np.float64)
        if ntot.shape != (n,):
            raise ValueError("ntot must have the same length as ln(PI)")
        n_pad = n + (n & 1)
        rows = [lnpi, ntot]
        coef_row, coef_kind = [], []
        for kind, row in coef:
            if isinstance(row, str) and row == "N":
                coef_row.append(1)
            else:
                coef_row.append(len(rows))
                rows.append(np.ascontiguousarray(row, dtype=np.float64))
            coef_kind.append(int(kind))
        if len(coef_row) > MAX_TERMS:
            raise ValueError("at most %d Taylor terms" % MAX_TERMS)
        n_term = 1 + len(sel_kinds)
        if n_term > MAX_TERMS:
            raise ValueError("at most %d terms per quantity" % MAX_TERMS)
        if len(sel) > MAX_SEL:
            raise ValueError("the fused sweep averages at most %d quantities; use phase_moments for more" % MAX_SEL)
        sel_row = []
        for q in sel:
            terms = [q] if n_term == 1 and not isinstance(q, (list, tuple)) else list(q)
            if len(terms) != n_term:
                raise ValueError("every quantity needs %d rows" % n_term)
            if n_term == 1 and isinstance(terms[0], str) and terms[0] == "N":
                sel_row.append(1)
                continue
            sel_row.append(len(rows))
            rows.extend(np.ascontiguousarray(r, dtype=np.float64) for r in terms)
        hull_row, hull_len = 0, 0
        if not coef:
            hull = upper_hull(ntot, lnpi)
            if hull is not None:
                slopes, verts = hull
                hull_row, hull_len = len(rows), len(verts)
                srow, vrow = np.zeros(n), np.zeros(n)
                srow[:len(slopes)] = slopes
                vrow[:len(verts)] = verts
                rows.extend([srow, vrow])
        blob = np.zeros((len(rows), n_pad), dtype=np.float64)
        for i, r in enumerate(rows):
            if r.shape != (n,):
                raise ValueError("row %d has shape %r, expected (%d,)" % (i, r.shape, n))
            blob[i, :n] = r
        self.n, self.n_pad, self.n_rows = n, n_pad, len(rows)
        self.blob_host = blob
        self.blob = t.from_numpy(blob).to(self.device)
        self.h2d_bytes = blob.nbytes
        self.n_sel, self.n_term = len(sel_row), n_term
        d = _lib.HistDesc()
        d.n, d.n_pad, d.n_rows = n, n_pad, len(rows)
        d.n_coef = len(coef_row)
        for i in range(len(coef_row)):
            d.coef_row[i], d.coef_kind[i] = coef_row[i], coef_kind[i]
        d.n_sel, d.n_term = len(sel_row), n_term
        for i, r in enumerate(sel_row):
            d.sel_row[i] = r
        d.sel_kind[0] = M_ONE
        for i, k in enumerate(sel_kinds):
            d.sel_kind[1 + i] = int(k)
        d.smooth, d.pmax, d.complete, d.compare_raw = int(smooth), 4, 0, 0
        d.cutoff, d.beta_ref, d.mu1_ref, d.dmu_ref = float(cutoff), float(beta_ref), float(mu1_ref), float(dmu_ref)
        d.hull_row, d.hull_len = hull_row, hull_len
        self.desc = d

    # ------------------------------------------------------------------------------------------
    def _desc(self, pmax, complete=False, compare_raw=False, cutoff=None, smooth=None):
        d = _lib.HistDesc.from_buffer_copy(self.desc)
        d.pmax = int(pmax)
        d.complete = 1 if complete else 0
        d.compare_raw = 1 if compare_raw else 0
        if cutoff is not None:
            d.cutoff = float(cutoff)
        if smooth is not None:
            d.smooth = int(smooth)
        return d

    def _dev_array(self, x):
        """1-D fp64 device tensor from a NumPy array / scalar / tensor (copies host data)."""
        t = torch()
        if x is None:
            return None
        if isinstance(x, t.Tensor):
            if x.dtype != t.float64 or not x.is_contiguous():
                x = x.to(t.float64).contiguous()
            return x.to(self.device)
        return t.from_numpy(np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)).to(self.device)

    def make_states(self, mu1, beta=None, dmu=None, grid=False):
        """Build the fhmc_states descriptor.  grid=False: flat lists (length-1 arrays broadcast).
        grid=True: state point = (mu1[i], beta[j], dmu[k]) with dmu fastest (temp_dmu_extrap_multi order)."""
        mu1_t, beta_t, dmu_t = self._dev_array(mu1), self._dev_array(beta), self._dev_array(dmu)
        st = _lib.States()
        nm = mu1_t.numel()
        nb = beta_t.numel() if beta_t is not None else 1
        nd = dmu_t.numel() if dmu_t is not None else 1
        if grid:
            S = nm * nb * nd
            st.mu1_div, st.beta_div, st.dmu_div = nb * nd, nd, 1
        else:
            S = max(nm, nb, nd)
            for k in (nm, nb, nd):
                if k not in (1, S):
                    raise ValueError("flat state lists must have equal length (or length 1)")
            st.mu1_div = st.beta_div = st.dmu_div = 1
        st.n_states = S
        st.mu1, st.n_mu1 = _ptr(mu1_t), nm
        st.beta, st.n_beta = _ptr(beta_t), nb
        st.dmu, st.n_dmu = _ptr(dmu_t), nd
        st._keep = (mu1_t, beta_t, dmu_t)  # keep the tensors alive
        return st

    def sweep(self, mu1, beta=None, dmu=None, grid=False, pmax=4, lanes=0, complete=False, compare_raw=False,
              cutoff=None, smooth=None, out=None, states=None):
        """K1+K3+K2 over all state points; returns a SweepResult of device tensors (asynchronous)."""
        L = _lib.load()
        st = states if states is not None else self.make_states(mu1, beta, dmu, grid)
        d = self._desc(pmax, complete, compare_raw, cutoff, smooth)
        if out is None:
            out = SweepResult(st.n_states, pmax, self.n_sel, self.device)
        elif out.n_states != st.n_states or out.pmax != pmax or out.n_sel != self.n_sel:
            raise ValueError("output buffers do not match the sweep")
        cs = out.c_struct()
        with torch().cuda.device(self.device):
            rc = L.fhmc_sweep_1d(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), ctypes.byref(cs), int(lanes),
                                 _stream_ptr(self.device))
        _lib.check(rc, "fhmc_sweep_1d")
        out._states = st
        return out

    def sweep_auto(self, mu1, beta=None, dmu=None, grid=False, pmax=4, **kw):
        """sweep() that grows pmax until no state point reports FHMC_E_CAPACITY (synchronises)."""
        while True:
            res = self.sweep(mu1, beta, dmu, grid, pmax=pmax, **kw)
            code = (res.status & ST_CODE_MASK)
            if not bool((code == _lib.E_CAPACITY).any().item()) or pmax > self.n:
                return res
            pmax = min(pmax * 4, self.n + 1)

    def lnpi_rows(self, result, states=None):
        """Normalised reweighted ln(PI) of every state point of ``result``: tensor [S][n]."""
        L = _lib.load()
        t = torch()
        st = states if states is not None else result._states
        out = t.empty((st.n_states, self.n), dtype=t.float64, device=self.device)
        d = self._desc(result.pmax)
        with t.cuda.device(self.device):
            rc = L.fhmc_lnpi_1d(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), _ptr(result.lnnorm), _ptr(out),
                                _stream_ptr(self.device))
        _lib.check(rc, "fhmc_lnpi_1d")
        return out

    def find_phase_eq(self, mu_guess, beta=None, dmu=None, lnz_tol=1e-10, mu_step=None, max_iter=200, pmax=4,
                      smooth=None, cutoff=None):
        """K4: one coexistence solve per entry of (mu_guess, beta, dmu) (flat lists).  Returns
        (SweepResult at coexistence with extra['mu_coex','dfe','iters'])."""
        L = _lib.load()
        t = torch()
        if self.n_sel < 1:
            raise ValueError("the solver needs quantity 0 to be N_tot (construct DeviceHistogram with sel=['N', ...])")
        st = self.make_states(mu_guess, beta, dmu, grid=False)
        d = self._desc(max(pmax, 2), False, False, cutoff, smooth)
        out = SweepResult(st.n_states, d.pmax, self.n_sel, self.device)
        T = st.n_states
        mu_coex = t.empty(T, dtype=t.float64, device=self.device)
        dfe = t.empty(T, dtype=t.float64, device=self.device)
        iters = t.empty(T, dtype=t.int32, device=self.device)
        if mu_step is None:
            mu_step = 1.0 / abs(self.desc.beta_ref)
        cs = out.c_struct()
        with t.cuda.device(self.device):
            rc = L.fhmc_find_phase_eq_1d(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), float(lnz_tol),
                                         float(mu_step), int(max_iter), _ptr(mu_coex), _ptr(dfe), _ptr(iters),
                                         ctypes.byref(cs), _stream_ptr(self.device))
        _lib.check(rc, "fhmc_find_phase_eq_1d")
        out.extra = {"mu_coex": mu_coex, "dfe": dfe, "iters": iters}
        out._states = st
        return out


def phase_moments(lnpi, mom, bounds, device=None):
    """K2 (drop-in thermo): phase averages of every row of ``mom`` [A][n] and the per-phase ln-sums.
    lnpi: normalised ln(PI) [n]; bounds [P][2].  Returns (avg [P][A], lnsum [P]) as NumPy arrays."""
    L = _lib.load()
    t = torch()
    dev = require_cuda(device)
    lnpi_t = lnpi if isinstance(lnpi, t.Tensor) else t.from_numpy(np.ascontiguousarray(lnpi, dtype=np.float64)).to(dev)
    n = lnpi_t.numel()
    b = t.from_numpy(np.ascontiguousarray(bounds, dtype=np.int32).reshape(-1, 2)).to(dev)
    P = b.shape[0]
    A = 0
    mom_t = avg = None
    if mom is not None:
        mom_t = mom if isinstance(mom, t.Tensor) else t.from_numpy(np.ascontiguousarray(mom, dtype=np.float64).reshape(-1, n)).to(dev)
        A = mom_t.shape[0]
        avg = t.empty((P, A), dtype=t.float64, device=dev)
    lnsum = t.empty(P, dtype=t.float64, device=dev)
    with t.cuda.device(dev):
        rc = L.fhmc_phase_moments(_ptr(lnpi_t), n, _ptr(mom_t), A, _ptr(b), P, _ptr(avg), _ptr(lnsum), _stream_ptr(dev))
    _lib.check(rc, "fhmc_phase_moments")
    return (avg.cpu().numpy() if avg is not None else np.zeros((P, 0))), lnsum.cpu().numpy()


def axpy_rows(arrays, weights, device=None):
    """out = sum_t weights[t] * arrays[t] on the device (moment extrapolation / mix).  NumPy in/out."""
    L = _lib.load()
    t = torch()
    dev = require_cuda(device)
    k = len(arrays)
    if k < 1 or k > MAX_TERMS or len(weights) != k:
        raise ValueError("need 1..%d arrays and as many weights" % MAX_TERMS)
    shape = np.shape(arrays[0])
    ts = [t.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(dev) for a in arrays]
    count = ts[0].numel()
    for x in ts:
        if x.numel() != count:
            raise ValueError("arrays must have equal size")
    out = t.empty(count, dtype=t.float64, device=dev)
    ptrs = (ctypes.c_void_p * k)(*[x.data_ptr() for x in ts])
    w = (ctypes.c_double * k)(*[float(x) for x in weights])
    with t.cuda.device(dev):
        rc = L.fhmc_axpy_rows(ptrs, w, k, count, _ptr(out), _stream_ptr(dev))
    _lib.check(rc, "fhmc_axpy_rows")
    return out.cpu().numpy().reshape(shape)


def reweight_2d(lnpi, bounds, op1, op2, a1, a2, props=None, device=None, return_device=False):
    """K5: 2-D joint histogram reweight for S state points; returns [S][3+n_prop] (lnZ, <op1>, <op2>, <prop>...)."""
    L = _lib.load()
    t = torch()
    dev = require_cuda(device)

    def dv(x, dtype):
        if isinstance(x, t.Tensor):
            return x.to(dev)
        return t.from_numpy(np.ascontiguousarray(x, dtype=dtype)).to(dev)

    lnpi_t = dv(lnpi, np.float64)
    n1, n2 = lnpi_t.shape
    b_t = dv(bounds, np.int32)
    o1, o2 = dv(op1, np.float64), dv(op2, np.float64)
    a1_t, a2_t = dv(np.atleast_1d(a1) if not isinstance(a1, t.Tensor) else a1, np.float64), \
        dv(np.atleast_1d(a2) if not isinstance(a2, t.Tensor) else a2, np.float64)
    S = a1_t.numel()
    n_prop = 0
    p_t = None
    if props is not None and len(props):
        p_t = dv(props, np.float64)
        n_prop = p_t.shape[0]
    out = t.empty((S, 3 + n_prop), dtype=t.float64, device=dev)
    ws_bytes = L.fhmc_reweight_2d_workspace(n1, n2, n_prop, S)
    ws = t.empty(max(ws_bytes // 8, 1), dtype=t.float64, device=dev)
    with t.cuda.device(dev):
        rc = L.fhmc_reweight_2d(_ptr(lnpi_t), _ptr(b_t), n1, n2, _ptr(o1), _ptr(o2), _ptr(p_t), n_prop, _ptr(a1_t),
                                _ptr(a2_t), S, _ptr(out), _ptr(ws), ws_bytes, _stream_ptr(dev))
    _lib.check(rc, "fhmc_reweight_2d")
    return out if return_device else out.cpu().numpy()


def measure_peaks(device=None, iters=20000):
    """Register-resident fp64 micro-benchmarks: returns dict(dfma_per_s, exp_per_s) measured with CUDA events."""
    L = _lib.load()
    t = torch()
    dev = require_cuda(device)
    sink = t.zeros(4, dtype=t.float64, device=dev)
    res = {}
    with t.cuda.device(dev):
        for name, fn in (("dfma_per_s", L.fhmc_bench_dfma), ("exp_per_s", L.fhmc_bench_exp)):
            best = 0.0
            for rep in range(4):
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record()
                ops = fn(int(iters), _ptr(sink), _stream_ptr(dev))
                e1.record()
                e1.synchronize()
                if ops < 0:
                    raise RuntimeError("micro-benchmark launch failed")
                if rep:
                    best = max(best, ops / (e0.elapsed_time(e1) * 1e-3))
            res[name] = best
    return res

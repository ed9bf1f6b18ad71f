"""Host-side engine: packs histograms into device blobs and drives the sm_100a kernels through the
C ABI (include/fhmc_b200.h).  PyTorch is used ONLY for device memory, streams and host<->device
copies; all arithmetic happens in libfhmc_b200.so.  No CPU fallback: without a CUDA device or the
built library every entry point raises.
"""
import ctypes
import os

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
    One-time host setup per histogram (qhull through scipy when available, else Andrew's monotone chain)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    if n < 3 or not np.all(np.diff(x) > 0) or not np.all(np.isfinite(y)):
        return None
    h = None
    try:
        from scipy.spatial import ConvexHull
        v = np.sort(ConvexHull(np.column_stack([x, y])).vertices)
        # keep the vertices on or above the chord from the first to the last point
        chord = y[0] + (y[-1] - y[0]) * (x[v] - x[0]) / (x[-1] - x[0])
        h = [int(i) for i in v[(y[v] >= chord) | (v == 0) | (v == n - 1)]]
    except Exception:
        h = None
    if h is None or len(h) < 2:
        h = []
        for i in range(n):
            while len(h) >= 2:
                a, b = h[-2], h[-1]
                if (y[b] - y[a]) * (x[i] - x[b]) <= (y[i] - y[b]) * (x[b] - x[a]):
                    h.pop()
                else:
                    break
            h.append(i)
    h = np.asarray(h, dtype=np.int64)
    slopes = np.diff(y[h]) / np.diff(x[h])
    while len(slopes) > 1 and not np.all(np.diff(slopes) < 0):   # drop collinear / numerically reflex vertices
        keep = np.concatenate([[True], np.diff(slopes) < 0, [True]])
        h = h[keep]
        slopes = np.diff(y[h]) / np.diff(x[h])
    return slopes, h.astype(np.float64)


class SweepResult(object):
    """Per-state-point records of one sweep.  All fields live in ONE device allocation (typed views), so the whole
    result comes back with a single device-to-host copy (``host()``)."""

    FIELDS = ("status", "nphase", "nmin", "lnnorm", "fe", "avg", "bounds", "max_idx", "min_idx")

    @staticmethod
    def layout(S, pmax, n_sel):
        spec = [("lnnorm", np.float64, (S,)), ("fe", np.float64, (S, pmax)), ("avg", np.float64, (S, pmax, n_sel)),
                ("status", np.int32, (S,)), ("nphase", np.int32, (S,)), ("nmin", np.int32, (S,)),
                ("bounds", np.int32, (S, pmax, 2)), ("max_idx", np.int32, (S, pmax)), ("min_idx", np.int32, (S, pmax + 1))]
        off, out = 0, []
        for name, dt, shape in spec:
            nb = int(np.prod(shape)) * np.dtype(dt).itemsize
            if name == "avg" and n_sel == 0:
                continue
            out.append((name, dt, shape, off, nb))
            off += (nb + 15) & ~15
        return out, max(off, 16)

    def __init__(self, n_states, pmax, n_sel, device, pinned=False):
        t = torch()
        S = int(n_states)
        self.n_states, self.pmax, self.n_sel, self.device = S, int(pmax), int(n_sel), device
        self._layout, total = self.layout(S, self.pmax, self.n_sel)
        self.buf = t.empty(total, dtype=t.uint8, device=device)
        self.avg = None
        tdt = {np.float64: t.float64, np.int32: t.int32}
        for name, dt, shape, off, nb in self._layout:
            setattr(self, name, self.buf[off:off + nb].view(tdt[dt]).view(shape))
        self.extra = {}

    def c_struct(self):
        return _lib.SweepOut(*[_ptr(getattr(self, k)) for k in self.FIELDS])

    def nbytes(self):
        return sum(nb for _, _, _, _, nb in self._layout)

    def host(self):
        hb = self.buf.cpu().numpy()
        out = {"avg": None}
        for name, dt, shape, off, nb in self._layout:
            out[name] = hb[off:off + nb].view(dt).reshape(shape)
        out["status"] = out["status"].view(np.uint32)
        out["code"] = (out["status"] & ST_CODE_MASK).astype(np.int32)
        out["safe"] = (out["status"] & ST_SAFE) != 0
        for k, v in self.extra.items():
            out[k] = v.cpu().numpy() if hasattr(v, "cpu") else v
        return out


def save_results(fname, host, attrs=None):
    """Write the dict of per-state-point arrays returned by ``SweepResult.host()`` / ``sweep_host`` /
    ``reweight_batch`` as variables of one HDF5 file (``io.hdf5_min.write_hdf5``); ``load_results`` reads it back."""
    from .io.hdf5_min import write_hdf5
    v = {}
    for k, a in host.items():
        if a is None:
            continue
        a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
        if a.dtype.kind in "fiub":
            v[k] = a
    write_hdf5(fname, v, dict(attrs or {}, producer="fhmcanalysis_b200"))


def load_results(fname):
    from .io.hdf5_min import File
    f = File(fname)
    out = {k: v.read() for k, v in f.variables.items()}
    if "safe" in out:
        out["safe"] = out["safe"].astype(bool)
    return out, dict(f.attrs)


def soa16_views(buf, S, pmax, nsel):
    """[S, ...] views of a uint8 tensor (host or device) laid out as fhmc_pack_phase_soa16 describes for S state points:
    status (int16), nphase (uint8), fe [S, pmax], avg [S, pmax, nsel], bounds [S, pmax, 2] (int16); phase-major underneath."""
    t = torch()
    nf = 1 + nsel
    out = {}
    out["status"] = t.as_strided(buf[:4 * S].view(t.int16), (S,), (2,))
    out["nphase"] = t.as_strided(buf[:4 * S], (S,), (4,), buf.storage_offset() + 2)
    out["path"] = t.as_strided(buf[:4 * S], (S,), (4,), buf.storage_offset() + 3)   # diagnostic: 1 = written by the tilt cells
    f0 = (4 * S + 15) & ~15
    F = buf[f0:f0 + pmax * S * nf * 8].view(t.float64)
    B = buf[f0 + pmax * S * nf * 8:f0 + pmax * S * nf * 8 + pmax * S * 4].view(t.int16)
    out["fe"] = t.as_strided(F, (S, pmax), (nf, S * nf), F.storage_offset())
    out["avg"] = t.as_strided(F, (S, pmax, nsel), (nf, S * nf, 1), F.storage_offset() + 1) if nsel else None
    out["bounds"] = t.as_strided(B, (S, pmax, 2), (2, 2 * S, 1), B.storage_offset())
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
        hull_row, hull_len = 0, 0          # the hull rows are added lazily (ensure_hull) by large mu sweeps only
        self._hull_possible = not coef
        # exp strength reduction of large pure-mu sweeps, fixed before the first one: 3 = product form (Horner blocks,
        # fhmc_fast_prod.cu) with two state points per thread where the sweep is large enough (> 512 points per SM), 2 = product form, one
        # point per thread, 1 = four multiplicative chains (fhmc_fast_rec.cu), 0/False = one true exp per bin
        self.use_recurrence = int(os.environ.get("FHMC_MU_RECURRENCE", "3"))
        # per-histogram tables for compact-record mu sweeps (ensure_mu_tables); False keeps the table-free kernel
        self.use_mu_tables = os.environ.get("FHMC_MU_TABLES", "1") != "0"
        # tilt cells on top of the tables for dense sweeps (ensure_mu_cells); False keeps the table walk
        self.use_mu_cells = os.environ.get("FHMC_MU_CELLS", "1") != "0"
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

    FAST_PATH_MIN_STATES = 4096   # = FHMC_FAST_MIN_STATES: below this the generic multi-lane kernels win and no hull is needed

    def ensure_hull(self):
        """Append the two hull rows the one-pass mu-sweep kernel needs (upper concave envelope of (N, lnPI)) and
        re-upload the blob.  Done once, on the first large pure-mu sweep."""
        if self.desc.hull_len or not self._hull_possible:
            return
        self._hull_possible = False
        hull = upper_hull(self.blob_host[1, :self.n], self.blob_host[0, :self.n])
        if hull is None:
            return
        slopes, verts = hull
        extra = np.zeros((2, self.n_pad), dtype=np.float64)
        extra[0, :len(slopes)] = slopes
        extra[1, :len(verts)] = verts
        self.desc.hull_row, self.desc.hull_len = self.n_rows, len(verts)
        # exp recurrence of the one-pass kernel: uniform N spacing and moderate 4-bin steps of lnPI
        Nrow, lrow = self.blob_host[1, :self.n], self.blob_host[0, :self.n]
        dN = np.diff(Nrow)
        if self.n > 8 and np.all(dN == dN[0]) and dN[0] > 0 and np.max(np.abs(lrow[4:] - lrow[:-4])) < 300.0 and self.use_recurrence:
            self.desc.mu_recurrence = min(3, max(1, int(self.use_recurrence)))
        self.blob_host = np.ascontiguousarray(np.vstack([self.blob_host, extra]))
        self.n_rows += 2
        self.desc.n_rows = self.n_rows
        self.blob = torch().from_numpy(self.blob_host).to(self.device)
        self.h2d_bytes = self.blob_host.nbytes
        if hasattr(self, "_blob_pin"):
            del self._blob_pin

    def ensure_mu_tables(self):
        """Per-histogram tables of the compact-record mu sweep (fhmc_mu_tables_build: product tables + the phase structure
        of every elementary tilt interval), built once on the device and handed to the kernels through desc.mu_tables.
        FHMC_MU_TABLES=0 keeps the table-free kernel (k_sweep_prod2)."""
        if getattr(self, "_mu_tables_tried", False) or not self.use_mu_tables:
            return
        self._mu_tables_tried = True
        self._mu_tables = None
        if not self.use_mu_tables or self.desc.mu_recurrence < 2 or not self.desc.hull_len:
            return
        L = _lib.load()
        t = torch()
        nbytes = int(L.fhmc_mu_tables_bytes(ctypes.byref(self.desc)))
        if nbytes == 0:
            return
        buf = t.empty(nbytes + 256, dtype=t.uint8, device=self.device)
        ptr = (buf.data_ptr() + 255) & ~255
        with t.cuda.device(self.device):
            rc = L.fhmc_mu_tables_build(ctypes.byref(self.desc), _ptr(self.blob), ctypes.c_void_p(ptr), nbytes, _stream_ptr(self.device))
        if rc == 2:
            return
        _lib.check(rc, "fhmc_mu_tables_build")
        self._mu_tables = buf
        self.desc.mu_tables = ptr

    CELLS_MIN_STATES = 1 << 17   # building the cells costs about as much as walking 1.4x10^5 state points
    CELLS_EXTRA = 32768          # cells beyond one per elementary interval the buffer has room for (0.8 KB each at two quantities)

    def ensure_mu_cells(self, mu):
        """Tilt cells of a dense compact-record mu sweep (fhmc_mu_cells_build_for): moment expansions of the per-phase sums about
        the centres of small tilt cells covering the range of the device tensor ``mu``, handed to the kernels through
        desc.mu_cells.  Everything happens on the device, in stream order (no read-back): the cells are rebuilt whenever the sweep
        is not the tensor they were built for -- (storage, length, version) -- and reused otherwise.  State points the cells do not
        cover are still evaluated (by the table walk), so a stale key costs speed, never correctness.  FHMC_MU_CELLS=0 /
        use_mu_cells = False keeps the table walk (k_sweep_tab2)."""
        if not self.use_mu_cells or not self.desc.mu_tables:
            return
        t = torch()
        if not isinstance(mu, t.Tensor) or not mu.is_cuda or mu.numel() == 0 or mu.dtype != t.float64 or not mu.is_contiguous():
            return
        key = (mu.data_ptr(), mu.numel(), mu._version)
        if key == getattr(self, "_cells_key", None) and self.desc.mu_cells:
            return
        L = _lib.load()
        d = self.desc
        if getattr(self, "_mu_cells", None) is None:
            nbytes = int(L.fhmc_mu_cells_bytes(ctypes.byref(d), self.CELLS_EXTRA))
            if nbytes == 0:
                self.use_mu_cells = False
                return
            self._mu_cells = t.empty(nbytes + 256, dtype=t.uint8, device=self.device)
            self._mu_cells_bytes = nbytes
        ptr = (self._mu_cells.data_ptr() + 255) & ~255
        with t.cuda.device(self.device):
            rc = L.fhmc_mu_cells_build_for(ctypes.byref(d), _ptr(self.blob), ctypes.c_void_p(ptr), self._mu_cells_bytes, self.CELLS_EXTRA,
                                           ctypes.c_void_p(mu.data_ptr()), mu.numel(), _stream_ptr(self.device))
        if rc == 2:
            self.use_mu_cells = False
            return
        _lib.check(rc, "fhmc_mu_cells_build_for")
        self._cells_key = key
        self._cells_host_key = None
        self.desc.mu_cells = ptr

    # ------------------------------------------------------------------------------------------
    def _desc(self, pmax, complete=False, compare_raw=False, cutoff=None, smooth=None):
        d = _lib.HistDesc.from_buffer_copy(self.desc)
        if not self.use_mu_cells:
            d.mu_cells = None
        if not self.use_mu_tables:
            d.mu_tables = None
            d.mu_cells = None
        d.pmax = int(pmax)
        d.complete = 1 if complete else 0
        d.compare_raw = 1 if compare_raw else 0
        if cutoff is not None:
            d.cutoff = float(cutoff)
        if smooth is not None:
            if int(smooth) != d.smooth:
                d.mu_tables = None      # the tables belong to one window width
                d.mu_cells = None
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
        if st.n_states >= self.FAST_PATH_MIN_STATES and lanes in (0, 1) and not complete:
            self.ensure_hull()
        d = self._desc(pmax, complete, compare_raw, cutoff, smooth)
        if out is None:
            out = SweepResult(st.n_states, pmax, self.n_sel, self.device)
        elif out.n_states < st.n_states or out.pmax != pmax or out.n_sel != self.n_sel:
            raise ValueError("output buffers do not match the sweep")
        cs = out.c_struct()
        with torch().cuda.device(self.device):
            rc = L.fhmc_sweep_1d(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), ctypes.byref(cs), int(lanes),
                                 _stream_ptr(self.device))
        _lib.check(rc, "fhmc_sweep_1d")
        out._states = st
        return out

    def sweep_compact(self, mu1, pmax=4, dst=None, n_total=None, first=0, fill_dead=True, max_nphase=None, states=None):
        """K1+K3+K2 with COMPACT records (fhmc_sweep_1d_compact): the state points leave the sweep kernel as phase-major narrow
        records {status i16, nphase u8, fe/avg f64, bounds i16} at indices first .. first + S - 1 of every destination.

        dst: None (a fresh device buffer for n_total records), a uint8 device tensor, or a list of raw device pointers
        (ints) -- e.g. the buffers of this GPU's NVLink peers, for a sweep fused with its gather (parallel.py).
        Returns dict(buf, status, nphase, fe, avg, bounds) of [n_total, ...] views of the first destination when it is a
        tensor; asynchronous."""
        t = torch()
        L = _lib.load()
        # A repeated call (same device tensor of mu -- storage, length, version --, same destination and options, cells still the
        # ones built for it): everything the C call needs was prepared by the previous one.  The host side of a dense sweep
        # otherwise costs more than its 40 us on the device (descriptor copy, state-point struct, six strided views per call).
        plan_key = None
        if (states is None and isinstance(mu1, t.Tensor) and mu1.is_cuda and isinstance(dst, t.Tensor)):
            mkey = (mu1.data_ptr(), mu1.numel(), mu1._version)
            plan_key = (mkey, int(pmax), dst.data_ptr(), dst.numel(), n_total, int(first), bool(fill_dead),
                        max_nphase.data_ptr() if max_nphase is not None else 0, self.desc.mu_tables,
                        self.desc.mu_cells if self.use_mu_cells and self.use_mu_tables else None)
            plan = getattr(self, "_compact_plan", None)
            if plan is not None and plan[0] == plan_key and getattr(self, "_cells_key", None) == mkey and plan[7] is self._cws:
                _, d, st, co, ws_ptr, ws_bytes, out, _ = plan
                if t.cuda.current_device() == self.device.index:
                    rc = L.fhmc_sweep_1d_compact(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), ctypes.byref(co),
                                                 ctypes.c_void_p(ws_ptr), ws_bytes, _stream_ptr(self.device))
                else:
                    with t.cuda.device(self.device):
                        rc = L.fhmc_sweep_1d_compact(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), ctypes.byref(co),
                                                     ctypes.c_void_p(ws_ptr), ws_bytes, _stream_ptr(self.device))
                _lib.check(rc, "fhmc_sweep_1d_compact")
                return out
        st = states if states is not None else self.make_states(mu1)
        S = int(st.n_states)
        if S >= self.FAST_PATH_MIN_STATES:
            self.ensure_hull()
            self.ensure_mu_tables()
            if S >= self.CELLS_MIN_STATES and not st.beta and not st.dmu:
                keep = getattr(st, "_keep", None)
                if keep is not None and st.n_mu1 == S and st.mu1_div == 1:
                    self.ensure_mu_cells(keep[0])
        n_total = S + int(first) if n_total is None else int(n_total)
        d = self._desc(pmax)
        nbytes = int(L.fhmc_pack_soa16_bytes(n_total, pmax, self.n_sel))
        buf = None
        if dst is None:
            buf = t.empty(nbytes, dtype=t.uint8, device=self.device)
            ptrs = [buf.data_ptr()]
        elif isinstance(dst, t.Tensor):
            if dst.numel() < nbytes or dst.dtype != t.uint8:
                raise ValueError("destination must be a uint8 tensor of >= %d bytes" % nbytes)
            buf = dst
            ptrs = [dst.data_ptr()]
        else:
            ptrs = [int(x) for x in dst]
        co = _lib.CompactOut()
        for k, ptr in enumerate(ptrs):
            co.dst[k] = ptr
        co.n_dst, co.n_total, co.first, co.fill_dead = len(ptrs), n_total, int(first), 1 if fill_dead else 0
        co.max_nphase = max_nphase.data_ptr() if max_nphase is not None else None
        ws_bytes = int(L.fhmc_sweep_compact_workspace(ctypes.byref(d), S))
        if getattr(self, "_cws", None) is None or self._cws.numel() < ws_bytes + 256:
            self._cws = t.empty(ws_bytes + 256, dtype=t.uint8, device=self.device)
        ws_ptr = (self._cws.data_ptr() + 255) & ~255
        with t.cuda.device(self.device):
            rc = L.fhmc_sweep_1d_compact(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), ctypes.byref(co),
                                         ctypes.c_void_p(ws_ptr), ws_bytes, _stream_ptr(self.device))
        _lib.check(rc, "fhmc_sweep_1d_compact")
        out = {"buf": buf, "_states": st, "n_total": n_total}
        if buf is not None:
            out.update(soa16_views(buf, n_total, pmax, self.n_sel))
        if plan_key is not None:
            # (keyed on what the descriptor points at NOW: the call above may have built tables / cells)
            mkey = plan_key[0]
            plan_key = plan_key[:8] + (self.desc.mu_tables, self.desc.mu_cells if self.use_mu_cells and self.use_mu_tables else None)
            if getattr(self, "_cells_key", None) == mkey and plan_key[4] == n_total:
                self._compact_plan = (plan_key, d, st, co, ws_ptr, ws_bytes, out, self._cws)
        return out

    def sweep_host(self, mu1, pmax=4, lanes=0, chunk=1 << 18, out=None, fields=None):
        """End-to-end mu sweep with HOST buffers (what `bench.py`'s e2e times): the state points are cut into chunks
        that alternate between two CUDA streams so that the pinned-host -> device copy of chunk k+1, the kernel of
        chunk k and the device -> pinned-host copy of the results of chunk k-1 overlap.  The histogram blob is
        uploaded on every call.

        mu1: 1-D float64 NumPy array or CPU tensor (pinned memory avoids one staging copy).
        out: optional dict of pinned CPU tensors [S, ...] (as returned by a previous call) to reuse.
        Returns the dict of CPU tensors (fields of SweepResult.FIELDS, or the subset `fields`)."""
        t = torch()
        L = _lib.load()
        dev = self.device
        mu_h = mu1 if isinstance(mu1, t.Tensor) else t.from_numpy(np.ascontiguousarray(mu1, dtype=np.float64))
        if not mu_h.is_pinned():
            mu_h = mu_h.pin_memory()
        S = mu_h.numel()
        names = [k for k in SweepResult.FIELDS if (fields is None or k in fields) and (k != "avg" or self.n_sel)]
        chunk = int(min(chunk, max(S, 1)))
        if S >= self.FAST_PATH_MIN_STATES and lanes in (0, 1):
            self.ensure_hull()
        if not hasattr(self, "_pipe") or self._pipe[0] != (chunk, pmax):
            self._pipe = ((chunk, pmax), [t.cuda.Stream(dev), t.cuda.Stream(dev)],
                          [t.empty(chunk, dtype=t.float64, device=dev) for _ in range(2)],
                          [SweepResult(chunk, pmax, self.n_sel, dev) for _ in range(2)])
        _, streams, mu_d, res = self._pipe
        if out is None:
            out = {}
            for k in names:
                ref = getattr(res[0], k)
                out[k] = t.empty((S,) + tuple(ref.shape[1:]), dtype=ref.dtype).pin_memory()
        d = self._desc(pmax)
        cur = t.cuda.current_stream(dev)
        if not hasattr(self, "_blob_pin"):
            self._blob_pin = t.from_numpy(self.blob_host).pin_memory()
        blob_d = t.empty_like(self.blob)
        blob_d.copy_(self._blob_pin, non_blocking=True)
        ready = t.cuda.Event()
        ready.record(cur)
        for st_ in streams:
            st_.wait_event(ready)
        for k, lo in enumerate(range(0, S, chunk)):
            hi = min(S, lo + chunk)
            b = k & 1
            with t.cuda.stream(streams[b]):
                mu_d[b][:hi - lo].copy_(mu_h[lo:hi], non_blocking=True)
                st = _lib.States()
                st.n_states = hi - lo
                st.mu1, st.n_mu1, st.mu1_div = _ptr(mu_d[b]), hi - lo, 1
                st.beta, st.n_beta, st.beta_div = None, 1, 1
                st.dmu, st.n_dmu, st.dmu_div = None, 1, 1
                cs = res[b].c_struct()
                rc = L.fhmc_sweep_1d(ctypes.byref(d), _ptr(blob_d), ctypes.byref(st), ctypes.byref(cs), int(lanes),
                                     ctypes.c_void_p(streams[b].cuda_stream))
                _lib.check(rc, "fhmc_sweep_1d")
                for name in names:
                    out[name][lo:hi].copy_(getattr(res[b], name)[:hi - lo], non_blocking=True)
        for st_ in streams:
            cur.wait_stream(st_)
        blob_d.record_stream(streams[0])
        blob_d.record_stream(streams[1])
        cur.synchronize()
        return out

    def sweep_host_compact(self, mu1, pmax=4, lanes=0, chunk=1 << 17, out=None, narrow=None):
        """Host buffers in, host buffers out: result set {status, nphase, fe, avg, bounds} of a mu sweep, with only the phase
        slots that exist crossing PCIe.  One call of ``fhmc_sweep_host_compact`` (csrc/fhmc_host_pipe.cu), which pipelines
        H2D(mu) -> sweep -> phase-major repack -> D2H in chunks on three private streams (upload, compute, download).  The returned CPU tensors are
        [S, pmax, ...] VIEWS of one pinned phase-major buffer; slots p >= nphase[s] hold NaN / -1.  ``out`` = a previous
        result to reuse (its pinned buffers are recycled).  ``narrow`` (default: whenever the bin indices fit int16) sends
        the records of ``fhmc_pack_phase_soa16``: status as int16, nphase as uint8, bounds as int16 -- 60 instead of 72
        bytes per two-phase state point with two averaged quantities; fe / avg stay fp64."""
        t = torch()
        L = _lib.load()
        dev = self.device
        narrow = (self.n <= 32767) if narrow is None else bool(narrow)
        mu_h = mu1 if isinstance(mu1, t.Tensor) else t.from_numpy(np.ascontiguousarray(mu1, dtype=np.float64))
        if not mu_h.is_pinned():
            mu_h = mu_h.pin_memory()
        S = mu_h.numel()
        nsel = self.n_sel
        rec = 16 + 8 * nsel
        chunk = int(min(chunk, max(S, 1)))
        if S >= self.FAST_PATH_MIN_STATES and lanes in (0, 1):
            self.ensure_hull()
            self.ensure_mu_tables()
            # tilt cells for the range of this host array, built once per (buffer, length) from a device copy: the chunks of the
            # pipeline then take k_sweep_cell (a changed content only costs speed: what the cells miss takes the table walk)
            if self.use_mu_cells and S >= self.CELLS_MIN_STATES and self.desc.mu_tables:
                hkey = (mu_h.data_ptr(), S)
                if getattr(self, "_cells_host_key", None) != hkey:
                    self.ensure_mu_cells(mu_h.to(dev, non_blocking=True))
                    self._cells_host_key = hkey
        n_chunks = (S + chunk - 1) // chunk
        if not hasattr(self, "_hpipe") or self._hpipe["key"] != (chunk, pmax):
            ws_bytes = int(L.fhmc_sweep_host_workspace(chunk, pmax, nsel))
            self._hpipe = {"key": (chunk, pmax), "ws": t.empty(ws_bytes + 256, dtype=t.uint8, device=dev), "ws_bytes": ws_bytes}
        hpipe = self._hpipe
        ws_ptr = (hpipe["ws"].data_ptr() + 255) & ~255
        if out is None or out.get("_key") != (S, pmax, nsel, chunk, narrow):
            out = self._host_result_views(S, pmax, nsel, chunk, narrow)
            out["_flags"] = t.zeros(n_chunks, dtype=t.int32).pin_memory()
        d = self._desc(pmax)
        if not hasattr(self, "_blob_pin"):
            self._blob_pin = t.from_numpy(self.blob_host).pin_memory()
        cur = t.cuda.current_stream(dev)
        blob_d = t.empty_like(self.blob)
        blob_d.copy_(self._blob_pin, non_blocking=True)     # the histogram travels with every call (it is host data too)
        top, moved = ctypes.c_int(0), ctypes.c_longlong(0)
        with t.cuda.device(dev):
            fn = L.fhmc_sweep_host_compact16 if narrow else L.fhmc_sweep_host_compact
            rc = fn(ctypes.byref(d), _ptr(blob_d), ctypes.c_void_p(mu_h.data_ptr()), S, int(lanes), chunk,
                                           ctypes.c_void_p(ws_ptr), hpipe["ws_bytes"], ctypes.c_void_p(out["_buf"].data_ptr()),
                                           ctypes.c_void_p(out["_flags"].data_ptr()), int(out.get("max_nphase", 1)),
                                           ctypes.byref(top), ctypes.byref(moved), ctypes.c_void_p(cur.cuda_stream))
        _lib.check(rc, "fhmc_sweep_host_compact")
        top = int(top.value)
        # phase blocks that no chunk of this call filled: NaN / -1, like the empty slots inside a block
        for p in range(top, max(top, out["_prev_top"])):
            out["fe"][:, p].fill_(float("nan"))
            if nsel:
                out["avg"][:, p].fill_(float("nan"))
            out["bounds"][:, p].fill_(-1)
        out["_prev_top"] = top
        out["max_nphase"] = top
        out["d2h_bytes"] = int(moved.value) + 4 * n_chunks       # + the per-chunk phase counts
        return out

    def _host_result_views(self, S, pmax, nsel, chunk, narrow=False):
        """One pinned phase-major buffer (layout of fhmc_pack_phase_major, or of fhmc_pack_phase_soa16 when ``narrow``, for
        S state points) and [S, pmax, ...] views of it."""
        t = torch()
        if narrow:
            L = _lib.load()
            nf = 1 + nsel
            buf = t.empty(int(L.fhmc_pack_soa16_bytes(S, pmax, nsel)), dtype=t.uint8).pin_memory()
            out = {"_key": (S, pmax, nsel, chunk, True), "_buf": buf, "_prev_top": pmax}
            out.update(soa16_views(buf, S, pmax, nsel))
            return out
        rec = 16 + 8 * nsel
        buf = t.empty(8 * S + pmax * rec * S, dtype=t.uint8).pin_memory()
        out = {"_key": (S, pmax, nsel, chunk, False), "_buf": buf, "_prev_top": pmax}   # every phase block starts out stale
        head = buf[:8 * S].view(t.int32)
        out["status"] = t.as_strided(head, (S,), (2,))
        out["nphase"] = t.as_strided(head, (S,), (2,), head.storage_offset() + 1)
        body64, body32 = buf[8 * S:].view(t.float64), buf[8 * S:].view(t.int32)
        out["fe"] = t.as_strided(body64, (S, pmax), (rec // 8, S * rec // 8))
        out["avg"] = t.as_strided(body64, (S, pmax, nsel), (rec // 8, S * rec // 8, 1), body64.storage_offset() + 1) if nsel else None
        out["bounds"] = t.as_strided(body32, (S, pmax, 2), (rec // 4, S * rec // 4, 1), body32.storage_offset() + 2 + 2 * nsel)
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

    def _solver_states(self, mu_guess, beta, dmu):
        """fhmc_states of a flat solve list from HOST arrays with ONE pinned staging buffer and one asynchronous upload
        (three pageable copies cost three synchronisations: 0.3 ms of a 1 ms launch)."""
        t = torch()
        cols = [np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64) for x in (mu_guess, beta, dmu) if x is not None]
        T = max(c.size for c in cols)
        for c in cols:
            if c.size not in (1, T):
                raise ValueError("flat state lists must have equal length (or length 1)")
        key = (len(cols), T)
        if getattr(self, "_sv_stage", None) is None or self._sv_stage[0] != key:
            self._sv_stage = (key, t.empty((len(cols), T), dtype=t.float64).pin_memory())
        stage = self._sv_stage[1]
        sn = stage.numpy()
        for k, c in enumerate(cols):
            sn[k, :] = c
        dev_t = t.empty((len(cols), T), dtype=t.float64, device=self.device)
        dev_t.copy_(stage, non_blocking=True)
        st = _lib.States()
        st.n_states = T
        st.mu1_div = st.beta_div = st.dmu_div = 1
        k = 0
        st.mu1, st.n_mu1 = ctypes.c_void_p(dev_t[0].data_ptr()), T
        k = 1
        if beta is not None:
            st.beta, st.n_beta = ctypes.c_void_p(dev_t[k].data_ptr()), T
            k += 1
        else:
            st.beta, st.n_beta = None, 1
        if dmu is not None:
            st.dmu, st.n_dmu = ctypes.c_void_p(dev_t[k].data_ptr()), T
        else:
            st.dmu, st.n_dmu = None, 1
        st._keep = (dev_t,)
        return st

    def find_phase_eq(self, mu_guess, beta=None, dmu=None, lnz_tol=1e-10, mu_step=None, max_iter=200, pmax=4,
                      smooth=None, cutoff=None, min_width=None, continuation=None, stride=32, seed_stride=64):
        """K4: one coexistence solve per entry of (mu_guess, beta, dmu) (flat lists).  Returns a SweepResult at coexistence
        with extra['mu_coex', 'dfe', 'iters'] (iters = evaluations of the last stage).

        continuation (default: automatic): when every solve starts from the SAME cold guess and the temperatures differ,
        every `stride`-th temperature (in order of beta) is solved first and the guesses of all solves are interpolated
        from those roots -- what a user's notebook does by hand when it feeds the previous temperature's mu into the next
        call, done hierarchically so that both stages stay batched.  The second stage then needs ~3 evaluations per
        solve instead of ~6.5 and never visits the monotone ln(PI) far from coexistence.  A list that is already ordered
        in beta takes ONE launch: every ``seed_stride``-th solve is a seed, the others wait inside the kernel for the two
        seeds around them and start from the interpolated root (``fhmc_find_phase_eq_curve``)."""
        t = torch()
        if self.n_sel < 1:
            raise ValueError("the solver needs quantity 0 to be N_tot (construct DeviceHistogram with sel=['N', ...])")
        host_in = not any(isinstance(x, t.Tensor) for x in (mu_guess, beta, dmu))
        if continuation is None:
            continuation = False
            if host_in and beta is not None and np.size(beta) >= 8 * stride and (dmu is None or np.size(dmu) == 1):
                g = np.atleast_1d(np.asarray(mu_guess, dtype=np.float64))
                continuation = bool(np.all(g == g.flat[0]))
        if continuation and host_in and beta is not None and np.size(beta) > stride and (dmu is None or np.size(dmu) == 1):
            b = np.atleast_1d(np.asarray(beta, dtype=np.float64))
            db = np.diff(b)
            if b.ndim == 1 and (np.all(db > 0) or np.all(db < 0)):
                # the list already runs along the curve: one launch, the seeds (every seed_stride-th solve) and the
                # interpolated guesses of the other solves are handled inside the kernel (fhmc_find_phase_eq_curve)
                out = self._find_phase_eq_once(mu_guess, beta, dmu, lnz_tol, mu_step, max_iter, pmax, smooth, cutoff, min_width,
                                               seed_stride=max(2, int(seed_stride)))
                out.extra["coarse_solves"] = (len(b) - 1) // max(2, int(seed_stride)) + 1
                return out
            g = np.broadcast_to(np.atleast_1d(np.asarray(mu_guess, dtype=np.float64)), b.shape)
            d = None if dmu is None else np.broadcast_to(np.atleast_1d(np.asarray(dmu, dtype=np.float64)), b.shape)
            order = np.argsort(b, kind="stable")
            pick = np.unique(np.concatenate([np.arange(0, len(b), stride), [len(b) - 1]]))
            ci = order[pick]
            rc = self._find_phase_eq_once(g[ci], b[ci], None if d is None else d[ci], lnz_tol, mu_step, max_iter, pmax, smooth,
                                          cutoff, min_width)
            code = (rc.status & ST_CODE_MASK).cpu().numpy()
            jump = (rc.status.cpu().numpy().view(np.uint32) & _lib.ST_JUMP) != 0
            mu_c = rc.extra["mu_coex"].cpu().numpy()
            good = (code == 0) & ~jump
            if good.sum() >= 2:
                g2 = np.empty_like(b)
                g2[order] = np.interp(np.arange(len(b)), pick[good], mu_c[good])   # rank space; clamped beyond the last root
                out = self._find_phase_eq_once(g2, b, d, lnz_tol, mu_step, max_iter, pmax, smooth, cutoff, min_width)
                out.extra["coarse_solves"] = len(ci)
                return out
        return self._find_phase_eq_once(mu_guess, beta, dmu, lnz_tol, mu_step, max_iter, pmax, smooth, cutoff, min_width)

    def _find_phase_eq_once(self, mu_guess, beta, dmu, lnz_tol, mu_step, max_iter, pmax, smooth, cutoff, min_width, seed_stride=0):
        L = _lib.load()
        t = torch()
        if any(isinstance(x, t.Tensor) for x in (mu_guess, beta, dmu)):
            st = self.make_states(mu_guess, beta, dmu, grid=False)
        else:
            st = self._solver_states(mu_guess, beta, dmu)
        d = self._desc(max(pmax, 2), False, False, cutoff, smooth)
        d.min_width = int(min_width) if min_width else 0    # 0: 2*smooth (N_tot histograms); the N_1 class passes smooth
        out = SweepResult(st.n_states, d.pmax, self.n_sel, self.device)
        T = st.n_states
        extra = t.empty((2, T), dtype=t.float64, device=self.device)
        mu_coex, dfe = extra[0], extra[1]
        iters = t.empty(T, dtype=t.int32, device=self.device)
        if mu_step is None:
            mu_step = 0.05 / abs(self.desc.beta_ref)   # first blind search step (doubles until <N> brackets the window)
        cs = out.c_struct()
        with t.cuda.device(self.device):
            if seed_stride and seed_stride > 1:
                rc = L.fhmc_find_phase_eq_curve(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), float(lnz_tol),
                                                float(mu_step), int(max_iter), int(seed_stride), _ptr(mu_coex), _ptr(dfe), _ptr(iters),
                                                ctypes.byref(cs), _stream_ptr(self.device))
            else:
                rc = L.fhmc_find_phase_eq_1d(ctypes.byref(d), _ptr(self.blob), ctypes.byref(st), float(lnz_tol),
                                             float(mu_step), int(max_iter), _ptr(mu_coex), _ptr(dfe), _ptr(iters),
                                             ctypes.byref(cs), _stream_ptr(self.device))
        _lib.check(rc, "fhmc_find_phase_eq_1d")
        out.extra = {"mu_coex": mu_coex, "dfe": dfe, "iters": iters}
        out._states = st
        return out


class ScalarPath(object):
    """Device-resident state behind the SCALAR drop-in calls (histogram.reweight / normalize / relextrema / thermo on one
    state point; SURVEY 8(b): "device buffers are caches keyed on array identity/version").

    One instance per (device, histogram length), shared by every histogram object of that length: the host NumPy arrays stay
    the source of truth (the reference's tests overwrite ``hist.data['ln(PI)']`` directly), so every call compares a cheap
    content key of ln(PI), N and the moment tensor with what the device holds and uploads only what changed; a call is ONE
    C-ABI call (``fhmc_scalar_point``: uploads, sweep record, normalised row, per-phase moment averages, one download, one
    synchronisation) into a pinned result buffer."""

    PMAX = 8
    MAX_CACHED = 16      # histogram lengths kept resident per process (least recently used goes first)
    _cache = {}          # insertion-ordered: (device index, n) -> ScalarPath.  Not thread-safe: one scalar call at a time per process

    @classmethod
    def get(cls, n, device=None):
        if device is None:
            t = torch()
            if cls._cache:                      # (a CUDA context exists: the cheap query is enough)
                key = (t.cuda.current_device(), int(n))
                sp = cls._cache.get(key)
                if sp is not None:
                    if len(cls._cache) > 1:
                        cls._cache[key] = cls._cache.pop(key)     # most recently used last
                    return sp
        dev = require_cuda(device)
        key = (dev.index if dev.index is not None else torch().cuda.current_device(), int(n))
        sp = cls._cache.get(key)
        if sp is None:
            while len(cls._cache) >= cls.MAX_CACHED:
                cls._cache.pop(next(iter(cls._cache)))
            sp = cls._cache[key] = cls(int(n), dev)
        return sp

    def __init__(self, n, device):
        t = torch()
        self.n, self.n_pad, self.device = n, n + (n & 1), device
        self._index = device.index if device.index is not None else t.cuda.current_device()
        self.blob = t.zeros((2, self.n_pad), dtype=t.float64, device=device)
        self.mu1_dev = t.zeros(1, dtype=t.float64, device=device)
        self.stage = t.zeros(2 * self.n_pad + 2, dtype=t.float64).pin_memory()     # ln(PI) | N | mu_1
        self.stage_np = self.stage.numpy()
        self.lnpi_key = self.ntot_key = self.mom_key = None
        self.mom_dev, self.n_arrays, self._w = None, 0, None
        self._alloc_out(0)
        d = _lib.HistDesc()
        d.n, d.n_pad, d.n_rows = n, self.n_pad, 2
        d.sel_kind[0] = M_ONE
        d.n_term = 1
        d.pmax = self.PMAX
        self.desc = d

    def _alloc_out(self, n_arrays):
        """One device range (and its pinned mirror) for the record, the normalised row, lnsum and the moment averages."""
        t = torch()
        layout, rec_bytes = SweepResult.layout(1, self.PMAX, 0)
        off = (rec_bytes + 15) & ~15
        self._off_row = off
        off += self.n_pad * 8
        self._off_lnsum = off
        off += self.PMAX * 8
        self._off_avg = off
        off += self.PMAX * max(n_arrays, 1) * 8
        self.out_bytes = off
        self.out_dev = t.zeros(off, dtype=t.uint8, device=self.device)
        self.out_host = t.zeros(off, dtype=t.uint8).pin_memory()
        self.out_np = self.out_host.numpy()
        self._layout = layout
        self._views = [(name, self.out_np[o:o + nb].view(dt).reshape(shape)[0] if len(shape) > 1 else self.out_np[o:o + nb].view(dt))
                       for name, dt, shape, o, nb in layout]
        base = self.out_dev.data_ptr()
        ptr = {name: base + o for name, _, _, o, _ in layout}
        self.rec = _lib.SweepOut(*[ptr.get(k) for k in SweepResult.FIELDS])
        self.n_arrays = n_arrays

    @staticmethod
    def _key(a):
        return (a.shape, hash(a.tobytes()))

    def _mom_key(self, mom2d):
        # exact arithmetic on the bit patterns: row sums modulo 2^64, combined with fixed odd weights per row.  Every bit of
        # every entry counts whatever the magnitudes (a floating-point checksum is blind to changes below the rounding of its
        # largest terms, the high-order moments); ~14 us for the 27 x 1001 tensor against ~110 us for hash(bytes)
        bits = mom2d.view(np.uint64)
        if self._w is None or self._w.shape[0] != bits.shape[0]:
            rng = np.random.default_rng(0x5EED)
            self._w = rng.integers(0, 2 ** 63, size=bits.shape[0], dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        return (mom2d.shape, int(np.dot(bits.sum(axis=1), self._w)))

    def point(self, lnpi, ntot, beta_ref, mu1_ref, smooth, mu1_target, complete=False, compare_raw=False, cutoff=10.0,
              want_row=False, mom=None):
        """One state point.  Returns a dict: code, status, nphase, nmin, lnnorm, fe, bounds, max_idx, min_idx (+ 'row' with
        want_row, + 'avg' [P][A] and 'lnsum' [P] with ``mom`` = the [A][n] moment rows)."""
        L = _lib.load()
        n, npad = self.n, self.n_pad
        io = _lib.ScalarIO()
        lnpi = np.ascontiguousarray(lnpi, dtype=np.float64)
        k = self._key(lnpi)
        if k != self.lnpi_key:
            self.stage_np[:n] = lnpi
            io.lnpi_host = self.stage.data_ptr()
            self.lnpi_key = k
        ntot = np.ascontiguousarray(ntot, dtype=np.float64)
        k = self._key(ntot)
        if k != self.ntot_key:
            self.stage_np[npad:npad + n] = ntot
            io.ntot_host = self.stage.data_ptr() + 8 * npad
            self.ntot_key = k
        want_mom = mom is not None
        if want_mom:
            mom = np.ascontiguousarray(mom, dtype=np.float64).reshape(-1, n)
            k = self._mom_key(mom)
            if k != self.mom_key:
                t = torch()
                if mom.shape[0] != self.n_arrays:
                    self._alloc_out(mom.shape[0])
                self.mom_dev = t.from_numpy(mom).to(self.device)
                self.mom_key = k
        d = self.desc
        d.smooth, d.complete, d.compare_raw = max(int(smooth), 1), 1 if complete else 0, 1 if compare_raw else 0
        d.cutoff, d.beta_ref, d.mu1_ref, d.dmu_ref = float(cutoff), float(beta_ref), float(mu1_ref), 0.0
        base = self.out_dev.data_ptr()
        io.blob, io.mu1_dev = self.blob.data_ptr(), self.mu1_dev.data_ptr()
        io.mu1_pinned, io.mu1 = self.stage.data_ptr() + 8 * 2 * npad, float(mu1_target)
        io.rec = self.rec
        io.row = base + self._off_row
        io.lnsum = base + self._off_lnsum
        io.avg = base + self._off_avg
        io.mom = self.mom_dev.data_ptr() if want_mom else None
        io.n_arrays = self.n_arrays if want_mom else 0
        io.out_dev, io.out_host, io.out_bytes = base, self.out_host.data_ptr(), self.out_bytes
        tc = torch().cuda
        if tc.current_device() == self._index:
            rc = L.fhmc_scalar_point(ctypes.byref(d), ctypes.byref(io), 1 if (want_row or want_mom) else 0, 1 if want_mom else 0,
                                     tc.current_stream().cuda_stream)
        else:
            with tc.device(self.device):
                rc = L.fhmc_scalar_point(ctypes.byref(d), ctypes.byref(io), 1 if (want_row or want_mom) else 0, 1 if want_mom else 0,
                                         tc.current_stream().cuda_stream)
        if rc != 0:
            self.lnpi_key = self.ntot_key = None     # the uploads may not have happened
        _lib.check(rc, "fhmc_scalar_point")
        hb = self.out_np
        out = {}
        for name, view in self._views:
            out[name] = view.copy()
        status = int(out["status"][0]) & 0xFFFFFFFF
        out["status"] = status
        out["code"] = status & ST_CODE_MASK
        out["safe"] = (status & ST_SAFE) != 0
        out["nphase"], out["nmin"], out["lnnorm"] = int(out["nphase"][0]), int(out["nmin"][0]), float(out["lnnorm"][0])
        if want_row or want_mom:
            out["row"] = hb[self._off_row:self._off_row + 8 * n].view(np.float64).copy()
        if want_mom:
            P = out["nphase"] if out["code"] == 0 else 0
            out["lnsum"] = hb[self._off_lnsum:self._off_lnsum + 8 * self.PMAX].view(np.float64)[:P].copy()
            A = self.n_arrays
            out["avg"] = hb[self._off_avg:self._off_avg + 8 * self.PMAX * A].view(np.float64).reshape(self.PMAX, A)[:P].copy()
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


def reweight_2d(lnpi, bounds, op1, op2, a1, a2, props=None, device=None, return_device=False, product=None):
    """K5: 2-D joint histogram reweight for S state points; returns [S][3+n_prop] (lnZ, <op1>, <op2>, <prop>...).
    product: use the product-form kernel (uniformly spaced op2, |a2| * span(op2) < 300); None = decide from host copies
    of op2 / a2 when they are host arrays, else the exp-per-bin kernel."""
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
    if product is None:
        product = False
        if not isinstance(op2, t.Tensor) and not isinstance(a2, t.Tensor) and n2 >= 8 and S >= 64:
            o2h, a2h = np.asarray(op2, dtype=np.float64), np.atleast_1d(np.asarray(a2, dtype=np.float64))
            d2 = np.diff(o2h)
            product = bool(np.all(d2 == d2[0]) and d2[0] > 0 and np.max(np.abs(a2h)) * (o2h[-1] - o2h[0]) < 300.0)
    ws_bytes = L.fhmc_reweight_2d_prod_workspace(n1, n2, n_prop, S) if product else 0
    if product and ws_bytes == 0:
        product = False
    if not product:
        ws_bytes = L.fhmc_reweight_2d_workspace(n1, n2, n_prop, S)
    ws = t.empty(max(ws_bytes // 8, 1), dtype=t.float64, device=dev)
    fn = L.fhmc_reweight_2d_prod if product else L.fhmc_reweight_2d
    with t.cuda.device(dev):
        rc = fn(_ptr(lnpi_t), _ptr(b_t), n1, n2, _ptr(o1), _ptr(o2), _ptr(p_t), n_prop, _ptr(a1_t),
                _ptr(a2_t), S, _ptr(out), _ptr(ws), ws_bytes, _stream_ptr(dev))
    _lib.check(rc, "fhmc_reweight_2d_prod" if product else "fhmc_reweight_2d")
    return out if return_device else out.cpu().numpy()


MASKED_2D_MAXPROP = 8


def masked_lse_2d(lnpi, mask=None, edge=None, props=None, shifted=False, device=None, peak_cap=64):
    """One-shot ragged / masked log-sum-exp of a surface lnPI[n1, n2] with averages of property matrices
    (pore_hist.normalize / pore_hist.thermo, two_dim/h_ntot/pore_hist.pyx:57-80, 154-184).

    mask  [n1, n2] bool (True = bin belongs to the region) or None;  edge [n1] = last valid column per row or None;
    props [n_prop, n1, n2] or None (any n_prop: processed eight matrices per launch).
    Returns dict(lnsum, max, avg[n_prop], peak_idx = (rows, cols) of the selected bins equal to the maximum,
    shifted = lnPI - lnsum when ``shifted``)."""
    L = _lib.load()
    t = torch()
    dev = require_cuda(device)
    lnpi_t = t.from_numpy(np.ascontiguousarray(lnpi, dtype=np.float64)).to(dev)
    if lnpi_t.dim() != 2:
        raise ValueError("lnpi must be two-dimensional")
    n1, n2 = lnpi_t.shape
    mask_t = None
    if mask is not None:
        mk = np.ascontiguousarray(mask).astype(np.uint8)
        if mk.shape != (n1, n2):
            raise ValueError("mask must have the shape of ln(PI)")
        mask_t = t.from_numpy(mk).to(dev)
    edge_t = None
    if edge is not None:
        ed = np.ascontiguousarray(edge, dtype=np.int32)
        if ed.shape != (n1,):
            raise ValueError("edge must hold one column index per row")
        edge_t = t.from_numpy(ed).to(dev)
    pr = None if props is None or len(props) == 0 else np.ascontiguousarray(props, dtype=np.float64)
    n_prop = 0 if pr is None else pr.shape[0]
    if pr is not None and pr.shape[1:] != (n1, n2):
        raise ValueError("props must be [n_prop, n1, n2]")
    avg = np.empty(n_prop, dtype=np.float64)
    shifted_t = t.empty_like(lnpi_t) if shifted else None
    res = None
    with t.cuda.device(dev):
        for q0 in range(0, max(n_prop, 1), MASKED_2D_MAXPROP):
            nq = min(MASKED_2D_MAXPROP, n_prop - q0) if n_prop else 0
            p_t = t.from_numpy(pr[q0:q0 + nq]).to(dev) if nq else None
            out = t.empty(2 + MASKED_2D_MAXPROP, dtype=t.float64, device=dev)
            peak = t.empty(1 + peak_cap, dtype=t.int64, device=dev)
            ws_bytes = L.fhmc_masked_lse_2d_workspace(n1, n2, nq)
            ws = t.empty(max(ws_bytes // 8, 2), dtype=t.float64, device=dev)
            rc = L.fhmc_masked_lse_2d(_ptr(lnpi_t), _ptr(mask_t), _ptr(edge_t), n1, n2, _ptr(p_t), nq, _ptr(out), _ptr(peak),
                                      peak_cap, _ptr(shifted_t) if (shifted and q0 == 0) else None, _ptr(ws), ws_bytes,
                                      _stream_ptr(dev))
            _lib.check(rc, "fhmc_masked_lse_2d")
            oh = out.cpu().numpy()
            avg[q0:q0 + nq] = oh[2:2 + nq]
            if res is None:
                pk = peak.cpu().numpy()
                if pk[0] > peak_cap:   # more ties with the maximum than the buffer holds: ask again with room for all
                    return masked_lse_2d(lnpi, mask, edge, props, shifted, device, peak_cap=int(pk[0]))
                flat = np.sort(pk[1:1 + pk[0]])
                res = {"lnsum": float(oh[0]), "max": float(oh[1]), "peak_idx": (flat // n2, flat % n2)}
    res["avg"] = avg
    if shifted:
        res["shifted"] = shifted_t.cpu().numpy()
    return res


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

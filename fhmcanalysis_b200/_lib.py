"""ctypes binding of libfhmc_b200.so (C ABI declared in include/fhmc_b200.h).

The library is built in-tree by ``fhmcanalysis_b200.build`` (nvcc, sm_100a).  There is NO CPU
fallback: if the library is missing or no CUDA device is present every compute entry point raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FHMC_LIB_PATH") or os.path.join(HERE, "libfhmc_b200.so")   # (override: A/B runs of experimental builds)

MAX_TERMS = 8
MAX_SEL = 4

# enum fhmc_monomial
M_DB, M_DD, M_DB2, M_DBDD, M_DD2, M_DB3, M_DB_MU1, M_ONE = range(8)

ST_CODE_MASK = 0xFF
ST_SAFE = 0x100
ST_GAP_FILL = 0x200
ST_SLOW_PATH = 0x400
ST_RESCUED = 0x800
ST_FAST = 0x1000
ST_JUMP = 0x2000
ST_LEAN = 0x4000
E_CAPACITY = 8
E_NO_COEX = 100

STATUS_TEXT = {
    0: "ok",
    1: "ln(PI) not long enough to analyze for relative extrema",
    2: "Bad relative extrema calculation",
    3: "Bad relative extrema calculation",
    4: "local maxima and local minima cannot be alternating, try adjusting the value of smooth",
    5: "Local maxima and minima not sorted correctly, try adjusting the value of smooth",
    6: "index out of bounds while assigning phase bounds",
    7: "tied extrema between neighbouring maxima/minima",
    8: "more extrema than pmax",
    100: "no pair of sufficiently wide phases / no bracket for coexistence",
}


class HistDesc(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int), ("n_pad", ctypes.c_int), ("n_rows", ctypes.c_int), ("n_coef", ctypes.c_int),
        ("coef_row", ctypes.c_int * MAX_TERMS), ("coef_kind", ctypes.c_int * MAX_TERMS),
        ("n_sel", ctypes.c_int), ("n_term", ctypes.c_int),
        ("sel_row", ctypes.c_int * MAX_SEL), ("sel_kind", ctypes.c_int * MAX_TERMS),
        ("smooth", ctypes.c_int), ("pmax", ctypes.c_int), ("complete", ctypes.c_int), ("compare_raw", ctypes.c_int),
        ("cutoff", ctypes.c_double), ("beta_ref", ctypes.c_double), ("mu1_ref", ctypes.c_double),
        ("dmu_ref", ctypes.c_double),
        ("hull_row", ctypes.c_int), ("hull_len", ctypes.c_int), ("mu_recurrence", ctypes.c_int), ("min_width", ctypes.c_int),
        ("mu_tables", ctypes.c_void_p), ("mu_cells", ctypes.c_void_p),
    ]


class States(ctypes.Structure):
    _fields_ = [
        ("n_states", ctypes.c_longlong),
        ("mu1", ctypes.c_void_p), ("n_mu1", ctypes.c_longlong), ("mu1_div", ctypes.c_longlong),
        ("beta", ctypes.c_void_p), ("n_beta", ctypes.c_longlong), ("beta_div", ctypes.c_longlong),
        ("dmu", ctypes.c_void_p), ("n_dmu", ctypes.c_longlong), ("dmu_div", ctypes.c_longlong),
    ]


class CompactOut(ctypes.Structure):
    _fields_ = [("dst", ctypes.c_void_p * 8), ("n_dst", ctypes.c_int), ("n_total", ctypes.c_longlong), ("first", ctypes.c_longlong),
                ("fill_dead", ctypes.c_int), ("max_nphase", ctypes.c_void_p)]


class SweepOut(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in
                ("status", "nphase", "nmin", "lnnorm", "fe", "avg", "bounds", "max_idx", "min_idx")]


class ScalarIO(ctypes.Structure):
    """fhmc_scalar_io (include/fhmc_b200.h): buffers of one fhmc_scalar_point call."""
    _fields_ = [("blob", ctypes.c_void_p), ("mu1_dev", ctypes.c_void_p), ("mu1_pinned", ctypes.c_void_p), ("mu1", ctypes.c_double),
                ("lnpi_host", ctypes.c_void_p), ("ntot_host", ctypes.c_void_p), ("rec", SweepOut), ("row", ctypes.c_void_p),
                ("mom", ctypes.c_void_p), ("n_arrays", ctypes.c_int), ("avg", ctypes.c_void_p), ("lnsum", ctypes.c_void_p),
                ("out_dev", ctypes.c_void_p), ("out_host", ctypes.c_void_p), ("out_bytes", ctypes.c_size_t)]


EXPORTS = [
    "fhmc_version", "fhmc_last_error", "fhmc_last_kernel", "fhmc_device_info", "fhmc_sweep_1d", "fhmc_lnpi_1d",
    "fhmc_phase_moments", "fhmc_axpy_rows", "fhmc_find_phase_eq_1d", "fhmc_reweight_2d",
    "fhmc_reweight_2d_workspace", "fhmc_pack_bytes", "fhmc_pack_phase_major",
    "fhmc_masked_lse_2d", "fhmc_masked_lse_2d_workspace", "fhmc_sweep_host_workspace", "fhmc_sweep_host_compact",
    "fhmc_sweep_host_compact16", "fhmc_pack_soa16_bytes", "fhmc_pack_phase_soa16",
    "fhmc_patch_shifts", "fhmc_reweight_2d_prod", "fhmc_reweight_2d_prod_workspace",
    "fhmc_bench_dfma", "fhmc_bench_exp", "fhmc_lean_stats", "fhmc_sweep_1d_compact", "fhmc_sweep_compact_workspace",
    "fhmc_mu_tables_bytes", "fhmc_mu_tables_build", "fhmc_mu_cells_bytes", "fhmc_mu_cells_build", "fhmc_mu_cells_build_for", "fhmc_phase_moments_dev", "fhmc_scalar_point", "fhmc_find_phase_eq_curve",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load libfhmc_b200.so; raises LibraryMissing (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "%s not found: build it with `python -m fhmcanalysis_b200.build` (nvcc, sm_100a). "
            "fhmcanalysis_b200 has no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_longlong
    L.fhmc_version.restype = ci
    L.fhmc_last_error.restype = ctypes.c_char_p
    L.fhmc_last_kernel.restype = ctypes.c_char_p
    L.fhmc_device_info.restype = ci
    L.fhmc_device_info.argtypes = [ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.fhmc_sweep_1d.restype = ci
    L.fhmc_sweep_1d.argtypes = [ctypes.POINTER(HistDesc), vp, ctypes.POINTER(States), ctypes.POINTER(SweepOut), ci, vp]
    L.fhmc_lnpi_1d.restype = ci
    L.fhmc_lnpi_1d.argtypes = [ctypes.POINTER(HistDesc), vp, ctypes.POINTER(States), vp, vp, vp]
    L.fhmc_phase_moments.restype = ci
    L.fhmc_phase_moments.argtypes = [vp, ci, vp, ci, vp, ci, vp, vp, vp]
    L.fhmc_pack_bytes.restype = cll
    L.fhmc_pack_bytes.argtypes = [cll, ci, ci]
    L.fhmc_pack_phase_major.restype = ci
    L.fhmc_pack_phase_major.argtypes = [ctypes.POINTER(SweepOut), cll, ci, ci, vp, vp, vp]
    L.fhmc_axpy_rows.restype = ci
    L.fhmc_axpy_rows.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(cd), ci, cll, vp, vp]
    L.fhmc_find_phase_eq_1d.restype = ci
    L.fhmc_find_phase_eq_1d.argtypes = [ctypes.POINTER(HistDesc), vp, ctypes.POINTER(States), cd, cd, ci,
                                        vp, vp, vp, ctypes.POINTER(SweepOut), vp]
    L.fhmc_find_phase_eq_curve.restype = ci
    L.fhmc_find_phase_eq_curve.argtypes = [ctypes.POINTER(HistDesc), vp, ctypes.POINTER(States), cd, cd, ci, ci,
                                           vp, vp, vp, ctypes.POINTER(SweepOut), vp]
    L.fhmc_reweight_2d.restype = ci
    L.fhmc_reweight_2d_workspace.restype = ctypes.c_size_t
    L.fhmc_reweight_2d_workspace.argtypes = [ci, ci, ci, cll]
    L.fhmc_reweight_2d.argtypes = [vp, vp, ci, ci, vp, vp, vp, ci, vp, vp, cll, vp, vp, ctypes.c_size_t, vp]
    L.fhmc_masked_lse_2d_workspace.restype = ctypes.c_size_t
    L.fhmc_masked_lse_2d_workspace.argtypes = [ci, ci, ci]
    L.fhmc_masked_lse_2d.restype = ci
    L.fhmc_masked_lse_2d.argtypes = [vp, vp, vp, ci, ci, vp, ci, vp, vp, ci, vp, vp, ctypes.c_size_t, vp]
    L.fhmc_sweep_host_workspace.restype = ctypes.c_size_t
    L.fhmc_sweep_host_workspace.argtypes = [cll, ci, ci]
    L.fhmc_sweep_host_compact.restype = ci
    L.fhmc_sweep_host_compact.argtypes = [ctypes.POINTER(HistDesc), vp, vp, cll, ci, cll, vp, ctypes.c_size_t, vp, vp, ci,
                                          ctypes.POINTER(ci), ctypes.POINTER(cll), vp]
    L.fhmc_reweight_2d_prod.restype = ci
    L.fhmc_reweight_2d_prod_workspace.restype = ctypes.c_size_t
    L.fhmc_reweight_2d_prod_workspace.argtypes = [ci, ci, ci, cll]
    L.fhmc_reweight_2d_prod.argtypes = [vp, vp, ci, ci, vp, vp, vp, ci, vp, vp, cll, vp, vp, ctypes.c_size_t, vp]
    L.fhmc_sweep_host_compact16.restype = ci
    L.fhmc_sweep_host_compact16.argtypes = L.fhmc_sweep_host_compact.argtypes
    L.fhmc_pack_soa16_bytes.restype = cll
    L.fhmc_pack_soa16_bytes.argtypes = [cll, ci, ci]
    L.fhmc_pack_phase_soa16.restype = ci
    L.fhmc_pack_phase_soa16.argtypes = [ctypes.POINTER(SweepOut), cll, ci, ci, vp, vp, vp]
    L.fhmc_patch_shifts.restype = ci
    L.fhmc_patch_shifts.argtypes = [vp, vp, vp, ci, vp, vp, vp]
    L.fhmc_bench_dfma.restype = cll
    L.fhmc_bench_dfma.argtypes = [ci, vp, vp]
    L.fhmc_bench_exp.restype = cll
    L.fhmc_bench_exp.argtypes = [ci, vp, vp]
    L.fhmc_sweep_compact_workspace.restype = ctypes.c_size_t
    L.fhmc_sweep_compact_workspace.argtypes = [ctypes.POINTER(HistDesc), cll]
    L.fhmc_sweep_1d_compact.restype = ci
    L.fhmc_sweep_1d_compact.argtypes = [ctypes.POINTER(HistDesc), vp, ctypes.POINTER(States), ctypes.POINTER(CompactOut), vp,
                                        ctypes.c_size_t, vp]
    L.fhmc_mu_tables_bytes.restype = ctypes.c_size_t
    L.fhmc_mu_tables_bytes.argtypes = [ctypes.POINTER(HistDesc)]
    L.fhmc_mu_tables_build.restype = ci
    L.fhmc_mu_tables_build.argtypes = [ctypes.POINTER(HistDesc), vp, vp, ctypes.c_size_t, vp]
    L.fhmc_mu_cells_bytes.restype = ctypes.c_size_t
    L.fhmc_mu_cells_bytes.argtypes = [ctypes.POINTER(HistDesc), ci]
    L.fhmc_mu_cells_build.restype = ci
    L.fhmc_mu_cells_build.argtypes = [ctypes.POINTER(HistDesc), vp, vp, ctypes.c_size_t, ci, ctypes.c_double, ctypes.c_double, vp]
    L.fhmc_mu_cells_build_for.restype = ci
    L.fhmc_mu_cells_build_for.argtypes = [ctypes.POINTER(HistDesc), vp, vp, ctypes.c_size_t, ci, vp, cll, vp]
    L.fhmc_phase_moments_dev.restype = ci
    L.fhmc_phase_moments_dev.argtypes = [vp, ci, vp, ci, vp, vp, vp, ci, vp, vp, vp]
    L.fhmc_scalar_point.restype = ci
    L.fhmc_scalar_point.argtypes = [ctypes.POINTER(HistDesc), ctypes.POINTER(ScalarIO), ci, ci, vp]
    L.fhmc_lean_stats.restype = ci
    L.fhmc_lean_stats.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ci]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, load().fhmc_last_error().decode("utf-8", "replace")))


def last_kernel():
    """Name of the sweep / solver kernel this thread launched last (diagnostic, fhmc_last_kernel)."""
    return load().fhmc_last_kernel().decode("ascii", "replace")


def lean_stats(reset=False):
    """Outcome counters of the lean evaluator (see fhmc_lean_stats in include/fhmc_b200.h) as a list of 8 ints."""
    buf = (ctypes.c_ulonglong * 24)()     # (8 counters; 8 more cycle counters when built with -DFHMC_LEAN_PROFILE)
    check(load().fhmc_lean_stats(buf, 1 if reset else 0), "fhmc_lean_stats")
    return [int(x) for x in buf]

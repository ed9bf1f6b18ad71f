"""CPU: the C-ABI library loads and exports every symbol include/fhmc_b200.h declares; host-side packing logic;
the built-in HDF5 reader; product code never touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fhmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fhmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fhmcanalysis_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libfhmc_b200.so not built (python -m fhmcanalysis_b200.build)"
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 11
    for name in names:
        assert hasattr(L, name), "missing export %s" % name
    assert sorted(_lib.EXPORTS) == names
    assert _lib.load().fhmc_version() == 1


def test_struct_layouts_match_header():
    from fhmcanalysis_b200 import _lib
    # sizes implied by the header: 4+8+8+2+4+8 ints + 4 ints + 4 doubles + hull_row/hull_len/mu_recurrence/min_width + mu_tables + mu_cells
    assert ctypes.sizeof(_lib.HistDesc) == (4 + 8 + 8 + 2 + 4 + 8 + 4) * 4 + 4 * 8 + 4 * 4 + 8 + 8
    assert ctypes.sizeof(_lib.States) == 10 * 8
    assert ctypes.sizeof(_lib.SweepOut) == 9 * 8


def test_no_cpu_fallback_without_gpu():
    import torch
    from fhmcanalysis_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.DeviceHistogram(np.zeros(8), np.arange(8), 1.0, 0.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fhmcanalysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("parity oracle", "").lower() or f == "build.py" and False, \
                    "%s mentions the oracle" % os.path.join(dirpath, f)


def test_hdf5_reader_against_reference_loader(golden, golden_meta):
    """T1:46-66 (test_load): shapes and values of unittests/reference/test.nc."""
    path = "/root/reference/unittests/reference/test.nc"
    if not os.path.exists(path):
        pytest.skip("/root/reference not present on this machine")
    from fhmcanalysis_b200.io.hdf5_min import Dataset
    d = Dataset(path)
    lnpi = np.array(d.variables["ln(PI)"][:], dtype=np.float64)
    assert lnpi.shape == (31,) and abs(lnpi[1] - 11.579287195849) < 1e-9
    assert np.array_equal(lnpi, golden["testnc/lnpi"])
    assert np.array_equal(d.variables["N_{tot}"][:], np.arange(31))
    assert d.variables["N_{i}^{j}*N_{k}^{m}*U^{p}"][:].shape == (2, 3, 2, 3, 3, 31)
    assert np.array_equal(d.variables["N_{i}^{j}*N_{k}^{m}*U^{p}"][:], golden["testnc/mom"])
    assert d.variables["P_{N_i}(N_{tot})"][:].shape == (2, 31, 122)
    assert d.variables["P_{U}(N_{tot})"][:].shape == (31, 122)
    assert int(d.nspec) == 2 and int(d.max_order) == 2 and float(d.volume) == 729.0
    assert d.history == golden_meta["testnc"]["history"]
    # superblock-v0 example composite
    sw = "/root/reference/example/ntot/square_well/T_0.90/composite.nc"
    d2 = Dataset(sw)
    assert np.array_equal(d2.variables["ln(PI)"][:], golden["sw/lnpi"])

"""Stand-in for the absent ``netCDF4`` package, used ONLY when the compiled reference oracle
(oracle/_ref) is imported: ``Dataset`` is served by the repo's minimal HDF5 reader so that the
reference's own ``histogram.reload`` (gc_hist.pyx:143-182) can read the fixture files."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from fhmcanalysis_b200.io.hdf5_min import Dataset  # noqa: E402,F401

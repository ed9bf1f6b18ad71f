#!/usr/bin/env python
"""Build the UNMODIFIED-ALGORITHM reference (jeetain/FHMCAnalysis) as a parity oracle / CPU baseline.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fhmcanalysis_b200/`` imports this; only ``tests/``,
``__graft_entry__`` and ``bench.py`` (cpu_baseline / --impl reference legs) may use the output.

What it does: copies the reference Cython sources
  moments/histogram/one_dim/ntot/gc_hist.pyx, gc_binary.pyx, moments/histogram/two_dim/joint_hist.pyx,
  moments/histogram/one_dim/n1/gc_hist.pyx (built as module ``gc_hist_n1``)
from ``/root/reference`` (or $FHMC_REFERENCE) into a scratch directory, applies the minimal
Python-2 -> Python-3 / NumPy-2 / Cython-3 *porting* edits listed in ``PATCHES`` below (none of
them touches arithmetic), cythonizes them with ``language_level=2`` and drops ONLY the compiled
extension modules into ``oracle/_ref/`` (git-ignored, shipped to the GPU box by gpurun).
No reference source is copied into the repository.

The build system of the reference (numpy.distutils, Python 2) is not runnable, hence this recipe.
If /root/reference is absent (GPU box) the script is a no-op and the prebuilt .so files are used.
"""

import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("FHMC_REFERENCE", "/root/reference")

SOURCES = {
    "gc_hist": "moments/histogram/one_dim/ntot/gc_hist.pyx",
    "gc_binary": "moments/histogram/one_dim/ntot/gc_binary.pyx",
    "joint_hist": "moments/histogram/two_dim/joint_hist.pyx",
    "gc_hist_n1": "moments/histogram/one_dim/n1/gc_hist.pyx",      # N_1 order parameter
    "pore_hist": "moments/histogram/two_dim/h_ntot/pore_hist.pyx",  # 2-D normalise / masked averages
    "fhmc_patch": "moments/win_patch/fhmc_patch.pyx",               # window patching shift solve
}

# (module, regex, replacement, expected count or None) -- porting edits only.
PATCHES = [
    # Py2 3-argument unbound-method binding (gc_hist.pyx:2565-2566) -> plain def wrappers
    ("gc_hist", r"histogram\._cy_normalize = types\.MethodType\(_cython_normalize, None, histogram\)",
     "def _py_normalize(self):\n\t_cython_normalize(self)\nhistogram._cy_normalize = _py_normalize", 1),
    ("gc_hist", r"histogram\._cy_reweight = types\.MethodType\(_cython_reweight, None, histogram\)",
     "def _py_reweight(self, mu1_new):\n\t_cython_reweight(self, mu1_new)\nhistogram._cy_reweight = _py_reweight", 1),
    # Cython 3 cannot auto-wrap a cdef function with a buffer argument as a Python callable for fmin
    ("gc_hist", r"cdef double phase_eq_error \(", "def phase_eq_error (", 1),
    # NumPy 2 removed the np.int / np.float aliases
    ("gc_hist", r"dtype=np\.int\)", "dtype=np.int64)", None),
    ("gc_hist", r"np\.float, np\.float64", "float, np.float64", 1),
    ("joint_hist", r"dtype=np\.int\)", "dtype=np.int64)", None),
]
# window patching (fhmc_patch.pyx:636, 640, 668, 472-473): the same Py2 method binding; the fmin objective and
# patch_window_pair are cdef functions -> def so that tests can drive them; np.float alias in window.__init__
PATCHES += [
    ("fhmc_patch", r"window\._cy_normalize = types\.MethodType\(_cython_normalize_lnPI, None, window\)",
     "def _py_normalize_lnPI(self):\n\t_cython_normalize_lnPI(self)\nwindow._cy_normalize = _py_normalize_lnPI", 1),
    ("fhmc_patch", r"cdef double window_patch_error \(", "def window_patch_error (", 1),
    ("fhmc_patch", r"cdef patch_window_pair \(", "def patch_window_pair (", 1),
    ("fhmc_patch", r"dtype=np\.float,", "dtype=float,", 2),
]
# the N_1 module carries the same Py2/NumPy-1 idioms as ntot/gc_hist.pyx (n1/gc_hist.pyx:1734-1735, 1739, 156, 112)
PATCHES += [("gc_hist_n1",) + p[1:] for p in PATCHES if p[0] == "gc_hist"]

SETUP = r'''
import numpy as np
from setuptools import setup, Extension
from Cython.Build import cythonize
exts = [Extension(n, [n + ".pyx"], include_dirs=[np.get_include()], libraries=["m"],
                  define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
                  extra_compile_args=["-O2", "-w"]) for n in %r]
setup(name="fhmc_ref", ext_modules=cythonize(exts, language_level=2, quiet=True,
      compiler_directives={"cpow": True}))
'''


def have_prebuilt():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT, n + suffix)) for n in SOURCES)


def build(force=False):
    """Compile the reference extension modules into oracle/_ref/. Returns True when available."""
    if not os.path.isdir(REF):
        return have_prebuilt()
    if have_prebuilt() and not force:
        newest_src = max(os.path.getmtime(os.path.join(REF, p)) for p in SOURCES.values())
        suffix = sysconfig.get_config_var("EXT_SUFFIX")
        oldest_so = min(os.path.getmtime(os.path.join(OUT, n + suffix)) for n in SOURCES)
        if oldest_so > max(newest_src, os.path.getmtime(os.path.abspath(__file__))):
            return True
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="fhmc_ref_build_")
    try:
        for mod, rel in SOURCES.items():
            with open(os.path.join(REF, rel), "r") as fh:
                src = fh.read()
            for pmod, pat, rep, count in PATCHES:
                if pmod != mod:
                    continue
                src, n = re.subn(pat, rep.replace("\\", "\\\\"), src)
                if (count is not None and n != count) or n == 0:
                    raise RuntimeError("patch %r matched %d times in %s" % (pat, n, rel))
            with open(os.path.join(work, mod + ".pyx"), "w") as fh:
                fh.write(src)
        with open(os.path.join(work, "setup.py"), "w") as fh:
            fh.write(SETUP % (sorted(SOURCES),))
        env = dict(os.environ)
        env["PYTHONPATH"] = os.path.join(HERE, "stubs") + os.pathsep + env.get("PYTHONPATH", "")
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace", "-q"], cwd=work, env=env)
        suffix = sysconfig.get_config_var("EXT_SUFFIX")
        for mod in SOURCES:
            shutil.copy2(os.path.join(work, mod + suffix), os.path.join(OUT, mod + suffix))
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref %s" % ("ready" if ok else "UNAVAILABLE (no /root/reference and no prebuilt .so)"))
    sys.exit(0 if ok else 1)

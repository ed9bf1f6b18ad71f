"""In-tree build of libfhmc_b200.so: nvcc, sm_100a only (`python -m fhmcanalysis_b200.build`)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = ["fhmc_b200.cu", "fhmc_solver.cu", "fhmc_solver_lean.cu", "fhmc_solver_lean2.cu", "fhmc_2d.cu", "fhmc_fast_taylor.cu", "fhmc_rowc.cu", "fhmc_fast_rec.cu", "fhmc_fast_prod.cu", "fhmc_fast_prod_compact.cu", "fhmc_tab.cu", "fhmc_cell.cu", "fhmc_masked2d.cu", "fhmc_host_pipe.cu", "fhmc_patch.cu"]
HEADERS = ["fhmc_common.cuh", "fhmc_point.cuh", "fhmc_fast.cuh", "fhmc_prod.cuh", "fhmc_tab.cuh", "fhmc_solver.cuh", "fhmc_lean.cuh", "fhmc_solver_lean.cuh"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _includes(path, seen=None):
    """Headers under csrc/ (transitively) included by ``path``: a source is rebuilt only when one of ITS headers changed."""
    seen = set() if seen is None else seen
    csrc = os.path.join(HERE, "csrc")
    try:
        text = open(path).read()
    except OSError:
        return seen
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("#include \"") and line.endswith("\""):
            h = os.path.join(csrc, line.split("\"")[1])
            if os.path.exists(h) and h not in seen:
                seen.add(h)
                _includes(h, seen)
    return seen


def build(force=False, verbose=False):
    """Compile every CUDA source to an object (in parallel) and link the shared library."""
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("FHMC_NVCC_FLAGS", "").split()   # e.g. -DFHMC_LEAN_PROFILE (cycle counters in the lean evaluator)
    csrc = os.path.join(HERE, "csrc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    common = [os.path.join(ROOT, "include", "fhmc_b200.h"), os.path.abspath(__file__)]
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [os.path.join(csrc, src)] + sorted(_includes(os.path.join(csrc, src))) + common):
            cmd = [nvcc] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(csrc, src), "-o", obj]
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    lib = os.path.join(HERE, "libfhmc_b200.so")
    if procs or not os.path.exists(lib):
        subprocess.check_call([nvcc, "-shared", "-cudart", "static", "-o", lib] + objs)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

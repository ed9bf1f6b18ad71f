"""CPU: the built-in HDF5 writer.  lookup3 is pinned by checksums found in the reference's own
fixture files (bytes committed in tests/golden/hdf5_checksums.json, made by tests/golden/make_golden.py); written files
round-trip through the built-in reader, and through ``histogram(fname, ...)`` exactly like test_load (T1:46-66)."""
import json
import os

import numpy as np
import pytest

from fhmcanalysis_b200.io import hdf5_min as h5

HERE = os.path.dirname(os.path.abspath(__file__))


def test_lookup3_against_reference_file_checksums():
    vec = json.load(open(os.path.join(HERE, "golden", "hdf5_checksums.json")))
    assert len(vec) >= 6
    for v in vec:
        assert h5.lookup3(bytes.fromhex(v["bytes"])) == v["checksum"], v["what"]
    # Jenkins' published self-test vectors for hashlittle()
    assert h5.lookup3(b"") == 0xDEADBEEF
    assert h5.lookup3(b"Four score and seven years ago", 0) == 0x17770551
    assert h5.lookup3(b"Four score and seven years ago", 1) == 0xCD628161


def test_written_headers_carry_valid_checksums(tmp_path):
    p = str(tmp_path / "x.h5")
    h5.write_hdf5(p, {"a": np.arange(5.0), "b": np.arange(6, dtype=np.int64).reshape(2, 3)}, {"note": "hi", "k": 3})
    buf = open(p, "rb").read()
    assert h5.lookup3(buf[:44]) == int.from_bytes(buf[44:48], "little")
    pos, seen = 0, 0
    while True:
        pos = buf.find(b"OHDR", pos)
        if pos < 0:
            break
        n = int.from_bytes(buf[pos + 6:pos + 10], "little")
        end = pos + 10 + n
        assert h5.lookup3(buf[pos:end]) == int.from_bytes(buf[end:end + 4], "little")
        seen += 1
        pos = end
    assert seen == 3
    assert int.from_bytes(buf[28:36], "little") == len(buf)          # end-of-file address


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32, np.uint32, np.uint8])
def test_roundtrip_dtypes_and_shapes(tmp_path, dtype):
    rng = np.random.default_rng(5)
    p = str(tmp_path / "t.h5")
    v = {"s": (rng.random(()) * 100).astype(dtype), "v": (rng.random(7) * 100).astype(dtype),
         "m": (rng.random((3, 4, 5)) * 100).astype(dtype), "e": np.zeros((0,), dtype=dtype),
         "ln(PI)": (rng.random(11) * 100).astype(dtype), "N_{i}^{j}*N_{k}^{m}*U^{p}": (rng.random((2, 3)) * 9).astype(dtype)}
    h5.write_hdf5(p, v, {"history": "made by a test", "volume": 729.0, "nspec": 2, "flag": True})
    f = h5.File(p)
    assert sorted(f.variables) == sorted(v)
    for k in v:
        a = f.variables[k].read()
        assert a.dtype == np.dtype(dtype) and a.shape == v[k].shape and np.array_equal(a, v[k]), k
    assert f.attrs == {"history": "made by a test", "volume": 729.0, "nspec": 2, "flag": 1}


def test_many_variables(tmp_path):
    p = str(tmp_path / "many.h5")
    v = {"var_%03d" % i: np.full(i % 5 + 1, float(i)) for i in range(120)}
    h5.write_hdf5(p, v)
    f = h5.File(p)
    assert len(f.variables) == 120 and all(np.array_equal(f.variables[k].read(), v[k]) for k in v)


def test_composite_roundtrip_through_histogram_loader(tmp_path, golden):
    """Write the T1 fixture's arrays (golden, from the reference loader) and reload them the way test_load does."""
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram
    p = str(tmp_path / "composite.nc")
    lnpi, mom = golden["testnc/lnpi"], golden["testnc/mom"]
    h5.write_composite(p, lnpi, np.arange(31), mom, 729.0, 2, 2, history="Created by test",
                       histograms={"P_{U}(N_{tot})": np.ones((31, 4)), "P_{U}(N_{tot})_{lb}": np.zeros(31),
                                   "P_{U}(N_{tot})_{ub}": np.ones(31), "P_{U}(N_{tot})_{bw}": np.full(31, 0.25)})
    hist = histogram(p, 1.0, [5.0, 0.0], 1)
    assert np.array_equal(hist.data["ln(PI)"], lnpi) and np.array_equal(hist.data["mom"], mom)
    assert hist.data["ntot"].dtype == np.int64 and hist.data["lb"] == 0 and hist.data["ub"] == 30
    assert hist.data["volume"] == 729.0 and hist.data["max_order"] == 2 and hist.metadata["file_history"] == "Created by test"
    assert hist.data["e_hist"]["hist"].shape == (31, 4) and hist.data["pk_hist"] == {}
    d = h5.Dataset(p)
    assert d.variables["i"][:].tolist() == [1, 2] and d.variables["p"][:].tolist() == [0, 1, 2]   # FP:586-591
    # second generation: histogram.to_nc -> reload
    p2 = str(tmp_path / "again.nc")
    hist.to_nc(p2)
    h2 = histogram(p2, 1.0, [5.0, 0.0], 1)
    assert np.array_equal(h2.data["ln(PI)"], lnpi) and np.array_equal(h2.data["mom"], mom)
    assert np.array_equal(h2.data["e_hist"]["bw"], np.full(31, 0.25))
    with pytest.raises(AssertionError):
        histogram(p2, 1.0, [5.0], 1)                                 # nspec mismatch, GH:149


def test_write_composite_rejects_bad_shapes(tmp_path):
    with pytest.raises(ValueError):
        h5.write_composite(str(tmp_path / "bad.nc"), np.zeros(5), np.arange(5), np.zeros((1, 3, 1, 3, 3, 4)), 1.0, 1, 2)


def test_results_roundtrip(tmp_path):
    from fhmcanalysis_b200 import engine
    rng = np.random.default_rng(0)
    host = {"status": rng.integers(0, 2**31, 9).astype(np.uint32), "code": np.zeros(9, np.int32), "safe": rng.random(9) < 0.5,
            "fe": rng.normal(size=(9, 4)), "avg": None, "bounds": rng.integers(0, 1000, (9, 4, 2)).astype(np.int32)}
    p = str(tmp_path / "res.h5")
    engine.save_results(p, host, {"beta": 1.25})
    back, attrs = engine.load_results(p)
    assert attrs["beta"] == 1.25 and attrs["producer"] == "fhmcanalysis_b200"
    for k in ("status", "code", "safe", "fe", "bounds"):
        assert np.array_equal(back[k], host[k]) and back[k].dtype == host[k].dtype, k


def test_every_example_composite_reads_and_round_trips(tmp_path):
    """All twelve composites under the reference's example/ directory (written by netCDF4 / libhdf5) through the built-in
    reader, then through write_composite and the reader again: identical arrays and attributes.  (Needs /root/reference;
    the GPU box skips it -- the vectors recorded from these files travel in tests/golden/examples_vectors.npz.)"""
    import glob
    from fhmcanalysis_b200.io.hdf5_min import Dataset, write_composite
    files = sorted(glob.glob("/root/reference/example/ntot/*/*/composite.nc") + glob.glob("/root/reference/example/ntot/*/*/*/composite.nc"))
    if not files:
        pytest.skip("/root/reference not present on this machine")
    assert len(files) == 12
    for k, f in enumerate(files):
        d = Dataset(f)
        lnpi = np.array(d.variables["ln(PI)"][:], dtype=np.float64)
        ntot = np.array(d.variables["N_{tot}"][:])
        mom = np.array(d.variables["N_{i}^{j}*N_{k}^{m}*U^{p}"][:], dtype=np.float64)
        nspec, max_order = int(d.nspec), int(d.max_order)
        assert ntot.dtype == np.int64 and np.array_equal(ntot, np.arange(len(lnpi)))
        assert mom.shape == (nspec, max_order + 1, nspec, max_order + 1, max_order + 1, len(lnpi))
        assert np.all(mom[0, 0, 0, 0, 0] == 1.0) and np.all(np.isfinite(lnpi)) and float(d.volume) > 0
        out = str(tmp_path / ("c%d.nc" % k))
        write_composite(out, lnpi, ntot, mom, float(d.volume), nspec, max_order, history=str(getattr(d, "history", "")))
        e = Dataset(out)
        assert np.array_equal(np.array(e.variables["ln(PI)"][:]), lnpi)
        assert np.array_equal(np.array(e.variables["N_{tot}"][:]), ntot)
        assert np.array_equal(np.array(e.variables["N_{i}^{j}*N_{k}^{m}*U^{p}"][:]), mom)
        assert float(e.volume) == float(d.volume) and int(e.nspec) == nspec and int(e.max_order) == max_order

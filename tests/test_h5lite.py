"""The h5py-free reader of the reference's training-set files (py_rmpe_server/h5lite.py):
against a file written by the HDF5 library itself, against the committed 3-sample fixture, and -- through the
REFERENCE's own RawDataIterator.read_data with this module standing in for h5py -- against the drop-in class."""
import importlib
import json
import os
import sys
import types

import numpy as np
import pytest

from cases import ROOT

import rmpe_b200  # noqa: F401  (registers the package under its importable alias)

h5lite = importlib.import_module("adapting-rgb-pose-estimation-to-new-domains_b200.py_rmpe_server.h5lite")
FIXTURE = os.path.join(ROOT, "tests", "golden", "datum_3samples.h5")


def _fixture_samples():
    spec = importlib.util.spec_from_file_location("make_h5_fixture", os.path.join(ROOT, "tests", "golden", "make_h5_fixture.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.samples()


def test_reads_a_file_written_by_the_hdf5_library():
    """scipy ships a MATLAB v7.3 test file = an HDF5 file (512-byte user block, superblock 0, symbol-table root group,
    version 1 object header, contiguous f64 data, fixed-length string attribute) written by libhdf5 1.6."""
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    with h5lite.File(path) as f:
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        assert d.attrs["MATLAB_class"] == "double"
        assert np.array_equal(d[()].ravel(), np.arange(9) * (np.pi / 4))


def test_fixture_round_trip():
    want = _fixture_samples()
    with h5lite.File(FIXTURE) as f:
        g = f["datum"]
        assert g.keys() == sorted(want) and len(g) == 3 and "0000001" in g
        for key, (arr, meta) in want.items():
            d = g[key]
            assert d.shape == arr.shape and d.dtype == np.uint8
            assert np.array_equal(d[()], arr) and np.array_equal(d.value, arr)
            assert "meta" in d.attrs and json.loads(d.attrs["meta"]) == meta
        with pytest.raises(KeyError):
            g["9999999"]


def test_many_keys_and_long_strings(tmp_path):
    rng = np.random.RandomState(1)
    samples = {"%07d" % i: (rng.randint(0, 256, size=(6, 8 + i % 5, 9 + i % 7)).astype(np.uint8),
                            {"joints": [[[1.0, 2.0, float(i % 3)]] * 17], "objpos": [[1, 2]], "scale_provided": [0.5],
                             "img_path": "p" * (i * 37 % 5000)}) for i in range(200)}
    path = str(tmp_path / "many.h5")
    h5lite.write_datum_file(path, samples)
    with h5lite.File(path) as f:
        g = f["datum"]
        assert g.keys() == sorted(samples)
        for key in ("0000000", "0000077", "0000199"):
            assert np.array_equal(g[key][()], samples[key][0])
            assert json.loads(g[key].attrs["meta"]) == samples[key][1]


def test_not_hdf5_and_chunked_are_reported(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(h5lite.H5FormatError):
        h5lite.File(str(p))


def test_raw_data_iterator_opens_the_fixture_like_the_reference():
    """RawDataIterator(h5file).read_data on the fixture: the drop-in (h5lite underneath) against the reference's own class
    (py_rmpe_data_iterator.py:13-19, 46-66) importing this reader under the name h5py."""
    from oracle import ref_shim
    it = rmpe_b200.data_iterator.RawDataIterator(FIXTURE, shuffle=False, augment=False)
    assert it.num_keys() == 3
    want = _fixture_samples()
    for key, (arr, meta) in want.items():
        img, mask_miss, m = it.read_data(key)
        assert np.array_equal(img, arr[0:3].transpose(1, 2, 0)) and np.array_equal(mask_miss, arr[4])
        assert m["objpos"] == meta["objpos"] and m["scale_provided"] == meta["scale_provided"]
        assert m["joints"].shape == (len(meta["joints"]), 18, 3)
    if not ref_shim.available():
        return
    ref_shim.load()
    had = sys.modules.get("h5py")
    sys.modules["h5py"] = types.SimpleNamespace(File=h5lite.File)
    try:
        sys.modules.pop("py_rmpe_server.py_rmpe_data_iterator", None)
        ref_it_mod = importlib.import_module("py_rmpe_server.py_rmpe_data_iterator")
        ref_it = ref_it_mod.RawDataIterator(FIXTURE, shuffle=False, augment=False)
        assert ref_it.num_keys() == it.num_keys()
        for key in want:
            a, b = ref_it.read_data(key), it.read_data(key)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            assert np.array_equal(a[2]["joints"], b[2]["joints"])
            assert a[2]["objpos"] == b[2]["objpos"] and a[2]["scale_provided"] == b[2]["scale_provided"]
    finally:
        if had is None:
            sys.modules.pop("h5py", None)
        else:
            sys.modules["h5py"] = had


def test_round_trip_property(tmp_path):
    """hypothesis: any set of keys, array shapes and (unicode, long, empty) meta strings survives write -> read."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    key_st = st.text(alphabet="0123456789abcdefXYZ_-", min_size=1, max_size=12)
    sample_st = st.tuples(st.integers(1, 6), st.integers(1, 9), st.integers(1, 11), st.integers(0, 2 ** 31 - 1),
                          st.text(max_size=300))
    counter = [0]

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.dictionaries(key_st, sample_st, min_size=1, max_size=40))
    def run(spec):
        samples = {}
        for k, (c, h, w, seed, text) in spec.items():
            arr = np.random.RandomState(seed).randint(0, 256, size=(c, h, w)).astype(np.uint8)
            samples[k] = (arr, json.dumps({"t": text, "n": seed}))
        counter[0] += 1
        path = str(tmp_path / ("p%d.h5" % counter[0]))
        h5lite.write_datum_file(path, samples)
        with h5lite.File(path) as f:
            g = f["datum"]
            assert g.keys() == sorted(samples, key=lambda s: s.encode("utf-8")) or sorted(g.keys()) == sorted(samples)
            for k, (arr, meta) in samples.items():
                d = g[k]
                assert d.shape == arr.shape and np.array_equal(d[()], arr)
                assert d.attrs["meta"] == meta

    run()

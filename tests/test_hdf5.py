"""DOLFIN-HDF5 ingestion: the numpy reader against the reference's own mesh file, and a
write -> read round trip of a tetrahedral mesh, its facet markers and a stored field through the HDF5File facade
(sandbox/brain-simulations/MPET-4networks-colin27.py:28-45,197-199,257-259).  Host logic only: runs without a GPU."""
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference/src/mpet/test/donut2D.h5"


def _hdf5_module():
    """waterscapes_b200.mpet.hdf5 without importing the package __init__ (which loads the CUDA library)."""
    import importlib.util
    root = os.path.join(os.path.dirname(__file__), "..", "waterscapes_b200", "mpet")
    if "waterscapes_b200.mpet" in sys.modules:
        from waterscapes_b200.mpet import hdf5
        return hdf5
    spec = importlib.util.spec_from_file_location("_wb_hdf5", os.path.join(root, "hdf5.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reader_against_golden_copy_of_the_reference_file(golden_dir, tmp_path):
    """tests/golden/donut2D.npz holds the arrays of the reference's donut2D.h5; where the reference checkout is
    present the parser must reproduce them from the file itself (a real DOLFIN-written HDF5 file)."""
    h = _hdf5_module()
    z = np.load(os.path.join(golden_dir, "donut2D.npz"))
    if os.path.exists(REF):
        r = h.H5Reader(REF)
        assert r.keys("/") == ["mesh"] and set(r.keys("/mesh")) >= {"coordinates", "topology"}
        assert np.array_equal(r.read("/mesh/coordinates"), z["coordinates"])
        assert np.array_equal(r.read("/mesh/topology"), z["topology"])
        assert r.attrs("/mesh/topology")["celltype"] == "triangle"
    # the same arrays through our writer and back
    path = str(tmp_path / "donut_copy.h5")
    h.write_h5(path, {"/mesh/coordinates": z["coordinates"], "/mesh/topology": (z["topology"], {"celltype": "triangle"})})
    r = h.H5Reader(path)
    assert np.array_equal(r.read("/mesh/coordinates"), z["coordinates"])
    assert np.array_equal(r.read("/mesh/topology"), z["topology"])
    assert r.attrs("/mesh/topology")["celltype"] == "triangle"


def test_round_trip_mesh_markers_and_vectors(tmp_path):
    h = _hdf5_module()
    rng = np.random.default_rng(0)
    tree = {"/mesh/coordinates": rng.standard_normal((50, 3)), "/mesh/topology": (rng.integers(0, 50, (120, 4)), {"celltype": "tetrahedron"}),
            "/boundaries/values": rng.integers(0, 3, 77), "/boundaries/topology": rng.integers(0, 50, (77, 3)),
            "/u/vector_0": (rng.standard_normal(31), {"timestamp": 0.125}), "/u/vector_1": (rng.standard_normal(31), {"timestamp": 0.25}),
            "/empty": np.zeros(0)}
    path = str(tmp_path / "t.h5")
    h.write_h5(path, tree)
    r = h.H5Reader(path)
    assert r.keys("/") == ["boundaries", "empty", "mesh", "u"] and r.keys("/u") == ["vector_0", "vector_1"]
    for k, v in tree.items():
        arr = v[0] if isinstance(v, tuple) else v
        got = r.read(k)
        assert got.shape == arr.shape and np.array_equal(got, arr), k
    assert abs(r.attrs("/u/vector_1")["timestamp"] - 0.25) < 1e-15
    assert r.has("/mesh/topology") and not r.has("/mesh/nope")

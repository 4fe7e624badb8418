"""The oracle against its committed golden outputs (tests/golden/oracle_cube.npz, written by
tests/golden/make_oracle_fixture.py): pattern and dof map bit-exact, values / right-hand side / solution to
round-off.  Freezes the oracle between rounds; the GPU twin of this test is in test_gpu_golden.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_oracle_fixture import case                                  # noqa: E402
from oracle.mpet import MPETOracle, MPETTotalPressureOracle           # noqa: E402


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("tag,cls", [("std", MPETOracle), ("tp", MPETTotalPressureOracle)])
def test_oracle_matches_golden(tag, cls, golden_dir):
    G = np.load(os.path.join(golden_dir, "oracle_cube.npz"))
    o, A, P, b, dofs, x = case(cls, 1, 0.0)
    assert np.array_equal(A.indptr, G[tag + "1_indptr"])
    assert np.array_equal(A.indices, G[tag + "1_indices"])
    assert np.array_equal(o.space.cell_dofs, G[tag + "1_cell_dofs"])
    assert np.array_equal(dofs, G[tag + "1_dofs"])
    assert _rel(A.data, G[tag + "1_A"]) < 1e-14
    assert _rel(P.data, G[tag + "1_P"]) < 1e-14
    assert _rel(b, G[tag + "1_b"]) < 1e-13
    assert _rel(x, G[tag + "1_x"]) < 1e-10
    o, A, P, b, dofs, x = case(cls, 2, 0.2)
    sums = np.array([A.nnz, np.abs(A.data).sum(), np.linalg.norm(A.data), np.abs(P.data).sum(),
                     np.linalg.norm(b), np.linalg.norm(x), float(dofs.sum())])
    assert np.allclose(sums, G[tag + "2_sums"], rtol=1e-10, atol=0)

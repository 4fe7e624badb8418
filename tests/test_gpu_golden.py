"""GPU against the committed golden vectors (tests/golden/oracle_cube.npz): pattern and dof map bit-exact,
assembled A and P <= 1e-12 relative Frobenius, both formulations, through the C-ABI."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.mesh import unit_cube_mesh
from tests.golden.make_oracle_fixture import PARAMS


@pytest.mark.parametrize("tag", ["std", "tp"])
def test_gpu_matches_golden(tag, golden_dir):
    from waterscapes_b200.engine import Engine
    G = np.load(os.path.join(golden_dir, "oracle_cube.npz"))
    mesh = unit_cube_mesh(1)
    tp = tag == "tp"
    eng = Engine(0)
    eng.set_mesh(mesh.coords, mesh.cells.astype(np.int32), PARAMS["J"] + (1 if tp else 0))
    args = (PARAMS["E"], PARAMS["nu"], PARAMS["alpha"], PARAMS["K"], PARAMS["S"], PARAMS["c"], 0.1, 0.5)
    (eng.set_params_total_pressure if tp else eng.set_params)(*args)
    rowptr, cols = eng.pattern()
    assert np.array_equal(rowptr.cpu().numpy(), G[tag + "1_indptr"])
    assert np.array_equal(cols.cpu().numpy(), G[tag + "1_indices"])
    assert np.array_equal(eng.cell_dofs().cpu().numpy(), G[tag + "1_cell_dofs"])
    eng.assemble_lhs()
    a = eng.values(0).cpu().numpy()
    assert np.linalg.norm(a - G[tag + "1_A"]) / np.linalg.norm(G[tag + "1_A"]) < 1e-12
    eng.assemble_prec()
    p = eng.values(3).cpu().numpy()
    assert np.linalg.norm(p - G[tag + "1_P"]) / np.linalg.norm(G[tag + "1_P"]) < 1e-12
    eng.close()

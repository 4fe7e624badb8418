"""GPU parity: device-built dof map / CSR pattern (bit-exact) and assembled values
(<= 1e-12 relative Frobenius) against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.mesh import unit_cube_mesh
from oracle.mpet import MPETOracle, Coef

PARAMS = {
    1: dict(J=1, E=500.0, nu=0.49, alpha=(1.0,), c=(1e-2,), K=(1e-5,), S=((0.0,),)),
    2: dict(J=2, E=2.2, nu=0.4545454545, alpha=(0.5, 0.5), c=(1.0, 1.0), K=(1.0, 1.0), S=((0, 1.0), (1.0, 0))),
    3: dict(J=3, E=1500.0, nu=0.4999, alpha=(0.49, 0.25, 0.01), c=(3.9e-4, 2.9e-4, 1.5e-5),
            K=(1.573e-5, 3.745e-2, 3.745e-2), S=((0, 0, 1e-6), (0, 0, 0), (1e-6, 0, 0))),
    4: dict(J=4, E=1500.0, nu=0.4999, alpha=(0.49, 0.25, 0.01, 0.25), c=(3.9e-4, 2.9e-4, 1.5e-5, 2.9e-4),
            K=(1.573e-5, 3.745e-2, 3.745e-2, 3.745e-2),
            S=((0, 0, 1e-6, 1e-6), (0, 0, 0, 1e-6), (1e-6, 0, 0, 1e-6), (1e-6, 1e-6, 1e-6, 0))),
}


def _engine(mesh, params, dt, theta):
    from waterscapes_b200.engine import Engine
    eng = Engine(0)
    eng.set_mesh(mesh.coords, mesh.cells.astype(np.int32), params["J"])
    eng.set_params(params["E"], params["nu"], params["alpha"], params["K"], params["S"], params["c"], dt, theta)
    return eng


@pytest.mark.parametrize("n,A,theta,jitter", [(3, 2, 1.0, 0.0), (4, 1, 0.5, 0.2), (5, 4, 0.5, 0.2), (8, 2, 1.0, 0.0),
                                              (2, 3, 0.5, 0.2), (1, 2, 1.0, 0.0)])
def test_pattern_dofmap_values(n, A, theta, jitter):
    mesh = unit_cube_mesh(n, jitter=jitter)
    params = PARAMS[A]
    dt = 0.05
    o = MPETOracle(mesh, params, dt=dt, theta=theta)
    eng = _engine(mesh, params, dt, theta)
    s = eng.sizes
    assert s["N"] == o.space.N and s["Ne"] == o.space.Ne
    # closed-form sizes of SURVEY.md section 8 (exact for UnitCubeMesh)
    nnz22 = 230 * n ** 3 + 138 * n ** 2 + 24 * n + 1
    nnz21 = 65 * n ** 3 + 57 * n ** 2 + 15 * n + 1
    nnz11 = 15 * n ** 3 + 21 * n ** 2 + 9 * n + 1
    assert (s["nnz22"], s["nnz21"], s["nnz11"]) == (nnz22, nnz21, nnz11)
    assert s["nnz"] == 9 * nnz22 + 6 * A * nnz21 + A * A * nnz11
    # dof map: bit-exact
    assert np.array_equal(eng.edges().cpu().numpy(), o.space.edge_vertices)
    assert np.array_equal(eng.cell_dofs().cpu().numpy(), o.space.cell_dofs)
    # CSR pattern: bit-exact
    rowptr, cols = eng.pattern()
    ip, ix = o.pattern()
    assert np.array_equal(rowptr.cpu().numpy(), ip)
    assert np.array_equal(cols.cpu().numpy(), ix)
    # values
    eng.assemble_lhs()
    vals = eng.values(0).cpu().numpy()
    Ao = o.on_pattern(o.assemble_lhs())
    rel = np.linalg.norm(vals - Ao.data) / np.linalg.norm(Ao.data)
    assert rel < 1e-12, rel
    # determinism: a second assembly is bit-identical
    eng.assemble_lhs()
    assert np.array_equal(eng.values(0).cpu().numpy(), vals)
    # preconditioner blocks
    eng.assemble_prec()
    pv = eng.values(3).cpu().numpy()
    Po = o.on_pattern(o.assemble_prec())
    assert np.linalg.norm(pv - Po.data) / np.linalg.norm(Po.data) < 1e-12
    eng.close()


def test_spmv_rhs_bc_export():
    n, A, theta, dt = 4, 2, 0.5, 0.1
    mesh = unit_cube_mesh(n, jitter=0.2)
    params = PARAMS[2]
    o = MPETOracle(mesh, params, dt=dt, theta=theta)
    eng = _engine(mesh, params, dt, theta)
    eng.assemble_lhs()
    N = o.space.N
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N)
    Ao = o.assemble_lhs()
    xd = torch.as_tensor(x, device="cuda")
    yd = torch.empty_like(xd)
    eng.spmv(xd, yd)
    y = yd.cpu().numpy()
    assert np.linalg.norm(y - Ao @ x) / np.linalg.norm(Ao @ x) < 1e-13
    # previous-state load L = rhs(F)
    o.up_ = x.copy()
    bo = o.assemble_L()
    bd = torch.empty_like(xd)
    eng.rhs_prev(xd, bd)
    assert np.linalg.norm(bd.cpu().numpy() - bo) / np.linalg.norm(bo) < 1e-13
    # mass operators / lumped loads
    w2 = eng.lumped(2).cpu().numpy()
    w1 = eng.lumped(1).cpu().numpy()
    f = Coef(value=(1.0, 0.0, 0.0))
    bo = o._cell_load(f, 0.0, "u")
    assert np.allclose(w2, bo[: o.space.N2], rtol=1e-12, atol=1e-15)
    bo = o._cell_load(Coef(value=1.0), 0.0, ("p", 0))
    assert np.allclose(w1, bo[o.space.p_dofs(0)], rtol=1e-12, atol=1e-15)
    # Dirichlet export variants
    o.momentum_markers[:] = 0
    o.continuity_markers[0][:] = 0
    dofs, _ = o.dirichlet(0.0)
    eng.set_dirichlet_dofs(dofs.astype(np.int32))
    v1 = eng.values(1).cpu().numpy()
    v2 = eng.values(2).cpu().numpy()
    A1 = o.on_pattern(o.apply_bc_matrix(o.on_pattern(Ao), dofs))
    A2 = o.on_pattern(o.apply_bc_symmetric(o.on_pattern(Ao), dofs))
    assert np.linalg.norm(v1 - A1.data) / np.linalg.norm(A1.data) < 1e-12
    assert np.linalg.norm(v2 - A2.data) / np.linalg.norm(A2.data) < 1e-12
    eng.close()

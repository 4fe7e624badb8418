"""CPU: host-side logic of the reference-facing API (no kernels): meshes, markers, expressions,
boundary dof sets, facet operators -- checked against the oracle's independent implementation."""
import numpy as np
import pytest

from oracle.mesh import unit_cube_mesh
from oracle.mpet import MPETOracle, Coef
from waterscapes_b200.mpet.dolfin_shim import (UnitCubeMesh, BoxMesh, Constant, Expression, CompiledSubDomain,
                                               MeshFunction, FacetNormal, NormalProduct, Parameters)
from waterscapes_b200.mpet.mpetproblem import MPETProblem, convert_to_E_nu, convert_to_mu_lmbda
from waterscapes_b200.workloads import sizes


def test_box_mesh_matches_oracle_generator():
    for n in (1, 2, 5):
        m = UnitCubeMesh(n)
        o = unit_cube_mesh(n)
        assert np.array_equal(m.cells, o.cells)
        assert np.allclose(m.coordinates, o.coords, atol=0)
    F = UnitCubeMesh(3).exterior_facets()
    Fo = unit_cube_mesh(3).exterior_facets()
    for k in ("vertices", "cell", "local"):
        assert np.array_equal(F[k], Fo[k])
    assert F["cell"].shape[0] == 6 * 2 * 9


def test_expression_strings_and_time_constant():
    t = Constant(0.25)
    e = Expression("mmHg2Pa*(3.0 + 2*sin(2*pi*t)) + x[0]*x[1]", mmHg2Pa=133.32, t=t, degree=1)
    x = np.array([[0.5, 2.0, 0.0], [1.0, 1.0, 1.0]])
    assert np.allclose(e.eval_points(x), 133.32 * (3 + 2 * np.sin(np.pi / 2)) + x[:, 0] * x[:, 1])
    t.assign(0.5)
    assert np.allclose(e.eval_points(x), 133.32 * 3 + x[:, 0] * x[:, 1])
    v = Expression(("x[0]", "pow(x[1], 2)", "t"), t=t, degree=2)
    assert v.value_shape() == (3,) and v.eval_points(x).shape == (2, 3)
    assert isinstance(e * FacetNormal(None), NormalProduct)


def test_subdomain_marking_and_problem_defaults():
    mesh = UnitCubeMesh(3)
    problem = MPETProblem(mesh, Constant(0.0), params=dict(J=2, E=1.0, nu=0.3, alpha=(1, 1), K=(1, 1),
                                                           S=((0, 0), (0, 0)), c=(1, 1)))
    import sys
    assert np.all(problem.momentum_boundary_markers.array() == sys.maxsize)
    CompiledSubDomain("on_boundary").mark(problem.momentum_boundary_markers, 0)
    CompiledSubDomain("on_boundary && near(x[0], 1.0)").mark(problem.momentum_boundary_markers, 1)
    arr = problem.momentum_boundary_markers.array()
    assert (arr == 1).sum() == 2 * 9 and (arr == 0).sum() == 5 * 2 * 9
    CompiledSubDomain("on_boundary && x[2] > 0.5 && x[1] < 0.5").mark(problem.continuity_boundary_markers[1], 2)
    assert (problem.continuity_boundary_markers[1].array() == 2).sum() > 0
    assert len(problem.g) == 2 and float(problem.g[0]) == 0.0
    E, nu = convert_to_E_nu(1.0, 10.0)
    mu, lm = convert_to_mu_lmbda(E, nu)
    assert abs(mu - 1.0) < 1e-14 and abs(lm - 10.0) < 1e-13


def test_sizes_formula():
    s = sizes(71, 4)
    assert s["dofs"] == 10265613 and s["nnz"] == 1400017525 and s["cells"] == 2147466
    p = Parameters("x"); p.add("dt", 0.1); p.update(dict(dt=0.2)); assert p["dt"] == 0.2


def test_elastic_stress_of_a_quadratic_field():
    """mpetproblem.py:18-24 mirror: sigma = 2 mu eps(u) + lambda div(u) I, exact for P2 displacements."""
    from waterscapes_b200.mpet.dolfin_shim import FunctionSpace, SubFunction
    from waterscapes_b200.mpet.mpetproblem import elastic_stress
    mesh = UnitCubeMesh(2)
    space = FunctionSpace.from_host(mesh, 1)
    X = space.node2_coordinates()
    u = np.stack([X[:, 0] ** 2 + 2 * X[:, 1], X[:, 1] * X[:, 2], 3 * X[:, 2] - X[:, 0]], axis=1)
    E, nu = 10.0, 0.3
    mu, lm = convert_to_mu_lmbda(E, nu)
    sig = elastic_stress(SubFunction(space, 0, u), E, nu).cell_values()
    xc = mesh.coordinates[mesh.cells].mean(axis=1)
    grad = np.zeros((xc.shape[0], 3, 3))
    grad[:, 0, 0] = 2 * xc[:, 0]; grad[:, 0, 1] = 2.0
    grad[:, 1, 1] = xc[:, 2]; grad[:, 1, 2] = xc[:, 1]
    grad[:, 2, 2] = 3.0; grad[:, 2, 0] = -1.0
    eps = 0.5 * (grad + np.swapaxes(grad, 1, 2))
    ref = 2 * mu * eps + lm * np.trace(grad, axis1=1, axis2=2)[:, None, None] * np.eye(3)
    assert np.allclose(sig, ref, rtol=1e-12, atol=1e-12)


def test_user_subdomain_inside_reference_style():
    """A Python ``SubDomain`` written the DOLFIN way (``x[0]`` = first coordinate of ONE point, ``and`` / ``or`` on
    coordinates, src/mpet/demo/*: ``near(x[0], 0.0) and on_boundary``) marks the same facets as the equivalent
    CompiledSubDomain string (it used to index the first POINT)."""
    from waterscapes_b200.mpet.dolfin_shim import SubDomain, near

    class Left(SubDomain):                       # vectorises: evaluated on all points at once
        def inside(self, x, on_boundary):
            return on_boundary and near(x[0], 0.0)

    class LeftLower(SubDomain):                  # `and` on two coordinate tests: only works point by point
        def inside(self, x, on_boundary):
            return near(x[0], 0.0) and x[2] < 0.5 + 1e-12

    mesh = UnitCubeMesh(4)
    a, b = MeshFunction("size_t", mesh, 2, 0), MeshFunction("size_t", mesh, 2, 0)
    Left().mark(a, 1)
    CompiledSubDomain("on_boundary && near(x[0], 0.0)").mark(b, 1)
    assert a.array_.sum() == 2 * 16 and np.array_equal(a.array_, b.array_)
    a, b = MeshFunction("size_t", mesh, 2, 0), MeshFunction("size_t", mesh, 2, 0)
    LeftLower().mark(a, 1)
    CompiledSubDomain("near(x[0], 0.0) && x[2] < 0.5 + 1e-12").mark(b, 1)
    assert a.array_.sum() == 2 * 8 and np.array_equal(a.array_, b.array_)


@pytest.mark.parametrize("degree", [0, 1, 3, 4])
def test_cell_load_operator_matches_oracle_quadrature(degree):
    """Expression(degree=d) source data (the reference's MMS tests use d = 3, test_convergence_mpetsolver.py:129-131):
    the host-built cell-lattice operator times the lattice values == the oracle's int f.v dx / int g q dx with the
    coefficient interpolated cell-wise into P_d and integrated by a rule exact for degree (test + d)."""
    import scipy.sparse as sps
    from waterscapes_b200.mpet.dolfin_shim import FunctionSpace, Mesh
    from waterscapes_b200.mpet.la import CellLoadOperator
    om = unit_cube_mesh(3, jitter=0.2)
    o = MPETOracle(om, dict(J=1, E=1.0, nu=0.3, alpha=(1.0,), c=(1.0,), K=(1.0,), S=((0.0,),)), dt=0.1, theta=1.0)
    space = FunctionSpace.from_host(Mesh(om.coords, om.cells), 1)
    f = lambda x, t=0.0: np.stack([np.sin(2 * x[:, 0]) * x[:, 1], np.exp(x[:, 2]) + x[:, 0] ** 3,
                                   np.cos(x[:, 0] + 2 * x[:, 1] * x[:, 2])], axis=1)
    g = lambda x, t=0.0: np.sin(3 * x[:, 0]) * np.cos(x[:, 1]) + x[:, 2] ** 4

    class E:                                     # stands in for dolfin_shim.Expression (only eval_points is used)
        def __init__(self, fn):
            self.fn = fn

        def eval_points(self, x):
            return self.fn(x)

    def apply(op, data, nrows):
        rp, ci, v = (t.numpy() for t in op.csr)
        return sps.csr_matrix((v, ci, rp), shape=(nrows, data.shape[-1])) @ data

    op2 = CellLoadOperator(space, True, degree, device="cpu")
    d2 = op2.data(E(f))
    ref = o._cell_load(Coef(fn=lambda x, t: f(x), degree=degree), 0.0, "u")
    for k in range(3):
        got = apply(op2, d2[k], space.N2)
        want = ref[o.space.u_dofs(k)]
        assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want), (degree, k)
    op1 = CellLoadOperator(space, False, degree, device="cpu")
    got = apply(op1, op1.data(E(g)), space.Nv)
    want = o._cell_load(Coef(fn=lambda x, t: g(x), degree=degree), 0.0, ("p", 0))[o.space.p_dofs(0)]
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want), degree


def test_rigid_motions_match_oracle_and_are_l2_orthonormal():
    """waterscapes_b200/mpet/rm_basis_L2.py (host-side twin of rm_basis_L2.py:10-74) against the oracle's
    independent implementation, and the defining property: the six motions are orthonormal in L2(Omega)."""
    from waterscapes_b200.mpet.dolfin_shim import Mesh
    from waterscapes_b200.mpet.rm_basis_L2 import rigid_motions
    om = unit_cube_mesh(3, jitter=0.2)
    om.coords[:] = om.coords * np.array([3.0, 2.0, 1.0]) + np.array([0.5, -1.0, 2.0])
    o = MPETOracle(om, dict(J=1, E=1.0, nu=0.3, alpha=(1.0,), c=(1.0,), K=(1.0,), S=((0.0,),)), dt=0.1, theta=1.0)
    Z = rigid_motions(Mesh(om.coords, om.cells))
    Zo = o.rigid_motions()
    x = np.random.default_rng(2).uniform(0, 3, size=(50, 3))
    assert len(Z) == len(Zo) == 6
    for a, b in zip(Z, Zo):      # eigenvectors are defined up to sign
        va, vb = a(x), b(x)
        assert min(np.abs(va - vb).max(), np.abs(va + vb).max()) < 1e-10
    # Gram matrix with the degree-2 rule (the motions are affine: exact)
    xc = om.coords[om.cells]
    vol = om.cell_volumes()
    a_, b_ = 0.1381966011250105, 0.5854101966249685
    lam = np.array([[b_, a_, a_, a_], [a_, b_, a_, a_], [a_, a_, b_, a_], [a_, a_, a_, b_]])
    xq = np.einsum("qv,cvd->cqd", lam, xc).reshape(-1, 3)
    w = np.repeat(vol / 4.0, 4)
    vals = [z(xq) for z in Z]
    G = np.array([[np.sum(w * np.einsum("pd,pd->p", vi, vj)) for vj in vals] for vi in vals])
    assert np.abs(G - np.eye(6)).max() < 1e-12

"""The reference-facing API (MPETProblem / MPETSolver) on the GPU against the oracle's time loop."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.mesh import unit_cube_mesh
from oracle.mpet import MPETOracle, Coef

PI = np.pi


def _mms_problem(n, J, theta, dt, T, nonsym=False, total_pressure=False):
    from waterscapes_b200.mpet import (MPETProblem, MPETSolver, MPETTotalPressureSolver, UnitCubeMesh, Constant,
                                       Expression, CompiledSubDomain, FacetNormal)
    S = ((0.0, 2.0), (1.0, 0.0)) if nonsym else ((0.0, 1.0), (1.0, 0.0))
    params = dict(J=J, E=2.2, nu=0.4545, alpha=(0.5, 0.5), c=(1.0, 1.0), K=(1.0, 1.0), S=S)
    mesh = UnitCubeMesh(n)
    time = Constant(0.0)
    problem = MPETProblem(mesh, time, params=params)
    u_e = ("0.1*cos(pi*x[0])*sin(pi*x[1])*sin(pi*x[2])*sin(pi*t)",
           "0.1*sin(pi*x[0])*cos(pi*x[1])*sin(pi*x[2])*sin(pi*t)",
           "0.1*sin(pi*x[0])*sin(pi*x[1])*cos(pi*x[2])*sin(pi*t)")
    p_e = ["%d*sin(pi*x[0])*cos(pi*x[1])*sin(pi*x[2])*sin(2*pi*t)" % (i + 1) for i in range(J)]
    problem.u_bar = Expression(u_e, t=time, degree=2)
    problem.p_bar = [Expression(p_e[i], t=time, degree=1) for i in range(J)]
    problem.f = Expression(("sin(pi*x[0])*t", "x[1]*x[2]", "cos(t)*x[0]"), t=time, degree=2)
    problem.g = [Expression("(1+%d)*x[0]*x[1]*sin(t)" % i, t=time, degree=1) for i in range(J)]
    n_ = FacetNormal(mesh)
    problem.s = Expression("0.3*t*x[1]", t=time, degree=2) * n_
    problem.I = [Expression("0.2*x[2]*(1+t)", t=time, degree=1), Constant(0.1)]
    problem.beta = [Constant(0.0), Constant(0.7)]
    problem.p_robin = [Constant(0.0), Expression("0.5*x[0]+t", t=time, degree=1)]
    on_boundary = CompiledSubDomain("on_boundary")
    right = CompiledSubDomain("on_boundary && near(x[0], 1.0)")
    top = CompiledSubDomain("on_boundary && near(x[2], 1.0)")
    on_boundary.mark(problem.momentum_boundary_markers, 0)
    right.mark(problem.momentum_boundary_markers, 1)
    for i in range(J):
        on_boundary.mark(problem.continuity_boundary_markers[i], 0)
    right.mark(problem.continuity_boundary_markers[0], 1)
    top.mark(problem.continuity_boundary_markers[1], 2)
    cls = MPETTotalPressureSolver if total_pressure else MPETSolver
    solver = cls(problem, dict(dt=dt, theta=theta, T=T))
    return mesh, params, problem, solver


def _oracle_twin(n, params, theta, dt, T, total_pressure=False):
    from oracle.mpet import MPETTotalPressureOracle
    mesh = unit_cube_mesh(n)
    o = (MPETTotalPressureOracle if total_pressure else MPETOracle)(mesh, params, dt=dt, theta=theta, T=T)
    J = params["J"]
    o.u_bar = Coef(fn=lambda x, t: 0.1 * np.stack([np.cos(PI * x[:, 0]) * np.sin(PI * x[:, 1]) * np.sin(PI * x[:, 2]),
                                                   np.sin(PI * x[:, 0]) * np.cos(PI * x[:, 1]) * np.sin(PI * x[:, 2]),
                                                   np.sin(PI * x[:, 0]) * np.sin(PI * x[:, 1]) * np.cos(PI * x[:, 2])],
                                                  1) * np.sin(PI * t))
    o.p_bar = [Coef(fn=lambda x, t, i=i: (i + 1) * np.sin(PI * x[:, 0]) * np.cos(PI * x[:, 1]) * np.sin(PI * x[:, 2])
                    * np.sin(2 * PI * t)) for i in range(J)]
    o.f = Coef(fn=lambda x, t: np.stack([np.sin(PI * x[:, 0]) * t, x[:, 1] * x[:, 2], np.cos(t) * x[:, 0]], 1), degree=2)
    o.g = [Coef(fn=lambda x, t, i=i: (1 + i) * x[:, 0] * x[:, 1] * np.sin(t), degree=1) for i in range(J)]
    o.s = Coef(fn=lambda x, t: 0.3 * t * x[:, 1], degree=2)
    o.s_times_normal = True
    o.I = [Coef(fn=lambda x, t: 0.2 * x[:, 2] * (1 + t), degree=1), Coef(value=0.1)]
    o.beta = [Coef(value=0.0), Coef(value=0.7)]
    o.p_robin = [Coef(value=0.0), Coef(fn=lambda x, t: 0.5 * x[:, 0] + t, degree=1)]
    F = o.facets
    xm = mesh.coords[F["vertices"]]
    right = np.all(np.abs(xm[:, :, 0] - 1.0) < 3e-16, axis=1)
    top = np.all(np.abs(xm[:, :, 2] - 1.0) < 3e-16, axis=1)
    o.momentum_markers[:] = 0
    o.momentum_markers[right] = 1
    for i in range(J):
        o.continuity_markers[i][:] = 0
    o.continuity_markers[0][right] = 1
    o.continuity_markers[1][top] = 2
    return o


def _init(solver, o):
    rng = np.random.default_rng(3)
    x0 = 0.05 * rng.standard_normal(o.space.N)
    o.up_ = x0.copy()
    solver.up_.vector().set_local(x0)


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("theta,nonsym", [(1.0, False), (0.5, False), (0.5, True)])
def test_solve_matches_oracle_time_loop(theta, nonsym):
    n, J, dt, T = 4, 2, 0.1, 0.3
    mesh, params, problem, solver = _mms_problem(n, J, theta, dt, T, nonsym=nonsym)
    o = _oracle_twin(n, params, theta, dt, T)
    _init(solver, o)
    # the pieces first: A (with Robin), b of the first step
    from waterscapes_b200.mpet import assemble
    A = solver._assemble_system().to_scipy()
    Ao = o.on_pattern(o.assemble_lhs())
    assert _rel(A.data, Ao.data) < 1e-12
    bo, dofs, vals = o.rhs(0.0)
    bcs = solver.bcs[0] + solver.bcs[1]
    solver._sync_dirichlet(bcs)
    b = solver._rhs(problem.time, 0.0, dt, theta, bcs).get_local()
    assert _rel(b, bo) < 1e-12, _rel(b, bo)
    problem.time.assign(0.0)
    ref = list((up.copy(), t) for up, t in o.solve_direct())
    k = 0
    for up, t in solver.solve():
        xo, to = ref[k]
        assert abs(t - to) < 1e-12
        x = up.vector().get_local()
        sp_ = o.space
        nu = 3 * sp_.N2
        errs = [_rel(x[:nu], xo[:nu])] + [_rel(x[sp_.p_dofs(i)], xo[sp_.p_dofs(i)]) for i in range(J)]
        assert max(errs) < 1e-8, (t, errs)
        k += 1
    assert k == len(ref) == 3
    print("niter", solver.solver_monitor["niter"])


def test_step_and_iterative_branch():
    n, J, dt, T, theta = 3, 2, 0.1, 0.2, 1.0
    mesh, params, problem, solver = _mms_problem(n, J, theta, dt, T)
    o = _oracle_twin(n, params, theta, dt, T)
    _init(solver, o)
    xo = o.step()
    solver.step(dt)
    x = solver.up.vector().get_local()
    assert _rel(x, xo) < 1e-8
    assert abs(float(problem.time) - dt) < 1e-14
    # iterative branch with PETSc-default tolerance: converges, error consistent with rtol
    mesh, params, problem, solver = _mms_problem(n, J, theta, dt, T)
    solver.params["direct_solver"] = False
    o = _oracle_twin(n, params, theta, dt, T)
    _init(solver, o)
    ref = [up.copy() for up, t in o.solve_direct()]
    outs = [up.vector().get_local() for up, t in solver.solve()]
    assert len(outs) == 2
    assert _rel(outs[-1], ref[-1]) < 1e-3
    assert all(nit > 0 for nit in solver.solver_monitor["niter"])
    u, p0, p1 = solver.up.split(deepcopy=True)
    assert u.values.shape == (o.space.N2, 3) and p0.values.shape == (o.space.Nv,)
    assert abs(p0((0.5, 0.5, 0.5)) - float(p0.values[np.argmin(np.linalg.norm(mesh.coordinates - 0.5, axis=1))])) < 0.5


def test_dg0_permeability_matches_oracle():
    """K_0 as a DG0 (cell-wise constant) function, the reference's
    sandbox/biot-robin/three_fields_precond.py:263-271, through MPETSolver: assembled A, preconditioner blocks,
    previous-state load (theta = 0.5 puts K into the right-hand side as well) and two time steps against the
    oracle's cell-wise restatement."""
    from waterscapes_b200.mpet import (MPETProblem, MPETSolver, UnitCubeMesh, Constant, Expression, CompiledSubDomain,
                                       CellFunction, assemble)
    n, J, theta, dt, T = 4, 2, 0.5, 0.1, 0.2
    mesh = UnitCubeMesh(n)
    rng = np.random.default_rng(11)
    kvals = 10.0 ** rng.uniform(-3, 0, mesh.num_cells())          # three decades of contrast
    base = dict(J=J, E=2.2, nu=0.4545, alpha=(0.5, 0.5), c=(1.0, 1.0), S=((0.0, 1.0), (1.0, 0.0)))
    time = Constant(0.0)
    problem = MPETProblem(mesh, time, params=dict(base, K=(CellFunction(mesh, kvals), 0.3)))
    problem.p_bar = [Expression("%d*sin(pi*x[0])*sin(2*pi*t)" % (i + 1), t=time, degree=1) for i in range(J)]
    problem.g = [Expression("(1+%d)*x[0]*x[1]*cos(t)" % i, t=time, degree=1) for i in range(J)]
    on_boundary = CompiledSubDomain("on_boundary")
    on_boundary.mark(problem.momentum_boundary_markers, 0)
    for i in range(J):
        on_boundary.mark(problem.continuity_boundary_markers[i], 0)
    solver = MPETSolver(problem, dict(dt=dt, theta=theta, T=T))
    o = MPETOracle(unit_cube_mesh(n), dict(base, K=(kvals, 0.3)), dt=dt, theta=theta, T=T)
    o.p_bar = [Coef(fn=lambda x, t, i=i: (i + 1) * np.sin(PI * x[:, 0]) * np.sin(2 * PI * t)) for i in range(J)]
    o.g = [Coef(fn=lambda x, t, i=i: (1 + i) * x[:, 0] * x[:, 1] * np.cos(t), degree=1) for i in range(J)]
    o.momentum_markers[:] = 0
    for i in range(J):
        o.continuity_markers[i][:] = 0
    _init(solver, o)
    A = solver._assemble_system().to_scipy()
    assert _rel(A.data, o.on_pattern(o.assemble_lhs()).data) < 1e-12
    P = assemble(solver.prec).to_scipy()
    Po = o.on_pattern(o.assemble_prec())
    assert _rel(P.data, Po.data) < 1e-12                       # no Dirichlet set pushed yet: P as assembled
    bd = torch.empty(o.space.N, dtype=torch.float64, device="cuda")
    solver.engine.rhs_prev(solver.up_.x, bd)
    assert _rel(bd.cpu().numpy(), o.assemble_L()) < 1e-12
    ref = [up.copy() for up, t in o.solve_direct()]
    for k, (up, t) in enumerate(solver.solve()):
        x = up.vector().get_local()
        sp_ = o.space
        nu = 3 * sp_.N2
        errs = [_rel(x[:nu], ref[k][:nu])] + [_rel(x[sp_.p_dofs(i)], ref[k][sp_.p_dofs(i)]) for i in range(J)]
        assert max(errs) < 1e-8, (k, errs)
    # back to a constant: the weighted stiffness is dropped again
    problem.params["K"] = (0.05, 0.3)
    A2 = solver._assemble_system().to_scipy()
    o2 = MPETOracle(unit_cube_mesh(n), dict(base, K=(0.05, 0.3)), dt=dt, theta=theta, T=T)
    assert _rel(A2.data, o2.on_pattern(o2.assemble_lhs()).data) < 1e-12


def test_hdf5_mesh_and_marker_ingestion(tmp_path):
    """a mesh and its facet markers written in DOLFIN's HDF5 layout, read back the way the
    reference's brain script does (MPET-4networks-colin27.py:28-45), and a step solved on the ingested mesh
    reproduces the step on the original one; fields are stored with HDF5File.write(u, "/u", t) (:257-259)."""
    from waterscapes_b200.mpet import (MPETProblem, MPETSolver, Mesh, MeshFunction, BoxMesh, Constant, Expression,
                                       CompiledSubDomain, HDF5File)
    from waterscapes_b200.mpet.hdf5 import H5Reader
    mesh0 = BoxMesh((0.0, 0.0, 0.0), (2.0, 1.0, 1.0), 4, 3, 3)
    bnd0 = MeshFunction("size_t", mesh0, 2)
    bnd0.set_all(1)
    CompiledSubDomain("on_boundary && near(x[0], 0.0)").mark(bnd0, 2)
    path = str(tmp_path / "box_boundaries.h5")
    with HDF5File(None, path, "w") as f:
        f.write(mesh0, "/mesh")
        f.write(bnd0, "/boundaries")

    def run(mesh, boundaries):
        time = Constant(0.0)
        params = dict(J=1, E=500.0, nu=0.35, alpha=(1.0,), c=(1e-2,), K=(1e-2,), S=((0.0,),))
        problem = MPETProblem(mesh, time, params=params)
        problem.p_bar = [Expression("x[0]*t", t=time, degree=1)]
        arr = boundaries.array()
        problem.momentum_boundary_markers.array()[:] = np.where(arr == 2, 0, 1)      # clamp the x = 0 face
        problem.continuity_boundary_markers[0].array()[:] = 0
        solver = MPETSolver(problem, dict(dt=0.1, T=0.1, theta=1.0))
        for up, t in solver.solve():
            pass
        return solver, up

    mesh = Mesh()
    f = HDF5File(None, path, "r")
    f.read(mesh, "/mesh", False)
    boundaries = MeshFunction("size_t", mesh, 2)
    f.read(boundaries, "/boundaries")
    f.close()
    assert np.array_equal(mesh.cells, mesh0.cells) and np.array_equal(mesh.coordinates, mesh0.coordinates)
    assert np.array_equal(boundaries.array(), bnd0.array()) and set(np.unique(bnd0.array())) == {1, 2}
    s0, up0 = run(mesh0, bnd0)
    s1, up1 = run(mesh, boundaries)
    assert np.array_equal(up0.vector().get_local(), up1.vector().get_local())
    out = str(tmp_path / "u.h5")
    with HDF5File(None, out, "w") as g:
        g.write(up1.split(deepcopy=True)[0], "/u", 0.1)
        g.write(up1, "/up", 0.1)
    r = H5Reader(out)
    assert np.array_equal(r.read("/up/vector_0"), up1.vector().get_local())
    assert abs(r.attrs("/u/vector_0")["timestamp"] - 0.1) < 1e-15


@pytest.mark.parametrize("df,dg", [(3, 3), (1, 2)])
def test_source_data_of_other_degrees_matches_oracle(df, dg):
    """Expression(degree=d) source data with d != the test space's degree (the reference's MMS tests use d = 3,
    test_convergence_mpetsolver.py:129-131): DOLFIN interpolates it cell-wise into P_d and integrates exactly; the GPU
    path applies the host-built cell-lattice operator with mpet_csr_spmv.  b of one step vs the oracle <= 1e-12."""
    from waterscapes_b200.mpet import Expression
    n, J, theta, dt, T = 3, 2, 0.5, 0.1, 0.1
    mesh, params, problem, solver = _mms_problem(n, J, theta, dt, T)
    o = _oracle_twin(n, params, theta, dt, T)
    time = problem.time
    problem.f = Expression(("sin(2*x[0])*x[1]*(1+t)", "exp(x[2]) + pow(x[0], 3)", "cos(x[0] + 2*x[1]*x[2])"), t=time,
                           degree=df)
    problem.g = [Expression("(1+%d)*sin(3*x[0])*cos(x[1])*(1+t) + pow(x[2], 4)" % i, t=time, degree=dg) for i in range(J)]
    o.f = Coef(fn=lambda x, t: np.stack([np.sin(2 * x[:, 0]) * x[:, 1] * (1 + t), np.exp(x[:, 2]) + x[:, 0] ** 3,
                                         np.cos(x[:, 0] + 2 * x[:, 1] * x[:, 2])], 1), degree=df)
    o.g = [Coef(fn=lambda x, t, i=i: (1 + i) * np.sin(3 * x[:, 0]) * np.cos(x[:, 1]) * (1 + t) + x[:, 2] ** 4, degree=dg)
           for i in range(J)]
    _init(solver, o)
    solver._assemble_system()
    bo, dofs, vals = o.rhs(0.0)
    bcs = solver.bcs[0] + solver.bcs[1]
    solver._sync_dirichlet(bcs)
    b = solver._rhs(problem.time, 0.0, dt, theta, bcs).get_local()
    assert _rel(b, bo) < 1e-12, _rel(b, bo)
    solver.engine.close()

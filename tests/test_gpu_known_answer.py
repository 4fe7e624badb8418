"""Known-answer (patch) test on the GPU, independent of the oracle: the 3-D analogue of the reference's
``constant_on_the_donut`` (src/mpet/test/test_donut.py:16-71).  J = 1, c = 0, alpha = 1, K = 1e-2, E = 500,
nu = 0.35, dt = 0.1, T = 0.2, pressure data p_bar = -t on the whole boundary.  The exact -- and exactly
representable -- solution is u = 0, p = -t.  (The reference loads the boundary with the traction s = t*n and
removes the rigid motions with Lagrange multipliers; here the displacement is clamped instead,
which has the same solution.)  Assertions and tolerances are the reference's (test_donut.py:70-71)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("solver_name", ["MPETSolver", "MPETTotalPressureSolver"])
def test_constant_pressure_on_the_cube(solver_name):
    import waterscapes_b200.mpet as mpet
    from waterscapes_b200.mpet import (MPETProblem, UnitCubeMesh, Constant, Expression, CompiledSubDomain)
    mesh = UnitCubeMesh(4)
    time = Constant(0.0)
    material = dict(J=1, c=(0.0,), alpha=(1.0,), K=(1.0e-2,), S=((0.0,),), E=500.0, nu=0.35)
    problem = MPETProblem(mesh, time, params=material)
    problem.p_bar = [Expression("-t", t=time, degree=0)]
    on_boundary = CompiledSubDomain("on_boundary")
    on_boundary.mark(problem.momentum_boundary_markers, 0)
    on_boundary.mark(problem.continuity_boundary_markers[0], 0)
    solver = getattr(mpet, solver_name)(problem, dict(dt=0.1, T=0.2, theta=1.0, direct_solver=True))
    k = 0
    for up, t in solver.solve():
        k += 1
    assert k == 2 and abs(t - 0.2) < 1e-12
    fields = up.split(deepcopy=True)
    u, p = fields[0], fields[-1]
    assert abs(p((0.5, 0.5, 0.5)) + 0.2) < 1e-8                       # test_donut.py:70
    # ||p||_L2 / sqrt(vol) = 0.2 (test_donut.py:71); p is P1, so the L2 norm follows from the nodal values
    assert np.max(np.abs(p.values + 0.2)) < 1e-8
    assert np.max(np.abs(u.values)) < 1e-8 * 0.2


def _traction_problem(J, p_dirichlet):
    from waterscapes_b200.mpet import (MPETProblem, BoxMesh, Constant, Expression, CompiledSubDomain, FacetNormal)
    mesh = BoxMesh((0.0, 0.0, 0.0), (100.0, 80.0, 60.0), 4, 4, 3)
    time = Constant(0.0)
    K = (1.0e-2, 1.0e-1)[:J]
    material = dict(J=J, c=(0.0,) * J, alpha=(1.0,) * J, K=K, S=tuple((0.0,) * J for _ in range(J)), E=500.0, nu=0.35)
    problem = MPETProblem(mesh, time, params=material)
    problem.s = Expression("t", t=time, degree=0) * FacetNormal(mesh)
    problem.u_has_nullspace = True
    on_boundary = CompiledSubDomain("on_boundary")
    on_boundary.mark(problem.momentum_boundary_markers, 1)          # traction everywhere: u up to rigid motions
    for i in range(J):
        on_boundary.mark(problem.continuity_boundary_markers[i], 0 if p_dirichlet else 1)
    if p_dirichlet:
        problem.p_bar = [Expression("-t", t=time, degree=0) for i in range(J)]
    else:
        problem.p_has_nullspace = [True] * J
    return mesh, problem


def test_constant_pressure_with_rigid_motion_multipliers():
    """src/mpet/test/test_donut.py:16-71 with its own data on a 3-D box: traction s = t*n on the whole boundary,
    u_has_nullspace = True (six rigid-motion Lagrange multipliers, rm_basis_L2.py), p_bar = -t.  Exact solution
    u = 0, p = -t; the reference's assertions and tolerances."""
    from waterscapes_b200.mpet import MPETSolver
    mesh, problem = _traction_problem(1, True)
    solver = MPETSolver(problem, dict(dt=0.1, T=0.2, theta=1.0))
    for up, t in solver.solve():
        pass
    (u, p, r) = up.split(deepcopy=True)
    assert len(up.vector().get_local()) == solver.VQ.N + 6 and r.values.shape == (6,)
    assert abs(p((10.0, 50.0, 30.0)) + 0.2) < 1.e-8, "Point value of p not matching reference"      # test_donut.py:70
    # ||p||_L2 / sqrt(volume) = 0.2 (test_donut.py:71): P1 mass norm through the library's mass operator
    pv = torch.as_tensor(p.values, device="cuda")
    Mp = torch.zeros_like(pv)
    solver.engine.mass_apply(1, 1.0, pv, Mp)
    l2 = float(torch.sqrt(torch.dot(pv, Mp)))
    assert abs(l2 / np.sqrt(100.0 * 80.0 * 60.0) - 0.2) < 1.e-10
    assert np.max(np.abs(u.values)) < 1e-8 * 0.2 and np.max(np.abs(r.values)) < 1e-8
    print("iterations", solver.solver_monitor["niter"])


def test_all_neumann_two_networks_with_pressure_multipliers():
    """src/mpet/test/test_donut.py:73-129 (smoke in the reference: no assertion there): two networks, every
    boundary Neumann, rigid-motion multipliers for u and one constant multiplier per pressure.  Here the result is
    also checked: the constraints hold (u orthogonal to the rigid motions, zero-mean pressures)."""
    from waterscapes_b200.mpet import MPETSolver
    from waterscapes_b200.mpet.rm_basis_L2 import rigid_motions
    mesh, problem = _traction_problem(2, False)
    solver = MPETSolver(problem, dict(dt=0.1, T=0.2, theta=1.0))
    for up, t in solver.solve():
        pass
    fields = up.split(deepcopy=True)
    assert len(fields) == 1 + 2 + 1 + 2                                   # u, p1, p2, r, two pressure multipliers
    x = up.vector().get_local()
    assert np.all(np.isfinite(x)) and x.shape[0] == solver.VQ.N + 8
    eng, sp = solver.engine, solver.VQ
    lump = eng.lumped(1).cpu().numpy()
    scale = 0.2        # the traction at t = 0.2: the pressures themselves vanish up to round-off in this problem
    for i in (1, 2):
        assert abs(lump @ fields[i].values) < 1e-8 * scale * lump.sum()  # int p_i dx = 0
    Z = rigid_motions(mesh)
    x2 = sp.node2_coordinates()
    u = fields[0].values
    for Zi in Z:
        z = Zi(x2)
        tot = 0.0
        for k in range(3):
            out = torch.zeros(sp.N2, dtype=torch.float64, device="cuda")
            eng.mass_apply(2, 1.0, torch.as_tensor(np.ascontiguousarray(z[:, k]), device="cuda"), out)
            tot += float(out.cpu().numpy() @ u[:, k])
        assert abs(tot) < 1e-8 * max(np.abs(u).max(), 1e-300) * np.sqrt(100.0 * 80.0 * 60.0)
    # uniform compression by the traction t*n: div u = -t / (lambda + 2 mu / 3) everywhere
    print("max |u|", np.abs(u).max(), "iterations", solver.solver_monitor["niter"])
    assert np.abs(u).max() > 1e-3

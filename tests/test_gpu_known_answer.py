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

"""Oracle-free convergence test of the GPU path: the 3-D twin of the reference's MMS rate tests
(src/mpet/test/test_convergence_mpetsolver.py:103-245, test_convergence_totalpressuresolver.py:103-253), run
through MPETSolver / MPETTotalPressureSolver exactly as the reference tests drive their solvers (same material
parameters, same boundary split -- Neumann traction sigma.n on x = 1, Dirichlet elsewhere --, same time grids),
with the reference's rate thresholds (:234-239 and :237-242) asserted on the finest pair of the reference's own
mesh sequence n = 8, 16, 32 (the pair 8 -> 16 is still pre-asymptotic in 3-D: pressure L2 rates 1.6-1.7).
Nothing here touches oracle/: exact solutions and error norms come from tests/mms3d.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests import mms3d


def _single_run(n, M, theta, total_pressure, df=3, dg=3):
    import waterscapes_b200.mpet as mpet
    from waterscapes_b200.mpet import (MPETProblem, UnitCubeMesh, Constant, Expression, CompiledSubDomain, FacetNormal,
                                       interpolate, assign)
    T = 1.0
    dt = float(T / M)
    if total_pressure:      # test_convergence_totalpressuresolver.py:110-118
        params = dict(J=2, c=(1.0, 1.0), alpha=(1.0, 1.0), K=(1.0, 1.0), S=((1.0, 1.0), (1.0, 1.0)), E=1.0, nu=0.35)
    else:                   # test_convergence_mpetsolver.py:110-118
        params = dict(J=2, c=(0.3, 0.4), alpha=(0.4, 0.6), K=(0.2, 0.3), S=((0.0, 2.0), (1.0, 0.0)), E=520.0, nu=0.47)
    ex = mms3d.exact_solutions(params)
    mesh = UnitCubeMesh(n)
    time = Constant(0.0)
    problem = MPETProblem(mesh, time, params=params)
    problem.f = Expression(lambda x, t: ex["f"](x, t), t=time, degree=df)
    problem.f.value_shape = lambda: (3,)
    problem.g = [Expression(lambda x, t, i=i: ex["g"][i](x, t), t=time, degree=dg) for i in range(2)]
    problem.u_bar = Expression(lambda x, t: ex["u"](x, t), t=time, degree=3)
    problem.u_bar.value_shape = lambda: (3,)
    problem.p_bar = [Expression(lambda x, t, i=i: ex["p"][i](x, t), t=time, degree=3) for i in range(2)]
    sigma = Expression(lambda x, t: ex["sigma"](x, t), t=time, degree=3)
    sigma.value_shape = lambda: (3, 3)
    problem.s = sigma * FacetNormal(mesh)
    on_boundary = CompiledSubDomain("on_boundary")
    right = CompiledSubDomain("near(x[0], 1.0) && on_boundary")
    on_boundary.mark(problem.momentum_boundary_markers, 0)
    right.mark(problem.momentum_boundary_markers, 1)
    for i in range(2):
        on_boundary.mark(problem.continuity_boundary_markers[i], 0)
    cls = mpet.MPETTotalPressureSolver if total_pressure else mpet.MPETSolver
    solver = cls(problem, dict(dt=dt, theta=theta, T=T, direct_solver=True))
    VP = solver.up_.function_space()
    time.assign(0.0)
    assign(solver.up_.sub(0), interpolate(problem.u_bar, VP.sub(0).collapse()))
    off = 1
    if total_pressure:
        p0 = Expression(lambda x, t: ex["p_total"](x, t), t=time, degree=3)
        assign(solver.up_.sub(1), interpolate(p0, VP.sub(1).collapse()))
        off = 2
    for i in range(2):
        assign(solver.up_.sub(i + off), interpolate(problem.p_bar[i], VP.sub(i + off).collapse()))
    for up, t in solver.solve():
        pass
    fields = up.split(deepcopy=True)
    u = fields[0].values
    ps = [fields[i + off].values for i in range(2)]
    err = mms3d.error_norms(mesh.coordinates, mesh.cells, VP.edge_index, VP.Nv, u, ps, ex, float(t))
    its = solver.solver_monitor["niter"]
    return err, mms3d.hmin(mesh.coordinates, mesh.cells), (min(its), max(its))


@pytest.mark.parametrize("solver_kind", ["standard", "total_pressure"])
@pytest.mark.parametrize("theta,ns,ms,df,dg", [(0.5, [8, 16, 32], [4, 8, 16], 3, 3), (1.0, [8, 16, 32], [8, 32, 128], 2, 1)])
def test_mms_convergence_rates_3d(solver_kind, theta, ns, ms, df, dg):
    """Source data f, g: ``Expression(degree=3)`` as in the reference's tests (:129-131) on the Crank-Nicolson grid
    -- cell-wise P3 interpolation integrated exactly, 912 673 distinct lattice points per evaluation at n = 32 --;
    the 128-step implicit-Euler grid takes nodal (P2 / P1) data, which keeps its host-side evaluation of the
    sympy-derived sources 5x shorter (the thresholds are the reference's either way)."""
    tp = solver_kind == "total_pressure"
    res = [_single_run(n, m, theta, tp, df, dg) for n, m in zip(ns, ms)]
    hs = [r[1] for r in res]
    u_L2 = mms3d.rates([r[0]["u_L2"] for r in res], hs)
    u_H1 = mms3d.rates([r[0]["u_H1"] for r in res], hs)
    p_L2 = [mms3d.rates([r[0]["p_L2"][i] for r in res], hs) for i in range(2)]
    p_H1 = [mms3d.rates([r[0]["p_H1"][i] for r in res], hs) for i in range(2)]
    print(solver_kind, theta, "u_L2", u_L2, "u_H1", u_H1, "p_L2", p_L2, "p_H1", p_H1, "iterations", [r[2] for r in res])
    assert u_L2[-1] > 1.70 and u_H1[-1] > 1.70
    if tp:      # test_convergence_totalpressuresolver.py:237-242
        assert p_L2[0][-1] > 1.85 and p_L2[1][-1] > 1.87
    else:       # test_convergence_mpetsolver.py:234-239
        assert p_L2[0][-1] > 1.70 and p_L2[1][-1] > 1.70
    assert p_H1[0][-1] > 0.95 and p_H1[1][-1] > 0.95

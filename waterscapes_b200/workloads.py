"""The BASELINE.json configurations as MPETProblem builders (synthetic meshes).

cfg1  UnitCubeMesh(8),  J=2, material of demo/demo_adaptive.py:96-101, MMS-type data, Dirichlet everywhere
cfg2  UnitCubeMesh(34), J=1 Biot, material of demo/demo_donut.py:31-38, clamped bottom, pulsating top traction
cfg3  120 mm cube n=28, J=4, parameters/boundary data of sandbox/brain-simulations/MPET-4networks-colin27.py:67-160
cfg4  as cfg3 with n=71 (~10.3 M dofs)
cfg5  as cfg3 restricted to the first three networks, n=72 per GPU (~10.3 M dofs per GPU), weak-scaled
"""
import math

from .mpet import (MPETProblem, MPETSolver, BoxMesh, UnitCubeMesh, Constant, Expression, CompiledSubDomain,
                   FacetNormal, convert_to_E_nu)

MMHG2PA = 133.32   # MPET-4networks-colin27.py:112

CONFIGS = {
    "cfg1": dict(n=8, J=2, dt=0.1, theta=1.0),
    "cfg2": dict(n=34, J=1, dt=0.05, theta=1.0),
    "cfg3": dict(n=28, J=4, dt=0.0125, theta=0.5),
    "cfg4": dict(n=71, J=4, dt=0.0125, theta=0.5),
    "cfg5": dict(n=72, J=3, dt=0.0125, theta=0.5),
}


def brain_params(J):
    c = (3.9e-4, 2.9e-4, 1.5e-5, 2.9e-4)
    alpha = (0.49, 0.25, 0.01, 0.25)
    kappa = (1.4e-14, 1.e-10, 1.e-10, 1.e-10)
    eta = (8.9e-4, 2.67e-3, 2.67e-3, 2.67e-3)
    K = [kappa[i] / eta[i] * 1.e6 for i in range(4)]
    s = 1.0e-6
    S = ((0.0, 0.0, s, s), (0.0, 0.0, 0.0, s), (s, 0.0, 0.0, s), (s, s, s, 0.0))
    return dict(J=J, alpha=alpha[:J], K=K[:J], S=tuple(row[:J] for row in S[:J]), c=c[:J], nu=0.4999, E=1500)


def make_problem(name, n=None, mesh=None):
    """Return (problem, solver_params, initial_condition_callable)."""
    cfg = dict(CONFIGS[name])
    if n is not None:
        cfg["n"] = n
    n, J = cfg["n"], cfg["J"]
    time = Constant(0.0)
    on_boundary = CompiledSubDomain("on_boundary")
    if name == "cfg1":
        mesh = mesh or UnitCubeMesh(n)
        E, nu = convert_to_E_nu(1.0, 10.0)
        params = dict(J=2, E=E, nu=nu, alpha=(0.5, 0.5), c=(1.0, 1.0), K=(1.0, 1.0), S=((0.0, 1.0), (1.0, 0.0)))
        problem = MPETProblem(mesh, time, params=params)
        problem.u_bar = Expression(("0.1*cos(pi*x[0])*sin(pi*x[1])*sin(pi*x[2])*sin(pi*t)",
                                    "0.1*sin(pi*x[0])*cos(pi*x[1])*sin(pi*x[2])*sin(pi*t)",
                                    "0.1*sin(pi*x[0])*sin(pi*x[1])*cos(pi*x[2])*sin(pi*t)"), t=time, degree=2)
        problem.p_bar = [Expression("%d*sin(pi*x[0])*cos(pi*x[1])*sin(pi*x[2])*sin(2*pi*t)" % (i + 1), t=time,
                                    degree=1) for i in range(2)]
        # f, g derived with sympy from the exact (u, p) above so that cfg1 doubles as a convergence check
        from .mms import cfg1_exact, standard_sources
        mms = standard_sources(params, *cfg1_exact())
        problem.f = Expression(lambda x, t: mms["f"](x, t), t=time, degree=2)
        problem.g = [Expression(lambda x, t, i=i: mms["g"][i](x, t), t=time, degree=1) for i in range(2)]
        problem.mms = mms
        on_boundary.mark(problem.momentum_boundary_markers, 0)
        for i in range(2):
            on_boundary.mark(problem.continuity_boundary_markers[i], 0)
    elif name == "cfg2":
        mesh = mesh or UnitCubeMesh(n)
        params = dict(J=1, E=500.0, nu=0.49, alpha=(1.0,), c=(1.0e-2,), K=(1.0e-5,), S=((0.0,),))
        problem = MPETProblem(mesh, time, params=params)
        bottom = CompiledSubDomain("on_boundary && near(x[2], 0.0)")
        top = CompiledSubDomain("on_boundary && near(x[2], 1.0)")
        on_boundary.mark(problem.momentum_boundary_markers, 1)
        bottom.mark(problem.momentum_boundary_markers, 0)
        problem.s = Expression("mmHg2Pa*0.15*sin(2*pi*t)", mmHg2Pa=133.322, t=time, degree=0) * FacetNormal(mesh)
        on_boundary.mark(problem.continuity_boundary_markers[0], 1)
        top.mark(problem.continuity_boundary_markers[0], 0)
    else:
        mesh = mesh or BoxMesh((0.0, 0.0, 0.0), (120.0, 120.0, 120.0), n, n, n)
        params = brain_params(J)
        problem = MPETProblem(mesh, time, params=params)
        p_e = Expression("mmHg2Pa*(3.0 + 2*sin(2*pi*t))", mmHg2Pa=MMHG2PA, t=time, degree=0)
        p_a = Expression("mmHg2Pa*(70.0 + 10.0*sin(2.0*pi*t))", mmHg2Pa=MMHG2PA, t=time, degree=0)
        p_v = Constant(MMHG2PA * 6.0)
        p_c = Constant(MMHG2PA * (6.0 + 70) / 2)
        problem.p_bar = [p_e, p_a, p_v, p_c][:J]
        on_boundary.mark(problem.momentum_boundary_markers, 0)
        for i in range(J):
            on_boundary.mark(problem.continuity_boundary_markers[i], 1 if i == 3 else 0)
    solver_params = dict(dt=cfg["dt"], theta=cfg["theta"], T=1.0)

    def initial_condition(solver):
        time.assign(0.0)
        if name == "cfg1":
            solver.up_.set_sub(0, problem.u_bar)
        if name != "cfg2":
            for i in range(J):
                solver.up_.set_sub(i + 1, problem.p_bar[i])

    return problem, solver_params, initial_condition


def sizes(n, J):
    """Closed-form sizes of the [P2]^3 x [P1]^J system on an n^3 box mesh."""
    nnz22 = 230 * n ** 3 + 138 * n ** 2 + 24 * n + 1
    nnz21 = 65 * n ** 3 + 57 * n ** 2 + 15 * n + 1
    nnz11 = 15 * n ** 3 + 21 * n ** 2 + 9 * n + 1
    N = 3 * (2 * n + 1) ** 3 + J * (n + 1) ** 3
    return dict(cells=6 * n ** 3, dofs=N, nnz=9 * nnz22 + 6 * J * nnz21 + J * J * nnz11,
                nnz22=nnz22, nnz21=nnz21, nnz11=nnz11)

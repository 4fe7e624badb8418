"""Pin the CPU oracle against every known-answer test the reference holds for the
assemble+solve path:

* src/mpet/test/test_donut.py:16-71   -- exact constant solution on donut2D.h5
* src/mpet/test/test_donut.py:73-129  -- all-Neumann two-network smoke run
* src/mpet/test/test_convergence_mpetsolver.py:103-245 -- MMS convergence rates

The asserts and tolerances are the reference's own.
"""
import math
import os

import numpy as np
import pytest

from oracle.mesh import SimplexMesh, unit_square_mesh
from oracle.mpet import MPETOracle, Coef


def _donut(golden_dir):
    z = np.load(os.path.join(golden_dir, "donut2D.npz"))
    mesh = SimplexMesh(z["coordinates"], z["topology"])
    # the fixture really is the reference's annulus: chi = 0, area ~ pi (100^2 - 30^2)
    ev, _ = mesh.edges()
    assert mesh.num_vertices - ev.shape[0] + mesh.num_cells == 0
    assert abs(mesh.cell_volumes().sum() / (math.pi * (100 ** 2 - 30 ** 2)) - 1) < 1e-3
    return mesh


def _point_value_p1(mesh, p, pt):
    x = mesh.coords[mesh.cells]
    J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)
    X = np.einsum("cij,cj->ci", np.linalg.inv(J), pt - x[:, 0])
    c = np.nonzero((X.min(1) > -1e-12) & (X.sum(1) < 1 + 1e-12))[0][0]
    lam = np.r_[1 - X[c].sum(), X[c]]
    return lam @ p[mesh.cells[c]]


def _l2_norm_p1(mesh, p):
    vol = mesh.cell_volumes()
    d = mesh.dim
    M = (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
    pl = p[mesh.cells]
    return math.sqrt(np.einsum("c,cm,mn,cn->", vol, pl, M, pl))


def test_constant_on_the_donut(golden_dir):
    """test_donut.py:16-71."""
    mesh = _donut(golden_dir)
    params = dict(J=1, c=(0.0,), alpha=(1.0,), K=(1.0e-2,), S=((0.0,),), E=500, nu=0.35)
    o = MPETOracle(mesh, params, dt=0.1, theta=1.0, T=0.2, u_has_nullspace=True)
    o.s = Coef(value=lambda t: t)          # Expression("t")*n
    o.s_times_normal = True
    o.momentum_markers[:] = 1              # on_boundary -> NEUMANN
    o.continuity_markers[0][:] = 0         # on_boundary -> DIRICHLET
    o.p_bar = [Coef(value=lambda t: -t)]
    for up, t in o.solve_direct():
        pass
    u, p = o.split(up)
    assert len(up) == o.space.nfe + 3      # (u, p, r): three rigid motions in 2-D
    volume = math.sqrt(mesh.cell_volumes().sum())
    p_x = _point_value_p1(mesh, p[0], np.array([0.0, 50.0]))
    assert abs(p_x + 0.2) < 1.e-8, "Point value of p not matching reference"
    assert abs(_l2_norm_p1(mesh, p[0]) / volume - 0.2) < 1.e-10


def test_constant_on_the_donut_nullspaces(golden_dir):
    """test_donut.py:73-129 (smoke only in the reference: no assert there)."""
    mesh = _donut(golden_dir)
    params = dict(J=2, c=(0.0, 0.0), alpha=(1.0, 1.0), K=(1.0e-2, 1.0e-1),
                  S=((0.0, 0.0), (0.0, 0.0)), E=500, nu=0.35)
    o = MPETOracle(mesh, params, dt=0.1, theta=1.0, T=0.2, u_has_nullspace=True,
                   p_has_nullspace=(True, True))
    o.s = Coef(value=lambda t: t)
    o.s_times_normal = True
    o.momentum_markers[:] = 1
    o.continuity_markers[0][:] = 1
    norms = [np.linalg.norm(up) for up, t in o.solve_direct()]
    assert len(norms) == 2 and all(np.isfinite(norms))


# ------------------------------------------------------------------------- MMS
def _exact_solutions(params):
    """test_convergence_mpetsolver.py:21-101, returned as numpy callables."""
    import sympy
    J, nu, E = params["J"], params["nu"], params["E"]
    alpha, c, K, S = params["alpha"], params["c"], params["K"], params["S"]
    lmbda = nu * E / ((1.0 - 2.0 * nu) * (1.0 + nu))
    mu = E / (2.0 * (1.0 + nu))
    pi = math.pi
    omega = 2 * pi
    sin, diff = sympy.sin, sympy.diff
    x = sympy.symbols("x0 x1")
    t = sympy.symbols("t")
    u = [sin(2.0 * pi * x[0] + pi / 2.0) * sin(2.0 * pi * x[1] + pi / 2.0) * sin(omega * t + t)] * 2
    p = [0]
    for i in range(1, J + 1):
        p += [-(i) * sin(2.0 * pi * x[0] + pi / 2.0) * sin(2.0 * pi * x[1] + pi / 2.0) * sin(omega * t + t)]
    d = 2
    div_u = sum(diff(u[i], x[i]) for i in range(d))
    p[0] = lmbda * div_u - sum(alpha[i] * p[i + 1] for i in range(J))
    grad_u = [[diff(u[i], x[j]) for j in range(d)] for i in range(d)]
    eps_u = [[0.5 * (grad_u[i][j] + grad_u[j][i]) for j in range(d)] for i in range(d)]
    grad_p = [[diff(p[i], x[j]) for j in range(d)] for i in range(J + 1)]
    sigma_ast = [[2 * mu * eps_u[i][j] for j in range(d)] for i in range(d)]
    sigma = [[2 * mu * eps_u[i][j] + (p[0] if i == j else 0) for j in range(d)] for i in range(d)]
    div_sigma_ast = [sum(diff(sigma_ast[i][j], x[j]) for j in range(d)) for i in range(d)]
    f = [-(div_sigma_ast[j] + diff(p[0], x[j])) for j in range(d)]
    g = []
    for i in range(J):
        g.append(-c[i] * diff(p[i + 1], t)
                 - alpha[i] / lmbda * diff(p[0] + sum(alpha[j] * p[j + 1] for j in range(J)), t)
                 + sum(diff(K[i] * grad_p[i + 1][j], x[j]) for j in range(d))
                 - sum(S[i][j] * (p[i + 1] - p[j + 1]) for j in range(J)))

    def fn(expr):
        f_ = sympy.lambdify((x[0], x[1], t), expr, "numpy")
        return lambda X, tt: np.broadcast_to(f_(X[:, 0], X[:, 1], tt), (X.shape[0],)).astype(float)

    def vec(exprs):
        fs = [fn(e) for e in exprs]
        return lambda X, tt: np.stack([f_(X, tt) for f_ in fs], axis=1)

    def ten(exprs):
        fs = [[fn(e) for e in row] for row in exprs]
        return lambda X, tt: np.stack([np.stack([f_(X, tt) for f_ in row], axis=1) for row in fs], axis=1)

    return dict(u=vec(u), grad_u=ten(grad_u), p=[fn(pi_) for pi_ in p[1:]], p_total=fn(p[0]),
                grad_p=[vec(gp) for gp in grad_p[1:]], f=vec(f), g=[fn(gi) for gi in g],
                sigma=ten(sigma))


def _single_run(n, M, theta):
    """test_convergence_mpetsolver.py:103-179."""
    T = 1.0
    dt = float(T / M)
    params = dict(J=2, c=(0.3, 0.4), alpha=(0.4, 0.6), K=(0.2, 0.3),
                  S=((0.0, 2.0), (1.0, 0.0)), E=520.0, nu=0.47)
    ex = _exact_solutions(params)
    mesh = unit_square_mesh(n)
    o = MPETOracle(mesh, params, dt=dt, theta=theta, T=T)
    o.f = Coef(fn=ex["f"], degree=3)
    o.g = [Coef(fn=ex["g"][i], degree=3) for i in range(2)]
    o.u_bar = Coef(fn=ex["u"], degree=3)
    o.s = Coef(fn=ex["sigma"], degree=4)       # sigma_ex * normal
    o.s_times_normal = True
    o.p_bar = [Coef(fn=ex["p"][i], degree=3) for i in range(2)]
    F = o.facets
    o.momentum_markers[:] = 0
    xm = mesh.coords[F["vertices"]]
    right = np.all(np.abs(xm[:, :, 0] - 1.0) < 3e-16, axis=1)     # near(x[0], 1.0)
    o.momentum_markers[right] = 1
    for i in range(2):
        o.continuity_markers[i][:] = 0
    # initial conditions: interpolate exact solution at t = 0
    sp_ = o.space
    x2 = sp_.node2_coords()
    u0 = ex["u"](x2, 0.0)
    for k in range(2):
        o.up_[sp_.u_dofs(k)] = u0[:, k]
    for i in range(2):
        o.up_[sp_.p_dofs(i)] = ex["p"][i](mesh.coords, 0.0)
    for up, t in o.solve_direct():
        pass
    err = o.error_norms(up, ex["u"], ex["p"], t, qdeg=7, grad_u=ex["grad_u"], grad_p=ex["grad_p"])
    h = 2 * min(_circumradius(mesh))
    return err["u_L2"], err["u_H1"], err["p_L2"], err["p_H1"], h


def _circumradius(mesh):
    x = mesh.coords[mesh.cells]
    a = np.linalg.norm(x[:, 1] - x[:, 2], axis=1)
    b = np.linalg.norm(x[:, 0] - x[:, 2], axis=1)
    c = np.linalg.norm(x[:, 0] - x[:, 1], axis=1)
    return a * b * c / (4 * mesh.cell_volumes())


def _rates(errors, hs):
    return [math.log(errors[i + 1] / errors[i]) / math.log(hs[i + 1] / hs[i]) for i in range(len(hs) - 1)]


@pytest.mark.parametrize("theta,ns,ms", [(0.5, [8, 16, 32], [4, 8, 16]),
                                         (1.0, [8, 16, 32], [8, 32, 128])])
def test_convergence_mpetsolver(theta, ns, ms):
    """test_convergence_mpetsolver.py:181-245 (thresholds :234-239)."""
    res = [_single_run(n, m, theta) for n, m in zip(ns, ms)]
    hs = [r[4] for r in res]
    u_L2 = _rates([r[0] for r in res], hs)
    u_H1 = _rates([r[1] for r in res], hs)
    p0_L2 = _rates([r[2][0] for r in res], hs)
    p1_L2 = _rates([r[2][1] for r in res], hs)
    p0_H1 = _rates([r[3][0] for r in res], hs)
    p1_H1 = _rates([r[3][1] for r in res], hs)
    print(u_L2, u_H1, p0_L2, p1_L2, p0_H1, p1_H1)
    assert u_H1[-1] > 1.70
    assert u_L2[-1] > 1.70
    assert p0_L2[-1] > 1.70
    assert p1_L2[-1] > 1.70
    assert p0_H1[-1] > 0.95
    assert p1_H1[-1] > 0.95


# ------------------------------------------------------------------------- total-pressure MMS
def _single_run_total_pressure(n, M, theta):
    """test_convergence_totalpressuresolver.py:103-182."""
    from oracle.mpet import MPETTotalPressureOracle
    T = 1.0
    dt = float(T / M)
    params = dict(J=2, c=(1.0, 1.0), alpha=(1.0, 1.0), K=(1.0, 1.0), S=((1.0, 1.0), (1.0, 1.0)), E=1.0, nu=0.35)
    ex = _exact_solutions(params)
    mesh = unit_square_mesh(n)
    o = MPETTotalPressureOracle(mesh, params, dt=dt, theta=theta, T=T)
    o.f = Coef(fn=ex["f"], degree=3)
    o.g = [Coef(fn=ex["g"][i], degree=3) for i in range(2)]
    o.u_bar = Coef(fn=ex["u"], degree=3)
    o.s = Coef(fn=ex["sigma"], degree=4)       # sigma_ex * normal, sigma = 2 mu eps(u) + p0 I
    o.s_times_normal = True
    o.p_bar = [Coef(fn=ex["p"][i], degree=3) for i in range(2)]
    F = o.facets
    o.momentum_markers[:] = 0
    xm = mesh.coords[F["vertices"]]
    right = np.all(np.abs(xm[:, :, 0] - 1.0) < 3e-16, axis=1)     # near(x[0], 1.0)
    o.momentum_markers[right] = 1
    for i in range(2):
        o.continuity_markers[i][:] = 0
    # initial conditions, total pressure included (test_convergence_totalpressuresolver.py:156-163)
    sp_ = o.space
    x2 = sp_.node2_coords()
    u0 = ex["u"](x2, 0.0)
    for k in range(2):
        o.up_[sp_.u_dofs(k)] = u0[:, k]
    o.up_[sp_.p_dofs(0)] = ex["p_total"](mesh.coords, 0.0)
    for i in range(2):
        o.up_[sp_.p_dofs(i + 1)] = ex["p"][i](mesh.coords, 0.0)
    for up, t in o.solve_direct():
        pass
    # errors of (u, p1, p2): drop the total pressure from the split
    err = o.error_norms(up, ex["u"], ex["p"], t, qdeg=7, grad_u=ex["grad_u"], grad_p=ex["grad_p"],
                        fields=[1, 2])
    h = 2 * min(_circumradius(mesh))
    return err["u_L2"], err["u_H1"], err["p_L2"], err["p_H1"], h


@pytest.mark.parametrize("theta,ns,ms", [(0.5, [8, 16, 32], [4, 8, 16]),
                                         (1.0, [8, 16, 32], [8, 32, 128])])
def test_convergence_totalpressuresolver(theta, ns, ms):
    """test_convergence_totalpressuresolver.py:184-253 (thresholds :237-242)."""
    res = [_single_run_total_pressure(n, m, theta) for n, m in zip(ns, ms)]
    hs = [r[4] for r in res]
    u_L2 = _rates([r[0] for r in res], hs)
    u_H1 = _rates([r[1] for r in res], hs)
    p0_L2 = _rates([r[2][0] for r in res], hs)
    p1_L2 = _rates([r[2][1] for r in res], hs)
    p0_H1 = _rates([r[3][0] for r in res], hs)
    p1_H1 = _rates([r[3][1] for r in res], hs)
    print(u_L2, u_H1, p0_L2, p1_L2, p0_H1, p1_H1)
    assert u_L2[-1] > 1.70
    assert u_H1[-1] > 1.70
    assert p0_L2[-1] > 1.85
    assert p1_L2[-1] > 1.87
    assert p0_H1[-1] > 0.95
    assert p1_H1[-1] > 0.95

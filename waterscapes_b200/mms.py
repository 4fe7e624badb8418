"""Manufactured-solution data for the standard MPET formulation (input data only: no engine, no oracle).

Given an exact displacement u(x, t) and network pressures p_i(x, t) as sympy expressions, derive with sympy the
body force f and the sources g_i that make them solve the reference's equations (mpetsolver.py:196-201 with the
right-hand sides L0 :233 and L1 :249; the continuity rows carry the reference's sign convention, "the MPET module
assumes -g", demo/demo_adaptive.py:48-49):

    f   = -div sigma(u) + sum_i alpha_i grad p_i ,        sigma(u) = 2 mu eps(u) + lambda div(u) I
    g_i = -c_i dp_i/dt - alpha_i d(div u)/dt + div(K_i grad p_i) - sum_j S_ij (p_i - p_j)

This is the derivation the reference's own MMS tests do (test/test_convergence_mpetsolver.py:21-101), written for
the two-field (u, p) unknowns.  Returned as numpy callables fn(X[npts, 3], t).
"""
import math

import numpy as np


def cfg1_exact():
    """The MMS fields of BASELINE.json configs[0]: the 3-D extension of demo/demo_adaptive.py:62-63."""
    import sympy
    x = sympy.symbols("x0 x1 x2")
    t = sympy.symbols("t")
    pi = sympy.pi
    s, c = sympy.sin, sympy.cos
    u = [0.1 * c(pi * x[0]) * s(pi * x[1]) * s(pi * x[2]) * s(pi * t),
         0.1 * s(pi * x[0]) * c(pi * x[1]) * s(pi * x[2]) * s(pi * t),
         0.1 * s(pi * x[0]) * s(pi * x[1]) * c(pi * x[2]) * s(pi * t)]
    p = [(i + 1) * s(pi * x[0]) * c(pi * x[1]) * s(pi * x[2]) * s(2 * pi * t) for i in range(2)]
    return x, t, u, p


def standard_sources(params, x, t, u, p):
    import sympy
    J, nu, E = int(params["J"]), float(params["nu"]), float(params["E"])
    alpha, cc, K, S = params["alpha"], params["c"], params["K"], params["S"]
    lmbda = nu * E / ((1.0 - 2.0 * nu) * (1.0 + nu))
    mu = E / (2.0 * (1.0 + nu))
    d = len(x)
    diff = sympy.diff
    div_u = sum(diff(u[i], x[i]) for i in range(d))
    eps = [[0.5 * (diff(u[i], x[j]) + diff(u[j], x[i])) for j in range(d)] for i in range(d)]
    sigma = [[2 * mu * eps[i][j] + (lmbda * div_u if i == j else 0) for j in range(d)] for i in range(d)]
    f = [-sum(diff(sigma[i][j], x[j]) for j in range(d)) + sum(float(alpha[k]) * diff(p[k], x[i]) for k in range(J))
         for i in range(d)]
    g = [-float(cc[i]) * diff(p[i], t) - float(alpha[i]) * diff(div_u, t)
         + sum(diff(float(K[i]) * diff(p[i], x[j]), x[j]) for j in range(d))
         - sum(float(S[i][j]) * (p[i] - p[j]) for j in range(J)) for i in range(J)]

    def fn(expr):
        f_ = sympy.lambdify(tuple(x) + (t,), expr, "numpy")
        return lambda X, tt: np.broadcast_to(np.asarray(f_(*[X[:, k] for k in range(d)], tt), dtype=float),
                                             (X.shape[0],)).copy()

    def vec(exprs):
        fs = [fn(e) for e in exprs]
        return lambda X, tt: np.stack([f_(X, tt) for f_ in fs], axis=1)

    return dict(u=vec(u), p=[fn(pi_) for pi_ in p], f=vec(f), g=[fn(gi) for gi in g])

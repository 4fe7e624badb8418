"""3-D manufactured solutions and error norms for the convergence tests (no oracle involved).

The exact solutions are the 3-D extension of the reference's
``src/mpet/test/test_convergence_mpetsolver.py:21-101`` (``exact_solutions``): every displacement component is
``sin(2 pi x + pi/2) sin(2 pi y + pi/2) sin(2 pi z + pi/2) sin(omega t + t)``, network pressure i is ``-i`` times the
same function, the total pressure is ``lambda div u - sum alpha_i p_i``; f, g and the total stress follow with sympy
exactly as in the reference.  ``error_norms`` integrates the true error with a Duffy (collapsed Gauss) rule, the
role DOLFIN's ``errornorm(..., degree_rise=3)`` plays in the reference test (:165-178).
"""
import math

import numpy as np

LEDGE = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))       # UFC local edges of a tet


def exact_solutions(params, d=3):
    import sympy
    J, nu, E = params["J"], params["nu"], params["E"]
    alpha, c, K, S = params["alpha"], params["c"], params["K"], params["S"]
    lmbda = nu * E / ((1.0 - 2.0 * nu) * (1.0 + nu))
    mu = E / (2.0 * (1.0 + nu))
    pi = math.pi
    omega = 2 * pi
    sin, diff = sympy.sin, sympy.diff
    x = sympy.symbols("x0 x1 x2")[:d]
    t = sympy.symbols("t")
    space = 1
    for k in range(d):
        space = space * sin(2.0 * pi * x[k] + pi / 2.0)
    u = [space * sin(omega * t + t)] * d
    p = [0]
    for i in range(1, J + 1):
        p += [-(i) * space * sin(omega * t + t)]
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
        f_ = sympy.lambdify(tuple(x) + (t,), expr, "numpy")
        return lambda X, tt: np.broadcast_to(f_(*[X[:, k] for k in range(d)], tt), (X.shape[0],)).astype(float)

    def vec(exprs):
        fs = [fn(e) for e in exprs]
        return lambda X, tt: np.stack([f_(X, tt) for f_ in fs], axis=1)

    def ten(exprs):
        fs = [[fn(e) for e in row] for row in exprs]
        return lambda X, tt: np.stack([np.stack([f_(X, tt) for f_ in row], axis=1) for row in fs], axis=1)

    return dict(u=vec(u), grad_u=ten(grad_u), p=[fn(pi_) for pi_ in p[1:]], p_total=fn(p[0]),
                grad_p=[vec(gp) for gp in grad_p[1:]], f=vec(f), g=[fn(gi) for gi in g], sigma=ten(sigma))


def duffy_rule(npts):
    """Collapsed Gauss rule on the reference tet (exact for polynomials of degree 2*npts - 3 at least)."""
    xg, wg = np.polynomial.legendre.leggauss(npts)
    xg, wg = 0.5 * (xg + 1.0), 0.5 * wg
    a, b, c = np.meshgrid(xg, xg, xg, indexing="ij")
    wa, wb, wc = np.meshgrid(wg, wg, wg, indexing="ij")
    X = np.stack([a, b * (1 - a), c * (1 - a) * (1 - b)], axis=-1).reshape(-1, 3)
    W = (wa * wb * wc * (1 - a) ** 2 * (1 - b)).ravel()
    return X, W


def _p2_tables(X):
    lam = np.stack([1 - X.sum(1), X[:, 0], X[:, 1], X[:, 2]], axis=1)            # [q, 4]
    dlam = np.array([[-1.0, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]])            # [4, 3]
    phi = [lam[:, i] * (2 * lam[:, i] - 1) for i in range(4)] + [4 * lam[:, a] * lam[:, b] for a, b in LEDGE]
    dphi = [(4 * lam[:, i] - 1)[:, None] * dlam[i][None] for i in range(4)] + \
           [4 * (lam[:, a, None] * dlam[b][None] + lam[:, b, None] * dlam[a][None]) for a, b in LEDGE]
    return lam, dlam, np.stack(phi, axis=1), np.stack(dphi, axis=1)               # phi [q,10], dphi [q,10,3]


def error_norms(coords, cells, edge_index, nv, u_vals, p_vals, ex, t, npts=5, chunk=4096):
    """L2 and H1 errors of the P2 displacement ``u_vals`` [N2, 3] and the P1 pressures ``p_vals`` (list of [Nv])
    against the exact solution dict ``ex`` at time t.  ``edge_index(lo, hi)`` -> edge number."""
    X, W = duffy_rule(npts)
    lam, dlam, phi, dphi = _p2_tables(X)
    cells = np.asarray(cells, dtype=np.int64)
    out = dict(u_L2=0.0, u_H1=0.0, p_L2=[0.0] * len(p_vals), p_H1=[0.0] * len(p_vals))
    for c0 in range(0, cells.shape[0], chunk):
        cv = cells[c0:c0 + chunk]
        xv = coords[cv]                                                     # [c, 4, 3]
        Jm = np.swapaxes(xv[:, 1:, :] - xv[:, :1, :], 1, 2)                  # J[i][j] = d x_i / d X_j
        det = np.abs(np.linalg.det(Jm))
        Ji = np.linalg.inv(Jm)                                              # dX_k / dx_m
        xq = xv[:, 0][:, None, :] + np.einsum("cij,qj->cqi", Jm, X)
        w = det[:, None] * W[None, :]
        nodes = np.concatenate([cv] + [(nv + edge_index(cv[:, a], cv[:, b]))[:, None] for a, b in LEDGE], axis=1)
        ul = u_vals[nodes]                                                  # [c, 10, 3]
        uh = np.einsum("qa,cak->cqk", phi, ul)
        gph = np.einsum("qaj,cjm->cqam", dphi, Ji)                           # physical gradients of the basis
        guh = np.einsum("cqam,cak->cqkm", gph, ul)
        flat = xq.reshape(-1, 3)
        ue = ex["u"](flat, t).reshape(uh.shape)
        gue = ex["grad_u"](flat, t).reshape(guh.shape)
        l2 = np.einsum("cq,cqk,cqk->", w, uh - ue, uh - ue)
        out["u_L2"] += l2
        out["u_H1"] += l2 + np.einsum("cq,cqkm,cqkm->", w, guh - gue, guh - gue)
        for i, pv in enumerate(p_vals):
            pl = pv[cv]
            ph = np.einsum("qa,ca->cq", lam, pl)
            gp = np.einsum("aj,cjm,ca->cm", dlam, Ji, pl)
            pe = ex["p"][i](flat, t).reshape(ph.shape)
            gpe = ex["grad_p"][i](flat, t).reshape(ph.shape + (3,))
            l2p = np.einsum("cq,cq,cq->", w, ph - pe, ph - pe)
            out["p_L2"][i] += l2p
            dg = gp[:, None, :] - gpe
            out["p_H1"][i] += l2p + np.einsum("cq,cqm,cqm->", w, dg, dg)
    out["u_L2"], out["u_H1"] = math.sqrt(out["u_L2"]), math.sqrt(out["u_H1"])
    out["p_L2"] = [math.sqrt(v) for v in out["p_L2"]]
    out["p_H1"] = [math.sqrt(v) for v in out["p_H1"]]
    return out


def hmin(coords, cells):
    """DOLFIN's mesh.hmin(): smallest cell diameter = 2 * circumradius of a tet."""
    x = coords[np.asarray(cells, dtype=np.int64)]
    a, b, c = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]
    vol6 = np.abs(np.einsum("ci,ci->c", a, np.cross(b, c)))
    num = (np.einsum("ci,ci->c", a, a)[:, None] * np.cross(b, c) + np.einsum("ci,ci->c", b, b)[:, None] * np.cross(c, a)
           + np.einsum("ci,ci->c", c, c)[:, None] * np.cross(a, b))
    return float(np.min(2.0 * np.linalg.norm(num, axis=1) / (2.0 * vol6)))


def rates(errors, hs):
    return [math.log(errors[i + 1] / errors[i]) / math.log(hs[i + 1] / hs[i]) for i in range(len(hs) - 1)]

"""CPU: the oracle's MINRES + block-AMG restatement agrees with its own direct solve."""
import numpy as np
import scipy.sparse.linalg as spla

from oracle.mesh import unit_cube_mesh
from oracle.mpet import MPETOracle, Coef
from oracle.krylov import minres, BlockAMG


def _problem(n=4, theta=1.0):
    mesh = unit_cube_mesh(n, jitter=0.2)
    params = dict(J=2, E=2.2, nu=0.4545, alpha=(0.5, 0.5), c=(1.0, 1.0), K=(1.0, 1.0), S=((0, 1.0), (1.0, 0)))
    o = MPETOracle(mesh, params, dt=0.1, theta=theta, T=0.2)
    o.momentum_markers[:] = 0
    for i in range(2):
        o.continuity_markers[i][:] = 0
    o.p_bar = [Coef(fn=lambda x, t, i=i: (i + 1) * np.sin(np.pi * x[:, 0]) * np.sin(2 * np.pi * t)) for i in range(2)]
    o.f = Coef(value=(0.3, -0.2, 1.0))
    o.g = [Coef(value=0.5), Coef(value=1.5)]
    return o


def test_minres_amg_matches_direct():
    o = _problem()
    A = o.assemble_lhs()
    b, dofs, vals = o.rhs(0.0)
    As, bs = o.apply_bc_symmetric(A, dofs, b)
    xref = spla.splu(As.tocsc()).solve(bs)
    M = BlockAMG(o, dofs)
    mask = np.zeros(o.space.N, bool)
    mask[dofs] = True
    x0 = np.zeros(o.space.N)
    x0[dofs] = vals
    x, info = minres(A, b, x0, M, mask=mask, rtol=1e-12, maxit=2000)
    assert info["converged"] and info["niter"] < 200
    assert np.linalg.norm(x - xref) / np.linalg.norm(xref) < 1e-9
    # SPD + symmetric preconditioner
    rng = np.random.default_rng(0)
    a = rng.standard_normal(o.space.N); a[mask] = 0
    c = rng.standard_normal(o.space.N); c[mask] = 0
    za, zc = M(a), M(c)
    assert abs(za @ c - a @ zc) < 1e-10 * abs(za @ c)
    assert za @ a > 0


def test_iterative_time_loop_tracks_direct():
    o1, o2 = _problem(3), _problem(3)
    ref = [up.copy() for up, t in o1.solve_direct()]
    mon = []
    out = [up.copy() for up, t in o2.solve_iterative(rtol=1e-11, monitor=mon)]
    assert len(out) == len(ref) == 2
    assert np.linalg.norm(out[-1] - ref[-1]) / np.linalg.norm(ref[-1]) < 1e-8
    assert all(m["converged"] for m in mon)

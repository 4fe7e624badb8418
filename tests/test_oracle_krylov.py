"""CPU: the oracle's MINRES + block-AMG restatement agrees with its own direct solve."""
import numpy as np
import pytest
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


def test_lanczos_bounds_and_polynomial_mode():
    """The spectral bounds that select the polynomial mode (oracle/krylov.py, mirrored by amg.cu) against
    scipy's eigensolver, and the accuracy of the resulting Chebyshev inverse."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle.krylov import lanczos_bounds, poly_degree, ScalarAMG, POLY_TARGET
    o = _problem(5)
    o.K = [1e-4, 1e-4]                       # mass-dominated network blocks (cond(D^-1 A) ~ 5)
    M = o.assemble_prec().tocsr()
    d = o.space.p_dofs(0)
    A = M[d][:, d].tocsr()
    dinv = 1.0 / A.diagonal()
    lmin, lmax = lanczos_bounds(A, dinv)
    Dm = sp.diags(np.sqrt(dinv))
    ev = np.linalg.eigvalsh((Dm @ A @ Dm).toarray())
    # Ritz values converge from inside the spectrum; the bounds are padded by 3 % / 2 % where they are used
    assert -1e-10 <= lmin - ev[0] < 1e-3 * ev[0] and -1e-10 <= ev[-1] - lmax < 1e-3 * ev[-1]
    h = ScalarAMG(A, allow_polynomial=True, cycles=2, degree=4)
    assert h.poly is not None and h.poly[2] == poly_degree(h.poly[0], h.poly[1])
    rng = np.random.default_rng(0)
    b = rng.standard_normal(A.shape[0])
    x = h.apply(b)[:, 0]
    xe = spla.spsolve(A.tocsc(), b)
    # error in the A-norm relative to the solution: below the residual-polynomial target (with slack)
    e = x - xe
    assert np.sqrt(e @ (A @ e)) < 5 * POLY_TARGET * np.sqrt(xe @ (A @ xe))
    # the polynomial preconditioner is symmetric positive definite
    c = rng.standard_normal(A.shape[0])
    assert abs(h.apply(b)[:, 0] @ c - b @ h.apply(c)[:, 0]) < 1e-10 * abs(b @ h.apply(c)[:, 0]) + 1e-12
    assert b @ x > 0


def test_drop_small_preserves_row_sums_and_symmetry():
    import scipy.sparse as sp
    from oracle.krylov import drop_small
    rng = np.random.default_rng(1)
    n = 200
    B = sp.random(n, n, density=0.05, random_state=3, format="csr")
    B.data[np.abs(B.data) < 0.5] *= 0.01          # many tiny entries
    S = (B + B.T).tocsr()
    A = (S + sp.diags(np.asarray(abs(S).sum(axis=1)).ravel() + 1.0)).tocsr()   # strictly diagonally dominant: SPD
    D = drop_small(A, 0.01)
    assert D.nnz < A.nnz
    assert np.allclose(np.asarray(D.sum(axis=1)).ravel(), np.asarray(A.sum(axis=1)).ravel())
    assert abs(D - D.T).max() < 1e-14
    assert np.all(np.linalg.eigvalsh(D.toarray()) > 0)


@pytest.mark.parametrize("A", [1, 2, 3])
def test_openmp_cell_loop_matches_numpy_assembly(A):
    """oracle/csrc/cpu_kernels.c:oracle_assemble_lhs_tet (the CPU timing arm's assemble(a)) == the numpy
    restatement, to round-off, on a jittered mesh."""
    from oracle import omp
    from tests.test_gpu_assembly import PARAMS
    mesh = unit_cube_mesh(3, jitter=0.2)
    o = MPETOracle(mesh, PARAMS[A], dt=0.05, theta=0.5)
    ref = o.on_pattern(o.assemble_lhs())
    got = omp.LhsAssembler(o)()
    assert np.array_equal(got.indices, ref.indices) and np.array_equal(got.indptr, ref.indptr)
    assert np.linalg.norm(got.data - ref.data) / np.linalg.norm(ref.data) < 1e-14

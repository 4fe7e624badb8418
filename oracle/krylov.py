"""Oracle restatement of the Krylov path (test infrastructure; see oracle/__init__.py).

What the reference asks PETSc/hypre to do -- ``PETScKrylovSolver("minres", "hypre_amg")`` with
``set_operators(Acopy, P)`` (mpetsolver.py:507,553-556; mpettotalpressuresolver.py:448,492-493) --
restated with scipy.sparse matvecs:

* ``minres``: PETSc's KSPMINRES recurrence [EXT: petsc/src/ksp/ksp/impls/minres/minres.c], one operator
  and one preconditioner application per iteration, convergence test on the preconditioned norm
  ``|eta| <= max(rtol * ||b||_B, atol)`` with b the right-hand side after ``apply_symmetric``
  (bc_symmetric.py:11-22) -- PETSc's KSPConvergedDefault takes the norm of b, not of the initial residual,
  as reference when the initial guess is nonzero [EXT: petsc/src/ksp/ksp/interface/iterativ.c] --,
  Dirichlet rows handled as in ``apply_symmetric``; stagnation exit as in krylov.cu.
* ``BlockAMG``: the block-diagonal V-cycle of DESIGN.md section "Preconditioner" (hypre BoomerAMG itself is
  an un-vendored dependency; its defaults are not reproducible here).  This class restates, loop for
  loop, the hierarchy the CUDA library builds (p-coarsening P2 -> P1, greedy smoothed aggregation,
  degree-2 Chebyshev smoothing, dense coarsest solve) so that iteration counts can be compared.
"""
import numpy as np
import scipy.sparse as sp

CHEB_DEGREE = 2
CHEB_RATIO = 4.0
# well-conditioned blocks (cond(D^-1 A) <= POLY_KAPPA_MAX, e.g. the mass-dominated network blocks) are inverted
# by a Chebyshev polynomial on the WHOLE spectrum instead of a V-cycle (amg.cu: same constants)
POLY_KAPPA_MAX = 32.0
POLY_TARGET = 1.0e-4
POLY_MAX_DEGREE = 28
LANCZOS_STEPS = 40
# Galerkin operators of the aggregated levels: entries below DROP_TOL * sqrt(a_ii a_jj) are lumped into the
# diagonal (smoothed aggregation fills in quickly: 850 entries per row two levels below the mesh on cfg5)
DROP_TOL = 0.01
# P1-field blocks that keep a hierarchy: P1_CYCLES stationary cycles, Chebyshev degree P1_DEGREE (amg.cu)
P1_CYCLES = 2
P1_DEGREE = 4
COARSE_MAX = 300
DENSE_MAX = 700
MAX_LEVELS = 12
STAG_WINDOW = 256      # krylov.cu: kStagWindow / kStagFactor
STAG_FACTOR = 0.99


# ------------------------------------------------------------------------------------------ MINRES
def minres(A, b, x0, M, mask=None, rtol=1e-5, atol=1e-50, maxit=10000):
    """Preconditioned MINRES.  ``A``: scipy matrix WITHOUT boundary conditions; ``mask``: bool array of
    Dirichlet rows (x0 must carry the boundary values there); ``M(r)``: SPD preconditioner.
    Returns (x, info dict)."""
    x = x0.copy()
    free = np.ones(b.shape[0], dtype=bool) if mask is None else ~mask

    def op(v):
        y = A @ v
        if mask is not None:
            y[mask] = v[mask]
        return y

    r = np.where(free, b - A @ x, 0.0)
    z = M(r)
    dp = r @ z
    if dp < 0:
        raise RuntimeError("indefinite preconditioner")
    beta = np.sqrt(dp)
    norm0 = beta
    # reference norm: sqrt(b . B b) of the symmetrically eliminated right-hand side
    if mask is not None:
        xb = np.where(mask, x, 0.0)
        rb = np.where(free, b - A @ xb, xb)
    else:
        rb = b
    dpb = float(rb @ M(rb))
    bnorm = np.sqrt(dpb) if dpb > 0 else beta
    tol = max(rtol * bnorm, atol)
    stag_ref = 1e300
    info = dict(niter=0, converged=False, res0=norm0, rel_res=1.0, bnorm=bnorm, reason=-3)
    if beta <= tol or beta == 0.0:
        info.update(converged=True, rel_res=0.0 if beta == 0 else beta / bnorm, reason=2)
        return x, info
    eta = beta
    c = c_old = 1.0
    s = s_old = 0.0
    v, u = r / beta, z / beta
    v_old = np.zeros_like(v)
    u_old = np.zeros_like(v)
    w1 = np.zeros_like(v)
    w2 = np.zeros_like(v)
    for it in range(1, maxit + 1):
        r = op(u)
        alpha = r @ u
        z = M(r)
        r = r - alpha * v - beta * v_old
        z = z - alpha * u - beta * u_old
        beta_old = beta
        dp = max(r @ z, 0.0)
        beta = np.sqrt(dp)
        c_oold, s_oold = c_old, s_old
        c_old, s_old = c, s
        rho0 = c_old * alpha - c_oold * s_old * beta_old
        rho1 = np.sqrt(rho0 * rho0 + beta * beta)
        rho2 = s_old * alpha + c_oold * c_old * beta_old
        rho3 = s_oold * beta_old
        c, s = rho0 / rho1, beta / rho1
        wn = (u - rho2 * w1 - rho3 * w2) / rho1
        w2, w1 = w1, wn
        x = x + c * eta * wn
        eta = -s * eta
        v_old, u_old = v, u
        if beta != 0.0:
            v, u = r / beta, z / beta
        info["niter"] = it
        info["rel_res"] = abs(eta) / bnorm
        if abs(eta) <= tol:
            info["converged"] = True
            info["reason"] = 2
            break
        if it % STAG_WINDOW == 0:
            if abs(eta) > STAG_FACTOR * stag_ref:
                info["reason"] = -5
                break
            stag_ref = abs(eta)
        if beta == 0.0:
            break
    return x, info


# ------------------------------------------------------------------------------------------ AMG
def _sa_prolongator(A, theta):
    """Greedy aggregation + one damped-Jacobi smoothing step (same passes, same order as amg.cu)."""
    A = A.tocsr()
    n = A.shape[0]
    rp, ci, av = A.indptr, A.indices, A.data
    diag = A.diagonal().copy()
    diag[diag == 0] = 1.0
    rows = np.repeat(np.arange(n), np.diff(rp))
    off = (ci != rows) & (av != 0.0)
    scale = np.sqrt(np.abs(diag[rows] * diag[ci]))
    strong = off & (np.abs(av) >= theta * scale)
    isolated = np.bincount(rows[off], minlength=n) == 0
    S = sp.csr_matrix((np.where(strong, np.abs(av) / scale, 0.0)[strong], ci[strong],
                       np.concatenate([[0], np.cumsum(np.bincount(rows[strong], minlength=n))])), shape=(n, n))
    srp, sci, sval = S.indptr, S.indices, S.data
    agg = np.full(n, -1, dtype=np.int64)
    nagg = 0
    for r in range(n):                                    # pass 1
        if agg[r] != -1 or isolated[r] or srp[r + 1] == srp[r]:
            continue
        nb = sci[srp[r]:srp[r + 1]]
        if np.any(agg[nb] != -1):
            continue
        agg[r] = nagg
        agg[nb] = nagg
        nagg += 1
    agg2 = agg.copy()
    for r in range(n):                                    # pass 2
        if agg[r] != -1 or isolated[r]:
            continue
        nb = sci[srp[r]:srp[r + 1]]
        ok = agg[nb] != -1
        if ok.any():
            w = np.where(ok, sval[srp[r]:srp[r + 1]], -1.0)
            agg2[r] = agg[nb[np.argmax(w)]]
    agg = agg2
    for r in range(n):                                    # pass 3
        if agg[r] != -1 or isolated[r]:
            continue
        agg[r] = nagg
        nb = sci[srp[r]:srp[r + 1]]
        take = nb[(agg[nb] == -1) & ~isolated[nb]]
        agg[take] = nagg
        nagg += 1
    if nagg == 0:
        return None
    have = agg >= 0
    T = sp.csr_matrix((np.ones(have.sum()), (np.nonzero(have)[0], agg[have])), shape=(n, nagg))
    x = 1.0 + 0.37 * np.sin(1.7 * np.arange(n))
    rho = 1.0
    for _ in range(15):
        x = x / np.linalg.norm(x)
        y = (A @ x) / diag
        rho = np.linalg.norm(y)
        x = y
    rho *= 1.1
    omega = (4.0 / 3.0) / rho
    Dinv = sp.diags(np.where(isolated, 0.0, 1.0 / diag))
    keep = sp.diags(np.where(isolated, 0.0, 1.0))
    P = keep @ T - omega * (Dinv @ (A @ T))
    return P.tocsr()


class _Level:
    pass


def drop_small(A, tol=None):
    """Lump off-diagonal entries with |a_ij| < tol * sqrt(|a_ii a_jj|) into the diagonal (symmetric criterion,
    row sums preserved).  amg.cu:drop_small is the same pass."""
    tol = DROP_TOL if tol is None else tol
    A = A.tocsr()
    if tol <= 0:
        return A
    d = A.diagonal()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    small = (rows != A.indices) & (np.abs(A.data) < tol * np.sqrt(np.abs(d[rows] * d[A.indices])))
    lump = np.bincount(rows[small], weights=A.data[small], minlength=A.shape[0])
    keep = ~small
    out = sp.csr_matrix((A.data[keep], (rows[keep], A.indices[keep])), shape=A.shape)
    out = (out + sp.diags(lump)).tocsr()
    out.sort_indices()
    return out


def lanczos_bounds(A, dinv, steps=LANCZOS_STEPS):
    """Extreme Ritz values of D^-1 A from `steps` Lanczos steps in the D inner product, deterministic start
    vector (amg.cu:lanczos_bounds runs the same recurrence on the device)."""
    n = A.shape[0]
    d = 1.0 / dinv
    i = np.arange(n)
    v = 1.0 + 0.5 * np.sin(0.37 * i + 0.1 * (i % 7))
    v = v / np.sqrt(np.sum(d * v * v))
    v_prev = np.zeros(n)
    beta = 0.0
    al, be = [], []
    for _ in range(min(steps, n)):
        w = dinv * (A @ v)
        a = float(np.sum(d * w * v))
        w = w - a * v - beta * v_prev
        al.append(a)
        b2 = float(np.sum(d * w * w))
        if b2 <= 1e-28 * max(1.0, a * a):
            break
        beta = np.sqrt(b2)
        be.append(beta)
        v_prev, v = v, w / beta
    m = len(al)
    T = np.diag(al) + np.diag(be[:m - 1], 1) + np.diag(be[:m - 1], -1)
    ev = np.linalg.eigvalsh(T)
    return float(ev[0]), float(ev[-1])


def poly_degree(lo, hi):
    """Smallest Chebyshev degree whose residual polynomial is below POLY_TARGET on [lo, hi]."""
    kappa = hi / lo
    sigma = (np.sqrt(kappa) - 1.0) / (np.sqrt(kappa) + 1.0)
    k = 1
    while k < POLY_MAX_DEGREE and 2.0 * sigma ** k / (1.0 + sigma ** (2 * k)) > POLY_TARGET:
        k += 1
    return k


class ScalarAMG:
    """V-cycle hierarchy for one scalar SPD block (identity rows on eliminated Dirichlet dofs)."""

    def __init__(self, A0, first_coarse=None, allow_polynomial=False, cycles=1, degree=None):
        self.levels = []
        self.poly = None
        self.cycles = cycles
        self.degree = CHEB_DEGREE if degree is None else degree
        self._push(A0.tocsr(), None)
        self.coarse_inv = None
        if allow_polynomial:
            lmin, lmax = lanczos_bounds(self.levels[0].A, self.levels[0].dinv)
            lo, hi = 0.97 * lmin, 1.02 * lmax
            if lmin > 0 and hi / lo <= POLY_KAPPA_MAX:
                self.poly = (lo, hi, poly_degree(lo, hi))
                return
        if first_coarse is not None:
            P, A1 = first_coarse
            self._push(A1.tocsr(), P.tocsr())
        A = self.levels[-1].A
        theta = 0.08
        while A.shape[0] > COARSE_MAX and len(self.levels) < MAX_LEVELS:
            P = _sa_prolongator(A, theta)
            if P is None or P.shape[1] > 0.8 * A.shape[0]:
                break
            Ac = drop_small((P.T @ (A @ P)).tocsr())
            Ac.sort_indices()
            self._push(Ac, P)
            A = Ac
            theta *= 0.5
        self.coarse_inv = None
        if A.shape[0] <= DENSE_MAX:
            self.coarse_inv = np.linalg.inv(A.toarray())

    def _push(self, A, P):
        L = _Level()
        L.A, L.P = A, P
        d = A.diagonal().copy()
        d[d == 0] = 1.0
        L.dinv = 1.0 / d
        n = A.shape[0]
        gersh = float(np.max(np.asarray(abs(A).sum(axis=1)).ravel() * np.abs(L.dinv)))
        i = np.arange(n)
        x = 1.0 + 0.5 * np.sin(0.37 * i + 0.1 * (i % 7))
        lam = 0.0
        for _ in range(20):
            y = (A @ x) * L.dinv
            lam = np.linalg.norm(y)
            if lam == 0:
                break
            x = y / lam
        est = 1.1 * lam
        L.lmax = gersh if (est <= 0 or est > gersh) else est
        self.levels.append(L)

    def _cheb(self, L, b, x, bounds=None, degree=None):
        lmax, lmin = (L.lmax, L.lmax / CHEB_RATIO) if bounds is None else (bounds[1], bounds[0])
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho_old = 1.0 / sigma
        d = None
        for k in range(self.degree if degree is None else degree):
            if k == 0:
                r = b if x is None else b - L.A @ x
                d = (1.0 / theta) * (L.dinv[:, None] * r)
                x = d.copy() if x is None else x + d
            else:
                rho = 1.0 / (2.0 * sigma - rho_old)
                r = b - L.A @ x
                d = rho * rho_old * d + (2.0 * rho / delta) * (L.dinv[:, None] * r)
                x = x + d
                rho_old = rho
        return x

    def _vcycle(self, lev, b):
        L = self.levels[lev]
        if lev == len(self.levels) - 1:
            if self.coarse_inv is not None:
                return self.coarse_inv @ b
            return self._cheb(L, b, None)
        x = self._cheb(L, b, None)
        r = b - L.A @ x
        C = self.levels[lev + 1]
        xc = self._vcycle(lev + 1, C.P.T @ r)
        x = x + C.P @ xc
        return self._cheb(L, b, x)

    def apply(self, B):
        """B: [n, nrhs] -> approximate A^-1 B."""
        B = B.reshape(B.shape[0], -1)
        if self.poly is not None:
            lo, hi, deg = self.poly
            return self._cheb(self.levels[0], B, None, bounds=(lo, hi), degree=deg)
        x = self._vcycle(0, B)
        for _ in range(1, self.cycles):          # stationary iteration x <- x + B (b - A x)
            x = x + self._vcycle(0, B - self.levels[0].A @ x)
        return x


def _eliminate(M, mask):
    """Symmetric Dirichlet elimination with unit diagonal (apply_symmetric(bc, P), mpetsolver.py:503-504)."""
    M = M.tocsr(copy=True)
    rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
    kill = mask[rows] | mask[M.indices]
    M.data[kill] = 0.0
    M = M + sp.diags(mask.astype(float))
    M = M.tocsr()
    M.sort_indices()
    return M


class BlockAMG:
    """Block-diagonal preconditioner: shared scalar P2 block mu*(grad,grad) for the three displacement
    components (p-coarsened to P1, then aggregation) + one hierarchy per P1 field, which degenerates to a
    Chebyshev polynomial on the whole spectrum when the field's block is well conditioned (amg.cu)."""

    def __init__(self, oracle, dirichlet_dofs, light=False):
        """``light``: one cycle with degree-2 smoothing for the P1 fields that keep a hierarchy (what the CUDA
        library does for relative tolerances >= 1e-8, amg.cu: g_p_degree / krylov.cu: light_tolerance), and
        degree-1 smoothing in the displacement cycle (amg.cu: g_u_degree)."""
        o = oracle
        sp_ = o.space
        assert o.d == 3 and sp_.nreal == 0 and sp_.perm is None
        from .mpet import convert_to_mu_lmbda
        mu, _ = convert_to_mu_lmbda(o.E, o.nu)
        N2, Nv, A = sp_.N2, sp_.Nv, o.nfields
        mask = np.zeros(sp_.N, dtype=bool)
        mask[dirichlet_dofs] = True
        for k in (1, 2):
            assert np.array_equal(mask[:N2], mask[k * N2:(k + 1) * N2])
        P = o.assemble_prec().tocsr()
        m2 = mask[:N2]
        A2 = _eliminate(P[:N2, :N2], m2)
        # P1 stiffness * mu on the vertex block, eliminated with the vertex part of the mask
        ones = MPETOracleScalar(o)
        A1 = _eliminate(mu * ones.p1_stiffness(), m2[:Nv])
        ev = sp_.edge_vertices
        rows, cols, vals = [], [], []
        for r in range(N2):
            if m2[r]:
                continue
            if r < Nv:
                rows.append(r); cols.append(r); vals.append(1.0)
            else:
                for v in ev[r - Nv]:
                    if not m2[v]:
                        rows.append(r); cols.append(v); vals.append(0.5)
        P21 = sp.csr_matrix((vals, (rows, cols)), shape=(N2, Nv))
        self.u = ScalarAMG(A2, first_coarse=(P21, A1), degree=1 if light else None)
        self.p = []
        for i in range(A):
            d = sp_.p_dofs(i)
            self.p.append(ScalarAMG(_eliminate(P[d][:, d], mask[d]), allow_polynomial=True,
                                    cycles=1 if light else P1_CYCLES, degree=2 if light else P1_DEGREE))
        self.N2, self.Nv, self.A = N2, Nv, A

    def __call__(self, r):
        N2, Nv = self.N2, self.Nv
        z = np.empty_like(r)
        U = self.u.apply(r[:3 * N2].reshape(3, N2).T)
        z[:3 * N2] = U.T.reshape(-1)
        for i in range(self.A):
            lo = 3 * N2 + i * Nv
            z[lo:lo + Nv] = self.p[i].apply(r[lo:lo + Nv])[:, 0]
        return z

    def num_levels(self):
        return [len(self.u.levels)] + [len(h.levels) for h in self.p]


class MPETOracleScalar:
    """Scalar P1 operators on the oracle's mesh (helper for the p-coarsened level)."""

    def __init__(self, oracle):
        self.o = oracle

    def p1_stiffness(self):
        o = self.o
        Nv = o.space.Nv
        nc = o.mesh.num_cells
        rows, cols, vals = [], [], []
        for c0 in range(0, nc, 8192):
            cells = np.arange(c0, min(nc, c0 + 8192))
            _, _, _, Lp = o.element_blocks(cells)
            cv = o.mesh.cells[cells]
            rows.append(np.repeat(cv[:, :, None], 4, axis=2).ravel())
            cols.append(np.repeat(cv[:, None, :], 4, axis=1).ravel())
            vals.append(Lp.ravel())
        M = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(Nv, Nv))
        return M.tocsr()

"""Oracle restatement of the MPET forms, assembly and time loop (test infrastructure).

Follows, line by line, what ``src/mpet/mpet/mpetsolver.py`` asks DOLFIN to do:

* ``create_variational_forms`` :134-298 -- main form ``F`` :196-201, nullspace
  Lagrange terms :203-215, ``L0`` :233, ``L1``/Robin :244-254, ``a = lhs(F)``,
  ``L = rhs(F)`` :260-261, preconditioner :264-277 (well-formed analogue
  ``mpettotalpressuresolver.py:276-283``).
* ``create_dirichlet_bcs`` :63-84, ``bc.apply`` [EXT: row zero + unit diagonal,
  rhs := boundary value], ``bc_symmetric.py:6-22`` (MatZeroRowsColumns).
* ``solve_direct`` :382-462 (time loop order) and ``step`` :317-379.
* ``mpetproblem.py:13-24`` (Lame parameters / elastic stress),
  ``rm_basis_L2.py:10-74`` (L2-orthonormal rigid motions).

Element-level formulas are SURVEY.md Appendix A.  Rows = test functions,
columns = trial functions.  Everything is fp64.
"""
import sys
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .fem import MixedSpace, simplex_quadrature, tabulate, lagrange_nodes

DIRICHLET_MARKER = 0      # mpetsolver.py:17
NEUMANN_MARKER = 1        # mpetsolver.py:18
ROBIN_MARKER = 2          # mpetsolver.py:19
INVALID = sys.maxsize     # mpetproblem.py:153


def convert_to_mu_lmbda(E, nu):
    """mpetproblem.py:13-16."""
    mu = E / (2.0 * (1.0 + nu))
    lmbda = nu * E / ((1.0 - 2.0 * nu) * (1.0 + nu))
    return mu, lmbda


class Coef:
    """A form coefficient.

    ``value``: spatially constant (number / tuple, or callable of t).
    ``fn(x[npts,d], t)``: spatially varying; ``degree`` = the DOLFIN
    ``Expression(..., degree=k)`` interpolation degree (per-cell nodal
    interpolation into P_k before quadrature [EXT]); ``degree=None`` evaluates
    exactly at the quadrature points.
    """

    def __init__(self, value=None, fn=None, degree=None):
        self.value, self.fn, self.degree = value, fn, degree

    @property
    def is_constant(self):
        return self.fn is None

    def qdegree(self):
        if self.is_constant:
            return 0
        return 2 if self.degree is None else self.degree

    def const(self, t):
        v = self.value(t) if callable(self.value) else self.value
        return np.asarray(v, dtype=float)

    def is_zero(self, t):
        return self.is_constant and not np.any(self.const(t))

    def at_points(self, x, t):
        """Exact point evaluation at physical points x [npts, d]."""
        if self.is_constant:
            v = self.const(t)
            return np.broadcast_to(v, (x.shape[0],) + v.shape).copy()
        return np.asarray(self.fn(x, t), dtype=float)

    def at_cell_points(self, mesh, cells, ref_pts, t):
        """Values at reference points of given cells: [nc, nq, *shape]."""
        xc = mesh.coords[mesh.cells[cells]]                      # [nc, d+1, d]
        d = mesh.dim

        def phys(pts):
            lam = np.concatenate([1 - pts.sum(1, keepdims=True), pts], axis=1)
            return np.einsum("qv,cvd->cqd", lam, xc)

        if self.is_constant or self.degree is None:
            xq = phys(ref_pts)
            v = self.at_points(xq.reshape(-1, d), t)
            return v.reshape(xq.shape[:2] + v.shape[1:])
        nodes = lagrange_nodes(d, self.degree)
        xn = phys(nodes)
        vn = np.asarray(self.fn(xn.reshape(-1, d), t), dtype=float)
        vn = vn.reshape(xn.shape[:2] + vn.shape[1:])             # [nc, nnode, ...]
        N, _ = tabulate(d, self.degree, ref_pts)                 # [nq, nnode]
        return np.einsum("qn,cn...->cq...", N, vn)


ZERO = Coef(value=0.0)


def _refined_solve(lu, A, b, sweeps=2):
    """Sparse LU + fixed-precision iterative refinement (what MUMPS' ICNTL(10) does).  With nu = 0.4999 and
    pressures eight orders of magnitude above the displacement, a bare SuperLU answer carries a relative
    error of ~1e-5 in the displacement (two column orderings disagree by that much); two sweeps bring the
    residual to round-off and the answer is then reproducible to ~1e-10."""
    x = lu.solve(b)
    for _ in range(sweeps):
        x = x + lu.solve(b - A @ x)
    return x


class MPETOracle:
    """Oracle twin of ``MPETProblem`` + ``MPETSolver`` (standard formulation)."""

    P_SHIFT = 0      # P1 field index of network i is i + P_SHIFT (the total-pressure twin puts p0 first)

    def __init__(self, mesh, params, dt=0.1, theta=1.0, t=0.0, T=1.0,
                 u_has_nullspace=False, p_has_nullspace=None, perm=None):
        self.mesh = mesh
        d = mesh.dim
        self.d = d
        self.A = int(params["J"])
        A = self.A
        self.E, self.nu = float(params["E"]), float(params["nu"])
        self.alpha = [float(v) for v in params["alpha"]]
        # K_i: a number, or an array with one value per cell (DG0 permeability, as in the reference's
        # sandbox/biot-robin/three_fields_precond.py:263-271)
        self.K_cell = {i: np.asarray(v, dtype=float) for i, v in enumerate(params["K"]) if np.ndim(v) == 1}
        self.K = [float(np.mean(v)) for v in params["K"]]
        self.c = [float(v) for v in params["c"]]
        self.S = np.array(params["S"], dtype=float).reshape(A, A)
        self.dt, self.theta, self.t, self.T = float(dt), float(theta), float(t), float(T)
        self.u_has_nullspace = bool(u_has_nullspace)
        self.p_has_nullspace = list(p_has_nullspace) if p_has_nullspace is not None else [False] * A
        self.Z = self.rigid_motions() if self.u_has_nullspace else []
        nreal = len(self.Z) + sum(self.p_has_nullspace)
        self.nfields = A + self.P_SHIFT                       # P1 fields of the mixed space
        self.space = MixedSpace(mesh, self.nfields, nreal=nreal, perm=perm)
        # data (mpetproblem.py:141-149 defaults)
        zero_vec = Coef(value=(0.0,) * d)
        self.f, self.s, self.u_bar = zero_vec, zero_vec, zero_vec
        self.s_times_normal = False
        self.g = [ZERO] * A
        self.I = [ZERO] * A
        self.beta = [ZERO] * A
        self.p_robin = [ZERO] * A
        self.p_bar = [ZERO] * A
        self.facets = mesh.exterior_facets()
        nf = self.facets["cell"].shape[0]
        self.momentum_markers = np.full(nf, INVALID, dtype=np.int64)
        self.continuity_markers = [np.full(nf, INVALID, dtype=np.int64) for _ in range(A)]
        self.up_ = np.zeros(self.space.N)
        self.up = np.zeros(self.space.N)
        self._geom = None

    # ------------------------------------------------------------------ geometry
    def geometry(self):
        """Per-cell affine geometry: Jinv^T-applied gradients need Jinv; detJ."""
        if self._geom is None:
            x = self.mesh.coords[self.mesh.cells]
            J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)     # J[c, i, j] = d x_i / d X_j
            self._geom = (np.linalg.inv(J), np.abs(np.linalg.det(J)))
        return self._geom

    def facet_geometry(self, sel):
        """Outward unit normal [nf, d] and facet measure [nf] for facets ``sel``."""
        F = self.facets
        fv = self.mesh.coords[F["vertices"][sel]]                # [nf, d, d]
        cells = F["cell"][sel]
        opp = self.mesh.coords[self.mesh.cells[cells, F["local"][sel]]]
        if self.d == 2:
            tvec = fv[:, 1] - fv[:, 0]
            n = np.stack([tvec[:, 1], -tvec[:, 0]], axis=1)
            meas = np.linalg.norm(tvec, axis=1)
        else:
            n = np.cross(fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0])
            meas = 0.5 * np.linalg.norm(n, axis=1)
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
        flip = np.einsum("fd,fd->f", n, fv[:, 0] - opp) < 0
        n[flip] *= -1
        return n, meas

    # ------------------------------------------------------------------ rigid motions
    def rigid_motions(self):
        """rm_basis_L2.py:10-74 as closures Z_i(x) -> [npts, d]."""
        d = self.d
        vol_c = self.mesh.cell_volumes()
        volume = vol_c.sum()
        xc = self.mesh.coords[self.mesh.cells]
        cen = (xc.mean(axis=1) * vol_c[:, None]).sum(0) / volume   # exact for linear x
        pts, wts = simplex_quadrature(d, 2)
        lam = np.concatenate([1 - pts.sum(1, keepdims=True), pts], axis=1)
        xq = np.einsum("qv,cvd->cqd", lam, xc) - cen
        fact = 2.0 if d == 2 else 6.0
        wq = wts[None, :] * vol_c[:, None] * fact
        if d == 2:
            r = np.einsum("cq,cqd,cqd->", wq, xq, xq)
            s, a = np.sqrt(volume), np.sqrt(r)
            return [lambda x: np.tile([1 / s, 0.0], (x.shape[0], 1)),
                    lambda x: np.tile([0.0, 1 / s], (x.shape[0], 1)),
                    lambda x: np.stack([-(x[:, 1] - cen[1]) / a, (x[:, 0] - cen[0]) / a], 1)]
        R = np.zeros((3, 3))
        eye = np.eye(3)
        for i in range(3):
            ci = np.cross(xq, eye[i])
            for j in range(i, 3):
                cj = np.cross(xq, eye[j])
                R[i, j] = R[j, i] = np.einsum("cq,cqd,cqd->", wq, ci, cj)
        eigw, eigv = np.linalg.eigh(R)
        eigv = eigv.T
        Z = [(lambda x, v=v: np.tile(v / np.sqrt(volume), (x.shape[0], 1))) for v in eigv]
        Z += [(lambda x, v=v, w=w: np.cross(x - cen, v) / np.sqrt(w)) for v, w in zip(eigv, eigw)]
        return Z

    # ------------------------------------------------------------------ element tables
    def _tables(self, qdeg):
        d = self.d
        pts, wts = simplex_quadrature(d, qdeg)
        N2, dN2 = tabulate(d, 2, pts)
        N1, dN1 = tabulate(d, 1, pts)
        return pts, wts, N2, dN2, N1, dN1

    def _Kc(self, i, cells):
        """K_i on ``cells``: the scalar, or its DG0 values broadcast over the element matrix."""
        if i in self.K_cell:
            return self.K_cell[i][cells][:, None, None]
        return self.K[i]

    def element_blocks(self, cells):
        """Per-cell integrals used by every form (degree-2 exact)."""
        d = self.d
        Jinv, detJ = self.geometry()
        Jinv, detJ = Jinv[cells], detJ[cells]
        pts, wts, N2, dN2, N1, dN1 = self._tables(2)
        g2 = np.einsum("qak,ckm->cqam", dN2, Jinv)     # d phi_a / d x_m
        g1 = np.einsum("qak,ckm->cqam", dN1, Jinv)
        w = wts[None, :] * detJ[:, None]
        G = np.einsum("cq,cqam,cqbn->cabmn", w, g2, g2)   # int d_m phi_a d_n phi_b
        D = np.einsum("cq,cqak,qm->camk", w, g2, N1)      # int d_k phi_a psi_m
        M = np.einsum("cq,qm,qn->cmn", w, N1, N1)         # P1 mass
        Lp = np.einsum("cq,cqmk,cqnk->cmn", w, g1, g1)    # P1 stiffness
        return G, D, M, Lp

    def element_matrix_lhs(self, cells):
        """``a = lhs(F)`` per cell, [nc, nloc, nloc]  (mpetsolver.py:196-201,260)."""
        d, A = self.d, self.A
        sp_ = self.space
        n2, n1, nloc = sp_.n2, sp_.n1, sp_.nloc
        mu, lmbda = convert_to_mu_lmbda(self.E, self.nu)
        dt, th = self.dt, self.theta
        G, D, M, Lp = self.element_blocks(cells)
        Ae = np.zeros((len(cells), nloc, nloc))
        trG = np.einsum("cabmm->cab", G)
        for k in range(d):
            for l in range(d):
                blk = mu * G[:, :, :, l, k] + lmbda * G[:, :, :, k, l]
                if k == l:
                    blk = blk + mu * trG
                Ae[:, k * n2:(k + 1) * n2, l * n2:(l + 1) * n2] = blk
        for i in range(A):
            pi = d * n2 + i * n1
            for k in range(d):
                Ae[:, k * n2:(k + 1) * n2, pi:pi + n1] = -self.alpha[i] * D[:, :, :, k]
                Ae[:, pi:pi + n1, k * n2:(k + 1) * n2] = \
                    -self.alpha[i] * np.swapaxes(D[:, :, :, k], 1, 2)
            offsum = sum(self.S[i, j] for j in range(A) if j != i)
            Ae[:, pi:pi + n1, pi:pi + n1] = (-self.c[i] * M - dt * th * self._Kc(i, cells) * Lp
                                             - dt * th * offsum * M)
            for j in range(A):
                if j != i:
                    pj = d * n2 + j * n1
                    Ae[:, pi:pi + n1, pj:pj + n1] = dt * th * self.S[i, j] * M
        return Ae

    def element_matrix_prev(self, cells):
        """Operator of ``L = rhs(F)`` acting on the previous state, per cell
        (pressure rows only; mpetsolver.py:198-201,261)."""
        d, A = self.d, self.A
        sp_ = self.space
        n2, n1, nloc = sp_.n2, sp_.n1, sp_.nloc
        dt, th = self.dt, self.theta
        G, D, M, Lp = self.element_blocks(cells)
        Be = np.zeros((len(cells), nloc, nloc))
        for i in range(A):
            pi = d * n2 + i * n1
            for k in range(d):
                Be[:, pi:pi + n1, k * n2:(k + 1) * n2] = \
                    -self.alpha[i] * np.swapaxes(D[:, :, :, k], 1, 2)
            offsum = sum(self.S[i, j] for j in range(A) if j != i)
            Be[:, pi:pi + n1, pi:pi + n1] = (-self.c[i] * M + dt * (1 - th) * self._Kc(i, cells) * Lp
                                             + dt * (1 - th) * offsum * M)
            for j in range(A):
                if j != i:
                    pj = d * n2 + j * n1
                    Be[:, pi:pi + n1, pj:pj + n1] = -dt * (1 - th) * self.S[i, j] * M
        return Be

    def element_matrix_prec(self, cells, total_pressure_mass=False):
        """Block-diagonal SPD preconditioner form (intended mpetsolver.py:268-272)."""
        d, A = self.d, self.A
        sp_ = self.space
        n2, n1, nloc = sp_.n2, sp_.n1, sp_.nloc
        mu, lmbda = convert_to_mu_lmbda(self.E, self.nu)
        dt, th = self.dt, self.theta
        G, D, M, Lp = self.element_blocks(cells)
        Pe = np.zeros((len(cells), nloc, nloc))
        trG = np.einsum("cabmm->cab", G)
        for k in range(d):
            Pe[:, k * n2:(k + 1) * n2, k * n2:(k + 1) * n2] = mu * trG
        for i in range(A):
            pi = d * n2 + i * n1
            offsum = sum(self.S[i, j] for j in range(A) if j != i)
            mass = self.c[i] + dt * th * offsum
            if total_pressure_mass:
                mass += self.alpha[i] ** 2 / lmbda
            Pe[:, pi:pi + n1, pi:pi + n1] = mass * M + dt * th * self._Kc(i, cells) * Lp
        return Pe

    # ------------------------------------------------------------------ global assembly
    def _assemble_cells(self, elem_fn, chunk=4096):
        N = self.space.N
        cd = self.space.cell_dofs
        nc = self.mesh.num_cells
        acc = None
        for c0 in range(0, nc, chunk):
            cells = np.arange(c0, min(nc, c0 + chunk))
            Ae = elem_fn(cells)
            rows = np.repeat(cd[cells][:, :, None], cd.shape[1], axis=2)
            cols = np.repeat(cd[cells][:, None, :], cd.shape[1], axis=1)
            part = sp.coo_matrix((Ae.ravel(), (rows.ravel(), cols.ravel())), shape=(N, N)).tocsr()
            acc = part if acc is None else acc + part
        return acc

    def pattern(self):
        """CSR sparsity = union over cells of celldofs x celldofs, structural zeros
        kept, columns ascending (DOLFIN SparsityPatternBuilder + PETSc AIJ [EXT]).
        Real dofs (if any) couple with every field dof they multiply."""
        cd = self.space.cell_dofs
        N = self.space.N
        rows = np.repeat(cd[:, :, None], cd.shape[1], axis=2).ravel()
        cols = np.repeat(cd[:, None, :], cd.shape[1], axis=1).ravel()
        P = sp.coo_matrix((np.ones(rows.shape[0], dtype=np.int8), (rows, cols)), shape=(N, N)).tocsr()
        P.sort_indices()
        return P.indptr.astype(np.int64), P.indices.astype(np.int32)

    def on_pattern(self, M):
        """Return M stored on the full pattern (explicit structural zeros kept)."""
        indptr, indices = self.pattern()
        N = self.space.N
        rows = np.repeat(np.arange(N), np.diff(indptr))
        Mc = M.tocoo()
        out = sp.coo_matrix((np.concatenate([Mc.data, np.zeros(indices.shape[0])]),
                             (np.concatenate([Mc.row, rows]),
                              np.concatenate([Mc.col, indices]))), shape=(N, N)).tocsr()
        out.sort_indices()
        assert np.array_equal(out.indptr, indptr) and np.array_equal(out.indices, indices)
        return out

    def assemble_lhs(self):
        """``A = assemble(a) + sum_i assemble(a_robin[i])`` (mpetsolver.py:412-415),
        plus the Lagrange-multiplier border when nullspaces are flagged."""
        A = self._assemble_cells(self.element_matrix_lhs)
        A = A + self._robin_matrix() + self._nullspace_border()
        A = A.tocsr()
        A.sort_indices()
        return A

    def assemble_prev_operator(self):
        B = self._assemble_cells(self.element_matrix_prev).tocsr()
        B.sort_indices()
        return B

    def assemble_prec(self, **kw):
        P = self._assemble_cells(lambda c: self.element_matrix_prec(c, **kw)).tocsr()
        if self.space.nreal:
            P = P + sp.diags(np.r_[np.zeros(self.space.nfe), np.ones(self.space.nreal)])
        P = P.tocsr()
        P.sort_indices()
        return P

    def _robin_matrix(self):
        """``-dt*theta*beta_i int_{marker 2} p_i q_i`` (mpetsolver.py:252-253)."""
        N = self.space.N
        out = sp.csr_matrix((N, N))
        for i in range(self.A):
            sel = np.nonzero(self.continuity_markers[i] == ROBIN_MARKER)[0]
            if sel.size == 0 or self.beta[i].is_zero(self.t):
                continue
            beta = float(self.beta[i].const(self.t))
            _, meas = self.facet_geometry(sel)
            d = self.d
            Mref = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))   # P1 facet mass / measure
            fv = self.facets["vertices"][sel]
            dofs = self.space._p(d * self.space.N2 + (i + self.P_SHIFT) * self.space.Nv + fv)
            vals = -self.dt * self.theta * beta * meas[:, None, None] * Mref[None]
            rows = np.repeat(dofs[:, :, None], d, axis=2)
            cols = np.repeat(dofs[:, None, :], d, axis=1)
            out = out + sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(N, N)).tocsr()
        return out

    def _nullspace_border(self):
        """mpetsolver.py:203-215: rows/cols of the Real-space multipliers."""
        N = self.space.N
        if self.space.nreal == 0:
            return sp.csr_matrix((N, N))
        rows, cols, vals = [], [], []
        r = self.space.nfe
        for Zi in self.Z:
            vec = self._cell_load(Coef(fn=lambda x, t, Zi=Zi: Zi(x), degree=None), 0.0, "u", qdeg=3)
            nz = np.nonzero(vec)[0]
            rows += [np.full(nz.size, r), nz]
            cols += [nz, np.full(nz.size, r)]
            vals += [vec[nz], vec[nz]]
            r += 1
        for i, flag in enumerate(self.p_has_nullspace):
            if not flag:
                continue
            vec = self._cell_load(Coef(value=1.0), 0.0, ("p", i + self.P_SHIFT))
            nz = np.nonzero(vec)[0]
            rows += [np.full(nz.size, r), nz]
            cols += [nz, np.full(nz.size, r)]
            vals += [vec[nz], vec[nz]]
            r += 1
        return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                             shape=(N, N)).tocsr()

    # ------------------------------------------------------------------ load vectors
    def _cell_load(self, coef, t, test, qdeg=None, chunk=8192):
        """int coef * test dx over all cells.  test = "u" (vector coef) or ("p", i)."""
        sp_ = self.space
        d = self.d
        b = np.zeros(sp_.N)
        if coef.is_zero(t):
            return b
        tdeg = 2 if test == "u" else 1
        qd = tdeg + coef.qdegree() if qdeg is None else qdeg
        pts, wts = simplex_quadrature(d, qd)
        Nt, _ = tabulate(d, tdeg, pts)
        _, detJ = self.geometry()
        nc = self.mesh.num_cells
        for c0 in range(0, nc, chunk):
            cells = np.arange(c0, min(nc, c0 + chunk))
            v = coef.at_cell_points(self.mesh, cells, pts, t)         # [nc, nq, ...]
            w = wts[None, :] * detJ[cells][:, None]
            if test == "u":
                be = np.einsum("cq,cqk,qa->cka", w, v, Nt)            # [nc, d, n2]
                for k in range(d):
                    np.add.at(b, sp_._p(k * sp_.N2 + sp_.cell_nodes2[cells]), be[:, k])
            else:
                i = test[1]
                be = np.einsum("cq,cq,qm->cm", w, v, Nt)
                np.add.at(b, sp_._p(d * sp_.N2 + i * sp_.Nv + self.mesh.cells[cells]), be)
        return b

    def _facet_load(self, coef, t, test, sel, times_normal=False):
        """int coef * test ds over exterior facets ``sel`` (ds(marker))."""
        sp_ = self.space
        d = self.d
        b = np.zeros(sp_.N)
        if sel.size == 0 or coef.is_zero(t):
            return b
        tdeg = 2 if test == "u" else 1
        qd = tdeg + coef.qdegree()
        fpts, fw = simplex_quadrature(d - 1, qd)                     # on reference facet
        F = self.facets
        cells, loc = F["cell"][sel], F["local"][sel]
        n, meas = self.facet_geometry(sel)
        ref_meas = 1.0 if d == 2 else 0.5
        verts = np.vstack([np.zeros(d), np.eye(d)])
        for lf in range(d + 1):
            m = np.nonzero(loc == lf)[0]
            if m.size == 0:
                continue
            fv = verts[[k for k in range(d + 1) if k != lf]]         # facet vertices in cell ref coords
            lamf = np.concatenate([1 - fpts.sum(1, keepdims=True), fpts], axis=1)
            cpts = lamf @ fv                                         # [nq, d]
            v = coef.at_cell_points(self.mesh, cells[m], cpts, t)
            if times_normal:
                v = np.einsum("cq...d,cd->cq...", v, n[m]) if v.ndim == 4 else v[..., None] * n[m][:, None, :]
            Nt, _ = tabulate(d, tdeg, cpts)
            w = fw[None, :] / ref_meas * meas[m][:, None]
            if test == "u":
                be = np.einsum("cq,cqk,qa->cka", w, v, Nt)
                for k in range(d):
                    np.add.at(b, sp_._p(k * sp_.N2 + sp_.cell_nodes2[cells[m]]), be[:, k])
            else:
                i = test[1]
                be = np.einsum("cq,cq,qm->cm", w, v, Nt)
                np.add.at(b, sp_._p(d * sp_.N2 + i * sp_.Nv + self.mesh.cells[cells[m]]), be)
        return b

    def assemble_L(self, B=None):
        """``assemble(L)``: previous-state load (mpetsolver.py:433)."""
        if B is None:
            B = self.assemble_prev_operator()
        return B @ self.up_

    def assemble_L1(self, i, t):
        """``assemble(L1[i])`` at time t (mpetsolver.py:249-254,440-442)."""
        dt, th = self.dt, self.theta
        f = i + self.P_SHIFT
        b = dt * self._cell_load(self.g[i], t, ("p", f))
        sel = np.nonzero(self.continuity_markers[i] == NEUMANN_MARKER)[0]
        b += dt * self._facet_load(self.I[i], t, ("p", f), sel)
        sel = np.nonzero(self.continuity_markers[i] == ROBIN_MARKER)[0]
        if sel.size and not self.beta[i].is_zero(t):
            beta = float(self.beta[i].const(t))
            b -= dt * beta * self._facet_load(self.p_robin[i], t, ("p", f), sel)
            if th != 1.0:
                pdofs = self.space.p_dofs(f)
                pprev = self.up_[pdofs]
                # (1-theta) * beta * p_prev is P1: integrate exactly with the facet mass
                fv = self.facets["vertices"][sel]
                _, meas = self.facet_geometry(sel)
                d = self.d
                Mref = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
                be = np.einsum("f,mn,fn->fm", meas, Mref, pprev[fv])
                np.add.at(b, pdofs[fv], dt * beta * (1 - th) * be)
        return b

    def assemble_L0(self, t):
        """``assemble(L0)`` at time t (mpetsolver.py:233,448)."""
        b = self._cell_load(self.f, t, "u")
        sel = np.nonzero(self.momentum_markers == NEUMANN_MARKER)[0]
        b += self._facet_load(self.s, t, "u", sel, times_normal=self.s_times_normal)
        return b

    # ------------------------------------------------------------------ Dirichlet
    def dirichlet(self, t):
        """``create_dirichlet_bcs`` + ``get_boundary_values``: (dofs, values) with the
        boundary data evaluated at time t; topological search over facets marked 0."""
        sp_ = self.space
        d = self.d
        x2 = sp_.node2_coords()
        dofs, vals = [], []
        sel = np.nonzero(self.momentum_markers == DIRICHLET_MARKER)[0]
        if sel.size:
            F = {k: v[sel] for k, v in self.facets.items()}
            nodes = np.unique(sp_.facet_nodes2(F))
            ub = self.u_bar.at_points(x2[nodes], t)
            for k in range(d):
                dofs.append(sp_._p(k * sp_.N2 + nodes))
                vals.append(ub[:, k])
        for i in range(self.A):
            sel = np.nonzero(self.continuity_markers[i] == DIRICHLET_MARKER)[0]
            if sel.size:
                verts = np.unique(self.facets["vertices"][sel])
                dofs.append(sp_._p(d * sp_.N2 + (i + self.P_SHIFT) * sp_.Nv + verts))
                vals.append(self.p_bar[i].at_points(self.mesh.coords[verts], t))
        if not dofs:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        return np.concatenate(dofs), np.concatenate(vals)

    @staticmethod
    def apply_bc_matrix(A, dofs):
        """``bc.apply(A)``: zero rows, unit diagonal (non-symmetric) [EXT MatZeroRows]."""
        A = A.tocsr(copy=True)
        mask = np.zeros(A.shape[0], dtype=bool)
        mask[dofs] = True
        rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
        A.data[mask[rows]] = 0.0
        A.data[mask[rows] & (rows == A.indices)] = 1.0
        return A

    @staticmethod
    def apply_bc_symmetric(A, dofs, b=None):
        """``apply_symmetric(bc, A, b)`` (bc_symmetric.py:11-22) == PETSc
        MatZeroRowsColumns(dofs, 1.0, x=b, b=b): b -= A[:, D] x_D, rows+cols zeroed,
        unit diagonal, b_D = x_D (b already holds the boundary values)."""
        A = A.tocsr(copy=True)
        N = A.shape[0]
        mask = np.zeros(N, dtype=bool)
        mask[dofs] = True
        if b is not None:
            xD = np.where(mask, b, 0.0)
            b = b - A @ xD
            b[mask] = xD[mask]
        rows = np.repeat(np.arange(N), np.diff(A.indptr))
        kill = mask[rows] | mask[A.indices]
        A.data[kill] = 0.0
        A.data[mask[rows] & (rows == A.indices)] = 1.0
        return (A, b) if b is not None else A

    # ------------------------------------------------------------------ time loop
    def rhs(self, t0, B=None):
        """b for the step t0 -> t0+dt with BC values inserted (mpetsolver.py:424-453)."""
        t_theta = t0 + self.theta * self.dt
        t1 = t0 + self.dt
        b = self.assemble_L(B)
        for i in range(self.A):
            b = b + self.assemble_L1(i, t_theta)
        b = b + self.assemble_L0(t1)
        dofs, vals = self.dirichlet(t1)
        b[dofs] = vals
        return b, dofs, vals

    def solve_direct(self):
        """Generator twin of ``MPETSolver.solve_direct`` (mpetsolver.py:382-462);
        SuperLU stands in for MUMPS."""
        A = self.assemble_lhs()
        B = self.assemble_prev_operator()
        dofs, _ = self.dirichlet(self.t)
        A = self.apply_bc_matrix(A, dofs)
        A = A.tocsc()
        lu = spla.splu(A)
        while self.t < self.T - 1e-9:
            b, _, _ = self.rhs(self.t, B)
            self.t = self.t + self.dt
            self.up = _refined_solve(lu, A, b)
            yield self.up, self.t
            self.up_ = self.up.copy()

    def step(self):
        """Twin of ``MPETSolver.step`` (mpetsolver.py:317-379): re-assemble, solve once."""
        A = self.assemble_lhs()
        dofs, _ = self.dirichlet(self.t)
        A = self.apply_bc_matrix(A, dofs)
        b, _, _ = self.rhs(self.t)
        self.t = self.t + self.dt
        A = A.tocsc()
        self.up = _refined_solve(spla.splu(A), A, b)
        return self.up

    # ------------------------------------------------------------------ post-processing
    def split(self, up=None):
        up = self.up if up is None else up
        sp_ = self.space
        u = np.stack([up[sp_.u_dofs(k)] for k in range(self.d)], axis=1)
        p = [up[sp_.p_dofs(i)] for i in range(self.nfields)]
        return u, p

    def error_norms(self, up, u_exact, p_exact, t, qdeg=8, grad_u=None, grad_p=None, fields=None):
        """L2 (and H1 when exact gradients are given) errors, high-order quadrature
        (the reference uses ``errornorm(..., degree_rise=5)``).  ``fields``: P1 field index of every
        entry of ``p_exact`` (default 0, 1, ...)."""
        d = self.d
        sp_ = self.space
        pts, wts = simplex_quadrature(d, qdeg)
        N2, dN2 = tabulate(d, 2, pts)
        N1, dN1 = tabulate(d, 1, pts)
        Jinv, detJ = self.geometry()
        xc = self.mesh.coords[self.mesh.cells]
        lam = np.concatenate([1 - pts.sum(1, keepdims=True), pts], axis=1)
        xq = np.einsum("qv,cvd->cqd", lam, xc)
        w = wts[None, :] * detJ[:, None]
        u, p = self.split(up)
        uh = np.einsum("qa,cak->cqk", N2, u[sp_.cell_nodes2])
        ue = u_exact(xq.reshape(-1, d), t).reshape(uh.shape)
        out = {"u_L2": np.sqrt(np.einsum("cq,cqk,cqk->", w, uh - ue, uh - ue))}
        if grad_u is not None:
            g2 = np.einsum("qak,ckm->cqam", dN2, Jinv)
            guh = np.einsum("cqam,cak->cqkm", g2, u[sp_.cell_nodes2])
            gue = grad_u(xq.reshape(-1, d), t).reshape(guh.shape)
            out["u_H1"] = np.sqrt(out["u_L2"] ** 2 + np.einsum("cq,cqkm,cqkm->", w, guh - gue, guh - gue))
        out["p_L2"], out["p_H1"] = [], []
        fields = list(range(len(p_exact))) if fields is None else fields
        p = [p[f] for f in fields]
        for i in range(len(p_exact)):
            ph = np.einsum("qm,cm->cq", N1, p[i][self.mesh.cells])
            pe = p_exact[i](xq.reshape(-1, d), t).reshape(ph.shape)
            l2 = np.sqrt(np.einsum("cq,cq,cq->", w, ph - pe, ph - pe))
            out["p_L2"].append(l2)
            if grad_p is not None:
                g1 = np.einsum("qak,ckm->cqam", dN1, Jinv)
                gph = np.einsum("cqam,ca->cqm", g1, p[i][self.mesh.cells])
                gpe = grad_p[i](xq.reshape(-1, d), t).reshape(gph.shape)
                out["p_H1"].append(np.sqrt(l2 ** 2 + np.einsum("cq,cqm,cqm->", w, gph - gpe, gph - gpe)))
        return out


# ---------------------------------------------------------------------------------- iterative twin
def _solve_iterative(self, rtol=1e-5, atol=1e-50, maxit=10000, monitor=None, reassemble=False):
    """Generator twin of ``MPETSolver.solve_iterative`` (mpetsolver.py:465-569): MINRES preconditioned
    by the block-diagonal AMG V-cycle (oracle/krylov.py).  ``reassemble=True`` re-assembles A every
    step like ``MPETSolver.step`` (mpetsolver.py:335)."""
    from .krylov import minres, BlockAMG
    A = self.assemble_lhs()
    B = self.assemble_prev_operator()
    dofs, _ = self.dirichlet(self.t)
    M = BlockAMG(self, dofs, light=rtol >= 1e-8)
    mask = np.zeros(self.space.N, dtype=bool)
    mask[dofs] = True
    self.up = self.up_.copy()
    while self.t < self.T - 1e-9:
        if reassemble:
            A = self.assemble_lhs()
        b, _, vals = self.rhs(self.t, B)
        x0 = self.up.copy()
        x0[dofs] = vals
        self.t = self.t + self.dt
        self.up, info = minres(A, b, x0, M, mask=mask, rtol=rtol, atol=atol, maxit=maxit)
        if monitor is not None:
            monitor.append(info)
        yield self.up, self.t
        self.up_ = self.up.copy()


MPETOracle.solve_iterative = _solve_iterative


# ---------------------------------------------------------------------------------- total pressure
class MPETTotalPressureOracle(MPETOracle):
    """Oracle twin of ``MPETTotalPressureSolver`` (mpettotalpressuresolver.py:164-328): unknowns
    (u, p0, p_1..p_J) with p0 the total pressure; P1 field 0 is p0, field i + 1 is network i.

    F (mpettotalpressuresolver.py:267-274), split with lhs/rhs:
      2 mu (eps(u), eps(v)) + (p0, div v)
      + (div u, w0) - 1/lambda (sum_i alpha_i p_i + p0, w0)
      + sum_i [ -c_i (p_i - p_i^-, w_i) - alpha_i/lambda (p0 - p0^- + sum_j alpha_j (p_j - p_j^-), w_i)
                - dt K_i (grad pm_i, grad w_i) - dt sum_j S_ij (pm_i - pm_j, w_i) ],   pm = theta p + (1-theta) p^-.
    The time loop (rhs groups and their times) is the standard one; see ``MPETOracle.rhs``."""

    P_SHIFT = 1

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        assert not self.u_has_nullspace and not any(self.p_has_nullspace), \
            "nullspace multipliers are not restated for the total-pressure twin"

    def element_matrix_lhs(self, cells):
        """``a = lhs(F)`` per cell (mpettotalpressuresolver.py:267-274)."""
        d, J = self.d, self.A
        sp_ = self.space
        n2, n1, nloc = sp_.n2, sp_.n1, sp_.nloc
        mu, lmbda = convert_to_mu_lmbda(self.E, self.nu)
        dt, th = self.dt, self.theta
        G, D, M, Lp = self.element_blocks(cells)
        Ae = np.zeros((len(cells), nloc, nloc))
        trG = np.einsum("cabmm->cab", G)
        for k in range(d):
            for l in range(d):
                blk = mu * G[:, :, :, l, k]               # 2 mu eps:eps = mu (grad u:grad v + grad u^T:grad v)
                if k == l:
                    blk = blk + mu * trG
                Ae[:, k * n2:(k + 1) * n2, l * n2:(l + 1) * n2] = blk
        p0 = d * n2
        for k in range(d):
            Ae[:, k * n2:(k + 1) * n2, p0:p0 + n1] = D[:, :, :, k]                       # + p0 div v
            Ae[:, p0:p0 + n1, k * n2:(k + 1) * n2] = np.swapaxes(D[:, :, :, k], 1, 2)    # + div u w0
        Ae[:, p0:p0 + n1, p0:p0 + n1] = -(1.0 / lmbda) * M
        for i in range(J):
            pi = d * n2 + (i + 1) * n1
            Ae[:, p0:p0 + n1, pi:pi + n1] = -(self.alpha[i] / lmbda) * M                 # row w0
            Ae[:, pi:pi + n1, p0:p0 + n1] = -(self.alpha[i] / lmbda) * M                 # row w_i, col p0
            offsum = sum(self.S[i, j] for j in range(J) if j != i)
            for j in range(J):
                pj = d * n2 + (j + 1) * n1
                blk = -(self.alpha[i] * self.alpha[j] / lmbda) * M
                if j == i:
                    blk = blk - self.c[i] * M - dt * th * self._Kc(i, cells) * Lp - dt * th * offsum * M
                else:
                    blk = blk + dt * th * self.S[i, j] * M
                Ae[:, pi:pi + n1, pj:pj + n1] = blk
        return Ae

    def element_matrix_prev(self, cells):
        """Operator of ``L = rhs(F)`` on the previous state (rows w_i only; w0 and v rows are zero)."""
        d, J = self.d, self.A
        sp_ = self.space
        n2, n1, nloc = sp_.n2, sp_.n1, sp_.nloc
        _, lmbda = convert_to_mu_lmbda(self.E, self.nu)
        dt, th = self.dt, self.theta
        G, D, M, Lp = self.element_blocks(cells)
        Be = np.zeros((len(cells), nloc, nloc))
        p0 = d * n2
        for i in range(J):
            pi = d * n2 + (i + 1) * n1
            Be[:, pi:pi + n1, p0:p0 + n1] = -(self.alpha[i] / lmbda) * M
            offsum = sum(self.S[i, j] for j in range(J) if j != i)
            for j in range(J):
                pj = d * n2 + (j + 1) * n1
                blk = -(self.alpha[i] * self.alpha[j] / lmbda) * M
                if j == i:
                    blk = blk - self.c[i] * M + dt * (1 - th) * self._Kc(i, cells) * Lp + dt * (1 - th) * offsum * M
                else:
                    blk = blk - dt * (1 - th) * self.S[i, j] * M
                Be[:, pi:pi + n1, pj:pj + n1] = blk
        return Be

    def element_matrix_prec(self, cells, **kw):
        """``prec = pu + pp + ppt`` (mpettotalpressuresolver.py:276-283)."""
        d, J = self.d, self.A
        sp_ = self.space
        n2, n1, nloc = sp_.n2, sp_.n1, sp_.nloc
        mu, lmbda = convert_to_mu_lmbda(self.E, self.nu)
        dt, th = self.dt, self.theta
        G, D, M, Lp = self.element_blocks(cells)
        Pe = np.zeros((len(cells), nloc, nloc))
        trG = np.einsum("cabmm->cab", G)
        for k in range(d):
            Pe[:, k * n2:(k + 1) * n2, k * n2:(k + 1) * n2] = mu * trG
        p0 = d * n2
        Pe[:, p0:p0 + n1, p0:p0 + n1] = M
        for i in range(J):
            pi = d * n2 + (i + 1) * n1
            offsum = sum(self.S[i, j] for j in range(J) if j != i)
            mass = self.alpha[i] ** 2 / lmbda + self.c[i] + dt * th * offsum
            Pe[:, pi:pi + n1, pi:pi + n1] = mass * M + dt * th * self._Kc(i, cells) * Lp
        return Pe

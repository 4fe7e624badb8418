"""Linear-algebra stand-ins bound to the B200 engine: forms, assemble(), Matrix, Krylov / LU solver
objects and the symmetric-BC helpers of the reference's bc_symmetric.py.

Reference call sites mirrored here (src/mpet/mpet/mpetsolver.py):
  assemble(a) :335,412,496 | assemble(a_robin[i]) + A.axpy :336-338 | assemble(L) :356 |
  assemble(L1[i]) :363-365 | assemble(L0) :371 | assemble(prec) :502 | bc.apply(A|b) :343-344,375-376 |
  LUSolver(A,"mumps").solve :347,379 | PETScKrylovSolver("minres","hypre_amg") :507,553-556 |
  apply_symmetric / get_bc_dofs / zero_rows_cols  (src/mpet/mpet/bc_symmetric.py:6-22).
Every numerical operation below is a call into libmpet_b200 (see engine.py); numpy is used only for
O(boundary) bookkeeping (facet lists, boundary values).
"""
import numpy as np
import torch

from .dolfin_shim import Constant, Expression, NormalProduct, Vector, DirichletBC

NEUMANN_MARKER = 1
ROBIN_MARKER = 2

# reference P2 triangle mass / |T| with node order v0, v1, v2, e(0,1), e(0,2), e(1,2)
_M6 = np.array([[6, -1, -1, 0, 0, -4],
                [-1, 6, -1, 0, -4, 0],
                [-1, -1, 6, -4, 0, 0],
                [0, 0, -4, 32, 16, 16],
                [0, -4, 0, 16, 32, 16],
                [-4, 0, 0, 16, 16, 32]], dtype=float) / 180.0
_M3 = (np.ones((3, 3)) + np.eye(3)) / 12.0


class Form:
    """Tag for one of the solver's variational forms (what UFL form objects are in the reference)."""

    def __init__(self, solver, kind, index=None):
        self.solver, self.kind, self.index = solver, kind, index

    def __repr__(self):
        return "Form(%s%s)" % (self.kind, "" if self.index is None else "[%d]" % self.index)


def _host_csr(rows, cols, vals, nrows, device):
    order = np.argsort(rows, kind="stable")
    rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    return (torch.as_tensor(rowptr, device=device), torch.as_tensor(cols.astype(np.int32), device=device),
            torch.as_tensor(vals, device=device))


def _lattice(degree):
    """Nodes of P_degree on the reference tet (degree 0: barycentre), as reference coordinates [nd, 3]."""
    if degree == 0:
        return np.full((1, 3), 0.25)
    pts = [(i, j, k) for i in range(degree + 1) for j in range(degree + 1 - i) for k in range(degree + 1 - i - j)]
    return np.asarray(pts, dtype=float) / degree


def _monomial_coefficients(nodes, degree):
    """Lagrange basis on ``nodes`` in the monomial basis of total degree <= degree: column n = basis function n."""
    expo = [(i, j, k) for i in range(degree + 1) for j in range(degree + 1 - i) for k in range(degree + 1 - i - j)]
    V = np.stack([nodes[:, 0] ** a * nodes[:, 1] ** b * nodes[:, 2] ** c for a, b, c in expo], axis=1)
    return np.linalg.inv(V), expo


def reference_mass(test_degree, data_degree):
    """int over the reference tet of (P_test basis a) * (P_data basis n), exactly (monomial integrals
    a! b! c! / (a + b + c + 3)!).  Test basis in UFC order: vertices, then the edges (2,3), (1,3), (1,2), (0,3),
    (0,2), (0,1); data basis on the lattice of _lattice(data_degree)."""
    from math import factorial
    vert = np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    if test_degree == 1:
        tn = vert
    else:
        mids = [0.5 * (vert[a] + vert[b]) for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))]
        tn = np.vstack([vert, mids])
    Ct, et = _monomial_coefficients(tn, test_degree)
    Cd, ed = _monomial_coefficients(_lattice(data_degree), data_degree)
    G = np.array([[factorial(a + p) * factorial(b + q) * factorial(c + r) / factorial(a + p + b + q + c + r + 3)
                   for (p, q, r) in ed] for (a, b, c) in et])
    return Ct.T @ G @ Cd


class CellLoadOperator:
    """int coef * test dx for a coefficient whose DOLFIN interpolation degree d differs from the test space's:
    ``Expression(..., degree=d)`` is interpolated CELL BY CELL into P_d (discontinuous across cells) and the product
    with the test function integrated exactly -- what FFC's quadrature of degree (test + d) does for the
    reference's ``f``, ``g`` (test_convergence_mpetsolver.py:129-131 uses degree 3).  Stored as a device CSR
    operator [test nodes] x [Nc * nd cell-lattice points] with entries |det J_c| * Mhat[a, n], built once per
    (space, degree) on the host and applied with mpet_csr_spmv, like the facet operators."""

    def __init__(self, space, p2, degree, device=None):
        mesh = space.mesh
        cells = mesh.cells.astype(np.int64)
        x = mesh.coordinates
        nc = cells.shape[0]
        xc = x[cells]                                               # [nc, 4, 3]
        detj = np.abs(np.linalg.det(xc[:, 1:, :] - xc[:, :1, :]))
        if p2:
            nodes = np.concatenate([cells] + [space.Nv + space.edge_index(cells[:, a], cells[:, b])[:, None]
                                              for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))], axis=1)
            nrows = space.N2
        else:
            nodes, nrows = cells, space.Nv
        M = reference_mass(2 if p2 else 1, degree)                  # [nt, nd]
        nt, nd = M.shape
        ref = _lattice(degree)
        lam = np.concatenate([1.0 - ref.sum(axis=1, keepdims=True), ref], axis=1)      # barycentric [nd, 4]
        pts = np.einsum("nv,cvd->cnd", lam, xc).reshape(-1, 3)                         # [nc * nd, 3]
        # lattice points on shared vertices / edges / faces coincide between cells (P3: 120 Nv cell-wise points, ~27 Nv
        # distinct ones): the coefficient is evaluated once per distinct point.  Points are identified on a grid of
        # 1e-10 of the mesh extent -- far below the lattice spacing; two copies of one point that straddle a grid
        # line are simply both evaluated.
        ext = float(np.max(x.max(axis=0) - x.min(axis=0))) or 1.0
        q = np.ascontiguousarray(np.round((pts - x.min(axis=0)) / (1e-10 * ext)).astype(np.int64))
        _, first, inverse = np.unique(q.view([("", np.int64)] * 3).ravel(), return_index=True, return_inverse=True)
        self.points = pts[first]                                                       # distinct points
        self.expand = inverse.ravel()                                                  # cell-lattice point -> distinct
        rows = np.repeat(nodes[:, :, None], nd, axis=2).ravel()
        cols = np.repeat((np.arange(nc)[:, None] * nd + np.arange(nd)[None, :])[:, None, :], nt, axis=1).ravel()
        vals = (detj[:, None, None] * M[None]).ravel()
        self.csr = _host_csr(rows, cols, vals, nrows, device if device is not None else space.engine.device)
        self.nd = nd

    def data(self, coef):
        """Values of ``coef`` at the cell-lattice points (scalar: [nc*nd]; vector: [3, nc*nd])."""
        v = np.asarray(coef.eval_points(self.points), dtype=float)[self.expand]
        return v.T if v.ndim == 2 else v

    def apply(self, engine, data, y, scale=1.0):
        """y += scale * Q data   (y: device slice over the test space's nodes)."""
        d = torch.as_tensor(np.ascontiguousarray(data * scale), device=engine.device)
        engine.csr_spmv(self.csr[0], self.csr[1], self.csr[2], d, y, beta=1.0)


class FacetOperator:
    """int_{facets with marker} data * test ds as a device CSR operator acting on per-facet nodal data.
    Built once per marker set on the host (O(N^(2/3)) facets), applied on the device every step."""

    def __init__(self, space, marker_array, marker_id, p2):
        mesh = space.mesh
        F = mesh.exterior_facets()
        self.sel = np.nonzero(marker_array == marker_id)[0]
        self.p2 = p2
        self.space = space
        nf = self.sel.size
        self.nf = nf
        if nf == 0:
            return
        fv = F["vertices"][self.sel]
        x = mesh.coordinates
        a, b, c = x[fv[:, 0]], x[fv[:, 1]], x[fv[:, 2]]
        nrm = np.cross(b - a, c - a)
        area = 0.5 * np.linalg.norm(nrm, axis=1)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        opp = x[mesh.cells[F["cell"][self.sel], F["local"][self.sel]]]
        flip = np.einsum("fd,fd->f", nrm, a - opp) < 0
        nrm[flip] *= -1
        self.normal, self.area = nrm, area
        if p2:
            nodes = np.concatenate([fv] + [space.Nv + space.edge_index(fv[:, i], fv[:, j])[:, None]
                                           for i, j in ((0, 1), (0, 2), (1, 2))], axis=1)
            M = _M6
            self.points = space.node2_coordinates()[nodes]          # [nf, 6, 3]
            nrows = space.N2
        else:
            nodes = fv
            M = _M3
            self.points = x[nodes]
            nrows = space.Nv
        k = nodes.shape[1]
        self.nodes = nodes
        rows = np.repeat(nodes[:, :, None], k, axis=2).ravel()
        cols = np.repeat((np.arange(nf)[:, None] * k + np.arange(k)[None, :])[:, None, :], k, axis=1).ravel()
        vals = (area[:, None, None] * M[None]).ravel()
        self.csr = _host_csr(rows, cols, vals, nrows, space.engine.device)
        self.k = k

    def data(self, coef):
        """Per-facet nodal values of ``coef`` (scalar: [nf*k]; vector: [3, nf*k])."""
        pts = self.points.reshape(-1, 3)
        if isinstance(coef, NormalProduct):
            v = np.asarray(coef.expr.eval_points(pts), dtype=float)
            n = np.repeat(self.normal, self.k, axis=0)
            if v.ndim == 1:
                return (v[:, None] * n).T
            return np.einsum("pij,pj->ip", v, n)
        v = np.asarray(coef.eval_points(pts), dtype=float)
        return v.T if v.ndim == 2 else v

    def apply(self, engine, data, y, scale=1.0):
        """y += scale * F data   (y: device slice of the right length)."""
        d = torch.as_tensor(np.ascontiguousarray(data * scale), device=engine.device)
        engine.csr_spmv(self.csr[0], self.csr[1], self.csr[2], d, y, beta=1.0)


def _is_zero(coef):
    if isinstance(coef, Constant):
        return not np.any(coef.values())
    if isinstance(coef, NormalProduct) and isinstance(coef.expr, Constant):
        return not np.any(coef.expr.values())
    return False


class RobinEntries:
    """assemble(a_robin[i]): -dt*theta*beta_i int_{marker 2} p q ds as unique (row, col, value) triplets."""

    def __init__(self, rows, cols, vals):
        self.rows, self.cols, self.vals = rows, cols, vals


class Matrix:
    """Handle on the block system (kind "A") or the block-diagonal preconditioner (kind "P") held by
    the engine.  The engine never overwrites assembled values with boundary conditions: applying a
    DirichletBC records the constrained dofs, and the Krylov kernels mask those rows on the fly."""

    def __init__(self, solver, kind):
        self.solver, self.kind = solver, kind
        self.bcs = []
        self.symmetric = False

    def axpy(self, a, other, same_nonzero_pattern=False):
        assert isinstance(other, RobinEntries) and self.kind == "A"
        if other.rows.size:
            self.solver.engine.add_entries(other.rows, other.cols, a * other.vals)

    def copy(self):
        m = Matrix(self.solver, self.kind)
        m.bcs = list(self.bcs)
        return m

    def _apply_dirichlet(self, bc):
        self.bcs.append(bc)

    def size(self, dim):
        return self.solver.engine.sizes["N"]

    def nnz(self):
        return self.solver.engine.sizes["nnz"]

    def to_scipy(self):
        """Host copy as scipy CSR (inspection / tests): boundary conditions recorded on this handle
        are applied the way the reference would have (rows only, or rows + columns)."""
        import scipy.sparse as sp
        eng = self.solver.engine
        if self.bcs:
            dofs = np.unique(np.concatenate([bc.dofs() for bc in self.bcs]))
            eng.set_dirichlet_dofs(dofs.astype(np.int32))
            self.solver._engine_bc_dofs = None        # force a refresh before the next solve
        which = 3 if self.kind == "P" else (0 if not self.bcs else (2 if self.symmetric else 1))
        rowptr, cols = eng.pattern()
        vals = eng.values(which)
        N = eng.sizes["N"]
        return sp.csr_matrix((vals.cpu().numpy(), cols.cpu().numpy(), rowptr.cpu().numpy()), shape=(N, N))


class AssembledVector(Vector):
    def _apply_dirichlet(self, bc):
        d = bc.dofs()
        if d.size:
            self.t[torch.as_tensor(d, device=self.t.device)] = torch.as_tensor(bc.values(), device=self.t.device)


def assemble(form):
    """Device assembly of one of the solver's forms."""
    return form.solver._assemble(form)


# bc_symmetric.py mirror: lives in its own module like the reference's, re-exported here for the solvers
from .bc_symmetric import get_bc_dofs, zero_rows_cols, apply_symmetric  # noqa: E402,F401


# --------------------------------------------------------------------------- solver objects
class PETScKrylovSolver:
    """``PETScKrylovSolver("minres", "hypre_amg")`` (mpetsolver.py:507): MINRES/GMRES with the
    block-diagonal AMG V-cycle, PETSc's default tolerances unless ``parameters`` are changed."""

    def __init__(self, method="minres", preconditioner="hypre_amg"):
        self.method = method
        self.pc = {"hypre_amg": "amg", "amg": "amg", "jacobi": "jacobi", "none": "none"}[preconditioner]
        self.parameters = {"relative_tolerance": 1e-5, "absolute_tolerance": 1e-50,
                           "maximum_iterations": 10000, "nonzero_initial_guess": False,
                           "gmres_restart": 30, "error_on_nonconvergence": True,
                           # reference norm of the test: "b" = PETSc default, "min_b_r0" = min(|b|, |r0|)
                           "convergence_norm": "b",
                           # a solve that stagnates (round-off floor) below this |r|/|b| counts as converged
                           "accept_stagnation_below": 0.0}
        self.A = self.P = None
        self.last_info = None

    def set_operators(self, A, P):
        self.A, self.P = A, P

    def solve(self, x, b):
        solver = self.A.solver
        eng = solver.engine
        p = self.parameters
        cfg = (self.method, self.pc, p["relative_tolerance"], p["absolute_tolerance"],
               p["maximum_iterations"], p["gmres_restart"], p["convergence_norm"])
        solver._sync_dirichlet(self.A.bcs)
        if cfg != solver._krylov_cfg or solver._pc_dirty:
            eng.krylov_setup(self.method, self.pc, rtol=p["relative_tolerance"], atol=p["absolute_tolerance"],
                             maxit=p["maximum_iterations"], restart=p["gmres_restart"],
                             reference_norm=p["convergence_norm"])
            if self.pc != "none":
                import time as _time
                torch.cuda.synchronize(eng.device)
                t0 = _time.perf_counter()
                eng.pc_setup()                      # synchronous: hierarchy set-up (once per mesh / dt / Dirichlet set)
                solver.solver_monitor["pc_setup_s"] = _time.perf_counter() - t0
            solver._pc_dirty = False
            solver._krylov_cfg = cfg
        xt = x.t if isinstance(x, Vector) else x
        bt = b.t if isinstance(b, Vector) else b
        if not p["nonzero_initial_guess"]:
            xt.zero_()
        solver.solver_monitor["method"] = self.method
        self.last_info = info = eng.solve(bt, xt)
        rel_b = info["rel_res"] * info["refnorm"] / info["bnorm"] if info["bnorm"] > 0 else 0.0
        if not info["converged"] and info["reason"] == -5 and rel_b <= p["accept_stagnation_below"]:
            info["converged"] = True        # round-off floor reached: as good as this arithmetic gets
        if not info["converged"] and p["error_on_nonconvergence"]:
            # DOLFIN's default (error_on_nonconvergence=True): a failed Krylov solve must not be stepped over
            why = {-3: "iteration limit reached", -4: "breakdown (indefinite preconditioner)",
                   -5: "stagnation"}.get(info["reason"], "reason %d" % info["reason"])
            raise RuntimeError("Krylov solve (%s) did not converge after %d iterations: %s; ||r||/||b|| = %.3e"
                               % (self.method, info["niter"], why, info["rel_res"]))
        return info["niter"]


class LUSolver:
    """``LUSolver(A, "mumps")`` (mpetsolver.py:347,422).  A sparse LU does not exist on this path
    (it is memory-infeasible at the target sizes); the same linear system is solved by
    the preconditioned Krylov kernels to a tolerance tight enough to agree with a direct solve to
    1e-8 ("parity" preset).  GMRES is chosen automatically when S is not symmetric."""

    def __init__(self, A, method="mumps"):
        self.A = A
        solver = A.solver
        sym = solver._exchange_is_symmetric()
        self.krylov = PETScKrylovSolver("minres" if sym else "gmres", "hypre_amg")
        self.krylov.parameters.update(relative_tolerance=1e-12, absolute_tolerance=1e-50,
                                      maximum_iterations=50000, nonzero_initial_guess=True, gmres_restart=100,
                                      convergence_norm="min_b_r0", accept_stagnation_below=1e-12)
        solver._ensure_prec()

    def solve(self, A, x, b):
        self.krylov.set_operators(A, None)
        return self.krylov.solve(x, b)

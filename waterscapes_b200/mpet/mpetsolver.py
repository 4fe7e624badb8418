"""MPETSolver on the B200 engine -- same surface as the reference's MPETSolver
(src/mpet/mpet/mpetsolver.py:21-569): default_params, create_function_spaces,
create_variational_forms, create_dirichlet_bcs, solve, step, solve_direct, solve_iterative; attributes
up_, up, dt, a, a_robin, L, L0, L1, prec, bcs, solver_monitor.

What changes underneath: "forms" are tags that the device assembler understands, ``assemble`` runs the
sm_100a element/gather kernels, and both ``solve_direct`` and ``solve_iterative`` end in the device
Krylov solver (see la.LUSolver for why there is no sparse LU).  The time loop, the order in which the
right-hand side is built (L at t0 with up_, L1 at t0+theta*dt, L0 at t0+dt, boundary values at t0+dt)
and the generator protocol are the reference's.
"""
import numpy as np
import torch

from ..engine import Engine
from .dolfin_shim import (Constant, Expression, NormalProduct, Function, FunctionSpace, DirichletBC,
                          Parameters, CellFunction, info, warning)
from .la import (Form, Matrix, AssembledVector, RobinEntries, FacetOperator, CellLoadOperator, assemble, LUSolver,
                 PETScKrylovSolver, apply_symmetric, _is_zero, _M3, NEUMANN_MARKER as _N, ROBIN_MARKER as _R)
from .mpetproblem import convert_to_mu_lmbda

# Marker conventions (mpetsolver.py:17-19)
DIRICHLET_MARKER = 0
NEUMANN_MARKER = 1
ROBIN_MARKER = 2


def _f(v, what="material parameter"):
    """Constant coefficient as a number.  The reference lets E, nu, alpha, c, K, S and beta be UFL coefficients
    (Expression / Function); the kernels here integrate constant coefficients exactly (K additionally as a DG0
    CellFunction), so anything else is refused loudly instead of being mis-read."""
    try:
        return float(v)
    except (TypeError, ValueError):
        raise NotImplementedError("%s must be a number or Constant on the B200 path (got %s); spatially varying "
                                  "K is supported as a DG0 CellFunction" % (what, type(v).__name__)) from None


class MPETSolver(object):

    def __init__(self, problem, params=None, device=0, partition=None):
        "Create solver with given MPET problem and parameters."
        self.problem = problem
        self.params = self.default_params()
        if params is not None:
            self.params.update(params)
        self.solver_monitor = {}
        if (problem.u_has_nullspace or any(problem.p_has_nullspace)) and partition is not None:
            raise NotImplementedError("nullspace Lagrange multipliers are single-GPU for now")
        if self.params["u_degree"] != 2 or self.params["p_degree"] != 1:
            raise NotImplementedError("the B200 kernels implement the Taylor-Hood pair P2-P1 only")
        self.engine = Engine(device)
        self.partition = partition          # waterscapes_b200.parallel.Partition (multi-GPU) or None
        self._engine_bc_dofs = None
        self._cell_ops = {}                 # (scalar space, degree) -> CellLoadOperator
        self._pc_dirty = True
        self._krylov_cfg = None
        self._prec_assembled_for = None
        self._facet_ops = {}
        self._lumped = {}
        self._border_ready = False
        self.create_variational_forms()
        if self.partition is not None:
            self.partition.setup(self.engine, self.VQ, self.problem.mesh)
        self.create_dirichlet_bcs()

    # ------------------------------------------------------------------ reference surface
    @staticmethod
    def default_params():
        "Define default solver parameters (mpetsolver.py:86-98)."
        params = Parameters("MPETSolver")
        params.add("dt", 0.1)
        params.add("t", 0.0)
        params.add("T", 1.0)
        params.add("theta", 1.0)
        params.add("u_degree", 2)
        params.add("p_degree", 1)
        params.add("direct_solver", True)
        params.add("testing", False)
        # B200-path extras (not in the reference): Krylov controls of the iterative branch
        params.add("krylov_rtol", 1e-5)      # PETSc default [EXT]; the reference sets none
        params.add("krylov_atol", 1e-50)
        params.add("krylov_maxit", 10000)
        return params

    # sub-space index of network i in the mixed space (0 is the displacement) and number of P1 fields;
    # the total-pressure solver shifts both by one (its sub-space 1 is the total pressure)
    _first_network_sub = 1

    def _num_p1_fields(self):
        return int(self.problem.params["J"])

    def _net_sub(self, i):
        return i + self._first_network_sub

    def create_function_spaces(self, mesh):
        "Mixed space [P2]^3 x [P1]^J; the dof map and sparsity graph are built on the device."
        J = self._num_p1_fields()
        if self.engine.sizes is None:
            self.engine.set_mesh(mesh.coordinates, mesh.cells, J)
        # Real spaces of the reference (mpetsolver.py:115-124): six rigid-motion multipliers when the displacement has
        # a nullspace, one multiplier per pressure that is only determined up to a constant.  They are appended to
        # every dof vector and enter the solver as a dense border (mpet_set_border), not as CSR rows.
        nreal = (6 if self.problem.u_has_nullspace else 0) + sum(bool(f) for f in self.problem.p_has_nullspace)
        return FunctionSpace(mesh, J, self.engine, nreal=nreal)

    def create_variational_forms(self, include_preconditioner=False):
        mesh = self.problem.mesh
        dt = Constant(self.params["dt"])
        J = int(self.problem.params["J"])
        VQ = self.create_function_spaces(mesh)
        self.VQ = VQ
        up_ = Function(VQ)
        up = Function(VQ)
        self.up_ = up_
        self.up = up
        self.dt = dt
        self.a = Form(self, "a")
        self.a_robin = [Form(self, "a_robin", i) for i in range(J)]
        self.L = Form(self, "L")
        self.L0 = Form(self, "L0")
        self.L1 = [Form(self, "L1", i) for i in range(J)]
        # the reference only builds this when asked (and its text is malformed, mpetsolver.py:270);
        # here the well-formed block-diagonal form is always available to the Krylov path
        self.prec = Form(self, "prec")
        forms = (self.a, self.a_robin, self.L, self.L0, self.L1)
        return (forms, self.prec, (up_, up), dt)

    def create_dirichlet_bcs(self):
        """DirichletBCs from the problem's markers (mpetsolver.py:63-84)."""
        VP = self.up.function_space()
        bcs0 = [DirichletBC(VP.sub(0), self.problem.u_bar, self.problem.momentum_boundary_markers,
                            DIRICHLET_MARKER)]
        bcs1 = []
        for i in range(int(self.problem.params["J"])):
            bcs1 += [DirichletBC(VP.sub(i + 1), self.problem.p_bar[i],
                                 self.problem.continuity_boundary_markers[i], DIRICHLET_MARKER)]
        self.bcs = [bcs0, bcs1]
        return [bcs0, bcs1]

    # ------------------------------------------------------------------ engine plumbing
    def _push_params(self):
        p = self.problem.params
        J = int(p["J"])
        # a DG0 permeability travels as per-cell values; its scalar slot is then a multiplier of 1
        dg0 = {i: v for i, v in enumerate(p["K"]) if isinstance(v, CellFunction)}
        Ks = [1.0 if i in dg0 else _f(v) for i, v in enumerate(p["K"])]
        cells_key = tuple((i, hash(v.values.tobytes())) for i, v in sorted(dg0.items()))
        vals = (_f(p["E"], "E"), _f(p["nu"], "nu"), [_f(v, "alpha") for v in p["alpha"]], Ks,
                [[_f(v) for v in row] for row in p["S"]], [_f(v) for v in p["c"]], float(self.dt),
                float(self.params["theta"]))
        if getattr(self, "_pushed", None) != vals + (cells_key,):
            self._engine_set_params(*vals)
            if cells_key != getattr(self, "_cells_key", ()):
                for i in range(J):
                    self.engine.set_cell_coefficient(self._net_sub(i) - 1, dg0[i].values if i in dg0 else None)
                self._cells_key = cells_key
            self._pushed = vals + (cells_key,)
            self._pc_dirty = True
        return J

    def _engine_set_params(self, *vals):
        self.engine.set_params(*vals)

    def _exchange_is_symmetric(self):
        S = np.array([[_f(v) for v in row] for row in self.problem.params["S"]], dtype=float)
        return bool(np.allclose(S, S.T, rtol=0, atol=0))

    def _ensure_prec(self):
        self._push_params()
        key = self._pushed
        if self._prec_assembled_for != key:
            if self.VQ.nreal and not self._border_ready:
                self._set_prec_shifts()
            self.engine.assemble_prec()
            self._prec_assembled_for = key
            self._pc_dirty = True
        if self.VQ.nreal and not self._border_ready:
            self._setup_border()

    # ------------------------------------------------------------------ nullspaces (mpetsolver.py:203-215)
    def _set_prec_shifts(self):
        """Mass shifts that make the preconditioner blocks of the constrained fields definite: the reference's
        (commented-out) ``inner(u, v)*dx`` (mpetsolver.py:276-277) and ``p[k]*q[k]*dx`` (:274), scaled with the
        block's own stiffness coefficient over the squared domain diameter."""
        x = self.problem.mesh.coordinates
        diam2 = float(np.sum((x.max(axis=0) - x.min(axis=0)) ** 2))
        p = self.problem.params
        mu, _ = convert_to_mu_lmbda(float(p["E"]), float(p["nu"]))
        dth = float(self.dt) * float(self.params["theta"])
        nf = self._num_p1_fields()
        shift_p = [0.0] * nf
        for i, flag in enumerate(self.problem.p_has_nullspace):
            if flag:
                shift_p[self._net_sub(i) - 1] = 10.0 * dth * float(p["K"][i]) / diam2
        self.engine.set_prec_shift(10.0 * mu / diam2 if self.problem.u_has_nullspace else 0.0, shift_p)

    def _setup_border(self):
        """Columns of the Lagrange-multiplier border: c_i = int Z_i . v dx for the rigid motions Z_i
        (rm_basis_L2.py), c_k = int q_i dx for a pressure with nullspace."""
        from .rm_basis_L2 import rigid_motions
        sp, eng = self.VQ, self.engine
        C = torch.zeros((sp.nreal, sp.N), dtype=torch.float64, device=eng.device)
        r = 0
        if self.problem.u_has_nullspace:
            x2 = sp.node2_coordinates()
            for Z in rigid_motions(self.problem.mesh):
                z = np.asarray(Z(x2), dtype=float)
                for k in range(3):
                    zk = torch.as_tensor(np.ascontiguousarray(z[:, k]), device=eng.device)
                    eng.mass_apply(2, 1.0, zk, C[r, k * sp.N2:(k + 1) * sp.N2])
                r += 1
        for i, flag in enumerate(self.problem.p_has_nullspace):
            if flag:
                lo, hi = sp.sub_range(self._net_sub(i))
                C[r, lo:hi] = self._lumped_vec(1)
                r += 1
        eng.set_border(C)
        self._border_ready = True
        self._pc_dirty = True

    def _sync_dirichlet(self, bcs):
        """Constrained dof set -> engine (only when it changed)."""
        dofs = [bc.dofs() for bc in bcs]
        dofs = np.concatenate(dofs) if dofs else np.zeros(0, dtype=np.int64)
        key = dofs.tobytes()
        if key != self._engine_bc_dofs:
            assert np.unique(dofs).size == dofs.size
            self.engine.set_dirichlet_dofs(dofs.astype(np.int32))
            self._engine_bc_dofs = key
            self._pc_dirty = True
            vals = [bc.values() for bc in bcs]
            self.engine.set_dirichlet_values(np.concatenate(vals) if vals else np.zeros(0))

    def _push_dirichlet_values(self, bcs):
        vals = [bc.values() for bc in bcs]
        self.engine.set_dirichlet_values(np.concatenate(vals) if vals else np.zeros(0))

    def _facet_op(self, which, index, marker_id, p2):
        markers = (self.problem.momentum_boundary_markers if which == "m"
                   else self.problem.continuity_boundary_markers[index])
        key = (which, index, marker_id, p2, markers.array().tobytes())
        op = self._facet_ops.get(key)
        if op is None:
            op = FacetOperator(self.VQ, markers.array(), marker_id, p2)
            self._facet_ops[key] = op
        return op

    def _lumped_vec(self, space):
        if space not in self._lumped:
            self._lumped[space] = self.engine.lumped(space)
        return self._lumped[space]

    def _cell_load(self, coef, space, y, scale):
        """y += scale * int coef * test dx for the scalar space (2: P2, 1: P1); y is a device slice."""
        pts = self.VQ.node2_coordinates() if space == 2 else self.problem.mesh.coordinates
        if isinstance(coef, Constant):
            v = float(coef)
            if v != 0.0:
                y.add_(self._lumped_vec(space), alpha=scale * v)
        elif self._cell_op_degree(coef, space) is not None:
            op = self._cell_op(space, self._cell_op_degree(coef, space))
            op.apply(self.engine, op.data(coef), y, scale=scale)
        else:
            vals = torch.as_tensor(np.asarray(coef.eval_points(pts), dtype=float), device=y.device)
            self.engine.mass_apply(space, scale, vals, y)

    @staticmethod
    def _cell_op_degree(coef, space):
        """Interpolation degree of a coefficient that needs the cell-lattice operator: DOLFIN interpolates
        ``Expression(degree=d)`` into P_d cell by cell; for d == the test space's degree (or no degree given) that is
        the nodal data the mass operator already takes."""
        d = getattr(coef, "degree", None)
        if d is None or int(d) == (2 if space == 2 else 1):
            return None
        return int(d)

    def _cell_op(self, space, degree):
        key = (space, degree)
        if key not in self._cell_ops:
            self._cell_ops[key] = CellLoadOperator(self.VQ, space == 2, degree)
        return self._cell_ops[key]

    # ------------------------------------------------------------------ pieces of the forms
    def _robin_entries(self, i):
        """-dt*theta*beta_i int_{marker 2} p_i q_i ds as unique (row, col, value) triplets
        (mpetsolver.py:252-253; lhs(L2[i]) in mpettotalpressuresolver.py:323)."""
        sp = self.VQ
        dt, theta = float(self.dt), float(self.params["theta"])
        op = self._facet_op("c", i, ROBIN_MARKER, False)
        beta = _f(self.problem.beta[i], "the Robin coefficient beta")
        if op.nf == 0 or beta == 0.0:
            return RobinEntries(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
        lo, _ = sp.sub_range(self._net_sub(i))
        rows = np.repeat(op.nodes[:, :, None], 3, axis=2).ravel() + lo
        cols = np.repeat(op.nodes[:, None, :], 3, axis=1).ravel() + lo
        vals = (-dt * theta * beta * op.area[:, None, None] * _M3[None]).ravel()
        key = rows * sp.N + cols
        uk, inv = np.unique(key, return_inverse=True)
        acc = np.zeros(uk.shape[0])
        np.add.at(acc, inv, vals)
        return RobinEntries((uk // sp.N).astype(np.int32), (uk % sp.N).astype(np.int32), acc)

    def _load_sources(self, i, b):
        """b += dt*int g_i q dx + dt*int_{marker 1} I_i q ds on network i's rows (mpetsolver.py:249)."""
        sp, eng = self.VQ, self.engine
        dt = float(self.dt)
        lo, hi = sp.sub_range(self._net_sub(i))
        y = b[lo:hi]
        self._cell_load(self.problem.g[i], 1, y, dt)
        if not _is_zero(self.problem.I[i]):
            op = self._facet_op("c", i, NEUMANN_MARKER, False)
            if op.nf:
                op.apply(eng, op.data(self.problem.I[i]), y, scale=dt)

    def _load_robin_rhs(self, i, b):
        """b += dt*beta_i int_{marker 2} ((1-theta) p_i^- - p_robin_i) q ds (mpetsolver.py:254;
        rhs(L2[i]) in mpettotalpressuresolver.py:323)."""
        sp, eng = self.VQ, self.engine
        dt, theta = float(self.dt), float(self.params["theta"])
        beta = _f(self.problem.beta[i], "the Robin coefficient beta")
        if beta == 0.0:
            return
        op = self._facet_op("c", i, ROBIN_MARKER, False)
        if not op.nf:
            return
        lo, hi = sp.sub_range(self._net_sub(i))
        y = b[lo:hi]
        if not _is_zero(self.problem.p_robin[i]):
            op.apply(eng, op.data(self.problem.p_robin[i]), y, scale=-dt * beta)
        if theta != 1.0:
            pprev = self.up_.x[lo:hi][torch.as_tensor(op.nodes.ravel(), device=eng.device)]
            eng.csr_spmv(op.csr[0], op.csr[1], op.csr[2], (dt * beta * (1.0 - theta)) * pprev, y, beta=1.0)

    # ------------------------------------------------------------------ assemble(form)
    def _assemble(self, form):
        eng = self.engine
        J = self._push_params()
        sp = self.VQ
        dt, theta = float(self.dt), float(self.params["theta"])
        kind = form.kind
        if kind == "a":
            eng.assemble_lhs()
            return Matrix(self, "A")
        if kind == "prec":
            self._ensure_prec()
            return Matrix(self, "P")
        if kind == "a_robin":
            return self._robin_entries(form.index)
        b = torch.zeros(sp.N + sp.nreal, dtype=torch.float64, device=eng.device)
        if kind == "L":
            eng.rhs_prev(self.up_.x, b)
        elif kind == "L1":
            self._load_sources(form.index, b)
            self._load_robin_rhs(form.index, b)
        elif kind == "L0":
            f, s = self.problem.f, self.problem.s
            if isinstance(f, Constant):
                fv = f.values()
                for k in range(3):
                    if fv[k] != 0.0:
                        b[k * sp.N2:(k + 1) * sp.N2].add_(self._lumped_vec(2), alpha=float(fv[k]))
            elif self._cell_op_degree(f, 2) is not None:
                op = self._cell_op(2, self._cell_op_degree(f, 2))
                d = op.data(f)
                for k in range(3):
                    op.apply(eng, d[k], b[k * sp.N2:(k + 1) * sp.N2])
            else:
                vals = np.asarray(f.eval_points(sp.node2_coordinates()), dtype=float)
                for k in range(3):
                    vk = torch.as_tensor(np.ascontiguousarray(vals[:, k]), device=eng.device)
                    eng.mass_apply(2, 1.0, vk, b[k * sp.N2:(k + 1) * sp.N2])
            if not _is_zero(s):
                op = self._facet_op("m", None, NEUMANN_MARKER, True)
                if op.nf:
                    d = op.data(s)
                    for k in range(3):
                        op.apply(eng, d[k], b[k * sp.N2:(k + 1) * sp.N2])
        else:
            raise ValueError("unknown form %r" % (form,))
        return AssembledVector(b)

    # ------------------------------------------------------------------ time stepping
    def solve(self):
        """Solve to the end time T, yielding (up, t) at each step.  Users must set up_ first.
        The reference always takes its direct branch here (mpetsolver.py:308-309); so does this, unless
        direct_solver=False selects the MINRES/GMRES branch the reference left unreachable."""
        if self.params["direct_solver"]:
            return self.solve_direct()
        return self.solve_iterative()

    def _rhs(self, time, t0, dt, theta, bcs):
        """b for one step, in the reference's order (mpetsolver.py:424-453)."""
        t_theta = t0 + theta * dt
        t1 = t0 + dt
        b = assemble(self.L)
        time.assign(t_theta)
        for l in self.L1:
            b.axpy(1.0, assemble(l))
        time.assign(t1)
        b.axpy(1.0, assemble(self.L0))
        self._push_dirichlet_values(bcs)
        self.engine.apply_dirichlet_rhs(b.t)      # == for bc in bcs: bc.apply(b)
        return b

    def _assemble_system(self):
        A = assemble(self.a)
        for a_a in self.a_robin:
            A.axpy(1.0, assemble(a_a), False)
        return A

    def step(self, dt=None, up_=None):
        """One step t -> t + dt: re-assemble A, build b, solve (mpetsolver.py:317-379).  Updates up
        and time, not up_."""
        theta = self.params["theta"]
        time = self.problem.time
        if dt is None:
            dt = self.params["dt"]
        self.params["dt"] = float(dt)
        self.dt.assign(dt)
        if up_ is not None and up_ is not self.up_:
            self.up_.assign(up_)
        A = self._assemble_system()
        (bcs0, bcs1) = self.bcs
        bcs = bcs0 + bcs1
        for bc in bcs:
            bc.apply(A)
        self._sync_dirichlet(bcs)
        if self.params["direct_solver"]:
            solver = LUSolver(A, "mumps")
            krylov = solver.krylov
        else:       # the reference's unreachable iterative recipe (mpetsolver.py:507), reference tolerance
            self._ensure_prec()
            krylov = PETScKrylovSolver("minres" if self._exchange_is_symmetric() else "gmres", "hypre_amg")
            krylov.parameters.update(relative_tolerance=self.params["krylov_rtol"],
                                     absolute_tolerance=self.params["krylov_atol"],
                                     maximum_iterations=self.params["krylov_maxit"], nonzero_initial_guess=True)
        b = self._rhs(time, float(time), float(dt), theta, bcs)
        self._last_b = b                          # kept for residual checks (bench.py: true_residual)
        self.up.x.copy_(self.up_.x)               # initial guess; boundary entries are overwritten
        krylov.set_operators(A, None)
        niter = krylov.solve(self.up.vector(), b)
        self.solver_monitor.setdefault("niter", []).append(niter)
        self.solver_monitor["last"] = krylov.last_info

    def solve_direct(self):
        """Generator twin of mpetsolver.py:382-462 (A assembled once, one solve per step)."""
        dt = self.params["dt"]
        T = self.params["T"]
        theta = self.params["theta"]
        time = self.problem.time
        self.dt.assign(dt)
        [bcs0, bcs1] = self.create_dirichlet_bcs()
        bcs = bcs0 + bcs1
        A = self._assemble_system()
        for bc in bcs:
            bc.apply(A)
        self._sync_dirichlet(bcs)
        solver = LUSolver(A, "mumps")
        self.solver_monitor["niter"] = []
        while (float(time) < (T - 1.e-9)):
            b = self._rhs(time, float(time), float(dt), theta, bcs)
            self.up.x.copy_(self.up_.x)
            niter = solver.solve(A, self.up.vector(), b)
            self.solver_monitor["niter"] += [niter]
            self.solver_monitor["last"] = solver.krylov.last_info
            yield self.up, float(time)
            self.up_.assign(self.up)

    def solve_iterative(self):
        """Generator twin of mpetsolver.py:465-569: MINRES (GMRES for non-symmetric S) preconditioned
        by the block-diagonal AMG V-cycle built from the ``prec`` form, PETSc default tolerances."""
        dt = self.params["dt"]
        T = self.params["T"]
        theta = self.params["theta"]
        time = self.problem.time
        self.dt.assign(dt)
        [bcs0, bcs1] = self.create_dirichlet_bcs()
        bcs = bcs0 + bcs1
        A = self._assemble_system()
        P = assemble(self.prec)
        for bc in bcs:
            apply_symmetric(bc, P)
        self._sync_dirichlet(bcs)
        method = "minres" if self._exchange_is_symmetric() else "gmres"
        solver = PETScKrylovSolver(method, "hypre_amg")
        solver.parameters.update(relative_tolerance=self.params["krylov_rtol"],
                                 absolute_tolerance=self.params["krylov_atol"],
                                 maximum_iterations=self.params["krylov_maxit"],
                                 nonzero_initial_guess=True)
        self.solver_monitor["niter"] = []
        if self.params["testing"]:
            g = torch.Generator(device="cpu").manual_seed(0)
            self.up.x.copy_(torch.randn(self.up.x.numel(), generator=g, dtype=torch.float64))
        else:
            self.up.assign(self.up_)
        while (float(time) < (T - 1.e-9)):
            Acopy = A.copy()        # a handle: the engine keeps A free of boundary conditions
            b = self._rhs(time, float(time), float(dt), theta, bcs)
            for bc in bcs:
                apply_symmetric(bc, Acopy, b)
            solver.set_operators(Acopy, P)
            niter = solver.solve(self.up.vector(), b)
            self.solver_monitor["niter"] += [niter]
            self.solver_monitor["last"] = solver.last_info
            yield self.up, float(time)
            self.up_.assign(self.up)
        self.solver_monitor["P"] = P
        self.solver_monitor["A"] = Acopy

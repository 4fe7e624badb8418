"""MPETTotalPressureSolver on the B200 engine -- same surface as the reference's
MPETTotalPressureSolver (src/mpet/mpet/mpettotalpressuresolver.py:15-512): unknowns (u, p0, p_1..p_J)
with p0 the total pressure, attributes F, L0, L1, L2, prec, up_, up, methods create_dirichlet_bcs,
default_params, create_variational_forms, solve_direct, solve_iterative, solve.

The device side is the SAME library path as the standard solver: the mixed space is [P2]^3 x [P1]^(J+1)
and ``mpet_set_params_total_pressure`` (include/mpet_b200.h) fills the block-coefficient tables of the
generic Taylor-Hood system with the terms of F (mpettotalpressuresolver.py:267-274) and of the
block-diagonal preconditioner (:276-283).  The time loops below follow the reference line by line,
including its order of right-hand-side groups: L at t0 with up_, L1 and rhs(L2) at t0 + theta*dt,
L0 and the boundary values at t0 + dt.
"""
import numpy as np
import torch

from .dolfin_shim import Constant, Function, DirichletBC, Parameters
from .la import Form, Matrix, AssembledVector, assemble, LUSolver, PETScKrylovSolver, apply_symmetric
from .mpetsolver import MPETSolver, DIRICHLET_MARKER


def system(F):
    """``(a, L) = system(self.F)`` (mpettotalpressuresolver.py:343,428)."""
    return Form(F.solver, "a"), Form(F.solver, "L")


def lhs(form):
    """``lhs(L2i)`` (mpettotalpressuresolver.py:357): the Robin boundary mass added to A."""
    if form.kind == "F":
        return Form(form.solver, "a")
    assert form.kind == "L2"
    return Form(form.solver, "a_robin", form.index)


def rhs(form):
    """``rhs(L2i)`` (mpettotalpressuresolver.py:388)."""
    if form.kind == "F":
        return Form(form.solver, "L")
    assert form.kind == "L2"
    return Form(form.solver, "L2rhs", form.index)


class MPETTotalPressureSolver(MPETSolver):

    _first_network_sub = 2          # sub-space 1 is the total pressure (mpettotalpressuresolver.py:137)

    def __init__(self, problem, params=None, device=0, partition=None):
        "Create solver with given MPET problem and parameters."
        self.monitor = False
        super().__init__(problem, params, device=device, partition=partition)

    @staticmethod
    def default_params():
        "Define default solver parameters (mpettotalpressuresolver.py:146-162)."
        params = Parameters("MPETSolver")
        params.add("dt", 0.05)
        params.add("t", 0.0)
        params.add("T", 1.0)
        params.add("theta", 0.5)
        params.add("u_degree", 2)
        params.add("p_degree", 1)
        params.add("direct_solver", True)
        params.add("testing", False)
        # B200-path extras (not in the reference): Krylov controls of the iterative branch
        params.add("krylov_rtol", 1e-5)      # PETSc default [EXT]; the reference sets none
        params.add("krylov_atol", 1e-50)
        params.add("krylov_maxit", 10000)
        return params

    def _num_p1_fields(self):
        return int(self.problem.params["J"]) + 1

    def _engine_set_params(self, *vals):
        self.engine.set_params_total_pressure(*vals)

    def create_variational_forms(self, include_preconditioner=False):
        mesh = self.problem.mesh
        self.dt = Constant(self.params["dt"])
        J = int(self.problem.params["J"])
        VW = self.create_function_spaces(mesh)
        self.VQ = VW
        self.up_ = Function(VW)
        self.up = Function(VW)
        self.F = Form(self, "F")
        self.L0 = Form(self, "L0")
        self.L1 = [Form(self, "L1", i) for i in range(J)]
        self.L2 = [Form(self, "L2", i) for i in range(J)]
        self.prec = Form(self, "prec")
        # names of the standard solver, so that shared plumbing (and users of either class) find them
        self.a, self.L = system(self.F)
        self.a_robin = [lhs(l2) for l2 in self.L2]
        return self.F, self.L0, self.L1, self.L2, self.prec, self.up_, self.up

    def create_dirichlet_bcs(self):
        """mpettotalpressuresolver.py:124-143: no condition on the total pressure."""
        VP = self.up.function_space()
        bcs0 = [DirichletBC(VP.sub(0), self.problem.u_bar, self.problem.momentum_boundary_markers,
                            DIRICHLET_MARKER)]
        bcs1 = []
        for i in range(int(self.problem.params["J"])):
            bcs1 += [DirichletBC(VP.sub(i + 2), self.problem.p_bar[i],
                                 self.problem.continuity_boundary_markers[i], DIRICHLET_MARKER)]
        self.bcs = [bcs0, bcs1]
        return [bcs0, bcs1]

    # ------------------------------------------------------------------ assemble(form)
    def _assemble(self, form):
        kind = form.kind
        if kind == "L1":            # dt*g_i*w dx + dt*I_i*w ds(1)   (mpettotalpressuresolver.py:321)
            self._push_params()
            b = torch.zeros(self.VQ.N, dtype=torch.float64, device=self.engine.device)
            self._load_sources(form.index, b)
            return AssembledVector(b)
        if kind == "L2rhs":         # rhs of dt*beta_i*(-pm_i + p_robin_i)*w ds(2)   (:323)
            self._push_params()
            b = torch.zeros(self.VQ.N, dtype=torch.float64, device=self.engine.device)
            self._load_robin_rhs(form.index, b)
            return AssembledVector(b)
        return super()._assemble(form)

    def _rhs(self, time, t0, dt, theta, bcs):
        """b for one step in the reference's order (mpettotalpressuresolver.py:371-401)."""
        b = assemble(self.L)
        t_theta = t0 + theta * dt
        time.assign(t_theta)
        for L1i in self.L1:
            b.axpy(1.0, assemble(L1i))
        for L2i in self.L2:
            b.axpy(1.0, assemble(rhs(L2i)))
        t = float(time) + (1.0 - theta) * dt
        time.assign(t)
        b.axpy(1.0, assemble(self.L0))
        self._push_dirichlet_values(bcs)
        self.engine.apply_dirichlet_rhs(b.t)      # == for bc in bcs: bc.apply(b)
        return b, t

    def step(self, dt=None, up_=None):
        """One step t -> t + dt with re-assembly of A (the total-pressure class of the reference has no
        ``step``; provided with the semantics of MPETSolver.step, mpetsolver.py:317-379, so that the
        two formulations can be timed on the same unit of work)."""
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
            krylov = LUSolver(A, "mumps").krylov
        else:
            self._ensure_prec()
            krylov = PETScKrylovSolver("minres" if self._exchange_is_symmetric() else "gmres", "hypre_amg")
            krylov.parameters.update(relative_tolerance=self.params["krylov_rtol"],
                                     absolute_tolerance=self.params["krylov_atol"],
                                     maximum_iterations=self.params["krylov_maxit"], nonzero_initial_guess=True)
        b, _ = self._rhs(time, float(time), float(dt), theta, bcs)
        self._last_b = b
        self.up.x.copy_(self.up_.x)
        krylov.set_operators(A, None)
        niter = krylov.solve(self.up.vector(), b)
        self.solver_monitor.setdefault("niter", []).append(niter)
        self.solver_monitor["last"] = krylov.last_info

    def solve_direct(self):
        """Generator twin of mpettotalpressuresolver.py:330-411."""
        dt = self.params["dt"]
        T = self.params["T"]
        theta = self.params["theta"]
        time = self.problem.time
        self.dt.assign(dt)
        (a, L) = system(self.F)
        [bcs0, bcs1] = self.create_dirichlet_bcs()
        bcs = bcs0 + bcs1
        A = assemble(a)
        for L2i in self.L2:
            A2 = assemble(lhs(L2i))
            A.axpy(1.0, A2, False)
        for bc in bcs:
            bc.apply(A)
        self._sync_dirichlet(bcs)
        solver = LUSolver(A, "mumps")
        self.up.assign(self.up_)
        self.solver_monitor["niter"] = []
        while (float(time) < (T - 1.e-9)):
            b, t = self._rhs(time, float(time), float(dt), theta, bcs)
            self.up.x.copy_(self.up_.x)
            niter = solver.solve(A, self.up.vector(), b)
            self.solver_monitor["niter"] += [niter]
            self.solver_monitor["last"] = solver.krylov.last_info
            yield self.up, float(time)
            self.up_.assign(self.up)
            time.assign(t)

    def solve_iterative(self):
        """Generator twin of mpettotalpressuresolver.py:414-505: MINRES + block-diagonal AMG on
        ``prec = mu*(grad u, grad v) + sum_i (alpha_i^2/lambda + c_i + dt*theta*sum_j S_ij)(p_i, w_i)
        + dt*theta*K_i (grad p_i, grad w_i) + (p0, w0)`` (:276-283)."""
        dt = self.params["dt"]
        T = self.params["T"]
        theta = self.params["theta"]
        time = self.problem.time
        self.dt.assign(dt)
        (a, L) = system(self.F)
        [bcs0, bcs1] = self.create_dirichlet_bcs()
        bcs = bcs0 + bcs1
        A = assemble(a)
        for L2i in self.L2:
            A2 = assemble(lhs(L2i))
            A.axpy(1.0, A2, False)
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
        Acopy = A
        while (float(time) < (T - 1.e-9)):
            Acopy = A.copy()
            b, t = self._rhs(time, float(time), float(dt), theta, bcs)
            for bc in bcs:
                apply_symmetric(bc, Acopy, b)
            solver.set_operators(Acopy, P)
            niter = solver.solve(self.up.vector(), b)
            self.solver_monitor["niter"] += [niter]
            self.solver_monitor["last"] = solver.last_info
            yield self.up, float(time)
            self.up_.assign(self.up)
            time.assign(t)
        self.solver_monitor["P"] = P
        self.solver_monitor["A"] = Acopy

    def solve(self):
        """mpettotalpressuresolver.py:508-512."""
        if self.params["direct_solver"]:
            return self.solve_direct()
        return self.solve_iterative()
